// pf_kernels.cuh — plan formats shared by the host-side plan builder and the sm_100a kernels.
#pragma once
#include "pf_device.cuh"

namespace pf {

// ---------------------------------------------------------------------------------------------
// Sweep plan. A dirty component is a tree over individuals and marriages, rooted at one
// individual. Per marriage m with up-individual u(m) (towards the root) two tasks exist:
//   FAMILY_UP(m):   for every down member x: up_i(x) = local(x) * prod_{m' below x} up_m(m')
//                   (stored), then up_m(m) = message m -> u(m)            [_PassingIVtoIF@0x10001c490
//                                                                          + _PassingIFtoIV@0x10001c8f0]
//   FAMILY_DOWN(m): in_u = local(u) * dn_m(u) * prod_{other m' below u} up_m(m'); for every down
//                   member x: dn_m(x) = message m -> x, stub(x) = up_i(x) * dn_m(x)   [_FillInStubs@0x10001e1a0];
//                   prong(m) [_FillInProngs@0x10001dd70]
//   ROOT(r):        stub(r) = local(r) * prod up_m; Z = sum_g stub(r)[g]; log Z summed over loci
//                   is the component's LL [_UpdateProb@0x10001da20]
// Tasks are sorted by (group, level); a thread block owns one 32-locus tile of one group and
// walks the levels with __syncthreads() between them; its warps share the tasks of a level.
// Independent components are spread over groups so large sweeps fill the GPU.
//
// A block first stages, in shared memory, its group's slice of the plan and the genotype data
// of every individual of the group for its 32 loci (one fully parallel load phase); transient
// message rows live in shared memory too (as many as fit; the rest spill to an L2-resident
// scratch array). The level loop therefore runs on shared-memory latency; the only global
// traffic inside it is the streaming store of stubs and prongs.
//
// One task per family and direction keeps the level count at (tree radius) x 2 and lets a warp
// load every member's message once; a task is executed by ONE warp per 32-locus tile, the other
// warps of the block take the other families of the level:
//   LEAF_IN(x)   cache a leaf's local factor in a shared-memory row             level 0
//   FAM_UP(m)    up_i of the non-leaf members (stored), then up_m(m)             level 1 + height
//   ROOT(r)      stub(r), log Z                                                  first level of the down band
//   FAM_DOWN(m)  in_u(m); dn_m(x) and stub(x) = up_i(x) * dn_m(x) for every down
//                member x; prong(m)                                              level base + depth
// Task records live in an int32 pool. Row references: plain = persistent row index; bit 30 =
// transient row (group-local index: shared memory below the group's n_rows_smem, else spilled);
// bit 31 (LEAF_REF_BIT) = no row: recompute the leaf's local factor from its genotype tile.
// A sibship is a list of items {type, value}: type 0 = one kid (value = its up reference), type 1 =
// a chunk of 2..FAM_CHUNK leaf kids whose factor product a KFOLD task left in rows value, value + 1.
//   ROOT:      member, stub_row, out_idx, nD, up_m_row[nD]
//   LEAF_IN:   member, up_i_row
//   KFOLD:     out_row (2 rows), n, kid_ref[n]
//   FAM_UP:    flags, up_m_row, nC, { member, up_i_row, nD, below_up_m_row[nD] } * nC, pa_ref, ma_ref,
//              n_items, item[n_items]
//   FAM_DOWN:  flags, prong_row, u_member, u_dn_row (-1 = root), nSib, sib_up_m_row[nSib],
//              pa{ref, stub_row, dn_row}, ma{ref, stub_row, dn_row}, n_items, tmp_row, item[n_items]
//              (+ stub_row, dn_row of the kid when the sibship is one single kid: tmp_row = -1;
//              ref = NO_REF: that parent is the up individual / absent; dn_row = -1: not needed)
//   FAM_EXCL:  n_items (>= 2), out_row, item[n_items]
//   KID_CHUNK: flags, ab_row (a, b, Pu: 4 rows), out_row (-1 = the only item), n, { up_ref, stub_row, dn_row } * n
// member = role | founder << 2 | (group-local genotype entry index) << 3; roles: 0 father,
// 1 mother, 2 kid. flags = selfing | (role of u(m) in m) << 1.
// (This is the first-generation plan; the compiled records of sweep2_kernel are in pf_sweep2.cuh.)
enum TaskKind : int { TASK_ROOT = 0, TASK_LEAF_IN = 1, TASK_FAM_UP = 2, TASK_FAM_DOWN = 3, TASK_KID_CHUNK = 4, TASK_KFOLD = 5, TASK_FAM_EXCL = 6 };
constexpr int NO_REF = 0x1fffffff;
#ifndef PF_FAM_CHUNK
#define PF_FAM_CHUNK 3
#endif
#ifndef PF_FAM_SMALL_MAX
#define PF_FAM_SMALL_MAX 1
#endif
constexpr int FAM_CHUNK = PF_FAM_CHUNK;           // leaf kids per sibship chunk (one KFOLD and one KID_CHUNK task each)
constexpr int FAM_SMALL_MAX = PF_FAM_SMALL_MAX;   // sibships up to this size are finished inside FAM_DOWN
constexpr int SMEM_ROW_BIT = 1 << 29;
// in place of an up_i row reference: the member is a leaf (no marriage below it), its up message
// is its local factor, recomputed from the staged genotype tile: LEAF_REF_BIT | member
constexpr unsigned LEAF_REF_BIT = 0x80000000u;

struct SweepTask { int kind; int off; };

// genotype entry of one individual of a group
struct EmisEnt {
    unsigned kind;        // EmisKind
    unsigned denom;       // EMIS_DEC16 denominator
    unsigned ptr_lo, ptr_hi;   // device pointer of the slot's genotype row
    unsigned smem_off;    // byte offset of the staged 32-locus tile, or EMIS_NOT_STAGED
    unsigned pad;
    double rcp;           // RN(1 / denom): num / denom is evaluated exactly as q = num * rcp;
                          // q += fma(-denom, q, num) * rcp (verified exhaustively on the host)
};
constexpr unsigned EMIS_NOT_STAGED = 0xffffffffu;

struct SweepGroup {
    int task_begin, task_end;     // sorted task range
    int pool_begin, pool_len;     // slice of the int pool
    int emis_begin, emis_count;   // slice of the genotype entries
    int emis_bytes;               // staged genotype bytes
    int n_rows, n_rows_smem;      // transient rows, and how many of them live in shared memory
    int scratch_base;             // first global scratch row of the group
    int staged;                   // plan slice staged in shared memory?
    int pad;
};

struct SweepParams {
    DevView D;
    const SweepTask *tasks;
    const int *pool;
    const int *level_off;   // [n_groups][n_levels + 1] offsets into tasks
    const EmisEnt *emis;
    const SweepGroup *groups;
    int n_groups, n_levels;
    double *partial;        // [n_tiles][n_out] per-tile log Z sums
    int n_out;
    long long *clocks;      // measurement aid: per-level clock64() of block (0, largest group), or null
    int clock_group;
    double *out;            // fused reduction (small n_out): the last block to finish adds the tiles, else null
    unsigned *counter;      // arrival counter of the fused reduction (zero between launches)
};

constexpr int TILE_LOCI = 32;
constexpr int SMEM_ROW_BYTES = 3 * TILE_LOCI * 8;

// The sweep kernel is built in two shapes. A block is latency-bound (its levels are a chain of
// single-warp tasks), so what matters is how many blocks an SM can overlap:
//   latency shape    16 warps, 1 block per SM, 195 KB of shared memory: the fastest single block;
//                    used while the launch's heavy blocks fit on the SMs in one wave (a Gibbs-step
//                    sweep of a 2 000-locus shard)
//   throughput shape  8 warps, 2 blocks per SM, 98 KB each: a block is ~1.4x slower (more message
//                    rows spill to L2, fewer warps for the wide levels) but two run side by side;
//                    used when there are more heavy blocks than SMs (10 000 loci, many chains)
// The budgets keep the total under the 196 KB carve-out so that 32 KB of L1 remain.
struct SweepShape { int warps, blocks_per_sm, smem_max, plan_smem_max, emis_smem_max, leaf_row_budget; };
#ifndef PF_LAT_WARPS
#define PF_LAT_WARPS 16
#endif
#ifndef PF_LEAF_BUDGET
#define PF_LEAF_BUDGET 400
#endif
constexpr SweepShape SWEEP_LATENCY{PF_LAT_WARPS, 1, 195 * 1024, 56 * 1024, 72 * 1024, PF_LEAF_BUDGET};
#ifndef PF_T_LEAF
#define PF_T_LEAF 64
#endif
constexpr SweepShape SWEEP_THROUGHPUT{8, 2, 98 * 1024, 28 * 1024, 36 * 1024, PF_T_LEAF};

// staged bytes of one individual's 32-locus genotype tile
__host__ __device__ inline int emis_tile_bytes(unsigned kind)
{
    return kind == EMIS_HARD ? 8 : kind == EMIS_DEC16 ? 3 * TILE_LOCI * 2 : kind == EMIS_F32 ? 3 * TILE_LOCI * 4
         : kind == EMIS_F64 ? 3 * TILE_LOCI * 8 : 0;
}

// shared-memory layout of a sweep block (same arithmetic on host and device)
struct SweepSmem { int tasks, pool, loff, emis, prior, eps, edata, rows, total; };
__host__ __device__ inline SweepSmem sweep_smem_layout(const SweepGroup &g, int n_levels)
{
    SweepSmem L;
    int o = 0;
    L.prior = o; o += 3 * TILE_LOCI * 8;
    L.eps = o; o += TILE_LOCI * 8;
    L.loff = o; o += ((n_levels + 1) * 4 + 15) & ~15;
    L.tasks = o; if (g.staged) o += ((g.task_end - g.task_begin) * 8 + 15) & ~15;
    L.pool = o; if (g.staged) o += (g.pool_len * 4 + 15) & ~15;
    L.emis = o; if (g.staged) o += (g.emis_count * (int)sizeof(EmisEnt) + 15) & ~15;
    L.edata = o; o += (g.emis_bytes + 15) & ~15;
    L.rows = o; o += g.n_rows_smem * SMEM_ROW_BYTES;
    L.total = o;
    return L;
}

// ---------------------------------------------------------------------------------------------
// Eval plan. A group = up to EVAL_G proposals of the same focal kid, handled by one warp over
// one chunk of loci (lanes = consecutive loci, so every row read is a coalesced 256 B segment).
//   TRIO:    v = sum_gm xm[gm] sum_gp A_kid[gm][gp] xp[gp]      (_CalculateLikeByStubs@0x10001f342)
//   SELFING: v = sum_gk kid[gk] sum_g T[gk][g][g] xp[g]         (_CalculateLikeByStubsSelfing@0x10001f7ab)
//   PRONG:   v = sum_g prong[g] kid[g]                          (_CalculateLikeByProngs@0x10001fb44)
// partial[chunk][proposal] = log of the product of v over the chunk's loci.
constexpr int EVAL_G = 4;        // general proposals per warp
#ifndef PF_EVAL_GS
#define PF_EVAL_GS 4
#endif
constexpr int EVAL_GS = PF_EVAL_GS;   // single-parent proposals per warp (one row and three FMAs each, two loci per lane)
constexpr int EVAL_WARPS = 4;

struct EvalProp { int kind; int row_p; int row_m; int out; };    // out: index of the proposal in the caller's list
struct EvalGroup { int kid_row; int begin; int n; int single; };

struct EvalParams {
    DevView D;
    const EvalGroup *groups;
    const EvalProp *props;
    int n_groups, n_props;
    int chunk_loci;         // multiple of 32
    double *partial;        // [n_chunks][n_props]
    double *out;            // [n_props] sums over chunks (written by the last block of each column)
    unsigned *counters;     // [group blocks] arrival counters, zero between launches
    // optional fused result delivery: the block that finishes the last column copies the call's whole
    // result vector (pub_n doubles at pub_src: a preceding sweep's ll_cc and this launch's sums) to the
    // host's mapped mailbox and raises its sequence flag — no publish launch
    const double *pub_src; double *pub_dst; int pub_n;
    const double *pub_src2; int pub_n2;            // a second piece, appended (deferred swap results)
    unsigned long long *pub_flag, pub_seq;
    unsigned *pub_counter;
};

struct ReduceParams {
    const double *partial;  // [n_parts][n]
    double *out;            // [n]
    int n_parts, n;
};

// cross-GPU sum over peer memory (xreduce_kernel)
constexpr int XREDUCE_MAX = 8192;        // doubles per (parity, rank) mailbox slot
constexpr int XREDUCE_MAX_WORLD = 8;
struct XReduceParams {
    const double *src;                   // this rank's partial sums [n]
    double *dst;                         // [n] sums over all ranks (may alias src)
    int n, rank, world;
    unsigned long long seq0;             // sequence number of the first round (n / XREDUCE_MAX rounds follow)
    double *peer_mail[XREDUCE_MAX_WORLD];            // every rank's mailbox data [2][world][XREDUCE_MAX]
    unsigned long long *peer_flags[XREDUCE_MAX_WORLD];   // every rank's flags [world]
    double *host_dst;                    // mapped pinned result vector, or null
    unsigned long long *host_flag, host_seq;
};
void launch_xreduce(const XReduceParams &X, cudaStream_t st);

void launch_sweep(const SweepParams &P, int n_tiles, int smem_bytes, bool staged, bool throughput_shape, cudaStream_t st);
void launch_eval(const EvalParams &P, int n_chunks, cudaStream_t st);
int eval_resident_blocks();
constexpr int EVAL_INLINE_MAX = 28672;
void set_eval_pdl(bool on);
// the plan (groups, then proposals at off_props) inside the kernel's parameters; false: too large, nothing launched
bool launch_eval_inline(const EvalParams &P, const void *blob, size_t bytes, size_t off_props, int n_chunks, cudaStream_t st);
void launch_reduce(const ReduceParams &P, cudaStream_t st);
void launch_publish(const double *src, double *dst_mapped, int n, unsigned long long *flag_mapped, unsigned long long seq, cudaStream_t st);

} // namespace pf
