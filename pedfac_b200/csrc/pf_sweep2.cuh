// pf_sweep2.cuh — plan format of the second-generation sum-product sweep (sweep2_kernel).
//
// Same mathematics and the same level-synchronous schedule as pf_kernels.cuh's sweep (one block =
// one 32-locus tile of one task group, warps share the tasks of a level, lanes = loci), but the plan
// is *compiled* on the host instead of interpreted on the device:
//   * every operand of a task is pre-resolved to a 16-bit shared-memory row number (row 0 is a
//     constant all-ones row, so an absent factor costs a multiply instead of a branch); persistent
//     rows (stubs, prongs) are only ever stored and carry their 32-bit global row number;
//   * an individual's up message is never stored: wherever it is needed it is re-composed in
//     registers as  decode(genotype tile) x prior? x up_m(row0) x up_m(row1) x up_m(row2)  — a
//     "cmp" operand of three words. That removes the up_i rows, the leaf cache rows and the LEAF_IN /
//     KFOLD level of the first sweep: a component needs one row per marriage plus two small banks;
//   * a family task multiplies up to four kids in line; only larger sibships use chunk tasks
//     (KFOLD2 at the up band; KIDCHUNK2 at the level of the marriage's FAMDOWN2 and independent of it: it
//     re-composes the parents' messages itself, so nothing is parked between the two);
//   * a block's warps are alone on their schedulers, so a sweep is bound by instruction issue and
//     instruction fetch (ncu: stall_no_inst first, then stall_wait): task bodies are straight-line
//     code, one instance per number of kids, whose only branches are warp-uniform tests of the
//     record header (branches inside the operands, to skip the rows of a leaf or the prior of a
//     non-founder, were measured slower than the multiplications they save);
//   * no local memory at all: shared memory takes the whole L1, so a spill or an indexed local array
//     is an L2 round trip per access (measured: it doubled the time of every level).
// A launch whose groups do not fit in shared memory (rows + genotype tiles + plan slice) falls back
// to the first-generation kernel, which can spill rows to the L2-resident scratch array.
//
// Reference routines restated by the tasks: FAMUP2 = _PassingIVtoIF@0x10001c490 +
// _PassingIFtoIV@0x10001c8f0 towards the root; FAMDOWN2 / KIDCHUNK2 = the same two away from the
// root + _FillInProngs@0x10001dd70 + _FillInStubs@0x10001e1a0; ROOT2 = _UpdateProb@0x10001da20.
#pragma once
#include "pf_kernels.cuh"

namespace pf {

enum Task2Kind : uint32_t { T2_ROOT = 0, T2_FAMUP = 1, T2_FAMDOWN = 2, T2_KFOLD = 3, T2_KIDCHUNK = 4 };

// cmp operand, 3 words:
//   w0 = where the individual's genotype data is | founder << 30
//        launches compiled for one genotype format (hard calls / 16-bit numerators over a common
//        denominator): the byte offset of the individual's staged tile (0 = the constant "no data"
//        tile); mixed launches: kind (3 bits, EmisKind) | genotype entry << 3
//   w1 = row0 | row1 << 16        (absent rows: row 0)
//   w2 = row2 | extra << 16       (extra != 0: word offset, from the group's pool, of {n, row...} more
//                                  rows; the record's header then carries T2_GENERAL)
constexpr uint32_t CMP_FOUNDER = 1u << 30, CMP_OFF_MASK = (1u << 30) - 1u;
constexpr int CMP_WORDS = 3;
constexpr uint32_t ROW2_ONES = 0;            // shared-memory row 0: all ones
constexpr uint32_t ROW2_TRASH = 1;           // shared-memory row 1: written instead of "no row", never read
// header flags (bits 8..15 of w0): selfing, urole (2 bits), then
constexpr uint32_t T2_HAS_A = 1u << 11, T2_HAS_B = 1u << 12;
// T2_GENERAL: a selfing marriage, or some operand of the record names more than three rows
constexpr uint32_t T2_GENERAL = 1u << 15;
constexpr int FAM2_INLINE_KIDS = 4;          // kids a family task handles itself
constexpr int CHUNK2_KIDS = 3;               // kids per chunk of a larger sibship

// Record layouts (32-bit words). header w0 = kind | flags << 8 | nk << 16 | np << 24
//   flags = selfing | urole << 1 (role of the marriage's up individual: 0 father, 1 mother, 2 kid)
//   ROOT2     w0 (flags bit 0: store the stub), w1 stub persistent row, w2 output index,
//             w3 transient destination row (up-only sweeps), cmp
//   FAMUP2    w0, w1 destination row, cmp A, cmp B, nk x cmp, np pair rows (two per word)
//   FAMDOWN2  w0, w1 prong persistent row, cmp U (the up individual without this marriage, with the
//             message from above among its rows), cmp A, w8 A stub row, cmp B, w12 B stub row
//             (T2_HAS_A / T2_HAS_B: that parent is a down member), w13 = A dn row | B dn row << 16,
//             nk x { cmp, stub persistent row, dn row }, np pair rows
//   KFOLD2    w0 (nk kids), w1 destination pair row, nk x cmp
//   KIDCHUNK2 w0 (nk own kids, np pairs of the sibship), w1 = ni | own chunk << 8, cmp U, cmp A,
//             cmp B, nk x { cmp, stub persistent row, dn row }, ni x cmp (the kids FAMDOWN2 handles),
//             np pair rows. Same level as the marriage's FAMDOWN2, independent of it.
struct Sweep2Task { uint32_t off; };          // word offset of the record in the group's pool

struct Emis2 {            // genotype entry of one individual of a group (16 bytes: one LDS.128)
    uint32_t smem_off;    // byte offset of its staged 32-locus tile
    uint32_t denom;       // EMIS_DEC16 denominator
    double rcp;           // RN(1 / denom)
};
struct Stage2 {           // what to copy into shared memory for one individual (read from global memory)
    unsigned long long ptr;
    uint32_t kind, smem_off;
};

struct Sweep2Group {
    int task_begin, task_end;     // sorted task range (Sweep2Task)
    int pool_begin, pool_len;     // record words
    int emis_begin, emis_count;   // Emis2 / Stage2 entries
    int emis_tab;                 // Emis2 entries the kernel needs in shared memory (generic format: all; else 0)
    int emis_bytes;               // staged genotype bytes
    int n_rows;                   // shared-memory rows, rows 0 (ones) and 1 (trash) included
};

struct Sweep2Params {
    DevView D;
    const Sweep2Task *tasks;
    const uint32_t *pool;
    const int *level_off;         // [n_groups][n_levels + 1] offsets into tasks
    const Emis2 *emis;
    const Stage2 *stage;
    const Sweep2Group *groups;
    int n_groups, n_levels;
    double *partial;              // [n_tiles][n_out]
    int n_out;
    long long *clocks;
    int clock_group;
    double *out;                  // fused reduction of the tiles (small n_out), else null
    unsigned *counter;
    int debug_flags;              // measurement aid: bit 0 = skip the persistent stores
    int format;                   // Fmt2 of the launch: 0 generic, 1 hard calls, 2 dec16 with one denominator
    uint32_t dec16_denom;
    double dec16_rcp;
};

struct Sweep2Smem { int loff, tasks, pool, emis, edata, rows, total; };
__host__ __device__ inline Sweep2Smem sweep2_smem_layout(const Sweep2Group &g, int n_levels)
{
    Sweep2Smem L;
    int o = 0;
    L.loff = o; o += ((n_levels + 1) * 4 + 15) & ~15;
    L.tasks = o; o += ((g.task_end - g.task_begin) * 4 + 15) & ~15;
    L.pool = o; o += (g.pool_len * 4 + 15) & ~15;
    L.emis = o; o += g.emis_tab * (int)sizeof(Emis2);
    L.edata = o; o += (g.emis_bytes + 15) & ~15;
    L.rows = o; o += g.n_rows * SMEM_ROW_BYTES;
    L.total = o;
    return L;
}

#ifndef PF_SWEEP2_WARPS
#define PF_SWEEP2_WARPS 8
#endif
constexpr int SWEEP2_WARPS = PF_SWEEP2_WARPS, SWEEP2_WARPS_SMALL = 4;
constexpr int SWEEP2_SMEM_MAX = 224 * 1024;      // one block per SM
constexpr int SWEEP2_SMEM_SMALL = 72 * 1024;     // three blocks per SM

void launch_sweep2(const Sweep2Params &P, int n_tiles, int smem_bytes, cudaStream_t st);

} // namespace pf
