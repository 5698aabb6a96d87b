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
// Task records live in an int32 pool:
//   ROOT:        slot, founder, stub_row, out_idx, nD, up_m_row[nD]
//   FAMILY_UP:   flags, up_m_row, n_down, { role, slot, founder, up_i_row, nD, up_m_row[nD] } * n_down
//   FAMILY_DOWN: flags, prong_row, u_slot, u_founder, u_dn_row (-1 = root), nSib, sib_up_m_row[nSib],
//                n_down, { role, up_i_row, stub_row, dn_row (-1 = not needed) } * n_down
// flags = selfing | (role of u(m) in m) << 1, roles: 0 father, 1 mother, 2 kid.
enum TaskKind : int { TASK_ROOT = 0, TASK_FAMILY_UP = 1, TASK_FAMILY_DOWN = 2 };

struct SweepTask { int kind; int off; };

struct SweepParams {
    DevView D;
    const SweepTask *tasks;
    const int *pool;
    const int *level_off;   // [n_groups][n_levels + 1] offsets into tasks
    int n_groups, n_levels;
    double *partial;        // [n_tiles][n_out] per-tile log Z sums
    int n_out;
};

constexpr int SWEEP_WARPS = 8;
constexpr int TILE_LOCI = 32;

// ---------------------------------------------------------------------------------------------
// Eval plan. A group = up to EVAL_G proposals of the same focal kid, handled by one warp over
// one chunk of loci (lanes = consecutive loci, so every row read is a coalesced 256 B segment).
//   TRIO:    v = sum_gm xm[gm] sum_gp A_kid[gm][gp] xp[gp]      (_CalculateLikeByStubs@0x10001f342)
//   SELFING: v = sum_gk kid[gk] sum_g T[gk][g][g] xp[g]         (_CalculateLikeByStubsSelfing@0x10001f7ab)
//   PRONG:   v = sum_g prong[g] kid[g]                          (_CalculateLikeByProngs@0x10001fb44)
// partial[chunk][proposal] = log of the product of v over the chunk's loci.
constexpr int EVAL_G = 4;
constexpr int EVAL_WARPS = 4;

struct EvalProp { int kind; int row_p; int row_m; int pad; };
struct EvalGroup { int kid_row; int begin; int n; int pad; };

struct EvalParams {
    DevView D;
    const EvalGroup *groups;
    const EvalProp *props;
    int n_groups, n_props;
    int chunk_loci;         // multiple of 32
    double *partial;        // [n_chunks][n_props]
};

struct ReduceParams {
    const double *partial;  // [n_parts][n]
    double *out;            // [n]
    int n_parts, n;
};

void launch_sweep(const SweepParams &P, int n_tiles, cudaStream_t st);
void launch_eval(const EvalParams &P, int n_chunks, cudaStream_t st);
void launch_reduce(const ReduceParams &P, cudaStream_t st);

} // namespace pf
