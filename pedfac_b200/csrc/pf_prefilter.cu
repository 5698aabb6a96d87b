// pf_prefilter.cu — the N_kid x N_cand x S single-parent scan as a register-tiled streaming
// kernel. For kid k and candidate c:  score = sum_s log( sum_g C_k[s][g] * X_c[s][g] ) + bias_c
// with C_k[s][g] = sum_h A_k[s][g][h] * prior[s][h] (kid stub contracted with _TPROB and the HWE
// prior of the unknown other parent) and X_c the candidate's stub. Shaped like a GEMM over the
// (locus, state) axis but the combine is a product (log-sum), so it runs on the fp64 pipe:
// tiles of C and X are staged in shared memory, every thread keeps a 2 x 4 block of (kid, cand)
// running products in registers, and the only HBM traffic is one pass over the stub rows per tile.
#include "pf_prefilter.cuh"

namespace pf {

constexpr int PF_KT = 32;   // kids per block
constexpr int PF_CT = 64;   // candidates per block
constexpr int PF_LT = 16;   // loci per stage
constexpr int PF_RK = 2, PF_RC = 4;

__global__ void __launch_bounds__(256) prefilter_scores_kernel(const __grid_constant__ PrefilterParams P)
{
    const DevView &D = P.D;
    __shared__ double Cs[3][PF_LT][PF_KT];
    __shared__ double Xs[3][PF_LT][PF_CT + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int kid0 = blockIdx.y * PF_KT, cand0 = blockIdx.x * PF_CT;
    const double *prior = D.rows + (size_t)ROW_PRIOR * 3 * D.Sp;

    LogProd acc[PF_RK][PF_RC];
#pragma unroll
    for (int r = 0; r < PF_RK; r++)
#pragma unroll
        for (int c = 0; c < PF_RC; c++) lp_init(acc[r][c]);

    for (int s0 = 0; s0 < D.Sl; s0 += PF_LT) {
        // stage C: 32 kids x 16 loci, two items per thread (lanes along the locus axis)
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int idx = tid + 256 * it, l = idx & 15, k = idx >> 4;
            const int s = s0 + l, kid = kid0 + k;
            Tri c{1.0, 0.0, 0.0};
            if (s < D.Sl && kid < P.n_kids) {
                const Tri ks = ld_tri(row_ptr(D, P.kid_rows[kid]), D.Sp, s);
                c = sym_mv(kid_matrix(ks), ld_tri(prior, D.Sp, s));
            }
            Cs[0][l][k] = c.x; Cs[1][l][k] = c.y; Cs[2][l][k] = c.z;
        }
        // stage X: 64 candidates x 16 loci x 3 states, 12 doubles per thread
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int idx = tid + 256 * it, l = idx & 15, c = idx >> 4;
            const int s = s0 + l, cand = cand0 + c;
            Tri x{1.0, 0.0, 0.0};
            if (s < D.Sl && cand < P.n_cands) x = ld_tri(row_ptr(D, P.cand_rows[cand]), D.Sp, s);
            Xs[0][l][c] = x.x; Xs[1][l][c] = x.y; Xs[2][l][c] = x.z;
        }
        __syncthreads();
#pragma unroll 4
        for (int l = 0; l < PF_LT; l++) {
            Tri ck[PF_RK], xc[PF_RC];
#pragma unroll
            for (int r = 0; r < PF_RK; r++) ck[r] = Tri{Cs[0][l][ty + 16 * r], Cs[1][l][ty + 16 * r], Cs[2][l][ty + 16 * r]};
#pragma unroll
            for (int c = 0; c < PF_RC; c++) xc[c] = Tri{Xs[0][l][tx + 16 * c], Xs[1][l][tx + 16 * c], Xs[2][l][tx + 16 * c]};
#pragma unroll
            for (int r = 0; r < PF_RK; r++)
#pragma unroll
                for (int c = 0; c < PF_RC; c++) lp_mul(acc[r][c], dot(ck[r], xc[c]));
        }
        __syncthreads();
        // keep the running mantissas in range: at most 16 factors in [1,2) per stage
        if (((s0 / PF_LT) & 31) == 31) {
#pragma unroll
            for (int r = 0; r < PF_RK; r++)
#pragma unroll
                for (int c = 0; c < PF_RC; c++) lp_renorm(acc[r][c]);
        }
    }
#pragma unroll
    for (int r = 0; r < PF_RK; r++)
#pragma unroll
        for (int c = 0; c < PF_RC; c++) {
            const int kid = kid0 + ty + 16 * r, cand = cand0 + tx + 16 * c;
            if (kid < P.n_kids && cand < P.n_cands)
                P.scores[(size_t)kid * P.n_cands + cand] = lp_log(acc[r][c]) + P.cand_bias[cand];
        }
}

// One warp per kid: top_k rounds of (max score, earliest candidate) selection; chosen entries
// are overwritten with NaN. A candidate that is the kid itself is never eligible.
__global__ void __launch_bounds__(128) prefilter_topk_kernel(const __grid_constant__ PrefilterParams P)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= P.n_kids) return;
    double *sc = P.scores + (size_t)warp * P.n_cands;
    const int kid_row = P.kid_rows[warp];
    for (int j = lane; j < P.n_cands; j += 32)
        if (P.cand_rows[j] == kid_row) sc[j] = __longlong_as_double(0x7ff8000000000000LL);
    __syncwarp();
    for (int t = 0; t < P.top_k; t++) {
        double best = 0.0; int bj = -1;
        for (int j = lane; j < P.n_cands; j += 32) {
            const double v = sc[j];
            if (v != v) continue;
            if (bj < 0 || v > best) { best = v; bj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (oj >= 0 && (bj < 0 || ob > best || (ob == best && oj < bj))) { best = ob; bj = oj; }
        }
        if (lane == 0) {
            P.out_idx[(size_t)warp * P.top_k + t] = bj;
            P.out_score[(size_t)warp * P.top_k + t] = bj >= 0 ? best : 0.0;
            if (bj >= 0) sc[bj] = __longlong_as_double(0x7ff8000000000000LL);
        }
        __syncwarp();
    }
}

int launch_prefilter_scores(const PrefilterParams &P, cudaStream_t st)
{
    if (P.n_kids <= 0 || P.n_cands <= 0) return 0;
    dim3 grid((P.n_cands + PF_CT - 1) / PF_CT, (P.n_kids + PF_KT - 1) / PF_KT);
    prefilter_scores_kernel<<<grid, 256, 0, st>>>(P);
    return 1;
}

int launch_prefilter_topk(const PrefilterParams &P, cudaStream_t st)
{
    if (P.n_kids <= 0) return 0;
    prefilter_topk_kernel<<<(P.n_kids * 32 + 127) / 128, 128, 0, st>>>(P);
    return 1;
}

} // namespace pf
