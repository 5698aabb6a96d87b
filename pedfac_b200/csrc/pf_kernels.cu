// pf_kernels.cu — hand-written sm_100a kernels of libpedfac_cuda: the level-synchronous
// sum-product sweep, the batched proposal evaluation and the fixed-order partial reduction.
// All arithmetic is fp64 (3-state contractions: no tensor-core shape exists here); loci are the
// contiguous axis of every row so warps issue full-sector coalesced accesses.
#include "pf_kernels.cuh"

namespace pf {

// ---- sweep ----------------------------------------------------------------------------------

__device__ __forceinline__ Tri prod_rows(const DevView &D, const int *rows, int n, int s, Tri acc)
{
    for (int j = 0; j < n; j++) acc = mul(acc, ld_tri(row_ptr(D, rows[j]), D.Sp, s));
    return acc;
}

// fold one member's incoming message into the family tensor
__device__ __forceinline__ void fold(Mat3 &W, int role, Tri in)
{
    if (role == 0) mat_mul_pa(W, in);
    else if (role == 1) mat_mul_ma(W, in);
    else mat_mul_kid(W, in);
}
__device__ __forceinline__ Tri emit(const Mat3 &W, int role)
{
    return role == 0 ? mat_to_pa(W) : role == 1 ? mat_to_ma(W) : mat_to_kid(W);
}
__device__ __forceinline__ void fold_self(Tri &w, int role, Tri in)
{
    w = mul(w, role == 0 ? in : self_kid_factor(in));
}
__device__ __forceinline__ Tri emit_self(Tri w, int role) { return role == 0 ? w : self_to_kid(w); }

__device__ __forceinline__ void ones9(Mat3 &W)
{
#pragma unroll
    for (int i = 0; i < 9; i++) W.m[i] = 1.0;
}

__device__ void task_root(const SweepParams &P, const int *t, int s, bool valid, int tile, int lane)
{
    const DevView &D = P.D;
    const int slot = t[0], founder = t[1], stub_row = t[2], out_idx = t[3], nD = t[4];
    LogProd z; lp_init(z);
    if (valid) {
        Tri v = prod_rows(D, t + 5, nD, s, local_factor(D, slot, founder, s));
        st_tri(row_ptr(D, stub_row), D.Sp, s, v);
        lp_mul(z, (v.x + v.y) + v.z);
    }
    if (out_idx >= 0) {
        z = lp_warp_reduce(z);
        if (lane == 0) P.partial[(size_t)tile * P.n_out + out_idx] = lp_log(z);
    }
}

__device__ void task_family_up(const SweepParams &P, const int *t, int s, bool valid)
{
    if (!valid) return;
    const DevView &D = P.D;
    const int flags = t[0], up_m_row = t[1], n_down = t[2];
    const int selfing = flags & 1, urole = flags >> 1;
    const int *q = t + 3;
    Mat3 W; Tri w{1.0, 1.0, 1.0};
    ones9(W);
    for (int x = 0; x < n_down; x++) {
        const int role = q[0], slot = q[1], founder = q[2], up_i_row = q[3], nD = q[4];
        Tri in = prod_rows(D, q + 5, nD, s, local_factor(D, slot, founder, s));
        st_tri(row_ptr(D, up_i_row), D.Sp, s, in);
        if (selfing) fold_self(w, role, in); else fold(W, role, in);
        q += 5 + nD;
    }
    st_tri(row_ptr(D, up_m_row), D.Sp, s, selfing ? emit_self(w, urole) : emit(W, urole));
}

__device__ void task_family_down(const SweepParams &P, const int *t, int s, bool valid)
{
    if (!valid) return;
    const DevView &D = P.D;
    const int flags = t[0], prong_row = t[1], u_slot = t[2], u_founder = t[3], u_dn_row = t[4], nSib = t[5];
    const int selfing = flags & 1, urole = flags >> 1;
    Tri in_u = local_factor(D, u_slot, u_founder, s);
    if (u_dn_row >= 0) in_u = mul(in_u, ld_tri(row_ptr(D, u_dn_row), D.Sp, s));
    in_u = prod_rows(D, t + 6, nSib, s, in_u);
    const int *q = t + 6 + nSib;
    const int n_down = q[0];
    const int *mem = q + 1;   // { role, up_i_row, stub_row, dn_row } * n_down

    if (!selfing) {
        Mat3 Wu; ones9(Wu); fold(Wu, urole, in_u);
        Mat3 Wall = Wu;
        for (int x = 0; x < n_down; x++) {
            Mat3 W = Wu;
            for (int y = 0; y < n_down; y++) {
                if (y == x) continue;
                fold(W, mem[4 * y], ld_tri(row_ptr(D, mem[4 * y + 1]), D.Sp, s));
            }
            const Tri up = ld_tri(row_ptr(D, mem[4 * x + 1]), D.Sp, s);
            const Tri dn = emit(W, mem[4 * x]);
            st_tri(row_ptr(D, mem[4 * x + 2]), D.Sp, s, mul(up, dn));
            if (mem[4 * x + 3] >= 0) st_tri(row_ptr(D, mem[4 * x + 3]), D.Sp, s, dn);
            fold(Wall, mem[4 * x], up);
        }
        st_tri(row_ptr(D, prong_row), D.Sp, s, mat_to_kid(Wall));
    } else {
        Tri wu{1.0, 1.0, 1.0}; fold_self(wu, urole, in_u);
        Tri wall = wu;
        for (int x = 0; x < n_down; x++) {
            Tri w = wu;
            for (int y = 0; y < n_down; y++) {
                if (y == x) continue;
                fold_self(w, mem[4 * y], ld_tri(row_ptr(D, mem[4 * y + 1]), D.Sp, s));
            }
            const Tri up = ld_tri(row_ptr(D, mem[4 * x + 1]), D.Sp, s);
            const Tri dn = emit_self(w, mem[4 * x]);
            st_tri(row_ptr(D, mem[4 * x + 2]), D.Sp, s, mul(up, dn));
            if (mem[4 * x + 3] >= 0) st_tri(row_ptr(D, mem[4 * x + 3]), D.Sp, s, dn);
            fold_self(wall, mem[4 * x], up);
        }
        st_tri(row_ptr(D, prong_row), D.Sp, s, self_to_kid(wall));
    }
}

__global__ void __launch_bounds__(SWEEP_WARPS * 32) sweep_kernel(const __grid_constant__ SweepParams P)
{
    const int tile = blockIdx.x, group = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = tile * TILE_LOCI + lane;
    const bool valid = s < P.D.Sl;
    const int *loff = P.level_off + (size_t)group * (P.n_levels + 1);
    for (int lev = 0; lev < P.n_levels; lev++) {
        const int beg = loff[lev], end = loff[lev + 1];
        for (int ti = beg + warp; ti < end; ti += SWEEP_WARPS) {
            const SweepTask tk = P.tasks[ti];
            const int *t = P.pool + tk.off;
            if (tk.kind == TASK_ROOT) task_root(P, t, s, valid, tile, lane);
            else if (tk.kind == TASK_FAMILY_UP) task_family_up(P, t, s, valid);
            else task_family_down(P, t, s, valid);
        }
        if (lev + 1 < P.n_levels) __syncthreads();
    }
}

void launch_sweep(const SweepParams &P, int n_tiles, cudaStream_t st)
{
    dim3 grid(n_tiles, P.n_groups);
    sweep_kernel<<<grid, SWEEP_WARPS * 32, 0, st>>>(P);
}

// ---- proposal evaluation ---------------------------------------------------------------------

__global__ void __launch_bounds__(EVAL_WARPS * 32) eval_kernel(const __grid_constant__ EvalParams P)
{
    const DevView &D = P.D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.y * EVAL_WARPS + warp;
    if (gi >= P.n_groups) return;
    const EvalGroup g = P.groups[gi];
    const int chunk = blockIdx.x;
    const int s_beg = chunk * P.chunk_loci;
    const int s_end = min(s_beg + P.chunk_loci, D.Sl);

    EvalProp pr[EVAL_G];
    const double *rp[EVAL_G], *rm[EVAL_G];
    LogProd acc[EVAL_G];
#pragma unroll
    for (int j = 0; j < EVAL_G; j++) {
        lp_init(acc[j]);
        if (j < g.n) {
            pr[j] = P.props[g.begin + j];
            rp[j] = row_ptr(D, pr[j].row_p);
            rm[j] = row_ptr(D, pr[j].row_m);
        } else {
            pr[j].kind = -1; rp[j] = rm[j] = nullptr;
        }
    }
    const double *kid_row = row_ptr(D, g.kid_row);

    for (int s = s_beg + lane; s < s_end; s += 32) {
        const Tri kid = ld_tri(kid_row, D.Sp, s);
        const Sym3 A = kid_matrix(kid);
#pragma unroll
        for (int j = 0; j < EVAL_G; j++) {
            if (pr[j].kind < 0) continue;
            double v;
            if (pr[j].kind == 0) {
                const Tri xp = ld_tri(rp[j], D.Sp, s), xm = ld_tri(rm[j], D.Sp, s);
                v = dot(xm, sym_mv(A, xp));
            } else if (pr[j].kind == 1) {
                const Tri xp = ld_tri(rp[j], D.Sp, s);
                v = dot(kid, self_to_kid(xp));
            } else {
                v = dot(ld_tri(rp[j], D.Sp, s), kid);
            }
            lp_mul(acc[j], v);
        }
    }
#pragma unroll
    for (int j = 0; j < EVAL_G; j++) {
        if (pr[j].kind < 0) continue;
        const LogProd r = lp_warp_reduce(acc[j]);
        if (lane == 0) P.partial[(size_t)chunk * P.n_props + g.begin + j] = lp_log(r);
    }
}

void launch_eval(const EvalParams &P, int n_chunks, cudaStream_t st)
{
    dim3 grid(n_chunks, (P.n_groups + EVAL_WARPS - 1) / EVAL_WARPS);
    eval_kernel<<<grid, EVAL_WARPS * 32, 0, st>>>(P);
}

// ---- fixed-order reduction of per-tile / per-chunk partial sums ----------------------------------

__global__ void __launch_bounds__(128) reduce_kernel(const __grid_constant__ ReduceParams P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    double acc = 0.0;
    for (int c = 0; c < P.n_parts; c++) acc += P.partial[(size_t)c * P.n + i];
    P.out[i] = acc;
}

void launch_reduce(const ReduceParams &P, cudaStream_t st)
{
    reduce_kernel<<<(P.n + 127) / 128, 128, 0, st>>>(P);
}

} // namespace pf
