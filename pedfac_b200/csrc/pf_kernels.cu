// pf_kernels.cu — hand-written sm_100a kernels of libpedfac_cuda: the level-synchronous
// sum-product sweep, the batched proposal evaluation and the fixed-order partial reduction.
// All arithmetic is fp64 (3-state contractions: no tensor-core shape exists here); loci are the
// contiguous axis of every row so warps issue full-sector coalesced accesses.
#include "pf_kernels.cuh"

namespace pf {

// ---- sweep ----------------------------------------------------------------------------------

// Per-thread context: where the staged plan, genotype tiles and transient rows live. The three
// row bases already include this thread's locus, so resolving a reference is one multiply-add.
struct SweepCtx {
    const EmisEnt *emis;      // group's genotype entries (shared or global)
    const char *edata;        // staged genotype tiles
    const double *prior;      // [3][32] staged HWE prior of the tile
    const double *eps;        // [32]
    double *smem_lane;        // transient rows in shared memory, + lane
    double *scratch_lane;     // spilled transient rows of this group, + locus
    double *rows_lane;        // persistent rows, + locus
    int Sp, Sl, s, lane, tile;
};

// Row access. A reference is either a transient row in shared memory, a spilled transient row or
// a persistent row in global memory; the branch is warp-uniform and keeps each access in a known
// address space (LDS / LDG instead of generic loads with 64-bit address selects).
__device__ __forceinline__ Tri ld_row(const SweepCtx &C, int ref)
{
    const int idx = ref & (SMEM_ROW_BIT - 1);
    if (ref & SMEM_ROW_BIT) {
        const double *p = C.smem_lane + idx * (3 * TILE_LOCI);
        return Tri{p[0], p[TILE_LOCI], p[2 * TILE_LOCI]};
    }
    const double *p = ((ref & SCRATCH_BIT) ? C.scratch_lane : C.rows_lane) + (size_t)idx * 3 * C.Sp;
    return Tri{p[0], p[C.Sp], p[2 * C.Sp]};
}
__device__ __forceinline__ void st_row(const SweepCtx &C, int ref, Tri v)
{
    const int idx = ref & (SMEM_ROW_BIT - 1);
    if (ref & SMEM_ROW_BIT) {
        double *p = C.smem_lane + idx * (3 * TILE_LOCI);
        p[0] = v.x; p[TILE_LOCI] = v.y; p[2 * TILE_LOCI] = v.z;
        return;
    }
    double *p = ((ref & SCRATCH_BIT) ? C.scratch_lane : C.rows_lane) + (size_t)idx * 3 * C.Sp;
    p[0] = v.x; p[C.Sp] = v.y; p[2 * C.Sp] = v.z;
}

// exact num / den for 16-bit numerators: Markstein's correction step on the rounded reciprocal
__device__ __forceinline__ double div_exact(double num, double den, double rcp)
{
    const double q = num * rcp;
    return fma(fma(-den, q, num), rcp, q);
}

// one locus of a genotype row: planes of `ps` loci starting at p, this thread's locus index i
// (_ReadPedfile@0x10000b1ca-0x10000b6e2 emission model)
__device__ __forceinline__ Tri decode_emis(const char *p, unsigned kind, int i, int ps, const EmisEnt &E, double ep)
{
    if (kind == EMIS_HARD) {
        const uint32_t w = reinterpret_cast<const uint32_t *>(p)[i >> 4];
        const int c = (w >> ((i & 15) * 2)) & 3;
        if (c == 3) return Tri{1.0, 1.0, 1.0};
        const double lo = ep / 2.0, hi = 1.0 - ep;
        return Tri{c == 0 ? hi : lo, c == 1 ? hi : lo, c == 2 ? hi : lo};
    }
    if (kind == EMIS_DEC16) {
        const uint16_t *u = reinterpret_cast<const uint16_t *>(p);
        const double den = (double)E.denom, rcp = E.rcp;
        return Tri{div_exact((double)u[i], den, rcp), div_exact((double)u[ps + i], den, rcp), div_exact((double)u[2 * ps + i], den, rcp)};
    }
    if (kind == EMIS_F32) {
        const float *f = reinterpret_cast<const float *>(p);
        return Tri{(double)f[i], (double)f[ps + i], (double)f[2 * ps + i]};
    }
    const double *f = reinterpret_cast<const double *>(p);
    return Tri{f[i], f[ps + i], f[2 * ps + i]};
}

// local(i) = emission x (founder ? HWE prior : 1) from the staged genotype tile
// (_ReadPedfile@0x10000b1ca-0x10000b6e2 emission model; _PassingOFtoIV@0x10001c2d0)
__device__ __forceinline__ Tri local_of(const SweepCtx &C, int member)
{
    const EmisEnt &E = C.emis[member >> 3];
    const unsigned kind = E.kind, off = E.smem_off;
    Tri e{1.0, 1.0, 1.0};
    const int l = C.lane;
    if (kind != EMIS_NONE) {
        if (off != EMIS_NOT_STAGED) e = decode_emis(C.edata + off, kind, l, TILE_LOCI, E, C.eps[l]);
        else e = decode_emis(reinterpret_cast<const char *>(((unsigned long long)E.ptr_hi << 32) | E.ptr_lo), kind, C.s, C.Sp, E, C.eps[l]);
    }
    if (member & 4) e = mul(e, Tri{C.prior[l], C.prior[TILE_LOCI + l], C.prior[2 * TILE_LOCI + l]});
    return e;
}

__device__ __forceinline__ Tri prod_rows(const SweepCtx &C, const int *rows, int n, Tri acc)
{
    for (int j = 0; j < n; j++) acc = mul(acc, ld_row(C, rows[j]));
    return acc;
}

// fold one member's incoming message into the family tensor W[gm][gp]; one code path for all
// roles: the member's factor matrix is in[gp] (father), in[gm] (mother) or K(gm,gp) (kid).
// A selfing marriage uses the same tensor and reads only its diagonal (gm == gp).
__device__ __forceinline__ void fold(Mat3 &W, int role, Tri in)
{
    // role is warp-uniform (one task per warp), so these branches never diverge
    if (role == 2) mat_mul_kid(W, in);
    else if (role == 0) mat_mul_pa(W, in);
    else mat_mul_ma(W, in);
}
__device__ __forceinline__ Tri emit(const Mat3 &W, int role, int selfing)
{
    if (selfing) return role == 0 ? mat_diag(W) : self_to_kid(mat_diag(W));
    return role == 0 ? mat_to_pa(W) : role == 1 ? mat_to_ma(W) : mat_to_kid(W);
}

__device__ __forceinline__ void ones9(Mat3 &W)
{
#pragma unroll
    for (int i = 0; i < 9; i++) W.m[i] = 1.0;
}

// fold the listed { role, up_i_row } members into W
// up message of a family member: a stored row, or recomputed for a leaf
__device__ __forceinline__ Tri up_msg(const SweepCtx &C, int ref)
{
    return ref < 0 ? local_of(C, ref & 0x7fffffff) : ld_row(C, ref);
}
__device__ __forceinline__ void fold_members(const SweepCtx &C, Mat3 &W, const int *mem, int n)
{
    for (int y = 0; y < n; y++) fold(W, mem[2 * y], up_msg(C, mem[2 * y + 1]));
}

__device__ __forceinline__ void task_root(const SweepParams &P, const SweepCtx &C, const int *t)
{
    const Tri v = prod_rows(C, t + 4, t[3], local_of(C, t[0]));
    st_row(C, t[1], v);
    if (t[2] >= 0) {
        LogProd z; lp_init(z);
        if (C.s < C.Sl) lp_mul(z, (v.x + v.y) + v.z);
        z = lp_warp_reduce(z);
        if (C.lane == 0) P.partial[(size_t)C.tile * P.n_out + t[2]] = lp_log(z);
    }
}

__device__ __forceinline__ void task_up_in(const SweepCtx &C, const int *t)
{
    st_row(C, t[1], prod_rows(C, t + 3, t[2], local_of(C, t[0])));
}

__device__ __forceinline__ void task_up_m(const SweepCtx &C, const int *t)
{
    Mat3 W; ones9(W);
    fold_members(C, W, t + 3, t[2]);
    st_row(C, t[1], emit(W, t[0] >> 1, t[0] & 1));
}

__device__ __forceinline__ void task_dn_in(const SweepCtx &C, const int *t)
{
    Tri in_u = local_of(C, t[0]);
    if (t[2] >= 0) in_u = mul(in_u, ld_row(C, t[2]));
    st_row(C, t[1], prod_rows(C, t + 4, t[3], in_u));
}

__device__ __forceinline__ void task_dn_out(const SweepCtx &C, const int *t)
{
    const int flags = t[0];
    Mat3 W; ones9(W);
    fold(W, (flags >> 1) & 3, ld_row(C, t[1]));
    fold_members(C, W, t + 6, t[5]);
    const Tri dn = emit(W, flags >> 3, flags & 1);
    st_row(C, t[3], mul(up_msg(C, t[2]), dn));
    if (t[4] >= 0) st_row(C, t[4], dn);
}

__device__ __forceinline__ void task_prong(const SweepCtx &C, const int *t)
{
    const int flags = t[0];
    Mat3 W; ones9(W);
    fold(W, (flags >> 1) & 3, ld_row(C, t[1]));
    fold_members(C, W, t + 4, t[3]);
    st_row(C, t[2], (flags & 1) ? self_to_kid(mat_diag(W)) : mat_to_kid(W));
}

// STAGED: every group's plan slice lives in shared memory (the normal case); otherwise plans are
// read from global memory. Two instantiations keep every pointer in a known address space.
template <bool STAGED>
__global__ void __launch_bounds__(SWEEP_WARPS * 32, 1) sweep_kernel(const __grid_constant__ SweepParams P)
{
    extern __shared__ __align__(16) char smem[];
    const DevView &D = P.D;
    const int tile = blockIdx.x;
    const SweepGroup G = P.groups[blockIdx.y];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthr = SWEEP_WARPS * 32;
    const SweepSmem L = sweep_smem_layout(G, P.n_levels);

    // ---- stage the plan slice, the tile's prior/eps and every individual's genotype tile ----
    int *s_loff = reinterpret_cast<int *>(smem + L.loff);
    for (int i = tid; i <= P.n_levels; i += nthr) s_loff[i] = P.level_off[(size_t)blockIdx.y * (P.n_levels + 1) + i];
    double *s_prior = reinterpret_cast<double *>(smem + L.prior);
    double *s_eps = reinterpret_cast<double *>(smem + L.eps);
    if (tid < 3 * TILE_LOCI) s_prior[tid] = D.rows[(size_t)ROW_PRIOR * 3 * D.Sp + (size_t)(tid / TILE_LOCI) * D.Sp + tile * TILE_LOCI + (tid % TILE_LOCI)];
    else if (tid < 4 * TILE_LOCI) s_eps[tid - 3 * TILE_LOCI] = D.eps[tile * TILE_LOCI + (tid - 3 * TILE_LOCI)];
    const SweepTask *tasks = P.tasks + G.task_begin;
    const int *pool = P.pool + G.pool_begin;
    const EmisEnt *emis = P.emis + G.emis_begin;
    if (STAGED) {
        int *d;
        const int *src;
        d = reinterpret_cast<int *>(smem + L.tasks); src = reinterpret_cast<const int *>(tasks);
        for (int i = tid; i < (G.task_end - G.task_begin) * 2; i += nthr) d[i] = src[i];
        d = reinterpret_cast<int *>(smem + L.pool);
        for (int i = tid; i < G.pool_len; i += nthr) d[i] = pool[i];
        d = reinterpret_cast<int *>(smem + L.emis); src = reinterpret_cast<const int *>(emis);
        for (int i = tid; i < G.emis_count * (int)(sizeof(EmisEnt) / 4); i += nthr) d[i] = src[i];
    }
    if (STAGED) {
        tasks = reinterpret_cast<const SweepTask *>(smem + L.tasks);
        pool = reinterpret_cast<const int *>(smem + L.pool);
        emis = reinterpret_cast<const EmisEnt *>(smem + L.emis);
    }
    // genotype tiles: one warp per individual, lanes along the loci (global entries: the copy in
    // shared memory may not be visible yet, so read the entry from global memory here)
    {
        const EmisEnt *gemis = P.emis + G.emis_begin;
        char *edata = smem + L.edata;
        const int s = tile * TILE_LOCI + lane;
        for (int i = warp; i < G.emis_count; i += SWEEP_WARPS) {
            const EmisEnt E = gemis[i];
            if (E.smem_off == EMIS_NOT_STAGED || E.kind == EMIS_NONE) continue;
            const void *gp = reinterpret_cast<const void *>(((unsigned long long)E.ptr_hi << 32) | E.ptr_lo);
            char *dst = edata + E.smem_off;
            if (E.kind == EMIS_HARD) {
                if (lane < 2) reinterpret_cast<uint32_t *>(dst)[lane] = static_cast<const uint32_t *>(gp)[tile * 2 + lane];
            } else if (E.kind == EMIS_DEC16) {
                const uint16_t *u = static_cast<const uint16_t *>(gp);
                uint16_t *o = reinterpret_cast<uint16_t *>(dst);
                o[lane] = u[s]; o[TILE_LOCI + lane] = u[D.Sp + s]; o[2 * TILE_LOCI + lane] = u[2 * D.Sp + s];
            } else if (E.kind == EMIS_F32) {
                const float *f = static_cast<const float *>(gp);
                float *o = reinterpret_cast<float *>(dst);
                o[lane] = f[s]; o[TILE_LOCI + lane] = f[D.Sp + s]; o[2 * TILE_LOCI + lane] = f[2 * D.Sp + s];
            } else {
                const double *f = static_cast<const double *>(gp);
                double *o = reinterpret_cast<double *>(dst);
                o[lane] = f[s]; o[TILE_LOCI + lane] = f[D.Sp + s]; o[2 * TILE_LOCI + lane] = f[2 * D.Sp + s];
            }
        }
    }
    __syncthreads();

    SweepCtx C;
    C.emis = emis; C.edata = smem + L.edata; C.prior = s_prior; C.eps = s_eps;
    C.lane = lane; C.tile = tile; C.s = tile * TILE_LOCI + lane; C.Sp = D.Sp; C.Sl = D.Sl;
    C.smem_lane = reinterpret_cast<double *>(smem + L.rows) + lane;
    C.scratch_lane = D.scratch + (size_t)G.scratch_base * 3 * D.Sp + C.s;
    C.rows_lane = D.rows + C.s;

    for (int lev = 0; lev < P.n_levels; lev++) {
        const int beg = s_loff[lev] - G.task_begin, end = s_loff[lev + 1] - G.task_begin;
        for (int ti = beg + warp; ti < end; ti += SWEEP_WARPS) {
            const SweepTask tk = tasks[ti];
            const int *t = pool + (tk.off - G.pool_begin);
            switch (tk.kind) {
            case TASK_ROOT: task_root(P, C, t); break;
            case TASK_UP_IN: task_up_in(C, t); break;
            case TASK_UP_M: task_up_m(C, t); break;
            case TASK_DN_IN: task_dn_in(C, t); break;
            case TASK_DN_OUT: task_dn_out(C, t); break;
            default: task_prong(C, t); break;
            }
        }
        if (lev + 1 < P.n_levels) __syncthreads();
    }
}

void launch_sweep(const SweepParams &P, int n_tiles, int smem_bytes, bool staged, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM_MAX);
        cudaFuncSetAttribute(sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM_MAX);
        attr_set = true;
    }
    dim3 grid(n_tiles, P.n_groups);
    if (staged) sweep_kernel<true><<<grid, SWEEP_WARPS * 32, smem_bytes, st>>>(P);
    else sweep_kernel<false><<<grid, SWEEP_WARPS * 32, smem_bytes, st>>>(P);
}

// ---- proposal evaluation ---------------------------------------------------------------------

// All row loads of a locus (kid + two rows per proposal of the group) are issued up front and
// unconditionally — a proposal that has no second row points at the prior row — so one warp keeps
// 27 independent 256 B segments in flight; the three contraction forms are then all computed and
// the proposal's kind selects one (a few spare flops on a memory-bound kernel).
__global__ void __launch_bounds__(EVAL_WARPS * 32, 4) eval_kernel(const __grid_constant__ EvalParams P)
{
    const DevView &D = P.D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.y * EVAL_WARPS + warp;
    const int chunk = blockIdx.x;
    EvalGroup g{ROW_PRIOR, 0, 0, 0};
    if (gi < P.n_groups) g = P.groups[gi];
    const int s_beg = chunk * P.chunk_loci;
    const int s_end = min(s_beg + P.chunk_loci, D.Sl);
    const size_t row_elems = (size_t)3 * D.Sp;

    int kind[EVAL_G];
    const double *rp[EVAL_G], *rm[EVAL_G];
    LogProd acc[EVAL_G];
#pragma unroll
    for (int j = 0; j < EVAL_G; j++) {
        lp_init(acc[j]);
        EvalProp pr{0, ROW_PRIOR, ROW_PRIOR, 0};
        if (j < g.n) pr = P.props[g.begin + j];
        kind[j] = pr.kind;
        rp[j] = D.rows + (size_t)pr.row_p * row_elems;
        rm[j] = D.rows + (size_t)pr.row_m * row_elems;
    }
    const double *kid_row = D.rows + (size_t)g.kid_row * row_elems;

    for (int s = s_beg + lane; s < s_end; s += 32) {
        const Tri kid = ld_tri(kid_row, D.Sp, s);
        Tri xp[EVAL_G], xm[EVAL_G];
#pragma unroll
        for (int j = 0; j < EVAL_G; j++) { xp[j] = ld_tri(rp[j], D.Sp, s); xm[j] = ld_tri(rm[j], D.Sp, s); }
        const Sym3 A = kid_matrix(kid);
#pragma unroll
        for (int j = 0; j < EVAL_G; j++) {
            const double vt = dot(xm[j], sym_mv(A, xp[j]));         // trio
            const double vs = dot(kid, self_to_kid(xp[j]));        // selfing
            const double vp = dot(xp[j], kid);                      // prong
            const double v = kind[j] == 0 ? vt : kind[j] == 1 ? vs : vp;
            if (j < g.n) lp_mul(acc[j], v);
        }
    }
#pragma unroll
    for (int j = 0; j < EVAL_G; j++) {
        if (j >= g.n) continue;
        const LogProd r = lp_warp_reduce(acc[j]);
        if (lane == 0) { P.partial[(size_t)chunk * P.n_props + g.begin + j] = lp_log(r); __threadfence(); }
    }

    // fused fixed-order reduction: the last block of this column of the grid to finish adds the
    // chunks' partial sums in chunk order (deterministic whichever block happens to be last)
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&P.counters[blockIdx.y], 1u);
        s_last = done == gridDim.x - 1;
        if (s_last) P.counters[blockIdx.y] = 0;     // ready for the next launch
    }
    __syncthreads();
    if (s_last && threadIdx.x < EVAL_WARPS * EVAL_G) {
        const int gi2 = blockIdx.y * EVAL_WARPS + threadIdx.x / EVAL_G, j = threadIdx.x % EVAL_G;
        if (gi2 < P.n_groups) {
            const EvalGroup g2 = P.groups[gi2];
            if (j < g2.n) {
                const double *src = P.partial + g2.begin + j;
                double sum = 0.0;
#pragma unroll 8
                for (int c = 0; c < (int)gridDim.x; c++) sum += __ldcg(src + (size_t)c * P.n_props);
                P.out[g2.begin + j] = sum;
            }
        }
    }
}

void launch_eval(const EvalParams &P, int n_chunks, cudaStream_t st)
{
    dim3 grid(n_chunks, (P.n_groups + EVAL_WARPS - 1) / EVAL_WARPS);
    eval_kernel<<<grid, EVAL_WARPS * 32, 0, st>>>(P);
}

// ---- fixed-order reduction of per-tile partial sums (sweeps) ----------------------------------------

// One warp per output: lanes add parts lane, lane+32, ... (independent L2 loads), then a fixed
// xor-butterfly adds the 32 lane sums — the order depends only on n_parts, so it is deterministic.
__global__ void __launch_bounds__(128) reduce_kernel(const __grid_constant__ ReduceParams P)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= P.n) return;
    double acc = 0.0;
#pragma unroll 4
    for (int c = lane; c < P.n_parts; c += 32) acc += __ldcg(P.partial + (size_t)c * P.n + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) P.out[i] = acc;
}

void launch_reduce(const ReduceParams &P, cudaStream_t st)
{
    reduce_kernel<<<(P.n * 32 + 127) / 128, 128, 0, st>>>(P);
}

} // namespace pf
