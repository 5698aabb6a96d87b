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
    int Sp, Sl, s, lane, tile, n_rows_smem;
};


// Row access. A reference is either a transient row in shared memory, a spilled transient row or
// a persistent row in global memory; the branch is warp-uniform and keeps each access in a known
// address space (LDS / LDG instead of generic loads with 64-bit address selects).
__device__ __forceinline__ Tri ld_row(const SweepCtx &C, int ref)
{
    const int idx = ref & (SMEM_ROW_BIT - 1);
    if ((ref & SCRATCH_BIT) && idx < C.n_rows_smem) {
        const double *p = C.smem_lane + idx * (3 * TILE_LOCI);
        return Tri{p[0], p[TILE_LOCI], p[2 * TILE_LOCI]};
    }
    const double *p = ((ref & SCRATCH_BIT) ? C.scratch_lane : C.rows_lane) + (size_t)idx * 3 * C.Sp;
    return Tri{p[0], p[C.Sp], p[2 * C.Sp]};
}
__device__ __forceinline__ void st_row(const SweepCtx &C, int ref, Tri v)
{
    const int idx = ref & (SMEM_ROW_BIT - 1);
    if ((ref & SCRATCH_BIT) && idx < C.n_rows_smem) {
        double *p = C.smem_lane + idx * (3 * TILE_LOCI);
        p[0] = v.x; p[TILE_LOCI] = v.y; p[2 * TILE_LOCI] = v.z;
        return;
    }
    double *p = ((ref & SCRATCH_BIT) ? C.scratch_lane : C.rows_lane) + (size_t)idx * 3 * C.Sp;
    p[0] = v.x; p[C.Sp] = v.y; p[2 * C.Sp] = v.z;
}

// Specialised accessors (fewer tests on the task's critical path than the generic pair above):
// persistent rows (stubs, prongs) are only ever stored; transient rows are in shared memory or
// spilled to the scratch array, never persistent.
__device__ __forceinline__ void st_persist(const SweepCtx &C, int row, Tri v)
{
    double *p = C.rows_lane + (size_t)row * 3 * C.Sp;
    p[0] = v.x; p[C.Sp] = v.y; p[2 * C.Sp] = v.z;
}
__device__ __forceinline__ Tri ld_t(const SweepCtx &C, int ref)
{
    const int idx = ref & (SMEM_ROW_BIT - 1);
    if (idx < C.n_rows_smem) {
        const double *p = C.smem_lane + idx * (3 * TILE_LOCI);
        return Tri{p[0], p[TILE_LOCI], p[2 * TILE_LOCI]};
    }
    const double *p = C.scratch_lane + (size_t)idx * 3 * C.Sp;
    return Tri{p[0], p[C.Sp], p[2 * C.Sp]};
}
__device__ __forceinline__ void st_t(const SweepCtx &C, int ref, Tri v)
{
    const int idx = ref & (SMEM_ROW_BIT - 1);
    if (idx < C.n_rows_smem) {
        double *p = C.smem_lane + idx * (3 * TILE_LOCI);
        p[0] = v.x; p[TILE_LOCI] = v.y; p[2 * TILE_LOCI] = v.z;
        return;
    }
    double *p = C.scratch_lane + (size_t)idx * 3 * C.Sp;
    p[0] = v.x; p[C.Sp] = v.y; p[2 * C.Sp] = v.z;
}
// transient row number idx of this group (tasks that walk consecutive rows of a bank)
__device__ __forceinline__ Tri ld_i(const SweepCtx &C, int idx)
{
    if (idx < C.n_rows_smem) {
        const double *p = C.smem_lane + idx * (3 * TILE_LOCI);
        return Tri{p[0], p[TILE_LOCI], p[2 * TILE_LOCI]};
    }
    const double *p = C.scratch_lane + (size_t)idx * 3 * C.Sp;
    return Tri{p[0], p[C.Sp], p[2 * C.Sp]};
}
__device__ __forceinline__ void st_i(const SweepCtx &C, int idx, Tri v)
{
    if (idx < C.n_rows_smem) {
        double *p = C.smem_lane + idx * (3 * TILE_LOCI);
        p[0] = v.x; p[TILE_LOCI] = v.y; p[2 * TILE_LOCI] = v.z;
        return;
    }
    double *p = C.scratch_lane + (size_t)idx * 3 * C.Sp;
    p[0] = v.x; p[C.Sp] = v.y; p[2 * C.Sp] = v.z;
}

// reference to transient row idx of this group (used when a task walks consecutive scratch rows)
__device__ __forceinline__ int scratch_ref(const SweepCtx &C, int idx)
{
    return SCRATCH_BIT | idx | (idx < C.n_rows_smem ? SMEM_ROW_BIT : 0);
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
    for (int j = 0; j < n; j++) acc = mul(acc, ld_t(C, rows[j]));
    return acc;
}

// up message of a family member: a stored row, or recomputed for a leaf without a cached row
__device__ __forceinline__ Tri up_msg(const SweepCtx &C, int ref)
{
    return ref < 0 ? local_of(C, ref & 0x7fffffff) : ld_t(C, ref);
}

__device__ __forceinline__ void task_root(const SweepParams &P, const SweepCtx &C, const int *t)
{
    const Tri v = prod_rows(C, t + 4, t[3], local_of(C, t[0]));
    st_row(C, t[1], v);
    if (t[2] >= 0) {
        LogProd z; lp_init(z);
        if (C.s < C.Sl) lp_mul(z, (v.x + v.y) + v.z);
        z = lp_warp_reduce(z);
        if (C.lane == 0) { P.partial[(size_t)C.tile * P.n_out + t[2]] = lp_log(z); __threadfence(); }
    }
}

// LEAF_IN: member, up_i_row — decode a leaf's local factor once into its cached row (level 0)
__device__ __forceinline__ void task_leaf_in(const SweepCtx &C, const int *t)
{
    st_t(C, t[1], local_of(C, t[0]));
}

// product of the kid factors K(up_msg(ref)) of refs[0..n), four at a time so that the twelve
// shared-memory loads of a chunk are in flight together (a long sibship is otherwise one
// dependent load -> fma chain per kid)
__device__ __forceinline__ void fold_kids(const SweepCtx &C, Sym3 &P, const int *refs, int stride, int n)
{
    constexpr int B = 4;
    int j = 0;
    for (; j + B <= n; j += B) {
        int ref[B];
        Tri in[B];
#pragma unroll
        for (int c = 0; c < B; c++) ref[c] = refs[(j + c) * stride];
#pragma unroll
        for (int c = 0; c < B; c++) in[c] = up_msg(C, ref[c]);
#pragma unroll
        for (int c = 0; c < B; c++) sym_mul(P, kid_matrix(in[c]));
    }
    for (; j < n; j++) sym_mul(P, kid_matrix(up_msg(C, refs[j * stride])));
}

// A sibship is a list of items {type, value}: type 0 = one kid, value = its up reference (a row, or
// a leaf to decode); type 1 = a chunk of 2..FAM_CHUNK leaf kids whose factor product a KFOLD task
// (level 0, in parallel on the block's warps) left in the transient rows value, value + 1. The
// family tasks then work per item: a sibship of n leaves costs them n / FAM_CHUNK row loads
// instead of n decodes, and the per-kid work of the down band is done by the KID_CHUNK tasks.
__device__ __forceinline__ Sym3 item_factor(const SweepCtx &C, const int *it)
{
    if (it[0] == 0) return kid_matrix(up_msg(C, it[1]));
    const Tri lo = ld_i(C, it[1]), hi = ld_i(C, it[1] + 1);
    return Sym3{lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
}

// KFOLD: out_row (2 consecutive transient rows), n, kid_ref[n]
__device__ __forceinline__ void task_kfold(const SweepCtx &C, const int *t)
{
    Sym3 P = sym_ones();
    fold_kids(C, P, t + 2, 1, t[1]);
    st_i(C, t[0], Tri{P.a00, P.a01, P.a02});
    st_i(C, t[0] + 1, Tri{P.a11, P.a12, P.a22});
}

// FAM_UP: flags, up_m_row, nC, { member, up_i_row, nD, below[nD] } * nC, pa_ref, ma_ref, n_items, item[n_items]
//   the nC non-leaf down members get their up message computed and stored first; pa_ref / ma_ref
//   are NO_REF when that parent is the up individual (or the marriage is selfing)
__device__ __forceinline__ void task_fam_up(const SweepCtx &C, const int *t)
{
    const int flags = t[0], selfing = flags & 1, urole = flags >> 1, nC = t[2];
    const int *q = t + 3;
    for (int x = 0; x < nC; x++) {
        st_t(C, q[1], prod_rows(C, q + 3, q[2], local_of(C, q[0])));
        q += 3 + q[2];
    }
    const Tri one{1.0, 1.0, 1.0};
    const Tri a = q[0] == NO_REF ? one : up_msg(C, q[0]);
    const Tri b = q[1] == NO_REF ? one : up_msg(C, q[1]);
    Sym3 P = sym_ones();
    const int n_items = q[2];
    for (int i = 0; i < n_items; i++) sym_mul(P, item_factor(C, q + 3 + 2 * i));
    const Tri out = urole == 0 ? fam_to_pa(P, b, selfing) : urole == 1 ? fam_to_pa(P, a, 0) : fam_to_kid(P, a, b, selfing);
    st_t(C, t[1], out);
}

// message + stub of one down member
__device__ __forceinline__ void put_down(const SweepCtx &C, const int *ent, Tri up, Tri dn)
{
    st_persist(C, ent[1], mul(up, dn));
    if (ent[2] >= 0) st_t(C, ent[2], dn);
}

// FAM_DOWN: flags, prong_row, u_member, u_dn_row (-1 = root), nSib, sib[nSib],
//           pa{ref, stub_row, dn_row}, ma{ref, stub_row, dn_row}, n_items, tmp_row, item[n_items]
//           (+ stub_row, dn_row of the kid when the sibship is one single kid: tmp_row = -1)
//   pa.ref / ma.ref = NO_REF when that parent is the up individual (or selfing mother);
//   tmp_row = first of 4 + 2 * n_items bank rows: a, b, Pu (the factor of the up individual when it
//   is a kid of the marriage, else ones; 2 rows), then per item the product of the OTHER items'
//   factors, which a FAM_EXCL task computes on another warp at the same level; the items'
//   KID_CHUNK tasks on the next level turn them into messages
__device__ __forceinline__ void task_fam_down(const SweepCtx &C, const int *t)
{
    const int flags = t[0], selfing = flags & 1, urole = flags >> 1, nSib = t[4];
    Tri in_u = local_of(C, t[2]);
    if (t[3] >= 0) in_u = mul(in_u, ld_t(C, t[3]));
    in_u = prod_rows(C, t + 5, nSib, in_u);
    const int *pa = t + 5 + nSib, *ma = pa + 3;
    const int n_items = ma[3], tmp = ma[4];
    const int *item = ma + 5;
    const Tri one{1.0, 1.0, 1.0};
    const Tri a = urole == 0 ? in_u : pa[0] == NO_REF ? one : up_msg(C, pa[0]);
    const Tri b = urole == 1 ? in_u : ma[0] == NO_REF ? one : up_msg(C, ma[0]);
    Sym3 Pu = sym_ones();
    if (urole == 2) sym_mul(Pu, kid_matrix(in_u));

    Sym3 Pall = Pu;
    if (tmp < 0) {
        // no down kid, or one single kid: finished here
        Tri kmsg = one;
        if (n_items) { kmsg = up_msg(C, item[1]); sym_mul(Pall, kid_matrix(kmsg)); }
        if (n_items) put_down(C, item + 1, kmsg, fam_to_kid(Pu, a, b, selfing));   // item + 1 = {ref, stub_row, dn_row}
    } else {
        st_i(C, tmp, a);
        st_i(C, tmp + 1, b);
        st_i(C, tmp + 2, Tri{Pu.a00, Pu.a01, Pu.a02});
        st_i(C, tmp + 3, Tri{Pu.a11, Pu.a12, Pu.a22});
        for (int i = 0; i < n_items; i++) sym_mul(Pall, item_factor(C, item + 2 * i));
    }
    st_persist(C, t[1], fam_to_kid(Pall, a, b, selfing));                     // prong
    if (urole != 0 && pa[0] != NO_REF) put_down(C, pa, a, fam_to_pa(Pall, b, selfing));
    if (urole != 1 && ma[0] != NO_REF) put_down(C, ma, b, fam_to_pa(Pall, a, 0));
}

// FAM_EXCL: n_items (>= 2), out_row, item[n_items] — for every item the product of the factors of
//   all OTHER items (exclusive prefix x exclusive suffix), left in rows out_row + 2 i, + 2 i + 1.
//   Independent of the messages from above, so it runs beside FAM_DOWN instead of inside it.
__device__ __forceinline__ void task_fam_excl(const SweepCtx &C, const int *t)
{
    const int n_items = t[0], out = t[1];
    const int *item = t + 2;
    Sym3 run = sym_ones();
    for (int i = 0; i < n_items; i++) {
        const int r = out + 2 * i;
        st_i(C, r, Tri{run.a00, run.a01, run.a02});
        st_i(C, r + 1, Tri{run.a11, run.a12, run.a22});
        if (i + 1 < n_items) sym_mul(run, item_factor(C, item + 2 * i));
    }
    Sym3 suf = item_factor(C, item + 2 * (n_items - 1));
    for (int i = n_items - 2; i >= 0; i--) {
        const int r = out + 2 * i;
        const Tri p0 = ld_i(C, r), p1 = ld_i(C, r + 1);
        st_i(C, r, Tri{p0.x * suf.a00, p0.y * suf.a01, p0.z * suf.a02});
        st_i(C, r + 1, Tri{p1.x * suf.a11, p1.y * suf.a12, p1.z * suf.a22});
        if (i > 0) sym_mul(suf, item_factor(C, item + 2 * i));
    }
}

// KID_CHUNK: flags, ab_row (a, b, Pu at ab_row .. ab_row + 3), out_row (product of the other items, 2 rows;
//            -1 = the only item), n, { up_ref, stub_row, dn_row (-1 = not needed) } * n   (n <= FAM_CHUNK kids)
__device__ __forceinline__ void task_kid_chunk(const SweepCtx &C, const int *t)
{
    const int selfing = t[0] & 1, n = t[3];
    const int ab = t[1], orow = t[2];
    const Tri a = ld_i(C, ab), b = ld_i(C, ab + 1);
    const Tri u0 = ld_i(C, ab + 2), u1 = ld_i(C, ab + 3);
    Sym3 Pout{u0.x, u0.y, u0.z, u1.x, u1.y, u1.z};
    if (orow >= 0) {
        const Tri p0 = ld_i(C, orow), p1 = ld_i(C, orow + 1);
        sym_mul(Pout, Sym3{p0.x, p0.y, p0.z, p1.x, p1.y, p1.z});
    }
    const int *mem = t + 4;
    Tri km[FAM_CHUNK];
#pragma unroll
    for (int j = 0; j < FAM_CHUNK; j++)
        if (j < n) km[j] = up_msg(C, mem[3 * j]);
#pragma unroll
    for (int j = 0; j < FAM_CHUNK; j++) {
        if (j >= n) break;
        Sym3 Pm = Pout;
#pragma unroll
        for (int i = 0; i < FAM_CHUNK; i++)
            if (i != j && i < n) sym_mul(Pm, kid_matrix(km[i]));
        put_down(C, mem + 3 * j, km[j], fam_to_kid(Pm, a, b, selfing));
    }
}

__device__ __forceinline__ void cp_async4(unsigned smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

// STAGED: every group's plan slice lives in shared memory (the normal case); otherwise plans are
// read from global memory. Two instantiations keep every pointer in a known address space.
template <bool STAGED, int SWEEP_WARPS, int MIN_BLOCKS>
__global__ void __launch_bounds__(SWEEP_WARPS * 32, MIN_BLOCKS) sweep_kernel(const __grid_constant__ SweepParams P)
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
    __syncthreads();
    // genotype tiles: fire-and-forget asynchronous copies (cp.async, 4 or 8 bytes per lane), one
    // warp per individual, all of a warp's individuals in flight at once; one wait at the end.
    // (A blocking load per individual made this phase a chain of global-memory latencies.)
    {
        char *edata = smem + L.edata;
        const size_t Sp = D.Sp;
        for (int i = warp; i < G.emis_count; i += SWEEP_WARPS) {
            const EmisEnt E = emis[i];
            if (E.smem_off == EMIS_NOT_STAGED || E.kind == EMIS_NONE) continue;
            const char *gp = reinterpret_cast<const char *>(((unsigned long long)E.ptr_hi << 32) | E.ptr_lo);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(edata + E.smem_off);
            if (E.kind == EMIS_HARD) {                      // 8 bytes: 2 words of 16 calls
                if (lane < 2) cp_async4(dst + 4 * lane, gp + ((size_t)tile * 2 + lane) * 4);
            } else if (E.kind == EMIS_DEC16) {              // 3 planes x 64 bytes
                const int pl = lane >> 4, w4 = lane & 15;   // lanes 0-15 plane 0, 16-31 plane 1
                cp_async4(dst + pl * 64 + 4 * w4, gp + ((size_t)pl * Sp + (size_t)tile * TILE_LOCI) * 2 + 4 * w4);
                if (lane < 16) cp_async4(dst + 128 + 4 * w4, gp + ((size_t)2 * Sp + (size_t)tile * TILE_LOCI) * 2 + 4 * w4);
            } else if (E.kind == EMIS_F32) {                // 3 planes x 128 bytes
#pragma unroll
                for (int pl = 0; pl < 3; pl++)
                    cp_async4(dst + pl * 128 + 4 * lane, gp + ((size_t)pl * Sp + (size_t)tile * TILE_LOCI + lane) * 4);
            } else {                                        // 3 planes x 256 bytes
#pragma unroll
                for (int pl = 0; pl < 3; pl++)
                    cp_async8(dst + pl * 256 + 8 * lane, gp + ((size_t)pl * Sp + (size_t)tile * TILE_LOCI + lane) * 8);
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();

    SweepCtx C;
    C.emis = emis; C.edata = smem + L.edata; C.prior = s_prior; C.eps = s_eps;
    C.lane = lane; C.tile = tile; C.s = tile * TILE_LOCI + lane; C.Sp = D.Sp; C.Sl = D.Sl; C.n_rows_smem = G.n_rows_smem;
    C.smem_lane = reinterpret_cast<double *>(smem + L.rows) + lane;
    C.scratch_lane = D.scratch + (size_t)G.scratch_base * 3 * D.Sp + C.s;
    C.rows_lane = D.rows + C.s;

    const bool clk = P.clocks != nullptr && blockIdx.x == 0 && (int)blockIdx.y == P.clock_group && tid == 0;
    if (clk) P.clocks[0] = clock64();
    for (int lev = 0; lev < P.n_levels; lev++) {
        const int beg = s_loff[lev] - G.task_begin, end = s_loff[lev + 1] - G.task_begin;
        for (int ti = beg + warp; ti < end; ti += SWEEP_WARPS) {
            const SweepTask tk = tasks[ti];
            const int *t = pool + (tk.off - G.pool_begin);
            switch (tk.kind) {
            case TASK_ROOT: task_root(P, C, t); break;
            case TASK_LEAF_IN: task_leaf_in(C, t); break;
            case TASK_FAM_UP: task_fam_up(C, t); break;
            case TASK_FAM_DOWN: task_fam_down(C, t); break;
            case TASK_KID_CHUNK: task_kid_chunk(C, t); break;
            case TASK_FAM_EXCL: task_fam_excl(C, t); break;
            default: task_kfold(C, t); break;
            }
        }
        if (lev + 1 < P.n_levels) __syncthreads();
        if (clk) P.clocks[lev + 1] = clock64();
    }
    // fused fixed-order reduction of the per-tile log Z sums (Gibbs-step sweeps have a handful of
    // outputs; a separate launch cost ~6 us): the last block of the grid to finish adds the tiles
    // exactly as reduce_kernel does — lanes take tiles lane, lane+32, ..., then an xor butterfly
    if (P.out != nullptr) {
        __shared__ bool s_last;
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned done = atomicAdd(P.counter, 1u);
            s_last = done == gridDim.x * gridDim.y - 1;
            if (s_last) *P.counter = 0;             // ready for the next launch
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            const int n_parts = gridDim.x;
            for (int i = warp; i < P.n_out; i += SWEEP_WARPS) {
                double acc = 0.0;
#pragma unroll 4
                for (int c = lane; c < n_parts; c += 32) acc += __ldcg(P.partial + (size_t)c * P.n_out + i);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) P.out[i] = acc;
            }
        }
    }
}

void launch_sweep(const SweepParams &P, int n_tiles, int smem_bytes, bool staged, bool throughput_shape, cudaStream_t st)
{
    constexpr SweepShape L = SWEEP_LATENCY, T = SWEEP_THROUGHPUT;
    static unsigned long long attr_set = 0;      // one bit per device: the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_set >> (dev & 63) & 1)) {
        cudaFuncSetAttribute(sweep_kernel<true, L.warps, L.blocks_per_sm>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem_max);
        cudaFuncSetAttribute(sweep_kernel<false, L.warps, L.blocks_per_sm>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem_max);
        cudaFuncSetAttribute(sweep_kernel<true, T.warps, T.blocks_per_sm>, cudaFuncAttributeMaxDynamicSharedMemorySize, T.smem_max);
        cudaFuncSetAttribute(sweep_kernel<false, T.warps, T.blocks_per_sm>, cudaFuncAttributeMaxDynamicSharedMemorySize, T.smem_max);
        attr_set |= 1ull << (dev & 63);
    }
    dim3 grid(n_tiles, P.n_groups);
    if (!throughput_shape) {
        if (staged) sweep_kernel<true, L.warps, L.blocks_per_sm><<<grid, L.warps * 32, smem_bytes, st>>>(P);
        else sweep_kernel<false, L.warps, L.blocks_per_sm><<<grid, L.warps * 32, smem_bytes, st>>>(P);
    } else {
        if (staged) sweep_kernel<true, T.warps, T.blocks_per_sm><<<grid, T.warps * 32, smem_bytes, st>>>(P);
        else sweep_kernel<false, T.warps, T.blocks_per_sm><<<grid, T.warps * 32, smem_bytes, st>>>(P);
    }
}

// ---- proposal evaluation ---------------------------------------------------------------------

// A group is one warp's work: up to EVAL_G general proposals, or up to EVAL_GS single-parent
// proposals, of one focal kid over one chunk of loci (lanes = consecutive loci, so every row read is
// a coalesced 256 B segment). All row loads of a locus are issued up front and unconditionally, so a
// warp keeps 27 (general) or 30 (singles) independent segments in flight.
//   general: the three contraction forms are all computed and the proposal's kind selects one;
//   singles: a proposal with one given parent x (or none: x = the prior) and an HWE founder as the
//            other is  v = x . (A prior)  with A = kid_matrix(kid stub) symmetric — one row and three
//            FMAs per proposal instead of two rows and twelve, and these are the proposals whose rows
//            come from HBM (every unobserved individual and the prefilter list are single-parent
//            candidates; the pairs re-read the ten best rows from L2). Twice the proposals per warp:
//            twice the useful bytes in flight.
template <int G, bool SINGLE>
__device__ __forceinline__ void eval_group(const EvalParams &P, const EvalGroup &g, int chunk, int lane)
{
    const DevView &D = P.D;
    const int s_beg = chunk * P.chunk_loci;
    const int s_end = min(s_beg + P.chunk_loci, D.Sl);
    const size_t row_elems = (size_t)3 * D.Sp;
    int kind[G];
    const double *rp[G], *rm[SINGLE ? 1 : G];
    LogProd acc[G];
#pragma unroll
    for (int j = 0; j < G; j++) {
        lp_init(acc[j]);
        EvalProp pr{0, ROW_PRIOR, ROW_PRIOR, 0};
        if (j < g.n) pr = P.props[g.begin + j];
        kind[j] = pr.kind;
        rp[j] = D.rows + (size_t)pr.row_p * row_elems;
        if (!SINGLE) rm[j] = D.rows + (size_t)pr.row_m * row_elems;
    }
    const double *kid_row = D.rows + (size_t)g.kid_row * row_elems;
    const double *prior_row = D.rows + (size_t)ROW_PRIOR * row_elems;
    for (int s = s_beg + lane; s < s_end; s += 32) {
        const Tri kid = ld_tri(kid_row, D.Sp, s);
        Tri xp[G], xm[SINGLE ? 1 : G];
#pragma unroll
        for (int j = 0; j < G; j++) { xp[j] = ld_tri(rp[j], D.Sp, s); if (!SINGLE) xm[j] = ld_tri(rm[j], D.Sp, s); }
        const Sym3 A = kid_matrix(kid);
        if (SINGLE) {
            const Tri c = sym_mv(A, ld_tri(prior_row, D.Sp, s));
#pragma unroll
            for (int j = 0; j < G; j++)
                if (j < g.n) lp_mul(acc[j], dot(xp[j], c));
        } else {
#pragma unroll
            for (int j = 0; j < G; j++) {
                const double vt = dot(xm[j], sym_mv(A, xp[j]));         // trio
                const double vs = dot(kid, self_to_kid(xp[j]));        // selfing
                const double vp = dot(xp[j], kid);                      // prong
                const double v = kind[j] == 0 ? vt : kind[j] == 1 ? vs : vp;
                if (j < g.n) lp_mul(acc[j], v);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < G; j++) {
        if (j >= g.n) continue;
        const LogProd r = lp_warp_reduce(acc[j]);
        if (lane == 0) P.partial[(size_t)chunk * P.n_props + g.begin + j] = lp_log(r);
    }
    if (lane == 0) __threadfence();            // one fence behind the warp's partial sums (ncu: a fence per sum was a quarter of the stall time)
}

// singles, two loci per lane: every row read is one 128-bit load per lane (a 512 B segment per warp
// and plane), so half as many load instructions carry the same bytes
__device__ __forceinline__ void eval_singles2(const EvalParams &P, const EvalGroup &g, int chunk, int lane)
{
    constexpr int G = EVAL_GS;
    const DevView &D = P.D;
    const int s_beg = chunk * P.chunk_loci;
    const int s_end = min(s_beg + P.chunk_loci, D.Sl);
    const size_t row_elems = (size_t)3 * D.Sp;
    const double *rp[G];
    LogProd acc[G];
#pragma unroll
    for (int j = 0; j < G; j++) {
        lp_init(acc[j]);
        rp[j] = D.rows + (size_t)(j < g.n ? P.props[g.begin + j].row_p : ROW_PRIOR) * row_elems;
    }
    const double *kid_row = D.rows + (size_t)g.kid_row * row_elems;
    const double *prior_row = D.rows + (size_t)ROW_PRIOR * row_elems;
    auto ld2 = [&](const double *row, int pl, int s) { return *reinterpret_cast<const double2 *>(row + (size_t)pl * D.Sp + s); };
    for (int s = s_beg + 2 * lane; s < s_end; s += 64) {           // s even, Sp a multiple of 64: 16-byte aligned, in bounds
        double2 x[G][3];
#pragma unroll
        for (int j = 0; j < G; j++)
#pragma unroll
            for (int pl = 0; pl < 3; pl++) x[j][pl] = ld2(rp[j], pl, s);
        const double2 k0 = ld2(kid_row, 0, s), k1 = ld2(kid_row, 1, s), k2 = ld2(kid_row, 2, s);
        const double2 q0 = ld2(prior_row, 0, s), q1 = ld2(prior_row, 1, s), q2 = ld2(prior_row, 2, s);
        const Tri ca = sym_mv(kid_matrix(Tri{k0.x, k1.x, k2.x}), Tri{q0.x, q1.x, q2.x});
        const Tri cb = sym_mv(kid_matrix(Tri{k0.y, k1.y, k2.y}), Tri{q0.y, q1.y, q2.y});
        const bool second = s + 1 < s_end;
#pragma unroll
        for (int j = 0; j < G; j++) {
            if (j >= g.n) continue;
            lp_mul(acc[j], dot(Tri{x[j][0].x, x[j][1].x, x[j][2].x}, ca));
            if (second) lp_mul(acc[j], dot(Tri{x[j][0].y, x[j][1].y, x[j][2].y}, cb));
        }
    }
#pragma unroll
    for (int j = 0; j < G; j++) {
        if (j >= g.n) continue;
        const LogProd r = lp_warp_reduce(acc[j]);
        if (lane == 0) P.partial[(size_t)chunk * P.n_props + g.begin + j] = lp_log(r);
    }
    if (lane == 0) __threadfence();            // one fence behind the warp's partial sums (ncu: a fence per sum was a quarter of the stall time)
}

template <bool DEPENDENT>
__device__ __forceinline__ void eval_body(const EvalParams &P)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.y * EVAL_WARPS + warp;
    const int chunk = blockIdx.x;
    EvalGroup g{ROW_PRIOR, 0, 0, 0};
    if (gi < P.n_groups) g = P.groups[gi];
    // launched as a programmatic dependent of the kernel in front of it (the sweep that writes the rows
    // read below): the blocks are resident and their plan is loaded while that kernel drains
    if (DEPENDENT) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (g.single) eval_singles2(P, g, chunk, lane);
    else eval_group<EVAL_G, false>(P, g, chunk, lane);

    // fused fixed-order reduction: the last block of this column of the grid to finish adds the
    // chunks' partial sums in chunk order (deterministic whichever block happens to be last) and
    // scatters them to the caller's proposal order
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&P.counters[blockIdx.y], 1u);
        s_last = done == gridDim.x - 1;
        if (s_last) P.counters[blockIdx.y] = 0;     // ready for the next launch
    }
    __syncthreads();
    if (s_last && threadIdx.x < EVAL_WARPS * EVAL_GS) {
        const int gi2 = blockIdx.y * EVAL_WARPS + threadIdx.x / EVAL_GS, j = threadIdx.x % EVAL_GS;
        if (gi2 < P.n_groups) {
            const EvalGroup g2 = P.groups[gi2];
            if (j < g2.n) {
                const double *src = P.partial + g2.begin + j;
                double sum = 0.0;
#pragma unroll 8
                for (int c = 0; c < (int)gridDim.x; c++) sum += __ldcg(src + (size_t)c * P.n_props);
                P.out[P.props[g2.begin + j].out] = sum;
            }
        }
    }
    if (s_last && P.pub_flag != nullptr) {          // s_last is uniform over the block
        __shared__ bool s_pub;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned t = atomicAdd(P.pub_counter, 1u);
            s_pub = t == gridDim.y - 1;
            if (s_pub) *P.pub_counter = 0;
        }
        __syncthreads();
        if (s_pub) {
            __threadfence();
            for (int i = threadIdx.x; i < P.pub_n; i += blockDim.x) P.pub_dst[i] = __ldcg(P.pub_src + i);
            for (int i = threadIdx.x; i < P.pub_n2; i += blockDim.x) P.pub_dst[P.pub_n + i] = __ldcg(P.pub_src2 + i);
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0) { *reinterpret_cast<volatile unsigned long long *>(P.pub_flag) = P.pub_seq; __threadfence_system(); }
        }
    }
}

__global__ void __launch_bounds__(EVAL_WARPS * 32, 4) eval_kernel(const __grid_constant__ EvalParams P) { eval_body<false>(P); }

// Small proposal lists travel in the kernel's parameter space (up to 28 KB; the limit is 32 764 bytes
// since CUDA 12.1): no copy-engine launch in front of the kernel and no PCIe read inside it. The address
// of a __grid_constant__ parameter is a valid generic pointer to read through.
static bool g_eval_pdl = true;       // PF_PDL=0: plain stream order
void set_eval_pdl(bool on) { g_eval_pdl = on; }

template <int BYTES>
struct EvalInline { EvalParams P; unsigned off_props; alignas(16) unsigned char blob[BYTES]; };

template <int BYTES>
__global__ void __launch_bounds__(EVAL_WARPS * 32, 4) eval_kernel_inl(const __grid_constant__ EvalInline<BYTES> A)
{
    EvalParams P = A.P;
    P.groups = reinterpret_cast<const EvalGroup *>(A.blob);
    P.props = reinterpret_cast<const EvalProp *>(A.blob + A.off_props);
    eval_body<true>(P);
}

template <int BYTES>
static void launch_eval_inl(const EvalParams &P, const void *blob, size_t bytes, size_t off_props, dim3 grid, cudaStream_t st)
{
    static thread_local EvalInline<BYTES> A;           // built in place: 16 KB on the stack of a fiber would be unkind
    A.P = P; A.off_props = (unsigned)off_props;
    memcpy(A.blob, blob, bytes);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(EVAL_WARPS * 32); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_eval_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, eval_kernel_inl<BYTES>, A);
}

bool launch_eval_inline(const EvalParams &P, const void *blob, size_t bytes, size_t off_props, int n_chunks, cudaStream_t st)
{
    dim3 grid(n_chunks, (P.n_groups + EVAL_WARPS - 1) / EVAL_WARPS);
    if (bytes <= 2048) launch_eval_inl<2048>(P, blob, bytes, off_props, grid, st);
    else if (bytes <= 4096) launch_eval_inl<4096>(P, blob, bytes, off_props, grid, st);
    else if (bytes <= 8192) launch_eval_inl<8192>(P, blob, bytes, off_props, grid, st);
    else if (bytes <= 16384) launch_eval_inl<16384>(P, blob, bytes, off_props, grid, st);
    else if (bytes <= EVAL_INLINE_MAX) launch_eval_inl<EVAL_INLINE_MAX>(P, blob, bytes, off_props, grid, st);
    else return false;
    return true;
}

// resident eval blocks on the current device (SM count x blocks per SM from the occupancy
// calculator): run_evals sizes its grid to one full wave of exactly this many blocks
int eval_resident_blocks()
{
    int dev = 0, sms = 148, per_sm = 4;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, eval_kernel, EVAL_WARPS * 32, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    return sms * per_sm;
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

// ---- result delivery -------------------------------------------------------------------------
// Small result vectors go to the host through mapped pinned memory: one block copies them and then
// raises a sequence flag the host spins on. This replaces a device-to-host copy plus a blocking
// stream synchronisation (~10 us per call) by one tiny launch.
__global__ void __launch_bounds__(256) publish_kernel(const double *src, double *dst, int n, unsigned long long *flag, unsigned long long seq)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { *reinterpret_cast<volatile unsigned long long *>(flag) = seq; __threadfence_system(); }
}

// ---- cross-GPU sum of a short result vector over NVLink peer memory -----------------------------
// Loci are sharded over `world` GPUs (one process each); a Gibbs step ends with the per-proposal /
// per-component sums of every shard added up. The vectors are a few KB, so the exchange is pure
// latency: ONE kernel on the engine's stream writes this rank's partial vector straight into every
// peer's mailbox (peer-to-peer stores through NVSwitch), raises a sequence flag there, waits for the
// flags of the other ranks in its own mailbox, adds the `world` vectors in rank order — the same
// order on every rank, so all ranks hold bit-identical results and their host logic stays in lock
// step — and publishes the sums to the host. No NCCL call, no extra launch, no host round trip.
// Mailboxes are double-buffered by the parity of the sequence number: a rank can be at most one
// exchange ahead of a peer (it cannot finish exchange s + 1 before that peer has sent s + 1, which
// the peer does only after it has read everything of exchange s).
__global__ void __launch_bounds__(1024) xreduce_kernel(const __grid_constant__ XReduceParams X)
{
    const int tid = threadIdx.x;
    for (int c0 = 0, round = 0; c0 < X.n; c0 += XREDUCE_MAX, round++) {
        const unsigned long long seq = X.seq0 + (unsigned long long)round;
        const int par = (int)(seq & 1ull), cnt = min(XREDUCE_MAX, X.n - c0);
        for (int p = 0; p < X.world; p++) {
            double *dst = X.peer_mail[p] + ((size_t)par * X.world + X.rank) * XREDUCE_MAX;
            for (int i = tid; i < cnt; i += blockDim.x) dst[i] = X.src[c0 + i];
        }
        __threadfence_system();
        __syncthreads();
        if (tid < X.world) {
            *reinterpret_cast<volatile unsigned long long *>(X.peer_flags[tid] + X.rank) = seq;
            __threadfence_system();
            volatile unsigned long long *mine = X.peer_flags[X.rank] + tid;
            const long long t0 = clock64();
            while (*mine < seq)
                if (clock64() - t0 > (20ll << 30)) __trap();      // a peer died: ~10 s at 2 GHz, then fail loudly
        }
        __syncthreads();
        __threadfence_system();
        const double *in = X.peer_mail[X.rank] + (size_t)par * X.world * XREDUCE_MAX;
        for (int i = tid; i < cnt; i += blockDim.x) {
            double acc = 0.0;
            for (int q = 0; q < X.world; q++) acc += *reinterpret_cast<const volatile double *>(in + (size_t)q * XREDUCE_MAX + i);
            X.dst[c0 + i] = acc;
            if (X.host_dst) X.host_dst[c0 + i] = acc;
        }
        __syncthreads();
    }
    if (X.host_flag) {
        __threadfence_system();
        __syncthreads();
        if (tid == 0) { *reinterpret_cast<volatile unsigned long long *>(X.host_flag) = X.host_seq; __threadfence_system(); }
    }
}

void launch_xreduce(const XReduceParams &X, cudaStream_t st)
{
    xreduce_kernel<<<1, 1024, 0, st>>>(X);
}

void launch_publish(const double *src, double *dst_mapped, int n, unsigned long long *flag_mapped, unsigned long long seq, cudaStream_t st)
{
    publish_kernel<<<1, 256, 0, st>>>(src, dst_mapped, n, flag_mapped, seq);
}

void launch_reduce(const ReduceParams &P, cudaStream_t st)
{
    reduce_kernel<<<(P.n * 32 + 127) / 128, 128, 0, st>>>(P);
}

} // namespace pf
