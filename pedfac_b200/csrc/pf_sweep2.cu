// pf_sweep2.cu — second-generation sum-product sweep for sm_100a: compiled task records with
// pre-resolved shared-memory operands (format and rationale: pf_sweep2.cuh). All arithmetic fp64.
#include "pf_sweep2.cuh"

namespace pf {

namespace {

constexpr int ROW_D = 3 * TILE_LOCI;            // doubles per shared-memory row

// Genotype format of a launch. When every individual of the launch carries hard calls (or none), or
// 16-bit numerators over one common denominator (or none), the kernel is compiled for that format:
// an operand then holds the byte offset of the individual's staged tile itself (an individual without
// data points at a constant "missing" tile), decoding is straight-line code, and the compiler can
// issue the loads of all operands of a task before any arithmetic. Mixed launches use FMT_GENERIC,
// which looks the individual up in the Emis2 table and branches on its kind.
enum Fmt2 : int { FMT_GENERIC = 0, FMT_HARD = 1, FMT_DEC16 = 2 };

struct Ctx2 {
    const uint32_t *pool;     // group's record words (shared memory)
    const Emis2 *emis;        // group's genotype entries (shared memory; FMT_GENERIC only)
    const char *edata_lane;   // staged genotype tiles, + this lane's offset inside a hard / dec16 tile
    const char *edata;        // staged genotype tiles
    double *rows_lane;        // shared-memory rows + lane
    double *persist_lane;     // persistent rows + locus
    Tri prior;                // HWE prior of this lane's locus
    double lo, hi;            // eps / 2, 1 - eps of this lane's locus
    double den, rcp;          // FMT_DEC16: the common denominator and RN(1 / den)
    size_t plane;             // Sp
    int lane, shift;          // shift: bit position of this lane's 2-bit code inside its word
    bool no_persist;          // measurement aid
};

__device__ __forceinline__ Tri ld2(const Ctx2 &C, uint32_t id)
{
    const double *p = C.rows_lane + id * ROW_D;
    return Tri{p[0], p[TILE_LOCI], p[2 * TILE_LOCI]};
}
__device__ __forceinline__ void st2(const Ctx2 &C, uint32_t id, Tri v)
{
    double *p = C.rows_lane + id * ROW_D;
    p[0] = v.x; p[TILE_LOCI] = v.y; p[2 * TILE_LOCI] = v.z;
}
__device__ __forceinline__ Sym3 ldpair(const Ctx2 &C, uint32_t id)
{
    const Tri a = ld2(C, id), b = ld2(C, id + 1);
    return Sym3{a.x, a.y, a.z, b.x, b.y, b.z};
}
__device__ __forceinline__ void stpair(const Ctx2 &C, uint32_t id, const Sym3 &P)
{
    st2(C, id, Tri{P.a00, P.a01, P.a02});
    st2(C, id + 1, Tri{P.a11, P.a12, P.a22});
}
__device__ __forceinline__ void persist(const Ctx2 &C, uint32_t row, Tri v)
{
    if (C.no_persist) return;
    double *p = C.persist_lane + (size_t)row * 3 * C.plane;
    p[0] = v.x; p[C.plane] = v.y; p[2 * C.plane] = v.z;
}

// local factor of an individual from its staged genotype tile
// (_ReadPedfile@0x10000b1ca-0x10000b6e2 emission model; _PassingOFtoIV@0x10001c2d0 founder prior)
template <int FMT>
__device__ __forceinline__ Tri decode2(const Ctx2 &C, uint32_t w0)
{
    Tri v{1.0, 1.0, 1.0};
    if (FMT == FMT_HARD) {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(C.edata_lane + (w0 & CMP_OFF_MASK));
        const int c = (w >> C.shift) & 3;
        v = Tri{c == 0 ? C.hi : C.lo, c == 1 ? C.hi : C.lo, c == 2 ? C.hi : C.lo};
        if (c == 3) v = Tri{1.0, 1.0, 1.0};
    } else if (FMT == FMT_DEC16) {
        const uint16_t *u = reinterpret_cast<const uint16_t *>(C.edata_lane + (w0 & CMP_OFF_MASK));
        const unsigned u0 = u[0], u1 = u[TILE_LOCI], u2 = u[2 * TILE_LOCI];
        v = Tri{div_exact((double)u0, C.den, C.rcp), div_exact((double)u1, C.den, C.rcp), div_exact((double)u2, C.den, C.rcp)};
    } else {
        const uint32_t kind = w0 & 7u;
        if (kind != EMIS_NONE) {
            const Emis2 E = C.emis[(w0 & CMP_OFF_MASK) >> 3];
            const char *p = C.edata + E.smem_off;
            const int l = C.lane;
            if (kind == EMIS_HARD) {
                const uint32_t w = reinterpret_cast<const uint32_t *>(p)[l >> 4];
                const int c = (w >> ((l & 15) * 2)) & 3;
                if (c != 3) v = Tri{c == 0 ? C.hi : C.lo, c == 1 ? C.hi : C.lo, c == 2 ? C.hi : C.lo};
            } else if (kind == EMIS_DEC16) {
                const uint16_t *u = reinterpret_cast<const uint16_t *>(p);
                const double den = (double)E.denom, rcp = E.rcp;
                v = Tri{div_exact((double)u[l], den, rcp), div_exact((double)u[TILE_LOCI + l], den, rcp), div_exact((double)u[2 * TILE_LOCI + l], den, rcp)};
            } else if (kind == EMIS_F32) {
                const float *f = reinterpret_cast<const float *>(p);
                v = Tri{(double)f[l], (double)f[TILE_LOCI + l], (double)f[2 * TILE_LOCI + l]};
            } else {
                const double *f = reinterpret_cast<const double *>(p);
                v = Tri{f[l], f[TILE_LOCI + l], f[2 * TILE_LOCI + l]};
            }
        }
    }
    const bool founder = (w0 & CMP_FOUNDER) != 0;
    return mul(v, Tri{founder ? C.prior.x : 1.0, founder ? C.prior.y : 1.0, founder ? C.prior.z : 1.0});
}

// message of an individual into one of its marriages: local factor x the messages of its other
// marriages (_PassingIVtoIF@0x10001c490). Absent rows are row 0 (ones) and the prior of a
// non-founder is a select: no branch (measured: warp-uniform tests that skip the rows of a leaf and
// the prior of a non-founder made every level 15-40 % slower — a block's warps are alone on their
// schedulers and pay the full branch latency). G (general records only): the operand may name more
// than three rows.
template <int FMT, bool G>
__device__ __forceinline__ Tri cmp2(const Ctx2 &C, const uint32_t *c)
{
    const uint32_t w0 = c[0], w1 = c[1], w2 = c[2];
    const Tri r0 = ld2(C, w1 & 0xffffu), r1 = ld2(C, w1 >> 16), r2 = ld2(C, w2 & 0xffffu);
    Tri v = decode2<FMT>(C, w0);
    v = mul(mul(v, r0), mul(r1, r2));
    if (G && (w2 >> 16)) {                                   // an individual with more than three other marriages
        const uint32_t *x = C.pool + (w2 >> 16);
        const uint32_t n = x[0];
#pragma unroll 1
        for (uint32_t j = 0; j < n; j++) v = mul(v, ld2(C, x[1 + j]));
    }
    return v;
}

__device__ __forceinline__ uint32_t pair_id(const uint32_t *w, int j) { return (w[j >> 1] >> ((j & 1) * 16)) & 0xffffu; }

// product of the chunk pairs of a large sibship, `skip` (if >= 0) left out
__device__ __forceinline__ void mul_pairs(const Ctx2 &C, Sym3 &P, const uint32_t *w, int np, int skip)
{
#pragma unroll 1
    for (int j = 0; j < np; j++)
        if (j != skip) sym_mul(P, ldpair(C, pair_id(w, j)));
}

// Task bodies. One instance per number of kids handled in registers (NK) and per G: G = false is the
// common record (no selfing, every operand within three rows: `selfing` folds to 0 and the extra-row
// test disappears); G = true records carry T2_GENERAL in their header.

// ROOT2 (_UpdateProb@0x10001da20: Z = sum_g belief, log Z summed over loci is the component's LL)
template <int FMT, bool G>
__device__ __forceinline__ void task2_root(const Sweep2Params &P, const Ctx2 &C, const uint32_t *r, int tile, bool in_range)
{
    const Tri v = cmp2<FMT, G>(C, r + 4);
    if ((r[0] >> 8) & 1u) persist(C, r[1], v); else st2(C, r[3], v);
    LogProd z; lp_init(z);
    if (in_range) lp_mul(z, (v.x + v.y) + v.z);
    z = lp_warp_reduce(z);
    if (C.lane == 0) { P.partial[(size_t)tile * P.n_out + r[2]] = lp_log(z); __threadfence(); }
}

// FAMUP2: message of a marriage towards the root
template <int FMT, int NK, bool G>
__device__ __forceinline__ void task2_famup(const Ctx2 &C, const uint32_t *r)
{
    const uint32_t w0 = r[0];
    const int selfing = G ? (w0 >> 8) & 1 : 0, urole = (w0 >> 9) & 3, np = w0 >> 24;
    const Tri a = cmp2<FMT, G>(C, r + 2), b = cmp2<FMT, G>(C, r + 2 + CMP_WORDS);
    const uint32_t *k = r + 2 + 2 * CMP_WORDS;
    Sym3 Pall = sym_ones();
#pragma unroll
    for (int j = 0; j < NK; j++) sym_mul(Pall, kid_matrix(cmp2<FMT, G>(C, k + j * CMP_WORDS)));
    if (np) mul_pairs(C, Pall, k + NK * CMP_WORDS, np, -1);
    Tri out;
    if (urole == 2) out = fam_to_kid(Pall, a, b, selfing);
    else out = fam_to_pa(Pall, urole == 0 ? b : a, urole == 0 ? selfing : 0);
    st2(C, r[1] & 0xffffu, out);
}

// KFOLD2: product of the kid factors of one chunk of a large sibship (a pair of rows)
template <int FMT, int NK, bool G>
__device__ __forceinline__ void task2_kfold(const Ctx2 &C, const uint32_t *r)
{
    Sym3 Pk = sym_ones();
#pragma unroll
    for (int j = 0; j < NK; j++) sym_mul(Pk, kid_matrix(cmp2<FMT, G>(C, r + 2 + j * CMP_WORDS)));
    stpair(C, r[1] & 0xffffu, Pk);
}

// messages and stubs of NK kids (entries of CMP_WORDS + 2 words at k: cmp, stub row, message row)
// given a, b and the product Pout of every factor outside this set; returns the product of all
// factors. The kids' factors stay in registers (NK is a template parameter: indexed arrays would go
// to local memory). Exclusive products: suffixes first, the prefix runs along. A kid without a
// message row of its own stores the message into row 1, which nobody reads.
// (_PassingIFtoIV@0x10001c8f0 towards the kids; stub = _FillInStubs@0x10001e1a0)
template <int FMT, int NK, bool G>
__device__ __forceinline__ Sym3 kids_down(const Ctx2 &C, const uint32_t *k, Tri a, Tri b, const Sym3 &Pout, int selfing)
{
    Tri km[NK > 0 ? NK : 1];
    Sym3 K[NK > 0 ? NK : 1];
#pragma unroll
    for (int j = 0; j < NK; j++) { km[j] = cmp2<FMT, G>(C, k + j * (CMP_WORDS + 2)); K[j] = kid_matrix(km[j]); }
    Sym3 suf[NK > 1 ? NK : 1];
    suf[NK > 1 ? NK - 1 : 0] = sym_ones();
#pragma unroll
    for (int j = NK - 2; j >= 0; j--) { suf[j] = suf[j + 1]; sym_mul(suf[j], K[j + 1]); }
    Sym3 pre = Pout;
#pragma unroll
    for (int j = 0; j < NK; j++) {
        Sym3 Pm = pre;
        if (NK > 1) sym_mul(Pm, suf[j]);
        const Tri dn = fam_to_kid(Pm, a, b, selfing);
        const uint32_t *e = k + j * (CMP_WORDS + 2);
        persist(C, e[CMP_WORDS], mul(km[j], dn));
        st2(C, e[CMP_WORDS + 1] & 0xffffu, dn);
        sym_mul(pre, K[j]);
    }
    return pre;
}

// FAMDOWN2: messages away from the root through one marriage: its prong, the stubs and messages of
// the parents below it and of the kids it handles itself (_PassingIFtoIV@0x10001c8f0 away from the
// root, _FillInProngs@0x10001dd70, _FillInStubs@0x10001e1a0)
template <int FMT, int NK, bool G>
__device__ __forceinline__ void task2_famdown(const Ctx2 &C, const uint32_t *r)
{
    const uint32_t w0 = r[0];
    const int selfing = G ? (w0 >> 8) & 1 : 0, urole = (w0 >> 9) & 3, np = w0 >> 24;
    const Tri in_u = cmp2<FMT, G>(C, r + 2);
    const Tri ca = cmp2<FMT, G>(C, r + 5), cb = cmp2<FMT, G>(C, r + 9);
    const uint32_t *k = r + 14;
    const Tri a = urole == 0 ? in_u : ca, b = urole == 1 ? in_u : cb;
    Sym3 Pout = sym_ones();                                        // the up individual's own factor when it is a kid here
    if (urole == 2) Pout = kid_matrix(in_u);
    if (np) mul_pairs(C, Pout, k + NK * (CMP_WORDS + 2), np, -1);  // the chunks of a large sibship
    const Sym3 Pall = kids_down<FMT, NK, G>(C, k, a, b, Pout, selfing);
    persist(C, r[1], fam_to_kid(Pall, a, b, selfing));             // prong
    const uint32_t dnr = r[13];
    if (w0 & T2_HAS_A) {
        const Tri dn = fam_to_pa(Pall, b, selfing);
        persist(C, r[8], mul(ca, dn));
        st2(C, dnr & 0xffffu, dn);
    }
    if (w0 & T2_HAS_B) {
        const Tri dn = fam_to_pa(Pall, a, 0);
        persist(C, r[12], mul(cb, dn));
        st2(C, dnr >> 16, dn);
    }
}

// KIDCHUNK2: the kids of one chunk of a large sibship, at the level of the marriage's FAMDOWN2 and
// independent of it: it composes the parents' messages itself and multiplies the factors of the
// kids FAMDOWN2 handles (ni of them, as cmps) and the other chunks' pairs (chunk `self` left out)
template <int FMT, int NK, bool G>
__device__ __forceinline__ void task2_kidchunk(const Ctx2 &C, const uint32_t *r)
{
    const uint32_t w0 = r[0], w1 = r[1];
    const int selfing = G ? (w0 >> 8) & 1 : 0, urole = (w0 >> 9) & 3, np = w0 >> 24;
    const int ni = w1 & 0xff, self = (w1 >> 8) & 0xff;
    const Tri in_u = cmp2<FMT, G>(C, r + 2);
    const Tri ca = cmp2<FMT, G>(C, r + 5), cb = cmp2<FMT, G>(C, r + 8);
    const Tri a = urole == 0 ? in_u : ca, b = urole == 1 ? in_u : cb;
    const uint32_t *k = r + 11;
    const uint32_t *oi = k + NK * (CMP_WORDS + 2);
    Sym3 Pout = sym_ones();
    if (urole == 2) Pout = kid_matrix(in_u);
#pragma unroll 1
    for (int j = 0; j < ni; j++) sym_mul(Pout, kid_matrix(cmp2<FMT, G>(C, oi + j * CMP_WORDS)));
    mul_pairs(C, Pout, oi + ni * CMP_WORDS, np, self);
    kids_down<FMT, NK, G>(C, k, a, b, Pout, selfing);
}

template <int FMT, bool G>
__device__ __forceinline__ void run_task2g(const Sweep2Params &P, const Ctx2 &C, const uint32_t *r, uint32_t w0, int tile, bool in_range)
{
    const int nk = (w0 >> 16) & 0xff;
    switch (w0 & 0xffu) {
    case T2_FAMUP:
        switch (nk) {
        case 0: task2_famup<FMT, 0, G>(C, r); break;
        case 1: task2_famup<FMT, 1, G>(C, r); break;
        case 2: task2_famup<FMT, 2, G>(C, r); break;
        case 3: task2_famup<FMT, 3, G>(C, r); break;
        default: task2_famup<FMT, 4, G>(C, r); break;
        }
        break;
    case T2_FAMDOWN:
        switch (nk) {
        case 0: task2_famdown<FMT, 0, G>(C, r); break;
        case 1: task2_famdown<FMT, 1, G>(C, r); break;
        case 2: task2_famdown<FMT, 2, G>(C, r); break;
        case 3: task2_famdown<FMT, 3, G>(C, r); break;
        default: task2_famdown<FMT, 4, G>(C, r); break;
        }
        break;
    case T2_KFOLD:
        switch (nk) {
        case 1: task2_kfold<FMT, 1, G>(C, r); break;
        case 2: task2_kfold<FMT, 2, G>(C, r); break;
        default: task2_kfold<FMT, 3, G>(C, r); break;
        }
        break;
    case T2_KIDCHUNK:
        switch (nk) {
        case 1: task2_kidchunk<FMT, 1, G>(C, r); break;
        case 2: task2_kidchunk<FMT, 2, G>(C, r); break;
        default: task2_kidchunk<FMT, 3, G>(C, r); break;
        }
        break;
    default: task2_root<FMT, G>(P, C, r, tile, in_range); break;
    }
}

template <int FMT>
__device__ __forceinline__ void run_task2(const Sweep2Params &P, const Ctx2 &C, const uint32_t *r, int tile, bool in_range)
{
    const uint32_t w0 = r[0];
    if (w0 & T2_GENERAL) run_task2g<FMT, true>(P, C, r, w0, tile, in_range);
    else run_task2g<FMT, false>(P, C, r, w0, tile, in_range);
}

// ---- asynchronous staging of the genotype tiles ------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// bulk asynchronous copy (the TMA engine's 1-D form): global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

} // namespace

template <int FMT, int WARPS, int MIN_BLOCKS>
__global__ void __launch_bounds__(WARPS * 32, MIN_BLOCKS) sweep2_kernel(const __grid_constant__ Sweep2Params P)
{
    extern __shared__ __align__(16) char smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    // a proposal kernel queued behind this sweep as a programmatic dependent may become resident now; it
    // waits (griddepcontrol.wait) for this grid to complete before it reads a row
    asm volatile("griddepcontrol.launch_dependents;");
    const DevView &D = P.D;
    const int tile = blockIdx.x;
    const Sweep2Group G = P.groups[blockIdx.y];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NTHR = WARPS * 32;
    const Sweep2Smem L = sweep2_smem_layout(G, P.n_levels);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);

    if (tid == 0) {
        mbar_init(bar, WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- stage the plan slice; row 0 = ones ----
    int *s_loff = reinterpret_cast<int *>(smem + L.loff);
    for (int i = tid; i <= P.n_levels; i += NTHR) s_loff[i] = P.level_off[(size_t)blockIdx.y * (P.n_levels + 1) + i] - G.task_begin;
    {
        uint32_t *d = reinterpret_cast<uint32_t *>(smem + L.tasks);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(P.tasks + G.task_begin);
        for (int i = tid; i < G.task_end - G.task_begin; i += NTHR) d[i] = src[i];
        d = reinterpret_cast<uint32_t *>(smem + L.pool); src = P.pool + G.pool_begin;
        for (int i = tid; i < G.pool_len; i += NTHR) d[i] = src[i];
        if (FMT == FMT_GENERIC) {
            uint4 *d4 = reinterpret_cast<uint4 *>(smem + L.emis);
            const uint4 *s4 = reinterpret_cast<const uint4 *>(P.emis + G.emis_begin);
            for (int i = tid; i < G.emis_count; i += NTHR) d4[i] = s4[i];
        }
        double *ones = reinterpret_cast<double *>(smem + L.rows);
        if (tid < ROW_D) ones[tid] = 1.0;
        // tile 0 of the staged genotype data: "no data" (all codes 3 / every numerator = the denominator)
        if (FMT == FMT_HARD && tid < 2) reinterpret_cast<uint32_t *>(smem + L.edata)[tid] = 0xffffffffu;
        if (FMT == FMT_DEC16 && tid < ROW_D) reinterpret_cast<uint16_t *>(smem + L.edata)[tid] = (uint16_t)P.dec16_denom;
    }
    __syncthreads();                                         // the barrier is initialised before anyone uses it
    // ---- genotype tiles: lane j of a warp owns one individual of a batch of 32; soft-call planes go
    // through the bulk-copy engine (64 / 128 / 256 B each, completion counted on the mbarrier), the 8
    // bytes of a hard-call tile through cp.async; everything is in flight before anyone waits ----
    {
        const unsigned edata = (unsigned)__cvta_generic_to_shared(smem + L.edata);
        const size_t Sp = D.Sp;
        for (int i0 = warp * 32; i0 < G.emis_count; i0 += WARPS * 32) {
            const int i = i0 + lane;
            Stage2 S{0ull, EMIS_NONE, 0u};
            if (i < G.emis_count) S = P.stage[G.emis_begin + i];
            const unsigned esz = S.kind == EMIS_DEC16 ? 2u : S.kind == EMIS_F32 ? 4u : S.kind == EMIS_F64 ? 8u : 0u;
            unsigned bytes = 3u * TILE_LOCI * esz, total = bytes;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            if (lane == 0 && total) mbar_expect_tx(bar, total);
            __syncwarp();
            const char *gp = reinterpret_cast<const char *>(S.ptr);
            if (esz) {
#pragma unroll
                for (int pl = 0; pl < 3; pl++)
                    bulk_g2s(edata + S.smem_off + pl * TILE_LOCI * esz, gp + ((size_t)pl * Sp + (size_t)tile * TILE_LOCI) * esz, TILE_LOCI * esz, bar);
            } else if (S.kind == EMIS_HARD) {
                cp_async8(edata + S.smem_off, gp + (size_t)tile * 8);
            }
        }
        if (lane == 0) mbar_arrive(bar);
        asm volatile("cp.async.wait_all;" ::: "memory");
        mbar_wait(bar, 0);
    }
    __syncthreads();

    Ctx2 C;
    C.pool = reinterpret_cast<const uint32_t *>(smem + L.pool);
    C.emis = reinterpret_cast<const Emis2 *>(smem + L.emis);
    C.edata = smem + L.edata;
    C.edata_lane = smem + L.edata + (FMT == FMT_HARD ? 4 * (lane >> 4) : 2 * lane);
    C.shift = (lane & 15) * 2;
    C.den = (double)P.dec16_denom; C.rcp = P.dec16_rcp;
    C.lane = lane;
    C.no_persist = (P.debug_flags & 1) != 0;
    C.plane = (size_t)D.Sp;
    const int s = tile * TILE_LOCI + lane;
    C.rows_lane = reinterpret_cast<double *>(smem + L.rows) + lane;
    C.persist_lane = D.rows + s;
    C.prior = ld_tri(D.rows + (size_t)ROW_PRIOR * 3 * D.Sp, D.Sp, s);
    {
        const double ep = D.eps[s];
        C.lo = ep / 2.0; C.hi = 1.0 - ep;
    }
    const bool in_range = s < D.Sl;
    const uint32_t *tasks = reinterpret_cast<const uint32_t *>(smem + L.tasks);

    const bool clk = P.clocks != nullptr && blockIdx.x == 0 && (int)blockIdx.y == P.clock_group && tid == 0;
    if (clk) P.clocks[0] = clock64();
    for (int lev = 0; lev < P.n_levels; lev++) {
        const int beg = s_loff[lev], end = s_loff[lev + 1];
        for (int ti = beg + warp; ti < end; ti += WARPS) run_task2<FMT>(P, C, C.pool + tasks[ti], tile, in_range);
        if (lev + 1 < P.n_levels) __syncthreads();
        if (clk) P.clocks[lev + 1] = clock64();
    }
    // fused fixed-order reduction of the per-tile log Z sums: the last block to finish adds the tiles
    // exactly as reduce_kernel does (lanes take tiles lane, lane + 32, ..., then an xor butterfly)
    if (P.out != nullptr) {
        __shared__ bool s_last;
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned done = atomicAdd(P.counter, 1u);
            s_last = done == gridDim.x * gridDim.y - 1;
            if (s_last) *P.counter = 0;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            const int n_parts = gridDim.x;
            for (int i = warp; i < P.n_out; i += WARPS) {
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

template <int FMT>
static void launch_fmt(const Sweep2Params &P, dim3 grid, int smem_bytes, cudaStream_t st)
{
    static unsigned long long attr_set = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_set >> (dev & 63) & 1)) {
        cudaFuncSetAttribute(sweep2_kernel<FMT, SWEEP2_WARPS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP2_SMEM_MAX);
        cudaFuncSetAttribute(sweep2_kernel<FMT, SWEEP2_WARPS_SMALL, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP2_SMEM_SMALL);
        attr_set |= 1ull << (dev & 63);
    }
    // Register-resident task bodies need ~170 registers; anything less spills, and with shared memory
    // taking the L1 a spill is an L2 round trip. So: a block of a large group has the SM to itself
    // (8 warps), blocks of small groups have 4 warps and share an SM three at a time.
    if (smem_bytes <= SWEEP2_SMEM_SMALL) sweep2_kernel<FMT, SWEEP2_WARPS_SMALL, 3><<<grid, SWEEP2_WARPS_SMALL * 32, smem_bytes, st>>>(P);
    else sweep2_kernel<FMT, SWEEP2_WARPS, 1><<<grid, SWEEP2_WARPS * 32, smem_bytes, st>>>(P);
}

void launch_sweep2(const Sweep2Params &P, int n_tiles, int smem_bytes, cudaStream_t st)
{
    dim3 grid(n_tiles, P.n_groups);
    if (P.format == FMT_HARD) launch_fmt<FMT_HARD>(P, grid, smem_bytes, st);
    else if (P.format == FMT_DEC16) launch_fmt<FMT_DEC16>(P, grid, smem_bytes, st);
    else launch_fmt<FMT_GENERIC>(P, grid, smem_bytes, st);
}

} // namespace pf
