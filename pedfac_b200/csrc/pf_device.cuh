// pf_device.cuh — device-side data model and arithmetic helpers of libpedfac_cuda (sm_100a).
//
// HBM layout (one engine = one GPU = one locus shard of Sl loci, padded to Sp):
//   rows      fp64 [n_rows][3][Sp]   SoA planes per genotype state, locus contiguous, so a warp
//                                    reads 32 consecutive loci of one plane as one 256 B segment.
//             row 0 = HWE founder prior, row 1 = all ones, then per chain: one stub row per
//             individual slot and one prong row per marriage slot (24 B per (node, locus)).
//   hard      u32  [cap_indiv][Sp/16] 2-bit genotype codes, 16 loci per word (0.25 B per call)
//   soft      per-slot SoA planes [3][Sp] of fp32 (12 B), fp64 (24 B) or u16 numerators (6 B)
//   scratch   fp64 rows of the same shape for the transient per-edge messages of one sweep
//             (the reference persists 2 x 24 B per edge and locus on every mnode; we do not).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pf {

enum EmisKind : uint8_t { EMIS_NONE = 0, EMIS_HARD = 1, EMIS_F32 = 2, EMIS_F64 = 3, EMIS_DEC16 = 4 };

constexpr int ROW_PRIOR = 0;
constexpr int ROW_ONES = 1;
constexpr int ROW_CHAIN0 = 2;

struct DevView {
    double *rows;            // persistent rows
    double *scratch;         // transient message rows
    const double *eps;       // [Sp]
    const uint8_t *ekind;    // [cap_indiv]
    const void *const *eptr; // [cap_indiv] device pointer to the slot's genotype row
    const uint32_t *edenom;  // [cap_indiv] denominator of EMIS_DEC16 rows
    int Sl, Sp;
};

// Row references inside plans: bit 30 set = scratch row, else persistent row.
constexpr int SCRATCH_BIT = 1 << 30;

__device__ __forceinline__ double *row_ptr(const DevView &D, int ref)
{
    return (ref & SCRATCH_BIT) ? D.scratch + (size_t)(ref & ~SCRATCH_BIT) * 3 * D.Sp
                               : D.rows + (size_t)ref * 3 * D.Sp;
}

struct Tri { double x, y, z; };

__device__ __forceinline__ Tri ld_tri(const double *row, int Sp, int s)
{
    return Tri{row[s], row[Sp + s], row[2 * Sp + s]};
}
__device__ __forceinline__ void st_tri(double *row, int Sp, int s, Tri v)
{
    row[s] = v.x; row[Sp + s] = v.y; row[2 * Sp + s] = v.z;
}
__device__ __forceinline__ Tri mul(Tri a, Tri b) { return Tri{a.x * b.x, a.y * b.y, a.z * b.z}; }

// Emission of individual `slot` at local locus s (_ReadPedfile@0x10000b1ca-0x10000b6e2):
// missing / unobserved -> 1; hard call c -> eps/2 everywhere, 1-eps at c; soft -> the triple.
__device__ __forceinline__ Tri emission(const DevView &D, int slot, int s)
{
    const uint8_t k = D.ekind[slot];
    if (k == EMIS_NONE) return Tri{1.0, 1.0, 1.0};
    const void *p = D.eptr[slot];
    if (k == EMIS_HARD) {
        const uint32_t w = static_cast<const uint32_t *>(p)[s >> 4];
        const int c = (w >> ((s & 15) * 2)) & 3;
        if (c == 3) return Tri{1.0, 1.0, 1.0};
        const double e = D.eps[s];
        const double lo = e / 2.0, hi = 1.0 - e;
        return Tri{c == 0 ? hi : lo, c == 1 ? hi : lo, c == 2 ? hi : lo};
    }
    if (k == EMIS_F32) {
        const float *f = static_cast<const float *>(p);
        return Tri{(double)f[s], (double)f[D.Sp + s], (double)f[2 * D.Sp + s]};
    }
    if (k == EMIS_F64) {
        const double *f = static_cast<const double *>(p);
        return Tri{f[s], f[D.Sp + s], f[2 * D.Sp + s]};
    }
    const uint16_t *u = static_cast<const uint16_t *>(p);
    const double den = (double)D.edenom[slot];
    return Tri{(double)u[s] / den, (double)u[D.Sp + s] / den, (double)u[2 * D.Sp + s] / den};
}

// local(i) = emission x (founder ? HWE prior : 1)  (_PassingOFtoIV@0x10001c2d0)
__device__ __forceinline__ Tri local_factor(const DevView &D, int slot, int founder, int s)
{
    Tri e = emission(D, slot, s);
    if (founder) e = mul(e, ld_tri(D.rows + (size_t)ROW_PRIOR * 3 * D.Sp, D.Sp, s));
    return e;
}

// ---- Mendelian transmission (_TPROB@0x100035b10), T[gk][gm][gp], symmetric in the parents ----
// K(gm,gp) = sum_gk T[gk][gm][gp] * k[gk]: 6 distinct values.
struct Sym3 { double a00, a01, a02, a11, a12, a22; };

__device__ __forceinline__ Sym3 kid_matrix(Tri k)
{
    Sym3 A;
    A.a00 = k.x;
    A.a01 = 0.5 * k.x + 0.5 * k.y;
    A.a02 = k.y;
    A.a11 = (0.25 * k.x + 0.5 * k.y) + 0.25 * k.z;
    A.a12 = 0.5 * k.y + 0.5 * k.z;
    A.a22 = k.z;
    return A;
}
// y[g] = sum_h A[g][h] x[h]
__device__ __forceinline__ Tri sym_mv(const Sym3 &A, Tri x)
{
    return Tri{A.a00 * x.x + A.a01 * x.y + A.a02 * x.z,
               A.a01 * x.x + A.a11 * x.y + A.a12 * x.z,
               A.a02 * x.x + A.a12 * x.y + A.a22 * x.z};
}
__device__ __forceinline__ double dot(Tri a, Tri b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// W[gm][gp] as a full 3x3 (parents' own factors break the symmetry)
struct Mat3 { double m[9]; };   // index gm*3+gp

__device__ __forceinline__ void mat_mul_kid(Mat3 &W, Tri k)
{
    const Sym3 K = kid_matrix(k);
    W.m[0] *= K.a00; W.m[1] *= K.a01; W.m[2] *= K.a02;
    W.m[3] *= K.a01; W.m[4] *= K.a11; W.m[5] *= K.a12;
    W.m[6] *= K.a02; W.m[7] *= K.a12; W.m[8] *= K.a22;
}
__device__ __forceinline__ void mat_mul_pa(Mat3 &W, Tri p)   // column factor: depends on gp
{
    W.m[0] *= p.x; W.m[1] *= p.y; W.m[2] *= p.z;
    W.m[3] *= p.x; W.m[4] *= p.y; W.m[5] *= p.z;
    W.m[6] *= p.x; W.m[7] *= p.y; W.m[8] *= p.z;
}
__device__ __forceinline__ void mat_mul_ma(Mat3 &W, Tri q)   // row factor: depends on gm
{
    W.m[0] *= q.x; W.m[1] *= q.x; W.m[2] *= q.x;
    W.m[3] *= q.y; W.m[4] *= q.y; W.m[5] *= q.y;
    W.m[6] *= q.z; W.m[7] *= q.z; W.m[8] *= q.z;
}
// message to the father: sum over gm; to the mother: sum over gp; to a kid: contract with T.
__device__ __forceinline__ Tri mat_to_pa(const Mat3 &W) { return Tri{W.m[0] + W.m[3] + W.m[6], W.m[1] + W.m[4] + W.m[7], W.m[2] + W.m[5] + W.m[8]}; }
__device__ __forceinline__ Tri mat_to_ma(const Mat3 &W) { return Tri{W.m[0] + W.m[1] + W.m[2], W.m[3] + W.m[4] + W.m[5], W.m[6] + W.m[7] + W.m[8]}; }
__device__ __forceinline__ Tri mat_to_kid(const Mat3 &W)
{
    // out[gk] = sum_{gm,gp} T[gk][gm][gp] W[gm][gp]
    const double s01 = W.m[1] + W.m[3], s02 = W.m[2] + W.m[6], s12 = W.m[5] + W.m[7];
    return Tri{W.m[0] + 0.5 * s01 + 0.25 * W.m[4],
               0.5 * s01 + s02 + 0.5 * W.m[4] + 0.5 * s12,
               0.25 * W.m[4] + 0.5 * s12 + W.m[8]};
}
// Selfing marriage on the same 3x3 tensor: only the diagonal gm == gp is meaningful.
__device__ __forceinline__ Tri mat_diag(const Mat3 &W) { return Tri{W.m[0], W.m[4], W.m[8]}; }
// Selfing marriage: one parent state g, T[gk][g][g] = {1,.25,0 | 0,.5,0 | 0,.25,1} over g.
__device__ __forceinline__ Tri self_kid_factor(Tri k)   // K(g,g)
{
    return Tri{k.x, (0.25 * k.x + 0.5 * k.y) + 0.25 * k.z, k.z};
}
__device__ __forceinline__ Tri self_to_kid(Tri w)       // out[gk] = sum_g T[gk][g][g] w[g]
{
    return Tri{w.x + 0.25 * w.y, 0.5 * w.y, 0.25 * w.y + w.z};
}

// exact num / den for 16-bit numerators: Markstein's correction step on the rounded reciprocal
__device__ __forceinline__ double div_exact(double num, double den, double rcp)
{
    const double q = num * rcp;
    return fma(fma(-den, q, num), rcp, q);
}

// ---- family contractions -----------------------------------------------------------------------
// A marriage with father message a[gp], mother message b[gm] and kid messages k_j contracts
//     W[gm][gp] = a[gp] * b[gm] * prod_j K_j(gm,gp),   K_j(gm,gp) = sum_gk T[gk][gm][gp] k_j[gk].
// Every K_j is symmetric in (gm,gp) and so is their elementwise product P (6 values); only the
// parents' own factors break the symmetry. Messages out of the marriage:
//     to the father  out[gp] = sum_gm b[gm] P[gm][gp]        (a left out)
//     to the mother  out[gm] = sum_gp a[gp] P[gm][gp]        (b left out)
//     to kid j       out[gk] = sum T[gk][gm][gp] a[gp] b[gm] P_{-j}[gm][gp]
//     prong          the same with all kids in P                (_FillInProngs@0x10001dd70)
// A selfing marriage has one parent state g = gm = gp: only the diagonal of P and a are used.
__device__ __forceinline__ Sym3 sym_ones() { return Sym3{1.0, 1.0, 1.0, 1.0, 1.0, 1.0}; }
__device__ __forceinline__ void sym_mul(Sym3 &P, const Sym3 &K)
{
    P.a00 *= K.a00; P.a01 *= K.a01; P.a02 *= K.a02; P.a11 *= K.a11; P.a12 *= K.a12; P.a22 *= K.a22;
}
__device__ __forceinline__ Tri fam_to_pa(const Sym3 &P, Tri b, int selfing)
{
    if (selfing) return Tri{P.a00, P.a11, P.a22};
    return Tri{b.x * P.a00 + b.y * P.a01 + b.z * P.a02, b.x * P.a01 + b.y * P.a11 + b.z * P.a12, b.x * P.a02 + b.y * P.a12 + b.z * P.a22};
}
__device__ __forceinline__ Tri fam_to_kid(const Sym3 &P, Tri a, Tri b, int selfing)
{
    if (selfing) return self_to_kid(Tri{a.x * P.a00, a.y * P.a11, a.z * P.a22});
    Mat3 W;     // W[gm][gp] = b[gm] * a[gp] * P[gm][gp]
    W.m[0] = b.x * a.x * P.a00; W.m[1] = b.x * a.y * P.a01; W.m[2] = b.x * a.z * P.a02;
    W.m[3] = b.y * a.x * P.a01; W.m[4] = b.y * a.y * P.a11; W.m[5] = b.y * a.z * P.a12;
    W.m[6] = b.z * a.x * P.a02; W.m[7] = b.z * a.y * P.a12; W.m[8] = b.z * a.z * P.a22;
    return mat_to_kid(W);
}

// ---- log of a product without one log() per factor ------------------------------------------
// value = m * 2^e * exp(slow). Positive normal factors are split into mantissa [1,2) and
// exponent with integer ops; zero, denormal, negative, inf and nan factors take the log() path
// so the result is exactly what sum-of-logs would give (-inf, nan, ...).
struct LogProd { double m; int e; double slow; };

__device__ __forceinline__ void lp_init(LogProd &a) { a.m = 1.0; a.e = 0; a.slow = 0.0; }

// out of line on purpose: log() is ~1 KB of SASS and this path is cold (zero, denormal, negative,
// inf or nan factor); inlining it at every product site blew the kernels past the instruction cache
static __device__ __noinline__ double lp_slow_log(double x) { return log(x); }

__device__ __forceinline__ void lp_mul(LogProd &a, double x)
{
    const int hi = __double2hiint(x);
    const unsigned se = (unsigned)hi >> 20;              // sign + exponent
    if (se - 1u < 0x7feu) {                              // positive, normal
        const double mant = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
        a.m *= mant;
        a.e += (int)se - 1023;
    } else {
        a.slow += lp_slow_log(x);
    }
}
// bring the mantissa back to [1,2); call at least every 512 lp_mul
__device__ __forceinline__ void lp_renorm(LogProd &a)
{
    const int hi = __double2hiint(a.m);
    const int ex = (hi >> 20) & 0x7ff;
    a.m = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(a.m));
    a.e += ex - 1023;
}
__device__ __forceinline__ void lp_merge(LogProd &a, const LogProd &b)
{
    a.m *= b.m; a.e += b.e; a.slow += b.slow;
}
__device__ __forceinline__ double lp_log(const LogProd &a)
{
    return (log(a.m) + (double)a.e * 0.693147180559945309417232121458) + a.slow;
}
// product over the 32 lanes of a warp, result valid in every lane (fixed butterfly order)
__device__ __forceinline__ LogProd lp_warp_reduce(LogProd a)
{
    lp_renorm(a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        LogProd b;
        b.m = __shfl_xor_sync(0xffffffffu, a.m, o);
        b.e = __shfl_xor_sync(0xffffffffu, a.e, o);
        b.slow = __shfl_xor_sync(0xffffffffu, a.slow, o);
        lp_merge(a, b);
    }
    return a;
}

} // namespace pf
