/*
 * pedfac_oracle.c — CPU restatement of the pedFac `pedigraph` likelihood layer (see header:
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED BY THE REFERENCE — pinned by _TPROB bytes,
 * brute-force enumeration and SURVEY.md App. E).
 *
 * Every function names the reference routine it follows as pedigraph@<VM address> <symbol>
 * (addresses of the Mach-O binary /root/reference/exec/pedigraph).
 * Arithmetic is IEEE fp64 in the reference's operation order: messages are raw products with
 * no normalisation, one log() per (node | proposal, locus), O(k^2) family messages (K_k is
 * recomputed for every out-edge), full re-contraction per proposal.
 *
 * Differences that do not change any value: flat arrays instead of pointer-of-pointer rows;
 * per-edge messages live in a per-sweep arena instead of persisting on the mnode (the reference
 * recomputes all of them for a dirty component anyway); the swap-with-last worklist
 * bookkeeping is replaced by repeated scans (a tree message depends only on its inputs, so the
 * schedule cannot change a bit of the result).
 */
#include "pedfac_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* pedigraph@0x100035b10 _TPROB: P(kid = gk | ma = gm, pa = gp), index gk*9 + gm*3 + gp.
 * tests/golden/tprob.json holds the 216 bytes extracted from the binary. */
static const double TPROB[27] = {
    1.0, 0.5, 0.0, 0.5, 0.25, 0.0, 0.0, 0.0, 0.0,
    0.0, 0.5, 1.0, 0.5, 0.5,  0.5, 1.0, 0.5, 0.0,
    0.0, 0.0, 0.0, 0.0, 0.25, 0.5, 0.0, 0.5, 1.0,
};

static __thread char g_err[512];
#define FAIL(code, ...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return (code); } while (0)

const char *pfo_last_error(void) { return g_err; }
const double *pfo_tprob(void) { return TPROB; }

struct pfo_engine {
    int S, s0, s1, Sl;      /* data-set loci, shard, local count */
    int capI, capM;
    double *eps;            /* [Sl] */
    double *prior;          /* [Sl][3]  ped->prior */
    double **emis;          /* [capI] -> [Sl][3] or NULL (= all ones)  indiv->lik */
    double **stub;          /* [capI] -> [Sl][3]  indiv->stub */
    double **prong;         /* [capM] -> [Sl][3]  mnode->prong */
    int *loc;               /* [capI] slot -> index in the current view, else -1 */
    char *arena; size_t arena_cap, arena_used;
};

static void *arena_get(pfo_engine *e, size_t bytes)
{
    bytes = (bytes + 63) & ~(size_t)63;
    if (e->arena_used + bytes > e->arena_cap) return NULL;
    void *p = e->arena + e->arena_used;
    e->arena_used += bytes;
    return p;
}

int pfo_create(pfo_engine **out, const pf_config *cfg)
{
    if (!out || !cfg) FAIL(PF_EINVAL, "null argument");
    if (cfg->n_loci <= 0 || cfg->locus_begin < 0 || cfg->locus_end > cfg->n_loci ||
        cfg->locus_begin >= cfg->locus_end || cfg->cap_indiv <= 0 || cfg->cap_marr < 0)
        FAIL(PF_EINVAL, "bad config");
    pfo_engine *e = calloc(1, sizeof *e);
    if (!e) FAIL(PF_ENOMEM, "oom");
    e->S = cfg->n_loci; e->s0 = cfg->locus_begin; e->s1 = cfg->locus_end; e->Sl = e->s1 - e->s0;
    e->capI = cfg->cap_indiv; e->capM = cfg->cap_marr;
    e->eps = calloc(e->Sl, sizeof(double));
    e->prior = calloc((size_t)e->Sl * 3, sizeof(double));
    e->emis = calloc(e->capI, sizeof(double *));
    e->stub = calloc(e->capI, sizeof(double *));
    e->prong = calloc(e->capM > 0 ? e->capM : 1, sizeof(double *));
    e->loc = malloc(sizeof(int) * e->capI);
    if (!e->eps || !e->prior || !e->emis || !e->stub || !e->prong || !e->loc) FAIL(PF_ENOMEM, "oom");
    for (int i = 0; i < e->capI; i++) e->loc[i] = -1;
    *out = e;
    return PF_OK;
}

int pfo_destroy(pfo_engine *e)
{
    if (!e) return PF_OK;
    for (int i = 0; i < e->capI; i++) { free(e->emis[i]); free(e->stub[i]); }
    for (int m = 0; m < e->capM; m++) free(e->prong[m]);
    free(e->eps); free(e->prior); free(e->emis); free(e->stub); free(e->prong); free(e->loc);
    free(e->arena); free(e);
    return PF_OK;
}

/* pedigraph@0x1000099d2-0x100009ad4 (_ReadPedfile, key aFreq): q = 1 - atof(tok);
 * prior = [q*q, (2*(1-q))*q, (1-q)*(1-q)]. Note (1-q) is recomputed, it is not `a`. */
int pfo_set_locus_params(pfo_engine *e, const double *eps, const double *afreq)
{
    if (!e || !eps || !afreq) FAIL(PF_EINVAL, "null argument");
    for (int s = 0; s < e->Sl; s++) {
        double q = 1.0 - afreq[e->s0 + s];
        e->eps[s] = eps[e->s0 + s];
        e->prior[s * 3 + 0] = q * q;
        e->prior[s * 3 + 1] = (2.0 * (1.0 - q)) * q;
        e->prior[s * 3 + 2] = (1.0 - q) * (1.0 - q);
    }
    return PF_OK;
}

static double *emis_row(pfo_engine *e, int slot)
{
    if (slot < 0 || slot >= e->capI) return NULL;
    if (!e->emis[slot]) e->emis[slot] = malloc(sizeof(double) * 3 * e->Sl);
    return e->emis[slot];
}

/* pedigraph@0x10000b49e-0x10000b69b (_ReadPedfile, integer token): 3 -> [1,1,1];
 * c in {0,1,2} -> all three eps/2.0, then e[c] = 1.0 - eps. */
int pfo_set_genotypes_hard(pfo_engine *e, int32_t slot, const uint8_t *codes)
{
    double *r = e ? emis_row(e, slot) : NULL;
    if (!r || !codes) FAIL(PF_EINVAL, "bad slot %d", slot);
    for (int s = 0; s < e->Sl; s++) {
        int c = codes[e->s0 + s];
        if (c > 3) FAIL(PF_EINVAL, "Acceptable genotype values are 0, 1, 2, or 3 (slot %d locus %d: %d)", slot, e->s0 + s, c);
        if (c == 3) { r[s * 3] = r[s * 3 + 1] = r[s * 3 + 2] = 1.0; continue; }
        for (int g = 0; g < 3; g++) r[s * 3 + g] = e->eps[s] / 2.0;
        r[s * 3 + c] = 1.0 - e->eps[s];
    }
    return PF_OK;
}

/* pedigraph@0x10000b363-0x10000b45e (_ReadPedfile, comma token): the three numbers verbatim. */
int pfo_set_genotypes_soft_f64(pfo_engine *e, int32_t slot, const double *lik)
{
    double *r = e ? emis_row(e, slot) : NULL;
    if (!r || !lik) FAIL(PF_EINVAL, "bad slot %d", slot);
    memcpy(r, lik + (size_t)e->s0 * 3, sizeof(double) * 3 * e->Sl);
    return PF_OK;
}

int pfo_set_genotypes_soft(pfo_engine *e, int32_t slot, const float *lik)
{
    double *r = e ? emis_row(e, slot) : NULL;
    if (!r || !lik) FAIL(PF_EINVAL, "bad slot %d", slot);
    for (int k = 0; k < 3 * e->Sl; k++) r[k] = (double)lik[(size_t)e->s0 * 3 + k];
    return PF_OK;
}

int pfo_set_genotypes_soft_dec16(pfo_engine *e, int32_t slot, const uint16_t *num, uint32_t denom)
{
    double *r = e ? emis_row(e, slot) : NULL;
    if (!r || !num || denom == 0) FAIL(PF_EINVAL, "bad slot %d", slot);
    for (int k = 0; k < 3 * e->Sl; k++) r[k] = (double)num[(size_t)e->s0 * 3 + k] / (double)denom;
    return PF_OK;
}

/* pedigraph@0x10000b1d0-0x10000b262 / _AttachUnobsIndiv@0x100012cfa: emission all ones. */
int pfo_set_genotypes_unobserved(pfo_engine *e, int32_t slot)
{
    if (!e || slot < 0 || slot >= e->capI) FAIL(PF_EINVAL, "bad slot %d", slot);
    free(e->emis[slot]); e->emis[slot] = NULL;
    return PF_OK;
}

/* ---------------------------------------------------------------------------------------- */
/* Sweep = pedigraph@0x10001ba10 _Propagating_msg restricted to the view.                    */

typedef struct {
    int nM;        /* indiv->nM */
    int *mar;      /* local marriage index  (indiv->M[k]) */
    int *edge;     /* edge index in that marriage: 0 pa, 1 ma, >= 2 kid (indiv->edgeIdxInM[k]) */
} iadj;

typedef struct {
    int ne;            /* 2 + kids (mnode->nEdges); edge 1 unused when selfing */
    int selfing;
    double *in;        /* [ne][Sl][3]  mnode->in  (individual -> marriage) */
    double *out;       /* [ne][Sl][3]  mnode->out (marriage -> individual) */
    char *passIn, *passOut;
} madj;

int pfo_sweep(pfo_engine *e, int32_t chain, const pf_graph_view *v, double *ll_cc)
{
    (void)chain;
    if (!e || !v || !ll_cc) FAIL(PF_EINVAL, "null argument");
    const int Sl = e->Sl, nI = v->n_indiv, nM = v->n_marr;
    const size_t T3 = (size_t)Sl * 3;
    int rc = PF_OK;

    /* size the arena */
    size_t nedges = 0;
    for (int j = 0; j < nM; j++) nedges += 2 + (v->marr_kid_begin[j + 1] - v->marr_kid_begin[j]);
    size_t need = sizeof(double) * T3 * (6 * (size_t)nI + 2 * nedges) + sizeof(double) * nI
                + (sizeof(iadj) + 64) * (size_t)(nI + 1) + (sizeof(madj) + 192) * (size_t)(nM + 1)
                + (sizeof(int) * 2 + 2) * (nedges + 16) + 4096;
    if (need > e->arena_cap) {
        free(e->arena);
        e->arena = malloc(need + need / 4);
        if (!e->arena) { e->arena_cap = 0; FAIL(PF_ENOMEM, "oom (%zu bytes)", need); }
        e->arena_cap = need + need / 4;
    }
    e->arena_used = 0;

    iadj *ia = arena_get(e, sizeof(iadj) * (nI + 1));
    madj *ma = arena_get(e, sizeof(madj) * (nM + 1));
    for (int i = 0; i < nI; i++) {
        int slot = v->indiv[i];
        if (slot < 0 || slot >= e->capI || e->loc[slot] != -1) { rc = PF_EINVAL; snprintf(g_err, sizeof g_err, "bad or duplicate individual slot %d in view", slot); goto done_reset; }
        if (v->indiv_cc[i] < 0 || v->indiv_cc[i] >= v->n_cc) { rc = PF_EINVAL; snprintf(g_err, sizeof g_err, "component label out of range"); goto done_reset; }
        e->loc[slot] = i;
        ia[i].nM = 0;
    }
    /* count attachments, then fill (order = order of appearance in the marriage list) */
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            for (int i = 0; i < nI; i++) {
                ia[i].mar = arena_get(e, sizeof(int) * (ia[i].nM + 1));
                ia[i].edge = arena_get(e, sizeof(int) * (ia[i].nM + 1));
                ia[i].nM = 0;
            }
        }
        for (int j = 0; j < nM; j++) {
            int nk = v->marr_kid_begin[j + 1] - v->marr_kid_begin[j];
            int ne = 2 + nk;
            for (int ed = 0; ed < ne; ed++) {
                if (ed == 1 && v->marr_selfing[j]) continue;
                int slot = ed == 0 ? v->marr_pa[j] : ed == 1 ? v->marr_ma[j] : v->marr_kids[v->marr_kid_begin[j] + ed - 2];
                if (slot < 0 || slot >= e->capI || e->loc[slot] < 0) { rc = PF_EINVAL; snprintf(g_err, sizeof g_err, "view not closed: marriage %d references unlisted individual %d", v->marr[j], slot); goto done_reset; }
                int i = e->loc[slot];
                if (pass == 1) { ia[i].mar[ia[i].nM] = j; ia[i].edge[ia[i].nM] = ed; }
                ia[i].nM++;
            }
        }
    }
    for (int j = 0; j < nM; j++) {
        int ne = 2 + (v->marr_kid_begin[j + 1] - v->marr_kid_begin[j]);
        if (v->marr[j] < 0 || v->marr[j] >= e->capM) { rc = PF_EINVAL; snprintf(g_err, sizeof g_err, "bad marriage slot"); goto done_reset; }
        ma[j].ne = ne; ma[j].selfing = v->marr_selfing[j];
        ma[j].in = arena_get(e, sizeof(double) * T3 * ne);
        ma[j].out = arena_get(e, sizeof(double) * T3 * ne);
        ma[j].passIn = arena_get(e, ne); ma[j].passOut = arena_get(e, ne);   /* _ResetMsg@0x10001c210 */
        memset(ma[j].passIn, 0, ne); memset(ma[j].passOut, 0, ne);
    }
    double *pOut = arena_get(e, sizeof(double) * T3 * nI);   /* pnode->out */
    double *gOut = arena_get(e, sizeof(double) * T3 * nI);   /* gnode->out */
    double *pIn = arena_get(e, sizeof(double) * T3 * nI);    /* pnode->in  */
    double *gIn = arena_get(e, sizeof(double) * T3 * nI);    /* gnode->in  */
    double *marg = arena_get(e, sizeof(double) * T3 * nI);   /* ped->f150->marg */
    double *asPar = arena_get(e, sizeof(double) * T3 * nI);  /* indiv->asParent */
    double *LLi = arena_get(e, sizeof(double) * nI);
    if (!asPar || !LLi) { rc = PF_ENOMEM; snprintf(g_err, sizeof g_err, "arena too small"); goto done_reset; }

    /* _PassingOFtoIV@0x10001c2d0: p-node sends the HWE prior for founders else 1; g-node the emission */
    for (int i = 0; i < nI; i++) {
        const double *em = e->emis[v->indiv[i]];
        for (int s = 0; s < Sl; s++) {
            for (int g = 0; g < 3; g++) pOut[i * T3 + s * 3 + g] = v->indiv_founder[i] ? e->prior[s * 3 + g] : 1.0;
            for (int g = 0; g < 3; g++) gOut[i * T3 + s * 3 + g] = em ? em[s * 3 + g] : 1.0;
        }
    }

    /* alternate _PassingIVtoIF@0x10001c490 / _PassingIFtoIV@0x10001c8f0 until no marriage waits */
    int *iremain = arena_get(e, sizeof(int) * (nI + 1));
    for (int i = 0; i < nI; i++) iremain[i] = ia[i].nM;
    int done = 0, rounds = 0;
    while (!done) {
        /* individual -> marriage k, once every other attached marriage has sent to the individual */
        for (int i = nI - 1; i >= 0; i--) {
            if (!iremain[i]) continue;
            for (int k = ia[i].nM - 1; k >= 0; k--) {
                madj *mk = &ma[ia[i].mar[k]];
                int ek = ia[i].edge[k];
                if (mk->passIn[ek]) continue;
                int ready = 1;
                for (int j = 0; j < ia[i].nM; j++)
                    if (j != k) ready = ready && ma[ia[i].mar[j]].passOut[ia[i].edge[j]];
                if (!ready) continue;
                for (int s = 0; s < Sl; s++)
                    for (int g = 0; g < 3; g++) {
                        double prod = 1.0;
                        for (int j = 0; j < ia[i].nM; j++)
                            if (j != k) prod = prod * ma[ia[i].mar[j]].out[ia[i].edge[j] * T3 + s * 3 + g];
                        mk->in[ek * T3 + s * 3 + g] = (prod * pOut[i * T3 + s * 3 + g]) * gOut[i * T3 + s * 3 + g];
                    }
                mk->passIn[ek] = 1;
                iremain[i]--;
            }
        }
        /* marriage -> individual on edge ed, once all other edges have arrived */
        done = 1;
        for (int j = nM - 1; j >= 0; j--) {
            madj *m = &ma[j];
            const int ne = m->ne, self = m->selfing;
            for (int ed = ne - 1; ed >= 0; ed--) {
                if (ed == 1 && self) continue;
                if (m->passOut[ed]) continue;
                int ready = 1;
                for (int x = 0; x < ne; x++) {
                    if (x == 1 && self) continue;
                    if (x != ed) ready = ready && m->passIn[x];
                }
                if (!ready) { done = 0; continue; }
                for (int s = 0; s < Sl; s++) {
                    const double *in = m->in;
                    for (int go = 0; go < 3; go++) {
                        double acc;
                        if (ed == 0 && !self) {             /* to father: sum over mother's state */
                            acc = 0.0;
                            for (int gm = 0; gm < 3; gm++) {
                                double t = in[1 * T3 + s * 3 + gm];
                                for (int k = 2; k < ne; k++) {
                                    double K = 0.0;
                                    for (int gk = 0; gk < 3; gk++) K = K + TPROB[gk * 9 + gm * 3 + go] * in[k * T3 + s * 3 + gk];
                                    t = t * K;
                                }
                                acc = acc + t;
                            }
                        } else if (ed == 0) {               /* selfing parent */
                            acc = 1.0;
                            for (int k = 2; k < ne; k++) {
                                double K = 0.0;
                                for (int gk = 0; gk < 3; gk++) K = K + TPROB[gk * 9 + go * 3 + go] * in[k * T3 + s * 3 + gk];
                                acc = acc * K;
                            }
                        } else if (ed == 1) {               /* to mother: sum over father's state */
                            acc = 0.0;
                            for (int gp = 0; gp < 3; gp++) {
                                double t = in[0 * T3 + s * 3 + gp];
                                for (int k = 2; k < ne; k++) {
                                    double K = 0.0;
                                    for (int gk = 0; gk < 3; gk++) K = K + TPROB[gk * 9 + go * 3 + gp] * in[k * T3 + s * 3 + gk];
                                    t = t * K;
                                }
                                acc = acc + t;
                            }
                        } else if (!self) {                 /* to kid `ed` */
                            acc = 0.0;
                            for (int gp = 0; gp < 3; gp++)
                                for (int gm = 0; gm < 3; gm++) {
                                    double t = (TPROB[go * 9 + gm * 3 + gp] * in[0 * T3 + s * 3 + gp]) * in[1 * T3 + s * 3 + gm];
                                    for (int k = 2; k < ne; k++) {
                                        if (k == ed) continue;
                                        double K = 0.0;
                                        for (int gk = 0; gk < 3; gk++) K = K + TPROB[gk * 9 + gm * 3 + gp] * in[k * T3 + s * 3 + gk];
                                        t = t * K;
                                    }
                                    acc = acc + t;
                                }
                        } else {                            /* to kid of a selfing parent */
                            acc = 0.0;
                            for (int g = 0; g < 3; g++) {
                                double t = TPROB[go * 9 + g * 3 + g] * in[0 * T3 + s * 3 + g];
                                for (int k = 2; k < ne; k++) {
                                    if (k == ed) continue;
                                    double K = 0.0;
                                    for (int gk = 0; gk < 3; gk++) K = K + TPROB[gk * 9 + g * 3 + g] * in[k * T3 + s * 3 + gk];
                                    t = t * K;
                                }
                                acc = acc + t;
                            }
                        }
                        m->out[ed * T3 + s * 3 + go] = acc;
                    }
                }
                m->passOut[ed] = 1;
            }
        }
        /* pedigraph@0x10001ba99-0x10001baf2: more rounds than individuals => not a forest */
        if (rounds++ > nI) { rc = PF_ECYCLE; snprintf(g_err, sizeof g_err, "propagation step exceeds the acyclic upper limit"); goto done_reset; }
    }

    /* _PassingIVtoOF@0x10001d840 */
    for (int i = 0; i < nI; i++)
        for (int s = 0; s < Sl; s++)
            for (int g = 0; g < 3; g++) {
                double B = 1.0;
                for (int j = 0; j < ia[i].nM; j++) B = B * ma[ia[i].mar[j]].out[ia[i].edge[j] * T3 + s * 3 + g];
                gIn[i * T3 + s * 3 + g] = B * pOut[i * T3 + s * 3 + g];
                pIn[i * T3 + s * 3 + g] = B * gOut[i * T3 + s * 3 + g];
            }

    /* _UpdateProb@0x10001da20: Z, LL[i] = sum_s log Z, marginals; LLcc[cc(i)] = LL[i] (last writer) */
    for (int c = 0; c < v->n_cc; c++) ll_cc[c] = 0.0;
    for (int i = 0; i < nI; i++) {
        LLi[i] = 0.0;
        for (int s = 0; s < Sl; s++) {
            const double *gi = gIn + i * T3 + s * 3, *go = gOut + i * T3 + s * 3;
            double Z = (gi[0] * go[0] + gi[1] * go[1]) + gi[2] * go[2];
            LLi[i] = LLi[i] + log(Z);
            for (int g = 0; g < 3; g++) marg[i * T3 + s * 3 + g] = (gi[g] * go[g]) / Z;
        }
        ll_cc[v->indiv_cc[i]] = LLi[i];
    }

    /* _FillInProngs@0x10001dd70: the message a NEW kid of the marriage would receive */
    for (int j = 0; j < nM; j++) {
        madj *m = &ma[j];
        int slot = v->marr[j];
        if (!e->prong[slot]) e->prong[slot] = malloc(sizeof(double) * T3);
        double *pr = e->prong[slot];
        const double *in = m->in;
        for (int s = 0; s < Sl; s++)
            for (int go = 0; go < 3; go++) {
                double acc = 0.0;
                if (!m->selfing) {
                    for (int gp = 0; gp < 3; gp++)
                        for (int gm = 0; gm < 3; gm++) {
                            double t = (TPROB[go * 9 + gm * 3 + gp] * in[0 * T3 + s * 3 + gp]) * in[1 * T3 + s * 3 + gm];
                            for (int k = 2; k < m->ne; k++) {
                                double K = 0.0;
                                for (int gk = 0; gk < 3; gk++) K = K + TPROB[gk * 9 + gm * 3 + gp] * in[k * T3 + s * 3 + gk];
                                t = t * K;
                            }
                            acc = acc + t;
                        }
                } else {
                    for (int g = 0; g < 3; g++) {
                        double t = TPROB[go * 9 + g * 3 + g] * in[0 * T3 + s * 3 + g];
                        for (int k = 2; k < m->ne; k++) {
                            double K = 0.0;
                            for (int gk = 0; gk < 3; gk++) K = K + TPROB[gk * 9 + g * 3 + g] * in[k * T3 + s * 3 + gk];
                            t = t * K;
                        }
                        acc = acc + t;
                    }
                }
                pr[s * 3 + go] = acc;
            }
    }

    /* _FillInStubs@0x10001e1a0 (+ _fillInMsgOutMnodeAsParent@0x10001e440, whose output nothing reads) */
    for (int i = 0; i < nI; i++) {
        int slot = v->indiv[i];
        if (!e->stub[slot]) e->stub[slot] = malloc(sizeof(double) * T3);
        double *st = e->stub[slot];
        for (int s = 0; s < Sl; s++)
            for (int gk = 0; gk < 3; gk++) {
                double a = 0.0;
                for (int gp = 0; gp < 3; gp++)
                    for (int gm = 0; gm < 3; gm++)
                        a = a + (TPROB[gk * 9 + gm * 3 + gp] * marg[i * T3 + s * 3 + gp]) * e->prior[s * 3 + gm];
                asPar[i * T3 + s * 3 + gk] = a;
            }
        for (int s = 0; s < Sl; s++)
            for (int g = 0; g < 3; g++) {
                double B = 1.0;
                for (int j = 0; j < ia[i].nM; j++) B = B * ma[ia[i].mar[j]].out[ia[i].edge[j] * T3 + s * 3 + g];
                st[s * 3 + g] = (B * pOut[i * T3 + s * 3 + g]) * gOut[i * T3 + s * 3 + g];
            }
    }

done_reset:
    for (int i = 0; i < nI; i++) {
        int slot = v->indiv[i];
        if (slot >= 0 && slot < e->capI) e->loc[slot] = -1;
    }
    return rc;
}

/* ---------------------------------------------------------------------------------------- */
/* Proposal data terms.                                                                      */

static const double *stub_or_prior(const pfo_engine *e, int slot)
{
    return slot == -1 ? e->prior : e->stub[slot];
}

/* One locus of _CalculateLikeByStubs@0x10001f342-0x10001f4fc: the TPROB entry is multiplied by
 * the father's factor then the mother's, summed gp-outer / gm-inner, then weighted by the kid's stub. */
static inline double trio_locus(const double *xp, const double *xm, const double *kid)
{
    double tot = 0.0;
    for (int gk = 0; gk < 3; gk++) {
        double inner = 0.0;
        for (int gp = 0; gp < 3; gp++)
            for (int gm = 0; gm < 3; gm++) {
                double t = TPROB[gk * 9 + gm * 3 + gp];
                t = t * xp[gp];
                t = t * xm[gm];
                inner = inner + t;
            }
        tot = tot + inner * kid[gk];
    }
    return tot;
}

int pfo_eval(pfo_engine *e, int32_t chain, int32_t kid, int32_t n, const pf_proposal *props, double *lsum)
{
    (void)chain;
    if (!e || !props || !lsum) FAIL(PF_EINVAL, "null argument");
    if (kid < 0 || kid >= e->capI || !e->stub[kid]) FAIL(PF_EINVAL, "kid %d has no stub (sweep it first)", kid);
    const int Sl = e->Sl;
    const double *ks = e->stub[kid];
    for (int p = 0; p < n; p++) {
        const pf_proposal *q = &props[p];
        double acc = 0.0;
        if (q->kind == PF_PROP_TRIO) {
            if ((q->pa != -1 && (q->pa < 0 || q->pa >= e->capI || !e->stub[q->pa])) ||
                (q->ma != -1 && (q->ma < 0 || q->ma >= e->capI || !e->stub[q->ma])))
                FAIL(PF_EINVAL, "proposal %d: parent without stub", p);
            const double *xp = stub_or_prior(e, q->pa), *xm = stub_or_prior(e, q->ma);
            for (int s = 0; s < Sl; s++) acc = acc + log(trio_locus(xp + s * 3, xm + s * 3, ks + s * 3));
        } else if (q->kind == PF_PROP_SELFING) {   /* _CalculateLikeByStubsSelfing@0x10001f7ab-0x10001f921 */
            if (q->pa != -1 && (q->pa < 0 || q->pa >= e->capI || !e->stub[q->pa])) FAIL(PF_EINVAL, "proposal %d: parent without stub", p);
            const double *xp = stub_or_prior(e, q->pa);
            for (int s = 0; s < Sl; s++) {
                double tot = 0.0;
                for (int gk = 0; gk < 3; gk++) {
                    double inner = 0.0;
                    for (int g = 0; g < 3; g++) inner = inner + TPROB[gk * 9 + g * 3 + g] * xp[s * 3 + g];
                    tot = tot + inner * ks[s * 3 + gk];
                }
                acc = acc + log(tot);
            }
        } else if (q->kind == PF_PROP_PRONG) {     /* _CalculateLikeByProngs@0x10001fb44-0x10001fc20 */
            if (q->marr < 0 || q->marr >= e->capM || !e->prong[q->marr]) FAIL(PF_EINVAL, "proposal %d: marriage without prong", p);
            const double *pr = e->prong[q->marr];
            for (int s = 0; s < Sl; s++) {
                double tot = 0.0;
                for (int g = 0; g < 3; g++) tot = tot + pr[s * 3 + g] * ks[s * 3 + g];
                acc = acc + log(tot);
            }
        } else FAIL(PF_EINVAL, "proposal %d: unknown kind %d", p, q->kind);
        lsum[p] = acc;
    }
    return PF_OK;
}

/* _RefineParentOffPair@0x100014d20: score every (kid, candidate != kid) with
 * _CalculateLLByStubs@0x10001eef0 (candidate as the single given parent, the other a new founder),
 * sort descending (_compareLLonGC), keep top_k. The age-window test is the caller's (it chooses
 * the candidate set); the -LLcc terms of @0x10001f183-0x10001f241 arrive as cand_bias.
 * Ties: earlier candidate first. */
int pfo_prefilter(pfo_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids,
                  const uint8_t *as_father, int32_t n_cands, const int32_t *cands,
                  const double *cand_bias, int32_t top_k, int32_t *out_slot, double *out_score)
{
    if (!e || !kids || !as_father || (n_cands > 0 && !cands) || !out_slot || !out_score || top_k <= 0)
        FAIL(PF_EINVAL, "bad argument");
    for (int k = 0; k < n_kids; k++) {
        for (int t = 0; t < top_k; t++) { out_slot[k * top_k + t] = -1; out_score[k * top_k + t] = -INFINITY; }
        for (int c = 0; c < n_cands; c++) {
            pf_proposal pr = { PF_PROP_TRIO, -1, -1, -1 };
            int cs = cands[c];
            if (cs == kids[k]) continue;
            if (as_father[k]) pr.pa = cs; else pr.ma = cs;
            double v;
            int rc = pfo_eval(e, chain, kids[k], 1, &pr, &v);
            if (rc) return rc;
            if (cand_bias) v = v + cand_bias[c];
            /* insertion into the descending list; strict > keeps earlier candidates first on ties */
            int pos = top_k;
            while (pos > 0 && (out_slot[k * top_k + pos - 1] == -1 || v > out_score[k * top_k + pos - 1])) pos--;
            if (pos < top_k) {
                for (int t = top_k - 1; t > pos; t--) { out_slot[k * top_k + t] = out_slot[k * top_k + t - 1]; out_score[k * top_k + t] = out_score[k * top_k + t - 1]; }
                out_slot[k * top_k + pos] = cs; out_score[k * top_k + pos] = v;
            }
        }
    }
    return PF_OK;
}

int pfo_get_stub(pfo_engine *e, int32_t chain, int32_t slot, double *out)
{
    (void)chain;
    if (!e || !out || slot < 0 || slot >= e->capI || !e->stub[slot]) FAIL(PF_EINVAL, "no stub for slot %d", slot);
    memcpy(out, e->stub[slot], sizeof(double) * 3 * e->Sl);
    return PF_OK;
}

int pfo_get_prong(pfo_engine *e, int32_t chain, int32_t marr, double *out)
{
    (void)chain;
    if (!e || !out || marr < 0 || marr >= e->capM || !e->prong[marr]) FAIL(PF_EINVAL, "no prong for marriage %d", marr);
    memcpy(out, e->prong[marr], sizeof(double) * 3 * e->Sl);
    return PF_OK;
}
