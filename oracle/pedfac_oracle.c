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
 * the swap-with-last worklist
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
    double **m_in, **m_out; /* [capM] -> [ne][Sl][3]  mnode->in / mnode->out of the last sweep of that marriage */
    int *m_ne;              /* [capM] edge capacity of the stored message arrays */
    int *loc;               /* [capI] slot -> index in the current view, else -1 */
    char *arena; size_t arena_cap, arena_used;
    double *deferred; int n_deferred;   /* swap values held for pfo_deferred_results */
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
    e->m_in = calloc(e->capM > 0 ? e->capM : 1, sizeof(double *));
    e->m_out = calloc(e->capM > 0 ? e->capM : 1, sizeof(double *));
    e->m_ne = calloc(e->capM > 0 ? e->capM : 1, sizeof(int));
    e->loc = malloc(sizeof(int) * e->capI);
    if (!e->eps || !e->prior || !e->emis || !e->stub || !e->prong || !e->loc || !e->m_in || !e->m_out || !e->m_ne) FAIL(PF_ENOMEM, "oom");
    for (int i = 0; i < e->capI; i++) e->loc[i] = -1;
    *out = e;
    return PF_OK;
}

int pfo_destroy(pfo_engine *e)
{
    if (!e) return PF_OK;
    for (int i = 0; i < e->capI; i++) { free(e->emis[i]); free(e->stub[i]); }
    for (int m = 0; m < e->capM; m++) { free(e->prong[m]); free(e->m_in[m]); free(e->m_out[m]); }
    free(e->m_in); free(e->m_out); free(e->m_ne);
    free(e->eps); free(e->prior); free(e->emis); free(e->stub); free(e->prong); free(e->loc);
    free(e->arena); free(e->deferred); free(e);
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
    size_t need = sizeof(double) * T3 * (6 * (size_t)nI) + sizeof(double) * nI
                + (sizeof(iadj) + 256) * (size_t)(nI + 1) + (sizeof(madj) + 256) * (size_t)(nM + 1)
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
                if (!ia[i].mar || !ia[i].edge) { rc = PF_ENOMEM; snprintf(g_err, sizeof g_err, "arena too small"); goto done_reset; }
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
        {   /* mnode->in / mnode->out persist between sweeps (the swap proposals read them) */
            const int slot = v->marr[j];
            if (e->m_ne[slot] < ne || !e->m_in[slot]) {
                free(e->m_in[slot]); free(e->m_out[slot]);
                e->m_in[slot] = malloc(sizeof(double) * T3 * (ne + 2));
                e->m_out[slot] = malloc(sizeof(double) * T3 * (ne + 2));
                if (!e->m_in[slot] || !e->m_out[slot]) { rc = PF_ENOMEM; snprintf(g_err, sizeof g_err, "oom"); goto done_reset; }
                e->m_ne[slot] = ne + 2;
            }
            ma[j].in = e->m_in[slot]; ma[j].out = e->m_out[slot];
        }
        ma[j].passIn = arena_get(e, ne); ma[j].passOut = arena_get(e, ne);   /* _ResetMsg@0x10001c210 */
        if (!ma[j].passIn || !ma[j].passOut) { rc = PF_ENOMEM; snprintf(g_err, sizeof g_err, "arena too small"); goto done_reset; }
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
int pfo_sweep_eval(pfo_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc,
                   int32_t kid, int32_t n, const pf_proposal *props, double *lsum)
{
    int rc = pfo_sweep(e, chain, view, ll_cc);
    if (rc != PF_OK || n == 0) return rc;
    return pfo_eval(e, chain, kid, n, props, lsum);
}

/* ---------------------------------------------------------------------------------------- */
/* Swap proposals: _ProposeSwapping@0x1000223d0 and its helpers, on the per-edge messages the
 * last sweep left on the marriages (mnode->in / mnode->out).                                 */

typedef struct {
    pfo_engine *e;
    const pf_graph_view *v;
    int *mloc;      /* marriage slot -> index in the view, else -1 */
    iadj *ia;       /* per view individual: attached marriages (view index) and edge */
    size_t T3;
} swap_ctx;

static int view_edge_slot(const pf_graph_view *v, int j, int ed)
{
    return ed == 0 ? v->marr_pa[j] : ed == 1 ? v->marr_ma[j] : v->marr_kids[v->marr_kid_begin[j] + ed - 2];
}

/* _initSwapMsg@0x100021cd0: the message a kid receives from a family whose father message comes
 * from marriage mPa's edge 0 (or the stub of iPa when mPa == -1), mother message from mMa's edge 1
 * (or the stub of iMa), siblings = the kids of mKids except exclKid (none when mKids == -1). */
static void init_swap_msg(const swap_ctx *c, double *holder, int mPa, int mMa, int mKids, int iPa, int iMa, int exclKid)
{
    const pfo_engine *e = c->e;
    const size_t T3 = c->T3;
    const double *paMsg = e->stub[iPa], *maMsg = e->stub[iMa];
    if (mPa != -1) paMsg = e->m_in[mPa];                    /* @0x100021d47-0x100021d64 */
    if (mMa != -1) maMsg = e->m_in[mMa] + T3;               /* @0x100021d68-0x100021d8a */
    const int jk = mKids == -1 ? -1 : c->mloc[mKids];
    const int ne = jk < 0 ? 2 : 2 + (c->v->marr_kid_begin[jk + 1] - c->v->marr_kid_begin[jk]);
    for (int s = 0; s < e->Sl; s++)
        for (int gk = 0; gk < 3; gk++) {
            double acc = 0.0;
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) {
                    double t = (TPROB[gk * 9 + b * 3 + a] * paMsg[s * 3 + a]) * maMsg[s * 3 + b];
                    for (int k = 2; k < ne; k++) {          /* @0x100021f51-0x100022059 */
                        if (view_edge_slot(c->v, jk, k) == exclKid) continue;
                        double K = 0.0;
                        for (int g = 0; g < 3; g++) K = K + TPROB[g * 9 + b * 3 + a] * e->m_in[mKids][k * T3 + s * 3 + g];
                        t = t * K;
                    }
                    acc = acc + t;
                }
            holder[s * 3 + gk] = acc;
        }
}

/* _directpassIV2IF@0x100020810: push the holder through individual `slot`: multiply by what its
 * other marriages (all but mA, mB) send it, by its prior message and by its emission. */
static void directpass_iv2if(const swap_ctx *c, double *holder, int slot, int mA, int mB)
{
    const pfo_engine *e = c->e;
    const size_t T3 = c->T3;
    const int i = e->loc[slot];
    const double *em = e->emis[slot];
    const int founder = c->v->indiv_founder[i];
    for (int s = 0; s < e->Sl; s++)
        for (int g = 0; g < 3; g++) {
            double t = holder[s * 3 + g];
            for (int k = 0; k < c->ia[i].nM; k++) {
                const int m = c->v->marr[c->ia[i].mar[k]];
                if (m == mA || m == mB) continue;
                t = t * e->m_out[m][c->ia[i].edge[k] * T3 + s * 3 + g];
            }
            holder[s * 3 + g] = (t * (founder ? e->prior[s * 3 + g] : 1.0)) * (em ? em[s * 3 + g] : 1.0);
        }
}

/* _directpassIF2IV@0x100020a70: push the holder through marriage M from member src to member dst:
 * the holder replaces M's incoming message on src's edge (the reference saves and restores that
 * row through ped->f160), then the ordinary marriage -> individual formula for dst's edge. */
static int directpass_if2iv(const swap_ctx *c, double *holder, double *tmp, int M, int src, int dst)
{
    const pfo_engine *e = c->e;
    const size_t T3 = c->T3;
    const int j = c->mloc[M];
    const pf_graph_view *v = c->v;
    const int ne = 2 + (v->marr_kid_begin[j + 1] - v->marr_kid_begin[j]);
    const int self = v->marr_selfing[j];
    int dst_ed = 0;                                             /* @0x100020afb-0x100020bab */
    if (v->marr_ma[j] == dst) dst_ed = 1;
    else for (int k = 2; k < ne; k++) if (view_edge_slot(v, j, k) == dst) { dst_ed = k; break; }
    int src_ed = -1;                                            /* @0x100020bb0-0x100020f13 */
    if (v->marr_pa[j] == src) src_ed = 0;
    else if (v->marr_ma[j] == src) src_ed = 1;
    else for (int k = 2; k < ne; k++) if (view_edge_slot(v, j, k) == src) { src_ed = k; break; }
    if (src_ed < 0) return -1;
    memcpy(tmp, holder, sizeof(double) * T3);
    const double **in = malloc(sizeof(double *) * (size_t)ne);
    if (!in) return -1;
    for (int k = 0; k < ne; k++) in[k] = k == src_ed ? tmp : e->m_in[M] + k * T3;
    for (int s = 0; s < e->Sl; s++)
        for (int go = 0; go < 3; go++) {
            double acc;
            if (dst_ed == 0 && !self) {                         /* @0x100020f44-0x1000210db */
                acc = 0.0;
                for (int gm = 0; gm < 3; gm++) {
                    double t = in[1][s * 3 + gm];
                    for (int k = 2; k < ne; k++) {
                        double K = 0.0;
                        for (int gk = 0; gk < 3; gk++) K = K + TPROB[gk * 9 + gm * 3 + go] * in[k][s * 3 + gk];
                        t = t * K;
                    }
                    acc = acc + t;
                }
            } else if (dst_ed == 0) {                           /* @0x1000210e0-0x10002126e */
                acc = 1.0;
                for (int k = 2; k < ne; k++) {
                    double K = 0.0;
                    for (int gk = 0; gk < 3; gk++) K = K + TPROB[gk * 9 + go * 3 + go] * in[k][s * 3 + gk];
                    acc = acc * K;
                }
            } else if (dst_ed == 1) {                           /* @0x100021282-0x100021496 */
                acc = 0.0;
                for (int gp = 0; gp < 3; gp++) {
                    double t = in[0][s * 3 + gp];
                    for (int k = 2; k < ne; k++) {
                        double K = 0.0;
                        for (int gk = 0; gk < 3; gk++) K = K + TPROB[gk * 9 + go * 3 + gp] * in[k][s * 3 + gk];
                        t = t * K;
                    }
                    acc = acc + t;
                }
            } else if (!self) {                                 /* @0x1000214b8-0x100021764 */
                acc = 0.0;
                for (int gp = 0; gp < 3; gp++)
                    for (int gm = 0; gm < 3; gm++) {
                        double t = (TPROB[go * 9 + gm * 3 + gp] * in[0][s * 3 + gp]) * in[1][s * 3 + gm];
                        for (int k = 2; k < ne; k++) {
                            if (k == dst_ed) continue;
                            double K = 0.0;
                            for (int gk = 0; gk < 3; gk++) K = K + TPROB[gk * 9 + gm * 3 + gp] * in[k][s * 3 + gk];
                            t = t * K;
                        }
                        acc = acc + t;
                    }
            } else {                                            /* @0x100021769-0x1000219b6 */
                acc = 0.0;
                for (int g = 0; g < 3; g++) {
                    double t = TPROB[go * 9 + g * 3 + g] * in[0][s * 3 + g];
                    for (int k = 2; k < ne; k++) {
                        if (k == dst_ed) continue;
                        double K = 0.0;
                        for (int gk = 0; gk < 3; gk++) K = K + TPROB[gk * 9 + g * 3 + g] * in[k][s * 3 + gk];
                        t = t * K;
                    }
                    acc = acc + t;
                }
            }
            holder[s * 3 + go] = acc;
        }
    free((void *)in);
    return 0;
}

int pfo_eval_swaps(pfo_engine *e, int32_t chain, const pf_graph_view *v, int32_t n, const pf_swap *swaps, double *lsum)
{
    (void)chain;
    if (!e || !v || n < 0 || (n > 0 && (!swaps || !lsum))) FAIL(PF_EINVAL, "null argument");
    const int nI = v->n_indiv, nM = v->n_marr;
    const size_t T3 = (size_t)e->Sl * 3;
    int rc = PF_OK;
    swap_ctx c = {e, v, NULL, NULL, T3};
    c.mloc = malloc(sizeof(int) * (e->capM + 1));
    c.ia = calloc(nI + 1, sizeof(iadj));
    size_t nedges = 0;
    for (int j = 0; j < nM; j++) nedges += 2 + (v->marr_kid_begin[j + 1] - v->marr_kid_begin[j]);
    int *adj = malloc(sizeof(int) * 2 * (nedges + 1));
    double *h0 = malloc(sizeof(double) * T3), *h1 = malloc(sizeof(double) * T3), *tmp = malloc(sizeof(double) * T3);
    int registered = 0;
    if (!c.mloc || !c.ia || !adj || !h0 || !h1 || !tmp) { rc = PF_ENOMEM; snprintf(g_err, sizeof g_err, "oom"); goto done; }
    for (int m = 0; m < e->capM; m++) c.mloc[m] = -1;
    for (int i = 0; i < nI; i++) {
        const int slot = v->indiv[i];
        if (slot < 0 || slot >= e->capI || e->loc[slot] != -1) { rc = PF_EINVAL; snprintf(g_err, sizeof g_err, "bad or duplicate individual slot %d in view", slot); goto done; }
        e->loc[slot] = i; registered = i + 1;
    }
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            int *p = adj;
            for (int i = 0; i < nI; i++) { c.ia[i].mar = p; p += c.ia[i].nM; c.ia[i].edge = p; p += c.ia[i].nM; c.ia[i].nM = 0; }
        }
        for (int j = 0; j < nM; j++) {
            const int ne = 2 + (v->marr_kid_begin[j + 1] - v->marr_kid_begin[j]);
            if (v->marr[j] < 0 || v->marr[j] >= e->capM) { rc = PF_EINVAL; snprintf(g_err, sizeof g_err, "bad marriage slot"); goto done; }
            c.mloc[v->marr[j]] = j;
            for (int ed = 0; ed < ne; ed++) {
                if (ed == 1 && v->marr_selfing[j]) continue;
                const int slot = view_edge_slot(v, j, ed);
                if (slot < 0 || slot >= e->capI || e->loc[slot] < 0) { rc = PF_EINVAL; snprintf(g_err, sizeof g_err, "view not closed"); goto done; }
                const int i = e->loc[slot];
                if (pass == 1) { c.ia[i].mar[c.ia[i].nM] = j; c.ia[i].edge[c.ia[i].nM] = ed; }
                c.ia[i].nM++;
            }
        }
    }
    for (int q = 0; q < n; q++) {
        const pf_swap *w = &swaps[q];
        const int Mz = w->marr_z, Mx = w->marr_x, M0 = w->prev_marr, P = w->prev_pa, Q = w->prev_ma, x = w->x, z = w->z;
        if (Mz < 0 || Mz >= e->capM || c.mloc[Mz] < 0 || Mx < 0 || Mx >= e->capM || c.mloc[Mx] < 0 || (M0 != -1 && (M0 < 0 || M0 >= e->capM || c.mloc[M0] < 0)) ||
            x < 0 || x >= e->capI || e->loc[x] < 0 || z < 0 || z >= e->capI || e->loc[z] < 0 || P < 0 || P >= e->capI || e->loc[P] < 0 || Q < 0 || Q >= e->capI || e->loc[Q] < 0 ||
            w->kind < 1 || w->kind > 3 || (w->kind == 3 && !w->move_kids)) { rc = PF_EINVAL; snprintf(g_err, sizeof g_err, "swap %d references something outside the view", q); goto done; }
        if (v->marr_selfing[c.mloc[Mz]] || v->marr_selfing[c.mloc[Mx]] || (M0 != -1 && v->marr_selfing[c.mloc[M0]])) { rc = PF_EINVAL; snprintf(g_err, sizeof g_err, "swap %d touches a selfing marriage", q); goto done; }
        const int paz = v->marr_pa[c.mloc[Mz]], maz = v->marr_ma[c.mloc[Mz]];
        {
            int z_in = 0;
            for (int k = v->marr_kid_begin[c.mloc[Mz]]; k < v->marr_kid_begin[c.mloc[Mz] + 1]; k++) z_in = z_in || v->marr_kids[k] == z;
            if (!z_in) { rc = PF_EINVAL; snprintf(g_err, sizeof g_err, "swap %d: z is not a kid of marr_z", q); goto done; }
        }
        /* _ProposeSwapping@0x100022422-0x100022a9e: which messages feed the two holders */
        if (w->kind == 1 && !w->move_kids)      { init_swap_msg(&c, h0, M0, Mz, Mz, P, maz, z);   init_swap_msg(&c, h1, Mz, M0, M0, paz, Q, x); }
        else if (w->kind == 1)                  { init_swap_msg(&c, h0, M0, Mz, M0, P, maz, x);   init_swap_msg(&c, h1, Mz, M0, Mz, paz, Q, z); }
        else if (w->kind == 2 && !w->move_kids) { init_swap_msg(&c, h0, Mz, M0, Mz, paz, Q, z);   init_swap_msg(&c, h1, M0, Mz, M0, P, maz, x); }
        else if (w->kind == 2)                  { init_swap_msg(&c, h0, Mz, M0, M0, paz, Q, x);   init_swap_msg(&c, h1, M0, Mz, Mz, P, maz, z); }
        else                                    { init_swap_msg(&c, h0, M0, M0, M0, P, Q, x);     init_swap_msg(&c, h1, Mz, Mz, Mz, paz, maz, z); }
        directpass_iv2if(&c, h0, z, Mz, Mx);                    /* path index 1: through z           */
        if (directpass_if2iv(&c, h0, tmp, Mx, z, x)) { rc = PF_EINVAL; snprintf(g_err, sizeof g_err, "swap %d: z is not a member of marr_x", q); goto done; }
        directpass_iv2if(&c, h0, x, Mx, Mx);                    /* path index 2: through x           */
        double acc = 0.0;                                       /* _calculateBundlell@0x1000221b5-0x100022291 */
        for (int s = 0; s < e->Sl; s++) {
            double d = 0.0;
            for (int g = 0; g < 3; g++) d = d + h1[s * 3 + g] * h0[s * 3 + g];
            acc = acc + log(d);
        }
        lsum[q] = acc;
    }
done:
    for (int i = 0; i < registered; i++) e->loc[v->indiv[i]] = -1;
    free(c.mloc); free(c.ia); free(adj); free(h0); free(h1); free(tmp);
    return rc;
}

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

/* the two halves of pfo_prefilter for locus shards: dense partial scores (host memory here), then
 * the ranking of the summed scores with the same tie rule (earlier candidate first; the kid itself
 * is never a candidate) */
int pfo_prefilter_scores_dev(pfo_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids,
                             const uint8_t *as_father, int32_t n_cands, const int32_t *cands,
                             const double *cand_bias, double *scores)
{
    if (!e || (n_kids > 0 && (!kids || !as_father)) || (n_cands > 0 && !cands) || (n_kids > 0 && n_cands > 0 && !scores)) FAIL(PF_EINVAL, "bad argument");
    for (int k = 0; k < n_kids; k++)
        for (int c = 0; c < n_cands; c++) {
            pf_proposal pr = { PF_PROP_TRIO, -1, -1, -1 };
            if (as_father[k]) pr.pa = cands[c]; else pr.ma = cands[c];
            double v = 0.0;
            if (cands[c] != kids[k]) { int rc = pfo_eval(e, chain, kids[k], 1, &pr, &v); if (rc) return rc; }
            scores[(size_t)k * n_cands + c] = cand_bias ? v + cand_bias[c] : v;
        }
    return PF_OK;
}

int pfo_prefilter_topk_dev(pfo_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids,
                           int32_t n_cands, const int32_t *cands, double *scores, int32_t top_k,
                           int32_t *out_slot, double *out_score)
{
    (void)chain;
    if (!e || (n_kids > 0 && (!kids || !out_slot || !out_score)) || (n_cands > 0 && (!cands || !scores)) || top_k <= 0) FAIL(PF_EINVAL, "bad argument");
    for (int k = 0; k < n_kids; k++) {
        for (int t = 0; t < top_k; t++) { out_slot[k * top_k + t] = -1; out_score[k * top_k + t] = -INFINITY; }
        for (int c = 0; c < n_cands; c++) {
            if (cands[c] == kids[k]) continue;
            const double v = scores[(size_t)k * n_cands + c];
            int pos = top_k;
            while (pos > 0 && (out_slot[k * top_k + pos - 1] == -1 || v > out_score[k * top_k + pos - 1])) pos--;
            if (pos < top_k) {
                for (int t = top_k - 1; t > pos; t--) { out_slot[k * top_k + t] = out_slot[k * top_k + t - 1]; out_score[k * top_k + t] = out_score[k * top_k + t - 1]; }
                out_slot[k * top_k + pos] = cands[c]; out_score[k * top_k + pos] = v;
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

int pfo_io_bytes(const pfo_engine *e, uint64_t *h2d, uint64_t *d2h)
{
    (void)e;
    if (h2d) *h2d = 0;
    if (d2h) *d2h = 0;
    return 0;
}

/* The CUDA library queues the swap proposals behind the sweep and the first proposals and hands their
 * values over later; here the same entry points are the three calls in a row, the values kept until
 * asked for. */
int pfo_sweep_eval_swaps(pfo_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc, int32_t kid, int32_t n,
                         const pf_proposal *props, double *lsum, const pf_graph_view *swap_view, int32_t n_swaps, const pf_swap *swaps)
{
    if (!e || n_swaps < 0) FAIL(PF_EINVAL, "null argument");
    if (n_swaps > 0 && e->n_deferred > 0) FAIL(PF_EINVAL, "the swap results of an earlier pfo_sweep_eval_swaps were not collected");
    int rc = pfo_sweep_eval(e, chain, view, ll_cc, kid, n, props, lsum);
    if (rc != PF_OK || n_swaps == 0) return rc;
    double *d = (double *)realloc(e->deferred, (size_t)n_swaps * sizeof(double));
    if (!d) FAIL(PF_ENOMEM, "out of memory");
    e->deferred = d;
    rc = pfo_eval_swaps(e, chain, swap_view, n_swaps, swaps, d);
    if (rc == PF_OK) e->n_deferred = n_swaps;
    return rc;
}

int pfo_deferred_results(pfo_engine *e, int32_t n, double *lsum_swaps)
{
    if (!e || n < 0 || (n > 0 && !lsum_swaps)) FAIL(PF_EINVAL, "null argument");
    if (n != e->n_deferred) FAIL(PF_EINVAL, "%d deferred results asked for, %d are held", n, e->n_deferred);
    if (n > 0) memcpy(lsum_swaps, e->deferred, (size_t)n * sizeof(double));
    e->n_deferred = 0;
    return PF_OK;
}

int pfo_eval_swaps_batch(pfo_engine *e, int32_t n_sets, const int32_t *chains, const pf_graph_view *views,
                         const int32_t *swap_begin, const pf_swap *swaps, double *lsum)
{
    if (!e || n_sets < 0 || (n_sets > 0 && (!views || !swap_begin))) FAIL(PF_EINVAL, "null argument");
    for (int s = 0; s < n_sets; s++) {
        const int n = swap_begin[s + 1] - swap_begin[s];
        if (n < 0) FAIL(PF_EINVAL, "swap_begin not monotone");
        if (n == 0) continue;
        int rc = pfo_eval_swaps(e, chains ? chains[s] : 0, &views[s], n, swaps + swap_begin[s], lsum + swap_begin[s]);
        if (rc != PF_OK) return rc;
    }
    return PF_OK;
}
