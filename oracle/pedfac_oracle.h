/*
 * pedfac_oracle.h — CPU restatement (plain C, fp64) of the pedFac `pedigraph` likelihood layer.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in pedfac_b200/ (the product) may link, import or execute
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, as the checker or as the timed CPU baseline.
 *
 * PARITY UNPINNED BY THE REFERENCE: ngthomas/pedFac ships neither source nor tests nor golden
 * outputs for this path; exec/pedigraph is a macOS Mach-O binary that
 * cannot run here. This file restates the algorithm recovered from that binary's -O0
 * disassembly, function by function, in the reference's operation order (citations:
 * pedigraph@<VM address> <symbol>). It is pinned by (1) the 27 doubles of _TPROB extracted
 * from the binary (tests/golden/tprob.json), (2) exhaustive enumeration over all joint
 * genotypes of small pedigrees (tests/bruteforce.py), (3) the SURVEY.md App. E known answers.
 *
 * The entry points mirror include/pedfac_cuda.h one to one (pfo_ instead of pf_), single chain.
 */
#ifndef PEDFAC_ORACLE_H
#define PEDFAC_ORACLE_H

#include "../include/pedfac_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pfo_engine pfo_engine;

const char *pfo_last_error(void);
const double *pfo_tprob(void); /* 27 doubles, [gk][gm][gp] */

int pfo_create(pfo_engine **out, const pf_config *cfg);
int pfo_destroy(pfo_engine *e);
/* no device: nothing is ever copied (keeps the sampler source identical for both backends) */
int pfo_io_bytes(const pfo_engine *e, uint64_t *h2d, uint64_t *d2h);
int pfo_set_locus_params(pfo_engine *e, const double *eps, const double *afreq);
int pfo_set_genotypes_hard(pfo_engine *e, int32_t slot, const uint8_t *codes);
int pfo_set_genotypes_soft(pfo_engine *e, int32_t slot, const float *lik);
int pfo_set_genotypes_soft_f64(pfo_engine *e, int32_t slot, const double *lik);
int pfo_set_genotypes_soft_dec16(pfo_engine *e, int32_t slot, const uint16_t *num, uint32_t denom);
int pfo_set_genotypes_unobserved(pfo_engine *e, int32_t slot);
int pfo_sweep(pfo_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc);
int pfo_eval(pfo_engine *e, int32_t chain, int32_t kid, int32_t n, const pf_proposal *props,
             double *lsum);
int pfo_sweep_eval(pfo_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc,
                   int32_t kid, int32_t n, const pf_proposal *props, double *lsum);
int pfo_sweep_eval_swaps(pfo_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc, int32_t kid, int32_t n,
                         const pf_proposal *props, double *lsum, const pf_graph_view *swap_view, int32_t n_swaps, const pf_swap *swaps);
int pfo_eval_swaps_batch(pfo_engine *e, int32_t n_sets, const int32_t *chains, const pf_graph_view *views,
                         const int32_t *swap_begin, const pf_swap *swaps, double *lsum);
int pfo_deferred_results(pfo_engine *e, int32_t n, double *lsum_swaps);
int pfo_eval_swaps(pfo_engine *e, int32_t chain, const pf_graph_view *view, int32_t n,
                   const pf_swap *swaps, double *lsum);
int pfo_prefilter(pfo_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids,
                  const uint8_t *as_father, int32_t n_cands, const int32_t *cands,
                  const double *cand_bias, int32_t top_k, int32_t *out_slot, double *out_score);
/* the oracle has no device: the *_dev pointers are host pointers here */
int pfo_prefilter_scores_dev(pfo_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids,
                             const uint8_t *as_father, int32_t n_cands, const int32_t *cands,
                             const double *cand_bias, double *scores);
int pfo_prefilter_topk_dev(pfo_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids,
                           int32_t n_cands, const int32_t *cands, double *scores, int32_t top_k,
                           int32_t *out_slot, double *out_score);
int pfo_get_stub(pfo_engine *e, int32_t chain, int32_t slot, double *out);
int pfo_get_prong(pfo_engine *e, int32_t chain, int32_t marr, double *out);

#ifdef __cplusplus
}
#endif
#endif
