/*
 * pedfac_cuda.h — C ABI of libpedfac_cuda, the B200 (sm_100a) likelihood engine for the
 * pedFac `pedigraph` sampler's hot path: per-locus sum-product over the pedigree factor
 * graph and per-proposal log-likelihood evaluation.
 *
 * The reference has no FFI seam: its MCMC driver (L2) calls the likelihood
 * layer (L1) directly on a `ped` struct. Each entry point below replaces one L2→L1 call
 * site of the reference binary `exec/pedigraph` (cited as pedigraph@<VM address> <symbol>).
 *
 * Conventions: plain C; every function returns PF_OK (0) or a negative PF_E* code and
 * leaves a message retrievable with pf_last_error(); all pointer arguments are caller-owned
 * HOST buffers unless the name ends in `_dev`; an engine is not thread-safe; there is no CPU
 * fallback — without a usable CUDA device pf_create fails.
 *
 * Indexing: "slot" = index into the reference's ped->indiv[] / ped->mnode[] arrays
 * (observed individuals first, unobserved after). Genotype state
 * g in {0,1,2} = copies of the "1" allele. Triples are stored [locus][g] on the host side.
 */
#ifndef PEDFAC_CUDA_H
#define PEDFAC_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_OK            0
#define PF_EINVAL      (-1)   /* bad argument / malformed graph view */
#define PF_ECUDA       (-2)   /* CUDA runtime error (message has the detail) */
#define PF_ENOMEM      (-3)
#define PF_ENODEVICE   (-4)   /* no usable GPU: there is deliberately no CPU fallback */
#define PF_ECYCLE      (-5)   /* the graph view is not a forest (reference: "exceeds the acyclic upper limit") */

typedef struct pf_engine pf_engine;

typedef struct pf_config {
    int32_t n_loci;       /* S: loci in the data set (prior.txt nSNP) */
    int32_t locus_begin;  /* this engine evaluates loci [locus_begin, locus_end) — the */
    int32_t locus_end;    /*   multi-GPU shard; (0, S) for a single GPU                */
    int32_t cap_indiv;    /* individual slots (ped->capIndiv, _CalculateMaxSize@0x100007ba5) */
    int32_t cap_marr;     /* marriage slots */
    int32_t n_chains;     /* independent MCMC chains sharing the genotype arrays (>= 1) */
    int32_t device;       /* CUDA device ordinal */
    int32_t reserved;
} pf_config;

/* Proposal kinds. One proposal ≙ one entry appended to the reference's `choice` list. */
enum {
    PF_PROP_TRIO    = 0,  /* _CalculateLikeByStubs@0x10001f260 / _CalculateLLByStubs@0x10001eef0:
                             kid gets parents (pa, ma); -1 = a new unobserved founder (HWE prior) */
    PF_PROP_SELFING = 1,  /* _CalculateLikeByStubsSelfing@0x10001f6d0: one parent `pa` (or -1) selfs */
    PF_PROP_PRONG   = 2   /* _CalculateLikeByProngs@0x10001fa20: kid joins existing marriage `marr` */
};

typedef struct pf_proposal {
    int32_t kind;
    int32_t pa;    /* individual slot or -1 */
    int32_t ma;    /* individual slot or -1 */
    int32_t marr;  /* marriage slot (PF_PROP_PRONG) */
} pf_proposal;

/*
 * The sub-forest to (re)propagate: exactly the two worklists _TurnOnSpotlight@0x10001bb90
 * builds (active individuals whose component is dirty; active marriages whose father's
 * component is dirty) plus the adjacency the message passes read through ped->mnode[].
 * The view must be closed: every individual referenced by a listed marriage is listed.
 */
typedef struct pf_graph_view {
    int32_t        n_indiv;
    const int32_t *indiv;          /* [n_indiv] slots                                         */
    const int32_t *indiv_cc;       /* [n_indiv] component label, value in [0, n_cc)           */
    const uint8_t *indiv_founder;  /* [n_indiv] indiv->isFounder (p-node sends the HWE prior) */
    int32_t        n_marr;
    const int32_t *marr;           /* [n_marr] marriage slots                                 */
    const int32_t *marr_pa;        /* [n_marr] father slot (edge 0)                           */
    const int32_t *marr_ma;        /* [n_marr] mother slot (edge 1); ignored when selfing     */
    const uint8_t *marr_selfing;   /* [n_marr] mnode->selfing                                 */
    const int32_t *marr_kid_begin; /* [n_marr + 1] CSR offsets into marr_kids (edges 2..)     */
    const int32_t *marr_kids;
    int32_t        n_cc;           /* labels used by indiv_cc are 0..n_cc-1 (local to the view) */
} pf_graph_view;

const char *pf_last_error(void);

/* Engine lifetime. ≙ _AllocIndiv@0x100007dd0 (storage), minus every per-edge message array. */
int pf_create(pf_engine **out, const pf_config *cfg);
int pf_destroy(pf_engine *e);

/* Run the engine's work on an existing CUDA stream (cudaStream_t passed as void*); default:
 * a private non-blocking stream. */
int pf_set_stream(pf_engine *e, void *cuda_stream);

/* Per-locus parameters: genotyping error eps[S] and "1"-allele frequency afreq[S]
 * (prior.txt rows `epsilon`, `aFreq`). The HWE founder prior is built as
 * _ReadPedfile@0x1000099d2-0x100009ad4 does: q = 1-a; [q*q, (2*(1-q))*q, (1-q)*(1-q)]. */
int pf_set_locus_params(pf_engine *e, const double *eps, const double *afreq);

/* Genotype data of one individual slot, all S loci of the data set (the engine keeps its
 * shard). ≙ the emission construction in _ReadPedfile@0x10000b1ca-0x10000b6e2.
 *   hard:  codes[s] in {0,1,2,3}; 3 = missing -> [1,1,1]; c -> e[c] = 1-eps[s], others eps[s]/2.
 *          Stored as 2-bit codes.
 *   soft (f32):  lik[s][3] genotype likelihoods stored as fp32 triples; emission = (double)lik.
 *   soft (f64):  exact doubles (what strtod returned in the reference).
 *   soft (dec16): num[s][3] / denom, the exact value of a decimal token with <= 4 fractional
 *          digits (R writes round(x, 4): R/renderGeno.R:114-118); 6 bytes per locus.
 *   unobserved: emission [1,1,1] (_AttachUnobsIndiv@0x100012cfa). */
int pf_set_genotypes_hard(pf_engine *e, int32_t slot, const uint8_t *codes);
int pf_set_genotypes_soft(pf_engine *e, int32_t slot, const float *lik);
int pf_set_genotypes_soft_f64(pf_engine *e, int32_t slot, const double *lik);
int pf_set_genotypes_soft_dec16(pf_engine *e, int32_t slot, const uint16_t *num, uint32_t denom);
int pf_set_genotypes_unobserved(pf_engine *e, int32_t slot);

/* ≙ _Propagating_msg@0x10001ba10 restricted to the view: recomputes the stub of every listed
 * individual (_FillInStubs@0x10001e1a0), the prong of every listed marriage
 * (_FillInProngs@0x10001dd70) and, per component label c, ll_cc[c] = sum over this engine's
 * loci of log Z (_UpdateProb@0x10001da20). ll_cc has view->n_cc entries. With a locus shard
 * the values are partial sums to be added across shards. */
int pf_sweep(pf_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc);

/* ≙ the CalculateLikeBy* call sites in _EvalMoves@0x100015010 / _ProposeSingletMove@0x1000149d0:
 * for focal individual `kid` and each proposal p, lsum[p] = sum over this engine's loci of
 * log(contraction) — the data term of the reference's ll. The caller adds the component
 * bookkeeping with pf_compose_ll(). */
int pf_eval(pf_engine *e, int32_t chain, int32_t kid, int32_t n, const pf_proposal *props,
            double *lsum);

/* pf_sweep followed by pf_eval in one submission with a single host wait: what a Gibbs step does
 * first (_Propagating_msg@0x10001ba10 after _PruneIndiv, then the single-parent proposals of
 * _EvalMoves@0x100015010, whose enumeration depends on the graph only). Same results as the two
 * calls; n may be 0. */
int pf_sweep_eval(pf_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc,
                  int32_t kid, int32_t n, const pf_proposal *props, double *lsum);
/* The same with the results left in device memory on the engine's stream:
 * out_dev = [ll_cc (view->n_cc) | lsum (n)]; on a sharded engine one cross-GPU sum covers both. */
int pf_sweep_eval_dev(pf_engine *e, int32_t chain, const pf_graph_view *view, int32_t kid, int32_t n,
                      const pf_proposal *props, double *out_dev);

/* Same as pf_sweep / pf_eval but results stay on the device (ll_cc_dev / lsum_dev are device
 * pointers, written on the engine's stream, no host synchronisation): the form used when an
 * NCCL all-reduce over locus shards follows. */
int pf_sweep_dev(pf_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc_dev);
int pf_eval_dev(pf_engine *e, int32_t chain, int32_t kid, int32_t n, const pf_proposal *props,
                double *lsum_dev);

/* Batched forms: one launch sequence for several chains (config D: 64 chains per GPU).
 * views[i] is swept on chains[i]; ll_cc receives the concatenation of the per-view outputs.
 * Evals: request i = (chains[i], kids[i], props[prop_begin[i] .. prop_begin[i+1])). */
int pf_sweep_batch(pf_engine *e, int32_t n, const int32_t *chains, const pf_graph_view *views,
                   double *ll_cc);
int pf_eval_batch(pf_engine *e, int32_t n, const int32_t *chains, const int32_t *kids,
                  const int32_t *prop_begin, const pf_proposal *props, double *lsum);
/* device-output forms of the batched calls (no host synchronisation) */
int pf_sweep_batch_dev(pf_engine *e, int32_t n, const int32_t *chains, const pf_graph_view *views,
                       double *ll_cc_dev);
int pf_eval_batch_dev(pf_engine *e, int32_t n, const int32_t *chains, const int32_t *kids,
                      const int32_t *prop_begin, const pf_proposal *props, double *lsum_dev);

/*
 * Swap proposals ≙ _ProposeSwapping@0x1000223d0 with its helpers _initSwapMsg@0x100021cd0,
 * _directpassIV2IF@0x100020810, _directpassIF2IV@0x100020a70 and the data term of
 * _calculateBundlell@0x1000220e0 (enumerated by _EvalMoves@0x100015989-0x100016651).
 *
 * x is the focal individual (already pruned from its previous parents prev_pa / prev_ma; prev_marr is
 * their marriage if it still exists, i.e. x had siblings, else -1), z a spouse of x in marriage
 * marr_x, and marr_z the marriage in which z is a kid, with parents (pa_z, ma_z). With K0 = the kids
 * left in prev_marr and Kz = the siblings of z, the proposal scores the pedigree in which
 *     kind 1: the fathers are exchanged   marr_z = (prev_pa, ma_z, .)   x's family = (pa_z, prev_ma, .)
 *     kind 2: the mothers are exchanged   marr_z = (pa_z, prev_ma, .)   x's family = (prev_pa, ma_z, .)
 *     kind 3: both                        marr_z = (prev_pa, prev_ma, .) x's family = (pa_z, ma_z, .)
 * and the sibships stay (move_kids = 0: marr_z keeps Kz + z, x joins K0; only kinds 1 and 2) or are
 * exchanged too (move_kids = 1: marr_z gets K0 + z, x's family gets Kz + x; always for kind 3).
 * The reference's kind 4 (two spouses) cannot be reached: its enumeration looks up the parents of x
 * itself, which has none at that point (_EvalMoves@0x100016136-0x100016292).
 *
 * `view` must hold the current components of x, prev_pa and prev_ma (the messages of the last
 * sweep of those components are the inputs, as in the reference). lsum[p] = sum over this engine's
 * loci of log sum_g h0[g]*h1[g] = log Z of the merged, rewired component. The engine evaluates it as
 * an upward-only sweep of that hypothetical component; nothing persisted is modified. The caller
 * adds the bookkeeping with pf_compose_swap_ll().
 */
typedef struct pf_swap {
    int32_t kind;       /* 1, 2 or 3 */
    int32_t move_kids;  /* 0 / 1 (the reference's choice aux[1]: -1 / z) */
    int32_t x, z;       /* individual slots */
    int32_t marr_z;     /* marriage where z is a kid */
    int32_t marr_x;     /* marriage of x and z */
    int32_t prev_pa, prev_ma;
    int32_t prev_marr;  /* or -1 */
    int32_t reserved;
} pf_swap;

int pf_eval_swaps(pf_engine *e, int32_t chain, const pf_graph_view *view, int32_t n,
                  const pf_swap *swaps, double *lsum);
int pf_eval_swaps_dev(pf_engine *e, int32_t chain, const pf_graph_view *view, int32_t n,
                      const pf_swap *swaps, double *lsum_dev);

/* Several lists of swap proposals in one submission (the multi-chain sampler: one list per chain that
 * is at its swap block): list s is swaps[swap_begin[s] .. swap_begin[s+1]) on chain chains[s]
 * (NULL = chain 0) and view views[s]; lsum is indexed like swaps. Values identical to pf_eval_swaps. */
int pf_eval_swaps_batch(pf_engine *e, int32_t n_sets, const int32_t *chains, const pf_graph_view *views,
                        const int32_t *swap_begin, const pf_swap *swaps, double *lsum);

/* The whole first submission of a Gibbs step: pf_sweep_eval, and behind it the swap proposals of the
 * step (≙ the swap block of _EvalMoves@0x100015989-0x100016651, whose enumeration also depends on the
 * graph only) as in pf_eval_swaps on `swap_view`. ll_cc and lsum are returned as soon as the proposal
 * kernel has finished; the swap proposals are planned on the host while the sweep runs, execute while
 * the caller prepares its next call, and their n_swaps values are DEFERRED: they travel in the same
 * delivery as the results of the next host-result call on this engine (pf_eval, pf_sweep, ...) and
 * are then read with pf_deferred_results, which delivers them itself if no call has carried them.
 * At most one set of deferred values may be outstanding. Values identical to pf_eval_swaps. */
int pf_sweep_eval_swaps(pf_engine *e, int32_t chain, const pf_graph_view *view, double *ll_cc,
                        int32_t kid, int32_t n, const pf_proposal *props, double *lsum,
                        const pf_graph_view *swap_view, int32_t n_swaps, const pf_swap *swaps);
int pf_deferred_results(pf_engine *e, int32_t n, double *lsum_swaps);

/* ≙ _RefineParentOffPair@0x100014d20 (its N_obs^2 x S _CalculateLLByStubs calls), dense form:
 * for every kid k and every candidate c (c != kid) the score
 *     lsum(kid k; c as the single given parent, the other parent a new founder) + cand_bias[c]
 * where cand_bias (may be NULL = 0) carries the -LLcc terms of _CalculateLLByStubs@0x10001f183-0x10001f241.
 * c is the father if as_father[k] else the mother (the value is the same: _TPROB is symmetric
 * in the parents). Reduced on the device to the top_k highest scores per kid (ties: earlier
 * candidate first; top_k <= 32). out_slot/out_score are [n_kids][top_k]; unused entries are
 * -1 / -inf. The caller groups kids that share a candidate set (same generation time: the age
 * window test of @0x100014dbf-0x100014def). On a locus shard this needs the communicator
 * (pf_comm_connect): the dense partial scores are summed over the shards before the ranking. */
int pf_prefilter(pf_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids,
                 const uint8_t *as_father, int32_t n_cands, const int32_t *cands,
                 const double *cand_bias, int32_t top_k, int32_t *out_slot, double *out_score);

/* pf_prefilter in two halves for engines that hold a locus shard: every shard writes
 * its dense partial scores [n_kids][n_cands] (sum over its loci; cand_bias is added as given — pass it
 * on one shard only, NULL on the others) to device memory on the engine's stream, the caller sums
 * them over the shards (NCCL all-reduce), and any shard ranks the sums. scores_dev is overwritten by
 * the ranking. */
int pf_prefilter_scores_dev(pf_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids,
                            const uint8_t *as_father, int32_t n_cands, const int32_t *cands,
                            const double *cand_bias, double *scores_dev);
int pf_prefilter_topk_dev(pf_engine *e, int32_t chain, int32_t n_kids, const int32_t *kids,
                          int32_t n_cands, const int32_t *cands, double *scores_dev, int32_t top_k,
                          int32_t *out_slot, double *out_score);

/* Read-back of persisted per-locus results (tests, debugging): out[s][3] for this engine's loci. */
int pf_get_stub(pf_engine *e, int32_t chain, int32_t slot, double *out);
int pf_get_prong(pf_engine *e, int32_t chain, int32_t marr, double *out);

/*
 * Locus shards on several GPUs of one node, one process per GPU (the reference has no
 * counterpart: it is one thread on one core). Every rank creates its engine with its own
 * [locus_begin, locus_end) and makes the identical sequence of calls. After pf_comm_connect every
 * result that is a sum over loci (ll_cc, lsum, swap sums, prefilter scores) is the sum over ALL
 * shards, bit-identical on every rank: the library adds the ranks' partial vectors itself, in rank
 * order, inside one kernel that writes to the peers' mailboxes over NVLink (no NCCL, no host hop).
 * Set-up: each rank calls pf_comm_local_handle (allocates its mailbox, returns an opaque
 * PF_COMM_HANDLE_BYTES-byte handle = cudaIpcMemHandle_t), the caller exchanges the handles between
 * the processes by any means (pedigraph: files in the run directory; Python: torch.distributed),
 * then pf_comm_connect(rank, world, handles of ranks 0..world-1 concatenated). world <= 8.
 */
#define PF_COMM_HANDLE_BYTES 64
int pf_comm_local_handle(pf_engine *e, int32_t world, void *handle_out);
int pf_comm_connect(pf_engine *e, int32_t rank, int32_t world, const void *handles);
int64_t pf_comm_reduce_count(const pf_engine *e);   /* cross-GPU sums performed so far */
/* Optional overlap for the device-buffer (_dev) entry points: with asynchronous sums enabled their
 * cross-GPU sums run on a second stream, ordered after the kernels that produce their input, while
 * the engine's stream goes on with the next call; the exchange is pure latency (and waits for the
 * slowest rank), so this hides it. The caller must not read or reuse an output buffer before
 * pf_comm_fence, which makes the engine's stream wait for every sum issued so far. */
int pf_comm_set_async(pf_engine *e, int32_t enable);
int pf_comm_fence(pf_engine *e);

/* Counters: kernels launched by this engine since creation. */
int64_t pf_launch_count(const pf_engine *e);
/* Counters: bytes the host-buffer entry points copied to the device (plans, views, proposal lists)
 * and back (results) since creation — what bench.py reports as h2d / d2h bytes per step. */
int pf_io_bytes(const pf_engine *e, uint64_t *h2d, uint64_t *d2h);

/* Measurement aid: while enabled, every kernel launch is bracketed by CUDA events on the
 * engine's stream. pf_profile_read synchronises, returns the summed device time in
 * milliseconds and the launch count per kernel kind (0 sweep, 1 eval, 2 reduce, 3 prefilter;
 * ms[4], count[4]) and clears the records. */
/* Shape of the most recent sweep plan, out[8] = { levels, task groups, tasks, shared-memory bytes
 * per block, transient message rows, of which in shared memory, up-band levels, locus tiles }. */
int pf_plan_stats(const pf_engine *e, int32_t *out);
int pf_profile(pf_engine *e, int32_t enable);
int pf_profile_read(pf_engine *e, double *ms, int64_t *count);

/*
 * The scalar bookkeeping the reference wraps around the data term (host, fp64), in the
 * reference's order (_CalculateLikeByStubs@0x10001f5df-0x10001f6ae):
 *   ll = ((-LLcc[cc(kid)] + -(use_pa)*LLcc[cc(pa)]) + -(use_ma)*LLcc[cc(ma)]) + (LLtot + lsum)
 * use_pa = pa != -1 && cc(pa) != cc(kid); use_ma = ma != -1 && cc(ma) != cc(pa) && cc(ma) != cc(kid).
 */
static inline double pf_compose_ll(double ll_tot, double lsum, double ll_cc_kid,
                                   int use_pa, double ll_cc_pa, int use_ma, double ll_cc_ma)
{
    double a = -ll_cc_kid;
    a = a + (use_pa ? -ll_cc_pa : 0.0);
    a = a + (use_ma ? -ll_cc_ma : 0.0);
    return a + (ll_tot + lsum);
}

/* Bookkeeping of a swap proposal, in the order of _calculateBundlell@0x100022291-0x1000223a8:
 *   ll = ((-LLcc[cc(x)] + -LLcc[cc(prev_pa)]) + (double)(-use_ma) * LLcc[cc(prev_ma)]) + (LLtot + lsum)
 * use_ma = cc(prev_pa) != cc(prev_ma) && cc(prev_pa) != cc(x). */
static inline double pf_compose_swap_ll(double ll_tot, double lsum, double ll_cc_x, double ll_cc_pa,
                                        int use_ma, double ll_cc_ma)
{
    const double a = (-ll_cc_x) + (-ll_cc_pa);
    return (a + (double)(-(use_ma ? 1 : 0)) * ll_cc_ma) + (ll_tot + lsum);
}

#ifdef __cplusplus
}
#endif
#endif /* PEDFAC_CUDA_H */
