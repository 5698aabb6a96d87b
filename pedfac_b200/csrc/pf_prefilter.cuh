// pf_prefilter.cuh — dense kid x candidate single-parent scan (_RefineParentOffPair@0x100014d20).
#pragma once
#include "pf_device.cuh"

namespace pf {

constexpr int PREF_MAX_TOPK = 32;

struct PrefilterParams {
    DevView D;
    const int *kid_rows;      // [n_kids] stub rows of the focal kids
    const int *cand_rows;     // [n_cands] stub rows of the candidate parents
    const double *cand_bias;  // [n_cands]
    int n_kids, n_cands, top_k;
    double *scores;           // [n_kids][n_cands] scratch
    double *out_score;        // [n_kids][top_k]
    int *out_idx;             // [n_kids][top_k] index into cand_rows or -1
};

// each returns the number of kernels launched
int launch_prefilter_scores(const PrefilterParams &P, cudaStream_t st);
int launch_prefilter_topk(const PrefilterParams &P, cudaStream_t st);

} // namespace pf
