// mle.cuh — batched maximum-likelihood tree-scale fit (score-msa --strategy mle).
// Placeholder until the batched Brent / batched expm kernels land.
#pragma once

#include <string>

#include "../../include/phylocsf_b200.h"
#include "kernels.cuh"

namespace pcsf {

struct MleBatch {
    int n_aln;
    const int64_t *d_win_start, *d_col_start, *d_len;
    WinSpace ws;
    int64_t nwin;
    bool want_anc;
    float *d_phylo, *d_anc;
};

struct DevBufFwd;

inline pcsf_status mle_setup(const ModelHost &) { return PCSF_OK; }

template <class Buf>
inline pcsf_status mle_run(const ModelHost &, const MleBatch &, double *const *, const float *, const int32_t *,
                           const int32_t *, Buf &, int, cudaStream_t, std::string &err) {
    err = "score-msa MLE strategy is not built yet";
    return PCSF_ERR_UNSUPPORTED;
}

}  // namespace pcsf
