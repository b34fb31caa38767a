// Tensor-core (tcgen05 / TMEM) full-rank scoring + top-k.  Placeholder until the
// UMMA kernel lands: it reports LGCN_ERR_UNSUPPORTED, it never falls back.
#include "common.cuh"

namespace lgcn {

int score_topk_tc(const float*, const float*, const int64_t*, int64_t, int64_t, int, const int64_t*,
                  const int32_t*, int, float, int32_t*, float*, cudaStream_t) {
  set_last_error("precision LGCN_BF16 (tcgen05 path) is not built yet");
  return LGCN_ERR_UNSUPPORTED;
}

}  // namespace lgcn
