// Shared by the per-sample loss kernels (bpr.cu, ssm.cu): CTA size and the deterministic final
// reduction of the per-sample loss / regulariser terms.
#pragma once
#include "common.cuh"

namespace lgcn {

constexpr int kBprBlock = 256;

// Deterministic final reduction of the per-sample loss / reg terms by the last CTA to arrive.
__device__ __forceinline__ void bpr_finish(int batch, float inv_b, float decay, float* __restrict__ loss_out,
                                           float* __restrict__ work, int32_t* __restrict__ work_counter) {
  __shared__ int s_last;
  __shared__ float s_red[2][kBprBlock / 32];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int old = atomicAdd(work_counter, 1);
    s_last = (old == (int)gridDim.x - 1);
    if (s_last) *work_counter = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float a = 0.f, r = 0.f;
  for (int i = threadIdx.x; i < batch; i += kBprBlock) {
    a += __ldcg(work + i);
    r += __ldcg(work + batch + i);
  }
  a = group_sum<32>(a, 0xffffffffu);
  r = group_sum<32>(r, 0xffffffffu);
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = a;
    s_red[1][threadIdx.x >> 5] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.f, tr = 0.f;
    for (int w = 0; w < kBprBlock / 32; ++w) {
      ta += s_red[0][w];
      tr += s_red[1][w];
    }
    const float loss = ta * inv_b;
    const float reg = 0.5f * tr * inv_b;
    loss_out[0] = loss;
    loss_out[1] = reg;
    loss_out[2] = loss + decay * reg;
    loss_out[3] += loss + decay * reg;  // running epoch sum (OneEpoch, model/lgcn.py:149)
  }
}

}  // namespace lgcn
