// Ranking metrics on device: hit test of the top-k lists against the test CSR and
// the recall / precision / hit-rate / NDCG sums.
//
// Replaces utils.getLabel (reference utils.py:40-48, a Python `in list` per
// element), RecallPrecision_ATk (metric.py:60-72) and NDCGatK_r (metric.py:84-103);
// the caller divides the sums by len(users) like trainer.py:169-170.
#include "common.cuh"

namespace lgcn {

constexpr int kMaxKs = 8;

__device__ __forceinline__ bool sorted_contains(const int32_t* __restrict__ a, int n, int32_t x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo < n && __ldg(a + lo) == x;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct KsArg { int ks[kMaxKs]; int n; };

__global__ void __launch_bounds__(128)
rank_metrics_kernel(const int32_t* __restrict__ topk, int n_eval, int k,
                    const int64_t* __restrict__ user_ids, const int64_t* __restrict__ test_rowptr,
                    const int32_t* __restrict__ test_sorted, KsArg ks, double* __restrict__ sums,
                    uint8_t* __restrict__ hits) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[4][kMaxKs];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int i = 0; i < kMaxKs; ++i) acc[m][i] = 0.0;

  if (r < n_eval) {
    const int64_t u = user_ids[r];
    const int64_t b = test_rowptr[u];
    const int n_gt = (int)(test_rowptr[u + 1] - b);
    const int32_t* gt = test_sorted + b;
    double right = 0.0, dcg = 0.0, idcg = 0.0;
    int next = 0;
    for (int j = 0; j < k; ++j) {
      const int32_t item = topk[(int64_t)r * k + j];
      const bool hit = item >= 0 && sorted_contains(gt, n_gt, item);
      if (hits != nullptr) hits[(int64_t)r * k + j] = hit ? 1 : 0;
      const double disc = 1.0 / log2((double)(j + 2));
      if (hit) { right += 1.0; dcg += disc; }
      if (j < n_gt) idcg += disc;
      // close every cut-off that ends at position j
      while (next < ks.n && ks.ks[next] == j + 1) {
        acc[0][next] = right / ((double)n_gt + 1e-6);           // metric.py:69
        acc[1][next] = right / (double)(j + 1);                 // metric.py:70
        acc[2][next] = right >= 1.0 ? 1.0 : 0.0;                // metric.py:71
        acc[3][next] = dcg / (idcg == 0.0 ? 1.0 : idcg);        // metric.py:97-102
        ++next;
      }
    }
  }

  __shared__ double s_part[4][kMaxKs][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int m = 0; m < 4; ++m)
    for (int i = 0; i < ks.n; ++i) {
      const double v = warp_sum_d(acc[m][i]);
      if (lane == 0) s_part[m][i][warp] = v;
    }
  __syncthreads();
  if (threadIdx.x < 4 * kMaxKs) {
    const int m = threadIdx.x / kMaxKs, i = threadIdx.x % kMaxKs;
    if (i < ks.n) {
      const double v = s_part[m][i][0] + s_part[m][i][1] + s_part[m][i][2] + s_part[m][i][3];
      atomicAdd(sums + m * ks.n + i, v);
    }
  }
}

}  // namespace lgcn

using namespace lgcn;

extern "C" int lgcn_rank_metrics(const int32_t* topk, int64_t n_eval, int k,
                                 const int64_t* user_ids, const int64_t* test_rowptr,
                                 const int32_t* test_sorted, const int32_t* ks_host, int n_ks,
                                 double* sums, uint8_t* hits, lgcn_stream_t stream) {
  LGCN_CHECK_ARG(topk && user_ids && test_rowptr && test_sorted && ks_host && sums,
                 "null pointer argument");
  LGCN_CHECK_ARG(n_eval >= 0 && n_eval < 0x7fffffffLL, "n_eval out of range");
  LGCN_CHECK_ARG(n_ks >= 1 && n_ks <= kMaxKs, "n_ks must be in [1, %d]", kMaxKs);
  KsArg ks;
  ks.n = n_ks;
  for (int i = 0; i < n_ks; ++i) {
    ks.ks[i] = ks_host[i];
    LGCN_CHECK_ARG(ks_host[i] >= 1 && ks_host[i] <= k, "cut-off %d outside [1, k=%d]", ks_host[i], k);
    LGCN_CHECK_ARG(i == 0 || ks_host[i] > ks_host[i - 1], "cut-offs must be strictly ascending");
  }
  if (n_eval == 0) return 0;
  const int grid = (int)((n_eval + 127) / 128);
  rank_metrics_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(topk, (int)n_eval, k, user_ids,
                                                              test_rowptr, test_sorted, ks, sums, hits);
  LGCN_LAUNCH_OK();
  return 0;
}
