// Full-rank evaluation: user x item scores fused with train-positive masking and
// a per-row top-k, so the [U, m] score matrix of the reference never exists.
//
// Replaces getUsersRating's matmul (reference model/lgcn.py:124), the host-built
// exclude lists + index_put of -(1<<10) (trainer.py:132-137), torch.topk
// (trainer.py:138) and the discarded 1.6 GB device->host copy (trainer.py:139).
//
// This file holds the EXACT fp32 path (precision = LGCN_F32): scores are fp32 FMA
// dot products, so the selected ids can be checked bit-for-bit against a stable
// sort of the same scores.  The tensor-core path lives in score_topk_tc.cu.
//
// Selection: one thread owns one user row for the whole sweep over the items.
// Items arrive in ascending id order, so a candidate can enter the top-k only if
// it is STRICTLY greater than the current k-th value (equal scores lose the tie
// to the earlier, lower id).  Candidates that beat the running threshold are
// appended to a per-row shared-memory buffer; when the buffer is full the row is
// re-selected down to k entries (ordered by score desc, id asc) and the threshold
// is raised.  Expected appends per row are O(k log(m/k)), not O(m).
#include "common.cuh"

namespace lgcn {

int score_topk_tc(const float* user_emb, const float* item_emb, const int64_t* user_ids,
                  int64_t n_eval, int64_t m_items, int d, const int64_t* pos_rowptr,
                  const int32_t* pos_sorted, int k, float mask_value, int32_t* out_idx,
                  float* out_val, float* dense, void* workspace, size_t workspace_bytes, int acc16,
                  cudaStream_t st);
size_t score_topk_tc_workspace(int64_t n_eval, int64_t m_items, int d);

constexpr int kTU = 64;    // users per CTA
constexpr int kTI = 128;   // items per tile
constexpr int kThreads = 256;

__device__ __forceinline__ bool pos_contains(const int32_t* __restrict__ a, int n, int32_t x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo < n && __ldg(a + lo) == x;
}

// Re-select row `row`'s buffer down to min(k, cnt) entries sorted by (val desc, idx asc).
__device__ __forceinline__ void reselect(float* cval, int* cidx, int row, int& cnt, int k) {
  const int kk = k < cnt ? k : cnt;
  for (int r = 0; r < kk; ++r) {
    int best = r;
    float bv = cval[r * kTU + row];
    int bi = cidx[r * kTU + row];
    for (int j = r + 1; j < cnt; ++j) {
      const float v = cval[j * kTU + row];
      const int i = cidx[j * kTU + row];
      if (v > bv || (v == bv && i < bi)) { best = j; bv = v; bi = i; }
    }
    if (best != r) {
      cval[best * kTU + row] = cval[r * kTU + row];
      cidx[best * kTU + row] = cidx[r * kTU + row];
      cval[r * kTU + row] = bv;
      cidx[r * kTU + row] = bi;
    }
  }
  cnt = kk;
}

template <int D>
__global__ void __launch_bounds__(kThreads)
score_topk_f32_kernel(const float* __restrict__ user_emb, const float* __restrict__ item_emb,
                      const int64_t* __restrict__ user_ids, int n_eval, int m_items,
                      const int64_t* __restrict__ pos_rowptr, const int32_t* __restrict__ pos_sorted,
                      int k, int cap, float mask_value, int32_t* __restrict__ out_idx,
                      float* __restrict__ out_val) {
  constexpr int LD = D + 4;        // padded row: 16B aligned, conflict-free LDS.128
  constexpr int QD = D / 4;        // float4 per row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sA = reinterpret_cast<float*>(smem_raw);   // [kTU][LD]
  float* sB = sA + kTU * LD;                        // [kTI][LD]
  float* sS = sB + kTI * LD;                        // [kTU][kTI+1]
  float* cval = sS + kTU * (kTI + 1);               // [cap][kTU]
  int* cidx = reinterpret_cast<int*>(cval + cap * kTU);

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * kTU;

  // ---- resident user tile ----
  for (int i = tid; i < kTU * QD; i += kThreads) {
    const int r = i / QD, q = i % QD;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < n_eval) v = ld_f4(user_emb + user_ids[row0 + r] * D + 4 * q);
    st_f4(sA + r * LD + 4 * q, v);
  }

  // ---- per-row selection state (threads 0..kTU-1) ----
  float thr = -INFINITY;
  int cnt = 0;
  const int32_t* my_pos = pos_sorted;
  int my_npos = 0;
  const bool selector = tid < kTU && row0 + tid < n_eval;
  if (selector) {
    const int64_t u = user_ids[row0 + tid];
    const int64_t b = pos_rowptr[u];
    my_pos = pos_sorted + b;
    my_npos = (int)(pos_rowptr[u + 1] - b);
  }

  for (int it0 = 0; it0 < m_items; it0 += kTI) {
    // ---- item tile -> shared ----
    for (int i = tid; i < kTI * QD; i += kThreads) {
      const int r = i / QD, q = i % QD;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (it0 + r < m_items) v = ldg_f4_stream(item_emb + (int64_t)(it0 + r) * D + 4 * q);
      st_f4(sB + r * LD + 4 * q, v);
    }
    __syncthreads();  // sB ready; previous tile's selection finished

    // ---- 4 x 8 register tile: rows ty*4+i, cols tx+16*j ----
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int q = 0; q < QD; ++q) {
      float4 a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = ld_f4(sA + (ty * 4 + i) * LD + 4 * q);
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = ld_f4(sB + (tx + 16 * j) * LD + 4 * q);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) sS[(ty * 4 + i) * (kTI + 1) + tx + 16 * j] = acc[i][j];
    __syncthreads();  // scores visible; sB free for the next tile

    // ---- threshold-filtered selection, one thread per row ----
    if (selector) {
      const int ncols = (m_items - it0) < kTI ? (m_items - it0) : kTI;
      const float* srow = sS + tid * (kTI + 1);
      for (int c = 0; c < ncols; ++c) {
        const float s = srow[c];
        if (s > thr) {
          const int item = it0 + c;
          const float v = pos_contains(my_pos, my_npos, item) ? mask_value : s;
          if (v > thr) {
            cval[cnt * kTU + tid] = v;
            cidx[cnt * kTU + tid] = item;
            if (++cnt == cap) {
              reselect(cval, cidx, tid, cnt, k);
              thr = cval[(k - 1) * kTU + tid];
            }
          }
        }
      }
    }
  }

  if (selector) {
    reselect(cval, cidx, tid, cnt, k);
    const int64_t o = (int64_t)(row0 + tid) * k;
    for (int r = 0; r < k; ++r) {
      out_idx[o + r] = r < cnt ? cidx[r * kTU + tid] : -1;
      out_val[o + r] = r < cnt ? cval[r * kTU + tid] : -INFINITY;
    }
  }
}

// Debug aid: dense scores with the SAME fp32 FMA order as the fused kernel.
__global__ void __launch_bounds__(256)
score_dense_f32_kernel(const float* __restrict__ user_emb, const float* __restrict__ item_emb,
                       const int64_t* __restrict__ user_ids, int64_t n_eval, int64_t m_items, int d,
                       float* __restrict__ scores) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_eval * m_items) return;
  const int64_t r = t / m_items, c = t % m_items;
  const float* a = user_emb + user_ids[r] * d;
  const float* b = item_emb + c * d;
  float acc = 0.f;
  for (int q = 0; q < d; ++q) acc = fmaf(a[q], b[q], acc);
  scores[t] = acc;
}

template <int D>
static int launch_f32(const float* user_emb, const float* item_emb, const int64_t* user_ids,
                      int n_eval, int m_items, const int64_t* pos_rowptr,
                      const int32_t* pos_sorted, int k, float mask_value, int32_t* out_idx,
                      float* out_val, cudaStream_t st) {
  int cap = 2 * k;
  cap = ((cap + 31) / 32) * 32;
  if (cap < 64) cap = 64;
  const size_t smem = sizeof(float) * ((size_t)kTU * (D + 4) + (size_t)kTI * (D + 4) +
                                       (size_t)kTU * (kTI + 1) + 2 * (size_t)cap * kTU);
  if (smem > 227 * 1024) {
    set_last_error("k=%d needs %zu bytes of shared memory", k, smem);
    return LGCN_ERR_UNSUPPORTED;
  }
  auto kern = score_topk_f32_kernel<D>;
  LGCN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (n_eval + kTU - 1) / kTU;
  kern<<<grid, kThreads, smem, st>>>(user_emb, item_emb, user_ids, n_eval, m_items, pos_rowptr,
                                     pos_sorted, k, cap, mask_value, out_idx, out_val);
  LGCN_LAUNCH_OK();
  return 0;
}

}  // namespace lgcn

using namespace lgcn;

static int score_topk_impl(const float* user_emb, const float* item_emb, const int64_t* user_ids,
                           int64_t n_eval, int64_t m_items, int d, const int64_t* pos_rowptr,
                           const int32_t* pos_sorted, int k, float mask_value, int precision,
                           int32_t* out_idx, float* out_val, float* dense, void* workspace,
                           size_t workspace_bytes, cudaStream_t st) {
  LGCN_CHECK_ARG(user_emb && item_emb && user_ids && pos_rowptr && pos_sorted && out_idx && out_val,
                 "null pointer argument");
  LGCN_CHECK_ARG(n_eval >= 0 && n_eval < 0x7fffffffLL, "n_eval out of range");
  LGCN_CHECK_ARG(m_items > 0 && m_items < 0x7fffffffLL, "m_items out of range");
  LGCN_CHECK_ARG(k >= 1 && k <= 128 && k <= m_items, "k must be in [1, min(128, m_items)]");
  if (n_eval == 0) return 0;
  if (precision == LGCN_BF16 || precision == LGCN_F16)
    return score_topk_tc(user_emb, item_emb, user_ids, n_eval, m_items, d, pos_rowptr, pos_sorted,
                         k, mask_value, out_idx, out_val, dense, workspace, workspace_bytes,
                         precision == LGCN_F16 ? 1 : 0, st);
  LGCN_CHECK_ARG(precision == LGCN_F32, "precision must be LGCN_F32, LGCN_BF16 or LGCN_F16");
  if (dense != nullptr) {
    const int rc = lgcn_score_dense_f32(user_emb, item_emb, user_ids, n_eval, m_items, d, dense, st);
    if (rc != 0) return rc;
  }
  switch (d) {
    case 32: return launch_f32<32>(user_emb, item_emb, user_ids, (int)n_eval, (int)m_items, pos_rowptr, pos_sorted, k, mask_value, out_idx, out_val, st);
    case 64: return launch_f32<64>(user_emb, item_emb, user_ids, (int)n_eval, (int)m_items, pos_rowptr, pos_sorted, k, mask_value, out_idx, out_val, st);
    case 128: return launch_f32<128>(user_emb, item_emb, user_ids, (int)n_eval, (int)m_items, pos_rowptr, pos_sorted, k, mask_value, out_idx, out_val, st);
    default:
      set_last_error("unsupported embedding width d=%d (supported: 32, 64, 128)", d);
      return LGCN_ERR_UNSUPPORTED;
  }
}

extern "C" int64_t lgcn_score_topk_workspace_bytes(int64_t n_eval, int64_t m_items, int d,
                                                   int precision) {
  if ((precision != LGCN_BF16 && precision != LGCN_F16) || n_eval <= 0 || m_items <= 0) return 0;
  return (int64_t)score_topk_tc_workspace(n_eval, m_items, d);
}

extern "C" int lgcn_score_topk(const float* user_emb, const float* item_emb,
                               const int64_t* user_ids, int64_t n_eval, int64_t m_items, int d,
                               const int64_t* pos_rowptr, const int32_t* pos_sorted, int k,
                               float mask_value, int precision, int32_t* out_idx, float* out_val,
                               void* workspace, int64_t workspace_bytes, lgcn_stream_t stream) {
  return score_topk_impl(user_emb, item_emb, user_ids, n_eval, m_items, d, pos_rowptr, pos_sorted, k,
                         mask_value, precision, out_idx, out_val, nullptr, workspace,
                         (size_t)(workspace_bytes < 0 ? 0 : workspace_bytes), (cudaStream_t)stream);
}

extern "C" int lgcn_score_topk_debug(const float* user_emb, const float* item_emb,
                                     const int64_t* user_ids, int64_t n_eval, int64_t m_items, int d,
                                     const int64_t* pos_rowptr, const int32_t* pos_sorted, int k,
                                     float mask_value, int precision, int32_t* out_idx,
                                     float* out_val, void* workspace, int64_t workspace_bytes,
                                     float* dense_scores, lgcn_stream_t stream) {
  LGCN_CHECK_ARG(dense_scores != nullptr, "dense_scores is null");
  return score_topk_impl(user_emb, item_emb, user_ids, n_eval, m_items, d, pos_rowptr, pos_sorted, k,
                         mask_value, precision, out_idx, out_val, dense_scores, workspace,
                         (size_t)(workspace_bytes < 0 ? 0 : workspace_bytes), (cudaStream_t)stream);
}

extern "C" int lgcn_score_dense_f32(const float* user_emb, const float* item_emb,
                                    const int64_t* user_ids, int64_t n_eval, int64_t m_items, int d,
                                    float* scores, lgcn_stream_t stream) {
  LGCN_CHECK_ARG(user_emb && item_emb && user_ids && scores, "null pointer argument");
  LGCN_CHECK_ARG(n_eval >= 0 && m_items > 0 && d > 0, "bad shape");
  const int64_t total = n_eval * m_items;
  if (total == 0) return 0;
  const int64_t blocks = (total + 255) / 256;
  LGCN_CHECK_ARG(blocks < 0x7fffffffLL, "score block too large for the debug kernel");
  score_dense_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      user_emb, item_emb, user_ids, n_eval, m_items, d, scores);
  LGCN_LAUNCH_OK();
  return 0;
}
