// Uniform BPR negative sampler (Philox4x32-10, counter based) + order-preserving
// compaction.  Replaces the Python loop of negative_sample.UniformSample
// (reference negative_sample.py:98-134); decision procedure per sample:
//   user   = randint(n_users)                                  (:107)
//   empty positive list -> sample dropped                      (:116-117)
//   pos    = allPos[user][randint(len)]   FILE ORDER           (:119-120)
//   neg    = first randint(m_items) not in allPos[user]        (:121-126)
// with randint(k) := mulhi32(r, k) on successive Philox words.
#include "common.cuh"

namespace lgcn {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

__device__ __forceinline__ bool contains_sorted(const int32_t* __restrict__ a, int64_t n, int32_t x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int32_t v = __ldg(a + mid);
    if (v < x) lo = mid + 1; else hi = mid;
  }
  return lo < n && __ldg(a + lo) == x;
}

// Weighted positive pick (negative_sample.py:53-56, np.random.choice(len(pos), p=probs[user])):
// numpy draws one uniform u and returns searchsorted(cumsum(p) / sum, u, side='right').  Here
// u = (r >> 8) * 2^-24 from the sample's second Philox word and cdf is the user's inclusive,
// normalised fp32 cumulative table: the first j with cdf[j] > u (clamped to the last entry).
__device__ __forceinline__ int64_t pick_by_cdf(const float* __restrict__ cdf, int64_t n, uint32_t r) {
  const float u = (float)(r >> 8) * 5.9604644775390625e-08f;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(cdf + mid) > u) hi = mid; else lo = mid + 1;
  }
  return lo < n ? lo : n - 1;
}

__global__ void __launch_bounds__(256)
uniform_sample_kernel(const int64_t* __restrict__ pos_rowptr, const int32_t* __restrict__ pos_file,
                      const int32_t* __restrict__ pos_sorted, const float* __restrict__ pos_cdf,
                      uint32_t n_users, uint32_t m_items,
                      int64_t first, int64_t count, int n_neg, uint32_t seed_lo, uint32_t seed_hi,
                      uint32_t epoch, int64_t* __restrict__ triples, uint8_t* __restrict__ valid) {
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= count) return;
  const int64_t i = first + slot;  // global sample index == Philox counter
  uint32_t r[4] = {(uint32_t)(i & 0xffffffffu), (uint32_t)((uint64_t)i >> 32), 0u, epoch};
  philox4x32_10(r, seed_lo, seed_hi);
  const int64_t user = __umulhi(r[0], n_users);
  const int64_t b = pos_rowptr[user], e = pos_rowptr[user + 1];
  const int64_t len = e - b;
  int64_t* o = triples + 3 * slot * n_neg;
  uint8_t* vo = valid + slot * n_neg;
  if (len == 0) {
    for (int t = 0; t < n_neg; ++t) {
      vo[t] = 0;
      o[3 * t] = user; o[3 * t + 1] = -1; o[3 * t + 2] = -1;
    }
    return;
  }
  const int64_t positem = pos_cdf != nullptr ? pos_file[b + pick_by_cdf(pos_cdf + b, len, r[1])]
                                             : pos_file[b + __umulhi(r[1], (uint32_t)len)];
  uint32_t j = 2, blk = 0;
  bool exhausted = false;
  for (int t = 0; t < n_neg; ++t) {   // n_neg > 1: the sampled-softmax layout, flat (u, pos, neg_t) rows
    int32_t negitem = -1;
    // The reference's `while True` (negative_sample.py:121-126) never ends for a user whose positives
    // cover every item; on the host that is a Ctrl-C, on the GPU a wedged device.  After
    // LGCN_MAX_NEG_TRIES rejected candidates in a row the sample (and the rest of its negatives) is
    // dropped like an empty user.  With deg(u) <= m/2 the odds of that are < 2^-256: parity untouched.
    for (int tries = 0; !exhausted; ++tries) {
      if (tries == LGCN_MAX_NEG_TRIES) { exhausted = true; break; }
      if ((j >> 2) != blk) {
        blk = j >> 2;
        r[0] = (uint32_t)(i & 0xffffffffu); r[1] = (uint32_t)((uint64_t)i >> 32); r[2] = blk; r[3] = epoch;
        philox4x32_10(r, seed_lo, seed_hi);
      }
      negitem = (int32_t)__umulhi(r[j & 3u], m_items);
      ++j;
      if (!contains_sorted(pos_sorted + b, len, negitem)) break;
    }
    vo[t] = exhausted ? 0 : 1;
    o[3 * t] = user;
    o[3 * t + 1] = exhausted ? -1 : positem;
    o[3 * t + 2] = exhausted ? -1 : negitem;
  }
}

// ---- order-preserving compaction: count per 1024-tile, scan tiles, scatter ----
constexpr int kTile = 1024;

__global__ void __launch_bounds__(256)
tile_count_kernel(const uint8_t* __restrict__ valid, int64_t count, int64_t* __restrict__ tile_cnt) {
  const int64_t base = (int64_t)blockIdx.x * kTile;
  int c = 0;
  for (int t = threadIdx.x; t < kTile; t += 256) {
    const int64_t i = base + t;
    if (i < count) c += valid[i] ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ int s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s[w];
    tile_cnt[blockIdx.x] = t;
  }
}

// exclusive scan of the tile counts by ONE CTA of 1024 threads (1024 tiles per pass)
__global__ void __launch_bounds__(kTile)
tile_scan_kernel(int64_t* tile_cnt, int64_t n_tiles, int64_t* n_out) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t s_run;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_tiles; base += kTile) {
    const int64_t t = base + threadIdx.x;
    const int64_t c = t < n_tiles ? tile_cnt[t] : 0;
    int64_t v = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int64_t w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int64_t run = s_run;
    const int64_t incl = run + (warp ? warp_tot[warp - 1] : 0) + v;
    if (t < n_tiles) tile_cnt[t] = incl - c;
    __syncthreads();
    if (threadIdx.x == kTile - 1) s_run = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = s_run;
}

__global__ void __launch_bounds__(kTile)
tile_scatter_kernel(const int64_t* __restrict__ triples, const uint8_t* __restrict__ valid,
                    int64_t count, const int64_t* __restrict__ tile_off, int64_t* __restrict__ out) {
  __shared__ int warp_tot[kTile / 32];
  const int64_t i = (int64_t)blockIdx.x * kTile + threadIdx.x;
  const bool ok = i < count && valid[i];
  const unsigned bal = __ballot_sync(0xffffffffu, ok);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_tot[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    int v = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    warp_tot[lane] = v;  // inclusive
  }
  __syncthreads();
  if (!ok) return;
  const int64_t dst = tile_off[blockIdx.x] + (warp ? warp_tot[warp - 1] : 0) +
                      __popc(bal & ((1u << lane) - 1u));
  out[3 * dst + 0] = triples[3 * i + 0];
  out[3 * dst + 1] = triples[3 * i + 1];
  out[3 * dst + 2] = triples[3 * i + 2];
}

}  // namespace lgcn

using namespace lgcn;

extern "C" int lgcn_uniform_sample_weighted(const int64_t* pos_rowptr, const int32_t* pos_file,
                                            const int32_t* pos_sorted, const float* pos_cdf, int64_t n_users,
                                            int64_t m_items, int64_t first, int64_t count, int n_neg,
                                            uint64_t seed, uint32_t epoch, int64_t* triples, uint8_t* valid,
                                            lgcn_stream_t stream);

extern "C" int lgcn_uniform_sample(const int64_t* pos_rowptr, const int32_t* pos_file,
                                   const int32_t* pos_sorted, int64_t n_users, int64_t m_items,
                                   int64_t first, int64_t count, int n_neg, uint64_t seed,
                                   uint32_t epoch, int64_t* triples, uint8_t* valid,
                                   lgcn_stream_t stream) {
  return lgcn_uniform_sample_weighted(pos_rowptr, pos_file, pos_sorted, nullptr, n_users, m_items, first, count,
                                      n_neg, seed, epoch, triples, valid, stream);
}

extern "C" int lgcn_uniform_sample_weighted(const int64_t* pos_rowptr, const int32_t* pos_file,
                                            const int32_t* pos_sorted, const float* pos_cdf, int64_t n_users,
                                            int64_t m_items, int64_t first, int64_t count, int n_neg,
                                            uint64_t seed, uint32_t epoch, int64_t* triples, uint8_t* valid,
                                            lgcn_stream_t stream) {
  LGCN_CHECK_ARG(pos_rowptr && pos_file && pos_sorted && triples && valid, "null pointer argument");
  LGCN_CHECK_ARG(n_users > 0 && n_users < 0xffffffffLL, "n_users out of range");
  LGCN_CHECK_ARG(m_items > 0 && m_items < 0x7fffffffLL, "m_items out of range");
  LGCN_CHECK_ARG(count >= 0 && first >= 0, "negative count/first");
  LGCN_CHECK_ARG(n_neg >= 1 && n_neg <= 4096, "n_neg must be in [1, 4096]");
  if (count == 0) return 0;
  const int64_t blocks = (count + 255) / 256;
  LGCN_CHECK_ARG(blocks < 0x7fffffffLL, "count too large");
  uniform_sample_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      pos_rowptr, pos_file, pos_sorted, pos_cdf, (uint32_t)n_users, (uint32_t)m_items, first, count, n_neg,
      (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), epoch, triples, valid);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_compact_triples(const int64_t* triples, const uint8_t* valid, int64_t count,
                                    int64_t* out, int64_t* n_out, int64_t* scratch,
                                    lgcn_stream_t stream) {
  LGCN_CHECK_ARG(triples && valid && out && n_out && scratch, "null pointer argument");
  LGCN_CHECK_ARG(count >= 0, "negative count");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_tiles = (count + kTile - 1) / kTile;
  if (n_tiles > 0) {
    tile_count_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(valid, count, scratch);
    LGCN_LAUNCH_OK();
  }
  tile_scan_kernel<<<1, kTile, 0, st>>>(scratch, n_tiles, n_out);
  LGCN_LAUNCH_OK();
  if (n_tiles > 0) {
    tile_scatter_kernel<<<(unsigned)n_tiles, kTile, 0, st>>>(triples, valid, count, scratch, out);
    LGCN_LAUNCH_OK();
  }
  return 0;
}
