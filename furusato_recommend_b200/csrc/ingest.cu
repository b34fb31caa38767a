// Device-side ingest of the reference's interaction files.
//
// Replaces the host Python loop of Loader.__init__ (reference dataloader.py:93-124 for train,
// :126-150 for test): every line of `train{suffix}.txt` is "uid item item ...\n"; the loop does
// `l = line.strip('\n').split(' ')`, `uid = int(l[0])`, `items = [int(i) for i in l[1:]]` and
// appends `[uid] * len(items)` / `items` to trainUser / trainItem (file order, duplicates kept).
//
// Here the raw bytes of the file sit in HBM and three kernels turn them into the same two int64
// arrays (byte work, HBM-bound: every byte is read twice, 16 output bytes per interaction):
//   1. ingest_count_kernel   per 4 KiB tile: number of tokens (maximal digit runs) and of uid tokens
//                            (the first token of a line)
//   2. tile scan             exclusive scans of both counts (one CTA, as in the sampler's compaction)
//   3. ingest_emit_kernel    every token start parses its digits; a uid token is written to
//                            line_uid[ordinal of its line]; any other token is an item and goes to
//                            out slot (token index - uid tokens before it) — no further scan is
//                            needed — together with its line ordinal
//   4. ingest_users_kernel   user[j] = line_uid[line_of[j]]
// A line is terminated by '\n' (a missing final newline is fine).  Bytes other than digits are
// separators; a '-' or any other non-digit, non-space, non-newline byte sets the error flag (the
// reference's int() would raise ValueError).  Lines without any token are skipped like blank lines
// in our host parser; a line with a uid but no items emits nothing (the reference fails at
// `max(items)` there, dataloader.py:119).
#include "common.cuh"

namespace lgcn {

constexpr int kIngestTile = 4096;
constexpr int kIngestThreads = 256;
constexpr int kBytesPerThread = kIngestTile / kIngestThreads;   // 16 consecutive bytes per thread

__device__ __forceinline__ bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }

// Has the line that contains byte `base` already started a token before `base`?  Walk back to the
// previous newline; separators are single spaces, so the walk stops after a byte or two.
__device__ __forceinline__ bool token_seen_before(const unsigned char* __restrict__ text, int64_t base) {
  for (int64_t q = base - 1; q >= 0; --q) {
    const unsigned char c = text[q];
    if (c == '\n') return false;
    if (is_digit(c)) return true;
  }
  return false;
}

// The thread's 16 consecutive bytes: one 16-byte load (the buffer is 16-byte aligned), '\n' past the end.
__device__ __forceinline__ void load_chunk(const unsigned char* __restrict__ text, int64_t n_bytes, int64_t base,
                                           unsigned char (&buf)[kBytesPerThread]) {
  if (base + kBytesPerThread <= n_bytes) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + base));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < kBytesPerThread; ++i) buf[i] = (unsigned char)((w[i >> 2] >> (8 * (i & 3))) & 0xffu);
  } else {
#pragma unroll
    for (int i = 0; i < kBytesPerThread; ++i) buf[i] = base + i < n_bytes ? text[base + i] : (unsigned char)'\n';
  }
}

// Per tile: (tokens, uid tokens) that START inside [base, base + kIngestTile); a uid token is the
// first token of its line.
__global__ void __launch_bounds__(kIngestThreads)
ingest_count_kernel(const unsigned char* __restrict__ text, int64_t n_bytes, int64_t* __restrict__ tile_tok,
                    int64_t* __restrict__ tile_uid, int32_t* __restrict__ err) {
  const int64_t base = (int64_t)blockIdx.x * kIngestTile + (int64_t)threadIdx.x * kBytesPerThread;
  int tok = 0, uid = 0;
  if (base < n_bytes) {
    unsigned char prev = base > 0 ? text[base - 1] : (unsigned char)'\n';
    bool seen = token_seen_before(text, base);
    unsigned char buf[kBytesPerThread];
    load_chunk(text, n_bytes, base, buf);
#pragma unroll
    for (int i = 0; i < kBytesPerThread; ++i) {
      const int64_t p = base + i;
      if (p >= n_bytes) break;
      const unsigned char c = buf[i];
      const bool dg = is_digit(c);
      if (dg && !is_digit(prev)) {
        ++tok;
        uid += !seen;
        seen = true;
      }
      if (c == '\n') seen = false;
      if (!dg && c != ' ' && c != '\n' && c != '\r' && c != '\t') atomicOr(err, 1);
      prev = c;
    }
  }
  __shared__ int s_tok[kIngestThreads / 32], s_uid[kIngestThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tok += __shfl_xor_sync(0xffffffffu, tok, o);
    uid += __shfl_xor_sync(0xffffffffu, uid, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_tok[threadIdx.x >> 5] = tok;
    s_uid[threadIdx.x >> 5] = uid;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, b = 0;
    for (int w = 0; w < kIngestThreads / 32; ++w) {
      a += s_tok[w];
      b += s_uid[w];
    }
    tile_tok[blockIdx.x] = a;
    tile_uid[blockIdx.x] = b;
  }
}

// exclusive scan of two int64 arrays by ONE CTA of 1024 threads; totals to totals[0..1]
__global__ void __launch_bounds__(1024)
ingest_scan_kernel(int64_t* a, int64_t* b, int64_t n, int64_t* totals) {
  __shared__ int64_t wa[32], wb[32];
  __shared__ int64_t run_a, run_b;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { run_a = 0; run_b = 0; }
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t t = base + threadIdx.x;
    const int64_t ca = t < n ? a[t] : 0, cb = t < n ? b[t] : 0;
    int64_t va = ca, vb = cb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t na = __shfl_up_sync(0xffffffffu, va, o), nb = __shfl_up_sync(0xffffffffu, vb, o);
      if (lane >= o) { va += na; vb += nb; }
    }
    if (lane == 31) { wa[warp] = va; wb[warp] = vb; }
    __syncthreads();
    if (warp == 0) {
      int64_t xa = wa[lane], xb = wb[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t na = __shfl_up_sync(0xffffffffu, xa, o), nb = __shfl_up_sync(0xffffffffu, xb, o);
        if (lane >= o) { xa += na; xb += nb; }
      }
      wa[lane] = xa;
      wb[lane] = xb;
    }
    __syncthreads();
    const int64_t ia = run_a + (warp ? wa[warp - 1] : 0) + va, ib = run_b + (warp ? wb[warp - 1] : 0) + vb;
    if (t < n) { a[t] = ia - ca; b[t] = ib - cb; }
    __syncthreads();
    if (threadIdx.x == 1023) { run_a = ia; run_b = ib; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { totals[0] = run_a; totals[1] = run_b; }
}

__global__ void __launch_bounds__(kIngestThreads)
ingest_emit_kernel(const unsigned char* __restrict__ text, int64_t n_bytes, const int64_t* __restrict__ tile_tok,
                   const int64_t* __restrict__ tile_uid, int64_t* __restrict__ line_uid, int64_t n_lines_cap,
                   int64_t* __restrict__ item, int64_t* __restrict__ line_of, int64_t cap_items,
                   int32_t* __restrict__ err) {
  const int64_t base = (int64_t)blockIdx.x * kIngestTile + (int64_t)threadIdx.x * kBytesPerThread;
  // pass 1: this thread's counts, then an exclusive scan over the CTA's threads
  int tok = 0, uid = 0;
  const bool seen0 = base < n_bytes ? token_seen_before(text, base) : false;
  const unsigned char prev0 = base > 0 && base <= n_bytes ? text[base - 1] : (unsigned char)'\n';
  unsigned char buf[kBytesPerThread];
  load_chunk(text, n_bytes, base < n_bytes ? base : n_bytes, buf);
  {
    unsigned char prev = prev0;
    bool seen = seen0;
#pragma unroll
    for (int i = 0; i < kBytesPerThread; ++i) {
      const int64_t p = base + i;
      const unsigned char c = buf[i];
      if (p < n_bytes && is_digit(c) && !is_digit(prev)) {
        ++tok;
        uid += !seen;
        seen = true;
      }
      if (c == '\n') seen = false;
      prev = c;
    }
  }
  __shared__ int s_tok[kIngestThreads / 32], s_uid[kIngestThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int it = tok, iu = uid;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, it, o), b = __shfl_up_sync(0xffffffffu, iu, o);
    if (lane >= o) { it += a; iu += b; }
  }
  if (lane == 31) { s_tok[warp] = it; s_uid[warp] = iu; }
  __syncthreads();
  int64_t tok_idx = tile_tok[blockIdx.x] + (it - tok);   // tokens that start before this thread's bytes
  int64_t uid_cnt = tile_uid[blockIdx.x] + (iu - uid);   // uid tokens (= non-empty lines) before them
  for (int w = 0; w < warp; ++w) { tok_idx += s_tok[w]; uid_cnt += s_uid[w]; }
  // pass 2: emit.  Line ordinal of a token = (uid tokens up to and including its own line's) - 1;
  // an item token's output slot = its token index minus the uid tokens before it.
  unsigned char prev = prev0;
  bool seen = seen0;
#pragma unroll
  for (int i = 0; i < kBytesPerThread; ++i) {
    const int64_t p = base + i;
    if (p >= n_bytes) break;
    const unsigned char c = buf[i];
    if (is_digit(c) && !is_digit(prev)) {
      int64_t v = 0;
      int nd = 0;
      for (int64_t q = p; q < n_bytes && is_digit(text[q]); ++q, ++nd) v = v * 10 + (text[q] - '0');
      if (nd > 18) atomicOr(err, 2);
      if (!seen) {
        if (uid_cnt < n_lines_cap) line_uid[uid_cnt] = v; else atomicOr(err, 4);
        ++uid_cnt;
        seen = true;
      } else {
        const int64_t o = tok_idx - uid_cnt;
        if (o >= 0 && o < cap_items) {
          item[o] = v;
          line_of[o] = uid_cnt - 1;
        } else {
          atomicOr(err, 4);
        }
      }
      ++tok_idx;
    }
    if (c == '\n') seen = false;
    prev = c;
  }
}

__global__ void __launch_bounds__(256)
ingest_users_kernel(const int64_t* __restrict__ line_uid, const int64_t* __restrict__ line_of, int64_t n,
                    int64_t* __restrict__ user) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    user[i] = line_uid[line_of[i]];
}

}  // namespace lgcn

using namespace lgcn;

extern "C" int64_t lgcn_ingest_tiles(int64_t n_bytes) { return (n_bytes + kIngestTile - 1) / kIngestTile; }

extern "C" int lgcn_ingest_count(const uint8_t* text, int64_t n_bytes, int64_t* tile_tok, int64_t* tile_uid,
                                 int64_t* totals, int32_t* err, lgcn_stream_t stream) {
  LGCN_CHECK_ARG(text && tile_tok && tile_uid && totals && err, "null pointer argument");
  LGCN_CHECK_ARG(n_bytes >= 0, "negative size");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_tiles = (n_bytes + kIngestTile - 1) / kIngestTile;
  LGCN_CHECK_ARG(n_tiles < 0x7fffffffLL, "file too large for one launch");
  if (n_tiles > 0) {
    ingest_count_kernel<<<(unsigned)n_tiles, kIngestThreads, 0, st>>>(text, n_bytes, tile_tok, tile_uid, err);
    LGCN_LAUNCH_OK();
  }
  ingest_scan_kernel<<<1, 1024, 0, st>>>(tile_tok, tile_uid, n_tiles, totals);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_ingest_emit(const uint8_t* text, int64_t n_bytes, const int64_t* tile_tok, const int64_t* tile_uid,
                                int64_t* line_uid, int64_t n_lines_cap, int64_t* user, int64_t* item, int64_t* line_of,
                                int64_t cap_items, int32_t* err, lgcn_stream_t stream) {
  LGCN_CHECK_ARG(text && tile_tok && tile_uid && line_uid && user && item && line_of && err, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_tiles = (n_bytes + kIngestTile - 1) / kIngestTile;
  if (n_tiles == 0) return 0;
  ingest_emit_kernel<<<(unsigned)n_tiles, kIngestThreads, 0, st>>>(text, n_bytes, tile_tok, tile_uid, line_uid,
                                                                  n_lines_cap, item, line_of, cap_items, err);
  LGCN_LAUNCH_OK();
  if (cap_items > 0) {
    int64_t blocks = (cap_items + 255) / 256;
    if (blocks > (int64_t)kSmCount * 16) blocks = (int64_t)kSmCount * 16;
    ingest_users_kernel<<<(unsigned)blocks, 256, 0, st>>>(line_uid, line_of, cap_items, user);
    LGCN_LAUNCH_OK();
  }
  return 0;
}
