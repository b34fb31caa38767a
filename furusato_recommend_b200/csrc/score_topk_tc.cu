// Tensor-core full-rank scoring + masked top-k (precision = LGCN_BF16).
//
// Replaces getUsersRating's matmul (reference model/lgcn.py:124), the -(1<<10)
// exclude-list index_put (trainer.py:132-137) and torch.topk (trainer.py:138).
// The [U, m] score matrix only ever exists as 128 x TN fp32 accumulator tiles in
// tensor memory.
//
// Structure (one CTA = 128 users, sweeps every item tile; sm_100a only):
//   pack kernels   fp32 rows -> bf16 in the UMMA canonical K-major / no-swizzle
//                  layout, one contiguous block per tile: [k-chunk(16 B)][row],
//                  i.e. 8x16 B core matrices, SBO = 128 B, LBO = rows*16 B.
//   warp 0         producer: one cp.async.bulk per tile (16/32 KB) -> smem ring,
//                  completion on an mbarrier (expect_tx).
//   warps 1 (, 3)  MMA issuers, one warp per user tile: the whole warp runs the loop and an
//                  elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=TN, K=16)
//                  d/16 times per tile into one of two TMEM accumulator stages; tcgen05.commit
//                  frees the smem stage and publishes the accumulator.
//   warp 2         allocates / frees the 2*TN TMEM columns.
//   warps 4..      epilogue: thread == TMEM lane == user row.  tcgen05.ld 64 columns at a
//                  time; two trees of 3-input maxima against the row's running threshold and
//                  one warp vote reject almost every chunk in ~0.6 instructions per score.
//                  Candidate path (some lane has a score above its threshold): the lane finds
//                  the column of its chunk maximum by descent through the maxima the fast path
//                  already holds and appends it; the chunk's second largest value tells whether
//                  the lane has more candidates (rare: those columns are re-read from TMEM).
//                  The per-row shared-memory buffer keeps the top-k set as a binary heap with the
//                  worst entry at the root (fold = replace-root + sift-down, warp-synchronous);
//                  the set is ordered once, at the end of the sweep.  Train positives are masked
//                  by walking the user's sorted list in step with the sweep, one load ahead.
//                  Layout m2g2 (default, k <= 24, d <= 64): two user tiles per CTA, one epilogue
//                  group of 4 warps per user tile; other layouts split the columns of a tile
//                  over two groups and merge at the end.  The other accumulator stage is being
//                  filled meanwhile.  DESIGN.md 4.4 has the measurements behind each choice.
// Ties: items arrive in ascending id inside a thread, filters are strict, the
// final merge orders by (score desc, id asc) => lowest id wins, as in the fp32 path.
#include <cuda_fp16.h>
#include <stdlib.h>

#include <string>

#include "common.cuh"

namespace lgcn {
namespace tc {

constexpr int kUM = 128;      // users per CTA == UMMA M
constexpr int kFrontWarps = 4;

// -DLGCN_TC_PROF=1 (tuning builds only): CTA 0 accumulates clock64() deltas of the pipeline phases of one epilogue
// warp and one MMA-issuing warp; lgcn_debug_tc_prof() reads them back.  [0..7] epilogue: wait-full, TMEM read,
// max tree + vote, candidate path, hand-back, tiles, candidate entries, total; [8..12] issuer: wait B tile,
// wait accumulator-empty, issue, tiles, total.
#ifdef LGCN_TC_PROF
__device__ long long g_tc_prof[24];   // [16..22]: candidate path split — masks, single take, multi path, compaction, #multi, #compactions
// absolute clock64() stamps of CTA 0 for item tiles [kProfT0, kProfT0 + 8): issuer 0 {accumulator free, MMAs issued},
// epilogue warp 0 {accumulator full, TMEM read done, handed back}
__device__ long long g_tc_stamp[8][5];
constexpr int kProfT0 = 6000;
#define LGCN_STAMP(j, slot, cond) do { if ((cond) && (j) >= kProfT0 && (j) < kProfT0 + 8) g_tc_stamp[(j) - kProfT0][slot] = clock64(); } while (0)
#define LGCN_PROF_T(var) const long long var = clock64()
#define LGCN_PROF_ADD(slot, t0, t1) pf[slot] += (t1) - (t0)
#else
#define LGCN_PROF_T(var)
#define LGCN_PROF_ADD(slot, t0, t1)
#define LGCN_STAMP(j, slot, cond)
#endif

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// accumulator hand-off waits (epilogue <-> MMA issuers): the suspend-time hint is a tuning knob
__device__ __forceinline__ void mbar_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  if (mbar_try_wait_hint(bar, parity, hint_ns)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, hint_ns)) {
    if ((++spins & 0xfffffu) == 0 && clock64() - t0 > 8000000000LL) __trap();
  }
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  // the last operand is a suspend-time hint: a waiting thread sleeps in hardware instead of
  // burning issue slots the epilogue warps of the same SM sub-partition need
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps after a few seconds instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfffu) == 0 && clock64() - t0 > 8000000000LL) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// Cluster variants: one CTA fetches a 1/CL slice of the item tile and the copy engine delivers it to the
// same shared-memory offset of every CTA of the cluster (each destination's mbarrier gets the bytes).
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
          "r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// the same arrive delivered to the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// 64 columns of 16-bit accumulators (one per 32-bit TMEM cell) packed two per register
__device__ __forceinline__ void tc_ld32_pack16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// one column: every lane gets its own row's value (used to re-read a candidate from TMEM)
__device__ __forceinline__ uint32_t tc_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
  return v;
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (sm_100 "version 1").
//   bits [0,14)  start address >> 4        bits [16,30) leading (K) byte offset >> 4
//   bits [32,46) stride (M/N) byte offset >> 4   bits [46,48) version = 1   [61,64) layout = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool acc16) {
  // f16 accumulation is only legal with f16 operands (kind::f16: D=f16 requires A=B=f16)
  return ((acc16 ? 0u : 1u) << 4) | ((acc16 ? 0u : 1u) << 7) | ((acc16 ? 0u : 1u) << 10) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- packing
// dst tile t: [c = 0..D/8)[r = 0..R) 16-byte vectors = rows t*R+r, elements 8c..8c+7 (bf16)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <bool F16>
__global__ void __launch_bounds__(256)
pack_bf16_kernel(const float* __restrict__ src, const int64_t* __restrict__ ids, int64_t n_rows, int D,
                 int R, int64_t n_tiles, uint4* __restrict__ dst) {
  const int cpr = D / 8;
  const int64_t total = n_tiles * R * cpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cpr);
    const int64_t row = i / cpr;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (row < n_rows) {
      const int64_t srow = ids ? ids[row] : row;
      const float4 a = ld_f4(src + srow * D + 8 * c);
      const float4 b = ld_f4(src + srow * D + 8 * c + 4);
      if (F16) {
        o.x = pack_f16x2(a.x, a.y); o.y = pack_f16x2(a.z, a.w);
        o.z = pack_f16x2(b.x, b.y); o.w = pack_f16x2(b.z, b.w);
      } else {
        o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
        o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
      }
    }
    const int64_t t = row / R;
    const int r = (int)(row % R);
    dst[(t * cpr + c) * R + r] = o;
  }
}

// ---------------------------------------------------------------- selection
struct Sel {
  float thr;   // admission threshold: the smallest value of the held top-k set (-inf until k entries exist)
  int cnt;     // entries in the buffer: [0, have) is the set, [have, cnt) were appended since the last fold
  int have;    // size of the set (<= k).  Once have == k the set is a binary heap with its WORST entry in slot 0
               // (smallest value; among equal values the largest id — it arrived last and the final order is
               // (value desc, id asc)); sel_finalize orders it once, at the end of the sweep.
};

// true when (av, ai) ranks below (bv, bi) in the final order
__device__ __forceinline__ bool sel_worse(float av, int ai, float bv, int bi) {
  return av < bv || (av == bv && ai > bi);
}

// Slot i of the heap [0, k) is vacant: sink (v, id) from there to its place.
__device__ __forceinline__ void sel_sift_down(float* cval, int* cidx, int NT, int k, int i, float v, int id) {
  for (;;) {
    int c = 2 * i + 1;
    if (c >= k) break;
    float cv = cval[c * NT];
    int ci = cidx[c * NT];
    if (c + 1 < k) {
      const float dv = cval[(c + 1) * NT];
      const int di = cidx[(c + 1) * NT];
      if (sel_worse(dv, di, cv, ci)) {
        ++c;
        cv = dv;
        ci = di;
      }
    }
    if (!sel_worse(cv, ci, v, id)) break;
    cval[i * NT] = cv;
    cidx[i * NT] = ci;
    i = c;
  }
  cval[i * NT] = v;
  cidx[i * NT] = id;
}

// Fold the appended entries [have, cnt) into the top-k set.  An admitted entry replaces the heap's root (the
// worst entry held) and sinks: ~log2(k) levels of four independent shared-memory loads.  The sorted insertion
// this replaces cost ~14 000 clocks per warp-synchronous call (k dependent shared-memory round trips per
// inserted entry; LGCN_TC_PROF build), a linear scan for the minimum ~10 000.  Appended entries carry larger
// ids than everything already held, so on an equal value the newcomer loses (strict >).
__device__ __noinline__ Sel sel_compact(Sel s, float* cval, int* cidx, int NT, int k) {
  int e = s.have;
  if (s.have < k) {   // fill phase: the appended entries sit right behind the set
    s.have = s.cnt < k ? s.cnt : k;
    e = s.have;
    if (s.have < k) {
      s.cnt = s.have;
      s.thr = -INFINITY;
      return s;
    }
    for (int i = k / 2 - 1; i >= 0; --i) sel_sift_down(cval, cidx, NT, k, i, cval[i * NT], cidx[i * NT]);   // heapify
    s.thr = cval[0];
  }
  for (; e < s.cnt; ++e) {
    const float v = cval[e * NT];
    if (v > s.thr) {
      sel_sift_down(cval, cidx, NT, k, 0, v, cidx[e * NT]);
      s.thr = cval[0];
    }
  }
  s.cnt = k;
  return s;
}

// End of the sweep: fold, then order the set by (value desc, id asc) — once per row.
__device__ __noinline__ Sel sel_finalize(Sel s, float* cval, int* cidx, int NT, int k) {
  s = sel_compact(s, cval, cidx, NT, k);
  for (int e = 1; e < s.have; ++e) {
    const float v = cval[e * NT];
    const int id = cidx[e * NT];
    int p = e;
    while (p > 0) {
      const float u = cval[(p - 1) * NT];
      const int ui = cidx[(p - 1) * NT];
      if (!sel_worse(u, ui, v, id)) break;
      cval[p * NT] = u;
      cidx[p * NT] = ui;
      --p;
    }
    cval[p * NT] = v;
    cidx[p * NT] = id;
  }
  s.cnt = s.have;
  return s;
}

// State handed back by sel_append (f16-accumulator path): the candidate buffer and the walk of the user's sorted
// train positives.  Candidates arrive in ascending item id, so membership is a pointer advance, not a search.
struct SelWalk {
  Sel s;
  int pp, next;
};

// Out of line on purpose: the call sites sit in the hot loop and must stay small
// (the unrolled epilogue is instruction-cache bound otherwise).
__device__ __noinline__ SelWalk sel_append(Sel s, int pp, int next, const int32_t* pos, int npos, float raw,
                                           int item, int m_items, float mask_value, float* cval, int* cidx,
                                           int NT, int cap, int k) {
  SelWalk out;
  while (next < item) {
    ++pp;
    next = pp < npos ? __ldg(pos + pp) : 0x7fffffff;
  }
  out.pp = pp;
  out.next = next;
  const float v = next == item ? mask_value : raw;  // trainer.py:137  rating[...] = -(1<<10)
  if (item < m_items && v > s.thr) {                // item >= m_items: zero padding of the last tile
    if (s.cnt == cap) s = sel_compact(s, cval, cidx, NT, k);  // rare: threshold still -inf
    if (v > s.thr) {
      cval[s.cnt * NT] = v;
      cidx[s.cnt * NT] = item;
      ++s.cnt;
    }
  }
  out.s = s;
  return out;
}

// 3-input max (FMNMX3, PTX ISA 8.6 / sm_100+): the fast-path filter is issue bound, and a tree of
// 3-input maxima needs 18 instructions per 32 scores instead of 31
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// max over 8-element blocks (m4[b] covers r[8b .. 8b+7]) and over the whole chunk
__device__ __forceinline__ float max32(const uint32_t (&r)[32], float (&m4)[4]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const float x = fmax3(__uint_as_float(r[8 * b]), __uint_as_float(r[8 * b + 1]), __uint_as_float(r[8 * b + 2]));
    const float y = fmax3(__uint_as_float(r[8 * b + 3]), __uint_as_float(r[8 * b + 4]), __uint_as_float(r[8 * b + 5]));
    m4[b] = fmax3(x, y, fmaxf(__uint_as_float(r[8 * b + 6]), __uint_as_float(r[8 * b + 7])));
  }
  return fmaxf(fmax3(m4[0], m4[1], m4[2]), m4[3]);
}

// ---- argmax of a 64-column chunk by descent through the maxima the fast path already holds.
// A lane that hit in a chunk almost always has exactly ONE candidate, and that candidate is the chunk maximum
// `cm`; only its column is unknown.  Static code walks the halves whose maximum equals cm (left first, so the
// lowest column wins a tie) — 6 compare+branch steps and ~8 max instructions on the path — and collects the
// maxima of the halves it did NOT enter: their maximum is the second largest value of the chunk, which tells
// whether the lane has more than one candidate (then the general path runs).
template <int I>
__device__ __forceinline__ float colv(const uint32_t (&r)[32], const uint32_t (&r2)[32]) {
  if constexpr (I < 32) return __uint_as_float(r[I]); else return __uint_as_float(r2[I - 32]);
}
template <int LO, int N>
__device__ __forceinline__ float rangemax(const uint32_t (&r)[32], const uint32_t (&r2)[32], const float (&m4a)[4],
                                          const float (&m4b)[4], float cma, float cmb) {
  if constexpr (N == 32) {
    return LO == 0 ? cma : cmb;
  } else if constexpr (N == 16) {
    if constexpr (LO < 32) return fmaxf(m4a[LO / 8], m4a[LO / 8 + 1]); else return fmaxf(m4b[(LO - 32) / 8], m4b[(LO - 32) / 8 + 1]);
  } else if constexpr (N == 8) {
    if constexpr (LO < 32) return m4a[LO / 8]; else return m4b[(LO - 32) / 8];
  } else if constexpr (N == 4) {
    return fmaxf(fmax3(colv<LO>(r, r2), colv<LO + 1>(r, r2), colv<LO + 2>(r, r2)), colv<LO + 3>(r, r2));
  } else if constexpr (N == 2) {
    return fmaxf(colv<LO>(r, r2), colv<LO + 1>(r, r2));
  } else {
    return colv<LO>(r, r2);
  }
}
template <int LO, int N>
__device__ __forceinline__ void descend(const uint32_t (&r)[32], const uint32_t (&r2)[32], const float (&m4a)[4],
                                        const float (&m4b)[4], float cma, float cmb, float cm, float& second,
                                        int& col) {
  if constexpr (N == 1) {
    col = LO;
  } else {
    const float ml = rangemax<LO, N / 2>(r, r2, m4a, m4b, cma, cmb);
    const float mr = rangemax<LO + N / 2, N / 2>(r, r2, m4a, m4b, cma, cmb);
    if (ml == cm) {
      second = fmaxf(second, mr);
      descend<LO, N / 2>(r, r2, m4a, m4b, cma, cmb, cm, second, col);
    } else {
      second = fmaxf(second, ml);
      descend<LO + N / 2, N / 2>(r, r2, m4a, m4b, cma, cmb, cm, second, col);
    }
  }
}

__device__ __forceinline__ __half2 as_h2(uint32_t x) { return *reinterpret_cast<__half2*>(&x); }
__device__ __forceinline__ float h2_lo(uint32_t x) { return __half2float(__ushort_as_half((unsigned short)(x & 0xffffu))); }
__device__ __forceinline__ float h2_hi(uint32_t x) { return __half2float(__ushort_as_half((unsigned short)(x >> 16))); }

// packed-f16 variant: register i holds columns 2i (low) and 2i+1 (high); m4[b] covers registers 8b..8b+7
__device__ __forceinline__ float max32_h(const uint32_t (&r)[32], float (&m4)[4]) {
  __half2 m[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = __hmax2(as_h2(r[2 * i]), as_h2(r[2 * i + 1]));
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = __hmax2(m[2 * i], m[2 * i + 1]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 q = __hmax2(m[2 * i], m[2 * i + 1]);
    m4[i] = fmaxf(__low2float(q), __high2float(q));
  }
  return fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
}

struct Params {
  const uint4* a_packed;   // user tiles, [n_utiles][D/8][128]
  const uint4* b_packed;   // item tiles, [n_itiles][D/8][TN]
  const int64_t* user_ids;
  int n_eval, m_items;
  const int64_t* pos_rowptr;
  const int32_t* pos_sorted;
  int k, cap, stages;
  float mask_value;
  int32_t* out_idx;
  float* out_val;
  float* dense;            // optional [n_eval, m_items] dump of the accumulators (tests)
  int trig;                // > 0: compaction trigger override (tuning)
  int debug_mode;          // 0 = normal; pipeline experiments: 1 = epilogue skips the TMEM reads,
                           // 2 = TMEM reads only, 3 = reads + max tree, no candidate passes
  int acc16;               // 1: f16 accumulators, read back two per register (tcgen05.ld pack::16b)
  int split;               // bulk copies per item tile (tuning: LGCN_TC_SPLIT)
  uint32_t wait_hint;      // suspend-time hint (ns) of the accumulator hand-off waits (tuning: LGCN_TC_WAIT_NS)
};

// MT user tiles (128 rows each) share every item tile: the B operand stream is what bounds the MMA
// pipeline at d = 64 (L2 -> SM bulk-copy bandwidth, ~43 B/clk/SM chip-wide: a 32 KB tile per 512
// MMA clocks does not fit), so MT = 2 halves the bytes per flop.  GROUPS epilogue groups of 4 warps:
// group g works on user tile g / CG, column slice g % CG of every item tile (CG = GROUPS / MT).
// NST accumulator stages in TMEM: an epilogue warp that runs into candidates (the rare slow path)
// only holds back ITS stage; with 4 stages the other warps and the MMA issuer run ahead and the
// variance averages out instead of costing every tile the slowest warp's time.
// RS ("register-staged", layout "m2rl", opt-in): the epilogue copies the whole 128-column accumulator of its
// user tile into registers and hands the TMEM stage back BEFORE any selection work; candidates are handled
// lane-locally from the registers.  Parity-tested; slower than m2g2 at d = 64 (the next MMA then overlaps the
// other stage's read-out: TMEM port contention).
// CL > 1 ("m2c2" / "m2c4"): CL CTAs form a thread-block cluster that SHARES the item-tile stream — every
// CTA keeps its own users, accumulators and epilogue, but each item tile is fetched from L2 once per
// cluster (CTA r loads slice r and multicasts it), and a ring stage is recycled when the MMAs of all CL
// CTAs that read it have retired (multicast tcgen05.commit onto every CTA's `empty` barrier).  ncu of the
// round-1 kernel shows the MMA issuer sleeping on the `full` barrier on every tile: the tile supply, not
// the epilogue, bounds the pipeline (148 CTAs pulling the same 16 KB from L2 at the same time).
template <int D, int TN, int GROUPS, bool DUMP, bool ACC16, int MT, int NST, bool RS = false, int CL = 1>
__global__ void __launch_bounds__((kFrontWarps + 4 * GROUPS) * 32, 1)
score_topk_tc_kernel(const Params p) {
  static_assert(!RS || (!ACC16 && GROUPS == MT && TN == 128), "register staging: fp32, one group per user tile, TN 128");
  constexpr int NT = GROUPS * 128;              // epilogue threads
  constexpr int CG = GROUPS / MT;               // column groups per user tile
  constexpr int kTmemCols = NST * MT * TN;
  static_assert(GROUPS % MT == 0 && kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols >= 32,
                "TMEM holds NST stages x MT accumulators of TN columns (power of two, <= 512)");
  constexpr uint32_t kABytes = kUM * D * 2;
  constexpr uint32_t kBBytes = TN * D * 2;
  constexpr int kKSteps = D / 16;
  constexpr uint32_t kIdesc = make_idesc(kUM, TN, ACC16);
  constexpr int kMaxStages = 8;

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;
  unsigned char* sB = sA + MT * kABytes;
  float* cval = reinterpret_cast<float*>(sB + (size_t)p.stages * kBBytes);
  int* cidx = reinterpret_cast<int*>(cval + (size_t)p.cap * NT);
  int* ccnt = cidx + (size_t)p.cap * NT;                      // [NT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ccnt + NT);    // 8-byte aligned: all sizes are multiples of 8
  // bars: full[kMaxStages], empty[kMaxStages], tfull[kMaxStages], tempty[kMaxStages], afull
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * kMaxStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kMaxStages);
  const uint32_t bar_tfull = smem_u32(bars + 2 * kMaxStages), bar_tempty = smem_u32(bars + 3 * kMaxStages);
  const uint32_t bar_afull = smem_u32(bars + 4 * kMaxStages);
  const int n_tiles = (p.m_items + TN - 1) / TN;
  const int S = p.stages;
  constexpr uint16_t kClusterMask = (uint16_t)((1u << CL) - 1u);
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, CL * MT);   // one commit per issuing warp (user tile) per CTA that reads the stage
    }
    // one full/empty pair per (accumulator stage, user tile): the groups of a user tile hand their
    // accumulator back without waiting for the other user tile's warps
    for (int a = 0; a < NST * MT; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4 * CG);   // one elected lane per epilogue warp arrives
    }
    mbar_init(bar_afull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // every CTA's barriers exist before a peer signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ producer
    if (lane == 0) {
      mbar_expect_tx(bar_afull, MT * kABytes);   // the CTA's MT user tiles are contiguous in a_packed
      bulk_g2s(smem_u32(sA), reinterpret_cast<const unsigned char*>(p.a_packed) + (size_t)blockIdx.x * MT * kABytes,
               MT * kABytes, bar_afull);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % S;
        if (j >= S) mbar_wait(bar_empty + 8 * s, ((j / S) - 1) & 1);
        mbar_expect_tx(bar_full + 8 * s, kBBytes);   // the whole tile lands here, from CL senders
        if constexpr (CL > 1) {
          constexpr uint32_t kSlice = kBBytes / CL;
          bulk_g2s_mc(smem_u32(sB + (size_t)s * kBBytes) + cta_rank * kSlice,
                      reinterpret_cast<const unsigned char*>(p.b_packed) + (size_t)j * kBBytes + cta_rank * kSlice,
                      kSlice, bar_full + 8 * s, kClusterMask);
        } else {
          const uint32_t piece = kBBytes / (uint32_t)p.split;
          for (int c = 0; c < p.split; ++c)
            bulk_g2s(smem_u32(sB + (size_t)s * kBBytes) + c * piece,
                     reinterpret_cast<const unsigned char*>(p.b_packed) + (size_t)j * kBBytes + c * piece, piece,
                     bar_full + 8 * s);
        }
      }
      // tail: do not exit while a tcgen05.commit may still arrive on an smem barrier
      for (int j = n_tiles > S ? n_tiles - S : 0; j < n_tiles; ++j)
        mbar_wait(bar_empty + 8 * (j % S), (j / S) & 1);
    }
  } else if (warp == 1 || (MT == 2 && warp == 3)) {
    // ------------------------------------------------ MMA issuers: one WARP per user tile
    // ncu of the round-1 kernel (profiles/r02_tc_topk_before_summary.txt): the epilogue warps sleep on
    // the accumulator barrier on every tile and the single issuing THREAD is busy the whole time — ~250
    // dependent instructions per item tile (runtime j % stages divisions, R2UR + ELECT + BRA.U.ANY around
    // every UTCHMMA because the thread ran under `if (lane == 0)`), i.e. ~1100 clocks of issue latency
    // for 512 clocks of tensor work.  Now: the whole warp runs the loop (uniform control flow, operands
    // stay in uniform registers), stage / phase counters are carried instead of divided, descriptors
    // are hoisted, and the two user tiles have their own issuing warp.
    const int mt = warp == 1 ? 0 : 1;
    mbar_wait(bar_afull, 0);
    uint64_t adesc[kKSteps];
#pragma unroll
    for (int kk = 0; kk < kKSteps; ++kk)   // one K=16 step = two 16-byte k-chunks = 2*LBO bytes further along
      adesc[kk] = make_desc(smem_u32(sA + (size_t)mt * kABytes), kUM * 16, 128) + (uint64_t)((kk * 2 * kUM * 16) >> 4);
    const uint64_t bdesc_base = make_desc(smem_u32(sB), TN * 16, 128);
    int s = 0, a = 0;
    uint32_t full_parity = 0, acc_round = 0;
#ifdef LGCN_TC_PROF
    long long pf[4] = {0, 0, 0, 0};
    const long long pf_begin = clock64();
#endif
    for (int j = 0; j < n_tiles; ++j) {
      LGCN_PROF_T(q0);
      mbar_wait(bar_full + 8 * s, full_parity);
      LGCN_PROF_T(q1);
      if (acc_round > 0) mbar_wait_hint(bar_tempty + 8 * (a * MT + mt), (acc_round - 1) & 1, p.wait_hint);
      LGCN_PROF_T(q2);
      LGCN_PROF_ADD(0, q0, q1);
      LGCN_PROF_ADD(1, q1, q2);
      LGCN_STAMP(j, 0, blockIdx.x == 0 && mt == 0 && lane == 0);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bdesc0 = bdesc_base + (uint64_t)(((uint32_t)s * kBBytes) >> 4);
        const uint32_t tacc = tmem_base + (uint32_t)((a * MT + mt) * TN);
#pragma unroll
        for (int kk = 0; kk < kKSteps; ++kk)
          tc_mma_bf16(tacc, adesc[kk], bdesc0 + (uint64_t)((kk * 2 * TN * 16) >> 4), kIdesc, kk > 0 ? 1u : 0u);
        tc_commit(bar_tfull + 8 * (a * MT + mt));   // this user tile's accumulator is ready for its groups
        // smem stage reusable once the MMAs of BOTH issuers retire — in every CTA that received the tile
        if constexpr (CL > 1) tc_commit_mc(bar_empty + 8 * s, kClusterMask); else tc_commit(bar_empty + 8 * s);
      }
      __syncwarp();
      LGCN_PROF_T(q3);
      LGCN_PROF_ADD(2, q2, q3);
      LGCN_STAMP(j, 1, blockIdx.x == 0 && mt == 0 && lane == 0);
      if (++s == S) { s = 0; full_parity ^= 1u; }
      if (++a == NST) { a = 0; ++acc_round; }
    }
#ifdef LGCN_TC_PROF
    if (blockIdx.x == 0 && mt == 0 && lane == 0) {
      g_tc_prof[8] = pf[0]; g_tc_prof[9] = pf[1]; g_tc_prof[10] = pf[2]; g_tc_prof[11] = n_tiles;
      g_tc_prof[12] = clock64() - pf_begin;
    }
#endif
  } else if (warp >= kFrontWarps) {
    // ------------------------------------------------ epilogue: thread == user row
    const int ew = warp - kFrontWarps;
    const int grp = ew >> 2;
    const int q = warp & 3;                       // TMEM sub-partition this warp may read
    const int row = 32 * q + lane;                // row inside the user tile == TMEM lane
    const int t = grp * 128 + row;                // slot in the candidate arrays
    const int mt = grp / CG, cg = grp % CG;       // user tile and column slice of this group
    const int64_t grow = ((int64_t)blockIdx.x * MT + mt) * kUM + row;
    const bool live = grow < p.n_eval;
    const int32_t* my_pos = p.pos_sorted;
    int my_npos = 0;
    if (live) {
      const int64_t u = p.user_ids[grow];
      const int64_t b = p.pos_rowptr[u];
      my_pos = p.pos_sorted + b;
      my_npos = (int)(p.pos_rowptr[u + 1] - b);
    }
    float* mv = cval + t;
    int* mi = cidx + t;
    Sel sel;
    const int dbg = p.debug_mode;   // read once: the parameter load does not belong in the per-chunk loop
    const int n_chunks_dbg = dbg == 1 ? 0 : 1;
    sel.thr = (live && dbg < 3) ? -INFINITY : INFINITY;   // debug 3: max tree only, no candidate ever passes
    sel.cnt = 0;
    sel.have = 0;
    // warp-synchronous compaction once any lane holds k + 6 entries (at most cap - 4): folding early
    // keeps the thresholds tight, and every candidate that is not admitted saves a slow-path trip
    // for the whole warp (trigger 22 / 26 / 30 / 36 / 44 at k = 20: 626 / 663 / 657 / 645 / 632 TFLOP/s)
    const int trig = min(p.cap - 4, p.trig > 0 ? p.trig : p.k + 6);
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;
    // walk of the user's sorted train positives, in step with the item sweep
    int pp = 0;
    int next_pos = my_npos > 0 ? __ldg(my_pos) : 0x7fffffff;
    // the positive after next_pos, loaded one step ahead: an advance of the walk is a register move, the
    // global load it issues is only needed by the advance after it (its ~500 clocks sat on every candidate
    // that stepped over a positive)
    int ahead_pos = my_npos > 1 ? __ldg(my_pos + 1) : 0x7fffffff;
    // the column groups of a user tile split every item tile; group cg owns chunks [cg*CPG, (cg+1)*CPG).
    // A chunk is 64 columns: two tcgen05.ld.x32 in flight (fp32), or one packed-f16 load (ACC16).
    // Two independent max trees per trip halve the exposed latency chain (TMEM load -> tree ->
    // vote -> branch) per score; with 2 epilogue warps per SM sub-partition that chain, not the
    // TMEM bandwidth (tools/tmem_read_bw.cu), is what bounds the fast path.
    constexpr int COLS = 64;
    constexpr int CPG = TN / COLS / CG;
    static_assert(CPG >= 1 && CPG * COLS * CG == TN, "every epilogue group needs whole chunks of the tile");
    const int c0 = cg * CPG;

    if constexpr (RS) {
      // Register-staged epilogue ("m2rl"): the warp copies its 32 rows x 128 columns of the accumulator into
      // registers and hands the TMEM stage back BEFORE any selection work, so a warp that runs into
      // candidates no longer holds back its user tile's next MMA (with two TMEM stages the hand-off is
      // otherwise coupled to the slowest of the four warps on every tile).  Candidates are handled
      // lane-locally from the registers: a 16-bit mask of the 8-column blocks this lane hit, a dense
      // switch that moves the hit block into 8 fixed registers, a predicated scan of those — no warp
      // collectives and no TMEM re-reads on the candidate path.
#ifdef LGCN_TC_PROF
      long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const long long pf_begin = clock64();
#endif
      for (int j = 0; j < n_tiles; ++j) {
        const int a = j % NST;
        LGCN_PROF_T(q0);
        mbar_wait_hint(bar_tfull + 8 * (a * MT + mt), (j / NST) & 1, p.wait_hint);
        LGCN_PROF_T(q1);
        LGCN_PROF_ADD(0, q0, q1);
        LGCN_STAMP(j, 2, blockIdx.x == 0 && ew == 0 && lane == 0);
        tc_fence_after();
        const uint32_t tbase = tmem_base + lane_base + (uint32_t)((a * MT + mt) * TN);
        uint32_t r[4][32];
        __syncwarp();
        if (dbg != 1) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) tc_ld32(tbase + (uint32_t)(32 * q4), r[q4]);
          tc_wait_ld();
        }
        tc_fence_before();
        __syncwarp();
        LGCN_STAMP(j, 3, blockIdx.x == 0 && ew == 0 && lane == 0);
        if (lane == 0) mbar_arrive(bar_tempty + 8 * (a * MT + mt));   // the accumulator lives in registers now
        LGCN_STAMP(j, 4, blockIdx.x == 0 && ew == 0 && lane == 0);
        LGCN_PROF_T(q2);
        LGCN_PROF_ADD(1, q1, q2);
        if (dbg == 1) continue;
        if (dbg == 2) {   // pipeline experiment: TMEM reads only, no selection work
          if ((r[0][0] ^ r[1][31] ^ r[2][0] ^ r[3][31]) == 0x7fc12345u) sel.cnt = 1;
          continue;
        }
        const int item_tile0 = j * TN;
        if (DUMP) {
          if (live) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (item_tile0 + 32 * q4 + i < p.m_items)
                  p.dense[grow * p.m_items + item_tile0 + 32 * q4 + i] = __uint_as_float(r[q4][i]);
          }
        }
        float m4[4][4], cm[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) cm[q4] = max32(r[q4], m4[q4]);
        const float cmax = fmaxf(fmax3(cm[0], cm[1], cm[2]), cm[3]);
        const bool slow_entry = __any_sync(0xffffffffu, cmax > sel.thr);
        LGCN_PROF_T(q3);
        LGCN_PROF_ADD(2, q2, q3);
        if (slow_entry) {
          if (cmax > sel.thr) {
            uint32_t bm = 0;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
              for (int b = 0; b < 4; ++b) bm |= (m4[q4][b] > sel.thr) ? (1u << (4 * q4 + b)) : 0u;
            while (bm) {   // hit blocks in ascending column order
              const int blk = __ffs(bm) - 1;
              bm &= bm - 1;
              uint32_t v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = 0xff800000u;   // -inf
              switch (blk) {
#define LGCN_BLK(n)                                                             \
  case n:                                                                       \
    _Pragma("unroll") for (int i = 0; i < 8; ++i) v[i] = r[(n) >> 2][8 * ((n) & 3) + i]; \
    break;
                LGCN_BLK(0) LGCN_BLK(1) LGCN_BLK(2) LGCN_BLK(3) LGCN_BLK(4) LGCN_BLK(5) LGCN_BLK(6) LGCN_BLK(7)
                LGCN_BLK(8) LGCN_BLK(9) LGCN_BLK(10) LGCN_BLK(11) LGCN_BLK(12) LGCN_BLK(13) LGCN_BLK(14) LGCN_BLK(15)
#undef LGCN_BLK
              }
              const int ib = item_tile0 + 8 * blk;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float raw = __uint_as_float(v[i]);
                if (raw > sel.thr) {
                  const int item = ib + i;
                  while (next_pos < item) {   // walk of the sorted train positives (ascending sweep)
                    ++pp;
                    next_pos = ahead_pos;
                    ahead_pos = pp + 1 < my_npos ? __ldg(my_pos + pp + 1) : 0x7fffffff;
                  }
                  const float vv = next_pos == item ? p.mask_value : raw;  // trainer.py:137
                  if (item < p.m_items && vv > sel.thr) {   // item >= m_items: zero padding of the last tile
                    if (sel.cnt == p.cap) sel = sel_compact(sel, mv, mi, NT, p.k);  // rare: threshold still -inf
                    if (vv > sel.thr) {
                      mv[sel.cnt * NT] = vv;
                      mi[sel.cnt * NT] = item;
                      ++sel.cnt;
                    }
                  }
                }
              }
            }
          }
          __syncwarp();
          if (__any_sync(0xffffffffu, sel.cnt >= trig)) sel = sel_compact(sel, mv, mi, NT, p.k);
          LGCN_PROF_T(q4);
          LGCN_PROF_ADD(3, q3, q4);
#ifdef LGCN_TC_PROF
          pf[6] += 1;
#endif
        }
      }
#ifdef LGCN_TC_PROF
      if (blockIdx.x == 0 && ew == 0 && lane == 0) {
        for (int i = 0; i < 5; ++i) g_tc_prof[i] = pf[i];
        g_tc_prof[5] = n_tiles; g_tc_prof[6] = pf[6]; g_tc_prof[7] = clock64() - pf_begin;
      }
#endif
    } else {
#ifdef LGCN_TC_PROF
    long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long ps[6] = {0, 0, 0, 0, 0, 0};
    const long long pf_begin = clock64();
#endif
    // Early hand-back.  With two TMEM stages per user tile the time a stage is HELD (accumulator full -> all four
    // warps done) bounds the whole pipeline (clock64 timeline, profiles/r02_tc_prof_timeline.log: issue 300-540,
    // commit -> epilogue 230-500, hold ~800, hand-back -> issuer 140 clocks per stage cycle).  Once every row of
    // the warp holds k entries (`warm`, warp-uniform) the stage goes back right after the LAST chunk's
    // tcgen05.ld completed: that chunk's max tree, vote and candidate handling work on registers only (the
    // rare several-candidates-in-one-lane case then picks its values with a per-lane switch instead of the
    // TMEM re-read), so half of all candidate entries leave the critical path.
    // Measured (profiles/r02_tc_experiments.log, job r02ac): +5 % at d = 128 (965 -> 1011 TFLOP/s), but -4 % at
    // d = 64 — there the next MMA then overlaps the other stage's read-out and the tcgen05.ld wait grows from
    // 78 to 189 clocks per tile (TMEM port contention between the accumulating MMA and the loads) — so it is
    // compiled in for d >= 128 only.
    constexpr bool kEarlyHandBack = !ACC16 && D >= 128;
    bool warm = dbg == 3;
    for (int j = 0; j < n_tiles; ++j) {
      const int a = j % NST;
      bool released = false;
      LGCN_PROF_T(q0);
      mbar_wait_hint(bar_tfull + 8 * (a * MT + mt), (j / NST) & 1, p.wait_hint);
      LGCN_PROF_T(q1);
      LGCN_PROF_ADD(0, q0, q1);
      LGCN_STAMP(j, 2, blockIdx.x == 0 && ew == 0 && lane == 0);
      tc_fence_after();
      const int item_tile0 = j * TN + c0 * COLS;
      const uint32_t tbase = tmem_base + lane_base + (uint32_t)((a * MT + mt) * TN + c0 * COLS);
#pragma unroll 1
      for (int cc = 0; cc < CPG * n_chunks_dbg; ++cc) {
        uint32_t r[32], r2[32];
        __syncwarp();
        LGCN_PROF_T(c0t);
        if (ACC16) {
          tc_ld32_pack16(tbase + (uint32_t)(cc * COLS), r);
        } else {
          tc_ld32(tbase + (uint32_t)(cc * COLS), r);
          tc_ld32(tbase + (uint32_t)(cc * COLS + 32), r2);
        }
        tc_wait_ld();
        if (kEarlyHandBack && warm && cc == CPG - 1) {   // warp-uniform
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * (a * MT + mt));
          released = true;
        }
        LGCN_PROF_T(c1t);
        LGCN_PROF_ADD(1, c0t, c1t);
        LGCN_STAMP(j, 3, blockIdx.x == 0 && ew == 0 && lane == 0);
#ifdef LGCN_TC_PROF   // tuning builds only: the test does not belong in the per-chunk loop of the product
        if (dbg == 2) {   // pipeline experiment: TMEM reads only, no selection work
          if ((r[0] ^ r[31] ^ (ACC16 ? 0u : r2[0] ^ r2[31])) == 0x7fc12345u) sel.cnt = 1;
          continue;
        }
#endif
        const int item0 = item_tile0 + cc * COLS;
        if (DUMP) {
          if (live) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (ACC16) {
                if (item0 + 2 * i < p.m_items) p.dense[grow * p.m_items + item0 + 2 * i] = h2_lo(r[i]);
                if (item0 + 2 * i + 1 < p.m_items) p.dense[grow * p.m_items + item0 + 2 * i + 1] = h2_hi(r[i]);
              } else {
                if (item0 + i < p.m_items) p.dense[grow * p.m_items + item0 + i] = __uint_as_float(r[i]);
                if (item0 + 32 + i < p.m_items) p.dense[grow * p.m_items + item0 + 32 + i] = __uint_as_float(r2[i]);
              }
            }
          }
        }
        if (!ACC16) {
          // Fast path: max trees of the two 32-column halves against the row's threshold, one vote.
          // Slow path (some lane of the warp has a candidate), per half: build the per-lane hit mask —
          // after this the registers are dead, so nothing has to be kept alive across the candidate
          // handling — then walk the columns any lane hit in ascending order, RE-READING that column
          // from TMEM (1-register tcgen05.ld; the accumulator stage is still ours) instead of
          // indexing live registers.
          float m4a[4], m4b[4];
          const float cma = max32(r, m4a);
          const float cmb = max32(r2, m4b);
          const float cm = fmaxf(cma, cmb);
          const bool slow_entry = __any_sync(0xffffffffu, cm > sel.thr);
          LGCN_PROF_T(c2t);
          LGCN_PROF_ADD(2, c1t, c2t);
          if (slow_entry) {
            // one candidate of this lane: positives walk, mask, append (ascending item ids)
            auto take = [&](int item, uint32_t raw) {
              while (next_pos < item) {   // walk of the sorted train positives (ascending sweep)
                ++pp;
                next_pos = ahead_pos;
                ahead_pos = pp + 1 < my_npos ? __ldg(my_pos + pp + 1) : 0x7fffffff;
              }
              const float v = next_pos == item ? p.mask_value : __uint_as_float(raw);  // trainer.py:137
              if (item < p.m_items && v > sel.thr) {   // item >= m_items: zero padding of the last tile
                if (sel.cnt == p.cap) sel = sel_compact(sel, mv, mi, NT, p.k);  // rare: threshold still -inf
                if (v > sel.thr) {
                  mv[sel.cnt * NT] = v;
                  mi[sel.cnt * NT] = item;
                  ++sel.cnt;
                }
              }
            };
            // Every slow-path instruction of ONE warp is on the critical path of its user tile (two TMEM stages:
            // the four warps of a tile hand the accumulator back together) and runs at 6-15 clocks per
            // instruction (cold, branchy, one warp) — measured 1600 clocks per entry before this version, 500 of
            // them in per-lane hit masks.  The common case is lane-local and short: a lane with a candidate finds
            // the column of its chunk maximum by descent (above); if the second largest value of the chunk is
            // below the threshold that maximum is the lane's only candidate and is taken as is.
            const bool mine = cm > sel.thr;
            int col = 0;
            float second = -INFINITY;
            if (mine) descend<0, 64>(r, r2, m4a, m4b, cma, cmb, cm, second, col);
            const bool multi = mine && second > sel.thr;
            if (mine && !multi) {
              // the lane's only candidate.  No capacity check: entries end with a fold as soon as any lane holds
              // `trig` (<= cap - 4) entries and this path appends one, so the buffer cannot be full here.
              const int item = item0 + col;
              while (next_pos < item) {   // walk of the sorted train positives (ascending sweep)
                ++pp;
                next_pos = ahead_pos;
                ahead_pos = pp + 1 < my_npos ? __ldg(my_pos + pp + 1) : 0x7fffffff;
              }
              const float v = next_pos == item ? p.mask_value : cm;  // trainer.py:137
              if (item < p.m_items && v > sel.thr) {   // item >= m_items: zero padding of the last tile
                mv[sel.cnt * NT] = v;
                mi[sel.cnt * NT] = item;
                ++sel.cnt;
              }
            }
#ifdef LGCN_TC_PROF
            __syncwarp();
            const long long s1 = clock64();
            ps[0] += s1 - c2t;
            const long long s2 = s1;
#endif
            // several hits in one lane (the first tiles of a sweep, then rare): walk the columns those lanes hit
            // in ascending order, RE-READING each column from TMEM (1-register tcgen05.ld with a warp-uniform
            // column; the accumulator stage is still ours) instead of indexing live registers
            if (__any_sync(0xffffffffu, multi || sel.cnt >= trig)) {   // one vote guards both rare paths
            if (__any_sync(0xffffffffu, multi)) {
              // per-lane hit masks of the lanes with several candidates, built only for the 8-column blocks they hit
              uint32_t hma = 0, hmb = 0;
              if (multi) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                  if (m4a[b] > sel.thr) {
#pragma unroll
                    for (int i = 8 * b; i < 8 * b + 8; ++i) hma |= (__uint_as_float(r[i]) > sel.thr) ? (1u << i) : 0u;
                  }
                  if (m4b[b] > sel.thr) {
#pragma unroll
                    for (int i = 8 * b; i < 8 * b + 8; ++i) hmb |= (__uint_as_float(r2[i]) > sel.thr) ? (1u << i) : 0u;
                  }
                }
              }
              auto multi_half = [&](uint32_t hm, int item_base, uint32_t tcol) {
                uint32_t any = __reduce_or_sync(0xffffffffu, hm);
                while (any) {   // two columns per round: both re-reads are in flight before the one wait
                  const int ca = __ffs(any) - 1;
                  any &= any - 1;
                  const bool two = any != 0;
                  const int cb = two ? __ffs(any) - 1 : ca;
                  any &= any - 1;   // no-op when any == 0
                  __syncwarp();     // the candidate handling below diverges; tcgen05.ld is warp-collective
                  const uint32_t ra = tc_ld1(tcol + (uint32_t)ca);
                  const uint32_t rb = tc_ld1(tcol + (uint32_t)cb);
                  tc_wait_ld();
                  if ((hm >> ca) & 1u) take(item_base + ca, ra);
                  if (two && ((hm >> cb) & 1u)) take(item_base + cb, rb);
                }
              };
              if (!released) {
                multi_half(hma, item0, tbase + (uint32_t)(cc * COLS));
                multi_half(hmb, item0 + 32, tbase + (uint32_t)(cc * COLS + 32));
              } else if (multi) {
                // the stage is gone: per-lane walk of the hit columns, values picked from the registers by a dense switch
                auto multi_regs = [&](uint32_t hm, const uint32_t(&x)[32], int item_base) {
                  while (hm) {
                    const int c = __ffs(hm) - 1;
                    hm &= hm - 1;
                    uint32_t raw = 0;
                    switch (c) {
#define LGCN_PICK(i) case i: raw = x[i]; break;
                      LGCN_PICK(0) LGCN_PICK(1) LGCN_PICK(2) LGCN_PICK(3) LGCN_PICK(4) LGCN_PICK(5) LGCN_PICK(6) LGCN_PICK(7)
                      LGCN_PICK(8) LGCN_PICK(9) LGCN_PICK(10) LGCN_PICK(11) LGCN_PICK(12) LGCN_PICK(13) LGCN_PICK(14) LGCN_PICK(15)
                      LGCN_PICK(16) LGCN_PICK(17) LGCN_PICK(18) LGCN_PICK(19) LGCN_PICK(20) LGCN_PICK(21) LGCN_PICK(22) LGCN_PICK(23)
                      LGCN_PICK(24) LGCN_PICK(25) LGCN_PICK(26) LGCN_PICK(27) LGCN_PICK(28) LGCN_PICK(29) LGCN_PICK(30) LGCN_PICK(31)
#undef LGCN_PICK
                    }
                    take(item_base + c, raw);
                  }
                };
                multi_regs(hma, r, item0);
                multi_regs(hmb, r2, item0 + 32);
              }
#ifdef LGCN_TC_PROF
              ps[4] += 1;
#endif
            }
#ifdef LGCN_TC_PROF
            const long long s3 = clock64();
            ps[2] += s3 - s2;
            const bool do_compact = __any_sync(0xffffffffu, sel.cnt >= trig);
            if (do_compact) {
              sel = sel_compact(sel, mv, mi, NT, p.k);
              ps[5] += 1;
              if (!warm) warm = __all_sync(0xffffffffu, sel.have == p.k || !live);
            }
            ps[3] += clock64() - s3;
#else
            if (__any_sync(0xffffffffu, sel.cnt >= trig)) {
              sel = sel_compact(sel, mv, mi, NT, p.k);
              if (!warm) warm = __all_sync(0xffffffffu, sel.have == p.k || !live);
            }
#endif
            }
            LGCN_PROF_T(c3t);
            LGCN_PROF_ADD(3, c2t, c3t);
#ifdef LGCN_TC_PROF
            pf[6] += 1;
#endif
          }
        } else {
          float m4[4];
          if (max32_h(r, m4) > sel.thr) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              if (m4[b] > sel.thr) {
#pragma unroll
                for (int i = 8 * b; i < 8 * b + 8; ++i) {
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    const float v = h == 0 ? h2_lo(r[i]) : h2_hi(r[i]);
                    if (v > sel.thr) {
                      const SelWalk w = sel_append(sel, pp, next_pos, my_pos, my_npos, v, item0 + 2 * i + h,
                                                   p.m_items, p.mask_value, mv, mi, NT, p.cap, p.k);
                      sel = w.s;
                      pp = w.pp;
                      next_pos = w.next;
                      ahead_pos = pp + 1 < my_npos ? __ldg(my_pos + pp + 1) : 0x7fffffff;
                    }
                  }
                }
              }
            }
          }
          if (__any_sync(0xffffffffu, sel.cnt >= trig)) sel = sel_compact(sel, mv, mi, NT, p.k);
        }
      }
      // one arrival per warp (128 same-address mbarrier arrives per hand-off serialise in shared memory)
      LGCN_PROF_T(q4);
      if (!released) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * (a * MT + mt));
      }
      LGCN_PROF_T(q5);
      LGCN_PROF_ADD(4, q4, q5);
      LGCN_STAMP(j, 4, blockIdx.x == 0 && ew == 0 && lane == 0);
    }
#ifdef LGCN_TC_PROF
    if (blockIdx.x == 0 && ew == 0 && lane == 0) {
      for (int i = 0; i < 5; ++i) g_tc_prof[i] = pf[i];
      g_tc_prof[5] = n_tiles; g_tc_prof[6] = pf[6]; g_tc_prof[7] = clock64() - pf_begin;
      for (int i = 0; i < 6; ++i) g_tc_prof[16 + i] = ps[i];
    }
#endif
    }

    sel = sel_finalize(sel, mv, mi, NT, p.k);
    ccnt[t] = sel.cnt;
    if (GROUPS > 1) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
    if (cg == 0 && live) {
      // merge the CG sorted partial lists of this row (the column slices of its user tile) by
      // (value desc, item id asc); slice g of this row lives in slot t + 128 * g
      int head[CG], num[CG];
#pragma unroll
      for (int g = 0; g < CG; ++g) {
        head[g] = 0;
        num[g] = g == 0 ? sel.cnt : ccnt[t + 128 * g];
      }
      for (int o = 0; o < p.k; ++o) {
        float v = -INFINITY;
        int id = -1, from = -1;
#pragma unroll
        for (int g = 0; g < CG; ++g) {
          if (head[g] < num[g]) {
            const float vg = cval[head[g] * NT + t + 128 * g];
            const int xg = cidx[head[g] * NT + t + 128 * g];
            if (from < 0 || vg > v || (vg == v && xg < id)) { v = vg; id = xg; from = g; }
          }
        }
#pragma unroll
        for (int g = 0; g < CG; ++g)
          if (g == from) ++head[g];
        p.out_idx[grow * p.k + o] = id;
        p.out_val[grow * p.k + o] = from < 0 ? -INFINITY : v;
      }
    }
  }

  // ------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // no CTA leaves while a peer may still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

constexpr size_t kSmemLimit = 227 * 1024;

// Tuning overrides (A-B runs, pipeline experiments): read from the environment ONCE, at the first
// launch of the process, never on the per-launch path.
struct Tuning {
  int debug_mode, trig, split, stages;
  uint32_t wait_hint;
  std::string layout;
};
static const Tuning& tuning() {
  static const Tuning t = [] {
    Tuning x;
    const char* dbg = getenv("LGCN_TC_DEBUG");
    x.debug_mode = dbg ? atoi(dbg) : 0;
    const char* tg = getenv("LGCN_TC_TRIG");
    x.trig = tg ? atoi(tg) : 0;
    const char* sp = getenv("LGCN_TC_SPLIT");
    x.split = sp ? atoi(sp) : 1;
    if (x.split != 1 && x.split != 2 && x.split != 4 && x.split != 8) x.split = 1;
    const char* sg = getenv("LGCN_TC_STAGES");
    x.stages = sg ? atoi(sg) : 0;
    const char* wh = getenv("LGCN_TC_WAIT_NS");
    x.wait_hint = wh ? (uint32_t)atoll(wh) : 0x989680u;
    const char* le = getenv("LGCN_TC_LAYOUT");
    x.layout = le ? le : "auto";
    return x;
  }();
  return t;
}

template <int D, int TN, int GROUPS, int MT, int NST, bool RS = false, int CL = 1>
static int launch(const Params& p0, cudaStream_t st) {
  Params p = p0;
  constexpr int NT = GROUPS * 128;
  const size_t a_bytes = (size_t)MT * kUM * D * 2, b_bytes = (size_t)TN * D * 2;
  const size_t cand = (size_t)p.cap * NT * 8 + (size_t)NT * 4;
  const size_t tail = (4 * 8 + 1) * 8 + 16;
  if (a_bytes + cand + tail + 2 * b_bytes > kSmemLimit) {
    set_last_error("k=%d does not fit the tensor-core top-k shared-memory budget", p.k);
    return LGCN_ERR_UNSUPPORTED;
  }
  int stages = (int)((kSmemLimit - a_bytes - cand - tail) / b_bytes);
  if (stages > 8) stages = 8;   // kMaxStages barriers
  if (tuning().stages > 1 && tuning().stages < stages) stages = tuning().stages;
  p.stages = stages;
  const size_t smem = a_bytes + (size_t)stages * b_bytes + cand + tail;
  const int grid = ((p.n_eval + kUM * MT * CL - 1) / (kUM * MT * CL)) * CL;   // whole clusters
  const int threads = (kFrontWarps + 4 * GROUPS) * 32;
#define LGCN_TC_LAUNCH(DUMP_, ACC_)                                                                     \
  do {                                                                                                    \
    auto kern = score_topk_tc_kernel<D, TN, GROUPS, DUMP_, ACC_, MT, NST, RS && !ACC_, CL>;               \
    LGCN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    cudaLaunchConfig_t lc = {};                                                                           \
    lc.gridDim = dim3((unsigned)grid);                                                                    \
    lc.blockDim = dim3((unsigned)threads);                                                                \
    lc.dynamicSmemBytes = smem;                                                                           \
    lc.stream = st;                                                                                       \
    cudaLaunchAttribute at[1];                                                                            \
    at[0].id = cudaLaunchAttributeClusterDimension;                                                       \
    at[0].val.clusterDim.x = CL;                                                                          \
    at[0].val.clusterDim.y = 1;                                                                           \
    at[0].val.clusterDim.z = 1;                                                                           \
    lc.attrs = at;                                                                                        \
    lc.numAttrs = CL > 1 ? 1 : 0;                                                                         \
    LGCN_CUDA_OK(cudaLaunchKernelEx(&lc, kern, p));                                                       \
  } while (0)
  constexpr bool kAcc16Ok = (TN / 64) % (GROUPS / MT) == 0;   // a group needs whole 64-column chunks
  if (p.acc16) {
    if constexpr (kAcc16Ok) {
      if (p.dense != nullptr) LGCN_TC_LAUNCH(true, true); else LGCN_TC_LAUNCH(false, true);
    } else {
      set_last_error("f16 accumulators need TN/64 divisible by the column groups per user tile");
      return LGCN_ERR_UNSUPPORTED;
    }
  } else {
    if (p.dense != nullptr) LGCN_TC_LAUNCH(true, false); else LGCN_TC_LAUNCH(false, false);
  }
#undef LGCN_TC_LAUNCH
  LGCN_LAUNCH_OK();
  return 0;
}

static size_t smem_need(int d, int tn, int groups, int mt, int cap) {
  return (size_t)mt * kUM * d * 2 + (size_t)cap * groups * 128 * 8 + (size_t)groups * 128 * 4 + (4 * 8 + 1) * 8 + 16 +
         2 * (size_t)tn * d * 2;
}

// One configuration: pack both operands for (TN, MT) and launch.
template <int D, int TN, int GROUPS, int MT, int NST, bool RS = false, int CL = 1>
static int run_cfg(const float* user_emb, const float* item_emb, const int64_t* user_ids, int n_eval,
                   int m_items, const int64_t* pos_rowptr, const int32_t* pos_sorted, int k, int cap,
                   float mask_value, int32_t* out_idx, float* out_val, float* dense, void* workspace,
                   size_t workspace_bytes, int acc16, cudaStream_t st) {
  const int64_t n_ut = ((n_eval + kUM * MT * CL - 1) / (kUM * MT * CL)) * MT * CL, n_it = ((int64_t)m_items + TN - 1) / TN;
  const size_t a_total = (size_t)n_ut * kUM * D * 2, b_total = (size_t)n_it * TN * D * 2;
  if (workspace == nullptr || workspace_bytes < a_total + b_total) {
    set_last_error("score_topk (bf16) needs a %zu-byte workspace, got %zu", a_total + b_total,
                   workspace_bytes);
    return LGCN_ERR_INVALID_ARG;
  }
  uint4* a_packed = reinterpret_cast<uint4*>(workspace);
  uint4* b_packed = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(workspace) + a_total);
  const int cap_blocks = kSmCount * 8;
  {
    const int64_t tot = n_ut * kUM * (D / 8);
    const int64_t blocks = (tot + 255) / 256;
    if (acc16)
      pack_bf16_kernel<true><<<(unsigned)(blocks < cap_blocks ? blocks : cap_blocks), 256, 0, st>>>(
          user_emb, user_ids, n_eval, D, kUM, n_ut, a_packed);
    else
      pack_bf16_kernel<false><<<(unsigned)(blocks < cap_blocks ? blocks : cap_blocks), 256, 0, st>>>(
          user_emb, user_ids, n_eval, D, kUM, n_ut, a_packed);
    LGCN_LAUNCH_OK();
  }
  {
    const int64_t tot = n_it * TN * (D / 8);
    const int64_t blocks = (tot + 255) / 256;
    if (acc16)
      pack_bf16_kernel<true><<<(unsigned)(blocks < cap_blocks ? blocks : cap_blocks), 256, 0, st>>>(
          item_emb, nullptr, m_items, D, TN, n_it, b_packed);
    else
      pack_bf16_kernel<false><<<(unsigned)(blocks < cap_blocks ? blocks : cap_blocks), 256, 0, st>>>(
          item_emb, nullptr, m_items, D, TN, n_it, b_packed);
    LGCN_LAUNCH_OK();
  }
  Params p;
  p.a_packed = a_packed; p.b_packed = b_packed; p.user_ids = user_ids;
  p.n_eval = n_eval; p.m_items = m_items; p.pos_rowptr = pos_rowptr; p.pos_sorted = pos_sorted;
  p.k = k; p.cap = cap; p.mask_value = mask_value; p.out_idx = out_idx; p.out_val = out_val; p.dense = dense;
  p.stages = 2;
  p.acc16 = acc16;
  p.debug_mode = tuning().debug_mode;
  p.trig = tuning().trig;
  p.split = tuning().split;
  p.wait_hint = tuning().wait_hint;
  return launch<D, TN, GROUPS, MT, NST, RS, CL>(p, st);
}

#define LGCN_TC_ARGS user_emb, item_emb, user_ids, n_eval, m_items, pos_rowptr, pos_sorted, k
#define LGCN_TC_TAIL mask_value, out_idx, out_val, dense, workspace, workspace_bytes, acc16, st

// Configuration choice.  TN1 = item-tile width of the one-user-tile layouts (256; 128 at d = 128).
// Layout names (LGCN_TC_LAYOUT overrides the automatic choice for A-B runs); measured at
// 75 776 users x 2 M items, d = 64, k = 20 with the final candidate path (profiles/r02_tc_experiments.log):
//   m2g2  MT = 2 user tiles x TN = 128, one epilogue group per user tile, 6-deep operand ring
//         (default for k <= 24, d <= 64, fp32 accumulators)                           1054-1061 TFLOP/s
//   m2g4  the same with two column groups per user tile (16 epilogue warps, k <= 20)       865
//   m2rl  m2g2 with the register-staged epilogue                                           740-786
//   m2s4  four TMEM stages x 64-column tiles (N = 64 MMAs are issue bound)                 698 (mid-round)
//   m2c2 / m2c4  m2g2 in clusters of 2 / 4 CTAs sharing the item-tile stream (multicast)   770 / 415 (mid-round)
//   g2    one user tile x TN1, two column groups (the round-1 layout; d = 128 uses it)     560 (mid-round)
// 24 < k <= 112: one group.
template <int D, int TN1>
static int run(const float* user_emb, const float* item_emb, const int64_t* user_ids, int n_eval,
               int m_items, const int64_t* pos_rowptr, const int32_t* pos_sorted, int k,
               float mask_value, int32_t* out_idx, float* out_val, float* dense, void* workspace,
               size_t workspace_bytes, int acc16, cudaStream_t st) {
  const std::string& layout = tuning().layout;
  if constexpr (D <= 64) {
    if (k <= 24 && !acc16) {
      if (layout == "m2rl") return run_cfg<D, 128, 2, 2, 2, true>(LGCN_TC_ARGS, 48, LGCN_TC_TAIL);
      if (layout == "m2s4") return run_cfg<D, 64, 2, 2, 4>(LGCN_TC_ARGS, 48, LGCN_TC_TAIL);   // 4 TMEM stages x 64 columns
      if (layout == "m2c2") return run_cfg<D, 128, 2, 2, 2, false, 2>(LGCN_TC_ARGS, 48, LGCN_TC_TAIL);
      if (layout == "m2c4") return run_cfg<D, 128, 2, 2, 2, false, 4>(LGCN_TC_ARGS, 48, LGCN_TC_TAIL);
      if (layout == "m2g2" || layout == "auto") return run_cfg<D, 128, 2, 2, 2>(LGCN_TC_ARGS, 48, LGCN_TC_TAIL);
    }
    if (k <= 20 && !acc16 && layout == "m2g4") return run_cfg<D, 128, 4, 2, 2>(LGCN_TC_ARGS, 32, LGCN_TC_TAIL);
  }
  if constexpr ((TN1 / 64) % 4 == 0) {
    if (k <= 20 && layout == "g4" && smem_need(D, TN1, 4, 1, 32) <= kSmemLimit)
      return run_cfg<D, TN1, 4, 1, 2>(LGCN_TC_ARGS, 32, LGCN_TC_TAIL);
  }
  if (k <= 24) return run_cfg<D, TN1, 2, 1, 2>(LGCN_TC_ARGS, 48, LGCN_TC_TAIL);
  int cap = 2 * k;
  if (cap < 64) cap = 64;
  if (cap > 128) cap = 128;
  if (k > cap - 16) {
    set_last_error("tensor-core top-k supports k <= 112 (got %d); use precision LGCN_F32", k);
    return LGCN_ERR_UNSUPPORTED;
  }
  return run_cfg<D, TN1, 1, 1, 2>(LGCN_TC_ARGS, cap, LGCN_TC_TAIL);
}
#undef LGCN_TC_ARGS
#undef LGCN_TC_TAIL

}  // namespace tc

#ifdef LGCN_TC_PROF
extern "C" int lgcn_debug_tc_prof(long long* out16) {
  return (int)cudaMemcpyFromSymbol(out16, tc::g_tc_prof, sizeof(long long) * 24);
}
extern "C" int lgcn_debug_tc_stamps(long long* out40) {
  return (int)cudaMemcpyFromSymbol(out40, tc::g_tc_stamp, sizeof(long long) * 40);
}
#endif

size_t score_topk_tc_workspace(int64_t n_eval, int64_t m_items, int d) {
  // covers every layout: user tiles rounded up to whole clusters (MT = 2 x CL = 4), item tiles at the widest TN
  const int tn = d >= 128 ? 128 : 256;
  const int64_t n_ut = ((n_eval + 8 * tc::kUM - 1) / (8 * tc::kUM)) * 8, n_it = (m_items + tn - 1) / tn;
  return (size_t)n_ut * tc::kUM * d * 2 + (size_t)n_it * tn * d * 2;
}

int score_topk_tc(const float* user_emb, const float* item_emb, const int64_t* user_ids,
                  int64_t n_eval, int64_t m_items, int d, const int64_t* pos_rowptr,
                  const int32_t* pos_sorted, int k, float mask_value, int32_t* out_idx,
                  float* out_val, float* dense, void* workspace, size_t workspace_bytes, int acc16,
                  cudaStream_t st) {
  switch (d) {
    case 32: return tc::run<32, 256>(user_emb, item_emb, user_ids, (int)n_eval, (int)m_items, pos_rowptr, pos_sorted, k, mask_value, out_idx, out_val, dense, workspace, workspace_bytes, acc16, st);
    case 64: return tc::run<64, 256>(user_emb, item_emb, user_ids, (int)n_eval, (int)m_items, pos_rowptr, pos_sorted, k, mask_value, out_idx, out_val, dense, workspace, workspace_bytes, acc16, st);
    case 128: return tc::run<128, 128>(user_emb, item_emb, user_ids, (int)n_eval, (int)m_items, pos_rowptr, pos_sorted, k, mask_value, out_idx, out_val, dense, workspace, workspace_bytes, acc16, st);
    default:
      set_last_error("unsupported embedding width d=%d (supported: 32, 64, 128)", d);
      return LGCN_ERR_UNSUPPORTED;
  }
}

}  // namespace lgcn
