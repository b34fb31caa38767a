// K-layer LightGCN propagation: CSR SpMM with the symmetric normalisation folded
// into dinv, fused with the layer-sum / Horner / Adam row epilogue.
//
// Replaces torch.sparse.mm (reference model/MF.py:200,204), PyG LGConv
// (model/lgcn.py:82), the layer mean (model/lgcn.py:83-84) and, in grad_mode 2,
// optim.Adam.step() (model/lgcn.py:132).  See include/lgcn_b200.h for the maths.
//
// Work decomposition (HBM/L2-gather bound, no tensor cores):
//   * a "group" of LPR = row_bytes/16 lanes owns one light row (deg <= HUB_DEG):
//     every lane keeps 16 bytes of the row in registers, so one gathered
//     neighbour row is one fully coalesced LPR*16-byte request;
//   * column indices are fetched LPR at a time (one coalesced load per group),
//     broadcast with shuffles, and the next chunk is prefetched while up to
//     UNROLL neighbour rows are in flight per lane;
//   * rows longer than HUB_DEG are cut into CTA-wide segments (<= SEG_EDGES
//     edges): the CTA's groups stride over the segment, reduce through shared
//     memory in a fixed order, and multi-segment hubs are finished by the last
//     arriving CTA which adds the partials in segment order (deterministic);
//   * light rows are visited in degree-descending order so the longest chains
//     start first and the tail of the grid is made of cheap rows; each has a packed
//     16-byte descriptor {row, degree, first edge} so the chain is descriptor ->
//     column ids -> neighbour rows, with the epilogue operands prefetched up front.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace lgcn {

static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// 128-thread CTAs, 8 per SM: the same 32 resident warps as 4 x 256 but the grid drains in finer
// steps (cfg-2: 44.8 vs 49.0 us per launch)
#ifndef LGCN_SPMM_BLOCK
#define LGCN_SPMM_BLOCK 128
#endif
constexpr int kBlock = LGCN_SPMM_BLOCK;
#ifndef LGCN_SPMM_UNROLL
#define LGCN_SPMM_UNROLL 4
#endif
#ifndef LGCN_SPMM_MINBLOCKS
#define LGCN_SPMM_MINBLOCKS (2048 / LGCN_SPMM_BLOCK / 2)   // 32 warps per SM at <= 64 registers
#endif

template <int D, bool SRC_BF16>
struct RowCfg {
  static constexpr int kRowBytes = D * (SRC_BF16 ? 2 : 4);
  static constexpr int kLPR = kRowBytes / 16;          // lanes per row
  static constexpr int kEPL = SRC_BF16 ? 8 : 4;        // elements per lane
  static constexpr int kGroups = kBlock / kLPR;        // groups per CTA
  static constexpr int kUnroll = kLPR >= LGCN_SPMM_UNROLL ? LGCN_SPMM_UNROLL : kLPR; // neighbour rows in flight
  static_assert(kRowBytes % 16 == 0 && kLPR >= 2 && kLPR <= 32, "unsupported row width");
  static_assert((kLPR & (kLPR - 1)) == 0, "lanes per row must be a power of two");
};

struct GraphDev {
  const int64_t* rowptr;
  const int32_t* col;
  const float* dinv;
  const int4* light_desc;
  int64_t n_light;
  const int32_t* seg_row;
  const int64_t* seg_begin;
  const int32_t* seg_len;
  const int32_t* seg_hub;
  int n_seg;
  const int32_t* hub_seg0;
  const int32_t* hub_nseg;
  int32_t* hub_counter;
  float* partial;
};

struct EpiDev {
  const float* edge_w;     // [nnz] per-slot weight (edge dropout: mask / keep_prob) or NULL
  const float* src_scale;  // w_j of the gather (first layer) and the pre-scale of dst; dinv unless overridden
  const float* dst_scale;  // x_i = dst_scale[i] * s_i; dinv unless overridden
  const void* src;
  void* dst;
  const float* base;
  const float* acc_in;
  float* acc_out;
  float acc_scale;
  int grad_mode;
  float inv_layers;
  float reg_coef;
  int32_t* cnt;
  float* emb;
  float* grad;
  float* adam_m;
  float* adam_v;
  const float* adam_hp;
  float one_minus_b1, beta2, one_minus_b2, eps;
  int zero_base;
  int push_emb;
  int n_peer;
  int64_t dst_row_off;
  void* peer[LGCN_MAX_PEERS];
  void* mc;   // multicast mapping of the gathered buffer (NVLS) or NULL
  int64_t route_rows;  // > 0: row i goes to peer[i / route_rows] only, at row dst_row_off + i % route_rows
};

// Sum of w_j * SRC[col[e]] over e = e0, e0+1, .. inside [e0, e_end) taken in
// chunks of LPR edges that advance by `stride` edges (stride == LPR: the whole
// range; stride == kGroups*LPR: this group's share of a hub segment).
//
// The kernel issues on ~50 % of the cycles (ncu), so the inner loop is kept to
// SHFL + IMAD.WIDE + LDG.128 + ISETP + 4 predicated FADD/FFMA per neighbour row:
//   * ALL 32 lanes call this together and every shuffle uses the full mask; trip counts are
//     made warp-uniform (max over the warp's groups), so there is no divergence bookkeeping
//     (MATCH/REDUX/VOTE + BRA.DIV per shuffle with a partial mask);
//   * loads are never predicated and need no zero-fill select: a slot past the end of the
//     neighbour list re-reads row 0 of SRC (an L1 hit) and is accumulated with weight 0;
//   * the row address is one mad.wide on a lane-folded base pointer.
// WMODE selects the slot weight at compile time: 0 none (pre-scaled source, plain sum), 1 dinv[col]
// (scale_src), 2 edge_w[slot] (edge dropout on a pre-scaled source), 3 both.
template <int D, bool SRC_BF16, int WMODE>
__device__ __forceinline__ void gather_sum(const void* __restrict__ src,
                                           const int32_t* __restrict__ col,
                                           const float* __restrict__ dinv,
                                           const float* __restrict__ edge_w, int64_t e0,
                                           int64_t e_end, int64_t stride, int lig,
                                           float (&acc)[RowCfg<D, SRC_BF16>::kEPL]) {
  constexpr bool SCALE_SRC = WMODE != 0;   // a weighted gather (the name the loop below uses)
  using C = RowCfg<D, SRC_BF16>;
  constexpr int LPR = C::kLPR, EPL = C::kEPL, U = C::kUnroll;
  constexpr unsigned kFull = 0xffffffffu;
  const char* base = reinterpret_cast<const char*>(src) + lig * 16;

  const int64_t span = e_end - e0;
  int n_chunks = span > 0 ? (int)((span + stride - 1) / stride) : 0;
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) n_chunks = max(n_chunks, __shfl_xor_sync(kFull, n_chunks, o));

  int nxt_c = 0;
  float nxt_w = 0.f;
  if (e0 + lig < e_end) {
    nxt_c = ldg_stream_i32(col + e0 + lig);
    if (WMODE == 1) nxt_w = __ldg(dinv + nxt_c);
    if (WMODE == 2) nxt_w = __ldg(edge_w + e0 + lig);
    if (WMODE == 3) nxt_w = __ldg(dinv + nxt_c) * __ldg(edge_w + e0 + lig);
  }
  int64_t e = e0;
  for (int ch = 0; ch < n_chunks; ++ch, e += stride) {
    const int64_t left = e_end - e;
    const int cnt = left <= 0 ? 0 : (left < LPR ? (int)left : LPR);
    const int cur_c = nxt_c;
    const float cur_w = nxt_w;
    nxt_c = 0;
    nxt_w = 0.f;
    if (e + stride + lig < e_end) {  // prefetch the next chunk of column ids
      nxt_c = ldg_stream_i32(col + e + stride + lig);
      if (WMODE == 1) nxt_w = __ldg(dinv + nxt_c);
      if (WMODE == 2) nxt_w = __ldg(edge_w + e + stride + lig);
      if (WMODE == 3) nxt_w = __ldg(dinv + nxt_c) * __ldg(edge_w + e + stride + lig);
    }
    int nb = (cnt + U - 1) / U;
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) nb = max(nb, __shfl_xor_sync(kFull, nb, o));
    for (int bt = 0; bt < nb; ++bt) {
      uint4 v[U];
      float w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = __shfl_sync(kFull, cur_c, bt * U + u, LPR);
        if (SCALE_SRC) w[u] = __shfl_sync(kFull, cur_w, bt * U + u, LPR);
        const char* rp;
        asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(rp) : "r"(c), "n"(C::kRowBytes), "l"(base));
        v[u] = ldg_row16(rp);   // L1-allocating: bypassing L1 costs 30 % (hot item rows repeat)
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // weight 0 retires a padding slot without a predicated accumulator write; a real slot of a
        // pre-scaled source has weight 1, and fma(1, f, acc) == acc + f exactly
        const float wu = SCALE_SRC ? w[u] : (bt * U + u < cnt ? 1.f : 0.f);
        if (SRC_BF16) {
          float f[8];
          unpack_bf16x8(v[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j % EPL] = fmaf(wu, f[j], acc[j % EPL]);
        } else {
          const float f[4] = {__uint_as_float(v[u].x), __uint_as_float(v[u].y),
                              __uint_as_float(v[u].z), __uint_as_float(v[u].w)};
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = fmaf(wu, f[j], acc[j]);
        }
      }
    }
  }
}

// Row epilogue; lane `lig` holds s[0..EPL) = elements [lig*EPL, lig*EPL+EPL) of row.
// Operands of the row epilogue that do not depend on the gather: fetched BEFORE it so their
// latency hides behind the neighbour loads instead of extending the per-row dependency chain.
template <int EPL>
struct RowPre {
  float di;      // dst_scale[row]
  float ds;      // src_scale[row] (== di for the symmetric normalisation)
  float v[EPL];  // base[row] (backward) or acc_in[row] (forward)
};

template <int D, int EPL>
__device__ __forceinline__ RowPre<EPL> row_prefetch(const EpiDev& p, int64_t row, int lig) {
  RowPre<EPL> r;
  r.di = __ldg(p.dst_scale + row);
  r.ds = p.src_scale == p.dst_scale ? r.di : __ldg(p.src_scale + row);
  const float* src = p.base != nullptr ? p.base : p.acc_in;
  const int64_t off = row * D + lig * EPL;
#pragma unroll
  for (int q = 0; q < EPL / 4; ++q) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src != nullptr) b = ld_f4(src + off + 4 * q);
    r.v[4 * q + 0] = b.x; r.v[4 * q + 1] = b.y; r.v[4 * q + 2] = b.z; r.v[4 * q + 3] = b.w;
  }
  return r;
}

// dst[row] = z.  n_peer == 0: plain local store.  n_peer > 0: the all-gather is fused here — the row
// goes to every reader's gathered buffer over NVLink (peer-mapped pointers), at this rank's row block.
template <int D, int EPL, bool DST_BF16>
__device__ __forceinline__ void store_dst_row(const EpiDev& p, int64_t row, int lig, const float (&z)[EPL]) {
  if (p.mc != nullptr) {
    // one store, replicated by the NVSwitch into every rank's gathered buffer
    const int64_t moff = (p.dst_row_off + row) * D + lig * EPL;
    if (DST_BF16) {
      __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.mc) + moff;
      if (EPL == 8) {
        asm volatile("multimem.st.weak.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dp), "r"(pack_bf16x2(z[0], z[1])),
                     "r"(pack_bf16x2(z[2], z[3])), "r"(pack_bf16x2(z[4 % EPL], z[5 % EPL])),
                     "r"(pack_bf16x2(z[6 % EPL], z[7 % EPL]))
                     : "memory");
      } else {
        asm volatile("multimem.st.weak.global.v2.bf16x2 [%0], {%1, %2};" ::"l"(dp), "r"(pack_bf16x2(z[0], z[1])),
                     "r"(pack_bf16x2(z[2], z[3]))
                     : "memory");
      }
    } else {
      float* dp = reinterpret_cast<float*>(p.mc) + moff;
#pragma unroll
      for (int qq = 0; qq < EPL / 4; ++qq)
        asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp + 4 * qq), "f"(z[4 * qq]),
                     "f"(z[4 * qq + 1]), "f"(z[4 * qq + 2]), "f"(z[4 * qq + 3])
                     : "memory");
    }
    return;
  }
  // route_rows: every row has ONE destination, the peer that owns it (partial sums of the reduce partition)
  const bool routed = p.route_rows > 0 && p.n_peer > 0;
  const int q_routed = routed ? (int)(row / p.route_rows) : 0;
  const int64_t row_d = routed ? row - (int64_t)q_routed * p.route_rows : row;
  const int n_dst = routed ? 1 : (p.n_peer > 0 ? p.n_peer : 1);
  const int64_t doff = p.n_peer > 0 ? (p.dst_row_off + row_d) * D + lig * EPL : row * D + lig * EPL;
  for (int q = 0; q < n_dst; ++q) {
    void* base_ptr = p.n_peer > 0 ? p.peer[routed ? q_routed : q] : p.dst;
    if (DST_BF16) {
      __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(base_ptr) + doff;
      if (EPL == 8) {
        float z8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) z8[j] = z[j % EPL];
        st_u4(dp, pack_bf16x8(z8));
      } else {
        uint2 w2;
        w2.x = pack_bf16x2(z[0], z[1]);
        w2.y = pack_bf16x2(z[2], z[3]);
        *reinterpret_cast<uint2*>(dp) = w2;
      }
    } else {
      float* dp = reinterpret_cast<float*>(base_ptr) + doff;
#pragma unroll
      for (int qq = 0; qq < EPL / 4; ++qq)
        st_f4(dp + 4 * qq, make_float4(z[4 * qq], z[4 * qq + 1], z[4 * qq + 2], z[4 * qq + 3]));
    }
  }
}

template <int D, int EPL, bool DST_BF16>
__device__ __forceinline__ void row_epilogue(const EpiDev& p, const RowPre<EPL>& pre, int64_t row,
                                             int lig, unsigned gmask, const float (&s)[EPL]) {
  const float di = pre.di;
  const int64_t off = row * D + lig * EPL;
  float x[EPL], t[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) x[j] = di * s[j];

  if (p.base != nullptr) {
#pragma unroll
    for (int j = 0; j < EPL; ++j) t[j] = pre.v[j] + x[j];
  } else {
#pragma unroll
    for (int j = 0; j < EPL; ++j) t[j] = x[j];
  }

  // grad_mode 2 with push_emb: dst receives the pre-scaled UPDATED embedding row instead (below)
  if (p.dst != nullptr && !(p.grad_mode == 2 && p.push_emb)) {
    float z[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) z[j] = pre.ds * t[j];
    store_dst_row<D, EPL, DST_BF16>(p, row, lig, z);
  }

  if (p.acc_out != nullptr) {  // base == nullptr on this path, so pre.v holds acc_in[row]
#pragma unroll
    for (int q = 0; q < EPL / 4; ++q)
      st_f4(p.acc_out + off + 4 * q,
            make_float4((pre.v[4 * q + 0] + x[4 * q + 0]) * p.acc_scale,
                        (pre.v[4 * q + 1] + x[4 * q + 1]) * p.acc_scale,
                        (pre.v[4 * q + 2] + x[4 * q + 2]) * p.acc_scale,
                        (pre.v[4 * q + 3] + x[4 * q + 3]) * p.acc_scale));
  }

  if (p.grad_mode != 0) {
    const float rc = p.reg_coef * float(p.cnt[row]);
    float step_size = 0.f, bc2_sqrt = 1.f;
    if (p.grad_mode == 2) {
      step_size = __ldg(p.adam_hp);
      bc2_sqrt = __ldg(p.adam_hp + 1);
    }
    float znew[EPL];   // push_emb: src_scale[row] * updated embedding
#pragma unroll
    for (int q = 0; q < EPL / 4; ++q) {
      const float4 e4 = ld_f4(p.emb + off + 4 * q);
      const float e[4] = {e4.x, e4.y, e4.z, e4.w};
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = fmaf(rc, e[j], t[4 * q + j] * p.inv_layers);
      if (p.grad_mode == 1) {
        st_f4(p.grad + off + 4 * q, make_float4(g[0], g[1], g[2], g[3]));
      } else {
        const float4 m4 = ld_f4(p.adam_m + off + 4 * q);
        const float4 v4 = ld_f4(p.adam_v + off + 4 * q);
        float m[4] = {m4.x, m4.y, m4.z, m4.w};
        float v[4] = {v4.x, v4.y, v4.z, v4.w};
        float en[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          m[j] = fmaf(g[j] - m[j], p.one_minus_b1, m[j]);
          v[j] = fmaf(p.one_minus_b2 * g[j], g[j], v[j] * p.beta2);
          const float denom = sqrtf(v[j]) / bc2_sqrt + p.eps;
          en[j] = e[j] - step_size * (m[j] / denom);
        }
        st_f4(p.adam_m + off + 4 * q, make_float4(m[0], m[1], m[2], m[3]));
        st_f4(p.adam_v + off + 4 * q, make_float4(v[0], v[1], v[2], v[3]));
        st_f4(p.emb + off + 4 * q, make_float4(en[0], en[1], en[2], en[3]));
#pragma unroll
        for (int j = 0; j < 4; ++j) znew[4 * q + j] = pre.ds * en[j];
        if (p.zero_base)  // leave G clean for the next step's scatter
          st_f4(const_cast<float*>(p.base) + off + 4 * q, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    if (p.grad_mode == 2) {
      // next step's pre-scaled layer-0 source, straight from the registers that hold the new row
      if (p.push_emb && p.dst != nullptr) store_dst_row<D, EPL, DST_BF16>(p, row, lig, znew);
      __syncwarp(gmask);  // every lane of the group has read cnt[row]
      if (lig == 0) p.cnt[row] = 0;
    }
  }
}

template <int D, bool SRC_BF16, bool DST_BF16, int WMODE>
__global__ void __launch_bounds__(kBlock, LGCN_SPMM_MINBLOCKS)
spmm_layer_kernel(const GraphDev g, const EpiDev p) {
  using C = RowCfg<D, SRC_BF16>;
  constexpr int LPR = C::kLPR, EPL = C::kEPL, NG = C::kGroups;
  const int lig = threadIdx.x % LPR;   // lane in group
  const int grp = threadIdx.x / LPR;   // group in CTA
  const unsigned gmask = group_mask(LPR);

  float acc[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) acc[j] = 0.f;

  if ((int)blockIdx.x >= g.n_seg) {
    // ---------------- light rows: one group per row ----------------
    const int64_t gid = (int64_t)(blockIdx.x - g.n_seg) * NG + grp;
    const bool live = gid < g.n_light;  // a dead group still takes part in the warp's shuffles
    int4 dsc = make_int4(0, 0, 0, 0);
    if (live) dsc = __ldg(g.light_desc + gid);  // {row, degree, first edge lo, hi}: one hop, no rowptr
    const int64_t row = dsc.x;
    const int64_t b = (int64_t)(((uint64_t)(uint32_t)dsc.w << 32) | (uint32_t)dsc.z);
    RowPre<EPL> pre;
    if (live) pre = row_prefetch<D, EPL>(p, row, lig);
    gather_sum<D, SRC_BF16, WMODE>(p.src, g.col, p.src_scale, p.edge_w, b, b + dsc.y, LPR, lig, acc);
    if (live) row_epilogue<D, EPL, DST_BF16>(p, pre, row, lig, gmask, acc);
    return;
  }

  // ---------------- hub segment: the whole CTA on <= SEG_EDGES edges ----------------
  __shared__ float red[NG][D];
  __shared__ int s_last;
  const int seg = blockIdx.x;
  const int64_t row = g.seg_row[seg];
  const int64_t b = g.seg_begin[seg];
  const int64_t e = b + g.seg_len[seg];
  gather_sum<D, SRC_BF16, WMODE>(p.src, g.col, p.src_scale, p.edge_w, b + (int64_t)grp * LPR, e,
                                 (int64_t)NG * LPR, lig, acc);
#pragma unroll
  for (int j = 0; j < EPL; ++j) red[grp][lig * EPL + j] = acc[j];
  __syncthreads();
  float colsum = 0.f;
  if (threadIdx.x < D) {
#pragma unroll 8
    for (int q = 0; q < NG; ++q) colsum += red[q][threadIdx.x];
  }
  const int hub = g.seg_hub[seg];
  const int nseg = g.hub_nseg[hub];
  if (nseg > 1) {
    const int seg0 = g.hub_seg0[hub];
    if (threadIdx.x < D) g.partial[(int64_t)seg * D + threadIdx.x] = colsum;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int old = atomicAdd(g.hub_counter + hub, 1);
      s_last = (old == nseg - 1);
      if (old == nseg - 1) g.hub_counter[hub] = 0;  // self-reset for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < D) {
      // up to ~1000 segment partials of a hub row: accumulated in fp64 (a sequential fp32 sum of n
      // equal-sized terms drifts by ~n * 2^-24; 3e-5 on the largest cfg-3 hubs), still in segment order
      double cs = 0.0;
      for (int q = 0; q < nseg; ++q)
        cs += (double)__ldcg(g.partial + (int64_t)(seg0 + q) * D + threadIdx.x);
      colsum = (float)cs;
    }
  }
  __syncthreads();  // everyone is done reading red[*] before it is overwritten
  if (threadIdx.x < D) red[0][threadIdx.x] = colsum;
  __syncthreads();
  if (grp == 0) {
    float s[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) s[j] = red[0][lig * EPL + j];
    const RowPre<EPL> pre = row_prefetch<D, EPL>(p, row, lig);
    row_epilogue<D, EPL, DST_BF16>(p, pre, row, lig, gmask, s);
  }
}

template <int D, bool SRC_BF16, bool DST_BF16, int WMODE>
static int launch_layer(const GraphDev& g, const EpiDev& p, cudaStream_t st) {
  using C = RowCfg<D, SRC_BF16>;
  static_assert(D <= kBlock, "hub reduction assumes D <= block size");
  const int64_t light_blocks = (g.n_light + C::kGroups - 1) / C::kGroups;
  const int64_t grid = (int64_t)g.n_seg + light_blocks;
  if (grid == 0) return 0;
  if (grid > 0x7fffffffLL) {
    set_last_error("grid too large: %lld", (long long)grid);
    return LGCN_ERR_UNSUPPORTED;
  }
  spmm_layer_kernel<D, SRC_BF16, DST_BF16, WMODE><<<(unsigned)grid, kBlock, 0, st>>>(g, p);
  LGCN_LAUNCH_OK();
  return 0;
}

template <int D, bool SB, bool DB>
static int dispatch_wmode(int wmode, const GraphDev& g, const EpiDev& p, cudaStream_t st) {
  switch (wmode) {
    case 0: return launch_layer<D, SB, DB, 0>(g, p, st);
    case 1: if constexpr (!SB) return launch_layer<D, SB, DB, 1>(g, p, st); else break;
    case 2: return launch_layer<D, SB, DB, 2>(g, p, st);
    case 3: if constexpr (!SB) return launch_layer<D, SB, DB, 3>(g, p, st); else break;
  }
  set_last_error("scale_src needs an fp32 source");
  return LGCN_ERR_UNSUPPORTED;
}

template <int D>
static int dispatch_dtype(const lgcn_layer_args_t* a, const GraphDev& g, const EpiDev& p,
                          cudaStream_t st) {
  const bool sb = a->src_dtype == LGCN_BF16, db = a->dst_dtype == LGCN_BF16;
  const int wmode = (a->scale_src ? 1 : 0) | (a->edge_w != nullptr ? 2 : 0);
  if (!sb && !db) return dispatch_wmode<D, false, false>(wmode, g, p, st);
  if (!sb && db) return dispatch_wmode<D, false, true>(wmode, g, p, st);
  if (sb && db) return dispatch_wmode<D, true, true>(wmode, g, p, st);
  // bf16 source, fp32 destination: raw fp32 partial sums of a bf16-stored activation (reduce partition)
  if (wmode == 0) return launch_layer<D, true, false, 0>(g, p, st);
  set_last_error("a bf16 source with an fp32 destination supports neither scale_src nor edge_w");
  return LGCN_ERR_UNSUPPORTED;
}

}  // namespace lgcn

using namespace lgcn;

namespace lgcn {
struct PeerPtrs { void* p[LGCN_MAX_PEERS]; };

template <bool DST_BF16>
__global__ void __launch_bounds__(256)
scale_rows_push_kernel(const float* __restrict__ x, const float* __restrict__ dinv, int64_t n_vec, int d4,
                       PeerPtrs peers, int n_peer, int64_t dst_elem_off) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / d4;
    const float di = __ldg(dinv + row);
    const float4 v = ld_f4(x + 4 * i);
    const float z0 = di * v.x, z1 = di * v.y, z2 = di * v.z, z3 = di * v.w;
    for (int q = 0; q < n_peer; ++q) {
      if (DST_BF16) {
        uint2 w2;
        w2.x = pack_bf16x2(z0, z1);
        w2.y = pack_bf16x2(z2, z3);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(peers.p[q]) + dst_elem_off + 4 * i) = w2;
      } else {
        st_f4(reinterpret_cast<float*>(peers.p[q]) + dst_elem_off + 4 * i, make_float4(z0, z1, z2, z3));
      }
    }
  }
}
}  // namespace lgcn

extern "C" int lgcn_scale_rows_push(const float* x, const float* dinv, int64_t n_rows, int d, int dst_dtype,
                                    void* const* dst_peers, int n_dst_peers, int64_t dst_row_offset,
                                    lgcn_stream_t stream) {
  LGCN_CHECK_ARG(x && dinv && dst_peers, "null pointer argument");
  LGCN_CHECK_ARG(n_dst_peers >= 1 && n_dst_peers <= LGCN_MAX_PEERS, "n_dst_peers out of range");
  LGCN_CHECK_ARG(d > 0 && d % 4 == 0 && n_rows >= 0, "bad shape");
  if (n_rows == 0) return 0;
  lgcn::PeerPtrs pp;
  for (int q = 0; q < LGCN_MAX_PEERS; ++q) pp.p[q] = q < n_dst_peers ? dst_peers[q] : nullptr;
  const int64_t n_vec = n_rows * (d / 4);
  int64_t blocks = (n_vec + 255) / 256;
  if (blocks > (int64_t)lgcn::kSmCount * 16) blocks = (int64_t)lgcn::kSmCount * 16;
  if (dst_dtype == LGCN_BF16)
    lgcn::scale_rows_push_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        x, dinv, n_vec, d / 4, pp, n_dst_peers, dst_row_offset * d);
  else
    lgcn::scale_rows_push_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        x, dinv, n_vec, d / 4, pp, n_dst_peers, dst_row_offset * d);
  LGCN_LAUNCH_OK();
  return 0;
}

namespace lgcn {
// rows[i] = [tab_a[loc_i] | tab_b[loc_i]] for the ids this rank owns, stored to EVERY peer's
// [n_ids, width] buffer at row i.  Each id has exactly one owner, so after a barrier every rank
// holds all rows without a reduction (the sync-free replacement of the 3B-row all-reduce).
__global__ void __launch_bounds__(256)
exchange_rows_push_kernel(const float* __restrict__ tab_a, const float* __restrict__ tab_b, int d4,
                          const int64_t* __restrict__ padded_ids, int64_t n_ids, int64_t R, int rank,
                          PeerPtrs peers, int n_peer) {
  const int w4 = tab_b != nullptr ? 2 * d4 : d4;  // float4 per output row
  const int64_t total = n_ids * w4;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / w4;
    const int c = (int)(t - i * w4);
    const int64_t id = __ldg(padded_ids + i);
    if (id / R != rank) continue;
    const int64_t loc = id - (int64_t)rank * R;
    const float4 v = c < d4 ? ld_f4(tab_a + (loc * d4 + c) * 4) : ld_f4(tab_b + (loc * d4 + (c - d4)) * 4);
    for (int q = 0; q < n_peer; ++q) st_f4(reinterpret_cast<float*>(peers.p[q]) + t * 4, v);
  }
}
}  // namespace lgcn

extern "C" int lgcn_exchange_rows_push(const float* tab_a, const float* tab_b, int d, const int64_t* padded_ids,
                                       int64_t n_ids, int64_t rows_per_rank, int rank, void* const* dst_peers,
                                       int n_dst_peers, lgcn_stream_t stream) {
  LGCN_CHECK_ARG(tab_a && padded_ids && dst_peers, "null pointer argument");
  LGCN_CHECK_ARG(n_dst_peers >= 1 && n_dst_peers <= LGCN_MAX_PEERS, "n_dst_peers out of range");
  LGCN_CHECK_ARG(d > 0 && d % 4 == 0 && n_ids >= 0 && rows_per_rank > 0 && rank >= 0, "bad shape");
  if (n_ids == 0) return 0;
  lgcn::PeerPtrs pp;
  for (int q = 0; q < LGCN_MAX_PEERS; ++q) pp.p[q] = q < n_dst_peers ? dst_peers[q] : nullptr;
  const int64_t total = n_ids * (tab_b ? 2 : 1) * (d / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)lgcn::kSmCount * 16) blocks = (int64_t)lgcn::kSmCount * 16;
  lgcn::exchange_rows_push_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      tab_a, tab_b, d / 4, padded_ids, n_ids, rows_per_rank, rank, pp, n_dst_peers);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_abi_version(void) { return LGCN_ABI_VERSION; }
extern "C" const char* lgcn_last_error(void) { return lgcn::last_error(); }

namespace lgcn {
static int fill_epilogue(const lgcn_layer_args_t* a, const float* dinv, EpiDev& p);
}

extern "C" int lgcn_propagate_layer(const lgcn_graph_t* gh, const lgcn_layer_args_t* a,
                                    lgcn_stream_t stream) {
  LGCN_CHECK_ARG(gh != nullptr && a != nullptr, "null graph/args");
  LGCN_CHECK_ARG(gh->n_nodes >= 0 && gh->n_nodes < 0x7fffffffLL, "n_nodes must fit int32");
  LGCN_CHECK_ARG(a->src != nullptr, "src is null");
  LGCN_CHECK_ARG(gh->n_seg < 0x7fffffffLL, "too many hub segments");
  LGCN_CHECK_ARG(a->acc_out == nullptr || a->acc_in != nullptr, "acc_out needs acc_in");
  LGCN_CHECK_ARG(a->grad_mode >= 0 && a->grad_mode <= 2, "grad_mode must be 0, 1 or 2");
  if (a->grad_mode != 0) {
    LGCN_CHECK_ARG(a->emb != nullptr && a->cnt != nullptr, "grad_mode needs emb and cnt");
    LGCN_CHECK_ARG(a->grad_mode != 1 || a->grad != nullptr, "grad_mode 1 needs grad");
    LGCN_CHECK_ARG(a->grad_mode != 2 || (a->adam_m && a->adam_v && a->adam_hp && a->base),
                   "grad_mode 2 needs adam_m, adam_v, adam_hp and base");
  }
  if (gh->n_nodes == 0) return 0;

  GraphDev g;
  g.rowptr = gh->rowptr; g.col = gh->col; g.dinv = gh->dinv;
  g.light_desc = reinterpret_cast<const int4*>(gh->light_desc); g.n_light = gh->n_light;
  g.seg_row = gh->seg_row; g.seg_begin = gh->seg_begin; g.seg_len = gh->seg_len;
  g.seg_hub = gh->seg_hub; g.n_seg = (int)gh->n_seg;
  g.hub_seg0 = gh->hub_seg0; g.hub_nseg = gh->hub_nseg; g.hub_counter = gh->hub_counter;
  g.partial = gh->partial;

  EpiDev p;
  if (int rc = fill_epilogue(a, gh->dinv, p)) return rc;

  cudaStream_t st = (cudaStream_t)stream;
  switch (a->d) {
    case 32: return dispatch_dtype<32>(a, g, p, st);
    case 64: return dispatch_dtype<64>(a, g, p, st);
    case 128: return dispatch_dtype<128>(a, g, p, st);
    default:
      set_last_error("unsupported embedding width d=%d (supported: 32, 64, 128)", a->d);
      return LGCN_ERR_UNSUPPORTED;
  }
}

namespace lgcn {
// lgcn_layer_args_t -> the epilogue parameter block both kernels use
static int fill_epilogue(const lgcn_layer_args_t* a, const float* dinv, EpiDev& p) {
  p.edge_w = a->edge_w;
  p.src_scale = a->src_scale != nullptr ? a->src_scale : dinv;
  p.dst_scale = a->dst_scale != nullptr ? a->dst_scale : dinv;
  p.src = a->src; p.dst = a->dst; p.base = a->base;
  p.acc_in = a->acc_in; p.acc_out = a->acc_out; p.acc_scale = a->acc_scale;
  p.grad_mode = a->grad_mode; p.inv_layers = a->inv_layers; p.reg_coef = a->reg_coef;
  p.cnt = a->cnt; p.emb = a->emb; p.grad = a->grad;
  p.adam_m = a->adam_m; p.adam_v = a->adam_v; p.adam_hp = a->adam_hp;
  p.one_minus_b1 = (float)(1.0 - a->beta1);
  p.beta2 = (float)a->beta2;
  p.one_minus_b2 = (float)(1.0 - a->beta2);
  p.eps = (float)a->eps;
  p.zero_base = a->zero_base;
  p.push_emb = a->push_emb;
  p.mc = a->dst_multicast;
  LGCN_CHECK_ARG(a->dst_multicast == nullptr || a->dst != nullptr, "dst_multicast needs dst (it fixes the dtype)");
  LGCN_CHECK_ARG(!a->push_emb || (a->grad_mode == 2 && a->dst != nullptr), "push_emb needs grad_mode 2 and dst");
  LGCN_CHECK_ARG(a->n_dst_peers >= 0 && a->n_dst_peers <= LGCN_MAX_PEERS, "n_dst_peers out of range");
  p.n_peer = a->n_dst_peers;
  p.dst_row_off = a->dst_row_offset;
  for (int q = 0; q < LGCN_MAX_PEERS; ++q) p.peer[q] = q < a->n_dst_peers ? a->dst_peers[q] : nullptr;
  p.route_rows = a->dst_route_rows;
  LGCN_CHECK_ARG(a->dst_route_rows >= 0 && (a->dst_route_rows == 0 || a->n_dst_peers > 0), "dst_route_rows needs dst_peers");
  LGCN_CHECK_ARG(!(a->zero_base && a->base == a->src), "zero_base with src == base races");
  return 0;
}

// Owner-side half of the reduce partition: s_i = sum_q partials[q][i] (fixed order: deterministic), then the
// SAME row epilogue as the SpMM (x_i = dst_scale_i * s_i, layer sum, next layer's pre-scaled row pushed to the
// readers, Horner base, Adam).  One group of D/4 lanes per row, 16 bytes per lane.
template <int D, bool DST_BF16>
__global__ void __launch_bounds__(kBlock)
reduce_rows_kernel(const float* __restrict__ partials, int n_parts, int64_t part_rows, int64_t n_rows, const EpiDev p) {
  constexpr int LPR = D / 4, NG = kBlock / LPR;
  const int lig = threadIdx.x % LPR;
  const int64_t row = (int64_t)blockIdx.x * NG + threadIdx.x / LPR;
  if (row >= n_rows) return;
  const unsigned gmask = group_mask(LPR);
  const RowPre<4> pre = row_prefetch<D, 4>(p, row, lig);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int q = 0; q < n_parts; ++q) {
    const float4 v = ldg_f4_stream(partials + ((int64_t)q * part_rows + row) * D + lig * 4);
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
  }
  row_epilogue<D, 4, DST_BF16>(p, pre, row, lig, gmask, s);
}
}  // namespace lgcn

extern "C" int lgcn_reduce_rows(const float* partials, int n_parts, int64_t part_rows, int64_t n_rows, const float* dinv,
                                const lgcn_layer_args_t* a, lgcn_stream_t stream) {
  LGCN_CHECK_ARG(partials != nullptr && dinv != nullptr && a != nullptr, "null pointer argument");
  LGCN_CHECK_ARG(n_parts >= 1 && n_parts <= LGCN_MAX_PEERS && part_rows >= n_rows && n_rows >= 0, "bad shape");
  LGCN_CHECK_ARG(a->acc_out == nullptr || a->acc_in != nullptr, "acc_out needs acc_in");
  LGCN_CHECK_ARG(a->grad_mode >= 0 && a->grad_mode <= 2, "grad_mode must be 0, 1 or 2");
  if (a->grad_mode != 0) {
    LGCN_CHECK_ARG(a->emb != nullptr && a->cnt != nullptr, "grad_mode needs emb and cnt");
    LGCN_CHECK_ARG(a->grad_mode != 1 || a->grad != nullptr, "grad_mode 1 needs grad");
    LGCN_CHECK_ARG(a->grad_mode != 2 || (a->adam_m && a->adam_v && a->adam_hp && a->base),
                   "grad_mode 2 needs adam_m, adam_v, adam_hp and base");
  }
  if (n_rows == 0) return 0;
  lgcn::EpiDev p;
  if (int rc = lgcn::fill_epilogue(a, dinv, p)) return rc;
  p.src = partials;
  cudaStream_t st = (cudaStream_t)stream;
  const bool db = a->dst != nullptr && a->dst_dtype == LGCN_BF16;
#define LGCN_REDUCE(D_)                                                                                     \
  case D_: {                                                                                                \
    constexpr int NG = lgcn::kBlock / (D_ / 4);                                                             \
    const unsigned grid = (unsigned)((n_rows + NG - 1) / NG);                                               \
    if (db) lgcn::reduce_rows_kernel<D_, true><<<grid, lgcn::kBlock, 0, st>>>(partials, n_parts, part_rows, n_rows, p); \
    else lgcn::reduce_rows_kernel<D_, false><<<grid, lgcn::kBlock, 0, st>>>(partials, n_parts, part_rows, n_rows, p);   \
    break;                                                                                                  \
  }
  switch (a->d) {
    LGCN_REDUCE(32)
    LGCN_REDUCE(64)
    LGCN_REDUCE(128)
    default:
      set_last_error("unsupported embedding width d=%d (supported: 32, 64, 128)", a->d);
      return LGCN_ERR_UNSUPPORTED;
  }
#undef LGCN_REDUCE
  LGCN_LAUNCH_OK();
  return 0;
}

