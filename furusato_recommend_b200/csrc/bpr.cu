// Fused BPR loss forward + gradient seed scatter, and the Adam bookkeeping.
//
// Replaces the six row gathers, the ~12 elementwise/reduction launches and the
// autograd index_put scatter of the reference's getEmbedding/bpr_loss/backward
// (model/lgcn.py:88-118,131) with ONE kernel: a group of d/4 lanes owns one
// (user, pos, neg) sample, keeps the three propagated rows in registers, forms
// the two dot products with shuffles, and issues 16-byte vector reds into G.
// The loss / reg sums are reduced in a fixed order by the last CTA to finish,
// so the reported loss is bit-stable run to run.
#include "common.cuh"
#include "loss_common.cuh"

namespace lgcn {

template <int D>
__global__ void __launch_bounds__(kBprBlock)
bpr_kernel(const float* __restrict__ out, const float* __restrict__ emb,
           const int64_t* __restrict__ users, const int64_t* __restrict__ pos,
           const int64_t* __restrict__ neg, int batch, int64_t n_users, int64_t m_items, float decay,
           float loss_scale, float* __restrict__ G, int32_t* __restrict__ cnt,
           float* __restrict__ loss_out, float* __restrict__ work,
           int32_t* __restrict__ work_counter) {
  constexpr int LPR = D / 4;
  constexpr int NG = kBprBlock / LPR;
  const int lig = threadIdx.x % LPR;
  const int grp = threadIdx.x / LPR;
  const unsigned gmask = group_mask(LPR);
  const int b = blockIdx.x * NG + grp;
  const float inv_b = 1.0f / float(batch);

  // An id outside the table is an IndexError in the reference (model/lgcn.py:90-95).  Here it would be
  // an out-of-bounds vector red into G: the sample is skipped, counted in work_counter[1] (sticky;
  // the host raises from it) and the step's loss is poisoned with NaN.
  bool in_range = false;
  if (b < batch) {
    const int64_t uu = users[b], pi = pos[b], qi = neg[b];
    in_range = uu >= 0 && uu < n_users && pi >= 0 && pi < m_items && qi >= 0 && qi < m_items;
    if (!in_range && lig == 0) {
      atomicAdd(work_counter + 1, 1);
      work[b] = __int_as_float(0x7fc00000);
      work[batch + b] = 0.f;
    }
  }
  if (in_range) {
    const int64_t u = users[b];
    const int64_t p = n_users + pos[b];
    const int64_t q = n_users + neg[b];
    const int eo = lig * 4;
    const float4 ue = ld_f4(out + u * D + eo);
    const float4 pe = ld_f4(out + p * D + eo);
    const float4 qe = ld_f4(out + q * D + eo);
    const float4 u0 = ld_f4(emb + u * D + eo);
    const float4 p0 = ld_f4(emb + p * D + eo);
    const float4 q0 = ld_f4(emb + q * D + eo);

    float pos_s = ue.x * pe.x + ue.y * pe.y + ue.z * pe.z + ue.w * pe.w;
    float neg_s = ue.x * qe.x + ue.y * qe.y + ue.z * qe.z + ue.w * qe.w;
    float rg = u0.x * u0.x + u0.y * u0.y + u0.z * u0.z + u0.w * u0.w +
               p0.x * p0.x + p0.y * p0.y + p0.z * p0.z + p0.w * p0.w +
               q0.x * q0.x + q0.y * q0.y + q0.z * q0.z + q0.w * q0.w;
    pos_s = group_sum<LPR>(pos_s, gmask);
    neg_s = group_sum<LPR>(neg_s, gmask);
    rg = group_sum<LPR>(rg, gmask);

    const float x = neg_s - pos_s;
    // torch.nn.functional.softplus: beta=1, threshold=20
    const float sp = x > 20.f ? x : log1pf(expf(x));
    const float sig = 1.f / (1.f + expf(-x));
    const float s = loss_scale * sig * inv_b;

    red_add_f4(G + u * D + eo, s * (qe.x - pe.x), s * (qe.y - pe.y), s * (qe.z - pe.z),
               s * (qe.w - pe.w));
    red_add_f4(G + p * D + eo, -s * ue.x, -s * ue.y, -s * ue.z, -s * ue.w);
    red_add_f4(G + q * D + eo, s * ue.x, s * ue.y, s * ue.z, s * ue.w);
    if (lig == 0) {
      atomicAdd(cnt + u, 1);
      atomicAdd(cnt + p, 1);
      atomicAdd(cnt + q, 1);
      work[b] = sp;
      work[batch + b] = rg;
    }
  }

  bpr_finish(batch, inv_b, decay, loss_out, work, work_counter);
}

// ---------------------------------------------------------------------------------------------
// Row-partitioned variant: the 3B rows of the batch live in a compact table
// rows[3B][2d] = [light_out row | embedding row] that their owners filled over NVLink
// (lgcn_exchange_rows_push); sample b uses rows b (user), B + b (positive), 2B + b (negative).
// Same arithmetic as bpr_kernel; the gradient rows go
//   * into this rank's G / cnt shard, for the rows it owns (padded id / R == rank), and
//   * pre-scaled by deg^-1/2, into row `padded id` of g0_full [W*R, d] fp32 — the layer-0 source
//     of the backward pass, built locally on every rank instead of all-gathering the (almost
//     empty) G shards.
// Replaces, on the distributed step, two strided copies, two memsets, the compact-table BPR
// launch, two index_add_ and a where() of the round-1 torch glue.
template <int D>
__global__ void __launch_bounds__(kBprBlock)
bpr_rows_kernel(const float* __restrict__ rows, const int64_t* __restrict__ padded_ids, int batch,
                int64_t R, int rank, float decay, float loss_scale, float* __restrict__ G,
                int32_t* __restrict__ cnt, float* __restrict__ g0_full, const float* __restrict__ dinv_pad,
                float* __restrict__ loss_out, float* __restrict__ work, int32_t* __restrict__ work_counter) {
  constexpr int LPR = D / 4;
  constexpr int NG = kBprBlock / LPR;
  const int lig = threadIdx.x % LPR;
  const int grp = threadIdx.x / LPR;
  const unsigned gmask = group_mask(LPR);
  const int b = blockIdx.x * NG + grp;
  const float inv_b = 1.0f / float(batch);
  if (b < batch) {
    const int eo = lig * 4;
    const float* ru = rows + (int64_t)b * 2 * D;
    const float* rp = rows + ((int64_t)batch + b) * 2 * D;
    const float* rq = rows + (2 * (int64_t)batch + b) * 2 * D;
    const float4 ue = ld_f4(ru + eo), pe = ld_f4(rp + eo), qe = ld_f4(rq + eo);
    const float4 u0 = ld_f4(ru + D + eo), p0 = ld_f4(rp + D + eo), q0 = ld_f4(rq + D + eo);
    float pos_s = ue.x * pe.x + ue.y * pe.y + ue.z * pe.z + ue.w * pe.w;
    float neg_s = ue.x * qe.x + ue.y * qe.y + ue.z * qe.z + ue.w * qe.w;
    float rg = u0.x * u0.x + u0.y * u0.y + u0.z * u0.z + u0.w * u0.w +
               p0.x * p0.x + p0.y * p0.y + p0.z * p0.z + p0.w * p0.w +
               q0.x * q0.x + q0.y * q0.y + q0.z * q0.z + q0.w * q0.w;
    pos_s = group_sum<LPR>(pos_s, gmask);
    neg_s = group_sum<LPR>(neg_s, gmask);
    rg = group_sum<LPR>(rg, gmask);
    const float x = neg_s - pos_s;
    const float sp = x > 20.f ? x : log1pf(expf(x));
    const float sig = 1.f / (1.f + expf(-x));
    const float s = loss_scale * sig * inv_b;
    const float4 gu = make_float4(s * (qe.x - pe.x), s * (qe.y - pe.y), s * (qe.z - pe.z), s * (qe.w - pe.w));
    const float4 gp = make_float4(-s * ue.x, -s * ue.y, -s * ue.z, -s * ue.w);
    const float4 gq = make_float4(s * ue.x, s * ue.y, s * ue.z, s * ue.w);
    const int64_t id[3] = {__ldg(padded_ids + b), __ldg(padded_ids + batch + b),
                           __ldg(padded_ids + 2 * (int64_t)batch + b)};
    const float4 gr[3] = {gu, gp, gq};
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      if (g0_full != nullptr) {
        const float di = __ldg(dinv_pad + id[t]);
        red_add_f4(g0_full + id[t] * D + eo, di * gr[t].x, di * gr[t].y, di * gr[t].z, di * gr[t].w);
      }
      const int64_t loc = id[t] - (int64_t)rank * R;
      if (loc >= 0 && loc < R) {
        red_add_f4(G + loc * D + eo, gr[t].x, gr[t].y, gr[t].z, gr[t].w);
        if (lig == 0) atomicAdd(cnt + loc, 1);
      }
    }
    if (lig == 0) {
      work[b] = sp;
      work[batch + b] = rg;
    }
  }
  bpr_finish(batch, inv_b, decay, loss_out, work, work_counter);
}

// Global (user, positive, negative) ids -> padded ids of the row partition (parallel.RowPartition):
// padded = owner * R + offset of the side inside the owner's block + (id - first row of the range).
// cuts: [2][W + 1] global row boundaries per side (users, items), side_off: [2][W].
__global__ void __launch_bounds__(256)
padded_ids_kernel(const int64_t* __restrict__ users, const int64_t* __restrict__ pos,
                  const int64_t* __restrict__ neg, int batch, int64_t n_users, int64_t n_nodes,
                  const int64_t* __restrict__ cuts, const int64_t* __restrict__ side_off, int world,
                  int64_t R, int64_t* __restrict__ out, int32_t* __restrict__ status) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * batch) return;
  const int which = t / batch, b = t - which * batch;
  const int64_t raw = which == 0 ? users[b] : (which == 1 ? pos[b] : neg[b]);
  const int64_t lim = which == 0 ? n_users : n_nodes - n_users;
  int64_t g = which == 0 ? raw : n_users + raw;
  if (raw < 0 || raw >= lim) {   // the reference's IndexError: counted, mapped to a valid row
    atomicAdd(status, 1);
    g = which == 0 ? 0 : n_users;
  }
  const int side = which == 0 ? 0 : 1;
  const int64_t* c = cuts + side * (world + 1);
  int r = 0;
  while (r + 1 < world && g >= c[r + 1]) ++r;
  out[t] = (int64_t)r * R + side_off[side * world + r] + (g - c[r]);
}

template <int D4>
__global__ void __launch_bounds__(256)
zero_rows_kernel(float* __restrict__ tab, const int64_t* __restrict__ ids, int64_t n_ids) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_ids * D4) return;
  const int64_t i = t / D4;
  st_f4(tab + __ldg(ids + i) * (4 * D4) + 4 * (t - i * D4), make_float4(0.f, 0.f, 0.f, 0.f));
}

__global__ void adam_tick_kernel(int64_t* step, float* hp, double lr, double b1, double b2) {
  const int64_t t = *step + 1;
  *step = t;
  const double bc1 = 1.0 - pow(b1, (double)t);
  const double bc2 = 1.0 - pow(b2, (double)t);
  hp[0] = (float)(lr / bc1);
  hp[1] = (float)sqrt(bc2);
}

__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                 float* __restrict__ v, int64_t n4, const float* __restrict__ hp, float omb1,
                 float b2, float omb2, float eps) {
  const float step_size = __ldg(hp), bc2_sqrt = __ldg(hp + 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g4 = ld_f4(grad + 4 * i);
    const float4 m4 = ld_f4(m + 4 * i);
    const float4 v4 = ld_f4(v + 4 * i);
    const float4 p4 = ld_f4(param + 4 * i);
    const float g[4] = {g4.x, g4.y, g4.z, g4.w};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w};
    float vv[4] = {v4.x, v4.y, v4.z, v4.w};
    float pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mm[j] = fmaf(g[j] - mm[j], omb1, mm[j]);
      vv[j] = fmaf(omb2 * g[j], g[j], vv[j] * b2);
      const float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
      pp[j] = pp[j] - step_size * (mm[j] / denom);
    }
    st_f4(m + 4 * i, make_float4(mm[0], mm[1], mm[2], mm[3]));
    st_f4(v + 4 * i, make_float4(vv[0], vv[1], vv[2], vv[3]));
    st_f4(param + 4 * i, make_float4(pp[0], pp[1], pp[2], pp[3]));
  }
}

template <int D>
static int launch_bpr(const float* out, const float* emb, const int64_t* users, const int64_t* pos,
                      const int64_t* neg, int batch, int64_t n_users, int64_t m_items, float decay, float loss_scale,
                      float* G, int32_t* cnt, float* loss_out, float* work, int32_t* work_counter,
                      cudaStream_t st) {
  constexpr int NG = kBprBlock / (D / 4);
  const int grid = (batch + NG - 1) / NG;
  bpr_kernel<D><<<grid, kBprBlock, 0, st>>>(out, emb, users, pos, neg, batch, n_users, m_items, decay,
                                            loss_scale, G, cnt, loss_out, work, work_counter);
  LGCN_LAUNCH_OK();
  return 0;
}

}  // namespace lgcn

using namespace lgcn;

extern "C" int lgcn_bpr_fwd_bwd(const float* out, const float* emb, const int64_t* users,
                                const int64_t* pos, const int64_t* neg, int64_t batch,
                                int64_t n_users, int64_t n_nodes, int d, float decay,
                                float loss_scale, float* G, int32_t* cnt, float* loss_out,
                                float* work, int32_t* work_counter, lgcn_stream_t stream) {
  LGCN_CHECK_ARG(out && emb && users && pos && neg && G && cnt && loss_out && work && work_counter,
                 "null pointer argument");
  LGCN_CHECK_ARG(batch > 0 && batch < (1 << 30), "batch out of range: %lld", (long long)batch);
  LGCN_CHECK_ARG(n_users >= 0 && n_users <= n_nodes, "n_users out of range");
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 32: return launch_bpr<32>(out, emb, users, pos, neg, (int)batch, n_users, n_nodes - n_users, decay, loss_scale, G, cnt, loss_out, work, work_counter, st);
    case 64: return launch_bpr<64>(out, emb, users, pos, neg, (int)batch, n_users, n_nodes - n_users, decay, loss_scale, G, cnt, loss_out, work, work_counter, st);
    case 128: return launch_bpr<128>(out, emb, users, pos, neg, (int)batch, n_users, n_nodes - n_users, decay, loss_scale, G, cnt, loss_out, work, work_counter, st);
    default:
      set_last_error("unsupported embedding width d=%d (supported: 32, 64, 128)", d);
      return LGCN_ERR_UNSUPPORTED;
  }
}

extern "C" int lgcn_bpr_fwd_bwd_rows(const float* rows, const int64_t* padded_ids, int64_t batch, int d,
                                     int64_t rows_per_rank, int rank, float decay, float loss_scale, float* G,
                                     int32_t* cnt, float* g0_full, const float* dinv_pad, float* loss_out,
                                     float* work, int32_t* work_counter, lgcn_stream_t stream) {
  LGCN_CHECK_ARG(rows && padded_ids && G && cnt && loss_out && work && work_counter, "null pointer argument");
  LGCN_CHECK_ARG(g0_full == nullptr || dinv_pad != nullptr, "g0_full needs dinv_pad");
  LGCN_CHECK_ARG(batch > 0 && batch < (1 << 30), "batch out of range: %lld", (long long)batch);
  LGCN_CHECK_ARG(rows_per_rank > 0 && rank >= 0, "bad partition");
  cudaStream_t st = (cudaStream_t)stream;
#define LGCN_BPR_ROWS(D_)                                                                                   \
  case D_: {                                                                                                \
    constexpr int NG = kBprBlock / (D_ / 4);                                                                \
    bpr_rows_kernel<D_><<<((int)batch + NG - 1) / NG, kBprBlock, 0, st>>>(                                  \
        rows, padded_ids, (int)batch, rows_per_rank, rank, decay, loss_scale, G, cnt, g0_full, dinv_pad,    \
        loss_out, work, work_counter);                                                                      \
    break;                                                                                                  \
  }
  switch (d) {
    LGCN_BPR_ROWS(32)
    LGCN_BPR_ROWS(64)
    LGCN_BPR_ROWS(128)
    default:
      set_last_error("unsupported embedding width d=%d (supported: 32, 64, 128)", d);
      return LGCN_ERR_UNSUPPORTED;
  }
#undef LGCN_BPR_ROWS
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_padded_ids(const int64_t* users, const int64_t* pos, const int64_t* neg, int64_t batch,
                               int64_t n_users, int64_t n_nodes, const int64_t* cuts, const int64_t* side_off,
                               int world, int64_t rows_per_rank, int64_t* padded_ids, int32_t* status,
                               lgcn_stream_t stream) {
  LGCN_CHECK_ARG(users && pos && neg && cuts && side_off && padded_ids && status, "null pointer argument");
  LGCN_CHECK_ARG(batch > 0 && batch < (1 << 29) && world >= 1 && world <= LGCN_MAX_PEERS, "bad shape");
  const int n = 3 * (int)batch;
  padded_ids_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      users, pos, neg, (int)batch, n_users, n_nodes, cuts, side_off, world, rows_per_rank, padded_ids, status);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_zero_rows(float* table, int d, const int64_t* ids, int64_t n_ids, lgcn_stream_t stream) {
  LGCN_CHECK_ARG(table && ids && n_ids >= 0, "null pointer argument");
  if (n_ids == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 32: zero_rows_kernel<8><<<(unsigned)((n_ids * 8 + 255) / 256), 256, 0, st>>>(table, ids, n_ids); break;
    case 64: zero_rows_kernel<16><<<(unsigned)((n_ids * 16 + 255) / 256), 256, 0, st>>>(table, ids, n_ids); break;
    case 128: zero_rows_kernel<32><<<(unsigned)((n_ids * 32 + 255) / 256), 256, 0, st>>>(table, ids, n_ids); break;
    default:
      set_last_error("unsupported embedding width d=%d (supported: 32, 64, 128)", d);
      return LGCN_ERR_UNSUPPORTED;
  }
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_adam_tick(int64_t* step, float* adam_hp, double lr, double beta1, double beta2,
                              lgcn_stream_t stream) {
  LGCN_CHECK_ARG(step && adam_hp, "null pointer argument");
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, adam_hp, lr, beta1, beta2);
  LGCN_LAUNCH_OK();
  return 0;
}

extern "C" int lgcn_adam_step(float* param, const float* grad, float* m, float* v, int64_t numel,
                              const float* adam_hp, double beta1, double beta2, double eps,
                              lgcn_stream_t stream) {
  LGCN_CHECK_ARG(param && grad && m && v && adam_hp, "null pointer argument");
  LGCN_CHECK_ARG(numel >= 0 && numel % 4 == 0, "numel must be a multiple of 4");
  if (numel == 0) return 0;
  const int64_t n4 = numel / 4;
  int64_t blocks = (n4 + 255) / 256;
  const int64_t cap = (int64_t)kSmCount * 16;
  if (blocks > cap) blocks = cap;
  adam_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      param, grad, m, v, n4, adam_hp, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
      (float)eps);
  LGCN_LAUNCH_OK();
  return 0;
}
