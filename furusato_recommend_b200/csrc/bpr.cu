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

namespace lgcn {

constexpr int kBprBlock = 256;

template <int D>
__global__ void __launch_bounds__(kBprBlock)
bpr_kernel(const float* __restrict__ out, const float* __restrict__ emb,
           const int64_t* __restrict__ users, const int64_t* __restrict__ pos,
           const int64_t* __restrict__ neg, int batch, int64_t n_users, float decay,
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

  if (b < batch) {
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

  // ---- deterministic final reduction by the last CTA ----
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
                      const int64_t* neg, int batch, int64_t n_users, float decay, float loss_scale,
                      float* G, int32_t* cnt, float* loss_out, float* work, int32_t* work_counter,
                      cudaStream_t st) {
  constexpr int NG = kBprBlock / (D / 4);
  const int grid = (batch + NG - 1) / NG;
  bpr_kernel<D><<<grid, kBprBlock, 0, st>>>(out, emb, users, pos, neg, batch, n_users, decay,
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
    case 32: return launch_bpr<32>(out, emb, users, pos, neg, (int)batch, n_users, decay, loss_scale, G, cnt, loss_out, work, work_counter, st);
    case 64: return launch_bpr<64>(out, emb, users, pos, neg, (int)batch, n_users, decay, loss_scale, G, cnt, loss_out, work, work_counter, st);
    case 128: return launch_bpr<128>(out, emb, users, pos, neg, (int)batch, n_users, decay, loss_scale, G, cnt, loss_out, work, work_counter, st);
    default:
      set_last_error("unsupported embedding width d=%d (supported: 32, 64, 128)", d);
      return LGCN_ERR_UNSUPPORTED;
  }
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
