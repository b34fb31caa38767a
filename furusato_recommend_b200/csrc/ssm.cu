// Sampled-softmax loss forward + gradient seed (BASELINE configs[3]).
//
// The reference's model/lgcnssm.py names its loss `softmax_loss` (lgcnssm.py:98-118) but the body is
// byte-for-byte the BPR softplus loss, and its OneEpoch raises NameError (lgcnssm.py:141): there is
// NO reference arithmetic for a sampled softmax — PARITY UNPINNED.  What the file keeps is the batch
// layout: flat (user, pos, neg) triples, J = neg_size consecutive rows per (user, positive).  This
// kernel computes the objective that layout is for (our own specification):
//
//   z_0 = <u_b, p_b> / tau,  z_j = <u_b, q_bj> / tau  (j = 1..J, rows of light_out)
//   loss = mean_b [ logsumexp(z_0..z_J) - z_0 ]
//   reg  = 0.5 * sum_b (|E[u_b]|^2 + |E[n+pos_b]|^2 + sum_j |E[n+neg_bj]|^2) / B
//   with w = softmax(z):   G[u_b]      += c * (sum_k w_k row_k - p_b)
//                          G[n+pos_b]  += c * (w_0 - 1) * u_b
//                          G[n+neg_bj] += c * w_j * u_b,         c = loss_scale / (tau * B)
//   cnt[row] += 1 per occurrence (the L2 term is applied by the propagate epilogue, as for BPR).
//
// One group of d/4 lanes owns one (user, positive): the user row stays in registers, the J negative
// rows stream through ONCE with a flash-style running (max, sum, weighted row sum), their logits are
// parked in shared memory, and a second loop — no gathers — scatters w_j * u_b with 16-byte vector
// reds.  The loss / reg sums are reduced in a fixed order by the last CTA (bpr_finish).
#include "common.cuh"
#include "loss_common.cuh"

namespace lgcn {

template <int D>
__global__ void __launch_bounds__(kBprBlock)
ssm_kernel(const float* __restrict__ out, const float* __restrict__ emb, const int64_t* __restrict__ users,
           const int64_t* __restrict__ pos, const int64_t* __restrict__ neg, int batch, int J, int64_t n_users,
           int64_t m_items, float inv_tau, float decay, float loss_scale, float* __restrict__ G,
           int32_t* __restrict__ cnt, float* __restrict__ loss_out, float* __restrict__ work,
           int32_t* __restrict__ work_counter) {
  constexpr int LPR = D / 4;
  constexpr int NG = kBprBlock / LPR;
  extern __shared__ float s_z[];   // [NG][J] logits of the group's negatives
  const int lig = threadIdx.x % LPR;
  const int grp = threadIdx.x / LPR;
  const unsigned gmask = group_mask(LPR);
  const int b = blockIdx.x * NG + grp;
  const float inv_b = 1.0f / float(batch);
  float* zs = s_z + (size_t)grp * J;

  bool ok = false;
  if (b < batch) {
    const int64_t uu = users[(int64_t)b * J], pi = pos[(int64_t)b * J];
    ok = uu >= 0 && uu < n_users && pi >= 0 && pi < m_items;
    for (int j = lig; j < J && ok; j += LPR) {
      const int64_t qi = neg[(int64_t)b * J + j];
      ok = qi >= 0 && qi < m_items;
    }
    ok = __all_sync(gmask, ok);
    if (!ok && lig == 0) {   // the reference's IndexError: skipped, counted, loss poisoned
      atomicAdd(work_counter + 1, 1);
      work[b] = __int_as_float(0x7fc00000);
      work[batch + b] = 0.f;
    }
  }
  if (ok) {
    const int eo = lig * 4;
    const int64_t u = users[(int64_t)b * J];
    const int64_t p = n_users + pos[(int64_t)b * J];
    const float4 ue = ld_f4(out + u * D + eo);
    const float4 pe = ld_f4(out + p * D + eo);
    const float4 u0 = ld_f4(emb + u * D + eo);
    const float4 p0 = ld_f4(emb + p * D + eo);
    float rg = u0.x * u0.x + u0.y * u0.y + u0.z * u0.z + u0.w * u0.w +
               p0.x * p0.x + p0.y * p0.y + p0.z * p0.z + p0.w * p0.w;
    const float z0 = group_sum<LPR>(ue.x * pe.x + ue.y * pe.y + ue.z * pe.z + ue.w * pe.w, gmask) * inv_tau;
    // running softmax state over k = 0..J (k = 0 is the positive)
    float m = z0, l = 1.f;
    float4 acc = pe;
    const int64_t* nb = neg + (int64_t)b * J;
    for (int j = 0; j < J; ++j) {
      const int64_t q = n_users + __ldg(nb + j);
      const float4 qe = ld_f4(out + q * D + eo);
      const float4 q0 = ld_f4(emb + q * D + eo);
      rg += q0.x * q0.x + q0.y * q0.y + q0.z * q0.z + q0.w * q0.w;
      const float z = group_sum<LPR>(ue.x * qe.x + ue.y * qe.y + ue.z * qe.z + ue.w * qe.w, gmask) * inv_tau;
      if (lig == 0) zs[j] = z;
      const float mn = fmaxf(m, z);
      const float sc = expf(m - mn), w = expf(z - mn);
      acc.x = acc.x * sc + w * qe.x; acc.y = acc.y * sc + w * qe.y;
      acc.z = acc.z * sc + w * qe.z; acc.w = acc.w * sc + w * qe.w;
      l = l * sc + w;
      m = mn;
    }
    rg = group_sum<LPR>(rg, gmask);
    const float inv_l = 1.f / l;
    const float c = loss_scale * inv_tau * inv_b;
    red_add_f4(G + u * D + eo, c * (acc.x * inv_l - pe.x), c * (acc.y * inv_l - pe.y), c * (acc.z * inv_l - pe.z),
               c * (acc.w * inv_l - pe.w));
    const float w0 = c * (expf(z0 - m) * inv_l - 1.f);
    red_add_f4(G + p * D + eo, w0 * ue.x, w0 * ue.y, w0 * ue.z, w0 * ue.w);
    __syncwarp(gmask);   // zs[] written by lane 0 of the group
    for (int j = 0; j < J; ++j) {
      const int64_t q = n_users + __ldg(nb + j);
      const float wj = c * expf(zs[j] - m) * inv_l;
      red_add_f4(G + q * D + eo, wj * ue.x, wj * ue.y, wj * ue.z, wj * ue.w);
      if (lig == 0) atomicAdd(cnt + q, 1);
    }
    if (lig == 0) {
      atomicAdd(cnt + u, 1);
      atomicAdd(cnt + p, 1);
      work[b] = logf(l) + m - z0;
      work[batch + b] = rg;
    }
  }
  bpr_finish(batch, inv_b, decay, loss_out, work, work_counter);
}

}  // namespace lgcn

using namespace lgcn;

extern "C" int lgcn_ssm_fwd_bwd(const float* out, const float* emb, const int64_t* users, const int64_t* pos,
                                const int64_t* neg, int64_t batch, int n_neg, int64_t n_users, int64_t n_nodes, int d,
                                float tau, float decay, float loss_scale, float* G, int32_t* cnt, float* loss_out,
                                float* work, int32_t* work_counter, lgcn_stream_t stream) {
  LGCN_CHECK_ARG(out && emb && users && pos && neg && G && cnt && loss_out && work && work_counter,
                 "null pointer argument");
  LGCN_CHECK_ARG(batch > 0 && batch < (1 << 30), "batch out of range: %lld", (long long)batch);
  LGCN_CHECK_ARG(n_neg >= 1 && n_neg <= 2048, "n_neg must be in [1, 2048]");
  LGCN_CHECK_ARG(n_users >= 0 && n_users <= n_nodes, "n_users out of range");
  LGCN_CHECK_ARG(tau > 0.f, "tau must be positive");
  cudaStream_t st = (cudaStream_t)stream;
#define LGCN_SSM(D_)                                                                                           \
  case D_: {                                                                                                   \
    constexpr int NG = kBprBlock / (D_ / 4);                                                                   \
    const size_t smem = (size_t)NG * n_neg * sizeof(float);                                                    \
    LGCN_CUDA_OK(cudaFuncSetAttribute(ssm_kernel<D_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    ssm_kernel<D_><<<((int)batch + NG - 1) / NG, kBprBlock, smem, st>>>(                                       \
        out, emb, users, pos, neg, (int)batch, n_neg, n_users, n_nodes - n_users, 1.f / tau, decay, loss_scale, \
        G, cnt, loss_out, work, work_counter);                                                                 \
    break;                                                                                                     \
  }
  switch (d) {
    LGCN_SSM(32)
    LGCN_SSM(64)
    LGCN_SSM(128)
    default:
      set_last_error("unsupported embedding width d=%d (supported: 32, 64, 128)", d);
      return LGCN_ERR_UNSUPPORTED;
  }
#undef LGCN_SSM
  LGCN_LAUNCH_OK();
  return 0;
}
