/* lgcn_b200.h — C ABI of the B200-native LightGCN hot path (liblgcn_b200.so).
 *
 * The reference (HiromasaYamanishi/furusato_recommend) is pure Python and has no
 * FFI boundary: its seam is the duck-typed model/sampler/dataset API.  Each entry
 * point below replaces the torch / PyG / numpy call sites named in its comment
 * (paths relative to the reference root).  The Python host layer in
 * furusato_recommend_b200/ binds these with ctypes and keeps the reference's
 * method names (LightGCN.computer/forward/bpr_loss/getUsersRating/stageOne,
 * UniformSample, Loader.getSparseGraph).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless it says HOST;
 *     outputs are pre-allocated by the caller; no allocation happens inside;
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*) and can be
 *     captured into a CUDA graph.  The library keeps no state of its own, but some ARGUMENTS are
 *     mutable scratch the caller owns — `hub_counter` / `partial` inside lgcn_graph_t, `work` /
 *     `work_counter` of the loss kernels: launches that share them must be stream-ordered (two
 *     streams on ONE lgcn_graph_t race; give each stream its own scratch);
 *   - return value: 0 = ok, > 0 = a cudaError_t, < 0 = LGCN_ERR_*; the message is
 *     available from lgcn_last_error() (thread-local).  Nothing throws.
 *   - embedding rows are row-major [N, d] with users first, then items
 *     (model/lgcn.py:71-74), 16-byte aligned, d in {32, 64, 128} (bf16: also 256).
 */
#ifndef LGCN_B200_H
#define LGCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGCN_ABI_VERSION 4

#define LGCN_ERR_INVALID_ARG (-1)
#define LGCN_ERR_UNSUPPORTED (-2)

#define LGCN_F32 0
#define LGCN_BF16 1
#define LGCN_F16 2 /* score_topk only: f16 operands AND f16 tensor-core accumulators */

/* rows with more than LGCN_HUB_DEG edges are split into CTA-wide segments of at
 * most LGCN_SEG_EDGES edges (host side builds the lists, see graph.py) */
#define LGCN_HUB_DEG 256
#define LGCN_SEG_EDGES 1024
#define LGCN_MAX_PEERS 8
/* the sampler gives up on a negative after this many rejected candidates in a row (the sample is
 * then dropped like an empty user) instead of spinning forever like negative_sample.py:121-126 */
#define LGCN_MAX_NEG_TRIES 256

typedef void* lgcn_stream_t; /* cudaStream_t */

int lgcn_abi_version(void);
const char* lgcn_last_error(void);

/* ------------------------------------------------------------------------
 * Graph: CSR of the bipartite adjacency A = [[0,R],[R^T,0]] WITHOUT values.
 * Replaces the COO FloatTensor of Loader.getSparseGraph (dataloader.py:215-258):
 * the normalisation D^-1/2 A D^-1/2 is folded into `dinv` (dataloader.py:236-238),
 * multi-edges are repeated column entries (csr_matrix sums duplicates, :164).
 * ------------------------------------------------------------------------ */
typedef struct lgcn_graph {
  int64_t n_nodes;           /* N = n_users + m_items                          */
  int64_t nnz;               /* directed entries (= 2 * train interactions)    */
  const int64_t* rowptr;     /* [N+1]                                          */
  const int32_t* col;        /* [nnz] neighbour node ids, sorted inside a row  */
  const float* dinv;         /* [N] deg^-1/2 (fp32), 0 where deg == 0          */
  /* work decomposition (degree-descending so long rows start first) */
  const int32_t* light_desc; /* [n_light][4] = {row, degree, first edge lo, hi}  */
                             /* of the rows with deg <= LGCN_HUB_DEG, 16 B aligned */
  int64_t n_light;
  const int32_t* seg_row;    /* [n_seg] hub row of each CTA segment            */
  const int64_t* seg_begin;  /* [n_seg] first edge                             */
  const int32_t* seg_len;    /* [n_seg] edges in the segment (<= SEG_EDGES)    */
  const int32_t* seg_hub;    /* [n_seg] index into the hub arrays              */
  int64_t n_seg;
  const int32_t* hub_seg0;   /* [n_hub] first segment of the hub               */
  const int32_t* hub_nseg;   /* [n_hub] number of segments                     */
  int32_t* hub_counter;      /* [n_hub] zero-initialised; self-resetting       */
  int64_t n_hub;
  float* partial;            /* [n_seg * d] fp32 scratch for multi-segment hubs */
} lgcn_graph_t;

/* ------------------------------------------------------------------------
 * One propagation layer  s_i = sum_{j in N(i)} w_j * SRC[j]  fused with a row
 * epilogue.  Replaces torch.sparse.mm(G, all_emb) (model/MF.py:200,204) ==
 * LGConv()(x, edge_index) (model/lgcn.py:82), the running layer sum
 * (model/lgcn.py:83-84, model/MF.py:206-208), their autograd transposes
 * (A_hat is symmetric, so backward is the same kernel in Horner form) and,
 * optionally, optim.Adam.step() (model/lgcn.py:63,132).
 *
 *   w_j = dinv[j] if scale_src else 1 (then SRC must already hold dinv (.) X)
 *   x_i = dinv[i] * s_i                       (= (A_hat X)[i])
 *   t_i = base ? base[i] + x_i : x_i          (backward: base = G)
 *   if dst:      dst[i]     = dinv[i] * t_i   (pre-scaled source of the next layer)
 *   if acc_out:  acc_out[i] = (acc_in[i] + x_i) * acc_scale   (running layer sum;
 *                last layer: acc_scale = 1/(K+1) gives light_out)
 *   if grad_mode: g_i = t_i * inv_layers + reg_coef * cnt[i] * emb[i]
 *        1: grad[i] = g_i
 *        2: Adam(emb[i], m[i], v[i], g_i) with the step sizes in adam_hp, then
 *           cnt[i] = 0 and, if zero_base, base[i] = 0 (so the next step starts
 *           clean; zero_base is illegal when src == base, i.e. K == 1)
 * ------------------------------------------------------------------------ */
/* SRC contract: row 0 of `src` must hold finite values — gather slots past the end of a
 * neighbour list re-read it (an L1 hit, no predicate on the load) and add it with weight 0. */
typedef struct lgcn_layer_args {
  int d;               /* embedding width                                       */
  int src_dtype;       /* LGCN_F32 | LGCN_BF16                                  */
  int dst_dtype;       /* LGCN_F32 | LGCN_BF16                                  */
  int scale_src;       /* 1: multiply neighbours by dinv[j] on the fly (f32 src) */
  const void* src;     /* [N,d]                                                 */
  void* dst;           /* [N,d] or NULL                                         */
  const float* base;   /* [N,d] fp32 or NULL                                    */
  const float* acc_in; /* [N,d] fp32 or NULL                                    */
  float* acc_out;      /* [N,d] fp32 or NULL (may alias acc_in)                 */
  float acc_scale;
  int grad_mode;       /* 0 | 1 | 2                                             */
  float inv_layers;    /* 1/(K+1)                                               */
  float reg_coef;      /* decay / B                                             */
  int32_t* cnt;        /* [N] occurrences of row i in the batch (bpr kernel)    */
  float* emb;          /* [N,d] fp32 embedding table E                          */
  float* grad;         /* [N,d] fp32 (grad_mode 1)                              */
  float* adam_m;       /* [N,d] fp32 (grad_mode 2)                              */
  float* adam_v;       /* [N,d] fp32 (grad_mode 2)                              */
  const float* adam_hp;/* device float[2]: {lr/bc1, sqrt(bc2)} (lgcn_adam_tick)  */
  double beta1, beta2, eps;
  int zero_base;       /* grad_mode 2: clear base[i] after use                  */
  /* fused all-gather: when n_dst_peers > 0 the dst row is stored to EVERY peer buffer at
   * row (dst_row_offset + i) instead of dst[i] (dst must still be non-NULL to enable the
   * write).  The pointers are peer-mapped device memory (NVLink P2P / symmetric memory). */
  int n_dst_peers;
  int64_t dst_row_offset;
  void* dst_peers[LGCN_MAX_PEERS];
  /* Asymmetric normalisation (rAdjGCN, model/radj.py:32-45: edge weight deg_src^-r * deg_dst^-(1-r)).
   * NULL = the graph's dinv.  With both set the layer computes
   *   w_j = src_scale[j],  x_i = dst_scale[i] * s_i,  dst[i] = src_scale[i] * t_i,
   * i.e. diag(dst_scale) A diag(src_scale); the transpose for the backward pass is the same
   * call with the two vectors swapped (A itself is symmetric). */
  const float* src_scale; /* [N] fp32 or NULL */
  const float* dst_scale; /* [N] fp32 or NULL */
  /* Per-slot weights, aligned with the graph's `col` array, or NULL: s_i = sum_e edge_w[e] * w_j *
   * SRC[col[e]].  Edge dropout (model/MF.py:158-176: every coalesced entry of A_hat is kept with
   * probability keep_prob and rescaled by 1/keep_prob, independently per direction) passes
   * mask/keep_prob here; the backward pass passes the weights of the REVERSE entries. */
  const float* edge_w;
  /* grad_mode 2 only: dst[i] (or the peers' rows, with n_dst_peers) receives src_scale[i] * E_new[i]
   * — the pre-scaled layer-0 source of the NEXT step's forward pass — instead of the pre-scaled
   * t_i.  The first layer of the next propagation then gathers it with scale_src = 0 (and, in the
   * row-partitioned case, needs no separate exchange of the updated table). */
  int push_emb;
  /* NVSwitch multicast (NVLS) form of the fused all-gather: when non-NULL the dst row is written ONCE, with
   * multimem.st, to row (dst_row_offset + i) of this multicast mapping — the switch replicates it into the
   * gathered buffer of every rank of the mapping — instead of n_dst_peers unicast peer stores (the
   * per-rank NVLink egress drops from n_dst_peers rows to one).  dst must still be non-NULL. */
  void* dst_multicast;
  /* > 0 (with dst_peers): every dst row has ONE destination — row i is stored to peer i / dst_route_rows only,
   * at row dst_row_offset + i % dst_route_rows.  Used for the partial sums of the reduce partition
   * (lgcn_reduce_rows): the rows of this launch are all items, in blocks of dst_route_rows per owner. */
  int64_t dst_route_rows;
} lgcn_layer_args_t;

int lgcn_propagate_layer(const lgcn_graph_t* g /*HOST*/, const lgcn_layer_args_t* a /*HOST*/,
                         lgcn_stream_t stream);

/* Owner-side half of the "reduce" partition (bipartite graphs with many more users than items): instead of
 * all-gathering the user rows so that item owners can gather them, every rank sums ITS users' rows per item
 * (lgcn_propagate_layer on the item x local-user CSR with unit scales and dst_route_rows: raw partial sums pushed
 * to the item's owner) and the owner adds the n_parts partials and runs the usual row epilogue:
 *   s_i = sum_q partials[q * part_rows + i]   (q ascending: deterministic),  i < n_rows
 * then exactly lgcn_propagate_layer's epilogue with x_i = dst_scale_i * s_i (a->src is ignored; dinv is the
 * owner's deg^-1/2 vector for these rows).  The per-layer exchange shrinks from the whole table to the item
 * table plus one item-table-sized partial per rank. */
int lgcn_reduce_rows(const float* partials, int n_parts, int64_t part_rows, int64_t n_rows, const float* dinv,
                     const lgcn_layer_args_t* a /*HOST*/, lgcn_stream_t stream);

/* dst_p[row_offset + i] = dinv[i] * x[i] on every peer p: the pre-scaled source of the first
 * layer, pushed straight into the peers' gathered buffers (no separate all-gather). */
int lgcn_scale_rows_push(const float* x, const float* dinv, int64_t n_rows, int d, int dst_dtype,
                         void* const* dst_peers /*HOST array*/, int n_dst_peers,
                         int64_t dst_row_offset, lgcn_stream_t stream);

/* Row exchange of the partitioned BPR step (the getEmbedding gathers of model/lgcn.py:88-96 when
 * the table is row-partitioned): for every i with padded_ids[i] / rows_per_rank == rank, the row
 * [tab_a[loc] | tab_b[loc]] (loc = padded_ids[i] % rows_per_rank; tab_b may be NULL) is stored to
 * row i of EVERY peer buffer ([n_ids, 2d] or [n_ids, d] fp32, peer-mapped).  Each id has one owner,
 * so after a cross-GPU barrier all ranks hold all rows; no reduction, no NCCL call. */
int lgcn_exchange_rows_push(const float* tab_a, const float* tab_b, int d, const int64_t* padded_ids,
                            int64_t n_ids, int64_t rows_per_rank, int rank,
                            void* const* dst_peers /*HOST array*/, int n_dst_peers, lgcn_stream_t stream);

/* ------------------------------------------------------------------------
 * Fused BPR forward + backward seed.  Replaces getEmbedding + bpr_loss
 * (model/lgcn.py:88-118) and the autograd scatter of loss.backward()
 * (model/lgcn.py:131):
 *   x_b = <out[u_b], out[n+neg_b]> - <out[u_b], out[n+pos_b]>
 *   loss_out[0] = mean_b softplus(x_b)  (torch threshold 20)
 *   loss_out[1] = 0.5 * sum_b(|E[u_b]|^2+|E[n+pos_b]|^2+|E[n+neg_b]|^2) / B
 *   loss_out[2] = loss_out[0] + decay * loss_out[1]
 *   loss_out[3] += loss_out[2]          (running epoch sum, model/lgcn.py:149)
 *   G[u_b] += s_b (out[n+neg_b] - out[n+pos_b]);  G[n+pos_b] -= s_b out[u_b];
 *   G[n+neg_b] += s_b out[u_b],  s_b = loss_scale * sigmoid(x_b) / B
 *   cnt[row]  += 1 for each of the 3 rows of every sample
 * G and cnt must be zero on entry.  work: float[2*B]; work_counter: int32[2], zeroed once by the
 * caller: [0] CTA arrival counter (self-resetting), [1] number of samples SKIPPED because an id
 * was outside [0, n_users) / [0, n_nodes - n_users) — the reference's IndexError; sticky, the
 * caller reads and clears it.  A step with a skipped sample reports NaN in loss_out[0..3].
 * ------------------------------------------------------------------------ */
int lgcn_bpr_fwd_bwd(const float* out, const float* emb, const int64_t* users, const int64_t* pos,
                     const int64_t* neg, int64_t batch, int64_t n_users, int64_t n_nodes, int d,
                     float decay, float loss_scale, float* G, int32_t* cnt, float* loss_out,
                     float* work, int32_t* work_counter, lgcn_stream_t stream);

/* ------------------------------------------------------------------------
 * Row-partitioned BPR step (the reference's ddp_lgcn.py:476-515 trains replicas, the
 * partition is _split_A_hat, dataloader.py:195-205, promoted to ranks).
 *
 * lgcn_padded_ids: global (user, pos, neg) ids -> padded row ids of the partition, ids[3B] =
 *   [users | n+pos | n+neg]: padded = owner * rows_per_rank + side_off[side][owner] + (id - cuts[side][owner]).
 *   cuts int64[2][world+1], side_off int64[2][world] (device).  Out-of-range ids are counted in
 *   *status (the reference's IndexError) and mapped to a valid row.
 * lgcn_bpr_fwd_bwd_rows: lgcn_bpr_fwd_bwd on the compact table rows[3B][2d] = [light_out | E] that
 *   lgcn_exchange_rows_push filled (sample b = rows b, B+b, 2B+b).  Gradient rows are added to this
 *   rank's G / cnt shard where padded / rows_per_rank == rank and, scaled by dinv_pad[padded], to
 *   row `padded` of g0_full [world*rows_per_rank, d] fp32 (may be NULL) — the pre-scaled layer-0
 *   source of the backward pass, built locally instead of exchanging the G shards.
 * lgcn_zero_rows: table[ids[i]] = 0 (clears exactly the rows the step touched in g0_full).
 * ------------------------------------------------------------------------ */
int lgcn_padded_ids(const int64_t* users, const int64_t* pos, const int64_t* neg, int64_t batch,
                    int64_t n_users, int64_t n_nodes, const int64_t* cuts, const int64_t* side_off,
                    int world, int64_t rows_per_rank, int64_t* padded_ids, int32_t* status,
                    lgcn_stream_t stream);
int lgcn_bpr_fwd_bwd_rows(const float* rows, const int64_t* padded_ids, int64_t batch, int d,
                          int64_t rows_per_rank, int rank, float decay, float loss_scale, float* G,
                          int32_t* cnt, float* g0_full, const float* dinv_pad, float* loss_out,
                          float* work, int32_t* work_counter, lgcn_stream_t stream);
int lgcn_zero_rows(float* table, int d, const int64_t* ids, int64_t n_ids, lgcn_stream_t stream);

/* ------------------------------------------------------------------------
 * Sampled-softmax forward + backward seed (BASELINE configs[3]).  PARITY UNPINNED: the
 * reference's model/lgcnssm.py:98-118 `softmax_loss` is the BPR softplus loss and its OneEpoch raises
 * NameError (:141); only the batch layout is the reference's — flat triples, n_neg consecutive rows
 * per (user, positive): users[b*n_neg], pos[b*n_neg], neg[b*n_neg + j].  Specification (ours):
 *   z_0 = <out[u], out[n+pos]>/tau, z_j = <out[u], out[n+neg_j]>/tau,  w = softmax(z_0..z_J)
 *   loss_out[0] = mean_b[logsumexp(z) - z_0]
 *   loss_out[1] = 0.5 * sum_b(|E[u]|^2 + |E[n+pos]|^2 + sum_j |E[n+neg_j]|^2) / B
 *   loss_out[2] = [0] + decay * [1];  loss_out[3] += [2]
 *   G[u] += c (sum_k w_k row_k - out[n+pos]); G[n+pos] += c (w_0 - 1) out[u]; G[n+neg_j] += c w_j out[u],
 *   c = loss_scale / (tau * B); cnt[row] += 1 per occurrence.
 * batch = number of (user, positive) pairs B.  work: float[2*B]; work_counter as lgcn_bpr_fwd_bwd.
 * ------------------------------------------------------------------------ */
int lgcn_ssm_fwd_bwd(const float* out, const float* emb, const int64_t* users, const int64_t* pos,
                     const int64_t* neg, int64_t batch, int n_neg, int64_t n_users, int64_t n_nodes, int d,
                     float tau, float decay, float loss_scale, float* G, int32_t* cnt, float* loss_out,
                     float* work, int32_t* work_counter, lgcn_stream_t stream);

/* Adam bookkeeping: ++(*step) and adam_hp = {lr / (1-beta1^t), sqrt(1-beta2^t)}
 * in double precision like torch.optim.Adam's Python scalars. */
int lgcn_adam_tick(int64_t* step, float* adam_hp, double lr, double beta1, double beta2,
                   lgcn_stream_t stream);

/* Dense Adam on a [numel] fp32 tensor given its gradient — optim.Adam.step()
 * (model/lgcn.py:132) for callers that bring their own gradient (autograd path). */
int lgcn_adam_step(float* param, const float* grad, float* m, float* v, int64_t numel,
                   const float* adam_hp, double beta1, double beta2, double eps,
                   lgcn_stream_t stream);

/* ------------------------------------------------------------------------
 * Uniform negative sampler.  Replaces negative_sample.UniformSample
 * (negative_sample.py:98-134).  Sample i draws word j%4 of
 * Philox4x32-10(key=(seed_lo,seed_hi), ctr=(i_lo,i_hi,j/4,epoch)); draw 0 picks
 * the user (mulhi32(r,n_users)), draw 1 the positive from the FILE-ORDER list
 * (:119-120), draws 2.. the first item not contained in the user's positives
 * (:121-126, membership by binary search in the sorted copy; after LGCN_MAX_NEG_TRIES
 * rejections in a row the sample is dropped, valid[i]=0).  Users with an
 * empty list get valid[i]=0 (:116-117).  Samples [first, first+count) are drawn
 * (a shard is a counter offset).  n_neg > 1 draws n_neg negatives per (user, positive) from the
 * following Philox words and emits n_neg flat rows (u, pos, neg_t) per sample — the batch layout
 * of model/lgcnssm.py:141.  triples: int64[count*n_neg, 3], valid: uint8[count*n_neg].
 * ------------------------------------------------------------------------ */
int lgcn_uniform_sample(const int64_t* pos_rowptr, const int32_t* pos_file, const int32_t* pos_sorted,
                        int64_t n_users, int64_t m_items, int64_t first, int64_t count,
                        int n_neg, uint64_t seed, uint32_t epoch, int64_t* triples,
                        uint8_t* valid, lgcn_stream_t stream);

/* Same sampler with a weighted positive pick — UniformSampling.sample_parallel with
 * config['sample_pow'] != 0 (negative_sample.py:53-56: np.random.choice(len(pos), p=probs[user])).
 * pos_cdf: fp32[nnz_pos], aligned with pos_file: per user the inclusive cumulative sum of its
 * probabilities, normalised so the last entry is 1.  The pick is the first j with cdf[j] > u,
 * u = (word >> 8) * 2^-24 of the sample's second Philox word (numpy: searchsorted(cdf, u, 'right')).
 * pos_cdf == NULL is lgcn_uniform_sample. */
int lgcn_uniform_sample_weighted(const int64_t* pos_rowptr, const int32_t* pos_file, const int32_t* pos_sorted,
                                 const float* pos_cdf, int64_t n_users, int64_t m_items, int64_t first,
                                 int64_t count, int n_neg, uint64_t seed, uint32_t epoch, int64_t* triples,
                                 uint8_t* valid, lgcn_stream_t stream);

/* Order-preserving compaction of the valid triples (np.array(S), :134).
 * scratch: int64[ceil(count/1024) + 1]; n_out: device int64[1]. */
int lgcn_compact_triples(const int64_t* triples, const uint8_t* valid, int64_t count,
                         int64_t* out, int64_t* n_out, int64_t* scratch, lgcn_stream_t stream);

/* ------------------------------------------------------------------------
 * Full-rank scoring fused with train-positive masking and top-k.  Replaces
 * getUsersRating's matmul (model/lgcn.py:124), the exclude-list index_put of
 * -(1<<10) (trainer.py:132-137) and torch.topk (trainer.py:138); ties broken by
 * lowest item id.  The [U,m] score matrix is never written.
 *   user_emb/item_emb: propagated embeddings (light_out halves), fp32 [*,d]
 *   user_ids: int64[n_eval] rows of user_emb to score
 *   precision: LGCN_F32 exact fp32 FMA scores | LGCN_BF16 tcgen05 tensor cores (fp32
 *   accumulate) | LGCN_F16 tcgen05 with f16 operands and f16 accumulators (half the TMEM read-out)
 * out_idx int32[n_eval,k] (sorted by score desc, id asc), out_val fp32[n_eval,k].
 * workspace: lgcn_score_topk_workspace_bytes() bytes of device scratch (may be NULL for F32).
 * ------------------------------------------------------------------------ */
int lgcn_score_topk(const float* user_emb, const float* item_emb, const int64_t* user_ids,
                    int64_t n_eval, int64_t m_items, int d, const int64_t* pos_rowptr,
                    const int32_t* pos_sorted, int k, float mask_value, int precision,
                    int32_t* out_idx, float* out_val, void* workspace, int64_t workspace_bytes,
                    lgcn_stream_t stream);

/* Bytes of device scratch lgcn_score_topk needs (bf16 operand tiles of the tensor-core
 * path; 0 for LGCN_F32). */
int64_t lgcn_score_topk_workspace_bytes(int64_t n_eval, int64_t m_items, int d, int precision);

/* Test aid: same as lgcn_score_topk, and also dumps the scores the selection saw
 * (the TMEM accumulators for LGCN_BF16) to dense_scores[n_eval, m_items]. */
int lgcn_score_topk_debug(const float* user_emb, const float* item_emb, const int64_t* user_ids,
                          int64_t n_eval, int64_t m_items, int d, const int64_t* pos_rowptr,
                          const int32_t* pos_sorted, int k, float mask_value, int precision,
                          int32_t* out_idx, float* out_val, void* workspace,
                          int64_t workspace_bytes, float* dense_scores, lgcn_stream_t stream);

/* Debug/test aid: the dense fp32 score block the F32 top-k path selects from
 * (same FMA order), scores[n_eval, m_items].  Small shapes only. */
int lgcn_score_dense_f32(const float* user_emb, const float* item_emb, const int64_t* user_ids,
                         int64_t n_eval, int64_t m_items, int d, float* scores,
                         lgcn_stream_t stream);

/* ------------------------------------------------------------------------
 * Ranking metrics on device.  Replaces utils.getLabel (utils.py:40-48),
 * RecallPrecision_ATk (metric.py:60-72) and NDCGatK_r (metric.py:84-103):
 * sums[4][n_ks] (double) += recall, precision, hr, ndcg sums over the users.
 * ------------------------------------------------------------------------ */
int lgcn_rank_metrics(const int32_t* topk, int64_t n_eval, int k, const int64_t* user_ids,
                      const int64_t* test_rowptr, const int32_t* test_sorted,
                      const int32_t* ks /*HOST, ascending*/, int n_ks, double* sums, uint8_t* hits /* [n_eval,k] or NULL */,
                      lgcn_stream_t stream);

/* ------------------------------------------------------------------------
 * Device-side ingest of the reference's interaction files.  Replaces the host
 * Python loop of Loader.__init__ (dataloader.py:93-124 train, :126-150 test): each line is
 * "uid item item ...\n"; every (uid, item) pair is appended to trainUser / trainItem in file
 * order, duplicates kept.  `text` is the raw file in device memory (16-byte aligned).
 *   lgcn_ingest_tiles(n_bytes)  -> number of 4 KiB tiles (sizes tile_tok / tile_uid)
 *   lgcn_ingest_count           -> exclusive scans of the per-tile token / uid-token counts and
 *                                  totals[0] = tokens, totals[1] = uid tokens (= non-empty lines);
 *                                  interactions = totals[0] - totals[1]
 *   lgcn_ingest_emit            -> user[i], item[i] (int64, file order), line_uid[l] = uid of the
 *                                  l-th non-empty line, line_of[i] = line ordinal of interaction i
 * err (device int32, zero on entry): bit 0 = a byte that is neither digit, space, tab, CR nor
 * newline (the reference's int() raises ValueError), bit 1 = a token longer than 18 digits,
 * bit 2 = counts and capacities disagree.
 * ------------------------------------------------------------------------ */
int64_t lgcn_ingest_tiles(int64_t n_bytes);
int lgcn_ingest_count(const uint8_t* text, int64_t n_bytes, int64_t* tile_tok, int64_t* tile_uid,
                      int64_t* totals, int32_t* err, lgcn_stream_t stream);
int lgcn_ingest_emit(const uint8_t* text, int64_t n_bytes, const int64_t* tile_tok, const int64_t* tile_uid,
                     int64_t* line_uid, int64_t n_lines_cap, int64_t* user, int64_t* item, int64_t* line_of,
                     int64_t cap_items, int32_t* err, lgcn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LGCN_B200_H */
