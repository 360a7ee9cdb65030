/*
 * gnnrag_b200.h -- C ABI of libgnnrag_b200.so: the B200 (sm_100a) implementation of GNN-RAG's GNN
 * retrieval hot path (ReaRev / NSM multi-hop message passing + answer scoring + candidate ranking).
 *
 * The reference (cmavro/GNN-RAG) has no FFI; its boundary is the Python nn.Module contract that
 * gnn/train_model.py:49-57,222 and gnn/evaluate.py:160 call.  The Python mirror in gnn_rag_b200/ keeps
 * that contract and calls the entry points below through ctypes.  Each entry point names the
 * reference code it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the CALLER allocates every input, output and workspace; the library never allocates, frees
 *     or retains a pointer; it is stateless and re-entrant per stream;
 *   - `stream` is a cudaStream_t passed as void* (e.g. torch.cuda.current_stream().cuda_stream);
 *   - all work is enqueued asynchronously on `stream`; nothing synchronises;
 *   - return value: 0 = GR_OK, negative = error (see gr_status); never throws;
 *     gr_last_error() returns a thread-local message for the last failing call;
 *   - node ids are GLOBAL rows b*N + local (gnn/dataset_load.py:483); index arrays produced by the
 *     library are int32; floating point is fp32 unless stated;
 *   - edge arrays (src/rel/w/fact) must be allocated with capacity gr_pad4(F) elements and row
 *     pointer arrays with capacity gr_pad4(Nt + 1) elements (the staging copies read whole 16-byte
 *     chunks).
 */
#ifndef GNNRAG_B200_H_
#define GNNRAG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GR_ABI_VERSION 1

typedef enum gr_status {
  GR_OK = 0,
  GR_ERR_INVALID_ARG = -1,   /* null pointer, bad size, unsupported combination            */
  GR_ERR_CUDA = -2,          /* a CUDA runtime call failed (message in gr_last_error)        */
  GR_ERR_WORKSPACE = -3,     /* workspace too small                                          */
  GR_ERR_UNSUPPORTED = -4    /* shape/dtype not handled by this build                        */
} gr_status;

/* flags for gr_linear */
#define GR_LINEAR_RELU 1u        /* C = relu(A W^T + b)                                      */
#define GR_LINEAR_EXACT_FP32 2u  /* force the fp32 SIMT kernel (no split-bf16 tensor-core path) */
#define GR_LINEAR_BF16_SINGLE 8u /* gr_linear_tc_planes: bf16 activation storage -- ONE product A_hi W_hi (A_lo ignored, may
                                    be NULL); bf16-input / fp32-accumulate accuracy, a third of the tensor work   */
#define GR_LINEAR_W_PRESPLIT 4u  /* gr_linear_tc_planes: the workspace still holds the bf16 hi/lo split of the same
                                  * W (same N, K, k_seg, k_seg_pitch) from an earlier call: skip the conversion pass.
                                  * For inference with fixed weights (weight pre-formatting, done once per weight
                                  * version by the caller). */

int gr_abi_version(void);
const char* gr_last_error(void);
/* runtime switches: "agg_tma" (0|1: stage CSR slices with bulk TMA copies), "linear_tc" (0|1: split-bf16
 * tcgen05 GEMM for gr_linear when the shape allows), "tc_cluster" (1|2: CTAs per cluster that share the W
 * tiles of the tcgen05 GEMM through TMA multicast), "agg_abs_ws" (0|1: persistent warp-specialised build of the |v| aggregation
 * kernel).  Process-wide; set before launching work. */
int gr_set_option(const char* name, int64_t value);
static inline int64_t gr_pad4(int64_t n) { return (n + 3) & ~(int64_t)3; }

/* ------------------------------------------------------------------------------------------------
 * CSR batching.  Replaces BaseGNNLayer.build_matrix (gnn/modules/kg_reasoning/base_gnn.py:19-51: seven
 * uncoalesced COO tensors) and the index part of TypeLayer.forward (gnn/modules/layer_init.py:32-37).
 * Input: the batched fact list of SingleDataLoader._build_fact_mat (gnn/dataset_load.py:473-527),
 * copied to the device as-is (idx_bytes = 8 for the reference's int64 arrays, 4 for int32).
 * Output: in-edges grouped by TAIL (forward messages: src = head) and by HEAD (inverse messages:
 * src = tail); inside a row, edges keep the ORIGINAL FACT ORDER (stable), so per-row reductions are a
 * pure function of the row's own fact sequence (deterministic; ties stay ties).
 *   rowptr_*: int32[Nt+1]; src_*, rel_*, fact_*: int32[F] (fact_* = original fact id of each slot).
 *   status: int32[1], set non-zero on device if an id is out of range (ids are clamped).
 *   nfacts: optional device int32[1].  When given, only the first min(F, *nfacts) fact slots are read: F is then the
 *   CAPACITY of fixed-shape input buffers (CUDA-graph replay over batches of different fact counts); everything
 *   downstream sees the live facts only, through the row pointers.
 * Workspace: gr_csr_build_workspace_bytes(F, Nt).
 */
size_t gr_csr_build_workspace_bytes(int64_t F, int64_t Nt);
int gr_csr_build(const void* heads, const void* rels, const void* tails, int idx_bytes,
                 int64_t F, int64_t Nt, int64_t num_rel_rows,
                 int32_t* rowptr_t, int32_t* src_t, int32_t* rel_t, int32_t* fact_t,
                 int32_t* rowptr_h, int32_t* src_h, int32_t* rel_h, int32_t* fact_h,
                 int32_t* status, const int32_t* nfacts, void* workspace, size_t workspace_bytes, void* stream);

/* out[e] = in[fact[e]] -- permute a per-fact fp32 array (weight_list / weight_rel_list of
 * gnn/dataset_load.py:509-517) into CSR slot order. */
int gr_gather_f32(const float* in, const int32_t* fact, float* out, int64_t F, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense linear layer  C[m, n] = act( sum_k A[m,k] W[n,k] + bias[n] ) (+ addend[m, n] for m < addend_rows)
 * A: [M,K] row stride lda; W: [N,K] row stride ldw (torch nn.Linear layout); C: row stride ldc.
 * Used for the HOISTED relation projection rel_linear_k(rel_features) (reasongnn.py:79,105 apply it to
 * F gathered rows; the R1 distinct rows suffice), `addend` = pos_emb rows (reasongnn.py:75-77), and for
 * e2e_linear (reasongnn.py:163, nsm_gnn.py:63).  bias / addend may be NULL.
 */
int gr_linear(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
              const float* addend, int64_t ld_addend, int64_t addend_rows,
              float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, uint32_t flags, void* stream);

/* Tensor-core variant of gr_linear for the big e2e_linear GEMMs (reasongnn.py:163, nsm_gnn.py:63):
 * fp32 in / fp32 out with fp32-class accuracy through the 3-product split-bf16 scheme on tcgen05
 * (x = hi + lo in bf16; A W^T ~= A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T, fp32 accumulation in TMEM,
 * dropped term <= 2^-18 relative).  Persistent, TMA-fed 128 x N tiles over 32- or 64-column k-blocks ("tc_bk"),
 * W shared across a CTA pair by TMA multicast ("tc_cluster"), TMA-store epilogue ("tc_tma_store").  Requires
 * 8 <= N <= 256.  The workspace (256-byte aligned, gr_linear_tc_workspace_bytes) holds the bf16 planes.
 * flags: GR_LINEAR_RELU. */
size_t gr_linear_tc_workspace_bytes(int64_t M, int64_t N, int64_t K);
int gr_linear_tc(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                 float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, uint32_t flags,
                 void* workspace, size_t workspace_bytes, void* stream);
/* Same GEMM with the A operand already in split-bf16 planes (written by gr_aggregate_dual / gr_type_layer /
 * a previous call): no conversion pass over A.  Outputs (each optional, at least one): fp32 C; bf16 planes
 * C_hi/C_lo (row stride ldc16) = the node-embedding columns of the NEXT layer's A operand; and
 * dots (float[2*M]): the score_func dot product (reasongnn.py:165) fused into the epilogue as two partial
 * sums over the lower / upper half of the output columns, dots[m] + dots[M+m] = sum_n C[m,n] * w_score[n]
 * (gr_masked_softmax adds them).  Persistent kernel, TMEM accumulators double-buffered (epilogue overlaps the next tile).
 * Segmented K: when k_seg_pitch > k_seg > 0 the A planes hold K/k_seg_pitch segments of k_seg valid columns
 * at pitch k_seg_pitch (zero padding in between) while W is the dense [N, (K/k_seg_pitch)*k_seg] torch weight;
 * the W planes are built in the padded layout.  K is the padded length.
 * Workspace: gr_linear_tc_planes_workspace_bytes(N, K) (the W planes), 256-byte aligned.
 * flags: GR_LINEAR_RELU, GR_LINEAR_W_PRESPLIT (the workspace still holds this W's planes: skip the conversion). */
size_t gr_linear_tc_planes_workspace_bytes(int64_t N, int64_t K);
int gr_linear_tc_planes(const void* A_hi, const void* A_lo, int64_t lda16, const float* W, int64_t ldw,
                        const float* bias, float* C, int64_t ldc, void* C_hi, void* C_lo, int64_t ldc16,
                        const float* w_score, float* dots, int64_t M, int64_t N, int64_t K,
                        int64_t k_seg, int64_t k_seg_pitch,
                        uint32_t flags, void* workspace, size_t workspace_bytes, void* stream);
/* fp32 [M,K] (row stride lda) -> bf16 hi/lo planes (row stride ld_out, multiple of 8). */
int gr_split_bf16(const float* A, int64_t lda, int64_t M, int64_t K, void* hi, void* lo, int64_t ld_out,
                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * The aggregation kernel family.
 *
 * gr_aggregate: one direction.  For every destination row n and instruction j < I
 *     out[n, out_col0 + j*seg_stride + d] = sum_{e in row n} relu(table[rel_e, d] * ins[b(n), j, d]) * c_e
 *     c_e = w_e * (w_e * prior[src_e])        (w_e = 1 when w == NULL)
 * = ReasonGNNLayer.reason_layer (reasongnn.py:61-89) with the tail CSR, reason_layer_inv (:91-116) with
 * the head CSR, NSMLayer.reason_layer (nsm_gnn.py:87-112) with I = 1.  `table` is the hoisted
 * rel_linear(rel_features) [R1, D].  ins: [B, I, D] contiguous.  b(n) = n / N.
 * possible (optional, float[Nt]): 1.0 where sum_e c_e > 1e-10 (nsm_gnn.py:101-103).
 * Edges with c_e == 0 are skipped exactly (relu(x)*0 = 0 for finite x).
 *
 * gr_aggregate_dual: both directions of one ReaRev GNN layer in one launch; instruction j writes
 *     forward  (tail CSR, table_fwd) -> columns out_col0 + (2j  )*D
 *     inverse  (head CSR, table_inv) -> columns out_col0 + (2j+1)*D
 * which is the concat order of ReasonGNNLayer.forward (reasongnn.py:150-161).  seg_pitch (0 = D) is the
 * column distance between consecutive segments: the bf16 planes use a pitch rounded up to 16 columns so that
 * every segment starts on a 32-byte sector (measured on B200: a 16-byte-misaligned segment start halves the
 * achievable write bandwidth, scripts/agg_probe.py).
 *
 * Split-bf16 planes (optional, gr_aggregate_dual / gr_type_layer): when out_hi/out_lo are non-NULL the
 * result is ALSO (or, with out == NULL, only) written as two bf16 matrices with row stride ld_planes and
 * the same column indexing as `out`: hi = bf16(y), lo = bf16(y - hi), so hi + lo = y to 2^-18 relative.
 * That is the A-operand layout gr_linear_tc_planes consumes, so the fp32 concat buffer of
 * reasongnn.py:158-161 never has to exist.
 *
 * gr_type_layer: out[n,:] = relu( sum_{tail CSR} w_e table[rel_e] + sum_{head CSR} w_e table[rel_e] ),
 * TypeLayer.forward (layer_init.py:46-57) with table = kb_self_linear(rel_features).
 */
int gr_aggregate(const int32_t* rowptr, const int32_t* src, const int32_t* rel, const float* w,
                 const float* prior, const float* table, const float* ins,
                 float* out, int64_t out_row_stride, int64_t out_col0, int64_t seg_stride,
                 float* possible, int B, int N, int D, int I, int64_t F, void* stream);

/* Backward of gr_aggregate (csrc/aggregate_bwd.cu; the kernel behind model(batch, training=True), gnn/train_model.py:222):
 * given grad_out[n, grad_col0 + j*seg_stride + d] = dL/dout it ACCUMULATES (+=, caller zeroes)
 *     grad_table[r, :] += sum_{e: rel_e = r} c_e sum_j grad_out[n_e, j, :] * ins[b, j, :] * [table[r] * ins[b, j] > 0]
 *     grad_ins[b, j, :] += sum_{e in b} c_e grad_out[n_e, j, :] * table[rel_e, :] * [table[rel_e] * ins[b, j] > 0]
 *     grad_prior[s]     += sum_{e: src_e = s} w_e^2 sum_j <grad_out[n_e, j, :], relu(table[rel_e] * ins[b, j])>
 * over the same destination CSR as the forward call.  D <= 256, I <= 4.  fp32 atomics: summation order is not
 * deterministic (neither is the reference's sparse.mm backward on CUDA). */
int gr_aggregate_backward(const int32_t* rowptr, const int32_t* src, const int32_t* rel, const float* w,
                          const float* prior, const float* table, const float* ins, const float* grad_out,
                          int64_t grad_row_stride, int64_t grad_col0, int64_t seg_stride, float* grad_table,
                          float* grad_ins, float* grad_prior, int B, int N, int D, int I, int64_t F, void* stream);

int gr_aggregate_dual(const int32_t* rowptr_t, const int32_t* src_t, const int32_t* rel_t,
                      const float* w_t, const int32_t* rowptr_h, const int32_t* src_h,
                      const int32_t* rel_h, const float* w_h, const float* prior,
                      const float* table_fwd, const float* table_inv, const float* ins,
                      float* out, int64_t out_row_stride, int64_t out_col0, int64_t seg_pitch,
                      void* out_hi, void* out_lo, int64_t ld_planes,
                      int B, int N, int D, int I, int64_t F, void* stream);

/* Specialised variant of gr_aggregate_dual for the hot shape (csrc/aggregate_abs.cu).  The hoisted relation table is
 * copied once per layer into a zero-padded 256-column layout (gr_pad_table256: table [rows, D] fp32, row stride ldt
 * -> out [rows][256] fp32, 16-byte aligned) so every lane of the gather is in-bounds, and the edge loop accumulates
 * sum c*v and sum c*|v| (|.| is a free FFMA2 source modifier on sm_100) instead of taking relu of every gathered
 * element: sum c*relu(+-v) = (Q +- S)/2.  Output: the split-bf16 planes only.  This build specialises D = 200,
 * seg_pitch = 208, N >= 64 (gr_aggregate_dual_abs_supported); other shapes use gr_aggregate_dual.
 * Same reference lines: reasongnn.py:61-116. */
int gr_pad_table256(const float* table, int64_t ldt, int64_t rows, int D, float* pn, void* stream);
int gr_aggregate_dual_abs_supported(int N, int D, int64_t seg_pitch, int64_t R1);
int gr_aggregate_dual_abs(const int32_t* rowptr_t, const int32_t* src_t, const int32_t* rel_t, const float* w_t,
                         const int32_t* rowptr_h, const int32_t* src_h, const int32_t* rel_h, const float* w_h,
                         const float* prior, const float* pn_fwd, const float* pn_inv, int64_t table_rows,
                         const float* ins, void* out_hi, void* out_lo, int64_t ld_planes, int64_t out_col0,
                         int64_t seg_pitch, int B, int N, int D, int I, int64_t F, int32_t* tile_counter, void* stream);
/* table_rows: rows (R1) of each padded table.
 * tile_counter: 4 bytes of device scratch (zeroed by the call) -> the persistent, warp-specialised kernel (a staging
 * warp prepares the next tile of rows while the consumer warps aggregate the current one; dynamic tile scheduler);
 * NULL -> one CTA per 64-row tile.  gr_set_option("agg_abs_ws", v) picks the persistent variant: 1 = 8 consumer warps /
 * 64-row tiles, 2 (default) = 9 consumer warps / 72-row tiles, 3 = gather through TMA tile::gather4 copies into
 * shared-memory rings (bit-identical results; slower at cfg2, kept as the measured alternative). */

/* One dense-prior ReaRev layer as ONE kernel (csrc/fused_layer.cu): the aggregation of both directions and all
 * instructions (reason_layer / reason_layer_inv, reasongnn.py:61-116) is produced straight into the shared-memory
 * operand slots of the tcgen05 e2e GEMM (torch.cat + e2e_linear + relu, reasongnn.py:158-163; score_func dot :165), so
 * the 2*I neighbour segments never reach HBM.  Replaces the pair gr_aggregate_dual_abs -> gr_linear_tc_planes.
 *   h_hi / h_lo    bf16 planes of the layer input h: [B*N, >= seg_pitch] with row stride ldh16 (only the first
 *                  seg_pitch columns are read; columns D .. seg_pitch-1 must be zero)
 *   pn_fwd/pn_inv  zero-padded 256-column relation tables of this layer (gr_pad_table256)
 *   W              e2e_linear weight [N_out, (2I+1)*D] fp32 (row stride ldw); it is re-ordered and split into bf16 hi/lo
 *                  planes inside `workspace` (gr_fused_layer_workspace_bytes, 256-byte aligned) unless flags carries
 *                  GR_LINEAR_W_PRESPLIT (workspace kept from an earlier call with the same W)
 *   outputs        any of C (fp32 [B*N, N_out]), C_hi / C_lo (bf16 planes, row stride ldc16), dots [2*B*N]
 *                  (dots[m] = <out[m], w_score>, dots[B*N + m] = 0: the layout gr_masked_softmax takes)
 * Supported (gr_fused_layer_supported): I <= 2, N >= 128, seg_pitch % 16 == 0, seg_pitch <= 224, N_out <= 256 and the
 * operand stages must fit shared memory (D = N_out = 200 does).  The A operand is bit-identical to the unfused pair;
 * the tensor core accumulates the k-blocks in a different order (fp32 rounding). */
/* Diagnostic: per-CTA wait-cycle counters of the last gr_fused_layer launch made with gr_set_option("fused_debug", 32)
 * (16 uint64 per CTA, slot meaning in csrc/fused_layer.cu). */
int gr_fused_profile_read(unsigned long long* out, int n);
int gr_fused_layer_supported(int64_t N_nodes, int64_t D, int64_t seg_pitch, int I, int64_t N_out);
size_t gr_fused_layer_workspace_bytes(int64_t D, int64_t seg_pitch, int I, int64_t N_out);
/* Once per batch: both CSRs of gr_csr_build re-laid out slot-major per quad of 4 rows ("quad ELL", csrc/fused_layer.cu)
 * into `ell` (gr_fused_ell_bytes, 256-byte aligned); gr_fused_layer reads it and keeps its per-layer {table offset,
 * coefficient} scratch inside the same buffer. */
size_t gr_fused_ell_bytes(int B, int N_nodes, int64_t F);
int gr_fused_ell_build(const int32_t* rowptr_t, const int32_t* src_t, const int32_t* rel_t, const float* w_t,
                       const int32_t* rowptr_h, const int32_t* src_h, const int32_t* rel_h, const float* w_h,
                       int B, int N_nodes, int64_t F, void* ell, size_t ell_bytes, void* stream);
int gr_fused_layer(const int32_t* rowptr_t, const int32_t* src_t, const int32_t* rel_t, const float* w_t,
                   const int32_t* rowptr_h, const int32_t* src_h, const int32_t* rel_h, const float* w_h,
                   const float* prior, const float* pn_fwd, const float* pn_inv, const float* ins,
                   const void* h_hi, const void* h_lo, int64_t ldh16, int64_t seg_pitch, const float* W,
                   int64_t ldw, const float* bias, float* C, int64_t ldc, void* C_hi, void* C_lo,
                   int64_t ldc16, const float* w_score, float* dots, int B, int N_nodes, int D, int I,
                   int64_t N_out, int64_t F, uint32_t flags, void* workspace, size_t workspace_bytes,
                   void* ell, size_t ell_bytes, void* stream);

/* Diagnostic only (scripts/agg_probe.py): replays the aggregation kernel's store pattern without any edge work. */
int gr_debug_store_probe(void* hi, void* lo, int64_t Nt, int64_t ld, int col_start, int ncols, int mode,
                         void* stream);

int gr_type_layer(const int32_t* rowptr_t, const int32_t* rel_t, const float* w_t,
                  const int32_t* rowptr_h, const int32_t* rel_h, const float* w_h,
                  const float* table, float* out, int64_t out_row_stride,
                  void* out_hi, void* out_lo, int64_t ld_planes,
                  int B, int N, int D, int64_t F, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sparse-prior fast path for one ReaRev layer (the first layer of every iteration sees the seed distribution,
 * rearev.py:208).  Rows none of whose in-edges carries prior mass get exactly zero neighbour messages, so
 * h_new = relu(W[:, :D] h + b) there (gr_linear_tc_planes with K = one segment).  gr_frontier_rows lists the
 * other rows (exact for any prior); gr_frontier_fixup recomputes those rows in full -- aggregation of both
 * directions for every instruction (same edge order/arithmetic as gr_aggregate_dual), e2e linear, relu, score
 * dot -- and overwrites h_new in the next planes / fp32 h / dots.  list: int32[Nt], count: int32[1] (device).
 */
int gr_frontier_rows(const int32_t* rowptr_t, const int32_t* src_t, const int32_t* rowptr_h,
                     const int32_t* src_h, const float* prior, int64_t Nt, int32_t* list, int32_t* count,
                     void* stream);
int gr_frontier_fixup(const int32_t* rowptr_t, const int32_t* src_t, const int32_t* rel_t, const float* w_t,
                      const int32_t* rowptr_h, const int32_t* src_h, const int32_t* rel_h, const float* w_h,
                      const float* prior, const float* table_fwd, const float* table_inv, const float* ins,
                      const void* cur_hi, const void* cur_lo, int64_t ld_cur, const float* W, int64_t ldw,
                      const float* bias, const float* w_score, void* nxt_hi, void* nxt_lo, int64_t ld_nxt,
                      float* h32, float* dots, const int32_t* list, const int32_t* count,
                      int B, int N, int D, int I, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Scoring: logits[b,n] = dot(h[b,n,:], w_score) + b_score + (1 - mask[b,n]) * (-1e11);
 * dist = softmax_n(logits).  reasongnn.py:165-169 / nsm_gnn.py:67-74.  One CTA per question.
 * h row stride ldh.  mask: float[B*N] (local_entity != num_entity, times possible_tail for NSM
 * reason_kb).  logits_out optional.
 */
int gr_score_softmax(const float* h, int64_t ldh, const float* w_score, const float* b_score,
                     const float* mask, float* dist, float* logits_out, int B, int N, int D,
                     void* stream);
/* Same, starting from precomputed score dots (gr_linear_tc_planes epilogue): logit = dots[n] (+ dots2[n] if
 * dots2 != NULL) + b_score + (1-mask)*VERY_NEG. */
int gr_masked_softmax(const float* dots, const float* dots2, const float* b_score, const float* mask,
                      float* dist, int B, int N, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Question-side updates, one CTA per question (csrc/question.cu).  In the reference each is a chain of 10-20
 * tiny torch ops on [B, D] tensors; fused here because at B200 speeds they are pure launch latency.
 * Pointer arrays named *_host are HOST arrays of device pointers (one per instruction, <= 8).
 *
 * gr_instructions: LSTMInstruction.forward after the encoder (gnn/modules/question_encoding/
 * base_encoder.py:73-114): for i < I:  q_i = question_linear_i(qnode);  cq = cq_linear([ri, q_i, q_i-ri, q_i*ri]);
 * attn = softmax_q(ca_linear(cq * hidden[q]) + (1-mask[q])*VERY_NEG);  ri = sum_q attn[q]*hidden[q];
 * out[b,i,:] = ri (ri starts at 0).  hidden [B,Q,D], qnode [B,D], qtext int64 [B,Q] (mask = qtext != pad_id).
 * attn_out optional [B,I,Q]. */
int gr_instructions(const float* hidden, const float* qnode, const int64_t* qtext, int64_t pad_id,
                    const float* const* Wq_host, const float* const* bq_host, const float* Wcq,
                    const float* bcq, const float* wca, const float* bca, float* out, float* attn_out,
                    int B, int Q, int D, int I, void* stream);
/* gr_query_reform: the instruction update after every iteration (gnn/models/ReaRev/rearev.py:214-221 ->
 * QueryReform.forward, gnn/modules/query_update.py:18-44, Fusion :6-16): y = seed_info[b] @ h[b] (seed rows
 * only, index order); for j < I: z = [x_j, y, x_j - y]; g = sigmoid(G_j z); out_j = g * (R_j z) + (1-g) * x_j.
 * ins_in/ins_out [B,I,D] (may not alias); Wr/Wg: fusion.r / fusion.g weights [D,3D]; seed_out optional [B,D]. */
int gr_query_reform(const float* seed_info, const float* h, int64_t ldh, const float* ins_in,
                    const float* const* Wr_host, const float* const* Wg_host, float* ins_out,
                    float* seed_out, int B, int N, int D, int I, void* stream);
/* gr_kl_loss_pred: BaseModel.calc_loss_label with loss_type "kl" (gnn/models/base_model.py:186-215,
 * rearev.py:156-160,228-232) and pred = argmax_n dist[b,n] (lowest index on ties):
 * loss = sum_b valid_b * sum_n kl_div(log(dist+1e-8), teacher/len_b) / B.  loss_q: float[B] scratch/output. */
int gr_kl_loss_pred(const float* dist, const float* teacher, float* loss_q, float* loss, int64_t* pred,
                    int B, int N, void* stream);

/* gr_lstm_forward: recurrence of the one-layer LSTM question encoder (gnn/modules/question_encoding/
 * lstm_encoder.py:27-36: nn.LSTM(batch_first=True), h0 = c0 = 0, gate order i,f,g,o) for all Q tokens in ONE
 * launch (8-CTA clusters, W_hh slices resident in shared memory, h exchanged through distributed shared memory).
 * gates_x [B,Q,4D] = x_t W_ih^T + b_ih (computed by the caller); W_hh [4D,D]; b_hh [4D] or NULL;
 * hidden [B,Q,D] out (h_n = hidden[:, Q-1]).  D <= gr_lstm_max_hidden() (256). */
size_t gr_lstm_max_hidden(void);
int gr_lstm_forward(const float* gates_x, const float* W_hh, const float* b_hh, float* hidden, int B, int Q,
                    int D, void* stream);

/* seed_retrieve[b,:] = sum_n seed_info[b,n] * h[b,n,:]  (torch.bmm in QueryReform.forward,
 * gnn/modules/query_update.py:40); only rows with seed_info != 0 are read, in index order. */
int gr_seed_retrieve(const float* seed_info, const float* h, int64_t ldh, float* out,
                     int B, int N, int D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Candidate ranking = the retrieved answer-node set, Evaluator.evaluate + f1_and_hits
 * (gnn/evaluate.py:156, 188-209, 25-50).  Per question: drop seeds (query_entities == 1), pads
 * (local_entity == pad_id) and p < (1-eps)/N (compared in double); stable sort by p descending
 * (ties keep local-index order); keep the prefix up to and including the item at which the
 * sequential float64 running sum exceeds eps.
 *   cand_idx:  int32[B, N]  local indices in retrieval order (first cand_count[b] valid)
 *   cand_count:int32[B]; cand_total:int32[B] = number of candidates before the eps cut.
 * Workspace: gr_rank_workspace_bytes(B, N).
 */
size_t gr_rank_workspace_bytes(int B, int N);
int gr_rank_candidates(const float* dist, const int64_t* local_entity, const float* query_entities,
                       int64_t pad_id, double eps, int32_t* cand_idx, int32_t* cand_count,
                       int32_t* cand_total, int B, int N, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Shortest-path node sets: nodes lying on any shortest path between any seed and
 * any retrieved candidate in the UNDIRECTED subgraph -- build_graph + get_truth_paths,
 * llm/src/utils/graph_utils.py:10-21,49-75.  Uses both CSRs of a question batch; one CTA per question.
 *   sources / targets: float[B*N] indicator arrays (non-zero = member).
 *   on_path: uint8[B*N] output; dist_src: int32[B*N] workspace/output (BFS depth from the sources... see DESIGN.md)
 */
size_t gr_paths_workspace_bytes(int B, int N, int max_sources, int max_targets);
int gr_shortest_path_nodes(const int32_t* rowptr_t, const int32_t* src_t,
                           const int32_t* rowptr_h, const int32_t* src_h,
                           const int32_t* source_idx, const int32_t* source_cnt, int max_sources,
                           const int32_t* target_idx, const int32_t* target_cnt, int max_targets,
                           uint8_t* on_path, int32_t* pair_dist, int B, int N,
                           void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GNNRAG_B200_H_ */
