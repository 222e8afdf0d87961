// Internal launch prototypes shared by the kernel translation units and the engine.
#pragma once
#include "common.cuh"

namespace capb200 {

// An activation matrix [rows, cols] kept as fp32 and (tensor-core modes) as split-fp16 planes, common pitch `ld`.
struct ActView {
    float* f = nullptr;
    __half* hi = nullptr;
    __half* lo = nullptr;
    long ld = 0;
};

// One state tensor to reorder by parent row: dst[r, :] = src[src_row[r], :]
struct StateCopy {
    const float* src = nullptr;
    long ld_src = 0;
    ActView dst;
};

// ---- pointwise.cu
int state_gather_embed_launch(int rows, const int* tokens, const int* src_row, const float* emb, long ld_emb, int E, int relu,
                              ActView xt, int H, int nstate, StateCopy sc0, StateCopy sc1, cudaStream_t stream);
int lstm_pointwise_launch(int rows, int H, const float* gates, long ld_g, const int* src_row, const float* c_prev, long ld_cp,
                          float* c_out, long ld_co, ActView h_out, const float* gather_bias, long ld_gb, const int* gather_idx,
                          cudaStream_t stream);
int relu_copy_launch(const float* x, long n, ActView out_flat, cudaStream_t stream);   // out = relu(x) (+ split planes), flat
int lstm_ln_launch(int rows, int H, const float* gates, long ld_g, const float* c_prev, long ld_cp, float* c_out, long ld_co, float* h_out, long ld_h,
                   const float* ln_a, const float* ln_b, float eps, float* ln_out, long ld_ln, cudaStream_t stream);   // LSTM cell + LayerNorm of h
int maxout_pointwise_launch(int rows, int H, const float* sums, long ld_s, const int* src_row, const float* c_prev, long ld_cp,
                            float* c_out, long ld_co, ActView h_out, cudaStream_t stream);
int additive_attention_launch(int n_images, int rpi, int R, int A, int H, const float* att_h, long ld_ah, const float* p_att, long ld_pa,
                              const float* att, long ld_at, const float* mask, long ld_mask, const float* alpha_w, const float* alpha_b,
                              float* score_scratch /*[rows, R]*/, ActView out, cudaStream_t stream, float* alpha_out = nullptr /*[rows, R]*/);
int mask_rows_launch(ActView x, int n_images, int R, int cols, const float* mask, long ld_mask, cudaStream_t stream);
// scheduled sampling: tok_out[r] = label[r, col] or, with probability prob, a draw from exp(prev_logp[r, :])
int ss_select_launch(int rows, int V1, const float* prev_logp, long ld, const long long* labels, long ld_labels, int col, unsigned long long seed, float prob,
                     int* tok_out, cudaStream_t stream);

// Edits of a log-prob row before the next word is chosen (capb200_decode_edits in include/capb200.h)
struct DecodeEdits {
    int constraint = 0;
    int unk_col = -1;
    int n_bad = 0;
    const int* bad = nullptr;
    int trigrams = 0;
    int trigram_rows = 0;
    __host__ __device__ bool any() const { return constraint != 0 || unk_col >= 0 || n_bad > 0 || trigrams != 0; }
    __host__ __device__ int kinds() const { return (constraint ? 1 : 0) + (unk_col >= 0 ? 1 : 0) + (n_bad > 0 ? 1 : 0); }
};

// ---- vocab.cu : log-softmax over the vocabulary + candidate selection
struct VocabStepArgs {
    int rows = 0;
    int V1 = 0;
    float* logits = nullptr;      // [rows, V1] in/out: overwritten with log-probs (pitch ld)
    long ld = 0;
    int twice = 0;                // beam search renormalises the log-probs a second time (CaptionModel.py:204)
    float2* stats = nullptr;      // beam search: keep the raw logits in place and write (max, log-sum-exp) per row here instead
    // top-k output for beam search (k <= 16)
    int topk = 0;
    float* top_val = nullptr;     // [rows, topk]
    int* top_idx = nullptr;       // [rows, topk]
    // greedy / multinomial selection for _sample
    int select = 0;               // 0 none, 1 greedy argmax, 2 multinomial (Gumbel-max on logp / temperature), 3 forced tokens,
                                  // 4 top-k sampling, 5 nucleus (top-p) sampling
    float top = 0.f;              // k (select 4) or p (select 5)
    DecodeEdits edits;            // applied after the log-softmax, before the selection; the edited row is what is stored
    const int* prev_tokens = nullptr;   // [rows] the word fed into this step (for the edits; only read when t > 0)
    float temperature = 1.0f;
    unsigned long long seed = 0;
    unsigned long long step = 0;  // Philox offset: one independent stream per (row, step)
    const int* forced = nullptr;  // [rows] when select == 3
    int* unfinished = nullptr;    // [rows] in/out (nullptr at beam search); rows already finished emit pad and a zero row
    int first_step = 0;
    int* tokens_out = nullptr;    // [rows] next input token
    long long* seq_out = nullptr; // seq[row * ld_seq + t] = token (int64, reference dtype)
    long ld_seq = 0;
    int t = 0;
    float* picked_lp = nullptr;   // optional: picked_lp[row * ld_picked] = log-prob of the chosen token
    long ld_picked = 1;
};
int vocab_step_launch(const VocabStepArgs& a, cudaStream_t stream);
// beam search with decode edits: the per-row candidate list [rows, k_in] (k_in = beam + edits.kinds()) is edited and cut to [rows, beam]
int beam_edit_launch(int rows, int k_in, int beam, int t, const DecodeEdits& ed, const int* prev_tokens, const float* val_in, const int* idx_in,
                     float* val_out, int* idx_out, cudaStream_t stream);
int scale_rows_launch(float* x, long ld, int rows, int cols, float factor, cudaStream_t stream);

// ---- beam.cu
struct BeamState {
    int B = 0, beam = 0, T = 0, V1 = 0;
    float* sums = nullptr;        // [B, beam]
    int* seq_a = nullptr;         // [B, beam, T] ping
    int* seq_b = nullptr;         // [B, beam, T] pong
    int* hist_a = nullptr;        // [B, beam, T] row index into the step-s log-prob slab, ping
    int* hist_b = nullptr;
    int* done_cnt = nullptr;      // [B]
    int* done_seq = nullptr;      // [B, beam*T, T]
    int* done_hist = nullptr;     // [B, beam*T, T]
    int* done_len = nullptr;      // [B, beam*T]
    double* done_p = nullptr;     // [B, beam*T]  length-penalised score (the reference keeps Python floats)
    float* done_raw = nullptr;    // [B, beam*T]  raw sum of log-probs
    int* tokens = nullptr;        // [B*beam] next input tokens
    int* src_row = nullptr;       // [B*beam] parent row (index into the previous step's rows)
};
int beam_step_launch(const BeamState& s, int t, int live, const float* top_val, const int* top_idx, int penalty_kind, float penalty_alpha,
                     cudaStream_t stream);
// sorts each image's finished beams by score, writes the best `keep` records:
//   out_seq [B*keep, T] int64, out_len/out_p [B*keep], out_hist [B*keep, T] (slab rows, -1 beyond length)
int beam_finalize_launch(const BeamState& s, int keep, long long* out_seq, int* out_len, float* out_p, float* out_raw, int* out_hist,
                         cudaStream_t stream);
// dst[k, s, :] = slab[s][hist[k, s], :] (zeros where hist < 0); slab step stride `step_stride` elements
// `stats` (optional): the slab holds raw logits and stats[s * stats_stride + row] = (max, log-sum-exp); rows are normalised on the fly
// (log_softmax once at s == 0, twice afterwards, exactly as the search scored them).
// `seqs` ([nseq, T] int64, the sequences the rows belong to) + `ed`: re-apply the decode edits the search made to each row (optional)
int gather_logprob_rows_launch(const float* slab, long step_stride, long ld_slab, const int* hist, int nseq, int T, int V1, float* dst,
                               const float2* stats, long stats_stride, cudaStream_t stream, const long long* seqs = nullptr,
                               const DecodeEdits* ed = nullptr);

// ---- transformer.cu
int layer_norm_launch(int rows, int D, const float* x, long ld_x, const float* a, const float* b, float eps, ActView out, cudaStream_t st);
int embed_pe_launch(int rows, int D, const int* tokens, const float* lut, const float* pe_row, float scale, ActView out, cudaStream_t st);
int enc_self_attention_launch(int B, int R, int heads, int dk, const float* q, const float* k, const float* v, long ld, const float* mask,
                              long ld_mask, ActView out, cudaStream_t st);
int dec_self_attention_launch(int rows, int heads, int dk, int t, const float* qkv, long ld_qkv, float* kcache, float* vcache, long step_stride,
                              long ld_c, const int* anc, long ld_anc, const long long* labels, long ld_lab, ActView out, cudaStream_t st);
int cross_attention_launch(int rows, int rpi, int heads, int dk, int R, const float* q, long ld_q, const float* kk, const float* vv, long ld_kv,
                           const float* mask, long ld_mask, ActView out, cudaStream_t st);

int glu_launch(int rows, int H, const float* t, long ld_t, const float* residual, long ld_res, ActView out, cudaStream_t st);
int masked_mean_launch(int B, int R, int H, const float* x, long ld_x, const float* mask, long ld_mask, ActView out, cudaStream_t st);

// ---- gemm_generic.cu / scst_kernels.cu (training step)
int gemm_generic_launch(int ta, int tb, int M, int N, int K, const float* A, long lda, const float* B, long ldb, float* C, long ldc, int accumulate,
                        const float* bias, cudaStream_t st);
int gemm_skinny_launch(int M, int N, int nseg, const float* const* A, const long* lda, const float* const* B, const long* ldb, const int* K, const int* tb,
                       float* C, long ldc, const float* bias, const float* row_bias, long ld_rb, int rpg, int accumulate, float* scratch,
                       size_t scratch_floats, int mode, cudaStream_t st);   // mode 0 = fp32 CUDA cores, 1 = 3xTF32 mma.sync
int gemm_wgrad_launch(int M, int N, int K, const float* dY, long ld_dy, const float* X, long ld_x, float* G, long ld_g, int accumulate, int mode,
                      cudaStream_t st);   // G[M,N] (+)= dY[K,M]^T X[K,N]; mode 1 = 3xTF32 tensor cores
int colsum_launch(int rows, int cols, const float* x, long ld, float* out, int accumulate, cudaStream_t st);
// ---- gemm_tf32.cu: tcgen05 kind::tf32 3-pass GEMM on fp32 operands (in-kernel hi/lo split), split-K over a cluster with a DSMEM reduction
struct Tf32Context;      // per-engine cache of encoded tensor maps and transposed operands
Tf32Context* tf32_context_create();
void tf32_context_destroy(Tf32Context* c);
void tf32_context_new_step(Tf32Context* c);          // weights may have changed: per-step transposes are rebuilt on next use
long tf32_context_launches(const Tf32Context* c);
bool gemm_tf32_supported(int nseg, const float* const* X, const long* ldx, const float* const* W, const long* ldw, const int* K);
int gemm_tf32_launch(Tf32Context* ctx, int M, int N, int nseg, const float* const* X, const long* ldx, const float* const* W, const long* ldw, const int* K,
                     float* C, long ldc, const float* bias, const float* row_bias, long ld_rb, int rpg, int accumulate, cudaStream_t st);
const float* tf32_transposed(Tf32Context* ctx, const float* src, long ld_src, int rows, int cols, bool per_step, long* ld_dst, cudaStream_t st);
int dropout_apply_launch(float* x, int rows, int cols, long ld, unsigned long long seed, unsigned site, unsigned step, float p, cudaStream_t st);
int dropout_mask_launch(float* m, long n, unsigned long long seed, unsigned site, unsigned step, float p, cudaStream_t st);
int dropout_copy_launch(const float* x, long ld_x, float* y, long ld_y, int rows, int cols, unsigned long long seed, unsigned site, unsigned step, float p,
                        cudaStream_t st);
int embed_relu_dropout_launch(int rows, int E, const int* tokens, const float* emb, float* xt, unsigned long long seed, unsigned step, float p,
                              cudaStream_t st);
int scst_dlogits_launch(const float* logp, long ld_row, const long long* seq, const float* reward, const float* mask_sum, float upstream, int N, int T, int V1,
                        float* dl, cudaStream_t st, const float* row_coef = nullptr);
// drop_worst on the sampled-loss rows: row_msum / row_coef [N] scratch; loss[0] = mean of the `keep` smallest row losses
int scst_drop_worst_launch(const long long* seq, const float* row_loss, int N, int T, int keep, float upstream, float* row_msum, float* row_coef, float* loss,
                           cudaStream_t st);
int xe_loss_backward_launch(const float* logp, long ld_row, const long long* labels, long ld_l, const float* masks, long ld_m, int N, int steps, int Ls, int V1,
                            float smoothing, float upstream, float* mask_sum, float* item_loss, float* dl, float* loss, cudaStream_t st, int keep = 0,
                            float* row_loss = nullptr, float* row_msum = nullptr, float* row_coef = nullptr);
int lstm_cell_backward_launch(int rows, int H, const float* gates, const float* c_prev, const float* c_new, const float* dh, const float* dh_extra,
                              long ld_extra, unsigned drop_site, unsigned drop_step, unsigned long long seed, float p, float* dc_carry, float* dgates,
                              cudaStream_t st);
int attention_backward_launch(int n_images, int rpi, int R, int A, int H, const float* d_out, const float* alpha, const float* att_h, const float* p_att,
                              const float* att, const float* w, float* d_att_h, float* d_att, float* d_p_att, float* d_w, float* d_b, float* d_alpha_scratch,
                              cudaStream_t st);
int relu_dropout_backward_launch(long n, const float* x, const float* dy, float* dx, float scale, cudaStream_t st);
int embed_backward_launch(int rows, int E, const int* tokens, const float* xt, const float* dxt, long ld_dxt, float scale, float* d_emb, cudaStream_t st);
int per_image_sum_launch(int steps, int rows, int rpi, int cols, const float* x, float* out, cudaStream_t st);
int add_strided_launch(float* a, const float* b, long ld_b, int rows, int cols, cudaStream_t st);

// ---- aoa_train_kernels.cu (AoANet training step)
int ln_backward_launch(int rows, int D, const float* x, long ld_x, const float* a, const float* dy, long ld_dy, float eps, float* dx, long ld_dx, int accumulate,
                       float* stats, float* da, float* db, int accumulate_params, cudaStream_t st, const float* add1 = nullptr, long ld_a1 = 0,
                       const float* add2 = nullptr, long ld_a2 = 0);
// fused element-wise steps of the AoANet decoder loop (aoa_train_kernels.cu)
int aoa_step_inputs_launch(int rows, int E, int H, int rpi, const int* tok_src, int* tok_dst, const float* emb, float* xt, const float* mean, long ld_mean,
                           const float* out_prev, float* x1c, unsigned long long seed, int step, float p_lm, float p_ctx, cudaStream_t st);
int glu_dropout_launch(int rows, int H, const float* t, long ld_t, float* out, long ld_o, float* outd, long ld_d, unsigned long long seed, int step, float p,
                       cudaStream_t st);
int glu_backward_fused_launch(int rows, int H, const float* t, long ld_t, const float* d_outd, long ld_dd, const float* dctx, float* dt, long ld_dt,
                              unsigned long long seed, int step, float p, cudaStream_t st);
int glu_backward_launch(int rows, int H, const float* t, long ld_t, const float* dy, long ld_dy, float* dt, long ld_dt, cudaStream_t st);
// ---- seed salt (dropout.cuh): effective seed of every dropout / sampling kernel = seed argument XOR salt; uploaded in stream order
int dropout_salt_set_scst(unsigned long long salt, cudaStream_t st);
int dropout_salt_set_aoa(unsigned long long salt, cudaStream_t st);
int dropout_salt_set_tfm(unsigned long long salt, cudaStream_t st);
int dropout_salt_set_vocab(unsigned long long salt, cudaStream_t st);
inline int dropout_salt_set_all(unsigned long long salt, cudaStream_t st) {
    return dropout_salt_set_scst(salt, st) | dropout_salt_set_aoa(salt, st) | dropout_salt_set_tfm(salt, st) | dropout_salt_set_vocab(salt, st);
}
// ---- tfm_train_kernels.cu: element-wise pieces of the Transformer training steps on TIME-major rows (row = t * rps + n; dropout keyed by (t, n))
int embed_pe_dropout_launch(int rows, int rps, int D, const int* tok, const float* lut, const float* pe, float scale, int t0, unsigned long long seed, int site,
                            float p, float* x, long ld, cudaStream_t st);
int embed_pe_backward_launch(int rows, int rps, int D, const int* tok, float scale, int t0, unsigned long long seed, int site, float p, const float* dx, long ld,
                             float* dlut, cudaStream_t st);
int add_dropout_rows_launch(int rows, int rps, int cols, int t0, const float* a, long ld_a, const float* b, long ld_b, float* out, long ld_o, unsigned long long seed,
                            int site, float p, cudaStream_t st);
int dropout_rows_copy_launch(int rows, int rps, int cols, int t0, const float* src, long ld_s, float* dst, long ld_d, unsigned long long seed, int site, float p,
                             const float* relu_of, long ld_r, cudaStream_t st);
int relu_dropout_rows_launch(int rows, int rps, int cols, int t0, float* h, long ld, unsigned long long seed, int site, float p, cudaStream_t st);
int permute_rows_launch(int L, int N, int D, const float* src, long ld_s, float* dst, long ld_d, int to_seq_major, cudaStream_t st);
int load_tokens_tm_launch(const long long* labels, long ld, int N, int L, int* tok, float* key_mask, long ld_m, cudaStream_t st);
// sequence self-attention with replayable dropout (aoa_train_kernels.cu): row(b, pos) = b * b_stride + pos * p_stride
int seq_attn_train_launch(int seqs, int n_keys, int q_lo, int q_hi, int heads, int dk, int causal, int idx_L, long b_stride, long p_stride, const float* q,
                          const float* k, const float* v, long ld, unsigned long long seed, int site, float p, float* out, long ld_out, const float* key_mask,
                          long ld_mask, cudaStream_t st);
int seq_attn_backward_launch(int seqs, int n_keys, int heads, int dk, int causal, int idx_L, long b_stride, long p_stride, const float* q, const float* k,
                             const float* v, long ld, unsigned long long seed, int site, float p, const float* d_out, long ld_do, float* dq, float* dk_,
                             float* dv, long ld_d, const float* key_mask, long ld_mask, cudaStream_t st);
int enc_attn_train_launch(int B, int R, int heads, int dk, const float* q, const float* k, const float* v, long ld, unsigned long long seed, int site, float p,
                          float* out, long ld_out, cudaStream_t st, const float* mask = nullptr, long ld_mask = 0);
int enc_attn_backward_launch(int B, int R, int heads, int dk, const float* q, const float* k, const float* v, long ld, unsigned long long seed, int site, float p,
                             const float* d_out, long ld_do, float* dq, float* dk_, float* dv, long ld_d, cudaStream_t st, const float* mask = nullptr,
                             long ld_mask = 0);
int cross_attn_train_launch(int rows, int rpi, int heads, int dk, int R, const float* q, long ld_q, const float* kk, const float* vv, long ld_kv,
                            unsigned long long seed, int site, int step, float p, float* out, long ld_out, float* probs, cudaStream_t st,
                            const float* mask = nullptr, long ld_mask = 0, int row_mod = 0);
int cross_attn_backward_launch(int B, int rpi, int heads, int dk, int R, const float* q, long ld_q, const float* kk, const float* vv, long ld_kv,
                               unsigned long long seed, int site, int step, float p, const float* probs, const float* d_out, long ld_do, float* dq, long ld_dq,
                               float* dkk, float* dvv, long ld_dkv, cudaStream_t st, int n_steps = 1, int row_mod = 0);
int mean_backward_launch(int B, int R, int H, const float* d_mean, long ld_dm, float* dx, long ld_dx, cudaStream_t st, const float* mask = nullptr,
                         long ld_mask = 0);
int add_dropout_launch(int rows, int cols, const float* a, long ld_a, const float* b, long ld_b, float* out, long ld_o, unsigned long long seed, int site, int step,
                       float p, cudaStream_t st);
int cat_dropout_launch(int rows, int c1, int c2, const float* a, long ld_a, const float* b, long ld_b, float* out, long ld_o, unsigned long long seed, int site,
                       int step, float p, cudaStream_t st);

// ---- reward.cu (CIDEr-D) and criterion
struct CiderTable;   // device hash table of n-gram -> idf
CiderTable* cider_table_create(const int* keys, const double* df, long n, double ref_len, cudaStream_t stream);
void cider_table_destroy(CiderTable* t);
int cider_reward_launch(const CiderTable* t, const long long* sampled, int S, const long long* greedy, int B, int T, const int* refs,
                        const int* ref_offsets, int L, double* scores, float* reward, long ld_reward, int reward_cols, cudaStream_t stream);
int reward_criterion_fwd_launch(const float* logprobs, long ld_row, long ld_t, const long long* seq, const float* reward, int N, int T,
                                float* loss_mean, float* loss_rows, float* mask_sum, cudaStream_t stream);
int reward_criterion_bwd_launch(const long long* seq, const float* reward, int N, int T, const float* mask_sum, float upstream,
                                float* grad, long ld_row, long ld_t, cudaStream_t stream);

}  // namespace capb200
