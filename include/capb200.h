/* capb200 -- C ABI of the B200-native caption-decoding / SCST engine.
 *
 * Drop-in boundary.  The reference (ruotianluo/ImageCaptioning.pytorch) is pure Python and has no FFI layer; its boundary
 * for this path is two Python call surfaces:
 *     CaptionModel.forward(..., mode='sample'|'forward')            captioning/models/CaptionModel.py:29-33
 *       -> AttModel._sample / _sample_beam / _forward              captioning/models/AttModel.py:258,218,126
 *     LossWrapper.forward(..., sc_flag=True)                        captioning/modules/loss_wrapper.py:56-73
 * The Python adapter in imagecaptioning.pytorch_b200 keeps those signatures and calls the entry points below through
 * ctypes.  Every entry point takes plain device pointers and sizes (no torch types), an explicit cudaStream_t passed as
 * void*, is asynchronous with respect to the host unless stated, returns 0 on success and non-zero on failure with a
 * message available from capb200_last_error().  PyTorch owns all tensors; the engine owns only packed weight copies and
 * workspaces.  One engine per device; different engines may be driven from different host threads.
 *
 * All matrices are row-major fp32 unless noted; token ids are int64 (the reference's torch.long) at the boundary.
 */
#ifndef CAPB200_H_
#define CAPB200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define CAPB200_ABI_VERSION 1

/* numeric modes of the dense contractions */
#define CAPB200_MODE_SIMT_FP32 0 /* fp32 FFMA on CUDA cores */
#define CAPB200_MODE_TC_F16X3 1  /* tcgen05 kind::f16, split-fp16 operands, 3 MMA passes, fp32 accumulate (parity grade) */
#define CAPB200_MODE_TC_F16X1 2  /* tcgen05 kind::f16, single pass (throughput mode, not parity grade) */
#define CAPB200_MODE_SKINNY_TF32X3 3 /* capb200_linear only: the training step's split-K GEMM, 3xTF32 mma.sync on the fp32 weights */
#define CAPB200_MODE_SKINNY_FP32 4   /* capb200_linear only: same split-K GEMM on CUDA cores */
#define CAPB200_MODE_TF32X3_TC 5       /* capb200_linear only: the training steps' tcgen05 kind::tf32 3-pass GEMM on fp32 operands (gemm_tf32.cu) */
#define CAPB200_MODE_TF32X3_TC_DGRAD 6 /* same kernel, input-gradient form: y[M,N] = x[M,K] * w[K,N]  (w row-major [K,N], transposed internally) */
#define CAPB200_MODE_TF32X3_TC_WGRAD 7 /* same kernel, weight-gradient form: y[M,N] = x[K,M]^T * w[K,N] (both row-major, transposed internally) */

#define CAPB200_FAMILY_UPDOWN 0 /* UpDownModel  captioning/models/AttModel.py:868 */
#define CAPB200_FAMILY_NEWFC 1  /* NewFCModel   captioning/models/AttModel.py:904 */

typedef struct capb200_engine capb200_engine;
typedef struct capb200_cider_table capb200_cider_table;

const char* capb200_last_error(void);
int capb200_abi_version(void);
/* Numeric precondition of the tensor-core modes: every value that is converted to split-fp16 planes (weights at bind time, the fc / att
 * feature tiles of each call) must be finite with |x| < 65504.  A violation sets a process-wide flag; bind_weights checks it synchronously,
 * every later entry point fails with a message while it is set.  Returns the flag (valid once the stream of the offending call has been
 * synchronised); reset != 0 clears it. */
int capb200_range_status(int reset);

/* ------------------------------------------------------------------------------------------------------------------
 * Operator level (each replaces one library call of the reference's per-timestep core; used by the parity tests)
 * ---------------------------------------------------------------------------------------------------------------- */

/* y[M,N] = x[M,K] * w[N,K]^T + b[N] (optional ReLU)             nn.Linear call sites AttModel.py:74-95,172,733
 * mode selects the arithmetic; the tensor-core modes split x and w into fp16 planes in scratch memory first. */
int capb200_linear(const float* x, long ldx, const float* w, long ldw, const float* b, float* y, long ldy, int M, int N, int K,
                   int relu, int mode, void* stream);

/* Same contraction with the operands split once, then `iters` back-to-back launches timed with CUDA events on `stream`
 * (synchronous; y holds the result afterwards).  Used by bench.py / the tiling sweeps. */
int capb200_bench_linear(const float* x, const float* w, const float* b, float* y, int M, int N, int K, int mode, int iters, float* ms_per_launch,
                         void* stream);

/* Diagnostics: one traced launch of the decode GEMM y[M,N] = x[M,K] w[N,K]^T (tc_f16x3, CTA-pair kernel).  trace_host receives 296 x 16
 * %globaltimer stamps (ns), one row per CTA: 0 set-up done, 1 first operands landed, 2/3 all MMAs of the CTA pair's first / second tile
 * issued, 4/5 accumulator of tile 0 / 1 complete (epilogue starts), 6/7 epilogue of tile 0 / 1 done, 8 kernel end (tools/gemm_trace.py). */
int capb200_gemm_trace(const float* x, const float* w, float* y, int M, int N, int K, unsigned long long* trace_host, int n_slots, void* stream);

/* nn.LSTMCell: gates = x*w_ih^T + b_ih + h*w_hh^T + b_hh; (i,f,g,o)           AttModel.py:628,635
 * x[M,Kx], h/c[M,H] -> h_out/c_out[M,H] */
int capb200_lstm_cell(const float* x, int Kx, const float* h, const float* c, const float* w_ih, const float* w_hh, const float* b_ih,
                      const float* b_hh, float* h_out, float* c_out, int M, int H, int mode, void* stream);

/* Attention.forward (AttModel.py:728-748) with per-image features: row r uses image r / rows_per_image.
 * att_h[rows,A] = h2att(h) incl. bias; p_att[B,R,A]; att[B,R,H]; mask[B,R] or NULL; alpha_w[A], alpha_b[1] -> out[rows,H] */
int capb200_additive_attention(const float* att_h, const float* p_att, const float* att, const float* mask, const float* alpha_w,
                               const float* alpha_b, float* out, int n_images, int rows_per_image, int R, int A, int H, void* stream);

/* In-place log_softmax over each row of logits[rows,V1] (twice != 0 applies it a second time, CaptionModel.py:204) and
 * the per-row top-k (values and indices, descending, lowest index first on ties). top_val/top_idx may be NULL if k == 0. */
int capb200_log_softmax_topk(float* logits, long ld, int rows, int V1, int twice, int k, float* top_val, int* top_idx, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Engine level
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
    int family;              /* CAPB200_FAMILY_* */
    int vocab_size;          /* V; logits have V+1 entries, id 0 = BOS = EOS = PAD (AttModel.py:65-67) */
    int input_encoding_size; /* E */
    int rnn_size;            /* H */
    int att_hid_size;        /* A */
    int fc_feat_size;        /* F_fc */
    int att_feat_size;       /* F_att */
    int seq_length;          /* T = max_length (AttModel.py:60) */
    int numeric_mode;        /* CAPB200_MODE_* */
} capb200_model_cfg;

/* Borrowed fp32 device pointers in the reference's state_dict layouts (SURVEY.md section 8b key list). */
typedef struct {
    const float* embed;                                                           /* [V+1,E]  embed.0.weight | embed.weight */
    const float *fc_embed_w, *fc_embed_b;                                         /* [H,F_fc] (newfc: [E,F_fc]) */
    const float *att_embed_w, *att_embed_b;                                       /* [H,F_att] */
    const float *ctx2att_w, *ctx2att_b;                                           /* [A,H] */
    const float *logit_w, *logit_b;                                               /* [V+1,H] */
    const float *att_lstm_w_ih, *att_lstm_w_hh, *att_lstm_b_ih, *att_lstm_b_hh;   /* [4H,E+2H] [4H,H] [4H] [4H] */
    const float *lang_lstm_w_ih, *lang_lstm_w_hh, *lang_lstm_b_ih, *lang_lstm_b_hh; /* [4H,2H] [4H,H] [4H] [4H] */
    const float *h2att_w, *h2att_b;                                               /* [A,H] */
    const float *alpha_w, *alpha_b;                                               /* [1,A] [1] */
    const float *i2h_w, *i2h_b, *h2h_w, *h2h_b;                                   /* newfc _core: [5H,E] [5H] [5H,H] [5H] */
} capb200_weights;

capb200_engine* capb200_engine_create(const capb200_model_cfg* cfg);
void capb200_engine_destroy(capb200_engine* e);
/* (Re)binds the parameter tensors; call again after every optimizer step.  Tensor-core modes repack the fp16 planes here. */
int capb200_engine_bind_weights(capb200_engine* e, const capb200_weights* w, void* stream);

/* Per-step edits of the log-prob rows before the next word is chosen (the reference's decode options; the edited row is also what the
 * reference stores in seqLogprobs / done_beams[...]['logps'], and so do we).  All zero / -1 / NULL = none. */
typedef struct {
    int decoding_constraint;   /* t > 0: log-prob of the previous word = -inf (CaptionModel.py:154-155, AttModel.py:294-297) */
    int unk_col;               /* beam search: this column's log-prob is lowered by 1000 at every step (suppress_UNK with 'UNK' as the last
                                  word, or unk_idx: CaptionModel.py:159-162); -1 = none */
    int n_bad_endings;         /* remove_bad_endings: t > 0 and previous word in the list -> log-prob of the end token (column 0) = -inf
                                  (CaptionModel.py:156-157, AttModel.py:299-304) */
    const int* bad_endings;    /* device, n_bad_endings word ids (model.bad_endings_ix) */
    int block_trigrams;        /* _sample only, t >= 3: every earlier occurrence of (w[t-2], w[t-1], x) lowers x by 2 ln 2 (AttModel.py:306-332) */
    int trigram_rows;          /* rows 0 .. trigram_rows-1 get it (the reference loops over batch_size rows, also when sample_n > 1) */
} capb200_decode_edits;

typedef struct {
    int beam_size;      /* b, 1..16 (with edits: b + number of active edit kinds <= 16) */
    int sample_n;       /* 1 or beam_size (AttModel.py:223) */
    int penalty_kind;   /* 0 '' (identity), 1 'wu_<alpha>', 2 'avg_<alpha>'   captioning/utils/misc.py:133-151 */
    float penalty_alpha;
    float temperature;  /* log_softmax(logprobs / temperature) from the second step on (CaptionModel.py:204); 0 is read as 1 */
    capb200_decode_edits edits;
} capb200_beam_opts;

/* AttModel._sample_beam + CaptionModel.beam_search (group_size 1).
 * fc[B,F_fc], att[B,R,F_att] contiguous, mask[B,R] or NULL.
 * seq[B*sample_n,T] int64 zero padded; seq_logprobs[B*sample_n,T,V+1] (NULL to skip the gather);
 * done_seq[B,b,T] int64, done_len[B,b], done_p[B,b], done_raw[B,b]: each image's finished beams sorted by score (may be NULL). */
int capb200_decode_beam(capb200_engine* e, const float* fc, const float* att, const float* mask, int B, int R, const capb200_beam_opts* opts,
                        long long* seq, float* seq_logprobs, long long* done_seq, int* done_len, float* done_p, float* done_raw, void* stream);
/* Full log-prob rows [len, V+1] of finished beam `rank` of image `image` from the most recent capb200_decode_beam call
 * (done_beams[image][rank]['logps'], CaptionModel.py:192); dst must hold T*(V+1) floats, rows beyond the length are zeroed. */
int capb200_beam_record_logprobs(capb200_engine* e, int image, int rank, float* dst, void* stream);

#define CAPB200_SAMPLE_GREEDY 0
#define CAPB200_SAMPLE_MULTINOMIAL 1
#define CAPB200_SAMPLE_FORCED 2  /* replay given tokens (parity checks against another sampler's draw) */
#define CAPB200_SAMPLE_TEACHER 3 /* AttModel._forward: feed labels[:, t] at step t, no finished-row masking */
#define CAPB200_SAMPLE_TOPK 4    /* sample_method 'top<k>', k >= 1: multinomial over the k most likely words of logprobs / temperature
                                    (CaptionModel.py:398-402); `top` = k */
#define CAPB200_SAMPLE_TOPP 5    /* sample_method 'top<p>', 0 < p < 1: nucleus sampling (CaptionModel.py:388-397); `top` = p */
typedef struct {
    int sample_n;             /* rows per image */
    int method;               /* CAPB200_SAMPLE_* */
    float temperature;
    unsigned long long seed;  /* Philox key for the sampling methods */
    int steps;                /* TEACHER: number of label columns to run (<= the label width) */
    float top;                /* TOPK: k, TOPP: p */
    capb200_decode_edits edits;
} capb200_sample_opts;

/* AttModel._sample (greedy / multinomial) and AttModel._forward (teacher forcing).
 * tokens_in[N,ld_tok] int64: forced tokens (FORCED) or labels (TEACHER), else NULL.  N = B*sample_n.
 * seq[N,T] int64; seq_logprobs[N,T_out,V+1] where T_out = T (sampling) or ld_tok (TEACHER); picked[N,T] optional. */
int capb200_decode_sample(capb200_engine* e, const float* fc, const float* att, const float* mask, int B, int R, const capb200_sample_opts* opts,
                          const long long* tokens_in, long ld_tok, long long* seq, float* seq_logprobs, float* picked, void* stream);

/* Number of this library's kernel launches issued through the engine since creation (bench.py reports it). */
long capb200_engine_launch_count(const capb200_engine* e);

/* Optional device-side timing of the dense contractions (cudaEvent pairs recorded on the launching stream around every
 * GEMM launch).  ids: 0 fc_embed, 1 att_embed, 2 ctx2att, 3 fc->gate bias, 4 att_lstm gates, 5 h2att, 6 lang_lstm gates,
 * 7 logit, 8 newfc core.  read_profile synchronises the device and returns accumulated milliseconds, algorithmic FLOPs
 * (2*M*N*K) and launch counts per id; n must be >= 9. */
int capb200_engine_set_profiling(capb200_engine* e, int enable);
int capb200_engine_read_profile(capb200_engine* e, int reset, double* ms, double* flops, long* calls, int n);

/* ------------------------------------------------------------------------------------------------------------------
 * Transformer captioner (TransformerModel, captioning/models/TransformerModel.py:237-363)
 *   prologue = att_embed + N_enc encoder layers (TransformerModel.py:305-338); decode keeps a K/V cache per layer instead of
 *   re-running all t tokens each step (:351-363) -- identical results because the decoder mask is causal.
 * ---------------------------------------------------------------------------------------------------------------- */
#define CAPB200_TFM_MAX_LAYERS 8
typedef struct capb200_tfm_engine capb200_tfm_engine;
typedef struct {
    int vocab_size, d_model, d_ff, heads, n_enc, n_dec, att_feat_size, seq_length, numeric_mode;
} capb200_tfm_cfg;
typedef struct { const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b; } capb200_mha_weights;          /* linears.0..3 */
typedef struct {
    capb200_mha_weights self_attn;
    const float *w1_w, *w1_b, *w2_w, *w2_b;                                                                   /* feed_forward.w_1 / w_2 */
    const float *ln0_a, *ln0_b, *ln1_a, *ln1_b;                                                               /* sublayer.{0,1}.norm.{a_2,b_2} */
} capb200_tfm_enc_layer;
typedef struct {
    capb200_mha_weights self_attn, src_attn;
    const float *w1_w, *w1_b, *w2_w, *w2_b;
    const float *ln0_a, *ln0_b, *ln1_a, *ln1_b, *ln2_a, *ln2_b;
} capb200_tfm_dec_layer;
typedef struct {
    const float *att_embed_w, *att_embed_b;                     /* att_embed.0 [D, F_att] */
    capb200_tfm_enc_layer enc[CAPB200_TFM_MAX_LAYERS];          /* model.encoder.layers.i */
    const float *enc_norm_a, *enc_norm_b;                       /* model.encoder.norm */
    capb200_tfm_dec_layer dec[CAPB200_TFM_MAX_LAYERS];          /* model.decoder.layers.i */
    const float *dec_norm_a, *dec_norm_b;                       /* model.decoder.norm */
    const float *lut, *pe;                                      /* model.tgt_embed.0.lut.weight [V+1, D]; model.tgt_embed.1.pe [1, 5000, D] */
    const float *gen_w, *gen_b;                                 /* model.generator.proj [V+1, D] */
} capb200_tfm_weights;

capb200_tfm_engine* capb200_tfm_create(const capb200_tfm_cfg* cfg);
void capb200_tfm_destroy(capb200_tfm_engine* e);
int capb200_tfm_bind_weights(capb200_tfm_engine* e, const capb200_tfm_weights* w, void* stream);
/* same contracts as capb200_decode_beam / capb200_beam_record_logprobs / capb200_decode_sample; fc features are unused */
int capb200_tfm_decode_beam(capb200_tfm_engine* e, const float* att, const float* mask, int B, int R, const capb200_beam_opts* opts, long long* seq,
                            float* seq_logprobs, long long* done_seq, int* done_len, float* done_p, float* done_raw, void* stream);
int capb200_tfm_beam_record_logprobs(capb200_tfm_engine* e, int image, int rank, float* dst, void* stream);
int capb200_tfm_decode_sample(capb200_tfm_engine* e, const float* att, const float* mask, int B, int R, const capb200_sample_opts* opts,
                              const long long* tokens_in, long ld_tok, long long* seq, float* seq_logprobs, float* picked, void* stream);
long capb200_tfm_launch_count(const capb200_tfm_engine* e);

/* ------------------------------------------------------------------------------------------------------------------
 * AoANet (AoAModel, captioning/models/AoAModel.py:188-226; configs/aoa.yml: refine=1, refine_aoa=1, use_ff=0,
 * decoder_type=AoA, use_multi_head=2, multi_head_scale=1, mean_feats=1)
 * ---------------------------------------------------------------------------------------------------------------- */
#define CAPB200_AOA_REFINER_LAYERS 6
typedef struct capb200_aoa_engine capb200_aoa_engine;
typedef struct {
    int vocab_size, input_encoding_size, rnn_size, heads, att_feat_size, seq_length, numeric_mode;
} capb200_aoa_cfg;
typedef struct {
    const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b;   /* refiner.layers.i.self_attn.linears.{0,1,2} [H,H] */
    const float *aoa_w, *aoa_b;                       /* refiner.layers.i.self_attn.aoa_layer.0 [2H,2H] */
    const float *ln_a, *ln_b;                         /* refiner.layers.i.sublayer.0.norm.{a_2,b_2} */
} capb200_aoa_refiner_layer;
typedef struct {
    const float* embed;                               /* embed.0.weight [V+1,E] */
    const float *att_embed_w, *att_embed_b;           /* att_embed.0 [H,F_att] */
    capb200_aoa_refiner_layer refiner[CAPB200_AOA_REFINER_LAYERS];
    const float *refiner_norm_a, *refiner_norm_b;     /* refiner.norm */
    const float *ctx2att_w, *ctx2att_b;               /* ctx2att [2H,H] */
    const float *att_lstm_w_ih, *att_lstm_w_hh, *att_lstm_b_ih, *att_lstm_b_hh;   /* core.att_lstm [4H,E+H] [4H,H] */
    const float *attn_norm_a, *attn_norm_b;           /* core.attention.norm */
    const float *attn_q_w, *attn_q_b;                 /* core.attention.linears.0 [H,H] */
    const float *att2ctx_w, *att2ctx_b;               /* core.att2ctx.0 [2H,2H] */
    const float *logit_w, *logit_b;                   /* logit [V+1,H] */
} capb200_aoa_weights;

capb200_aoa_engine* capb200_aoa_create(const capb200_aoa_cfg* cfg);
void capb200_aoa_destroy(capb200_aoa_engine* e);
int capb200_aoa_bind_weights(capb200_aoa_engine* e, const capb200_aoa_weights* w, void* stream);
int capb200_aoa_decode_beam(capb200_aoa_engine* e, const float* att, const float* mask, int B, int R, const capb200_beam_opts* opts, long long* seq,
                            float* seq_logprobs, long long* done_seq, int* done_len, float* done_p, float* done_raw, void* stream);
int capb200_aoa_beam_record_logprobs(capb200_aoa_engine* e, int image, int rank, float* dst, void* stream);
int capb200_aoa_decode_sample(capb200_aoa_engine* e, const float* att, const float* mask, int B, int R, const capb200_sample_opts* opts,
                              const long long* tokens_in, long ld_tok, long long* seq, float* seq_logprobs, float* picked, void* stream);
long capb200_aoa_launch_count(const capb200_aoa_engine* e);

/* ------------------------------------------------------------------------------------------------------------------
 * SCST reward and criterion
 * ---------------------------------------------------------------------------------------------------------------- */
/* Document-frequency table in the scripts/prepro_ngrams.py format, flattened: keys[n,4] int32 token ids padded with -1,
 * df[n] float64, ref_len = number of reference images (ciderD_scorer.py:108-111).  Host pointers; synchronous. */
capb200_cider_table* capb200_cider_table_create(const int* keys, const double* df, long n, double ref_len, void* stream);
void capb200_cider_table_destroy(capb200_cider_table* t);

/* get_self_critical_reward (captioning/utils/rewards.py:41-81) with CIDEr-D only:
 * sampled[S,T], greedy[B,T] int64 device; refs[n_refs_total,L] int32 device (0 padded), ref_offsets[B+1] int32 device;
 * scores[S+B] float64 device (CIDEr-D of every hypothesis); reward[S,T] fp32 = score(sample) - score(greedy of its image). */
int capb200_self_critical_reward(const capb200_cider_table* t, const long long* sampled, int S, const long long* greedy, int B, int T,
                                 const int* refs, const int* ref_offsets, int L, double* scores, float* reward, void* stream);

/* get_scores (captioning/utils/rewards.py:83-114) with cider_reward_weight = 1: CIDEr-D of S = B*n sampled captions against their
 * image's references -> scores[S] float64.  reward (optional, [S,T] fp32): score minus the mean score of the image's other samples,
 * the per-token weight of the 'new_self_critical' structure loss (losses.py:168-187). */
int capb200_cider_scores(const capb200_cider_table* t, const long long* sampled, int S, int B, int T, const int* refs, const int* ref_offsets,
                         int L, double* scores, float* reward, void* stream);

/* RewardCriterion.forward (captioning/modules/losses.py:22-37). logprobs[N,T,V1]; seq[N,T] int64; reward[N,T].
 * loss_mean[1], loss_rows[N] (reduction 'none'), mask_sum[1]; any output may be NULL. */
int capb200_reward_criterion_forward(const float* logprobs, const long long* seq, const float* reward, int N, int T, int V1, float* loss_mean,
                                     float* loss_rows, float* mask_sum, void* stream);
/* d loss_mean / d logprobs, scaled by `upstream`; grad[N,T,V1] must be zero-filled by the caller. */
int capb200_reward_criterion_backward(const long long* seq, const float* reward, int N, int T, int V1, const float* mask_sum, float upstream,
                                      float* grad, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * One self-critical training step of the UpDown model (LossWrapper.forward with sc_flag, loss_wrapper.py:56-73, plus the
 * loss.backward() of tools/train.py:189): eval-mode greedy baseline, train-mode multinomial samples (dropout on, AttModel.py:74-88,
 * :637), CIDEr-D self-critical reward, RewardCriterion, then back-propagation through time into every parameter gradient.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
    int sample_n;              /* opt.train_sample_n */
    float temperature;
    unsigned long long seed;   /* Philox key of the sampler and of the dropout masks */
    float drop_prob;           /* drop_prob_lm; 0 disables dropout */
    float upstream;            /* d(total loss)/d(this loss), normally 1 */
    int baseline;              /* CAPB200_BASELINE_GREEDY (self-critical, loss_wrapper.py:56-73) or CAPB200_BASELINE_LEAVE_ONE_OUT
                                  (structure loss 'new_self_critical', losses.py:168-187: no greedy decode, greedy_seq may be NULL) */
    const long long* forced_tokens; /* optional [B*sample_n, T] int64 device: replay these samples instead of drawing them (parity
                                  checks against the reference's own multinomial draw); NULL = sample */
    const float* att_masks;    /* optional [B, R] fp32 device (1 = valid region): variable region counts (AttModel.py:44-49,106-112,742-744;
                                  dataloader.py:230-241); the caller has clipped R to the longest valid length; NULL = all regions valid */
    int keep_rows;             /* drop_worst (tools/train.py:187-191): 0 = reduction 'mean'; k > 0 = the criterion runs with reduction 'none' (one loss per
                                  caption row) and the k rows with the smallest loss are averaged: loss[0] = that mean, gradients accordingly */
    float* row_loss;           /* optional [rows] output of the per-row losses (what LossWrapper returns as out['loss'] under drop_worst_flag) */
} capb200_scst_opts;
#define CAPB200_BASELINE_GREEDY 0
#define CAPB200_BASELINE_LEAVE_ONE_OUT 1
/* Gradient buffers, one per capb200_weights field (same shapes, fp32, device); every one is OVERWRITTEN. */
typedef struct {
    float* embed;
    float *fc_embed_w, *fc_embed_b, *att_embed_w, *att_embed_b, *ctx2att_w, *ctx2att_b, *logit_w, *logit_b;
    float *att_lstm_w_ih, *att_lstm_w_hh, *att_lstm_b_ih, *att_lstm_b_hh, *lang_lstm_w_ih, *lang_lstm_w_hh, *lang_lstm_b_ih, *lang_lstm_b_hh;
    float *h2att_w, *h2att_b, *alpha_w, *alpha_b;
} capb200_updown_grads;
/* fc[B,F_fc], att[B,R,F_att] (opts->att_masks for variable region counts); refs as in capb200_self_critical_reward.
 * Outputs: sample_seq[B*n,T] int64, greedy_seq[B,T] int64, sample_logprobs[B*n,T,V+1] (caller zero-fills), reward[B*n,T], loss[1].
 * Execution: the whole step (~900-4900 launches, none of them data dependent) is captured into ONE CUDA graph the second time a configuration
 * -- shapes, every pointer argument, every option except the seed -- is seen, and replayed afterwards (the *_scst_step entry points of all
 * three families; CAPB200_SCST_GRAPH=0 disables it).  For that the step runs on an engine-owned stream that first waits for `stream` and that
 * `stream` is made to wait for before the call returns; the features (and the region mask) are copied into an engine-owned staging buffer, so
 * they may live anywhere, while refs / ref_offsets / the output and gradient buffers should keep their addresses from step to step (a changed
 * address is a new configuration: one eager step, one capture).  A replay draws new samples and masks from opts->seed exactly as the eager
 * step would (the seed reaches the kernels through a device-side salt), so results do not depend on whether a step was replayed. */
int capb200_updown_scst_step(capb200_engine* e, const float* fc, const float* att, int B, int R, const capb200_scst_opts* opts,
                             const capb200_cider_table* table, const int* refs, const int* ref_offsets, int L, const capb200_updown_grads* grads,
                             long long* sample_seq, long long* greedy_seq, float* sample_logprobs, float* reward, float* loss, void* stream);
/* ------------------------------------------------------------------------------------------------------------------
 * One cross-entropy (XE) training step of the UpDown model: the teacher-forced AttModel._forward (AttModel.py:126-164; train mode,
 * dropout on, no scheduled sampling) over labels[..., :-1], LanguageModelCriterion or LabelSmoothing (losses.py:204-265) against
 * labels[..., 1:] / masks[..., 1:] with reduction 'mean' (loss_wrapper.py:54-55), and back-propagation through time.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
    int seq_per_img;           /* label rows per image (the reference repeats the features, utils.repeat_tensors) */
    int steps;                 /* columns actually evaluated: the reference stops at the first column i >= 1 whose tokens are all 0 */
    unsigned long long seed;   /* Philox key of the dropout masks */
    float drop_prob;
    float label_smoothing;     /* 0 = LanguageModelCriterion, > 0 = LabelSmoothing(smoothing) */
    float upstream;
    const float* att_masks;    /* optional [B, R] region mask, see capb200_scst_opts */
    float ss_prob;             /* scheduled sampling (AttModel.py:145-154): from the second step on, each row's input word is drawn from the model's
                                  previous prediction with this probability; 0 = teacher forcing */
    long long* tokens_used;    /* optional [N, label_cols-1] int64: the words actually fed (labels, or the draws where scheduled sampling hit) */
    int keep_rows;             /* drop_worst (tools/train.py:187-191): 0 = reduction 'mean'; k > 0 = the criterion runs with reduction 'none' (one loss per
                                  caption row) and the k rows with the smallest loss are averaged: loss[0] = that mean, gradients accordingly */
    float* row_loss;           /* optional [rows] output of the per-row losses (what LossWrapper returns as out['loss'] under drop_worst_flag) */
} capb200_xe_opts;
/* labels[N, label_cols] int64 (column 0 = BOS = 0), masks[N, label_cols] fp32, N = B * seq_per_img.
 * Outputs: logprobs[N, label_cols-1, V+1] (caller zero-fills; columns >= steps stay zero), loss[1], every grads buffer overwritten. */
int capb200_updown_xe_step(capb200_engine* e, const float* fc, const float* att, int B, int R, const capb200_xe_opts* opts, const long long* labels,
                           const float* masks, int label_cols, const capb200_updown_grads* grads, float* logprobs, float* loss, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * One self-critical training step of the AoANet model (BASELINE configs[3]): LossWrapper.forward with sc_flag (loss_wrapper.py:56-73)
 * over AoAModel (refiner + AoA decoder, AoAModel.py:56-226) in train mode, and its backward.  Dropout sites (replayable through
 * capb200_dropout_mask with the same seed): 1 att_embed [B*R,H]; 2 word embedding at `step` [N,E]; 3 core output at `step` [N,H];
 * 4 ctx_drop at `step` [N,H]; 5 decoder attention probabilities at `step` [N,heads,R]; 10+l refiner attention probabilities
 * [B,heads,R,R]; 20+l AoA-layer input [B*R,2H]; 30+l refiner SublayerConnection [B*R,H]   (l = refiner layer 0..5).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
    int sample_n;
    float temperature;
    unsigned long long seed;
    float upstream;
    int baseline;              /* CAPB200_BASELINE_* */
    float drop_prob_lm;        /* att_embed, word embedding, ctx_drop, out_drop (opt.drop_prob_lm) */
    float drop_attn;           /* MultiHeadedDotAttention dropout on the probabilities (0.1, AoAModel.py:18) */
    float drop_aoa;            /* dropout_aoa (0.3) */
    float drop_sublayer;       /* refiner SublayerConnection (0.1, AoAModel.py:119) */
    int ctx_drop;              /* opt.ctx_drop */
    const long long* forced_tokens; /* optional [B*sample_n, T] int64 device: replay these samples (see capb200_scst_opts) */
    const float* att_masks;    /* optional [B, R] region mask (refiner self-attention keys, masked mean pooling AoAModel.py:216-219, decoder
                                  attention keys) */
    int keep_rows;             /* drop_worst (tools/train.py:187-191): 0 = reduction 'mean'; k > 0 = the criterion runs with reduction 'none' (one loss per
                                  caption row) and the k rows with the smallest loss are averaged: loss[0] = that mean, gradients accordingly */
    float* row_loss;           /* optional [rows] output of the per-row losses (what LossWrapper returns as out['loss'] under drop_worst_flag) */
} capb200_aoa_scst_opts;
/* Gradient buffers, laid out field by field like the weights struct above: parameter shapes, fp32, device; every one is OVERWRITTEN. */
typedef struct {
    float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *aoa_w, *aoa_b, *ln_a, *ln_b;
} capb200_aoa_refiner_layer_grads;
typedef struct {
    float* embed;
    float *att_embed_w, *att_embed_b;
    capb200_aoa_refiner_layer_grads refiner[CAPB200_AOA_REFINER_LAYERS];
    float *refiner_norm_a, *refiner_norm_b;
    float *ctx2att_w, *ctx2att_b;
    float *att_lstm_w_ih, *att_lstm_w_hh, *att_lstm_b_ih, *att_lstm_b_hh;
    float *attn_norm_a, *attn_norm_b;
    float *attn_q_w, *attn_q_b;
    float *att2ctx_w, *att2ctx_b;
    float *logit_w, *logit_b;
} capb200_aoa_grads;
/* att[B,R,F_att] (opts->att_masks for variable region counts); outputs as in capb200_updown_scst_step. */
int capb200_aoa_scst_step(capb200_aoa_engine* e, const float* att, int B, int R, const capb200_aoa_scst_opts* opts, const capb200_cider_table* table,
                          const int* refs, const int* ref_offsets, int L, const capb200_aoa_grads* grads, long long* sample_seq, long long* greedy_seq,
                          float* sample_logprobs, float* reward, float* loss, void* stream);

/* One cross-entropy training step of AoANet (teacher-forced AttModel._forward over AoAModel in train mode + LanguageModelCriterion /
 * LabelSmoothing + backward); arguments as capb200_updown_xe_step, dropout sites as capb200_aoa_scst_step. */
typedef struct {
    int seq_per_img;
    int steps;                 /* columns actually evaluated (early break of AttModel.py:158-159) */
    unsigned long long seed;
    float label_smoothing;
    float upstream;
    float drop_prob_lm, drop_attn, drop_aoa, drop_sublayer;
    int ctx_drop;
    const float* att_masks;    /* optional [B, R] region mask */
    float ss_prob;             /* scheduled sampling, see capb200_xe_opts */
    long long* tokens_used;
    int keep_rows;             /* drop_worst (tools/train.py:187-191): 0 = reduction 'mean'; k > 0 = the criterion runs with reduction 'none' (one loss per
                                  caption row) and the k rows with the smallest loss are averaged: loss[0] = that mean, gradients accordingly */
    float* row_loss;           /* optional [rows] output of the per-row losses (what LossWrapper returns as out['loss'] under drop_worst_flag) */
} capb200_aoa_xe_opts;
int capb200_aoa_xe_step(capb200_aoa_engine* e, const float* att, int B, int R, const capb200_aoa_xe_opts* opts, const long long* labels,
                        const float* masks, int label_cols, const capb200_aoa_grads* grads, float* logprobs, float* loss, void* stream);

/* Gradient-group events for an overlapped data-parallel all-reduce (tools/train_pl.py:479: DDP buckets the reference's gradients the
 * same way).  A training step finishes its gradient buffers in a fixed order of groups; after the last write of group k it records
 * events[k] (cudaEvent_t, caller-owned) on the step's stream, so a communication stream can all-reduce group k while the rest of the
 * backward pass still runs.  n = 0 or events = NULL switches the recording off.  Groups:
 *   UpDown (capb200_engine_set_grad_events, n <= 2): 0 logit.{weight,bias}; 1 every other parameter.
 *   AoANet (capb200_aoa_set_grad_events, n <= 10):  0 logit; 1 decoder (att2ctx, attention q-projection and norm, att_lstm) + embed;
 *           2 ctx2att + refiner.norm; 3..8 refiner layers 5..0; 9 att_embed. */
int capb200_engine_set_grad_events(capb200_engine* e, void* const* events, int n);
int capb200_aoa_set_grad_events(capb200_aoa_engine* e, void* const* events, int n);

/* The dropout keep/scale mask (0 or 1/(1-p)) of one site and step, for tests that replay it in the oracle:
 * site 0 = fc_embed [B,H], 1 = att_embed [B*R,H], 2 = word embedding at `step` [N,E], 3 = core output at `step` [N,H]. */
int capb200_dropout_mask(float* mask, long n, unsigned long long seed, int site, int step, float p, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training steps of the Transformer captioner (TransformerModel.py:262-363 under LossWrapper, loss_wrapper.py:25-73, + loss.backward()).
 *   capb200_tfm_xe_step    the teacher-forced TransformerModel._forward (:340-348: one pass over all label_cols - 1 positions; keys that
 *                          are pad / eos are masked, position 0 never, :323-328) + LanguageModelCriterion / LabelSmoothing + backward;
 *                          labels / masks / logprobs / loss as capb200_updown_xe_step (logprobs [N, label_cols - 1, V+1], all positions)
 *   capb200_tfm_scst_step  eval-mode greedy baseline (or leave-one-out), train-mode multinomial samples, CIDEr-D reward, RewardCriterion,
 *                          backward; arguments as capb200_aoa_scst_step
 * The gradient table has the field layout of the weight table (every pointer is written; `pe` is a buffer and is ignored).
 * dropout = the Transformer's own rate (attention probabilities, SublayerConnections, feed-forward, positional encoding: 0.1);
 * drop_prob_lm = att_embed's Dropout.  Replayable through capb200_dropout_mask(seed, site, step, ...) with, per site, step = the position t
 * and the element index n * cols + c for decoder tensors [N, cols] (encoder tensors [B*R, cols]: step 0): 1 att_embed; 2 target embedding;
 * encoder layer l: 10+l attention probabilities [B, heads, R, R], 20+l / 40+l SublayerConnections, 30+l feed-forward hidden;
 * decoder layer l: 50+l self-attention probabilities (step 0, index ((n*heads + h)*(T+2) + t)*(T+2) + s), 60+l / 80+l / 100+l
 * SublayerConnections, 70+l source-attention probabilities [N, heads, R] at step t, 90+l feed-forward hidden.
 * The encoder runs once per image (the reference's _forward runs it once per caption: same values, seq_per_img x the work).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct { float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b; } capb200_mha_grads;
typedef struct {
    capb200_mha_grads self_attn;
    float *w1_w, *w1_b, *w2_w, *w2_b;
    float *ln0_a, *ln0_b, *ln1_a, *ln1_b;
} capb200_tfm_enc_layer_grads;
typedef struct {
    capb200_mha_grads self_attn, src_attn;
    float *w1_w, *w1_b, *w2_w, *w2_b;
    float *ln0_a, *ln0_b, *ln1_a, *ln1_b, *ln2_a, *ln2_b;
} capb200_tfm_dec_layer_grads;
typedef struct {
    float *att_embed_w, *att_embed_b;
    capb200_tfm_enc_layer_grads enc[CAPB200_TFM_MAX_LAYERS];
    float *enc_norm_a, *enc_norm_b;
    capb200_tfm_dec_layer_grads dec[CAPB200_TFM_MAX_LAYERS];
    float *dec_norm_a, *dec_norm_b;
    float *lut, *pe;
    float *gen_w, *gen_b;
} capb200_tfm_grads;
typedef struct {
    int seq_per_img;
    unsigned long long seed;
    float label_smoothing, upstream, drop_prob_lm, dropout;
    const float* att_masks;        /* [B, R] or NULL */
    int keep_rows;                 /* drop_worst: > 0 keeps the `keep_rows` rows with the smallest loss (loss_wrapper.py:47-49, 75-77) */
    float* row_loss;               /* [N] or NULL */
} capb200_tfm_xe_opts;
typedef struct {
    int sample_n;
    float temperature;
    unsigned long long seed;
    float upstream;
    int baseline;                  /* CAPB200_BASELINE_GREEDY / CAPB200_BASELINE_LEAVE_ONE_OUT */
    float drop_prob_lm, dropout;
    const long long* forced_tokens;   /* [N, T] or NULL: replay these samples instead of drawing (tests) */
    const float* att_masks;
    int keep_rows;
    float* row_loss;
} capb200_tfm_scst_opts;
int capb200_tfm_xe_step(capb200_tfm_engine* e, const float* att, int B, int R, const capb200_tfm_xe_opts* opts, const long long* labels, const float* masks,
                        int label_cols, const capb200_tfm_grads* grads, float* logprobs, float* loss, void* stream);
int capb200_tfm_scst_step(capb200_tfm_engine* e, const float* att, int B, int R, const capb200_tfm_scst_opts* opts, const capb200_cider_table* table,
                          const int* refs, const int* ref_offsets, int L, const capb200_tfm_grads* grads, long long* sample_seq, long long* greedy_seq,
                          float* sample_logprobs, float* reward, float* loss, void* stream);
/* gradient groups (see capb200_engine_set_grad_events), n <= 2: 0 generator + decoder + target embedding; 1 encoder + att_embed */
int capb200_tfm_set_grad_events(capb200_tfm_engine* e, void* const* events, int n);

/* ------------------------------------------------------------------------------------------------------------------
 * Optimizer step of the training loop: utils.clip_gradient(optimizer, grad_clip_value) (captioning/utils/misc.py:156-160, called at
 * tools/train.py:193) + torch.optim.Adam.step() (built by build_optimizer, misc.py:186-205; tools/train.py:196) in ONE launch.
 *   table  [n_tensors][4] device pointers {param, grad, exp_avg, exp_avg_sq} (fp32, contiguous), itself in device memory
 *   numel  [n_tensors] element counts (device)
 *   chunks [n_chunks][2] int32 {tensor index, chunk index}; a chunk is capb200_adam_chunk_elems() elements (device)
 *   step   the 1-based step count AFTER this update (bias corrections 1 - beta^step);  clip_value <= 0 disables the clamp;
 *   write_clamped != 0 stores the clamped gradient back (clip_gradient's in-place side effect).  weight_decay is Adam's L2 term.
 * ---------------------------------------------------------------------------------------------------------------- */
int capb200_adam_chunk_elems(void);
int capb200_adam_step(const unsigned long long* table, const long long* numel, const int* chunks, int n_chunks, double lr, double beta1, double beta2,
                      double eps, double weight_decay, long step, double clip_value, int write_clamped, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAPB200_H_ */
