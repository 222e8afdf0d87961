// AoANet engine (C ABI capb200_aoa_* in include/capb200.h).
//
// Reference: captioning/models/AoAModel.py
//   _prepare_feature :207-226  att_embed -> 6 AoA refiner layers (:100-126; MultiHeadedDotAttention with project_k_v=1, do_aoa=1
//                              :56-98) -> LayerNorm -> mean pooling (mean_feats) -> ctx2att (H -> 2H = K | V of the decoder attention)
//   AoA_Decoder_Core :163-186  att_lstm(cat[xt, mean + ctx_prev]) -> LayerNorm(h) -> Linear -> 8-head dot attention over the image's
//                              K | V -> GLU(Linear(cat[att, h_att])) = the new context vector, which is also the output and is
//                              carried in state[0][1]; state[1][1] is never touched
// B200 specifics: the mean-feature term of the LSTM gates is contracted once per image (row bias), the word term comes from the
// per-token gate table, the LSTM cell is applied in the GEMM epilogue (tensor-core modes), K | V are indexed per image.
#include <vector>

#include "../../include/capb200.h"
#include "common.cuh"
#include "engine_common.cuh"
#include "kernels.cuh"

using namespace capb200;

namespace capb200 {
__global__ void capb_add_vec_kernel(const float* a, const float* b, float* o, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] + b[i];
}
__global__ void capb_interleave_gates_kernel(const float* src, float* dst, int H) {      // dst[4*j+g] = src[g*H + j]
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 * H) dst[i] = src[(i & 3) * H + (i >> 2)];
}
}  // namespace capb200

struct capb200_aoa_engine {
    capb200_aoa_cfg cfg{};
    capb200_aoa_weights w{};
    int V1 = 0, E = 0, H = 0, heads = 0, dk = 0, F = 0, T = 0, mode = 0;
    bool tc = false, bound = false;
    long launches = 0;

    char* wblock = nullptr;
    float *r_qkv_w[CAPB200_AOA_REFINER_LAYERS] = {}, *r_qkv_b[CAPB200_AOA_REFINER_LAYERS] = {};
    float *bsum = nullptr, *bsum_il = nullptr;
    float* xgate = nullptr;
    long ld_xgate = 0;
    Planes p_att, p_ctx, p_logit, p_ih_x, p_ih_c, p_hh, p_q, p_a2c_a, p_a2c_h;
    Planes pr_qkv[CAPB200_AOA_REFINER_LAYERS], pr_aoa_a[CAPB200_AOA_REFINER_LAYERS], pr_aoa_q[CAPB200_AOA_REFINER_LAYERS];

    char* ws = nullptr;
    int capB = 0, capRows = 0, capR = 0, capBeam = 0;
    Planes in_att;
    Act rx, rln, rqkv, ratt, rt, att_e, mean, p_att_kv, g_mean;     // prologue activations
    Act h0_in, h0_out, ctx_in, ctx_out, xt, gates, qln, qproj, att, t2;   // decoder activations [rows, .]
    float* c0[2] = {nullptr, nullptr};
    long ld_c = 0;
    int core_cur = 0;
    DecodeBuffers d;
    std::vector<GemmTcPlan*> plans;
};

namespace {

enum Site { A_ATT = 0, A_CTX, A_GMEAN, A_LSTM, A_Q, A_A2C, A_LOGIT, A_REF /* + 2*l: qkv, aoa */, A_COUNT = A_REF + 2 * CAPB200_AOA_REFINER_LAYERS };

void destroy_plans(capb200_aoa_engine* e) {
    for (auto& p : e->plans) { if (p) gemm_tc_plan_destroy(p); p = nullptr; }
}

int gemm(capb200_aoa_engine* e, int site, GemmProblem& g, int plan_rows, cudaStream_t st) {
    e->launches++;
    return run_gemm_mode(e->mode, &e->plans[site], g, plan_rows, st);
}

void layout_weights(capb200_aoa_engine* e, Arena& a) {
    const int H = e->H, E = e->E;
    for (int l = 0; l < CAPB200_AOA_REFINER_LAYERS; ++l) { e->r_qkv_w[l] = a.take<float>((long)3 * H * H); e->r_qkv_b[l] = a.take<float>(3 * H); }
    e->bsum = a.take<float>(4 * H);
    e->bsum_il = a.take<float>(4 * H);
    e->ld_xgate = round_up(4 * H, 8);
    e->xgate = a.take<float>((long)e->V1 * e->ld_xgate);
    if (!e->tc) return;
    e->p_att = carve_planes(a, H, e->F);
    e->p_ctx = carve_planes(a, 2 * H, H);
    e->p_logit = carve_planes(a, e->V1, H);
    e->p_ih_x = carve_planes(a, 4 * H, E);
    e->p_ih_c = carve_planes(a, 4 * H, H);
    e->p_hh = carve_planes(a, 4 * H, H);
    e->p_q = carve_planes(a, H, H);
    e->p_a2c_a = carve_planes(a, 2 * H, H);
    e->p_a2c_h = carve_planes(a, 2 * H, H);
    for (int l = 0; l < CAPB200_AOA_REFINER_LAYERS; ++l) {
        e->pr_qkv[l] = carve_planes(a, 3 * H, H);
        e->pr_aoa_a[l] = carve_planes(a, 2 * H, H);
        e->pr_aoa_q[l] = carve_planes(a, 2 * H, H);
    }
}

void layout_workspace(capb200_aoa_engine* e, Arena& a, int B, int rows, int R, int beam) {
    const int H = e->H, E = e->E, T = e->T;
    const bool tc = e->tc;
    const long BR = (long)B * R;
    if (tc) e->in_att = carve_planes(a, BR, e->F);
    e->rx.carve(a, BR, H, false);
    e->rln.carve(a, BR, H, tc);
    e->rqkv.carve(a, BR, 3 * H, false);
    e->ratt.carve(a, BR, H, tc);
    e->rt.carve(a, BR, 2 * H, false);
    e->att_e.carve(a, BR, H, tc);
    e->mean.carve(a, B, H, tc);
    e->p_att_kv.carve(a, BR, 2 * H, false);
    e->g_mean.carve(a, B, 4 * H, false);
    e->h0_in.carve(a, rows, H, tc);
    e->h0_out.carve(a, rows, H, tc);
    e->ctx_in.carve(a, rows, H, tc);
    e->ctx_out.carve(a, rows, H, tc);
    e->xt.carve(a, rows, E, tc);
    e->gates.carve(a, rows, 4 * H, false);
    e->qln.carve(a, rows, H, tc);
    e->qproj.carve(a, rows, H, false);
    e->att.carve(a, rows, H, tc);
    e->t2.carve(a, rows, 2 * H, false);
    e->ld_c = round_up(H, 8);
    for (int i = 0; i < 2; ++i) e->c0[i] = a.take<float>((long)rows * e->ld_c);
    e->d.carve(a, B, rows, beam, T);
}

int ensure_workspace(capb200_aoa_engine* e, int B, int rows, int R, int beam, cudaStream_t st) {
    if (B <= e->capB && rows <= e->capRows && R <= e->capR && beam <= e->capBeam && e->ws != nullptr) return 0;
    const int nB = B > e->capB ? B : e->capB, nRows = rows > e->capRows ? rows : e->capRows;
    const int nR = R > e->capR ? R : e->capR, nBeam = beam > e->capBeam ? beam : e->capBeam;
    Arena dry;
    layout_workspace(e, dry, nB, nRows, nR, nBeam);
    const size_t need = dry.off + 256;
    CAPB_CHECK_CUDA(cudaStreamSynchronize(st));
    destroy_plans(e);
    if (e->ws) CAPB_CHECK_CUDA(cudaFree(e->ws));
    e->ws = nullptr;
    CAPB_CHECK_CUDA(cudaMalloc(&e->ws, need));
    Arena real;
    real.base = e->ws;
    layout_workspace(e, real, nB, nRows, nR, nBeam);
    e->capB = nB; e->capRows = nRows; e->capR = nR; e->capBeam = nBeam;
    CAPB_CHECK_CUDA(cudaMemsetAsync(e->ws, 0, need, st));
    return fill_int_launch(e->d.neg1, nRows, -1, st);
}

int pack(capb200_aoa_engine* e, const float* w, long ldw, int rows, int cols, const Planes& p, cudaStream_t st) {
    e->launches++;
    return split_planes_launch(w, ldw, rows, cols, p.hi, p.lo, p.ld, st);
}
int pack_gates(capb200_aoa_engine* e, const float* w, long ldw, int H, int cols, const Planes& p, cudaStream_t st) {
    e->launches++;
    return split_planes_interleave_launch(w, ldw, H, cols, p.hi, p.lo, p.ld, st);
}

int prepare(capb200_aoa_engine* e, const float* att, const float* mask, int B, int R, cudaStream_t st) {
    const int H = e->H, E = e->E, BR = B * R, capBR = e->capB * e->capR;
    const capb200_aoa_weights& w = e->w;
    ActView in; in.f = const_cast<float*>(att); in.ld = e->F;
    if (e->tc) {
        e->launches++;
        if (split_planes_launch(att, e->F, BR, e->F, e->in_att.hi, e->in_att.lo, e->in_att.ld, st)) return 1;
        in.hi = e->in_att.hi; in.lo = e->in_att.lo;
    }
    {
        GemmProblem g;
        g.M = BR; g.N = H; g.nseg = 1;
        g.seg[0] = seg_of(in, w.att_embed_w, e->F, e->p_att, e->F);
        g.seg[0].lda_h = e->in_att.ld;
        g.epi.bias = w.att_embed_b; g.epi.relu = 1;
        g.epi.C = e->rx.v.f; g.epi.ldc = e->rx.v.ld;
        if (gemm(e, A_ATT, g, capBR, st)) return 1;
    }
    if (mask != nullptr) { e->launches++; if (mask_rows_launch(e->rx.v, B, R, H, mask, R, st)) return 1; }
    for (int l = 0; l < CAPB200_AOA_REFINER_LAYERS; ++l) {
        const capb200_aoa_refiner_layer& L = w.refiner[l];
        e->launches++;
        if (layer_norm_launch(BR, H, e->rx.v.f, e->rx.v.ld, L.ln_a, L.ln_b, 1e-6f, e->rln.v, st)) return 1;
        {
            GemmProblem g;
            g.M = BR; g.N = 3 * H; g.nseg = 1;
            g.seg[0] = seg_of(e->rln.v, e->r_qkv_w[l], H, e->pr_qkv[l], H);
            g.epi.bias = e->r_qkv_b[l];
            g.epi.C = e->rqkv.v.f; g.epi.ldc = e->rqkv.v.ld;
            if (gemm(e, A_REF + 2 * l, g, capBR, st)) return 1;
        }
        e->launches++;
        if (enc_self_attention_launch(B, R, e->heads, e->dk, e->rqkv.v.f, e->rqkv.v.f + H, e->rqkv.v.f + 2 * H, e->rqkv.v.ld, mask, R, e->ratt.v, st)) return 1;
        {   // AoA: GLU(Linear(cat[attended, query])) with the normed layer input as the query
            GemmProblem g;
            g.M = BR; g.N = 2 * H; g.nseg = 2;
            g.seg[0] = seg_of(e->ratt.v, L.aoa_w, 2 * H, e->pr_aoa_a[l], H);
            g.seg[1] = seg_of(e->rln.v, L.aoa_w + H, 2 * H, e->pr_aoa_q[l], H);
            g.epi.bias = L.aoa_b;
            g.epi.C = e->rt.v.f; g.epi.ldc = e->rt.v.ld;
            if (gemm(e, A_REF + 2 * l + 1, g, capBR, st)) return 1;
        }
        e->launches++;
        ActView xo = e->rx.v;
        if (glu_launch(BR, H, e->rt.v.f, e->rt.v.ld, e->rx.v.f, e->rx.v.ld, xo, st)) return 1;
    }
    e->launches++;
    if (layer_norm_launch(BR, H, e->rx.v.f, e->rx.v.ld, w.refiner_norm_a, w.refiner_norm_b, 1e-6f, e->att_e.v, st)) return 1;
    e->launches++;
    if (masked_mean_launch(B, R, H, e->att_e.v.f, e->att_e.v.ld, mask, R, e->mean.v, st)) return 1;
    {   // ctx2att: K | V of the decoder attention, per image
        GemmProblem g;
        g.M = BR; g.N = 2 * H; g.nseg = 1;
        g.seg[0] = seg_of(e->att_e.v, w.ctx2att_w, H, e->p_ctx, H);
        g.epi.bias = w.ctx2att_b;
        g.epi.C = e->p_att_kv.v.f; g.epi.ldc = e->p_att_kv.v.ld;
        if (gemm(e, A_CTX, g, capBR, st)) return 1;
    }
    {   // time-invariant gate term: mean_feats * W_ih[:, E:]^T + b_ih + b_hh
        GemmProblem g;
        g.M = B; g.N = 4 * H; g.nseg = 1;
        g.seg[0] = seg_of(e->mean.v, w.att_lstm_w_ih + E, E + H, e->p_ih_c, H);
        g.epi.bias = e->tc ? e->bsum_il : e->bsum;
        g.epi.C = e->g_mean.v.f; g.epi.ldc = e->g_mean.v.ld;
        if (gemm(e, A_GMEAN, g, e->capB, st)) return 1;
    }
    return 0;
}

int core_step(capb200_aoa_engine* e, int rows, int rpi, const int* tokens, const int* src_row, float* logits, long ld_logits, int B, int R,
              const float* mask, cudaStream_t st) {
    const int H = e->H, E = e->E;
    const capb200_aoa_weights& w = e->w;
    StateCopy s0, s1;
    s0.src = e->h0_out.v.f; s0.ld_src = e->h0_out.v.ld; s0.dst = e->h0_in.v;
    s1.src = e->ctx_out.v.f; s1.ld_src = e->ctx_out.v.ld; s1.dst = e->ctx_in.v;
    e->launches++;
    if (state_gather_embed_launch(rows, tokens, src_row, w.embed, E, 0, 1, e->xt.v, H, 2, s0, s1, st)) return 1;
    const int cur = e->core_cur, nxt = cur ^ 1;
    {   // att_lstm gates: ctx_prev and h_att_prev segments + per-image mean term + per-token word term
        GemmProblem g;
        g.M = rows; g.N = 4 * H; g.nseg = 2;
        g.seg[0] = seg_of(e->ctx_in.v, w.att_lstm_w_ih + E, E + H, e->p_ih_c, H);
        g.seg[1] = seg_of(e->h0_in.v, w.att_lstm_w_hh, H, e->p_hh, H);
        g.epi.row_bias = e->g_mean.v.f; g.epi.ld_row_bias = e->g_mean.v.ld; g.epi.rows_per_group = rpi;
        if (e->tc) {
            g.epi.lstm = 1; g.epi.H = H;
            g.epi.c_prev = e->c0[cur]; g.epi.ld_cprev = e->ld_c; g.epi.src_row = src_row;
            g.epi.c_out = e->c0[nxt]; g.epi.ld_cout = e->ld_c;
            g.epi.gather_bias = e->xgate; g.epi.ld_gb = e->ld_xgate; g.epi.gather_idx = tokens;
            g.epi.h_f = e->h0_out.v.f; g.epi.h_hi = e->h0_out.v.hi; g.epi.h_lo = e->h0_out.v.lo; g.epi.ld_h = e->h0_out.v.ld;
        } else {
            g.epi.C = e->gates.v.f; g.epi.ldc = e->gates.v.ld;
        }
        if (gemm(e, A_LSTM, g, e->capRows, st)) return 1;
    }
    if (!e->tc) {
        e->launches++;
        if (lstm_pointwise_launch(rows, H, e->gates.v.f, e->gates.v.ld, src_row, e->c0[cur], e->ld_c, e->c0[nxt], e->ld_c, e->h0_out.v, e->xgate,
                                  e->ld_xgate, tokens, st)) return 1;
    }
    e->core_cur = nxt;
    // multi-head dot attention: LayerNorm(h_att) -> Linear -> heads over the image's K | V (no output layer, no AoA here)
    e->launches++;
    if (layer_norm_launch(rows, H, e->h0_out.v.f, e->h0_out.v.ld, w.attn_norm_a, w.attn_norm_b, 1e-6f, e->qln.v, st)) return 1;
    {
        GemmProblem g;
        g.M = rows; g.N = H; g.nseg = 1;
        g.seg[0] = seg_of(e->qln.v, w.attn_q_w, H, e->p_q, H);
        g.epi.bias = w.attn_q_b;
        g.epi.C = e->qproj.v.f; g.epi.ldc = e->qproj.v.ld;
        if (gemm(e, A_Q, g, e->capRows, st)) return 1;
    }
    e->launches++;
    // AoAModel.py:168 passes (query, value = p_att[..., :H], key = p_att[..., H:]): the first half of ctx2att's output is V, the second K
    if (cross_attention_launch(rows, rpi, e->heads, e->dk, R, e->qproj.v.f, e->qproj.v.ld, e->p_att_kv.v.f + H, e->p_att_kv.v.f, e->p_att_kv.v.ld, mask, R,
                               e->att.v, st)) return 1;
    {   // att2ctx: GLU(Linear(cat[att, h_att]))
        GemmProblem g;
        g.M = rows; g.N = 2 * H; g.nseg = 2;
        g.seg[0] = seg_of(e->att.v, w.att2ctx_w, 2 * H, e->p_a2c_a, H);
        g.seg[1] = seg_of(e->h0_out.v, w.att2ctx_w + H, 2 * H, e->p_a2c_h, H);
        g.epi.bias = w.att2ctx_b;
        g.epi.C = e->t2.v.f; g.epi.ldc = e->t2.v.ld;
        if (gemm(e, A_A2C, g, e->capRows, st)) return 1;
    }
    e->launches++;
    if (glu_launch(rows, H, e->t2.v.f, e->t2.v.ld, nullptr, 0, e->ctx_out.v, st)) return 1;
    GemmProblem g;
    g.M = rows; g.N = e->V1; g.nseg = 1;
    g.seg[0] = seg_of(e->ctx_out.v, w.logit_w, H, e->p_logit, H);
    g.epi.bias = w.logit_b;
    g.epi.C = logits; g.epi.ldc = ld_logits;
    (void)B;
    return gemm(e, A_LOGIT, g, e->capRows, st);
}

int check_ready(capb200_aoa_engine* e) {
    CAPB_REQUIRE(e != nullptr, "null engine");
    CAPB_REQUIRE(e->bound, "capb200_aoa_bind_weights has not been called");
    return 0;
}

}  // namespace

extern "C" {

capb200_aoa_engine* capb200_aoa_create(const capb200_aoa_cfg* c) {
    if (c == nullptr) { set_error("null cfg"); return nullptr; }
    if (c->heads < 1 || c->rnn_size % c->heads != 0) { set_error("rnn_size must be divisible by the head count"); return nullptr; }
    if (c->numeric_mode < 0 || c->numeric_mode > 2) { set_error("unknown numeric mode"); return nullptr; }
    if (c->seq_length < 1 || c->seq_length > 64) { set_error("seq_length must be in 1..64"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: the capb200 engine has no CPU fallback"); return nullptr; }
    capb200_aoa_engine* e = new capb200_aoa_engine();
    e->cfg = *c;
    e->V1 = c->vocab_size + 1; e->E = c->input_encoding_size; e->H = c->rnn_size; e->heads = c->heads; e->dk = c->rnn_size / c->heads;
    e->F = c->att_feat_size; e->T = c->seq_length; e->mode = c->numeric_mode;
    e->tc = c->numeric_mode != CAPB200_MODE_SIMT_FP32;
    e->plans.assign(A_COUNT, nullptr);
    return e;
}

void capb200_aoa_destroy(capb200_aoa_engine* e) {
    if (e == nullptr) return;
    destroy_plans(e);
    cudaFree(e->wblock);
    cudaFree(e->ws);
    cudaFree(e->d.slab);
    delete e;
}

long capb200_aoa_launch_count(const capb200_aoa_engine* e) { return e ? e->launches : 0; }

int capb200_aoa_bind_weights(capb200_aoa_engine* e, const capb200_aoa_weights* w, void* stream) {
    CAPB_REQUIRE(e != nullptr && w != nullptr, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CAPB_REQUIRE(w->embed && w->att_embed_w && w->ctx2att_w && w->att_lstm_w_ih && w->att_lstm_w_hh && w->attn_q_w && w->att2ctx_w && w->logit_w,
                 "missing AoA weights");
    e->w = *w;
    const int H = e->H, E = e->E, V1 = e->V1;
    if (e->wblock == nullptr) {
        Arena dry;
        layout_weights(e, dry);
        CAPB_CHECK_CUDA(cudaMalloc(&e->wblock, dry.off + 256));
        Arena real;
        real.base = e->wblock;
        layout_weights(e, real);
    }
    const long hh = (long)H * H;
    for (int l = 0; l < CAPB200_AOA_REFINER_LAYERS; ++l) {
        const capb200_aoa_refiner_layer& L = w->refiner[l];
        CAPB_CHECK_CUDA(cudaMemcpyAsync(e->r_qkv_w[l], L.q_w, sizeof(float) * hh, cudaMemcpyDeviceToDevice, st));
        CAPB_CHECK_CUDA(cudaMemcpyAsync(e->r_qkv_w[l] + hh, L.k_w, sizeof(float) * hh, cudaMemcpyDeviceToDevice, st));
        CAPB_CHECK_CUDA(cudaMemcpyAsync(e->r_qkv_w[l] + 2 * hh, L.v_w, sizeof(float) * hh, cudaMemcpyDeviceToDevice, st));
        CAPB_CHECK_CUDA(cudaMemcpyAsync(e->r_qkv_b[l], L.q_b, sizeof(float) * H, cudaMemcpyDeviceToDevice, st));
        CAPB_CHECK_CUDA(cudaMemcpyAsync(e->r_qkv_b[l] + H, L.k_b, sizeof(float) * H, cudaMemcpyDeviceToDevice, st));
        CAPB_CHECK_CUDA(cudaMemcpyAsync(e->r_qkv_b[l] + 2 * H, L.v_b, sizeof(float) * H, cudaMemcpyDeviceToDevice, st));
    }
    capb_add_vec_kernel<<<cdiv(4 * H, 256), 256, 0, st>>>(w->att_lstm_b_ih, w->att_lstm_b_hh, e->bsum, 4 * H);
    capb_interleave_gates_kernel<<<cdiv(4 * H, 256), 256, 0, st>>>(e->bsum, e->bsum_il, H);
    CAPB_CHECK_CUDA(cudaGetLastError());
    e->launches += 2;
    if (e->tc) {
        int rc = pack(e, w->att_embed_w, e->F, H, e->F, e->p_att, st) | pack(e, w->ctx2att_w, H, 2 * H, H, e->p_ctx, st) |
                 pack(e, w->logit_w, H, V1, H, e->p_logit, st) | pack(e, w->attn_q_w, H, H, H, e->p_q, st) |
                 pack(e, w->att2ctx_w, 2 * H, 2 * H, H, e->p_a2c_a, st) | pack(e, w->att2ctx_w + H, 2 * H, 2 * H, H, e->p_a2c_h, st) |
                 pack_gates(e, w->att_lstm_w_ih, E + H, H, E, e->p_ih_x, st) | pack_gates(e, w->att_lstm_w_ih + E, E + H, H, H, e->p_ih_c, st) |
                 pack_gates(e, w->att_lstm_w_hh, H, H, H, e->p_hh, st);
        for (int l = 0; l < CAPB200_AOA_REFINER_LAYERS; ++l) {
            rc |= pack(e, e->r_qkv_w[l], H, 3 * H, H, e->pr_qkv[l], st) | pack(e, w->refiner[l].aoa_w, 2 * H, 2 * H, H, e->pr_aoa_a[l], st) |
                  pack(e, w->refiner[l].aoa_w + H, 2 * H, 2 * H, H, e->pr_aoa_q[l], st);
        }
        if (rc) return 1;
    }
    {   // per-token gate table: relu(embed) * W_ih[:, 0:E]^T
        const long ldE = round_up(E, 8);
        char* tmp = nullptr;
        const size_t tmp_bytes = (size_t)V1 * ldE * (sizeof(float) + (e->tc ? 2 * sizeof(__half) : 0)) + 1024;
        CAPB_CHECK_CUDA(cudaMallocAsync(&tmp, tmp_bytes, st));
        CAPB_CHECK_CUDA(cudaMemsetAsync(tmp, 0, tmp_bytes, st));
        ActView ev;
        ev.ld = ldE;
        ev.f = reinterpret_cast<float*>(tmp);
        if (e->tc) { ev.hi = reinterpret_cast<__half*>(tmp + (size_t)V1 * ldE * sizeof(float)); ev.lo = ev.hi + (size_t)V1 * ldE; }
        int rc = 0;
        if (ldE == E) rc = relu_copy_launch(w->embed, (long)V1 * E, ev, st);
        else {
            for (int v = 0; v < V1 && !rc; ++v) {
                ActView rv = ev;
                rv.f += (long)v * ldE; if (rv.hi) { rv.hi += (long)v * ldE; rv.lo += (long)v * ldE; }
                rc = relu_copy_launch(w->embed + (long)v * E, E, rv, st);
            }
        }
        GemmProblem g;
        g.M = V1; g.N = 4 * H; g.nseg = 1;
        g.seg[0] = seg_of(ev, w->att_lstm_w_ih, E + H, e->p_ih_x, E);
        g.epi.C = e->xgate; g.epi.ldc = e->ld_xgate;
        if (!rc) {
            if (!e->tc) rc = gemm_simt_launch(g, st);
            else {
                GemmTcPlan* plan = gemm_tc_plan_create(g, e->mode == CAPB200_MODE_TC_F16X3 ? 3 : 1);
                rc = plan ? gemm_tc_plan_launch(plan, nullptr, 0, st) : 1;
                if (plan) gemm_tc_plan_destroy(plan);
            }
        }
        e->launches += 2;
        cudaFreeAsync(tmp, st);
        if (rc) return 1;
    }
    e->bound = true;
    return 0;
}

int capb200_aoa_decode_beam(capb200_aoa_engine* e, const float* att, const float* mask, int B, int R, const capb200_beam_opts* opts, long long* seq,
                            float* seq_logprobs, long long* done_seq, int* done_len, float* done_p, float* done_raw, void* stream) {
    if (check_ready(e)) return 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CAPB_REQUIRE(opts != nullptr && att != nullptr && seq != nullptr && B >= 1 && R >= 1, "bad argument");
    const int beam = opts->beam_size, keep = opts->sample_n;
    CAPB_REQUIRE(beam >= 1 && beam <= 16 && beam <= e->V1, "beam_size must be in 1..16 and <= V+1");
    CAPB_REQUIRE(keep == 1 || keep == beam, "sample_n must be 1 or beam_size (AttModel.py:223)");
    if (ensure_workspace(e, B, B * beam, R, beam, st)) return 1;
    if (prepare(e, att, mask, B, R, st)) return 1;
    e->core_cur = 0;
    auto core = [&](int nrows, int live, const int* tokens, const int* src_row, int /*t*/, float* logits, long ld) {
        return core_step(e, nrows, live, tokens, src_row, logits, ld, B, R, mask, st);
    };
    return beam_decode_driver(e->d, e->V1, e->T, B, beam, keep, opts->penalty_kind, opts->penalty_alpha, seq, seq_logprobs, done_seq, done_len, done_p,
                              done_raw, core, &e->launches, st);
}

int capb200_aoa_beam_record_logprobs(capb200_aoa_engine* e, int image, int rank, float* dst, void* stream) {
    if (check_ready(e)) return 1;
    return beam_record_logprobs(e->d, e->V1, e->T, image, rank, dst, static_cast<cudaStream_t>(stream));
}

int capb200_aoa_decode_sample(capb200_aoa_engine* e, const float* att, const float* mask, int B, int R, const capb200_sample_opts* opts,
                              const long long* tokens_in, long ld_tok, long long* seq, float* seq_logprobs, float* picked, void* stream) {
    if (check_ready(e)) return 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CAPB_REQUIRE(opts != nullptr && att != nullptr && seq_logprobs != nullptr && B >= 1 && R >= 1, "bad argument");
    const int n = opts->sample_n, method = opts->method;
    CAPB_REQUIRE(n >= 1 && method >= 0 && method <= 3, "bad sampling options");
    if (method == CAPB200_SAMPLE_FORCED || method == CAPB200_SAMPLE_TEACHER) CAPB_REQUIRE(tokens_in != nullptr && ld_tok >= 1, "token matrix required");
    if (method != CAPB200_SAMPLE_TEACHER) CAPB_REQUIRE(seq != nullptr, "seq output required");
    if (method == CAPB200_SAMPLE_MULTINOMIAL) CAPB_REQUIRE(opts->temperature > 0.f, "temperature must be positive");
    const int rows = B * n;
    const int steps = (method == CAPB200_SAMPLE_TEACHER) ? opts->steps : e->T;
    const long t_out = (method == CAPB200_SAMPLE_TEACHER) ? ld_tok : e->T;
    CAPB_REQUIRE(steps >= 0 && steps <= t_out, "steps out of range");
    if (ensure_workspace(e, B, rows, R, 1, st)) return 1;
    if (prepare(e, att, mask, B, R, st)) return 1;
    e->core_cur = 0;
    auto core = [&](int nrows, int /*live*/, const int* tokens, const int* src_row, int /*t*/, float* logits, long ld) {
        return core_step(e, nrows, n, tokens, src_row, logits, ld, B, R, mask, st);
    };
    return sample_decode_driver(e->d, e->V1, e->T, rows, method, opts->temperature, opts->seed, steps, tokens_in, ld_tok, seq, seq_logprobs, picked,
                                core, &e->launches, st);
}

}  // extern "C"
