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
    char* tape = nullptr;          // SCST training tape (owned, grown on demand)
    size_t tape_bytes = 0;
    Tf32Context* tf32 = nullptr;   // tensor maps + transposed operands of the training GEMMs (tensor-core modes)
    cudaEvent_t grad_events[10] = {};   // caller-owned: recorded when a gradient group is complete (capb200_aoa_set_grad_events)
    cudaStream_t side = nullptr;        // the greedy baseline of the SCST step runs here, concurrently with the sampling forward
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    StepGraph sg;                       // CUDA graph of the whole SCST step (engine_common.cuh)
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
    CAPB_NVTX("capb200 aoa prepare_feature (att_embed, refiner, ctx2att)");
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
    CAPB_CHECK_RANGE();
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
    if (e->d.loop_exec) cudaGraphExecDestroy(e->d.loop_exec);
    cudaFree(e->d.slab);
    cudaFree(e->tape);
    e->sg.destroy();
    tf32_context_destroy(e->tf32);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->side) cudaStreamDestroy(e->side);
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
    if (e->tc && !e->bound) {        // first binding only: a re-binding must not stall the training loop (see capb200_engine_bind_weights)
        CAPB_CHECK_CUDA(cudaStreamSynchronize(st));
        CAPB_CHECK_RANGE();
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
                              done_raw, core, &e->launches, st, loop_graph_key(e->ws, e->wblock, mask, R, 9), to_edits(opts->edits), opts->temperature);
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
    CAPB_REQUIRE(n >= 1 && method >= 0 && method <= 5, "bad sampling options");
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
                                core, &e->launches, st, to_edits(opts->edits), opts->top);
}

}  // extern "C"

// =====================================================================================================================
// SCST training step (AoANet)
// =====================================================================================================================
namespace {

constexpr int NL = CAPB200_AOA_REFINER_LAYERS;

struct ATape {
    float *x[NL + 1], *ln[NL], *qkv[NL], *ratt[NL], *catd[NL], *t[NL], *g[NL];
    float *att_e, *mean, *kv;
    int* tok;
    float *xt, *x1c, *gates, *c, *h, *qln, *qp, *probs, *att, *t2, *out, *outd;
    float *DL, *dOUTD, *D_T2, *D_QP, *DG, *dctx, *dh, *dc, *dX2, *d_qln, *d_hatt, *dxt, *d_x1c, *S, *d_mean, *d_kv, *d_att_e, *d_x, *d_g, *d_t,
        *d_catd, *d_qkv, *d_ln, *dpre, *stats, *mask_sum, *skinny, *glp, *item_loss;
    size_t skinny_floats;
    double* scores;
    int *s_tokens, *s_unfinished, *s_forced;   // sampling-loop state (own copies: the greedy baseline runs concurrently on the decode workspace)
    float *row_loss, *row_msum, *row_coef;     // drop_worst
};

void layout_atape(ATape& tp, Arena& a, int B, int R, int N, int T, int E, int H, int heads, int V1) {
    const long BR = (long)B * R, TN = (long)T * N;
    for (int l = 0; l <= NL; ++l) tp.x[l] = a.take<float>(BR * H);
    for (int l = 0; l < NL; ++l) {
        tp.ln[l] = a.take<float>(BR * H); tp.qkv[l] = a.take<float>(BR * 3 * H); tp.ratt[l] = a.take<float>(BR * H);
        tp.catd[l] = a.take<float>(BR * 2 * H); tp.t[l] = a.take<float>(BR * 2 * H); tp.g[l] = a.take<float>(BR * H);
    }
    tp.att_e = a.take<float>(BR * H); tp.mean = a.take<float>((long)B * H); tp.kv = a.take<float>(BR * 2 * H);
    tp.tok = a.take<int>(TN);
    tp.xt = a.take<float>(TN * E); tp.x1c = a.take<float>(TN * H); tp.gates = a.take<float>(TN * 4 * H); tp.c = a.take<float>(TN * H);
    tp.h = a.take<float>(TN * H); tp.qln = a.take<float>(TN * H); tp.qp = a.take<float>(TN * H); tp.probs = a.take<float>(TN * heads * R);
    tp.att = a.take<float>(TN * H); tp.t2 = a.take<float>(TN * 2 * H); tp.out = a.take<float>(TN * H); tp.outd = a.take<float>(TN * H);
    tp.DL = a.take<float>(TN * V1); tp.dOUTD = a.take<float>(TN * H); tp.D_T2 = a.take<float>(TN * 2 * H); tp.D_QP = a.take<float>(TN * H);
    tp.DG = a.take<float>(TN * 4 * H);
    const long NH = (long)N * H;
    tp.dctx = a.take<float>(NH); tp.dh = a.take<float>(NH); tp.dc = a.take<float>(NH); tp.dX2 = a.take<float>(2 * NH);
    tp.d_qln = a.take<float>(NH); tp.d_hatt = a.take<float>(NH); tp.dxt = a.take<float>((long)N * E); tp.d_x1c = a.take<float>(NH);
    tp.S = a.take<float>((long)B * 4 * H); tp.d_mean = a.take<float>((long)B * H); tp.d_kv = a.take<float>(BR * 2 * H); tp.d_att_e = a.take<float>(BR * H);
    tp.d_x = a.take<float>(BR * H); tp.d_g = a.take<float>(BR * H); tp.d_t = a.take<float>(BR * 2 * H); tp.d_catd = a.take<float>(BR * 2 * H);
    tp.d_qkv = a.take<float>(BR * 3 * H); tp.d_ln = a.take<float>(BR * H); tp.dpre = a.take<float>(BR * H);
    tp.stats = a.take<float>(2 * (BR > N ? BR : N)); tp.mask_sum = a.take<float>(8);
    tp.skinny_floats = (size_t)4 << 20;
    tp.skinny = a.take<float>((long)tp.skinny_floats);
    tp.glp = a.take<float>((long)B * T * V1);
    tp.item_loss = a.take<float>(TN);
    tp.scores = a.take<double>((long)N + B);
    tp.s_tokens = a.take<int>(N); tp.s_unfinished = a.take<int>(N); tp.s_forced = a.take<int>(N);
    tp.row_loss = a.take<float>(N); tp.row_msum = a.take<float>(N); tp.row_coef = a.take<float>(N);
}

// dW[out, in] (+)= dY[rows, out]^T * X[rows, in]
inline int wgrad_mode(int out_f, int in_f, int rows, const float* dY, long ld_dy, const float* X, long ld_x, float* G, long ld_g, int accumulate, int mode,
                      cudaStream_t st) {
    return gemm_wgrad_launch(out_f, in_f, rows, dY, ld_dy, X, ld_x, G, ld_g, accumulate, mode, st);
}

}  // namespace

namespace {

struct AoaTrainArgs {
    bool xe = false;
    int n = 1, T = 0, Tl = 0;              // rows per image, steps evaluated, log-prob columns
    float p_lm = 0.f, p_at = 0.f, p_aoa = 0.f, p_sub = 0.f, temperature = 1.f, upstream = 1.f, smoothing = 0.f;
    int ctx_drop = 0;
    unsigned long long seed = 0;
    bool greedy_baseline = true;
    const capb200_cider_table* table = nullptr;
    const int* refs = nullptr; const int* ref_offsets = nullptr; int L = 0;
    long long* sample_seq = nullptr; long long* greedy_seq = nullptr; float* reward = nullptr;
    const long long* forced = nullptr;
    const float* mask = nullptr;           // [B, R] region mask or null
    float ss_prob = 0.f;
    long long* tokens_used = nullptr;
    int keep = 0;
    float* row_loss = nullptr;
    const long long* labels = nullptr; long ld_labels = 0; const float* masks = nullptr; long ld_masks = 0;
    float* logprobs = nullptr; float* loss = nullptr;
};

int aoa_train_step(capb200_aoa_engine* e, const float* att, int B, int R, const AoaTrainArgs& ta, const capb200_aoa_grads* grads, cudaStream_t st) {
    void* stream = static_cast<void*>(st);
    const bool greedy_baseline = !ta.xe && ta.greedy_baseline;
    const int n = ta.n, N = B * n, T = ta.T, E = e->E, H = e->H, V1 = e->V1, F = e->F, heads = e->heads, dk = e->dk;
    const float p_lm = ta.p_lm, p_at = ta.p_at, p_aoa = ta.p_aoa, p_sub = ta.p_sub;
    const float p_ctx = ta.ctx_drop ? p_lm : 0.f;
    const float keep_lm = 1.0f / (1.0f - p_lm);
    const unsigned long long seed = ta.seed;
    const capb200_aoa_weights& w = e->w;
    const capb200_aoa_grads& G = *grads;
    const long BR = (long)B * R, NH = (long)N * H, TN = (long)T * N;
    const long ld_lp = (long)ta.Tl * V1;
    float* const sample_logprobs = ta.logprobs;
    long long* const sample_seq = ta.sample_seq;
    long long* const greedy_seq = ta.greedy_seq;
    float* const reward = ta.reward;
    float* const loss = ta.loss;

    {
        Arena dry; ATape t0; layout_atape(t0, dry, B, R, N, T, E, H, heads, V1);
        if (dry.off + 256 > e->tape_bytes) {
            CAPB_CHECK_CUDA(cudaStreamSynchronize(st));
            if (e->tape) CAPB_CHECK_CUDA(cudaFree(e->tape));
            e->tape = nullptr;
            CAPB_CHECK_CUDA(cudaMalloc(&e->tape, dry.off + 256));
            e->tape_bytes = dry.off + 256;
        }
    }
    Arena ar; ar.base = e->tape;
    ATape tp; layout_atape(tp, ar, B, R, N, T, E, H, heads, V1);

    // ---- (1) greedy baseline, eval mode: the regular decode path, on a side stream (joins before the reward): it and the train-mode
    // sampling forward are independent chains of small latency-bound kernels
    if (ensure_workspace(e, B, N, R, 1, st)) return 1;         // decode workspace sized before anything is in flight
    bool greedy_on_side = false;
    cudaStream_t gs_enqueue = st;
    capb200_sample_opts so;
    if (greedy_baseline) {
        CAPB_NVTX("capb200 aoa scst: greedy baseline (eval mode, side stream)");
        memset(&so, 0, sizeof(so)); so.edits.unk_col = -1; so.sample_n = 1; so.method = CAPB200_SAMPLE_GREEDY; so.temperature = 1.f; so.seed = 0; so.steps = T;
        cudaStream_t gs = st;
        static const bool serial = getenv("CAPB200_SCST_SERIAL_GREEDY") != nullptr;
        if (!serial) {
            bool ok = true;
            if (e->side == nullptr) ok = create_side_stream(&e->side) == cudaSuccess;
            if (ok && e->ev_fork == nullptr) ok = cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) == cudaSuccess;
            if (ok && e->ev_join == nullptr) ok = cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming) == cudaSuccess;
            if (ok) {
                CAPB_CHECK_CUDA(cudaEventRecord(e->ev_fork, st));
                CAPB_CHECK_CUDA(cudaStreamWaitEvent(e->side, e->ev_fork, 0));
                gs = e->side;
                greedy_on_side = true;
            } else (void)cudaGetLastError();
        }
        gs_enqueue = gs;
    }
    if (e->tc && e->tf32 == nullptr) e->tf32 = tf32_context_create();
    tf32_context_new_step(e->tf32);
    const long tf32_l0 = tf32_context_launches(e->tf32);
    Skinny sk{tp.skinny, tp.skinny_floats, e->tc ? 1 : 0, st};
    sk.ctx = e->tf32;
    auto act = [](float* p, long ld) { ActView v; v.f = p; v.hi = nullptr; v.lo = nullptr; v.ld = ld; return v; };
    const int wmode = e->tc ? 1 : 0;
    auto wgrad = [&](int out_f, int in_f, int rows, const float* dY, long ld_dy, const float* X, long ld_x, float* Gp, long ld_g, int accumulate, cudaStream_t) {
        return sk.wgrad(out_f, in_f, rows, dY, ld_dy, X, ld_x, Gp, ld_g, accumulate);
    };
    (void)wmode;

    // ---- (2) train-mode prologue: att_embed (+dropout), six refiner layers, final norm, mean pooling, ctx2att
    nvtxRangePushA("capb200 aoa train step: forward on the tape");
    if (sk.lin(att, F, w.att_embed_w, F, w.att_embed_b, tp.x[0], H, (int)BR, H, F, 0)) return 1;
    if (relu_copy_launch(tp.x[0], BR * H, act(tp.x[0], H), st)) return 1;
    if (dropout_apply_launch(tp.x[0], (int)BR, H, H, seed, 1, 0, p_lm, st)) return 1;
    if (ta.mask != nullptr) {      // pack_wrapper(att_embed): padded regions embed to exactly zero (AoAModel.py:211, AttModel.py:44-49)
        if (mask_rows_launch(act(tp.x[0], H), B, R, H, ta.mask, R, st)) return 1;
        e->launches++;
    }
    for (int l = 0; l < NL; ++l) {
        const capb200_aoa_refiner_layer& Lw = w.refiner[l];
        if (layer_norm_launch((int)BR, H, tp.x[l], H, Lw.ln_a, Lw.ln_b, 1e-6f, act(tp.ln[l], H), st)) return 1;
        if (sk.lin(tp.ln[l], H, e->r_qkv_w[l], H, e->r_qkv_b[l], tp.qkv[l], 3 * H, (int)BR, 3 * H, H, 0)) return 1;
        if (enc_attn_train_launch(B, R, heads, dk, tp.qkv[l], tp.qkv[l] + H, tp.qkv[l] + 2 * H, 3 * H, seed, 10 + l, p_at, tp.ratt[l], H, st, ta.mask, R)) return 1;
        if (cat_dropout_launch((int)BR, H, H, tp.ratt[l], H, tp.ln[l], H, tp.catd[l], 2 * H, seed, 20 + l, 0, p_aoa, st)) return 1;
        if (sk.lin(tp.catd[l], 2 * H, Lw.aoa_w, 2 * H, Lw.aoa_b, tp.t[l], 2 * H, (int)BR, 2 * H, 2 * H, 0)) return 1;
        if (glu_launch((int)BR, H, tp.t[l], 2 * H, nullptr, 0, act(tp.g[l], H), st)) return 1;
        if (add_dropout_launch((int)BR, H, tp.x[l], H, tp.g[l], H, tp.x[l + 1], H, seed, 30 + l, 0, p_sub, st)) return 1;
        e->launches += 9;
    }
    if (layer_norm_launch((int)BR, H, tp.x[NL], H, w.refiner_norm_a, w.refiner_norm_b, 1e-6f, act(tp.att_e, H), st)) return 1;
    if (masked_mean_launch(B, R, H, tp.att_e, H, ta.mask, R, act(tp.mean, H), st)) return 1;
    if (sk.lin(tp.att_e, H, w.ctx2att_w, H, w.ctx2att_b, tp.kv, 2 * H, (int)BR, 2 * H, H, 0)) return 1;
    e->launches += 8;

    // ---- the greedy baseline's ~220 launches are enqueued only now: the side stream forked at the top of the step (it does not wait for the
    // prologue), but the host needs ~0.7 ms to enqueue them, and the main stream should be busy with the prologue meanwhile, not idle
    if (greedy_baseline) {
        CAPB_CHECK_CUDA(cudaMemsetAsync(tp.glp, 0, sizeof(float) * (size_t)B * T * V1, gs_enqueue));
        CAPB_CHECK_CUDA(cudaMemsetAsync(greedy_seq, 0, sizeof(long long) * (size_t)B * T, gs_enqueue));
        if (capb200_aoa_decode_sample(e, att, ta.mask, B, R, &so, nullptr, 0, greedy_seq, tp.glp, nullptr, static_cast<void*>(gs_enqueue))) return 1;
        if (greedy_on_side) CAPB_CHECK_CUDA(cudaEventRecord(e->ev_join, e->side));
    
    }
    // ---- (3) T sampling steps with the tape
    CAPB_CHECK_CUDA(cudaMemsetAsync(tp.s_tokens, 0, sizeof(int) * N, st));
    for (int t = 0; t < T; ++t) {
        int* tok = tp.tok + (long)t * N;
        if (ta.xe) {
            if (t >= 1 && ta.ss_prob > 0.f) {      // scheduled sampling (AttModel.py:145-154)
                if (ss_select_launch(N, V1, sample_logprobs + (long)(t - 1) * V1, ld_lp, ta.labels, ta.ld_labels, t, seed, ta.ss_prob, tok, st)) return 1;
            } else if (load_token_column_launch(ta.labels, ta.ld_labels, t, N, tok, st)) return 1;
            if (ta.tokens_used != nullptr && store_token_column_launch(tok, N, ta.tokens_used, ta.Tl, t, st)) return 1;
        }
        float* xt = tp.xt + (long)t * N * E;
        float* x1c = tp.x1c + (long)t * NH;
        float* gates = tp.gates + (long)t * N * 4 * H;
        float *c_t = tp.c + (long)t * NH, *h_t = tp.h + (long)t * NH, *qln = tp.qln + (long)t * NH, *qp = tp.qp + (long)t * NH;
        float *att_t = tp.att + (long)t * NH, *t2 = tp.t2 + (long)t * 2 * NH, *out_t = tp.out + (long)t * NH;
        const float* h_prev = t ? tp.h + (long)(t - 1) * NH : nullptr;
        const float* c_prev = t ? tp.c + (long)(t - 1) * NH : nullptr;
        // one launch: the step's tokens (sampling: the previous step's draw), xt = dropout(relu(embed)), x1c = mean[img] + ctx_drop(previous context)
        if (aoa_step_inputs_launch(N, E, H, n, ta.xe ? tok : tp.s_tokens, ta.xe ? nullptr : tok, w.embed, xt, tp.mean, H, t ? tp.out + (long)(t - 1) * NH : nullptr, x1c,
                                   seed, t, p_lm, p_ctx, st)) return 1;
        {
            GemmProblem g; g.M = N; g.N = 4 * H; g.nseg = 2;
            g.seg[0].A = xt; g.seg[0].lda = E; g.seg[0].W = w.att_lstm_w_ih; g.seg[0].ldw = E + H; g.seg[0].K = E;
            g.seg[1].A = x1c; g.seg[1].lda = H; g.seg[1].W = w.att_lstm_w_ih + E; g.seg[1].ldw = E + H; g.seg[1].K = H;
            if (t) { g.seg[2].A = h_prev; g.seg[2].lda = H; g.seg[2].W = w.att_lstm_w_hh; g.seg[2].ldw = H; g.seg[2].K = H; g.nseg = 3; }
            g.epi.bias = e->bsum; g.epi.C = gates; g.epi.ldc = 4 * H;
            if (sk.gates(g)) return 1;
        }
        if (lstm_ln_launch(N, H, gates, 4 * H, c_prev, H, c_t, H, h_t, H, w.attn_norm_a, w.attn_norm_b, 1e-6f, qln, H, st)) return 1;      // cell + attention.norm
        if (sk.lin(qln, H, w.attn_q_w, H, w.attn_q_b, qp, H, N, H, H, 0)) return 1;
        // AoAModel.py:168 passes (query, value = p_att[..., :H], key = p_att[..., H:])
        if (cross_attn_train_launch(N, n, heads, dk, R, qp, H, tp.kv + H, tp.kv, 2 * H, seed, 5, t, p_at, att_t, H, tp.probs + (long)t * N * heads * R, st,
                                    ta.mask, R)) return 1;
        {
            GemmProblem g; g.M = N; g.N = 2 * H; g.nseg = 2;
            g.seg[0].A = att_t; g.seg[0].lda = H; g.seg[0].W = w.att2ctx_w; g.seg[0].ldw = 2 * H; g.seg[0].K = H;
            g.seg[1].A = h_t; g.seg[1].lda = H; g.seg[1].W = w.att2ctx_w + H; g.seg[1].ldw = 2 * H; g.seg[1].K = H;
            g.epi.bias = w.att2ctx_b; g.epi.C = t2; g.epi.ldc = 2 * H;
            if (sk.gates(g)) return 1;
        }
        float* outd = tp.outd + (long)t * H;                                   // [N][T][H]: batched logit backward
        if (glu_dropout_launch(N, H, t2, 2 * H, out_t, H, outd, (long)T * H, seed, t, p_lm, st)) return 1;
        float* logits = sample_logprobs + (long)t * V1;
        if (sk.lin(outd, (long)T * H, w.logit_w, H, w.logit_b, logits, ld_lp, N, V1, H, 0)) return 1;
        VocabStepArgs va;
        va.rows = N; va.V1 = V1; va.logits = logits; va.ld = ld_lp;
        if (!ta.xe) {
            va.select = 2; va.temperature = ta.temperature; va.seed = seed; va.step = (unsigned long long)t;
            va.unfinished = tp.s_unfinished; va.first_step = (t == 0); va.tokens_out = tp.s_tokens;
            va.seq_out = sample_seq; va.ld_seq = T; va.t = t;
            if (ta.forced != nullptr) {
                if (load_token_column_launch(ta.forced, T, t, N, tp.s_forced, st)) return 1;
                va.select = 3; va.forced = tp.s_forced;
            }
        }
        if (vocab_step_launch(va, st)) return 1;
        e->launches += 15;
    }

    // ---- (4) reward and loss
    nvtxRangePop();
    CAPB_NVTX("capb200 aoa train step: reward, loss, backward, weight gradients");
    if (ta.xe) {
        if (xe_loss_backward_launch(sample_logprobs, ld_lp, ta.labels, ta.ld_labels, ta.masks, ta.ld_masks, N, T, ta.Tl, V1, ta.smoothing, ta.upstream,
                                    tp.mask_sum, tp.item_loss, tp.DL, loss, st, ta.keep, ta.row_loss ? ta.row_loss : tp.row_loss, tp.row_msum, tp.row_coef)) return 1;
    } else {
        if (greedy_on_side) CAPB_CHECK_CUDA(cudaStreamWaitEvent(st, e->ev_join, 0));      // join: the reward needs the baseline captions
        if (cider_reward_launch(ta.table->t, sample_seq, N, greedy_baseline ? greedy_seq : nullptr, B, T, ta.refs, ta.ref_offsets, ta.L, tp.scores, reward, T, T,
                                st)) return 1;
        float* rl = ta.keep > 0 ? (ta.row_loss ? ta.row_loss : tp.row_loss) : nullptr;
        if (reward_criterion_fwd_launch(sample_logprobs, ld_lp, V1, sample_seq, reward, N, T, loss, rl, tp.mask_sum, st)) return 1;
        if (ta.keep > 0 && scst_drop_worst_launch(sample_seq, rl, N, T, ta.keep, ta.upstream, tp.row_msum, tp.row_coef, loss, st)) return 1;
        // ---- (5) backward through the decoder
        if (scst_dlogits_launch(sample_logprobs, ld_lp, sample_seq, reward, tp.mask_sum, ta.upstream, N, T, V1, tp.DL, st, ta.keep > 0 ? tp.row_coef : nullptr)) return 1;
    }
    if (sk.dgrad((int)TN, H, V1, tp.DL, V1, w.logit_w, H, tp.dOUTD, H, 0)) return 1;
    if (wgrad(V1, H, (int)TN, tp.DL, V1, tp.outd, H, G.logit_w, H, 0, st)) return 1;
    if (colsum_launch((int)TN, V1, tp.DL, V1, G.logit_b, 0, st)) return 1;
    auto group_done = [&](int k) -> int {
        if (record_group_event(e->grad_events[k], st)) return 1;
        return 0;
    };
    if (group_done(0)) return 1;                                        // logit
    CAPB_CHECK_CUDA(cudaMemsetAsync(tp.dctx, 0, sizeof(float) * NH, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(tp.dh, 0, sizeof(float) * NH, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(tp.dc, 0, sizeof(float) * NH, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(tp.d_kv, 0, sizeof(float) * BR * 2 * H, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(G.embed, 0, sizeof(float) * (size_t)V1 * E, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(G.attn_norm_a, 0, sizeof(float) * H, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(G.attn_norm_b, 0, sizeof(float) * H, st));
    for (int t = T - 1; t >= 0; --t) {
        const float* c_prev = t ? tp.c + (long)(t - 1) * NH : nullptr;
        float* d_t2 = tp.D_T2 + (long)t * 2 * NH;
        float* d_qp = tp.D_QP + (long)t * NH;
        float* dg = tp.DG + (long)t * N * 4 * H;
        // d out_t = out_drop-masked logit gradient + what step t+1 received through its (dropped) context input
        if (glu_backward_fused_launch(N, H, tp.t2 + (long)t * 2 * NH, 2 * H, tp.dOUTD + (long)t * H, (long)T * H, tp.dctx, d_t2, 2 * H, seed, t, p_lm, st)) return 1;
        if (sk.dgrad(N, 2 * H, 2 * H, d_t2, 2 * H, w.att2ctx_w, 2 * H, tp.dX2, 2 * H, 0)) return 1;             // [d att | d h_att]
        // attention: d att is the first half of dX2's rows (pitch 2H)
        if (cross_attn_backward_launch(B, n, heads, dk, R, tp.qp + (long)t * NH, H, tp.kv + H, tp.kv, 2 * H, seed, 5, t, p_at, tp.probs + (long)t * N * heads * R,
                                       tp.dX2, 2 * H, d_qp, H, tp.d_kv + H, tp.d_kv, 2 * H, st)) return 1;
        if (sk.dgrad(N, H, H, d_qp, H, w.attn_q_w, H, tp.d_qln, H, 0)) return 1;
        // d h_att = carried (from W_hh of step t+1) + att2ctx's h_att half + through the query LayerNorm
        if (ln_backward_launch(N, H, tp.h + (long)t * NH, H, w.attn_norm_a, tp.d_qln, H, 1e-6f, tp.d_hatt, H, 0, tp.stats, G.attn_norm_a, G.attn_norm_b, 1, st,
                               tp.dX2 + H, 2 * H, tp.dh, H)) return 1;
        if (lstm_cell_backward_launch(N, H, tp.gates + (long)t * N * 4 * H, c_prev, tp.c + (long)t * NH, tp.d_hatt, nullptr, 0, 0, 0, seed, 0.f, tp.dc, dg, st)) return 1;
        if (sk.dgrad(N, E, 4 * H, dg, 4 * H, w.att_lstm_w_ih, E + H, tp.dxt, E, 0)) return 1;
        if (sk.dgrad(N, H, 4 * H, dg, 4 * H, w.att_lstm_w_ih + E, E + H, tp.d_x1c, H, 0)) return 1;
        if (sk.dgrad(N, H, 4 * H, dg, 4 * H, w.att_lstm_w_hh, H, tp.dh, H, 0)) return 1;                             // carried to step t-1
        if (embed_backward_launch(N, E, tp.tok + (long)t * N, tp.xt + (long)t * N * E, tp.dxt, E, keep_lm, G.embed, st)) return 1;
        // the context input of step t is ctx_drop(out_{t-1}): its gradient flows to out_{t-1}
        if (t > 0 && dropout_copy_launch(tp.d_x1c, H, tp.dctx, H, N, H, seed, 4, (unsigned)t, p_ctx, st)) return 1;
        e->launches += 17;
    }
    // weight gradients batched over time (K = T*N rows)
    const int TN1 = (int)((long)(T - 1) * N);
    int rc = 0;
    rc |= wgrad(2 * H, H, (int)TN, tp.D_T2, 2 * H, tp.att, H, G.att2ctx_w, 2 * H, 0, st);
    rc |= wgrad(2 * H, H, (int)TN, tp.D_T2, 2 * H, tp.h, H, G.att2ctx_w + H, 2 * H, 0, st);
    rc |= colsum_launch((int)TN, 2 * H, tp.D_T2, 2 * H, G.att2ctx_b, 0, st);
    rc |= wgrad(H, H, (int)TN, tp.D_QP, H, tp.qln, H, G.attn_q_w, H, 0, st);
    rc |= colsum_launch((int)TN, H, tp.D_QP, H, G.attn_q_b, 0, st);
    rc |= wgrad(4 * H, E, (int)TN, tp.DG, 4 * H, tp.xt, E, G.att_lstm_w_ih, E + H, 0, st);
    rc |= wgrad(4 * H, H, (int)TN, tp.DG, 4 * H, tp.x1c, H, G.att_lstm_w_ih + E, E + H, 0, st);
    if (TN1 > 0) rc |= wgrad(4 * H, H, TN1, tp.DG + (long)N * 4 * H, 4 * H, tp.h, H, G.att_lstm_w_hh, H, 0, st);
    else CAPB_CHECK_CUDA(cudaMemsetAsync(G.att_lstm_w_hh, 0, sizeof(float) * 4 * H * H, st));
    rc |= colsum_launch((int)TN, 4 * H, tp.DG, 4 * H, G.att_lstm_b_ih, 0, st);
    rc |= colsum_launch((int)TN, 4 * H, tp.DG, 4 * H, G.att_lstm_b_hh, 0, st);
    // mean_feats enters every step's gate input: d mean[img] = (sum over steps and the image's rows of d gates) * W_ih[:, E:]
    rc |= per_image_sum_launch(T, N, n, 4 * H, tp.DG, tp.S, st);
    rc |= sk.dgrad(B, H, 4 * H, tp.S, 4 * H, w.att_lstm_w_ih + E, E + H, tp.d_mean, H, 0);
    if (rc) return 1;
    if (group_done(1)) return 1;                                        // decoder + embed

    // ---- (6) backward through the prologue
    rc |= sk.dgrad((int)BR, H, 2 * H, tp.d_kv, 2 * H, w.ctx2att_w, H, tp.d_att_e, H, 0);
    rc |= wgrad(2 * H, H, (int)BR, tp.d_kv, 2 * H, tp.att_e, H, G.ctx2att_w, H, 0, st);
    rc |= colsum_launch((int)BR, 2 * H, tp.d_kv, 2 * H, G.ctx2att_b, 0, st);
    rc |= mean_backward_launch(B, R, H, tp.d_mean, H, tp.d_att_e, H, st, ta.mask, R);
    rc |= ln_backward_launch((int)BR, H, tp.x[NL], H, w.refiner_norm_a, tp.d_att_e, H, 1e-6f, tp.d_x, H, 0, tp.stats, G.refiner_norm_a, G.refiner_norm_b, 0, st);
    if (rc) return 1;
    if (group_done(2)) return 1;                                        // ctx2att + refiner.norm
    for (int l = NL - 1; l >= 0; --l) {
        const capb200_aoa_refiner_layer& Lw = w.refiner[l];
        const capb200_aoa_refiner_layer_grads& Lg = G.refiner[l];
        // x[l+1] = x[l] + dropout(g): d_x carries to x[l] unchanged; d g = d_x * mask
        rc |= dropout_copy_launch(tp.d_x, H, tp.d_g, H, (int)BR, H, seed, 30 + l, 0, p_sub, st);
        rc |= glu_backward_launch((int)BR, H, tp.t[l], 2 * H, tp.d_g, H, tp.d_t, 2 * H, st);
        rc |= wgrad(2 * H, 2 * H, (int)BR, tp.d_t, 2 * H, tp.catd[l], 2 * H, Lg.aoa_w, 2 * H, 0, st);
        rc |= colsum_launch((int)BR, 2 * H, tp.d_t, 2 * H, Lg.aoa_b, 0, st);
        rc |= sk.dgrad((int)BR, 2 * H, 2 * H, tp.d_t, 2 * H, Lw.aoa_w, 2 * H, tp.d_catd, 2 * H, 0);
        rc |= dropout_apply_launch(tp.d_catd, (int)BR, 2 * H, 2 * H, seed, 20 + l, 0, p_aoa, st);               // [d attended | d ln (query half)]
        // self-attention backward needs a contiguous d attended
        CAPB_CHECK_CUDA(cudaMemcpy2DAsync(tp.d_g, sizeof(float) * H, tp.d_catd, sizeof(float) * 2 * H, sizeof(float) * H, BR, cudaMemcpyDeviceToDevice, st));
        rc |= enc_attn_backward_launch(B, R, heads, dk, tp.qkv[l], tp.qkv[l] + H, tp.qkv[l] + 2 * H, 3 * H, seed, 10 + l, p_at, tp.d_g, H, tp.d_qkv, tp.d_qkv + H,
                                       tp.d_qkv + 2 * H, 3 * H, st, ta.mask, R);
        // q | k | v gradients: one GEMM / one column reduction when the caller laid the three tensors out back to back (the Python
        // mirror's flat gradient buffer does), else three
        if (Lg.k_w == Lg.q_w + (long)H * H && Lg.v_w == Lg.k_w + (long)H * H) {
            rc |= wgrad(3 * H, H, (int)BR, tp.d_qkv, 3 * H, tp.ln[l], H, Lg.q_w, H, 0, st);
        } else {
            rc |= wgrad(H, H, (int)BR, tp.d_qkv, 3 * H, tp.ln[l], H, Lg.q_w, H, 0, st);
            rc |= wgrad(H, H, (int)BR, tp.d_qkv + H, 3 * H, tp.ln[l], H, Lg.k_w, H, 0, st);
            rc |= wgrad(H, H, (int)BR, tp.d_qkv + 2 * H, 3 * H, tp.ln[l], H, Lg.v_w, H, 0, st);
        }
        if (Lg.k_b == Lg.q_b + H && Lg.v_b == Lg.k_b + H) {
            rc |= colsum_launch((int)BR, 3 * H, tp.d_qkv, 3 * H, Lg.q_b, 0, st);
        } else {
            rc |= colsum_launch((int)BR, H, tp.d_qkv, 3 * H, Lg.q_b, 0, st);
            rc |= colsum_launch((int)BR, H, tp.d_qkv + H, 3 * H, Lg.k_b, 0, st);
            rc |= colsum_launch((int)BR, H, tp.d_qkv + 2 * H, 3 * H, Lg.v_b, 0, st);
        }
        // d ln = query half of the AoA input + through the packed q|k|v projection
        CAPB_CHECK_CUDA(cudaMemcpy2DAsync(tp.d_ln, sizeof(float) * H, tp.d_catd + H, sizeof(float) * 2 * H, sizeof(float) * H, BR, cudaMemcpyDeviceToDevice, st));
        rc |= sk.dgrad((int)BR, H, 3 * H, tp.d_qkv, 3 * H, e->r_qkv_w[l], H, tp.d_ln, H, 1);
        rc |= ln_backward_launch((int)BR, H, tp.x[l], H, Lw.ln_a, tp.d_ln, H, 1e-6f, tp.d_x, H, 1, tp.stats, Lg.ln_a, Lg.ln_b, 0, st);
        if (rc) return 1;
        if (group_done(3 + (NL - 1 - l))) return 1;                     // refiner layer l (layers finish 5 -> 0)
        e->launches += 22;
    }
    rc |= relu_dropout_backward_launch(BR * H, tp.x[0], tp.d_x, tp.dpre, keep_lm, st);
    rc |= wgrad(H, F, (int)BR, tp.dpre, H, att, F, G.att_embed_w, F, 0, st);
    rc |= colsum_launch((int)BR, H, tp.dpre, H, G.att_embed_b, 0, st);
    e->launches += tf32_context_launches(e->tf32) - tf32_l0;
    if (!rc && group_done(9)) return 1;                                 // att_embed
    return rc;
}

}  // namespace

extern "C" int capb200_aoa_scst_step(capb200_aoa_engine* e, const float* att, int B, int R, const capb200_aoa_scst_opts* opts,
                                     const capb200_cider_table* table, const int* refs, const int* ref_offsets, int L, const capb200_aoa_grads* grads,
                                     long long* sample_seq, long long* greedy_seq, float* sample_logprobs, float* reward, float* loss, void* stream) {
    if (check_ready(e)) return 1;
    CAPB_REQUIRE(opts && att && table && refs && ref_offsets && grads && sample_seq && sample_logprobs && reward && loss, "null argument");
    const bool greedy_baseline = opts->baseline == CAPB200_BASELINE_GREEDY;
    CAPB_REQUIRE(greedy_baseline || opts->baseline == CAPB200_BASELINE_LEAVE_ONE_OUT, "unknown baseline");
    CAPB_REQUIRE(!greedy_baseline || greedy_seq != nullptr, "the greedy baseline needs greedy_seq");
    const int n = opts->sample_n;
    CAPB_REQUIRE(n >= 1 && n <= 16 && (greedy_baseline || n >= 2) && B >= 1 && R >= 1, "sample_n must be in 1..16 (>= 2 for the leave-one-out baseline)");
    const float p_lm = opts->drop_prob_lm, p_at = opts->drop_attn, p_aoa = opts->drop_aoa, p_sub = opts->drop_sublayer;
    CAPB_REQUIRE(p_lm >= 0.f && p_lm < 1.f && p_at >= 0.f && p_at < 1.f && p_aoa >= 0.f && p_aoa < 1.f && p_sub >= 0.f && p_sub < 1.f, "dropout rates must be in [0, 1)");
    AoaTrainArgs ta;
    ta.n = n; ta.T = e->T; ta.Tl = e->T; ta.p_lm = p_lm; ta.p_at = p_at; ta.p_aoa = p_aoa; ta.p_sub = p_sub; ta.temperature = opts->temperature;
    ta.upstream = opts->upstream; ta.ctx_drop = opts->ctx_drop; ta.seed = opts->seed; ta.greedy_baseline = greedy_baseline; ta.table = table;
    ta.refs = refs; ta.ref_offsets = ref_offsets; ta.L = L; ta.sample_seq = sample_seq; ta.greedy_seq = greedy_seq; ta.reward = reward;
    ta.logprobs = sample_logprobs; ta.loss = loss; ta.forced = opts->forced_tokens; ta.mask = opts->att_masks; ta.keep = opts->keep_rows; ta.row_loss = opts->row_loss;
    CAPB_REQUIRE(ta.keep >= 0 && ta.keep <= B * n, "keep_rows must be in 0..rows");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // ---- the step as ONE CUDA graph.  Its ~1100 kernels are 5-30 us each and every launch boundary costs ~2 us on the stream, the host
    // needs ~3 ms to enqueue them, and nothing about the sequence depends on data: captured the second time a configuration is seen,
    // replayed afterwards with a fresh seed (dropout.cuh: seed salt).  The features are copied to an engine-owned buffer first so that the
    // graph reads a stable address.  The gradient-group events a data-parallel caller listens to (overlapped all-reduce) become external
    // event-record nodes of the graph (record_group_event); the event handles are part of the key.  Not used when forced tokens are replayed.
    static const bool graph_with_listener = !(getenv("CAPB200_SCST_GRAPH_SYNC") != nullptr && atoi(getenv("CAPB200_SCST_GRAPH_SYNC")) == 0);
    bool listening = false;
    for (int i = 0; i < 10; ++i) listening = listening || e->grad_events[i] != nullptr;
    if (!StepGraph::enabled() || !e->tc || (listening && !graph_with_listener) || ta.forced != nullptr || e->sg.broken) {
        if (dropout_salt_set_all(0ull, st)) return 1;
        return aoa_train_step(e, att, B, R, ta, grads, st);
    }
    cudaStream_t gst = e->sg.enter(st);             // a capturable engine-owned stream, ordered after the caller's stream
    const void* srcs[2] = {att, ta.mask};
    const size_t bytes[2] = {sizeof(float) * (size_t)B * R * e->F, ta.mask ? sizeof(float) * (size_t)B * R : 0};
    size_t off[2];
    if (e->sg.stage_inputs(2, srcs, bytes, off, gst)) return 1;
    const float* att_s = reinterpret_cast<const float*>(e->sg.stage + off[0]);
    if (ta.mask) ta.mask = reinterpret_cast<const float*>(e->sg.stage + off[1]);
    unsigned long long key = 1469598103934665603ull;
    capb200_aoa_scst_opts o2 = *opts; o2.seed = 0; o2.att_masks = ta.mask;
    StepGraph::mix(key, &o2, sizeof(o2)); StepGraph::mix(key, grads, sizeof(*grads)); StepGraph::mix(key, &e->w, sizeof(e->w));
    const void* ptrs[] = {table, refs, ref_offsets, sample_seq, greedy_seq, sample_logprobs, reward, loss, e->tape, e->ws, e->wblock, e->sg.stage, gst};
    StepGraph::mix(key, ptrs, sizeof(ptrs));
    StepGraph::mix(key, e->grad_events, sizeof(e->grad_events));
    const int dims[] = {B, R, L};
    StepGraph::mix(key, dims, sizeof(dims));
    const int rc_graph = run_step_graph(e->sg, key, opts->seed, &e->launches, gst, [&]() { return aoa_train_step(e, att_s, B, R, ta, grads, gst); });
    if (e->sg.leave(st, gst)) return 1;
    return rc_graph;
}

extern "C" int capb200_aoa_set_grad_events(capb200_aoa_engine* e, void* const* events, int n) {
    CAPB_REQUIRE(e != nullptr && n >= 0 && n <= 10, "AoANet has 10 gradient groups");
    for (int i = 0; i < 10; ++i) e->grad_events[i] = (events != nullptr && i < n) ? static_cast<cudaEvent_t>(events[i]) : nullptr;
    return 0;
}

extern "C" int capb200_aoa_xe_step(capb200_aoa_engine* e, const float* att, int B, int R, const capb200_aoa_xe_opts* opts, const long long* labels,
                                   const float* masks, int label_cols, const capb200_aoa_grads* grads, float* logprobs, float* loss, void* stream) {
    if (check_ready(e)) return 1;
    CAPB_REQUIRE(opts && att && labels && masks && grads && logprobs && loss, "null argument");
    CAPB_REQUIRE(opts->seq_per_img >= 1 && opts->seq_per_img <= 16 && B >= 1 && R >= 1, "seq_per_img must be in 1..16");
    const float p_lm = opts->drop_prob_lm, p_at = opts->drop_attn, p_aoa = opts->drop_aoa, p_sub = opts->drop_sublayer;
    CAPB_REQUIRE(p_lm >= 0.f && p_lm < 1.f && p_at >= 0.f && p_at < 1.f && p_aoa >= 0.f && p_aoa < 1.f && p_sub >= 0.f && p_sub < 1.f, "dropout rates must be in [0, 1)");
    CAPB_REQUIRE(opts->label_smoothing >= 0.f && opts->label_smoothing < 1.f, "label_smoothing must be in [0, 1)");
    CAPB_REQUIRE(label_cols >= 2 && label_cols <= e->T + 2, "labels are [N, seq_length + 2] (BOS, words, EOS padding)");
    CAPB_REQUIRE(opts->steps >= 1 && opts->steps <= label_cols - 1, "steps must be in 1..label_cols-1");
    AoaTrainArgs ta;
    ta.xe = true;
    ta.n = opts->seq_per_img; ta.T = opts->steps; ta.Tl = label_cols - 1; ta.p_lm = p_lm; ta.p_at = p_at; ta.p_aoa = p_aoa; ta.p_sub = p_sub;
    ta.upstream = opts->upstream; ta.ctx_drop = opts->ctx_drop; ta.seed = opts->seed; ta.smoothing = opts->label_smoothing;
    ta.labels = labels; ta.ld_labels = label_cols; ta.masks = masks; ta.ld_masks = label_cols; ta.logprobs = logprobs; ta.loss = loss;
    ta.mask = opts->att_masks; ta.ss_prob = opts->ss_prob; ta.tokens_used = opts->tokens_used; ta.keep = opts->keep_rows; ta.row_loss = opts->row_loss;
    CAPB_REQUIRE(ta.ss_prob >= 0.f && ta.ss_prob <= 1.f, "ss_prob must be in [0, 1]");
    CAPB_REQUIRE(ta.keep >= 0 && ta.keep <= B * opts->seq_per_img, "keep_rows must be in 0..rows");
    if (dropout_salt_set_all(0ull, static_cast<cudaStream_t>(stream))) return 1;      // eager step: the seed arguments are the effective seeds
    return aoa_train_step(e, att, B, R, ta, grads, static_cast<cudaStream_t>(stream));
}
