// Transformer captioner engine (C ABI capb200_tfm_* in include/capb200.h).
//
// Reference: captioning/models/TransformerModel.py -- _prepare_feature :305-338 (att_embed + N_enc pre-norm encoder layers),
// core :351-363 (the reference re-runs the whole decoder over all t tokens every step; here every layer keeps a K/V cache and a
// step touches one token per row -- 20 token-layers instead of 210 per caption, same results since the decoder mask is causal),
// Generator :50-57.  Beam search / sampling bookkeeping is the shared driver of engine_common.cuh; a beam row reads its
// ancestors' cache entries through the search history, so the cache is never reordered by parent beam (the reference reorders
// its token history, CaptionModel.py:105-108).
#include <vector>

#include "../../include/capb200.h"
#include "common.cuh"
#include "engine_common.cuh"
#include "kernels.cuh"

using namespace capb200;

namespace capb200 {
const char* last_error_cstr();
}

struct capb200_tfm_engine {
    capb200_tfm_cfg cfg{};
    capb200_tfm_weights w{};
    int V1 = 0, D = 0, Dff = 0, H = 0, dk = 0, NE = 0, ND = 0, F = 0, T = 0, mode = 0;
    bool tc = false, bound = false;
    long launches = 0;

    // bind-time (owned): concatenated projections and their split planes
    char* wblock = nullptr;
    float *enc_qkv_w[CAPB200_TFM_MAX_LAYERS] = {}, *enc_qkv_b[CAPB200_TFM_MAX_LAYERS] = {};
    float *dec_qkv_w[CAPB200_TFM_MAX_LAYERS] = {}, *dec_qkv_b[CAPB200_TFM_MAX_LAYERS] = {};
    float *dec_skv_w[CAPB200_TFM_MAX_LAYERS] = {}, *dec_skv_b[CAPB200_TFM_MAX_LAYERS] = {};
    Planes p_att, p_gen;
    Planes pe_qkv[CAPB200_TFM_MAX_LAYERS], pe_o[CAPB200_TFM_MAX_LAYERS], pe_w1[CAPB200_TFM_MAX_LAYERS], pe_w2[CAPB200_TFM_MAX_LAYERS];
    Planes pd_qkv[CAPB200_TFM_MAX_LAYERS], pd_o[CAPB200_TFM_MAX_LAYERS], pd_qs[CAPB200_TFM_MAX_LAYERS], pd_skv[CAPB200_TFM_MAX_LAYERS],
        pd_os[CAPB200_TFM_MAX_LAYERS], pd_w1[CAPB200_TFM_MAX_LAYERS], pd_w2[CAPB200_TFM_MAX_LAYERS];

    // workspace (owned)
    char* ws = nullptr;
    int capB = 0, capRows = 0, capR = 0, capBeam = 0;
    Planes in_att;
    Act ex, eln, eqkv, eatt, eh, mem;            // encoder activations [B*R, .]
    float* skv[CAPB200_TFM_MAX_LAYERS] = {};     // per decoder layer [B*R, 2D]: K | V of the memory
    Act x, ln, qkv, att, qs, hh;                 // decoder activations [rows, .]
    float *kc[CAPB200_TFM_MAX_LAYERS] = {}, *vc[CAPB200_TFM_MAX_LAYERS] = {};   // [T][rows][D]
    long cache_step_stride = 0;
    DecodeBuffers d;
    std::vector<GemmTcPlan*> plans;

    // training steps (capb200_tfm_xe_step / capb200_tfm_scst_step)
    char* tape = nullptr;
    size_t tape_bytes = 0;
    Tf32Context* tf32 = nullptr;
    cudaEvent_t grad_events[2] = {};
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    StepGraph sg;                       // CUDA graph of the whole SCST step (engine_common.cuh)
};

namespace {

enum Site { S_ATT = 0, S_GEN = 1, S_ENC = 2 /* + 4*l: qkv,o,w1,w2 */, S_SKV = 2 + 4 * CAPB200_TFM_MAX_LAYERS /* + l */,
            S_DEC = S_SKV + CAPB200_TFM_MAX_LAYERS /* + 6*l: qkv,o,qs,os,w1,w2 */, S_COUNT = S_DEC + 6 * CAPB200_TFM_MAX_LAYERS };

void destroy_plans(capb200_tfm_engine* e) {
    for (auto& p : e->plans) { if (p) gemm_tc_plan_destroy(p); p = nullptr; }
}

int gemm(capb200_tfm_engine* e, int site, GemmProblem& g, int plan_rows, cudaStream_t st) {
    e->launches++;
    return run_gemm_mode(e->mode, &e->plans[site], g, plan_rows, st);
}

// y = act(x * W^T + b) (+ residual); x given as an ActView, W as fp32 pointer + planes
int linear(capb200_tfm_engine* e, int site, const ActView& x, int M, int K, const float* w, const Planes& wp, const float* b, int N, ActView out,
           bool relu, const float* residual, long ld_res, int plan_rows, cudaStream_t st) {
    GemmProblem g;
    g.M = M; g.N = N; g.nseg = 1;
    g.seg[0] = seg_of(x, w, K, wp, K);
    g.epi.bias = b; g.epi.relu = relu ? 1 : 0;
    g.epi.residual = residual; g.epi.ld_res = ld_res;
    g.epi.C = out.f; g.epi.ldc = out.ld; g.epi.C_hi = out.hi; g.epi.C_lo = out.lo; g.epi.ldcs = out.ld;
    return gemm(e, site, g, plan_rows, st);
}

void layout_weights(capb200_tfm_engine* e, Arena& a) {
    const int D = e->D, Dff = e->Dff;
    for (int l = 0; l < e->NE; ++l) { e->enc_qkv_w[l] = a.take<float>((long)3 * D * D); e->enc_qkv_b[l] = a.take<float>(3 * D); }
    for (int l = 0; l < e->ND; ++l) {
        e->dec_qkv_w[l] = a.take<float>((long)3 * D * D); e->dec_qkv_b[l] = a.take<float>(3 * D);
        e->dec_skv_w[l] = a.take<float>((long)2 * D * D); e->dec_skv_b[l] = a.take<float>(2 * D);
    }
    if (!e->tc) return;
    e->p_att = carve_planes(a, D, e->F);
    e->p_gen = carve_planes(a, e->V1, D);
    for (int l = 0; l < e->NE; ++l) {
        e->pe_qkv[l] = carve_planes(a, 3 * D, D); e->pe_o[l] = carve_planes(a, D, D);
        e->pe_w1[l] = carve_planes(a, Dff, D); e->pe_w2[l] = carve_planes(a, D, Dff);
    }
    for (int l = 0; l < e->ND; ++l) {
        e->pd_qkv[l] = carve_planes(a, 3 * D, D); e->pd_o[l] = carve_planes(a, D, D); e->pd_qs[l] = carve_planes(a, D, D);
        e->pd_skv[l] = carve_planes(a, 2 * D, D); e->pd_os[l] = carve_planes(a, D, D);
        e->pd_w1[l] = carve_planes(a, Dff, D); e->pd_w2[l] = carve_planes(a, D, Dff);
    }
}

void layout_workspace(capb200_tfm_engine* e, Arena& a, int B, int rows, int R, int beam) {
    const int D = e->D, Dff = e->Dff, T = e->T;
    const bool tc = e->tc;
    const long BR = (long)B * R;
    if (tc) e->in_att = carve_planes(a, BR, e->F);
    e->ex.carve(a, BR, D, false);
    e->eln.carve(a, BR, D, tc);
    e->eqkv.carve(a, BR, 3 * D, false);
    e->eatt.carve(a, BR, D, tc);
    e->eh.carve(a, BR, Dff, tc);
    e->mem.carve(a, BR, D, tc);
    for (int l = 0; l < e->ND; ++l) e->skv[l] = a.take<float>(BR * 2 * D);
    e->x.carve(a, rows, D, false);
    e->ln.carve(a, rows, D, tc);
    e->qkv.carve(a, rows, 3 * D, false);
    e->att.carve(a, rows, D, tc);
    e->qs.carve(a, rows, D, false);
    e->hh.carve(a, rows, Dff, tc);
    e->cache_step_stride = (long)rows * D;
    for (int l = 0; l < e->ND; ++l) {
        e->kc[l] = a.take<float>((long)(T + 1) * rows * D);      // T + 1 positions: teacher forcing feeds bos + T labels
        e->vc[l] = a.take<float>((long)(T + 1) * rows * D);
    }
    e->d.carve(a, B, rows, beam, T);
}

int ensure_workspace(capb200_tfm_engine* e, int B, int rows, int R, int beam, cudaStream_t st) {
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

int pack(capb200_tfm_engine* e, const float* w, int rows, int cols, const Planes& p, cudaStream_t st) {
    e->launches++;
    return split_planes_launch(w, cols, rows, cols, p.hi, p.lo, p.ld, st);
}

int concat_rows(float* dst, const float* a, const float* b, const float* c, long n_each, cudaStream_t st) {
    CAPB_CHECK_CUDA(cudaMemcpyAsync(dst, a, sizeof(float) * n_each, cudaMemcpyDeviceToDevice, st));
    CAPB_CHECK_CUDA(cudaMemcpyAsync(dst + n_each, b, sizeof(float) * n_each, cudaMemcpyDeviceToDevice, st));
    if (c) CAPB_CHECK_CUDA(cudaMemcpyAsync(dst + 2 * n_each, c, sizeof(float) * n_each, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// _prepare_feature: att_embed (+ReLU), N_enc pre-norm encoder layers, final LayerNorm, then K/V of every decoder layer's src_attn
int prepare(capb200_tfm_engine* e, const float* att, const float* mask, int B, int R, cudaStream_t st) {
    const int D = e->D, Dff = e->Dff, BR = B * R, capBR = e->capB * e->capR;
    const capb200_tfm_weights& w = e->w;
    ActView in; in.f = const_cast<float*>(att); in.ld = e->F;
    if (e->tc) {
        e->launches++;
        if (split_planes_launch(att, e->F, BR, e->F, e->in_att.hi, e->in_att.lo, e->in_att.ld, st)) return 1;
        in.hi = e->in_att.hi; in.lo = e->in_att.lo;
        // the planes have their own pitch: route through an explicit segment below
    }
    {
        GemmProblem g;
        g.M = BR; g.N = D; g.nseg = 1;
        g.seg[0] = seg_of(in, w.att_embed_w, e->F, e->p_att, e->F);
        g.seg[0].lda_h = e->in_att.ld;
        g.epi.bias = w.att_embed_b; g.epi.relu = 1;
        g.epi.C = e->ex.v.f; g.epi.ldc = e->ex.v.ld;
        if (gemm(e, S_ATT, g, capBR, st)) return 1;
    }
    if (mask != nullptr) { e->launches++; if (mask_rows_launch(e->ex.v, B, R, D, mask, R, st)) return 1; }
    for (int l = 0; l < e->NE; ++l) {
        const capb200_tfm_enc_layer& L = w.enc[l];
        e->launches++;
        if (layer_norm_launch(BR, D, e->ex.v.f, e->ex.v.ld, L.ln0_a, L.ln0_b, 1e-6f, e->eln.v, st)) return 1;
        if (linear(e, S_ENC + 4 * l, e->eln.v, BR, D, e->enc_qkv_w[l], e->pe_qkv[l], e->enc_qkv_b[l], 3 * D, e->eqkv.v, false, nullptr, 0, capBR, st)) return 1;
        e->launches++;
        if (enc_self_attention_launch(B, R, e->H, e->dk, e->eqkv.v.f, e->eqkv.v.f + D, e->eqkv.v.f + 2 * D, e->eqkv.v.ld, mask, R, e->eatt.v, st)) return 1;
        ActView xo = e->ex.v; xo.hi = xo.lo = nullptr;
        if (linear(e, S_ENC + 4 * l + 1, e->eatt.v, BR, D, L.self_attn.o_w, e->pe_o[l], L.self_attn.o_b, D, xo, false, e->ex.v.f, e->ex.v.ld, capBR, st)) return 1;
        e->launches++;
        if (layer_norm_launch(BR, D, e->ex.v.f, e->ex.v.ld, L.ln1_a, L.ln1_b, 1e-6f, e->eln.v, st)) return 1;
        if (linear(e, S_ENC + 4 * l + 2, e->eln.v, BR, D, L.w1_w, e->pe_w1[l], L.w1_b, Dff, e->eh.v, true, nullptr, 0, capBR, st)) return 1;
        if (linear(e, S_ENC + 4 * l + 3, e->eh.v, BR, Dff, L.w2_w, e->pe_w2[l], L.w2_b, D, xo, false, e->ex.v.f, e->ex.v.ld, capBR, st)) return 1;
    }
    e->launches++;
    if (layer_norm_launch(BR, D, e->ex.v.f, e->ex.v.ld, w.enc_norm_a, w.enc_norm_b, 1e-6f, e->mem.v, st)) return 1;
    for (int l = 0; l < e->ND; ++l) {
        ActView o; o.f = e->skv[l]; o.ld = 2 * D;
        if (linear(e, S_SKV + l, e->mem.v, BR, D, e->dec_skv_w[l], e->pd_skv[l], e->dec_skv_b[l], 2 * D, o, false, nullptr, 0, capBR, st)) return 1;
    }
    return 0;
}

// one decoder step for `rows` rows at position t
int core_step(capb200_tfm_engine* e, int rows, int rpi, const int* tokens, const int* anc, const long long* labels, long ld_lab, int t, float* logits,
              long ld_logits, int R, const float* mask, cudaStream_t st) {
    const int D = e->D, Dff = e->Dff, capRows = e->capRows;
    const capb200_tfm_weights& w = e->w;
    e->launches++;
    if (embed_pe_launch(rows, D, tokens, w.lut, w.pe + (long)t * D, sqrtf((float)D), e->x.v, st)) return 1;
    ActView xo = e->x.v; xo.hi = xo.lo = nullptr;
    for (int l = 0; l < e->ND; ++l) {
        const capb200_tfm_dec_layer& L = w.dec[l];
        const int s0 = S_DEC + 6 * l;
        e->launches++;
        if (layer_norm_launch(rows, D, e->x.v.f, e->x.v.ld, L.ln0_a, L.ln0_b, 1e-6f, e->ln.v, st)) return 1;
        if (linear(e, s0, e->ln.v, rows, D, e->dec_qkv_w[l], e->pd_qkv[l], e->dec_qkv_b[l], 3 * D, e->qkv.v, false, nullptr, 0, capRows, st)) return 1;
        e->launches++;
        if (dec_self_attention_launch(rows, e->H, e->dk, t, e->qkv.v.f, e->qkv.v.ld, e->kc[l], e->vc[l], e->cache_step_stride, D, anc, e->T, labels,
                                      ld_lab, e->att.v, st)) return 1;
        if (linear(e, s0 + 1, e->att.v, rows, D, L.self_attn.o_w, e->pd_o[l], L.self_attn.o_b, D, xo, false, e->x.v.f, e->x.v.ld, capRows, st)) return 1;
        e->launches++;
        if (layer_norm_launch(rows, D, e->x.v.f, e->x.v.ld, L.ln1_a, L.ln1_b, 1e-6f, e->ln.v, st)) return 1;
        if (linear(e, s0 + 2, e->ln.v, rows, D, L.src_attn.q_w, e->pd_qs[l], L.src_attn.q_b, D, e->qs.v, false, nullptr, 0, capRows, st)) return 1;
        e->launches++;
        if (cross_attention_launch(rows, rpi, e->H, e->dk, R, e->qs.v.f, e->qs.v.ld, e->skv[l], e->skv[l] + D, 2 * D, mask, R, e->att.v, st)) return 1;
        if (linear(e, s0 + 3, e->att.v, rows, D, L.src_attn.o_w, e->pd_os[l], L.src_attn.o_b, D, xo, false, e->x.v.f, e->x.v.ld, capRows, st)) return 1;
        e->launches++;
        if (layer_norm_launch(rows, D, e->x.v.f, e->x.v.ld, L.ln2_a, L.ln2_b, 1e-6f, e->ln.v, st)) return 1;
        if (linear(e, s0 + 4, e->ln.v, rows, D, L.w1_w, e->pd_w1[l], L.w1_b, Dff, e->hh.v, true, nullptr, 0, capRows, st)) return 1;
        if (linear(e, s0 + 5, e->hh.v, rows, Dff, L.w2_w, e->pd_w2[l], L.w2_b, D, xo, false, e->x.v.f, e->x.v.ld, capRows, st)) return 1;
    }
    e->launches++;
    if (layer_norm_launch(rows, D, e->x.v.f, e->x.v.ld, w.dec_norm_a, w.dec_norm_b, 1e-6f, e->ln.v, st)) return 1;
    ActView lo; lo.f = logits; lo.ld = ld_logits;
    return linear(e, S_GEN, e->ln.v, rows, D, w.gen_w, e->p_gen, w.gen_b, e->V1, lo, false, nullptr, 0, capRows, st);
}

int check_ready(capb200_tfm_engine* e) {
    CAPB_REQUIRE(e != nullptr, "null engine");
    CAPB_REQUIRE(e->bound, "capb200_tfm_bind_weights has not been called");
    CAPB_CHECK_RANGE();
    return 0;
}

}  // namespace

extern "C" {

capb200_tfm_engine* capb200_tfm_create(const capb200_tfm_cfg* c) {
    if (c == nullptr) { set_error("null cfg"); return nullptr; }
    if (c->n_enc < 0 || c->n_enc > CAPB200_TFM_MAX_LAYERS || c->n_dec < 1 || c->n_dec > CAPB200_TFM_MAX_LAYERS) { set_error("layer count must be within 1..8"); return nullptr; }
    if (c->heads < 1 || c->d_model % c->heads != 0) { set_error("d_model must be divisible by the head count"); return nullptr; }
    if (c->numeric_mode < 0 || c->numeric_mode > 2) { set_error("unknown numeric mode"); return nullptr; }
    if (c->seq_length < 1 || c->seq_length > 31) { set_error("seq_length must be in 1..31 on the transformer path"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: the capb200 engine has no CPU fallback"); return nullptr; }
    capb200_tfm_engine* e = new capb200_tfm_engine();
    e->cfg = *c;
    e->V1 = c->vocab_size + 1; e->D = c->d_model; e->Dff = c->d_ff; e->H = c->heads; e->dk = c->d_model / c->heads;
    e->NE = c->n_enc; e->ND = c->n_dec; e->F = c->att_feat_size; e->T = c->seq_length; e->mode = c->numeric_mode;
    e->tc = c->numeric_mode != CAPB200_MODE_SIMT_FP32;
    e->plans.assign(S_COUNT, nullptr);
    return e;
}

void capb200_tfm_destroy(capb200_tfm_engine* e) {
    if (e == nullptr) return;
    destroy_plans(e);
    cudaFree(e->wblock);
    cudaFree(e->ws);
    cudaFree(e->tape);
    e->sg.destroy();
    tf32_context_destroy(e->tf32);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->side) cudaStreamDestroy(e->side);
    if (e->d.loop_exec) cudaGraphExecDestroy(e->d.loop_exec);
    cudaFree(e->d.slab);
    delete e;
}

long capb200_tfm_launch_count(const capb200_tfm_engine* e) { return e ? e->launches : 0; }

int capb200_tfm_bind_weights(capb200_tfm_engine* e, const capb200_tfm_weights* w, void* stream) {
    CAPB_REQUIRE(e != nullptr && w != nullptr, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CAPB_REQUIRE(w->att_embed_w && w->att_embed_b && w->lut && w->pe && w->gen_w && w->gen_b && w->dec_norm_a && w->dec_norm_b, "missing weights");
    e->w = *w;
    const int D = e->D, Dff = e->Dff;
    if (e->wblock == nullptr) {
        Arena dry;
        layout_weights(e, dry);
        CAPB_CHECK_CUDA(cudaMalloc(&e->wblock, dry.off + 256));
        Arena real;
        real.base = e->wblock;
        layout_weights(e, real);
    }
    const long dd = (long)D * D;
    for (int l = 0; l < e->NE; ++l) {
        const capb200_mha_weights& a = w->enc[l].self_attn;
        if (concat_rows(e->enc_qkv_w[l], a.q_w, a.k_w, a.v_w, dd, st) || concat_rows(e->enc_qkv_b[l], a.q_b, a.k_b, a.v_b, D, st)) return 1;
    }
    for (int l = 0; l < e->ND; ++l) {
        const capb200_mha_weights& a = w->dec[l].self_attn;
        const capb200_mha_weights& s = w->dec[l].src_attn;
        if (concat_rows(e->dec_qkv_w[l], a.q_w, a.k_w, a.v_w, dd, st) || concat_rows(e->dec_qkv_b[l], a.q_b, a.k_b, a.v_b, D, st)) return 1;
        if (concat_rows(e->dec_skv_w[l], s.k_w, s.v_w, nullptr, dd, st) || concat_rows(e->dec_skv_b[l], s.k_b, s.v_b, nullptr, D, st)) return 1;
    }
    if (e->tc) {
        int rc = pack(e, w->att_embed_w, D, e->F, e->p_att, st) | pack(e, w->gen_w, e->V1, D, e->p_gen, st);
        for (int l = 0; l < e->NE; ++l) {
            rc |= pack(e, e->enc_qkv_w[l], 3 * D, D, e->pe_qkv[l], st) | pack(e, w->enc[l].self_attn.o_w, D, D, e->pe_o[l], st);
            rc |= pack(e, w->enc[l].w1_w, Dff, D, e->pe_w1[l], st) | pack(e, w->enc[l].w2_w, D, Dff, e->pe_w2[l], st);
        }
        for (int l = 0; l < e->ND; ++l) {
            rc |= pack(e, e->dec_qkv_w[l], 3 * D, D, e->pd_qkv[l], st) | pack(e, w->dec[l].self_attn.o_w, D, D, e->pd_o[l], st);
            rc |= pack(e, w->dec[l].src_attn.q_w, D, D, e->pd_qs[l], st) | pack(e, e->dec_skv_w[l], 2 * D, D, e->pd_skv[l], st);
            rc |= pack(e, w->dec[l].src_attn.o_w, D, D, e->pd_os[l], st);
            rc |= pack(e, w->dec[l].w1_w, Dff, D, e->pd_w1[l], st) | pack(e, w->dec[l].w2_w, D, Dff, e->pd_w2[l], st);
        }
        if (rc) return 1;
        if (!e->bound) {        // first binding only: a re-binding must not stall the training loop (see capb200_engine_bind_weights)
            CAPB_CHECK_CUDA(cudaStreamSynchronize(st));
            CAPB_CHECK_RANGE();
        }
    }
    e->bound = true;
    return 0;
}

int capb200_tfm_decode_beam(capb200_tfm_engine* e, const float* att, const float* mask, int B, int R, const capb200_beam_opts* opts, long long* seq,
                            float* seq_logprobs, long long* done_seq, int* done_len, float* done_p, float* done_raw, void* stream) {
    if (check_ready(e)) return 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CAPB_REQUIRE(opts != nullptr && att != nullptr && seq != nullptr && B >= 1 && R >= 1, "bad argument");
    const int beam = opts->beam_size, keep = opts->sample_n;
    CAPB_REQUIRE(beam >= 1 && beam <= 16 && beam <= e->V1, "beam_size must be in 1..16 and <= V+1");
    CAPB_REQUIRE(keep == 1 || keep == beam, "sample_n must be 1 or beam_size (AttModel.py:223)");
    if (ensure_workspace(e, B, B * beam, R, beam, st)) return 1;
    if (prepare(e, att, mask, B, R, st)) return 1;
    auto core = [&](int nrows, int live, const int* tokens, const int* /*src_row*/, int t, float* logits, long ld) {
        const int* anc = (t == 0) ? nullptr : beam_ancestors(e->d.bs, t);
        return core_step(e, nrows, live, tokens, anc, nullptr, 0, t, logits, ld, R, mask, st);
    };
    return beam_decode_driver(e->d, e->V1, e->T, B, beam, keep, opts->penalty_kind, opts->penalty_alpha, seq, seq_logprobs, done_seq, done_len, done_p,
                              done_raw, core, &e->launches, st, loop_graph_key(e->ws, e->wblock, mask, R, 7), to_edits(opts->edits), opts->temperature);
}

int capb200_tfm_beam_record_logprobs(capb200_tfm_engine* e, int image, int rank, float* dst, void* stream) {
    if (check_ready(e)) return 1;
    return beam_record_logprobs(e->d, e->V1, e->T, image, rank, dst, static_cast<cudaStream_t>(stream));
}

int capb200_tfm_decode_sample(capb200_tfm_engine* e, const float* att, const float* mask, int B, int R, const capb200_sample_opts* opts,
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
    CAPB_REQUIRE(steps >= 0 && steps <= t_out && steps <= e->T + 1 && steps <= 31, "steps out of range");
    if (ensure_workspace(e, B, rows, R, 1, st)) return 1;
    if (prepare(e, att, mask, B, R, st)) return 1;
    const long long* labels = (method == CAPB200_SAMPLE_TEACHER) ? tokens_in : nullptr;
    auto core = [&](int nrows, int /*live*/, const int* tokens, const int* /*src_row*/, int t, float* logits, long ld) {
        return core_step(e, nrows, n, tokens, nullptr, labels, ld_tok, t, logits, ld, R, mask, st);
    };
    return sample_decode_driver(e->d, e->V1, e->T, rows, method, opts->temperature, opts->seed, steps, tokens_in, ld_tok, seq, seq_logprobs, picked,
                                core, &e->launches, st, to_edits(opts->edits), opts->top);
}

}  // extern "C"


// =====================================================================================================================
// Training steps of the Transformer captioner.
//
// Reference: LossWrapper.forward (captioning/modules/loss_wrapper.py:25-73) over TransformerModel -- the XE branch calls
// TransformerModel._forward (TransformerModel.py:340-348: ONE teacher-forced pass over all positions, seq_mask = pad/eos keys masked +
// subsequent mask, :319-328) and LanguageModelCriterion / LabelSmoothing; the sc branch samples in train mode through core (:351-363,
// which re-runs the whole prefix every step -- causal, so the result equals a K/V-cached step) and applies RewardCriterion; then
// loss.backward() (tools/train.py:189).
//
// Shape of the implementation: decoder activations live on a TIME-major tape (row = t * N + n), so
//   * the teacher-forced pass runs every kernel once over all L * N rows,
//   * the sampling pass runs the same kernels on the N rows of one position per step (the tape's earlier K/V rows are its cache),
//   * the backward pass is always batched over the L * N rows: every contraction is a tcgen05 kind::tf32 GEMM (gemm_tf32.cu).
// Dropout masks are functions of (seed, site, position, element), so both forward forms draw the same masks.  Sites: 1 att_embed;
// 2 target embedding + positional encoding; encoder layer l: 10+l attention probabilities, 20+l / 40+l the two SublayerConnections,
// 30+l the feed-forward hidden layer; decoder layer l: 50+l self-attention probabilities, 60+l / 80+l / 100+l the three
// SublayerConnections, 70+l source-attention probabilities, 90+l the feed-forward hidden layer.
// One deviation, documented in DESIGN.md: _forward repeats the image features per caption BEFORE the encoder (:329-333), so the
// reference runs the encoder seq_per_img times per image with independent dropout masks; here the encoder runs once per image and its
// output is shared by the image's captions (identical when dropout is off; with dropout on, one encoder mask per image instead of five).
// =====================================================================================================================
namespace {

constexpr int TML = CAPB200_TFM_MAX_LAYERS;

struct TTape {
    // encoder, rows b * R + r
    float *X[TML + 1], *eln0[TML], *eqkv[TML], *eatt[TML], *xm[TML], *eln1[TML], *ehd[TML], *mem, *skv[TML];
    // decoder, rows t * N + n
    float *Y[TML + 1], *dln0[TML], *dqkv[TML], *datt[TML], *ym1[TML], *dln1[TML], *dqs[TML], *probs[TML], *dcatt[TML], *ym2[TML], *dln2[TML], *dhd[TML];
    float *yln_tm, *yln_nm;
    int* tok;
    float* key_mask;
    // gradients / scratch
    float *DL, *d_yln_nm, *dY, *d_tmp, *d_h, *d_ln, *d_att, *d_qs, *d_qkv, *d_skv[TML], *d_mem, *dX, *tmp, *stats, *mask_sum, *item_loss, *glp, *skinny;
    size_t skinny_floats;
    double* scores;
    int *s_tokens, *s_unfinished, *s_forced;
    float *row_loss, *row_msum, *row_coef;
};

void layout_ttape(TTape& tp, Arena& a, int B, int R, int N, int L, int T, int D, int Dff, int heads, int V1, int NE, int ND, bool scst) {
    const long BR = (long)B * R, LN = (long)L * N;
    const long big = BR > LN ? BR : LN;
    for (int l = 0; l <= NE; ++l) tp.X[l] = a.take<float>(BR * D);
    for (int l = 0; l < NE; ++l) {
        tp.eln0[l] = a.take<float>(BR * D); tp.eqkv[l] = a.take<float>(BR * 3 * D); tp.eatt[l] = a.take<float>(BR * D); tp.xm[l] = a.take<float>(BR * D);
        tp.eln1[l] = a.take<float>(BR * D); tp.ehd[l] = a.take<float>(BR * Dff);
    }
    tp.mem = a.take<float>(BR * D);
    for (int l = 0; l < ND; ++l) { tp.skv[l] = a.take<float>(BR * 2 * D); tp.d_skv[l] = a.take<float>(BR * 2 * D); }
    for (int l = 0; l <= ND; ++l) tp.Y[l] = a.take<float>(LN * D);
    for (int l = 0; l < ND; ++l) {
        tp.dln0[l] = a.take<float>(LN * D); tp.dqkv[l] = a.take<float>(LN * 3 * D); tp.datt[l] = a.take<float>(LN * D); tp.ym1[l] = a.take<float>(LN * D);
        tp.dln1[l] = a.take<float>(LN * D); tp.dqs[l] = a.take<float>(LN * D); tp.probs[l] = a.take<float>(LN * heads * R); tp.dcatt[l] = a.take<float>(LN * D);
        tp.ym2[l] = a.take<float>(LN * D); tp.dln2[l] = a.take<float>(LN * D); tp.dhd[l] = a.take<float>(LN * Dff);
    }
    tp.yln_tm = a.take<float>(LN * D); tp.yln_nm = a.take<float>(LN * D);
    tp.tok = a.take<int>(LN);
    tp.key_mask = a.take<float>(LN);
    tp.DL = a.take<float>(LN * V1); tp.d_yln_nm = a.take<float>(LN * D); tp.dY = a.take<float>(LN * D);
    tp.d_tmp = a.take<float>(big * D); tp.d_h = a.take<float>(big * Dff); tp.d_ln = a.take<float>(big * D); tp.d_att = a.take<float>(big * D);
    tp.d_qs = a.take<float>(LN * D); tp.d_qkv = a.take<float>(big * 3 * D); tp.d_mem = a.take<float>(BR * D); tp.dX = a.take<float>(BR * D);
    tp.tmp = a.take<float>(big * D);
    tp.stats = a.take<float>(2 * big); tp.mask_sum = a.take<float>(8); tp.item_loss = a.take<float>(LN);
    tp.glp = a.take<float>(scst ? (long)B * T * V1 : 1);
    tp.skinny_floats = (size_t)4 << 20;
    tp.skinny = a.take<float>((long)tp.skinny_floats);
    tp.scores = a.take<double>((long)N + B);
    tp.s_tokens = a.take<int>(N); tp.s_unfinished = a.take<int>(N); tp.s_forced = a.take<int>(N);
    tp.row_loss = a.take<float>(N); tp.row_msum = a.take<float>(N); tp.row_coef = a.take<float>(N);
}

struct TfmTrainArgs {
    bool xe = false;
    int n = 1, L = 0;                      // rows per image; positions evaluated (XE: label_cols - 1, SCST: seq_length)
    float p_lm = 0.f, p = 0.f, temperature = 1.f, upstream = 1.f, smoothing = 0.f;
    unsigned long long seed = 0;
    bool greedy_baseline = true;
    const capb200_cider_table* table = nullptr;
    const int* refs = nullptr; const int* ref_offsets = nullptr; int Lref = 0;
    long long* sample_seq = nullptr; long long* greedy_seq = nullptr; float* reward = nullptr;
    const long long* forced = nullptr;
    const float* mask = nullptr;
    int keep = 0;
    float* row_loss = nullptr;
    const long long* labels = nullptr; long ld_labels = 0; const float* masks = nullptr; long ld_masks = 0;
    float* logprobs = nullptr; float* loss = nullptr;
};

int tfm_train_step(capb200_tfm_engine* e, const float* att, int B, int R, const TfmTrainArgs& ta, const capb200_tfm_grads* grads, cudaStream_t st) {
    const int n = ta.n, N = B * n, L = ta.L, D = e->D, Dff = e->Dff, V1 = e->V1, F = e->F, heads = e->H, dk = e->dk, NE = e->NE, ND = e->ND, T = e->T;
    const int BR = B * R, LNr = L * N;
    const int idxL = T + 2;                            // fixed pitch of the self-attention dropout index (positions never reach it)
    const float p = ta.p, p_lm = ta.p_lm;
    const unsigned long long seed = ta.seed;
    const capb200_tfm_weights& w = e->w;
    const capb200_tfm_grads& G = *grads;
    const float emb_scale = sqrtf((float)D);
    CAPB_REQUIRE(L >= 1 && L <= T + 1 && L < 32, "positions out of range");
    {
        Arena dry; TTape t0; layout_ttape(t0, dry, B, R, N, L, T, D, Dff, heads, V1, NE, ND, !ta.xe);
        if (dry.off + 256 > e->tape_bytes) {
            CAPB_CHECK_CUDA(cudaStreamSynchronize(st));
            if (e->tape) CAPB_CHECK_CUDA(cudaFree(e->tape));
            e->tape = nullptr;
            CAPB_CHECK_CUDA(cudaMalloc(&e->tape, dry.off + 256));
            e->tape_bytes = dry.off + 256;
        }
    }
    Arena ar; ar.base = e->tape;
    TTape tp; layout_ttape(tp, ar, B, R, N, L, T, D, Dff, heads, V1, NE, ND, !ta.xe);

    // ---- greedy baseline (eval mode): the regular K/V-cached decode on a side stream, joined before the reward
    const bool greedy_baseline = !ta.xe && ta.greedy_baseline;
    bool greedy_on_side = false;
    cudaStream_t gs_enqueue = st;
    capb200_sample_opts so;
    if (greedy_baseline) {
        if (ensure_workspace(e, B, B, R, 1, st)) return 1;
        memset(&so, 0, sizeof(so)); so.edits.unk_col = -1; so.sample_n = 1; so.method = CAPB200_SAMPLE_GREEDY; so.temperature = 1.f; so.steps = T;
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
    auto act = [](float* ptr, long ld) { ActView v; v.f = ptr; v.hi = nullptr; v.lo = nullptr; v.ld = ld; return v; };
    long& nl = e->launches;
    int rc = 0;

    // ---- encoder forward on the tape (train mode)
    rc |= sk.lin(att, F, w.att_embed_w, F, w.att_embed_b, tp.X[0], D, BR, D, F, 0);
    rc |= relu_dropout_rows_launch(BR, BR, D, 0, tp.X[0], D, seed, 1, p_lm, st);
    if (ta.mask != nullptr) { rc |= mask_rows_launch(act(tp.X[0], D), B, R, D, ta.mask, R, st); nl++; }
    nl += 2;
    for (int l = 0; l < NE && !rc; ++l) {
        const capb200_tfm_enc_layer& Lw = w.enc[l];
        rc |= layer_norm_launch(BR, D, tp.X[l], D, Lw.ln0_a, Lw.ln0_b, 1e-6f, act(tp.eln0[l], D), st);
        rc |= sk.lin(tp.eln0[l], D, e->enc_qkv_w[l], D, e->enc_qkv_b[l], tp.eqkv[l], 3 * D, BR, 3 * D, D, 0);
        rc |= seq_attn_train_launch(B, R, 0, R, heads, dk, 0, R, R, 1, tp.eqkv[l], tp.eqkv[l] + D, tp.eqkv[l] + 2 * D, 3 * D, seed, 10 + l, p, tp.eatt[l], D, ta.mask, R, st);
        rc |= sk.lin(tp.eatt[l], D, Lw.self_attn.o_w, D, Lw.self_attn.o_b, tp.tmp, D, BR, D, D, 0);
        rc |= add_dropout_rows_launch(BR, BR, D, 0, tp.X[l], D, tp.tmp, D, tp.xm[l], D, seed, 20 + l, p, st);
        rc |= layer_norm_launch(BR, D, tp.xm[l], D, Lw.ln1_a, Lw.ln1_b, 1e-6f, act(tp.eln1[l], D), st);
        rc |= sk.lin(tp.eln1[l], D, Lw.w1_w, D, Lw.w1_b, tp.ehd[l], Dff, BR, Dff, D, 0);
        rc |= relu_dropout_rows_launch(BR, BR, Dff, 0, tp.ehd[l], Dff, seed, 30 + l, p, st);
        rc |= sk.lin(tp.ehd[l], Dff, Lw.w2_w, Dff, Lw.w2_b, tp.tmp, D, BR, D, Dff, 0);
        rc |= add_dropout_rows_launch(BR, BR, D, 0, tp.xm[l], D, tp.tmp, D, tp.X[l + 1], D, seed, 40 + l, p, st);
        nl += 10;
    }
    rc |= layer_norm_launch(BR, D, tp.X[NE], D, w.enc_norm_a, w.enc_norm_b, 1e-6f, act(tp.mem, D), st);
    for (int l = 0; l < ND; ++l) rc |= sk.lin(tp.mem, D, e->dec_skv_w[l], D, e->dec_skv_b[l], tp.skv[l], 2 * D, BR, 2 * D, D, 0);
    nl += 1 + ND;
    if (rc) return 1;

    // ---- the greedy baseline's launches are enqueued only now: its stream forked at the top of the step, and while the host enqueues
    // them the main stream is busy with the encoder instead of idle
    if (greedy_baseline) {
        CAPB_CHECK_CUDA(cudaMemsetAsync(tp.glp, 0, sizeof(float) * (size_t)B * T * V1, gs_enqueue));
        CAPB_CHECK_CUDA(cudaMemsetAsync(ta.greedy_seq, 0, sizeof(long long) * (size_t)B * T, gs_enqueue));
        if (capb200_tfm_decode_sample(e, att, ta.mask, B, R, &so, nullptr, 0, ta.greedy_seq, tp.glp, nullptr, static_cast<void*>(gs_enqueue))) return 1;
        if (greedy_on_side) CAPB_CHECK_CUDA(cudaEventRecord(e->ev_join, e->side));
    }

    // ---- decoder forward over positions [t0, t1)
    const float* key_mask = ta.xe ? tp.key_mask : nullptr;
    auto dec_forward = [&](int t0, int t1) -> int {
        const long r0 = (long)t0 * N;
        const int nr = (t1 - t0) * N;
        int r = 0;
        r |= embed_pe_dropout_launch(nr, N, D, tp.tok + r0, w.lut, w.pe, emb_scale, t0, seed, 2, p, tp.Y[0] + r0 * D, D, st);
        for (int l = 0; l < ND && !r; ++l) {
            const capb200_tfm_dec_layer& Lw = w.dec[l];
            r |= layer_norm_launch(nr, D, tp.Y[l] + r0 * D, D, Lw.ln0_a, Lw.ln0_b, 1e-6f, act(tp.dln0[l] + r0 * D, D), st);
            r |= sk.lin(tp.dln0[l] + r0 * D, D, e->dec_qkv_w[l], D, e->dec_qkv_b[l], tp.dqkv[l] + r0 * 3 * D, 3 * D, nr, 3 * D, D, 0);
            r |= seq_attn_train_launch(N, t1, t0, t1, heads, dk, 1, idxL, 1, N, tp.dqkv[l], tp.dqkv[l] + D, tp.dqkv[l] + 2 * D, 3 * D, seed, 50 + l, p, tp.datt[l], D,
                                       key_mask, L, st);
            r |= sk.lin(tp.datt[l] + r0 * D, D, Lw.self_attn.o_w, D, Lw.self_attn.o_b, tp.tmp, D, nr, D, D, 0);
            r |= add_dropout_rows_launch(nr, N, D, t0, tp.Y[l] + r0 * D, D, tp.tmp, D, tp.ym1[l] + r0 * D, D, seed, 60 + l, p, st);
            r |= layer_norm_launch(nr, D, tp.ym1[l] + r0 * D, D, Lw.ln1_a, Lw.ln1_b, 1e-6f, act(tp.dln1[l] + r0 * D, D), st);
            r |= sk.lin(tp.dln1[l] + r0 * D, D, Lw.src_attn.q_w, D, Lw.src_attn.q_b, tp.dqs[l] + r0 * D, D, nr, D, D, 0);
            r |= cross_attn_train_launch(nr, n, heads, dk, R, tp.dqs[l] + r0 * D, D, tp.skv[l], tp.skv[l] + D, 2 * D, seed, 70 + l, t0, p, tp.dcatt[l] + r0 * D, D,
                                         tp.probs[l] + r0 * heads * R, st, ta.mask, R, N);
            r |= sk.lin(tp.dcatt[l] + r0 * D, D, Lw.src_attn.o_w, D, Lw.src_attn.o_b, tp.tmp, D, nr, D, D, 0);
            r |= add_dropout_rows_launch(nr, N, D, t0, tp.ym1[l] + r0 * D, D, tp.tmp, D, tp.ym2[l] + r0 * D, D, seed, 80 + l, p, st);
            r |= layer_norm_launch(nr, D, tp.ym2[l] + r0 * D, D, Lw.ln2_a, Lw.ln2_b, 1e-6f, act(tp.dln2[l] + r0 * D, D), st);
            r |= sk.lin(tp.dln2[l] + r0 * D, D, Lw.w1_w, D, Lw.w1_b, tp.dhd[l] + r0 * Dff, Dff, nr, Dff, D, 0);
            r |= relu_dropout_rows_launch(nr, N, Dff, t0, tp.dhd[l] + r0 * Dff, Dff, seed, 90 + l, p, st);
            r |= sk.lin(tp.dhd[l] + r0 * Dff, Dff, Lw.w2_w, Dff, Lw.w2_b, tp.tmp, D, nr, D, Dff, 0);
            r |= add_dropout_rows_launch(nr, N, D, t0, tp.ym2[l] + r0 * D, D, tp.tmp, D, tp.Y[l + 1] + r0 * D, D, seed, 100 + l, p, st);
            nl += 16;
        }
        r |= layer_norm_launch(nr, D, tp.Y[ND] + r0 * D, D, w.dec_norm_a, w.dec_norm_b, 1e-6f, act(tp.yln_tm + r0 * D, D), st);
        nl += 2;
        return r;
    };

    const long ld_lp = (long)L * V1;                   // log-prob row pitch of one sequence: [N, L, V1]
    if (ta.xe) {
        if (load_tokens_tm_launch(ta.labels, ta.ld_labels, N, L, tp.tok, tp.key_mask, L, st)) return 1;
        if (dec_forward(0, L)) return 1;
        rc |= permute_rows_launch(L, N, D, tp.yln_tm, D, tp.yln_nm, D, 1, st);
        rc |= sk.lin(tp.yln_nm, D, w.gen_w, D, w.gen_b, ta.logprobs, V1, LNr, V1, D, 0);
        VocabStepArgs va; va.rows = LNr; va.V1 = V1; va.logits = ta.logprobs; va.ld = V1;
        rc |= vocab_step_launch(va, st);
        nl += 4;
        if (rc) return 1;
        if (xe_loss_backward_launch(ta.logprobs, ld_lp, ta.labels, ta.ld_labels, ta.masks, ta.ld_masks, N, L, L, V1, ta.smoothing, ta.upstream, tp.mask_sum,
                                    tp.item_loss, tp.DL, ta.loss, st, ta.keep, ta.row_loss ? ta.row_loss : tp.row_loss, tp.row_msum, tp.row_coef)) return 1;
    } else {
        CAPB_CHECK_CUDA(cudaMemsetAsync(tp.s_tokens, 0, sizeof(int) * N, st));
        for (int t = 0; t < L; ++t) {
            CAPB_CHECK_CUDA(cudaMemcpyAsync(tp.tok + (long)t * N, tp.s_tokens, sizeof(int) * N, cudaMemcpyDeviceToDevice, st));
            if (dec_forward(t, t + 1)) return 1;
            float* logits = ta.logprobs + (long)t * V1;
            if (sk.lin(tp.yln_tm + (long)t * N * D, D, w.gen_w, D, w.gen_b, logits, ld_lp, N, V1, D, 0)) return 1;
            VocabStepArgs va;
            va.rows = N; va.V1 = V1; va.logits = logits; va.ld = ld_lp;
            va.select = 2; va.temperature = ta.temperature; va.seed = seed; va.step = (unsigned long long)t;
            va.unfinished = tp.s_unfinished; va.first_step = (t == 0); va.tokens_out = tp.s_tokens;
            va.seq_out = ta.sample_seq; va.ld_seq = L; va.t = t;
            if (ta.forced != nullptr) {
                if (load_token_column_launch(ta.forced, L, t, N, tp.s_forced, st)) return 1;
                va.select = 3; va.forced = tp.s_forced;
            }
            if (vocab_step_launch(va, st)) return 1;
            nl += 3;
        }
        if (permute_rows_launch(L, N, D, tp.yln_tm, D, tp.yln_nm, D, 1, st)) return 1;
        if (greedy_on_side) CAPB_CHECK_CUDA(cudaStreamWaitEvent(st, e->ev_join, 0));
        if (cider_reward_launch(ta.table->t, ta.sample_seq, N, greedy_baseline ? ta.greedy_seq : nullptr, B, L, ta.refs, ta.ref_offsets, ta.Lref, tp.scores, ta.reward,
                                L, L, st)) return 1;
        float* rl = ta.keep > 0 ? (ta.row_loss ? ta.row_loss : tp.row_loss) : nullptr;
        if (reward_criterion_fwd_launch(ta.logprobs, ld_lp, V1, ta.sample_seq, ta.reward, N, L, ta.loss, rl, tp.mask_sum, st)) return 1;
        if (ta.keep > 0 && scst_drop_worst_launch(ta.sample_seq, rl, N, L, ta.keep, ta.upstream, tp.row_msum, tp.row_coef, ta.loss, st)) return 1;
        if (scst_dlogits_launch(ta.logprobs, ld_lp, ta.sample_seq, ta.reward, tp.mask_sum, ta.upstream, N, L, V1, tp.DL, st, ta.keep > 0 ? tp.row_coef : nullptr)) return 1;
        nl += 5;
    }

    // ---- backward: generator and the final LayerNorm
    auto colsum = [&](int rows, int cols, const float* x, long ld, float* out) { nl++; return colsum_launch(rows, cols, x, ld, out, 0, st); };
    // q | k | v gradients of a self-attention block: one GEMM / one column reduction when the caller laid the three tensors out back to back
    // (the Python mirror's flat gradient buffer does), else three
    auto qkv_grads = [&](int rows, const float* x, const capb200_mha_grads& ag) -> int {
        int r = 0;
        if (ag.k_w == ag.q_w + (long)D * D && ag.v_w == ag.k_w + (long)D * D) r |= sk.wgrad(3 * D, D, rows, tp.d_qkv, 3 * D, x, D, ag.q_w, D, 0);
        else {
            r |= sk.wgrad(D, D, rows, tp.d_qkv, 3 * D, x, D, ag.q_w, D, 0);
            r |= sk.wgrad(D, D, rows, tp.d_qkv + D, 3 * D, x, D, ag.k_w, D, 0);
            r |= sk.wgrad(D, D, rows, tp.d_qkv + 2 * D, 3 * D, x, D, ag.v_w, D, 0);
        }
        if (ag.k_b == ag.q_b + D && ag.v_b == ag.k_b + D) r |= colsum(rows, 3 * D, tp.d_qkv, 3 * D, ag.q_b);
        else {
            r |= colsum(rows, D, tp.d_qkv, 3 * D, ag.q_b);
            r |= colsum(rows, D, tp.d_qkv + D, 3 * D, ag.k_b);
            r |= colsum(rows, D, tp.d_qkv + 2 * D, 3 * D, ag.v_b);
        }
        return r;
    };
    rc |= sk.dgrad(LNr, D, V1, tp.DL, V1, w.gen_w, D, tp.d_yln_nm, D, 0);
    rc |= sk.wgrad(V1, D, LNr, tp.DL, V1, tp.yln_nm, D, G.gen_w, D, 0);
    rc |= colsum(LNr, V1, tp.DL, V1, G.gen_b);
    rc |= permute_rows_launch(L, N, D, tp.d_yln_nm, D, tp.d_tmp, D, 0, st);
    rc |= ln_backward_launch(LNr, D, tp.Y[ND], D, w.dec_norm_a, tp.d_tmp, D, 1e-6f, tp.dY, D, 0, tp.stats, G.dec_norm_a, G.dec_norm_b, 0, st);
    nl += 3;
    for (int l = 0; l < ND; ++l) CAPB_CHECK_CUDA(cudaMemsetAsync(tp.d_skv[l], 0, sizeof(float) * (size_t)BR * 2 * D, st));
    if (rc) return 1;
    for (int l = ND - 1; l >= 0 && !rc; --l) {
        const capb200_tfm_dec_layer& Lw = w.dec[l];
        const capb200_tfm_dec_layer_grads& Lg = G.dec[l];
        // feed-forward sublayer: Y[l+1] = ym2 + dropout(w2(dropout(relu(w1(ln2(ym2))))))
        rc |= dropout_rows_copy_launch(LNr, N, D, 0, tp.dY, D, tp.d_tmp, D, seed, 100 + l, p, nullptr, 0, st);
        rc |= sk.wgrad(D, Dff, LNr, tp.d_tmp, D, tp.dhd[l], Dff, Lg.w2_w, Dff, 0);
        rc |= colsum(LNr, D, tp.d_tmp, D, Lg.w2_b);
        rc |= sk.dgrad(LNr, Dff, D, tp.d_tmp, D, Lw.w2_w, Dff, tp.d_h, Dff, 0);
        rc |= dropout_rows_copy_launch(LNr, N, Dff, 0, tp.d_h, Dff, tp.d_h, Dff, seed, 90 + l, p, tp.dhd[l], Dff, st);
        rc |= sk.wgrad(Dff, D, LNr, tp.d_h, Dff, tp.dln2[l], D, Lg.w1_w, D, 0);
        rc |= colsum(LNr, Dff, tp.d_h, Dff, Lg.w1_b);
        rc |= sk.dgrad(LNr, D, Dff, tp.d_h, Dff, Lw.w1_w, D, tp.d_ln, D, 0);
        rc |= ln_backward_launch(LNr, D, tp.ym2[l], D, Lw.ln2_a, tp.d_ln, D, 1e-6f, tp.dY, D, 1, tp.stats, Lg.ln2_a, Lg.ln2_b, 0, st);
        // source attention sublayer: ym2 = ym1 + dropout(o(attention(q(ln1(ym1)), memory)))
        rc |= dropout_rows_copy_launch(LNr, N, D, 0, tp.dY, D, tp.d_tmp, D, seed, 80 + l, p, nullptr, 0, st);
        rc |= sk.wgrad(D, D, LNr, tp.d_tmp, D, tp.dcatt[l], D, Lg.src_attn.o_w, D, 0);
        rc |= colsum(LNr, D, tp.d_tmp, D, Lg.src_attn.o_b);
        rc |= sk.dgrad(LNr, D, D, tp.d_tmp, D, Lw.src_attn.o_w, D, tp.d_att, D, 0);
        rc |= cross_attn_backward_launch(B, n, heads, dk, R, tp.dqs[l], D, tp.skv[l], tp.skv[l] + D, 2 * D, seed, 70 + l, 0, p, tp.probs[l], tp.d_att, D, tp.d_qs, D,
                                         tp.d_skv[l], tp.d_skv[l] + D, 2 * D, st, L, N);
        rc |= sk.wgrad(D, D, LNr, tp.d_qs, D, tp.dln1[l], D, Lg.src_attn.q_w, D, 0);
        rc |= colsum(LNr, D, tp.d_qs, D, Lg.src_attn.q_b);
        rc |= sk.dgrad(LNr, D, D, tp.d_qs, D, Lw.src_attn.q_w, D, tp.d_ln, D, 0);
        rc |= ln_backward_launch(LNr, D, tp.ym1[l], D, Lw.ln1_a, tp.d_ln, D, 1e-6f, tp.dY, D, 1, tp.stats, Lg.ln1_a, Lg.ln1_b, 0, st);
        // self-attention sublayer: ym1 = Y[l] + dropout(o(causal attention(q|k|v(ln0(Y[l])))))
        rc |= dropout_rows_copy_launch(LNr, N, D, 0, tp.dY, D, tp.d_tmp, D, seed, 60 + l, p, nullptr, 0, st);
        rc |= sk.wgrad(D, D, LNr, tp.d_tmp, D, tp.datt[l], D, Lg.self_attn.o_w, D, 0);
        rc |= colsum(LNr, D, tp.d_tmp, D, Lg.self_attn.o_b);
        rc |= sk.dgrad(LNr, D, D, tp.d_tmp, D, Lw.self_attn.o_w, D, tp.d_att, D, 0);
        rc |= seq_attn_backward_launch(N, L, heads, dk, 1, idxL, 1, N, tp.dqkv[l], tp.dqkv[l] + D, tp.dqkv[l] + 2 * D, 3 * D, seed, 50 + l, p, tp.d_att, D, tp.d_qkv,
                                       tp.d_qkv + D, tp.d_qkv + 2 * D, 3 * D, key_mask, L, st);
        rc |= qkv_grads(LNr, tp.dln0[l], Lg.self_attn);
        rc |= sk.dgrad(LNr, D, 3 * D, tp.d_qkv, 3 * D, e->dec_qkv_w[l], D, tp.d_ln, D, 0);
        rc |= ln_backward_launch(LNr, D, tp.Y[l], D, Lw.ln0_a, tp.d_ln, D, 1e-6f, tp.dY, D, 1, tp.stats, Lg.ln0_a, Lg.ln0_b, 0, st);
        nl += 14;
    }
    if (rc) return 1;
    CAPB_CHECK_CUDA(cudaMemsetAsync(G.lut, 0, sizeof(float) * (size_t)V1 * D, st));
    rc |= embed_pe_backward_launch(LNr, N, D, tp.tok, emb_scale, 0, seed, 2, p, tp.dY, D, G.lut, st);
    // memory: K | V projections of every decoder layer
    for (int l = 0; l < ND; ++l) {
        const capb200_tfm_dec_layer_grads& Lg = G.dec[l];
        rc |= sk.dgrad(BR, D, 2 * D, tp.d_skv[l], 2 * D, e->dec_skv_w[l], D, tp.d_mem, D, l > 0 ? 1 : 0);
        if (Lg.src_attn.v_w == Lg.src_attn.k_w + (long)D * D) rc |= sk.wgrad(2 * D, D, BR, tp.d_skv[l], 2 * D, tp.mem, D, Lg.src_attn.k_w, D, 0);
        else {
            rc |= sk.wgrad(D, D, BR, tp.d_skv[l], 2 * D, tp.mem, D, Lg.src_attn.k_w, D, 0);
            rc |= sk.wgrad(D, D, BR, tp.d_skv[l] + D, 2 * D, tp.mem, D, Lg.src_attn.v_w, D, 0);
        }
        if (Lg.src_attn.v_b == Lg.src_attn.k_b + D) rc |= colsum(BR, 2 * D, tp.d_skv[l], 2 * D, Lg.src_attn.k_b);
        else {
            rc |= colsum(BR, D, tp.d_skv[l], 2 * D, Lg.src_attn.k_b);
            rc |= colsum(BR, D, tp.d_skv[l] + D, 2 * D, Lg.src_attn.v_b);
        }
    }
    if (rc) return 1;
    if (record_group_event(e->grad_events[0], st)) return 1;                                    // generator + decoder + target embedding
    rc |= ln_backward_launch(BR, D, tp.X[NE], D, w.enc_norm_a, tp.d_mem, D, 1e-6f, tp.dX, D, 0, tp.stats, G.enc_norm_a, G.enc_norm_b, 0, st);
    for (int l = NE - 1; l >= 0 && !rc; --l) {
        const capb200_tfm_enc_layer& Lw = w.enc[l];
        const capb200_tfm_enc_layer_grads& Lg = G.enc[l];
        rc |= dropout_rows_copy_launch(BR, BR, D, 0, tp.dX, D, tp.d_tmp, D, seed, 40 + l, p, nullptr, 0, st);
        rc |= sk.wgrad(D, Dff, BR, tp.d_tmp, D, tp.ehd[l], Dff, Lg.w2_w, Dff, 0);
        rc |= colsum(BR, D, tp.d_tmp, D, Lg.w2_b);
        rc |= sk.dgrad(BR, Dff, D, tp.d_tmp, D, Lw.w2_w, Dff, tp.d_h, Dff, 0);
        rc |= dropout_rows_copy_launch(BR, BR, Dff, 0, tp.d_h, Dff, tp.d_h, Dff, seed, 30 + l, p, tp.ehd[l], Dff, st);
        rc |= sk.wgrad(Dff, D, BR, tp.d_h, Dff, tp.eln1[l], D, Lg.w1_w, D, 0);
        rc |= colsum(BR, Dff, tp.d_h, Dff, Lg.w1_b);
        rc |= sk.dgrad(BR, D, Dff, tp.d_h, Dff, Lw.w1_w, D, tp.d_ln, D, 0);
        rc |= ln_backward_launch(BR, D, tp.xm[l], D, Lw.ln1_a, tp.d_ln, D, 1e-6f, tp.dX, D, 1, tp.stats, Lg.ln1_a, Lg.ln1_b, 0, st);
        rc |= dropout_rows_copy_launch(BR, BR, D, 0, tp.dX, D, tp.d_tmp, D, seed, 20 + l, p, nullptr, 0, st);
        rc |= sk.wgrad(D, D, BR, tp.d_tmp, D, tp.eatt[l], D, Lg.self_attn.o_w, D, 0);
        rc |= colsum(BR, D, tp.d_tmp, D, Lg.self_attn.o_b);
        rc |= sk.dgrad(BR, D, D, tp.d_tmp, D, Lw.self_attn.o_w, D, tp.d_att, D, 0);
        rc |= seq_attn_backward_launch(B, R, heads, dk, 0, R, R, 1, tp.eqkv[l], tp.eqkv[l] + D, tp.eqkv[l] + 2 * D, 3 * D, seed, 10 + l, p, tp.d_att, D, tp.d_qkv,
                                       tp.d_qkv + D, tp.d_qkv + 2 * D, 3 * D, ta.mask, R, st);
        rc |= qkv_grads(BR, tp.eln0[l], Lg.self_attn);
        rc |= sk.dgrad(BR, D, 3 * D, tp.d_qkv, 3 * D, e->enc_qkv_w[l], D, tp.d_ln, D, 0);
        rc |= ln_backward_launch(BR, D, tp.X[l], D, Lw.ln0_a, tp.d_ln, D, 1e-6f, tp.dX, D, 1, tp.stats, Lg.ln0_a, Lg.ln0_b, 0, st);
        nl += 8;
    }
    rc |= dropout_rows_copy_launch(BR, BR, D, 0, tp.dX, D, tp.d_tmp, D, seed, 1, p_lm, tp.X[0], D, st);
    rc |= sk.wgrad(D, F, BR, tp.d_tmp, D, att, F, G.att_embed_w, F, 0);
    rc |= colsum(BR, D, tp.d_tmp, D, G.att_embed_b);
    nl += 2 + tf32_context_launches(e->tf32) - tf32_l0;
    if (!rc && record_group_event(e->grad_events[1], st)) return 1;                             // encoder + att_embed
    return rc;
}

}  // namespace

extern "C" int capb200_tfm_set_grad_events(capb200_tfm_engine* e, void* const* events, int n) {
    CAPB_REQUIRE(e != nullptr && n >= 0 && n <= 2, "the transformer has 2 gradient groups");
    for (int i = 0; i < 2; ++i) e->grad_events[i] = (events != nullptr && i < n) ? static_cast<cudaEvent_t>(events[i]) : nullptr;
    return 0;
}

extern "C" int capb200_tfm_xe_step(capb200_tfm_engine* e, const float* att, int B, int R, const capb200_tfm_xe_opts* opts, const long long* labels,
                                   const float* masks, int label_cols, const capb200_tfm_grads* grads, float* logprobs, float* loss, void* stream) {
    if (check_ready(e)) return 1;
    CAPB_REQUIRE(opts && att && labels && masks && grads && logprobs && loss, "null argument");
    CAPB_REQUIRE(opts->seq_per_img >= 1 && opts->seq_per_img <= 16 && B >= 1 && R >= 1, "seq_per_img must be in 1..16");
    CAPB_REQUIRE(opts->drop_prob_lm >= 0.f && opts->drop_prob_lm < 1.f && opts->dropout >= 0.f && opts->dropout < 1.f, "dropout rates must be in [0, 1)");
    CAPB_REQUIRE(opts->label_smoothing >= 0.f && opts->label_smoothing < 1.f, "label_smoothing must be in [0, 1)");
    CAPB_REQUIRE(label_cols >= 2 && label_cols <= e->T + 2, "labels are [N, seq_length + 2] (BOS, words, EOS padding)");
    TfmTrainArgs ta;
    ta.xe = true; ta.n = opts->seq_per_img; ta.L = label_cols - 1; ta.p_lm = opts->drop_prob_lm; ta.p = opts->dropout; ta.upstream = opts->upstream;
    ta.seed = opts->seed; ta.smoothing = opts->label_smoothing; ta.labels = labels; ta.ld_labels = label_cols; ta.masks = masks; ta.ld_masks = label_cols;
    ta.logprobs = logprobs; ta.loss = loss; ta.mask = opts->att_masks; ta.keep = opts->keep_rows; ta.row_loss = opts->row_loss;
    CAPB_REQUIRE(ta.keep >= 0 && ta.keep <= B * ta.n, "keep_rows must be in 0..rows");
    if (dropout_salt_set_all(0ull, static_cast<cudaStream_t>(stream))) return 1;      // eager step: the seed arguments are the effective seeds
    return tfm_train_step(e, att, B, R, ta, grads, static_cast<cudaStream_t>(stream));
}

extern "C" int capb200_tfm_scst_step(capb200_tfm_engine* e, const float* att, int B, int R, const capb200_tfm_scst_opts* opts, const capb200_cider_table* table,
                                     const int* refs, const int* ref_offsets, int L, const capb200_tfm_grads* grads, long long* sample_seq, long long* greedy_seq,
                                     float* sample_logprobs, float* reward, float* loss, void* stream) {
    if (check_ready(e)) return 1;
    CAPB_REQUIRE(opts && att && table && refs && ref_offsets && grads && sample_seq && sample_logprobs && reward && loss, "null argument");
    const bool greedy_baseline = opts->baseline == CAPB200_BASELINE_GREEDY;
    CAPB_REQUIRE(greedy_baseline || opts->baseline == CAPB200_BASELINE_LEAVE_ONE_OUT, "unknown baseline");
    CAPB_REQUIRE(!greedy_baseline || greedy_seq != nullptr, "the greedy baseline needs greedy_seq");
    const int n = opts->sample_n;
    CAPB_REQUIRE(n >= 1 && n <= 16 && (greedy_baseline || n >= 2) && B >= 1 && R >= 1, "sample_n must be in 1..16 (>= 2 for the leave-one-out baseline)");
    CAPB_REQUIRE(opts->drop_prob_lm >= 0.f && opts->drop_prob_lm < 1.f && opts->dropout >= 0.f && opts->dropout < 1.f, "dropout rates must be in [0, 1)");
    CAPB_REQUIRE(opts->temperature > 0.f, "temperature must be positive");
    TfmTrainArgs ta;
    ta.n = n; ta.L = e->T; ta.p_lm = opts->drop_prob_lm; ta.p = opts->dropout; ta.temperature = opts->temperature; ta.upstream = opts->upstream; ta.seed = opts->seed;
    ta.greedy_baseline = greedy_baseline; ta.table = table; ta.refs = refs; ta.ref_offsets = ref_offsets; ta.Lref = L; ta.sample_seq = sample_seq;
    ta.greedy_seq = greedy_seq; ta.reward = reward; ta.logprobs = sample_logprobs; ta.loss = loss; ta.forced = opts->forced_tokens; ta.mask = opts->att_masks;
    ta.keep = opts->keep_rows; ta.row_loss = opts->row_loss;
    CAPB_REQUIRE(ta.keep >= 0 && ta.keep <= B * n, "keep_rows must be in 0..rows");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the whole step (~4900 launches at 6 + 6 layers) as one CUDA graph: see capb200_aoa_scst_step and engine_common.cuh (StepGraph)
    if (!StepGraph::enabled() || !e->tc || ta.forced != nullptr || e->sg.broken) {
        if (dropout_salt_set_all(0ull, st)) return 1;      // eager step: the seed arguments are the effective seeds
        return tfm_train_step(e, att, B, R, ta, grads, st);
    }
    cudaStream_t gst = e->sg.enter(st);             // a capturable engine-owned stream, ordered after the caller's stream
    const void* srcs[2] = {att, ta.mask};
    const size_t bytes[2] = {sizeof(float) * (size_t)B * R * e->F, ta.mask ? sizeof(float) * (size_t)B * R : 0};
    size_t off[2];
    if (e->sg.stage_inputs(2, srcs, bytes, off, gst)) return 1;
    const float* att_s = reinterpret_cast<const float*>(e->sg.stage + off[0]);
    if (ta.mask) ta.mask = reinterpret_cast<const float*>(e->sg.stage + off[1]);
    unsigned long long key = 1469598103934665603ull;
    capb200_tfm_scst_opts o2 = *opts; o2.seed = 0; o2.att_masks = ta.mask;
    StepGraph::mix(key, &o2, sizeof(o2)); StepGraph::mix(key, grads, sizeof(*grads)); StepGraph::mix(key, &e->w, sizeof(e->w));
    const void* ptrs[] = {table, refs, ref_offsets, sample_seq, greedy_seq, sample_logprobs, reward, loss, e->tape, e->ws, e->wblock, e->sg.stage, gst};
    StepGraph::mix(key, ptrs, sizeof(ptrs));
    StepGraph::mix(key, e->grad_events, sizeof(e->grad_events));
    const int dims[] = {B, R, L};
    StepGraph::mix(key, dims, sizeof(dims));
    const int rc_graph = run_step_graph(e->sg, key, opts->seed, &e->launches, gst, [&]() { return tfm_train_step(e, att_s, B, R, ta, grads, gst); });
    if (e->sg.leave(st, gst)) return 1;
    return rc_graph;
}
