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
        CAPB_CHECK_CUDA(cudaStreamSynchronize(st));
        CAPB_CHECK_RANGE();
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
