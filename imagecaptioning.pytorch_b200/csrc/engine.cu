// Engine: owns packed weights + workspaces and drives the whole decode (prologue, T timesteps, beam bookkeeping) as a
// host-sync-free sequence of kernel launches on the caller's stream.  C ABI declared in include/capb200.h.
//
// Reference call stack this replaces:
//   AttModel._sample_beam  AttModel.py:218-256  -> _prepare_feature :114-124, get_logprobs_state :166-176,
//                                                  CaptionModel.beam_search CaptionModel.py:35-209
//   AttModel._sample       AttModel.py:258-352  -> sample_next_word CaptionModel.py:370-407
//   AttModel._forward      AttModel.py:126-164
// Differences that matter for speed, none for results:
//   * image features (fc', att', p_att) are indexed per image, never replicated per beam (repeat_tensors, AttModel.py:241),
//   * the fc' contribution to the attention-LSTM gates is constant over time, so it is contracted once per image in the
//     prologue and enters every step as a per-image row bias (the reference re-multiplies it every step, AttModel.py:626),
//   * no per-step host synchronisation: EOS handling, finished-beam records and history live on the device.
#include <map>
#include <mutex>
#include <vector>

#include "../../include/capb200.h"
#include "common.cuh"
#include "kernels.cuh"
#include "engine_common.cuh"

namespace capb200 {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
const char* last_error_cstr() { return g_last_error.c_str(); }

__global__ void capb_fill_int_kernel(int* p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void capb_load_token_column_kernel(const long long* src, long ld, int col, int n, int* dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (int)src[(long)i * ld + col];
}
__global__ void capb_store_token_column_kernel(const int* src, int n, long long* dst, long ld, int col) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[(long)i * ld + col] = src[i];
}
int store_token_column_launch(const int* src, int n, long long* dst, long ld, int col, cudaStream_t st) {
    if (n <= 0) return 0;
    capb_store_token_column_kernel<<<cdiv(n, 256), 256, 0, st>>>(src, n, dst, ld, col);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int fill_int_launch(int* p, int n, int v, cudaStream_t st) {
    if (n <= 0) return 0;
    capb_fill_int_kernel<<<cdiv(n, 256), 256, 0, st>>>(p, n, v);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int load_token_column_launch(const long long* src, long ld, int col, int n, int* dst, cudaStream_t st) {
    if (n <= 0) return 0;
    capb_load_token_column_kernel<<<cdiv(n, 256), 256, 0, st>>>(src, ld, col, n, dst);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

namespace {

__global__ void add_vec_kernel(const float* a, const float* b, float* o, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] + b[i];
}
__global__ void interleave_gates_kernel(const float* src, float* dst, int H) {      // dst[4*j+g] = src[g*H + j]
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 * H) dst[i] = src[(i & 3) * H + (i >> 2)];
}

__global__ void iota_div_kernel(int* p, int n, int div) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i / div;
}


enum GemmId { G_FC = 0, G_ATT, G_CTX, G_GFC, G_LSTM1, G_H2ATT, G_LSTM2, G_LOGIT, G_CORE, G_LSTM2A, G_LSTM2B, G_COUNT };
constexpr int G_REPORTED = 9;      // ids exposed through capb200_engine_read_profile (the split language-LSTM launches are folded into G_LSTM2)

}  // namespace
}  // namespace capb200

using namespace capb200;


struct capb200_engine {
    capb200_model_cfg cfg{};
    capb200_weights w{};
    int V1 = 0, E = 0, H = 0, A = 0, T = 0;
    int mode = 0;
    bool tc = false;
    bool bound = false;
    long launches = 0;

    // bind-time buffers (owned)
    char* wblock = nullptr;
    size_t wblock_bytes = 0;
    float *bsum_att = nullptr, *bsum_lang = nullptr, *bsum_core = nullptr;
    float *bsum_att_il = nullptr, *bsum_lang_il = nullptr;   // gate-interleaved copies for the fused LSTM epilogue (tensor-core modes)
    float* xgate = nullptr;          // [V+1, 4H] relu(embed) * W_ih[:, 2H:]^T: per-token gate contribution (eval-mode decode)
    long ld_xgate = 0;
    bool use_xgate = true;
    Planes p_fc, p_attw, p_ctx, p_logit, p_a_ih_h, p_a_ih_fc, p_a_ih_x, p_a_hh, p_l_ih_a, p_l_ih_h, p_l_hh, p_h2att, p_i2h, p_h2h;

    // workspace (owned)
    char* ws = nullptr;
    size_t ws_bytes = 0;
    int capB = 0, capRows = 0, capR = 0, capBeam = 0;
    Planes in_fc, in_att;                        // split copies of the user inputs (tensor-core modes)
    Act fc_e, att_e, p_att, g_fc, xt, h0_in, h1_in, h0_out, h1_out, att_res, att_h, gates;
    float *c0[2] = {nullptr, nullptr}, *c1[2] = {nullptr, nullptr};
    long ld_c = 0;
    int* img_of_row = nullptr;
    float* att_score = nullptr;   // [rows, R] attention scores
    DecodeBuffers d;              // tokens / beam state / slab shared with the other families' engines

    GemmTcPlan* plans[G_COUNT] = {nullptr};
    int core_cur = 0;   // which c buffer currently holds the state

    // SCST training tape (owned, grown on demand)
    char* tape = nullptr;
    size_t tape_bytes = 0;
    Tf32Context* tf32 = nullptr;       // tensor maps + transposed operands of the training GEMMs (tensor-core modes)
    StepGraph sg;                                      // CUDA graph of the whole SCST step
    cudaEvent_t grad_events[2] = {nullptr, nullptr};   // caller-owned: recorded when a gradient group is complete (capb200_engine_set_grad_events)
    // Optional overlap of the additive attention (SFU / FMA pipes) with the tensor-core work of the language LSTM that does not depend on it:
    // gates = W_h h_att + W_hh h_lang_prev (side stream, while the attention runs) and then + W_a att_res with the fused cell (main stream).
    bool split_lang = false;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_gfork = nullptr, ev_gjoin = nullptr;     // fork / join of the SCST step's concurrent greedy baseline

    // optional per-GEMM device timing (cudaEvent pairs on the launching stream), off by default
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_ids;          // GEMM id of each recorded pair
    std::vector<double> ev_flops;     // algorithmic FLOPs of each recorded launch
    size_t ev_used = 0;
    double prof_ms[G_COUNT] = {0};
    double prof_flops[G_COUNT] = {0};
    long prof_calls[G_COUNT] = {0};
};

namespace {

void destroy_plans(capb200_engine* e) {
    for (int i = 0; i < G_COUNT; ++i) {
        if (e->plans[i]) { gemm_tc_plan_destroy(e->plans[i]); e->plans[i] = nullptr; }
    }
}

void layout_weights(capb200_engine* e, Arena& a) {
    const int H = e->H, E = e->E, A = e->A, V1 = e->V1;
    const bool updown = e->cfg.family == CAPB200_FAMILY_UPDOWN;
    e->bsum_att = a.take<float>(4 * H);
    e->bsum_lang = a.take<float>(4 * H);
    e->bsum_core = a.take<float>(5 * H);
    e->bsum_att_il = a.take<float>(4 * H);
    e->bsum_lang_il = a.take<float>(4 * H);
    if (updown) { e->ld_xgate = round_up(4 * H, 8); e->xgate = a.take<float>((long)V1 * e->ld_xgate); }
    if (!e->tc) return;
    e->p_logit = carve_planes(a, V1, H);
    if (updown) {
        e->p_fc = carve_planes(a, H, e->cfg.fc_feat_size);
        e->p_attw = carve_planes(a, H, e->cfg.att_feat_size);
        e->p_ctx = carve_planes(a, A, H);
        e->p_a_ih_h = carve_planes(a, 4 * H, H);
        e->p_a_ih_fc = carve_planes(a, 4 * H, H);
        e->p_a_ih_x = carve_planes(a, 4 * H, E);
        e->p_a_hh = carve_planes(a, 4 * H, H);
        e->p_l_ih_a = carve_planes(a, 4 * H, H);
        e->p_l_ih_h = carve_planes(a, 4 * H, H);
        e->p_l_hh = carve_planes(a, 4 * H, H);
        e->p_h2att = carve_planes(a, A, H);
    } else {
        e->p_fc = carve_planes(a, E, e->cfg.fc_feat_size);
        e->p_i2h = carve_planes(a, 5 * H, E);
        e->p_h2h = carve_planes(a, 5 * H, H);
    }
}

int pack(capb200_engine* e, const float* w, long ldw, int rows, int cols, const Planes& p, cudaStream_t st) {
    e->launches++;
    return split_planes_launch(w, ldw, rows, cols, p.hi, p.lo, p.ld, st);
}
// LSTM weight blocks [4H, cols] are stored gate-interleaved (row 4*j+g) so the GEMM epilogue can apply the cell directly
int pack_gates(capb200_engine* e, const float* w, long ldw, int H, int cols, const Planes& p, cudaStream_t st) {
    e->launches++;
    return split_planes_interleave_launch(w, ldw, H, cols, p.hi, p.lo, p.ld, st);
}

void layout_workspace(capb200_engine* e, Arena& a, int B, int rows, int R, int beam) {
    const int H = e->H, E = e->E, A = e->A, T = e->T;
    const bool updown = e->cfg.family == CAPB200_FAMILY_UPDOWN;
    const bool tc = e->tc;
    if (tc) {
        e->in_fc = carve_planes(a, B, e->cfg.fc_feat_size);
        if (updown) e->in_att = carve_planes(a, (long)B * R, e->cfg.att_feat_size);
    }
    e->fc_e.carve(a, B, updown ? H : E, tc);
    if (updown) {
        e->att_e.carve(a, (long)B * R, H, tc);
        e->p_att.carve(a, (long)B * R, A, false);
        e->g_fc.carve(a, B, 4 * H, false);
        e->h1_in.carve(a, rows, H, tc);
        e->h1_out.carve(a, rows, H, tc);
        e->att_res.carve(a, rows, H, tc);
        e->att_h.carve(a, rows, A, false);
    }
    e->xt.carve(a, rows, E, tc);
    e->h0_in.carve(a, rows, H, tc);
    e->h0_out.carve(a, rows, H, tc);
    e->gates.carve(a, rows, 5 * H, false);
    e->ld_c = round_up(H, 8);
    for (int i = 0; i < 2; ++i) {
        e->c0[i] = a.take<float>((long)rows * e->ld_c);
        e->c1[i] = a.take<float>((long)rows * e->ld_c);
    }
    e->img_of_row = a.take<int>(rows);
    e->att_score = a.take<float>((long)rows * (R > 0 ? R : 1));
    e->d.carve(a, B, rows, beam, T);
}

int ensure_workspace(capb200_engine* e, int B, int rows, int R, int beam, cudaStream_t st) {
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
    e->ws_bytes = need;
    Arena real;
    real.base = e->ws;
    layout_workspace(e, real, nB, nRows, nR, nBeam);
    e->capB = nB; e->capRows = nRows; e->capR = nR; e->capBeam = nBeam;
    CAPB_CHECK_CUDA(cudaMemsetAsync(e->ws, 0, need, st));
    return fill_int_launch(e->d.neg1, nRows, -1, st);
}

// ---- GEMM dispatch ------------------------------------------------------------------------------------------------
// `plan_rows` is the row capacity the tensor maps are encoded for; M the rows valid in this launch.
int run_gemm_inner(capb200_engine* e, int id, GemmProblem& g, int plan_rows, cudaStream_t st) {
    e->launches++;
    if (!e->tc) return gemm_simt_launch(g, st);
    if (e->plans[id] == nullptr) {
        GemmProblem planned = g;
        planned.M = plan_rows;
        e->plans[id] = gemm_tc_plan_create(planned, e->mode == CAPB200_MODE_TC_F16X3 ? 3 : 1);
        if (e->plans[id] == nullptr) return 1;
    }
    return gemm_tc_plan_launch(e->plans[id], &g.epi, g.M, st);
}

int run_gemm(capb200_engine* e, int id, GemmProblem& g, int plan_rows, cudaStream_t st) {
    if (!e->profiling) return run_gemm_inner(e, id, g, plan_rows, st);
    if (e->ev_used + 2 > e->ev_pool.size()) {
        for (int i = 0; i < 64; ++i) {
            cudaEvent_t ev;
            CAPB_CHECK_CUDA(cudaEventCreate(&ev));
            e->ev_pool.push_back(ev);
        }
    }
    double k_total = 0;
    for (int s = 0; s < g.nseg; ++s) k_total += g.seg[s].K;
    CAPB_CHECK_CUDA(cudaEventRecord(e->ev_pool[e->ev_used], st));
    const int rc = run_gemm_inner(e, id, g, plan_rows, st);
    CAPB_CHECK_CUDA(cudaEventRecord(e->ev_pool[e->ev_used + 1], st));
    e->ev_used += 2;
    e->ev_ids.push_back(id);
    e->ev_flops.push_back(2.0 * g.M * g.N * k_total);
    return rc;
}

// ---- prologue: _prepare_feature ---------------------------------------------------------------------------------------
int prepare(capb200_engine* e, const float* fc, const float* att, const float* mask, int B, int R, cudaStream_t st) {
    CAPB_NVTX("capb200 prepare_feature (fc_embed, att_embed, ctx2att)");
    const int H = e->H, E = e->E, A = e->A;
    const bool updown = e->cfg.family == CAPB200_FAMILY_UPDOWN;
    const capb200_weights& w = e->w;
    ActView fc_in;  fc_in.f = const_cast<float*>(fc);  fc_in.ld = e->cfg.fc_feat_size;
    if (e->tc) {
        e->launches++;
        if (split_planes_launch(fc, e->cfg.fc_feat_size, B, e->cfg.fc_feat_size, e->in_fc.hi, e->in_fc.lo, e->in_fc.ld, st)) return 1;
    }
    {   // fc_embed: Linear (+ReLU for the attention models; NewFC has a bare Linear, AttModel.py:907)
        GemmProblem g;
        g.M = B; g.N = updown ? H : E; g.nseg = 1;
        g.seg[0] = seg_of(fc_in, w.fc_embed_w, e->cfg.fc_feat_size, e->p_fc, e->cfg.fc_feat_size);
        g.seg[0].A_hi = e->in_fc.hi; g.seg[0].A_lo = e->in_fc.lo; g.seg[0].lda_h = e->in_fc.ld;
        g.epi.bias = w.fc_embed_b; g.epi.relu = updown ? 1 : 0;
        g.epi.C = e->fc_e.v.f; g.epi.ldc = e->fc_e.v.ld; g.epi.C_hi = e->fc_e.v.hi; g.epi.C_lo = e->fc_e.v.lo; g.epi.ldcs = e->fc_e.v.ld;
        if (run_gemm(e, G_FC, g, e->capB, st)) return 1;
    }
    if (!updown) return 0;
    ActView att_in; att_in.f = const_cast<float*>(att); att_in.ld = e->cfg.att_feat_size;
    if (e->tc) {
        e->launches++;
        if (split_planes_launch(att, e->cfg.att_feat_size, B * R, e->cfg.att_feat_size, e->in_att.hi, e->in_att.lo, e->in_att.ld, st)) return 1;
    }
    {   // att_embed: the one consumer of the [B,R,2048] bottom-up tile (AttModel.py:119)
        GemmProblem g;
        g.M = B * R; g.N = H; g.nseg = 1;
        g.seg[0] = seg_of(att_in, w.att_embed_w, e->cfg.att_feat_size, e->p_attw, e->cfg.att_feat_size);
        g.seg[0].A_hi = e->in_att.hi; g.seg[0].A_lo = e->in_att.lo; g.seg[0].lda_h = e->in_att.ld;
        g.epi.bias = w.att_embed_b; g.epi.relu = 1;
        g.epi.C = e->att_e.v.f; g.epi.ldc = e->att_e.v.ld; g.epi.C_hi = e->att_e.v.hi; g.epi.C_lo = e->att_e.v.lo; g.epi.ldcs = e->att_e.v.ld;
        if (run_gemm(e, G_ATT, g, e->capB * e->capR, st)) return 1;
    }
    if (mask != nullptr) {   // pack_wrapper zero-pads the rows of invalid regions (AttModel.py:44-49)
        e->launches++;
        if (mask_rows_launch(e->att_e.v, B, R, H, mask, R, st)) return 1;
    }
    {   // ctx2att
        GemmProblem g;
        g.M = B * R; g.N = A; g.nseg = 1;
        g.seg[0] = seg_of(e->att_e.v, w.ctx2att_w, H, e->p_ctx, H);
        g.epi.bias = w.ctx2att_b;
        g.epi.C = e->p_att.v.f; g.epi.ldc = e->p_att.v.ld;
        if (run_gemm(e, G_CTX, g, e->capB * e->capR, st)) return 1;
    }
    {   // time-invariant part of the attention-LSTM gates: fc' * W_ih[:, H:2H]^T + b_ih + b_hh
        GemmProblem g;
        g.M = B; g.N = 4 * H; g.nseg = 1;
        g.seg[0] = seg_of(e->fc_e.v, w.att_lstm_w_ih + H, E + 2 * H, e->p_a_ih_fc, H);
        g.epi.bias = e->tc ? e->bsum_att_il : e->bsum_att;     // tensor-core modes keep every gate quantity interleaved
        g.epi.C = e->g_fc.v.f; g.epi.ldc = e->g_fc.v.ld;
        if (run_gemm(e, G_GFC, g, e->capB, st)) return 1;
    }
    return 0;
}

// ---- one application of the recurrent core on `rows` rows (rpi rows per image) -----------------------------------------
// tokens: input word per row; src_row: parent row per row (nullptr = identity, e->neg1 = fresh zero state)
int core_step(capb200_engine* e, int rows, int rpi, const int* tokens, const int* src_row, float* logits, long ld_logits, int n_images, int R,
              const float* mask, cudaStream_t st) {
    const int H = e->H, E = e->E, A = e->A, V1 = e->V1;
    const capb200_weights& w = e->w;
    if (e->cfg.family == CAPB200_FAMILY_UPDOWN) {
        StateCopy s0, s1;
        s0.src = e->h0_out.v.f; s0.ld_src = e->h0_out.v.ld; s0.dst = e->h0_in.v;
        s1.src = e->h1_out.v.f; s1.ld_src = e->h1_out.v.ld; s1.dst = e->h1_in.v;
        const bool xg = e->use_xgate;     // word contribution comes from the per-token table instead of a K-segment
        e->launches++;
        if (state_gather_embed_launch(rows, tokens, src_row, w.embed, E, xg ? 0 : E, 1, e->xt.v, H, 2, s0, s1, st)) return 1;
        const int cur = e->core_cur, nxt = cur ^ 1;
        {   // attention LSTM gates: [h_lang_prev | (xt) | h_att_prev] segments + per-image fc' term
            GemmProblem g;
            g.M = rows; g.N = 4 * H;
            g.seg[0] = seg_of(e->h1_in.v, w.att_lstm_w_ih, E + 2 * H, e->p_a_ih_h, H);
            g.seg[1] = seg_of(e->h0_in.v, w.att_lstm_w_hh, H, e->p_a_hh, H);
            g.nseg = 2;
            if (!xg) { g.seg[2] = seg_of(e->xt.v, w.att_lstm_w_ih + 2 * H, E + 2 * H, e->p_a_ih_x, E); g.nseg = 3; }
            g.epi.row_bias = e->g_fc.v.f; g.epi.ld_row_bias = e->g_fc.v.ld; g.epi.rows_per_group = rpi;
            if (e->tc) {     // cell applied in the GEMM epilogue: no gate buffer, no point-wise launch
                g.epi.lstm = 1; g.epi.H = H;
                g.epi.c_prev = e->c0[cur]; g.epi.ld_cprev = e->ld_c; g.epi.src_row = src_row;
                g.epi.c_out = e->c0[nxt]; g.epi.ld_cout = e->ld_c;
                g.epi.gather_bias = xg ? e->xgate : nullptr; g.epi.ld_gb = e->ld_xgate; g.epi.gather_idx = tokens;
                g.epi.h_f = e->h0_out.v.f; g.epi.h_hi = e->h0_out.v.hi; g.epi.h_lo = e->h0_out.v.lo; g.epi.ld_h = e->h0_out.v.ld;
            } else {
                g.epi.C = e->gates.v.f; g.epi.ldc = e->gates.v.ld;
            }
            if (run_gemm(e, G_LSTM1, g, e->capRows, st)) return 1;
        }
        if (!e->tc) {
            e->launches++;
            if (lstm_pointwise_launch(rows, H, e->gates.v.f, e->gates.v.ld, src_row, e->c0[cur], e->ld_c, e->c0[nxt], e->ld_c, e->h0_out.v,
                                      xg ? e->xgate : nullptr, e->ld_xgate, tokens, st)) return 1;
        }
        {   // h2att
            GemmProblem g;
            g.M = rows; g.N = A; g.nseg = 1;
            g.seg[0] = seg_of(e->h0_out.v, w.h2att_w, H, e->p_h2att, H);
            g.epi.bias = w.h2att_b;
            g.epi.C = e->att_h.v.f; g.epi.ldc = e->att_h.v.ld;
            if (run_gemm(e, G_H2ATT, g, e->capRows, st)) return 1;
        }
        const bool split = e->split_lang && e->tc && e->side != nullptr;
        if (split) {   // fork: the part of the language-LSTM gates that does not need the attention result, on the side stream
            CAPB_CHECK_CUDA(cudaEventRecord(e->ev_fork, st));
            CAPB_CHECK_CUDA(cudaStreamWaitEvent(e->side, e->ev_fork, 0));
            GemmProblem g;
            g.M = rows; g.N = 4 * H; g.nseg = 2;
            g.seg[0] = seg_of(e->h0_out.v, w.lang_lstm_w_ih + H, 2 * H, e->p_l_ih_h, H);
            g.seg[1] = seg_of(e->h1_in.v, w.lang_lstm_w_hh, H, e->p_l_hh, H);
            g.epi.bias = e->bsum_lang_il;
            g.epi.C = e->gates.v.f; g.epi.ldc = e->gates.v.ld;          // gate-interleaved partial sums [rows, 4H]
            if (run_gemm(e, G_LSTM2A, g, e->capRows, e->side)) return 1;
            CAPB_CHECK_CUDA(cudaEventRecord(e->ev_join, e->side));
        }
        e->launches += 2;
        if (additive_attention_launch(n_images, rpi, R, A, H, e->att_h.v.f, e->att_h.v.ld, e->p_att.v.f, e->p_att.v.ld, e->att_e.v.f, e->att_e.v.ld,
                                      mask, R, w.alpha_w, w.alpha_b, e->att_score, e->att_res.v, st)) return 1;
        if (split) {   // join: add the attention term and apply the cell
            CAPB_CHECK_CUDA(cudaStreamWaitEvent(st, e->ev_join, 0));
            GemmProblem g;
            g.M = rows; g.N = 4 * H; g.nseg = 1;
            g.seg[0] = seg_of(e->att_res.v, w.lang_lstm_w_ih, 2 * H, e->p_l_ih_a, H);
            g.epi.residual = e->gates.v.f; g.epi.ld_res = e->gates.v.ld;
            g.epi.lstm = 1; g.epi.H = H;
            g.epi.c_prev = e->c1[cur]; g.epi.ld_cprev = e->ld_c; g.epi.src_row = src_row;
            g.epi.c_out = e->c1[nxt]; g.epi.ld_cout = e->ld_c;
            g.epi.h_f = e->h1_out.v.f; g.epi.h_hi = e->h1_out.v.hi; g.epi.h_lo = e->h1_out.v.lo; g.epi.ld_h = e->h1_out.v.ld;
            if (run_gemm(e, G_LSTM2B, g, e->capRows, st)) return 1;
        } else {   // language LSTM gates: [att_res | h_att | h_lang_prev]
            GemmProblem g;
            g.M = rows; g.N = 4 * H; g.nseg = 3;
            g.seg[0] = seg_of(e->att_res.v, w.lang_lstm_w_ih, 2 * H, e->p_l_ih_a, H);
            g.seg[1] = seg_of(e->h0_out.v, w.lang_lstm_w_ih + H, 2 * H, e->p_l_ih_h, H);
            g.seg[2] = seg_of(e->h1_in.v, w.lang_lstm_w_hh, H, e->p_l_hh, H);
            if (e->tc) {
                g.epi.bias = e->bsum_lang_il;
                g.epi.lstm = 1; g.epi.H = H;
                g.epi.c_prev = e->c1[cur]; g.epi.ld_cprev = e->ld_c; g.epi.src_row = src_row;
                g.epi.c_out = e->c1[nxt]; g.epi.ld_cout = e->ld_c;
                g.epi.h_f = e->h1_out.v.f; g.epi.h_hi = e->h1_out.v.hi; g.epi.h_lo = e->h1_out.v.lo; g.epi.ld_h = e->h1_out.v.ld;
            } else {
                g.epi.bias = e->bsum_lang;
                g.epi.C = e->gates.v.f; g.epi.ldc = e->gates.v.ld;
            }
            if (run_gemm(e, G_LSTM2, g, e->capRows, st)) return 1;
        }
        if (!e->tc) {
            e->launches++;
            if (lstm_pointwise_launch(rows, H, e->gates.v.f, e->gates.v.ld, src_row, e->c1[cur], e->ld_c, e->c1[nxt], e->ld_c, e->h1_out.v,
                                      nullptr, 0, nullptr, st)) return 1;
        }
        e->core_cur = nxt;
        {   // vocabulary projection straight into the caller's log-prob storage
            GemmProblem g;
            g.M = rows; g.N = V1; g.nseg = 1;
            g.seg[0] = seg_of(e->h1_out.v, w.logit_w, H, e->p_logit, H);
            g.epi.bias = w.logit_b;
            g.epi.C = logits; g.epi.ldc = ld_logits;
            if (run_gemm(e, G_LOGIT, g, e->capRows, st)) return 1;
        }
        return 0;
    }
    // ---- NewFC: maxout LSTM; a fresh state first consumes the image embedding (AttModel.py:925-936)
    const bool fresh = (src_row == e->d.neg1);
    for (int pass = fresh ? 0 : 1; pass < 2; ++pass) {
        StateCopy s0, s1;
        s0.src = e->h0_out.v.f; s0.ld_src = e->h0_out.v.ld; s0.dst = e->h0_in.v;
        const int* srcs = (pass == 0) ? e->d.neg1 : (fresh ? nullptr : src_row);
        e->launches++;
        if (pass == 0) {
            if (state_gather_embed_launch(rows, e->img_of_row, srcs, e->fc_e.v.f, e->fc_e.v.ld, E, 0, e->xt.v, H, 1, s0, s1, st)) return 1;
        } else {
            if (state_gather_embed_launch(rows, tokens, srcs, w.embed, E, E, 0, e->xt.v, H, 1, s0, s1, st)) return 1;
        }
        const int cur = e->core_cur, nxt = cur ^ 1;
        GemmProblem g;
        g.M = rows; g.N = 5 * H; g.nseg = 2;
        g.seg[0] = seg_of(e->xt.v, w.i2h_w, E, e->p_i2h, E);
        g.seg[1] = seg_of(e->h0_in.v, w.h2h_w, H, e->p_h2h, H);
        g.epi.bias = e->bsum_core;
        g.epi.C = e->gates.v.f; g.epi.ldc = e->gates.v.ld;
        if (run_gemm(e, G_CORE, g, e->capRows, st)) return 1;
        e->launches++;
        if (maxout_pointwise_launch(rows, H, e->gates.v.f, e->gates.v.ld, srcs, e->c0[cur], e->ld_c, e->c0[nxt], e->ld_c, e->h0_out.v, st)) return 1;
        e->core_cur = nxt;
    }
    GemmProblem g;
    g.M = rows; g.N = V1; g.nseg = 1;
    g.seg[0] = seg_of(e->h0_out.v, w.logit_w, H, e->p_logit, H);
    g.epi.bias = w.logit_b;
    g.epi.C = logits; g.epi.ldc = ld_logits;
    return run_gemm(e, G_LOGIT, g, e->capRows, st);
}

bool ensure_side(capb200_engine* e) {
    if (e->side != nullptr && e->ev_fork != nullptr && e->ev_join != nullptr && e->ev_gfork != nullptr && e->ev_gjoin != nullptr) return true;
    if (e->side == nullptr && create_side_stream(&e->side) != cudaSuccess) { (void)cudaGetLastError(); e->side = nullptr; return false; }
    if (e->ev_fork == nullptr && cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); e->ev_fork = nullptr; return false; }
    if (e->ev_join == nullptr && cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); e->ev_join = nullptr; return false; }
    if (e->ev_gfork == nullptr && cudaEventCreateWithFlags(&e->ev_gfork, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); e->ev_gfork = nullptr; return false; }
    if (e->ev_gjoin == nullptr && cudaEventCreateWithFlags(&e->ev_gjoin, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); e->ev_gjoin = nullptr; return false; }
    return true;
}

int check_ready(capb200_engine* e) {
    CAPB_REQUIRE(e != nullptr, "null engine");
    CAPB_REQUIRE(e->bound, "capb200_engine_bind_weights has not been called");
    CAPB_CHECK_RANGE();
    return 0;
}

}  // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" {

const char* capb200_last_error(void) { return g_last_error.c_str(); }
int capb200_abi_version(void) { return CAPB200_ABI_VERSION; }
int capb200_range_status(int reset) { return range_flag_read(reset); }

capb200_engine* capb200_engine_create(const capb200_model_cfg* cfg) {
    if (cfg == nullptr) { set_error("null cfg"); return nullptr; }
    if (cfg->family != CAPB200_FAMILY_UPDOWN && cfg->family != CAPB200_FAMILY_NEWFC) { set_error("unknown model family"); return nullptr; }
    if (cfg->numeric_mode < 0 || cfg->numeric_mode > 2) { set_error("unknown numeric mode"); return nullptr; }
    if (cfg->seq_length < 1 || cfg->seq_length > 64) { set_error("seq_length must be in 1..64"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: the capb200 engine has no CPU fallback");
        return nullptr;
    }
    capb200_engine* e = new capb200_engine();
    e->cfg = *cfg;
    e->V1 = cfg->vocab_size + 1;
    e->E = cfg->input_encoding_size;
    e->H = cfg->rnn_size;
    e->A = cfg->att_hid_size;
    e->T = cfg->seq_length;
    e->mode = cfg->numeric_mode;
    e->tc = cfg->numeric_mode != CAPB200_MODE_SIMT_FP32;
    if (e->tc && getenv("CAPB200_SPLIT_LANG") != nullptr && atoi(getenv("CAPB200_SPLIT_LANG")) != 0) {
        if (create_side_stream(&e->side) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming) == cudaSuccess) e->split_lang = true;
        else (void)cudaGetLastError();
    }
    return e;
}

void capb200_engine_destroy(capb200_engine* e) {
    if (e == nullptr) return;
    destroy_plans(e);
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    cudaFree(e->wblock);
    cudaFree(e->ws);
    if (e->d.loop_exec) cudaGraphExecDestroy(e->d.loop_exec);
    cudaFree(e->d.slab);
    cudaFree(e->tape);
    e->sg.destroy();
    tf32_context_destroy(e->tf32);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->ev_gfork) cudaEventDestroy(e->ev_gfork);
    if (e->ev_gjoin) cudaEventDestroy(e->ev_gjoin);
    if (e->side) cudaStreamDestroy(e->side);
    delete e;
}

long capb200_engine_launch_count(const capb200_engine* e) { return e ? e->launches : 0; }

int capb200_engine_set_profiling(capb200_engine* e, int enable) {
    CAPB_REQUIRE(e != nullptr, "null engine");
    e->profiling = enable != 0;
    return 0;
}

int capb200_engine_read_profile(capb200_engine* e, int reset, double* ms, double* flops, long* calls, int n) {
    CAPB_REQUIRE(e != nullptr && n >= G_REPORTED, "need room for 9 GEMM ids");
    CAPB_CHECK_CUDA(cudaDeviceSynchronize());
    for (size_t i = 0; i < e->ev_ids.size(); ++i) {
        float t = 0.f;
        CAPB_CHECK_CUDA(cudaEventElapsedTime(&t, e->ev_pool[2 * i], e->ev_pool[2 * i + 1]));
        e->prof_ms[e->ev_ids[i]] += t;
        e->prof_flops[e->ev_ids[i]] += e->ev_flops[i];
        e->prof_calls[e->ev_ids[i]] += 1;
    }
    e->ev_ids.clear();
    e->ev_flops.clear();
    e->ev_used = 0;
    for (int i = G_LSTM2A; i <= G_LSTM2B; ++i) {       // the split language-LSTM launches report under the language-LSTM id
        e->prof_ms[G_LSTM2] += e->prof_ms[i]; e->prof_flops[G_LSTM2] += e->prof_flops[i]; e->prof_calls[G_LSTM2] += e->prof_calls[i];
        e->prof_ms[i] = 0; e->prof_flops[i] = 0; e->prof_calls[i] = 0;
    }
    for (int i = 0; i < G_REPORTED; ++i) {
        if (ms) ms[i] = e->prof_ms[i];
        if (flops) flops[i] = e->prof_flops[i];
        if (calls) calls[i] = e->prof_calls[i];
        if (reset) { e->prof_ms[i] = 0; e->prof_flops[i] = 0; e->prof_calls[i] = 0; }
    }
    return 0;
}

int capb200_engine_bind_weights(capb200_engine* e, const capb200_weights* w, void* stream) {
    CAPB_REQUIRE(e != nullptr && w != nullptr, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool updown = e->cfg.family == CAPB200_FAMILY_UPDOWN;
    CAPB_REQUIRE(w->embed && w->fc_embed_w && w->fc_embed_b && w->logit_w && w->logit_b, "missing shared weights");
    if (updown) {
        CAPB_REQUIRE(w->att_embed_w && w->att_embed_b && w->ctx2att_w && w->ctx2att_b && w->att_lstm_w_ih && w->att_lstm_w_hh && w->att_lstm_b_ih &&
                         w->att_lstm_b_hh && w->lang_lstm_w_ih && w->lang_lstm_w_hh && w->lang_lstm_b_ih && w->lang_lstm_b_hh && w->h2att_w &&
                         w->h2att_b && w->alpha_w && w->alpha_b, "missing UpDown weights");
    } else {
        CAPB_REQUIRE(w->i2h_w && w->i2h_b && w->h2h_w && w->h2h_b, "missing NewFC weights");
    }
    e->w = *w;
    const int H = e->H, E = e->E, A = e->A, V1 = e->V1;
    if (e->wblock == nullptr) {
        Arena dry;
        layout_weights(e, dry);
        e->wblock_bytes = dry.off + 256;
        CAPB_CHECK_CUDA(cudaMalloc(&e->wblock, e->wblock_bytes));
        CAPB_CHECK_CUDA(cudaMemsetAsync(e->wblock, 0, e->wblock_bytes, st));
        Arena real;
        real.base = e->wblock;
        layout_weights(e, real);
    }
    if (updown) {
        add_vec_kernel<<<cdiv(4 * H, 256), 256, 0, st>>>(w->att_lstm_b_ih, w->att_lstm_b_hh, e->bsum_att, 4 * H);
        add_vec_kernel<<<cdiv(4 * H, 256), 256, 0, st>>>(w->lang_lstm_b_ih, w->lang_lstm_b_hh, e->bsum_lang, 4 * H);
        interleave_gates_kernel<<<cdiv(4 * H, 256), 256, 0, st>>>(e->bsum_att, e->bsum_att_il, H);
        interleave_gates_kernel<<<cdiv(4 * H, 256), 256, 0, st>>>(e->bsum_lang, e->bsum_lang_il, H);
        e->launches += 4;
    } else {
        add_vec_kernel<<<cdiv(5 * H, 256), 256, 0, st>>>(w->i2h_b, w->h2h_b, e->bsum_core, 5 * H);
        e->launches += 1;
    }
    CAPB_CHECK_CUDA(cudaGetLastError());
    if (e->tc) {
        int rc = pack(e, w->logit_w, H, V1, H, e->p_logit, st);
        if (updown) {
            rc |= pack(e, w->fc_embed_w, e->cfg.fc_feat_size, H, e->cfg.fc_feat_size, e->p_fc, st);
            rc |= pack(e, w->att_embed_w, e->cfg.att_feat_size, H, e->cfg.att_feat_size, e->p_attw, st);
            rc |= pack(e, w->ctx2att_w, H, A, H, e->p_ctx, st);
            rc |= pack_gates(e, w->att_lstm_w_ih, E + 2 * H, H, H, e->p_a_ih_h, st);
            rc |= pack_gates(e, w->att_lstm_w_ih + H, E + 2 * H, H, H, e->p_a_ih_fc, st);
            rc |= pack_gates(e, w->att_lstm_w_ih + 2 * H, E + 2 * H, H, E, e->p_a_ih_x, st);
            rc |= pack_gates(e, w->att_lstm_w_hh, H, H, H, e->p_a_hh, st);
            rc |= pack_gates(e, w->lang_lstm_w_ih, 2 * H, H, H, e->p_l_ih_a, st);
            rc |= pack_gates(e, w->lang_lstm_w_ih + H, 2 * H, H, H, e->p_l_ih_h, st);
            rc |= pack_gates(e, w->lang_lstm_w_hh, H, H, H, e->p_l_hh, st);
            rc |= pack(e, w->h2att_w, H, A, H, e->p_h2att, st);
        } else {
            rc |= pack(e, w->fc_embed_w, e->cfg.fc_feat_size, E, e->cfg.fc_feat_size, e->p_fc, st);
            rc |= pack(e, w->i2h_w, E, 5 * H, E, e->p_i2h, st);
            rc |= pack(e, w->h2h_w, H, 5 * H, H, e->p_h2h, st);
        }
        if (rc) return 1;
    }
    if (updown) {
        // per-token gate table: relu(embed)[V+1,E] * W_ih[:, 2H:2H+E]^T  (one GEMM per bind; replaces a K=E segment in every step)
        const long n = (long)V1 * E;
        const long ldE = round_up(E, 8);
        char* tmp = nullptr;
        const size_t tmp_bytes = (size_t)V1 * ldE * (sizeof(float) + (e->tc ? 2 * sizeof(__half) : 0)) + 1024;
        CAPB_CHECK_CUDA(cudaMallocAsync(&tmp, tmp_bytes, st));
        ActView ev;
        ev.ld = ldE;
        ev.f = reinterpret_cast<float*>(tmp);
        if (e->tc) {
            ev.hi = reinterpret_cast<__half*>(tmp + (size_t)V1 * ldE * sizeof(float));
            ev.lo = ev.hi + (size_t)V1 * ldE;
        }
        int rc = 0;
        if (ldE == E) {
            rc = relu_copy_launch(w->embed, n, ev, st);
        } else {
            CAPB_CHECK_CUDA(cudaMemsetAsync(tmp, 0, tmp_bytes, st));
            for (int v = 0; v < V1 && !rc; ++v) {     // ragged pitch (tiny test configs only): row by row
                ActView rv = ev;
                rv.f += (long)v * ldE; if (rv.hi) { rv.hi += (long)v * ldE; rv.lo += (long)v * ldE; }
                rc = relu_copy_launch(w->embed + (long)v * E, E, rv, st);
            }
        }
        GemmProblem g;
        g.M = V1; g.N = 4 * H; g.nseg = 1;
        g.seg[0] = seg_of(ev, w->att_lstm_w_ih + 2 * H, E + 2 * H, e->p_a_ih_x, E);
        g.epi.C = e->xgate; g.epi.ldc = e->ld_xgate;
        if (!rc) {
            if (!e->tc) {
                rc = gemm_simt_launch(g, st);
            } else {
                GemmTcPlan* plan = gemm_tc_plan_create(g, e->mode == CAPB200_MODE_TC_F16X3 ? 3 : 1);
                rc = plan ? gemm_tc_plan_launch(plan, nullptr, 0, st) : 1;
                if (plan) gemm_tc_plan_destroy(plan);
            }
        }
        e->launches += 2;
        cudaFreeAsync(tmp, st);
        if (rc) return 1;
    }
    if (e->tc && !e->bound) {   // first binding: wait for the conversions and refuse weights outside the fp16 range of the split planes.
        // Re-bindings (a training loop changes the weights every step) must not stall the host: the range flag is host-mapped and every later
        // entry point checks it (check_ready), so an overflow introduced by an optimizer step is reported by the next call instead.
        CAPB_CHECK_CUDA(cudaStreamSynchronize(st));
        CAPB_CHECK_RANGE();
    }
    e->bound = true;
    return 0;
}

int capb200_decode_beam(capb200_engine* e, const float* fc, const float* att, const float* mask, int B, int R, const capb200_beam_opts* opts,
                        long long* seq, float* seq_logprobs, long long* done_seq, int* done_len, float* done_p, float* done_raw, void* stream) {
    if (check_ready(e)) return 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CAPB_REQUIRE(opts != nullptr && fc != nullptr && seq != nullptr, "null argument");
    const int beam = opts->beam_size, keep = opts->sample_n;
    CAPB_REQUIRE(beam >= 1 && beam <= 16 && beam <= e->V1, "beam_size must be in 1..16 and <= V+1");
    CAPB_REQUIRE(keep == 1 || keep == beam, "sample_n must be 1 or beam_size (AttModel.py:223)");
    CAPB_REQUIRE(B >= 1, "empty batch");
    const bool updown = e->cfg.family == CAPB200_FAMILY_UPDOWN;
    if (updown) CAPB_REQUIRE(att != nullptr && R >= 1, "attention features required");
    if (!updown) R = 1;
    const int T = e->T, V1 = e->V1;
    const int rows = B * beam;
    if (ensure_workspace(e, B, rows, R, beam, st)) return 1;
    if (!updown) { iota_div_kernel<<<cdiv(rows, 256), 256, 0, st>>>(e->img_of_row, rows, 1); e->launches++; }
    if (prepare(e, fc, att, mask, B, R, st)) return 1;
    e->core_cur = 0;
    auto core = [&](int nrows, int live, const int* tokens, const int* src_row, int /*t*/, float* logits, long ld) {
        return core_step(e, nrows, live, tokens, src_row, logits, ld, B, R, mask, st);
    };
    return beam_decode_driver(e->d, V1, T, B, beam, keep, opts->penalty_kind, opts->penalty_alpha, seq, seq_logprobs, done_seq, done_len, done_p,
                              done_raw, core, &e->launches, st, e->profiling ? 0ull : loop_graph_key(e->ws, e->wblock, mask, R, (int)e->cfg.family),
                              to_edits(opts->edits), opts->temperature);
}

int capb200_beam_record_logprobs(capb200_engine* e, int image, int rank, float* dst, void* stream) {
    if (check_ready(e)) return 1;
    return beam_record_logprobs(e->d, e->V1, e->T, image, rank, dst, static_cast<cudaStream_t>(stream));
}

int capb200_decode_sample(capb200_engine* e, const float* fc, const float* att, const float* mask, int B, int R, const capb200_sample_opts* opts,
                          const long long* tokens_in, long ld_tok, long long* seq, float* seq_logprobs, float* picked, void* stream) {
    if (check_ready(e)) return 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CAPB_REQUIRE(opts != nullptr && fc != nullptr && seq_logprobs != nullptr, "null argument");
    const int n = opts->sample_n;
    CAPB_REQUIRE(n >= 1 && B >= 1, "empty batch");
    const bool updown = e->cfg.family == CAPB200_FAMILY_UPDOWN;
    if (updown) CAPB_REQUIRE(att != nullptr && R >= 1, "attention features required");
    if (!updown) R = 1;
    const int method = opts->method;
    CAPB_REQUIRE(method >= 0 && method <= 5, "unknown sampling method");
    if (method == CAPB200_SAMPLE_FORCED || method == CAPB200_SAMPLE_TEACHER) CAPB_REQUIRE(tokens_in != nullptr && ld_tok >= 1, "token matrix required");
    if (method != CAPB200_SAMPLE_TEACHER) CAPB_REQUIRE(seq != nullptr, "seq output required");
    if (method == CAPB200_SAMPLE_MULTINOMIAL || method >= CAPB200_SAMPLE_TOPK) CAPB_REQUIRE(opts->temperature > 0.f, "temperature must be positive");
    const int T = e->T, V1 = e->V1;
    const int rows = B * n;
    const int steps = (method == CAPB200_SAMPLE_TEACHER) ? opts->steps : T;
    const long t_out = (method == CAPB200_SAMPLE_TEACHER) ? ld_tok : T;
    CAPB_REQUIRE(steps >= 0 && steps <= t_out, "steps out of range");
    if (ensure_workspace(e, B, rows, R, 1, st)) return 1;
    if (!updown) { iota_div_kernel<<<cdiv(rows, 256), 256, 0, st>>>(e->img_of_row, rows, n); e->launches++; }
    if (prepare(e, fc, att, mask, B, R, st)) return 1;
    e->core_cur = 0;
    auto core = [&](int nrows, int /*live*/, const int* tokens, const int* src_row, int /*t*/, float* logits, long ld) {
        return core_step(e, nrows, n, tokens, src_row, logits, ld, B, R, mask, st);
    };
    return sample_decode_driver(e->d, V1, T, rows, method, opts->temperature, opts->seed, steps, tokens_in, ld_tok, seq, seq_logprobs, picked, core,
                                &e->launches, st, to_edits(opts->edits), opts->top);
}

// ---------------------------------------------------------------------------------------------------------------------
// operator-level entry points
// ---------------------------------------------------------------------------------------------------------------------
int capb200_linear(const float* x, long ldx, const float* w, long ldw, const float* b, float* y, long ldy, int M, int N, int K, int relu, int mode,
                   void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CAPB_REQUIRE(x && w && y && M > 0 && N > 0 && K > 0, "bad argument");
    GemmProblem g;
    g.M = M; g.N = N; g.nseg = 1;
    g.seg[0].A = x; g.seg[0].lda = ldx; g.seg[0].W = w; g.seg[0].ldw = ldw; g.seg[0].K = K;
    g.epi.bias = b; g.epi.relu = relu; g.epi.C = y; g.epi.ldc = ldy;
    if (mode == CAPB200_MODE_SIMT_FP32) return gemm_simt_launch(g, st);
    if (mode == CAPB200_MODE_SKINNY_TF32X3 || mode == CAPB200_MODE_SKINNY_FP32) {
        // the training step's split-K GEMM (no relu epilogue); scratch for the partial sums is allocated per call here
        CAPB_REQUIRE(!relu, "the skinny GEMM has no relu epilogue");
        const size_t cap = (size_t)4 << 20;
        float* part = nullptr;
        CAPB_CHECK_CUDA(cudaMallocAsync(&part, cap * sizeof(float), st));
        const int tb = 1;
        const int rc = gemm_skinny_launch(M, N, 1, &x, &ldx, &w, &ldw, &K, &tb, y, ldy, b, nullptr, 0, 1, 0, part, cap, mode == CAPB200_MODE_SKINNY_TF32X3, st);
        cudaFreeAsync(part, st);
        return rc;
    }
    if (mode == CAPB200_MODE_TF32X3_TC || mode == CAPB200_MODE_TF32X3_TC_DGRAD || mode == CAPB200_MODE_TF32X3_TC_WGRAD) {
        // the training steps' tcgen05 kind::tf32 kernel (gemm_tf32.cu) through the same helper the engines use:
        //   TF32X3_TC        y[M,N]  = x[M,K] w[N,K]^T + b                         (forward)
        //   TF32X3_TC_DGRAD  y[M,N]  = x[M,K] w'[K,N]       with w' passed as `w`  (input gradient: W stored [out = K, in = N], cached transpose)
        //   TF32X3_TC_WGRAD  y[M,N]  = x'[K,M]^T w'[K,N]    with x', w' row-major   (weight gradient: dY = x', X = w', transposed per call)
        CAPB_REQUIRE(!relu, "the tf32 GEMM has no relu epilogue");
        Tf32Context* ctx = tf32_context_create();
        Skinny sk{nullptr, 0, 1, st};
        sk.ctx = ctx;
        int rc;
        if (mode == CAPB200_MODE_TF32X3_TC) rc = sk.lin(x, ldx, w, ldw, b, y, ldy, M, N, K, 0);
        else if (mode == CAPB200_MODE_TF32X3_TC_DGRAD) rc = sk.dgrad(M, N, K, x, ldx, w, ldw, y, ldy, 0);
        else rc = sk.wgrad(M, N, K, x, ldx, w, ldw, y, ldy, 0);
        if (tf32_context_launches(ctx) == 0 && !rc) { set_error("capb200_linear: the operands are not TMA-compatible, the tcgen05 tf32 kernel did not run"); rc = 1; }
        cudaStreamSynchronize(st);
        tf32_context_destroy(ctx);
        return rc;
    }
    CAPB_REQUIRE(mode == CAPB200_MODE_TC_F16X3 || mode == CAPB200_MODE_TC_F16X1, "unknown mode");
    const long ldh = round_up(K, 64);
    __half* scratch = nullptr;
    const size_t elems = (size_t)(M + N) * ldh * 2;
    CAPB_CHECK_CUDA(cudaMallocAsync(&scratch, elems * sizeof(__half), st));
    __half* xh = scratch; __half* xl = xh + (size_t)M * ldh;
    __half* wh = xl + (size_t)M * ldh; __half* wl = wh + (size_t)N * ldh;
    int rc = split_planes_launch(x, ldx, M, K, xh, xl, ldh, st) | split_planes_launch(w, ldw, N, K, wh, wl, ldh, st);
    g.seg[0].A_hi = xh; g.seg[0].A_lo = xl; g.seg[0].lda_h = ldh;
    g.seg[0].W_hi = wh; g.seg[0].W_lo = wl; g.seg[0].ldw_h = ldh;
    GemmTcPlan* plan = rc ? nullptr : gemm_tc_plan_create(g, mode == CAPB200_MODE_TC_F16X3 ? 3 : 1);
    if (plan == nullptr) rc = 1;
    if (!rc) rc = gemm_tc_plan_launch(plan, nullptr, 0, st);
    if (plan) gemm_tc_plan_destroy(plan);
    cudaFreeAsync(scratch, st);
    return rc;
}

int capb200_bench_linear(const float* x, const float* w, const float* b, float* y, int M, int N, int K, int mode, int iters, float* ms_per_launch,
                         void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CAPB_REQUIRE(x && w && y && ms_per_launch && M > 0 && N > 0 && K > 0 && iters > 0, "bad argument");
    if (mode == CAPB200_MODE_TF32X3_TC || mode == CAPB200_MODE_SKINNY_TF32X3) {
        // the training GEMMs: tcgen05 kind::tf32 kernel (gemm_tf32.cu) or the mma.sync split-K kernel it replaced, timed back to back
        Tf32Context* ctx = mode == CAPB200_MODE_TF32X3_TC ? tf32_context_create() : nullptr;
        float* scratch = nullptr;
        const size_t cap = (size_t)4 << 20;
        CAPB_CHECK_CUDA(cudaMalloc(&scratch, cap * sizeof(float)));
        Skinny sk{scratch, cap, 1, st};
        sk.ctx = ctx;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        int rc = 0;
        for (int i = 0; i < 3 + iters && !rc; ++i) {
            if (i == 3) cudaEventRecord(e0, st);
            rc = sk.lin(x, K, w, K, b, y, N, M, N, K, 0);
        }
        cudaEventRecord(e1, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { set_error(std::string("bench_linear: ") + cudaGetErrorString(cudaGetLastError())); rc = 1; }
        float ms = 0.f;
        if (!rc) { cudaEventElapsedTime(&ms, e0, e1); *ms_per_launch = ms / iters; }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaFree(scratch);
        tf32_context_destroy(ctx);
        return rc;
    }
    GemmProblem g;
    g.M = M; g.N = N; g.nseg = 1;
    g.seg[0].A = x; g.seg[0].lda = K; g.seg[0].W = w; g.seg[0].ldw = K; g.seg[0].K = K;
    g.epi.bias = b; g.epi.C = y; g.epi.ldc = N;
    const long ldh = round_up(K, 64);
    __half* scratch = nullptr;
    GemmTcPlan* plan = nullptr;
    int rc = 0;
    if (mode != CAPB200_MODE_SIMT_FP32) {
        CAPB_CHECK_CUDA(cudaMalloc(&scratch, (size_t)(M + N) * ldh * 2 * sizeof(__half)));
        __half* xh = scratch; __half* xl = xh + (size_t)M * ldh;
        __half* wh = xl + (size_t)M * ldh; __half* wl = wh + (size_t)N * ldh;
        rc = split_planes_launch(x, K, M, K, xh, xl, ldh, st) | split_planes_launch(w, K, N, K, wh, wl, ldh, st);
        g.seg[0].A_hi = xh; g.seg[0].A_lo = xl; g.seg[0].lda_h = ldh;
        g.seg[0].W_hi = wh; g.seg[0].W_lo = wl; g.seg[0].ldw_h = ldh;
        plan = rc ? nullptr : gemm_tc_plan_create(g, mode == CAPB200_MODE_TC_F16X3 ? 3 : 1);
        if (plan == nullptr) rc = 1;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3 + iters && !rc; ++i) {
        if (i == 3) cudaEventRecord(e0, st);
        rc = plan ? gemm_tc_plan_launch(plan, nullptr, 0, st) : gemm_simt_launch(g, st);
    }
    cudaEventRecord(e1, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { set_error(std::string("bench_linear: ") + cudaGetErrorString(cudaGetLastError())); rc = 1; }
    float ms = 0.f;
    if (!rc) { cudaEventElapsedTime(&ms, e0, e1); *ms_per_launch = ms / iters; }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (plan) gemm_tc_plan_destroy(plan);
    if (scratch) cudaFree(scratch);
    return rc;
}

int capb200_gemm_trace(const float* x, const float* w, float* y, int M, int N, int K, unsigned long long* trace_host, int n_slots, void* stream) {
    // one traced launch of the decode GEMM (3-pass tcgen05, CTA pairs) after three warm launches: trace_host[296][16] %globaltimer stamps
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CAPB_REQUIRE(x && w && y && trace_host && n_slots >= 296 * 16, "bad argument");
    GemmProblem g;
    g.M = M; g.N = N; g.nseg = 1;
    g.seg[0].A = x; g.seg[0].lda = K; g.seg[0].W = w; g.seg[0].ldw = K; g.seg[0].K = K;
    g.epi.C = y; g.epi.ldc = N;
    const long ldh = round_up(K, 64);
    __half* scratch = nullptr;
    unsigned long long* trace = nullptr;
    CAPB_CHECK_CUDA(cudaMalloc(&scratch, (size_t)(M + N) * ldh * 2 * sizeof(__half)));
    CAPB_CHECK_CUDA(cudaMalloc(&trace, sizeof(unsigned long long) * 296 * 16));
    CAPB_CHECK_CUDA(cudaMemsetAsync(trace, 0, sizeof(unsigned long long) * 296 * 16, st));
    __half* xh = scratch; __half* xl = xh + (size_t)M * ldh;
    __half* wh = xl + (size_t)M * ldh; __half* wl = wh + (size_t)N * ldh;
    int rc = split_planes_launch(x, K, M, K, xh, xl, ldh, st) | split_planes_launch(w, K, N, K, wh, wl, ldh, st);
    g.seg[0].A_hi = xh; g.seg[0].A_lo = xl; g.seg[0].lda_h = ldh;
    g.seg[0].W_hi = wh; g.seg[0].W_lo = wl; g.seg[0].ldw_h = ldh;
    GemmTcPlan* plan = rc ? nullptr : gemm_tc_plan_create(g, 3);
    if (plan == nullptr) rc = 1;
    for (int i = 0; i < 3 && !rc; ++i) rc = gemm_tc_plan_launch(plan, nullptr, 0, st);
    if (!rc) {
        GemmEpilogue ep = g.epi;
        ep.trace = trace;
        rc = gemm_tc_plan_launch(plan, &ep, 0, st);
    }
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) { set_error("gemm_trace: kernel failed"); rc = 1; }
    if (!rc) CAPB_CHECK_CUDA(cudaMemcpy(trace_host, trace, sizeof(unsigned long long) * 296 * 16, cudaMemcpyDeviceToHost));
    if (plan) gemm_tc_plan_destroy(plan);
    cudaFree(scratch);
    cudaFree(trace);
    return rc;
}

int capb200_lstm_cell(const float* x, int Kx, const float* h, const float* c, const float* w_ih, const float* w_hh, const float* b_ih,
                      const float* b_hh, float* h_out, float* c_out, int M, int H, int mode, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CAPB_REQUIRE(x && h && c && w_ih && w_hh && b_ih && b_hh && h_out && c_out, "null argument");
    float* gates = nullptr;
    CAPB_CHECK_CUDA(cudaMallocAsync(&gates, sizeof(float) * ((size_t)M * 4 * H + 4 * H), st));
    float* bsum = gates + (size_t)M * 4 * H;
    add_vec_kernel<<<cdiv(4 * H, 256), 256, 0, st>>>(b_ih, b_hh, bsum, 4 * H);
    int rc = capb200_linear(x, Kx, w_ih, Kx, bsum, gates, 4 * H, M, 4 * H, Kx, 0, mode, stream);
    float* g2 = nullptr;
    if (!rc) {
        // second contraction accumulated through a temporary: gates += h * w_hh^T
        CAPB_CHECK_CUDA(cudaMallocAsync(&g2, sizeof(float) * (size_t)M * 4 * H, st));
        rc = capb200_linear(h, H, w_hh, H, nullptr, g2, 4 * H, M, 4 * H, H, 0, mode, stream);
        if (!rc) add_vec_kernel<<<cdiv(M * 4 * H, 256), 256, 0, st>>>(gates, g2, gates, M * 4 * H);
    }
    if (!rc) {
        ActView ho; ho.f = h_out; ho.ld = H;
        rc = lstm_pointwise_launch(M, H, gates, 4 * H, nullptr, c, H, c_out, H, ho, nullptr, 0, nullptr, st);
    }
    if (g2) cudaFreeAsync(g2, st);
    cudaFreeAsync(gates, st);
    return rc;
}

int capb200_additive_attention(const float* att_h, const float* p_att, const float* att, const float* mask, const float* alpha_w,
                               const float* alpha_b, float* out, int n_images, int rows_per_image, int R, int A, int H, void* stream) {
    CAPB_REQUIRE(att_h && p_att && att && alpha_w && alpha_b && out, "null argument");
    ActView o; o.f = out; o.ld = H;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* scratch = nullptr;
    CAPB_CHECK_CUDA(cudaMallocAsync(&scratch, sizeof(float) * (size_t)n_images * rows_per_image * R, st));
    const int rc = additive_attention_launch(n_images, rows_per_image, R, A, H, att_h, A, p_att, A, att, H, mask, R, alpha_w, alpha_b, scratch, o, st);
    cudaFreeAsync(scratch, st);
    return rc;
}

int capb200_log_softmax_topk(float* logits, long ld, int rows, int V1, int twice, int k, float* top_val, int* top_idx, void* stream) {
    CAPB_REQUIRE(logits != nullptr && rows > 0 && V1 > 0, "bad argument");
    VocabStepArgs va;
    va.rows = rows; va.V1 = V1; va.logits = logits; va.ld = ld; va.twice = twice; va.topk = k; va.top_val = top_val; va.top_idx = top_idx;
    return vocab_step_launch(va, static_cast<cudaStream_t>(stream));
}

// =====================================================================================================================
// SCST training step (UpDown)
// =====================================================================================================================
namespace {

struct Tape {
    int* tok; float *xt, *g1, *h0, *c0, *atth, *alpha, *attres, *g2, *h1, *c1, *out;          // forward, [T][N][.] except out [N][T][H]
    float *fc_e, *att_e, *p_att, *g_fc, *gl, *glp;                                              // prologue + greedy scratch
    float *DL, *dOUT, *DG1, *DG2, *DATTH, *dh0, *dc0, *dh1, *dc1, *tmpH, *dX2, *dxt, *d_att_e, *d_p_att, *S, *d_fc_e, *dpre_att, *dpre_fc, *mask_sum, *dalpha, *skinny, *item_loss;
    size_t skinny_floats;
    double* scores;
    long long* gseq_dummy;
    int *s_tokens, *s_unfinished, *s_forced;   // sampling-loop state of the train step (its own copies: the greedy baseline runs concurrently)
    float* s_att_score;
    float *row_loss, *row_msum, *row_coef;     // drop_worst: per-row loss, mask count and gradient coefficient
};

void layout_tape(Tape& tp, Arena& a, int B, int R, int N, int T, int E, int H, int A, int V1, int F_att, int F_fc) {
    const long TN = (long)T * N, BR = (long)B * R;
    tp.tok = a.take<int>(TN);
    tp.xt = a.take<float>(TN * E); tp.g1 = a.take<float>(TN * 4 * H); tp.h0 = a.take<float>(TN * H); tp.c0 = a.take<float>(TN * H);
    tp.atth = a.take<float>(TN * A); tp.alpha = a.take<float>(TN * R); tp.attres = a.take<float>(TN * H);
    tp.g2 = a.take<float>(TN * 4 * H); tp.h1 = a.take<float>(TN * H); tp.c1 = a.take<float>(TN * H); tp.out = a.take<float>(TN * H);
    tp.fc_e = a.take<float>((long)B * H); tp.att_e = a.take<float>(BR * H); tp.p_att = a.take<float>(BR * A); tp.g_fc = a.take<float>((long)B * 4 * H);
    tp.gl = a.take<float>((long)N * H); tp.glp = a.take<float>((long)B * T * V1);
    tp.DL = a.take<float>(TN * V1); tp.dOUT = a.take<float>(TN * H); tp.DG1 = a.take<float>(TN * 4 * H); tp.DG2 = a.take<float>(TN * 4 * H);
    tp.DATTH = a.take<float>(TN * A);
    tp.dh0 = a.take<float>((long)N * H); tp.dc0 = a.take<float>((long)N * H); tp.dh1 = a.take<float>((long)N * H); tp.dc1 = a.take<float>((long)N * H);
    tp.tmpH = a.take<float>((long)N * H); tp.dX2 = a.take<float>((long)N * 2 * H); tp.dxt = a.take<float>((long)N * E);
    tp.d_att_e = a.take<float>(BR * H); tp.d_p_att = a.take<float>(BR * A); tp.S = a.take<float>((long)B * 4 * H); tp.d_fc_e = a.take<float>((long)B * H);
    tp.dpre_att = a.take<float>(BR * H); tp.dpre_fc = a.take<float>((long)B * H); tp.mask_sum = a.take<float>(8);
    tp.scores = a.take<double>((long)N + B);
    tp.dalpha = a.take<float>((long)N * R);
    tp.item_loss = a.take<float>(TN);
    tp.skinny_floats = (size_t)4 << 20;                       // split-K partial sums (16 MB)
    tp.skinny = a.take<float>((long)tp.skinny_floats);
    tp.s_tokens = a.take<int>(N); tp.s_unfinished = a.take<int>(N); tp.s_forced = a.take<int>(N);
    tp.s_att_score = a.take<float>((long)N * R);
    tp.row_loss = a.take<float>(N); tp.row_msum = a.take<float>(N); tp.row_coef = a.take<float>(N);
    (void)F_att; (void)F_fc;
}

}  // namespace

extern "C" int capb200_dropout_mask(float* mask, long n, unsigned long long seed, int site, int step, float p, void* stream) {
    CAPB_REQUIRE(mask != nullptr && n > 0, "bad argument");
    if (dropout_salt_set_all(0ull, static_cast<cudaStream_t>(stream))) return 1;      // a graph replay of an SCST step may have left a salt behind
    return dropout_mask_launch(mask, n, seed, (unsigned)site, (unsigned)step, p, static_cast<cudaStream_t>(stream));
}

namespace {

// One training step on the tape: SCST (sampled tokens, reward-weighted loss) or XE (teacher-forced tokens, cross-entropy).
struct TrainArgs {
    bool xe = false;
    int n = 1;                 // rows per image: train_sample_n (SCST) or seq_per_img (XE)
    int T = 0;                 // steps evaluated (and columns of the tape)
    int Tl = 0;                // columns of the log-prob output [N, Tl, V1]
    float p = 0.f, temperature = 1.f, upstream = 1.f, smoothing = 0.f;
    unsigned long long seed = 0;
    // SCST
    bool greedy_baseline = true;
    const capb200_cider_table* table = nullptr;
    const int* refs = nullptr; const int* ref_offsets = nullptr; int L = 0;
    long long* sample_seq = nullptr; long long* greedy_seq = nullptr; float* reward = nullptr;
    const long long* forced = nullptr;      // replay these samples instead of drawing
    const float* mask = nullptr;            // [B, R] region mask or null
    float ss_prob = 0.f;                    // XE: scheduled sampling probability
    long long* tokens_used = nullptr;       // XE: optional [N, Tl] record of the words fed
    int keep = 0;                           // drop_worst: rows kept (0 = reduction 'mean')
    float* row_loss = nullptr;              // drop_worst: optional per-row loss output
    // XE
    const long long* labels = nullptr; long ld_labels = 0; const float* masks = nullptr; long ld_masks = 0;
    float* logprobs = nullptr; float* loss = nullptr;
};

int updown_train_step(capb200_engine* e, const float* fc, const float* att, int B, int R, const TrainArgs& ta, const capb200_updown_grads* grads,
                      cudaStream_t st) {
    void* stream = static_cast<void*>(st);
    const int n = ta.n, N = B * n, T = ta.T, E = e->E, H = e->H, A = e->A, V1 = e->V1;
    const int Fa = e->cfg.att_feat_size, Ff = e->cfg.fc_feat_size;
    const float p = ta.p;
    const float keep_scale = 1.0f / (1.0f - p);
    const unsigned long long seed = ta.seed;
    const capb200_weights& w = e->w;
    float* const sample_logprobs = ta.logprobs;
    const long ld_lp = (long)ta.Tl * V1;
    long long* const sample_seq = ta.sample_seq;
    long long* const greedy_seq = ta.greedy_seq;
    float* const reward = ta.reward;
    float* const loss = ta.loss;

    // ---- (1) greedy baseline, eval mode (no dropout): the regular decode path
    {
        Arena dry; Tape t0; layout_tape(t0, dry, B, R, N, T, E, H, A, V1, Fa, Ff);
        if (dry.off + 256 > e->tape_bytes) {
            CAPB_CHECK_CUDA(cudaStreamSynchronize(st));
            if (e->tape) CAPB_CHECK_CUDA(cudaFree(e->tape));
            e->tape = nullptr;
            CAPB_CHECK_CUDA(cudaMalloc(&e->tape, dry.off + 256));
            e->tape_bytes = dry.off + 256;
        }
    }
    Arena ar; ar.base = e->tape;
    Tape tp; layout_tape(tp, ar, B, R, N, T, E, H, A, V1, Fa, Ff);
    if (ensure_workspace(e, B, N, R, 1, st)) return 1;        // decode workspace sized before anything is in flight
    bool greedy_on_side = false;
    if (!ta.xe && ta.greedy_baseline) {
        // The eval-mode greedy baseline (B rows, the regular decode path with its own workspace) and the train-mode sampling forward (B*n rows,
        // on the tape) are independent chains of small, latency-bound kernels: the baseline runs on a side stream and joins before the reward.
        CAPB_NVTX("capb200 scst: greedy baseline (eval mode, side stream)");
        capb200_sample_opts so; memset(&so, 0, sizeof(so)); so.edits.unk_col = -1; so.sample_n = 1; so.method = CAPB200_SAMPLE_GREEDY; so.temperature = 1.f; so.seed = 0; so.steps = T;
        cudaStream_t gs = st;
        static const bool serial = getenv("CAPB200_SCST_SERIAL_GREEDY") != nullptr;
        if (!serial && ensure_side(e)) {
            CAPB_CHECK_CUDA(cudaEventRecord(e->ev_gfork, st));
            CAPB_CHECK_CUDA(cudaStreamWaitEvent(e->side, e->ev_gfork, 0));
            gs = e->side;
            greedy_on_side = true;
        }
        CAPB_CHECK_CUDA(cudaMemsetAsync(tp.glp, 0, sizeof(float) * (size_t)B * T * V1, gs));
        CAPB_CHECK_CUDA(cudaMemsetAsync(greedy_seq, 0, sizeof(long long) * (size_t)B * T, gs));
        if (capb200_decode_sample(e, fc, att, ta.mask, B, R, &so, nullptr, 0, greedy_seq, tp.glp, nullptr, static_cast<void*>(gs))) return 1;
        if (greedy_on_side) CAPB_CHECK_CUDA(cudaEventRecord(e->ev_gjoin, e->side));
    }
    if (e->tc && e->tf32 == nullptr) e->tf32 = tf32_context_create();
    tf32_context_new_step(e->tf32);                                          // the weights may have changed since the last step
    const long tf32_l0 = tf32_context_launches(e->tf32);
    Skinny sk{tp.skinny, tp.skinny_floats, e->tc ? 1 : 0, st};              // tcgen05 3xTF32 GEMMs unless the engine is in simt_fp32 mode
    sk.ctx = e->tf32;

    // ---- (2) train-mode prologue: fc_embed / att_embed with dropout, ctx2att, per-image gate term
    nvtxRangePushA("capb200 train step: forward on the tape");
    const long BR = (long)B * R;
    if (sk.lin(fc, Ff, w.fc_embed_w, Ff, w.fc_embed_b, tp.fc_e, H, B, H, Ff, 0)) return 1;
    if (relu_copy_launch(tp.fc_e, (long)B * H, ActView{tp.fc_e, nullptr, nullptr, H}, st)) return 1;
    if (dropout_apply_launch(tp.fc_e, B, H, H, seed, 0, 0, p, st)) return 1;
    if (sk.lin(att, Fa, w.att_embed_w, Fa, w.att_embed_b, tp.att_e, H, (int)BR, H, Fa, 0)) return 1;
    if (relu_copy_launch(tp.att_e, BR * H, ActView{tp.att_e, nullptr, nullptr, H}, st)) return 1;
    if (dropout_apply_launch(tp.att_e, (int)BR, H, H, seed, 1, 0, p, st)) return 1;
    if (ta.mask != nullptr) {      // pack_wrapper: the embedding of a padded region is exactly zero (AttModel.py:44-49); relu'(0) = 0 keeps its gradient zero
        if (mask_rows_launch(ActView{tp.att_e, nullptr, nullptr, H}, B, R, H, ta.mask, R, st)) return 1;
        e->launches++;
    }
    if (sk.lin(tp.att_e, H, w.ctx2att_w, H, w.ctx2att_b, tp.p_att, A, (int)BR, A, H, 0)) return 1;
    if (sk.lin(tp.fc_e, H, w.att_lstm_w_ih + H, E + 2 * H, e->bsum_att, tp.g_fc, 4 * H, B, 4 * H, H, 0)) return 1;
    e->launches += 8;

    // ---- (3) T sampling steps with the tape
    const long NH = (long)N * H;
    CAPB_CHECK_CUDA(cudaMemsetAsync(tp.s_tokens, 0, sizeof(int) * N, st));
    for (int t = 0; t < T; ++t) {
        int* tok = tp.tok + (long)t * N;
        if (ta.xe) {
            if (t >= 1 && ta.ss_prob > 0.f) {      // scheduled sampling: draw from the model's own previous prediction (AttModel.py:145-154)
                if (ss_select_launch(N, V1, sample_logprobs + (long)(t - 1) * V1, ld_lp, ta.labels, ta.ld_labels, t, seed, ta.ss_prob, tok, st)) return 1;
            } else if (load_token_column_launch(ta.labels, ta.ld_labels, t, N, tok, st)) return 1;
            if (ta.tokens_used != nullptr && store_token_column_launch(tok, N, ta.tokens_used, ta.Tl, t, st)) return 1;
        }
        else CAPB_CHECK_CUDA(cudaMemcpyAsync(tok, tp.s_tokens, sizeof(int) * N, cudaMemcpyDeviceToDevice, st));
        float* xt = tp.xt + (long)t * N * E;
        float* g1 = tp.g1 + (long)t * N * 4 * H;
        float* g2 = tp.g2 + (long)t * N * 4 * H;
        const float* h0p = t ? tp.h0 + (long)(t - 1) * NH : nullptr;
        const float* h1p = t ? tp.h1 + (long)(t - 1) * NH : nullptr;
        const float* c0p = t ? tp.c0 + (long)(t - 1) * NH : nullptr;
        const float* c1p = t ? tp.c1 + (long)(t - 1) * NH : nullptr;
        float *h0 = tp.h0 + (long)t * NH, *c0 = tp.c0 + (long)t * NH, *h1 = tp.h1 + (long)t * NH, *c1 = tp.c1 + (long)t * NH;
        if (embed_relu_dropout_launch(N, E, tok, w.embed, xt, seed, (unsigned)t, p, st)) return 1;
        // gates1 = g_fc[img] + xt W_x^T (+ h_lang_prev W_h^T + h_att_prev W_hh^T)
        {
            GemmProblem g; g.M = N; g.N = 4 * H; g.nseg = 1;
            g.seg[0].A = xt; g.seg[0].lda = E; g.seg[0].W = w.att_lstm_w_ih + 2 * H; g.seg[0].ldw = E + 2 * H; g.seg[0].K = E;
            if (t) {
                g.seg[1].A = h1p; g.seg[1].lda = H; g.seg[1].W = w.att_lstm_w_ih; g.seg[1].ldw = E + 2 * H; g.seg[1].K = H;
                g.seg[2].A = h0p; g.seg[2].lda = H; g.seg[2].W = w.att_lstm_w_hh; g.seg[2].ldw = H; g.seg[2].K = H;
                g.nseg = 3;
            }
            g.epi.row_bias = tp.g_fc; g.epi.ld_row_bias = 4 * H; g.epi.rows_per_group = n;
            g.epi.C = g1; g.epi.ldc = 4 * H;
            if (sk.gates(g)) return 1;
        }
        if (lstm_pointwise_launch(N, H, g1, 4 * H, nullptr, c0p, H, c0, H, ActView{h0, nullptr, nullptr, H}, nullptr, 0, nullptr, st)) return 1;
        float* atth = tp.atth + (long)t * N * A;
        if (sk.lin(h0, H, w.h2att_w, H, w.h2att_b, atth, A, N, A, H, 0)) return 1;
        float* attres = tp.attres + (long)t * NH;
        if (additive_attention_launch(B, n, R, A, H, atth, A, tp.p_att, A, tp.att_e, H, ta.mask, R, w.alpha_w, w.alpha_b, tp.s_att_score,
                                      ActView{attres, nullptr, nullptr, H}, st, tp.alpha + (long)t * N * R)) return 1;
        {
            GemmProblem g; g.M = N; g.N = 4 * H; g.nseg = 2;
            g.seg[0].A = attres; g.seg[0].lda = H; g.seg[0].W = w.lang_lstm_w_ih; g.seg[0].ldw = 2 * H; g.seg[0].K = H;
            g.seg[1].A = h0; g.seg[1].lda = H; g.seg[1].W = w.lang_lstm_w_ih + H; g.seg[1].ldw = 2 * H; g.seg[1].K = H;
            if (t) { g.seg[2].A = h1p; g.seg[2].lda = H; g.seg[2].W = w.lang_lstm_w_hh; g.seg[2].ldw = H; g.seg[2].K = H; g.nseg = 3; }
            g.epi.bias = e->bsum_lang;
            g.epi.C = g2; g.epi.ldc = 4 * H;
            if (sk.gates(g)) return 1;
        }
        if (lstm_pointwise_launch(N, H, g2, 4 * H, nullptr, c1p, H, c1, H, ActView{h1, nullptr, nullptr, H}, nullptr, 0, nullptr, st)) return 1;
        // core output = dropout(h_lang), stored in (n, t) order for the batched logit backward
        float* out = tp.out + (long)t * H;
        if (dropout_copy_launch(h1, H, out, (long)T * H, N, H, seed, 3, (unsigned)t, p, st)) return 1;
        float* logits = sample_logprobs + (long)t * V1;
        if (sk.lin(out, (long)T * H, w.logit_w, H, w.logit_b, logits, ld_lp, N, V1, H, 0)) return 1;
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
        e->launches += 12;
    }

    // ---- (4) reward and loss
    nvtxRangePop();
    CAPB_NVTX("capb200 train step: reward, loss, backward through time, weight gradients");
    const long TN = (long)T * N;
    const capb200_updown_grads& G = *grads;
    if (ta.xe) {
        if (xe_loss_backward_launch(sample_logprobs, ld_lp, ta.labels, ta.ld_labels, ta.masks, ta.ld_masks, N, T, ta.Tl, V1, ta.smoothing, ta.upstream,
                                    tp.mask_sum, tp.item_loss, tp.DL, loss, st, ta.keep, ta.row_loss ? ta.row_loss : tp.row_loss, tp.row_msum, tp.row_coef)) return 1;
    } else {
        if (greedy_on_side) CAPB_CHECK_CUDA(cudaStreamWaitEvent(st, e->ev_gjoin, 0));     // join: the reward needs the baseline captions
        if (cider_reward_launch(ta.table->t, sample_seq, N, ta.greedy_baseline ? greedy_seq : nullptr, B, T, ta.refs, ta.ref_offsets, ta.L, tp.scores, reward,
                                T, T, st)) return 1;
        float* rl = ta.keep > 0 ? (ta.row_loss ? ta.row_loss : tp.row_loss) : nullptr;
        if (reward_criterion_fwd_launch(sample_logprobs, ld_lp, V1, sample_seq, reward, N, T, loss, rl, tp.mask_sum, st)) return 1;
        if (ta.keep > 0 && scst_drop_worst_launch(sample_seq, rl, N, T, ta.keep, ta.upstream, tp.row_msum, tp.row_coef, loss, st)) return 1;
        // ---- (5) backward: logit layer, batched over all (n, t)
        if (scst_dlogits_launch(sample_logprobs, ld_lp, sample_seq, reward, tp.mask_sum, ta.upstream, N, T, V1, tp.DL, st, ta.keep > 0 ? tp.row_coef : nullptr)) return 1;
    }
    e->launches += 3;
    if (sk.dgrad((int)TN, H, V1, tp.DL, V1, w.logit_w, H, tp.dOUT, H, 0)) return 1;          // dOUT = DL * W
    if (sk.wgrad(V1, H, (int)TN, tp.DL, V1, tp.out, H, G.logit_w, H, 0)) return 1;            // dW = DL^T * OUT
    if (colsum_launch((int)TN, V1, tp.DL, V1, G.logit_b, 0, st)) return 1;
    if (record_group_event(e->grad_events[0], st)) return 1;                                  // group 0 (logit) is final
    CAPB_CHECK_CUDA(cudaMemsetAsync(tp.dh0, 0, sizeof(float) * NH, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(tp.dc0, 0, sizeof(float) * NH, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(tp.dh1, 0, sizeof(float) * NH, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(tp.dc1, 0, sizeof(float) * NH, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(tp.d_att_e, 0, sizeof(float) * BR * H, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(tp.d_p_att, 0, sizeof(float) * BR * A, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(G.alpha_w, 0, sizeof(float) * A, st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(G.alpha_b, 0, sizeof(float), st));
    CAPB_CHECK_CUDA(cudaMemsetAsync(G.embed, 0, sizeof(float) * (size_t)V1 * E, st));
    for (int t = T - 1; t >= 0; --t) {
        const float* c0p = t ? tp.c0 + (long)(t - 1) * NH : nullptr;
        const float* c1p = t ? tp.c1 + (long)(t - 1) * NH : nullptr;
        float* dg1 = tp.DG1 + (long)t * N * 4 * H;
        float* dg2 = tp.DG2 + (long)t * N * 4 * H;
        // language LSTM: dh = carried dh1 + dropout-masked dOUT[:, t]
        if (lstm_cell_backward_launch(N, H, tp.g2 + (long)t * N * 4 * H, c1p, tp.c1 + (long)t * NH, tp.dh1, tp.dOUT + (long)t * H, (long)T * H, 3, (unsigned)t,
                                      seed, p, tp.dc1, dg2, st)) return 1;
        if (sk.dgrad(N, 2 * H, 4 * H, dg2, 4 * H, w.lang_lstm_w_ih, 2 * H, tp.dX2, 2 * H, 0)) return 1;   // [d att_res | d h_att]
        if (sk.dgrad(N, H, 4 * H, dg2, 4 * H, w.lang_lstm_w_hh, H, tp.dh1, H, 0)) return 1;              // carried dh_lang
        // attention: needs a contiguous d att_res
        CAPB_CHECK_CUDA(cudaMemcpy2DAsync(tp.tmpH, sizeof(float) * H, tp.dX2, sizeof(float) * 2 * H, sizeof(float) * H, N, cudaMemcpyDeviceToDevice, st));
        float* datth = tp.DATTH + (long)t * N * A;
        if (attention_backward_launch(B, n, R, A, H, tp.tmpH, tp.alpha + (long)t * N * R, tp.atth + (long)t * N * A, tp.p_att, tp.att_e, w.alpha_w, datth,
                                      tp.d_att_e, tp.d_p_att, G.alpha_w, G.alpha_b, tp.dalpha, st)) return 1;
        // dh_att = carried + d h_att from the language LSTM input + d att_h * W_h2att
        if (add_strided_launch(tp.dh0, tp.dX2 + H, 2 * H, N, H, st)) return 1;
        if (sk.dgrad(N, H, A, datth, A, w.h2att_w, H, tp.dh0, H, 1)) return 1;
        if (lstm_cell_backward_launch(N, H, tp.g1 + (long)t * N * 4 * H, c0p, tp.c0 + (long)t * NH, tp.dh0, nullptr, 0, 0, 0, seed, p, tp.dc0, dg1, st)) return 1;
        // inputs of the attention LSTM: [h_lang_prev | fc' | xt] and h_att_prev
        if (sk.dgrad(N, H, 4 * H, dg1, 4 * H, w.att_lstm_w_ih, E + 2 * H, tp.dh1, H, 1)) return 1;         // += d h_lang_prev
        if (sk.dgrad(N, E, 4 * H, dg1, 4 * H, w.att_lstm_w_ih + 2 * H, E + 2 * H, tp.dxt, E, 0)) return 1;
        if (sk.dgrad(N, H, 4 * H, dg1, 4 * H, w.att_lstm_w_hh, H, tp.dh0, H, 0)) return 1;                // carried dh_att
        if (embed_backward_launch(N, E, tp.tok + (long)t * N, tp.xt + (long)t * N * E, tp.dxt, E, keep_scale, G.embed, st)) return 1;
        e->launches += 12;
    }
    // weight gradients, batched over time (K = T*N)
    const int TN1 = (int)((long)(T - 1) * N);
    const float* DG1s = tp.DG1 + (long)N * 4 * H;     // steps 1..T-1 pair with the previous step's hidden states
    const float* DG2s = tp.DG2 + (long)N * 4 * H;
    int rc = 0;
    rc |= sk.wgrad(4 * H, H, (int)TN, tp.DG2, 4 * H, tp.attres, H, G.lang_lstm_w_ih, 2 * H, 0);
    rc |= sk.wgrad(4 * H, H, (int)TN, tp.DG2, 4 * H, tp.h0, H, G.lang_lstm_w_ih + H, 2 * H, 0);
    rc |= sk.wgrad(4 * H, H, TN1, DG2s, 4 * H, tp.h1, H, G.lang_lstm_w_hh, H, 0);
    rc |= colsum_launch((int)TN, 4 * H, tp.DG2, 4 * H, G.lang_lstm_b_ih, 0, st);
    rc |= colsum_launch((int)TN, 4 * H, tp.DG2, 4 * H, G.lang_lstm_b_hh, 0, st);
    rc |= sk.wgrad(4 * H, H, TN1, DG1s, 4 * H, tp.h1, H, G.att_lstm_w_ih, E + 2 * H, 0);
    rc |= sk.wgrad(4 * H, E, (int)TN, tp.DG1, 4 * H, tp.xt, E, G.att_lstm_w_ih + 2 * H, E + 2 * H, 0);
    rc |= sk.wgrad(4 * H, H, TN1, DG1s, 4 * H, tp.h0, H, G.att_lstm_w_hh, H, 0);
    rc |= colsum_launch((int)TN, 4 * H, tp.DG1, 4 * H, G.att_lstm_b_ih, 0, st);
    rc |= colsum_launch((int)TN, 4 * H, tp.DG1, 4 * H, G.att_lstm_b_hh, 0, st);
    rc |= per_image_sum_launch(T, N, n, 4 * H, tp.DG1, tp.S, st);
    rc |= sk.wgrad(4 * H, H, B, tp.S, 4 * H, tp.fc_e, H, G.att_lstm_w_ih + H, E + 2 * H, 0);               // fc' block
    rc |= sk.dgrad(B, H, 4 * H, tp.S, 4 * H, w.att_lstm_w_ih + H, E + 2 * H, tp.d_fc_e, H, 0);             // d fc'
    rc |= sk.wgrad(A, H, (int)TN, tp.DATTH, A, tp.h0, H, G.h2att_w, H, 0);
    rc |= colsum_launch((int)TN, A, tp.DATTH, A, G.h2att_b, 0, st);
    // prologue
    rc |= sk.dgrad((int)BR, H, A, tp.d_p_att, A, w.ctx2att_w, H, tp.d_att_e, H, 1);
    rc |= sk.wgrad(A, H, (int)BR, tp.d_p_att, A, tp.att_e, H, G.ctx2att_w, H, 0);
    rc |= colsum_launch((int)BR, A, tp.d_p_att, A, G.ctx2att_b, 0, st);
    rc |= relu_dropout_backward_launch(BR * H, tp.att_e, tp.d_att_e, tp.dpre_att, keep_scale, st);
    rc |= sk.wgrad(H, Fa, (int)BR, tp.dpre_att, H, att, Fa, G.att_embed_w, Fa, 0);
    rc |= colsum_launch((int)BR, H, tp.dpre_att, H, G.att_embed_b, 0, st);
    rc |= relu_dropout_backward_launch((long)B * H, tp.fc_e, tp.d_fc_e, tp.dpre_fc, keep_scale, st);
    rc |= sk.wgrad(H, Ff, B, tp.dpre_fc, H, fc, Ff, G.fc_embed_w, Ff, 0);
    rc |= colsum_launch(B, H, tp.dpre_fc, H, G.fc_embed_b, 0, st);
    e->launches += 30 + (tf32_context_launches(e->tf32) - tf32_l0);     // + transposes of the tcgen05 path
    if (!rc && record_group_event(e->grad_events[1], st)) return 1;
    return rc;
}

}  // namespace

extern "C" int capb200_updown_scst_step(capb200_engine* e, const float* fc, const float* att, int B, int R, const capb200_scst_opts* opts,
                                        const capb200_cider_table* table, const int* refs, const int* ref_offsets, int L,
                                        const capb200_updown_grads* grads, long long* sample_seq, long long* greedy_seq, float* sample_logprobs,
                                        float* reward, float* loss, void* stream) {
    if (check_ready(e)) return 1;
    CAPB_REQUIRE(e->cfg.family == CAPB200_FAMILY_UPDOWN, "the SCST step is implemented for the UpDown family");
    CAPB_REQUIRE(opts && fc && att && table && refs && ref_offsets && grads && sample_seq && sample_logprobs && reward && loss, "null argument");
    const bool greedy_baseline = opts->baseline == CAPB200_BASELINE_GREEDY;
    CAPB_REQUIRE(greedy_baseline || opts->baseline == CAPB200_BASELINE_LEAVE_ONE_OUT, "unknown baseline");
    CAPB_REQUIRE(!greedy_baseline || greedy_seq != nullptr, "the greedy baseline needs greedy_seq");
    CAPB_REQUIRE(greedy_baseline || opts->sample_n >= 2, "the leave-one-out baseline needs sample_n >= 2");
    CAPB_REQUIRE(opts->sample_n >= 1 && opts->sample_n <= 16 && B >= 1 && R >= 1, "sample_n must be in 1..16");
    CAPB_REQUIRE(opts->drop_prob >= 0.f && opts->drop_prob < 1.f, "drop_prob must be in [0, 1)");
    TrainArgs ta;
    ta.n = opts->sample_n; ta.T = e->T; ta.Tl = e->T; ta.p = opts->drop_prob; ta.temperature = opts->temperature; ta.upstream = opts->upstream;
    ta.seed = opts->seed; ta.greedy_baseline = greedy_baseline; ta.table = table; ta.refs = refs; ta.ref_offsets = ref_offsets; ta.L = L;
    ta.sample_seq = sample_seq; ta.greedy_seq = greedy_seq; ta.reward = reward; ta.logprobs = sample_logprobs; ta.loss = loss;
    ta.forced = opts->forced_tokens; ta.mask = opts->att_masks; ta.keep = opts->keep_rows; ta.row_loss = opts->row_loss;
    CAPB_REQUIRE(ta.keep >= 0 && ta.keep <= B * opts->sample_n, "keep_rows must be in 0..rows");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the whole step as one CUDA graph (see capb200_aoa_scst_step and engine_common.cuh: StepGraph)
    if (!StepGraph::enabled() || !e->tc || ta.forced != nullptr || e->sg.broken) {
        if (dropout_salt_set_all(0ull, st)) return 1;      // eager step: the seed arguments are the effective seeds
        return updown_train_step(e, fc, att, B, R, ta, grads, st);
    }
    cudaStream_t gst = e->sg.enter(st);             // a capturable engine-owned stream, ordered after the caller's stream
    const void* srcs[3] = {fc, att, ta.mask};
    const size_t bytes[3] = {sizeof(float) * (size_t)B * e->cfg.fc_feat_size, sizeof(float) * (size_t)B * R * e->cfg.att_feat_size,
                             ta.mask ? sizeof(float) * (size_t)B * R : 0};
    size_t off[3];
    if (e->sg.stage_inputs(3, srcs, bytes, off, gst)) return 1;
    const float* fc_s = reinterpret_cast<const float*>(e->sg.stage + off[0]);
    const float* att_s = reinterpret_cast<const float*>(e->sg.stage + off[1]);
    if (ta.mask) ta.mask = reinterpret_cast<const float*>(e->sg.stage + off[2]);
    unsigned long long key = 1469598103934665603ull;
    capb200_scst_opts o2 = *opts; o2.seed = 0; o2.att_masks = ta.mask;
    StepGraph::mix(key, &o2, sizeof(o2)); StepGraph::mix(key, grads, sizeof(*grads)); StepGraph::mix(key, &e->w, sizeof(e->w));
    const void* ptrs[] = {table, refs, ref_offsets, sample_seq, greedy_seq, sample_logprobs, reward, loss, e->tape, e->ws, e->wblock, e->sg.stage, gst};
    StepGraph::mix(key, ptrs, sizeof(ptrs));
    StepGraph::mix(key, e->grad_events, sizeof(e->grad_events));
    const int dims[] = {B, R, L};
    StepGraph::mix(key, dims, sizeof(dims));
    const int rc_graph = run_step_graph(e->sg, key, opts->seed, &e->launches, gst, [&]() { return updown_train_step(e, fc_s, att_s, B, R, ta, grads, gst); });
    if (e->sg.leave(st, gst)) return 1;
    return rc_graph;
}

extern "C" int capb200_engine_set_grad_events(capb200_engine* e, void* const* events, int n) {
    CAPB_REQUIRE(e != nullptr && n >= 0 && n <= 2, "UpDown has 2 gradient groups");
    for (int i = 0; i < 2; ++i) e->grad_events[i] = (events != nullptr && i < n) ? static_cast<cudaEvent_t>(events[i]) : nullptr;
    return 0;
}

extern "C" int capb200_updown_xe_step(capb200_engine* e, const float* fc, const float* att, int B, int R, const capb200_xe_opts* opts,
                                      const long long* labels, const float* masks, int label_cols, const capb200_updown_grads* grads, float* logprobs,
                                      float* loss, void* stream) {
    if (check_ready(e)) return 1;
    CAPB_REQUIRE(e->cfg.family == CAPB200_FAMILY_UPDOWN, "the XE step is implemented for the UpDown family");
    CAPB_REQUIRE(opts && fc && att && labels && masks && grads && logprobs && loss, "null argument");
    CAPB_REQUIRE(opts->seq_per_img >= 1 && opts->seq_per_img <= 16 && B >= 1 && R >= 1, "seq_per_img must be in 1..16");
    CAPB_REQUIRE(opts->drop_prob >= 0.f && opts->drop_prob < 1.f, "drop_prob must be in [0, 1)");
    CAPB_REQUIRE(opts->label_smoothing >= 0.f && opts->label_smoothing < 1.f, "label_smoothing must be in [0, 1)");
    CAPB_REQUIRE(label_cols >= 2 && label_cols <= e->T + 2, "labels are [N, seq_length + 2] (BOS, words, EOS padding)");
    CAPB_REQUIRE(opts->steps >= 1 && opts->steps <= label_cols - 1, "steps must be in 1..label_cols-1");
    TrainArgs ta;
    ta.xe = true;
    ta.n = opts->seq_per_img; ta.T = opts->steps; ta.Tl = label_cols - 1; ta.p = opts->drop_prob; ta.upstream = opts->upstream; ta.seed = opts->seed;
    ta.smoothing = opts->label_smoothing;
    ta.labels = labels; ta.ld_labels = label_cols; ta.masks = masks; ta.ld_masks = label_cols; ta.logprobs = logprobs; ta.loss = loss;
    ta.mask = opts->att_masks; ta.ss_prob = opts->ss_prob; ta.tokens_used = opts->tokens_used; ta.keep = opts->keep_rows; ta.row_loss = opts->row_loss;
    CAPB_REQUIRE(ta.ss_prob >= 0.f && ta.ss_prob <= 1.f, "ss_prob must be in [0, 1]");
    CAPB_REQUIRE(ta.keep >= 0 && ta.keep <= B * opts->seq_per_img, "keep_rows must be in 0..rows");
    if (dropout_salt_set_all(0ull, static_cast<cudaStream_t>(stream))) return 1;      // eager step: the seed arguments are the effective seeds
    return updown_train_step(e, fc, att, B, R, ta, grads, static_cast<cudaStream_t>(stream));
}


capb200_cider_table* capb200_cider_table_create(const int* keys, const double* df, long n, double ref_len, void* stream) {
    if (keys == nullptr || df == nullptr || n < 0 || ref_len <= 0) { set_error("bad CIDEr-D table arguments"); return nullptr; }
    CiderTable* t = cider_table_create(keys, df, n, ref_len, static_cast<cudaStream_t>(stream));
    if (t == nullptr) return nullptr;
    capb200_cider_table* h = new capb200_cider_table();
    h->t = t;
    return h;
}

void capb200_cider_table_destroy(capb200_cider_table* t) {
    if (t == nullptr) return;
    cider_table_destroy(t->t);
    delete t;
}

int capb200_self_critical_reward(const capb200_cider_table* t, const long long* sampled, int S, const long long* greedy, int B, int T,
                                 const int* refs, const int* ref_offsets, int L, double* scores, float* reward, void* stream) {
    CAPB_REQUIRE(t != nullptr && sampled && greedy && refs && ref_offsets && scores, "null argument");
    return cider_reward_launch(t->t, sampled, S, greedy, B, T, refs, ref_offsets, L, scores, reward, T, T, static_cast<cudaStream_t>(stream));
}

int capb200_cider_scores(const capb200_cider_table* t, const long long* sampled, int S, int B, int T, const int* refs, const int* ref_offsets,
                         int L, double* scores, float* reward, void* stream) {
    CAPB_REQUIRE(t != nullptr && sampled && refs && ref_offsets && scores, "null argument");
    return cider_reward_launch(t->t, sampled, S, nullptr, B, T, refs, ref_offsets, L, scores, reward, T, T, static_cast<cudaStream_t>(stream));
}

int capb200_reward_criterion_forward(const float* logprobs, const long long* seq, const float* reward, int N, int T, int V1, float* loss_mean,
                                     float* loss_rows, float* mask_sum, void* stream) {
    CAPB_REQUIRE(logprobs && seq && reward && N > 0 && T > 0, "bad argument");
    return reward_criterion_fwd_launch(logprobs, (long)T * V1, V1, seq, reward, N, T, loss_mean, loss_rows, mask_sum, static_cast<cudaStream_t>(stream));
}

int capb200_reward_criterion_backward(const long long* seq, const float* reward, int N, int T, int V1, const float* mask_sum, float upstream,
                                      float* grad, void* stream) {
    CAPB_REQUIRE(seq && reward && mask_sum && grad, "null argument");
    return reward_criterion_bwd_launch(seq, reward, N, T, mask_sum, upstream, grad, (long)T * V1, V1, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
