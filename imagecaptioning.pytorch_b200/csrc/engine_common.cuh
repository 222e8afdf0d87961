// Helpers shared by the engines (UpDown/NewFC in engine.cu, Transformer in tfm_engine.cu, AoA in aoa_engine.cu).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/capb200.h"
#include "common.cuh"
#include "kernels.cuh"
#include "nvtx.cuh"

// opaque C handle of the CIDEr-D document-frequency table (shared by the UpDown and AoA training steps)
struct capb200_cider_table {
    capb200::CiderTable* t = nullptr;
};

namespace capb200 {

struct Planes {
    __half* hi = nullptr;
    __half* lo = nullptr;
    long ld = 0;
};

// bump allocator over one cudaMalloc'ed block; a dry run (base == nullptr) measures the size
struct Arena {
    char* base = nullptr;
    size_t off = 0;
    template <typename T>
    T* take(size_t n) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

struct Act {
    ActView v;
    void carve(Arena& a, long rows, long cols, bool planes) {
        v.ld = round_up(cols, 64);          // 128-byte plane rows: every TMA box row is one aligned L2 line
        v.f = a.take<float>(rows * v.ld);
        if (planes) {
            v.hi = a.take<__half>(rows * v.ld);
            v.lo = a.take<__half>(rows * v.ld);
        } else {
            v.hi = v.lo = nullptr;
        }
    }
};

inline DecodeEdits to_edits(const capb200_decode_edits& c) {
    DecodeEdits e;
    e.constraint = c.decoding_constraint;
    e.unk_col = c.unk_col;
    e.n_bad = (c.bad_endings != nullptr && c.n_bad_endings > 0) ? c.n_bad_endings : 0;
    e.bad = c.bad_endings;
    e.trigrams = c.block_trigrams;
    e.trigram_rows = c.trigram_rows;
    return e;
}

inline Planes carve_planes(Arena& a, long rows, long cols) {
    Planes p;
    p.ld = round_up(cols, 64);
    p.hi = a.take<__half>(rows * p.ld);
    p.lo = a.take<__half>(rows * p.ld);
    return p;
}


inline GemmSeg seg_of(const ActView& a, const float* w, long ldw, const Planes& wp, int K) {
    GemmSeg s;
    s.A = a.f; s.lda = a.ld; s.W = w; s.ldw = ldw;
    s.A_hi = a.hi; s.A_lo = a.lo; s.lda_h = a.ld;
    s.W_hi = wp.hi; s.W_lo = wp.lo; s.ldw_h = wp.ld;
    s.K = K;
    return s;
}


// Runs one GEMM in the engine's numeric mode.  `plan` caches the encoded TMA maps of this call site (tensor-core modes);
// `plan_rows` is the row capacity the maps are encoded for, g.M the rows valid in this launch.
inline int run_gemm_mode(int mode, GemmTcPlan** plan, GemmProblem& g, int plan_rows, cudaStream_t st) {
    if (mode == 0) return gemm_simt_launch(g, st);
    if (*plan == nullptr) {
        GemmProblem planned = g;
        planned.M = plan_rows;
        *plan = gemm_tc_plan_create(planned, mode == 1 ? 3 : 1);
        if (*plan == nullptr) return 1;
    }
    return gemm_tc_plan_launch(*plan, &g.epi, g.M, st);
}


// ---------------------------------------------------------------------------------------------------------------------
// Search / sampling state shared by every model family, and the two decode drivers.  A family supplies only its recurrent
// core as a callable:  core(rows, rows_per_image, tokens, src_row, t, logits, ld_logits) -> 0 on success, which must leave the
// step's raw logits [rows, V1] at `logits` (row pitch ld_logits).
// ---------------------------------------------------------------------------------------------------------------------
struct DecodeBuffers {
    int *tokens = nullptr, *src_row = nullptr, *neg1 = nullptr, *unfinished = nullptr, *forced = nullptr;
    float* top_val = nullptr;
    int* top_idx = nullptr;
    float* top_val_e = nullptr;       // [rows, 16] candidate lists after the decode edits (beam search with options)
    int* top_idx_e = nullptr;
    float2* slab_stats = nullptr;     // [T, rows]
    BeamState bs;
    long long* rec_seq = nullptr;     // [B, beam, T] sorted records of the last beam decode
    int *rec_len = nullptr, *rec_hist = nullptr, *out_hist = nullptr, *tmp_len = nullptr;
    float *rec_p = nullptr, *rec_raw = nullptr, *tmp_p = nullptr, *tmp_raw = nullptr;
    float* slab = nullptr;            // separately allocated: [T, rows, V1] raw logits of a beam search
    size_t slab_bytes = 0;
    long slab_step_stride = 0;
    int last_B = 0, last_beam = 0;
    DecodeEdits last_edits;           // edits of the last beam decode (re-applied when a finished beam's rows are materialised later)
    // CUDA graph of the T-step beam loop (every launch of it is static for a given shape / workspace): captured the second time a
    // configuration is seen, replayed afterwards; the launch gaps of ~180 serial kernels are ~5 % of a decode
    cudaGraphExec_t loop_exec = nullptr;
    unsigned long long loop_key[8] = {0, 0, 0, 0, 0, 0, 0, 0}, seen_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long loop_launches = 0;
    bool graph_broken = false;

    void carve(Arena& a, int B, int rows, int beam, int T) {
        tokens = a.take<int>(rows);
        src_row = a.take<int>(rows);
        neg1 = a.take<int>(rows);
        unfinished = a.take<int>(rows);
        forced = a.take<int>(rows);
        top_val = a.take<float>((long)rows * 16);
        top_idx = a.take<int>((long)rows * 16);
        top_val_e = a.take<float>((long)rows * 16);
        top_idx_e = a.take<int>((long)rows * 16);
        slab_stats = a.take<float2>((long)rows * T);
        const long rec = (long)B * beam * T;
        bs.sums = a.take<float>((long)B * beam);
        bs.seq_a = a.take<int>(rec);
        bs.seq_b = a.take<int>(rec);
        bs.hist_a = a.take<int>(rec);
        bs.hist_b = a.take<int>(rec);
        bs.done_cnt = a.take<int>(B);
        bs.done_seq = a.take<int>(rec * T);
        bs.done_hist = a.take<int>(rec * T);
        bs.done_len = a.take<int>(rec);
        bs.done_p = a.take<double>(rec);
        bs.done_raw = a.take<float>(rec);
        bs.tokens = tokens;
        bs.src_row = src_row;
        rec_seq = a.take<long long>(rec);
        rec_hist = a.take<int>(rec);
        out_hist = a.take<int>(rec);
        rec_len = a.take<int>((long)B * beam);
        rec_p = a.take<float>((long)B * beam);
        rec_raw = a.take<float>((long)B * beam);
        tmp_len = a.take<int>((long)B * beam);
        tmp_p = a.take<float>((long)B * beam);
        tmp_raw = a.take<float>((long)B * beam);
    }
};

__global__ void capb_fill_int_kernel(int* p, int n, int v);
__global__ void capb_load_token_column_kernel(const long long* src, long ld, int col, int n, int* dst);
int fill_int_launch(int* p, int n, int v, cudaStream_t st);
int load_token_column_launch(const long long* src, long ld, int col, int n, int* dst, cudaStream_t st);
int store_token_column_launch(const int* src, int n, long long* dst, long ld, int col, cudaStream_t st);

// ancestors of the current rows at step t of a beam search (valid for positions < t): the table beam_step(t-1) wrote
inline const int* beam_ancestors(const BeamState& s, int t) { return ((t - 1) & 1) ? s.hist_a : s.hist_b; }

// AttModel._sample_beam + CaptionModel.beam_search (see engine.cu header for the reference lines)
// Key of everything outside the driver that the captured beam-loop launches depend on (0 disables graph capture)
inline unsigned long long loop_graph_key(const void* ws, const void* wblock, const void* mask, int R, int family) {
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { h ^= v; h *= 1099511628211ull; };
    mix(reinterpret_cast<uintptr_t>(ws)); mix(reinterpret_cast<uintptr_t>(wblock)); mix(reinterpret_cast<uintptr_t>(mask));
    mix((unsigned long long)R); mix((unsigned long long)family + 17);
    return h | 1ull;
}

template <class CoreFn>
int beam_decode_driver(DecodeBuffers& d, int V1, int T, int B, int beam, int keep, int penalty_kind, float penalty_alpha, long long* seq,
                       float* seq_logprobs, long long* done_seq, int* done_len, float* done_p, float* done_raw, CoreFn core, long* launches,
                       cudaStream_t st, unsigned long long graph_key = 0, const DecodeEdits& ed = DecodeEdits(), float temperature = 1.0f) {
    const int rows = B * beam;
    const bool edits = ed.any();
    const int k_in = beam + ed.kinds();
    CAPB_REQUIRE(!ed.trigrams, "block_trigrams applies to _sample only (AttModel.py:306)");
    CAPB_REQUIRE(k_in <= 16, "beam_size + number of active decode edits (decoding_constraint, remove_bad_endings, UNK suppression) must be <= 16");
    if (temperature == 0.f) temperature = 1.0f;
    CAPB_REQUIRE(temperature > 0.f, "temperature must be positive");
    const size_t slab_need = (size_t)T * rows * V1 * sizeof(float);
    if (slab_need > d.slab_bytes) {
        CAPB_CHECK_CUDA(cudaStreamSynchronize(st));
        if (d.slab) CAPB_CHECK_CUDA(cudaFree(d.slab));
        d.slab = nullptr;
        CAPB_CHECK_CUDA(cudaMalloc(&d.slab, slab_need));
        d.slab_bytes = slab_need;
    }
    d.slab_step_stride = (long)rows * V1;
    d.last_B = B;
    d.last_beam = beam;
    BeamState s = d.bs;
    s.B = B; s.beam = beam; s.T = T; s.V1 = V1;
    auto run_loop = [&]() -> int {
        CAPB_NVTX("capb200 beam loop (T steps: core, vocab stats, beam step)");
        CAPB_CHECK_CUDA(cudaMemsetAsync(s.sums, 0, sizeof(float) * B * beam, st));
        CAPB_CHECK_CUDA(cudaMemsetAsync(s.done_cnt, 0, sizeof(int) * B, st));
        CAPB_CHECK_CUDA(cudaMemsetAsync(d.tokens, 0, sizeof(int) * rows, st));      // <bos> = 0
        for (int t = 0; t < T; ++t) {
            const int live = (t == 0) ? 1 : beam;
            const int nrows = B * live;
            float* logits = d.slab + (long)t * d.slab_step_stride;
            if (core(nrows, live, d.tokens, t == 0 ? d.neg1 : d.src_row, t, logits, (long)V1)) return 1;
            if (t > 0 && temperature != 1.0f) {      // log_softmax(logprobs / temperature) = log_softmax(logits / temperature)  (CaptionModel.py:204)
                if (scale_rows_launch(logits, V1, nrows, V1, 1.0f / temperature, st)) return 1;
                *launches += 1;
            }
            VocabStepArgs va;
            va.rows = nrows; va.V1 = V1; va.logits = logits; va.ld = V1;
            va.twice = (t > 0) ? 1 : 0;      // init_logprobs went through one log_softmax only (AttModel.py:239, CaptionModel.py:204)
            va.topk = edits ? k_in : beam; va.top_val = d.top_val; va.top_idx = d.top_idx;
            va.stats = d.slab_stats + (long)t * rows;
            if (vocab_step_launch(va, st)) return 1;
            const float* tv = d.top_val;
            const int* ti = d.top_idx;
            if (edits) {      // drop / lower the edited candidates, keep the `beam` best (the edits of CaptionModel.py:154-162)
                if (beam_edit_launch(nrows, k_in, beam, t, ed, d.tokens, d.top_val, d.top_idx, d.top_val_e, d.top_idx_e, st)) return 1;
                tv = d.top_val_e; ti = d.top_idx_e;
                *launches += 1;
            }
            if (beam_step_launch(s, t, live, tv, ti, penalty_kind, penalty_alpha, st)) return 1;
            *launches += 2;
        }
        return 0;
    };
    // graph key: everything the captured launches depend on (caller's key covers workspace / weight / mask pointers and R)
    unsigned long long key[8] = {graph_key, (unsigned long long)B, (unsigned long long)beam, (unsigned long long)T, (unsigned long long)V1,
                                 (unsigned long long)penalty_kind, 0ull, (unsigned long long)reinterpret_cast<uintptr_t>(d.slab)};
    memcpy(&key[6], &penalty_alpha, sizeof(float));
    memcpy(reinterpret_cast<char*>(&key[6]) + 4, &temperature, sizeof(float));
    // decode edits are baked into the captured launches too
    key[5] ^= ((unsigned long long)(ed.constraint & 1) << 8) ^ ((unsigned long long)(unsigned)(ed.unk_col + 1) << 16) ^ ((unsigned long long)ed.n_bad << 48);
    key[0] ^= (unsigned long long)reinterpret_cast<uintptr_t>(ed.bad) * 0x9E3779B97F4A7C15ull;
    d.last_edits = ed;
    static const bool graphs_off = getenv("CAPB200_NO_GRAPH") != nullptr;
    const bool try_graph = graph_key != 0 && !graphs_off && !d.graph_broken;
    if (try_graph && d.loop_exec != nullptr && memcmp(key, d.loop_key, sizeof(key)) == 0) {
        CAPB_CHECK_CUDA(cudaGraphLaunch(d.loop_exec, st));
        *launches += d.loop_launches;
    } else if (try_graph && memcmp(key, d.seen_key, sizeof(key)) == 0) {
        // second decode with this configuration: every lazy initialisation has happened, capture the loop and replay it from now on
        if (d.loop_exec != nullptr) { cudaGraphExecDestroy(d.loop_exec); d.loop_exec = nullptr; }
        const long l0 = *launches;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        const int rc = ok ? run_loop() : 1;
        if (ok) ok = cudaStreamEndCapture(st, &graph) == cudaSuccess && graph != nullptr && rc == 0;
        if (ok) ok = cudaGraphInstantiate(&d.loop_exec, graph, 0) == cudaSuccess;
        if (graph != nullptr) cudaGraphDestroy(graph);
        if (!ok) {
            (void)cudaGetLastError();
            d.loop_exec = nullptr;
            d.graph_broken = true;          // fall back to eager launches for good
            *launches = l0;
            if (run_loop()) return 1;
        } else {
            d.loop_launches = *launches - l0;
            memcpy(d.loop_key, key, sizeof(key));
            CAPB_CHECK_CUDA(cudaGraphLaunch(d.loop_exec, st));
        }
    } else {
        memcpy(d.seen_key, key, sizeof(key));
        if (run_loop()) return 1;
    }
    // all finished beams of every image, best first
    CAPB_NVTX("capb200 beam finalize + log-prob rows");
    if (beam_finalize_launch(s, beam, d.rec_seq, d.rec_len, d.rec_p, d.rec_raw, d.rec_hist, st)) return 1;
    *launches += 1;
    if (keep == beam) {
        CAPB_CHECK_CUDA(cudaMemcpyAsync(seq, d.rec_seq, sizeof(long long) * B * beam * T, cudaMemcpyDeviceToDevice, st));
        if (seq_logprobs) {
            *launches += 1;
            if (gather_logprob_rows_launch(d.slab, d.slab_step_stride, V1, d.rec_hist, B * beam, T, V1, seq_logprobs, d.slab_stats, rows, st, d.rec_seq, &ed)) return 1;
        }
    } else {
        *launches += 1;
        if (beam_finalize_launch(s, 1, seq, d.tmp_len, d.tmp_p, d.tmp_raw, d.out_hist, st)) return 1;
        if (seq_logprobs) {
            *launches += 1;
            if (gather_logprob_rows_launch(d.slab, d.slab_step_stride, V1, d.out_hist, B, T, V1, seq_logprobs, d.slab_stats, rows, st, seq, &ed)) return 1;
        }
    }
    if (done_seq) CAPB_CHECK_CUDA(cudaMemcpyAsync(done_seq, d.rec_seq, sizeof(long long) * B * beam * T, cudaMemcpyDeviceToDevice, st));
    if (done_len) CAPB_CHECK_CUDA(cudaMemcpyAsync(done_len, d.rec_len, sizeof(int) * B * beam, cudaMemcpyDeviceToDevice, st));
    if (done_p) CAPB_CHECK_CUDA(cudaMemcpyAsync(done_p, d.rec_p, sizeof(float) * B * beam, cudaMemcpyDeviceToDevice, st));
    if (done_raw) CAPB_CHECK_CUDA(cudaMemcpyAsync(done_raw, d.rec_raw, sizeof(float) * B * beam, cudaMemcpyDeviceToDevice, st));
    return 0;
}

inline int beam_record_logprobs(DecodeBuffers& d, int V1, int T, int image, int rank, float* dst, cudaStream_t st) {
    CAPB_REQUIRE(d.slab != nullptr && image >= 0 && image < d.last_B && rank >= 0 && rank < d.last_beam, "no such finished beam");
    return gather_logprob_rows_launch(d.slab, d.slab_step_stride, V1, d.rec_hist + ((long)image * d.last_beam + rank) * T, 1, T, V1, dst,
                                      d.slab_stats, (long)d.last_B * d.last_beam, st, d.rec_seq + ((long)image * d.last_beam + rank) * T, &d.last_edits);
}

// AttModel._sample (greedy / multinomial / forced replay) and AttModel._forward (teacher forcing); method codes = CAPB200_SAMPLE_*
template <class CoreFn>
int sample_decode_driver(DecodeBuffers& d, int V1, int T, int rows, int method, float temperature, unsigned long long seed, int steps,
                         const long long* tokens_in, long ld_tok, long long* seq, float* seq_logprobs, float* picked, CoreFn core, long* launches,
                         cudaStream_t st, const DecodeEdits& ed = DecodeEdits(), float top = 0.f) {
    const bool teacher = method == 3, forced = method == 2;
    CAPB_NVTX("capb200 sample / teacher-forcing loop");
    CAPB_REQUIRE(ed.unk_col < 0, "UNK suppression is a beam-search option (CaptionModel.py:159-162)");
    if (method == 4) CAPB_REQUIRE(top >= 1.f, "top-k sampling needs k >= 1");
    if (method == 5) CAPB_REQUIRE(top > 0.f && top < 1.f, "nucleus sampling needs 0 < p < 1");
    const long t_out = teacher ? ld_tok : T;
    CAPB_CHECK_CUDA(cudaMemsetAsync(d.tokens, 0, sizeof(int) * rows, st));
    for (int t = 0; t < steps; ++t) {
        if (teacher) { if (load_token_column_launch(tokens_in, ld_tok, t, rows, d.tokens, st)) return 1; *launches += 1; }
        else if (forced) { if (load_token_column_launch(tokens_in, ld_tok, t, rows, d.forced, st)) return 1; *launches += 1; }
        float* logits = seq_logprobs + (long)t * V1;
        if (core(rows, 0, d.tokens, t == 0 ? d.neg1 : nullptr, t, logits, t_out * V1)) return 1;
        VocabStepArgs va;
        va.rows = rows; va.V1 = V1; va.logits = logits; va.ld = t_out * V1;
        va.twice = 0;
        if (!teacher) {
            va.select = (method == 0) ? 1 : (method == 1 ? 2 : (method == 2 ? 3 : method));      // 4 top-k, 5 nucleus
            va.top = top;
            va.edits = ed;
            va.prev_tokens = d.tokens;        // the word fed into this step (read before tokens_out is rewritten at the end of the kernel)
            va.temperature = temperature;
            va.seed = seed;
            va.step = (unsigned long long)t;
            va.forced = d.forced;
            va.unfinished = d.unfinished;
            va.first_step = (t == 0);
            va.tokens_out = d.tokens;
            va.seq_out = seq; va.ld_seq = T; va.t = t;
            va.picked_lp = picked ? picked + t : nullptr;      // picked is [N,T]
            va.ld_picked = T;
        }
        *launches += 1;
        if (vocab_step_launch(va, st)) return 1;
    }
    return 0;
}

// Training-step GEMMs on the raw fp32 PyTorch weights (always current, no repack after optimizer steps).  With a Tf32Context (tensor-core
// engines) every call runs on the tcgen05 kind::tf32 3-pass kernel of gemm_tf32.cu; operands that are not K-major in HBM (W for the input
// gradients, dY / X for the weight gradients) go through cached transposes.  Without a context (simt_fp32 engines), when an operand is not
// TMA-compatible (rows not 16-byte aligned: tiny test shapes), or with CAPB200_SKINNY_LEGACY set, the split-K kernels of gemm_generic.cu run.
// CUDA graph of a whole fused training step (see capb200_aoa_scst_step) + the engine-owned staging buffer that gives the graph stable input
// addresses.  CAPB200_SCST_GRAPH=0 keeps the steps eager.
struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    unsigned long long key = 0, seen = 0, cap_seed = 0;
    long launches = 0, replays = 0;
    char* stage = nullptr;
    size_t stage_bytes = 0;
    bool broken = false;
    // The caller's stream is usually torch's legacy default stream, which cannot be captured: the step runs on an engine-owned stream that
    // waits for the caller's stream on entry (enter) and that the caller's stream waits for on exit (leave).
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    cudaStream_t enter(cudaStream_t caller) {
        if (stream == nullptr) {
            if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); broken = true; return caller; }
        }
        if (cudaEventRecord(ev_in, caller) != cudaSuccess || cudaStreamWaitEvent(stream, ev_in, 0) != cudaSuccess) { (void)cudaGetLastError(); broken = true; return caller; }
        return stream;
    }
    int leave(cudaStream_t caller, cudaStream_t used) {
        if (used == caller) return 0;
        CAPB_CHECK_CUDA(cudaEventRecord(ev_out, used));
        CAPB_CHECK_CUDA(cudaStreamWaitEvent(caller, ev_out, 0));
        return 0;
    }
    static bool enabled() { static const bool v = !(getenv("CAPB200_SCST_GRAPH") != nullptr && atoi(getenv("CAPB200_SCST_GRAPH")) == 0); return v; }
    void reset() { if (exec) cudaGraphExecDestroy(exec); exec = nullptr; key = 0; }
    void destroy() {
        if (getenv("CAPB200_GRAPH_DEBUG") != nullptr && (exec || replays)) fprintf(stderr, "capb200: step graph replayed %ld times\n", replays);
        reset(); if (stage) cudaFree(stage); stage = nullptr; stage_bytes = 0;
        if (ev_in) cudaEventDestroy(ev_in);
        if (ev_out) cudaEventDestroy(ev_out);
        if (stream) cudaStreamDestroy(stream);
        ev_in = ev_out = nullptr; stream = nullptr;
    }
    // copies up to four buffers back to back (256-byte aligned) into the staging buffer in stream order; off[i] = where buffer i landed
    int stage_inputs(int n, const void* const* src, const size_t* bytes, size_t* off, cudaStream_t st) {
        size_t need = 256;
        for (int i = 0; i < n; ++i) { off[i] = need; need += (bytes[i] + 255) & ~size_t(255); }
        if (need > stage_bytes) {
            CAPB_CHECK_CUDA(cudaStreamSynchronize(st));
            reset();
            if (stage) CAPB_CHECK_CUDA(cudaFree(stage));
            stage = nullptr;
            CAPB_CHECK_CUDA(cudaMalloc(&stage, need));
            stage_bytes = need;
        }
        for (int i = 0; i < n; ++i)
            if (src[i] != nullptr && bytes[i]) CAPB_CHECK_CUDA(cudaMemcpyAsync(stage + off[i], src[i], bytes[i], cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    static void mix(unsigned long long& h, const void* p, size_t nbytes) {
        const unsigned char* c = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < nbytes; ++i) { h ^= c[i]; h *= 1099511628211ull; }
    }
};

// Runs `run()` (which enqueues one whole training step on `st`, reading its inputs from the staging buffer) eagerly the first time `key` is
// seen, captures it into a graph the second time, and replays the graph afterwards with the seed carried by the salt.
template <class Run>
int run_step_graph(StepGraph& sg, unsigned long long key, unsigned long long seed, long* launches, cudaStream_t st, Run run) {
    if (sg.exec != nullptr && sg.key == key) {
        sg.replays++;
        if (dropout_salt_set_all(sg.cap_seed ^ seed, st)) return 1;
        CAPB_CHECK_CUDA(cudaGraphLaunch(sg.exec, st));
        *launches += sg.launches;
        return 0;
    }
    if (dropout_salt_set_all(0ull, st)) return 1;
    if (sg.seen != key) {               // first sighting: eager (it also performs every first-use allocation)
        sg.seen = key;
        return run();
    }
    sg.reset();
    const long l0 = *launches;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { (void)cudaGetLastError(); sg.broken = true; return run(); }
    const int rc = run();
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != 0 || ce != cudaSuccess || graph == nullptr) {
        (void)cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        sg.broken = true;               // something in the step is not capturable here: stay eager from now on
        *launches = l0;
        return run();
    }
    const cudaError_t ie = cudaGraphInstantiate(&sg.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { (void)cudaGetLastError(); sg.exec = nullptr; sg.broken = true; *launches = l0; return run(); }
    sg.key = key; sg.cap_seed = seed; sg.launches = *launches - l0;
    if (getenv("CAPB200_GRAPH_DEBUG") != nullptr) fprintf(stderr, "capb200: training step captured into a CUDA graph (%ld launches)\n", sg.launches);
    CAPB_CHECK_CUDA(cudaGraphLaunch(sg.exec, st));
    return 0;
}

// Records a caller-owned "gradient group complete" event.  Inside a stream capture (the step is being turned into a graph) the record
// becomes an EXTERNAL event-record node: every replay records the event when the node's dependencies have executed, and a
// cudaStreamWaitEvent issued on another stream after cudaGraphLaunch waits for exactly that (the overlapped all-reduce of grad_sync.py).
inline int record_group_event(cudaEvent_t ev, cudaStream_t st) {
    if (ev == nullptr) return 0;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    CAPB_CHECK_CUDA(cudaStreamIsCapturing(st, &cs));
    CAPB_CHECK_CUDA(cudaEventRecordWithFlags(ev, st, cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault));
    return 0;
}

// Side stream of the SCST steps (the eval-mode greedy baseline runs on it while the train-mode sampling pass runs on the caller's stream).
// Lowest priority by default: both chains are latency-bound and compete for SMs (a persistent GEMM CTA owns its SM's shared memory), and
// the sampling pass is the critical path -- its pending CTAs should be placed first.  CAPB200_SIDE_PRIORITY=0 gives both equal priority.
inline cudaError_t create_side_stream(cudaStream_t* s) {
    static const bool equal = getenv("CAPB200_SIDE_PRIORITY") != nullptr && atoi(getenv("CAPB200_SIDE_PRIORITY")) == 0;
    int least = 0, greatest = 0;
    if (!equal && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess) return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, least);
    return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
}

struct Skinny {
    float* scratch; size_t cap; int mode; cudaStream_t st;
    Tf32Context* ctx = nullptr;
    static bool legacy() { static const bool v = getenv("CAPB200_SKINNY_LEGACY") != nullptr; return v; }
    bool tc() const { return ctx != nullptr && mode != 0 && !legacy(); }
    // y = x * W^T (+ b)          (nn.Linear forward; W stored [N, K])
    int lin(const float* x, long ldx, const float* w, long ldw, const float* b, float* y, long ldy, int M, int N, int K, int accumulate) const {
        if (tc() && gemm_tf32_supported(1, &x, &ldx, &w, &ldw, &K))
            return gemm_tf32_launch(ctx, M, N, 1, &x, &ldx, &w, &ldw, &K, y, ldy, b, nullptr, 0, 1, accumulate, st);
        const int tb = 1;
        return gemm_skinny_launch(M, N, 1, &x, &ldx, &w, &ldw, &K, &tb, y, ldy, b, nullptr, 0, 1, accumulate, scratch, cap, mode, st);
    }
    // dx = dy * W                (nn.Linear input gradient; W stored [K, N] = [out, in])
    int dgrad(int M, int N, int K, const float* dy, long lddy, const float* w, long ldw, float* dx, long lddx, int accumulate) const {
        if (tc() && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (lddy & 3) == 0) {
            long ldt = 0;
            const float* wt = tf32_transposed(ctx, w, ldw, K, N, true, &ldt, st);          // W^T [N, K]: rebuilt once per training step
            if (wt == nullptr) return 1;
            return gemm_tf32_launch(ctx, M, N, 1, &dy, &lddy, &wt, &ldt, &K, dx, lddx, nullptr, nullptr, 0, 1, accumulate, st);
        }
        const int tb = 0;
        return gemm_skinny_launch(M, N, 1, &dy, &lddy, &w, &ldw, &K, &tb, dx, lddx, nullptr, nullptr, 0, 1, accumulate, scratch, cap, mode, st);
    }
    // the K-segmented gate GEMM of the decode path, on fp32 weights
    int gates(const GemmProblem& g) const {
        const float* A[3]; const float* B[3]; long lda[3], ldb[3]; int K[3], tb[3];
        for (int i = 0; i < g.nseg; ++i) { A[i] = g.seg[i].A; lda[i] = g.seg[i].lda; B[i] = g.seg[i].W; ldb[i] = g.seg[i].ldw; K[i] = g.seg[i].K; tb[i] = 1; }
        if (tc() && gemm_tf32_supported(g.nseg, A, lda, B, ldb, K))
            return gemm_tf32_launch(ctx, g.M, g.N, g.nseg, A, lda, B, ldb, K, g.epi.C, g.epi.ldc, g.epi.bias, g.epi.row_bias, g.epi.ld_row_bias,
                                    g.epi.rows_per_group, 0, st);
        return gemm_skinny_launch(g.M, g.N, g.nseg, A, lda, B, ldb, K, tb, g.epi.C, g.epi.ldc, g.epi.bias, g.epi.row_bias, g.epi.ld_row_bias,
                                  g.epi.rows_per_group, 0, scratch, cap, mode, st);
    }
    // dW[out, in] (+)= dY[rows, out]^T * X[rows, in]      (weight gradient, batched over time: rows = T * N)
    int wgrad(int out_f, int in_f, int rows, const float* dY, long ld_dy, const float* X, long ld_x, float* G, long ld_g, int accumulate) const {
        if (tc() && rows >= 64) {
            long ld_a = 0, ld_b = 0;
            const float* dyt = tf32_transposed(ctx, dY, ld_dy, rows, out_f, false, &ld_a, st);     // [out, rows]
            const float* xt = dyt ? tf32_transposed(ctx, X, ld_x, rows, in_f, false, &ld_b, st) : nullptr;     // [in, rows]
            if (xt == nullptr) return 1;
            return gemm_tf32_launch(ctx, out_f, in_f, 1, &dyt, &ld_a, &xt, &ld_b, &rows, G, ld_g, nullptr, nullptr, 0, 1, accumulate, st);
        }
        return gemm_wgrad_launch(out_f, in_f, rows, dY, ld_dy, X, ld_x, G, ld_g, accumulate, mode, st);
    }
};


}  // namespace capb200
