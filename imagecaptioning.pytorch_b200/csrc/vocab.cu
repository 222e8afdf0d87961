// Vocabulary epilogue: log-softmax (once for _sample, twice for beam search) and candidate selection.
//
// Replaces, per decode step:  F.log_softmax(self.logit(output))            AttModel.py:172
//                             F.log_softmax(logprobs / temperature)        CaptionModel.py:204   (beam search only; T = 1)
//                             torch.sort(b*(V+1) candidates)[:b]           CaptionModel.py:80-81 (per-row top-b, merged in beam.cu)
//                             torch.max / Categorical(logits).sample()     CaptionModel.py:372,405
//                             finished-row masking                         AttModel.py:340-347
// One CTA per row; the row lives in shared memory between the passes so HBM sees one read and one write.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace capb200 {

// seed salt of the sampling kernels (see dropout.cuh: lets a captured CUDA graph of the SCST step draw fresh samples on every replay)
static __device__ unsigned long long g_vocab_seed_salt = 0ull;
int dropout_salt_set_vocab(unsigned long long salt, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(g_vocab_seed_salt, &salt, sizeof(salt), 0, cudaMemcpyHostToDevice, st) == cudaSuccess ? 0 : 1;
}

namespace {

constexpr int VT = 256;   // threads per row

template <int NT = VT>
__device__ __forceinline__ float block_max(float v, float* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = scratch[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) r = fmaxf(r, scratch[w]);
    return r;
}

template <int NT = VT>
__device__ __forceinline__ float block_sum(float v, float* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) r += scratch[w];
    return r;
}

// arg-max with lowest-index tie-break; result broadcast to all threads
template <int NT = VT>
__device__ __forceinline__ void block_argmax(float v, int i, float* sval, int* sidx, float& out_v, int& out_i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sval[threadIdx.x >> 5] = v; sidx[threadIdx.x >> 5] = i; }
    __syncthreads();
    out_v = sval[0];
    out_i = sidx[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) {
        const float ov = sval[w];
        const int oi = sidx[w];
        if (ov > out_v || (ov == out_v && oi < out_i)) { out_v = ov; out_i = oi; }
    }
}

// Philox4x32-10 counter-based generator (Salmon et al. 2011): one 128-bit block per (element, row, step).
__device__ __forceinline__ uint32_t philox_first(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c0;
}

// Passes over the shared-memory copy of the row: (1) load + max, (2) sum exp, (3) log-probs + second-normalisation sum +
// per-thread sorted top-KMAX, (4) final values to HBM; then k block-wide arg-max rounds over the per-thread list heads.
// The second log_softmax only shifts the row by a constant, so candidates are ranked on the first-pass values.
// NT threads per row: 256 in general; 1024 for the few-row sampling / greedy steps of the training loops, where one CTA per row leaves the
// machine nearly empty and the row passes (37 elements per thread at 256 threads, a Philox draw each) are the whole cost
template <int KMAX, int NT = VT>
__global__ void __launch_bounds__(NT) vocab_step_kernel(const VocabStepArgs a) {
    extern __shared__ float row[];                // [V1]
    __shared__ float s_red[NT / 32];
    __shared__ int s_idx[NT / 32];
    const int r = blockIdx.x;
    const int V1 = a.V1;
    float* g = a.logits + (long)r * a.ld;

    if (a.unfinished != nullptr && !a.first_step && a.unfinished[r] == 0) {
        // sequence already ended: emit pad and a zero log-prob row (AttModel.py:342-344)
        for (int v = threadIdx.x; v < V1; v += NT) g[v] = 0.f;
        if (threadIdx.x == 0) {
            if (a.tokens_out) a.tokens_out[r] = 0;
            if (a.seq_out) a.seq_out[(long)r * a.ld_seq + a.t] = 0;
            if (a.picked_lp) a.picked_lp[(long)r * a.ld_picked] = 0.f;
        }
        return;
    }

    float mx = -INFINITY;
    for (int v = threadIdx.x; v < V1; v += NT) { const float x = g[v]; row[v] = x; mx = fmaxf(mx, x); }
    mx = block_max<NT>(mx, s_red);
    float sum = 0.f;
    for (int v = threadIdx.x; v < V1; v += NT) sum += __expf(row[v] - mx);      // ex2.approx path: relative error ~1e-7 on the sum
    sum = block_sum<NT>(sum, s_red);
    const float lsum = logf(sum);
    const float m2 = (mx - mx) - lsum;            // max of the log-probs (second log_softmax)
    const int k_eff = a.topk > 0 ? a.topk : (a.select == 1 ? 1 : 0);
    float tv[KMAX];
    int ti[KMAX];
#pragma unroll
    for (int q = 0; q < KMAX; ++q) { tv[q] = -INFINITY; ti[q] = 0x7fffffff; }
    auto consider = [&](float lp, int v) {
        if (k_eff > 0 && lp > tv[KMAX - 1]) {
            float cv = lp;
            int ci = v;
#pragma unroll
            for (int q = 0; q < KMAX; ++q) {
                if (cv > tv[q]) { const float t0 = tv[q]; const int t1 = ti[q]; tv[q] = cv; ti[q] = ci; cv = t0; ci = t1; }
            }
        }
    };
    const bool edit = a.edits.any();
    for (int v = threadIdx.x; v < V1; v += NT) {
        const float lp = (row[v] - mx) - lsum;
        row[v] = lp;
        if (!edit) consider(lp, v);
    }
    if (edit) {
        // The reference's decode options edit the log-prob row AFTER the log-softmax, and the edited row is both what the next word is
        // chosen from and what is stored (AttModel.py:294-332).  A handful of columns change: one thread applies them to the shared copy.
        __syncthreads();
        if (threadIdx.x == 0) {
            const int t = a.t;
            const int prev = (t > 0 && a.prev_tokens != nullptr) ? a.prev_tokens[r] : -1;
            if (a.edits.constraint && prev >= 0 && prev < V1) row[prev] = -INFINITY;                   // never repeat the previous word
            bool prev_bad = false;
            for (int i = 0; i < a.edits.n_bad; ++i) prev_bad |= (t > 0 && a.edits.bad[i] == prev);
            if (prev_bad) row[0] = -INFINITY;                                                          // no end token after a bad ending
            if (a.edits.trigrams && t >= 3 && r < a.edits.trigram_rows && a.seq_out != nullptr) {
                const long long* sq = a.seq_out + (long)r * a.ld_seq;
                const long long p0 = sq[t - 2], p1 = sq[t - 1];
                for (int j = 0; j + 2 <= t - 1; ++j) {
                    if (sq[j] != p0 || sq[j + 1] != p1) continue;
                    const long long w = sq[j + 2];
                    bool first = true;
                    int count = 0;
                    for (int i = 0; i + 2 <= t - 1; ++i) {
                        if (sq[i] == p0 && sq[i + 1] == p1 && sq[i + 2] == w) { if (i < j) first = false; ++count; }
                    }
                    if (first && w >= 0 && w < V1) row[w] += ((float)count * -0.693f) * 2.0f;          // mask * -0.693 * alpha, alpha = 2 (AttModel.py:330-332)
                }
            }
        }
        __syncthreads();
        for (int v = threadIdx.x; v < V1; v += NT) consider(row[v], v);
    }
    // Second log_softmax (beam search, CaptionModel.py:204): its max is m2 = -lsum, so exp(lp - m2) = exp(x - mx) term by term and
    // its normaliser is the first pass's `sum` again (up to one rounding, ~1e-7 on the log-prob); no second exp pass is needed.
    const float l2 = lsum;
    if (a.twice) {
        for (int v = threadIdx.x; v < V1; v += NT) g[v] = (row[v] - m2) - l2;
    } else {
        for (int v = threadIdx.x; v < V1; v += NT) g[v] = row[v];
    }

    int greedy_tok = 0;
    for (int k = 0; k < k_eff; ++k) {
        float ov;
        int oi;
        block_argmax<NT>(tv[0], ti[0], s_red, s_idx, ov, oi);
        if (ti[0] == oi) {                        // the owner pops its head
#pragma unroll
            for (int q = 0; q + 1 < KMAX; ++q) { tv[q] = tv[q + 1]; ti[q] = ti[q + 1]; }
            tv[KMAX - 1] = -INFINITY;
            ti[KMAX - 1] = 0x7fffffff;
        }
        if (k == 0) greedy_tok = oi;
        if (a.topk > 0 && threadIdx.x == 0) {
            a.top_val[(long)r * a.topk + k] = a.twice ? (ov - m2) - l2 : ov;
            a.top_idx[(long)r * a.topk + k] = oi;
        }
    }

    if (a.select != 0) {
        int tok;
        if (a.select == 3) {
            tok = a.forced[r];
        } else if (a.select == 1) {
            tok = greedy_tok;
        } else {
            __syncthreads();
            const float inv_t = 1.0f / a.temperature;
            // order-preserving map float -> uint32 (for the threshold searches of the truncated samplers)
            auto okey = [](float x) { const uint32_t u = __float_as_uint(x); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); };
            uint32_t keep_from = 0;                    // sample among the words whose key is >= keep_from
            if (a.select == 4) {
                // top-k (CaptionModel.py:398-402): threshold = k-th largest log-prob, by bisection on the key bits (exact; ties at the threshold are all kept)
                const float kf = floorf(a.top);
                for (int bit = 31; bit >= 0; --bit) {
                    const uint32_t cand = keep_from | (1u << bit);
                    float cnt = 0.f;
                    for (int v = threadIdx.x; v < V1; v += NT) cnt += (okey(row[v]) >= cand) ? 1.f : 0.f;
                    cnt = block_sum<NT>(cnt, s_red);
                    if (cnt >= kf) keep_from = cand;
                }
            } else if (a.select == 5) {
                // nucleus (CaptionModel.py:388-397): a word is kept iff the probability mass of the strictly more likely words is < p
                float mxl = -INFINITY;
                for (int v = threadIdx.x; v < V1; v += NT) mxl = fmaxf(mxl, row[v]);
                mxl = block_max<NT>(mxl, s_red);
                float z = 0.f;
                for (int v = threadIdx.x; v < V1; v += NT) z += __expf((row[v] - mxl) * inv_t);
                z = block_sum<NT>(z, s_red);
                const float target = a.top * z;
                auto mass_above = [&](uint32_t key) {   // sum over the words with key > `key`
                    float m = 0.f;
                    for (int v = threadIdx.x; v < V1; v += NT) m += (okey(row[v]) > key) ? __expf((row[v] - mxl) * inv_t) : 0.f;
                    return block_sum<NT>(m, s_red);
                };
                // largest key F with mass_above(F) >= target; everything above F is kept (the most likely word always is)
                uint32_t F = 0;
                if (mass_above(0u) < target) keep_from = 0;
                else {
                    for (int bit = 31; bit >= 0; --bit) {
                        const uint32_t cand = F | (1u << bit);
                        if (mass_above(cand) >= target) F = cand;
                    }
                    keep_from = F + 1u;
                }
            }
            float bv = -INFINITY;
            int bi = 0x7fffffff;
            const unsigned long long sd = a.seed ^ g_vocab_seed_salt;           // see dropout.cuh: graph replays of the SCST step
            const uint32_t k0 = (uint32_t)sd, k1 = (uint32_t)(sd >> 32);
            for (int v = threadIdx.x; v < V1; v += NT) {
                if (okey(row[v]) < keep_from) continue;
                const uint32_t bits = philox_first((uint32_t)v, (uint32_t)r, (uint32_t)a.step, (uint32_t)(a.step >> 32), k0, k1);
                const float u = ((float)(bits >> 9) + 0.5f) * (1.0f / 8388608.0f);      // (0,1), 23 bits
                const float x = row[v] * inv_t - logf(-logf(u));                         // Gumbel-max sample of softmax(logp / T)
                if (x > bv) { bv = x; bi = v; }
            }
            float ov;
            block_argmax<NT>(bv, bi, s_red, s_idx, ov, tok);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (a.unfinished) a.unfinished[r] = (tok != 0) ? 1 : 0;
            if (a.tokens_out) a.tokens_out[r] = tok;
            if (a.seq_out) a.seq_out[(long)r * a.ld_seq + a.t] = tok;
            if (a.picked_lp) a.picked_lp[(long)r * a.ld_picked] = row[tok];
        }
    }
}

// Beam-search variant: the raw logits stay where the GEMM wrote them (they are normalised lazily, only for the rows that end up in
// the output); this kernel streams each row twice from L2 (max, then sum-exp + per-thread top-k) and writes only the row statistics
// and the top-k candidates.  Halves the HBM traffic of the step's vocabulary epilogue.
// Each thread sees only ~V1/256 elements, so it keeps just its two best; the k block-wide arg-max rounds pop list heads and a
// thread whose list runs dry (it owned >= 3 of the global top-k: rare) rescans its elements for the next one.
__global__ void __launch_bounds__(VT) vocab_stats_kernel(const VocabStepArgs a) {
    __shared__ float s_red[VT / 32];
    __shared__ int s_idx[VT / 32];
    const int r = blockIdx.x;
    const int V1 = a.V1;
    const float* g = a.logits + (long)r * a.ld;
    const bool vec = ((V1 & 3) == 0) && ((a.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.logits) & 15) == 0);
    float mx = -INFINITY;
    if (vec) {
        const float4* g4 = reinterpret_cast<const float4*>(g);
        for (int v = threadIdx.x; v < V1 / 4; v += VT) { const float4 x = g4[v]; mx = fmaxf(mx, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w))); }
    } else {
        for (int v = threadIdx.x; v < V1; v += VT) mx = fmaxf(mx, g[v]);
    }
    mx = block_max(mx, s_red);
    float t0v = -INFINITY, t1v = -INFINITY;
    int t0i = 0x7fffffff, t1i = 0x7fffffff;
    float sum = 0.f;
    auto visit = [&](float x, int v) {
        sum += __expf(x - mx);
        if (x > t1v) {                              // strict: earlier (lower) indices win ties
            if (x > t0v) { t1v = t0v; t1i = t0i; t0v = x; t0i = v; }
            else { t1v = x; t1i = v; }
        }
    };
    if (vec) {
        const float4* g4 = reinterpret_cast<const float4*>(g);
        for (int v = threadIdx.x; v < V1 / 4; v += VT) {
            const float4 x = g4[v];
            visit(x.x, 4 * v); visit(x.y, 4 * v + 1); visit(x.z, 4 * v + 2); visit(x.w, 4 * v + 3);
        }
    } else {
        for (int v = threadIdx.x; v < V1; v += VT) visit(g[v], v);
    }
    sum = block_sum(sum, s_red);
    const float lsum = logf(sum);
    const float m2 = (mx - mx) - lsum, l2 = lsum;
    if (threadIdx.x == 0) a.stats[r] = make_float2(mx, lsum);
    int popped = 0;
    for (int k = 0; k < a.topk; ++k) {
        float ov;
        int oi;
        block_argmax(t0v, t0i, s_red, s_idx, ov, oi);
        if (t0i == oi && oi != 0x7fffffff) {
            const float lastv = t0v;
            const int lasti = t0i;
            t0v = t1v; t0i = t1i;
            t1v = -INFINITY; t1i = 0x7fffffff;
            if (++popped >= 2 && t0i == 0x7fffffff) {
                // rescan my elements for the best one ordered after (lastv, lasti)
                auto consider = [&](float x, int v) {
                    const bool after = (x < lastv) || (x == lastv && v > lasti);
                    if (after && (x > t0v || (x == t0v && v < t0i))) { t0v = x; t0i = v; }
                };
                if (vec) {
                    const float4* g4 = reinterpret_cast<const float4*>(g);
                    for (int v = threadIdx.x; v < V1 / 4; v += VT) {
                        const float4 x = g4[v];
                        consider(x.x, 4 * v); consider(x.y, 4 * v + 1); consider(x.z, 4 * v + 2); consider(x.w, 4 * v + 3);
                    }
                } else {
                    for (int v = threadIdx.x; v < V1; v += VT) consider(g[v], v);
                }
            }
        }
        if (threadIdx.x == 0) {
            const float lp = (ov - mx) - lsum;
            a.top_val[(long)r * a.topk + k] = a.twice ? (lp - m2) - l2 : lp;
            a.top_idx[(long)r * a.topk + k] = oi;
        }
    }
}

// Single-pass variant of vocab_stats_kernel (one CTA per row, loads straight from global memory, 8 CTAs per SM): per-thread online
// softmax (running max, partial sum rescaled when the max grows) and one max-of-four test in front of the top-2 bookkeeping, so the
// row is read once and the common path is ~5 instructions per element.
__global__ void __launch_bounds__(VT) vocab_stats_online_kernel(const VocabStepArgs a) {
    __shared__ float s_red[VT / 32];
    __shared__ int s_idx[VT / 32];
    const int r = blockIdx.x;
    const int n4 = a.V1 >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(a.logits + (long)r * a.ld);
    constexpr float kL2E = 1.4426950408889634f;
    float t0v = -INFINITY, t1v = -INFINITY;
    int t0i = 0x7fffffff, t1i = 0x7fffffff;
    float m = -INFINITY, mL = -INFINITY, part = 0.f;
    for (int v = threadIdx.x; v < n4; v += VT) {
        const float4 x = g4[v];
        const float m4 = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
        if (m4 > m) {
            float sc;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sc) : "f"((m - m4) * kL2E));
            part = (m == -INFINITY) ? 0.f : part * sc;
            m = m4;
            mL = m4 * kL2E;
        }
        float e0, e1, e2, e3;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(x.x, kL2E, -mL)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(x.y, kL2E, -mL)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(x.z, kL2E, -mL)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fmaf(x.w, kL2E, -mL)));
        part += (e0 + e1) + (e2 + e3);
        if (m4 > t1v) {                             // strict: earlier (lower) indices win ties
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (xs[u] > t1v) {
                    if (xs[u] > t0v) { t1v = t0v; t1i = t0i; t0v = xs[u]; t0i = 4 * v + u; }
                    else { t1v = xs[u]; t1i = 4 * v + u; }
                }
            }
        }
    }
    const float mx = block_max(m, s_red);
    float sum = (m == -INFINITY) ? 0.f : part * __expf(m - mx);
    sum = block_sum(sum, s_red);
    const float lsum = logf(sum);
    const float m2 = (mx - mx) - lsum, l2 = lsum;
    if (threadIdx.x == 0) a.stats[r] = make_float2(mx, lsum);
    int popped = 0;
    for (int k = 0; k < a.topk; ++k) {
        float ov;
        int oi;
        block_argmax(t0v, t0i, s_red, s_idx, ov, oi);
        if (t0i == oi && oi != 0x7fffffff) {
            const float lastv = t0v;
            const int lasti = t0i;
            t0v = t1v; t0i = t1i;
            t1v = -INFINITY; t1i = 0x7fffffff;
            if (++popped >= 2 && t0i == 0x7fffffff) {
                auto consider = [&](float x, int v) {
                    const bool after = (x < lastv) || (x == lastv && v > lasti);
                    if (after && (x > t0v || (x == t0v && v < t0i))) { t0v = x; t0i = v; }
                };
                for (int v = threadIdx.x; v < n4; v += VT) {
                    const float4 x = g4[v];
                    consider(x.x, 4 * v); consider(x.y, 4 * v + 1); consider(x.z, 4 * v + 2); consider(x.w, 4 * v + 3);
                }
            }
        }
        if (threadIdx.x == 0) {
            const float lp = (ov - mx) - lsum;
            a.top_val[(long)r * a.topk + k] = a.twice ? (lp - m2) - l2 : lp;
            a.top_idx[(long)r * a.topk + k] = oi;
        }
    }
}

// 128-thread form of the single-pass kernel: 16 CTAs per SM (all 1280 rows of the headline shape resident at once, no tail wave) and
// four independent 128-bit loads in flight per thread.
constexpr int VT2 = 128;
__device__ __forceinline__ float block_max4(float v, float* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    return fmaxf(fmaxf(scratch[0], scratch[1]), fmaxf(scratch[2], scratch[3]));
}
__device__ __forceinline__ float block_sum4(float v, float* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    return (scratch[0] + scratch[1]) + (scratch[2] + scratch[3]);
}
__device__ __forceinline__ void block_argmax4(float v, int i, float* sval, int* sidx, float& out_v, int& out_i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sval[threadIdx.x >> 5] = v; sidx[threadIdx.x >> 5] = i; }
    __syncthreads();
    out_v = sval[0];
    out_i = sidx[0];
#pragma unroll
    for (int w = 1; w < VT2 / 32; ++w) {
        const float ov = sval[w];
        const int oi = sidx[w];
        if (ov > out_v || (ov == out_v && oi < out_i)) { out_v = ov; out_i = oi; }
    }
}

__global__ void __launch_bounds__(VT2) vocab_stats_online128_kernel(const VocabStepArgs a) {
    __shared__ float s_red[VT2 / 32];
    __shared__ int s_idx[VT2 / 32];
    const int r = blockIdx.x;
    const int n4 = a.V1 >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(a.logits + (long)r * a.ld);
    constexpr float kL2E = 1.4426950408889634f;
    float t0v = -INFINITY, t1v = -INFINITY;
    int t0i = 0x7fffffff, t1i = 0x7fffffff;
    float m = -INFINITY, mL = -INFINITY, part = 0.f;
    auto consume = [&](const float4 x, int v) {
        const float m4 = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
        if (m4 > m) {
            float sc;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sc) : "f"((m - m4) * kL2E));
            part = (m == -INFINITY) ? 0.f : part * sc;
            m = m4;
            mL = m4 * kL2E;
        }
        float e0, e1, e2, e3;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(x.x, kL2E, -mL)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(x.y, kL2E, -mL)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(x.z, kL2E, -mL)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fmaf(x.w, kL2E, -mL)));
        part += (e0 + e1) + (e2 + e3);
        if (m4 > t1v) {                             // strict: earlier (lower) indices win ties
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (xs[u] > t1v) {
                    if (xs[u] > t0v) { t1v = t0v; t1i = t0i; t0v = xs[u]; t0i = 4 * v + u; }
                    else { t1v = xs[u]; t1i = 4 * v + u; }
                }
            }
        }
    };
    int v = threadIdx.x;
    for (; v + 3 * VT2 < n4; v += 4 * VT2) {        // four loads in flight, consumed in index order (tie order is preserved)
        const float4 x0 = g4[v], x1 = g4[v + VT2], x2 = g4[v + 2 * VT2], x3 = g4[v + 3 * VT2];
        consume(x0, v); consume(x1, v + VT2); consume(x2, v + 2 * VT2); consume(x3, v + 3 * VT2);
    }
    for (; v < n4; v += VT2) consume(g4[v], v);
    const float mx = block_max4(m, s_red);
    float sum = (m == -INFINITY) ? 0.f : part * __expf(m - mx);
    sum = block_sum4(sum, s_red);
    const float lsum = logf(sum);
    const float m2 = (mx - mx) - lsum, l2 = lsum;
    if (threadIdx.x == 0) a.stats[r] = make_float2(mx, lsum);
    int popped = 0;
    for (int k = 0; k < a.topk; ++k) {
        float ov;
        int oi;
        block_argmax4(t0v, t0i, s_red, s_idx, ov, oi);
        if (t0i == oi && oi != 0x7fffffff) {
            const float lastv = t0v;
            const int lasti = t0i;
            t0v = t1v; t0i = t1i;
            t1v = -INFINITY; t1i = 0x7fffffff;
            if (++popped >= 2 && t0i == 0x7fffffff) {
                auto consider = [&](float x, int w) {
                    const bool after = (x < lastv) || (x == lastv && w > lasti);
                    if (after && (x > t0v || (x == t0v && w < t0i))) { t0v = x; t0i = w; }
                };
                for (int w = threadIdx.x; w < n4; w += VT2) {
                    const float4 x = g4[w];
                    consider(x.x, 4 * w); consider(x.y, 4 * w + 1); consider(x.z, 4 * w + 2); consider(x.w, 4 * w + 3);
                }
            }
        }
        if (threadIdx.x == 0) {
            const float lp = (ov - mx) - lsum;
            a.top_val[(long)r * a.topk + k] = a.twice ? (lp - m2) - l2 : lp;
            a.top_idx[(long)r * a.topk + k] = oi;
        }
    }
}

// Register-resident variant for rows of up to VT * 4 * NV elements (16-byte aligned): the row is read from L2/HBM exactly once, with
// all of a thread's loads in flight together; max, sum-exp, per-thread top-2 and the rare rescan then work on registers.  Same
// arithmetic (and the same tie order) as vocab_stats_kernel.
template <int NV>
__global__ void __launch_bounds__(VT) vocab_stats_reg_kernel(const VocabStepArgs a) {
    __shared__ float s_red[VT / 32];
    __shared__ int s_idx[VT / 32];
    const int r = blockIdx.x;
    const int n4 = a.V1 >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(a.logits + (long)r * a.ld);
    float4 xs[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = threadIdx.x + i * VT;
        xs[i] = v < n4 ? g4[v] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) mx = fmaxf(mx, fmaxf(fmaxf(xs[i].x, xs[i].y), fmaxf(xs[i].z, xs[i].w)));
    mx = block_max(mx, s_red);
    float t0v = -INFINITY, t1v = -INFINITY;
    int t0i = 0x7fffffff, t1i = 0x7fffffff;
    float sum = 0.f;
    auto visit = [&](float x, int v) {
        sum += __expf(x - mx);                      // padding lanes hold -inf: exp -> 0, never a candidate
        if (x > t1v) {                              // strict: earlier (lower) indices win ties
            if (x > t0v) { t1v = t0v; t1i = t0i; t0v = x; t0i = v; }
            else { t1v = x; t1i = v; }
        }
    };
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = 4 * (threadIdx.x + i * VT);
        visit(xs[i].x, v); visit(xs[i].y, v + 1); visit(xs[i].z, v + 2); visit(xs[i].w, v + 3);
    }
    sum = block_sum(sum, s_red);
    const float lsum = logf(sum);
    const float m2 = (mx - mx) - lsum, l2 = lsum;
    if (threadIdx.x == 0) a.stats[r] = make_float2(mx, lsum);
    int popped = 0;
    for (int k = 0; k < a.topk; ++k) {
        float ov;
        int oi;
        block_argmax(t0v, t0i, s_red, s_idx, ov, oi);
        if (t0i == oi && oi != 0x7fffffff) {
            const float lastv = t0v;
            const int lasti = t0i;
            t0v = t1v; t0i = t1i;
            t1v = -INFINITY; t1i = 0x7fffffff;
            if (++popped >= 2 && t0i == 0x7fffffff) {
                auto consider = [&](float x, int v) {
                    const bool after = (x < lastv) || (x == lastv && v > lasti);
                    if (after && (x > t0v || (x == t0v && v < t0i))) { t0v = x; t0i = v; }
                };
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int v = 4 * (threadIdx.x + i * VT);
                    if (v < a.V1) { consider(xs[i].x, v); consider(xs[i].y, v + 1); consider(xs[i].z, v + 2); consider(xs[i].w, v + 3); }
                }
            }
        }
        if (threadIdx.x == 0) {
            const float lp = (ov - mx) - lsum;
            a.top_val[(long)r * a.topk + k] = a.twice ? (lp - m2) - l2 : lp;
            a.top_idx[(long)r * a.topk + k] = oi;
        }
    }
}

// Streaming variant: persistent CTAs walk the rows; each row (V1 * 4 bytes, 16-byte aligned) arrives in shared memory through one
// cp.async.bulk while the previous row is being reduced (double buffer), so the L2/HBM read of row i+1 overlaps the arithmetic of
// row i and no thread ever waits on its own global loads.  Same arithmetic and tie order as vocab_stats_kernel.
__global__ void __launch_bounds__(VT) vocab_stats_stream_kernel(const VocabStepArgs a) {
    extern __shared__ __align__(16) unsigned char vs_smem[];
    __shared__ float s_red[VT / 32];
    __shared__ int s_idx[VT / 32];
    __shared__ __align__(8) uint64_t bar[2];
    const int V1 = a.V1, n4 = V1 >> 2;
    const uint32_t row_bytes = (uint32_t)V1 * 4u;
    float* buf0 = reinterpret_cast<float*>(vs_smem);
    float* buf1 = buf0 + V1;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar[0], 1);
        ptx::mbar_init(&bar[1], 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();
    int r = blockIdx.x;
    if (threadIdx.x == 0 && r < a.rows) {
        ptx::mbar_arrive_expect_tx(&bar[0], row_bytes);
        ptx::bulk_load_1d(buf0, a.logits + (long)r * a.ld, row_bytes, &bar[0]);
    }
    for (int it = 0; r < a.rows; r += gridDim.x, ++it) {
        const int cur = it & 1;
        const float* g = cur ? buf1 : buf0;
        const int rn = r + gridDim.x;
        if (threadIdx.x == 0 && rn < a.rows) {          // the other buffer was released by the __syncthreads that ended iteration it-1
            ptx::mbar_arrive_expect_tx(&bar[cur ^ 1], row_bytes);
            ptx::bulk_load_1d(cur ? buf0 : buf1, a.logits + (long)rn * a.ld, row_bytes, &bar[cur ^ 1]);
        }
        ptx::mbar_wait(&bar[cur], (it >> 1) & 1);
        const float4* g4 = reinterpret_cast<const float4*>(g);
        // One pass, online softmax per thread: running max m with the partial sum rescaled when it grows (rare after the first few
        // elements), one max-of-four test guards the top-2 bookkeeping.  The r01f capture showed the two-pass form issue-bound at
        // ~40 thread-instructions per element; this form needs about half.
        constexpr float kL2E = 1.4426950408889634f;
        float t0v = -INFINITY, t1v = -INFINITY;
        int t0i = 0x7fffffff, t1i = 0x7fffffff;
        float m = -INFINITY, mL = -INFINITY, part = 0.f;
        for (int v = threadIdx.x; v < n4; v += VT) {
            const float4 x = g4[v];
            const float m4 = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
            if (m4 > m) {
                float sc;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sc) : "f"((m - m4) * kL2E));
                part = (m == -INFINITY) ? 0.f : part * sc;
                m = m4;
                mL = m4 * kL2E;
            }
            float e0, e1, e2, e3;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(x.x, kL2E, -mL)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(x.y, kL2E, -mL)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(x.z, kL2E, -mL)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fmaf(x.w, kL2E, -mL)));
            part += (e0 + e1) + (e2 + e3);
            if (m4 > t1v) {                             // strict: earlier (lower) indices win ties
                const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (xs[u] > t1v) {
                        if (xs[u] > t0v) { t1v = t0v; t1i = t0i; t0v = xs[u]; t0i = 4 * v + u; }
                        else { t1v = xs[u]; t1i = 4 * v + u; }
                    }
                }
            }
        }
        const float mx = block_max(m, s_red);
        float sum = (m == -INFINITY) ? 0.f : part * __expf(m - mx);
        sum = block_sum(sum, s_red);
        const float lsum = logf(sum);
        const float m2 = (mx - mx) - lsum, l2 = lsum;
        if (threadIdx.x == 0) a.stats[r] = make_float2(mx, lsum);
        int popped = 0;
        for (int k = 0; k < a.topk; ++k) {
            float ov;
            int oi;
            block_argmax(t0v, t0i, s_red, s_idx, ov, oi);
            if (t0i == oi && oi != 0x7fffffff) {
                const float lastv = t0v;
                const int lasti = t0i;
                t0v = t1v; t0i = t1i;
                t1v = -INFINITY; t1i = 0x7fffffff;
                if (++popped >= 2 && t0i == 0x7fffffff) {
                    auto consider = [&](float x, int v) {
                        const bool after = (x < lastv) || (x == lastv && v > lasti);
                        if (after && (x > t0v || (x == t0v && v < t0i))) { t0v = x; t0i = v; }
                    };
                    for (int v = threadIdx.x; v < n4; v += VT) {
                        const float4 x = g4[v];
                        consider(x.x, 4 * v); consider(x.y, 4 * v + 1); consider(x.z, 4 * v + 2); consider(x.w, 4 * v + 3);
                    }
                }
            }
            if (threadIdx.x == 0) {
                const float lp = (ov - mx) - lsum;
                a.top_val[(long)r * a.topk + k] = a.twice ? (lp - m2) - l2 : lp;
                a.top_idx[(long)r * a.topk + k] = oi;
            }
        }
        __syncthreads();                                // everyone is done with buffer `cur` before it is refilled
    }
}

// Scheduled sampling (AttModel.py:145-154): with probability `prob` the word fed into step `col` is drawn from the model's own previous
// prediction exp(logprobs[:, col-1]) instead of the label.  One CTA per row; per-row uniform and the Gumbel-max draw come from the Philox
// stream (seed, site 6, step col).  Non-differentiable by construction (the reference samples from a detached tensor).
__global__ void __launch_bounds__(VT) ss_select_kernel(int V1, const float* __restrict__ prev_logp, long ld, const long long* __restrict__ labels,
                                                      long ld_labels, int col, unsigned long long seed, float prob, int* __restrict__ tok_out) {
    __shared__ float s_red[VT / 32];
    __shared__ int s_idx[VT / 32];
    const int r = blockIdx.x;
    seed ^= g_vocab_seed_salt;
    const uint32_t k0 = (uint32_t)seed ^ 0x6a09e667u, k1 = (uint32_t)(seed >> 32) ^ 0xbb67ae85u;
    const uint32_t ubits = philox_first(0xffffffffu, (uint32_t)r, (uint32_t)col, 6u, k0, k1);
    const float u_row = ((float)(ubits >> 9) + 0.5f) * (1.0f / 8388608.0f);
    if (!(u_row < prob)) {                       // keep the label (uniform over the whole CTA: no divergence at the barriers below)
        if (threadIdx.x == 0) tok_out[r] = (int)labels[(long)r * ld_labels + col];
        return;
    }
    const float* g = prev_logp + (long)r * ld;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int v = threadIdx.x; v < V1; v += VT) {
        const uint32_t bits = philox_first((uint32_t)v, (uint32_t)r, (uint32_t)col, 6u, k0, k1);
        const float u = ((float)(bits >> 9) + 0.5f) * (1.0f / 8388608.0f);
        const float x = g[v] - logf(-logf(u));                  // Gumbel-max draw from softmax(logp) = exp(logp)
        if (x > bv) { bv = x; bi = v; }
    }
    float ov;
    int tok;
    block_argmax(bv, bi, s_red, s_idx, ov, tok);
    if (threadIdx.x == 0) tok_out[r] = tok;
}

__global__ void mask_rows_kernel(ActView x, int R, int cols, const float* __restrict__ mask, long ld_mask) {
    const int row = blockIdx.x;              // row = img * R + r
    const int img = row / R, r = row % R;
    if (mask[(long)img * ld_mask + r] != 0.f) return;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        x.f[(long)row * x.ld + c] = 0.f;
        if (x.hi) { x.hi[(long)row * x.ld + c] = __float2half(0.f); x.lo[(long)row * x.ld + c] = __float2half(0.f); }
    }
}

}  // namespace

int vocab_step_launch(const VocabStepArgs& a, cudaStream_t stream) {
    if (a.rows <= 0) return 0;
    CAPB_REQUIRE(a.topk <= 16, "beam size up to 16");
    if (a.stats != nullptr) {
        CAPB_REQUIRE(a.select == 0 && a.topk > 0, "stats mode is the beam-search epilogue");
        const bool vec = ((a.V1 & 3) == 0) && ((a.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.logits) & 15) == 0);
        const size_t stream_smem = sizeof(float) * 2 * (size_t)a.V1;
        // default: single-pass online kernel; "stream" / "reg" / "plain" select the other variants for A/B timing (profiles/)
        static const char* variant = getenv("CAPB200_VOCAB_STATS");
        const char vsel = variant ? variant[0] : 'o';
        if (vec && vsel == 'o' && !(variant && variant[1] == '2')) {
            vocab_stats_online128_kernel<<<a.rows, VT2, 0, stream>>>(a);
        } else if (vec && vsel == 'o') {                                  // "o2": the 256-thread form
            vocab_stats_online_kernel<<<a.rows, VT, 0, stream>>>(a);
        } else if (vec && vsel == 's' && stream_smem <= 100 * 1024) {
            static std::atomic<unsigned long long> configured{0};
            if (first_use_on_device(configured)) {
                CAPB_CHECK_CUDA(cudaFuncSetAttribute(vocab_stats_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(100 * 1024)));
            }
            const int grid = a.rows < 2 * 148 ? a.rows : 2 * 148;       // two resident CTAs per SM, each double-buffering one row
            vocab_stats_stream_kernel<<<grid, VT, stream_smem, stream>>>(a);
        } else if (vec && vsel == 'r' && a.V1 <= VT * 4 * 10) {
            vocab_stats_reg_kernel<10><<<a.rows, VT, 0, stream>>>(a);
        } else {
            vocab_stats_kernel<<<a.rows, VT, 0, stream>>>(a);
        }
        CAPB_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    const size_t smem = sizeof(float) * (size_t)a.V1;
    CAPB_REQUIRE(smem <= 200 * 1024, "vocabulary larger than 51200 entries needs the multi-pass variant");
    static std::atomic<unsigned long long> configured{0};
    if (first_use_on_device(configured)) {
        CAPB_CHECK_CUDA(cudaFuncSetAttribute(vocab_step_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        CAPB_CHECK_CUDA(cudaFuncSetAttribute(vocab_step_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        CAPB_CHECK_CUDA(cudaFuncSetAttribute(vocab_step_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    }
    if (a.topk <= 2 && a.select != 0 && a.rows <= 2 * 148) {
        static std::atomic<unsigned long long> configured2{0};
        if (first_use_on_device(configured2)) {
            CAPB_CHECK_CUDA(cudaFuncSetAttribute(vocab_step_kernel<2, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        }
        vocab_step_kernel<2, 1024><<<a.rows, 1024, smem, stream>>>(a);
    }
    else if (a.topk <= 2) vocab_step_kernel<2><<<a.rows, VT, smem, stream>>>(a);
    else if (a.topk <= 8) vocab_step_kernel<8><<<a.rows, VT, smem, stream>>>(a);
    else vocab_step_kernel<16><<<a.rows, VT, smem, stream>>>(a);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int ss_select_launch(int rows, int V1, const float* prev_logp, long ld, const long long* labels, long ld_labels, int col, unsigned long long seed, float prob,
                     int* tok_out, cudaStream_t stream) {
    if (rows <= 0) return 0;
    ss_select_kernel<<<rows, VT, 0, stream>>>(V1, prev_logp, ld, labels, ld_labels, col, seed, prob, tok_out);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int mask_rows_launch(ActView x, int n_images, int R, int cols, const float* mask, long ld_mask, cudaStream_t stream) {
    if (n_images <= 0 || mask == nullptr) return 0;
    mask_rows_kernel<<<n_images * R, 128, 0, stream>>>(x, R, cols, mask, ld_mask);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace capb200
