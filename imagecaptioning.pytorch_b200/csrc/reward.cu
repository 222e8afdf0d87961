// SCST reward on the device: CIDEr-D over token ids, self-critical baseline, and RewardCriterion forward / backward.
//
// Replaces the host-side Python path the reference takes every SCST step:
//   get_self_critical_reward   captioning/utils/rewards.py:41-81   (.cpu().numpy(), str() of every id, dict loops)
//   CiderD.compute_score       cider/pyciderevalcap/ciderD/ciderD.py:31-56 -> ciderD_scorer.py:17-32,53-79,156-208
//   RewardCriterion.forward    captioning/modules/losses.py:22-37
// Arithmetic follows the reference exactly, in float64 like numpy:
//   * a caption is its tokens up to and INCLUDING the first 0 (array_to_str keeps "0"),
//   * tf-idf weight of an n-gram = tf * (log(ref_len) - log(max(1, df))), df from the preprocessed table,
//   * per order n: sum over the hypothesis' distinct n-grams of min(w_h, w_r) * w_r, divided by |h||r| when both are
//     non-zero, times exp(-(len_h - len_r)^2 / (2 * 6^2)) where len counts BIGRAMS (ciderD_scorer.py:177-178),
//   * score = 10 * mean over references of the mean over n = 1..4.
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace capb200 {

constexpr int CIDER_MAXL = 64;     // max tokens in a caption incl. the closing 0
constexpr int CIDER_N = 4;

struct CiderSlot {
    int key[4];
    double idf;
};

struct CiderTable {
    CiderSlot* slots = nullptr;   // device, open addressing, key[0] == -2 marks an empty slot
    unsigned long long mask = 0;  // capacity - 1
    double log_ref_len = 0.0;
    long entries = 0;
};

__host__ __device__ inline unsigned long long cider_hash(int a, int b, int c, int d) {
    unsigned long long h = 0x9E3779B97F4A7C15ull;
    const int k[4] = {a, b, c, d};
    for (int i = 0; i < 4; ++i) {
        h ^= (unsigned long long)(unsigned int)k[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 31;
    }
    return h;
}

CiderTable* cider_table_create(const int* keys, const double* df, long n, double ref_len, cudaStream_t stream) {
    unsigned long long cap = 64;
    while (cap < (unsigned long long)(2 * n + 1)) cap <<= 1;
    std::vector<CiderSlot> host(cap);
    for (auto& s : host) { s.key[0] = -2; s.key[1] = s.key[2] = s.key[3] = -2; s.idf = 0.0; }
    const double log_ref = log(ref_len);
    for (long i = 0; i < n; ++i) {
        const int* k = keys + 4 * i;
        unsigned long long h = cider_hash(k[0], k[1], k[2], k[3]) & (cap - 1);
        while (host[h].key[0] != -2) {
            if (host[h].key[0] == k[0] && host[h].key[1] == k[1] && host[h].key[2] == k[2] && host[h].key[3] == k[3]) break;
            h = (h + 1) & (cap - 1);
        }
        host[h].key[0] = k[0]; host[h].key[1] = k[1]; host[h].key[2] = k[2]; host[h].key[3] = k[3];
        host[h].idf = log_ref - log(df[i] > 1.0 ? df[i] : 1.0);
    }
    CiderTable* t = new CiderTable();
    t->mask = cap - 1;
    t->log_ref_len = log_ref;
    t->entries = n;
    if (cudaMalloc(&t->slots, cap * sizeof(CiderSlot)) != cudaSuccess) {
        set_error("cider_table_create: cudaMalloc failed");
        delete t;
        return nullptr;
    }
    if (cudaMemcpyAsync(t->slots, host.data(), cap * sizeof(CiderSlot), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess) {
        set_error("cider_table_create: upload failed");
        cudaFree(t->slots);
        delete t;
        return nullptr;
    }
    return t;
}

void cider_table_destroy(CiderTable* t) {
    if (t == nullptr) return;
    cudaFree(t->slots);
    delete t;
}

namespace {

__device__ __forceinline__ double cider_idf(const CiderSlot* __restrict__ slots, unsigned long long mask, double log_ref_len, const int* tok, int n) {
    const int k0 = tok[0], k1 = n > 1 ? tok[1] : -1, k2 = n > 2 ? tok[2] : -1, k3 = n > 3 ? tok[3] : -1;
    unsigned long long h = cider_hash(k0, k1, k2, k3) & mask;
    for (;;) {
        const CiderSlot& s = slots[h];
        if (s.key[0] == -2) return log_ref_len;                 // unseen n-gram: df = 0 -> log(max(1, 0)) = 0
        if (s.key[0] == k0 && s.key[1] == k1 && s.key[2] == k2 && s.key[3] == k3) return s.idf;
        h = (h + 1) & mask;
    }
}

__device__ __forceinline__ bool same_gram(const int* a, const int* b, int n) {
    for (int i = 0; i < n; ++i) if (a[i] != b[i]) return false;
    return true;
}

// Builds the tf-idf description of one caption held in shared memory.
//   gram index g = n * MAXL + p (order n+1 starting at position p);  w[g] = tf * idf for the first occurrence, else 0
//   nrm[n] = sqrt(sum w^2);  returns through `w`, `valid` (1 = distinct n-gram present)
__device__ void cider_vectorise(const CiderSlot* slots, unsigned long long mask, double log_ref_len, const int* tok, int len, double* w,
                                unsigned char* valid, double* nrm) {
    for (int g = threadIdx.x; g < CIDER_N * CIDER_MAXL; g += blockDim.x) {
        const int n = g / CIDER_MAXL + 1, p = g % CIDER_MAXL;
        double wv = 0.0;
        unsigned char ok = 0;
        if (p + n <= len) {
            bool first = true;
            int tf = 0;
            for (int q = 0; q + n <= len; ++q) {
                if (same_gram(tok + p, tok + q, n)) {
                    if (q < p) { first = false; break; }
                    ++tf;
                }
            }
            if (first) {
                wv = (double)tf * cider_idf(slots, mask, log_ref_len, tok + p, n);
                ok = 1;
            }
        }
        w[g] = wv;
        valid[g] = ok;
    }
    __syncthreads();
    if (threadIdx.x < CIDER_N) {
        double s = 0.0;
        for (int p = 0; p < CIDER_MAXL; ++p) {
            const int g = threadIdx.x * CIDER_MAXL + p;
            if (valid[g]) s += w[g] * w[g];
        }
        nrm[threadIdx.x] = sqrt(s);
    }
    __syncthreads();
}

// one CTA per hypothesis: hyps 0..S-1 are the samples (image i / n), S..S+B-1 the greedy captions (image i - S)
__global__ void __launch_bounds__(256) cider_score_kernel(const CiderSlot* __restrict__ slots, unsigned long long mask, double log_ref_len,
                                                          const long long* __restrict__ sampled, int S, const long long* __restrict__ greedy, int B,
                                                          int T, const int* __restrict__ refs, const int* __restrict__ ref_offsets, int L,
                                                          double* __restrict__ scores) {
    __shared__ int h_tok[CIDER_MAXL], r_tok[CIDER_MAXL];
    __shared__ double h_w[CIDER_N * CIDER_MAXL], r_w[CIDER_N * CIDER_MAXL], contrib[CIDER_N * CIDER_MAXL];
    __shared__ unsigned char h_valid[CIDER_N * CIDER_MAXL], r_valid[CIDER_N * CIDER_MAXL];
    __shared__ double h_nrm[CIDER_N], r_nrm[CIDER_N], acc[CIDER_N];
    __shared__ int h_len, r_len;
    const int hyp = blockIdx.x;
    const int n_per = (B > 0) ? S / B : 1;
    const int img = hyp < S ? hyp / n_per : hyp - S;
    const long long* src = hyp < S ? sampled + (long)hyp * T : greedy + (long)(hyp - S) * T;
    if (threadIdx.x == 0) {
        int len = 0;
        for (int i = 0; i < T && i < CIDER_MAXL; ++i) { const int v = (int)src[i]; h_tok[len++] = v; if (v == 0) break; }
        h_len = len;
    }
    __syncthreads();
    cider_vectorise(slots, mask, log_ref_len, h_tok, h_len, h_w, h_valid, h_nrm);
    const int hl = h_len > 1 ? h_len - 1 : 0;               // "length" = number of bigrams
    const int r0 = ref_offsets[img], r1 = ref_offsets[img + 1];
    double total = 0.0;                                       // only thread 0 uses it
    for (int r = r0; r < r1; ++r) {
        if (threadIdx.x == 0) {
            int len = 0;
            for (int i = 0; i < L && i < CIDER_MAXL; ++i) { const int v = refs[(long)r * L + i]; r_tok[len++] = v; if (v == 0) break; }
            r_len = len;
        }
        __syncthreads();
        cider_vectorise(slots, mask, log_ref_len, r_tok, r_len, r_w, r_valid, r_nrm);
        for (int g = threadIdx.x; g < CIDER_N * CIDER_MAXL; g += blockDim.x) {
            double c = 0.0;
            if (h_valid[g]) {
                const int n = g / CIDER_MAXL + 1, p = g % CIDER_MAXL;
                double wr = 0.0;
                for (int q = 0; q + n <= r_len; ++q) {
                    const int gr = (n - 1) * CIDER_MAXL + q;
                    if (r_valid[gr] && same_gram(h_tok + p, r_tok + q, n)) { wr = r_w[gr]; break; }
                }
                c = fmin(h_w[g], wr) * wr;
            }
            contrib[g] = c;
        }
        __syncthreads();
        if (threadIdx.x < CIDER_N) {
            const int n = threadIdx.x;
            double v = 0.0;
            for (int p = 0; p < CIDER_MAXL; ++p) v += contrib[n * CIDER_MAXL + p];
            if (h_nrm[n] != 0.0 && r_nrm[n] != 0.0) v /= (h_nrm[n] * r_nrm[n]);
            const int rl = r_len > 1 ? r_len - 1 : 0;
            const double delta = (double)(hl - rl);
            v *= exp(-(delta * delta) / (2.0 * 6.0 * 6.0));
            acc[n] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) total += (acc[0] + acc[1] + acc[2] + acc[3]) / (double)CIDER_N;
        __syncthreads();
    }
    if (threadIdx.x == 0) scores[hyp] = (r1 > r0) ? total / (double)(r1 - r0) * 10.0 : 0.0;
}

__global__ void cider_reward_kernel(const double* __restrict__ scores, int S, int B, float* __restrict__ reward, long ld, int cols) {
    const int i = blockIdx.x;
    const int n_per = S / B;
    const float rwd = (float)(scores[i] - scores[S + i / n_per]);     // fp64 difference, then the .to(float32) of loss_wrapper.py:71
    for (int c = threadIdx.x; c < cols; c += blockDim.x) reward[(long)i * ld + c] = rwd;
}

// new_self_critical (losses.py:168-187): reward_i = s_i - mean of the image's other samples, in fp32 after scores.type_as(input) (:62)
__global__ void cider_reward_loo_kernel(const double* __restrict__ scores, int n_per, float* __restrict__ reward, long ld, int cols) {
    const int i = blockIdx.x;
    const int first = (i / n_per) * n_per;
    float sum = 0.f;
    for (int j = 0; j < n_per; ++j) sum += (float)scores[first + j];
    const float s = (float)scores[i];
    const float rwd = s - (sum - s) / (float)(n_per - 1);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) reward[(long)i * ld + c] = rwd;
}

// RewardCriterion (losses.py:22-37): single CTA, deterministic tree reduction
__global__ void __launch_bounds__(256) reward_criterion_fwd_kernel(const float* __restrict__ lp, long ld_row, long ld_t, const long long* __restrict__ seq,
                                                                   const float* __restrict__ reward, int N, int T, float* loss_mean,
                                                                   float* loss_rows, float* mask_sum) {
    __shared__ float s_out[256], s_msk[256];
    float o_acc = 0.f, m_acc = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float ro = 0.f, rm = 0.f;
        for (int t = 0; t < T; ++t) {
            const float m = (t == 0 || seq[(long)n * T + t - 1] > 0) ? 1.f : 0.f;
            const long long tok = seq[(long)n * T + t];
            const float v = -lp[(long)n * ld_row + (long)t * ld_t + tok] * reward[(long)n * T + t] * m;
            ro += v;
            rm += m;
        }
        if (loss_rows) loss_rows[n] = ro / rm;
        o_acc += ro;
        m_acc += rm;
    }
    s_out[threadIdx.x] = o_acc;
    s_msk[threadIdx.x] = m_acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { s_out[threadIdx.x] += s_out[threadIdx.x + w]; s_msk[threadIdx.x] += s_msk[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (loss_mean) *loss_mean = s_out[0] / s_msk[0];
        if (mask_sum) *mask_sum = s_msk[0];
    }
}

__global__ void reward_criterion_bwd_kernel(const long long* __restrict__ seq, const float* __restrict__ reward, int N, int T,
                                            const float* __restrict__ mask_sum, float upstream, float* __restrict__ grad, long ld_row, long ld_t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * T) return;
    const int n = i / T, t = i % T;
    const float m = (t == 0 || seq[(long)n * T + t - 1] > 0) ? 1.f : 0.f;
    const long long tok = seq[i];
    grad[(long)n * ld_row + (long)t * ld_t + tok] = -reward[i] * m / (*mask_sum) * upstream;
}

}  // namespace

int cider_reward_launch(const CiderTable* t, const long long* sampled, int S, const long long* greedy, int B, int T, const int* refs,
                        const int* ref_offsets, int L, double* scores, float* reward, long ld_reward, int reward_cols, cudaStream_t stream) {
    CAPB_REQUIRE(t != nullptr, "CIDEr-D table not initialised (init_scorer)");
    CAPB_REQUIRE(B > 0 && S % B == 0, "sample rows must be a multiple of the image count");
    CAPB_REQUIRE(T <= CIDER_MAXL && L <= CIDER_MAXL, "caption length above 64 tokens");
    // greedy == nullptr: score the samples only; the reward baseline is then the mean of the image's other samples
    const int hyps = greedy != nullptr ? S + B : S;
    if (hyps == 0) return 0;
    cider_score_kernel<<<hyps, 256, 0, stream>>>(t->slots, t->mask, t->log_ref_len, sampled, S, greedy, B, T, refs, ref_offsets, L, scores);
    CAPB_CHECK_CUDA(cudaGetLastError());
    if (reward != nullptr && S > 0) {
        if (greedy != nullptr) {
            cider_reward_kernel<<<S, 32, 0, stream>>>(scores, S, B, reward, ld_reward, reward_cols);
        } else {
            CAPB_REQUIRE(S / B >= 2, "the leave-one-out baseline needs at least two samples per image");
            cider_reward_loo_kernel<<<S, 32, 0, stream>>>(scores, S / B, reward, ld_reward, reward_cols);
        }
        CAPB_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

int reward_criterion_fwd_launch(const float* logprobs, long ld_row, long ld_t, const long long* seq, const float* reward, int N, int T,
                                float* loss_mean, float* loss_rows, float* mask_sum, cudaStream_t stream) {
    reward_criterion_fwd_kernel<<<1, 256, 0, stream>>>(logprobs, ld_row, ld_t, seq, reward, N, T, loss_mean, loss_rows, mask_sum);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int reward_criterion_bwd_launch(const long long* seq, const float* reward, int N, int T, const float* mask_sum, float upstream,
                                float* grad, long ld_row, long ld_t, cudaStream_t stream) {
    if (N * T <= 0) return 0;
    reward_criterion_bwd_kernel<<<cdiv(N * T, 256), 256, 0, stream>>>(seq, reward, N, T, mask_sum, upstream, grad, ld_row, ld_t);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace capb200
