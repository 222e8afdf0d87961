// Element-wise / small-reduction kernels of the per-timestep core.
//
//   state_gather_embed   embedding lookup (+ReLU) and parent-beam state reorder   AttModel.py:168 (embed), CaptionModel.py:105-108
//   lstm_pointwise       nn.LSTMCell gate math (i,f,g,o)                          AttModel.py:628,635
//   maxout_pointwise     NewFC maxout LSTM core                                   FCModel.py:25-42
//   additive_attention   softmax_r(w . tanh(p_att[r] + W_h h)) then sum_r a_r v_r AttModel.py:728-748
//
// All kernels read per-IMAGE features (p_att, att) by row / rows_per_image instead of the beam-replicated copies the
// reference materialises with repeat_tensors (AttModel.py:241-243).
#include "common.cuh"
#include "kernels.cuh"

namespace capb200 {

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void store_act(const ActView& o, long row, int col, float v) {
    o.f[row * o.ld + col] = v;
    if (o.hi != nullptr) {
        __half h, l;
        split_f32(v, h, l);
        o.hi[row * o.ld + col] = h;
        o.lo[row * o.ld + col] = l;
    }
}

// grid = rows, block = 256.  src_row < 0 means "fresh zero state" (first step).
__global__ void state_gather_embed_kernel(int rows, const int* __restrict__ tokens, const int* __restrict__ src_row,
                                          const float* __restrict__ emb, long ld_emb, int E, int relu, ActView xt,
                                          int H, int nstate, StateCopy sc0, StateCopy sc1) {
    const int r = blockIdx.x;
    if (r >= rows) return;
    const int tok = tokens[r];
    const float* e = emb + (long)tok * ld_emb;
    for (int c = threadIdx.x; c < E; c += blockDim.x) {
        float v = __ldg(e + c);
        if (relu) v = fmaxf(v, 0.f);
        store_act(xt, r, c, v);
    }
    const int src = src_row ? src_row[r] : r;
    for (int s = 0; s < nstate; ++s) {
        const StateCopy& sc = (s == 0) ? sc0 : sc1;
        for (int c = threadIdx.x; c < H; c += blockDim.x) {
            const float v = (src < 0) ? 0.f : sc.src[(long)src * sc.ld_src + c];
            store_act(sc.dst, r, c, v);
        }
    }
}

// gates [rows, 4H] in the nn.LSTMCell order i,f,g,o (bias already added by the GEMM epilogue).
__global__ void lstm_pointwise_kernel(int rows, int H, const float* __restrict__ gates, long ld_g, const int* __restrict__ src_row,
                                      const float* __restrict__ c_prev, long ld_cp, float* __restrict__ c_out, long ld_co, ActView h_out) {
    const long total = (long)rows * H;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int r = (int)(i / H), c = (int)(i % H);
        const float* g = gates + (long)r * ld_g;
        const float gi = g[c], gf = g[H + c], gg = g[2 * H + c], go = g[3 * H + c];
        const int src = src_row ? src_row[r] : r;
        const float cp = (src < 0 || c_prev == nullptr) ? 0.f : c_prev[(long)src * ld_cp + c];
        const float cn = sigmoidf_(gf) * cp + sigmoidf_(gi) * tanhf(gg);
        const float hn = sigmoidf_(go) * tanhf(cn);
        c_out[(long)r * ld_co + c] = cn;
        store_act(h_out, r, c, hn);
    }
}

// sums [rows, 5H]: sigmoid(i), sigmoid(f), sigmoid(o), then two candidates whose max is the cell input.
__global__ void maxout_pointwise_kernel(int rows, int H, const float* __restrict__ sums, long ld_s, const int* __restrict__ src_row,
                                        const float* __restrict__ c_prev, long ld_cp, float* __restrict__ c_out, long ld_co, ActView h_out) {
    const long total = (long)rows * H;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int r = (int)(i / H), c = (int)(i % H);
        const float* s = sums + (long)r * ld_s;
        const float gi = sigmoidf_(s[c]), gf = sigmoidf_(s[H + c]), go = sigmoidf_(s[2 * H + c]);
        const float gg = fmaxf(s[3 * H + c], s[4 * H + c]);
        const int src = src_row ? src_row[r] : r;
        const float cp = (src < 0 || c_prev == nullptr) ? 0.f : c_prev[(long)src * ld_cp + c];
        const float cn = gf * cp + gi * gg;
        const float hn = go * tanhf(cn);
        c_out[(long)r * ld_co + c] = cn;
        store_act(h_out, r, c, hn);
    }
}

// One CTA per image; it serves the image's `rpi` rows (beams / samples) so p_att[i] and att[i] are read once per CTA.
//   att_h  [rows, A]      h2att(h) incl. bias
//   p_att  [B, R, A]      ctx2att(att) incl. bias        (per image)
//   att    [B, R, H]      att_embed output               (per image)
//   mask   [B, R] or null 1 = valid region
constexpr int ATT_JB = 5;       // rows handled per pass (beam 5 = one pass)
constexpr int ATT_THREADS = 256;
__global__ void __launch_bounds__(ATT_THREADS) additive_attention_kernel(int rpi, int R, int A, int H, const float* __restrict__ att_h, long ld_ah,
                                                                          const float* __restrict__ p_att, long ld_pa,
                                                                          const float* __restrict__ att, long ld_at,
                                                                          const float* __restrict__ mask, long ld_mask,
                                                                          const float* __restrict__ alpha_w, const float* __restrict__ alpha_b_ptr, ActView out) {
    extern __shared__ float sm[];
    float* s_ah = sm;                        // [ATT_JB][A]
    float* s_w = s_ah + ATT_JB * A;          // [A]
    float* s_score = s_w + A;                // [ATT_JB][R]
    const int img = blockIdx.x;
    const float alpha_b = __ldg(alpha_b_ptr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = ATT_THREADS / 32;
    for (int a = threadIdx.x; a < A; a += ATT_THREADS) s_w[a] = __ldg(alpha_w + a);
    for (int j0 = 0; j0 < rpi; j0 += ATT_JB) {
        const int nj = min(ATT_JB, rpi - j0);
        __syncthreads();
        for (int i = threadIdx.x; i < nj * A; i += ATT_THREADS) {
            const int j = i / A, a = i % A;
            s_ah[j * A + a] = att_h[(long)(img * rpi + j0 + j) * ld_ah + a];
        }
        __syncthreads();
        // scores: warp per region
        for (int r = warp; r < R; r += nwarp) {
            const float* pr = p_att + ((long)img * R + r) * ld_pa;
            float part[ATT_JB];
#pragma unroll
            for (int j = 0; j < ATT_JB; ++j) part[j] = 0.f;
            for (int a = lane; a < A; a += 32) {
                const float pv = __ldg(pr + a);
                const float w = s_w[a];
#pragma unroll
                for (int j = 0; j < ATT_JB; ++j)
                    if (j < nj) part[j] = fmaf(w, tanhf(pv + s_ah[j * A + a]), part[j]);
            }
#pragma unroll
            for (int j = 0; j < ATT_JB; ++j) {
                float v = part[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && j < nj) s_score[j * R + r] = v + alpha_b;
            }
        }
        __syncthreads();
        // softmax over regions (+ optional mask renormalisation): warp j handles row j
        if (warp < nj) {
            float* sc = s_score + warp * R;
            float mx = -INFINITY;
            for (int r = lane; r < R; r += 32) mx = fmaxf(mx, sc[r]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.f;
            for (int r = lane; r < R; r += 32) { const float e = expf(sc[r] - mx); sc[r] = e; sum += e; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float inv = 1.0f / sum;
            float msum = 0.f;
            for (int r = lane; r < R; r += 32) {
                float w = sc[r] * inv;
                if (mask != nullptr) { w *= mask[(long)img * ld_mask + r]; msum += w; }
                sc[r] = w;
            }
            if (mask != nullptr) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
                for (int r = lane; r < R; r += 32) sc[r] = sc[r] / msum;
            }
        }
        __syncthreads();
        // weighted sum of the image's region features; each thread owns feature columns c, c+256, ...
        for (int c = threadIdx.x; c < H; c += ATT_THREADS) {
            float acc[ATT_JB];
#pragma unroll
            for (int j = 0; j < ATT_JB; ++j) acc[j] = 0.f;
            const float* ap = att + (long)img * R * ld_at + c;
            for (int r = 0; r < R; ++r) {
                const float v = __ldg(ap + (long)r * ld_at);
#pragma unroll
                for (int j = 0; j < ATT_JB; ++j)
                    if (j < nj) acc[j] = fmaf(s_score[j * R + r], v, acc[j]);
            }
#pragma unroll
            for (int j = 0; j < ATT_JB; ++j)
                if (j < nj) store_act(out, (long)img * rpi + j0 + j, c, acc[j]);
        }
    }
}

}  // namespace

int state_gather_embed_launch(int rows, const int* tokens, const int* src_row, const float* emb, long ld_emb, int E, int relu,
                              ActView xt, int H, int nstate, StateCopy sc0, StateCopy sc1, cudaStream_t stream) {
    if (rows <= 0) return 0;
    state_gather_embed_kernel<<<rows, 256, 0, stream>>>(rows, tokens, src_row, emb, ld_emb, E, relu, xt, H, nstate, sc0, sc1);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int pw_blocks(long total) {
    long b = (total + 255) / 256;
    return (int)(b > 148 * 8 ? 148 * 8 : b);
}

int lstm_pointwise_launch(int rows, int H, const float* gates, long ld_g, const int* src_row, const float* c_prev, long ld_cp,
                          float* c_out, long ld_co, ActView h_out, cudaStream_t stream) {
    if (rows <= 0) return 0;
    lstm_pointwise_kernel<<<pw_blocks((long)rows * H), 256, 0, stream>>>(rows, H, gates, ld_g, src_row, c_prev, ld_cp, c_out, ld_co, h_out);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int maxout_pointwise_launch(int rows, int H, const float* sums, long ld_s, const int* src_row, const float* c_prev, long ld_cp,
                            float* c_out, long ld_co, ActView h_out, cudaStream_t stream) {
    if (rows <= 0) return 0;
    maxout_pointwise_kernel<<<pw_blocks((long)rows * H), 256, 0, stream>>>(rows, H, sums, ld_s, src_row, c_prev, ld_cp, c_out, ld_co, h_out);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int additive_attention_launch(int n_images, int rpi, int R, int A, int H, const float* att_h, long ld_ah, const float* p_att, long ld_pa,
                              const float* att, long ld_at, const float* mask, long ld_mask, const float* alpha_w, const float* alpha_b,
                              ActView out, cudaStream_t stream) {
    if (n_images <= 0 || rpi <= 0) return 0;
    const size_t smem = sizeof(float) * ((size_t)ATT_JB * A + A + (size_t)ATT_JB * R);
    CAPB_REQUIRE(smem <= 48 * 1024, "attention: att_hid_size / region count too large for the shared-memory staging");
    additive_attention_kernel<<<n_images, ATT_THREADS, smem, stream>>>(rpi, R, A, H, att_h, ld_ah, p_att, ld_pa, att, ld_at, mask, ld_mask,
                                                                        alpha_w, alpha_b, out);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace capb200
