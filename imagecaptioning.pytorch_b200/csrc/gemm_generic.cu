// Generic fp32 CUDA-core GEMM for the SCST backward pass:  C[M,N] = op(A)[M,K] * op(B)[K,N] (+ C if accumulate)
//
//   TA = 0: A stored [M, K] (pitch lda)      TA = 1: A stored [K, M]   (dW = dY^T * X reads dY this way)
//   TB = 0: B stored [K, N] (pitch ldb)      TB = 1: B stored [N, K]   (nn.Linear forward: x * W^T)
// The training shapes are skinny (M = B * sample_n = 50..60 rows, or K = T * rows ~ 1000 for the weight gradients), i.e. weight-
// streaming bound; 64x64x16 tiles with 4x4 register blocks keep enough CTAs in flight for those shapes.
#include <cstring>

#include "common.cuh"
#include "kernels.cuh"

namespace capb200 {

namespace {

constexpr int GT = 64, GK = 16;

template <int TA, int TB>
__global__ void __launch_bounds__(256) gemm_generic_kernel(int M, int N, int K, const float* __restrict__ A, long lda, const float* __restrict__ B,
                                                           long ldb, float* __restrict__ C, long ldc, int accumulate, const float* __restrict__ bias) {
    __shared__ float As[GK][GT + 1];
    __shared__ float Bs[GK][GT + 1];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += GK) {
        // 64 x 16 elements per operand tile, 4 per thread
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = tid + 256 * u;
            int mm, kk;
            if (TA) { mm = idx & 63; kk = idx >> 6; } else { kk = idx & 15; mm = idx >> 4; }       // contiguous index follows the storage order
            float v = 0.f;
            if (m0 + mm < M && k0 + kk < K) v = TA ? A[(long)(k0 + kk) * lda + m0 + mm] : A[(long)(m0 + mm) * lda + k0 + kk];
            As[kk][mm] = v;
            int nn, kb;
            if (TB) { kb = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kb = idx >> 6; }
            float w = 0.f;
            if (n0 + nn < N && k0 + kb < K) w = TB ? B[(long)(n0 + nn) * ldb + k0 + kb] : B[(long)(k0 + kb) * ldb + n0 + nn];
            Bs[kb][nn] = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + ty * 4 + i;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + tx * 4 + j;
            if (col >= N) continue;
            float v = acc[i][j];
            if (bias != nullptr) v += bias[col];
            float* c = C + (long)row * ldc + col;
            *c = accumulate ? (*c + v) : v;
        }
    }
}

// column sums: out[c] (+)= sum_r x[r, c]   (bias gradients)
__global__ void colsum_kernel(int rows, int cols, const float* __restrict__ x, long ld, float* __restrict__ out, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += x[(long)r * ld + c];
    out[c] = accumulate ? out[c] + s : s;
}

}  // namespace

int gemm_generic_launch(int ta, int tb, int M, int N, int K, const float* A, long lda, const float* B, long ldb, float* C, long ldc, int accumulate,
                        const float* bias, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    dim3 grid(cdiv(N, GT), cdiv(M, GT));
    if (!ta && !tb) gemm_generic_kernel<0, 0><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, bias);
    else if (!ta && tb) gemm_generic_kernel<0, 1><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, bias);
    else if (ta && !tb) gemm_generic_kernel<1, 0><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, bias);
    else gemm_generic_kernel<1, 1><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, bias);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int colsum_launch(int rows, int cols, const float* x, long ld, float* out, int accumulate, cudaStream_t st) {
    if (cols <= 0) return 0;
    colsum_kernel<<<cdiv(cols, 128), 128, 0, st>>>(rows, cols, x, ld, out, accumulate);
    CAPB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace capb200

// ---------------------------------------------------------------------------------------------------------------------
// Skinny split-K GEMM for the training step:  C[M,N] (+)= sum_s A_s[M,K_s] * op(B_s)  (+ bias[N] + row_bias[row / rpg, N])
// M = B * sample_n is 50..60 rows, so a conventional tiling leaves most SMs idle while one CTA walks a 3000-deep K; here the K-steps of
// all segments are split across blockIdx.y so ~2 CTAs per SM stream disjoint slices of the weights, and a second pass adds the partial
// sums in a fixed order (deterministic, no atomics).
// ---------------------------------------------------------------------------------------------------------------------
namespace capb200 {

namespace {

struct SkinnyParams {
    const float* A[3];
    const float* B[3];
    long lda[3], ldb[3];
    int K[3], tb[3];
    int nseg, M, N, ksteps_total, ksplit;
    float* out;          // C (ksplit == 1) or the partial buffer [ksplit][M][N]
    long ldo;
    const float* bias;
    const float* row_bias;
    long ld_rb;
    int rpg, accumulate;
};

__global__ void __launch_bounds__(256) gemm_skinny_kernel(const SkinnyParams p) {
    __shared__ float As[GK][GT + 1];
    __shared__ float Bs[GK][GT + 1];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * GT, m0 = blockIdx.z * GT;
    const int per = (p.ksteps_total + p.ksplit - 1) / p.ksplit;
    const int ks0 = blockIdx.y * per, ks1 = min(p.ksteps_total, ks0 + per);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    int seg = 0, seg_first = 0;
    for (int ks = ks0; ks < ks1; ++ks) {
        while (seg < p.nseg - 1 && ks >= seg_first + (p.K[seg] + GK - 1) / GK) { seg_first += (p.K[seg] + GK - 1) / GK; ++seg; }
        const int k0 = (ks - seg_first) * GK;
        const int K = p.K[seg];
        const float* A = p.A[seg];
        const float* B = p.B[seg];
        const long lda = p.lda[seg], ldb = p.ldb[seg];
        const int tb = p.tb[seg];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = tid + 256 * u;
            const int kk = idx & 15, mm = idx >> 4;
            As[kk][mm] = (m0 + mm < p.M && k0 + kk < K) ? A[(long)(m0 + mm) * lda + k0 + kk] : 0.f;
            int nn, kb;
            if (tb) { kb = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kb = idx >> 6; }
            float w = 0.f;
            if (n0 + nn < p.N && k0 + kb < K) w = tb ? B[(long)(n0 + nn) * ldb + k0 + kb] : B[(long)(k0 + kb) * ldb + n0 + nn];
            Bs[kb][nn] = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + ty * 4 + i;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + tx * 4 + j;
            if (col >= p.N) continue;
            float v = acc[i][j];
            if (p.ksplit == 1) {
                if (p.bias) v += p.bias[col];
                if (p.row_bias) v += p.row_bias[(long)(row / p.rpg) * p.ld_rb + col];
                float* c = p.out + (long)row * p.ldo + col;
                *c = p.accumulate ? (*c + v) : v;
            } else {
                p.out[((long)blockIdx.y * p.M + row) * p.N + col] = v;
            }
        }
    }
}

__global__ void skinny_reduce_kernel(int M, int N, int ksplit, const float* __restrict__ part, float* __restrict__ C, long ldc, const float* __restrict__ bias,
                                     const float* __restrict__ row_bias, long ld_rb, int rpg, int accumulate) {
    const long total = (long)M * N;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int row = (int)(i / N), col = (int)(i % N);
        float v = 0.f;
        for (int s = 0; s < ksplit; ++s) v += part[(long)s * total + i];
        if (bias) v += bias[col];
        if (row_bias) v += row_bias[(long)(row / rpg) * ld_rb + col];
        float* c = C + (long)row * ldc + col;
        *c = accumulate ? (*c + v) : v;
    }
}

}  // namespace

int gemm_skinny_launch(int M, int N, int nseg, const float* const* A, const long* lda, const float* const* B, const long* ldb, const int* K, const int* tb,
                       float* C, long ldc, const float* bias, const float* row_bias, long ld_rb, int rpg, int accumulate, float* scratch,
                       size_t scratch_floats, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    CAPB_REQUIRE(nseg >= 1 && nseg <= 3, "1..3 segments");
    SkinnyParams p;
    memset(&p, 0, sizeof(p));
    p.nseg = nseg; p.M = M; p.N = N;
    int ksteps = 0;
    for (int s = 0; s < nseg; ++s) {
        p.A[s] = A[s]; p.B[s] = B[s]; p.lda[s] = lda[s]; p.ldb[s] = ldb[s]; p.K[s] = K[s]; p.tb[s] = tb[s];
        ksteps += cdiv(K[s], GK);
    }
    p.ksteps_total = ksteps;
    const int tiles = cdiv(N, GT) * cdiv(M, GT);
    int ksplit = (2 * 148 + tiles - 1) / tiles;
    if (ksplit > ksteps / 4) ksplit = ksteps / 4;                  // at least 4 K-steps per CTA
    if (ksplit < 1) ksplit = 1;
    while (ksplit > 1 && (size_t)ksplit * M * N > scratch_floats) --ksplit;
    p.ksplit = ksplit;
    p.bias = bias; p.row_bias = row_bias; p.ld_rb = ld_rb; p.rpg = rpg < 1 ? 1 : rpg; p.accumulate = accumulate;
    if (ksplit == 1) { p.out = C; p.ldo = ldc; } else { p.out = scratch; p.ldo = N; }
    dim3 grid(cdiv(N, GT), ksplit, cdiv(M, GT));
    gemm_skinny_kernel<<<grid, 256, 0, st>>>(p);
    CAPB_CHECK_CUDA(cudaGetLastError());
    if (ksplit > 1) {
        long blocks = ((long)M * N + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        skinny_reduce_kernel<<<(int)blocks, 256, 0, st>>>(M, N, ksplit, scratch, C, ldc, bias, row_bias, ld_rb, p.rpg, accumulate);
        CAPB_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace capb200
