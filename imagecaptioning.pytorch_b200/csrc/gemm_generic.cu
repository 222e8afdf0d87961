// Generic fp32 CUDA-core GEMM for the SCST backward pass:  C[M,N] = op(A)[M,K] * op(B)[K,N] (+ C if accumulate)
//
//   TA = 0: A stored [M, K] (pitch lda)      TA = 1: A stored [K, M]   (dW = dY^T * X reads dY this way)
//   TB = 0: B stored [K, N] (pitch ldb)      TB = 1: B stored [N, K]   (nn.Linear forward: x * W^T)
// The training shapes are skinny (M = B * sample_n = 50..60 rows, or K = T * rows ~ 1000 for the weight gradients), i.e. weight-
// streaming bound; 64x64x16 tiles with 4x4 register blocks keep enough CTAs in flight for those shapes.
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
