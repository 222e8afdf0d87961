// Probe: what does a kernel boundary cost on this GPU for a chain of small dependent kernels, and how much of it does programmatic dependent
// launch (griddepcontrol) hide?  Variants: plain stream launches, a CUDA graph of them, PDL launches, a graph captured from PDL launches.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_probe pdl_probe.cu     run: ./pdl_probe [work_iters]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <bool PDL>
__global__ void __launch_bounds__(256) step_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int iters) {
    if (PDL) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;");
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = in[i];
    for (int k = 0; k < iters; ++k) v = fmaf(v, 1.0000001f, 1e-7f);
    out[i] = v;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <bool PDL>
void enqueue(float* a, float* b, int n, int iters, int count, int grid, cudaStream_t st) {
    for (int k = 0; k < count; ++k) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = PDL ? 1 : 0;
        const float* in = (k & 1) ? b : a;
        float* out = (k & 1) ? a : b;
        CK(cudaLaunchKernelEx(&cfg, step_kernel<PDL>, in, out, n, iters));
    }
}

template <bool PDL>
float run(bool graph, float* a, float* b, int n, int iters, int count, int grid, cudaStream_t st) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    cudaGraphExec_t exec = nullptr;
    if (graph) {
        cudaGraph_t g;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        enqueue<PDL>(a, b, n, iters, count, grid, st);
        CK(cudaStreamEndCapture(st, &g));
        CK(cudaGraphInstantiate(&exec, g, 0));
        CK(cudaGraphDestroy(g));
    }
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaStreamSynchronize(st));
        CK(cudaEventRecord(e0, st));
        if (graph) CK(cudaGraphLaunch(exec, st)); else enqueue<PDL>(a, b, n, iters, count, grid, st);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    if (exec) CK(cudaGraphExecDestroy(exec));
    return best;
}

int main(int argc, char** argv) {
    const int count = 1000;
    cudaStream_t st; CK(cudaStreamCreate(&st));
    const int n = 148 * 256 * 4;
    float *a, *b; CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
    const int iters_list[] = {0, 2000, 8000};
    const int grids[] = {148, 592};
    for (int grid : grids)
        for (int iters : iters_list) {
            const int nn = grid * 256;
            const float p = run<false>(false, a, b, nn, iters, count, grid, st), pg = run<false>(true, a, b, nn, iters, count, grid, st);
            const float d = run<true>(false, a, b, nn, iters, count, grid, st), dg = run<true>(true, a, b, nn, iters, count, grid, st);
            printf("grid %4d CTAs, %5d fma/thread: per kernel  stream %.2f us | graph %.2f us | PDL stream %.2f us | PDL graph %.2f us\n", grid, iters, p * 1000 / count,
                   pg * 1000 / count, d * 1000 / count, dg * 1000 / count);
        }
    // correctness of the PDL chain: a += 1e-7-ish per step is not checked numerically here; the dependency is (k reads what k-1 wrote)
    return 0;
}
