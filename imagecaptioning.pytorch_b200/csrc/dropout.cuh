// Replayable dropout: keep(seed, site, step, element) is a pure function (Philox4x32-10), so masks are never stored -- the backward
// pass (and the tests, through capb200_dropout_mask) re-evaluate them.
#pragma once
#include <cstdint>

namespace capb200 {

__device__ __forceinline__ uint32_t philox_word(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
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

// Seed salt.  Every kernel receives its seed by value, which a captured CUDA graph freezes; the fused SCST step is replayed from a graph with a
// different seed every step, so the effective seed is (seed XOR salt) with the salt in device memory: one copy per translation unit that
// includes this header (static), refreshed in stream order before every training step by dropout_salt_set_all (kernels.cuh).  Eager steps
// upload salt 0, graph replays upload captured_seed XOR requested_seed: the masks depend on the requested seed only, whichever way the step ran.
static __device__ unsigned long long g_capb_seed_salt = 0ull;
#define CAPB_DEFINE_SALT_SETTER(fn)                                                                                                   \
    int fn(unsigned long long salt, cudaStream_t st) {                                                                                \
        return cudaMemcpyToSymbolAsync(g_capb_seed_salt, &salt, sizeof(salt), 0, cudaMemcpyHostToDevice, st) == cudaSuccess ? 0 : 1;  \
    }

// 0 (dropped) or 1/(1-p) (kept): inverted dropout like nn.Dropout
__device__ __forceinline__ float drop_scale(unsigned long long seed, uint32_t site, uint32_t step, uint32_t idx, float p) {
    if (p <= 0.f) return 1.f;
    seed ^= g_capb_seed_salt;
    const uint32_t bits = philox_word(idx, site, step, 0x5C57u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float u = (float)(bits >> 8) * (1.0f / 16777216.0f);        // [0, 1)
    return u < p ? 0.f : 1.0f / (1.0f - p);
}

}  // namespace capb200
