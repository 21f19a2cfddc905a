// Philox4x32-10 (Salmon et al., SC'11) counter-based generator shared by the dropout kernels: counter = (group of four
// elements, stream offset), key = seed.  Layer_Dropout (Dropout.py:15-30) draws mask = Bernoulli(keep) / keep from it.
#pragma once
#include <cstdint>

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}

// keep flags of the four elements of group q: bit u set = element 4q + u survives (u uniform from 24 random bits: exact in fp32)
__device__ __forceinline__ uint32_t dropout_keep4(unsigned long long q, unsigned long long seed, unsigned long long offset, float keep) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    uint32_t bits = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) bits |= ((float)(rr[u] >> 8) * (1.f / 16777216.f) < keep) ? (1u << u) : 0u;
    return bits;
}
