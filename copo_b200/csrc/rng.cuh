// Counter-based standard normals shared by the sampling kernels (heads.cu, mlp_tc.cu): a hash of (seed, step, row)
// feeds Box-Muller.  Not part of any parity claim - tests inject their own eps where exact values matter.
#pragma once
#include <stdint.h>

namespace b2c {

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
// two standard normals from a counter (Box-Muller on two hashed 24-bit uniforms)
__device__ __forceinline__ void normal2(uint32_t seed, uint32_t ctr_hi, uint32_t ctr_lo, float& n0, float& n1) {
    uint32_t a = hash32(seed ^ hash32(ctr_hi * 0x9E3779B1u + 0x85EBCA77u) ^ (ctr_lo * 0xC2B2AE3Du));
    uint32_t b = hash32(a + 0x27D4EB2Fu);
    float u0 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float u1 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float r = sqrtf(-2.0f * logf(u0));
    float s, c;
    sincosf(6.28318530717958647692f * u1, &s, &c);
    n0 = r * c; n1 = r * s;
}

}  // namespace b2c
