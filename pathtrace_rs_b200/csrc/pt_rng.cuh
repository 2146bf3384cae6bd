// pt_rng.cuh — per-pixel RNG of the render loop, device side.
//
// Replaces `Xoshiro256Plus::seed_from_u64(..)` + `rng.gen::<f32>()` as used by src/scene.rs:96-108
// (crates rand 0.8.5 / rand_core 0.6.3 / rand_xoshiro 0.6.0 — third-party, Cargo.lock:850-882):
//   seed_from_u64 : four SplitMix64 outputs fill the 256-bit state
//   next_u64      : xoshiro256+  (result = s0 + s3, then the xorshift/rotate state update)
//   gen::<f32>()  : top 24 bits of next_u32() (= high half of next_u64) scaled by 2^-24
// The state lives in 8 registers per lane for the whole life of a pixel, so every pixel consumes
// exactly the stream the reference would (same seed, same draw order).
#pragma once
#include <stdint.h>

namespace pt {

struct Rng {
    uint64_t s0, s1, s2, s3;
};

__device__ __forceinline__ uint64_t splitmix64_next(uint64_t& x) {
    x += 0x9e3779b97f4a7c15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ void rng_seed(Rng& r, uint64_t seed) {
    r.s0 = splitmix64_next(seed);
    r.s1 = splitmix64_next(seed);
    r.s2 = splitmix64_next(seed);
    r.s3 = splitmix64_next(seed);
}

__device__ __forceinline__ uint64_t rng_next_u64(Rng& r) {
    const uint64_t result = r.s0 + r.s3;
    const uint64_t t = r.s1 << 17;
    r.s2 ^= r.s0;
    r.s3 ^= r.s1;
    r.s1 ^= r.s2;
    r.s0 ^= r.s3;
    r.s2 ^= t;
    r.s3 = (r.s3 << 45) | (r.s3 >> 19);
    return result;
}

__device__ __forceinline__ float rng_f32(Rng& r) {
    const uint32_t hi = (uint32_t)(rng_next_u64(r) >> 32);
    return (float)(hi >> 8) * (1.0f / 16777216.0f);
}

// src/scene.rs:99-101
__device__ __forceinline__ uint64_t pixel_seed(uint32_t x, uint32_t y, uint32_t frame_num) {
    return ((uint64_t)x * 1973ULL + (uint64_t)y * 9277ULL + (uint64_t)frame_num * 26699ULL) | 1ULL;
}

}  // namespace pt
