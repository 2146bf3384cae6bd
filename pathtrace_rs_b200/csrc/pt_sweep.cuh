// pt_sweep.cuh — nearest-hit sweep of one ray (one lane) over the sphere SoA in shared memory.
//
// Semantic spec: `SpheresSoA::hit_scalar` (src/collision/spheres_soa.rs:105-155), i.e. for unit
// directions the same nearest hit as the live `HitableList::ray_hit` -> `Sphere::ray_hit`
// (src/collision/hitable_list.rs:40-56, src/collision/sphere.rs:29-66).
//
// Layout: "blocks" of 4 spheres, 64 B each: float4 X (cx0..3), Y (cy0..3), Z (cz0..3), R (r^2 0..3).
// A lane reads a block with four broadcast LDS.128 and tests it as two packed pairs with the
// sm_100 f32x2 pipe (FADD2 / FMUL2 / FFMA2, ray components broadcast from scalar registers):
//
//   co  = c - o                       3 FADD2        pre-filter, fused:
//   nb  = co . d                      FMUL2 + 2 FFMA2     lhs = nb*nb + r^2
//   rhs = co . co                     FMUL2 + 2 FFMA2     hit candidate  <=>  lhs > rhs   (disc > 0)
//   lhs = nb*nb + r^2                 1 FFMA2
//
// = 10 packed FP instructions per 2 (ray,sphere) tests, 16 flop per test.  Only when the
// pre-filter fires does the lane re-evaluate that one sphere with the reference's exact unfused
// expression order (spheres_soa.rs:116-129), so accepted hits and their `t` round like the oracle's.
// Padding spheres have centre = FLT_MAX, r^2 = 0 (spheres_soa.rs:53-61): rhs = +inf, never a candidate.
#pragma once
#include <stdint.h>
#include <float.h>

namespace pt {

typedef unsigned long long u64_t;

__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
    u64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<u64_t*>(&a)), "l"(*reinterpret_cast<u64_t*>(&b)), "l"(*reinterpret_cast<u64_t*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 f2_sub(float2 a, float2 b) {
    u64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<u64_t*>(&a)), "l"(*reinterpret_cast<u64_t*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
    u64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<u64_t*>(&a)), "l"(*reinterpret_cast<u64_t*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}

constexpr float kMinT = 0.001f;   // src/scene.rs:16
constexpr float kMaxT = FLT_MAX;  // src/scene.rs:15

// exact re-test of one sphere, reference expression order (spheres_soa.rs:116-129), unfused
__device__ __forceinline__ void sweep_exact(float cox, float coy, float coz, float r2, float dx, float dy, float dz,
                                            int index, float& hit_t, int& hit_index) {
    const float nb = (cox * dx + coy * dy) + coz * dz;
    const float c = ((cox * cox + coy * coy) + coz * coz) - r2;
    const float discriminant = nb * nb - c;
    if (discriminant > 0.0f) {
        const float discriminant_sqrt = sqrtf(discriminant);
        float t = nb - discriminant_sqrt;
        if (t < kMinT) t = nb + discriminant_sqrt;
        if (t > kMinT && t < hit_t) {
            hit_t = t;
            hit_index = index;
        }
    }
}

// one packed pair (spheres base_index, base_index+1)
__device__ __forceinline__ void sweep_pair(float2 cx, float2 cy, float2 cz, float2 r2, float ox, float oy, float oz,
                                           float dx, float dy, float dz, int base_index, float& hit_t, int& hit_index) {
    const float2 cox = f2_sub(cx, make_float2(ox, ox));
    const float2 coy = f2_sub(cy, make_float2(oy, oy));
    const float2 coz = f2_sub(cz, make_float2(oz, oz));
    const float2 nb = f2_fma(coz, make_float2(dz, dz), f2_fma(coy, make_float2(dy, dy), f2_mul(cox, make_float2(dx, dx))));
    const float2 rhs = f2_fma(coz, coz, f2_fma(coy, coy, f2_mul(cox, cox)));
    const float2 lhs = f2_fma(nb, nb, r2);
    if (lhs.x > rhs.x || lhs.y > rhs.y) {
        if (lhs.x > rhs.x) sweep_exact(cox.x, coy.x, coz.x, r2.x, dx, dy, dz, base_index, hit_t, hit_index);
        if (lhs.y > rhs.y) sweep_exact(cox.y, coy.y, coz.y, r2.y, dx, dy, dz, base_index + 1, hit_t, hit_index);
    }
}

// sweep blocks [0, n_blocks) of `blk` (shared memory); sphere index of block j, slot e is first_index + 4*j + e
template <int UNROLL>
__device__ __forceinline__ void sweep_blocks(const float4* __restrict__ blk, int n_blocks, int first_index, float ox, float oy,
                                             float oz, float dx, float dy, float dz, float& hit_t, int& hit_index) {
#pragma unroll UNROLL
    for (int j = 0; j < n_blocks; ++j) {
        const float4 X = blk[4 * j + 0], Y = blk[4 * j + 1], Z = blk[4 * j + 2], R = blk[4 * j + 3];
        const int base = first_index + 4 * j;
        sweep_pair(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y), make_float2(R.x, R.y), ox, oy, oz, dx, dy,
                   dz, base, hit_t, hit_index);
        sweep_pair(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w), make_float2(R.z, R.w), ox, oy, oz, dx, dy,
                   dz, base + 2, hit_t, hit_index);
    }
}

}  // namespace pt
