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
// = 10 packed FP instructions per 2 (ray,sphere) tests, 16 flop per test.  A group of
// 4*GROUP spheres shares ONE branch (the compares are OR-ed into a single predicate), so the common path
// is straight-line code with 2*GROUP independent dependency chains.  Only when the pre-filter fires does
// the lane re-evaluate the flagged spheres with the reference's exact unfused expression order
// (spheres_soa.rs:116-129), so accepted hits and their `t` round like the oracle's.
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
        // strict `<` in ascending index order (spheres_soa.rs:126) == lowest index among equal t; written so that the
        // result does not depend on the order candidates are visited in
        if (t > kMinT && (t < hit_t || (t == hit_t && index < hit_index))) {
            hit_t = t;
            hit_index = index;
        }
    }
}

// pre-filter of one packed pair: returns lhs - rhs ordering inputs (kept live only until the group's branch)
struct PairTest {
    float2 lhs, rhs;
};
__device__ __forceinline__ PairTest sweep_pair(float2 cx, float2 cy, float2 cz, float2 r2, float ox, float oy, float oz, float dx,
                                               float dy, float dz) {
    const float2 cox = f2_sub(cx, make_float2(ox, ox));
    const float2 coy = f2_sub(cy, make_float2(oy, oy));
    const float2 coz = f2_sub(cz, make_float2(oz, oz));
    const float2 nb = f2_fma(coz, make_float2(dz, dz), f2_fma(coy, make_float2(dy, dy), f2_mul(cox, make_float2(dx, dx))));
    PairTest t;
    // conservative margin: the fused pre-filter and the unfused exact test round differently (a few ulp of |co|^2);
    // shrinking rhs by 2^-19 keeps every hit the exact test would accept flagged (11th packed instruction of the pair)
    const float2 rhs = f2_fma(coz, coz, f2_fma(coy, coy, f2_mul(cox, cox)));
    t.rhs = f2_mul(rhs, make_float2(1.0f - 1.9073486328125e-06f, 1.0f - 1.9073486328125e-06f));
    t.lhs = f2_fma(nb, nb, r2);
    return t;
}

// rare path: exact re-test of the spheres of block `j` whose bit is set in `mask` (bit e = slot e), ascending
// index order so that ties keep the lowest index like the scalar loop (spheres_soa.rs:126-129)
__device__ __forceinline__ void sweep_candidates(const float4* __restrict__ blk, int j, int first_index, unsigned mask, float ox,
                                                 float oy, float oz, float dx, float dy, float dz, float& hit_t, int& hit_index) {
    const float* bf = reinterpret_cast<const float*>(blk + 4 * j);
#pragma unroll 1
    while (mask != 0u) {
        const int e = __ffs(mask) - 1;
        mask &= mask - 1u;
        const float cox = bf[e] - ox, coy = bf[4 + e] - oy, coz = bf[8 + e] - oz;
        sweep_exact(cox, coy, coz, bf[12 + e], dx, dy, dz, first_index + 4 * j + e, hit_t, hit_index);
    }
}

// sweep blocks [0, n_blocks) of `blk` (shared memory); sphere index of block j, slot e is first_index + 4*j + e.
// GROUP blocks (4*GROUP spheres, 2*GROUP independent packed chains) are tested per trip with ONE branch.
template <int GROUP>
__device__ __forceinline__ void sweep_blocks(const float4* __restrict__ blk, int n_blocks, int first_index, float ox, float oy,
                                             float oz, float dx, float dy, float dz, float& hit_t, int& hit_index) {
    int j = 0;
#pragma unroll 1
    for (; j + GROUP <= n_blocks; j += GROUP) {
        PairTest t[2 * GROUP];
#pragma unroll
        for (int g = 0; g < GROUP; ++g) {
            const float4 X = blk[4 * (j + g) + 0], Y = blk[4 * (j + g) + 1], Z = blk[4 * (j + g) + 2], R = blk[4 * (j + g) + 3];
            t[2 * g] = sweep_pair(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y), make_float2(R.x, R.y), ox, oy,
                                  oz, dx, dy, dz);
            t[2 * g + 1] = sweep_pair(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w), make_float2(R.z, R.w), ox,
                                      oy, oz, dx, dy, dz);
        }
        bool any = false;
#pragma unroll
        for (int q = 0; q < 2 * GROUP; ++q) any = any | (t[q].lhs.x > t[q].rhs.x) | (t[q].lhs.y > t[q].rhs.y);
        if (any) {
#pragma unroll
            for (int g = 0; g < GROUP; ++g) {
                const unsigned mask = (t[2 * g].lhs.x > t[2 * g].rhs.x ? 1u : 0u) | (t[2 * g].lhs.y > t[2 * g].rhs.y ? 2u : 0u) |
                                      (t[2 * g + 1].lhs.x > t[2 * g + 1].rhs.x ? 4u : 0u) | (t[2 * g + 1].lhs.y > t[2 * g + 1].rhs.y ? 8u : 0u);
                sweep_candidates(blk, j + g, first_index, mask, ox, oy, oz, dx, dy, dz, hit_t, hit_index);
            }
        }
    }
#pragma unroll 1
    for (; j < n_blocks; ++j) {  // ragged tail: one block at a time
        const float4 X = blk[4 * j + 0], Y = blk[4 * j + 1], Z = blk[4 * j + 2], R = blk[4 * j + 3];
        const PairTest a = sweep_pair(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y), make_float2(R.x, R.y), ox, oy, oz, dx, dy, dz);
        const PairTest b = sweep_pair(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w), make_float2(R.z, R.w), ox, oy, oz, dx, dy, dz);
        const unsigned mask = (a.lhs.x > a.rhs.x ? 1u : 0u) | (a.lhs.y > a.rhs.y ? 2u : 0u) | (b.lhs.x > b.rhs.x ? 4u : 0u) | (b.lhs.y > b.rhs.y ? 8u : 0u);
        if (mask) sweep_candidates(blk, j, first_index, mask, ox, oy, oz, dx, dy, dz, hit_t, hit_index);
    }
}


// =====================================================================================================
// Constant-bank sweep (scenes of up to kMaxConstSpheres spheres — every preset of the reference).
//
// Measured on B200 (tools/probe_forms.cu, profiles/probe_forms_r1.txt): a packed FP32 instruction costs
// ~2.1 clk/warp when it reads one register pair, ~2.5 with two, >3 with two pairs + a scalar — the register
// file, not the FMA pipe, bounds the LDS form above at ~52 % of peak.  Sphere data is warp-uniform, so here it
// is read through the uniform datapath (LDCU from the constant bank -> uniform registers) and every packed
// instruction has the shape  FFMA2 Rpair, Rscalar(ray), URpair(2 spheres), Rpair|Rscalar : one pair read.
// To make every operation of that shape the discriminant is expanded around the ray instead of the sphere:
//
//   disc = (c.d - o.d)^2 + 2 c.o + (r^2 - |c|^2) - |o|^2
//   A = c.d - o.d            3 FFMA2   (ray: dx,dy,dz, -o.d)
//   B = 2 c.o + k            3 FFMA2   (ray: 2ox,2oy,2oz; sphere: k = r^2 - |c|^2 + slack)
//   L = A*A + B              1 FFMA2   candidate  <=>  L > |o|^2 (1 - 2^-19)
//
// = 7 packed instructions per 2 tests (14 flop executed per test; the algorithmic count stays 16).
// The expansion cancels badly (|c|^2 against r^2), so it is used ONLY as a conservative pre-filter: `slack` =
// 2^-19 (|c|^2 + r^2) on the sphere side and the 2^-19 relative margin on the ray side dominate the rounding
// error of both this form and the reference's (32 ulp of the largest terms), so every hit the exact test would
// accept is flagged.
//
// The sweep itself is BRANCH-FREE: a group of 8 spheres (2 blocks) reduces its 8 tests with FMNMX3 to one compare
// that sets one bit of a per-lane flag word (32 groups = 256 spheres per word, words live in registers).  After
// the whole sweep the lanes walk their set bits in parallel: the flagged group is re-filtered from the shared-
// memory copy (sweep_pair, LDS form) and the surviving spheres are re-tested with the reference's exact unfused
// expression (sweep_exact).  Divergence costs a few trips of one SIMD loop per sweep instead of a serialized
// branch per (lane, sphere) event, and the hot loop is straight-line code the compiler keeps in uniform registers.
// =====================================================================================================
constexpr int kSweepThreads = 256;  // == kCtaThreads (queue stride)
#ifndef PT_CONST_GROUP
#define PT_CONST_GROUP 2
#endif
constexpr int kConstGroupBlocks = PT_CONST_GROUP;          // blocks (of 4 spheres) per pre-filter branch / queue entry
constexpr int kEntryMaskBits = 4 * kConstGroupBlocks;       // one flag bit per sphere of the group
constexpr int kMaxConstBlocks = 1000;                      // 4000 spheres * 16 B = 64 000 B of the 64 KB constant bank
constexpr int kMaxConstSpheres = 4 * kMaxConstBlocks;
__constant__ float4 c_prefilter[4 * kMaxConstBlocks];      // per block: X(cx0..3) Y Z K(k0..3)

// re-filter + exact test of one flagged group (blocks j, j+1 of the shared-memory copy)
__device__ __forceinline__ void sweep_resolve_group(const float4* __restrict__ blk, int j, float ox, float oy, float oz, float dx,
                                                    float dy, float dz, float& hit_t, int& hit_index) {
#pragma unroll
    for (int g = 0; g < kConstGroupBlocks; ++g) {
        const float4 X = blk[4 * (j + g) + 0], Y = blk[4 * (j + g) + 1], Z = blk[4 * (j + g) + 2], R = blk[4 * (j + g) + 3];
        const PairTest a = sweep_pair(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y), make_float2(R.x, R.y), ox, oy, oz, dx, dy, dz);
        const PairTest b = sweep_pair(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w), make_float2(R.z, R.w), ox, oy, oz, dx, dy, dz);
        const unsigned mask = (a.lhs.x > a.rhs.x ? 1u : 0u) | (a.lhs.y > a.rhs.y ? 2u : 0u) | (b.lhs.x > b.rhs.x ? 4u : 0u) | (b.lhs.y > b.rhs.y ? 8u : 0u);
        if (mask) sweep_candidates(blk, j + g, 0, mask, ox, oy, oz, dx, dy, dz, hit_t, hit_index);
    }
}

// n_groups = n_blocks / 2; both the constant image and the shared-memory copy are padded to whole groups.
// Candidate queue: one 32-bit entry per flagged group = (first block << kEntryMaskBits) | one flag bit per sphere, kQueueCap entries per lane in
// shared memory ([entry][thread] layout, conflict-free).  A full queue (rare) tests the group on the spot.
constexpr int kQueueCap = 12;

__device__ __forceinline__ void sweep_resolve_entry(const float4* __restrict__ blk, uint32_t entry, float ox, float oy, float oz, float dx,
                                                    float dy, float dz, float& hit_t, int& hit_index) {
    const int base = (int)(entry >> kEntryMaskBits) * 4;
    uint32_t mask = entry & ((1u << kEntryMaskBits) - 1u);
#pragma unroll 1
    while (mask != 0u) {
        const int index = base + __ffs(mask) - 1;
        mask &= mask - 1u;
        const float* bf = reinterpret_cast<const float*>(blk) + (index >> 2) * 16 + (index & 3);
        sweep_exact(bf[0] - ox, bf[4] - oy, bf[8] - oz, bf[12], dx, dy, dz, index, hit_t, hit_index);
    }
}

template <int WORDS>
__device__ __forceinline__ void sweep_const(int n_groups, const float4* __restrict__ blk, const float2* __restrict__ ksm, uint32_t* __restrict__ q,
                                            float ox, float oy, float oz, float dx, float dy, float dz, float& hit_t, int& hit_index) {
    const int n_blocks_padded = n_groups * kConstGroupBlocks;
    const float nod = -((ox * dx + oy * dy) + oz * dz);
    const float o2x = ox + ox, o2y = oy + oy, o2z = oz + oz;
    const float oo = ((ox * ox + oy * oy) + oz * oz) * (1.0f - 1.9073486328125e-06f);
    int cnt = 0;
#pragma unroll 1
    for (int j = 0; j < n_blocks_padded; j += kConstGroupBlocks) {
        float2 L[2 * kConstGroupBlocks];
#pragma unroll
        for (int g = 0; g < kConstGroupBlocks; ++g) {
            const float4 X = c_prefilter[4 * (j + g) + 0], Y = c_prefilter[4 * (j + g) + 1], Z = c_prefilter[4 * (j + g) + 2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 cx = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                const float2 cy = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                const float2 cz = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
                const float2 k = ksm[2 * (j + g) + h];  // LDS.64 broadcast: k in a vector pair keeps every FFMA2 at one uniform operand
                const float2 A = f2_fma(cz, make_float2(dz, dz), f2_fma(cy, make_float2(dy, dy), f2_fma(cx, make_float2(dx, dx), make_float2(nod, nod))));
                const float2 B = f2_fma(cz, make_float2(o2z, o2z), f2_fma(cy, make_float2(o2y, o2y), f2_fma(cx, make_float2(o2x, o2x), k)));
                L[2 * g + h] = f2_fma(A, A, B);
            }
        }
        bool any = false;
#pragma unroll
        for (int p = 0; p < 2 * kConstGroupBlocks; ++p) any = any | (L[p].x > oo) | (L[p].y > oo);
        if (any) {
            uint32_t mask = 0u;
#pragma unroll
            for (int p = 0; p < 2 * kConstGroupBlocks; ++p) mask |= (L[p].x > oo ? 1u << (2 * p) : 0u) | (L[p].y > oo ? 2u << (2 * p) : 0u);
            const uint32_t entry = ((uint32_t)j << kEntryMaskBits) | mask;
            if (cnt < kQueueCap) {
                q[cnt * kSweepThreads] = entry;
                cnt += 1;
            } else {  // queue full (rare): test this group now; sweep_exact's tie rule makes the visiting order irrelevant
                sweep_resolve_entry(blk, entry, ox, oy, oz, dx, dy, dz, hit_t, hit_index);
            }
        }
    }
    // resolve: every lane walks its own queue, all lanes in parallel
#pragma unroll 1
    for (int i = 0; i < cnt; ++i) sweep_resolve_entry(blk, q[i * kSweepThreads], ox, oy, oz, dx, dy, dz, hit_t, hit_index);
}

// =====================================================================================================
// Expanded pre-filter with shared-memory operands (scenes beyond the constant bank: resident and streamed kernels).
//
// Same algebra as sweep_const (A = c.d - o.d, B = 2 c.o + k, L = A*A + B, candidate <=> L > |o|^2 (1 - 2^-19)) but the
// sphere pairs come from the pre-filter image staged in shared memory (LDS.128 broadcast): 7 packed instructions per
// 2 tests of the shape FFMA2 Rpair, Rpair(spheres), Rscalar(ray), Rpair — against 11 for the sphere-relative form of
// sweep_blocks, whose packed instructions mostly read two or three register pairs (profiles/probe_forms_r1.txt).
// `pf` holds blocks [first_block, first_block + n_blocks) of the image (n_blocks a multiple of the group size);
// flagged groups go to the lane's queue as (absolute block << kEntryMaskBits | flags) and are re-tested by the
// caller with the reference's exact expression against `exact` (global memory for the streamed kernel).
// =====================================================================================================
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sweep_expanded(const float4* __restrict__ pf, int n_blocks, int first_block, const float4* __restrict__ exact,
                                               uint32_t* __restrict__ q, int& cnt, float ox, float oy, float oz, float dx, float dy, float dz,
                                               float nod, float o2x, float o2y, float o2z, float oo, float& hit_t, int& hit_index) {
    // pin the per-ray operands in registers: without this ptxas rematerialises them (9 scalar FP instructions) in every trip
    asm volatile("" : "+f"(nod), "+f"(o2x), "+f"(o2y), "+f"(o2z), "+f"(oo));
    // explicit shared-window address, advanced by one group per trip (keeps the loop's address arithmetic to one add)
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(pf);
    asm volatile("" : "+r"(addr));
#pragma unroll 1
    for (int j = 0; j < n_blocks; j += kConstGroupBlocks, addr += 64u * kConstGroupBlocks) {
        float2 L[2 * kConstGroupBlocks];
#pragma unroll
        for (int g = 0; g < kConstGroupBlocks; ++g) {
            const float4 X = lds128(addr + 64u * g), Y = lds128(addr + 64u * g + 16u), Z = lds128(addr + 64u * g + 32u), K = lds128(addr + 64u * g + 48u);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 cx = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                const float2 cy = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                const float2 cz = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
                const float2 k = h ? make_float2(K.z, K.w) : make_float2(K.x, K.y);
                const float2 A = f2_fma(cz, make_float2(dz, dz), f2_fma(cy, make_float2(dy, dy), f2_fma(cx, make_float2(dx, dx), make_float2(nod, nod))));
                const float2 B = f2_fma(cz, make_float2(o2z, o2z), f2_fma(cy, make_float2(o2y, o2y), f2_fma(cx, make_float2(o2x, o2x), k)));
                L[2 * g + h] = f2_fma(A, A, B);
            }
        }
        bool any = false;
#pragma unroll
        for (int p = 0; p < 2 * kConstGroupBlocks; ++p) any = any | (L[p].x > oo) | (L[p].y > oo);
        if (any) {
            uint32_t mask = 0u;
#pragma unroll
            for (int p = 0; p < 2 * kConstGroupBlocks; ++p) mask |= (L[p].x > oo ? 1u << (2 * p) : 0u) | (L[p].y > oo ? 2u << (2 * p) : 0u);
            const uint32_t entry = ((uint32_t)(first_block + j) << kEntryMaskBits) | mask;
            if (cnt < kQueueCap) {
                q[cnt * kSweepThreads] = entry;
                cnt += 1;
            } else {  // queue full (rare): test this group now; sweep_exact's tie rule makes the visiting order irrelevant
                sweep_resolve_entry(exact, entry, ox, oy, oz, dx, dy, dz, hit_t, hit_index);
            }
        }
    }
}

// drain the lane's queue: exact re-test of every flagged sphere (all lanes in parallel)
__device__ __forceinline__ void sweep_drain(const float4* __restrict__ exact, const uint32_t* __restrict__ q, int& cnt, float ox, float oy, float oz,
                                            float dx, float dy, float dz, float& hit_t, int& hit_index) {
#pragma unroll 1
    for (int i = 0; i < cnt; ++i) sweep_resolve_entry(exact, q[i * kSweepThreads], ox, oy, oz, dx, dy, dz, hit_t, hit_index);
    cnt = 0;
}

}  // namespace pt
