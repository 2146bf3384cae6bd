// pt_sweep.cuh — nearest-hit sweep of one ray (one lane) over the whole sphere SoA, brute force.
//
// Semantic spec: `SpheresSoA::hit_scalar` (src/collision/spheres_soa.rs:105-155), i.e. for unit
// directions the same nearest hit as the live `HitableList::ray_hit` -> `Sphere::ray_hit`
// (src/collision/hitable_list.rs:40-56, src/collision/sphere.rs:29-66).
//
// Two-stage test.  Stage 1 is a conservative pre-filter over ALL spheres in packed FP32 (FFMA2, two spheres per
// instruction, the ray broadcast from scalar registers): the discriminant expanded around the ray,
//
//   disc = (c.d - o.d)^2 + 2 c.o + (r^2 - |c|^2) - |o|^2 ,   A = c.d - o.d (3 FFMA2), B = 2 c.o + k (3 FFMA2),
//   L = A*A + B (1 FFMA2),   candidate  <=>  L > |o|^2 (1 - 2^-19)
//
// = 7 packed instructions per 2 (ray, sphere) tests on the "pre-filter image" X, Y, Z, K (k = r^2 - |c|^2 + slack) of a
// block of 4 spheres.  Stage 2 re-tests only the flagged spheres with the reference's exact unfused expression
// (sweep_exact = spheres_soa.rs:116-129) on the exact blocks X, Y, Z, R^2, so accepted hits and their `t` round like the
// oracle's.  Two operand paths for stage 1: sweep_const reads the image through the constant bank into uniform
// registers (scenes <= 4000 spheres), sweep_expanded reads it with broadcast LDS.128 from a TMA-staged shared-memory
// tile (resident up to ~13 k spheres, streamed beyond).  Padding: k = -3e38 in the image (never a candidate),
// centre = FLT_MAX, r^2 = 0 in the exact blocks (spheres_soa.rs:53-61).
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
#define PT_CONST_GROUP 4
#endif
constexpr int kConstGroupBlocks = PT_CONST_GROUP;          // blocks (of 4 spheres) per pre-filter branch / queue entry
constexpr int kEntryMaskBits = 4 * kConstGroupBlocks;       // one flag bit per sphere of the group
constexpr int kMaxConstBlocks = 1000;                      // 4000 spheres * 16 B = 64 000 B of the 64 KB constant bank
constexpr int kMaxConstSpheres = 4 * kMaxConstBlocks;
__constant__ float4 c_prefilter[4 * kMaxConstBlocks];      // per block: X(cx0..3) Y Z K(k0..3)

// n_groups = n_blocks / 2; both the constant image and the shared-memory copy are padded to whole groups.
// Candidate queue: one 32-bit entry per flagged group = (first block << kEntryMaskBits) | one flag bit per sphere, kQueueCap entries per lane in
// shared memory ([entry][thread] layout, conflict-free).  A full queue (rare) tests the group on the spot.
constexpr int kQueueCap = 12;

template <int MASK_BITS>
__device__ __forceinline__ void sweep_resolve_entry(const float4* __restrict__ blk, uint32_t entry, float ox, float oy, float oz, float dx,
                                                    float dy, float dz, float& hit_t, int& hit_index) {
    const int base = (int)(entry >> MASK_BITS) * 4;
    uint32_t mask = entry & ((1u << MASK_BITS) - 1u);
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
                sweep_resolve_entry<kEntryMaskBits>(blk, entry, ox, oy, oz, dx, dy, dz, hit_t, hit_index);
            }
        }
    }
    // resolve: every lane walks its own queue, all lanes in parallel
#pragma unroll 1
    for (int i = 0; i < cnt; ++i) sweep_resolve_entry<kEntryMaskBits>(blk, q[i * kSweepThreads], ox, oy, oz, dx, dy, dz, hit_t, hit_index);
}

// =====================================================================================================
// Expanded pre-filter with shared-memory operands (scenes beyond the constant bank: resident and streamed kernels).
//
// Same algebra as sweep_const (A = c.d - o.d, B = 2 c.o + k, L = A*A + B, candidate <=> L > |o|^2 (1 - 2^-19)) but the
// sphere pairs come from the pre-filter image staged in shared memory (LDS.128 broadcast): 7 packed instructions per
// 2 tests of the shape FFMA2 Rpair, Rpair(spheres), Rscalar(ray), Rpair — against 11 for the sphere-relative form of
// sweep_blocks, whose packed instructions mostly read two or three register pairs (profiles/probe_forms_r1.txt).
// `pf` holds blocks [first_block, first_block + n_blocks) of the image (n_blocks a multiple of the group size);
// flagged groups go to the lane's queue as (absolute block << kLdsMaskBits | flags) and are re-tested by the
// caller with the reference's exact expression against `exact` (global memory for the streamed kernel).
// =====================================================================================================
constexpr int kLdsGroupBlocks = 2;                 // blocks per pre-filter branch / queue entry of the LDS sweep
constexpr int kLdsMaskBits = 4 * kLdsGroupBlocks;  // entry = absolute block << 8 | flags: up to 2^24 blocks
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
template <bool PIPE>
__device__ __forceinline__ void sweep_expanded(const float4* __restrict__ pf, int n_blocks, int first_block, const float4* __restrict__ exact,
                                               uint32_t* __restrict__ q, int& cnt, float ox, float oy, float oz, float dx, float dy, float dz,
                                               float nod, float o2x, float o2y, float o2z, float oo, float& hit_t, int& hit_index) {
    // pin the per-ray operands (and the trip count) in registers: without this ptxas rematerialises them — 9 scalar FP
    // instructions and a constant-bank reload of n_blocks — in every trip
    asm volatile("" : "+f"(nod), "+f"(o2x), "+f"(o2y), "+f"(o2z), "+f"(oo), "+r"(n_blocks));
    // explicit shared-window address, advanced by one group per trip (keeps the loop's address arithmetic to one add)
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(pf);
    asm volatile("" : "+r"(addr));
    // PIPE: software pipeline, one block deep — block b+1 is in flight while block b is tested (LDS latency off the
    // FFMA2 chain) at the price of 16 registers; pays when the kernel is at 2 CTAs/SM anyway (streamed tiles)
    const uint32_t last = addr + 64u * (uint32_t)(n_blocks > 0 ? n_blocks - 1 : 0);
    float4 cX, cY, cZ, cK;
    if (PIPE) {
        cX = lds128(addr); cY = lds128(addr + 16u); cZ = lds128(addr + 32u); cK = lds128(addr + 48u);
    }
    const uint32_t base = addr, end = addr + 64u * (uint32_t)n_blocks;
#pragma unroll 1
    for (; addr < end; addr += 64u * kLdsGroupBlocks) {
        // (a warp-uniform single-branch form — vote.any on the group's flag — measured 1.5 % slower: the vote serialises
        // the back-branch behind the whole FFMA2 chain, whereas this back-branch depends on the address alone)
        float2 L[2 * kLdsGroupBlocks];
        bool any;
        {
#pragma unroll
            for (int g = 0; g < kLdsGroupBlocks; ++g) {
                float4 X, Y, Z, K;
                if (PIPE) {
                    X = cX; Y = cY; Z = cZ; K = cK;
                    const uint32_t na = min(addr + 64u * (g + 1), last);  // next block (clamped at the tile's last block)
                    cX = lds128(na); cY = lds128(na + 16u); cZ = lds128(na + 32u); cK = lds128(na + 48u);
                } else {
                    X = lds128(addr + 64u * g); Y = lds128(addr + 64u * g + 16u); Z = lds128(addr + 64u * g + 32u); K = lds128(addr + 64u * g + 48u);
                }
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
            any = false;
#pragma unroll
            for (int p = 0; p < 2 * kLdsGroupBlocks; ++p) any = any | (L[p].x > oo) | (L[p].y > oo);
        }
        if (any) {
            uint32_t mask = 0u;
#pragma unroll
            for (int p = 0; p < 2 * kLdsGroupBlocks; ++p) mask |= (L[p].x > oo ? 1u << (2 * p) : 0u) | (L[p].y > oo ? 2u << (2 * p) : 0u);
            const uint32_t entry = (((uint32_t)first_block + ((addr - base) >> 6)) << kLdsMaskBits) | mask;
            if (cnt < kQueueCap) {
                q[cnt * kSweepThreads] = entry;
                cnt += 1;
            } else {  // queue full (rare): test this group now; sweep_exact's tie rule makes the visiting order irrelevant
                sweep_resolve_entry<kLdsMaskBits>(exact, entry, ox, oy, oz, dx, dy, dz, hit_t, hit_index);
            }
        }
    }
}

// drain the lane's queue: exact re-test of every flagged sphere (all lanes in parallel)
__device__ __forceinline__ void sweep_drain(const float4* __restrict__ exact, const uint32_t* __restrict__ q, int& cnt, float ox, float oy, float oz,
                                            float dx, float dy, float dz, float& hit_t, int& hit_index) {
#pragma unroll 1
    for (int i = 0; i < cnt; ++i) sweep_resolve_entry<kLdsMaskBits>(exact, q[i * kSweepThreads], ox, oy, oz, dx, dy, dz, hit_t, hit_index);
    cnt = 0;
}

}  // namespace pt
