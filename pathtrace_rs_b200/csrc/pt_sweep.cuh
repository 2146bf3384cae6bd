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
//   L = A*A + B (1 FFMA2),   candidate  <=>  L > |o|^2 (1 - 2^-18)
//
// = 7 packed instructions per 2 (ray, sphere) tests on the "pre-filter image" X, Y, Z, K (k = r^2 - |c|^2 + slack) of a
// block of 4 spheres.  Stage 2 re-tests only the flagged spheres with the reference's exact unfused expression
// (sweep_exact = spheres_soa.rs:116-129) on the exact blocks X, Y, Z, R^2, so accepted hits and their `t` round like the
// oracle's.  Stage 1 (sweep_expanded) reads the image with broadcast LDS.128 from a TMA-staged shared-memory tile
// (resident up to ~13 k spheres, streamed beyond).  Padding: k = -3e38 in the image (never a candidate),
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

// groups per trip of the pre-filter loop.  Two in the streamed kernel (all sweep, registers to spare: cfg5 +5.6 %); one in
// the resident kernel, where the longer body costs more around the loop than the saved back-branch gives (cfg2 -1.4 %)
#ifndef PT_GROUP_UNROLL_STREAMED
#define PT_GROUP_UNROLL_STREAMED 2
#endif
#ifndef PT_GROUP_UNROLL_RESIDENT
#define PT_GROUP_UNROLL_RESIDENT 1
#endif

constexpr float kMinT = 0.001f;   // src/scene.rs:16
constexpr float kMaxT = FLT_MAX;  // src/scene.rs:15

// Slack of the conservative pre-filter.  In exact arithmetic L - |o|^2 equals the reference's discriminant
// D = (co.d)^2 - |co|^2 + r^2; in f32 the expanded evaluation (7 fused steps on terms as large as (|c| + |o|)^2, the
// unfused o.d and |o|^2) and the reference's own sphere-relative evaluation differ from D by at most about
//   2^-24 (29 |c|^2 + 31 |o|^2 + 7 r^2)      (every rounding taken at its worst and with the same sign),
// i.e. 2^-19 (0.9 |c|^2 + 0.97 |o|^2 + 0.22 r^2).  The filter adds kSlack (|c|^2 + r^2) on the sphere side (baked into K
// by the host, rounded up) and kSlack |o|^2 on the ray side: 2^-18 is twice that bound, so a sphere the exact expression
// accepts is always a candidate.  Domain: finite geometry, |camera| <= 1e17; a sphere with |c|^2 + r^2 >= 1e24 is stored
// with K = +inf (always a candidate, decided by the exact test alone).  Checked per ray, not assumed: pt_debug_hits mode 0
// against mode 1 (tests/test_gpu_hits.py), including origins and centres swept to 1e6.
constexpr float kSlack = 3.814697265625e-06f;  // 2^-18
constexpr double kSlackSphere = 3.814697265625e-06;

// exact re-test of one sphere, reference expression order (spheres_soa.rs:116-129), unfused
// `order` (may be null): the scene's spheres are stored in a spatial order (ptgpu.cu), order[i] = the sphere's position
// in the caller's list; equal-t ties go to the lower ORIGINAL position, as the reference's in-order walk decides them
__device__ __noinline__ bool tie_goes_to(const uint32_t* __restrict__ order, int index, int hit_index) {
    if (order == nullptr || hit_index < 0) return index < hit_index;
    return __ldg(order + index) < __ldg(order + hit_index);
}
// ORDERED: the kernel runs on a spatially ordered scene (resident kernel); otherwise stored order = list order and the
// tie rule is the plain index comparison
template <bool ORDERED = false>
__device__ __forceinline__ void sweep_exact(float cox, float coy, float coz, float r2, float dx, float dy, float dz,
                                            int index, float& hit_t, int& hit_index, const uint32_t* __restrict__ order = nullptr) {
    const float nb = (cox * dx + coy * dy) + coz * dz;
    const float c = ((cox * cox + coy * coy) + coz * coz) - r2;
    const float discriminant = nb * nb - c;
    if (discriminant > 0.0f) {
        const float discriminant_sqrt = sqrtf(discriminant);
        float t = nb - discriminant_sqrt;
        if (t < kMinT) t = nb + discriminant_sqrt;
        // strict `<` in ascending index order (spheres_soa.rs:126) == lowest index among equal t; written so that the
        // result does not depend on the order candidates are visited in
        if (t > kMinT && (t < hit_t || (t == hit_t && (ORDERED ? tie_goes_to(order, index, hit_index) : index < hit_index)))) {
            hit_t = t;
            hit_index = index;
        }
    }
}

// ---- Hitable::MovingSphere (src/collision/moving_sphere.rs) ----
// A moving sphere enters the pre-filter as the static sphere that bounds its whole sweep (centre0 + delta/2, radius
// r + |delta|/2), so stage 1 never looks at the ray's time.  Its exact block carries centre0 and a NEGATIVE r^2 as the
// "moving" tag; stage 2 then evaluates the reference's own expression with the centre at ray.time.
struct __align__(16) DevMotion {  // 32 B, one record per sphere (zeros for static spheres), read only on the rare path
    float dx, dy, dz;        // centre_delta = centre1 - centre0   (moving_sphere.rs:21)
    float time_start;        // moving_sphere.rs:23
    float inv_time_delta;    // 1 / (time1 - time0)               (moving_sphere.rs:24)
    float radius;            // signed, as given
    float _pad[2];
};

// moving_sphere.rs:38-73 for one sphere: returns the accepted root (near if inside (t_min, t_max), else far), or -1.
// Out of line on purpose: it is rare, and inlining it into the resolve loop costs the hot kernels registers.
__device__ __forceinline__ float moving_sphere_hit_t(const DevMotion* __restrict__ mo, float time, float c0x, float c0y, float c0z, float ox,
                                                  float oy, float oz, float dx, float dy, float dz) {
    const DevMotion m = *mo;
    const float s = (time - m.time_start) * m.inv_time_delta;  // moving_sphere.rs:29
    const float cx = c0x + s * m.dx, cy = c0y + s * m.dy, cz = c0z + s * m.dz;
    const float ocx = ox - cx, ocy = oy - cy, ocz = oz - cz;
    const float a = (dx * dx + dy * dy) + dz * dz;
    const float b = (ocx * dx + ocy * dy) + ocz * dz;
    const float c = ((ocx * ocx + ocy * ocy) + ocz * ocz) - m.radius * m.radius;
    const float discriminant = b * b - a * c;
    if (discriminant > 0.0f) {
        const float discriminant_sqrt = sqrtf(discriminant);
        float t = (-b - discriminant_sqrt) / a;
        if (t < kMaxT && t > kMinT) return t;
        t = (-b + discriminant_sqrt) / a;
        if (t < kMaxT && t > kMinT) return t;
    }
    return -1.0f;
}

// what the rare paths need to know about motion: the table and where this lane keeps its ray's time
struct MotionCtx {
    const DevMotion* table;        // nullptr when the scene has no moving sphere
    const float* time;             // this path's ray.time (a shared-memory slot; NOT volatile: a volatile shared load in the
                                   // re-test path makes ptxas drop the sweep's uniform operands, see pt_wave.cuh)
    const uint32_t* order;         // stored index -> position in the caller's sphere list (nullptr: identity)
};

#ifndef PT_CTA_THREADS
#define PT_CTA_THREADS 256
#endif
constexpr int kSweepThreads = PT_CTA_THREADS;  // == kCtaThreads (queue stride)

// Candidate queue: one 32-bit entry per flagged group = (group index << MASK_BITS) | one flag bit per sphere of the group,
// kQueueCap entries per lane in shared memory ([entry][thread] layout, conflict-free).  A full queue (rare) tests the
// group on the spot.  Stage 2: every lane walks its own queue after the sweep, all lanes in parallel — divergence costs
// a few trips of one SIMD loop per sweep instead of a serialized branch body per (lane, sphere) event.
constexpr int kQueueCap = 12;

template <int MASK_BITS, bool MOTION, bool ORDERED>
__device__ __forceinline__ void sweep_resolve_entry(const float4* __restrict__ blk, const MotionCtx& mc, uint32_t entry, float ox, float oy, float oz,
                                                    float dx, float dy, float dz, float& hit_t, int& hit_index, unsigned& flagged) {
    const int base = (int)(entry >> MASK_BITS) * MASK_BITS;  // MASK_BITS spheres per group
    uint32_t mask = entry & ((1u << MASK_BITS) - 1u);
    flagged += (unsigned)__popc(mask);  // diagnostic (pt_debug_hits); dead code in the render kernels
#pragma unroll 1
    while (mask != 0u) {
        const int index = base + __ffs(mask) - 1;
        mask &= mask - 1u;
        const float* bf = reinterpret_cast<const float*>(blk) + (index >> 2) * 16 + (index & 3);
        const float r2 = bf[12];
        if (MOTION && r2 < 0.0f) {  // MovingSphere tag (MOTION is a kernel template parameter: static scenes compile none of this)
            const float t = moving_sphere_hit_t(mc.table + index, *mc.time, bf[0], bf[4], bf[8], ox, oy, oz, dx, dy, dz);
            // nearest hit, lowest index among equal t (the strict `<` of the in-order walk, hitable_list.rs:49-54)
            if (t > 0.0f && (t < hit_t || (t == hit_t && (ORDERED ? tie_goes_to(mc.order, index, hit_index) : index < hit_index)))) {
                hit_t = t;
                hit_index = index;
            }
        } else {
            sweep_exact<ORDERED>(bf[0] - ox, bf[4] - oy, bf[8] - oz, r2, dx, dy, dz, index, hit_t, hit_index, mc.order);
        }
    }
}

// =====================================================================================================
// Expanded pre-filter with shared-memory operands (resident and streamed kernels).
//
// A = c.d - o.d, B = 2 c.o + k, L = A*A + B, candidate <=> L > |o|^2 (1 - 2^-18); the sphere pairs come from the
// pre-filter image staged in shared memory (LDS.128 broadcast): 7 packed instructions per 2 tests of the shape
// FFMA2 Rpair, Rpair(spheres), Rscalar(ray), Rpair — against 11 for the sphere-relative form co = c - o, whose packed
// instructions mostly read two or three register pairs (profiles/probe_forms_r1.txt; history in DESIGN.md §4.1).
// `pf` holds blocks [first_block, first_block + n_blocks) of the image (n_blocks a multiple of the group size);
// flagged groups go to the lane's queue as (absolute block << kLdsMaskBits | flags) and are re-tested by the
// caller with the reference's exact expression against `exact` (global memory for the streamed kernel).
// =====================================================================================================
#ifndef PT_LDS_GROUP
#define PT_LDS_GROUP 4
#endif
constexpr int kLdsGroupBlocks = PT_LDS_GROUP;      // blocks (of 4 spheres) per pre-filter branch / queue entry
constexpr int kLdsMaskBits = 4 * kLdsGroupBlocks;  // entry = absolute group << kLdsMaskBits | flags
constexpr long long kMaxSweepBlocks = (1ll << (32 - kLdsMaskBits)) * kLdsGroupBlocks;  // what the entry's group field can address
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
template <bool PIPE, bool MOTION>
__device__ __forceinline__ void sweep_expanded(const float4* __restrict__ pf, int n_blocks, int first_block, const float4* __restrict__ exact,
                                               const MotionCtx& mc, uint32_t* __restrict__ q, int& cnt, float ox, float oy, float oz, float dx, float dy, float dz,
                                               float nod, float o2x, float o2y, float o2z, float oo, float& hit_t, int& hit_index, unsigned& flagged) {
    // pin the per-ray operands (and the trip count) in registers: without this ptxas rematerialises them — 9 scalar FP
    // instructions and a constant-bank reload of n_blocks — in every trip
    asm volatile("" : "+f"(nod), "+f"(o2x), "+f"(o2y), "+f"(o2z), "+f"(oo), "+r"(n_blocks));
    // explicit shared-window address, advanced by one group per trip (keeps the loop's address arithmetic to one add)
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(pf);
    asm volatile("" : "+r"(addr));
    // the lane's queue base as a pinned shared-window address: left to itself ptxas rebuilds it from %tid, the CTA's
    // shared window and n_blocks (12 instructions incl. S2R/S2UR) inside every candidate push — 27 % of the group
    // iterations on cfg2.  Resident kernel only: the streamed kernel (PIPE) has no register to spare for it (-3 % there).
    uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(q);
    if (!PIPE) asm volatile("" : "+r"(qaddr));
    // PIPE: software pipeline, one block deep — block b+1 is in flight while block b is tested (LDS latency off the
    // FFMA2 chain) at the price of 16 registers; pays when the kernel is at 2 CTAs/SM anyway (streamed tiles)
    const uint32_t last = addr + 64u * (uint32_t)(n_blocks > 0 ? n_blocks - 1 : 0);
    float4 cX, cY, cZ, cK;
    if (PIPE) {
        cX = lds128(addr); cY = lds128(addr + 16u); cZ = lds128(addr + 32u); cK = lds128(addr + 48u);
    }
    const uint32_t base = addr, end = addr + 64u * (uint32_t)n_blocks;
    constexpr int kGroupUnroll = PIPE ? PT_GROUP_UNROLL_STREAMED : PT_GROUP_UNROLL_RESIDENT;
#pragma unroll kGroupUnroll
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
            // flag bits from sign bits: oo - L is negative exactly when L > oo (a float difference is zero only for equal
            // operands), so one packed subtract per pair and one funnel shift per sphere build the mask — half the
            // instructions of a compare + select + add per sphere.  Highest sphere first, so bit e = sphere e of the group.
            uint32_t mask = 0u;
#pragma unroll
            for (int p = 2 * kLdsGroupBlocks - 1; p >= 0; --p) {
                const float2 d = f2_fma(L[p], make_float2(-1.0f, -1.0f), make_float2(oo, oo));
                mask = __funnelshift_l(__float_as_uint(d.y), mask, 1);
                mask = __funnelshift_l(__float_as_uint(d.x), mask, 1);
            }
            const uint32_t entry = ((((uint32_t)first_block + ((addr - base) >> 6)) / kLdsGroupBlocks) << kLdsMaskBits) | mask;
            if (cnt < kQueueCap) {
                if (!PIPE) asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + (uint32_t)cnt * (uint32_t)(kSweepThreads * 4)), "r"(entry) : "memory");
                else q[cnt * kSweepThreads] = entry;
                cnt += 1;
            } else {  // queue full (rare): test this group now; sweep_exact's tie rule makes the visiting order irrelevant
                sweep_resolve_entry<kLdsMaskBits, MOTION, !PIPE>(exact, mc, entry, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
            }
        }
    }
}

// =====================================================================================================
// Two rays per lane, sphere pairs as UNIFORM operands (resident kernel, round 2).
//
// Measured on B200 (tools/probe_forms.cu, tools/probe_sweep2.cu, profiles/probe_*_r2*.txt): the packed FMA is bound by
// register-file reads, not by the FMA pipe.  FFMA2 Rpair(spheres) * Rscalar(ray) + Rpair(acc) reads five register words
// and issues every 3.1 clk per SM sub-partition; FFMA2 Rscalar(ray) * URpair(spheres) + Rpair(acc) reads three and issues
// every 2.1 clk, the pipe's own rate.  Sphere data is the same for every lane, so the X, Y, Z planes of the pre-filter
// image travel as a KERNEL PARAMETER (constant bank 0, `ConstImage`) and reach the FMA through the uniform datapath
// (LDCU.64 -> UR pair); only K (the addend of the B chain: an instruction takes one uniform operand) is read from shared
// memory.  The uniform loads are the next limit (24 LDCU.64 per 16 spheres), so every lane carries TWO rays through the
// sweep and each loaded sphere pair serves both: loop in isolation 8.95-9.4 clk per 32 tests (85-89 % of the FP32 peak at
// 16 flop per test) against 11.0-11.5 (69-73 %) for the round-1 form.
//
// Scenes whose image does not fit the parameter space (> kMaxConstSpheres) run the same two-ray loop with LDS.128
// sphere pairs (CONSTIMG = false): each pair then serves four consecutive FFMA2 through the operand reuse cache (76-81 %).
//
// Queue: the lane's two rays share one queue of QCAP entries ([entry][QSTRIDE threads]), ray 0 filling it from the front and ray 1 from the
// back, so an entry needs no ray tag and each ray is drained with its own registers.  overflow[r] (initialised to n_blocks
// by the caller) comes back as the first block of ray r's first flagged group that found the queue full.
// =====================================================================================================
constexpr int kMaxConstBlocks = 512;                     // 2048 spheres: 24 KB of the 32 764-byte kernel parameter space
constexpr int kMaxConstSpheres = 4 * kMaxConstBlocks;
template <bool CONSTIMG>
struct ConstImageT {
    float4 v[CONSTIMG ? 3 * kMaxConstBlocks : 1];  // per block of 4 spheres: X(cx0..3), Y, Z
};

template <bool CONSTIMG, int QSTRIDE, int QCAP>
__device__ __forceinline__ void sweep_two(const ConstImageT<CONSTIMG>& ci, const float4* __restrict__ pf, int n_blocks, uint32_t* __restrict__ q, int& cnt0, int& cnt1,
                                          const float (&dx)[2], const float (&dy)[2], const float (&dz)[2], float (&o2x)[2], float (&o2y)[2],
                                          float (&o2z)[2], float (&nod)[2], float (&oo)[2], int (&overflow)[2]) {
    // pin the per-ray operands in registers (ptxas otherwise rematerialises them inside the loop)
    asm volatile("" : "+f"(nod[0]), "+f"(o2x[0]), "+f"(o2y[0]), "+f"(o2z[0]), "+f"(oo[0]));
    asm volatile("" : "+f"(nod[1]), "+f"(o2x[1]), "+f"(o2y[1]), "+f"(o2z[1]), "+f"(oo[1]));
    uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(q);
    asm volatile("" : "+r"(qaddr));
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(pf);
    if constexpr (!CONSTIMG) asm volatile("" : "+r"(addr));
    // CONSTIMG: the loop counter must stay provably uniform — no inline-asm operand may depend on it (an "r" constraint
    // forces a vector register and ptxas then falls back to LDC into vector registers: 17.8 clk per 32 tests)
#pragma unroll 1
    for (int j = 0; j < n_blocks; j += kLdsGroupBlocks) {
        float2 L[2][2 * kLdsGroupBlocks];
#pragma unroll
        for (int g = 0; g < kLdsGroupBlocks; ++g) {
            float4 X, Y, Z, K;
            if constexpr (CONSTIMG) {
                X = ci.v[3 * (j + g) + 0]; Y = ci.v[3 * (j + g) + 1]; Z = ci.v[3 * (j + g) + 2];
                K = pf[j + g];  // K plane only (16 B per block), plain shared load with a uniform index
            } else {
                const uint32_t a4 = addr + 64u * (uint32_t)(j + g);
                X = lds128(a4); Y = lds128(a4 + 16u); Z = lds128(a4 + 32u); K = lds128(a4 + 48u);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 cx = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                const float2 cy = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                const float2 cz = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
                const float2 k = h ? make_float2(K.z, K.w) : make_float2(K.x, K.y);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const float2 A = f2_fma(cz, make_float2(dz[r], dz[r]), f2_fma(cy, make_float2(dy[r], dy[r]), f2_fma(cx, make_float2(dx[r], dx[r]), make_float2(nod[r], nod[r]))));
                    const float2 B = f2_fma(cz, make_float2(o2z[r], o2z[r]), f2_fma(cy, make_float2(o2y[r], o2y[r]), f2_fma(cx, make_float2(o2x[r], o2x[r]), k)));
                    L[r][2 * g + h] = f2_fma(A, A, B);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            bool any = false;
#pragma unroll
            for (int p = 0; p < 2 * kLdsGroupBlocks; ++p) any = any | (L[r][p].x > oo[r]) | (L[r][p].y > oo[r]);
            if (any) {
                // flag bits from sign bits (see sweep_expanded): bit e = sphere e of the group
                uint32_t mask = 0u;
#pragma unroll
                for (int p = 2 * kLdsGroupBlocks - 1; p >= 0; --p) {
                    const float2 d = f2_fma(L[r][p], make_float2(-1.0f, -1.0f), make_float2(oo[r], oo[r]));
                    mask = __funnelshift_l(__float_as_uint(d.y), mask, 1);
                    mask = __funnelshift_l(__float_as_uint(d.x), mask, 1);
                }
                const uint32_t entry = (((uint32_t)j / kLdsGroupBlocks) << kLdsMaskBits) | mask;
                if (cnt0 + cnt1 < QCAP) {
                    const uint32_t slot = r == 0 ? (uint32_t)cnt0 : (uint32_t)(QCAP - 1 - cnt1);
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + slot * (uint32_t)(QSTRIDE * 4)), "r"(entry) : "memory");
                    if (r == 0) cnt0 += 1; else cnt1 += 1;
                } else {
                    // queue full (rare): remember the first group that did not fit; the caller re-tests every sphere from
                    // there on with the exact expression (sweep_overflow).  Nothing else may sit in this loop: ptxas keeps
                    // the sphere operands in uniform registers only while it can prove that the warp runs the loop in
                    // lockstep, and a data-dependent inner loop or a call in here makes it fall back to vector loads.
                    overflow[r] = min(overflow[r], j);
                }
            }
        }
    }
}

// groups [first_block / kLdsGroupBlocks, n_blocks / kLdsGroupBlocks): the exact test on every sphere (queue overflow)
template <bool MOTION>
__device__ __forceinline__ void sweep_overflow(const float4* __restrict__ exact, const MotionCtx& mc, int first_block, int n_blocks, float ox, float oy, float oz,
                                            float dx, float dy, float dz, float& hit_t, int& hit_index, unsigned& flagged) {
    for (int g = first_block / kLdsGroupBlocks; g < n_blocks / kLdsGroupBlocks; ++g)
        sweep_resolve_entry<kLdsMaskBits, MOTION, true>(exact, mc, ((uint32_t)g << kLdsMaskBits) | ((1u << kLdsMaskBits) - 1u), ox, oy, oz, dx, dy, dz, hit_t, hit_index,
                                                        flagged);
}

// drain `cnt` queue entries starting at entry `first` (the lane's [entry][QSTRIDE threads] queue), all lanes in parallel
template <bool MOTION, int QSTRIDE>
__device__ __forceinline__ void sweep_drain_range(const float4* __restrict__ exact, const MotionCtx& mc, const uint32_t* __restrict__ q, int first, int cnt, float ox,
                                                  float oy, float oz, float dx, float dy, float dz, float& hit_t, int& hit_index, unsigned& flagged) {
    uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(q) + (uint32_t)first * (uint32_t)(QSTRIDE * 4);
    asm volatile("" : "+r"(qaddr));
    const uint32_t qend = qaddr + (uint32_t)cnt * (uint32_t)(QSTRIDE * 4);
#pragma unroll 1
    for (; qaddr != qend; qaddr += (uint32_t)(QSTRIDE * 4)) {
        uint32_t entry;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(entry) : "r"(qaddr) : "memory");
        sweep_resolve_entry<kLdsMaskBits, MOTION, true>(exact, mc, entry, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
    }
}

// drain the lane's queue: exact re-test of every flagged sphere (all lanes in parallel).  PIN: walk the queue through a
// pinned shared-window address (resident kernel; see the push in sweep_expanded)
template <bool MOTION, bool PIN>
__device__ __forceinline__ void sweep_drain(const float4* __restrict__ exact, const MotionCtx& mc, const uint32_t* __restrict__ q, int& cnt, float ox, float oy, float oz,
                                            float dx, float dy, float dz, float& hit_t, int& hit_index, unsigned& flagged) {
    if (PIN) {
        uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(q);
        asm volatile("" : "+r"(qaddr));
        const uint32_t qend = qaddr + (uint32_t)cnt * (uint32_t)(kSweepThreads * 4);
#pragma unroll 1
        for (; qaddr != qend; qaddr += (uint32_t)(kSweepThreads * 4)) {
            uint32_t entry;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(entry) : "r"(qaddr) : "memory");
            sweep_resolve_entry<kLdsMaskBits, MOTION, true>(exact, mc, entry, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
        }
    } else {
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) sweep_resolve_entry<kLdsMaskBits, MOTION, false>(exact, mc, q[i * kSweepThreads], ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
    }
    cnt = 0;
}

}  // namespace pt
