// pt_sweep_mma.cuh — stage 1 of the sweep (the conservative pre-filter over ALL spheres, pt_sweep.cuh) on the warp-level
// tensor path: mma.sync.m16n8k16 f16 -> HMMA.16816.F32 on sm_100a.
//
// Why.  The FP32 form of the pre-filter is bound by register-file reads of the packed FMA (3.1 clk per FFMA2 where the pipe
// needs 2: DESIGN.md §5.2, tools/probe_forms.cu); its loop runs at 11.0-11.5 clk per 32 (ray, sphere) tests per SM
// sub-partition.  But the pre-filter is two small dot products per test followed by one FMA and a sign test — a GEMM of the
// sphere table against the warp's 32 rays with K = 5 — and the exact stage 2 decides every hit anyway, so stage 1 only has
// to be CONSERVATIVE, not f32-exact.  The same loop on the tensor path: 6.5-7.5 clk per 32 tests, measured in isolation with
// the whole operand pipeline and the candidate push (tools/probe_mma_sweep.cu, profiles/probe_mma_sweep_r2*.txt).
//
// Form.  sigma and s are powers of two chosen per scene (host: build_mma_image, ptgpu.cu):
//   S   = [sigma cx, sigma cy, sigma cz, s, K'/s]          K' = sigma^2 (r^2 - |c|^2 + slack (|c|^2 + r^2)) + abs_slack
//   R_A = [dx, dy, dz, sigma (-o.d) / s, 0]
//   R_B = [2 sigma ox, 2 sigma oy, 2 sigma oz, -sigma^2 |o|^2 (1 - slack) / s, s]
//   A' = S.R_A = sigma (c.d - o.d),   B' = S.R_B = sigma^2 (2 c.o - |o|^2 (1 - slack)) + K',   candidate <=> A'^2 + B' > 0
// which is sigma^2 times the FP32 filter's L - |o|^2 (1 - slack).  Every f32 element x is split into two f16 pieces,
// hi = rn(x), lo = rn(x - hi) (22 significant bits), and the three products hi*hi + hi*lo + lo*hi are laid along K:
//   K index 0..4 S_hi R_hi, 5..9 S_hi R_lo, 10..14 S_lo R_hi, 15 unused  ->  ONE m16n8k16 MMA per dot product.
// sigma maps the scene's extent to 16384 (so |2 sigma o| <= 32768 < 65504, the f16 maximum), s keeps K' and |o|^2 there.
//
// Error budget (what `slack` must cover; all relative to |c|^2 + r^2 + |o|^2 in scaled units): the split drops lo*lo and
// rounds lo, <= 3 * 2^-22 per product; the tensor core accumulates the exact f16 products in f32, measured <= 2^-21 of the
// sum of |terms| (tools/probe_mma.cu); 2 |A'| dA' + dB' then stays below 2^-17.1, and the reference's own f32 discriminant
// differs from the true one by <= 2^-19 (pt_sweep.cuh).  kMmaSlack = 2^-15 is more than three times the sum.  f16
// subnormals (elements below 2^-14 after scaling) add at most 2^-12 in absolute terms: abs_slack = 2^-7.  Checked per ray,
// not assumed: pt_debug_hits through this sweep against the exact test on every sphere (tests/test_gpu_hits.py).
//
// Layout.  D[ray][sphere]: the rays are the A operand (row-major 16 x 16, loop-invariant fragments), the spheres the B
// operand (16 x 8, one LDS.128 per lane per 16 spheres from the fragment-ordered image).  Lane (g, t) = (lane >> 2, lane & 3)
// puts its ray in row 8 (t & 1) + g of row block t >> 1, so the four rays whose results a quad holds (rows g, g + 8 of both
// row blocks) are the four rays its lanes own: a lane that finds a candidate pushes (step, 16 flag bits) to its OWN queue —
// no ballots, no atomics — and in stage 2 every lane walks the queues of its quad and takes the bits of its own ray.
//   value (rb, sg, c) of lane (g, t): ray slot 2 rb + (c >> 1) of the quad (= the owner's t), sphere 16 step + 8 sg + 2 t + (c & 1)
//   flag bit b = 8 rb + 4 (c >> 1) + 2 sg + (c & 1): nibble `slot` of the entry belongs to the quad's lane t = slot
#pragma once
#include <cuda_fp16.h>
#include "pt_sweep.cuh"

namespace pt {

// steps per trip of the loop (an experiment knob: 2 measured no faster on cfg2/cfg4/cfg5, profiles/variants_r2_mma.txt)
#ifndef PT_MMA_UNROLL
#define PT_MMA_UNROLL 1
#endif
constexpr int kMmaUnroll = PT_MMA_UNROLL;
constexpr float kMmaSlack = 3.0517578125e-05f;    // 2^-15
constexpr double kMmaSlackSphere = 3.0517578125e-05;
constexpr double kMmaAbsSlack = 1.0 / 128.0;       // scaled units
constexpr double kMmaExtentTarget = 16384.0;       // sigma * extent <= this
constexpr int kMmaStageRows = 16;                  // rows of kSweepThreads words the ray fragments pass through

struct MmaScale {
    float sigma, s, inv_s;
    float max_o2;      // rays with |o - t|^2 beyond this (the square of the extent the scales were chosen for) bypass stage 1
    float tx, ty, tz;  // the scene's offset: the filter works on c - t and o - t (the discriminant does not depend on t).  An axis
                       // is shifted only when the scene lies at more than four extents from the origin on it, so that o - t is
                       // exact for every origin inside the extent (Sterbenz) and the ray tested is the ray given, to the bit
};

__device__ __forceinline__ void hmma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "f"(0.0f));
}
__device__ __forceinline__ uint32_t mma_pack(__half a, __half b) { return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16); }

// the 16 words of one ray's operand: words 0..7 the R_A column (K halves 2i, 2i+1 in word i), words 8..15 the R_B column.
// (ox, oy, oz) is the origin relative to the scene's offset.  A lane without a path in flight gets the parked operand:
// A' = 0, B' = K' - 65504 s < 0 for every sphere.
__device__ __forceinline__ void mma_ray_operand(const MmaScale sc, bool active, float ox, float oy, float oz, float dx, float dy, float dz, uint32_t (&w)[16]) {
    const float nod = -((ox * dx + oy * dy) + oz * dz);
    const float oo = ((ox * ox + oy * oy) + oz * oz) * (1.0f - kMmaSlack);
    float ra[5] = {dx, dy, dz, sc.sigma * nod * sc.inv_s, 0.0f};
    float rb[5] = {2.0f * sc.sigma * ox, 2.0f * sc.sigma * oy, 2.0f * sc.sigma * oz, -(sc.sigma * sc.sigma) * oo * sc.inv_s, sc.s};
    if (!active) {
#pragma unroll
        for (int e = 0; e < 5; ++e) ra[e] = rb[e] = 0.0f;
        rb[3] = -65504.0f;
        rb[4] = sc.s;
    }
    __half v[2][16];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int e = 0; e < 5; ++e) {
            const float x = c ? rb[e] : ra[e];
            const __half hi = __float2half_rn(x);
            const __half lo = __float2half_rn(x - __half2float(hi));
            v[c][e] = hi;
            v[c][5 + e] = lo;
            v[c][10 + e] = hi;
        }
        v[c][15] = __float2half_rn(0.0f);
    }
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) w[c * 8 + i] = mma_pack(v[c][2 * i], v[c][2 * i + 1]);
}

// exact re-test of ONE stored sphere (stage 2): the body of sweep_resolve_entry's loop
template <bool MOTION>
__device__ __forceinline__ void sweep_test_index(const float4* __restrict__ blk, const MotionCtx& mc, int index, float ox, float oy, float oz, float dx, float dy,
                                                 float dz, float& hit_t, int& hit_index) {
    const float* bf = reinterpret_cast<const float*>(blk) + (index >> 2) * 16 + (index & 3);
    const float r2 = bf[12];
    if (MOTION && r2 < 0.0f) {
        const float t = moving_sphere_hit_t(mc.table + index, *mc.time, bf[0], bf[4], bf[8], ox, oy, oz, dx, dy, dz);
        if (t > 0.0f && (t < hit_t || (t == hit_t && tie_goes_to(mc.order, index, hit_index)))) {
            hit_t = t;
            hit_index = index;
        }
    } else {
        sweep_exact<true>(bf[0] - ox, bf[4] - oy, bf[8] - oz, r2, dx, dy, dz, index, hit_t, hit_index, mc.order);
    }
}

// The lane's ray as A-operand fragments of both row blocks (loop-invariant over the sphere loop).
struct MmaRayFrags {
    uint4 fa[2], fb[2];
    bool in_range;  // the ray lies inside the operands' validity domain (otherwise it takes no part in stage 1)
};

// stage: kMmaStageRows rows of kSweepThreads words, of which this warp uses its own 32 columns (a transpose buffer).
__device__ __forceinline__ void mma_ray_fragments(uint32_t* __restrict__ stage, const MmaScale sc, bool active, float ox, float oy, float oz, float dx, float dy,
                                                  float dz, MmaRayFrags& f) {
    const unsigned lane = threadIdx.x & 31u, g = lane >> 2, t = lane & 3u;
    uint32_t* warp_cols = stage + (threadIdx.x & ~31u);
    // validity domain of the f16 operands: origin inside the extent the scales were chosen for, direction of unit size.
    // A ray outside it (or a non-finite one) takes no part in stage 1 and gets the exact test on every sphere instead.
    const float sx = ox - sc.tx, sy = oy - sc.ty, sz = oz - sc.tz;
    f.in_range = ((sx * sx + sy * sy) + sz * sz) <= sc.max_o2 && ((dx * dx + dy * dy) + dz * dz) <= 4.0f;
    // ray operands -> A fragments.  Fragment (quad g, row block rb, t, column type c) is four consecutive words
    // {ray0.word[t], ray1.word[t], ray0.word[t+4], ray1.word[t+4]} at row c*8 + rb*4 + t, columns 4g..4g+3 of the warp:
    // the owner of ray (rb = t >> 1, half = t & 1) scatters its 16 words, every lane reads its four fragments with LDS.128.
    {
        uint32_t w[16];
        mma_ray_operand(sc, active && f.in_range, sx, sy, sz, dx, dy, dz, w);
        uint32_t* base = warp_cols + 4u * g + (t & 1u);
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int k = 0; k < 8; ++k) base[(c * 8 + (int)(t >> 1) * 4 + (k & 3)) * kSweepThreads + 2 * (k >> 2)] = w[c * 8 + k];
    }
    __syncwarp();
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
        f.fa[rb] = *reinterpret_cast<const uint4*>(warp_cols + (rb * 4 + (int)t) * kSweepThreads + 4u * g);
        f.fb[rb] = *reinterpret_cast<const uint4*>(warp_cols + (8 + rb * 4 + (int)t) * kSweepThreads + 4u * g);
    }
    __syncwarp();
}

// Stage 1 over steps [first_step, first_step + n_steps) of the image, whose fragments start at simg (the loop fetches one
// step ahead: 512 readable bytes must follow the last step).  q: this lane's queue, [kQueueCap][kSweepThreads]; cnt: entries
// in it; ovf_step: lowered to the first step whose entry found the queue full.
__device__ __forceinline__ void sweep_mma_steps(const uint4* __restrict__ simg, int first_step, int n_steps, const MmaRayFrags& f, uint32_t* __restrict__ q, int& cnt,
                                                int& ovf_step) {
    const unsigned lane = threadIdx.x & 31u;
    uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(q);
    asm volatile("" : "+r"(qaddr));
    const uint4* p = simg + lane;
    uint4 sp = *p;
    const int end = first_step + n_steps;
#pragma unroll kMmaUnroll
    for (int s = first_step; s < end; ++s) {
        p += 32;
        const uint4 cur = sp;
        float A[2][2][4], B[2][2][4];
#pragma unroll
        for (int rb = 0; rb < 2; ++rb)
#pragma unroll
            for (int sg = 0; sg < 2; ++sg) {
                hmma16816(A[rb][sg], f.fa[rb], sg ? cur.z : cur.x, sg ? cur.w : cur.y);
                hmma16816(B[rb][sg], f.fb[rb], sg ? cur.z : cur.x, sg ? cur.w : cur.y);
            }
        sp = *p;  // next step's sphere fragments
        // N = -(A'^2 + B'): candidate <=> N < 0 <=> sign bit (a -0.0 is a harmless false positive)
        float2 N[2][2][2];
        float m = 3.0e38f;
#pragma unroll
        for (int rb = 0; rb < 2; ++rb)
#pragma unroll
            for (int sg = 0; sg < 2; ++sg)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 a2 = make_float2(A[rb][sg][2 * h], A[rb][sg][2 * h + 1]);
                    N[rb][sg][h] = f2_fma(make_float2(-a2.x, -a2.y), a2, make_float2(-B[rb][sg][2 * h], -B[rb][sg][2 * h + 1]));
                    m = fminf(fminf(N[rb][sg][h].x, N[rb][sg][h].y), m);
                }
        if (m < 0.0f) {  // this lane holds a candidate (operands are finite and inside the f16 range by construction: no NaN)
            uint32_t mask = 0u;  // bit b = 8 rb + 4 h + 2 sg + e, built from the highest bit down
#pragma unroll
            for (int rb = 1; rb >= 0; --rb)
#pragma unroll
                for (int h = 1; h >= 0; --h)
#pragma unroll
                    for (int sg = 1; sg >= 0; --sg) {
                        mask = __funnelshift_l(__float_as_uint(N[rb][sg][h].y), mask, 1);
                        mask = __funnelshift_l(__float_as_uint(N[rb][sg][h].x), mask, 1);
                    }
            mask &= 0xffffu;
            if (cnt < kQueueCap) {
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + (uint32_t)cnt * (uint32_t)(kSweepThreads * 4)), "r"(((uint32_t)s << 16) | mask) : "memory");
                cnt += 1;
            } else {
                ovf_step = min(ovf_step, s);
            }
        }
    }
}

// Stage 1 for the warp's 32 rays against a resident image of n_steps steps ((n_steps + 1) x 32 uint4: one step of padding).
// Returns the number of entries pushed; ovf_step comes back as the first step that needs the exact test on everything
// (n_steps: none; 0: the ray is outside the operands' domain).
__device__ __forceinline__ int sweep_mma(const uint4* __restrict__ simg, int n_steps, uint32_t* __restrict__ stage, uint32_t* __restrict__ q, const MmaScale sc,
                                         bool active, float ox, float oy, float oz, float dx, float dy, float dz, int& ovf_step) {
    MmaRayFrags f;
    mma_ray_fragments(stage, sc, active, ox, oy, oz, dx, dy, dz, f);
    int cnt = 0;
    ovf_step = (active && !f.in_range) ? 0 : n_steps;
    sweep_mma_steps(simg, 0, n_steps, f, q, cnt, ovf_step);
    return cnt;
}

// Stage 2 for this lane's ray over the steps [first_step, end_step) stage 1 has just covered: the exact test on every sphere
// one of the quad's lanes flagged for it.  qbase: the queue array WITHOUT the thread offset ([kQueueCap][kSweepThreads]).
// An entry's step field has 16 bits: scenes of up to 65535 steps (1 M spheres).
template <bool MOTION>
__device__ __forceinline__ void sweep_mma_drain(const float4* __restrict__ exact, const MotionCtx& mc, const uint32_t* __restrict__ qbase, int cnt, int ovf_step,
                                                int first_step, int end_step, float ox, float oy, float oz, float dx, float dy, float dz, float& hit_t, int& hit_index,
                                                unsigned& flagged) {
    const unsigned lane = threadIdx.x & 31u, t = lane & 3u;
    const unsigned quad_thread = threadIdx.x & ~3u;
    __syncwarp();  // the quad's pushes are visible
    // a full queue somewhere in the quad, or a ray outside the operands' domain (both rare): every sphere from that step on
    // gets the exact test below, and queue entries from there on are skipped (they would only repeat it)
    int ovf = ovf_step;
    ovf = min(ovf, __shfl_xor_sync(0xffffffffu, ovf, 1));
    ovf = min(ovf, __shfl_xor_sync(0xffffffffu, ovf, 2));
    ovf = max(ovf, first_step);
#pragma unroll 1
    for (unsigned tf = 0; tf < 4u; ++tf) {
        const int n = __shfl_sync(0xffffffffu, cnt, (int)((lane & ~3u) + tf));
        const uint32_t* col = qbase + quad_thread + tf;
#pragma unroll 1
        for (int e = 0; e < n; ++e) {
            const uint32_t entry = col[e * kSweepThreads];
            const int step = (int)(entry >> 16);
            uint32_t nib = step < ovf ? (entry >> (4u * t)) & 15u : 0u;
            flagged += (unsigned)__popc(nib);
            const int base = step * 16 + 2 * (int)tf;
#pragma unroll 1
            while (nib != 0u) {
                const int b = __ffs(nib) - 1;
                nib &= nib - 1u;
                sweep_test_index<MOTION>(exact, mc, base + 8 * (b >> 1) + (b & 1), ox, oy, oz, dx, dy, dz, hit_t, hit_index);
            }
        }
    }
    if (ovf < end_step) sweep_overflow<MOTION>(exact, mc, ovf * kLdsGroupBlocks, end_step * kLdsGroupBlocks, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
    __syncwarp();  // nobody overwrites a queue its quad is still reading
}

}  // namespace pt
