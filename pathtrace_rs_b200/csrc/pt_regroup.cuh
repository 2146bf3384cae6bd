// pt_regroup.cuh — the resident kernel: ONE path per lane and a CTA-level regroup of the paths between sweep and shading.
// The default for every scene that fits in shared memory, in two instantiations of the same kernel: stage 1 of the sweep on
// the tensor path (MMA = true, pt_sweep_mma.cuh: cfg2 70.6 %, cfg4 72.9 % of the FP32 peak in algorithmic flop) or in packed
// FP32 (MMA = false, pt_sweep.cuh: 55.2 % / 56.7 %; small scenes, ill-scaled scenes, cameras outside the scene's extent).
//
// Round 2 also built two alternatives around a cheaper FP32 sweep — two paths per lane with the sphere pairs as uniform
// operands from a kernel-parameter image (pt_megakernel_resident, 44 %) and an asynchronous wavefront form with a path pool
// and per-material queues in shared memory (pt_wave.cuh, 38 %) — and measured both slower than this kernel on the real
// workloads although their sweep loops are 10-23 % faster in isolation (tools/probe_sweep2.cu): what they save in the loop
// they lose around it (DESIGN.md §4.6 / §5.2 have the numbers and the ncu evidence).  They stay selectable through
// PtOptions.resident_kernel.
#pragma once
#include "pt_megakernel.cuh"

namespace pt {

// =====================================================================================================
// CTA-level regrouping of paths by what they do next.
//
// After the sweep the 256 lanes of a CTA are about to run different code: Lambertian / textured Lambertian / metal /
// dielectric scatter, or end their path (miss, light, depth limit) and start a new sample at the next refill.  Left in
// place, every warp executes the union of those branches with a quarter of its lanes (ncu, cfg2: 8.5 of 32 lanes active
// outside the sweep).  Lanes are interchangeable — a lane is only the register home of one path's state — so once per
// trip the CTA counting-sorts its paths by category through shared memory: ballots give each lane its rank inside its
// warp, one shared-memory atomicAdd per (warp, category) reserves the warp's range inside the category, and the whole
// path state (30 words) is written to its new slot and read back by the thread that now owns it.  Every path still
// consumes exactly its own RNG stream and performs exactly the same arithmetic, so images stay bit-identical; only the
// assignment of paths to lanes changes.  Finished lanes collect in whole warps, which then skip the sweep.
// =====================================================================================================
// PT_REGROUP_DOMAINS: the CTA's paths are regrouped within this many independent groups of warps (named barriers), an
// experiment knob: 2 halves the number of warps that wait for each other at the price of a coarser sort
#ifndef PT_REGROUP_DOMAINS
#define PT_REGROUP_DOMAINS 1
#endif
#ifndef PT_REGROUP_PERIOD
#define PT_REGROUP_PERIOD 1
#endif
constexpr int kRegroupCats = 6;
constexpr int kRegroupWords = 28;  // 8 rng + 6 ray + 3 thr + 3 col + px, py, sample, depth, flags, hit_t, hit_index, time
enum { CAT_LAMBERT_CONST = 0, CAT_LAMBERT_TEX = 1, CAT_METAL = 2, CAT_DIELECTRIC = 3, CAT_ENDING = 4, CAT_IDLE = 5 };

// what a hit on stored sphere `index` does next (the scatter code its material runs)
__device__ __forceinline__ int sphere_category(const KernelArgs& a, int index) {
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(a.shade + index) + 1);
    const int kind = __float_as_int(s1.y);
    if (kind == MAT_LAMBERTIAN) return __float_as_int(s1.z) < 0 ? CAT_LAMBERT_CONST : CAT_LAMBERT_TEX;
    if (kind == MAT_METAL) return CAT_METAL;
    if (kind == MAT_DIELECTRIC) return CAT_DIELECTRIC;
    return CAT_ENDING;  // DiffuseLight
}
// cat_table: one byte per stored sphere in shared memory (EXACT_SMEM kernels), or nullptr: read the shading record
__device__ __forceinline__ int lane_category(const KernelArgs& a, const Lane& L, int hit_index, const uint8_t* __restrict__ cat_table) {
    if (!L.active) return CAT_IDLE;
    if (hit_index < 0 || L.depth >= a.max_depth) return CAT_ENDING;
    return cat_table ? (int)cat_table[hit_index] : sphere_category(a, hit_index);
}

// xchg: [kRegroupWords][kCtaThreads] words; cat_count: this trip's [kRegroupCats] counters (two sets alternate: the set
// used by trip t is cleared after trip t's second barrier and next touched after trip t+1's first barrier).
// Two CTA barriers per trip; the second also ORs "some lane still has work" over the CTA and returns it.  Trip t+1's
// first barrier separates trip t's reads of xchg from trip t+1's writes.
constexpr int kRegroupDomainThreads = kCtaThreads / PT_REGROUP_DOMAINS;
__device__ __forceinline__ void regroup_barrier(unsigned dom) {
    if (PT_REGROUP_DOMAINS == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(1u + dom), "r"((unsigned)kRegroupDomainThreads) : "memory");
}
__device__ __forceinline__ bool regroup_barrier_or(unsigned dom, bool flag) {
    if (PT_REGROUP_DOMAINS == 1) return __syncthreads_or(flag ? 1 : 0) != 0;
    unsigned out;
    asm volatile("{ .reg .pred p, q; setp.ne.u32 p, %1, 0; bar.red.or.pred q, %2, %3, p; selp.u32 %0, 1, 0, q; }"
                 : "=r"(out) : "r"(flag ? 1u : 0u), "r"(1u + dom), "r"((unsigned)kRegroupDomainThreads) : "memory");
    return out != 0u;
}
__device__ __forceinline__ bool cta_regroup(const KernelArgs& a, Lane& L, float& hit_t, int& hit_index, uint32_t* __restrict__ xchg,
                                            uint32_t* __restrict__ cat_count, unsigned lane_id, const uint8_t* __restrict__ cat_table = nullptr) {
    const unsigned dom = PT_REGROUP_DOMAINS == 1 ? 0u : threadIdx.x / (unsigned)kRegroupDomainThreads;
    const unsigned dom_base = dom * (unsigned)kRegroupDomainThreads;
    cat_count += dom * 16u;
    const int cat = lane_category(a, L, hit_index, cat_table);
    unsigned mine = 0u;     // ballot of this lane's category
    unsigned warp_off = 0u;  // where this warp's lanes of that category start inside the category
#pragma unroll
    for (int c = 0; c < kRegroupCats; ++c) {
        const unsigned b = __ballot_sync(kFullMask, cat == c);
        unsigned off = 0u;
        if (lane_id == 0u && b != 0u) off = atomicAdd(&cat_count[c], (unsigned)__popc(b));
        off = __shfl_sync(kFullMask, off, 0);
        if (cat == c) {
            mine = b;
            warp_off = off;
        }
    }
    regroup_barrier(dom);  // all counts are final
    unsigned base = dom_base;
#pragma unroll
    for (int c = 0; c < kRegroupCats - 1; ++c) base += (c < cat) ? cat_count[c] : 0u;
    const unsigned dest = base + warp_off + (unsigned)__popc(mine & ((1u << lane_id) - 1u));
    uint32_t* w = xchg + dest;
    const uint32_t flags = (L.active ? 1u : 0u) | (L.have_pixel ? 2u : 0u) | (L.finished ? 4u : 0u) | (L.pend ? 8u : 0u);
    w[0 * kCtaThreads] = (uint32_t)L.rng.s0;  w[1 * kCtaThreads] = (uint32_t)(L.rng.s0 >> 32);
    w[2 * kCtaThreads] = (uint32_t)L.rng.s1;  w[3 * kCtaThreads] = (uint32_t)(L.rng.s1 >> 32);
    w[4 * kCtaThreads] = (uint32_t)L.rng.s2;  w[5 * kCtaThreads] = (uint32_t)(L.rng.s2 >> 32);
    w[6 * kCtaThreads] = (uint32_t)L.rng.s3;  w[7 * kCtaThreads] = (uint32_t)(L.rng.s3 >> 32);
    w[8 * kCtaThreads] = __float_as_uint(L.o.x);  w[9 * kCtaThreads] = __float_as_uint(L.o.y);  w[10 * kCtaThreads] = __float_as_uint(L.o.z);
    w[11 * kCtaThreads] = __float_as_uint(L.d.x); w[12 * kCtaThreads] = __float_as_uint(L.d.y); w[13 * kCtaThreads] = __float_as_uint(L.d.z);
    w[14 * kCtaThreads] = __float_as_uint(L.thr.x); w[15 * kCtaThreads] = __float_as_uint(L.thr.y); w[16 * kCtaThreads] = __float_as_uint(L.thr.z);
    w[17 * kCtaThreads] = __float_as_uint(L.col.x); w[18 * kCtaThreads] = __float_as_uint(L.col.y); w[19 * kCtaThreads] = __float_as_uint(L.col.z);
    w[20 * kCtaThreads] = L.px;  w[21 * kCtaThreads] = L.py;  w[22 * kCtaThreads] = L.sample;  w[23 * kCtaThreads] = L.depth;
    w[24 * kCtaThreads] = flags;
    w[25 * kCtaThreads] = __float_as_uint(hit_t);
    w[26 * kCtaThreads] = (uint32_t)hit_index;
    w[27 * kCtaThreads] = __float_as_uint(L.time);
    const bool live = regroup_barrier_or(dom, !L.finished);  // every path is in its new slot
    if (threadIdx.x - dom_base < (unsigned)kRegroupCats) cat_count[threadIdx.x - dom_base] = 0u;
    const uint32_t* r = xchg + threadIdx.x;
    L.rng.s0 = (uint64_t)r[0 * kCtaThreads] | ((uint64_t)r[1 * kCtaThreads] << 32);
    L.rng.s1 = (uint64_t)r[2 * kCtaThreads] | ((uint64_t)r[3 * kCtaThreads] << 32);
    L.rng.s2 = (uint64_t)r[4 * kCtaThreads] | ((uint64_t)r[5 * kCtaThreads] << 32);
    L.rng.s3 = (uint64_t)r[6 * kCtaThreads] | ((uint64_t)r[7 * kCtaThreads] << 32);
    L.o = v3(__uint_as_float(r[8 * kCtaThreads]), __uint_as_float(r[9 * kCtaThreads]), __uint_as_float(r[10 * kCtaThreads]));
    L.d = v3(__uint_as_float(r[11 * kCtaThreads]), __uint_as_float(r[12 * kCtaThreads]), __uint_as_float(r[13 * kCtaThreads]));
    L.thr = v3(__uint_as_float(r[14 * kCtaThreads]), __uint_as_float(r[15 * kCtaThreads]), __uint_as_float(r[16 * kCtaThreads]));
    L.col = v3(__uint_as_float(r[17 * kCtaThreads]), __uint_as_float(r[18 * kCtaThreads]), __uint_as_float(r[19 * kCtaThreads]));
    L.px = r[20 * kCtaThreads];  L.py = r[21 * kCtaThreads];  L.sample = r[22 * kCtaThreads];  L.depth = r[23 * kCtaThreads];
    const uint32_t f = r[24 * kCtaThreads];
    L.active = (f & 1u) != 0u;  L.have_pixel = (f & 2u) != 0u;  L.finished = (f & 4u) != 0u;
    L.pend = (f & 8u) != 0u;
    hit_t = __uint_as_float(r[25 * kCtaThreads]);
    hit_index = (int)r[26 * kCtaThreads];
    L.time = __uint_as_float(r[27 * kCtaThreads]);
    return live;
}

// =====================================================================================================
// One path per lane; the whole pre-filter image stays in shared memory for the life of the CTA (one TMA bulk copy).
// The lane's `pend` flag and ray.time sit in shared-memory slots across the sweep (the kernel runs at 78 of the 80
// registers that three CTAs per SM allow) and travel with the path through the regroup.
// =====================================================================================================
struct RegroupSmem {
    float4* pf;
    PerlinSmem* P;
    uint32_t* queue;           // this lane's candidate queue: [kQueueCap][kCtaThreads]
    uint32_t* queue_base;      // the queue array without the lane offset (the tensor-path drain reads the quad's queues)
    volatile uint32_t* pend;   // this lane's slot
    volatile float* tslot;     // this lane's slot
    uint32_t* xchg;            // [kRegroupWords][kCtaThreads]
    uint32_t* cat_count;       // two sets of kRegroupCats counters, 8 words apart
    float4* exact;             // EXACT_SMEM kernels: the exact blocks (stage 2 + shading, 16 B per sphere), behind the counters,
                               // followed by one category byte per stored sphere (n_blocks * 4 bytes)
    // image_bytes: n_blocks * 64 (FP32 pre-filter image) or (n_steps + 1) * 512 (tensor-path fragment image)
    __device__ __forceinline__ RegroupSmem(unsigned char* raw, size_t image_bytes) {
        pf = reinterpret_cast<float4*>(raw);
        P = reinterpret_cast<PerlinSmem*>(raw + image_bytes);
        queue_base = reinterpret_cast<uint32_t*>(P + 1);
        queue = queue_base + threadIdx.x;
        pend = queue + kQueueCap * kCtaThreads;
        tslot = reinterpret_cast<volatile float*>(pend + kCtaThreads);
        xchg = reinterpret_cast<uint32_t*>(P + 1) + (kQueueCap + 2) * kCtaThreads;
        cat_count = xchg + kRegroupWords * kCtaThreads;
        exact = reinterpret_cast<float4*>(cat_count + 16 * PT_REGROUP_DOMAINS);
    }
};
template <bool MMA>
__device__ __forceinline__ size_t regroup_image_bytes(const KernelArgs& a) {
    return MMA ? (size_t)(a.n_steps + 1) * 512 : (size_t)a.n_blocks * 64;
}
template <bool MMA, bool EXACT_SMEM = false>
__device__ __forceinline__ void regroup_stage(const KernelArgs& a, const RegroupSmem& sm, uint64_t* bar) {
    *sm.pend = 0u;
    *sm.tslot = 0.0f;
    if (threadIdx.x < 16 * PT_REGROUP_DOMAINS) sm.cat_count[threadIdx.x] = 0u;
    const uint32_t bytes = (uint32_t)regroup_image_bytes<MMA>(a);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && bytes != 0u) {
        const uint32_t exact_bytes = EXACT_SMEM ? (uint32_t)a.n_blocks * 64u : 0u;
        mbar_arrive_expect_tx(bar, bytes + exact_bytes);
        tma_bulk_g2s_chunked(sm.pf, MMA ? reinterpret_cast<const void*>(a.mma_image) : reinterpret_cast<const void*>(a.prefilter), bytes, bar);
        if (exact_bytes != 0u) tma_bulk_g2s_chunked(sm.exact, a.blocks, exact_bytes, bar);
    }
    stage_perlin(a, sm.P);
    if (EXACT_SMEM) {  // what the regroup sorts by, one byte per stored sphere (padding spheres are never hit)
        uint8_t* tab = reinterpret_cast<uint8_t*>(sm.exact + (size_t)a.n_blocks * 4);
        for (int i = (int)threadIdx.x; i < a.n_blocks * 4; i += (int)blockDim.x) tab[i] = i < a.n_spheres ? (uint8_t)sphere_category(a, i) : (uint8_t)CAT_ENDING;
    }
    __syncthreads();
    if (bytes != 0u) mbar_wait(bar, 0);
}
// the sweep phase of one trip for the lane's ray (already replaced by the parked ray for lanes without a path in flight)
// MMA: stage 1 on the tensor path (pt_sweep_mma.cuh); the ray fragments pass through the warp's own 32 columns of the
// exchange buffer, which nobody else touches between this warp's read-back in cta_regroup and the next trip's first barrier.
// `active`: the lane has a path in flight (its ray is the parked ray otherwise).
// EXACT_SMEM: stage 2 reads the exact blocks from shared memory (tensor-path kernel, where the loop leaves the LSU nearly idle:
// cfg2 +1.8 %, cfg4 +1.3 %; chosen by the host when it does not cost a resident CTA).
template <bool MOTION, bool MMA, bool EXACT_SMEM = false>
__device__ __forceinline__ void regroup_sweep(const KernelArgs& a, const RegroupSmem& sm, const MotionCtx& mc, bool active, float ox, float oy, float oz, float dx,
                                              float dy, float dz, float& hit_t, int& hit_index, unsigned& flagged) {
    hit_t = kMaxT;
    hit_index = -1;
    if (MMA) {
        int ovf_step;
        const int n = sweep_mma(reinterpret_cast<const uint4*>(sm.pf), a.n_steps, sm.xchg, sm.queue, a.mma, active, ox, oy, oz, dx, dy, dz, ovf_step);
        sweep_mma_drain<MOTION>(EXACT_SMEM ? sm.exact : a.blocks, mc, sm.queue_base, n, ovf_step, 0, a.n_steps, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
        return;
    }
    const float nod = -((ox * dx + oy * dy) + oz * dz);
    const float oo = ((ox * ox + oy * oy) + oz * oz) * (1.0f - kSlack);
    int cnt = 0;
    sweep_expanded<false, MOTION>(sm.pf, a.n_blocks, 0, a.blocks, mc, sm.queue, cnt, ox, oy, oz, dx, dy, dz, nod, ox + ox, oy + oy, oz + oz, oo, hit_t, hit_index, flagged);
    sweep_drain<MOTION, true>(a.blocks, mc, sm.queue, cnt, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
}

template <bool MOTION, bool MMA, bool EXACT_SMEM = false>
__global__ void __launch_bounds__(kCtaThreads) pt_megakernel_regroup(const __grid_constant__ KernelArgs a) {
    static_assert(MMA || !EXACT_SMEM, "the exact blocks move to shared memory only in the tensor-path kernel");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const RegroupSmem sm(smem_raw, regroup_image_bytes<MMA>(a));
    regroup_stage<MMA, EXACT_SMEM>(a, sm, &bar);
    const MotionCtx mc{a.motion, const_cast<const float*>(sm.tslot), a.order};

    const unsigned lane_id = threadIdx.x & 31u;
    Lane L;
    lane_init(L);
    unsigned long long rays = 0ULL;
    unsigned sweeps = 0u;
    lane_refill<MOTION>(a, L, lane_id);
    *sm.pend = L.pend ? 1u : 0u;
    *sm.tslot = L.time;
    for (uint32_t trip = 0;; ++trip) {
        float ox = L.o.x, oy = L.o.y, oz = L.o.z, dx = L.d.x, dy = L.d.y, dz = L.d.z;
        if (!L.active) {  // parked lane: |o|^2 = 1e36 dwarfs every L, d = 0 -> never a candidate
            ox = 0.0f; oy = 1.0e18f; oz = 0.0f;
            dx = dy = dz = 0.0f;
        }
        float hit_t = kMaxT;
        int hit_index = -1;
        __syncwarp();
        if (__any_sync(kFullMask, L.active)) {  // regrouping collects idle lanes in whole warps: they skip the sweep
            sweeps += 1u;
            unsigned flagged = 0u;
            regroup_sweep<MOTION, MMA, EXACT_SMEM>(a, sm, mc, L.active, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
        }
        __syncwarp();
        L.pend = *sm.pend != 0u;  // (parked in shared memory across the sweep)
        L.time = *sm.tslot;
        // (PT_REGROUP_PERIOD > 1, an experiment knob: regroup only every n-th trip; between two regroups a path is shaded by the lane that swept it)
        if (PT_REGROUP_PERIOD == 1 || trip % PT_REGROUP_PERIOD == 0u)
            if (!cta_regroup(a, L, hit_t, hit_index, sm.xchg, sm.cat_count + ((trip / PT_REGROUP_PERIOD) & 1u) * 8u, lane_id,
                             EXACT_SMEM ? reinterpret_cast<const uint8_t*>(sm.exact + (size_t)a.n_blocks * 4) : nullptr))
                break;
        if (L.active) {
            rays += 1ULL;  // scene.rs:57
            lane_shade<MOTION>(a, L, EXACT_SMEM ? sm.exact : a.blocks, *sm.P, mc, hit_t, hit_index);
        }
        lane_refill<MOTION>(a, L, lane_id);  // ended paths sit side by side now: next sample / next ticket together
        *sm.pend = L.pend ? 1u : 0u;
        *sm.tslot = L.time;
        __syncwarp();
    }
    flush_ray_count(a, rays, lane_id, sweeps);
}

// pt_debug_hits for scenes this kernel renders: caller-supplied rays through regroup_sweep (same staging, operands, queue,
// re-tests)
template <bool MOTION, bool MMA>
__global__ void __launch_bounds__(kCtaThreads) pt_debug_hits_regroup(const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const RegroupSmem sm(smem_raw, regroup_image_bytes<MMA>(a));
    regroup_stage<MMA>(a, sm, &bar);
    const MotionCtx mc{a.motion, const_cast<const float*>(sm.tslot), a.order};
    for (uint32_t base = blockIdx.x * kCtaThreads; base < a.dbg_n; base += gridDim.x * kCtaThreads) {
        const uint32_t i = base + threadIdx.x;
        float ox = 0.0f, oy = 1.0e18f, oz = 0.0f, dx = 0.0f, dy = 0.0f, dz = 0.0f, time = 0.0f;
        if (i < a.dbg_n) {
            const float* ray = a.dbg_rays + (size_t)i * 6;
            ox = ray[0]; oy = ray[1]; oz = ray[2]; dx = ray[3]; dy = ray[4]; dz = ray[5];
            if (a.dbg_times) time = a.dbg_times[i];
        }
        *sm.tslot = time;
        __syncwarp();
        float hit_t;
        int hit_index;
        unsigned flagged = 0u;
        regroup_sweep<MOTION, MMA>(a, sm, mc, i < a.dbg_n, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
        if (i < a.dbg_n) {
            a.dbg_idx[i] = hit_index < 0 ? -1 : (a.order ? (int32_t)__ldg(a.order + hit_index) : hit_index);
            a.dbg_t[i] = hit_t;
            if (a.dbg_flagged) a.dbg_flagged[i] = flagged;
        }
    }
}

}  // namespace pt
