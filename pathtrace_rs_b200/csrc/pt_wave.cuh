// pt_wave.cuh — the wavefront form of the resident kernel: one persistent CTA per SM, a pool of paths in shared memory,
// warps as asynchronous workers that pull JOBS from per-category queues.
//
// Why.  ncu of the two-paths-per-lane kernel (profiles/ncu_r2_v1_cfg2_*): the sweep loop is 58 % of the executed
// instructions and runs with 32 of 32 lanes, everything else runs with 12.5 of 32 — after the sweep the 32 paths of a warp
// want five different pieces of code (Lambertian, textured Lambertian, metal, dielectric, path end + next camera ray) and
// the warp executes all of them.  Per 64 rays the kernel needs ~9 900 issue slots where the FMA-bound sweep needs ~9 200 clk:
// it is ISSUE-bound by divergent shading.  Round 1 sorted the CTA's paths between two CTA barriers per trip and lost to the
// barriers what the sort won; one barrier per trip in the two-path kernel costs 14 % (variant `sync1`).
//
// Here nothing waits for anything.  A path is a 112-byte record in the CTA's pool; a queue entry is a 16-bit slot index.
//   SWEEP job   64 paths from the needs-sweep queue: sweep_two (uniform sphere operands, two rays per lane), exact re-tests,
//               nearest hit into the record, then every path is filed under what it does next —
//   SHADE job   32 paths of ONE category: lane_shade + lane_refill run convergent; paths with a new ray go back to the
//               needs-sweep queue, paths whose pixel/ticket supply ended retire.
// Queues are rings in shared memory guarded by one CTA-wide spin lock that lane 0 of a warp holds for a few dozen
// instructions per job (a 64-ray sweep is ~10 000 clk).  A warp that finds no full batch takes the fullest partial one
// unless other warps are still busy (they will refill the queues), so the CTA can never stall with work outstanding.
// Every path still consumes exactly its own pixel's RNG stream and the same arithmetic: images are bit-identical.
#pragma once
#include "pt_megakernel.cuh"

namespace pt {

#ifndef PT_WAVE_THREADS
#define PT_WAVE_THREADS 640
#endif
constexpr int kWaveThreads = PT_WAVE_THREADS;
constexpr int kWaveWarps = kWaveThreads / 32;
constexpr int kWaveRecWords = 28;  // 7 x 16 bytes: the 112-byte stride spreads LDS.128 of random slots over all banks
constexpr int kWaveMaxPool = 2048 - 64;
constexpr int kWaveCandCap = 8;    // candidate-queue entries per lane, shared by its two rays (overflow: sweep_overflow)
enum { WQ_SWEEP = 0, WQ_END = 1, WQ_LAMBERT = 2, WQ_LAMBERT_TEX = 3, WQ_METAL = 4, WQ_DIELECTRIC = 5, kWaveQueues = 6 };
enum { WJ_NONE = -1, WJ_EXIT = -2 };

// record: r0,r1 = generator | r2 = origin, hit_t | r3 = direction, hit_index | r4 = throughput, ray.time |
//         r5 = colour sum, px | r6 = py, sample, depth, flags
struct WaveCtl {
    unsigned lock;
    unsigned head[kWaveQueues];
    unsigned count[kWaveQueues];
    int live;  // paths that have not retired
    int busy;  // warps currently holding a job
};

struct WaveSmem {
    float4* kplane;
    PerlinSmem* P;
    uint32_t* cand;    // this lane's candidate queue: [kWaveCandCap][kWaveThreads]
    uint16_t* queues;  // [kWaveQueues][cap]
    uint4* pool;       // [pool_paths][7]
    uint32_t cap;      // ring capacity of every job queue (>= pool_paths + 64)
    __device__ __forceinline__ WaveSmem(unsigned char* raw, const KernelArgs& a) {
        const uint32_t image_bytes = ((uint32_t)a.n_blocks * 16u + 127u) & ~127u;
        kplane = reinterpret_cast<float4*>(raw);
        P = reinterpret_cast<PerlinSmem*>(raw + image_bytes);
        cand = reinterpret_cast<uint32_t*>(P + 1) + threadIdx.x;
        cap = (uint32_t)a.wave_pool + 64u;
        queues = reinterpret_cast<uint16_t*>(reinterpret_cast<uint32_t*>(P + 1) + kWaveCandCap * kWaveThreads);
        const uintptr_t after = reinterpret_cast<uintptr_t>(queues + (size_t)kWaveQueues * cap);
        pool = reinterpret_cast<uint4*>((after + 15) & ~(uintptr_t)15);
    }
    __device__ __forceinline__ uint4* rec(uint32_t slot) const { return pool + (size_t)slot * 7; }
    __device__ __forceinline__ uint16_t* queue(int q) const { return queues + (size_t)q * cap; }
};

// The lock is taken by the WARP, and no lane does anything the others do not.  ptxas keeps the sweep's sphere operands in
// uniform registers only while it can prove that the warp reaches the sweep converged, and (measured, CUDA 12.9) it gives
// that up as soon as a lane-dependent branch with a side effect — `if (lane == 0) atomicCAS(..)`, `if (lane == 0) ctl->x = ..`
// — sits on the path from the top of the job loop to the sweep; the loop then falls back to LDC into vector registers and
// runs 1.7x slower.  So every lane executes the same instructions: the compare-and-swap of lanes 1..31 compares against a
// value the lock never holds, the release is an AND with all ones for them, and the queue bookkeeping inside the critical
// section is computed and stored redundantly by all 32 lanes (same values, same addresses).
__device__ __forceinline__ void wave_lock(volatile WaveCtl* ctl, unsigned lane_id) {
    for (unsigned spins = 0u;; ++spins) {
        const unsigned old = atomicCAS(const_cast<unsigned*>(&ctl->lock), lane_id == 0u ? 0u : 0xffffffffu, 1u);
        if (__ballot_sync(kFullMask, lane_id == 0u && old == 0u) != 0u) break;
        if (spins > (1u << 26)) __trap();  // watchdog: a lost lock must end the launch with an error, not hang the GPU
        __nanosleep(32);
    }
    __threadfence_block();
}
__device__ __forceinline__ void wave_unlock(volatile WaveCtl* ctl, unsigned lane_id) {
    __threadfence_block();
    __syncwarp();
    atomicAnd(const_cast<unsigned*>(&ctl->lock), lane_id == 0u ? 0u : 0xffffffffu);
}

__device__ __forceinline__ void wave_store(uint4* r, const Lane& L) {
    r[0] = make_uint4((uint32_t)L.rng.s0, (uint32_t)(L.rng.s0 >> 32), (uint32_t)L.rng.s1, (uint32_t)(L.rng.s1 >> 32));
    r[1] = make_uint4((uint32_t)L.rng.s2, (uint32_t)(L.rng.s2 >> 32), (uint32_t)L.rng.s3, (uint32_t)(L.rng.s3 >> 32));
    r[2] = make_uint4(__float_as_uint(L.o.x), __float_as_uint(L.o.y), __float_as_uint(L.o.z), 0u);
    r[3] = make_uint4(__float_as_uint(L.d.x), __float_as_uint(L.d.y), __float_as_uint(L.d.z), 0u);
    r[4] = make_uint4(__float_as_uint(L.thr.x), __float_as_uint(L.thr.y), __float_as_uint(L.thr.z), __float_as_uint(L.time));
    r[5] = make_uint4(__float_as_uint(L.col.x), __float_as_uint(L.col.y), __float_as_uint(L.col.z), L.px);
    r[6] = make_uint4(L.py, L.sample, L.depth, (L.active ? 1u : 0u) | (L.have_pixel ? 2u : 0u) | (L.finished ? 4u : 0u) | (L.pend ? 8u : 0u));
}
__device__ __forceinline__ void wave_load(const uint4* r, Lane& L, float& hit_t, int& hit_index) {
    const uint4 a = r[0], b = r[1], c = r[2], d = r[3], e = r[4], f = r[5], g = r[6];
    L.rng.s0 = (uint64_t)a.x | ((uint64_t)a.y << 32);
    L.rng.s1 = (uint64_t)a.z | ((uint64_t)a.w << 32);
    L.rng.s2 = (uint64_t)b.x | ((uint64_t)b.y << 32);
    L.rng.s3 = (uint64_t)b.z | ((uint64_t)b.w << 32);
    L.o = v3(__uint_as_float(c.x), __uint_as_float(c.y), __uint_as_float(c.z));
    hit_t = __uint_as_float(c.w);
    L.d = v3(__uint_as_float(d.x), __uint_as_float(d.y), __uint_as_float(d.z));
    hit_index = (int)d.w;
    L.thr = v3(__uint_as_float(e.x), __uint_as_float(e.y), __uint_as_float(e.z));
    L.time = __uint_as_float(e.w);
    L.col = v3(__uint_as_float(f.x), __uint_as_float(f.y), __uint_as_float(f.z));
    L.px = f.w;
    L.py = g.x;  L.sample = g.y;  L.depth = g.z;
    L.active = (g.w & 1u) != 0u;  L.have_pixel = (g.w & 2u) != 0u;  L.finished = (g.w & 4u) != 0u;  L.pend = (g.w & 8u) != 0u;
}

// what a path does after its sweep (material.rs:138-159 by kind; a miss, a light or the depth limit end the path)
__device__ __forceinline__ int wave_category(const KernelArgs& a, int hit_index, uint32_t depth) {
    if (hit_index < 0 || depth >= a.max_depth) return WQ_END;
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(a.shade + hit_index) + 1);
    const int kind = __float_as_int(s1.y);
    if (kind == MAT_LAMBERTIAN) return __float_as_int(s1.z) < 0 ? WQ_LAMBERT : WQ_LAMBERT_TEX;
    if (kind == MAT_METAL) return WQ_METAL;
    if (kind == MAT_DIELECTRIC) return WQ_DIELECTRIC;
    return WQ_END;  // DiffuseLight
}

// Take a job (warp-uniform result: every lane computes it): .x = queue to serve (or WJ_EXIT), .y = entries taken, .z = ring
// position of the first one.  OUT OF LINE on purpose: with these spin loops inlined into the job loop ptxas stops keeping the
// sweep's sphere operands in uniform registers (it can no longer prove that the warp reaches the sweep converged; measured
// with CUDA 12.9: LDC into vector registers instead of LDCU, the loop 1.7x slower — tests/test_host_and_abi.py checks the
// SASS).  A call is a convergence point it does understand.
__device__ __noinline__ int4 wave_get_job(volatile WaveCtl* ctl, unsigned cap, unsigned lane_id, bool was_busy) {
    unsigned tries = 0u;
    int4 job = make_int4(WJ_NONE, 0, 0, 0);
    for (unsigned spins = 0;; ++spins) {
        const unsigned old = atomicCAS(const_cast<unsigned*>(&ctl->lock), lane_id == 0u ? 0u : 0xffffffffu, 1u);
        if (__ballot_sync(kFullMask, lane_id == 0u && old == 0u) != 0u) {
            __threadfence_block();
            // fullest shading queue
            int best = WQ_END;
            unsigned best_n = ctl->count[WQ_END];
#pragma unroll
            for (int q = WQ_END + 1; q < kWaveQueues; ++q) {
                const unsigned c = ctl->count[q];
                if (c > best_n) {
                    best_n = c;
                    best = q;
                }
            }
            const unsigned ns = ctl->count[WQ_SWEEP];
            int busy = ctl->busy - (was_busy ? 1 : 0);
            was_busy = false;
            int pick = WJ_NONE;
            if (best_n >= 32u) pick = best;                 // a full, convergent shading batch: short, and it feeds the sweep queue
            else if (ns >= 64u) pick = WQ_SWEEP;            // a full sweep
            else if (ns + best_n != 0u && (busy == 0 || tries >= 4u)) pick = (ns * 32u >= best_n * 64u) ? WQ_SWEEP : best;  // nobody will add to the queues, or waited long enough: the fuller partial batch
            else if (ns + best_n == 0u && busy == 0 && ctl->live == 0) pick = WJ_EXIT;
            // (unconditional stores: `pick` is warp-uniform, but ptxas cannot know that, and a store under a branch it takes
            // for divergent costs the sweep its uniform registers)
            const int qi = pick >= 0 ? pick : 0;
            const unsigned have = ctl->count[qi];
            const unsigned h = ctl->head[qi];
            const int cnt = pick >= 0 ? (int)min(have, pick == WQ_SWEEP ? 64u : 32u) : 0;
            busy += pick >= 0 ? 1 : 0;
            __syncwarp();  // all lanes have read the old values
            ctl->head[qi] = (h + (unsigned)cnt) % cap;
            ctl->count[qi] = have - (unsigned)cnt;
            ctl->busy = busy;
            wave_unlock(ctl, lane_id);
            job = make_int4(pick, cnt, (int)h, 0);
            tries += 1u;
            if (__ballot_sync(kFullMask, pick != WJ_NONE) != 0u) break;
            __nanosleep(200);
        } else {
            __nanosleep(32);
        }
        if (spins > (1u << 26)) __trap();  // watchdog (seconds): a lost lock, or queues empty with nobody busy and paths still alive
    }
    return job;
}

// File this warp's paths: lane-level (slot, queue) pairs, up to two per lane (queue < 0: nothing).  One critical section.
// Out of line for the same reason as wave_get_job.
__device__ __noinline__ void wave_push2(volatile WaveCtl* ctl, const WaveSmem& sm, unsigned lane_id, int q0, uint32_t slot0, int q1, uint32_t slot1, int retired) {
    unsigned b0[kWaveQueues], b1[kWaveQueues];
#pragma unroll
    for (int q = 0; q < kWaveQueues; ++q) {
        b0[q] = __ballot_sync(kFullMask, q0 == q);
        b1[q] = __ballot_sync(kFullMask, q1 == q);
    }
    wave_lock(ctl, lane_id);
    unsigned tail = 0u;
    if (lane_id < (unsigned)kWaveQueues) tail = (ctl->head[lane_id] + ctl->count[lane_id]) % sm.cap;
    const unsigned below = (1u << lane_id) - 1u;
#pragma unroll
    for (int q = 0; q < kWaveQueues; ++q) {
        const unsigned n0 = (unsigned)__popc(b0[q]), n1 = (unsigned)__popc(b1[q]);
        if (n0 + n1 == 0u) continue;  // warp-uniform
        const unsigned t = __shfl_sync(kFullMask, tail, q);
        uint16_t* ring = sm.queue(q);
        if (q0 == q) ring[(t + (unsigned)__popc(b0[q] & below)) % sm.cap] = (uint16_t)slot0;
        if (q1 == q) ring[(t + n0 + (unsigned)__popc(b1[q] & below)) % sm.cap] = (uint16_t)slot1;
        if (lane_id == 0u) ctl->count[q] += n0 + n1;
    }
    if (lane_id == 0u && retired != 0) ctl->live -= retired;
    wave_unlock(ctl, lane_id);
}

// Second half of a SWEEP job, out of line (see wave_get_job): exact re-tests of the flagged spheres of the lane's two rays
// (their origin and direction are read back from the records), nearest hits into the records, every path filed under what
// it does next.  slot < 0: the lane has no path in that row.
template <bool MOTION>
__device__ __noinline__ void wave_sweep_finish(const KernelArgs& a, const WaveSmem& sm, volatile WaveCtl* ctl, unsigned lane_id, int slot0, int slot1, int cnt0,
                                               int cnt1, int overflow0, int overflow1) {
    int cat[2] = {-1, -1};
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
        const int slot = r ? slot1 : slot0;
        if (slot < 0) continue;
        uint32_t* w = reinterpret_cast<uint32_t*>(sm.rec((uint32_t)slot));
        const float ox = __uint_as_float(w[8]), oy = __uint_as_float(w[9]), oz = __uint_as_float(w[10]);
        const float dx = __uint_as_float(w[12]), dy = __uint_as_float(w[13]), dz = __uint_as_float(w[14]);
        const MotionCtx mc{a.motion, reinterpret_cast<const float*>(w + 19), a.order};  // ray.time of the path
        float hit_t = kMaxT;
        int hit_index = -1;
        unsigned flagged = 0u;
        const int cnt = r ? cnt1 : cnt0;
        sweep_drain_range<MOTION, kWaveThreads>(a.blocks, mc, sm.cand, r ? kWaveCandCap - cnt : 0, cnt, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
        const int first = r ? overflow1 : overflow0;
        if (first < a.n_blocks) sweep_overflow<MOTION>(a.blocks, mc, first, a.n_blocks, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
        w[11] = __float_as_uint(hit_t);
        w[15] = (uint32_t)hit_index;
        const int c = wave_category(a, hit_index, w[26]);
        if (r) cat[1] = c; else cat[0] = c;
    }
    __syncwarp();
    wave_push2(ctl, sm, lane_id, cat[0], (uint32_t)max(slot0, 0), cat[1], (uint32_t)max(slot1, 0), 0);
}

// SHADE job: up to 32 paths of ONE category (queue `kind`), convergent.  Returns the number of rays traced (0 or 1 per lane).
// Out of line: the job loop stays small enough for ptxas to keep the sweep in uniform registers (see wave_get_job).
template <bool MOTION>
__device__ __noinline__ unsigned wave_shade_job(const KernelArgs& a, const WaveSmem& sm, volatile WaveCtl* ctl, unsigned lane_id, int kind, int n, unsigned head) {
    const bool valid = lane_id < (unsigned)n;
    const uint32_t slot = valid ? sm.queue(kind)[(head + lane_id) % sm.cap] : 0u;
    Lane L;
    lane_init(L);
    L.finished = true;  // lanes without a path take part in the collectives of lane_refill and nothing else
    float hit_t = kMaxT;
    int hit_index = -1;
    unsigned rays = 0u;
    if (valid) wave_load(sm.rec(slot), L, hit_t, hit_index);
    if (L.active) {
        rays = 1u;  // scene.rs:57
        const MotionCtx mc{a.motion, nullptr, a.order};
        lane_shade<MOTION>(a, L, a.blocks, *sm.P, mc, hit_t, hit_index);
    }
    lane_refill<MOTION>(a, L, lane_id);
    int q = -1, retired = 0;
    if (valid) {
        wave_store(sm.rec(slot), L);
        if (L.active) q = WQ_SWEEP;          // a new ray (next bounce, or the next sample's camera ray)
        else if (!L.finished) q = WQ_END;    // holds a ticket whose predecessor chunk is not published yet: ask again
        else retired = 1;                    // no tickets left
    }
    const int n_retired = __popc(__ballot_sync(kFullMask, retired != 0));
    wave_push2(ctl, sm, lane_id, q, slot, -1, 0u, n_retired);
    return rays;
}

template <bool MOTION>
__global__ void __launch_bounds__(kWaveThreads, 1) pt_megakernel_wave(const __grid_constant__ KernelArgs a, const __grid_constant__ ConstImageT<true> ci) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ WaveCtl ctl;
    const WaveSmem sm(smem_raw, a);
    const uint32_t image_bytes = (uint32_t)a.n_blocks * 16u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        ctl.lock = 0u;
        for (int q = 0; q < kWaveQueues; ++q) ctl.head[q] = ctl.count[q] = 0u;
        ctl.count[WQ_END] = (unsigned)a.wave_pool;  // every slot starts as a path that needs its first ticket
        ctl.live = a.wave_pool;
        ctl.busy = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0 && image_bytes != 0u) {
        mbar_arrive_expect_tx(&bar, image_bytes);
        tma_bulk_g2s_chunked(sm.kplane, a.kplane, image_bytes, &bar);
    }
    stage_perlin(a, sm.P);
    {
        Lane L;
        lane_init(L);
        for (uint32_t slot = threadIdx.x; slot < (uint32_t)a.wave_pool; slot += kWaveThreads) {
            wave_store(sm.rec(slot), L);
            sm.queue(WQ_END)[slot] = (uint16_t)slot;
        }
    }
    __syncthreads();
    if (image_bytes != 0u) mbar_wait(&bar, 0);

    const unsigned lane_id = threadIdx.x & 31u;
    unsigned long long rays = 0ULL;
    unsigned sweeps = 0u;
    int4 job = wave_get_job(&ctl, sm.cap, lane_id, false);
    for (;;) {
        const int kind = job.x, n = job.y;
        const unsigned head = (unsigned)job.z;
        // branch on VOTES: ptxas keeps the sweep's sphere operands in uniform registers only inside control flow it can prove
        // warp-uniform.  The next job is fetched at the BOTTOM of the loop for the same reason (see wave_lock).
        if (__ballot_sync(kFullMask, kind == WJ_EXIT) != 0u) break;
        if (__ballot_sync(kFullMask, kind == WQ_SWEEP) != 0u) {
            // ---- 64 rays: lane l carries entries l and l + 32 ----
            float ox[2], oy[2], oz[2], dx[2], dy[2], dz[2];
            uint32_t slot[2];
            bool valid[2];
            const uint16_t* ring = sm.queue(WQ_SWEEP);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const unsigned i = lane_id + 32u * (unsigned)r;
                valid[r] = i < (unsigned)n;
                slot[r] = valid[r] ? ring[(head + i) % sm.cap] : 0u;
                ox[r] = 0.0f; oy[r] = 1.0e18f; oz[r] = 0.0f;  // parked ray: never a candidate
                dx[r] = dy[r] = dz[r] = 0.0f;
                if (valid[r]) {
                    const uint4 c = sm.rec(slot[r])[2], d = sm.rec(slot[r])[3];
                    ox[r] = __uint_as_float(c.x); oy[r] = __uint_as_float(c.y); oz[r] = __uint_as_float(c.z);
                    dx[r] = __uint_as_float(d.x); dy[r] = __uint_as_float(d.y); dz[r] = __uint_as_float(d.z);
                }
            }
            sweeps += 1u + (n > 32 ? 1u : 0u);
            float o2x[2], o2y[2], o2z[2], nod[2], oo[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                nod[r] = -((ox[r] * dx[r] + oy[r] * dy[r]) + oz[r] * dz[r]);
                oo[r] = ((ox[r] * ox[r] + oy[r] * oy[r]) + oz[r] * oz[r]) * (1.0f - kSlack);
                o2x[r] = ox[r] + ox[r]; o2y[r] = oy[r] + oy[r]; o2z[r] = oz[r] + oz[r];
            }
            int cnt0 = 0, cnt1 = 0;
            int overflow[2] = {a.n_blocks, a.n_blocks};
            sweep_two<true, kWaveThreads, kWaveCandCap>(ci, sm.kplane, a.n_blocks, sm.cand, cnt0, cnt1, dx, dy, dz, o2x, o2y, o2z, nod, oo, overflow);
            wave_sweep_finish<MOTION>(a, sm, &ctl, lane_id, valid[0] ? (int)slot[0] : -1, valid[1] ? (int)slot[1] : -1, cnt0, cnt1, overflow[0], overflow[1]);
        } else {
            rays += wave_shade_job<MOTION>(a, sm, &ctl, lane_id, kind, n, head);
        }
        job = wave_get_job(&ctl, sm.cap, lane_id, true);
    }
    flush_ray_count(a, rays, lane_id, sweeps);
}

}  // namespace pt
