// pt_wave.cuh — the wavefront form of the resident kernel: one persistent CTA per SM, a pool of paths in shared memory,
// warps that trade paths through per-material queues so that every warp SHADES 32 paths of one kind.
//
// Why.  ncu of the two-paths-per-lane kernel (profiles/ncu_r2_v1_cfg2_*): the sweep loop is 58 % of the executed
// instructions and runs with 32 of 32 lanes, everything else runs with 12.5 of 32 — after the sweep the 32 paths of a warp
// want five different pieces of code (Lambertian, textured Lambertian, metal, dielectric, path end + next camera ray) and
// the warp executes all of them.  Per 64 rays the kernel needs ~9 900 issue slots where the FMA-bound sweep needs ~9 200 clk:
// it is ISSUE-bound by divergent shading.  Round 1 sorted the CTA's paths between two CTA barriers per trip and lost to the
// barriers what the sort won; one barrier per trip in the two-path kernel costs 14 % (variant `sync1`).
//
// Here nothing waits for anything.  A path is a 112-byte record in the CTA's pool; a queue entry is a 16-bit slot index.
// A warp's cycle:
//   SWEEP     its hand of <= 64 paths (two per lane): sweep_two with uniform sphere operands, exact re-tests, nearest hit
//             into the record, every path classified by what it does next;
//   EXCHANGE  one critical section: the 64 classified paths go into the category queues, two batches of <= 32 paths of ONE
//             category each come out (round-robin over the categories that have a full batch);
//   SHADE     each batch runs lane_shade + lane_refill convergent; the paths that have a ray afterwards are the next hand.
// The pool holds 64 paths per warp plus a few hundred in the queues, which is what makes full single-category batches
// available; one lock acquisition per warp per ~10 000-clk cycle keeps the lock idle.  Every path still consumes exactly its
// own pixel's RNG stream and the same arithmetic: images are bit-identical to the lockstep kernels'.
#pragma once
#include "pt_megakernel.cuh"

namespace pt {

#ifndef PT_WAVE_THREADS
#define PT_WAVE_THREADS 640
#endif
// PT_WAVE_LDS=1: the sweep reads its sphere pairs from a staged image in shared memory (LDS.128) instead of the
// kernel-parameter image through the uniform datapath
#ifndef PT_WAVE_LDS
#define PT_WAVE_LDS 0
#endif
constexpr bool kWaveConstImg = PT_WAVE_LDS == 0;
constexpr int kWaveThreads = PT_WAVE_THREADS;
constexpr int kWaveWarps = kWaveThreads / 32;
constexpr int kWaveRecWords = 28;  // 7 x 16 bytes: the 112-byte stride spreads LDS.128 of random slots over all banks
constexpr int kWaveMaxPool = 2048 - 64;
constexpr int kWaveCandCap = 8;    // candidate-queue entries per lane, shared by its two rays (overflow: sweep_overflow)
enum { WQ_END = 0, WQ_LAMBERT = 1, WQ_LAMBERT_TEX = 2, WQ_METAL = 3, WQ_DIELECTRIC = 4, kWaveQueues = 5 };

// record: r0,r1 = generator | r2 = origin, hit_t | r3 = direction, hit_index | r4 = throughput, ray.time |
//         r5 = colour sum, px | r6 = py, sample, depth, flags
struct WaveCtl {
    unsigned tail[kWaveQueues];   // ring positions handed to producers (reserved with atomicAdd, monotone, taken modulo cap)
    unsigned head[kWaveQueues];   // ring positions handed to consumers
    int avail[kWaveQueues];       // published entries not yet claimed (a consumer claims by subtracting; may dip below 0 briefly)
    unsigned rr;                  // round-robin start of the batch selection (racy on purpose: a hint)
    int live;                     // paths that have not retired
    unsigned abort;               // a watchdog fired: every warp leaves its loop, the host reports the launch as failed
    unsigned int* status;         // global words the watchdog reports to (KernelArgs.status)
    unsigned scratch[32];         // where lanes 1..31 aim their share of every atomic (see wave_atomic_add)
};
// Watchdogs.  The queues cannot deadlock by construction (a path is always in exactly one queue or in a warp's hand), but a
// bug here would hang the GPU, so every wait is bounded: after ~seconds the warp records why, raises `abort`, and the whole
// CTA drains out; the host turns a non-zero status word into PT_ERR_CUDA.
enum { WAVE_LOST_ENTRY = 1, WAVE_STARVED = 2 };
__device__ __forceinline__ void wave_abort(volatile WaveCtl* ctl, unsigned code) {
    if (atomicOr(ctl->status, code) == 0u) {  // first report: a snapshot of the queue state for the host's error message
        unsigned int* dbg = ctl->status;
        dbg[1] = (unsigned)ctl->live;
        dbg[2] = 0u;
        for (int q = 0; q < kWaveQueues; ++q) dbg[3 + q] = (unsigned)ctl->avail[q];
        dbg[8] = blockIdx.x;
        dbg[9] = threadIdx.x;
    }
    ctl->abort = 1u;
}

struct WaveSmem {
    float4* kplane;
    PerlinSmem* P;
    uint32_t* cand;    // this lane's candidate queue: [kWaveCandCap][kWaveThreads]
    uint16_t* queues;  // [kWaveQueues][cap]
    uint4* pool;       // [pool_paths][7]
    uint32_t cap;      // ring capacity of every queue (>= pool_paths + 64)
    __device__ __forceinline__ WaveSmem(unsigned char* raw, const KernelArgs& a) {
        const uint32_t image_bytes = ((uint32_t)a.n_blocks * (kWaveConstImg ? 16u : 64u) + 127u) & ~127u;
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

// One atomic add on a control word on behalf of the warp; returns the old value to every lane.  No lane does anything the
// others do not: ptxas keeps the sweep's sphere operands in uniform registers only while it can prove that the warp reaches
// the sweep converged, and (measured, CUDA 12.9) it gives that up as soon as a lane-dependent branch with a side effect —
// `if (lane == 0) atomicAdd(..)` — or a spin loop that a single lane executes is reachable in the warp's loop, even inside
// an out-of-line callee; the sweep then falls back to LDC into vector registers and runs 1.7x slower.  So lanes 1..31 execute
// the same atomic on a scratch word of their own with an increment of zero.
__device__ __forceinline__ int wave_atomic_add(volatile WaveCtl* ctl, volatile void* word, int inc, unsigned lane_id) {
    int* p = const_cast<int*>(lane_id == 0u ? reinterpret_cast<volatile int*>(word) : reinterpret_cast<volatile int*>(&ctl->scratch[lane_id]));
    const int old = atomicAdd(p, lane_id == 0u ? inc : 0);
    return __shfl_sync(kFullMask, old, 0);
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

// EXCHANGE: file the warp's hand (lane-level (slot, category) pairs, two per lane; category < 0: nothing) and take two
// batches of up to 32 paths of one category each.  Returns {category A, n A, ring head A} in .x .y .z and B packed in .w as
// category | n << 8 | head << 16 (the rings hold fewer than 65 536 entries); .x = -2: everything has retired, leave.
//
// Lock-free.  A queue is a ring of 16-bit slot indices (0xFFFF = empty) with three counters.  A producer reserves ring
// positions with one atomicAdd on `tail`, writes its entries, fences, and publishes their NUMBER with an atomicAdd on
// `avail`.  A consumer claims a number with an atomicAdd(-n) on `avail` (giving it back if it overdrew), takes the positions
// with an atomicAdd on `head`, and — because two producers may publish out of order — waits for each of its positions to
// turn non-empty before reading it and marking it empty again; that wait is a handful of instructions of another warp.
// The ring is larger than the pool, so a reserved position has always been consumed.
//
// OUT OF LINE on purpose: with these loops inlined into the warp's loop ptxas stops keeping the sweep's sphere operands in
// uniform registers (measured with CUDA 12.9: LDC into vector registers instead of LDCU, the sweep 1.7x slower —
// tests/test_host_and_abi.py checks the SASS).  A call is a convergence point it does understand.
constexpr uint16_t kWaveEmpty = 0xFFFFu;
__device__ __noinline__ int4 wave_exchange(volatile WaveCtl* ctl, uint16_t* queues, unsigned cap, unsigned lane_id, int q0, uint32_t slot0, int q1, uint32_t slot1,
                                           int retired) {
    const unsigned below = (1u << lane_id) - 1u;
    // ---- push ----
#pragma unroll
    for (int q = 0; q < kWaveQueues; ++q) {
        const unsigned b0 = __ballot_sync(kFullMask, q0 == q), b1 = __ballot_sync(kFullMask, q1 == q);
        const unsigned n0 = (unsigned)__popc(b0), n1 = (unsigned)__popc(b1);
        if (n0 + n1 == 0u) continue;  // (a vote: warp-uniform)
        const unsigned t = (unsigned)wave_atomic_add(ctl, &ctl->tail[q], (int)(n0 + n1), lane_id);
        volatile uint16_t* ring = queues + (size_t)q * cap;
        volatile uint16_t* e0 = ring + (t + (unsigned)__popc(b0 & below)) % cap;
        volatile uint16_t* e1 = ring + (t + n0 + (unsigned)__popc(b1 & below)) % cap;
        // a reserved position is normally long free (the ring is larger than the pool); if its last consumer has claimed it
        // but not emptied it yet, wait for that
        for (unsigned spins = 0u;; ++spins) {
            const bool busy = (q0 == q && *e0 != kWaveEmpty) || (q1 == q && *e1 != kWaveEmpty);
            if (__ballot_sync(kFullMask, busy) == 0u) break;
            if (spins > (1u << 24)) wave_abort(ctl, WAVE_LOST_ENTRY);
            if (ctl->abort != 0u) break;
        }
        if (q0 == q) *e0 = (uint16_t)slot0;
        if (q1 == q) *e1 = (uint16_t)slot1;
        __threadfence_block();
        __syncwarp();
        wave_atomic_add(ctl, &ctl->avail[q], (int)(n0 + n1), lane_id);
    }
    if (retired != 0) wave_atomic_add(ctl, &ctl->live, -retired, lane_id);
    // ---- pop two batches: a category with a full batch, round-robin; else the fullest one ----
    int cat[2] = {-1, -1};
    unsigned n[2] = {0u, 0u}, h[2] = {0u, 0u};
    unsigned rr = ctl->rr;
    for (unsigned tries = 0u;; ++tries) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (n[k] != 0u) continue;
            int pick = -1, best = 0;
#pragma unroll
            for (int i = 0; i < kWaveQueues; ++i) {
                const int q = (int)((rr + (unsigned)i) % (unsigned)kWaveQueues);
                const int c = ctl->avail[q];
                const int score = c >= 32 ? 1000 - i : c;  // first full batch in round-robin order, else the fullest
                if (score > best) {
                    best = score;
                    pick = q;
                }
            }
            if (__ballot_sync(kFullMask, pick >= 0) == 0u) continue;
            const int seen = ctl->avail[pick];
            const int want = min(max(seen, 1), 32);
            const int old = wave_atomic_add(ctl, &ctl->avail[pick], -want, lane_id);
            if (old < want) {  // somebody else was faster: give it back and look again
                wave_atomic_add(ctl, &ctl->avail[pick], want, lane_id);
                continue;
            }
            cat[k] = pick;
            n[k] = (unsigned)want;
            h[k] = (unsigned)wave_atomic_add(ctl, &ctl->head[pick], want, lane_id) % cap;
            rr = (unsigned)pick + 1u;
        }
        if (__ballot_sync(kFullMask, n[0] + n[1] != 0u) != 0u) break;
        // empty-handed: either everything has retired, or the remaining paths are in other warps' hands — wait for them
        if (ctl->live == 0 || ctl->abort != 0u) {  // (`live` only ever falls: 0 is final)
            cat[0] = -2;
            break;
        }
        if (tries > (1u << 21)) wave_abort(ctl, WAVE_STARVED);
        __nanosleep(1000);
    }
    ctl->rr = rr % (unsigned)kWaveQueues;
    return make_int4(cat[0], (int)n[0], (int)h[0], (cat[1] & 0xff) | (int)(n[1] << 8) | (int)(h[1] << 16));
}

// the slot index in ring position (head + lane) of queue `cat`: waits for the producer's write, then frees the position
__device__ __forceinline__ uint32_t wave_take_entry(volatile WaveCtl* ctl, uint16_t* queues, unsigned cap, int cat, unsigned head, bool valid, unsigned lane_id) {
    volatile uint16_t* e = queues + (size_t)cat * cap + (head + lane_id) % cap;
    uint16_t v = kWaveEmpty;
    for (unsigned spins = 0u;; ++spins) {
        if (valid && v == kWaveEmpty) v = *e;
        if (__ballot_sync(kFullMask, valid && v == kWaveEmpty) == 0u) break;
        if (spins > (1u << 24)) wave_abort(ctl, WAVE_LOST_ENTRY);
        if (ctl->abort != 0u) break;
    }
    if (valid) *e = kWaveEmpty;
    return valid && v != kWaveEmpty ? (uint32_t)v : 0u;
}

// Second half of the SWEEP, out of line (see wave_exchange): exact re-tests of the flagged spheres of the lane's two rays
// (origin and direction are read back from the records), nearest hits into the records; returns what the two paths do next
// (category of row 0 in the low byte, of row 1 in the next; 0xff = no path).  slot < 0: the lane has no path in that row.
template <bool MOTION>
__device__ __noinline__ unsigned wave_sweep_finish(const KernelArgs& a, const WaveSmem& sm, int slot0, int slot1, int cnt0, int cnt1, int overflow0, int overflow1) {
    unsigned cats = 0xffffu;
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
        const unsigned c = (unsigned)wave_category(a, hit_index, w[26]);
        cats = r ? ((cats & 0x00ffu) | (c << 8)) : ((cats & 0xff00u) | c);
    }
    return cats;
}

// SHADE: up to 32 paths of ONE category (ring entries [head, head + n) of queue `cat`), convergent.  Returns the slot the
// lane now holds a ray for (next hand) or -1, plus 0x10000 if the lane traced a ray and 0x20000 if its path retired.
// Out of line: the warp's loop stays small enough for ptxas to keep the sweep in uniform registers (see wave_exchange).
template <bool MOTION>
__device__ __noinline__ int wave_shade_batch(const KernelArgs& a, const WaveSmem& sm, volatile WaveCtl* ctl, unsigned lane_id, int cat, int n, unsigned head,
                                             int& requeue_slot) {
    const bool valid = lane_id < (unsigned)n;
    const uint32_t slot = wave_take_entry(ctl, sm.queues, sm.cap, cat, head, valid, lane_id);
    Lane L;
    lane_init(L);
    L.finished = true;  // lanes without a path take part in the collectives of lane_refill and nothing else
    float hit_t = kMaxT;
    int hit_index = -1;
    int result = -1;
    requeue_slot = -1;
    if (valid) wave_load(sm.rec(slot), L, hit_t, hit_index);
    if (L.active) {
        result = 0x10000;  // scene.rs:57: one more ray traced
        const MotionCtx mc{a.motion, nullptr, a.order};
        lane_shade<MOTION>(a, L, a.blocks, *sm.P, mc, hit_t, hit_index);
    } else {
        result = 0;
    }
    lane_refill<MOTION>(a, L, lane_id);
    if (valid) {
        wave_store(sm.rec(slot), L);
        if (L.active) result = (result & 0x10000) | (int)slot | 0x40000;  // a new ray (next bounce, or the next sample's camera ray)
        else if (!L.finished) requeue_slot = (int)slot;                   // holds a ticket whose predecessor chunk is not published yet: ask again
        else result |= 0x20000;                                           // no tickets left: the path retires
    }
    return result;
}

template <bool MOTION>
__global__ void __launch_bounds__(kWaveThreads, 1) pt_megakernel_wave(const __grid_constant__ KernelArgs a, const __grid_constant__ ConstImageT<true> ci) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ WaveCtl ctl;
    const WaveSmem sm(smem_raw, a);
    const uint32_t image_bytes = (uint32_t)a.n_blocks * (kWaveConstImg ? 16u : 64u);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        for (int q = 0; q < kWaveQueues; ++q) {
            ctl.head[q] = ctl.tail[q] = 0u;
            ctl.avail[q] = 0;
        }
        ctl.tail[WQ_END] = (unsigned)a.wave_pool;  // every slot starts as a path that needs its first ticket
        ctl.avail[WQ_END] = a.wave_pool;
        ctl.rr = 0u;
        ctl.live = a.wave_pool;
        ctl.abort = 0u;
        ctl.status = a.status;
        for (int i = 0; i < 32; ++i) ctl.scratch[i] = 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0 && image_bytes != 0u) {
        mbar_arrive_expect_tx(&bar, image_bytes);
        tma_bulk_g2s_chunked(sm.kplane, kWaveConstImg ? a.kplane : a.prefilter, image_bytes, &bar);
    }
    stage_perlin(a, sm.P);
    {
        Lane L;
        lane_init(L);
        for (uint32_t i = threadIdx.x; i < kWaveQueues * sm.cap; i += kWaveThreads) sm.queues[i] = kWaveEmpty;
        __syncthreads();
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
    // the hand: the path each lane holds in row 0 / row 1 (-1: none) and what it does next (category, < 0: nothing to file)
    int slot[2] = {-1, -1}, cat[2] = {-1, -1}, retired = 0;
    for (;;) {
        // ---- EXCHANGE ----
        const int4 ex = wave_exchange(&ctl, sm.queues, sm.cap, lane_id, cat[0], (uint32_t)max(slot[0], 0), cat[1], (uint32_t)max(slot[1], 0), retired);
        // branch on VOTES: ptxas keeps the sweep's sphere operands in uniform registers only inside control flow it can prove
        // warp-uniform, and it knows that of a ballot
        if (__ballot_sync(kFullMask, ex.x == -2) != 0u) break;
        // ---- SHADE the two batches; the paths that come out with a ray are the next hand ----
        retired = 0;
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
            const int bc = k ? ((ex.w & 0xff) == 0xff ? -1 : (ex.w & 0xff)) : ex.x;
            const int bn = k ? ((ex.w >> 8) & 0xff) : ex.y;
            const unsigned bh = k ? ((unsigned)ex.w >> 16) : (unsigned)ex.z;
            int keep = -1, requeue = -1;
            if (bn > 0) {
                const int res = wave_shade_batch<MOTION>(a, sm, &ctl, lane_id, bc, bn, bh, requeue);
                rays += (res & 0x10000) ? 1ULL : 0ULL;
                if (res & 0x40000) keep = res & 0xffff;
                retired += __popc(__ballot_sync(kFullMask, (res & 0x20000) != 0)) * (lane_id == 0u ? 1 : 0);
            }
            // a path without a ray that is still waiting for its pixel's previous chunk goes back to the END queue
            const int s_ = keep >= 0 ? keep : requeue;
            const int c_ = keep >= 0 ? -1 : (requeue >= 0 ? WQ_END : -1);
            if (k) { slot[1] = s_; cat[1] = c_; } else { slot[0] = s_; cat[0] = c_; }
        }
        retired = __shfl_sync(kFullMask, retired, 0);
        // ---- SWEEP the hand (paths with cat < 0 and slot >= 0 hold a ray) ----
        float ox[2], oy[2], oz[2], dx[2], dy[2], dz[2];
        bool valid[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            valid[r] = slot[r] >= 0 && cat[r] < 0;
            ox[r] = 0.0f; oy[r] = 1.0e18f; oz[r] = 0.0f;  // parked ray: never a candidate
            dx[r] = dy[r] = dz[r] = 0.0f;
            if (valid[r]) {
                const uint4 c = sm.rec((uint32_t)slot[r])[2], d = sm.rec((uint32_t)slot[r])[3];
                ox[r] = __uint_as_float(c.x); oy[r] = __uint_as_float(c.y); oz[r] = __uint_as_float(c.z);
                dx[r] = __uint_as_float(d.x); dy[r] = __uint_as_float(d.y); dz[r] = __uint_as_float(d.z);
            }
        }
        const unsigned m0 = __ballot_sync(kFullMask, valid[0]), m1 = __ballot_sync(kFullMask, valid[1]);
        if ((m0 | m1) != 0u) {
            sweeps += (m0 != 0u ? 1u : 0u) + (m1 != 0u ? 1u : 0u);
            float o2x[2], o2y[2], o2z[2], nod[2], oo[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                nod[r] = -((ox[r] * dx[r] + oy[r] * dy[r]) + oz[r] * dz[r]);
                oo[r] = ((ox[r] * ox[r] + oy[r] * oy[r]) + oz[r] * oz[r]) * (1.0f - kSlack);
                o2x[r] = ox[r] + ox[r]; o2y[r] = oy[r] + oy[r]; o2z[r] = oz[r] + oz[r];
            }
            int cnt0 = 0, cnt1 = 0;
            int overflow[2] = {a.n_blocks, a.n_blocks};
#if PT_WAVE_LDS
            const ConstImageT<false> none{};
            sweep_two<false, kWaveThreads, kWaveCandCap>(none, sm.kplane, a.n_blocks, sm.cand, cnt0, cnt1, dx, dy, dz, o2x, o2y, o2z, nod, oo, overflow);
#else
            sweep_two<true, kWaveThreads, kWaveCandCap>(ci, sm.kplane, a.n_blocks, sm.cand, cnt0, cnt1, dx, dy, dz, o2x, o2y, o2z, nod, oo, overflow);
#endif
            const unsigned cats = wave_sweep_finish<MOTION>(a, sm, valid[0] ? slot[0] : -1, valid[1] ? slot[1] : -1, cnt0, cnt1, overflow[0], overflow[1]);
            if (valid[0]) cat[0] = (int)(cats & 0xffu);
            if (valid[1]) cat[1] = (int)((cats >> 8) & 0xffu);
        }
    }
    flush_ray_count(a, rays, lane_id, sweeps);
}

}  // namespace pt
