// pt_megakernel.cuh — the persistent render kernel: device side of `Scene::update`
// (src/scene.rs:73-121) with `ray_trace` (src/scene.rs:49-71) unrolled into a per-lane state machine.
//
// Mapping:
//   * one path = one pixel's current sample.  A path owns a pixel for a whole chunk of its samples (so the pixel's
//     xoshiro256+ stream is consumed in exactly the reference's order), then pulls the next ticket from a global atomic
//     queue (warp-aggregated) — the rayon `par_iter_mut` of scene.rs:90-93.
//   * every trip of the main loop is ONE full sphere sweep for all live paths of the warp (convergent), followed by a
//     short divergent shading step.  A path that ended starts its next sample (or next ticket) in the same trip, so the
//     sweep always runs with all live paths.  The default resident kernel (pt_regroup.cuh) keeps one path per lane and
//     regroups the CTA's paths by material between sweep and shading; pt_megakernel_resident carries two paths per lane.
//   * the recursion `emitted + attenuation * ray_trace(..)` becomes `colour += throughput * emitted;
//     throughput *= attenuation`.
//   * stage 1 of the sweep (the conservative pre-filter over all spheres) runs on the tensor path (mma.sync f16 split
//     operands, pt_sweep_mma.cuh) or in packed FP32 (pt_sweep.cuh); its image is staged in shared memory once per CTA with
//     a TMA bulk copy (cp.async.bulk + mbarrier), streamed tile by tile through L2 for scenes that do not fit
//     (pt_megakernel_streamed), or — two-paths-per-lane flavour — read through the uniform datapath from a kernel parameter.
#pragma once
#include "pt_shade.cuh"
#include "pt_sweep.cuh"
#include "pt_sweep_mma.cuh"

namespace pt {

struct KernelArgs {
    const float4* blocks;  // n_blocks * 4 float4 (see pt_sweep.cuh)
    int n_blocks;
    int n_spheres;
    const DevShade* shade;
    const DevTexture* tex;
    const uint8_t* images;     // RGB8 pool of the scene's Image textures (texture.rs:6-37), nullptr when there is none
    const PerlinSmem* perlin;  // global copy, staged to shared memory when has_noise
    const uint32_t* order;     // stored sphere index -> position in the caller's list (equal-t ties), nullptr = identity
    const DevMotion* motion;   // per-sphere MovingSphere records, nullptr when the scene has none (moving_sphere.rs)
    const float4* prefilter;   // pre-filter image X,Y,Z,K per block (global copy; staged/streamed by the LDS kernels)
    const float4* kplane;      // K plane alone, one float4 per block (resident kernel with the X,Y,Z planes in its parameter image)
    const uint4* mma_image;    // tensor-path pre-filter: fragment-ordered sphere operand, (n_steps + 1) x 32 uint4 (pt_sweep_mma.cuh)
    int n_steps;               // ... steps of 16 spheres (= n_blocks / 4)
    MmaScale mma;              // ... its scene scale
    int single_row;            // resident kernel: fewer pixels than lanes, only path row 0 takes work
    int wave_pool;             // wavefront kernel: paths in the CTA's shared-memory pool
    unsigned int* status;      // wavefront kernel: watchdog report (0 = clean), see pt_wave.cuh
    int has_noise;
    DevCamera cam;
    uint32_t width, height, samples, max_depth, frame_num;
    float inv_nx, inv_ny, inv_ns, mix_prev, mix_new;
    int has_sky;
    V3 sky;
    uint32_t tile_rows, part_index, part_count, n_owned_pixels;
    int random_seed;
    uint64_t seed_salt;
    float* rgb;
    unsigned long long* ray_count;
    unsigned long long* sweep_count;  // warp-level sweeps performed (x32 = lane slots offered; rays / that = lane efficiency)
    unsigned int* next_pixel;  // ticket counter of the chunk queue (see lane_refill)
    // chunk queue: ticket t = chunk (t / n_owned_pixels) of owned pixel (t % n_owned_pixels); chunk c covers samples
    // [c * chunk_samples, min((c+1) * chunk_samples, samples)).  chunk_samples is a power of two and chunk_mask =
    // chunk_samples - 1, or chunk_samples >= samples and chunk_mask = 0xffffffff (one chunk per pixel, no state table).
    // pixstate: kPixStateWords words per owned pixel.
    uint32_t chunk_samples, chunk_mask, n_tickets;
    uint32_t* pixstate;
    // streamed variant only
    int tile_blocks;  // blocks per shared-memory tile
    int n_tiles;
    // pt_debug_hits only: caller-supplied rays through the kernels' own sweep phase
    const float* dbg_rays;   // n x (ox, oy, oz, dx, dy, dz)
    const float* dbg_times;  // n ray times, or nullptr (= 0)
    uint32_t dbg_n;
    int32_t* dbg_idx;        // nearest hit as a position in the CALLER's sphere list, -1 = miss
    float* dbg_t;
    uint32_t* dbg_flagged;   // spheres the pre-filter passed on to the exact test (nullptr: not wanted)
};

constexpr unsigned kFullMask = 0xffffffffu;

// ---- mbarrier / TMA bulk-copy helpers (PTX) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy (TMA, non-tensor form); bytes % 16 == 0, 16-byte aligned both sides
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// issue `bytes` as <= 64 KB bulk copies on one barrier (caller has already done arrive.expect_tx(bytes))
__device__ __forceinline__ void tma_bulk_g2s_chunked(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    const uint32_t kChunk = 65536u;
    for (uint32_t off = 0; off < bytes; off += kChunk) {
        const uint32_t n = bytes - off < kChunk ? bytes - off : kChunk;
        tma_bulk_g2s(reinterpret_cast<char*>(dst_smem) + off, reinterpret_cast<const char*>(src_gmem) + off, n, bar);
    }
}

// ---- per-lane path state ----
struct Lane {
    Rng rng;
    V3 o, d;        // current ray
    V3 thr;         // product of attenuations so far
    V3 col;         // sum of radiance over this pixel's samples so far
    uint32_t px, py;
    uint32_t sample;  // samples started for this pixel
    uint32_t depth;
    bool active;      // a path is in flight
    bool have_pixel;
    bool finished;    // queue exhausted
    bool pend;        // holding a ticket whose predecessor chunk is not published yet (see lane_refill)
    float time;       // ray.time of the path in flight (camera.rs:59): read only by MovingSphere, constant along a path
};

// Pixel-state table (global memory, 48 B per owned pixel): what a pixel carries from one chunk of samples to the
// next.  The RNG stream of a pixel is sequential (scene.rs:96-110: one generator per pixel, consumed sample after
// sample), so chunks of one pixel run one after another, but on whichever lane pulls the ticket: words 0-7 xoshiro256+
// state, 8-10 colour sum so far, 11 = number of samples completed (published with st.release, read with ld.acquire).
constexpr int kPixStateWords = 12;
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
#ifdef EXP_NOASM
    return __ldcg(p);
#else
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
#ifdef EXP_NOASM
    __threadfence(); __stcg(p, v);
#else
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

__device__ __forceinline__ void owned_pixel_to_xy(const KernelArgs& a, uint32_t j, uint32_t& x, uint32_t& y) {
    const uint32_t r = j / a.width;
    x = j - r * a.width;
    if (a.part_count <= 1) {
        y = r;
    } else {
        const uint32_t tl = r / a.tile_rows;
        y = (tl * a.part_count + a.part_index) * a.tile_rows + (r - tl * a.tile_rows);
    }
}

__device__ __forceinline__ uint32_t xy_to_owned_pixel(const KernelArgs& a, uint32_t x, uint32_t y) {
    if (a.part_count <= 1) return y * a.width + x;
    const uint32_t tile = y / a.tile_rows;
    const uint32_t tl = tile / a.part_count;
    return (tl * a.tile_rows + (y - tile * a.tile_rows)) * a.width + x;
}

// scene.rs:113-116 — blend the finished pixel into the accumulation buffer
__device__ __forceinline__ void write_pixel(const KernelArgs& a, const Lane& L) {
    float* out = a.rgb + ((size_t)L.py * a.width + L.px) * 3;
    const V3 col = L.col * a.inv_ns;
    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
    if (a.frame_num != 0) {
        p0 = out[0];
        p1 = out[1];
        p2 = out[2];
    }
    out[0] = p0 * a.mix_prev + col.x * a.mix_new;
    out[1] = p1 * a.mix_prev + col.y * a.mix_new;
    out[2] = p2 * a.mix_prev + col.z * a.mix_new;
}

// hand the pixel to whoever pulls its next chunk: state first, then the sample count with release semantics
__device__ __forceinline__ void publish_pixel_state(const KernelArgs& a, const Lane& L) {
    uint32_t* st = a.pixstate + (size_t)xy_to_owned_pixel(a, L.px, L.py) * kPixStateWords;
    uint4* v = reinterpret_cast<uint4*>(st);
    __stcg(v, make_uint4((uint32_t)L.rng.s0, (uint32_t)(L.rng.s0 >> 32), (uint32_t)L.rng.s1, (uint32_t)(L.rng.s1 >> 32)));
    __stcg(v + 1, make_uint4((uint32_t)L.rng.s2, (uint32_t)(L.rng.s2 >> 32), (uint32_t)L.rng.s3, (uint32_t)(L.rng.s3 >> 32)));
    __stcg(reinterpret_cast<uint2*>(st + 8), make_uint2(__float_as_uint(L.col.x), __float_as_uint(L.col.y)));
    __stcg(st + 10, __float_as_uint(L.col.z));
    st_release_u32(st + 11, L.sample);
}
// true when the predecessor chunk has been published; loads the state (L2 loads: another SM wrote it)
__device__ __forceinline__ bool acquire_pixel_state(const KernelArgs& a, Lane& L) {
    const uint32_t* st = a.pixstate + (size_t)xy_to_owned_pixel(a, L.px, L.py) * kPixStateWords;
    if (ld_acquire_u32(st + 11) != L.sample) return false;
    const uint4* v = reinterpret_cast<const uint4*>(st);
    const uint4 r0 = __ldcg(v), r1 = __ldcg(v + 1), c = __ldcg(v + 2);
    L.rng.s0 = (uint64_t)r0.x | ((uint64_t)r0.y << 32);
    L.rng.s1 = (uint64_t)r0.z | ((uint64_t)r0.w << 32);
    L.rng.s2 = (uint64_t)r1.x | ((uint64_t)r1.y << 32);
    L.rng.s3 = (uint64_t)r1.z | ((uint64_t)r1.w << 32);
    L.col = v3(__uint_as_float(c.x), __uint_as_float(c.y), __uint_as_float(c.z));
    return true;
}

// top-of-trip bookkeeping: finish chunks/pixels, pull new tickets, start the next sample (scene.rs:94-110).
//
// Work queue.  The reference hands whole pixels to its rayon workers (scene.rs:90-93).  With ~10^5 lanes a pixel
// (samples x ~2.7 sweeps, strictly sequential because of its RNG stream) is too coarse a unit: once the queue runs dry
// every warp idles lane by lane for up to one pixel's duration.  So the queue holds CHUNKS of `chunk_samples` samples in
// sample-major order (all pixels' chunk 0, then all pixels' chunk 1, ...): every pixel advances at the same pace and the
// tail of a launch is one chunk, not one pixel.  A pixel's generator and colour sum travel through the pixel-state
// table, so each pixel still consumes exactly the reference's stream and sums its samples in the reference's order.
// A lane whose ticket's predecessor chunk is not published yet (rare: the predecessor was handed out n_owned_pixels
// tickets earlier) parks for a trip and asks again — it never spins, the predecessor may be a lane of the same warp.
// `pend` is that lane's "holding an unready ticket" flag; the end-of-chunk test uses the uniform chunk_mask rather than a
// per-lane bound.  No path state is carried across the sweep in registers: it lives in the path's shared-memory record
// (resident kernel) or in the lane's shared-memory slots (streamed kernel: pend, time).
template <bool MOTION>
__device__ __forceinline__ void lane_refill(const KernelArgs& a, Lane& L, unsigned lane_id) {
    bool want_pixel = false;
    if (!L.active && !L.finished) {
        if (L.have_pixel && ((L.sample & a.chunk_mask) == 0u || L.sample >= a.samples)) {
            if (L.sample >= a.samples)
                write_pixel(a, L);
            else
                publish_pixel_state(a, L);
            L.have_pixel = false;
        }
        want_pixel = !L.have_pixel && !L.pend;
    }
    const unsigned need = __ballot_sync(kFullMask, want_pixel);
    if (need != 0u) {
        const int leader = __ffs(need) - 1;
        unsigned base = 0;
        if ((int)lane_id == leader) base = atomicAdd(a.next_pixel, (unsigned)__popc(need));
        base = __shfl_sync(kFullMask, base, leader);
        if (want_pixel) {
            const unsigned t = base + (unsigned)__popc(need & ((1u << lane_id) - 1u));
            if (t < a.n_tickets) {
                const uint32_t chunk = t / a.n_owned_pixels;
                const uint32_t j = t - chunk * a.n_owned_pixels;
                owned_pixel_to_xy(a, j, L.px, L.py);
                uint64_t seed;
                if (a.random_seed) {
                    // scene.rs:96-97: an independent host-entropy seed per pixel per frame
                    uint64_t h = a.seed_salt + ((uint64_t)L.py * a.width + L.px) * 0x9e3779b97f4a7c15ULL +
                                 (uint64_t)a.frame_num * 0xd1b54a32d192ed03ULL;
                    seed = splitmix64_next(h);
                } else {
                    seed = pixel_seed(L.px, L.py, a.frame_num);
                }
                rng_seed(L.rng, seed);  // kept by chunk 0 only: later chunks load the published state
                L.col = v3(0.0f, 0.0f, 0.0f);
                L.sample = chunk * a.chunk_samples;
                L.have_pixel = chunk == 0u;
                L.pend = chunk != 0u;
            } else {
                L.finished = true;
            }
        }
    }
    if (!L.active && !L.finished && !L.have_pixel) {
        if (L.pend && acquire_pixel_state(a, L)) {  // not yet published: the lane sits this trip out and asks again
            L.pend = false;
            L.have_pixel = true;
        }
    }
    if (!L.active && L.have_pixel) {
        // scene.rs:107-110
        const float u = ((float)L.px + rng_f32(L.rng)) * a.inv_nx;
        const float v = ((float)L.py + rng_f32(L.rng)) * a.inv_ny;
        float time;
        camera_get_ray(a.cam, u, v, L.rng, L.o, L.d, time);
        // ray.time is read only by MovingSphere (moving_sphere.rs:39) and is constant along a path (material.rs:62,83,117)
        if (MOTION) L.time = time;
        L.thr = v3(1.0f, 1.0f, 1.0f);
        L.depth = 0;
        L.sample += 1;
        L.active = true;
    }
}

// after the sweep: scene.rs:57-70 for the lane's current ray
template <bool MOTION>
__device__ __forceinline__ void lane_shade(const KernelArgs& a, Lane& L, const float4* __restrict__ blk, const PerlinSmem& P,
                                           const MotionCtx& mc, float hit_t, int hit_index) {
    if (hit_index < 0) {
        L.col = L.col + L.thr * sky_colour(a.has_sky != 0, a.sky, L.d);
        L.active = false;
        return;
    }
    // spheres_soa.rs:132-139 epilogue
    const V3 point = L.o + (hit_t * L.d);
    const float* bf = reinterpret_cast<const float*>(blk) + (hit_index >> 2) * 16 + (hit_index & 3);
    const V3 centre = v3(bf[0], bf[4], bf[8]);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(a.shade + hit_index));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(a.shade + hit_index) + 1);
    DevShade m;
    m.ar = s0.x; m.ag = s0.y; m.ab = s0.z; m.param = s0.w;
    m.rinv = s1.x; m.kind = __float_as_int(s1.y); m.tex = __float_as_int(s1.z);
    V3 normal;
    const bool moving = MOTION && __float_as_int(s1.w) != 0;
    if (moving) {  // MovingSphere: centre at ray.time, normal = (p - centre) / radius (moving_sphere.rs:28-31,49)
        const DevMotion mo = mc.table[hit_index];
        const float s = (L.time - mo.time_start) * mo.inv_time_delta;
        const V3 c = v3(centre.x + s * mo.dx, centre.y + s * mo.dy, centre.z + s * mo.dz);
        normal = v3((point.x - c.x) / mo.radius, (point.y - c.y) / mo.radius, (point.z - c.z) / mo.radius);
    } else {
        normal = (point - centre) * m.rinv;
    }
    if (m.kind == MAT_DIFFUSE_LIGHT) {  // material.rs:161-167 (+ :157: lights do not scatter)
        const V3 em = m.tex < 0 ? v3(m.ar, m.ag, m.ab) : texture_value(TexCtx{a.tex, a.images, !moving}, P, m.tex, point, normal);
        L.col = L.col + L.thr * em;
        L.active = false;
        return;
    }
    if (L.depth < a.max_depth) {
        V3 att, sd;
        if (material_scatter(m, TexCtx{a.tex, a.images, !moving}, P, L.d, point, normal, L.rng, att, sd)) {
            L.thr = L.thr * att;
            L.o = point;
            L.d = sd;
            L.depth += 1;
            return;
        }
    }
    L.active = false;  // absorbed, or depth limit: contributes `emitted` = 0
}

__device__ __forceinline__ void lane_init(Lane& L) {
    L.active = false;
    L.have_pixel = false;
    L.finished = false;
    L.sample = 0;
    L.depth = 0;
    L.px = L.py = 0;
    L.o = v3(0.0f, 0.0f, 0.0f);
    L.d = v3(0.0f, 0.0f, 0.0f);
    L.thr = v3(0.0f, 0.0f, 0.0f);
    L.col = v3(0.0f, 0.0f, 0.0f);
    L.rng.s0 = L.rng.s1 = L.rng.s2 = L.rng.s3 = 0;
    L.pend = false;
    L.time = 0.0f;
}

__device__ __forceinline__ void flush_ray_count(const KernelArgs& a, unsigned long long rays, unsigned lane_id, unsigned sweeps) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) rays += __shfl_xor_sync(kFullMask, rays, off);
    if (lane_id == 0 && rays != 0ULL) atomicAdd(a.ray_count, rays);
    if (lane_id == 0 && sweeps != 0u) atomicAdd(a.sweep_count, (unsigned long long)sweeps);
}

__device__ __forceinline__ void stage_perlin(const KernelArgs& a, PerlinSmem* P) {
    if (a.has_noise) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.perlin);
        uint32_t* dst = reinterpret_cast<uint32_t*>(P);
        for (unsigned i = threadIdx.x; i < sizeof(PerlinSmem) / 4; i += blockDim.x) dst[i] = src[i];
    }
}

#ifndef PT_CTA_THREADS
#define PT_CTA_THREADS 256
#endif
constexpr int kCtaThreads = PT_CTA_THREADS;


// =====================================================================================================
// Resident variant: the whole scene stays on chip for the life of the CTA, and every lane carries TWO paths.
//
// Why two.  The sweep is bound by register-file reads and by the uniform loads that feed it (pt_sweep.cuh, sweep_two):
// a sphere pair fetched once — LDCU.64 into a uniform register pair from the kernel-parameter image, or one LDS.128 for
// scenes beyond it — serves both of the lane's rays.  A path's state (26 words: generator, ray, throughput, colour sum,
// pixel, counters, flags, ray.time) lives in the path's own record in shared memory, [word][path] with path = row * T + tid:
// the lane that sweeps a path is the lane that shades it, so the records are lane-private — no barrier and no atomics
// anywhere in the main loop; warps drift apart and one warp's shading overlaps another's sweep.  Only the two rays
// (8 derived words each) and the nearest hit so far are in registers during the sweep; a record is loaded, advanced by
// lane_shade / lane_refill and stored back once per trip.  Every path still consumes exactly its own pixel's RNG stream
// in the reference's order (scene.rs:96-110), so images do not depend on any of this.
//
// (Round 1 sorted the CTA's 256 paths by material through shared memory between sweep and shading — two CTA barriers per
// trip, 14 % of the warp time at BAR.SYNC for 11 % fewer instructions: a wash on cfg2.  The records make a warp-local sort
// possible later without moving state; it is not needed to beat that kernel.)
// =====================================================================================================
constexpr int kPathWords = 26;  // 8 rng + 3 o + 3 d + 3 thr + 3 col + px, py, sample, depth, flags, time
constexpr int kPathRows = 2;    // paths per lane

// record of path (row, tid): word w at rec[w * kPathRows * kCtaThreads], rec = state + row * kCtaThreads + tid
__device__ __forceinline__ void path_store(uint32_t* __restrict__ rec, const Lane& L) {
    constexpr int S = kPathRows * kCtaThreads;
    rec[0 * S] = (uint32_t)L.rng.s0;  rec[1 * S] = (uint32_t)(L.rng.s0 >> 32);
    rec[2 * S] = (uint32_t)L.rng.s1;  rec[3 * S] = (uint32_t)(L.rng.s1 >> 32);
    rec[4 * S] = (uint32_t)L.rng.s2;  rec[5 * S] = (uint32_t)(L.rng.s2 >> 32);
    rec[6 * S] = (uint32_t)L.rng.s3;  rec[7 * S] = (uint32_t)(L.rng.s3 >> 32);
    rec[8 * S] = __float_as_uint(L.o.x);  rec[9 * S] = __float_as_uint(L.o.y);  rec[10 * S] = __float_as_uint(L.o.z);
    rec[11 * S] = __float_as_uint(L.d.x); rec[12 * S] = __float_as_uint(L.d.y); rec[13 * S] = __float_as_uint(L.d.z);
    rec[14 * S] = __float_as_uint(L.thr.x); rec[15 * S] = __float_as_uint(L.thr.y); rec[16 * S] = __float_as_uint(L.thr.z);
    rec[17 * S] = __float_as_uint(L.col.x); rec[18 * S] = __float_as_uint(L.col.y); rec[19 * S] = __float_as_uint(L.col.z);
    rec[20 * S] = L.px;  rec[21 * S] = L.py;  rec[22 * S] = L.sample;  rec[23 * S] = L.depth;
    rec[24 * S] = (L.active ? 1u : 0u) | (L.have_pixel ? 2u : 0u) | (L.finished ? 4u : 0u) | (L.pend ? 8u : 0u);
    rec[25 * S] = __float_as_uint(L.time);
}
__device__ __forceinline__ void path_load(const uint32_t* __restrict__ rec, Lane& L) {
    constexpr int S = kPathRows * kCtaThreads;
    L.rng.s0 = (uint64_t)rec[0 * S] | ((uint64_t)rec[1 * S] << 32);
    L.rng.s1 = (uint64_t)rec[2 * S] | ((uint64_t)rec[3 * S] << 32);
    L.rng.s2 = (uint64_t)rec[4 * S] | ((uint64_t)rec[5 * S] << 32);
    L.rng.s3 = (uint64_t)rec[6 * S] | ((uint64_t)rec[7 * S] << 32);
    L.o = v3(__uint_as_float(rec[8 * S]), __uint_as_float(rec[9 * S]), __uint_as_float(rec[10 * S]));
    L.d = v3(__uint_as_float(rec[11 * S]), __uint_as_float(rec[12 * S]), __uint_as_float(rec[13 * S]));
    L.thr = v3(__uint_as_float(rec[14 * S]), __uint_as_float(rec[15 * S]), __uint_as_float(rec[16 * S]));
    L.col = v3(__uint_as_float(rec[17 * S]), __uint_as_float(rec[18 * S]), __uint_as_float(rec[19 * S]));
    L.px = rec[20 * S];  L.py = rec[21 * S];  L.sample = rec[22 * S];  L.depth = rec[23 * S];
    const uint32_t f = rec[24 * S];
    L.active = (f & 1u) != 0u;  L.have_pixel = (f & 2u) != 0u;  L.finished = (f & 4u) != 0u;  L.pend = (f & 8u) != 0u;
    L.time = __uint_as_float(rec[25 * S]);
}

#ifdef PT_PAIR_MIN_CTAS
#define PT_PAIR_LAUNCH_BOUNDS __launch_bounds__(kCtaThreads, PT_PAIR_MIN_CTAS)
#else
#define PT_PAIR_LAUNCH_BOUNDS __launch_bounds__(kCtaThreads, 768 / kCtaThreads)  // 24 warps per SM at 80 registers
#endif
// PT_TRIP_SYNC=1: the warps of a CTA start every trip together (one CTA barrier per trip)
#ifndef PT_TRIP_SYNC
#define PT_TRIP_SYNC 0
#endif

// shared-memory layout of the resident kernels:
// [pre-filter image: K plane (16 B per block) or X,Y,Z,K (64 B per block)] [Perlin tables] [queues] [path records]
template <bool CONSTIMG>
struct ResidentSmem {
    float4* pf;
    PerlinSmem* P;
    uint32_t* queue;  // this lane's queue: [kQueueCap][kCtaThreads]
    uint32_t* state;  // this lane's records: [kPathWords][kPathRows][kCtaThreads]
    uint32_t image_bytes;
    __device__ __forceinline__ ResidentSmem(unsigned char* raw, int n_blocks) {
        image_bytes = (uint32_t)n_blocks * (CONSTIMG ? 16u : 64u);
        pf = reinterpret_cast<float4*>(raw);
        P = reinterpret_cast<PerlinSmem*>(raw + ((image_bytes + 127u) & ~127u));
        queue = reinterpret_cast<uint32_t*>(P + 1) + threadIdx.x;
        state = reinterpret_cast<uint32_t*>(P + 1) + kQueueCap * kCtaThreads + threadIdx.x;
    }
    // where path row r keeps its ray.time (MotionCtx reads it on the rare paths)
    __device__ __forceinline__ const float* time_slot(int row) const {
        return reinterpret_cast<const float*>(state + 25 * kPathRows * kCtaThreads + row * kCtaThreads);
    }
};
// stage the image with one TMA bulk copy (thread 0 issues, everyone waits on the mbarrier after the caller's own set-up)
template <bool CONSTIMG>
__device__ __forceinline__ void resident_stage_begin(const KernelArgs& a, const ResidentSmem<CONSTIMG>& sm, uint64_t* bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && sm.image_bytes != 0u) {
        mbar_arrive_expect_tx(bar, sm.image_bytes);
        tma_bulk_g2s_chunked(sm.pf, CONSTIMG ? a.kplane : a.prefilter, sm.image_bytes, bar);
    }
}

// The sweep phase of one trip: both rays of the lane against the whole scene, nearest hits into hit_t / hit_index.
// o/d: the two rays, already replaced by the parked ray for paths that are not in flight.  pf: the staged image (K plane
// for CONSTIMG); queue: this lane's candidate queue, [QCAP][QSTRIDE threads].
template <bool CONSTIMG, bool MOTION, int QSTRIDE, int QCAP>
__device__ __forceinline__ void resident_sweep_core(const KernelArgs& a, const ConstImageT<CONSTIMG>& ci, const float4* __restrict__ pf, uint32_t* __restrict__ queue,
                                                    const MotionCtx& mc0, const MotionCtx& mc1, const float (&ox)[2], const float (&oy)[2], const float (&oz)[2],
                                                    const float (&dx)[2], const float (&dy)[2], const float (&dz)[2], float (&hit_t)[2], int (&hit_index)[2],
                                                    unsigned (&flagged)[2]) {
    float o2x[2], o2y[2], o2z[2], nod[2], oo[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        nod[r] = -((ox[r] * dx[r] + oy[r] * dy[r]) + oz[r] * dz[r]);
        oo[r] = ((ox[r] * ox[r] + oy[r] * oy[r]) + oz[r] * oz[r]) * (1.0f - kSlack);
        o2x[r] = ox[r] + ox[r]; o2y[r] = oy[r] + oy[r]; o2z[r] = oz[r] + oz[r];  // exact: o is recovered as 0.5 * o2 for the re-tests
        hit_t[r] = kMaxT;
        hit_index[r] = -1;
    }
    int cnt0 = 0, cnt1 = 0;
    int overflow[2] = {a.n_blocks, a.n_blocks};
    sweep_two<CONSTIMG, QSTRIDE, QCAP>(ci, pf, a.n_blocks, queue, cnt0, cnt1, dx, dy, dz, o2x, o2y, o2z, nod, oo, overflow);
    sweep_drain_range<MOTION, QSTRIDE>(a.blocks, mc0, queue, 0, cnt0, 0.5f * o2x[0], 0.5f * o2y[0], 0.5f * o2z[0], dx[0], dy[0], dz[0], hit_t[0], hit_index[0], flagged[0]);
    sweep_drain_range<MOTION, QSTRIDE>(a.blocks, mc1, queue, QCAP - cnt1, cnt1, 0.5f * o2x[1], 0.5f * o2y[1], 0.5f * o2z[1], dx[1], dy[1], dz[1], hit_t[1], hit_index[1], flagged[1]);
    if (min(overflow[0], overflow[1]) < a.n_blocks) {  // a queue overflowed (rare): one out-of-line pass per affected ray
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
            const int first = r ? overflow[1] : overflow[0];
            if (first >= a.n_blocks) continue;
            float ht = r ? hit_t[1] : hit_t[0];
            int hi = r ? hit_index[1] : hit_index[0];
            unsigned fl = 0u;
            sweep_overflow<MOTION>(a.blocks, r ? mc1 : mc0, first, a.n_blocks, 0.5f * (r ? o2x[1] : o2x[0]), 0.5f * (r ? o2y[1] : o2y[0]), 0.5f * (r ? o2z[1] : o2z[0]),
                                   r ? dx[1] : dx[0], r ? dy[1] : dy[0], r ? dz[1] : dz[0], ht, hi, fl);
            if (r) { hit_t[1] = ht; hit_index[1] = hi; flagged[1] += fl; } else { hit_t[0] = ht; hit_index[0] = hi; flagged[0] += fl; }
        }
    }
}
template <bool CONSTIMG, bool MOTION>
__device__ __forceinline__ void resident_sweep(const KernelArgs& a, const ConstImageT<CONSTIMG>& ci, const ResidentSmem<CONSTIMG>& sm, const MotionCtx& mc0,
                                               const MotionCtx& mc1, const float (&ox)[2], const float (&oy)[2], const float (&oz)[2], const float (&dx)[2],
                                               const float (&dy)[2], const float (&dz)[2], float (&hit_t)[2], int (&hit_index)[2], unsigned (&flagged)[2]) {
    resident_sweep_core<CONSTIMG, MOTION, kCtaThreads, kQueueCap>(a, ci, sm.pf, sm.queue, mc0, mc1, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
}

// CONSTIMG: X, Y, Z planes of the pre-filter image arrive as the kernel parameter `ci` (scenes of up to kMaxConstSpheres
// spheres — every preset of the reference) and only the K plane is staged in shared memory; otherwise the whole image is
// staged.  MOTION: the scene has Hitable::MovingSphere entries.
// a.single_row: fewer pixels than lanes — row 1 stays empty (its lanes never pull a ticket)
template <bool CONSTIMG, bool MOTION>
__global__ void PT_PAIR_LAUNCH_BOUNDS pt_megakernel_resident(const __grid_constant__ KernelArgs a, const __grid_constant__ ConstImageT<CONSTIMG> ci) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const ResidentSmem<CONSTIMG> sm(smem_raw, a.n_blocks);
    resident_stage_begin<CONSTIMG>(a, sm, &bar);
    stage_perlin(a, sm.P);
    const unsigned lane_id = threadIdx.x & 31u;
    {
        Lane L;
        lane_init(L);
#pragma unroll 1
        for (int row = 0; row < kPathRows; ++row) {
            Lane N = L;
            // small images: row 1 never takes work (the sweep still carries its parked ray; what counts there is latency)
            N.finished = a.single_row != 0 && row != 0;
            lane_refill<MOTION>(a, N, lane_id);
            path_store(sm.state + row * kCtaThreads, N);
        }
    }
    __syncthreads();
    if (sm.image_bytes != 0u) mbar_wait(&bar, 0);
    const MotionCtx mc0{a.motion, sm.time_slot(0), a.order};
    const MotionCtx mc1{a.motion, sm.time_slot(1), a.order};

    unsigned long long rays = 0ULL;
    unsigned sweeps = 0u;
    for (;;) {
        // ---- the two rays of this lane ----
        float ox[2], oy[2], oz[2], dx[2], dy[2], dz[2], hit_t[2];
        int hit_index[2];
        bool act[2], fin = true;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t* rec = sm.state + r * kCtaThreads;
            constexpr int S = kPathRows * kCtaThreads;
            const uint32_t f = rec[24 * S];
            act[r] = (f & 1u) != 0u;
            fin = fin && (f & 4u) != 0u;
            ox[r] = __uint_as_float(rec[8 * S]); oy[r] = __uint_as_float(rec[9 * S]); oz[r] = __uint_as_float(rec[10 * S]);
            dx[r] = __uint_as_float(rec[11 * S]); dy[r] = __uint_as_float(rec[12 * S]); dz[r] = __uint_as_float(rec[13 * S]);
            if (!act[r]) {  // parked path: |o|^2 = 1e36 dwarfs every L, d = 0 -> never a candidate
                ox[r] = 0.0f; oy[r] = 1.0e18f; oz[r] = 0.0f;
                dx[r] = dy[r] = dz[r] = 0.0f;
            }
            hit_t[r] = kMaxT;
            hit_index[r] = -1;
        }
#if PT_TRIP_SYNC
        if (__syncthreads_and(fin ? 1 : 0)) break;  // every path of this CTA has run out of tickets
#else
        if (__all_sync(kFullMask, fin)) break;  // every path of this warp has run out of tickets
#endif
        const unsigned m0 = __ballot_sync(kFullMask, act[0]), m1 = __ballot_sync(kFullMask, act[1]);
        if ((m0 | m1) != 0u) {
            sweeps += (m0 != 0u ? 1u : 0u) + (m1 != 0u ? 1u : 0u);
            unsigned flagged[2] = {0u, 0u};
            resident_sweep<CONSTIMG, MOTION>(a, ci, sm, mc0, mc1, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
        }
        // ---- shade / refill, one path row after the other (same code, the row only moves the record's address) ----
#pragma unroll 1
        for (int row = 0; row < kPathRows; ++row) {
            uint32_t* rec = sm.state + row * kCtaThreads;
            Lane L;
            path_load(rec, L);
            const float ht = row == 0 ? hit_t[0] : hit_t[1];
            const int hi = row == 0 ? hit_index[0] : hit_index[1];
            if (L.active) {
                rays += 1ULL;  // scene.rs:57
                lane_shade<MOTION>(a, L, a.blocks, *sm.P, row == 0 ? mc0 : mc1, ht, hi);
            }
            lane_refill<MOTION>(a, L, lane_id);
            path_store(rec, L);
        }
    }
    flush_ray_count(a, rays, lane_id, sweeps);
}

// pt_debug_hits, resident scenes: caller-supplied rays through resident_sweep — the same staging, operands, queue and
// re-tests as the render kernel.  Ray i of the batch sits in path row (i / T) % 2 of lane i % T of CTA i / 2T.
template <bool CONSTIMG, bool MOTION>
__global__ void PT_PAIR_LAUNCH_BOUNDS pt_debug_hits_resident(const __grid_constant__ KernelArgs a, const __grid_constant__ ConstImageT<CONSTIMG> ci) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const ResidentSmem<CONSTIMG> sm(smem_raw, a.n_blocks);
    resident_stage_begin<CONSTIMG>(a, sm, &bar);
    __syncthreads();
    if (sm.image_bytes != 0u) mbar_wait(&bar, 0);
    const MotionCtx mc0{a.motion, sm.time_slot(0), a.order};
    const MotionCtx mc1{a.motion, sm.time_slot(1), a.order};
    constexpr uint32_t kBatch = kPathRows * kCtaThreads;
    for (uint32_t base = blockIdx.x * kBatch; base < a.dbg_n; base += gridDim.x * kBatch) {
        float ox[2], oy[2], oz[2], dx[2], dy[2], dz[2], hit_t[2];
        int hit_index[2];
        unsigned flagged[2] = {0u, 0u};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t i = base + r * kCtaThreads + threadIdx.x;
            ox[r] = 0.0f; oy[r] = 1.0e18f; oz[r] = 0.0f;
            dx[r] = dy[r] = dz[r] = 0.0f;
            float time = 0.0f;
            if (i < a.dbg_n) {
                const float* ray = a.dbg_rays + (size_t)i * 6;
                ox[r] = ray[0]; oy[r] = ray[1]; oz[r] = ray[2]; dx[r] = ray[3]; dy[r] = ray[4]; dz[r] = ray[5];
                if (a.dbg_times) time = a.dbg_times[i];
            }
            *const_cast<float*>(sm.time_slot(r)) = time;
        }
        __syncwarp();
        resident_sweep<CONSTIMG, MOTION>(a, ci, sm, mc0, mc1, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t i = base + r * kCtaThreads + threadIdx.x;
            if (i < a.dbg_n) {
                a.dbg_idx[i] = hit_index[r] < 0 ? -1 : (a.order ? (int32_t)__ldg(a.order + hit_index[r]) : hit_index[r]);
                a.dbg_t[i] = hit_t[r];
                if (a.dbg_flagged) a.dbg_flagged[i] = flagged[r];
            }
        }
    }
}

// pt_debug_hits mode 1: the reference's exact expression on EVERY stored sphere, no pre-filter (any scene size) — what the
// two-stage sweep must reproduce.  One ray per thread, spheres in stored order, ties to the lower position in the caller's list.
template <bool MOTION>
__global__ void pt_debug_hits_exact_all(const __grid_constant__ KernelArgs a) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.dbg_n; i += gridDim.x * blockDim.x) {
        const float* ray = a.dbg_rays + (size_t)i * 6;
        const float ox = ray[0], oy = ray[1], oz = ray[2], dx = ray[3], dy = ray[4], dz = ray[5];
        const float time = a.dbg_times ? a.dbg_times[i] : 0.0f;
        const MotionCtx mc{a.motion, &time, a.order};
        float hit_t = kMaxT;
        int hit_index = -1;
        unsigned flagged = 0u;
        for (int g = 0; g < a.n_blocks / kLdsGroupBlocks; ++g)
            sweep_resolve_entry<kLdsMaskBits, MOTION, true>(a.blocks, mc, ((uint32_t)g << kLdsMaskBits) | ((1u << kLdsMaskBits) - 1u), ox, oy, oz, dx, dy, dz, hit_t,
                                                            hit_index, flagged);
        a.dbg_idx[i] = hit_index < 0 ? -1 : (a.order ? (int32_t)__ldg(a.order + hit_index) : hit_index);
        a.dbg_t[i] = hit_t;
        if (a.dbg_flagged) a.dbg_flagged[i] = flagged;
    }
}

// =====================================================================================================
// Streamed variant: the SoA does not fit in shared memory.  All warps of the CTA sweep the same tile;
// tiles are double-buffered with TMA bulk copies (full/empty mbarrier pair per buffer), so each trip of
// the main loop streams the whole SoA once through L2 for every live lane of the CTA.
// =====================================================================================================
#ifdef PT_STREAM_NOPIPE
constexpr bool kStreamPipe = false;
#else
constexpr bool kStreamPipe = true;  // LDS one block ahead (see sweep_expanded)
#endif
#ifdef PT_STREAM_MIN_CTAS
#define PT_STREAM_LAUNCH_BOUNDS __launch_bounds__(kCtaThreads, PT_STREAM_MIN_CTAS)
#else
#define PT_STREAM_LAUNCH_BOUNDS __launch_bounds__(kCtaThreads)
#endif

// the double-buffered tile pipeline of one CTA: full/empty mbarrier pair per buffer, thread 0 produces
struct TileStream {
    uint64_t* full_bar;   // [2]
    uint64_t* empty_bar;  // [2]
    unsigned char* buf0;  // tile buffer b lives at buf0 + b * tile_bytes (computed, not looked up: an array of pointers would
    uint32_t tile_bytes;  // be demoted to local memory and the sweep's loads would lose their shared-memory address space)
    const char* src;      // the image being streamed: the FP32 pre-filter image (64 B per block of 4 spheres) or the
    uint32_t block_bytes; // tensor-path fragment image (128 B per block)
    uint32_t full_phase[2];   // parity to wait for on full_bar[b]
    uint32_t empty_phase[2];  // producer side: parity to wait for on empty_bar[b]
    uint32_t produced[2];     // producer: how many times buffer b has been filled
    __device__ __forceinline__ float4* tile_buf(int b) const { return reinterpret_cast<float4*>(buf0 + (size_t)b * tile_bytes); }
    __device__ __forceinline__ void init(uint64_t* full, uint64_t* empty, unsigned char* raw, uint32_t bytes, const void* image, uint32_t bytes_per_block) {
        full_bar = full; empty_bar = empty; buf0 = raw; tile_bytes = bytes;
        src = reinterpret_cast<const char*>(image); block_bytes = bytes_per_block;
        full_phase[0] = full_phase[1] = empty_phase[0] = empty_phase[1] = produced[0] = produced[1] = 0u;
        if (threadIdx.x == 0) {
            mbar_init(&full_bar[0], 1);
            mbar_init(&full_bar[1], 1);
            mbar_init(&empty_bar[0], kCtaThreads);  // every lane releases the buffer itself (no elected lane: each thread's
            mbar_init(&empty_bar[1], kCtaThreads);  // own reads are ordered before its own arrive)
            fence_mbar_init();
        }
    }
    __device__ __forceinline__ void produce(const KernelArgs& a, int tile) {  // thread 0 only
        const int b = tile & 1;
        if (produced[b] != 0u) {  // wait until every warp has released the previous contents
            mbar_wait(&empty_bar[b], empty_phase[b]);
            empty_phase[b] ^= 1u;
        }
        produced[b] += 1u;
        const int first = tile * a.tile_blocks;
        const int nb = min(a.tile_blocks, a.n_blocks - first);
        const uint32_t bytes = (uint32_t)nb * block_bytes;
        mbar_arrive_expect_tx(&full_bar[b], bytes);
        tma_bulk_g2s_chunked(tile_buf(b), src + (size_t)first * block_bytes, bytes, &full_bar[b]);
    }
};

// The sweep phase of one trip of the streamed kernels: the CTA streams the whole image once through its two tile buffers and
// every lane tests its ray against each tile.  Must be called by ALL threads of the CTA (parked lanes carry the parked ray).
template <bool MOTION>
__device__ __forceinline__ void streamed_sweep(const KernelArgs& a, TileStream& ts, const MotionCtx& mc, uint32_t* queue, float ox, float oy, float oz, float dx,
                                               float dy, float dz, float& hit_t, int& hit_index, unsigned& flagged) {
    hit_t = kMaxT;
    hit_index = -1;
    const float nod = -((ox * dx + oy * dy) + oz * dz);
    const float oo = ((ox * ox + oy * oy) + oz * oz) * (1.0f - kSlack);
    int cnt = 0;
    if (threadIdx.x == 0) {
        ts.produce(a, 0);
        if (a.n_tiles > 1) ts.produce(a, 1);
    }
    for (int tile = 0; tile < a.n_tiles; ++tile) {
        const int b = tile & 1;
        mbar_wait(&ts.full_bar[b], ts.full_phase[b]);
        ts.full_phase[b] ^= 1u;
        __syncwarp();
        const int first = tile * a.tile_blocks;
        const int nb = min(a.tile_blocks, a.n_blocks - first);
        sweep_expanded<kStreamPipe, MOTION>(ts.tile_buf(b), nb, first, a.blocks, mc, queue, cnt, ox, oy, oz, dx, dy, dz, nod, ox + ox, oy + oy, oz + oz, oo, hit_t, hit_index, flagged);
        mbar_arrive(&ts.empty_bar[b]);
        __syncwarp();
        if (threadIdx.x == 0 && tile + 2 < a.n_tiles) ts.produce(a, tile + 2);
        // this tile's candidates: exact re-test against the global SoA (L2), off the tile buffer's critical path
        sweep_drain<MOTION, false>(a.blocks, mc, queue, cnt, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
    }
}

// Same with stage 1 on the tensor path (pt_sweep_mma.cuh): the tiles are steps of the fragment image, the ray's A fragments are
// built once per trip (stage: the CTA's transpose buffer) and every tile's candidates are resolved by the quad before the next.
// The loop fetches one step ahead: past a tile's last step it reads the first bytes of whatever follows the buffer in shared
// memory (the other tile buffer or the Perlin tables) and never uses them.
template <bool MOTION>
__device__ __forceinline__ void streamed_sweep_mma(const KernelArgs& a, TileStream& ts, const MotionCtx& mc, uint32_t* queue, const uint32_t* queue_base,
                                                   uint32_t* stage, bool active, float ox, float oy, float oz, float dx, float dy, float dz, float& hit_t,
                                                   int& hit_index, unsigned& flagged) {
    hit_t = kMaxT;
    hit_index = -1;
    if (threadIdx.x == 0) {
        ts.produce(a, 0);
        if (a.n_tiles > 1) ts.produce(a, 1);
    }
    MmaRayFrags f;
    mma_ray_fragments(stage, a.mma, active, ox, oy, oz, dx, dy, dz, f);
    const int tile_steps = a.tile_blocks / kLdsGroupBlocks;
    for (int tile = 0; tile < a.n_tiles; ++tile) {
        const int b = tile & 1;
        mbar_wait(&ts.full_bar[b], ts.full_phase[b]);
        ts.full_phase[b] ^= 1u;
        __syncwarp();
        const int first_step = tile * tile_steps;
        const int ns = min(tile_steps, a.n_steps - first_step);
        int cnt = 0;
        int ovf_step = (active && !f.in_range) ? 0 : first_step + ns;
        sweep_mma_steps(reinterpret_cast<const uint4*>(ts.tile_buf(b)), first_step, ns, f, queue, cnt, ovf_step);
        mbar_arrive(&ts.empty_bar[b]);
        __syncwarp();
        if (threadIdx.x == 0 && tile + 2 < a.n_tiles) ts.produce(a, tile + 2);
        sweep_mma_drain<MOTION>(a.blocks, mc, queue_base, cnt, ovf_step, first_step, first_step + ns, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
    }
}

// shared-memory layout of the streamed kernels: [tile 0][tile 1][Perlin tables][queues][pend][ray.time][MMA: transpose buffer]
template <bool MMA>
struct StreamedSmem {
    uint32_t tile_bytes;
    PerlinSmem* P;
    uint32_t* queue_base;
    uint32_t* queue;
    volatile uint32_t* pend;
    volatile float* tslot;
    uint32_t* stage;
    __device__ __forceinline__ StreamedSmem(unsigned char* raw, const KernelArgs& a) {
        tile_bytes = (uint32_t)a.tile_blocks * (MMA ? 128u : 64u);
        P = reinterpret_cast<PerlinSmem*>(raw + 2 * (size_t)tile_bytes);
        queue_base = reinterpret_cast<uint32_t*>(P + 1);
        queue = queue_base + threadIdx.x;  // [kQueueCap][kCtaThreads] candidate queues
        pend = queue + kQueueCap * kCtaThreads;
        tslot = reinterpret_cast<volatile float*>(pend + kCtaThreads);  // [kCtaThreads] ray.time per lane
        stage = queue_base + (kQueueCap + 2) * kCtaThreads;              // [kMmaStageRows][kCtaThreads], MMA only
    }
};

// (the tensor-path flavour states its two CTAs per SM: left alone ptxas squeezes it into 80 registers with 128 bytes of spills)
template <bool MOTION, bool MMA>
__global__ void __launch_bounds__(kCtaThreads, MMA ? 2 : 0) pt_megakernel_streamed(const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[2];
    __shared__ __align__(8) uint64_t empty_bar[2];
    __shared__ int cta_live;  // lanes not finished, recomputed every trip
    const StreamedSmem<MMA> sm(smem_raw, a);
    PerlinSmem* P = sm.P;
    uint32_t* queue = sm.queue;
    volatile uint32_t* pend = sm.pend;
    volatile float* tslot = sm.tslot;
    *pend = 0u;
    *tslot = 0.0f;
    const MotionCtx mc{a.motion, const_cast<const float*>(tslot), a.order};
    TileStream ts;
    ts.init(full_bar, empty_bar, smem_raw, sm.tile_bytes, MMA ? reinterpret_cast<const void*>(a.mma_image) : reinterpret_cast<const void*>(a.prefilter), MMA ? 128u : 64u);
    if (threadIdx.x == 0) cta_live = 0;
    stage_perlin(a, P);
    __syncthreads();

    const unsigned lane_id = threadIdx.x & 31u;
    Lane L;
    lane_init(L);
    unsigned long long rays = 0ULL;
    unsigned sweeps = 0u;
    for (;;) {
        L.pend = *pend != 0u;  // the streamed kernel parks these two in shared memory across the sweep (registers)
        L.time = *tslot;
        lane_refill<MOTION>(a, L, lane_id);
        *pend = L.pend ? 1u : 0u;
        *tslot = L.time;
        // CTA-wide liveness: every warp must keep consuming tiles while any warp still has work
        const bool warp_live = !__all_sync(kFullMask, L.finished);
        __syncthreads();  // previous trip's cta_live reads are done
        if (threadIdx.x == 0) cta_live = 0;
        __syncthreads();
        if (lane_id == 0 && warp_live) atomicAdd(&cta_live, 1);
        __syncthreads();
        if (cta_live == 0) break;

        float ox = L.o.x, oy = L.o.y, oz = L.o.z, dx = L.d.x, dy = L.d.y, dz = L.d.z;
        if (!L.active) {  // parked lane: |o|^2 = 1e36 dwarfs every L, d = 0 -> never a candidate
            ox = 0.0f; oy = 1.0e18f; oz = 0.0f;
            dx = dy = dz = 0.0f;
        }
        float hit_t;
        int hit_index;
        unsigned flagged = 0u;
        sweeps += 1u;  // every warp of the CTA sweeps every tile of every trip
        if (MMA) streamed_sweep_mma<MOTION>(a, ts, mc, queue, sm.queue_base, sm.stage, L.active, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
        else streamed_sweep<MOTION>(a, ts, mc, queue, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
        if (L.active) {
            rays += 1ULL;
            L.time = *tslot;  // (re-read: nothing of the lane's bookkeeping is carried across the sweep in a register)
            // the hit sphere's centre is no longer in shared memory: read it from the global SoA
            lane_shade<MOTION>(a, L, a.blocks, *P, mc, hit_t, hit_index);
        }
    }
    flush_ray_count(a, rays, lane_id, sweeps);
}

// pt_debug_hits, streamed scenes: caller-supplied rays through streamed_sweep (same tiles, barriers, operands, re-tests)
template <bool MOTION, bool MMA>
__global__ void PT_STREAM_LAUNCH_BOUNDS pt_debug_hits_streamed(const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[2];
    __shared__ __align__(8) uint64_t empty_bar[2];
    const StreamedSmem<MMA> sm(smem_raw, a);
    uint32_t* queue = sm.queue;
    volatile float* tslot = sm.tslot;
    const MotionCtx mc{a.motion, const_cast<const float*>(tslot), a.order};
    TileStream ts;
    ts.init(full_bar, empty_bar, smem_raw, sm.tile_bytes, MMA ? reinterpret_cast<const void*>(a.mma_image) : reinterpret_cast<const void*>(a.prefilter), MMA ? 128u : 64u);
    __syncthreads();
    // `base` is uniform within the CTA: all its threads run the same trips (the sweep is CTA-collective)
    for (uint32_t base = blockIdx.x * kCtaThreads; base < a.dbg_n; base += gridDim.x * kCtaThreads) {
        const uint32_t i = base + threadIdx.x;
        float ox = 0.0f, oy = 1.0e18f, oz = 0.0f, dx = 0.0f, dy = 0.0f, dz = 0.0f, time = 0.0f;
        if (i < a.dbg_n) {
            const float* ray = a.dbg_rays + (size_t)i * 6;
            ox = ray[0]; oy = ray[1]; oz = ray[2]; dx = ray[3]; dy = ray[4]; dz = ray[5];
            if (a.dbg_times) time = a.dbg_times[i];
        }
        *tslot = time;
        float hit_t;
        int hit_index;
        unsigned flagged = 0u;
        if (MMA) streamed_sweep_mma<MOTION>(a, ts, mc, queue, sm.queue_base, sm.stage, i < a.dbg_n, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
        else streamed_sweep<MOTION>(a, ts, mc, queue, ox, oy, oz, dx, dy, dz, hit_t, hit_index, flagged);
        if (i < a.dbg_n) {
            a.dbg_idx[i] = hit_index < 0 ? -1 : (a.order ? (int32_t)__ldg(a.order + hit_index) : hit_index);
            a.dbg_t[i] = hit_t;
            if (a.dbg_flagged) a.dbg_flagged[i] = flagged;
        }
    }
}

// =====================================================================================================
// Output stage: src/offline.rs:43-51 + src/math.rs:36-48 (row flip, linear->sRGB, 8-bit pack)
// =====================================================================================================
__global__ void pt_srgb8_kernel(const float* __restrict__ rgb, uint32_t width, uint32_t height, uint8_t* __restrict__ out) {
    const size_t n = (size_t)width * height;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t row = (uint32_t)(i / width);
        const uint32_t x = (uint32_t)(i - (size_t)row * width);
        const size_t src = ((size_t)(height - 1 - row) * width + x) * 3;  // `.rev()` over rows
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float lin = fmaxf(rgb[src + c], 0.0f);
            float s = 1.055f * powf(lin, 0.41666666f) - 0.055f;
            s = fminf(fmaxf(s, 0.0f), 1.0f);
            out[i * 3 + c] = (uint8_t)(s * 255.99f);
        }
    }
}

// Measurement helper: dependent-chain FFMA kernel, 8 independent accumulators per lane.
__global__ void pt_ffma_peak_kernel(float* out, int iters, float a, float b) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3f + (float)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rn(acc[i], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace pt
