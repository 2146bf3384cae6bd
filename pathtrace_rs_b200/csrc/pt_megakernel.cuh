// pt_megakernel.cuh — the persistent render kernel: device side of `Scene::update`
// (src/scene.rs:73-121) with `ray_trace` (src/scene.rs:49-71) unrolled into a per-lane state machine.
//
// Mapping:
//   * one lane = one pixel's path.  A lane owns a pixel for all of its `samples` (so the pixel's
//     xoshiro256+ stream is consumed in exactly the reference's order), then pulls the next pixel
//     from a global atomic queue (warp-aggregated) — the rayon `par_iter_mut` of scene.rs:90-93.
//   * every trip of the main loop is ONE full sphere sweep for all 32 lanes (convergent, FP32-bound),
//     followed by a short divergent shading step.  A lane whose path ended starts its next sample
//     (or next pixel) at the top of the next trip, so the sweep always runs with all live lanes.
//   * the recursion `emitted + attenuation * ray_trace(..)` becomes `colour += throughput * emitted;
//     throughput *= attenuation` carried in registers.
//   * sphere SoA is staged into shared memory once per CTA with TMA bulk copies (cp.async.bulk +
//     mbarrier); scenes that do not fit are streamed tile by tile (pt_megakernel_streamed below).
#pragma once
#include "pt_shade.cuh"
#include "pt_sweep.cuh"

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
// `pend` is that lane's "holding an unready ticket" flag; it lives in shared memory, and the end-of-chunk test uses the
// uniform chunk_mask rather than a per-lane bound: the kernel sits at 78 of the 80 registers that three CTAs per SM
// allow, so nothing that can live elsewhere is carried across the sweep in a register.
template <bool MOTION>
__device__ __forceinline__ void lane_refill(const KernelArgs& a, Lane& L, unsigned lane_id, volatile uint32_t* pend, volatile float* tslot) {
    bool want_pixel = false;
    if (!L.active && !L.finished) {
        if (L.have_pixel && ((L.sample & a.chunk_mask) == 0u || L.sample >= a.samples)) {
            if (L.sample >= a.samples)
                write_pixel(a, L);
            else
                publish_pixel_state(a, L);
            L.have_pixel = false;
        }
        want_pixel = !L.have_pixel && *pend == 0u;
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
                *pend = chunk != 0u ? 1u : 0u;
            } else {
                L.finished = true;
            }
        }
    }
    if (!L.active && !L.finished && !L.have_pixel) {
        if (*pend != 0u && acquire_pixel_state(a, L)) {  // not yet published: the lane sits this trip out and asks again
            *pend = 0u;
            L.have_pixel = true;
        }
    }
    if (!L.active && L.have_pixel) {
        // scene.rs:107-110
        const float u = ((float)L.px + rng_f32(L.rng)) * a.inv_nx;
        const float v = ((float)L.py + rng_f32(L.rng)) * a.inv_ny;
        float time;
        camera_get_ray(a.cam, u, v, L.rng, L.o, L.d, time);
        // ray.time is read only by MovingSphere (moving_sphere.rs:39) and is constant along a path (material.rs:62,83,117):
        // it lives in the lane's shared-memory slot, not in a register carried across the sweep
        if (MOTION) *tslot = time;
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
        const float s = (*mc.time - mo.time_start) * mo.inv_time_delta;
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


// Optional phase profile (compile with -DPT_PROFILE; tools/phase_profile.py): per-warp clock64 deltas of the
// three phases of a trip and trip/lane counters, summed into a global array.  Not compiled into the product.
#ifdef PT_PROFILE
__device__ unsigned long long g_prof[8];  // 0 refill clk, 1 sweep clk, 2 shade clk, 3 warp trips, 4 active lane-trips, 5 total clk
#define PT_PROF_DECL unsigned long long pf_t0 = 0, pf_refill = 0, pf_sweep = 0, pf_shade = 0, pf_trips = 0, pf_lanes = 0, pf_start = clock64();
#define PT_PROF_TICK() (pf_t0 = clock64())
#define PT_PROF_TOCK(acc) do { unsigned long long t_ = clock64(); acc += t_ - pf_t0; pf_t0 = t_; } while (0)
#define PT_PROF_FLUSH(lane_id) do { if ((lane_id) == 0) { atomicAdd(&g_prof[0], pf_refill); atomicAdd(&g_prof[1], pf_sweep); atomicAdd(&g_prof[2], pf_shade); \
    atomicAdd(&g_prof[3], pf_trips); atomicAdd(&g_prof[5], clock64() - pf_start); } atomicAdd(&g_prof[4], pf_lanes); } while (0)
#else
#define PT_PROF_DECL
#define PT_PROF_TICK()
#define PT_PROF_TOCK(acc)
#define PT_PROF_FLUSH(lane_id)
#endif


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
#ifndef PT_REGROUP
#define PT_REGROUP 1
#endif
constexpr int kRegroupCats = 6;
constexpr int kRegroupWords = 28;  // 8 rng + 6 ray + 3 thr + 3 col + px, py, sample, depth, flags, hit_t, hit_index, time
enum { CAT_LAMBERT_CONST = 0, CAT_LAMBERT_TEX = 1, CAT_METAL = 2, CAT_DIELECTRIC = 3, CAT_ENDING = 4, CAT_IDLE = 5 };

__device__ __forceinline__ int lane_category(const KernelArgs& a, const Lane& L, int hit_index) {
    if (!L.active) return CAT_IDLE;
    if (hit_index < 0 || L.depth >= a.max_depth) return CAT_ENDING;
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(a.shade + hit_index) + 1);
    const int kind = __float_as_int(s1.y);
    if (kind == MAT_LAMBERTIAN) return __float_as_int(s1.z) < 0 ? CAT_LAMBERT_CONST : CAT_LAMBERT_TEX;
    if (kind == MAT_METAL) return CAT_METAL;
    if (kind == MAT_DIELECTRIC) return CAT_DIELECTRIC;
    return CAT_ENDING;  // DiffuseLight
}

// xchg: [kRegroupWords][kCtaThreads] words; cat_count: this trip's [kRegroupCats] counters (two sets alternate: the set
// used by trip t is cleared after trip t's second barrier and next touched after trip t+1's first barrier).
// Two CTA barriers per trip; the second also ORs "some lane still has work" over the CTA and returns it.  Trip t+1's
// first barrier separates trip t's reads of xchg from trip t+1's writes.
__device__ __forceinline__ bool cta_regroup(const KernelArgs& a, Lane& L, float& hit_t, int& hit_index, volatile uint32_t* pend,
                                            volatile float* tslot, uint32_t* __restrict__ xchg, uint32_t* __restrict__ cat_count,
                                            unsigned lane_id) {
    const int cat = lane_category(a, L, hit_index);
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
    __syncthreads();  // all counts are final
    unsigned base = 0u;
#pragma unroll
    for (int c = 0; c < kRegroupCats - 1; ++c) base += (c < cat) ? cat_count[c] : 0u;
    const unsigned dest = base + warp_off + (unsigned)__popc(mine & ((1u << lane_id) - 1u));
    uint32_t* w = xchg + dest;
    const uint32_t flags = (L.active ? 1u : 0u) | (L.have_pixel ? 2u : 0u) | (L.finished ? 4u : 0u) | (*pend != 0u ? 8u : 0u);
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
    w[27 * kCtaThreads] = __float_as_uint(*tslot);
    const bool live = __syncthreads_or(L.finished ? 0 : 1) != 0;  // every path is in its new slot
    if (threadIdx.x < kRegroupCats) cat_count[threadIdx.x] = 0u;
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
    *pend = (f >> 3) & 1u;
    hit_t = __uint_as_float(r[25 * kCtaThreads]);
    hit_index = (int)r[26 * kCtaThreads];
    *tslot = __uint_as_float(r[27 * kCtaThreads]);
    return live;
}

// =====================================================================================================
// Resident variant: the whole sphere SoA lives in shared memory for the life of the CTA.
// =====================================================================================================
#ifndef PT_EXACT_SMEM
#define PT_EXACT_SMEM 0
#endif
#ifdef PT_RES_PIPE
constexpr bool kResidentPipe = true;  // LDS one block ahead in the resident kernel too (needs the registers: see PT_LDS_MIN_CTAS)
#else
constexpr bool kResidentPipe = false;
#endif
#ifdef PT_LDS_MIN_CTAS
#define PT_LDS_LAUNCH_BOUNDS __launch_bounds__(kCtaThreads, PT_LDS_MIN_CTAS)
#else
#define PT_LDS_LAUNCH_BOUNDS __launch_bounds__(kCtaThreads)
#endif
// MOTION: the scene has Hitable::MovingSphere entries (compiled out of the static instantiation)
template <bool MOTION>
__global__ void PT_LDS_LAUNCH_BOUNDS pt_megakernel_resident(const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    float4* pf = reinterpret_cast<float4*>(smem_raw);  // pre-filter image of the whole scene
#if PT_EXACT_SMEM
    // exact blocks right behind it: the candidate re-tests and the hit epilogue read them with LDS instead of LDG
    const float4* ex = reinterpret_cast<const float4*>(smem_raw + (size_t)a.n_blocks * 64);
    PerlinSmem* P = reinterpret_cast<PerlinSmem*>(smem_raw + (size_t)a.n_blocks * 128);
#else
    const float4* ex = a.blocks;  // exact blocks stay in global/L2
    PerlinSmem* P = reinterpret_cast<PerlinSmem*>(smem_raw + (size_t)a.n_blocks * 64);
#endif
    uint32_t* queue = reinterpret_cast<uint32_t*>(P + 1) + threadIdx.x;  // [kQueueCap][kCtaThreads] candidate queues
    volatile uint32_t* pend = queue + kQueueCap * kCtaThreads;
    volatile float* tslot = reinterpret_cast<volatile float*>(pend + kCtaThreads);  // [kCtaThreads] ray.time per lane
    uint32_t* xchg = const_cast<uint32_t*>(reinterpret_cast<volatile uint32_t*>(tslot)) - threadIdx.x + kCtaThreads;  // [kRegroupWords][kCtaThreads]
    uint32_t* cat_count = xchg + kRegroupWords * kCtaThreads;                                                          // [kRegroupCats]
    *pend = 0u;
    *tslot = 0.0f;
    if (threadIdx.x < 16) cat_count[threadIdx.x] = 0u;  // two sets of kRegroupCats counters, 8 words apart
    const MotionCtx mc{a.motion, tslot, a.order};

    const uint32_t bytes = (uint32_t)a.n_blocks * 64u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && bytes != 0u) {
#if PT_EXACT_SMEM
        mbar_arrive_expect_tx(&bar, 2u * bytes);
        tma_bulk_g2s_chunked(pf, a.prefilter, bytes, &bar);
        tma_bulk_g2s_chunked(const_cast<float4*>(ex), a.blocks, bytes, &bar);
#else
        mbar_arrive_expect_tx(&bar, bytes);
        tma_bulk_g2s_chunked(pf, a.prefilter, bytes, &bar);
#endif
    }
    stage_perlin(a, P);
    __syncthreads();
    if (bytes != 0u) mbar_wait(&bar, 0);

    const unsigned lane_id = threadIdx.x & 31u;
    Lane L;
    lane_init(L);
    unsigned long long rays = 0ULL;
    unsigned sweeps = 0u;
    PT_PROF_DECL

#if PT_REGROUP
    lane_refill<MOTION>(a, L, lane_id, pend, tslot);
#endif
    for (uint32_t trip = 0;; ++trip) {
        PT_PROF_TICK();
#if !PT_REGROUP
        lane_refill<MOTION>(a, L, lane_id, pend, tslot);
        if (__all_sync(kFullMask, L.finished)) break;
#endif
        float ox = L.o.x, oy = L.o.y, oz = L.o.z, dx = L.d.x, dy = L.d.y, dz = L.d.z;
        if (!L.active) {  // parked lane: |o|^2 = 1e36 dwarfs every L, d = 0 -> never a candidate
            ox = 0.0f; oy = 1.0e18f; oz = 0.0f;
            dx = dy = dz = 0.0f;
        }
        float hit_t = kMaxT;
        int hit_index = -1;
        __syncwarp();
        PT_PROF_TOCK(pf_refill);
        if (__any_sync(kFullMask, L.active)) {  // regrouping collects idle lanes in whole warps: they skip the sweep
            sweeps += 1u;
            const float nod = -((ox * dx + oy * dy) + oz * dz);
            const float oo = ((ox * ox + oy * oy) + oz * oz) * (1.0f - 1.9073486328125e-06f);
            int cnt = 0;
            sweep_expanded<kResidentPipe, MOTION>(pf, a.n_blocks, 0, ex, mc, queue, cnt, ox, oy, oz, dx, dy, dz, nod, ox + ox, oy + oy, oz + oz, oo, hit_t, hit_index);
            sweep_drain<MOTION, true>(ex, mc, queue, cnt, ox, oy, oz, dx, dy, dz, hit_t, hit_index);
        }
        __syncwarp();
        PT_PROF_TOCK(pf_sweep);
#if PT_REGROUP
        if (!cta_regroup(a, L, hit_t, hit_index, pend, tslot, xchg, cat_count + (trip & 1u) * 8u, lane_id)) break;
#endif
        if (L.active) {
            rays += 1ULL;  // scene.rs:57
            lane_shade<MOTION>(a, L, ex, *P, mc, hit_t, hit_index);
        }
#if PT_REGROUP
        lane_refill<MOTION>(a, L, lane_id, pend, tslot);  // ended paths sit side by side now: next sample / next ticket together
#endif
        __syncwarp();
        PT_PROF_TOCK(pf_shade);
#ifdef PT_PROFILE
        pf_trips += 1;
#endif
    }
#ifdef PT_PROFILE
    pf_lanes = rays;
#endif
    PT_PROF_FLUSH(lane_id);
    flush_ray_count(a, rays, lane_id, sweeps);
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
template <bool MOTION>
__global__ void PT_STREAM_LAUNCH_BOUNDS pt_megakernel_streamed(const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[2];
    __shared__ __align__(8) uint64_t empty_bar[2];
    __shared__ int cta_live;  // lanes not finished, recomputed every trip
    const uint32_t tile_bytes = (uint32_t)a.tile_blocks * 64u;
    // tile buffer b lives at smem_raw + b * tile_bytes (computed, not looked up: an array of pointers would be demoted to
    // local memory and the sweep's loads would lose their shared-memory address space)
    auto tile_buf = [&](int b) { return reinterpret_cast<float4*>(smem_raw + (size_t)b * tile_bytes); };
    PerlinSmem* P = reinterpret_cast<PerlinSmem*>(smem_raw + 2 * (size_t)tile_bytes);
    uint32_t* queue = reinterpret_cast<uint32_t*>(P + 1) + threadIdx.x;  // [kQueueCap][kCtaThreads] candidate queues
    volatile uint32_t* pend = queue + kQueueCap * kCtaThreads;
    volatile float* tslot = reinterpret_cast<volatile float*>(pend + kCtaThreads);  // [kCtaThreads] ray.time per lane
    *pend = 0u;
    *tslot = 0.0f;
    const MotionCtx mc{a.motion, tslot, a.order};

    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        mbar_init(&empty_bar[0], kCtaThreads);  // every lane releases the buffer itself (no elected lane: each thread's
        mbar_init(&empty_bar[1], kCtaThreads);  // own reads are ordered before its own arrive)
        fence_mbar_init();
        cta_live = 0;
    }
    stage_perlin(a, P);
    __syncthreads();

    const unsigned lane_id = threadIdx.x & 31u;
    Lane L;
    lane_init(L);
    unsigned long long rays = 0ULL;
    unsigned sweeps = 0u;
    uint32_t full_phase[2] = {0u, 0u};   // parity to wait for on full_bar[b]
    uint32_t empty_phase[2] = {0u, 0u};  // producer side: parity to wait for on empty_bar[b]
    uint32_t produced[2] = {0u, 0u};     // producer: how many times buffer b has been filled

    auto produce = [&](int tile) {  // thread 0 only
        const int b = tile & 1;
        if (produced[b] != 0u) {  // wait until every warp has released the previous contents
            mbar_wait(&empty_bar[b], empty_phase[b]);
            empty_phase[b] ^= 1u;
        }
        produced[b] += 1u;
        const int first = tile * a.tile_blocks;
        const int nb = min(a.tile_blocks, a.n_blocks - first);
        const uint32_t bytes = (uint32_t)nb * 64u;
        mbar_arrive_expect_tx(&full_bar[b], bytes);
        tma_bulk_g2s_chunked(tile_buf(b), a.prefilter + (size_t)first * 4, bytes, &full_bar[b]);
    };

    for (;;) {
        lane_refill<MOTION>(a, L, lane_id, pend, tslot);
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
        float hit_t = kMaxT;
        int hit_index = -1;
        const float nod = -((ox * dx + oy * dy) + oz * dz);
        const float oo = ((ox * ox + oy * oy) + oz * oz) * (1.0f - 1.9073486328125e-06f);
        int cnt = 0;
        sweeps += 1u;  // every warp of the CTA sweeps every tile of every trip
        if (threadIdx.x == 0) {
            produce(0);
            if (a.n_tiles > 1) produce(1);
        }
        for (int tile = 0; tile < a.n_tiles; ++tile) {
            const int b = tile & 1;
            mbar_wait(&full_bar[b], full_phase[b]);
            full_phase[b] ^= 1u;
            __syncwarp();
            const int first = tile * a.tile_blocks;
            const int nb = min(a.tile_blocks, a.n_blocks - first);
            sweep_expanded<kStreamPipe, MOTION>(tile_buf(b), nb, first, a.blocks, mc, queue, cnt, ox, oy, oz, dx, dy, dz, nod, ox + ox, oy + oy, oz + oz, oo, hit_t, hit_index);
            mbar_arrive(&empty_bar[b]);
            __syncwarp();
            if (threadIdx.x == 0 && tile + 2 < a.n_tiles) produce(tile + 2);
            // this tile's candidates: exact re-test against the global SoA (L2), off the tile buffer's critical path
            sweep_drain<MOTION, false>(a.blocks, mc, queue, cnt, ox, oy, oz, dx, dy, dz, hit_t, hit_index);
        }
        if (L.active) {
            rays += 1ULL;
            // the hit sphere's centre is no longer in shared memory: read it from the global SoA
            lane_shade<MOTION>(a, L, a.blocks, *P, mc, hit_t, hit_index);
        }
    }
    flush_ray_count(a, rays, lane_id, sweeps);
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
