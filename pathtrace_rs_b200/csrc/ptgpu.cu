// ptgpu.cu — host side of libptgpu.so: the C ABI declared in include/ptgpu.h.
//
// Scene flattening (the GPU arm of `Params::new_scene`, src/params.rs:29-46, plus `SpheresSoA::new`,
// src/collision/spheres_soa.rs:26-74), kernel launch (`Scene::update`, src/scene.rs:73-121) and the
// host<->device traffic around it.  No CPU fallback: every entry point fails without a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ptgpu.h"
#include "pt_megakernel.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define PT_CUDA(call)                                                                                        \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return fail(PT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr size_t kMaxDynSmem = 227 * 1024 - 1024;  // leave room for static shared + driver reservation

}  // namespace

struct PtScene {
    int device = 0;
    int sm_count = 0;
    uint32_t n_spheres = 0;
    int n_blocks = 0;
    bool has_noise = false;
    bool has_sky = false;
    pt::V3 sky{0, 0, 0};
    float4* d_blocks = nullptr;
    pt::DevShade* d_shade = nullptr;
    pt::DevTexture* d_tex = nullptr;
    uint8_t* d_images = nullptr;  // RGB8 pool of the Image textures (nullptr: none)
    uint32_t* d_order = nullptr;  // stored sphere index -> position in the caller's list (nullptr: stored in list order)
    pt::PerlinSmem* d_perlin = nullptr;
    pt::DevMotion* d_motion = nullptr;  // MovingSphere records (nullptr: none)
    float motion_t_lo = 0.0f, motion_t_hi = 0.0f;  // intersection of the moving spheres' [time0, time1]
    float4* d_prefilter = nullptr;  // pre-filter image X,Y,Z,K per block (LDS kernels stage/stream it)
    // per-render scratch
    unsigned long long* d_ray_count = nullptr;  // [0] ray count
    unsigned long long* d_sweep_count = nullptr;  // warp-level sweeps of the last launch (lane-efficiency diagnostic)
    unsigned int* d_next_pixel = nullptr;
    uint32_t* d_pixstate = nullptr;  // chunk queue: 12 words per owned pixel (pt_megakernel.cuh, PixState)
    size_t d_pixstate_pixels = 0;
    float* d_rgb = nullptr;  // device image for the host-buffer entry points
    size_t d_rgb_floats = 0;
    uint8_t* d_rgb8 = nullptr;
    size_t d_rgb8_bytes = 0;
    // pt_render_progressive: what the resident image currently holds
    bool prog_valid = false;
    uint32_t prog_w = 0, prog_h = 0, prog_next_frame = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // launch geometry
    bool resident = true;
    int tile_blocks = 0, n_tiles = 0;
    size_t smem_bytes = 0;
    int ctas_per_sm = 0;
    PtRenderStats stats{};
};

namespace {

int check_device(int device, cudaDeviceProp* prop_out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(PT_ERR_NO_DEVICE, "no CUDA device: %s", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(PT_ERR_INVALID, "device %d out of range (have %d)", device, n);
    cudaDeviceProp prop;
    PT_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(PT_ERR_NO_DEVICE, "device %d is sm_%d%d; libptgpu is built for sm_100a only", device, prop.major, prop.minor);
    if (prop_out) *prop_out = prop;
    return PT_OK;
}

// kernel selection + shared-memory sizing for a scene
template <typename K>
int configure_kernel(K kernel, size_t smem, int* ctas_per_sm) {
    PT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, pt::kCtaThreads, smem));
    if (occ < 1) return fail(PT_ERR_TOO_LARGE, "kernel does not fit on an SM with %zu bytes of shared memory", smem);
    *ctas_per_sm = occ;
    return PT_OK;
}

int configure_streamed(PtScene* s) {
    return s->d_motion ? configure_kernel(pt::pt_megakernel_streamed<true>, s->smem_bytes, &s->ctas_per_sm)
                       : configure_kernel(pt::pt_megakernel_streamed<false>, s->smem_bytes, &s->ctas_per_sm);
}

// shared-memory budget of the two kernels (one place: plan_launch and the storage-order decision both ask)
constexpr size_t kQueueBytes = (size_t)pt::kQueueCap * pt::kCtaThreads * sizeof(uint32_t);  // per-lane candidate queues
constexpr size_t kPerlinBytes = sizeof(pt::PerlinSmem) + 2 * pt::kCtaThreads * sizeof(uint32_t) + kQueueBytes;  // Perlin tables + `pend` words + ray.time slots + queues
constexpr size_t kRegroupBytes = (size_t)pt::kRegroupWords * pt::kCtaThreads * sizeof(uint32_t) + 64;  // path-state exchange + category counters
int forced_stream_tile() {  // test hook: PTGPU_FORCE_STREAM_TILE_BLOCKS=<n> runs any scene through the streamed kernel with n-block tiles
    const char* env = std::getenv("PTGPU_FORCE_STREAM_TILE_BLOCKS");
    return env ? std::atoi(env) : 0;
}
bool fits_resident(int n_blocks) {
    if (forced_stream_tile() > 0 && n_blocks > 0) return false;
    const size_t exact_bytes = PT_EXACT_SMEM ? (size_t)n_blocks * 64 : 0;  // resident kernel: exact blocks in shared memory too
    return (size_t)n_blocks * 64 + kPerlinBytes + kRegroupBytes + exact_bytes <= kMaxDynSmem;
}

int plan_launch(PtScene* s) {
    const size_t perlin_bytes = kPerlinBytes;
    const size_t all = (size_t)s->n_blocks * 64 + perlin_bytes;
    int forced_tile = forced_stream_tile();
    if (forced_tile > 0 && s->n_blocks > 0) {
        s->resident = false;
        forced_tile = (forced_tile + pt::kLdsGroupBlocks - 1) / pt::kLdsGroupBlocks * pt::kLdsGroupBlocks;  // whole groups
        s->tile_blocks = std::min(forced_tile, s->n_blocks);
        s->n_tiles = (s->n_blocks + s->tile_blocks - 1) / s->tile_blocks;
        s->smem_bytes = 2 * (size_t)s->tile_blocks * 64 + perlin_bytes;
        return configure_streamed(s);
    }
    if (fits_resident(s->n_blocks)) {
        s->resident = true;
        s->smem_bytes = all + kRegroupBytes + (PT_EXACT_SMEM ? (size_t)s->n_blocks * 64 : 0);
        s->tile_blocks = s->n_blocks;
        s->n_tiles = 1;
        return s->d_motion ? configure_kernel(pt::pt_megakernel_resident<true>, s->smem_bytes, &s->ctas_per_sm)
                           : configure_kernel(pt::pt_megakernel_resident<false>, s->smem_bytes, &s->ctas_per_sm);
    }
    // streamed: two tile buffers; 2 CTAs per SM keeps the FP32 pipe fed while one CTA waits on a barrier
    s->resident = false;
    int stream_ctas = 2;
    if (const char* env = std::getenv("PTGPU_STREAM_CTAS")) stream_ctas = std::max(1, std::min(4, std::atoi(env)));  // tuning hook
    const size_t per_cta = (kMaxDynSmem + 1024) / stream_ctas - 2048;
    const size_t tile_bytes = ((per_cta - perlin_bytes) / 2) & ~(size_t)1023;
    s->tile_blocks = (int)(tile_bytes / 64);
    s->n_tiles = (s->n_blocks + s->tile_blocks - 1) / s->tile_blocks;
    s->smem_bytes = 2 * (size_t)s->tile_blocks * 64 + perlin_bytes;
    return configure_streamed(s);
}

uint32_t owned_rows(uint32_t height, const PtPartition& p) {
    if (p.part_count <= 1) return height;
    const uint32_t n_tiles = (height + p.tile_rows - 1) / p.tile_rows;
    uint32_t rows = 0;
    for (uint32_t k = p.part_index; k < n_tiles; k += p.part_count) rows += std::min(p.tile_rows, height - k * p.tile_rows);
    return rows;
}

int normalise_partition(const PtPartition* in, PtPartition* out) {
    PtPartition p{4, 0, 1, 0};
    if (in) {
        p = *in;
        if (p.tile_rows == 0) p.tile_rows = 4;
        if (p.part_count == 0) p.part_count = 1;
        if (p.part_index >= p.part_count) return fail(PT_ERR_INVALID, "partition index %u >= count %u", p.part_index, p.part_count);
    }
    *out = p;
    return PT_OK;
}

int validate_params(const PtParams* params, const PtCamera* camera) {
    if (!params || !camera) return fail(PT_ERR_INVALID, "null params/camera");
    if (params->width == 0 || params->height == 0) return fail(PT_ERR_INVALID, "zero-sized image %ux%u", params->width, params->height);
    if ((uint64_t)params->width * params->height > 0xffffffffULL / 4) return fail(PT_ERR_TOO_LARGE, "image too large");
    if (params->use_bvh) return fail(PT_ERR_UNSUPPORTED, "use_bvh is not supported on the GPU path (flat sphere list only, params.rs:36-43)");
    return PT_OK;
}

// enqueue one Scene::update on `stream`; d_rgb is the full-size device image
int launch_update(PtScene* s, const PtParams* params, const PtCamera* cam, uint32_t frame_num, const PtPartition& part,
                  float* d_rgb, unsigned long long* d_ray_count, cudaStream_t stream) {
    pt::KernelArgs a{};
    a.blocks = s->d_blocks;
    a.n_blocks = s->n_blocks;
    a.n_spheres = (int)s->n_spheres;
    a.shade = s->d_shade;
    a.tex = s->d_tex;
    a.images = s->d_images;
    a.order = s->d_order;
    a.perlin = s->d_perlin;
    a.prefilter = s->d_prefilter;
    a.motion = s->d_motion;
    a.has_noise = s->has_noise ? 1 : 0;
    auto V = [](const float* f) { return pt::V3{f[0], f[1], f[2]}; };
    a.cam.origin = V(cam->origin);
    a.cam.llc = V(cam->lower_left_corner);
    a.cam.horizontal = V(cam->horizontal);
    a.cam.vertical = V(cam->vertical);
    a.cam.u = V(cam->u);
    a.cam.v = V(cam->v);
    a.cam.time0 = cam->time0;
    a.cam.time1 = cam->time1;
    a.cam.lens_radius = cam->lens_radius;
    a.width = params->width;
    a.height = params->height;
    a.samples = params->samples;
    a.max_depth = params->max_depth;
    a.frame_num = frame_num;
    // scene.rs:82-87
    a.inv_nx = 1.0f / (float)params->width;
    a.inv_ny = 1.0f / (float)params->height;
    a.inv_ns = 1.0f / (float)params->samples;
    a.mix_prev = (float)frame_num / (float)(frame_num + 1);
    a.mix_new = 1.0f - a.mix_prev;
    a.has_sky = s->has_sky ? 1 : 0;
    a.sky = s->sky;
    a.tile_rows = part.tile_rows;
    a.part_index = part.part_index;
    a.part_count = part.part_count;
    a.n_owned_pixels = owned_rows(params->height, part) * params->width;
    a.random_seed = params->random_seed ? 1 : 0;
    a.seed_salt = params->seed_salt;
    a.rgb = d_rgb;
    a.ray_count = d_ray_count;
    a.sweep_count = s->d_sweep_count;
    a.next_pixel = s->d_next_pixel;
    a.tile_blocks = s->tile_blocks;
    a.n_tiles = s->n_tiles;

    if (s->d_motion && !(cam->time0 >= s->motion_t_lo && cam->time1 <= s->motion_t_hi && cam->time0 <= cam->time1))
        return fail(PT_ERR_UNSUPPORTED, "camera shutter [%g, %g] is not inside the moving spheres' interval [%g, %g] (the pre-filter bounds their sweep over that interval)",
                    cam->time0, cam->time1, s->motion_t_lo, s->motion_t_hi);
    PT_CUDA(cudaMemsetAsync(s->d_next_pixel, 0, sizeof(unsigned int), stream));
    PT_CUDA(cudaMemsetAsync(d_ray_count, 0, sizeof(unsigned long long), stream));
    PT_CUDA(cudaMemsetAsync(s->d_sweep_count, 0, sizeof(unsigned long long), stream));
    s->stats.kernel_launches = 0;
    s->stats.grid_ctas = 0;
    if (a.n_owned_pixels == 0) return PT_OK;

    // persistent grid: one wave of CTAs, never more lanes than pixels
    const uint32_t want = (a.n_owned_pixels + pt::kCtaThreads - 1) / pt::kCtaThreads;
    const uint32_t grid = std::min<uint32_t>((uint32_t)(s->sm_count * s->ctas_per_sm), want);

    // chunk queue (pt_megakernel.cuh, lane_refill): samples are handed out in chunks, sample-major, so that all pixels
    // finish together.  With fewer than two pixels per lane every pixel starts at once and the launch lasts as long as
    // its slowest pixel whatever the unit: one chunk per pixel then, which skips the state table altogether.
    uint32_t chunk = 0;  // 0 = one chunk per pixel
    const uint64_t lanes = (uint64_t)grid * pt::kCtaThreads;
    if (params->samples > 8 && (uint64_t)a.n_owned_pixels >= 2 * lanes) {
        const uint32_t max_chunks = std::max<uint32_t>(1u, std::min<uint32_t>(64u, 0xF0000000u / a.n_owned_pixels));
        const uint32_t at_least = std::max<uint32_t>((params->samples + max_chunks - 1) / max_chunks, 8u);
        chunk = 8;
        while (chunk < at_least) chunk *= 2;
    }
    if (const char* env = std::getenv("PTGPU_CHUNK_SAMPLES")) {  // test/tuning hook: a power of two, 0 = whole pixels
        const long v = std::atol(env);
        chunk = 0;
        if (v > 0) {
            chunk = 1;
            while ((long)chunk < v && chunk < (1u << 30)) chunk *= 2;
        }
    }
    uint32_t n_chunks = 1;
    if (chunk != 0 && chunk < params->samples) {
        n_chunks = (params->samples + chunk - 1) / chunk;
        if ((uint64_t)n_chunks * a.n_owned_pixels > 0xF0000000ull)
            return fail(PT_ERR_TOO_LARGE, "%u chunks x %u pixels overflow the ticket counter", n_chunks, a.n_owned_pixels);
        a.chunk_samples = chunk;
        a.chunk_mask = chunk - 1;
    } else {
        a.chunk_samples = std::max<uint32_t>(params->samples, 1u);
        a.chunk_mask = 0xffffffffu;
    }
    a.n_tickets = n_chunks * a.n_owned_pixels;
    a.pixstate = nullptr;
    if (n_chunks > 1) {
        if (s->d_pixstate_pixels < a.n_owned_pixels) {
            if (s->d_pixstate) cudaFree(s->d_pixstate);
            s->d_pixstate = nullptr;
            s->d_pixstate_pixels = 0;
            PT_CUDA(cudaMalloc(&s->d_pixstate, (size_t)a.n_owned_pixels * pt::kPixStateWords * sizeof(uint32_t)));
            s->d_pixstate_pixels = a.n_owned_pixels;
        }
        // word 11 of every record (= samples completed) must read 0 before the first chunk is published
        PT_CUDA(cudaMemsetAsync(s->d_pixstate, 0, (size_t)a.n_owned_pixels * pt::kPixStateWords * sizeof(uint32_t), stream));
        a.pixstate = s->d_pixstate;
    }
    if (s->resident) {
        if (s->d_motion) pt::pt_megakernel_resident<true><<<grid, pt::kCtaThreads, s->smem_bytes, stream>>>(a);
        else pt::pt_megakernel_resident<false><<<grid, pt::kCtaThreads, s->smem_bytes, stream>>>(a);
    } else {
        if (s->d_motion) pt::pt_megakernel_streamed<true><<<grid, pt::kCtaThreads, s->smem_bytes, stream>>>(a);
        else pt::pt_megakernel_streamed<false><<<grid, pt::kCtaThreads, s->smem_bytes, stream>>>(a);
    }
    PT_CUDA(cudaGetLastError());
    s->stats.kernel_launches = 1;
    s->stats.grid_ctas = grid;
    s->stats.cta_threads = pt::kCtaThreads;
    s->stats.smem_bytes = (uint32_t)s->smem_bytes;
    s->stats.resident = s->resident ? 1u : 0u;
    s->stats.n_spheres = s->n_spheres;
    return PT_OK;
}

int ensure_image(PtScene* s, size_t floats) {
    if (s->d_rgb_floats >= floats) return PT_OK;
    if (s->d_rgb) cudaFree(s->d_rgb);
    s->d_rgb = nullptr;
    s->d_rgb_floats = 0;
    PT_CUDA(cudaMalloc(&s->d_rgb, floats * sizeof(float)));
    s->d_rgb_floats = floats;
    return PT_OK;
}

// copy the rows a partition owns between host and device images (same layout on both sides)
int copy_owned_rows(PtScene* s, const PtParams* params, const PtPartition& part, float* host, bool to_device, uint64_t* bytes_out) {
    const size_t row_bytes = (size_t)params->width * 3 * sizeof(float);
    uint64_t bytes = 0;
    if (part.part_count <= 1) {
        bytes = row_bytes * params->height;
        PT_CUDA(cudaMemcpyAsync(to_device ? (void*)s->d_rgb : (void*)host, to_device ? (const void*)host : (const void*)s->d_rgb, bytes,
                                to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, s->stream));
    } else {
        const uint32_t n_tiles = (params->height + part.tile_rows - 1) / part.tile_rows;
        for (uint32_t k = part.part_index; k < n_tiles; k += part.part_count) {
            const uint32_t r0 = k * part.tile_rows;
            const uint32_t nr = std::min(part.tile_rows, params->height - r0);
            const size_t off = (size_t)r0 * params->width * 3;
            const size_t nbytes = row_bytes * nr;
            PT_CUDA(cudaMemcpyAsync(to_device ? (void*)(s->d_rgb + off) : (void*)(host + off),
                                    to_device ? (const void*)(host + off) : (const void*)(s->d_rgb + off), nbytes,
                                    to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, s->stream));
            bytes += nbytes;
        }
    }
    *bytes_out = bytes;
    return PT_OK;
}

// Storage order of a scene's spheres (see pt_scene_create).  Fills order_of[j] = position in the caller's list of the
// sphere stored at j and returns the mode: 0 the caller's order, 1 Morton, 2 large spheres first, then Morton.
int storage_order(const PtSceneDesc* desc, bool any_moving, std::vector<uint32_t>& order_of) {
    const uint32_t n = desc->n_spheres;
    auto is_moving = [&](uint32_t i) { return any_moving && desc->motion[i].moving != 0; };
    // ---- storage order.  The sweep flags candidates per group of 16 consecutive spheres, and a ray's candidates are
    // spatially coherent, so spheres are stored along a Morton curve through their centres: a group becomes a compact
    // patch instead of a strip of the caller's list, and a ray (and the 32 neighbouring rays of its warp) touches fewer
    // groups.  order_of[j] = position in the caller's list of the sphere stored at j; hits are decided exactly as before
    // (same expression per sphere, equal-t ties to the lower ORIGINAL position), so images do not change.
    order_of.resize(n);
    for (uint32_t i = 0; i < n; ++i) order_of[i] = i;
    int order_mode = n > 64 ? 2 : 0;  // 0: the caller's order; 1: Morton; 2: large spheres first, then Morton
    if (const char* env = std::getenv("PTGPU_SPATIAL_ORDER")) order_mode = n > 1 ? std::max(0, std::min(2, std::atoi(env))) : 0;  // tuning hook
    // resident kernel only: the streamed kernel keeps list order and the plain index tie rule (with ~10^5 spheres few groups
    // are flagged anyway: 66.0 -> 66.5 % on cfg5, and the out-of-line tie rule costs that kernel more than it gains)
    const int n_blocks_planned = (int)(((n + 3) / 4 + pt::kLdsGroupBlocks - 1) / pt::kLdsGroupBlocks * pt::kLdsGroupBlocks);
    if (!fits_resident(n_blocks_planned)) order_mode = 0;
#ifdef PT_RES_PIPE
    order_mode = 0;  // that experimental build instantiates the resident sweep without the ordered tie rule
#endif
    if (order_mode != 0) {
        auto centre_of = [&](uint32_t i, int axis) -> double {
            const float* c = axis == 0 ? desc->centre_x : (axis == 1 ? desc->centre_y : desc->centre_z);
            double v = c[i];
            if (is_moving(i)) v += 0.5 * ((double)desc->motion[i].centre1[axis] - v);
            return v;
        };
        double lo[3], hi[3];
        for (int ax = 0; ax < 3; ++ax) {  // robust bounds: the 2nd..98th percentile of the centres (a 1000-radius ground sphere must not stretch the grid)
            std::vector<double> v(n);
            for (uint32_t i = 0; i < n; ++i) v[i] = centre_of(i, ax);
            std::sort(v.begin(), v.end());
            lo[ax] = v[(size_t)(0.02 * (n - 1))];
            hi[ax] = v[(size_t)(0.98 * (n - 1))];
            if (!(hi[ax] > lo[ax])) hi[ax] = lo[ax] + 1.0;
        }
        auto spread = [](uint32_t v) {  // 10 bits -> every third bit
            v &= 0x3ffu;
            v = (v | (v << 16)) & 0x030000ffu;
            v = (v | (v << 8)) & 0x0300f00fu;
            v = (v | (v << 4)) & 0x030c30c3u;
            v = (v | (v << 2)) & 0x09249249u;
            return v;
        };
        std::vector<uint32_t> code(n);
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t q[3];
            for (int ax = 0; ax < 3; ++ax) {
                double t = (centre_of(i, ax) - lo[ax]) / (hi[ax] - lo[ax]);
                t = std::isfinite(t) ? std::min(1.0, std::max(0.0, t)) : 0.0;
                q[ax] = (uint32_t)(t * 1023.0);
            }
            code[i] = spread(q[0]) | (spread(q[1]) << 1) | (spread(q[2]) << 2);
        }
        // spheres much larger than the typical one (the ground, the three big RTIOW spheres) are candidates for a large
        // share of all rays wherever they are stored: keep them together in the leading group(s) instead of letting each
        // of them turn another group into an "always flagged" one
        std::vector<uint8_t> large(n, 0);
        if (order_mode == 2) {
            std::vector<float> radii(n);
            for (uint32_t i = 0; i < n; ++i) radii[i] = std::fabs(desc->radius[i]);
            std::nth_element(radii.begin(), radii.begin() + n / 2, radii.end());
            const float median = radii[n / 2];
            for (uint32_t i = 0; i < n; ++i) large[i] = std::fabs(desc->radius[i]) > 3.0f * median ? 1 : 0;
        }
        std::stable_sort(order_of.begin(), order_of.end(), [&](uint32_t a_, uint32_t b_) {
            if (large[a_] != large[b_]) return large[a_] > large[b_];
            if (large[a_]) return a_ < b_;
            return code[a_] < code[b_];
        });
    }
    return order_mode;
}

}  // namespace

extern "C" {

int pt_abi_version(void) { return PT_ABI_VERSION; }
const char* pt_last_error(void) { return g_last_error.c_str(); }

uint32_t pt_abi_struct_size(int which) {
    switch (which) {
        case 0: return sizeof(PtParams);
        case 1: return sizeof(PtCamera);
        case 2: return sizeof(PtTexture);
        case 3: return sizeof(PtMaterial);
        case 4: return sizeof(PtPerlin);
        case 5: return sizeof(PtSceneDesc);
        case 6: return sizeof(PtPartition);
        case 7: return sizeof(PtDeviceInfo);
        case 8: return sizeof(PtRenderStats);
        case 9: return sizeof(PtMotion);
        case 10: return sizeof(PtImage);
        default: return 0;
    }
}

uint32_t pt_partition_rows(const PtPartition* part_in, uint32_t height, uint32_t* rows_out, uint32_t cap) {
    PtPartition p;
    if (normalise_partition(part_in, &p) != PT_OK) return 0;
    uint32_t count = 0;
    const uint32_t n_tiles = (height + p.tile_rows - 1) / p.tile_rows;
    for (uint32_t k = p.part_index; k < n_tiles; k += p.part_count) {
        const uint32_t r0 = k * p.tile_rows;
        const uint32_t nr = std::min(p.tile_rows, height - r0);
        for (uint32_t r = 0; r < nr; ++r) {
            if (rows_out && count < cap) rows_out[count] = r0 + r;
            ++count;
        }
    }
    return count;
}

uint32_t pt_scene_storage_order(const PtSceneDesc* desc, uint32_t* order_out, uint32_t cap) {
    if (!desc || desc->struct_size != sizeof(PtSceneDesc)) return 0;
    const uint32_t n = desc->n_spheres;
    if (n > 0 && (!desc->centre_x || !desc->centre_y || !desc->centre_z || !desc->radius)) return 0;
    bool any_moving = false;
    if (desc->motion)
        for (uint32_t i = 0; i < n; ++i) any_moving = any_moving || desc->motion[i].moving != 0;
    std::vector<uint32_t> order_of;
    const int mode = storage_order(desc, any_moving, order_of);
    for (uint32_t j = 0; j < n && j < cap && order_out; ++j) order_out[j] = order_of[j];
    return (uint32_t)mode;
}

int pt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int pt_device_info(int device, PtDeviceInfo* out) {
    if (!out) return fail(PT_ERR_INVALID, "null out");
    cudaDeviceProp prop;
    int rc = check_device(device, &prop);
    if (rc != PT_OK) return rc;
    std::memset(out, 0, sizeof(*out));
    std::strncpy(out->name, prop.name, sizeof(out->name) - 1);
    out->sm_count = prop.multiProcessorCount;
    out->cc_major = prop.major;
    out->cc_minor = prop.minor;
    int khz = 0;
    PT_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
    out->sm_clock_khz = khz;
    out->fp32_fma_peak_flops = (double)prop.multiProcessorCount * 128.0 * 2.0 * (double)khz * 1e3;
    out->global_mem_bytes = prop.totalGlobalMem;
    return PT_OK;
}

int pt_scene_create(const PtSceneDesc* desc, int device, PtScene** out) {
    if (!desc || !out) return fail(PT_ERR_INVALID, "null desc/out");
    *out = nullptr;
    if (desc->struct_size != sizeof(PtSceneDesc)) return fail(PT_ERR_INVALID, "PtSceneDesc.struct_size %u != %zu (ABI mismatch)", desc->struct_size, sizeof(PtSceneDesc));
    const uint32_t n = desc->n_spheres;
    if (n > 0 && (!desc->centre_x || !desc->centre_y || !desc->centre_z || !desc->radius || !desc->material_index))
        return fail(PT_ERR_INVALID, "null sphere arrays with n_spheres = %u", n);
    if (n > 0 && (desc->n_materials == 0 || !desc->materials)) return fail(PT_ERR_INVALID, "spheres without materials");
    if (desc->n_textures > 0 && !desc->textures) return fail(PT_ERR_INVALID, "null textures");
    if (n > (1u << 28)) return fail(PT_ERR_TOO_LARGE, "too many spheres: %u", n);
    if (desc->n_images > 0 && !desc->images) return fail(PT_ERR_INVALID, "null images with n_images = %u", desc->n_images);

    // ---- images (texture.rs:6-37): one pool of packed RGB8 pixels, each image at a 16-byte aligned offset ----
    std::vector<size_t> image_offset(desc->n_images, 0);
    size_t image_pool_bytes = 0;
    for (uint32_t i = 0; i < desc->n_images; ++i) {
        const PtImage& im = desc->images[i];
        if (im.width == 0 || im.height == 0 || !im.data) return fail(PT_ERR_INVALID, "image %u: empty (%ux%u) or null data", i, im.width, im.height);
        if (im.width > (1u << 20) || im.height > (1u << 20)) return fail(PT_ERR_TOO_LARGE, "image %u: %ux%u is too large", i, im.width, im.height);
        image_offset[i] = image_pool_bytes;
        image_pool_bytes += ((size_t)im.width * im.height * 3 + 15) & ~(size_t)15;
        if (image_pool_bytes > 0x7fffffffULL) return fail(PT_ERR_TOO_LARGE, "image textures exceed 2 GB");
    }

    // ---- validate + flatten materials/textures ----
    bool uses_noise = false;
    for (uint32_t t = 0; t < desc->n_textures; ++t) {
        const PtTexture& tx = desc->textures[t];
        if (tx.kind == PT_TEX_CHECKER) {
            if (tx.odd < 0 || tx.even < 0 || (uint32_t)tx.odd >= desc->n_textures || (uint32_t)tx.even >= desc->n_textures)
                return fail(PT_ERR_INVALID, "texture %u: checker child index out of range", t);
        } else if (tx.kind == PT_TEX_NOISE) {
            uses_noise = true;
        } else if (tx.kind == PT_TEX_IMAGE) {
            if (tx.image < 0 || (uint32_t)tx.image >= desc->n_images) return fail(PT_ERR_INVALID, "texture %u: image index %d out of range", t, tx.image);
        } else if (tx.kind != PT_TEX_CONSTANT) {
            return fail(PT_ERR_UNSUPPORTED, "texture %u: kind %d is not a Texture variant (texture.rs:40-55)", t, tx.kind);
        }
    }
    if (uses_noise && !desc->perlin) return fail(PT_ERR_INVALID, "a Noise texture is present but desc.perlin is NULL");
    for (uint32_t m = 0; m < desc->n_materials; ++m) {
        const PtMaterial& mt = desc->materials[m];
        if (mt.kind < PT_MAT_LAMBERTIAN || mt.kind > PT_MAT_DIFFUSE_LIGHT)
            return fail(PT_ERR_UNSUPPORTED, "material %u: kind %d is not supported (Isotropic needs ConstantMedium, out of scope)", m, mt.kind);
        if ((mt.kind == PT_MAT_LAMBERTIAN || mt.kind == PT_MAT_DIFFUSE_LIGHT) && (mt.texture < 0 || (uint32_t)mt.texture >= desc->n_textures))
            return fail(PT_ERR_INVALID, "material %u: texture index %d out of range", m, mt.texture);
    }

    bool any_moving = false;
    float t_lo = -FLT_MAX, t_hi = FLT_MAX;
    if (desc->motion) {
        for (uint32_t i = 0; i < n; ++i) {
            const PtMotion& mo = desc->motion[i];
            if (!mo.moving) continue;
            if (!(mo.time1 > mo.time0)) return fail(PT_ERR_INVALID, "sphere %u: MovingSphere needs time1 > time0 (got %g, %g)", i, mo.time0, mo.time1);
            any_moving = true;
            t_lo = std::max(t_lo, mo.time0);
            t_hi = std::min(t_hi, mo.time1);
        }
    }
    auto is_moving = [&](uint32_t i) { return any_moving && desc->motion[i].moving != 0; };

    std::vector<uint32_t> order_of;
    const bool spatial = storage_order(desc, any_moving, order_of) != 0;

    cudaDeviceProp prop;
    int rc = check_device(device, &prop);
    if (rc != PT_OK) return rc;
    PT_CUDA(cudaSetDevice(device));

    PtScene* s = new PtScene();
    s->device = device;
    s->sm_count = prop.multiProcessorCount;
    s->n_spheres = n;
    s->n_blocks = (int)((n + 3) / 4);
    s->n_blocks = (s->n_blocks + pt::kLdsGroupBlocks - 1) / pt::kLdsGroupBlocks * pt::kLdsGroupBlocks;  // whole groups; padding spheres can never be hit
    if (s->n_blocks > pt::kMaxSweepBlocks) { delete s; return fail(PT_ERR_TOO_LARGE, "too many spheres for the candidate-queue encoding: %u", n); }
    s->has_noise = uses_noise;
    s->has_sky = desc->has_sky != 0;
    s->sky = pt::V3{desc->sky[0], desc->sky[1], desc->sky[2]};

    // ---- sphere blocks: X,Y,Z,R^2 for 4 spheres; padding = (FLT_MAX centre, r^2 = 0) spheres_soa.rs:53-61 ----
    std::vector<float4> blocks((size_t)std::max(s->n_blocks, 1) * 4);
    std::vector<pt::DevShade> shade(std::max<uint32_t>(n, 1));
    for (int j = 0; j < s->n_blocks; ++j) {
        float* f = reinterpret_cast<float*>(&blocks[(size_t)j * 4]);
        for (int e = 0; e < 4; ++e) {
            const bool valid = (uint32_t)j * 4 + e < n;
            const uint32_t i = valid ? order_of[(uint32_t)j * 4 + e] : 0u;  // position in the caller's list
            f[0 + e] = valid ? desc->centre_x[i] : FLT_MAX;
            f[4 + e] = valid ? desc->centre_y[i] : FLT_MAX;
            f[8 + e] = valid ? desc->centre_z[i] : FLT_MAX;
            f[12 + e] = valid ? desc->radius[i] * desc->radius[i] : 0.0f;  // spheres_soa.rs:46
            if (valid && is_moving(i)) f[12 + e] = -std::max(f[12 + e], FLT_MIN);  // negative r^2 tags a MovingSphere (pt_sweep.cuh)
        }
    }
    // pre-filter image (pt_sweep.cuh): X, Y, Z, K = r^2 - |c|^2 + 2^-19 (|c|^2 + r^2), padded to the group
    std::vector<float4> h_prefilter;
    {
        h_prefilter.assign((size_t)std::max(s->n_blocks, 1) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
        for (int j = 0; j < s->n_blocks; ++j) {
            float* f = reinterpret_cast<float*>(&h_prefilter[(size_t)j * 4]);
            for (int e = 0; e < 4; ++e) {
                f[0 + e] = f[4 + e] = f[8 + e] = 0.0f;
                f[12 + e] = -3.0e38f;  // padding: never a candidate
                if ((uint32_t)j * 4 + e >= n) continue;
                const uint32_t i = order_of[(uint32_t)j * 4 + e];
                double cx = desc->centre_x[i], cy = desc->centre_y[i], cz = desc->centre_z[i], r = std::fabs((double)desc->radius[i]);
                if (is_moving(i)) {  // static bound of the whole sweep: centre0 + delta/2, radius r + |delta|/2 (+ f32 rounding of the lerp)
                    const double ex = desc->motion[i].centre1[0] - cx, ey = desc->motion[i].centre1[1] - cy, ez = desc->motion[i].centre1[2] - cz;
                    cx += 0.5 * ex; cy += 0.5 * ey; cz += 0.5 * ez;
                    r += 0.5 * std::sqrt(ex * ex + ey * ey + ez * ez) * (1.0 + 1e-5) + 1e-6 * (std::fabs(cx) + std::fabs(cy) + std::fabs(cz) + r);
                }
                const double c2 = cx * cx + cy * cy + cz * cz, r2 = r * r;
                if (!(c2 + r2 < 1.0e24)) {  // too large (or NaN) for the expanded form: always a candidate, the exact test decides
                    f[12 + e] = INFINITY;
                    continue;
                }
                f[0 + e] = (float)cx;
                f[4 + e] = (float)cy;
                f[8 + e] = (float)cz;
                const double k = r2 - c2 + 1.9073486328125e-06 * (c2 + r2);
                f[12 + e] = std::nextafter((float)k, INFINITY);  // round towards "candidate"
            }
        }
    }
    for (uint32_t j = 0; j < n; ++j) {
        const uint32_t i = order_of[j];
        const int32_t mi = desc->material_index[i];
        if (mi < 0 || (uint32_t)mi >= desc->n_materials) {
            delete s;
            return fail(PT_ERR_INVALID, "sphere %u: material index %d out of range", i, mi);
        }
        const PtMaterial& mt = desc->materials[mi];
        pt::DevShade d{};
        d.rinv = 1.0f / desc->radius[i];  // spheres_soa.rs:47
        d.kind = mt.kind;
        d.tex = -1;
        d.moving = is_moving(i) ? 1 : 0;
        if (mt.kind == PT_MAT_LAMBERTIAN || mt.kind == PT_MAT_DIFFUSE_LIGHT) {
            const PtTexture& tx = desc->textures[mt.texture];
            if (tx.kind == PT_TEX_CONSTANT) {  // fold the constant colour into the per-sphere record
                d.ar = tx.color[0];
                d.ag = tx.color[1];
                d.ab = tx.color[2];
            } else {
                d.tex = mt.texture;
            }
        } else if (mt.kind == PT_MAT_METAL) {
            d.ar = mt.albedo[0];
            d.ag = mt.albedo[1];
            d.ab = mt.albedo[2];
            d.param = mt.fuzz;
        } else {
            d.param = mt.ref_idx;
        }
        shade[j] = d;
    }
    std::vector<pt::DevTexture> tex(std::max<uint32_t>(desc->n_textures, 1));
    for (uint32_t t = 0; t < desc->n_textures; ++t) {
        const PtTexture& tx = desc->textures[t];
        pt::DevTexture d{};
        d.r = tx.color[0];
        d.g = tx.color[1];
        d.b = tx.color[2];
        d.scale = tx.scale;
        d.kind = tx.kind;
        d.odd = tx.odd;
        d.even = tx.even;
        if (tx.kind == PT_TEX_IMAGE) {
            d.odd = (int32_t)desc->images[tx.image].width;
            d.even = (int32_t)desc->images[tx.image].height;
            d.offset = (int32_t)image_offset[tx.image];
        }
        tex[t] = d;
    }

    auto cleanup_fail = [&](int code) {
        pt_scene_destroy(s);
        return code;
    };
#define PT_CUDA_S(call)                                                                                     \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) return cleanup_fail(fail(PT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_))); \
    } while (0)
    PT_CUDA_S(cudaMalloc(&s->d_blocks, blocks.size() * sizeof(float4)));
    PT_CUDA_S(cudaMemcpy(s->d_blocks, blocks.data(), blocks.size() * sizeof(float4), cudaMemcpyHostToDevice));
    PT_CUDA_S(cudaMalloc(&s->d_shade, shade.size() * sizeof(pt::DevShade)));
    PT_CUDA_S(cudaMemcpy(s->d_shade, shade.data(), shade.size() * sizeof(pt::DevShade), cudaMemcpyHostToDevice));
    PT_CUDA_S(cudaMalloc(&s->d_tex, tex.size() * sizeof(pt::DevTexture)));
    PT_CUDA_S(cudaMemcpy(s->d_tex, tex.data(), tex.size() * sizeof(pt::DevTexture), cudaMemcpyHostToDevice));
    if (spatial) {
        PT_CUDA_S(cudaMalloc(&s->d_order, order_of.size() * sizeof(uint32_t)));
        PT_CUDA_S(cudaMemcpy(s->d_order, order_of.data(), order_of.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    if (image_pool_bytes > 0) {
        std::vector<uint8_t> pool(image_pool_bytes, 0);
        for (uint32_t i = 0; i < desc->n_images; ++i)
            std::memcpy(pool.data() + image_offset[i], desc->images[i].data, (size_t)desc->images[i].width * desc->images[i].height * 3);
        PT_CUDA_S(cudaMalloc(&s->d_images, pool.size()));
        PT_CUDA_S(cudaMemcpy(s->d_images, pool.data(), pool.size(), cudaMemcpyHostToDevice));
    }
    if (any_moving) {
        std::vector<pt::DevMotion> motion(n);
        for (uint32_t j = 0; j < n; ++j) {
            const uint32_t i = order_of[j];
            pt::DevMotion m{};
            if (is_moving(i)) {  // MovingSphere::new, moving_sphere.rs:16-26
                const PtMotion& mo = desc->motion[i];
                m.dx = mo.centre1[0] - desc->centre_x[i];
                m.dy = mo.centre1[1] - desc->centre_y[i];
                m.dz = mo.centre1[2] - desc->centre_z[i];
                m.time_start = mo.time0;
                m.inv_time_delta = 1.0f / (mo.time1 - mo.time0);
                m.radius = desc->radius[i];
            }
            motion[j] = m;
        }
        PT_CUDA_S(cudaMalloc(&s->d_motion, motion.size() * sizeof(pt::DevMotion)));
        PT_CUDA_S(cudaMemcpy(s->d_motion, motion.data(), motion.size() * sizeof(pt::DevMotion), cudaMemcpyHostToDevice));
        s->motion_t_lo = t_lo;
        s->motion_t_hi = t_hi;
    }
    PT_CUDA_S(cudaMalloc(&s->d_prefilter, h_prefilter.size() * sizeof(float4)));
    PT_CUDA_S(cudaMemcpy(s->d_prefilter, h_prefilter.data(), h_prefilter.size() * sizeof(float4), cudaMemcpyHostToDevice));
    PT_CUDA_S(cudaMalloc(&s->d_perlin, sizeof(pt::PerlinSmem)));
    {
        std::vector<unsigned char> raw(sizeof(pt::PerlinSmem), 0);
        pt::PerlinSmem* ps = reinterpret_cast<pt::PerlinSmem*>(raw.data());
        if (desc->perlin) {
            for (int i = 0; i < 256; ++i) {
                ps->randvec[i] = make_float4(desc->perlin->randvec[i][0], desc->perlin->randvec[i][1], desc->perlin->randvec[i][2], 0.0f);
                if (desc->perlin->perm_x[i] > 255 || desc->perlin->perm_y[i] > 255 || desc->perlin->perm_z[i] > 255)
                    return cleanup_fail(fail(PT_ERR_INVALID, "perlin permutation entry %d out of range", i));
                ps->perm_x[i] = (uint8_t)desc->perlin->perm_x[i];
                ps->perm_y[i] = (uint8_t)desc->perlin->perm_y[i];
                ps->perm_z[i] = (uint8_t)desc->perlin->perm_z[i];
            }
        }
        PT_CUDA_S(cudaMemcpy(s->d_perlin, raw.data(), raw.size(), cudaMemcpyHostToDevice));
    }
    PT_CUDA_S(cudaMalloc(&s->d_ray_count, sizeof(unsigned long long)));
    PT_CUDA_S(cudaMalloc(&s->d_sweep_count, sizeof(unsigned long long)));
    PT_CUDA_S(cudaMalloc(&s->d_next_pixel, sizeof(unsigned int)));
    PT_CUDA_S(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (auto& e : s->ev) PT_CUDA_S(cudaEventCreate(&e));
#undef PT_CUDA_S
    rc = plan_launch(s);
    if (rc != PT_OK) return cleanup_fail(rc);
    *out = s;
    return PT_OK;
}

void pt_scene_destroy(PtScene* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    cudaFree(s->d_blocks);
    cudaFree(s->d_shade);
    cudaFree(s->d_tex);
    cudaFree(s->d_images);
    cudaFree(s->d_order);
    cudaFree(s->d_perlin);
    cudaFree(s->d_prefilter);
    cudaFree(s->d_motion);
    cudaFree(s->d_ray_count);
    cudaFree(s->d_sweep_count);
    cudaFree(s->d_next_pixel);
    cudaFree(s->d_pixstate);
    cudaFree(s->d_rgb);
    cudaFree(s->d_rgb8);
    for (auto& e : s->ev)
        if (e) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int pt_render_part(PtScene* s, const PtParams* params, const PtCamera* camera, uint32_t frame_num, const PtPartition* part_in,
                   float* rgb_inout, uint64_t* ray_count_out) {
    if (!s || !rgb_inout) return fail(PT_ERR_INVALID, "null scene/buffer");
    int rc = validate_params(params, camera);
    if (rc != PT_OK) return rc;
    PtPartition part;
    rc = normalise_partition(part_in, &part);
    if (rc != PT_OK) return rc;
    PT_CUDA(cudaSetDevice(s->device));
    const size_t floats = (size_t)params->width * params->height * 3;
    rc = ensure_image(s, floats);
    if (rc != PT_OK) return rc;

    s->stats = PtRenderStats{};
    s->prog_valid = false;  // this call reuses the scene's device image
    uint64_t h2d = 0, d2h = 0;
    PT_CUDA(cudaEventRecord(s->ev[0], s->stream));
    if (frame_num != 0) {  // the blend reads the previous frame (scene.rs:114-116); frame 0 has mix_prev = 0
        rc = copy_owned_rows(s, params, part, rgb_inout, true, &h2d);
        if (rc != PT_OK) return rc;
    }
    PT_CUDA(cudaEventRecord(s->ev[1], s->stream));
    rc = launch_update(s, params, camera, frame_num, part, s->d_rgb, s->d_ray_count, s->stream);
    if (rc != PT_OK) return rc;
    PT_CUDA(cudaEventRecord(s->ev[2], s->stream));
    rc = copy_owned_rows(s, params, part, rgb_inout, false, &d2h);
    if (rc != PT_OK) return rc;
    unsigned long long rays = 0;
    PT_CUDA(cudaMemcpyAsync(&rays, s->d_ray_count, sizeof(rays), cudaMemcpyDeviceToHost, s->stream));
    PT_CUDA(cudaEventRecord(s->ev[3], s->stream));
    PT_CUDA(cudaStreamSynchronize(s->stream));
    float ms = 0;
    PT_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
    s->stats.h2d_ms = ms;
    PT_CUDA(cudaEventElapsedTime(&ms, s->ev[1], s->ev[2]));
    s->stats.kernel_ms = ms;
    PT_CUDA(cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]));
    s->stats.d2h_ms = ms;
    s->stats.h2d_bytes = h2d;
    s->stats.d2h_bytes = d2h + sizeof(rays);
    s->stats.ray_count = rays;
    if (ray_count_out) *ray_count_out = rays;
    if (std::getenv("PTGPU_DEBUG_SWEEPS")) {  // diagnostic: lane efficiency of the sweep = rays / (32 x warp sweeps)
        unsigned long long sweeps = 0;
        cudaMemcpy(&sweeps, s->d_sweep_count, sizeof(sweeps), cudaMemcpyDeviceToHost);
        std::fprintf(stderr, "[ptgpu] warp sweeps %llu, rays %llu, lane efficiency %.3f\n", sweeps, rays, sweeps ? (double)rays / (32.0 * (double)sweeps) : 0.0);
    }
    return PT_OK;
}

int pt_render(PtScene* s, const PtParams* params, const PtCamera* camera, uint32_t frame_num, float* rgb_inout, uint64_t* ray_count_out) {
    return pt_render_part(s, params, camera, frame_num, nullptr, rgb_inout, ray_count_out);
}

int pt_render_progressive(PtScene* s, const PtParams* params, const PtCamera* camera, uint32_t frame_num, float* rgb_out, uint8_t* rgb8_out,
                          uint64_t* ray_count_out) {
    if (!s) return fail(PT_ERR_INVALID, "null scene");
    int rc = validate_params(params, camera);
    if (rc != PT_OK) return rc;
    if (frame_num != 0 && !(s->prog_valid && s->prog_w == params->width && s->prog_h == params->height && s->prog_next_frame == frame_num))
        return fail(PT_ERR_INVALID, "frame %u does not continue the resident accumulation (have %ux%u, next frame %u)", frame_num, s->prog_w,
                    s->prog_h, s->prog_valid ? s->prog_next_frame : 0u);
    PT_CUDA(cudaSetDevice(s->device));
    const size_t n = (size_t)params->width * params->height;
    s->prog_valid = false;  // any failure below leaves the resident image undefined
    rc = ensure_image(s, n * 3);
    if (rc != PT_OK) return rc;
    const PtPartition whole{4, 0, 1, 0};
    s->stats = PtRenderStats{};
    PT_CUDA(cudaEventRecord(s->ev[1], s->stream));
    rc = launch_update(s, params, camera, frame_num, whole, s->d_rgb, s->d_ray_count, s->stream);
    if (rc != PT_OK) return rc;
    PT_CUDA(cudaEventRecord(s->ev[2], s->stream));
    uint64_t d2h = 0;
    if (rgb_out) {
        PT_CUDA(cudaMemcpyAsync(rgb_out, s->d_rgb, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
        d2h += n * 3 * sizeof(float);
    }
    if (rgb8_out) {
        if (s->d_rgb8_bytes < n * 3) {
            if (s->d_rgb8) cudaFree(s->d_rgb8);
            s->d_rgb8 = nullptr;
            s->d_rgb8_bytes = 0;
            PT_CUDA(cudaMalloc(&s->d_rgb8, n * 3));
            s->d_rgb8_bytes = n * 3;
        }
        rc = pt_srgb8_device(s, s->d_rgb, params->width, params->height, s->d_rgb8, s->stream);
        if (rc != PT_OK) return rc;
        PT_CUDA(cudaMemcpyAsync(rgb8_out, s->d_rgb8, n * 3, cudaMemcpyDeviceToHost, s->stream));
        d2h += n * 3;
    }
    unsigned long long rays = 0;
    PT_CUDA(cudaMemcpyAsync(&rays, s->d_ray_count, sizeof(rays), cudaMemcpyDeviceToHost, s->stream));
    PT_CUDA(cudaEventRecord(s->ev[3], s->stream));
    PT_CUDA(cudaStreamSynchronize(s->stream));
    float ms = 0;
    PT_CUDA(cudaEventElapsedTime(&ms, s->ev[1], s->ev[2]));
    s->stats.kernel_ms = ms;
    PT_CUDA(cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]));
    s->stats.d2h_ms = ms;
    s->stats.d2h_bytes = d2h + sizeof(rays);
    s->stats.ray_count = rays;
    if (ray_count_out) *ray_count_out = rays;
    s->prog_valid = true;
    s->prog_w = params->width;
    s->prog_h = params->height;
    s->prog_next_frame = frame_num + 1;
    return PT_OK;
}

int pt_render_device(PtScene* s, const PtParams* params, const PtCamera* camera, uint32_t frame_num, const PtPartition* part_in,
                     float* d_rgb_inout, uint64_t* d_ray_count, void* cuda_stream) {
    if (!s || !d_rgb_inout || !d_ray_count) return fail(PT_ERR_INVALID, "null scene/buffer");
    int rc = validate_params(params, camera);
    if (rc != PT_OK) return rc;
    PtPartition part;
    rc = normalise_partition(part_in, &part);
    if (rc != PT_OK) return rc;
    PT_CUDA(cudaSetDevice(s->device));
    s->stats = PtRenderStats{};
    return launch_update(s, params, camera, frame_num, part, d_rgb_inout, reinterpret_cast<unsigned long long*>(d_ray_count),
                         reinterpret_cast<cudaStream_t>(cuda_stream));
}

int pt_srgb8_device(PtScene* s, const float* d_rgb, uint32_t width, uint32_t height, uint8_t* d_rgb8_out, void* cuda_stream) {
    if (!s || !d_rgb || !d_rgb8_out || width == 0 || height == 0) return fail(PT_ERR_INVALID, "null/empty argument");
    PT_CUDA(cudaSetDevice(s->device));
    const size_t n = (size_t)width * height;
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((n + threads - 1) / threads, (size_t)s->sm_count * 8);
    pt::pt_srgb8_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(d_rgb, width, height, d_rgb8_out);
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_srgb8(PtScene* s, const float* rgb, uint32_t width, uint32_t height, uint8_t* rgb8_out) {
    if (!s || !rgb || !rgb8_out || width == 0 || height == 0) return fail(PT_ERR_INVALID, "null/empty argument");
    PT_CUDA(cudaSetDevice(s->device));
    const size_t n = (size_t)width * height;
    int rc = ensure_image(s, n * 3);
    if (rc != PT_OK) return rc;
    if (s->d_rgb8_bytes < n * 3) {
        if (s->d_rgb8) cudaFree(s->d_rgb8);
        s->d_rgb8 = nullptr;
        s->d_rgb8_bytes = 0;
        PT_CUDA(cudaMalloc(&s->d_rgb8, n * 3));
        s->d_rgb8_bytes = n * 3;
    }
    s->prog_valid = false;  // this call reuses the scene's device image
    PT_CUDA(cudaMemcpyAsync(s->d_rgb, rgb, n * 3 * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    rc = pt_srgb8_device(s, s->d_rgb, width, height, s->d_rgb8, s->stream);
    if (rc != PT_OK) return rc;
    PT_CUDA(cudaMemcpyAsync(rgb8_out, s->d_rgb8, n * 3, cudaMemcpyDeviceToHost, s->stream));
    PT_CUDA(cudaStreamSynchronize(s->stream));
    return PT_OK;
}

int pt_scene_stats(const PtScene* s, PtRenderStats* out) {
    if (!s || !out) return fail(PT_ERR_INVALID, "null argument");
    *out = s->stats;
    return PT_OK;
}

#ifdef PT_PROFILE
// profile build only (tools/phase_profile.py): read and clear the phase counters
int pt_profile_read(unsigned long long* out8) {
    PT_CUDA(cudaMemcpyFromSymbol(out8, pt::g_prof, sizeof(unsigned long long) * 8));
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    PT_CUDA(cudaMemcpyToSymbol(pt::g_prof, z, sizeof(z)));
    return PT_OK;
}
#endif

int pt_probe_fp32_peak(int device, double* flops_out) {
    if (!flops_out) return fail(PT_ERR_INVALID, "null out");
    cudaDeviceProp prop;
    int rc = check_device(device, &prop);
    if (rc != PT_OK) return rc;
    PT_CUDA(cudaSetDevice(device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 20000;
    float* d_out = nullptr;
    PT_CUDA(cudaMalloc(&d_out, sizeof(float) * threads * blocks));
    cudaEvent_t e0, e1;
    PT_CUDA(cudaEventCreate(&e0));
    PT_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        PT_CUDA(cudaEventRecord(e0));
        pt::pt_ffma_peak_kernel<<<blocks, threads>>>(d_out, iters, 1.0001f, 0.5f);
        PT_CUDA(cudaEventRecord(e1));
        PT_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        PT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *flops_out = 2.0 * 8.0 * iters * (double)threads * blocks / (best * 1e-3);
    return PT_OK;
}

}  // extern "C"
