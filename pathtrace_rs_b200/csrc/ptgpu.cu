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
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ptgpu.h"
#include "pt_regroup.cuh"
#include "pt_wave.cuh"

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
#ifndef PT_MMA_RESIDENT_MIN_CTAS
#define PT_MMA_RESIDENT_MIN_CTAS 2  // the tensor-path regroup kernel is chosen while at least this many CTAs fit on an SM
#endif

}  // namespace

// Host image of a flattened scene, built once per pt_scene_create* and uploaded to every device of the scene.
struct FlatScene {
    uint32_t n_spheres = 0;
    int n_blocks = 0;
    bool has_noise = false, has_sky = false, spatial = false, any_moving = false;
    pt::V3 sky{0, 0, 0};
    float motion_t_lo = 0.0f, motion_t_hi = 0.0f;
    std::vector<float4> blocks, prefilter, kplane;
    std::vector<pt::DevShade> shade;
    std::vector<pt::DevTexture> tex;
    std::vector<uint32_t> order_of;
    std::vector<uint8_t> image_pool;
    std::vector<pt::DevMotion> motion;
    std::vector<unsigned char> perlin_raw;
    std::unique_ptr<pt::ConstImageT<true>> const_image;  // X,Y,Z planes for the kernel-parameter image (n_blocks <= kMaxConstBlocks)
    // tensor-path pre-filter (pt_sweep_mma.cuh): fragment-ordered f16 sphere operand + the scene's scales; mma_ok = the
    // scene is one the tensor path is worth using on (enough spheres, absolute slack small against the spheres' size)
    std::vector<uint4> mma_image;
    std::vector<uint16_t> mma_rows;  // per stored sphere (incl. padding) its 16 K halves, before the fragment shuffle (pt_scene_mma_operand)
    pt::MmaScale mma{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bool mma_ok = false;
};

// One device's copy of a scene plus its per-render scratch.
struct Replica {
    int device = 0;
    int sm_count = 0;
    uint32_t n_spheres = 0;
    int n_blocks = 0;
    bool has_noise = false;
    bool has_sky = false;
    pt::V3 sky{0, 0, 0};
    PtOptions opt{};
    float4* d_blocks = nullptr;
    pt::DevShade* d_shade = nullptr;
    pt::DevTexture* d_tex = nullptr;
    uint8_t* d_images = nullptr;  // RGB8 pool of the Image textures (nullptr: none)
    uint32_t* d_order = nullptr;  // stored sphere index -> position in the caller's list (nullptr: stored in list order)
    pt::PerlinSmem* d_perlin = nullptr;
    pt::DevMotion* d_motion = nullptr;  // MovingSphere records (nullptr: none)
    float motion_t_lo = 0.0f, motion_t_hi = 0.0f;  // intersection of the moving spheres' [time0, time1]
    float4* d_prefilter = nullptr;  // pre-filter image X,Y,Z,K per block (LDS kernels stage/stream it)
    float4* d_kplane = nullptr;     // its K plane alone (resident kernel with the X,Y,Z planes in the parameter image)
    bool const_image = false;       // resident kernel reads X,Y,Z through the kernel-parameter image
    bool wave = false;              // ... in its wavefront form (pt_wave.cuh): one CTA per SM, wave_pool paths in shared memory
    bool regroup = false;           // resident scene rendered by the one-path-per-lane kernel with the CTA regroup (pt_regroup.cuh): the default
    bool mma = false;               // ... with stage 1 of its sweep on the tensor path (pt_sweep_mma.cuh)
    bool mma_ok = false;            // the scene has a usable tensor-path image
    bool exact_smem = false;        // tensor-path regroup kernel with the exact blocks in shared memory as well
    uint4* d_mma_image = nullptr;
    pt::MmaScale mma_scale{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    size_t fp32_smem_bytes = 0;     // the FP32 regroup kernel's launch geometry, kept as the per-render fallback of the tensor path
    int fp32_ctas_per_sm = 0;       // (camera outside the extent the f16 operands were scaled for)
    int fp32_tile_blocks = 0, fp32_n_tiles = 0;
    size_t mma_smem_bytes_exact = 0;  // shared memory of the exact_smem flavour
    int wave_pool = 0;
    const pt::ConstImageT<true>* h_const_image = nullptr;  // owned by the PtScene; passed by value at every launch (24 KB)
    // per-render scratch.  At most one render is in flight per replica: every launch waits for the previous one's
    // event (ev_busy), whatever stream it was issued on, so the scratch below is never shared by two launches.
    unsigned long long* d_ray_count = nullptr;  // [0] ray count
    unsigned long long* d_sweep_count = nullptr;  // warp-level sweeps of the last launch (lane-efficiency diagnostic)
    unsigned int* d_next_pixel = nullptr;
    unsigned int* d_status = nullptr;  // wavefront kernel's watchdog word
    uint32_t* d_pixstate = nullptr;  // chunk queue: 12 words per owned pixel (pt_megakernel.cuh, PixState)
    size_t d_pixstate_pixels = 0;
    float* d_rgb = nullptr;  // device image for the host-buffer entry points
    size_t d_rgb_floats = 0;
    uint8_t* d_rgb8 = nullptr;
    size_t d_rgb8_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_busy = nullptr;
    bool busy_recorded = false;
    // launch geometry
    bool resident = true;
    int tile_blocks = 0, n_tiles = 0;
    size_t smem_bytes = 0;
    int ctas_per_sm = 0;
    PtRenderStats stats{};
};

struct PtScene {
    std::vector<Replica*> reps;  // one per device, in the caller's device order
    PtOptions opt{};
    FlatScene flat;              // kept for the parameter image (and cheap: the scene is small)
    PtRenderStats stats{};       // aggregate of the last render
    // pt_render_progressive: what the resident images currently hold
    bool prog_valid = false;
    uint32_t prog_w = 0, prog_h = 0, prog_next_frame = 0;
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
    if (ctas_per_sm) *ctas_per_sm = occ;
    return PT_OK;
}

// shared-memory budget of the kernels (one place: plan_launch and the storage-order decision both ask)
constexpr size_t kQueueBytes = (size_t)pt::kQueueCap * pt::kCtaThreads * sizeof(uint32_t);  // per-lane candidate queues
constexpr size_t kStreamedFixedBytes = sizeof(pt::PerlinSmem) + 2 * pt::kCtaThreads * sizeof(uint32_t) + kQueueBytes;  // streamed kernel: Perlin tables + `pend` words + ray.time slots + queues
constexpr size_t kPathBytes = (size_t)pt::kPathWords * pt::kPathRows * pt::kCtaThreads * sizeof(uint32_t);            // resident kernel: path records
size_t resident_smem(int n_blocks, bool const_image) {
    const size_t image = ((size_t)n_blocks * (const_image ? 16 : 64) + 127) & ~(size_t)127;
    return image + sizeof(pt::PerlinSmem) + kQueueBytes + kPathBytes;
}
constexpr size_t kRegroupBytes = (size_t)pt::kRegroupWords * pt::kCtaThreads * sizeof(uint32_t) + 64 * PT_REGROUP_DOMAINS;  // regroup kernel: path-state exchange + category counters
size_t regroup_smem(int n_blocks) { return (size_t)n_blocks * 64 + kStreamedFixedBytes + kRegroupBytes; }
size_t regroup_mma_smem(int n_blocks) { return (size_t)(n_blocks / pt::kLdsGroupBlocks + 1) * 512 + kStreamedFixedBytes + kRegroupBytes; }  // 32 B per sphere + one step of padding
// which resident kernel renders the scene: 0 automatic = the tensor-path regroup kernel where the scene suits it (5), else the FP32 one (4)
uint32_t resident_choice(const PtOptions& opt, bool mma_ok, int n_blocks) {
    uint32_t k = opt.resident_kernel;
    if (k == 0) k = 5;
    if (k == 5 && !(mma_ok && regroup_mma_smem(n_blocks) <= kMaxDynSmem / PT_MMA_RESIDENT_MIN_CTAS)) k = 4;  // two CTAs per SM at least
    return k;
}
bool fits_resident(int n_blocks, const PtOptions& opt) {
    if (opt.force_stream_tile_blocks > 0 && n_blocks > 0) return false;
    if (opt.resident_kernel == 0 || opt.resident_kernel == 4 || opt.resident_kernel == 5) return regroup_smem(n_blocks) <= kMaxDynSmem;
    return resident_smem(n_blocks, n_blocks <= pt::kMaxConstBlocks && opt.resident_kernel != 3) <= kMaxDynSmem;
}

int plan_launch(Replica* s) {
    int forced_tile = s->opt.force_stream_tile_blocks;
    // Mid-size scenes (about 2 000 - 13 000 spheres): the FP32 image still fits in shared memory, the tensor-path image no longer
    // does at two CTAs per SM.  Measured on B200 (tools/midsize_bench.py, profiles/midsize_r2.txt): the L2-streamed kernel with
    // the tensor-path pre-filter renders them 1.5-1.75x faster than the resident kernel with the FP32 pre-filter (and faster
    // than the tensor-path resident kernel at one CTA per SM), so the automatic choice streams them.
    const bool stream_for_mma = s->opt.resident_kernel == 0 && s->mma_ok && s->n_blocks / pt::kLdsGroupBlocks < 65536 &&
                                regroup_mma_smem(s->n_blocks) > kMaxDynSmem / PT_MMA_RESIDENT_MIN_CTAS;
    if (fits_resident(s->n_blocks, s->opt) && !stream_for_mma) {
        s->resident = true;
        s->const_image = s->n_blocks <= pt::kMaxConstBlocks && s->opt.resident_kernel != 3;
        s->smem_bytes = resident_smem(s->n_blocks, s->const_image);
        s->tile_blocks = s->n_blocks;
        s->n_tiles = 1;
        int rc;
        const uint32_t choice = resident_choice(s->opt, s->mma_ok, s->n_blocks);
        if (choice == 4 || choice == 5) {  // one path per lane + CTA regroup (pt_regroup.cuh), the fastest measured
            s->regroup = true;
            s->const_image = false;
            s->smem_bytes = regroup_smem(s->n_blocks);
            rc = s->d_motion ? configure_kernel(pt::pt_megakernel_regroup<true, false>, s->smem_bytes, &s->ctas_per_sm)
                             : configure_kernel(pt::pt_megakernel_regroup<false, false>, s->smem_bytes, &s->ctas_per_sm);
            if (rc == PT_OK)
                rc = s->d_motion ? configure_kernel(pt::pt_debug_hits_regroup<true, false>, s->smem_bytes, nullptr)
                                 : configure_kernel(pt::pt_debug_hits_regroup<false, false>, s->smem_bytes, nullptr);
            s->fp32_smem_bytes = s->smem_bytes;
            s->fp32_ctas_per_sm = s->ctas_per_sm;
            if (rc == PT_OK && choice == 5) {  // stage 1 of the sweep on the tensor path; the FP32 kernel above stays configured as its fallback
                s->mma = true;
                s->smem_bytes = regroup_mma_smem(s->n_blocks);
                rc = s->d_motion ? configure_kernel(pt::pt_megakernel_regroup<true, true>, s->smem_bytes, &s->ctas_per_sm)
                                 : configure_kernel(pt::pt_megakernel_regroup<false, true>, s->smem_bytes, &s->ctas_per_sm);
                if (rc == PT_OK)
                    rc = s->d_motion ? configure_kernel(pt::pt_debug_hits_regroup<true, true>, s->smem_bytes, nullptr)
                                     : configure_kernel(pt::pt_debug_hits_regroup<false, true>, s->smem_bytes, nullptr);
                // the exact blocks (16 B per sphere) in shared memory too, when that does not cost a resident CTA (measured: cfg2 +1.8 %)
                const size_t with_exact = s->smem_bytes + (size_t)s->n_blocks * 64 + (((size_t)s->n_blocks * 4 + 15) & ~(size_t)15);  // + one category byte per sphere
                if (rc == PT_OK && with_exact <= kMaxDynSmem) {
                    int ctas = 0;
                    rc = s->d_motion ? configure_kernel(pt::pt_megakernel_regroup<true, true, true>, with_exact, &ctas)
                                     : configure_kernel(pt::pt_megakernel_regroup<false, true, true>, with_exact, &ctas);
                    if (rc == PT_OK && ctas >= s->ctas_per_sm) {
                        s->exact_smem = true;
                        s->mma_smem_bytes_exact = with_exact;
                    }
                }
            }
            return rc;
        }
        if (s->const_image && s->opt.resident_kernel == 1) {
            // wavefront kernel: everything that is left of the SM's shared memory becomes the path pool
            const size_t fixed = (((size_t)s->n_blocks * (pt::kWaveConstImg ? 16 : 64) + 127) & ~(size_t)127) + sizeof(pt::PerlinSmem) + (size_t)pt::kWaveCandCap * pt::kWaveThreads * sizeof(uint32_t) + 64;
            const size_t per_path = (size_t)pt::kWaveRecWords * 4 + pt::kWaveQueues * sizeof(uint16_t);
            long pool = ((long)kMaxDynSmem - (long)fixed - (long)(pt::kWaveQueues * 64 * sizeof(uint16_t))) / (long)per_path;
            pool = std::min<long>(pool / 64 * 64, pt::kWaveMaxPool / 64 * 64);
            if (pool >= 2 * 64 * pt::kWaveWarps / 2) {  // at least 64 paths per warp
                s->wave = true;
                s->wave_pool = (int)pool;
                s->smem_bytes = fixed + (size_t)pt::kWaveQueues * ((size_t)pool + 64) * sizeof(uint16_t) + (size_t)pool * pt::kWaveRecWords * 4;
                PT_CUDA(cudaFuncSetAttribute(pt::pt_megakernel_wave<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
                PT_CUDA(cudaFuncSetAttribute(pt::pt_megakernel_wave<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
                s->ctas_per_sm = 1;
                // pt_debug_hits runs the lockstep kernel's sweep phase (same sweep_two / re-test code)
                const size_t dbg_smem = resident_smem(s->n_blocks, true);
                return s->d_motion ? configure_kernel(pt::pt_debug_hits_resident<true, true>, dbg_smem, nullptr)
                                   : configure_kernel(pt::pt_debug_hits_resident<true, false>, dbg_smem, nullptr);
            }
        }
        if (s->const_image) {
            rc = s->d_motion ? configure_kernel(pt::pt_megakernel_resident<true, true>, s->smem_bytes, &s->ctas_per_sm)
                             : configure_kernel(pt::pt_megakernel_resident<true, false>, s->smem_bytes, &s->ctas_per_sm);
            if (rc == PT_OK)
                rc = s->d_motion ? configure_kernel(pt::pt_debug_hits_resident<true, true>, s->smem_bytes, nullptr)
                                 : configure_kernel(pt::pt_debug_hits_resident<true, false>, s->smem_bytes, nullptr);
        } else {
            rc = s->d_motion ? configure_kernel(pt::pt_megakernel_resident<false, true>, s->smem_bytes, &s->ctas_per_sm)
                             : configure_kernel(pt::pt_megakernel_resident<false, false>, s->smem_bytes, &s->ctas_per_sm);
            if (rc == PT_OK)
                rc = s->d_motion ? configure_kernel(pt::pt_debug_hits_resident<false, true>, s->smem_bytes, nullptr)
                                 : configure_kernel(pt::pt_debug_hits_resident<false, false>, s->smem_bytes, nullptr);
        }
        return rc;
    }
    s->resident = false;
    s->const_image = false;
    // streamed kernel: two tile buffers; 2 CTAs per SM keeps the pipes fed while one CTA waits on a barrier.  The FP32 flavour
    // is always configured (it is the per-render fallback of the tensor path); the tensor-path flavour streams the fragment
    // image (128 B per block of 4 spheres instead of 64) and carries the CTA's transpose buffer.
    if (forced_tile > 0) forced_tile = (forced_tile + pt::kLdsGroupBlocks - 1) / pt::kLdsGroupBlocks * pt::kLdsGroupBlocks;  // whole groups
    auto geometry = [&](bool mma, int* tile_blocks, int* n_tiles, size_t* smem) {
        const size_t fixed = kStreamedFixedBytes + (mma ? (size_t)pt::kMmaStageRows * pt::kCtaThreads * sizeof(uint32_t) : 0);
        const size_t block_bytes = mma ? 128 : 64;
        if (forced_tile > 0 && s->n_blocks > 0) {  // PtOptions: any scene through the streamed kernel with n-block tiles
            *tile_blocks = std::min(forced_tile, s->n_blocks);
        } else {
            const int stream_ctas = s->opt.stream_ctas > 0 ? std::min(4, s->opt.stream_ctas) : 2;
            const size_t per_cta = (kMaxDynSmem + 1024) / stream_ctas - 2048;
            const size_t tile_bytes = ((per_cta - fixed) / 2) & ~(size_t)1023;
            *tile_blocks = (int)(tile_bytes / block_bytes) / pt::kLdsGroupBlocks * pt::kLdsGroupBlocks;
        }
        *n_tiles = (s->n_blocks + *tile_blocks - 1) / *tile_blocks;
        *smem = 2 * (size_t)*tile_blocks * block_bytes + fixed;
    };
    geometry(false, &s->tile_blocks, &s->n_tiles, &s->smem_bytes);
    int rc = s->d_motion ? configure_kernel(pt::pt_megakernel_streamed<true, false>, s->smem_bytes, &s->ctas_per_sm)
                         : configure_kernel(pt::pt_megakernel_streamed<false, false>, s->smem_bytes, &s->ctas_per_sm);
    if (rc == PT_OK)
        rc = s->d_motion ? configure_kernel(pt::pt_debug_hits_streamed<true, false>, s->smem_bytes, nullptr)
                         : configure_kernel(pt::pt_debug_hits_streamed<false, false>, s->smem_bytes, nullptr);
    s->fp32_smem_bytes = s->smem_bytes;
    s->fp32_ctas_per_sm = s->ctas_per_sm;
    s->fp32_tile_blocks = s->tile_blocks;
    s->fp32_n_tiles = s->n_tiles;
    const bool want_mma = s->mma_ok && (s->opt.resident_kernel == 0 || s->opt.resident_kernel == 5) && s->n_blocks / pt::kLdsGroupBlocks < 65536;
    if (rc == PT_OK && want_mma) {
        s->mma = true;
        geometry(true, &s->tile_blocks, &s->n_tiles, &s->smem_bytes);
        rc = s->d_motion ? configure_kernel(pt::pt_megakernel_streamed<true, true>, s->smem_bytes, &s->ctas_per_sm)
                         : configure_kernel(pt::pt_megakernel_streamed<false, true>, s->smem_bytes, &s->ctas_per_sm);
        if (rc == PT_OK)
            rc = s->d_motion ? configure_kernel(pt::pt_debug_hits_streamed<true, true>, s->smem_bytes, nullptr)
                             : configure_kernel(pt::pt_debug_hits_streamed<false, true>, s->smem_bytes, nullptr);
    }
    return rc;
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

int normalise_options(const PtOptions* in, PtOptions* out) {
    PtOptions o{};
    o.struct_size = sizeof(PtOptions);
    o.spatial_order = -1;
    if (in) {
        if (in->struct_size != sizeof(PtOptions)) return fail(PT_ERR_INVALID, "PtOptions.struct_size %u != %zu (ABI mismatch)", in->struct_size, sizeof(PtOptions));
        o = *in;
        if (o.force_stream_tile_blocks < 0 || o.stream_ctas < 0 || o.stream_ctas > 4 || o.chunk_samples < -1 || o.spatial_order < -1 || o.spatial_order > 2 || o.resident_kernel > 5)
            return fail(PT_ERR_INVALID, "PtOptions field out of range");
    }
    if (o.tile_rows == 0) o.tile_rows = 4;
    *out = o;
    return PT_OK;
}

int validate_params(const PtParams* params, const PtCamera* camera) {
    if (!params || !camera) return fail(PT_ERR_INVALID, "null params/camera");
    if (params->width == 0 || params->height == 0) return fail(PT_ERR_INVALID, "zero-sized image %ux%u", params->width, params->height);
    if ((uint64_t)params->width * params->height > 0xffffffffULL / 4) return fail(PT_ERR_TOO_LARGE, "image too large");
    // the reference divides by `samples` (scene.rs:85): zero samples would blend 0 * inf = NaN into every pixel
    if (params->samples == 0) return fail(PT_ERR_INVALID, "samples must be at least 1 (Scene::update divides by it, scene.rs:85)");
    if (params->use_bvh) return fail(PT_ERR_UNSUPPORTED, "use_bvh is not supported on the GPU path (flat sphere list only, params.rs:36-43)");
    return PT_OK;
}

// The pre-filter's validity domain (pt_sweep.cuh): its slack covers the rounding of the expanded discriminant only while
// the ray origin is not astronomically far from the scene.  Origins are the camera or points on sphere surfaces, so the
// camera is the one input to check per call: |origin| and the lens offset must stay below 2^20 times the scene's extent
// (and below 1e18, where |o|^2 meets the parked-lane sentinel).
int validate_camera_domain(const Replica* s, const PtCamera* cam) {
    for (int i = 0; i < 3; ++i) {
        const float v[7] = {cam->origin[i], cam->lower_left_corner[i], cam->horizontal[i], cam->vertical[i], cam->u[i], cam->v[i], cam->lens_radius};
        for (float f : v)
            if (!std::isfinite(f) || std::fabs(f) > 1.0e17f)
                return fail(PT_ERR_INVALID, "camera component %g is outside the sweep's validity domain (finite, |x| <= 1e17)", f);
    }
    (void)s;
    return PT_OK;
}

void fill_scene_args(const Replica* s, pt::KernelArgs& a) {
    a.blocks = s->d_blocks;
    a.n_blocks = s->n_blocks;
    a.n_spheres = (int)s->n_spheres;
    a.shade = s->d_shade;
    a.tex = s->d_tex;
    a.images = s->d_images;
    a.order = s->d_order;
    a.perlin = s->d_perlin;
    a.prefilter = s->d_prefilter;
    a.kplane = s->d_kplane;
    a.mma_image = s->d_mma_image;
    a.n_steps = s->n_blocks / pt::kLdsGroupBlocks;
    a.mma = s->mma_scale;
    a.motion = s->d_motion;
    a.has_noise = s->has_noise ? 1 : 0;
    a.tile_blocks = s->tile_blocks;
    a.n_tiles = s->n_tiles;
}

// every launch on a replica waits for the previous one (whatever stream it ran on) and leaves its own event behind
int serialise_begin(Replica* s, cudaStream_t stream) {
    if (s->busy_recorded) PT_CUDA(cudaStreamWaitEvent(stream, s->ev_busy, 0));
    return PT_OK;
}
int serialise_end(Replica* s, cudaStream_t stream) {
    PT_CUDA(cudaEventRecord(s->ev_busy, stream));
    s->busy_recorded = true;
    return PT_OK;
}

// enqueue one Scene::update on `stream`; d_rgb is the full-size device image
int launch_update(Replica* s, const PtParams* params, const PtCamera* cam, uint32_t frame_num, const PtPartition& part,
                  float* d_rgb, unsigned long long* d_ray_count, cudaStream_t stream) {
    pt::KernelArgs a{};
    fill_scene_args(s, a);
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
    a.status = s->d_status;

    int rc = validate_camera_domain(s, cam);
    if (rc != PT_OK) return rc;
    if (s->d_motion && !(cam->time0 >= s->motion_t_lo && cam->time1 <= s->motion_t_hi && cam->time0 <= cam->time1))
        return fail(PT_ERR_UNSUPPORTED, "camera shutter [%g, %g] is not inside the moving spheres' interval [%g, %g] (the pre-filter bounds their sweep over that interval)",
                    cam->time0, cam->time1, s->motion_t_lo, s->motion_t_hi);
    rc = serialise_begin(s, stream);
    if (rc != PT_OK) return rc;
    PT_CUDA(cudaMemsetAsync(s->d_next_pixel, 0, sizeof(unsigned int), stream));
    PT_CUDA(cudaMemsetAsync(s->d_status, 0, 16 * sizeof(unsigned int), stream));
    PT_CUDA(cudaMemsetAsync(d_ray_count, 0, sizeof(unsigned long long), stream));
    PT_CUDA(cudaMemsetAsync(s->d_sweep_count, 0, sizeof(unsigned long long), stream));
    s->stats.kernel_launches = 0;
    s->stats.grid_ctas = 0;
    if (a.n_owned_pixels == 0) return serialise_end(s, stream);

    // tensor-path sweep: its f16 operands are scaled for ray origins inside the scene's extent (bounce rays start on sphere
    // surfaces).  A camera outside it would send every primary ray down the exact-test-on-everything path (correct, slow):
    // such a render goes to the FP32 kernel instead.
    bool use_mma = s->mma;
    if (use_mma) {
        double reach = 0.0;
        const float off[3] = {s->mma_scale.tx, s->mma_scale.ty, s->mma_scale.tz};
        for (int i = 0; i < 3; ++i) {
            const double c = std::fabs((double)cam->origin[i] - (double)off[i]) + (double)std::fabs(cam->lens_radius) * (std::fabs(cam->u[i]) + std::fabs(cam->v[i]));
            reach += c * c;
        }
        if (!(reach <= (double)s->mma_scale.max_o2)) use_mma = false;
    }
    const int ctas_per_sm = (s->mma && !use_mma) ? s->fp32_ctas_per_sm : s->ctas_per_sm;
    const size_t smem_bytes = (s->mma && !use_mma) ? s->fp32_smem_bytes : s->smem_bytes;
    if (s->mma && !use_mma && !s->resident) {
        a.tile_blocks = s->fp32_tile_blocks;
        a.n_tiles = s->fp32_n_tiles;
    }

    // persistent grid: one wave of CTAs, never more lanes than pixels.  The resident kernel carries two paths per lane;
    // an image with fewer pixels than the machine has lanes keeps one path per lane (latency, not throughput, is what
    // counts there) and the second row stays parked.
    const uint32_t max_ctas = (uint32_t)(s->sm_count * ctas_per_sm);
    const uint32_t paths_per_cta = s->wave ? (uint32_t)s->wave_pool : (uint32_t)pt::kCtaThreads * (s->resident && !s->regroup ? pt::kPathRows : 1);
    a.wave_pool = s->wave_pool;
    uint32_t want = (a.n_owned_pixels + paths_per_cta - 1) / paths_per_cta;
    a.single_row = 0;
    if (s->resident && !s->wave && !s->regroup && (uint64_t)a.n_owned_pixels <= (uint64_t)max_ctas * pt::kCtaThreads) {
        a.single_row = 1;
        want = (a.n_owned_pixels + pt::kCtaThreads - 1) / pt::kCtaThreads;
    }
    const uint32_t grid = std::min<uint32_t>(max_ctas, want);

    // chunk queue (pt_megakernel.cuh, lane_refill): samples are handed out in chunks, sample-major, so that all pixels
    // finish together.  With fewer than two pixels per path every pixel starts at once and the launch lasts as long as
    // its slowest pixel whatever the unit: one chunk per pixel then, which skips the state table altogether.
    uint32_t chunk = 0;  // 0 = one chunk per pixel
    const uint64_t paths = (uint64_t)grid * (a.single_row ? pt::kCtaThreads : paths_per_cta);
    if (params->samples > 8 && (uint64_t)a.n_owned_pixels >= 2 * paths) {
        const uint32_t max_chunks = std::max<uint32_t>(1u, std::min<uint32_t>(64u, 0xF0000000u / a.n_owned_pixels));
        const uint32_t at_least = std::max<uint32_t>((params->samples + max_chunks - 1) / max_chunks, 8u);
        chunk = 8;
        while (chunk < at_least) chunk *= 2;
    }
    if (s->opt.chunk_samples != 0) {  // explicit PtOptions choice: a power of two, -1 = whole pixels
        chunk = 0;
        if (s->opt.chunk_samples > 0) {
            chunk = 1;
            while ((long)chunk < (long)s->opt.chunk_samples && chunk < (1u << 30)) chunk *= 2;
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
        a.chunk_samples = params->samples;
        a.chunk_mask = 0xffffffffu;
    }
    a.n_tickets = n_chunks * a.n_owned_pixels;
    a.pixstate = nullptr;
    if (n_chunks > 1) {
        if (s->d_pixstate_pixels < a.n_owned_pixels) {
            // (cudaFree synchronises the device: growing the table is the one blocking step of an otherwise asynchronous launch)
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
    if (s->regroup && use_mma && s->exact_smem) {
        if (s->d_motion) pt::pt_megakernel_regroup<true, true, true><<<grid, pt::kCtaThreads, s->mma_smem_bytes_exact, stream>>>(a);
        else pt::pt_megakernel_regroup<false, true, true><<<grid, pt::kCtaThreads, s->mma_smem_bytes_exact, stream>>>(a);
    } else if (s->regroup && use_mma) {
        if (s->d_motion) pt::pt_megakernel_regroup<true, true><<<grid, pt::kCtaThreads, smem_bytes, stream>>>(a);
        else pt::pt_megakernel_regroup<false, true><<<grid, pt::kCtaThreads, smem_bytes, stream>>>(a);
    } else if (s->regroup) {
        if (s->d_motion) pt::pt_megakernel_regroup<true, false><<<grid, pt::kCtaThreads, smem_bytes, stream>>>(a);
        else pt::pt_megakernel_regroup<false, false><<<grid, pt::kCtaThreads, smem_bytes, stream>>>(a);
    } else if (s->wave) {
        if (s->d_motion) pt::pt_megakernel_wave<true><<<grid, pt::kWaveThreads, s->smem_bytes, stream>>>(a, *s->h_const_image);
        else pt::pt_megakernel_wave<false><<<grid, pt::kWaveThreads, s->smem_bytes, stream>>>(a, *s->h_const_image);
    } else if (s->resident) {
        if (s->const_image) {
            if (s->d_motion) pt::pt_megakernel_resident<true, true><<<grid, pt::kCtaThreads, s->smem_bytes, stream>>>(a, *s->h_const_image);
            else pt::pt_megakernel_resident<true, false><<<grid, pt::kCtaThreads, s->smem_bytes, stream>>>(a, *s->h_const_image);
        } else {
            const pt::ConstImageT<false> none{};
            if (s->d_motion) pt::pt_megakernel_resident<false, true><<<grid, pt::kCtaThreads, s->smem_bytes, stream>>>(a, none);
            else pt::pt_megakernel_resident<false, false><<<grid, pt::kCtaThreads, s->smem_bytes, stream>>>(a, none);
        }
    } else if (use_mma) {
        if (s->d_motion) pt::pt_megakernel_streamed<true, true><<<grid, pt::kCtaThreads, smem_bytes, stream>>>(a);
        else pt::pt_megakernel_streamed<false, true><<<grid, pt::kCtaThreads, smem_bytes, stream>>>(a);
    } else {
        if (s->d_motion) pt::pt_megakernel_streamed<true, false><<<grid, pt::kCtaThreads, smem_bytes, stream>>>(a);
        else pt::pt_megakernel_streamed<false, false><<<grid, pt::kCtaThreads, smem_bytes, stream>>>(a);
    }
    PT_CUDA(cudaGetLastError());
    rc = serialise_end(s, stream);
    if (rc != PT_OK) return rc;
    s->stats.kernel_launches = 1;
    s->stats.grid_ctas = grid;
    s->stats.cta_threads = s->wave ? pt::kWaveThreads : pt::kCtaThreads;
    s->stats.smem_bytes = (uint32_t)((s->regroup && use_mma && s->exact_smem) ? s->mma_smem_bytes_exact : smem_bytes);
    s->stats.resident = s->resident ? ((s->regroup && use_mma) ? 2u : 1u) : (use_mma ? 3u : 0u);
    s->stats.n_spheres = s->n_spheres;
    return PT_OK;
}

int ensure_image(Replica* s, size_t floats) {
    if (s->d_rgb_floats >= floats) return PT_OK;
    if (s->d_rgb) cudaFree(s->d_rgb);
    s->d_rgb = nullptr;
    s->d_rgb_floats = 0;
    PT_CUDA(cudaMalloc(&s->d_rgb, floats * sizeof(float)));
    s->d_rgb_floats = floats;
    return PT_OK;
}
int ensure_rgb8(Replica* s, size_t bytes) {
    if (s->d_rgb8_bytes >= bytes) return PT_OK;
    if (s->d_rgb8) cudaFree(s->d_rgb8);
    s->d_rgb8 = nullptr;
    s->d_rgb8_bytes = 0;
    PT_CUDA(cudaMalloc(&s->d_rgb8, bytes));
    s->d_rgb8_bytes = bytes;
    return PT_OK;
}

// copy the rows a partition owns between host and device images (same layout on both sides)
int copy_owned_rows(Replica* s, const PtParams* params, const PtPartition& part, float* host, bool to_device, uint64_t* bytes_out) {
    const size_t row_bytes = (size_t)params->width * 3 * sizeof(float);
    uint64_t bytes = 0;
    if (part.part_count <= 1) {
        bytes = row_bytes * params->height;
        PT_CUDA(cudaMemcpyAsync(to_device ? (void*)s->d_rgb : (void*)host, to_device ? (const void*)host : (const void*)s->d_rgb, bytes,
                                to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, s->stream));
    } else {
        // one strided copy for all owned tiles: tile k of this part starts part_count * tile_rows rows after tile k-1
        const uint32_t n_tiles = (params->height + part.tile_rows - 1) / part.tile_rows;
        const uint32_t full_tiles = params->height / part.tile_rows;  // tiles with all their rows
        uint32_t n_mine_full = 0;
        for (uint32_t k = part.part_index; k < full_tiles; k += part.part_count) ++n_mine_full;
        const size_t first = (size_t)part.part_index * part.tile_rows * params->width * 3;
        if (n_mine_full > 0) {
            const size_t pitch = row_bytes * part.tile_rows * part.part_count, wbytes = row_bytes * part.tile_rows;
            PT_CUDA(cudaMemcpy2DAsync(to_device ? (void*)(s->d_rgb + first) : (void*)(host + first), pitch,
                                      to_device ? (const void*)(host + first) : (const void*)(s->d_rgb + first), pitch, wbytes, n_mine_full,
                                      to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, s->stream));
            bytes += (uint64_t)wbytes * n_mine_full;
        }
        if (n_tiles > full_tiles && (n_tiles - 1) % part.part_count == part.part_index) {  // the ragged last tile is ours
            const uint32_t r0 = (n_tiles - 1) * part.tile_rows;
            const size_t off = (size_t)r0 * params->width * 3, nbytes = row_bytes * (params->height - r0);
            PT_CUDA(cudaMemcpyAsync(to_device ? (void*)(s->d_rgb + off) : (void*)(host + off), to_device ? (const void*)(host + off) : (const void*)(s->d_rgb + off),
                                    nbytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, s->stream));
            bytes += nbytes;
        }
    }
    *bytes_out = bytes;
    return PT_OK;
}

// Tensor-path pre-filter image (pt_sweep_mma.cuh).  bound: per stored sphere the centre and radius the filter tests.
// extent = twice the reach of the spheres (so a camera up to one scene-size away still renders on this path); sigma maps it
// to kMmaExtentTarget, s keeps K' / s and sigma^2 |o|^2 / s inside the f16 range.  Not built (mma_ok = false) for scenes
// where it cannot pay: fewer than 128 spheres (one or two loop steps: the FP32 loop has less set-up per sweep), or spheres so
// small against the scene's extent that the absolute slack of the f16 subnormal range would flag them for every nearby ray.
void build_mma_image(FlatScene& fs, const std::vector<double>& bound) {
    fs.mma_ok = false;
    const uint32_t n = fs.n_spheres;
    if (n < 128) return;
    // the scene's offset t: per axis the middle of the spheres' bounding box, used only where the scene lies at more than four
    // times its own reach from the origin on that axis (then o - t is exact in f32 for every origin inside the extent, so the
    // filter sees exactly the ray the exact test sees); 0 elsewhere — every preset of the reference.
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (uint32_t j = 0; j < n; ++j)
        for (int a = 0; a < 3; ++a) {
            const double* b = &bound[(size_t)j * 4];
            if (!std::isfinite(b[a]) || !std::isfinite(b[3])) return;
            lo[a] = std::min(lo[a], b[a] - b[3]);
            hi[a] = std::max(hi[a], b[a] + b[3]);
        }
    double t[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
    double centred_reach = 0.0;
    for (uint32_t j = 0; j < n; ++j) {
        const double* b = &bound[(size_t)j * 4];
        centred_reach = std::max(centred_reach, std::sqrt((b[0] - t[0]) * (b[0] - t[0]) + (b[1] - t[1]) * (b[1] - t[1]) + (b[2] - t[2]) * (b[2] - t[2])) + b[3]);
    }
    for (int a = 0; a < 3; ++a) {
        t[a] = (double)(float)t[a];  // the kernel subtracts the f32 value
        if (!(std::fabs(t[a]) > 4.0 * 2.0 * centred_reach)) t[a] = 0.0;  // (extent = twice the reach)
    }
    double reach = 0.0;
    std::vector<double> r2s;
    r2s.reserve(n);
    std::vector<double> shifted(bound);
    for (uint32_t j = 0; j < n; ++j) {
        double* b = &shifted[(size_t)j * 4];
        for (int a = 0; a < 3; ++a) b[a] -= t[a];
        const double d = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]) + b[3];
        if (!std::isfinite(d)) return;
        reach = std::max(reach, d);
        r2s.push_back(b[3] * b[3]);
    }
    const double extent = 2.0 * reach;
    if (!(extent > 1.0e-12 && extent < 1.0e15)) return;
    const double sigma = std::exp2(std::floor(std::log2(pt::kMmaExtentTarget / extent)));
    const double s = std::exp2(std::ceil(std::log2(sigma * sigma * extent * extent / 32768.0)));
    std::nth_element(r2s.begin(), r2s.begin() + r2s.size() / 2, r2s.end());
    const double median_r2 = r2s[r2s.size() / 2];
    if (!(pt::kMmaAbsSlack / (sigma * sigma) <= 0.05 * median_r2)) return;
    fs.mma.sigma = (float)sigma;
    fs.mma.s = (float)s;
    fs.mma.inv_s = (float)(1.0 / s);
    fs.mma.max_o2 = (float)(extent * extent * (1.0 - 1e-6));
    fs.mma.tx = (float)t[0];
    fs.mma.ty = (float)t[1];
    fs.mma.tz = (float)t[2];
    const int n_steps = fs.n_blocks / pt::kLdsGroupBlocks;
    auto h16 = [](float x) { const __half h = __float2half_rn(x); unsigned short u; std::memcpy(&u, &h, 2); return u; };
    auto h2f = [](unsigned short u) { __half h; std::memcpy(&h, &u, 2); return __half2float(h); };
    std::vector<unsigned short> row((size_t)n_steps * 16 * 16, 0);  // per stored sphere its 16 K halves
    for (int j = 0; j < n_steps * 16; ++j) {
        float S[5] = {0.0f, 0.0f, 0.0f, (float)s, -65504.0f};  // padding: B' <= -65504 s + slack terms < 0 for every ray in range
        if ((uint32_t)j < n) {
            const double* b = &shifted[(size_t)j * 4];
            const double c2 = b[0] * b[0] + b[1] * b[1] + b[2] * b[2], r2 = b[3] * b[3];
            const double Kp = sigma * sigma * (r2 - c2 + pt::kMmaSlackSphere * (c2 + r2)) + pt::kMmaAbsSlack;
            S[0] = (float)(sigma * b[0]);
            S[1] = (float)(sigma * b[1]);
            S[2] = (float)(sigma * b[2]);
            S[4] = std::nextafter((float)(Kp / s), INFINITY);  // round towards "candidate"
        }
        for (int e = 0; e < 5; ++e) {
            const unsigned short hi = h16(S[e]);
            const unsigned short lo = h16(S[e] - h2f(hi));
            row[(size_t)j * 16 + e] = hi;
            row[(size_t)j * 16 + 5 + e] = hi;
            row[(size_t)j * 16 + 10 + e] = lo;
        }
    }
    // B-operand fragments: lane (g, t) of step st holds {b0, b1} of sphere group 0 (sphere 16 st + g) and of group 1
    // (sphere 16 st + 8 + g): K halves (2t, 2t+1) and (2t+8, 2t+9).  One zero step of padding (the loop fetches ahead).
    fs.mma_image.assign((size_t)(n_steps + 1) * 32, make_uint4(0u, 0u, 0u, 0u));
    for (int st = 0; st < n_steps; ++st)
        for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3;
            auto pk = [&](int r, int k) { return (uint32_t)row[(size_t)(st * 16 + r) * 16 + k] | ((uint32_t)row[(size_t)(st * 16 + r) * 16 + k + 1] << 16); };
            fs.mma_image[(size_t)st * 32 + lane] = make_uint4(pk(g, 2 * t), pk(g, 2 * t + 8), pk(8 + g, 2 * t), pk(8 + g, 2 * t + 8));
        }
    fs.mma_rows.assign(row.begin(), row.end());
    fs.mma_ok = true;
}


// Storage order of a scene's spheres (see pt_scene_create).  Fills order_of[j] = position in the caller's list of the
// sphere stored at j and returns the mode: 0 the caller's order, 1 Morton, 2 large spheres first, then Morton.
int storage_order(const PtSceneDesc* desc, bool any_moving, const PtOptions& opt, std::vector<uint32_t>& order_of) {
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
    if (opt.spatial_order >= 0) order_mode = n > 1 ? std::min(2, opt.spatial_order) : 0;  // explicit PtOptions choice
    // resident kernel only: the streamed kernel keeps list order and the plain index tie rule (with ~10^5 spheres few groups
    // are flagged anyway: 66.0 -> 66.5 % on cfg5, and the out-of-line tie rule costs that kernel more than it gains)
    const int n_blocks_planned = (int)(((n + 3) / 4 + pt::kLdsGroupBlocks - 1) / pt::kLdsGroupBlocks * pt::kLdsGroupBlocks);
    if (!fits_resident(n_blocks_planned, opt)) order_mode = 0;
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

// ---- flatten the caller's scene once on the host (the GPU arm of `Params::new_scene` + `SpheresSoA::new`) ----
int flatten_scene(const PtSceneDesc* desc, const PtOptions& opt, FlatScene& fs) {
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

    // ---- validate materials/textures ----
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
    // the sweep's validity domain (pt_sweep.cuh): finite geometry; magnitudes are handled per sphere below (a sphere too
    // large or too far for the expanded form is flagged for every ray and decided by the exact test alone)
    for (uint32_t i = 0; i < n; ++i) {
        if (!std::isfinite(desc->centre_x[i]) || !std::isfinite(desc->centre_y[i]) || !std::isfinite(desc->centre_z[i]) || !std::isfinite(desc->radius[i]))
            return fail(PT_ERR_INVALID, "sphere %u: centre/radius must be finite", i);
        if (is_moving(i) && !(std::isfinite(desc->motion[i].centre1[0]) && std::isfinite(desc->motion[i].centre1[1]) && std::isfinite(desc->motion[i].centre1[2])))
            return fail(PT_ERR_INVALID, "sphere %u: centre1 must be finite", i);
    }

    fs.any_moving = any_moving;
    fs.spatial = storage_order(desc, any_moving, opt, fs.order_of) != 0;
    const std::vector<uint32_t>& order_of = fs.order_of;
    fs.n_spheres = n;
    fs.n_blocks = (int)((n + 3) / 4);
    fs.n_blocks = (fs.n_blocks + pt::kLdsGroupBlocks - 1) / pt::kLdsGroupBlocks * pt::kLdsGroupBlocks;  // whole groups; padding spheres can never be hit
    if (fs.n_blocks > pt::kMaxSweepBlocks) return fail(PT_ERR_TOO_LARGE, "too many spheres for the candidate-queue encoding: %u", n);
    fs.has_noise = uses_noise;
    fs.has_sky = desc->has_sky != 0;
    fs.sky = pt::V3{desc->sky[0], desc->sky[1], desc->sky[2]};
    fs.motion_t_lo = t_lo;
    fs.motion_t_hi = t_hi;

    // ---- sphere blocks: X,Y,Z,R^2 for 4 spheres; padding = (FLT_MAX centre, r^2 = 0) spheres_soa.rs:53-61 ----
    fs.blocks.assign((size_t)std::max(fs.n_blocks, 1) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    fs.shade.assign(std::max<uint32_t>(n, 1), pt::DevShade{});
    for (int j = 0; j < fs.n_blocks; ++j) {
        float* f = reinterpret_cast<float*>(&fs.blocks[(size_t)j * 4]);
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
    // pre-filter image (pt_sweep.cuh): X, Y, Z, K = r^2 - |c|^2 + slack, padded to the group.
    // slack = 2^-18 (|c|^2 + r^2): twice the worst-case rounding budget of the expanded form against the reference's
    // sphere-relative one for |o| <~ |c| (7 fused steps + the unfused o.d and |o|^2 + the reference's own discriminant
    // error), so every sphere the exact expression accepts is a candidate.  Proven per ray, not assumed:
    // pt_debug_hits mode 0 against mode 1 in tests/test_gpu_hits.py.
    fs.prefilter.assign((size_t)std::max(fs.n_blocks, 1) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    std::vector<double> bound((size_t)n * 4, 0.0);  // per stored sphere: centre and radius of what the filter tests (a MovingSphere's static bound)
    for (int j = 0; j < fs.n_blocks; ++j) {
        float* f = reinterpret_cast<float*>(&fs.prefilter[(size_t)j * 4]);
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
            double* bd = &bound[((size_t)j * 4 + e) * 4];
            bd[0] = cx; bd[1] = cy; bd[2] = cz; bd[3] = r;
            if (!(c2 + r2 < 1.0e24)) {  // too large (or NaN) for the expanded form: always a candidate, the exact test decides
                f[12 + e] = INFINITY;
                continue;
            }
            f[0 + e] = (float)cx;
            f[4 + e] = (float)cy;
            f[8 + e] = (float)cz;
            const double k = r2 - c2 + pt::kSlackSphere * (c2 + r2);
            f[12 + e] = std::nextafter((float)k, INFINITY);  // round towards "candidate"
        }
    }
    build_mma_image(fs, bound);
    // K plane alone + the kernel-parameter image of the X, Y, Z planes (resident kernel, small scenes)
    fs.kplane.assign((size_t)std::max(fs.n_blocks, 1), make_float4(0.f, 0.f, 0.f, 0.f));
    for (int j = 0; j < fs.n_blocks; ++j) fs.kplane[j] = fs.prefilter[(size_t)j * 4 + 3];
    if (fs.n_blocks <= pt::kMaxConstBlocks) {
        fs.const_image.reset(new pt::ConstImageT<true>());
        std::memset(fs.const_image.get(), 0, sizeof(pt::ConstImageT<true>));
        for (int j = 0; j < fs.n_blocks; ++j)
            for (int c = 0; c < 3; ++c) fs.const_image->v[3 * j + c] = fs.prefilter[(size_t)j * 4 + c];
    }
    for (uint32_t j = 0; j < n; ++j) {
        const uint32_t i = order_of[j];
        const int32_t mi = desc->material_index[i];
        if (mi < 0 || (uint32_t)mi >= desc->n_materials) return fail(PT_ERR_INVALID, "sphere %u: material index %d out of range", i, mi);
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
        fs.shade[j] = d;
    }
    fs.tex.assign(std::max<uint32_t>(desc->n_textures, 1), pt::DevTexture{});
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
        fs.tex[t] = d;
    }
    if (image_pool_bytes > 0) {
        fs.image_pool.assign(image_pool_bytes, 0);
        for (uint32_t i = 0; i < desc->n_images; ++i)
            std::memcpy(fs.image_pool.data() + image_offset[i], desc->images[i].data, (size_t)desc->images[i].width * desc->images[i].height * 3);
    }
    if (any_moving) {
        fs.motion.assign(n, pt::DevMotion{});
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
            fs.motion[j] = m;
        }
    }
    fs.perlin_raw.assign(sizeof(pt::PerlinSmem), 0);
    if (desc->perlin) {
        pt::PerlinSmem* ps = reinterpret_cast<pt::PerlinSmem*>(fs.perlin_raw.data());
        for (int i = 0; i < 256; ++i) {
            ps->randvec[i] = make_float4(desc->perlin->randvec[i][0], desc->perlin->randvec[i][1], desc->perlin->randvec[i][2], 0.0f);
            if (desc->perlin->perm_x[i] > 255 || desc->perlin->perm_y[i] > 255 || desc->perlin->perm_z[i] > 255)
                return fail(PT_ERR_INVALID, "perlin permutation entry %d out of range", i);
            ps->perm_x[i] = (uint8_t)desc->perlin->perm_x[i];
            ps->perm_y[i] = (uint8_t)desc->perlin->perm_y[i];
            ps->perm_z[i] = (uint8_t)desc->perlin->perm_z[i];
        }
    }
    return PT_OK;
}

void destroy_replica(Replica* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->busy_recorded) cudaEventSynchronize(s->ev_busy);
    cudaFree(s->d_blocks);
    cudaFree(s->d_shade);
    cudaFree(s->d_tex);
    cudaFree(s->d_images);
    cudaFree(s->d_order);
    cudaFree(s->d_perlin);
    cudaFree(s->d_prefilter);
    cudaFree(s->d_kplane);
    cudaFree(s->d_mma_image);
    cudaFree(s->d_motion);
    cudaFree(s->d_ray_count);
    cudaFree(s->d_sweep_count);
    cudaFree(s->d_next_pixel);
    cudaFree(s->d_status);
    cudaFree(s->d_pixstate);
    cudaFree(s->d_rgb);
    cudaFree(s->d_rgb8);
    for (auto& e : s->ev)
        if (e) cudaEventDestroy(e);
    if (s->ev_busy) cudaEventDestroy(s->ev_busy);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

template <typename T>
int upload(T** dst, const std::vector<T>& src) {
    if (src.empty()) return PT_OK;
    PT_CUDA(cudaMalloc(dst, src.size() * sizeof(T)));
    PT_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return PT_OK;
}

int create_replica(const FlatScene& fs, int device, const PtOptions& opt, Replica** out) {
    cudaDeviceProp prop;
    int rc = check_device(device, &prop);
    if (rc != PT_OK) return rc;
    PT_CUDA(cudaSetDevice(device));
    Replica* s = new Replica();
    s->device = device;
    s->sm_count = prop.multiProcessorCount;
    s->n_spheres = fs.n_spheres;
    s->n_blocks = fs.n_blocks;
    s->has_noise = fs.has_noise;
    s->has_sky = fs.has_sky;
    s->sky = fs.sky;
    s->opt = opt;
    s->motion_t_lo = fs.motion_t_lo;
    s->motion_t_hi = fs.motion_t_hi;
    s->h_const_image = fs.const_image.get();
    rc = upload(&s->d_blocks, fs.blocks);
    if (rc == PT_OK) rc = upload(&s->d_shade, fs.shade);
    if (rc == PT_OK) rc = upload(&s->d_tex, fs.tex);
    if (rc == PT_OK && fs.spatial) rc = upload(&s->d_order, fs.order_of);
    if (rc == PT_OK) rc = upload(&s->d_images, fs.image_pool);
    if (rc == PT_OK) rc = upload(&s->d_motion, fs.motion);
    if (rc == PT_OK) rc = upload(&s->d_prefilter, fs.prefilter);
    if (rc == PT_OK) rc = upload(&s->d_kplane, fs.kplane);
    if (rc == PT_OK && fs.mma_ok) rc = upload(&s->d_mma_image, fs.mma_image);
    s->mma_ok = fs.mma_ok;
    s->mma_scale = fs.mma;
    if (rc == PT_OK) {
        std::vector<unsigned char> raw = fs.perlin_raw;
        unsigned char* d = nullptr;
        rc = upload(&d, raw);
        s->d_perlin = reinterpret_cast<pt::PerlinSmem*>(d);
    }
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (rc == PT_OK && e != cudaSuccess) rc = fail(PT_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    };
    if (rc == PT_OK) cuda_ok(cudaMalloc(&s->d_ray_count, sizeof(unsigned long long)), "cudaMalloc");
    if (rc == PT_OK) cuda_ok(cudaMalloc(&s->d_sweep_count, sizeof(unsigned long long)), "cudaMalloc");
    if (rc == PT_OK) cuda_ok(cudaMalloc(&s->d_next_pixel, sizeof(unsigned int)), "cudaMalloc");
    if (rc == PT_OK) cuda_ok(cudaMalloc(&s->d_status, 16 * sizeof(unsigned int)), "cudaMalloc");
    if (rc == PT_OK) cuda_ok(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    for (auto& e : s->ev)
        if (rc == PT_OK) cuda_ok(cudaEventCreate(&e), "cudaEventCreate");
    if (rc == PT_OK) cuda_ok(cudaEventCreateWithFlags(&s->ev_busy, cudaEventDisableTiming), "cudaEventCreate");
    if (rc == PT_OK) rc = plan_launch(s);
    if (rc != PT_OK) {
        const std::string keep = g_last_error;
        destroy_replica(s);
        g_last_error = keep;
        return rc;
    }
    *out = s;
    return PT_OK;
}

// one Scene::update of `part` on one replica with HOST buffers (H2D of the previous frame, kernel, D2H)
int render_part_host(Replica* s, const PtParams* params, const PtCamera* camera, uint32_t frame_num, const PtPartition& part, float* rgb_inout,
                     uint64_t* ray_count_out) {
    PT_CUDA(cudaSetDevice(s->device));
    const size_t floats = (size_t)params->width * params->height * 3;
    int rc = ensure_image(s, floats);
    if (rc != PT_OK) return rc;
    s->stats = PtRenderStats{};
    uint64_t h2d = 0, d2h = 0;
    PT_CUDA(cudaEventRecord(s->ev[0], s->stream));
    if (frame_num != 0) {  // the blend reads the previous frame (scene.rs:114-116); frame 0 has mix_prev = 0
        rc = serialise_begin(s, s->stream);  // the device image may still be read by an earlier asynchronous launch
        if (rc != PT_OK) return rc;
        rc = copy_owned_rows(s, params, part, rgb_inout, true, &h2d);
        if (rc != PT_OK) return rc;
    }
    PT_CUDA(cudaEventRecord(s->ev[1], s->stream));
    rc = launch_update(s, params, camera, frame_num, part, s->d_rgb, s->d_ray_count, s->stream);
    if (rc != PT_OK) return rc;
    PT_CUDA(cudaEventRecord(s->ev[2], s->stream));
    rc = copy_owned_rows(s, params, part, rgb_inout, false, &d2h);
    if (rc != PT_OK) return rc;
    unsigned long long rays = 0, sweeps = 0;
    PT_CUDA(cudaMemcpyAsync(&rays, s->d_ray_count, sizeof(rays), cudaMemcpyDeviceToHost, s->stream));
    PT_CUDA(cudaMemcpyAsync(&sweeps, s->d_sweep_count, sizeof(sweeps), cudaMemcpyDeviceToHost, s->stream));
    unsigned int status_words[16] = {0};
    PT_CUDA(cudaMemcpyAsync(status_words, s->d_status, sizeof(status_words), cudaMemcpyDeviceToHost, s->stream));
    PT_CUDA(cudaEventRecord(s->ev[3], s->stream));
    PT_CUDA(cudaStreamSynchronize(s->stream));
    if (status_words[0] != 0)
        return fail(PT_ERR_CUDA, "render kernel watchdog fired (code %u: %s; live %d, lock %u, queues %u %u %u %u %u, CTA %u thread %u): the image is incomplete",
                    status_words[0], status_words[0] & 1u ? "lost queue lock" : "warp starved with paths outstanding", (int)status_words[1], status_words[2],
                    status_words[3], status_words[4], status_words[5], status_words[6], status_words[7], status_words[8], status_words[9]);
    s->stats.warp_sweeps = sweeps;
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
    return PT_OK;
}

// one frame of the resident progressive accumulation of `part` on one replica (image stays on the device)
int render_part_progressive(Replica* s, const PtParams* params, const PtCamera* camera, uint32_t frame_num, const PtPartition& part, float* rgb_out,
                            uint8_t* rgb8_out, uint64_t* ray_count_out) {
    PT_CUDA(cudaSetDevice(s->device));
    const size_t n = (size_t)params->width * params->height;
    int rc = ensure_image(s, n * 3);
    if (rc != PT_OK) return rc;
    s->stats = PtRenderStats{};
    PT_CUDA(cudaEventRecord(s->ev[1], s->stream));
    rc = launch_update(s, params, camera, frame_num, part, s->d_rgb, s->d_ray_count, s->stream);
    if (rc != PT_OK) return rc;
    PT_CUDA(cudaEventRecord(s->ev[2], s->stream));
    uint64_t d2h = 0;
    if (rgb_out) {
        rc = copy_owned_rows(s, params, part, rgb_out, false, &d2h);
        if (rc != PT_OK) return rc;
    }
    if (rgb8_out) {
        rc = ensure_rgb8(s, n * 3);
        if (rc != PT_OK) return rc;
        const int threads = 256;
        const int blocks = (int)std::min<size_t>((n + threads - 1) / threads, (size_t)s->sm_count * 8);
        pt::pt_srgb8_kernel<<<blocks, threads, 0, s->stream>>>(s->d_rgb, params->width, params->height, s->d_rgb8);
        PT_CUDA(cudaGetLastError());
        // the sRGB image is top-down: bottom-up row y of the accumulation buffer is row height-1-y there
        const size_t row8 = (size_t)params->width * 3;
        if (part.part_count <= 1) {
            PT_CUDA(cudaMemcpyAsync(rgb8_out, s->d_rgb8, n * 3, cudaMemcpyDeviceToHost, s->stream));
            d2h += n * 3;
        } else {
            const uint32_t n_tiles = (params->height + part.tile_rows - 1) / part.tile_rows;
            for (uint32_t k = part.part_index; k < n_tiles; k += part.part_count) {
                const uint32_t r0 = k * part.tile_rows, nr = std::min(part.tile_rows, params->height - r0);
                const size_t off = (size_t)(params->height - r0 - nr) * row8;
                PT_CUDA(cudaMemcpyAsync(rgb8_out + off, s->d_rgb8 + off, row8 * nr, cudaMemcpyDeviceToHost, s->stream));
                d2h += row8 * nr;
            }
        }
    }
    unsigned long long rays = 0, sweeps = 0;
    PT_CUDA(cudaMemcpyAsync(&rays, s->d_ray_count, sizeof(rays), cudaMemcpyDeviceToHost, s->stream));
    PT_CUDA(cudaMemcpyAsync(&sweeps, s->d_sweep_count, sizeof(sweeps), cudaMemcpyDeviceToHost, s->stream));
    unsigned int status_words[16] = {0};
    PT_CUDA(cudaMemcpyAsync(status_words, s->d_status, sizeof(status_words), cudaMemcpyDeviceToHost, s->stream));
    PT_CUDA(cudaEventRecord(s->ev[3], s->stream));
    PT_CUDA(cudaStreamSynchronize(s->stream));
    if (status_words[0] != 0)
        return fail(PT_ERR_CUDA, "render kernel watchdog fired (code %u: %s; live %d, lock %u, queues %u %u %u %u %u, CTA %u thread %u): the image is incomplete",
                    status_words[0], status_words[0] & 1u ? "lost queue lock" : "warp starved with paths outstanding", (int)status_words[1], status_words[2],
                    status_words[3], status_words[4], status_words[5], status_words[6], status_words[7], status_words[8], status_words[9]);
    s->stats.warp_sweeps = sweeps;
    float ms = 0;
    PT_CUDA(cudaEventElapsedTime(&ms, s->ev[1], s->ev[2]));
    s->stats.kernel_ms = ms;
    PT_CUDA(cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]));
    s->stats.d2h_ms = ms;
    s->stats.d2h_bytes = d2h + sizeof(rays);
    s->stats.ray_count = rays;
    if (ray_count_out) *ray_count_out = rays;
    return PT_OK;
}

// Run `fn(replica, part_of_that_replica, &rays)` for every replica of the scene: inline for one device, one host thread
// per GPU otherwise.  Ray counts are summed; the first failure wins and its message reaches the caller's
// thread-local pt_last_error().
template <typename F>
int for_each_replica(PtScene* sc, F fn, uint64_t* ray_count_out) {
    const uint32_t n = (uint32_t)sc->reps.size();
    std::vector<int> rc(n, PT_OK);
    std::vector<uint64_t> rays(n, 0);
    std::vector<std::string> msg(n);
    auto part_of = [&](uint32_t i) { return n == 1 ? PtPartition{4, 0, 1, 0} : PtPartition{sc->opt.tile_rows, i, n, 0}; };
    if (n == 1) {
        rc[0] = fn(sc->reps[0], part_of(0), &rays[0]);
    } else {
        std::vector<std::thread> th;
        th.reserve(n);
        for (uint32_t i = 0; i < n; ++i)
            th.emplace_back([&, i] {
                rc[i] = fn(sc->reps[i], part_of(i), &rays[i]);
                if (rc[i] != PT_OK) msg[i] = g_last_error;  // this worker's thread-local message
            });
        for (auto& t : th) t.join();
    }
    // aggregate statistics: times are the slowest device's, bytes and launches add up
    sc->stats = PtRenderStats{};
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const PtRenderStats& st = sc->reps[i]->stats;
        sc->stats.kernel_ms = std::max(sc->stats.kernel_ms, st.kernel_ms);
        sc->stats.h2d_ms = std::max(sc->stats.h2d_ms, st.h2d_ms);
        sc->stats.d2h_ms = std::max(sc->stats.d2h_ms, st.d2h_ms);
        sc->stats.h2d_bytes += st.h2d_bytes;
        sc->stats.d2h_bytes += st.d2h_bytes;
        sc->stats.kernel_launches += st.kernel_launches;
        sc->stats.grid_ctas += st.grid_ctas;
        sc->stats.cta_threads = st.cta_threads;
        sc->stats.smem_bytes = st.smem_bytes;
        sc->stats.resident = st.resident;
        sc->stats.n_spheres = st.n_spheres;
        sc->stats.warp_sweeps += st.warp_sweeps;
        total += rays[i];
    }
    sc->stats.ray_count = total;
    for (uint32_t i = 0; i < n; ++i)
        if (rc[i] != PT_OK) {
            if (n > 1) g_last_error = "device " + std::to_string(sc->reps[i]->device) + ": " + msg[i];
            return rc[i];
        }
    if (ray_count_out) *ray_count_out = total;
    return PT_OK;
}

int need_single_device(const PtScene* sc, const char* what) {
    if (sc->reps.size() != 1) return fail(PT_ERR_UNSUPPORTED, "%s takes device pointers and needs a one-device scene (this one spans %zu devices)", what, sc->reps.size());
    return PT_OK;
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
        case 11: return sizeof(PtOptions);
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

uint32_t pt_scene_storage_order(const PtSceneDesc* desc, const PtOptions* options, uint32_t* order_out, uint32_t cap) {
    if (!desc || desc->struct_size != sizeof(PtSceneDesc)) return 0;
    const uint32_t n = desc->n_spheres;
    if (n > 0 && (!desc->centre_x || !desc->centre_y || !desc->centre_z || !desc->radius)) return 0;
    bool any_moving = false;
    if (desc->motion)
        for (uint32_t i = 0; i < n; ++i) any_moving = any_moving || desc->motion[i].moving != 0;
    std::vector<uint32_t> order_of;
    PtOptions opt;
    if (normalise_options(options, &opt) != PT_OK) return 0;
    const int mode = storage_order(desc, any_moving, opt, order_of);
    for (uint32_t j = 0; j < n && j < cap && order_out; ++j) order_out[j] = order_of[j];
    return (uint32_t)mode;
}

uint32_t pt_scene_mma_operand(const PtSceneDesc* desc, const PtOptions* options, uint16_t* rows_out, uint32_t cap, float* scale_out, uint32_t* order_out) {
    if (!desc) return 0;
    PtOptions opt;
    if (normalise_options(options, &opt) != PT_OK) return 0;
    FlatScene fs;
    if (flatten_scene(desc, opt, fs) != PT_OK || !fs.mma_ok) return 0;
    const uint32_t stored = (uint32_t)(fs.mma_rows.size() / 16);
    for (uint32_t j = 0; j < stored && j < cap && rows_out; ++j) std::memcpy(rows_out + (size_t)j * 16, &fs.mma_rows[(size_t)j * 16], 16 * sizeof(uint16_t));
    if (scale_out) {
        scale_out[0] = fs.mma.sigma;
        scale_out[1] = fs.mma.s;
        scale_out[2] = fs.mma.inv_s;
        scale_out[3] = fs.mma.max_o2;
        scale_out[4] = fs.mma.tx;
        scale_out[5] = fs.mma.ty;
        scale_out[6] = fs.mma.tz;
    }
    for (uint32_t j = 0; j < fs.n_spheres && j < cap && order_out; ++j) order_out[j] = fs.order_of[j];
    return stored;
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

int pt_scene_create_multi(const PtSceneDesc* desc, const int* devices, uint32_t n_devices, const PtOptions* options, PtScene** out) {
    if (!desc || !out) return fail(PT_ERR_INVALID, "null desc/out");
    *out = nullptr;
    if (!devices || n_devices == 0) return fail(PT_ERR_INVALID, "empty device list");
    if (n_devices > 64) return fail(PT_ERR_INVALID, "too many devices: %u", n_devices);
    PtOptions opt;
    int rc = normalise_options(options, &opt);
    if (rc != PT_OK) return rc;
    std::unique_ptr<PtScene> sc(new PtScene());
    sc->opt = opt;
    rc = flatten_scene(desc, opt, sc->flat);
    if (rc != PT_OK) return rc;
    // devices first: "no device" is the answer on a machine without a B200, whatever else is wrong
    for (uint32_t i = 0; i < n_devices && rc == PT_OK; ++i) {
        Replica* r = nullptr;
        rc = create_replica(sc->flat, devices[i], opt, &r);
        if (rc == PT_OK) sc->reps.push_back(r);
    }
    if (rc != PT_OK) {
        const std::string keep = g_last_error;
        for (Replica* r : sc->reps) destroy_replica(r);
        g_last_error = keep;
        return rc;
    }
    *out = sc.release();
    return PT_OK;
}

int pt_scene_create(const PtSceneDesc* desc, int device, PtScene** out) { return pt_scene_create_multi(desc, &device, 1, nullptr, out); }

void pt_scene_destroy(PtScene* sc) {
    if (!sc) return;
    for (Replica* r : sc->reps) destroy_replica(r);
    delete sc;
}

uint32_t pt_scene_device_count(const PtScene* sc) { return sc ? (uint32_t)sc->reps.size() : 0u; }

int pt_scene_device_stats(const PtScene* sc, uint32_t index, PtRenderStats* out) {
    if (!sc || !out || index >= sc->reps.size()) return fail(PT_ERR_INVALID, "null argument or device slot out of range");
    *out = sc->reps[index]->stats;
    return PT_OK;
}

int pt_render_part(PtScene* sc, const PtParams* params, const PtCamera* camera, uint32_t frame_num, const PtPartition* part_in,
                   float* rgb_inout, uint64_t* ray_count_out) {
    if (!sc || !rgb_inout) return fail(PT_ERR_INVALID, "null scene/buffer");
    int rc = validate_params(params, camera);
    if (rc != PT_OK) return rc;
    PtPartition part;
    rc = normalise_partition(part_in, &part);
    if (rc != PT_OK) return rc;
    sc->prog_valid = false;  // this call reuses the scene's device images
    if (sc->reps.size() > 1) {
        if (part.part_count > 1) return fail(PT_ERR_UNSUPPORTED, "a multi-device scene partitions the image itself: pass part = NULL (or use one-device scenes with pt_render_part)");
        return for_each_replica(sc, [&](Replica* r, const PtPartition& p, uint64_t* rays) { return render_part_host(r, params, camera, frame_num, p, rgb_inout, rays); },
                                ray_count_out);
    }
    return for_each_replica(sc, [&](Replica* r, const PtPartition&, uint64_t* rays) { return render_part_host(r, params, camera, frame_num, part, rgb_inout, rays); },
                            ray_count_out);
}

int pt_render(PtScene* s, const PtParams* params, const PtCamera* camera, uint32_t frame_num, float* rgb_inout, uint64_t* ray_count_out) {
    return pt_render_part(s, params, camera, frame_num, nullptr, rgb_inout, ray_count_out);
}

int pt_render_progressive(PtScene* sc, const PtParams* params, const PtCamera* camera, uint32_t frame_num, float* rgb_out, uint8_t* rgb8_out,
                          uint64_t* ray_count_out) {
    if (!sc) return fail(PT_ERR_INVALID, "null scene");
    int rc = validate_params(params, camera);
    if (rc != PT_OK) return rc;
    if (frame_num != 0 && !(sc->prog_valid && sc->prog_w == params->width && sc->prog_h == params->height && sc->prog_next_frame == frame_num))
        return fail(PT_ERR_INVALID, "frame %u does not continue the resident accumulation (have %ux%u, next frame %u)", frame_num, sc->prog_w,
                    sc->prog_h, sc->prog_valid ? sc->prog_next_frame : 0u);
    sc->prog_valid = false;  // any failure below leaves the resident images undefined
    rc = for_each_replica(sc, [&](Replica* r, const PtPartition& p, uint64_t* rays) { return render_part_progressive(r, params, camera, frame_num, p, rgb_out, rgb8_out, rays); },
                          ray_count_out);
    if (rc != PT_OK) return rc;
    sc->prog_valid = true;
    sc->prog_w = params->width;
    sc->prog_h = params->height;
    sc->prog_next_frame = frame_num + 1;
    return PT_OK;
}

int pt_render_device(PtScene* sc, const PtParams* params, const PtCamera* camera, uint32_t frame_num, const PtPartition* part_in,
                     float* d_rgb_inout, uint64_t* d_ray_count, void* cuda_stream) {
    if (!sc || !d_rgb_inout || !d_ray_count) return fail(PT_ERR_INVALID, "null scene/buffer");
    int rc = need_single_device(sc, "pt_render_device");
    if (rc != PT_OK) return rc;
    rc = validate_params(params, camera);
    if (rc != PT_OK) return rc;
    PtPartition part;
    rc = normalise_partition(part_in, &part);
    if (rc != PT_OK) return rc;
    Replica* s = sc->reps[0];
    PT_CUDA(cudaSetDevice(s->device));
    s->stats = PtRenderStats{};
    rc = launch_update(s, params, camera, frame_num, part, d_rgb_inout, reinterpret_cast<unsigned long long*>(d_ray_count),
                       reinterpret_cast<cudaStream_t>(cuda_stream));
    sc->stats = s->stats;
    return rc;
}

int pt_srgb8_device(PtScene* sc, const float* d_rgb, uint32_t width, uint32_t height, uint8_t* d_rgb8_out, void* cuda_stream) {
    if (!sc || !d_rgb || !d_rgb8_out || width == 0 || height == 0) return fail(PT_ERR_INVALID, "null/empty argument");
    int rc = need_single_device(sc, "pt_srgb8_device");
    if (rc != PT_OK) return rc;
    Replica* s = sc->reps[0];
    PT_CUDA(cudaSetDevice(s->device));
    const size_t n = (size_t)width * height;
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((n + threads - 1) / threads, (size_t)s->sm_count * 8);
    pt::pt_srgb8_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(d_rgb, width, height, d_rgb8_out);
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_srgb8(PtScene* sc, const float* rgb, uint32_t width, uint32_t height, uint8_t* rgb8_out) {
    if (!sc || !rgb || !rgb8_out || width == 0 || height == 0) return fail(PT_ERR_INVALID, "null/empty argument");
    Replica* s = sc->reps[0];  // the output stage is 0.1 ms of work: the first device does it
    PT_CUDA(cudaSetDevice(s->device));
    const size_t n = (size_t)width * height;
    int rc = ensure_image(s, n * 3);
    if (rc != PT_OK) return rc;
    rc = ensure_rgb8(s, n * 3);
    if (rc != PT_OK) return rc;
    sc->prog_valid = false;  // this call reuses the scene's device image
    rc = serialise_begin(s, s->stream);
    if (rc != PT_OK) return rc;
    PT_CUDA(cudaMemcpyAsync(s->d_rgb, rgb, n * 3 * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((n + threads - 1) / threads, (size_t)s->sm_count * 8);
    pt::pt_srgb8_kernel<<<blocks, threads, 0, s->stream>>>(s->d_rgb, width, height, s->d_rgb8);
    PT_CUDA(cudaGetLastError());
    PT_CUDA(cudaMemcpyAsync(rgb8_out, s->d_rgb8, n * 3, cudaMemcpyDeviceToHost, s->stream));
    PT_CUDA(cudaStreamSynchronize(s->stream));
    return PT_OK;
}

int pt_scene_stats(const PtScene* sc, PtRenderStats* out) {
    if (!sc || !out) return fail(PT_ERR_INVALID, "null argument");
    *out = sc->stats;
    return PT_OK;
}

int pt_host_register(void* ptr, uint64_t bytes) {
    if (!ptr || bytes == 0) return fail(PT_ERR_INVALID, "null/empty buffer");
    PT_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return PT_OK;
}
int pt_host_unregister(void* ptr) {
    if (!ptr) return fail(PT_ERR_INVALID, "null buffer");
    PT_CUDA(cudaHostUnregister(ptr));
    return PT_OK;
}

int pt_debug_hits(PtScene* sc, const float* rays6, const float* times, uint32_t n, int32_t mode, int32_t* idx_out, float* t_out, uint32_t* flagged_out) {
    if (!sc || !rays6 || !idx_out || !t_out) return fail(PT_ERR_INVALID, "null argument");
    if (mode != 0 && mode != 1) return fail(PT_ERR_INVALID, "mode %d (0 = shipped two-stage sweep, 1 = exact test on every sphere)", mode);
    if (n == 0) return PT_OK;
    Replica* s = sc->reps[0];
    PT_CUDA(cudaSetDevice(s->device));
    float *d_rays = nullptr, *d_times = nullptr, *d_t = nullptr;
    int32_t* d_idx = nullptr;
    uint32_t* d_flagged = nullptr;
    int rc = PT_OK;
    auto step = [&](cudaError_t e, const char* what) {
        if (rc == PT_OK && e != cudaSuccess) rc = fail(PT_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    };
    step(cudaMalloc(&d_rays, (size_t)n * 6 * sizeof(float)), "cudaMalloc");
    step(cudaMalloc(&d_t, (size_t)n * sizeof(float)), "cudaMalloc");
    step(cudaMalloc(&d_idx, (size_t)n * sizeof(int32_t)), "cudaMalloc");
    if (times) step(cudaMalloc(&d_times, (size_t)n * sizeof(float)), "cudaMalloc");
    if (flagged_out) step(cudaMalloc(&d_flagged, (size_t)n * sizeof(uint32_t)), "cudaMalloc");
    if (rc == PT_OK) step(cudaMemcpyAsync(d_rays, rays6, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, s->stream), "cudaMemcpyAsync");
    if (rc == PT_OK && times) step(cudaMemcpyAsync(d_times, times, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->stream), "cudaMemcpyAsync");
    if (rc == PT_OK) {
        pt::KernelArgs a{};
        fill_scene_args(s, a);
        a.dbg_rays = d_rays;
        a.dbg_times = d_times;
        a.dbg_n = n;
        a.dbg_idx = d_idx;
        a.dbg_t = d_t;
        a.dbg_flagged = d_flagged;
        rc = serialise_begin(s, s->stream);
        const uint32_t max_ctas = (uint32_t)(s->sm_count * s->ctas_per_sm);
        if (rc == PT_OK && mode == 1) {
            const uint32_t grid = std::min<uint32_t>((n + 255u) / 256u, (uint32_t)s->sm_count * 8u);
            if (s->d_motion) pt::pt_debug_hits_exact_all<true><<<grid, 256, 0, s->stream>>>(a);
            else pt::pt_debug_hits_exact_all<false><<<grid, 256, 0, s->stream>>>(a);
        } else if (rc == PT_OK && s->regroup) {
            const uint32_t grid = std::min<uint32_t>((n + pt::kCtaThreads - 1) / pt::kCtaThreads, max_ctas);
            if (s->mma) {
                if (s->d_motion) pt::pt_debug_hits_regroup<true, true><<<grid, pt::kCtaThreads, s->smem_bytes, s->stream>>>(a);
                else pt::pt_debug_hits_regroup<false, true><<<grid, pt::kCtaThreads, s->smem_bytes, s->stream>>>(a);
            } else {
                if (s->d_motion) pt::pt_debug_hits_regroup<true, false><<<grid, pt::kCtaThreads, s->smem_bytes, s->stream>>>(a);
                else pt::pt_debug_hits_regroup<false, false><<<grid, pt::kCtaThreads, s->smem_bytes, s->stream>>>(a);
            }
        } else if (rc == PT_OK && s->resident) {
            const uint32_t batch = (uint32_t)pt::kCtaThreads * pt::kPathRows;
            const uint32_t grid = std::min<uint32_t>((n + batch - 1) / batch, (uint32_t)s->sm_count * 2u);
            if (s->const_image) {
                const size_t dbg_smem = resident_smem(s->n_blocks, true);
                if (s->d_motion) pt::pt_debug_hits_resident<true, true><<<grid, pt::kCtaThreads, dbg_smem, s->stream>>>(a, *s->h_const_image);
                else pt::pt_debug_hits_resident<true, false><<<grid, pt::kCtaThreads, dbg_smem, s->stream>>>(a, *s->h_const_image);
            } else {
                const pt::ConstImageT<false> none{};
                if (s->d_motion) pt::pt_debug_hits_resident<false, true><<<grid, pt::kCtaThreads, s->smem_bytes, s->stream>>>(a, none);
                else pt::pt_debug_hits_resident<false, false><<<grid, pt::kCtaThreads, s->smem_bytes, s->stream>>>(a, none);
            }
        } else if (rc == PT_OK) {
            const uint32_t grid = std::min<uint32_t>((n + pt::kCtaThreads - 1) / pt::kCtaThreads, max_ctas);
            if (s->mma) {
                if (s->d_motion) pt::pt_debug_hits_streamed<true, true><<<grid, pt::kCtaThreads, s->smem_bytes, s->stream>>>(a);
                else pt::pt_debug_hits_streamed<false, true><<<grid, pt::kCtaThreads, s->smem_bytes, s->stream>>>(a);
            } else {
                if (s->d_motion) pt::pt_debug_hits_streamed<true, false><<<grid, pt::kCtaThreads, s->smem_bytes, s->stream>>>(a);
                else pt::pt_debug_hits_streamed<false, false><<<grid, pt::kCtaThreads, s->smem_bytes, s->stream>>>(a);
            }
        }
        if (rc == PT_OK) step(cudaGetLastError(), "kernel launch");
        if (rc == PT_OK) rc = serialise_end(s, s->stream);
    }
    if (rc == PT_OK) step(cudaMemcpyAsync(idx_out, d_idx, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream), "cudaMemcpyAsync");
    if (rc == PT_OK) step(cudaMemcpyAsync(t_out, d_t, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->stream), "cudaMemcpyAsync");
    if (rc == PT_OK && flagged_out) step(cudaMemcpyAsync(flagged_out, d_flagged, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream), "cudaMemcpyAsync");
    if (rc == PT_OK) step(cudaStreamSynchronize(s->stream), "cudaStreamSynchronize");
    cudaFree(d_rays);
    cudaFree(d_times);
    cudaFree(d_t);
    cudaFree(d_idx);
    cudaFree(d_flagged);
    return rc;
}

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
