/* ptgpu.h — C ABI of libptgpu.so: the B200 (sm_100a) implementation of pathtrace-rs's per-pixel
 * path-tracing loop (`Scene::update` and everything it calls).
 *
 * This header is the drop-in boundary.  Every entry point names the reference interface it replaces
 * (paths relative to the pathtrace-rs source tree).  Plain C types only: pointers, sizes, PODs.
 * There is no CPU fallback: every compute entry point fails with PT_ERR_NO_DEVICE when no CUDA
 * device is usable.
 *
 * Image convention (src/scene.rs:94-95,108; src/camera.rs:41-44): rgb buffers are `width*height`
 * packed (r,g,b) f32 triples, row-major, BOTTOM-UP (row 0 is the bottom of the picture), exactly the
 * `&mut [(f32,f32,f32)]` that `Scene::update` blends into.
 */
#ifndef PTGPU_H
#define PTGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PT_ABI_VERSION 4

/* ---- status codes (the Rust shim `expect`s on non-zero, matching the reference's panic-on-error
 *      convention: src/offline.rs:10,32,59) ---- */
enum {
    PT_OK = 0,
    PT_ERR_INVALID = 1,      /* null pointer, bad index, zero-sized image ... */
    PT_ERR_UNSUPPORTED = 2,  /* e.g. use_bvh, or a material/texture kind outside the path */
    PT_ERR_NO_DEVICE = 3,    /* no CUDA device / wrong architecture (needs sm_100) */
    PT_ERR_CUDA = 4,         /* a CUDA runtime call failed; see pt_last_error() */
    PT_ERR_TOO_LARGE = 5
};

/* ---- Params: src/params.rs:11-18 (`pub struct Params`) ---- */
typedef struct PtParams {
    uint32_t width;
    uint32_t height;
    uint32_t samples;
    uint32_t max_depth;
    uint8_t random_seed; /* bool: per-pixel seeds from host entropy instead of (x,y,frame): src/scene.rs:96-102 */
    uint8_t use_bvh;     /* bool: must be 0 — the GPU path is the flat list (src/params.rs:36-43 List arm) */
    uint8_t _pad[6];
    uint64_t seed_salt;  /* only read when random_seed != 0: the host's `rand::random()` entropy */
} PtParams;

/* ---- Camera: src/camera.rs:8-19 (10 private fields, built on the host by Camera::new :22-54) ---- */
typedef struct PtCamera {
    float origin[3];
    float lower_left_corner[3];
    float horizontal[3];
    float vertical[3];
    float u[3];
    float v[3];
    float w[3]; /* unused by get_ray (src/camera.rs:56-68); carried for completeness */
    float time0;
    float time1;
    float lens_radius;
} PtCamera;

/* ---- Texture: src/texture.rs:40-55 (arena references become indices into PtSceneDesc.textures / .images) ---- */
enum { PT_TEX_CONSTANT = 0, PT_TEX_CHECKER = 1, PT_TEX_NOISE = 2, PT_TEX_IMAGE = 3 };
typedef struct PtTexture {
    int32_t kind;
    float color[3]; /* Constant */
    int32_t odd;    /* Checker: texture index */
    int32_t even;   /* Checker: texture index */
    float scale;    /* Noise */
    int32_t image;  /* Image: index into PtSceneDesc.images (ignored by the other kinds) */
} PtTexture;

/* ---- RgbImage: src/texture.rs:6-37 (`image.to_rgb8().into_raw()`: width*height packed 8-bit RGB, row 0 = top).
 * Looked up by `RgbImage::value(u, v)` (texture.rs:27-36) with the sphere (u, v) of `get_sphere_uv`
 * (src/material.rs:41-49).  (u, v) are computed only when the Image texture is the material's own albedo / emit
 * texture (material.rs:169-180); an Image nested inside a Checker is sampled at (0, 0), as in the reference. ---- */
typedef struct PtImage {
    uint32_t width;
    uint32_t height;
    const uint8_t* data; /* width*height*3 bytes; copied by pt_scene_create */
} PtImage;

/* ---- Material: src/material.rs:13-19 ---- */
enum { PT_MAT_LAMBERTIAN = 0, PT_MAT_METAL = 1, PT_MAT_DIELECTRIC = 2, PT_MAT_DIFFUSE_LIGHT = 3 };
typedef struct PtMaterial {
    int32_t kind;
    int32_t texture; /* Lambertian.albedo / DiffuseLight.emit: texture index; -1 otherwise */
    float albedo[3]; /* Metal */
    float fuzz;      /* Metal */
    float ref_idx;   /* Dielectric */
    int32_t _pad;
} PtMaterial;

/* ---- Perlin: src/perlin.rs:7-12 (tables generated on the host by Perlin::new :44-51) ---- */
typedef struct PtPerlin {
    float randvec[256][3];
    uint32_t perm_x[256];
    uint32_t perm_y[256];
    uint32_t perm_z[256];
} PtPerlin;

/* ---- MovingSphere: src/collision/moving_sphere.rs:7-31 (`MovingSphere::new(centre0, centre1, time0, time1, radius)`).
 * One record per sphere when PtSceneDesc.motion != NULL; centre0 and radius are the sphere's centre_x/y/z and radius.
 * centre(t) = centre0 + ((t - time0) / (time1 - time0)) * (centre1 - centre0), evaluated at the ray's time
 * (src/camera.rs:59).  The camera's shutter interval must lie inside [time0, time1] of every moving sphere. ---- */
typedef struct PtMotion {
    float centre1[3];
    float time0;
    float time1;
    uint32_t moving; /* 0: Hitable::Sphere (the other fields are ignored); 1: Hitable::MovingSphere */
} PtMotion;

/* ---- Scene: src/scene.rs:18-22 (`world` flattened) + src/collision/spheres_soa.rs:12-23 (SoA arrays).
 * The flattener walks Hitable::List and must reject anything that is not Hitable::Sphere or Hitable::MovingSphere
 * (mirrors the panic at spheres_soa.rs:49-51). `radius` keeps its sign (hollow spheres: presets.rs:265). */
typedef struct PtSceneDesc {
    uint32_t struct_size; /* = sizeof(PtSceneDesc) */
    uint32_t n_spheres;
    const float* centre_x;
    const float* centre_y;
    const float* centre_z;
    const float* radius;
    const int32_t* material_index; /* per sphere, into materials[] */
    uint32_t n_materials;
    uint32_t n_textures;
    const PtMaterial* materials;
    const PtTexture* textures;
    const PtPerlin* perlin; /* may be NULL when no Noise texture is referenced */
    uint32_t has_sky;       /* Scene.sky: Option<Vec3>  (src/scene.rs:20,39-47) */
    float sky[3];
    const PtMotion* motion; /* per sphere, or NULL when the scene has no Hitable::MovingSphere (src/collision/hitable.rs:17) */
    uint32_t n_images;      /* Storage.image_arena entries referenced by Image textures (src/storage.rs) */
    uint32_t _pad;
    const PtImage* images;  /* may be NULL when n_images == 0 */
} PtSceneDesc;

/* ---- partition of one image over several GPUs / calls: interleaved row tiles.
 * Tile k = rows [k*tile_rows, (k+1)*tile_rows); this part owns tile k iff k % part_count == part_index.
 * Pixel seeds depend only on (x, y, frame) (src/scene.rs:99-101), so the image is identical for any split. */
typedef struct PtPartition {
    uint32_t tile_rows;  /* 0 -> default (4) */
    uint32_t part_index;
    uint32_t part_count; /* 0 or 1 -> whole image */
    uint32_t _pad;
} PtPartition;

typedef struct PtDeviceInfo {
    char name[64];
    int32_t sm_count;
    int32_t cc_major;
    int32_t cc_minor;
    int32_t sm_clock_khz;       /* max SM clock */
    double fp32_fma_peak_flops; /* sm_count * 128 lanes * 2 flop * max SM clock */
    uint64_t global_mem_bytes;
} PtDeviceInfo;

/* per-render statistics of the last pt_render* call on a scene (measurement) */
typedef struct PtRenderStats {
    double kernel_ms;   /* CUDA-event time of the megakernel launch(es) (0 for pt_render_device: caller times) */
    double h2d_ms;      /* host->device copy of the previous frame (0 when frame_num == 0) */
    double d2h_ms;      /* device->host copy of the result */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t ray_count;
    uint32_t n_spheres;
    uint32_t kernel_launches;
    uint32_t grid_ctas;
    uint32_t cta_threads;
    uint32_t smem_bytes;
    uint32_t resident; /* 0: sphere image streamed through L2 in tiles; 1: resident in shared memory, pre-filter in packed FP32;
                          2: resident, pre-filter's dot products on the tensor path (PtOptions.resident_kernel 5); 3: streamed, tensor path */
    uint64_t warp_sweeps; /* warp-level sweeps of 32 ray slots performed: ray_count / (32 * warp_sweeps) = lane efficiency of the sweep */
} PtRenderStats;

/* ---- explicit launch options (replace the process-global environment hooks of ABI v3).  Every field has a
 * "library decides" value, and a NULL PtOptions* means all of them.  The options never change an image: sphere order,
 * kernel flavour and chunking are internal (hits, equal-t ties and every pixel's RNG stream are the reference's). ---- */
typedef struct PtOptions {
    uint32_t struct_size;             /* = sizeof(PtOptions) */
    int32_t force_stream_tile_blocks; /* 0: resident kernel whenever the scene fits in shared memory; n > 0: always the
                                         L2-streamed kernel with n-block tiles (1 block = 4 spheres) */
    int32_t stream_ctas;              /* 0: default (2); 1..4: CTAs per SM of the streamed kernel */
    int32_t chunk_samples;            /* 0: automatic; -1: whole pixels (the reference's work unit, scene.rs:90-93);
                                         n > 0: samples per work-queue ticket, rounded up to a power of two */
    int32_t spatial_order;            /* -1: automatic; 0: store spheres in the caller's order; 1: Morton order;
                                         2: large spheres first, then Morton order */
    uint32_t tile_rows;               /* multi-device scenes: rows per interleaved row tile (0 -> 4) */
    uint32_t resident_kernel;         /* which kernel renders a scene that fits in shared memory.  0 automatic: 5 where the
                                         scene suits it (>= 128 spheres of a size comparable to the scene's extent), else 4;
                                         a suitable scene whose tensor-path image no longer fits twice per SM (about 2 000
                                         spheres and up) is streamed through L2 on the tensor path instead (measured faster).
                                         4 one path per lane + CTA regroup, pre-filter in packed FP32; 5 the same kernel with
                                         the pre-filter's dot products on the tensor path (mma.sync f16 split operands,
                                         pt_sweep_mma.cuh; a render whose camera lies outside the scene's extent falls back
                                         to 4); 2 / 3 two paths per lane with the sphere pairs as uniform operands from a
                                         kernel-parameter image / from shared memory; 1 wavefront form (path pool +
                                         per-material queues).  1 and 2 need a scene of at most 2048 spheres and fall back
                                         to 3 beyond.  4 / 5 also choose the pre-filter of the L2-streamed kernel (scenes beyond
                                         shared memory).  All produce the same image. */
} PtOptions;

typedef struct PtScene PtScene; /* opaque: device copies of one scene on one or several GPUs */

int pt_abi_version(void);
/* sizeof() of the ABI structs as this library was compiled, so a binding can assert its own layout:
 * which = 0 PtParams, 1 PtCamera, 2 PtTexture, 3 PtMaterial, 4 PtPerlin, 5 PtSceneDesc, 6 PtPartition,
 * 7 PtDeviceInfo, 8 PtRenderStats, 9 PtMotion, 10 PtImage, 11 PtOptions; anything else -> 0. */
uint32_t pt_abi_struct_size(int which);
const char* pt_last_error(void); /* thread-local message of the last failing call */
int pt_device_count(void);
int pt_device_info(int device, PtDeviceInfo* out);

/* Replaces `Params::new_scene` (src/params.rs:29-46, List arm) + `SpheresSoA::new`
 * (src/collision/spheres_soa.rs:26-74): validates and uploads the flat scene to `device`. */
int pt_scene_create(const PtSceneDesc* desc, int device, PtScene** out);
void pt_scene_destroy(PtScene* scene);

/* Same, replicated on `n_devices` GPUs of this host (the scene is small and read-only: SURVEY §8e).  The ordinary entry
 * points then split every `Scene::update` over the devices by interleaved row tiles — tile k of `options->tile_rows`
 * rows belongs to device k mod n_devices — with one host thread per GPU inside the call and every GPU copying its own
 * rows straight from / into the caller's buffer: the caller still makes ONE call (src/offline.rs:29,
 * src/glium_window.rs:102) and there is no collective in the data path.  Pixel seeds depend only on (x, y, frame)
 * (src/scene.rs:99-101), so the image and the ray count are identical for any device list.
 * `options` may be NULL.  pt_render_device / pt_srgb8_device take device pointers and therefore need a one-device scene.
 * A device may be listed more than once: every entry is an independent replica with its own stream and buffers (no use in
 * production, but it lets a one-GPU machine exercise the whole fan-out). */
int pt_scene_create_multi(const PtSceneDesc* desc, const int* devices, uint32_t n_devices, const PtOptions* options,
                          PtScene** out);
uint32_t pt_scene_device_count(const PtScene* scene);
/* statistics of the last render on device slot `index` (0 .. pt_scene_device_count-1): per-GPU kernel time and bytes,
 * what a caller needs to see the tail of the interleaved partition */
int pt_scene_device_stats(const PtScene* scene, uint32_t index, PtRenderStats* out);

/* Replaces `Scene::update(&self, &Params, &Camera, frame_num, &mut [(f32,f32,f32)]) -> usize`
 * (src/scene.rs:73-121; call sites src/offline.rs:29, src/glium_window.rs:102).
 * rgb_inout: HOST buffer, width*height*3 f32; read when frame_num > 0 (running mean, scene.rs:86-87,
 * 113-116), always written.  *ray_count_out = number of ray_trace calls (scene.rs:57,118-120). */
int pt_render(PtScene* scene, const PtParams* params, const PtCamera* camera, uint32_t frame_num,
              float* rgb_inout, uint64_t* ray_count_out);

/* Same, but only the rows of `part` are rendered, read and written (multi-GPU row-tile mode:
 * one call per GPU into the same host image). part == NULL -> whole image. */
int pt_render_part(PtScene* scene, const PtParams* params, const PtCamera* camera, uint32_t frame_num,
                   const PtPartition* part, float* rgb_inout, uint64_t* ray_count_out);

/* Device-resident variant (progressive / windowed use, src/glium_window.rs:98-131: the buffer stays
 * on the GPU across frames).  d_rgb_inout: DEVICE buffer width*height*3 f32 on the scene's device;
 * d_ray_count: DEVICE u64, overwritten with this call's ray count.  Asynchronous on `cuda_stream`
 * (a cudaStream_t; NULL = default stream); no host synchronisation is performed.
 * At most one render is in flight per PtScene: its ticket counter and pixel-state table are per scene, so every launch —
 * through this or any other entry point, on whatever stream — first waits (cudaStreamWaitEvent) for the scene's previous
 * launch.  Calls on one scene must still come from one host thread at a time. */
int pt_render_device(PtScene* scene, const PtParams* params, const PtCamera* camera, uint32_t frame_num,
                     const PtPartition* part, float* d_rgb_inout, uint64_t* d_ray_count, void* cuda_stream);

/* Progressive accumulation with the image RESIDENT on the device between calls — the worker loop of the windowed mode,
 * src/glium_window.rs:96-131 (`ray_count += scene.update(&params, &camera, frame_num, &mut rgb_buffer); frame_num += 1`),
 * without the per-frame upload/download of pt_render.  frame_num == 0 (re)starts the accumulation (scene.rs:86-87:
 * mix_prev = 0); frame_num > 0 blends into the image the previous call left on the device and must follow it
 * (same width/height, frame_num == previous + 1), otherwise PT_ERR_INVALID.  Either output may be NULL:
 *   rgb_out   width*height*3 f32, bottom-up  — the accumulation buffer itself (what Scene::update leaves in `buffer`)
 *   rgb8_out  width*height*3 u8, top-down sRGB — what the window uploads / offline.rs:43-51 writes                    */
int pt_render_progressive(PtScene* scene, const PtParams* params, const PtCamera* camera, uint32_t frame_num, float* rgb_out,
                          uint8_t* rgb8_out, uint64_t* ray_count_out);

/* Output stage of `render_offline` (src/offline.rs:43-51 + src/math.rs:36-48): rows flipped to
 * top-down, linear -> sRGB (1.055*x^0.41666666-0.055, clamped, *255.99 truncated), packed RGB8.
 * Host-buffer and device-buffer forms. */
int pt_srgb8(PtScene* scene, const float* rgb, uint32_t width, uint32_t height, uint8_t* rgb8_out);
int pt_srgb8_device(PtScene* scene, const float* d_rgb, uint32_t width, uint32_t height, uint8_t* d_rgb8_out,
                    void* cuda_stream);

int pt_scene_stats(const PtScene* scene, PtRenderStats* out);

/* Pin / unpin a caller-owned host buffer (cudaHostRegister).  pt_render accepts any host memory; a pageable
 * `Vec<(f32,f32,f32)>` (src/offline.rs:25) is copied through the driver's staging buffers, a registered one by DMA at full
 * PCIe rate and concurrently from all devices of a multi-device scene.  Worth it for the progressive loop
 * (src/glium_window.rs:96-131: the same buffer every frame) and for 4K images; the library never registers behind the
 * caller's back because it cannot know how long the buffer lives. */
int pt_host_register(void* ptr, uint64_t bytes);
int pt_host_unregister(void* ptr);

/* Per-ray entry of the sphere sweep — the GPU twin of `SpheresSoA::hit` / `HitableList::ray_hit` as the reference's own
 * benches call them (src/collision/spheres_soa.rs:464-485, src/bench.rs:17-26: one unit-direction ray against the whole
 * list, t in (0.001, f32::MAX)).  Runs the SHIPPED two-stage sweep of the render kernel (packed pre-filter over all
 * spheres, exact re-test of the flagged ones) for `n` caller-supplied rays on the scene's first device:
 *   rays6  n x (ox, oy, oz, dx, dy, dz), |d| = 1 as everywhere on the path;  times: n ray times for
 *   Hitable::MovingSphere (src/camera.rs:59), NULL = 0;
 *   idx_out[i] = position in the CALLER's sphere list of the nearest hit, -1 on a miss;  t_out[i] = its t (FLT_MAX on a miss);
 *   flagged_out (may be NULL): number of spheres the pre-filter passed on to the exact test for ray i.
 * mode 0 = the shipped sweep; mode 1 = the exact test on EVERY sphere with no pre-filter (the differential check that
 * the pre-filter never drops a sphere the exact expression accepts). */
int pt_debug_hits(PtScene* scene, const float* rays6, const float* times, uint32_t n, int32_t mode, int32_t* idx_out,
                  float* t_out, uint32_t* flagged_out);

/* Rows of an image of `height` rows that `part` owns, ascending (the order the device buffer is walked).
 * Writes at most `cap` row indices to rows_out (may be NULL) and returns the total count: what a
 * multi-GPU caller needs to gather per-GPU results. */
uint32_t pt_partition_rows(const PtPartition* part, uint32_t height, uint32_t* rows_out, uint32_t cap);

/* Diagnostic, host only (no GPU needed): the order in which pt_scene_create* stores the spheres of `desc` on the device
 * under `options` (may be NULL).
 * order_out[j] = position in the caller's list of the sphere stored at j (at most `cap` entries are written; order_out
 * may be NULL).  Returns the mode: 0 = the caller's order (small scenes, and scenes that exceed shared memory), 1 = Morton
 * order of the centres, 2 = spheres much larger than the median first, then Morton order.  The order is internal: nearest
 * hits and equal-t ties (first in the caller's list: src/collision/spheres_soa.rs:126, hitable_list.rs:49-54) do not
 * depend on it. */
uint32_t pt_scene_storage_order(const PtSceneDesc* desc, const PtOptions* options, uint32_t* order_out, uint32_t cap);

/* Diagnostic, host only (no GPU needed): the sphere operand of the tensor-path pre-filter (pt_sweep_mma.cuh) as
 * pt_scene_create* builds it for `desc` under `options`, before the shuffle into MMA fragments.
 * rows_out: per STORED sphere (padding spheres included, at most `cap` of them) 16 f16 bit patterns = the K halves
 * [S_hi(5) | S_hi(5) | S_lo(5) | 0] of S = [sigma (c - t), s, K'/s]; scale_out[7] = sigma, s, 1/s, max |o - t|^2, tx, ty, tz;
 * order_out[j] = position in the caller's list of stored sphere j (n_spheres entries).  Any of the three may be NULL.
 * Returns the number of stored spheres (a multiple of 16), or 0 when the scene does not take the tensor path (fewer than
 * 128 spheres, or spheres tiny against the scene's extent).  tests/test_host_and_abi.py evaluates the filter from these
 * halves in numpy and checks it against the reference's exact expression. */
uint32_t pt_scene_mma_operand(const PtSceneDesc* desc, const PtOptions* options, uint16_t* rows_out, uint32_t cap, float* scale_out,
                              uint32_t* order_out);

/* Measurement helper: sustained FP32 FFMA throughput of `device` (flop/s) from a pure-FMA kernel,
 * so bench.py can print the measured ceiling beside the nominal sm_count*128*2*clock figure. */
int pt_probe_fp32_peak(int device, double* flops_out);

#ifdef __cplusplus
}
#endif
#endif /* PTGPU_H */
