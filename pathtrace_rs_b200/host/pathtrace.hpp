// pathtrace.hpp — C++ host mirror of the pathtrace-rs API that sits above the C ABI (include/ptgpu.h).
//
// The reference is a Rust binary crate and this image has no Rust toolchain, so the host side that the
// north star asks for ("keeps the reference's Params/Scene/Camera/Material/Texture API so it drops in
// behind offline.rs and main.rs") is written in C++ with the same type names, method names, argument
// meaning and error behaviour (the reference panics via expect/unwrap; this mirror throws
// std::runtime_error, and the offline driver aborts with the same messages).  The Rust binding a
// maintainer would add instead is in INTEGRATION.md and rust/.
//
//   Params              src/params.rs:11-46        Camera            src/camera.rs:8-68 (new only; get_ray is device code)
//   Xoshiro256Plus      rand_xoshiro 0.6.0          Texture/Material  src/texture.rs:40-72, src/material.rs:13-39
//   Perlin (tables)     src/perlin.rs:15-51         Sphere/Hitable    src/collision/sphere.rs:7-26, hitable.rs:12-21
//   Storage             src/storage.rs:12-96        Scene             src/scene.rs:18-31,73-121  (update -> pt_render)
//   presets::from_name  src/presets.rs:13-38        offline::render_offline  src/offline.rs:16-60
//
// Everything here is scene ASSEMBLY (host work in the reference too).  All per-pixel arithmetic runs in
// libptgpu.so; there is no CPU rendering path in this library.
#pragma once
#include <cstdint>
#include <deque>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/ptgpu.h"

namespace pathtrace {

struct Vec3 {
    float x = 0, y = 0, z = 0;
    constexpr Vec3() = default;
    constexpr Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    static constexpr Vec3 splat(float v) { return Vec3(v, v, v); }
};
Vec3 operator+(Vec3 a, Vec3 b);
Vec3 operator-(Vec3 a, Vec3 b);
Vec3 operator*(Vec3 a, float s);
Vec3 operator*(float s, Vec3 a);
float dot(Vec3 a, Vec3 b);
Vec3 cross(Vec3 a, Vec3 b);
float length(Vec3 a);
Vec3 normalize(Vec3 a);

// rand_xoshiro::Xoshiro256Plus with rand_core's seed_from_u64 and rand's Standard f32 sampling
class Xoshiro256Plus {
public:
    static Xoshiro256Plus seed_from_u64(uint64_t seed);
    uint64_t next_u64();
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    float gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }  // rng.gen::<f32>()
    const uint64_t* state() const { return s_; }

private:
    uint64_t s_[4] = {0, 0, 0, 0};
};

// src/camera.rs
class Camera {
public:
    Camera() = default;
    static Camera create(Vec3 lookfrom, Vec3 lookat, Vec3 vup, float vfov, float aspect, float aperture, float focus_dist,
                         float time0, float time1);  // Camera::new
    PtCamera to_ffi() const;                          // the `pub(crate) fn to_ffi` INTEGRATION.md adds to camera.rs

private:
    Vec3 origin, lower_left_corner, horizontal, vertical, u, v, w;
    float time0 = 0, time1 = 0, lens_radius = 0;
};

// src/perlin.rs (table generation only)
class Perlin {
public:
    explicit Perlin(Xoshiro256Plus& rng);  // Perlin::new
    const PtPerlin& tables() const { return t_; }

private:
    PtPerlin t_;
};

// src/texture.rs:6-25 — width*height packed 8-bit RGB, row 0 = top (`image.to_rgb8().into_raw()`).  The lookup
// (`RgbImage::value`, texture.rs:27-36) is device code.
class RgbImage {
public:
    // RgbImage::open.  The reference decodes any format through the `image` crate; this mirror decodes binary PPM ("P6",
    // maxval 255) whatever the file is called, and throws for a missing file (the reference `unwrap`s) or another format.
    static RgbImage open(const std::string& path);
    static RgbImage from_raw(uint32_t width, uint32_t height, std::vector<uint8_t> data);
    uint32_t width() const { return width_; }
    uint32_t height() const { return height_; }
    const std::vector<uint8_t>& data() const { return data_; }

private:
    uint32_t width_ = 0, height_ = 0;
    std::vector<uint8_t> data_;
};

// src/texture.rs:40-72
struct Texture {
    enum Kind { Constant = PT_TEX_CONSTANT, Checker = PT_TEX_CHECKER, Noise = PT_TEX_NOISE, Image = PT_TEX_IMAGE } kind = Constant;
    Vec3 color;                   // Constant
    const Texture* odd = nullptr;  // Checker
    const Texture* even = nullptr;
    const Perlin* noise = nullptr;  // Noise
    float scale = 0;
    const RgbImage* image = nullptr;  // Image
};
namespace texture {
Texture constant(Vec3 color);
Texture checker(const Texture* odd, const Texture* even);
Texture noise(const Perlin* noise, float scale);
Texture rgb_image(const RgbImage* image);
}  // namespace texture

// src/material.rs:13-39
struct Material {
    enum Kind { Lambertian = PT_MAT_LAMBERTIAN, Metal = PT_MAT_METAL, Dielectric = PT_MAT_DIELECTRIC, DiffuseLight = PT_MAT_DIFFUSE_LIGHT } kind = Lambertian;
    const Texture* albedo_tex = nullptr;  // Lambertian.albedo / DiffuseLight.emit
    Vec3 albedo;                          // Metal
    float fuzz = 0;
    float ref_idx = 0;
};
namespace material {
Material lambertian(const Texture* albedo);
Material metal(Vec3 albedo, float fuzz);
Material dielectric(float ref_idx);
Material diffuse_light(const Texture* emit);
}  // namespace material

// src/collision/sphere.rs:7-26
class Sphere {
public:
    Sphere(Vec3 centre, float radius) : centre_(centre), radius_(radius) {}
    Vec3 centre() const { return centre_; }
    float radius() const { return radius_; }

private:
    Vec3 centre_;
    float radius_;
};

// src/collision/moving_sphere.rs:7-36
class MovingSphere {
public:
    MovingSphere(Vec3 centre0, Vec3 centre1, float time0, float time1, float radius)  // MovingSphere::new
        : centre0_(centre0), centre1_(centre1), time0_(time0), time1_(time1), radius_(radius) {}
    Vec3 centre0() const { return centre0_; }
    Vec3 centre1() const { return centre1_; }
    float time0() const { return time0_; }
    float time1() const { return time1_; }
    float radius() const { return radius_; }

private:
    Vec3 centre0_, centre1_;
    float time0_, time1_, radius_;
};

// src/collision/hitable.rs:12-21 — only the arms the GPU path accepts are constructible; anything else
// is represented as `Unsupported` so the flattener can reject it the way spheres_soa.rs:49-51 panics.
struct Hitable {
    enum Kind { SphereKind, MovingSphereKind, Unsupported } kind = SphereKind;
    const Sphere* sphere = nullptr;
    const Material* material = nullptr;
    std::string what;  // name of the unsupported variant
    const MovingSphere* moving_sphere = nullptr;
    static Hitable make_sphere(const Sphere* s, const Material* m) { return Hitable{SphereKind, s, m, {}, nullptr}; }
    static Hitable make_moving_sphere(const MovingSphere* s, const Material* m) { return Hitable{MovingSphereKind, nullptr, m, {}, s}; }
    static Hitable unsupported(std::string name) { return Hitable{Unsupported, nullptr, nullptr, std::move(name), nullptr}; }
};

// src/storage.rs:12-96 — arenas (pointer-stable deques) + the one Perlin table
class Storage {
public:
    explicit Storage(Xoshiro256Plus& rng) : perlin_noise(rng) {}  // Storage::new: draws the Perlin tables first
    const Texture* alloc_texture(Texture t) { textures_.push_back(t); return &textures_.back(); }
    const Material* alloc_material(Material m) { materials_.push_back(m); return &materials_.back(); }
    const Sphere* alloc_sphere(Sphere s) { spheres_.push_back(s); return &spheres_.back(); }
    const MovingSphere* alloc_moving_sphere(MovingSphere s) { moving_spheres_.push_back(s); return &moving_spheres_.back(); }
    const RgbImage* alloc_image(RgbImage i) { images_.push_back(std::move(i)); return &images_.back(); }  // storage.rs: image_arena
    Perlin perlin_noise;

private:
    std::deque<Texture> textures_;
    std::deque<Material> materials_;
    std::deque<Sphere> spheres_;
    std::deque<MovingSphere> moving_spheres_;
    std::deque<RgbImage> images_;
};

struct Params;

// src/scene.rs:18-31,73-121.  Owns the device copy of the flattened world.
class Scene {
public:
    Scene(const std::vector<Hitable>& world, std::optional<Vec3> sky, int device = 0);  // Scene::new + flatten + upload
    // the same scene replicated on several GPUs: `update` stays ONE call and the library splits the image by interleaved
    // row tiles, one host thread per GPU (pt_scene_create_multi); `options` may be null
    Scene(const std::vector<Hitable>& world, std::optional<Vec3> sky, const std::vector<int>& devices, const PtOptions* options = nullptr);
    ~Scene();
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;

    // Scene::update — same arguments, same return (ray count).  buffer: width*height (r,g,b) triples, bottom-up.
    size_t update(const Params& params, const Camera& camera, uint32_t frame_num, float* buffer, size_t buffer_len_pixels) const;
    size_t update_part(const Params& params, const Camera& camera, uint32_t frame_num, const PtPartition& part, float* buffer,
                       size_t buffer_len_pixels) const;
    // the windowed worker loop (glium_window.rs:96-131): frame_num = 0, 1, 2, ... with the accumulation buffer resident on
    // the device(s); `buffer` (may be null) receives the running mean, `rgb8` (may be null) the top-down sRGB8 picture
    size_t update_progressive(const Params& params, const Camera& camera, uint32_t frame_num, float* buffer, uint8_t* rgb8) const;
    PtScene* handle() const { return scene_; }
    size_t len() const { return n_spheres_; }

private:
    PtScene* scene_ = nullptr;
    size_t n_spheres_ = 0;
};

// src/params.rs:11-46
struct Params {
    uint32_t width = 1280, height = 720, samples = 4, max_depth = 10;  // main.rs:78-85 defaults
    bool random_seed = false, use_bvh = false;
    uint64_t seed_salt = 0;  // entropy for random_seed (rand::random() in the reference)
    Xoshiro256Plus new_rng() const;
    std::unique_ptr<Scene> new_scene(Xoshiro256Plus& rng, const Storage& storage, std::vector<Hitable> hitables,
                                     std::optional<Vec3> sky, int device = 0) const;
    std::unique_ptr<Scene> new_scene(Xoshiro256Plus& rng, const Storage& storage, std::vector<Hitable> hitables,
                                     std::optional<Vec3> sky, const std::vector<int>& devices, const PtOptions* options = nullptr) const;
    PtParams to_ffi() const;
};

namespace presets {
using Preset = std::tuple<std::vector<Hitable>, Camera, std::optional<Vec3>>;
// presets::from_name — sphere-only presets: random (moving spheres), random_spheres, small, two_perlin_spheres, smallpt, earth,
// final, plus the synthetic stress100k.  Other names -> nullopt ("unrecognised preset").
// `earth` opens "media/earthmap.jpg" like presets.rs:583 (the asset is not part of the reference tree); the environment
// variable PATHTRACE_EARTHMAP names another file.
std::optional<Preset> from_name(const std::string& name, const Params& params, Xoshiro256Plus& rng, Storage& storage, bool quiet = false);
}  // namespace presets

namespace offline {
// src/offline.rs:16-60.  Returns (elapsed seconds, ray count); writes `output_png` unless empty.
std::pair<double, size_t> render_offline(const std::string& preset, const Params& params, const std::string& output_png = "output.png",
                                         int device = 0);
// the same over a list of GPUs (`-G 0-7`), and `frames` > 1 = the progressive loop of the windowed mode run headless
// (frame_num = 0 .. frames-1 accumulated on the device, glium_window.rs:98-131)
std::pair<double, size_t> render_offline(const std::string& preset, const Params& params, const std::string& output_png,
                                         const std::vector<int>& devices, uint32_t frames = 1);
}  // namespace offline

void linear_to_srgb8_image(const Scene& scene, const float* rgb, uint32_t width, uint32_t height, std::vector<uint8_t>& out);
bool write_png_rgb8(const std::string& path, const uint8_t* rgb, uint32_t width, uint32_t height);

}  // namespace pathtrace
