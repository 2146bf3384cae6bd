// pathtrace.cpp — implementation of the C++ host mirror (see pathtrace.hpp) + its C interface for ctypes.
#include "pathtrace.hpp"

#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <unordered_map>

namespace pathtrace {

// ---- glam-like Vec3 (unfused; this library is built with -ffp-contract=off) ----
Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
Vec3 operator*(Vec3 a, float s) { return Vec3(a.x * s, a.y * s, a.z * s); }
Vec3 operator*(float s, Vec3 a) { return Vec3(s * a.x, s * a.y, s * a.z); }
float dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
Vec3 cross(Vec3 a, Vec3 b) { return Vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
float length(Vec3 a) { return std::sqrt(dot(a, a)); }
Vec3 normalize(Vec3 a) { return a * (1.0f / length(a)); }

// ---- RNG ----
Xoshiro256Plus Xoshiro256Plus::seed_from_u64(uint64_t state) {
    Xoshiro256Plus r;
    for (uint64_t& word : r.s_) {  // SplitMix64, one output per state word
        state += 0x9e3779b97f4a7c15ULL;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        word = z ^ (z >> 31);
    }
    return r;
}
uint64_t Xoshiro256Plus::next_u64() {
    const uint64_t out = s_[0] + s_[3];
    const uint64_t t = s_[1] << 17;
    s_[2] ^= s_[0];
    s_[3] ^= s_[1];
    s_[1] ^= s_[2];
    s_[0] ^= s_[3];
    s_[2] ^= t;
    s_[3] = (s_[3] << 45) | (s_[3] >> 19);
    return out;
}

// ---- Camera::new (src/camera.rs:22-54) ----
Camera Camera::create(Vec3 lookfrom, Vec3 lookat, Vec3 vup, float vfov, float aspect, float aperture, float focus_dist,
                      float time0, float time1) {
    const float theta = vfov * 3.14159265358979323846f / 180.0f;
    const float half_height = std::tan(theta * 0.5f);
    const float half_width = aspect * half_height;
    Camera c;
    c.w = normalize(lookfrom - lookat);
    c.u = normalize(cross(vup, c.w));
    c.v = cross(c.w, c.u);
    c.origin = lookfrom;
    c.lower_left_corner = lookfrom - half_width * focus_dist * c.u - half_height * focus_dist * c.v - focus_dist * c.w;
    c.horizontal = 2.0f * half_width * focus_dist * c.u;
    c.vertical = 2.0f * half_height * focus_dist * c.v;
    c.time0 = time0;
    c.time1 = time1;
    c.lens_radius = aperture * 0.5f;
    return c;
}
PtCamera Camera::to_ffi() const {
    PtCamera f{};
    auto put = [](float* dst, Vec3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; };
    put(f.origin, origin);
    put(f.lower_left_corner, lower_left_corner);
    put(f.horizontal, horizontal);
    put(f.vertical, vertical);
    put(f.u, u);
    put(f.v, v);
    put(f.w, w);
    f.time0 = time0;
    f.time1 = time1;
    f.lens_radius = lens_radius;
    return f;
}

// ---- Perlin::new (src/perlin.rs:15-51): randvec, then perm_x, perm_y, perm_z ----
Perlin::Perlin(Xoshiro256Plus& rng) {
    for (auto& rv : t_.randvec) {
        const float a = -1.0f + 2.0f * rng.gen_f32();
        const float b = -1.0f + 2.0f * rng.gen_f32();
        const float c = -1.0f + 2.0f * rng.gen_f32();
        const Vec3 n = normalize(Vec3(a, b, c));
        rv[0] = n.x; rv[1] = n.y; rv[2] = n.z;
    }
    for (uint32_t* perm : {t_.perm_x, t_.perm_y, t_.perm_z}) {
        for (uint32_t i = 0; i < 256; ++i) perm[i] = i;
        for (int i = 255; i >= 0; --i) {  // perlin.rs:28-33
            const size_t target = (size_t)std::floor(rng.gen_f32() * (float)(i + 1));
            std::swap(perm[i], perm[target]);
        }
    }
}

namespace texture {
Texture constant(Vec3 color) { Texture t; t.kind = Texture::Constant; t.color = color; return t; }
Texture checker(const Texture* odd, const Texture* even) { Texture t; t.kind = Texture::Checker; t.odd = odd; t.even = even; return t; }
Texture noise(const Perlin* n, float scale) { Texture t; t.kind = Texture::Noise; t.noise = n; t.scale = scale; return t; }
Texture rgb_image(const RgbImage* image) { Texture t; t.kind = Texture::Image; t.image = image; return t; }
}  // namespace texture

// ---- RgbImage (src/texture.rs:12-25) ----
RgbImage RgbImage::from_raw(uint32_t width, uint32_t height, std::vector<uint8_t> data) {
    if (width == 0 || height == 0 || data.size() != (size_t)width * height * 3) throw std::runtime_error("RgbImage: data length != width*height*3");
    RgbImage im;
    im.width_ = width;
    im.height_ = height;
    im.data_ = std::move(data);
    return im;
}
RgbImage RgbImage::open(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("RgbImage::open: cannot open '" + path + "'");  // image::open(path).unwrap()
    std::vector<uint8_t> bytes;
    uint8_t chunk[65536];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0) bytes.insert(bytes.end(), chunk, chunk + got);
    std::fclose(f);
    // binary PPM: "P6" <ws> width <ws> height <ws> maxval <single ws> pixels; '#' starts a comment up to end of line
    size_t pos = 0;
    auto token = [&]() -> std::string {
        for (;;) {
            while (pos < bytes.size() && std::isspace(bytes[pos])) ++pos;
            if (pos < bytes.size() && bytes[pos] == '#') { while (pos < bytes.size() && bytes[pos] != '\n') ++pos; continue; }
            break;
        }
        std::string t;
        while (pos < bytes.size() && !std::isspace(bytes[pos])) t.push_back((char)bytes[pos++]);
        return t;
    };
    if (token() != "P6") throw std::runtime_error("RgbImage::open: '" + path + "' is not a binary PPM (the only format this mirror decodes; the reference uses the image crate)");
    const long w = std::atol(token().c_str()), h = std::atol(token().c_str()), maxval = std::atol(token().c_str());
    pos += 1;  // the single whitespace byte after maxval
    if (w <= 0 || h <= 0 || maxval != 255 || bytes.size() < pos + (size_t)w * h * 3) throw std::runtime_error("RgbImage::open: '" + path + "': bad or truncated PPM");
    return from_raw((uint32_t)w, (uint32_t)h, std::vector<uint8_t>(bytes.begin() + pos, bytes.begin() + pos + (size_t)w * h * 3));
}
namespace material {
Material lambertian(const Texture* albedo) { Material m; m.kind = Material::Lambertian; m.albedo_tex = albedo; return m; }
Material metal(Vec3 albedo, float fuzz) { Material m; m.kind = Material::Metal; m.albedo = albedo; m.fuzz = fuzz; return m; }
Material dielectric(float ref_idx) { Material m; m.kind = Material::Dielectric; m.ref_idx = ref_idx; return m; }
Material diffuse_light(const Texture* emit) { Material m; m.kind = Material::DiffuseLight; m.albedo_tex = emit; return m; }
}  // namespace material

// ---- Scene: flatten Hitable list -> PtSceneDesc -> pt_scene_create ----
namespace {
struct Flattener {
    std::vector<float> cx, cy, cz, radius;
    std::vector<PtMotion> motion;  // one per sphere; handed to the library only when some hitable moves
    bool any_moving = false;
    std::vector<int32_t> material_index;
    std::vector<PtMaterial> materials;
    std::vector<PtTexture> textures;
    std::unordered_map<const Material*, int32_t> mat_ids;
    std::unordered_map<const Texture*, int32_t> tex_ids;
    std::vector<PtImage> images;
    std::unordered_map<const RgbImage*, int32_t> image_ids;
    const Perlin* perlin = nullptr;

    int32_t image_id(const RgbImage* im) {
        if (!im) throw std::runtime_error("null image reference");
        auto it = image_ids.find(im);
        if (it != image_ids.end()) return it->second;
        images.push_back(PtImage{im->width(), im->height(), im->data().data()});
        const int32_t id = (int32_t)images.size() - 1;
        image_ids.emplace(im, id);
        return id;
    }

    int32_t texture_id(const Texture* t) {
        if (!t) throw std::runtime_error("null texture reference");
        auto it = tex_ids.find(t);
        if (it != tex_ids.end()) return it->second;
        PtTexture f{};
        f.kind = t->kind;
        f.odd = f.even = -1;
        f.image = -1;
        if (t->kind == Texture::Image) {
            f.image = image_id(t->image);
        } else if (t->kind == Texture::Constant) {
            f.color[0] = t->color.x; f.color[1] = t->color.y; f.color[2] = t->color.z;
        } else if (t->kind == Texture::Checker) {
            f.odd = texture_id(t->odd);
            f.even = texture_id(t->even);
        } else {
            f.scale = t->scale;
            if (perlin && perlin != t->noise) throw std::runtime_error("more than one Perlin table in a scene");
            perlin = t->noise;
        }
        textures.push_back(f);
        const int32_t id = (int32_t)textures.size() - 1;
        tex_ids.emplace(t, id);
        return id;
    }
    int32_t material_id(const Material* m) {
        if (!m) throw std::runtime_error("null material reference");
        auto it = mat_ids.find(m);
        if (it != mat_ids.end()) return it->second;
        PtMaterial f{};
        f.kind = m->kind;
        f.texture = -1;
        if (m->kind == Material::Lambertian || m->kind == Material::DiffuseLight) f.texture = texture_id(m->albedo_tex);
        f.albedo[0] = m->albedo.x; f.albedo[1] = m->albedo.y; f.albedo[2] = m->albedo.z;
        f.fuzz = m->fuzz;
        f.ref_idx = m->ref_idx;
        materials.push_back(f);
        const int32_t id = (int32_t)materials.size() - 1;
        mat_ids.emplace(m, id);
        return id;
    }
};
}  // namespace

Scene::Scene(const std::vector<Hitable>& world, std::optional<Vec3> sky, int device) : Scene(world, sky, std::vector<int>{device}, nullptr) {}
Scene::Scene(const std::vector<Hitable>& world, std::optional<Vec3> sky, const std::vector<int>& devices, const PtOptions* options) {
    Flattener fl;
    for (const Hitable& h : world) {
        PtMotion mo{};
        if (h.kind == Hitable::MovingSphereKind && h.moving_sphere) {  // hitable.rs:17,52-58
            const MovingSphere& ms = *h.moving_sphere;
            fl.cx.push_back(ms.centre0().x);
            fl.cy.push_back(ms.centre0().y);
            fl.cz.push_back(ms.centre0().z);
            fl.radius.push_back(ms.radius());
            mo.centre1[0] = ms.centre1().x; mo.centre1[1] = ms.centre1().y; mo.centre1[2] = ms.centre1().z;
            mo.time0 = ms.time0();
            mo.time1 = ms.time1();
            mo.moving = 1;
            fl.any_moving = true;
        } else if (h.kind == Hitable::SphereKind && h.sphere) {
            fl.cx.push_back(h.sphere->centre().x);
            fl.cy.push_back(h.sphere->centre().y);
            fl.cz.push_back(h.sphere->centre().z);
            fl.radius.push_back(h.sphere->radius());
        } else {  // spheres_soa.rs:49-51
            throw std::runtime_error("Expected Hitable::Sphere, got " + (h.what.empty() ? std::string("<unknown>") : h.what));
        }
        fl.motion.push_back(mo);
        fl.material_index.push_back(fl.material_id(h.material));
    }
    PtSceneDesc d{};
    d.struct_size = sizeof(PtSceneDesc);
    d.n_spheres = (uint32_t)fl.cx.size();
    d.centre_x = fl.cx.data();
    d.centre_y = fl.cy.data();
    d.centre_z = fl.cz.data();
    d.radius = fl.radius.data();
    d.material_index = fl.material_index.data();
    d.n_materials = (uint32_t)fl.materials.size();
    d.materials = fl.materials.data();
    d.n_textures = (uint32_t)fl.textures.size();
    d.textures = fl.textures.data();
    d.perlin = fl.perlin ? &fl.perlin->tables() : nullptr;
    d.motion = fl.any_moving ? fl.motion.data() : nullptr;
    d.n_images = (uint32_t)fl.images.size();
    d.images = fl.images.empty() ? nullptr : fl.images.data();
    d.has_sky = sky.has_value() ? 1u : 0u;
    if (sky) { d.sky[0] = sky->x; d.sky[1] = sky->y; d.sky[2] = sky->z; }
    n_spheres_ = d.n_spheres;
    if (pt_scene_create_multi(&d, devices.data(), (uint32_t)devices.size(), options, &scene_) != PT_OK)
        throw std::runtime_error(std::string("pt_scene_create: ") + pt_last_error());
}
Scene::~Scene() { pt_scene_destroy(scene_); }

size_t Scene::update(const Params& params, const Camera& camera, uint32_t frame_num, float* buffer, size_t buffer_len_pixels) const {
    PtPartition whole{0, 0, 1, 0};
    return update_part(params, camera, frame_num, whole, buffer, buffer_len_pixels);
}
size_t Scene::update_part(const Params& params, const Camera& camera, uint32_t frame_num, const PtPartition& part, float* buffer,
                          size_t buffer_len_pixels) const {
    if (buffer_len_pixels != (size_t)params.width * params.height) throw std::runtime_error("buffer length != width*height");
    const PtParams p = params.to_ffi();
    const PtCamera c = camera.to_ffi();
    uint64_t rays = 0;
    if (pt_render_part(scene_, &p, &c, frame_num, &part, buffer, &rays) != PT_OK)
        throw std::runtime_error(std::string("pt_render: ") + pt_last_error());
    return (size_t)rays;
}

size_t Scene::update_progressive(const Params& params, const Camera& camera, uint32_t frame_num, float* buffer, uint8_t* rgb8) const {
    const PtParams p = params.to_ffi();
    const PtCamera c = camera.to_ffi();
    uint64_t rays = 0;
    if (pt_render_progressive(scene_, &p, &c, frame_num, buffer, rgb8, &rays) != PT_OK)
        throw std::runtime_error(std::string("pt_render_progressive: ") + pt_last_error());
    return (size_t)rays;
}

// ---- Params (src/params.rs:21-46) ----
Xoshiro256Plus Params::new_rng() const { return Xoshiro256Plus::seed_from_u64(random_seed ? seed_salt : 0); }
std::unique_ptr<Scene> Params::new_scene(Xoshiro256Plus&, const Storage&, std::vector<Hitable> hitables, std::optional<Vec3> sky,
                                         int device) const {
    if (use_bvh) throw std::runtime_error("use_bvh: the BVH arm stays on the CPU reference (params.rs:36-40); the GPU path is the flat list");
    return std::make_unique<Scene>(hitables, sky, device);
}
std::unique_ptr<Scene> Params::new_scene(Xoshiro256Plus&, const Storage&, std::vector<Hitable> hitables, std::optional<Vec3> sky,
                                         const std::vector<int>& devices, const PtOptions* options) const {
    if (use_bvh) throw std::runtime_error("use_bvh: the BVH arm stays on the CPU reference (params.rs:36-40); the GPU path is the flat list");
    return std::make_unique<Scene>(hitables, sky, devices, options);
}
PtParams Params::to_ffi() const {
    PtParams p{};
    p.width = width; p.height = height; p.samples = samples; p.max_depth = max_depth;
    p.random_seed = random_seed; p.use_bvh = use_bvh; p.seed_salt = seed_salt;
    return p;
}

// ---- presets (src/presets.rs) ----
namespace presets {
namespace {
Camera rtiow_camera(const Params& params, float aperture, float time1) {  // presets.rs:95-109 / :275-289
    return Camera::create(Vec3(13.0f, 2.0f, 3.0f), Vec3(0.0f, 0.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 20.0f,
                          (float)params.width / (float)params.height, aperture, 10.0f, 0.0f, time1);
}
Hitable sphere(Storage& st, Vec3 centre, float radius, Material m) {
    return Hitable::make_sphere(st.alloc_sphere(Sphere(centre, radius)), st.alloc_material(m));
}
const Texture* constant(Storage& st, Vec3 c) { return st.alloc_texture(texture::constant(c)); }

Hitable moving_sphere(Storage& st, Vec3 centre0, Vec3 centre1, float radius, Material m) {  // presets.rs:122-127
    return Hitable::make_moving_sphere(st.alloc_moving_sphere(MovingSphere(centre0, centre1, 0.0f, 1.0f, radius)), st.alloc_material(m));
}

// presets.rs:89-215.  only_spheres = true is `random_spheres`, false is `random` (the Lambertian spheres move).
// grid_half = 11 is the reference preset; 158 is stress100k.
Preset random_impl(const Params& params, Xoshiro256Plus& rng, Storage& st, int grid_half, bool only_spheres = true) {
    Camera camera = rtiow_camera(params, 0.1f, 1.0f);
    std::vector<Hitable> hitables;
    hitables.reserve((size_t)4 * grid_half * grid_half + 4);
    const Texture* odd = constant(st, Vec3(0.2f, 0.3f, 0.1f));
    const Texture* even = constant(st, Vec3(0.9f, 0.9f, 0.9f));
    hitables.push_back(sphere(st, Vec3(0.0f, -1000.0f, 0.0f), 1000.0f, material::lambertian(st.alloc_texture(texture::checker(odd, even)))));
    for (int a = -grid_half; a < grid_half; ++a) {
        for (int b = -grid_half; b < grid_half; ++b) {
            const float choose_material = rng.gen_f32();
            const float x = (float)a + 0.9f * rng.gen_f32();
            const float z = (float)b + 0.9f * rng.gen_f32();
            const Vec3 centre(x, 0.2f, z);
            if (choose_material < 0.8f) {
                const Vec3 centre1 = centre + Vec3(0.0f, 0.5f * rng.gen_f32(), 0.0f);  // drawn even for the static variant (presets.rs:150)
                float ch[3];
                for (float& c : ch) {
                    const float p = rng.gen_f32();
                    const float q = rng.gen_f32();
                    c = p * q;
                }
                const Material m = material::lambertian(constant(st, Vec3(ch[0], ch[1], ch[2])));
                hitables.push_back(only_spheres ? sphere(st, centre, 0.2f, m) : moving_sphere(st, centre, centre1, 0.2f, m));
            } else if (choose_material < 0.95f) {
                float ch[3];
                for (float& c : ch) c = 0.5f * (1.0f + rng.gen_f32());
                const float fuzz = 0.5f * rng.gen_f32();
                hitables.push_back(sphere(st, centre, 0.2f, material::metal(Vec3(ch[0], ch[1], ch[2]), fuzz)));
            } else {
                hitables.push_back(sphere(st, centre, 0.2f, material::dielectric(1.5f)));
            }
        }
    }
    hitables.push_back(sphere(st, Vec3(0.0f, 1.0f, 0.0f), 1.0f, material::dielectric(1.5f)));
    hitables.push_back(sphere(st, Vec3(-4.0f, 1.0f, 0.0f), 1.0f, material::lambertian(constant(st, Vec3(0.4f, 0.2f, 0.1f)))));
    hitables.push_back(sphere(st, Vec3(4.0f, 1.0f, 0.0f), 1.0f, material::metal(Vec3(0.7f, 0.6f, 0.5f), 0.0f)));
    return Preset(std::move(hitables), camera, std::nullopt);
}
Preset small(const Params& params, Storage& st) {  // presets.rs:217-269
    const Vec3 lookfrom(3.0f, 3.0f, 2.0f), lookat(0.0f, 0.0f, -1.0f);
    Camera camera = Camera::create(lookfrom, lookat, Vec3(0.0f, 1.0f, 0.0f), 20.0f, (float)params.width / (float)params.height, 0.1f,
                                   length(lookfrom - lookat), 0.0f, 1.0f);
    std::vector<Hitable> h;
    h.push_back(sphere(st, Vec3(0.0f, 0.0f, -1.0f), 0.5f, material::lambertian(constant(st, Vec3(0.1f, 0.2f, 0.5f)))));
    h.push_back(sphere(st, Vec3(0.0f, -100.5f, -1.0f), 100.0f, material::lambertian(constant(st, Vec3(0.8f, 0.8f, 0.0f)))));
    h.push_back(sphere(st, Vec3(1.0f, 0.0f, -1.0f), 0.5f, material::metal(Vec3(0.8f, 0.6f, 0.2f), 0.0f)));
    h.push_back(sphere(st, Vec3(-1.0f, 0.0f, -1.0f), 0.5f, material::dielectric(1.5f)));
    h.push_back(sphere(st, Vec3(-1.0f, 0.0f, -1.0f), -0.45f, material::dielectric(1.5f)));
    return Preset(std::move(h), camera, std::nullopt);
}
Preset two_perlin_spheres(const Params& params, Storage& st) {  // presets.rs:271-315
    Camera camera = rtiow_camera(params, 0.0f, 0.0f);
    const Texture* noise_texture = st.alloc_texture(texture::noise(&st.perlin_noise, 4.0f));
    std::vector<Hitable> h;
    h.push_back(sphere(st, Vec3(0.0f, -1000.0f, 0.0f), 1000.0f, material::lambertian(noise_texture)));
    h.push_back(sphere(st, Vec3(0.0f, 2.0f, 0.0f), 2.0f, material::lambertian(noise_texture)));
    return Preset(std::move(h), camera, std::nullopt);
}
Preset smallpt(const Params& params, Storage& st) {  // presets.rs:853-930
    Camera camera = Camera::create(Vec3(50.0f, 52.0f, 295.6f), Vec3(50.0f, 33.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 30.0f,
                                   (float)params.width / (float)params.height, 0.05f, 100.0f, 0.0f, 1.0f);
    auto wall = [&](Vec3 c, Vec3 albedo) { return sphere(st, c, 1e3f, material::lambertian(constant(st, albedo))); };
    std::vector<Hitable> h;
    h.push_back(wall(Vec3(1e3f + 1.0f, 40.8f, 81.6f), Vec3(0.75f, 0.25f, 0.25f)));
    h.push_back(wall(Vec3(-1e3f + 99.0f, 40.8f, 81.6f), Vec3(0.25f, 0.25f, 0.75f)));
    h.push_back(wall(Vec3(50.0f, 40.8f, 1e3f), Vec3(0.75f, 0.75f, 0.75f)));
    h.push_back(wall(Vec3(50.0f, 1e3f, 81.6f), Vec3(0.75f, 0.75f, 0.75f)));
    h.push_back(wall(Vec3(50.0f, -1e3f + 81.6f, 81.6f), Vec3(0.75f, 0.75f, 0.75f)));
    h.push_back(sphere(st, Vec3(27.0f, 16.5f, 47.0f), 16.5f, material::metal(Vec3(1.0f, 1.0f, 1.0f) * 0.999f, 0.0f)));
    h.push_back(sphere(st, Vec3(73.0f, 16.5f, 78.0f), 16.5f, material::dielectric(1.5f)));
    h.push_back(sphere(st, Vec3(50.0f, 81.6f - 16.5f, 81.6f), 1.5f, material::diffuse_light(constant(st, Vec3(4.0f, 4.0f, 4.0f) * 100.0f))));
    return Preset(std::move(h), camera, Vec3(0.0f, 0.0f, 0.0f));
}
Preset earth(const Params& params, Storage& st) {  // presets.rs:555-594
    Camera camera = rtiow_camera(params, 0.0f, 0.0f);
    const char* env = std::getenv("PATHTRACE_EARTHMAP");
    const RgbImage* earth_image = st.alloc_image(RgbImage::open(env && *env ? env : "media/earthmap.jpg"));
    const Texture* earth_texture = st.alloc_texture(texture::rgb_image(earth_image));
    std::vector<Hitable> h;
    h.push_back(sphere(st, Vec3(0.0f, 0.0f, 0.0f), 2.0f, material::lambertian(earth_texture)));
    return Preset(std::move(h), camera, std::nullopt);
}
}  // namespace

std::optional<Preset> from_name(const std::string& name, const Params& params, Xoshiro256Plus& rng, Storage& storage, bool quiet) {
    if (!quiet)  // presets.rs:19-22
        std::printf("generating '%s' preset at %ux%u with %u samples per pixel\n", name.c_str(), params.width, params.height, params.samples);
    if (name == "random") return random_impl(params, rng, storage, 11, false);
    if (name == "random_spheres") return random_impl(params, rng, storage, 11);
    if (name == "stress100k") return random_impl(params, rng, storage, 158);
    if (name == "small") return small(params, storage);
    if (name == "smallpt") return smallpt(params, storage);
    if (name == "two_perlin_spheres") return two_perlin_spheres(params, storage);
    if (name == "earth") return earth(params, storage);
    if (name == "final") return Preset({}, rtiow_camera(params, 0.1f, 1.0f), std::nullopt);  // presets.rs:40-71: an empty stub
    return std::nullopt;
}
}  // namespace presets

// ---- output stage + PNG ----
void linear_to_srgb8_image(const Scene& scene, const float* rgb, uint32_t width, uint32_t height, std::vector<uint8_t>& out) {
    out.resize((size_t)width * height * 3);
    if (pt_srgb8(scene.handle(), rgb, width, height, out.data()) != PT_OK) throw std::runtime_error(std::string("pt_srgb8: ") + pt_last_error());
}

namespace {
uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return crc;
}
void put_be32(std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }
void png_chunk(std::vector<uint8_t>& file, const char* type, const std::vector<uint8_t>& data) {
    put_be32(file, (uint32_t)data.size());
    std::vector<uint8_t> body(type, type + 4);
    body.insert(body.end(), data.begin(), data.end());
    file.insert(file.end(), body.begin(), body.end());
    put_be32(file, crc32_update(0xffffffffu, body.data(), body.size()) ^ 0xffffffffu);
}
}  // namespace

// Minimal PNG encoder (8-bit RGB, zlib "stored" blocks) — the reference uses the `image` crate (offline.rs:52-59).
bool write_png_rgb8(const std::string& path, const uint8_t* rgb, uint32_t width, uint32_t height) {
    std::vector<uint8_t> raw;
    raw.reserve(((size_t)width * 3 + 1) * height);
    for (uint32_t y = 0; y < height; ++y) {
        raw.push_back(0);  // filter: none
        raw.insert(raw.end(), rgb + (size_t)y * width * 3, rgb + (size_t)(y + 1) * width * 3);
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
        const size_t n = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + n >= raw.size() ? 1 : 0);
        z.push_back((uint8_t)(n & 0xff)); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xff)); z.push_back((uint8_t)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        for (size_t i = 0; i < n; ++i) { a = (a + raw[off + i]) % 65521; b = (b + a) % 65521; }
        if (raw.empty()) break;
    }
    put_be32(z, (b << 16) | a);
    std::vector<uint8_t> file = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width); put_be32(ihdr, height);
    ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0});
    png_chunk(file, "IHDR", ihdr);
    png_chunk(file, "IDAT", z);
    png_chunk(file, "IEND", {});
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(file.data(), 1, file.size(), f) == file.size();
    std::fclose(f);
    return ok;
}

// ---- offline::render_offline (src/offline.rs:16-60) ----
namespace offline {
std::pair<double, size_t> render_offline(const std::string& preset, const Params& params, const std::string& output_png, int device) {
    return render_offline(preset, params, output_png, std::vector<int>{device}, 1);
}
std::pair<double, size_t> render_offline(const std::string& preset, const Params& params, const std::string& output_png,
                                         const std::vector<int>& devices, uint32_t frames) {
    Xoshiro256Plus rng = params.new_rng();
    Storage storage(rng);
    auto built = presets::from_name(preset, params, rng, storage);
    if (!built) throw std::runtime_error("unrecognised preset");
    auto& [hitables, camera, sky] = *built;
    std::unique_ptr<Scene> scene = params.new_scene(rng, storage, std::move(hitables), sky, devices, nullptr);
    std::vector<float> rgb_buffer((size_t)params.width * params.height * 3, 0.0f);

    const auto start_time = std::chrono::steady_clock::now();
    size_t ray_count = 0;
    if (frames <= 1) {
        const uint32_t frame_num = 0;  // only ever processing 1 frame in offline
        ray_count = scene->update(params, camera, frame_num, rgb_buffer.data(), (size_t)params.width * params.height);
    } else {
        // glium_window.rs:98-131 without the window: the buffer stays on the device, only the last frame comes back
        for (uint32_t frame_num = 0; frame_num < frames; ++frame_num)
            ray_count += scene->update_progressive(params, camera, frame_num, frame_num + 1 == frames ? rgb_buffer.data() : nullptr, nullptr);
    }
    const double elapsed_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start_time).count();
    std::printf("%.2fsecs %zurays %.2fMrays/s\n", elapsed_secs, ray_count, (double)ray_count / 1000000.0 / elapsed_secs);

    if (!output_png.empty()) {
        std::vector<uint8_t> image_bytes;
        linear_to_srgb8_image(*scene, rgb_buffer.data(), params.width, params.height, image_bytes);
        if (!write_png_rgb8(output_png, image_bytes.data(), params.width, params.height)) throw std::runtime_error("Failed to save output image");
    }
    return {elapsed_secs, ray_count};
}
}  // namespace offline

}  // namespace pathtrace

// ================================================================================================
// C interface (ctypes: pathtrace_rs_b200/host.py).  Handles own Storage + Scene + Camera of one preset.
// ================================================================================================
namespace {
struct PresetHandle {
    pathtrace::Params params;
    pathtrace::Xoshiro256Plus rng;
    std::unique_ptr<pathtrace::Storage> storage;
    std::vector<pathtrace::Hitable> hitables;
    pathtrace::Camera camera;
    std::optional<pathtrace::Vec3> sky;
    std::unique_ptr<pathtrace::Scene> scene;  // created lazily (needs a GPU)
};
thread_local std::string g_err;
}  // namespace

extern "C" {

struct PthParams {
    uint32_t width, height, samples, max_depth;
    uint32_t random_seed, use_bvh;
    uint64_t seed_salt;
};

const char* pth_last_error(void) { return g_err.c_str(); }

static pathtrace::Params to_params(const PthParams* p) {
    pathtrace::Params q;
    q.width = p->width; q.height = p->height; q.samples = p->samples; q.max_depth = p->max_depth;
    q.random_seed = p->random_seed != 0; q.use_bvh = p->use_bvh != 0; q.seed_salt = p->seed_salt;
    return q;
}

// offline.rs:17-21: rng, Storage, preset — host only, no GPU touched
void* pth_preset_build(const char* name, const PthParams* p) {
    try {
        auto h = std::make_unique<PresetHandle>();
        h->params = to_params(p);
        h->rng = h->params.new_rng();
        h->storage = std::make_unique<pathtrace::Storage>(h->rng);
        auto built = pathtrace::presets::from_name(name, h->params, h->rng, *h->storage, true);
        if (!built) { g_err = "unrecognised preset"; return nullptr; }
        h->hitables = std::move(std::get<0>(*built));
        h->camera = std::get<1>(*built);
        h->sky = std::get<2>(*built);
        return h.release();
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void pth_preset_free(void* hv) { delete (PresetHandle*)hv; }
uint32_t pth_preset_len(void* hv) { return (uint32_t)((PresetHandle*)hv)->hitables.size(); }
void pth_preset_camera(void* hv, PtCamera* out) { *out = ((PresetHandle*)hv)->camera.to_ffi(); }
// next f32 of the scene rng after the preset was built (draw-order check)
float pth_preset_next_f32(void* hv) { pathtrace::Xoshiro256Plus r = ((PresetHandle*)hv)->rng; return r.gen_f32(); }

// flat dump of the host scene in the oracle's dump layout (tests compare the two builders bit for bit)
void pth_preset_spheres(void* hv, float* centre_radius, int32_t* kind, float* params5) {
    auto* h = (PresetHandle*)hv;
    for (size_t i = 0; i < h->hitables.size(); ++i) {
        const auto& hit = h->hitables[i];
        const bool mv = hit.kind == pathtrace::Hitable::MovingSphereKind;
        const pathtrace::Vec3 c0 = mv ? hit.moving_sphere->centre0() : hit.sphere->centre();
        centre_radius[4 * i] = c0.x; centre_radius[4 * i + 1] = c0.y;
        centre_radius[4 * i + 2] = c0.z; centre_radius[4 * i + 3] = mv ? hit.moving_sphere->radius() : hit.sphere->radius();
        const pathtrace::Material& m = *hit.material;
        kind[i] = m.kind;
        pathtrace::Vec3 c = m.albedo;
        if ((m.kind == pathtrace::Material::Lambertian || m.kind == pathtrace::Material::DiffuseLight) && m.albedo_tex->kind == pathtrace::Texture::Constant)
            c = m.albedo_tex->color;
        params5[5 * i] = c.x; params5[5 * i + 1] = c.y; params5[5 * i + 2] = c.z; params5[5 * i + 3] = m.fuzz; params5[5 * i + 4] = m.ref_idx;
    }
}
// per sphere: centre1 (3), time0, time1, moving flag — same layout as the oracle's orc_scene_motion
void pth_preset_motion(void* hv, float* out6) {
    auto* h = (PresetHandle*)hv;
    for (size_t i = 0; i < h->hitables.size(); ++i) {
        const auto& hit = h->hitables[i];
        const bool mv = hit.kind == pathtrace::Hitable::MovingSphereKind;
        const pathtrace::Vec3 c1 = mv ? hit.moving_sphere->centre1() : hit.sphere->centre();
        out6[6 * i] = c1.x; out6[6 * i + 1] = c1.y; out6[6 * i + 2] = c1.z;
        out6[6 * i + 3] = mv ? hit.moving_sphere->time0() : 0.0f;
        out6[6 * i + 4] = mv ? hit.moving_sphere->time1() : 0.0f;
        out6[6 * i + 5] = mv ? 1.0f : 0.0f;
    }
}
void pth_preset_perlin(void* hv, PtPerlin* out) { *out = ((PresetHandle*)hv)->storage->perlin_noise.tables(); }
int32_t pth_preset_sky(void* hv, float* sky3) {
    auto* h = (PresetHandle*)hv;
    if (!h->sky) return 0;
    sky3[0] = h->sky->x; sky3[1] = h->sky->y; sky3[2] = h->sky->z;
    return 1;
}

// params.new_scene (GPU upload).  Returns 0 on success.
int32_t pth_scene_create(void* hv, int32_t device) {
    auto* h = (PresetHandle*)hv;
    try {
        h->scene = h->params.new_scene(h->rng, *h->storage, h->hitables, h->sky, device);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
// the same on a list of devices with explicit PtOptions (may be null)
int32_t pth_scene_create_multi(void* hv, const int32_t* devices, uint32_t n_devices, const PtOptions* options) {
    auto* h = (PresetHandle*)hv;
    try {
        h->scene = h->params.new_scene(h->rng, *h->storage, h->hitables, h->sky, std::vector<int>(devices, devices + n_devices), options);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
void* pth_scene_handle(void* hv) { auto* h = (PresetHandle*)hv; return h->scene ? h->scene->handle() : nullptr; }

// Scene::update through the mirror (host buffer)
int32_t pth_scene_update(void* hv, const PthParams* p, uint32_t frame_num, const PtPartition* part, float* buffer, uint64_t* rays_out) {
    auto* h = (PresetHandle*)hv;
    try {
        if (!h->scene) throw std::runtime_error("scene not created");
        const pathtrace::Params q = to_params(p);
        const PtPartition whole{0, 0, 1, 0};
        *rays_out = h->scene->update_part(q, h->camera, frame_num, part ? *part : whole, buffer, (size_t)q.width * q.height);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int32_t pth_render_offline(const char* preset, const PthParams* p, const char* output_png, int32_t device, double* secs, uint64_t* rays) {
    try {
        auto r = pathtrace::offline::render_offline(preset, to_params(p), output_png ? output_png : "", device);
        if (secs) *secs = r.first;
        if (rays) *rays = r.second;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int32_t pth_render_offline_multi(const char* preset, const PthParams* p, const char* output_png, const int32_t* devices, uint32_t n_devices, uint32_t frames,
                                 double* secs, uint64_t* rays) {
    try {
        auto r = pathtrace::offline::render_offline(preset, to_params(p), output_png ? output_png : "", std::vector<int>(devices, devices + n_devices), frames);
        if (secs) *secs = r.first;
        if (rays) *rays = r.second;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// error-path probe: a world containing a non-sphere hitable must be rejected (spheres_soa.rs:49-51)
int32_t pth_flatten_rejects_non_sphere(void) {
    try {
        std::vector<pathtrace::Hitable> world{pathtrace::Hitable::unsupported("Rect")};
        pathtrace::Scene s(world, std::nullopt, 0);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// RgbImage::open (texture.rs:14-25): width/height, and the pixels when rgb_out != nullptr (cap bytes available)
int32_t pth_image_open(const char* path, uint32_t* width, uint32_t* height, uint8_t* rgb_out, uint64_t cap) {
    try {
        const pathtrace::RgbImage im = pathtrace::RgbImage::open(path);
        *width = im.width();
        *height = im.height();
        if (rgb_out) {
            if (cap < im.data().size()) throw std::runtime_error("buffer too small");
            std::memcpy(rgb_out, im.data().data(), im.data().size());
        }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int32_t pth_write_png(const char* path, const uint8_t* rgb, uint32_t w, uint32_t h) { return pathtrace::write_png_rgb8(path, rgb, w, h) ? 0 : 1; }

}  // extern "C"
