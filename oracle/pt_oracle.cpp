// pt_oracle.cpp — CPU ORACLE for the pathtrace-rs per-pixel path-tracing loop.
//
// TEST INFRASTRUCTURE ONLY.  This file is a plain C++17 restatement of the reference's CPU
// algorithm (bitshifter/pathtrace-rs, `Scene::update` and everything it calls).  It exists so
// the CUDA path can be checked against it.  Only tests/, __graft_entry__.smoke() and the
// `cpu_baseline` / `--impl reference` legs of bench.py may load it.  Nothing under
// pathtrace_rs_b200/ links, imports or calls it, and the product has no CPU fallback.
//
// PARITY STATUS: "parity unpinned" by the reference itself — the reference has no tests, no golden
// vectors and no fixtures, and it cannot be built here (no Rust toolchain).  What
// pins this restatement instead (tests/test_oracle_*.py):
//   * upstream known-answer vectors for SplitMix64 / xoshiro256+ (the RNG lives in the crates
//     rand 0.8.5 / rand_core 0.6.3 / rand_xoshiro 0.6.0, Cargo.lock:850-882, not in the tree);
//   * the cross-implementation equivalence the reference itself benchmarks (flat list vs SoA scalar
//     vs SoA AVX2 must return the same nearest hit for the bench fixture ray, src/bench.rs:17-26);
//   * closed-form invariants of Scene::update (empty scene, max_depth 0, frame blending).
//
// Build (see oracle/Makefile): g++ -O2 -ffp-contract=off ... (Rust never contracts a*b+c to FMA).
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#if defined(__AVX2__)
#include <immintrin.h>
#endif

namespace orc {

// ------------------------------------------------------------------------------------------------
// RNG — rand_xoshiro 0.6.0 Xoshiro256Plus + rand_core 0.6.3 seed_from_u64 + rand 0.8.5 Standard<f32>
// (third-party crates, source not in /root/reference; algorithm = published xoshiro256+ / SplitMix64;
//  call sites: src/params.rs:21-27, src/scene.rs:96-102, every `rng.gen::<f32>()`).
// ------------------------------------------------------------------------------------------------
struct Rng {
    uint64_t s[4];

    static uint64_t splitmix64(uint64_t& x) {
        x += 0x9e3779b97f4a7c15ULL;
        uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    // Xoshiro256Plus::seed_from_u64: four SplitMix64 outputs, little-endian fill.
    static Rng seed_from_u64(uint64_t seed) {
        Rng r;
        for (int i = 0; i < 4; ++i) r.s[i] = splitmix64(seed);
        return r;
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next_u64() {
        const uint64_t result = s[0] + s[3];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    // rand 0.8.5 `Standard` for f32: 24 high bits of a u32, scaled by 2^-24 -> [0,1)
    float gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
};

// ------------------------------------------------------------------------------------------------
// glam 0.20.5 Vec3 (scalar f32 math; no FMA): dot = (x*x + y*y) + z*z, normalize = v * (1/length)
// ------------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
static inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 splat(float a) { return V3{a, a, a}; }
static inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
static inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
static inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
static inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float length(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 normalize(V3 a) { return a * (1.0f / length(a)); }
static inline V3 cross(V3 a, V3 b) {
    return V3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}

// ------------------------------------------------------------------------------------------------
// simd.rs:107-208 — sinf_cosf (Cephes / sse_mathfun polynomial), lane 0 only (simd.rs:111-112)
// ------------------------------------------------------------------------------------------------
static inline void sinf_cosf(float xin, float& s_out, float& c_out) {
    uint32_t xb;
    std::memcpy(&xb, &xin, 4);
    uint32_t sign_bit_sin = xb & 0x80000000u;       // simd.rs:126
    xb &= 0x7fffffffu;                              // simd.rs:124
    float x;
    std::memcpy(&x, &xb, 4);
    float y = x * 1.27323954473516f;                // simd.rs:129  (4/pi)
    int32_t j = (int32_t)y;                         // cvttps (truncate)  :132
    j = (j + 1) & ~1;                               // :135-136
    y = (float)j;                                   // :137
    int32_t j4 = j;
    uint32_t swap_sign_bit_sin = ((uint32_t)(j & 4)) << 29;   // :142-144
    bool poly_mask = (j & 2) == 0;                  // :147-149
    // extended precision modular arithmetic  :153-161
    x = x + y * -0.78515625f;
    x = x + y * -2.4187564849853515625e-4f;
    x = x + y * -3.77489497744594108e-8f;
    j4 = j4 - 2;                                    // :163
    uint32_t sign_bit_cos = ((~(uint32_t)j4) & 4u) << 29;     // :164-166
    sign_bit_sin ^= swap_sign_bit_sin;              // :168
    float z = x * x;                                // :171
    float yc = 2.443315711809948E-005f;             // :172-181
    yc = yc * z;
    yc = yc + -1.388731625493765E-003f;
    yc = yc * z;
    yc = yc + 4.166664568298827E-002f;
    yc = yc * z;
    yc = yc * z;
    float tmp = z * 0.5f;
    yc = yc - tmp;
    yc = yc + 1.0f;
    float ys = -1.9515295891E-4f;                   // :184-191
    ys = ys * z;
    ys = ys + 8.3321608736E-3f;
    ys = ys * z;
    ys = ys + -1.6666654611E-1f;
    ys = ys * z;
    ys = ys * x;
    ys = ys + x;
    // select  :194-201   sin takes y2 where poly_mask else y; cos takes the other
    float sinv = poly_mask ? ys : yc;
    float cosv = poly_mask ? yc : ys;
    uint32_t sb, cb;
    std::memcpy(&sb, &sinv, 4);
    std::memcpy(&cb, &cosv, 4);
    sb ^= sign_bit_sin;                             // :204-207
    cb ^= sign_bit_cos;
    std::memcpy(&s_out, &sb, 4);
    std::memcpy(&c_out, &cb, 4);
}

// ------------------------------------------------------------------------------------------------
// math.rs
// ------------------------------------------------------------------------------------------------
// math.rs:6-13
static inline V3 random_in_unit_disk(Rng& rng) {
    for (;;) {
        float a = rng.gen_f32();
        float b = rng.gen_f32();
        V3 p = 2.0f * v3(a, b, 0.0f) - v3(1.0f, 1.0f, 0.0f);
        if (dot(p, p) < 1.0f) return p;
    }
}
// math.rs:15-26
static inline V3 random_in_unit_sphere(Rng& rng) {
    for (;;) {
        float a = 2.0f * rng.gen_f32() - 1.0f;
        float b = 2.0f * rng.gen_f32() - 1.0f;
        float c = 2.0f * rng.gen_f32() - 1.0f;
        V3 p = v3(a, b, c);
        if (dot(p, p) < 1.0f) return p;
    }
}
// math.rs:28-34
static inline V3 random_unit_vector(Rng& rng) {
    float z = rng.gen_f32() * 2.0f - 1.0f;
    float a = rng.gen_f32() * 2.0f * 3.14159265358979323846f;
    float r = std::sqrt(1.0f - z * z);
    float sina, cosa;
    sinf_cosf(a, sina, cosa);
    return v3(r * cosa, r * sina, z);
}
// math.rs:36-48
static inline void linear_to_srgb(const float rgb[3], uint8_t out[3]) {
    for (int i = 0; i < 3; ++i) {
        float c = std::max(rgb[i], 0.0f);
        float s = 1.055f * std::pow(c, 0.41666666f) - 0.055f;
        s = std::min(std::max(s, 0.0f), 1.0f);
        out[i] = (uint8_t)(s * 255.99f);
    }
}
// math.rs:61-63
static inline V3 reflect(V3 v, V3 n) { return v - 2.0f * dot(v, n) * n; }
// math.rs:65-73
static inline bool refract(V3 v, V3 n, float ni_over_nt, V3& out) {
    float dt = dot(v, n);
    float discriminant = 1.0f - (ni_over_nt * ni_over_nt) * (1.0f - (dt * dt));
    if (discriminant > 0.0f) {
        out = ni_over_nt * (v - n * dt) - n * std::sqrt(discriminant);
        return true;
    }
    return false;
}
// math.rs:76-80
static inline float schlick(float cosine, float ref_idx) {
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * std::pow(1.0f - cosine, 5.0f);
}

// ------------------------------------------------------------------------------------------------
// perlin.rs
// ------------------------------------------------------------------------------------------------
struct Perlin {
    V3 randvec[256];
    uint32_t perm_x[256], perm_y[256], perm_z[256];

    // perlin.rs:28-42
    static void generate_perm(Rng& rng, uint32_t* perm) {
        for (int i = 0; i < 256; ++i) perm[i] = (uint32_t)i;
        for (int i = 255; i >= 0; --i) {
            size_t target = (size_t)std::floor(rng.gen_f32() * (float)(i + 1));
            std::swap(perm[i], perm[target]);
        }
    }
    // perlin.rs:15-26, :44-51  (draw order: randvec, perm_x, perm_y, perm_z)
    void init(Rng& rng) {
        for (int i = 0; i < 256; ++i) {
            float a = -1.0f + 2.0f * rng.gen_f32();
            float b = -1.0f + 2.0f * rng.gen_f32();
            float c = -1.0f + 2.0f * rng.gen_f32();
            randvec[i] = normalize(v3(a, b, c));
        }
        generate_perm(rng, perm_x);
        generate_perm(rng, perm_y);
        generate_perm(rng, perm_z);
    }
    // Rust `f32 as usize` saturates: negative / NaN -> 0 (perlin.rs:96-98)
    static size_t sat_usize(float f) {
        if (!(f > 0.0f)) return 0;
        if (f >= 18446744073709551616.0f) return ~(size_t)0;
        return (size_t)f;
    }
    // perlin.rs:54-74
    static float interpolate(const V3 c[2][2][2], float u, float v, float w) {
        float uu = u * u * (3.0f - 2.0f * u);
        float vv = v * v * (3.0f - 2.0f * v);
        float ww = w * w * (3.0f - 2.0f * w);
        float accum = 0.0f;
        for (int i = 0; i < 2; ++i) {
            float ii = (float)i;
            for (int j = 0; j < 2; ++j) {
                float jj = (float)j;
                for (int k = 0; k < 2; ++k) {
                    float kk = (float)k;
                    V3 weight = v3(u - ii, v - jj, w - kk);
                    accum += (ii * uu + (1.0f - ii) * (1.0f - uu)) * (jj * vv + (1.0f - jj) * (1.0f - vv)) *
                             (kk * ww + (1.0f - kk) * (1.0f - ww)) * dot(c[i][j][k], weight);
                }
            }
        }
        return accum;
    }
    // perlin.rs:89-111
    float noise(V3 p) const {
        float fx = std::floor(p.x), fy = std::floor(p.y), fz = std::floor(p.z);
        float u = p.x - fx, v = p.y - fy, w = p.z - fz;
        size_t i = sat_usize(fx), j = sat_usize(fy), k = sat_usize(fz);
        V3 c[2][2][2];
        for (size_t di = 0; di < 2; ++di)
            for (size_t dj = 0; dj < 2; ++dj)
                for (size_t dk = 0; dk < 2; ++dk)
                    c[di][dj][dk] = randvec[perm_x[(i + di) & 255] ^ perm_y[(j + dj) & 255] ^ perm_z[(k + dk) & 255]];
        return interpolate(c, u, v, w);
    }
    // perlin.rs:76-87
    float turb(V3 p) const {
        float accum = 0.0f;
        V3 temp_p = p;
        float weight = 1.0f;
        for (int d = 0; d < 7; ++d) {
            accum += weight * noise(temp_p);
            weight *= 0.5f;
            temp_p = temp_p * 2.0f;
        }
        return std::fabs(accum);
    }
};

// ------------------------------------------------------------------------------------------------
// texture.rs / material.rs data model (arena refs become indices)
// ------------------------------------------------------------------------------------------------
enum TexKind : int32_t { TEX_CONSTANT = 0, TEX_CHECKER = 1, TEX_NOISE = 2, TEX_IMAGE = 3 };
struct Texture {
    int32_t kind;
    V3 color;           // Constant
    int32_t odd, even;  // Checker (indices)
    float scale;        // Noise
    int32_t image = -1; // Image (index into Scene::images)
};
// Rust `f32 as i32`: truncation towards zero, saturating, NaN -> 0
static inline int32_t rust_f32_as_i32(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}
// texture.rs:6-37
struct RgbImage {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> data;  // image.to_rgb8().into_raw(): row 0 = top
    V3 value(float u, float v) const {  // texture.rs:27-36
        int32_t i = rust_f32_as_i32(u * (float)width);
        int32_t j = rust_f32_as_i32((1.0f - v) * (float)height - 0.001f);
        size_t ii = (size_t)std::min(std::max(i, 0), (int32_t)width - 1);
        size_t jj = (size_t)std::min(std::max(j, 0), (int32_t)height - 1);
        float r = (float)data[3 * ii + 3 * (size_t)width * jj] / 255.0f;
        float g = (float)data[3 * ii + 3 * (size_t)width * jj + 1] / 255.0f;
        float b = (float)data[3 * ii + 3 * (size_t)width * jj + 2] / 255.0f;
        return v3(r, g, b);
    }
};
// material.rs:41-49 (`x.atan2(y)` = atan2(x, y))
static inline void get_sphere_uv(V3 normal, float& u, float& v) {
    const float PI = 3.14159265358979323846f, FRAC_1_2PI = 1.0f / (2.0f * PI);
    float phi = std::atan2(normal.x, normal.y);
    float theta = std::asin(normal.y);
    u = 1.0f - (phi + PI) * FRAC_1_2PI;
    v = (theta + 1.57079632679489661923f) * 0.318309886183790671538f;
}
enum MatKind : int32_t { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3 };
struct Material {
    int32_t kind;
    int32_t tex;   // Lambertian albedo / DiffuseLight emit
    V3 albedo;     // Metal
    float fuzz;    // Metal
    float ref_idx; // Dielectric
};
struct Sphere {
    V3 centre;  // Sphere.centre / MovingSphere.centre_start
    float radius;
    int32_t material;
    // collision/moving_sphere.rs:7-26 — MovingSphere::new(centre0, centre1, time0, time1, radius)
    bool moving = false;
    V3 centre_delta = {0, 0, 0};
    float time_start = 0.0f, inv_time_delta = 0.0f;
    // moving_sphere.rs:28-31
    V3 centre_at(float time) const { return centre + ((time - time_start) * inv_time_delta) * centre_delta; }
};

// camera.rs:8-19
struct Camera {
    V3 origin, lower_left_corner, horizontal, vertical, u, v, w;
    float time0, time1, lens_radius;
};
// camera.rs:22-54
static Camera camera_new(V3 lookfrom, V3 lookat, V3 vup, float vfov, float aspect, float aperture, float focus_dist,
                         float time0, float time1) {
    float theta = vfov * 3.14159265358979323846f / 180.0f;
    float half_height = std::tan(theta * 0.5f);
    float half_width = aspect * half_height;
    V3 w = normalize(lookfrom - lookat);
    V3 u = normalize(cross(vup, w));
    V3 v = cross(w, u);
    Camera c;
    c.origin = lookfrom;
    c.lower_left_corner = lookfrom - half_width * focus_dist * u - half_height * focus_dist * v - focus_dist * w;
    c.horizontal = 2.0f * half_width * focus_dist * u;
    c.vertical = 2.0f * half_height * focus_dist * v;
    c.u = u;
    c.v = v;
    c.w = w;
    c.time0 = time0;
    c.time1 = time1;
    c.lens_radius = aperture * 0.5f;
    return c;
}

// collision/ray.rs:4-26 (rcp_direction is only used by AABB tests — out of scope)
struct Ray {
    V3 origin, direction;
    float time;
};
static inline V3 point_at(const Ray& r, float t) { return r.origin + (t * r.direction); }

// camera.rs:56-68
static inline Ray get_ray(const Camera& c, float s, float t, Rng& rng) {
    V3 rd = c.lens_radius * random_in_unit_disk(rng);
    V3 offset = c.u * rd.x + c.v * rd.y;
    float time = c.time0 + rng.gen_f32() * (c.time1 - c.time0);
    Ray r;
    r.origin = c.origin + offset;
    r.direction = normalize(c.lower_left_corner + s * c.horizontal + t * c.vertical - c.origin - offset);
    r.time = time;
    return r;
}

// collision/ray.rs:43-50
struct RayHit {
    V3 point, normal;
    float t, u, v;
};

// ------------------------------------------------------------------------------------------------
// Scene = flat sphere list + materials + textures + perlin + sky  (scene.rs:18-22, hitable_list.rs:9-11)
// ------------------------------------------------------------------------------------------------
enum HitMode : int32_t { HIT_LIST = 0, HIT_SOA_SCALAR = 1, HIT_SOA_AVX2 = 2 };

// debug recorder (tools/ray_stats.py): every ray handed to ray_hit, in trace order
struct RayRecorder {
    float* out = nullptr;  // 6 floats per ray
    int64_t cap = 0, n = 0;
    std::vector<float>* vec = nullptr;   // growable alternative to `out` (orc_record_rays): 6 floats per ray ...
    std::vector<float>* tvec = nullptr;  // ... and the ray's time
};
static thread_local RayRecorder* g_recorder = nullptr;

struct Scene {
    std::vector<Sphere> spheres;
    std::vector<Material> materials;
    std::vector<Texture> textures;
    std::vector<RgbImage> images;
    Perlin perlin;
    bool has_sky = false;
    V3 sky = {0, 0, 0};
    Camera camera;
    // SoA mirror (spheres_soa.rs:12-23, :26-74), padded to 8 with centre=f32::MAX, r^2=0
    std::vector<float> cx, cy, cz, rsq, rinv;
    size_t soa_len = 0;
    std::vector<int32_t> moving_index;  // Hitable::MovingSphere entries (hybrid SoA modes, see ray_hit)

    void build_soa() {
        size_t n = spheres.size();
        soa_len = (n + 7) & ~(size_t)7;  // align_to(num, 8) math.rs:83-85 with the AVX2 chunk
        cx.assign(soa_len, std::numeric_limits<float>::max());
        cy.assign(soa_len, std::numeric_limits<float>::max());
        cz.assign(soa_len, std::numeric_limits<float>::max());
        rsq.assign(soa_len, 0.0f);
        rinv.assign(soa_len, 0.0f);
        moving_index.clear();
        for (size_t i = 0; i < n; ++i) {
            if (spheres[i].moving) {  // not representable in SpheresSoA (spheres_soa.rs:49-51 panics): tested by the live form
                moving_index.push_back((int32_t)i);
                continue;
            }
            cx[i] = spheres[i].centre.x;
            cy[i] = spheres[i].centre.y;
            cz[i] = spheres[i].centre.z;
            rsq[i] = spheres[i].radius * spheres[i].radius;
            rinv[i] = 1.0f / spheres[i].radius;
        }
    }

    // texture.rs:74-91
    V3 tex_value(int32_t ti, float u, float v, V3 p) const {
        const Texture& t = textures[ti];
        switch (t.kind) {
            case TEX_IMAGE: return images[t.image].value(u, v);
            case TEX_CONSTANT: return t.color;
            case TEX_CHECKER: {
                V3 s = v3(10.0f, 10.0f, 10.0f) * p;
                float sines = std::sin(s.x) * std::sin(s.y) * std::sin(s.z);
                return sines < 0.0f ? tex_value(t.odd, u, v, p) : tex_value(t.even, u, v, p);
            }
            default:  // TEX_NOISE
                return v3(1.0f, 1.0f, 1.0f) * 0.5f * (1.0f + std::sin(t.scale * p.z + 10.0f * perlin.turb(p)));
        }
    }

    // collision/sphere.rs:29-66 (live AoS test; glam Vec3A dot has the same association as Vec3)
    static bool sphere_hit(const Sphere& s, const Ray& ray, float t_min, float t_max, RayHit& out) {
        V3 oc = ray.origin - s.centre;
        float a = dot(ray.direction, ray.direction);
        float b = dot(oc, ray.direction);
        float c = dot(oc, oc) - s.radius * s.radius;
        float discriminant = b * b - a * c;
        if (discriminant > 0.0f) {
            float dsq = std::sqrt(discriminant);
            float t = (-b - dsq) / a;
            if (t < t_max && t > t_min) {
                out.point = point_at(ray, t);
                out.normal = (out.point - s.centre) / s.radius;
                out.t = t;
                out.u = 0.0f;
                out.v = 0.0f;
                return true;
            }
            t = (-b + dsq) / a;
            if (t < t_max && t > t_min) {
                out.point = point_at(ray, t);
                out.normal = (out.point - s.centre) / s.radius;
                out.t = t;
                out.u = 0.0f;
                out.v = 0.0f;
                return true;
            }
        }
        return false;
    }
    // collision/moving_sphere.rs:38-73 (same roots test as Sphere::ray_hit, centre evaluated at ray.time)
    static bool moving_sphere_hit(const Sphere& s, const Ray& ray, float t_min, float t_max, RayHit& out) {
        V3 centre = s.centre_at(ray.time);
        V3 oc = ray.origin - centre;
        float a = dot(ray.direction, ray.direction);
        float b = dot(oc, ray.direction);
        float c = dot(oc, oc) - s.radius * s.radius;
        float discriminant = b * b - a * c;
        if (discriminant > 0.0f) {
            float dsq = std::sqrt(discriminant);
            float t = (-b - dsq) / a;
            if (t < t_max && t > t_min) {
                out.point = point_at(ray, t);
                out.normal = (out.point - centre) / s.radius;
                out.t = t;
                out.u = 0.0f;
                out.v = 0.0f;
                return true;
            }
            t = (-b + dsq) / a;
            if (t < t_max && t > t_min) {
                out.point = point_at(ray, t);
                out.normal = (out.point - centre) / s.radius;
                out.t = t;
                out.u = 0.0f;
                out.v = 0.0f;
                return true;
            }
        }
        return false;
    }
    // collision/hitable.rs:39-65 — the enum dispatch for the two arms in scope
    static bool hitable_hit(const Sphere& s, const Ray& ray, float t_min, float t_max, RayHit& out) {
        return s.moving ? moving_sphere_hit(s, ray, t_min, t_max, out) : sphere_hit(s, ray, t_min, t_max, out);
    }
    // collision/hitable_list.rs:40-56
    bool hit_list(const Ray& ray, float t_min, float t_max, RayHit& hit, int32_t& index) const {
        bool any = false;
        float closest = t_max;
        RayHit h;
        for (size_t i = 0; i < spheres.size(); ++i) {
            if (hitable_hit(spheres[i], ray, t_min, closest, h)) {
                any = true;
                closest = h.t;
                hit = h;
                index = (int32_t)i;
            }
        }
        return any;
    }
    // material.rs:169-180 — (u, v) exist only when the material's own texture is an Image; (0, 0) otherwise
    void material_sphere_uv(int32_t sphere_index, V3 normal, float& u, float& v) const {
        u = v = 0.0f;
        const Material& m = materials[spheres[sphere_index].material];
        if ((m.kind == MAT_LAMBERTIAN || m.kind == MAT_DIFFUSE_LIGHT) && textures[m.tex].kind == TEX_IMAGE) get_sphere_uv(normal, u, v);
    }
    // spheres_soa.rs:132-154 epilogue.  NOTE the live path differs: Sphere::ray_hit returns u = v = 0 for every material
    // (sphere.rs:44-45,56-57), so `earth` renders one texel there; the SoA form — the GPU path's spec — asks the material.
    bool soa_epilogue(const Ray& ray, float hit_t, size_t hit_index, RayHit& hit, int32_t& index) const {
        if (hit_index >= soa_len || hit_index >= spheres.size()) return false;
        hit.point = point_at(ray, hit_t);
        V3 centre = v3(cx[hit_index], cy[hit_index], cz[hit_index]);
        hit.normal = (hit.point - centre) * rinv[hit_index];
        hit.t = hit_t;
        material_sphere_uv((int32_t)hit_index, hit.normal, hit.u, hit.v);
        index = (int32_t)hit_index;
        return true;
    }
    // spheres_soa.rs:105-155
    bool hit_soa_scalar(const Ray& ray, float t_min, float t_max, RayHit& hit, int32_t& index) const {
        float hit_t = t_max;
        size_t hit_index = soa_len;
        for (size_t i = 0; i < soa_len; ++i) {
            V3 co = v3(cx[i], cy[i], cz[i]) - ray.origin;
            float nb = dot(co, ray.direction);
            float c = dot(co, co) - rsq[i];
            float discriminant = nb * nb - c;
            if (discriminant > 0.0f) {
                float dsq = std::sqrt(discriminant);
                float t = nb - dsq;
                if (t < t_min) t = nb + dsq;
                if (t > t_min && t < hit_t) {
                    hit_t = t;
                    hit_index = i;
                }
            }
        }
        return soa_epilogue(ray, hit_t, hit_index, hit, index);
    }
#if defined(__AVX2__)
    // spheres_soa.rs:274-391 — 8 spheres per iteration, per-lane running (hit_t, hit_index),
    // horizontal min + lowest matching lane (:354-359).  dot3 = (x*x + y*y) + z*z, no FMA (simd.rs:271-283).
    bool hit_soa_avx2(const Ray& ray, float t_min, float t_max, RayHit& hit, int32_t& index) const {
        const __m256 t_min8 = _mm256_set1_ps(t_min);
        const __m256 ox = _mm256_set1_ps(ray.origin.x), oy = _mm256_set1_ps(ray.origin.y), oz = _mm256_set1_ps(ray.origin.z);
        const __m256 dx = _mm256_set1_ps(ray.direction.x), dy = _mm256_set1_ps(ray.direction.y), dz = _mm256_set1_ps(ray.direction.z);
        __m256 hit_t8 = _mm256_set1_ps(t_max);
        __m256i hit_i8 = _mm256_set1_epi32(-1);
        __m256i idx8 = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7);
        const __m256i eight = _mm256_set1_epi32(8);
        for (size_t i = 0; i < soa_len; i += 8) {
            __m256 cox = _mm256_sub_ps(_mm256_loadu_ps(&cx[i]), ox);
            __m256 coy = _mm256_sub_ps(_mm256_loadu_ps(&cy[i]), oy);
            __m256 coz = _mm256_sub_ps(_mm256_loadu_ps(&cz[i]), oz);
            __m256 nb = _mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(cox, dx), _mm256_mul_ps(coy, dy)), _mm256_mul_ps(coz, dz));
            __m256 cc = _mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(cox, cox), _mm256_mul_ps(coy, coy)), _mm256_mul_ps(coz, coz));
            __m256 c = _mm256_sub_ps(cc, _mm256_loadu_ps(&rsq[i]));
            __m256 disc = _mm256_sub_ps(_mm256_mul_ps(nb, nb), c);
            __m256 pos = _mm256_cmp_ps(disc, _mm256_setzero_ps(), _CMP_GT_OQ);
            if (_mm256_movemask_ps(pos) != 0) {
                __m256 dsq = _mm256_sqrt_ps(disc);
                __m256 t0 = _mm256_sub_ps(nb, dsq);
                __m256 t1 = _mm256_add_ps(nb, dsq);
                __m256 t = _mm256_blendv_ps(t1, t0, _mm256_cmp_ps(t0, t_min8, _CMP_GT_OQ));
                __m256 mask = _mm256_and_ps(pos, _mm256_and_ps(_mm256_cmp_ps(t, t_min8, _CMP_GT_OQ), _mm256_cmp_ps(t, hit_t8, _CMP_LT_OQ)));
                hit_i8 = _mm256_castps_si256(_mm256_blendv_ps(_mm256_castsi256_ps(hit_i8), _mm256_castsi256_ps(idx8), mask));
                hit_t8 = _mm256_blendv_ps(hit_t8, t, mask);
            }
            idx8 = _mm256_add_epi32(idx8, eight);
        }
        alignas(32) float ht[8];
        alignas(32) int32_t hi[8];
        _mm256_store_ps(ht, hit_t8);
        _mm256_store_si256((__m256i*)hi, hit_i8);
        float m = ht[0];
        for (int l = 1; l < 8; ++l) m = std::min(m, ht[l]);
        if (!(m < t_max)) return false;
        for (int l = 0; l < 8; ++l) {
            if (ht[l] == m) {  // lowest lane among equal minima (cttz of the equality mask)
                if (hi[l] < 0) return false;
                return soa_epilogue(ray, m, (size_t)hi[l], hit, index);
            }
        }
        return false;
    }
#endif
    bool ray_hit(int mode, const Ray& ray, float t_min, float t_max, RayHit& hit, int32_t& index) const {
        if (g_recorder && g_recorder->vec) {
            const float v[6] = {ray.origin.x, ray.origin.y, ray.origin.z, ray.direction.x, ray.direction.y, ray.direction.z};
            g_recorder->vec->insert(g_recorder->vec->end(), v, v + 6);
            g_recorder->tvec->push_back(ray.time);
        } else if (g_recorder && g_recorder->n < g_recorder->cap) {
            float* o = g_recorder->out + 6 * g_recorder->n++;
            o[0] = ray.origin.x; o[1] = ray.origin.y; o[2] = ray.origin.z;
            o[3] = ray.direction.x; o[4] = ray.direction.y; o[5] = ray.direction.z;
        }
        bool found;
        switch (mode) {
            case HIT_SOA_SCALAR: found = hit_soa_scalar(ray, t_min, t_max, hit, index); break;
#if defined(__AVX2__)
            case HIT_SOA_AVX2: found = hit_soa_avx2(ray, t_min, t_max, hit, index); break;
#endif
            default: return hit_list(ray, t_min, t_max, hit, index);
        }
        // Hybrid (NOT a reference configuration: SpheresSoA::new panics on a MovingSphere, spheres_soa.rs:49-51).  The GPU
        // path tests static spheres in the SoA form and moving spheres in the live form of moving_sphere.rs:38-73; this
        // restates exactly that so the two can be compared bit for bit.  Nearest hit, lowest index among equal t — what
        // the strict `<` of the in-order list walk (hitable_list.rs:49-54) yields.
        for (int32_t mi : moving_index) {
            RayHit h;
            if (moving_sphere_hit(spheres[mi], ray, t_min, t_max, h) && (!found || h.t < hit.t || (h.t == hit.t && mi < index))) {
                found = true;
                hit = h;
                index = mi;
            }
        }
        return found;
    }

    // material.rs:52-67
    bool scatter_lambertian(const Material& m, const Ray& ray_in, const RayHit& h, Rng& rng, V3& att, Ray& sc) const {
        V3 target = h.point + h.normal + random_unit_vector(rng);
        att = tex_value(m.tex, h.u, h.v, h.point);
        sc.origin = h.point;
        sc.direction = normalize(target - h.point);
        sc.time = ray_in.time;
        return true;
    }
    // material.rs:69-89
    bool scatter_metal(const Material& m, const Ray& ray_in, const RayHit& h, Rng& rng, V3& att, Ray& sc) const {
        V3 reflected = reflect(ray_in.direction, h.normal);
        if (dot(reflected, h.normal) > 0.0f) {
            att = m.albedo;
            sc.origin = h.point;
            sc.direction = normalize(reflected + m.fuzz * random_in_unit_sphere(rng));
            sc.time = ray_in.time;
            return true;
        }
        return false;
    }
    // material.rs:91-124
    bool scatter_dielectric(const Material& m, const Ray& ray_in, const RayHit& h, Rng& rng, V3& att, Ray& sc) const {
        att = v3(1.0f, 1.0f, 1.0f);
        float ref_idx = m.ref_idx;
        float rdotn = dot(ray_in.direction, h.normal);
        V3 outward_normal;
        float ni_over_nt, cosine;
        if (rdotn > 0.0f) {
            cosine = rdotn / length(ray_in.direction);
            cosine = std::sqrt(1.0f - ref_idx * ref_idx * (1.0f - cosine * cosine));
            outward_normal = -h.normal;
            ni_over_nt = ref_idx;
        } else {
            cosine = -rdotn / length(ray_in.direction);
            outward_normal = h.normal;
            ni_over_nt = 1.0f / ref_idx;
        }
        V3 refracted;
        if (refract(ray_in.direction, outward_normal, ni_over_nt, refracted)) {
            float reflect_prob = schlick(cosine, ref_idx);
            if (rng.gen_f32() > reflect_prob) {
                sc.origin = h.point;
                sc.direction = normalize(refracted);
                sc.time = ray_in.time;
                return true;
            }
        }
        sc.origin = h.point;
        sc.direction = normalize(reflect(ray_in.direction, h.normal));
        sc.time = ray_in.time;
        return true;
    }
    // material.rs:138-159
    bool scatter(const Material& m, const Ray& ray_in, const RayHit& h, Rng& rng, V3& att, Ray& sc) const {
        switch (m.kind) {
            case MAT_LAMBERTIAN: return scatter_lambertian(m, ray_in, h, rng, att, sc);
            case MAT_METAL: return scatter_metal(m, ray_in, h, rng, att, sc);
            case MAT_DIELECTRIC: return scatter_dielectric(m, ray_in, h, rng, att, sc);
            default: return false;  // DiffuseLight
        }
    }
    // material.rs:161-167
    V3 emitted(const Material& m, float u, float v, V3 p) const {
        if (m.kind == MAT_DIFFUSE_LIGHT) return tex_value(m.tex, u, v, p);
        return v3(0, 0, 0);
    }
    // scene.rs:39-47
    V3 sky_colour(const Ray& ray) const {
        if (has_sky) return sky;
        float t = 0.5f * (ray.direction.y + 1.0f);
        return splat(1.0f - t) + t * v3(0.5f, 0.7f, 1.0f) * 0.3f;
    }
    // scene.rs:49-71 (recursive, exactly as the reference: emitted + attenuation * L(scattered))
    V3 ray_trace(int mode, const Ray& ray_in, uint32_t depth, uint32_t max_depth, Rng& rng, uint64_t& ray_count) const {
        ray_count += 1;
        RayHit h;
        int32_t idx = -1;
        if (ray_hit(mode, ray_in, 0.001f, std::numeric_limits<float>::max(), h, idx)) {
            const Material& m = materials[spheres[idx].material];
            V3 em = emitted(m, h.u, h.v, h.point);
            if (depth < max_depth) {
                V3 att;
                Ray sc;
                if (scatter(m, ray_in, h, rng, att, sc)) {
                    return em + att * ray_trace(mode, sc, depth + 1, max_depth, rng, ray_count);
                }
            }
            return em;
        }
        return sky_colour(ray_in);
    }
};

// Same estimator as ray_trace, evaluated front to back the way the GPU kernel carries it in registers
// (colour += throughput * emitted; throughput *= attenuation).  Identical draws, identical hits; only the
// association of the final products differs from the recursion (last-ulp).  Used to compare the CUDA path
// with the oracle pixel by pixel; the recursive form above stays the reference-faithful one.
static V3 ray_trace_iterative(const Scene& sc, int mode, Ray ray, uint32_t max_depth, Rng& rng, uint64_t& ray_count) {
    V3 col = v3(0, 0, 0), thr = v3(1.0f, 1.0f, 1.0f);
    for (uint32_t depth = 0;; ++depth) {
        ray_count += 1;
        RayHit h;
        int32_t idx = -1;
        if (!sc.ray_hit(mode, ray, 0.001f, std::numeric_limits<float>::max(), h, idx)) return col + thr * sc.sky_colour(ray);
        const Material& m = sc.materials[sc.spheres[idx].material];
        if (m.kind == MAT_DIFFUSE_LIGHT) return col + thr * sc.emitted(m, h.u, h.v, h.point);
        if (depth >= max_depth) return col;
        V3 att;
        Ray next;
        if (!sc.scatter(m, ray, h, rng, att, next)) return col;
        thr = thr * att;
        ray = next;
    }
}

struct Params {  // params.rs:11-18
    uint32_t width, height, samples, max_depth;
    bool random_seed, use_bvh;
};

// scene.rs:94-116 — one pixel of Scene::update
static inline uint64_t pixel_seed(uint32_t x, uint32_t y, uint32_t frame_num) {
    return ((uint64_t)x * 1973 + (uint64_t)y * 9277 + (uint64_t)frame_num * 26699) | 1;  // scene.rs:99-101
}

static void update_pixel(const Scene& sc, const Camera& cam, const Params& p, int mode, uint32_t frame_num, size_t i,
                         float* out, uint64_t& rays) {
    const float inv_nx = 1.0f / (float)p.width, inv_ny = 1.0f / (float)p.height, inv_ns = 1.0f / (float)p.samples;
    const float mix_prev = (float)frame_num / (float)(frame_num + 1);
    const float mix_new = 1.0f - mix_prev;
    uint32_t y = (uint32_t)i / p.width;
    uint32_t x = (uint32_t)i - y * p.width;
    Rng rng = Rng::seed_from_u64(pixel_seed(x, y, frame_num));
    uint64_t ray_count = 0;
    V3 col = v3(0, 0, 0);
    for (uint32_t s = 0; s < p.samples; ++s) {
        float u = ((float)x + rng.gen_f32()) * inv_nx;
        float v = ((float)y + rng.gen_f32()) * inv_ny;
        Ray ray = get_ray(cam, u, v, rng);
        if (mode & 0x100) {
            // GPU association: the pixel sum absorbs each term as it is produced
            V3 L = ray_trace_iterative(sc, mode & 0xff, ray, p.max_depth, rng, ray_count);
            col = col + L;
        } else {
            col = col + sc.ray_trace(mode, ray, 0, p.max_depth, rng, ray_count);
        }
    }
    col = col * inv_ns;
    out[0] = out[0] * mix_prev + col.x * mix_new;
    out[1] = out[1] * mix_prev + col.y * mix_new;
    out[2] = out[2] * mix_prev + col.z * mix_new;
    rays += ray_count;
}

// scene.rs:73-121 — rayon par_iter_mut over pixels -> std::thread pool pulling row chunks
static uint64_t update(const Scene& sc, const Camera& cam, const Params& p, int mode, uint32_t frame_num, float* buffer,
                       int nthreads, uint32_t row_begin, uint32_t row_end) {
    std::atomic<uint64_t> total{0};
    std::atomic<uint32_t> next_row{row_begin};
    if (nthreads < 1) nthreads = 1;
    auto worker = [&]() {
        uint64_t rays = 0;
        for (;;) {
            uint32_t y = next_row.fetch_add(1);
            if (y >= row_end) break;
            for (uint32_t x = 0; x < p.width; ++x) {
                size_t i = (size_t)y * p.width + x;
                update_pixel(sc, cam, p, mode, frame_num, i, buffer + 3 * i, rays);
            }
        }
        total.fetch_add(rays, std::memory_order_relaxed);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return total.load();
}

// ------------------------------------------------------------------------------------------------
// presets.rs (sphere-only presets) — Storage::new draws the Perlin tables first (storage.rs:28-43)
// ------------------------------------------------------------------------------------------------
static int32_t add_tex_constant(Scene& s, V3 c) {
    Texture t{};
    t.kind = TEX_CONSTANT;
    t.color = c;
    t.odd = t.even = -1;
    s.textures.push_back(t);
    return (int32_t)s.textures.size() - 1;
}
static int32_t add_tex_checker(Scene& s, int32_t odd, int32_t even) {
    Texture t{};
    t.kind = TEX_CHECKER;
    t.odd = odd;
    t.even = even;
    s.textures.push_back(t);
    return (int32_t)s.textures.size() - 1;
}
static int32_t add_tex_noise(Scene& s, float scale) {
    Texture t{};
    t.kind = TEX_NOISE;
    t.scale = scale;
    t.odd = t.even = -1;
    s.textures.push_back(t);
    return (int32_t)s.textures.size() - 1;
}
static int32_t add_mat(Scene& s, int32_t kind, int32_t tex, V3 albedo, float fuzz, float ref_idx) {
    Material m{};
    m.kind = kind;
    m.tex = tex;
    m.albedo = albedo;
    m.fuzz = fuzz;
    m.ref_idx = ref_idx;
    s.materials.push_back(m);
    return (int32_t)s.materials.size() - 1;
}
static void add_sphere(Scene& s, V3 c, float r, int32_t mat) {
    Sphere sp;
    sp.centre = c; sp.radius = r; sp.material = mat;
    s.spheres.push_back(sp);
}
// presets.rs:122-127 + moving_sphere.rs:16-26
static void add_moving_sphere(Scene& s, V3 c0, V3 c1, float t0, float t1, float r, int32_t mat) {
    Sphere sp;
    sp.centre = c0; sp.radius = r; sp.material = mat;
    sp.moving = true;
    sp.centre_delta = c1 - c0;
    sp.time_start = t0;
    sp.inv_time_delta = 1.0f / (t1 - t0);
    s.spheres.push_back(sp);
}

static Camera rtiow_camera(const Params& p, float aperture, float t1) {  // presets.rs:95-109, :275-289
    return camera_new(v3(13, 2, 3), v3(0, 0, 0), v3(0, 1, 0), 20.0f, (float)p.width / (float)p.height, aperture, 10.0f,
                      0.0f, t1);
}

// presets.rs:89-215; only_spheres=true is `random_spheres`, false is `random` (Lambertian spheres move, :150-172);
// `half` = 11 for the reference presets, 158 for stress100k
static void preset_random_spheres(Scene& s, const Params& p, Rng& rng, int half, bool only_spheres = true) {
    s.camera = rtiow_camera(p, 0.1f, 1.0f);
    int32_t odd = add_tex_constant(s, v3(0.2f, 0.3f, 0.1f));
    int32_t even = add_tex_constant(s, v3(0.9f, 0.9f, 0.9f));
    int32_t chk = add_tex_checker(s, odd, even);
    add_sphere(s, v3(0.0f, -1000.0f, 0.0f), 1000.0f, add_mat(s, MAT_LAMBERTIAN, chk, v3(0, 0, 0), 0, 0));
    for (int a = -half; a < half; ++a) {
        for (int b = -half; b < half; ++b) {
            float choose_material = rng.gen_f32();
            float cxv = (float)a + 0.9f * rng.gen_f32();
            float czv = (float)b + 0.9f * rng.gen_f32();
            V3 centre = v3(cxv, 0.2f, czv);
            if (choose_material < 0.8f) {
                V3 centre1 = centre + v3(0.0f, 0.5f * rng.gen_f32(), 0.0f);  // drawn even when only_spheres (presets.rs:150)
                float r0 = rng.gen_f32(); float r1 = rng.gen_f32();
                float g0 = rng.gen_f32(); float g1 = rng.gen_f32();
                float b0 = rng.gen_f32(); float b1 = rng.gen_f32();
                int32_t t = add_tex_constant(s, v3(r0 * r1, g0 * g1, b0 * b1));
                int32_t m = add_mat(s, MAT_LAMBERTIAN, t, v3(0, 0, 0), 0, 0);
                if (only_spheres) add_sphere(s, centre, 0.2f, m);
                else add_moving_sphere(s, centre, centre1, 0.0f, 1.0f, 0.2f, m);
            } else if (choose_material < 0.95f) {
                float r = 0.5f * (1.0f + rng.gen_f32());
                float g = 0.5f * (1.0f + rng.gen_f32());
                float bb = 0.5f * (1.0f + rng.gen_f32());
                float fuzz = 0.5f * rng.gen_f32();
                add_sphere(s, centre, 0.2f, add_mat(s, MAT_METAL, -1, v3(r, g, bb), fuzz, 0));
            } else {
                add_sphere(s, centre, 0.2f, add_mat(s, MAT_DIELECTRIC, -1, v3(0, 0, 0), 0, 1.5f));
            }
        }
    }
    add_sphere(s, v3(0.0f, 1.0f, 0.0f), 1.0f, add_mat(s, MAT_DIELECTRIC, -1, v3(0, 0, 0), 0, 1.5f));
    add_sphere(s, v3(-4.0f, 1.0f, 0.0f), 1.0f,
               add_mat(s, MAT_LAMBERTIAN, add_tex_constant(s, v3(0.4f, 0.2f, 0.1f)), v3(0, 0, 0), 0, 0));
    add_sphere(s, v3(4.0f, 1.0f, 0.0f), 1.0f, add_mat(s, MAT_METAL, -1, v3(0.7f, 0.6f, 0.5f), 0.0f, 0));
}
// presets.rs:217-269
static void preset_small(Scene& s, const Params& p) {
    V3 lookfrom = v3(3, 3, 2), lookat = v3(0, 0, -1);
    s.camera = camera_new(lookfrom, lookat, v3(0, 1, 0), 20.0f, (float)p.width / (float)p.height, 0.1f,
                          length(lookfrom - lookat), 0.0f, 1.0f);
    add_sphere(s, v3(0, 0, -1), 0.5f, add_mat(s, MAT_LAMBERTIAN, add_tex_constant(s, v3(0.1f, 0.2f, 0.5f)), v3(0, 0, 0), 0, 0));
    add_sphere(s, v3(0, -100.5f, -1), 100.0f, add_mat(s, MAT_LAMBERTIAN, add_tex_constant(s, v3(0.8f, 0.8f, 0.0f)), v3(0, 0, 0), 0, 0));
    add_sphere(s, v3(1, 0, -1), 0.5f, add_mat(s, MAT_METAL, -1, v3(0.8f, 0.6f, 0.2f), 0.0f, 0));
    add_sphere(s, v3(-1, 0, -1), 0.5f, add_mat(s, MAT_DIELECTRIC, -1, v3(0, 0, 0), 0, 1.5f));
    add_sphere(s, v3(-1, 0, -1), -0.45f, add_mat(s, MAT_DIELECTRIC, -1, v3(0, 0, 0), 0, 1.5f));
}
// presets.rs:271-315
static void preset_two_perlin_spheres(Scene& s, const Params& p) {
    s.camera = rtiow_camera(p, 0.0f, 0.0f);
    int32_t nt = add_tex_noise(s, 4.0f);
    add_sphere(s, v3(0, -1000, 0), 1000.0f, add_mat(s, MAT_LAMBERTIAN, nt, v3(0, 0, 0), 0, 0));
    add_sphere(s, v3(0, 2, 0), 2.0f, add_mat(s, MAT_LAMBERTIAN, nt, v3(0, 0, 0), 0, 0));
}
// presets.rs:853-930
static void preset_smallpt(Scene& s, const Params& p) {
    s.camera = camera_new(v3(50.0f, 52.0f, 295.6f), v3(50.0f, 33.0f, 0.0f), v3(0, 1, 0), 30.0f,
                          (float)p.width / (float)p.height, 0.05f, 100.0f, 0.0f, 1.0f);
    auto lam = [&](V3 c) { return add_mat(s, MAT_LAMBERTIAN, add_tex_constant(s, c), v3(0, 0, 0), 0, 0); };
    add_sphere(s, v3(1e3f + 1.0f, 40.8f, 81.6f), 1e3f, lam(v3(0.75f, 0.25f, 0.25f)));
    add_sphere(s, v3(-1e3f + 99.0f, 40.8f, 81.6f), 1e3f, lam(v3(0.25f, 0.25f, 0.75f)));
    add_sphere(s, v3(50.0f, 40.8f, 1e3f), 1e3f, lam(v3(0.75f, 0.75f, 0.75f)));
    add_sphere(s, v3(50.0f, 1e3f, 81.6f), 1e3f, lam(v3(0.75f, 0.75f, 0.75f)));
    add_sphere(s, v3(50.0f, -1e3f + 81.6f, 81.6f), 1e3f, lam(v3(0.75f, 0.75f, 0.75f)));
    add_sphere(s, v3(27.0f, 16.5f, 47.0f), 16.5f, add_mat(s, MAT_METAL, -1, v3(1.0f, 1.0f, 1.0f) * 0.999f, 0.0f, 0));
    add_sphere(s, v3(73.0f, 16.5f, 78.0f), 16.5f, add_mat(s, MAT_DIELECTRIC, -1, v3(0, 0, 0), 0, 1.5f));
    add_sphere(s, v3(50.0f, 81.6f - 16.5f, 81.6f), 1.5f,
               add_mat(s, MAT_DIFFUSE_LIGHT, add_tex_constant(s, v3(4.0f, 4.0f, 4.0f) * 100.0f), v3(0, 0, 0), 0, 0));
    s.has_sky = true;
    s.sky = v3(0, 0, 0);
}

static int32_t add_tex_image(Scene& s, int32_t image) {
    Texture t{};
    t.kind = TEX_IMAGE;
    t.odd = t.even = -1;
    t.image = image;
    s.textures.push_back(t);
    return (int32_t)s.textures.size() - 1;
}
// presets.rs:555-594.  The asset (media/earthmap.jpg) is not in the reference tree: the caller supplies the decoded
// pixels (orc_set_earth_image) — what `RgbImage::open` would have produced.
static RgbImage g_earth_image;
static bool preset_earth(Scene& s, const Params& p) {
    if (g_earth_image.data.empty()) return false;
    s.camera = rtiow_camera(p, 0.0f, 0.0f);
    s.images.push_back(g_earth_image);
    int32_t tex = add_tex_image(s, 0);
    add_sphere(s, v3(0, 0, 0), 2.0f, add_mat(s, MAT_LAMBERTIAN, tex, v3(0, 0, 0), 0, 0));
    return true;
}

// offline.rs:16-23 — rng = seed_from_u64(0) (params.rs:21-27); Storage::new (Perlin) first; then the preset.
static Scene* build_preset(const char* name, const Params& p) {
    Scene* s = new Scene();
    Rng rng = Rng::seed_from_u64(0);
    s->perlin.init(rng);
    std::string n(name);
    if (n == "random_spheres") preset_random_spheres(*s, p, rng, 11);
    else if (n == "random") preset_random_spheres(*s, p, rng, 11, false);
    else if (n == "stress100k") preset_random_spheres(*s, p, rng, 158);
    else if (n == "small") preset_small(*s, p);
    else if (n == "two_perlin_spheres") preset_two_perlin_spheres(*s, p);
    else if (n == "smallpt") preset_smallpt(*s, p);
    else if (n == "earth") { if (!preset_earth(*s, p)) { delete s; return nullptr; } }
    else if (n == "final") s->camera = rtiow_camera(p, 0.1f, 1.0f);  // presets.rs:40-71: empty stub, same camera
    else { delete s; return nullptr; }
    s->build_soa();
    return s;
}

}  // namespace orc

// ================================================================================================
// C interface for ctypes (tests / bench cpu_baseline only)
// ================================================================================================
extern "C" {

struct OrcParams {
    uint32_t width, height, samples, max_depth;
    uint32_t random_seed, use_bvh;
};

void* orc_scene_build(const char* preset, const OrcParams* p) {
    orc::Params pp{p->width, p->height, p->samples, p->max_depth, p->random_seed != 0, p->use_bvh != 0};
    return orc::build_preset(preset, pp);
}
void orc_scene_free(void* h) { delete (orc::Scene*)h; }
// the decoded pixels the next orc_scene_build("earth") uses (RGB8, row 0 = top)
void orc_set_earth_image(uint32_t width, uint32_t height, const uint8_t* rgb) {
    orc::g_earth_image.width = width;
    orc::g_earth_image.height = height;
    orc::g_earth_image.data.assign(rgb, rgb + (size_t)width * height * 3);
}
// Append an image / an Image texture / a Checker texture to a scene and point a sphere's material at a texture
// (randomised-scene parity tests of the Image path).  Return the new index.
int32_t orc_scene_add_image(void* h, uint32_t width, uint32_t height, const uint8_t* rgb) {
    auto* s = (orc::Scene*)h;
    orc::RgbImage im;
    im.width = width; im.height = height;
    im.data.assign(rgb, rgb + (size_t)width * height * 3);
    s->images.push_back(std::move(im));
    return (int32_t)s->images.size() - 1;
}
int32_t orc_scene_add_image_texture(void* h, int32_t image) { return orc::add_tex_image(*(orc::Scene*)h, image); }
int32_t orc_scene_add_checker_texture(void* h, int32_t odd, int32_t even) { return orc::add_tex_checker(*(orc::Scene*)h, odd, even); }
void orc_scene_set_sphere_texture(void* h, int32_t sphere, int32_t tex) {
    auto* s = (orc::Scene*)h;
    s->materials[s->spheres[sphere].material].tex = tex;
}
int32_t orc_scene_image_count(void* h) { return (int32_t)((orc::Scene*)h)->images.size(); }
// width/height of image `image`; copies the pixels when rgb_out != nullptr
void orc_scene_image(void* h, int32_t image, uint32_t* width, uint32_t* height, uint8_t* rgb_out) {
    const orc::RgbImage& im = ((orc::Scene*)h)->images[image];
    *width = im.width;
    *height = im.height;
    if (rgb_out) std::memcpy(rgb_out, im.data.data(), im.data.size());
}
// RgbImage::value and get_sphere_uv on their own (unit pins)
void orc_image_value(void* h, int32_t image, float u, float v, float* out3) {
    orc::V3 c = ((orc::Scene*)h)->images[image].value(u, v);
    out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}
void orc_sphere_uv(float nx, float ny, float nz, float* uv2) { orc::get_sphere_uv(orc::v3(nx, ny, nz), uv2[0], uv2[1]); }

// A scene from flat arrays (randomised-scene parity tests): per sphere centre(3)+radius, material kind, colour(3)+fuzz+
// ref_idx (Lambertian/DiffuseLight get a Constant texture of that colour), optional motion rows (centre1(3), time0,
// time1, moving flag) and a Camera::new argument list: lookfrom(3), lookat(3), vup(3), vfov, aspect, aperture,
// focus_dist, time0, time1 (camera.rs:22-32).  sky3 == nullptr means the default gradient (scene.rs:43-46).
void* orc_scene_custom(int32_t n, const float* centre_radius, const int32_t* kind, const float* params5, const float* motion6,
                       const float* cam15, const float* sky3) {
    using namespace orc;
    Scene* s = new Scene();
    Rng rng = Rng::seed_from_u64(0);
    s->perlin.init(rng);
    for (int32_t i = 0; i < n; ++i) {
        const float* p5 = params5 + 5 * i;
        int32_t tex = -1;
        if (kind[i] == MAT_LAMBERTIAN || kind[i] == MAT_DIFFUSE_LIGHT) tex = add_tex_constant(*s, v3(p5[0], p5[1], p5[2]));
        const int32_t m = add_mat(*s, kind[i], tex, v3(p5[0], p5[1], p5[2]), p5[3], p5[4]);
        const V3 c = v3(centre_radius[4 * i], centre_radius[4 * i + 1], centre_radius[4 * i + 2]);
        if (motion6 && motion6[6 * i + 5] != 0.0f)
            add_moving_sphere(*s, c, v3(motion6[6 * i], motion6[6 * i + 1], motion6[6 * i + 2]), motion6[6 * i + 3], motion6[6 * i + 4],
                              centre_radius[4 * i + 3], m);
        else
            add_sphere(*s, c, centre_radius[4 * i + 3], m);
    }
    s->camera = camera_new(v3(cam15[0], cam15[1], cam15[2]), v3(cam15[3], cam15[4], cam15[5]), v3(cam15[6], cam15[7], cam15[8]), cam15[9],
                           cam15[10], cam15[11], cam15[12], cam15[13], cam15[14]);
    if (sky3) {
        s->has_sky = true;
        s->sky = v3(sky3[0], sky3[1], sky3[2]);
    }
    s->build_soa();
    return s;
}
int32_t orc_scene_counts(void* h, int32_t* n_spheres, int32_t* n_materials, int32_t* n_textures) {
    auto* s = (orc::Scene*)h;
    *n_spheres = (int32_t)s->spheres.size();
    *n_materials = (int32_t)s->materials.size();
    *n_textures = (int32_t)s->textures.size();
    return 0;
}
// flat dumps, layouts documented in tests/orc.py
void orc_scene_spheres(void* h, float* centre_radius /*n*4*/, int32_t* material /*n*/) {
    auto* s = (orc::Scene*)h;
    for (size_t i = 0; i < s->spheres.size(); ++i) {
        centre_radius[4 * i + 0] = s->spheres[i].centre.x;
        centre_radius[4 * i + 1] = s->spheres[i].centre.y;
        centre_radius[4 * i + 2] = s->spheres[i].centre.z;
        centre_radius[4 * i + 3] = s->spheres[i].radius;
        material[i] = s->spheres[i].material;
    }
}
// per sphere: centre1 (3), time0, time1, moving flag (as float) — moving_sphere.rs:16-26 inverted to its constructor arguments
void orc_scene_motion(void* h, float* out /*n*6*/) {
    auto* s = (orc::Scene*)h;
    for (size_t i = 0; i < s->spheres.size(); ++i) {
        const orc::Sphere& sp = s->spheres[i];
        orc::V3 c1 = sp.moving ? sp.centre + sp.centre_delta : sp.centre;
        out[6 * i + 0] = c1.x; out[6 * i + 1] = c1.y; out[6 * i + 2] = c1.z;
        out[6 * i + 3] = sp.time_start;
        out[6 * i + 4] = sp.moving ? sp.time_start + 1.0f / sp.inv_time_delta : 0.0f;
        out[6 * i + 5] = sp.moving ? 1.0f : 0.0f;
    }
}
void orc_scene_materials(void* h, int32_t* kind_tex /*n*2*/, float* albedo_fuzz_ref /*n*5*/) {
    auto* s = (orc::Scene*)h;
    for (size_t i = 0; i < s->materials.size(); ++i) {
        const auto& m = s->materials[i];
        kind_tex[2 * i] = m.kind;
        kind_tex[2 * i + 1] = m.tex;
        albedo_fuzz_ref[5 * i + 0] = m.albedo.x;
        albedo_fuzz_ref[5 * i + 1] = m.albedo.y;
        albedo_fuzz_ref[5 * i + 2] = m.albedo.z;
        albedo_fuzz_ref[5 * i + 3] = m.fuzz;
        albedo_fuzz_ref[5 * i + 4] = m.ref_idx;
    }
}
void orc_scene_textures(void* h, int32_t* kind_odd_even /*n*3*/, float* color_scale /*n*4*/) {
    auto* s = (orc::Scene*)h;
    for (size_t i = 0; i < s->textures.size(); ++i) {
        const auto& t = s->textures[i];
        kind_odd_even[3 * i] = t.kind;
        kind_odd_even[3 * i + 1] = t.kind == orc::TEX_IMAGE ? t.image : t.odd;  // Image: the image index rides in the `odd` slot
        kind_odd_even[3 * i + 2] = t.even;
        color_scale[4 * i + 0] = t.color.x;
        color_scale[4 * i + 1] = t.color.y;
        color_scale[4 * i + 2] = t.color.z;
        color_scale[4 * i + 3] = t.scale;
    }
}
void orc_scene_perlin(void* h, float* randvec /*256*3*/, uint32_t* perm_xyz /*3*256*/) {
    auto* s = (orc::Scene*)h;
    for (int i = 0; i < 256; ++i) {
        randvec[3 * i] = s->perlin.randvec[i].x;
        randvec[3 * i + 1] = s->perlin.randvec[i].y;
        randvec[3 * i + 2] = s->perlin.randvec[i].z;
        perm_xyz[i] = s->perlin.perm_x[i];
        perm_xyz[256 + i] = s->perlin.perm_y[i];
        perm_xyz[512 + i] = s->perlin.perm_z[i];
    }
}
// camera as 24 floats: origin, llc, horizontal, vertical, u, v, w (7*3) + time0, time1, lens_radius
void orc_scene_camera(void* h, float* out24) {
    auto* s = (orc::Scene*)h;
    const orc::Camera& c = s->camera;
    const orc::V3* v[7] = {&c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical, &c.u, &c.v, &c.w};
    for (int i = 0; i < 7; ++i) {
        out24[3 * i] = v[i]->x;
        out24[3 * i + 1] = v[i]->y;
        out24[3 * i + 2] = v[i]->z;
    }
    out24[21] = c.time0;
    out24[22] = c.time1;
    out24[23] = c.lens_radius;
}
void orc_scene_sky(void* h, int32_t* has_sky, float* sky3) {
    auto* s = (orc::Scene*)h;
    *has_sky = s->has_sky;
    sky3[0] = s->sky.x;
    sky3[1] = s->sky.y;
    sky3[2] = s->sky.z;
}

// Scene::update over rows [row_begin,row_end) of the image (full image: 0,height). mode: 0 list, 1 soa scalar, 2 soa avx2
uint64_t orc_update(void* h, const OrcParams* p, uint32_t frame_num, float* rgb_inout, int32_t mode, int32_t nthreads,
                    uint32_t row_begin, uint32_t row_end) {
    auto* s = (orc::Scene*)h;
    orc::Params pp{p->width, p->height, p->samples, p->max_depth, p->random_seed != 0, p->use_bvh != 0};
    return orc::update(*s, s->camera, pp, mode, frame_num, rgb_inout, nthreads, row_begin, std::min(row_end, p->height));
}
int32_t orc_has_avx2() {
#if defined(__AVX2__)
    return 1;
#else
    return 0;
#endif
}

// --- small probes used by the pinning tests ---
void orc_rng_seed(uint64_t seed, uint64_t* state4) {
    orc::Rng r = orc::Rng::seed_from_u64(seed);
    std::memcpy(state4, r.s, 32);
}
void orc_rng_u64(uint64_t* state4, uint64_t* out, int32_t n) {
    orc::Rng r;
    std::memcpy(r.s, state4, 32);
    for (int i = 0; i < n; ++i) out[i] = r.next_u64();
    std::memcpy(state4, r.s, 32);
}
void orc_rng_f32(uint64_t* state4, float* out, int32_t n) {
    orc::Rng r;
    std::memcpy(r.s, state4, 32);
    for (int i = 0; i < n; ++i) out[i] = r.gen_f32();
    std::memcpy(state4, r.s, 32);
}
void orc_sincos(const float* x, float* s, float* c, int32_t n) {
    for (int i = 0; i < n; ++i) orc::sinf_cosf(x[i], s[i], c[i]);
}
float orc_turb(void* h, float x, float y, float z) { return ((orc::Scene*)h)->perlin.turb(orc::v3(x, y, z)); }
float orc_noise(void* h, float x, float y, float z) { return ((orc::Scene*)h)->perlin.noise(orc::v3(x, y, z)); }
void orc_tex_value(void* h, int32_t tex, float x, float y, float z, float* out3) {
    orc::V3 v = ((orc::Scene*)h)->tex_value(tex, 0.f, 0.f, orc::v3(x, y, z));
    out3[0] = v.x; out3[1] = v.y; out3[2] = v.z;
}
void orc_srgb(const float* rgb, uint8_t* out, int32_t npix) {
    for (int i = 0; i < npix; ++i) orc::linear_to_srgb(rgb + 3 * i, out + 3 * i);
}
// nearest hit for explicit rays (o,d as 6 floats each); returns index (-1 miss) and t
void orc_hit(void* h, int32_t mode, const float* rays6, int32_t n, int32_t* idx_out, float* t_out) {
    auto* s = (orc::Scene*)h;
    for (int i = 0; i < n; ++i) {
        orc::Ray r;
        r.origin = orc::v3(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]);
        r.direction = orc::v3(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]);
        r.time = 0;
        orc::RayHit hit;
        int32_t idx = -1;
        bool ok = s->ray_hit(mode, r, 0.001f, std::numeric_limits<float>::max(), hit, idx);
        idx_out[i] = ok ? idx : -1;
        t_out[i] = ok ? hit.t : std::numeric_limits<float>::max();
    }
}
// src/bench.rs:17-26 fixture: scene rng continues after the preset; ray = camera.get_ray(0.5, 0.5, rng).
// The oracle's build_preset does not keep the rng, so the fixture re-derives it: same draws, same order.
void orc_bench_fixture_ray(const OrcParams* p, float* ray6) {
    orc::Params pp{p->width, p->height, p->samples, p->max_depth, false, false};
    orc::Scene s;
    orc::Rng rng = orc::Rng::seed_from_u64(0);
    s.perlin.init(rng);
    orc::preset_random_spheres(s, pp, rng, 11);
    orc::Ray r = orc::get_ray(s.camera, 0.5f, 0.5f, rng);
    ray6[0] = r.origin.x; ray6[1] = r.origin.y; ray6[2] = r.origin.z;
    ray6[3] = r.direction.x; ray6[4] = r.direction.y; ray6[5] = r.direction.z;
}
// first f32 the scene rng yields after building `random_spheres` (SURVEY Appendix B cross-check)
float orc_next_f32_after_random_spheres() {
    orc::Params pp{200, 100, 10, 10, false, false};
    orc::Scene s;
    orc::Rng rng = orc::Rng::seed_from_u64(0);
    s.perlin.init(rng);
    orc::preset_random_spheres(s, pp, rng, 11);
    return rng.gen_f32();
}
// debug: trace one pixel and record its rays in order; returns the number of rays recorded
int64_t orc_trace_pixel_rays(void* h, const OrcParams* p, uint32_t x, uint32_t y, float* rays6, int64_t cap) {
    auto* s = (orc::Scene*)h;
    orc::Params pp{p->width, p->height, p->samples, p->max_depth, false, false};
    orc::RayRecorder rec;
    rec.out = rays6;
    rec.cap = cap;
    orc::g_recorder = &rec;
    float px[3] = {0, 0, 0};
    uint64_t rays = 0;
    orc::update_pixel(*s, s->camera, pp, orc::HIT_SOA_SCALAR | 0x100, 0, (size_t)y * p->width + x, px, rays);
    orc::g_recorder = nullptr;
    return rec.n;
}
// every ray the update of pixels [pix0, pix1) (row-major indices) hands to ray_hit, in trace order pixel after pixel, with its
// ray.time; at most `cap` rays are written.  Threads record disjoint pixel ranges; the result does not depend on nthreads.
int64_t orc_record_rays(void* h, const OrcParams* p, uint32_t frame_num, uint64_t pix0, uint64_t pix1, float* rays6, float* times, int64_t cap,
                        int32_t nthreads) {
    auto* s = (orc::Scene*)h;
    const orc::Params pp{p->width, p->height, p->samples, p->max_depth, false, false};
    nthreads = std::max(1, nthreads);
    std::vector<std::vector<float>> rv(nthreads), tv(nthreads);
    std::vector<std::thread> th;
    const uint64_t n = pix1 > pix0 ? pix1 - pix0 : 0;
    for (int k = 0; k < nthreads; ++k)
        th.emplace_back([&, k] {
            orc::RayRecorder rec;
            rec.vec = &rv[k];
            rec.tvec = &tv[k];
            orc::g_recorder = &rec;
            for (uint64_t i = pix0 + n * k / nthreads; i < pix0 + n * (k + 1) / nthreads; ++i) {
                float px[3] = {0, 0, 0};
                uint64_t rays = 0;
                orc::update_pixel(*s, s->camera, pp, orc::HIT_SOA_SCALAR | 0x100, frame_num, (size_t)i, px, rays);
            }
            orc::g_recorder = nullptr;
        });
    for (auto& t : th) t.join();
    int64_t out = 0;
    for (int k = 0; k < nthreads && out < cap; ++k) {
        const int64_t take = std::min<int64_t>((int64_t)tv[k].size(), cap - out);
        std::memcpy(rays6 + 6 * out, rv[k].data(), (size_t)take * 6 * sizeof(float));
        if (times) std::memcpy(times + out, tv[k].data(), (size_t)take * sizeof(float));
        out += take;
    }
    return out;
}
// orc_hit with per-ray times (Hitable::MovingSphere) and threads
void orc_hit_times(void* h, int32_t mode, const float* rays6, const float* times, int64_t n, int32_t* idx_out, float* t_out, int32_t nthreads) {
    auto* s = (orc::Scene*)h;
    nthreads = std::max(1, nthreads);
    std::vector<std::thread> th;
    for (int k = 0; k < nthreads; ++k)
        th.emplace_back([&, k] {
            for (int64_t i = n * k / nthreads; i < n * (k + 1) / nthreads; ++i) {
                orc::Ray r;
                r.origin = orc::v3(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]);
                r.direction = orc::v3(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]);
                r.time = times ? times[i] : 0.0f;
                orc::RayHit hit;
                int32_t idx = -1;
                const bool ok = s->ray_hit(mode, r, 0.001f, std::numeric_limits<float>::max(), hit, idx);
                idx_out[i] = ok ? idx : -1;
                t_out[i] = ok ? hit.t : std::numeric_limits<float>::max();
            }
        });
    for (auto& t : th) t.join();
}
int32_t orc_hw_threads() { return (int32_t)std::thread::hardware_concurrency(); }

}  // extern "C"
