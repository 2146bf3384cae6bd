// pt_shade.cuh — everything a path does between two sphere sweeps: camera ray generation,
// sampling, materials, textures, Perlin noise, sky.  Device side of
//   src/camera.rs:56-68 (get_ray)          src/math.rs:6-34,61-80 (samplers, reflect/refract/schlick)
//   src/material.rs:52-124,138-167          src/texture.rs:74-91
//   src/perlin.rs:54-111                    src/scene.rs:39-47 (sky)       src/simd.rs:107-208 (sinf_cosf)
//
// The translation unit is compiled with -fmad=false: every expression here rounds exactly like the
// reference's unfused f32 arithmetic (IEEE sqrt and division are nvcc's defaults), so a path only
// leaves the reference's trajectory when a last-ulp difference in sinf()/x^5 flips a branch.
// Fused arithmetic is used only where it is written explicitly (the sweep's pre-filter, pt_sweep.cuh).
#pragma once
#include <stdint.h>
#include <float.h>
#include "pt_rng.cuh"

#ifndef PT_PERLIN_SELECT
#define PT_PERLIN_SELECT 1
#endif

namespace pt {

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
// glam: dot = (x*x + y*y) + z*z ; normalize = v * (1 / length)
__device__ __forceinline__ float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 normalize(V3 a) { return a * (1.0f / length(a)); }

// ---- device-side scene records (built by the host flattener in ptgpu.cu) ----
enum : int32_t { TEX_CONSTANT = 0, TEX_CHECKER = 1, TEX_NOISE = 2, TEX_IMAGE = 3 };
enum : int32_t { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3 };

struct __align__(16) DevTexture {  // 32 B
    float r, g, b;   // Constant colour
    float scale;     // Noise
    int32_t kind;
    int32_t odd, even;  // Checker: child texture indices; Image: width, height
    int32_t offset;     // Image: byte offset of the RGB8 pixels in the scene's image pool
};

// one record per sphere, read only when that sphere is the nearest hit
struct __align__(16) DevShade {  // 32 B
    float ar, ag, ab;  // Lambertian/DiffuseLight with a Constant texture: the colour; Metal: albedo
    float param;       // Metal: fuzz; Dielectric: ref_idx
    float rinv;        // 1 / radius (signed) — spheres_soa.rs:47
    int32_t kind;      // MAT_*
    int32_t tex;       // texture index when the texture is not Constant, else -1
    int32_t moving;    // 1: Hitable::MovingSphere (centre and normal follow ray.time, DevMotion in pt_sweep.cuh)
};

struct DevCamera {  // src/camera.rs:8-19 minus `w`
    V3 origin, llc, horizontal, vertical, u, v;
    float time0, time1, lens_radius;
};

// Perlin tables staged in shared memory: randvec as float4 (one LDS.128 per corner), perms as bytes.
struct PerlinSmem {
    float4 randvec[256];
    uint8_t perm_x[256], perm_y[256], perm_z[256];
};

// ---- src/simd.rs:120-208: Cephes sinf/cosf, scalar restatement of the SSE2 lane ----
__device__ __forceinline__ void sinf_cosf_cephes(float xin, float& s_out, float& c_out) {
    uint32_t xb = __float_as_uint(xin);
    uint32_t sign_bit_sin = xb & 0x80000000u;
    float x = __uint_as_float(xb & 0x7fffffffu);
    float y = x * 1.27323954473516f;
    int32_t j = (int32_t)y;  // truncation, like cvttps
    j = (j + 1) & ~1;
    y = (float)j;
    const uint32_t swap_sign_bit_sin = ((uint32_t)(j & 4)) << 29;
    const bool poly_mask = (j & 2) == 0;
    x = x + y * -0.78515625f;
    x = x + y * -2.4187564849853515625e-4f;
    x = x + y * -3.77489497744594108e-8f;
    const uint32_t sign_bit_cos = ((~(uint32_t)(j - 2)) & 4u) << 29;
    sign_bit_sin ^= swap_sign_bit_sin;
    const float z = x * x;
    float yc = 2.443315711809948E-005f;
    yc = yc * z;
    yc = yc + -1.388731625493765E-003f;
    yc = yc * z;
    yc = yc + 4.166664568298827E-002f;
    yc = yc * z;
    yc = yc * z;
    yc = yc - z * 0.5f;
    yc = yc + 1.0f;
    float ys = -1.9515295891E-4f;
    ys = ys * z;
    ys = ys + 8.3321608736E-3f;
    ys = ys * z;
    ys = ys + -1.6666654611E-1f;
    ys = ys * z;
    ys = ys * x;
    ys = ys + x;
    const float sinv = poly_mask ? ys : yc;
    const float cosv = poly_mask ? yc : ys;
    s_out = __uint_as_float(__float_as_uint(sinv) ^ sign_bit_sin);
    c_out = __uint_as_float(__float_as_uint(cosv) ^ sign_bit_cos);
}

// ---- src/math.rs ----
__device__ __forceinline__ V3 random_in_unit_disk(Rng& rng) {  // math.rs:6-13
    for (;;) {
        const float a = rng_f32(rng);
        const float b = rng_f32(rng);
        const V3 p = 2.0f * v3(a, b, 0.0f) - v3(1.0f, 1.0f, 0.0f);
        if (dot(p, p) < 1.0f) return p;
    }
}
__device__ __forceinline__ V3 random_in_unit_sphere(Rng& rng) {  // math.rs:15-26
    for (;;) {
        const float a = 2.0f * rng_f32(rng) - 1.0f;
        const float b = 2.0f * rng_f32(rng) - 1.0f;
        const float c = 2.0f * rng_f32(rng) - 1.0f;
        const V3 p = v3(a, b, c);
        if (dot(p, p) < 1.0f) return p;
    }
}
__device__ __forceinline__ V3 random_unit_vector(Rng& rng) {  // math.rs:28-34
    const float z = rng_f32(rng) * 2.0f - 1.0f;
    const float a = rng_f32(rng) * 2.0f * 3.14159265358979323846f;
    const float r = sqrtf(1.0f - z * z);
    float sina, cosa;
    sinf_cosf_cephes(a, sina, cosa);
    return v3(r * cosa, r * sina, z);
}
__device__ __forceinline__ V3 reflect(V3 v, V3 n) { return v - 2.0f * dot(v, n) * n; }  // math.rs:61-63
__device__ __forceinline__ bool refract(V3 v, V3 n, float ni_over_nt, V3& out) {          // math.rs:65-73
    const float dt = dot(v, n);
    const float discriminant = 1.0f - (ni_over_nt * ni_over_nt) * (1.0f - (dt * dt));
    if (discriminant > 0.0f) {
        out = ni_over_nt * (v - n * dt) - n * sqrtf(discriminant);
        return true;
    }
    return false;
}
__device__ __forceinline__ float schlick(float cosine, float ref_idx) {  // math.rs:76-80
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    r0 = r0 * r0;
    // `(1.0 - cosine).powf(5.0)`: Rust calls the platform's powf.  x^5 through a double-precision multiply chain, rounded once,
    // is the correctly rounded power; glibc's powf (<= 0.52 ulp) returns the same float for 99.93 % of x in [0, 1] and the
    // neighbouring one otherwise (measured over 4*10^5 arguments), and the value is only compared against one 24-bit uniform
    // draw (material.rs:109): a coin flip differs about once per 10^10 evaluations.  (The f32 chain x2*x2*x used in round 1
    // was off by up to 3 ulp on half of the arguments.)  Dielectric hits are ~4 % of the rays: the four DMULs do not show.
    const double x = (double)(1.0f - cosine);
    const double x2 = x * x;
    const float x5 = (float)(x2 * x2 * x);
    return r0 + (1.0f - r0) * x5;
}

// ---- src/perlin.rs ----
__device__ __forceinline__ float perlin_noise(const PerlinSmem& P, V3 p) {  // perlin.rs:89-111 + :54-74
    const float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    const float u = p.x - fx, v = p.y - fy, w = p.z - fz;
    // Rust `as usize` saturates negatives/NaN to 0; __float2uint_rz does the same for u32 and the
    // index is masked to 8 bits (floats >= 2^32 have their low 8 bits zero in both widths, and so does
    // the saturated usize::MAX+1 wrap: (MAX + 1) & 255 == 0 while MAX & 255 == 255 — see below).
    uint32_t i = __float2uint_rz(fx), j = __float2uint_rz(fy), k = __float2uint_rz(fz);
    // For |coordinate| >= 2^32 Rust yields the exact integer (low 8 bits are 0 for f32 >= 2^32) or, when
    // saturated (>= 2^64), usize::MAX (low bits 255).  u32 saturation gives 0xffffffff (low bits 255) for
    // everything >= 2^32; patch the non-saturated band so both agree.
    if (fx >= 4294967296.0f && fx < 18446744073709551616.0f) i = 0u;
    if (fy >= 4294967296.0f && fy < 18446744073709551616.0f) j = 0u;
    if (fz >= 4294967296.0f && fz < 18446744073709551616.0f) k = 0u;
    const uint32_t px0 = P.perm_x[i & 255u], px1 = P.perm_x[(i + 1u) & 255u];
    const uint32_t py0 = P.perm_y[j & 255u], py1 = P.perm_y[(j + 1u) & 255u];
    const uint32_t pz0 = P.perm_z[k & 255u], pz1 = P.perm_z[(k + 1u) & 255u];
    const float uu = u * u * (3.0f - 2.0f * u);
    const float vv = v * v * (3.0f - 2.0f * v);
    const float ww = w * w * (3.0f - 2.0f * w);
    float accum = 0.0f;
#pragma unroll
    for (int di = 0; di < 2; ++di) {
#pragma unroll
        for (int dj = 0; dj < 2; ++dj) {
#pragma unroll
            for (int dk = 0; dk < 2; ++dk) {
                const uint32_t idx = (di ? px1 : px0) ^ (dj ? py1 : py0) ^ (dk ? pz1 : pz0);
                const float4 c = P.randvec[idx];
                const float ii = (float)di, jj = (float)dj, kk = (float)dk;
                const V3 weight = v3(u - ii, v - jj, w - kk);
#if PT_PERLIN_SELECT
                // i*uu + (1-i)*(1-uu) with i in {0, 1} is uu or 1-uu exactly (0*x = +0 for the finite x in [0, 1] that reach
                // here, and y + 0 = y); spelled as a select it saves the 0*x products the compiler must otherwise keep
                // (-fmad=false, no fast-math: ~26 SASS instructions per noise call).  Validated on hardware against the product
                // form (-DPT_PERLIN_SELECT=0): bit-identical images (profiles/variants_r2_*.txt).
                accum += (di ? uu : 1.0f - uu) * (dj ? vv : 1.0f - vv) * (dk ? ww : 1.0f - ww) * dot(v3(c.x, c.y, c.z), weight);
#else
                accum += (ii * uu + (1.0f - ii) * (1.0f - uu)) * (jj * vv + (1.0f - jj) * (1.0f - vv)) *
                         (kk * ww + (1.0f - kk) * (1.0f - ww)) * dot(v3(c.x, c.y, c.z), weight);
#endif
            }
        }
    }
    return accum;
}
__device__ __forceinline__ float perlin_turb(const PerlinSmem& P, V3 p) {  // perlin.rs:76-87
    float accum = 0.0f;
    V3 temp_p = p;
    float weight = 1.0f;
#pragma unroll 1
    for (int d = 0; d < 7; ++d) {
        accum += weight * perlin_noise(P, temp_p);
        weight *= 0.5f;
        temp_p = temp_p * 2.0f;
    }
    return fabsf(accum);
}

// ---- src/material.rs:41-49: get_sphere_uv.  `x.atan2(y)` is atan2(x, y) — the reference's argument order. ----
__device__ __forceinline__ void get_sphere_uv(V3 normal, float& u, float& v) {
    const float phi = atan2f(normal.x, normal.y);
    const float theta = asinf(normal.y);
    u = 1.0f - (phi + 3.14159265358979323846f) * (1.0f / (2.0f * 3.14159265358979323846f));
    v = (theta + 1.57079632679489661923f) * 0.318309886183790671538f;
}
// ---- src/texture.rs:27-36: RgbImage::value.  Rust `as i32` saturates and maps NaN to 0, like cvt.rzi.s32.f32. ----
__device__ __noinline__ V3 image_value(const uint8_t* __restrict__ pixels, int32_t width, int32_t height, float u, float v) {
    int32_t i = __float2int_rz(u * (float)width);
    int32_t j = __float2int_rz((1.0f - v) * (float)height - 0.001f);
    i = min(max(i, 0), width - 1);
    j = min(max(j, 0), height - 1);
    const uint8_t* t = pixels + 3 * (size_t)i + 3 * (size_t)width * (size_t)j;
    return v3((float)__ldg(t) / 255.0f, (float)__ldg(t + 1) / 255.0f, (float)__ldg(t + 2) / 255.0f);
}

// what texture_value needs beyond the texture table: the scene's image pool, and whether this hit carries sphere (u, v)
// (Sphere via the SoA epilogue spheres_soa.rs:141 does; MovingSphere::ray_hit returns u = v = 0, moving_sphere.rs:53-54)
struct TexCtx {
    const DevTexture* __restrict__ tex;
    const uint8_t* __restrict__ images;
    bool sphere_uv;
};

// ---- src/texture.rs:74-91 ---- (Checker recursion unrolled into a bounded walk: the arena graph
// is a DAG of at most n_textures nodes; 16 levels is far beyond any preset).  `normal` feeds get_sphere_uv, which the
// reference evaluates only when the material's OWN texture is an Image (material.rs:169-180: level 0 here); an Image
// reached through a Checker sees (u, v) = (0, 0).
__device__ __forceinline__ V3 texture_value(const TexCtx& tc, const PerlinSmem& P, int32_t ti, V3 p, V3 normal) {
    const DevTexture* __restrict__ tex = tc.tex;
#pragma unroll 1
    for (int level = 0; level < 16; ++level) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(tex + ti));
        const int4 b = __ldg(reinterpret_cast<const int4*>(tex + ti) + 1);
        if (b.x == TEX_CONSTANT) return v3(a.x, a.y, a.z);
        if (b.x == TEX_IMAGE) {
            float u = 0.0f, v = 0.0f;
            if (level == 0 && tc.sphere_uv) get_sphere_uv(normal, u, v);
            return image_value(tc.images + (uint32_t)b.w, b.y, b.z, u, v);
        }
        if (b.x == TEX_CHECKER) {
            const V3 s = v3(10.0f, 10.0f, 10.0f) * p;
            const float sines = sinf(s.x) * sinf(s.y) * sinf(s.z);
            ti = sines < 0.0f ? b.y : b.z;
            continue;
        }
        // Noise
        const float g = 0.5f * (1.0f + sinf(a.w * p.z + 10.0f * perlin_turb(P, p)));
        return v3(g, g, g);  // vec3(1,1,1) * 0.5 * (...) : 1*0.5 == 0.5 exactly
    }
    return v3(0.0f, 0.0f, 0.0f);
}

// ---- src/scene.rs:39-47 ----
__device__ __forceinline__ V3 sky_colour(bool has_sky, V3 sky, V3 dir) {
    if (has_sky) return sky;
    const float t = 0.5f * (dir.y + 1.0f);
    const float a = 1.0f - t;
    return v3(a, a, a) + t * v3(0.5f, 0.7f, 1.0f) * 0.3f;
}

// ---- src/camera.rs:56-68 ----
__device__ __forceinline__ void camera_get_ray(const DevCamera& c, float s, float t, Rng& rng, V3& o, V3& d, float& time) {
    const V3 rd = c.lens_radius * random_in_unit_disk(rng);
    const V3 offset = c.u * rd.x + c.v * rd.y;
    time = c.time0 + rng_f32(rng) * (c.time1 - c.time0);
    o = c.origin + offset;
    d = normalize(c.llc + s * c.horizontal + t * c.vertical - c.origin - offset);
}

// ---- src/material.rs:138-159: scatter.  Returns false when the path is absorbed. ----
__device__ __forceinline__ bool material_scatter(const DevShade& m, const TexCtx& tex, const PerlinSmem& P,
                                                 V3 rd, V3 point, V3 normal, Rng& rng, V3& attenuation, V3& scattered) {
    if (m.kind == MAT_LAMBERTIAN) {  // material.rs:52-67
        const V3 target = point + normal + random_unit_vector(rng);
        attenuation = m.tex < 0 ? v3(m.ar, m.ag, m.ab) : texture_value(tex, P, m.tex, point, normal);
        scattered = normalize(target - point);
        return true;
    }
    if (m.kind == MAT_METAL) {  // material.rs:69-89
        const V3 reflected = reflect(rd, normal);
        if (dot(reflected, normal) > 0.0f) {
            attenuation = v3(m.ar, m.ag, m.ab);
            scattered = normalize(reflected + m.param * random_in_unit_sphere(rng));
            return true;
        }
        return false;
    }
    if (m.kind == MAT_DIELECTRIC) {  // material.rs:91-124
        const float ref_idx = m.param;
        attenuation = v3(1.0f, 1.0f, 1.0f);
        const float rdotn = dot(rd, normal);
        V3 outward_normal;
        float ni_over_nt, cosine;
        if (rdotn > 0.0f) {
            cosine = rdotn / length(rd);
            cosine = sqrtf(1.0f - ref_idx * ref_idx * (1.0f - cosine * cosine));
            outward_normal = -normal;
            ni_over_nt = ref_idx;
        } else {
            cosine = -rdotn / length(rd);
            outward_normal = normal;
            ni_over_nt = 1.0f / ref_idx;
        }
        V3 refracted;
        if (refract(rd, outward_normal, ni_over_nt, refracted)) {
            const float reflect_prob = schlick(cosine, ref_idx);
            if (rng_f32(rng) > reflect_prob) {
                scattered = normalize(refracted);
                return true;
            }
        }
        scattered = normalize(reflect(rd, normal));
        return true;
    }
    return false;  // DiffuseLight: material.rs:157
}

}  // namespace pt
