// probe_mma.cu — can the warp-level tensor path (mma.sync, legacy HMMA pipe on sm_100a) carry the pre-filter's dot products?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o probe_mma probe_mma.cu
//
// The pre-filter (pt_sweep.cuh) is two dot products per (sphere, ray) test, A = [cx cy cz 1 k].[dx dy dz nod 0] and
// B' = [cx cy cz 1 k].[2ox 2oy 2oz -oo 1], followed by L' = A*A + B' > 0.  With every f32 operand split into two TF32
// pieces (hi = rna(x), lo = rna(x - hi)) and the three products hi*hi + hi*lo + lo*hi, K = 15 -> 16: two m16n8k8 TF32
// MMAs per dot product per 16 spheres x 8 rays.  This probe measures
//   (1) the issue rate of mma.sync m16n8k8 tf32 (and m16n8k16 bf16 for reference), clk per MMA per SM sub-partition,
//   (2) one full "step" of the would-be loop: 2 non-broadcast LDS.128 sphere fragments, 16 MMAs (4 ray groups x {A, B'} x
//       2 k-steps), 8 FFMA2 (L'), 8 FMNMX3-equivalents, one vote — 512 tests — in clk per 32 tests per SMSP (the unit of
//       tools/probe_sweep2.cu: the shipped FP32 loop runs at 11.0-11.5),
//   (3) the error of the split evaluation against f64 on RTIOW-like operands, normalised by |c|^2 + r^2 + |o|^2 (the
//       quantity the pre-filter's slack is proportional to): what slack a tensor-core pre-filter would need.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
typedef unsigned long long u64;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32_z(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {  // D = A*B (zero accumulator)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(0.0f));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*(u64*)&a), "l"(*(u64*)&b), "l"(*(u64*)&c)); return *(float2*)&d; }
__device__ __forceinline__ uint32_t tf32_rna(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }

constexpr int CH = 8;
template <int KIND>
__global__ void __launch_bounds__(256) k_rate(float* out, const uint32_t* in, int iters) {
    uint32_t a[4], b[2];
    float d[CH][4];
    for (int i = 0; i < 4; ++i) a[i] = in[threadIdx.x + i];
    for (int i = 0; i < 2; ++i) b[i] = in[threadIdx.x + 7 + i];
    for (int c = 0; c < CH; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.0f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) { if (KIND == 0) mma_tf32(d[c], a, b); else mma_bf16(d[c], a, b); }
    }
    float r = 0;
    for (int c = 0; c < CH; ++c) for (int i = 0; i < 4; ++i) r += d[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// one step of the would-be sweep loop per iteration: 16 spheres x 32 rays
//   sphere fragments: 2 x LDS.128 per lane, lane-specific addresses (fragment order), advancing through a 32 KB image
//   ray fragments: 4 ray groups x {A, B'} x 2 k-steps x 2 registers = 32 loop-invariant registers
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_step(float* out, const uint32_t* in, const float4* img, int n_steps, int trips) {
    extern __shared__ float4 simg[];  // n_steps x 2 x 32 float4
    for (int i = threadIdx.x; i < n_steps * 64; i += blockDim.x) simg[i] = img[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t rb[4][2][2][2];
    for (int q = 0; q < 4; ++q) for (int w = 0; w < 2; ++w) for (int ks = 0; ks < 2; ++ks) for (int i = 0; i < 2; ++i) rb[q][w][ks][i] = in[threadIdx.x + q * 8 + w * 4 + ks * 2 + i];
    unsigned flagged = 0u;
    for (int t = 0; t < trips; ++t) {
        float runmax = -3.0e38f;
        const float4* p = simg + lane;
#pragma unroll 1
        for (int s = 0; s < n_steps; ++s, p += 64) {
            const float4 f0 = p[0], f1 = p[32];
            const uint32_t a0[4] = {__float_as_uint(f0.x), __float_as_uint(f0.y), __float_as_uint(f0.z), __float_as_uint(f0.w)};
            const uint32_t a1[4] = {__float_as_uint(f1.x), __float_as_uint(f1.y), __float_as_uint(f1.z), __float_as_uint(f1.w)};
            float m = -3.0e38f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float A[4], B[4];
                mma_tf32_z(A, a0, rb[q][0][0]);
                mma_tf32(A, a1, rb[q][0][1]);
                mma_tf32_z(B, a0, rb[q][1][0]);
                mma_tf32(B, a1, rb[q][1][1]);
                const float2 L0 = fma2(make_float2(A[0], A[1]), make_float2(A[0], A[1]), make_float2(B[0], B[1]));
                const float2 L1 = fma2(make_float2(A[2], A[3]), make_float2(A[2], A[3]), make_float2(B[2], B[3]));
                m = fmaxf(fmaxf(L0.x, L0.y), m);
                m = fmaxf(fmaxf(L1.x, L1.y), m);
            }
            if (__any_sync(0xffffffffu, m > 0.0f)) { flagged += 1u; runmax = fmaxf(runmax, m); }
        }
        rb[0][0][0][0] ^= (t & 1) << 13;  // a slightly different ray next trip: nothing hoisted
        if (runmax == 12345.0f) flagged += 7u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)flagged;
}

// accuracy: one warp = 16 spheres x 8 rays, the real operand layout
//   S = [cx cy cz 1 k], R_A = [dx dy dz nod 0], R_B = [2ox 2oy 2oz -oo 1];  K index: 0..4 hi*hi, 5..9 hi*lo, 10..14 lo*hi, 15 zero
__global__ void k_acc(const float* S, const float* RA, const float* RB, float* outA, float* outB, int n_tiles) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < n_tiles; tile += gridDim.x * (blockDim.x >> 5)) {
        const float* s = S + (size_t)tile * 16 * 5;
        const float* ra = RA + (size_t)tile * 8 * 5;
        const float* rbv = RB + (size_t)tile * 8 * 5;
        auto s_elem = [&](int row, int k) -> uint32_t {  // sphere operand element (row, k)
            if (k == 15) return 0u;
            const int piece = k / 5, e = k % 5;
            const float x = s[row * 5 + e];
            const uint32_t hi = tf32_rna(x);
            if (piece < 2) return hi;
            return tf32_rna(x - __uint_as_float(hi));
        };
        auto r_elem = [&](const float* r, int col, int k) -> uint32_t {
            if (k == 15) return 0u;
            const int piece = k / 5, e = k % 5;
            const float x = r[col * 5 + e];
            const uint32_t hi = tf32_rna(x);
            if (piece != 1) return hi;
            return tf32_rna(x - __uint_as_float(hi));
        };
        float A[4] = {0, 0, 0, 0}, B[4] = {0, 0, 0, 0};
        for (int ks = 0; ks < 2; ++ks) {
            const uint32_t a[4] = {s_elem(g, ks * 8 + t), s_elem(g + 8, ks * 8 + t), s_elem(g, ks * 8 + t + 4), s_elem(g + 8, ks * 8 + t + 4)};
            const uint32_t ba[2] = {r_elem(ra, g, ks * 8 + t), r_elem(ra, g, ks * 8 + t + 4)};
            const uint32_t bb[2] = {r_elem(rbv, g, ks * 8 + t), r_elem(rbv, g, ks * 8 + t + 4)};
            mma_tf32(A, a, ba);
            mma_tf32(B, a, bb);
        }
        float* oa = outA + (size_t)tile * 128;
        float* ob = outB + (size_t)tile * 128;
        oa[g * 8 + 2 * t] = A[0]; oa[g * 8 + 2 * t + 1] = A[1]; oa[(g + 8) * 8 + 2 * t] = A[2]; oa[(g + 8) * 8 + 2 * t + 1] = A[3];
        ob[g * 8 + 2 * t] = B[0]; ob[g * 8 + 2 * t + 1] = B[1]; ob[(g + 8) * 8 + 2 * t] = B[2]; ob[(g + 8) * 8 + 2 * t + 1] = B[3];
    }
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double clk = clk_khz * 1e3;
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, SM clock %.0f MHz\n", prop.name, sms, clk / 1e6);
    uint32_t* d_in; float* d_out;
    std::vector<uint32_t> h_in(4096);
    std::mt19937 gen(1);
    for (auto& v : h_in) { float f = (float)(gen() % 2000) / 1000.0f - 1.0f; v = *(uint32_t*)&f & 0xffffe000u; }
    CK(cudaMalloc(&d_in, h_in.size() * 4)); CK(cudaMemcpy(d_in, h_in.data(), h_in.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, (size_t)sms * 8 * 256 * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // (1) issue rate
    for (int kind = 0; kind < 2; ++kind) for (int ctas = 1; ctas <= 4; ctas *= 2) {
        const int iters = 4000;
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            if (kind == 0) k_rate<0><<<sms * ctas, 256>>>(d_out, d_in, iters); else k_rate<1><<<sms * ctas, 256>>>(d_out, d_in, iters);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
        }
        const double mmas_per_smsp = (double)iters * CH * (ctas * 8) / 4.0;
        const double clk_per = best * 1e-3 * clk / mmas_per_smsp;
        const double macs = kind == 0 ? 1024.0 : 2048.0;
        printf("%s, %d warps/SMSP: %.3f ms, %.2f clk per MMA per SMSP, %.1f TFLOP/s dense\n", kind == 0 ? "mma.sync m16n8k8 tf32 " : "mma.sync m16n8k16 bf16", ctas * 2, best,
               clk_per, 2.0 * macs * mmas_per_smsp * 4 * sms / (best * 1e-3) / 1e12);
    }
    // (2) the loop step
    {
        const int n_steps = 32;  // 512 spheres
        float4* d_img; std::vector<float4> h_img(n_steps * 64);
        for (auto& v : h_img) { v.x = (float)(gen() % 2000) / 100.0f - 10.0f; v.y = v.x * 0.5f; v.z = -v.x; v.w = -300.0f; uint32_t* u = (uint32_t*)&v; for (int i = 0; i < 4; ++i) u[i] &= 0xffffe000u; }
        CK(cudaMalloc(&d_img, h_img.size() * 16)); CK(cudaMemcpy(d_img, h_img.data(), h_img.size() * 16, cudaMemcpyHostToDevice));
        const size_t smem = (size_t)n_steps * 64 * 16;
        const int trips = 64;
        for (int minb = 1; minb <= 3; ++minb) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0));
                if (minb == 1) { CK(cudaFuncSetAttribute(k_step<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_step<1><<<sms * 1, 256, smem>>>(d_out, d_in, d_img, n_steps, trips); }
                if (minb == 2) { CK(cudaFuncSetAttribute(k_step<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_step<2><<<sms * 2, 256, smem>>>(d_out, d_in, d_img, n_steps, trips); }
                if (minb == 3) { CK(cudaFuncSetAttribute(k_step<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_step<3><<<sms * 3, 256, smem>>>(d_out, d_in, d_img, n_steps, trips); }
                CK(cudaGetLastError());
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
            }
            const double tests = (double)sms * minb * 8 /*warps*/ * trips * n_steps * 512.0;
            const double clk_per32 = best * 1e-3 * clk * (sms * 4.0) / (tests / 32.0);
            printf("loop step (2 LDS.128 + 16 MMA + 8 FFMA2 + max + vote per 512 tests), %d CTAs/SM: %.3f ms, %.1f Gtests/s, %.2f clk per 32 tests per SMSP (FP32 loop: 11.0-11.5), x%.2f\n",
                   minb, best, tests / (best * 1e-3) / 1e9, clk_per32, 11.2 / clk_per32);
        }
    }
    // (3) accuracy of the split evaluation on RTIOW-like operands
    {
        const int n_tiles = 1 << 16;
        std::vector<float> S((size_t)n_tiles * 16 * 5), RA((size_t)n_tiles * 8 * 5), RB((size_t)n_tiles * 8 * 5);
        std::uniform_real_distribution<float> U(0.0f, 1.0f);
        const float kSlack = 1.0f / 16384.0f;  // 2^-14, the candidate value
        for (int tile = 0; tile < n_tiles; ++tile) {
            const int flavour = tile & 7;  // 0..5 small spheres around the origin, 6 the ground sphere, 7 far geometry
            for (int r = 0; r < 16; ++r) {
                float cx, cy, cz, rad;
                if (flavour == 6) { cx = 0; cy = -1000.0f; cz = 0; rad = 1000.0f; }
                else if (flavour == 7) { cx = (U(gen) - 0.5f) * 2.0e4f; cy = (U(gen) - 0.5f) * 2.0e4f; cz = (U(gen) - 0.5f) * 2.0e4f; rad = 1.0f + 100.0f * U(gen); }
                else { cx = (U(gen) - 0.5f) * 22.0f; cy = 0.2f + U(gen); cz = (U(gen) - 0.5f) * 22.0f; rad = 0.2f + 0.8f * (flavour == 0) * U(gen); }
                const double c2 = (double)cx * cx + (double)cy * cy + (double)cz * cz, r2 = (double)rad * rad;
                float* s = &S[((size_t)tile * 16 + r) * 5];
                s[0] = cx; s[1] = cy; s[2] = cz; s[3] = 1.0f; s[4] = (float)(r2 - c2);
            }
            for (int c = 0; c < 8; ++c) {
                float ox, oy, oz;
                if (flavour == 7) { ox = (U(gen) - 0.5f) * 2.0e4f; oy = (U(gen) - 0.5f) * 2.0e4f; oz = (U(gen) - 0.5f) * 2.0e4f; }
                else { ox = (U(gen) - 0.5f) * 26.0f; oy = U(gen) * 3.0f; oz = (U(gen) - 0.5f) * 26.0f; }
                float dx = U(gen) - 0.5f, dy = U(gen) - 0.5f, dz = U(gen) - 0.5f;
                const float n = sqrtf(dx * dx + dy * dy + dz * dz); dx /= n; dy /= n; dz /= n;
                float* ra = &RA[((size_t)tile * 8 + c) * 5];
                float* rb = &RB[((size_t)tile * 8 + c) * 5];
                ra[0] = dx; ra[1] = dy; ra[2] = dz; ra[3] = -((ox * dx + oy * dy) + oz * dz); ra[4] = 0.0f;
                rb[0] = ox + ox; rb[1] = oy + oy; rb[2] = oz + oz; rb[3] = -(((ox * ox + oy * oy) + oz * oz)); rb[4] = 1.0f;
            }
        }
        float *dS, *dRA, *dRB, *dA, *dB;
        CK(cudaMalloc(&dS, S.size() * 4)); CK(cudaMalloc(&dRA, RA.size() * 4)); CK(cudaMalloc(&dRB, RB.size() * 4));
        CK(cudaMalloc(&dA, (size_t)n_tiles * 128 * 4)); CK(cudaMalloc(&dB, (size_t)n_tiles * 128 * 4));
        CK(cudaMemcpy(dS, S.data(), S.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dRA, RA.data(), RA.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dRB, RB.data(), RB.size() * 4, cudaMemcpyHostToDevice));
        k_acc<<<sms * 4, 256>>>(dS, dRA, dRB, dA, dB, n_tiles);
        CK(cudaDeviceSynchronize());
        std::vector<float> hA((size_t)n_tiles * 128), hB((size_t)n_tiles * 128);
        CK(cudaMemcpy(hA.data(), dA, hA.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hB.data(), dB, hB.size() * 4, cudaMemcpyDeviceToHost));
        double worst[8] = {0}, worstA[8] = {0}, worstB[8] = {0};
        for (int tile = 0; tile < n_tiles; ++tile) for (int r = 0; r < 16; ++r) for (int c = 0; c < 8; ++c) {
            const float* s = &S[((size_t)tile * 16 + r) * 5];
            const float* ra = &RA[((size_t)tile * 8 + c) * 5];
            const float* rb = &RB[((size_t)tile * 8 + c) * 5];
            double A = 0, B = 0, sa = 0, sb = 0;
            for (int e = 0; e < 5; ++e) { A += (double)s[e] * ra[e]; B += (double)s[e] * rb[e]; sa += fabs((double)s[e] * ra[e]); sb += fabs((double)s[e] * rb[e]); }
            const double gA = hA[(size_t)tile * 128 + r * 8 + c], gB = hB[(size_t)tile * 128 + r * 8 + c];
            const double Lx = A * A + B, Lg = gA * gA + gB;  // the f32 FFMA that forms L' adds at most half an ulp of max(A^2, |B|): ignored here, budgeted in the slack
            const double c2 = (double)s[0] * s[0] + (double)s[1] * s[1] + (double)s[2] * s[2];
            const double r2 = s[4] + c2, o2 = -(double)rb[3];
            const double scale = c2 + fabs(r2) + o2;
            const int f = tile & 7;
            worst[f] = fmax(worst[f], fabs(Lg - Lx) / scale);
            worstA[f] = fmax(worstA[f], fabs(gA - A) / (sa + 1e-30));
            worstB[f] = fmax(worstB[f], fabs(gB - B) / (sb + 1e-30));
        }
        for (int f = 0; f < 8; ++f)
            printf("accuracy, operand flavour %d (%s): max |L'_mma - L'_f64| / (|c|^2+r^2+|o|^2) = %.3g = 2^%.1f   (A: %.3g of sum|terms|, B': %.3g)   candidate slack 2^-14 = %.3g\n", f,
                   f == 6 ? "ground sphere" : f == 7 ? "far geometry 1e4" : "small spheres", worst[f], log2(worst[f] + 1e-300), worstA[f], worstB[f], (double)kSlack);
    }
    return 0;
}
