// probe_fp32.cu — micro-benchmarks that decide the shape of the sphere sweep on sm_100a.
// Not part of the product path; kept so the design numbers in DESIGN.md can be re-measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o probe_fp32 probe_fp32.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

typedef unsigned long long u64;
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*(u64*)&a), "l"(*(u64*)&b), "l"(*(u64*)&c)); return *(float2*)&d; }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return *(float2*)&d; }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return *(float2*)&d; }

// ---------------- pure pipe peaks -----------------
template <int CH>
__global__ void k_ffma(float* out, int iters, float a, float b) {
    float acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0; 
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 3 distinct register operands per FFMA (no immediates / uniform regs)
template <int CH>
__global__ void k_ffma3(float* out, int iters, const float* in) {
    float acc[CH], m[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { acc[i] = threadIdx.x * 1e-3f + i; m[i] = in[threadIdx.x + i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fmaf(acc[i], m[i], m[(i + 1) % CH]);
    }
    float s = 0; 
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_ffma2(float* out, int iters, const float* in) {
    float2 acc[CH], m[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { acc[i] = make_float2(threadIdx.x * 1e-3f + i, i); m[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + i + 7]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fma2(acc[i], m[i], m[(i + 1) % CH]);
    }
    float s = 0; 
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------- sweep variants -----------------
struct RayIn { float ox, oy, oz, dx, dy, dz; };

// V0: scalar, AoS float4 (cx,cy,cz,r2) per sphere
__global__ void __launch_bounds__(256) k_sweep_scalar(const float4* __restrict__ sp, int n, const RayIn* __restrict__ rays, int rays_per_thread, float* out_t, int* out_i) {
    extern __shared__ float4 s[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = sp[i];
    __syncthreads();
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < rays_per_thread; ++r) {
        RayIn ry = rays[(size_t)r * gridDim.x * blockDim.x + tid];
        float ht = 3.402823466e38f; int hi = -1;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            float4 S = s[j];
            float cx = S.x - ry.ox, cy = S.y - ry.oy, cz = S.z - ry.oz;
            float nb = fmaf(cz, ry.dz, fmaf(cy, ry.dy, cx * ry.dx));
            float cc = fmaf(cz, cz, fmaf(cy, cy, fmaf(cx, cx, -S.w)));
            float disc = fmaf(nb, nb, -cc);
            if (disc > 0.f) {
                float sq = sqrtf(disc);
                float t = nb - sq;
                if (t < 0.001f) t = nb + sq;
                if (t > 0.001f && t < ht) { ht = t; hi = j; }
            }
        }
        out_t[(size_t)r * gridDim.x * blockDim.x + tid] = ht;
        out_i[(size_t)r * gridDim.x * blockDim.x + tid] = hi;
    }
}

// V1: packed: block of 4 spheres = 4 float4: X(cx0..3) Y Z R(r2)
template <int RPL, bool SLOW>  // rays per lane processed together (LDS amortisation); SLOW=false: count positives only
__global__ void __launch_bounds__(256) k_sweep_packed(const float4* __restrict__ sp, int n4, const RayIn* __restrict__ rays, int rays_per_thread, float* out_t, int* out_i) {
    extern __shared__ float4 s[];
    for (int i = threadIdx.x; i < n4 * 4; i += blockDim.x) s[i] = sp[i];
    __syncthreads();
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < rays_per_thread; r += RPL) {
        RayIn ry[RPL]; float ht[RPL]; int hi[RPL];
#pragma unroll
        for (int q = 0; q < RPL; ++q) { ry[q] = rays[(size_t)(r + q) * stride + tid]; ht[q] = 3.402823466e38f; hi[q] = -1; }
#pragma unroll 2
        for (int j = 0; j < n4; ++j) {
            float4 X = s[4 * j], Y = s[4 * j + 1], Z = s[4 * j + 2], R = s[4 * j + 3];
#pragma unroll
            for (int q = 0; q < RPL; ++q) {
                float2 o_x = make_float2(ry[q].ox, ry[q].ox), o_y = make_float2(ry[q].oy, ry[q].oy), o_z = make_float2(ry[q].oz, ry[q].oz);
                float2 d_x = make_float2(ry[q].dx, ry[q].dx), d_y = make_float2(ry[q].dy, ry[q].dy), d_z = make_float2(ry[q].dz, ry[q].dz);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float2 cx = sub2(h ? make_float2(X.z, X.w) : make_float2(X.x, X.y), o_x);
                    float2 cy = sub2(h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y), o_y);
                    float2 cz = sub2(h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y), o_z);
                    float2 r2 = h ? make_float2(R.z, R.w) : make_float2(R.x, R.y);
                    float2 nb = fma2(cz, d_z, fma2(cy, d_y, mul2(cx, d_x)));
                    float2 rhs = fma2(cz, cz, fma2(cy, cy, mul2(cx, cx)));
                    float2 lhs = fma2(nb, nb, r2);
                    if (!SLOW) { hi[q] += (lhs.x > rhs.x) + (lhs.y > rhs.y); }
                    else if (lhs.x > rhs.x || lhs.y > rhs.y) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            float l = e ? lhs.y : lhs.x, rr = e ? rhs.y : rhs.x, b = e ? nb.y : nb.x;
                            float disc = l - rr;
                            if (disc > 0.f) {
                                float sq = sqrtf(disc);
                                float t = b - sq;
                                if (t < 0.001f) t = b + sq;
                                if (t > 0.001f && t < ht[q]) { ht[q] = t; hi[q] = 4 * j + 2 * h + e; }
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < RPL; ++q) { out_t[(size_t)(r + q) * stride + tid] = ht[q]; out_i[(size_t)(r + q) * stride + tid] = hi[q]; }
    }
}

// V3: 1 ray per lane, sphere PAIRS as uniform operands from the constant bank, expanded algebra (7 FFMA2 per pair):
//   A = c.d - o.d ; B = 2 c.o + k (k = r^2 - |c|^2 + slack) ; candidate <=> A*A + B > |o|^2 (1 - 2^-19)
__constant__ float4 c_blk[4000];
template <int GROUP, bool SLOW>
__global__ void __launch_bounds__(256) k_sweep_const(const float4* __restrict__ sp_aos, int n4, const RayIn* __restrict__ rays, int rays_per_thread, float* out_t, int* out_i) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < rays_per_thread; ++r) {
        RayIn ry = rays[(size_t)r * stride + tid];
        const float nod = -(ry.ox * ry.dx + ry.oy * ry.dy + ry.oz * ry.dz);
        const float o2x = 2.f * ry.ox, o2y = 2.f * ry.oy, o2z = 2.f * ry.oz;
        const float oo = (ry.ox * ry.ox + ry.oy * ry.oy + ry.oz * ry.oz) * (1.0f - 1.9073486e-6f);
        float ht = 3.402823466e38f; int hi = SLOW ? -1 : 0;
        for (int j = 0; j < n4; j += GROUP) {
            float2 L[2 * GROUP];
#pragma unroll
            for (int g = 0; g < GROUP; ++g) {
                const float4 X = c_blk[4 * (j + g)], Y = c_blk[4 * (j + g) + 1], Z = c_blk[4 * (j + g) + 2], K = c_blk[4 * (j + g) + 3];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 cx = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y), cy = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                    const float2 cz = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y), k = h ? make_float2(K.z, K.w) : make_float2(K.x, K.y);
                    const float2 A = fma2(cz, make_float2(ry.dz, ry.dz), fma2(cy, make_float2(ry.dy, ry.dy), fma2(cx, make_float2(ry.dx, ry.dx), make_float2(nod, nod))));
                    const float2 B = fma2(cz, make_float2(o2z, o2z), fma2(cy, make_float2(o2y, o2y), fma2(cx, make_float2(o2x, o2x), k)));
                    L[2 * g + h] = fma2(A, A, B);
                }
            }
            bool any = false;
#pragma unroll
            for (int q = 0; q < 2 * GROUP; ++q) any = any | (L[q].x > oo) | (L[q].y > oo);
            if (!SLOW) { hi += any; }
            else if (any) {
#pragma unroll
                for (int q = 0; q < 2 * GROUP; ++q) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if ((e ? L[q].y : L[q].x) > oo) {
                            const int idx = 4 * j + 2 * q + e;
                            const float4 S = sp_aos[idx];
                            float cx = S.x - ry.ox, cy = S.y - ry.oy, cz = S.z - ry.oz;
                            float nb = fmaf(cz, ry.dz, fmaf(cy, ry.dy, cx * ry.dx));
                            float cc = fmaf(cz, cz, fmaf(cy, cy, fmaf(cx, cx, -S.w)));
                            float disc = fmaf(nb, nb, -cc);
                            if (disc > 0.f) { float sq = sqrtf(disc); float t = nb - sq; if (t < 0.001f) t = nb + sq; if (t > 0.001f && t < ht) { ht = t; hi = idx; } }
                        }
                    }
                }
            }
        }
        out_t[(size_t)r * stride + tid] = ht; out_i[(size_t)r * stride + tid] = hi;
    }
}

static float frand(uint64_t& s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (float)((s >> 40) & 0xFFFFFF) / 16777216.0f; }

int main(int argc, char** argv) {
    bool quick = argc > 1;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount; int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device %s sms %d clock_attr %.0f MHz\n", prop.name, sms, clk_khz / 1e3);
    double peak = sms * 128.0 * 2.0 * 1.965e9;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float* d_out; CK(cudaMalloc(&d_out, sizeof(float) * sms * 32 * 1024)); float* d_in; CK(cudaMalloc(&d_in, 4096 * 4)); CK(cudaMemset(d_in, 0, 4096 * 4));
    auto timeit = [&](auto launch, int reps) { launch(); CK(cudaDeviceSynchronize()); float best = 1e30f; for (int i = 0; i < reps; ++i) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; } return best; };
    const int iters = 20000;
    for (int wps : {4, 8, 16, 32}) {  // warps per SM
        if (quick) break;
        int threads = 256, blocks = sms * wps * 32 / threads; if (blocks < 1) { blocks = sms; threads = wps * 32; }
        if (wps * 32 < 256) { threads = wps * 32; blocks = sms; }
        float ms;
        ms = timeit([&] { k_ffma<8><<<blocks, threads>>>(d_out, iters, 1.0001f, 0.5f); }, 3);
        printf("ffma_imm   warps/SM %2d : %.2f TFLOP/s (%.1f%% of 74.45)\n", wps, 2.0 * 8 * iters * (double)blocks * threads / ms / 1e9, 100 * 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / peak);
        ms = timeit([&] { k_ffma3<8><<<blocks, threads>>>(d_out, iters, d_in); }, 3);
        printf("ffma_3reg  warps/SM %2d : %.2f TFLOP/s (%.1f%%)\n", wps, 2.0 * 8 * iters * (double)blocks * threads / ms / 1e9, 100 * 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / peak);
        ms = timeit([&] { k_ffma2<8><<<blocks, threads>>>(d_out, iters, d_in); }, 3);
        printf("ffma2_3reg warps/SM %2d : %.2f TFLOP/s (%.1f%%)\n", wps, 4.0 * 8 * iters * (double)blocks * threads / ms / 1e9, 100 * 4.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / peak);
    }
    // scene: ground + 22x22 grid + 3 big = 488
    std::vector<float> cx, cy, cz, rr; uint64_t seed = 12345;
    cx.push_back(0); cy.push_back(-1000); cz.push_back(0); rr.push_back(1000);
    for (int a = -11; a < 11; ++a) for (int b = -11; b < 11; ++b) { cx.push_back(a + 0.9f * frand(seed)); cy.push_back(0.2f); cz.push_back(b + 0.9f * frand(seed)); rr.push_back(0.2f); }
    cx.push_back(0); cy.push_back(1); cz.push_back(0); rr.push_back(1);
    cx.push_back(-4); cy.push_back(1); cz.push_back(0); rr.push_back(1);
    cx.push_back(4); cy.push_back(1); cz.push_back(0); rr.push_back(1);
    int n = (int)cx.size(); int n4 = (n + 3) / 4;
    std::vector<float4> aos(n4 * 4), blk(n4 * 4);
    for (int i = 0; i < n4 * 4; ++i) { bool v = i < n; aos[i] = make_float4(v ? cx[i] : 3.0e38f, v ? cy[i] : 3.0e38f, v ? cz[i] : 3.0e38f, v ? rr[i] * rr[i] : 0.f); }
    for (int j = 0; j < n4; ++j) { float* X = (float*)&blk[4 * j]; for (int e = 0; e < 4; ++e) { int i = 4 * j + e; bool v = i < n; X[e] = v ? cx[i] : 1.0e18f; X[4 + e] = v ? cy[i] : 1.0e18f; X[8 + e] = v ? cz[i] : 1.0e18f; X[12 + e] = v ? rr[i] * rr[i] : 0.f; } }
    {
        std::vector<float4> cb(((n4 + 3) / 4 * 4 + 4) * 4);
        for (size_t j = 0; j < cb.size() / 4; ++j) { float* X = (float*)&cb[4 * j]; for (int e = 0; e < 4; ++e) { size_t i = 4 * j + e; bool v = i < (size_t)n;
            double c2 = v ? (double)cx[i] * cx[i] + (double)cy[i] * cy[i] + (double)cz[i] * cz[i] : 0.0; double r2 = v ? (double)rr[i] * rr[i] : 0.0;
            X[e] = v ? cx[i] : 0.f; X[4 + e] = v ? cy[i] : 0.f; X[8 + e] = v ? cz[i] : 0.f; X[12 + e] = v ? (float)(r2 - c2 + 32 * 5.96e-8 * (c2 + r2)) : -3.0e38f; } }
        CK(cudaMemcpyToSymbol(c_blk, cb.data(), sizeof(float4) * cb.size()));
    }
    float4 *d_aos, *d_blk; CK(cudaMalloc(&d_aos, sizeof(float4) * n4 * 4)); CK(cudaMalloc(&d_blk, sizeof(float4) * n4 * 4));
    CK(cudaMemcpy(d_aos, aos.data(), sizeof(float4) * n4 * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_blk, blk.data(), sizeof(float4) * n4 * 4, cudaMemcpyHostToDevice));
    // rays: half primary-like from (13,2,3) towards origin region, half secondary from the ground plane into the upper hemisphere
    const int RPT = 64;
    for (int mix = 0; mix < 2; ++mix) {
    for (int wps : {8, 16, 24, 32}) {
        if (quick && !(mix == 1 && wps == 24)) continue;
        int threads = 256, blocks = sms * wps * 32 / threads; size_t nthreads = (size_t)blocks * threads; size_t nr = nthreads * RPT;
        std::vector<RayIn> rays(nr);
        for (size_t i = 0; i < nr; ++i) {
            RayIn r; bool prim = mix == 0 ? true : (frand(seed) < 0.4f);
            if (prim) { r.ox = 13 + 0.05f * (frand(seed) - 0.5f); r.oy = 2 + 0.05f * (frand(seed) - 0.5f); r.oz = 3; float tx = 12 * (frand(seed) - 0.5f), ty = 3 * (frand(seed) - 0.3f), tz = 12 * (frand(seed) - 0.5f); float dx = tx - r.ox, dy = ty - r.oy, dz = tz - r.oz; float il = 1.f / sqrtf(dx * dx + dy * dy + dz * dz); r.dx = dx * il; r.dy = dy * il; r.dz = dz * il; }
            else { r.ox = 20 * (frand(seed) - 0.5f); r.oz = 20 * (frand(seed) - 0.5f); r.oy = 0.0f; float dx, dy, dz, l2; do { dx = 2 * frand(seed) - 1; dy = frand(seed); dz = 2 * frand(seed) - 1; l2 = dx * dx + dy * dy + dz * dz; } while (l2 > 1 || l2 < 1e-4f); float il = 1.f / sqrtf(l2); r.dx = dx * il; r.dy = dy * il; r.dz = dz * il; }
            rays[i] = r;
        }
        RayIn* d_rays; CK(cudaMalloc(&d_rays, sizeof(RayIn) * nr)); CK(cudaMemcpy(d_rays, rays.data(), sizeof(RayIn) * nr, cudaMemcpyHostToDevice));
        float* d_t; int* d_i; CK(cudaMalloc(&d_t, 4 * nr)); CK(cudaMalloc(&d_i, 4 * nr));
        std::vector<int> i0(nr), i1(nr);
        size_t smem = sizeof(float4) * n4 * 4;
        double tests = (double)nr * n;
        float ms = timeit([&] { k_sweep_scalar<<<blocks, threads, smem>>>(d_aos, n4 * 4, d_rays, RPT, d_t, d_i); }, 3);
        CK(cudaMemcpy(i0.data(), d_i, 4 * nr, cudaMemcpyDeviceToHost));
        printf("mix %d sweep_scalar   warps/SM %2d : %.3f ms  %.1f Gtests/s  %.2f TFLOP/s (%.1f%% of peak)\n", mix, wps, ms, tests / ms / 1e6, 16 * tests / ms / 1e9, 100 * 16 * tests / (ms * 1e-3) / peak);
        ms = timeit([&] { k_sweep_packed<1, true><<<blocks, threads, smem>>>(d_blk, n4, d_rays, RPT, d_t, d_i); }, 3);
        CK(cudaMemcpy(i1.data(), d_i, 4 * nr, cudaMemcpyDeviceToHost)); size_t diff = 0, hits = 0; for (size_t i = 0; i < nr; ++i) { diff += i0[i] != i1[i]; hits += i0[i] >= 0; }
        printf("mix %d sweep_packed<1> warps/SM %2d : %.3f ms  %.1f Gtests/s  %.2f TFLOP/s (%.1f%% of peak)  mismatches %zu / %zu (hit frac %.3f)\n", mix, wps, ms, tests / ms / 1e6, 16 * tests / ms / 1e9, 100 * 16 * tests / (ms * 1e-3) / peak, diff, nr, (double)hits / nr);
        ms = timeit([&] { k_sweep_packed<2, true><<<blocks, threads, smem>>>(d_blk, n4, d_rays, RPT, d_t, d_i); }, 3);
        CK(cudaMemcpy(i1.data(), d_i, 4 * nr, cudaMemcpyDeviceToHost)); diff = 0; for (size_t i = 0; i < nr; ++i) diff += i0[i] != i1[i];
        printf("mix %d sweep_packed<2> warps/SM %2d : %.3f ms  %.1f Gtests/s  %.2f TFLOP/s (%.1f%% of peak)  mismatches %zu\n", mix, wps, ms, tests / ms / 1e6, 16 * tests / ms / 1e9, 100 * 16 * tests / (ms * 1e-3) / peak, diff);
        ms = timeit([&] { k_sweep_packed<1, false><<<blocks, threads, smem>>>(d_blk, n4, d_rays, RPT, d_t, d_i); }, 3);
        printf("mix %d sweep_packed<1,noslow> warps/SM %2d : %.3f ms  %.1f Gtests/s  %.2f TFLOP/s (%.1f%% of peak)\n", mix, wps, ms, tests / ms / 1e6, 16 * tests / ms / 1e9, 100 * 16 * tests / (ms * 1e-3) / peak);
        ms = timeit([&] { k_sweep_packed<2, false><<<blocks, threads, smem>>>(d_blk, n4, d_rays, RPT, d_t, d_i); }, 3);
        printf("mix %d sweep_packed<2,noslow> warps/SM %2d : %.3f ms  %.1f Gtests/s  %.2f TFLOP/s (%.1f%% of peak)\n", mix, wps, ms, tests / ms / 1e6, 16 * tests / ms / 1e9, 100 * 16 * tests / (ms * 1e-3) / peak);
        ms = timeit([&] { k_sweep_const<2, true><<<blocks, threads>>>(d_aos, (n4 + 1) / 2 * 2, d_rays, RPT, d_t, d_i); }, 3);
        CK(cudaMemcpy(i1.data(), d_i, 4 * nr, cudaMemcpyDeviceToHost)); diff = 0; for (size_t i = 0; i < nr; ++i) diff += i0[i] != i1[i];
        printf("mix %d sweep_const<2>      warps/SM %2d : %.3f ms  %.1f Gtests/s  %.2f TFLOP/s (%.1f%% of peak)  mismatches %zu\n", mix, wps, ms, tests / ms / 1e6, 16 * tests / ms / 1e9, 100 * 16 * tests / (ms * 1e-3) / peak, diff);
        ms = timeit([&] { k_sweep_const<4, true><<<blocks, threads>>>(d_aos, (n4 + 3) / 4 * 4, d_rays, RPT, d_t, d_i); }, 3);
        CK(cudaMemcpy(i1.data(), d_i, 4 * nr, cudaMemcpyDeviceToHost)); diff = 0; for (size_t i = 0; i < nr; ++i) diff += i0[i] != i1[i];
        printf("mix %d sweep_const<4>      warps/SM %2d : %.3f ms  %.1f Gtests/s  %.2f TFLOP/s (%.1f%% of peak)  mismatches %zu\n", mix, wps, ms, tests / ms / 1e6, 16 * tests / ms / 1e9, 100 * 16 * tests / (ms * 1e-3) / peak, diff);
        ms = timeit([&] { k_sweep_const<2, false><<<blocks, threads>>>(d_aos, (n4 + 1) / 2 * 2, d_rays, RPT, d_t, d_i); }, 3);
        printf("mix %d sweep_const<2,noslow> warps/SM %2d : %.3f ms  %.1f Gtests/s  %.2f TFLOP/s (%.1f%% of peak)\n", mix, wps, ms, tests / ms / 1e6, 16 * tests / ms / 1e9, 100 * 16 * tests / (ms * 1e-3) / peak);
        CK(cudaFree(d_rays)); CK(cudaFree(d_t)); CK(cudaFree(d_i));
    }
    }
    return 0;
}
