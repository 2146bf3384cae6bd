// probe_sweep2.cu — throughput of candidate shapes of the pre-filter loop (pt_sweep.cuh) in isolation, round 2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o probe_sweep2 probe_sweep2.cu
//
// Every kernel sweeps the same staged pre-filter image (X,Y,Z,K blocks of 4 spheres in shared memory, LDS.128 broadcast)
// and flags groups of 16 spheres exactly like the product loop (max-tree + one branch, push of a 32-bit entry).
//   one   : the round-1 form — one ray per lane, FFMA2 spherePair * rayScalar + accPair (7 packed FMAs per 2 spheres)
//   two   : two rays per lane — FFMA2 sphereScalar * rayPair + accPair (7 packed FMAs per sphere = 2 tests); the
//           64-bit ray operand is loop-invariant and sits in the same operand slot of consecutive instructions, so the
//           operand-reuse cache can serve it; the per-instruction register reads drop from 5 words to 3
// Reported: clk per (ray, sphere) test per SM sub-partition and the fraction of the FP32 peak at 16 flop per test.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cstring>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*(u64*)&a), "l"(*(u64*)&b), "l"(*(u64*)&c)); return *(float2*)&d; }
__device__ __forceinline__ float4 lds128(uint32_t addr) { float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)); return v; }
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }

constexpr int kThreads = 256;
constexpr int kQueueCap = 12;

// ---------------- one ray per lane (round-1 product loop, group of 16 spheres) ----------------
template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_one(const float4* __restrict__ pfg, int n_blocks, const float* __restrict__ rays, int trips, unsigned* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* pf = reinterpret_cast<float4*>(smem);
    uint32_t* q = reinterpret_cast<uint32_t*>(smem + (size_t)n_blocks * 64) + threadIdx.x;
    for (int i = threadIdx.x; i < n_blocks * 4; i += kThreads) pf[i] = pfg[i];
    __syncthreads();
    const int gid = blockIdx.x * kThreads + threadIdx.x;
    float ox = rays[gid * 6 + 0], oy = rays[gid * 6 + 1], oz = rays[gid * 6 + 2], dx = rays[gid * 6 + 3], dy = rays[gid * 6 + 4], dz = rays[gid * 6 + 5];
    unsigned total = 0;
    for (int t = 0; t < trips; ++t) {
        float nod = -((ox * dx + oy * dy) + oz * dz);
        float oo = ((ox * ox + oy * oy) + oz * oz) * (1.0f - 1.9073486328125e-06f);
        float o2x = ox + ox, o2y = oy + oy, o2z = oz + oz;
        asm volatile("" : "+f"(nod), "+f"(o2x), "+f"(o2y), "+f"(o2z), "+f"(oo), "+r"(n_blocks));
        uint32_t addr = (uint32_t)__cvta_generic_to_shared(pf);
        asm volatile("" : "+r"(addr));
        uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(q);
        asm volatile("" : "+r"(qaddr));
        const uint32_t base = addr, end = addr + 64u * (uint32_t)n_blocks;
        int cnt = 0;
#pragma unroll 1
        for (; addr < end; addr += 64u * 4) {
            float2 L[8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 X = lds128(addr + 64u * g), Y = lds128(addr + 64u * g + 16u), Z = lds128(addr + 64u * g + 32u), K = lds128(addr + 64u * g + 48u);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 cx = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                    const float2 cy = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                    const float2 cz = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
                    const float2 k = h ? make_float2(K.z, K.w) : make_float2(K.x, K.y);
                    const float2 A = fma2(cz, bc(dz), fma2(cy, bc(dy), fma2(cx, bc(dx), bc(nod))));
                    const float2 B = fma2(cz, bc(o2z), fma2(cy, bc(o2y), fma2(cx, bc(o2x), k)));
                    L[2 * g + h] = fma2(A, A, B);
                }
            }
            bool any = false;
#pragma unroll
            for (int p = 0; p < 8; ++p) any = any | (L[p].x > oo) | (L[p].y > oo);
            if (any) {
                uint32_t mask = 0u;
#pragma unroll
                for (int p = 7; p >= 0; --p) {
                    const float2 d = fma2(L[p], bc(-1.0f), bc(oo));
                    mask = __funnelshift_l(__float_as_uint(d.y), mask, 1);
                    mask = __funnelshift_l(__float_as_uint(d.x), mask, 1);
                }
                const uint32_t entry = ((((addr - base) >> 6) / 4) << 16) | mask;
                if (cnt < kQueueCap) {
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + (uint32_t)cnt * (uint32_t)(kThreads * 4)), "r"(entry) : "memory");
                    cnt += 1;
                }
            }
        }
        total += cnt;
        // next trip: a slightly different ray, so nothing is hoisted out of the trip loop
        ox += 1e-3f * dx; oy += 1e-3f * dy; oz += 1e-3f * dz;
    }
    out[gid] = total;
}

// ---------------- one ray per lane, sphere pairs through the constant bank / uniform registers ----------------
constexpr int kMaxConstBlocks = 1000;
__constant__ float4 c_pf[4 * kMaxConstBlocks];
template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_const(const float4* __restrict__ pfg, int n_blocks, const float* __restrict__ rays, int trips, unsigned* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* pf = reinterpret_cast<float4*>(smem);
    uint32_t* q = reinterpret_cast<uint32_t*>(smem + (size_t)n_blocks * 64) + threadIdx.x;
    for (int i = threadIdx.x; i < n_blocks * 4; i += kThreads) pf[i] = pfg[i];
    __syncthreads();
    const int gid = blockIdx.x * kThreads + threadIdx.x;
    float ox = rays[gid * 6 + 0], oy = rays[gid * 6 + 1], oz = rays[gid * 6 + 2], dx = rays[gid * 6 + 3], dy = rays[gid * 6 + 4], dz = rays[gid * 6 + 5];
    unsigned total = 0;
    for (int t = 0; t < trips; ++t) {
        float nod = -((ox * dx + oy * dy) + oz * dz);
        float oo = ((ox * ox + oy * oy) + oz * oz) * (1.0f - 1.9073486328125e-06f);
        float o2x = ox + ox, o2y = oy + oy, o2z = oz + oz;
        asm volatile("" : "+f"(nod), "+f"(o2x), "+f"(o2y), "+f"(o2z), "+f"(oo));
        uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(q);
        asm volatile("" : "+r"(qaddr));
        int cnt = 0;
#pragma unroll 1
        for (int j = 0; j < n_blocks; j += 4) {
            float2 L[8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 X = c_pf[4 * (j + g) + 0], Y = c_pf[4 * (j + g) + 1], Z = c_pf[4 * (j + g) + 2];
                const float4 K = pf[4 * (j + g) + 3];  // plain shared load with a uniform index (an asm "r" operand would force j into a vector register)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 cx = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                    const float2 cy = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                    const float2 cz = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
                    const float2 k = h ? make_float2(K.z, K.w) : make_float2(K.x, K.y);
                    const float2 A = fma2(cz, bc(dz), fma2(cy, bc(dy), fma2(cx, bc(dx), bc(nod))));
                    const float2 B = fma2(cz, bc(o2z), fma2(cy, bc(o2y), fma2(cx, bc(o2x), k)));
                    L[2 * g + h] = fma2(A, A, B);
                }
            }
            bool any = false;
#pragma unroll
            for (int p = 0; p < 8; ++p) any = any | (L[p].x > oo) | (L[p].y > oo);
            if (any) {
                uint32_t mask = 0u;
#pragma unroll
                for (int p = 7; p >= 0; --p) {
                    const float2 d = fma2(L[p], bc(-1.0f), bc(oo));
                    mask = __funnelshift_l(__float_as_uint(d.y), mask, 1);
                    mask = __funnelshift_l(__float_as_uint(d.x), mask, 1);
                }
                const uint32_t entry = ((uint32_t)(j / 4) << 16) | mask;
                if (cnt < kQueueCap) {
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + (uint32_t)cnt * (uint32_t)(kThreads * 4)), "r"(entry) : "memory");
                    cnt += 1;
                }
            }
        }
        total += cnt;
        ox += 1e-3f * dx; oy += 1e-3f * dy; oz += 1e-3f * dz;
    }
    out[gid] = total;
}

template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_const_diag(const float4* __restrict__ pfg, int n_blocks, const float* __restrict__ rays, int trips, unsigned* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* pf = reinterpret_cast<float4*>(smem);
    uint32_t* q = reinterpret_cast<uint32_t*>(smem + (size_t)n_blocks * 64) + threadIdx.x;
    for (int i = threadIdx.x; i < n_blocks * 4; i += kThreads) pf[i] = pfg[i];
    __syncthreads();
    const int gid = blockIdx.x * kThreads + threadIdx.x;
    float ox = rays[gid * 6 + 0], oy = rays[gid * 6 + 1], oz = rays[gid * 6 + 2], dx = rays[gid * 6 + 3], dy = rays[gid * 6 + 4], dz = rays[gid * 6 + 5];
    unsigned total = 0;
    for (int t = 0; t < trips; ++t) {
        float nod = -((ox * dx + oy * dy) + oz * dz);
        float oo = ((ox * ox + oy * oy) + oz * oz) * (1.0f - 1.9073486328125e-06f);
        float o2x = ox + ox, o2y = oy + oy, o2z = oz + oz;
        asm volatile("" : "+f"(nod), "+f"(o2x), "+f"(o2y), "+f"(o2z), "+f"(oo));
        uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(q);
        asm volatile("" : "+r"(qaddr));
        int cnt = 0;
#pragma unroll 1
        for (int j = 0; j < n_blocks; j += 4) {
            float2 L[8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 X = c_pf[4 * (j + g) + 0]; const float4 Y = make_float4(X.y, X.x, X.w, X.z), Z = make_float4(X.z, X.w, X.x, X.y);
                const float4 K = pf[4 * (j + g) + 3];  // plain shared load with a uniform index (an asm "r" operand would force j into a vector register)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 cx = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                    const float2 cy = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                    const float2 cz = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
                    const float2 k = h ? make_float2(K.z, K.w) : make_float2(K.x, K.y);
                    const float2 A = fma2(cz, bc(dz), fma2(cy, bc(dy), fma2(cx, bc(dx), bc(nod))));
                    const float2 B = fma2(cz, bc(o2z), fma2(cy, bc(o2y), fma2(cx, bc(o2x), k)));
                    L[2 * g + h] = fma2(A, A, B);
                }
            }
            bool any = false;
#pragma unroll
            for (int p = 0; p < 8; ++p) any = any | (L[p].x > oo) | (L[p].y > oo);
            if (any) {
                uint32_t mask = 0u;
#pragma unroll
                for (int p = 7; p >= 0; --p) {
                    const float2 d = fma2(L[p], bc(-1.0f), bc(oo));
                    mask = __funnelshift_l(__float_as_uint(d.y), mask, 1);
                    mask = __funnelshift_l(__float_as_uint(d.x), mask, 1);
                }
                const uint32_t entry = ((uint32_t)(j / 4) << 16) | mask;
                if (cnt < kQueueCap) {
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + (uint32_t)cnt * (uint32_t)(kThreads * 4)), "r"(entry) : "memory");
                    cnt += 1;
                }
            }
        }
        total += cnt;
        ox += 1e-3f * dx; oy += 1e-3f * dy; oz += 1e-3f * dz;
    }
    out[gid] = total;
}

template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_const2(const float4* __restrict__ pfg, int n_blocks, const float* __restrict__ rays, int trips, unsigned* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* pf = reinterpret_cast<float4*>(smem);
    uint32_t* q = reinterpret_cast<uint32_t*>(smem + (size_t)n_blocks * 64) + threadIdx.x;
    for (int i = threadIdx.x; i < n_blocks * 4; i += kThreads) pf[i] = pfg[i];
    __syncthreads();
    const int gid = blockIdx.x * kThreads + threadIdx.x;
    const int n_lanes = gridDim.x * kThreads;
    float ox[2], oy[2], oz[2], dx[2], dy[2], dz[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int id = gid + r * n_lanes;
        ox[r] = rays[id * 6 + 0]; oy[r] = rays[id * 6 + 1]; oz[r] = rays[id * 6 + 2]; dx[r] = rays[id * 6 + 3]; dy[r] = rays[id * 6 + 4]; dz[r] = rays[id * 6 + 5];
    }
    unsigned total = 0;
    for (int t = 0; t < trips; ++t) {
        float nod[2], oo[2], o2x[2], o2y[2], o2z[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            nod[r] = -((ox[r] * dx[r] + oy[r] * dy[r]) + oz[r] * dz[r]);
            oo[r] = ((ox[r] * ox[r] + oy[r] * oy[r]) + oz[r] * oz[r]) * (1.0f - 1.9073486328125e-06f);
            o2x[r] = ox[r] + ox[r]; o2y[r] = oy[r] + oy[r]; o2z[r] = oz[r] + oz[r];
            asm volatile("" : "+f"(nod[r]), "+f"(o2x[r]), "+f"(o2y[r]), "+f"(o2z[r]), "+f"(oo[r]));
        }
        uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(q);
        asm volatile("" : "+r"(qaddr));
        int cnt = 0;
#pragma unroll 1
        for (int j = 0; j < n_blocks; j += 4) {
            float2 L[2][8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 X = c_pf[4 * (j + g) + 0], Y = c_pf[4 * (j + g) + 1], Z = c_pf[4 * (j + g) + 2];
                const float4 K = pf[4 * (j + g) + 3];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 cx = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                    const float2 cy = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                    const float2 cz = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
                    const float2 k = h ? make_float2(K.z, K.w) : make_float2(K.x, K.y);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float2 A = fma2(cz, bc(dz[r]), fma2(cy, bc(dy[r]), fma2(cx, bc(dx[r]), bc(nod[r]))));
                        const float2 B = fma2(cz, bc(o2z[r]), fma2(cy, bc(o2y[r]), fma2(cx, bc(o2x[r]), k)));
                        L[r][2 * g + h] = fma2(A, A, B);
                    }
                }
            }
            bool any = false;
#pragma unroll
            for (int p = 0; p < 8; ++p) any = any | (L[0][p].x > oo[0]) | (L[0][p].y > oo[0]) | (L[1][p].x > oo[1]) | (L[1][p].y > oo[1]);
            if (any) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    uint32_t mask = 0u;
#pragma unroll
                    for (int p = 7; p >= 0; --p) {
                        const float2 d = fma2(L[r][p], bc(-1.0f), bc(oo[r]));
                        mask = __funnelshift_l(__float_as_uint(d.y), mask, 1);
                        mask = __funnelshift_l(__float_as_uint(d.x), mask, 1);
                    }
                    const uint32_t entry = ((uint32_t)(j / 4) << 16) | mask;
                    if (mask != 0u && cnt < kQueueCap) {
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + (uint32_t)cnt * (uint32_t)(kThreads * 4)), "r"(entry) : "memory");
                        cnt += 1;
                    }
                }
            }
        }
        total += cnt;
#pragma unroll
        for (int r = 0; r < 2; ++r) { ox[r] += 1e-3f * dx[r]; oy[r] += 1e-3f * dy[r]; oz[r] += 1e-3f * dz[r]; }
    }
    out[gid] = total;
}

template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_lds2(const float4* __restrict__ pfg, int n_blocks, const float* __restrict__ rays, int trips, unsigned* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* pf = reinterpret_cast<float4*>(smem);
    uint32_t* q = reinterpret_cast<uint32_t*>(smem + (size_t)n_blocks * 64) + threadIdx.x;
    for (int i = threadIdx.x; i < n_blocks * 4; i += kThreads) pf[i] = pfg[i];
    __syncthreads();
    const int gid = blockIdx.x * kThreads + threadIdx.x;
    const int n_lanes = gridDim.x * kThreads;
    float ox[2], oy[2], oz[2], dx[2], dy[2], dz[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int id = gid + r * n_lanes;
        ox[r] = rays[id * 6 + 0]; oy[r] = rays[id * 6 + 1]; oz[r] = rays[id * 6 + 2]; dx[r] = rays[id * 6 + 3]; dy[r] = rays[id * 6 + 4]; dz[r] = rays[id * 6 + 5];
    }
    unsigned total = 0;
    for (int t = 0; t < trips; ++t) {
        float nod[2], oo[2], o2x[2], o2y[2], o2z[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            nod[r] = -((ox[r] * dx[r] + oy[r] * dy[r]) + oz[r] * dz[r]);
            oo[r] = ((ox[r] * ox[r] + oy[r] * oy[r]) + oz[r] * oz[r]) * (1.0f - 1.9073486328125e-06f);
            o2x[r] = ox[r] + ox[r]; o2y[r] = oy[r] + oy[r]; o2z[r] = oz[r] + oz[r];
            asm volatile("" : "+f"(nod[r]), "+f"(o2x[r]), "+f"(o2y[r]), "+f"(o2z[r]), "+f"(oo[r]));
        }
        uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(q);
        asm volatile("" : "+r"(qaddr));
        int cnt = 0;
        uint32_t addr = (uint32_t)__cvta_generic_to_shared(pf);
        asm volatile("" : "+r"(addr));
#pragma unroll 1
        for (int j = 0; j < n_blocks; j += 4) {
            float2 L[2][8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 X = lds128(addr + 64u * (uint32_t)(j + g)), Y = lds128(addr + 64u * (uint32_t)(j + g) + 16u), Z = lds128(addr + 64u * (uint32_t)(j + g) + 32u), K = lds128(addr + 64u * (uint32_t)(j + g) + 48u);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 cx = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                    const float2 cy = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                    const float2 cz = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
                    const float2 k = h ? make_float2(K.z, K.w) : make_float2(K.x, K.y);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float2 A = fma2(cz, bc(dz[r]), fma2(cy, bc(dy[r]), fma2(cx, bc(dx[r]), bc(nod[r]))));
                        const float2 B = fma2(cz, bc(o2z[r]), fma2(cy, bc(o2y[r]), fma2(cx, bc(o2x[r]), k)));
                        L[r][2 * g + h] = fma2(A, A, B);
                    }
                }
            }
            bool any = false;
#pragma unroll
            for (int p = 0; p < 8; ++p) any = any | (L[0][p].x > oo[0]) | (L[0][p].y > oo[0]) | (L[1][p].x > oo[1]) | (L[1][p].y > oo[1]);
            if (any) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    uint32_t mask = 0u;
#pragma unroll
                    for (int p = 7; p >= 0; --p) {
                        const float2 d = fma2(L[r][p], bc(-1.0f), bc(oo[r]));
                        mask = __funnelshift_l(__float_as_uint(d.y), mask, 1);
                        mask = __funnelshift_l(__float_as_uint(d.x), mask, 1);
                    }
                    const uint32_t entry = ((uint32_t)(j / 4) << 16) | mask;
                    if (mask != 0u && cnt < kQueueCap) {
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + (uint32_t)cnt * (uint32_t)(kThreads * 4)), "r"(entry) : "memory");
                        cnt += 1;
                    }
                }
            }
        }
        total += cnt;
#pragma unroll
        for (int r = 0; r < 2; ++r) { ox[r] += 1e-3f * dx[r]; oy[r] += 1e-3f * dy[r]; oz[r] += 1e-3f * dz[r]; }
    }
    out[gid] = total;
}

// ---------------- same, image passed by value as a kernel parameter (constant bank 0) ----------------
struct PfParam { float4 v[4 * 128]; };  // 512 spheres = 8 KB of kernel parameters
template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_param(const __grid_constant__ PfParam cp, const float4* __restrict__ pfg, int n_blocks, const float* __restrict__ rays, int trips, unsigned* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* pf = reinterpret_cast<float4*>(smem);
    uint32_t* q = reinterpret_cast<uint32_t*>(smem + (size_t)n_blocks * 64) + threadIdx.x;
    for (int i = threadIdx.x; i < n_blocks * 4; i += kThreads) pf[i] = pfg[i];
    __syncthreads();
    const int gid = blockIdx.x * kThreads + threadIdx.x;
    float ox = rays[gid * 6 + 0], oy = rays[gid * 6 + 1], oz = rays[gid * 6 + 2], dx = rays[gid * 6 + 3], dy = rays[gid * 6 + 4], dz = rays[gid * 6 + 5];
    unsigned total = 0;
    for (int t = 0; t < trips; ++t) {
        float nod = -((ox * dx + oy * dy) + oz * dz);
        float oo = ((ox * ox + oy * oy) + oz * oz) * (1.0f - 1.9073486328125e-06f);
        float o2x = ox + ox, o2y = oy + oy, o2z = oz + oz;
        asm volatile("" : "+f"(nod), "+f"(o2x), "+f"(o2y), "+f"(o2z), "+f"(oo));
        uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(q);
        asm volatile("" : "+r"(qaddr));
        int cnt = 0;
#pragma unroll 1
        for (int j = 0; j < n_blocks; j += 4) {
            float2 L[8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 X = cp.v[4 * (j + g) + 0], Y = cp.v[4 * (j + g) + 1], Z = cp.v[4 * (j + g) + 2];
                const float4 K = pf[4 * (j + g) + 3];  // plain shared load with a uniform index (an asm "r" operand would force j into a vector register)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 cx = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                    const float2 cy = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                    const float2 cz = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
                    const float2 k = h ? make_float2(K.z, K.w) : make_float2(K.x, K.y);
                    const float2 A = fma2(cz, bc(dz), fma2(cy, bc(dy), fma2(cx, bc(dx), bc(nod))));
                    const float2 B = fma2(cz, bc(o2z), fma2(cy, bc(o2y), fma2(cx, bc(o2x), k)));
                    L[2 * g + h] = fma2(A, A, B);
                }
            }
            bool any = false;
#pragma unroll
            for (int p = 0; p < 8; ++p) any = any | (L[p].x > oo) | (L[p].y > oo);
            if (any) {
                uint32_t mask = 0u;
#pragma unroll
                for (int p = 7; p >= 0; --p) {
                    const float2 d = fma2(L[p], bc(-1.0f), bc(oo));
                    mask = __funnelshift_l(__float_as_uint(d.y), mask, 1);
                    mask = __funnelshift_l(__float_as_uint(d.x), mask, 1);
                }
                const uint32_t entry = ((uint32_t)(j / 4) << 16) | mask;
                if (cnt < kQueueCap) {
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + (uint32_t)cnt * (uint32_t)(kThreads * 4)), "r"(entry) : "memory");
                    cnt += 1;
                }
            }
        }
        total += cnt;
        ox += 1e-3f * dx; oy += 1e-3f * dy; oz += 1e-3f * dz;
    }
    out[gid] = total;
}

// ---------------- two rays per lane ----------------
// GB = blocks (of 4 spheres) per group; the flag word of a group holds GB*4 bits per ray
template <int MINB, int GB, int ORDER>
__global__ void __launch_bounds__(kThreads, MINB) k_two(const float4* __restrict__ pfg, int n_blocks, const float* __restrict__ rays, int trips, unsigned* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* pf = reinterpret_cast<float4*>(smem);
    uint32_t* q = reinterpret_cast<uint32_t*>(smem + (size_t)n_blocks * 64) + threadIdx.x;
    for (int i = threadIdx.x; i < n_blocks * 4; i += kThreads) pf[i] = pfg[i];
    __syncthreads();
    const int gid = blockIdx.x * kThreads + threadIdx.x;
    const int n_lanes = gridDim.x * kThreads;
    float2 ox, oy, oz, dx, dy, dz;  // .x = ray A, .y = ray B
    ox = make_float2(rays[gid * 6 + 0], rays[(gid + n_lanes) * 6 + 0]);
    oy = make_float2(rays[gid * 6 + 1], rays[(gid + n_lanes) * 6 + 1]);
    oz = make_float2(rays[gid * 6 + 2], rays[(gid + n_lanes) * 6 + 2]);
    dx = make_float2(rays[gid * 6 + 3], rays[(gid + n_lanes) * 6 + 3]);
    dy = make_float2(rays[gid * 6 + 4], rays[(gid + n_lanes) * 6 + 4]);
    dz = make_float2(rays[gid * 6 + 5], rays[(gid + n_lanes) * 6 + 5]);
    unsigned total = 0;
    for (int t = 0; t < trips; ++t) {
        float2 nod, oo, o2x, o2y, o2z;
        nod.x = -((ox.x * dx.x + oy.x * dy.x) + oz.x * dz.x);
        nod.y = -((ox.y * dx.y + oy.y * dy.y) + oz.y * dz.y);
        oo.x = ((ox.x * ox.x + oy.x * oy.x) + oz.x * oz.x) * (1.0f - 1.9073486328125e-06f);
        oo.y = ((ox.y * ox.y + oy.y * oy.y) + oz.y * oz.y) * (1.0f - 1.9073486328125e-06f);
        o2x = make_float2(ox.x + ox.x, ox.y + ox.y);
        o2y = make_float2(oy.x + oy.x, oy.y + oy.y);
        o2z = make_float2(oz.x + oz.x, oz.y + oz.y);
        u64 &rnod = *(u64*)&nod, &ro2x = *(u64*)&o2x, &ro2y = *(u64*)&o2y, &ro2z = *(u64*)&o2z, &rdx = *(u64*)&dx, &rdy = *(u64*)&dy, &rdz = *(u64*)&dz;
        asm volatile("" : "+l"(rnod), "+l"(ro2x), "+l"(ro2y), "+l"(ro2z), "+l"(rdx), "+l"(rdy), "+l"(rdz), "+r"(n_blocks));
        asm volatile("" : "+f"(oo.x), "+f"(oo.y));
        uint32_t addr = (uint32_t)__cvta_generic_to_shared(pf);
        asm volatile("" : "+r"(addr));
        uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(q);
        asm volatile("" : "+r"(qaddr));
        const uint32_t base = addr, end = addr + 64u * (uint32_t)n_blocks;
        int cnt = 0;
#pragma unroll 1
        for (; addr < end; addr += 64u * GB) {
            float2 L[4 * GB];
#pragma unroll
            for (int g = 0; g < GB; ++g) {
                const float4 X = lds128(addr + 64u * g), Y = lds128(addr + 64u * g + 16u), Z = lds128(addr + 64u * g + 32u), K = lds128(addr + 64u * g + 48u);
                const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w}, ks[4] = {K.x, K.y, K.z, K.w};
                if (ORDER == 0) {  // sphere-major: the 7 FMAs of one sphere back to back
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 A = fma2(bc(zs[e]), dz, fma2(bc(ys[e]), dy, fma2(bc(xs[e]), dx, nod)));
                        const float2 B = fma2(bc(zs[e]), o2z, fma2(bc(ys[e]), o2y, fma2(bc(xs[e]), o2x, bc(ks[e]))));
                        L[4 * g + e] = fma2(A, A, B);
                    }
                } else {  // operand-major: one ray operand serves 4 consecutive FMAs (reuse cache)
                    float2 A[4], B[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) A[e] = fma2(bc(xs[e]), dx, nod);
#pragma unroll
                    for (int e = 0; e < 4; ++e) B[e] = fma2(bc(xs[e]), o2x, bc(ks[e]));
#pragma unroll
                    for (int e = 0; e < 4; ++e) A[e] = fma2(bc(ys[e]), dy, A[e]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) B[e] = fma2(bc(ys[e]), o2y, B[e]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) A[e] = fma2(bc(zs[e]), dz, A[e]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) B[e] = fma2(bc(zs[e]), o2z, B[e]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) L[4 * g + e] = fma2(A[e], A[e], B[e]);
                }
            }
            float mA = L[0].x, mB = L[0].y;
#pragma unroll
            for (int p = 1; p < 4 * GB; ++p) { mA = fmaxf(mA, L[p].x); mB = fmaxf(mB, L[p].y); }
            if ((mA > oo.x) | (mB > oo.y)) {
                uint32_t maskA = 0u, maskB = 0u;
#pragma unroll
                for (int p = 4 * GB - 1; p >= 0; --p) {
                    const float2 d = fma2(L[p], bc(-1.0f), oo);
                    maskA = __funnelshift_l(__float_as_uint(d.x), maskA, 1);
                    maskB = __funnelshift_l(__float_as_uint(d.y), maskB, 1);
                }
                const uint32_t entry = ((((addr - base) >> 6) / GB) << 16) ^ maskA ^ (maskB << (GB == 4 ? 0 : 8));  // probe: one word
                if (cnt < kQueueCap) {
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + (uint32_t)cnt * (uint32_t)(kThreads * 4)), "r"(entry) : "memory");
                    cnt += 1;
                }
            }
        }
        total += cnt;
        ox.x += 1e-3f * dx.x; oy.x += 1e-3f * dy.x; oz.x += 1e-3f * dz.x;
        ox.y += 1e-3f * dx.y; oy.y += 1e-3f * dy.y; oz.y += 1e-3f * dz.y;
    }
    out[gid] = total;
}

struct Result { float ms; double tests; };
template <typename K>
Result run(K k, int ctas_per_sm, int sms, int rays_per_lane, const float4* d_pf, int n_blocks, const float* d_rays, unsigned* d_out, int trips) {
    const size_t smem = (size_t)n_blocks * 64 + kQueueCap * kThreads * 4;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kThreads, smem));
    const int blocks = sms * (occ < ctas_per_sm ? occ : ctas_per_sm);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<<<blocks, kThreads, smem>>>(d_pf, n_blocks, d_rays, trips, d_out); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); k<<<blocks, kThreads, smem>>>(d_pf, n_blocks, d_rays, trips, d_out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    Result res; res.ms = best; res.tests = (double)blocks * kThreads * rays_per_lane * trips * (double)n_blocks * 4;
    printf("   occupancy %d CTAs/SM (asked %d), grid %d", occ, ctas_per_sm, blocks);
    return res;
}
template <typename K>
Result run_param(K k, int ctas_per_sm, int sms, const PfParam& cp, const float4* d_pf, int n_blocks, const float* d_rays, unsigned* d_out, int trips) {
    const size_t smem = (size_t)n_blocks * 64 + kQueueCap * kThreads * 4;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kThreads, smem));
    const int blocks = sms * (occ < ctas_per_sm ? occ : ctas_per_sm);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<<<blocks, kThreads, smem>>>(cp, d_pf, n_blocks, d_rays, trips, d_out); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); k<<<blocks, kThreads, smem>>>(cp, d_pf, n_blocks, d_rays, trips, d_out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    Result res; res.ms = best; res.tests = (double)blocks * kThreads * trips * (double)n_blocks * 4;
    printf("   occupancy %d CTAs/SM (asked %d), grid %d", occ, ctas_per_sm, blocks);
    return res;
}
void report(const char* name, Result r, int sms, double clk_hz, unsigned cand) {
    const double tests_per_s = r.tests / (r.ms * 1e-3);
    const double peak = sms * 128.0 * 2.0 * clk_hz;
    printf("\n%-34s %.3f ms  %.1f Gtests/s  %.2f clk/test/SMSP(warp)  %.1f%% of FP32 peak at 16 flop/test  [pushes %u]\n", name, r.ms, tests_per_s / 1e9,
           clk_hz * (sms * 4.0) * 32.0 / tests_per_s, 100.0 * tests_per_s * 16.0 / peak, cand);
}

int main(int argc, char** argv) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); const int sms = prop.multiProcessorCount;
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double clk_hz = khz * 1e3;
    const int n_spheres = 512, n_blocks = n_spheres / 4, trips = 64;
    // scene: the RTIOW layout — a grid of small spheres on a big ground sphere
    std::vector<float> pf((size_t)n_blocks * 16);
    srand(1);
    auto rnd = []() { return (float)rand() / (float)RAND_MAX; };
    for (int j = 0; j < n_blocks; ++j) for (int e = 0; e < 4; ++e) {
        const int i = j * 4 + e;
        double cx = (i % 22) - 11 + 0.9 * rnd(), cy = 0.2, cz = (i / 22) - 11 + 0.9 * rnd(), r = 0.2;
        if (i == 0) { cx = 0; cy = -1000; cz = 0; r = 1000; }
        const double c2 = cx * cx + cy * cy + cz * cz, r2 = r * r;
        pf[j * 16 + 0 + e] = (float)cx; pf[j * 16 + 4 + e] = (float)cy; pf[j * 16 + 8 + e] = (float)cz;
        pf[j * 16 + 12 + e] = (float)(r2 - c2 + 1.9073486328125e-06 * (c2 + r2));
    }
    for (int mode = 0; mode < 2; ++mode) {  // 0: rays that flag nothing (pure loop), 1: camera-like rays (realistic push rate)
        const int max_lanes = sms * 4 * kThreads;
        std::vector<float> rays((size_t)max_lanes * 2 * 6);
        for (size_t i = 0; i < rays.size() / 6; ++i) {
            float ox = 13, oy = 2, oz = 3, dx, dy, dz;
            if (mode == 0) { ox = 0; oy = 50; oz = 0; dx = 0.1f * (rnd() - 0.5f); dy = 1; dz = 0.1f * (rnd() - 0.5f); }
            else { dx = -13 + 8 * (rnd() - 0.5f); dy = -2 + 4 * (rnd() - 0.5f); dz = -3 + 8 * (rnd() - 0.5f); }
            const float l = std::sqrt(dx * dx + dy * dy + dz * dz);
            rays[i * 6 + 0] = ox; rays[i * 6 + 1] = oy; rays[i * 6 + 2] = oz; rays[i * 6 + 3] = dx / l; rays[i * 6 + 4] = dy / l; rays[i * 6 + 5] = dz / l;
        }
        float4* d_pf; float* d_rays; unsigned* d_out;
        CK(cudaMalloc(&d_pf, pf.size() * 4)); CK(cudaMemcpy(d_pf, pf.data(), pf.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpyToSymbol(c_pf, pf.data(), pf.size() * 4));
        CK(cudaMalloc(&d_rays, rays.size() * 4)); CK(cudaMemcpy(d_rays, rays.data(), rays.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_out, (size_t)max_lanes * 4));
        printf("==== mode %d (%s), %d spheres, %d trips, SM clock %.0f MHz, %d SMs\n", mode, mode ? "camera-like rays" : "no candidates", n_spheres, trips, clk_hz / 1e6, sms);
        auto pushes = [&](int lanes) { std::vector<unsigned> h(lanes); CK(cudaMemcpy(h.data(), d_out, (size_t)lanes * 4, cudaMemcpyDeviceToHost)); unsigned long long s = 0; for (unsigned v : h) s += v; return (unsigned)(s / (lanes ? lanes : 1)); };
#define RUN1(K, CT, NAME) { Result r = run(K, CT, sms, 1, d_pf, n_blocks, d_rays, d_out, trips); report(NAME, r, sms, clk_hz, pushes(sms * kThreads)); }
#define RUN2(K, CT, NAME) { Result r = run(K, CT, sms, 2, d_pf, n_blocks, d_rays, d_out, trips); report(NAME, r, sms, clk_hz, pushes(sms * kThreads)); }
        RUN1(k_one<3>, 3, "one ray/lane, 3 CTAs/SM");
        RUN1(k_one<2>, 2, "one ray/lane, 2 CTAs/SM");
        RUN1(k_const<3>, 3, "one ray/lane const-bank UR, 3 CTAs/SM");
        RUN1(k_const<2>, 2, "one ray/lane const-bank UR, 2 CTAs/SM");
        RUN1(k_const_diag<3>, 3, "DIAG const-bank, X plane only, 3 CTAs/SM");
        RUN2(k_const2<2>, 2, "two rays/lane const-bank UR, 2 CTAs/SM");
        RUN2(k_const2<3>, 3, "two rays/lane const-bank UR, 3 CTAs/SM");
        RUN2(k_lds2<2>, 2, "two rays/lane LDS sphere pairs, 2 CTAs/SM");
        RUN2(k_lds2<3>, 3, "two rays/lane LDS sphere pairs, 3 CTAs/SM");
        { PfParam cp; memcpy(cp.v, pf.data(), sizeof(cp.v)); Result r = run_param(k_param<3>, 3, sms, cp, d_pf, n_blocks, d_rays, d_out, trips); report("one ray/lane kernel-param UR, 3 CTAs/SM", r, sms, clk_hz, pushes(sms * kThreads)); }
        RUN2((k_two<2, 4, 0>), 2, "two rays/lane g16 sphere-major 2CTA");
        RUN2((k_two<2, 4, 1>), 2, "two rays/lane g16 operand-major 2CTA");
        RUN2((k_two<2, 2, 0>), 2, "two rays/lane g8 sphere-major 2CTA");
        RUN2((k_two<2, 2, 1>), 2, "two rays/lane g8 operand-major 2CTA");
        RUN2((k_two<3, 2, 1>), 3, "two rays/lane g8 operand-major 3CTA");
        RUN2((k_two<1, 4, 1>), 1, "two rays/lane g16 operand-major 1CTA");
        cudaFree(d_pf); cudaFree(d_rays); cudaFree(d_out);
    }
    return 0;
}
