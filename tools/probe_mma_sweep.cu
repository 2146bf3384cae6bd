// probe_mma_sweep.cu — prototype of the pre-filter on the warp-level tensor path (mma.sync m16n8k16 f16, HMMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o probe_mma_sweep probe_mma_sweep.cu
//
// The pre-filter of pt_sweep.cuh is two dot products per (sphere, ray) test followed by one FMA and a sign test:
//   A' = S . R_A,  B' = S . R_B,  L' = A'*A' + B' > 0      with (sigma, s: powers of two, see below)
//   S   = [sigma cx, sigma cy, sigma cz, s, K'/s],   K' = sigma^2 (r^2 - |c|^2 + slack (|c|^2 + r^2)) + abs_slack
//   R_A = [dx, dy, dz, sigma (-o.d) / s, 0],   R_B = [2 sigma ox, 2 sigma oy, 2 sigma oz, -sigma^2 |o|^2 (1 - slack) / s, s]
// so L' = sigma^2 (disc + slack terms): the same conservative test as the FP32 loop.  Every f32 element is split into two f16
// pieces (hi = rn(x), lo = rn(x - hi): 22 significant bits) and the products hi*hi + hi*lo + lo*hi are laid along K:
//   K index 0..4 S_hi*R_hi, 5..9 S_hi*R_lo, 10..14 S_lo*R_hi, 15 unused   ->  ONE m16n8k16 MMA per dot product per 16 spheres x 8 rays.
// sigma scales the scene into the f16 range (|sigma x| <= 16384), s keeps K' and |o|^2 there.
//
// The probe runs the whole would-be sweep (sphere fragments by non-broadcast LDS.128, ray fragments through a shared-memory
// transpose, 8 HMMA + 8 FFMA2 + max + vote per 512 tests, ballots -> per-ray 16-bit masks -> per-lane queue) and reports
//   * clk per 32 tests per SM sub-partition (the FP32 loop of the shipped kernel: 11.0-11.5, tools/probe_sweep2.cu),
//   * conservativeness against the reference's exact unfused f32 discriminant on every (ray, sphere) pair of a sample,
//   * candidates per ray next to the FP32 filter's own count (same slack rule at 2^-18).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
typedef unsigned long long u64;
constexpr int kThreads = 256;
constexpr int kQueueCap = 12;

__device__ __forceinline__ void hmma(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "f"(0.0f));
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*(u64*)&a), "l"(*(u64*)&b), "l"(*(u64*)&c)); return *(float2*)&d; }
__device__ __forceinline__ void split16(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack(__half a, __half b) { return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16); }

struct Scale { float sigma, s, inv_s, slack; };

// the 16 words (A column: words 0..7, B column: words 8..15) of one ray's operand
__device__ __forceinline__ void ray_operand(const Scale sc, float ox, float oy, float oz, float dx, float dy, float dz, uint32_t (&w)[16]) {
    const float nod = -((ox * dx + oy * dy) + oz * dz);
    const float oo = ((ox * ox + oy * oy) + oz * oz) * (1.0f - sc.slack);
    const float ra[5] = {dx, dy, dz, sc.sigma * nod * sc.inv_s, 0.0f};
    const float rb[5] = {2.0f * sc.sigma * ox, 2.0f * sc.sigma * oy, 2.0f * sc.sigma * oz, -(sc.sigma * sc.sigma) * oo * sc.inv_s, sc.s};
    __half v[2][16];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int e = 0; e < 5; ++e) {
            __half hi, lo;
            split16(c ? rb[e] : ra[e], hi, lo);
            v[c][e] = hi; v[c][5 + e] = lo; v[c][10 + e] = hi;
        }
        v[c][15] = __float2half_rn(0.0f);
    }
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) w[c * 8 + i] = pack(v[c][2 * i], v[c][2 * i + 1]);
}

// ---- layout "rows = rays": D[ray][sphere].  Per MMA 16 rays x 8 spheres; the rays are the A operand (loop-invariant
// fragments), the spheres the B operand (one LDS.128 per lane per 16 spheres).  Lane (g, t) of the warp owns ray
// 16 (t >> 1) + 8 (t & 1) + g: the four rays whose results a quad holds are the four rays its lanes own, so a lane that
// finds a candidate pushes it to its OWN queue (no ballots, no atomics) and the owners later read the queues of their quad.
// Value v = (rb, sg, c): ray slot 2 rb + (c >> 1) (= the owner's t), sphere 16 step + 8 sg + 2 t + (c & 1); mask bit
// b = 8 rb + 4 (c >> 1) + 2 sg + (c & 1).  N = -(A'^2 + B') is what is computed: candidate <=> N < 0 <=> sign bit.
template <int MINB, bool VERIFY, int PIPE>
__global__ void __launch_bounds__(kThreads, MINB) k_rows(const uint4* __restrict__ img, int n_steps, const float* __restrict__ rays, int trips, Scale sc,
                                                         unsigned* __restrict__ out, uint16_t* __restrict__ bitmap) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint4* simg = reinterpret_cast<uint4*>(smem);                                // (n_steps + 1) x 32 uint4 (one step of padding for the prefetch)
    uint32_t* stage = reinterpret_cast<uint32_t*>(simg + (n_steps + 1) * 32);    // [16 words][kThreads]
    uint32_t* queue = stage + 16 * kThreads + threadIdx.x;                       // [kQueueCap][kThreads]
    for (int i = threadIdx.x; i < (n_steps + 1) * 32; i += blockDim.x) simg[i] = i < n_steps * 32 ? img[i] : make_uint4(0, 0, 0, 0);
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
        const int my_ray = 16 * (t >> 1) + 8 * (t & 1) + g;  // within the warp
    const int gid = blockIdx.x * blockDim.x + (threadIdx.x & ~31) + my_ray;
    float ox = rays[gid * 6 + 0], oy = rays[gid * 6 + 1], oz = rays[gid * 6 + 2], dx = rays[gid * 6 + 3], dy = rays[gid * 6 + 4], dz = rays[gid * 6 + 5];
    unsigned total = 0u, overflow = 0u;
    for (int trip = 0; trip < trips; ++trip) {
        // ray operands -> A fragments.  Fragment (quad g, row block rb, t, column type c) is four consecutive words
        // {ray0.word[t], ray1.word[t], ray0.word[t+4], ray1.word[t+4]} at row c*8 + rb*4 + t, column 4g of the warp's
        // 32 columns: the owner of ray (rb_w = t >> 1, half = t & 1) scatters its 16 words, every lane reads 4 x LDS.128.
        {
            uint32_t w[16];
            ray_operand(sc, ox, oy, oz, dx, dy, dz, w);
            uint32_t* base = stage + (threadIdx.x & ~31) + 4 * g + (t & 1);
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int k = 0; k < 8; ++k) base[(c * 8 + (t >> 1) * 4 + (k & 3)) * kThreads + 2 * (k >> 2)] = w[c * 8 + k];
        }
        __syncwarp();
        uint4 fa[2], fb[2];
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
            fa[rb] = *reinterpret_cast<const uint4*>(stage + (rb * 4 + t) * kThreads + (threadIdx.x & ~31) + 4 * g);
            fb[rb] = *reinterpret_cast<const uint4*>(stage + (8 + rb * 4 + t) * kThreads + (threadIdx.x & ~31) + 4 * g);
        }
        __syncwarp();
        int cnt = 0;
        const uint4* p = simg + lane;
        uint4 sp = *p;
#pragma unroll 1
        for (int s = 0; s < n_steps; ++s) {
            p += 32;
            const uint4 cur = sp;
            float A[2][2][4], B[2][2][4];
#define HA(rb, sg) hmma(A[rb][sg], fa[rb], sg ? cur.z : cur.x, sg ? cur.w : cur.y)
#define HB(rb, sg) hmma(B[rb][sg], fb[rb], sg ? cur.z : cur.x, sg ? cur.w : cur.y)
            if (PIPE == 0) { HA(0, 0); HB(0, 0); HA(0, 1); HB(0, 1); HA(1, 0); HB(1, 0); HA(1, 1); HB(1, 1); }       // sphere operand shared by pairs
            else if (PIPE == 1) { HA(0, 0); HA(0, 1); HB(0, 0); HB(0, 1); HA(1, 0); HA(1, 1); HB(1, 0); HB(1, 1); }  // ray operand shared by pairs
            else { HA(0, 0); HA(0, 1); HB(0, 1); HB(0, 0); HA(1, 0); HA(1, 1); HB(1, 1); HB(1, 0); }                // snake: every MMA shares one operand with its predecessor
            sp = *p;  // next step's sphere fragments (the image carries one step of padding)
            float2 N[2][2][2];
            float m = 3.0e38f;
#pragma unroll
            for (int rb = 0; rb < 2; ++rb)
#pragma unroll
                for (int sg = 0; sg < 2; ++sg)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float2 a2 = make_float2(A[rb][sg][2 * h], A[rb][sg][2 * h + 1]);
                        N[rb][sg][h] = fma2(make_float2(-a2.x, -a2.y), a2, make_float2(-B[rb][sg][2 * h], -B[rb][sg][2 * h + 1]));
                        m = fminf(fminf(N[rb][sg][h].x, N[rb][sg][h].y), m);
                    }
            if (m < 0.0f) {  // lane-divergent: this lane holds a candidate
                uint32_t mask = 0u;  // bit b = 8 rb + 4 h + 2 sg + e, built from the highest bit down
#pragma unroll
                for (int rb = 1; rb >= 0; --rb)
#pragma unroll
                    for (int h = 1; h >= 0; --h)
#pragma unroll
                        for (int sg = 1; sg >= 0; --sg) {
                            mask = __funnelshift_l(__float_as_uint(N[rb][sg][h].y), mask, 1);
                            mask = __funnelshift_l(__float_as_uint(N[rb][sg][h].x), mask, 1);
                        }
                mask &= 0xffffu;
                if (VERIFY) {  // scatter the bits to the owners' bitmaps: [ray][step] 16 bits, bit = sphere within the step
                    for (int b = 0; b < 16; ++b) if ((mask >> b) & 1u) {
                        const int slot = b >> 2, sph = 8 * ((b >> 1) & 1) + 2 * t + (b & 1);
                        const int ray = 16 * (slot >> 1) + 8 * (slot & 1) + g;
                        atomicOr(reinterpret_cast<unsigned*>(bitmap) + (((size_t)(blockIdx.x * blockDim.x + (threadIdx.x & ~31) + ray) * n_steps + s) >> 1),
                                 (1u << sph) << (16 * ((((size_t)(blockIdx.x * blockDim.x + (threadIdx.x & ~31) + ray) * n_steps + s)) & 1)));
                    }
                }
                if (cnt < kQueueCap) { queue[cnt * kThreads] = ((unsigned)s << 16) | mask; cnt += 1; } else overflow += 1u;
                total += (unsigned)__popc(mask);
            }
        }
        ox += 1e-3f * dx; oy += 1e-3f * dy; oz += 1e-3f * dz;
        if (cnt > 0) total += queue[0] >> 31;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
    out[gridDim.x * blockDim.x + blockIdx.x * blockDim.x + threadIdx.x] = overflow;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); const int sms = prop.multiProcessorCount;
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double clk_hz = khz * 1e3;
    const int n_spheres = 512, n_steps = n_spheres / 16, trips = 64;
    srand(1);
    auto rnd = []() { return (float)rand() / (float)RAND_MAX; };
    // scene: the RTIOW layout — a grid of small spheres on a big ground sphere (same as probe_sweep2)
    std::vector<float> cx(n_spheres), cy(n_spheres), cz(n_spheres), rr(n_spheres);
    double extent = 0;
    for (int i = 0; i < n_spheres; ++i) {
        cx[i] = (float)((i % 22) - 11 + 0.9 * rnd()); cy[i] = 0.2f; cz[i] = (float)((i / 22) - 11 + 0.9 * rnd()); rr[i] = 0.2f;
        if (i == 0) { cx[i] = 0; cy[i] = -1000; cz[i] = 0; rr[i] = 1000; }
        if (i >= 1 && i <= 3) { cx[i] = -4.0f + 4.0f * (i - 1); cy[i] = 1.0f; cz[i] = 0.0f; rr[i] = 1.0f; }
        extent = fmax(extent, sqrt((double)cx[i] * cx[i] + (double)cy[i] * cy[i] + (double)cz[i] * cz[i]) + rr[i]);
    }
    extent = fmax(extent, 64.0);  // camera / ray origins
    Scale sc;
    sc.sigma = (float)exp2(floor(log2(16384.0 / extent)));
    sc.s = (float)exp2(ceil(log2((double)sc.sigma * sc.sigma * extent * extent / 32768.0)));
    if (sc.s < 1.0f) sc.s = 1.0f;
    sc.inv_s = 1.0f / sc.s;
    for (int slack_log2 = 15; slack_log2 <= 15; slack_log2 += 1) {
        sc.slack = (float)exp2(-slack_log2);
        const double abs_slack = 1.0 / 128.0;  // scaled units
        printf("==== extent %.0f, sigma %g, s %g, slack 2^-%d (+ %.3g absolute in scaled units)\n", extent, sc.sigma, sc.s, slack_log2, abs_slack);
        // sphere operand image in fragment order: step s, lane (g, t) -> a0 (row g, k 2t..), a1 (row g+8, k 2t..), a2 (row g, k 2t+8..), a3 (row g+8, k 2t+8..)
        std::vector<uint32_t> img((size_t)n_steps * 32 * 4);
        auto h16 = [](float x) { __half h = __float2half_rn(x); unsigned short u; memcpy(&u, &h, 2); return u; };
        auto h2f = [](unsigned short u) { __half h; memcpy(&h, &u, 2); return __half2float(h); };
        std::vector<unsigned short> rowv((size_t)n_spheres * 16);
        for (int i = 0; i < n_spheres; ++i) {
            const double c2 = (double)cx[i] * cx[i] + (double)cy[i] * cy[i] + (double)cz[i] * cz[i], r2 = (double)rr[i] * rr[i];
            const double Kp = (double)sc.sigma * sc.sigma * (r2 - c2 + (double)sc.slack * (c2 + r2)) + abs_slack;
            const float S[5] = {sc.sigma * cx[i], sc.sigma * cy[i], sc.sigma * cz[i], sc.s, (float)(Kp / sc.s)};
            for (int e = 0; e < 5; ++e) {
                const unsigned short hi = h16(S[e]);
                // round the K' piece UP so that hi + lo >= the f32 value (the slack must not be eaten by the split)
                unsigned short lo = h16(S[e] - h2f(hi));
                rowv[(size_t)i * 16 + e] = hi; rowv[(size_t)i * 16 + 5 + e] = hi; rowv[(size_t)i * 16 + 10 + e] = lo;
            }
            rowv[(size_t)i * 16 + 15] = 0;
        }
        for (int s = 0; s < n_steps; ++s) for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3;
            auto pk = [&](int row, int k) { return (uint32_t)rowv[(size_t)(s * 16 + row) * 16 + k] | ((uint32_t)rowv[(size_t)(s * 16 + row) * 16 + k + 1] << 16); };
            uint32_t* o = &img[((size_t)s * 32 + lane) * 4];
            o[0] = pk(g, 2 * t); o[1] = pk(g, 2 * t + 8); o[2] = pk(8 + g, 2 * t); o[3] = pk(8 + g, 2 * t + 8);
        }
        uint4* d_img; CK(cudaMalloc(&d_img, img.size() * 4)); CK(cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
        for (int mode = 0; mode < 3; ++mode) {  // 0: no candidates, 1: camera-like rays, 2: bounce-like rays (origins on spheres, random directions)
            const int max_lanes = sms * 5 * kThreads;
            std::vector<float> rays((size_t)max_lanes * 6);
            for (size_t i = 0; i < rays.size() / 6; ++i) {
                float ox = 13, oy = 2, oz = 3, dx, dy, dz;
                if (mode == 0) { ox = 0; oy = 50; oz = 0; dx = 0.1f * (rnd() - 0.5f); dy = 1; dz = 0.1f * (rnd() - 0.5f); }
                else if (mode == 1) { dx = -13 + 8 * (rnd() - 0.5f); dy = -2 + 4 * (rnd() - 0.5f); dz = -3 + 8 * (rnd() - 0.5f); }
                else {
                    const int k = 1 + rand() % (n_spheres - 1);
                    float nx = rnd() - 0.5f, ny = rnd() - 0.5f, nz = rnd() - 0.5f; float l = std::sqrt(nx * nx + ny * ny + nz * nz); nx /= l; ny /= l; nz /= l;
                    if (rand() % 3 == 0) { ox = (rnd() - 0.5f) * 22; oy = 0.0f; oz = (rnd() - 0.5f) * 22; nx = 0; ny = 1; nz = 0; }  // on the ground
                    else { ox = cx[k] + rr[k] * nx; oy = cy[k] + rr[k] * ny; oz = cz[k] + rr[k] * nz; }
                    dx = nx + (rnd() - 0.5f); dy = ny + (rnd() - 0.5f); dz = nz + (rnd() - 0.5f);
                }
                const float l = std::sqrt(dx * dx + dy * dy + dz * dz);
                rays[i * 6 + 0] = ox; rays[i * 6 + 1] = oy; rays[i * 6 + 2] = oz; rays[i * 6 + 3] = dx / l; rays[i * 6 + 4] = dy / l; rays[i * 6 + 5] = dz / l;
            }
            float* d_rays; unsigned* d_out; uint16_t* d_bm;
            CK(cudaMalloc(&d_rays, rays.size() * 4)); CK(cudaMemcpy(d_rays, rays.data(), rays.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMalloc(&d_out, (size_t)max_lanes * 8));
            const size_t smem = (size_t)(n_steps + 1) * 512 + 16 * kThreads * 4 + kQueueCap * kThreads * 4;
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            for (int st = 0; st < 3; ++st) for (int ctas = 2; ctas <= 3; ++ctas) {
                auto launch = [&](int blocks) {
#define L_(S, C) if (st == S && ctas == C) { CK(cudaFuncSetAttribute(k_rows<C, false, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_rows<C, false, S><<<blocks, kThreads, smem>>>(d_img, n_steps, d_rays, trips, sc, d_out, nullptr); }
                    L_(0, 2) L_(0, 3) L_(1, 2) L_(1, 3) L_(2, 2) L_(2, 3)
                };
                const int blocks = sms * ctas;
                launch(blocks); CK(cudaDeviceSynchronize());
                float best = 1e30f;
                for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(e0)); launch(blocks); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms); }
                std::vector<unsigned> h((size_t)max_lanes * 2); CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
                double cand = 0, ovf = 0; const int lanes = blocks * kThreads;
                for (int i = 0; i < lanes; ++i) { cand += h[i]; ovf += h[(size_t)lanes + i]; }
                const double tests = (double)lanes * trips * n_spheres;
                const double clk32 = best * 1e-3 * clk_hz * (sms * 4.0) / (tests / 32.0);
                printf("mode %d order %d, %d CTAs/SM: %.3f ms  %.1f Gtests/s  %.2f clk per 32 tests per SMSP  (x%.2f of the FP32 loop's 11.2)  candidates/ray %.2f, queue overflows/ray %.3f\n", mode, st, ctas,
                       best, tests / (best * 1e-3) / 1e9, clk32, 11.2 / clk32, cand / lanes / trips, ovf / lanes / trips);
            }
            // verification on the first CTAs: every sphere the reference's exact f32 expression accepts must be flagged
            {
                const int blocks = 64, lanes = blocks * kThreads;
                CK(cudaMalloc(&d_bm, (size_t)lanes * n_steps * 2)); CK(cudaMemset(d_bm, 0, (size_t)lanes * n_steps * 2));
                CK(cudaFuncSetAttribute(k_rows<1, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_rows<1, true, 1><<<blocks, kThreads, smem>>>(d_img, n_steps, d_rays, 1, sc, d_out, d_bm); CK(cudaDeviceSynchronize());
                std::vector<uint16_t> bm((size_t)lanes * n_steps); CK(cudaMemcpy(bm.data(), d_bm, bm.size() * 2, cudaMemcpyDeviceToHost));
                long long exact_hits = 0, missed = 0, flagged = 0, fp32_flagged = 0; double worst_margin = 1e30;
                for (int i = 0; i < lanes; ++i) {
                    const float* r = &rays[(size_t)i * 6];
                    for (int k = 0; k < n_spheres; ++k) {
                        const int s = k / 16, p = k % 16;
                        const bool flag = (bm[(size_t)i * n_steps + s] >> p) & 1;
                        flagged += flag;
                        // spheres_soa.rs:116-121, unfused f32
                        const float cox = cx[k] - r[0], coy = cy[k] - r[1], coz = cz[k] - r[2];
                        const float nb = (cox * r[3] + coy * r[4]) + coz * r[5];
                        const float c = ((cox * cox + coy * coy) + coz * coz) - rr[k] * rr[k];
                        const float disc = nb * nb - c;
                        // the FP32 filter's rule at slack 2^-18 in f64 (its candidate count, for comparison)
                        const double c2 = (double)cx[k] * cx[k] + (double)cy[k] * cy[k] + (double)cz[k] * cz[k], r2 = (double)rr[k] * rr[k], o2 = (double)r[0] * r[0] + (double)r[1] * r[1] + (double)r[2] * r[2];
                        const double co_d = ((double)cx[k] - r[0]) * r[3] + ((double)cy[k] - r[1]) * r[4] + ((double)cz[k] - r[2]) * r[5];
                        const double D = co_d * co_d - (((double)cx[k] - r[0]) * ((double)cx[k] - r[0]) + ((double)cy[k] - r[1]) * ((double)cy[k] - r[1]) + ((double)cz[k] - r[2]) * ((double)cz[k] - r[2])) + r2;
                        fp32_flagged += D + exp2(-18.0) * (c2 + r2 + o2) > 0;
                        if (disc > 0.0f) { exact_hits += 1; if (!flag) missed += 1; }
                        if (!flag) worst_margin = fmin(worst_margin, -D / (c2 + r2 + o2));  // how far below zero the nearest unflagged pair sits, in slack units
                    }
                }
                printf("mode %d verify: %d rays x %d spheres: exact-expression hits %lld, MISSED %lld, flagged %.3f per ray (FP32 filter at 2^-18: %.3f), nearest unflagged pair at -D/(|c|^2+r^2+|o|^2) = 2^%.1f\n", mode, lanes,
                       n_spheres, exact_hits, missed, (double)flagged / lanes, (double)fp32_flagged / lanes, log2(worst_margin));
                cudaFree(d_bm);
            }
            cudaFree(d_rays); cudaFree(d_out);
        }
        cudaFree(d_img);
    }
    return 0;
}
