// probe_forms.cu — cost (clk per warp-instruction per SM sub-partition) of the packed-FP32 instruction forms the
// sweep can be built from.  Each kernel runs CH independent chains so latency is hidden; 8 warps per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_forms probe_forms.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*(u64*)&a), "l"(*(u64*)&b), "l"(*(u64*)&c)); return *(float2*)&d; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return *(float2*)&d; }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return *(float2*)&d; }
constexpr int CH = 8;
#define PROLOG float2 acc[CH], p[CH], q[CH]; float s[CH]; \
  _Pragma("unroll") for (int i = 0; i < CH; ++i) { acc[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + i + 1]); p[i] = make_float2(in[threadIdx.x + 2 * i + 3], in[threadIdx.x + i + 9]); q[i] = make_float2(in[threadIdx.x + 3 * i + 5], in[threadIdx.x + i + 17]); s[i] = in[threadIdx.x + 5 * i + 2]; }
#define EPILOG float r = 0; _Pragma("unroll") for (int i = 0; i < CH; ++i) r += acc[i].x + acc[i].y + p[i].x + q[i].y + s[i]; out[blockIdx.x * blockDim.x + threadIdx.x] = r;

__global__ void k_fma2_ppp(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = fma2(acc[i], p[i], q[i]); } EPILOG }
__global__ void k_fma2_psp(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = fma2(p[i], make_float2(s[i], s[i]), acc[i]); } EPILOG }   // pair * scalar(distinct per chain) + pair
__global__ void k_fma2_psp_shared(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = fma2(p[i], make_float2(s[0], s[0]), acc[i]); } EPILOG }   // same scalar in all chains (reuse-able)
__global__ void k_fma2_pup(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = fma2(p[i], make_float2(u0, u0), acc[i]); } EPILOG }       // pair * uniform scalar + pair
__global__ void k_fma2_aac(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = fma2(p[i], p[i], acc[i]); } EPILOG }                      // pair^2 + pair
__global__ void k_fma2_aau(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = fma2(acc[i], acc[i], make_float2(u0, u1)); } EPILOG }     // pair^2 + uniform pair
__global__ void k_add2_ps(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = add2(acc[i], make_float2(s[i], s[i])); } EPILOG }
__global__ void k_add2_pu(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = add2(acc[i], make_float2(u0, u0)); } EPILOG }
__global__ void k_add2_pp(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = add2(acc[i], p[i]); } EPILOG }
__global__ void k_mul2_ps(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = mul2(acc[i], make_float2(s[i], s[i])); } EPILOG }
__global__ void k_mul2_aa(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = mul2(acc[i], acc[i]); } EPILOG }
// round 2: forms for two rays per lane — sphere scalar x ray pair + acc pair
__global__ void k_fma2_sPp(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = fma2(make_float2(s[i], s[i]), p[0], acc[i]); } EPILOG }   // scalar(distinct) * pair(shared by all chains) + pair
__global__ void k_fma2_Psp(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = fma2(p[0], make_float2(s[i], s[i]), acc[i]); } EPILOG }   // same, pair in slot a
__global__ void k_fma2_spp(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = fma2(make_float2(s[i], s[i]), p[i], acc[i]); } EPILOG }   // scalar * pair(distinct) + pair
__global__ void k_fma2_sPs(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = fma2(make_float2(s[i], s[i]), p[0], make_float2(acc[i].x, acc[i].x)); } EPILOG }   // scalar * shared pair + scalar
__global__ void k_fma2_sup(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = fma2(make_float2(s[i], s[i]), make_float2(u0, u1), acc[i]); } EPILOG }   // scalar * uniform pair + pair
// scalar forms
__global__ void k_ffma_rrr(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) { acc[i].x = __fmaf_rn(acc[i].x, p[i].x, q[i].x); acc[i].y = __fmaf_rn(acc[i].y, p[i].y, q[i].y); } } EPILOG }
__global__ void k_ffma_rur(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) { acc[i].x = __fmaf_rn(p[i].x, u0, acc[i].x); acc[i].y = __fmaf_rn(p[i].y, u1, acc[i].y); } } EPILOG }
__global__ void k_ffma_rrr2(float* out, const float* in, int iters, float u0, float u1) { PROLOG for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) { acc[i].x = __fmaf_rn(p[i].x, p[i].x, acc[i].x); acc[i].y = __fmaf_rn(p[i].y, p[i].y, acc[i].y); } } EPILOG }
__global__ void k_fsetp(float* out, const float* in, int iters, float u0, float u1) { PROLOG int cnt = 0; for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int i = 0; i < CH; ++i) { cnt += (acc[i].x > p[i].x) ? 1 : 0; acc[i].x += 1.0f; } } EPILOG out[0] = cnt; }

template <typename K> void run(const char* name, K k, int sms, float* d_out, float* d_in, int per_iter, int fma_per_instr) {
    const int iters = 4096, threads = 256, blocks = sms * 4;  // 32 warps/SM = 8 per SMSP
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<<<blocks, threads>>>(d_out, d_in, iters, 1.0001f, 0.9999f); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(e0)); k<<<blocks, threads>>>(d_out, d_in, iters, 1.0001f, 0.9999f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    double warp_instr_per_smsp = (double)iters * per_iter * 8;  // 8 warps per SMSP
    double clk = best * 1e-3 * 1.965e9;
    printf("%-22s %.3f ms  %.2f clk per warp-instr per SMSP  (%.1f%% of FP32 peak if every instr were %d FMA/lane)\n", name, best, clk / warp_instr_per_smsp,
           100.0 * fma_per_instr * 32 * warp_instr_per_smsp / clk / 32.0, fma_per_instr);
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); int sms = prop.multiProcessorCount;
    float *d_out, *d_in; CK(cudaMalloc(&d_out, 4 * sms * 4 * 256)); CK(cudaMalloc(&d_in, 4 * 4096)); CK(cudaMemset(d_in, 0, 4 * 4096));
    run("FFMA2 pair,pair,pair", k_fma2_ppp, sms, d_out, d_in, CH, 2);
    run("FFMA2 pair,scal,pair", k_fma2_psp, sms, d_out, d_in, CH, 2);
    run("FFMA2 pair,scal*,pair", k_fma2_psp_shared, sms, d_out, d_in, CH, 2);
    run("FFMA2 pair,unif,pair", k_fma2_pup, sms, d_out, d_in, CH, 2);
    run("FFMA2 a,a,pair", k_fma2_aac, sms, d_out, d_in, CH, 2);
    run("FFMA2 a,a,unifpair", k_fma2_aau, sms, d_out, d_in, CH, 2);
    run("FADD2 pair,scal", k_add2_ps, sms, d_out, d_in, CH, 2);
    run("FADD2 pair,unif", k_add2_pu, sms, d_out, d_in, CH, 2);
    run("FADD2 pair,pair", k_add2_pp, sms, d_out, d_in, CH, 2);
    run("FMUL2 pair,scal", k_mul2_ps, sms, d_out, d_in, CH, 2);
    run("FMUL2 a,a", k_mul2_aa, sms, d_out, d_in, CH, 2);
    run("FFMA2 scal,PAIR*,pair", k_fma2_sPp, sms, d_out, d_in, CH, 2);
    run("FFMA2 PAIR*,scal,pair", k_fma2_Psp, sms, d_out, d_in, CH, 2);
    run("FFMA2 scal,pair,pair", k_fma2_spp, sms, d_out, d_in, CH, 2);
    run("FFMA2 scal,PAIR*,scal", k_fma2_sPs, sms, d_out, d_in, CH, 2);
    run("FFMA2 scal,unifpair,pair", k_fma2_sup, sms, d_out, d_in, CH, 2);
    run("FFMA r,r,r (x2)", k_ffma_rrr, sms, d_out, d_in, 2 * CH, 1);
    run("FFMA r,unif,r (x2)", k_ffma_rur, sms, d_out, d_in, 2 * CH, 1);
    run("FFMA a,a,r (x2)", k_ffma_rrr2, sms, d_out, d_in, 2 * CH, 1);
    run("FSETP+FADD", k_fsetp, sms, d_out, d_in, 2 * CH, 1);
    return 0;
}
