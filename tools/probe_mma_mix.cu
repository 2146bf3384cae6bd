// probe_mma_mix.cu — does the legacy HMMA (mma.sync) issue overlap with FP32 work of the same SM sub-partition?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_mma_mix probe_mma_mix.cu
// Per iteration: 8 independent HMMA.16816.F32 (zero accumulator) + NF independent scalar FFMAs, 6 warps per SMSP.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void hmma(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "f"(0.0f));
}
template <int NM, int NF>
__global__ void __launch_bounds__(256) k_mix(float* out, const uint32_t* in, int iters) {
    uint4 a = make_uint4(in[threadIdx.x], in[threadIdx.x + 1], in[threadIdx.x + 2], in[threadIdx.x + 3]);
    uint32_t b[8][2];
    for (int i = 0; i < 8; ++i) { b[i][0] = in[threadIdx.x + 4 + i]; b[i][1] = in[threadIdx.x + 12 + i]; }
    float f[16], acc = 0.0f;
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(in[threadIdx.x + 20 + i]);
    const float u = __uint_as_float(in[0]), v = __uint_as_float(in[1]);
    for (int it = 0; it < iters; ++it) {
        float d[8][4];
#pragma unroll
        for (int i = 0; i < NM; ++i) hmma(d[i], a, b[i][0], b[i][1]);
#pragma unroll
        for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i & 15]) : "f"(u), "f"(v));
#pragma unroll
        for (int i = 0; i < NM; ++i) acc += d[i][0];  // one consumer per MMA (FADD)
        a.x ^= (uint32_t)it & 1u;
    }
    for (int i = 0; i < 16; ++i) acc += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int NM, int NF>
void run(float* d_out, uint32_t* d_in, int sms, double clk) {
    const int iters = 4000, ctas = 3;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0)); k_mix<NM, NF><<<sms * ctas, 256>>>(d_out, d_in, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
    }
    const double per_iter = best * 1e-3 * clk / ((double)iters * ctas * 8 / 4.0);
    printf("%d HMMA + %2d FFMA (+%d FADD) per iteration: %.1f clk per iteration per SMSP\n", NM, NF, NM, per_iter);
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); int khz; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    uint32_t* d_in; float* d_out; CK(cudaMalloc(&d_in, 4096 * 4)); CK(cudaMemset(d_in, 0x3c, 4096 * 4)); CK(cudaMalloc(&d_out, (size_t)prop.multiProcessorCount * 3 * 256 * 4));
    const int sms = prop.multiProcessorCount; const double clk = khz * 1e3;
    run<8, 0>(d_out, d_in, sms, clk); run<8, 16>(d_out, d_in, sms, clk); run<8, 32>(d_out, d_in, sms, clk); run<8, 64>(d_out, d_in, sms, clk);
    run<0, 64>(d_out, d_in, sms, clk); run<4, 32>(d_out, d_in, sms, clk); run<0, 32>(d_out, d_in, sms, clk);
    return 0;
}
