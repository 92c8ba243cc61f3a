// Issue-rate microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

__global__ void k_scalar(float *out, float a, float b) {
    float x[2 * ILP];
    for (int i = 0; i < 2 * ILP; i++) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 2 * ILP; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
    for (int i = 0; i < 2 * ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_packed(float *out, float a, float b) {
    unsigned long long x[ILP], aa, bb;
    float2 av = make_float2(a, a), bv = make_float2(b, b);
    aa = *reinterpret_cast<unsigned long long *>(&av);
    bb = *reinterpret_cast<unsigned long long *>(&bv);
    for (int i = 0; i < ILP; i++) {
        float2 v = make_float2(threadIdx.x * 1e-3f + 2 * i, threadIdx.x * 1e-3f + 2 * i + 1);
        x[i] = *reinterpret_cast<unsigned long long *>(&v);
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
    }
    float s = 0;
    for (int i = 0; i < ILP; i++) {
        float2 v = *reinterpret_cast<float2 *>(&x[i]);
        s += v.x + v.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k_scalar<<<sms * 8, 256>>>(out, 0.999f, 0.001f);
            else k_packed<<<sms * 8, 256>>>(out, 0.999f, 0.001f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        double fmas = (double)sms * 8 * 256 * ITERS * 2 * ILP;
        printf("%s: %.3f ms  %.2f TFLOP/s  (%.1f fma lanes/clk/SM at 1.9 GHz)\n",
               mode ? "FFMA2 (f32x2)" : "FFMA scalar ", best, 2 * fmas / best / 1e9,
               fmas / (best * 1e-3) / sms / 1.9e9);
    }
    return 0;
}
