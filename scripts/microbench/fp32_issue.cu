// Measures FP32 issue rates on sm_100a: scalar FFMA/FADD vs packed FFMA2/FADD2 (fma.rn.f32x2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_issue fp32_issue.cu ; run on a B200.
// Output: warp-instructions / cycle / SM and flop/cycle/SM for each variant (JSON lines).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;  // independent dependency chains per thread

template <int MODE>
__global__ void __launch_bounds__(1024, 1) issue_kernel(float* out, float s0, float s1, long long* cyc) {
    float a[CH], b[CH];
    unsigned long long A[CH], B[CH];
    for (int i = 0; i < CH; i++) {
        a[i] = threadIdx.x * 0.001f + i; b[i] = 1.0f + i * 1e-3f;
        float2 t = make_float2(a[i], b[i]);
        A[i] = *reinterpret_cast<unsigned long long*>(&t);
        t = make_float2(b[i], a[i]);
        B[i] = *reinterpret_cast<unsigned long long*>(&t);
    }
    float2 w2 = make_float2(s0, s1);
    unsigned long long W = *reinterpret_cast<unsigned long long*>(&w2);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s0), "f"(b[i]));
            if (MODE == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(W), "l"(B[i]));
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B[i]));
            if (MODE == 4) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            if (MODE == 5) {  // FFMA with both multiplicands in per-thread registers (3 distinct regs + dst)
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) % CH]));
            }
            if (MODE == 6) {  // packed with 3 per-thread register pairs
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(B[i]), "l"(B[(i + 1) % CH]));
            }
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < CH; i++) {
        float2 t = *reinterpret_cast<float2*>(&A[i]);
        acc += a[i] + t.x + t.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int flop_per_instr, int threads) {
    int dev_sms = 0;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * dev_sms * 1024);
    cudaMalloc(&cyc, sizeof(long long) * dev_sms);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) issue_kernel<MODE><<<dev_sms, threads>>>(out, 1.0001f, 0.9999f, cyc);
    cudaEventRecord(e0);
    issue_kernel<MODE><<<dev_sms, threads>>>(out, 1.0001f, 0.9999f, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * dev_sms, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < dev_sms; i++) c += h[i]; c /= dev_sms;
    double winstr = double(ITERS) * CH * (threads / 32);
    printf("{\"variant\": \"%s\", \"warps_per_sm\": %d, \"warp_instr_per_cycle_per_sm\": %.3f, "
           "\"flop_per_cycle_per_sm\": %.1f, \"cycles\": %.0f, \"ms\": %.4f, \"implied_mhz\": %.0f, \"err\": \"%s\"}\n",
           name, threads / 32, winstr / c, winstr / c * 32 * flop_per_instr, c, ms, c / ms / 1e3,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {256, 512, 1024}) {
        run<0>("FFMA r,r,U,r", 2, threads);
        run<5>("FFMA r,r,r,r", 2, threads);
        run<1>("FADD", 1, threads);
        run<4>("FMUL", 1, threads);
        run<2>("FFMA2 r,r,U,r", 4, threads);
        run<6>("FFMA2 r,r,r,r", 4, threads);
        run<3>("FADD2", 2, threads);
    }
    return 0;
}
