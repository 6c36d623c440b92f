// Measures FP32 issue rates on sm_100a: scalar FFMA/FADD/FMUL vs packed FFMA2/FADD2/FMUL2
// (fma/add/mul.rn.f32x2) in the operand forms the harmonic-energy FFT uses, plus mixes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_issue.bin fp32_issue.cu ; run on a B200.
// Output: JSON lines; rates are warp-instructions / cycle / SM (4 schedulers => 4.0 is the issue limit)
// and the "pair rate" = complex-element operations per cycle per SM (a packed op = 1 pair, a scalar op = 1/2).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 1024;
constexpr int UNR = 8;  // loop body = UNR * CH instructions, 3 instructions of loop overhead
constexpr int CH = 8;   // independent dependency chains per thread

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

enum { FFMA_IMM, FFMA_RRR, FADD_RR, FMUL_RR, FFMA2_IMM, FFMA2_RRR, FADD2_RR, FMUL2_RR, FFMA2_ONE, FFMA2_SWZ,
       MIX_FFMA2_FADD, MIX_FFMA2_FADD2, FADD2_SWZ, FFMA2_BCAST, MIX_FFMA2_LDS, NMODES };
static const char* kNames[NMODES] = {
    "FFMA r,r,imm,r", "FFMA r,r,r,r", "FADD r,r,r", "FMUL r,r,r", "FFMA2 r,r,imm,r", "FFMA2 r,r,r,r", "FADD2 r,r,r",
    "FMUL2 r,r,r", "FFMA2 r,b,U(1.0),a (add via fma)", "FFMA2 r,r.LO_HI.NP,imm,r", "mix 1 FFMA2imm : 1 FADD",
    "mix 1 FFMA2imm : 1 FADD2", "FADD2 r,r,r.LO_HI.NP", "FFMA2 r,r,r.F32(bcast),r", "mix 4 FFMA2imm : 1 LDS.64"};
static const int kFlop[NMODES] = {2, 2, 1, 1, 4, 4, 2, 2, 4, 4, 0, 0, 2, 4, 0};

template <int MODE>
__global__ void __launch_bounds__(1024, 1) issue_kernel(float* out, float s0, float s1, float one, long long* cyc) {
    __shared__ u64 sm[1024];
    float a[CH], b[CH];
    u64 A[CH], B[CH];
    for (int i = 0; i < CH; i++) {
        a[i] = threadIdx.x * 0.001f + i; b[i] = s0 + i * 1e-3f;
        A[i] = pk(a[i], b[i]);
        B[i] = pk(b[i] * s1, a[i] * s0);
    }
    sm[threadIdx.x] = A[0];
    __syncthreads();
    const u64* sp = sm + (threadIdx.x & 31);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UNR; u++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (MODE == FFMA_IMM) asm volatile("fma.rn.f32 %0, %0, 0f3F7FF972, %1;" : "+f"(a[i]) : "f"(b[i]));
                if (MODE == FFMA_RRR) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) % CH]));
                if (MODE == FADD_RR) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
                if (MODE == FMUL_RR) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
                if (MODE == FFMA2_IMM) {
                    u64 c; asm("mov.b64 %0, {0f3F7FF972, 0f3F7FF972};" : "=l"(c));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(c), "l"(B[i]));
                }
                if (MODE == FFMA2_RRR) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(B[i]), "l"(B[(i + 1) % CH]));
                if (MODE == FADD2_RR) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B[i]));
                if (MODE == FMUL2_RR) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B[i]));
                if (MODE == FFMA2_ONE) {
                    u64 c = pk(one, one);  // runtime 1.0 in a uniform register: ptxas cannot fold it into FADD2
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(B[i]), "l"(c));
                }
                if (MODE == FFMA2_SWZ) {
                    float x, y; upk(A[i], x, y);
                    u64 sw = pk(y, -x), c; asm("mov.b64 %0, {0f3F7FF972, 0f3F7FF972};" : "=l"(c));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(A[i]) : "l"(sw), "l"(c), "l"(B[i]));
                }
                if (MODE == FADD2_SWZ) {
                    float x, y; upk(A[i], x, y);
                    u64 sw = pk(y, -x);
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(A[i]) : "l"(B[i]), "l"(sw));
                }
                if (MODE == FFMA2_BCAST) {
                    u64 c = pk(b[i], b[i]);
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(B[i]), "l"(c));
                }
                if (MODE == MIX_FFMA2_FADD) {
                    u64 c; asm("mov.b64 %0, {0f3F7FF972, 0f3F7FF972};" : "=l"(c));
                    if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(c), "l"(B[i]));
                    else asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
                }
                if (MODE == MIX_FFMA2_FADD2) {
                    u64 c; asm("mov.b64 %0, {0f3F7FF972, 0f3F7FF972};" : "=l"(c));
                    if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(c), "l"(B[i]));
                    else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B[i]));
                }
                if (MODE == MIX_FFMA2_LDS) {
                    u64 c; asm("mov.b64 %0, {0f3F7FF972, 0f3F7FF972};" : "=l"(c));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(c), "l"(B[i]));
                    if ((i & 3) == 3) { u64 l; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(l) : "l"(sp + 32 * (i + u))); B[i] ^= l & 1; }
                }
            }
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < CH; i++) {
        float x, y; upk(A[i], x, y);
        acc += a[i] + x + y;
        upk(B[i], x, y);
        acc += x + y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(int threads) {
    int dev_sms = 0;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * dev_sms * 1024);
    cudaMalloc(&cyc, sizeof(long long) * dev_sms);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) issue_kernel<MODE><<<dev_sms, threads>>>(out, 1.0001f, 0.9999f, 1.0f, cyc);
    cudaEventRecord(e0);
    issue_kernel<MODE><<<dev_sms, threads>>>(out, 1.0001f, 0.9999f, 1.0f, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * dev_sms, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < dev_sms; i++) c += h[i]; c /= dev_sms;
    double winstr = double(ITERS) * UNR * CH * (threads / 32);
    printf("{\"variant\": \"%s\", \"warps_per_sm\": %d, \"warp_instr_per_cycle_per_sm\": %.3f, "
           "\"flop_per_cycle_per_sm\": %.1f, \"cycles\": %.0f, \"ms\": %.4f, \"implied_mhz\": %.0f, \"err\": \"%s\"}\n",
           kNames[MODE], threads / 32, winstr / c, winstr / c * 32 * kFlop[MODE], c, ms, c / ms / 1e3,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

template <int M>
void run_all(int threads) {
    run<M>(threads);
    if constexpr (M + 1 < NMODES) run_all<M + 1>(threads);
}

int main() {
    for (int threads : {512, 1024}) run_all<0>(threads);
    return 0;
}
