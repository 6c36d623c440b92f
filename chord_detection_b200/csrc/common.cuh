// Shared host/device helpers for libchordb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/chordb200.h"

struct HePlan;
struct EsacfPlan;
struct IterF0Plan;
struct PrimePlan;

struct cdb_handle {
  int device = 0;
  int num_sms = 0;
  int smem_optin = 0;
  std::string err;
  int64_t launches = 0;
  std::map<std::string, HePlan*> he_plans;
  std::map<std::string, EsacfPlan*> esacf_plans;
  std::map<std::string, IterF0Plan*> iterf0_plans;
  std::map<std::string, PrimePlan*> prime_plans;
  std::vector<void*> owned;  // device allocations freed by cdb_destroy
};

// plan destructors live with their kernels
void cdb_free_he_plans(cdb_handle* h);
void cdb_free_esacf_plans(cdb_handle* h);
void cdb_free_iterf0_plans(cdb_handle* h);
void cdb_free_prime_plans(cdb_handle* h);

int cdb_fail(cdb_handle* h, int code, const char* fmt, ...);

#define CDB_CUDA(h, call)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return cdb_fail((h), (int)e__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                      __FILE__, __LINE__);                                             \
  } while (0)

template <typename T>
static inline std::string pod_key(const T& v) {
  return std::string(reinterpret_cast<const char*>(&v), sizeof(T));
}

template <typename T>
int cdb_upload(cdb_handle* h, const std::vector<T>& v, T** out) {
  *out = nullptr;
  if (v.empty()) return 0;
  T* d = nullptr;
  CDB_CUDA(h, cudaMalloc(&d, v.size() * sizeof(T)));
  h->owned.push_back(d);
  CDB_CUDA(h, cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}

// ---------------------------------------------------------------- device side
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP / SYNCS)
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif  // __CUDACC__
