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
  // grow-only scratch shared by every ESACF parameter set of this handle (one call at a time per
  // handle: include/chordb200.h, Conventions)
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // harmonic-energy finalisation scratch (he.cu): [0..11] grid-wide fp64 accumulators, [12] a CTA
  // ticket counter; zero between launches (the last CTA of a launch cleans up after itself)
  double* he_scratch = nullptr;
  // one-shot all-reduce over peer (NVLink) memory, fused into the last CTA of a kernel (comm.cu)
  struct Comm* comm = nullptr;
  int opt_esacf_fit_warps = 0;  // cdb_set_option("esacf_fit_warps"): 0 = the built-in default
  // optional per-kernel timing (cdb_profile_enable / cdb_profile_report): one CUDA event after
  // every launch on the caller's stream; off by default (no events, no cost)
  bool prof_on = false;
  std::vector<std::pair<const char*, cudaEvent_t>> prof_marks;
  std::vector<cudaEvent_t> prof_pool;
};

// marks the end of stage `name` (and the start of the next one) on stream st when profiling is on
void cdb_mark(cdb_handle* h, cudaStream_t st, const char* name);

#define CDB_MAX_PEERS 16
#define CDB_MAIL_STRIDE 16  // doubles per (parity, source rank) mailbox slot: 12 values + flag + pad

// Mailboxes of a one-shot all-reduce of 12 doubles (comm.cu).  Every rank owns
// mail[2][world][CDB_MAIL_STRIDE] in cudaMalloc'ed memory that all peers map through CUDA IPC.
struct Comm {
  int rank = 0, world = 1;
  double* mail[CDB_MAX_PEERS] = {};  // [q] = rank q's mailbox as mapped into this process
  void* local = nullptr;             // this rank's own allocation (== mail[rank])
  unsigned long long seq = 0;        // collective sequence number (host side, lock-step on all ranks)
  int* d_status = nullptr;           // device word: != 0 after a peer timed out
};

// kernel-side view, passed by value
struct CommArgs {
  int rank, world;
  unsigned long long seq;  // 0: no all-reduce in this launch
  double* mail[CDB_MAX_PEERS];
  int* status;
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

// Plan-cache key built from the individual fields (never from raw struct bytes: padding in a C
// caller's struct is uninitialised and would miss the cache on every call, leaking tables).
template <typename... Ts>
static inline std::string cdb_key(const Ts&... v) {
  std::string s;
  (s.append(reinterpret_cast<const char*>(&v), sizeof(v)), ...);
  return s;
}

// Uploads a host table into a new device allocation owned by the handle.  cudaMemcpy from pageable
// memory may return before the DMA has landed and is not ordered against the caller's non-blocking
// stream, so the device is synchronised before the table can be used (plan creation only).
template <typename T>
int cdb_upload(cdb_handle* h, const std::vector<T>& v, T** out) {
  *out = nullptr;
  if (v.empty()) return 0;
  T* d = nullptr;
  CDB_CUDA(h, cudaMalloc(&d, v.size() * sizeof(T)));
  h->owned.push_back(d);
  CDB_CUDA(h, cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  CDB_CUDA(h, cudaDeviceSynchronize());
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

// ---- one-shot all-reduce of 12 doubles through peer mailboxes, executed by ONE warp (the last CTA
// of a launch).  v = this rank's value for lane < 12.  Every rank stores its 12 values and then a
// flag (the sequence number) into slot [seq & 1][rank] of every peer's mailbox over NVLink, waits
// until the world's flags have arrived in its own mailbox, and sums the slots in rank order (so all
// ranks obtain bit-identical sums).  Two slot parities suffice: a rank can only start collective
// s + 2 after every rank has finished s + 1, i.e. has consumed s.  A peer that does not show up
// within ~10 s sets *status and the result becomes NaN (never a silent hang).
__device__ __forceinline__ double comm_allreduce12(const CommArgs& c, double v, int lane) {
  const int par = (int)(c.seq & 1ull);
  const size_t slot = ((size_t)par * c.world + c.rank) * CDB_MAIL_STRIDE;
  if (lane < 12) {
    for (int q = 0; q < c.world; ++q)
      asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(c.mail[q] + slot + lane), "d"(v)
                   : "memory");
    __threadfence_system();
  }
  __syncwarp();
  if (lane < c.world)
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(c.mail[lane] + slot + 12), "l"(c.seq)
                 : "memory");
  bool ok = true;
  // after one time-out the communicator is dead: later collectives do not wait another 10 s each
  const bool dead = c.status && *reinterpret_cast<volatile int*>(c.status) != 0;
  if (dead) ok = false;
  if (!dead && lane < c.world) {
    const double* flag = c.mail[c.rank] + ((size_t)par * c.world + lane) * CDB_MAIL_STRIDE + 12;
    unsigned long long got = 0, t0 = 0, t1 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(flag) : "memory");
      if (got == c.seq) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ull) {
        ok = false;
        break;
      }
      __nanosleep(64);
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  double sum = 0.0;
  if (lane < 12) {
    for (int q = 0; q < c.world; ++q) {
      double x;
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];"
                   : "=d"(x)
                   : "l"(c.mail[c.rank] + ((size_t)par * c.world + q) * CDB_MAIL_STRIDE + lane)
                   : "memory");
      sum += x;
    }
  }
  if (!ok) {
    if (lane == 0 && c.status) atomicExch(c.status, 1);
    sum = __longlong_as_double(0x7ff8000000000000ll);
  }
  return sum;
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
