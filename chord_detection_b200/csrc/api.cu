// libchordb200: handle management and shared C-ABI glue (include/chordb200.h).
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

int cdb_fail(cdb_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

extern "C" {

int cdb_version(void) { return CDB_VERSION; }

int cdb_create(cdb_handle** out, int device) {
  if (!out) return CDB_E_NULL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return CDB_E_NOGPU;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CDB_E_NOGPU;
  if (prop.major != 10) return CDB_E_NOGPU;  // built for sm_100a only; no fallback path
  if (cudaSetDevice(device) != cudaSuccess) return CDB_E_NOGPU;
  cdb_handle* h = new cdb_handle();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->smem_optin = (int)prop.sharedMemPerBlockOptin;
  *out = h;
  return 0;
}

int cdb_destroy(cdb_handle* h) {
  if (!h) return CDB_E_NULL;
  cudaSetDevice(h->device);
  cdb_free_he_plans(h);
  cdb_free_esacf_plans(h);
  cdb_free_iterf0_plans(h);
  cdb_free_prime_plans(h);
  for (void* p : h->owned) cudaFree(p);
  delete h;
  return 0;
}

const char* cdb_last_error(cdb_handle* h) { return h ? h->err.c_str() : "null handle"; }

int64_t cdb_launch_count(cdb_handle* h) { return h ? h->launches : -1; }

int64_t cdb_num_frames(int64_t clip_len, int frame_size, int hop) {
  if (clip_len <= 0 || frame_size <= 0) return 0;
  int64_t hp = hop > 0 ? hop : frame_size;
  return (clip_len + hp - 1) / hp;  // dsp/frame.py:9-10 when hop == frame_size
}

}  // extern "C"
