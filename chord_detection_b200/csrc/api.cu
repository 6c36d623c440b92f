// libchordb200: handle management and shared C-ABI glue (include/chordb200.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>

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

void cdb_mark(cdb_handle* h, cudaStream_t st, const char* name) {
  if (!h || !h->prof_on) return;
  cudaEvent_t e = nullptr;
  if (!h->prof_pool.empty()) {
    e = h->prof_pool.back();
    h->prof_pool.pop_back();
  } else if (cudaEventCreate(&e) != cudaSuccess) {
    return;
  }
  cudaEventRecord(e, st);
  h->prof_marks.emplace_back(name, e);
}

extern "C" {

int cdb_profile_enable(cdb_handle* h, int on) {
  if (!h) return CDB_E_NULL;
  for (auto& m : h->prof_marks) h->prof_pool.push_back(m.second);
  h->prof_marks.clear();
  h->prof_on = on != 0;
  return 0;
}

// "name ms\n" per distinct stage, summed over everything recorded since cdb_profile_enable(h, 1);
// a stage's time is the gap between its mark and the previous mark on the stream ("begin" marks
// open a call and are not reported).  Waits for the recorded events.
int64_t cdb_profile_report(cdb_handle* h, char* buf, int64_t buf_len) {
  if (!h || !buf || buf_len < 1) return CDB_E_NULL;
  std::vector<std::pair<std::string, double>> acc;
  for (size_t i = 1; i < h->prof_marks.size(); ++i) {
    const std::string name = h->prof_marks[i].first;
    if (name == "begin") continue;
    if (cudaEventSynchronize(h->prof_marks[i].second) != cudaSuccess) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->prof_marks[i - 1].second, h->prof_marks[i].second) !=
        cudaSuccess)
      continue;
    bool found = false;
    for (auto& kv : acc)
      if (kv.first == name) {
        kv.second += ms;
        found = true;
      }
    if (!found) acc.emplace_back(name, (double)ms);
  }
  std::string out;
  char line[160];
  for (auto& kv : acc) {
    snprintf(line, sizeof(line), "%s %.6f\n", kv.first.c_str(), kv.second);
    out += line;
  }
  if ((int64_t)out.size() + 1 > buf_len) return CDB_E_INVALID;
  memcpy(buf, out.c_str(), out.size() + 1);
  return (int64_t)out.size();
}

int cdb_version(void) { return CDB_VERSION; }

int cdb_set_option(cdb_handle* h, const char* name, int value) {
  if (!h || !name) return CDB_E_NULL;
  if (std::strcmp(name, "esacf_fit_warps") == 0) {
    if (value < 0 || value > 8) return cdb_fail(h, CDB_E_INVALID, "esacf_fit_warps %d", value);
    h->opt_esacf_fit_warps = value;
    return 0;
  }
  return cdb_fail(h, CDB_E_INVALID, "unknown option %s", name);
}

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
  // finalisation scratch of the harmonic-energy kernel: must be zero before the first launch
  if (cudaMalloc(&h->he_scratch, 16 * sizeof(double)) != cudaSuccess ||
      cudaMemset(h->he_scratch, 0, 16 * sizeof(double)) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    delete h;
    return CDB_E_NOGPU;
  }
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
  cdb_comm_destroy(h);
  if (h->he_scratch) cudaFree(h->he_scratch);
  for (void* p : h->owned) cudaFree(p);
  if (h->ws) cudaFree(h->ws);
  for (auto& m : h->prof_marks) cudaEventDestroy(m.second);
  for (auto e : h->prof_pool) cudaEventDestroy(e);
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
