// libchordb200: batched Chromagram._pack + detect_key (chromagram.py:50-126), SURVEY.md 8f-1.
// One thread per chroma vector; the arithmetic lives in chroma_pack.cuh (host/device).
#include "chroma_pack.cuh"
#include "common.cuh"

__global__ void pack_key_kernel(const double* __restrict__ chroma, int64_t n,
                                uint8_t* __restrict__ digits, int32_t* __restrict__ key) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double c[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) c[j] = chroma[i * 12 + j];
  if (digits) {
    uint8_t d[12];
    cpk::pack_digits(c, d);
#pragma unroll
    for (int j = 0; j < 12; ++j) digits[i * 12 + j] = d[j];
  }
  if (key) key[i] = cpk::key_code(c);
}

extern "C" {

int cdb_pack_and_key(cdb_handle* h, const double* d_chroma, int64_t n, uint8_t* d_digits,
                     int32_t* d_key, void* stream) {
  if (!h) return CDB_E_NULL;
  if (!d_chroma || n < 0) return cdb_fail(h, CDB_E_INVALID, "bad arguments");
  if (n == 0) return 0;
  CDB_CUDA(h, cudaSetDevice(h->device));
  pack_key_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_chroma, n,
                                                                                 d_digits, d_key);
  h->launches += 1;
  CDB_CUDA(h, cudaGetLastError());
  return 0;
}

int cdb_host_pack_and_key(const double* chroma, int64_t n, uint8_t* digits, int32_t* key) {
  if (!chroma || n < 0) return CDB_E_INVALID;
  for (int64_t i = 0; i < n; ++i) {
    if (digits) cpk::pack_digits(chroma + i * 12, digits + i * 12);
    if (key) key[i] = cpk::key_code(chroma + i * 12);
  }
  return 0;
}

double cdb_host_py_round3(double v) { return cpk::py_round3(v); }

}  // extern "C"
