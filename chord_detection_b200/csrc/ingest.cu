// Ingestion helpers: the decode step of librosa.load (multipitch.py:25) for
// PCM16 audio.  soundfile returns int16 PCM as float32 s / 32768 and librosa.to_mono averages the
// channels in float32; both are exact here (the sum of up to 256 16-bit values is exact in
// float32, the division is IEEE), so feeding the device int16 samples halves the host->device
// bytes without changing a single bit of the float32 signal the kernels see.
// Resampling to 22 050 Hz: a polyphase FIR kernel with scipy.signal.resample_poly's semantics
// (the host path of chord_detection_b200/audio.py; librosa's own default resampler, soxr_hq, is an
// unpinned third-party dependency).  The taps are designed on the host
// (scipy.signal.firwin, Kaiser 5.0) and passed in.
#include "common.cuh"

// One output sample of upfirdn(h, x, up, down)[n_pre_remove + m] with h already scaled by `up`
// and shifted by n_pre_pad zeros (resample_poly):  sum_k h[k] x_up[t0 - k],  t0 = (m +
// n_pre_remove) * down - n_pre_pad,  x_up[i * up] = x[i].  FP64 accumulation, one rounding.
__host__ __device__ inline float resample_poly_sample(const float* x, int64_t n_in, int up, int down,
                                                      const float* taps, int n_taps, int n_pre_pad,
                                                      int n_pre_remove, int64_t m) {
  const int64_t t0 = (m + n_pre_remove) * (int64_t)down - n_pre_pad;
  int64_t k = t0 % up;  // first tap whose x_up index is a multiple of up
  if (k < 0) k += up;
  double acc = 0.0;
  for (; k < n_taps; k += up) {
    const int64_t i = (t0 - k) / up;  // exact
    if (i < 0) break;                 // i decreases with k
    if (i < n_in) acc += (double)taps[k] * (double)x[i];
  }
  return (float)acc;
}

__global__ void __launch_bounds__(256) resample_poly_kernel(const float* __restrict__ x, int64_t n_in,
                                                            int up, int down,
                                                            const float* __restrict__ taps, int n_taps,
                                                            int n_pre_pad, int n_pre_remove,
                                                            float* __restrict__ y, int64_t n_out,
                                                            int use_smem) {
  // taps staged in shared memory when they fit, read through L1 / L2 otherwise (very large rate
  // ratios: more than ~58 K taps)
  extern __shared__ float s_taps[];
  const float* t = taps;
  if (use_smem) {
    for (int i = threadIdx.x; i < n_taps; i += blockDim.x) s_taps[i] = taps[i];
    __syncthreads();
    t = s_taps;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < n_out; m += stride)
    y[m] = resample_poly_sample(x, n_in, up, down, t, n_taps, n_pre_pad, n_pre_remove, m);
}

__global__ void __launch_bounds__(256) pcm16_to_mono_kernel(const int16_t* __restrict__ in,
                                                            int64_t n_frames, int channels,
                                                            float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (channels == 1) {
    // 8 samples (one 128-bit load, two 128-bit stores) per thread and step
    const int64_t n8 = n_frames >> 3;
    const bool vec = ((reinterpret_cast<uintptr_t>(in) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec) {
      for (int64_t i = i0; i < n8; i += stride) {
        const int4 q = __ldg(reinterpret_cast<const int4*>(in) + i);
        const int w[4] = {q.x, q.y, q.z, q.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[2 * j] = (float)(short)(w[j] & 0xffff) * (1.0f / 32768.0f);
          f[2 * j + 1] = (float)(short)(w[j] >> 16) * (1.0f / 32768.0f);
        }
        float4* o = reinterpret_cast<float4*>(out) + 2 * i;
        o[0] = make_float4(f[0], f[1], f[2], f[3]);
        o[1] = make_float4(f[4], f[5], f[6], f[7]);
      }
      for (int64_t i = (n8 << 3) + i0; i < n_frames; i += stride)
        out[i] = (float)in[i] * (1.0f / 32768.0f);
    } else {
      for (int64_t i = i0; i < n_frames; i += stride) out[i] = (float)in[i] * (1.0f / 32768.0f);
    }
    return;
  }
  const float inv = (float)channels;
  for (int64_t i = i0; i < n_frames; i += stride) {
    const int16_t* p = in + i * channels;
    float s = 0.0f;
    for (int c = 0; c < channels; ++c) s += (float)p[c] * (1.0f / 32768.0f);
    out[i] = __fdiv_rn(s, inv);  // numpy.mean: float32 sum, then one division
  }
}

extern "C" int cdb_pcm16_to_mono_f32(cdb_handle* h, const int16_t* d_pcm, int64_t n_frames,
                                     int channels, float* d_out, void* stream) {
  if (!h) return CDB_E_NULL;
  if (n_frames < 0 || channels < 1 || channels > 256)
    return cdb_fail(h, CDB_E_INVALID, "n_frames %lld, channels %d", (long long)n_frames, channels);
  if (n_frames == 0) return 0;
  if (!d_pcm || !d_out) return cdb_fail(h, CDB_E_NULL, "null input / output");
  CDB_CUDA(h, cudaSetDevice(h->device));
  const int64_t work = channels == 1 ? (n_frames + 7) / 8 : n_frames;
  const int64_t grid = std::min<int64_t>((work + 255) / 256, (int64_t)h->num_sms * 8);
  pcm16_to_mono_kernel<<<(unsigned)std::max<int64_t>(grid, 1), 256, 0, (cudaStream_t)stream>>>(
      d_pcm, n_frames, channels, d_out);
  h->launches += 1;
  CDB_CUDA(h, cudaGetLastError());
  return 0;
}

// resample_poly on the device.  taps[n_taps] = firwin(2*10*max(up,down)+1, 1/max(up,down),
// window=('kaiser', 5.0)) * up as float32; n_pre_pad / n_pre_remove / n_out as scipy computes them
// (chord_detection_b200/audio.py: resample_plan).
extern "C" int cdb_resample_poly_f32(cdb_handle* h, const float* d_x, int64_t n_in, int up, int down,
                                     const float* d_taps, int n_taps, int n_pre_pad,
                                     int n_pre_remove, float* d_y, int64_t n_out, void* stream) {
  if (!h) return CDB_E_NULL;
  if (n_in < 0 || n_out < 0 || up < 1 || down < 1 || n_taps < 1 || n_pre_pad < 0 ||
      n_pre_remove < 0)
    return cdb_fail(h, CDB_E_INVALID, "invalid resampling plan (up %d, down %d, taps %d)", up, down,
                    n_taps);
  if (n_out == 0) return 0;
  if (!d_x || !d_taps || !d_y) return cdb_fail(h, CDB_E_NULL, "null input / taps / output");
  CDB_CUDA(h, cudaSetDevice(h->device));
  const int64_t grid = std::min<int64_t>((n_out + 255) / 256, (int64_t)h->num_sms * 16);
  // 20*max(up,down)+1 taps: 12 801 for 32 kHz or 96 kHz -> 22.05 kHz, i.e. more than the default
  // 48 KB of dynamic shared memory: opt in up to the device limit, beyond that read from L2
  size_t smem = (size_t)n_taps * sizeof(float);
  const int use_smem = smem <= (size_t)h->smem_optin ? 1 : 0;
  if (!use_smem) smem = 0;
  if (smem > 48 * 1024)
    CDB_CUDA(h, cudaFuncSetAttribute(resample_poly_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  resample_poly_kernel<<<(unsigned)std::max<int64_t>(grid, 1), 256, smem, (cudaStream_t)stream>>>(
      d_x, n_in, up, down, d_taps, n_taps, n_pre_pad, n_pre_remove, d_y, n_out, use_smem);
  h->launches += 1;
  CDB_CUDA(h, cudaGetLastError());
  return 0;
}

// host execution of the same per-sample function (CPU tests, no GPU)
extern "C" int cdb_host_resample_poly_f32(const float* x, int64_t n_in, int up, int down,
                                          const float* taps, int n_taps, int n_pre_pad,
                                          int n_pre_remove, float* y, int64_t n_out) {
  if (!x || !taps || !y || up < 1 || down < 1 || n_taps < 1) return -1;
  for (int64_t m = 0; m < n_out; ++m)
    y[m] = resample_poly_sample(x, n_in, up, down, taps, n_taps, n_pre_pad, n_pre_remove, m);
  return 0;
}
