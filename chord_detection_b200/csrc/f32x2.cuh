// Packed FP32x2 arithmetic for sm_100a (PTX fma/add/sub/mul.rn.f32x2 -> SASS FFMA2 / FADD2 / FMUL2).
// A complex number lives in ONE aligned 64-bit register pair: re in the low half, im in the high
// half.  ptxas folds the pack / unpack / swap / negate helpers below into operand modifiers of
// the packed instruction (.LO_HI swizzle, per-half sign, 32-bit scalar or immediate broadcast),
// so a twiddled radix-2 butterfly is 3 instructions and a complex multiply is 2.
// The arithmetic helpers are __host__ __device__: the host bodies emulate the packed instructions
// with scalar fmaf / + / - / * so that kernels written on top of them can be executed thread by
// thread in CPU tests (same rounding: every packed lane is an IEEE single operation).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef unsigned long long c64;  // packed (lo = re, hi = im)

#define F32X2_HD __host__ __device__ __forceinline__

F32X2_HD c64 pk(float lo, float hi) {
  c64 r;
#ifdef __CUDA_ARCH__
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
#else
  uint32_t a, b;
  memcpy(&a, &lo, 4);
  memcpy(&b, &hi, 4);
  r = (c64)a | ((c64)b << 32);
#endif
  return r;
}
F32X2_HD void upk(c64 v, float& lo, float& hi) {
#ifdef __CUDA_ARCH__
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
#else
  const uint32_t a = (uint32_t)v, b = (uint32_t)(v >> 32);
  memcpy(&lo, &a, 4);
  memcpy(&hi, &b, 4);
#endif
}
F32X2_HD c64 bc(float s) { return pk(s, s); }
F32X2_HD c64 fma2(c64 a, c64 b, c64 c) {
#ifdef __CUDA_ARCH__
  c64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
#else
  float ax, ay, bx, by, cx, cy;
  upk(a, ax, ay);
  upk(b, bx, by);
  upk(c, cx, cy);
  return pk(fmaf(ax, bx, cx), fmaf(ay, by, cy));
#endif
}
F32X2_HD c64 add2(c64 a, c64 b) {
#ifdef __CUDA_ARCH__
  c64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
#else
  float ax, ay, bx, by;
  upk(a, ax, ay);
  upk(b, bx, by);
  return pk(ax + bx, ay + by);
#endif
}
F32X2_HD c64 sub2(c64 a, c64 b) {
#ifdef __CUDA_ARCH__
  c64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
#else
  float ax, ay, bx, by;
  upk(a, ax, ay);
  upk(b, bx, by);
  return pk(ax - bx, ay - by);
#endif
}
F32X2_HD c64 mul2(c64 a, c64 b) {
#ifdef __CUDA_ARCH__
  c64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
#else
  float ax, ay, bx, by;
  upk(a, ax, ay);
  upk(b, bx, by);
  return pk(ax * bx, ay * by);
#endif
}
F32X2_HD c64 neg2(c64 a) {
  float x, y;
  upk(a, x, y);
  return pk(-x, -y);
}
F32X2_HD c64 conj2(c64 a) {  // (re, -im)
  float x, y;
  upk(a, x, y);
  return pk(x, -y);
}
F32X2_HD c64 swp(c64 a) {  // (im, re)
  float x, y;
  upk(a, x, y);
  return pk(y, x);
}
F32X2_HD c64 mul_mi(c64 a) {  // -i * a = (im, -re)
  float x, y;
  upk(a, x, y);
  return pk(y, -x);
}
F32X2_HD c64 mul_pi(c64 a) {  // +i * a = (-im, re)
  float x, y;
  upk(a, x, y);
  return pk(-y, x);
}
// z * w for packed z and w = (wr, wi): (zr wr - zi wi, zi wr + zr wi)
F32X2_HD c64 cmul2(c64 z, c64 w) {
  float wr, wi;
  upk(w, wr, wi);
  return fma2(mul_pi(z), bc(wi), mul2(z, bc(wr)));
}

// 64-bit shared-memory store of a packed value from its two halves: written this way ptxas stores
// straight from the FFMA2 result registers (a plain 64-bit store of the asm result costs two MOVs).
__device__ __forceinline__ void sts2(void* smem_ptr, c64 v) {
  float x, y;
  upk(v, x, y);
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr))),
               "f"(x), "f"(y)
               : "memory");
}

// MUFU.SQRT (relative error ~2^-22): the 4th root of the window maxima needs no IEEE rounding
F32X2_HD float sqrt_approx(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}

// 64-bit shared-memory load that keeps its program order relative to the other volatile
// shared-memory accesses (used to software-pipeline table loads ahead of their use).
__device__ __forceinline__ c64 lds2v(const void* smem_ptr) {
  c64 r;
  asm volatile("ld.shared.b64 %0, [%1];"
               : "=l"(r)
               : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr)))
               : "memory");
  return r;
}
