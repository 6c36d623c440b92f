// Packed FP32x2 arithmetic for sm_100a (PTX fma/add/sub/mul.rn.f32x2 -> SASS FFMA2 / FADD2 / FMUL2).
// A complex number lives in ONE aligned 64-bit register pair: re in the low half, im in the high
// half.  ptxas folds the pack / unpack / swap / negate helpers below into operand modifiers of
// the packed instruction (.LO_HI swizzle, per-half sign, 32-bit scalar or immediate broadcast),
// so a twiddled radix-2 butterfly is 3 instructions and a complex multiply is 2.
#pragma once
#include <stdint.h>

typedef unsigned long long c64;  // packed (lo = re, hi = im)

__device__ __forceinline__ c64 pk(float lo, float hi) {
  c64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(c64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ c64 bc(float s) { return pk(s, s); }
__device__ __forceinline__ c64 fma2(c64 a, c64 b, c64 c) {
  c64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ c64 add2(c64 a, c64 b) {
  c64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c64 sub2(c64 a, c64 b) {
  c64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c64 mul2(c64 a, c64 b) {
  c64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c64 neg2(c64 a) {
  float x, y;
  upk(a, x, y);
  return pk(-x, -y);
}
__device__ __forceinline__ c64 conj2(c64 a) {  // (re, -im)
  float x, y;
  upk(a, x, y);
  return pk(x, -y);
}
__device__ __forceinline__ c64 swp(c64 a) {  // (im, re)
  float x, y;
  upk(a, x, y);
  return pk(y, x);
}
__device__ __forceinline__ c64 mul_mi(c64 a) {  // -i * a = (im, -re)
  float x, y;
  upk(a, x, y);
  return pk(y, -x);
}
__device__ __forceinline__ c64 mul_pi(c64 a) {  // +i * a = (-im, re)
  float x, y;
  upk(a, x, y);
  return pk(-y, x);
}
// z * w for packed z and w = (wr, wi): (zr wr - zi wi, zi wr + zr wi)
__device__ __forceinline__ c64 cmul2(c64 z, c64 w) {
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
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
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
