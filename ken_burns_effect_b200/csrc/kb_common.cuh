// kb_common.cuh -- shared helpers for libkb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "kb200.h"

#ifndef __CUDA_ARCH__
#define KB_HOST_ONLY 1
#endif
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libkb200 is written for sm_100a (B200) only"
#endif

namespace kb {

// ---- error plumbing ------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int check_launch(const char *what);
void count_launch(int n = 1);

#define KB_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      kb::set_error(__VA_ARGS__);      \
      return KB_EINVAL;                \
    }                                  \
  } while (0)

static inline unsigned int cdiv(long a, long b) { return (unsigned int)((a + b - 1) / b); }

// ---- splat geometry, utils/common.py:447-484 (identical in all three point kernels) ----------------
// The rounding of every operation is spelled out with intrinsics so that nvcc cannot re-associate or
// contract differently from the reference's NVRTC build (whose SASS shows: FADD for num/den, a full
// precision fp32 division, one FFMA per intersection component, DADD/DADD + F2F for ox/oy, an IEEE fp64
// division + DADD + F2F for err).
//
// The reference evaluates several sub-expressions in fp64 (double literals pasted into float code).  ncu
// showed the float<->double / float<->int conversions of a literal transcription saturating the XU pipe
// (16 lanes/clk/SM; profiles/ncu_r01b_summary.md), so every such expression is replaced here by an fp32
// form that is PROVABLY the same function (tests/test_exact_arith.py brute-forces each identity on the CPU):
//   (double)z < 0.001                    ==  z < 0.001f          ((float)0.001 is the smallest float >= 0.001)
//   (float)((double)i + 0.5*W - 0.5)     ==  i + (float)(0.5*W - 0.5)   for W >= 2: both double additions are
//                                            exact or far below half an ulp of the result, so one fp32
//                                            rounding of the exact sum is what the reference computes too
//   (int)floor(o), (float)(int)          ==  magic-number rounding + fix-up, for |o| < 2^22 (farther out a
//                                            point touches no pixel of any image this library accepts)
//   (double)a >= (double)b + 1.0         ==  sign of the exact difference a - b - 1 (TwoSum), for 1 <= |b| <= 1e15
// Only err keeps its fp64 division (its quotient is rounded twice; no fp32 shortcut reproduces that).
struct Proj {
  float ox, oy, err;
  int nwx, nwy;
  float wnw, wne, wsw, wse;
};

struct Camera {
  float f32;      // (float)focal            -- make_float3(0, 0, focal)
  double fB;      // focal * baseline        -- the double constant the reference's compiler folds
  double halfW;   // 0.5 * W
  double halfH;   // 0.5 * H
  float cx, cy;   // (float)(0.5*W - 0.5), (float)(0.5*H - 0.5): exact for W, H < 2^24
  int W, H;
};

constexpr float kFloorRange = 4194304.0f;   // 2^22; images are limited to 2^22 - 2 pixels per side

// floor(v) as float and int without conversion instructions; requires |v| < 2^22.
__device__ __forceinline__ void floor_fi(float v, float &f, int &i) {
  const float magic = 12582912.0f;                         // 1.5 * 2^23: ulp 1 around it
  const float r = __fadd_rn(v, magic);                     // magic + rint(v)
  i = __float_as_int(r) - 0x4B400000;
  f = __fsub_rn(r, magic);
  if (f > v) {
    f = __fsub_rn(f, 1.0f);
    i -= 1;
  }
}

// float -> double for a normal, finite float: pure integer work (the F2F instruction runs on the XU pipe).
__device__ __forceinline__ double widen_normal(float v) {
  const unsigned b = __float_as_uint(v);
  const unsigned hi = (b & 0x80000000u) | (((b & 0x7FFFFFFFu) >> 3) + 0x38000000u);
  return __hiloint2double((int)hi, (int)(b << 29));
}

__device__ __forceinline__ bool project(float x, float y, float z, const Camera &cam, Proj &p) {
  if (!(z >= 0.001f)) return false;                          // :453  (double)z < 0.001; NaN is culled like there
  const float nx = __fsub_rn(0.0f, x);                       // dblLineVector = 0 - P      :451
  const float ny = __fsub_rn(0.0f, y);
  const float num = __fsub_rn(cam.f32, z);                   // dot(planePoint - P, n)     :457
  const float den = __fsub_rn(0.0f, z);                      // dot(lineVector, n)         :458
  const float t = __fdiv_rn(num, den);                       //                            :459
  // :461 fabs(den) < 0.001 cannot fire once z >= 0.001 held
  const float ix = __fmaf_rn(t, nx, x);                      // P + t * lineVector         :465
  const float iy = __fmaf_rn(t, ny, y);
  if (cam.W >= 2 && cam.H >= 2) {
    p.ox = __fadd_rn(ix, cam.cx);                            // :467-468, see the identities above
    p.oy = __fadd_rn(iy, cam.cy);
  } else {                                                   // 1-pixel-wide images: the literal form
    p.ox = __double2float_rn(__dadd_rn(__dadd_rn((double)ix, cam.halfW), -0.5));
    p.oy = __double2float_rn(__dadd_rn(__dadd_rn((double)iy, cam.halfH), -0.5));
  }
  if (!(fabsf(p.ox) < kFloorRange && fabsf(p.oy) < kFloorRange)) return false;   // touches no pixel (also NaN)
  const double zd = z < 3.0e38f ? widen_normal(z) : (double)z;
  p.err = __double2float_rn(__dsub_rn(1000000.0, __ddiv_rn(cam.fB, __dadd_rn(zd, 0.0000001))));  // :470
  float x0, y0;
  floor_fi(p.ox, x0, p.nwx);                                 // :472-479
  floor_fi(p.oy, y0, p.nwy);
  const float x1 = __fadd_rn(x0, 1.0f), y1 = __fadd_rn(y0, 1.0f);   // (float)(nwx + 1): exact below 2^24
  const float ax = __fsub_rn(x1, p.ox), bx = __fsub_rn(p.ox, x0);
  const float ay = __fsub_rn(y1, p.oy), by = __fsub_rn(p.oy, y0);
  p.wnw = __fmul_rn(ax, ay);                                 // :481-484
  p.wne = __fmul_rn(bx, ay);
  p.wsw = __fmul_rn(ax, by);
  p.wse = __fmul_rn(bx, by);
  return true;
}

// Exact sign tests of the reference's fp64 comparisons "(double)a >= (double)b + 1.0" (:556-561) and
// "(double)a <= (double)b + 1.0" (:639).  d = fl(a - b) decides unless it is exactly 1; then TwoSum's
// error term e (a - b = d + e exactly) does.  b + 1.0 is exact in fp64 for 1 <= |b| <= 2^52, which is what
// makes the exact difference the reference's own result; other magnitudes take the literal form.
__device__ __forceinline__ float twosum_err(float a, float b, float d) {   // d = fl(a - b)
  const float bb = __fsub_rn(d, a);
  return __fadd_rn(__fsub_rn(a, __fsub_rn(d, bb)), __fsub_rn(__fsub_rn(0.0f, b), bb));
}
__device__ __forceinline__ bool cmp_exact_domain(float a, float b) {
  return (fabsf(b) >= 1.0f) & (fabsf(b) <= 1.0e15f) & (fabsf(a) <= 1.0e15f);
}
__device__ __forceinline__ bool ge_plus_one(float a, float b) {            // (double)a >= (double)b + 1.0
  if (!cmp_exact_domain(a, b)) return (double)a >= __dadd_rn((double)b, 1.0);
  const float d = __fsub_rn(a, b);
  if (d != 1.0f) return d > 1.0f;
  return twosum_err(a, b, d) >= 0.0f;
}
__device__ __forceinline__ bool le_plus_one(float a, float b) {            // (double)a <= (double)b + 1.0
  if (!cmp_exact_domain(a, b)) return (double)a <= __dadd_rn((double)b, 1.0);
  const float d = __fsub_rn(a, b);
  if (d != 1.0f) return d < 1.0f;
  return twosum_err(a, b, d) <= 0.0f;
}

// process_shift's tensor half (utils/common.py:104-109) applied to one point: the reference runs
// x *= z / (z + 1e-7) and x += shift as separate fp32 torch kernels, so multiply and add stay unfused.
__device__ __forceinline__ void shift_point(float &x, float &y, float &z, float sx, float sy, float sz) {
  const float r = __fdiv_rn(z, __fadd_rn(z, 0.0000001f));
  x = __fadd_rn(__fmul_rn(x, r), sx);
  y = __fadd_rn(__fmul_rn(y, r), sy);
  z = __fadd_rn(z, sz);
}

// Which of the four neighbours updateZee votes for (0 NW, 1 NE, 2 SW, 3 SE, -1 none), :486-506.
__device__ __forceinline__ int pick_neighbour(const Proj &p) {
  const float a = p.wnw, b = p.wne, c = p.wsw, d = p.wse;
  if ((a >= b) & (a >= c) & (a >= d)) return 0;
  if ((b >= a) & (b >= c) & (b >= d)) return 1;
  if ((c >= a) & (c >= b) & (c >= d)) return 2;
  if ((d >= a) & (d >= b) & (d >= c)) return 3;
  return -1;
}

// float min on a z-buffer cell.  Non-negative floats order like their int bit patterns, so the common
// case is one native RED.MIN.S32; negative values (depth < focal*baseline/1e6) take the reference's CAS
// loop (utils/common.py:275-283).  Both keep "cell = min of everything written so far".
__device__ __forceinline__ void zmin(float *cell, float v) {
  const int iv = __float_as_int(v);
  if (iv >= 0) {
    atomicMin(reinterpret_cast<int *>(cell), iv);
  } else {
    int old = *reinterpret_cast<volatile int *>(cell);
    while (__int_as_float(old) > v) {
      const int seen = atomicCAS(reinterpret_cast<int *>(cell), old, iv);
      if (seen == old) break;
      old = seen;
    }
  }
}

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// The z gate of updateOutput, :639: (double)err <= (double)zee + 1.0.
__device__ __forceinline__ bool z_gate(float err, float zee) {
  return le_plus_one(err, zee);
}

// The 16 ray directions of fill_disocclusion (utils/common.py:859-867) after the reference's per-thread
// fp32 normalisation d / sqrt(dx*dx + dy*dy); the values are exact fp32 results, written as hex floats.
static __constant__ float c_dirx[16] = {
    -0x1.6a09e6p-1f, 0x0.0p+0f, 0x1.6a09e6p-1f, 0x1.0p+0f, -0x1.c9f25cp-2f, 0x1.c9f25cp-2f, 0x1.c9f25cp-1f, 0x1.c9f25cp-1f,
    -0x1.1c01aap-1f, -0x1.43d136p-2f, 0x1.43d136p-2f, 0x1.1c01aap-1f, 0x1.aa028p-1f, 0x1.e5b9dp-1f, 0x1.e5b9dp-1f, 0x1.aa028p-1f};
static __constant__ float c_diry[16] = {
    0x1.6a09e6p-1f, 0x1.0p+0f, 0x1.6a09e6p-1f, 0x0.0p+0f, 0x1.c9f25cp-1f, 0x1.c9f25cp-1f, 0x1.c9f25cp-2f, -0x1.c9f25cp-2f,
    0x1.aa028p-1f, 0x1.e5b9dp-1f, 0x1.e5b9dp-1f, 0x1.aa028p-1f, 0x1.1c01aap-1f, 0x1.43d136p-2f, -0x1.43d136p-2f, -0x1.1c01aap-1f};

}  // namespace kb
