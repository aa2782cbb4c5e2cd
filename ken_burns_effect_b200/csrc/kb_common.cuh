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
  int W, H;
};

__device__ __forceinline__ bool project(float x, float y, float z, const Camera &cam, Proj &p) {
  if ((double)z < 0.001) return false;                       // :453 (float promoted to double)
  const float nx = __fsub_rn(0.0f, x);                       // dblLineVector = 0 - P      :451
  const float ny = __fsub_rn(0.0f, y);
  const float num = __fsub_rn(cam.f32, z);                   // dot(planePoint - P, n)     :457
  const float den = __fsub_rn(0.0f, z);                      // dot(lineVector, n)         :458
  const float t = __fdiv_rn(num, den);                       //                            :459
  // :461 fabs(den) < 0.001 cannot fire once z >= 0.001 held
  const float ix = __fmaf_rn(t, nx, x);                      // P + t * lineVector         :465
  const float iy = __fmaf_rn(t, ny, y);
  p.ox = __double2float_rn(__dadd_rn(__dadd_rn((double)ix, cam.halfW), -0.5));   // :467
  p.oy = __double2float_rn(__dadd_rn(__dadd_rn((double)iy, cam.halfH), -0.5));   // :468
  p.err = __double2float_rn(__dsub_rn(1000000.0, __ddiv_rn(cam.fB, __dadd_rn((double)z, 0.0000001))));  // :470
  p.nwx = (int)floorf(p.ox);                                 // :472-479
  p.nwy = (int)floorf(p.oy);
  const float x0 = (float)p.nwx, x1 = (float)(p.nwx + 1);
  const float y0 = (float)p.nwy, y1 = (float)(p.nwy + 1);
  const float ax = __fsub_rn(x1, p.ox), bx = __fsub_rn(p.ox, x0);
  const float ay = __fsub_rn(y1, p.oy), by = __fsub_rn(p.oy, y0);
  p.wnw = __fmul_rn(ax, ay);                                 // :481-484
  p.wne = __fmul_rn(bx, ay);
  p.wsw = __fmul_rn(ax, by);
  p.wse = __fmul_rn(bx, by);
  return true;
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
  return (double)err <= __dadd_rn((double)zee, 1.0);
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
