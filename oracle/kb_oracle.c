/*
 * kb_oracle.c -- CPU restatement of the reference's novel-view render path.   TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call this
 * library; the product (ken_burns_effect_b200) never links or loads it.
 *
 * Every function restates one piece of pierlj/ken-burns-effect (paths relative to /root/reference) and
 * keeps the reference's floating-point expression tree: which sub-expressions are fp32, which are
 * promoted to fp64 by the double literals the reference pastes into its CUDA source, and where NVRTC
 * contracts a*b+c into one FMA (checked in the SASS of the reference cubins built by oracle/build_ref.py).
 * Compile with -ffp-contract=off so the C compiler adds no contraction of its own (see Makefile).
 *
 * Parity pin: tests/test_gpu_render.py and tests/test_gpu_mask.py run the reference's own kernels (oracle/_ref/ cubins) on
 * the B200 and compares them with this file bit-for-bit (z-buffer) / to fp32 summation-order noise
 * (accumulators); tests/golden/ holds outputs of those reference kernels for the CPU-only suite.
 *
 * Threading: every kernel is an OpenMP parallel-for over points or pixels (one thread per element like
 * the CUDA grid); results are independent of the thread count except for the fp32 summation order of
 * kbo_splat_accum, which is the same freedom the reference's atomicAdd has.  Tests run it with 1 thread.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KBO_API __attribute__((visibility("default")))

KBO_API int kbo_version(void) { return 1; }

KBO_API void kbo_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

KBO_API int kbo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- float atomics on plain memory (the reference's atomicMin is a CAS loop, common.py:275-283) ---- */
/* Grow-only scratch buffers, one per slot: a frame needs ~75 MB of temporaries, and getting them from malloc() every frame means
 * fresh mmaps and ~18 000 page faults per frame -- which made the SAME loop measure 36 frames/s in a fresh process and 55 in a
 * process whose heap had been stretched before (bench.py's two arms).  Single caller at a time (the kernels inside are OpenMP). */
static void *kbo_scratch(int slot, size_t bytes) {
  static void *buf[16];
  static size_t cap[16];
  if (bytes > cap[slot]) {
    free(buf[slot]);
    buf[slot] = malloc(bytes);
    cap[slot] = buf[slot] ? bytes : 0;
  }
  return buf[slot];
}

static inline void atomic_min_f32(float *addr, float v) {
  int32_t *ia = (int32_t *)addr;
  int32_t old = __atomic_load_n(ia, __ATOMIC_RELAXED);
  for (;;) {
    float fo;
    memcpy(&fo, &old, 4);
    if (!(fo > v)) return;
    int32_t nv;
    memcpy(&nv, &v, 4);
    if (__atomic_compare_exchange_n(ia, &old, nv, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return;
  }
}

static inline void atomic_add_f32(float *addr, float v) {
  int32_t *ia = (int32_t *)addr;
  int32_t old = __atomic_load_n(ia, __ATOMIC_RELAXED);
  for (;;) {
    float fo, fn;
    memcpy(&fo, &old, 4);
    fn = fo + v;
    int32_t nv;
    memcpy(&nv, &fn, 4);
    if (__atomic_compare_exchange_n(ia, &old, nv, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Splat geometry shared by the three point kernels (common.py:447-484, :599-636, :710-751).
 *   z test        : (double)z < 0.001                      (:453)   float vs double literal
 *   num, den      : fp32 dot products with (0,0,1)         (:457-458, helper_math.h:1244); the x,y terms
 *                   are exact zeros for finite input, so num = f32(focal) - z and den = -z
 *   t             : fp32 IEEE division                     (:459)
 *   I = P + t*(-P): one FMA per component                  (:465, helper_math.h:345,814; FFMA in the SASS)
 *   ox, oy        : ((double)I.x + 0.5*W) - 0.5 -> float   (:467-468)
 *   err           : 1e6 - (focal*baseline)/((double)z + 1e-7) -> float   (:470), focal*baseline in double
 *   weights       : fp32 products of fp32 differences      (:481-484)
 * Returns 0 when the point is culled.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  float ox, oy, err;
  int nwx, nwy;
  float wnw, wne, wsw, wse;
} kbo_proj;

static inline int kbo_project(float x, float y, float z, double focal, double baseline, int W, int H,
                              kbo_proj *p) {
  if ((double)z < 0.001) return 0;
  const float f32 = (float)focal;
  const float nx = 0.0f - x, ny = 0.0f - y, nz = 0.0f - z; /* dblLineVector = 0 - P */
  const float num = f32 - z;
  const float den = nz;
  const float t = num / den;
  if (fabs((double)den) < 0.001) return 0;
  const float ix = fmaf(t, nx, x);
  const float iy = fmaf(t, ny, y);
  p->ox = (float)(((double)ix + 0.5 * (double)W) - 0.5);
  p->oy = (float)(((double)iy + 0.5 * (double)H) - 0.5);
  p->err = (float)(1000000.0 - ((focal * baseline) / ((double)z + 0.0000001)));
  p->nwx = (int)floorf(p->ox);
  p->nwy = (int)floorf(p->oy);
  const float sex = (float)(p->nwx + 1), sey = (float)(p->nwy + 1);
  const float nwxf = (float)p->nwx, nwyf = (float)p->nwy;
  p->wnw = (sex - p->ox) * (sey - p->oy);
  p->wne = (p->ox - nwxf) * (sey - p->oy);
  p->wsw = (sex - p->ox) * (p->oy - nwyf);
  p->wse = (p->ox - nwxf) * (p->oy - nwyf);
  return 1;
}

/* process_shift's tensor half (common.py:104-109): clone; x,y *= z/(z+1e-7) in fp32; += shift.
 * Each step is its own torch kernel, so no contraction between the multiply and the add. */
KBO_API void kbo_shift_points(const float *xyz, long N, const float *shift3, float *out) {
#pragma omp parallel for schedule(static)
  for (long n = 0; n < N; ++n) {
    const float x = xyz[n], y = xyz[N + n], z = xyz[2 * N + n];
    const float r = z / (z + 0.0000001f);
    const float xs = x * r, ys = y * r;
    out[n] = xs + shift3[0];
    out[N + n] = ys + shift3[1];
    out[2 * N + n] = z + shift3[2];
  }
}

/* Index of the neighbour updateZee picks (0 NW, 1 NE, 2 SW, 3 SE, -1 none), common.py:486-506. */
static inline int kbo_pick(const kbo_proj *p) {
  const float a = p->wnw, b = p->wne, c = p->wsw, d = p->wse;
  if ((a >= b) & (a >= c) & (a >= d)) return 0;
  if ((b >= a) & (b >= c) & (b >= d)) return 1;
  if ((c >= a) & (c >= b) & (c >= d)) return 2;
  if ((d >= a) & (d >= b) & (d >= c)) return 3;
  return -1; /* NaN weights */
}

/* kernel_pointrender_updateZee (common.py:434-507).  xyz [B,3,N]; zee [B,H,W] is filled with 1e6 here
 * (common.py:430).  pix_idx (optional, [B,N]) receives y*W+x of the chosen pixel or -1: the integer index
 * map of SURVEY.md 8(c). */
KBO_API void kbo_splat_min(const float *xyz, int B, long N, double focal, double baseline, float *zee, int H,
                           int W, int32_t *pix_idx) {
  const long P = (long)H * W;
  for (long i = 0; i < B * P; ++i) zee[i] = 1000000.0f;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)B * N; ++i) {
    const long b = i / N, n = i % N;
    const float *s = xyz + b * 3 * N;
    kbo_proj p;
    if (pix_idx) pix_idx[i] = -1;
    if (!kbo_project(s[n], s[N + n], s[2 * N + n], focal, baseline, W, H, &p)) continue;
    const int k = kbo_pick(&p);
    if (k < 0) continue;
    const int px = p.nwx + (k & 1), py = p.nwy + (k >> 1);
    if ((px >= 0) & (px < W) & (py >= 0) & (py < H)) {
      atomic_min_f32(&zee[b * P + (long)py * W + px], p.err);
      if (pix_idx) pix_idx[i] = py * W + px;
    }
  }
}

/* generate_mask's kernel (common.py:696-827), sequential semantics.  The reference runs one thread per point: a point that
 * finds the cell of the pixel it votes for strictly behind its own err lowers the cell, sets its own mask to 1, swaps its
 * index into id_memory and clears the mask of the point it displaced (unless that was point 0: `pid > 0`, :759); a point
 * that does not lower the cell clears its own mask.  Threads race (check-then-atomicMin), so the reference's result
 * depends on scheduling; this restatement executes the points in index order, one outcome the race allows and the
 * canonical one this repository adopts: mask[n] = 1 iff n is the LOWEST-indexed point among those with the minimal err
 * at its pixel -- except that a displaced point 0 keeps its 1, like in the reference.  mask [B,N] (N = H*W at the call
 * site, :829), zee [B,H,W] scratch. */
KBO_API void kbo_mask_zee(const float *xyz, int B, long N, double focal, double baseline, int H, int W, float *mask,
                          float *zee) {
  const long P = (long)H * W;
  int32_t *ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
  for (long b = 0; b < B; ++b) {
    float *z = zee + b * P, *m = mask + b * N;
    const float *s = xyz + b * 3 * N;
    for (long i = 0; i < P; ++i) { z[i] = 1000000.0f; ids[i] = -1; }
    for (long n = 0; n < N; ++n) m[n] = 0.0f;
    for (long n = 0; n < N; ++n) {
      kbo_proj p;
      if (!kbo_project(s[n], s[N + n], s[2 * N + n], focal, baseline, W, H, &p)) continue;
      const int k = kbo_pick(&p);
      if (k < 0) continue;
      const int px = p.nwx + (k & 1), py = p.nwy + (k >> 1);
      if (!((px >= 0) & (px < W) & (py >= 0) & (py < H))) continue;
      const long pix = (long)py * W + px;
      if (z[pix] > p.err) {
        z[pix] = p.err;
        m[n] = 1.0f;
        const int32_t pid = ids[pix];
        ids[pix] = (int32_t)n;
        if (pid > 0) m[pid] = 0.0f;
      } else {
        m[n] = 0.0f;
      }
    }
  }
  free(ids);
}

/* kernel_pointrender_updateDegrid (common.py:524-568).
 * The reference updates zee IN PLACE while other threads read it (a benign race).  mode 0 restates it
 * race-free: read zee_in, write zee_out (what every thread would see if all reads came first) -- the
 * canonical semantics this repository adopts.  mode 1 is the sequential in-place raster scan,
 * another outcome the race allows; tests use both to bound how many pixels the race can touch. */
KBO_API void kbo_degrid(const float *zee_in, float *zee_out, int B, int H, int W, int mode) {
  static const int ox[4] = {1, 0, 1, 1};
  static const int oy[4] = {0, 1, 1, -1};
  const long P = (long)H * W;
  if (mode == 1) {
    if (zee_out != zee_in) memcpy(zee_out, zee_in, sizeof(float) * B * P);
    zee_in = zee_out;
  }
#pragma omp parallel for schedule(static) if (mode == 0)
  for (long i = 0; i < B * P; ++i) {
    const long b = i / P;
    const int y = (int)((i % P) / W), x = (int)(i % W);
    const float *zi = zee_in + b * P;
    const float c = zi[(long)y * W + x];
    int count = 0;
    float sum = 0.0f;
    for (int k = 0; k < 4; ++k) {
      const int x1 = x + ox[k], y1 = y + oy[k], x2 = x - ox[k], y2 = y - oy[k];
      if ((x1 < 0) | (x1 >= W) | (y1 < 0) | (y1 >= H)) continue;
      if ((x2 < 0) | (x2 >= W) | (y2 < 0) | (y2 >= H)) continue;
      const float a = zi[(long)y1 * W + x1], d = zi[(long)y2 * W + x2];
      if ((double)c >= (double)a + 1.0) {
        if ((double)c >= (double)d + 1.0) {
          count += 2;
          sum += a;
          sum += d;
        }
      }
    }
    float r = c;
    if (count > 0) r = fminf(c, sum / (float)count);
    zee_out[i] = r;
  }
}

/* kernel_pointrender_updateOutput (common.py:585-669).  data [B,C,N]; out [B,C+1,H,W] zero-filled here
 * (common.py:431); the last plane is the ones channel the reference concatenates (common.py:429). */
KBO_API void kbo_splat_accum(const float *xyz, const float *data, int B, long N, int C, double focal,
                             double baseline, const float *zee, float *out, int H, int W) {
  const long P = (long)H * W;
  memset(out, 0, sizeof(float) * B * (C + 1) * P);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)B * N; ++i) {
    const long b = i / N, n = i % N;
    const float *s = xyz + b * 3 * N;
    const float *d = data + b * C * N;
    kbo_proj p;
    if (!kbo_project(s[n], s[N + n], s[2 * N + n], focal, baseline, W, H, &p)) continue;
    const float w4[4] = {p.wnw, p.wne, p.wsw, p.wse};
    for (int k = 0; k < 4; ++k) {
      const int px = p.nwx + (k & 1), py = p.nwy + (k >> 1);
      if (!((px >= 0) & (px < W) & (py >= 0) & (py < H))) continue;
      const long pix = (long)py * W + px;
      if (!((double)p.err <= (double)zee[b * P + pix] + 1.0)) continue;
      float *o = out + b * (C + 1) * P + pix;
      for (int c = 0; c < C; ++c) atomic_add_f32(o + c * P, d[c * N + n] * w4[k]);
      atomic_add_f32(o + C * P, 1.0f * w4[k]);
    }
  }
}

/* common.py:686 -- render = out[:C] / (out[C] + 1e-7), existing = out[C].  (1e-7 is a python float added
 * to a float tensor: fp32 add of (float)1e-7.) */
KBO_API void kbo_normalize(const float *out, int B, int C, int H, int W, float *render, float *existing) {
  const long P = (long)H * W;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < B * P; ++i) {
    const long b = i / P, pix = i % P;
    const float w = out[(b * (C + 1) + C) * P + pix];
    const float den = w + 0.0000001f;
    for (int c = 0; c < C; ++c) render[(b * C + c) * P + pix] = out[(b * (C + 1) + c) * P + pix] / den;
    existing[i] = w;
  }
}

/* render_pointcloud (common.py:428-686) end to end.  degrid_mode as in kbo_degrid.  zee_out (optional,
 * [B,H,W]) receives the z-buffer AFTER the degrid pass; zee_raw (optional) the one before. */
KBO_API int kbo_render_pointcloud(const float *xyz, const float *data, int B, long N, int C, int W, int H,
                                  double focal, double baseline, int degrid_mode, float *render,
                                  float *existing, float *zee_raw, float *zee_out) {
  const long P = (long)H * W;
  float *z0 = (float *)kbo_scratch(0, sizeof(float) * B * P);
  float *z1 = (float *)kbo_scratch(1, sizeof(float) * B * P);
  float *acc = (float *)kbo_scratch(2, sizeof(float) * B * (C + 1) * P);
  if (!z0 || !z1 || !acc) {
    return -1;
  }
  kbo_splat_min(xyz, B, N, focal, baseline, z0, H, W, NULL);
  if (zee_raw) memcpy(zee_raw, z0, sizeof(float) * B * P);
  kbo_degrid(z0, z1, B, H, W, degrid_mode);
  if (zee_out) memcpy(zee_out, z1, sizeof(float) * B * P);
  kbo_splat_accum(xyz, data, B, N, C, focal, baseline, z1, acc, H, W);
  kbo_normalize(acc, B, C, H, W, render, existing);
  return 0;
}

/* kernel_discfill_updateOutput (common.py:837-924).  input [B,C,H,W], depth [B,1,H,W] -> output (a clone
 * of input with hole pixels, depth <= 0, overwritten from the farther end of the shortest of 16 rays).
 * fill_xy (optional, [B,H,W,2] int32) receives the source pixel chosen for each hole (-1,-1 otherwise). */
static void fill_impl(const float *input, const float *depth, int B, int C, int H, int W, float *output,
                      int32_t *fill_xy, int32_t *ends) {
  static const float dx0[16] = {-1, 0, 1, 1, -1, 1, 2, 2, -2, -1, 1, 2, 3, 3, 3, 3};
  static const float dy0[16] = {1, 1, 1, 0, 2, 2, 1, -1, 3, 3, 3, 3, 2, 1, -1, -2};
  float dirx[16], diry[16];
  for (int k = 0; k < 16; ++k) { /* :862-867 */
    const float nrm = sqrtf((dx0[k] * dx0[k]) + (dy0[k] * dy0[k]));
    dirx[k] = dx0[k] / nrm;
    diry[k] = dy0[k] / nrm;
  }
  const long P = (long)H * W;
  memcpy(output, input, sizeof(float) * B * C * P);
#pragma omp parallel for schedule(dynamic, 1024)
  for (long i = 0; i < B * P; ++i) {
    const long b = i / P;
    const int y = (int)((i % P) / W), x = (int)(i % W);
    const float *dep = depth + b * P;
    if (fill_xy) fill_xy[2 * i] = fill_xy[2 * i + 1] = -1;
    if (ends) ends[4 * i] = ends[4 * i + 1] = ends[4 * i + 2] = ends[4 * i + 3] = -1;
    if (dep[(long)y * W + x] > 0.0f) continue;
    float shortest = 1000000.0f;
    int fx = -1, fy = -1;
    int e0 = -1, e1 = -1, e2 = -1, e3 = -1;
    for (int k = 0; k < 16; ++k) {
      float ax = (float)x, ay = (float)y, bx = (float)x, by = (float)y;
      int iax = 0, iay = 0, ibx = 0, iby = 0;
      for (;;) { /* :876-883 */
        ax -= dirx[k]; iax = (int)roundf(ax);
        ay -= diry[k]; iay = (int)roundf(ay);
        if ((iax < 0) | (iax >= W)) break;
        if ((iay < 0) | (iay >= H)) break;
        if (dep[(long)iay * W + iax] > 0.0f) break;
      }
      if ((iax < 0) | (iax >= W)) continue;
      if ((iay < 0) | (iay >= H)) continue;
      for (;;) { /* :887-894 */
        bx += dirx[k]; ibx = (int)roundf(bx);
        by += diry[k]; iby = (int)roundf(by);
        if ((ibx < 0) | (ibx >= W)) break;
        if ((iby < 0) | (iby >= H)) break;
        if (dep[(long)iby * W + ibx] > 0.0f) break;
      }
      if ((ibx < 0) | (ibx >= W)) continue;
      if ((iby < 0) | (iby >= H)) continue;
      const float ddx = (float)(ibx - iax), ddy = (float)(iby - iay);
      const float dist = sqrtf(ddx * ddx + ddy * ddy); /* powf(int,2): exact for these magnitudes */
      if (shortest > dist) {
        fx = iax; fy = iay;
        if (dep[(long)iay * W + iax] < dep[(long)iby * W + ibx]) { fx = ibx; fy = iby; }
        shortest = dist;
        e0 = iax; e1 = iay; e2 = ibx; e3 = iby;
      }
    }
    if (fx == -1 || fy == -1) continue;
    if (fill_xy) { fill_xy[2 * i] = fx; fill_xy[2 * i + 1] = fy; }
    if (ends) { ends[4 * i] = e0; ends[4 * i + 1] = e1; ends[4 * i + 2] = e2; ends[4 * i + 3] = e3; }
    for (int c = 0; c < C; ++c)
      output[(b * C + c) * P + (long)y * W + x] = input[(b * C + c) * P + (long)fy * W + fx];
  }
}

KBO_API void kbo_fill(const float *input, const float *depth, int B, int C, int H, int W, float *output,
                      int32_t *fill_xy) {
  fill_impl(input, depth, B, C, H, W, output, fill_xy, NULL);
}

/* Same, also reporting both end points (from-x, from-y, to-x, to-y; -1 for non-holes) of the winning ray of every hole: the
 * only decision of the fill that depends on floating-point noise is which of the two is farther (:904-907), so a test can
 * tell which hole pixels are decided by a depth tie. */
KBO_API void kbo_fill_ends(const float *input, const float *depth, int B, int C, int H, int W, float *output,
                           int32_t *fill_xy, int32_t *ends) {
  fill_impl(input, depth, B, C, H, W, output, fill_xy, ends);
}

/* common.py:255 -- (render[0,0:3].transpose(1,2,0) * 255.0).clip(0,255).astype(uint8): fp32 multiply,
 * clip, truncation toward zero.  render is [>=3 planes of P]; out is HWC uint8. */
KBO_API void kbo_to_uint8(const float *render, int H, int W, uint8_t *out) {
  const long P = (long)H * W;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < P; ++i)
    for (int c = 0; c < 3; ++c) {
      float v = render[c * P + i] * 255.0f;
      v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
      out[3 * i + c] = (uint8_t)v;
    }
}

/* ---- OpenCV restatements (third-party: opencv-python 4.13.0, the version in this image; pinned against
 * cv2 itself in tests/test_oracle_tail.py) ------------------------------------------------------- */

static inline int cv_round(double v) { return (int)lrint(v); } /* round-half-even, like cvRound */
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }

/* cv2.getRectSubPix for 8UC3 -> 8UC3 (common.py:256): fixed-point bilinear, 16 fractional bits,
 * replicated border.  src HWC [sh,sw,3], dst [ph,pw,3]. */
KBO_API void kbo_getrectsubpix_8u3(const uint8_t *src, int sh, int sw, int pw, int ph, double cx, double cy,
                                   uint8_t *dst) {
  float fx = (float)cx, fy = (float)cy;
  fx -= (pw - 1) * 0.5f;
  fy -= (ph - 1) * 0.5f;
  const int ipx = cv_floor(fx), ipy = cv_floor(fy);
  const float a = fx - ipx, b = fy - ipy;
  const int a11 = cv_round((1.f - a) * (1.f - b) * (1 << 16));
  const int a12 = cv_round(a * (1.f - b) * (1 << 16));
  const int a21 = cv_round((1.f - a) * b * (1 << 16));
  const int a22 = cv_round(a * b * (1 << 16));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < ph; ++i) {
    int y0 = ipy + i, y1 = ipy + i + 1;
    y0 = y0 < 0 ? 0 : (y0 > sh - 1 ? sh - 1 : y0);
    y1 = y1 < 0 ? 0 : (y1 > sh - 1 ? sh - 1 : y1);
    for (int j = 0; j < pw; ++j) {
      int x0 = ipx + j, x1 = ipx + j + 1;
      x0 = x0 < 0 ? 0 : (x0 > sw - 1 ? sw - 1 : x0);
      x1 = x1 < 0 ? 0 : (x1 > sw - 1 ? sw - 1 : x1);
      for (int c = 0; c < 3; ++c) {
        const int s = src[((long)y0 * sw + x0) * 3 + c] * a11 + src[((long)y0 * sw + x1) * 3 + c] * a12 +
                      src[((long)y1 * sw + x0) * 3 + c] * a21 + src[((long)y1 * sw + x1) * 3 + c] * a22;
        dst[((long)i * pw + j) * 3 + c] = (uint8_t)((s + (1 << 15)) >> 16);
      }
    }
  }
}

/* Per-axis tables of cv2.resize(INTER_LINEAR) for 8-bit input: source index and 11-bit coefficients.
 * OpenCV clamps the fractional part only on the x axis (is_x=1); on the y axis it keeps the fraction
 * and clips the two row indices when they are used, so border rows blend a row with itself. */
KBO_API void kbo_resize_tables(int ssize, int dsize, int is_x, int32_t *ofs, int16_t *coef /*[dsize*2]*/) {
  const double inv_scale = (double)dsize / ssize;
  const double scale = 1. / inv_scale;
  for (int d = 0; d < dsize; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = cv_floor(f);
    f -= s;
    if (is_x) {
      if (s < 0) { f = 0; s = 0; }
      if (s >= ssize - 1) { f = 0; s = ssize - 1; }
    }
    ofs[d] = s;
    int c0 = cv_round((1.f - f) * 2048), c1 = cv_round(f * 2048);
    coef[2 * d] = (int16_t)c0;
    coef[2 * d + 1] = (int16_t)c1;
  }
}

/* cv2.resize(src, (dw,dh), interpolation=INTER_LINEAR) for 8UC3 (common.py:257). */
KBO_API void kbo_resize_linear_8u3(const uint8_t *src, int sh, int sw, int dh, int dw, uint8_t *dst) {
  int32_t *xo = (int32_t *)malloc(sizeof(int32_t) * dw), *yo = (int32_t *)malloc(sizeof(int32_t) * dh);
  int16_t *xa = (int16_t *)malloc(sizeof(int16_t) * 2 * dw), *ya = (int16_t *)malloc(sizeof(int16_t) * 2 * dh);
  kbo_resize_tables(sw, dw, 1, xo, xa);
  kbo_resize_tables(sh, dh, 0, yo, ya);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < dh; ++y) {
    int y0 = yo[y], y1 = yo[y] + 1;
    y0 = y0 < 0 ? 0 : (y0 > sh - 1 ? sh - 1 : y0);
    y1 = y1 < 0 ? 0 : (y1 > sh - 1 ? sh - 1 : y1);
    const int b0 = ya[2 * y], b1 = ya[2 * y + 1];
    for (int x = 0; x < dw; ++x) {
      const int x0 = xo[x], x1 = x0 + 1 < sw ? x0 + 1 : sw - 1;
      const int a0 = xa[2 * x], a1 = xa[2 * x + 1];
      for (int c = 0; c < 3; ++c) {
        const int r0 = src[((long)y0 * sw + x0) * 3 + c] * a0 + src[((long)y0 * sw + x1) * 3 + c] * a1;
        const int r1 = src[((long)y1 * sw + x0) * 3 + c] * a0 + src[((long)y1 * sw + x1) * 3 + c] * a1;
        dst[((long)y * dw + x) * 3 + c] = (uint8_t)((((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2);
      }
    }
  }
  free(xo); free(yo); free(xa); free(ya);
}

/* The per-frame body of process_kenburns' loop (common.py:238-257) on an already shifted cloud:
 * render (C=4: RGB + depth) -> fill_disocclusion(render, render[3]*(existing>0)) -> uint8 -> crop -> resize.
 * frame: uint8 [H,W,3].  scratch-free convenience used by the CPU baseline. */
KBO_API int kbo_frame(const float *xyz_shifted, const float *rgbd, long N, int W, int H, double focal,
                      double baseline, int crop_w, int crop_h, int degrid_mode, uint8_t *frame) {
  const long P = (long)H * W;
  float *render = (float *)kbo_scratch(3, sizeof(float) * 4 * P);
  float *existing = (float *)kbo_scratch(4, sizeof(float) * P);
  float *dmask = (float *)kbo_scratch(5, sizeof(float) * P);
  float *filled = (float *)kbo_scratch(6, sizeof(float) * 4 * P);
  uint8_t *u8 = (uint8_t *)kbo_scratch(7, 3 * P);
  uint8_t *patch = (uint8_t *)kbo_scratch(8, (size_t)3 * crop_w * crop_h);
  if (!render || !existing || !dmask || !filled || !u8 || !patch) return -1;
  int rc = kbo_render_pointcloud(xyz_shifted, rgbd, 1, N, 4, W, H, focal, baseline, degrid_mode, render,
                                 existing, NULL, NULL);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < P; ++i) dmask[i] = render[3 * P + i] * (existing[i] > 0.0f ? 1.0f : 0.0f);
  kbo_fill(render, dmask, 1, 4, H, W, filled, NULL);
  kbo_to_uint8(filled, H, W, u8);
  kbo_getrectsubpix_8u3(u8, H, W, crop_w, crop_h, W / 2.0, H / 2.0, patch);
  kbo_resize_linear_8u3(patch, crop_h, crop_w, H, W, frame);
  return rc;
}

/* spatial_filter(x, 'median-5') on a {0,1} map (common.py:417-421): reflect pad 2, 25-way median ==
 * (5x5 box count >= 13).  Used by pointcloud_inpainting.py:208-209. */
KBO_API void kbo_median5_binary(const float *in, int H, int W, float *out) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int cnt = 0;
      for (int dy = -2; dy <= 2; ++dy)
        for (int dx = -2; dx <= 2; ++dx) {
          int yy = y + dy, xx = x + dx;
          yy = yy < 0 ? -yy : (yy >= H ? 2 * H - 2 - yy : yy);
          xx = xx < 0 ? -xx : (xx >= W ? 2 * W - 2 - xx : xx);
          cnt += in[(long)yy * W + xx] > 0.5f;
        }
      out[(long)y * W + x] = cnt >= 13 ? 1.0f : 0.0f;
    }
}
