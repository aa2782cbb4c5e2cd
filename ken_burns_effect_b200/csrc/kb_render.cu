// kb_render.cu -- render_pointcloud / fill_disocclusion of the reference as sm_100a kernels behind the C ABI.
//
// Reference: utils/common.py:428-686 (render_pointcloud: updateZee, updateDegrid, updateOutput, epilogue)
//            utils/common.py:833-937 (fill_disocclusion), :104-109 (process_shift tensor half),
//            :417-421 (median-5, here for binary masks).
// These are the general (any B, any C, NCHW in/out) operators behind the reference's op API.  The fused
// multi-pose frame loop lives in kb_frames.cu.
//
// All kernels are scatter / stencil / gather work bound by L2+HBM traffic and atomic throughput, not by
// math: no tensor cores here.  Design notes (DESIGN.md has the numbers):
//   * points are read coalesced from the reference's own SoA layout [B,3,N] / [B,C,N];
//   * the z-buffer min is one native RED.MIN.S32 per point (positive floats order like ints);
//   * accumulators are channels-last so that one point/neighbour issues ceil((C+1)/4) 16-byte
//     RED.ADD.F32x4 instead of C+1 scalar atomics into planes H*W apart (20 -> 8 at C=4, 276 -> 72 at C=68);
//   * grids are sized in whole waves of 148 SMs where the element count allows.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "kb_common.cuh"

namespace kb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_fill_f32(float *__restrict__ p, long n, float v) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

__global__ void __launch_bounds__(256) k_shift_points(const float *__restrict__ xyz, const float *__restrict__ shift,
                                                      float *__restrict__ out, long N) {
  const int b = blockIdx.y;
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float *s = xyz + (long)b * 3 * N;
  float x = s[n], y = s[N + n], z = s[2 * N + n];
  shift_point(x, y, z, shift[b * 3 + 0], shift[b * 3 + 1], shift[b * 3 + 2]);
  float *o = out + (long)b * 3 * N;
  o[n] = x;
  o[N + n] = y;
  o[2 * N + n] = z;
}

struct Shift3 {
  float x, y, z;
  int on;
};

// updateZee: one thread per point.
__global__ void __launch_bounds__(256) k_splat_min(const float *__restrict__ xyz, long N, Shift3 sh, Camera cam,
                                                   float *__restrict__ zee, int32_t *__restrict__ pix_idx) {
  const int b = blockIdx.y;
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float *s = xyz + (long)b * 3 * N;
  float x = __ldg(s + n), y = __ldg(s + N + n), z = __ldg(s + 2 * N + n);
  if (sh.on) shift_point(x, y, z, sh.x, sh.y, sh.z);
  int32_t chosen = -1;
  Proj p;
  if (project(x, y, z, cam, p)) {
    const int k = pick_neighbour(p);
    if (k >= 0) {
      const int px = p.nwx + (k & 1), py = p.nwy + (k >> 1);
      if ((px >= 0) & (px < cam.W) & (py >= 0) & (py < cam.H)) {
        chosen = py * cam.W + px;
        zmin(zee + (long)b * cam.H * cam.W + chosen, p.err);
      }
    }
  }
  if (pix_idx) pix_idx[(long)b * N + n] = chosen;
}

// updateDegrid, race-free (read zin, write zout).  One thread per pixel; the 3x3 neighbourhood comes
// through L1 (each value is reused by 9 threads of the same CTA row band).
__global__ void __launch_bounds__(256) k_degrid(const float *__restrict__ zin, float *__restrict__ zout, int H, int W) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const long base = (long)blockIdx.z * H * W;
  const float *z = zin + base;
  const float c = z[(long)y * W + x];
  int count = 0;
  float sum = 0.0f;
  const int ox[4] = {1, 0, 1, 1};
  const int oy[4] = {0, 1, 1, -1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x1 = x + ox[k], y1 = y + oy[k], x2 = x - ox[k], y2 = y - oy[k];
    if ((x1 < 0) | (x1 >= W) | (y1 < 0) | (y1 >= H)) continue;
    if ((x2 < 0) | (x2 >= W) | (y2 < 0) | (y2 >= H)) continue;
    const float a = z[(long)y1 * W + x1], d = z[(long)y2 * W + x2];
    if (ge_plus_one(c, a) && ge_plus_one(c, d)) {   // :556-561, exact fp32 form of the fp64 comparison (kb_common.cuh)
      count += 2;
      sum = __fadd_rn(sum, a);
      sum = __fadd_rn(sum, d);
    }
  }
  float r = c;
  if (count > 0) r = fminf(c, __fdiv_rn(sum, (float)count));
  zout[base + (long)y * W + x] = r;
}

// updateOutput: one thread per point; channels in groups of 4 -> one RED.ADD.F32x4 per neighbour and group.
// data is the reference's [B,C,N]; the ones channel is synthesised (weight itself), :429.
__global__ void __launch_bounds__(256) k_splat_accum(const float *__restrict__ xyz, const float *__restrict__ data,
                                                     long N, int C, int Cp, Shift3 sh, Camera cam,
                                                     const float *__restrict__ zee, float *__restrict__ accum) {
  const int b = blockIdx.y;
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float *s = xyz + (long)b * 3 * N;
  float x = __ldg(s + n), y = __ldg(s + N + n), z = __ldg(s + 2 * N + n);
  if (sh.on) shift_point(x, y, z, sh.x, sh.y, sh.z);
  Proj p;
  if (!project(x, y, z, cam, p)) return;
  const long P = (long)cam.H * cam.W;
  const float *zb = zee + (long)b * P;
  float w[4] = {p.wnw, p.wne, p.wsw, p.wse};
  long pix[4];
  bool on[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int px = p.nwx + (k & 1), py = p.nwy + (k >> 1);
    on[k] = (px >= 0) & (px < cam.W) & (py >= 0) & (py < cam.H);
    pix[k] = on[k] ? (long)py * cam.W + px : 0;
    if (on[k]) on[k] = z_gate(p.err, __ldg(zb + pix[k]));
    // a zero weight adds exact zeros to every channel: skipping it changes no sum
    if (on[k]) on[k] = (w[k] != 0.0f);
  }
  if (!(on[0] | on[1] | on[2] | on[3])) return;
  const float *d = data + (long)b * C * N + n;
  float *ab = accum + (long)b * P * Cp;
  for (int c0 = 0; c0 < Cp; c0 += 4) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + j;
      v[j] = (c < C) ? __ldg(d + (long)c * N) : (c == C ? 1.0f : 0.0f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!on[k]) continue;
      red_add_v4(ab + pix[k] * Cp + c0, __fmul_rn(v[0], w[k]), __fmul_rn(v[1], w[k]), __fmul_rn(v[2], w[k]),
                 __fmul_rn(v[3], w[k]));
    }
  }
}


// updateOutput with the per-point data given as ROWS (point-major, `rs` floats per point, channels contiguous): the layout the
// convolutions produce (NHWC), so that the 64 context features of pointcloud_inpainting (:199-206) go from the conv epilogue
// into the splat without the NHWC -> NCHW -> torch.cat round trip (0.26 ms and 430 MB of temporaries per inpainting pass).
__global__ void __launch_bounds__(256) k_splat_accum_rows(const float *__restrict__ xyz, const float *__restrict__ rows, long rs,
                                                          long N, int C, int Cp, Shift3 sh, Camera cam,
                                                          const float *__restrict__ zee, float *__restrict__ accum) {
  const int b = blockIdx.y;
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float *s = xyz + (long)b * 3 * N;
  float x = __ldg(s + n), y = __ldg(s + N + n), z = __ldg(s + 2 * N + n);
  if (sh.on) shift_point(x, y, z, sh.x, sh.y, sh.z);
  Proj p;
  if (!project(x, y, z, cam, p)) return;
  const long P = (long)cam.H * cam.W;
  const float *zb = zee + (long)b * P;
  float w[4] = {p.wnw, p.wne, p.wsw, p.wse};
  long pix[4];
  bool on[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int px = p.nwx + (k & 1), py = p.nwy + (k >> 1);
    on[k] = (px >= 0) & (px < cam.W) & (py >= 0) & (py < cam.H);
    pix[k] = on[k] ? (long)py * cam.W + px : 0;
    if (on[k]) on[k] = z_gate(p.err, __ldg(zb + pix[k]));
    if (on[k]) on[k] = (w[k] != 0.0f);
  }
  if (!(on[0] | on[1] | on[2] | on[3])) return;
  const float *d = rows + ((long)b * N + n) * rs;
  float *ab = accum + (long)b * P * Cp;
  for (int c0 = 0; c0 < Cp; c0 += 4) {
    float v[4];
    if (c0 + 4 <= C) {
      const float4 t = __ldg(reinterpret_cast<const float4 *>(d + c0));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        v[j] = (c < C) ? __ldg(d + c) : (c == C ? 1.0f : 0.0f);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!on[k]) continue;
      red_add_v4(ab + pix[k] * Cp + c0, __fmul_rn(v[0], w[k]), __fmul_rn(v[1], w[k]), __fmul_rn(v[2], w[k]), __fmul_rn(v[3], w[k]));
    }
  }
}

// The weight channel of channels-last accumulators as a [B,1,H,W] map (tensorExisting of :686 before any thresholding).
__global__ void __launch_bounds__(256) k_accum_weight(const float *__restrict__ accum, int C, int Cp, long total, float *__restrict__ w) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) w[i] = accum[i * Cp + C];
}

// In place: accum[p][c] = accum[p][c] / (accum[p][C] + 1e-7) * mask[p] for c < C (:686 and `render * existing`,
// pointcloud_inpainting.py:210), then accum[p][C] = mask[p] -- the rows become the network input cat([data, mask]) (:135).
__global__ void __launch_bounds__(256) k_normalize_rows(float *__restrict__ accum, int C, int Cp, long total, const float *__restrict__ mask) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;     // one thread per (pixel, group of 4 channels)
  const int groups = Cp / 4;
  const long pix = i / groups;
  const int c0 = (int)(i - pix * groups) * 4;
  if (pix >= total) return;
  float *row = accum + pix * Cp;
  const float den = __fadd_rn(row[C], 0.0000001f);
  const float m = mask ? mask[pix] : 1.0f;
  float4 v = *reinterpret_cast<float4 *>(row + c0);
  float a[4] = {v.x, v.y, v.z, v.w};
  // the group that holds channel C writes the weight back unchanged, so the other groups of the pixel may read it at any
  // time; the mask replaces it in a second launch (kb_normalize_rows)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + j;
    if (c < C) a[j] = __fmul_rn(__fdiv_rn(a[j], den), m);
  }
  *reinterpret_cast<float4 *>(row + c0) = make_float4(a[0], a[1], a[2], a[3]);
}
__global__ void __launch_bounds__(256) k_set_mask_channel(float *__restrict__ accum, int C, int Cp, long total, const float *__restrict__ mask) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float *row = accum + i * Cp;
  row[C] = mask ? mask[i] : 1.0f;
  for (int c = C + 1; c < Cp; ++c) row[c] = 0.0f;
}

// epilogue :686 -- channels-last accumulators -> NCHW render + existing.  A CTA handles 32 pixels x all
// channels through shared memory so that both the channels-last reads and the planar writes coalesce.
__global__ void __launch_bounds__(256) k_normalize(const float *__restrict__ accum, int C, int Cp, long P,
                                                   float *__restrict__ render, float *__restrict__ existing) {
  extern __shared__ float tile[];  // [32][Cp+1]
  const int b = blockIdx.y;
  const long p0 = (long)blockIdx.x * 32;
  const int npx = (int)min((long)32, P - p0);
  const float *src = accum + ((long)b * P + p0) * Cp;
  const int ld = Cp + 1;
  for (int i = threadIdx.x; i < npx * Cp; i += blockDim.x) tile[(i / Cp) * ld + (i % Cp)] = src[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane < npx) {
    const float wsum = tile[lane * ld + C];
    const float den = __fadd_rn(wsum, 0.0000001f);
    for (int c = wrp; c < C; c += nw) render[((long)b * C + c) * P + p0 + lane] = __fdiv_rn(tile[lane * ld + c], den);
    if (wrp == 0) existing[(long)b * P + p0 + lane] = wsum;
  }
}

// fill_disocclusion :837-924, one thread per pixel (only hole pixels do any work).
__global__ void __launch_bounds__(256) k_fill(const float *__restrict__ input, const float *__restrict__ depth,
                                              float *__restrict__ output, int C, int H, int W) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const int b = blockIdx.z;
  const long P = (long)H * W;
  const float *dep = depth + (long)b * P;
  const float *in = input + (long)b * C * P;
  float *out = output + (long)b * C * P;
  const long me = (long)y * W + x;
  int fx = x, fy = y;
  if (!(dep[me] > 0.0f)) {
    float shortest = 1000000.0f;
    fx = -1;
    fy = -1;
    for (int k = 0; k < 16; ++k) {
      const float dx = c_dirx[k], dy = c_diry[k];
      float ax = (float)x, ay = (float)y, bx = (float)x, by = (float)y;
      int iax, iay, ibx, iby;
      for (;;) {
        ax = __fsub_rn(ax, dx); iax = (int)roundf(ax);
        ay = __fsub_rn(ay, dy); iay = (int)roundf(ay);
        if ((iax < 0) | (iax >= W)) break;
        if ((iay < 0) | (iay >= H)) break;
        if (dep[(long)iay * W + iax] > 0.0f) break;
      }
      if ((iax < 0) | (iax >= W)) continue;
      if ((iay < 0) | (iay >= H)) continue;
      for (;;) {
        bx = __fadd_rn(bx, dx); ibx = (int)roundf(bx);
        by = __fadd_rn(by, dy); iby = (int)roundf(by);
        if ((ibx < 0) | (ibx >= W)) break;
        if ((iby < 0) | (iby >= H)) break;
        if (dep[(long)iby * W + ibx] > 0.0f) break;
      }
      if ((ibx < 0) | (ibx >= W)) continue;
      if ((iby < 0) | (iby >= H)) continue;
      const float ddx = (float)(ibx - iax), ddy = (float)(iby - iay);
      const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)));
      if (shortest > dist) {
        fx = iax; fy = iay;
        if (dep[(long)iay * W + iax] < dep[(long)iby * W + ibx]) { fx = ibx; fy = iby; }
        shortest = dist;
      }
    }
    if (fx == -1 || fy == -1) { fx = x; fy = y; }   // no ray found: keep the clone's value
  }
  const long from = (long)fy * W + fx;
  for (int c = 0; c < C; ++c) out[(long)c * P + me] = in[(long)c * P + from];
}


// spatial_filter(x, 'laplacian'), utils/common.py:398-409: the reference's asymmetric 5-tap kernel (taps (0,1) = -1,
// (0,2) = -1, (1,1) = 4, (1,0) = -1, (2,0) = -1 of a 3x3 window) on a replicate-padded map, one channel at a time.
// The reference runs it as F.conv2d (cuDNN); the sum below keeps cuDNN's-agnostic left-to-right order of the five products.
__global__ void __launch_bounds__(256) k_laplacian5(const float *__restrict__ in, float *__restrict__ out, int H, int W) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const float *p = in + (long)blockIdx.z * H * W;
  const int xm = max(x - 1, 0), xp = min(x + 1, W - 1), ym = max(y - 1, 0), yp = min(y + 1, H - 1);
  // window rows: y-1 -> taps (0,1) at x, (0,2) at x+1;  y -> (1,0) at x-1, (1,1) at x;  y+1 -> (2,0) at x-1
  float acc = -p[(long)ym * W + x];
  acc = __fsub_rn(acc, p[(long)ym * W + xp]);
  acc = __fsub_rn(acc, p[(long)y * W + xm]);
  acc = __fmaf_rn(4.0f, p[(long)y * W + x], acc);
  acc = __fsub_rn(acc, p[(long)yp * W + xm]);
  out[(long)blockIdx.z * H * W + (long)y * W + x] = acc;
}

// median-5 on a {0,1} map with reflect padding == (5x5 count >= 13), :417-421.
__global__ void __launch_bounds__(256) k_median5_binary(const float *__restrict__ in, float *__restrict__ out, int H, int W) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const long base = (long)blockIdx.z * H * W;
  int cnt = 0;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
    int yy = y + dy;
    yy = yy < 0 ? -yy : (yy >= H ? 2 * H - 2 - yy : yy);
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      int xx = x + dx;
      xx = xx < 0 ? -xx : (xx >= W ? 2 * W - 2 - xx : xx);
      cnt += in[base + (long)yy * W + xx] > 0.5f;
    }
  }
  out[base + (long)y * W + x] = cnt >= 13 ? 1.0f : 0.0f;
}



// ---- generate_mask, utils/common.py:689-830 -----------------------------------------------------------------------------
// The reference marks, per point, whether it ended up owning the z-buffer cell it votes for, with a check-then-atomicMin
// and an atomicExch of point ids that race.  Deterministic form of the same bookkeeping, three passes:
// k_splat_min (z-buffer + the pixel each point votes for), k_mask_winner (lowest point index among the points that hold the
// minimal err of their pixel), k_mask_write (mask = 1 for the winner; point 0 keeps a 1 whenever it ever lowered its cell,
// the reference's `pid > 0` quirk, :759).  This is what the reference computes when its threads happen to run in index order.
__global__ void __launch_bounds__(256) k_mask_winner(const float *__restrict__ xyz, long N, Camera cam, const float *__restrict__ zee,
                                                     const int32_t *__restrict__ pix_idx, int *__restrict__ winner) {
  const int b = blockIdx.y;
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int pix = pix_idx[(long)b * N + n];
  if (pix < 0) return;
  const float *s = xyz + (long)b * 3 * N;
  Proj p;
  if (!project(__ldg(s + n), __ldg(s + N + n), __ldg(s + 2 * N + n), cam, p)) return;
  const long P = (long)cam.H * cam.W;
  if (p.err == zee[(long)b * P + pix]) atomicMin(winner + (long)b * P + pix, (int)n);
}

__global__ void __launch_bounds__(256) k_mask_write(long N, long P, const int32_t *__restrict__ pix_idx, const int *__restrict__ winner,
                                                    float *__restrict__ mask) {
  const int b = blockIdx.y;
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int pix = pix_idx[(long)b * N + n];
  float m = 0.0f;
  if (pix >= 0) {
    // in index order point 0 is the first to reach its pixel, so it always lowers the cell (1e6 > err) and is never cleared
    m = (winner[(long)b * P + pix] == (int)n || n == 0) ? 1.0f : 0.0f;
  }
  mask[(long)b * N + n] = m;
}

__global__ void __launch_bounds__(256) k_fill_i32(int *__restrict__ p, long n, int v) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

// ---- self-test of the exact fp32 replacements in kb_common.cuh / kb_frames.cu against the literal forms -----------
__device__ __forceinline__ unsigned char st_quant(float acc, float den) {
  float v = __fmul_rn(__fdiv_rn(acc, den), 255.0f);
  v = fminf(fmaxf(v, 0.0f), 255.0f);
  return (unsigned char)v;
}
__device__ __forceinline__ unsigned char st_quant_shared(float acc, float den) {
  // same sequence as kb_frames.cu: quant_shared()
  if (!((den >= 0x1p-60f) & (den <= 0x1p60f) & (fabsf(acc) <= 0x1p60f))) return st_quant(acc, den);
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(den));
  const float e = __fmaf_rn(-den, r0, 1.0f);
  const float r = __fmaf_rn(r0, e, r0);
  const float q0 = __fmul_rn(acc, r);
  const float rem = __fmaf_rn(-den, q0, acc);
  float v = __fmul_rn(__fmaf_rn(r, rem, q0), 255.0f);
  v = fminf(fmaxf(v, 0.0f), 255.0f);
  return (unsigned char)v;
}
__device__ __forceinline__ int st_round_away_i(float v) {
  const float magic = 12582912.0f;
  const float r = __fadd_rn(v, magic);
  int i = __float_as_int(r) - 0x4B400000;
  const float diff = __fsub_rn(v, __fsub_rn(r, magic));
  i += ((diff == 0.5f) & (v > 0.0f)) ? 1 : 0;
  i -= ((diff == -0.5f) & (v < 0.0f)) ? 1 : 0;
  return i;
}

__global__ void __launch_bounds__(256) k_selftest(int which, const float *__restrict__ a, const float *__restrict__ b, long n,
                                                  int W, unsigned long long *__restrict__ mismatches) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = a[i], y = b[i];
  bool bad = false;
  switch (which) {
    case 0:   // shared-reciprocal quantisation vs IEEE division; y plays the weight sum (>= 0)
      bad = st_quant_shared(x, __fadd_rn(fabsf(y), 0.0000001f)) != st_quant(x, __fadd_rn(fabsf(y), 0.0000001f));
      break;
    case 1:   // exact comparisons vs the fp64 forms
      bad = (ge_plus_one(x, y) != ((double)x >= __dadd_rn((double)y, 1.0))) ||
            (le_plus_one(x, y) != ((double)x <= __dadd_rn((double)y, 1.0)));
      break;
    case 2: { // floor / round without conversion instructions, |x| < 2^22
      if (!(fabsf(x) < kFloorRange)) break;
      float f;
      int k;
      floor_fi(x, f, k);
      bad = (f != floorf(x)) || (k != (int)floorf(x)) || (st_round_away_i(x) != (int)roundf(x));
      break;
    }
    case 3: { // pixel coordinate: fp32 sum vs the reference's two fp64 additions (W >= 2); z threshold; widen
      const float lit = __double2float_rn(__dadd_rn(__dadd_rn((double)x, 0.5 * (double)W), -0.5));
      bad = (W >= 2 && lit != __fadd_rn(x, (float)(0.5 * (double)W - 0.5))) || ((x >= 0.001f) != !((double)x < 0.001)) ||
            (x >= 0.001f && x < 3.0e38f && widen_normal(x) != (double)x);
      break;
    }
    default:
      break;
  }
  if (bad) atomicAdd(mismatches, 1ULL);
}

// ---------------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------------

Camera make_camera(double focal, double baseline, int H, int W) {
  Camera c;
  c.f32 = (float)focal;
  c.fB = focal * baseline;
  c.halfW = 0.5 * (double)W;
  c.halfH = 0.5 * (double)H;
  c.cx = (float)(0.5 * (double)W - 0.5);
  c.cy = (float)(0.5 * (double)H - 0.5);
  c.W = W;
  c.H = H;
  return c;
}

static Shift3 make_shift(const float *shift_host) {
  Shift3 s{0.f, 0.f, 0.f, 0};
  if (shift_host) {
    s.x = shift_host[0];
    s.y = shift_host[1];
    s.z = shift_host[2];
    s.on = 1;
  }
  return s;
}

}  // namespace kb

using namespace kb;

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int kb_version(void) { return 100; }
const char *kb_last_error(void) { return g_err; }
long long kb_launch_count(void) { return g_launches.load(); }

int kb_shift_points(const float *xyz, const float *shift, float *out, int B, long N, kb_stream_t stream) {
  KB_REQUIRE(xyz && shift && out && B > 0 && N > 0, "kb_shift_points: bad arguments");
  dim3 grid(cdiv(N, 256), B);
  k_shift_points<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, shift, out, N);
  count_launch();
  return check_launch("kb_shift_points");
}

int kb_splat_min(const float *xyz, int B, long N, const float *shift_host, double focal, double baseline,
                 float *zee, int H, int W, int32_t *pix_idx, kb_stream_t stream) {
  KB_REQUIRE(xyz && zee && B > 0 && N > 0 && H > 0 && W > 0, "kb_splat_min: bad arguments");
  KB_REQUIRE((long)H * W < (1L << 31) && H <= KB_MAX_SIDE && W <= KB_MAX_SIDE, "kb_splat_min: image too large");
  cudaStream_t st = (cudaStream_t)stream;
  const long nz = (long)B * H * W;
  k_fill_f32<<<min(cdiv(nz, 256), 148u * 8u), 256, 0, st>>>(zee, nz, 1000000.0f);
  dim3 grid(cdiv(N, 256), B);
  k_splat_min<<<grid, 256, 0, st>>>(xyz, N, make_shift(shift_host), make_camera(focal, baseline, H, W), zee, pix_idx);
  count_launch(2);
  return check_launch("kb_splat_min");
}

int kb_degrid(const float *zee_in, float *zee_out, int B, int H, int W, kb_stream_t stream) {
  KB_REQUIRE(zee_in && zee_out && zee_in != zee_out && B > 0 && H > 0 && W > 0, "kb_degrid: bad arguments");
  dim3 grid(cdiv(W, 32), cdiv(H, 8), B);
  k_degrid<<<grid, 256, 0, (cudaStream_t)stream>>>(zee_in, zee_out, H, W);
  count_launch();
  return check_launch("kb_degrid");
}

int kb_accum_channels(int C) { return (C + 1 + 3) & ~3; }

int kb_splat_accum(const float *xyz, const float *data, int B, long N, int C, const float *shift_host,
                   double focal, double baseline, const float *zee, float *accum, int H, int W,
                   kb_stream_t stream) {
  KB_REQUIRE(xyz && data && zee && accum && B > 0 && N > 0 && C > 0 && H > 0 && W > 0, "kb_splat_accum: bad arguments");
  KB_REQUIRE((long)H * W < (1L << 31) && H <= KB_MAX_SIDE && W <= KB_MAX_SIDE, "kb_splat_accum: image too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int Cp = kb_accum_channels(C);
  cudaError_t e = cudaMemsetAsync(accum, 0, sizeof(float) * (size_t)B * H * W * Cp, st);
  if (e != cudaSuccess) {
    set_error("kb_splat_accum memset: %s", cudaGetErrorString(e));
    return (int)e;
  }
  dim3 grid(cdiv(N, 256), B);
  k_splat_accum<<<grid, 256, 0, st>>>(xyz, data, N, C, Cp, make_shift(shift_host), make_camera(focal, baseline, H, W),
                                      zee, accum);
  count_launch(2);
  return check_launch("kb_splat_accum");
}

int kb_splat_accum_rows(const float *xyz, const float *data_rows, long row_stride, int B, long N, int C, const float *shift_host,
                        double focal, double baseline, const float *zee, float *accum, int H, int W, kb_stream_t stream) {
  KB_REQUIRE(xyz && data_rows && zee && accum && B > 0 && N > 0 && C > 0 && H > 0 && W > 0, "kb_splat_accum_rows: bad arguments");
  KB_REQUIRE(row_stride >= C && row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(data_rows) & 15) == 0,
             "kb_splat_accum_rows: rows must be 16-byte aligned with a stride that is a multiple of 4 floats and >= C");
  KB_REQUIRE((long)H * W < (1L << 31) && H <= KB_MAX_SIDE && W <= KB_MAX_SIDE, "kb_splat_accum_rows: image too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int Cp = kb_accum_channels(C);
  cudaError_t e = cudaMemsetAsync(accum, 0, sizeof(float) * (size_t)B * H * W * Cp, st);
  if (e != cudaSuccess) {
    set_error("kb_splat_accum_rows memset: %s", cudaGetErrorString(e));
    return (int)e;
  }
  dim3 grid(cdiv(N, 256), B);
  k_splat_accum_rows<<<grid, 256, 0, st>>>(xyz, data_rows, row_stride, N, C, Cp, make_shift(shift_host),
                                           make_camera(focal, baseline, H, W), zee, accum);
  count_launch(2);
  return check_launch("kb_splat_accum_rows");
}

int kb_accum_weight(const float *accum, int B, int C, int H, int W, float *weight, kb_stream_t stream) {
  KB_REQUIRE(accum && weight && B > 0 && C > 0 && H > 0 && W > 0, "kb_accum_weight: bad arguments");
  const long total = (long)B * H * W;
  k_accum_weight<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(accum, C, kb_accum_channels(C), total, weight);
  count_launch();
  return check_launch("kb_accum_weight");
}

int kb_normalize_rows(float *accum, int B, int C, int H, int W, const float *mask, kb_stream_t stream) {
  KB_REQUIRE(accum && B > 0 && C > 0 && H > 0 && W > 0, "kb_normalize_rows: bad arguments");
  const int Cp = kb_accum_channels(C);
  const long total = (long)B * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  k_normalize_rows<<<cdiv(total * (Cp / 4), 256), 256, 0, st>>>(accum, C, Cp, total, mask);
  k_set_mask_channel<<<cdiv(total, 256), 256, 0, st>>>(accum, C, Cp, total, mask);   // after every group has read the weight
  count_launch(2);
  return check_launch("kb_normalize_rows");
}

int kb_normalize(const float *accum, int B, int C, int H, int W, float *render, float *existing,
                 kb_stream_t stream) {
  KB_REQUIRE(accum && render && existing && B > 0 && C > 0 && H > 0 && W > 0, "kb_normalize: bad arguments");
  const int Cp = kb_accum_channels(C);
  const long P = (long)H * W;
  dim3 grid(cdiv(P, 32), B);
  const size_t smem = sizeof(float) * 32 * (Cp + 1);
  k_normalize<<<grid, 256, smem, (cudaStream_t)stream>>>(accum, C, Cp, P, render, existing);
  count_launch();
  return check_launch("kb_normalize");
}

size_t kb_render_workspace_bytes(int B, int C, int H, int W) {
  const size_t P = (size_t)H * W;
  return sizeof(float) * (size_t)B * P * (2 + (size_t)kb_accum_channels(C));
}

int kb_render_pointcloud(const float *xyz, const float *data, int B, long N, int C, int W, int H,
                         double focal, double baseline, float *render, float *existing, void *workspace,
                         kb_stream_t stream) {
  KB_REQUIRE(workspace, "kb_render_pointcloud: workspace is null");
  const size_t P = (size_t)H * W;
  float *z0 = (float *)workspace;
  float *z1 = z0 + (size_t)B * P;
  float *acc = z1 + (size_t)B * P;
  int rc = kb_splat_min(xyz, B, N, nullptr, focal, baseline, z0, H, W, nullptr, stream);
  if (rc) return rc;
  rc = kb_degrid(z0, z1, B, H, W, stream);
  if (rc) return rc;
  rc = kb_splat_accum(xyz, data, B, N, C, nullptr, focal, baseline, z1, acc, H, W, stream);
  if (rc) return rc;
  return kb_normalize(acc, B, C, H, W, render, existing, stream);
}

int kb_fill(const float *input, const float *depth, float *output, int B, int C, int H, int W,
            kb_stream_t stream) {
  KB_REQUIRE(input && depth && output && input != output && B > 0 && C > 0 && H > 0 && W > 0, "kb_fill: bad arguments");
  dim3 grid(cdiv(W, 32), cdiv(H, 8), B);
  k_fill<<<grid, 256, 0, (cudaStream_t)stream>>>(input, depth, output, C, H, W);
  count_launch();
  return check_launch("kb_fill");
}

int kb_median5_binary(const float *in, float *out, int B, int H, int W, kb_stream_t stream) {
  KB_REQUIRE(in && out && in != out && B > 0 && H > 2 && W > 2, "kb_median5_binary: bad arguments");
  dim3 grid(cdiv(W, 32), cdiv(H, 8), B);
  k_median5_binary<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, H, W);
  count_launch();
  return check_launch("kb_median5_binary");
}

size_t kb_mask_workspace_bytes(int B, long N, int H, int W) {
  return sizeof(float) * (size_t)B * H * W + sizeof(int) * (size_t)B * H * W + sizeof(int32_t) * (size_t)B * (size_t)N;
}

int kb_generate_mask(const float *xyz, int B, long N, double focal, double baseline, int H, int W, float *mask, void *workspace,
                     kb_stream_t stream) {
  KB_REQUIRE(xyz && mask && workspace && B > 0 && N > 0 && H > 0 && W > 0 && N < (1L << 31), "kb_generate_mask: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long P = (long)H * W;
  float *zee = (float *)workspace;
  int *winner = (int *)(zee + (size_t)B * P);
  int32_t *pix = (int32_t *)(winner + (size_t)B * P);
  int rc = kb_splat_min(xyz, B, N, nullptr, focal, baseline, zee, H, W, pix, stream);
  if (rc) return rc;
  k_fill_i32<<<min(cdiv((long)B * P, 256), 148u * 8u), 256, 0, st>>>(winner, (long)B * P, 0x7fffffff);
  dim3 grid(cdiv(N, 256), B);
  k_mask_winner<<<grid, 256, 0, st>>>(xyz, N, make_camera(focal, baseline, H, W), zee, pix, winner);
  k_mask_write<<<grid, 256, 0, st>>>(N, P, pix, winner, mask);
  count_launch(3);
  return check_launch("kb_generate_mask");
}

int kb_laplacian5(const float *in, float *out, int planes, int H, int W, kb_stream_t stream) {
  KB_REQUIRE(in && out && in != out && planes > 0 && H > 0 && W > 0, "kb_laplacian5: bad arguments");
  dim3 grid(cdiv(W, 32), cdiv(H, 8), planes);
  k_laplacian5<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, H, W);
  count_launch();
  return check_launch("kb_laplacian5");
}

int kb_selftest_arith(int which, const float *a, const float *b, long n, int W, unsigned long long *mismatches,
                      kb_stream_t stream) {
  KB_REQUIRE(a && b && mismatches && n > 0 && which >= 0 && which <= 3, "kb_selftest_arith: bad arguments");
  k_selftest<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(which, a, b, n, W, mismatches);
  count_launch();
  return check_launch("kb_selftest_arith");
}

}  // extern "C"

// ---- image front end: kbe.py:96-114, :181 of the reference ------------------------------------------------------------
// cv2.imread's uint8 HWC image -> transforms.ToTensor() (x / 255) -> transforms.Normalize(.5, .5) ((x - 0.5) / 0.5) -> crop of
// height and width to multiples of 4 -> (x + 1) / 2, as ONE pass on the device: the host sends 3 bytes per pixel instead of
// building three float images and sending 12.  Every operation is the IEEE fp32 operation torch's CPU kernels perform, in the
// same order, so the tensor is bit-identical to the reference's.
namespace kb {
__global__ void __launch_bounds__(256) k_image_front_end(const unsigned char *__restrict__ src, int H, int W, int Hc, int Wc,
                                                         int swap_rb, float *__restrict__ dst) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long Pc = (long)Hc * Wc;
  if (i >= Pc) return;
  const int y = (int)(i / Wc), x = (int)(i - (long)y * Wc);
  const unsigned char *px = src + ((long)y * W + x) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float u = (float)px[swap_rb ? 2 - c : c];
    const float t = __fdiv_rn(u, 255.0f);                                   // ToTensor
    const float n = __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f);                    // Normalize(mean .5, std .5)
    dst[c * Pc + i] = __fdiv_rn(__fadd_rn(n, 1.0f), 2.0f);                  // (tensorImage + 1) / 2, kbe.py:181
  }
}
}  // namespace kb

extern "C" int kb_image_front_end(const unsigned char *src, int H, int W, int swap_rb, float *dst, kb_stream_t stream) {
  KB_REQUIRE(src && dst && H >= 4 && W >= 4, "kb_image_front_end: bad arguments");
  const int Hc = H - H % 4, Wc = W - W % 4;
  kb::k_image_front_end<<<kb::cdiv((long)Hc * Wc, 256), 256, 0, (cudaStream_t)stream>>>(src, H, W, Hc, Wc, swap_rb, dst);
  kb::count_launch();
  return kb::check_launch("kb_image_front_end");
}
