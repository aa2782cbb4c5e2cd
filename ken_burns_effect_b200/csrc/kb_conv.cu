// kb_conv.cu -- 2-D convolution of the reference's conv stacks as an implicit GEMM on the 5th-gen tensor cores.
//
// Reference: every nn.Conv2d of models/disparity_estimation.py, models/disparity_refinement*.py,
// models/pointcloud_inpainting.py (3x3 stride 1/2, 1x1 shortcuts, the 7x7 stride-2 stem) together with
// what surrounds it there: bias, the per-channel PReLU that follows (or precedes the next conv), the
// residual / skip sum of the GridNet (pointcloud_inpainting.py:141-172).  cuDNN runs these as fp32 NCHW
// convolutions plus separate elementwise kernels; PyTorch's default on this GPU lets cuDNN use TF32.
//
// B200 design (one CTA = one tile of 128 output pixels x Npad output channels):
//   * activations are NHWC fp32 in HBM.  For every filter tap (r,s) and every 32-channel slice the A
//     operand of the GEMM is the input tile shifted by the tap -- one TMA box load {32 ch, tile_w, tile_h}
//     at (x0*stride + s - pad, y0*stride + r - pad); TMA zero-fills everything outside the image, which is
//     the convolution's zero padding, and its traversal stride is the convolution stride.  No im2col
//     buffer, no index arithmetic in the kernel.
//   * weights are pre-packed per (tap, 32-channel slice) as K-major [Cout][32] panels, TMA-loaded next to A.
//   * both operands land in shared memory in the 128-byte swizzled K-major layout that tcgen05.mma reads
//     through shared-memory descriptors; the accumulator (128 lanes x Npad fp32 columns) lives in TMEM.
//   * warp 0 = TMA producer, then the MMA issuer(s) (all run their loops as whole warps, one elected lane issues),
//     then the epilogue warps: tcgen05.ld the accumulator, + bias, partial-conv renormalisation, + residual, then up to
//     three outputs, each optionally passed through the next layer's PReLU and the consumer's mask, so that
//     "PReLU -> conv" chains never need a separate elementwise pass.
//   * a ring of `stages` shared-memory slots with full/empty mbarriers decouples TMA from the tensor pipe.
//   * operands are TF32 (fp32 tensors) or, per call, fp16 (kb_conv_args.x_f16: fp16 activations, fp16 filter panels, 64
//     channels per 128-byte chunk, kind::f16); accumulation, bias, residual and fp32 outputs are the same in both.
// Two kernels: k_conv_tf32 (one TMA load per filter tap, any filter; warps 0 / 1 / 2-9; two CTAs per SM so that one CTA's
// epilogue overlaps the other's main loop) and the persistent k_conv_halo_tf32 (stride 1, k <= 3: one halo load per tile, taps
// are descriptor offsets, resident filters, accumulator ring in TMEM, two MMA issuers, epilogue in teams of four warps) --
// see the comment above it.
// Environment switches (experiments; the defaults are what the measurements in DESIGN.md section 5 chose):
//   KB_CONV_DEBUG   bit mask, results are WRONG with any bit set: 1 no epilogue stores, 2 no MMAs, 4 no activation loads,
//                   8 no TMEM reads (1 and 8 select the generic epilogue) -- how "what bounds the kernel" was measured
//   KB_CONV_NO_LEAN=1 generic epilogue everywhere   KB_CONV_TEAMS=n epilogue teams (<= 4)   KB_CONV_ONE_ISSUER=1
//   KB_CONV_NO_WIDE=1 16-byte instead of 32-byte global loads / stores   KB_CONV_NO_RESIDENT=1   KB_CONV_ALGO=1|2
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <mutex>

#include "kb_common.cuh"

namespace kb {

// ---- PTX wrappers ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate; issued by ONE thread for the whole CTA.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with each shared-memory descriptor given as its two 32-bit halves: only the low word (start address) changes from
// one MMA to the next, so a step costs one 32-bit add instead of 64-bit arithmetic.
__device__ __forceinline__ void umma_tf32_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 (fp16 operands, fp32 accumulate): K = 16 halves = the same 32 bytes per MMA as kind::tf32, twice the MACs.
__device__ __forceinline__ void umma_f16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Operand kind is a template parameter of the kernels: a run-time choice inside the issue loop cost the TF32 path 10 %
// (two predicated UTCHMMA per step on the one thread everything waits for).
template <bool F16>
__device__ __forceinline__ void umma_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                        uint32_t accumulate) {
  if (F16) umma_f16_lh(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
  else umma_tf32_lh(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
}
// One lane of a converged warp.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// mbarrier arrive when every MMA issued so far by this thread has finished reading smem / writing TMEM.
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread l of the warp receives row (lane base + l).  The load is asynchronous:
// the registers are valid after tmem_ld_wait(), which lets the caller put other loads in flight in between.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
  // the "+r" ties make every use of r[] depend on the wait
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// K-major, 128-byte swizzled operand tile whose rows are 128 bytes: 8-row groups are 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) /* LBO (unused with swizzle) */ |
         (64ull << 32) /* SBO = 1024 B */ | (1ull << 46) /* descriptor version: sm_100 */ | (2ull << 61) /* SWIZZLE_128B */;
}

__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// ---- kernel ------------------------------------------------------------------------------------------------
constexpr int kTileM = 128;            // output pixels per CTA = UMMA M
constexpr int kChunk = 32;             // input channels per K step (32 fp32 = one 128-byte swizzle row)
constexpr int kABytes = kTileM * kChunk * 4;
constexpr int kConvThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
// persistent kernel: warp 0 TMA, warps 1-2 MMA issuers, then the epilogue warps (8 with the generic epilogue, teams of 4 with the lean one)
constexpr int kHaloIssuers = 2, kHaloEpiWarp0 = 1 + kHaloIssuers;
constexpr int kHaloThreads = 32 * kHaloEpiWarp0 + 256;        // generic epilogue: 8 warps on one tile
constexpr int kHaloMaxTeams = 4;                               // lean epilogue: up to four teams of four warps
constexpr int kHaloLeanThreads = 32 * kHaloEpiWarp0 + kHaloMaxTeams * 128;
constexpr int kMaxOut = 3;

struct ConvOut {
  float *ptr;
  long stride;          // floats per pixel
  const float *slope;   // per-channel PReLU applied to this output (nullptr = identity)
  const float *mul;     // per-pixel factor applied last (the {0,1} mask a partial-conv consumer multiplies its input with)
  int round_tf32;       // round to TF32 (the value only feeds convolutions)
  int f16;              // store as fp16 (ptr points at halves, stride counts halves): the input of a kind::f16 convolution
  int wide;             // every pixel of this output starts on a 32-byte boundary: the lean epilogue stores 32 bytes per instruction
};

struct ConvKernelParams {
  int Ho, Wo;           // stored output size (possibly cropped)
  int tile_w, tile_h, tiles_x, tiles_y;
  int chunks, ksize, stride, pad;
  int Cout, Cout4;      // real output channels, and rounded up to 4 (allocated)
  int last_ksteps;      // 8-channel MMA steps that hold real channels in the LAST 32-channel slice (1..4): the rest is zero padding
  int cpad;             // Cout rounded up to a whole number of N blocks (size of the epilogue tables in shared memory)
  int Npad;             // UMMA N of this launch
  int stages;
  uint32_t tmem_cols;
  const float *bias;
  const float *res;
  long res_stride;
  const float *pc_ratio;   // partial convolution (utils/partial_conv.py:62-77): per-pixel mask_ratio and update_mask, or nullptr
  const float *pc_um;
  int n_out;
  int vec;              // bias / slope arrays can be read as float4 (Cout % 4 == 0, 16-byte aligned)
  ConvOut out[kMaxOut];
  // persistent halo kernel only
  int pitch;            // halo row pitch in pixels (>= tile_w + ksize - 1)
  int a_stage_bytes;    // one halo slice (32 channels), rounded up to 1024
  int a_stages, b_stages;
  int acc_stages;       // accumulator ring depth in TMEM (2..kAccMax)
  int epi_off;          // byte offset of the epilogue tables from the aligned shared-memory base
  int resident;         // all weight panels stay in shared memory for the life of the CTA
  int n_blocks;
  long work_items;      // tiles * n_blocks
  int debug;            // KB_CONV_DEBUG bit mask (profiling experiments only, see the file header)
  int f16;              // operands are fp16 (activations and packed filters): kind::f16, 64 channels per 128-byte chunk
  int cpc;              // channels per chunk: 32 (tf32) or 64 (f16) -- a chunk is always one 128-byte swizzle row per pixel
  int fast;             // lean epilogue (Cout % 16 == 0, no debug bits that touch the epilogue): 1 dense, 2 with partial-conv
                        // renormalisation / per-pixel output factors; 0 = generic epilogue
  int res_wide;         // every pixel of the residual starts on a 32-byte boundary
  int issuers;          // MMA-issuing warps: 2 with resident filters (items alternate; each issuer has its own half of the
                        // activation ring, its own accumulator stages and its own epilogue team), else 1
  int teams;            // lean epilogue: teams of 4 warps, team t takes tiles t, t + teams, ... of the CTA (block = 96 + 128 * teams threads)
};


// Epilogue of one accumulator row (= one output pixel): TMEM -> registers, + bias, + residual, then every requested
// output (own PReLU, optional TF32 rounding) as 16-byte stores.  Warp-collective (tcgen05.ld).  Eight epilogue warps
// share a tile: warp e reads TMEM lanes 32*(e%4).. and the 16-column chunks of parity e/4.
// Channels Cout..round_up(Cout,4) come out as exact zeros without masking: their filter rows and bias are zero padding
// and the residual's own padding channels are zero by the same rule.
//
// ncu (profiles/ncu_conv_r01f_*): with bias / PReLU slopes read from global memory inside the loop, the eight epilogue
// warps spent the tile waiting on those (L2-latency) loads and the kernel ran epilogue-bound at 14-35 % tensor-pipe
// activity.  They now live in shared memory (staged once per CTA, zero-padded so every read is an unconditional float4),
// the residual of a chunk is requested BEFORE the accumulator is waited for, and the TMEM load overlaps both.
// PReLU as torch computes it (x >= 0 ? x : w * x): a compare and a predicated multiply.
__device__ __forceinline__ float prelu1(float v, float s) { return v < 0.f ? __fmul_rn(v, s) : v; }

// 4 consecutive channels of one output pixel as fp16 (8 bytes), saturated to the fp16 range
__device__ __forceinline__ void store_half4(float *base_as_float, long half_index, float4 w) {
  const float lim = 65504.0f;
  w.x = fminf(fmaxf(w.x, -lim), lim); w.y = fminf(fmaxf(w.y, -lim), lim);
  w.z = fminf(fmaxf(w.z, -lim), lim); w.w = fminf(fmaxf(w.w, -lim), lim);
  const __half2 lo = __floats2half2_rn(w.x, w.y), hi = __floats2half2_rn(w.z, w.w);
  uint2 v;
  v.x = *reinterpret_cast<const uint32_t *>(&lo);
  v.y = *reinterpret_cast<const uint32_t *>(&hi);
  *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(base_as_float) + half_index) = v;
}

struct EpiSmem {
  const float *bias;              // [cpad]   (zeros when the layer has no bias); the slopes of output o follow at
  int cpad;                       //          bias + (1 + o) * cpad (only valid where p.out[o].slope != nullptr)
};

// floats of shared memory the epilogue tables take for `cpad` (padded) output channels
__host__ __device__ constexpr int epi_smem_floats(int cpad) { return (1 + kMaxOut) * cpad; }

// Called by all nthr epilogue threads (tid 0..nthr-1) before their first tile; ends with a barrier among them.
__device__ __forceinline__ EpiSmem epi_stage(const ConvKernelParams &p, float *base, int cpad, int tid, int nthr = 256) {
  EpiSmem e;
  e.bias = base;
  e.cpad = cpad;
  for (int c = tid; c < cpad; c += nthr) base[c] = (p.bias && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o) {
    float *dst = base + (1 + o) * cpad;
    if (o < p.n_out && p.out[o].slope)
      for (int c = tid; c < cpad; c += nthr) dst[c] = c < p.Cout ? __ldg(p.out[o].slope + c) : 1.f;
  }
  asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
  return e;
}

struct EpiPixel {
  bool inside;
  long pix;               // index of this thread's output pixel in [N, Ho, Wo]
  const float *res;
  float ratio, um;        // partial-conv scalars of this pixel (1, 1 for a dense convolution)
};

__device__ __forceinline__ EpiPixel epi_pixel(const ConvKernelParams &p, int img, int oy, int ox) {
  EpiPixel e;
  e.inside = (oy < p.Ho) & (ox < p.Wo);
  e.pix = e.inside ? ((long)img * p.Ho + oy) * p.Wo + ox : 0;
  e.res = p.res ? p.res + e.pix * p.res_stride : nullptr;
  e.ratio = p.pc_ratio ? __ldg(p.pc_ratio + e.pix) : 1.f;      // requested here, consumed after the accumulator barrier
  e.um = p.pc_um ? __ldg(p.pc_um + e.pix) : 1.f;
  return e;
}

__device__ __forceinline__ void epi_load_res(const ConvKernelParams &p, const EpiPixel &px, int c, float4 (&rr)[4]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    rr[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (px.res && px.inside && c + 4 * g < p.Cout4) rr[g] = *reinterpret_cast<const float4 *>(px.res + c + 4 * g);
  }
}

// The residual of the FIRST chunk must already be in rr (epi_load_res before the accumulator barrier).
// Structure: the 16 accumulator values of a chunk are finished once (bias, partial-conv renormalisation, residual), then a
// ROLLED loop runs over the n_out outputs.  With the outputs unrolled the compiler predicated the code of all three instead
// of branching around the absent ones, and the eight epilogue warps issued ~800 instructions per tile -- after the MMA issue
// loop was fixed THEY were what the tensor pipe waited for (profiles/ncu_conv_r01q_0).
__device__ __forceinline__ void epilogue_rows(const ConvKernelParams &p, const EpiSmem &es, const EpiPixel &px, uint32_t taddr,
                                              int n0, int half, float4 (&rr)[4]) {
  for (int c0 = 16 * half; c0 < p.Npad; c0 += 32) {
    if (n0 + c0 >= p.Cout4) break;           // warp-uniform
    uint32_t raw[16];
    if (p.debug & 8) {   // experiment: no TMEM reads
#pragma unroll
      for (int i = 0; i < 16; ++i) raw[i] = 0;
    } else {
      tmem_ld16_issue(taddr + (uint32_t)c0, raw);
    }
    float4 rnext[4];
    const bool more = (c0 + 32 < p.Npad) && (n0 + c0 + 32 < p.Cout4);
    if (more) epi_load_res(p, px, n0 + c0 + 32, rnext);
    float4 bs[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) bs[g] = *reinterpret_cast<const float4 *>(es.bias + n0 + c0 + 4 * g);
    tmem_ld_wait(raw);
    if (px.inside) {
      float4 a[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        a[g] = make_float4(__uint_as_float(raw[4 * g]) + bs[g].x, __uint_as_float(raw[4 * g + 1]) + bs[g].y,
                           __uint_as_float(raw[4 * g + 2]) + bs[g].z, __uint_as_float(raw[4 * g + 3]) + bs[g].w);
      }
      if (p.pc_ratio) {
        // output = ((raw_out - bias) * mask_ratio + bias) * update_mask, utils/partial_conv.py:74-77, same operation order
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          a[g].x = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(a[g].x, bs[g].x), px.ratio), bs[g].x), px.um);
          a[g].y = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(a[g].y, bs[g].y), px.ratio), bs[g].y), px.um);
          a[g].z = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(a[g].z, bs[g].z), px.ratio), bs[g].z), px.um);
          a[g].w = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(a[g].w, bs[g].w), px.ratio), bs[g].w), px.um);
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        a[g].x += rr[g].x; a[g].y += rr[g].y; a[g].z += rr[g].z; a[g].w += rr[g].w;
      }
      const int ngroups = min(4, (p.Cout4 - (n0 + c0)) >> 2);      // float4 groups of this chunk that exist in the outputs
#pragma unroll 1
      for (int o = 0; o < p.n_out; ++o) {
        const ConvOut &out = p.out[o];
        float *dst = out.ptr + px.pix * out.stride + n0 + c0;
        const long hidx = px.pix * out.stride + n0 + c0;           // the same position counted in halves (out.f16)
        const float *slope = out.slope ? es.bias + (1 + o) * es.cpad + n0 + c0 : nullptr;
        const float m = out.mul ? __ldg(out.mul + px.pix) : 1.f;
        const bool rnd = out.round_tf32 != 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g >= ngroups) break;
          float4 w = a[g];
          if (slope) {
            const float4 sl = *reinterpret_cast<const float4 *>(slope + 4 * g);
            w.x = prelu1(w.x, sl.x); w.y = prelu1(w.y, sl.y); w.z = prelu1(w.z, sl.z); w.w = prelu1(w.w, sl.w);
          }
          if (out.mul) { w.x *= m; w.y *= m; w.z *= m; w.w *= m; }
          if (out.f16) {
            if (!(p.debug & 1)) store_half4(out.ptr, hidx + 4 * g, w);
            continue;
          }
          if (rnd) { w.x = round_tf32(w.x); w.y = round_tf32(w.y); w.z = round_tf32(w.z); w.w = round_tf32(w.w); }
          if (!(p.debug & 1)) *reinterpret_cast<float4 *>(dst + 4 * g) = w;
        }
      }
    }
    if (more) {
#pragma unroll
      for (int g = 0; g < 4; ++g) rr[g] = rnext[g];
    }
  }
}

// ---- the LEAN epilogue of the persistent kernel --------------------------------------------------------------------
// ncu source page of the 32 -> 32 layer at 768 x 1024 (profiles/ncu_conv_r02v_f16_0_summary.md): nothing on the SM is busy
// (tensor pipe 12 %, LSU 34 %, issue slots 43 %), yet a tile takes 2 700 cycles -- the time ONE epilogue warp needs for the
// ~440 dependent instructions epilogue_rows() costs it per tile (16 values per thread): generic-address loads of the
// shared-memory tables, 64-bit index arithmetic per output and per 4 channels, clamp pairs before the fp16 pack, copies
// that merge the PReLU / no-PReLU paths.  Two warps per scheduler cannot hide that chain, and dealing tiles to teams of
// warps does not shorten it.  The common layer (dense convolution, Cout a multiple of 16, no per-pixel factors) takes this
// form instead: explicit ld.shared of the tables, one address per output, cvt.satfinite for the fp16 pack, the two PReLU
// cases as separate straight-line paths.  Same arithmetic, same order, bit-identical results.
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// cvt.rna.tf32.f32 for every finite and infinite input in two integer instructions (ptxas expands the cvt into three)
__device__ __forceinline__ float round_tf32_bits(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
// two floats -> packed halves (lo in the low 16 bits), saturated to +-65504 like store_half4
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
  uint32_t h;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(hi), "f"(lo));
  return h;
}

// 32 bytes in one instruction (sm_100 has 256-bit global accesses).  With two 16-byte stores per 32-byte sector the layer ran at
// the speed of its stores: every warp-level store touches 32 different lines (one per pixel), each sector arrives in L2 in two
// halves, and with MMAs and loads switched off (KB_CONV_DEBUG=6) the 32 -> 32 layer still took 25 us (fp16 output) / 40 us (fp32
// output) of the 37 / 56 us of the whole kernel.
__device__ __forceinline__ void stg256(void *ptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g,
                                       uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f),
               "r"(g), "r"(h)
               : "memory");
}

// 16 consecutive floats of the residual: two 32-byte loads when the pixel starts on a 32-byte boundary, else four 16-byte ones
__device__ __forceinline__ void lean_load_res(const float *src, bool wide, float4 (&r)[4]) {
  if (wide) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=f"(r[2 * h].x), "=f"(r[2 * h].y), "=f"(r[2 * h].z), "=f"(r[2 * h].w), "=f"(r[2 * h + 1].x),
                     "=f"(r[2 * h + 1].y), "=f"(r[2 * h + 1].z), "=f"(r[2 * h + 1].w)
                   : "l"(src + 8 * h)
                   : "memory");
  } else {
#pragma unroll
    for (int g = 0; g < 4; ++g) r[g] = *(reinterpret_cast<const float4 *>(src) + g);
  }
}

__device__ __forceinline__ void lean_store16(const ConvOut &out, long pix, int c, const float (&w)[16]) {
  if (out.f16) {
    uint4 *d = reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(out.ptr) + pix * out.stride + c);
    const uint32_t h0 = pack_half2_sat(w[0], w[1]), h1 = pack_half2_sat(w[2], w[3]), h2 = pack_half2_sat(w[4], w[5]),
                   h3 = pack_half2_sat(w[6], w[7]), h4 = pack_half2_sat(w[8], w[9]), h5 = pack_half2_sat(w[10], w[11]),
                   h6 = pack_half2_sat(w[12], w[13]), h7 = pack_half2_sat(w[14], w[15]);
    if (out.wide) {
      stg256(d, h0, h1, h2, h3, h4, h5, h6, h7);
    } else {
      d[0] = make_uint4(h0, h1, h2, h3);
      d[1] = make_uint4(h4, h5, h6, h7);
    }
  } else {
    float4 *d = reinterpret_cast<float4 *>(out.ptr + pix * out.stride + c);
    uint32_t v[16];
    if (out.round_tf32) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(round_tf32_bits(w[i]));
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(w[i]);
    }
    if (out.wide) {
      stg256(d, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
      stg256(d + 2, v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g) reinterpret_cast<uint4 *>(d)[g] = make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    }
  }
}

// A value the epilogue reads every tile, pinned in a register: left to itself the compiler re-reads kernel parameters from the
// constant bank inside the dependent chains of the tile loop (20-40 cycles each with two warps per scheduler to hide them).
__device__ __forceinline__ int pin_reg(int v) { asm volatile("" : "+r"(v)); return v; }

// One 16-column chunk of this thread's accumulator row: channels c .. c+15 of pixel `pix`.  `cur` holds the residual of this
// chunk (loaded one chunk ahead), the residual of the next chunk (16 channels on) is requested into `nxt`.
// PCMUL: the layer has a partial-convolution renormalisation and / or per-pixel output factors (compiled out of the dense form:
// as run-time branches they cost the dense layers 2-3 % of a network forward -- registers, two spilled values, extra issue slots)
template <bool PCMUL>
__device__ __forceinline__ void lean_chunk(const ConvKernelParams &p, uint32_t tab, int cpad, int n_out, bool inside, long pix,
                                           const float *res, uint32_t taddr, int c, bool more, const float4 (&cur)[4],
                                           float4 (&nxt)[4], bool pc, float ratio, float um) {
  uint32_t raw[16];
  tmem_ld16_issue(taddr, raw);
  if (more && res && inside) lean_load_res(res + c + 16, p.res_wide != 0, nxt);
  float4 bs[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) bs[g] = lds_f4(tab + (uint32_t)(c + 4 * g) * 4u);
  tmem_ld_wait(raw);
  if (!inside) return;
  float a[16];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    a[4 * g] = __uint_as_float(raw[4 * g]) + bs[g].x;
    a[4 * g + 1] = __uint_as_float(raw[4 * g + 1]) + bs[g].y;
    a[4 * g + 2] = __uint_as_float(raw[4 * g + 2]) + bs[g].z;
    a[4 * g + 3] = __uint_as_float(raw[4 * g + 3]) + bs[g].w;
  }
  if (PCMUL && pc) {
    // output = ((raw_out - bias) * mask_ratio + bias) * update_mask, utils/partial_conv.py:74-77, same operation order
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      a[4 * g] = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(a[4 * g], bs[g].x), ratio), bs[g].x), um);
      a[4 * g + 1] = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(a[4 * g + 1], bs[g].y), ratio), bs[g].y), um);
      a[4 * g + 2] = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(a[4 * g + 2], bs[g].z), ratio), bs[g].z), um);
      a[4 * g + 3] = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(a[4 * g + 3], bs[g].w), ratio), bs[g].w), um);
    }
  }
  if (res) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      a[4 * g] += cur[g].x; a[4 * g + 1] += cur[g].y; a[4 * g + 2] += cur[g].z; a[4 * g + 3] += cur[g].w;
    }
  }
#pragma unroll 1
  for (int o = 0; o < n_out; ++o) {
    // `a` is loop-invariant, and the compiler hoists the TF32 rounding and the fp16 pack of the no-PReLU case out of the loop --
    // 40 instructions per chunk executed whether or not any output wants them.  The empty asm makes `a` opaque per iteration.
    asm volatile("" : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]), "+f"(a[4]), "+f"(a[5]), "+f"(a[6]), "+f"(a[7]), "+f"(a[8]),
                      "+f"(a[9]), "+f"(a[10]), "+f"(a[11]), "+f"(a[12]), "+f"(a[13]), "+f"(a[14]), "+f"(a[15]));
    const ConvOut &out = p.out[o];
    if (out.slope) {
      const uint32_t st = tab + (uint32_t)((1 + o) * cpad + c) * 4u;
      float w[16];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 sl = lds_f4(st + 16u * g);
        w[4 * g] = prelu1(a[4 * g], sl.x); w[4 * g + 1] = prelu1(a[4 * g + 1], sl.y);
        w[4 * g + 2] = prelu1(a[4 * g + 2], sl.z); w[4 * g + 3] = prelu1(a[4 * g + 3], sl.w);
      }
      if (PCMUL && out.mul) {
        const float m = __ldg(out.mul + pix);
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] *= m;
      }
      lean_store16(out, pix, c, w);
    } else if (PCMUL && out.mul) {
      const float m = __ldg(out.mul + pix);
      float w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = a[i] * m;
      lean_store16(out, pix, c, w);
    } else {
      lean_store16(out, pix, c, a);
    }
  }
}

// All 16-column chunks of one accumulator tile that belong to this warp: columns 0, 16, 32, ... of the N block at n0 (a team
// of four warps owns a tile, one warp per TMEM lane quadrant), two chunks per trip so that the residual buffers swap roles
// instead of being copied.  `tab` is the shared-memory address of the epilogue tables (bias, then the slopes of output o at
// (1 + o) * cpad floats); rr holds the residual of the first chunk.
template <bool PCMUL>
__device__ __forceinline__ void epilogue_chunks_lean(const ConvKernelParams &p, uint32_t tab, int cpad, int n_out, int Npad, int Cout,
                                                     bool inside, long pix, uint32_t taddr, int n0, float4 (&rr)[4], float ratio,
                                                     float um) {
  const float *res = p.res ? p.res + pix * p.res_stride : nullptr;
  const bool pc = PCMUL && p.pc_ratio != nullptr;
  float4 r2[4];
  int c0 = 0;
  for (;;) {
    bool more = (c0 + 16 < Npad) && (n0 + c0 + 16 < Cout);        // warp-uniform (Cout % 16 == 0: a chunk is whole or absent)
    lean_chunk<PCMUL>(p, tab, cpad, n_out, inside, pix, res, taddr + (uint32_t)c0, n0 + c0, more, rr, r2, pc, ratio, um);
    if (!more) break;
    c0 += 16;
    more = (c0 + 16 < Npad) && (n0 + c0 + 16 < Cout);
    lean_chunk<PCMUL>(p, tab, cpad, n_out, inside, pix, res, taddr + (uint32_t)c0, n0 + c0, more, r2, rr, pc, ratio, um);
    if (!more) break;
    c0 += 16;
  }
}

// ---- the same epilogue with the outputs UNROLLED, for the per-tap kernel: it is not persistent and runs two CTAs per SM,
// which the register budget of this form (96) allows and that of the rolled form (115) does not; there the epilogue is a
// tail of each CTA, not a steady-state stage, and fewer registers win.
struct EpiPixelU {
  bool inside;
  const float *res;
  float *dst[kMaxOut];
  float ratio, um;        // partial-conv scalars of this pixel (1, 1 for a dense convolution)
  float mul[kMaxOut];     // trailing per-pixel factor of each output (1 = none)
};

__device__ __forceinline__ EpiPixelU epi_pixel_u(const ConvKernelParams &p, int img, int oy, int ox) {
  EpiPixelU e;
  e.inside = (oy < p.Ho) & (ox < p.Wo);
  const long pix = e.inside ? ((long)img * p.Ho + oy) * p.Wo + ox : 0;
  e.res = p.res ? p.res + pix * p.res_stride : nullptr;
  e.ratio = p.pc_ratio ? __ldg(p.pc_ratio + pix) : 1.f;      // requested here, consumed after the accumulator barrier
  e.um = p.pc_um ? __ldg(p.pc_um + pix) : 1.f;
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o) {
    e.dst[o] = nullptr;
    if (o < p.n_out)
      e.dst[o] = p.out[o].f16 ? reinterpret_cast<float *>(reinterpret_cast<__half *>(p.out[o].ptr) + pix * p.out[o].stride)
                              : p.out[o].ptr + pix * p.out[o].stride;
    e.mul[o] = (o < p.n_out && p.out[o].mul) ? __ldg(p.out[o].mul + pix) : 1.f;
  }
  return e;
}

__device__ __forceinline__ void epi_load_res_u(const ConvKernelParams &p, const EpiPixelU &px, int c, float4 (&rr)[4]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    rr[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (px.res && px.inside && c + 4 * g < p.Cout4) rr[g] = *reinterpret_cast<const float4 *>(px.res + c + 4 * g);
  }
}

// The residual of the FIRST chunk must already be in rr (epi_load_res before the accumulator barrier).
__device__ __forceinline__ void epilogue_rows_u(const ConvKernelParams &p, const EpiSmem &es, const EpiPixelU &px, uint32_t taddr,
                                              int n0, int half, float4 (&rr)[4]) {
  for (int c0 = 16 * half; c0 < p.Npad; c0 += 32) {
    if (n0 + c0 >= p.Cout4) break;           // warp-uniform
    uint32_t raw[16];
    if (p.debug & 8) {   // experiment: no TMEM reads
#pragma unroll
      for (int i = 0; i < 16; ++i) raw[i] = 0;
    } else {
      tmem_ld16_issue(taddr + (uint32_t)c0, raw);
    }
    float4 rnext[4];
    const bool more = (c0 + 32 < p.Npad) && (n0 + c0 + 32 < p.Cout4);
    if (more) epi_load_res_u(p, px, n0 + c0 + 32, rnext);
    float4 bs[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) bs[g] = *reinterpret_cast<const float4 *>(es.bias + n0 + c0 + 4 * g);
    tmem_ld_wait(raw);
    if (px.inside) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c = n0 + c0 + 4 * g;
        if (c >= p.Cout4) break;
        float4 a = make_float4(__uint_as_float(raw[4 * g]), __uint_as_float(raw[4 * g + 1]), __uint_as_float(raw[4 * g + 2]),
                               __uint_as_float(raw[4 * g + 3]));
        a.x += bs[g].x; a.y += bs[g].y; a.z += bs[g].z; a.w += bs[g].w;
        if (p.pc_ratio) {
          // output = ((raw_out - bias) * mask_ratio + bias) * update_mask, utils/partial_conv.py:74-77, same operation order
          a.x = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(a.x, bs[g].x), px.ratio), bs[g].x), px.um);
          a.y = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(a.y, bs[g].y), px.ratio), bs[g].y), px.um);
          a.z = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(a.z, bs[g].z), px.ratio), bs[g].z), px.um);
          a.w = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(a.w, bs[g].w), px.ratio), bs[g].w), px.um);
        }
        a.x += rr[g].x; a.y += rr[g].y; a.z += rr[g].z; a.w += rr[g].w;
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
          if (o >= p.n_out) break;
          float4 w = a;
          if (p.out[o].slope) {
            const float4 sl = *reinterpret_cast<const float4 *>(es.bias + (1 + o) * es.cpad + c);
            w.x = prelu1(w.x, sl.x); w.y = prelu1(w.y, sl.y); w.z = prelu1(w.z, sl.z); w.w = prelu1(w.w, sl.w);
          }
          if (p.out[o].mul) {
            w.x *= px.mul[o]; w.y *= px.mul[o]; w.z *= px.mul[o]; w.w *= px.mul[o];
          }
          if (p.out[o].f16) {
            if (!(p.debug & 1)) store_half4(px.dst[o], c, w);
            continue;
          }
          if (p.out[o].round_tf32) {
            w.x = round_tf32(w.x); w.y = round_tf32(w.y); w.z = round_tf32(w.z); w.w = round_tf32(w.w);
          }
          if (!(p.debug & 1)) *reinterpret_cast<float4 *>(px.dst[o] + c) = w;
        }
      }
    }
    if (more) {
#pragma unroll
      for (int g = 0; g < 4; ++g) rr[g] = rnext[g];
    }
  }
}


// two CTAs per SM: the per-tap kernel is not persistent, one CTA's epilogue overlaps the other's main loop
template <bool F16>
__global__ void __launch_bounds__(kConvThreads, 2) k_conv_tf32(const __grid_constant__ CUtensorMap map_a,
                                                            const __grid_constant__ CUtensorMap map_b,
                                                            const ConvKernelParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB | B Npad*128] then barriers
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.Npad * kChunk * 4;
  const int stage_bytes = kABytes + b_bytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
  uint64_t *empty = full + p.stages;
  uint64_t *acc_full = empty + p.stages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);
  float *epi_tab = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile -> (image, y0, x0)
  int t = blockIdx.x;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int img = t / p.tiles_y;
  const int x0 = tx * p.tile_w, y0 = ty * p.tile_h;
  const int n0 = blockIdx.y * p.Npad;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int taps = p.ksize * p.ksize;
  const int J = taps * p.chunks;

  // Producer and MMA issuer are whole warps running their loops convergently, with one elected lane issuing the TMA /
  // tcgen05 instructions: ring state and descriptors then live in uniform registers (see k_conv_halo_tf32).
  if (warp == 0) {
    // ===== TMA producer =====
    const bool leader = elect_one();
    uint32_t s = 0, ph = 0;
    int j = 0;
    const int cx = x0 * p.stride - p.pad, cy = y0 * p.stride - p.pad;
    for (int r = 0; r < p.ksize; ++r)
      for (int q = 0; q < p.ksize; ++q)
        for (int ck = 0; ck < p.chunks; ++ck, ++j) {
          mbar_wait(empty + s, ph ^ 1u);
          uint8_t *a_dst = smem + (size_t)s * stage_bytes;
          if (leader) {
            mbar_expect_tx(full + s, (uint32_t)stage_bytes);
            tma_load_4d(&map_a, full + s, a_dst, ck * p.cpc, cx + q, cy + r, img);
            tma_load_3d(&map_b, full + s, a_dst + kABytes, 0, n0, j);
          }
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
        }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const bool leader = elect_one();
    const uint32_t fmt = F16 ? 0u : 2u;        // operand format field: 0 = f16, 2 = tf32
    const uint32_t idesc = (1u << 4) /* D fp32 */ | (fmt << 7) /* A */ | (fmt << 10) /* B */ |
                           ((uint32_t)(p.Npad >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const uint64_t tmpl = umma_desc_sw128(0);
    const uint32_t d_hi = (uint32_t)(tmpl >> 32), d_tl = (uint32_t)tmpl;
    const uint32_t base_lo = smem_u32(smem) >> 4, stage_lo = (uint32_t)stage_bytes >> 4;
    uint32_t s = 0, ph = 0, lo = base_lo;
    int ck = 0;
    for (int j = 0; j < J; ++j) {
      mbar_wait(full + s, ph);
      tc_fence_after();
      const int ksteps = ck == p.chunks - 1 ? p.last_ksteps : kChunk / 8;   // skip MMA steps made of padding channels only
      if (++ck == p.chunks) ck = 0;
      if (leader) {
#pragma unroll
        for (int k = 0; k < kChunk / 8; ++k)   // UMMA K = 8 tf32 = 32 bytes: advance the start address inside the swizzle row
          if (k < ksteps)
            umma_lh<F16>(tmem_base, d_tl | ((lo + 2 * k) & 0x3FFF), d_hi, d_tl | ((lo + (kABytes >> 4) + 2 * k) & 0x3FFF), d_hi, idesc,
                    (uint32_t)((j | k) != 0));
        umma_commit(empty + s);                // slot reusable once these MMAs have read it
      }
      lo += stage_lo;
      if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; lo = base_lo; }
    }
    if (leader) umma_commit(acc_full);         // accumulator complete
  } else {
    // ===== epilogue: warps 2..9; warp w may only touch TMEM lanes 32*(w%4) .. +31 =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int m = q * 32 + lane;               // accumulator row = pixel of the tile
    const int py = m / p.tile_w, px = m - py * p.tile_w;
    const EpiSmem es = epi_stage(p, epi_tab, p.cpad, threadIdx.x - 64);
    const EpiPixelU ep = epi_pixel_u(p, img, y0 + py, x0 + px);
    float4 rr[4];
    epi_load_res_u(p, ep, n0 + 16 * half, rr);
    mbar_wait(acc_full, 0);
    tc_fence_after();
    epilogue_rows_u(p, es, ep, tmem_base + ((uint32_t)(q * 32) << 16), n0, half, rr);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}


// ---- persistent kernel for stride-1 filters: one halo load per tile instead of one load per tap -----------------------
// The A operand of tap (r,s) is the SAME shared-memory halo tile {32 ch, pitch, tile_h+k-1} read through a descriptor whose
// start address is moved by (r*pitch + s) pixels (128 bytes each): 8 consecutive output pixels of one image row are 8
// consecutive 128-byte rows of the swizzle atom, and consecutive image rows are `pitch` pixels apart (the descriptor's
// stride byte offset).  This cuts the L2 -> shared-memory traffic of a 3x3 filter from 9 tiles to ~1.4 tiles per output
// tile.  CTAs are persistent: the TMA producer runs ahead across tiles, the accumulator is double-buffered in TMEM so the
// epilogue of tile i overlaps the MMAs of tile i+1, and filters small enough stay resident in shared memory.
constexpr int kHaloTileW = 8, kHaloTileH = 16;
constexpr int kAccMax = 4;   // accumulator ring in TMEM: up to 4 tiles between the MMA issuer and the epilogue

// Work item w = ((img * tiles_y + ty) * tiles_x + tx) * n_blocks + nb, visited as w = blockIdx.x, += gridDim.x, ...
// The coordinates are carried along as mixed-radix digits: the 64-bit divisions of a direct decomposition cost each of
// the eight epilogue warps several hundred instructions per tile, which made THEM the bottleneck of small tiles.
struct TileIter {
  int nb, tx, ty, img;
  int d_nb, d_tx, d_ty, d_img;
  int left;
  // visits w = blockIdx.x + first * gridDim.x, += every * gridDim.x, ...
  __device__ __forceinline__ void init(const ConvKernelParams &p, int first = 0, int every = 1) {
    long w = blockIdx.x + (long)first * gridDim.x, step = (long)every * gridDim.x;
    left = w < p.work_items ? (int)((p.work_items - w + step - 1) / step) : 0;
    nb = (int)(w % p.n_blocks); w /= p.n_blocks;
    tx = (int)(w % p.tiles_x); w /= p.tiles_x;
    ty = (int)(w % p.tiles_y); img = (int)(w / p.tiles_y);
    d_nb = (int)(step % p.n_blocks); step /= p.n_blocks;
    d_tx = (int)(step % p.tiles_x); step /= p.tiles_x;
    d_ty = (int)(step % p.tiles_y); d_img = (int)(step / p.tiles_y);
  }
  __device__ __forceinline__ void next(const ConvKernelParams &p) {
    --left;
    nb += d_nb;
    int c = nb >= p.n_blocks;
    nb -= c ? p.n_blocks : 0;
    tx += d_tx + c;
    c = tx >= p.tiles_x;
    tx -= c ? p.tiles_x : 0;
    ty += d_ty + c;
    c = ty >= p.tiles_y;
    ty -= c ? p.tiles_y : 0;
    img += d_img + c;
  }
};

// Measured on B200 (tools/probe_halo.py, profiles/probe_halo_r01.jsonl): the tensor core applies the 128-byte swizzle XOR
// to the ABSOLUTE shared-memory address bits [7,10), exactly like TMA does when it writes the tile.  A descriptor may
// therefore start at any 128-byte row of a TMA-written tile and use any multiple of 128 bytes as its stride between
// 8-row groups; the "matrix base offset" field must stay 0 (setting it to (addr >> 7) & 7 corrupts the result).
__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// All taps of one 128-byte channel chunk, NK K-steps each (NK * 8 tf32 or NK * 16 fp16 channels hold real data), issued by one
// thread.  A0 / B0: low descriptor words of the halo slice and of the chunk's first filter panel (panels of consecutive taps
// are b_step_lo apart); acc0 = 0 zeroes the accumulator with the very first MMA.
template <int KS, bool F16, int NK>
__device__ __forceinline__ void issue_taps(uint32_t d_tmem, uint32_t A0, uint32_t a_hi, uint32_t B0, uint32_t b_hi, uint32_t idesc,
                                           uint32_t acc0, uint32_t b_step_lo) {
  constexpr int kPitch = kHaloTileW + KS - 1;
#pragma unroll
  for (int tap = 0; tap < KS * KS; ++tap) {
    const uint32_t a_off = (uint32_t)(((tap / KS) * kPitch + (tap % KS)) * kChunk * 4) >> 4;   // compile-time
#pragma unroll
    for (int k = 0; k < NK; ++k)
      umma_lh<F16>(d_tmem, A0 + a_off + 2 * k, a_hi, B0 + 2 * k, b_hi, idesc, (tap | k) ? 1u : acc0);
    B0 += b_step_lo;
  }
}

// EPI 1 / 2: the lean epilogue in teams of four warps (96 registers, up to 608 threads), 2 = with partial-convolution renormalisation
// and per-pixel output factors; EPI 0: the generic epilogue, 8 warps on one tile.  Instantiations rather than run-time branches:
// together the epilogues need 121 registers, which caps the block at 512 threads.
template <int KS, bool F16, int EPI>
__global__ void __launch_bounds__(EPI ? kHaloLeanThreads : kHaloThreads, 1) k_conv_halo_tf32(const __grid_constant__ CUtensorMap map_a,
                                                                    const __grid_constant__ CUtensorMap map_b,
                                                                    const ConvKernelParams p) {
  constexpr int kTaps = KS * KS;
  constexpr int kPitch = kHaloTileW + KS - 1;                         // halo row pitch in pixels
  constexpr int kBoxBytes = kPitch * (kHaloTileH + KS - 1) * kChunk * 4;   // expect-tx of one halo slice
  constexpr int kAStage = (kBoxBytes + 1023) & ~1023;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.Npad * kChunk * 4;
  uint8_t *smem_a = smem;
  uint8_t *smem_b = smem + (size_t)p.a_stages * kAStage;
  uint64_t *a_full = reinterpret_cast<uint64_t *>(smem_b + (size_t)p.b_stages * b_bytes);
  uint64_t *a_empty = a_full + p.a_stages;
  uint64_t *b_full = a_empty + p.a_stages;
  uint64_t *b_empty = b_full + p.b_stages;
  uint64_t *acc_full = b_empty + p.b_stages;
  uint64_t *acc_empty = acc_full + kAccMax;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + kAccMax);
  float *epi_tab = reinterpret_cast<float *>(smem + p.epi_off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
    for (int s = 0; s < kAccMax; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, EPI ? 4 : 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (whole warp, one elected lane issues) =====
    const bool leader = elect_one();
    uint32_t sb = 0, phb = 0;
    // activation ring(s): one per MMA issuer (ring r = slots r * ring .. + ring - 1), items alternate between them.  Every barrier
    // then has one producer and one consumer that visit all of its phases in order; with a shared ring an issuer would wait for
    // a phase two ahead of the last one it saw, which a parity wait cannot tell from "already complete".
    const uint32_t ring = (uint32_t)(p.a_stages / p.issuers);
    uint32_t ra[2] = {0, 0}, rph[2] = {0, 0};
    uint32_t par = 0;
    bool first = true;
    TileIter it;
    it.init(p);
    if (p.resident && it.left > 0) {
      // Resident filters: ALL panels first, in slot order (chunk-major).  Interleaved with the activation chunks (as the streamed
      // path does) the producer would wait for a free activation stage while the MMA warp waits for the last panels whenever a
      // tile has more chunks than there are activation stages -- e.g. 144 fp16 channels = 3 chunks with 2 stages left next to
      // 27 resident panels.
      const int cn = it.nb * p.Npad;
      for (int ck = 0; ck < p.chunks; ++ck) {
        int j = ck;
        for (int tap = 0; tap < kTaps; ++tap, j += p.chunks) {
          mbar_wait(b_empty + sb, phb ^ 1u);
          if (leader) {
            mbar_expect_tx(b_full + sb, (uint32_t)b_bytes);
            tma_load_3d(&map_b, b_full + sb, smem_b + (size_t)sb * b_bytes, 0, cn, j);
          }
          if (++sb == (uint32_t)p.b_stages) { sb = 0; phb ^= 1u; }
        }
      }
      first = false;
    }
    for (; it.left > 0; it.next(p)) {
      const int img = it.img;
      const int cx = it.tx * kHaloTileW - p.pad, cy = it.ty * kHaloTileH - p.pad, cn = it.nb * p.Npad;
      for (int ck = 0; ck < p.chunks; ++ck) {
        const uint32_t pos = par ? ra[1] : ra[0], pha = par ? rph[1] : rph[0];
        const uint32_t sa = par * ring + pos;
        mbar_wait(a_empty + sa, pha ^ 1u);
        if (leader) {
          if (p.debug & 4) {   // experiment: no activation loads, barriers only
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(a_full + sa)) : "memory");
          } else {
            mbar_expect_tx(a_full + sa, (uint32_t)kBoxBytes);
            tma_load_4d(&map_a, a_full + sa, smem_a + (size_t)sa * kAStage, ck * p.cpc, cx, cy, img);
          }
        }
        const bool wrap = pos + 1 == ring;
        if (par) { ra[1] = wrap ? 0 : pos + 1; rph[1] ^= wrap ? 1u : 0u; }
        else { ra[0] = wrap ? 0 : pos + 1; rph[0] ^= wrap ? 1u : 0u; }
        if (!p.resident || first) {
          int j = ck;                                   // packed weights: panel (tap, ck) at index tap*chunks + ck
          for (int tap = 0; tap < kTaps; ++tap, j += p.chunks) {
            mbar_wait(b_empty + sb, phb ^ 1u);
            if (leader) {
              mbar_expect_tx(b_full + sb, (uint32_t)b_bytes);
              tma_load_3d(&map_b, b_full + sb, smem_b + (size_t)sb * b_bytes, 0, cn, j);
            }
            if (++sb == (uint32_t)p.b_stages) { sb = 0; phb ^= 1u; }
          }
        }
      }
      first = false;
      if (p.issuers == 2) par ^= 1u;
    }
  } else if (warp < kHaloEpiWarp0) {
    // ===== MMA issuers =====
    // The WHOLE warp runs this loop and one elected lane issues the tcgen05 instructions.  With the loop inside
    // `if (lane == 0)` the compiler keeps the ring state in per-thread registers and moves every descriptor into uniform
    // registers one MMA at a time: ncu showed ~17 dependent instructions (~128 cycles) per MMA on the issuing thread, i.e.
    // the tensor pipe (16-64 cycles per MMA) waited for its own issue loop and every other role waited for the tensor pipe.
    const bool leader = elect_one();
    const uint32_t fmt = F16 ? 0u : 2u;        // operand format field: 0 = f16, 2 = tf32
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const uint64_t tmpl_a = umma_desc_sw128_sbo(0, kPitch * kChunk * 4), tmpl_b = umma_desc_sw128(0);
    const uint32_t a_hi = (uint32_t)(tmpl_a >> 32), b_hi = (uint32_t)(tmpl_b >> 32);
    const uint32_t a_tl = (uint32_t)tmpl_a, b_tl = (uint32_t)tmpl_b;          // low words without the address field
    const uint32_t a_base_lo = smem_u32(smem_a) >> 4, b_base_lo = smem_u32(smem_b) >> 4, b_step_lo = (uint32_t)b_bytes >> 4;
    uint32_t sa = 0, pha = 0, a_lo = a_base_lo;                 // (a position inside the issuer's own ring when filters are resident)
    uint32_t sb = 0, phb = 0, b_lo = b_base_lo;
    uint32_t as = 0, phacc = 0;
    const int n_items = blockIdx.x < p.work_items ? (int)((p.work_items - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    // Shared-memory addresses fit the 14-bit (>> 4) address field, whose bits are zero in the templates: "template | address"
    // is "template + address", and the step from one MMA to the next is a compile-time constant added to a per-slot base.
    if (p.resident) {
      // Filters stay in shared memory: wait once for all panels, then the loop has no weight bookkeeping at all.
      // TWO issuing warps, items alternating between them.  The tensor core's queue is short: while a single issuer ran its
      // per-tile bookkeeping (two barrier waits, fences, two commits, ring arithmetic: ~500 cycles of dependent latency) the
      // queue drained, and a 32 -> 32 fp16 tile took 1 650 cycles for 18 MMAs the pipe finishes in 860
      // (tools/micro/umma_rate.cu, 2 K-steps per tap).  With two issuers one's bookkeeping overlaps the other's MMAs; the tiles
      // are independent (own accumulator stages, own activation ring, own epilogue team) and a commit tracks the MMAs of its own
      // thread only.  MMAs of different issuers may complete out of order, which is why nothing is shared between the two chains.
      const int issuer = warp - 1;
      if (issuer < p.issuers) {
      if (n_items > 0)
        for (int j = 0; j < p.b_stages; ++j) mbar_wait(b_full + j, 0);
      tc_fence_after();
      const uint32_t ring = (uint32_t)(p.a_stages / p.issuers), slot0 = (uint32_t)issuer * ring;
      const uint32_t a_step_lo = kAStage >> 4;
      as = (uint32_t)issuer;                                  // accumulator stages issuer, issuer + issuers, ... (acc_stages is even)
      for (int item = issuer; item < n_items; item += p.issuers) {
        mbar_wait(acc_empty + as, phacc ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)p.Npad;
        uint32_t B0 = b_tl + b_base_lo;                       // panel (ck, tap) at slot ck * taps + tap
        for (int ck = 0; ck < p.chunks; ++ck) {
          mbar_wait(a_full + slot0 + sa, pha);
          tc_fence_after();
          const uint32_t A0 = a_tl + a_base_lo + (slot0 + sa) * a_step_lo;
          const int ksteps = ck == p.chunks - 1 ? p.last_ksteps : kChunk / 8;   // skip MMA steps made of padding channels only
          // The issue loop is straight-line code per K-step count: with `if (k < ksteps)` inside one unrolled body the
          // descriptor arithmetic of the absent steps is still executed (predicated off), and a 32-channel fp16 layer -- two
          // steps of four -- spent ~110 cycles of issue per MMA (profiles/ncu_conv_r02w_f16_0).
          if (leader && !(p.debug & 2)) {
            const uint32_t acc0 = (uint32_t)(ck != 0);
            if (ksteps == 4) issue_taps<KS, F16, 4>(d_tmem, A0, a_hi, B0, b_hi, idesc, acc0, b_step_lo);
            else if (ksteps == 2) issue_taps<KS, F16, 2>(d_tmem, A0, a_hi, B0, b_hi, idesc, acc0, b_step_lo);
            else if (ksteps == 1) issue_taps<KS, F16, 1>(d_tmem, A0, a_hi, B0, b_hi, idesc, acc0, b_step_lo);
            else issue_taps<KS, F16, 3>(d_tmem, A0, a_hi, B0, b_hi, idesc, acc0, b_step_lo);
          }
          B0 += kTaps * b_step_lo;
          if (leader) umma_commit(a_empty + slot0 + sa);
          if (++sa == ring) { sa = 0; pha ^= 1u; }
        }
        if (leader) umma_commit(acc_full + as);
        as += (uint32_t)p.issuers;
        if (as >= (uint32_t)p.acc_stages) { as -= p.acc_stages; phacc ^= 1u; }
      }
      }
    } else if (warp == 1) {
      for (int item = 0; item < n_items; ++item) {
        mbar_wait(acc_empty + as, phacc ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)p.Npad;
        for (int ck = 0; ck < p.chunks; ++ck) {
          mbar_wait(a_full + sa, pha);
          tc_fence_after();
          const uint32_t A0 = a_tl + a_lo;
#pragma unroll
          for (int tap = 0; tap < kTaps; ++tap) {
            mbar_wait(b_full + sb, phb);
            tc_fence_after();
            const uint32_t a_off = (uint32_t)(((tap / KS) * kPitch + (tap % KS)) * kChunk * 4) >> 4;   // compile-time
            if (leader) {
              const uint32_t B0 = b_tl + b_lo;
              if (!(p.debug & 2)) {
#pragma unroll
                for (int k = 0; k < kChunk / 8; ++k)   // (streamed filters = wide layers: no padding steps worth skipping)
                    umma_lh<F16>(d_tmem, A0 + a_off + 2 * k, a_hi, B0 + 2 * k, b_hi, idesc, (uint32_t)((ck != 0) | (tap != 0) | (k != 0)));
              }
              umma_commit(b_empty + sb);
            }
            b_lo += b_step_lo;
            if (++sb == (uint32_t)p.b_stages) { sb = 0; phb ^= 1u; b_lo = b_base_lo; }
          }
          if (leader) umma_commit(a_empty + sa);
          a_lo += kAStage >> 4;
          if (++sa == (uint32_t)p.a_stages) { sa = 0; pha ^= 1u; a_lo = a_base_lo; }
        }
        if (leader) umma_commit(acc_full + as);
        if (++as == (uint32_t)p.acc_stages) { as = 0; phacc ^= 1u; }
      }
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3, half = (warp - kHaloEpiWarp0) >> 2;
    const int m = q * 32 + lane;
    const int py = m / kHaloTileW, px = m - py * kHaloTileW;
    const EpiSmem es = epi_stage(p, epi_tab, p.cpad, threadIdx.x - 32 * kHaloEpiWarp0, (int)blockDim.x - 32 * kHaloEpiWarp0);
    uint32_t as = 0, phacc = 0;
    TileIter it;
    if constexpr (EPI != 0) {
      // Teams of four warps (one per TMEM lane quadrant), team t on tiles t, t + teams, ...: the per-tile bookkeeping of a warp
      // (tile coordinates, pixel index, barrier wait) is paid once for all the chunks of its rows, and `teams` tiles are in
      // flight, which is what hides the latency of each warp's dependent chain (two or three warps per scheduler).
      const int team = (warp - kHaloEpiWarp0) >> 2, teams = pin_reg(p.teams);
      const uint32_t tab = smem_u32(epi_tab);
      const int Npad = pin_reg(p.Npad), Cout = pin_reg(p.Cout), n_out = pin_reg(p.n_out), cpad = pin_reg(p.cpad);
      const int Ho = pin_reg(p.Ho), Wo = pin_reg(p.Wo), acc_stages = pin_reg(p.acc_stages);
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
      as = (uint32_t)team;
      while (as >= (uint32_t)acc_stages) { as -= acc_stages; phacc ^= 1u; }
      for (it.init(p, team, teams); it.left > 0; it.next(p)) {
        const int n0 = it.nb * Npad, oy = it.ty * kHaloTileH + py, ox = it.tx * kHaloTileW + px;
        const bool inside = (oy < Ho) & (ox < Wo);
        const long pix = inside ? ((long)it.img * Ho + oy) * Wo + ox : 0;
        float4 rr[4];
        if (p.res && inside) lean_load_res(p.res + pix * p.res_stride + n0, p.res_wide != 0, rr);   // in flight while the MMAs finish
        float ratio = 1.f, um = 1.f;                           // partial convolution: per-pixel mask_ratio and update_mask
        if (EPI == 2 && p.pc_ratio) { ratio = __ldg(p.pc_ratio + pix); um = __ldg(p.pc_um + pix); }
        mbar_wait(acc_full + as, phacc);
        tc_fence_after();
        epilogue_chunks_lean<EPI == 2>(p, tab, cpad, n_out, Npad, Cout, inside, pix, lane_base + as * (uint32_t)Npad, n0, rr, ratio, um);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc_empty + as)) : "memory");
        as += (uint32_t)teams;
        while (as >= (uint32_t)acc_stages) { as -= acc_stages; phacc ^= 1u; }
      }
    } else {
      for (it.init(p); it.left > 0; it.next(p)) {
        const int n0 = it.nb * p.Npad;
        const EpiPixel ep = epi_pixel(p, it.img, it.ty * kHaloTileH + py, it.tx * kHaloTileW + px);
        float4 rr[4];
        epi_load_res(p, ep, n0 + 16 * half, rr);                 // in flight while the MMAs of this tile finish
        mbar_wait(acc_full + as, phacc);
        tc_fence_after();
        epilogue_rows(p, es, ep, tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)p.Npad, n0, half, rr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc_empty + as)) : "memory");
        if (++as == (uint32_t)p.acc_stages) { as = 0; phacc ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---- weight packing: OIHW fp32 -> [tap][chunk][Cout_pad16][32] fp32 rounded to TF32 ------------------------
__global__ void __launch_bounds__(256) k_pack_weights(const float *__restrict__ w, int Cout, int Cin, int ksize, int chunks,
                                                      int Cout_pad, const float *__restrict__ out_scale, float *__restrict__ dst,
                                                      long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ci = (int)(i % kChunk);
  long r = i / kChunk;
  const int o = (int)(r % Cout_pad);
  r /= Cout_pad;
  const int ck = (int)(r % chunks);
  const int tap = (int)(r / chunks);
  const int c = ck * kChunk + ci;
  float v = 0.f;
  if (o < Cout && c < Cin) {
    v = w[((long)o * Cin + c) * ksize * ksize + tap];
    if (out_scale) v *= out_scale[o];
  }
  dst[i] = round_tf32(v);
}

// The same filters for kind::f16: [tap][chunk of 64 channels][Cout_pad16][64] halves (one 128-byte row per output channel and chunk).
__global__ void __launch_bounds__(256) k_pack_weights_f16(const float *__restrict__ w, int Cout, int Cin, int ksize, int chunks,
                                                          int Cout_pad, const float *__restrict__ out_scale, __half *__restrict__ dst,
                                                          long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cpc = 2 * kChunk;
  const int ci = (int)(i % cpc);
  long r = i / cpc;
  const int o = (int)(r % Cout_pad);
  r /= Cout_pad;
  const int ck = (int)(r % chunks);
  const int tap = (int)(r / chunks);
  const int c = ck * cpc + ci;
  float v = 0.f;
  if (o < Cout && c < Cin) {
    v = w[((long)o * Cin + c) * ksize * ksize + tap];
    if (out_scale) v *= out_scale[o];
  }
  dst[i] = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
}

// ---- elementwise companions of the conv stacks (NHWC) ---------------------------------------------------------
// bilinear x2 (align_corners=False) followed by the per-channel PReLU of the Upsample block
// (models/pointcloud_inpainting.py:70-72); output optionally cropped to (Ho, Wo) <= (2H, 2W).
__global__ void __launch_bounds__(256) k_upsample2x_prelu(const float *__restrict__ x, long xs, int H, int W, int C4,
                                                          const float *__restrict__ slope, int C, float *__restrict__ y, long ys,
                                                          int Ho, int Wo, int round, const float *__restrict__ mul) {
  // One thread per (input cell, 4-channel group): the cell between input pixels (ix, ix + 1) x (iy, iy + 1), ix = -1 .. W - 1,
  // holds the 2 x 2 output pixels (2 ix + 1, 2 ix + 2) x (2 iy + 1, 2 iy + 2), all interpolated from those four inputs:
  //   src = (dst + 0.5) / 2 - 0.5, clamped at 0 (PyTorch area_pixel_compute_source_index, align_corners=False)
  //   dst = 2 i + 1 -> src = i + 0.25;  dst = 2 i + 2 -> src = i + 0.75;  dst = 0 -> src = 0 (cell -1, weight 1 on pixel 0).
  // Four 16-byte loads feed four outputs (the one-output-per-thread form loaded four per output and was bound by its own
  // instruction stream).  grid: x over (cell column, channel group), y over (image, cell row): 32-bit index arithmetic only.
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (unsigned)(W + 1) * (unsigned)C4) return;
  const int cx = (int)(t / (unsigned)C4), cg = (int)(t - (unsigned)cx * (unsigned)C4);
  const int n = (int)(blockIdx.y / (unsigned)(H + 1)), cy = (int)(blockIdx.y - (unsigned)n * (unsigned)(H + 1));
  const int ix = cx - 1, iy = cy - 1;
  const int x0 = max(ix, 0), x1 = min(ix + 1, W - 1), y0 = max(iy, 0), y1 = min(iy + 1, H - 1);
  const float *b = x + (long)n * H * W * xs + 4 * cg;
  const float4 p00 = *reinterpret_cast<const float4 *>(b + ((long)y0 * W + x0) * xs);
  const float4 p01 = *reinterpret_cast<const float4 *>(b + ((long)y0 * W + x1) * xs);
  const float4 p10 = *reinterpret_cast<const float4 *>(b + ((long)y1 * W + x0) * xs);
  const float4 p11 = *reinterpret_cast<const float4 *>(b + ((long)y1 * W + x1) * xs);
  const float a00[4] = {p00.x, p00.y, p00.z, p00.w}, a01[4] = {p01.x, p01.y, p01.z, p01.w};
  const float a10[4] = {p10.x, p10.y, p10.z, p10.w}, a11[4] = {p11.x, p11.y, p11.z, p11.w};
  float sl[4] = {1.f, 1.f, 1.f, 1.f};
  if (slope) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (4 * cg + e < C) sl[e] = __ldg(slope + 4 * cg + e);
  }
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    const int oy = 2 * iy + 1 + dy;
    if (oy < 0 || oy >= Ho) continue;
    // weight of the lower row: 0.25 / 0.75, and 0 for output row 0 (clamped source)
    const float ly = dy == 0 ? 0.25f : (iy < 0 ? 0.f : 0.75f), hy = 1.f - ly;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const int ox = 2 * ix + 1 + dx;
      if (ox < 0 || ox >= Wo) continue;
      const float lx = dx == 0 ? 0.25f : (ix < 0 ? 0.f : 0.75f), hx = 1.f - lx;
      const long opix = ((long)n * Ho + oy) * Wo + ox;
      const float m = mul ? __ldg(mul + opix) : 1.f;
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // same association as ATen's upsample_bilinear2d: hy*(hx*p00 + lx*p01) + ly*(hx*p10 + lx*p11)
        float v = hy * (hx * a00[e] + lx * a01[e]) + ly * (hx * a10[e] + lx * a11[e]);
        const int c = 4 * cg + e;
        if (slope && c < C) v = v > 0.f ? v : v * sl[e];
        if (c >= C) v = 0.f;
        if (mul) v *= m;
        o[e] = round == 1 ? round_tf32(v) : v;
      }
      if (round == 2) store_half4(y, opix * ys + 4 * cg, make_float4(o[0], o[1], o[2], o[3]));   // fp16 output
      else *reinterpret_cast<float4 *>(y + opix * ys + 4 * cg) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// y = prelu(x) (slope per channel; nullptr = copy), NHWC with independent pixel strides.
__global__ void __launch_bounds__(256) k_prelu_nhwc(const float *__restrict__ x, long xs, const float *__restrict__ slope, int C,
                                                    int C4, float *__restrict__ y, long ys, int round, long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cg = (int)(i % C4);
  const long pix = i / C4;
  const float4 v = *reinterpret_cast<const float4 *>(x + pix * xs + 4 * cg);
  float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = 4 * cg + e;
    if (slope && c < C) a[e] = a[e] > 0.f ? a[e] : a[e] * __ldg(slope + c);
    if (c >= C) a[e] = 0.f;
    if (round == 1) a[e] = round_tf32(a[e]);
  }
  if (round == 2) store_half4(y, pix * ys + 4 * cg, make_float4(a[0], a[1], a[2], a[3]));                                 // fp16 output
  else *reinterpret_cast<float4 *>(y + pix * ys + 4 * cg) = make_float4(a[0], a[1], a[2], a[3]);
}

// 2x2 max-pool, stride 2, ceil_mode=True (models/disparity_estimation.py:90), NHWC.
__global__ void __launch_bounds__(256) k_maxpool2_ceil(const float *__restrict__ x, long xs, int H, int W, int C4,
                                                       float *__restrict__ y, long ys, int Ho, int Wo, long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cg = (int)(i % C4);
  long r = i / C4;
  const int ox = (int)(r % Wo);
  r /= Wo;
  const int oy = (int)(r % Ho);
  const int n = (int)(r / Ho);
  const float *b = x + (long)n * H * W * xs + 4 * cg;
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int dy = 0; dy < 2; ++dy)
    for (int dx = 0; dx < 2; ++dx) {
      const int iy = 2 * oy + dy, ix = 2 * ox + dx;
      if (iy >= H || ix >= W) continue;
      const float4 v = *reinterpret_cast<const float4 *>(b + ((long)iy * W + ix) * xs);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  *reinterpret_cast<float4 *>(y + (((long)n * Ho + oy) * Wo + ox) * ys + 4 * cg) = m;
}

// NCHW [N,C,H,W] -> NHWC with pixel stride ys (>= C, channels beyond C zeroed up to C4*4), optional affine
// per-tensor (x - sub) * mul used for the per-sample normalisation of Refine / Inpaint.
__global__ void __launch_bounds__(256) k_nchw_to_nhwc(const float *__restrict__ x, int C, long HW, float *__restrict__ y, long ys,
                                                      int C4, float sub, float mul, long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long pix = i % HW;          // consecutive threads -> consecutive pixels: coalesced plane reads
  long r = i / HW;
  const int cg = (int)(r % C4);
  const int n = (int)(r / C4);
  float a[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = 4 * cg + e;
    a[e] = c < C ? (x[((long)n * C + c) * HW + pix] - sub) * mul : 0.f;
  }
  *reinterpret_cast<float4 *>(y + ((long)n * HW + pix) * ys + 4 * cg) = make_float4(a[0], a[1], a[2], a[3]);
}

// NHWC (pixel stride xs, channel offset applied by the caller) -> NCHW [N,C,H,W], y = x * mul + add.
__global__ void __launch_bounds__(256) k_nhwc_to_nchw(const float *__restrict__ x, long xs, int C, long HW, float *__restrict__ y,
                                                      float mul, float add, long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long pix = i % HW;
  long r = i / HW;
  const int c = (int)(r % C);
  const int n = (int)(r / C);
  y[i] = x[((long)n * HW + pix) * xs + c] * mul + add;
}


// ---- PartialConv2d mask bookkeeping, utils/partial_conv.py:43-69 --------------------------------------------------------
// With multi_channel=True the reference convolves a [N,Cin,H,W] mask with an all-ones [Cout,Cin,k,k] filter.  Every mask of
// models/partial_inpainting.py has identical channels (it starts as tensorMasks.expand_as(data), :152, and each update
// is again channel-independent), so that convolution is Cin times the k x k box sum of ONE channel -- a sum of 0/1 values,
// exact in fp32 in any order.  Per output pixel: S = Cin * box(mask), update_mask = clamp(S, 0, 1),
// mask_ratio = (Cin*k*k / (S + 1e-8)) * update_mask.  mask == nullptr is the reference's "no mask" case (all ones).
__global__ void __launch_bounds__(256) k_pconv_mask(const float *__restrict__ mask, int H, int W, int Cin, int ksize, int stride,
                                                    int pad, int Ho, int Wo, float *__restrict__ ratio, float *__restrict__ um,
                                                    long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ox = (int)(i % Wo);
  long r = i / Wo;
  const int oy = (int)(r % Ho);
  const int n = (int)(r / Ho);
  const float *m = mask ? mask + (long)n * H * W : nullptr;
  float box = 0.f;
  for (int dy = 0; dy < ksize; ++dy) {
    const int iy = oy * stride - pad + dy;
    if (iy < 0 || iy >= H) continue;
    for (int dx = 0; dx < ksize; ++dx) {
      const int ix = ox * stride - pad + dx;
      if (ix < 0 || ix >= W) continue;
      box += m ? m[(long)iy * W + ix] : 1.f;
    }
  }
  const float S = __fmul_rn(box, (float)Cin);
  const float u = fminf(fmaxf(S, 0.f), 1.f);
  // `self.slide_winsize / (self.update_mask + 1e-8)` is int / Tensor, which torch evaluates as reciprocal(tensor) * int
  ratio[i] = __fmul_rn(__fmul_rn(__frcp_rn(__fadd_rn(S, 1e-8f)), (float)(Cin * ksize * ksize)), u);
  um[i] = u;
}

// ---- host ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

static int pow2_at_least(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace kb

using namespace kb;

extern "C" {

long kb_conv_packed_floats(int Cout, int Cin, int ksize) {
  if (Cout <= 0 || Cin <= 0 || ksize <= 0) return 0;
  const long chunks = (Cin + kChunk - 1) / kChunk, cout_pad = (Cout + 15) / 16 * 16;
  return (long)ksize * ksize * chunks * cout_pad * kChunk;
}

int kb_conv_pack_weights(const float *w_oihw, int Cout, int Cin, int ksize, const float *out_scale, float *packed,
                         kb_stream_t stream) {
  KB_REQUIRE(w_oihw && packed && Cout > 0 && Cin > 0 && ksize > 0, "kb_conv_pack_weights: bad arguments");
  const int chunks = (Cin + kChunk - 1) / kChunk, cout_pad = (Cout + 15) / 16 * 16;
  const long total = kb_conv_packed_floats(Cout, Cin, ksize);
  k_pack_weights<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, ksize, chunks, cout_pad, out_scale,
                                                                    packed, total);
  count_launch();
  return check_launch("kb_conv_pack_weights");
}



long kb_conv_packed_halves(int Cout, int Cin, int ksize) {
  if (Cout <= 0 || Cin <= 0 || ksize <= 0) return 0;
  const long cpc = 2 * kChunk, chunks = (Cin + cpc - 1) / cpc, cout_pad = (Cout + 15) / 16 * 16;
  return (long)ksize * ksize * chunks * cout_pad * cpc;
}

int kb_conv_pack_weights_f16(const float *w_oihw, int Cout, int Cin, int ksize, const float *out_scale, void *packed,
                             kb_stream_t stream) {
  KB_REQUIRE(w_oihw && packed && Cout > 0 && Cin > 0 && ksize > 0, "kb_conv_pack_weights_f16: bad arguments");
  const int cpc = 2 * kChunk, chunks = (Cin + cpc - 1) / cpc, cout_pad = (Cout + 15) / 16 * 16;
  const long total = kb_conv_packed_halves(Cout, Cin, ksize);
  k_pack_weights_f16<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, ksize, chunks, cout_pad, out_scale,
                                                                        reinterpret_cast<__half *>(packed), total);
  count_launch();
  return check_launch("kb_conv_pack_weights_f16");
}

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v ? atoi(v) : dflt;
}

static int make_weight_map(EncodeTiledFn enc, const float *w_packed, int cout_pad, int J, int npad, int f16, CUtensorMap *map) {
  const int cpc = f16 ? 2 * kChunk : kChunk;      // elements per 128-byte row
  cuuint64_t dims[3] = {(cuuint64_t)cpc, (cuuint64_t)cout_pad, (cuuint64_t)J};
  cuuint64_t strides[2] = {(cuuint64_t)kChunk * 4, (cuuint64_t)kChunk * 4 * cout_pad};
  cuuint32_t box[3] = {(cuuint32_t)cpc, (cuuint32_t)npad, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(w_packed), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("kb_conv2d: cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
    return KB_EINVAL;
  }
  return 0;
}

// NHWC activations as a 4-D tensor {C, W, H, N}; box {32, box_w, box_h, 1} traversed with the convolution stride.
static int make_act_map(EncodeTiledFn enc, const kb_conv_args *a, int box_w, int box_h, int stride, CUtensorMap *map) {
  const int esz = a->x_f16 ? 2 : 4, cpc = a->x_f16 ? 2 * kChunk : kChunk;
  cuuint64_t dims[4] = {(cuuint64_t)a->Cin, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->N};
  cuuint64_t strides[3] = {(cuuint64_t)a->x_stride * esz, (cuuint64_t)a->x_stride * esz * a->W,
                           (cuuint64_t)a->x_stride * esz * a->W * a->H};
  cuuint32_t box[4] = {(cuuint32_t)cpc, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  // (L2 promotion NONE / 64B / 128B / 256B measure the same here)
  CUresult r = enc(map, a->x_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(a->x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("kb_conv2d: cuTensorMapEncodeTiled(activations) failed with CUresult %d", (int)r);
    return KB_EINVAL;
  }
  return 0;
}

// cudaFuncSetAttribute and the SM count are per DEVICE: one process may drive several GPUs, so the caches are keyed by it.
static constexpr int kMaxDevices = 64;
static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

static int raise_smem_limit() {
  static bool done_dev[kMaxDevices] = {};
  bool &done = done_dev[current_device()];
  if (done) return 0;
  const int big = 227 * 1024;
  cudaError_t e = cudaFuncSetAttribute(k_conv_tf32<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_tf32<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_halo_tf32<1, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_halo_tf32<1, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_halo_tf32<1, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_halo_tf32<1, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_halo_tf32<1, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_halo_tf32<1, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_halo_tf32<3, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_halo_tf32<3, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_halo_tf32<3, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_halo_tf32<3, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_halo_tf32<3, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_halo_tf32<3, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e != cudaSuccess) {
    set_error("kb_conv2d: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
    return (int)e;
  }
  done = true;
  return 0;
}

static int sm_count() {
  static int n_dev[kMaxDevices] = {};
  const int dev = current_device();
  int &n = n_dev[dev];
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

int kb_conv2d(const kb_conv_args *a, kb_stream_t stream) {
  KB_REQUIRE(a && a->x && a->w_packed, "kb_conv2d: null argument");
  KB_REQUIRE(a->N > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, "kb_conv2d: bad shape");
  KB_REQUIRE(a->x_stride >= a->Cin && a->x_stride % (a->x_f16 ? 8 : 4) == 0,
             "kb_conv2d: x_stride must be a multiple of 16 bytes and >= Cin");
  KB_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & 15) == 0, "kb_conv2d: x must be 16-byte aligned");
  KB_REQUIRE(a->ksize >= 1 && a->ksize <= 7 && (a->stride == 1 || a->stride == 2) && a->pad >= 0, "kb_conv2d: unsupported filter");
  KB_REQUIRE(a->n_out >= 1 && a->n_out <= kMaxOut, "kb_conv2d: need 1..3 outputs");
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("kb_conv2d: cuTensorMapEncodeTiled is not available from the driver");
    return KB_ENOSUP;
  }
  ConvKernelParams p;
  memset(&p, 0, sizeof(p));
  p.Ho = (a->H + 2 * a->pad - a->ksize) / a->stride + 1;
  p.Wo = (a->W + 2 * a->pad - a->ksize) / a->stride + 1;
  KB_REQUIRE(p.Ho > 0 && p.Wo > 0, "kb_conv2d: empty output");
  if (a->out_H > 0) {
    KB_REQUIRE(a->out_H <= p.Ho, "kb_conv2d: out_H exceeds the convolution output");
    p.Ho = a->out_H;
  }
  if (a->out_W > 0) {
    KB_REQUIRE(a->out_W <= p.Wo, "kb_conv2d: out_W exceeds the convolution output");
    p.Wo = a->out_W;
  }
  // algo: 1 = one TMA load per filter tap (any filter), 2 = persistent halo kernel (stride 1, k <= 3)
  const bool halo_ok = a->stride == 1 && (a->ksize == 1 || a->ksize == 3) && a->pad == a->ksize / 2;
  int algo = a->algo > 0 ? a->algo : env_int("KB_CONV_ALGO", 0);
  if (algo == 0) algo = halo_ok ? 2 : 1;
  KB_REQUIRE(algo == 1 || (algo == 2 && halo_ok), "kb_conv2d: algo 2 needs stride 1, ksize <= 3, 'same' padding");

  p.f16 = a->x_f16 ? 1 : 0;
  p.cpc = p.f16 ? 2 * kChunk : kChunk;                 // channels per 128-byte chunk
  const int kstep_ch = p.cpc / 4;                      // channels per MMA K step (32 bytes): 8 tf32 or 16 f16
  p.chunks = (a->Cin + p.cpc - 1) / p.cpc;
  p.last_ksteps = ((a->Cin - (p.chunks - 1) * p.cpc) + kstep_ch - 1) / kstep_ch;
  p.ksize = a->ksize;
  p.stride = a->stride;
  p.pad = a->pad;
  p.Cout = a->Cout;
  p.Cout4 = (a->Cout + 3) & ~3;
  const int cout_pad = (a->Cout + 15) / 16 * 16;
  const int J = a->ksize * a->ksize * p.chunks;
  if (algo == 1) {
    p.tile_w = a->tile_w > 0 ? a->tile_w : (p.Wo >= 32 ? 32 : pow2_at_least(p.Wo));
    KB_REQUIRE(p.tile_w <= kTileM && (p.tile_w & (p.tile_w - 1)) == 0, "kb_conv2d: tile_w must be a power of two <= 128");
    p.tile_h = kTileM / p.tile_w;
  } else {
    p.tile_w = kHaloTileW;
    p.tile_h = kHaloTileH;
  }
  p.tiles_x = (p.Wo + p.tile_w - 1) / p.tile_w;
  p.tiles_y = (p.Ho + p.tile_h - 1) / p.tile_h;
  const long tiles = (long)p.tiles_x * p.tiles_y * a->N;
  int npad = a->n_block > 0 ? a->n_block : cout_pad;
  if (a->n_block <= 0) {
    if (npad > 256) npad = 256;
    // few tiles (deep, low-resolution rows): split the output channels over more CTAs to fill the SMs
    while (npad > 64 && npad % 32 == 0 && tiles * ((cout_pad + npad - 1) / npad) < sm_count()) npad /= 2;
  }
  KB_REQUIRE(npad % 16 == 0 && npad >= 16 && npad <= 256, "kb_conv2d: n_block must be a multiple of 16 in [16,256]");
  p.Npad = npad;
  const int n_blocks = (cout_pad + npad - 1) / npad;
  p.n_blocks = n_blocks;
  p.bias = a->bias;
  p.res = a->res;
  p.res_stride = a->res_stride;
  if (a->res) KB_REQUIRE(a->res_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(a->res) & 15) == 0, "kb_conv2d: residual alignment");
  p.res_wide = a->res && a->res_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(a->res) & 31) == 0 && !env_int("KB_CONV_NO_WIDE", 0);
  p.n_out = a->n_out;
  p.vec = (a->Cout % 4 == 0) && (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0;
  for (int o = 0; o < a->n_out; ++o) {
    if (reinterpret_cast<uintptr_t>(a->out[o].slope) & 15) p.vec = 0;
    KB_REQUIRE(a->out[o].ptr && a->out[o].pixel_stride % (a->out[o].store_f16 ? 8 : 4) == 0 && a->out[o].pixel_stride >= p.Cout4 &&
                   (reinterpret_cast<uintptr_t>(a->out[o].ptr) & 15) == 0,
               "kb_conv2d: output %d: pointer / stride must be 16-byte aligned and hold %d channels", o, p.Cout4);
    p.out[o].ptr = a->out[o].ptr;
    p.out[o].stride = a->out[o].pixel_stride;
    p.out[o].slope = a->out[o].slope;
    p.out[o].mul = a->out[o].mul;
    p.out[o].round_tf32 = a->out[o].round_tf32;
    p.out[o].f16 = a->out[o].store_f16;
    p.out[o].wide = (reinterpret_cast<uintptr_t>(a->out[o].ptr) & 31) == 0 &&
                    (a->out[o].pixel_stride * (a->out[o].store_f16 ? 2 : 4)) % 32 == 0 && !env_int("KB_CONV_NO_WIDE", 0);
  }
  KB_REQUIRE((a->pc_ratio == nullptr) == (a->pc_um == nullptr), "kb_conv2d: pc_ratio and pc_um come together");
  p.pc_ratio = a->pc_ratio;
  p.pc_um = a->pc_um;
  int rc = raise_smem_limit();
  if (rc) return rc;
  CUtensorMap map_a, map_b;
  rc = make_weight_map(enc, a->w_packed, cout_pad, J, npad, p.f16, &map_b);
  if (rc) return rc;
  const int b_bytes = npad * kChunk * 4;
  const size_t smem_cap = 227 * 1024;
  p.cpad = n_blocks * npad;
  const size_t epi_bytes = sizeof(float) * (size_t)epi_smem_floats(p.cpad);

  if (algo == 1) {
    p.tmem_cols = (uint32_t)max(32, pow2_at_least(npad));
    const int stage_bytes = kABytes + b_bytes;
    int stages = a->stages > 0 ? a->stages : (stage_bytes * 4 <= 100 * 1024 ? 4 : (int)((200 * 1024 - epi_bytes) / stage_bytes));
    stages = max(2, min(min(stages, 8), max(J, 2)));
    p.stages = stages;
    rc = make_act_map(enc, a, p.tile_w * a->stride, p.tile_h * a->stride, a->stride, &map_a);
    if (rc) return rc;
    const size_t smem = 1024 + (size_t)stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 32 + epi_bytes;
    KB_REQUIRE(smem <= smem_cap, "kb_conv2d: pipeline does not fit shared memory");
    dim3 grid((unsigned)tiles, (unsigned)n_blocks);
    if (p.f16) k_conv_tf32<true><<<grid, kConvThreads, smem, (cudaStream_t)stream>>>(map_a, map_b, p);
    else k_conv_tf32<false><<<grid, kConvThreads, smem, (cudaStream_t)stream>>>(map_a, map_b, p);
    count_launch();
    return check_launch("kb_conv2d");
  }

  // ---- persistent halo kernel ----
  const int halo_w = kHaloTileW + a->ksize - 1, halo_h = kHaloTileH + a->ksize - 1;
  p.pitch = halo_w;
  p.debug = env_int("KB_CONV_DEBUG", 0);
  p.fast = ((p.debug & ~6) == 0 && a->Cout % 16 == 0 && !env_int("KB_CONV_NO_LEAN", 0)) ? 1 : 0;
  if (p.fast && a->pc_ratio) p.fast = 2;
  for (int o = 0; o < a->n_out; ++o)
    if (p.fast && a->out[o].mul) p.fast = 2;
  const int box_bytes = p.pitch * halo_h * kChunk * 4;
  p.a_stage_bytes = (box_bytes + 1023) & ~1023;
  p.acc_stages = min(kAccMax, 512 / npad);
  KB_REQUIRE(p.acc_stages >= 2, "kb_conv2d: accumulator does not fit TMEM");
  const size_t bar_bytes = 1024;                         // barriers + TMEM slot
  const size_t fixed = 1024 + bar_bytes + epi_bytes;     // alignment slack + barriers + epilogue tables
  const size_t budget = smem_cap - fixed;
  p.resident = (n_blocks == 1 && (size_t)J * b_bytes + 2 * (size_t)p.a_stage_bytes <= budget) ? 1 : 0;
  if (env_int("KB_CONV_NO_RESIDENT", 0)) p.resident = 0;
  if (p.resident) {
    p.b_stages = J;
    p.a_stages = (int)min((size_t)6, (budget - (size_t)J * b_bytes) / p.a_stage_bytes);
  } else {
    p.a_stages = p.chunks >= 3 ? 3 : 2;
    p.b_stages = (int)min((size_t)12, (budget - (size_t)p.a_stages * p.a_stage_bytes) / b_bytes);
    KB_REQUIRE(p.b_stages >= 2, "kb_conv2d: weight ring does not fit shared memory");
  }
  if (a->stages > 0) p.a_stages = max(2, min(a->stages, p.a_stages));
  // Two issuers halve the activation ring each of them sees: worth it only while a half still holds a whole tile plus one slot
  // to load ahead into (measured: 64 -> 64 TF32 -- 2 chunks, 3 slots next to 144 KB of resident filters -- ran 146 us with two
  // issuers on one slot each, 114 us with one issuer on three).
  p.issuers = (p.resident && p.a_stages / 2 >= max(2, p.chunks) && !env_int("KB_CONV_ONE_ISSUER", 0)) ? kHaloIssuers : 1;
  if (p.issuers == 2) {
    // item i belongs to issuer i % 2, accumulator stage i % acc_stages and epilogue team i % 2: even ring sizes keep every
    // barrier between ONE issuer and ONE team (or the producer), each visiting all of its phases in order
    p.acc_stages &= ~1;
    p.a_stages &= ~1;
    p.teams = p.fast ? ((p.acc_stages >= 4 && env_int("KB_CONV_TEAMS", kHaloMaxTeams) >= 4) ? 4 : 2) : 0;
  } else {
    // a team's first wait must be for the first phase of its accumulator stage: teams <= acc_stages
    p.teams = p.fast ? max(1, min(min(env_int("KB_CONV_TEAMS", kHaloMaxTeams), kHaloMaxTeams), p.acc_stages)) : 0;
  }
  p.tmem_cols = (uint32_t)max(32, pow2_at_least(p.acc_stages * npad));
  const int halo_threads = p.fast ? 32 * kHaloEpiWarp0 + 128 * p.teams : kHaloThreads;
  KB_REQUIRE((2 * p.a_stages + 2 * p.b_stages + 2 * kAccMax) * sizeof(uint64_t) + 16 <= bar_bytes, "kb_conv2d: too many pipeline stages");
  p.epi_off = (int)((size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * b_bytes + bar_bytes);
  p.work_items = tiles * n_blocks;
  rc = make_act_map(enc, a, p.pitch, halo_h, 1, &map_a);
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * b_bytes + bar_bytes + epi_bytes;
  KB_REQUIRE(smem <= smem_cap, "kb_conv2d: pipeline does not fit shared memory");
  const unsigned grid = (unsigned)min((long)sm_count(), p.work_items);
  const cudaStream_t st = (cudaStream_t)stream;
#define KB_HALO_LAUNCH(KS_, F16_, EPI_) k_conv_halo_tf32<KS_, F16_, EPI_><<<grid, halo_threads, smem, st>>>(map_a, map_b, p)
#define KB_HALO_EPI(KS_, F16_)                         \
  do {                                                 \
    if (p.fast == 2) KB_HALO_LAUNCH(KS_, F16_, 2);     \
    else if (p.fast == 1) KB_HALO_LAUNCH(KS_, F16_, 1); \
    else KB_HALO_LAUNCH(KS_, F16_, 0);                 \
  } while (0)
  if (a->ksize == 1) {
    if (p.f16) KB_HALO_EPI(1, true); else KB_HALO_EPI(1, false);
  } else {
    if (p.f16) KB_HALO_EPI(3, true); else KB_HALO_EPI(3, false);
  }
#undef KB_HALO_EPI
#undef KB_HALO_LAUNCH
  count_launch();
  return check_launch("kb_conv2d");
}

int kb_upsample2x_prelu(const float *x, long x_stride, int N, int H, int W, int C, const float *slope, float *y, long y_stride,
                        int Ho, int Wo, int round_tf32, const float *mul, kb_stream_t stream) {
  KB_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0 && Ho <= 2 * H && Wo <= 2 * W,
             "kb_upsample2x_prelu: bad arguments");
  KB_REQUIRE(x_stride % 4 == 0 && y_stride % (round_tf32 == 2 ? 8 : 4) == 0 &&
                 ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
             "kb_upsample2x_prelu: pointers must be 16-byte aligned, strides multiples of 16 bytes");
  const int C4 = (C + 3) / 4;
  KB_REQUIRE((long)N * (H + 1) <= 65535 && (long)(W + 1) * C4 < (1L << 31), "kb_upsample2x_prelu: N * (H + 1) must fit a grid dimension");
  const dim3 grid((unsigned)cdiv((long)(W + 1) * C4, 256), (unsigned)(N * (H + 1)));
  k_upsample2x_prelu<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_stride, H, W, C4, slope, C, y, y_stride, Ho, Wo, round_tf32, mul);
  count_launch();
  return check_launch("kb_upsample2x_prelu");
}

int kb_prelu_nhwc(const float *x, long x_stride, long pixels, int C, const float *slope, float *y, long y_stride, int round_tf32,
                  kb_stream_t stream) {
  KB_REQUIRE(x && y && pixels > 0 && C > 0 && x_stride % 4 == 0 && y_stride % (round_tf32 == 2 ? 8 : 4) == 0 &&
                 ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
             "kb_prelu_nhwc: bad arguments (16-byte aligned pointers, strides multiples of 16 bytes)");
  const int C4 = (C + 3) / 4;
  const long total = pixels * C4;
  k_prelu_nhwc<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, x_stride, slope, C, C4, y, y_stride, round_tf32, total);
  count_launch();
  return check_launch("kb_prelu_nhwc");
}

int kb_maxpool2_ceil(const float *x, long x_stride, int N, int H, int W, int C, float *y, long y_stride, kb_stream_t stream) {
  KB_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && x_stride % 4 == 0 && y_stride % 4 == 0 &&
                 ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
             "kb_maxpool2_ceil: bad arguments (16-byte aligned pointers, strides multiples of 4 floats)");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, C4 = (C + 3) / 4;
  const long total = (long)N * Ho * Wo * C4;
  k_maxpool2_ceil<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, x_stride, H, W, C4, y, y_stride, Ho, Wo, total);
  count_launch();
  return check_launch("kb_maxpool2_ceil");
}

int kb_nchw_to_nhwc(const float *x, int N, int C, int H, int W, float *y, long y_stride, float sub, float mul,
                    kb_stream_t stream) {
  KB_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0 && y_stride % 4 == 0 && y_stride >= ((C + 3) & ~3),
             "kb_nchw_to_nhwc: bad arguments");
  KB_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, "kb_nchw_to_nhwc: destination must be 16-byte aligned");
  const int C4 = (C + 3) / 4;
  const long HW = (long)H * W, total = (long)N * C4 * HW;
  k_nchw_to_nhwc<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, C, HW, y, y_stride, C4, sub, mul, total);
  count_launch();
  return check_launch("kb_nchw_to_nhwc");
}

int kb_nhwc_to_nchw(const float *x, long x_stride, int N, int C, int H, int W, float *y, float mul, float add,
                    kb_stream_t stream) {
  KB_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0, "kb_nhwc_to_nchw: bad arguments");
  const long HW = (long)H * W, total = (long)N * C * HW;
  k_nhwc_to_nchw<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, x_stride, C, HW, y, mul, add, total);
  count_launch();
  return check_launch("kb_nhwc_to_nchw");
}

int kb_pconv_mask(const float *mask, int N, int H, int W, int Cin, int ksize, int stride, int pad, float *ratio, float *update_mask,
                  kb_stream_t stream) {
  KB_REQUIRE(ratio && update_mask && N > 0 && H > 0 && W > 0 && Cin > 0 && ksize >= 1 && ksize <= 7 && stride >= 1 && pad >= 0,
             "kb_pconv_mask: bad arguments");
  const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
  KB_REQUIRE(Ho > 0 && Wo > 0, "kb_pconv_mask: empty output");
  const long total = (long)N * Ho * Wo;
  k_pconv_mask<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(mask, H, W, Cin, ksize, stride, pad, Ho, Wo, ratio, update_mask,
                                                                  total);
  count_launch();
  return check_launch("kb_pconv_mask");
}

}  // extern "C"
