/*
 * kb200.h -- C ABI of libkb200.so: the B200 (sm_100a) implementation of the novel-view synthesis hot path
 * of pierlj/ken-burns-effect.  This is the drop-in boundary: plain pointers and sizes, an explicit CUDA
 * stream, no torch types.  Paths below are relative to the reference tree.
 *
 * The reference has no FFI of its own for this path: it JIT-compiles CUDA source strings through cupy and
 * launches them with raw `tensor.data_ptr()` integers (utils/common.py:377-380, :516-521).  Each entry
 * point here replaces one such launch site (or one torch/numpy/OpenCV stage of the per-frame loop,
 * utils/common.py:222-260); the Python mirror that binds them with ctypes is
 * ken_burns_effect_b200/utils/common.py, and INTEGRATION.md shows the stub a reference maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - tensors are fp32, contiguous, laid out as the reference lays them out (NCHW / [B,C,N]) unless a
 *     parameter says otherwise;
 *   - one call = enqueue on `stream` only (no allocation, no synchronisation, re-entrant);
 *   - return value: 0 on success, a cudaError_t (>0) from the launch, or a negative KB_E* argument error;
 *     kb_last_error() returns a thread-local description of the last failure;
 *   - focal and baseline are doubles because the reference pastes them as double literals into its
 *     kernels (utils/common.py:447,470) and parts of the arithmetic are evaluated in fp64 there.
 */
#ifndef KB200_H
#define KB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define KB_EINVAL (-1)   /* bad argument (null pointer, non-positive size, unsupported C) */
#define KB_ENOSUP (-2)   /* valid request this build does not support */

#define KB_MAX_SIDE 4194302 /* largest image width / height (the splat rounds pixel coordinates exactly below 2^22) */

typedef void *kb_stream_t; /* cudaStream_t */

/* Library identification.  kb_version() = 10000*major + 100*minor + patch. */
int kb_version(void);
const char *kb_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
long long kb_launch_count(void);

/* ---- render_pointcloud, utils/common.py:428-686 ------------------------------------------------ */

/* process_shift's tensor half, utils/common.py:104-109: out = clone(xyz); out.xy *= z/(z+1e-7); out += shift.
 * xyz, out: [B,3,N]; shift: [B,3] (device) */
int kb_shift_points(const float *xyz, const float *shift, float *out, int B, long N, kb_stream_t stream);

/* kernel_pointrender_updateZee, utils/common.py:434-521.
 * xyz [B,3,N].  shift_host: optional HOST pointer to 3 floats added to every point inside the kernel after
 * the process_shift rescale (fuses utils/common.py:104-109; pass NULL for already shifted points).
 * zee [B,H,W] is (re)initialised to 1e6 by this call (utils/common.py:430) and then min-reduced.
 * pix_idx: optional [B,N] int32, receives y*W+x of the pixel each point votes for, or -1 (index map). */
int kb_splat_min(const float *xyz, int B, long N, const float *shift_host, double focal, double baseline,
                 float *zee, int H, int W, int32_t *pix_idx, kb_stream_t stream);

/* kernel_pointrender_updateDegrid, utils/common.py:524-582, race-free: reads zee_in, writes zee_out
 * (the reference updates in place while neighbours are being read; see DESIGN.md "degrid"). */
int kb_degrid(const float *zee_in, float *zee_out, int B, int H, int W, kb_stream_t stream);

/* kernel_pointrender_updateOutput, utils/common.py:585-684.
 * data [B,C,N]; zee [B,H,W] (degridded); accum [B,H,W,Cp] channels-last, Cp = kb_accum_channels(C),
 * zero-initialised by this call; channel C holds the weight sum (the reference's appended ones channel,
 * utils/common.py:429), channels > C are padding. */
int kb_accum_channels(int C);
int kb_splat_accum(const float *xyz, const float *data, int B, long N, int C, const float *shift_host,
                   double focal, double baseline, const float *zee, float *accum, int H, int W,
                   kb_stream_t stream);

/* utils/common.py:686: render[B,C,H,W] = accum[..., :C] / (accum[..., C] + 1e-7); existing[B,1,H,W] = accum[..., C]. */
int kb_normalize(const float *accum, int B, int C, int H, int W, float *render, float *existing,
                 kb_stream_t stream);

/* The same accumulation with the per-point data as ROWS: data_rows[(b*N + n) * row_stride + c], channels contiguous (the NHWC
 * layout kb_conv2d writes), so features go from a convolution epilogue into the splat without a layout change
 * (models/pointcloud_inpainting.py:199-206). */
int kb_splat_accum_rows(const float *xyz, const float *data_rows, long row_stride, int B, long N, int C, const float *shift_host,
                        double focal, double baseline, const float *zee, float *accum, int H, int W, kb_stream_t stream);
/* weight[B,1,H,W] = accum[..., C]  (tensorExisting of utils/common.py:686, before thresholding). */
int kb_accum_weight(const float *accum, int B, int C, int H, int W, float *weight, kb_stream_t stream);
/* In place: accum[..., c] = accum[..., c] / (accum[..., C] + 1e-7) * mask for c < C, then accum[..., C] = mask (mask [B,H,W],
 * NULL = ones): utils/common.py:686 followed by `render * existing` and torch.cat([data, mask]) of
 * models/pointcloud_inpainting.py:210, :135 -- the accumulator becomes the NHWC input of the inpainting GridNet. */
int kb_normalize_rows(float *accum, int B, int C, int H, int W, const float *mask, kb_stream_t stream);

/* The whole of render_pointcloud (four launches above).  workspace: kb_render_workspace_bytes() bytes. */
size_t kb_render_workspace_bytes(int B, int C, int H, int W);
int kb_render_pointcloud(const float *xyz, const float *data, int B, long N, int C, int W, int H,
                         double focal, double baseline, float *render, float *existing, void *workspace,
                         kb_stream_t stream);

/* ---- fill_disocclusion, utils/common.py:833-937 -------------------------------------------------- */
/* input [B,C,H,W], depth [B,1,H,W] -> output [B,C,H,W] (input with hole pixels, depth<=0, filled). */
int kb_fill(const float *input, const float *depth, float *output, int B, int C, int H, int W,
            kb_stream_t stream);

/* ---- spatial_filter on masks, utils/common.py:417-421 as used by pointcloud_inpainting.py:208-209 -- */
/* out = median5x5(in) for in in {0,1} (reflect padding) == (5x5 box count >= 13). in/out [B,1,H,W]. */
int kb_median5_binary(const float *in, float *out, int B, int H, int W, kb_stream_t stream);

/* ---- generate_mask, utils/common.py:689-830 (training-time disocclusion mask) -------------------------------------- */
/* xyz [B,3,N] (points already shifted, :690) -> mask [B,N]: 1 where the point owns the z-buffer cell of the pixel it votes for.
 * The reference's check-then-atomicMin / atomicExch bookkeeping races; this is its outcome for threads running in index order
 * (the lowest-indexed point among those with the minimal err wins a pixel).  The caller applies the median-5
 * (:829, kb_median5_binary).  workspace: kb_mask_workspace_bytes() bytes. */
size_t kb_mask_workspace_bytes(int B, long N, int H, int W);
int kb_generate_mask(const float *xyz, int B, long N, double focal, double baseline, int H, int W, float *mask, void *workspace,
                     kb_stream_t stream);

/* spatial_filter(x, 'laplacian'), utils/common.py:398-409: the reference's 5-tap kernel on a replicate-padded map, applied to
 * each of `planes` = B*C planes [H,W] independently (the reference builds a block-diagonal conv2d for it). */
int kb_laplacian5(const float *in, float *out, int planes, int H, int W, kb_stream_t stream);

/* ---- the per-frame loop of process_kenburns, utils/common.py:222-260, fused ---------------------- */

typedef struct kb_pose {
  float shift[3];   /* tensorShift of process_shift (utils/common.py:98-102), already rounded to fp32 */
  float _pad;
  double focal;     /* currentFocal (utils/common.py:225-229) */
} kb_pose;

typedef struct kb_frame_params {
  int H, W;              /* render / output frame size (objectCommon intHeight/intWidth) */
  int crop_w, crop_h;    /* getRectSubPix patch size (utils/common.py:256) */
  double baseline;       /* objectCommon['dblBaseline'] */
} kb_frame_params;

#define KB_MAX_POSES 32

/* Bytes of device scratch needed to render K poses at once. */
size_t kb_frames_workspace_bytes(const kb_frame_params *p, int K);

/* Renders K (<= KB_MAX_POSES) frames of one point cloud in one call: for each pose
 *   process_shift -> render_pointcloud(C=4: RGB+depth) -> fill_disocclusion -> *255/clip/uint8 ->
 *   getRectSubPix(crop) -> resize(W,H)                      (utils/common.py:238-257)
 * xyz [3,N] UNSHIFTED points (tensorInpaPoints), rgbd [4,N] (tensorInpaImage ++ tensorInpaDepth),
 * poses_host: K poses in HOST memory (copied into kernel parameters, no H2D transfer),
 * frames: device (or mapped pinned host) buffer [K,H,W,3] uint8, RGB order as the reference's numpyOutput. */
int kb_render_frames(const float *xyz, const float *rgbd, long N, const kb_pose *poses_host, int K,
                     const kb_frame_params *p, void *workspace, uint8_t *frames, kb_stream_t stream);

/* ---- the conv stacks: every nn.Conv2d of models/disparity_estimation.py, models/disparity_refinement*.py,
 *      models/pointcloud_inpainting.py and what surrounds it (bias, PReLU, residual / skip sums) ------------
 * Activations are NHWC fp32 ("pixel stride" = floats between consecutive pixels, a multiple of 4, which lets a
 * tensor be a channel slice of a wider concat buffer -- torch.cat of the reference, disparity_refinement.py:100-104).
 * Arithmetic: TF32 operands, fp32 accumulation on the tcgen05 tensor cores (what cuDNN does for the reference under
 * PyTorch's default torch.backends.cudnn.allow_tf32=True on this GPU). */

/* Number of floats kb_conv_pack_weights writes for a [Cout,Cin,k,k] filter. */
long kb_conv_packed_floats(int Cout, int Cin, int ksize);
/* nn.Conv2d.weight [Cout,Cin,k,k] (device) -> the layout kb_conv2d reads, rounded to TF32.  out_scale: optional
 * [Cout] factor folded into the filter (eval-mode BatchNorm of the VGG trunk, disparity_estimation.py:86-105). */
int kb_conv_pack_weights(const float *w_oihw, int Cout, int Cin, int ksize, const float *out_scale, float *packed,
                         kb_stream_t stream);

/* The same filters as fp16 panels for kb_conv_args.x_f16 (halves: kb_conv_packed_halves). */
long kb_conv_packed_halves(int Cout, int Cin, int ksize);
int kb_conv_pack_weights_f16(const float *w_oihw, int Cout, int Cin, int ksize, const float *out_scale, void *packed,
                             kb_stream_t stream);

typedef struct kb_conv_out {
  float *ptr;            /* [N,Ho,Wo,pixel_stride] */
  long pixel_stride;     /* floats, multiple of 4, >= round_up(Cout,4); channels Cout..round_up(Cout,4) are written as 0 */
  const float *slope;    /* per-channel PReLU weight applied to THIS output (NULL: none) -- the activation that
                            precedes the consumer's convolution ('relu-conv-relu-conv', pointcloud_inpainting.py:12-17) */
  const float *mul;      /* optional [N,Ho,Wo] per-pixel factor applied last: the {0,1} mask a PartialConv2d consumer multiplies
                            its input with (utils/partial_conv.py:71: conv(input * mask)) */
  int round_tf32;        /* 1: round to TF32 (value is only ever read by another convolution) */
  int store_f16;         /* 1: store as fp16 (ptr points at halves, pixel_stride counts halves, multiple of 8): the value is the
                            input of a kind::f16 convolution (x_f16 below); saturated to +-65504 */
} kb_conv_out;

typedef struct kb_conv_args {
  const float *x;        /* [N,H,W,x_stride] */
  int N, H, W, Cin;
  long x_stride;
  const float *w_packed; /* kb_conv_pack_weights output */
  const float *bias;     /* [Cout] or NULL */
  int Cout, ksize, stride, pad;   /* ksize 1..7, stride 1 or 2 */
  const float *res;      /* optional tensor added before the outputs' PReLUs (residual / GridNet skip sum), [N,Ho,Wo,res_stride] */
  long res_stride;
  const float *pc_ratio; /* PartialConv2d (utils/partial_conv.py:62-77): per-pixel mask_ratio and update_mask [N,Ho,Wo] from */
  const float *pc_um;    /* kb_pconv_mask; the conv output becomes ((conv + b - b) * ratio + b) * update_mask.  NULL: dense conv */
  int n_out;             /* 1..3 */
  kb_conv_out out[3];
  int out_H, out_W;      /* 0 = the convolution's own output size; smaller: only the top-left out_H x out_W pixels exist in
                            out/res (the reference crops an up-sampled odd-sized map with F.pad(-1), pointcloud_inpainting.py:154) */
  int tile_w;            /* 0 = auto; output tile is tile_w x (128/tile_w) pixels */
  int n_block;           /* 0 = auto; output channels per CTA (multiple of 16, <= 256) */
  int stages;            /* 0 = auto; shared-memory pipeline depth */
  int algo;              /* 0 = auto; 1 = one TMA load per filter tap (any filter); 2 = persistent halo kernel (stride 1, k <= 3) */
  int x_f16;             /* 1: x holds fp16 (x_stride counts halves, multiple of 8) and w_packed comes from kb_conv_pack_weights_f16:
                            tcgen05 kind::f16 -- the same 10 mantissa bits the TF32 path keeps of its operands, twice the MACs per
                            MMA and half the operand bytes; accumulation, bias, residual and un-flagged outputs stay fp32 */
} kb_conv_args;

/* out_o = prelu_o(pc(conv(x, w) + bias) + res) * mul_o   for o < n_out;  one launch. */
int kb_conv2d(const kb_conv_args *args, kb_stream_t stream);

/* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) followed by the block's first PReLU
 * (models/pointcloud_inpainting.py:70-72); the output may be cropped to Ho x Wo (the reference's F.pad(-1), :154-155). */
int kb_upsample2x_prelu(const float *x, long x_stride, int N, int H, int W, int C, const float *slope, float *y, long y_stride,
                        int Ho, int Wo, int round_tf32 /* 1: round to TF32; 2: y holds fp16, y_stride counts halves */,
                        const float *mul /* optional [N,Ho,Wo] factor, as kb_conv_out.mul */, kb_stream_t stream);
/* Mask bookkeeping of PartialConv2d(multi_channel=True), utils/partial_conv.py:43-69, for masks whose channels are identical
 * (all masks of models/partial_inpainting.py): mask [N,H,W] of {0,1} (NULL = no mask = ones) ->
 * update_mask = clamp(Cin * box_k(mask), 0, 1) and ratio = Cin*k*k / (Cin * box_k(mask) + 1e-8) * update_mask, both [N,Ho,Wo]. */
int kb_pconv_mask(const float *mask, int N, int H, int W, int Cin, int ksize, int stride, int pad, float *ratio, float *update_mask,
                  kb_stream_t stream);
/* y = prelu(x) per channel (slope NULL: strided copy). */
int kb_prelu_nhwc(const float *x, long x_stride, long pixels, int C, const float *slope, float *y, long y_stride, int round_tf32,
                  kb_stream_t stream);
/* nn.MaxPool2d(2, 2, ceil_mode=True), models/disparity_estimation.py:90. */
int kb_maxpool2_ceil(const float *x, long x_stride, int N, int H, int W, int C, float *y, long y_stride, kb_stream_t stream);
/* Layout changes at the module boundary: y_nhwc = (x_nchw - sub) * mul ; y_nchw = x_nhwc * mul + add. */
int kb_nchw_to_nhwc(const float *x, int N, int C, int H, int W, float *y, long y_stride, float sub, float mul, kb_stream_t stream);
int kb_nhwc_to_nchw(const float *x, long x_stride, int N, int C, int H, int W, float *y, float mul, float add, kb_stream_t stream);

/* ---- image front end, kbe.py:96-114 and :181 ----------------------------------------------------- */
/* src: uint8 [H,W,3] as cv2.imread returns it (device memory); dst: float [3, H - H%4, W - W%4] = ((ToTensor -> Normalize(.5,.5))
 * cropped to multiples of 4, + 1) / 2 -- the tensor kbe.py hands to Pipeline.__call__, bit for bit.  swap_rb: the
 * cv2.cvtColor(BGR2RGB) of `--pretrained-estim` (kbe.py:97-98). */
int kb_image_front_end(const unsigned char *src, int H, int W, int swap_rb, float *dst, kb_stream_t stream);

/* ---- process_autozoom, utils/common.py:114-170 -------------------------------------------------- */
/* For K (<= KB_MAX_POSES) camera shifts of one cloud xyz [3,N] (unshifted): counts[k] = number of pixels where the `existing` map
 * of render_pointcloud(process_shift(points, shift_k), ...) is > 0 (:154-160: the score the reference maximises over a 16 x 16
 * grid of candidate windows with one full render each).  counts: device int[K]; workspace: kb_coverage_workspace_bytes. */
size_t kb_coverage_workspace_bytes(int H, int W, int K);
int kb_coverage(const float *xyz, long N, const kb_pose *poses_host, int K, int H, int W, double baseline, void *workspace,
                int *counts, kb_stream_t stream);

/* ---- self-test ------------------------------------------------------------------------------------ */
/* The kernels replace the fp64 sub-expressions the reference's source substitution creates (utils/common.py:453,
 * :467-470, :556-561, :639) and the IEEE divisions of the frame tail (:686, :255) by cheaper fp32 sequences that are
 * the same functions.  This entry point evaluates both forms on the device for n pairs (a[i], b[i]) and adds the
 * number of disagreements to *mismatches (device, caller-zeroed):
 *   which 0: quantised a / (|b| + 1e-7) through the shared reciprocal vs the IEEE division
 *         1: a >= b + 1 and a <= b + 1 (exact fp32) vs the fp64 comparisons
 *         2: floor / round-half-away of a without conversions vs floorf / roundf (|a| < 2^22)
 *         3: a + (0.5*W - 0.5) vs the two fp64 additions; a >= 0.001f vs (double)a < 0.001; float -> double widening */
int kb_selftest_arith(int which, const float *a, const float *b, long n, int W, unsigned long long *mismatches,
                      kb_stream_t stream);

/* ---- measurement hooks (bench.py) ---------------------------------------------------------------- */
/* Stages of kb_render_frames, in launch order: 0 memset(accumulators) 1 init(z-buffer + resize tables)
 * 2 splat_min 3 degrid 4 splat_accum 5 resolve (normalise+quantise+hole list) 6 fill 7 crop+resize. */
#define KB_FRAME_STAGES 8
/* While enabled, kb_render_frames records CUDA events on its stream around every stage. */
int kb_profile_enable(int on);
/* Synchronises on the recorded events, returns the summed milliseconds per stage over all calls since the
 * last read (stage_ms[KB_FRAME_STAGES]) and the number of calls; clears the record. */
int kb_profile_read(double *stage_ms, long long *calls);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* KB200_H */
