/*
 * rvsr_b200.h -- C ABI of the B200-native RealVSR/EDVR hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry
 * point names the reference interface it replaces (paths relative to the reference
 * repo IanYeung/RealVSR, codes/models/archs/).
 *
 * Conventions (all entry points):
 *   - return 0 on success, a negative RVSR_E_* code on failure; never throws.
 *     rvsr_last_error() returns a thread-local message for the last failure.
 *   - all device work is enqueued on the cudaStream_t passed as `stream` (void*),
 *     no host synchronisation, no allocation inside per-call entry points; the caller
 *     owns every buffer including the workspace.
 *   - dtype codes: RVSR_F32 = 0, RVSR_F16 = 1, RVSR_BF16 = 2 (DCN operator and training entry points).
 *   - "NCHW" tensors are contiguous like the reference's; the inference engine's
 *     activation layout (channel-blocked [N][C/8][H][W][8]) stays private to it.  Only the
 *     training entry points (rvsr_c8_*) take channel-blocked bf16 tensors.
 */
#ifndef RVSR_B200_H
#define RVSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVSR_F32 0
#define RVSR_F16 1
#define RVSR_BF16 2 /* DCN operator (rvsr_mdcn_fwd / rvsr_mdcn_bwd) and the rvsr_c8_* training entry points: bfloat16 tensors in and out -- the dtype of a
                       torch.autocast(bfloat16) training step (BASELINE cfg5).  EDVR's shape class (64 -> 64, 3x3, stride 1,
                       8 deformable groups) runs on tcgen05 in both directions (bf16 / fp16 operands, fp32 accumulate); other
                       shapes are widened to the fp32 CUDA-core kernels and rounded once.  The reference's extension has no
                       bf16 dispatch at all (deform_conv_cuda_kernel.cu:781) */

#define RVSR_OK 0
#define RVSR_E_INVALID (-1)     /* bad argument / unsupported shape (reference: AT_ERROR / shape_check) */
#define RVSR_E_CUDA (-2)        /* CUDA runtime error (reference only printf'd these, .cu:794-798) */
#define RVSR_E_WORKSPACE (-3)   /* workspace too small */
#define RVSR_E_STATE (-4)       /* engine not finalised / weight missing */
#define RVSR_E_UNSUPPORTED (-5) /* valid in the reference but not built here yet */

#define RVSR_ACT_NONE 0
#define RVSR_ACT_LRELU 1 /* LeakyReLU(0.1), EDVR_arch.py:96 */
#define RVSR_ACT_RELU 2

/* ------------------------------------------------------------------ misc */
int rvsr_version(void);
const char *rvsr_last_error(void);
/* 1 if the current device is sm_100 (B200); 0 otherwise; <0 on CUDA error. */
int rvsr_device_ok(void);

/* ------------------------------------------------------------------ DCNv2 operator
 * Replaces the reference's pybind module `deform_conv_cuda`
 * (dcn/src/deform_conv_cuda.cpp:687-701), v2 functions.
 */

/* modulated_deform_conv_cuda_forward (deform_conv_cuda.cpp:490-569) +
 * modulated_deformable_im2col_gpu_kernel (deform_conv_cuda_kernel.cu:571-633), fused:
 * gather and contraction happen in one kernel, no `columns` buffer, bias added in the
 * epilogue.  All tensors NCHW of `dtype`; bias may be NULL (with_bias=false).
 * offset [B, dg*2*kh*kw, Ho, Wo], mask [B, dg*kh*kw, Ho, Wo] (already sigmoid-ed).
 * workspace: rvsr_mdcn_fwd_workspace_bytes(). */
size_t rvsr_mdcn_fwd_workspace_bytes(int B, int C, int H, int W, int Cout, int kh, int kw,
                                     int stride, int pad, int dil, int groups, int dg, int dtype);
int rvsr_mdcn_fwd(const void *input, const void *offset, const void *mask, const void *weight,
                  const void *bias, void *output, int B, int C, int H, int W, int Cout, int kh,
                  int kw, int stride, int pad, int dil, int groups, int dg, int dtype,
                  void *workspace, size_t workspace_bytes, void *stream);

/* modulated_deform_conv_cuda_backward (deform_conv_cuda.cpp:571-685) + the col2im /
 * col2im_coord kernels (deform_conv_cuda_kernel.cu:635-767).  grad_input, grad_offset,
 * grad_mask are overwritten; grad_weight and grad_bias (may be NULL) must be zeroed by
 * the caller and are accumulated into, like the reference (cpp:659-671).  RVSR_F32, or RVSR_BF16 (all ten tensors
 * bfloat16; gradients are computed and accumulated in fp32 and rounded once on the way out). */
size_t rvsr_mdcn_bwd_workspace_bytes(int B, int C, int H, int W, int Cout, int kh, int kw,
                                     int stride, int pad, int dil, int groups, int dg, int dtype);
int rvsr_mdcn_bwd(const void *input, const void *offset, const void *mask, const void *weight,
                  const void *grad_output, void *grad_input, void *grad_offset, void *grad_mask,
                  void *grad_weight, void *grad_bias, int B, int C, int H, int W, int Cout, int kh,
                  int kw, int stride, int pad, int dil, int groups, int dg, int dtype,
                  void *workspace, size_t workspace_bytes, void *stream);

/* ModulatedDeformConvPack.forward with extra_offset_mask=True (dcn/deform_conv.py:274-292):
 * conv_offset_mask(feat) -> chunk/cat/sigmoid -> modulated deform conv of x, optional
 * LeakyReLU(0.1) (EDVR_arch.py:107,:130).  NCHW in / NCHW out; 3x3, stride 1, pad 1, dil 1,
 * groups 1 (the only configuration EDVR_arch.py:73-94 instantiates). */
size_t rvsr_mdcn_pack_fwd_workspace_bytes(int B, int C, int H, int W, int Cout, int dg, int dtype);
int rvsr_mdcn_pack_fwd(const void *x, const void *feat, const void *w_offset_mask,
                       const void *b_offset_mask, const void *weight, const void *bias, void *y,
                       int B, int C, int H, int W, int Cout, int dg, int act, int dtype,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------ convolution operator
 * One nn.Conv2d site of EDVR_arch.py (:71-91, :146-164, :229-253) with what surrounds it fused:
 * torch.cat([x1, x2], 1) on the input (x2 may be NULL), bias, activation, residual add (after
 * the activation; may be NULL) and optionally nn.PixelShuffle(2) on the output
 * (y is then [B, Cout/4, 2Ho, 2Wo]).  ks in {1,3}, pad = ks/2, stride in {1,2}.  NCHW tensors
 * of `dtype`.  use_tc = 1 selects the tcgen05 kernel (fp16 only; RVSR_E_UNSUPPORTED if the
 * shape is not covered), 0 the CUDA-core kernel. */
size_t rvsr_conv2d_fwd_workspace_bytes(int B, int C1, int C2, int H, int W, int Cout, int ks, int dtype);
int rvsr_conv2d_fwd(const void *x1, const void *x2, const void *weight, const void *bias,
                    const void *residual, void *y, int B, int C1, int C2, int H, int W, int Cout,
                    int ks, int stride, int act, int shuffle, int dtype, int use_tc, void *workspace,
                    size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------ training path (bf16, channel-blocked tensors)
 * BASELINE cfg5 (bf16 training step): the plain convolutions of EDVR_arch.py forward AND backward on this library's
 * tcgen05 kernels, called from torch.autograd.Functions (realvsr_b200/train_c8.py) that replace nn.Conv2d's autograd
 * (EDVR_arch.py:71-91, :229-253; arch_util.py:121-139), F.interpolate x2 (:109-121) and PixelShuffle(2) + lrelu (:313-314).
 * "C8" tensors are [N][ceil(C/8)][H][W][8] bfloat16, contiguous per image (x_image_stride in elements lets a source be a
 * channel slice of a wider tensor); channels beyond C are zero.
 *   rvsr_c8_conv_pack_weight  fp32 OIHW nn.Parameter -> the kernels' bf16 operand layout.  mode 0: forward convolution over
 *                             input channels [w_c0, w_c0 + Cin) of rows with w_cin_total channels.  mode 1: the DATA GRADIENT
 *                             of input channels [w_c0, w_c0 + Cout) as a convolution Cin (= forward Cout) -> Cout with the
 *                             transposed, flipped weights (dX = conv(dY, W^T flipped); stride 1 only).  layouts: bit 0 = the
 *                             single-CTA kernels' operand layout, bit 1 = the CTA-pair kernels' (read by 3x3 launches of >= 4
 *                             tiles); 3 packs both.
 *   rvsr_c8_conv_fwd          y = act(conv(cat(x[0..nsrc)), w) + bias) [+ residual], optional fused PixelShuffle(2); nsrc sources
 *                             of C channels each (C % 16 == 0, C <= 64); also runs every data gradient (with mode-1 weights).
 *                             residual_mode 0: y += residual.  residual_mode 2: `residual` is the OUTPUT of the activation this
 *                             data gradient flows into, y *= (residual > 0 ? 1 : residual_slope) -- its gradient, fused.
 *   rvsr_c8_conv_wgrad        njobs (<= 8) weight gradients of one geometry in one launch (the sources of a torch.cat convolution,
 *                             the two convolutions of a fused pair).  Job j: gw[j][co][c0[j] + ci][ky][kx] = sum_pixels
 *                             x[j][pixel + (ky, kx) - pad][ci] * g[j][pixel][co] for the 64 input channels of source x[j] (OIHW fp32
 *                             [Cout][cin_total[j]][ks][ks]; written, fixed summation order), db[j][co] = sum_pixels g[j] (may be
 *                             NULL).  Cin = 64 per job, ks in {1, 3}.
 *   rvsr_c8_act_bwd           out = y > 0 ? g : slope * g  (LeakyReLU(0.1) / ReLU given the layer OUTPUT y)
 *   rvsr_c8_unshuffle2_act_bwd  gradient through lrelu(PixelShuffle(2)(.)): g, y [N][C/4 ch][2H][2W] -> out [N][C ch][H][W]
 *   rvsr_c8_upsample2x        bilinear x2 (align_corners=False) times `scale`, or (backward = 1) its adjoint. */
/* NCHW (RVSR_BF16 / RVSR_F32) <-> C8 with `planes` >= ceil(C / 8) channel blocks per image: only blocks [0, ceil(C / 8)) are
 * written (channels C .. of the last one as zeros) / read */
int rvsr_c8_from_nchw(const void *src, int src_dtype, void *dst_c8, int N, int C, int H, int W, int planes, void *stream);
int rvsr_c8_to_nchw(const void *src_c8, void *dst, int dst_dtype, int N, int C, int H, int W, int planes, void *stream);
size_t rvsr_c8_conv_weight_bytes(int Cout, int Cin, int ks, int shuffle);
int rvsr_c8_conv_pack_weight(const float *weight, void *dst, int Cout, int Cin, int ks, int shuffle, int mode, int w_cin_total,
                             int w_c0, int layouts, void *stream);
/* rvsr_c8_conv_pack_weight for several views of ONE weight tensor in one launch (a layer's forward operand and the operands of its
 * data gradients): spec holds 8 ints per view {Cout, Cin, ks, shuffle, mode, w_cin_total, w_c0, layouts}, dst[k] is sized by
 * rvsr_c8_conv_weight_bytes of view k; at most 8 layouts in total */
int rvsr_c8_conv_pack_weights(const float *weight, int nviews, const int *spec, void *const *dst, void *stream);
/* which operand layout (rvsr_c8_conv_pack_weight's `layouts` bits) a rvsr_c8_conv_fwd launch of this shape reads: 1, 2, or 0 when the
 * shape is not covered -- a caller may pack just that one */
int rvsr_c8_conv_layouts(int nsrc, int C, int N, int H, int W, int Cout, int ks, int shuffle);
int rvsr_c8_conv_fwd(const void *const *x, const long long *x_image_stride, int nsrc, int C, const void *w_packed, const float *bias,
                     const void *residual, void *y, int N, int H, int W, int Cout, int ks, int stride, int act, int shuffle,
                     int residual_mode, float residual_slope, void *stream);
size_t rvsr_c8_conv_wgrad_workspace_bytes(int njobs, int N, int H, int W, int Cout);
int rvsr_c8_conv_wgrad(int njobs, const void *const *x, const long long *x_image_stride, const void *const *g, float *const *gw,
                       float *const *db, const int *cin_total, const int *c0, int N, int H, int W, int Cin, int Cout, int ks,
                       void *workspace, size_t workspace_bytes, void *stream);
int rvsr_c8_act_bwd(const void *g, const void *y, void *out, long long n_elems, int act, void *stream);
int rvsr_c8_unshuffle2_act_bwd(const void *g, const void *y, void *out, int N, int C, int H, int W, int act, void *stream);
int rvsr_c8_upsample2x(const void *src, void *dst, long long planes, int H, int W, float scale, int backward, void *stream);
/* TSA_Fusion's temporal attention (EDVR_arch.py:170-181) and its gradient on 64-channel C8 tensors: aligned, emb [B * N][8][H][W][8],
 * emb_ref [B][8][H][W][8]; prob[b][n][pixel] = sigmoid(<emb[b, n], emb_ref[b]>) (fp32, kept for the backward);
 * out[n] ([B][8][H][W][8], one tensor per frame = the N sources of the 1x1 fusion convolutions) = aligned[b, n] * prob.
 * bwd: gout[n] may be NULL (no gradient for that frame); g_aligned, g_emb, g_emb_ref are written. */
int rvsr_c8_tsa_temporal(const void *aligned, const void *emb, const void *emb_ref, void *const *out, float *prob, int B, int N, int C,
                         int H, int W, void *stream);
int rvsr_c8_tsa_temporal_bwd(const void *const *gout, const void *aligned, const void *emb, const void *emb_ref, const float *prob,
                             void *g_aligned, void *g_emb, void *g_emb_ref, int B, int N, int C, int H, int W, void *stream);
/* TSA_Fusion's spatial attention helpers on C8 tensors (`planes` = images x channel blocks): MaxPool2d(3, 2, 1) + AvgPool2d(3, 2, 1)
 * of one tensor in one pass (EDVR_arch.py:154-155) and the gradient w.r.t. that tensor (g_max / g_avg may be NULL; the max
 * gradient goes to the first maximum of a window, torch's rule); out = fea * sigmoid(att) * 2 + att_add (:206-207) and its
 * gradients g_fea, g_att (the gradient of att_add is g itself). */
int rvsr_c8_pool_maxavg(const void *src, void *dst_max, void *dst_avg, long long planes, int H, int W, void *stream);
int rvsr_c8_pool_maxavg_bwd(const void *src, const void *g_max, const void *g_avg, void *g_src, long long planes, int H, int W, void *stream);
int rvsr_c8_tsa_final(const void *fea, const void *att, const void *att_add, void *out, long long n_elems, void *stream);
int rvsr_c8_tsa_final_bwd(const void *g, const void *fea, const void *att, void *g_fea, void *g_att, long long n_elems, void *stream);
/* ModulatedDeformConvPack.forward / its autograd (deform_conv.py:274-292, :97-153) on C8 tensors, for EDVR's shape class
 * (64 -> 64 channels, 3x3, stride 1, pad 1, 8 deformable groups).  `om` is the output of the conv_offset_mask convolution
 * as a 256-channel C8 tensor (144 offsets, 72 mask logits, 40 zero channels: the 216 outputs padded to whole 64-wide tiles);
 * the chunk / cat / sigmoid of the reference happen inside.  fwd: y = act(dcn(x, offset, sigmoid(mask)) + bias) on
 * dcn_tc_kernel.  bwd (dcn_bwd_tc_kernel): gx, gom (C8 bf16), gw [64][64][3][3] and gb [64] (fp32), all written. */
size_t rvsr_c8_mdcn_workspace_bytes(int N, int H, int W, int backward);
int rvsr_c8_mdcn_fwd(const void *x, const void *om, const float *weight, const float *bias, void *y, int N, int H, int W, int act,
                     void *workspace, size_t workspace_bytes, void *stream);
int rvsr_c8_mdcn_bwd(const void *x, const void *om, const float *weight, const void *g, const void *y, void *gx, void *gom, float *gw,
                     float *gb, int N, int H, int W, int act, void *workspace, size_t workspace_bytes, void *stream);

/* F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False) * scale on a contiguous NCHW tensor (planes = N * C) of
 * dtype RVSR_F32 / F16 / BF16, or (backward = 1) its adjoint: src = gradient of the [2H][2W] result, dst = gradient of the input.
 * The reference calls it at EDVR_arch.py:53-57, :109-121, :195-202 -- torch's NCHW kernel for it is 16 % of an fp32 training step. */
int rvsr_upsample2x_nchw(const void *src, void *dst, long long planes, int H, int W, float scale, int backward, int dtype, void *stream);

/* ------------------------------------------------------------------ EDVR engine
 * Replaces EDVR.forward / EDVR_NoUp.forward (EDVR_arch.py:258-320, :358-404) and everything
 * they call (PCD_Align :98-132, TSA_Fusion :168-208, ResidualBlock_noBN arch_util.py:135-139)
 * for inference.  Constructor arguments mirror EDVR.__init__ (EDVR_arch.py:212-213). */
typedef struct rvsr_edvr_config {
    int nf, nc, nframes, groups, front_RBs, back_RBs;
    int center;    /* -1 => nframes / 2 (EDVR_arch.py:217) */
    int predeblur; /* 1: Predeblur_ResNet_Pyramid + conv_1x1 in front of the feature extraction (EDVR_arch.py:15-59, :226-227, :262-266) */
    int HR_in;     /* 1: frames arrive at the OUTPUT resolution; stem conv_first_1 -> _2 (stride 2) -> _3 (stride 2), base = the
                      centre frame itself (EDVR_arch.py:228-231, :267-274, :315-316).  Both are ignored when upsample = 0 (EDVR_NoUp) */
    int w_TSA;
    int upsample;  /* 1 = EDVR (x4 pixel-shuffle tail), 0 = EDVR_NoUp */
    int precision; /* RVSR_F32: fp32 storage, SIMT kernels (strict parity);
                      RVSR_F16: fp16 storage, fp32 accumulate, tcgen05 kernels */
} rvsr_edvr_config;

typedef struct rvsr_engine rvsr_engine;

int rvsr_engine_create(const rvsr_edvr_config *cfg, rvsr_engine **out);
void rvsr_engine_destroy(rvsr_engine *e);
/* Hand the engine one state_dict tensor (reference key names) as a
 * contiguous fp32 DEVICE buffer in PyTorch OIHW order; the engine copies/repacks it into
 * its private layout on `stream`.  Unknown names -> RVSR_E_INVALID. */
int rvsr_engine_set_weight(rvsr_engine *e, const char *name, const float *dev_ptr,
                           const int64_t *shape, int ndim, void *stream);
/* Check every tensor of the state_dict contract has been supplied; build packed weights. */
int rvsr_engine_finalize(rvsr_engine *e, void *stream);
/* Number of expected state_dict tensors and their names (for strict-load checks). */
int rvsr_engine_num_weights(const rvsr_engine *e);
const char *rvsr_engine_weight_name(const rvsr_engine *e, int i);

size_t rvsr_engine_workspace_bytes(const rvsr_engine *e, int B, int H, int W);
/* x: [B, nframes, nc, H, W] NCHW DEVICE tensor of x_dtype; out: [B, nc, sH, sW] of out_dtype
 * (s = 4 for EDVR, 1 for EDVR_NoUp and for EDVR with HR_in).  H and W must be multiples of 4 (16 with HR_in). */
int rvsr_engine_forward(rvsr_engine *e, const void *x, int x_dtype, void *out, int out_dtype,
                        int B, int H, int W, void *workspace, size_t workspace_bytes,
                        void *stream);
/* Same, but x/out are HOST buffers (pinned for overlap): stages through `dev_in`/`dev_out`
 * device staging buffers supplied by the caller (sizes = the tensors'), H2D + forward + D2H on
 * `stream`; the caller synchronises.  This is what util.single_forward
 * (codes/utils/util.py:222-237) amounts to: .to(device) -> model -> .float().cpu(). */
int rvsr_engine_forward_host(rvsr_engine *e, const void *x_host, int x_dtype, void *out_host,
                             int out_dtype, int B, int H, int W, void *dev_in, void *dev_out,
                             void *workspace, size_t workspace_bytes, void *stream);
/* Sliding-window video inference with a per-frame feature cache.
 * The reference slides its window one frame at a time and recomputes everything
 * (test_RealVSR_wi_GT.py:114-119), but the feature pyramid of a frame (conv_first, front ResBlocks,
 * fea_L2/L3 convs, EDVR_arch.py:276-283) depends on that frame only: extract it once per frame into a
 * cache of n_slots frames, then run alignment + fusion + reconstruction on windows given as slot
 * indices.  Results are bit-identical to rvsr_engine_forward on the same frames.
 *   cache   : caller-owned device buffer of rvsr_engine_cache_bytes(n_slots, H, W)
 *   extract : frames [F, nc, H, W] NCHW of `dtype` -> slots [slot0, slot0 + F)
 *   forward : window_slots = HOST array [B * nframes] (window-major, frame order as in forward's x);
 *             frames = the LQ frames of ALL slots [n_slots, nc, H, W] (for the base of the output);
 *             workspace: rvsr_engine_workspace_bytes(B, H, W) is sufficient. */
size_t rvsr_engine_cache_bytes(const rvsr_engine *e, int n_slots, int H, int W);
size_t rvsr_engine_extract_workspace_bytes(const rvsr_engine *e, int F, int H, int W);
int rvsr_engine_extract_features(rvsr_engine *e, const void *frames, int dtype, int F, int H, int W,
                                 void *cache, int n_slots, int slot0, void *workspace,
                                 size_t workspace_bytes, void *stream);
int rvsr_engine_forward_cached(rvsr_engine *e, const void *cache, int n_slots, const int *window_slots,
                               const void *frames, int x_dtype, void *out, int out_dtype, int B,
                               int H, int W, void *workspace, size_t workspace_bytes, void *stream);
/* How many kernels the last rvsr_engine_forward enqueued (bench.py's gpu_launches). */
int rvsr_engine_last_launch_count(const rvsr_engine *e);
/* Per-launch profiling: when on, every kernel launch of rvsr_engine_forward is bracketed by
 * CUDA events on the launching stream.  profile_collect() waits for them and returns the
 * number of entries (or <0); profile_entry() reads one: label ("tc:<weight>", "simt:<weight>",
 * "glue:<op>"), device milliseconds, algorithmic FLOPs and bytes of that launch. */
int rvsr_engine_set_profiling(rvsr_engine *e, int on);
int rvsr_engine_profile_collect(rvsr_engine *e);
int rvsr_engine_profile_entry(const rvsr_engine *e, int i, char *label, int label_cap, float *ms,
                              double *flops, double *bytes);
/* Debug/parity taps: copy an internal activation (C-blocked) out as NCHW fp32.
 * names: "L1","L2","L3","aligned","fused".  Valid after a forward with the same workspace. */
int rvsr_engine_read_tap(rvsr_engine *e, const char *name, float *dst_dev, size_t dst_elems,
                         void *stream);

/* ---- image I/O around the model, on the device ------------------------------------
 * rvsr_frames_from_u8 replaces, for T frames at once, data/util.py::read_img (:87-101: uint8 -> float32 / 255) and
 * ::read_img_seq (:104-122: channel reversal [2, 1, 0] when reverse_channels != 0, HWC -> CHW, stack):
 *   u8_thwc [T, H, W, C] uint8 (file order, as cv2.imread returns it)  ->  out_tchw [T, C, H, W] of out_dtype.
 * rvsr_frames_to_u8 replaces the reference test loop's egress (test_RealVSR_wi_GT.py:121-128):
 *   color_mode 0: utils/util.py::tensor2img(out_type=uint8, reverse_channel=True) (:151-181)
 *   color_mode 1: tensor2img(out_type=float32, reverse_channel=False) + data/util.py::ycbcr2bgr (:397-416)
 *                 + (np.clip(., 0, 1) * 255.).round().astype(uint8)
 *   in_bchw [B, 3, H, W] (fp32 / fp16)  ->  u8_bhwc_bgr [B, H, W, 3] uint8, ready for cv2.imwrite.  Bit-exact with
 *   the numpy arithmetic (float32 / float64 steps and round-half-to-even reproduced). */
int rvsr_frames_from_u8(const void *u8_thwc, void *out_tchw, int T, int C, int H, int W, int reverse_channels,
                        int out_dtype, void *stream);
int rvsr_frames_to_u8(const void *in_bchw, int in_dtype, void *u8_bhwc_bgr, int B, int C, int H, int W,
                      int color_mode, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RVSR_B200_H */
