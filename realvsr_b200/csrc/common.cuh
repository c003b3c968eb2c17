// common.cuh -- shared helpers for the rvsr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/rvsr_b200.h"

namespace rvsr {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
const char *get_error();

#define RVSR_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            ::rvsr::set_error(__VA_ARGS__);       \
            return RVSR_E_INVALID;                \
        }                                         \
    } while (0)

#define RVSR_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::rvsr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                              __FILE__, __LINE__);                                        \
            return RVSR_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)

#define RVSR_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            ::rvsr::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                        \
            return RVSR_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of the engine path is launched with programmatic stream serialization: its CTAs may become
// resident while the previous kernel of the stream is still draining, run their prologue (barrier init, TMEM
// allocation, weight staging -- nothing produced by the previous kernel) and then block in pdl_wait() until the
// previous grid has completed and its writes are visible.  pdl_trigger() lets the NEXT kernel do the same.
// A kernel launched through launch_k() MUST execute pdl_wait() in every CTA before touching activations.
// RVSR_PDL=0 launches the same kernels fully serialized (the two griddepcontrol instructions become no-ops).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Only launches made inside a PdlScope (the engine's forward plan, whose weights were packed long before) use it;
// the operator-level entry points pack weights right before their kernel and stay fully serialized.
bool pdl_enabled();  // engine.cu
struct PdlScope {
    PdlScope();
    ~PdlScope();
};
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device): remember which pairs were set (a process may
// drive several devices, e.g. the reference's nn.DataParallel wrapper, VideoSR_AllPair_model_YCbCr_Split.py:33-36)
int ensure_max_dynamic_smem(const void *func, int bytes);  // engine.cu

#define RVSR_TRY(expr)            \
    do {                          \
        int _rc = (expr);         \
        if (_rc != RVSR_OK) return _rc; \
    } while (0)

// ---------------------------------------------------------------- scalar type helpers
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// One pixel of one channel block: 8 channels.  16 B for half, 32 B for float.
template <typename T> struct alignas(sizeof(T) * 8) Vec8 { T v[8]; };

template <typename T> __device__ __forceinline__ void load8(const T *p, float (&f)[8]);
template <> __device__ __forceinline__ void load8<float>(const float *p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
    const float4 b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__half>(const __half *p, float (&f)[8]) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
    const __half2 *h = reinterpret_cast<const __half2 *>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(h[i]);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}
template <typename T> __device__ __forceinline__ void store8(T *p, const float (&f)[8]);
template <> __device__ __forceinline__ void store8<float>(float *p, const float (&f)[8]) {
    reinterpret_cast<float4 *>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4 *>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void store8<__half>(__half *p, const float (&f)[8]) {
    uint4 a;
    __half2 *h = reinterpret_cast<__half2 *>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4 *>(p) = a;
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == RVSR_ACT_LRELU) return v > 0.f ? v : 0.1f * v;
    if (act == RVSR_ACT_RELU) return v > 0.f ? v : 0.f;
    return v;
}
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

// Bilinear sample of one channel block (8 channels) at (py, px); zero unless
// -1 < py < H and -1 < px < W; each corner individually bounds-checked
// (deform_conv_cuda_kernel.cu:480-495, :618).
template <typename T>
__device__ __forceinline__ void sample8(const T *plane, int H, int W, float py, float px, float (&v)[8]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = 0.f;
    if (!(py > -1.f && px > -1.f && py < (float)H && px < (float)W)) return;
    const float fy = floorf(py), fx = floorf(px);
    const int y0 = (int)fy, x0 = (int)fx, y1 = y0 + 1, x1 = x0 + 1;
    const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
    float t[8];
    if (y0 >= 0 && x0 >= 0) {
        load8<T>(plane + ((long long)y0 * W + x0) * 8, t);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = fmaf(hy * hx, t[c], v[c]);
    }
    if (y0 >= 0 && x1 <= W - 1) {
        load8<T>(plane + ((long long)y0 * W + x1) * 8, t);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = fmaf(hy * lx, t[c], v[c]);
    }
    if (y1 <= H - 1 && x0 >= 0) {
        load8<T>(plane + ((long long)y1 * W + x0) * 8, t);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = fmaf(ly * hx, t[c], v[c]);
    }
    if (y1 <= H - 1 && x1 <= W - 1) {
        load8<T>(plane + ((long long)y1 * W + x1) * 8, t);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = fmaf(ly * lx, t[c], v[c]);
    }
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------- op descriptors
// A conv input: a channel-blocked tensor [n][C/8][H][W][8].  `frames`/`fixed_frame`
// implement the PCD "reference frame" broadcast (EDVR_arch.py:291-303): with
// fixed_frame >= 0, image n reads image (n / frames) * frames + fixed_frame.
struct Src {
    const void *ptr;
    long long image_stride;  // elements between consecutive images
    int C;                   // channels (multiple of 8 in storage)
    int frames, fixed_frame;
    const int *map;          // optional device table: image n reads image map[n] (feature-cache slots)
    int map_images;          // number of images addressable through `map` (for the TMA tensor map)
};
#define RVSR_MAX_SRC 7        // CUDA-core kernels; EDVR's widest concat is nframes <= 7 frames
#define RVSR_MAX_SRC_TC 14    // tcgen05 kernels: a 128-channel source (nf = 128) is fed as two 64-channel sources

enum OutMode {
    OUT_C8 = 0,           // channel-blocked, same resolution
    OUT_C8_SHUFFLE2 = 1,  // nn.PixelShuffle(2) fused into the store (EDVR_arch.py:311-312)
    OUT_PLANAR_F32 = 2,   // [n][Cout][H][W] fp32; channels >= sig_from get a sigmoid (offset/mask conv)
    OUT_NCHW_T = 3,       // [n][Cout][H][W] of T (operator-level API)
    OUT_OM24 = 4,         // DCN offsets+mask for the tcgen05 gather kernel: per pixel and deformable group
                          // 24 words = 3 blocks of 32 B: [dy0 dx0 .. dy3 dx3][dy4 dx4 .. dy7 dx7]
                          // [dy8 dx8 m01 m23 m45 m67 m8_ 0], offsets fp32, sigmoid(mask) as fp16 pairs;
                          // memory [n][dg*3][H][W][8 words]
    OUT_FINAL = 5,        // network output (tcgen05 conv_last only): NCHW of fin.out_dtype, plus the bilinear x`scale`
                          // upsample of the centre LQ frame (EDVR_arch.py:315-319) -- no conv_last tensor, no extra pass
};

// OUT_FINAL: where the base frame comes from (same arithmetic as final_add_kernel)
struct FinalAdd {
    const void *x;           // LQ clip, NCHW [B * frames (or all cached frames)][nc][H / scale][W / scale]
    const int *center_map;   // optional: image b -> index of its centre frame in x (feature-cache mode)
    int x_dtype, out_dtype;  // RVSR_F32 / RVSR_F16
    int frames, center, nc, scale;
};

// How a weight-packing kernel addresses its fp32 source: element (co, cin, tap) of the convolution being packed is read at
// base + co * s_co + cin * s_ci + tap * s_tap.  OIHW: {0, Cin * KK, KK, 1}.  Data gradient of input channels [c0, c0 + C) of a
// convolution with CinTot inputs (dX = conv(dY, W^T flipped)): {c0 * KK + KK - 1, KK, CinTot * KK, -1}.
struct WeightView {
    long long base, s_co, s_ci, s_tap;
    int bf16;  // destination format: 0 = fp16, 1 = bfloat16
};

#define RVSR_PACK_MAX_VIEWS 8

struct ConvOp {
    Src src[RVSR_MAX_SRC_TC];
    int nsrc;
    const float *w_simt;   // packed fp32 [sum chunks][taps][8][CoutPad]
    const void *w_tc;      // packed fp16 UMMA layout (tc kernels) or null
    const void *w_tc2;     // packed fp16 layout of the CTA-pair kernel (64-wide 3x3 convs) or null
    const float *bias;     // [Cout] fp32 or null
    void *out;
    long long out_image_stride;
    const void *residual;  // same layout as out (OUT_C8 only), added after the activation (res_pre: before bias + activation)
    long long res_image_stride;
    int res_pre;           // 1: out = act(conv + residual + bias) -- the other half of a split torch.cat convolution
    int res_div;           // image n reads residual image n / res_div (0 = 1): one residual per window of res_div frames
    int N, H, W;           // images, input height/width
    int Cout, ks, stride;  // pad = ks / 2
    int act, out_mode, sig_from;
    int dg;                // OUT_OM24: deformable groups (Cout == 27 * dg)
    FinalAdd fin;          // OUT_FINAL
    int bf16;              // tcgen05 kernels: sources, residual and output are bfloat16 (w_tc / w_tc2 packed as bfloat16)
    float res_slope;       // res_pre == 2 (tcgen05 kernels): `residual` is a mask, out = v * (residual > 0 ? 1 : res_slope)
};

struct DcnOp {
    Src x;                 // sampled features
    const float *offset;   // planar fp32 [n][dg*2*K][Ho][Wo]
    const float *mask;     // planar fp32 [n][dg*K][Ho][Wo], already sigmoid-ed
    long long offset_image_stride, mask_image_stride;
    const void *om24;      // if non-null: offsets+mask in OUT_OM24 format (tcgen05 path), offset/mask unused
    long long om24_image_stride;  // in 4-byte words
    const float *w_simt;   // packed like ConvOp
    const void *w_tc;
    const void *w_tc_hi;   // nf = 128: contraction weights of input channels [64, 128) (w_tc: [0, 64)), pack_weight_dcn_tc_half
    const float *bias;
    void *out;
    long long out_image_stride;
    int N, H, W, Cout, kh, kw, stride, pad, dil, dg;
    int act, out_mode;
};

// DCN backward (fp32, CUDA-core): one launch produces all five gradients
struct DcnBwdOp {
    const float *x_c8;                  // [N][C8][H][W][8]
    const float *offset, *mask, *gout;  // NCHW planar fp32
    const float *w_dense;               // [Cout][C][K] (conv groups expanded to dense)
    float *gx_c8;                       // [N][C8][H][W][8], zeroed by the caller
    float *goffset, *gmask;             // NCHW planar, zeroed by the caller
    float *gw_dense;                    // [Cout][C][K], accumulated into
    float *gbias;                       // [Cout] or null, accumulated into
    int N, C, H, W, Cout, kh, kw, stride, pad, dil, dg;
};
int launch_dcn_bwd_simt(const DcnBwdOp &op, cudaStream_t s);
int fold_grouped_weight(const float *gd, float *gw, int Cout, int C, int K, int groups, cudaStream_t s);

// ---------------------------------------------------------------- kernel launchers (simt_kernels.cu)
template <typename T> int launch_conv_simt(const ConvOp &op, cudaStream_t s);
template <typename T> int launch_dcn_simt(const DcnOp &op, cudaStream_t s);
template <typename T, typename Tin>
int launch_pack_nchw(const Tin *src, T *dst, int N, int C, int H, int W, cudaStream_t s, int Cdst = 0);
int pad_weight_cin(const float *w, float *dst, int Cout, int Cin, int Cpad, int KK, cudaStream_t s);
template <typename T, typename Tout>
int launch_unpack_nchw(const T *src, Tout *dst, int N, int C, int H, int W, cudaStream_t s);
template <typename Tout>
int launch_frames_from_u8(const uint8_t *src, Tout *dst, int T, int C, int H, int W, int reverse, cudaStream_t s);
template <typename Tin>
int launch_frames_to_u8(const Tin *src, uint8_t *dst, int B, int H, int W, int mode, cudaStream_t s);
template <typename T>
int launch_upsample2x(const T *src, T *dst, int N, int C, int H, int W, float scale, cudaStream_t s, const T *add = nullptr);
template <typename T>
int launch_pool_maxavg(const T *src, T *dst_max, T *dst_avg, int N, int C, int H, int W, cudaStream_t s);
template <typename T>
int launch_tsa_temporal(const T *emb, const T *emb_ref, const T *aligned, T *out, int B, int frames,
                        int C, int H, int W, cudaStream_t s);
template <typename T>
int launch_tsa_final(const T *fea, const T *att, const T *att_add, T *out, long long n, cudaStream_t s);
template <typename T, typename Tin, typename Tout>
int launch_final_add(const T *res_c8, const Tin *x, Tout *out, int B, int frames, int center, int nc,
                     int H, int W, int scale, cudaStream_t s, const int *center_map = nullptr);
int launch_convert_f16_f32(const void *src, float *dst, long long n, cudaStream_t s);
// planar offsets [n][dg*18][H][W] + sigmoid-ed mask [n][dg*9][H][W] of `dtype` -> OUT_OM24 (see OutMode)
int launch_om24_from_planar(const void *offset, const void *mask, int dtype, void *om24, int N, int dg, int H, int W, cudaStream_t s);
int launch_convert_bf16_f32(const void *src, float *dst, long long n, cudaStream_t s);
int launch_convert_f32_bf16(const float *src, void *dst, long long n, cudaStream_t s);
int launch_fill_f32(float *dst, float v, long long n, cudaStream_t s);

// weight packing (host-callable, device work on stream)
int pack_weight_simt(const float *w_oihw, float *dst, int Cout, int Cin_total, int ks,
                     const int *src_channels, int nsrc, int cout_pad, cudaStream_t s);
int pack_weight_dcn_simt(const float *w_oihw, float *dst, int Cout, int C, int K, int cout_pad,
                         cudaStream_t s);

}  // namespace rvsr
