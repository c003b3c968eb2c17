// api.cu -- the extern "C" surface declared in include/rvsr_b200.h.
#include <new>

#include "engine.cuh"

struct rvsr_engine {
    rvsr::Engine impl;
    explicit rvsr_engine(const rvsr_edvr_config &c) : impl(c) {}
};

namespace rvsr {
int expand_grouped_weight(const float *w, float *dst, int Cout, int C, int K, int groups, cudaStream_t s);


namespace {
struct Carver {  // carve the caller's workspace (or just count, when base == nullptr)
    char *base;
    size_t cap, off = 0;
    bool ok = true;
    void *take(size_t bytes) {
        const size_t a = align_up(off, 256);
        off = a + align_up(bytes, 256);
        if (base == nullptr) return nullptr;
        if (off > cap) { ok = false; return nullptr; }
        return base + a;
    }
};

struct DcnDims {
    int B, C, H, W, Cout, kh, kw, stride, pad, dil, groups, dg, Ho, Wo, K;
};
int dcn_dims(DcnDims &d, int B, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil,
             int groups, int dg) {
    // the reference's checks: deform_conv_cuda.cpp:511-516 and the shape math at :518-521
    RVSR_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0 && Cout > 0, "dcn: bad tensor sizes");
    RVSR_CHECK_ARG(kh > 0 && kw > 0 && stride > 0 && pad >= 0 && dil > 0, "dcn: bad conv parameters");
    RVSR_CHECK_ARG(groups > 0 && C % groups == 0 && Cout % groups == 0,
                   "Input shape and kernel channels wont match: (%d vs groups %d)", C, groups);
    RVSR_CHECK_ARG(dg > 0 && C % dg == 0, "dcn: channels %d not divisible by deformable groups %d", C, dg);
    d = {B, C, H, W, Cout, kh, kw, stride, pad, dil, groups, dg, 0, 0, kh * kw};
    d.Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
    d.Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
    RVSR_CHECK_ARG(d.Ho > 0 && d.Wo > 0, "dcn: convolution input is too small");
    RVSR_CHECK_ARG(d.K == 9 || d.K == 1, "dcn: only 3x3 / 1x1 kernels are built");
    return RVSR_OK;
}

struct DcnWs {
    void *x_c8;
    float *off32, *mask32, *w32, *b32, *wdense, *wpack;
};
void carve_dcn(Carver &cv, const DcnDims &d, int dtype, DcnWs &ws) {
    const size_t es = dtype == RVSR_F16 ? 2 : 4;
    const size_t po = (size_t)d.Ho * d.Wo;
    ws.x_c8 = cv.take((size_t)d.B * cdiv(d.C, 8) * d.H * d.W * 8 * es);
    ws.off32 = ws.mask32 = ws.w32 = ws.b32 = nullptr;
    if (dtype == RVSR_F16) {
        ws.off32 = (float *)cv.take((size_t)d.B * d.dg * 2 * d.K * po * 4);
        ws.mask32 = (float *)cv.take((size_t)d.B * d.dg * d.K * po * 4);
        ws.w32 = (float *)cv.take((size_t)d.Cout * (d.C / d.groups) * d.K * 4);
        ws.b32 = (float *)cv.take((size_t)d.Cout * 4);
    }
    ws.wdense = d.groups > 1 ? (float *)cv.take((size_t)d.Cout * d.C * d.K * 4) : nullptr;
    ws.wpack = (float *)cv.take((size_t)cdiv(d.C, 8) * d.K * 8 * cdiv(d.Cout, 64) * 64 * 4);
}

bool mdcn_fwd_tc_enabled() {  // RVSR_MDCN_FWD_TC=0: 16-bit rvsr_mdcn_fwd stays on the CUDA-core kernel (A/B runs)
    static const bool off = getenv("RVSR_MDCN_FWD_TC") != nullptr && getenv("RVSR_MDCN_FWD_TC")[0] == '0';
    return !off;
}

template <typename T>
int mdcn_fwd_t(const DcnDims &d, const void *input, const void *offset, const void *mask, const void *weight,
               const void *bias, void *output, int dtype, const DcnWs &ws, int act, cudaStream_t s) {
    const size_t po = (size_t)d.Ho * d.Wo;
    RVSR_TRY((launch_pack_nchw<T, T>((const T *)input, (T *)ws.x_c8, d.B, d.C, d.H, d.W, s)));
    const float *off = (const float *)offset, *msk = (const float *)mask, *w = (const float *)weight,
                *b = (const float *)bias;
    if (dtype == RVSR_F16) {
        RVSR_TRY(launch_convert_f16_f32(offset, ws.off32, (long long)d.B * d.dg * 2 * d.K * po, s));
        RVSR_TRY(launch_convert_f16_f32(mask, ws.mask32, (long long)d.B * d.dg * d.K * po, s));
        RVSR_TRY(launch_convert_f16_f32(weight, ws.w32, (long long)d.Cout * (d.C / d.groups) * d.K, s));
        if (bias) RVSR_TRY(launch_convert_f16_f32(bias, ws.b32, d.Cout, s));
        off = ws.off32; msk = ws.mask32; w = ws.w32; b = bias ? ws.b32 : nullptr;
    }
    if (d.groups > 1) {
        RVSR_TRY(expand_grouped_weight(w, ws.wdense, d.Cout, d.C, d.K, d.groups, s));
        w = ws.wdense;
    }
    const int cout_pad = cdiv(d.Cout, 64) * 64, cin = d.C;
    RVSR_TRY(pack_weight_simt(w, ws.wpack, d.Cout, d.C, d.kh, &cin, 1, cout_pad, s));
    DcnOp op = {};
    op.x.ptr = ws.x_c8; op.x.image_stride = (long long)cdiv(d.C, 8) * d.H * d.W * 8; op.x.C = d.C;
    op.x.frames = 1; op.x.fixed_frame = -1;
    op.offset = off; op.mask = msk;
    op.offset_image_stride = (long long)d.dg * 2 * d.K * po;
    op.mask_image_stride = (long long)d.dg * d.K * po;
    op.w_simt = ws.wpack; op.bias = b; op.out = output; op.out_image_stride = (long long)d.Cout * po;
    op.N = d.B; op.H = d.H; op.W = d.W; op.Cout = d.Cout; op.kh = d.kh; op.kw = d.kw; op.stride = d.stride;
    op.pad = d.pad; op.dil = d.dil; op.dg = d.dg; op.act = act; op.out_mode = OUT_NCHW_T;
    return launch_dcn_simt<T>(op, s);
}
}  // namespace
}  // namespace rvsr

using namespace rvsr;

extern "C" {

int rvsr_version(void) { return 100; }
const char *rvsr_last_error(void) { return get_error(); }

int rvsr_device_ok(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        set_error("no usable CUDA device");
        return RVSR_E_CUDA;
    }
    return p.major == 10 ? 1 : 0;
}

size_t rvsr_mdcn_fwd_workspace_bytes(int B, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad,
                                     int dil, int groups, int dg, int dtype) {
    DcnDims d;
    if (dcn_dims(d, B, C, H, W, Cout, kh, kw, stride, pad, dil, groups, dg) != RVSR_OK) return 0;
    Carver cv{nullptr, 0};
    DcnWs ws;
    size_t tc_bytes = 0;
    if ((dtype == RVSR_F16 || dtype == RVSR_BF16) && C == 64 && Cout == 64 && kh == 3 && kw == 3 && dg == 8) {  // tcgen05 path of rvsr_mdcn_fwd
        Carver ct{nullptr, 0};
        const size_t px = (size_t)B * H * W;
        for (size_t n : {px * 64 * 2, px * 64 * 2, px * 8 * 96, (size_t)64 * 64 * 9 * 4, (size_t)64 * 4, tc_dcn_weight_bytes(64, 64, 9) + 16}) ct.take(n);
        tc_bytes = ct.off + 256;
    }
    if (dtype == RVSR_BF16) {  // fp32 copies of the six tensors (see rvsr_mdcn_fwd)
        const size_t po = (size_t)d.Ho * d.Wo;
        for (size_t n : {(size_t)d.B * d.C * d.H * d.W, (size_t)d.B * d.dg * 2 * d.K * po, (size_t)d.B * d.dg * d.K * po,
                         (size_t)d.Cout * (d.C / d.groups) * d.K, (size_t)d.Cout, (size_t)d.B * d.Cout * po})
            cv.take(n * 4);
        dtype = RVSR_F32;
    }
    carve_dcn(cv, d, dtype, ws);
    return (cv.off > tc_bytes ? cv.off : tc_bytes) + 256;
}

int rvsr_mdcn_fwd(const void *input, const void *offset, const void *mask, const void *weight, const void *bias,
                  void *output, int B, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil,
                  int groups, int dg, int dtype, void *workspace, size_t workspace_bytes, void *stream) {
    DcnDims d;
    RVSR_TRY(dcn_dims(d, B, C, H, W, Cout, kh, kw, stride, pad, dil, groups, dg));
    RVSR_CHECK_ARG(dtype == RVSR_F32 || dtype == RVSR_F16 || dtype == RVSR_BF16, "dcn: bad dtype %d", dtype);
    if (B == 0) return RVSR_OK;
    RVSR_CHECK_ARG(input && offset && mask && weight && output && workspace, "dcn: null buffer");
    Carver cv{(char *)workspace, workspace_bytes};
    const size_t mis = (size_t)((uintptr_t)workspace % 256);
    if (mis) { cv.base += 256 - mis; cv.cap -= 256 - mis; }
    cudaStream_t s = (cudaStream_t)stream;
    // EDVR's shape class with 16-bit tensors: gather -> UMMA on tcgen05 (dcn_tc_kernel; fp16 operands, fp32 coordinates and
    // accumulate -- a bf16 caller's values are exactly representable up to fp16's range, which activations do not leave)
    if ((dtype == RVSR_F16 || dtype == RVSR_BF16) && C == 64 && Cout == 64 && kh == 3 && kw == 3 && stride == 1 && pad == 1 && dil == 1 &&
        groups == 1 && dg == 8 && mdcn_fwd_tc_enabled()) {
        const size_t px = (size_t)B * H * W;
        __half *x8 = (__half *)cv.take(px * 64 * 2), *y8 = (__half *)cv.take(px * 64 * 2);
        void *om = cv.take(px * 8 * 96);
        float *w32 = (float *)cv.take((size_t)64 * 64 * 9 * 4), *b32 = (float *)cv.take(64 * 4);
        void *wtc = cv.take(tc_dcn_weight_bytes(64, 64, 9) + 16);
        if (!cv.ok) { set_error("dcn: workspace too small (%zu bytes)", workspace_bytes); return RVSR_E_WORKSPACE; }
        if (dtype == RVSR_F16) {
            RVSR_TRY((launch_pack_nchw<__half, __half>((const __half *)input, x8, B, 64, H, W, s)));
            RVSR_TRY(launch_convert_f16_f32(weight, w32, 64 * 64 * 9, s));
            if (bias) RVSR_TRY(launch_convert_f16_f32(bias, b32, 64, s));
        } else {
            RVSR_TRY((launch_pack_nchw<__half, __nv_bfloat16>((const __nv_bfloat16 *)input, x8, B, 64, H, W, s)));
            RVSR_TRY(launch_convert_bf16_f32(weight, w32, 64 * 64 * 9, s));
            if (bias) RVSR_TRY(launch_convert_bf16_f32(bias, b32, 64, s));
        }
        RVSR_TRY(launch_om24_from_planar(offset, mask, dtype, om, B, 8, H, W, s));
        RVSR_TRY(pack_weight_dcn_tc(w32, wtc, 64, 64, 9, s));
        DcnOp op = {};
        op.x = Src{x8, (long long)64 * H * W, 64, 1, -1}; op.om24 = om; op.om24_image_stride = (long long)8 * 24 * H * W;
        op.w_tc = wtc; op.bias = bias ? b32 : nullptr; op.out = y8; op.out_image_stride = (long long)64 * H * W;
        op.N = B; op.H = H; op.W = W; op.Cout = 64; op.kh = op.kw = 3; op.stride = 1; op.pad = 1; op.dil = 1; op.dg = 8;
        op.act = RVSR_ACT_NONE; op.out_mode = OUT_C8;
        if (tc_dcn_supported(op)) {
            RVSR_TRY(launch_dcn_tc(op, s));
            if (dtype == RVSR_F16) return launch_unpack_nchw<__half, __half>(y8, (__half *)output, B, 64, H, W, s);
            return launch_unpack_nchw<__half, __nv_bfloat16>(y8, (__nv_bfloat16 *)output, B, 64, H, W, s);
        }
        cv.off = 0;  // not covered after all (no tensor-map encoder): fall through to the CUDA-core paths
    }
    if (dtype == RVSR_BF16) {  // bfloat16 tensors, fp32 arithmetic: widen, run the fp32 operator, round the result once
        const size_t po = (size_t)d.Ho * d.Wo;
        const long long n_in = (long long)d.B * d.C * d.H * d.W, n_off = (long long)d.B * d.dg * 2 * d.K * po, n_msk = n_off / 2,
                        n_w = (long long)d.Cout * (d.C / d.groups) * d.K, n_out = (long long)d.B * d.Cout * po;
        float *in32 = (float *)cv.take(n_in * 4), *off32 = (float *)cv.take(n_off * 4), *msk32 = (float *)cv.take(n_msk * 4);
        float *w32 = (float *)cv.take(n_w * 4), *b32 = (float *)cv.take((size_t)d.Cout * 4), *out32 = (float *)cv.take(n_out * 4);
        DcnWs ws;
        carve_dcn(cv, d, RVSR_F32, ws);
        if (!cv.ok) { set_error("dcn: workspace too small (%zu bytes)", workspace_bytes); return RVSR_E_WORKSPACE; }
        RVSR_TRY(launch_convert_bf16_f32(input, in32, n_in, s));
        RVSR_TRY(launch_convert_bf16_f32(offset, off32, n_off, s));
        RVSR_TRY(launch_convert_bf16_f32(mask, msk32, n_msk, s));
        RVSR_TRY(launch_convert_bf16_f32(weight, w32, n_w, s));
        if (bias) RVSR_TRY(launch_convert_bf16_f32(bias, b32, d.Cout, s));
        RVSR_TRY(mdcn_fwd_t<float>(d, in32, off32, msk32, w32, bias ? b32 : nullptr, out32, RVSR_F32, ws, RVSR_ACT_NONE, s));
        return launch_convert_f32_bf16(out32, output, n_out, s);
    }
    DcnWs ws;
    carve_dcn(cv, d, dtype, ws);
    if (!cv.ok) { set_error("dcn: workspace too small (%zu bytes)", workspace_bytes); return RVSR_E_WORKSPACE; }
    if (dtype == RVSR_F16)
        return mdcn_fwd_t<__half>(d, input, offset, mask, weight, bias, output, dtype, ws, RVSR_ACT_NONE, s);
    return mdcn_fwd_t<float>(d, input, offset, mask, weight, bias, output, dtype, ws, RVSR_ACT_NONE, s);
}

static size_t mdcn_bwd_ws(const DcnDims &d) {
    size_t n = 2 * align_up((size_t)d.B * cdiv(d.C, 8) * d.H * d.W * 8 * 4, 256);   // x and grad_x channel-blocked
    n += 2 * align_up((size_t)d.Cout * d.C * d.K * 4, 256);                        // dense weight + dense grad
    return n + 1024;
}
size_t rvsr_mdcn_bwd_workspace_bytes(int B, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil,
                                     int groups, int dg, int dtype) {
    DcnDims d;
    if ((dtype != RVSR_F32 && dtype != RVSR_BF16) || dcn_dims(d, B, C, H, W, Cout, kh, kw, stride, pad, dil, groups, dg) != RVSR_OK) return 0;
    if (dtype == RVSR_BF16) {  // fp32 copies of the five inputs and the five gradients (same carving as rvsr_mdcn_bwd)
        const size_t po = (size_t)d.Ho * d.Wo;
        const size_t n_in = (size_t)d.B * d.C * d.H * d.W, n_off = (size_t)d.B * d.dg * 2 * d.K * po, n_w = (size_t)d.Cout * (d.C / d.groups) * d.K;
        Carver cv{nullptr, 0};
        for (size_t n : {n_in, n_off, n_off / 2, n_w, (size_t)d.B * d.Cout * po, n_in, n_off, n_off / 2, n_w, (size_t)d.Cout}) cv.take(n * 4);
        const size_t simt = cv.off + mdcn_bwd_ws(d) + 1024, tc = dcn_bwd_tc_workspace_bytes(d.B, d.H, d.W) + 512;
        return simt > tc ? simt : tc;
    }
    return mdcn_bwd_ws(d);
}
int rvsr_mdcn_bwd(const void *input, const void *offset, const void *mask, const void *weight, const void *grad_output,
                  void *grad_input, void *grad_offset, void *grad_mask, void *grad_weight, void *grad_bias, int B, int C,
                  int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil, int groups, int dg, int dtype,
                  void *workspace, size_t workspace_bytes, void *stream) {
    DcnDims d;
    RVSR_TRY(dcn_dims(d, B, C, H, W, Cout, kh, kw, stride, pad, dil, groups, dg));
    if (dtype != RVSR_F32 && dtype != RVSR_BF16) {
        set_error("rvsr_mdcn_bwd: fp32 or bf16 tensors (gradients are computed in fp32; convert fp16 on the host side)");
        return RVSR_E_UNSUPPORTED;
    }
    if (B == 0) return RVSR_OK;
    RVSR_CHECK_ARG(input && offset && mask && weight && grad_output && grad_input && grad_offset && grad_mask &&
                       grad_weight && workspace, "dcn bwd: null buffer");
    if (dtype == RVSR_BF16 && dcn_bwd_tc_supported(C, Cout, kh, kw, stride, pad, dil, groups, dg) &&
        workspace_bytes >= dcn_bwd_tc_workspace_bytes(B, H, W) + 256)
        // EDVR's shape class: grad_col / grad_weight contractions on tcgen05, bf16 operands (dcn_bwd_tc.cu)
        return launch_dcn_bwd_tc(input, offset, mask, weight, grad_output, grad_input, grad_offset, grad_mask, grad_weight, grad_bias, B, H, W,
                                 workspace, workspace_bytes, (cudaStream_t)stream);
    if (dtype == RVSR_BF16) {  // other shapes: widen the inputs, run the fp32 operator on fp32 scratch, round every gradient once
        Carver cb{(char *)workspace, workspace_bytes};
        const size_t misb = (size_t)((uintptr_t)workspace % 256);
        if (misb) { cb.base += 256 - misb; cb.cap -= 256 - misb; }
        const size_t po = (size_t)d.Ho * d.Wo;
        const long long n_in = (long long)d.B * d.C * d.H * d.W, n_off = (long long)d.B * d.dg * 2 * d.K * po, n_msk = n_off / 2,
                        n_w = (long long)d.Cout * (d.C / d.groups) * d.K, n_out = (long long)d.B * d.Cout * po;
        float *in32 = (float *)cb.take(n_in * 4), *off32 = (float *)cb.take(n_off * 4), *msk32 = (float *)cb.take(n_msk * 4);
        float *w32 = (float *)cb.take(n_w * 4), *go32 = (float *)cb.take(n_out * 4);
        float *gin32 = (float *)cb.take(n_in * 4), *goff32 = (float *)cb.take(n_off * 4), *gmsk32 = (float *)cb.take(n_msk * 4);
        float *gw32 = (float *)cb.take(n_w * 4), *gb32 = (float *)cb.take((size_t)d.Cout * 4);
        if (!cb.ok || cb.off + mdcn_bwd_ws(d) > cb.cap) { set_error("dcn bwd: workspace too small"); return RVSR_E_WORKSPACE; }
        cudaStream_t sb = (cudaStream_t)stream;
        RVSR_TRY(launch_convert_bf16_f32(input, in32, n_in, sb));
        RVSR_TRY(launch_convert_bf16_f32(offset, off32, n_off, sb));
        RVSR_TRY(launch_convert_bf16_f32(mask, msk32, n_msk, sb));
        RVSR_TRY(launch_convert_bf16_f32(weight, w32, n_w, sb));
        RVSR_TRY(launch_convert_bf16_f32(grad_output, go32, n_out, sb));
        // grad_weight / grad_bias: summed in fp32 scratch and WRITTEN (a bf16 accumulator would round every partial sum);
        // for the caller-zeroed tensors the contract asks for, that is the same result
        RVSR_CUDA(cudaMemsetAsync(gw32, 0, n_w * 4, sb));
        RVSR_CUDA(cudaMemsetAsync(gb32, 0, (size_t)d.Cout * 4, sb));
        RVSR_TRY(rvsr_mdcn_bwd(in32, off32, msk32, w32, go32, gin32, goff32, gmsk32, gw32, grad_bias ? gb32 : nullptr, B, C, H, W, Cout, kh, kw,
                               stride, pad, dil, groups, dg, RVSR_F32, cb.base + cb.off, cb.cap - cb.off, stream));
        RVSR_TRY(launch_convert_f32_bf16(gin32, grad_input, n_in, sb));
        RVSR_TRY(launch_convert_f32_bf16(goff32, grad_offset, n_off, sb));
        RVSR_TRY(launch_convert_f32_bf16(gmsk32, grad_mask, n_msk, sb));
        RVSR_TRY(launch_convert_f32_bf16(gw32, grad_weight, n_w, sb));
        if (grad_bias) RVSR_TRY(launch_convert_f32_bf16(gb32, grad_bias, d.Cout, sb));
        return RVSR_OK;
    }
    Carver cv{(char *)workspace, workspace_bytes};
    const size_t mis = (size_t)((uintptr_t)workspace % 256);
    if (mis) { cv.base += 256 - mis; cv.cap -= 256 - mis; }
    const size_t c8e = (size_t)d.B * cdiv(d.C, 8) * d.H * d.W * 8;
    float *x8 = (float *)cv.take(c8e * 4), *gx8 = (float *)cv.take(c8e * 4);
    float *wd = (float *)cv.take((size_t)d.Cout * d.C * d.K * 4), *gwd = (float *)cv.take((size_t)d.Cout * d.C * d.K * 4);
    if (!cv.ok) { set_error("dcn bwd: workspace too small"); return RVSR_E_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t po = (size_t)d.Ho * d.Wo;
    RVSR_TRY((launch_pack_nchw<float, float>((const float *)input, x8, d.B, d.C, d.H, d.W, s)));
    RVSR_CUDA(cudaMemsetAsync(gx8, 0, c8e * 4, s));
    RVSR_CUDA(cudaMemsetAsync(gwd, 0, (size_t)d.Cout * d.C * d.K * 4, s));
    RVSR_CUDA(cudaMemsetAsync(grad_offset, 0, (size_t)d.B * d.dg * 2 * d.K * po * 4, s));
    RVSR_CUDA(cudaMemsetAsync(grad_mask, 0, (size_t)d.B * d.dg * d.K * po * 4, s));
    const float *w = (const float *)weight;
    if (d.groups > 1) {
        RVSR_TRY(expand_grouped_weight(w, wd, d.Cout, d.C, d.K, d.groups, s));
        w = wd;
    }
    DcnBwdOp op = {};
    op.x_c8 = x8; op.offset = (const float *)offset; op.mask = (const float *)mask; op.gout = (const float *)grad_output;
    op.w_dense = w; op.gx_c8 = gx8; op.goffset = (float *)grad_offset; op.gmask = (float *)grad_mask; op.gw_dense = gwd;
    op.gbias = (float *)grad_bias;
    op.N = d.B; op.C = d.C; op.H = d.H; op.W = d.W; op.Cout = d.Cout; op.kh = d.kh; op.kw = d.kw; op.stride = d.stride;
    op.pad = d.pad; op.dil = d.dil; op.dg = d.dg;
    RVSR_TRY(launch_dcn_bwd_simt(op, s));
    RVSR_TRY(fold_grouped_weight(gwd, (float *)grad_weight, d.Cout, d.C, d.K, d.groups, s));
    return launch_unpack_nchw<float, float>(gx8, (float *)grad_input, d.B, d.C, d.H, d.W, s);
}

size_t rvsr_mdcn_pack_fwd_workspace_bytes(int B, int C, int H, int W, int Cout, int dg, int dtype) {
    if (B < 0 || C <= 0 || H <= 0 || W <= 0 || Cout <= 0 || dg <= 0) return 0;
    const size_t es = dtype == RVSR_F16 ? 2 : 4, px = (size_t)B * H * W, c8 = (size_t)cdiv(C, 8) * 8;
    size_t n = 2 * align_up(px * c8 * es, 256);                       // x, feat channel-blocked
    n += align_up(px * 27 * dg * 4, 256);                             // offsets + mask (planar fp32 >= OM24)
    n += align_up(px * (size_t)cdiv(Cout, 8) * 8 * es, 256);          // output channel-blocked
    n += 2 * align_up((size_t)27 * dg * C * 9 * 4, 256) + 2 * align_up((size_t)Cout * C * 9 * 4, 256);  // fp32 copies
    n += align_up((size_t)cdiv(C, 8) * 72 * cdiv(27 * dg, 64) * 64 * 4, 256) + align_up((size_t)cdiv(C, 8) * 72 * cdiv(Cout, 64) * 64 * 4, 256);
    n += 2 * align_up(tc_conv_weight_bytes(27 * dg, C, 3, 2) + 256, 256) + align_up(tc_conv_weight_bytes(Cout, C, 3) + 256, 256);
    n += 2 * align_up((size_t)(27 * dg + Cout) * 4, 256);
    return n + 8192;
}

}  // extern "C"

template <typename T>
static int mdcn_pack_fwd_t(const void *x, const void *feat, const void *w_om, const void *b_om, const void *weight,
                           const void *bias, void *y, int B, int C, int H, int W, int Cout, int dg, int act, int dtype,
                           Carver &cv, cudaStream_t s) {
    const size_t es = sizeof(T), px = (size_t)B * H * W;
    const int Com = 27 * dg, C8 = cdiv(C, 8);
    T *xa = (T *)cv.take(px * C8 * 8 * es), *fa = (T *)cv.take(px * C8 * 8 * es);
    float *om = (float *)cv.take(px * Com * 4);
    T *oa = (T *)cv.take(px * cdiv(Cout, 8) * 8 * es);
    float *wom32 = (float *)cv.take((size_t)Com * C * 9 * 4), *w32 = (float *)cv.take((size_t)Cout * C * 9 * 4);
    float *bom32 = (float *)cv.take((size_t)Com * 4), *b32 = (float *)cv.take((size_t)Cout * 4);
    const int cp_om = cdiv(Com, 64) * 64, cp = cdiv(Cout, 64) * 64;
    float *wom_simt = (float *)cv.take((size_t)C8 * 72 * cp_om * 4), *w_simt = (float *)cv.take((size_t)C8 * 72 * cp * 4);
    const bool tc = dtype == RVSR_F16 && C == 64 && Cout == 64 && (C / dg) % 8 == 0 && tc_conv_weight_bytes(Com, C, 3, 2) > 0;
    void *wom_tc = tc ? cv.take(tc_conv_weight_bytes(Com, C, 3, 2) + 16) : nullptr;
    void *w_tc = tc ? cv.take(tc_conv_weight_bytes(Cout, C, 3) + 16) : nullptr;
    void *wom_tc2 = (tc && tc2_weight_bytes(Com, C, 3, 2) > 0) ? cv.take(tc2_weight_bytes(Com, C, 3, 2) + 16) : nullptr;
    if (!cv.ok) { set_error("mdcn_pack: workspace too small"); return RVSR_E_WORKSPACE; }
    RVSR_TRY((launch_pack_nchw<T, T>((const T *)x, xa, B, C, H, W, s)));
    RVSR_TRY((launch_pack_nchw<T, T>((const T *)feat, fa, B, C, H, W, s)));
    const float *wom = (const float *)w_om, *bom = (const float *)b_om, *w = (const float *)weight, *b = (const float *)bias;
    if (dtype == RVSR_F16) {
        RVSR_TRY(launch_convert_f16_f32(w_om, wom32, (long long)Com * C * 9, s)); wom = wom32;
        RVSR_TRY(launch_convert_f16_f32(weight, w32, (long long)Cout * C * 9, s)); w = w32;
        if (b_om) { RVSR_TRY(launch_convert_f16_f32(b_om, bom32, Com, s)); bom = bom32; }
        if (bias) { RVSR_TRY(launch_convert_f16_f32(bias, b32, Cout, s)); b = b32; }
    }
    const int cin = C;
    const long long img = (long long)C8 * H * W * 8;
    ConvOp co = {};
    co.src[0] = Src{fa, img, C, 1, -1}; co.nsrc = 1; co.bias = bom; co.out = om;
    co.N = B; co.H = H; co.W = W; co.Cout = Com; co.ks = 3; co.stride = 1; co.act = RVSR_ACT_NONE; co.sig_from = 18 * dg;
    DcnOp d = {};
    d.x = Src{xa, img, C, 1, -1}; d.bias = b; d.N = B; d.H = H; d.W = W; d.Cout = Cout; d.kh = d.kw = 3; d.stride = 1;
    d.pad = 1; d.dil = 1; d.dg = dg; d.act = act;
    if (tc && dg == 8 && bom != nullptr && tc_pack_om_weight_bytes(Com, C, dg) > 0) {
        // ONE kernel: offset/mask conv -> TMEM -> gather -> contraction (dcn_fused.cu); the OM tensor never exists.
        // The workspace slots of the two-kernel path's weight layouts hold this kernel's (same sizes).
        PackFusedOp f = {};
        f.x = Src{xa, img, C, 1, -1}; f.feat = fa; f.w_om = wom_tc; f.bias_om = bom; f.w_dcn2 = w_tc; f.bias = b;
        f.out = oa; f.out_image_stride = (long long)cdiv(Cout, 8) * H * W * 8;
        f.N = B; f.H = H; f.W = W; f.Cout = Cout; f.dg = dg; f.act = act;
        if (tc_pack_fused_supported(f)) {
            RVSR_TRY(pack_weight_om_stream(wom, wom_tc, Com, C, dg, s));
            RVSR_TRY(pack_weight_tc2(w, w_tc, Cout, C, 3, 0, s));
            RVSR_TRY(launch_pack_fused(f, s));
            return launch_unpack_nchw<T, T>(oa, (T *)y, B, Cout, H, W, s);
        }
    }
    if (tc) {
        RVSR_TRY(pack_weight_tc(wom, wom_tc, Com, C, 3, 2, s));
        RVSR_TRY(pack_weight_tc(w, w_tc, Cout, C, 3, 0, s));
        if (wom_tc2 != nullptr) RVSR_TRY(pack_weight_tc2(wom, wom_tc2, Com, C, 3, 2, s));
        co.w_tc = wom_tc; co.w_tc2 = wom_tc2; co.out_mode = OUT_OM24; co.dg = dg; co.out_image_stride = (long long)dg * 24 * H * W;
        RVSR_CHECK_ARG(tc_conv_supported(co), "mdcn_pack: offset conv not covered by the tcgen05 kernel");
        RVSR_TRY(launch_conv_tc(co, s));
        d.w_tc = w_tc; d.om24 = om; d.om24_image_stride = co.out_image_stride; d.out = oa;
        d.out_image_stride = (long long)cdiv(Cout, 8) * H * W * 8; d.out_mode = OUT_C8;
        RVSR_CHECK_ARG(tc_dcn_supported(d), "mdcn_pack: DCN not covered by the tcgen05 kernel");
        RVSR_TRY(launch_dcn_tc(d, s));
        return launch_unpack_nchw<T, T>(oa, (T *)y, B, Cout, H, W, s);
    }
    RVSR_TRY(pack_weight_simt(wom, wom_simt, Com, C, 3, &cin, 1, cp_om, s));
    RVSR_TRY(pack_weight_simt(w, w_simt, Cout, C, 3, &cin, 1, cp, s));
    co.w_simt = wom_simt; co.out_mode = OUT_PLANAR_F32; co.out_image_stride = (long long)Com * H * W;
    RVSR_TRY(launch_conv_simt<T>(co, s));
    d.w_simt = w_simt; d.offset = om; d.mask = om + (long long)18 * dg * H * W;
    d.offset_image_stride = d.mask_image_stride = (long long)Com * H * W;
    d.out = y; d.out_image_stride = (long long)Cout * H * W; d.out_mode = OUT_NCHW_T;
    return launch_dcn_simt<T>(d, s);
}

extern "C" {

int rvsr_mdcn_pack_fwd(const void *x, const void *feat, const void *w_offset_mask, const void *b_offset_mask,
                       const void *weight, const void *bias, void *y, int B, int C, int H, int W, int Cout, int dg,
                       int act, int dtype, void *workspace, size_t workspace_bytes, void *stream) {
    RVSR_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0 && Cout > 0, "mdcn_pack: bad sizes");
    RVSR_CHECK_ARG(dg > 0 && C % dg == 0, "mdcn_pack: channels %d not divisible by deformable groups %d", C, dg);
    RVSR_CHECK_ARG(dtype == RVSR_F32 || dtype == RVSR_F16, "mdcn_pack: bad dtype");
    RVSR_CHECK_ARG(act == RVSR_ACT_NONE || act == RVSR_ACT_LRELU || act == RVSR_ACT_RELU, "mdcn_pack: bad activation");
    if (B == 0) return RVSR_OK;
    RVSR_CHECK_ARG(x && feat && w_offset_mask && weight && y && workspace, "mdcn_pack: null buffer");
    Carver cv{(char *)workspace, workspace_bytes};
    const size_t mis = (size_t)((uintptr_t)workspace % 256);
    if (mis) { cv.base += 256 - mis; cv.cap -= 256 - mis; }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == RVSR_F16)
        return mdcn_pack_fwd_t<__half>(x, feat, w_offset_mask, b_offset_mask, weight, bias, y, B, C, H, W, Cout, dg, act,
                                       dtype, cv, s);
    return mdcn_pack_fwd_t<float>(x, feat, w_offset_mask, b_offset_mask, weight, bias, y, B, C, H, W, Cout, dg, act, dtype,
                                  cv, s);
}

size_t rvsr_conv2d_fwd_workspace_bytes(int B, int C1, int C2, int H, int W, int Cout, int ks, int dtype) {
    const size_t es = dtype == RVSR_F16 ? 2 : 4;
    const size_t px = (size_t)B * H * W;
    const int Cin = C1 + C2;
    size_t n = 0;
    n += align_up(px * (size_t)cdiv(C1 > 16 ? C1 : 16, 8) * 8 * es, 256) + align_up(px * (size_t)cdiv(C2 > 0 ? C2 : 1, 8) * 8 * es, 256);
    n += 2 * align_up(px * 4 * (size_t)cdiv(Cout, 8) * 8 * es, 256);            // out (maybe shuffled) + residual
    n += 2 * align_up((size_t)Cout * (Cin + 16) * ks * ks * 4, 256);             // fp32 weight copy + padded copy
    n += align_up((size_t)cdiv(Cin + 16, 8) * ks * ks * 8 * cdiv(Cout, 64) * 64 * 4, 256);  // simt pack
    n += 2 * align_up(tc_conv_weight_bytes(Cout, Cin + 16, ks) + 256, 256) + align_up((size_t)Cout * 4, 256);
    return n + 4096;
}

}  // extern "C"

template <typename T>
static int conv2d_fwd_t(const void *x1, const void *x2, const void *weight, const void *bias, const void *residual,
                        void *y, int B, int C1, int C2, int H, int W, int Cout, int ks, int stride, int act,
                        int shuffle, int dtype, int use_tc, Carver &cv, cudaStream_t s) {
    const size_t es = sizeof(T);
    const int Cin = C1 + C2, KK = ks * ks;
    const int Ho = stride == 1 ? H : (H - 1) / 2 + 1, Wo = stride == 1 ? W : (W - 1) / 2 + 1;
    const bool tc = use_tc && dtype == RVSR_F16;
    const int C1s = (tc && C2 == 0 && C1 < 16) ? 16 : C1;   // tensor-core K granularity (see engine.cu finalize)
    const size_t px = (size_t)B * H * W, pxo = (size_t)B * Ho * Wo;
    T *a1 = (T *)cv.take(px * cdiv(C1s, 8) * 8 * es);
    T *a2 = C2 > 0 ? (T *)cv.take(px * cdiv(C2, 8) * 8 * es) : nullptr;
    const int Cst = shuffle ? Cout / 4 : Cout;
    const size_t out_elems = (shuffle ? pxo * 4 : pxo) * cdiv(Cst, 8) * 8;
    T *o = (T *)cv.take(out_elems * es);
    T *r = residual ? (T *)cv.take(out_elems * es) : nullptr;
    float *w32 = (float *)cv.take((size_t)Cout * Cin * KK * 4);
    float *wpad = (float *)cv.take((size_t)Cout * (Cin + 16) * KK * 4);
    float *b32 = (float *)cv.take((size_t)Cout * 4);
    const int CinS = C1s + C2;
    const int cout_pad = cdiv(Cout, 64) * 64;
    float *wsimt = (float *)cv.take((size_t)cdiv(CinS, 8) * KK * 8 * cout_pad * 4);
    void *wtc = tc ? cv.take(tc_conv_weight_bytes(Cout, CinS, ks) + 16) : nullptr;
    void *wtc2 = (tc && tc2_weight_bytes(Cout, CinS, ks, shuffle ? 1 : 0) > 0) ? cv.take(tc2_weight_bytes(Cout, CinS, ks, shuffle ? 1 : 0) + 16) : nullptr;
    if (!cv.ok) { set_error("conv2d: workspace too small"); return RVSR_E_WORKSPACE; }
    RVSR_TRY((launch_pack_nchw<T, T>((const T *)x1, a1, B, C1, H, W, s, C1s)));
    if (C2 > 0) RVSR_TRY((launch_pack_nchw<T, T>((const T *)x2, a2, B, C2, H, W, s)));
    if (residual) RVSR_TRY((launch_pack_nchw<T, T>((const T *)residual, r, B, Cout, Ho, Wo, s)));
    const float *w = (const float *)weight, *b = (const float *)bias;
    if (dtype == RVSR_F16) {
        RVSR_TRY(launch_convert_f16_f32(weight, w32, (long long)Cout * Cin * KK, s));
        w = w32;
        if (bias) { RVSR_TRY(launch_convert_f16_f32(bias, b32, Cout, s)); b = b32; }
    }
    if (C1s != C1) { RVSR_TRY(pad_weight_cin(w, wpad, Cout, C1, C1s, KK, s)); w = wpad; }
    const int cins = CinS;
    RVSR_CHECK_ARG(C1 % 8 == 0 || C2 == 0, "conv2d: first source must have a multiple of 8 channels when concatenating");
    RVSR_TRY(pack_weight_simt(w, wsimt, Cout, CinS, ks, &cins, 1, cout_pad, s));
    if (tc && tc_conv_weight_bytes(Cout, CinS, ks) > 0) RVSR_TRY(pack_weight_tc(w, wtc, Cout, CinS, ks, shuffle, s));
    if (wtc2 != nullptr) RVSR_TRY(pack_weight_tc2(w, wtc2, Cout, CinS, ks, shuffle ? 1 : 0, s));
    ConvOp op = {};
    op.src[0] = Src{a1, (long long)cdiv(C1s, 8) * H * W * 8, C1s, 1, -1};
    op.nsrc = 1;
    if (C2 > 0) { op.src[1] = Src{a2, (long long)cdiv(C2, 8) * H * W * 8, C2, 1, -1}; op.nsrc = 2; }
    op.w_simt = wsimt; op.w_tc = (tc && tc_conv_weight_bytes(Cout, CinS, ks) > 0) ? wtc : nullptr; op.w_tc2 = wtc2; op.bias = b;
    op.out = o; op.out_image_stride = (long long)(out_elems / B);
    op.residual = r; op.res_image_stride = op.out_image_stride;
    op.N = B; op.H = H; op.W = W; op.Cout = Cout; op.ks = ks; op.stride = stride; op.act = act;
    op.out_mode = shuffle ? OUT_C8_SHUFFLE2 : OUT_C8; op.sig_from = 1 << 30;
    if (tc) {
        if (!tc_conv_supported(op)) { set_error("conv2d: configuration not covered by the tcgen05 kernel"); return RVSR_E_UNSUPPORTED; }
        RVSR_TRY(launch_conv_tc(op, s));
    } else {
        RVSR_TRY(launch_conv_simt<T>(op, s));
    }
    if (shuffle) return launch_unpack_nchw<T, T>(o, (T *)y, B, Cst, 2 * Ho, 2 * Wo, s);
    return launch_unpack_nchw<T, T>(o, (T *)y, B, Cout, Ho, Wo, s);
}

extern "C" {

int rvsr_conv2d_fwd(const void *x1, const void *x2, const void *weight, const void *bias, const void *residual, void *y,
                    int B, int C1, int C2, int H, int W, int Cout, int ks, int stride, int act, int shuffle, int dtype,
                    int use_tc, void *workspace, size_t workspace_bytes, void *stream) {
    RVSR_CHECK_ARG(B >= 0 && C1 > 0 && C2 >= 0 && H > 0 && W > 0 && Cout > 0, "conv2d: bad sizes");
    RVSR_CHECK_ARG(ks == 1 || ks == 3, "conv2d: kernel size %d not built", ks);
    RVSR_CHECK_ARG(stride == 1 || stride == 2, "conv2d: stride %d not built", stride);
    RVSR_CHECK_ARG(dtype == RVSR_F32 || dtype == RVSR_F16, "conv2d: bad dtype");
    RVSR_CHECK_ARG(!shuffle || (Cout % 4 == 0 && residual == nullptr), "conv2d: pixel-shuffle needs Cout %% 4 == 0");
    if (B == 0) return RVSR_OK;
    RVSR_CHECK_ARG(x1 && weight && y && workspace, "conv2d: null buffer");
    Carver cv{(char *)workspace, workspace_bytes};
    const size_t mis = (size_t)((uintptr_t)workspace % 256);
    if (mis) { cv.base += 256 - mis; cv.cap -= 256 - mis; }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == RVSR_F16)
        return conv2d_fwd_t<__half>(x1, x2, weight, bias, residual, y, B, C1, C2, H, W, Cout, ks, stride, act, shuffle,
                                    dtype, use_tc, cv, s);
    return conv2d_fwd_t<float>(x1, x2, weight, bias, residual, y, B, C1, C2, H, W, Cout, ks, stride, act, shuffle, dtype,
                               0, cv, s);
}

int rvsr_engine_create(const rvsr_edvr_config *cfg, rvsr_engine **out) {
    RVSR_CHECK_ARG(cfg != nullptr && out != nullptr, "engine_create: null argument");
    RVSR_CHECK_ARG(cfg->nf > 0 && cfg->nc > 0 && cfg->nc <= 8 && cfg->nframes > 0 && cfg->groups > 0,
                   "engine_create: bad config");
    RVSR_CHECK_ARG(cfg->front_RBs >= 0 && cfg->back_RBs >= 0, "engine_create: bad block counts");
    RVSR_CHECK_ARG(cfg->precision == RVSR_F32 || cfg->precision == RVSR_F16, "engine_create: bad precision");
    // predeblur / HR_in only exist in EDVR (upsample = 1); EDVR_NoUp stores and ignores them (EDVR_arch.py:335-339)
    if (cfg->nf % 8 != 0 || cfg->nf % cfg->groups != 0 || cfg->nframes > RVSR_MAX_SRC) {
        set_error("engine_create: needs nf %% 8 == 0, nf %% groups == 0, nframes <= %d", RVSR_MAX_SRC);
        return RVSR_E_UNSUPPORTED;
    }
    if (!cfg->upsample && cfg->nf != 64) {
        set_error("engine_create: EDVR_NoUp needs nf == 64 (HRconv is hard-wired to 64 channels, EDVR_arch.py:348)");
        return RVSR_E_INVALID;
    }
    RVSR_CHECK_ARG(cfg->center < cfg->nframes, "engine_create: center %d >= nframes", cfg->center);
    rvsr_engine *e = new (std::nothrow) rvsr_engine(*cfg);
    RVSR_CHECK_ARG(e != nullptr, "engine_create: out of host memory");
    *out = e;
    return RVSR_OK;
}
void rvsr_engine_destroy(rvsr_engine *e) { delete e; }
int rvsr_engine_set_weight(rvsr_engine *e, const char *name, const float *dev_ptr, const int64_t *shape, int ndim,
                           void *stream) {
    RVSR_CHECK_ARG(e != nullptr, "null engine");
    return e->impl.set_weight(name, dev_ptr, shape, ndim, (cudaStream_t)stream);
}
int rvsr_engine_finalize(rvsr_engine *e, void *stream) {
    RVSR_CHECK_ARG(e != nullptr, "null engine");
    return e->impl.finalize((cudaStream_t)stream);
}
int rvsr_engine_num_weights(const rvsr_engine *e) { return e ? (int)e->impl.names().size() : 0; }
const char *rvsr_engine_weight_name(const rvsr_engine *e, int i) {
    if (e == nullptr || i < 0 || i >= (int)e->impl.names().size()) return nullptr;
    return e->impl.names()[i].c_str();
}
size_t rvsr_engine_workspace_bytes(const rvsr_engine *e, int B, int H, int W) {
    return e ? const_cast<rvsr_engine *>(e)->impl.workspace_bytes(B, H, W) : 0;
}
int rvsr_engine_forward(rvsr_engine *e, const void *x, int x_dtype, void *out, int out_dtype, int B, int H, int W,
                        void *workspace, size_t workspace_bytes, void *stream) {
    RVSR_CHECK_ARG(e != nullptr, "null engine");
    return e->impl.forward(x, x_dtype, out, out_dtype, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
}
int rvsr_engine_forward_host(rvsr_engine *e, const void *x_host, int x_dtype, void *out_host, int out_dtype, int B,
                             int H, int W, void *dev_in, void *dev_out, void *workspace, size_t workspace_bytes,
                             void *stream) {
    RVSR_CHECK_ARG(e != nullptr, "null engine");
    RVSR_CHECK_ARG(x_host && out_host && dev_in && dev_out, "forward_host: null buffer");
    const rvsr_edvr_config &c = e->impl.cfg();
    cudaStream_t s = (cudaStream_t)stream;
    const int sc = c.upsample ? 4 : 1;
    const size_t in_bytes = (size_t)B * c.nframes * c.nc * H * W * (x_dtype == RVSR_F16 ? 2 : 4);
    const size_t out_bytes = (size_t)B * c.nc * H * sc * W * sc * (out_dtype == RVSR_F16 ? 2 : 4);
    RVSR_CUDA(cudaMemcpyAsync(dev_in, x_host, in_bytes, cudaMemcpyHostToDevice, s));
    RVSR_TRY(e->impl.forward(dev_in, x_dtype, dev_out, out_dtype, B, H, W, workspace, workspace_bytes, s));
    RVSR_CUDA(cudaMemcpyAsync(out_host, dev_out, out_bytes, cudaMemcpyDeviceToHost, s));
    return RVSR_OK;
}
size_t rvsr_engine_cache_bytes(const rvsr_engine *e, int n_slots, int H, int W) { return e ? e->impl.cache_bytes(n_slots, H, W) : 0; }
size_t rvsr_engine_extract_workspace_bytes(const rvsr_engine *e, int F, int H, int W) {
    return e ? const_cast<rvsr_engine *>(e)->impl.extract_workspace_bytes(F, H, W) : 0;
}
int rvsr_engine_extract_features(rvsr_engine *e, const void *frames, int dtype, int F, int H, int W, void *cache, int n_slots,
                                 int slot0, void *workspace, size_t workspace_bytes, void *stream) {
    RVSR_CHECK_ARG(e != nullptr, "null engine");
    return e->impl.extract_features(frames, dtype, F, H, W, cache, n_slots, slot0, workspace, workspace_bytes, (cudaStream_t)stream);
}
int rvsr_engine_forward_cached(rvsr_engine *e, const void *cache, int n_slots, const int *window_slots, const void *frames,
                               int x_dtype, void *out, int out_dtype, int B, int H, int W, void *workspace,
                               size_t workspace_bytes, void *stream) {
    RVSR_CHECK_ARG(e != nullptr, "null engine");
    return e->impl.forward_cached(cache, n_slots, window_slots, frames, x_dtype, out, out_dtype, B, H, W, workspace,
                                  workspace_bytes, (cudaStream_t)stream);
}
int rvsr_engine_last_launch_count(const rvsr_engine *e) { return e ? e->impl.last_launches() : 0; }
int rvsr_engine_set_profiling(rvsr_engine *e, int on) {
    RVSR_CHECK_ARG(e != nullptr, "null engine");
    e->impl.set_profiling(on != 0);
    return RVSR_OK;
}
int rvsr_engine_profile_collect(rvsr_engine *e) {
    RVSR_CHECK_ARG(e != nullptr, "null engine");
    return e->impl.prof_collect();
}
int rvsr_engine_profile_entry(const rvsr_engine *e, int i, char *label, int label_cap, float *ms, double *flops,
                              double *bytes) {
    RVSR_CHECK_ARG(e != nullptr && i >= 0 && i < (int)e->impl.prof().size(), "profile_entry: bad index");
    const ProfEntry &p = e->impl.prof()[i];
    if (label != nullptr && label_cap > 0) snprintf(label, (size_t)label_cap, "%s", p.label.c_str());
    if (ms) *ms = p.ms;
    if (flops) *flops = p.flops;
    if (bytes) *bytes = p.bytes;
    return RVSR_OK;
}
int rvsr_frames_from_u8(const void *u8_thwc, void *out_tchw, int T, int C, int H, int W, int reverse_channels, int out_dtype,
                        void *stream) {
    RVSR_CHECK_ARG(T >= 0 && C >= 1 && C <= 4 && H > 0 && W > 0, "frames_from_u8: bad shape T=%d C=%d H=%d W=%d", T, C, H, W);
    RVSR_CHECK_ARG(out_dtype == RVSR_F32 || out_dtype == RVSR_F16, "frames_from_u8: bad dtype %d", out_dtype);
    if (T == 0) return RVSR_OK;
    RVSR_CHECK_ARG(u8_thwc != nullptr && out_tchw != nullptr, "frames_from_u8: null buffer");
    if (out_dtype == RVSR_F32)
        return launch_frames_from_u8<float>((const uint8_t *)u8_thwc, (float *)out_tchw, T, C, H, W, reverse_channels, (cudaStream_t)stream);
    return launch_frames_from_u8<__half>((const uint8_t *)u8_thwc, (__half *)out_tchw, T, C, H, W, reverse_channels, (cudaStream_t)stream);
}
int rvsr_frames_to_u8(const void *in_bchw, int in_dtype, void *u8_bhwc_bgr, int B, int C, int H, int W, int color_mode, void *stream) {
    RVSR_CHECK_ARG(B >= 0 && C == 3 && H > 0 && W > 0, "frames_to_u8: expected [B, 3, H, W], got C=%d", C);
    RVSR_CHECK_ARG(in_dtype == RVSR_F32 || in_dtype == RVSR_F16, "frames_to_u8: bad dtype %d", in_dtype);
    RVSR_CHECK_ARG(color_mode == 0 || color_mode == 1, "frames_to_u8: color_mode must be 0 (RGB) or 1 (YCbCr), got %d", color_mode);
    if (B == 0) return RVSR_OK;
    RVSR_CHECK_ARG(in_bchw != nullptr && u8_bhwc_bgr != nullptr, "frames_to_u8: null buffer");
    if (in_dtype == RVSR_F32)
        return launch_frames_to_u8<float>((const float *)in_bchw, (uint8_t *)u8_bhwc_bgr, B, H, W, color_mode, (cudaStream_t)stream);
    return launch_frames_to_u8<__half>((const __half *)in_bchw, (uint8_t *)u8_bhwc_bgr, B, H, W, color_mode, (cudaStream_t)stream);
}
int rvsr_engine_read_tap(rvsr_engine *e, const char *name, float *dst_dev, size_t dst_elems, void *stream) {
    RVSR_CHECK_ARG(e != nullptr, "null engine");
    return e->impl.read_tap(name, dst_dev, dst_elems, (cudaStream_t)stream);
}

}  // extern "C"
