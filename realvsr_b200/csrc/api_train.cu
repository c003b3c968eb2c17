// api_train.cu -- extern "C" surface of the bf16 training path (rvsr_c8_* in include/rvsr_b200.h): thin argument checks
// around the launchers of train_kernels.cu and the bf16 mode of the tcgen05 convolution kernels (tc_kernels.cu).
#include "engine.cuh"

namespace rvsr {
bool conv_wgrad_tc_supported(int Cin, int ks, int stride);
size_t conv_wgrad_tc_workspace_bytes(int njobs, int N, int H, int W, int Cout);
int launch_conv_wgrad_tc(int njobs, const void *const *x_c8, const long long *x_image_stride, const void *const *g_c8, float *const *gw,
                         float *const *db, const int *cin_total, const int *c0, int N, int H, int W, int Cout, int ks, void *workspace,
                         size_t workspace_bytes, cudaStream_t s);
int launch_act_bwd_c8(const void *g, const void *y, void *out, long long n_elems, int act, cudaStream_t s);
int launch_unshuffle2_act_bwd_c8(const void *g, const void *y, void *out, int N, int C, int H, int W, int act, cudaStream_t s);
int launch_upsample2x_c8(const void *src, void *dst, long long planes, int H, int W, float scale, int backward, cudaStream_t s);
int launch_nchw_to_c8_bf16(const void *src, int src_dtype, void *dst, int N, int C, int H, int W, int planes, cudaStream_t s);
int launch_c8_to_nchw_bf16(const void *src, void *dst, int dst_dtype, int N, int C, int H, int W, int planes, cudaStream_t s);
size_t c8_mdcn_workspace_bytes(int N, int H, int W, int backward);
int c8_mdcn_fwd(const void *x, const void *om, const float *weight, const float *bias, void *y, int N, int H, int W, int act,
                void *workspace, size_t workspace_bytes, cudaStream_t s);
int c8_mdcn_bwd(const void *x, const void *om, const float *weight, const void *g, const void *y, void *gx, void *gom, float *gw,
                float *gb, int N, int H, int W, int act, void *workspace, size_t workspace_bytes, cudaStream_t s);
int launch_tsa_temporal_c8(const void *aligned, const void *emb, const void *emb_ref, void *const *out, float *prob, int B, int N, int H,
                           int W, cudaStream_t s);
int launch_tsa_temporal_bwd_c8(const void *const *gout, const void *aligned, const void *emb, const void *emb_ref, const float *prob,
                               void *g_aligned, void *g_emb, void *g_emb_ref, int B, int N, int H, int W, cudaStream_t s);
int launch_pool_maxavg_c8(const void *src, void *dmax, void *davg, long long planes, int H, int W, cudaStream_t s);
int launch_pool_maxavg_bwd_c8(const void *src, const void *gmax, const void *gavg, void *gin, long long planes, int H, int W, cudaStream_t s);
int launch_tsa_final_c8(const void *fea, const void *att, const void *add, void *out, long long n_elems, cudaStream_t s);
int launch_tsa_final_bwd_c8(const void *g, const void *fea, const void *att, void *g_fea, void *g_att, long long n_elems, cudaStream_t s);
int launch_upsample2x_nchw(const void *src, void *dst, long long planes, int H, int W, float scale, int backward, int dtype, cudaStream_t s);
}  // namespace rvsr

using namespace rvsr;

extern "C" {

int rvsr_c8_from_nchw(const void *src, int src_dtype, void *dst_c8, int N, int C, int H, int W, int planes, void *stream) {
    RVSR_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0, "c8_from_nchw: bad sizes");
    RVSR_CHECK_ARG(N == 0 || (src && dst_c8), "c8_from_nchw: null buffer");
    return launch_nchw_to_c8_bf16(src, src_dtype, dst_c8, N, C, H, W, planes, (cudaStream_t)stream);
}
int rvsr_c8_to_nchw(const void *src_c8, void *dst, int dst_dtype, int N, int C, int H, int W, int planes, void *stream) {
    RVSR_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0, "c8_to_nchw: bad sizes");
    RVSR_CHECK_ARG(N == 0 || (src_c8 && dst), "c8_to_nchw: null buffer");
    return launch_c8_to_nchw_bf16(src_c8, dst, dst_dtype, N, C, H, W, planes, (cudaStream_t)stream);
}

size_t rvsr_c8_conv_weight_bytes(int Cout, int Cin, int ks, int shuffle) {
    const size_t a = tc_conv_weight_bytes(Cout, Cin, ks, shuffle ? 1 : 0);
    if (a == 0) return 0;
    return align_up(a, 256) + align_up(tc2_weight_bytes(Cout, Cin, ks, shuffle ? 1 : 0), 256) + 256;
}

int rvsr_c8_conv_pack_weight(const float *weight, void *dst, int Cout, int Cin, int ks, int shuffle, int mode, int w_cin_total,
                             int w_c0, int layouts, void *stream) {
    RVSR_CHECK_ARG(weight && dst && Cout > 0 && Cin > 0 && (ks == 1 || ks == 3), "c8 pack weight: bad arguments");
    RVSR_CHECK_ARG(mode == 0 || mode == 1, "c8 pack weight: mode %d", mode);
    RVSR_CHECK_ARG(((uintptr_t)dst & 255) == 0, "c8 pack weight: destination must be 256-byte aligned");
    const size_t a = tc_conv_weight_bytes(Cout, Cin, ks, shuffle ? 1 : 0);
    if (a == 0) { set_error("c8 pack weight: shape %d x %d x %d not covered by the tcgen05 kernels", Cout, Cin, ks); return RVSR_E_UNSUPPORTED; }
    const long long KK = ks * ks;
    WeightView wv;
    if (mode == 0) {  // forward: weight[co][w_c0 + cin][tap], rows of w_cin_total input channels
        RVSR_CHECK_ARG(w_c0 >= 0 && w_c0 + Cin <= w_cin_total, "c8 pack weight: channel slice");
        wv = WeightView{w_c0 * KK, w_cin_total * KK, KK, 1, 1};
    } else {          // data gradient of input channels [w_c0, w_c0 + Cout): weight[cin][w_c0 + co][KK - 1 - tap]
        RVSR_CHECK_ARG(w_c0 >= 0 && w_c0 + Cout <= w_cin_total, "c8 pack weight: channel slice");
        wv = WeightView{w_c0 * KK + KK - 1, KK, w_cin_total * KK, -1, 1};
    }
    cudaStream_t s = (cudaStream_t)stream;
    // layouts: bit 0 = single-CTA kernels' layout, bit 1 = CTA-pair layout (the one launches of >= 4 tiles read; 3x3 only)
    if (layouts & 1) RVSR_TRY(pack_weight_tc(weight, dst, Cout, Cin, ks, shuffle ? 1 : 0, s, &wv));
    if ((layouts & 2) && tc2_weight_bytes(Cout, Cin, ks, shuffle ? 1 : 0) > 0)
        RVSR_TRY(pack_weight_tc2(weight, (char *)dst + align_up(a, 256), Cout, Cin, ks, shuffle ? 1 : 0, s, &wv));
    return RVSR_OK;
}

int rvsr_c8_conv_pack_weights(const float *weight, int nviews, const int *spec, void *const *dst, void *stream) {
    RVSR_CHECK_ARG(weight && spec && dst && nviews >= 1, "c8 pack weights: bad arguments");
    int dims[RVSR_PACK_MAX_VIEWS][5];
    WeightView wv[RVSR_PACK_MAX_VIEWS];
    void *out[RVSR_PACK_MAX_VIEWS];
    int n = 0;
    for (int k = 0; k < nviews; ++k) {
        const int *v = spec + 8 * k;
        const int Cout = v[0], Cin = v[1], ks = v[2], shuffle = v[3] ? 1 : 0, mode = v[4], tot = v[5], c0 = v[6], layouts = v[7];
        RVSR_CHECK_ARG(Cout > 0 && Cin > 0 && (ks == 1 || ks == 3) && (mode == 0 || mode == 1) && dst[k] != nullptr, "c8 pack weights: view %d", k);
        RVSR_CHECK_ARG(((uintptr_t)dst[k] & 255) == 0, "c8 pack weights: destination must be 256-byte aligned");
        const size_t a = tc_conv_weight_bytes(Cout, Cin, ks, shuffle);
        if (a == 0) { set_error("c8 pack weights: shape %d x %d x %d not covered by the tcgen05 kernels", Cout, Cin, ks); return RVSR_E_UNSUPPORTED; }
        const long long KK = ks * ks;
        RVSR_CHECK_ARG(c0 >= 0 && c0 + (mode == 0 ? Cin : Cout) <= tot, "c8 pack weights: channel slice");
        const WeightView w = mode == 0 ? WeightView{c0 * KK, tot * KK, KK, 1, 1} : WeightView{c0 * KK + KK - 1, KK, tot * KK, -1, 1};
        for (int lay = 1; lay <= 2; lay <<= 1) {
            if (!(layouts & lay) || (lay == 2 && tc2_weight_bytes(Cout, Cin, ks, shuffle) == 0)) continue;
            RVSR_CHECK_ARG(n < RVSR_PACK_MAX_VIEWS, "c8 pack weights: more than %d layouts in one call", RVSR_PACK_MAX_VIEWS);
            dims[n][0] = Cout; dims[n][1] = Cin; dims[n][2] = ks; dims[n][3] = shuffle; dims[n][4] = lay == 2;
            wv[n] = w;
            out[n] = lay == 2 ? (char *)dst[k] + align_up(a, 256) : dst[k];
            ++n;
        }
    }
    if (n == 0) return RVSR_OK;
    return pack_weight_views(weight, n, dims, wv, out, (cudaStream_t)stream);
}

int rvsr_c8_conv_layouts(int nsrc, int C, int N, int H, int W, int Cout, int ks, int shuffle) {
    if (nsrc < 1 || nsrc > RVSR_MAX_SRC_TC || C <= 0 || N <= 0 || H <= 0 || W <= 0 || Cout <= 0) return 0;
    const int mode = shuffle ? 1 : 0;
    if (tc_conv_weight_bytes(Cout, nsrc * C, ks, mode) == 0) return 0;
    ConvOp op = {};
    static const char dummy = 0;   // the plan only asks whether the layouts exist, it never reads them
    for (int i = 0; i < nsrc; ++i) op.src[i] = Src{&dummy, (long long)cdiv(C, 8) * 8 * H * W, C, 1, -1};
    op.nsrc = nsrc;
    op.w_tc = &dummy;
    op.w_tc2 = tc2_weight_bytes(Cout, nsrc * C, ks, mode) > 0 ? &dummy : nullptr;
    op.N = N; op.H = H; op.W = W; op.Cout = Cout; op.ks = ks; op.stride = 1;
    op.out_mode = shuffle ? OUT_C8_SHUFFLE2 : OUT_C8; op.bf16 = 1;
    return tc_conv_layouts(op);
}

int rvsr_c8_conv_fwd(const void *const *x, const long long *x_image_stride, int nsrc, int C, const void *w_packed, const float *bias,
                     const void *residual, void *y, int N, int H, int W, int Cout, int ks, int stride, int act, int shuffle,
                     int residual_mode, float residual_slope, void *stream) {
    RVSR_CHECK_ARG(x && x_image_stride && nsrc >= 1 && nsrc <= RVSR_MAX_SRC_TC, "c8 conv: 1..%d sources", RVSR_MAX_SRC_TC);
    RVSR_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0 && Cout > 0 && (ks == 1 || ks == 3) && (stride == 1 || stride == 2), "c8 conv: bad sizes");
    RVSR_CHECK_ARG(!shuffle || (Cout % 4 == 0 && residual == nullptr && stride == 1), "c8 conv: pixel-shuffle needs Cout %% 4 == 0, no residual");
    RVSR_CHECK_ARG(residual_mode == 0 || (residual_mode == 2 && residual != nullptr), "c8 conv: residual_mode 0 (add) or 2 (mask)");
    if (N == 0) return RVSR_OK;
    RVSR_CHECK_ARG(w_packed && y, "c8 conv: null buffer");
    const int mode = shuffle ? 1 : 0;
    const size_t a = tc_conv_weight_bytes(Cout, nsrc * C, ks, mode);
    if (a == 0) { set_error("c8 conv: shape not covered by the tcgen05 kernels"); return RVSR_E_UNSUPPORTED; }
    ConvOp op = {};
    for (int i = 0; i < nsrc; ++i) {
        RVSR_CHECK_ARG(x[i] != nullptr, "c8 conv: null source %d", i);
        op.src[i] = Src{x[i], x_image_stride[i], C, 1, -1};
    }
    op.nsrc = nsrc;
    op.w_tc = w_packed;
    op.w_tc2 = tc2_weight_bytes(Cout, nsrc * C, ks, mode) > 0 ? (const char *)w_packed + align_up(a, 256) : nullptr;
    op.bias = bias;
    const int Ho = stride == 1 ? H : H / 2, Wo = stride == 1 ? W : W / 2;
    op.out = y;
    op.out_image_stride = shuffle ? (long long)cdiv(Cout / 4, 8) * 8 * 4 * Ho * Wo : (long long)cdiv(Cout, 8) * 8 * Ho * Wo;
    op.residual = residual; op.res_image_stride = op.out_image_stride;
    op.res_pre = residual_mode; op.res_slope = residual_slope;
    op.N = N; op.H = H; op.W = W; op.Cout = Cout; op.ks = ks; op.stride = stride; op.act = act;
    op.out_mode = shuffle ? OUT_C8_SHUFFLE2 : OUT_C8; op.sig_from = 1 << 30;
    op.bf16 = 1;
    if (!tc_conv_supported(op)) { set_error("c8 conv: configuration not covered by the tcgen05 kernels"); return RVSR_E_UNSUPPORTED; }
    return launch_conv_tc(op, (cudaStream_t)stream);
}

size_t rvsr_c8_conv_wgrad_workspace_bytes(int njobs, int N, int H, int W, int Cout) {
    return (njobs > 0 && N > 0 && H > 0 && W > 0 && Cout > 0) ? conv_wgrad_tc_workspace_bytes(njobs, N, H, W, Cout) : 512;
}
int rvsr_c8_conv_wgrad(int njobs, const void *const *x, const long long *x_image_stride, const void *const *g, float *const *gw,
                       float *const *db, const int *cin_total, const int *c0, int N, int H, int W, int Cin, int Cout, int ks,
                       void *workspace, size_t workspace_bytes, void *stream) {
    RVSR_CHECK_ARG(njobs >= 1 && x && x_image_stride && g && gw && db && cin_total && c0, "c8 conv wgrad: bad job arrays");
    RVSR_CHECK_ARG(N >= 0 && H > 0 && W > 0 && Cout > 0, "c8 conv wgrad: bad sizes");
    if (!conv_wgrad_tc_supported(Cin, ks, 1)) { set_error("c8 conv wgrad: built for 64 input channels, 3x3 / 1x1, stride 1"); return RVSR_E_UNSUPPORTED; }
    RVSR_CHECK_ARG(N == 0 || workspace, "c8 conv wgrad: null workspace");
    return launch_conv_wgrad_tc(njobs, x, x_image_stride, g, gw, db, cin_total, c0, N, H, W, Cout, ks, workspace, workspace_bytes,
                                (cudaStream_t)stream);
}

int rvsr_c8_act_bwd(const void *g, const void *y, void *out, long long n_elems, int act, void *stream) {
    RVSR_CHECK_ARG(n_elems >= 0 && (n_elems == 0 || (g && y && out)), "c8 act bwd: bad arguments");
    return launch_act_bwd_c8(g, y, out, n_elems, act, (cudaStream_t)stream);
}
int rvsr_c8_unshuffle2_act_bwd(const void *g, const void *y, void *out, int N, int C, int H, int W, int act, void *stream) {
    RVSR_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0 && g && out && (act == RVSR_ACT_NONE || y), "c8 unshuffle: bad arguments");
    return launch_unshuffle2_act_bwd_c8(g, y, out, N, C, H, W, act, (cudaStream_t)stream);
}
int rvsr_c8_upsample2x(const void *src, void *dst, long long planes, int H, int W, float scale, int backward, void *stream) {
    RVSR_CHECK_ARG(planes >= 0 && H > 0 && W > 0 && (planes == 0 || (src && dst)), "c8 upsample2x: bad arguments");
    return launch_upsample2x_c8(src, dst, planes, H, W, scale, backward, (cudaStream_t)stream);
}

size_t rvsr_c8_mdcn_workspace_bytes(int N, int H, int W, int backward) { return c8_mdcn_workspace_bytes(N, H, W, backward); }
int rvsr_c8_mdcn_fwd(const void *x, const void *om, const float *weight, const float *bias, void *y, int N, int H, int W, int act,
                     void *workspace, size_t workspace_bytes, void *stream) {
    RVSR_CHECK_ARG(N >= 0 && H > 0 && W > 0, "c8 mdcn fwd: bad sizes");
    RVSR_CHECK_ARG(N == 0 || (x && om && weight && y && workspace), "c8 mdcn fwd: null buffer");
    return c8_mdcn_fwd(x, om, weight, bias, y, N, H, W, act, workspace, workspace_bytes, (cudaStream_t)stream);
}
int rvsr_c8_mdcn_bwd(const void *x, const void *om, const float *weight, const void *g, const void *y, void *gx, void *gom, float *gw,
                     float *gb, int N, int H, int W, int act, void *workspace, size_t workspace_bytes, void *stream) {
    RVSR_CHECK_ARG(N >= 0 && H > 0 && W > 0, "c8 mdcn bwd: bad sizes");
    RVSR_CHECK_ARG(gw && (N == 0 || (x && om && weight && g && gx && gom && workspace)), "c8 mdcn bwd: null buffer");
    return c8_mdcn_bwd(x, om, weight, g, y, gx, gom, gw, gb, N, H, W, act, workspace, workspace_bytes, (cudaStream_t)stream);
}

int rvsr_c8_tsa_temporal(const void *aligned, const void *emb, const void *emb_ref, void *const *out, float *prob, int B, int N, int C,
                         int H, int W, void *stream) {
    RVSR_CHECK_ARG(B >= 0 && N > 0 && H > 0 && W > 0, "c8 tsa temporal: bad sizes");
    if (C != 64) { set_error("c8 tsa temporal: built for 64 channels"); return RVSR_E_UNSUPPORTED; }
    RVSR_CHECK_ARG(B == 0 || (aligned && emb && emb_ref && out && prob), "c8 tsa temporal: null buffer");
    return launch_tsa_temporal_c8(aligned, emb, emb_ref, out, prob, B, N, H, W, (cudaStream_t)stream);
}
int rvsr_c8_tsa_temporal_bwd(const void *const *gout, const void *aligned, const void *emb, const void *emb_ref, const float *prob,
                             void *g_aligned, void *g_emb, void *g_emb_ref, int B, int N, int C, int H, int W, void *stream) {
    RVSR_CHECK_ARG(B >= 0 && N > 0 && H > 0 && W > 0, "c8 tsa temporal bwd: bad sizes");
    if (C != 64) { set_error("c8 tsa temporal: built for 64 channels"); return RVSR_E_UNSUPPORTED; }
    RVSR_CHECK_ARG(B == 0 || (gout && aligned && emb && emb_ref && prob && g_aligned && g_emb && g_emb_ref), "c8 tsa temporal bwd: null buffer");
    return launch_tsa_temporal_bwd_c8(gout, aligned, emb, emb_ref, prob, g_aligned, g_emb, g_emb_ref, B, N, H, W, (cudaStream_t)stream);
}

int rvsr_c8_pool_maxavg(const void *src, void *dst_max, void *dst_avg, long long planes, int H, int W, void *stream) {
    RVSR_CHECK_ARG(planes >= 0 && H > 0 && W > 0 && (planes == 0 || (src && dst_max && dst_avg)), "c8 pool: bad arguments");
    return launch_pool_maxavg_c8(src, dst_max, dst_avg, planes, H, W, (cudaStream_t)stream);
}
int rvsr_c8_pool_maxavg_bwd(const void *src, const void *g_max, const void *g_avg, void *g_src, long long planes, int H, int W, void *stream) {
    RVSR_CHECK_ARG(planes >= 0 && H > 0 && W > 0 && (planes == 0 || (src && g_src)), "c8 pool bwd: bad arguments");
    return launch_pool_maxavg_bwd_c8(src, g_max, g_avg, g_src, planes, H, W, (cudaStream_t)stream);
}
int rvsr_c8_tsa_final(const void *fea, const void *att, const void *att_add, void *out, long long n_elems, void *stream) {
    RVSR_CHECK_ARG(n_elems >= 0 && n_elems % 8 == 0 && (n_elems == 0 || (fea && att && att_add && out)), "c8 tsa final: bad arguments");
    return launch_tsa_final_c8(fea, att, att_add, out, n_elems, (cudaStream_t)stream);
}
int rvsr_c8_tsa_final_bwd(const void *g, const void *fea, const void *att, void *g_fea, void *g_att, long long n_elems, void *stream) {
    RVSR_CHECK_ARG(n_elems >= 0 && n_elems % 8 == 0 && (n_elems == 0 || (g && fea && att && g_fea && g_att)), "c8 tsa final bwd: bad arguments");
    return launch_tsa_final_bwd_c8(g, fea, att, g_fea, g_att, n_elems, (cudaStream_t)stream);
}

int rvsr_upsample2x_nchw(const void *src, void *dst, long long planes, int H, int W, float scale, int backward, int dtype, void *stream) {
    RVSR_CHECK_ARG(planes >= 0 && H > 0 && W > 0 && (planes == 0 || (src && dst)), "upsample2x (NCHW): bad arguments");
    return launch_upsample2x_nchw(src, dst, planes, H, W, scale, backward, dtype, (cudaStream_t)stream);
}

}  // extern "C"
