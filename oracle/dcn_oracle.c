/*
 * oracle/dcn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's modulated deformable convolution
 * (DCNv2), forward and backward.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this; the product path never does.
 *
 * What it follows in the reference (/root/reference, read-only):
 *   forward  sample   : codes/models/archs/dcn/src/deform_conv_cuda_kernel.cu:467-497 (dmcn_im2col_bilinear)
 *   forward  gather   : ...deform_conv_cuda_kernel.cu:571-633 (modulated_deformable_im2col_gpu_kernel)
 *   forward  contract : codes/models/archs/dcn/src/deform_conv_cuda.cpp:539-568 (addmm_ + bias)
 *   backward grad_col : ...deform_conv_cuda.cpp:617-626
 *   backward coord    : ...deform_conv_cuda_kernel.cu:526-568, :695-767 (grad_offset, grad_mask)
 *   backward input    : ...deform_conv_cuda_kernel.cu:499-524, :635-693 (scatter to <=4 corners)
 *   backward params   : ...deform_conv_cuda.cpp:647-671 (grad_weight, grad_bias; accumulate over batch)
 *
 * It is written from the operator definition, not
 * translated line by line: one loop nest per output pixel that builds the
 * modulated sample vector and contracts it, instead of im2col + GEMM.
 *
 * Pinning: tests/test_oracle.py checks it against torchvision.ops.deform_conv2d
 * (same MSRA lineage, CPU, fp64) and against the golden fixtures generated
 * from the reference's own EDVR_arch.py (tests/golden/make_golden.py); on the
 * GPU box tests/test_dcn_gpu.py also checks it against the reference's own
 * compiled CUDA extension (oracle/_ref) when that was built.
 *
 * Layouts (all contiguous, NCHW like the reference):
 *   input  [B, C, H, W]
 *   offset [B, dg*2*kh*kw, Ho, Wo]   channel = g*2*kh*kw + 2*(i*kw+j) + {0: dy, 1: dx}
 *   mask   [B, dg*kh*kw,   Ho, Wo]   channel = g*kh*kw + i*kw + j   (already sigmoid-ed)
 *   weight [Cout, C/groups, kh, kw],  bias [Cout] or NULL
 *   output [B, Cout, Ho, Wo]
 * All arithmetic is done in double; the _f32 entry points convert at the edges.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int B, C, H, W, Cout, kh, kw, stride, pad, dil, groups, dg, Ho, Wo;
} dcn_shape;

static int dcn_shape_init(dcn_shape *s, int B, int C, int H, int W, int Cout, int kh, int kw,
                          int stride, int pad, int dil, int groups, int dg) {
    if (B < 0 || C <= 0 || H <= 0 || W <= 0 || Cout <= 0 || kh <= 0 || kw <= 0 || stride <= 0 ||
        pad < 0 || dil <= 0 || groups <= 0 || dg <= 0)
        return -1;
    if (C % groups || Cout % groups || C % dg) return -2;
    s->B = B; s->C = C; s->H = H; s->W = W; s->Cout = Cout; s->kh = kh; s->kw = kw;
    s->stride = stride; s->pad = pad; s->dil = dil; s->groups = groups; s->dg = dg;
    s->Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
    s->Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
    if (s->Ho <= 0 || s->Wo <= 0) return -3;
    return 0;
}

/* The four bilinear corners of a sampling point and their weights.  A corner
 * that falls outside the image keeps weight but has valid=0 (contributes 0),
 * matching the per-corner bounds checks of the reference sampler. */
typedef struct {
    int inside;          /* py > -1 && px > -1 && py < H && px < W */
    int y[2], x[2];      /* low / high */
    double wy[2], wx[2]; /* (1-ly, ly), (1-lx, lx) */
    int vy[2], vx[2];    /* corner row/col inside the image? */
} corners;

static void corners_at(corners *c, double py, double px, int H, int W) {
    c->inside = (py > -1.0 && px > -1.0 && py < (double)H && px < (double)W);
    double fy = floor(py), fx = floor(px);
    c->y[0] = (int)fy; c->y[1] = c->y[0] + 1;
    c->x[0] = (int)fx; c->x[1] = c->x[0] + 1;
    double ly = py - fy, lx = px - fx;
    c->wy[0] = 1.0 - ly; c->wy[1] = ly;
    c->wx[0] = 1.0 - lx; c->wx[1] = lx;
    c->vy[0] = c->y[0] >= 0;      c->vy[1] = c->y[1] <= H - 1;
    c->vx[0] = c->x[0] >= 0;      c->vx[1] = c->x[1] <= W - 1;
}

static double sample_plane(const double *plane, const corners *c, int W) {
    if (!c->inside) return 0.0;
    double v = 0.0;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            if (c->vy[a] && c->vx[b]) v += c->wy[a] * c->wx[b] * plane[c->y[a] * W + c->x[b]];
    return v;
}

int dcn_oracle_fwd_f64(const double *input, const double *offset, const double *mask,
                       const double *weight, const double *bias, double *output, int B, int C,
                       int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil,
                       int groups, int dg) {
    dcn_shape s;
    int rc = dcn_shape_init(&s, B, C, H, W, Cout, kh, kw, stride, pad, dil, groups, dg);
    if (rc) return rc;
    const int K = kh * kw, cpg = C / dg, cin_g = C / groups, cout_g = Cout / groups;
    const long plane_o = (long)s.Ho * s.Wo;
    /* rows are independent: one scratch vector per thread (OpenMP is only used
     * so the bench's CPU-baseline leg can use all host cores) */
#pragma omp parallel
    {
    double *col = (double *)malloc(sizeof(double) * (size_t)C * K);
#pragma omp for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < s.Ho; ++y)
            for (int x = 0; col && x < s.Wo; ++x) {
                const long pix = (long)y * s.Wo + x;
                /* modulated sample vector col[c*K + k] for this output pixel */
                for (int g = 0; g < dg; ++g)
                    for (int k = 0; k < K; ++k) {
                        const int i = k / kw, j = k % kw;
                        const double dy = offset[((long)b * dg * 2 * K + g * 2 * K + 2 * k) * plane_o + pix];
                        const double dx = offset[((long)b * dg * 2 * K + g * 2 * K + 2 * k + 1) * plane_o + pix];
                        const double m = mask[((long)b * dg * K + g * K + k) * plane_o + pix];
                        corners cs;
                        corners_at(&cs, y * stride - pad + i * dil + dy, x * stride - pad + j * dil + dx, H, W);
                        for (int cc = 0; cc < cpg; ++cc) {
                            const int c = g * cpg + cc;
                            col[c * K + k] = m * sample_plane(input + ((long)b * C + c) * H * W, &cs, W);
                        }
                    }
                /* contraction with the (grouped) weight, plus bias */
                for (int o = 0; o < Cout; ++o) {
                    const int cg = o / cout_g;
                    double acc = 0.0;
                    const double *wrow = weight + (long)o * cin_g * K;
                    const double *crow = col + (long)cg * cin_g * K;
                    for (int t = 0; t < cin_g * K; ++t) acc += wrow[t] * crow[t];
                    output[((long)b * Cout + o) * plane_o + pix] = acc + (bias ? bias[o] : 0.0);
                }
            }
    free(col);
    }
    return 0;
}

/* Gradients.  grad_input/grad_offset/grad_mask are overwritten; grad_weight and
 * grad_bias are ACCUMULATED into (the reference accumulates them across the
 * batch into caller-zeroed tensors, deform_conv_cuda.cpp:659-671). */
int dcn_oracle_bwd_f64(const double *input, const double *offset, const double *mask,
                       const double *weight, const double *grad_output, double *grad_input,
                       double *grad_offset, double *grad_mask, double *grad_weight,
                       double *grad_bias, int B, int C, int H, int W, int Cout, int kh, int kw,
                       int stride, int pad, int dil, int groups, int dg) {
    dcn_shape s;
    int rc = dcn_shape_init(&s, B, C, H, W, Cout, kh, kw, stride, pad, dil, groups, dg);
    if (rc) return rc;
    const int K = kh * kw, cpg = C / dg, cin_g = C / groups, cout_g = Cout / groups;
    const long plane_o = (long)s.Ho * s.Wo, plane_i = (long)H * W;
    memset(grad_input, 0, sizeof(double) * (size_t)B * C * plane_i);
    memset(grad_offset, 0, sizeof(double) * (size_t)B * dg * 2 * K * plane_o);
    memset(grad_mask, 0, sizeof(double) * (size_t)B * dg * K * plane_o);
    double *gcol = (double *)malloc(sizeof(double) * (size_t)C * K);
    if (!gcol) return -4;
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < s.Ho; ++y)
            for (int x = 0; x < s.Wo; ++x) {
                const long pix = (long)y * s.Wo + x;
                /* grad wrt the modulated sample vector: W^T . grad_out */
                for (int t = 0; t < C * K; ++t) gcol[t] = 0.0;
                for (int o = 0; o < Cout; ++o) {
                    const int cg = o / cout_g;
                    const double go = grad_output[((long)b * Cout + o) * plane_o + pix];
                    const double *wrow = weight + (long)o * cin_g * K;
                    double *grow = gcol + (long)cg * cin_g * K;
                    for (int t = 0; t < cin_g * K; ++t) grow[t] += wrow[t] * go;
                    if (grad_bias) grad_bias[o] += go;
                }
                for (int g = 0; g < dg; ++g)
                    for (int k = 0; k < K; ++k) {
                        const int i = k / kw, j = k % kw;
                        const long oc = (long)b * dg * 2 * K + g * 2 * K + 2 * k;
                        const long mc = (long)b * dg * K + g * K + k;
                        const double dy = offset[oc * plane_o + pix];
                        const double dx = offset[(oc + 1) * plane_o + pix];
                        const double m = mask[mc * plane_o + pix];
                        corners cs;
                        corners_at(&cs, y * stride - pad + i * dil + dy, x * stride - pad + j * dil + dx, H, W);
                        double g_dy = 0.0, g_dx = 0.0, g_m = 0.0;
                        for (int cc = 0; cc < cpg; ++cc) {
                            const int c = g * cpg + cc;
                            const double *plane = input + ((long)b * C + c) * plane_i;
                            double *gplane = grad_input + ((long)b * C + c) * plane_i;
                            const double gc = gcol[c * K + k];
                            const double val = sample_plane(plane, &cs, W);
                            /* weight grad sees the modulated sample */
                            for (int o = 0; o < Cout; ++o) {
                                if (o / cout_g != c / cin_g) continue;
                                const double go = grad_output[((long)b * Cout + o) * plane_o + pix];
                                grad_weight[((long)o * cin_g + (c % cin_g)) * K + k] += go * m * val;
                            }
                            if (!cs.inside) continue; /* out-of-range tap: zero gradient everywhere */
                            g_m += gc * val;
                            for (int a = 0; a < 2; ++a)
                                for (int bb = 0; bb < 2; ++bb) {
                                    if (!(cs.vy[a] && cs.vx[bb])) continue;
                                    const double v = plane[cs.y[a] * W + cs.x[bb]];
                                    /* d(sample)/d(py): -wx on the low row, +wx on the high row */
                                    g_dy += gc * m * (a ? 1.0 : -1.0) * cs.wx[bb] * v;
                                    g_dx += gc * m * (bb ? 1.0 : -1.0) * cs.wy[a] * v;
                                    gplane[cs.y[a] * W + cs.x[bb]] += gc * m * cs.wy[a] * cs.wx[bb];
                                }
                        }
                        grad_offset[oc * plane_o + pix] = g_dy;
                        grad_offset[(oc + 1) * plane_o + pix] = g_dx;
                        grad_mask[mc * plane_o + pix] = g_m;
                    }
            }
    free(gcol);
    return 0;
}

/* ---- float32 edges (convert, run in double, convert back) ---- */
static double *to_f64(const float *p, size_t n) {
    if (!p) return NULL;
    double *d = (double *)malloc(sizeof(double) * (n ? n : 1));
    for (size_t i = 0; d && i < n; ++i) d[i] = (double)p[i];
    return d;
}
static void from_f64(float *dst, const double *src, size_t n) {
    for (size_t i = 0; i < n; ++i) dst[i] = (float)src[i];
}

int dcn_oracle_fwd_f32(const float *input, const float *offset, const float *mask,
                       const float *weight, const float *bias, float *output, int B, int C, int H,
                       int W, int Cout, int kh, int kw, int stride, int pad, int dil, int groups,
                       int dg) {
    dcn_shape s;
    int rc = dcn_shape_init(&s, B, C, H, W, Cout, kh, kw, stride, pad, dil, groups, dg);
    if (rc) return rc;
    const size_t K = (size_t)kh * kw, po = (size_t)s.Ho * s.Wo;
    double *i64 = to_f64(input, (size_t)B * C * H * W);
    double *o64 = to_f64(offset, (size_t)B * dg * 2 * K * po);
    double *m64 = to_f64(mask, (size_t)B * dg * K * po);
    double *w64 = to_f64(weight, (size_t)Cout * (C / groups) * K);
    double *b64 = to_f64(bias, (size_t)Cout);
    double *y64 = (double *)malloc(sizeof(double) * (size_t)B * Cout * po + 8);
    rc = dcn_oracle_fwd_f64(i64, o64, m64, w64, b64, y64, B, C, H, W, Cout, kh, kw, stride, pad,
                            dil, groups, dg);
    if (!rc) from_f64(output, y64, (size_t)B * Cout * po);
    free(i64); free(o64); free(m64); free(w64); free(b64); free(y64);
    return rc;
}

int dcn_oracle_bwd_f32(const float *input, const float *offset, const float *mask,
                       const float *weight, const float *grad_output, float *grad_input,
                       float *grad_offset, float *grad_mask, float *grad_weight, float *grad_bias,
                       int B, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad,
                       int dil, int groups, int dg) {
    dcn_shape s;
    int rc = dcn_shape_init(&s, B, C, H, W, Cout, kh, kw, stride, pad, dil, groups, dg);
    if (rc) return rc;
    const size_t K = (size_t)kh * kw, po = (size_t)s.Ho * s.Wo, pi = (size_t)H * W;
    const size_t n_in = (size_t)B * C * pi, n_off = (size_t)B * dg * 2 * K * po,
                 n_m = (size_t)B * dg * K * po, n_w = (size_t)Cout * (C / groups) * K;
    double *i64 = to_f64(input, n_in), *o64 = to_f64(offset, n_off), *m64 = to_f64(mask, n_m);
    double *w64 = to_f64(weight, n_w), *g64 = to_f64(grad_output, (size_t)B * Cout * po);
    double *gi = (double *)malloc(sizeof(double) * n_in + 8);
    double *go = (double *)malloc(sizeof(double) * n_off + 8);
    double *gm = (double *)malloc(sizeof(double) * n_m + 8);
    double *gw = to_f64(grad_weight, n_w);
    double *gb = to_f64(grad_bias, (size_t)Cout);
    rc = dcn_oracle_bwd_f64(i64, o64, m64, w64, g64, gi, go, gm, gw, gb, B, C, H, W, Cout, kh, kw,
                            stride, pad, dil, groups, dg);
    if (!rc) {
        from_f64(grad_input, gi, n_in); from_f64(grad_offset, go, n_off);
        from_f64(grad_mask, gm, n_m);   from_f64(grad_weight, gw, n_w);
        if (grad_bias) from_f64(grad_bias, gb, (size_t)Cout);
    }
    free(i64); free(o64); free(m64); free(w64); free(g64);
    free(gi); free(go); free(gm); free(gw); free(gb);
    return rc;
}
