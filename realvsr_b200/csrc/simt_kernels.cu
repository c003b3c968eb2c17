// simt_kernels.cu -- CUDA-core kernels of the rvsr_b200 hot path.
//
// Two families live here:
//  (1) the fp32-accumulate "tile GEMM" convolution / deformable-convolution kernels that
//      serve the strict-parity fp32 mode, odd shapes (Cin = 3, stride 2, tiny nf) and act as
//      the cross-check for the tcgen05 kernels in tc_kernels.cu;
//  (2) the bandwidth-bound glue ops (layout pack/unpack, x2 bilinear upsample, 3/2/1 pools,
//      TSA temporal attention, final fusion) used by BOTH precisions: 128-bit vector
//      accesses over the channel-blocked layout, one pass, nothing materialised twice.
//
// Reference semantics being reproduced (IanYeung/RealVSR, codes/models/archs/):
//   DCN sample/gather  dcn/src/deform_conv_cuda_kernel.cu:467-497, :571-633
//   DCN contraction    dcn/src/deform_conv_cuda.cpp:539-568
//   upsample / pools / attention   EDVR_arch.py:111-124, :154-155, :175-207
#include <cuda_bf16.h>

#include "common.cuh"

namespace rvsr {

static constexpr int TILE_W = 8, TILE_H = 8, TILE_P = 64;  // output pixels per CTA
static constexpr int TILE_CO = 64;                         // output channels per CTA
static constexpr int NTHREADS = 256;

__device__ __forceinline__ long long src_image(const Src &s, int n) {
    const int m = s.map != nullptr ? __ldg(s.map + n) : (s.fixed_frame >= 0 ? (n / s.frames) * s.frames + s.fixed_frame : n);
    return (long long)m * s.image_stride;
}

// ---------------------------------------------------------------- shared consumer
// acc[j][k]: output channel co0 + j, pixel 4*pg + k of the tile.
template <typename T>
__device__ __forceinline__ void tile_epilogue(float (&acc)[4][4], const float *bias, void *out_,
                                              long long out_image_stride, const void *res_,
                                              long long res_image_stride, int n, int Cout, int Ho,
                                              int Wo, int oy0, int ox0, int pg, int co0, int act,
                                              int out_mode, int sig_from) {
    const int Co8 = (Cout + 7) / 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int p = pg * 4 + k;
        const int oy = oy0 + p / TILE_W, ox = ox0 + p % TILE_W;
        if (oy >= Ho || ox >= Wo) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + j;
            float t = acc[j][k] + ((bias != nullptr && co < Cout) ? bias[co] : 0.f);
            v[j] = apply_act(t, act);
        }
        if (out_mode == OUT_C8) {
            if (co0 >= Co8 * 8) continue;
            const long long e = ((((long long)(co0 / 8)) * Ho + oy) * Wo + ox) * 8 + (co0 % 8);
            T *o = reinterpret_cast<T *>(out_) + (long long)n * out_image_stride + e;
            if (res_ != nullptr) {
                const T *r = reinterpret_cast<const T *>(res_) + (long long)n * res_image_stride + e;
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] += to_f<T>(r[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = from_f<T>(co0 + j < Cout ? v[j] : 0.f);
        } else if (out_mode == OUT_C8_SHUFFLE2) {
            // out[n, c, 2y+i, 2x+j] = in[n, 4c+2i+j, y, x]
            if (co0 >= Cout) continue;
            const int c = co0 / 4, C2 = Cout / 4, C28 = (C2 + 7) / 8;
            (void)C28;
            T *o = reinterpret_cast<T *>(out_) + (long long)n * out_image_stride;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int yy = 2 * oy + (j >> 1), xx = 2 * ox + (j & 1);
                o[((((long long)(c / 8)) * (2 * Ho) + yy) * (2 * Wo) + xx) * 8 + (c % 8)] = from_f<T>(v[j]);
            }
        } else if (out_mode == OUT_PLANAR_F32) {
            float *o = reinterpret_cast<float *>(out_) + (long long)n * out_image_stride;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int co = co0 + j;
                if (co < Cout) o[((long long)co * Ho + oy) * Wo + ox] = co >= sig_from ? sigmoidf_(v[j]) : v[j];
            }
        } else {  // OUT_NCHW_T
            T *o = reinterpret_cast<T *>(out_) + (long long)n * out_image_stride;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int co = co0 + j;
                if (co < Cout) o[((long long)co * Ho + oy) * Wo + ox] = from_f<T>(v[j]);
            }
        }
    }
}

template <int ROWS>
__device__ __forceinline__ void tile_fma(const float (*col)[TILE_P], const float (*wt)[TILE_CO],
                                         float (&acc)[4][4], int pg, int cg) {
#pragma unroll 8
    for (int r = 0; r < ROWS; ++r) {
        const float4 a = *reinterpret_cast<const float4 *>(&col[r][pg * 4]);
        const float4 b = *reinterpret_cast<const float4 *>(&wt[r][cg * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[j][k] = fmaf(bv[j], av[k], acc[j][k]);
    }
}

template <int ROWS>
__device__ __forceinline__ void load_weight_tile(float (*wt)[TILE_CO], const float *w, long long row0,
                                                 int cout_pad, int cb) {
    for (int i = threadIdx.x; i < ROWS * (TILE_CO / 4); i += NTHREADS) {
        const int r = i / (TILE_CO / 4), c4 = i % (TILE_CO / 4);
        *reinterpret_cast<float4 *>(&wt[r][c4 * 4]) =
            __ldg(reinterpret_cast<const float4 *>(w + (row0 + r) * cout_pad + cb * TILE_CO) + c4);
    }
}

// ---------------------------------------------------------------- convolution (ks x ks, pad ks/2)
template <typename T, int KS>
__global__ void __launch_bounds__(NTHREADS) conv_simt_kernel(const ConvOp op) {
    constexpr int KK = KS * KS, ROWS = KK * 8;
    __shared__ __align__(16) float col[ROWS][TILE_P];
    __shared__ __align__(16) float wt[ROWS][TILE_CO];
    const int pad = KS / 2;
    const int Ho = (op.H + 2 * pad - KS) / op.stride + 1, Wo = (op.W + 2 * pad - KS) / op.stride + 1;
    const int ntx = (Wo + TILE_W - 1) / TILE_W;
    const int ox0 = (blockIdx.x % ntx) * TILE_W, oy0 = (blockIdx.x / ntx) * TILE_H;
    const int cb = blockIdx.y, n = blockIdx.z;
    const int cout_pad = ((op.Cout + TILE_CO - 1) / TILE_CO) * TILE_CO;
    const int pg = threadIdx.x % 16, cg = threadIdx.x / 16;
    float acc[4][4] = {};
    int chunk_base = 0;
    for (int s = 0; s < op.nsrc; ++s) {
        const Src &src = op.src[s];
        const int C8 = (src.C + 7) / 8;
        const T *base = reinterpret_cast<const T *>(src.ptr) + src_image(src, n);
        for (int q = 0; q < C8; ++q) {
            __syncthreads();
            load_weight_tile<ROWS>(wt, op.w_simt, (long long)(chunk_base + q) * ROWS, cout_pad, cb);
            for (int i = threadIdx.x; i < TILE_P * KK; i += NTHREADS) {
                const int p = i % TILE_P, t = i / TILE_P;
                const int oy = oy0 + p / TILE_W, ox = ox0 + p % TILE_W;
                const int iy = oy * op.stride - pad + t / KS, ix = ox * op.stride - pad + t % KS;
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (oy < Ho && ox < Wo && iy >= 0 && iy < op.H && ix >= 0 && ix < op.W)
                    load8<T>(base + ((((long long)q) * op.H + iy) * op.W + ix) * 8, v);
#pragma unroll
                for (int c = 0; c < 8; ++c) col[t * 8 + c][p] = v[c];
            }
            __syncthreads();
            tile_fma<ROWS>(col, wt, acc, pg, cg);
        }
        chunk_base += C8;
    }
    tile_epilogue<T>(acc, op.bias, op.out, op.out_image_stride, op.residual, op.res_image_stride, n,
                     op.Cout, Ho, Wo, oy0, ox0, pg, cb * TILE_CO + cg * 4, op.act, op.out_mode,
                     op.sig_from);
}

template <typename T> int launch_conv_simt(const ConvOp &op, cudaStream_t s) {
    RVSR_CHECK_ARG(op.ks == 1 || op.ks == 3, "conv: kernel size %d not supported", op.ks);
    RVSR_CHECK_ARG(op.stride == 1 || op.stride == 2, "conv: stride %d not supported", op.stride);
    RVSR_CHECK_ARG(op.nsrc >= 1 && op.nsrc <= RVSR_MAX_SRC, "conv: %d sources", op.nsrc);
    RVSR_CHECK_ARG(op.out_mode != OUT_C8_SHUFFLE2 || op.Cout % 4 == 0, "conv: shuffle needs Cout%%4==0");
    const int pad = op.ks / 2;
    const int Ho = (op.H + 2 * pad - op.ks) / op.stride + 1, Wo = (op.W + 2 * pad - op.ks) / op.stride + 1;
    dim3 grid(cdiv(Wo, TILE_W) * cdiv(Ho, TILE_H), cdiv(op.Cout, TILE_CO), op.N);
    if (grid.x == 0 || grid.z == 0) return RVSR_OK;
    RVSR_CHECK_ARG(grid.z <= 65535 && grid.y <= 65535, "conv: too many images");
    if (op.ks == 3)
        conv_simt_kernel<T, 3><<<grid, NTHREADS, 0, s>>>(op);
    else
        conv_simt_kernel<T, 1><<<grid, NTHREADS, 0, s>>>(op);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
template int launch_conv_simt<float>(const ConvOp &, cudaStream_t);
template int launch_conv_simt<__half>(const ConvOp &, cudaStream_t);

// ---------------------------------------------------------------- modulated deformable conv
template <typename T, int K>
__global__ void __launch_bounds__(NTHREADS) dcn_simt_kernel(const DcnOp op) {
    constexpr int ROWS = K * 8;
    __shared__ __align__(16) float col[ROWS][TILE_P];
    __shared__ __align__(16) float wt[ROWS][TILE_CO];
    const int Ho = (op.H + 2 * op.pad - (op.dil * (op.kh - 1) + 1)) / op.stride + 1;
    const int Wo = (op.W + 2 * op.pad - (op.dil * (op.kw - 1) + 1)) / op.stride + 1;
    const int ntx = (Wo + TILE_W - 1) / TILE_W;
    const int ox0 = (blockIdx.x % ntx) * TILE_W, oy0 = (blockIdx.x / ntx) * TILE_H;
    const int cb = blockIdx.y, n = blockIdx.z;
    const int cout_pad = ((op.Cout + TILE_CO - 1) / TILE_CO) * TILE_CO;
    const int pg = threadIdx.x % 16, cg = threadIdx.x / 16;
    const int C = op.x.C, C8 = (C + 7) / 8, cpg = C / op.dg;
    const long long plane = (long long)Ho * Wo;
    const T *xb = reinterpret_cast<const T *>(op.x.ptr) + src_image(op.x, n);
    const float *off = op.offset + (long long)n * op.offset_image_stride;
    const float *msk = op.mask + (long long)n * op.mask_image_stride;
    float acc[4][4] = {};
    for (int q = 0; q < C8; ++q) {
        __syncthreads();
        load_weight_tile<ROWS>(wt, op.w_simt, (long long)q * ROWS, cout_pad, cb);
        const T *xq = xb + (long long)q * op.H * op.W * 8;
        for (int i = threadIdx.x; i < TILE_P * K; i += NTHREADS) {
            const int p = i % TILE_P, t = i / TILE_P;
            const int oy = oy0 + p / TILE_W, ox = ox0 + p % TILE_W;
            float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (oy < Ho && ox < Wo) {
                const long long pix = (long long)oy * Wo + ox;
                const float by = (float)(oy * op.stride - op.pad + (t / op.kw) * op.dil);
                const float bx = (float)(ox * op.stride - op.pad + (t % op.kw) * op.dil);
                int gprev = -1;
                float m = 0.f, v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int ch = q * 8 + c;
                    if (ch >= C) break;
                    const int g = ch / cpg;
                    if (g != gprev) {  // one coordinate pair + mask per (deformable group, tap)
                        gprev = g;
                        const float dy = __ldg(off + ((long long)g * 2 * K + 2 * t) * plane + pix);
                        const float dx = __ldg(off + ((long long)g * 2 * K + 2 * t + 1) * plane + pix);
                        m = __ldg(msk + ((long long)g * K + t) * plane + pix);
                        sample8<T>(xq, op.H, op.W, by + dy, bx + dx, v);
                    }
                    r[c] = m * v[c];
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) col[t * 8 + c][p] = r[c];
        }
        __syncthreads();
        tile_fma<ROWS>(col, wt, acc, pg, cg);
    }
    tile_epilogue<T>(acc, op.bias, op.out, op.out_image_stride, nullptr, 0, n, op.Cout, Ho, Wo, oy0, ox0,
                     pg, cb * TILE_CO + cg * 4, op.act, op.out_mode, 0);
}

template <typename T> int launch_dcn_simt(const DcnOp &op, cudaStream_t s) {
    const int K = op.kh * op.kw;
    RVSR_CHECK_ARG(K == 9 || K == 1, "dcn: only 3x3 and 1x1 kernels are built (got %dx%d)", op.kh, op.kw);
    RVSR_CHECK_ARG(op.dg > 0 && op.x.C % op.dg == 0, "dcn: channels %d not divisible by deformable groups %d",
                   op.x.C, op.dg);
    const int Ho = (op.H + 2 * op.pad - (op.dil * (op.kh - 1) + 1)) / op.stride + 1;
    const int Wo = (op.W + 2 * op.pad - (op.dil * (op.kw - 1) + 1)) / op.stride + 1;
    RVSR_CHECK_ARG(Ho > 0 && Wo > 0, "dcn: empty output");
    dim3 grid(cdiv(Wo, TILE_W) * cdiv(Ho, TILE_H), cdiv(op.Cout, TILE_CO), op.N);
    if (grid.z == 0) return RVSR_OK;
    RVSR_CHECK_ARG(grid.z <= 65535, "dcn: too many images");
    if (K == 9)
        dcn_simt_kernel<T, 9><<<grid, NTHREADS, 0, s>>>(op);
    else
        dcn_simt_kernel<T, 1><<<grid, NTHREADS, 0, s>>>(op);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
template int launch_dcn_simt<float>(const DcnOp &, cudaStream_t);
template int launch_dcn_simt<__half>(const DcnOp &, cudaStream_t);

// ---------------------------------------------------------------- weight packing
// dst[((chunk*KK + t)*8 + ci) * cout_pad + co] = w[co][cin(chunk, ci)][t]   (zero padded)
__global__ void pack_weight_simt_kernel(const float *__restrict__ w, float *__restrict__ dst, int Cout,
                                        int Cin_total, int KK, int nsrc, int c0, int c1, int c2, int c3,
                                        int c4, int c5, int c6, int cout_pad, long long total) {
    const int cs[RVSR_MAX_SRC] = {c0, c1, c2, c3, c4, c5, c6};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % cout_pad);
        long long r = i / cout_pad;
        const int ci = (int)(r % 8);
        r /= 8;
        const int t = (int)(r % KK);
        int chunk = (int)(r / KK);
        int cin_off = 0, cin = -1;
        for (int s = 0; s < nsrc; ++s) {
            const int c8 = (cs[s] + 7) / 8;
            if (chunk < c8) {
                if (chunk * 8 + ci < cs[s]) cin = cin_off + chunk * 8 + ci;
                break;
            }
            chunk -= c8;
            cin_off += cs[s];
        }
        dst[i] = (cin >= 0 && co < Cout) ? w[((long long)co * Cin_total + cin) * KK + t] : 0.f;
    }
}

int pack_weight_simt(const float *w_oihw, float *dst, int Cout, int Cin_total, int ks, const int *src_channels,
                     int nsrc, int cout_pad, cudaStream_t s) {
    RVSR_CHECK_ARG(nsrc >= 1 && nsrc <= RVSR_MAX_SRC, "pack: %d sources", nsrc);
    int c[RVSR_MAX_SRC] = {0, 0, 0, 0, 0, 0, 0};
    int chunks = 0, sum = 0;
    for (int i = 0; i < nsrc; ++i) {
        c[i] = src_channels[i];
        chunks += (c[i] + 7) / 8;
        sum += c[i];
    }
    RVSR_CHECK_ARG(sum == Cin_total, "pack: source channels %d != weight input channels %d", sum, Cin_total);
    const long long total = (long long)chunks * ks * ks * 8 * cout_pad;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    pack_weight_simt_kernel<<<blocks, 256, 0, s>>>(w_oihw, dst, Cout, Cin_total, ks * ks, nsrc, c[0], c[1],
                                                   c[2], c[3], c[4], c[5], c[6], cout_pad, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

// ---------------------------------------------------------------- layout pack / unpack
template <typename T, typename Tin>
__global__ void pack_nchw_kernel(const Tin *__restrict__ src, T *__restrict__ dst, int C, int H, int W) {
    pdl_trigger();
    pdl_wait();
    // grid = (pixels / 256, channel blocks, images): 32-bit index math only (three 64-bit divisions per element made
    // this 46 MB copy take 34 us)
    const int pix = blockIdx.x * blockDim.x + threadIdx.x, HW = H * W;
    if (pix >= HW) return;
    const int q = blockIdx.y, C8 = gridDim.y;
    const long long n = blockIdx.z;
    const Tin *sp = src + (n * C + q * 8) * (long long)HW + pix;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = q * 8 + c < C ? to_f<Tin>(sp[(long long)c * HW]) : 0.f;
    store8<T>(dst + ((n * C8 + q) * (long long)HW + pix) * 8, v);
}
template <typename T, typename Tin>
int launch_pack_nchw(const Tin *src, T *dst, int N, int C, int H, int W, cudaStream_t s, int Cdst) {
    if (Cdst < C) Cdst = C;
    const int C8 = (Cdst + 7) / 8;
    if ((long long)N * C8 * H * W == 0) return RVSR_OK;
    RVSR_CHECK_ARG(N <= 65535 && C8 <= 65535, "pack_nchw: too many images / channel blocks (%d, %d)", N, C8);
    launch_k(pack_nchw_kernel<T, Tin>, dim3((H * W + 255) / 256, C8, N), dim3(256), 0, s, src, dst, C, H, W);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
template int launch_pack_nchw<float, float>(const float *, float *, int, int, int, int, cudaStream_t, int);
template int launch_pack_nchw<__half, float>(const float *, __half *, int, int, int, int, cudaStream_t, int);
template int launch_pack_nchw<float, __half>(const __half *, float *, int, int, int, int, cudaStream_t, int);
template int launch_pack_nchw<__half, __half>(const __half *, __half *, int, int, int, int, cudaStream_t, int);
template int launch_pack_nchw<__half, __nv_bfloat16>(const __nv_bfloat16 *, __half *, int, int, int, int, cudaStream_t, int);

template <typename T, typename Tout>
__global__ void unpack_nchw_kernel(const T *__restrict__ src, Tout *__restrict__ dst, int C, int H, int W,
                                   long long total) {
    const int C8 = (C + 7) / 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        long long r = i / W;
        const int y = (int)(r % H);
        r /= H;
        const int ch = (int)(r % C);
        const long long n = r / C;
        dst[i] = (Tout)to_f<T>(src[((((n * C8 + ch / 8) * H + y) * W) + x) * 8 + ch % 8]);
    }
}
template <typename T, typename Tout>
int launch_unpack_nchw(const T *src, Tout *dst, int N, int C, int H, int W, cudaStream_t s) {
    const long long total = (long long)N * C * H * W;
    if (total == 0) return RVSR_OK;
    unpack_nchw_kernel<T, Tout><<<(int)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192), 256, 0, s>>>(
        src, dst, C, H, W, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
template int launch_unpack_nchw<float, float>(const float *, float *, int, int, int, int, cudaStream_t);
template int launch_unpack_nchw<__half, float>(const __half *, float *, int, int, int, int, cudaStream_t);
template int launch_unpack_nchw<__half, __half>(const __half *, __half *, int, int, int, int, cudaStream_t);
template int launch_unpack_nchw<__half, __nv_bfloat16>(const __half *, __nv_bfloat16 *, int, int, int, int, cudaStream_t);

__global__ void convert_f16_f32_kernel(const __half *__restrict__ src, float *__restrict__ dst, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = __half2float(src[i]);
}
int launch_convert_f16_f32(const void *src, float *dst, long long n, cudaStream_t s) {
    if (n == 0) return RVSR_OK;
    convert_f16_f32_kernel<<<(int)((n + 255) / 256 < 8192 ? (n + 255) / 256 : 8192), 256, 0, s>>>(
        reinterpret_cast<const __half *>(src), dst, n);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
// planar offsets / mask -> OUT_OM24: per pixel and deformable group three 32-byte blocks
// [dy0 dx0 .. dy3 dx3][dy4 dx4 .. dy7 dx7][dy8 dx8 m01 m23 m45 m67 m8_ 0], offsets fp32, mask as fp16 pairs
template <typename Tin>
__global__ void om24_from_planar_kernel(const Tin *__restrict__ off, const Tin *__restrict__ msk, uint4 *__restrict__ om, int dg, int HW) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const int g = blockIdx.y;
    const long long n = blockIdx.z;
    const Tin *op = off + (n * dg * 18 + g * 18) * (long long)HW + pix;
    const Tin *mp = msk + (n * dg * 9 + g * 9) * (long long)HW + pix;
    float o[18], m[10];
#pragma unroll
    for (int j = 0; j < 18; ++j) o[j] = to_f<Tin>(op[(long long)j * HW]);
#pragma unroll
    for (int j = 0; j < 9; ++j) m[j] = to_f<Tin>(mp[(long long)j * HW]);
    m[9] = 0.f;
    uint32_t mw[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const __half2 h = __floats2half2_rn(m[2 * k], m[2 * k + 1]);
        mw[k] = *reinterpret_cast<const uint32_t *>(&h);
    }
    uint4 *d = om + ((n * dg * 3 + g * 3) * (long long)HW + pix) * 2;
    auto f = [](float v) { return __float_as_uint(v); };
    d[0] = make_uint4(f(o[0]), f(o[1]), f(o[2]), f(o[3])); d[1] = make_uint4(f(o[4]), f(o[5]), f(o[6]), f(o[7]));
    d += (long long)HW * 2;
    d[0] = make_uint4(f(o[8]), f(o[9]), f(o[10]), f(o[11])); d[1] = make_uint4(f(o[12]), f(o[13]), f(o[14]), f(o[15]));
    d += (long long)HW * 2;
    d[0] = make_uint4(f(o[16]), f(o[17]), mw[0], mw[1]); d[1] = make_uint4(mw[2], mw[3], mw[4], 0u);
}
int launch_om24_from_planar(const void *offset, const void *mask, int dtype, void *om24, int N, int dg, int H, int W, cudaStream_t s) {
    if ((long long)N * dg * H * W == 0) return RVSR_OK;
    const dim3 grid((H * W + 127) / 128, dg, N);
    if (dtype == RVSR_F16)
        om24_from_planar_kernel<__half><<<grid, 128, 0, s>>>((const __half *)offset, (const __half *)mask, (uint4 *)om24, dg, H * W);
    else if (dtype == RVSR_BF16)
        om24_from_planar_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>((const __nv_bfloat16 *)offset, (const __nv_bfloat16 *)mask, (uint4 *)om24, dg, H * W);
    else
        om24_from_planar_kernel<float><<<grid, 128, 0, s>>>((const float *)offset, (const float *)mask, (uint4 *)om24, dg, H * W);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
__global__ void convert_bf16_f32_kernel(const __nv_bfloat16 *__restrict__ src, float *__restrict__ dst, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __bfloat162float(src[i]);
}
__global__ void convert_f32_bf16_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16_rn(src[i]);
}
int launch_convert_bf16_f32(const void *src, float *dst, long long n, cudaStream_t s) {
    if (n == 0) return RVSR_OK;
    convert_bf16_f32_kernel<<<(int)((n + 255) / 256 < 8192 ? (n + 255) / 256 : 8192), 256, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16 *>(src), dst, n);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
int launch_convert_f32_bf16(const float *src, void *dst, long long n, cudaStream_t s) {
    if (n == 0) return RVSR_OK;
    convert_f32_bf16_kernel<<<(int)((n + 255) / 256 < 8192 ? (n + 255) / 256 : 8192), 256, 0, s>>>(
        src, reinterpret_cast<__nv_bfloat16 *>(dst), n);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
__global__ void fill_f32_kernel(float *dst, float v, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = v;
}
int launch_fill_f32(float *dst, float v, long long n, cudaStream_t s) {
    if (n == 0) return RVSR_OK;
    fill_f32_kernel<<<(int)((n + 255) / 256 < 8192 ? (n + 255) / 256 : 8192), 256, 0, s>>>(dst, v, n);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

// ---------------------------------------------------------------- x2 bilinear upsample (align_corners=False)
// F.interpolate(scale_factor=2, mode='bilinear') optionally times `scale`
// (EDVR_arch.py:111-112: upsampled offsets are multiplied by 2 after interpolation).
// One thread per SOURCE pixel of one channel block: it produces the 2x2 output quad from the 3x3 source
// neighbourhood (9 x 16 B loads for 4 x 16 B stores instead of 16 loads; the two outputs of a row are one
// contiguous 32 B).  F.interpolate(scale_factor=2, bilinear, align_corners=False): output 2k reads sources
// (k-1, k) with weights (0.25, 0.75), output 2k+1 reads (k, k+1) with (0.75, 0.25); indices clamp at the borders
// (EDVR_arch.py:111-112, :120-121 multiply the upsampled OFFSETS by 2 -- `scale`).  grid.y = (image, block) plane.
template <typename T>
__global__ void upsample2x_kernel(const T *__restrict__ src, T *__restrict__ dst, int H, int W, float scale, const T *__restrict__ add) {
    pdl_trigger();
    pdl_wait();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * W) return;
    const int y = idx / W, x = idx - y * W;
    const long long pl = blockIdx.y;
    const T *p = src + pl * H * W * 8;
    const int ym = max(y - 1, 0), yp = min(y + 1, H - 1), xm = max(x - 1, 0), xp = min(x + 1, W - 1);
    const int rows[3] = {ym, y, yp};
    float hl[3][8], hr[3][8];  // horizontally blended rows: left output (2x) and right output (2x + 1)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float a[8], b[8], c[8];
        const T *row = p + (long long)rows[r] * W * 8;
        load8<T>(row + xm * 8, a);
        load8<T>(row + x * 8, b);
        load8<T>(row + xp * 8, c);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            hl[r][k] = 0.25f * a[k] + 0.75f * b[k];
            hr[r][k] = 0.75f * b[k] + 0.25f * c[k];
        }
    }
    const long long o0 = (pl * (2 * H) + 2 * y) * (long long)(2 * W) * 8 + (long long)(2 * x) * 8;
    T *o = dst + o0;
    const T *ad = add != nullptr ? add + o0 : nullptr;  // optional: dst = scale * up(src) + add (Predeblur pyramid, EDVR_arch.py:53-57)
    float t[8], u[8];
    auto put = [&](long long off) {
        if (ad != nullptr) {
            load8<T>(ad + off, u);
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] += u[k];
        }
        store8<T>(o + off, t);
    };
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = scale * (0.25f * hl[0][k] + 0.75f * hl[1][k]);
    put(0);
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = scale * (0.25f * hr[0][k] + 0.75f * hr[1][k]);
    put(8);
    const long long row = (long long)(2 * W) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = scale * (0.75f * hl[1][k] + 0.25f * hl[2][k]);
    put(row);
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = scale * (0.75f * hr[1][k] + 0.25f * hr[2][k]);
    put(row + 8);
}

template <typename T>
int launch_upsample2x(const T *src, T *dst, int N, int C, int H, int W, float scale, cudaStream_t s, const T *add) {
    const int planes = N * ((C + 7) / 8);
    if (planes == 0 || H == 0 || W == 0) return RVSR_OK;
    RVSR_CHECK_ARG(planes <= 65535, "upsample2x: too many (image, channel block) planes: %d", planes);
    launch_k(upsample2x_kernel<T>, dim3((H * W + 127) / 128, planes), dim3(128), 0, s, src, dst, H, W, scale, add);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
template int launch_upsample2x<float>(const float *, float *, int, int, int, int, float, cudaStream_t, const float *);
template int launch_upsample2x<__half>(const __half *, __half *, int, int, int, int, float, cudaStream_t, const __half *);

// ---------------------------------------------------------------- MaxPool2d(3,2,1) + AvgPool2d(3,2,1) in one pass
// max pads with -inf; avg divides by 9 always (count_include_pad=True) -- EDVR_arch.py:154-155.
template <typename T>
__global__ void pool_maxavg_kernel(const T *__restrict__ src, T *__restrict__ dmax, T *__restrict__ davg, int H,
                                   int W, long long total) {
    pdl_trigger();
    pdl_wait();
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wo);
        long long r = i / Wo;
        const int oy = (int)(r % Ho);
        const long long pl = r / Ho;
        const T *p = src + pl * H * W * 8;
        float mx[8], sm[8], v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { mx[k] = -INFINITY; sm[k] = 0.f; }
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int y = 2 * oy + dy, x = 2 * ox + dx;
                if (y < 0 || y >= H || x < 0 || x >= W) continue;
                load8<T>(p + ((long long)y * W + x) * 8, v);
#pragma unroll
                for (int k = 0; k < 8; ++k) { mx[k] = fmaxf(mx[k], v[k]); sm[k] += v[k]; }
            }
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[k] *= (1.f / 9.f);
        store8<T>(dmax + i * 8, mx);
        store8<T>(davg + i * 8, sm);
    }
}
template <typename T>
int launch_pool_maxavg(const T *src, T *dst_max, T *dst_avg, int N, int C, int H, int W, cudaStream_t s) {
    const long long total = (long long)N * ((C + 7) / 8) * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1);
    if (total == 0) return RVSR_OK;
    launch_k(pool_maxavg_kernel<T>, dim3((int)((total + 255) / 256 < 16384 ? (total + 255) / 256 : 16384)), dim3(256), 0, s, src, dst_max, dst_avg, H, W, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
template int launch_pool_maxavg<float>(const float *, float *, float *, int, int, int, int, cudaStream_t);
template int launch_pool_maxavg<__half>(const __half *, __half *, __half *, int, int, int, int, cudaStream_t);

// ---------------------------------------------------------------- TSA temporal attention (EDVR_arch.py:175-181)
// One warp handles 32 consecutive pixels of one (batch, frame); each lane owns a pixel and
// walks the channel blocks: cor = sum_c emb[c]*emb_ref[c]; out = aligned * sigmoid(cor).
template <typename T>
__global__ void tsa_temporal_kernel(const T *__restrict__ emb, const T *__restrict__ emb_ref,
                                    const T *__restrict__ aligned, T *__restrict__ out, int frames, int C8,
                                    long long HW, long long total) {
    pdl_trigger();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i % HW;
        const long long bn = i / HW;  // b * frames + frame
        const long long b = bn / frames;
        const T *e = emb + bn * C8 * HW * 8, *er = emb_ref + b * C8 * HW * 8;
        float cor = 0.f, u[8], w[8];
        for (int q = 0; q < C8; ++q) {
            load8<T>(e + (q * HW + pix) * 8, u);
            load8<T>(er + (q * HW + pix) * 8, w);
#pragma unroll
            for (int k = 0; k < 8; ++k) cor = fmaf(u[k], w[k], cor);
        }
        const float pr = sigmoidf_(cor);
        const T *a = aligned + bn * C8 * HW * 8;
        T *o = out + bn * C8 * HW * 8;
        for (int q = 0; q < C8; ++q) {
            load8<T>(a + (q * HW + pix) * 8, u);
#pragma unroll
            for (int k = 0; k < 8; ++k) u[k] *= pr;
            store8<T>(o + (q * HW + pix) * 8, u);
        }
    }
}
template <typename T>
int launch_tsa_temporal(const T *emb, const T *emb_ref, const T *aligned, T *out, int B, int frames, int C,
                        int H, int W, cudaStream_t s) {
    const long long HW = (long long)H * W, total = (long long)B * frames * HW;
    if (total == 0) return RVSR_OK;
    launch_k(tsa_temporal_kernel<T>, dim3((int)((total + 127) / 128 < 16384 ? (total + 127) / 128 : 16384)), dim3(128), 0, s, emb, emb_ref, aligned, out, frames, (C + 7) / 8, HW, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
template int launch_tsa_temporal<float>(const float *, const float *, const float *, float *, int, int, int,
                                        int, int, cudaStream_t);
template int launch_tsa_temporal<__half>(const __half *, const __half *, const __half *, __half *, int, int,
                                         int, int, int, cudaStream_t);

// fea * sigmoid(att) * 2 + att_add   (EDVR_arch.py:205-207)
template <typename T>
__global__ void tsa_final_kernel(const T *__restrict__ fea, const T *__restrict__ att,
                                 const T *__restrict__ att_add, T *__restrict__ out, long long n8) {
    pdl_trigger();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
         i += (long long)gridDim.x * blockDim.x) {
        float f[8], a[8], d[8];
        load8<T>(fea + i * 8, f);
        load8<T>(att + i * 8, a);
        load8<T>(att_add + i * 8, d);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = f[k] * sigmoidf_(a[k]) * 2.f + d[k];
        store8<T>(out + i * 8, f);
    }
}
template <typename T>
int launch_tsa_final(const T *fea, const T *att, const T *att_add, T *out, long long n, cudaStream_t s) {
    const long long n8 = n / 8;
    if (n8 == 0) return RVSR_OK;
    launch_k(tsa_final_kernel<T>, dim3((int)((n8 + 255) / 256 < 16384 ? (n8 + 255) / 256 : 16384)), dim3(256), 0, s, fea, att, att_add, out, n8);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
template int launch_tsa_final<float>(const float *, const float *, const float *, float *, long long, cudaStream_t);
template int launch_tsa_final<__half>(const __half *, const __half *, const __half *, __half *, long long,
                                      cudaStream_t);

// ---------------------------------------------------------------- out = conv_last + base  (EDVR_arch.py:314-319, :401-403)
// res_c8: [B][1][sH][sW][8] (nc <= 8 channels used); x: [B][frames][nc][H][W] NCHW;
// base = x4 bilinear (align_corners=False) of the centre LQ frame, or the frame itself (scale 1).
template <typename T, typename Tin, typename Tout>
__global__ void final_add_kernel(const T *__restrict__ res, const Tin *__restrict__ x, Tout *__restrict__ out,
                                 int frames, int center, int nc, int H, int W, int scale, long long total,
                                 const int *__restrict__ center_map) {
    pdl_trigger();
    pdl_wait();
    const int Ho = H * scale, Wo = W * scale;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wo);
        long long r = i / Wo;
        const int oy = (int)(r % Ho);
        const long long b = r / Ho;
        float v[8];
        load8<T>(res + i * 8, v);
        const long long cimg = center_map != nullptr ? (long long)__ldg(center_map + b) : b * frames + center;
        const Tin *xc = x + (cimg * nc) * (long long)H * W;
        if (scale == 1) {
            for (int c = 0; c < nc; ++c)
                out[((b * nc + c) * Ho + oy) * (long long)Wo + ox] =
                    (Tout)(v[c] + to_f<Tin>(xc[((long long)c * H + oy) * W + ox]));
        } else {
            const float inv = 1.f / scale;
            const float sy = fmaxf(inv * (oy + 0.5f) - 0.5f, 0.f), sx = fmaxf(inv * (ox + 0.5f) - 0.5f, 0.f);
            const int y0 = (int)sy, x0 = (int)sx;
            const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
            const float ly = sy - y0, lx = sx - x0;
            for (int c = 0; c < nc; ++c) {
                const Tin *pc = xc + (long long)c * H * W;
                const float base =
                    (1.f - ly) * ((1.f - lx) * to_f<Tin>(pc[y0 * W + x0]) + lx * to_f<Tin>(pc[y0 * W + x1])) +
                    ly * ((1.f - lx) * to_f<Tin>(pc[y1 * W + x0]) + lx * to_f<Tin>(pc[y1 * W + x1]));
                out[((b * nc + c) * Ho + oy) * (long long)Wo + ox] = (Tout)(v[c] + base);
            }
        }
    }
}
template <typename Tout> __device__ __forceinline__ Tout cast_out(float v);
template <typename T, typename Tin, typename Tout>
int launch_final_add(const T *res_c8, const Tin *x, Tout *out, int B, int frames, int center, int nc, int H,
                     int W, int scale, cudaStream_t s, const int *center_map) {
    RVSR_CHECK_ARG(nc <= 8, "final_add: nc %d > 8", nc);
    const long long total = (long long)B * H * scale * W * scale;
    if (total == 0) return RVSR_OK;
    launch_k(final_add_kernel<T, Tin, Tout>, dim3((int)((total + 255) / 256 < 16384 ? (total + 255) / 256 : 16384)), dim3(256), 0, s, res_c8, x, out, frames, center, nc, H, W, scale, total, center_map);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
#define INST_FINAL(T, Tin, Tout) \
    template int launch_final_add<T, Tin, Tout>(const T *, const Tin *, Tout *, int, int, int, int, int, int, int, cudaStream_t, const int *);
INST_FINAL(float, float, float)
INST_FINAL(float, __half, float)
INST_FINAL(float, float, __half)
INST_FINAL(float, __half, __half)
INST_FINAL(__half, float, float)
INST_FINAL(__half, __half, float)
INST_FINAL(__half, float, __half)
INST_FINAL(__half, __half, __half)

}  // namespace rvsr

// zero-pad the input-channel dimension of an OIHW weight (conv_first: 3 -> 16 for the tensor-core path)
namespace rvsr {
__global__ void pad_cin_kernel(const float *__restrict__ w, float *__restrict__ dst, int Cin, int Cpad, int KK, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % KK);
        const int c = (int)((i / KK) % Cpad);
        const long long co = i / ((long long)KK * Cpad);
        dst[i] = c < Cin ? w[(co * Cin + c) * KK + t] : 0.f;
    }
}
int pad_weight_cin(const float *w, float *dst, int Cout, int Cin, int Cpad, int KK, cudaStream_t s) {
    const long long total = (long long)Cout * Cpad * KK;
    pad_cin_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(w, dst, Cin, Cpad, KK, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
}  // namespace rvsr

// ---------------------------------------------------------------- grouped -> dense weight (operator-level API only)
namespace rvsr {
__global__ void expand_grouped_weight_kernel(const float *__restrict__ w, float *__restrict__ dst, int Cout, int C,
                                             int K, int groups, long long total) {
    const int cin_g = C / groups, cout_g = Cout / groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % K);
        const int c = (int)((i / K) % C);
        const int co = (int)(i / ((long long)K * C));
        dst[i] = (c / cin_g == co / cout_g) ? w[((long long)co * cin_g + c % cin_g) * K + t] : 0.f;
    }
}
int expand_grouped_weight(const float *w, float *dst, int Cout, int C, int K, int groups, cudaStream_t s) {
    const long long total = (long long)Cout * C * K;
    expand_grouped_weight_kernel<<<(int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096), 256, 0, s>>>(
        w, dst, Cout, C, K, groups, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
}  // namespace rvsr

// ---------------------------------------------------------------- modulated deformable conv, backward (fp32)
// Reference: modulated_deform_conv_cuda_backward (deform_conv_cuda.cpp:571-685) with its three kernels
// (deform_conv_cuda_kernel.cu:635-767) and three GEMMs per sample.  Here one kernel per
// (64-pixel tile, 8-channel block): it rebuilds the tile's modulated samples, forms
// grad_col = W^T . grad_out for the tile in shared memory, and from that produces grad_input (atomic
// scatter to <= 4 corners), grad_offset / grad_mask (atomic: several channel blocks can share a
// deformable group), grad_weight (tile-local grad_out . col^T, then atomics) and grad_bias.
// No columns buffer, no per-sample loop.  Summation order is nondeterministic like the reference's
// atomicAdd col2im (deform_conv_cuda_kernel.cu:688).
namespace rvsr {

template <int K>
__global__ void __launch_bounds__(NTHREADS) dcn_bwd_simt_kernel(const DcnBwdOp op) {
    constexpr int ROWS = K * 8, RP = 5, CO = 64;  // rows per thread group (16 groups x 5 >= 72)
    extern __shared__ __align__(16) float sm[];
    float(*col)[TILE_P + 1] = reinterpret_cast<float(*)[TILE_P + 1]>(sm);                       // m * sample
    float(*gcol)[TILE_P + 1] = reinterpret_cast<float(*)[TILE_P + 1]>(sm + ROWS * (TILE_P + 1));
    float(*go)[TILE_P + 1] = reinterpret_cast<float(*)[TILE_P + 1]>(sm + 2 * ROWS * (TILE_P + 1));  // [CO][px]
    float(*wt)[ROWS + 1] = reinterpret_cast<float(*)[ROWS + 1]>(sm + (2 * ROWS + CO) * (TILE_P + 1));  // [CO][row]
    const int Ho = (op.H + 2 * op.pad - (op.dil * (op.kh - 1) + 1)) / op.stride + 1;
    const int Wo = (op.W + 2 * op.pad - (op.dil * (op.kw - 1) + 1)) / op.stride + 1;
    const int ntx = (Wo + TILE_W - 1) / TILE_W;
    const int ox0 = (blockIdx.x % ntx) * TILE_W, oy0 = (blockIdx.x / ntx) * TILE_H;
    const int q = blockIdx.y, n = blockIdx.z;
    const int C = op.C, C8 = (C + 7) / 8, cpg = C / op.dg;
    const long long plane = (long long)Ho * Wo, iplane = (long long)op.H * op.W;
    const float *xq = op.x_c8 + ((long long)n * C8 + q) * iplane * 8;
    float *gxq = op.gx_c8 + ((long long)n * C8 + q) * iplane * 8;
    const float *off = op.offset + (long long)n * op.dg * 2 * K * plane;
    const float *msk = op.mask + (long long)n * op.dg * K * plane;
    const float *gout = op.gout + (long long)n * op.Cout * plane;

    // 1. modulated samples of this tile / channel block (same producer as the forward kernel)
    for (int i = threadIdx.x; i < TILE_P * K; i += NTHREADS) {
        const int p = i % TILE_P, t = i / TILE_P;
        const int oy = oy0 + p / TILE_W, ox = ox0 + p % TILE_W;
        float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (oy < Ho && ox < Wo) {
            const long long pix = (long long)oy * Wo + ox;
            const float by = (float)(oy * op.stride - op.pad + (t / op.kw) * op.dil);
            const float bx = (float)(ox * op.stride - op.pad + (t % op.kw) * op.dil);
            int gprev = -1;
            float m = 0.f, v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int ch = q * 8 + c;
                if (ch >= C) break;
                const int g = ch / cpg;
                if (g != gprev) {
                    gprev = g;
                    m = __ldg(msk + ((long long)g * K + t) * plane + pix);
                    sample8<float>(xq, op.H, op.W, by + __ldg(off + ((long long)g * 2 * K + 2 * t) * plane + pix),
                                   bx + __ldg(off + ((long long)g * 2 * K + 2 * t + 1) * plane + pix), v);
                }
                r[c] = m * v[c];
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) col[t * 8 + c][p] = r[c];
    }
    // 2. loop over output-channel slices: grad_col += W^T gout ; grad_weight += gout col^T ; grad_bias
    const int pg = threadIdx.x % 16, rg = threadIdx.x / 16;
    float acc[RP][4] = {};
    for (int cs = 0; cs < op.Cout; cs += CO) {
        __syncthreads();
        for (int i = threadIdx.x; i < CO * TILE_P; i += NTHREADS) {
            const int o = i / TILE_P, p = i % TILE_P;
            const int oy = oy0 + p / TILE_W, ox = ox0 + p % TILE_W;
            go[o][p] = (cs + o < op.Cout && oy < Ho && ox < Wo) ? __ldg(gout + (long long)(cs + o) * plane + (long long)oy * Wo + ox) : 0.f;
        }
        for (int i = threadIdx.x; i < CO * ROWS; i += NTHREADS) {
            const int o = i / ROWS, r = i % ROWS, t = r / 8, ch = q * 8 + r % 8;
            wt[o][r] = (cs + o < op.Cout && ch < C) ? __ldg(op.w_dense + ((long long)(cs + o) * C + ch) * K + t) : 0.f;
        }
        __syncthreads();
        for (int o = 0; o < CO; ++o) {
            float a[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = go[o][pg * 4 + k];
#pragma unroll
            for (int j = 0; j < RP; ++j) {
                const int r = rg * RP + j;
                const float w = r < ROWS ? wt[o][r] : 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[j][k] = fmaf(w, a[k], acc[j][k]);
            }
        }
        // grad_weight: thread -> 4 output channels x 5 rows, reduce over the tile's 64 pixels
        {
            const int og = threadIdx.x / 16, rr = threadIdx.x % 16;
            float gw[4][RP] = {};
            for (int p = 0; p < TILE_P; ++p) {
                float cv[RP];
#pragma unroll
                for (int j = 0; j < RP; ++j) cv[j] = rr * RP + j < ROWS ? col[rr * RP + j][p] : 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float g = go[og * 4 + i][p];
#pragma unroll
                    for (int j = 0; j < RP; ++j) gw[i][j] = fmaf(g, cv[j], gw[i][j]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < RP; ++j) {
                    const int o = cs + og * 4 + i, r = rr * RP + j;
                    if (o < op.Cout && r < ROWS && q * 8 + r % 8 < C && gw[i][j] != 0.f)
                        atomicAdd(op.gw_dense + ((long long)o * C + q * 8 + r % 8) * K + r / 8, gw[i][j]);
                }
        }
        if (q == 0 && op.gbias != nullptr && threadIdx.x < CO && cs + threadIdx.x < op.Cout) {
            float sb = 0.f;
            for (int p = 0; p < TILE_P; ++p) sb += go[threadIdx.x][p];
            atomicAdd(op.gbias + cs + threadIdx.x, sb);
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RP; ++j)
        if (rg * RP + j < ROWS)
#pragma unroll
            for (int k = 0; k < 4; ++k) gcol[rg * RP + j][pg * 4 + k] = acc[j][k];
    __syncthreads();
    // 3. scatter: per (pixel, tap) walk the block's channels; coordinates once per deformable group
    for (int i = threadIdx.x; i < TILE_P * K; i += NTHREADS) {
        const int p = i % TILE_P, t = i / TILE_P;
        const int oy = oy0 + p / TILE_W, ox = ox0 + p % TILE_W;
        if (oy >= Ho || ox >= Wo) continue;
        const long long pix = (long long)oy * Wo + ox;
        const float by = (float)(oy * op.stride - op.pad + (t / op.kw) * op.dil);
        const float bx = (float)(ox * op.stride - op.pad + (t % op.kw) * op.dil);
        int c = 0;
        while (c < 8 && q * 8 + c < C) {
            const int g = (q * 8 + c) / cpg;
            int cend = c;
            while (cend < 8 && q * 8 + cend < C && (q * 8 + cend) / cpg == g) ++cend;
            const long long oc = ((long long)g * 2 * K + 2 * t) * plane + pix, mc = ((long long)g * K + t) * plane + pix;
            const float py = by + __ldg(off + oc), px = bx + __ldg(off + oc + plane), m = __ldg(msk + mc);
            if (py > -1.f && px > -1.f && py < (float)op.H && px < (float)op.W) {
                const float fy = floorf(py), fx = floorf(px);
                const int y0 = (int)fy, x0 = (int)fx, y1 = y0 + 1, x1 = x0 + 1;
                const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
                const bool vy0 = y0 >= 0, vy1 = y1 <= op.H - 1, vx0 = x0 >= 0, vx1 = x1 <= op.W - 1;
                float v00[8] = {}, v01[8] = {}, v10[8] = {}, v11[8] = {};
                if (vy0 && vx0) load8<float>(xq + ((long long)y0 * op.W + x0) * 8, v00);
                if (vy0 && vx1) load8<float>(xq + ((long long)y0 * op.W + x1) * 8, v01);
                if (vy1 && vx0) load8<float>(xq + ((long long)y1 * op.W + x0) * 8, v10);
                if (vy1 && vx1) load8<float>(xq + ((long long)y1 * op.W + x1) * 8, v11);
                float g_dy = 0.f, g_dx = 0.f, g_m = 0.f;
                const bool whole = c == 0 && cend == 8;  // the group covers the whole channel block (EDVR: 8 channels per group):
                float gmv[8];                            // its 8 contributions to a corner go out as two vector reductions
                for (int cc = c; cc < cend; ++cc) {
                    const float gc = gcol[t * 8 + cc][p];
                    const float val = hy * (hx * v00[cc] + lx * v01[cc]) + ly * (hx * v10[cc] + lx * v11[cc]);
                    g_m += gc * val;
                    g_dy += gc * m * (hx * (v10[cc] - v00[cc]) + lx * (v11[cc] - v01[cc]));
                    g_dx += gc * m * (hy * (v01[cc] - v00[cc]) + ly * (v11[cc] - v10[cc]));
                    const float gm = gc * m;
                    gmv[cc] = gm;
                    if (whole) continue;
                    if (vy0 && vx0) atomicAdd(gxq + ((long long)y0 * op.W + x0) * 8 + cc, gm * hy * hx);
                    if (vy0 && vx1) atomicAdd(gxq + ((long long)y0 * op.W + x1) * 8 + cc, gm * hy * lx);
                    if (vy1 && vx0) atomicAdd(gxq + ((long long)y1 * op.W + x0) * 8 + cc, gm * ly * hx);
                    if (vy1 && vx1) atomicAdd(gxq + ((long long)y1 * op.W + x1) * 8 + cc, gm * ly * lx);
                }
                if (whole) {  // 32 scalar atomics -> 8 red.global.add.v4.f32 (same values, same fp32 adds)
                    auto corner = [&](bool ok, int yy, int xx, float w) {
                        if (!ok) return;
                        float *d = gxq + ((long long)yy * op.W + xx) * 8;
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(gmv[0] * w), "f"(gmv[1] * w), "f"(gmv[2] * w),
                                     "f"(gmv[3] * w) : "memory");
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4), "f"(gmv[4] * w), "f"(gmv[5] * w), "f"(gmv[6] * w),
                                     "f"(gmv[7] * w) : "memory");
                    };
                    corner(vy0 && vx0, y0, x0, hy * hx);
                    corner(vy0 && vx1, y0, x1, hy * lx);
                    corner(vy1 && vx0, y1, x0, ly * hx);
                    corner(vy1 && vx1, y1, x1, ly * lx);
                }
                atomicAdd(op.goffset + (long long)n * op.dg * 2 * K * plane + oc, g_dy);
                atomicAdd(op.goffset + (long long)n * op.dg * 2 * K * plane + oc + plane, g_dx);
                atomicAdd(op.gmask + (long long)n * op.dg * K * plane + mc, g_m);
            }
            c = cend;
        }
    }
}

// dense [Cout][C][K] gradient -> grouped [Cout][C/groups][K], accumulated
__global__ void fold_grouped_weight_kernel(const float *__restrict__ gd, float *__restrict__ gw, int Cout, int C, int K,
                                           int groups, long long total) {
    const int cin_g = C / groups, cout_g = Cout / groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % K);
        const int cl = (int)((i / K) % cin_g);
        const int co = (int)(i / ((long long)K * cin_g));
        gw[i] += gd[((long long)co * C + (co / cout_g) * cin_g + cl) * K + t];
    }
}

int launch_dcn_bwd_simt(const DcnBwdOp &op, cudaStream_t s) {
    const int K = op.kh * op.kw;
    RVSR_CHECK_ARG(K == 9 || K == 1, "dcn bwd: only 3x3 and 1x1 kernels are built");
    const int Ho = (op.H + 2 * op.pad - (op.dil * (op.kh - 1) + 1)) / op.stride + 1;
    const int Wo = (op.W + 2 * op.pad - (op.dil * (op.kw - 1) + 1)) / op.stride + 1;
    dim3 grid(cdiv(Wo, TILE_W) * cdiv(Ho, TILE_H), cdiv(op.C, 8), op.N);
    if (grid.z == 0) return RVSR_OK;
    RVSR_CHECK_ARG(grid.z <= 65535 && grid.y <= 65535, "dcn bwd: too many images / channels");
    const int ROWS = K * 8;
    const size_t smem = ((size_t)(2 * ROWS + 64) * (TILE_P + 1) + (size_t)64 * (ROWS + 1)) * sizeof(float);
    if (K == 9) {
        RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&dcn_bwd_simt_kernel<9>), (int)smem));
        dcn_bwd_simt_kernel<9><<<grid, NTHREADS, smem, s>>>(op);
    } else {
        dcn_bwd_simt_kernel<1><<<grid, NTHREADS, smem, s>>>(op);
    }
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
int fold_grouped_weight(const float *gd, float *gw, int Cout, int C, int K, int groups, cudaStream_t s) {
    const long long total = (long long)Cout * (C / groups) * K;
    fold_grouped_weight_kernel<<<(int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096), 256, 0, s>>>(gd, gw, Cout, C, K, groups, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

// ---------------------------------------------------------------- image I/O around the model
// Ingest: what data/util.py::read_img (:87-101, uint8 -> float32 / 255) + read_img_seq (:104-122, channel reversal
// [2, 1, 0], HWC -> CHW, stack) do on the host, for T frames at once.
template <typename Tout>
__global__ void frames_from_u8_kernel(const uint8_t *__restrict__ src, Tout *__restrict__ dst, int C, int H, int W,
                                      int reverse, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        long long r = i / W;
        const int y = (int)(r % H);
        r /= H;
        const int c = (int)(r % C);
        const long long t = r / C;
        const int cs = reverse ? C - 1 - c : c;
        const float v = (float)src[((t * H + y) * W + x) * C + cs] / 255.f;  // float32 division, like numpy
        dst[i] = from_f<Tout>(v);
    }
}
template <typename Tout>
int launch_frames_from_u8(const uint8_t *src, Tout *dst, int T, int C, int H, int W, int reverse, cudaStream_t s) {
    const long long total = (long long)T * C * H * W;
    if (total == 0) return RVSR_OK;
    frames_from_u8_kernel<Tout><<<(int)((total + 255) / 256 < 16384 ? (total + 255) / 256 : 16384), 256, 0, s>>>(src, dst, C, H, W, reverse, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
template int launch_frames_from_u8<float>(const uint8_t *, float *, int, int, int, int, int, cudaStream_t);
template int launch_frames_from_u8<__half>(const uint8_t *, __half *, int, int, int, int, int, cudaStream_t);

// Egress: network output [B, 3, H, W] -> uint8 BGR images [B, H, W, 3], the arithmetic of the reference's test loop:
//   mode 0 (RGB model)   utils/util.py::tensor2img(out_type=uint8, reverse_channel=True) (:151-181):
//                        clamp to [0, 1], RGB -> BGR, (x * 255.0).round() in float32
//   mode 1 (YCbCr model) tensor2img(out_type=float32, reverse_channel=False), then data/util.py::ycbcr2bgr (:397-416)
//                        and (np.clip(., 0, 1) * 255.).round() (test_RealVSR_wi_GT.py:122-123).  ycbcr2bgr multiplies
//                        the float32 image by a float64 matrix: that part runs in double here as well.
// np.round / ndarray.round are round-half-to-even: rintf / rint.
template <typename Tin>
__global__ void frames_to_u8_kernel(const Tin *__restrict__ src, uint8_t *__restrict__ dst, int H, int W, int mode, long long total) {
    const long long plane = (long long)H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / plane, pix = i % plane;
        const Tin *p = src + b * 3 * plane + pix;
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = fminf(fmaxf(to_f<Tin>(p[c * plane]), 0.f), 1.f);  // clamp_(0, 1); (x - 0) / (1 - 0)
        uint8_t *o = dst + i * 3;
        if (mode == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) o[2 - c] = (uint8_t)rintf(v[c] * 255.0f);
        } else {
            const double M[3][3] = {{0.00456621, 0.00456621, 0.00456621}, {0.00791071, -0.00153632, 0}, {0, -0.00318811, 0.00625893}};
            const double off[3] = {-276.836, 135.576, -222.921};
            float img[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) img[c] = v[c] * 255.f;  // img *= 255. on the float32 array
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                // numpy matmul (float32 promoted to float64), K = 3: plain multiply-adds in index order, no fma contraction
                double acc = __dmul_rn((double)img[0], M[0][k]);
                acc = __dadd_rn(acc, __dmul_rn((double)img[1], M[1][k]));
                acc = __dadd_rn(acc, __dmul_rn((double)img[2], M[2][k]));
                const double rlt = __ddiv_rn(__dadd_rn(__dmul_rn(acc, 255.0), off[k]), 255.0);
                const float f = fminf(fmaxf((float)rlt, 0.f), 1.f);  // astype(float32); np.clip(., 0, 1)
                o[k] = (uint8_t)rintf(f * 255.f);
            }
        }
    }
}
template <typename Tin>
int launch_frames_to_u8(const Tin *src, uint8_t *dst, int B, int H, int W, int mode, cudaStream_t s) {
    const long long total = (long long)B * H * W;
    if (total == 0) return RVSR_OK;
    frames_to_u8_kernel<Tin><<<(int)((total + 255) / 256 < 16384 ? (total + 255) / 256 : 16384), 256, 0, s>>>(src, dst, H, W, mode, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
template int launch_frames_to_u8<float>(const float *, uint8_t *, int, int, int, int, cudaStream_t);
template int launch_frames_to_u8<__half>(const __half *, uint8_t *, int, int, int, int, cudaStream_t);

}  // namespace rvsr
