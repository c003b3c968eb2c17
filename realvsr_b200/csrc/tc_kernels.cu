// tc_kernels.cu -- tcgen05 / TMEM / TMA kernels of the rvsr_b200 hot path (sm_100a only).
//
//   conv_tc_kernel<KS, NT>   3x3 / 1x1 convolution as an im2col-free implicit GEMM:
//       * one TMA load per (tile, source) brings a (4+KS-1) x 32-pixel halo of every 8-channel
//         block into shared memory, ONCE; the nine taps are nine shifted *views* of that tile,
//         expressed purely through the UMMA shared-memory descriptor start address;
//       * tcgen05.mma (M=128 pixels, N=NT output channels, K=16 channels) accumulates all taps
//         and all sources (fused torch.cat) into a TMEM accumulator, double buffered;
//       * four epilogue warps drain TMEM (tcgen05.ld), add bias, activation, residual, and
//         store 128-bit channel blocks (optionally pixel-shuffled, subsampled for stride 2,
//         or as planar fp32 + sigmoid for the DCN offset/mask prediction).
//   dcn_tc_kernel            modulated deformable 3x3 conv: gather warps sample the input
//       bilinearly (fp32 coordinates and blend), scale by the mask, and write fp16 A-operand
//       tiles straight into the UMMA core-matrix layout in shared memory -- the reference's
//       9x-inflated `columns` buffer never exists; tcgen05.mma contracts each tap as soon as
//       it is gathered.
//
// Shared-memory operand layout (both kernels): K-major, SWIZZLE_NONE "core matrices" of
// 8 rows x 16 bytes (8 fp16 channels).  Activations are stored channel-blocked
// [N][C/8][H][W][8], so a pixel of a channel block IS one core-matrix row: rows (pixels) are
// 16 B apart, 8-row groups 128 B apart (SBO), channel blocks one plane apart (LBO).  A tap
// (dy, dx) is the same tile with the start address advanced by (dy*32 + dx) * 16 bytes.
//
// Reference semantics: nn.Conv2d sites of EDVR_arch.py (:71-91, :146-164, :229-253) and the
// DCN of dcn/src/deform_conv_cuda_kernel.cu:467-497, :571-633 + deform_conv_cuda.cpp:539-568.
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "tc_common.cuh"

namespace rvsr {

// ---------------------------------------------------------------- debug: kernel-boundary timeline (RVSR_TC_STAMPS=1)
// %globaltimer stamps of CTA 0 of every pair-kernel launch: 0 entry, 1 prologue done (cluster sync), 2 pdl_wait
// returned, 3 first stage full (issuer 0), 4 first accumulator full (epilogue), 5 last epilogue tile done, 6 exit.
constexpr int STAMP_SLOTS = 512;
__device__ unsigned long long g_stamps[STAMP_SLOTS][8];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
static int g_stamp_next = 0;
static char g_stamp_label[STAMP_SLOTS][48];
static bool stamps_on() {
    static const bool on = getenv("RVSR_TC_STAMPS") != nullptr;
    return on;
}
void tc_stamps_dump() {
    if (!stamps_on() || g_stamp_next == 0) return;
    cudaDeviceSynchronize();
    static unsigned long long h[STAMP_SLOTS][8];
    cudaMemcpyFromSymbol(h, g_stamps, sizeof(h));
    const int n = g_stamp_next < STAMP_SLOTS ? g_stamp_next : STAMP_SLOTS;
    printf("[stamps] launch  gap-from-prev-exit | prologue  pdl-wait  first-full  first-acc  ...last-epi  exit | total (us)\n");
    for (int i = 0; i < n; ++i) {
        const unsigned long long *t = h[i];
        const double gap = i > 0 ? ((double)t[0] - (double)h[i - 1][6]) * 1e-3 : 0.0;
        printf("[stamps] %2d %-40s %7.2f | %6.2f %6.2f %6.2f %6.2f %8.2f %6.2f | %8.2f\n", i, g_stamp_label[i], gap,
               (t[1] - t[0]) * 1e-3, (t[2] - t[1]) * 1e-3, (double)((long long)(t[3] - t[2])) * 1e-3, (double)((long long)(t[4] - t[3])) * 1e-3,
               (double)((long long)(t[5] - t[4])) * 1e-3, (double)((long long)(t[6] - t[5])) * 1e-3, (t[6] - t[0]) * 1e-3);
    }
    g_stamp_next = 0;
}

// Pixel-shuffle column order of one N-pass (NT columns = NT/4 channels x 4 sub-pixels ij = 2i + j):
// column n = i * (NT/2) + kblk * 16 + j * 8 + e  <->  conv channel 4 * (pss * NT/4 + kblk * 8 + e) + 2i + j
__host__ __device__ __forceinline__ int shuffle_col_to_channel(int n, int pss, int NT) {
    const int i = n / (NT / 2), r = n % (NT / 2), kblk = r / 16, j = (r % 16) / 8, e = r % 8;
    return 4 * (pss * (NT / 4) + kblk * 8 + e) + 2 * i + j;
}

// ---------------------------------------------------------------- shared epilogue
struct EpiArgs {
    const float *bias_s;  // smem, NT floats for this pass (0 beyond Cout)
    void *out;
    long long out_image_stride;
    const __half *residual;
    long long res_image_stride;
    int H, W, Cout, act, out_mode, sig_from, subsample, dg;
    const FinalAdd *fin;  // OUT_FINAL
    int res_pre, res_div;
    int bf16;  // 16-bit storage format of out / residual: 0 = fp16, 1 = bfloat16 (training path)
    float res_slope;  // res_pre == 2: `residual` is a MASK tensor, out = v * (residual > 0 ? 1 : res_slope) -- the gradient through the
                      // LeakyReLU / ReLU that produced it, fused into the data-gradient convolution feeding it (training path)
};

// 16-bit pair <-> fp32 pair in either storage format
template <bool BF> __device__ __forceinline__ float2 unpack16x2(uint32_t u) {
    if (BF) return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
    return __half22float2(*reinterpret_cast<const __half2 *>(&u));
}
template <bool BF> __device__ __forceinline__ uint32_t pack16x2(float a, float b) {
    if (BF) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<const uint32_t *>(&h);
    }
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// Epilogue of one pixel (= TMEM lane) of one tile.  With NH = 2 two warps share a TMEM lane
// quarter and take the lower / upper half of the NT accumulator columns.  Order of events:
//   1. residual blocks are fetched BEFORE waiting for the accumulator (latency overlaps the MMAs);
//   2. all tcgen05.ld of the warp's columns are issued, one wait, and the TMEM buffer is handed
//      back to the MMA warp immediately;
//   3. bias / activation / residual / fp16 pack / 128-bit stores run out of registers.
// ACT is a template parameter for the common OUT_C8 mode (ACT = -1: runtime e.act).
template <int ACT> __device__ __forceinline__ float act_t(float v, int act_rt) {
    if (ACT == RVSR_ACT_NONE) return v;
    if (ACT == RVSR_ACT_LRELU) return fmaxf(v, 0.1f * v);
    if (ACT == RVSR_ACT_RELU) return fmaxf(v, 0.f);
    return apply_act(v, act_rt);
}

template <int NT, int NH, bool BF = false> struct EpiTile {  // BF: out / residual are bfloat16 (training path) instead of fp16
    static constexpr int HALFC = NH == 1 ? NT : ((NT / NH + 15) / 16) * 16;  // columns per warp (upper bound)
    static constexpr int CW = HALFC >= 32 ? 32 : 16;                         // columns per register chunk
    static constexpr int NCH = HALFC / CW;                                   // chunks per warp
    static constexpr bool RES = NT <= 64;  // residual add is only built for narrow tiles (register budget)
    uint4 res[RES ? HALFC / 8 : 1];
    uint32_t acc[CW / 16][16];
    bool has_res;

    __device__ __forceinline__ void prefetch(const EpiArgs &e, int half, int pss, int n, int y, int x, bool valid) {
        has_res = false;
        if (!RES || e.out_mode != OUT_C8 || e.residual == nullptr || !valid) return;
        has_res = true;
        if constexpr (RES) {
            const int Co8 = (e.Cout + 7) / 8;
            const uint4 *r = reinterpret_cast<const uint4 *>(e.residual + (long long)(n / e.res_div) * e.res_image_stride) + (long long)y * e.W + x;
            const long long plane = (long long)e.H * e.W;
#pragma unroll
            for (int j = 0; j < HALFC / 8; ++j) {
                const int c0 = half * HALFC + j * 8, q = (pss * NT + c0) / 8;
                res[j] = make_uint4(0, 0, 0, 0);
                if (c0 < NT && q < Co8) res[j] = __ldg(r + q * plane);
            }
        }
    }
    // TMEM -> registers for chunk `ch` of this warp's columns
    __device__ __forceinline__ void load(uint32_t taddr, int half, int ch) {
#pragma unroll
        for (int g = 0; g < CW / 16; ++g)
            if (half * HALFC + ch * CW + g * 16 < NT) tmem_ld16_nowait(taddr + half * HALFC + ch * CW + g * 16, acc[g]);
        tmem_ld_wait();
    }

    template <int ACT>
    __device__ __forceinline__ void store_c8(const EpiArgs &e, int half, int ch, int pss, int n, int y, int x) {
        int Ho = e.H, Wo = e.W;
        if (e.subsample) {
            if ((y | x) & 1) return;
            y >>= 1; x >>= 1; Ho = (e.H - 1) / 2 + 1; Wo = (e.W - 1) / 2 + 1;
        }
        const int Co8 = (e.Cout + 7) / 8;
        const long long plane = (long long)Ho * Wo;
        uint4 *o = reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(e.out) + (long long)n * e.out_image_stride) +
                   (long long)y * Wo + x;
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) {
            const int c0 = half * HALFC + ch * CW + j * 8, q = (pss * NT + c0) / 8;
            if (c0 >= NT || q >= Co8) continue;
            const float4 b0 = *reinterpret_cast<const float4 *>(e.bias_s + c0);
            const float4 b1 = *reinterpret_cast<const float4 *>(e.bias_s + c0 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float v[8], rr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (RES && has_res) {
                const uint32_t *h = reinterpret_cast<const uint32_t *>(&res[RES ? ch * (CW / 8) + j : 0]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = unpack16x2<BF>(h[i]);
                    rr[2 * i] = f.x; rr[2 * i + 1] = f.y;
                }
            }
            const bool pre = e.res_pre == 1, mask = BF && e.res_pre == 2 && RES && has_res;  // the mask mode exists on the training path only
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a = __uint_as_float(acc[j >> 1][(j & 1) * 8 + i]) + bb[i];
                if (mask) { const float t = act_t<ACT>(a, e.act); v[i] = rr[i] > 0.f ? t : e.res_slope * t; }
                else v[i] = pre ? act_t<ACT>(a + rr[i], e.act) : act_t<ACT>(a, e.act) + rr[i];
            }
            uint4 pk;
            uint32_t *h = reinterpret_cast<uint32_t *>(&pk);
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = pack16x2<BF>(v[2 * i], v[2 * i + 1]);
            o[q * plane] = pk;
        }
    }

    // nn.PixelShuffle(2) fused into the store: out[n, c, 2y + i, 2x + j] = conv[n, 4c + 2i + j, y, x].  Columns were
    // permuted at pack time (shuffle_col_to_channel) so that 16 consecutive columns are the j = 0 and j = 1 sub-pixels
    // of one 8-channel block: they are adjacent in memory and go out as ONE 256-bit store (full sectors).
    template <int ACT>
    __device__ __forceinline__ void store_shuffle(const EpiArgs &e, int half, int ch, int pss, int n, int y, int x) {
        constexpr int CP = NT / 4;  // channels per sub-pixel in one N-pass
        const int C2 = e.Cout / 4;
        const long long plane2 = (long long)(2 * e.H) * (2 * e.W);  // (pixel, block) cells per output channel block
        const int c0 = half * HALFC + ch * CW;
        if (c0 >= NT) return;
        const int i = c0 / (NT / 2), kb0 = (c0 % (NT / 2)) / 16;
        uint4 *ob = reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(e.out) + (long long)n * e.out_image_stride) +
                    (long long)(pss * (CP / 8) + kb0) * plane2 + (long long)(2 * y + i) * (2 * e.W) + 2 * x;
#pragma unroll
        for (int g = 0; g < CW / 16; ++g) {
            if ((pss * (CP / 8) + kb0 + g) * 8 >= C2) break;
            const float *b = e.bias_s + c0 + g * 16;
            uint4 pk[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t *h = reinterpret_cast<uint32_t *>(&pk[j]);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    h[k] = pack16x2<BF>(act_t<ACT>(__uint_as_float(acc[g][j * 8 + 2 * k]) + b[j * 8 + 2 * k], e.act),
                                        act_t<ACT>(__uint_as_float(acc[g][j * 8 + 2 * k + 1]) + b[j * 8 + 2 * k + 1], e.act));
            }
            st_global_256(ob + (long long)g * plane2, pk[0], pk[1]);
        }
    }

    // planar fp32 [n][Cout][H][W] with sigmoid from channel sig_from on (CUDA-core DCN's input format)
    __device__ __forceinline__ void store_planar(const EpiArgs &e, int half, int ch, int pss, int n, int y, int x) {
        const long long plane = (long long)e.H * e.W;
#pragma unroll
        for (int g = 0; g < CW / 16; ++g) {
            const int c0 = half * HALFC + ch * CW + g * 16;
            if (c0 >= NT) break;
            const int co0 = pss * NT + c0;
            float *o = reinterpret_cast<float *>(e.out) + (long long)n * e.out_image_stride + (long long)co0 * plane +
                       (long long)y * e.W + x;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float v = apply_act(__uint_as_float(acc[g][i]) + e.bias_s[c0 + i], e.act);
                if (co0 + i >= e.sig_from) v = __fdividef(1.f, 1.f + __expf(-v));
                if (co0 + i < e.Cout) o[i * plane] = v;
            }
        }
    }

    // OUT_OM24 (NT = 128): one register chunk = one deformable group = 32 columns
    // [dy0 dx0 .. dy8 dx8 m0 .. m8 pad*5] (column order fixed at weight-pack time).
    __device__ __forceinline__ void store_om24(const EpiArgs &e, int half, int ch, int pss, int n, int y, int x) {
        if constexpr (NT == 128 && CW == 32) {
            const int g = pss * 4 + (half * HALFC + ch * CW) / 32;
            if (g >= e.dg) return;
            const long long plane = (long long)e.H * e.W;
            const float *b = e.bias_s + half * HALFC + ch * CW;
            float v[27];
#pragma unroll
            for (int j = 0; j < 27; ++j) v[j] = __uint_as_float(acc[j >> 4][j & 15]) + b[j];
#pragma unroll
            for (int j = 18; j < 27; ++j) v[j] = __fdividef(1.f, 1.f + __expf(-v[j]));
            uint4 *og = reinterpret_cast<uint4 *>(e.out) + ((long long)n * e.out_image_stride) / 4 +
                        ((long long)(g * 3) * plane + (long long)y * e.W + x) * 2;
            st_global_256(og, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])),
                          make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
            og += plane * 2;
            st_global_256(og, make_uint4(__float_as_uint(v[8]), __float_as_uint(v[9]), __float_as_uint(v[10]), __float_as_uint(v[11])),
                          make_uint4(__float_as_uint(v[12]), __float_as_uint(v[13]), __float_as_uint(v[14]), __float_as_uint(v[15])));
            og += plane * 2;
            uint32_t m[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const __half2 h = __floats2half2_rn(v[18 + 2 * k], k < 4 ? v[19 + 2 * k] : 0.f);
                m[k] = *reinterpret_cast<const uint32_t *>(&h);
            }
            st_global_256(og, make_uint4(__float_as_uint(v[16]), __float_as_uint(v[17]), m[0], m[1]), make_uint4(m[2], m[3], m[4], 0u));
        }
    }

    // OUT_FINAL: out[n, c, y, x] = conv_last + bias + bilinear(center LQ frame) -- the reference's last two lines
    // (EDVR_arch.py:315-319) without materialising conv_last's output; base arithmetic as in final_add_kernel.
    __device__ __forceinline__ void store_final(const EpiArgs &e, int half, int ch, int n, int y, int x) {
        if (half != 0 || ch != 0) return;
        const FinalAdd &f = *e.fin;
        const int Hl = e.H / f.scale, Wl = e.W / f.scale;
        const long long cimg = f.center_map != nullptr ? (long long)__ldg(f.center_map + n) : (long long)n * f.frames + f.center;
        const long long lplane = (long long)Hl * Wl, hplane = (long long)e.H * e.W;
        int y0 = y, x0 = x, y1 = y, x1 = x;
        float ly = 0.f, lx = 0.f;
        if (f.scale != 1) {
            const float inv = 1.f / f.scale;
            const float sy = fmaxf(inv * (y + 0.5f) - 0.5f, 0.f), sx = fmaxf(inv * (x + 0.5f) - 0.5f, 0.f);
            y0 = (int)sy; x0 = (int)sx;
            y1 = y0 + (y0 < Hl - 1 ? 1 : 0); x1 = x0 + (x0 < Wl - 1 ? 1 : 0);
            ly = sy - y0; lx = sx - x0;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c >= f.nc) break;
            const long long pc = (cimg * f.nc + c) * lplane;
            float p00, p01, p10, p11;
            if (f.x_dtype == RVSR_F32) {
                const float *xp = reinterpret_cast<const float *>(f.x) + pc;
                p00 = __ldg(xp + y0 * Wl + x0); p01 = __ldg(xp + y0 * Wl + x1); p10 = __ldg(xp + y1 * Wl + x0); p11 = __ldg(xp + y1 * Wl + x1);
            } else {
                const __half *xp = reinterpret_cast<const __half *>(f.x) + pc;
                p00 = __half2float(__ldg(xp + y0 * Wl + x0)); p01 = __half2float(__ldg(xp + y0 * Wl + x1));
                p10 = __half2float(__ldg(xp + y1 * Wl + x0)); p11 = __half2float(__ldg(xp + y1 * Wl + x1));
            }
            const float base = f.scale == 1 ? p00 : (1.f - ly) * ((1.f - lx) * p00 + lx * p01) + ly * ((1.f - lx) * p10 + lx * p11);
            const float v = apply_act(__uint_as_float(acc[0][c]) + e.bias_s[c], e.act) + base;
            const long long oi = ((long long)n * f.nc + c) * hplane + (long long)y * e.W + x;
            if (f.out_dtype == RVSR_F32) reinterpret_cast<float *>(e.out)[oi] = v;
            else reinterpret_cast<__half *>(e.out)[oi] = __float2half_rn(v);
        }
    }

    __device__ __forceinline__ void store(const EpiArgs &e, int half, int ch, int pss, int n, int y, int x, bool valid) {
        if (!valid) return;
        if (e.out_mode == OUT_FINAL) { store_final(e, half, ch, n, y, x); return; }
        if (e.out_mode == OUT_C8) {
            if (e.act == RVSR_ACT_LRELU) store_c8<RVSR_ACT_LRELU>(e, half, ch, pss, n, y, x);
            else if (e.act == RVSR_ACT_RELU) store_c8<RVSR_ACT_RELU>(e, half, ch, pss, n, y, x);
            else store_c8<RVSR_ACT_NONE>(e, half, ch, pss, n, y, x);
        } else if (e.out_mode == OUT_OM24) {
            store_om24(e, half, ch, pss, n, y, x);
        } else if (e.out_mode == OUT_C8_SHUFFLE2) {
            if (e.act == RVSR_ACT_LRELU) store_shuffle<RVSR_ACT_LRELU>(e, half, ch, pss, n, y, x);
            else store_shuffle<-1>(e, half, ch, pss, n, y, x);
        } else {
            store_planar(e, half, ch, pss, n, y, x);
        }
    }

    // Drain this warp's part of one accumulator: chunk by chunk (32 columns in registers at a time); `release`
    // hands the TMEM buffer back to the MMA issuer right after the last chunk has been loaded.
    template <typename Release>
    __device__ __forceinline__ void run(const EpiArgs &e, uint32_t taddr, int half, int pss, int n, int y, int x, bool valid,
                                        bool do_load, bool do_store, Release &&release) {
        bool released = false;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
            if (half * HALFC + ch * CW >= NT) break;
            if (do_load) load(taddr, half, ch);
            if (ch == NCH - 1 || half * HALFC + (ch + 1) * CW >= NT) { release(); released = true; }
            if (do_store) store(e, half, ch, pss, n, y, x, valid);
        }
        if (!released) release();  // a warp without columns (NT < NH * 16) still owes its arrival
    }
};

// Lean epilogue for the dominant case: 64-wide tile, OUT_C8, stride 1 (every 64 -> 64 / 128 -> 64 convolution of the
// network).  The epilogue's instruction stream is ~1/6 of the step's energy (the step is power-capped), so: activation
// as a template parameter, one 64-bit output pointer advanced by a constant plane stride, residual blocks fetched
// before the accumulator wait, all 32 columns of the warp in registers with ONE tcgen05.wait::ld, and the TMEM
// buffer handed back before any arithmetic.  Exact integer division of the tile index by precomputed magic
// numbers (host: magic_div) replaces three hardware divisions per tile and warp.
template <int ACT, bool BF, typename Release>
__device__ __forceinline__ void epi_c8_fast(const float *bias_s, __half *out, long long out_image_stride, const __half *residual,
                                            long long res_image_stride, int H, int W, int Co8, uint32_t taddr, int half,
                                            int qbase, int n, int y, int x, bool valid, uint32_t full_bar, uint32_t full_par,
                                            Release &&release, int n_res = -1, int res_mode = 0, float res_slope = 0.f) {
    constexpr int NBLK = 4;  // this warp's 32 columns = 4 channel blocks
    const long long plane = (long long)H * W, pix = (long long)y * W + x;
    const int q0 = qbase + half * NBLK;  // first output channel block of this warp (qbase: blocks of earlier N-passes)
    uint4 res[NBLK];
    const bool has_res = residual != nullptr && valid;
    if (has_res) {
        const uint4 *r = reinterpret_cast<const uint4 *>(residual + (long long)(n_res < 0 ? n : n_res) * res_image_stride) + q0 * plane + pix;
#pragma unroll
        for (int j = 0; j < NBLK; ++j) res[j] = q0 + j < Co8 ? __ldg(r + j * plane) : make_uint4(0, 0, 0, 0);
    }
    mbar_wait(full_bar, full_par);
    tc_fence_after();
    uint32_t a0[16], a1[16];
    tmem_ld16_nowait(taddr + half * 32, a0);
    tmem_ld16_nowait(taddr + half * 32 + 16, a1);
    tmem_ld_wait();
    release();
    if (!valid) return;
    uint4 *o = reinterpret_cast<uint4 *>(out + (long long)n * out_image_stride) + q0 * plane + pix;
    const float *b = bias_s + half * 32;
#pragma unroll
    for (int j = 0; j < NBLK; ++j) {
        if (q0 + j >= Co8) break;
        const float4 b0 = *reinterpret_cast<const float4 *>(b + j * 8), b1 = *reinterpret_cast<const float4 *>(b + j * 8 + 4);
        const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
        const uint32_t *a = (j < 2 ? a0 : a1) + (j & 1) * 8;
        float2 v[4];
        const bool pre = has_res && res_mode == 1;  // split-cat convolution: the other half's partial sums join before bias + activation
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // packed fp32 pairs (FADD2 / FMUL2): same IEEE results, half the instructions
            v[i] = add2(make_float2(__uint_as_float(a[2 * i]), __uint_as_float(a[2 * i + 1])), bb[i]);
            if (pre) v[i] = add2(v[i], unpack16x2<BF>(reinterpret_cast<const uint32_t *>(&res[j])[i]));
            if (ACT == RVSR_ACT_LRELU) {
                const float2 t = mul2(v[i], make_float2(0.1f, 0.1f));
                v[i] = make_float2(fmaxf(v[i].x, t.x), fmaxf(v[i].y, t.y));
            } else if (ACT == RVSR_ACT_RELU) {
                v[i] = make_float2(fmaxf(v[i].x, 0.f), fmaxf(v[i].y, 0.f));
            }
        }
        if (has_res && !pre) {
            const uint32_t *h = reinterpret_cast<const uint32_t *>(&res[j]);
            if (BF && res_mode == 2) {  // mask: gradient through the activation that produced `residual` (training path only)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 m = unpack16x2<BF>(h[i]);
                    v[i] = make_float2(m.x > 0.f ? v[i].x : res_slope * v[i].x, m.y > 0.f ? v[i].y : res_slope * v[i].y);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = add2(v[i], unpack16x2<BF>(h[i]));
            }
        }
        uint4 pk;
        uint32_t *h = reinterpret_cast<uint32_t *>(&pk);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = pack16x2<BF>(v[i].x, v[i].y);
        *o = pk;
        o += plane;
    }
}

// ---------------------------------------------------------------- convolution
struct alignas(64) TcConvParams {
    CUtensorMap tmap[RVSR_MAX_SRC_TC];
    int src_pstride[RVSR_MAX_SRC_TC];  // channel-block planes between consecutive images of a source
    int src_frames[RVSR_MAX_SRC_TC], src_fixed[RVSR_MAX_SRC_TC];
    const int *src_map[RVSR_MAX_SRC_TC];  // optional image -> slot tables (feature cache)
    int nsrc, C8s, nstages;
    const __half *w;  // [pass][tap][Q][NT][8]
    const float *bias;
    void *out;
    long long out_image_stride;
    const __half *residual;
    long long res_image_stride;
    int res_pre, res_div;
    float res_slope;
    int N, H, W, Cout, act, out_mode, sig_from, subsample, dg;
    int tiles_x, tiles_y, num_tiles;
    TileDiv td;
    FinalAdd fin;
    int bf16;   // operands, residual and output are bfloat16 instead of fp16 (training path; same kernels, other instruction-descriptor formats)
    int stamp;  // slot in g_stamps or -1
    int debug;  // RVSR_TC_DEBUG bit mask for timing experiments only (results become garbage):
                // 1 = issue no MMAs, 2 = no epilogue stores, 4 = no TMA halo loads, 8 = no TMEM loads
};

constexpr int TC_EPI_WARPS = 16, TC_EPI_GROUPS = 2;  // two groups of 8 epilogue warps take alternate tiles
// warp 0: TMA producer; warps 1-2: MMA issuers (alternate tiles -- one thread cannot issue tcgen05.mma
// fast enough to keep the tensor pipe busy at N = 64); warp 3: TMEM allocator; warps 4-11: epilogue
constexpr int TC_EPI_WARP0 = 4;
constexpr int TC_THREADS = 32 * (TC_EPI_WARP0 + TC_EPI_WARPS);
__host__ __device__ constexpr int acc_stride(int NT) { return NT <= 32 ? 32 : (NT <= 64 ? 64 : 128); }

template <int KS, int NT, bool BF = false>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcConvParams p) {
    constexpr int KK = KS * KS, PAD = KS / 2, VALID = TC_TW - (KS - 1), HALO_ROWS = TC_ROWS + KS - 1;
    constexpr int PLANE_BYTES = HALO_ROWS * TC_TW * 16;
    constexpr int ACC = acc_stride(NT);
    // Issuer warps (1..3): one thread cannot issue tcgen05.mma fast enough, and while an issuer handles its
    // per-tile barriers / commits (~1.5k cycles) the others must keep the pipe full.  Tiles are dealt
    // round-robin.  mbarrier waits are parity based, so every barrier must have ONE consumer that sees
    // each phase in order: each issuer has its own "full" barrier per stage (the producer arms the one
    // of the tile's issuer), and the number of TMEM accumulators is a multiple of the issuer count so an
    // accumulator is always filled by the same issuer.
    // The epilogue is latency bound per tile (~1.3k cycles), so two groups of epilogue warps take alternate
    // tiles; the accumulator count is even so a buffer always belongs to the same group.
    // Narrow tiles: 3 issuers, 6 accumulators, 2 epilogue groups.  Wide tiles (128 columns): 3 issuers,
    // 3 accumulators (384 of 512 TMEM columns), ONE epilogue group of 16 warps (32 columns per warp).
    constexpr int MMAW = 3;
    constexpr int EG = ACC <= 64 ? TC_EPI_GROUPS : 1;
    constexpr int NB = ACC <= 64 ? 6 : 3;
    static_assert(NB % EG == 0 && NB % MMAW == 0, "accumulator -> issuer / epilogue group mapping must be fixed");
    constexpr int TMEM_COLS = NB * ACC <= 32 ? 32 : NB * ACC <= 64 ? 64 : NB * ACC <= 128 ? 128 : NB * ACC <= 256 ? 256 : 512;
    static_assert(NB * ACC <= 512 && NB % MMAW == 0, "TMEM accumulator plan");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int Q = p.nsrc * p.C8s;  // channel blocks over all sources
    const uint32_t w_bytes = (uint32_t)Q * KK * NT * 16;
    const uint32_t stage_bytes = (uint32_t)p.C8s * PLANE_BYTES;
    uint8_t *w_s = smem;
    uint8_t *stage_s = smem + w_bytes;
    float *bias_s = reinterpret_cast<float *>(stage_s + (size_t)p.nstages * stage_bytes + 128);  // +128: tap overrun slack
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + NT);
    // bars: MMAW x S full (per issuer), S empty, weights-full, NB accumulator-full, NB accumulator-empty, TMEM base
    const int S = p.nstages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + (MMAW + 1) * S + 1 + 2 * NB);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    auto FULL = [&](int w, int st) { return BAR(w * S + st); };
    auto EMPTY = [&](int st) { return BAR(MMAW * S + st); };
    const uint32_t WFULL = BAR((MMAW + 1) * S);
    auto TFULL = [&](int b) { return BAR((MMAW + 1) * S + 1 + b); };
    auto TEMPTY = [&](int b) { return BAR((MMAW + 1) * S + 1 + NB + b); };
    const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int pss = blockIdx.y;
    // RVSR_TC_DEBUG & 16: per-role phase timing (clock64 sums over all tiles of CTA 0), printed at exit
    __shared__ long long ph_acc[3][8];
    const bool stamp = (p.debug & 16) && blockIdx.x == 0 && blockIdx.y == 0;
    if (threadIdx.x < 24) ph_acc[threadIdx.x / 8][threadIdx.x % 8] = 0;
    long long tprev = 0;
#define PH_BEGIN() do { if (stamp && lane == 0) tprev = clock64(); } while (0)
#define PH(role, k) do { if (stamp && lane == 0) { const long long tn = clock64(); ph_acc[role][k] += tn - tprev; tprev = tn; } } while (0)

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int i = 0; i < (MMAW + 1) * S + 1 + NB; ++i) mbar_init(BAR(i), 1);
        for (int i = 0; i < NB; ++i) mbar_init(TEMPTY(i), TC_EPI_WARPS / EG);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < NT; i += TC_THREADS) {
        const int co = pss * NT + i;
        float b = 0.f;
        if (p.bias != nullptr) {
            if (p.out_mode == OUT_C8_SHUFFLE2) {
                const int cc = shuffle_col_to_channel(i, pss, NT);
                b = cc < p.Cout ? p.bias[cc] : 0.f;
            } else if (p.out_mode == OUT_OM24) {
                const int g = pss * 4 + i / 32, j = i % 32;
                if (g < p.dg && j < 27) b = p.bias[j < 18 ? g * 18 + j : 18 * p.dg + g * 9 + (j - 18)];
            } else if (co < p.Cout) {
                b = p.bias[co];
            }
        }
        bias_s[i] = b;
    }
    if (warp == 3) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);

    if (warp == 0) {
        if (lane == 0) {
            // ---- weights for this N-pass: resident for the CTA's lifetime
            mbar_expect_tx(WFULL, w_bytes);
            const uint8_t *wg = reinterpret_cast<const uint8_t *>(p.w) + (size_t)pss * w_bytes;
            for (uint32_t o = 0; o < w_bytes; o += 32768) {
                const uint32_t n = w_bytes - o < 32768 ? w_bytes - o : 32768;
                bulk_load(smem_u32(w_s + o), wg + o, n, WFULL);
            }
            for (int s = 0; s < p.nsrc; ++s) prefetch_tensormap(&p.tmap[s]);
            pdl_wait();  // weights are static; everything below reads the previous kernel's output
            // ---- halo tiles
            uint32_t it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
                const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, n = tile / (p.tiles_x * p.tiles_y);
                const int w = (int)(tl % MMAW);  // issuer that owns this tile
                for (int s = 0; s < p.nsrc; ++s, ++it) {
                    const int st = it % S;
                    PH_BEGIN();
                    mbar_wait(EMPTY(st), ((it / S) & 1) ^ 1);
                    PH(0, 0);
                    if (p.debug & 4) { mbar_arrive(FULL(w, st)); PH(0, 1); continue; }
                    mbar_expect_tx(FULL(w, st), stage_bytes);
                    const int img = p.src_map[s] != nullptr ? __ldg(p.src_map[s] + n)
                                    : (p.src_fixed[s] >= 0 ? (n / p.src_frames[s]) * p.src_frames[s] + p.src_fixed[s] : n);
                    tma_load_3d(smem_u32(stage_s + (size_t)st * stage_bytes), &p.tmap[s], FULL(w, st),
                                (tx * VALID - PAD) * 8, ty * TC_ROWS - PAD, img * p.src_pstride[s]);
                    PH(0, 1);
                }
            }
        }
    } else if (warp <= MMAW) {
        {
            // ---- MMA issuer(s).  The issue stream is the critical path of the whole kernel (the tensor pipe
            // needs a new N=64 MMA every 48 cycles), so: whole warp on uniform values (see elect_one()), no
            // divisions, no 64-bit descriptor rebuilds, running counters instead of modulo, everything unrolled.
            constexpr uint32_t idesc = make_idesc(NT) | (BF ? (1u << 7) | (1u << 10) : 0u);  // a/b format: 0 = F16, 1 = BF16
            const uint32_t mw = (uint32_t)(warp - 1);
            const uint32_t nsrc = (uint32_t)p.nsrc, C8s = (uint32_t)p.C8s;
            const uint64_t adesc0 = make_desc(smem_u32(stage_s), PLANE_BYTES, 128);
            const uint64_t bdesc0 = make_desc(smem_u32(w_s), NT * 16, 128);
            const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
            const uint32_t a_base = (uint32_t)adesc0, b_base = (uint32_t)bdesc0;
            const uint32_t stage_units = stage_bytes >> 4;                 // descriptor address units are 16 B
            const uint32_t b_src_step = C8s * (NT * 16 / 16), b_tap_step = (uint32_t)Q * (NT * 16 / 16);
            const int nk = (p.debug & 1) ? 0 : (int)C8s / 2;
            // stage ring position of this issuer's first tile, then advanced by MMAW tiles at a time;
            // `par` holds, per stage, the parity of the next phase of THIS issuer's full barrier
            uint32_t st = (mw * nsrc) % (uint32_t)S, par = 0;
            mbar_wait(WFULL, 0);
            uint32_t t = mw;
            for (int tile = blockIdx.x + (int)mw * gridDim.x; tile < p.num_tiles; tile += MMAW * gridDim.x, t += MMAW) {
                const uint32_t buf = t % NB;
                if (mw == 0) PH_BEGIN();
                mbar_wait(TEMPTY(buf), ((t / NB) & 1) ^ 1);  // epilogue has drained this accumulator
                if (mw == 0) PH(1, 0);
                tc_fence_after();
                if (mw == 0) PH(1, 1);
                const uint32_t d = tmem_base + buf * ACC;
                uint32_t b_lo0 = b_base;
                for (uint32_t s = 0; s < nsrc; ++s) {
                    mbar_wait(FULL(mw, st), (par >> st) & 1);
                    par ^= 1u << st;
                    if (mw == 0) PH(1, 2);
                    tc_fence_after();
                    if (mw == 0) PH(1, 3);
                    const uint32_t a_lo0 = a_base + st * stage_units;
                    if (elect_one()) {
#pragma unroll
                        for (int tap = 0; tap < KK; ++tap) {
                            const uint32_t a_lo = a_lo0 + (uint32_t)((tap / KS) * TC_TW + (tap % KS));
                            const uint32_t b_lo = b_lo0 + (uint32_t)tap * b_tap_step;
                            if (nk == 4) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma_f16(d, ((uint64_t)a_hi << 32) | (a_lo + (uint32_t)kk * (2 * PLANE_BYTES / 16)),
                                             ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)kk * (2 * NT)), idesc,
                                             (tap | kk) ? 1u : (s ? 1u : 0u));
                            } else {
                                for (int kk = 0; kk < nk; ++kk)
                                    umma_f16(d, ((uint64_t)a_hi << 32) | (a_lo + (uint32_t)kk * (2 * PLANE_BYTES / 16)),
                                             ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)kk * (2 * NT)), idesc,
                                             (tap | kk) ? 1u : (s ? 1u : 0u));
                            }
                        }
                        umma_commit(EMPTY(st));  // stage reusable once these MMAs have read it
                        if (s + 1 == nsrc) umma_commit(TFULL(buf));  // accumulator complete
                    }
                    __syncwarp();
                    if (mw == 0) PH(1, 4);
                    b_lo0 += b_src_step;
                    if (++st == (uint32_t)S) st = 0;
                }
                if (MMAW > 1) {  // skip the stages of the tiles the other issuer(s) own
                    st += (MMAW - 1) * nsrc;
                    while (st >= (uint32_t)S) st -= (uint32_t)S;
                }
            }
        }
    } else if (warp >= TC_EPI_WARP0) {
        pdl_wait();  // residual reads / output writes (the issuers only consume what the producer loaded after its wait)
        const int lq = warp & 3;                         // TMEM lane quarter this warp may access == tile row
        constexpr int WPG = TC_EPI_WARPS / EG;                       // warps per epilogue group
        const int eg = (warp - TC_EPI_WARP0) / WPG;                  // this warp's group: tiles t == eg (mod groups)
        const int half = ((warp - TC_EPI_WARP0) % WPG) >> 2;         // WPG / 4 warps per lane quarter split the columns
        EpiArgs e{bias_s, p.out, p.out_image_stride, p.residual, p.res_image_stride, p.H, p.W, p.Cout, p.act,
                  p.out_mode, p.sig_from, p.subsample, p.dg, &p.fin, p.res_pre, p.res_div, p.bf16, p.res_slope};
        EpiTile<NT, WPG / 4, BF> ep;
        uint32_t t = (uint32_t)eg;
        for (int tile = blockIdx.x + eg * gridDim.x; tile < p.num_tiles; tile += EG * gridDim.x, t += EG) {
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, n = tile / (p.tiles_x * p.tiles_y);
            const uint32_t buf = t % NB;
            const int y = ty * TC_ROWS + lq, x = tx * VALID + lane;
            const bool valid = lane < VALID && y < p.H && x < p.W;
            const bool es = warp == TC_EPI_WARP0 && lane == 0;
            if (es) PH_BEGIN();
            ep.prefetch(e, half, pss, n, y, x, valid);
            if (es) PH(2, 0);
            mbar_wait(TFULL(buf), (t / NB) & 1);
            if (es) PH(2, 1);
            tc_fence_after();
            if (es) PH(2, 2);
            ep.run(e, tmem_base + buf * ACC + ((uint32_t)(lq * 32) << 16), half, pss, n, y, x, valid, !(p.debug & 8), !(p.debug & 2), [&] {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(TEMPTY(buf));  // accumulator is in registers: the issuer may reuse the buffer
            });
            if (es) PH(2, 4);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc(tmem_base, TMEM_COLS);
    if (stamp && threadIdx.x == 0 && p.num_tiles > 1500) {
        const int nt = (p.num_tiles + gridDim.x - 1) / gridDim.x;
        printf("[tc conv KS=%d NT=%d nsrc=%d] tiles/CTA %d | per tile: producer wait-empty %lld issue %lld | issuer0 (per own tile) wait-tempty %lld fence %lld "
               "wait-full %lld fence %lld mma-issue %lld commit-stage %lld commit-acc %lld | epilogue prefetch %lld wait-tfull %lld tmem-ld %lld arrive %lld store %lld\n",
               KS, NT, p.nsrc, nt, ph_acc[0][0] / nt, ph_acc[0][1] / nt, ph_acc[1][0] * MMAW / nt, ph_acc[1][1] * MMAW / nt, ph_acc[1][2] * MMAW / nt,
               ph_acc[1][3] * MMAW / nt, ph_acc[1][4] * MMAW / nt, ph_acc[1][5] * MMAW / nt, ph_acc[1][6] * MMAW / nt, ph_acc[2][0] / nt,
               ph_acc[2][1] / nt, ph_acc[2][2] / nt, ph_acc[2][3] / nt, ph_acc[2][4] / nt);
    }
#undef PH
#undef PH_BEGIN
}

// ---------------------------------------------------------------- convolution, CTA pair (cta_group::2)
// Same algorithm as conv_tc_kernel<3, 64> on a cluster of two CTAs (two SMs of one TPC): one
// tcgen05.mma.cta_group::2 covers M = 256 pixels (each CTA's own 128-pixel tile, its own halo in its own
// shared memory) x N = 64 output channels, and the B operand (weights) is split between the pair -- each
// CTA keeps and feeds only 32 of the 64 rows.  Per SM and MMA that is 4 KB of A + 1 KB of B from shared
// memory instead of 4 + 2 KB: 40 instead of 48 cycles, i.e. the N = 64 shape's ceiling moves from 2/3 to 0.8
// of the tensor peak.  Only the leader CTA (cluster rank 0) issues MMAs.  Synchronisation:
//   FULL(issuer, stage)  leader's barrier, 2 arrivals (leader: arrive.expect_tx for both tiles' bytes, peer:
//                        remote arrive) + both CTAs' TMA (cp.async.bulk.tensor ... cta_group::2) complete_tx
//   EMPTY(stage), TFULL(acc)  tcgen05.commit ... multicast::cluster -> the same barrier in BOTH CTAs
//   TEMPTY(acc)          leader's barrier, arrivals from the epilogue warps of both CTAs (peer: remote arrive)
// S2 = true: stride-2 convolution as a REAL implicit GEMM (fea_L2/L3_conv1, the HR_in stem, the predeblur pyramid:
// EDVR_arch.py:279,:282,:229-231,:31-32).  out(y, x) = sum in(2y + dy - 1, 2x + dx - 1): split the input into its four
// (row parity p, column parity q) phase images; tap dy reads phase p = 1, 0, 1 at phase row y - 1, y, y, and the same for dx.
// The four phase tiles (5 rows x 32 pixels each) come straight from the un-split tensor through a 4-D tensor map with
// element strides {1, 2, 2, 1}; each tap is then a view of one phase tile shifted by (0 | 1, 0 | 1), exactly like the
// stride-1 taps.  Kernel geometry (H, W, tiles) is that of the OUTPUT.  (Round 1 computed these layers at full resolution
// and stored every other pixel: 4x the MMAs.)
template <int KS, int NT, bool S2 = false, bool BF = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) conv_tc2_kernel(const __grid_constant__ TcConvParams p) {
    constexpr int NH2 = NT / 2;
    constexpr int KK = KS * KS, PAD = KS / 2, VALID = TC_TW - (KS - 1), HALO_ROWS = S2 ? TC_ROWS + 1 : TC_ROWS + KS - 1;
    constexpr int PLANE_BYTES = HALO_ROWS * TC_TW * 16;
    static_assert(!S2 || (KS == 3 && NT == 64), "stride-2 path: 3x3, 64-wide tiles");
    // one MMA instruction covers two tiles here, so two issuers suffice for the wide tiles, which leaves room for
    // 4 accumulators (all 512 TMEM columns) and two epilogue groups (the OM24 / pixel-shuffle epilogues are the
    // slower side of those kernels)
    constexpr int ACC = acc_stride(NT), MMAW = ACC <= 64 ? 3 : 2, EG = TC_EPI_GROUPS, NB = ACC <= 64 ? 6 : 4, TMEM_COLS = 512;
    static_assert(NT == 64 || NT == 128, "pair kernel is built for 64- and 128-wide tiles");
    static_assert(NB % MMAW == 0 && NB % EG == 0 && NB * ACC <= 512, "accumulator plan");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int Q = p.nsrc * p.C8s;
    const uint32_t w_bytes = (uint32_t)Q * KK * NH2 * 16;  // this CTA's half of the weights
    const uint32_t phase_bytes = (uint32_t)p.C8s * PLANE_BYTES;                // S2: one phase tile (all channel blocks)
    const uint32_t stage_bytes = S2 ? 4 * phase_bytes : phase_bytes;
    uint8_t *w_s = smem;
    uint8_t *stage_s = smem + w_bytes;
    float *bias_s = reinterpret_cast<float *>(stage_s + (size_t)p.nstages * stage_bytes + 128);
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + NT);
    const int S = p.nstages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + (MMAW + 1) * S + 2 + 2 * NB);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    auto FULL = [&](int w, int st) { return BAR(w * S + st); };
    auto EMPTY = [&](int st) { return BAR(MMAW * S + st); };
    const uint32_t WFULL = BAR((MMAW + 1) * S);
    auto TFULL = [&](int b) { return BAR((MMAW + 1) * S + 1 + b); };
    auto TEMPTY = [&](int b) { return BAR((MMAW + 1) * S + 1 + NB + b); };
    const uint32_t WPEER = BAR((MMAW + 1) * S + 1 + 2 * NB);  // leader only: the peer's weight half has landed
    const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const uint32_t rank = blockIdx.x & 1u;  // == %cluster_ctarank for __cluster_dims__(2, 1, 1); provably uniform
    const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int npairs = (p.num_tiles + 1) / 2;
    const int pss = blockIdx.y;

    const bool stamp = p.stamp >= 0 && blockIdx.x == 0 && blockIdx.y == 0;
#define STAMP(k) do { if (stamp) g_stamps[p.stamp][k] = globaltimer_ns(); } while (0)
    if (threadIdx.x == 0) STAMP(0);
    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int i = 0; i < MMAW * S; ++i) mbar_init(BAR(i), 2);                       // FULL: leader + peer arrive
        for (int i = MMAW * S; i < (MMAW + 1) * S + 1 + NB; ++i) mbar_init(BAR(i), 1);  // EMPTY, WFULL, TFULL
        for (int i = 0; i < NB; ++i) mbar_init(TEMPTY(i), 2 * (TC_EPI_WARPS / EG));   // epilogue warps of both CTAs
        mbar_init(WPEER, 1);
        fence_barrier_init();
        // this CTA's half of the weights: resident for the CTA's lifetime
        mbar_expect_tx(WFULL, w_bytes);
        const uint8_t *wg = reinterpret_cast<const uint8_t *>(p.w) + ((size_t)pss * 2 + rank) * w_bytes;
        for (uint32_t o = 0; o < w_bytes; o += 32768) {
            const uint32_t n = w_bytes - o < 32768 ? w_bytes - o : 32768;
            bulk_load(smem_u32(w_s + o), wg + o, n, WFULL);
        }
        for (int s = 0; s < p.nsrc; ++s) prefetch_tensormap(&p.tmap[s]);
    }
    // Kept OFF the prologue's critical path (every launch pays it): the bias is staged by the epilogue warps while
    // the first halo tiles are in flight, and the weights are awaited by the issuers -- the leader needs both halves:
    // the peer forwards "my half has landed" to the leader's WPEER barrier.
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    if (threadIdx.x == 0) STAMP(1);
    pdl_wait();  // prologue (barriers, weights, TMEM) overlapped the previous kernel's tail; activations from here on
    if (threadIdx.x == 0) STAMP(2);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0, tl = 0;
            for (int pr = cid; pr < npairs; pr += nclusters, ++tl) {
                int tile = 2 * pr + (int)rank;
                if (tile >= p.num_tiles) tile = p.num_tiles - 1;  // odd tile count: the peer recomputes the last tile, never stores it
                const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, n = tile / (p.tiles_x * p.tiles_y);
                const int w = (int)(tl % MMAW);
                for (int s = 0; s < p.nsrc; ++s, ++it) {
                    const int st = it % S;
                    mbar_wait(EMPTY(st), ((it / S) & 1) ^ 1);
                    const uint32_t full0 = mapa_rank0(FULL(w, st));
                    if (p.debug & 4) { mbar_arrive_cluster(full0); continue; }  // timing experiment: no halo loads
                    if (rank == 0)
                        mbar_expect_tx(FULL(w, st), 2 * stage_bytes);
                    else
                        mbar_arrive_cluster(full0);
                    const int img = p.src_map[s] != nullptr ? __ldg(p.src_map[s] + n)
                                    : (p.src_fixed[s] >= 0 ? (n / p.src_frames[s]) * p.src_frames[s] + p.src_fixed[s] : n);
                    if (S2) {
#pragma unroll
                        for (int ph = 0; ph < 4; ++ph)  // phase (row parity ph >> 1, column parity ph & 1), origin one phase pixel up / left
                            tma_load_4d_2sm(smem_u32(stage_s + (size_t)st * stage_bytes + ph * phase_bytes), &p.tmap[s], full0, 0,
                                            2 * (tx * VALID - 1) + (ph & 1), 2 * (ty * TC_ROWS - 1) + (ph >> 1), img * p.src_pstride[s]);
                    } else {
                        tma_load_3d_2sm(smem_u32(stage_s + (size_t)st * stage_bytes), &p.tmap[s], full0,
                                        (tx * VALID - PAD) * 8, ty * TC_ROWS - PAD, img * p.src_pstride[s]);
                    }
                }
            }
        }
    } else if (warp <= MMAW) {
        if (rank != 0) {
            if (warp == 1 && lane == 0) {
                mbar_wait(WFULL, 0);
                mbar_arrive_cluster(mapa_rank0(WPEER));
            }
        } else {  // whole warp, uniform values; only the tcgen05 instructions are predicated on one lane
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((256u >> 4) << 24) |  // M = 256 over the pair
                                       (BF ? (1u << 7) | (1u << 10) : 0u);
            const uint32_t mw = (uint32_t)(warp - 1);
            const uint32_t nsrc = (uint32_t)p.nsrc, C8s = (uint32_t)p.C8s;
            const uint64_t adesc0 = make_desc(smem_u32(stage_s), PLANE_BYTES, 128);
            const uint64_t bdesc0 = make_desc(smem_u32(w_s), NH2 * 16, 128);
            const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
            const uint32_t a_base = (uint32_t)adesc0, b_base = (uint32_t)bdesc0;
            const uint32_t stage_units = stage_bytes >> 4, phase_units = phase_bytes >> 4;
            const uint32_t b_src_step = C8s * (NH2 * 16 / 16), b_tap_step = (uint32_t)Q * (NH2 * 16 / 16);
            const int nk = (p.debug & 1) ? 0 : (int)C8s / 2;
            uint32_t st = (mw * nsrc) % (uint32_t)S, par = 0;
            uint32_t t = mw;
            mbar_wait(WFULL, 0);  // both weight halves resident (phase 0 of these two barriers completes exactly once)
            mbar_wait(WPEER, 0);
            for (int pr = cid + (int)mw * nclusters; pr < npairs; pr += MMAW * nclusters, t += MMAW) {
                const uint32_t buf = t % NB;
                mbar_wait(TEMPTY(buf), ((t / NB) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + buf * ACC;
                uint32_t b_lo0 = b_base;
                for (uint32_t s = 0; s < nsrc; ++s) {
                    mbar_wait(FULL(mw, st), (par >> st) & 1);
                    par ^= 1u << st;
                    tc_fence_after();
                    if (mw == 0 && t == 0 && s == 0 && lane == 0) STAMP(3);
                    const uint32_t a_lo0 = a_base + st * stage_units;
                    if (elect_one()) {
#pragma unroll
                        for (int tap = 0; tap < KK; ++tap) {
                            // stride 1: tap (dy, dx) = the halo tile advanced by dy rows and dx pixels.  stride 2: phase tile
                            // (dy != 1, dx != 1) advanced by (dy != 0) rows and (dx != 0) pixels (see the kernel's header)
                            const uint32_t a_lo = S2 ? a_lo0 + (uint32_t)(((tap / KS) != 1 ? 2 : 0) + ((tap % KS) != 1 ? 1 : 0)) * phase_units +
                                                           (uint32_t)(((tap / KS) != 0 ? TC_TW : 0) + ((tap % KS) != 0 ? 1 : 0))
                                                     : a_lo0 + (uint32_t)((tap / KS) * TC_TW + (tap % KS));
                            const uint32_t b_lo = b_lo0 + (uint32_t)tap * b_tap_step;
                            if (nk == 4) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma_f16_2sm(d, ((uint64_t)a_hi << 32) | (a_lo + (uint32_t)kk * (2 * PLANE_BYTES / 16)),
                                                 ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)kk * (2 * NH2)), idesc,
                                                 (tap | kk) ? 1u : (s ? 1u : 0u));
                            } else {
                                for (int kk = 0; kk < nk; ++kk)
                                    umma_f16_2sm(d, ((uint64_t)a_hi << 32) | (a_lo + (uint32_t)kk * (2 * PLANE_BYTES / 16)),
                                                 ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)kk * (2 * NH2)), idesc,
                                                 (tap | kk) ? 1u : (s ? 1u : 0u));
                            }
                        }
                        umma_commit_2sm(EMPTY(st));
                        if (s + 1 == nsrc) umma_commit_2sm(TFULL(buf));
                    }
                    __syncwarp();
                    b_lo0 += b_src_step;
                    if (++st == (uint32_t)S) st = 0;
                }
                st += (MMAW - 1) * nsrc;
                while (st >= (uint32_t)S) st -= (uint32_t)S;
            }
        }
    } else if (warp >= TC_EPI_WARP0) {
        for (int i = threadIdx.x - 32 * TC_EPI_WARP0; i < NT; i += 32 * TC_EPI_WARPS) {  // column -> channel maps as in conv_tc_kernel
            const int co = pss * NT + i;
            float b = 0.f;
            if (p.bias != nullptr) {
                if (p.out_mode == OUT_C8_SHUFFLE2) {
                    const int cc = shuffle_col_to_channel(i, pss, NT);
                    b = cc < p.Cout ? p.bias[cc] : 0.f;
                } else if (p.out_mode == OUT_OM24) {
                    const int g = pss * 4 + i / 32, j = i % 32;
                    if (g < p.dg && j < 27) b = p.bias[j < 18 ? g * 18 + j : 18 * p.dg + g * 9 + (j - 18)];
                } else if (co < p.Cout) {
                    b = p.bias[co];
                }
            }
            bias_s[i] = b;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");  // epilogue warps only
        constexpr int WPG = TC_EPI_WARPS / EG;
        const int lq = warp & 3;
        const int eg = (warp - TC_EPI_WARP0) / WPG;
        const int half = ((warp - TC_EPI_WARP0) % WPG) >> 2;
        EpiArgs e{bias_s, p.out, p.out_image_stride, p.residual, p.res_image_stride, p.H, p.W, p.Cout, p.act,
                  p.out_mode, p.sig_from, p.subsample, p.dg, &p.fin, p.res_pre, p.res_div, p.bf16, p.res_slope};
        EpiTile<NT, WPG / 4, BF> ep;
        uint32_t t = (uint32_t)eg;
        if (NT == 64 && WPG == 8 && p.out_mode == OUT_C8 && !p.subsample && p.debug == 0) {
            // lean path (see epi_c8_fast): every 64-wide stride-1 convolution of the network
            auto tiles = [&](auto act_tag) {
                constexpr int ACT = decltype(act_tag)::value;
                const int Co8 = (p.Cout + 7) / 8;
                uint32_t buf = (uint32_t)eg, par = 0;  // buf = t % NB, par = (t / NB) & 1 without divisions
                for (int pr = cid + eg * nclusters; pr < npairs; pr += EG * nclusters) {
                    const int tile = 2 * pr + (int)rank;
                    const bool real = tile < p.num_tiles;
                    int tx, ty, n;
                    tile_coords(p.td, real ? tile : p.num_tiles - 1, tx, ty, n);
                    const int y = ty * TC_ROWS + lq, x = tx * VALID + lane;
                    const bool valid = real && lane < VALID && y < p.H && x < p.W;
                    const uint32_t tempty0 = mapa_rank0(TEMPTY(buf));
                    epi_c8_fast<ACT, BF>(bias_s, reinterpret_cast<__half *>(p.out), p.out_image_stride, p.residual, p.res_image_stride,
                                     p.H, p.W, Co8, tmem_base + buf * ACC + ((uint32_t)(lq * 32) << 16), half, pss * (NT / 8), n, y, x,
                                     valid, TFULL(buf), par, [&] {
                                         tc_fence_before();
                                         __syncwarp();
                                         if (lane == 0) mbar_arrive_cluster(tempty0);
                                     }, n / p.res_div, p.res_pre, p.res_slope);
                    if (pr == cid + eg * nclusters && eg == 0 && warp == TC_EPI_WARP0 && lane == 0) STAMP(4);
                    buf += EG;
                    if (buf >= (uint32_t)NB) { buf -= NB; par ^= 1u; }
                }
            };
            if (p.act == RVSR_ACT_LRELU) tiles(std::integral_constant<int, RVSR_ACT_LRELU>{});
            else if (p.act == RVSR_ACT_RELU) tiles(std::integral_constant<int, RVSR_ACT_RELU>{});
            else tiles(std::integral_constant<int, RVSR_ACT_NONE>{});
        } else
        for (int pr = cid + eg * nclusters; pr < npairs; pr += EG * nclusters, t += EG) {
            const int tile = 2 * pr + (int)rank;
            const bool real = tile < p.num_tiles;
            const int tcl = real ? tile : p.num_tiles - 1;
            const int tx = tcl % p.tiles_x, ty = (tcl / p.tiles_x) % p.tiles_y, n = tcl / (p.tiles_x * p.tiles_y);
            const uint32_t buf = t % NB;
            const int y = ty * TC_ROWS + lq, x = tx * VALID + lane;
            const bool valid = real && lane < VALID && y < p.H && x < p.W;
            ep.prefetch(e, half, pss, n, y, x, valid);
            mbar_wait(TFULL(buf), (t / NB) & 1);
            tc_fence_after();
            ep.run(e, tmem_base + buf * ACC + ((uint32_t)(lq * 32) << 16), half, pss, n, y, x, valid, !(p.debug & 8), !(p.debug & 2), [&] {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_rank0(TEMPTY(buf)));  // the leader's issuer waits for both CTAs
            });
        }
    }
    if (warp == TC_EPI_WARP0 && lane == 0) STAMP(5);
    tc_fence_before();
    cluster_sync_all();   // no CTA exits (or frees TMEM) while its partner can still touch its barriers / operands
    if (threadIdx.x == 0) STAMP(6);
#undef STAMP
    if (warp == 3) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// ---------------------------------------------------------------- host side: tensor maps, packing, launch
static int tc_pick_nt(int Cout, int mode = 0) {
    if (mode == 2) return (Cout % 27 == 0) ? 128 : 0;  // OUT_OM24: 4 deformable groups x 32 columns per pass
    if (Cout <= 16) return 16;
    if (Cout <= 64) return 64;
    if (Cout % 128 == 0) return 128;
    if (Cout % 64 == 0) return 64;
    return 0;
}
static int tc_passes(int Cout, int NT, int mode = 0) { return mode == 2 ? (Cout / 27 + 3) / 4 : (Cout + NT - 1) / NT; }
static int pad16(int c) { return (c + 15) / 16 * 16; }

size_t tc_conv_weight_bytes(int Cout, int Cin, int ks, int mode) {
    const int NT = tc_pick_nt(Cout, mode);
    if (NT == 0 || (ks != 1 && ks != 3) || (mode == 2 && ks != 3)) return 0;
    return (size_t)tc_passes(Cout, NT, mode) * ks * ks * (pad16(Cin) / 8) * NT * 16;
}
size_t tc_dcn_weight_bytes(int Cout, int C, int K) { return (Cout == 64 && C == 64 && K == 9) ? tc_conv_weight_bytes(64, 64, 3) : 0; }

// Weight element (co, cin, tap) is read at w[base + co * s_co + cin * s_ci + tap * s_tap]: OIHW is (Cin * KK, KK, 1, 0); the
// transposed, flipped slice a data-gradient convolution needs is another stride set (WeightView, common.cuh).
__device__ __forceinline__ uint16_t to_16(float v, int bf) {
    return bf ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
}
// conv output channel of column n of N-pass pss (modes: 0 plain, 1 pixel-shuffle order, 2 OUT_OM24 order); >= Cout: padding
__device__ __forceinline__ int pack_col_channel(int n, int pss, int NT, int mode, int Cout) {
    if (mode == 1) return shuffle_col_to_channel(n, pss, NT);
    if (mode == 2) {  // OUT_OM24: column = local group * 32 + [dy0 dx0 .. dy8 dx8 m0 .. m8]
        const int dg = Cout / 27, g = pss * 4 + n / 32, j = n % 32;
        return (g < dg && j < 27) ? (j < 18 ? g * 18 + j : 18 * dg + g * 9 + (j - 18)) : Cout;
    }
    return pss * NT + n;
}
// element i of the single-CTA layout [pass][tap][Q][NT][8]
__device__ __forceinline__ uint16_t pack_tc_elem(const float *__restrict__ w, long long i, int Cout, int Cin, int KK, int Q, int NT, int mode,
                                                 const WeightView &wv) {
    const int e = (int)(i % 8);
    long long r = i / 8;
    const int n = (int)(r % NT);
    r /= NT;
    const int q = (int)(r % Q);
    r /= Q;
    const int tap = (int)(r % KK);
    const int pss = (int)(r / KK);
    const int cin = q * 8 + e, co = pack_col_channel(n, pss, NT, mode, Cout);
    return to_16((co < Cout && cin < Cin) ? w[wv.base + (long long)co * wv.s_co + (long long)cin * wv.s_ci + (long long)tap * wv.s_tap] : 0.f, wv.bf16);
}
// element i of the CTA-pair layout [pass][rank][tap][Q][NT/2][8] (3x3)
__device__ __forceinline__ uint16_t pack_tc2_elem(const float *__restrict__ w, long long i, int Cout, int Q, int NT, int mode, const WeightView &wv) {
    const int NH = NT / 2;
    const int e = (int)(i % 8);
    long long r = i / 8;
    const int nrow = (int)(r % NH);
    r /= NH;
    const int q = (int)(r % Q);
    r /= Q;
    const int tap = (int)(r % 9);
    r /= 9;
    const int rank = (int)(r % 2);
    const int pss = (int)(r / 2);
    const int cin = q * 8 + e, co = pack_col_channel(rank * NH + nrow, pss, NT, mode, Cout);
    return to_16(co < Cout ? w[wv.base + (long long)co * wv.s_co + (long long)cin * wv.s_ci + (long long)tap * wv.s_tap] : 0.f, wv.bf16);
}
__global__ void pack_weight_tc_kernel(const float *__restrict__ w, uint16_t *__restrict__ dst, int Cout, int Cin, int KK,
                                      int Q, int NT, int mode, long long total, WeightView wv) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        dst[i] = pack_tc_elem(w, i, Cout, Cin, KK, Q, NT, mode, wv);
}
int pack_weight_tc(const float *w_oihw, void *dst, int Cout, int Cin, int ks, int mode, cudaStream_t s, const WeightView *view) {
    const WeightView wv = view ? *view : WeightView{0, (long long)Cin * ks * ks, ks * ks, 1, 0};
    const int NT = tc_pick_nt(Cout, mode);
    RVSR_CHECK_ARG(NT != 0, "tc pack: unsupported Cout %d (mode %d)", Cout, mode);
    const int Q = pad16(Cin) / 8;
    const long long total = (long long)tc_passes(Cout, NT, mode) * ks * ks * Q * NT * 8;
    pack_weight_tc_kernel<<<(int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096), 256, 0, s>>>(
        w_oihw, reinterpret_cast<uint16_t *>(dst), Cout, Cin, ks * ks, Q, NT, mode, total, wv);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
int pack_weight_dcn_tc(const float *w_oihw, void *dst, int Cout, int C, int K, cudaStream_t s) {
    RVSR_CHECK_ARG(K == 9, "tc dcn pack: 3x3 only");
    return pack_weight_tc(w_oihw, dst, Cout, C, 3, 0, s);
}

// CTA-pair layout: [pass][rank][tap][Q][NT/2][8] -- each CTA of the pair owns half of the pass's columns
size_t tc2_weight_bytes(int Cout, int Cin, int ks, int mode) {
    const int NT = tc_pick_nt(Cout, mode);
    if (ks != 3 || (NT != 64 && NT != 128) || Cin % 16 != 0) return 0;
    return tc_conv_weight_bytes(Cout, Cin, ks, mode);
}
__global__ void pack_weight_tc2_kernel(const float *__restrict__ w, uint16_t *__restrict__ dst, int Cout, int Cin, int Q, int NT,
                                       int mode, long long total, WeightView wv) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        dst[i] = pack_tc2_elem(w, i, Cout, Q, NT, mode, wv);
}
// Several views of ONE weight tensor in one launch (grid.y = view): the forward operand of a convolution and the transposed,
// flipped operands of its data gradients (training path: one tiny launch per layer and step instead of one per view)
struct PackView {
    uint16_t *dst;
    int Cout, Cin, KK, Q, NT, mode, pair;  // pair: CTA-pair layout (else single-CTA)
    long long total;
    WeightView wv;
};
struct PackViews { PackView v[RVSR_PACK_MAX_VIEWS]; };
__global__ void pack_weight_views_kernel(const float *__restrict__ w, const __grid_constant__ PackViews pv) {
    const PackView &j = pv.v[blockIdx.y];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < j.total; i += (long long)gridDim.x * blockDim.x)
        j.dst[i] = j.pair ? pack_tc2_elem(w, i, j.Cout, j.Q, j.NT, j.mode, j.wv) : pack_tc_elem(w, i, j.Cout, j.Cin, j.KK, j.Q, j.NT, j.mode, j.wv);
}
// views[k]: {Cout, Cin, ks, shuffle mode, pair layout?} + WeightView; dst[k] sized by tc_conv_weight_bytes
int pack_weight_views(const float *w, int n, const int (*dims)[5], const WeightView *wv, void *const *dst, cudaStream_t s) {
    RVSR_CHECK_ARG(n >= 1 && n <= RVSR_PACK_MAX_VIEWS, "tc pack: 1..%d views", RVSR_PACK_MAX_VIEWS);
    PackViews pv;
    memset(&pv, 0, sizeof(pv));
    long long most = 0;
    for (int k = 0; k < n; ++k) {
        const int Cout = dims[k][0], Cin = dims[k][1], ks = dims[k][2], mode = dims[k][3], pair = dims[k][4];
        const int NT = tc_pick_nt(Cout, mode);
        RVSR_CHECK_ARG(NT != 0 && (!pair || tc2_weight_bytes(Cout, Cin, ks, mode) > 0), "tc pack: unsupported view %d", k);
        PackView &j = pv.v[k];
        j.dst = reinterpret_cast<uint16_t *>(dst[k]); j.Cout = Cout; j.Cin = Cin; j.KK = ks * ks; j.NT = NT; j.mode = mode; j.pair = pair;
        j.Q = pair ? Cin / 8 : pad16(Cin) / 8;
        j.total = pair ? (long long)tc_passes(Cout, NT, mode) * 2 * 9 * (Cin / 8) * (NT / 2) * 8
                       : (long long)tc_passes(Cout, NT, mode) * ks * ks * j.Q * NT * 8;
        j.wv = wv[k];
        if (j.total > most) most = j.total;
    }
    const int gx = (int)((most + 255) / 256 < 1024 ? (most + 255) / 256 : 1024);
    pack_weight_views_kernel<<<dim3(gx, n), 256, 0, s>>>(w, pv);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
int pack_weight_tc2(const float *w_oihw, void *dst, int Cout, int Cin, int ks, int mode, cudaStream_t s, const WeightView *view) {
    const WeightView wv = view ? *view : WeightView{0, (long long)Cin * 9, 9, 1, 0};
    RVSR_CHECK_ARG(tc2_weight_bytes(Cout, Cin, ks, mode) > 0, "tc2 pack: unsupported shape");
    const int NT = tc_pick_nt(Cout, mode);
    const long long total = (long long)tc_passes(Cout, NT, mode) * 2 * 9 * (Cin / 8) * (NT / 2) * 8;
    pack_weight_tc2_kernel<<<(int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096), 256, 0, s>>>(
        w_oihw, reinterpret_cast<uint16_t *>(dst), Cout, Cin, Cin / 8, NT, mode, total, wv);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}


struct TcConvPlan {
    int NT, passes, KS, C8s, nstages;
    size_t smem;
    bool pair_only;  // the weights only fit split over a CTA pair (cta_group::2 kernel)
};
static bool tc_conv_plan(const ConvOp &op, TcConvPlan &pl) {
    if (op.w_tc == nullptr || (op.ks != 1 && op.ks != 3) || op.nsrc < 1) return false;
    if (!(op.stride == 1 || (op.stride == 2 && op.ks == 3 && op.out_mode == OUT_C8 && op.H % 2 == 0 && op.W % 2 == 0)))
        return false;
    if (op.out_mode != OUT_C8 && op.out_mode != OUT_C8_SHUFFLE2 && op.out_mode != OUT_PLANAR_F32 &&
        op.out_mode != OUT_OM24 && op.out_mode != OUT_FINAL)
        return false;
    if (op.out_mode == OUT_FINAL && (op.Cout > 8 || op.Cout != op.fin.nc || op.stride != 1 || op.fin.x == nullptr ||
                                     op.fin.scale < 1 || op.H % op.fin.scale != 0 || op.W % op.fin.scale != 0))
        return false;
    const int mode = op.out_mode == OUT_OM24 ? 2 : 0;
    if (mode == 2 && (op.ks != 3 || op.stride != 1 || op.dg <= 0 || op.Cout != 27 * op.dg)) return false;
    if (op.residual != nullptr && (op.out_mode != OUT_C8 || op.Cout > 64)) return false;
    const int C = op.src[0].C;
    for (int i = 0; i < op.nsrc; ++i)
        if (op.src[i].C != C) return false;
    if (C % 16 != 0 || C > 64) return false;
    pl.NT = tc_pick_nt(op.Cout, mode);
    if (pl.NT == 0) return false;
    if (op.out_mode == OUT_C8_SHUFFLE2 && (pl.NT % 64 != 0 || op.Cout % pl.NT != 0)) return false;
    pl.passes = tc_passes(op.Cout, pl.NT, mode);
    pl.KS = op.ks;
    pl.C8s = C / 8;
    const size_t wb = (size_t)op.nsrc * pl.C8s * op.ks * op.ks * pl.NT * 16;
    const size_t stage = (size_t)pl.C8s * (TC_ROWS + op.ks - 1) * TC_TW * 16;
    size_t fixed = wb + 128 + pl.NT * 4 + 512;
    pl.pair_only = false;
    if (fixed + 2 * stage > TC_SMEM_LIMIT) {
        // the CTA-pair kernel keeps half of the weight rows per CTA (nf = 128: 256 -> 64 and 128 -> 216 contractions)
        const bool pair = op.w_tc2 != nullptr && op.ks == 3 && op.stride == 1 && (pl.NT == 128 || (pl.NT == 64 && op.out_mode == OUT_C8));
        fixed = wb / 2 + 128 + pl.NT * 4 + 512;
        if (!pair || fixed + 2 * stage > TC_SMEM_LIMIT) return false;
        pl.pair_only = true;
    }
    pl.nstages = (int)((TC_SMEM_LIMIT - fixed) / stage);
    if (pl.nstages > 6) pl.nstages = 6;
    pl.smem = fixed + pl.nstages * stage + 1024;
    return true;
}
bool tc_conv_supported(const ConvOp &op) {
    TcConvPlan pl;
    return tc_conv_plan(op, pl) && get_encode() != nullptr;
}

template <int KS, int NT, bool BF> static int launch_conv_tc_t(const TcConvParams &p, const TcConvPlan &pl, int sms, cudaStream_t s) {
    RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&conv_tc_kernel<KS, NT, BF>), (int)TC_SMEM_LIMIT + 1024));
    int gx = sms / pl.passes;
    if (gx < 1) gx = 1;
    if (gx > p.num_tiles) gx = p.num_tiles;
    launch_k(conv_tc_kernel<KS, NT, BF>, dim3(gx, pl.passes), dim3(TC_THREADS), pl.smem, s, p);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

static bool tc_two_cta_enabled() {
    static const bool on = !(getenv("RVSR_TC_2CTA") != nullptr && getenv("RVSR_TC_2CTA")[0] == '0');
    return on;
}
// Does a stride-1 launch of `op` run on the CTA-pair kernel, i.e. read the w_tc2 layout (else: w_tc)?  One decision, used by
// launch_conv_tc and by callers that pack only the layout a launch will read (rvsr_c8_conv_layouts).
static bool tc_conv_pair_stride1(const ConvOp &op, const TcConvPlan &pl) {
    const int valid = TC_TW - (op.ks - 1);
    const long long num_tiles = (long long)cdiv(op.W, valid) * cdiv(op.H, TC_ROWS) * op.N;
    return (tc_two_cta_enabled() || pl.pair_only) && op.w_tc2 != nullptr && (pl.NT == 64 || pl.NT == 128) && op.ks == 3 &&
           (num_tiles >= 4 || pl.pair_only) && (pl.NT == 128 || op.out_mode == OUT_C8) && op.stride == 1;
}
// bit 0: the launch reads op.w_tc, bit 1: op.w_tc2; 0: not covered.  op.w_tc / w_tc2 only need to be non-null where that layout exists.
int tc_conv_layouts(const ConvOp &op) {
    TcConvPlan pl;
    if (!tc_conv_plan(op, pl)) return 0;
    return tc_conv_pair_stride1(op, pl) ? 2 : 1;
}

int launch_conv_tc(const ConvOp &op, cudaStream_t s) {
    TcConvPlan pl;
    RVSR_CHECK_ARG(tc_conv_plan(op, pl), "tc conv: unsupported configuration");
    RVSR_CHECK_ARG(!op.bf16 || (op.stride == 1 && (op.out_mode == OUT_C8 || op.out_mode == OUT_C8_SHUFFLE2)),
                   "tc conv: the bf16 instances cover stride 1 with channel-blocked output");
    EncodeTiledFn enc = get_encode();
    RVSR_CHECK_ARG(enc != nullptr, "tc conv: cuTensorMapEncodeTiled unavailable");
    TcConvParams p;
    memset(&p, 0, sizeof(p));
    const int halo_rows = TC_ROWS + op.ks - 1;
    const long long plane = (long long)op.H * op.W * 8;
    for (int i = 0; i < op.nsrc; ++i) {
        const Src &sr = op.src[i];
        RVSR_CHECK_ARG(sr.image_stride % plane == 0, "tc conv: image stride is not a whole number of planes");
        p.src_pstride[i] = (int)(sr.image_stride / plane);
        p.src_frames[i] = sr.frames > 0 ? sr.frames : 1;
        p.src_fixed[i] = sr.fixed_frame;
        p.src_map[i] = sr.map;
        const int reach = sr.map != nullptr ? sr.map_images : op.N;  // images addressable through this source
        const cuuint64_t dims[3] = {(cuuint64_t)op.W * 8, (cuuint64_t)op.H,
                                    (cuuint64_t)(reach - 1) * p.src_pstride[i] + pl.C8s};
        const cuuint64_t strides[2] = {(cuuint64_t)op.W * 16, (cuuint64_t)op.H * op.W * 16};
        const cuuint32_t box[3] = {(cuuint32_t)TC_TW * 8, (cuuint32_t)halo_rows, (cuuint32_t)pl.C8s};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&p.tmap[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(sr.ptr), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tc conv: cuTensorMapEncodeTiled failed (%d) W=%d H=%d planes=%llu", (int)r, op.W, op.H,
                      (unsigned long long)dims[2]);
            return RVSR_E_CUDA;
        }
    }
    p.nsrc = op.nsrc; p.C8s = pl.C8s; p.nstages = pl.nstages;
    p.w = reinterpret_cast<const __half *>(op.w_tc); p.bias = op.bias;
    p.out = op.out; p.out_image_stride = op.out_image_stride;
    p.residual = reinterpret_cast<const __half *>(op.residual); p.res_image_stride = op.res_image_stride;
    p.res_pre = op.res_pre; p.res_div = op.res_div > 0 ? op.res_div : 1; p.bf16 = op.bf16; p.res_slope = op.res_slope;
    p.N = op.N; p.H = op.H; p.W = op.W; p.Cout = op.Cout; p.act = op.act; p.out_mode = op.out_mode;
    p.sig_from = op.sig_from; p.subsample = op.stride == 2 ? 1 : 0; p.dg = op.dg; p.fin = op.fin;
    const int valid = TC_TW - (op.ks - 1);
    p.tiles_x = cdiv(op.W, valid); p.tiles_y = cdiv(op.H, TC_ROWS);
    p.num_tiles = p.tiles_x * p.tiles_y * op.N;
    if (p.num_tiles == 0) return RVSR_OK;
    p.td.tpi = (uint32_t)(p.tiles_x * p.tiles_y); p.td.m_tpi = magic_div(p.td.tpi, (uint32_t)p.num_tiles);
    p.td.tx = (uint32_t)p.tiles_x; p.td.m_tx = magic_div(p.td.tx, p.td.tpi);
    static const int dbg = getenv("RVSR_TC_DEBUG") ? atoi(getenv("RVSR_TC_DEBUG")) : 0;
    p.debug = dbg;
    p.stamp = -1;
    const int sms = sm_count();
    // CTA-pair kernels (cta_group::2) for the 3x3 convolutions with 64- and 128-wide tiles (the bulk of the network)
    const bool two_cta = tc_two_cta_enabled();
    // stride 2 as a real implicit GEMM over the four phase images (RVSR_S2=0: compute at full resolution and subsample)
    static const bool s2_on = !(getenv("RVSR_S2") != nullptr && getenv("RVSR_S2")[0] == '0');
    const bool s2 = s2_on && two_cta && op.stride == 2 && op.w_tc2 != nullptr && pl.NT == 64 && op.ks == 3 && op.out_mode == OUT_C8 &&
                    op.residual == nullptr && op.H % 2 == 0 && op.W % 2 == 0 &&
                    // two stages of four phase tiles next to this CTA's weight half (else: full resolution + subsampled store)
                    (size_t)op.nsrc * pl.C8s * 9 * 32 * 16 + 128 + 64 * 4 + 512 + 2 * (size_t)4 * pl.C8s * (TC_ROWS + 1) * TC_TW * 16 <= TC_SMEM_LIMIT;
    if (s2) {
        const int Ho = op.H / 2, Wo = op.W / 2;
        for (int i = 0; i < op.nsrc; ++i) {  // 4-D maps {8 channels, W, H, planes}, every other pixel of every other row
            const Src &sr = op.src[i];
            const int reach = sr.map != nullptr ? sr.map_images : op.N;
            const cuuint64_t dims[4] = {8, (cuuint64_t)op.W, (cuuint64_t)op.H, (cuuint64_t)(reach - 1) * p.src_pstride[i] + pl.C8s};
            const cuuint64_t strides[3] = {16, (cuuint64_t)op.W * 16, (cuuint64_t)op.H * op.W * 16};
            const cuuint32_t box[4] = {8, 2 * TC_TW, 2 * (TC_ROWS + 1), (cuuint32_t)pl.C8s};
            const cuuint32_t estr[4] = {1, 2, 2, 1};
            CUresult r = enc(&p.tmap[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(sr.ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("tc conv: cuTensorMapEncodeTiled (stride 2) failed (%d)", (int)r); return RVSR_E_CUDA; }
        }
        p.H = Ho; p.W = Wo; p.subsample = 0;
        p.tiles_x = cdiv(Wo, valid); p.tiles_y = cdiv(Ho, TC_ROWS);
        p.num_tiles = p.tiles_x * p.tiles_y * op.N;
        p.td.tpi = (uint32_t)(p.tiles_x * p.tiles_y); p.td.m_tpi = magic_div(p.td.tpi, (uint32_t)p.num_tiles);
        p.td.tx = (uint32_t)p.tiles_x; p.td.m_tx = magic_div(p.td.tx, p.td.tpi);
    }
    if (tc_conv_pair_stride1(op, pl) ||
        (s2 && (two_cta || pl.pair_only) && op.w_tc2 != nullptr && pl.NT == 64 && op.ks == 3 && (p.num_tiles >= 4 || pl.pair_only) && op.out_mode == OUT_C8)) {
        const size_t wb2 = (size_t)op.nsrc * pl.C8s * 9 * (pl.NT / 2) * 16;
        const size_t stage = s2 ? (size_t)4 * pl.C8s * (TC_ROWS + 1) * TC_TW * 16 : (size_t)pl.C8s * (TC_ROWS + 2) * TC_TW * 16;
        const size_t fixed = wb2 + 128 + pl.NT * 4 + 512;
        int st2 = (int)((TC_SMEM_LIMIT - fixed) / stage);
        if (st2 > 6) st2 = 6;
        RVSR_CHECK_ARG(st2 >= 2, "tc conv: not enough shared memory for two stages");
        p.nstages = st2;
        p.w = reinterpret_cast<const __half *>(op.w_tc2);
        const size_t smem2 = fixed + (size_t)st2 * stage + 1024;
        if (stamps_on() && g_stamp_next < STAMP_SLOTS) {
            p.stamp = g_stamp_next;
            snprintf(g_stamp_label[g_stamp_next], sizeof(g_stamp_label[0]), "N%d %dx%d src%d co%d m%d tiles%d", op.N, op.H, op.W, op.nsrc,
                     op.Cout, op.out_mode, p.num_tiles);
            ++g_stamp_next;
        }
        const int npairs = (p.num_tiles + 1) / 2;
        int clusters = (sms / 2) / pl.passes;
        if (clusters < 1) clusters = 1;
        if (clusters > npairs) clusters = npairs;
        if (s2) {
            RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&conv_tc2_kernel<3, 64, true>), (int)TC_SMEM_LIMIT + 1024));
            launch_k(conv_tc2_kernel<3, 64, true>, dim3(2 * clusters, pl.passes), dim3(TC_THREADS), smem2, s, p);
        } else if (pl.NT == 64 && op.bf16) {  // bf16 operands / storage (training path): separate instances, the fp16 code is untouched
            RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&conv_tc2_kernel<3, 64, false, true>), (int)TC_SMEM_LIMIT + 1024));
            launch_k(conv_tc2_kernel<3, 64, false, true>, dim3(2 * clusters, pl.passes), dim3(TC_THREADS), smem2, s, p);
        } else if (pl.NT == 64) {
            RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&conv_tc2_kernel<3, 64>), (int)TC_SMEM_LIMIT + 1024));
            launch_k(conv_tc2_kernel<3, 64>, dim3(2 * clusters, pl.passes), dim3(TC_THREADS), smem2, s, p);
        } else if (op.bf16) {
            RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&conv_tc2_kernel<3, 128, false, true>), (int)TC_SMEM_LIMIT + 1024));
            launch_k(conv_tc2_kernel<3, 128, false, true>, dim3(2 * clusters, pl.passes), dim3(TC_THREADS), smem2, s, p);
        } else {
            RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&conv_tc2_kernel<3, 128>), (int)TC_SMEM_LIMIT + 1024));
            launch_k(conv_tc2_kernel<3, 128>, dim3(2 * clusters, pl.passes), dim3(TC_THREADS), smem2, s, p);
        }
        RVSR_LAUNCH_CHECK();
        return RVSR_OK;
    }
    RVSR_CHECK_ARG(!pl.pair_only, "tc conv: this shape needs the CTA-pair kernel");
#define RVSR_TC_CASE(KS_, NT_) \
    if (op.ks == KS_ && pl.NT == NT_) return op.bf16 ? launch_conv_tc_t<KS_, NT_, true>(p, pl, sms, s) : launch_conv_tc_t<KS_, NT_, false>(p, pl, sms, s);
    RVSR_TC_CASE(3, 16) RVSR_TC_CASE(3, 64) RVSR_TC_CASE(3, 128)
    RVSR_TC_CASE(1, 16) RVSR_TC_CASE(1, 64) RVSR_TC_CASE(1, 128)
#undef RVSR_TC_CASE
    set_error("tc conv: no kernel instance for ks=%d NT=%d", op.ks, pl.NT);
    return RVSR_E_UNSUPPORTED;
}

// ---------------------------------------------------------------- modulated deformable conv (gather -> UMMA)
struct TcDcnParams {
    const __half *x;
    long long x_image_stride;
    const int *x_map;         // optional image -> slot table (feature cache)
    const uint4 *om;          // OUT_OM24 offsets + mask
    long long om_stride;      // uint4 units per image
    const __half *w;
    const float *bias;
    __half *out;
    long long out_image_stride;
    int N, H, W, cpg, act;
    int tiles_x, tiles_y, num_tiles;
    int debug;  // RVSR_DCN_DEBUG timing experiments (results wrong): 1 = no fence.proxy.async after the gather stores
    // nf = 128 (NT = 128): one launch contracts 64 input channels (8 blocks starting at blk0) into all 128 outputs; the
    // second launch adds its result to the first one's partial sums (fp16, no bias) and finishes with bias + activation
    int blk0, accum, finish;
};
constexpr int DCN_GATHER_WARPS = 16, DCN_THREADS = 32 * (1 + DCN_GATHER_WARPS + 4), DCN_STAGES = 3;  // 2..4 measure the same; 6 is slower (L1 capacity)
constexpr int DCN_TAP_BYTES = 8 * 128 * 16;  // 8 channel blocks x 128 pixels x 16 B

template <bool BLEND16, int NT>
__global__ void __launch_bounds__(DCN_THREADS, 1) dcn_tc_kernel(const __grid_constant__ TcDcnParams p) {
    constexpr int K = 9, Q = 8, ACC = NT, TMEM_COLS = 2 * NT;
    static_assert(NT == 64 || NT == 128, "DCN contraction tile: 64 or 128 output channels");
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t w_bytes = Q * K * NT * 16;
    uint8_t *w_s = smem;
    uint8_t *tap_s = smem + w_bytes;
    float *bias_s = reinterpret_cast<float *>(tap_s + DCN_STAGES * DCN_TAP_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + NT);
    constexpr int S = DCN_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 5);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) mbar_init(BAR(i), DCN_GATHER_WARPS);  // full: one arrive per gather warp
        for (int i = S; i < 2 * S + 3; ++i) mbar_init(BAR(i), 1);
        mbar_init(BAR(2 * S + 3), 4);
        mbar_init(BAR(2 * S + 4), 4);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < NT; i += DCN_THREADS) bias_s[i] = p.bias != nullptr ? p.bias[i] : 0.f;
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);

    if (warp == 0) {
        {   // whole warp on uniform values; the tcgen05 / bulk-copy instructions are predicated on one elected lane
            if (elect_one()) {
                mbar_expect_tx(BAR(2 * S), w_bytes);
                for (uint32_t o = 0; o < w_bytes; o += 24576)
                    bulk_load(smem_u32(w_s + o), reinterpret_cast<const uint8_t *>(p.w) + o, 24576, BAR(2 * S));
            }
            __syncwarp();
            constexpr uint32_t idesc = make_idesc(NT);
            mbar_wait(BAR(2 * S), 0);
            uint32_t t = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++t) {
                const uint32_t buf = t & 1;
                mbar_wait(BAR(2 * S + 3 + buf), ((t >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + buf * ACC;
                // 18 gather steps per tile (2 channel halves x 9 taps), two steps per ring stage
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    constexpr int dummy = 0; (void)dummy;
                    const int st = j % S;
                    mbar_wait_idle(BAR(st), (t * 3 + j / S) & 1);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int part = 0; part < 2; ++part) {
                            const int s18 = 2 * j + part, h = s18 / 9, tap = s18 % 9;
                            const uint32_t a0 = smem_u32(tap_s + st * DCN_TAP_BYTES + part * (DCN_TAP_BYTES / 2));
                            const uint32_t b0 = smem_u32(w_s) + (uint32_t)(tap * Q + 4 * h) * (NT * 16);
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk)
                                umma_f16(d, make_desc(a0 + (uint32_t)(2 * kk) * 2048, 2048, 128),
                                         make_desc(b0 + (uint32_t)(2 * kk) * (NT * 16), NT * 16, 128), idesc, (s18 | kk) ? 1u : 0u);
                        }
                        umma_commit(BAR(S + st));
                        if (j == 8) umma_commit(BAR(2 * S + 1 + buf));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp <= DCN_GATHER_WARPS) {
        // ---- gather: thread -> pixel m of the tile and channel block qq of the current half (4 blocks).
        // A tile is 18 steps (half 0 taps 0..8, half 1 taps 0..8), one bilinear sample per thread and step.
        // Software pipeline: the four corner loads of step s+1 (also across halves and tiles) are issued
        // before step s is blended, so the loads of a warp are never drained.  Branch-free: corner addresses
        // are clamped into the image and out-of-range corners get a zero weight.  The loop is fully unrolled
        // so the offset / mask registers (three 32 B blocks per pixel and group) are indexed statically.
        static_assert(S == 3, "stage schedule below assumes 9 two-step stages over a ring of 3");
        pdl_wait();  // features and offsets / mask come from the previous kernels
        const int gt = threadIdx.x - 32;
        const int m = gt & 127, qq = gt >> 7;
        const long long plane = (long long)p.H * p.W;
        const float Hf = (float)p.H, Wf = (float)p.W;
        const int Hm1 = p.H - 1, Wm1 = p.W - 1;
        struct Unit { const uint4 *pl; const uint4 *om; float by, bx; bool valid; };
        struct Samp { uint4 c[4]; uint32_t w[4]; };
        auto unit_of = [&](int tile, int half, bool pf) {
            Unit u;
            if (tile >= p.num_tiles) tile = p.num_tiles - 1;  // prefetch past the end: a harmless re-read
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, n = tile / (p.tiles_x * p.tiles_y);
            const int y = ty * TC_ROWS + (m >> 5), x = tx * TC_TW + (m & 31);
            u.valid = y < p.H && x < p.W;
            const long long pix = u.valid ? (long long)y * p.W + x : 0;
            const int blk = p.blk0 + half * 4 + qq, g = (blk * 8) / p.cpg;
            const long long img = p.x_map != nullptr ? __ldg(p.x_map + n) : n;
            u.pl = reinterpret_cast<const uint4 *>(p.x + img * p.x_image_stride + (long long)blk * plane * 8);
            u.om = p.om + (long long)n * p.om_stride + ((long long)(g * 3) * plane + pix) * 2;
            u.by = (float)(y - 1); u.bx = (float)(x - 1);
            if (pf) {
                // several steps ahead of the first use: the unit's offset/mask blocks into L2 (streamed from
                // DRAM) and the undeformed 3x3 neighbourhood of its feature planes into L1
                prefetch_l2(u.om); prefetch_l2(u.om + 2 * plane); prefetch_l2(u.om + 4 * plane);
                const int xc = min(x, Wm1), yc = min(y, Hm1);
                prefetch_l1(u.pl + (max(yc - 1, 0) * p.W + xc));
                prefetch_l1(u.pl + (yc * p.W + xc));
                prefetch_l1(u.pl + (min(yc + 1, Hm1) * p.W + xc));
            }
            return u;
        };
        auto issue = [&](Samp &sm, const Unit &u, int tap, float dy, float dx, float mk) {
            const float py = u.by + (float)(tap / 3) + dy, px = u.bx + (float)(tap % 3) + dx;
            const bool inside = u.valid && py > -1.f && px > -1.f && py < Hf && px < Wf;
            const float fy = floorf(inside ? py : 0.f), fx = floorf(inside ? px : 0.f);
            const int y0 = (int)fy, x0 = (int)fx;
            const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
            const float m_in = inside ? mk : 0.f;  // mask folded into the weights
            const float wy0 = y0 >= 0 ? hy * m_in : 0.f, wy1 = y0 + 1 <= Hm1 ? ly * m_in : 0.f;
            const float wx0 = x0 >= 0 ? hx : 0.f, wx1 = x0 + 1 <= Wm1 ? lx : 0.f;
            const float w0 = wy0 * wx0, w1 = wy0 * wx1, w2 = wy1 * wx0, w3 = wy1 * wx1;
            const int r0 = max(y0, 0) * p.W, r1 = min(y0 + 1, Hm1) * p.W, x0c = max(x0, 0), x1c = min(x0 + 1, Wm1);
            sm.c[0] = __ldg(u.pl + (r0 + x0c)); sm.c[1] = __ldg(u.pl + (r0 + x1c));
            sm.c[2] = __ldg(u.pl + (r1 + x0c)); sm.c[3] = __ldg(u.pl + (r1 + x1c));
            if (BLEND16) {
                // corner weights (computed in fp32) rounded to fp16x2 for the HFMA2 blend
                __half2 h0 = __float2half2_rn(w0), h1 = __float2half2_rn(w1), h2 = __float2half2_rn(w2), h3 = __float2half2_rn(w3);
                sm.w[0] = *reinterpret_cast<uint32_t *>(&h0); sm.w[1] = *reinterpret_cast<uint32_t *>(&h1);
                sm.w[2] = *reinterpret_cast<uint32_t *>(&h2); sm.w[3] = *reinterpret_cast<uint32_t *>(&h3);
            } else {
                sm.w[0] = __float_as_uint(w0); sm.w[1] = __float_as_uint(w1);
                sm.w[2] = __float_as_uint(w2); sm.w[3] = __float_as_uint(w3);
            }
        };
        auto blend = [&](const Samp &sm) {
            uint4 pk;
            if (BLEND16) {
                // fp16x2 blend: 16 HFMA2 instead of 32 conversions + 32 FFMA; adds <= 3 fp16 roundings to a
                // value that is stored as fp16 anyway.
                __half2 *o = reinterpret_cast<__half2 *>(&pk);
                const __half2 *wk = reinterpret_cast<const __half2 *>(sm.w);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    __half2 a = __hmul2(wk[0], reinterpret_cast<const __half2 *>(&sm.c[0])[j]);
                    a = __hfma2(wk[1], reinterpret_cast<const __half2 *>(&sm.c[1])[j], a);
                    a = __hfma2(wk[2], reinterpret_cast<const __half2 *>(&sm.c[2])[j], a);
                    o[j] = __hfma2(wk[3], reinterpret_cast<const __half2 *>(&sm.c[3])[j], a);
                }
            } else {
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const __half2 *h = reinterpret_cast<const __half2 *>(&sm.c[k]);
                    const float wk = __uint_as_float(sm.w[k]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __half22float2(h[j]);
                        v[2 * j] = fmaf(wk, f.x, v[2 * j]);
                        v[2 * j + 1] = fmaf(wk, f.y, v[2 * j + 1]);
                    }
                }
                __half2 *h = reinterpret_cast<__half2 *>(&pk);
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            }
            return pk;
        };
        // offsets / mask registers: A = taps 0..3, B = taps 4..7, C = [dy8 dx8 m01 m23][m45 m67 m8_ 0]
        uint4 A[2][2], B[2], C[2][2];
        Unit U[2];
        Samp SM[2];
        auto off_y = [&](int h, int t) -> float {
            const uint4 &r = t < 4 ? A[h][t >> 1] : (t < 8 ? B[(t - 4) >> 1] : C[h][0]);
            return __uint_as_float(t == 8 ? r.x : ((t & 1) ? r.z : r.x));
        };
        auto off_x = [&](int h, int t) -> float {
            const uint4 &r = t < 4 ? A[h][t >> 1] : (t < 8 ? B[(t - 4) >> 1] : C[h][0]);
            return __uint_as_float(t == 8 ? r.y : ((t & 1) ? r.w : r.y));
        };
        auto mask_of = [&](int h, int t) -> float {
            const int wi = t >> 1;
            const uint32_t word = wi == 0 ? C[h][0].z : wi == 1 ? C[h][0].w : wi == 2 ? C[h][1].x : wi == 3 ? C[h][1].y : C[h][1].z;
            const __half2 mh = *reinterpret_cast<const __half2 *>(&word);
            return (t & 1) ? __high2float(mh) : __low2float(mh);
        };
        auto load_AC = [&](int h) {
            ld_global_nc_256(U[h].om, A[h][0], A[h][1]);
            ld_global_nc_256(U[h].om + 4 * plane, C[h][0], C[h][1]);
        };
        U[0] = unit_of(blockIdx.x, 0, false);
        load_AC(0);
        issue(SM[0], U[0], 0, off_y(0, 0), off_x(0, 0), mask_of(0, 0));
        uint32_t t = 0;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++t) {
#pragma unroll
            for (int s18 = 0; s18 < 18; ++s18) {
                const int h = s18 / 9, tap = s18 % 9, st = (s18 >> 1) % S, part = s18 & 1;
                if (tap == 0) ld_global_nc_256(U[h].om + 2 * plane, B[0], B[1]);
                if (tap == 1)  // set up the next unit (other half of this tile, or the first half of the next tile)
                    U[h ^ 1] = h == 0 ? unit_of(tile, 1, true) : unit_of(tile + gridDim.x, 0, true);
                if (tap == 6) load_AC(h ^ 1);
                if (tap < 8) issue(SM[(s18 + 1) & 1], U[h], tap + 1, off_y(h, tap + 1), off_x(h, tap + 1), mask_of(h, tap + 1));
                else issue(SM[(s18 + 1) & 1], U[h ^ 1], 0, off_y(h ^ 1, 0), off_x(h ^ 1, 0), mask_of(h ^ 1, 0));
                const uint4 pk = blend(SM[s18 & 1]);
                if (part == 0) mbar_wait(BAR(S + st), ((t * 3 + (s18 >> 1) / S) & 1) ^ 1);  // stage free (its MMAs completed)
                *reinterpret_cast<uint4 *>(tap_s + st * DCN_TAP_BYTES + part * (DCN_TAP_BYTES / 2) + qq * 2048 + m * 16) = pk;
                if (part == 1) {
                    if (!(p.debug & 1)) fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(st));
                }
            }
        }
    } else {
        pdl_wait();
        const int lq = warp & 3;
        EpiArgs e{bias_s, p.out, p.out_image_stride, nullptr, 0, p.H, p.W, NT, p.act, OUT_C8, 0, 0, 0, nullptr, 0, 1};
        EpiTile<64, 1> ep;
        ep.has_res = false;
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++t) {
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, n = tile / (p.tiles_x * p.tiles_y);
            const uint32_t buf = t & 1;
            mbar_wait_idle(BAR(2 * S + 1 + buf), (t >> 1) & 1);  // a tile takes ~10k cycles to gather
            tc_fence_after();
            const int y = ty * TC_ROWS + lq, x = tx * TC_TW + lane;
            const uint32_t taddr = tmem_base + buf * ACC + ((uint32_t)(lq * 32) << 16);
            auto release = [&] {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(2 * S + 3 + buf));
            };
            if constexpr (NT == 64) {
                ep.run(e, taddr, 0, 0, n, y, x, y < p.H && x < p.W, true, true, release);
            } else {
                // 128 columns, 32 at a time; partial sums of the first input half are read back from `out` (same thread, in place)
                const bool valid = y < p.H && x < p.W;
                uint4 *o = reinterpret_cast<uint4 *>(p.out + (long long)n * p.out_image_stride) + (long long)y * p.W + x;
                const long long plane = (long long)p.H * p.W;
#pragma unroll 1
                for (int hc = 0; hc < NT / 32; ++hc) {
                    uint32_t a[2][16];
                    tmem_ld16_nowait(taddr + hc * 32, a[0]);
                    tmem_ld16_nowait(taddr + hc * 32 + 16, a[1]);
                    tmem_ld_wait();
                    if (hc == NT / 32 - 1) release();
                    if (!valid) continue;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(a[q >> 1][(q & 1) * 8 + i]);
                        uint4 *dst = o + (long long)(hc * 4 + q) * plane;
                        if (p.accum) {
                            const uint4 prev = *dst;
                            const __half2 *h = reinterpret_cast<const __half2 *>(&prev);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 f = __half22float2(h[i]);
                                v[2 * i] += f.x; v[2 * i + 1] += f.y;
                            }
                        }
                        if (p.finish) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = apply_act(v[i] + bias_s[hc * 32 + q * 8 + i], p.act);
                        }
                        uint4 pk;
                        __half2 *h = reinterpret_cast<__half2 *>(&pk);
#pragma unroll
                        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                        *dst = pk;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

bool tc_dcn_supported(const DcnOp &op) {
    const bool c64 = op.x.C == 64 && op.Cout == 64, c128 = op.x.C == 128 && op.Cout == 128 && op.w_tc_hi != nullptr;
    if (op.w_tc == nullptr || !(c64 || c128) || op.kh != 3 || op.kw != 3) return false;
    if (op.stride != 1 || op.pad != 1 || op.dil != 1 || op.out_mode != OUT_C8) return false;
    const int cpg = op.x.C / op.dg;
    if (cpg * op.dg != op.x.C || cpg % 8 != 0) return false;
    if (op.x.fixed_frame >= 0 || op.om24 == nullptr) return false;
    return true;
}

int launch_dcn_tc(const DcnOp &op, cudaStream_t s) {
    RVSR_CHECK_ARG(tc_dcn_supported(op), "tc dcn: unsupported configuration");
    TcDcnParams p;
    p.x = reinterpret_cast<const __half *>(op.x.ptr); p.x_image_stride = op.x.image_stride; p.x_map = op.x.map;
    p.om = reinterpret_cast<const uint4 *>(op.om24); p.om_stride = op.om24_image_stride / 4;
    p.w = reinterpret_cast<const __half *>(op.w_tc); p.bias = op.bias;
    p.out = reinterpret_cast<__half *>(op.out); p.out_image_stride = op.out_image_stride;
    p.N = op.N; p.H = op.H; p.W = op.W; p.cpg = op.x.C / op.dg; p.act = op.act;
    p.tiles_x = cdiv(op.W, TC_TW); p.tiles_y = cdiv(op.H, TC_ROWS); p.num_tiles = p.tiles_x * p.tiles_y * op.N;
    if (p.num_tiles == 0) return RVSR_OK;
    static const int ddbg = getenv("RVSR_DCN_DEBUG") ? atoi(getenv("RVSR_DCN_DEBUG")) : 0;
    p.debug = ddbg;
    static const bool blend32 = getenv("RVSR_DCN_BLEND") != nullptr && strcmp(getenv("RVSR_DCN_BLEND"), "fp32") == 0;
    int gx = sm_count();
    if (gx > p.num_tiles) gx = p.num_tiles;
    p.blk0 = 0; p.accum = 0; p.finish = 1;
    if (op.x.C == 64) {
        const size_t smem = 8 * 9 * 64 * 16 + DCN_STAGES * DCN_TAP_BYTES + 64 * 4 + 256 + 1024;
        RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&dcn_tc_kernel<true, 64>), (int)smem));
        RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&dcn_tc_kernel<false, 64>), (int)smem));
        // RVSR_DCN_BLEND=fp32 keeps the bilinear blend in fp32 (one rounding per sample instead of four)
        if (blend32)
            launch_k(dcn_tc_kernel<false, 64>, dim3(gx), dim3(DCN_THREADS), smem, s, p);
        else
            launch_k(dcn_tc_kernel<true, 64>, dim3(gx), dim3(DCN_THREADS), smem, s, p);
        RVSR_LAUNCH_CHECK();
        return RVSR_OK;
    }
    // nf = 128: the 128 x 128 x 9 weights (295 KB) do not fit: two launches over the input-channel halves
    const size_t smem = 8 * 9 * 128 * 16 + DCN_STAGES * DCN_TAP_BYTES + 128 * 4 + 256 + 1024;
    RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&dcn_tc_kernel<true, 128>), (int)smem));
    RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&dcn_tc_kernel<false, 128>), (int)smem));
    for (int h = 0; h < 2; ++h) {
        p.w = reinterpret_cast<const __half *>(h == 0 ? op.w_tc : op.w_tc_hi);
        p.blk0 = 8 * h; p.accum = h; p.finish = h;
        if (blend32)
            launch_k(dcn_tc_kernel<false, 128>, dim3(gx), dim3(DCN_THREADS), smem, s, p);
        else
            launch_k(dcn_tc_kernel<true, 128>, dim3(gx), dim3(DCN_THREADS), smem, s, p);
        RVSR_LAUNCH_CHECK();
    }
    return RVSR_OK;
}

// contraction weights of one input-channel half (nf = 128): [tap][8 blocks][128 rows][8], value = w[n][h * 64 + q * 8 + e][tap]
size_t tc_dcn_half_weight_bytes() { return (size_t)9 * 8 * 128 * 16; }
__global__ void pack_weight_dcn_half_kernel(const float *__restrict__ w, __half *__restrict__ dst, int h, int total) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int e = i % 8, n = (i / 8) % 128, q = (i / 1024) % 8, tap = i / 8192;
        dst[i] = __float2half_rn(w[((long long)n * 128 + h * 64 + q * 8 + e) * 9 + tap]);
    }
}
int pack_weight_dcn_tc_half(const float *w_oihw, void *dst, int h, cudaStream_t s) {
    const int total = 9 * 8 * 128 * 8;
    pack_weight_dcn_half_kernel<<<(total + 255) / 256, 256, 0, s>>>(w_oihw, reinterpret_cast<__half *>(dst), h, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}


// ---------------------------------------------------------------- conv_last, "taps in N" (+ base frame)
// The last convolution has nc (<= 3) output channels: as an ordinary implicit GEMM it needs 36 MMAs per tile that
// each read a full 4 KB A operand for 16 (3 useful) output columns -- bound by the A-operand shared-memory reads.
// Here the nine taps move into the N dimension instead:
//     P[pixel][tap * nc + co] = sum_c in[pixel][c] * W[co][c][tap]          ONE unshifted view, K = 64: 4 MMAs (N = 32)
//     out[co][y][x]           = sum_tap P[(y + dy - 1, x + dx - 1)][tap * nc + co]   shift-and-add in the epilogue
// A 6 x 32 halo tile is covered by two M = 128 views (rows 0-3 and rows 2-5): 8 MMAs per 4 x 30 output tile instead
// of 36.  The epilogue moves the partial sums TMEM -> registers -> shared memory ([column][row][x] fp32, conflict
// free), synchronises its 8 warps, and every thread then gathers the 9 * nc partials of its output pixel, adds bias
// and the bilinearly upsampled centre LQ frame (EDVR_arch.py:315-319) and writes the NCHW result.
struct alignas(64) TcTapnParams {
    CUtensorMap tmap;
    const __half *w;     // [C8s][32][8]: row n = tap * nc + co
    const float *bias;   // [nc]
    void *out;           // NCHW, fin.out_dtype
    FinalAdd fin;
    int N, H, W, nc, C8s, nstages;
    int tiles_x, tiles_y, num_tiles;
    TileDiv td;
    int debug;  // RVSR_TC_DEBUG timing experiments: 1 no MMAs, 2 no output phase, 4 no halo loads, 8 no TMEM drain
};
// three epilogue groups of 8 warps take tiles round-robin: one tile's epilogue is a ~2.4k-cycle latency chain
// (accumulator wait, TMEM drain, two group barriers, shared-memory gather, stores)
constexpr int TAPN_EPI_WARP0 = 4, TAPN_EG = 3, TAPN_EPI_WARPS = 8 * TAPN_EG, TAPN_THREADS = 32 * (TAPN_EPI_WARP0 + TAPN_EPI_WARPS);
constexpr int TAPN_NB = 2 * TAPN_EG, TAPN_XCH = 27 * 6 * 32;  // accumulator pairs in flight; floats of one exchange buffer
constexpr int TAPN_TMEM_COLS = TAPN_NB * 64 <= 256 ? 256 : 512;

__global__ void __launch_bounds__(TAPN_THREADS, 1) conv_tapn_kernel(const __grid_constant__ TcTapnParams p) {
    constexpr int HALO_ROWS = 6, PLANE_BYTES = HALO_ROWS * TC_TW * 16, VALID = TC_TW - 2, NB = TAPN_NB, EG = TAPN_EG, WPG = TAPN_EPI_WARPS / EG;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int S = p.nstages;
    const uint32_t w_bytes = (uint32_t)p.C8s * 32 * 16, stage_bytes = (uint32_t)p.C8s * PLANE_BYTES;
    uint8_t *w_s = smem;
    uint8_t *stage_s = smem + w_bytes;
    float *xch = reinterpret_cast<float *>(stage_s + (size_t)S * stage_bytes + 128);
    float *bias_s = xch + EG * TAPN_XCH;
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + 8);
    // bars: S full, S empty, weights-full, NB accumulator-full, NB accumulator-empty
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 1 + 2 * NB);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    auto FULL = [&](int st) { return BAR(st); };
    auto EMPTY = [&](int st) { return BAR(S + st); };
    const uint32_t WFULL = BAR(2 * S);
    auto TFULL = [&](int b) { return BAR(2 * S + 1 + b); };
    auto TEMPTY = [&](int b) { return BAR(2 * S + 1 + NB + b); };
    const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * S + 1 + NB; ++i) mbar_init(BAR(i), 1);
        for (int i = 0; i < NB; ++i) mbar_init(TEMPTY(i), WPG);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), TAPN_TMEM_COLS);  // NB x (2 views x 32 columns)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(WFULL, w_bytes);
            bulk_load(smem_u32(w_s), p.w, w_bytes, WFULL);
            prefetch_tensormap(&p.tmap);
            pdl_wait();
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                int tx, ty, n;
                tile_coords(p.td, tile, tx, ty, n);
                const int st = it % S;
                mbar_wait(EMPTY(st), ((it / S) & 1) ^ 1);
                if (p.debug & 4) { mbar_arrive(FULL(st)); continue; }
                mbar_expect_tx(FULL(st), stage_bytes);
                tma_load_3d(smem_u32(stage_s + (size_t)st * stage_bytes), &p.tmap, FULL(st), (tx * VALID - 1) * 8, ty * TC_ROWS - 1,
                            n * p.C8s);
            }
        }
    } else if (warp == 1) {
        // ---- issuer: whole warp on uniform values (see elect_one)
        constexpr uint32_t idesc = make_idesc(32);
        const uint64_t adesc0 = make_desc(smem_u32(stage_s), PLANE_BYTES, 128);
        const uint64_t bdesc0 = make_desc(smem_u32(w_s), 32 * 16, 128);
        const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
        const uint32_t a_base = (uint32_t)adesc0, b_base = (uint32_t)bdesc0;
        const uint32_t stage_units = stage_bytes >> 4;
        const int nk = (p.debug & 1) ? 0 : p.C8s / 2;
        mbar_wait(WFULL, 0);
        uint32_t t = 0, st = 0, sph = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++t) {
            const uint32_t buf = t % NB;
            mbar_wait(TEMPTY(buf), ((t / NB) & 1) ^ 1);
            mbar_wait(FULL(st), sph);
            tc_fence_after();
            const uint32_t d = tmem_base + buf * 64;
            const uint32_t a_lo0 = a_base + st * stage_units;
            if (elect_one()) {
#pragma unroll
                for (int v = 0; v < 2; ++v)  // view v: tile rows 2v .. 2v + 3
                    for (int kk = 0; kk < nk; ++kk)
                        umma_f16(d + (uint32_t)v * 32, ((uint64_t)a_hi << 32) | (a_lo0 + (uint32_t)v * (2 * TC_TW) + (uint32_t)kk * (2 * PLANE_BYTES / 16)),
                                 ((uint64_t)b_hi << 32) | (b_base + (uint32_t)kk * (2 * 32)), idesc, kk ? 1u : 0u);
                umma_commit(EMPTY(st));
                umma_commit(TFULL(buf));
            }
            __syncwarp();
            if (++st == (uint32_t)S) { st = 0; sph ^= 1u; }
        }
    } else if (warp >= TAPN_EPI_WARP0) {
        pdl_wait();
        const int wi = (warp - TAPN_EPI_WARP0) % WPG, eg = (warp - TAPN_EPI_WARP0) / WPG;
        const int lq = warp & 3, hv = wi >> 2;       // TMEM lane quarter; view this warp drains (phase 1) / channel set (phase 2)
        const int nc = p.nc, ncol = 9 * nc;
        if (threadIdx.x - 32 * TAPN_EPI_WARP0 < 8) bias_s[threadIdx.x - 32 * TAPN_EPI_WARP0] =
            (p.bias != nullptr && (int)(threadIdx.x - 32 * TAPN_EPI_WARP0) < nc) ? p.bias[threadIdx.x - 32 * TAPN_EPI_WARP0] : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TAPN_EPI_WARPS) : "memory");
        float *xg = xch + eg * TAPN_XCH;
        const FinalAdd &f = p.fin;
        const int Hl = p.H / f.scale, Wl = p.W / f.scale;
        const long long lplane = (long long)Hl * Wl, hplane = (long long)p.H * p.W;
        const bool reader = hv == 0 || lq >= 2;      // view 1's lanes 0..63 repeat tile rows 2, 3
        const int row = hv == 0 ? lq : lq + 2;       // tile row this warp's TMEM lanes hold
        // ---- base frame pixels of this thread's output pixel: the 4 bilinear corners per channel are fetched one tile
        // AHEAD (raw values in registers), so their global-memory latency is hidden behind the previous tile's work;
        // 8 warps per tile: channels 0, 1 on the view-0 warps, channel 2 on the view-1 warps
        const int co0 = hv == 0 ? 0 : 2, co1 = min(nc, hv == 0 ? 2 : 3);
        struct Pix { uint32_t c[2][4]; float ly, lx; int n, y, x; bool valid; };  // c: RAW loaded bits (converting here would wait for the load)
        auto fetch = [&](int tile, Pix &q) {
            int tx, ty;
            tile_coords(p.td, tile, tx, ty, q.n);
            q.y = ty * TC_ROWS + lq; q.x = tx * VALID + lane;
            q.valid = lane < VALID && q.y < p.H && q.x < p.W;
            q.ly = q.lx = 0.f;
            if (!q.valid) return;
            const long long cimg = f.center_map != nullptr ? (long long)__ldg(f.center_map + q.n) : (long long)q.n * f.frames + f.center;
            int y0 = q.y, x0 = q.x, y1 = q.y, x1 = q.x;
            if (f.scale != 1) {
                const float inv = 1.f / f.scale;
                const float sy = fmaxf(inv * (q.y + 0.5f) - 0.5f, 0.f), sx = fmaxf(inv * (q.x + 0.5f) - 0.5f, 0.f);
                y0 = (int)sy; x0 = (int)sx;
                y1 = y0 + (y0 < Hl - 1 ? 1 : 0); x1 = x0 + (x0 < Wl - 1 ? 1 : 0);
                q.ly = sy - y0; q.lx = sx - x0;
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int co = co0 + k;
                if (co >= co1) break;
                const long long pc = (cimg * nc + co) * lplane;
                if (f.x_dtype == RVSR_F32) {
                    const uint32_t *xp = reinterpret_cast<const uint32_t *>(f.x) + pc;
                    q.c[k][0] = __ldg(xp + y0 * Wl + x0); q.c[k][1] = __ldg(xp + y0 * Wl + x1);
                    q.c[k][2] = __ldg(xp + y1 * Wl + x0); q.c[k][3] = __ldg(xp + y1 * Wl + x1);
                } else {
                    const unsigned short *xp = reinterpret_cast<const unsigned short *>(f.x) + pc;
                    q.c[k][0] = __ldg(xp + y0 * Wl + x0); q.c[k][1] = __ldg(xp + y0 * Wl + x1);
                    q.c[k][2] = __ldg(xp + y1 * Wl + x0); q.c[k][3] = __ldg(xp + y1 * Wl + x1);
                }
            }
        };
        // per tile: `cur` was fetched an iteration ago, `nxt` is fetched for the following tile; the two buffers swap
        // roles by unrolling -- copying nxt into cur would wait for the loads that were just issued.
        Pix px[2];
        const int stride = EG * (int)gridDim.x;
        int tile0 = blockIdx.x + eg * gridDim.x;
        uint32_t t0 = (uint32_t)eg;
        if (tile0 < p.num_tiles) fetch(tile0, px[0]);
        for (; tile0 < p.num_tiles; tile0 += 2 * stride, t0 += 2 * EG) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {  // fully unrolled: px[u] / px[u ^ 1] stay in registers
            const int tile = tile0 + u * stride;
            if (tile >= p.num_tiles) break;
            const uint32_t t = t0 + (uint32_t)u * EG;
            Pix &cur = px[u], &nxt = px[u ^ 1];
            const uint32_t buf = t % NB;
            mbar_wait(TFULL(buf), (t / NB) & 1);
            tc_fence_after();
            if (reader && !(p.debug & 8)) {
                uint32_t r0[16], r1[16];
                const uint32_t taddr = tmem_base + buf * 64 + (uint32_t)hv * 32 + ((uint32_t)(lq * 32) << 16);
                tmem_ld16_nowait(taddr, r0);
                tmem_ld16_nowait(taddr + 16, r1);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 27; ++c)
                    if (c < ncol) xg[(c * 6 + row) * 32 + lane] = __uint_as_float(c < 16 ? r0[c & 15] : r1[c & 15]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(TEMPTY(buf));
            asm volatile("bar.sync %0, %1;" ::"r"(2 + eg), "n"(32 * WPG) : "memory");  // partial sums of the tile are in xg
            if (cur.valid && !(p.debug & 2)) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int co = co0 + k;
                    if (co >= co1) break;
                    float acc = bias_s[co];
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap)
                        acc += xg[((tap * nc + co) * 6 + lq + tap / 3) * 32 + lane + tap % 3];
                    float c4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        c4[i] = f.x_dtype == RVSR_F32 ? __uint_as_float(cur.c[k][i]) : __half2float(__ushort_as_half((unsigned short)cur.c[k][i]));
                    const float base = f.scale == 1 ? c4[0] : (1.f - cur.ly) * ((1.f - cur.lx) * c4[0] + cur.lx * c4[1]) +
                                                                  cur.ly * ((1.f - cur.lx) * c4[2] + cur.lx * c4[3]);
                    const long long oi = ((long long)cur.n * nc + co) * hplane + (long long)cur.y * p.W + cur.x;
                    if (f.out_dtype == RVSR_F32) reinterpret_cast<float *>(p.out)[oi] = acc + base;
                    else reinterpret_cast<__half *>(p.out)[oi] = __float2half_rn(acc + base);
                }
            }
            const int next = tile + EG * gridDim.x;
            if (next < p.num_tiles) fetch(next, nxt);
            asm volatile("bar.sync %0, %1;" ::"r"(2 + eg), "n"(32 * WPG) : "memory");  // xg may be overwritten by the next tile
          }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TAPN_TMEM_COLS);
}

size_t tc_tapn_weight_bytes(int Cout, int Cin, int ks) {
    return (ks == 3 && Cout >= 1 && Cout <= 3 && Cin % 16 == 0 && Cin >= 16 && Cin <= 64) ? (size_t)(Cin / 8) * 32 * 16 : 0;
}
__global__ void pack_weight_tapn_kernel(const float *__restrict__ w, __half *__restrict__ dst, int Cout, int Cin, int total) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int e = i % 8, n = (i / 8) % 32, q = i / 256;
        const int tap = n / Cout, co = n % Cout, cin = q * 8 + e;
        dst[i] = __float2half_rn(n < 9 * Cout ? w[((long long)co * Cin + cin) * 9 + tap] : 0.f);
    }
}
int pack_weight_tapn(const float *w_oihw, void *dst, int Cout, int Cin, cudaStream_t s) {
    RVSR_CHECK_ARG(tc_tapn_weight_bytes(Cout, Cin, 3) > 0, "tapn pack: unsupported shape %d <- %d", Cout, Cin);
    const int total = (Cin / 8) * 32 * 8;
    pack_weight_tapn_kernel<<<(total + 255) / 256, 256, 0, s>>>(w_oihw, reinterpret_cast<__half *>(dst), Cout, Cin, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

// op: the conv_last ConvOp (one source, OUT_FINAL, fin filled in); w_tapn from pack_weight_tapn
int launch_conv_tapn(const ConvOp &op, const void *w_tapn, cudaStream_t s) {
    const Src &sr = op.src[0];
    RVSR_CHECK_ARG(op.nsrc == 1 && op.ks == 3 && op.stride == 1 && op.out_mode == OUT_FINAL && w_tapn != nullptr &&
                       tc_tapn_weight_bytes(op.Cout, sr.C, 3) > 0 && op.fin.nc == op.Cout && op.fin.x != nullptr && op.fin.scale >= 1 &&
                       op.H % op.fin.scale == 0 && op.W % op.fin.scale == 0 && sr.map == nullptr && sr.fixed_frame < 0,
                   "tapn conv: unsupported configuration");
    EncodeTiledFn enc = get_encode();
    RVSR_CHECK_ARG(enc != nullptr, "tapn conv: cuTensorMapEncodeTiled unavailable");
    TcTapnParams p;
    memset(&p, 0, sizeof(p));
    p.C8s = sr.C / 8;
    const long long plane = (long long)op.H * op.W * 8;
    RVSR_CHECK_ARG(sr.image_stride == plane * p.C8s, "tapn conv: source must be densely packed");
    const cuuint64_t dims[3] = {(cuuint64_t)op.W * 8, (cuuint64_t)op.H, (cuuint64_t)op.N * p.C8s};
    const cuuint64_t strides[2] = {(cuuint64_t)op.W * 16, (cuuint64_t)op.H * op.W * 16};
    const cuuint32_t box[3] = {(cuuint32_t)TC_TW * 8, 6, (cuuint32_t)p.C8s};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(sr.ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("tapn conv: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return RVSR_E_CUDA;
    }
    p.w = reinterpret_cast<const __half *>(w_tapn); p.bias = op.bias; p.out = op.out; p.fin = op.fin;
    p.N = op.N; p.H = op.H; p.W = op.W; p.nc = op.Cout;
    static const int dbg = getenv("RVSR_TC_DEBUG") ? atoi(getenv("RVSR_TC_DEBUG")) : 0;
    p.debug = dbg;
    p.tiles_x = cdiv(op.W, TC_TW - 2); p.tiles_y = cdiv(op.H, TC_ROWS); p.num_tiles = p.tiles_x * p.tiles_y * op.N;
    if (p.num_tiles == 0) return RVSR_OK;
    p.td.tpi = (uint32_t)(p.tiles_x * p.tiles_y); p.td.m_tpi = magic_div(p.td.tpi, (uint32_t)p.num_tiles);
    p.td.tx = (uint32_t)p.tiles_x; p.td.m_tx = magic_div(p.td.tx, p.td.tpi);
    const size_t stage = (size_t)p.C8s * 6 * TC_TW * 16;
    const size_t fixed = (size_t)p.C8s * 32 * 16 + 128 + TAPN_EG * TAPN_XCH * sizeof(float) + 8 * sizeof(float) + 512;
    int st = (int)((TC_SMEM_LIMIT - fixed) / stage);
    if (st > 6) st = 6;
    RVSR_CHECK_ARG(st >= 2, "tapn conv: not enough shared memory");
    p.nstages = st;
    const size_t smem = fixed + (size_t)st * stage + 1024;
    RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&conv_tapn_kernel), (int)TC_SMEM_LIMIT + 1024));
    int gx = sm_count();
    if (gx > p.num_tiles) gx = p.num_tiles;
    launch_k(conv_tapn_kernel, dim3(gx), dim3(TAPN_THREADS), smem, s, p);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

}  // namespace rvsr
