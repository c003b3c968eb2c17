// train_kernels.cu -- the bf16 TRAINING path on channel-blocked tensors (BASELINE cfg5: forward + backward of the plain
// convolutions of EDVR_arch.py on this library's own kernels instead of cuDNN).
//
// Tensors are "C8": [N][C/8][H][W][8] bfloat16, the layout of the inference engine, held by PyTorch as ordinary 5-D tensors
// (realvsr_b200/train_c8.py) so that torch.autograd orchestrates and every Function below is one launch of:
//   forward / data gradient   conv_tc2_kernel / conv_tc_kernel (tc_kernels.cu) with bf16 operands; the data gradient of a
//                             stride-1 convolution IS a convolution with the transposed, flipped weights (WeightView)
//   weight (+ bias) gradient  conv_wgrad_tc_kernel (here): dW[tap][ci][co] = sum_pixels x[pixel + tap][ci] * g[pixel][co], a GEMM
//                             whose K dimension is the PIXEL axis, on tcgen05 with both operands MN-major
//   activation gradient, pixel-unshuffle, x2 bilinear upsample forward / adjoint, NCHW <-> C8: small CUDA-core kernels (HBM-bound)
// Reference semantics: nn.Conv2d autograd at EDVR_arch.py:71-91, :229-253, arch_util.py:121-139; F.interpolate(scale_factor=2,
// mode='bilinear', align_corners=False) at EDVR_arch.py:109-121; nn.PixelShuffle(2) at :313-314.
#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace rvsr {

size_t dcn_bwd_tc_wt_bytes();
int pack_wt_dcn_bwd_tc(const void *weight_bf16, void *wt, cudaStream_t s);
int launch_dcn_bwd_tc_core(const void *x8, const void *g8, const float *off32, const float *msk32, const void *wt, float *gx8, float *goff32,
                           float *gmsk32, float *gw32, int B, int H, int W, cudaStream_t s, void *gom_c8);
int launch_act_bwd_c8(const void *g, const void *y, void *out, long long n_elems, int act, cudaStream_t s);

namespace {

// ---------------------------------------------------------------- weight gradient on tcgen05
// Per 128-pixel tile (4 rows x 32 columns) and K step (16 consecutive pixels of one row):
//   B operand  = g tile   [8 co blocks][4 rows][32 px][8]   (TMA, dense box)            N = 64 output channels
//   A operand  = x halo   [7 rows][8 channel blocks][34 px][8], ROW-major on purpose: the 16 MN chunks of an M = 128 operand sit one
//                block stride (544 B) apart, so chunks 0-7 are the 8 channel blocks of halo row r and chunks 8-15 THE SAME blocks of
//                row r + 1 -- one instruction covers taps (dy, dx) [rows 0-63] and (dy + 1, dx) [rows 64-127] of all 64 input
//                channels from ONE copy of the halo.  A tap is a shifted start address, exactly as in the forward kernels.  The halo is 34 pixels
//                wide (K steps must not wrap at dx = +1), which no dense TMA box delivers (<= 256 elements per dimension), so
//                sixteen producer warps assemble it with 16-byte loads (zero fill outside the image).
//   D          = six 128 x 64 fp32 accumulators in TMEM, resident over ALL tiles of the CTA: (dy -1|0, dx -1), (.., dx 0),
//                (.., dx +1), then (dy +1|unused, dx -1..+1): 6 instructions per K step for 9 taps (75 % useful rows).
// One red.global.add.v4.f32 pass per CTA at the end into dW^T [9][64][Cout] fp32 (zeroed by the caller).  The bias gradient
// (column sums of g) is taken from the g tiles in shared memory by two otherwise idle warps.
constexpr int WG_PROD_WARPS = 16, WG_WARP_TMA = 16, WG_WARP_MMA = 17, WG_WARP_BIAS0 = 18, WG_BIAS_WARPS = 2;
constexpr int WG_AHEAD = 4;  // register sets of a producer thread = halo tiles whose loads are in flight ahead of the shared-memory ring
constexpr int WG_THREADS = 32 * (WG_WARP_BIAS0 + WG_BIAS_WARPS);
constexpr int WG_XCOLS = 34, WG_XROWS = 7;           // halo: rows -1 .. 5, columns -1 .. 32
constexpr int WG_XBLOCK = WG_XCOLS * 16;             // 544 B: one channel block of one row
constexpr int WG_XROW = 8 * WG_XBLOCK;               // 4352 B: one row, [8 channel blocks][34 px][8]
constexpr int WG_X_BYTES = WG_XROWS * WG_XROW;       // 30464
constexpr int WG_G_BYTES = 8 * 128 * 16;             // 16384
constexpr int WG_STAGE = WG_G_BYTES + WG_X_BYTES;    // 46848 (a multiple of 128: the TMA destination stays aligned)
constexpr int WG_STAGES = 4;
constexpr int WG_HALO_ELEMS = 8 * WG_XROWS * WG_XCOLS;  // 16-byte cells loaded per tile
constexpr int WG_PER_THREAD = (WG_HALO_ELEMS + 32 * WG_PROD_WARPS - 1) / (32 * WG_PROD_WARPS);

constexpr int WG_MAX_JOBS = 8;
// One launch runs up to WG_MAX_JOBS independent (x, g) pairs of the same geometry (blockIdx.z = job): the sources of a torch.cat
// convolution (same g), the two convolutions of a fused pair -- each job costs a launch's fixed ~10 us otherwise.
struct alignas(64) WgradParams {
    CUtensorMap tmap_g[WG_MAX_JOBS];        // [W * 8, H, N * g_planes] bf16, box {256, 4, 8}
    const uint4 *x[WG_MAX_JOBS];            // C8 bf16: 16-byte cells [N][planes][H][W]
    long long x_image_stride[WG_MAX_JOBS];  // cells between images
    int bias_jobs;             // bit j: job j also sums the bias gradient
    int g_planes;              // channel blocks per image of g
    float *partial;            // [jobs][passes][CTAs][6][128][64] fp32: every CTA's accumulators (summed by wgrad_reduce_kernel)
    float *bias_partial;       // [jobs][passes][CTAs][64] fp32
    int Cout, N, H, W;
    int tiles_x, tiles_y, num_tiles;
    TileDiv td;
    int ks;     // 3, or 1: only the centre tap (rows 64-127 of accumulator 1) is a weight
    int debug;  // RVSR_WG_DEBUG, timing experiments only (results become garbage): 1 = issue no MMAs, 2 = no halo loads / stores
};

// kind::f16 with bf16 A / B, fp32 D, both operands MN-major
__host__ __device__ constexpr uint32_t wg_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *stage_s = smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage_s + WG_STAGES * WG_STAGE);
    constexpr int B_FULL = 0, B_EMPTY = WG_STAGES, B_DONE = 2 * WG_STAGES, B_COUNT = B_DONE + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + B_COUNT);
    float *bias_s = reinterpret_cast<float *>(tmem_slot + 4);  // 2 x 64 floats
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int pass = blockIdx.y, job = blockIdx.z;
    const size_t cta = ((size_t)job * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;  // this CTA's slot in partial / bias_partial
    const bool want_bias = (p.bias_jobs >> job) & 1;
    const int T = blockIdx.x < (unsigned)p.num_tiles ? (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_STAGES; ++i) {
            mbar_init(BAR(B_FULL + i), WG_PROD_WARPS + 1);   // producer warps + the TMA thread's expect_tx arrival
            mbar_init(BAR(B_EMPTY + i), 1 + WG_BIAS_WARPS);  // tcgen05.commit + the bias warps
        }
        mbar_init(BAR(B_DONE), 1);
        fence_barrier_init();
        prefetch_tensormap(&p.tmap_g[job]);
    }
    if (warp == WG_WARP_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);

    if (warp < WG_PROD_WARPS) {
        // ---- x halo producers: rows -1 .. 5 of the tile, 34 columns from -1, all 8 channel blocks.
        // The loads of the next WG_AHEAD - 1 tiles are in flight while a tile is written to shared memory (WG_AHEAD register sets).
        const long long plane = (long long)p.H * p.W;
        int cell[WG_PER_THREAD];  // this thread's cells of the halo: (block, row, column) -> offsets, fixed for the whole kernel
        int off_s[WG_PER_THREAD], rowi[WG_PER_THREAD], coli[WG_PER_THREAD];
#pragma unroll
        for (int i = 0; i < WG_PER_THREAD; ++i) {
            const int idx = (int)threadIdx.x + i * 32 * WG_PROD_WARPS;
            const int pl = idx / (WG_XROWS * WG_XCOLS), rem = idx - pl * (WG_XROWS * WG_XCOLS);
            const int row = rem / WG_XCOLS, col = rem - row * WG_XCOLS;
            cell[i] = idx < WG_HALO_ELEMS ? pl : -1;
            rowi[i] = row; coli[i] = col;
            off_s[i] = (row * 8 + pl) * WG_XCOLS + col;  // 16-byte units in the stage's halo: [row][block][column]
        }
        auto load_tile = [&](int t, uint4 (&v)[WG_PER_THREAD]) {
            const int tile = blockIdx.x + t * gridDim.x;
            int tx, ty, n;
            tile_coords(p.td, tile, tx, ty, n);
            const int y0 = ty * TC_ROWS - 1, x0 = tx * TC_TW - 1;
            const uint4 *src = p.x[job] + (long long)n * p.x_image_stride[job];
#pragma unroll
            for (int i = 0; i < WG_PER_THREAD; ++i) {
                const int gy = y0 + rowi[i], gx = x0 + coli[i];
                v[i] = make_uint4(0, 0, 0, 0);
                if (cell[i] >= 0 && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W && !(p.debug & 2))
                    v[i] = __ldg(src + cell[i] * plane + (long long)gy * p.W + gx);
            }
        };
        auto store_tile = [&](int t, const uint4 (&v)[WG_PER_THREAD]) {
            const int st = t % WG_STAGES;
            mbar_wait(BAR(B_EMPTY + st), ((t / WG_STAGES) & 1) ^ 1);
            uint4 *xa = reinterpret_cast<uint4 *>(stage_s + st * WG_STAGE + WG_G_BYTES);
#pragma unroll
            for (int i = 0; i < WG_PER_THREAD; ++i) {
                if (cell[i] >= 0 && !(p.debug & 2)) xa[off_s[i]] = v[i];
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_FULL + st));
        };
        // WG_AHEAD tiles of loads in flight per thread: with few tiles per CTA (the 16-image layers) the kernel is bound by
        // the latency of these loads, not by their bandwidth
        uint4 v[WG_AHEAD][WG_PER_THREAD];
#pragma unroll
        for (int k = 0; k < WG_AHEAD - 1; ++k)
            if (k < T) load_tile(k, v[k]);
        for (int t0 = 0; t0 < T; t0 += WG_AHEAD) {
#pragma unroll
            for (int k = 0; k < WG_AHEAD; ++k) {
                const int t = t0 + k;
                if (t < T) {
                    if (t + WG_AHEAD - 1 < T) load_tile(t + WG_AHEAD - 1, v[(k + WG_AHEAD - 1) % WG_AHEAD]);
                    store_tile(t, v[k]);
                }
            }
        }
        // ---- flush: D row m = (tap half, ci), column = co, to this CTA's slice of `partial` (plain stores: 148 CTAs adding into the
        // same 36 864 addresses with red.global cost ~25 us per launch).  Warps 0-7: TMEM lane quarter warp & 3, column half warp >> 2.
        if (T > 0 && warp < 8) {
            mbar_wait(BAR(B_DONE), 0);
            tc_fence_after();
            const int q = warp & 3, hf = warp >> 2;
            const int m = q * 32 + lane;
            float *part = p.partial + (cta * 6 * 128 + m) * 64 + hf * 32;
#pragma unroll 1
            for (int b = 0; b < 6; ++b) {
                if (b >= 3 && q >= 2) break;  // rows 64-127 of the dy = +1 accumulators are not weights
                if (p.ks == 1 && (b != 1 || q < 2)) continue;  // 1x1: the centre tap only
                uint32_t a0[16], a1[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 64 + hf * 32);
                tmem_ld16_nowait(taddr, a0);
                tmem_ld16_nowait(taddr + 16, a1);
                tmem_ld_wait();
                float4 *d = reinterpret_cast<float4 *>(part + (size_t)b * 128 * 64);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    d[j] = make_float4(__uint_as_float(a0[4 * j]), __uint_as_float(a0[4 * j + 1]), __uint_as_float(a0[4 * j + 2]), __uint_as_float(a0[4 * j + 3]));
                    d[4 + j] = make_float4(__uint_as_float(a1[4 * j]), __uint_as_float(a1[4 * j + 1]), __uint_as_float(a1[4 * j + 2]), __uint_as_float(a1[4 * j + 3]));
                }
            }
        }
    } else if (warp == WG_WARP_TMA) {
        if (lane == 0) {
            for (int t = 0; t < T; ++t) {
                const int tile = blockIdx.x + t * gridDim.x, st = t % WG_STAGES;
                int tx, ty, n;
                tile_coords(p.td, tile, tx, ty, n);
                mbar_wait(BAR(B_EMPTY + st), ((t / WG_STAGES) & 1) ^ 1);
                mbar_expect_tx(BAR(B_FULL + st), WG_G_BYTES);
                tma_load_3d(smem_u32(stage_s + st * WG_STAGE), &p.tmap_g[job], BAR(B_FULL + st), tx * TC_TW * 8, ty * TC_ROWS,
                            n * p.g_planes + pass * 8);
            }
        }
    } else if (warp == WG_WARP_MMA) {
        constexpr uint32_t idesc = wg_idesc(128, 64);
        for (int t = 0; t < T; ++t) {
            const int st = t % WG_STAGES;
            mbar_wait(BAR(B_FULL + st), (t / WG_STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t g0 = smem_u32(stage_s + st * WG_STAGE), x0 = g0 + WG_G_BYTES;
#pragma unroll
                for (int r = 0; r < TC_ROWS; ++r)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint64_t bdesc = make_desc(g0 + (uint32_t)(r * TC_TW + h * 16) * 16, 128, 2048);
#pragma unroll
                        for (int b = 0; b < 6; ++b) {
                            if ((p.ks == 1 && b != 1) || (p.debug & 1)) continue;
                            const uint32_t a = x0 + (uint32_t)(r + (b >= 3 ? 2 : 0)) * WG_XROW + (uint32_t)(h * 16 + (b % 3)) * 16;
                            umma_f16(tmem_base + (uint32_t)b * 64, make_desc(a, 128, WG_XBLOCK), bdesc, idesc, (t | r | h) ? 1u : 0u);
                        }
                    }
                umma_commit(BAR(B_EMPTY + st));
            }
            __syncwarp();
        }
        if (T > 0 && elect_one()) umma_commit(BAR(B_DONE));
        __syncwarp();
    } else {
        // ---- bias gradient: column sums of the g tiles (pixels outside the image arrive as zeros).  64 threads walk the tile's
        // 1024 16-byte cells with consecutive lanes on consecutive cells (a strided walk made these loads 32-way bank
        // conflicts that saturated the LSU pipe the halo producers share -- ncu, profiles/r02_ncu_conv_wgrad.txt).
        const int j = (warp - WG_WARP_BIAS0) * 32 + lane;  // 0 .. 63
        float acc[8][8];                                    // [channel block][channel]
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[q][k] = 0.f;
        for (int t = 0; t < T; ++t) {
            const int st = t % WG_STAGES;
            mbar_wait_sleep(BAR(B_FULL + st), (t / WG_STAGES) & 1, 200);
            if (want_bias) {
                const uint4 *g = reinterpret_cast<const uint4 *>(stage_s + st * WG_STAGE) + j;
#pragma unroll
                for (int i = 0; i < 16; ++i) {  // cell i * 64 + j: channel block i / 2
                    const uint4 u = g[i * 64];
                    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        acc[i >> 1][2 * k] += __uint_as_float(w[k] << 16);
                        acc[i >> 1][2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_EMPTY + st));
        }
        if (want_bias) {
            float *bp = p.bias_partial + cta * 64;
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float s = acc[q][k];
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (lane == 0) bias_s[(warp - WG_WARP_BIAS0) * 64 + q * 8 + k] = s;
                }
            asm volatile("bar.sync 1, 64;" ::: "memory");  // the two bias warps
            bp[j] = bias_s[j] + bias_s[64 + j];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WG_WARP_MMA) tmem_dealloc(tmem_base, 512);
}

// Sum of the CTAs' accumulators -> the OIHW weight gradient gw[co][c0 + ci][ky][kx] (written, fp32) and db[co].  A block = 64
// groups of 4 consecutive output channels x 4 slices of the CTA range: 16-byte loads, 4 in flight per thread (the partials are
// up to 29 MB, mostly L2-resident: memory-level parallelism is what this kernel needs), then a fixed-order sum of the 4 slices
// (bit-reproducible gradients, which atomics would not give).
struct WgradOut {
    float *gw[WG_MAX_JOBS], *db[WG_MAX_JOBS];
    int cin_total[WG_MAX_JOBS], c0[WG_MAX_JOBS];
};
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ partial_all, const float *__restrict__ bias_partial_all,
                                                           const __grid_constant__ WgradOut out, int ctas, int Cout, int ks) {
    const int job = blockIdx.y, passes = (Cout + 63) / 64;
    const float *partial = partial_all + (size_t)job * passes * ctas * 6 * 128 * 64;
    const float *bias_partial = bias_partial_all + (size_t)job * passes * ctas * 64;
    float *gw = out.gw[job], *db = out.db[job];
    const int cin_total = out.cin_total[job], c0 = out.c0[job];
    const int KK = ks * ks, Cq = Cout / 4;
    const int grp = blockIdx.x * 64 + (threadIdx.x & 63), slice = threadIdx.x >> 6;
    __shared__ float4 sm[4][64];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool live = grp < KK * 64 * Cq;
    int co = 0, ci = 0, tap = 0;
    if (live) {
        co = (grp % Cq) * 4; ci = (grp / Cq) % 64; tap = grp / (Cq * 64);
        int b, m;
        if (ks == 1) { b = 1; m = 64 + ci; }
        else if (tap < 6) { b = tap % 3; m = (tap / 3) * 64 + ci; }
        else { b = 3 + tap % 3; m = ci; }
        const float4 *src = reinterpret_cast<const float4 *>(partial + ((size_t)(co >> 6) * ctas * 6 * 128 + (size_t)b * 128 + m) * 64 + (co & 63));
        const size_t stride = (size_t)6 * 128 * 64 / 4;  // float4s between consecutive CTAs' slices
        const int per = (ctas + 3) / 4, lo = slice * per, hi = min(ctas, lo + per);
        int c = lo;
        for (; c + 3 < hi; c += 4) {
            const float4 v0 = __ldg(src + (size_t)c * stride), v1 = __ldg(src + (size_t)(c + 1) * stride);
            const float4 v2 = __ldg(src + (size_t)(c + 2) * stride), v3 = __ldg(src + (size_t)(c + 3) * stride);
            acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
            acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
        }
        for (; c < hi; ++c) {
            const float4 v = __ldg(src + (size_t)c * stride);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    sm[slice][threadIdx.x & 63] = acc;
    __syncthreads();
    if (live && slice == 0) {
        const float4 a1 = sm[1][threadIdx.x], a2 = sm[2][threadIdx.x], a3 = sm[3][threadIdx.x];
        const float r[4] = {(acc.x + a1.x) + (a2.x + a3.x), (acc.y + a1.y) + (a2.y + a3.y), (acc.z + a1.z) + (a2.z + a3.z), (acc.w + a1.w) + (a2.w + a3.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) gw[((size_t)(co + k) * cin_total + c0 + ci) * KK + tap] = r[k];
    }
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (db != nullptr && idx < Cout) {
        const float *src = bias_partial + (size_t)(idx >> 6) * ctas * 64 + (idx & 63);
        float s_ = 0.f;
        for (int c = 0; c < ctas; ++c) s_ += src[(size_t)c * 64];
        db[idx] = s_;
    }
}

// ---------------------------------------------------------------- small CUDA-core kernels (bf16, 16-byte cells)
__device__ __forceinline__ void unpack8(const uint4 &u, float (&v)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint4 u;
    uint32_t *w = reinterpret_cast<uint32_t *>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t *>(&h);
    }
    return u;
}

// gradient through LeakyReLU(0.1) / ReLU given the layer's OUTPUT y (sign(y) == sign of the pre-activation; at 0 both
// frameworks take the negative branch): out = y > 0 ? g : slope * g
__global__ void act_bwd_c8_kernel(const uint4 *__restrict__ g, const uint4 *__restrict__ y, uint4 *__restrict__ out, long long n, float slope) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gv[8], yv[8];
        unpack8(__ldg(g + i), gv);
        unpack8(__ldg(y + i), yv);
#pragma unroll
        for (int k = 0; k < 8; ++k) gv[k] = yv[k] > 0.f ? gv[k] : slope * gv[k];
        out[i] = pack8(gv);
    }
}

// Gradient through lrelu(PixelShuffle(2)(conv)) (EDVR_arch.py:313-314): g, y are [N][C2/8][2H][2W][8] (C2 = C / 4 channels),
// out is the gradient of the conv output [N][C/8][H][W][8]: out[n, 4c + 2i + j, h, w] = g[n, c, 2h + i, 2w + j] * act'(y[...]).
// A lane owns one high-resolution pixel of one input block (one 16-byte load of g and of y); the four lanes of a 2x2 quad
// (lane = 4 * quad + 2i + j) then hold the 32 values of FOUR output cells (blocks 4P .. 4P + 3 at the low-resolution pixel): lane
// k = 2i + j of the quad assembles block 4P + k from everybody's channels 2k, 2k + 1 with 16 quad shuffles and stores 16 bytes.
// Loads are contiguous per half warp (16 pixels of a row), stores per 8 lanes.
__global__ void unshuffle2_act_bwd_kernel(const uint4 *__restrict__ g, const uint4 *__restrict__ y, uint4 *__restrict__ out, int P2, int H, int W,
                                          long long total_quads, float slope, int has_act) {
    const int lane = threadIdx.x & 31, k = lane & 3;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int Wq = (W + 7) / 8;  // a warp covers 8 low-resolution pixels of one row
    for (long long wq = warp0; wq * 8 < total_quads; wq += nwarps) {
        const int wx = (int)(wq % Wq);
        long long r = wq / Wq;
        const int h = (int)(r % H);
        r /= H;
        const int P = (int)(r % P2);
        const long long n = r / P2;
        const int w = wx * 8 + (lane >> 2);
        const bool ok = w < W;
        uint32_t v[4] = {0u, 0u, 0u, 0u};
        if (ok) {
            const long long src = ((n * P2 + P) * (2 * H) + (2 * h + (k >> 1))) * (long long)(2 * W) + (2 * w + (k & 1));
            const uint4 gv = __ldg(g + src);
            v[0] = gv.x; v[1] = gv.y; v[2] = gv.z; v[3] = gv.w;
            if (has_act) {
                const uint4 yv = __ldg(y + src);
                const uint32_t yy[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {  // two bf16 channels per word
                    const float g0 = __uint_as_float(v[q] << 16), g1 = __uint_as_float(v[q] & 0xffff0000u);
                    const float y0 = __uint_as_float(yy[q] << 16), y1 = __uint_as_float(yy[q] & 0xffff0000u);
                    const __nv_bfloat162 o = __floats2bfloat162_rn(y0 > 0.f ? g0 : slope * g0, y1 > 0.f ? g1 : slope * g1);
                    v[q] = *reinterpret_cast<const uint32_t *>(&o);
                }
            }
        }
        // word q of a lane = its channels (2q, 2q + 1) of block P, i.e. output block 4P + q, elements 4 * {0, 1} + (its k)
        uint32_t got[4];
#pragma unroll
        for (int src_k = 0; src_k < 4; ++src_k) {
            uint32_t t = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t s_ = __shfl_sync(0xffffffffu, v[q], (lane & ~3) | src_k);
                if (q == k) t = s_;
            }
            got[src_k] = t;  // lane src_k's channels (2k, 2k + 1)
        }
        if (ok) {
            // output cell of block 4P + k: element e = 4 * (channel parity) + sub-pixel: [c=2k: k'=0..3][c=2k+1: k'=0..3]
            uint4 o;
            o.x = (got[0] & 0xffffu) | (got[1] << 16);
            o.y = (got[2] & 0xffffu) | (got[3] << 16);
            o.z = (got[0] >> 16) | (got[1] & 0xffff0000u);
            o.w = (got[2] >> 16) | (got[3] & 0xffff0000u);
            out[((n * (4 * P2) + 4 * P + k) * H + h) * (long long)W + w] = o;
        }
    }
}

// F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) on [planes][H][W][8], times `scale`
__global__ void upsample2x_c8_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, int H, int W, long long total, float scale) {
    const int Ho = 2 * H, Wo = 2 * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wo);
        long long r = i / Wo;
        const int oy = (int)(r % Ho);
        const long long pl = r / Ho;
        const float sy = fmaxf(0.5f * (oy + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.5f * (ox + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)sy, x0 = (int)sx, y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
        const float ly = sy - y0, lx = sx - x0;
        const uint4 *p = src + pl * H * W;
        float a[8], b[8], c[8], d[8], o[8];
        unpack8(__ldg(p + (long long)y0 * W + x0), a);
        unpack8(__ldg(p + (long long)y0 * W + x1), b);
        unpack8(__ldg(p + (long long)y1 * W + x0), c);
        unpack8(__ldg(p + (long long)y1 * W + x1), d);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = scale * ((1.f - ly) * ((1.f - lx) * a[k] + lx * b[k]) + ly * ((1.f - lx) * c[k] + lx * d[k]));
        dst[i] = pack8(o);
    }
}
// its adjoint: gin[y][x] = scale * sum over the outputs whose two taps per axis include (y, x).  Output rows 2y-1 .. 2y+2 can
// reference input row y; each candidate recomputes its (y0, y1, ly) exactly as the forward does, so borders are exact.
__global__ void upsample2x_bwd_c8_kernel(const uint4 *__restrict__ g, uint4 *__restrict__ gin, int H, int W, long long total, float scale) {
    const int Ho = 2 * H, Wo = 2 * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        long long r = i / W;
        const int y = (int)(r % H);
        const long long pl = r / H;
        const uint4 *p = g + pl * Ho * Wo;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float wy[4], wx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int oy = 2 * y - 1 + k, ox = 2 * x - 1 + k;
            wy[k] = 0.f; wx[k] = 0.f;
            if (oy >= 0 && oy < Ho) {
                const float sy = fmaxf(0.5f * (oy + 0.5f) - 0.5f, 0.f);
                const int y0 = (int)sy, y1 = y0 + (y0 < H - 1 ? 1 : 0);
                const float ly = sy - y0;
                wy[k] = (y0 == y ? 1.f - ly : 0.f) + (y1 == y ? ly : 0.f);
            }
            if (ox >= 0 && ox < Wo) {
                const float sx = fmaxf(0.5f * (ox + 0.5f) - 0.5f, 0.f);
                const int x0 = (int)sx, x1 = x0 + (x0 < W - 1 ? 1 : 0);
                const float lx = sx - x0;
                wx[k] = (x0 == x ? 1.f - lx : 0.f) + (x1 == x ? lx : 0.f);
            }
        }
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            if (wy[ky] == 0.f) continue;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                if (wx[kx] == 0.f) continue;
                float v[8];
                unpack8(__ldg(p + (long long)(2 * y - 1 + ky) * Wo + (2 * x - 1 + kx)), v);
                const float w = wy[ky] * wx[kx];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += w * v[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= scale;
        gin[i] = pack8(acc);
    }
}

// NCHW (bf16 or fp32) -> C8 bf16 (channels beyond C are zero) and back; one thread per (pixel, channel block)
template <typename Tin>
__global__ void nchw_to_c8_kernel(const Tin *__restrict__ src, uint4 *__restrict__ dst, int C, int HW, int C8) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const int q = blockIdx.y;  // C8: channel blocks per image of the C8 tensor (>= gridDim.y)
    const long long n = blockIdx.z;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = q * 8 + e;
        v[e] = c < C ? to_f<Tin>(src[(n * C + c) * (long long)HW + i]) : 0.f;
    }
    dst[(n * C8 + q) * (long long)HW + i] = pack8(v);
}
template <typename Tout>
__global__ void c8_to_nchw_kernel(const uint4 *__restrict__ src, Tout *__restrict__ dst, int C, int HW, int C8) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const int q = blockIdx.y;
    const long long n = blockIdx.z;
    float v[8];
    unpack8(__ldg(src + (n * C8 + q) * (long long)HW + i), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = q * 8 + e;
        if (c < C) {
            if constexpr (sizeof(Tout) == 4) dst[(n * C + c) * (long long)HW + i] = v[e];
            else dst[(n * C + c) * (long long)HW + i] = __float2bfloat16_rn(v[e]);
        }
    }
}

// ---- ModulatedDeformConvPack on C8 tensors: layout glue around dcn_tc_kernel / dcn_bwd_tc_kernel
// om: the 64 -> 216 offset / mask convolution's output as a 256-channel C8 bf16 tensor (channels: 144 offsets g * 18 + 2 * tap +
// {dy, dx}, 72 mask logits 144 + g * 9 + tap, 40 zeros -- deform_conv.py:279-283).  One thread per (image, pixel) writes, group by group,
//   om24  the forward kernel's format (see om24_from_planar_kernel), mask = sigmoid(logit), and / or
//   off32 / msk32  planar fp32 [N][144][HW] / [N][72][HW] for the backward kernel.
__global__ void om_from_c8_kernel(const uint4 *__restrict__ om, uint4 *__restrict__ om24, float *__restrict__ off32,
                                  float *__restrict__ msk32, int HW) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const long long n = blockIdx.y;
    const uint4 *src = om + n * 32 * (long long)HW + pix;
    // One thread per pixel walks the 8 groups; every index below is a compile-time constant after unrolling, so a group's 27
    // values come out of at most six 16-byte loads (scalar 2-byte loads of the same cells were 3x slower).
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        float o[18], m[10];
        {
            const int p0 = (g * 18) / 8, p1 = (g * 18 + 17) / 8;
            float v[4][8];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (p0 + q <= p1) unpack8(__ldg(src + (long long)(p0 + q) * HW), v[q]);
#pragma unroll
            for (int j = 0; j < 18; ++j) o[j] = v[(g * 18 + j) / 8 - p0][(g * 18 + j) % 8];
            const int m0 = (144 + g * 9) / 8, m1 = (144 + g * 9 + 8) / 8;
            float u[2][8];
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (m0 + q <= m1) unpack8(__ldg(src + (long long)(m0 + q) * HW), u[q]);
#pragma unroll
            for (int j = 0; j < 9; ++j) m[j] = __fdividef(1.f, 1.f + __expf(-u[(144 + g * 9 + j) / 8 - m0][(144 + g * 9 + j) % 8]));
            m[9] = 0.f;
        }
        if (om24 != nullptr) {
            uint32_t mw[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const __half2 h = __floats2half2_rn(m[2 * k], m[2 * k + 1]);
                mw[k] = *reinterpret_cast<const uint32_t *>(&h);
            }
            uint4 *d = om24 + ((n * 24 + g * 3) * (long long)HW + pix) * 2;
            auto f = [](float v) { return __float_as_uint(v); };
            d[0] = make_uint4(f(o[0]), f(o[1]), f(o[2]), f(o[3])); d[1] = make_uint4(f(o[4]), f(o[5]), f(o[6]), f(o[7]));
            d += (long long)HW * 2;
            d[0] = make_uint4(f(o[8]), f(o[9]), f(o[10]), f(o[11])); d[1] = make_uint4(f(o[12]), f(o[13]), f(o[14]), f(o[15]));
            d += (long long)HW * 2;
            d[0] = make_uint4(f(o[16]), f(o[17]), mw[0], mw[1]); d[1] = make_uint4(mw[2], mw[3], mw[4], 0u);
        }
        if (off32 != nullptr) {
            float *po = off32 + (n * 144 + g * 18) * (long long)HW + pix, *pm = msk32 + (n * 72 + g * 9) * (long long)HW + pix;
#pragma unroll
            for (int j = 0; j < 18; ++j) po[(long long)j * HW] = o[j];
#pragma unroll
            for (int j = 0; j < 9; ++j) pm[(long long)j * HW] = m[j];
        }
    }
}
// gradient of the 256-channel convolution output from the operator's planar fp32 gradients: offsets as they are, mask logits
// through the sigmoid (s * (1 - s) with s = msk32), padding channels zero.  One thread per (image, channel block, pixel).
__global__ void om_grad_to_c8_kernel(const float *__restrict__ goff32, const float *__restrict__ gmsk32, const float *__restrict__ msk32,
                                     uint4 *__restrict__ gom, int HW) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const int q = blockIdx.y;
    const long long n = blockIdx.z;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = q * 8 + e;
        if (c < 144) {
            v[e] = goff32[(n * 144 + c) * (long long)HW + pix];
        } else if (c < 216) {
            const long long i = (n * 72 + (c - 144)) * (long long)HW + pix;
            const float sg = msk32[i];
            v[e] = gmsk32[i] * sg * (1.f - sg);
        } else {
            v[e] = 0.f;
        }
    }
    gom[(n * 32 + q) * (long long)HW + pix] = pack8(v);
}
// db[c] += sum over images and pixels of a 64-channel C8 tensor
__global__ void bias_grad_c8_kernel(const uint4 *__restrict__ g, float *__restrict__ db, int HW, int N) {
    const int q = blockIdx.y;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int n = blockIdx.z; n < N; n += gridDim.z)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
            float v[8];
            unpack8(__ldg(g + ((long long)n * 8 + q) * HW + i), v);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += v[k];
        }
    __shared__ float part[8][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float s_ = acc[k];
        for (int o = 16; o > 0; o >>= 1) s_ += __shfl_xor_sync(0xffffffffu, s_, o);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][k] = s_;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w][threadIdx.x];
        atomicAdd(db + q * 8 + threadIdx.x, t);
    }
}
// bf16 <-> fp16, same layout (dcn_tc_kernel's operands are fp16: more mantissa than bf16, and activations stay far inside its range)
__global__ void cvt_bf16_to_f16_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
        unpack8(__ldg(src + i), v);
        uint4 u;
        __half2 *h = reinterpret_cast<__half2 *>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
        dst[i] = u;
    }
}
__global__ void cvt_f16_to_bf16_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint4 u = __ldg(src + i);
        const __half2 *h = reinterpret_cast<const __half2 *>(&u);
        float v[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(h[k]);
            v[2 * k] = f.x; v[2 * k + 1] = f.y;
        }
        dst[i] = pack8(v);
    }
}

// ---- TSA temporal attention (EDVR_arch.py:170-181) on C8 tensors, 64 channels.  One thread per (batch, frame, pixel):
// prob = sigmoid(sum_c emb[b, n, c] * emb_ref[b, c]); out_n = aligned[b, n] * prob.  `out` is N separate [B][8][HW][8] tensors
// (the N sources of the 1x1 fusion convolutions), prob [B][N][HW] fp32 is kept for the backward.
struct TsaPtrs { void *p[RVSR_MAX_SRC_TC]; };
__global__ void tsa_temporal_c8_kernel(const uint4 *__restrict__ aligned, const uint4 *__restrict__ emb, const uint4 *__restrict__ emb_ref,
                                       TsaPtrs out, float *__restrict__ prob, int N, int HW, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int px = (int)(i % HW);
        const int n = (int)((i / HW) % N);
        const long long b = i / ((long long)HW * N);
        const uint4 *e = emb + ((b * N + n) * 8) * (long long)HW + px, *r = emb_ref + (b * 8) * (long long)HW + px;
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float ev[8], rv[8];
            unpack8(__ldg(e + (long long)q * HW), ev);
            unpack8(__ldg(r + (long long)q * HW), rv);
#pragma unroll
            for (int k = 0; k < 8; ++k) dot += ev[k] * rv[k];
        }
        const float pr = __fdividef(1.f, 1.f + __expf(-dot));
        prob[i] = pr;
        const uint4 *a = aligned + ((b * N + n) * 8) * (long long)HW + px;
        uint4 *o = reinterpret_cast<uint4 *>(out.p[n]) + (b * 8) * (long long)HW + px;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float av[8];
            unpack8(__ldg(a + (long long)q * HW), av);
#pragma unroll
            for (int k = 0; k < 8; ++k) av[k] *= pr;
            o[(long long)q * HW] = pack8(av);
        }
    }
}
// backward: one thread per (batch, pixel) walks the N frames.  g_aligned = g_out * prob; g_dot = (sum_c g_out * aligned) * p (1 - p);
// g_emb = g_dot * emb_ref; g_emb_ref = sum_n g_dot * emb[n]  (no atomics: the frame loop is inside the thread)
__global__ void tsa_temporal_bwd_c8_kernel(TsaPtrs gout, const uint4 *__restrict__ aligned, const uint4 *__restrict__ emb,
                                           const uint4 *__restrict__ emb_ref, const float *__restrict__ prob, uint4 *__restrict__ g_aligned,
                                           uint4 *__restrict__ g_emb, uint4 *__restrict__ g_emb_ref, int N, int HW, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int px = (int)(i % HW);
        const long long b = i / HW;
        const uint4 *r = emb_ref + (b * 8) * (long long)HW + px;
        float acc[8][8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[q][k] = 0.f;
        for (int n = 0; n < N; ++n) {
            const long long base = ((b * N + n) * 8) * (long long)HW + px;
            const float pr = prob[(b * N + n) * (long long)HW + px];
            const uint4 *go = reinterpret_cast<const uint4 *>(gout.p[n]);
            float gdot = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float gv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, av[8];
                if (go != nullptr) unpack8(__ldg(go + (b * 8 + q) * (long long)HW + px), gv);
                unpack8(__ldg(aligned + base + (long long)q * HW), av);
#pragma unroll
                for (int k = 0; k < 8; ++k) { gdot += gv[k] * av[k]; gv[k] *= pr; }
                g_aligned[base + (long long)q * HW] = pack8(gv);
            }
            gdot *= pr * (1.f - pr);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float ev[8], rv[8];
                unpack8(__ldg(emb + base + (long long)q * HW), ev);
                unpack8(__ldg(r + (long long)q * HW), rv);
#pragma unroll
                for (int k = 0; k < 8; ++k) { acc[q][k] += gdot * ev[k]; rv[k] *= gdot; }
                g_emb[base + (long long)q * HW] = pack8(rv);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) g_emb_ref[(b * 8 + q) * (long long)HW + px] = pack8(acc[q]);
    }
}

// ---- TSA spatial attention helpers (EDVR_arch.py:154-155, :186-208) on C8 tensors
// MaxPool2d(3, 2, 1) and AvgPool2d(3, 2, 1) (count_include_pad) of one tensor in one pass; cells = 16-byte (8-channel) units
__global__ void pool_maxavg_c8_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dmax, uint4 *__restrict__ davg, int H, int W,
                                      long long total) {
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wo);
        long long r = i / Wo;
        const int oy = (int)(r % Ho);
        const uint4 *p = src + (r / Ho) * H * W;
        float mx[8], sm[8], v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { mx[k] = -INFINITY; sm[k] = 0.f; }
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int y = 2 * oy + dy, x = 2 * ox + dx;
                if (y < 0 || y >= H || x < 0 || x >= W) continue;
                unpack8(__ldg(p + (long long)y * W + x), v);
#pragma unroll
                for (int k = 0; k < 8; ++k) { mx[k] = fmaxf(mx[k], v[k]); sm[k] += v[k]; }
            }
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[k] *= (1.f / 9.f);
        dmax[i] = pack8(mx);
        davg[i] = pack8(sm);
    }
}
// gradient of both pools w.r.t. their common input: one thread per INPUT cell visits the <= 4 windows that contain it; the max
// pool's gradient goes to the FIRST maximum of a window in row-major order (torch's rule), found by re-scanning the window
__global__ void pool_maxavg_bwd_c8_kernel(const uint4 *__restrict__ src, const uint4 *__restrict__ gmax, const uint4 *__restrict__ gavg,
                                          uint4 *__restrict__ gin, int H, int W, long long total) {
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        long long r = i / W;
        const int y = (int)(r % H);
        const long long pl = r / H;
        const uint4 *p = src + pl * H * W;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int oy = y / 2; oy <= (y + 1) / 2; ++oy) {
            if (oy >= Ho) continue;
            for (int ox = x / 2; ox <= (x + 1) / 2; ++ox) {
                if (ox >= Wo) continue;
                const long long o = (pl * Ho + oy) * Wo + ox;
                float ga[8], gm[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (gavg != nullptr) {
                    unpack8(__ldg(gavg + o), ga);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] += ga[k] * (1.f / 9.f);
                }
                if (gmax == nullptr) continue;
                unpack8(__ldg(gmax + o), gm);
                float best[8];
                int arg[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; arg[k] = -1; }
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int yy = 2 * oy + dy, xx = 2 * ox + dx;
                        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                        float v[8];
                        unpack8(__ldg(p + (long long)yy * W + xx), v);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (v[k] > best[k] || arg[k] < 0) { best[k] = v[k]; arg[k] = yy * W + xx; }
                    }
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (arg[k] == y * W + x) acc[k] += gm[k];
            }
        }
        gin[i] = pack8(acc);
    }
}
// out = fea * sigmoid(att) * 2 + att_add (EDVR_arch.py:206-207) and its gradient (g_att_add = g: no kernel needed)
__global__ void tsa_final_c8_kernel(const uint4 *__restrict__ fea, const uint4 *__restrict__ att, const uint4 *__restrict__ add,
                                    uint4 *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float f[8], a[8], d[8];
        unpack8(__ldg(fea + i), f); unpack8(__ldg(att + i), a); unpack8(__ldg(add + i), d);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = f[k] * (2.f * __fdividef(1.f, 1.f + __expf(-a[k]))) + d[k];
        out[i] = pack8(f);
    }
}
__global__ void tsa_final_bwd_c8_kernel(const uint4 *__restrict__ g, const uint4 *__restrict__ fea, const uint4 *__restrict__ att,
                                        uint4 *__restrict__ g_fea, uint4 *__restrict__ g_att, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gv[8], f[8], a[8], gf[8], ga[8];
        unpack8(__ldg(g + i), gv); unpack8(__ldg(fea + i), f); unpack8(__ldg(att + i), a);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float sg = __fdividef(1.f, 1.f + __expf(-a[k]));
            gf[k] = gv[k] * 2.f * sg;
            ga[k] = gv[k] * f[k] * 2.f * sg * (1.f - sg);
        }
        g_fea[i] = pack8(gf);
        g_att[i] = pack8(ga);
    }
}

// ---- F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) on NCHW planes (module path: fp32 / fp16 / bf16
// training on torch's convolutions).  torch's own NCHW kernel gives one thread an output PIXEL and loops over all images x
// channels inside it: 1024 threads for a [80, 64, 16, 16] tensor, 0.8-1.4 ms per call, 16 % of the fp32 cfg5 step.
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <typename T>
__global__ void upsample2x_nchw_kernel(const T *__restrict__ src, T *__restrict__ dst, int H, int W, long long total, float scale) {
    const int Ho = 2 * H, Wo = 2 * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wo);
        long long r = i / Wo;
        const int oy = (int)(r % Ho);
        const T *p = src + (r / Ho) * H * W;
        const float sy = fmaxf(0.5f * (oy + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.5f * (ox + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)sy, x0 = (int)sx, y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
        const float ly = sy - y0, lx = sx - x0;
        const float a = to_f<T>(p[(long long)y0 * W + x0]), b = to_f<T>(p[(long long)y0 * W + x1]);
        const float c = to_f<T>(p[(long long)y1 * W + x0]), d = to_f<T>(p[(long long)y1 * W + x1]);
        dst[i] = from_f32<T>(scale * ((1.f - ly) * ((1.f - lx) * a + lx * b) + ly * ((1.f - lx) * c + lx * d)));
    }
}
template <typename T>
__global__ void upsample2x_bwd_nchw_kernel(const T *__restrict__ g, T *__restrict__ gin, int H, int W, long long total, float scale) {
    const int Ho = 2 * H, Wo = 2 * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        long long r = i / W;
        const int y = (int)(r % H);
        const T *p = g + (r / H) * Ho * Wo;
        float wy[4], wx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // output rows / columns 2y-1 .. 2y+2 can reference input row / column y (see the C8 adjoint)
            const int oy = 2 * y - 1 + k, ox = 2 * x - 1 + k;
            wy[k] = 0.f; wx[k] = 0.f;
            if (oy >= 0 && oy < Ho) {
                const float sy = fmaxf(0.5f * (oy + 0.5f) - 0.5f, 0.f);
                const int y0 = (int)sy, y1 = y0 + (y0 < H - 1 ? 1 : 0);
                wy[k] = (y0 == y ? 1.f - (sy - y0) : 0.f) + (y1 == y ? sy - y0 : 0.f);
            }
            if (ox >= 0 && ox < Wo) {
                const float sx = fmaxf(0.5f * (ox + 0.5f) - 0.5f, 0.f);
                const int x0 = (int)sx, x1 = x0 + (x0 < W - 1 ? 1 : 0);
                wx[k] = (x0 == x ? 1.f - (sx - x0) : 0.f) + (x1 == x ? sx - x0 : 0.f);
            }
        }
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            if (wy[ky] == 0.f) continue;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx)
                if (wx[kx] != 0.f) acc += wy[ky] * wx[kx] * to_f<T>(p[(long long)(2 * y - 1 + ky) * Wo + (2 * x - 1 + kx)]);
        }
        gin[i] = from_f32<T>(scale * acc);
    }
}

static int ew_grid(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

bool conv_wgrad_tc_supported(int Cin, int ks, int stride) {
    return Cin == 64 && (ks == 3 || ks == 1) && stride == 1 && get_encode() != nullptr;
}

// CTAs per pass.  Measured cost model of one launch (us): 8 + 0.08 * CTAs (each CTA writes 147 KB of partial sums that the
// reduce kernel reads back) + 2 * tiles / CTAs  ->  minimum at CTAs = 5 * sqrt(tiles)
static int wgrad_ctas(int num_tiles, int passes) {
    static const double scale = getenv("RVSR_WG_CTA_SCALE") ? atof(getenv("RVSR_WG_CTA_SCALE")) : 5.0;
    int gx = (int)(scale * sqrt((double)num_tiles));
    const int cap = sm_count() / passes > 0 ? sm_count() / passes : 1;
    if (gx > cap) gx = cap;
    if (gx > num_tiles) gx = num_tiles;
    return gx < 1 ? 1 : gx;
}
size_t conv_wgrad_tc_workspace_bytes(int njobs, int N, int H, int W, int Cout) {
    const int passes = cdiv(Cout, 64), tiles = cdiv(W, TC_TW) * cdiv(H, TC_ROWS) * N;
    const size_t ctas = (size_t)wgrad_ctas(tiles, passes * njobs) * passes * njobs;
    return align_up(ctas * 6 * 128 * 64 * 4, 256) + align_up(ctas * 64 * 4, 256) + 512;
}
// njobs independent weight gradients of one geometry.  Job j: gw[j] (OIHW fp32 [Cout][cin_total[j]][ks][ks]) gets the slice of
// input channels [c0[j], c0[j] + 64) WRITTEN from source x[j] and output gradient g[j]; db[j] ([Cout] fp32, may be null) written.
int launch_conv_wgrad_tc(int njobs, const void *const *x_c8, const long long *x_image_stride, const void *const *g_c8, float *const *gw,
                         float *const *db, const int *cin_total, const int *c0, int N, int H, int W, int Cout, int ks, void *workspace,
                         size_t workspace_bytes, cudaStream_t s) {
    RVSR_CHECK_ARG(njobs >= 1 && njobs <= WG_MAX_JOBS, "conv wgrad: 1..%d jobs", WG_MAX_JOBS);
    RVSR_CHECK_ARG(Cout > 0 && Cout % 8 == 0, "conv wgrad: Cout %d is not a multiple of 8", Cout);
    RVSR_CHECK_ARG(workspace_bytes >= conv_wgrad_tc_workspace_bytes(njobs, N, H, W, Cout), "conv wgrad: workspace too small");
    for (int j = 0; j < njobs; ++j) {
        RVSR_CHECK_ARG(x_image_stride[j] % 8 == 0, "conv wgrad: image stride");
        RVSR_CHECK_ARG(c0[j] >= 0 && c0[j] + 64 <= cin_total[j], "conv wgrad: channel slice");
        RVSR_CHECK_ARG(gw[j] != nullptr && (N == 0 || (x_c8[j] != nullptr && g_c8[j] != nullptr)), "conv wgrad: null buffer (job %d)", j);
    }
    if (N == 0) {
        for (int j = 0; j < njobs; ++j) {
            RVSR_CHECK_ARG(c0[j] == 0 && cin_total[j] == 64, "conv wgrad: empty batch with a channel slice");
            RVSR_CUDA(cudaMemsetAsync(gw[j], 0, (size_t)Cout * 64 * ks * ks * 4, s));
            if (db[j]) RVSR_CUDA(cudaMemsetAsync(db[j], 0, (size_t)Cout * 4, s));
        }
        return RVSR_OK;
    }
    EncodeTiledFn enc = get_encode();
    RVSR_CHECK_ARG(enc != nullptr, "conv wgrad: cuTensorMapEncodeTiled unavailable");
    WgradParams p;
    memset(&p, 0, sizeof(p));
    WgradOut out;
    memset(&out, 0, sizeof(out));
    const int gpl = Cout / 8;
    for (int j = 0; j < njobs; ++j) {
        const cuuint64_t dims[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)N * gpl};
        const cuuint64_t strides[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
        const cuuint32_t box[3] = {(cuuint32_t)TC_TW * 8, (cuuint32_t)TC_ROWS, 8};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&p.tmap_g[j], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(g_c8[j]), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv wgrad: cuTensorMapEncodeTiled failed (%d)", (int)r); return RVSR_E_CUDA; }
        p.x[j] = reinterpret_cast<const uint4 *>(x_c8[j]); p.x_image_stride[j] = x_image_stride[j] / 8;
        if (db[j] != nullptr) p.bias_jobs |= 1 << j;
        out.gw[j] = gw[j]; out.db[j] = db[j]; out.cin_total[j] = cin_total[j]; out.c0[j] = c0[j];
    }
    p.tiles_x = cdiv(W, TC_TW); p.tiles_y = cdiv(H, TC_ROWS); p.num_tiles = p.tiles_x * p.tiles_y * N;
    const int passes = cdiv(Cout, 64), gx = wgrad_ctas(p.num_tiles, passes * njobs);
    char *wsp = reinterpret_cast<char *>(workspace);
    wsp += (256 - (size_t)((uintptr_t)wsp % 256)) % 256;
    p.partial = reinterpret_cast<float *>(wsp);
    p.bias_partial = reinterpret_cast<float *>(wsp + align_up((size_t)gx * passes * njobs * 6 * 128 * 64 * 4, 256));
    p.g_planes = gpl;
    p.Cout = Cout; p.N = N; p.H = H; p.W = W; p.ks = ks;
    static const int dbg = getenv("RVSR_WG_DEBUG") ? atoi(getenv("RVSR_WG_DEBUG")) : 0;
    p.debug = dbg;
    p.td.tpi = (uint32_t)(p.tiles_x * p.tiles_y); p.td.m_tpi = magic_div(p.td.tpi, (uint32_t)p.num_tiles);
    p.td.tx = (uint32_t)p.tiles_x; p.td.m_tx = magic_div(p.td.tx, p.td.tpi);
    const size_t smem = (size_t)WG_STAGES * WG_STAGE + 8 * 16 + 2 * 64 * 4 + 1024;
    RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&conv_wgrad_tc_kernel), (int)smem));
    conv_wgrad_tc_kernel<<<dim3(gx, passes, njobs), WG_THREADS, smem, s>>>(p);
    RVSR_LAUNCH_CHECK();
    const int groups = ks * ks * 64 * (Cout / 4);  // Cout % 8 == 0
    int rb = (groups + 63) / 64;
    if (rb * 256 < Cout) rb = (Cout + 255) / 256;  // the bias sums ride on the first threads
    wgrad_reduce_kernel<<<dim3(rb, njobs), 256, 0, s>>>(p.partial, p.bias_partial, out, gx, Cout, ks);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

int launch_act_bwd_c8(const void *g, const void *y, void *out, long long n_elems, int act, cudaStream_t s) {
    RVSR_CHECK_ARG(n_elems % 8 == 0, "act bwd: element count");
    RVSR_CHECK_ARG(act == RVSR_ACT_LRELU || act == RVSR_ACT_RELU, "act bwd: activation %d", act);
    if (n_elems == 0) return RVSR_OK;
    act_bwd_c8_kernel<<<ew_grid(n_elems / 8), 256, 0, s>>>((const uint4 *)g, (const uint4 *)y, (uint4 *)out, n_elems / 8,
                                                           act == RVSR_ACT_LRELU ? 0.1f : 0.f);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
// g, y: [N][C/32][2H][2W][8] -> out: [N][C/8][H][W][8]; act = RVSR_ACT_NONE: pure pixel-unshuffle (y unused)
int launch_unshuffle2_act_bwd_c8(const void *g, const void *y, void *out, int N, int C, int H, int W, int act, cudaStream_t s) {
    RVSR_CHECK_ARG(C % 32 == 0, "unshuffle: C %d is not a multiple of 32", C);
    const long long quads = (long long)N * (C / 32) * H * ((W + 7) / 8) * 8;  // low-resolution pixels x input blocks, rows padded to 8
    if (quads == 0) return RVSR_OK;
    unshuffle2_act_bwd_kernel<<<ew_grid(quads * 4), 256, 0, s>>>((const uint4 *)g, (const uint4 *)y, (uint4 *)out, C / 32, H, W, quads,
                                                                act == RVSR_ACT_LRELU ? 0.1f : 0.f, act != RVSR_ACT_NONE);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
int launch_upsample2x_c8(const void *src, void *dst, long long planes, int H, int W, float scale, int backward, cudaStream_t s) {
    if (planes == 0 || H == 0 || W == 0) return RVSR_OK;
    if (!backward) {
        const long long total = planes * 4 * H * W;
        upsample2x_c8_kernel<<<ew_grid(total), 256, 0, s>>>((const uint4 *)src, (uint4 *)dst, H, W, total, scale);
    } else {  // src: gradient of the [2H][2W] output, dst: gradient of the [H][W] input
        const long long total = planes * H * W;
        upsample2x_bwd_c8_kernel<<<ew_grid(total), 256, 0, s>>>((const uint4 *)src, (uint4 *)dst, H, W, total, scale);
    }
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
int launch_nchw_to_c8_bf16(const void *src, int src_dtype, void *dst, int N, int C, int H, int W, int planes, cudaStream_t s) {
    RVSR_CHECK_ARG(planes >= cdiv(C, 8), "nchw -> c8: %d channel blocks cannot hold %d channels", planes, C);
    if (N == 0) return RVSR_OK;
    const int HW = H * W;
    RVSR_CHECK_ARG(N <= 65535 && cdiv(C, 8) <= 65535, "nchw -> c8: too many images / blocks");
    const dim3 grid((HW + 255) / 256, cdiv(C, 8), N);
    if (src_dtype == RVSR_BF16) nchw_to_c8_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16 *)src, (uint4 *)dst, C, HW, planes);
    else if (src_dtype == RVSR_F32) nchw_to_c8_kernel<float><<<grid, 256, 0, s>>>((const float *)src, (uint4 *)dst, C, HW, planes);
    else { set_error("nchw -> c8: dtype %d", src_dtype); return RVSR_E_INVALID; }
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
int launch_c8_to_nchw_bf16(const void *src, void *dst, int dst_dtype, int N, int C, int H, int W, int planes, cudaStream_t s) {
    RVSR_CHECK_ARG(planes >= cdiv(C, 8), "c8 -> nchw: %d channel blocks do not hold %d channels", planes, C);
    if (N == 0) return RVSR_OK;
    const int HW = H * W;
    RVSR_CHECK_ARG(N <= 65535 && cdiv(C, 8) <= 65535, "c8 -> nchw: too many images / blocks");
    const dim3 grid((HW + 255) / 256, cdiv(C, 8), N);
    if (dst_dtype == RVSR_BF16) c8_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const uint4 *)src, (__nv_bfloat16 *)dst, C, HW, planes);
    else if (dst_dtype == RVSR_F32) c8_to_nchw_kernel<float><<<grid, 256, 0, s>>>((const uint4 *)src, (float *)dst, C, HW, planes);
    else { set_error("c8 -> nchw: dtype %d", dst_dtype); return RVSR_E_INVALID; }
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

template <typename T> static void up2_nchw_launch(const void *src, void *dst, long long planes, int H, int W, float scale, int backward, cudaStream_t s) {
    if (!backward) {
        const long long total = planes * 4 * H * W;
        upsample2x_nchw_kernel<T><<<ew_grid(total), 256, 0, s>>>((const T *)src, (T *)dst, H, W, total, scale);
    } else {
        const long long total = planes * H * W;
        upsample2x_bwd_nchw_kernel<T><<<ew_grid(total), 256, 0, s>>>((const T *)src, (T *)dst, H, W, total, scale);
    }
}
// src [planes][H][W] -> dst [planes][2H][2W] (backward = 0), or the adjoint: src = gradient of the [2H][2W] output, dst = gradient of
// the [H][W] input (backward = 1).  planes = N * C of a contiguous NCHW tensor.
int launch_upsample2x_nchw(const void *src, void *dst, long long planes, int H, int W, float scale, int backward, int dtype, cudaStream_t s) {
    if (planes == 0 || H == 0 || W == 0) return RVSR_OK;
    if (dtype == RVSR_F32) up2_nchw_launch<float>(src, dst, planes, H, W, scale, backward, s);
    else if (dtype == RVSR_F16) up2_nchw_launch<__half>(src, dst, planes, H, W, scale, backward, s);
    else if (dtype == RVSR_BF16) up2_nchw_launch<__nv_bfloat16>(src, dst, planes, H, W, scale, backward, s);
    else { set_error("upsample2x (NCHW): dtype %d", dtype); return RVSR_E_INVALID; }
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

int launch_pool_maxavg_c8(const void *src, void *dmax, void *davg, long long planes, int H, int W, cudaStream_t s) {
    const long long total = planes * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1);
    if (total == 0) return RVSR_OK;
    pool_maxavg_c8_kernel<<<ew_grid(total), 256, 0, s>>>((const uint4 *)src, (uint4 *)dmax, (uint4 *)davg, H, W, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
int launch_pool_maxavg_bwd_c8(const void *src, const void *gmax, const void *gavg, void *gin, long long planes, int H, int W, cudaStream_t s) {
    const long long total = planes * H * W;
    if (total == 0) return RVSR_OK;
    pool_maxavg_bwd_c8_kernel<<<ew_grid(total), 256, 0, s>>>((const uint4 *)src, (const uint4 *)gmax, (const uint4 *)gavg, (uint4 *)gin, H, W, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
int launch_tsa_final_c8(const void *fea, const void *att, const void *add, void *out, long long n_elems, cudaStream_t s) {
    if (n_elems == 0) return RVSR_OK;
    tsa_final_c8_kernel<<<ew_grid(n_elems / 8), 256, 0, s>>>((const uint4 *)fea, (const uint4 *)att, (const uint4 *)add, (uint4 *)out, n_elems / 8);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
int launch_tsa_final_bwd_c8(const void *g, const void *fea, const void *att, void *g_fea, void *g_att, long long n_elems, cudaStream_t s) {
    if (n_elems == 0) return RVSR_OK;
    tsa_final_bwd_c8_kernel<<<ew_grid(n_elems / 8), 256, 0, s>>>((const uint4 *)g, (const uint4 *)fea, (const uint4 *)att, (uint4 *)g_fea, (uint4 *)g_att, n_elems / 8);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

int launch_tsa_temporal_c8(const void *aligned, const void *emb, const void *emb_ref, void *const *out, float *prob, int B, int N, int H,
                           int W, cudaStream_t s) {
    RVSR_CHECK_ARG(N >= 1 && N <= RVSR_MAX_SRC_TC, "tsa temporal: 1..%d frames", RVSR_MAX_SRC_TC);
    const long long total = (long long)B * N * H * W;
    if (total == 0) return RVSR_OK;
    TsaPtrs o;
    for (int i = 0; i < N; ++i) o.p[i] = out[i];
    tsa_temporal_c8_kernel<<<ew_grid(total), 256, 0, s>>>((const uint4 *)aligned, (const uint4 *)emb, (const uint4 *)emb_ref, o, prob, N, H * W, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
int launch_tsa_temporal_bwd_c8(const void *const *gout, const void *aligned, const void *emb, const void *emb_ref, const float *prob,
                               void *g_aligned, void *g_emb, void *g_emb_ref, int B, int N, int H, int W, cudaStream_t s) {
    RVSR_CHECK_ARG(N >= 1 && N <= RVSR_MAX_SRC_TC, "tsa temporal: 1..%d frames", RVSR_MAX_SRC_TC);
    const long long total = (long long)B * H * W;
    if (total == 0) return RVSR_OK;
    TsaPtrs g;
    for (int i = 0; i < N; ++i) g.p[i] = const_cast<void *>(gout[i]);
    tsa_temporal_bwd_c8_kernel<<<(int)((total + 127) / 128), 128, 0, s>>>(g, (const uint4 *)aligned, (const uint4 *)emb, (const uint4 *)emb_ref, prob,
                                                                         (uint4 *)g_aligned, (uint4 *)g_emb, (uint4 *)g_emb_ref, N, H * W, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

// ---------------------------------------------------------------- ModulatedDeformConvPack on C8 tensors
// Shape class of every EDVR DCN with nf = 64: C = Cout = 64, 3x3, stride 1, pad 1, 8 deformable groups.
size_t c8_mdcn_workspace_bytes(int N, int H, int W, int backward) {
    const size_t px = (size_t)N * H * W;
    if (!backward)
        return 2 * align_up(px * 128, 256) + align_up(px * 768, 256) + align_up(tc_dcn_weight_bytes(64, 64, 9) + 16, 256) + 4096;
    return align_up(px * 128, 256) + 2 * (align_up(px * 576, 256) + align_up(px * 288, 256)) + align_up(px * 256, 256) +
           align_up(dcn_bwd_tc_wt_bytes(), 256) + align_up((size_t)64 * 64 * 9 * 2, 256) + 4096;
}
int c8_mdcn_fwd(const void *x, const void *om, const float *weight, const float *bias, void *y, int N, int H, int W, int act,
                void *workspace, size_t workspace_bytes, cudaStream_t s) {
    RVSR_CHECK_ARG(workspace_bytes >= c8_mdcn_workspace_bytes(N, H, W, 0), "c8 mdcn fwd: workspace too small");
    if (N == 0) return RVSR_OK;
    const size_t px = (size_t)N * H * W;
    char *wsp = reinterpret_cast<char *>(workspace);
    wsp += (256 - (size_t)((uintptr_t)wsp % 256)) % 256;
    auto take = [&](size_t bytes) { char *r = wsp; wsp += align_up(bytes, 256); return r; };
    void *x16 = take(px * 128), *y16 = take(px * 128), *om24 = take(px * 768), *wtc = take(tc_dcn_weight_bytes(64, 64, 9) + 16);
    const int HW = H * W;
    cvt_bf16_to_f16_kernel<<<ew_grid((long long)px * 8), 256, 0, s>>>((const uint4 *)x, (uint4 *)x16, (long long)px * 8);
    RVSR_CHECK_ARG(N <= 65535, "c8 mdcn: too many images");
    om_from_c8_kernel<<<dim3((HW + 127) / 128, N), 128, 0, s>>>((const uint4 *)om, (uint4 *)om24, nullptr, nullptr, HW);
    RVSR_LAUNCH_CHECK();
    RVSR_TRY(pack_weight_dcn_tc(weight, wtc, 64, 64, 9, s));
    DcnOp op = {};
    op.x = Src{x16, (long long)64 * HW, 64, 1, -1}; op.om24 = om24; op.om24_image_stride = (long long)8 * 24 * HW;
    op.w_tc = wtc; op.bias = bias; op.out = y16; op.out_image_stride = (long long)64 * HW;
    op.N = N; op.H = H; op.W = W; op.Cout = 64; op.kh = op.kw = 3; op.stride = 1; op.pad = 1; op.dil = 1; op.dg = 8;
    op.act = act; op.out_mode = OUT_C8;
    if (!tc_dcn_supported(op)) { set_error("c8 mdcn fwd: tcgen05 DCN kernel unavailable"); return RVSR_E_UNSUPPORTED; }
    RVSR_TRY(launch_dcn_tc(op, s));
    cvt_f16_to_bf16_kernel<<<ew_grid((long long)px * 8), 256, 0, s>>>((const uint4 *)y16, (uint4 *)y, (long long)px * 8);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}
// g: gradient of y (after the activation, y = the forward output, needed only when act != none).  gx: [N][8][H][W][8] bf16,
// gom: [N][32][H][W][8] bf16, gw: [64][64][3][3] fp32 and gb: [64] fp32 -- all WRITTEN (this call's values).
int c8_mdcn_bwd(const void *x, const void *om, const float *weight, const void *g, const void *y, void *gx, void *gom, float *gw,
                float *gb, int N, int H, int W, int act, void *workspace, size_t workspace_bytes, cudaStream_t s) {
    RVSR_CHECK_ARG(workspace_bytes >= c8_mdcn_workspace_bytes(N, H, W, 1), "c8 mdcn bwd: workspace too small");
    RVSR_CHECK_ARG(act == RVSR_ACT_NONE || y != nullptr, "c8 mdcn bwd: the activation gradient needs the forward output");
    RVSR_CUDA(cudaMemsetAsync(gw, 0, (size_t)64 * 64 * 9 * 4, s));
    if (gb != nullptr) RVSR_CUDA(cudaMemsetAsync(gb, 0, 64 * 4, s));
    if (N == 0) return RVSR_OK;
    const size_t px = (size_t)N * H * W;
    char *wsp = reinterpret_cast<char *>(workspace);
    wsp += (256 - (size_t)((uintptr_t)wsp % 256)) % 256;
    auto take = [&](size_t bytes) { char *r = wsp; wsp += align_up(bytes, 256); return r; };
    void *gact = take(px * 128);
    float *off32 = (float *)take(px * 576), *msk32 = (float *)take(px * 288), *goff32 = (float *)take(px * 576), *gmsk32 = (float *)take(px * 288);
    float *gx32 = (float *)take(px * 256);
    void *wt = take(dcn_bwd_tc_wt_bytes()), *wbf = take((size_t)64 * 64 * 9 * 2);
    const int HW = H * W;
    const void *gp = g;
    if (act != RVSR_ACT_NONE) {
        RVSR_TRY(launch_act_bwd_c8(g, y, gact, (long long)px * 64, act, s));
        gp = gact;
    }
    om_from_c8_kernel<<<dim3((HW + 127) / 128, N), 128, 0, s>>>((const uint4 *)om, nullptr, off32, msk32, HW);
    RVSR_LAUNCH_CHECK();
    RVSR_TRY(launch_convert_f32_bf16(weight, wbf, 64 * 64 * 9, s));
    RVSR_TRY(pack_wt_dcn_bwd_tc(wbf, wt, s));
    RVSR_CUDA(cudaMemsetAsync(gx32, 0, px * 256, s));
    // offset / mask gradients: written by the kernel straight into gom (bf16, channel-blocked, through the sigmoid); RVSR_DCN_BWD_GOM=0
    // keeps the planar fp32 round trip + om_grad_to_c8_kernel (bit-identical results) for A/B runs
    static const bool direct = !(getenv("RVSR_DCN_BWD_GOM") != nullptr && getenv("RVSR_DCN_BWD_GOM")[0] == '0');
    if (direct)  // the 40 padding channels (blocks 27..31 of every image) get no gradient
        RVSR_CUDA(cudaMemset2DAsync((char *)gom + (size_t)27 * HW * 16, (size_t)32 * HW * 16, 0, (size_t)5 * HW * 16, (size_t)N, s));
    RVSR_TRY(launch_dcn_bwd_tc_core(x, gp, off32, msk32, wt, gx32, goff32, gmsk32, gw, N, H, W, s, direct ? gom : nullptr));
    if (gb != nullptr) {
        bias_grad_c8_kernel<<<dim3(HW >= 4096 ? 16 : 1, 8, N < 8 ? N : 8), 256, 0, s>>>((const uint4 *)gp, gb, HW, N);
        RVSR_LAUNCH_CHECK();
    }
    RVSR_TRY(launch_convert_f32_bf16(gx32, gx, (long long)px * 64, s));
    if (!direct) {
        om_grad_to_c8_kernel<<<dim3((HW + 127) / 128, 32, N), 128, 0, s>>>(goff32, gmsk32, msk32, (uint4 *)gom, HW);
        RVSR_LAUNCH_CHECK();
    }
    return RVSR_OK;
}

}  // namespace rvsr
