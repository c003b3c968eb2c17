// dcn_fused.cu -- ModulatedDeformConvPack.forward (extra_offset_mask=True) as ONE tcgen05 kernel (sm_100a only).
//
// Reference: dcn/deform_conv.py:274-292 (conv_offset_mask -> chunk / cat / sigmoid -> modulated_deform_conv) ->
// deform_conv_cuda.cpp:490-569 -> deform_conv_cuda_kernel.cu:571-633 (+ cuBLAS addmm_).  The reference writes the 216-channel
// offset/mask tensor, a concatenated copy of the offsets, the sigmoid-ed mask and a 9x-inflated `columns` buffer to HBM;
// round 1 of this repo still wrote the offsets + mask once (768 B per pixel, OUT_OM24) between two kernels.  Here nothing
// but the two inputs and the output touches HBM:
//
//   feat halo tile --TMA--> smem --tcgen05.mma (64 -> 8 groups x 32 columns, 3x3, fp32 accumulate)--> TMEM  "OM"
//   OM --tcgen05.ld--> registers of the gather threads (+ bias, sigmoid on the 9 mask columns)
//   x --bilinear gather (fp32 coordinates, fp16x2 blend, mask folded into the corner weights)--> smem, UMMA operand layout
//   smem --tcgen05.mma (576 x 64 contraction)--> TMEM --tcgen05.ld--> bias, activation --> out
//
// The offset/mask convolution is tensor-pipe bound and the gather is issue / latency bound: in one kernel the former runs
// under the latter (the tensor pipe is idle 90 % of the time in a gather-only kernel).
//
// Shape of the kernel: a CTA PAIR (cluster of 2, tcgen05 cta_group::2).  Every MMA covers M = 256 pixels = each CTA's own
// 4 x 32 tile (30 valid columns: the nine taps of the offset conv are nine shifted VIEWS of one 6 x 32 halo tile, as in
// conv_tc2_kernel), and the B operands are split between the two CTAs.  The offset/mask weights (64 -> 216, 249 KB) do not
// fit in shared memory even split over the pair next to everything else, so they are STREAMED: per tile and CTA 18 chunks of
// 8 KB (2 halves of 4 deformable groups x 9 taps) through a 6-stage ring, always L2 hits -- 147 KB per tile instead of the
// 196 KB per tile the OUT_OM24 tensor moved through L2 and HBM.  TMEM: OM half A [0,128) | OM half B [128,256) | two output
// accumulators [256,384).  A gather thread owns one pixel and one deformable group per half tile: half A's OM columns are
// released as soon as every thread has pulled its 32 columns into registers, a full tile time before they are needed again.
//
// Warps (768 threads):  0 TMA producer (halo tiles + weight chunks) | 1 offset-conv MMA issuer (leader CTA) | 2 contraction
// MMA issuer (leader CTA) | 3 TMEM allocator | 4-19 gather (warp % 4 = TMEM lane quarter = tile row) | 20-23 epilogue.
#include "tc_common.cuh"

namespace rvsr {

namespace {

constexpr int FP_GATHER_WARPS = 16, FP_GATHER_WARP0 = 4, FP_EPI_WARP0 = FP_GATHER_WARP0 + FP_GATHER_WARPS;
constexpr int FP_THREADS = 32 * (FP_EPI_WARP0 + 4);
constexpr int FP_PLANE_BYTES = (TC_ROWS + 2) * TC_TW * 16;  // one channel block of one halo copy (6 rows x 32 pixels)
constexpr int FP_COPY_BYTES = 8 * FP_PLANE_BYTES;           // 24576: one halo copy
constexpr int FP_HALO_BYTES = 3 * FP_COPY_BYTES;            // three copies, shifted by dx = 0, 1, 2 pixels (all 32 columns valid)
constexpr int FP_WCHUNK = 8 * 64 * 16;                      // one (half, rank, tap) chunk of offset/mask weights: 8 KB
constexpr int FP_SW = 4;                                    // weight ring stages
constexpr int FP_STEP_BYTES = 4 * 128 * 16;                 // gathered A operand of one step: 4 channel blocks x 128 pixels x 16 B
constexpr int FP_SPS = 3;                                   // gather steps per ring stage
constexpr int FP_STAGE_BYTES = FP_SPS * FP_STEP_BYTES;
constexpr int FP_SA = 3;                                    // gather ring stages (6 stage uses per tile)
constexpr int FP_WDCN_BYTES = 9 * 8 * 32 * 16;              // this CTA's half of the contraction weights
constexpr int FP_TMEM_COLS = 512, FP_OM_COLS = 128, FP_D_COL0 = 256;

struct alignas(64) TcPackParams {
    CUtensorMap tmap_feat;   // [W * 8, H, N * 8 planes] fp16, box {256, 6, 8}: one halo copy
    CUtensorMap tmap_wom;    // offset/mask weights as [rows][256 halfs], box {256, 16} = one 8 KB chunk
    const __half *x;
    long long x_image_stride;
    const int *x_map;        // optional image -> slot table (feature cache)
    const __half *w_dcn;     // [rank][tap][8][32][8]
    const float *bias_om;    // [27 * 8], reference channel order
    const float *bias;       // [64] or null
    __half *out;
    long long out_image_stride;
    int N, H, W, act;
    int tiles_x, tiles_y, num_tiles;
    TileDiv td;
    int debug;               // RVSR_DCN_DEBUG timing experiments (results wrong): 1 no gather loads, 2 no offset-conv MMAs, 8 no MMA issue throttle
};

__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *tmap, uint32_t bar_rank0, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_rank0), "r"(c0), "r"(c1)
        : "memory");
}
// Offsets and mask of one (pixel, deformable group): 18 fp32 offsets (dy0 dx0 .. dy8 dx8) + sigmoid(mask) as 5 fp16 pairs.
struct OmRegs {
    float off[18];
    uint32_t mk[5];
};

template <bool BLEND16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FP_THREADS, 1) dcn_pack_fused_kernel(const __grid_constant__ TcPackParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *wdcn_s = smem;
    uint8_t *wom_s = wdcn_s + FP_WDCN_BYTES;
    uint8_t *halo_s = wom_s + FP_SW * FP_WCHUNK;
    uint8_t *tap_s = halo_s + FP_HALO_BYTES;
    float *bias_om_s = reinterpret_cast<float *>(tap_s + FP_SA * FP_STAGE_BYTES);
    float *bias_s = bias_om_s + 256;
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + 64);
    // barrier indices
    constexpr int B_HFULL = 0, B_HEMPTY = B_HFULL + 1, B_WFULL = B_HEMPTY + 1, B_WEMPTY = B_WFULL + FP_SW,
                  B_OMFULL = B_WEMPTY + FP_SW, B_OMEMPTY = B_OMFULL + 2, B_AFULL = B_OMEMPTY + 2, B_AEMPTY = B_AFULL + FP_SA,
                  B_DFULL = B_AEMPTY + FP_SA, B_DEMPTY = B_DFULL + 2, B_WDCN = B_DEMPTY + 2, B_WDPEER = B_WDCN + 1,
                  B_COUNT = B_WDPEER + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + B_COUNT);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const uint32_t rank = blockIdx.x & 1u;  // == %cluster_ctarank for __cluster_dims__(2, 1, 1); provably uniform
    const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int npairs = (p.num_tiles + 1) / 2;
    const int T = cid < npairs ? (npairs - cid + nclusters - 1) / nclusters : 0;  // tile pairs of this cluster
    const bool slow_poll = !(p.debug & 16);  // RVSR_DCN_DEBUG & 16: the idle roles poll with the suspend-time hint again (A/B)

    pdl_trigger();
    if (threadIdx.x == 0) {
        mbar_init(BAR(B_HFULL), 2); mbar_init(BAR(B_HEMPTY), 1);
        for (int i = 0; i < FP_SW; ++i) { mbar_init(BAR(B_WFULL + i), 2); mbar_init(BAR(B_WEMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(B_OMFULL + i), 1); mbar_init(BAR(B_OMEMPTY + i), 2 * FP_GATHER_WARPS);
            mbar_init(BAR(B_DFULL + i), 1); mbar_init(BAR(B_DEMPTY + i), 2 * 4);
        }
        for (int i = 0; i < FP_SA; ++i) { mbar_init(BAR(B_AFULL + i), 2 * FP_GATHER_WARPS); mbar_init(BAR(B_AEMPTY + i), 1); }
        mbar_init(BAR(B_WDCN), 1); mbar_init(BAR(B_WDPEER), 1);
        fence_barrier_init();
        // this CTA's half of the contraction weights: resident for the CTA's lifetime
        mbar_expect_tx(BAR(B_WDCN), FP_WDCN_BYTES);
        const uint8_t *wg = reinterpret_cast<const uint8_t *>(p.w_dcn) + (size_t)rank * FP_WDCN_BYTES;
        for (uint32_t o = 0; o < FP_WDCN_BYTES; o += 18432) bulk_load(smem_u32(wdcn_s + o), wg + o, 18432, BAR(B_WDCN));
        prefetch_tensormap(&p.tmap_feat);
        prefetch_tensormap(&p.tmap_wom);
    }
    // bias tables: offset/mask columns in TMEM order [half][local group][32] and the output bias
    for (int i = threadIdx.x; i < 256 + 64; i += FP_THREADS) {
        if (i < 256) {
            const int g = i >> 5, j = i & 31;
            bias_om_s[i] = j < 27 ? p.bias_om[j < 18 ? g * 18 + j : 18 * 8 + g * 9 + (j - 18)] : 0.f;
        } else {
            bias_s[i - 256] = p.bias != nullptr ? p.bias[i - 256] : 0.f;
        }
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(FP_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);

    // Register budget: 768 threads x 80 registers at launch; the producer / issuer warpgroup and the epilogue warpgroup hand
    // registers to the four gather warpgroups, whose 18-step unrolled pipeline keeps two (pixel, group) offset sets, two
    // in-flight samples and their addresses live (56 + 4 x 88 + 56 = 464 <= 6 x 80).
    if (warp < FP_GATHER_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
    if (warp == 0) {
        if (lane == 0) {
            // ---- producer: halo tiles of `feat` and the streamed offset/mask weight chunks (both complete on the LEADER's
            // full barriers: leader arrive.expect_tx for both CTAs' bytes, peer a plain remote arrive)
            uint32_t wk = 0;  // weight chunks issued so far
            auto wchunk = [&](int hf, int tap) {
                const int st = (int)(wk % FP_SW);
                if (slow_poll) mbar_wait_sleep(BAR(B_WEMPTY + st), ((wk / FP_SW) & 1) ^ 1, 100);
                else mbar_wait(BAR(B_WEMPTY + st), ((wk / FP_SW) & 1) ^ 1);
                const uint32_t full0 = mapa_rank0(BAR(B_WFULL + st));
                if (rank == 0) mbar_expect_tx(BAR(B_WFULL + st), 2 * FP_WCHUNK);
                else mbar_arrive_cluster(full0);
                tma_load_2d_2sm(smem_u32(wom_s + st * FP_WCHUNK), &p.tmap_wom, full0, 0, ((hf * 2 + (int)rank) * 9 + tap) * 16);
                ++wk;
            };
            auto halo = [&](int t) {
                int tile = 2 * (cid + t * nclusters) + (int)rank;
                if (tile >= p.num_tiles) tile = p.num_tiles - 1;  // odd tile count: the peer recomputes the last tile, never stores it
                int tx, ty, n;
                tile_coords(p.td, tile, tx, ty, n);
                if (slow_poll) mbar_wait_sleep(BAR(B_HEMPTY), ((uint32_t)t & 1) ^ 1, 400);  // one stage: the previous tile's offset conv has read it
                else mbar_wait(BAR(B_HEMPTY), ((uint32_t)t & 1) ^ 1);
                const uint32_t full0 = mapa_rank0(BAR(B_HFULL));
                if (rank == 0) mbar_expect_tx(BAR(B_HFULL), 2 * FP_HALO_BYTES);
                else mbar_arrive_cluster(full0);
                // The nine taps are shifted VIEWS of a halo tile (UMMA descriptor start address).  A view shifted by dx columns
                // wraps its last dx columns into the next row, so each dx gets its own copy, loaded at x0 - 1 + dx: tap (dy, dx)
                // = copy dx advanced by dy rows, and all 32 columns of the tile are valid.
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
                    tma_load_3d_2sm(smem_u32(halo_s + dx * FP_COPY_BYTES), &p.tmap_feat, full0, (tx * TC_TW - 1 + dx) * 8, ty * TC_ROWS - 1, n * 8);
            };
            constexpr int NPRE = FP_SW < 9 ? FP_SW : 9;  // weights are static: the first chunks go out before the dependency wait
            if (T > 0)
                for (int tap = 0; tap < NPRE; ++tap) wchunk(0, tap);
            pdl_wait();
            if (T > 0) halo(0);
            for (int t = 0; t < T; ++t) {
                for (int tap = (t == 0 ? NPRE : 0); tap < 9; ++tap) wchunk(0, tap);
                for (int tap = 0; tap < 9; ++tap) wchunk(1, tap);
                // one halo stage: tile t + 1's halo waits for the end of tile t's offset conv, so every weight chunk of
                // tile t must have been requested before (the conv cannot finish without them)
                if (t + 1 < T) halo(t + 1);
            }
        }
    } else if (warp == 1) {
        if (rank != 0) {
            if (lane == 0) {  // "my half of the contraction weights has landed" -> leader
                mbar_wait(BAR(B_WDCN), 0);
                mbar_arrive_cluster(mapa_rank0(BAR(B_WDPEER)));
            }
        } else {
            // ---- offset/mask convolution: per tile 2 halves x 9 taps x 4 k-steps, M = 256 (pair), N = 128, K = 16
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(FP_OM_COLS >> 3) << 17) | ((256u >> 4) << 24);
            const uint64_t adesc0 = make_desc(smem_u32(halo_s), FP_PLANE_BYTES, 128);
            const uint64_t bdesc0 = make_desc(smem_u32(wom_s), 64 * 16, 128);
            const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
            const uint32_t a_base = (uint32_t)adesc0, b_base = (uint32_t)bdesc0;
            const bool no_mma = (p.debug & 2) != 0;
            uint32_t wst = 0, wpar = 0, wc = 0;  // weight ring position / parity, chunks consumed
            for (int t = 0; t < T; ++t) {
                mbar_wait(BAR(B_HFULL), (uint32_t)t & 1);
                tc_fence_after();
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    if (slow_poll) mbar_wait_sleep(BAR(B_OMEMPTY + hf), ((uint32_t)t & 1) ^ 1, 400);  // every gather thread has pulled the previous tile's columns
                    else mbar_wait(BAR(B_OMEMPTY + hf), ((uint32_t)t & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)hf * FP_OM_COLS;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        // at most two taps (8 MMAs, ~0.5k cycles) queued: the contraction's MMAs share the tensor pipe in issue
                        // order, and the gather ring only buffers ~3k cycles of them
                        if (wc >= 2 && !(p.debug & 8)) mbar_wait(BAR(B_WEMPTY + (wc - 2) % FP_SW), ((wc - 2) / FP_SW) & 1);
                        mbar_wait(BAR(B_WFULL + wst), wpar);
                        tc_fence_after();
                        const uint32_t a_lo = a_base + (uint32_t)(tap % 3) * (FP_COPY_BYTES >> 4) + (uint32_t)((tap / 3) * TC_TW);
                        const uint32_t b_lo = b_base + wst * (FP_WCHUNK >> 4);
                        if (elect_one()) {
                            if (!no_mma) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma_f16_2sm(d, ((uint64_t)a_hi << 32) | (a_lo + (uint32_t)kk * (2 * FP_PLANE_BYTES / 16)),
                                                 ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)kk * (2 * 64)), idesc, (tap | kk) ? 1u : 0u);
                            }
                            umma_commit_2sm(BAR(B_WEMPTY + wst));
                            if (tap == 8) umma_commit_2sm(BAR(B_OMFULL + hf));
                            if (tap == 8 && hf == 1) umma_commit_2sm(BAR(B_HEMPTY));
                        }
                        __syncwarp();
                        ++wc;
                        if (++wst == FP_SW) { wst = 0; wpar ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 2) {
        if (rank == 0) {
            // ---- contraction: per tile 18 gather steps (2 channel halves x 9 taps), three per ring stage; M = 256, N = 64
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((256u >> 4) << 24);
            mbar_wait(BAR(B_WDCN), 0);
            mbar_wait(BAR(B_WDPEER), 0);
            for (int t = 0; t < T; ++t) {
                const uint32_t buf = (uint32_t)t & 1;
                mbar_wait(BAR(B_DEMPTY + buf), (((uint32_t)t >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + FP_D_COL0 + buf * 64;
#pragma unroll
                for (int u = 0; u < 18 / FP_SPS; ++u) {
                    const int st = u % FP_SA;
                    if (slow_poll) mbar_wait_sleep(BAR(B_AFULL + st), ((uint32_t)t * (18 / FP_SPS / FP_SA) + u / FP_SA) & 1, 200);
                    else mbar_wait_idle(BAR(B_AFULL + st), ((uint32_t)t * (18 / FP_SPS / FP_SA) + u / FP_SA) & 1);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int part = 0; part < FP_SPS; ++part) {
                            const int s18 = FP_SPS * u + part, h = s18 / 9, tap = s18 % 9;
                            const uint32_t a0 = smem_u32(tap_s + st * FP_STAGE_BYTES + part * FP_STEP_BYTES);
                            const uint32_t b0 = smem_u32(wdcn_s) + (uint32_t)(tap * 8 + 4 * h) * (32 * 16);
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk)
                                umma_f16_2sm(d, make_desc(a0 + (uint32_t)(2 * kk) * 2048, 2048, 128),
                                             make_desc(b0 + (uint32_t)(2 * kk) * (32 * 16), 32 * 16, 128), idesc, (s18 | kk) ? 1u : 0u);
                        }
                        umma_commit_2sm(BAR(B_AEMPTY + st));
                        if (u == 18 / FP_SPS - 1) umma_commit_2sm(BAR(B_DFULL + buf));
                    }
                    __syncwarp();
                }
            }
        }
    }
    } else if (warp < FP_EPI_WARP0) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 88;" ::: "memory");
        // ---- gather: thread -> pixel m of the CTA's tile (TMEM lane) and channel block qq of the current half (= one
        // deformable group of 8 channels).  18 steps per tile (half 0 taps 0..8, half 1 taps 0..8), one bilinear sample per
        // thread and step, corner loads one step ahead (across halves and tiles), branch-free border handling.
        pdl_wait();
        const int lq = warp & 3, qq = (warp - FP_GATHER_WARP0) >> 2;
        const int m = lq * 32 + lane;
        const long long plane = (long long)p.H * p.W;
        const float Hf = (float)p.H, Wf = (float)p.W;
        const int Hm1 = p.H - 1, Wm1 = p.W - 1;
        const bool no_loads = (p.debug & 1) != 0;
        struct Unit { const uint4 *pl; float by, bx; bool valid; };
        struct Samp { uint4 c[4]; uint32_t w[4]; };
        auto unit_of = [&](int t, int half) {
            Unit u;
            int tile = 2 * (cid + t * nclusters) + (int)rank;
            if (tile >= p.num_tiles) tile = p.num_tiles - 1;
            int tx, ty, n;
            tile_coords(p.td, tile, tx, ty, n);
            const int y = ty * TC_ROWS + lq, x = tx * TC_TW + lane;
            u.valid = y < p.H && x < p.W;
            const long long img = p.x_map != nullptr ? __ldg(p.x_map + n) : n;
            u.pl = reinterpret_cast<const uint4 *>(p.x + img * p.x_image_stride + (long long)(half * 4 + qq) * plane * 8);
            u.by = (float)(y - 1); u.bx = (float)(x - 1);
            // (an L1 prefetch of the undeformed neighbourhood here cost 4.5 %: the L1TEX pipe is the co-bottleneck of the gather)
            return u;
        };
        // this thread's 32 OM columns of half hf -> registers (+ bias, sigmoid on the mask columns); releases the columns.
        // A pixel outside the image (ragged last tile) gets a zero mask: its samples then contribute nothing, for free.
        auto load_om = [&](OmRegs &o, int hf, uint32_t parity, bool valid) {
            mbar_wait(BAR(B_OMFULL + hf), parity);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)hf * FP_OM_COLS + (uint32_t)qq * 32 + ((uint32_t)(lq * 32) << 16);
            const float *b = bias_om_s + hf * 128 + qq * 32;
            uint32_t r[16];
            tmem_ld16_nowait(taddr, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) o.off[j] = __uint_as_float(r[j]) + b[j];
            tmem_ld16_nowait(taddr + 16, r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_rank0(BAR(B_OMEMPTY + hf)));
            o.off[16] = __uint_as_float(r[0]) + b[16];
            o.off[17] = __uint_as_float(r[1]) + b[17];
            float mv[10];
#pragma unroll
            for (int j = 0; j < 9; ++j) mv[j] = valid ? __fdividef(1.f, 1.f + __expf(-(__uint_as_float(r[2 + j]) + b[18 + j]))) : 0.f;
            mv[9] = 0.f;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const __half2 h2 = __floats2half2_rn(mv[2 * k], mv[2 * k + 1]);
                o.mk[k] = *reinterpret_cast<const uint32_t *>(&h2);
            }
        };
        auto mask_of = [&](const OmRegs &o, int tap) -> float {
            const __half2 mh = *reinterpret_cast<const __half2 *>(&o.mk[tap >> 1]);
            return (tap & 1) ? __high2float(mh) : __low2float(mh);
        };
        // One bilinear sample (deform_conv_cuda_kernel.cu:467-497, :618): zero unless -1 < py < H and -1 < px < W, each corner
        // individually bounds-checked.  Branch-free: the coordinate is clamped to [-1, H] x [-1, W] -- at -1 and at H (W) both
        // rows (columns) get a zero weight (the far one because the fraction is 0, the near one because it is outside), so a
        // sample that leaves the image contributes exactly 0 without an `inside` test; corner addresses are clamped into the image.
        auto issue = [&](Samp &sm, const Unit &u, const OmRegs &o, int tap) {
            const float py = fminf(fmaxf(u.by + (float)(tap / 3) + o.off[2 * tap], -1.f), Hf);
            const float px = fminf(fmaxf(u.bx + (float)(tap % 3) + o.off[2 * tap + 1], -1.f), Wf);
            const float fy = floorf(py), fx = floorf(px);
            const int y0 = (int)fy, x0 = (int)fx;
            const float ly = py - fy, lx = px - fx;
            const float mk = mask_of(o, tap);  // mask folded into the row weights
            const float wy0 = (unsigned)y0 <= (unsigned)Hm1 ? (1.f - ly) * mk : 0.f, wy1 = y0 < Hm1 ? ly * mk : 0.f;
            const float wx0 = (unsigned)x0 <= (unsigned)Wm1 ? 1.f - lx : 0.f, wx1 = x0 < Wm1 ? lx : 0.f;
            const float w0 = wy0 * wx0, w1 = wy0 * wx1, w2 = wy1 * wx0, w3 = wy1 * wx1;
            const int r0 = min(max(y0, 0), Hm1) * p.W, r1 = min(y0 + 1, Hm1) * p.W, x0c = min(max(x0, 0), Wm1), x1c = min(x0 + 1, Wm1);
            if (!no_loads) {
                sm.c[0] = __ldg(u.pl + (r0 + x0c)); sm.c[1] = __ldg(u.pl + (r0 + x1c));
                sm.c[2] = __ldg(u.pl + (r1 + x0c)); sm.c[3] = __ldg(u.pl + (r1 + x1c));
            }
            if (BLEND16) {
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
        if (T > 0) {
            OmRegs OM[2];
            Unit U[2];
            Samp SM[2];
#pragma unroll
            for (int k = 0; k < 4; ++k) { SM[0].c[k] = make_uint4(0, 0, 0, 0); SM[1].c[k] = make_uint4(0, 0, 0, 0); }
            U[0] = unit_of(0, 0);
            load_om(OM[0], 0, 0, U[0].valid);
            issue(SM[0], U[0], OM[0], 0);
            uint8_t *const my_a = tap_s + qq * 2048 + m * 16;
            const uint32_t afull0 = mapa_rank0(BAR(B_AFULL));
#pragma unroll 1
            for (int t = 0; t < T; ++t) {
                const bool has_next = t + 1 < T;
#pragma unroll
                for (int s18 = 0; s18 < 18; ++s18) {
                    const int h = s18 / 9, tap = s18 % 9, u = s18 / FP_SPS, st = u % FP_SA, part = s18 % FP_SPS;
                    if (tap == 1) {  // set up the next unit (other half of this tile, or the first half of the next tile)
                        if (h == 0) U[1] = unit_of(t, 1);
                        else if (has_next) U[0] = unit_of(t + 1, 0);
                    }
                    if (tap == 6) {  // ... and pull its offsets / mask out of TMEM
                        if (h == 0) load_om(OM[1], 1, (uint32_t)t & 1, U[1].valid);
                        else if (has_next) load_om(OM[0], 0, ((uint32_t)t + 1) & 1, U[0].valid);
                    }
                    if (tap < 8) issue(SM[(s18 + 1) & 1], U[h], OM[h], tap + 1);
                    else if (h == 0 || has_next) issue(SM[(s18 + 1) & 1], U[h ^ 1], OM[h ^ 1], 0);
                    const uint4 pk = blend(SM[s18 & 1]);
                    if (part == 0)  // stage free (the MMAs of its previous use completed)
                        mbar_wait(BAR(B_AEMPTY + st), (((uint32_t)t * (18 / FP_SPS / FP_SA) + u / FP_SA) & 1) ^ 1);
                    *reinterpret_cast<uint4 *>(my_a + st * FP_STAGE_BYTES + part * FP_STEP_BYTES) = pk;
                    if (part == FP_SPS - 1) {
                        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(afull0 + 8u * (uint32_t)st);
                    }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
        pdl_wait();
        const int lq = warp & 3;
        for (int t = 0; t < T; ++t) {
            const int tile = 2 * (cid + t * nclusters) + (int)rank;
            const bool real = tile < p.num_tiles;
            int tx, ty, n;
            tile_coords(p.td, real ? tile : p.num_tiles - 1, tx, ty, n);
            const uint32_t buf = (uint32_t)t & 1;
            const int y = ty * TC_ROWS + lq, x = tx * TC_TW + lane;
            const bool valid = real && y < p.H && x < p.W;
            if (slow_poll) mbar_wait_sleep(BAR(B_DFULL + buf), ((uint32_t)t >> 1) & 1, 1000);  // a tile takes ~10k cycles to gather
            else mbar_wait_idle(BAR(B_DFULL + buf), ((uint32_t)t >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + FP_D_COL0 + buf * 64 + ((uint32_t)(lq * 32) << 16);
            uint4 *o = reinterpret_cast<uint4 *>(p.out + (long long)n * p.out_image_stride) + (long long)y * p.W + x;
            const long long plane = (long long)p.H * p.W;
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {  // 32 columns at a time (register budget of this warpgroup)
                uint32_t a[2][16];
                tmem_ld16_nowait(taddr + hc * 32, a[0]);
                tmem_ld16_nowait(taddr + hc * 32 + 16, a[1]);
                tmem_ld_wait();
                if (hc == 1) {  // accumulator fully in registers: hand the buffer back to the issuer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_rank0(BAR(B_DEMPTY + buf)));
                }
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 pk;
                        __half2 *h = reinterpret_cast<__half2 *>(&pk);
                        const float *b = bias_s + hc * 32 + q * 8;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float v0 = __uint_as_float(a[q >> 1][(q & 1) * 8 + 2 * i]) + b[2 * i];
                            float v1 = __uint_as_float(a[q >> 1][(q & 1) * 8 + 2 * i + 1]) + b[2 * i + 1];
                            if (p.act == RVSR_ACT_LRELU) { v0 = fmaxf(v0, 0.1f * v0); v1 = fmaxf(v1, 0.1f * v1); }
                            else if (p.act == RVSR_ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                            h[i] = __floats2half2_rn(v0, v1);
                        }
                        o[(hc * 4 + q) * plane] = pk;
                    }
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();  // no CTA exits (or frees TMEM) while its partner can still touch its barriers / operands
    if (warp == 3) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(FP_TMEM_COLS) : "memory");
}

// offset/mask weights, streamed layout: [half][rank][tap][8 blocks][64 rows][8]; row n of (half, rank): deformable group
// g = half * 4 + rank * 2 + n / 32, column j = n % 32 of [dy0 dx0 .. dy8 dx8 m0 .. m8 0 0 0 0 0]
__global__ void pack_weight_om_stream_kernel(const float *__restrict__ w, __half *__restrict__ dst, int Cin, int total) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int e = i % 8, n = (i / 8) % 64, q = (i / 512) % 8, tap = (i / 4096) % 9, rk = (i / 36864) % 2, hf = i / 73728;
        const int g = hf * 4 + rk * 2 + n / 32, j = n % 32, cin = q * 8 + e;
        const int co = j < 18 ? g * 18 + j : (j < 27 ? 18 * 8 + g * 9 + (j - 18) : -1);
        dst[i] = __float2half_rn(co >= 0 ? w[((long long)co * Cin + cin) * 9 + tap] : 0.f);
    }
}

}  // namespace

size_t tc_pack_om_weight_bytes(int Cout, int Cin, int dg) { return (Cout == 216 && Cin == 64 && dg == 8) ? (size_t)2 * 2 * 9 * FP_WCHUNK : 0; }
int pack_weight_om_stream(const float *w_oihw, void *dst, int Cout, int Cin, int dg, cudaStream_t s) {
    RVSR_CHECK_ARG(tc_pack_om_weight_bytes(Cout, Cin, dg) > 0, "fused pack: offset/mask conv must be 64 -> 216 (8 groups)");
    const int total = 2 * 2 * 9 * 8 * 64 * 8;
    pack_weight_om_stream_kernel<<<(total + 255) / 256, 256, 0, s>>>(w_oihw, reinterpret_cast<__half *>(dst), Cin, total);
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

bool tc_pack_fused_supported(const PackFusedOp &op) {
    static const bool off = getenv("RVSR_DCN_FUSED") != nullptr && getenv("RVSR_DCN_FUSED")[0] == '0';
    if (off || get_encode() == nullptr) return false;
    if (op.x.C != 64 || op.Cout != 64 || op.dg != 8 || op.w_om == nullptr || op.w_dcn2 == nullptr || op.bias_om == nullptr) return false;
    if (op.x.fixed_frame >= 0 || op.feat == nullptr) return false;
    return (long long)cdiv(op.W, TC_TW) * cdiv(op.H, TC_ROWS) * op.N >= 2;
}

int launch_pack_fused(const PackFusedOp &op, cudaStream_t s) {
    RVSR_CHECK_ARG(tc_pack_fused_supported(op), "fused pack: unsupported configuration");
    EncodeTiledFn enc = get_encode();
    TcPackParams p;
    memset(&p, 0, sizeof(p));
    {
        const cuuint64_t dims[3] = {(cuuint64_t)op.W * 8, (cuuint64_t)op.H, (cuuint64_t)op.N * 8};
        const cuuint64_t strides[2] = {(cuuint64_t)op.W * 16, (cuuint64_t)op.H * op.W * 16};
        const cuuint32_t box[3] = {(cuuint32_t)TC_TW * 8, (cuuint32_t)(TC_ROWS + 2), 8};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&p.tmap_feat, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(op.feat), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("fused pack: cuTensorMapEncodeTiled(feat) failed (%d)", (int)r); return RVSR_E_CUDA; }
    }
    {
        const cuuint64_t dims[2] = {256, (cuuint64_t)(2 * 2 * 9 * 16)};
        const cuuint64_t strides[1] = {512};
        const cuuint32_t box[2] = {256, 16};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&p.tmap_wom, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(op.w_om), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("fused pack: cuTensorMapEncodeTiled(weights) failed (%d)", (int)r); return RVSR_E_CUDA; }
    }
    p.x = reinterpret_cast<const __half *>(op.x.ptr); p.x_image_stride = op.x.image_stride; p.x_map = op.x.map;
    p.w_dcn = reinterpret_cast<const __half *>(op.w_dcn2); p.bias_om = op.bias_om; p.bias = op.bias;
    p.out = reinterpret_cast<__half *>(op.out); p.out_image_stride = op.out_image_stride;
    p.N = op.N; p.H = op.H; p.W = op.W; p.act = op.act;
    p.tiles_x = cdiv(op.W, TC_TW); p.tiles_y = cdiv(op.H, TC_ROWS); p.num_tiles = p.tiles_x * p.tiles_y * op.N;
    p.td.tpi = (uint32_t)(p.tiles_x * p.tiles_y); p.td.m_tpi = magic_div(p.td.tpi, (uint32_t)p.num_tiles);
    p.td.tx = (uint32_t)p.tiles_x; p.td.m_tx = magic_div(p.td.tx, p.td.tpi);
    static const int ddbg = getenv("RVSR_DCN_DEBUG") ? atoi(getenv("RVSR_DCN_DEBUG")) : 0;
    p.debug = ddbg;
    const size_t smem = FP_WDCN_BYTES + FP_SW * FP_WCHUNK + FP_HALO_BYTES + FP_SA * FP_STAGE_BYTES + (256 + 64) * 4 + 512 + 1024;
    static const bool blend32 = getenv("RVSR_DCN_BLEND") != nullptr && strcmp(getenv("RVSR_DCN_BLEND"), "fp32") == 0;
    const int npairs = (p.num_tiles + 1) / 2;
    int clusters = sm_count() / 2;
    if (clusters > npairs) clusters = npairs;
    if (blend32) {
        RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&dcn_pack_fused_kernel<false>), (int)smem));
        launch_k(dcn_pack_fused_kernel<false>, dim3(2 * clusters), dim3(FP_THREADS), smem, s, p);
    } else {
        RVSR_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void *>(&dcn_pack_fused_kernel<true>), (int)smem));
        launch_k(dcn_pack_fused_kernel<true>, dim3(2 * clusters), dim3(FP_THREADS), smem, s, p);
    }
    RVSR_LAUNCH_CHECK();
    return RVSR_OK;
}

}  // namespace rvsr
