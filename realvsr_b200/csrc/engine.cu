// engine.cu -- EDVR / EDVR_NoUp inference as one C++ call (see include/rvsr_b200.h).
//
// What it replaces in the reference (IanYeung/RealVSR, codes/models/archs/EDVR_arch.py):
//   EDVR.forward :258-320, EDVR_NoUp.forward :358-404, PCD_Align.forward :98-132,
//   TSA_Fusion.forward :168-208, ResidualBlock_noBN (arch_util.py:135-139),
//   ModulatedDeformConvPack.forward (dcn/deform_conv.py:274-292).
//
// Differences in structure (not in results):
//   * the N per-frame PCD passes the reference runs in a Python loop (:297-303) are one
//     batch of B*N images; the "reference frame" operand is a broadcast view, not a clone;
//   * torch.cat([a, b]) -> conv is a two-source convolution (no concat copy);
//   * activation, bias, residual add, pixel-shuffle are conv epilogues;
//   * activations live in a channel-blocked layout [N][C/8][H][W][8] so that one pixel of
//     one deformable group (nf/groups = 8 channels in every shipped config) is one 128-bit load.
#include "engine.cuh"

#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <set>
#include <utility>

namespace rvsr {

// ---------------------------------------------------------------- error string
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }
int ensure_max_dynamic_smem(const void *func, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<const void *, int>> done;
    int dev = 0;
    RVSR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (done.count({func, dev})) return RVSR_OK;
    RVSR_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.insert({func, dev});
    return RVSR_OK;
}
static thread_local int g_pdl_scope = 0;
PdlScope::PdlScope() { ++g_pdl_scope; }
PdlScope::~PdlScope() { --g_pdl_scope; }
bool pdl_enabled() {  // RVSR_PDL is read at every launch so that tests can compare both modes in one process
    const char *env = getenv("RVSR_PDL");
    return !(env != nullptr && env[0] == '0') && g_pdl_scope > 0;
}

// ---------------------------------------------------------------- state_dict contract
void Engine::expect(const std::string &name, std::vector<int64_t> shape) {
    names_.push_back(name);
    RawWeight w;
    w.shape = std::move(shape);
    raw_[name] = w;
}
void Engine::expect_conv(const std::string &name, int co, int ci, int k) {
    expect(name + ".weight", {co, ci, k, k});
    expect(name + ".bias", {co});
}

Engine::Engine(const rvsr_edvr_config &cfg) : cfg_(cfg) {
    if (cfg_.center < 0) cfg_.center = cfg_.nframes / 2;
    const int nf = cfg_.nf, N = cfg_.nframes;
    // Same order as the reference modules register their parameters.
    if (!cfg_.upsample) cfg_.predeblur = cfg_.HR_in = 0;  // EDVR_NoUp ignores both (EDVR_arch.py:335-339, :358-404)
    if (cfg_.predeblur) {  // Predeblur_ResNet_Pyramid (EDVR_arch.py:15-59) + conv_1x1 (:226-227)
        const std::string d = "pre_deblur.";
        if (cfg_.HR_in) {
            expect_conv(d + "conv_first_1", nf, 3, 3);
            expect_conv(d + "conv_first_2", nf, nf, 3);
            expect_conv(d + "conv_first_3", nf, nf, 3);
        } else {
            expect_conv(d + "conv_first", nf, 3, 3);
        }
        for (const char *n : {"RB_L1_1", "RB_L1_2", "RB_L1_3", "RB_L1_4", "RB_L1_5", "RB_L2_1", "RB_L2_2", "RB_L3_1"}) {
            expect_conv(d + n + ".conv1", nf, nf, 3);
            expect_conv(d + n + ".conv2", nf, nf, 3);
        }
        expect_conv(d + "deblur_L2_conv", nf, nf, 3);
        expect_conv(d + "deblur_L3_conv", nf, nf, 3);
        expect_conv("conv_1x1", nf, nf, 1);
    } else if (cfg_.HR_in) {  // EDVR_arch.py:228-231
        expect_conv("conv_first_1", nf, cfg_.nc, 3);
        expect_conv("conv_first_2", nf, nf, 3);
        expect_conv("conv_first_3", nf, nf, 3);
    } else {
        expect_conv("conv_first", nf, cfg_.nc, 3);
    }
    for (int i = 0; i < cfg_.front_RBs; ++i) {
        expect_conv("feature_extraction." + std::to_string(i) + ".conv1", nf, nf, 3);
        expect_conv("feature_extraction." + std::to_string(i) + ".conv2", nf, nf, 3);
    }
    for (const char *n : {"fea_L2_conv1", "fea_L2_conv2", "fea_L3_conv1", "fea_L3_conv2"}) expect_conv(n, nf, nf, 3);
    const std::string p = "pcd_align.";
    auto dcn = [&](const std::string &n) {
        expect(p + n + ".weight", {nf, nf, 3, 3});
        expect(p + n + ".bias", {nf});
        expect_conv(p + n + ".conv_offset_mask", cfg_.groups * 27, nf, 3);
    };
    expect_conv(p + "L3_offset_conv1", nf, 2 * nf, 3);
    expect_conv(p + "L3_offset_conv2", nf, nf, 3);
    dcn("L3_dcnpack");
    expect_conv(p + "L2_offset_conv1", nf, 2 * nf, 3);
    expect_conv(p + "L2_offset_conv2", nf, 2 * nf, 3);
    expect_conv(p + "L2_offset_conv3", nf, nf, 3);
    dcn("L2_dcnpack");
    expect_conv(p + "L2_fea_conv", nf, 2 * nf, 3);
    expect_conv(p + "L1_offset_conv1", nf, 2 * nf, 3);
    expect_conv(p + "L1_offset_conv2", nf, 2 * nf, 3);
    expect_conv(p + "L1_offset_conv3", nf, nf, 3);
    dcn("L1_dcnpack");
    expect_conv(p + "L1_fea_conv", nf, 2 * nf, 3);
    expect_conv(p + "cas_offset_conv1", nf, 2 * nf, 3);
    expect_conv(p + "cas_offset_conv2", nf, nf, 3);
    dcn("cas_dcnpack");
    if (cfg_.w_TSA) {
        const std::string t = "tsa_fusion.";
        expect_conv(t + "tAtt_1", nf, nf, 3);
        expect_conv(t + "tAtt_2", nf, nf, 3);
        expect_conv(t + "fea_fusion", nf, N * nf, 1);
        expect_conv(t + "sAtt_1", nf, N * nf, 1);
        expect_conv(t + "sAtt_2", nf, 2 * nf, 1);
        expect_conv(t + "sAtt_3", nf, nf, 3);
        expect_conv(t + "sAtt_4", nf, nf, 1);
        expect_conv(t + "sAtt_5", nf, nf, 3);
        expect_conv(t + "sAtt_L1", nf, nf, 1);
        expect_conv(t + "sAtt_L2", nf, 2 * nf, 3);
        expect_conv(t + "sAtt_L3", nf, nf, 3);
        expect_conv(t + "sAtt_add_1", nf, nf, 1);
        expect_conv(t + "sAtt_add_2", nf, nf, 1);
    } else {
        expect_conv("tsa_fusion", nf, N * nf, 1);
    }
    for (int i = 0; i < cfg_.back_RBs; ++i) {
        expect_conv("recon_trunk." + std::to_string(i) + ".conv1", nf, nf, 3);
        expect_conv("recon_trunk." + std::to_string(i) + ".conv2", nf, nf, 3);
    }
    if (cfg_.upsample) {
        expect_conv("upconv1", nf * 4, nf, 3);
        expect_conv("upconv2", 64 * 4, nf, 3);  // literal 64, EDVR_arch.py:250
    }
    expect_conv("HRconv", 64, 64, 3);
    expect_conv("conv_last", cfg_.nc, 64, 3);
}

Engine::~Engine() {
    prof_clear();
    for (void *p : owned_) cudaFree(p);
}

int Engine::set_weight(const char *name, const float *dev_ptr, const int64_t *shape, int ndim, cudaStream_t s) {
    auto it = raw_.find(name ? name : "");
    RVSR_CHECK_ARG(it != raw_.end(), "unexpected key in state_dict: \"%s\"", name ? name : "(null)");
    RawWeight &w = it->second;
    bool same = (int)w.shape.size() == ndim;
    for (int i = 0; same && i < ndim; ++i) same = w.shape[i] == shape[i];
    RVSR_CHECK_ARG(same, "size mismatch for %s", name);
    RVSR_CHECK_ARG(dev_ptr != nullptr, "null pointer for %s", name);
    size_t n = 1;
    for (int64_t d : w.shape) n *= (size_t)d;
    if (w.dev == nullptr) {
        RVSR_CUDA(cudaMalloc(&w.dev, n * sizeof(float)));
        owned_.push_back(w.dev);
    }
    RVSR_CUDA(cudaMemcpyAsync(w.dev, dev_ptr, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    w.set = true;
    finalized_ = false;
    return RVSR_OK;
}

int Engine::finalize(cudaStream_t s) {
    for (const auto &n : names_)
        if (!raw_[n].set) {
            set_error("missing key in state_dict: \"%s\"", n.c_str());
            return RVSR_E_STATE;
        }
    for (const auto &n : names_) {
        const size_t pos = n.rfind(".weight");
        if (pos == std::string::npos || pos + 7 != n.size()) continue;
        const std::string base = n.substr(0, pos);
        const RawWeight &w = raw_[n];
        PackedConv &pc = packed_[base];
        pc.Cout = (int)w.shape[0];
        pc.Cin = (int)w.shape[1];
        pc.ks = (int)w.shape[2];
        pc.bias = raw_[base + ".bias"].dev;
        const float *wsrc = w.dev;
        const bool first = base == "conv_first" || base == "conv_first_1" || base == "pre_deblur.conv_first" || base == "pre_deblur.conv_first_1";
        if (cfg_.precision == RVSR_F16 && first && pc.Cin < 16) {
            // tensor-core K granularity is 16 channels: the LQ frames are stored zero-padded to 16
            // channels and conv_first's weight gets matching zero input channels
            float *padded = nullptr;
            RVSR_CUDA(cudaMalloc(&padded, (size_t)pc.Cout * 16 * pc.ks * pc.ks * sizeof(float)));
            owned_.push_back(padded);
            RVSR_TRY(pad_weight_cin(w.dev, padded, pc.Cout, pc.Cin, 16, pc.ks * pc.ks, s));
            wsrc = padded;
            pc.Cin = 16;
        }
        const int cout_pad = cdiv(pc.Cout, 64) * 64;
        const size_t bytes = (size_t)cdiv(pc.Cin, 8) * pc.ks * pc.ks * 8 * cout_pad * sizeof(float);
        if (pc.w_simt == nullptr) {
            RVSR_CUDA(cudaMalloc(&pc.w_simt, bytes));
            owned_.push_back(pc.w_simt);
        }
        const int cin = pc.Cin;
        RVSR_TRY(pack_weight_simt(wsrc, pc.w_simt, pc.Cout, pc.Cin, pc.ks, &cin, 1, cout_pad, s));
        if (cfg_.precision == RVSR_F16) {
            const bool is_dcn = base.size() > 8 && base.compare(base.size() - 8, 8, "_dcnpack") == 0;
            const bool shuffle = (base == "upconv1" || base == "upconv2");
            const bool is_om = base.size() > 17 && base.compare(base.size() - 17, 17, ".conv_offset_mask") == 0;
            // the tcgen05 DCN kernels want offsets/mask in OUT_OM24 order (nf == 64 or 128, whole 8-channel
            // blocks per deformable group); otherwise the offset conv stays planar on the CUDA-core kernel
            const int mode = shuffle ? 1 : ((is_om && (cfg_.nf == 64 || cfg_.nf == 128) &&
                                             (cfg_.nf / cfg_.groups) % 8 == 0) ? 2 : 0);
            const size_t tb = is_dcn ? tc_dcn_weight_bytes(pc.Cout, pc.Cin, pc.ks * pc.ks)
                                     : ((is_om && mode != 2) ? 0 : tc_conv_weight_bytes(pc.Cout, pc.Cin, pc.ks, mode));
            if (tb > 0) {
                if (pc.w_tc == nullptr) {
                    RVSR_CUDA(cudaMalloc(&pc.w_tc, tb));
                    owned_.push_back(pc.w_tc);
                }
                if (is_dcn)
                    RVSR_TRY(pack_weight_dcn_tc(wsrc, pc.w_tc, pc.Cout, pc.Cin, pc.ks * pc.ks, s));
                else
                    RVSR_TRY(pack_weight_tc(wsrc, pc.w_tc, pc.Cout, pc.Cin, pc.ks, mode, s));
            }
            if (base == "conv_last" && tc_tapn_weight_bytes(pc.Cout, pc.Cin, pc.ks) > 0) {
                if (pc.w_tapn == nullptr) {
                    RVSR_CUDA(cudaMalloc(&pc.w_tapn, tc_tapn_weight_bytes(pc.Cout, pc.Cin, pc.ks)));
                    owned_.push_back(pc.w_tapn);
                }
                RVSR_TRY(pack_weight_tapn(wsrc, pc.w_tapn, pc.Cout, pc.Cin, s));
            }
            const bool dcn64 = is_dcn && pc.Cout == 64 && pc.Cin == 64 && pc.ks == 3;  // fused pack: contraction weights, pair layout
            const size_t tb2 = dcn64 ? tc2_weight_bytes(64, 64, 3, 0)
                                     : ((!is_dcn && !(is_om && mode != 2)) ? tc2_weight_bytes(pc.Cout, pc.Cin, pc.ks, mode) : 0);
            if (tb2 > 0) {
                if (pc.w_tc2 == nullptr) {
                    RVSR_CUDA(cudaMalloc(&pc.w_tc2, tb2));
                    owned_.push_back(pc.w_tc2);
                }
                RVSR_TRY(pack_weight_tc2(wsrc, pc.w_tc2, pc.Cout, pc.Cin, pc.ks, dcn64 ? 0 : mode, s));
            }
            if (is_om && tc_pack_om_weight_bytes(pc.Cout, pc.Cin, cfg_.groups) > 0) {
                if (pc.w_om_stream == nullptr) {
                    RVSR_CUDA(cudaMalloc(&pc.w_om_stream, tc_pack_om_weight_bytes(pc.Cout, pc.Cin, cfg_.groups)));
                    owned_.push_back(pc.w_om_stream);
                }
                RVSR_TRY(pack_weight_om_stream(wsrc, pc.w_om_stream, pc.Cout, pc.Cin, cfg_.groups, s));
            }
        }
    }
    // nf = 128 (BASELINE cfg4): the weights-resident tcgen05 kernels hold at most a 64-wide output tile of a 256-channel
    // contraction per CTA pair, so a 128-output convolution runs as two launches over the two halves of its output
    // channels ("#o0" / "#o1": rows [0, 64) / [64, 128) of the OIHW weight, contiguous), each reading its 128-channel
    // sources as pairs of 64-channel sources (Plan::conv).
    if (cfg_.precision == RVSR_F16) {
        for (const auto &n : names_) {
            const size_t pos = n.rfind(".weight");
            if (pos == std::string::npos || pos + 7 != n.size()) continue;
            const std::string base = n.substr(0, pos);
            const RawWeight &w = raw_[n];
            const bool is_dcn = base.size() > 8 && base.compare(base.size() - 8, 8, "_dcnpack") == 0;
            if (w.shape.size() != 4 || w.shape[0] != 128 || base == "upconv1" || base == "upconv2") continue;
            const PackedConv &full = packed_[base];   // Cin as packed (conv_first: padded to 16)
            const int Cin = full.Cin, ks = full.ks, KK = ks * ks;
            if (Cin % 16 != 0) continue;
            if (is_dcn) {
                if (Cin != 128 || ks != 3) continue;
                for (int h = 0; h < 2; ++h) {  // "#i0" / "#i1": input channels [0, 64) / [64, 128), all 128 outputs
                    PackedConv &pc = packed_[base + (h == 0 ? "#i0" : "#i1")];
                    pc.Cout = 128; pc.Cin = 64; pc.ks = 3; pc.bias = full.bias;
                    if (pc.w_tc == nullptr) {
                        RVSR_CUDA(cudaMalloc(&pc.w_tc, tc_dcn_half_weight_bytes()));
                        owned_.push_back(pc.w_tc);
                    }
                    RVSR_TRY(pack_weight_dcn_tc_half(w.dev, pc.w_tc, h, s));
                }
                continue;
            }
            const float *wsrc = w.dev;
            if (Cin != (int)w.shape[1]) {  // conv_first: the zero-padded copy made above is not kept; rebuild it
                float *padded = nullptr;
                RVSR_CUDA(cudaMalloc(&padded, (size_t)128 * Cin * KK * sizeof(float)));
                owned_.push_back(padded);
                RVSR_TRY(pad_weight_cin(w.dev, padded, 128, (int)w.shape[1], Cin, KK, s));
                wsrc = padded;
            }
            for (int h = 0; h < 2; ++h) {
                PackedConv &pc = packed_[base + (h == 0 ? "#o0" : "#o1")];
                pc.Cout = 64; pc.Cin = Cin; pc.ks = ks;
                pc.bias = full.bias != nullptr ? full.bias + h * 64 : nullptr;
                const float *wh = wsrc + (size_t)h * 64 * Cin * KK;
                const size_t tb = tc_conv_weight_bytes(64, Cin, ks, 0), tb2 = tc2_weight_bytes(64, Cin, ks, 0);
                if (tb > 0) {
                    if (pc.w_tc == nullptr) { RVSR_CUDA(cudaMalloc(&pc.w_tc, tb)); owned_.push_back(pc.w_tc); }
                    RVSR_TRY(pack_weight_tc(wh, pc.w_tc, 64, Cin, ks, 0, s));
                }
                if (tb2 > 0) {
                    if (pc.w_tc2 == nullptr) { RVSR_CUDA(cudaMalloc(&pc.w_tc2, tb2)); owned_.push_back(pc.w_tc2); }
                    RVSR_TRY(pack_weight_tc2(wh, pc.w_tc2, 64, Cin, ks, 0, s));
                }
            }
        }
    }
    // Split packs for the convolutions over torch.cat([neighbour, reference]) (PCD *_offset_conv1, EDVR_arch.py:99,:109,
    // :118,:127): conv(cat[a, b]) = conv_a(a) + conv_b(b), and b -- the reference (centre) frame's features -- is the same
    // for all N frames of a window.  The engine computes conv_b once per window and adds it to conv_a's accumulator
    // before bias + activation (ConvOp::res_pre), instead of recomputing it N times inside a 128-channel convolution.
    if (cfg_.precision == RVSR_F16 && cfg_.nf == 64) {
        for (const char *n : {"pcd_align.L3_offset_conv1", "pcd_align.L2_offset_conv1", "pcd_align.L1_offset_conv1", "pcd_align.cas_offset_conv1"}) {
            const std::string base = n;
            const RawWeight &w = raw_[base + ".weight"];
            if (w.shape.size() != 4 || w.shape[0] != 64 || w.shape[1] != 128 || w.shape[2] != 3) continue;
            for (int h = 0; h < 2; ++h) {
                PackedConv &pc = packed_[base + (h == 0 ? "#a" : "#b")];
                pc.Cout = 64; pc.Cin = 64; pc.ks = 3;
                pc.bias = h == 0 ? raw_[base + ".bias"].dev : nullptr;
                if (pc.w_simt == nullptr) {  // fp32 OIHW half [64][64][3][3], the source of the two tcgen05 packs (not a simt pack)
                    RVSR_CUDA(cudaMalloc(&pc.w_simt, (size_t)64 * 64 * 9 * sizeof(float)));
                    owned_.push_back(pc.w_simt);
                    RVSR_CUDA(cudaMalloc(&pc.w_tc, tc_conv_weight_bytes(64, 64, 3, 0)));
                    owned_.push_back(pc.w_tc);
                    RVSR_CUDA(cudaMalloc(&pc.w_tc2, tc2_weight_bytes(64, 64, 3, 0)));
                    owned_.push_back(pc.w_tc2);
                }
                RVSR_CUDA(cudaMemcpy2DAsync(pc.w_simt, (size_t)64 * 9 * sizeof(float), w.dev + (size_t)h * 64 * 9, (size_t)128 * 9 * sizeof(float),
                                            (size_t)64 * 9 * sizeof(float), 64, cudaMemcpyDeviceToDevice, s));
                RVSR_TRY(pack_weight_tc(pc.w_simt, pc.w_tc, 64, 64, 3, 0, s));
                RVSR_TRY(pack_weight_tc2(pc.w_simt, pc.w_tc2, 64, 64, 3, 0, s));
            }
        }
    }
    finalized_ = true;
    return RVSR_OK;
}

// ---------------------------------------------------------------- forward plan
namespace {

template <typename T> struct Plan {
    Engine *eng;
    Arena &ar;
    bool dry;
    cudaStream_t s;
    const std::map<std::string, PackedConv> &packed;
    bool use_tc;
    int launches = 0;
    int rc = RVSR_OK;

    Act make(int N, int C, int H, int W) {
        Act a;
        a.N = N; a.C = C; a.H = H; a.W = W;
        a.p = ar.alloc((size_t)a.elems() * sizeof(T));
        if (a.p == nullptr && rc == RVSR_OK) {
            set_error("workspace too small");
            rc = RVSR_E_WORKSPACE;
        }
        return a;
    }
    // the plan is done with this activation: its memory may be handed to a later layer (see Arena)
    void drop(Act &a) {
        static const bool keep = getenv("RVSR_ARENA_REUSE") != nullptr && getenv("RVSR_ARENA_REUSE")[0] == '0';
        if (!keep && a.p != nullptr) ar.free(a.p);
        a.p = nullptr;
    }
    static Src src_of(const Act &a) {
        Src r;
        r.ptr = a.p; r.image_stride = a.image_elems(); r.C = a.C; r.frames = 1; r.fixed_frame = -1;
        r.map = nullptr; r.map_images = 0;
        return r;
    }
    // image n reads the `frame`-th image of its group of `frames` (PCD reference features)
    static Src src_fixed(const Act &a, int frames, int frame) {
        Src r = src_of(a);
        r.frames = frames; r.fixed_frame = frame;
        return r;
    }
    // image n reads slot map[n] of a cache tensor with `a.N` slots (sliding-window feature cache)
    static Src src_mapped(const Act &a, const int *map) {
        Src r = src_of(a);
        r.map = map; r.map_images = a.N;
        return r;
    }
    // B images, image b reads image b*frames + frame of `a` (TSA: per-frame slices of a [B*N] tensor)
    static Src src_slice(const Act &a, int frames, int frame) {
        Src r = src_of(a);
        r.ptr = reinterpret_cast<const T *>(a.p) + (long long)frame * a.image_elems();
        r.image_stride = a.image_elems() * frames;
        return r;
    }
    const PackedConv *get(const std::string &name) {
        auto it = packed.find(name);
        if (it == packed.end()) {
            if (rc == RVSR_OK) { set_error("engine: no packed weights for %s", name.c_str()); rc = RVSR_E_STATE; }
            return nullptr;
        }
        return &it->second;
    }
    void note(int r) {
        ++launches;
        if (r != RVSR_OK && rc == RVSR_OK) rc = r;
    }
    // every kernel launch goes through here: counted, and (in profiling mode) bracketed by
    // CUDA events on the launching stream with its algorithmic flops / bytes attached
    template <typename F> void launch(const std::string &label, double flops, double bytes, F &&f) {
        ProfEntry *pe = eng->prof_begin(label, flops, bytes, s);
        note(f());
        eng->prof_end(pe, s);
        static const bool sync_each = getenv("RVSR_SYNC_EACH") != nullptr;  // debug: find the launch that faults
        if (sync_each && rc == RVSR_OK) {
            const cudaError_t e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) {
                set_error("launch '%s' failed: %s", label.c_str(), cudaGetErrorString(e));
                rc = RVSR_E_CUDA;
            }
        }
    }

    // generic convolution with fused epilogue; out allocated here
    Act conv(const std::string &name, std::initializer_list<Src> srcs, int N, int H, int W, int act,
             int stride = 1, int out_mode = OUT_C8, const Act *residual = nullptr, int sig_from = 1 << 30,
             int res_pre = 0, int res_div = 1, bool tc_only = false, double flops_alg = -1.0) {
        const PackedConv *pc = get(name);
        if (pc == nullptr) return Act();
        const int Ho = stride == 1 ? H : (H - 1) / 2 + 1, Wo = stride == 1 ? W : (W - 1) / 2 + 1;
        Act o;
        if (out_mode == OUT_C8_SHUFFLE2) {
            o = make(N, pc->Cout / 4, 2 * Ho, 2 * Wo);
        } else if (out_mode == OUT_PLANAR_F32 || out_mode == OUT_OM24) {
            o.N = N; o.C = pc->Cout; o.H = Ho; o.W = Wo;
            const size_t words = out_mode == OUT_OM24 ? (size_t)(pc->Cout / 27) * 24 : (size_t)pc->Cout;
            o.p = ar.alloc((size_t)N * words * Ho * Wo * sizeof(float));
            if (o.p == nullptr && rc == RVSR_OK) { set_error("workspace too small"); rc = RVSR_E_WORKSPACE; }
        } else {
            o = make(N, pc->Cout, Ho, Wo);
        }
        if (dry || rc != RVSR_OK) return o;
        ConvOp op = {};
        int cin = 0;
        for (const Src &sr : srcs) {
            cin += sr.C;
            if (use_tc && sizeof(T) == 2 && sr.C == 128 && op.nsrc + 2 <= RVSR_MAX_SRC_TC) {
                // tcgen05 kernels stage at most 64 channels per source: a 128-channel tensor = two sources, the second 8 planes in
                Src a = sr, b = sr;
                a.C = b.C = 64;
                b.ptr = reinterpret_cast<const T *>(sr.ptr) + (long long)64 * H * W;
                op.src[op.nsrc++] = a; op.src[op.nsrc++] = b;
            } else if (op.nsrc < RVSR_MAX_SRC_TC) {
                op.src[op.nsrc++] = sr;
            }
        }
        if (cin != pc->Cin) {
            set_error("engine: %s expects %d input channels, got %d", name.c_str(), pc->Cin, cin);
            rc = RVSR_E_INVALID;
            return o;
        }
        op.w_simt = pc->w_simt; op.w_tc = pc->w_tc; op.w_tc2 = pc->w_tc2; op.bias = pc->bias;
        op.out = o.p;
        op.out_image_stride = out_mode == OUT_PLANAR_F32 ? (long long)pc->Cout * Ho * Wo
                              : out_mode == OUT_OM24     ? (long long)(pc->Cout / 27) * 24 * Ho * Wo
                                                         : o.image_elems();
        op.dg = out_mode == OUT_OM24 ? pc->Cout / 27 : 0;
        if (residual != nullptr) { op.residual = residual->p; op.res_image_stride = residual->image_elems(); }
        op.res_pre = res_pre; op.res_div = res_div;
        op.N = N; op.H = H; op.W = W; op.Cout = pc->Cout; op.ks = pc->ks; op.stride = stride;
        op.act = act; op.out_mode = out_mode; op.sig_from = sig_from;
        const double px = (double)N * Ho * Wo;
        // algorithmic work of the REFERENCE's layer; flops_alg overrides it where the plan splits a layer
        const double flops = flops_alg >= 0 ? flops_alg : 2.0 * cin * pc->Cout * pc->ks * pc->ks * px;
        const double obytes = out_mode == OUT_PLANAR_F32 ? px * pc->Cout * 4
                              : out_mode == OUT_OM24     ? px * (pc->Cout / 27) * 96.0
                                                         : px * pc->Cout * sizeof(T);
        const double bytes = (double)N * H * W * cin * sizeof(T) + obytes + (residual ? obytes : 0);
        // 128 output channels: two 64-wide launches on the "#o0" / "#o1" half packs (see finalize)
        if (use_tc && sizeof(T) == 2 && pc->Cout == 128 && out_mode == OUT_C8 && packed.count(name + "#o0") && packed.count(name + "#o1")) {
            bool ok = true;
            ConvOp hv[2];
            for (int h = 0; h < 2 && ok; ++h) {
                const PackedConv &ph = packed.find(name + (h == 0 ? "#o0" : "#o1"))->second;
                hv[h] = op;
                hv[h].w_simt = nullptr; hv[h].w_tc = ph.w_tc; hv[h].w_tc2 = ph.w_tc2; hv[h].bias = ph.bias;
                hv[h].Cout = 64;
                hv[h].out = reinterpret_cast<T *>(o.p) + (long long)h * 64 * Ho * Wo;
                if (residual != nullptr) hv[h].residual = reinterpret_cast<const T *>(residual->p) + (long long)h * 64 * Ho * Wo;
                ok = tc_conv_supported(hv[h]);
            }
            if (ok) {
                const double px2 = (double)N * Ho * Wo;
                for (int h = 0; h < 2; ++h)
                    launch(std::string("tc:conv") + std::to_string(pc->ks) + "x" + std::to_string(pc->ks) + "_co128h:" + name,
                           (flops_alg >= 0 ? flops_alg : 2.0 * cin * pc->Cout * pc->ks * pc->ks * px2) / 2,
                           ((double)N * H * W * cin * sizeof(T)) + (px2 * 64 * sizeof(T)) * (residual ? 2 : 1),
                           [&] { return launch_conv_tc(hv[h], s); });
                return o;
            }
        }
        const bool tc = use_tc && op.nsrc <= RVSR_MAX_SRC_TC && tc_conv_supported(op);
        if (!tc) {  // the CUDA-core kernel takes whole sources (<= 7): undo the 64-channel split
            op.nsrc = 0;
            for (const Src &sr : srcs) op.src[op.nsrc++] = sr;
        }
        if (tc_only && !tc) {
            set_error("engine: %s needs the tcgen05 conv kernel", name.c_str());
            rc = RVSR_E_STATE;
            return o;
        }
        if (out_mode == OUT_OM24 && !tc) {
            set_error("engine: %s: OUT_OM24 needs the tcgen05 conv kernel", name.c_str());
            rc = RVSR_E_STATE;
            return o;
        }
        // label = <kernel family>:<shape class>:<weight name>; bench.py aggregates by the first two fields
        const std::string kind = std::string(tc ? "tc:" : "simt:") + "conv" + std::to_string(pc->ks) + "x" + std::to_string(pc->ks) +
                                 (out_mode == OUT_OM24 ? "_om24" : (out_mode == OUT_PLANAR_F32 ? "_planar" : "")) + "_co" +
                                 std::to_string(pc->Cout) + ":";
        launch(kind + name, flops, bytes, [&] { return tc ? launch_conv_tc(op, s) : launch_conv_simt<T>(op, s); });
        return o;
    }

    // ModulatedDeformConvPack with extra_offset_mask=True (deform_conv.py:274-292)
    Act dcn_pack(const std::string &name, const Act &x, const Act &feat, int dg, int act, const int *xmap = nullptr) {
        const int K = 9;
        const PackedConv *pc = get(name);
        const PackedConv *pom = get(name + ".conv_offset_mask");
        if (pc == nullptr || pom == nullptr) return Act();
        // tcgen05 pair: offset conv writes OUT_OM24, gather kernel consumes it.  Otherwise planar fp32
        // [N][27*dg][H][W]: first 18*dg = offsets (o1|o2 of chunk(3) concatenated back = the first two
        // thirds), last 9*dg = sigmoid(mask).
        const PackedConv *pi0 = packed.count(name + "#i0") ? &packed.find(name + "#i0")->second : nullptr;   // nf = 128 half packs
        const PackedConv *pi1 = packed.count(name + "#i1") ? &packed.find(name + "#i1")->second : nullptr;
        const bool c128 = x.C == 128 && pc->Cout == 128 && feat.C == 128 && pi0 != nullptr && pi1 != nullptr;
        const bool om24 = use_tc && pom->w_tc != nullptr && sizeof(T) == 2 && (x.C / dg) % 8 == 0 &&
                          ((pc->w_tc != nullptr && x.C == 64 && pc->Cout == 64 && feat.C == 64) || c128);
        // ONE kernel for the whole pack (dcn_fused.cu): offsets / mask go from the offset conv's TMEM accumulator straight into
        // the gather threads' registers.  RVSR_DCN_FUSED=0 keeps the round-1 kernel pair (OUT_OM24 tensor in HBM) for A/B runs.
        if (om24 && pom->w_om_stream != nullptr && pc->w_tc2 != nullptr && pom->bias != nullptr) {
            PackFusedOp f = {};
            f.x = xmap != nullptr ? src_mapped(x, xmap) : src_of(x);
            f.feat = feat.p; f.w_om = pom->w_om_stream; f.bias_om = pom->bias; f.w_dcn2 = pc->w_tc2; f.bias = pc->bias;
            f.N = feat.N; f.H = x.H; f.W = x.W; f.Cout = pc->Cout; f.dg = dg; f.act = act;
            if (tc_pack_fused_supported(f)) {
                Act o = make(feat.N, pc->Cout, x.H, x.W);
                if (dry || rc != RVSR_OK) return o;
                f.out = o.p; f.out_image_stride = o.image_elems();
                const double px = (double)feat.N * x.H * x.W;
                const double flops = (2.0 * feat.C * pom->Cout * K + 2.0 * x.C * pc->Cout * K + 8.0 * x.C * K) * px;
                const double bytes = px * (feat.C + x.C + pc->Cout) * sizeof(T);  // 22.1 MB per L1 image
                launch("tc:dcn_pack_fused:" + name, flops, bytes, [&] { return launch_pack_fused(f, s); });
                return o;
            }
        }
        Act om = conv(name + ".conv_offset_mask", {src_of(feat)}, feat.N, feat.H, feat.W, RVSR_ACT_NONE, 1,
                      om24 ? OUT_OM24 : OUT_PLANAR_F32, nullptr, 2 * dg * K);
        Act o = make(feat.N, pc->Cout, x.H, x.W);
        if (dry || rc != RVSR_OK) { drop(om); return o; }
        DcnOp op = {};
        op.x = xmap != nullptr ? src_mapped(x, xmap) : src_of(x);
        if (om24) {
            op.om24 = om.p;
            op.om24_image_stride = (long long)dg * 24 * x.H * x.W;
        } else {
            op.offset = reinterpret_cast<const float *>(om.p);
            op.mask = op.offset + (long long)2 * dg * K * x.H * x.W;
            op.offset_image_stride = op.mask_image_stride = (long long)3 * dg * K * x.H * x.W;
        }
        op.w_simt = pc->w_simt; op.w_tc = c128 ? pi0->w_tc : pc->w_tc; op.w_tc_hi = c128 ? pi1->w_tc : nullptr; op.bias = pc->bias;
        op.out = o.p; op.out_image_stride = o.image_elems();
        op.N = feat.N; op.H = x.H; op.W = x.W; op.Cout = pc->Cout;
        op.kh = op.kw = 3; op.stride = 1; op.pad = 1; op.dil = 1; op.dg = dg;
        op.act = act; op.out_mode = OUT_C8;
        const double px = (double)feat.N * x.H * x.W;
        const double flops = (2.0 * x.C * pc->Cout * K + 8.0 * x.C * K) * px;  // contraction + gather
        const double bytes = px * (x.C * sizeof(T) + (om24 ? 96.0 * dg : 3.0 * dg * K * 4) + pc->Cout * sizeof(T));
        const bool tc = om24 && tc_dcn_supported(op);
        if (om24 && !tc) {
            set_error("engine: %s: tcgen05 DCN kernel rejected an OUT_OM24 plan", name.c_str());
            rc = RVSR_E_STATE;
            return o;
        }
        launch(std::string(tc ? "tc:" : "simt:") + "dcn_gather_mma:" + name, flops, bytes,
               [&] { return tc ? launch_dcn_tc(op, s) : launch_dcn_simt<T>(op, s); });
        drop(om);
        return o;
    }
    Act up2(const Act &a, float scale) {
        Act o = make(a.N, a.C, 2 * a.H, 2 * a.W);
        if (!dry && rc == RVSR_OK)
            launch("glue:upsample2x:", 0, (double)a.elems() * sizeof(T) * 5, [&] {
                return launch_upsample2x<T>((const T *)a.p, (T *)o.p, a.N, a.C, a.H, a.W, scale, s); });
        return o;
    }
    void pools(const Act &a, Act &mx, Act &av) {
        mx = make(a.N, a.C, (a.H - 1) / 2 + 1, (a.W - 1) / 2 + 1);
        av = make(a.N, a.C, mx.H, mx.W);
        if (!dry && rc == RVSR_OK)
            launch("glue:pool_maxavg:", 0, (double)a.elems() * sizeof(T) * 1.5, [&] {
                return launch_pool_maxavg<T>((const T *)a.p, (T *)mx.p, (T *)av.p, a.N, a.C, a.H, a.W, s); });
    }
    // conv(cat([a, ref])) with `ref` constant over the `frames` images of a window (see Engine::finalize): split packs
    // when they exist (fp16, nf = 64), otherwise the two-source convolution.  refB: the reference features of the B windows.
    Act conv_cat_ref(const std::string &name, const Src &a, const Src &ref_all, const Src &refB, int NB_, int B_, int frames, int H, int W, int act) {
        static const bool split_on = !(getenv("RVSR_SPLIT_CAT") != nullptr && getenv("RVSR_SPLIT_CAT")[0] == '0');
        if (split_on && use_tc && sizeof(T) == 2 && packed.count(name + "#a") && packed.count(name + "#b") && NB_ == B_ * frames) {
            // profile accounting: the pair together is the reference's 128 -> 64 convolution over NB_ images -- its
            // algorithmic FLOPs are booked on the #a launch, none on #b (the roofline counts the reference's work)
            Act r = conv(name + "#b", {refB}, B_, H, W, RVSR_ACT_NONE, 1, OUT_C8, nullptr, 1 << 30, 0, 1, true, 0.0);
            Act o = conv(name + "#a", {a}, NB_, H, W, act, 1, OUT_C8, &r, 1 << 30, 1, frames, true,
                         2.0 * 128 * 64 * 9 * (double)NB_ * H * W);
            drop(r);
            return o;
        }
        return conv(name, {a, ref_all}, NB_, H, W, act);
    }
    // one ResidualBlock_noBN by its full prefix (Predeblur pyramid): x + conv2(relu(conv1(x)))
    // (consumes `cur`: its memory is released once the block's output exists)
    Act resblock1(const std::string &b, Act cur) {
        Act t = conv(b + ".conv1", {src_of(cur)}, cur.N, cur.H, cur.W, RVSR_ACT_RELU);
        Act o = conv(b + ".conv2", {src_of(t)}, cur.N, cur.H, cur.W, RVSR_ACT_NONE, 1, OUT_C8, &cur);
        drop(t); drop(cur);
        return o;
    }
    // a + up2(b)  (bilinear x2, align_corners=False), one pass
    Act up2_add(Act b, Act a) {  // consumes both
        Act o = make(b.N, b.C, 2 * b.H, 2 * b.W);
        if (!dry && rc == RVSR_OK)
            launch("glue:upsample2x_add:", 0, (double)b.elems() * sizeof(T) * 9, [&] {
                return launch_upsample2x<T>((const T *)b.p, (T *)o.p, b.N, b.C, b.H, b.W, 1.f, s, (const T *)a.p); });
        drop(a); drop(b);
        return o;
    }
    // `keep_input`: the caller still needs `cur` (a tap); every intermediate of the run is released as soon as it is dead
    Act resblocks(const std::string &prefix, int count, Act cur, bool keep_input = false) {
        for (int i = 0; i < count; ++i) {
            const std::string b = prefix + "." + std::to_string(i);
            Act t = conv(b + ".conv1", {src_of(cur)}, cur.N, cur.H, cur.W, RVSR_ACT_RELU);
            Act o = conv(b + ".conv2", {src_of(t)}, cur.N, cur.H, cur.W, RVSR_ACT_NONE, 1, OUT_C8, &cur);
            drop(t);
            if (!(keep_input && i == 0)) drop(cur);
            cur = o;
        }
        return cur;
    }
};

}  // namespace

template <typename T>
int Engine::run(Arena &ar, bool dry, const void *x, int x_dtype, void *out, int out_dtype, int B, int Hin, int Win,
                cudaStream_t s, const CacheArgs *ca) {
    const int nf = cfg_.nf, N = cfg_.nframes, nc = cfg_.nc, dg = cfg_.groups, ctr = cfg_.center;
    // HR_in: frames arrive at the OUTPUT resolution, two stride-2 convolutions bring the features to 1/4 (EDVR_arch.py:267-274)
    const bool hr = cfg_.HR_in != 0;
    const int H = hr ? Hin / 4 : Hin, W = hr ? Win / 4 : Win;
    const int LR = RVSR_ACT_LRELU, NONE = RVSR_ACT_NONE;
    // RVSR_DISABLE_TC=1 routes the fp16 engine through the CUDA-core kernels (debug / cross-check)
    static const bool tc_off = getenv("RVSR_DISABLE_TC") != nullptr && getenv("RVSR_DISABLE_TC")[0] == '1';
    Plan<T> P{this, ar, dry, s, packed_, cfg_.precision == RVSR_F16 && !tc_off};
    using PT = Plan<T>;
    PdlScope pdl_scope;  // programmatic dependent launch for every kernel of the plan (common.cuh)
    // Three modes share this plan:
    //   full    (ca == null)        x = [B, N, nc, H, W] windows -> out
    //   extract (ca->extract)       x = [F, nc, H, W] frames -> pyramid written into cache slots [slot0, slot0+F)
    //   cached  (ca && !extract)    windows given as slot indices into the cache; x = all cached LQ frames
    const bool extract = ca != nullptr && ca->extract, cached = ca != nullptr && !ca->extract;
    const int NB = extract ? B : B * N;   // images in the feature / alignment stages (extract: B = frame count)

    Act L1, L2, L3;
    const int *map_nbr = nullptr, *map_ref = nullptr, *map_ctr = nullptr;
    // dst = v, releasing what dst held (v was computed FROM the old dst: arguments are evaluated before the call)
    auto replace = [&](Act &dst, Act v) { P.drop(dst); dst = v; };
    if (!cached) {
        // ---- LQ frames -> channel-blocked
        const int nc_store = (cfg_.precision == RVSR_F16 && nc < 16) ? 16 : nc;  // see finalize(): K granularity
        Act xin = P.make(NB, nc_store, Hin, Win);
        if (!dry && P.rc == RVSR_OK)
            P.launch("glue:pack_input:", 0, (double)xin.elems() * sizeof(T) * 1.4, [&] {
                return x_dtype == RVSR_F32
                           ? launch_pack_nchw<T, float>((const float *)x, (T *)xin.p, NB, nc, Hin, Win, s, nc_store)
                           : launch_pack_nchw<T, __half>((const __half *)x, (T *)xin.p, NB, nc, Hin, Win, s, nc_store); });
        // ---- per-frame feature pyramid (EDVR_arch.py:262-283)
        auto stem = [&](const std::string &pre) {  // conv_first, or the HR_in stem conv_first_1 -> _2 (s2) -> _3 (s2)
            if (!hr) {
                Act o = P.conv(pre + "conv_first", {PT::src_of(xin)}, NB, Hin, Win, LR);
                P.drop(xin);
                return o;
            }
            Act a = P.conv(pre + "conv_first_1", {PT::src_of(xin)}, NB, Hin, Win, LR);
            P.drop(xin);
            replace(a, P.conv(pre + "conv_first_2", {PT::src_of(a)}, NB, Hin, Win, LR, 2));
            replace(a, P.conv(pre + "conv_first_3", {PT::src_of(a)}, NB, a.H, a.W, LR, 2));
            return a;
        };
        if (cfg_.predeblur) {  // Predeblur_ResNet_Pyramid.forward (EDVR_arch.py:43-59), then conv_1x1 without activation (:265)
            const std::string d = "pre_deblur.";
            Act l1 = stem(d);
            Act l2 = P.conv(d + "deblur_L2_conv", {PT::src_of(l1)}, NB, H, W, LR, 2);
            Act l3 = P.conv(d + "deblur_L3_conv", {PT::src_of(l2)}, NB, l2.H, l2.W, LR, 2);
            l3 = P.resblock1(d + "RB_L3_1", l3);
            l2 = P.up2_add(l3, P.resblock1(d + "RB_L2_1", l2));
            l2 = P.resblock1(d + "RB_L2_2", l2);
            l1 = P.up2_add(l2, P.resblock1(d + "RB_L1_2", P.resblock1(d + "RB_L1_1", l1)));
            for (const char *n : {"RB_L1_3", "RB_L1_4", "RB_L1_5"}) l1 = P.resblock1(d + n, l1);
            L1 = P.conv("conv_1x1", {PT::src_of(l1)}, NB, H, W, NONE);
            P.drop(l1);
        } else {
            L1 = stem("");
        }
        L1 = P.resblocks("feature_extraction", cfg_.front_RBs, L1);
        L2 = P.conv("fea_L2_conv1", {PT::src_of(L1)}, NB, H, W, LR, 2);
        replace(L2, P.conv("fea_L2_conv2", {PT::src_of(L2)}, NB, L2.H, L2.W, LR));
        L3 = P.conv("fea_L3_conv1", {PT::src_of(L2)}, NB, L2.H, L2.W, LR, 2);
        replace(L3, P.conv("fea_L3_conv2", {PT::src_of(L3)}, NB, L3.H, L3.W, LR));
        if (extract) {
            if (!dry && P.rc == RVSR_OK) {  // copy the pyramid of the F frames into their (contiguous) cache slots
                const Act *lv[3] = {&L1, &L2, &L3};
                char *dst = reinterpret_cast<char *>(ca->cache);
                for (int k = 0; k < 3; ++k) {
                    const size_t img = (size_t)lv[k]->image_elems() * sizeof(T);
                    if (cudaMemcpyAsync(dst + (size_t)ca->slot0 * img, lv[k]->p, img * (size_t)NB, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
                        set_error("extract_features: cache copy failed");
                        return RVSR_E_CUDA;
                    }
                    dst += (size_t)ca->n_slots * img;  // level regions are laid out [L1 slots][L2 slots][L3 slots]
                }
                launches_ = P.launches;
            }
            return P.rc;
        }
    } else {
        // ---- features come from the cache: level regions [L1 x n_slots][L2 x n_slots][L3 x n_slots]
        char *base = reinterpret_cast<char *>(ca->cache);
        L1.N = L2.N = L3.N = ca->n_slots; L1.C = L2.C = L3.C = nf;
        L1.H = H; L1.W = W; L2.H = H / 2; L2.W = W / 2; L3.H = H / 4; L3.W = W / 4;
        L1.p = base;
        L2.p = base + (size_t)L1.image_elems() * sizeof(T) * ca->n_slots;
        L3.p = reinterpret_cast<char *>(L2.p) + (size_t)L2.image_elems() * sizeof(T) * ca->n_slots;
        // device tables: neighbour slot of image b*N+i, reference (centre) slot of its window, centre slot per window
        int *maps = reinterpret_cast<int *>(ar.alloc((size_t)(2 * NB + B) * sizeof(int)));
        if (maps == nullptr && P.rc == RVSR_OK) { set_error("workspace too small"); P.rc = RVSR_E_WORKSPACE; }
        map_nbr = maps; map_ref = maps + NB; map_ctr = maps + 2 * NB;
        if (!dry && P.rc == RVSR_OK) {
            std::vector<int> h((size_t)2 * NB + B);
            for (int b = 0; b < B; ++b) {
                for (int i = 0; i < N; ++i) {
                    h[b * N + i] = ca->window_slots[b * N + i];
                    h[NB + b * N + i] = ca->window_slots[b * N + ctr];
                }
                h[2 * NB + b] = ca->window_slots[b * N + ctr];
            }
            host_maps_.push_back(std::move(h));  // keep alive until the async copy has been consumed
            if (cudaMemcpyAsync(maps, host_maps_.back().data(), host_maps_.back().size() * sizeof(int), cudaMemcpyHostToDevice, s) != cudaSuccess) {
                set_error("forward_cached: slot table upload failed");
                return RVSR_E_CUDA;
            }
        }
    }
    auto nbr = [&](const Act &a) { return cached ? PT::src_mapped(a, map_nbr) : PT::src_of(a); };
    auto ref = [&](const Act &a) { return cached ? PT::src_mapped(a, map_ref) : PT::src_fixed(a, N, ctr); };
    auto refB = [&](const Act &a) { return cached ? PT::src_mapped(a, map_ctr) : PT::src_slice(a, N, ctr); };  // one per window

    // ---- PCD alignment of all N frames at once (EDVR_arch.py:98-132, :297-303)
    const std::string p = "pcd_align.";
    // (L1 / L2 / L3, `aligned` and `fused` stay allocated: they are the parity taps of rvsr_engine_read_tap)
    Act o3 = P.conv_cat_ref(p + "L3_offset_conv1", nbr(L3), ref(L3), refB(L3), NB, B, N, L3.H, L3.W, LR);
    replace(o3, P.conv(p + "L3_offset_conv2", {PT::src_of(o3)}, NB, L3.H, L3.W, LR));
    Act f3 = P.dcn_pack(p + "L3_dcnpack", L3, o3, dg, LR, map_nbr);

    Act o2 = P.conv_cat_ref(p + "L2_offset_conv1", nbr(L2), ref(L2), refB(L2), NB, B, N, L2.H, L2.W, LR);
    Act o3u = P.up2(o3, 2.f);
    P.drop(o3);
    replace(o2, P.conv(p + "L2_offset_conv2", {PT::src_of(o2), PT::src_of(o3u)}, NB, L2.H, L2.W, LR));
    P.drop(o3u);
    replace(o2, P.conv(p + "L2_offset_conv3", {PT::src_of(o2)}, NB, L2.H, L2.W, LR));
    Act f2 = P.dcn_pack(p + "L2_dcnpack", L2, o2, dg, NONE, map_nbr);
    Act f3u = P.up2(f3, 1.f);
    P.drop(f3);
    replace(f2, P.conv(p + "L2_fea_conv", {PT::src_of(f2), PT::src_of(f3u)}, NB, L2.H, L2.W, LR));
    P.drop(f3u);

    Act o1 = P.conv_cat_ref(p + "L1_offset_conv1", nbr(L1), ref(L1), refB(L1), NB, B, N, H, W, LR);
    Act o2u = P.up2(o2, 2.f);
    P.drop(o2);
    replace(o1, P.conv(p + "L1_offset_conv2", {PT::src_of(o1), PT::src_of(o2u)}, NB, H, W, LR));
    P.drop(o2u);
    replace(o1, P.conv(p + "L1_offset_conv3", {PT::src_of(o1)}, NB, H, W, LR));
    Act f1 = P.dcn_pack(p + "L1_dcnpack", L1, o1, dg, NONE, map_nbr);
    P.drop(o1);
    Act f2u = P.up2(f2, 1.f);
    P.drop(f2);
    replace(f1, P.conv(p + "L1_fea_conv", {PT::src_of(f1), PT::src_of(f2u)}, NB, H, W, NONE));  // no lrelu (:125)
    P.drop(f2u);

    Act oc = P.conv_cat_ref(p + "cas_offset_conv1", PT::src_of(f1), ref(L1), refB(L1), NB, B, N, H, W, LR);
    replace(oc, P.conv(p + "cas_offset_conv2", {PT::src_of(oc)}, NB, H, W, LR));
    Act aligned = P.dcn_pack(p + "cas_dcnpack", f1, oc, dg, LR);
    P.drop(oc);
    P.drop(f1);

    // ---- fusion
    Act fused;
    auto frame_sources = [&](const Act &a) {
        std::vector<Src> v;
        for (int i = 0; i < N; ++i) v.push_back(PT::src_slice(a, N, i));
        return v;
    };
    auto conv_frames = [&](const std::string &name, const Act &a, int act) {
        // 1x1 conv over the N*nf channels of [B, N*C, H, W] == N sources of nf channels
        std::vector<Src> v = frame_sources(a);
        switch (N) {
            case 1: return P.conv(name, {v[0]}, B, H, W, act);
            case 2: return P.conv(name, {v[0], v[1]}, B, H, W, act);
            case 3: return P.conv(name, {v[0], v[1], v[2]}, B, H, W, act);
            case 4: return P.conv(name, {v[0], v[1], v[2], v[3]}, B, H, W, act);
            case 5: return P.conv(name, {v[0], v[1], v[2], v[3], v[4]}, B, H, W, act);
            case 6: return P.conv(name, {v[0], v[1], v[2], v[3], v[4], v[5]}, B, H, W, act);
            default: return P.conv(name, {v[0], v[1], v[2], v[3], v[4], v[5], v[6]}, B, H, W, act);
        }
    };
    if (cfg_.w_TSA) {
        const std::string t = "tsa_fusion.";
        Act emb_ref = P.conv(t + "tAtt_2", {PT::src_slice(aligned, N, ctr)}, B, H, W, NONE);
        Act emb = P.conv(t + "tAtt_1", {PT::src_of(aligned)}, NB, H, W, NONE);
        Act ali = P.make(NB, nf, H, W);
        if (!dry && P.rc == RVSR_OK)
            P.launch("glue:tsa_temporal:", 4.0 * ali.elems(), (double)ali.elems() * sizeof(T) * 4, [&] {
                return launch_tsa_temporal<T>((const T *)emb.p, (const T *)emb_ref.p, (const T *)aligned.p,
                                              (T *)ali.p, B, N, nf, H, W, s); });
        P.drop(emb); P.drop(emb_ref);
        Act fea = conv_frames(t + "fea_fusion", ali, LR);
        Act att = conv_frames(t + "sAtt_1", ali, LR);
        P.drop(ali);
        Act mx, av;
        P.pools(att, mx, av);
        replace(att, P.conv(t + "sAtt_2", {PT::src_of(mx), PT::src_of(av)}, B, mx.H, mx.W, LR));
        P.drop(mx); P.drop(av);
        Act attL = P.conv(t + "sAtt_L1", {PT::src_of(att)}, B, att.H, att.W, LR);
        Act mx2, av2;
        P.pools(attL, mx2, av2);
        replace(attL, P.conv(t + "sAtt_L2", {PT::src_of(mx2), PT::src_of(av2)}, B, mx2.H, mx2.W, LR));
        P.drop(mx2); P.drop(av2);
        replace(attL, P.conv(t + "sAtt_L3", {PT::src_of(attL)}, B, attL.H, attL.W, LR));
        Act attLu = P.up2(attL, 1.f);
        P.drop(attL);
        replace(att, P.conv(t + "sAtt_3", {PT::src_of(att)}, B, att.H, att.W, LR, 1, OUT_C8, &attLu));  // lrelu, then + att_L
        P.drop(attLu);
        replace(att, P.conv(t + "sAtt_4", {PT::src_of(att)}, B, att.H, att.W, LR));
        Act attu = P.up2(att, 1.f);
        replace(att, P.conv(t + "sAtt_5", {PT::src_of(attu)}, B, H, W, NONE));
        P.drop(attu);
        Act add = P.conv(t + "sAtt_add_1", {PT::src_of(att)}, B, H, W, LR);
        replace(add, P.conv(t + "sAtt_add_2", {PT::src_of(add)}, B, H, W, NONE));
        fused = P.make(B, nf, H, W);
        if (!dry && P.rc == RVSR_OK)
            P.launch("glue:tsa_final:", 0, (double)fused.elems() * sizeof(T) * 4, [&] {
                return launch_tsa_final<T>((const T *)fea.p, (const T *)att.p, (const T *)add.p, (T *)fused.p,
                                           fused.elems(), s); });
        P.drop(fea); P.drop(att); P.drop(add);
    } else {
        fused = conv_frames("tsa_fusion", aligned, NONE);
    }

    // ---- reconstruction (EDVR_arch.py:310-319 / :398-403)
    Act r = P.resblocks("recon_trunk", cfg_.back_RBs, fused, true);
    auto next_r = [&](Act v) {  // r = v; the old r is released unless it is still the `fused` tap (back_RBs == 0)
        if (r.p != fused.p) P.drop(r);
        r = v;
    };
    int scale = 1;  // output resolution / resolution of the frames the base is taken from
    if (cfg_.upsample) {
        next_r(P.conv("upconv1", {PT::src_of(r)}, B, H, W, LR, 1, OUT_C8_SHUFFLE2));
        next_r(P.conv("upconv2", {PT::src_of(r)}, B, r.H, r.W, LR, 1, OUT_C8_SHUFFLE2));
        scale = hr ? 1 : 4;  // HR_in: base = the centre frame itself (EDVR_arch.py:315-316)
    }
    next_r(P.conv("HRconv", {PT::src_of(r)}, B, r.H, r.W, LR));
    // conv_last + base frame: one tcgen05 launch that writes the NCHW result (OUT_FINAL) when the shape allows it
    bool fused_final = false;
    if (P.use_tc && sizeof(T) == 2 && nc <= 8) {
        const PackedConv *pc = P.get("conv_last");
        if (pc != nullptr && pc->w_tc != nullptr && r.C % 16 == 0 && r.C <= 64) {
            fused_final = true;
            if (!dry && P.rc == RVSR_OK) {
                ConvOp op = {};
                op.src[0] = PT::src_of(r); op.nsrc = 1;
                op.w_simt = pc->w_simt; op.w_tc = pc->w_tc; op.w_tc2 = nullptr; op.bias = pc->bias;
                op.out = out; op.out_image_stride = 0;
                op.N = B; op.H = r.H; op.W = r.W; op.Cout = pc->Cout; op.ks = pc->ks; op.stride = 1;
                op.act = NONE; op.out_mode = OUT_FINAL; op.sig_from = 1 << 30;
                op.fin.x = x; op.fin.center_map = map_ctr; op.fin.x_dtype = x_dtype; op.fin.out_dtype = out_dtype;
                op.fin.frames = N; op.fin.center = ctr; op.fin.nc = nc; op.fin.scale = scale;
                // RVSR_TAPN=0 keeps the ordinary implicit GEMM (36 MMAs per tile) for A/B timing
                static const bool tapn_on = !(getenv("RVSR_TAPN") != nullptr && getenv("RVSR_TAPN")[0] == '0');
                if (tapn_on && pc->w_tapn != nullptr) {
                    const double px = (double)B * r.H * r.W;
                    P.launch("tc:conv3x3_tapn_co" + std::to_string(pc->Cout) + ":conv_last", 2.0 * r.C * pc->Cout * 9 * px,
                             px * r.C * sizeof(T) + px * nc * (out_dtype == RVSR_F32 ? 4 : 2) + (double)B * nc * Hin * Win * (x_dtype == RVSR_F32 ? 4 : 2),
                             [&] { return launch_conv_tapn(op, pc->w_tapn, s); });
                } else if (!tc_conv_supported(op)) {
                    fused_final = false;
                } else {
                    const double px = (double)B * r.H * r.W;
                    P.launch("tc:conv3x3_final_co" + std::to_string(pc->Cout) + ":conv_last", 2.0 * r.C * pc->Cout * 9 * px,
                             px * r.C * sizeof(T) + px * nc * (out_dtype == RVSR_F32 ? 4 : 2) + (double)B * nc * Hin * Win * (x_dtype == RVSR_F32 ? 4 : 2),
                             [&] { return launch_conv_tc(op, s); });
                }
            }
        }
    }
    if (dry && fused_final) {
        // the dry run cannot ask tc_conv_supported (no pointers yet): reserve the unfused tensor so both plans agree
        fused_final = false;
    }
    Act last = fused_final ? Act() : P.conv("conv_last", {PT::src_of(r)}, B, r.H, r.W, NONE);
    if (!fused_final && !dry && P.rc == RVSR_OK) {
        const T *lp = (const T *)last.p;
        P.launch("glue:final_add_base:", 0, (double)last.elems() * sizeof(T) + (double)B * nc * last.H * last.W * 4, [&] {
            if (x_dtype == RVSR_F32 && out_dtype == RVSR_F32)
                return launch_final_add<T, float, float>(lp, (const float *)x, (float *)out, B, N, ctr, nc, Hin, Win, scale, s, map_ctr);
            if (x_dtype == RVSR_F32)
                return launch_final_add<T, float, __half>(lp, (const float *)x, (__half *)out, B, N, ctr, nc, Hin, Win, scale, s, map_ctr);
            if (out_dtype == RVSR_F32)
                return launch_final_add<T, __half, float>(lp, (const __half *)x, (float *)out, B, N, ctr, nc, Hin, Win, scale, s, map_ctr);
            return launch_final_add<T, __half, __half>(lp, (const __half *)x, (__half *)out, B, N, ctr, nc, Hin, Win, scale, s, map_ctr);
        });
    }
    if (!dry && !cached) {
        taps_.clear();
        taps_["L1"] = L1; taps_["L2"] = L2; taps_["L3"] = L3; taps_["aligned"] = aligned; taps_["fused"] = fused;
        tap_dtype_ = sizeof(T) == 4 ? RVSR_F32 : RVSR_F16;
        launches_ = P.launches;
    }
    return P.rc;
}

ProfEntry *Engine::prof_begin(const std::string &label, double flops, double bytes, cudaStream_t s) {
    if (!profiling_) return nullptr;
    prof_.emplace_back();
    ProfEntry &e = prof_.back();
    e.label = label; e.flops = flops; e.bytes = bytes;
    if (cudaEventCreate(&e.e0) != cudaSuccess || cudaEventCreate(&e.e1) != cudaSuccess) return nullptr;
    cudaEventRecord(e.e0, s);
    return &e;
}
void Engine::prof_end(ProfEntry *e, cudaStream_t s) {
    if (e != nullptr) cudaEventRecord(e->e1, s);
}
void Engine::prof_clear() {
    for (ProfEntry &e : prof_) {
        if (e.e0) cudaEventDestroy(e.e0);
        if (e.e1) cudaEventDestroy(e.e1);
    }
    prof_.clear();
}
int Engine::prof_collect() {
    for (ProfEntry &e : prof_) {
        if (e.e0 == nullptr || e.e1 == nullptr) continue;
        if (cudaEventSynchronize(e.e1) != cudaSuccess || cudaEventElapsedTime(&e.ms, e.e0, e.e1) != cudaSuccess) {
            set_error("profile: event timing failed: %s", cudaGetErrorString(cudaGetLastError()));
            return RVSR_E_CUDA;
        }
    }
    return (int)prof_.size();
}

static int check_dims(const rvsr_edvr_config &c, int B, int H, int W) {
    RVSR_CHECK_ARG(B >= 0 && H > 0 && W > 0, "engine: bad input size B=%d H=%d W=%d", B, H, W);
    const int m = c.HR_in && c.upsample ? 16 : 4;  // two stride-2 levels (+ two more in the HR_in stem)
    RVSR_CHECK_ARG(H % m == 0 && W % m == 0, "engine: H and W must be multiples of %d (got %dx%d)", m, H, W);
    return RVSR_OK;
}

size_t Engine::workspace_bytes(int B, int H, int W) {
    if (check_dims(cfg_, B, H, W) != RVSR_OK) return 0;
    Arena ar;
    if (cfg_.precision == RVSR_F16)
        run<__half>(ar, true, nullptr, RVSR_F32, nullptr, RVSR_F32, B, H, W, nullptr, nullptr);
    else
        run<float>(ar, true, nullptr, RVSR_F32, nullptr, RVSR_F32, B, H, W, nullptr, nullptr);
    return ar.peak + 4096;
}

int Engine::forward(const void *x, int x_dtype, void *out, int out_dtype, int B, int H, int W, void *ws,
                    size_t ws_bytes, cudaStream_t s) {
    if (!finalized_) {
        set_error("engine: weights not finalised");
        return RVSR_E_STATE;
    }
    RVSR_TRY(check_dims(cfg_, B, H, W));
    RVSR_CHECK_ARG(x_dtype == RVSR_F32 || x_dtype == RVSR_F16, "engine: bad x dtype %d", x_dtype);
    RVSR_CHECK_ARG(out_dtype == RVSR_F32 || out_dtype == RVSR_F16, "engine: bad out dtype %d", out_dtype);
    if (B == 0) return RVSR_OK;
    RVSR_CHECK_ARG(x != nullptr && out != nullptr && ws != nullptr, "engine: null buffer");
    prof_clear();
    prof_.reserve(1024);  // entries must not move while a forward holds pointers to them
    Arena ar;
    ar.base = reinterpret_cast<char *>(ws);
    ar.cap = ws_bytes;
    const size_t mis = (size_t)(reinterpret_cast<uintptr_t>(ws) % 1024);
    if (mis) { ar.base += 1024 - mis; ar.cap -= 1024 - mis; }
    host_maps_.clear();
    const int rc = cfg_.precision == RVSR_F16 ? run<__half>(ar, false, x, x_dtype, out, out_dtype, B, H, W, s, nullptr)
                                              : run<float>(ar, false, x, x_dtype, out, out_dtype, B, H, W, s, nullptr);
    tc_stamps_dump();  // debug only (RVSR_TC_STAMPS): synchronises
    return rc;
}

// ---------------------------------------------------------------- sliding-window feature cache
size_t Engine::cache_bytes(int n_slots, int H, int W) const {
    if (n_slots <= 0 || check_dims(cfg_, 0, H, W) != RVSR_OK) return 0;
    if (cfg_.HR_in && cfg_.upsample) { H /= 4; W /= 4; }  // the cache holds features: 1/4 of the HR_in frame size
    const size_t es = cfg_.precision == RVSR_F16 ? 2 : 4, c8 = (size_t)cdiv(cfg_.nf, 8) * 8;
    return (size_t)n_slots * c8 * es * ((size_t)H * W + (size_t)(H / 2) * (W / 2) + (size_t)(H / 4) * (W / 4));
}
size_t Engine::extract_workspace_bytes(int F, int H, int W) {
    if (check_dims(cfg_, F, H, W) != RVSR_OK) return 0;
    Arena ar;
    CacheArgs ca = {};
    ca.extract = true;
    if (cfg_.precision == RVSR_F16)
        run<__half>(ar, true, nullptr, RVSR_F32, nullptr, RVSR_F32, F, H, W, nullptr, &ca);
    else
        run<float>(ar, true, nullptr, RVSR_F32, nullptr, RVSR_F32, F, H, W, nullptr, &ca);
    return ar.peak + 4096;
}
static void arena_over(Arena &ar, void *ws, size_t ws_bytes) {
    ar.base = reinterpret_cast<char *>(ws);
    ar.cap = ws_bytes;
    const size_t mis = (size_t)(reinterpret_cast<uintptr_t>(ws) % 1024);
    if (mis) { ar.base += 1024 - mis; ar.cap -= 1024 - mis; }
}
int Engine::extract_features(const void *frames, int dtype, int F, int H, int W, void *cache, int n_slots, int slot0,
                             void *ws, size_t ws_bytes, cudaStream_t s) {
    if (!finalized_) { set_error("engine: weights not finalised"); return RVSR_E_STATE; }
    RVSR_TRY(check_dims(cfg_, F, H, W));
    RVSR_CHECK_ARG(dtype == RVSR_F32 || dtype == RVSR_F16, "extract_features: bad dtype %d", dtype);
    RVSR_CHECK_ARG(slot0 >= 0 && n_slots > 0 && slot0 + F <= n_slots, "extract_features: slots [%d, %d) outside the cache of %d", slot0, slot0 + F, n_slots);
    if (F == 0) return RVSR_OK;
    RVSR_CHECK_ARG(frames && cache && ws, "extract_features: null buffer");
    Arena ar;
    arena_over(ar, ws, ws_bytes);
    CacheArgs ca = {};
    ca.extract = true; ca.cache = cache; ca.n_slots = n_slots; ca.slot0 = slot0;
    prof_clear();
    if (cfg_.precision == RVSR_F16) return run<__half>(ar, false, frames, dtype, nullptr, RVSR_F32, F, H, W, s, &ca);
    return run<float>(ar, false, frames, dtype, nullptr, RVSR_F32, F, H, W, s, &ca);
}
int Engine::forward_cached(const void *cache, int n_slots, const int *window_slots, const void *frames, int x_dtype, void *out,
                           int out_dtype, int B, int H, int W, void *ws, size_t ws_bytes, cudaStream_t s) {
    if (!finalized_) { set_error("engine: weights not finalised"); return RVSR_E_STATE; }
    RVSR_TRY(check_dims(cfg_, B, H, W));
    RVSR_CHECK_ARG(x_dtype == RVSR_F32 || x_dtype == RVSR_F16, "forward_cached: bad x dtype %d", x_dtype);
    RVSR_CHECK_ARG(out_dtype == RVSR_F32 || out_dtype == RVSR_F16, "forward_cached: bad out dtype %d", out_dtype);
    if (B == 0) return RVSR_OK;
    RVSR_CHECK_ARG(cache && window_slots && frames && out && ws && n_slots > 0, "forward_cached: null buffer");
    for (int i = 0; i < B * cfg_.nframes; ++i)
        RVSR_CHECK_ARG(window_slots[i] >= 0 && window_slots[i] < n_slots, "forward_cached: slot %d outside the cache of %d", window_slots[i], n_slots);
    Arena ar;
    arena_over(ar, ws, ws_bytes);
    CacheArgs ca = {};
    ca.cache = const_cast<void *>(cache); ca.n_slots = n_slots; ca.window_slots = window_slots;
    prof_clear();
    prof_.reserve(1024);
    host_maps_.clear();
    if (cfg_.precision == RVSR_F16) return run<__half>(ar, false, frames, x_dtype, out, out_dtype, B, H, W, s, &ca);
    return run<float>(ar, false, frames, x_dtype, out, out_dtype, B, H, W, s, &ca);
}

int Engine::read_tap(const char *name, float *dst, size_t dst_elems, cudaStream_t s) {
    auto it = taps_.find(name ? name : "");
    RVSR_CHECK_ARG(it != taps_.end(), "engine: unknown tap %s", name ? name : "(null)");
    const Act &a = it->second;
    RVSR_CHECK_ARG(dst_elems >= (size_t)a.N * a.C * a.H * a.W, "engine: tap buffer too small");
    if (tap_dtype_ == RVSR_F32) return launch_unpack_nchw<float, float>((const float *)a.p, dst, a.N, a.C, a.H, a.W, s);
    return launch_unpack_nchw<__half, float>((const __half *)a.p, dst, a.N, a.C, a.H, a.W, s);
}

}  // namespace rvsr
