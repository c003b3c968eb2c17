// engine.cuh -- the EDVR inference engine behind rvsr_engine_* (see include/rvsr_b200.h).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace rvsr {

// Allocator over the caller-provided workspace: best-fit free list with coalescing, so that an activation's memory is
// reused as soon as the plan says it is dead (Plan::drop) -- the peak, not the sum, of the live tensors sizes the
// workspace (round 1's bump allocator needed ~6 GB for a batch of 4 cfg2 windows).  With base == nullptr it only does the
// bookkeeping, which is how rvsr_engine_workspace_bytes() sizes the workspace from the same code path that later runs
// the forward: both runs see the same sequence of alloc / free calls and therefore the same offsets.
// Reuse is safe without extra synchronisation: every kernel of a forward is enqueued on ONE stream, and a kernel only
// touches activations after griddepcontrol.wait, i.e. after every earlier kernel of the stream has completed.
struct Arena {
    struct Block { size_t off, size; bool free; };
    char *base = nullptr;
    size_t cap = 0, peak = 0;
    bool overflow = false;
    std::vector<Block> blocks;  // sorted by offset, contiguous from 0
    void *alloc(size_t bytes) {
        const size_t need = align_up(bytes == 0 ? 1 : bytes, 1024);
        int best = -1;
        for (int i = 0; i < (int)blocks.size(); ++i)
            if (blocks[i].free && blocks[i].size >= need && (best < 0 || blocks[i].size < blocks[best].size)) best = i;
        size_t off;
        if (best >= 0) {
            off = blocks[best].off;
            if (blocks[best].size > need) {
                const Block rest{off + need, blocks[best].size - need, true};
                blocks[best].size = need;
                blocks.insert(blocks.begin() + best + 1, rest);
            }
            blocks[best].free = false;
        } else {
            if (!blocks.empty() && blocks.back().free) {  // grow the free tail instead of leaving a hole
                off = blocks.back().off;
                blocks.back().size = need;
                blocks.back().free = false;
            } else {
                off = blocks.empty() ? 0 : blocks.back().off + blocks.back().size;
                blocks.push_back(Block{off, need, false});
            }
            if (off + need > peak) peak = off + need;
        }
        if (base == nullptr) return reinterpret_cast<void *>(size_t(1024) + off);  // dry run: non-null, offset recoverable
        if (off + need > cap) { overflow = true; return nullptr; }
        return base + off;
    }
    void free(const void *p) {
        if (p == nullptr) return;
        const size_t off = base == nullptr ? reinterpret_cast<size_t>(p) - 1024 : (size_t)(reinterpret_cast<const char *>(p) - base);
        for (int i = 0; i < (int)blocks.size(); ++i) {
            if (blocks[i].off != off || blocks[i].free) continue;
            blocks[i].free = true;
            if (i + 1 < (int)blocks.size() && blocks[i + 1].free) { blocks[i].size += blocks[i + 1].size; blocks.erase(blocks.begin() + i + 1); }
            if (i > 0 && blocks[i - 1].free) { blocks[i - 1].size += blocks[i].size; blocks.erase(blocks.begin() + i); }
            return;
        }
    }
};

// Channel-blocked activation [N][C8][H][W][8].
struct Act {
    void *p = nullptr;
    int N = 0, C = 0, H = 0, W = 0;
    int C8() const { return (C + 7) / 8; }
    long long image_elems() const { return (long long)C8() * H * W * 8; }
    long long elems() const { return image_elems() * N; }
};

struct PackedConv {
    float *w_simt = nullptr;  // [chunks][taps][8][cout_pad] fp32
    void *w_tc = nullptr;     // fp16 UMMA layout (tc_kernels.cu) or null
    void *w_tc2 = nullptr;    // fp16 CTA-pair layout or null
    void *w_tapn = nullptr;   // fp16 taps-in-N layout (conv_last) or null
    void *w_om_stream = nullptr;  // conv_offset_mask only: streamed layout of the fused pack kernel (dcn_fused.cu) or null
    float *bias = nullptr;    // fp32 [Cout]
    int Cout = 0, Cin = 0, ks = 0;
};

struct RawWeight {
    std::vector<int64_t> shape;
    float *dev = nullptr;
    bool set = false;
};

struct ProfEntry {
    std::string label;
    double flops = 0, bytes = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms = 0.f;
};

// Feature-cache modes of Engine::run (see engine.cu)
struct CacheArgs {
    bool extract;             // true: write the pyramid of the given frames into the cache and stop
    void *cache;              // [L1 x n_slots][L2 x n_slots][L3 x n_slots], channel-blocked
    int n_slots, slot0;
    const int *window_slots;  // host, [B * nframes] slot index of every frame of every window
};

class Engine {
  public:
    // per-launch profiling (CUDA events on the launching stream)
    void set_profiling(bool on) { profiling_ = on; }
    ProfEntry *prof_begin(const std::string &label, double flops, double bytes, cudaStream_t s);
    void prof_end(ProfEntry *e, cudaStream_t s);
    int prof_collect();
    void prof_clear();
    const std::vector<ProfEntry> &prof() const { return prof_; }

    explicit Engine(const rvsr_edvr_config &cfg);
    ~Engine();
    int set_weight(const char *name, const float *dev_ptr, const int64_t *shape, int ndim, cudaStream_t s);
    int finalize(cudaStream_t s);
    size_t workspace_bytes(int B, int H, int W);
    int forward(const void *x, int x_dtype, void *out, int out_dtype, int B, int H, int W, void *ws,
                size_t ws_bytes, cudaStream_t s);
    int read_tap(const char *name, float *dst, size_t dst_elems, cudaStream_t s);
    size_t cache_bytes(int n_slots, int H, int W) const;
    size_t extract_workspace_bytes(int F, int H, int W);
    int extract_features(const void *frames, int dtype, int F, int H, int W, void *cache, int n_slots, int slot0, void *ws,
                         size_t ws_bytes, cudaStream_t s);
    int forward_cached(const void *cache, int n_slots, const int *window_slots, const void *frames, int x_dtype, void *out,
                       int out_dtype, int B, int H, int W, void *ws, size_t ws_bytes, cudaStream_t s);
    const std::vector<std::string> &names() const { return names_; }
    int last_launches() const { return launches_; }
    const rvsr_edvr_config &cfg() const { return cfg_; }

  private:
    template <typename T>
    int run(Arena &ar, bool dry, const void *x, int x_dtype, void *out, int out_dtype, int B, int H, int W,
            cudaStream_t s, const CacheArgs *ca);
    void expect(const std::string &name, std::vector<int64_t> shape);
    void expect_conv(const std::string &name, int co, int ci, int k);

    rvsr_edvr_config cfg_;
    std::vector<std::string> names_;
    std::map<std::string, RawWeight> raw_;
    std::map<std::string, PackedConv> packed_;
    std::map<std::string, Act> taps_;
    int tap_dtype_ = RVSR_F32;
    bool finalized_ = false;
    int launches_ = 0;
    std::vector<void *> owned_;  // cudaMalloc'ed buffers
    bool profiling_ = false;
    std::vector<ProfEntry> prof_;
    std::vector<std::vector<int>> host_maps_;
};

// tc_kernels.cu: tcgen05 paths (fp16 storage).  Return RVSR_E_UNSUPPORTED when the shape is
// not covered so the caller can route the op to the CUDA-core kernel instead.
void tc_stamps_dump();  // RVSR_TC_STAMPS=1: print the kernel-boundary timeline of the launches since the last dump
bool tc_conv_supported(const ConvOp &op);
int launch_conv_tc(const ConvOp &op, cudaStream_t s);
size_t tc_conv_weight_bytes(int Cout, int Cin, int ks, int mode = 0);
// mode: 0 plain, 1 pixel-shuffle column order, 2 OUT_OM24 column order (Cout == 27 * dg)
int pack_weight_views(const float *w, int n, const int (*dims)[5], const WeightView *wv, void *const *dst, cudaStream_t s);
int tc_conv_layouts(const ConvOp &op);  // which packed weight layout a stride-1 launch reads: 1 = w_tc, 2 = w_tc2, 0 = not covered
int pack_weight_tc(const float *w_oihw, void *dst, int Cout, int Cin, int ks, int mode, cudaStream_t s, const WeightView *view = nullptr);
size_t tc2_weight_bytes(int Cout, int Cin, int ks, int mode = 0);
int pack_weight_tc2(const float *w_oihw, void *dst, int Cout, int Cin, int ks, int mode, cudaStream_t s, const WeightView *view = nullptr);
size_t tc_tapn_weight_bytes(int Cout, int Cin, int ks);   // conv_last "taps in N" kernel (Cout <= 3)
int pack_weight_tapn(const float *w_oihw, void *dst, int Cout, int Cin, cudaStream_t s);
int launch_conv_tapn(const ConvOp &op, const void *w_tapn, cudaStream_t s);
// dcn_fused.cu: ModulatedDeformConvPack as ONE kernel (offset/mask conv -> TMEM -> gather -> contraction)
struct PackFusedOp {
    Src x;                 // sampled features [N][8][H][W][8] fp16 (optionally through an image -> slot map)
    const void *feat;      // offset features, densely packed [N][8][H][W][8] fp16
    const void *w_om;      // pack_weight_om_stream layout
    const float *bias_om;  // [27 * dg] fp32, reference channel order
    const void *w_dcn2;    // pack_weight_tc2 layout (CTA pair) of the 64 x 64 x 3 x 3 contraction weights
    const float *bias;     // [64] or null
    void *out;
    long long out_image_stride;
    int N, H, W, Cout, dg, act;
};
size_t tc_pack_om_weight_bytes(int Cout, int Cin, int dg);
int pack_weight_om_stream(const float *w_oihw, void *dst, int Cout, int Cin, int dg, cudaStream_t s);
bool tc_pack_fused_supported(const PackFusedOp &op);
int launch_pack_fused(const PackFusedOp &op, cudaStream_t s);
// dcn_bwd_tc.cu: DCN backward on the tensor cores (bf16 tensors, EDVR's nf = 64 shape class)
bool dcn_bwd_tc_supported(int C, int Cout, int kh, int kw, int stride, int pad, int dil, int groups, int dg);
size_t dcn_bwd_tc_workspace_bytes(int B, int H, int W);
int launch_dcn_bwd_tc(const void *input, const void *offset, const void *mask, const void *weight, const void *grad_output,
                      void *grad_input, void *grad_offset, void *grad_mask, void *grad_weight, void *grad_bias, int B, int H, int W,
                      void *workspace, size_t workspace_bytes, cudaStream_t s);
bool tc_dcn_supported(const DcnOp &op);
int launch_dcn_tc(const DcnOp &op, cudaStream_t s);
size_t tc_dcn_weight_bytes(int Cout, int C, int K);
int pack_weight_dcn_tc(const float *w_oihw, void *dst, int Cout, int C, int K, cudaStream_t s);
size_t tc_dcn_half_weight_bytes();  // nf = 128: one input-channel half of the 128 x 128 x 3 x 3 contraction weights
int pack_weight_dcn_tc_half(const float *w_oihw, void *dst, int h, cudaStream_t s);

}  // namespace rvsr
