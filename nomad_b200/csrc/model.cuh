// Host-side model state (device weights in kernel-ready layouts) and the per-call batch plan.
#pragma once
#include <vector>

#include "gemm.cuh"

namespace nb {

static constexpr int CONV_DIM = 512;
static constexpr int EMBED = 768;
static constexpr int FFN = 3072;
static constexpr int HEADS = 12;
static constexpr int HEAD_DIM = 64;
static constexpr int LAYERS = 12;
static constexpr int POS_K = 128;
static constexpr int POS_G = 16;
static constexpr int POS_GC = 48;  // channels per group
static constexpr int EMB = 256;
static constexpr int NSTAT = 65;   // 10 tap sums + 55 tap cross-products of the waveform
static constexpr int STAT_CHUNK = 4096;  // conv0 frames per wave-stats block
static const int CONV_KERNEL[7] = {10, 3, 3, 3, 3, 2, 2};
static const int CONV_STRIDE[7] = {5, 2, 2, 2, 2, 2, 2};

// Per-utterance geometry, one entry per utterance, resident on the device for the kernels that need
// to know where an utterance starts and how much of it is valid.
struct UttMeta {
    long long wav_off;  // first sample in the packed waveform buffer
    int n;              // samples
    int T0;             // valid conv0 frames
    int row0;           // first row in the flat level-0 activation (multiple of 64)
    int rows0;          // rows reserved at level 0 (T0 rounded up to a multiple of 64)
    int T;              // valid frames after the conv encoder (level 6)
    int frame0;         // first row in the flat frame-level buffers = row0 / 64
    int frames;         // rows reserved at frame level = rows0 / 64
    int pos0;           // first row of this utterance in the zero-padded positional-conv layout
};

struct LayerWeights {
    op_t *w_qkv, *w_o, *w_fc1, *w_fc2;          // [N][K] K-major op_t
    float *b_qkv, *b_o, *b_fc1, *b_fc2;
    float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    // transposed copies for the dgrad GEMMs of the loss path
    op_t *wt_qkv, *wt_o, *wt_fc1, *wt_fc2;      // [K][N] -> used as [N'=K][K'=N]
};

struct Weights {
    float* conv0_w;      // [512][10]
    float *gn_g, *gn_b;  // [512]
    op_t* conv_w[7];     // l = 1..6: [512][k*512], K index = tap*512 + cin
    op_t* conv_wt[7];    // dgrad: [k][cin=512][cout=512] -> per tap [N'=cin][K'=cout]
    float *ln0_g, *ln0_b;  // LayerNorm(512)
    op_t* proj_w;        // [768][512]
    op_t* proj_wt;       // [512][768]
    float* proj_b;
    op_t* pos_w;         // [16][48][128*48], K index = tap*48 + cin
    op_t* pos_wt;        // dgrad: [16][48 (cin)][128*48], K index = tap'*48 + cout (taps flipped)
    float* pos_b;        // [768]
    float *lne_g, *lne_b;  // encoder LayerNorm(768)
    LayerWeights layer[LAYERS];
    float* head_wt;      // scoring head, transposed [768][256]
    float* head_b;       // [256]
    float* loss_head_wt; // loss head (LossNetLayers.embedding_layer), transposed [768][256]
    float* loss_head_w;  // [256][768] (backward)
    float* loss_head_b;
};

struct Plan {
    int B = 0;
    std::vector<UttMeta> utt;
    long long total_samples = 0;
    long long rows0 = 0;   // flat rows at level 0 (multiple of 64)
    long long frames = 0;  // flat rows at frame level = rows0 / 64
    long long pos_rows = 0;  // rows of the padded positional-conv layout (outputs)
    int max_T = 0;
    int max_chunks = 0;    // wave-stats chunks of the longest utterance
    bool uniform = false;  // all utterances the same length
};

// Device pointers carved from the caller's workspace for one forward pass.
struct Workspace {
    UttMeta* meta;       // [B]
    double* stat_part;   // [B][max_chunks][65]
    float* c0_fold;      // [B][512][12]: 10 folded taps, shift, pad
    op_t* act_a;         // level 0/2/4/6: (rows0 + 8) x 512
    op_t* act_b;         // level 1/3/5 and LN(512) output: (rows0/2 + 8) x 512
    float* x;            // residual stream, frames x 768 fp32
    op_t* xh;            // op_t copy (GEMM operand)
    float* pre;          // pre-LayerNorm sums, frames x 768 fp32
    op_t* pos_g;         // [16][pos_rows + 128][48]
    op_t* pos_y;         // [pos_rows][768]
    op_t* qkv;           // frames x 2304
    op_t* attn;          // frames x 768
    op_t* ffn_h;         // frames x 3072
    size_t bytes;
};

struct Handle {
    int device = 0;
    int gemm_impl = 0;
    Weights w{};
    std::vector<void*> allocs;  // everything cudaMalloc'ed for the weights
    bool has_loss_head = false;
    UttMeta* meta_host = nullptr;  // pinned staging for the per-call metadata
    int meta_cap = 0;
    cudaEvent_t meta_event = nullptr;  // last use of meta_host by an async copy
};

int make_plan(const int64_t* sample_offsets, int B, Plan* plan);
size_t carve_workspace(const Plan& p, void* base, Workspace* ws);

}  // namespace nb
