// Internal interface of the op_t tensor-core GEMM used by every dense contraction of the path
// (conv layers 1-6 as overlapping-row implicit GEMMs, feature projection, grouped positional conv,
// QKV / out-proj / FFN, their dgrad twins and the split-op_t Gram matrix of the distance kernel).
#pragma once
#include "common.cuh"

namespace nb {

enum : int {
    EPI_BIAS = 1,      // + bias[col]
    EPI_GELU = 2,      // exact erf GELU
    EPI_RESID = 4,     // + resid[row, col] (fp32)
    EPI_OUT_F32 = 8,   // store fp32
    EPI_OUT_H16 = 16, // store op_t
    EPI_MUL_AUX = 32,  // * aux[row, col] (op_t)  -- dgrad through GELU: aux holds gelu'(pre-activation)
    EPI_CDIST = 64,    // out = sqrt(max(na[row] + nb[col] - 2 acc, 0)); fp32 store + per-column-group partial row sums
    EPI_RESID_LN = 256,  // + LayerNorm(resid row) recomputed from saved (mean, rstd): (r - mean) * rstd * g + b
    // LayerNorm of the A operand folded into this GEMM: A holds the UN-normalised rows x, the weights carry gamma
    // (W' = W diag(gamma)), and out = rstd_r * (acc - mean_r * s[col]) + c[col] with s = row sums of W' and
    // c = W beta + bias -- exact algebra, so no LayerNorm kernel has to run between the two GEMMs.
    EPI_LN_FOLD = 512,
    // also emit per-row partial LayerNorm statistics of the fp32 output, one (mean, M2) pair per 64 columns, for
    // a later EPI_LN_FOLD / EPI_RESID_LN consumer (needs N % 256 == 0)
    EPI_STATS_OUT = 1024,
    // fp32-class ("precise") mode, only in the PREC kernel instantiations: operands are hi + lo fp16 planes and the
    // mainloop runs three K segments (A_hi B_hi + A_lo B_hi + A_hi B_lo, ~22 significant bits per product); the
    // epilogue first multiplies the accumulator by acc_scale (weights are stored times a power of two so that their
    // lo plane stays in fp16's normal range), uses libdevice erff for the GELU, and EPI_OUT_H16 stores the result as
    // hi (out_h) + lo (out_l) planes.
    EPI_PRECISE = 2048,
    EPI_SAVE_DGELU = 128,  // with EPI_GELU: also store gelu'(pre-activation) to aux_out (bf16/fp16), for the loss backward
};

// One operand: ``rows`` rows of K op_t values, row r of batch b starting at
// ptr + b * batch_stride + r * row_stride (elements).  row_stride may be SMALLER than K
// (overlapping rows): that is how a strided conv over a channels-last activation becomes a GEMM
// without an im2col buffer.
struct GemmOperand {
    const op_t* ptr;
    long long rows;
    long long row_stride;
    long long batch_stride;
    // 0: K index k of row r lives at r * row_stride + k (rows may overlap).
    // w > 0 ("wrapped K", the non-overlapping spelling of the same thing, needs w % 64 == 0):
    //   it lives at (r + k / w) * row_stride + k % w, with row_stride == w.
    int k_wrap;
    const op_t* lo;  // EPI_PRECISE: the lo plane (same layout as ptr); nullptr otherwise
};

struct GemmEpilogue {
    int flags;
    const float* bias;        // [batch][N]
    long long bias_bstride;
    const float* resid;       // fp32, element (row, col) at resid[row * ldr + col + b * resid_bstride]
    long long ldr;
    long long resid_bstride;
    const float* ln_stats;    // EPI_RESID_LN: [rows][2] = mean, rstd of the resid row (when ln_part is null)
    const float* ln_part;     // EPI_RESID_LN / EPI_LN_FOLD: [rows][LN_PARTS][2] partial (mean, M2) per 64 columns
    float* part_out;          // EPI_STATS_OUT target, same layout
    const float* fold_s;      // EPI_LN_FOLD: [N] row sums of the gamma-folded (rounded) weights
    const float* fold_c;      // EPI_LN_FOLD: [N] W beta + bias
    const float* ln_g;        // [N]
    const float* ln_b;        // [N]
    const op_t* aux;          // 16-bit, same indexing as out (ldo / out_bstride)
    op_t* aux_out;            // EPI_SAVE_DGELU target, same indexing as out
    float* out_f;             // element (row, col) of batch b at out[row * ldo + col + b * out_bstride]
    op_t* out_h;
    long long ldo;
    long long out_bstride;
    // EPI_CDIST: acc = cd_scale^-1 * <a_row, b_col> from split-fp16 operands
    const float* norm_a;      // [M] squared norms of the rows of A
    const float* norm_b;      // [N]
    double* row_sum;          // SIMT check kernel only: [M] += sum over cols of the distances (atomics)
    float* row_part;          // tensor-core kernels: [column groups][M] partial row sums, each written exactly once
                              // (group = 64- or 128-column slice owned by one epilogue warp) -> deterministic means
    const float* cd_a;        // [M][256] original fp32 rows (exact re-evaluation of near-zero distances)
    const float* cd_b;        // [N][256]
    float cd_inv_scale;       // acc * cd_inv_scale = dot product
    float acc_scale;          // EPI_PRECISE: accumulator scale (2^-k of the weight tensor)
    op_t* out_l;              // EPI_PRECISE + EPI_OUT_H16: lo plane of the output, same indexing as out_h
};

// C[b] = epilogue(A[b] (M x K) * B[b]^T (N x K)).  impl: 0 = tcgen05/TMA kernel, 1 = SIMT check kernel.
static constexpr int LN_PARTS = 12;  // 768 / 64

int gemm_h16(cudaStream_t st, const GemmOperand& A, const GemmOperand& B, int M, int N, int K, int batch,
              const GemmEpilogue& epi, int impl);

// C[b] = epilogue(sum_k A[m, k] B_{b % b_mod}[n, k + shift0 + (b / b_mod) * step]): ONE A tensor shared by all batches, b_mod B
// tensors B.batch_stride apart, B's K origin shifted per batch (out-of-range columns read as zero; shifts multiples of 8
// elements: TMA wants 16-byte aligned origins); B.row_stride = B's full row length.  N <= 128.
int gemm_h16_corr(cudaStream_t st, const GemmOperand& A, const GemmOperand& B, int M, int N, int K, int batch, int shift0,
                  int step, int b_mod, const GemmEpilogue& epi);

int device_sm_count();
// column groups (partial row sums per row) the EPI_CDIST tensor-core kernels write for an n x m problem
int cdist_row_groups(long long n, long long m);

// event-pair timing of every tensor-core GEMM launch while enabled (see bench.py)
void gemm_profile_enable(bool on);
bool gemm_profile_active();
int gemm_profile_read(double* total_ms, double* total_flops, long long* launches);

}  // namespace nb
