// op_t x op_t -> fp32 GEMM for sm_100a: TMA-staged operand tiles, tcgen05.mma with the accumulator
// in TMEM, warp-specialised persistent CTAs (1 TMA warp, 1 MMA warp, 8 epilogue warps), double-buffered
// accumulators so the epilogue of tile i overlaps the mainloop of tile i+1.
//
// C[b][m, n] = epilogue( sum_k A[b][m, k] * B[b][n, k] ),  A and B both K-major ("TN" GEMM: activations
// x weight^T, which is what nn.Linear / conv-as-GEMM need).
#include "gemm.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace nb {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 op_t = 128 B = one swizzle-128B row
static constexpr int UMMA_K = 16;
static constexpr int NUM_EPI_WARPS = 8;
static constexpr int GEMM_THREADS = 128 + NUM_EPI_WARPS * 32;
static constexpr float CD_REFINE = 1e-3f;  // squared distances below this are re-evaluated exactly (see cdist_chunk)

template <int BN>
struct TileCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192 ? 4 : (BN == 128 ? 5 : 7));
    static constexpr int ACC_STRIDE = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);  // TMEM columns per accumulator stage
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE;  // two accumulator stages (power of two)
    static constexpr int EPI_STAGE_BYTES = NUM_EPI_WARPS * 4096;  // coalescing buffers of the epilogue warps
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmArgs {
    int M, N, K, batch;
    int umma_n;  // N of the MMA instruction (<= BN, multiple of 16)
    int m_tiles, n_tiles;
    int a_wrap;  // GemmOperand::k_wrap of A
    // "shared operand" batching (single-CTA kernel): every batch reads the SAME A / B tensor (batch coordinate 0) and the
    // B operand's K origin moves by b_kshift0 + batch * b_kshift_step elements (negative / past-the-end columns read as
    // zero): batch = tap of a correlation sum_k A[m, k] B[n, k + tap] (the positional conv's weight gradient)
    int shared_ab, b_kshift0, b_kshift_step, b_mod;  // B batch coordinate = batch % b_mod, K shift = b_kshift0 + (batch / b_mod) * step
    int resid_tma;      // residual epilogues (prefetch variant): fetch the residual tiles with TMA instead of cp.async
    int resid_l2pf;     // residual epilogues: L2-prefetch the next tile's slice of the fp32 residual stream (tiles ahead, 0 = off)
    unsigned sleep_ns;  // back-off of the waiting TMA / MMA role threads (0 = spin); they share schedulers with epilogue warps
    GemmEpilogue epi;
};
__device__ __forceinline__ void mbar_wait_role(uint64_t* bar, uint32_t parity, unsigned ns) {
    if (mbar_try_wait(bar, parity)) return;
    while (!mbar_try_wait(bar, parity))
        if (ns) __nanosleep(ns);
}

// ------------------------------------------------------------------------------------------------
// Staged, coalesced global stores.  TMEM hands every lane one ROW of the tile, so a direct store instruction
// would touch 32 different rows with 16 B each: half-sector writes that L2 has to read-modify-write (measured:
// the N=768, K=768 out-projection ran at 370 TFLOP/s because of it).  Instead each epilogue warp owns a 4 KB
// shared-memory staging buffer (32 rows x 128 B, 16-byte pieces XOR-swizzled): lanes deposit their row, then
// the warp reads it back transposed so that consecutive lanes store consecutive 16 B of the SAME row --
// full 32 B sectors / 128 B lines per instruction.
// Epilogue for one warp x 32-column chunk: lane owns row `row` (v[] = its fp32 accumulators).  Executed by all
// 32 lanes (staging is warp-collective); lanes whose row is past M skip the loads and are never flushed.
// Per-row state an epilogue thread carries across the chunks of its tile.
struct RowCtx {
    float2 st;       // (mean, rstd) of the LayerNorm input row (EPI_LN_FOLD / EPI_RESID_LN)
    float pm, pM2;   // statistics of the previous (even) 32-column chunk (EPI_STATS_OUT)
    // Residual prefetch (8-warp pair kernel): this chunk's fp32 residual tile was requested with cp.async into
    // `rbuf` one chunk (or, for the first chunk of a tile, one mainloop) ago; as soon as it has been consumed the
    // next chunk's tile is requested from `rnext` (nullptr: nothing to prefetch / this chunk was not prefetched).
    float* rbuf;
    bool rhave;
    const float* rnext;
    // the same prefetch as ONE TMA tile load per chunk (box 32 x 32 fp32 in the staging swizzle, completion on this warp's own
    // mbarrier) instead of 8 cp.async per lane: nullptr = cp.async
    const CUtensorMap* tm_r;
    uint64_t* rbar;
    uint32_t rphase;
    const float* rbase;  // origin of the residual matrix (tile coordinates = (rnext - rbase) / ldr, % ldr)
    // Column vectors of this warp's column range (bias or c | s or gamma | beta), staged in shared memory once per
    // tile by epilogue_stage_vectors: the chunk loop then reads them with broadcast LDS instead of one L2 round
    // trip per chunk (the kernel's 200+ KB of shared memory leave almost no L1).  nullptr = read global memory.
    const float* vec;
    int vec_stride;  // floats per staged vector (= columns per epilogue warp)
    int vec_col0;    // first column of the staged range
    // TMA stores (pair kernel, TS instantiations): the staged tile leaves through cp.async.bulk.tensor instead of being
    // read back and stored by the lanes -- one instruction per tile instead of 8 (fp32) / 4 (16-bit) LDS + STG pairs per
    // lane, and ragged edges are clipped by the hardware.  The staging layouts (stage_put_*) ARE the 128-byte / 64-byte
    // TMA swizzles.  `pending` = 2 KB slots of this warp's staging area that an issued store may still be reading.
    const CUtensorMap* tm_f;   // fp32 output map (box 32 x 32, SWIZZLE_128B) or nullptr
    const CUtensorMap* tm_h;   // 16-bit output map (box 32 x 32, SWIZZLE_64B) or nullptr
    const CUtensorMap* tm_a;   // 16-bit aux_out map or nullptr
    uint32_t pending;
    int stage_bytes;           // 4096 or 6144 per warp
    int flip;                  // alternates the 16-bit slot when the area has no room for fp32 + 16-bit side by side
};
// residual tile request through TMA: the warp has finished reading rbuf (the caller's __syncwarp), lane 0 arms the barrier
__device__ __forceinline__ void resid_tma_issue(const CUtensorMap* map, uint64_t* bar, float* rbuf, const float* src,
                                                const float* base, long long ld, int lane) {
    if (lane == 0) {
        const long long off = src - base;
        const int row = (int)(off / ld), col = (int)(off - (long long)row * ld);
        mbar_expect_tx(bar, 4096);
        tma_load_2d_cta(rbuf, map, bar, col, row);
    }
}
// wait until the slots in `mask` are free again (warp-uniform), then mark them as about to be in flight
__device__ __forceinline__ void ts_acquire(RowCtx& rc, uint32_t mask, int lane) {
    if (rc.pending & mask) {
        if (lane == 0) bulk_wait_read();
        rc.pending = 0;
    }
    __syncwarp();
    rc.pending |= mask;
}
__device__ __forceinline__ void ts_issue(const CUtensorMap* map, const void* src, int col, long long row, int lane) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(map, src, col, (int)row);
        bulk_commit();
    }
}
// byte offset of the 16-bit slot for this store (fp32 owns [0, 4096))
__device__ __forceinline__ int ts_h16_slot(RowCtx& rc, bool f32_too) {
    if (rc.stage_bytes >= 6144) return 4096;
    if (f32_too) return 0;  // shares the fp32 area (rare combination in the 16-warp variants): serialised by ts_acquire
    rc.flip ^= 1;
    return rc.flip * 2048;
}
// (mean, rstd) of a 768-wide row from its LN_PARTS partial (mean, M2) pairs of 64 columns each (Chan et al.)
__device__ __forceinline__ float2 ln_row_stats(const float* part, long long row) {
    const float4* p = reinterpret_cast<const float4*>(part + row * (2 * LN_PARTS));
    float m[LN_PARTS], M2 = 0.f, mean = 0.f;
#pragma unroll
    for (int i = 0; i < LN_PARTS / 2; ++i) {
        const float4 t = __ldg(p + i);
        m[2 * i] = t.x; m[2 * i + 1] = t.z;
        M2 += t.y + t.w;
        mean += t.x + t.z;
    }
    mean *= 1.0f / LN_PARTS;
    float d2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_PARTS; ++i) d2 = fmaf(m[i] - mean, m[i] - mean, d2);
    const float var = fmaf(64.0f, d2, M2) * (1.0f / (64.0f * LN_PARTS));
    return make_float2(mean, rsqrtf(var + 1e-5f));
}

// (mean, rstd) of this thread's row for EPI_LN_FOLD / EPI_RESID_LN.  Called BEFORE the wait for the accumulator so
// the (L2-latency) loads overlap the mainloop of the tile instead of sitting at the head of its epilogue.
// EF: the epilogue flags as a compile-time constant (-1 = read them from the arguments).  The hot flag
// combinations of the scoring path get their own instantiation: no flag tests, no dead paths holding registers.
template <int EF>
__device__ __forceinline__ int epi_flags(const GemmEpilogue& e) { return EF >= 0 ? EF : e.flags; }

template <int EF>
__device__ __forceinline__ float2 epilogue_row_stats(const GemmArgs& args, long long row) {
    if (row >= args.M || !(epi_flags<EF>(args.epi) & (EPI_LN_FOLD | EPI_RESID_LN))) return make_float2(0.f, 1.f);
    return args.epi.ln_part != nullptr ? ln_row_stats(args.epi.ln_part, row)
                                       : __ldg(reinterpret_cast<const float2*>(args.epi.ln_stats) + row);
}

// Stage the column vectors of columns [c0, c0 + W) for one epilogue warp: slot 0 = bias (or fold_c), slot 1 = fold_s
// (or ln_g), slot 2 = ln_b.  Called once per tile, before the wait for the accumulator.  Only complete ranges are
// staged (the caller falls back to global loads otherwise).
template <int EF, int W>
__device__ __forceinline__ void epilogue_stage_vectors(const GemmEpilogue& e, float* vec, int c0, int lane) {
    const int flags = epi_flags<EF>(e);
    __syncwarp();  // the previous tile's chunk loop is done reading
    if (lane * 4 < W) {
        if (flags & EPI_BIAS) reinterpret_cast<float4*>(vec)[lane] = __ldg(reinterpret_cast<const float4*>(e.bias + c0) + lane);
        if (flags & EPI_LN_FOLD) {
            reinterpret_cast<float4*>(vec)[lane] = __ldg(reinterpret_cast<const float4*>(e.fold_c + c0) + lane);
            reinterpret_cast<float4*>(vec + W)[lane] = __ldg(reinterpret_cast<const float4*>(e.fold_s + c0) + lane);
        }
        if (flags & EPI_RESID_LN) {
            reinterpret_cast<float4*>(vec + W)[lane] = __ldg(reinterpret_cast<const float4*>(e.ln_g + c0) + lane);
            reinterpret_cast<float4*>(vec + 2 * W)[lane] = __ldg(reinterpret_cast<const float4*>(e.ln_b + c0) + lane);
        }
    }
    __syncwarp();
}

// FULL: all 32 rows of the warp are valid and the chunk has all 32 columns (compile-time: no predicates)
// SV / RP (flag-specialised pair kernels, whose launcher guarantees N % 256 == 0 and one batch): the column vectors are
// always staged / every FULL chunk's residual tile was prefetched -- as compile-time facts, so the global-load twins of
// those reads are not even issued predicated-off (they were 72 of ~860 issue slots per chunk of a latency-bound warp).
template <bool FULL, int EF, bool PREC = false, bool TS = false, bool SV = false, bool RP = false>
__device__ __forceinline__ void epilogue_chunk(const GemmEpilogue& e, float (&v)[32], long long row, bool row_ok,
                                               int rows_valid, int col0, int ncols, int b, float* stage, int lane,
                                               RowCtx& rc) {
    const int flags = epi_flags<EF>(e);
    const long long off = row * e.ldo + col0 + (long long)b * e.out_bstride;
    const long long woff = (row - lane) * e.ldo + col0 + (long long)b * e.out_bstride;  // this warp's first row
    if (PREC) {
        const float sc = e.acc_scale;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= sc;
    }
    if ((FULL || row_ok) && (flags & EPI_BIAS)) {
        const float4* bp = reinterpret_cast<const float4*>(e.bias + (long long)b * e.bias_bstride + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (FULL || j * 4 < ncols) {
                const float4 t = (FULL && (SV || rc.vec)) ? *reinterpret_cast<const float4*>(rc.vec + (col0 - rc.vec_col0) + 4 * j)
                                                          : __ldg(bp + j);
                v[4 * j + 0] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
            }
        }
    }
    if ((FULL || row_ok) && (flags & EPI_LN_FOLD)) {
        const float4* sp = reinterpret_cast<const float4*>(e.fold_s + col0);
        const float4* cp = reinterpret_cast<const float4*>(e.fold_c + col0);
        const float nm = -rc.st.x, rs = rc.st.y;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (FULL || j * 4 < ncols) {
                const bool sv = FULL && (SV || rc.vec != nullptr);
                const float4 a = sv ? *reinterpret_cast<const float4*>(rc.vec + rc.vec_stride + (col0 - rc.vec_col0) + 4 * j)
                                    : __ldg(sp + j);
                const float4 c = sv ? *reinterpret_cast<const float4*>(rc.vec + (col0 - rc.vec_col0) + 4 * j) : __ldg(cp + j);
#if NB_F32X2
                const float2 nm2 = make_float2(nm, nm), rs2 = make_float2(rs, rs);
                const float2 lo = ffma2(rs2, ffma2(nm2, make_float2(a.x, a.y), make_float2(v[4 * j + 0], v[4 * j + 1])), make_float2(c.x, c.y));
                const float2 hi = ffma2(rs2, ffma2(nm2, make_float2(a.z, a.w), make_float2(v[4 * j + 2], v[4 * j + 3])), make_float2(c.z, c.w));
                v[4 * j + 0] = lo.x; v[4 * j + 1] = lo.y; v[4 * j + 2] = hi.x; v[4 * j + 3] = hi.y;
#else
                v[4 * j + 0] = fmaf(rs, fmaf(nm, a.x, v[4 * j + 0]), c.x);
                v[4 * j + 1] = fmaf(rs, fmaf(nm, a.y, v[4 * j + 1]), c.y);
                v[4 * j + 2] = fmaf(rs, fmaf(nm, a.z, v[4 * j + 2]), c.z);
                v[4 * j + 3] = fmaf(rs, fmaf(nm, a.w, v[4 * j + 3]), c.w);
#endif
            }
        }
    }
    if (flags & EPI_SAVE_DGELU) {
        float g[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf_with_grad(v[j], g[j]);
        if (TS && rc.tm_a != nullptr) {
            const int slot = ts_h16_slot(rc, false);
            ts_acquire(rc, 1u << (slot >> 11), lane);
            op_t* sp = reinterpret_cast<op_t*>(reinterpret_cast<char*>(stage) + slot);
            stage_put_h16(sp, g, lane);
            ts_issue(rc.tm_a, sp, col0, row - lane, lane);
        } else {
            stage_put_h16(reinterpret_cast<op_t*>(stage), g, lane);
            stage_flush_h16(reinterpret_cast<op_t*>(stage), e.aux_out + woff, e.ldo, rows_valid, ncols, lane);
        }
    } else if (flags & EPI_GELU) {
        if (PREC || !NB_GELU_FAST) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = PREC ? gelu_erf_exact(v[j]) : gelu_act(v[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float2 t = gelu_fast2(make_float2(v[j], v[j + 1]));
                v[j] = t.x;
                v[j + 1] = t.y;
            }
        }
    }
    if ((FULL || row_ok) && (flags & EPI_MUL_AUX)) {
        const uint4* ap = reinterpret_cast<const uint4*>(e.aux + off);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (FULL || j * 8 < ncols) {
                const uint4 t = __ldg(ap + j);
                float2 f;
                f = unpack_op(t.x); v[8 * j + 0] *= f.x; v[8 * j + 1] *= f.y;
                f = unpack_op(t.y); v[8 * j + 2] *= f.x; v[8 * j + 3] *= f.y;
                f = unpack_op(t.z); v[8 * j + 4] *= f.x; v[8 * j + 5] *= f.y;
                f = unpack_op(t.w); v[8 * j + 6] *= f.x; v[8 * j + 7] *= f.y;
            }
        }
    }
    const bool rpf = FULL && (RP || rc.rhave);  // warp-uniform
    if (rpf) {
        if (rc.tm_r != nullptr) {
            mbar_wait(rc.rbar, rc.rphase);
            rc.rphase ^= 1;
        } else {
            stage_fill_wait();
        }
    }
    if ((FULL || row_ok) && (flags & EPI_RESID)) {
        const float4* rp =
            reinterpret_cast<const float4*>(e.resid + row * e.ldr + col0 + (long long)b * e.resid_bstride);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (FULL || j * 4 < ncols) {
                const float4 t = rpf ? stage_row_f32(rc.rbuf, lane, j) : __ldg(rp + j);
                v[4 * j + 0] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
            }
        }
    }
    if ((FULL || row_ok) && (flags & EPI_RESID_LN)) {
        // residual = LayerNorm(pre-LN row) rebuilt from its saved statistics: saves the fp32 write + read of the
        // normalised residual stream (the LayerNorm kernel then only emits the 16-bit GEMM operand)
        const float4* rp = reinterpret_cast<const float4*>(e.resid + row * e.ldr + col0);
        const float2 st = rc.st;
        const float4* gp = reinterpret_cast<const float4*>(e.ln_g + col0);
        const float4* bp = reinterpret_cast<const float4*>(e.ln_b + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (FULL || j * 4 < ncols) {
                const bool sv = FULL && (SV || rc.vec != nullptr);
                const float4 t = rpf ? stage_row_f32(rc.rbuf, lane, j) : __ldg(rp + j);
                const float4 g = sv ? *reinterpret_cast<const float4*>(rc.vec + rc.vec_stride + (col0 - rc.vec_col0) + 4 * j)
                                    : __ldg(gp + j);
                const float4 bb = sv ? *reinterpret_cast<const float4*>(rc.vec + 2 * rc.vec_stride + (col0 - rc.vec_col0) + 4 * j)
                                     : __ldg(bp + j);
                v[4 * j + 0] += fmaf((t.x - st.x) * st.y, g.x, bb.x);
                v[4 * j + 1] += fmaf((t.y - st.x) * st.y, g.y, bb.y);
                v[4 * j + 2] += fmaf((t.z - st.x) * st.y, g.z, bb.z);
                v[4 * j + 3] += fmaf((t.w - st.x) * st.y, g.w, bb.w);
            }
        }
    }
    if (rc.rbuf != nullptr) {  // warp-uniform
        if (rpf) __syncwarp();  // every lane has read its row
        rc.rhave = rc.rnext != nullptr;
        if (rc.rhave) {
            if (rc.tm_r != nullptr) resid_tma_issue(rc.tm_r, rc.rbar, rc.rbuf, rc.rnext, rc.rbase, e.ldr, lane);
            else stage_fill_f32_async(rc.rbuf, rc.rnext, e.ldr, lane);
        }
    }
    if (flags & EPI_STATS_OUT) {  // two-pass statistics of this chunk, merged pairwise into 64-column partials
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) s4[j & 3] += v[j];
        const float mc = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / 32.0f);
        float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) q4[j & 3] = fmaf(v[j] - mc, v[j] - mc, q4[j & 3]);
        const float M2c = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        if ((col0 & 32) == 0) {
            rc.pm = mc;
            rc.pM2 = M2c;
        } else if (FULL || row_ok) {
            const float d = rc.pm - mc;
            reinterpret_cast<float2*>(e.part_out)[row * LN_PARTS + (col0 >> 6)] =
                make_float2(0.5f * (rc.pm + mc), rc.pM2 + M2c + 16.0f * d * d);
        }
    }
    if (TS && rc.tm_f != nullptr && (flags & EPI_OUT_F32)) {
        ts_acquire(rc, 3u, lane);
        stage_put_f32(stage, v, lane);
        ts_issue(rc.tm_f, stage, col0, row - lane, lane);
    } else if (flags & EPI_OUT_F32) {
        stage_put_f32(stage, v, lane);
        stage_flush_f32(stage, e.out_f + woff, e.ldo, rows_valid, ncols, lane);
    }
    if (TS && rc.tm_h != nullptr && (flags & EPI_OUT_H16)) {
        const int slot = ts_h16_slot(rc, (flags & EPI_OUT_F32) != 0);
        ts_acquire(rc, 1u << (slot >> 11), lane);
        op_t* sp = reinterpret_cast<op_t*>(reinterpret_cast<char*>(stage) + slot);
        stage_put_h16(sp, v, lane);
        ts_issue(rc.tm_h, sp, col0, row - lane, lane);
    } else if (flags & EPI_OUT_H16) {
        stage_put_h16(reinterpret_cast<op_t*>(stage), v, lane);
        stage_flush_h16(reinterpret_cast<op_t*>(stage), e.out_h + woff, e.ldo, rows_valid, ncols, lane);
        if (PREC) {  // lo plane: what the fp16 rounding of the hi plane left over
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] -= op2f(f2op(v[j]));
            stage_put_h16(reinterpret_cast<op_t*>(stage), v, lane);
            stage_flush_h16(reinterpret_cast<op_t*>(stage), e.out_l + woff, e.ldo, rows_valid, ncols, lane);
        }
    }
}

// Scalar twin of the above (SIMT check kernel).
__device__ __forceinline__ void epilogue_scalar(const GemmEpilogue& e, float v, long long row, int col, int b) {
    const int flags = e.flags;
    if (flags & EPI_CDIST) {
        float d2 = e.norm_a[row] + e.norm_b[col] - 2.0f * (v * e.cd_inv_scale);
        if (d2 < CD_REFINE) {
            float s = 0.f;
            for (int k = 0; k < 256; ++k) {
                const float t = e.cd_a[row * 256 + k] - e.cd_b[(long long)col * 256 + k];
                s = fmaf(t, t, s);
            }
            d2 = s;
        }
        const float d = sqrtf(fmaxf(d2, 0.0f));
        if (e.out_f) e.out_f[row * e.ldo + col] = d;
        atomicAdd(e.row_sum + row, (double)d);
        return;
    }
    if (flags & EPI_BIAS) v += e.bias[(long long)b * e.bias_bstride + col];
    const long long off = row * e.ldo + col + (long long)b * e.out_bstride;
    if (flags & EPI_SAVE_DGELU) {
        float g;
        v = gelu_erf_with_grad(v, g);
        e.aux_out[off] = f2op(g);
    } else if (flags & EPI_GELU) {
        v = gelu_erf(v);
    }
    if (flags & EPI_MUL_AUX) v *= op2f(e.aux[off]);
    if (flags & EPI_RESID) v += e.resid[row * e.ldr + col + (long long)b * e.resid_bstride];
    if (flags & EPI_RESID_LN) {
        const float mean = e.ln_stats[2 * row], rstd = e.ln_stats[2 * row + 1];
        v += fmaf((e.resid[row * e.ldr + col] - mean) * rstd, e.ln_g[col], e.ln_b[col]);
    }
    if (flags & EPI_OUT_F32) e.out_f[off] = v;
    if (flags & EPI_OUT_H16) e.out_h[off] = f2op(v);
}

// ------------------------------------------------------------------------------------------------
// Distance epilogue: d = sqrt(max(|a|^2 + |b|^2 - 2 a.b, 0)).  The Gram form cancels catastrophically for
// near-duplicate rows, so squared distances below CD_REFINE are re-evaluated with fp32 direct differences
// (rare); above it the fp32 Gram error (~2e-7 on d^2) keeps d within 3e-6 of scipy's float64 result.

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;  // MUFU.SQRT: relative error ~2^-22, i.e. < 5e-7 on distances <= 2 (the 1e-5 bound has room for it)
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// exact fp32 re-evaluation of one squared distance (direct differences)
__device__ __noinline__ float cdist_exact_d2(const float* __restrict__ a, const float* __restrict__ b) {
    const float4* pa = reinterpret_cast<const float4*>(a);
    const float4* pb = reinterpret_cast<const float4*>(b);
    float s = 0.f;
    for (int k = 0; k < 64; ++k) {
        const float4 x = __ldg(pa + k), y = __ldg(pb + k);
        const float d0 = x.x - y.x, d1 = x.y - y.y, d3 = x.z - y.z, d4 = x.w - y.w;
        s = fmaf(d0, d0, s); s = fmaf(d1, d1, s); s = fmaf(d3, d3, s); s = fmaf(d4, d4, s);
    }
    return s;
}

__device__ __forceinline__ double cdist_chunk(const GemmEpilogue& e, float (&v)[32], long long row, bool row_ok,
                                              int rows_valid, int col0, int ncols, float* stage, int lane,
                                              RowCtx* ts = nullptr) {
    const float na = row_ok ? __ldg(e.norm_a + row) : 0.f;
    float s32 = 0.f;  // 32 distances <= 2 each: an fp32 partial sum is exact to ~1e-7; fp64 only across chunks
    const bool full = ncols == 32;  // warp-uniform
    float nbv[32];
    if (full) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(e.norm_b + col0) + j);  // col0 % 32 == 0
            nbv[4 * j] = t.x; nbv[4 * j + 1] = t.y; nbv[4 * j + 2] = t.z; nbv[4 * j + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) nbv[j] = j < ncols ? __ldg(e.norm_b + col0 + j) : 0.f;
    }
    // squared distances of the whole chunk first; ONE test decides whether any of them needs the exact path
    const float m2s = -2.0f * e.cd_inv_scale;
    float dmin = CD_REFINE;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        v[j] = fmaf(v[j], m2s, na + nbv[j]);
        if (full || j < ncols) dmin = fminf(dmin, v[j]);
    }
    if (row_ok && dmin < CD_REFINE) {  // rare: near-duplicate rows
#pragma unroll
        for (int j = 0; j < 32; ++j)  // static indices keep v[] in registers
            if (j < ncols && v[j] < CD_REFINE) v[j] = cdist_exact_d2(e.cd_a + row * 256, e.cd_b + (long long)(col0 + j) * 256);
    }
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        v[j] = sqrt_approx(fmaxf(v[j], 0.f));
        if (full || j < ncols) s4[j & 3] += v[j];
    }
    s32 = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    const double rs = row_ok ? (double)s32 : 0.0;
    if (e.out_f != nullptr && ts != nullptr && ts->tm_f != nullptr) {  // warp-uniform: tile leaves through TMA
        // two 4 KB halves of the staging area alternate, so the store of chunk c overlaps the math of chunk c + 1
        ts->flip ^= 1;
        float* sp = stage + (ts->stage_bytes >= 8192 ? ts->flip * 1024 : 0);
        ts_acquire(*ts, ts->stage_bytes >= 8192 ? (3u << (2 * ts->flip)) : 3u, lane);
        stage_put_f32(sp, v, lane);
        ts_issue(ts->tm_f, sp, col0, row - lane, lane);
    } else if (e.out_f != nullptr) {  // warp-uniform
        stage_put_f32(stage, v, lane);
        float* wout = e.out_f + (row - lane) * e.ldo + col0;
        if (full && (e.ldo & 3) == 0 && ((reinterpret_cast<uintptr_t>(wout) & 15) == 0)) {
            stage_flush_f32(stage, wout, e.ldo, rows_valid, ncols, lane);
        } else {  // ragged / unaligned: one row per instruction, lane = column (still full-line coalesced)
            __syncwarp();
            for (int rl = 0; rl < 32 && rl < rows_valid; ++rl)
                if (lane < ncols) wout[rl * e.ldo + lane] = stage[rl * 32 + ((((lane >> 2) ^ (rl & 7)) << 2) | (lane & 3))];
            __syncwarp();
        }
    }
    return rs;
}

// ------------------------------------------------------------------------------------------------
// Epilogue of one warp for one tile: CHUNKS x (32 lanes x 32 columns).  The accumulator stage is handed back
// to the MMA warp as soon as the last TMEM load has landed (before that chunk's math and stores).
// (Double-buffering the TMEM loads across chunks was measured SLOWER: 168 registers and less ILP in the GELU.)
// PREC (single-CTA kernel only): the tile's sum lives in FOUR TMEM accumulators PACC columns apart (see gemm_tc_kernel);
// they are added here in IEEE fp32.
struct StoreMaps {
    const CUtensorMap *f, *h, *a;
    const CUtensorMap* r;  // residual tile loads through TMA (nullptr = cp.async)
    uint64_t* rbar;
    uint32_t* rphase;      // carried across the tiles of a warp
    int stage_bytes;
    uint32_t* pending;  // carried across the tiles of a warp
    int* flip;
};

template <int CHUNKS, bool PAIR, bool CDIST, int EF, bool PREC = false, int PACC = 0, bool TS = false, bool SV = false, bool RP = false>
__device__ __forceinline__ void epilogue_tile(const GemmArgs& args, uint32_t taddr, long long row, int col_first,
                                              int ccol_first, int b, uint64_t* tmem_empty_bar, int lane, float* stage,
                                              float2 row_st, float* rbuf, bool& rhave, const float* next_tile_src,
                                              const float* vec = nullptr, int vec_stride = 0, int cd_group = 0,
                                              const StoreMaps* sm = nullptr) {
    const bool row_ok = row < args.M;
    const long long rv = (long long)args.M - (row - lane);
    const int rows_valid = rv > 32 ? 32 : (rv < 0 ? 0 : (int)rv);
    double row_sum = 0.0;
    RowCtx rc;
    rc.st = row_st;
    rc.pm = rc.pM2 = 0.f;
    rc.rbuf = rbuf;
    rc.rhave = rhave;
    rc.rnext = nullptr;
    rc.vec = vec;
    rc.vec_stride = vec_stride;
    rc.vec_col0 = col_first;
    rc.tm_f = rc.tm_h = rc.tm_a = nullptr;
    rc.tm_r = nullptr; rc.rbar = nullptr; rc.rphase = 0; rc.rbase = args.epi.resid;
    if (sm != nullptr && sm->r != nullptr) { rc.tm_r = sm->r; rc.rbar = sm->rbar; rc.rphase = *sm->rphase; }
    rc.pending = 0;
    rc.stage_bytes = 4096;
    rc.flip = 0;
    if (TS && sm != nullptr) {
        rc.tm_f = sm->f; rc.tm_h = sm->h; rc.tm_a = sm->a;
        rc.stage_bytes = sm->stage_bytes;
        rc.pending = *sm->pending;
        rc.flip = *sm->flip;
    }
#pragma unroll 1
    for (int c = 0; c < CHUNKS; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (PREC && PACC > 0) {
#pragma unroll 1
            for (int a = 1; a < 4; ++a) {
                uint32_t q[32];
                tmem_ld_32x32(taddr + a * PACC + c * 32, q);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(q[j]));
            }
        }
        if (c == CHUNKS - 1) {  // every TMEM read of this tile has landed: release the accumulator stage now
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(tmem_empty_bar, 0); else mbar_arrive(tmem_empty_bar);
            }
        }
        const int col0 = col_first + c * 32;
        int ncols = args.N - col0;
        ncols = ncols > 32 ? 32 : ncols;
        if (ccol_first + c * 32 >= args.umma_n) ncols = 0;
        if (!CDIST && rbuf != nullptr) {  // where the residual tile of the NEXT chunk (or tile) comes from
            rc.rnext = next_tile_src;
            if (c + 1 < CHUNKS) {
                const int col1 = col0 + 32;
                rc.rnext = (rows_valid == 32 && col1 + 32 <= args.N && ccol_first + (c + 1) * 32 + 32 <= args.umma_n)
                               ? args.epi.resid + (row - lane) * args.epi.ldr + col1
                               : nullptr;
            }
        }
        if (ncols > 0 && rows_valid > 0) {  // warp-uniform
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (CDIST) row_sum += cdist_chunk(args.epi, v, row, row_ok, rows_valid, col0, ncols, stage, lane, TS ? &rc : nullptr);
            else if (rows_valid == 32 && ncols == 32) epilogue_chunk<true, EF, PREC, TS, SV, RP>(args.epi, v, row, true, 32, col0, 32, b, stage, lane, rc);
            else epilogue_chunk<false, EF, PREC, TS>(args.epi, v, row, row_ok, rows_valid, col0, ncols, b, stage, lane, rc);
        }
    }
    rhave = rc.rhave;
    if (sm != nullptr && sm->r != nullptr) *sm->rphase = rc.rphase;
    if (TS && sm != nullptr) {
        *sm->pending = rc.pending;
        *sm->flip = rc.flip;
    }
    // one writer per (column group, row): the row means come out bit-identical from run to run (no atomics)
    if (CDIST && row_ok) args.epi.row_part[(long long)cd_group * args.M + row] = (float)row_sum;
}

// ------------------------------------------------------------------------------------------------
// PREC: operands are hi + lo planes (tmA / tmB = hi, tmA2 / tmB2 = lo) and the K loop runs three segments,
// A_hi B_hi, A_lo B_hi, A_hi B_lo, into the same accumulator (see EPI_PRECISE).
template <int BN, bool CDIST, bool PREC = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const GemmArgs args) {
    using Cfg = TileCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic on the __shared__ array keeps the address space (LDS / STS, not generic LD / ST)
    float* epi_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = args.m_tiles * args.n_tiles * args.batch;
    const int k_blocks = (args.K + BK - 1) / BK;
    const int k_total = PREC ? 3 * k_blocks : k_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], NUM_EPI_WARPS);
        }
        mbar_fence_init();
    }
    // PREC: four accumulators per tile and no double buffering.  The tensor core truncates its fp32 accumulator at
    // every instruction (measured: relative bias -2^-25 per k16 step, profiles/r02_precise_probe.log), so a K = 768
    // split GEMM accumulated in one chain of 144 steps is 5-10x less accurate than IEEE fp32.  The hi x hi segment is
    // therefore spread over three accumulators (k-block mod 3) and the two small cross terms go to a fourth, which
    // cuts the longest chain 9x; the epilogue adds the four in IEEE fp32.
    constexpr int TCOLS = PREC ? 4 * Cfg::ACC_STRIDE : Cfg::TMEM_COLS;
    if (warp == 2) {
        tmem_alloc(tmem_slot, TCOLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = Cfg::A_BYTES + (uint32_t)args.umma_n * BK * 2;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n_blk = tile % args.n_tiles;
                const int m_blk = (tile / args.n_tiles) % args.m_tiles;
                const int b = tile / (args.n_tiles * args.m_tiles);
                for (int kq = 0; kq < k_total; ++kq) {
                    const int seg = PREC ? kq / k_blocks : 0, kb = PREC ? kq - seg * k_blocks : kq;
                    const CUtensorMap* mA = (PREC && seg == 1) ? &tmA2 : &tmA;
                    const CUtensorMap* mB = (PREC && seg == 2) ? &tmB2 : &tmB;
                    mbar_wait_role(&empty_bar[stage], phase ^ 1, args.sleep_ns);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    mbar_expect_tx(&full_bar[stage], tx_bytes);
                    if (args.a_wrap > 0) {
                        const int kk = kb * BK;
                        tma_load_3d(sa, mA, &full_bar[stage], kk % args.a_wrap, m_blk * BM + kk / args.a_wrap, b);
                    } else {
                        tma_load_3d(sa, mA, &full_bar[stage], kb * BK, m_blk * BM, args.shared_ab ? 0 : b);
                    }
                    tma_load_3d(sb, mB, &full_bar[stage],
                                kb * BK + (args.shared_ab ? args.b_kshift0 + (b / args.b_mod) * args.b_kshift_step : 0), n_blk * BN,
                                args.shared_ab ? b % args.b_mod : b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_h16(BM, args.umma_n);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = PREC ? 0 : (it & 1);
                const uint32_t aphase = PREC ? (it & 1) : ((it >> 1) & 1);
                mbar_wait_role(&tmem_empty[as], aphase ^ 1, args.sleep_ns);
                tc_fence_after();
                uint32_t started = 0;  // PREC: accumulators that already hold a partial sum
                for (int kb = 0; kb < k_total; ++kb) {
                    mbar_wait_role(&full_bar[stage], phase, args.sleep_ns);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + Cfg::A_BYTES;
                    const uint64_t da = umma_desc_sw128(sa);
                    const uint64_t db = umma_desc_sw128(sb);
                    const int acc = PREC ? (kb < k_blocks ? kb % 3 : 3) : as;
                    const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
                    const uint32_t cont = PREC ? ((started >> acc) & 1u) : (kb ? 1u : 0u);
                    started |= 1u << acc;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advancing K by 16 op_t = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
                        umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (cont | k) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above retire
                    if (kb == k_total - 1) umma_commit(&tmem_full[as]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> registers -> global =====================
        const int ew = warp - 4;
        const int q = warp & 3;          // TMEM lane quarter this warp may access
        const int h = ew >> 2;           // column half
        constexpr int HALF = BN / 2;
        constexpr int CHUNKS = (HALF + 31) / 32;
        bool no_prefetch = false;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int n_blk = tile % args.n_tiles;
            const int m_blk = (tile / args.n_tiles) % args.m_tiles;
            const int b = tile / (args.n_tiles * args.m_tiles);
            const int as = PREC ? 0 : (it & 1);
            const uint32_t aphase = PREC ? (it & 1) : ((it >> 1) & 1);
            const long long row = (long long)m_blk * BM + q * 32 + lane;
            const float2 row_st = CDIST ? make_float2(0.f, 1.f) : epilogue_row_stats<-1>(args, row);
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            epilogue_tile<CHUNKS, false, CDIST, -1, PREC, Cfg::ACC_STRIDE>(args, tmem_base + (uint32_t)(as * Cfg::ACC_STRIDE + h * HALF) + ((uint32_t)(q * 32) << 16),
                                         row, n_blk * BN + h * HALF, h * HALF, b, &tmem_empty[as], lane, epi_stage + ew * 1024,
                                         row_st, nullptr, no_prefetch, nullptr, nullptr, 0, n_blk * 2 + h);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TCOLS);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster share one 256 x 256 output tile.  Each CTA
// stages its own 128 rows of A and HALF of the B tile (128 rows); the pair's tensor cores read both halves,
// so per SM the TMA traffic into shared memory drops from 48 to 32 KB per K-block and the B operand is read
// once per pair.  The single-CTA kernel is shared-memory-port bound at ~2/3 of the tensor peak (96 B/clk of
// operand reads + 96 B/clk of TMA writes against a 128 B/clk port); the pair halves the B share of both.
// The even CTA (leader) issues every MMA and owns the full / tmem_empty barriers; commits are multicast.
struct Pair256 {
    static constexpr int BN = 256;
    static constexpr int A_BYTES = BM * BK * 2;          // 16 KB: this CTA's 128 rows of A
    static constexpr int B_BYTES = (BN / 2) * BK * 2;    // 16 KB: this CTA's half of B
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = 5;
    // + one 4 KB residual-prefetch buffer per epilogue warp in the RPF instantiation (8 warps only: the 16-warp
    // variant has no room).  It is a separate instantiation so that GEMMs without a residual keep the smaller
    // shared-memory carve-out (and with it 60 KB instead of 28 KB of L1 for their epilogue's vector loads).
    static constexpr int resid_bytes(int epi_warps, bool rpf) { return rpf ? epi_warps * 4096 : 0; }
    // three staged column vectors per epilogue warp (256 / (warps / 4) columns each)
    static constexpr int vec_bytes(int epi_warps) { return epi_warps * 3 * (BN / (epi_warps / 4)) * 4; }
    // the mainloop is insensitive to 4 vs 5 vs 6 stages (measured, profiles/r01_gemm_probe_stages.log): 4 leaves room
    // for the epilogue buffers and, in the plain 8-warp variant, keeps the 196 KB carve-out (60 KB of L1)
    static constexpr int stages(int epi_warps, bool rpf) { return 4; }
    // staging area per epilogue warp: 4 KB fp32 tile (+ 2 KB 16-bit tile next to it in the 8-warp variants, so that a
    // chunk's fp32 and 16-bit TMA stores do not wait for each other)
    static constexpr int stage_bytes(int epi_warps, bool cdist) { return (epi_warps == 8 && !cdist) ? 6144 : 4096; }
    static constexpr int smem_bytes(int epi_warps, bool rpf, bool cdist = false) {
        return stages(epi_warps, rpf) * STAGE_BYTES + epi_warps * stage_bytes(epi_warps, cdist) + resid_bytes(epi_warps, rpf) +
               vec_bytes(epi_warps) + 1024 + 256;
    }
};

// TS: epilogue tiles leave through TMA stores (tmOutF / tmOutH / tmAux: fp32 output, 16-bit output, 16-bit aux_out).
template <int NEW, bool CDIST, bool RPF, int EF, bool PREC = false, bool TS = false>  // NEW = epilogue warps (8 or 16); RPF = residual prefetch through smem
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * NEW, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                    const __grid_constant__ CUtensorMap tmOutF, const __grid_constant__ CUtensorMap tmOutH,
                    const __grid_constant__ CUtensorMap tmAux, const __grid_constant__ CUtensorMap tmResid, const GemmArgs args) {
    using Cfg = Pair256;
    constexpr int STAGES = Cfg::stages(NEW, RPF);
    constexpr int BN = Cfg::BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic on the __shared__ array keeps the address space (LDS / STS, not generic LD / ST)
    float* epi_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
    constexpr int SB = Cfg::stage_bytes(NEW, CDIST);  // staging bytes per epilogue warp
    // flag-specialised instantiations are only launched for N % 256 == 0, one batch (launch_pair): vectors always staged
    constexpr bool SPEC = !CDIST && EF >= 0;
    constexpr bool SPEC_SV = SPEC && (EF & (EPI_BIAS | EPI_LN_FOLD | EPI_RESID_LN)) != 0;
    float* resid_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + NEW * SB);
    float* vec_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + NEW * SB + Cfg::resid_bytes(NEW, RPF));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + NEW * SB + Cfg::resid_bytes(NEW, RPF) +
                                                     Cfg::vec_bytes(NEW));
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* resid_bar = tmem_empty + 2;  // [NEW] one per epilogue warp (residual tiles through TMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(resid_bar + NEW);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_tiles = args.m_tiles * args.n_tiles * args.batch;  // m_tiles counts 256-row tiles here
    const int k_blocks = (args.K + BK - 1) / BK;
    const int k_total = PREC ? 3 * k_blocks : k_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 2 * NEW);  // epilogue warps of BOTH CTAs arrive on the leader's
        }
        for (int s = 0; s < NEW; ++s) mbar_init(&resid_bar[s], 1);
        mbar_fence_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tmem_slot, 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch, cluster
    // sync) ran while the previous kernel of the stream was still draining its last tiles; from here on this kernel reads
    // what that kernel wrote.  (No-ops when the kernel was launched without the attribute.)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
                const int n_blk = tile % args.n_tiles;
                const int m_blk = (tile / args.n_tiles) % args.m_tiles;
                const int b = tile / (args.n_tiles * args.m_tiles);
                const int row0 = m_blk * 256 + (int)rank * BM;
                const int col0 = n_blk * BN + (int)rank * (BN / 2);
                for (int kq = 0; kq < k_total; ++kq) {
                    const int seg = PREC ? kq / k_blocks : 0, kb = PREC ? kq - seg * k_blocks : kq;
                    const CUtensorMap* mA = (PREC && seg == 1) ? &tmA2 : &tmA;
                    const CUtensorMap* mB = (PREC && seg == 2) ? &tmB2 : &tmB;
                    mbar_wait_role(&empty_bar[stage], phase ^ 1, args.sleep_ns);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                    if (args.a_wrap > 0) {
                        const int kk = kb * BK;
                        tma_load_3d_pair(sa, mA, &full_bar[stage], kk % args.a_wrap, row0 + kk / args.a_wrap, b);
                    } else {
                        tma_load_3d_pair(sa, mA, &full_bar[stage], kb * BK, row0, b);
                    }
                    tma_load_3d_pair(sb, mB, &full_bar[stage], kb * BK, col0, b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            const uint32_t idesc = umma_idesc_h16(256, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait_role(&tmem_empty[as], aphase ^ 1, args.sleep_ns);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int kb = 0; kb < k_total; ++kb) {
                    mbar_wait_role(&full_bar[stage], phase, args.sleep_ns);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t da = umma_desc_sw128(sa);
                    const uint64_t db = umma_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_f16_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                    umma_commit_pair(&empty_bar[stage]);
                    if (kb == k_total - 1) umma_commit_pair(&tmem_full[as]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int q = warp & 3;
        const int h = ew >> 2;
        constexpr int HALF = BN / (NEW / 4);  // columns per epilogue warp (two or four column groups)
        // residual prefetch: first 32 x 32 fp32 tile of this warp in tile t (nullptr if it is ragged / absent)
        const int eflags = epi_flags<EF>(args.epi);
        const bool want_pf = RPF && (eflags & (EPI_RESID | EPI_RESID_LN)) != 0 && args.batch == 1;
        float* rbuf = want_pf ? resid_stage + ew * 1024 : nullptr;
        auto tile_src = [&](int t) -> const float* {
            if (!want_pf || t >= num_tiles) return nullptr;
            const int n2 = t % args.n_tiles, m2 = (t / args.n_tiles) % args.m_tiles;
            const long long r0 = (long long)m2 * 256 + rank * BM + q * 32;
            const int c0 = n2 * BN + h * HALF;
            return (r0 + 32 <= args.M && c0 + 32 <= args.N) ? args.epi.resid + r0 * args.epi.ldr + c0 : nullptr;
        };
        uint32_t ts_pending = 0;
        int ts_flip = 0;
        StoreMaps smaps;
        smaps.f = (TS && (eflags & (EPI_OUT_F32 | EPI_CDIST)) && args.epi.out_f != nullptr) ? &tmOutF : nullptr;
        smaps.h = (TS && (eflags & EPI_OUT_H16)) ? &tmOutH : nullptr;
        smaps.a = (TS && (eflags & EPI_SAVE_DGELU)) ? &tmAux : nullptr;
        smaps.stage_bytes = SB;
        smaps.pending = &ts_pending;
        smaps.flip = &ts_flip;
        uint32_t rphase = 0;
        smaps.r = (want_pf && args.resid_tma) ? &tmResid : nullptr;
        smaps.rbar = &resid_bar[ew];
        smaps.rphase = &rphase;
        bool rhave = false;
        if (want_pf) {
            const float* src = tile_src(pair_id);
            if (src != nullptr) {
                if (smaps.r != nullptr) resid_tma_issue(smaps.r, smaps.rbar, rbuf, src, args.epi.resid, args.epi.ldr, lane);
                else stage_fill_f32_async(rbuf, src, args.epi.ldr, lane);
                rhave = true;
            }
        }
        int it = 0;
        for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++it) {
            const int n_blk = tile % args.n_tiles;
            const int m_blk = (tile / args.n_tiles) % args.m_tiles;
            const int b = tile / (args.n_tiles * args.m_tiles);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const long long row = (long long)m_blk * 256 + rank * BM + q * 32 + lane;
            const float2 row_st = CDIST ? make_float2(0.f, 1.f) : epilogue_row_stats<EF>(args, row);
            if (!CDIST && (!want_pf || args.resid_l2pf > 0) && (eflags & (EPI_RESID | EPI_RESID_LN)) &&
                tile + (want_pf ? args.resid_l2pf : 1) * num_pairs < num_tiles) {
                // pull a LATER tile's slice of the fp32 residual stream into L2 a whole tile ahead of its use: the one-chunk
                // shared-memory prefetch of a warp keeps only 4 KB in flight, i.e. 32 KB per SM -- at the ~3 k cycles of a
                // loaded HBM round trip that is 2.6 TB/s for the whole GPU (measured: +65 us for the 157 MB residual of the
                // out-projection, profiles/r02_outproj_knockout.log); from L2 the same fills return in a fraction of that
                const int nt = tile + (want_pf ? args.resid_l2pf : 1) * num_pairs;
                const int n2 = nt % args.n_tiles, m2 = (nt / args.n_tiles) % args.m_tiles;
                const long long r2 = (long long)m2 * 256 + rank * BM + q * 32 + lane;
                const int c2 = n2 * BN + h * HALF;
                if (r2 < args.M && c2 + HALF <= args.N)
                    prefetch_l2_bulk(args.epi.resid + r2 * args.epi.ldr + c2, HALF * 4);
            }
            const float* next_src = tile_src(tile + num_pairs);
            // column vectors of this warp's HALF columns -> shared memory (only for complete column ranges, batch 1)
            float* vec = nullptr;
            if (!CDIST && args.batch == 1 && n_blk * BN + h * HALF + HALF <= args.N &&
                (eflags & (EPI_BIAS | EPI_LN_FOLD | EPI_RESID_LN)) != 0) {
                vec = vec_stage + ew * (3 * HALF);
                epilogue_stage_vectors<EF, HALF>(args.epi, vec, n_blk * BN + h * HALF, lane);
            }
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            epilogue_tile<HALF / 32, true, CDIST, EF, PREC, 0, TS, SPEC_SV, SPEC && RPF && (EF & (EPI_RESID | EPI_RESID_LN)) != 0>(args, tmem_base + (uint32_t)(as * 256 + h * HALF) + ((uint32_t)(q * 32) << 16), row,
                                           n_blk * BN + h * HALF, h * HALF, b, &tmem_empty[as], lane, epi_stage + ew * (SB / 4),
                                           row_st, rbuf, rhave, next_src, vec, HALF, n_blk * (NEW / 4) + h, &smaps);
        }
        if (TS && lane == 0) bulk_wait_all();  // this warp's bulk stores have left shared memory and are globally visible
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// SIMT check kernel: same operands, same epilogue, one thread per output element.  Used by the GPU
// parity tests to cross-check the tensor-core kernel on device (and selectable with
// NOMAD_B200_GEMM=simt for debugging); never the default.
__global__ void gemm_simt_kernel(GemmOperand A, GemmOperand B, GemmArgs args) {
    __shared__ float sa[16][17];
    __shared__ float sb[16][17];
    const int b = blockIdx.z;
    const long long row = (long long)blockIdx.y * 16 + threadIdx.y;
    const int col = blockIdx.x * 16 + threadIdx.x;
    const op_t* a = A.ptr + (long long)b * A.batch_stride;
    const op_t* w = B.ptr + (long long)b * B.batch_stride;
    float acc = 0.f;
    for (int k0 = 0; k0 < args.K; k0 += 16) {
        const long long ar = (long long)blockIdx.y * 16 + threadIdx.y;
        const int ak = k0 + threadIdx.x;
        const long long aoff = A.k_wrap > 0 ? (ar + ak / A.k_wrap) * A.row_stride + ak % A.k_wrap : ar * A.row_stride + ak;
        sa[threadIdx.y][threadIdx.x] = (ar < args.M && ak < args.K) ? op2f(a[aoff]) : 0.f;
        const int br = blockIdx.x * 16 + threadIdx.y;
        sb[threadIdx.y][threadIdx.x] = (br < args.N && ak < args.K) ? op2f(w[(long long)br * B.row_stride + ak]) : 0.f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = fmaf(sa[threadIdx.y][k], sb[threadIdx.x][k], acc);
        __syncthreads();
    }
    if (row < args.M && col < args.N) epilogue_scalar(args.epi, acc, row, col, b);
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// 3-D map over (K, rows, batch) of a op_t operand, box = (64, box_rows, 1), 128-byte swizzle.
static int make_operand_map(CUtensorMap* map, const GemmOperand& op, int K, int batch, int box_rows) {
    if (op.k_wrap > 0) {
        NB_CHECK(op.k_wrap % BK == 0 && op.row_stride == op.k_wrap, "wrapped-K operand needs k_wrap %% 64 == 0 and row_stride == k_wrap");
        K = op.k_wrap;
    }
    EncodeTiledFn fn = get_encode_fn();
    NB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    NB_CHECK((reinterpret_cast<uintptr_t>(op.ptr) & 15) == 0, "GEMM operand pointer must be 16-byte aligned");
    NB_CHECK((op.row_stride * 2) % 16 == 0 && (op.batch_stride * 2) % 16 == 0,
             "GEMM operand strides must be multiples of 16 bytes (row_stride=%lld batch_stride=%lld)", op.row_stride,
             op.batch_stride);
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)op.rows, (cuuint64_t)batch};
    cuuint64_t gstride[2] = {(cuuint64_t)op.row_stride * 2,
                             (cuuint64_t)(batch > 1 ? op.batch_stride : op.row_stride * op.rows) * 2};
    if (gstride[1] == 0) gstride[1] = 16;
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<op_t*>(op.ptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NB_CHECK(r == CUDA_SUCCESS,
             "cuTensorMapEncodeTiled failed (%d): K=%d rows=%lld batch=%d row_stride=%lld batch_stride=%lld box_rows=%d",
             (int)r, K, op.rows, batch, op.row_stride, op.batch_stride, box_rows);
    return 0;
}

int device_sm_count() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        sms[dev] = n;
    }
    return sms[dev];
}

static bool use_pair_kernel(int M, int N, int batch) {
    static const int use_pair = getenv("NOMAD_B200_PAIR") ? atoi(getenv("NOMAD_B200_PAIR")) : 1;
    return use_pair && N >= 256 && (long long)((M + 255) / 256) * ((N + 255) / 256) * batch >= device_sm_count() / 2;
}
// must mirror the kernel choice of gemm_h16 for EPI_CDIST: pair kernel = 16 epilogue warps (4 column groups per
// 256-wide tile); single-CTA kernels = 2 column groups per BN-wide tile (BN = 256 / 128 / 64; EPI_CDIST never takes 192)
int cdist_row_groups(long long n, long long m) {
    if (use_pair_kernel((int)n, (int)m, 1)) return (int)((m + 255) / 256) * 4;
    const int bn = m > 128 ? 256 : (m > 64 ? 128 : 64);
    return (int)((m + bn - 1) / bn) * 2;
}

// In-situ timing of the tensor-core GEMM launches (bench.py's roofline leg): one CUDA event pair per
// launch on the launching stream, read back after the timed region.
struct GemmProfile {
    bool on = false;
    std::vector<cudaEvent_t> ev;  // pairs
    std::vector<double> flops;
    size_t used = 0;
};
static GemmProfile g_prof;

void gemm_profile_enable(bool on) {
    g_prof.on = on;
    g_prof.used = 0;
    g_prof.flops.clear();
}
bool gemm_profile_active() { return g_prof.on; }
int gemm_profile_read(double* total_ms, double* total_flops, long long* launches) {
    double ms = 0.0, fl = 0.0;
    for (size_t i = 0; i < g_prof.used; ++i) {
        NB_CUDA(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
        float t = 0.f;
        NB_CUDA(cudaEventElapsedTime(&t, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
        ms += t;
        fl += g_prof.flops[i];
    }
    *total_ms = ms;
    *total_flops = fl;
    *launches = (long long)g_prof.used;
    g_prof.used = 0;
    g_prof.flops.clear();
    return 0;
}
static int prof_begin(cudaStream_t st, double flops) {
    if (!g_prof.on) return 0;
    if (g_prof.ev.size() < 2 * (g_prof.used + 1)) {
        cudaEvent_t a, b;
        NB_CUDA(cudaEventCreate(&a));
        NB_CUDA(cudaEventCreate(&b));
        g_prof.ev.push_back(a);
        g_prof.ev.push_back(b);
    }
    g_prof.flops.push_back(flops);
    NB_CUDA(cudaEventRecord(g_prof.ev[2 * g_prof.used], st));
    return 0;
}
static int prof_end(cudaStream_t st) {
    if (!g_prof.on) return 0;
    NB_CUDA(cudaEventRecord(g_prof.ev[2 * g_prof.used + 1], st));
    g_prof.used++;
    return 0;
}

static int lo_operand(const GemmOperand& op, GemmOperand* out) {
    *out = op;
    if (op.lo != nullptr) out->ptr = op.lo;
    return 0;
}

template <int BN, bool CDIST, bool PREC = false>
static int launch_tc_impl(cudaStream_t st, const GemmOperand& A, const GemmOperand& B, GemmArgs& args) {
    using Cfg = TileCfg<BN>;
    static bool attr_set[64] = {false};  // the attribute is per device
    if (bool* flag = device_once_flag(attr_set)) {
        NB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, CDIST, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        *flag = true;
    }
    CUtensorMap tmA, tmB, tmA2, tmB2;
    // shared-operand batching: one tensor per operand (no batch dimension); B's map spans its whole row so that the
    // per-batch K shift reads zeros outside it
    const int mb = args.shared_ab ? 1 : args.batch;
    const int kb_ext = args.shared_ab ? (int)B.row_stride : args.K;
    NB_TRY(make_operand_map(&tmA, A, args.K, mb, BM));
    NB_TRY(make_operand_map(&tmB, B, kb_ext, args.shared_ab ? args.b_mod : mb, args.umma_n));
    GemmOperand Al, Bl;
    lo_operand(A, &Al);
    lo_operand(B, &Bl);
    NB_TRY(make_operand_map(&tmA2, Al, args.K, mb, BM));
    NB_TRY(make_operand_map(&tmB2, Bl, kb_ext, args.shared_ab ? args.b_mod : mb, args.umma_n));
    args.m_tiles = (args.M + BM - 1) / BM;
    args.n_tiles = (args.N + BN - 1) / BN;
    const long long tiles = (long long)args.m_tiles * args.n_tiles * args.batch;
    int grid = device_sm_count();
    if (tiles < grid) grid = (int)tiles;
    NB_TRY(prof_begin(st, 2.0 * args.M * args.N * args.K * args.batch));
    gemm_tc_kernel<BN, CDIST, PREC><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, tmA2, tmB2, args);
    NB_LAUNCHED();
    NB_TRY(prof_end(st));
    return 0;
}
template <int BN>
static int launch_tc(cudaStream_t st, const GemmOperand& A, const GemmOperand& B, GemmArgs& args) {
    if constexpr (BN <= 128) {
        if (args.epi.flags & EPI_PRECISE) return launch_tc_impl<BN, false, true>(st, A, B, args);
    }
    return (args.epi.flags & EPI_CDIST) ? launch_tc_impl<BN, true>(st, A, B, args) : launch_tc_impl<BN, false>(st, A, B, args);
}

// 2-D map over an (rows x cols, leading dimension ld) output for TMA stores of 32 x 32 boxes in the staging swizzle
static int make_store_map(CUtensorMap* map, const void* ptr, bool f32, long long rows, long long cols, long long ld) {
    EncodeTiledFn fn = get_encode_fn();
    NB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    const size_t el = f32 ? 4 : 2;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * el};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim,
                    gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (store map) failed (%d): rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
    return 0;
}
// can every output of this epilogue leave through TMA?  (16-byte aligned base and pitch, one batch)
// Measured on the bench step (profiles/r02_tma_store_ab.txt, alternating runs in one box): TMA stores gain 2-6 % where
// the epilogue only emits one 16-bit tile per chunk and has no LayerNorm algebra (conv1-6, QKV) and on the distance
// matrix; they LOSE 3-5 % on the residual epilogues (out-proj, FC2: the fence + wait per chunk serialises warps that
// are already latency-bound on the residual stream) and on FC1.  NOMAD_B200_TMA_STORE: 0 = never, 1 = where it wins
// (default), 2 = every epilogue that can.
static bool tma_store_ok(const GemmArgs& args) {
    static const int enabled = getenv("NOMAD_B200_TMA_STORE") ? atoi(getenv("NOMAD_B200_TMA_STORE")) : 1;
    const GemmEpilogue& e = args.epi;
    if (!enabled || args.batch != 1 || (e.flags & EPI_PRECISE)) return false;
    if (enabled == 1 && ((e.flags & (EPI_RESID | EPI_RESID_LN | EPI_SAVE_DGELU | EPI_MUL_AUX)) ||
                         ((e.flags & EPI_LN_FOLD) && (e.flags & EPI_GELU))))
        return false;
    auto ok = [&](const void* p, size_t el) { return p != nullptr && ((uintptr_t)p & 15) == 0 && (e.ldo * el) % 16 == 0; };
    if ((e.flags & EPI_CDIST) && e.out_f != nullptr && !ok(e.out_f, 4)) return false;
    if ((e.flags & EPI_OUT_F32) && !ok(e.out_f, 4)) return false;
    if ((e.flags & EPI_OUT_H16) && !ok(e.out_h, 2)) return false;
    if ((e.flags & EPI_SAVE_DGELU) && !ok(e.aux_out, 2)) return false;
    return true;
}

template <int NEW, bool CDIST, bool RPF = false, int EF = -1, bool PREC = false, bool TS = true>
static int launch_pair_impl(cudaStream_t st, const GemmOperand& A, const GemmOperand& B, GemmArgs& args) {
    using Cfg = Pair256;
    if (TS && !tma_store_ok(args)) return launch_pair_impl<NEW, CDIST, RPF, EF, PREC, false>(st, A, B, args);
    constexpr int SMEM = Cfg::smem_bytes(NEW, RPF, CDIST);
    static bool attr_set[64] = {false};  // the attribute is per device
    if (bool* flag = device_once_flag(attr_set)) {
        NB_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<NEW, CDIST, RPF, EF, PREC, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        *flag = true;
    }
    CUtensorMap tmA, tmB, tmA2, tmB2;
    NB_TRY(make_operand_map(&tmA, A, args.K, args.batch, BM));
    NB_TRY(make_operand_map(&tmB, B, args.K, args.batch, Cfg::BN / 2));
    GemmOperand Al, Bl;
    lo_operand(A, &Al);
    lo_operand(B, &Bl);
    NB_TRY(make_operand_map(&tmA2, Al, args.K, args.batch, BM));
    NB_TRY(make_operand_map(&tmB2, Bl, args.K, args.batch, Cfg::BN / 2));
    args.umma_n = Cfg::BN;
    args.m_tiles = (args.M + 255) / 256;
    args.n_tiles = (args.N + Cfg::BN - 1) / Cfg::BN;
    const long long tiles = (long long)args.m_tiles * args.n_tiles * args.batch;
    int pairs = device_sm_count() / 2;
    if (tiles < pairs) pairs = (int)tiles;
    NB_TRY(prof_begin(st, 2.0 * args.M * args.N * args.K * args.batch));
    CUtensorMap tmF = tmA, tmH = tmA, tmX = tmA, tmR = tmA;  // placeholders when an output is absent
    if (RPF && args.resid_tma) NB_TRY(make_store_map(&tmR, args.epi.resid, true, args.M, args.N, args.epi.ldr));
    if (TS) {
        const GemmEpilogue& e = args.epi;
        if ((e.flags & (EPI_OUT_F32 | EPI_CDIST)) && e.out_f) NB_TRY(make_store_map(&tmF, e.out_f, true, args.M, args.N, e.ldo));
        if (e.flags & EPI_OUT_H16) NB_TRY(make_store_map(&tmH, e.out_h, false, args.M, args.N, e.ldo));
        if (e.flags & EPI_SAVE_DGELU) NB_TRY(make_store_map(&tmX, e.aux_out, false, args.M, args.N, e.ldo));
    }
    static const int pdl = getenv("NOMAD_B200_PDL") ? atoi(getenv("NOMAD_B200_PDL")) : 1;
    if (pdl) {  // programmatic dependent launch: this kernel's prologue overlaps the tail of the previous kernel
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(128 + 32 * NEW);
        cfg.dynamicSmemBytes = SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        NB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel<NEW, CDIST, RPF, EF, PREC, TS>, tmA, tmB, tmA2, tmB2, tmF, tmH, tmX, tmR, args));
    } else {
        gemm_tc_pair_kernel<NEW, CDIST, RPF, EF, PREC, TS><<<2 * pairs, 128 + 32 * NEW, SMEM, st>>>(tmA, tmB, tmA2, tmB2, tmF, tmH, tmX, tmR, args);
    }
    NB_LAUNCHED();
    NB_TRY(prof_end(st));
    return 0;
}
static int launch_pair(cudaStream_t st, const GemmOperand& A, const GemmOperand& B, GemmArgs& args) {
    if (args.epi.flags & EPI_CDIST) return launch_pair_impl<16, true>(st, A, B, args);  // epilogue-bound (sqrt, sums, 4 B/pair out)
    // 16 epilogue warps hide the latency of a math-heavy (GELU) epilogue when the mainloop is short (FC1, K = 768:
    // 887 vs 841 TFLOP/s); with a long mainloop (conv, K = 1536) the extra warps only cost registers (1075 vs 1107).
    // NOMAD_B200_EPI16 overrides (0 = never, 2 = always).
    static const int epi16 = getenv("NOMAD_B200_EPI16") ? atoi(getenv("NOMAD_B200_EPI16")) : 1;
    // A short mainloop whose epilogue streams an fp32 residual in and an fp32 tile out (out-proj, K = 768) also gains
    // from twice the warps keeping loads in flight (0.125 -> 0.106 ms) -- unless it additionally emits the 16-bit
    // copy and the LayerNorm statistics: that epilogue does not fit the 96-register budget of the 640-thread CTA
    // (in-model 231 us with 16 warps, 170 us with 8).
    const int fl = args.epi.flags;
    const bool heavy = args.K <= 1024 && ((fl & (EPI_GELU | EPI_SAVE_DGELU)) != 0 ||
                                          ((fl & (EPI_RESID | EPI_RESID_LN)) != 0 && !(fl & EPI_STATS_OUT)));
    // hot flag combinations of the scoring path run flag-specialised instantiations (NOMAD_B200_EPI_SPEC=0: generic)
    static const int spec_env = getenv("NOMAD_B200_EPI_SPEC") ? atoi(getenv("NOMAD_B200_EPI_SPEC")) : 1;
    const bool spec = spec_env && args.N % Pair256::BN == 0 && args.batch == 1;  // what SPEC_SV in the kernel relies on
    constexpr int F_FC1 = EPI_LN_FOLD | EPI_GELU | EPI_OUT_H16;
    constexpr int F_QKV = EPI_LN_FOLD | EPI_OUT_H16;
    constexpr int F_CONV = EPI_GELU | EPI_OUT_H16;
    constexpr int F_RES = EPI_BIAS | EPI_RESID_LN | EPI_OUT_F32 | EPI_OUT_H16 | EPI_STATS_OUT;
    // loss path (activations saved for the backward) and the layers at the ends of the stack
    constexpr int F_FC1_SAVE = EPI_BIAS | EPI_GELU | EPI_OUT_H16 | EPI_SAVE_DGELU;
    constexpr int F_CONV_SAVE = EPI_GELU | EPI_OUT_H16 | EPI_SAVE_DGELU;
    constexpr int F_LIN_H16 = EPI_BIAS | EPI_OUT_H16;
    constexpr int F_RES_PLAIN = EPI_BIAS | EPI_RESID_LN | EPI_OUT_F32;
    constexpr int F_BRES = EPI_BIAS | EPI_RESID | EPI_OUT_F32;   // loss forward: out-proj / FC2 (explicit LayerNorm kernels follow)
    constexpr int F_DG_RES = EPI_RESID | EPI_OUT_F32;            // dgrad + residual branch
    constexpr int F_DG_AUX = EPI_MUL_AUX | EPI_OUT_H16;          // dgrad through GELU (x gelu')
    constexpr int F_H16 = EPI_OUT_H16;                           // plain dgrad
    if (epi16 == 2 || (epi16 == 1 && heavy)) {
        if (spec && fl == F_BRES) return launch_pair_impl<16, false, false, F_BRES>(st, A, B, args);
        if (spec && fl == F_RES_PLAIN) return launch_pair_impl<16, false, false, F_RES_PLAIN>(st, A, B, args);
        if (spec && fl == F_FC1) return launch_pair_impl<16, false, false, F_FC1>(st, A, B, args);
        if (spec && fl == F_CONV) return launch_pair_impl<16, false, false, F_CONV>(st, A, B, args);
        if (spec && fl == F_FC1_SAVE) return launch_pair_impl<16, false, false, F_FC1_SAVE>(st, A, B, args);
        if (spec && fl == F_CONV_SAVE) return launch_pair_impl<16, false, false, F_CONV_SAVE>(st, A, B, args);
        return launch_pair_impl<16, false>(st, A, B, args);
    }
    // Measured and dropped (round 2, profiles/r02_experiments.md): this epilogue on 16 warps without the residual prefetch
    // buffer (96-register budget: 100 B of spills) runs out-proj at 192-212 us instead of 139 us.
    static const int rpf = getenv("NOMAD_B200_RESID_PREFETCH") ? atoi(getenv("NOMAD_B200_RESID_PREFETCH")) : 1;
    if (rpf && (fl & (EPI_RESID | EPI_RESID_LN)) != 0 && args.batch == 1) {
        if (spec && fl == F_RES) return launch_pair_impl<8, false, true, F_RES>(st, A, B, args);
        if (spec && fl == F_RES_PLAIN) return launch_pair_impl<8, false, true, F_RES_PLAIN>(st, A, B, args);
        if (spec && fl == F_BRES) return launch_pair_impl<8, false, true, F_BRES>(st, A, B, args);
        if (spec && fl == F_DG_RES) return launch_pair_impl<8, false, true, F_DG_RES>(st, A, B, args);
        return launch_pair_impl<8, false, true>(st, A, B, args);
    }
    if (spec && fl == F_QKV) return launch_pair_impl<8, false, false, F_QKV>(st, A, B, args);
    if (spec && fl == F_CONV) return launch_pair_impl<8, false, false, F_CONV>(st, A, B, args);
    if (spec && fl == F_CONV_SAVE) return launch_pair_impl<8, false, false, F_CONV_SAVE>(st, A, B, args);
    if (spec && fl == F_LIN_H16) return launch_pair_impl<8, false, false, F_LIN_H16>(st, A, B, args);
    if (spec && fl == F_DG_AUX) return launch_pair_impl<8, false, false, F_DG_AUX>(st, A, B, args);
    if (spec && fl == F_H16) return launch_pair_impl<8, false, false, F_H16>(st, A, B, args);
    return launch_pair_impl<8, false>(st, A, B, args);
}

static thread_local int g_corr_shift0 = 0, g_corr_step = 0, g_corr_on = 0, g_corr_mod = 1;

// C[b] = epilogue(sum_k A[m, k] B_{b % b_mod}[n, k + shift0 + (b / b_mod) * step]) for b in [0, batch): ONE A tensor shared by
// all batches, b_mod B tensors (B.batch_stride apart), the B operand's K origin shifted per batch (columns outside the row read
// as zero).  TMA needs the shifted origin 16-byte aligned: shift0 and step must be multiples of 8 elements -- a caller that
// needs every shift keeps 8 copies of B pre-shifted by 0..7 elements (b_mod = 8).  N <= 128 (single-CTA kernel).
int gemm_h16_corr(cudaStream_t st, const GemmOperand& A, const GemmOperand& B, int M, int N, int K, int batch, int shift0,
                  int step, int b_mod, const GemmEpilogue& epi) {
    NB_CHECK(N <= 128, "gemm_h16_corr: N <= 128");
    NB_CHECK(shift0 % 8 == 0 && step % 8 == 0 && b_mod >= 1, "gemm_h16_corr: shifts must be multiples of 8 elements");
    g_corr_on = 1; g_corr_shift0 = shift0; g_corr_step = step; g_corr_mod = b_mod;
    const int rc = gemm_h16(st, A, B, M, N, K, batch, epi, 0);
    g_corr_on = 0;
    return rc;
}

int gemm_h16(cudaStream_t st, const GemmOperand& A, const GemmOperand& B, int M, int N, int K, int batch,
              const GemmEpilogue& epi, int impl) {
    if (M <= 0 || N <= 0 || batch <= 0) return 0;
    NB_CHECK(K > 0 && K % 8 == 0, "GEMM K=%d must be a positive multiple of 8", K);
    GemmArgs args;
    args.M = M; args.N = N; args.K = K; args.batch = batch;
    args.epi = epi;
    args.umma_n = 0; args.m_tiles = args.n_tiles = 0;
    args.a_wrap = A.k_wrap;
    args.shared_ab = g_corr_on; args.b_kshift0 = g_corr_shift0; args.b_kshift_step = g_corr_step; args.b_mod = g_corr_mod;
    static const unsigned sleep_ns = getenv("NOMAD_B200_GEMM_SLEEP") ? (unsigned)atoi(getenv("NOMAD_B200_GEMM_SLEEP")) : 0u;
    args.sleep_ns = sleep_ns;
    static const int resid_l2pf = getenv("NOMAD_B200_RESID_L2PF") ? atoi(getenv("NOMAD_B200_RESID_L2PF")) : 0;
    args.resid_l2pf = resid_l2pf;
    static const int resid_tma = getenv("NOMAD_B200_RESID_TMA") ? atoi(getenv("NOMAD_B200_RESID_TMA")) : 0;
    args.resid_tma = (resid_tma && (epi.flags & (EPI_RESID | EPI_RESID_LN)) && epi.resid != nullptr && batch == 1 &&
                      ((uintptr_t)epi.resid & 15) == 0 && (epi.ldr * 4) % 16 == 0) ? 1 : 0;
    NB_CHECK(B.k_wrap == 0, "only the A operand may use wrapped K");
    if (epi.flags & EPI_PRECISE) {
        NB_CHECK(impl == 0 && A.lo != nullptr && B.lo != nullptr && epi.acc_scale > 0.f, "EPI_PRECISE needs hi + lo operand planes and a scale");
        NB_CHECK(!(epi.flags & EPI_OUT_H16) || epi.out_l != nullptr, "EPI_PRECISE with a 16-bit output needs the lo plane pointer");
        NB_CHECK(!(epi.flags & (EPI_CDIST | EPI_LN_FOLD | EPI_STATS_OUT | EPI_RESID_LN | EPI_SAVE_DGELU | EPI_MUL_AUX)),
                 "EPI_PRECISE supports bias / GELU / residual / fp32 and split 16-bit outputs only");
    }
    if (!(epi.flags & EPI_CDIST)) {
        NB_CHECK(N % 8 == 0 && epi.ldo % 8 == 0 && epi.out_bstride % 8 == 0,
                 "GEMM epilogue needs N, ldo, out_bstride multiples of 8 (N=%d ldo=%lld bstride=%lld)", N, epi.ldo,
                 epi.out_bstride);
        NB_CHECK(!(epi.flags & EPI_RESID) || (epi.ldr % 4 == 0 && epi.resid_bstride % 4 == 0),
                 "GEMM residual leading dimension must be a multiple of 4");
    }
    NB_CHECK(!(epi.flags & (EPI_LN_FOLD | EPI_STATS_OUT)) || impl == 0, "LayerNorm-folding epilogues exist only in the tensor-core kernel");
    NB_CHECK(!(epi.flags & EPI_STATS_OUT) || (N == 64 * LN_PARTS && batch == 1 && epi.part_out != nullptr),
             "EPI_STATS_OUT needs N = %d", 64 * LN_PARTS);
    NB_CHECK(!(epi.flags & EPI_LN_FOLD) || (epi.ln_part && epi.fold_s && epi.fold_c && N % 4 == 0 && batch == 1),
             "EPI_LN_FOLD needs partial statistics, s and c vectors");
    if (impl == 1) {
        dim3 grid((N + 15) / 16, (M + 15) / 16, batch), block(16, 16);
        gemm_simt_kernel<<<grid, block, 0, st>>>(A, B, args);
        NB_LAUNCHED();
        return 0;
    }
    // CTA pairs (cta_group::2) for the big GEMMs; NOMAD_B200_PAIR=0 falls back to the single-CTA kernel
    if (epi.flags & EPI_PRECISE) {  // four TMEM accumulators per tile: 128-wide (or 64-wide) single-CTA tiles
        NB_CHECK(K >= 3 * BK, "EPI_PRECISE needs K >= %d", 3 * BK);
        if (N > 64) {
            args.umma_n = 128;
            return launch_tc<128>(st, A, B, args);
        }
        args.umma_n = (N + 15) / 16 * 16;
        return launch_tc<64>(st, A, B, args);
    }
    if (use_pair_kernel(M, N, batch)) return launch_pair(st, A, B, args);
    if (N > 128) {
        // 256-wide tiles unless 192-wide ones waste enough fewer CTA-waves to pay for their lower per-tile
        // efficiency (measured 0.89 on B200, profiles/r01_gemm_probe_v2.log)
        static const int force_bn = getenv("NOMAD_B200_BN") ? atoi(getenv("NOMAD_B200_BN")) : 0;
        const double sms = device_sm_count();
        auto eff = [&](int bn, double tile_eff) {
            const double tiles = (double)((M + BM - 1) / BM) * ((N + bn - 1) / bn) * batch;
            const double used = (double)N / (((N + bn - 1) / bn) * (double)bn);
            return tile_eff * used * tiles / (std::ceil(tiles / sms) * sms);
        };
        const bool use192 = (epi.flags & (EPI_STATS_OUT | EPI_CDIST | EPI_PRECISE)) ? false : (force_bn ? force_bn == 192 : eff(192, 0.89) > eff(256, 1.0));
        if (use192) {
            args.umma_n = 192;
            return launch_tc<192>(st, A, B, args);
        }
        args.umma_n = 256;
        return launch_tc<256>(st, A, B, args);
    } else if (N > 64) {
        args.umma_n = 128;
        return launch_tc<128>(st, A, B, args);
    } else {
        args.umma_n = (N + 15) / 16 * 16;
        return launch_tc<64>(st, A, B, args);
    }
}

}  // namespace nb
