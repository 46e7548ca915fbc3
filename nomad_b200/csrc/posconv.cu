// Grouped positional convolution (768 ch, 16 groups of 48, k = 128, pad 64) as a dedicated tcgen05 kernel.
//
// Input: pos_g[g][p][48] (16-bit, every utterance surrounded by >= 64 zero rows, see pos_scatter_kernel).
// Output row m of group g:  y[m, g*48 + n] = epi( sum_{k<128} sum_{c<48} pos_g[g][m + k][c] * w[g][n][k*48 + c] ).
//
// Viewed as a GEMM the A operand has K = 6144 but consecutive output rows share 127/128 of it, so instead of
// streaming 128 x 6144 values per 128 output rows (what the generic overlapping-row GEMM does: ~10 GB through
// L2 per forward, L2-bound) this kernel loads a SLAB of 512 + 127 input rows x 48 channels ONCE per tile into
// shared memory in the un-swizzled core-matrix layout [8-channel chunk][row][16 B].  In that layout the row
// stride is a uniform 16 B, so the A descriptor of tap k is simply the slab base advanced by k rows: every
// (tap, 16-channel) step is one tcgen05.mma (M = 128, N = 48, K = 16) per 128-row sub-tile, 4 sub-tiles share
// each streamed weight block.  The weights (B operand, [48][6144] per group, K-major, 128-byte swizzle) stream
// through an 8-stage TMA ring exactly like the generic GEMM.  Accumulators: 4 sub-tiles x 48 columns in TMEM,
// double buffered so the epilogue (bias + erf-GELU, optional gelu' save) overlaps the next tile.
#include <mutex>

#include "kernels.cuh"

namespace nb {

static constexpr int PC_SUB = 4;                       // 128-row sub-tiles per CTA tile
static constexpr int PC_ROWS = 128 * PC_SUB;           // 512 output rows per tile
static constexpr int PC_SLAB_ROWS = PC_ROWS + 128;     // 640 (>= 512 + 127)
static constexpr int PC_CHUNKS = POS_GC / 8;           // 6 chunks of 8 channels (16 B)
static constexpr int PC_LBO = PC_SLAB_ROWS * 16;       // bytes between K-chunks
static constexpr int PC_SLAB_BYTES = PC_CHUNKS * PC_LBO;  // 61440
static constexpr int PC_BSTAGE_BYTES = POS_GC * 128;   // 48 rows x 128 B = 6144
static constexpr int PC_BSTAGES = 8;
static constexpr int PC_KBLOCKS = POS_K * POS_GC / 64; // 96
static constexpr int PC_THREADS = 256;
static constexpr int PC_SMEM = 2 * PC_SLAB_BYTES + PC_BSTAGES * PC_BSTAGE_BYTES + 1024 + 256;
static constexpr int PC_ACC_STRIDE = 64;               // TMEM columns per sub-tile accumulator

struct PosConvArgs {
    long long pos_rows;     // output rows per group
    int m_tiles;
    int flags;              // EPI_BIAS | EPI_GELU | EPI_SAVE_DGELU
    const float* bias;      // [768]
    op_t* out;              // [pos_rows][768]
    op_t* aux_out;          // [pos_rows][768] or nullptr
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// un-swizzled K-major operand: rows 16 B apart (SBO = 128 B per 8 rows), K-chunks LBO bytes apart
__device__ __forceinline__ uint64_t umma_desc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;  // layout type 0 = SWIZZLE_NONE
}

// tcgen05.mma with both descriptors passed as 32-bit halves (no 64-bit arithmetic at the call site)
__device__ __forceinline__ void umma_f16_parts(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(PC_THREADS, 1)
posconv_kernel(const __grid_constant__ CUtensorMap tmSlab, const __grid_constant__ CUtensorMap tmW, const PosConvArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic on the __shared__ array keeps the address space (LDS / STS, not generic LD / ST)
    uint8_t* slab[2] = {smem, smem + PC_SLAB_BYTES};
    uint8_t* bst = smem + 2 * PC_SLAB_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(bst + PC_BSTAGES * PC_BSTAGE_BYTES);
    uint64_t* b_full = bar;                       // [8]
    uint64_t* b_empty = bar + PC_BSTAGES;         // [8]
    uint64_t* slab_full = bar + 2 * PC_BSTAGES;   // [2]
    uint64_t* slab_empty = slab_full + 2;         // [2]
    uint64_t* tmem_full = slab_empty + 2;         // [2]
    uint64_t* tmem_empty = tmem_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = args.m_tiles * POS_G;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmSlab);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < PC_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&slab_full[s], 1);
            mbar_init(&slab_empty[s], 1);
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 4);
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int g = tile % POS_G, mt = tile / POS_G;
                const int sb = it & 1;
                const uint32_t sphase = (it >> 1) & 1;
                mbar_wait(&slab_empty[sb], sphase ^ 1);
                mbar_expect_tx(&slab_full[sb], PC_SLAB_BYTES);
#pragma unroll 1
                for (int c = 0; c < PC_CHUNKS; ++c)
#pragma unroll 1
                    for (int r = 0; r < PC_SLAB_ROWS / 128; ++r)
                        tma_load_3d(slab[sb] + c * PC_LBO + r * 128 * 16, &tmSlab, &slab_full[sb], c * 8,
                                    mt * PC_ROWS + r * 128, g);
                for (int kb = 0; kb < PC_KBLOCKS; ++kb) {
                    mbar_wait(&b_empty[stage], phase ^ 1);
                    mbar_expect_tx(&b_full[stage], PC_BSTAGE_BYTES);
                    tma_load_3d(bst + stage * PC_BSTAGE_BYTES, &tmW, &b_full[stage], kb * 64, 0, g);
                    if (++stage == PC_BSTAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_h16(128, POS_GC);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int sb = it & 1;
                const uint32_t sphase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[sb], sphase ^ 1);
                mbar_wait(&slab_full[sb], sphase);
                tc_fence_after();
                // Descriptors as (lo, hi) 32-bit halves: the start-address field sits in the low word and never
                // carries, so advancing to another tap / sub-tile / K-chunk is ONE 32-bit add per MMA.  The single
                // issuing thread has to sustain one N=48 MMA every ~30 cycles, so its instruction count matters.
                const uint64_t da0 = umma_desc_noswz(smem_u32(slab[sb]), PC_LBO, 128);
                const uint32_t da_lo = (uint32_t)da0, da_hi = (uint32_t)(da0 >> 32);
                const uint32_t d_base = tmem_base + sb * (PC_SUB * PC_ACC_STRIDE);
                uint32_t acc = 0;
#pragma unroll 1
                for (int kb3 = 0; kb3 < PC_KBLOCKS / 3; ++kb3) {
                    // 3 weight blocks = 12 K-steps of 16 = 4 taps x 3 channel thirds; everything below is unrolled
                    const uint32_t a_tap0 = da_lo + 4 * kb3;  // tap * 16 B >> 4
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        mbar_wait(&b_full[stage], phase);
                        tc_fence_after();
                        const uint64_t db0 = umma_desc_sw128(smem_u32(bst + stage * PC_BSTAGE_BYTES));
                        const uint32_t db_lo = (uint32_t)db0, db_hi = (uint32_t)(db0 >> 32);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const int t = 4 * j + k4;              // 0..11 (compile time)
                            const int tap = t / 3, cc = t % 3;
                            const uint32_t a_lo = a_tap0 + tap + (2 * cc) * (PC_LBO >> 4);
#pragma unroll
                            for (int sub = 0; sub < PC_SUB; ++sub)
                                umma_f16_parts(d_base + sub * PC_ACC_STRIDE, a_lo + sub * 128, da_hi, db_lo + 2 * k4, db_hi,
                                               idesc, acc);
                            acc = 1;
                        }
                        umma_commit(&b_empty[stage]);
                        if (++stage == PC_BSTAGES) { stage = 0; phase ^= 1; }
                    }
                }
                umma_commit(&slab_empty[sb]);
                umma_commit(&tmem_full[sb]);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int g = tile % POS_G, mt = tile / POS_G;
            const int sb = it & 1;
            const uint32_t sphase = (it >> 1) & 1;
            mbar_wait(&tmem_full[sb], sphase);
            tc_fence_after();
            float bias[POS_GC];
#pragma unroll
            for (int j = 0; j < POS_GC; ++j) bias[j] = (args.flags & EPI_BIAS) ? __ldg(args.bias + g * POS_GC + j) : 0.f;
#pragma unroll 1
            for (int sub = 0; sub < PC_SUB; ++sub) {
                const long long m = (long long)mt * PC_ROWS + sub * 128 + q * 32 + lane;
                const uint32_t taddr = tmem_base + (uint32_t)(sb * (PC_SUB * PC_ACC_STRIDE) + sub * PC_ACC_STRIDE) +
                                       ((uint32_t)(q * 32) << 16);
                uint32_t r0[32], r1[16];
                tmem_ld_32x32(taddr, r0);
                tmem_ld_32x16(taddr + 32, r1);
                tmem_ld_wait();
                if (m < args.pos_rows) {
                    float v[POS_GC], gr[POS_GC];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + bias[j];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[32 + j] = __uint_as_float(r1[j]) + bias[32 + j];
                    if (args.flags & EPI_SAVE_DGELU) {
#pragma unroll
                        for (int j = 0; j < POS_GC; ++j) v[j] = gelu_erf_with_grad(v[j], gr[j]);
                        uint4* ap = reinterpret_cast<uint4*>(args.aux_out + m * EMBED + g * POS_GC);
#pragma unroll
                        for (int j = 0; j < 6; ++j)
                            ap[j] = make_uint4(pack_op(gr[8 * j], gr[8 * j + 1]), pack_op(gr[8 * j + 2], gr[8 * j + 3]),
                                               pack_op(gr[8 * j + 4], gr[8 * j + 5]), pack_op(gr[8 * j + 6], gr[8 * j + 7]));
                    } else if (args.flags & EPI_GELU) {
#pragma unroll
                        for (int j = 0; j < POS_GC; ++j) v[j] = gelu_act(v[j]);
                    }
                    uint4* op = reinterpret_cast<uint4*>(args.out + m * EMBED + g * POS_GC);
#pragma unroll
                    for (int j = 0; j < 6; ++j)
                        op[j] = make_uint4(pack_op(v[8 * j], v[8 * j + 1]), pack_op(v[8 * j + 2], v[8 * j + 3]),
                                           pack_op(v[8 * j + 4], v[8 * j + 5]), pack_op(v[8 * j + 6], v[8 * j + 7]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[sb]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 encode_fn() {
    static EncodeTiledFn2 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn2>(p);
    });
    return fn;
}

// pos_g: [16][rows_alloc][48]; w: [16][48][6144]; out / aux_out: [pos_rows][768]
int launch_posconv(cudaStream_t st, const op_t* pos_g, long long rows_alloc, long long pos_rows, const op_t* w,
                   const float* bias, int flags, op_t* out, op_t* aux_out) {
    EncodeTiledFn2 fn = encode_fn();
    NB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    static bool attr_set[64] = {false};  // the attribute is per device
    if (bool* flag = device_once_flag(attr_set)) {
        NB_CUDA(cudaFuncSetAttribute(posconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM));
        *flag = true;
    }
    CUtensorMap tmSlab, tmW;
    {
        cuuint64_t gdim[3] = {(cuuint64_t)POS_GC, (cuuint64_t)rows_alloc, (cuuint64_t)POS_G};
        cuuint64_t gstr[2] = {(cuuint64_t)POS_GC * 2, (cuuint64_t)rows_alloc * POS_GC * 2};
        cuuint32_t box[3] = {8, 128, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = fn(&tmSlab, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<op_t*>(pos_g), gdim, gstr, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NB_CHECK(r == CUDA_SUCCESS, "posconv: slab tensor map failed (%d)", (int)r);
    }
    {
        const long long K = (long long)POS_K * POS_GC;
        cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)POS_GC, (cuuint64_t)POS_G};
        cuuint64_t gstr[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * POS_GC * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)POS_GC, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = fn(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<op_t*>(w), gdim, gstr, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NB_CHECK(r == CUDA_SUCCESS, "posconv: weight tensor map failed (%d)", (int)r);
    }
    PosConvArgs a;
    a.pos_rows = pos_rows;
    a.m_tiles = (int)((pos_rows + PC_ROWS - 1) / PC_ROWS);
    a.flags = flags;
    a.bias = bias;
    a.out = out;
    a.aux_out = aux_out;
    const int tiles = a.m_tiles * POS_G;
    int grid = device_sm_count();
    if (tiles < grid) grid = tiles;
    posconv_kernel<<<grid, PC_THREADS, PC_SMEM, st>>>(tmSlab, tmW, a);
    NB_LAUNCHED();
    return 0;
}

}  // namespace nb
