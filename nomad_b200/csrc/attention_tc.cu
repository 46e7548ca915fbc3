// Attention core on tcgen05 for utterances of at most 256 frames (<= 5.1 s; the bench workload has 199).
//
// One CTA (256 threads) = one (utterance, head, 128-query tile).  The whole key range fits one MMA:
//   S = Q K^T     : tcgen05.mma M=128, N=Tp (T rounded up to 16, <= 256), K=64  -> 128 x Tp fp32 in TMEM
//   softmax       : thread r owns query row r = TMEM lane r (two passes over the row straight from TMEM,
//                   no cross-thread reduction), P written as fp16 into shared memory in the 128B-swizzled
//                   K-major operand layout
//   O = P V       : tcgen05.mma M=128, N=64, K=Tp with V as an MN-major B operand (V's natural [key][d]
//                   layout, no transpose), accumulating over the S columns that are no longer needed
//   epilogue      : O / rowsum -> fp16 -> global; optional log-sum-exp for the backward
// Q/K/V tiles come in by TMA (128B swizzle).  256 TMEM columns and 96 KB of shared memory per CTA let two
// CTAs share an SM, so one CTA's softmax overlaps the other's MMAs without intra-CTA pipelining.
// Longer utterances use the streaming mma.sync kernel in encoder.cu.
#include <mutex>

#include "kernels.cuh"

namespace nb {

static constexpr int AT_MAXT = 256;
static constexpr int AT_SMEM = 96 * 1024 + 1024 + 64 + 2048;  // tiles + align slack + barriers + row reductions

struct AttnTcArgs {
    const UttMeta* meta;
    op_t* out;      // frames x 768
    float* lse;     // frames x 12 or nullptr
};

// kind::f16 instruction descriptor with selectable B major-ness (bit 16: 0 = K-major, 1 = MN-major)
__device__ __forceinline__ uint32_t idesc_h16(int m, int n, int b_mn_major) {
    return (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 256 threads: thread (quarter = warp & 3, lane) owns query row quarter * 32 + lane = TMEM lane; the two
// warps that share a TMEM lane quarter (warp and warp + 4) split the key columns between them (even / odd
// 32-column chunks) and exchange row max / row sum through shared memory.
__global__ void __launch_bounds__(256, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTcArgs args) {
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
    const UttMeta m = args.meta[b];
    const int T = m.T;
    if (T > AT_MAXT || q0 >= T) return;
    const int Tp = (T + 15) & ~15;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                 // 128 x 128 B
    uint8_t* sK = smem + 16384;         // 256 x 128 B
    uint8_t* sV = smem + 49152;         // 256 x 128 B
    // P: 4 key blocks of (128 rows x 128 B); blocks 0..2 overlay Q and K (dead once S is computed), block 3 is extra
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 98304);
    uint64_t* bar_load = bars;          // TMA landed
    uint64_t* bar_s = bars + 1;         // S = QK^T done
    uint64_t* bar_o = bars + 2;         // O = PV done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
    float* red = reinterpret_cast<float*>(smem + 98304 + 64);  // [2 halves][128 rows] max, then [2][128] sum
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quarter = warp & 3, half = warp >> 2;
    const int rowl = quarter * 32 + lane;  // local query row

    if (tid == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(bar_load, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_o, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (tid == 0) {
        const int kv_boxes = Tp > 128 ? 2 : 1;
        mbar_expect_tx(bar_load, 16384u * (1 + 2 * kv_boxes));
        const int row0 = m.frame0;
        tma_load_2d(sQ, &tmQKV, bar_load, h * HEAD_DIM, row0 + q0);
        for (int i = 0; i < kv_boxes; ++i) {
            tma_load_2d(sK + i * 16384, &tmQKV, bar_load, EMBED + h * HEAD_DIM, row0 + i * 128);
            tma_load_2d(sV + i * 16384, &tmQKV, bar_load, 2 * EMBED + h * HEAD_DIM, row0 + i * 128);
        }
        mbar_wait(bar_load, 0);
        tc_fence_after();
        const uint64_t dq = umma_desc_sw128(smem_u32(sQ)), dk = umma_desc_sw128(smem_u32(sK));
        const uint32_t id = idesc_h16(128, Tp, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), id, k ? 1u : 0u);
        umma_commit(bar_s);
    }
    mbar_wait(bar_s, 0);
    tc_fence_after();

    // ---- softmax over this thread's chunks (half, half + 2, ...) of its row
    const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
    const int chunks = (Tp + 31) >> 5;
    const float LOG2E = 1.4426950408889634f;
    float mx = -INFINITY;
    {
        // two register buffers, ping-ponged with static names so they stay in registers
        uint32_t ra[32], rb[32];
        auto scan = [&](const uint32_t (&r)[32], int c) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c * 32 + j < T) mx = fmaxf(mx, __uint_as_float(r[j]));
        };
        if (half < chunks) tmem_ld_32x32(trow + half * 32, ra);
        for (int c = half; c < chunks; c += 4) {
            tmem_ld_wait();
            if (c + 2 < chunks) tmem_ld_32x32(trow + (c + 2) * 32, rb);
            scan(ra, c);
            if (c + 2 < chunks) {
                tmem_ld_wait();
                if (c + 4 < chunks) tmem_ld_32x32(trow + (c + 4) * 32, ra);
                scan(rb, c + 2);
            }
        }
    }
    red[half * 128 + rowl] = mx;
    __syncthreads();
    mx = fmaxf(red[rowl], red[128 + rowl]);
    float sum = 0.f;
    const float mscaled = mx * LOG2E;
    {
        uint32_t ra[32], rb[32];
        auto emit = [&](const uint32_t (&r)[32], int c) {
            float p[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float e = ex2_approx(fmaf(__uint_as_float(r[j]), LOG2E, -mscaled));
                p[j] = (c * 32 + j < T) ? e : 0.f;
                sum += p[j];
            }
            // 32 keys = 4 pieces of 16 B in key block (c / 2), piece index base (c & 1) * 4
            const int kb = c >> 1;
            uint8_t* pb = (kb < 3 ? smem + kb * 16384 : smem + 81920) + (rowl >> 3) * 1024 + (rowl & 7) * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int ch = ((c & 1) * 4 + q) ^ (rowl & 7);
                *reinterpret_cast<uint4*>(pb + ch * 16) =
                    make_uint4(pack_op(p[8 * q], p[8 * q + 1]), pack_op(p[8 * q + 2], p[8 * q + 3]),
                               pack_op(p[8 * q + 4], p[8 * q + 5]), pack_op(p[8 * q + 6], p[8 * q + 7]));
            }
        };
        if (half < chunks) tmem_ld_32x32(trow + half * 32, ra);
        for (int c = half; c < chunks; c += 4) {
            tmem_ld_wait();
            if (c + 2 < chunks) tmem_ld_32x32(trow + (c + 2) * 32, rb);
            emit(ra, c);
            if (c + 2 < chunks) {
                tmem_ld_wait();
                if (c + 4 < chunks) tmem_ld_32x32(trow + (c + 4) * 32, ra);
                emit(rb, c + 2);
            }
        }
    }
    red[256 + half * 128 + rowl] = sum;
    // P (generic-proxy writes) must be visible to the tensor core (async proxy); S reads must be done
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t id = idesc_h16(128, HEAD_DIM, 1);  // B = V is MN-major: [key][d], d contiguous
        const int ksteps = Tp >> 4;
        for (int ks = 0; ks < ksteps; ++ks) {
            const int kb = ks >> 2;
            const uint32_t pa = smem_u32(kb < 3 ? smem + kb * 16384 : smem + 81920);
            const uint64_t dp = umma_desc_sw128(pa) + (uint64_t)(2 * (ks & 3));
            const uint64_t dv = umma_desc_sw128(smem_u32(sV) + ks * 2048);  // 16 keys = two 8-row groups
            umma_f16(tmem, dp, dv, id, ks ? 1u : 0u);
        }
        umma_commit(bar_o);
    }
    sum = red[256 + rowl] + red[384 + rowl];
    mbar_wait(bar_o, 0);
    tc_fence_after();
    // ---- epilogue: this thread's 32 of the 64 output columns; staged through the (now dead) Q/K region so
    // that every store instruction writes whole sectors
    {
        uint32_t r0[32];
        tmem_ld_32x32(trow + half * 32, r0);
        tmem_ld_wait();
        const float inv = 1.0f / sum;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) * inv;
        op_t* stage = reinterpret_cast<op_t*>(smem + warp * 2048);
        stage_put_h16(stage, v, lane);
        const int rows_valid = T - (q0 + quarter * 32);
        op_t* wout = args.out + ((long long)m.frame0 + q0 + quarter * 32) * EMBED + h * HEAD_DIM + half * 32;
        stage_flush_h16(stage, wout, EMBED, rows_valid > 32 ? 32 : rows_valid, 32, lane);
        if (args.lse != nullptr && half == 0 && q0 + rowl < T)
            args.lse[((long long)m.frame0 + q0 + rowl) * HEADS + h] = mx + __logf(sum);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int launch_attention_tc(cudaStream_t st, const op_t* qkv, const UttMeta* meta, int B, int max_T, long long frames,
                        op_t* out, float* lse) {
    static EncodeTiledFn3 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn3>(p);
    });
    NB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    static bool attr_set = false;
    if (!attr_set) {
        NB_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        attr_set = true;
    }
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)(3 * EMBED), (cuuint64_t)frames};
    cuuint64_t gstr[1] = {(cuuint64_t)(3 * EMBED) * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<op_t*>(qkv), gdim, gstr, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NB_CHECK(r == CUDA_SUCCESS, "attention: tensor map failed (%d)", (int)r);
    AttnTcArgs a{meta, out, lse};
    const int mt = max_T > AT_MAXT ? AT_MAXT : max_T;
    dim3 grid((mt + 127) / 128, HEADS, B);
    attention_tc_kernel<<<grid, 256, AT_SMEM, st>>>(tm, a);
    NB_LAUNCHED();
    return 0;
}

}  // namespace nb
