// Attention core softmax(Q K^T) V on tcgen05 for utterances of ANY length (non-causal, keys >= T masked).
// q is already scaled by head_dim^-0.5 (folded into the QKV weights).
//
// Work item = (utterance, head, 128-query tile).  A persistent CTA (192 threads, FOUR CTAs per SM: 49 KB of
// shared memory, 128 TMEM columns -- all 512 of the SM --, 80 registers) walks a host-built item list (longest
// utterances first) and streams 64-key tiles through a 2-stage TMA ring:
//   warp 0 (one lane)  TMA producer: Q tile (16 KB), then K_0, V_0, K_1, V_1, ... (8 KB boxes, 128B swizzle)
//   warp 1 (one lane)  MMA issuer:   S = Q K_j^T      tcgen05.mma M=128 N=64 K=64  -> TMEM columns [0, 64)
//                                    O (+)= P_j V_j   tcgen05.mma M=128 N=64 K<=64 -> TMEM columns [64, 128)
//                                    (V in its natural [key][d] layout as an MN-major B operand)
//   warps 2..5         softmax: thread r owns query row r = TMEM lane r.  Per key tile: row max straight from
//                      TMEM, online-softmax rescale of the O row in TMEM (only when some row of the warp saw a
//                      new maximum), P = exp2(.) as fp16 into shared memory in the K-major swizzled operand
//                      layout; after the last tile O / rowsum -> fp16 -> staged, sector-aligned global stores.
// Per CTA the chain S-MMA -> softmax -> PV-MMA is serial in the key tiles (S is single-buffered).  The kernel is
// bound by instruction issue in the softmax warps (ncu: ~9 instructions per score, IPC ~0.45 with two warps
// per scheduler), not by the MMAs or their latency -- double/triple-buffered S, lazy rescaling and deferred
// epilogues were all measured and bought nothing (profiles/r01_attention_notes.md) -- so the design goes for
// occupancy instead: small tiles let four CTAs (16 softmax warps) share an SM and fill each other's gaps
// (round 2: three -> four CTAs by halving the K/V ring, which was measured not to matter: -4..7 %).
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "kernels.cuh"

namespace nb {

static constexpr int FA_THREADS = 192;
static constexpr int FA_BK = 64;                                  // keys per tile
static constexpr int FA_CHUNKS = FA_BK / 32;
#ifndef NB_FA_RING
#define NB_FA_RING 2   // K, V stages of 8 KB: 2 leave room for a fourth CTA per SM (4 and 2 measured equal at 3 CTAs)
#endif
#ifndef NB_FA_CTAS
#define NB_FA_CTAS 4   // 80 registers (48 B of spills), 4 x 128 TMEM columns: 1.57-1.60 -> 1.45-1.53 ms per 12 layers
#endif
static constexpr int FA_RING = NB_FA_RING;
static constexpr int FA_Q_BYTES = 16384;                          // 128 rows x 128 B
static constexpr int FA_TILE_BYTES = FA_BK * 128;                 // K or V tile: FA_BK rows x 128 B
static constexpr int FA_P_BYTES = (FA_BK / 64) * 16384;           // P = 64-key blocks of 128 rows x 128 B
static constexpr int FA_OFF_RING = FA_Q_BYTES;                    // after Q
static constexpr int FA_OFF_P = FA_OFF_RING + FA_RING * FA_TILE_BYTES;
static constexpr int FA_OFF_BAR = FA_OFF_P + FA_P_BYTES;
static constexpr int FA_SMEM = FA_OFF_BAR + 128 + 1024;           // barriers + alignment slack
static constexpr uint32_t FA_TMEM_COLS = 128;
static constexpr uint32_t FA_O_COL = FA_BK;
static constexpr int FA_CTAS_PER_SM = NB_FA_CTAS;

// One (utterance, query tile) of the work list; every entry stands for HEADS work items.
struct FaEntry {
    int frame0;  // first row of the utterance in the frame-level buffers
    int T;       // valid frames
    int q0;      // first query row of the tile
    int utt;
};

struct AttnFaArgs {
    const FaEntry* items;  // longest utterances first
    int n_items;           // entries of items; work items = n_items * HEADS
    op_t* out;             // frames x 768
    float* lse;            // frames x 12 or nullptr
};

__device__ __forceinline__ uint32_t fa_idesc(int m, int n, int b_mn_major) {
    return (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void fa_tma_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

struct FaItem {
    int h, q0, T, n_kv;
    long long frame0;
};
__device__ __forceinline__ FaItem fa_item(const AttnFaArgs& a, int i) {
    const int4 e = __ldg(reinterpret_cast<const int4*>(a.items) + i / HEADS);
    FaItem it;
    it.h = i % HEADS;
    it.frame0 = e.x;
    it.T = e.y;
    it.q0 = e.z;
    it.n_kv = (e.y + FA_BK - 1) / FA_BK;
    return it;
}

// P2: packed fp32 arithmetic (FFMA2 / FADD2 / FMUL2) in the softmax warps -- 12 instead of 16 issue slots per 4 scores
template <bool P2>
__global__ void __launch_bounds__(FA_THREADS, FA_CTAS_PER_SM)
attention_fa_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmKV,
                    const AttnFaArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic on the __shared__ array keeps the address space (LDS / STS, not generic LD / ST)
    uint8_t* sQ = smem;
    uint8_t* sRing = smem + FA_OFF_RING;
    uint8_t* sP = smem + FA_OFF_P;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FA_OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* full = bars + 2;             // [FA_RING]
    uint64_t* empty = bars + 2 + FA_RING;  // [FA_RING]
    uint64_t* s_full = bars + 2 + 2 * FA_RING;
    uint64_t* sp_ready = s_full + 1;  // softmax done with S, P written (and O rescaled)
    uint64_t* pv_done = s_full + 2;   // O updated, P free
    uint64_t* o_free = s_full + 3;    // epilogue finished reading O
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total = args.n_items * HEADS;

    if (tid == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmKV);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < FA_RING; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(sp_ready, 4);
        mbar_init(pv_done, 1);
        mbar_init(o_free, 4);
        mbar_fence_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, FA_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t c = 0, it = 0;
            for (int i = blockIdx.x; i < total; i += gridDim.x, ++it) {
                const FaItem w = fa_item(args, i);
                mbar_wait(q_empty, (it & 1) ^ 1);
                mbar_expect_tx(q_full, FA_Q_BYTES);
                fa_tma_2d(sQ, &tmQKV, q_full, w.h * HEAD_DIM, (int)w.frame0 + w.q0);
                for (int j = 0; j < w.n_kv; ++j) {
#pragma unroll
                    for (int kv = 0; kv < 2; ++kv, ++c) {
                        const uint32_t st = c % FA_RING, ph = (c / FA_RING) & 1;
                        mbar_wait(&empty[st], ph ^ 1);
                        mbar_expect_tx(&full[st], FA_TILE_BYTES);
                        fa_tma_2d(sRing + st * FA_TILE_BYTES, &tmKV, &full[st], (1 + kv) * EMBED + w.h * HEAD_DIM,
                                  (int)w.frame0 + j * FA_BK);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t id_s = fa_idesc(128, FA_BK, 0);
            const uint32_t id_o = fa_idesc(128, HEAD_DIM, 1);  // B = V is MN-major: [key][d], d contiguous
            const uint64_t dq = umma_desc_sw128(smem_u32(sQ));
            uint32_t c = 0, it = 0, g = 0;
            auto issue_pv = [&](const FaItem& w, int j, uint32_t cv) {
                const uint32_t st = cv % FA_RING, ph = (cv / FA_RING) & 1;
                mbar_wait(&full[st], ph);
                if (j == 0 && it > 0) mbar_wait(o_free, (it - 1) & 1);
                tc_fence_after();
                const int nv = min(FA_BK, w.T - j * FA_BK);
                const int ksteps = ((nv + 31) >> 5) * 2;  // the softmax writes whole 32-key chunks (masked keys as 0)
                const uint32_t vbase = smem_u32(sRing + st * FA_TILE_BYTES);
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t dp = umma_desc_sw128(smem_u32(sP) + (ks >> 2) * 16384) + (uint64_t)(2 * (ks & 3));
                    const uint64_t dv = umma_desc_sw128(vbase + ks * 2048);  // 16 keys = two 8-row groups
                    umma_f16(tmem + FA_O_COL, dp, dv, id_o, (j | ks) ? 1u : 0u);
                }
                umma_commit(&empty[st]);
                umma_commit(pv_done);
            };
            for (int i = blockIdx.x; i < total; i += gridDim.x, ++it) {
                const FaItem w = fa_item(args, i);
                mbar_wait(q_full, it & 1);
                for (int j = 0; j < w.n_kv; ++j, ++g) {
                    const uint32_t ck = c + 2 * j;
                    const uint32_t st = ck % FA_RING, ph = (ck / FA_RING) & 1;
                    mbar_wait(&full[st], ph);
                    if (j > 0) mbar_wait(sp_ready, (g - 1) & 1);  // S consumed and P_{j-1} written
                    tc_fence_after();
                    const uint64_t dk = umma_desc_sw128(smem_u32(sRing + st * FA_TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(tmem, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), id_s, k ? 1u : 0u);
                    umma_commit(&empty[st]);
                    if (j == w.n_kv - 1) umma_commit(q_empty);
                    umma_commit(s_full);
                    if (j > 0) issue_pv(w, j - 1, ck - 1);
                }
                mbar_wait(sp_ready, (g - 1) & 1);
                issue_pv(w, w.n_kv - 1, c + 2 * (w.n_kv - 1) + 1);
                c += 2 * w.n_kv;
            }
        }
    } else {
        // ===================== softmax + epilogue (warps 2..5) =====================
        const int quarter = warp & 3;          // TMEM lane quarter this warp may access
        const int rowl = quarter * 32 + lane;  // query row inside the tile
        const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
        const float LOG2E = 1.4426950408889634f;
        uint8_t* prow = sP + (rowl >> 3) * 1024 + (rowl & 7) * 128;  // this row inside a 64-key block of P
        uint32_t it = 0, g = 0;
        for (int i = blockIdx.x; i < total; i += gridDim.x, ++it) {
            const FaItem w = fa_item(args, i);
            const int rows_valid = w.T - (w.q0 + quarter * 32);  // of this warp
            const bool active = rows_valid > 0;                  // warp-uniform
            float m_run = -INFINITY, l_run = 0.f;
            for (int j = 0; j < w.n_kv; ++j, ++g) {
                const int nv = min(FA_BK, w.T - j * FA_BK);
                const int chunks = (nv + 31) >> 5;
                mbar_wait(s_full, g & 1);
                tc_fence_after();
                float mx = -INFINITY;
                if (active) {
                    for (int c = 0; c < chunks; ++c) {
                        uint32_t r[32];
                        tmem_ld_32x32(trow + c * 32, r);
                        tmem_ld_wait();
                        if (c * 32 + 32 > nv) {  // partial chunk: mask once
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                if (c * 32 + k >= nv) r[k] = 0xff800000u;  // -inf
                        }
                        float m2[2] = {mx, -INFINITY};  // FMNMX3: two scores per instruction
#pragma unroll
                        for (int k = 0; k < 32; k += 4) {
                            m2[0] = max3(m2[0], __uint_as_float(r[k]), __uint_as_float(r[k + 1]));
                            m2[1] = max3(m2[1], __uint_as_float(r[k + 2]), __uint_as_float(r[k + 3]));
                        }
                        mx = fmaxf(m2[0], m2[1]);
                    }
                }
                const float m_new = fmaxf(m_run, mx);
                if (j > 0) {
                    mbar_wait(pv_done, (g - 1) & 1);  // P free again, O holds tiles 0..j-1
                    tc_fence_after();
                }
                if (active) {
                    if (j > 0 && __any_sync(0xffffffffu, m_new > m_run)) {
                        const float sc = ex2_approx((m_run - m_new) * LOG2E);
                        l_run *= sc;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            uint32_t r[32];
                            tmem_ld_32x32(trow + FA_O_COL + c * 32, r);
                            tmem_ld_wait();
                            if (P2) {
                                const float2 sc2 = make_float2(sc, sc);
#pragma unroll
                                for (int k = 0; k < 32; k += 2) {
                                    const float2 t = fmul2(make_float2(__uint_as_float(r[k]), __uint_as_float(r[k + 1])), sc2);
                                    r[k] = __float_as_uint(t.x);
                                    r[k + 1] = __float_as_uint(t.y);
                                }
                            } else {
#pragma unroll
                                for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * sc);
                            }
                            tmem_st_32x32(trow + FA_O_COL + c * 32, r);
                        }
                        tmem_st_wait();
                    }
                    const float ms = m_new * LOG2E;
                    float sum = 0.f;
                    for (int c = 0; c < chunks; ++c) {
                        uint32_t r[32];
                        tmem_ld_32x32(trow + c * 32, r);
                        tmem_ld_wait();
                        float p[32];
                        if (P2 && c * 32 + 32 <= nv) {
                            const float2 l2 = make_float2(LOG2E, LOG2E), nms = make_float2(-ms, -ms);
                            float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
                            for (int k = 0; k < 32; k += 4) {
                                const float2 e0 = ffma2(make_float2(__uint_as_float(r[k]), __uint_as_float(r[k + 1])), l2, nms);
                                const float2 e1 = ffma2(make_float2(__uint_as_float(r[k + 2]), __uint_as_float(r[k + 3])), l2, nms);
                                p[k] = ex2_approx(e0.x); p[k + 1] = ex2_approx(e0.y);
                                p[k + 2] = ex2_approx(e1.x); p[k + 3] = ex2_approx(e1.y);
                                a0 = fadd2(a0, make_float2(p[k], p[k + 1]));
                                a1 = fadd2(a1, make_float2(p[k + 2], p[k + 3]));
                            }
                            sum += (a0.x + a0.y) + (a1.x + a1.y);
                        } else if (c * 32 + 32 <= nv) {
                            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int k = 0; k < 32; ++k) {
                                p[k] = ex2_approx(fmaf(__uint_as_float(r[k]), LOG2E, -ms));
                                s4[k & 3] += p[k];
                            }
                            sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
                        } else {
#pragma unroll
                            for (int k = 0; k < 32; ++k) {
                                const float e = ex2_approx(fmaf(__uint_as_float(r[k]), LOG2E, -ms));
                                p[k] = (c * 32 + k < nv) ? e : 0.f;
                                sum += p[k];
                            }
                        }
                        // 32 keys = 4 pieces of 16 B in key block (c / 2), piece index base (c & 1) * 4
                        uint8_t* pb = prow + (c >> 1) * 16384;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int ch = ((c & 1) * 4 + q) ^ (rowl & 7);
                            *reinterpret_cast<uint4*>(pb + ch * 16) =
                                make_uint4(pack_op(p[8 * q], p[8 * q + 1]), pack_op(p[8 * q + 2], p[8 * q + 3]),
                                           pack_op(p[8 * q + 4], p[8 * q + 5]), pack_op(p[8 * q + 6], p[8 * q + 7]));
                        }
                    }
                    l_run += sum;
                    m_run = m_new;
                }
                // P (generic-proxy writes) must be visible to the tensor core (async proxy); S / O accesses done
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(sp_ready);
            }
            // ---- epilogue: O / rowsum -> fp16, staged through this warp's own (now free) rows of P
            mbar_wait(pv_done, (g - 1) & 1);
            tc_fence_after();
            if (active) {
                const float inv = 1.0f / l_run;
                op_t* stage = reinterpret_cast<op_t*>(sP + quarter * 4096);
                op_t* wout = args.out + (w.frame0 + w.q0 + quarter * 32) * EMBED + w.h * HEAD_DIM;
                const int rv = rows_valid > 32 ? 32 : rows_valid;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(trow + FA_O_COL + c * 32, r);
                    tmem_ld_wait();
                    float v[32];
                    if (P2) {
                        const float2 inv2 = make_float2(inv, inv);
#pragma unroll
                        for (int k = 0; k < 32; k += 2) {
                            const float2 t = fmul2(make_float2(__uint_as_float(r[k]), __uint_as_float(r[k + 1])), inv2);
                            v[k] = t.x;
                            v[k + 1] = t.y;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]) * inv;
                    }
                    stage_put_h16(stage, v, lane);
                    stage_flush_h16(stage, wout + c * 32, EMBED, rv, 32, lane);
                }
                if (args.lse != nullptr && lane < rows_valid)
                    args.lse[(w.frame0 + w.q0 + rowl) * HEADS + w.h] = m_run + __logf(l_run);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_free);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, FA_TMEM_COLS);
    }
}

// (utterance, query tile) entries (4 words each, see FaEntry), longest utterances first so the static round-robin over persistent CTAs
// behaves like longest-processing-time-first scheduling.
void build_attention_items(const Plan& p, std::vector<uint32_t>* items) {
    std::vector<int> order(p.B);
    for (int b = 0; b < p.B; ++b) order[b] = b;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return p.utt[a].T > p.utt[b].T; });
    items->clear();
    for (int b : order) {
        const int qt = (p.utt[b].T + 127) / 128;
        for (int q = 0; q < qt; ++q) {  // one FaEntry = 4 words
            items->push_back((uint32_t)p.utt[b].frame0);
            items->push_back((uint32_t)p.utt[b].T);
            items->push_back((uint32_t)(q * 128));
            items->push_back((uint32_t)b);
        }
    }
}

typedef CUresult (*EncodeTiledFnFa)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int launch_attention_fa(cudaStream_t st, const op_t* qkv, const uint32_t* items, int n_items,
                        long long frames, op_t* out, float* lse) {
    static EncodeTiledFnFa fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnFa>(p);
    });
    NB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    static bool attr_set[64] = {false};  // the attribute is per device
    if (bool* flag = device_once_flag(attr_set)) {
        NB_CUDA(cudaFuncSetAttribute(attention_fa_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
        NB_CUDA(cudaFuncSetAttribute(attention_fa_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
        *flag = true;
    }
    static const int packed = getenv("NOMAD_B200_FA_F32X2") ? atoi(getenv("NOMAD_B200_FA_F32X2")) : 1;
    if (n_items <= 0) return 0;
    CUtensorMap tmq, tmkv;
    cuuint64_t gdim[2] = {(cuuint64_t)(3 * EMBED), (cuuint64_t)frames};
    cuuint64_t gstr[1] = {(cuuint64_t)(3 * EMBED) * 2};
    cuuint32_t es[2] = {1, 1};
    for (int k = 0; k < 2; ++k) {
        cuuint32_t box[2] = {64, k == 0 ? 128u : (cuuint32_t)FA_BK};
        CUresult r = fn(k == 0 ? &tmq : &tmkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<op_t*>(qkv), gdim, gstr,
                        box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NB_CHECK(r == CUDA_SUCCESS, "attention: tensor map failed (%d)", (int)r);
    }
    AttnFaArgs a{reinterpret_cast<const FaEntry*>(items), n_items, out, lse};
    const long long total = (long long)n_items * HEADS;
    int grid = FA_CTAS_PER_SM * device_sm_count();
    if (total < grid) grid = (int)total;
    if (packed) attention_fa_kernel<true><<<grid, FA_THREADS, FA_SMEM, st>>>(tmq, tmkv, a);
    else attention_fa_kernel<false><<<grid, FA_THREADS, FA_SMEM, st>>>(tmq, tmkv, a);
    NB_LAUNCHED();
    return 0;
}

}  // namespace nb
