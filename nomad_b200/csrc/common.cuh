// Shared device/host helpers for the nomad_b200 sm_100a kernels.
// PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor) and tcgen05 (UMMA + TMEM).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace nb {

// 16-bit tensor-core operand type.  fp16 (11-bit significand) rather than bf16 (8-bit): same tcgen05
// kind::f16 rate, 8x less rounding noise -- that is what brings embeddings within 1e-3 of the fp32
// reference (bf16 operands measured 2-3e-3, profiles/r01_parity_probe_bf16.log).  The residual stream,
// all normalisation statistics, softmax and the accumulators stay fp32; conversions saturate.
typedef __half op_t;
static constexpr float OP_MAX = 65504.0f;
__host__ __device__ __forceinline__ op_t f2op(float v) {
    v = v > OP_MAX ? OP_MAX : (v < -OP_MAX ? -OP_MAX : v);
    return __float2half_rn(v);
}
__host__ __device__ __forceinline__ float op2f(op_t v) { return __half2float(v); }

// ---------------------------------------------------------------- error handling (host)
void set_error(const char* fmt, ...);
const char* get_error();

#define NB_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            nb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

#define NB_CHECK(cond, ...)                \
    do {                                   \
        if (!(cond)) {                     \
            nb::set_error(__VA_ARGS__);    \
            return 1;                      \
        }                                  \
    } while (0)

// every kernel launch of the library goes through this so bench.py can report gpu_launches
void count_launch();
void count_launches(long long n);  // graph replays: n kernels at once
#define NB_LAUNCHED()                 \
    do {                              \
        nb::count_launch();           \
        NB_CUDA(cudaGetLastError());  \
    } while (0)

// "done once per device" state for per-device settings (cudaFuncSetAttribute is per device/context): returns the flag
// of the CURRENT device if it is still unset, nullptr otherwise.
inline bool* device_once_flag(bool (&flags)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    return flags[dev] ? nullptr : &flags[dev];
}

#define NB_TRY(expr)            \
    do {                        \
        int _r = (expr);        \
        if (_r) return _r;      \
    } while (0)

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t.reg .b32 R1;\n\t"
        "elect.sync R1|P1, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float gelu_erf_exact(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// erf-GELU with erf from Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7): 2 MUFU + ~12 FMA-pipe
// instructions instead of libdevice erff's ~30.  With y = |x| sqrt(log2(e) / 2):
//   exp(-x^2/2) = 2^(-y^2),  t = 1 / (1 + p |x| / sqrt 2) = 1 / (1 + P y),  erf = 1 - poly(t) 2^(-y^2)
//   gelu(x) = 0.5 (x + |x| erf(|x| / sqrt 2))
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define NB_GELU_K1 0.84932180028801904272f  /* sqrt(log2(e) / 2) */
#define NB_GELU_P 0.2727374808792225f        /* 0.3275911 / sqrt(log2(e)) */
__device__ __forceinline__ float gelu_erf(float x) {
    const float ax = fabsf(x);
    const float y = ax * NB_GELU_K1;
    const float t = rcp_approx(fmaf(NB_GELU_P, y, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    const float e = fmaf(-p, ex2_approx(-y * y), 1.0f);
    return 0.5f * fmaf(ax, e, x);
}
// erf-GELU as x * sigmoid(2 u), u = x (c0 + c1 x^2 + c2 x^4) with x^2 clamped to 36: a minimax fit of the
// argument (tools/fit_gelu.py) whose max |error| against the exact erf form is 2.5e-5 over all x -- below the
// fp16 rounding of any output above 0.05 -- for 7 FMA-pipe + 2 MUFU instructions instead of 11 + 2.  The
// coefficients carry the factor -2 log2(e) so the sigmoid is one ex2 and one rcp; both ends saturate cleanly
// (ex2 -> 0 gives x, ex2 -> inf gives -0).  Used where only the activation is needed (scoring path); the loss
// path keeps gelu_erf_with_grad, which shares its erf / exp between the value and the derivative.
#ifndef NB_GELU_FAST
#define NB_GELU_FAST 1
#endif
__device__ __forceinline__ float gelu_fast(float x) {
    const float x2 = fminf(x * x, 36.0f);
    float p = fmaf(1.0142630552e-03f, x2, -1.0677572400e-01f);   // -2 log2(e) * (c2, c1)
    p = fmaf(p, x2, -2.3011213395e+00f);                           // -2 log2(e) * c0
    const float e = ex2_approx(x * p);
    return x * rcp_approx(1.0f + e);
}
// gelu_fast on a register pair with packed fp32 arithmetic: 6 packed + 2 FMNMX + 4 MUFU for two values instead of
// 2 x (7 + 2) scalar instructions (the GELU epilogues are bound by instruction issue)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c);
__device__ __forceinline__ float2 fadd2(float2 a, float2 b);
__device__ __forceinline__ float2 fmul2(float2 a, float2 b);
#ifndef NB_F32X2
#define NB_F32X2 1   // 0: scalar fp32 arithmetic in the GELU / LayerNorm-fold epilogues (A/B builds)
#endif
__device__ __forceinline__ float gelu_fast(float x);
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
#if !NB_F32X2
    return make_float2(gelu_fast(x.x), gelu_fast(x.y));
#endif
    float2 x2 = fmul2(x, x);
    x2.x = fminf(x2.x, 36.0f);
    x2.y = fminf(x2.y, 36.0f);
    float2 p = ffma2(make_float2(1.0142630552e-03f, 1.0142630552e-03f), x2, make_float2(-1.0677572400e-01f, -1.0677572400e-01f));
    p = ffma2(p, x2, make_float2(-2.3011213395e+00f, -2.3011213395e+00f));
    const float2 u = fmul2(x, p);
    const float2 e = fadd2(make_float2(ex2_approx(u.x), ex2_approx(u.y)), make_float2(1.0f, 1.0f));
    return fmul2(x, make_float2(rcp_approx(e.x), rcp_approx(e.y)));
}
__device__ __forceinline__ float gelu_act(float x) {
#if NB_GELU_FAST
    return gelu_fast(x);
#else
    return gelu_erf(x);
#endif
}
// gelu(x) and d gelu / dx = Phi(x) + x phi(x) from one erf / one exp evaluation
__device__ __forceinline__ float gelu_erf_with_grad(float x, float& grad) {
    const float ax = fabsf(x);
    const float y = ax * NB_GELU_K1;
    const float t = rcp_approx(fmaf(NB_GELU_P, y, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    const float ex = ex2_approx(-y * y);      // exp(-x^2 / 2)
    const float e = fmaf(-p, ex, 1.0f);       // erf(|x| / sqrt 2)
    const float cdf = 0.5f + copysignf(0.5f * e, x);
    grad = fmaf(x * 0.39894228040143267794f, ex, cdf);
    return 0.5f * fmaf(ax, e, x);
}
// d/dx [0.5 x (1 + erf(x/sqrt2))] = Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}
// Packed fp32 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2, two fp32 operations per issued instruction).  The softmax and
// epilogue warps of this library are bound by instruction issue, not by a pipe, so halving their FMA-pipe instruction
// count is a direct gain.  Operands are register pairs; adjacent registers pack for free.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    uint64_t ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    uint64_t ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    uint64_t ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint32_t pack_op(float lo, float hi) {
    uint32_t r;  // one F2FP.SATFINITE: round to nearest even, clamp to +-65504 instead of overflowing to inf
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 unpack_op(uint32_t u) {
    __half2 v = *reinterpret_cast<__half2*>(&u);
    return __half22float2(v);
}
// Two GELUs straight to a packed fp16 pair with ONE MUFU for both: sigma(2u) = (1 + tanh u) / 2 with the same
// minimax u(x), tanh.approx.f16x2 on the packed arguments, and the final 0.5 x (1 + t) as one HFMA2.  Error is that
// of fp16 tanh (~5e-4 absolute on t, i.e. <= 2.5e-4 |x| on the result): about one extra fp16 rounding.  Only used
// where the MUFU pipe is the bound and the result is stored as fp16 anyway (conv0: 1.7 G activations per step).
__device__ __forceinline__ uint32_t gelu_pair_h2(float x0, float x1) {
#if !NB_F32X2
    {
        const float a0 = fminf(x0 * x0, 36.0f), a1 = fminf(x1 * x1, 36.0f);
        float p0 = fmaf(-3.515167885e-04f, a0, 3.700564602e-02f), p1 = fmaf(-3.515167885e-04f, a1, 3.700564602e-02f);
        p0 = fmaf(p0, a0, 7.975078843e-01f);
        p1 = fmaf(p1, a1, 7.975078843e-01f);
        const uint32_t u = pack_op(x0 * p0, x1 * p1);
        uint32_t t;
        asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(u));
        const uint32_t hx = pack_op(0.5f * x0, 0.5f * x1);
        uint32_t r;
        asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(r) : "r"(hx), "r"(t));
        return r;
    }
#endif
    const float2 x = make_float2(x0, x1);
    float2 a = fmul2(x, x);
    a.x = fminf(a.x, 36.0f);
    a.y = fminf(a.y, 36.0f);
    float2 p = ffma2(make_float2(-3.515167885e-04f, -3.515167885e-04f), a, make_float2(3.700564602e-02f, 3.700564602e-02f));
    p = ffma2(p, a, make_float2(7.975078843e-01f, 7.975078843e-01f));
    const float2 xp = fmul2(x, p), hxf = fmul2(x, make_float2(0.5f, 0.5f));
    const uint32_t u = pack_op(xp.x, xp.y);
    uint32_t t;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(u));
    const uint32_t hx = pack_op(hxf.x, hxf.y);
    uint32_t r;
    asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(r) : "r"(hx), "r"(t));
    return r;
}

// ---------------------------------------------------------------- row-per-lane -> coalesced stores via smem
static constexpr int STAGE_BYTES_PER_WARP = 4096;

__device__ __forceinline__ void stage_put_f32(float* stage, const float (&v)[32], int lane) {
#pragma unroll
    for (int p = 0; p < 8; ++p)
        *reinterpret_cast<float4*>(stage + lane * 32 + ((p ^ (lane & 7)) << 2)) =
            make_float4(v[4 * p], v[4 * p + 1], v[4 * p + 2], v[4 * p + 3]);
}
// out points at (first row of this warp, first column of the chunk)
__device__ __forceinline__ void stage_flush_f32(const float* stage, float* out, long long ld, int rows_valid, int ncols,
                                                int lane) {
    __syncwarp();
    const int p = lane & 7;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int rl = k * 4 + (lane >> 3);
        if (rl < rows_valid && p * 4 < ncols)
            *reinterpret_cast<float4*>(out + rl * ld + p * 4) =
                *reinterpret_cast<const float4*>(stage + rl * 32 + ((p ^ (rl & 7)) << 2));
    }
    __syncwarp();
}
// Reverse direction for an fp32 tile of which every lane needs one ROW (the residual stream in a GEMM epilogue):
// cp.async moves 16-byte pieces global -> shared with consecutive lanes on consecutive pieces of the same row
// (whole sectors per request, no registers held while the data is in flight); after stage_fill_wait each lane reads
// its own row back (swizzled, conflict-free).  All 32 rows x 32 columns must be valid.
__device__ __forceinline__ void stage_fill_f32_async(float* stage, const float* src, long long ld, int lane) {
    const int p = lane & 7;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int rl = k * 4 + (lane >> 3);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(stage + rl * 32 + ((p ^ (rl & 7)) << 2))),
                     "l"(src + rl * ld + p * 4)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_fill_wait() {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
}
__device__ __forceinline__ float4 stage_row_f32(const float* stage, int lane, int p) {
    return *reinterpret_cast<const float4*>(stage + lane * 32 + ((p ^ (lane & 7)) << 2));
}
__device__ __forceinline__ void stage_put_h16(op_t* stage, const float (&v)[32], int lane) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
        *reinterpret_cast<uint4*>(stage + lane * 32 + ((p ^ ((lane >> 1) & 3)) << 3)) =
            make_uint4(pack_op(v[8 * p], v[8 * p + 1]), pack_op(v[8 * p + 2], v[8 * p + 3]),
                       pack_op(v[8 * p + 4], v[8 * p + 5]), pack_op(v[8 * p + 6], v[8 * p + 7]));
}
__device__ __forceinline__ void stage_flush_h16(const op_t* stage, op_t* out, long long ld, int rows_valid, int ncols,
                                                int lane) {
    __syncwarp();
    const int p = lane & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int rl = k * 8 + (lane >> 2);
        if (rl < rows_valid && p * 8 < ncols)
            *reinterpret_cast<uint4*>(out + rl * ld + p * 8) =
                *reinterpret_cast<const uint4*>(stage + rl * 32 + ((p ^ ((rl >> 1) & 3)) << 3));
    }
    __syncwarp();
}

// L2 prefetch of `bytes` (multiple of 16) contiguous global bytes: no registers, no completion to wait for
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// NB_MBAR_HINT_NS > 0: try_wait carries a suspend-time hint, so a waiting thread sleeps in hardware until the phase completes
// (or the hint runs out) instead of re-issuing the test: the spin loops of the attention kernel's softmax warps were HALF of
// all instructions it issued (profiles/r02b_ncu_misc: BRA + SYNCS + YIELD = 10 of 22 thread instructions per score).
#ifndef NB_MBAR_HINT_NS
#define NB_MBAR_HINT_NS 20000   // measured (alternating builds, one box): step 16.40-16.54 -> 16.16-16.17 ms, FC1 228-238 -> 214 us, QKV 171 -> 163 us
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
#if NB_MBAR_HINT_NS > 0
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)NB_MBAR_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
#endif
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tile load into this CTA's shared memory, completion on a CTA-local mbarrier
__device__ __forceinline__ void tma_load_2d_cta(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// TMA store of one box from shared memory (layout = the tensor map's swizzle) to global memory; coordinates past the
// tensor's extent are clipped by the hardware.  Issued by ONE lane after the warp's generic-proxy writes to the box
// have been fenced into the async proxy; the box may be overwritten once bulk_wait_read has returned.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- CTA pairs (cluster of 2, tcgen05 cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// In the shared::cluster window bit 24 of an address selects the odd CTA of a pair; clearing it turns "my
// barrier" into "the same barrier in the even (leader) CTA" (CUTLASS Sm100MmaPeerBitMask).
static constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
// TMA load issued by either CTA of a pair, completing on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                 int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// arrive on the same-named mbarrier of CTA `rank` of this cluster.  Default semantics (release at CTA scope), as
// CUTLASS's ClusterBarrier::arrive does: the waiter only needs the tcgen05-fenced TMEM reads ordered, and a
// cluster-scope release made lane 0 wait for every global store of its epilogue to drain (ncu: membar stalls).
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of the pair's MMAs, arriving on the same-named mbarrier in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
// D[tmem of both CTAs] (+)= A (256 x 16, 128 rows per CTA) * B^T (N x 16, N/2 rows per CTA); leader issues
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers fp16/bf16 operands with fp32 accumulate.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 128-byte-swizzled operand tile (rows of 128 B, 8-row groups 1024 B apart) -> UMMA smem descriptor.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address   [0,14)
    d |= (uint64_t)1 << 16;                        // LBO (ignored for swizzled K-major) [16,30)
    d |= (uint64_t)(1024u >> 4) << 32;             // SBO = 1024 B    [32,46)
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: fp16 x fp16 (a/b format 0) -> fp32 (c format 1), both operands K-major.
__host__ __device__ __forceinline__ uint32_t umma_idesc_h16(int m, int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = lane i of this warp's quarter).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace nb
