// NOMAD loss (reference nomad.py:142-146, 243-282) forward + backward to the estimate waveform.
//
// The 2B utterances (B estimates followed by B clean references, all N samples) run through ONE forward
// pass in "save" mode; the loss is the sum of 13 mean-L1 terms between the two halves; the backward
// walks the estimate half only (dgrad chain, no weight gradients).  All gradient tensors carry a
// power-of-two scale S (chosen so the L1 seeds are O(1)) to stay inside fp16's normal range; the final
// kernel divides it out.  fairseq's GradMultiply(feature_grad_mult) on the conv features is applied at
// the LayerNorm(512) boundary.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "../../include/nomad_b200.h"
#include "kernels.cuh"

namespace nb {

int launch_attention_bwd(cudaStream_t st, const op_t* qkv, const op_t* attn_out, const op_t* d_out, const float* lse,
                         float* D, const UttMeta* meta, int B, int max_T, long long frames, op_t* d_qkv);

// ---------------------------------------------------------------------------------------------
struct Row768f {
    float v[24];
};
__device__ __forceinline__ void load_row768(Row768f& r, const float* __restrict__ p, int lane) {
#pragma unroll
    for (int h = 0; h < 3; ++h) {
        const float4* xp = reinterpret_cast<const float4*>(p + (lane + 32 * h) * 8);
        const float4 a = __ldg(xp), c = __ldg(xp + 1);
        float* v = r.v + 8 * h;
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    }
}
__device__ __forceinline__ void load_vec768(Row768f& r, const float* __restrict__ p, int lane) { load_row768(r, p, lane); }
// x -> xhat (in place), returns rstd
__device__ __forceinline__ float normalise768(Row768f& r) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) s += r.v[i];
    const float mean = warp_sum(s) * (1.0f / EMBED);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) { r.v[i] = __fsub_rn(r.v[i], mean); q = fmaf(r.v[i], r.v[i], q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / EMBED) + 1e-5f);
    // explicit rounding: the compiler must not contract this product into a later (est - clean) subtraction,
    // or identical inputs stop giving an exactly zero L1 term / zero gradient sign
#pragma unroll
    for (int i = 0; i < 24; ++i) r.v[i] = __fmul_rn(r.v[i], rstd);
    return rstd;
}
__device__ __forceinline__ void store_row768(const Row768f& r, float* __restrict__ x, op_t* __restrict__ xh, int lane) {
#pragma unroll
    for (int h = 0; h < 3; ++h) {
        const int c0 = (lane + 32 * h) * 8;
        const float* v = r.v + 8 * h;
        if (x != nullptr) {
            float4* xo = reinterpret_cast<float4*>(x + c0);
            xo[0] = make_float4(v[0], v[1], v[2], v[3]);
            xo[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (xh != nullptr)
            *reinterpret_cast<uint4*>(xh + c0) =
                make_uint4(pack_op(v[0], v[1]), pack_op(v[2], v[3]), pack_op(v[4], v[5]), pack_op(v[6], v[7]));
    }
}

// ---------------------------------------------------------------------------------------------
// One L1 term between the two halves of a layer output x = LN(pre):  acc += sum |x_est - x_clean|
__global__ void __launch_bounds__(256) l1_layer_kernel(const float* __restrict__ pre, const UttMeta* __restrict__ meta,
                                                       int B_est, long long frames_est, const float* __restrict__ g,
                                                       double* __restrict__ acc) {
    const long long f = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    float tot = 0.f;
    if (f < frames_est) {
        const int b = find_utt_by_frame(meta, B_est, (int)f);
        if ((int)f - meta[b].frame0 < meta[b].T) {
            Row768f e, c, gg;
            load_row768(e, pre + f * EMBED, lane);
            load_row768(c, pre + (f + frames_est) * EMBED, lane);
            load_vec768(gg, g, lane);
            normalise768(e);
            normalise768(c);
#pragma unroll
            for (int i = 0; i < 24; ++i) tot += fabsf(__fsub_rn(e.v[i], c.v[i]) * gg.v[i]);  // beta cancels
        }
    }
    tot = warp_sum(tot);
    __shared__ float red[8];
    if (lane == 0) red[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w];
        if (s != 0.f) atomicAdd(acc, (double)s);
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm(768) backward on the estimate rows, with the optional L1 seed of this layer's output and an
// optional per-utterance broadcast gradient (mean-pool of the head).
//   g = g_in + seed * sign(x_est - x_clean) + g_pool[utt]
//   g_pre = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat))
__global__ void __launch_bounds__(256) ln768_bwd_kernel(const float* __restrict__ g_in, const float* __restrict__ pre,
                                                        const UttMeta* __restrict__ meta, int B_est, long long frames_est,
                                                        const float* __restrict__ gam, float seed,
                                                        const float* __restrict__ g_pool, float* __restrict__ g_out,
                                                        op_t* __restrict__ g_out_h) {
    const long long f = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (f >= frames_est) return;
    const int lane = threadIdx.x & 31;
    const int b = find_utt_by_frame(meta, B_est, (int)f);
    Row768f out;
    if ((int)f - meta[b].frame0 >= meta[b].T) {
#pragma unroll
        for (int i = 0; i < 24; ++i) out.v[i] = 0.f;
        store_row768(out, g_out ? g_out + f * EMBED : nullptr, g_out_h ? g_out_h + f * EMBED : nullptr, lane);
        return;
    }
    Row768f xh, gg, g;
    load_row768(xh, pre + f * EMBED, lane);
    load_vec768(gg, gam, lane);
    const float rstd = normalise768(xh);
    if (g_in != nullptr) {
        load_row768(g, g_in + f * EMBED, lane);
    } else {
#pragma unroll
        for (int i = 0; i < 24; ++i) g.v[i] = 0.f;
    }
    if (seed != 0.f) {
        Row768f xc;
        load_row768(xc, pre + (f + frames_est) * EMBED, lane);
        normalise768(xc);
#pragma unroll
        for (int i = 0; i < 24; ++i) {
            const float d = __fsub_rn(xh.v[i], xc.v[i]) * gg.v[i];  // x_est - x_clean (beta cancels)
            g.v[i] += d > 0.f ? seed : (d < 0.f ? -seed : 0.f);
        }
    }
    if (g_pool != nullptr) {
        Row768f gp;
        load_vec768(gp, g_pool + (long long)b * EMBED, lane);
#pragma unroll
        for (int i = 0; i < 24; ++i) g.v[i] += gp.v[i];
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        g.v[i] *= gg.v[i];
        s1 += g.v[i];
        s2 = fmaf(g.v[i], xh.v[i], s2);
    }
    s1 = warp_sum(s1) * (1.0f / EMBED);
    s2 = warp_sum(s2) * (1.0f / EMBED);
#pragma unroll
    for (int i = 0; i < 24; ++i) out.v[i] = rstd * (g.v[i] - s1 - xh.v[i] * s2);
    store_row768(out, g_out ? g_out + f * EMBED : nullptr, g_out_h ? g_out_h + f * EMBED : nullptr, lane);
}

// ---------------------------------------------------------------------------------------------
// Head term: e = normalize(W relu(pooled) + b) for both halves; acc += sum |e_est - e_clean|;
// g_pool[b][k] = d(term) / d x[t, k] for every valid frame t of estimate b (already divided by T).
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ pooled, const UttMeta* __restrict__ meta,
                                                       int B_est, const float* __restrict__ head_wt,
                                                       const float* __restrict__ head_w, const float* __restrict__ head_b,
                                                       float seed, double* __restrict__ acc, float* __restrict__ g_pool) {
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ float pe[EMBED], pc[EMBED], gy[EMB];
    __shared__ float red[3][8];
    for (int k = tid; k < EMBED; k += 256) {
        pe[k] = pooled[(long long)b * EMBED + k];
        pc[k] = pooled[(long long)(b + B_est) * EMBED + k];
    }
    __syncthreads();
    float ye = head_b[tid], yc = ye;
    for (int k = 0; k < EMBED; ++k) {
        const float w = __ldg(head_wt + k * EMB + tid);
        ye = fmaf(fmaxf(pe[k], 0.f), w, ye);
        yc = fmaf(fmaxf(pc[k], 0.f), w, yc);
    }
    float se = warp_sum(ye * ye), sc = warp_sum(yc * yc);
    if ((tid & 31) == 0) { red[0][tid >> 5] = se; red[1][tid >> 5] = sc; }
    __syncthreads();
    float ne = 0.f, nc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { ne += red[0][w]; nc += red[1][w]; }
    ne = fmaxf(sqrtf(ne), 1e-12f);
    nc = fmaxf(sqrtf(nc), 1e-12f);
    const float ee = ye / ne, ec = yc / nc;
    const float d = ee - ec;
    const float ge = d > 0.f ? seed : (d < 0.f ? -seed : 0.f);
    float l1 = warp_sum(fabsf(d)), dot = warp_sum(ge * ee);
    __syncthreads();
    if ((tid & 31) == 0) { red[0][tid >> 5] = l1; red[2][tid >> 5] = dot; }
    __syncthreads();
    float l1t = 0.f, dott = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { l1t += red[0][w]; dott += red[2][w]; }
    if (tid == 0) atomicAdd(acc, (double)l1t);
    gy[tid] = (ge - ee * dott) / ne;  // backward of x / max(||x||, eps)
    __syncthreads();
    const float invT = 1.0f / (float)meta[b].T;
    for (int k = tid; k < EMBED; k += 256) {
        float a = 0.f;
        for (int o = 0; o < EMB; ++o) a = fmaf(gy[o], __ldg(head_w + (long long)o * EMBED + k), a);
        g_pool[(long long)b * EMBED + k] = pe[k] > 0.f ? a * invT : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// Positional-conv backward, step 1: encoder LayerNorm backward (z = x0 + pos_y recomputed) -> g_z (fp32),
// and scatter of g_z * gelu'(pos pre-activation) into the grouped, zero-padded layout for the dgrad GEMM.
__global__ void __launch_bounds__(256) pos_bwd_prep_kernel(const float* __restrict__ g_x, const float* __restrict__ x0,
                                                           const op_t* __restrict__ pos_y, const op_t* __restrict__ pos_aux,
                                                           const UttMeta* __restrict__ meta, int B_est, long long frames_est,
                                                           const float* __restrict__ gam, long long pos_rows_alloc,
                                                           float* __restrict__ g_z, op_t* __restrict__ pos_g) {
    const long long f = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (f >= frames_est) return;
    const int lane = threadIdx.x & 31;
    const int b = find_utt_by_frame(meta, B_est, (int)f);
    const int t = (int)f - meta[b].frame0;
    Row768f out;
    if (t >= meta[b].T) {
#pragma unroll
        for (int i = 0; i < 24; ++i) out.v[i] = 0.f;
        store_row768(out, g_z + f * EMBED, nullptr, lane);
        return;
    }
    const long long m = (long long)meta[b].pos0 + t - POS_K / 2;  // row of this frame in pos_y / pos_aux
    Row768f z, g, gg;
    load_row768(z, x0 + f * EMBED, lane);
    float aux[24];
#pragma unroll
    for (int h = 0; h < 3; ++h) {
        const int c0 = (lane + 32 * h) * 8;
        const uint4 y = __ldg(reinterpret_cast<const uint4*>(pos_y + m * EMBED + c0));
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(pos_aux + m * EMBED + c0));
        float2 t0 = unpack_op(y.x), t1 = unpack_op(y.y), t2 = unpack_op(y.z), t3 = unpack_op(y.w);
        float* v = z.v + 8 * h;
        v[0] += t0.x; v[1] += t0.y; v[2] += t1.x; v[3] += t1.y; v[4] += t2.x; v[5] += t2.y; v[6] += t3.x; v[7] += t3.y;
        t0 = unpack_op(a.x); t1 = unpack_op(a.y); t2 = unpack_op(a.z); t3 = unpack_op(a.w);
        float* q = aux + 8 * h;
        q[0] = t0.x; q[1] = t0.y; q[2] = t1.x; q[3] = t1.y; q[4] = t2.x; q[5] = t2.y; q[6] = t3.x; q[7] = t3.y;
    }
    const float rstd = normalise768(z);
    load_row768(g, g_x + f * EMBED, lane);
    load_vec768(gg, gam, lane);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        g.v[i] *= gg.v[i];
        s1 += g.v[i];
        s2 = fmaf(g.v[i], z.v[i], s2);
    }
    s1 = warp_sum(s1) * (1.0f / EMBED);
    s2 = warp_sum(s2) * (1.0f / EMBED);
#pragma unroll
    for (int i = 0; i < 24; ++i) out.v[i] = rstd * (g.v[i] - s1 - z.v[i] * s2);
    store_row768(out, g_z + f * EMBED, nullptr, lane);
    const long long p = (long long)meta[b].pos0 + t;
#pragma unroll
    for (int h = 0; h < 3; ++h) {
        const int ch = (lane + 32 * h) * 8, grp = ch / POS_GC, cc = ch % POS_GC;
        const float* v = out.v + 8 * h;
        const float* q = aux + 8 * h;
        *reinterpret_cast<uint4*>(pos_g + ((long long)grp * pos_rows_alloc + p) * POS_GC + cc) =
            make_uint4(pack_op(v[0] * q[0], v[1] * q[1]), pack_op(v[2] * q[2], v[3] * q[3]),
                       pack_op(v[4] * q[4], v[5] * q[5]), pack_op(v[6] * q[6], v[7] * q[7]));
    }
}

// step 2 (after the dgrad GEMM): g_x0 = g_z + dgrad[pos0 + t - 63]  -> 16-bit operand of the projection dgrad
__global__ void __launch_bounds__(256) pos_bwd_finish_kernel(const float* __restrict__ g_z, const op_t* __restrict__ pos_dy,
                                                             const UttMeta* __restrict__ meta, int B_est,
                                                             long long frames_est, op_t* __restrict__ g_x0h) {
    const long long f = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (f >= frames_est) return;
    const int lane = threadIdx.x & 31;
    const int b = find_utt_by_frame(meta, B_est, (int)f);
    const int t = (int)f - meta[b].frame0;
    Row768f r;
    if (t >= meta[b].T) {
#pragma unroll
        for (int i = 0; i < 24; ++i) r.v[i] = 0.f;
    } else {
        load_row768(r, g_z + f * EMBED, lane);
        const long long m = (long long)meta[b].pos0 + t - (POS_K / 2 - 1);
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const uint4 y = __ldg(reinterpret_cast<const uint4*>(pos_dy + m * EMBED + (lane + 32 * h) * 8));
            const float2 t0 = unpack_op(y.x), t1 = unpack_op(y.y), t2 = unpack_op(y.z), t3 = unpack_op(y.w);
            float* v = r.v + 8 * h;
            v[0] += t0.x; v[1] += t0.y; v[2] += t1.x; v[3] += t1.y; v[4] += t2.x; v[5] += t2.y; v[6] += t3.x; v[7] += t3.y;
        }
    }
    store_row768(r, nullptr, g_x0h + f * EMBED, lane);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm(512) backward fused with GradMultiply and the GELU gradient of conv layer 6:
//   g_u6 = fgm * LN_bwd(g; y6) * gelu'(u6)      (aux6 is zero on padding rows)
__global__ void __launch_bounds__(256) ln512_bwd_kernel(const float* __restrict__ g_in, const op_t* __restrict__ y6,
                                                        const op_t* __restrict__ aux6, long long rows,
                                                        const float* __restrict__ gam, float fgm, op_t* __restrict__ g_u6) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float x[16], g[16], a[16];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c0 = (lane + 32 * h) * 8;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(y6 + row * CONV_DIM + c0));
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(aux6 + row * CONV_DIM + c0));
        float2 f;
        f = unpack_op(u.x); x[8 * h + 0] = f.x; x[8 * h + 1] = f.y;
        f = unpack_op(u.y); x[8 * h + 2] = f.x; x[8 * h + 3] = f.y;
        f = unpack_op(u.z); x[8 * h + 4] = f.x; x[8 * h + 5] = f.y;
        f = unpack_op(u.w); x[8 * h + 6] = f.x; x[8 * h + 7] = f.y;
        f = unpack_op(w.x); a[8 * h + 0] = f.x; a[8 * h + 1] = f.y;
        f = unpack_op(w.y); a[8 * h + 2] = f.x; a[8 * h + 3] = f.y;
        f = unpack_op(w.z); a[8 * h + 4] = f.x; a[8 * h + 5] = f.y;
        f = unpack_op(w.w); a[8 * h + 6] = f.x; a[8 * h + 7] = f.y;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(g_in + row * CONV_DIM + c0));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(g_in + row * CONV_DIM + c0 + 4));
        const float4 m0 = __ldg(reinterpret_cast<const float4*>(gam + c0)), m1 = __ldg(reinterpret_cast<const float4*>(gam + c0 + 4));
        g[8 * h + 0] = g0.x * m0.x; g[8 * h + 1] = g0.y * m0.y; g[8 * h + 2] = g0.z * m0.z; g[8 * h + 3] = g0.w * m0.w;
        g[8 * h + 4] = g1.x * m1.x; g[8 * h + 5] = g1.y * m1.y; g[8 * h + 6] = g1.z * m1.z; g[8 * h + 7] = g1.w * m1.w;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    const float mean = warp_sum(s) * (1.0f / CONV_DIM);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] -= mean; q = fmaf(x[i], x[i], q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / CONV_DIM) + 1e-5f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] *= rstd; s1 += g[i]; s2 = fmaf(g[i], x[i], s2); }
    s1 = warp_sum(s1) * (1.0f / CONV_DIM);
    s2 = warp_sum(s2) * (1.0f / CONV_DIM);
    const float k = rstd * fgm;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = k * (g[8 * h + i] - s1 - x[8 * h + i] * s2) * a[8 * h + i];
        *reinterpret_cast<uint4*>(g_u6 + row * CONV_DIM + (lane + 32 * h) * 8) =
            make_uint4(pack_op(r[0], r[1]), pack_op(r[2], r[3]), pack_op(r[4], r[5]), pack_op(r[6], r[7]));
    }
}

// ---------------------------------------------------------------------------------------------
// conv0 + GroupNorm backward.  G0[t, c] = (dL/du0)[t, c] * gamma_c * rstd_c is produced by the level-1
// dgrad GEMM (aux0 carries gelu' * gamma * rstd).  With uhat = (conv0 - mean) * rstd:
//   dL/dconv0[t, c] = G0 - mean_t(G0) - uhat * mean_t(G0 * uhat)
// Pass 1: per (utt, channel) sums of G0 and G0 * uhat.
__global__ void __launch_bounds__(256) conv0_bwd_stats_kernel(const op_t* __restrict__ G0, const float* __restrict__ wav,
                                                              const UttMeta* __restrict__ meta, int B_est,
                                                              const float* __restrict__ w0, const float* __restrict__ stat,
                                                              double* __restrict__ sums) {
    const int row_base = blockIdx.x * 64;
    const int b = find_utt_by_frame(meta, B_est, blockIdx.x);
    const UttMeta m = meta[b];
    const int t_base = row_base - m.row0;
    const int valid = min(64, m.T0 - t_base);
    if (valid <= 0) return;
    __shared__ float xs[64 * 5 + 8];
    const float* x = wav + m.wav_off;
    for (int i = threadIdx.x; i < 64 * 5 + 5; i += blockDim.x) {
        const long long s = (long long)t_base * 5 + i;
        xs[i] = (s < m.n) ? __ldg(x + s) : 0.f;
    }
    const int c = 2 * threadIdx.x;
    float wa[10], wb[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) { wa[j] = __ldg(w0 + c * 10 + j); wb[j] = __ldg(w0 + (c + 1) * 10 + j); }
    const float mean_a = stat[((long long)b * CONV_DIM + c) * 2], rstd_a = stat[((long long)b * CONV_DIM + c) * 2 + 1];
    const float mean_b = stat[((long long)b * CONV_DIM + c + 1) * 2], rstd_b = stat[((long long)b * CONV_DIM + c + 1) * 2 + 1];
    __syncthreads();
    float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
    const uint32_t* gp = reinterpret_cast<const uint32_t*>(G0 + (long long)row_base * CONV_DIM + c);
    for (int t = 0; t < valid; ++t) {
        const float2 g = unpack_op(gp[(long long)t * (CONV_DIM / 2)]);
        float ua = 0.f, ub = 0.f;
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const float xv = xs[5 * t + j];
            ua = fmaf(wa[j], xv, ua);
            ub = fmaf(wb[j], xv, ub);
        }
        ua = (ua - mean_a) * rstd_a;
        ub = (ub - mean_b) * rstd_b;
        s1a += g.x; s2a = fmaf(g.x, ua, s2a);
        s1b += g.y; s2b = fmaf(g.y, ub, s2b);
    }
    double* o = sums + ((long long)b * CONV_DIM + c) * 2;
    atomicAdd(o + 0, (double)s1a);
    atomicAdd(o + 1, (double)s2a);
    atomicAdd(o + 2, (double)s1b);
    atomicAdd(o + 3, (double)s2b);
}

// Per-utterance constants of the dgrad:  val(t, j) = V[t, j] - c1[j] + c2[j] - sum_j' x[5t + j'] Q[j'][j]
//   c1[j] = sum_c m1_c w[c,j];  c2[j] = sum_c mean_c rstd_c m2_c w[c,j];  Q[j'][j] = sum_c w[c,j'] rstd_c m2_c w[c,j]
// consts[b] = { cc[10] = c2 - c1, Q[100] }
__global__ void __launch_bounds__(128) conv0_bwd_consts_kernel(const double* __restrict__ sums, const UttMeta* __restrict__ meta,
                                                               const float* __restrict__ w0, const float* __restrict__ stat,
                                                               float* __restrict__ consts) {
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid >= 110) return;
    const double invT = 1.0 / (double)meta[b].T0;
    double a = 0.0;
    if (tid < 10) {
        const int j = tid;
        for (int c = 0; c < CONV_DIM; ++c) {
            const double m1 = sums[((long long)b * CONV_DIM + c) * 2] * invT, m2 = sums[((long long)b * CONV_DIM + c) * 2 + 1] * invT;
            const double mean = stat[((long long)b * CONV_DIM + c) * 2], rstd = stat[((long long)b * CONV_DIM + c) * 2 + 1];
            a += (mean * rstd * m2 - m1) * (double)w0[c * 10 + j];
        }
    } else {
        const int jp = (tid - 10) / 10, j = (tid - 10) % 10;
        for (int c = 0; c < CONV_DIM; ++c) {
            const double m2 = sums[((long long)b * CONV_DIM + c) * 2 + 1] * invT;
            const double rstd = stat[((long long)b * CONV_DIM + c) * 2 + 1];
            a += (double)w0[c * 10 + jp] * rstd * m2 * (double)w0[c * 10 + j];
        }
    }
    consts[(long long)b * 112 + tid] = (float)a;
}

// Overlap-add of the 10 taps (stride 5) into the waveform gradient, times out_scale (= 1 / S).
__global__ void __launch_bounds__(256) conv0_bwd_finish_kernel(const float* __restrict__ V, const float* __restrict__ wav,
                                                               const UttMeta* __restrict__ meta, const float* __restrict__ consts,
                                                               long long n_per_utt, float out_scale, float* __restrict__ d_wav) {
    const int b = blockIdx.y;
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_per_utt) return;
    const UttMeta m = meta[b];
    __shared__ float cs[112];
    if (threadIdx.x < 110) cs[threadIdx.x] = consts[(long long)b * 112 + threadIdx.x];
    __syncthreads();
    const float* x = wav + m.wav_off;
    float acc = 0.f;
    const int t_hi = (int)(s / 5), j_lo = (int)(s % 5);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int t = t_hi - k, j = j_lo + 5 * k;
        if (t >= 0 && t < m.T0) {
            float v = V[((long long)m.row0 + t) * 16 + j] + cs[j];
#pragma unroll
            for (int jp = 0; jp < 10; ++jp) v = fmaf(-__ldg(x + 5 * t + jp), cs[10 + jp * 10 + j], v);
            acc += v;
        }
    }
    d_wav[(long long)b * n_per_utt + s] = acc * out_scale;
}

__global__ void loss_finalize_kernel(const double* __restrict__ acc, double inv_layer, double inv_head, float* __restrict__ loss) {
    double s = 0.0;
    for (int i = 0; i < LAYERS; ++i) s += acc[i] * inv_layer;
    s += acc[LAYERS] * inv_head;
    *loss = (float)s;
}

// ---------------------------------------------------------------------------------------------
struct LossBufs {
    double* acc;        // [13] L1 sums
    float* pooled;      // [2B][768] pre-ReLU mean-pooled features
    float* emb;         // [2B][256] (not used by the loss value itself; keeps pool_head's contract)
    // backward (estimate half)
    float* g_pool;      // [B][768]
    float* g_a;         // frames_e x 768 fp32 (gradient wrt a layer output / running)
    float* g_b;         // frames_e x 768 fp32 (LN backward output, residual branch)
    op_t* g_bh;         // 16-bit copy
    float* g_c;         // frames_e x 768 fp32
    op_t* g_h;          // frames_e x 3072
    op_t* g_attn;       // frames_e x 768
    op_t* g_qkv;        // frames_e x 2304
    float* D;           // frames_e x 12
    op_t* pos_g;        // grouped padded layout (reuses geometry of the forward one, estimate half)
    op_t* pos_dy;       // [pos_rows_e][768]
    op_t* g_x0h;        // frames_e x 768
    float* g_ln0;       // frames_e x 512
    op_t* gu_a;         // conv-level gradients, levels 6/4/2/0: 8 zero rows + (rows0_e + 8) x 512
    op_t* gu_b;         // levels 5/3/1
    double* c0_sums;    // [B][512][2]
    float* c0_consts;   // [B][112]
    float* c0_V;        // rows0_e x 16
    size_t bytes;
};

static size_t carve_loss(const Plan& p, int B_est, long long frames_e, long long rows0_e, long long pos_rows_e,
                         bool with_grad, void* base, LossBufs* out) {
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = (o + bytes + 1023) / 1024 * 1024;
        return base ? (void*)((char*)base + at) : nullptr;
    };
    LossBufs L;
    memset(&L, 0, sizeof(L));
    L.acc = (double*)take(sizeof(double) * 16);
    L.pooled = (float*)take(4ull * EMBED * p.B);
    L.emb = (float*)take(4ull * EMB * p.B);
    L.g_pool = (float*)take(4ull * EMBED * B_est);
    if (with_grad) {
        L.g_a = (float*)take(4ull * EMBED * frames_e);
        L.g_b = (float*)take(4ull * EMBED * frames_e);
        L.g_bh = (op_t*)take(2ull * EMBED * frames_e);
        L.g_c = (float*)take(4ull * EMBED * frames_e);
        L.g_h = (op_t*)take(2ull * FFN * frames_e);
        L.g_attn = (op_t*)take(2ull * EMBED * frames_e);
        L.g_qkv = (op_t*)take(2ull * 3 * EMBED * frames_e);
        L.D = (float*)take(4ull * HEADS * frames_e);
        L.pos_g = (op_t*)take(2ull * POS_G * POS_GC * (pos_rows_e + POS_K));
        L.pos_dy = (op_t*)take(2ull * EMBED * pos_rows_e);
        L.g_x0h = (op_t*)take(2ull * EMBED * frames_e);
        L.g_ln0 = (float*)take(4ull * CONV_DIM * frames_e);
        L.gu_a = (op_t*)take(2ull * CONV_DIM * (rows0_e + 16));
        L.gu_b = (op_t*)take(2ull * CONV_DIM * (rows0_e / 2 + 16));
        L.c0_sums = (double*)take(8ull * 2 * CONV_DIM * B_est);
        L.c0_consts = (float*)take(4ull * 112 * B_est);
        L.c0_V = (float*)take(4ull * 16 * rows0_e);
    }
    L.bytes = o;
    if (out) *out = L;
    return o;
}

static int loss_plan(int B, int64_t N, Plan* p) {
    NB_CHECK(B > 0 && N >= NOMAD_B200_MIN_SAMPLES, "loss: need B > 0 and N >= %d samples", NOMAD_B200_MIN_SAMPLES);
    std::vector<int64_t> off(2 * (size_t)B + 1);
    for (int b = 0; b <= 2 * B; ++b) off[b] = (int64_t)b * N;
    return make_plan(off.data(), 2 * B, p);
}

}  // namespace nb

using namespace nb;

extern "C" {

// core workspace of one loss step (forward in save mode + backward buffers), without the CUDA-graph staging area
static size_t loss_core_bytes(int B, int64_t N, int with_grad) {
    Plan p;
    if (loss_plan(B, N, &p)) return 0;
    const size_t fwd = carve_workspace(p, nullptr, nullptr, true);
    const long long frames_e = p.frames / 2, rows0_e = p.rows0 / 2;
    const long long pos_rows_e = frames_e + (long long)POS_GAP * B + POS_K / 2;
    return fwd + carve_loss(p, B, frames_e, rows0_e, pos_rows_e, with_grad != 0, nullptr, nullptr) + 2048;
}
static size_t loss_stage_bytes(int B, int64_t N) { return ((size_t)3 * B * N * 4 + 4096 + 1023) / 1024 * 1024; }

size_t nomad_b200_loss_workspace_bytes(int B, int64_t N, int with_grad) {
    const size_t core = loss_core_bytes(B, N, with_grad);
    if (core == 0) return 0;
    // + estimate / clean / gradient / loss staging: the step is replayed as a CUDA graph over fixed addresses
    return (core + 1023) / 1024 * 1024 + loss_stage_bytes(B, N);
}

}  // extern "C"

// phase 0: the whole step; 1: only the per-call metadata upload (the part of a step that cannot live in a CUDA graph:
// it goes through the handle's pinned staging ring); 2: the whole step except that upload (what the graph holds)
static int loss_run(nomad_b200_handle* hh, const float* est_dev, const float* clean_dev, int B, int64_t N,
                    float feature_grad_mult, float* loss_dev, float* d_est_dev, void* workspace_dev, size_t workspace_bytes,
                    cudaStream_t st, int phase) {
    Handle* h = &hh->h;
    const bool with_grad = d_est_dev != nullptr;
    const Weights& w = h->w;
    const int impl = h->gemm_impl;

    Plan p;
    NB_TRY(loss_plan(B, N, &p));
    // utterances 0..B-1 = estimates, B..2B-1 = clean; both addressed relative to est_dev
    const long long clean_off = (long long)(((intptr_t)clean_dev - (intptr_t)est_dev) / 4);
    for (int b = 0; b < B; ++b) {
        p.utt[b].wav_off = (long long)b * N;
        p.utt[B + b].wav_off = clean_off + (long long)b * N;
    }
    const long long F = p.frames, Fe = F / 2, R0e = p.rows0 / 2;
    const long long pos_rows_e = Fe + (long long)POS_GAP * B + POS_K / 2;
    const int T = p.max_T;
    Workspace ws;
    const size_t fwd_bytes = carve_workspace(p, workspace_dev, &ws, true);
    LossBufs L;
    const size_t loss_bytes = carve_loss(p, B, Fe, R0e, pos_rows_e, with_grad, (char*)workspace_dev + fwd_bytes, &L);
    NB_CHECK(workspace_bytes >= fwd_bytes + loss_bytes, "loss: workspace too small (%zu < %zu bytes)", workspace_bytes,
             fwd_bytes + loss_bytes);
    NB_CHECK(((uintptr_t)workspace_dev & 1023) == 0, "loss: workspace must be 1024-byte aligned");

    // ------------------------------------------------------------------ forward (both halves, save mode)
    if (phase != 2) NB_TRY(upload_meta(h, p, ws, st));
    if (phase == 1) return 0;
    NB_CUDA(cudaMemsetAsync(L.acc, 0, sizeof(double) * 16, st));
    NB_TRY(forward_encoder(h, p, ws, est_dev, st, nullptr, 0));
    NB_TRY(launch_pool_head(st, ws.x, ws.meta, p.B, w.loss_head_wt, w.loss_head_b, L.emb, L.pooled));
    const unsigned row_blocks = (unsigned)((Fe + 7) / 8);
    for (int l = 0; l < LAYERS; ++l) {
        l1_layer_kernel<<<row_blocks, 256, 0, st>>>(ws.layer[l].pre2, ws.meta, B, Fe, w.layer[l].ln2_g, L.acc + l);
        NB_LAUNCHED();
    }
    // gradient scale: power of two with S / (B T 768) in [1, 2)
    const double numel = (double)B * T * EMBED;
    const float S = with_grad ? (float)std::exp2(std::ceil(std::log2(numel))) : 0.f;
    const float seed_layer = (float)(S / numel), seed_head = (float)(S / ((double)B * EMB));
    // head term (13th): value always, pooled-feature gradient when a backward follows (seed 0 otherwise)
    head_bwd_kernel<<<B, 256, 0, st>>>(L.pooled, ws.meta, B, w.loss_head_wt, w.loss_head_w, w.loss_head_b, seed_head,
                                       L.acc + LAYERS, L.g_pool);
    NB_LAUNCHED();
    loss_finalize_kernel<<<1, 1, 0, st>>>(L.acc, 1.0 / numel, 1.0 / ((double)B * EMB), loss_dev);
    NB_LAUNCHED();
    if (getenv("NOMAD_B200_DEBUG_LOSS")) {  // per-term L1 sums, for debugging only (synchronises)
        double hacc[16];
        NB_CUDA(cudaStreamSynchronize(st));
        NB_CUDA(cudaMemcpy(hacc, L.acc, sizeof(hacc), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 13; ++i) fprintf(stderr, "[nomad_b200] L1 term %2d: sum |diff| = %.9g\n", i, hacc[i]);
    }
    if (!with_grad) return 0;

    // ------------------------------------------------------------------ backward (estimate half)
    auto epi_grad = [&](int flags, const float* resid, float* out_f, op_t* out_h, const op_t* aux, long long ld) {
        GemmEpilogue e = epi_linear(flags, nullptr, resid, out_f, out_h, ld);
        e.aux = aux;
        return e;
    };
    NB_CUDA(cudaMemsetAsync(L.g_qkv, 0, 2ull * 3 * EMBED * Fe, st));
    const float* g_running = nullptr;  // gradient wrt the current layer's output from the layers above
    for (int l = LAYERS - 1; l >= 0; --l) {
        const LayerWeights& W = w.layer[l];
        const LayerBufs& Lb = ws.layer[l];
        // final_layer_norm backward (+ L1 seed of this layer's output, + pooled-head gradient on the top layer)
        ln768_bwd_kernel<<<row_blocks, 256, 0, st>>>(g_running, Lb.pre2, ws.meta, B, Fe, W.ln2_g, seed_layer,
                                                     l == LAYERS - 1 ? L.g_pool : nullptr, L.g_b, L.g_bh);
        NB_LAUNCHED();
        {   // fc2 dgrad, times gelu'(fc1 pre-activation)
            GemmOperand A{L.g_bh, Fe, EMBED, 0, 0};
            GemmOperand Bw{W.wt_fc2, FFN, EMBED, 0, 0};
            GemmEpilogue e = epi_grad(EPI_MUL_AUX | EPI_OUT_H16, nullptr, nullptr, L.g_h, Lb.ffn_aux, FFN);
            NB_TRY(gemm_h16(st, A, Bw, (int)Fe, FFN, EMBED, 1, e, impl));
        }
        {   // fc1 dgrad + residual branch
            GemmOperand A{L.g_h, Fe, FFN, 0, 0};
            GemmOperand Bw{W.wt_fc1, EMBED, FFN, 0, 0};
            GemmEpilogue e = epi_grad(EPI_RESID | EPI_OUT_F32, L.g_b, L.g_c, nullptr, nullptr, EMBED);
            NB_TRY(gemm_h16(st, A, Bw, (int)Fe, EMBED, FFN, 1, e, impl));
        }
        // self_attn_layer_norm backward
        ln768_bwd_kernel<<<row_blocks, 256, 0, st>>>(L.g_c, Lb.pre1, ws.meta, B, Fe, W.ln1_g, 0.f, nullptr, L.g_b, L.g_bh);
        NB_LAUNCHED();
        {   // out_proj dgrad
            GemmOperand A{L.g_bh, Fe, EMBED, 0, 0};
            GemmOperand Bw{W.wt_o, EMBED, EMBED, 0, 0};
            GemmEpilogue e = epi_grad(EPI_OUT_H16, nullptr, nullptr, L.g_attn, nullptr, EMBED);
            NB_TRY(gemm_h16(st, A, Bw, (int)Fe, EMBED, EMBED, 1, e, impl));
        }
        NB_TRY(launch_attention_bwd(st, Lb.qkv, Lb.attn, L.g_attn, Lb.lse, L.D, ws.meta, B, T, Fe, L.g_qkv));
        {   // fused q/k/v dgrad + residual branch -> gradient wrt this layer's input
            GemmOperand A{L.g_qkv, Fe, 3 * EMBED, 0, 0};
            GemmOperand Bw{W.wt_qkv, EMBED, 3 * EMBED, 0, 0};
            GemmEpilogue e = epi_grad(EPI_RESID | EPI_OUT_F32, L.g_b, L.g_a, nullptr, nullptr, EMBED);
            NB_TRY(gemm_h16(st, A, Bw, (int)Fe, EMBED, 3 * EMBED, 1, e, impl));
        }
        g_running = L.g_a;
    }
    // encoder LayerNorm + positional conv backward
    {
        const long long rows_alloc = pos_rows_e + POS_K;
        NB_CUDA(cudaMemsetAsync(L.pos_g, 0, 2ull * POS_G * POS_GC * rows_alloc, st));
        pos_bwd_prep_kernel<<<row_blocks, 256, 0, st>>>(L.g_a, ws.x0, ws.pos_y, ws.pos_aux, ws.meta, B, Fe, w.lne_g,
                                                        rows_alloc, L.g_b, L.pos_g);
        NB_LAUNCHED();
        if (impl == 0) {
            NB_TRY(launch_posconv(st, L.pos_g, rows_alloc, pos_rows_e, w.pos_wt, nullptr, 0, L.pos_dy, nullptr));
        } else {
            GemmOperand A{L.pos_g, pos_rows_e, POS_GC, rows_alloc * POS_GC, 0};
            GemmOperand Bw{w.pos_wt, POS_GC, (long long)POS_K * POS_GC, (long long)POS_GC * POS_K * POS_GC, 0};
            GemmEpilogue e = epi_grad(EPI_OUT_H16, nullptr, nullptr, L.pos_dy, nullptr, EMBED);
            e.out_bstride = POS_GC;
            NB_TRY(gemm_h16(st, A, Bw, (int)pos_rows_e, POS_GC, POS_K * POS_GC, POS_G, e, impl));
        }
        pos_bwd_finish_kernel<<<row_blocks, 256, 0, st>>>(L.g_b, L.pos_dy, ws.meta, B, Fe, L.g_x0h);
        NB_LAUNCHED();
    }
    {   // feature projection dgrad
        GemmOperand A{L.g_x0h, Fe, EMBED, 0, 0};
        GemmOperand Bw{w.proj_wt, CONV_DIM, EMBED, 0, 0};
        GemmEpilogue e = epi_grad(EPI_OUT_F32, nullptr, L.g_ln0, nullptr, nullptr, CONV_DIM);
        NB_TRY(gemm_h16(st, A, Bw, (int)Fe, CONV_DIM, EMBED, 1, e, impl));
    }
    // conv stack backward.  Gradient buffers start 8 rows into their allocation so that "row -1" reads zeros.
    NB_CUDA(cudaMemsetAsync(L.gu_a, 0, 2ull * CONV_DIM * 8, st));
    NB_CUDA(cudaMemsetAsync(L.gu_b, 0, 2ull * CONV_DIM * 8, st));
    op_t* gu[7];
    for (int l = 0; l < 7; ++l) gu[l] = ((l & 1) ? L.gu_b : L.gu_a) + 8 * CONV_DIM;
    ln512_bwd_kernel<<<row_blocks, 256, 0, st>>>(L.g_ln0, ws.y[6], ws.aux[6], Fe, w.ln0_g, feature_grad_mult, gu[6]);
    NB_LAUNCHED();
    for (int l = 6; l >= 1; --l) {
        const long long M = R0e >> l;  // rows at level l (estimate half)
        if (CONV_KERNEL[l] == 2) {
            // rows 2m and 2m+1 of level l-1 in one GEMM: N = 1024 = (tap, cin)
            GemmOperand A{gu[l], M, CONV_DIM, 0, 0};
            GemmOperand Bw{w.conv_wt[l], 2 * CONV_DIM, CONV_DIM, 0, 0};
            GemmEpilogue e = epi_grad(EPI_MUL_AUX | EPI_OUT_H16, nullptr, nullptr, gu[l - 1], ws.aux[l - 1], 2 * CONV_DIM);
            NB_TRY(gemm_h16(st, A, Bw, (int)M, 2 * CONV_DIM, CONV_DIM, 1, e, impl));
        } else {
            {   // even rows 2m: taps 2 (from row m-1) and 0 (row m): overlapping rows starting one row early
                GemmOperand A{gu[l] - CONV_DIM, M, CONV_DIM, 0, 0};
                GemmOperand Bw{w.conv_wte[l], CONV_DIM, 2 * CONV_DIM, 0, 0};
                GemmEpilogue e = epi_grad(EPI_MUL_AUX | EPI_OUT_H16, nullptr, nullptr, gu[l - 1], ws.aux[l - 1], 2 * CONV_DIM);
                NB_TRY(gemm_h16(st, A, Bw, (int)M, CONV_DIM, 2 * CONV_DIM, 1, e, impl));
            }
            {   // odd rows 2m+1: tap 1
                GemmOperand A{gu[l], M, CONV_DIM, 0, 0};
                GemmOperand Bw{w.conv_wt[l] + (size_t)CONV_DIM * CONV_DIM, CONV_DIM, CONV_DIM, 0, 0};
                GemmEpilogue e = epi_grad(EPI_MUL_AUX | EPI_OUT_H16, nullptr, nullptr, gu[l - 1] + CONV_DIM,
                                          ws.aux[l - 1] + CONV_DIM, 2 * CONV_DIM);
                NB_TRY(gemm_h16(st, A, Bw, (int)M, CONV_DIM, CONV_DIM, 1, e, impl));
            }
        }
    }
    // conv0 + GroupNorm backward: gu[0] = G0
    NB_CUDA(cudaMemsetAsync(L.c0_sums, 0, 8ull * 2 * CONV_DIM * B, st));
    conv0_bwd_stats_kernel<<<(unsigned)(R0e / 64), 256, 0, st>>>(gu[0], est_dev, ws.meta, B, w.conv0_w, ws.gn_stat, L.c0_sums);
    NB_LAUNCHED();
    conv0_bwd_consts_kernel<<<B, 128, 0, st>>>(L.c0_sums, ws.meta, w.conv0_w, ws.gn_stat, L.c0_consts);
    NB_LAUNCHED();
    {
        GemmOperand A{gu[0], R0e, CONV_DIM, 0, 0};
        GemmOperand Bw{w.conv0_wh, 16, CONV_DIM, 0, 0};
        GemmEpilogue e = epi_grad(EPI_OUT_F32, nullptr, L.c0_V, nullptr, nullptr, 16);
        NB_TRY(gemm_h16(st, A, Bw, (int)R0e, 16, CONV_DIM, 1, e, impl));
    }
    dim3 fgrid((unsigned)((N + 255) / 256), B);
    conv0_bwd_finish_kernel<<<fgrid, 256, 0, st>>>(L.c0_V, est_dev, ws.meta, L.c0_consts, N, 1.0f / S, d_est_dev);
    NB_LAUNCHED();
    return 0;
}

extern "C" {

// The step is ~250 small launches for B = 32 x 2 s (profiles/r02_loss_trace_before.log: 8 % of the step was launch gaps):
// after one eager call per (B, N, grad, feature_grad_mult, workspace) the launch sequence is captured into a CUDA graph
// over fixed addresses -- inputs are copied into a staging area at the end of the caller's workspace, results copied
// out -- and replayed.  NOMAD_B200_LOSS_GRAPH=0 keeps every call eager.
int nomad_b200_loss_fwd_bwd(nomad_b200_handle* hh, const float* est_dev, const float* clean_dev, int B, int64_t N,
                            float feature_grad_mult, float* loss_dev, float* d_est_dev, void* workspace_dev,
                            size_t workspace_bytes, void* stream) {
    NB_CHECK(hh != nullptr, "null nomad_b200 handle");
    Handle* h = &hh->h;
    NB_CHECK(est_dev && clean_dev && loss_dev && workspace_dev, "loss: null pointer");
    NB_CHECK(h->has_loss_head, "loss: call nomad_b200_set_loss_head first (LossNetLayers has its own head, nomad.py:238-241)");
    NB_CHECK((((uintptr_t)est_dev | (uintptr_t)clean_dev) & 3) == 0, "loss: waveform pointers must be 4-byte aligned");
    NB_CHECK(((uintptr_t)workspace_dev & 1023) == 0, "loss: workspace must be 1024-byte aligned");
    NB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int with_grad = d_est_dev != nullptr ? 1 : 0;
    static const int use_graph = getenv("NOMAD_B200_LOSS_GRAPH") ? atoi(getenv("NOMAD_B200_LOSS_GRAPH")) : 1;
    const size_t core = (loss_core_bytes(B, N, with_grad) + 1023) / 1024 * 1024;
    NB_CHECK(core != 0, "loss: %s", nomad_b200_last_error());
    const bool graph_ok = use_graph && h->gemm_impl == 0 && !getenv("NOMAD_B200_DEBUG_LOSS") && !gemm_profile_active() &&
                          workspace_bytes >= core + loss_stage_bytes(B, N);
    if (!graph_ok)
        return loss_run(hh, est_dev, clean_dev, B, N, feature_grad_mult, loss_dev, d_est_dev, workspace_dev, workspace_bytes, st, 0);

    const size_t bn = (size_t)B * N;
    float* s_est = (float*)((char*)workspace_dev + core);
    float* s_clean = s_est + bn;
    float* s_grad = s_clean + bn;
    float* s_loss = s_grad + bn;
    NB_CUDA(cudaMemcpyAsync(s_est, est_dev, bn * 4, cudaMemcpyDeviceToDevice, st));
    NB_CUDA(cudaMemcpyAsync(s_clean, clean_dev, bn * 4, cudaMemcpyDeviceToDevice, st));
    LossGraphEntry* ent = nullptr;
    for (auto& e : h->loss_graphs)
        if (e.B == B && e.N == N && e.with_grad == with_grad && e.fgm == feature_grad_mult && e.ws == workspace_dev) ent = &e;
    if (ent == nullptr) {  // first call with this shape: eager (also performs the one-off per-kernel attribute set-up)
        if (h->loss_graphs.size() >= 8) {
            for (auto& e : h->loss_graphs)
                if (e.exec) cudaGraphExecDestroy(e.exec);
            h->loss_graphs.clear();
        }
        h->loss_graphs.push_back(LossGraphEntry{B, (long long)N, with_grad, feature_grad_mult, workspace_dev, nullptr, 0, false});
        NB_TRY(loss_run(hh, s_est, s_clean, B, N, feature_grad_mult, s_loss, with_grad ? s_grad : nullptr, workspace_dev, core, st, 0));
    } else if (ent->exec == nullptr && !ent->failed) {  // second call: capture on the handle's own stream, then launch
        if (!h->graph_stream) NB_CUDA(cudaStreamCreateWithFlags(&h->graph_stream, cudaStreamNonBlocking));
        NB_TRY(loss_run(hh, s_est, s_clean, B, N, feature_grad_mult, s_loss, with_grad ? s_grad : nullptr, workspace_dev, core, st, 1));
        const long long n0 = nomad_b200_launch_count();
        cudaGraph_t graph = nullptr;
        NB_CUDA(cudaStreamBeginCapture(h->graph_stream, cudaStreamCaptureModeThreadLocal));
        const int rc = loss_run(hh, s_est, s_clean, B, N, feature_grad_mult, s_loss, with_grad ? s_grad : nullptr, workspace_dev,
                                core, h->graph_stream, 2);
        const cudaError_t ce = cudaStreamEndCapture(h->graph_stream, &graph);
        ent->kernels = nomad_b200_launch_count() - n0;
        if (rc == 0 && ce == cudaSuccess && graph != nullptr &&
            cudaGraphInstantiate(&ent->exec, graph, 0) == cudaSuccess) {
            cudaGraphDestroy(graph);
            NB_CUDA(cudaGraphLaunch(ent->exec, st));
        } else {  // capture not possible here: stay eager for this shape
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            ent->exec = nullptr;
            ent->failed = true;
            count_launches(-ent->kernels);
            NB_TRY(loss_run(hh, s_est, s_clean, B, N, feature_grad_mult, s_loss, with_grad ? s_grad : nullptr, workspace_dev, core, st, 2));
        }
    } else if (ent->exec != nullptr) {
        NB_TRY(loss_run(hh, s_est, s_clean, B, N, feature_grad_mult, s_loss, with_grad ? s_grad : nullptr, workspace_dev, core, st, 1));
        NB_CUDA(cudaGraphLaunch(ent->exec, st));
        count_launches(ent->kernels);
    } else {
        NB_TRY(loss_run(hh, s_est, s_clean, B, N, feature_grad_mult, s_loss, with_grad ? s_grad : nullptr, workspace_dev, core, st, 0));
    }
    NB_CUDA(cudaMemcpyAsync(loss_dev, s_loss, 4, cudaMemcpyDeviceToDevice, st));
    if (with_grad) NB_CUDA(cudaMemcpyAsync(d_est_dev, s_grad, bn * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // extern "C"
