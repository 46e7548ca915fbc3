// Backward of the attention core for the NOMAD loss path (dgrad only).
//   P = exp(S - lse),  S = Q K^T  (Q pre-scaled by head_dim^-0.5 via the fused QKV weights)
//   D[q]  = sum_d dO[q, d] O[q, d]
//   dS    = P * (dO V^T - D)
//   dQ = dS K,   dK = dS^T Q,   dV = P^T dO
// Two atomic-free, deterministic kernels: one CTA per 64-query tile produces dQ (streams over key tiles),
// one CTA per 64-key tile produces dK and dV (streams over query tiles).  S is recomputed in both.
// fp16 mma.sync m16n8k16 with fp32 accumulation, same tile/fragment conventions as the forward kernel.
#include "kernels.cuh"

namespace nb {

static constexpr int AB = 64;  // tile edge (queries or keys)

__device__ __forceinline__ void b_cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void b_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void b_ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void b_ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void b_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ op_t* b_tile_ptr(op_t* base, int row, int col) {
    return base + row * 64 + ((((col >> 3) ^ (row & 7)) << 3) | (col & 7));
}
// 64 rows x 64 cols of a row-major matrix with leading dimension ld (elements) -> swizzled smem tile
__device__ __forceinline__ void b_load_tile(op_t* tile, const op_t* __restrict__ src, long long ld, long long first_row,
                                            int row0, int last_valid, int col0, int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 128;
        const int r = idx >> 3, ch = idx & 7;
        int gr = row0 + r;
        gr = gr > last_valid ? last_valid : gr;
        b_cp_async16(b_tile_ptr(tile, r, ch * 8), src + (first_row + gr) * ld + col0 + ch * 8);
    }
}
// A-operand fragments (16 rows x 64 k) of rows [r0, r0+16) of a tile
__device__ __forceinline__ void b_load_a_frags(uint32_t (&f)[4][4], op_t* tile, int r0, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
        b_ldsm_x4(f[kk], b_tile_ptr(tile, r0 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 16 + (lane >> 4) * 8));
}
// acc[16 x 64] (+)= A[16 x 64 (k)] * T^T where T is a [64 (n)][64 (k)] tile ("n-major, k contiguous")
__device__ __forceinline__ void b_mma_nt(float (&acc)[8][4], const uint32_t (&a)[4][4], op_t* tile, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t bf[4];
            b_ldsm_x4(bf, b_tile_ptr(tile, np * 16 + (lane & 7) + (lane >> 4) * 8, kk * 16 + ((lane >> 3) & 1) * 8));
            b_mma(acc[2 * np], a[kk], bf[0], bf[1]);
            b_mma(acc[2 * np + 1], a[kk], bf[2], bf[3]);
        }
}
// acc[16 x 64 (n)] += A[16 x 64 (k)] * T where T is a [64 (k)][64 (n)] tile ("k rows, n contiguous")
__device__ __forceinline__ void b_mma_nn(float (&acc)[8][4], const uint32_t (&a)[4][4], op_t* tile, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t bf[4];
            b_ldsm_x4_t(bf, b_tile_ptr(tile, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 16 + (lane >> 4) * 8));
            b_mma(acc[2 * np], a[kk], bf[0], bf[1]);
            b_mma(acc[2 * np + 1], a[kk], bf[2], bf[3]);
        }
}
// accumulator layout (16 x 64) -> A-operand fragments of the same 16 x 64 matrix
__device__ __forceinline__ void b_acc_to_a(uint32_t (&a)[4][4], const float (&c)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i >> 1][(i & 1) * 2 + 0] = pack_op(c[i][0], c[i][1]);
        a[i >> 1][(i & 1) * 2 + 1] = pack_op(c[i][2], c[i][3]);
    }
}
__device__ __forceinline__ void b_zero(float (&c)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// D[f, h] = sum_d dO[f, h*64 + d] * O[f, h*64 + d]; one warp per (frame, head)
__global__ void __launch_bounds__(256) attn_bwd_d_kernel(const op_t* __restrict__ d_out, const op_t* __restrict__ o,
                                                         long long frames, float* __restrict__ D) {
    const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= frames * HEADS) return;
    const int lane = threadIdx.x & 31;
    const long long f = w / HEADS;
    const int h = (int)(w % HEADS);
    const uint32_t a = *reinterpret_cast<const uint32_t*>(d_out + f * EMBED + h * HEAD_DIM + 2 * lane);
    const uint32_t b = *reinterpret_cast<const uint32_t*>(o + f * EMBED + h * HEAD_DIM + 2 * lane);
    const float2 x = unpack_op(a), y = unpack_op(b);
    const float s = warp_sum(x.x * y.x + x.y * y.y);
    if (lane == 0) D[w] = s;
}

// ---------------------------------------------------------------------------------------------
// D consistent with the probabilities the backward itself uses: D[q] = sum_j P_qj dP_qj / sum_j P_qj with
// P = exp(S - lse) recomputed here.  The cheap form above takes O from the forward pass, where it was rounded to
// 16 bits: the row sums of dS = P (dP - D) are then off by ~5e-4 |dO||O| instead of zero, a first-order error on
// dQ / dK that swamps them whenever the true gradient wrt the scores is small (nearly uniform attention, or the
// per-utterance broadcast gradient of a mean-pooled objective: the triplet fine-tuning step).  One extra S / dP
// pass per query tile; used where parameter gradients are wanted.
__global__ void __launch_bounds__(128) attn_bwd_dcons_kernel(const op_t* __restrict__ qkv, const op_t* __restrict__ d_out,
                                                             const float* __restrict__ lse, const UttMeta* __restrict__ meta,
                                                             float* __restrict__ D) {
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AB;
    const int T = meta[b].T;
    if (q0 >= T) return;
    const long long f0 = meta[b].frame0;
    __shared__ __align__(128) op_t Qs[AB * 64];
    __shared__ __align__(128) op_t dOs[AB * 64];
    __shared__ __align__(128) op_t Ks[AB * 64];
    __shared__ __align__(128) op_t Vs[AB * 64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    b_load_tile(Qs, qkv, 3 * EMBED, f0, q0, T - 1, h * HEAD_DIM, tid);
    b_load_tile(dOs, d_out, EMBED, f0, q0, T - 1, h * HEAD_DIM, tid);
    b_cp_async_wait_all();
    __syncthreads();
    uint32_t qf[4][4], dof[4][4];
    b_load_a_frags(qf, Qs, warp * 16, lane);
    b_load_a_frags(dof, dOs, warp * 16, lane);
    const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
    const int c0 = r0 < T ? r0 : T - 1, c1 = r1 < T ? r1 : T - 1;
    const float lse0 = lse[(f0 + c0) * HEADS + h], lse1 = lse[(f0 + c1) * HEADS + h];
    float n0 = 0.f, n1 = 0.f, z0 = 0.f, z1 = 0.f;
    const int n_tiles = (T + AB - 1) / AB;
    for (int kt = 0; kt < n_tiles; ++kt) {
        __syncthreads();
        b_load_tile(Ks, qkv, 3 * EMBED, f0, kt * AB, T - 1, EMBED + h * HEAD_DIM, tid);
        b_load_tile(Vs, qkv, 3 * EMBED, f0, kt * AB, T - 1, 2 * EMBED + h * HEAD_DIM, tid);
        b_cp_async_wait_all();
        __syncthreads();
        float s[8][4], dp[8][4];
        b_zero(s);
        b_zero(dp);
        b_mma_nt(s, qf, Ks, lane);
        b_mma_nt(dp, dof, Vs, lane);
        const int key_base = kt * AB + 2 * (lane & 3);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k0 = key_base + i * 8;
            const bool v0 = k0 < T, v1 = k0 + 1 < T;
            const float p0 = v0 ? __expf(s[i][0] - lse0) : 0.f, p1 = v1 ? __expf(s[i][1] - lse0) : 0.f;
            const float p2 = v0 ? __expf(s[i][2] - lse1) : 0.f, p3 = v1 ? __expf(s[i][3] - lse1) : 0.f;
            n0 = fmaf(p0, dp[i][0], fmaf(p1, dp[i][1], n0));
            n1 = fmaf(p2, dp[i][2], fmaf(p3, dp[i][3], n1));
            z0 += p0 + p1;
            z1 += p2 + p3;
        }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        n0 += __shfl_xor_sync(0xffffffffu, n0, o); n1 += __shfl_xor_sync(0xffffffffu, n1, o);
        z0 += __shfl_xor_sync(0xffffffffu, z0, o); z1 += __shfl_xor_sync(0xffffffffu, z1, o);
    }
    if ((lane & 3) == 0) {
        if (r0 < T) D[(f0 + r0) * HEADS + h] = n0 / z0;
        if (r1 < T) D[(f0 + r1) * HEADS + h] = n1 / z1;
    }
}

// ---------------------------------------------------------------------------------------------
// dQ: one CTA per (64-query tile, head, utterance)
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const op_t* __restrict__ qkv, const op_t* __restrict__ d_out,
                                                          const float* __restrict__ lse, const float* __restrict__ D,
                                                          const UttMeta* __restrict__ meta, op_t* __restrict__ d_qkv) {
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AB;
    const int T = meta[b].T;
    if (q0 >= T) return;
    const long long f0 = meta[b].frame0;
    __shared__ __align__(128) op_t Qs[AB * 64];
    __shared__ __align__(128) op_t dOs[AB * 64];
    __shared__ __align__(128) op_t Ks[AB * 64];
    __shared__ __align__(128) op_t Vs[AB * 64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    b_load_tile(Qs, qkv, 3 * EMBED, f0, q0, T - 1, h * HEAD_DIM, tid);
    b_load_tile(dOs, d_out, EMBED, f0, q0, T - 1, h * HEAD_DIM, tid);
    b_cp_async_wait_all();
    __syncthreads();
    uint32_t qf[4][4], dof[4][4];
    b_load_a_frags(qf, Qs, warp * 16, lane);
    b_load_a_frags(dof, dOs, warp * 16, lane);
    const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
    const int c0 = r0 < T ? r0 : T - 1, c1 = r1 < T ? r1 : T - 1;
    const float lse0 = lse[(f0 + c0) * HEADS + h], lse1 = lse[(f0 + c1) * HEADS + h];
    const float D0 = D[(f0 + c0) * HEADS + h], D1 = D[(f0 + c1) * HEADS + h];
    float dq[8][4];
    b_zero(dq);
    const int n_tiles = (T + AB - 1) / AB;
    for (int kt = 0; kt < n_tiles; ++kt) {
        __syncthreads();
        b_load_tile(Ks, qkv, 3 * EMBED, f0, kt * AB, T - 1, EMBED + h * HEAD_DIM, tid);
        b_load_tile(Vs, qkv, 3 * EMBED, f0, kt * AB, T - 1, 2 * EMBED + h * HEAD_DIM, tid);
        b_cp_async_wait_all();
        __syncthreads();
        float s[8][4], dp[8][4];
        b_zero(s);
        b_zero(dp);
        b_mma_nt(s, qf, Ks, lane);    // S  = Q K^T
        b_mma_nt(dp, dof, Vs, lane);  // dP = dO V^T
        const int key_base = kt * AB + 2 * (lane & 3);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k0 = key_base + i * 8;
            const bool v0 = k0 < T, v1 = k0 + 1 < T;
            const float p0 = v0 ? __expf(s[i][0] - lse0) : 0.f, p1 = v1 ? __expf(s[i][1] - lse0) : 0.f;
            const float p2 = v0 ? __expf(s[i][2] - lse1) : 0.f, p3 = v1 ? __expf(s[i][3] - lse1) : 0.f;
            s[i][0] = p0 * (dp[i][0] - D0); s[i][1] = p1 * (dp[i][1] - D0);
            s[i][2] = p2 * (dp[i][2] - D1); s[i][3] = p3 * (dp[i][3] - D1);
        }
        uint32_t dsf[4][4];
        b_acc_to_a(dsf, s);
        b_mma_nn(dq, dsf, Ks, lane);  // dQ += dS K
    }
    op_t* ob = d_qkv + f0 * (3 * EMBED) + h * HEAD_DIM + 2 * (lane & 3);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (r0 < T) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * (3 * EMBED) + i * 8) = pack_op(dq[i][0], dq[i][1]);
        if (r1 < T) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * (3 * EMBED) + i * 8) = pack_op(dq[i][2], dq[i][3]);
    }
}

// ---------------------------------------------------------------------------------------------
// dK, dV: one CTA per (64-key tile, head, utterance); works on S^T so keys are the M dimension
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const op_t* __restrict__ qkv, const op_t* __restrict__ d_out,
                                                           const float* __restrict__ lse, const float* __restrict__ D,
                                                           const UttMeta* __restrict__ meta, op_t* __restrict__ d_qkv) {
    const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * AB;
    const int T = meta[b].T;
    if (k0 >= T) return;
    const long long f0 = meta[b].frame0;
    __shared__ __align__(128) op_t Ks[AB * 64];
    __shared__ __align__(128) op_t Vs[AB * 64];
    __shared__ __align__(128) op_t Qs[AB * 64];
    __shared__ __align__(128) op_t dOs[AB * 64];
    __shared__ float lse_s[AB], D_s[AB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    b_load_tile(Ks, qkv, 3 * EMBED, f0, k0, T - 1, EMBED + h * HEAD_DIM, tid);
    b_load_tile(Vs, qkv, 3 * EMBED, f0, k0, T - 1, 2 * EMBED + h * HEAD_DIM, tid);
    b_cp_async_wait_all();
    __syncthreads();
    uint32_t kf[4][4], vf[4][4];
    b_load_a_frags(kf, Ks, warp * 16, lane);
    b_load_a_frags(vf, Vs, warp * 16, lane);
    const int kr0 = k0 + warp * 16 + (lane >> 2), kr1 = kr0 + 8;  // key rows owned by this thread
    float dk[8][4], dv[8][4];
    b_zero(dk);
    b_zero(dv);
    const int n_tiles = (T + AB - 1) / AB;
    for (int qt = 0; qt < n_tiles; ++qt) {
        __syncthreads();
        b_load_tile(Qs, qkv, 3 * EMBED, f0, qt * AB, T - 1, h * HEAD_DIM, tid);
        b_load_tile(dOs, d_out, EMBED, f0, qt * AB, T - 1, h * HEAD_DIM, tid);
        if (tid < AB) {
            const int q = qt * AB + tid;
            const int qc = q < T ? q : T - 1;
            lse_s[tid] = lse[(f0 + qc) * HEADS + h];
            D_s[tid] = D[(f0 + qc) * HEADS + h];
        }
        b_cp_async_wait_all();
        __syncthreads();
        float st[8][4], dpt[8][4];
        b_zero(st);
        b_zero(dpt);
        b_mma_nt(st, kf, Qs, lane);    // S^T  = K Q^T
        b_mma_nt(dpt, vf, dOs, lane);  // dP^T = V dO^T
        const int q_base = qt * AB + 2 * (lane & 3);
        float pt[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int ql = i * 8 + 2 * (lane & 3);  // local query column
            const int q = q_base + i * 8;
            const bool qv0 = q < T, qv1 = q + 1 < T;
            const float l0 = lse_s[ql], l1 = lse_s[ql + 1], d0 = D_s[ql], d1 = D_s[ql + 1];
            const float p0 = (qv0 && kr0 < T) ? __expf(st[i][0] - l0) : 0.f;
            const float p1 = (qv1 && kr0 < T) ? __expf(st[i][1] - l1) : 0.f;
            const float p2 = (qv0 && kr1 < T) ? __expf(st[i][2] - l0) : 0.f;
            const float p3 = (qv1 && kr1 < T) ? __expf(st[i][3] - l1) : 0.f;
            pt[i][0] = p0; pt[i][1] = p1; pt[i][2] = p2; pt[i][3] = p3;
            st[i][0] = p0 * (dpt[i][0] - d0); st[i][1] = p1 * (dpt[i][1] - d1);
            st[i][2] = p2 * (dpt[i][2] - d0); st[i][3] = p3 * (dpt[i][3] - d1);
        }
        uint32_t af[4][4];
        b_acc_to_a(af, pt);
        b_mma_nn(dv, af, dOs, lane);  // dV += P^T dO
        b_acc_to_a(af, st);
        b_mma_nn(dk, af, Qs, lane);   // dK += dS^T Q
    }
    op_t* ob = d_qkv + f0 * (3 * EMBED) + h * HEAD_DIM + 2 * (lane & 3);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (kr0 < T) {
            *reinterpret_cast<uint32_t*>(ob + (long long)kr0 * (3 * EMBED) + EMBED + i * 8) = pack_op(dk[i][0], dk[i][1]);
            *reinterpret_cast<uint32_t*>(ob + (long long)kr0 * (3 * EMBED) + 2 * EMBED + i * 8) = pack_op(dv[i][0], dv[i][1]);
        }
        if (kr1 < T) {
            *reinterpret_cast<uint32_t*>(ob + (long long)kr1 * (3 * EMBED) + EMBED + i * 8) = pack_op(dk[i][2], dk[i][3]);
            *reinterpret_cast<uint32_t*>(ob + (long long)kr1 * (3 * EMBED) + 2 * EMBED + i * 8) = pack_op(dv[i][2], dv[i][3]);
        }
    }
}

int launch_attention_bwd(cudaStream_t st, const op_t* qkv, const op_t* attn_out, const op_t* d_out, const float* lse,
                         float* D, const UttMeta* meta, int B, int max_T, long long frames, op_t* d_qkv, bool consistent_d) {
    dim3 grid((max_T + AB - 1) / AB, HEADS, B);
    if (consistent_d) {
        attn_bwd_dcons_kernel<<<grid, 128, 0, st>>>(qkv, d_out, lse, meta, D);
    } else {
        attn_bwd_d_kernel<<<(unsigned)((frames * HEADS + 7) / 8), 256, 0, st>>>(d_out, attn_out, frames, D);
    }
    NB_LAUNCHED();
    attn_bwd_dq_kernel<<<grid, 128, 0, st>>>(qkv, d_out, lse, D, meta, d_qkv);
    NB_LAUNCHED();
    attn_bwd_dkv_kernel<<<grid, 128, 0, st>>>(qkv, d_out, lse, D, meta, d_qkv);
    NB_LAUNCHED();
    return 0;
}

}  // namespace nb
