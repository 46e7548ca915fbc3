// Pairwise Euclidean distance + row mean:  scipy.spatial.distance.cdist(test, nmr) followed by
// np.mean(axis=1) (reference nomad.py:108,111).
//
// fp32 direct-difference kernel: d = sqrt(sum_k (a_k - b_k)^2).  The direct form has no cancellation
// (unit-norm embeddings give distances down to ~0.1 where the Gram form ||a||^2 + ||b||^2 - 2ab loses
// 3-4 digits), so it meets the 1e-5 bound against scipy's float64 result without tricks.  Row sums are
// accumulated in fp64.
#include <cstring>

#include "kernels.cuh"

namespace nb {

static constexpr int CD_TILE = 64, CD_K = 32, CD_DIM = 256;

__global__ void __launch_bounds__(256) cdist_fp32_kernel(const float* __restrict__ a, long long n,
                                                         const float* __restrict__ b, long long m,
                                                         float* __restrict__ dm, double* __restrict__ row_part) {
    __shared__ __align__(16) float As[CD_K][CD_TILE + 4];
    __shared__ __align__(16) float Bs[CD_K][CD_TILE + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long long row0 = (long long)blockIdx.y * CD_TILE, col0 = (long long)blockIdx.x * CD_TILE;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < CD_DIM; k0 += CD_K) {
        // 64 rows x 32 k per operand = 512 float4 loads; 256 threads x 2
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int idx = threadIdx.x + it * 256;
            const int r = idx >> 3, kq = (idx & 7) * 4;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (row0 + r < n) va = __ldg(reinterpret_cast<const float4*>(a + (row0 + r) * CD_DIM + k0 + kq));
            if (col0 + r < m) vb = __ldg(reinterpret_cast<const float4*>(b + (col0 + r) * CD_DIM + k0 + kq));
            As[kq + 0][r] = va.x; As[kq + 1][r] = va.y; As[kq + 2][r] = va.z; As[kq + 3][r] = va.w;
            Bs[kq + 0][r] = vb.x; Bs[kq + 1][r] = vb.y; Bs[kq + 2][r] = vb.z; Bs[kq + 3][r] = vb.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CD_K; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w};
            const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float d = ar[i] - br[j];
                    acc[i][j] = fmaf(d, d, acc[i][j]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long row = row0 + ty * 4 + i;
        float d[4];
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            d[j] = sqrtf(acc[i][j]);
            if (col0 + tx * 4 + j < m) s += (double)d[j];
        }
        // reduce the 16 column-threads of this row (they are 16 consecutive lanes)
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (row < n) {
            if (dm != nullptr) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const long long col = col0 + tx * 4 + j;
                    if (col < m) dm[row * m + col] = d[j];
                }
            }
            if (tx == 0) row_part[(long long)blockIdx.x * n + row] = s;  // one writer per (column tile, row)
        }
    }
}


// Paired (diagonal-only) distances d[i] = ||a[i] - b[i]||_2 for the full-reference evaluation mode of the reference's
// harness (train_triplet.py:267-274 takes the diagonal of cdist): one warp per pair, fp64 accumulation of the fp32
// differences, so the result equals scipy's float64 cdist diagonal to rounding.
__global__ void __launch_bounds__(256) paired_dist_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          long long n, double* __restrict__ out) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    const float4* pa = reinterpret_cast<const float4*>(a + row * CD_DIM);
    const float4* pb = reinterpret_cast<const float4*>(b + row * CD_DIM);
    double s = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float4 x = __ldg(pa + lane + 32 * h), y = __ldg(pb + lane + 32 * h);
        const double d0 = (double)x.x - (double)y.x, d1 = (double)x.y - (double)y.y, d2 = (double)x.z - (double)y.z,
                     d3 = (double)x.w - (double)y.w;
        s += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = sqrt(s);
}

int launch_paired_dist(cudaStream_t st, const float* a, const float* b, long long n, double* out) {
    if (n <= 0) return 0;
    paired_dist_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(a, b, n, out);
    NB_LAUNCHED();
    return 0;
}

__global__ void scale_rows_kernel(double* v, long long n, double s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] *= s;
}

// Row means from the per-(column group, row) partial sums, added in a fixed order (np.mean(axis=1), nomad.py:111):
// bit-identical from run to run, unlike atomics whose order depends on tile scheduling.
template <typename T>
__global__ void __launch_bounds__(256) row_mean_finish_kernel(const T* __restrict__ part, long long n, int groups,
                                                              double inv_m, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int g = 0; g < groups; ++g) s += (double)part[(long long)g * n + i];
    out[i] = s * inv_m;
}

// ---------------------------------------------------------------------------------------------
// Tensor-core path: Gram matrix on tcgen05 from split-fp16 operands.
//   x = hi + lo (hi = fp16(S x), lo = fp16(S x - hi), S = 64 keeps lo out of the fp16 subnormals)
//   <a, b> ~ (a_hi.b_hi + a_lo.b_hi + a_hi.b_lo) / S^2   -- one K = 768 GEMM over [hi | lo | hi] x [hi | hi | lo]
// which carries ~21 significant bits, enough for the 1e-5 bound once near-zero distances are re-evaluated
// exactly in the epilogue (gemm.cu: cdist_chunk).  Norms are fp32 from the original rows.
static constexpr float CD_SCALE = 64.0f;

__global__ void __launch_bounds__(256) cdist_prep_kernel(const float* __restrict__ x, long long n, int is_b,
                                                         op_t* __restrict__ split, float* __restrict__ norm) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    const float4* p = reinterpret_cast<const float4*>(x + row * CD_DIM);
    const float4 a = __ldg(p + 2 * lane), b = __ldg(p + 2 * lane + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float s = 0.f;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float x0 = v[2 * i] * CD_SCALE, x1 = v[2 * i + 1] * CD_SCALE;
        s = fmaf(v[2 * i], v[2 * i], s);
        s = fmaf(v[2 * i + 1], v[2 * i + 1], s);
        hi[i] = pack_op(x0, x1);
        const float2 h = unpack_op(hi[i]);
        lo[i] = pack_op(x0 - h.x, x1 - h.y);
    }
    s = warp_sum(s);
    if (lane == 0) norm[row] = s;
    uint4* o = reinterpret_cast<uint4*>(split + row * (3 * CD_DIM));
    const uint4 H = make_uint4(hi[0], hi[1], hi[2], hi[3]), L = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    // A rows: [hi | lo | hi]   B rows: [hi | hi | lo]
    o[lane] = H;
    o[32 + lane] = is_b ? H : L;
    o[64 + lane] = is_b ? L : H;
}

size_t cdist_tc_workspace(long long n, long long m) {
    auto al = [](size_t v) { return (v + 1023) / 1024 * 1024; };
    return al((size_t)n * 3 * CD_DIM * 2) + al((size_t)m * 3 * CD_DIM * 2) + al((size_t)n * 4) + al((size_t)m * 4) +
           al((size_t)n * 4 * (size_t)cdist_row_groups(n, m)) + 1024;
}
size_t cdist_fp32_workspace(long long n, long long m) {
    return (size_t)n * 8 * (size_t)((m + CD_TILE - 1) / CD_TILE) + 1024;
}

int launch_cdist_tc(cudaStream_t st, const float* a, long long n, const float* b, long long m, float* dm,
                    double* row_mean, void* ws, size_t ws_bytes, int impl) {
    if (n <= 0) return 0;
    if (m <= 0) {  // np.mean of an empty row is NaN (the reference would print NaN scores)
        NB_CUDA(cudaMemsetAsync(row_mean, 0xFF, sizeof(double) * n, st));
        return 0;
    }
    NB_CHECK(ws != nullptr && ws_bytes >= cdist_tc_workspace(n, m), "cdist: workspace too small (%zu < %zu bytes)", ws_bytes,
             cdist_tc_workspace(n, m));
    NB_CHECK(n < (1LL << 31) && m < (1LL << 31), "cdist: too many rows for one call; chunk it");
    auto al = [](size_t v) { return (v + 1023) / 1024 * 1024; };
    char* base = (char*)ws;
    op_t* sa = (op_t*)base;
    op_t* sb = (op_t*)(base + al((size_t)n * 3 * CD_DIM * 2));
    float* na = (float*)((char*)sb + al((size_t)m * 3 * CD_DIM * 2));
    float* nbv = (float*)((char*)na + al((size_t)n * 4));
    float* part = (float*)((char*)nbv + al((size_t)m * 4));
    const int groups = cdist_row_groups(n, m);
    if (impl != 0) NB_CUDA(cudaMemsetAsync(row_mean, 0, sizeof(double) * n, st));  // SIMT check kernel: atomics
    cdist_prep_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(a, n, 0, sa, na);
    NB_LAUNCHED();
    cdist_prep_kernel<<<(unsigned)((m + 7) / 8), 256, 0, st>>>(b, m, 1, sb, nbv);
    NB_LAUNCHED();
    GemmOperand A{sa, n, 3 * CD_DIM, 0, 0};
    GemmOperand Bw{sb, m, 3 * CD_DIM, 0, 0};
    GemmEpilogue e;
    memset(&e, 0, sizeof(e));
    e.flags = EPI_CDIST;
    e.out_f = dm;
    e.ldo = m;
    e.norm_a = na;
    e.norm_b = nbv;
    e.row_sum = row_mean;
    e.row_part = part;
    e.cd_a = a;
    e.cd_b = b;
    e.cd_inv_scale = 1.0f / (CD_SCALE * CD_SCALE);
    NB_TRY(gemm_h16(st, A, Bw, (int)n, (int)m, 3 * CD_DIM, 1, e, impl));
    if (impl != 0)
        scale_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(row_mean, n, 1.0 / (double)m);
    else
        row_mean_finish_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, n, groups, 1.0 / (double)m, row_mean);
    NB_LAUNCHED();
    return 0;
}

int launch_cdist_fp32(cudaStream_t st, const float* a, long long n, const float* b, long long m, float* dm,
                      double* row_mean, void* ws, size_t ws_bytes) {
    if (n <= 0) return 0;
    if (m <= 0) {  // np.mean of an empty row is NaN
        NB_CUDA(cudaMemsetAsync(row_mean, 0xFF, sizeof(double) * n, st));
        return 0;
    }
    NB_CHECK(ws != nullptr && ws_bytes >= cdist_fp32_workspace(n, m), "cdist: workspace too small (%zu < %zu bytes)", ws_bytes,
             cdist_fp32_workspace(n, m));
    dim3 grid((unsigned)((m + CD_TILE - 1) / CD_TILE), (unsigned)((n + CD_TILE - 1) / CD_TILE));
    NB_CHECK(grid.y <= 65535u, "cdist: too many rows for one launch (%lld); chunk the call", n);
    double* part = (double*)ws;
    cdist_fp32_kernel<<<grid, 256, 0, st>>>(a, n, b, m, dm, part);
    NB_LAUNCHED();
    row_mean_finish_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, n, (int)grid.x, 1.0 / (double)m, row_mean);
    NB_LAUNCHED();
    return 0;
}

}  // namespace nb
