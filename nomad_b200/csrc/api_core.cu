// Error reporting, launch counting and the small stand-alone C-ABI entry points.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "../../include/nomad_b200.h"
#include "gemm.cuh"

namespace nb {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void count_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace nb

extern "C" {

const char* nomad_b200_last_error(void) { return nb::get_error(); }
const char* nomad_b200_version(void) { return "nomad_b200 0.1 (sm_100a)"; }
int64_t nomad_b200_launch_count(void) { return nb::g_launches.load(); }

int nomad_b200_profile_gemm(int enable) {
    nb::gemm_profile_enable(enable != 0);
    return 0;
}
int nomad_b200_profile_gemm_read(double* total_ms, double* total_flops, int64_t* launches) {
    long long n = 0;
    int r = nb::gemm_profile_read(total_ms, total_flops, &n);
    *launches = n;
    return r;
}

int nomad_b200_gemm_f16(const void* a_f16, int64_t a_rows, int64_t lda, int k_wrap, const void* b_f16, int m, int n,
                         int k, int batch, int64_t a_bstride, int64_t b_bstride, int64_t c_bstride, const float* bias,
                         const float* resid, float* c_f32, void* c_f16, int64_t ldc, int flags, int gemm_impl,
                         void* stream) {
    nb::GemmOperand A{(const nb::op_t*)a_f16, a_rows, lda, a_bstride, k_wrap};
    nb::GemmOperand B{(const nb::op_t*)b_f16, n, k, b_bstride, 0};
    nb::GemmEpilogue e;
    memset(&e, 0, sizeof(e));
    e.flags = flags & (nb::EPI_BIAS | nb::EPI_GELU | nb::EPI_RESID | nb::EPI_OUT_F32 | nb::EPI_OUT_H16);
    e.bias = bias;
    e.bias_bstride = n;
    e.resid = resid;
    e.ldr = ldc;
    e.resid_bstride = c_bstride;
    e.out_f = c_f32;
    e.out_h = (nb::op_t*)c_f16;
    e.ldo = ldc;
    e.out_bstride = c_bstride;
    return nb::gemm_h16((cudaStream_t)stream, A, B, m, n, k, batch, e, gemm_impl);
}

int nomad_b200_gemm_split(const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo, int m, int n,
                          int k, float acc_scale, const float* bias, float* c_f32, void* c_hi, void* c_lo, int64_t ldc,
                          int flags, void* stream) {
    nb::GemmOperand A{(const nb::op_t*)a_hi, m, lda, 0, 0, (const nb::op_t*)a_lo};
    nb::GemmOperand B{(const nb::op_t*)b_hi, n, k, 0, 0, (const nb::op_t*)b_lo};
    nb::GemmEpilogue e;
    memset(&e, 0, sizeof(e));
    e.flags = (flags & (nb::EPI_BIAS | nb::EPI_GELU | nb::EPI_OUT_F32 | nb::EPI_OUT_H16)) | nb::EPI_PRECISE;
    e.bias = bias;
    e.bias_bstride = n;
    e.out_f = c_f32;
    e.out_h = (nb::op_t*)c_hi;
    e.out_l = (nb::op_t*)c_lo;
    e.ldo = ldc;
    e.acc_scale = acc_scale;
    return nb::gemm_h16((cudaStream_t)stream, A, B, m, n, k, 1, e, 0);
}

}  // extern "C"
