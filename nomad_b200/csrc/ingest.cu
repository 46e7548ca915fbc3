// Waveform ingest on the device: what ``Nomad.load_processing`` (reference nomad.py:192-212) does on the host
// between ``torchaudio.load`` and the model -- PCM16 -> float in [-1, 1), mean of the first two channels when the
// file has more than one (nomad.py:199-200), ``torchaudio.transforms.Resample(sr, 16000)`` when the rates differ
// (nomad.py:203-205; torchaudio's default sinc-Hann kernel, lowpass_filter_width 6, rolloff 0.99), optional trim to
// 10 s (nomad.py:208-210).  The bytes that cross PCIe are the 16-bit samples at the file's own rate.
//
// Resampling follows torchaudio/functional/functional.py (_get_sinc_resample_kernel / _apply_sinc_resample_kernel):
// with o = sr / gcd, n = target / gcd, width = ceil(6 o / (0.99 min(o, n))), K = 2 width + o, output sample
// j = m n + p is  sum_k h[p][k] xpad[m o + k],  xpad = x padded by `width` zeros on the left.  The table h is built
// on the host in float64 exactly as torchaudio builds it (including the float32 rounding of -p / n that its
// int64-arange / int division produces) and cast to float32.
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/nomad_b200.h"
#include "common.cuh"

namespace nb {

struct ResampleTable {
    int o = 0, n = 0, width = 0, K = 0;
    float* dev = nullptr;  // [n][K]
    int2* range = nullptr; // [n]: first and one-past-last tap of phase p that is not (numerically) zero
};

static long long gcd_ll(long long a, long long b) {
    while (b) {
        const long long t = a % b;
        a = b;
        b = t;
    }
    return a;
}

static int get_table(int sr, int target, ResampleTable* out) {
    static std::mutex mu;
    static std::map<std::pair<int, long long>, ResampleTable> cache;  // per (device, rates)
    int dev = 0;
    NB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    const std::pair<int, long long> key(dev, ((long long)sr << 32) | (unsigned)target);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return 0;
    }
    const long long g = gcd_ll(sr, target);
    ResampleTable t;
    t.o = (int)(sr / g);
    t.n = (int)(target / g);
    const double lpw = 6.0, rolloff = 0.99;
    const double base = (double)(t.o < t.n ? t.o : t.n) * rolloff;
    t.width = (int)std::ceil(lpw * t.o / base);
    t.K = 2 * t.width + t.o;
    NB_CHECK((long long)t.n * t.K < (1LL << 26), "resample %d -> %d Hz needs a %d x %d filter bank: unsupported rate pair", sr,
             target, t.n, t.K);
    std::vector<float> h((size_t)t.n * t.K);
    const double scale = base / t.o;
    for (int p = 0; p < t.n; ++p) {
        const double tp = (double)(float)((double)(-p) / (double)t.n);  // torch: int64 arange / int -> float32
        for (int k = 0; k < t.K; ++k) {
            double tt = (tp + (double)(k - t.width) / (double)t.o) * base;
            tt = tt < -lpw ? -lpw : (tt > lpw ? lpw : tt);
            const double c = std::cos(tt * M_PI / lpw / 2.0);
            const double window = c * c;
            tt *= M_PI;
            const double s = tt == 0.0 ? 1.0 : std::sin(tt) / tt;
            h[(size_t)p * t.K + k] = (float)(s * window * scale);
        }
    }
    // Outside |t| < 6 the Hann window is cos^2(pi / 2): ~1e-33 in float64, never visible in an fp32 sum.  For
    // 44.1 kHz that leaves ~34 live taps of 475 per phase.
    std::vector<int2> rng(t.n);
    for (int p = 0; p < t.n; ++p) {
        int lo = 0, hi = t.K;
        while (lo < hi && std::fabs(h[(size_t)p * t.K + lo]) < 1e-30f) ++lo;
        while (hi > lo && std::fabs(h[(size_t)p * t.K + hi - 1]) < 1e-30f) --hi;
        rng[p] = make_int2(lo, hi);
    }
    NB_CUDA(cudaMalloc((void**)&t.dev, h.size() * sizeof(float)));
    NB_CUDA(cudaMemcpy(t.dev, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    NB_CUDA(cudaMalloc((void**)&t.range, rng.size() * sizeof(int2)));
    NB_CUDA(cudaMemcpy(t.range, rng.data(), rng.size() * sizeof(int2), cudaMemcpyHostToDevice));
    cache[key] = t;
    *out = t;
    return 0;
}

__device__ __forceinline__ float pcm_mono(const int16_t* __restrict__ pcm, long long i, int channels) {
    // one channel: s / 32768; more: (s0 / 32768 + s1 / 32768) / 2 -- all exact in fp32
    const int16_t* f = pcm + i * channels;
    return channels == 1 ? (float)f[0] * (1.0f / 32768.0f) : ((float)f[0] + (float)f[1]) * (1.0f / 65536.0f);
}

__global__ void __launch_bounds__(256) pcm_to_mono_kernel(const int16_t* __restrict__ pcm, long long n, int channels,
                                                          float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = pcm_mono(pcm, i, channels);
}

// One block = 256 consecutive output samples.  The input window they need is staged in shared memory as mono
// float (decoded once per block); each thread then runs its phase's K taps over it.
__global__ void __launch_bounds__(256) resample_kernel(const int16_t* __restrict__ pcm, long long n_in, int channels,
                                                       const float* __restrict__ h, const int2* __restrict__ range,
                                                       int o, int n, int width, int K, long long n_out,
                                                       float* __restrict__ out, int win_cap) {
    extern __shared__ float xs[];
    const long long j0 = (long long)blockIdx.x * 256;
    const long long m0 = j0 / n;                                  // first input group of the block
    const long long j_last = min(j0 + 255, n_out - 1);
    const long long m1 = j_last / n;
    const long long first = m0 * o - width;                        // input index of xs[0]
    const int win = (int)((m1 - m0) * o + K);
    for (int i = threadIdx.x; i < win && i < win_cap; i += 256) {
        const long long s = first + i;
        xs[i] = (s >= 0 && s < n_in) ? pcm_mono(pcm, s, channels) : 0.f;
    }
    __syncthreads();
    const long long j = j0 + threadIdx.x;
    if (j >= n_out) return;
    const long long m = j / n;
    const int p = (int)(j - m * n);
    const float* hp = h + (long long)p * K;
    const float* x = xs + (m - m0) * o;
    const int2 kr = __ldg(range + p);
    float acc = 0.f;
    for (int k = kr.x; k < kr.y; ++k) acc = fmaf(__ldg(hp + k), x[k], acc);
    out[j] = acc;
}

}  // namespace nb

using namespace nb;

extern "C" {

int64_t nomad_b200_ingest_out_samples(int64_t n_frames, int sr, int target_sr, int trim) {
    if (n_frames < 0 || sr <= 0 || target_sr <= 0) return -1;
    int64_t n = n_frames;
    if (sr != target_sr) {
        const long long g = gcd_ll(sr, target_sr);
        const long long o = sr / g, nn = target_sr / g;
        n = (nn * n_frames + o - 1) / o;  // ceil(new * length / orig)
    }
    if (trim && n > (int64_t)target_sr * 10) n = (int64_t)target_sr * 10;
    return n;
}

int nomad_b200_ingest_pcm16(const int16_t* pcm_dev, int64_t n_frames, int channels, int sr, int target_sr, int trim,
                            float* out_dev, void* stream) {
    NB_CHECK(pcm_dev && out_dev && n_frames >= 0 && channels >= 1 && sr > 0 && target_sr > 0, "ingest: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_out = nomad_b200_ingest_out_samples(n_frames, sr, target_sr, trim);
    if (n_out == 0) return 0;
    if (sr == target_sr) {
        pcm_to_mono_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(pcm_dev, n_out, channels, out_dev);
        NB_LAUNCHED();
        return 0;
    }
    ResampleTable t;
    NB_TRY(get_table(sr, target_sr, &t));
    // input window of one block: the groups its 256 outputs touch, plus the filter length
    const int win_cap = (256 / t.n + 2) * t.o + t.K;
    const size_t smem = (size_t)win_cap * sizeof(float);
    NB_CHECK(smem <= 200 * 1024, "ingest: resampling %d -> %d Hz needs %zu bytes of shared memory per block", sr, target_sr, smem);
    static size_t smem_set_dev[64] = {0};  // the attribute is per device
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0 || cur_dev >= 64) cur_dev = 0;
    size_t& smem_set = smem_set_dev[cur_dev];
    if (smem > 48 * 1024 && smem > smem_set) {
        NB_CUDA(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    resample_kernel<<<(unsigned)((n_out + 255) / 256), 256, smem, st>>>(pcm_dev, n_frames, channels, t.dev, t.range, t.o, t.n,
                                                                         t.width, t.K, n_out, out_dev, win_cap);
    NB_LAUNCHED();
    return 0;
}

}  // extern "C"
