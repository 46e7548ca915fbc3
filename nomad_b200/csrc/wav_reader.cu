// Native reader of 16-bit PCM RIFF/WAVE files: the decode step of ``Nomad.load_processing`` (reference nomad.py:192-212,
// ``torchaudio.load``) for the common corpus format, done by a few host threads straight into the caller's (pinned)
// staging buffer.  Python's ``wave`` module costs ~0.25 ms of interpreter time per file, which is what held
// ``nomad.predict`` from files at 31 k utterance-seconds per second against 50 k from memory (profiles/
// r02_final_bench_files.json).  Host code only (no CUDA).  Anything that is not plain 16-bit PCM (format tag 1, or
// WAVE_FORMAT_EXTENSIBLE with the PCM sub-format) is reported as "not handled" and takes the Python path.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/nomad_b200.h"
#include "common.cuh"

namespace nb {

struct WavInfo {
    int sr = 0, channels = 0;
    long long frames = -1, data_offset = 0;
};

static uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// parse the RIFF header; frames = -1 when the file is not 16-bit PCM wav
static WavInfo probe_one(const char* path) {
    WavInfo w;
    FILE* f = fopen(path, "rb");
    if (!f) return w;
    unsigned char h[12];
    if (fread(h, 1, 12, f) != 12 || memcmp(h, "RIFF", 4) != 0 || memcmp(h + 8, "WAVE", 4) != 0) {
        fclose(f);
        return w;
    }
    bool have_fmt = false;
    int bits = 0, block_align = 0;
    long long pos = 12;
    for (;;) {
        unsigned char ch[8];
        if (fseek(f, (long)pos, SEEK_SET) != 0 || fread(ch, 1, 8, f) != 8) break;
        const uint32_t size = rd32(ch + 4);
        if (memcmp(ch, "fmt ", 4) == 0) {
            unsigned char fm[40];
            const size_t n = size < 40 ? size : 40;
            if (n < 16 || fread(fm, 1, n, f) != n) break;
            uint16_t tag = rd16(fm);
            w.channels = rd16(fm + 2);
            w.sr = (int)rd32(fm + 4);
            block_align = rd16(fm + 12);
            bits = rd16(fm + 14);
            if (tag == 0xFFFE && n >= 26) tag = rd16(fm + 24);  // WAVE_FORMAT_EXTENSIBLE: first word of the sub-format GUID
            have_fmt = tag == 1 && bits == 16 && w.channels >= 1 && block_align == 2 * w.channels;
            if (!have_fmt) break;
        } else if (memcmp(ch, "data", 4) == 0) {
            if (have_fmt) {
                fseek(f, 0, SEEK_END);
                const long long file_size = ftell(f);
                long long bytes = size;
                if (pos + 8 + bytes > file_size) bytes = file_size - (pos + 8);  // truncated file / streaming header
                w.data_offset = pos + 8;
                w.frames = bytes / block_align;
            }
            break;
        }
        pos += 8 + (long long)size + (size & 1);
    }
    fclose(f);
    return w;
}

template <typename F>
static void parallel_for(long long n, int threads, F fn) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > n) threads = (int)(n > 0 ? n : 1);
    std::atomic<long long> next{0};
    auto work = [&] {
        for (long long i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}

}  // namespace nb

extern "C" {

int nomad_b200_wav_probe(const char* const* paths, int64_t n, int32_t* sample_rate, int32_t* channels, int64_t* frames,
                         int64_t* data_offset, int threads) {
    NB_CHECK(n >= 0 && (n == 0 || (paths && sample_rate && channels && frames && data_offset)), "wav_probe: bad arguments");
    nb::parallel_for(n, threads, [&](long long i) {
        const nb::WavInfo w = nb::probe_one(paths[i]);
        sample_rate[i] = w.sr;
        channels[i] = w.channels;
        frames[i] = w.frames;
        data_offset[i] = w.data_offset;
    });
    return 0;
}

int nomad_b200_wav_read_pcm16(const char* const* paths, int64_t n, const int64_t* data_offset, const int64_t* n_samples,
                              const int64_t* dst_offset, int16_t* dst, int threads) {
    NB_CHECK(n >= 0 && (n == 0 || (paths && data_offset && n_samples && dst_offset && dst)), "wav_read_pcm16: bad arguments");
    std::atomic<long long> failed{-1};
    nb::parallel_for(n, threads, [&](long long i) {
        if (n_samples[i] <= 0) return;
        FILE* f = fopen(paths[i], "rb");
        bool ok = f != nullptr && fseek(f, (long)data_offset[i], SEEK_SET) == 0 &&
                  fread(dst + dst_offset[i], 2, (size_t)n_samples[i], f) == (size_t)n_samples[i];
        if (f) fclose(f);
        if (!ok) failed.store(i);
    });
    NB_CHECK(failed.load() < 0, "wav_read_pcm16: short read on %s", paths[failed.load()]);
    return 0;
}

}  // extern "C"
