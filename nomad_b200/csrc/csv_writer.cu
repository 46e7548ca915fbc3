// Result formatting at scale: the text pandas writes for ``DataFrame(values).round(3)`` + ``to_csv`` (reference
// nomad.py:113-120, 138-139), produced directly from the float64 score matrix by a few host threads.  pandas needs
// seconds to minutes for a 1e5 x 1e3 frame; the bytes are the contract, not the tool.  Host code only (no CUDA).
//
// What has to be reproduced:
//   * ``round(3)``       numpy: rint(x * 1000) / 1000 in float64 (half-to-even on the scaled value)
//   * float formatting   Python's repr: the shortest digit string that round-trips, fixed notation when
//                        -4 <= exponent < 16 (always with a fractional part: "1.0"), else d.ddde-05 / 1e+16
//   * missing values     NaN -> empty field; +-inf -> "inf" / "-inf"
//   * csv.QUOTE_MINIMAL  labels containing , " CR or LF are quoted, quotes doubled; line terminator "\n"
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nomad_b200.h"
#include "common.cuh"

namespace nb {

// Python repr(float) for a finite value (float_repr_style 'short')
static void append_repr(std::string& out, double v) {
    if (std::isnan(v)) return;  // pandas: na_rep = ''
    if (std::isinf(v)) {
        out += v < 0 ? "-inf" : "inf";
        return;
    }
    char buf[64];
    // shortest round-trip digits in scientific form: d[.ddd]e[+-]XX
    auto res = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    const char* p = buf;
    const char* end = res.ptr;
    if (*p == '-') {
        out += '-';
        ++p;
    }
    char digits[32];
    int nd = 0;
    while (p < end && *p != 'e') {
        if (*p != '.') digits[nd++] = *p;
        ++p;
    }
    int exp10 = 0;
    if (p < end && *p == 'e') {
        ++p;
        const bool neg = *p == '-';
        if (*p == '+' || *p == '-') ++p;
        while (p < end) exp10 = exp10 * 10 + (*p++ - '0');
        if (neg) exp10 = -exp10;
    }
    const int decpt = exp10 + 1;  // position of the decimal point relative to the digit string
    if (decpt > -4 && decpt <= 16) {
        if (decpt <= 0) {
            out += "0.";
            out.append((size_t)(-decpt), '0');
            out.append(digits, (size_t)nd);
        } else if (decpt >= nd) {
            out.append(digits, (size_t)nd);
            out.append((size_t)(decpt - nd), '0');
            out += ".0";
        } else {
            out.append(digits, (size_t)decpt);
            out += '.';
            out.append(digits + decpt, (size_t)(nd - decpt));
        }
    } else {  // d[.ddd]e[+-]XX with at least two exponent digits
        out += digits[0];
        if (nd > 1) {
            out += '.';
            out.append(digits + 1, (size_t)(nd - 1));
        }
        char eb[16];
        snprintf(eb, sizeof(eb), "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
        out += eb;
    }
}

static void append_label(std::string& out, const char* s) {
    bool quote = false;
    for (const char* p = s; *p; ++p)
        if (*p == ',' || *p == '"' || *p == '\n' || *p == '\r') quote = true;
    if (!quote) {
        out += s;
        return;
    }
    out += '"';
    for (const char* p = s; *p; ++p) {
        if (*p == '"') out += '"';
        out += *p;
    }
    out += '"';
}

static inline double round_decimals(double v, int decimals, double pow10) {
    if (decimals < 0 || !std::isfinite(v)) return v;
    return std::nearbyint(v * pow10) / pow10;  // numpy.round for decimals >= 0
}

}  // namespace nb

using namespace nb;

extern "C" {

int nomad_b200_write_scores_csv(const char* path, const char* index_name, const char* const* row_labels, int64_t n_rows,
                                const char* const* col_labels, int64_t n_cols, const double* values, int decimals,
                                int threads) {
    NB_CHECK(path && index_name && n_rows >= 0 && n_cols >= 0 && (n_rows == 0 || row_labels) && (n_cols == 0 || col_labels) &&
                 (n_rows * n_cols == 0 || values),
             "write_scores_csv: bad arguments");
    FILE* f = fopen(path, "wb");
    NB_CHECK(f != nullptr, "write_scores_csv: cannot open %s", path);
    std::string head;
    append_label(head, index_name);
    for (int64_t c = 0; c < n_cols; ++c) {
        head += ',';
        append_label(head, col_labels[c]);
    }
    head += '\n';
    bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
    const double pow10 = std::pow(10.0, decimals < 0 ? 0 : decimals);
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    const int64_t block = 4096;  // rows formatted per thread per round (bounded memory, ordered output)
    for (int64_t r0 = 0; r0 < n_rows && ok; r0 += block * T) {
        std::vector<std::string> parts((size_t)T);
        std::vector<std::thread> pool;
        for (int t = 0; t < T; ++t) {
            const int64_t a = r0 + (int64_t)t * block, b = std::min<int64_t>(a + block, n_rows);
            if (a >= b) break;
            pool.emplace_back([&, t, a, b] {
                std::string& s = parts[(size_t)t];
                s.reserve((size_t)(b - a) * (size_t)(n_cols * 7 + 32));
                for (int64_t r = a; r < b; ++r) {
                    append_label(s, row_labels[r]);
                    const double* v = values + r * n_cols;
                    for (int64_t c = 0; c < n_cols; ++c) {
                        s += ',';
                        append_repr(s, round_decimals(v[c], decimals, pow10));
                    }
                    s += '\n';
                }
            });
        }
        for (auto& th : pool) th.join();
        for (auto& s : parts)
            if (!s.empty() && ok) ok = fwrite(s.data(), 1, s.size(), f) == s.size();
    }
    ok = (fclose(f) == 0) && ok;
    NB_CHECK(ok, "write_scores_csv: write to %s failed", path);
    return 0;
}

}  // extern "C"
