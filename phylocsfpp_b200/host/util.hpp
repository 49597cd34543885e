// util.hpp — small host utilities of the phylocsf_b200 command line tool.
//
// my_format mirrors my_fprintf (reference src/common.hpp:48-68): the value is narrowed to float, printed with
// the given format, trailing zeros are stripped but one decimal is kept ("24.834", "3.54", "2.0", "-0.0").
#pragma once

#include <emmintrin.h>
#include <sys/stat.h>

#include <algorithm>
#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace host {

[[noreturn]] inline void die(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fputs("\033[31mError: ", stderr);
    vfprintf(stderr, fmt, ap);
    fputs("\033[0m\n", stderr);
    va_end(ap);
    exit(1);
}

inline std::string lower(std::string s) {
    for (char &c : s) c = (char)tolower((unsigned char)c);
    return s;
}

inline std::vector<std::string> split(const std::string &s, char sep) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i <= s.size()) {
        size_t j = s.find(sep, i);
        if (j == std::string::npos) j = s.size();
        out.push_back(s.substr(i, j - i));
        i = j + 1;
    }
    return out;
}

// common.hpp:36-46 create_directory
inline bool create_directory(const std::string &path) {
    struct stat st;
    if (stat(path.c_str(), &st) == -1) {
        mkdir(path.c_str(), 0764);
        return true;
    }
    return false;
}

// Appends my_fprintf(f, fmt, d) + '\n' to out, decimals = 3 or 4, through snprintf (the reference's own route).
inline void my_format_printf(std::string &out, int decimals, float d) {
    char buf[48];
    int n = snprintf(buf, sizeof buf, decimals == 3 ? "%.3f" : "%.4f", d);
    for (int i = n; i >= 0; --i) {
        if (buf[i] == '.') {
            buf[i + 1] = '0';
            if (i + 2 > n) buf[i + 2] = 0;
            break;
        }
        if (isdigit((unsigned char)buf[i])) {
            if (buf[i] != '0') break;
            buf[i] = 0;
        }
    }
    out.append(buf);
    out.push_back('\n');
}

// Same text without snprintf.  A float has 24 significant bits and 10^3 = 2^3 * 125, 10^4 = 2^4 * 625 need 7 / 10 more,
// so (double)d * 10^k is exact and nearbyint (ties to even, the default rounding mode) is exactly the decimal rounding
// printf performs on the exact binary value.
inline void my_format(std::string &out, int decimals, float d) {
    const double scale = decimals == 3 ? 1000.0 : 10000.0;
    const double r = __builtin_nearbyint((double)d * scale);
    if (!(__builtin_fabs(r) < 9.0e15)) { my_format_printf(out, decimals, d); return; }
    char buf[40];
    char *e = buf + sizeof buf;
    char *p = e;
    *--p = '\n';
    uint64_t q = (uint64_t)__builtin_fabs(r);
    const uint64_t iscale = decimals == 3 ? 1000 : 10000;
    uint64_t frac = q % iscale, ip = q / iscale;
    // fractional digits, trailing zeros cut but one digit kept
    int nd = decimals;
    while (nd > 1 && frac % 10 == 0) { frac /= 10; --nd; }
    for (int i = 0; i < nd; ++i) { *--p = (char)('0' + frac % 10); frac /= 10; }
    *--p = '.';
    do { *--p = (char)('0' + ip % 10); ip /= 10; } while (ip);
    if (__builtin_signbit(r)) *--p = '-';
    out.append(p, (size_t)(e - p));
}

// my_format into a caller-provided buffer (at least 48 bytes free); returns the new end.  Same text as my_format: the decimal rounding is
// done by cvtsd2si (ties to even), digits written without loops for the common sizes.
inline char *my_format_to(char *p, int decimals, float d) {
    const double scale = decimals == 3 ? 1000.0 : 10000.0;
    const double x = (double)d * scale;
    if (!(__builtin_fabs(x) < 4.0e9)) {          // huge, inf, nan: the general route
        std::string tmp;
        my_format(tmp, decimals, d);
        memcpy(p, tmp.data(), tmp.size());
        return p + tmp.size();
    }
    // cvtsd2si rounds to nearest, ties to even (the default MXCSR mode): the decimal rounding printf applies to the exact binary value
    const long long ri = _mm_cvtsd_si64(_mm_set_sd(x));
    const uint32_t q = (uint32_t)(ri < 0 ? -ri : ri);
    if (__builtin_signbit(x)) *p++ = '-';          // also for values that round to zero: "-0.0", as nearbyint(-0.4) = -0.0 prints
    const uint32_t iscale = decimals == 3 ? 1000u : 10000u;
    uint32_t ip = q / iscale, frac = q - ip * iscale;
    if (ip < 10) {
        *p++ = (char)('0' + ip);
    } else if (ip < 100) {
        *p++ = (char)('0' + ip / 10);
        *p++ = (char)('0' + ip % 10);
    } else {
        char tmp[12];
        int n = 0;
        do { tmp[n++] = (char)('0' + ip % 10); ip /= 10; } while (ip);
        while (n) *p++ = tmp[--n];
    }
    *p++ = '.';
    if (decimals == 3) {
        const uint32_t d0 = frac / 100, d12 = frac - d0 * 100, d1 = d12 / 10, d2 = d12 - d1 * 10;
        *p++ = (char)('0' + d0);
        if (d12) { *p++ = (char)('0' + d1); if (d2) *p++ = (char)('0' + d2); }
    } else {
        const uint32_t d0 = frac / 1000, r1 = frac - d0 * 1000, d1 = r1 / 100, r2 = r1 - d1 * 100, d2 = r2 / 10, d3 = r2 - d2 * 10;
        *p++ = (char)('0' + d0);
        if (r1) { *p++ = (char)('0' + d1); if (r2) { *p++ = (char)('0' + d2); if (d3) *p++ = (char)('0' + d3); } }
    }
    *p++ = '\n';
    return p;
}

// Append-only text buffer over a std::string with raw-pointer writes (the wig text of a chain: millions of short lines).
struct TextOut {
    std::string &s;
    size_t n;
    explicit TextOut(std::string &str) : s(str), n(str.size()) {}
    inline char *room(size_t want) {          // pointer to at least `want` writable bytes at the end
        if (n + want > s.size()) s.resize(std::max(s.size() * 2, n + want + 4096));
        return &s[n];
    }
    inline void value(int decimals, float d) { char *p = room(48); n = (size_t)(my_format_to(p, decimals, d) - s.data()); }
    inline void text(const char *t) { const size_t l = strlen(t); memcpy(room(l), t, l); n += l; }
    void finish() { s.resize(n); }
};

}  // namespace host
