// util.hpp — small host utilities of the phylocsf_b200 command line tool.
//
// my_format mirrors my_fprintf (reference src/common.hpp:48-68): the value is narrowed to float, printed with
// the given format, trailing zeros are stripped but one decimal is kept ("24.834", "3.54", "2.0", "-0.0").
#pragma once

#include <sys/stat.h>

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

}  // namespace host
