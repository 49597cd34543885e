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

// Appends my_fprintf(f, fmt, d) + '\n' to out.  decimals = 3 or 4.
inline void my_format(std::string &out, int decimals, float d) {
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

}  // namespace host
