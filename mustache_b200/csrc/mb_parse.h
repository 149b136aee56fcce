// Native reader of the reference's text contact format (host code, no device): the parse half of read_pd()
// (mustache.py:254-263: get_sep, pd.read_csv(header=None), dropna, chromosome filter), multi-threaded over a memory map.
// Strict by design: it handles what Hi-C dumps look like (5 columns `chr pos chr pos value` or 3 columns `pos pos value`,
// integer positions, plain decimal values) and reports MB200_PARSE_UNSUPPORTED for anything pandas would treat specially
// (quotes, missing fields, ragged rows, non-integer positions, NaN tokens, more than 15 significant digits -- beyond which
// pandas' own converter is not correctly rounded), in which case the caller keeps using pandas: same arrays either way.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#ifndef MB200_PARSE_UNSUPPORTED
#define MB200_PARSE_UNSUPPORTED 1
#endif

struct mb200_contacts {
    std::vector<int64_t> a, b;
    std::vector<double> val;
    int ncols = 0;
    bool value_is_int = true;
    int64_t rows_total = 0;         // rows of the file (before the chromosome filter)
};

namespace mbparse {

// str(s).replace('chr', '') == str(c).replace('chr', '')   (is_chr, mustache.py:191-196)
inline std::string strip_chr(const char* p, size_t n) {
    std::string s(p, n), out;
    size_t i = 0;
    while (i < s.size()) {
        if (s.compare(i, 3, "chr") == 0) i += 3; else out.push_back(s[i++]);
    }
    return out;
}

// exact powers of ten: a double holding <= 15 digits times / divided by one of these is ONE correctly rounded operation,
// which is what pandas' precise_xstrtod and strtod both return
static const double kPow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                  1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// decimal token -> double; false when the token is not a plain number this parser reproduces exactly
inline bool parse_value(const char* p, const char* e, double& out, bool& is_int) {
    if (p == e) return false;
    bool neg = false;
    if (*p == '-' || *p == '+') { neg = *p == '-'; ++p; }
    if (p == e) return false;
    uint64_t mant = 0;
    int digits = 0, exp10 = 0;
    bool any = false, frac = false, seen_nonzero = false;
    for (; p < e; ++p) {
        const char c = *p;
        if (c >= '0' && c <= '9') {
            any = true;
            if (c != '0' || seen_nonzero) {
                seen_nonzero = true;
                if (++digits > 15) return false;
                mant = mant * 10 + (uint64_t)(c - '0');
            }
            if (frac) --exp10;
        } else if (c == '.' && !frac) {
            frac = true;
        } else {
            break;
        }
    }
    if (!any) return false;
    bool has_exp = false;
    if (p < e && (*p == 'e' || *p == 'E')) {
        has_exp = true;
        ++p;
        bool eneg = false;
        if (p < e && (*p == '-' || *p == '+')) { eneg = *p == '-'; ++p; }
        if (p == e) return false;
        int ev = 0;
        for (; p < e && *p >= '0' && *p <= '9'; ++p) {
            ev = ev * 10 + (*p - '0');
            if (ev > 400) return false;
        }
        exp10 += eneg ? -ev : ev;
    }
    if (p != e) return false;
    // trailing fractional zeros of the mantissa do not count as digits: fold them back so that "3.0" is 3 * 10^0
    while (mant != 0 && mant % 10 == 0 && exp10 < 0) { mant /= 10; ++exp10; }
    if (mant == 0) exp10 = 0;
    if (exp10 < -22 || exp10 > 22) return false;
    double v = (double)mant;
    v = exp10 >= 0 ? v * kPow10[exp10] : v / kPow10[-exp10];
    out = neg ? -v : v;
    is_int = !frac && !has_exp;
    return true;
}

inline bool parse_int(const char* p, const char* e, int64_t& out) {
    if (p == e) return false;
    bool neg = false;
    if (*p == '-') { neg = true; ++p; }
    if (p == e || e - p > 18) return false;
    int64_t v = 0;
    for (; p < e; ++p) {
        if (*p < '0' || *p > '9') return false;
        v = v * 10 + (*p - '0');
    }
    out = neg ? -v : v;
    return true;
}

struct Chunk {
    std::vector<int64_t> a, b;
    std::vector<double> val;
    bool all_int = true, bad = false;
    int64_t rows = 0;
};

inline void parse_range(const char* base, size_t lo, size_t hi, char sep, int ncols, const std::string& want, Chunk& out) {
    const char* p = base + lo;
    const char* end = base + hi;
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        const char* next = nl ? nl + 1 : end;
        if (le > p && le[-1] == '\r') --le;
        if (le == p) { p = next; continue; }                       // blank line (skip_blank_lines)
        const char* f[6];
        int nf = 0;
        f[0] = p;
        for (const char* q = p; q < le; ++q) {
            if (*q == '"') { out.bad = true; return; }
            if (*q == sep) {
                if (nf + 1 >= 5) { out.bad = true; return; }        // more fields than columns
                f[++nf] = q + 1;
            }
        }
        ++nf;
        f[nf] = le + 1;
        if (nf != ncols) { out.bad = true; return; }
        ++out.rows;
        const int ia = ncols == 5 ? 1 : 0, ib = ncols == 5 ? 3 : 1, iv = ncols == 5 ? 4 : 2;
        int64_t a, b;
        double v;
        bool vint = false;
        if (!parse_int(f[ia], f[ia + 1] - 1, a) || !parse_int(f[ib], f[ib + 1] - 1, b) || !parse_value(f[iv], f[iv + 1] - 1, v, vint)) {
            out.bad = true;
            return;
        }
        if (ncols == 5) {
            if (f[1] - 1 == f[0] || f[3] - 1 == f[2]) { out.bad = true; return; }      // empty chromosome field
            if (strip_chr(f[0], (size_t)(f[1] - 1 - f[0])) != want || strip_chr(f[2], (size_t)(f[3] - 1 - f[2])) != want) {
                out.all_int = out.all_int && vint;                  // dtype inference sees every row of the file
                p = next;
                continue;
            }
        }
        out.all_int = out.all_int && vint;
        out.a.push_back(a);
        out.b.push_back(b);
        out.val.push_back(v);
        p = next;
    }
}

}  // namespace mbparse

// 0 = ok, MB200_PARSE_UNSUPPORTED = let pandas do it, negative = I/O error
inline int mb200_parse_file(const char* path, const char* chromosome, int threads, mb200_contacts* out) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); return -1; }
    const size_t size = (size_t)sb.st_size;
    if (size == 0) { close(fd); return MB200_PARSE_UNSUPPORTED; }
    const char* base = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (base == MAP_FAILED) return -1;
    // get_sep (mustache.py:199-215): decided from the first line only
    const char* nl = (const char*)memchr(base, '\n', size);
    const size_t l0 = nl ? (size_t)(nl - base) : size;
    char sep = 0;
    if (memchr(base, '\t', l0)) sep = '\t';
    else if (memchr(base, ' ', l0)) sep = ' ';
    else if (memchr(base, ',', l0)) sep = ',';
    int rc = 0;
    if (!sep) rc = MB200_PARSE_UNSUPPORTED;
    int ncols = 1;
    for (size_t i = 0; i < l0 && !rc; ++i) ncols += base[i] == sep;
    if (!rc && ncols != 5 && ncols != 3) rc = MB200_PARSE_UNSUPPORTED;
    if (!rc) {
        const std::string want = mbparse::strip_chr(chromosome ? chromosome : "", chromosome ? strlen(chromosome) : 0);
        int nt = std::max(1, std::min(threads > 0 ? threads : (int)std::thread::hardware_concurrency(), 32));
        if (size < (1u << 20)) nt = 1;
        std::vector<size_t> cut(nt + 1, size);
        cut[0] = 0;
        for (int t = 1; t < nt; ++t) {
            size_t pos = size / nt * t;
            const char* q = (const char*)memchr(base + pos, '\n', size - pos);
            cut[t] = q ? (size_t)(q - base) + 1 : size;
        }
        std::vector<mbparse::Chunk> parts(nt);
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t)
            pool.emplace_back([&, t] { if (cut[t] < cut[t + 1]) mbparse::parse_range(base, cut[t], cut[t + 1], sep, ncols, want, parts[t]); });
        for (auto& th : pool) th.join();
        size_t total = 0;
        for (auto& c : parts) {
            if (c.bad) rc = MB200_PARSE_UNSUPPORTED;
            total += c.a.size();
        }
        if (!rc) {
            out->ncols = ncols;
            out->a.reserve(total); out->b.reserve(total); out->val.reserve(total);
            for (auto& c : parts) {
                out->a.insert(out->a.end(), c.a.begin(), c.a.end());
                out->b.insert(out->b.end(), c.b.begin(), c.b.end());
                out->val.insert(out->val.end(), c.val.begin(), c.val.end());
                out->value_is_int = out->value_is_int && c.all_int;
                out->rows_total += c.rows;
            }
        }
    }
    munmap((void*)base, size);
    return rc;
}
