// Path hand-off text format: "[n0, .., nL-1, o0, .., oL-1]\n" per walk.
// Writer side of the reference: gen_merw.cpp:188-206; reader side:
// PathNet_run.py:325-334 / :419-423 (`line[1:-2].split(",")`).
//
// The sampler produces walks ~1000x faster than one core can print them, so both directions run on all host threads:
// the walk range (writer) or the byte range cut at line ends (reader) is split into per-thread pieces, a first pass
// sizes each piece (bytes, or walks), a prefix sum places them, a second pass does the work in place.
// PN_HOST_THREADS caps the thread count (default: hardware concurrency, at most 64).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "common.h"

namespace {

int host_threads(int64_t work_items, int64_t min_per_thread)
{
    int t = (int)std::thread::hardware_concurrency();
    if (t <= 0) t = 1;
    if (t > 64) t = 64;
    if (const char *e = getenv("PN_HOST_THREADS")) { const int v = atoi(e); if (v >= 1) t = v < 256 ? v : 256; }
    const int64_t by_work = work_items / (min_per_thread > 0 ? min_per_thread : 1);
    if (by_work < t) t = (int)(by_work < 1 ? 1 : by_work);
    return t;
}

template <typename F>
void run_parallel(int nthreads, F &&fn)   // fn(thread index)
{
    if (nthreads <= 1) { fn(0); return; }
    std::vector<std::thread> th;
    th.reserve(nthreads - 1);
    for (int i = 1; i < nthreads; i++) th.emplace_back([&fn, i] { fn(i); });
    fn(0);
    for (auto &t : th) t.join();
}

const char DIGIT_PAIRS[201] =
    "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
    "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";

inline int ndigits(uint32_t x)
{
    if (x < 10u) return 1;
    if (x < 100u) return 2;
    if (x < 1000u) return 3;
    if (x < 10000u) return 4;
    if (x < 100000u) return 5;
    if (x < 1000000u) return 6;
    if (x < 10000000u) return 7;
    if (x < 100000000u) return 8;
    if (x < 1000000000u) return 9;
    return 10;
}
inline int int_len(int32_t v) { return v < 0 ? 1 + ndigits((uint32_t)(-(int64_t)v)) : ndigits((uint32_t)v); }

inline char *put_int(char *p, int32_t v)
{
    uint32_t x;
    if (v < 0) { *p++ = '-'; x = (uint32_t)(-(int64_t)v); } else x = (uint32_t)v;
    const int n = ndigits(x);
    char *q = p + n;
    while (x >= 100u) {
        const uint32_t r = x % 100u;
        x /= 100u;
        q -= 2;
        memcpy(q, DIGIT_PAIRS + 2 * r, 2);
    }
    if (x >= 10u) memcpy(q - 2, DIGIT_PAIRS + 2 * x, 2);
    else q[-1] = (char)('0' + x);
    return p + n;
}

// bytes of one walk's line: '[' + fields + ", " between the 2L fields + "]\n"
inline int64_t line_len(const int32_t *nd, const uint8_t *ty, int L)
{
    int64_t k = 3 + 2 * (2 * (int64_t)L - 1);
    for (int t = 0; t < L; t++) k += int_len(nd[t]) + ndigits(ty[t]);
    return k;
}

inline char *put_line(char *p, const int32_t *nd, const uint8_t *ty, int L)
{
    *p++ = '[';
    for (int t = 0; t < L; t++) { p = put_int(p, nd[t]); *p++ = ','; *p++ = ' '; }
    for (int t = 0; t < L; t++) {
        p = put_int(p, (int32_t)ty[t]);
        if (t != L - 1) { *p++ = ','; *p++ = ' '; }
    }
    *p++ = ']'; *p++ = '\n';
    return p;
}

struct ParseErr { int64_t at; char msg[160]; };

// sequential reader over txt[i0, i1): fills walks from index w0 (when nodes != null), returns the number of walks seen
// or -1 with *err set.  limit = walks the output arrays can still take from w0.
int64_t parse_range(const char *txt, int64_t i0, int64_t i1, int32_t L, int32_t *nodes, uint8_t *types, int64_t w0, int64_t limit,
                    ParseErr *err)
{
    auto fail = [&](int64_t at, const char *fmt, long long a, int b) {
        err->at = at;
        snprintf(err->msg, sizeof(err->msg), fmt, a, b);
        return (int64_t)-1;
    };
    int64_t w = 0, i = i0;
    while (i < i1) {
        while (i < i1 && (txt[i] == '\n' || txt[i] == '\r' || txt[i] == ' ')) i++;
        if (i >= i1) break;
        if (txt[i] != '[') return fail(i, "pn_paths_parse_txt: expected '[' at byte %lld", (long long)i, 0);
        i++;
        if (w >= limit) return fail(i, "pn_paths_parse_txt: more than %lld walks", (long long)(w0 + limit), 0);
        for (int t = 0; t < 2 * L; t++) {
            while (i < i1 && txt[i] == ' ') i++;
            bool neg = false;
            if (i < i1 && txt[i] == '-') { neg = true; i++; }
            if (i >= i1 || txt[i] < '0' || txt[i] > '9') return fail(i, "pn_paths_parse_txt: expected digit at byte %lld", (long long)i, 0);
            int64_t v = 0;
            while (i < i1 && txt[i] >= '0' && txt[i] <= '9') { v = v * 10 + (txt[i] - '0'); i++; }
            if (neg) v = -v;
            if (nodes) {
                if (t < L) nodes[(w0 + w) * L + t] = (int32_t)v;
                else types[(w0 + w) * L + (t - L)] = (uint8_t)v;
            }
            if (t != 2 * L - 1) {
                if (i >= i1 || txt[i] != ',')
                    return fail(i, "pn_paths_parse_txt: expected ',' at byte %lld (walk has fewer than %d fields)", (long long)i, 2 * L);
                i++;
            }
        }
        if (i >= i1 || txt[i] != ']') return fail(i, "pn_paths_parse_txt: expected ']' at byte %lld", (long long)i, 0);
        i++;
        w++;
    }
    return w;
}

}  // namespace

extern "C" int64_t pn_paths_format_txt(const int32_t *nodes, const uint8_t *types, int64_t walks,
                                       int32_t L, char *out, int64_t cap)
{
    if (!nodes || !types || walks < 0 || L < 1) {
        pn::set_error("pn_paths_format_txt: bad arguments");
        return PN_ERR_INVALID;
    }
    if (L > 64) { pn::set_error("pn_paths_format_txt: L > 64"); return PN_ERR_INVALID; }
    const int T = host_threads(walks, 1 << 15);
    std::vector<int64_t> off((size_t)T + 1, 0);
    auto range = [&](int i, int64_t &a, int64_t &b) { a = walks * i / T; b = walks * (i + 1) / T; };
    run_parallel(T, [&](int i) {
        int64_t a, b, k = 0;
        range(i, a, b);
        for (int64_t w = a; w < b; w++) k += line_len(nodes + w * L, types + w * L, L);
        off[(size_t)i + 1] = k;
    });
    for (int i = 0; i < T; i++) off[(size_t)i + 1] += off[(size_t)i];
    const int64_t total = off[(size_t)T];
    if (!out) return total;   // size query
    if (total > cap) { pn::set_error("pn_paths_format_txt: buffer too small (%lld bytes needed)", (long long)total); return PN_ERR_INVALID; }
    run_parallel(T, [&](int i) {
        int64_t a, b;
        range(i, a, b);
        char *p = out + off[(size_t)i];
        for (int64_t w = a; w < b; w++) p = put_line(p, nodes + w * L, types + w * L, L);
    });
    return total;
}

extern "C" int64_t pn_paths_parse_txt(const char *txt, int64_t bytes, int32_t L, int32_t *nodes,
                                      uint8_t *types, int64_t max_walks)
{
    if (!txt || bytes < 0 || L < 1 || !nodes || !types) {
        pn::set_error("pn_paths_parse_txt: bad arguments");
        return PN_ERR_INVALID;
    }
    const int T = host_threads(bytes, 1 << 20);
    if (T == 1) {   // one piece: a single pass that stores as it goes
        ParseErr e;
        e.at = -1;
        const int64_t k = parse_range(txt, 0, bytes, L, nodes, types, 0, max_walks, &e);
        if (k < 0) { pn::set_error("%s", e.msg); return PN_ERR_INVALID; }
        return k;
    }
    // cut points: piece i starts right after the first '\n' at or after bytes*i/T (a line never straddles two pieces)
    std::vector<int64_t> cut((size_t)T + 1, bytes);
    cut[0] = 0;
    for (int i = 1; i < T; i++) {
        int64_t c = bytes * i / T;
        if (c < cut[(size_t)i - 1]) c = cut[(size_t)i - 1];
        const void *nl = c < bytes ? memchr(txt + c, '\n', (size_t)(bytes - c)) : nullptr;
        cut[(size_t)i] = nl ? (const char *)nl - txt + 1 : bytes;
    }
    std::vector<int64_t> cnt((size_t)T + 1, 0);
    std::vector<ParseErr> errs((size_t)T);
    for (auto &e : errs) e.at = -1;
    // pass 1: validate and count walks per piece (nothing is stored)
    run_parallel(T, [&](int i) {
        const int64_t k = parse_range(txt, cut[(size_t)i], cut[(size_t)i + 1], L, nullptr, nullptr, 0, INT64_MAX, &errs[(size_t)i]);
        cnt[(size_t)i + 1] = k < 0 ? 0 : k;
    });
    for (int i = 0; i < T; i++) {   // the earliest malformed byte wins, as in a sequential read
        if (errs[(size_t)i].at >= 0) { pn::set_error("%s", errs[(size_t)i].msg); return PN_ERR_INVALID; }
        cnt[(size_t)i + 1] += cnt[(size_t)i];
        if (cnt[(size_t)i + 1] > max_walks) { pn::set_error("pn_paths_parse_txt: more than %lld walks", (long long)max_walks); return PN_ERR_INVALID; }
    }
    // pass 2: fill
    run_parallel(T, [&](int i) {
        parse_range(txt, cut[(size_t)i], cut[(size_t)i + 1], L, nodes, types, cnt[(size_t)i], cnt[(size_t)i + 1] - cnt[(size_t)i], &errs[(size_t)i]);
    });
    return cnt[(size_t)T];
}
