// Reader of the sampler's input files: `<name>.in` = header "n m", then m rows "u v p" (gen_merw.cpp:162-172), and the
// 2-column `<name>_nsl.in` of the uniform-RW variant, rows "u v" (gen.cpp:79-94).  The reference reads them with fscanf
// (`%d %d %lf`): a whitespace-separated token stream, values correctly rounded by strtod.  Here the body is cut at line
// ends into one piece per host thread (counted, then filled in place); std::from_chars is correctly rounded as well, so
// the doubles are bit-identical.  Anything that is not "one row per line" falls back to a sequential token-stream read
// with exactly the fscanf semantics.
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "common.h"

namespace {

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

int host_threads(int64_t bytes)
{
    int t = (int)std::thread::hardware_concurrency();
    if (t <= 0) t = 1;
    if (t > 64) t = 64;
    if (const char *e = getenv("PN_HOST_THREADS")) { const int v = atoi(e); if (v >= 1) t = v < 256 ? v : 256; }
    const int64_t by_work = bytes >> 20;   // at least 1 MiB of text per thread
    if (by_work < t) t = (int)(by_work < 1 ? 1 : by_work);
    return t;
}

template <typename F>
void run_parallel(int nthreads, F &&fn)
{
    if (nthreads <= 1) { fn(0); return; }
    std::vector<std::thread> th;
    th.reserve(nthreads - 1);
    for (int i = 1; i < nthreads; i++) th.emplace_back([&fn, i] { fn(i); });
    fn(0);
    for (auto &t : th) t.join();
}

// one token -> int32 / double; returns the position after it, or nullptr if the token is not a number
inline const char *tok_int(const char *p, const char *end, int32_t *out)
{
    const char *q = p;
    if (q < end && *q == '+') q++;
    long long v = 0;
    auto r = std::from_chars(q, end, v);
    if (r.ec != std::errc() || r.ptr == q) return nullptr;
    *out = (int32_t)v;
    return r.ptr;
}
inline const char *tok_double(const char *p, const char *end, double *out)
{
    const char *q = p;
    if (q < end && *q == '+') q++;
    double v = 0;
    auto r = std::from_chars(q, end, v);
    if (r.ec == std::errc::result_out_of_range) {   // strtod semantics: +-HUGE_VAL / 0 with the token consumed
        char tmp[64];
        const size_t k = (size_t)(r.ptr - p) < sizeof(tmp) - 1 ? (size_t)(r.ptr - p) : sizeof(tmp) - 1;
        memcpy(tmp, p, k); tmp[k] = 0;
        v = strtod(tmp, nullptr);
    } else if (r.ec != std::errc() || r.ptr == q) return nullptr;
    *out = v;
    return r.ptr;
}

// rows of a piece that is "one row per line" (exactly `cols` tokens on every non-blank line), at most `limit` of them.
// store == false: count only.  Returns the row count, or -1 if the piece is not in that form.
int64_t parse_lines(const char *p, const char *end, int cols, bool store, int32_t *u, int32_t *v, double *w, int64_t limit)
{
    int64_t rows = 0;
    while (p < end && rows < limit) {
        while (p < end && is_ws(*p)) p++;
        if (p >= end) break;
        int32_t a, b;
        double c = 0;
        const char *q = tok_int(p, end, &a);
        if (!q || q >= end || (*q != ' ' && *q != '\t')) return -1;
        while (q < end && (*q == ' ' || *q == '\t')) q++;
        q = tok_int(q, end, &b);
        if (!q) return -1;
        if (cols == 3) {
            if (q >= end || (*q != ' ' && *q != '\t')) return -1;
            while (q < end && (*q == ' ' || *q == '\t')) q++;
            q = tok_double(q, end, &c);
            if (!q) return -1;
        }
        while (q < end && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
        if (q < end && *q != '\n') return -1;   // more tokens on the line
        if (store) { u[rows] = a; v[rows] = b; if (w) w[rows] = c; }
        rows++;
        p = q;
    }
    return rows;
}

}  // namespace

extern "C" int64_t pn_edges_parse_txt(const char *txt, int64_t bytes, int32_t cols, int64_t max_rows, int32_t *n_out,
                                      int64_t *m_out, int32_t *u, int32_t *v, double *p)
{
    if (!txt || bytes < 0 || (cols != 2 && cols != 3) || !n_out || !m_out) {
        pn::set_error("pn_edges_parse_txt: bad arguments");
        return PN_ERR_INVALID;
    }
    const char *cur = txt, *end = txt + bytes;
    // header: n m
    int32_t n = 0;
    long long m = 0;
    {
        while (cur < end && is_ws(*cur)) cur++;
        const char *q = tok_int(cur, end, &n);
        if (!q) { pn::set_error("pn_edges_parse_txt: header 'n m' not found"); return PN_ERR_INVALID; }
        cur = q;
        while (cur < end && is_ws(*cur)) cur++;
        auto r = std::from_chars(cur, end, m);
        if (r.ec != std::errc() || r.ptr == cur || m < 0) { pn::set_error("pn_edges_parse_txt: header 'n m' not found"); return PN_ERR_INVALID; }
        cur = r.ptr;
    }
    *n_out = n;
    *m_out = m;
    if (!u || !v) return 0;   // header query
    if (m > max_rows) { pn::set_error("pn_edges_parse_txt: header says %lld rows, the arrays hold %lld", m, (long long)max_rows); return PN_ERR_INVALID; }
    if (cols == 3 && !p) { pn::set_error("pn_edges_parse_txt: p is NULL for a 3-column file"); return PN_ERR_INVALID; }
    // fast path: one row per line, pieces cut at line ends
    {
        const int64_t body = end - cur;
        const int T = host_threads(body);
        std::vector<const char *> cut((size_t)T + 1, end);
        cut[0] = cur;
        for (int i = 1; i < T; i++) {
            const char *c = cur + body * i / T;
            if (c < cut[(size_t)i - 1]) c = cut[(size_t)i - 1];
            const void *nl = c < end ? memchr(c, '\n', (size_t)(end - c)) : nullptr;
            cut[(size_t)i] = nl ? (const char *)nl + 1 : end;
        }
        std::vector<int64_t> cnt((size_t)T + 1, 0);
        std::vector<char> ok((size_t)T, 1);
        run_parallel(T, [&](int i) {
            const int64_t k = parse_lines(cut[(size_t)i], cut[(size_t)i + 1], cols, false, nullptr, nullptr, nullptr, INT64_MAX);
            if (k < 0) ok[(size_t)i] = 0; else cnt[(size_t)i + 1] = k;
        });
        bool regular = true;
        for (int i = 0; i < T; i++) { regular = regular && ok[(size_t)i]; cnt[(size_t)i + 1] += cnt[(size_t)i]; }
        if (regular && cnt[(size_t)T] >= m) {
            // rows beyond m are ignored like the reference's loop (for i < m)
            run_parallel(T, [&](int i) {
                const int64_t r0 = cnt[(size_t)i];
                if (r0 < m) parse_lines(cut[(size_t)i], cut[(size_t)i + 1], cols, true, u + r0, v + r0, p ? p + r0 : nullptr, m - r0);
            });
            return m;
        }
        if (regular) { pn::set_error("pn_edges_parse_txt: header says %lld rows, found %lld", m, (long long)cnt[(size_t)T]); return PN_ERR_INVALID; }
    }
    // general path: fscanf-like token stream
    for (long long r = 0; r < m; r++) {
        int32_t a, b;
        double c = 0;
        while (cur < end && is_ws(*cur)) cur++;
        const char *q = tok_int(cur, end, &a);
        if (q) { cur = q; while (cur < end && is_ws(*cur)) cur++; q = tok_int(cur, end, &b); }
        if (q && cols == 3) { cur = q; while (cur < end && is_ws(*cur)) cur++; q = tok_double(cur, end, &c); }
        if (!q) { pn::set_error("pn_edges_parse_txt: header says %lld rows, found %lld", m, r); return PN_ERR_INVALID; }
        cur = q;
        u[r] = a; v[r] = b; if (p) p[r] = c;
    }
    return m;
}
