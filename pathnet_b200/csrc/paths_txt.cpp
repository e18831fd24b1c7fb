// Path hand-off text format: "[n0, .., nL-1, o0, .., oL-1]\n" per walk.
// Writer side of the reference: gen_merw.cpp:188-206; reader side:
// PathNet_run.py:325-334 / :419-423 (`line[1:-2].split(",")`).
#include <cstring>

#include "common.h"

static inline char *put_int(char *p, int32_t v)
{
    char tmp[12];
    int k = 0;
    uint32_t x = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
    do { tmp[k++] = (char)('0' + x % 10); x /= 10; } while (x);
    if (v < 0) *p++ = '-';
    while (k) *p++ = tmp[--k];
    return p;
}

extern "C" int64_t pn_paths_format_txt(const int32_t *nodes, const uint8_t *types, int64_t walks,
                                       int32_t L, char *out, int64_t cap)
{
    if (!nodes || !types || walks < 0 || L < 1) {
        pn::set_error("pn_paths_format_txt: bad arguments");
        return PN_ERR_INVALID;
    }
    // worst case per walk: '[' + L*(11+2) + L*(3+2) + "]\n"
    char line[64 * 18 + 8];
    if (L > 64) { pn::set_error("pn_paths_format_txt: L > 64"); return PN_ERR_INVALID; }
    int64_t pos = 0;
    for (int64_t w = 0; w < walks; w++) {
        char *p = line;
        *p++ = '[';
        for (int t = 0; t < L; t++) { p = put_int(p, nodes[w * L + t]); *p++ = ','; *p++ = ' '; }
        for (int t = 0; t < L; t++) {
            p = put_int(p, (int32_t)types[w * L + t]);
            if (t != L - 1) { *p++ = ','; *p++ = ' '; }
        }
        *p++ = ']'; *p++ = '\n';
        int64_t k = p - line;
        if (out) {
            if (pos + k > cap) { pn::set_error("pn_paths_format_txt: buffer too small"); return PN_ERR_INVALID; }
            memcpy(out + pos, line, (size_t)k);
        }
        pos += k;
    }
    return pos;
}

extern "C" int64_t pn_paths_parse_txt(const char *txt, int64_t bytes, int32_t L, int32_t *nodes,
                                      uint8_t *types, int64_t max_walks)
{
    if (!txt || bytes < 0 || L < 1 || !nodes || !types) {
        pn::set_error("pn_paths_parse_txt: bad arguments");
        return PN_ERR_INVALID;
    }
    int64_t w = 0, i = 0;
    while (i < bytes) {
        while (i < bytes && (txt[i] == '\n' || txt[i] == '\r' || txt[i] == ' ')) i++;
        if (i >= bytes) break;
        if (txt[i] != '[') { pn::set_error("pn_paths_parse_txt: expected '[' at byte %lld", (long long)i); return PN_ERR_INVALID; }
        i++;
        if (w >= max_walks) { pn::set_error("pn_paths_parse_txt: more than %lld walks", (long long)max_walks); return PN_ERR_INVALID; }
        for (int t = 0; t < 2 * L; t++) {
            while (i < bytes && txt[i] == ' ') i++;
            bool neg = false;
            if (i < bytes && txt[i] == '-') { neg = true; i++; }
            if (i >= bytes || txt[i] < '0' || txt[i] > '9') { pn::set_error("pn_paths_parse_txt: expected digit at byte %lld", (long long)i); return PN_ERR_INVALID; }
            int64_t v = 0;
            while (i < bytes && txt[i] >= '0' && txt[i] <= '9') { v = v * 10 + (txt[i] - '0'); i++; }
            if (neg) v = -v;
            if (t < L) nodes[w * L + t] = (int32_t)v; else types[w * L + (t - L)] = (uint8_t)v;
            if (t != 2 * L - 1) {
                if (i >= bytes || txt[i] != ',') { pn::set_error("pn_paths_parse_txt: expected ',' at byte %lld (walk has fewer than %d fields)", (long long)i, 2 * L); return PN_ERR_INVALID; }
                i++;
            }
        }
        if (i >= bytes || txt[i] != ']') { pn::set_error("pn_paths_parse_txt: expected ']' at byte %lld", (long long)i); return PN_ERR_INVALID; }
        i++;
        w++;
    }
    return w;
}
