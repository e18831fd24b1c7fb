// Error TLS, version, launch counter, dropout hash (host side).
#include <cstdarg>
#include <cstdio>

#include "common.h"
#include "dropout_hash.h"

namespace pn {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
}  // namespace pn

extern "C" int pn_version(void) { return PN_VERSION; }
extern "C" const char *pn_last_error(void) { return pn::g_err; }
extern "C" uint64_t pn_launch_count(void) { return pn::g_launches.load(); }
extern "C" uint32_t pn_dropout_hash(uint64_t seed, uint32_t stream_id, uint64_t idx)
{
    return pn::dropout_hash(seed, stream_id, idx);
}
