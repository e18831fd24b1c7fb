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

// ---- NVTX ---------------------------------------------------------------------
#include <nvtx3/nvToolsExt.h>
namespace pn {
void nvtx_push(const char *name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }
}  // namespace pn

// ---- profiling slots --------------------------------------------------------
#include <map>
#include <mutex>
#include <string>
#include <vector>
namespace pn {
struct ProfSlot { std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev; cudaEvent_t open = nullptr; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::map<std::string, ProfSlot> g_prof;

void prof_begin(const char *slot, cudaStream_t st)
{
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfSlot &s = g_prof[slot];
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    s.open = e;
}
void prof_end(const char *slot, cudaStream_t st)
{
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfSlot &s = g_prof[slot];
    if (!s.open) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    s.ev.emplace_back(s.open, e);
    s.open = nullptr;
}
}  // namespace pn

extern "C" int pn_profile_enable(int on)
{
    std::lock_guard<std::mutex> lk(pn::g_prof_mu);
    for (auto &kv : pn::g_prof) {
        for (auto &pr : kv.second.ev) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
        if (kv.second.open) cudaEventDestroy(kv.second.open);
    }
    pn::g_prof.clear();
    pn::g_prof_on = on != 0;
    return PN_OK;
}

extern "C" int pn_profile_read(const char *name, double *ms, int64_t *launches)
{
    std::lock_guard<std::mutex> lk(pn::g_prof_mu);
    double total = 0; int64_t n = 0;
    auto it = pn::g_prof.find(name ? name : "");
    if (it != pn::g_prof.end())
        for (auto &pr : it->second.ev) {
            float t = 0.f;
            if (cudaEventSynchronize(pr.second) == cudaSuccess &&
                cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) { total += t; n++; }
        }
    if (ms) *ms = total;
    if (launches) *launches = n;
    return PN_OK;
}

extern "C" int pn_version(void) { return PN_VERSION; }
extern "C" const char *pn_last_error(void) { return pn::g_err; }
extern "C" uint64_t pn_launch_count(void) { return pn::g_launches.load(); }
extern "C" uint32_t pn_dropout_hash(uint64_t seed, uint32_t stream_id, uint64_t idx)
{
    return pn::dropout_hash(seed, stream_id, idx);
}
