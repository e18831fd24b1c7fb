/* Test infrastructure only (see oracle/README.md).
 * LD_PRELOAD shim: pins time() so the UNMODIFIED reference sampler
 * (/root/reference/preprocess/gen_merw.cpp:161 `srand(time(0))`) becomes
 * deterministic.  FAKE_TIME=<seed> selects the srand() seed. */
#include <stdlib.h>
#include <time.h>

time_t time(time_t *t)
{
    const char *s = getenv("FAKE_TIME");
    time_t v = s ? (time_t)atoll(s) : (time_t)12345;
    if (t) *t = v;
    return v;
}
