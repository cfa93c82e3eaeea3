// runtime.cu -- error string, version, launch counter of librepconc_b200.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace rc {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace rc

RC_API const char* rc_last_error(void) { return rc::g_err; }
RC_API const char* rc_version(void) { return "repconc_b200 0.1 sm_100a"; }
RC_API int64_t rc_launch_count(void) { return rc::g_launches.load(std::memory_order_relaxed); }
