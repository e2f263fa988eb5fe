// Error plumbing + version for the obia_b200 C ABI.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace obia {

char *err_buf()
{
    static thread_local char buf[512] = {0};
    return buf;
}

int set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// event pairs recorded around the profiled kernel while profiling is enabled
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_ev;
static cudaEvent_t g_prof_open = nullptr;

void prof_begin(cudaStream_t st)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_on) return;
    cudaEvent_t a;
    if (cudaEventCreate(&a) != cudaSuccess) return;
    cudaEventRecord(a, st);
    g_prof_open = a;
}

void prof_end(cudaStream_t st)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_on || !g_prof_open) return;
    cudaEvent_t b;
    if (cudaEventCreate(&b) != cudaSuccess) return;
    cudaEventRecord(b, st);
    g_prof_ev.emplace_back(g_prof_open, b);
    g_prof_open = nullptr;
}

}  // namespace obia

extern "C" int64_t obia_b200_launch_count(void) { return obia::g_launches.load(); }

extern "C" int obia_b200_profile_enable(int on)
{
    std::lock_guard<std::mutex> lk(obia::g_prof_mu);
    obia::g_prof_on = on != 0;
    return OBIA_B200_OK;
}

extern "C" int obia_b200_profile_read(double *total_ms, int64_t *launches)
{
    std::lock_guard<std::mutex> lk(obia::g_prof_mu);
    double tot = 0.0;
    int64_t n = 0;
    for (auto &pr : obia::g_prof_ev) {
        float ms = 0.f;
        if (cudaEventSynchronize(pr.second) == cudaSuccess &&
            cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
            tot += ms;
            ++n;
        }
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    obia::g_prof_ev.clear();
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    return OBIA_B200_OK;
}

extern "C" const char *obia_b200_last_error(void) { return obia::err_buf(); }
extern "C" int obia_b200_version(void) { return 100; }
