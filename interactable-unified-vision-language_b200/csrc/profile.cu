// Launch accounting and an optional CUDA-event profiler (used by bench.py for the live roofline numbers).
// Events are recorded on the stream the kernels are launched on, immediately around each launch.
#include "../../include/samvit_b200.h"
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace svb {
namespace {
struct Rec { int cat; double flops, bytes; cudaEvent_t e0, e1; };
std::mutex g_mu;
bool g_on = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
std::atomic<long long> g_launches{0};
thread_local int t_open = -1;

cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void prof_begin(int cat, double flops, double bytes, cudaStream_t st) {
    if (!g_on) return;
    std::lock_guard<std::mutex> lk(g_mu);
    Rec r{cat, flops, bytes, get_event(), get_event()};
    cudaEventRecord(r.e0, st);
    g_recs.push_back(r);
    t_open = (int)g_recs.size() - 1;
}
void prof_end(cudaStream_t st) {
    if (!g_on || t_open < 0) return;
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEventRecord(g_recs[t_open].e1, st);
    t_open = -1;
}
}  // namespace svb

extern "C" {
int svb_profile_start(void) {
    std::lock_guard<std::mutex> lk(svb::g_mu);
    svb::g_on = true;
    return 0;
}
int svb_profile_stop(double* ms, double* flops, double* bytes, int64_t* launches) {
    SVB_CHECK_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(svb::g_mu);
    svb::g_on = false;
    for (int c = 0; c < svb::PC_COUNT; ++c) { ms[c] = 0; flops[c] = 0; bytes[c] = 0; launches[c] = 0; }
    // SVB_PROF_DETAIL=1: per distinct (category, FLOPs, bytes) — i.e. per GEMM shape / kernel variant — launch count and mean time
    const bool detail = getenv("SVB_PROF_DETAIL") != nullptr;
    std::map<std::tuple<int, double, double>, std::pair<int, double>> agg;
    for (auto& r : svb::g_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
            ms[r.cat] += t; flops[r.cat] += r.flops; bytes[r.cat] += r.bytes; launches[r.cat] += 1;
            if (detail) { auto& a = agg[std::make_tuple(r.cat, r.flops, r.bytes)]; a.first += 1; a.second += t; }
        }
        svb::g_pool.push_back(r.e0);
        svb::g_pool.push_back(r.e1);
    }
    svb::g_recs.clear();
    for (auto& kv : agg) {
        const double fl = std::get<1>(kv.first), by = std::get<2>(kv.first), mean = kv.second.second / kv.second.first;
        fprintf(stderr, "prof cat %d  GFLOP %10.2f  MB %9.2f  n %4d  mean %9.1f us  total %8.2f ms  %7.1f TF/s %7.1f GB/s\n", std::get<0>(kv.first),
                fl / 1e9, by / 1e6, kv.second.first, mean * 1e3, kv.second.second, fl / mean / 1e9, by / mean / 1e6);
    }
    return 0;
}
int64_t svb_launch_count(void) { return svb::g_launches.load(); }
}
