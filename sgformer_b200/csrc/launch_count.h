// Host-side count of kernels launched by this library (exported as sgf_launch_count()).
#pragma once
#include <atomic>
#include <stdint.h>
namespace sgf {
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace sgf
