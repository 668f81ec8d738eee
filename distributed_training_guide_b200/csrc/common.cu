#include "common.cuh"
#include <atomic>

namespace dtg {
static std::atomic<unsigned long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
}  // namespace dtg
