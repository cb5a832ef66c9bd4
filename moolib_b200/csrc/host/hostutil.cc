// Small process-wide helpers of the host layer: the kernel-launch counter (bench.py's gpu_launches claim) and the
// MOOLIB_B200_TRACE phase watchdog.
#include "common.h"

#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <thread>

namespace mbh {

uint64_t& launch_counter() {
  static uint64_t n = 0;
  return n;
}

namespace {
std::atomic<const char*> g_phase{"idle"};
std::atomic<int64_t> g_phase_ns{0};
}  // namespace

void trace_phase(const char* phase) {
  static const bool enabled = [] {
    const char* e = std::getenv("MOOLIB_B200_TRACE");
    if (!e || !*e || *e == '0') return false;
    std::thread([] {
      const char* last = nullptr;
      while (true) {
        std::this_thread::sleep_for(std::chrono::seconds(1));
        const char* ph = g_phase.load();
        int64_t age = std::chrono::steady_clock::now().time_since_epoch().count() - g_phase_ns.load();
        if (age > 3000000000ll && ph != last) {
          fprintf(stderr, "[moolib_b200 trace pid %d] stuck %.1f s in phase '%s'\n", (int)getpid(), age / 1e9, ph);
          fflush(stderr);
          last = ph;
        } else if (age <= 3000000000ll) {
          last = nullptr;
        }
      }
    }).detach();
    return true;
  }();
  if (!enabled) return;
  g_phase.store(phase);
  g_phase_ns.store(std::chrono::steady_clock::now().time_since_epoch().count());
}

}  // namespace mbh
