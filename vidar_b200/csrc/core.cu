// Error reporting / identification for libvidar_b200.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace vidar {

std::atomic<int64_t> g_launches{0};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace vidar

extern "C" const char* vidar_last_error(void) { return vidar::err_buf(); }
extern "C" const char* vidar_version(void) { return "vidar_b200 0.1 sm_100a"; }
extern "C" int64_t vidar_launch_count(void) {
  return vidar::g_launches.load(std::memory_order_relaxed);
}
