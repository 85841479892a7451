// error state, counters, version
#include "common.cuh"

namespace cc {
std::atomic<long long> g_launches{0};
static thread_local std::string t_err;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_err = buf;
}
const std::string& last_error() { return t_err; }
}  // namespace cc

extern "C" const char* cc_last_error(void) { return cc::last_error().c_str(); }
extern "C" int cc_version(void) { return 1; }
extern "C" long long cc_launch_count(void) { return cc::g_launches.load(); }
extern "C" const char* cc_arch(void) { return "sm_100a"; }
