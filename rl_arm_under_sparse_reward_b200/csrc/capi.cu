// Library-wide plumbing of the C-ABI: version, thread-local error string, launch counter.
#include "common.cuh"

namespace bmi {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace bmi

extern "C" int bmi_abi_version(void) { return BMI_ABI_VERSION; }
extern "C" const char* bmi_last_error(void) { return bmi::g_err; }
extern "C" int64_t bmi_launch_count(void) { return (int64_t)bmi::g_launches.load(); }
