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

// Device-wide hint for the random gathers of the path (HER rows, self-collision pair tables): how many bytes L2 fetches from
// DRAM per miss (32 / 64 / 128; cudaLimitMaxL2FetchGranularity).  The driver may ignore it.
extern "C" int bmi_set_l2_fetch_granularity(int32_t bytes) {
  BMI_REQUIRE(bytes == 32 || bytes == 64 || bytes == 128, "bmi_set_l2_fetch_granularity: 32, 64 or 128");
  BMI_CUDA_CHECK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
  return BMI_OK;
}
