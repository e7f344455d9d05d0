// Library plumbing: error string, device checks.
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace avt {

static thread_local char g_err[512] = "";

void set_last_error(const char* what, const char* detail, const char* file, int line) {
  snprintf(g_err, sizeof g_err, "%s: %s (%s:%d)", what, detail, file, line);
}

static int g_sm_limit = 0;
static int g_pdl = -1;   // -1: read AVT_PDL from the environment on first use (default on)

bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("AVT_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}

int num_sms_physical();
int num_sms() {
  const int n = num_sms_physical();
  return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

int num_sms_physical() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
      n = 148;
  }
  return n;
}

static long long g_launches = 0;
void count_launch() { ++g_launches; }   // (host calls are single-threaded per process)

}  // namespace avt

extern "C" int avt_abi_version(void) { return 1; }
extern "C" long long avt_kernel_launch_count(void) { return avt::g_launches; }
extern "C" int avt_set_sm_limit(int n) {
  avt::g_sm_limit = n > 0 ? (n & ~1) : 0;  // even, so CTA pairs still tile the budget
  return AVT_OK;
}
extern "C" int avt_set_pdl(int enable) {
  avt::g_pdl = enable ? 1 : 0;
  return AVT_OK;
}
extern "C" const char* avt_last_error(void) { return avt::g_err; }
extern "C" int avt_check_device(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    avt::set_last_error("avt_check_device", "no CUDA device", __FILE__, __LINE__);
    return AVT_ERR_NO_GPU;
  }
  if (major != 10) {
    avt::set_last_error("avt_check_device", "device is not compute capability 10.x (B200)", __FILE__, __LINE__);
    return AVT_ERR_NO_GPU;
  }
  return AVT_OK;
}
