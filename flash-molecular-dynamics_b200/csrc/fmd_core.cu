// Error state, version, device info.
#include <mutex>
#include <string>

#include "fmd_common.cuh"

static thread_local std::string g_err;

extern "C" void fmd_set_error(const char* msg) { g_err = msg ? msg : ""; }
extern "C" const char* fmd_last_error(void) { return g_err.c_str(); }
extern "C" int fmd_version(void) { return 100; }

int fmd_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}
extern "C" int fmd_sm_count(void) { return fmd_num_sms(); }
