// C-ABI plumbing shared by every entry point: last-error string, device queries.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

extern "C" void mtl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* mtl_last_error_string(void) { return g_err; }

extern "C" int mtl_abi_version(void) { return 1; }

static long long g_launches = 0;
extern "C" void mtl_count_launch(void) { ++g_launches; }
// number of kernels this library has launched (or captured into a CUDA graph) so far
extern "C" long long mtl_launch_count(void) { return g_launches; }

int mtl_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
    sms = prop.multiProcessorCount;
  }
  return sms;
}

extern "C" int mtl_device_sm_count(void) { return mtl_num_sms(); }
