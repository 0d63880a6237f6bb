// kr_api.cu — error reporting + version for the C ABI (include/kokoro_b200.h).
#include "kr_common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void kr_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

static unsigned long long g_launches = 0;
void kr_count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

extern "C" const char* kr_last_error(void) { return g_err; }
// Kernels launched by this library since it was loaded (every launch site counts itself).
extern "C" long long kr_launch_count(void) { return (long long)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" int kr_abi_version(void) { return 1; }

// Device sanity probe: returns the compute capability major*10+minor of the current device, or <0.
extern "C" int kr_device_cc(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    kr_set_error("no CUDA device");
    return KR_ERR_CUDA;
  }
  return prop.major * 10 + prop.minor;
}
