// Library-wide C-ABI plumbing: error text, device check, launch counter.
#include "common.cuh"
#include <stdarg.h>
#include <mutex>

namespace pl {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int require_device() {
  static int state = 0;  // 0 unknown, 1 ok, -1 failed
  static char why[256];
  if (state == 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      snprintf(why, sizeof(why), "no CUDA device (%s): plslam_b200 has no CPU fallback", cudaGetErrorString(e));
      state = -1;
    } else {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceProp p;
      cudaGetDeviceProperties(&p, dev);
      if (p.major != 10) {
        snprintf(why, sizeof(why), "device %s is sm_%d%d; this library is built for sm_100a only", p.name, p.major, p.minor);
        state = -1;
      } else {
        state = 1;
      }
    }
  }
  if (state < 0) { set_error("%s", why); return PL_ERR_CUDA; }
  return PL_OK;
}
}  // namespace pl

extern "C" const char* pl_last_error(void) { return pl::g_err; }
extern "C" int pl_version(void) { return 100; }
extern "C" unsigned long long pl_launch_count(void) { return pl::g_launches; }
