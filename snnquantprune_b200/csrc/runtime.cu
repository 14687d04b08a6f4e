// Error reporting, device checks and launch accounting for the snnqp C-ABI.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace snnqp {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return SNNQP_ERR_CUDA;
}

int invalid(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return SNNQP_ERR_INVALID;
}

int unsupported(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return SNNQP_ERR_UNSUPPORTED;
}

struct DevInfo {
  int ok = -1;
  int sms = 0;
};
static DevInfo g_dev[64];

static int probe(int dev) {
  if (dev < 0 || dev >= 64) return 0;
  if (g_dev[dev].ok < 0) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
      g_dev[dev].ok = 0;
    } else {
      g_dev[dev].ok = (prop.major == 10) ? 1 : 0;
      g_dev[dev].sms = prop.multiProcessorCount;
    }
  }
  return g_dev[dev].ok;
}

int require_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice (no CUDA device: this library has no CPU fallback)");
  if (!probe(dev)) {
    set_error("device %d is not compute capability 10.x (sm_100a kernels only, no fallback)", dev);
    return SNNQP_ERR_CUDA;
  }
  return SNNQP_OK;
}

int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  probe(dev);
  return g_dev[dev].sms > 0 ? g_dev[dev].sms : 148;
}

void count_launch(int n) { g_launches += n; }

}  // namespace snnqp

extern "C" {

int snnqp_abi_version(void) { return SNNQP_ABI_VERSION; }

const char *snnqp_last_error(void) { return snnqp::g_err; }

int snnqp_device_ok(void) { return snnqp::require_device() == SNNQP_OK ? 1 : 0; }

int64_t snnqp_launch_count(int reset) {
  int64_t v = snnqp::g_launches;
  if (reset) snnqp::g_launches = 0;
  return v;
}

}  // extern "C"
