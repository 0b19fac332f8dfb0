// runtime.cu — host-side plumbing of libb200decode: error strings, launch
// counter, cached device properties, driver entry point for tensor maps.
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches += n; }

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

int pdl_level() {
  static const int level = [] {
    const char* e = getenv("B200_PDL");
    return e ? atoi(e) : 1;
  }();
  return level;
}

tensor_map_encode_fn get_tensor_map_encode() {
  static std::once_flag once;
  static tensor_map_encode_fn fn = nullptr;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tensor_map_encode_fn>(p);
  });
  return fn;
}

}  // namespace b200

extern "C" {

int b200_abi_version(void) { return 1; }

const char* b200_last_error(void) { return b200::g_err; }

int64_t b200_launch_count(void) { return b200::g_launches; }

void b200_launch_count_reset(void) { b200::g_launches = 0; }

}  // extern "C"
