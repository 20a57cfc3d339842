// ABI bookkeeping: version, error strings, device capability check.
#include "common.cuh"

extern "C" int lr2_abi_version(void) { return LR2_ABI_VERSION; }

extern "C" const char* lr2_last_error_string(int code) {
  switch (code) {
    case LR2_OK: return "ok";
    case LR2_ERR_BAD_SHAPE: return "bad shape or null pointer";
    case LR2_ERR_BAD_DTYPE: return "unsupported dtype";
    case LR2_ERR_MISALIGNED: return "pointer or leading dimension not 16-byte aligned";
    case LR2_ERR_WRONG_ARCH: return "device is not compute capability 10.x (sm_100a required)";
    case LR2_ERR_CUDA: return "CUDA launch/runtime error";
    case LR2_ERR_TMA: return "cuTensorMapEncodeTiled failed or unavailable";
    case LR2_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

extern "C" int lr2_check_device(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return LR2_ERR_CUDA;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return LR2_ERR_CUDA;
  return major == 10 ? LR2_OK : LR2_ERR_WRONG_ARCH;
}

#include <atomic>
static std::atomic<long long> g_launches{0};
extern "C" void lr2_note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" long long lr2_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
