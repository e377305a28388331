// Helpers shared by the extern "C" translation units (capi.cu: production ABI, capi_debug.cu: self-test library).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "ttt_internal.h"

static thread_local char g_err[512] = "";

static int fail(int code, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s", what);
  return code;
}
static int cuda_ret(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s [%s]: %s (%s)", where, tb::g_where, cudaGetErrorName(e), cudaGetErrorString(e));
  return (int)e;
}

// Device that owns `p` (-1 + error message if it is not a device pointer).
typedef CUresult (*PFN_ptrAttr)(void*, CUpointer_attribute, CUdeviceptr);
static int device_of(const void* p, int* ord_out) {
  static PFN_ptrAttr fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_ptrAttr>(ptr);
  }
  int ord = -1;
  if (!fn || fn(&ord, CU_POINTER_ATTRIBUTE_DEVICE_ORDINAL, (CUdeviceptr)(uintptr_t)p) != CUDA_SUCCESS || ord < 0) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess || a.type != cudaMemoryTypeDevice) {
      cudaGetLastError();
      return fail(-5, "pointer argument is not a device pointer");
    }
    ord = a.device;
  }
  *ord_out = ord;
  return 0;
}

// RAII: make the device that owns `p` current for the duration of one ABI call and restore the caller's device on exit.
// Nothing is cached: the host framework may switch the thread's current device behind our back between calls
// (torch.cuda.set_device / device guards), and PyTorch's autograd worker threads (or any fresh thread) have no current
// CUDA context in this library's statically linked runtime instance -- driver calls such as cuTensorMapEncodeTiled
// would fail with CUDA_ERROR_INVALID_CONTEXT and launches would go to device 0.
struct DeviceBinding {
  int prev = -1, rc = 0;
  explicit DeviceBinding(const void* p) {
    int ord = -1;
    rc = device_of(p, &ord);
    if (rc) return;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    cudaError_t e = cudaSetDevice(ord);  // also establishes the primary context on a fresh thread
    if (e != cudaSuccess) { rc = cuda_ret(e, "cudaSetDevice"); prev = -1; return; }
    if (prev == ord) prev = -1;  // nothing to restore
  }
  ~DeviceBinding() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceBinding(const DeviceBinding&) = delete;
  DeviceBinding& operator=(const DeviceBinding&) = delete;
};
#define TB_BIND_DEVICE(p)  \
  DeviceBinding bind__(p); \
  if (bind__.rc) return bind__.rc
