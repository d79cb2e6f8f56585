// Host-side helpers shared by the translation units of libsrb.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srb {

// thread-local error text returned by srb_last_error()
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define SRB_CUDA(call)                                   \
  do {                                                   \
    int _rc = ::srb::check_cuda((call), #call);          \
    if (_rc != 0) return _rc;                            \
  } while (0)

#define SRB_REQUIRE(cond, ...)                           \
  do {                                                   \
    if (!(cond)) {                                       \
      ::srb::set_error(__VA_ARGS__);                     \
      return -2;                                         \
    }                                                    \
  } while (0)

inline int after_launch(const char* name) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", name, cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

int num_sms();
#ifdef SRB_TRACE
unsigned long long* debug_trace_buffer();   // debug build only (tools/trace_kernels.py)
#endif

// Programmatic dependent launch: every kernel of the path is launched with programmatic stream serialization, so the
// CTAs of kernel N+1 start (barrier init, TMEM allocation, tensor-map prefetch, constant tables) on SMs that kernel N
// has already left and block in griddepcontrol.wait until kernel N has completed and flushed; kernels touch nothing a
// predecessor reads or writes before that wait.  Stream capture records these as programmatic edges, so the CUDA
// graph of a call keeps the overlap.  SRB_PDL=0 switches it off (A/B measurements).
bool pdl_enabled();

// An L2 access-policy window for one launch: accesses inside [base, base + bytes) are persisting (kept in the L2
// set-aside region), everything else streams as usual.  bytes == 0: no window.
struct L2Window {
  const void* base = nullptr;
  size_t bytes = 0;
};
// reserves the persisting share of L2 once per device (no-op afterwards); returns false when the device has none
bool l2_persist_reserve();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_window(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     L2Window win, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (win.bytes != 0 && l2_persist_reserve()) {
    attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[n].val.accessPolicyWindow.base_ptr = const_cast<void*>(win.base);
    attr[n].val.accessPolicyWindow.num_bytes = win.bytes;
    attr[n].val.accessPolicyWindow.hitRatio = 1.0f;
    attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  return launch_pdl_window(kernel, grid, block, smem, stream, L2Window(), static_cast<Args&&>(args)...);
}

}  // namespace srb
