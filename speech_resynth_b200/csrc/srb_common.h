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

}  // namespace srb
