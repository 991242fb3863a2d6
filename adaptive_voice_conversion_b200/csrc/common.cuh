// Shared device/host helpers for libavc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/avc_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libavc_b200 is written for sm_100a (B200) only"
#endif

// AVC_PDL=1 builds the programmatic-dependent-launch variant (libavc_b200_pdl.so): every kernel is
// launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with pdl_sync(), so the
// launch latency and prologue of kernel N+1 overlap the tail of kernel N (also as CUDA-graph edges).
// AVC_PDL=0 compiles to exactly the plain <<<>>> launches and kernels of the default library.
#ifndef AVC_PDL
#define AVC_PDL 0
#endif

namespace avc {

// Let the next kernel in the stream start launching, then wait until the previous one has completed
// and flushed.  MUST precede the first global-memory access of a kernel; kernels that allocate
// TMEM call it after the allocation (a dependent CTA parked in the wait must never hold TMEM
// columns that a CTA of the still-running primary kernel is waiting to allocate).
__device__ __forceinline__ void pdl_sync() {
#if AVC_PDL
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

#if AVC_PDL
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface in AVC_CHECK_LAUNCH
}
#define AVC_LAUNCH(kern, grid, block, smem, st, ...) avc::launch_pdl(kern, grid, block, smem, st, __VA_ARGS__)
#else
#define AVC_LAUNCH(kern, grid, block, smem, st, ...) kern<<<grid, block, smem, st>>>(__VA_ARGS__)
#endif

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int opt_tc_uniform_issue();  // runtime options, see avc_set_option
int opt_wgrad_reduce_v2();

#define AVC_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) {                   \
      avc::set_error(__VA_ARGS__);   \
      return (code);                 \
    }                                \
  } while (0)

#define AVC_CHECK_LAUNCH(name)                                                  \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      avc::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
      return AVC_ERR_CUDA;                                                      \
    }                                                                           \
    avc::count_launch();                                                        \
  } while (0)

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Source position of padded/zero-inserted input index p (may be out of range).
// L = logical length (Tin * ups).  Returns -1 for "reads zero".
__device__ __forceinline__ int src_pos(int p, int L, int pad_mode, int ups) {
  if (pad_mode == AVC_PAD_REFLECT) {
    if (p < 0) p = -p;
    if (p >= L) p = 2 * (L - 1) - p;
    if (p < 0 || p >= L) return -1;  // pad wider than the signal: undefined in torch, read 0
  } else {
    if (p < 0 || p >= L) return -1;
  }
  if (ups == 2) {
    if (p & 1) return -1;
    p >>= 1;
  }
  return p;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace avc
