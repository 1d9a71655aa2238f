// Shared helpers for the svk kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/svk.h"

#define SVK_API extern "C" __attribute__((visibility("default")))

void svk_set_error(const char* fmt, ...);
void svk_count_launch(int n = 1);

#define SVK_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      svk_set_error(__VA_ARGS__);               \
      return (code);                            \
    }                                           \
  } while (0)

// Check the launch that just happened; count it.
#define SVK_LAUNCH_CHECK(name)                                            \
  do {                                                                    \
    cudaError_t e__ = cudaGetLastError();                                 \
    if (e__ != cudaSuccess) {                                             \
      svk_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return (int)e__;                                                    \
    }                                                                     \
    svk_count_launch();                                                   \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ----------------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the library is launched with the programmatic-stream-serialization
// attribute: it may START (be scheduled, run its prologue: barrier init, TMEM allocation, tensor-map prefetch) while the
// previous kernel of the stream is still draining, and it calls pdl_wait() — which returns once that kernel has completed
// and its memory is visible — before its first access to global memory.  pdl_launch_dependents() at the top of a kernel
// lets the NEXT kernel's blocks be scheduled as soon as all blocks of this one have started.  The training step is ~250
// back-to-back kernels of 20-150 us; the ~2.5 us launch gap between two of them (profiles/r02_timeline_before_pdl.json:
// 0.58 ms of a 9.4 ms step) is what this hides.  RULE: every __global__ function calls pdl_wait() unconditionally, in
// every block, before touching global memory — a kernel that skipped it could finish before its predecessor and release
// ITS successor too early.  SVK_DISABLE_PDL=1 launches without the attribute (the device-side calls are then no-ops).
// WHERE the trigger goes matters (measured on one box, bench.py --steps 30, ms per step; profiles/r02_pdl.md):
//     no PDL 9.37 | no explicit trigger (dependents start when the last block EXITS) 9.12 | tensor-core kernels trigger at
//     their start 9.48 | every kernel triggers at its start 9.52.
// An early trigger makes the next kernel's blocks resident while this one runs; a waiting tensor-core CTA (200 KB of smem,
// its TMEM columns) keeps the side stream's weight-gradient CTAs off that SM, and those are meant to run beside the
// elementwise passes.  So nothing triggers early: the gain is the launch latency + memory flush that the pre-staged
// dependent skips, with the prologue above pdl_wait() overlapping it.  The macros keep the other placements buildable
// (csrc/build.py --variant NAME -DSVK_PDL_TC_EARLY=1).
#ifndef SVK_PDL_EW_EARLY
#define SVK_PDL_EW_EARLY 0
#endif
#ifndef SVK_PDL_TC_EARLY
#define SVK_PDL_TC_EARLY 0
#endif
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {      // tensor-core kernels
#if SVK_PDL_TC_EARLY
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_prologue() {               // elementwise / reduction kernels: no prologue worth overlapping
#if SVK_PDL_EW_EARLY
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
  pdl_wait();
}

bool svk_pdl_enabled();

template <typename... KArgs, typename... Args>
static inline void svk_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = svk_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // the caller checks cudaGetLastError (SVK_LAUNCH_CHECK)
}

static inline int svk_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ----------------------------------------------------------------------------------------------------------
// Storage-type helpers: T is float or __nv_bfloat16; VEC elements make one 16-byte access.
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  typedef float4 raw;
  __device__ static inline void unpack(const float4& r, float (&v)[4]) { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
  __device__ static inline void load(const float* p, float (&v)[4]) { unpack(*reinterpret_cast<const float4*>(p), v); }
  __device__ static inline void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  typedef uint4 raw;
  __device__ static inline void unpack(const uint4& r, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ static inline void load(const __nv_bfloat16* p, float (&v)[8]) { unpack(*reinterpret_cast<const uint4*>(p), v); }
  __device__ static inline void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

__device__ inline float to_f(float v) { return v; }
__device__ inline float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ inline T from_f(float v);
template <> __device__ inline float from_f<float>(float v) { return v; }
template <> __device__ inline __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
// value as it will read back from storage of type T
template <typename T> __device__ inline float round_to(float v) { return to_f(from_f<T>(v)); }

__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#define SVK_DISPATCH_DTYPE(dtype, name, ...)                                   \
  if ((dtype) == SVK_F32) { typedef float T; __VA_ARGS__ }                     \
  else if ((dtype) == SVK_BF16) { typedef __nv_bfloat16 T; __VA_ARGS__ }       \
  else { svk_set_error("%s: bad dtype %d", name, (int)(dtype)); return SVK_E_BADARG; }

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
