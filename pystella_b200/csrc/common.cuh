// Shared helpers for the pystella_b200 CUDA translation units (sm_100a only).
#pragma once

#ifndef __CUDACC_RTC__
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#else
// NVRTC (generated-RHS instantiation of the fused pipeline, see fused_stage_ws.cuh): no
// host headers; the few types the device code needs are spelled out here
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef struct CUtensorMap_st {
  alignas(64) unsigned long long opaque[16];
} CUtensorMap;
typedef struct CUstream_st* cudaStream_t;
#endif

#include "pystella_b200.h"

namespace psb {

void set_error(const char* fmt, ...);

#define PSB_CHECK_CUDA(expr)                                                   \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      psb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                     cudaGetErrorString(_e));                                  \
      return 2;                                                                \
    }                                                                          \
  } while (0)

#define PSB_REQUIRE(cond, ...)                                                 \
  do {                                                                         \
    if (!(cond)) {                                                             \
      psb::set_error(__VA_ARGS__);                                             \
      return 1;                                                                \
    }                                                                          \
  } while (0)

#ifndef __CUDACC_RTC__
// One launch path for the stage kernels: plain stream order, or (launch_flags &
// PSB_LAUNCH_PDL) with the programmatic-stream-serialization attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, int launch_flags, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (launch_flags & PSB_LAUNCH_PDL) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
#endif

// --------------------------------------------------------------------------
// finite-difference coefficient tables (pystella/derivs.py:127-131, 160-165),
// evaluated exactly like the reference's Python float divisions.
// --------------------------------------------------------------------------
template <int H> __host__ __device__ constexpr double lap_coef(int s) {
  if constexpr (H == 1) return s == 0 ? -2.0 : 1.0;
  if constexpr (H == 2) return s == 0 ? -30.0 / 12.0 : (s == 1 ? 16.0 / 12.0 : -1.0 / 12.0);
  if constexpr (H == 3)
    return s == 0 ? -490.0 / 180.0
                  : (s == 1 ? 270.0 / 180.0 : (s == 2 ? -27.0 / 180.0 : 2.0 / 180.0));
  if constexpr (H == 4)
    return s == 0 ? -14350.0 / 5040.0
                  : (s == 1 ? 8064.0 / 5040.0
                            : (s == 2 ? -1008.0 / 5040.0
                                      : (s == 3 ? 128.0 / 5040.0 : -9.0 / 5040.0)));
  return 0.0;
}
template <int H> __host__ __device__ constexpr double grad_coef(int s) {
  if constexpr (H == 1) return 1.0 / 2.0;
  if constexpr (H == 2) return s == 1 ? 8.0 / 12.0 : -1.0 / 12.0;
  if constexpr (H == 3) return s == 1 ? 45.0 / 60.0 : (s == 2 ? -9.0 / 60.0 : 1.0 / 60.0);
  if constexpr (H == 4)
    return s == 1 ? 672.0 / 840.0
                  : (s == 2 ? -168.0 / 840.0 : (s == 3 ? 32.0 / 840.0 : -3.0 / 840.0));
  return 0.0;
}

// --------------------------------------------------------------------------
// small vector type: VEC consecutive z-sites owned by one thread
// --------------------------------------------------------------------------
template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Vec {
  T v[VEC];
};

template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> ld_vec(const T* p) {
  return *reinterpret_cast<const Vec<T, VEC>*>(p);
}
template <typename T, int VEC>
__device__ __forceinline__ void st_vec(T* p, const Vec<T, VEC>& x) {
  *reinterpret_cast<Vec<T, VEC>*>(p) = x;
}

// --------------------------------------------------------------------------
// mbarrier + TMA (cp.async.bulk.tensor) wrappers
// --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// 4-D tiled load: coordinates innermost first (z, y, x, field)
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map,
                                            uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}

// L2 prefetch of a tile (no shared memory, no completion): the pipeline's producer issues it
// a few planes ahead of the real load, so that the load hits L2 (~3x shorter latency) and the
// bytes the rings can keep in flight go further.
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2,
                                                int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// Programmatic dependent launch (PSB_LAUNCH_PDL): `pdl_launch_dependents` lets the next
// kernel of the stream be scheduled once every CTA of this grid has executed it (i.e. once
// the whole grid has been dispatched); `pdl_wait` blocks until the previous kernel has
// completed and flushed.  Both are no-ops for launches without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map))
               : "memory");
}

// Host: encode a 4-D (field, x, y, z) tensor map over a stack of padded fields.
// Returns 0 on success, non-zero if the layout violates TMA constraints
// (caller then falls back to the cooperative loader).
int encode_field_tensor_map(CUtensorMap* map, int dtype, const void* base,
                            int64_t nfields, int64_t field_stride_elems, int64_t Px,
                            int64_t Py, int64_t Pz, int box_x, int box_y, int box_z);

}  // namespace psb
