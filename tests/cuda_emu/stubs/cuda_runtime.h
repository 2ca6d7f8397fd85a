// Test infrastructure: the few host-side CUDA runtime names the stand-alone translation units
// of pystella_b200/csrc use, so that they compile with g++ on top of cuda_emu.h.
#pragma once
#include <stdint.h>
#include <stdio.h>
typedef struct CUstream_st* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 3; return cudaSuccess; }
// programmatic-launch plumbing of common.cuh (template, never instantiated under emulation)
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 1 };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t {
  dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream;
  cudaLaunchAttribute* attrs; unsigned numAttrs;
};
template <typename... A, typename... B>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kern)(A...),
                                             B... args) {
  emu::launch_smem(cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes,
                   [&] { kern(static_cast<A>(args)...); });
  return cudaSuccess;
}
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// driver entry points: the one the library asks for, served by the tensor-map model
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0,
                                       cudaDriverEntryPointSymbolNotFound = 1 };
#define cudaEnableDefault 0
static inline cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long,
                                                  cudaDriverEntryPointQueryResult* q) {
  if (std::strcmp(name, "cuTensorMapEncodeTiled") == 0) {
    *fn = reinterpret_cast<void*>(&emu::encode_tiled);
    *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
  }
  *fn = nullptr;
  *q = cudaDriverEntryPointSymbolNotFound;
  return 1;
}
