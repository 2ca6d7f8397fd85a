// Test infrastructure: the driver-API names common.cu mentions (tensor-map encoding).  The
// entry-point query reports "not available", so every TMA descriptor request fails the way it
// does on a machine without the driver, and the callers take their non-TMA paths.
#pragma once
#include <stdint.h>
typedef struct CUtensorMap_st { alignas(64) unsigned long long opaque[16]; } CUtensorMap;
typedef int CUresult;
#define CUDA_SUCCESS 0
typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7, CU_TENSOR_MAP_DATA_TYPE_FLOAT64 = 8 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_256B = 3 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
