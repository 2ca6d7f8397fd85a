/*
 * pystella_b200 -- run-time specialisation (EXPERIMENTAL, opt-in: PSB_JIT=1).
 *
 * pystella turns arbitrary symbolic statements into kernels with loopy
 * (pystella/elementwise.py:164-235).  The hand-written menu kernels of
 * pystella_b200.h cover the sectors the reference ships; for everything else the
 * Python layer (pystella_b200/codegen.py) emits one small CUDA C kernel per
 * ElementWiseMap, which this library compiles with NVRTC for sm_100a and launches
 * through the driver API.  Kept in a separate shared object so that the main
 * library has no NVRTC dependency.
 */
#ifndef PYSTELLA_B200_JIT_H
#define PYSTELLA_B200_JIT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSB_JIT_MAX_ARRAYS 48
#define PSB_JIT_MAX_SCALARS 32

/* The single by-value parameter of every generated kernel. */
typedef struct psb_jit_args {
  void* arrays[PSB_JIT_MAX_ARRAYS];     /* device pointers (outer subscripts applied)  */
  int32_t halos[PSB_JIT_MAX_ARRAYS];    /* halo padding of each array                  */
  double scalars[PSB_JIT_MAX_SCALARS];  /* by-value scalars, cast to the kernel dtype  */
  int64_t Nx, Ny, Nz;                   /* rank (un-padded) shape                      */
} psb_jit_args;

/* NVRTC-compile `source` (must define `extern "C" __global__ void <kernel_name>(const
 * psb_jit_args)`) for `arch` (e.g. "sm_100a").  No GPU needed.  Returns a handle > 0, or 0
 * on failure (psb_jit_last_log() holds the compiler log / error). */
int64_t psb_jit_compile(const char* source, const char* kernel_name, const char* arch);
const char* psb_jit_last_log(void);
/* Size of the compiled cubin (0 for a bad handle). */
size_t psb_jit_cubin_size(int64_t handle);
/* Launch (loads the module on first use).  Returns 0 on success. */
int psb_jit_launch(int64_t handle, const psb_jit_args* args, unsigned gx, unsigned gy,
                   unsigned gz, unsigned bx, unsigned by, unsigned bz, void* stream);
int psb_jit_release(int64_t handle);

/* General form used for the generated-RHS instantiation of the fused stage pipeline
 * (pystella_b200/csrc/fused_stage_ws.cuh compiled with PSB_GENERATED_RHS):
 *  - `options`: NVRTC options (include paths, defines); the architecture is appended;
 *  - `header_names` / `header_sources`: in-memory headers (stand-ins for the host system
 *    headers the device sources include);
 *  - `name_expression`: a C++ name expression of a __global__ template instance (its
 *    lowered name is looked up), or NULL when `kernel_name` is an extern "C" symbol. */
int64_t psb_jit_compile_ex(const char* source, const char* kernel_name,
                           const char* name_expression, const char* arch,
                           const char* const* options, int noptions,
                           const char* const* header_names,
                           const char* const* header_sources, int nheaders);
/* Register an already compiled cubin (e.g. from the on-disk cache the Python layer keeps,
 * written with psb_jit_write_cubin) whose kernel has the lowered name `symbol`. */
int64_t psb_jit_load_cubin(const void* cubin, size_t size, const char* symbol);
/* Lowered (mangled) symbol name of the compiled kernel. */
const char* psb_jit_kernel_symbol(int64_t handle);
/* Number of registers per thread / static shared memory of the compiled kernel are not
 * available without a device; the cubin can be inspected with cuobjdump instead. */
int psb_jit_write_cubin(int64_t handle, const char* path);
/* Launch with an arbitrary parameter list: params[i] points to the i-th kernel parameter
 * (cuLaunchKernel semantics).  `smem_bytes` of dynamic shared memory are opted into on
 * first use. */
int psb_jit_launch_params(int64_t handle, void** params, unsigned gx, unsigned gy,
                          unsigned gz, unsigned bx, unsigned by, unsigned bz,
                          unsigned smem_bytes, void* stream);
/* Same, with launch flags (PSB_LAUNCH_PDL of pystella_b200.h). */
int psb_jit_launch_params_v(int64_t handle, void** params, unsigned gx, unsigned gy,
                            unsigned gz, unsigned bx, unsigned by, unsigned bz,
                            unsigned smem_bytes, int launch_flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif
