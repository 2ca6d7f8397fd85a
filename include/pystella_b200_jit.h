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

#ifdef __cplusplus
}
#endif
#endif
