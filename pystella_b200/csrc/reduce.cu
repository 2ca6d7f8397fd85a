// Polynomial reductions: the kernel side of Reduction.__call__
// (pystella/reduction.py:165-204, 237-255).  The reference computes
// tmp[r,j,k] = sum_i expr_r and then runs one pyopencl reduction + one blocking
// D2H read per expression; here every expression of the call is reduced in one
// pass (warp shuffle -> block -> per-block partials) and a single-block finalize
// kernel folds the partials in a fixed order (deterministic for a given grid).
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace psb {

constexpr int RED_MAX_BLOCKS = 1024;
constexpr int RED_THREADS = 256;

template <typename T> struct RedParams {
  const T* arrays[PSB_MAX_RED_ARRAYS];
  int halo[PSB_MAX_RED_ARRAYS];
  int narrays, nterms, noutputs;
  int Nx, Ny, Nz;
  double coef[PSB_MAX_RED_TERMS];
  int8_t power[PSB_MAX_RED_TERMS][PSB_MAX_RED_ARRAYS];
  int term_begin[PSB_MAX_RED_OUTPUTS + 1];  // terms sorted by output
  int op[PSB_MAX_RED_OUTPUTS];
  int flags[PSB_MAX_RED_OUTPUTS];
  double* partials;
  double* out;
  int variant;   // 0 = auto (vectorised kernel when the layout allows), 1 = generic kernel
};

__device__ __forceinline__ double red_identity(int op) {
  return op == 0 ? 0.0 : (op == 1 ? -DBL_MAX : DBL_MAX);
}
__device__ __forceinline__ double red_combine(int op, double a, double b) {
  return op == 0 ? a + b : (op == 1 ? fmax(a, b) : fmin(a, b));
}

// NARR / NOUT are compile-time bounds (rounded up by the host dispatch) so that the
// per-site loops are short and fully unrolled; term ranges and powers stay runtime
// (uniform, read from the kernel parameter space).
template <typename T, int NARR, int NOUT>
__global__ void __launch_bounds__(RED_THREADS)
    reduce_kernel(const __grid_constant__ RedParams<T> p) {
  double acc[NOUT];
#pragma unroll
  for (int r = 0; r < NOUT; ++r) acc[r] = red_identity(r < p.noutputs ? p.op[r] : 0);

  const int64_t rows = (int64_t)p.Nx * p.Ny;
  const int lane_z = threadIdx.x % 64, sub = threadIdx.x / 64;
  for (int64_t row = (int64_t)blockIdx.x * 4 + sub; row < rows;
       row += (int64_t)gridDim.x * 4) {
    const int i = (int)(row / p.Ny), j = (int)(row % p.Ny);
    const T* rp[NARR];
#pragma unroll
    for (int m = 0; m < NARR; ++m) {
      const int h = p.halo[m];
      rp[m] = (m < p.narrays)
                  ? p.arrays[m] + ((int64_t)(i + h) * (p.Ny + 2 * h) + (j + h)) *
                                          (p.Nz + 2 * h) + h
                  : p.arrays[0];
    }
    for (int k = lane_z; k < p.Nz; k += 64) {
      double v[NARR];
#pragma unroll
      for (int m = 0; m < NARR; ++m) v[m] = (double)__ldg(rp[m] + k);
#pragma unroll
      for (int r = 0; r < NOUT; ++r) {
        if (r < p.noutputs) {
          double val = 0.0;
          for (int t = p.term_begin[r]; t < p.term_begin[r + 1]; ++t) {
            double term = p.coef[t];
#pragma unroll
            for (int m = 0; m < NARR; ++m) {
              const int pw = p.power[t][m];
              if (pw >= 1) term *= v[m];
              if (pw >= 2) term *= v[m];
              for (int e = 2; e < pw; ++e) term *= v[m];
            }
            val += term;
          }
          if (p.flags[r] & 1) val = fabs(val);
          const int op = p.op[r];
          acc[r] = op == 0 ? acc[r] + val : red_combine(op, acc[r], val);
        }
      }
    }
  }

  __shared__ double sh[RED_THREADS / 32][NOUT];
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
#pragma unroll
  for (int r = 0; r < NOUT; ++r) {
    if (r < p.noutputs) {
      double x = acc[r];
      const int op = p.op[r];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1)
        x = red_combine(op, x, __shfl_xor_sync(0xffffffffu, x, d));
      if (lane == 0) sh[warp][r] = x;
    }
  }
  __syncthreads();
  if (threadIdx.x < p.noutputs) {
    const int r = threadIdx.x, op = p.op[r];
    double x = sh[0][r];
    for (int w = 1; w < RED_THREADS / 32; ++w) x = red_combine(op, x, sh[w][r]);
    p.partials[(int64_t)blockIdx.x * PSB_MAX_RED_OUTPUTS + r] = x;
  }
}

// Vectorised variant (default when the layout allows): every thread owns U row segments of
// two sites each (one 16-byte / 8-byte vector load per array and segment, all issued before
// the arithmetic starts), and the run-time loops over terms and powers are OUTSIDE the loop
// over those 2U sites, so their overhead (uniform loads of coefficients and powers, branches)
// is paid once per 2U sites.  Per site the association is the oracle's:
// term = ((coef * a) * a) ..., value = value + term in term order.  Needs even Nz and halos
// and 2-element aligned rows; otherwise reduce_kernel.
constexpr int RED_V2_UNROLL = 4;

template <typename T, int NARR, int NOUT>
__global__ void __launch_bounds__(RED_THREADS)
    reduce_kernel_v2(const __grid_constant__ RedParams<T> p) {
  double acc[NOUT];
#pragma unroll
  for (int r = 0; r < NOUT; ++r) acc[r] = red_identity(r < p.noutputs ? p.op[r] : 0);

  const int64_t rows = (int64_t)p.Nx * p.Ny;
  const int lane_z = threadIdx.x % 64, sub = threadIdx.x / 64;
  constexpr int U = RED_V2_UNROLL, S = 2 * U;
  for (int64_t row = (int64_t)blockIdx.x * 4 + sub; row < rows;
       row += (int64_t)gridDim.x * 4) {
    const int i = (int)(row / p.Ny), j = (int)(row % p.Ny);
    const T* rp[NARR];
#pragma unroll
    for (int m = 0; m < NARR; ++m) {
      const int h = p.halo[m];
      rp[m] = (m < p.narrays)
                  ? p.arrays[m] + ((int64_t)(i + h) * (p.Ny + 2 * h) + (j + h)) *
                                          (p.Nz + 2 * h) + h
                  : p.arrays[0];
    }
    for (int k0 = lane_z * 2; k0 < p.Nz; k0 += 64 * 2 * U) {
      double v[NARR][S];
      bool live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = k0 + u * 128;
        live[u] = k < p.Nz;
#pragma unroll
        for (int m = 0; m < NARR; ++m) {
          Vec<T, 2> raw;
          if (live[u]) {
            raw = ld_vec<T, 2>(rp[m] + k);
          } else {
            raw.v[0] = raw.v[1] = (T)0;
          }
          v[m][2 * u] = (double)raw.v[0];
          v[m][2 * u + 1] = (double)raw.v[1];
        }
      }
#pragma unroll
      for (int r = 0; r < NOUT; ++r) {
        if (r < p.noutputs) {
          double val[S];
#pragma unroll
          for (int s = 0; s < S; ++s) val[s] = 0.0;
          for (int t = p.term_begin[r]; t < p.term_begin[r + 1]; ++t) {
            double term[S];
            const double c = p.coef[t];
#pragma unroll
            for (int s = 0; s < S; ++s) term[s] = c;
#pragma unroll
            for (int m = 0; m < NARR; ++m) {
              const int pw = p.power[t][m];
              for (int e = 0; e < pw; ++e) {
#pragma unroll
                for (int s = 0; s < S; ++s) term[s] *= v[m][s];
              }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) val[s] += term[s];
          }
          const bool take_abs = (p.flags[r] & 1) != 0;
          const int op = p.op[r];
          if (op == 0) {
            double part = 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (live[u]) {
                const double a0 = take_abs ? fabs(val[2 * u]) : val[2 * u];
                const double a1 = take_abs ? fabs(val[2 * u + 1]) : val[2 * u + 1];
                part += a0 + a1;
              }
            }
            acc[r] += part;
          } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (live[u]) {
                const double a0 = take_abs ? fabs(val[2 * u]) : val[2 * u];
                const double a1 = take_abs ? fabs(val[2 * u + 1]) : val[2 * u + 1];
                acc[r] = red_combine(op, acc[r], red_combine(op, a0, a1));
              }
            }
          }
        }
      }
    }
  }

  __shared__ double sh[RED_THREADS / 32][NOUT];
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
#pragma unroll
  for (int r = 0; r < NOUT; ++r) {
    if (r < p.noutputs) {
      double x = acc[r];
      const int op = p.op[r];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1)
        x = red_combine(op, x, __shfl_xor_sync(0xffffffffu, x, d));
      if (lane == 0) sh[warp][r] = x;
    }
  }
  __syncthreads();
  if (threadIdx.x < p.noutputs) {
    const int r = threadIdx.x, op = p.op[r];
    double x = sh[0][r];
    for (int w = 1; w < RED_THREADS / 32; ++w) x = red_combine(op, x, sh[w][r]);
    p.partials[(int64_t)blockIdx.x * PSB_MAX_RED_OUTPUTS + r] = x;
  }
}

template <typename T> static bool reduce_v2_layout_ok(const RedParams<T>& p) {
  if (p.Nz % 2) return false;
  for (int m = 0; m < p.narrays; ++m) {
    if (p.halo[m] % 2) return false;
    if (reinterpret_cast<uintptr_t>(p.arrays[m]) % (2 * sizeof(T))) return false;
  }
  return true;
}

template <typename T, int NARR>
static void launch_reduce_nout(const RedParams<T>& p, int nblocks, cudaStream_t s) {
  if constexpr (NARR <= 4) {
    if (p.variant != 1 && reduce_v2_layout_ok(p)) {
      if (p.noutputs <= 2)
        reduce_kernel_v2<T, NARR, 2><<<nblocks, RED_THREADS, 0, s>>>(p);
      else if (p.noutputs <= 4)
        reduce_kernel_v2<T, NARR, 4><<<nblocks, RED_THREADS, 0, s>>>(p);
      else if (p.noutputs <= 8)
        reduce_kernel_v2<T, NARR, 8><<<nblocks, RED_THREADS, 0, s>>>(p);
      else
        reduce_kernel_v2<T, NARR, PSB_MAX_RED_OUTPUTS><<<nblocks, RED_THREADS, 0, s>>>(p);
      return;
    }
  }
  if (p.noutputs <= 2)
    reduce_kernel<T, NARR, 2><<<nblocks, RED_THREADS, 0, s>>>(p);
  else if (p.noutputs <= 4)
    reduce_kernel<T, NARR, 4><<<nblocks, RED_THREADS, 0, s>>>(p);
  else if (p.noutputs <= 8)
    reduce_kernel<T, NARR, 8><<<nblocks, RED_THREADS, 0, s>>>(p);
  else
    reduce_kernel<T, NARR, PSB_MAX_RED_OUTPUTS><<<nblocks, RED_THREADS, 0, s>>>(p);
}

template <typename T>
static void launch_reduce(const RedParams<T>& p, int nblocks, cudaStream_t s) {
  if (p.narrays <= 1)
    launch_reduce_nout<T, 1>(p, nblocks, s);
  else if (p.narrays <= 2)
    launch_reduce_nout<T, 2>(p, nblocks, s);
  else if (p.narrays <= 3)
    launch_reduce_nout<T, 3>(p, nblocks, s);
  else if (p.narrays <= 4)
    launch_reduce_nout<T, 4>(p, nblocks, s);
  else if (p.narrays <= 6)
    launch_reduce_nout<T, 6>(p, nblocks, s);
  else
    launch_reduce_nout<T, PSB_MAX_RED_ARRAYS>(p, nblocks, s);
}

struct FinalizeParams {
  const double* partials;
  double* out;
  int nblocks, noutputs;
  int op[PSB_MAX_RED_OUTPUTS];
};

__global__ void __launch_bounds__(256)
    finalize_kernel(const __grid_constant__ FinalizeParams p) {
  __shared__ double sh[8];
  for (int r = 0; r < p.noutputs; ++r) {
    const int op = p.op[r];
    double x = red_identity(op);
    for (int b = threadIdx.x; b < p.nblocks; b += blockDim.x)
      x = red_combine(op, x, p.partials[(int64_t)b * PSB_MAX_RED_OUTPUTS + r]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
      x = red_combine(op, x, __shfl_xor_sync(0xffffffffu, x, d));
    if (threadIdx.x % 32 == 0) sh[threadIdx.x / 32] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      double y = sh[0];
      for (int w = 1; w < 8; ++w) y = red_combine(op, y, sh[w]);
      p.out[r] = y;
    }
    __syncthreads();
  }
}

template <typename T>
static int reduce_typed(const psb_reduce_args* a, cudaStream_t s) {
  RedParams<T> p;
  memset(&p, 0, sizeof(p));
  p.narrays = a->narrays;
  p.noutputs = a->noutputs;
  p.Nx = (int)a->Nx;
  p.Ny = (int)a->Ny;
  p.Nz = (int)a->Nz;
  for (int m = 0; m < a->narrays; ++m) {
    p.arrays[m] = static_cast<const T*>(a->arrays[m]);
    p.halo[m] = a->halo[m];
  }
  // sort terms by output (stable), build ranges
  int nt = 0;
  for (int r = 0; r < a->noutputs; ++r) {
    p.term_begin[r] = nt;
    for (int t = 0; t < a->nterms; ++t) {
      if (a->term_output[t] == r) {
        p.coef[nt] = a->coef[t];
        for (int m = 0; m < PSB_MAX_RED_ARRAYS; ++m) p.power[nt][m] = a->power[t][m];
        ++nt;
      }
    }
    p.op[r] = a->op[r];
    p.flags[r] = a->flags[r];
  }
  p.term_begin[a->noutputs] = nt;
  p.nterms = nt;
  p.partials = a->partials;
  p.out = a->out;
  p.variant = a->tune_variant;

  const int64_t rows = a->Nx * a->Ny;
  int nblocks = (int)((rows + 3) / 4);
  const int cap = num_sms() * 6 < RED_MAX_BLOCKS ? num_sms() * 6 : RED_MAX_BLOCKS;
  if (nblocks > cap) nblocks = cap;
  launch_reduce<T>(p, nblocks, s);
  PSB_CHECK_CUDA(cudaGetLastError());

  FinalizeParams f;
  f.partials = a->partials;
  f.out = a->out;
  f.nblocks = nblocks;
  f.noutputs = a->noutputs;
  for (int r = 0; r < PSB_MAX_RED_OUTPUTS; ++r) f.op[r] = r < a->noutputs ? a->op[r] : 0;
  finalize_kernel<<<1, 256, 0, s>>>(f);
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace psb

extern "C" int64_t psb_reduce_workspace_doubles(int noutputs) {
  (void)noutputs;
  return (int64_t)psb::RED_MAX_BLOCKS * PSB_MAX_RED_OUTPUTS;
}

extern "C" int psb_reduce(const psb_reduce_args* a, void* stream) {
  using namespace psb;
  PSB_REQUIRE(a != nullptr, "psb_reduce: null args");
  PSB_REQUIRE(a->narrays >= 1 && a->narrays <= PSB_MAX_RED_ARRAYS,
              "psb_reduce: narrays=%d out of range", a->narrays);
  PSB_REQUIRE(a->nterms >= 0 && a->nterms <= PSB_MAX_RED_TERMS,
              "psb_reduce: nterms=%d out of range", a->nterms);
  PSB_REQUIRE(a->noutputs >= 1 && a->noutputs <= PSB_MAX_RED_OUTPUTS,
              "psb_reduce: noutputs=%d out of range", a->noutputs);
  PSB_REQUIRE(a->Nx > 0 && a->Ny > 0 && a->Nz > 0, "psb_reduce: empty grid");
  PSB_REQUIRE(a->partials && a->out, "psb_reduce: null workspace/output");
  for (int m = 0; m < a->narrays; ++m)
    PSB_REQUIRE(a->arrays[m] != nullptr, "psb_reduce: null array %d", m);
  for (int r = 0; r < a->noutputs; ++r)
    PSB_REQUIRE(a->op[r] >= 0 && a->op[r] <= 2, "psb_reduce: unknown op %d", a->op[r]);
  for (int t = 0; t < a->nterms; ++t) {
    PSB_REQUIRE(a->term_output[t] >= 0 && a->term_output[t] < a->noutputs,
                "psb_reduce: term %d feeds unknown output", t);
    for (int m = 0; m < a->narrays; ++m)
      PSB_REQUIRE(a->power[t][m] >= 0, "psb_reduce: negative power of array %d in term %d", m,
                  t);
  }
  for (int m = 0; m < a->narrays; ++m)
    PSB_REQUIRE(a->halo[m] >= 0, "psb_reduce: negative halo for array %d", m);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == PSB_F64) return reduce_typed<double>(a, s);
  if (a->dtype == PSB_F32) return reduce_typed<float>(a, s);
  set_error("psb_reduce: unknown dtype %d", a->dtype);
  return 1;
}
