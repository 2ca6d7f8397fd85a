/*
 * pystella_b200 -- C ABI of the B200-native lattice-PDE hot path.
 *
 * This is the drop-in boundary underneath the Python classes that mirror
 * zachjweiner/pystella (Stepper / FiniteDifferencer / DomainDecomposition /
 * Reduction).  pystella itself has no FFI layer (every kernel is generated at
 * run time by loopy -> OpenCL), so each entry point cites the reference
 * *Python* call site whose device work it replaces (paths relative to the
 * reference checkout).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise;
 *   - arrays are C-ordered with z fastest; "padded" arrays have shape
 *     (Nx+2h, Ny+2h, Nz+2h), "unpadded" arrays (Nx, Ny, Nz)
 *     (pystella/field/__init__.py:579-606);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - functions return 0 on success; on failure a non-zero code is returned
 *     and psb_last_error() describes it (thread-local string);
 *   - nothing here allocates device memory: workspaces are passed in.
 */
#ifndef PYSTELLA_B200_H
#define PYSTELLA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSB_ABI_VERSION 2

#define PSB_F64 0
#define PSB_F32 1

#define PSB_MAX_SCALARS 4      /* fields per ScalarSector handled by one launch   */
#define PSB_MAX_TERMS 16       /* monomials of dV/df_n per field                  */
#define PSB_ENERGY_STRIDE 16    /* doubles per energy-partials row (>= 2*MAX_SCALARS+1) */
#define PSB_ENERGY_SUMS_DOUBLES (16 + 64 * 16 + 2)
#define PSB_MAX_RED_ARRAYS 8   /* distinct input arrays of one reduction launch   */
#define PSB_MAX_RED_TERMS 32   /* monomials over all outputs of one reduction     */
#define PSB_MAX_RED_OUTPUTS 16 /* reductions per launch                           */

const char* psb_last_error(void);
int psb_abi_version(void);
/* Compiled with FMA contraction disabled (bit-comparable with the oracle)? */
int psb_strict_fp(void);

/* ------------------------------------------------------------------------
 * Peer tables of one node (symmetric memory, NVLink): see psb_expansion_stage_peer.
 * ---------------------------------------------------------------------- */
typedef struct psb_peer_args {
  int32_t nranks, rank;
  void* const* peer_slots;       /* device array of nranks pointers                  */
  void* const* peer_flags;       /* device array of nranks pointers                  */
  uint64_t* epoch;               /* device counter (this rank)                       */
} psb_peer_args;

/* ------------------------------------------------------------------------
 * In-kernel stage tail (optional, fused stage kernels): CTA 0 of the launch (the "watcher")
 * performs what the reference does on the host between two stages --
 * Reduction.__call__ -> get_rho_and_p -> Expansion.step (pystella/reduction.py:237-255,
 * pystella/sectors.py:211-229, pystella/expansion.py:140-157) and, multi-rank, the
 * allreduce of pystella/decomp.py:470-491 -- with no extra launch and no fence or atomic in
 * the other CTAs: every CTA publishes one row of PSB_ENERGY_STRIDE doubles when it is done
 * (its energy partial sums -- the energy_partials rows of the launch -- or, for kernels
 * without energy sums, a marker in `rows`); rows must hold the bit pattern
 * PSB_TAIL_SENTINEL in every double before the launch (the watcher re-arms them), a value
 * is valid when it differs from it.  The watcher adds the rows in a fixed order and then
 *   PSB_TAIL_FOLD    writes the sums to sums[0..2ns];
 *   PSB_TAIL_PEER    all-gathers sums[0..2ns] over peer memory + inter-GPU flag barrier
 *                    (exactly psb_expansion_stage_peer's protocol; with nscalars == 0 a
 *                    pure barrier that orders the peer halo stores of the kernel);
 *   PSB_TAIL_EXPAND  energies with the current scale factor and one low-storage stage of
 *                    (a, adot) as psb_expansion_stage.
 * Without PSB_TAIL_FOLD, sums must have been folded by an earlier kernel of the stage.
 * ---------------------------------------------------------------------- */
/* Launch flags of the fused stage kernels and of psb_expansion_stage_peer_v.
 * PSB_LAUNCH_PDL: programmatic dependent launch -- the kernel may be scheduled while the
 * previous kernel of the stream drains (its CTAs set up their barriers and descriptors, then
 * wait in `griddepcontrol.wait` until the previous kernel has completed and its stores are
 * visible).  Every kernel of this library that can precede such a launch executes
 * `griddepcontrol.launch_dependents` and `griddepcontrol.wait` itself, so chains of stage
 * kernels stay transitively ordered.  Capturable in CUDA graphs (programmatic edges). */
#define PSB_LAUNCH_PDL 1

#define PSB_TAIL_FOLD 1
#define PSB_TAIL_EXPAND 2
#define PSB_TAIL_PEER 4
#define PSB_TAIL_SENTINEL 0x7ff8dead0000beefULL
typedef struct psb_stage_tail {
  int32_t mode;                  /* OR of PSB_TAIL_*; 0 = none                        */
  int32_t nscalars;              /* fields in the energy sums (0: no sums)            */
  double* rows;                  /* marker rows, one per CTA (kernels without energy
                                    partials); sentinel-initialised                   */
  int64_t nrows;                 /* capacity of `rows` in rows                        */
  double* sums;                  /* PSB_ENERGY_STRIDE doubles                         */
  double grid_size, mpl, dt, A, B;   /* EXPAND: background stage coefficients         */
  double* state;                 /* {a, adot, hubble, ka, kadot}                      */
  double* energies;              /* 2 nscalars + 3 doubles                            */
  psb_peer_args peers;           /* PEER                                              */
} psb_stage_tail;

/* ------------------------------------------------------------------------
 * Scalar-sector Runge-Kutta stage.
 *
 * One launch performs, for every field n < nscalars and every site of the
 * x-range [x_begin, x_end) x [0,Ny) x [0,Nz):
 *
 *     rhs_f    = dfdt_n
 *     rhs_dfdt = lap(f_n) - 2*hubble*dfdt_n - a^2 * dV/df_n(f_0..f_{n-1})
 *     kf_n    <- A*kf_n    + dt*rhs_f    ;  f_n    <- f_n    + B*kf_n
 *     kdfdt_n <- A*kdfdt_n + dt*rhs_dfdt ;  dfdt_n <- dfdt_n + B*kdfdt_n
 *
 * i.e. pystella.ScalarSector.rhs_dict (pystella/sectors.py:117-131) advanced by
 * one stage of pystella.LowStorageRKStepper (pystella/step.py:473-491, coefficients
 * step.py:543-557); the wave equation of examples/wave_equation.py:49-53 is the
 * case nscalars=1, a=1, hubble=0, no potential.  dV/df_n is a polynomial
 * sum_t coef[n][t] * prod_m f_m^power[n][t][m].
 *
 * psb_scalar_stage_fused evaluates lap(f_n) in-kernel from `f_in` (which must
 * hold valid halos on the faces) with the centred stencil of radius h
 * (pystella/derivs.py:160-165, 291-296), writes the new f to `f_out` (a second
 * padded buffer: the stencil of neighbouring sites still needs the old f) and,
 * where wrap_* is set, also stores the periodic image of each boundary site
 * into the opposite face halo of f_out (replaces the P==1 branches of
 * DomainDecomposition.share_halos, pystella/decomp.py:387-395, 422-424, for
 * star stencils).  It replaces the reference's per-stage pair
 * FiniteDifferencer.__call__ (pystella/derivs.py:339-429) + Stepper.__call__
 * (pystella/step.py:158-170, 524-528).
 *
 * psb_scalar_stage_unfused reads lap(f_n) from `lap_f` (unpadded) and updates
 * f in place (f_out must equal f_in), exactly like the reference's
 * element-wise stage kernel.
 * ---------------------------------------------------------------------- */
typedef struct psb_scalar_stage_args {
  int32_t dtype;                 /* PSB_F64 | PSB_F32                                */
  int32_t nscalars;              /* 1..PSB_MAX_SCALARS                               */
  int32_t h;                     /* halo padding of f/dfdt/k == stencil radius, 1..4 */
  int32_t wrap_x, wrap_y, wrap_z;/* fused: write periodic images into f_out halos    */
  int64_t Nx, Ny, Nz;            /* rank (un-padded) shape                           */
  int64_t x_begin, x_end;        /* sites updated by this launch                     */
  int64_t x2_begin, x2_end;      /* fused: optional second x range of the same launch */
                                 /* (both boundary slabs of a rank in one grid)       */
  int64_t field_stride;          /* elements between fields in padded arrays         */
  int64_t lap_stride;            /* elements between fields in lap_f (unfused)       */
  const void* f_in;              /* padded, field 0                                  */
  void* f_out;                   /* padded, field 0 (fused: != f_in)                 */
  void* dfdt;                    /* padded, in place                                 */
  void* kf;                      /* padded, in place (the "_f_tmp" twin)             */
  void* kdfdt;                   /* padded, in place (the "_dfdt_tmp" twin)          */
  const void* lap_f;             /* unfused only: unpadded Laplacian                 */
  double dx[3];                  /* grid spacing (fused)                             */
  double dt, A, B;               /* stage coefficients                               */
  double a, hubble;              /* used when a_dev/hubble_dev are NULL              */
  const double* a_dev;           /* optional device scalars (stay on device)         */
  const double* hubble_dev;
  /* Fused compute + halo exchange over peer memory (x-slab decomposition, NVLink):
   * when non-NULL, the new f of the h lowest / highest site planes is ALSO stored
   * straight into the high / low x-halo of the neighbouring rank's f_out buffer
   * (peer-mapped device pointers to field 0 of that buffer).  Replaces the
   * pack -> D2H -> MPI.Sendrecv -> H2D -> unpack of pystella/decomp.py:402-419. */
  void* peer_lo_f_out;           /* rank rx-1 (its HIGH halo receives my low planes)  */
  void* peer_hi_f_out;           /* rank rx+1 (its LOW halo receives my high planes)  */
  int64_t peer_lo_Nx;            /* x extent of rank rx-1 (locates its high halo)     */
  int64_t peer_lo_stride;        /* field strides (elements) of the peers' buffers    */
  int64_t peer_hi_stride;
  int32_t nterms[PSB_MAX_SCALARS];
  double coef[PSB_MAX_SCALARS][PSB_MAX_TERMS];
  int8_t power[PSB_MAX_SCALARS][PSB_MAX_TERMS][PSB_MAX_SCALARS];
  /* Optional fused energy sums of the INPUT (pre-stage) state, fused kernel only
   * (ScalarSector.reducers, pystella/sectors.py:133-144, without the a-dependent
   * prefactors): per CTA  {sum dfdt_n^2 | n}, sum V(f), {sum -f_n*lap(f_n) | n}.
   * energy_partials (may be NULL) holds psb_scalar_stage_energy_slots() rows of
   * PSB_ENERGY_STRIDE doubles starting at row energy_slot_offset.  V is the
   * polynomial sum_t vcoef[t] * prod_m f_m^vpower[t][m]. */
  double* energy_partials;
  int64_t energy_slot_offset;
  int32_t nvterms;
  double vcoef[PSB_MAX_TERMS];
  int8_t vpower[PSB_MAX_TERMS][PSB_MAX_SCALARS];
  int32_t tune_xchunk;           /* 0 = default; planes marched per CTA              */
  int32_t tune_variant;          /* 0 = auto; see DESIGN.md (kernel variants)        */
  const psb_stage_tail* tail;    /* fused only; HOST pointer, NULL = no in-kernel tail */
  int32_t tune_prefetch;         /* fused: planes the producer prefetches into L2 ahead of
                                    its loads; 0 = default, < 0 = none                  */
  int32_t launch_flags;          /* OR of PSB_LAUNCH_*                                  */
  void* pws_sched;               /* fused, optional: 16 zeroed bytes of device memory (work
                                    scheduler of the persistent kernel variant, re-armed
                                    by the kernel); NULL = one-chunk-per-CTA kernels only */
} psb_scalar_stage_args;

int psb_scalar_stage_fused(const psb_scalar_stage_args* args, void* stream);
int psb_scalar_stage_unfused(const psb_scalar_stage_args* args, void* stream);

/* ------------------------------------------------------------------------
 * Classical Runge-Kutta stage (storage copies along a leading axis) of the scalar
 * sector, element-wise: RungeKuttaStepper.make_steps (pystella/step.py:202-239)
 * with the step_statements of RungeKutta4 / RK3 / RK2 (step.py:242-411).
 * f, dfdt: (ncopies, nscalars, padded); lap_f: (nscalars, N) computed by the caller
 * from copy `q_read` (0 at stage 0, else 1; step.py:231).  For both unknowns u:
 *     u[dst[s]] = sum_m coef[s][m] * u[src[s][m]] + dt_coef[s] * dt * rhs_u(copy q_read)
 * for s < nstatements, all right-hand sides and sources taken from the old state.
 * ---------------------------------------------------------------------- */
typedef struct psb_rk_stage_args {
  int32_t dtype, nscalars, h, ncopies;
  int64_t Nx, Ny, Nz;
  void* f;
  void* dfdt;
  const void* lap_f;
  int32_t q_read, nstatements;
  int32_t dst[2];
  int32_t ncomb[2];
  int32_t src[2][2];
  double coef[2][2];
  double dt_coef[2];
  double dt, a, hubble;
  int32_t nterms[PSB_MAX_SCALARS];
  double pcoef[PSB_MAX_SCALARS][PSB_MAX_TERMS];
  int8_t power[PSB_MAX_SCALARS][PSB_MAX_TERMS][PSB_MAX_SCALARS];
} psb_rk_stage_args;
int psb_scalar_rk_stage(const psb_rk_stage_args* args, void* stream);

/* ------------------------------------------------------------------------
 * Tensor-perturbation (gravitational-wave) sector stage, element-wise:
 * TensorPerturbationSector.rhs_dict (pystella/sectors.py:187-204) with the
 * trace-dropped scalar stress tensor (sectors.py:146-152) advanced by one
 * low-storage RK stage (step.py:473-491).  For the six packed components
 * c = tensor_index(i, j), i <= j (sectors.py:164-167):
 *     rhs_hij    = dhijdt_c
 *     rhs_dhijdt = lap_hij_c - 2*hubble*dhijdt_c
 *                  + source_coef * sum_n dfdx[n][i-1] * dfdx[n][j-1]
 * hij/dhijdt/k arrays are padded (6, Nx+2h, ..); lap_hij (6, N..) and
 * dfdx (nscalars, 3, N..) are unpadded and precomputed (FiniteDifferencer).
 * ---------------------------------------------------------------------- */
typedef struct psb_tensor_stage_args {
  int32_t dtype, nscalars, h, reserved;
  int64_t Nx, Ny, Nz;
  void* hij;
  void* dhijdt;
  void* khij;
  void* kdhijdt;
  const void* lap_hij;
  const void* dfdx;
  double dt, A, B, hubble, source_coef;
  const double* hubble_dev;
} psb_tensor_stage_args;
int psb_tensor_stage_unfused(const psb_tensor_stage_args* args, void* stream);

/* The same stage with lap(hij_c) and the gradients of the source fields evaluated
 * in-kernel from `hij_in` and `src` (the scalar fields' OLD values, padded, valid face
 * halos), in the TMA pipeline of psb_scalar_stage_fused; hij ping-pongs into `hij_out`
 * (write-through periodic halos where wrap_* is set).  This is the multi-array stencil of
 * pystella/sectors.py:187-204 without lap_hij / dfdx ever touching HBM: per site and
 * stage 6 x (read hij, dhijdt, 2 twins; write the same) = 48 S bytes + the source planes
 * (2 S, shared by the six components through L2).  x-slab decompositions: the new hij of
 * the h lowest / highest site planes is also stored into the neighbours' hij_out halos
 * over peer memory (peer_*; as psb_scalar_stage_args), and `tail` lets the launch's last
 * CTA run the inter-GPU barrier + background stage (psb_stage_tail; no PSB_TAIL_FOLD). */
typedef struct psb_tensor_fused_args {
  int32_t dtype, nscalars, h;
  int32_t wrap_x, wrap_y, wrap_z;
  int64_t Nx, Ny, Nz;
  const void* hij_in;
  void* hij_out;
  void* dhijdt;
  void* khij;
  void* kdhijdt;
  const void* src;               /* (nscalars, padded) source fields, field 0        */
  int64_t src_stride;            /* elements between source fields                   */
  double dx[3];
  double dt, A, B, hubble, source_coef;
  const double* hubble_dev;
  void* peer_lo_hij_out;         /* rank rx-1's hij_out, component 0 (or NULL)        */
  void* peer_hi_hij_out;         /* rank rx+1's hij_out, component 0 (or NULL)        */
  int64_t peer_lo_Nx;            /* x extent of rank rx-1                             */
  int64_t peer_lo_stride;        /* component strides (elements) of the peers' buffers */
  int64_t peer_hi_stride;
  const psb_stage_tail* tail;    /* HOST pointer, NULL = none                         */
  int32_t tune_xchunk;           /* 0 = default                                       */
  int32_t tune_variant;          /* 0 = auto; 1 = 8-byte vectors for fp32             */
  int32_t tune_prefetch;         /* L2 prefetch distance in planes; 0 = default, < 0 none */
  int32_t launch_flags;          /* OR of PSB_LAUNCH_*                                  */
} psb_tensor_fused_args;
int psb_tensor_stage_fused(const psb_tensor_fused_args* args, void* stream);
/* CTAs of that launch (= marker rows its stage tail needs); -1 on error. */
int64_t psb_tensor_stage_fused_rows(const psb_tensor_fused_args* args);

/* ------------------------------------------------------------------------
 * Finite differences: FiniteDifferencer.__call__ kernels lap_knl / grad_knl /
 * grad_lap_knl / pd{x,y,z}_knl / pd*_incr_knl (pystella/derivs.py:291-337).
 * `fx` is padded by `halo` >= h with valid halos; any of lap/pdx/pdy/pdz may be
 * NULL.  With accumulate != 0 the first derivatives are added to the output
 * (the `div += pd_i` kernels of FiniteDifferencer.divergence, derivs.py:462-468).
 * ---------------------------------------------------------------------- */
int psb_finite_difference(const void* fx, void* lap, void* pdx, void* pdy, void* pdz,
                          int dtype, int64_t Nx, int64_t Ny, int64_t Nz,
                          int h, int halo, const double dx[3], int accumulate,
                          void* stream);
/* The same with an explicit kernel choice: variant 0 = auto (the TMA pipeline of
 * csrc/fd_ws.cuh for lap / grad / lap+grad with even radii and TMA-able layouts, else the
 * generic tile kernel), 1 = generic tile kernel. */
int psb_finite_difference_v(const void* fx, void* lap, void* pdx, void* pdy, void* pdz,
                            int dtype, int64_t Nx, int64_t Ny, int64_t Nz,
                            int h, int halo, const double dx[3], int accumulate,
                            int variant, void* stream);

/* ------------------------------------------------------------------------
 * Periodic halo fill of one rank-local padded array, order z -> x -> y, each
 * pass spanning the padded extent of the other axes (edges/corners included):
 * the pack_unpack_{z,x,y}_knl kernels (pystella/decomp.py:181-182, 208-209,
 * 232-233) called by DomainDecomposition.share_halos (decomp.py:387-395, 422-424).
 * axes_mask: bit0 = x, bit1 = y, bit2 = z.
 * ---------------------------------------------------------------------- */
int psb_halo_wrap(void* arr, int dtype, int64_t Nx, int64_t Ny, int64_t Nz,
                  int hx, int hy, int hz, int axes_mask, void* stream);

/* Face pack/unpack for decomposed axes (pack_{x,y}_knl / unpack_{x,y}_knl,
 * pystella/decomp.py:177-180, 204-207).  buf has shape (2, hx, Ny+2hy, Nz+2hz)
 * for axis 0 and (2, Nx+2hx, hy, Nz+2hz) for axis 1; slot 0 = low face /
 * high halo, slot 1 = high face / low halo, as in the reference. */
int psb_halo_pack(const void* arr, void* buf, int dtype, int64_t Nx, int64_t Ny,
                  int64_t Nz, int hx, int hy, int hz, int axis, void* stream);
int psb_halo_unpack(void* arr, const void* buf, int dtype, int64_t Nx, int64_t Ny,
                    int64_t Nz, int hx, int hy, int hz, int axis, void* stream);

/* gather_knl / scatter_knl (pystella/decomp.py:253-254): strip / add padding. */
int psb_remove_halos(const void* padded, void* unpadded, int dtype, int64_t Nx,
                     int64_t Ny, int64_t Nz, int hx, int hy, int hz, void* stream);
int psb_restore_halos(const void* unpadded, void* padded, int dtype, int64_t Nx,
                      int64_t Ny, int64_t Nz, int hx, int hy, int hz, void* stream);

/* ------------------------------------------------------------------------
 * Polynomial reductions: Reduction.__call__ (pystella/reduction.py:165-204,
 * 237-255).  Output r is  op_r over sites of  sum_{t in terms(r)} coef[t] *
 * prod_j arr_j^power[t][j]  (abs value taken first if flags bit0 is set), with
 * op in {0: sum, 1: max, 2: min}.  Array j has halo padding halo[j] (0 for
 * unpadded arrays).  Two launches: per-block partials, then a fixed-order
 * finalize into out[r] (double).
 * ---------------------------------------------------------------------- */
typedef struct psb_reduce_args {
  int32_t dtype;
  int32_t narrays, nterms, noutputs;
  int64_t Nx, Ny, Nz;
  const void* arrays[PSB_MAX_RED_ARRAYS];
  int32_t halo[PSB_MAX_RED_ARRAYS];
  double coef[PSB_MAX_RED_TERMS];
  int8_t power[PSB_MAX_RED_TERMS][PSB_MAX_RED_ARRAYS];
  int32_t term_output[PSB_MAX_RED_TERMS]; /* which output each term feeds */
  int32_t op[PSB_MAX_RED_OUTPUTS];
  int32_t flags[PSB_MAX_RED_OUTPUTS];
  double* partials;              /* workspace: psb_reduce_workspace_doubles() */
  double* out;                   /* noutputs doubles (device)                 */
  int32_t tune_variant;          /* 0 = auto; 1 = generic (scalar-load) kernel */
  int32_t reserved0;
} psb_reduce_args;

int64_t psb_reduce_workspace_doubles(int noutputs);
int psb_reduce(const psb_reduce_args* args, void* stream);

/* ------------------------------------------------------------------------
 * Binning pass of PowerSpectra.bin_power (pystella/fourier/spectra.py:103-131, a
 * Histogrammer over momentum space, pystella/histogram.py:60-170):
 *     spectrum[round(|k| / bin_width)] += count * |k|^k_power * |fk|^2
 * over the rank-local momentum-space array fk (Kx, Ky, Kz), C order, complex128
 * (dtype PSB_F64) or complex64 (PSB_F32);  |k|^2 = sum_mu (dk[mu] * momenta_mu[i_mu])^2
 * with the three 1-D momenta arrays of DFT.sub_k (real, same precision as fk);
 * count = 2 where is_real and 0 < momenta_z < nyquist, else 1.  `spectrum` (num_bins
 * doubles) is accumulated into: the caller zeroes it.  One pass over fk: 2 S bytes per
 * mode.  Accumulation order across warps is not fixed (atomics, as in the reference).
 * ---------------------------------------------------------------------- */
typedef struct psb_bin_power_args {
  int32_t dtype;
  int32_t is_real;
  int64_t Kx, Ky, Kz;
  const void* fk;
  const void* momenta_x;
  const void* momenta_y;
  const void* momenta_z;
  double dk[3];
  double bin_width, k_power, nyquist;
  int32_t num_bins;
  int32_t tune_blocks;           /* 0 = automatic (4 CTAs per SM) */
  double* spectrum;
} psb_bin_power_args;

int psb_bin_power(const psb_bin_power_args* args, void* stream);

/* ------------------------------------------------------------------------
 * Launch plan of the fused stage for an externally compiled kernel instance
 * (EXPERIMENTAL; row f-2 of the scope table): pystella accepts arbitrary right-hand
 * sides (pystella/step.py:473-491); for systems outside the polynomial menu the
 * Python layer compiles pystella_b200/csrc/fused_stage_ws.cuh at run time with a
 * generated RHS functor (NVRTC, include/pystella_b200_jit.h) and launches that instance
 * with exactly the parameters the built-in launch would use.  This entry computes them
 * without launching: the template arguments of the instance that matches the layout,
 * the launch geometry, the by-value kernel parameter block and the four TMA
 * descriptors (f_in, dfdt, kf, kdfdt).  Returns 3 when the layout is not eligible for
 * the TMA pipeline (odd extents / unaligned pointers).
 * ---------------------------------------------------------------------- */
#define PSB_FUSED_PLAN_PARAM_BYTES 4096
typedef struct psb_fused_plan {
  int32_t dtype, h, nscalars;          /* T, H, NF                                    */
  int32_t by, tz, p, minb, vec, energy;/* BY, TZ, P, MINB, VEC, ENERGY                */
  uint32_t grid[3], threads, smem_bytes, params_bytes;
  int32_t reserved;                    /* in: non-zero = dry run (geometry and parameter
                                          block only, descriptors left zero; needs no
                                          driver).  Keeps the descriptors at offset 64 */
  unsigned char tensor_maps[4][128];   /* CUtensorMap x 4                             */
  unsigned char params[PSB_FUSED_PLAN_PARAM_BYTES]; /* psb::FusedParams<T, NF>        */
} psb_fused_plan;

int psb_scalar_stage_fused_plan(const psb_scalar_stage_args* args, psb_fused_plan* plan);

/* Number of energy-partial rows one psb_scalar_stage_fused launch with these
 * args writes (one per CTA). */
int64_t psb_scalar_stage_energy_slots(const psb_scalar_stage_args* args);

/* ------------------------------------------------------------------------
 * Device-resident FLRW background (replaces, per RK stage, the host round trip
 * Reduction.__call__ -> Expansion.step of examples/scalar_preheating.py:263-264:
 * pystella/reduction.py:237-255, pystella/sectors.py:211-229,
 * pystella/expansion.py:101-157).
 *
 * psb_energy_sums folds `nslots` partial rows into sums[0..nvals) in a fixed
 * order (two-level, one launch).  `sums` must hold PSB_ENERGY_SUMS_DOUBLES doubles,
 * zero-initialised once by the caller (result + scratch rows + a ticket counter).  psb_expansion_stage (single thread) then forms, with the
 * CURRENT scale factor a = state[0]:
 *     kinetic_n  = sums[n]        / (2 a^2) / grid_size
 *     potential  = sums[ns]                 / grid_size
 *     gradient_n = sums[ns+1+n]   / (2 a^2) / grid_size
 *     total = sum of all ; pressure = sum kinetic - sum gradient / 3 - potential
 * writes energies[0..2ns+3) = {kinetic.., potential, gradient.., total, pressure}
 * and advances (a, adot) by one low-storage RK stage of
 *     a' = adot ; adot' = 4 pi a^2 / (3 mpl^2) (total - 3 pressure) a
 * with twins (ka, kadot), then hubble = adot / a.
 * state = {a, adot, hubble, ka, kadot} (5 doubles on the device).
 * Between the two calls a multi-rank job all-reduces sums[0..nvals).
 * ---------------------------------------------------------------------- */
int psb_energy_sums(const double* partials, int64_t nslots, int nvals, double* sums,
                    void* stream);
int psb_expansion_stage(const double* sums, int nscalars, double grid_size, double mpl,
                        double dt, double A, double B, double* state, double* energies,
                        void* stream);
/* Multi-rank variant without NCCL: every rank stores its sums into slot `rank` of
 * each peer's symmetric slot array (peer_slots[q] -> rank q's array of
 * [2][nranks][PSB_ENERGY_STRIDE] doubles), raises its flag there
 * (peer_flags[q] -> rank q's uint64[2][nranks]), waits for all flags, adds the
 * slots in rank order (bit-identical on all ranks) and then performs
 * psb_expansion_stage.  The kernel doubles as the inter-GPU barrier that orders
 * the peer-memory halo stores of psb_scalar_stage_fused before the next stage.
 * `epoch` is a device counter the kernel increments (CUDA-graph friendly).
 * With nscalars == 0 only the barrier is performed (no energies / expansion). */
int psb_expansion_stage_peer(const psb_peer_args* peers, const double* sums, int nscalars,
                             double grid_size, double mpl, double dt, double A, double B,
                             double* state, double* energies, void* stream);
/* Same, with launch flags (PSB_LAUNCH_PDL). */
int psb_expansion_stage_peer_v(const psb_peer_args* peers, const double* sums, int nscalars,
                               double grid_size, double mpl, double dt, double A, double B,
                               double* state, double* energies, int launch_flags,
                               void* stream);

/* adot = sqrt(8 pi a^2 rho / (3 mpl^2)) a from energies (first Friedmann equation,
 * pystella/expansion.py:101-117): initialises state = {a, adot, hubble, 0, 0}. */
int psb_expansion_init(const double* sums, int nscalars, double grid_size, double mpl,
                       double a0, double* state, double* energies, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PYSTELLA_B200_H */
