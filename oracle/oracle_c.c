/*
 * TEST INFRASTRUCTURE ONLY -- plain-C/OpenMP restatement of the reference hot
 * path, used (i) as a second, independently written checker against the NumPy
 * oracle and (ii) as the timed CPU baseline of bench.py ("cpu_baseline" and
 * "--impl reference"; kind = "port": the reference's own pyopencl-on-POCL path
 * cannot run in this image).  It follows the reference's UNFUSED
 * per-stage schedule (examples/scalar_preheating.py:258-266):
 *
 *   share_halos(f)            pystella/decomp.py:387-395, 422-424
 *   lap_f = lap(f)            pystella/derivs.py:160-165, 291-296
 *   LSRK stage on (f, dfdt)   pystella/step.py:473-491, sectors.py:117-131
 *   energy reduction          pystella/reduction.py:165-173, sectors.py:133-144
 *
 * Layout: padded arrays (n, Nx+2h, Ny+2h, Nz+2h), z fastest; lap (n, Nx, Ny, Nz).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXS 4
#define MAXT 16

static const double LAPC[5][5] = {
    {0, 0, 0, 0, 0},
    {-2.0, 1.0, 0, 0, 0},
    {-30.0 / 12.0, 16.0 / 12.0, -1.0 / 12.0, 0, 0},
    {-490.0 / 180.0, 270.0 / 180.0, -27.0 / 180.0, 2.0 / 180.0, 0},
    {-14350.0 / 5040.0, 8064.0 / 5040.0, -1008.0 / 5040.0, 128.0 / 5040.0, -9.0 / 5040.0}};

#define IDX(i, j, k) (((int64_t)(i) * Py + (j)) * Pz + (k))

void orc_share_halos(double* a, int64_t Nx, int64_t Ny, int64_t Nz, int h) {
  const int64_t Px = Nx + 2 * h, Py = Ny + 2 * h, Pz = Nz + 2 * h;
  /* z, then x, then y; each over the padded extent of the other axes */
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t i = 0; i < Px; ++i)
    for (int64_t j = 0; j < Py; ++j)
      for (int k = 0; k < h; ++k) {
        a[IDX(i, j, k)] = a[IDX(i, j, Nz + k)];
        a[IDX(i, j, Nz + h + k)] = a[IDX(i, j, k + h)];
      }
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < h; ++i)
    for (int64_t j = 0; j < Py; ++j)
      for (int64_t k = 0; k < Pz; ++k) {
        a[IDX(i, j, k)] = a[IDX(Nx + i, j, k)];
        a[IDX(Nx + h + i, j, k)] = a[IDX(i + h, j, k)];
      }
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t i = 0; i < Px; ++i)
    for (int j = 0; j < h; ++j)
      for (int64_t k = 0; k < Pz; ++k) {
        a[IDX(i, j, k)] = a[IDX(i, Ny + j, k)];
        a[IDX(i, Ny + h + j, k)] = a[IDX(i, j + h, k)];
      }
}

void orc_laplacian(const double* f, double* lap, int64_t Nx, int64_t Ny, int64_t Nz, int h,
                   const double* dx) {
  const int64_t Py = Ny + 2 * h, Pz = Nz + 2 * h;
  const double wx = pow(dx[0], -2.0), wy = pow(dx[1], -2.0), wz = pow(dx[2], -2.0);
  const double* c = LAPC[h];
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t i = 0; i < Nx; ++i)
    for (int64_t j = 0; j < Ny; ++j) {
      const double* p = f + IDX(i + h, j + h, h);
      double* out = lap + (i * Ny + j) * Nz;
      for (int64_t k = 0; k < Nz; ++k) {
        double sx = c[0] * p[k], sy = c[0] * p[k], sz = c[0] * p[k];
        for (int s = 1; s <= h; ++s) {
          sx = sx + c[s] * p[k + s * Py * Pz];
          sx = sx + c[s] * p[k - s * Py * Pz];
          sy = sy + c[s] * p[k + s * Pz];
          sy = sy + c[s] * p[k - s * Pz];
          sz = sz + c[s] * p[k + s];
          sz = sz + c[s] * p[k - s];
        }
        out[k] = sx * wx + sy * wy + sz * wz;
      }
    }
}

typedef struct {
  int nscalars;
  int nterms[MAXS];          /* dV/df_n */
  double coef[MAXS][MAXT];
  int power[MAXS][MAXT][MAXS];
  int nvterms;               /* V */
  double vcoef[MAXT];
  int vpower[MAXT][MAXS];
} orc_potential;

static inline double poly(const double* fv, int ns, int nt, const double* coef,
                          const int (*power)[MAXS]) {
  double out = 0.0;
  for (int t = 0; t < nt; ++t) {
    double term = coef[t];
    for (int m = 0; m < ns; ++m)
      for (int e = 0; e < power[t][m]; ++e) term = term * fv[m];
    out = out + term;
  }
  return out;
}

void orc_lsrk_scalar_stage(double* f, double* dfdt, double* kf, double* kd, const double* lap,
                           int64_t Nx, int64_t Ny, int64_t Nz, int h, double dt, double A,
                           double B, double a, double hubble, const orc_potential* pot) {
  const int64_t Px = Nx + 2 * h, Py = Ny + 2 * h, Pz = Nz + 2 * h;
  const int64_t fs = Px * Py * Pz, ls = Nx * Ny * Nz;
  const int ns = pot->nscalars;
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t i = 0; i < Nx; ++i)
    for (int64_t j = 0; j < Ny; ++j)
      for (int64_t k = 0; k < Nz; ++k) {
        const int64_t o = IDX(i + h, j + h, k + h), ol = (i * Ny + j) * Nz + k;
        double fv[MAXS], dv[MAXS], rd[MAXS];
        for (int n = 0; n < ns; ++n) {
          fv[n] = f[o + n * fs];
          dv[n] = dfdt[o + n * fs];
        }
        for (int n = 0; n < ns; ++n) {
          const double dV = poly(fv, ns, pot->nterms[n], pot->coef[n], pot->power[n]);
          rd[n] = (lap[ol + n * ls] - (2.0 * hubble) * dv[n]) - (a * a) * dV;
        }
        for (int n = 0; n < ns; ++n) {
          const int64_t on = o + n * fs;
          const double kfv = A * kf[on] + dt * dv[n];
          const double kdv = A * kd[on] + dt * rd[n];
          kf[on] = kfv;
          kd[on] = kdv;
          f[on] = fv[n] + B * kfv;
          dfdt[on] = dv[n] + B * kdv;
        }
      }
}

/* out = {kinetic_n.., potential, gradient_n..} averaged over grid_size */
void orc_scalar_energy(const double* f, const double* dfdt, const double* lap, int64_t Nx,
                       int64_t Ny, int64_t Nz, int h, double a, double grid_size,
                       const orc_potential* pot, double* out) {
  const int64_t Px = Nx + 2 * h, Py = Ny + 2 * h, Pz = Nz + 2 * h;
  const int64_t fs = Px * Py * Pz, ls = Nx * Ny * Nz;
  const int ns = pot->nscalars;
  double kin[MAXS] = {0}, grad[MAXS] = {0}, potsum = 0.0;
#pragma omp parallel for collapse(2) schedule(static) reduction(+ : kin[:MAXS], grad[:MAXS], potsum)
  for (int64_t i = 0; i < Nx; ++i)
    for (int64_t j = 0; j < Ny; ++j)
      for (int64_t k = 0; k < Nz; ++k) {
        const int64_t o = IDX(i + h, j + h, k + h), ol = (i * Ny + j) * Nz + k;
        double fv[MAXS];
        for (int n = 0; n < ns; ++n) {
          fv[n] = f[o + n * fs];
          const double dv = dfdt[o + n * fs];
          kin[n] += dv * dv / 2 / (a * a);
          grad[n] += -fv[n] * lap[ol + n * ls] / 2 / (a * a);
        }
        potsum += poly(fv, ns, pot->nvterms, pot->vcoef, pot->vpower);
      }
  for (int n = 0; n < ns; ++n) {
    out[n] = kin[n] / grid_size;
    out[ns + 1 + n] = grad[n] / grid_size;
  }
  out[ns] = potsum / grid_size;
}

/* state = {a, adot, hubble, ka, kadot}; energies as orc_scalar_energy's out */
static void expansion_stage(double* st, const double* en, int ns, double mpl, double dt,
                            double A, double B) {
  double kin = 0, grad = 0;
  for (int n = 0; n < ns; ++n) {
    kin += en[n];
    grad += en[ns + 1 + n];
  }
  const double pot = en[ns];
  const double total = kin + pot + grad, pressure = kin - grad / 3.0 - pot;
  const double a = st[0], adot = st[1];
  const double rhs = 4.0 * M_PI * (a * a) / 3.0 / (mpl * mpl) * (total - 3.0 * pressure) * a;
  st[3] = A * st[3] + dt * adot;
  st[0] = a + B * st[3];
  st[4] = A * st[4] + dt * rhs;
  st[1] = adot + B * st[4];
  st[2] = st[1] / st[0];
}

/*
 * nsteps full low-storage RK steps in the reference's loop order
 * (stepper -> expand.step -> share_halos + lap + energy).  `energies` must hold
 * the energies of the initial state on entry (the caller computes them with
 * orc_share_halos/orc_laplacian/orc_scalar_energy) and holds the final ones on exit.
 */
void orc_preheating_steps(int nsteps, int nstages, const double* A, const double* B, double* f,
                          double* dfdt, double* kf, double* kd, double* lap, int64_t Nx,
                          int64_t Ny, int64_t Nz, int h, const double* dx, double dt,
                          double mpl, double grid_size, const orc_potential* pot,
                          double* state, double* energies, int with_expansion) {
  const int64_t fs = (Nx + 2 * h) * (Ny + 2 * h) * (Nz + 2 * h), ls = Nx * Ny * Nz;
  const int ns = pot->nscalars;
  for (int step = 0; step < nsteps; ++step)
    for (int s = 0; s < nstages; ++s) {
      if (!with_expansion) {
        /* examples/wave_equation.py:61-65: derivs then stepper */
        for (int n = 0; n < ns; ++n) {
          orc_share_halos(f + n * fs, Nx, Ny, Nz, h);
          orc_laplacian(f + n * fs, lap + n * ls, Nx, Ny, Nz, h, dx);
        }
        orc_lsrk_scalar_stage(f, dfdt, kf, kd, lap, Nx, Ny, Nz, h, dt, A[s], B[s], state[0],
                              state[2], pot);
        continue;
      }
      orc_lsrk_scalar_stage(f, dfdt, kf, kd, lap, Nx, Ny, Nz, h, dt, A[s], B[s], state[0],
                            state[2], pot);
      expansion_stage(state, energies, ns, mpl, dt, A[s], B[s]);
      for (int n = 0; n < ns; ++n) {
        orc_share_halos(f + n * fs, Nx, Ny, Nz, h);
        orc_laplacian(f + n * fs, lap + n * ls, Nx, Ny, Nz, h, dx);
      }
      orc_scalar_energy(f, dfdt, lap, Nx, Ny, Nz, h, state[0], grid_size, pot, energies);
    }
}

/* ---------------------------------------------------------------------------
 * Tensor-perturbation sector (config 4): first-derivative stencils
 * (pystella/derivs.py:127-131, 293-294; term order c1*f[+1], -c1*f[-1], c2*f[+2], ...),
 * TensorPerturbationSector.rhs_dict (pystella/sectors.py:187-204) and the loop of
 * examples/scalar_preheating.py:258-266 with gravitational waves enabled.
 * ------------------------------------------------------------------------- */
static const double GRADC[5][5] = {
    {0, 0, 0, 0, 0},
    {0, 1.0 / 2.0, 0, 0, 0},
    {0, 8.0 / 12.0, -1.0 / 12.0, 0, 0},
    {0, 45.0 / 60.0, -9.0 / 60.0, 1.0 / 60.0, 0},
    {0, 672.0 / 840.0, -168.0 / 840.0, 32.0 / 840.0, -3.0 / 840.0}};

/* grd: (3, Nx, Ny, Nz) */
void orc_gradient(const double* f, double* grd, int64_t Nx, int64_t Ny, int64_t Nz, int h,
                  const double* dx) {
  const int64_t Py = Ny + 2 * h, Pz = Nz + 2 * h, ls = Nx * Ny * Nz;
  const double wx = 1 / dx[0], wy = 1 / dx[1], wz = 1 / dx[2];
  const double* c = GRADC[h];
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t i = 0; i < Nx; ++i)
    for (int64_t j = 0; j < Ny; ++j) {
      const double* p = f + IDX(i + h, j + h, h);
      double* out = grd + (i * Ny + j) * Nz;
      for (int64_t k = 0; k < Nz; ++k) {
        double gx = 0, gy = 0, gz = 0;
        for (int s = 1; s <= h; ++s) {
          if (s == 1) {
            gx = c[s] * p[k + s * Py * Pz];
            gy = c[s] * p[k + s * Pz];
            gz = c[s] * p[k + s];
          } else {
            gx = gx + c[s] * p[k + s * Py * Pz];
            gy = gy + c[s] * p[k + s * Pz];
            gz = gz + c[s] * p[k + s];
          }
          gx = gx + (-1 * c[s]) * p[k - s * Py * Pz];
          gy = gy + (-1 * c[s]) * p[k - s * Pz];
          gz = gz + (-1 * c[s]) * p[k - s];
        }
        out[k] = gx * wx;
        out[k + ls] = gy * wy;
        out[k + 2 * ls] = gz * wz;
      }
    }
}

/* hij, dhij, kh, kdh: (6, padded); lap_hij: (6, N); dfdx: (ns, 3, N) */
void orc_lsrk_tensor_stage(double* hij, double* dhij, double* kh, double* kdh,
                           const double* lap_hij, const double* dfdx, int ns, int64_t Nx,
                           int64_t Ny, int64_t Nz, int h, double dt, double A, double B,
                           double hubble) {
  const int64_t Px = Nx + 2 * h, Py = Ny + 2 * h, Pz = Nz + 2 * h;
  const int64_t fs = Px * Py * Pz, ls = Nx * Ny * Nz;
  static const int CI[6] = {0, 0, 0, 1, 1, 2}, CJ[6] = {0, 1, 2, 1, 2, 2};
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t i = 0; i < Nx; ++i)
    for (int64_t j = 0; j < Ny; ++j)
      for (int64_t k = 0; k < Nz; ++k) {
        const int64_t o = IDX(i + h, j + h, k + h), ol = (i * Ny + j) * Nz + k;
        for (int c = 0; c < 6; ++c) {
          double S = 0;
          for (int n = 0; n < ns; ++n) {
            const double t = dfdx[ol + (n * 3 + CI[c]) * ls] * dfdx[ol + (n * 3 + CJ[c]) * ls];
            S = (n == 0) ? t : S + t;
          }
          const int64_t oc = o + c * fs;
          const double dv = dhij[oc];
          const double rhs = (lap_hij[ol + c * ls] - (2.0 * hubble) * dv) + (16.0 * M_PI) * S;
          const double khv = A * kh[oc] + dt * dv;
          const double kdv = A * kdh[oc] + dt * rhs;
          kh[oc] = khv;
          kdh[oc] = kdv;
          hij[oc] = hij[oc] + B * khv;
          dhij[oc] = dv + B * kdv;
        }
      }
}

/*
 * nsteps full steps of scalar fields + tensor perturbations + background, in the order of
 * examples/scalar_preheating.py:258-266: stepper (all sectors, from the old state) ->
 * expand.step -> derivs(f: lap, grd) + energies -> derivs(hij: lap).  On entry lap_f, dfdx,
 * lap_hij and energies must belong to the initial state.
 */
void orc_preheating_gw_steps(int nsteps, int nstages, const double* A, const double* B,
                             double* f, double* dfdt, double* kf, double* kd, double* lap_f,
                             double* dfdx, double* hij, double* dhij, double* kh, double* kdh,
                             double* lap_hij, int64_t Nx, int64_t Ny, int64_t Nz, int h,
                             const double* dx, double dt, double mpl, double grid_size,
                             const orc_potential* pot, double* state, double* energies,
                             int with_expansion) {
  const int64_t fs = (Nx + 2 * h) * (Ny + 2 * h) * (Nz + 2 * h), ls = Nx * Ny * Nz;
  const int ns = pot->nscalars;
  for (int step = 0; step < nsteps; ++step)
    for (int s = 0; s < nstages; ++s) {
      orc_lsrk_scalar_stage(f, dfdt, kf, kd, lap_f, Nx, Ny, Nz, h, dt, A[s], B[s], state[0],
                            state[2], pot);
      orc_lsrk_tensor_stage(hij, dhij, kh, kdh, lap_hij, dfdx, ns, Nx, Ny, Nz, h, dt, A[s],
                            B[s], state[2]);
      if (with_expansion) expansion_stage(state, energies, ns, mpl, dt, A[s], B[s]);
      for (int n = 0; n < ns; ++n) {
        orc_share_halos(f + n * fs, Nx, Ny, Nz, h);
        orc_laplacian(f + n * fs, lap_f + n * ls, Nx, Ny, Nz, h, dx);
        orc_gradient(f + n * fs, dfdx + n * 3 * ls, Nx, Ny, Nz, h, dx);
      }
      orc_scalar_energy(f, dfdt, lap_f, Nx, Ny, Nz, h, state[0], grid_size, pot, energies);
      for (int c = 0; c < 6; ++c) {
        orc_share_halos(hij + c * fs, Nx, Ny, Nz, h);
        orc_laplacian(hij + c * fs, lap_hij + c * ls, Nx, Ny, Nz, h, dx);
      }
    }
}

/* Parallel first touch: copies src into freshly allocated (untouched) dst with the same
 * static row partition the compute loops use, so that on a multi-socket host every thread
 * later works on pages of its own NUMA node (the timed CPU baseline otherwise swings by 2x
 * with wherever the single-threaded NumPy initialisation happened to place the pages). */
void orc_first_touch_copy(double* dst, const double* src, int64_t rows, int64_t row_len) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t k = 0; k < row_len; ++k) dst[r * row_len + k] = src ? src[r * row_len + k] : 0.0;
}

int orc_num_threads(void) {
  int n = 1;
#ifdef _OPENMP
#pragma omp parallel
  {
#pragma omp single
    n = omp_get_num_threads();
  }
#endif
  return n;
}
