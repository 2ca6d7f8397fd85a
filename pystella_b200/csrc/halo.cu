// Halo kernels of DomainDecomposition (pystella/decomp.py:157-254):
// in-place periodic wrap (pack_unpack_{z,x,y}), face pack/unpack for decomposed
// axes, and padding removal/restoration (gather_knl / scatter_knl).
#include "common.cuh"

namespace psb {

struct HaloGeom {
  int64_t Nx, Ny, Nz;
  int hx, hy, hz;
  int64_t Px, Py, Pz;
};

static HaloGeom make_geom(int64_t Nx, int64_t Ny, int64_t Nz, int hx, int hy, int hz) {
  HaloGeom g{Nx, Ny, Nz, hx, hy, hz, Nx + 2 * hx, Ny + 2 * hy, Nz + 2 * hz};
  return g;
}

// arr[.., .., k] = arr[.., .., Nz+k] ; arr[.., .., Nz+hz+k] = arr[.., .., k+hz]
// over ALL padded (i, j)   (decomp.py:232-233)
template <typename T>
__global__ void wrap_z_kernel(T* arr, HaloGeom g) {
  const int64_t rows = g.Px * g.Py;
  const int64_t total = rows * g.hz;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / g.hz;
    const int k = (int)(e % g.hz);
    T* r = arr + row * g.Pz;
    r[k] = r[g.Nz + k];
    r[g.Nz + g.hz + k] = r[k + g.hz];
  }
}

// arr[i, :, :] = arr[Nx+i, :, :] ; arr[Nx+hx+i, :, :] = arr[i+hx, :, :]  (decomp.py:181-182)
template <typename T>
__global__ void wrap_x_kernel(T* arr, HaloGeom g) {
  const int64_t plane = g.Py * g.Pz;
  const int64_t total = plane * g.hx;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / plane, r = e % plane;
    arr[i * plane + r] = arr[(g.Nx + i) * plane + r];
    arr[(g.Nx + g.hx + i) * plane + r] = arr[(i + g.hx) * plane + r];
  }
}

// arr[:, j, :] = arr[:, Ny+j, :] ; arr[:, Ny+hy+j, :] = arr[:, j+hy, :]  (decomp.py:208-209)
template <typename T>
__global__ void wrap_y_kernel(T* arr, HaloGeom g) {
  const int64_t total = g.Px * g.hy * g.Pz;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = e % g.Pz;
    const int64_t j = (e / g.Pz) % g.hy;
    const int64_t i = e / (g.Pz * g.hy);
    T* pl = arr + i * g.Py * g.Pz;
    pl[j * g.Pz + k] = pl[(g.Ny + j) * g.Pz + k];
    pl[(g.Ny + g.hy + j) * g.Pz + k] = pl[(j + g.hy) * g.Pz + k];
  }
}

// buf[1, i, j, k] = arr[Nx+i, j, k] ; buf[0, i, j, k] = arr[i+hx, j, k]  (decomp.py:177-178)
// unpack: arr[i, j, k] = buf[1, ...] ; arr[Nx+hx+i, j, k] = buf[0, ...]  (decomp.py:179-180)
template <typename T, bool PACK>
__global__ void face_x_kernel(T* arr, T* buf, HaloGeom g) {
  const int64_t plane = g.Py * g.Pz;
  const int64_t total = plane * g.hx;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / plane, r = e % plane;
    if (PACK) {
      buf[total + e] = arr[(g.Nx + i) * plane + r];
      buf[e] = arr[(i + g.hx) * plane + r];
    } else {
      arr[i * plane + r] = buf[total + e];
      arr[(g.Nx + g.hx + i) * plane + r] = buf[e];
    }
  }
}

// buf[1, i, j, k] = arr[i, Ny+j, k] ; buf[0, i, j, k] = arr[i, j+hy, k]  (decomp.py:204-205)
// unpack: arr[i, j, k] = buf[1, ...] ; arr[i, Ny+hy+j, k] = buf[0, ...]  (decomp.py:206-207)
template <typename T, bool PACK>
__global__ void face_y_kernel(T* arr, T* buf, HaloGeom g) {
  const int64_t total = g.Px * g.hy * g.Pz;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = e % g.Pz;
    const int64_t j = (e / g.Pz) % g.hy;
    const int64_t i = e / (g.Pz * g.hy);
    T* pl = arr + i * g.Py * g.Pz;
    if (PACK) {
      buf[total + e] = pl[(g.Ny + j) * g.Pz + k];
      buf[e] = pl[(j + g.hy) * g.Pz + k];
    } else {
      pl[j * g.Pz + k] = buf[total + e];
      pl[(g.Ny + g.hy + j) * g.Pz + k] = buf[e];
    }
  }
}

// arr[i, j, k] = subarr[i+hx, j+hy, k+hz]  /  the reverse  (decomp.py:253-254)
template <typename T, bool REMOVE>
__global__ void pad_kernel(T* padded, T* unpadded, HaloGeom g) {
  const int64_t total = g.Nx * g.Ny * g.Nz;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = e % g.Nz;
    const int64_t j = (e / g.Nz) % g.Ny;
    const int64_t i = e / (g.Nz * g.Ny);
    const int64_t o = ((i + g.hx) * g.Py + (j + g.hy)) * g.Pz + (k + g.hz);
    if (REMOVE)
      unpadded[e] = padded[o];
    else
      padded[o] = unpadded[e];
  }
}

static inline int grid_for(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

template <typename T>
static int wrap_typed(void* arr, const HaloGeom& g, int mask, cudaStream_t s) {
  T* a = static_cast<T*>(arr);
  constexpr int TH = 256;
  // order z -> x -> y (decomp.py:387-449)
  if ((mask & 4) && g.hz > 0)
    wrap_z_kernel<T><<<grid_for(g.Px * g.Py * g.hz, TH), TH, 0, s>>>(a, g);
  if ((mask & 1) && g.hx > 0)
    wrap_x_kernel<T><<<grid_for(g.Py * g.Pz * g.hx, TH), TH, 0, s>>>(a, g);
  if ((mask & 2) && g.hy > 0)
    wrap_y_kernel<T><<<grid_for(g.Px * g.hy * g.Pz, TH), TH, 0, s>>>(a, g);
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T, bool PACK>
static int face_typed(void* arr, void* buf, const HaloGeom& g, int axis, cudaStream_t s) {
  constexpr int TH = 256;
  T* a = static_cast<T*>(arr);
  T* b = static_cast<T*>(buf);
  if (axis == 0) {
    if (g.hx > 0)
      face_x_kernel<T, PACK><<<grid_for(g.Py * g.Pz * g.hx, TH), TH, 0, s>>>(a, b, g);
  } else if (axis == 1) {
    if (g.hy > 0)
      face_y_kernel<T, PACK><<<grid_for(g.Px * g.hy * g.Pz, TH), TH, 0, s>>>(a, b, g);
  } else {
    set_error("halo pack/unpack: axis %d not decomposable (z is never split)", axis);
    return 1;
  }
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T, bool REMOVE>
static int pad_typed(void* padded, void* unpadded, const HaloGeom& g, cudaStream_t s) {
  constexpr int TH = 256;
  pad_kernel<T, REMOVE><<<grid_for(g.Nx * g.Ny * g.Nz, TH), TH, 0, s>>>(
      static_cast<T*>(padded), static_cast<T*>(unpadded), g);
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace psb

using namespace psb;

#define PSB_GEOM_CHECKS(name)                                                  \
  PSB_REQUIRE(Nx > 0 && Ny > 0 && Nz > 0, name ": empty grid");                \
  PSB_REQUIRE(hx >= 0 && hy >= 0 && hz >= 0, name ": negative halo");          \
  PSB_REQUIRE(dtype == PSB_F64 || dtype == PSB_F32, name ": unknown dtype %d", dtype)

extern "C" int psb_halo_wrap(void* arr, int dtype, int64_t Nx, int64_t Ny, int64_t Nz,
                             int hx, int hy, int hz, int axes_mask, void* stream) {
  PSB_REQUIRE(arr != nullptr, "psb_halo_wrap: null array");
  PSB_GEOM_CHECKS("psb_halo_wrap");
  PSB_REQUIRE(Nx >= hx && Ny >= hy && Nz >= hz, "psb_halo_wrap: grid smaller than halo");
  HaloGeom g = make_geom(Nx, Ny, Nz, hx, hy, hz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dtype == PSB_F64 ? wrap_typed<double>(arr, g, axes_mask, s)
                          : wrap_typed<float>(arr, g, axes_mask, s);
}

extern "C" int psb_halo_pack(const void* arr, void* buf, int dtype, int64_t Nx, int64_t Ny,
                             int64_t Nz, int hx, int hy, int hz, int axis, void* stream) {
  PSB_REQUIRE(arr && buf, "psb_halo_pack: null pointer");
  PSB_GEOM_CHECKS("psb_halo_pack");
  HaloGeom g = make_geom(Nx, Ny, Nz, hx, hy, hz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  void* a = const_cast<void*>(arr);
  return dtype == PSB_F64 ? face_typed<double, true>(a, buf, g, axis, s)
                          : face_typed<float, true>(a, buf, g, axis, s);
}

extern "C" int psb_halo_unpack(void* arr, const void* buf, int dtype, int64_t Nx,
                               int64_t Ny, int64_t Nz, int hx, int hy, int hz, int axis,
                               void* stream) {
  PSB_REQUIRE(arr && buf, "psb_halo_unpack: null pointer");
  PSB_GEOM_CHECKS("psb_halo_unpack");
  HaloGeom g = make_geom(Nx, Ny, Nz, hx, hy, hz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  void* b = const_cast<void*>(buf);
  return dtype == PSB_F64 ? face_typed<double, false>(arr, b, g, axis, s)
                          : face_typed<float, false>(arr, b, g, axis, s);
}

extern "C" int psb_remove_halos(const void* padded, void* unpadded, int dtype, int64_t Nx,
                                int64_t Ny, int64_t Nz, int hx, int hy, int hz,
                                void* stream) {
  PSB_REQUIRE(padded && unpadded, "psb_remove_halos: null pointer");
  PSB_GEOM_CHECKS("psb_remove_halos");
  HaloGeom g = make_geom(Nx, Ny, Nz, hx, hy, hz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  void* p = const_cast<void*>(padded);
  return dtype == PSB_F64 ? pad_typed<double, true>(p, unpadded, g, s)
                          : pad_typed<float, true>(p, unpadded, g, s);
}

extern "C" int psb_restore_halos(const void* unpadded, void* padded, int dtype, int64_t Nx,
                                 int64_t Ny, int64_t Nz, int hx, int hy, int hz,
                                 void* stream) {
  PSB_REQUIRE(padded && unpadded, "psb_restore_halos: null pointer");
  PSB_GEOM_CHECKS("psb_restore_halos");
  HaloGeom g = make_geom(Nx, Ny, Nz, hx, hy, hz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  void* u = const_cast<void*>(unpadded);
  return dtype == PSB_F64 ? pad_typed<double, false>(padded, u, g, s)
                          : pad_typed<float, false>(padded, u, g, s);
}
