// NVRTC + driver-API runtime for generated element-wise kernels (EXPERIMENTAL).
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "pystella_b200.h"
#include "pystella_b200_jit.h"

namespace {

struct Entry {
  std::string cubin;
  std::string name;
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  bool loaded = false;
  bool live = false;
  unsigned smem_opted = 0;
};

std::vector<Entry> g_entries;
std::mutex g_mutex;
thread_local std::string g_log;

template <typename F> F driver(const char* sym) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(sym, &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<F>(p);
}

typedef CUresult (*ModuleLoadData)(CUmodule*, const void*);
typedef CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
typedef CUresult (*ModuleUnload)(CUmodule);
typedef CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned,
                                 unsigned, unsigned, CUstream, void**, void**);
typedef CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
typedef CUresult (*LaunchKernelEx)(const CUlaunchConfig*, CUfunction, void**, void**);

// loads the module of entry `e` on first use; g_mutex must be held
int ensure_loaded(Entry* e) {
  if (e->loaded) return 0;
  cudaFree(0);  // make sure the primary context exists
  auto load = driver<ModuleLoadData>("cuModuleLoadData");
  auto getf = driver<ModuleGetFunction>("cuModuleGetFunction");
  if (!load || !getf) {
    g_log = "driver entry points unavailable";
    return 2;
  }
  CUresult r = load(&e->mod, e->cubin.data());
  if (r != CUDA_SUCCESS) {
    g_log = "cuModuleLoadData failed (" + std::to_string((int)r) + ")";
    return 2;
  }
  r = getf(&e->fn, e->mod, e->name.c_str());
  if (r != CUDA_SUCCESS) {
    g_log = "cuModuleGetFunction(" + e->name + ") failed (" + std::to_string((int)r) + ")";
    return 2;
  }
  e->loaded = true;
  return 0;
}

}  // namespace

extern "C" const char* psb_jit_last_log(void) { return g_log.c_str(); }

extern "C" int64_t psb_jit_compile(const char* source, const char* kernel_name,
                                   const char* arch) {
  g_log.clear();
  if (!source || !kernel_name) {
    g_log = "psb_jit_compile: null source / kernel name";
    return 0;
  }
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, source, "psb_generated.cu", 0, nullptr, nullptr) !=
      NVRTC_SUCCESS) {
    g_log = "nvrtcCreateProgram failed";
    return 0;
  }
  std::string archopt = std::string("--gpu-architecture=") + (arch ? arch : "sm_100a");
  const char* opts[] = {archopt.c_str(), "--std=c++17", "-default-device",
                        "--fmad=true"};
  nvrtcResult rc = nvrtcCompileProgram(prog, 4, opts);
  size_t logsz = 0;
  nvrtcGetProgramLogSize(prog, &logsz);
  if (logsz > 1) {
    g_log.resize(logsz);
    nvrtcGetProgramLog(prog, &g_log[0]);
  }
  if (rc != NVRTC_SUCCESS) {
    g_log = std::string("NVRTC: ") + nvrtcGetErrorString(rc) + "\n" + g_log;
    nvrtcDestroyProgram(&prog);
    return 0;
  }
  size_t sz = 0;
  if (nvrtcGetCUBINSize(prog, &sz) != NVRTC_SUCCESS || sz == 0) {
    g_log = "NVRTC produced no cubin (is the architecture a real one, e.g. sm_100a?)";
    nvrtcDestroyProgram(&prog);
    return 0;
  }
  Entry e;
  e.cubin.resize(sz);
  nvrtcGetCUBIN(prog, &e.cubin[0]);
  nvrtcDestroyProgram(&prog);
  e.name = kernel_name;
  e.live = true;
  std::lock_guard<std::mutex> lock(g_mutex);
  g_entries.push_back(std::move(e));
  return (int64_t)g_entries.size();
}

extern "C" int64_t psb_jit_load_cubin(const void* cubin, size_t size, const char* symbol) {
  g_log.clear();
  if (!cubin || size == 0 || !symbol) {
    g_log = "psb_jit_load_cubin: null cubin / symbol";
    return 0;
  }
  Entry e;
  e.cubin.assign(static_cast<const char*>(cubin), size);
  e.name = symbol;
  e.live = true;
  std::lock_guard<std::mutex> lock(g_mutex);
  g_entries.push_back(std::move(e));
  return (int64_t)g_entries.size();
}

extern "C" int64_t psb_jit_compile_ex(const char* source, const char* kernel_name,
                                      const char* name_expression, const char* arch,
                                      const char* const* options, int noptions,
                                      const char* const* header_names,
                                      const char* const* header_sources, int nheaders) {
  g_log.clear();
  if (!source || (!kernel_name && !name_expression)) {
    g_log = "psb_jit_compile_ex: null source / kernel name";
    return 0;
  }
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, source, "psb_generated.cu", nheaders, header_sources,
                         header_names) != NVRTC_SUCCESS) {
    g_log = "nvrtcCreateProgram failed";
    return 0;
  }
  if (name_expression && nvrtcAddNameExpression(prog, name_expression) != NVRTC_SUCCESS) {
    g_log = "nvrtcAddNameExpression failed";
    nvrtcDestroyProgram(&prog);
    return 0;
  }
  std::string archopt = std::string("--gpu-architecture=") + (arch ? arch : "sm_100a");
  std::vector<const char*> opts(options, options + (noptions > 0 ? noptions : 0));
  opts.push_back(archopt.c_str());
  nvrtcResult rc = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
  size_t logsz = 0;
  nvrtcGetProgramLogSize(prog, &logsz);
  if (logsz > 1) {
    g_log.resize(logsz);
    nvrtcGetProgramLog(prog, &g_log[0]);
  }
  if (rc != NVRTC_SUCCESS) {
    g_log = std::string("NVRTC: ") + nvrtcGetErrorString(rc) + "\n" + g_log;
    nvrtcDestroyProgram(&prog);
    return 0;
  }
  Entry e;
  if (name_expression) {
    const char* lowered = nullptr;
    if (nvrtcGetLoweredName(prog, name_expression, &lowered) != NVRTC_SUCCESS || !lowered) {
      g_log = std::string("no lowered name for ") + name_expression;
      nvrtcDestroyProgram(&prog);
      return 0;
    }
    e.name = lowered;
  } else {
    e.name = kernel_name;
  }
  size_t sz = 0;
  if (nvrtcGetCUBINSize(prog, &sz) != NVRTC_SUCCESS || sz == 0) {
    g_log = "NVRTC produced no cubin (is the architecture a real one, e.g. sm_100a?)";
    nvrtcDestroyProgram(&prog);
    return 0;
  }
  e.cubin.resize(sz);
  nvrtcGetCUBIN(prog, &e.cubin[0]);
  nvrtcDestroyProgram(&prog);
  e.live = true;
  std::lock_guard<std::mutex> lock(g_mutex);
  g_entries.push_back(std::move(e));
  return (int64_t)g_entries.size();
}

extern "C" const char* psb_jit_kernel_symbol(int64_t handle) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (handle < 1 || handle > (int64_t)g_entries.size() || !g_entries[handle - 1].live)
    return "";
  return g_entries[handle - 1].name.c_str();
}

extern "C" int psb_jit_write_cubin(int64_t handle, const char* path) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (handle < 1 || handle > (int64_t)g_entries.size() || !g_entries[handle - 1].live || !path)
    return 1;
  FILE* f = fopen(path, "wb");
  if (!f) return 2;
  const std::string& c = g_entries[handle - 1].cubin;
  size_t n = fwrite(c.data(), 1, c.size(), f);
  fclose(f);
  return n == c.size() ? 0 : 2;
}

extern "C" int psb_jit_launch_params(int64_t handle, void** params, unsigned gx, unsigned gy,
                                     unsigned gz, unsigned bx, unsigned by, unsigned bz,
                                     unsigned smem_bytes, void* stream) {
  return psb_jit_launch_params_v(handle, params, gx, gy, gz, bx, by, bz, smem_bytes, 0, stream);
}

extern "C" int psb_jit_launch_params_v(int64_t handle, void** params, unsigned gx,
                                       unsigned gy, unsigned gz, unsigned bx, unsigned by,
                                       unsigned bz, unsigned smem_bytes, int launch_flags,
                                       void* stream) {
  g_log.clear();
  if (!params) {
    g_log = "psb_jit_launch_params: null parameter list";
    return 1;
  }
  Entry* e = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (handle < 1 || handle > (int64_t)g_entries.size() || !g_entries[handle - 1].live) {
      g_log = "psb_jit_launch_params: bad handle";
      return 1;
    }
    e = &g_entries[handle - 1];
    if (int rc = ensure_loaded(e)) return rc;
    if (smem_bytes > e->smem_opted) {
      auto setattr = driver<FuncSetAttribute>("cuFuncSetAttribute");
      if (!setattr ||
          setattr(e->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem_bytes) !=
              CUDA_SUCCESS) {
        g_log = "cuFuncSetAttribute(max dynamic shared memory) failed";
        return 2;
      }
      e->smem_opted = smem_bytes;
    }
  }
  if (launch_flags & PSB_LAUNCH_PDL) {
    // programmatic dependent launch (the generated instance of the fused pipeline executes
    // griddepcontrol.wait like the precompiled one)
    static LaunchKernelEx launch_ex = driver<LaunchKernelEx>("cuLaunchKernelEx");
    if (!launch_ex) {
      g_log = "psb_jit_launch_params_v: cuLaunchKernelEx unavailable";
      return 2;
    }
    CUlaunchConfig cfg = {};
    cfg.gridDimX = gx, cfg.gridDimY = gy, cfg.gridDimZ = gz;
    cfg.blockDimX = bx, cfg.blockDimY = by, cfg.blockDimZ = bz;
    cfg.sharedMemBytes = smem_bytes;
    cfg.hStream = static_cast<CUstream>(stream);
    CUlaunchAttribute attr[1];
    attr[0].id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
    attr[0].value.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUresult r = launch_ex(&cfg, e->fn, params, nullptr);
    if (r != CUDA_SUCCESS) {
      g_log = "cuLaunchKernelEx failed (" + std::to_string((int)r) + ")";
      return 2;
    }
    return 0;
  }
  static LaunchKernel launch = driver<LaunchKernel>("cuLaunchKernel");
  if (!launch) {
    g_log = "psb_jit_launch_params: cuLaunchKernel unavailable";
    return 2;
  }
  CUresult r = launch(e->fn, gx, gy, gz, bx, by, bz, smem_bytes,
                      static_cast<CUstream>(stream), params, nullptr);
  if (r != CUDA_SUCCESS) {
    g_log = "cuLaunchKernel failed (" + std::to_string((int)r) + ")";
    return 2;
  }
  return 0;
}

extern "C" size_t psb_jit_cubin_size(int64_t handle) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (handle < 1 || handle > (int64_t)g_entries.size() || !g_entries[handle - 1].live) return 0;
  return g_entries[handle - 1].cubin.size();
}

extern "C" int psb_jit_launch(int64_t handle, const psb_jit_args* args, unsigned gx,
                              unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                              void* stream) {
  g_log.clear();
  Entry* e = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (handle < 1 || handle > (int64_t)g_entries.size() || !g_entries[handle - 1].live) {
      g_log = "psb_jit_launch: bad handle";
      return 1;
    }
    e = &g_entries[handle - 1];
    if (int rc = ensure_loaded(e)) return rc;
  }
  if (!args) {
    g_log = "psb_jit_launch: null args";
    return 1;
  }
  static LaunchKernel launch = driver<LaunchKernel>("cuLaunchKernel");
  if (!launch) {
    g_log = "psb_jit_launch: cuLaunchKernel unavailable";
    return 2;
  }
  psb_jit_args copy = *args;
  void* params[] = {&copy};
  CUresult r = launch(e->fn, gx, gy, gz, bx, by, bz, 0, static_cast<CUstream>(stream), params,
                      nullptr);
  if (r != CUDA_SUCCESS) {
    g_log = "cuLaunchKernel failed (" + std::to_string((int)r) + ")";
    return 2;
  }
  return 0;
}

extern "C" int psb_jit_release(int64_t handle) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (handle < 1 || handle > (int64_t)g_entries.size() || !g_entries[handle - 1].live) return 1;
  Entry& e = g_entries[handle - 1];
  if (e.loaded) {
    auto unload = driver<ModuleUnload>("cuModuleUnload");
    if (unload) unload(e.mod);
  }
  e = Entry();
  return 0;
}
