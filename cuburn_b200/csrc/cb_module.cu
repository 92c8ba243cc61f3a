// Run-time compilation (NVRTC -> sm_100a cubin) and launch of the per-genome
// iterate module.  Replaces pycuda.compiler.compile + module_from_buffer +
// get_function/launch as used by the reference (cuburn/code/util.py:97-112,
// cuburn/render.py:232-246, 338-346).
//
// The CUDA driver API is reached through cudaGetDriverEntryPoint so the library
// has no link-time dependency on libcuda.so (it loads, and NVRTC compiles, on a
// machine without a GPU driver).
#include <cuda.h>
#include <nvrtc.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "cb_common.h"

namespace {

struct DriverApi {
    CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned,
                             unsigned, unsigned, unsigned, CUstream, void **,
                             void **) = nullptr;
    CUresult (*FuncGetAttribute)(int *, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int *, CUfunction, int,
                                                          size_t) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    CUresult (*ModuleGetGlobal)(CUdeviceptr *, size_t *, CUmodule, const char *) = nullptr;
    bool ready = false;
};

DriverApi g_drv;

template <typename F>
bool load_entry(const char *name, F *out) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        cb_set_error("cannot resolve driver entry point %s: %s", name,
                     e != cudaSuccess ? cudaGetErrorString(e) : "not found");
        return false;
    }
    *out = reinterpret_cast<F>(fn);
    return true;
}

int ensure_driver() {
    if (g_drv.ready) return CB_OK;
    CB_CUDA(cudaFree(0));
    if (!load_entry("cuModuleLoadData", &g_drv.ModuleLoadData) ||
        !load_entry("cuModuleUnload", &g_drv.ModuleUnload) ||
        !load_entry("cuModuleGetFunction", &g_drv.ModuleGetFunction) ||
        !load_entry("cuLaunchKernel", &g_drv.LaunchKernel) ||
        !load_entry("cuFuncGetAttribute", &g_drv.FuncGetAttribute) ||
        !load_entry("cuFuncSetAttribute", &g_drv.FuncSetAttribute) ||
        !load_entry("cuOccupancyMaxActiveBlocksPerMultiprocessor",
                    &g_drv.OccupancyMaxActiveBlocksPerMultiprocessor) ||
        !load_entry("cuGetErrorString", &g_drv.GetErrorString) ||
        !load_entry("cuModuleGetGlobal", &g_drv.ModuleGetGlobal))
        return CB_ERR_CUDA;
    g_drv.ready = true;
    return CB_OK;
}

const char *cu_err(CUresult r) {
    const char *s = nullptr;
    if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
    return s ? s : "unknown driver error";
}

#define CB_DRV(expr)                                                          \
    do {                                                                      \
        CUresult r__ = (expr);                                                \
        if (r__ != CUDA_SUCCESS) {                                            \
            cb_set_error("%s failed: %s", #expr, cu_err(r__));                \
            return CB_ERR_CUDA;                                               \
        }                                                                     \
    } while (0)

}  // namespace

struct cb_module_s {
    std::string name;
    std::vector<char> cubin;
    CUmodule mod = nullptr;     // loaded lazily: building needs no GPU
    std::map<std::string, CUfunction> funcs;
};

static int module_load(cb_module m) {
    if (m->mod) return CB_OK;
    int rc = ensure_driver();
    if (rc != CB_OK) return rc;
    CB_DRV(g_drv.ModuleLoadData(&m->mod, m->cubin.data()));
    return CB_OK;
}

static int module_function(cb_module m, const char *kernel, CUfunction *out) {
    int rc = module_load(m);
    if (rc != CB_OK) return rc;
    auto it = m->funcs.find(kernel);
    if (it == m->funcs.end()) {
        CUfunction f;
        CB_DRV(g_drv.ModuleGetFunction(&f, m->mod, kernel));
        it = m->funcs.emplace(kernel, f).first;
    }
    *out = it->second;
    return CB_OK;
}

extern "C" {

int cb_module_build(const char *source, const char *name,
                    const char *const *headers, const char *const *header_names,
                    int nheaders, const char *const *options, int noptions,
                    cb_module *out) {
    CB_REQUIRE(source && out, "source/out is null");
    nvrtcProgram prog;
    nvrtcResult r = nvrtcCreateProgram(&prog, source, name ? name : "cb_module.cu",
                                       nheaders, headers, header_names);
    if (r != NVRTC_SUCCESS) {
        cb_set_error("nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
        return CB_ERR_NVRTC;
    }
    r = nvrtcCompileProgram(prog, noptions, options);
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        nvrtcGetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        nvrtcGetProgramLog(prog, &log[0]);
        cb_set_error("NVRTC compilation of %s failed (%s):\n%s",
                     name ? name : "module", nvrtcGetErrorString(r), log.c_str());
        nvrtcDestroyProgram(&prog);
        return CB_ERR_NVRTC;
    }
    size_t n = 0;
    r = nvrtcGetCUBINSize(prog, &n);
    if (r != NVRTC_SUCCESS || n == 0) {
        cb_set_error("NVRTC produced no cubin (%s); is --gpu-architecture=sm_100a set?",
                     nvrtcGetErrorString(r));
        nvrtcDestroyProgram(&prog);
        return CB_ERR_NVRTC;
    }
    cb_module m = new cb_module_s();
    m->name = name ? name : "module";
    m->cubin.resize(n);
    nvrtcGetCUBIN(prog, m->cubin.data());
    nvrtcDestroyProgram(&prog);
    *out = m;
    return CB_OK;
}

int cb_module_destroy(cb_module m) {
    if (!m) return CB_OK;
    if (m->mod && g_drv.ready) g_drv.ModuleUnload(m->mod);
    delete m;
    return CB_OK;
}

int cb_module_get_cubin(cb_module m, const void **cubin, size_t *size) {
    CB_REQUIRE(m && cubin && size, "null argument");
    *cubin = m->cubin.data();
    *size = m->cubin.size();
    return CB_OK;
}

int cb_module_kernel_info(cb_module m, const char *kernel, int block_threads,
                          int *num_regs, int *static_smem, int *ctas_per_sm) {
    CB_REQUIRE(m && kernel, "null argument");
    CUfunction f;
    int rc = module_function(m, kernel, &f);
    if (rc != CB_OK) return rc;
    if (num_regs) CB_DRV(g_drv.FuncGetAttribute(num_regs, CU_FUNC_ATTRIBUTE_NUM_REGS, f));
    if (static_smem)
        CB_DRV(g_drv.FuncGetAttribute(static_smem, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, f));
    if (ctas_per_sm)
        CB_DRV(g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, f,
                                                               block_threads, 0));
    return CB_OK;
}

int cb_module_kernel_local_bytes(cb_module m, const char *kernel, int *local_bytes) {
    CB_REQUIRE(m && kernel && local_bytes, "null argument");
    CUfunction f;
    int rc = module_function(m, kernel, &f);
    if (rc != CB_OK) return rc;
    CB_DRV(g_drv.FuncGetAttribute(local_bytes, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, f));
    return CB_OK;
}

int cb_module_launch(cb_module m, const char *kernel, int gx, int gy, int gz,
                     int bx, int by, int bz, int dyn_smem, void **args,
                     cb_stream s) {
    CB_REQUIRE(m && kernel, "null argument");
    CUfunction f;
    int rc = module_function(m, kernel, &f);
    if (rc != CB_OK) return rc;
    CB_REQUIRE(dyn_smem >= 0 && dyn_smem <= 227 * 1024, "dynamic shared memory beyond 227 KB");
    if (dyn_smem > 48 * 1024)       // opt in to the large carve-out (up to 227 KB per CTA)
        CB_DRV(g_drv.FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                      dyn_smem));
    CB_DRV(g_drv.LaunchKernel(f, gx, gy, gz, bx, by, bz, dyn_smem,
                              (CUstream)cb_cs(s), args, nullptr));
    g_cb_launches++;
    return CB_OK;
}

int cb_module_set_global(cb_module m, const char *symbol, cb_dptr src, size_t bytes,
                         cb_stream s) {
    CB_REQUIRE(m && symbol, "null argument");
    int rc = module_load(m);
    if (rc != CB_OK) return rc;
    CUdeviceptr dst = 0;
    size_t size = 0;
    CB_DRV(g_drv.ModuleGetGlobal(&dst, &size, m->mod, symbol));
    CB_REQUIRE(bytes <= size, "source larger than the module global");
    CB_CUDA(cudaMemcpyAsync((void *)dst, (const void *)src, bytes,
                            cudaMemcpyDeviceToDevice, cb_cs(s)));
    return CB_OK;
}

int cb_iterate(cb_module m, const cb_iter_args *args, int grid_ctas, cb_stream s) {
    CB_REQUIRE(m && args, "null argument");
    CB_REQUIRE(grid_ctas > 0, "grid_ctas must be positive");
    CB_REQUIRE(args->first_sample % 32768ull == 0, "first_sample must be unit aligned");
    CB_REQUIRE(args->nts > 0 && args->pal_rows > 0, "bad temporal sample counts");
    CB_REQUIRE(args->tickets || (!args->dynamic && !args->spill),
               "tickets scratch needed for dynamic units and for the spill sweep");
    cb_iter_args a = *args;
    if (a.tickets)
        CB_CUDA(cudaMemsetAsync((void *)a.tickets, 0, 8, cb_cs(s)));
    void *kargs[1] = {&a};
    return cb_module_launch(m, "cb_iter", grid_ctas, 1, 1, 256, 1, 1, 0, kargs, s);
}

}  // extern "C"
