// Run-time build of user-defined targets: pb2_user_target.cuh + the user's CUDA source are compiled with NVRTC for
// sm_100a (libnvrtc is bound at run time with dlopen -- libpb2 has no link-time dependency on it), the cubin is loaded
// through the CUDA runtime's library API and its four kernels (log-prob + gradient, leapfrog, HMC, NUTS) are launched
// by the same launcher as the offline-compiled chain kernels (pb2_chain_kernels.cu: launch_chain).
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstring>
#include <string>
#include <vector>

#include "pb2_internal.h"

namespace pb2 {

namespace {

struct Nvrtc {
  void* so = nullptr;
  decltype(&nvrtcCreateProgram) create = nullptr;
  decltype(&nvrtcCompileProgram) compile = nullptr;
  decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
  decltype(&nvrtcGetProgramLog) log = nullptr;
  decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
  decltype(&nvrtcGetCUBIN) cubin = nullptr;
  decltype(&nvrtcDestroyProgram) destroy = nullptr;
  decltype(&nvrtcGetErrorString) err = nullptr;
};

template <class F>
bool bind(void* so, const char* name, F& f) {
  f = reinterpret_cast<F>(dlsym(so, name));
  return f != nullptr;
}

const Nvrtc* nvrtc() {
  static Nvrtc n;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
      n.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (n.so) break;
    }
    if (n.so) {
      const bool ok = bind(n.so, "nvrtcCreateProgram", n.create) && bind(n.so, "nvrtcCompileProgram", n.compile) &&
                      bind(n.so, "nvrtcGetProgramLogSize", n.log_size) && bind(n.so, "nvrtcGetProgramLog", n.log) &&
                      bind(n.so, "nvrtcGetCUBINSize", n.cubin_size) && bind(n.so, "nvrtcGetCUBIN", n.cubin) &&
                      bind(n.so, "nvrtcDestroyProgram", n.destroy) && bind(n.so, "nvrtcGetErrorString", n.err);
      if (!ok) { dlclose(n.so); n.so = nullptr; }
    }
  }
  return n.so ? &n : nullptr;
}

}  // namespace

int user_elements_per_lane(int dim) {
  for (int e : {1, 2, 4, 8})
    if (dim <= 32 * e) return e;
  return 0;
}

// Compile the user's source for a D-dimensional target.  On failure the compiler's log is the error message.
int user_compile(pb2_ctx* ctx, const char* source, int dim, const char* include_dir, int variant, int flags,
                 std::vector<char>& cubin) {
  if (!source || !include_dir) return set_error(ctx, PB2_ERR_INVALID, "user target: NULL source / include directory");
  const int E = user_elements_per_lane(dim);
  if (dim < 1 || E == 0) return set_error(ctx, PB2_ERR_UNSUPPORTED, "user target: need 1 <= D <= 256");
  const Nvrtc* n = nvrtc();
  if (!n) return set_error(ctx, PB2_ERR_UNSUPPORTED, "user target: libnvrtc.so.12 could not be loaded");
  std::string tu = "#include \"pb2_user_target.cuh\"\n#line 1 \"user_target.cu\"\n";
  tu += source;
  tu += "\n";
  nvrtcProgram prog;
  nvrtcResult r = n->create(&prog, tu.c_str(), "pb2_user_target_tu.cu", 0, nullptr, nullptr);
  if (r != NVRTC_SUCCESS) return set_error(ctx, PB2_ERR_CUDA, std::string("nvrtcCreateProgram: ") + n->err(r));
  const std::string inc = std::string("-I") + include_dir;
  const std::string edef = "-DPB2_USER_E=" + std::to_string(E);
  const std::string vdef = "-DPB2_USER_VARIANT=" + std::to_string(variant);
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", inc.c_str(), edef.c_str(), vdef.c_str(),
                        "-DPB2_USER_COOPERATIVE"};
  const int nopts = (int)(sizeof(opts) / sizeof(opts[0])) - ((flags & PB2_USER_COOPERATIVE) ? 0 : 1);
  r = n->compile(prog, nopts, opts);
  if (r != NVRTC_SUCCESS) {
    size_t ls = 0;
    n->log_size(prog, &ls);
    std::string log(ls ? ls : 1, '\0');
    if (ls) n->log(prog, &log[0]);
    n->destroy(&prog);
    return set_error(ctx, PB2_ERR_INVALID, std::string("user target does not compile (") + n->err(r) + "):\n" + log.c_str());
  }
  size_t cs = 0;
  n->cubin_size(prog, &cs);
  cubin.resize(cs);
  r = n->cubin(prog, cubin.data());
  n->destroy(&prog);
  if (r != NVRTC_SUCCESS || cs == 0) return set_error(ctx, PB2_ERR_CUDA, "nvrtcGetCUBIN failed");
  return PB2_OK;
}

// Build (once) and load variant `variant` of the target's kernels: 0 plain, 1 ScaledT, 2 TransformedT.
int user_target_ensure(pb2_ctx* ctx, pb2_target* t, int variant) {
  if (variant < 0 || variant >= kUserVariants) return set_error(ctx, PB2_ERR_INVALID, "user target: bad variant");
  if (t->user_lib[variant]) return PB2_OK;
  std::vector<char> cubin;
  if (int rc = user_compile(ctx, t->user_source.c_str(), t->dim, t->user_include.c_str(), variant, t->user_flags, cubin)) return rc;
  cudaLibrary_t lib = nullptr;
  if (int rc = check_cuda(ctx, cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0),
                          "cudaLibraryLoadData(user target)"))
    return rc;
  static const char* names[4] = {"pb2_user_logp_grad", "pb2_user_leapfrog", "pb2_user_hmc", "pb2_user_nuts"};
  for (int m = 0; m < 4; ++m) {
    cudaKernel_t k = nullptr;
    if (int rc = check_cuda(ctx, cudaLibraryGetKernel(&k, lib, names[m]), "cudaLibraryGetKernel(user target)")) {
      cudaLibraryUnload(lib);
      return rc;
    }
    t->user_kernels[variant][m] = (void*)k;
  }
  t->user_lib[variant] = (void*)lib;
  return PB2_OK;
}

void user_target_unload(pb2_target* t) {
  for (int v = 0; v < kUserVariants; ++v) {
    if (t->user_lib[v]) cudaLibraryUnload((cudaLibrary_t)t->user_lib[v]);
    t->user_lib[v] = nullptr;
  }
}

}  // namespace pb2
