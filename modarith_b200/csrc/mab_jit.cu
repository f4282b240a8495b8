// mab_jit.cu -- NVRTC front end and kernel cache of mab_<P>_modprog_jit (see mab_jit.h).
//
// NVRTC is loaded with dlopen on first use, so the library itself links against nothing but the CUDA runtime
// and loads on machines without the compiler; a JIT call there fails with MAB_ERR_NOJIT (there is no fall-back:
// mab_<P>_modprog, the interpreter, is a separate entry point the caller chooses).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include "modarith_b200.h"
#include "mab_jit.h"

namespace {

// the few NVRTC entry points used, by their documented C signatures (nvrtc.h)
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* handle = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool tried = false;
  std::string why;
};
Nvrtc g_nvrtc;
std::mutex g_nvrtc_mutex;

const Nvrtc* nvrtc() {
  std::lock_guard<std::mutex> g(g_nvrtc_mutex);
  Nvrtc& N = g_nvrtc;
  if (N.tried) return N.handle ? &N : nullptr;
  N.tried = true;
  const char* env = getenv("MAB_NVRTC");                                   // explicit path, if the default search fails
  const char* cands[] = {env, "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                         "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so.13"};
  for (const char* c : cands) {
    if (!c || !*c) continue;
    N.handle = dlopen(c, RTLD_NOW | RTLD_LOCAL);
    if (N.handle) break;
    const char* e = dlerror();
    N.why += std::string(c) + ": " + (e ? e : "?") + "\n";
  }
  if (!N.handle) return nullptr;
  bool ok = true;
  auto sym = [&](const char* name) { void* p = dlsym(N.handle, name); if (!p) { ok = false; N.why += std::string("missing symbol ") + name + "\n"; } return p; };
  N.CreateProgram = (decltype(N.CreateProgram))sym("nvrtcCreateProgram");
  N.DestroyProgram = (decltype(N.DestroyProgram))sym("nvrtcDestroyProgram");
  N.CompileProgram = (decltype(N.CompileProgram))sym("nvrtcCompileProgram");
  N.GetCUBINSize = (decltype(N.GetCUBINSize))sym("nvrtcGetCUBINSize");
  N.GetCUBIN = (decltype(N.GetCUBIN))sym("nvrtcGetCUBIN");
  N.GetProgramLogSize = (decltype(N.GetProgramLogSize))sym("nvrtcGetProgramLogSize");
  N.GetProgramLog = (decltype(N.GetProgramLog))sym("nvrtcGetProgramLog");
  N.GetErrorString = (decltype(N.GetErrorString))sym("nvrtcGetErrorString");
  if (!ok) { dlclose(N.handle); N.handle = nullptr; return nullptr; }
  return &N;
}

thread_local std::string t_log;

struct Entry {
  cudaLibrary_t lib;
  cudaKernel_t kernel;
};
std::unordered_map<std::string, Entry> g_cache;
std::mutex g_cache_mutex;

// what <stdint.h> / <stddef.h> give the generated headers (NVRTC has no host headers)
const char* kStdint =
    "#pragma once\n"
    "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
    "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n"
    "typedef unsigned long uintptr_t;\n";
const char* kStddef = "#pragma once\n";

}  // namespace

void mab_jit_set_log(const std::string& s) { t_log = s; }

int mab_jit_compile(const std::string& source, const MabJitHeader* headers, int nheaders, std::vector<char>* cubin) {
  const Nvrtc* N = nvrtc();
  if (!N) {
    t_log = "NVRTC could not be loaded (set MAB_NVRTC to the path of libnvrtc.so):\n" + g_nvrtc.why;
    return MAB_ERR_NOJIT;
  }
  std::vector<const char*> names, texts;
  for (int i = 0; i < nheaders; i++) { names.push_back(headers[i].name); texts.push_back(headers[i].text); }
  names.push_back("stdint.h"); texts.push_back(kStdint);
  names.push_back("stddef.h"); texts.push_back(kStddef);
  nvrtcProgram prog = nullptr;
  int rc = N->CreateProgram(&prog, source.c_str(), "mab_modprog_jit.cu", (int)names.size(), texts.data(), names.data());
  if (rc != 0) { t_log = std::string("nvrtcCreateProgram: ") + N->GetErrorString(rc); return MAB_ERR_JIT; }
  // -default-device: the generated headers' unannotated helpers (F::name()) are device code here
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-lineinfo"};
  rc = N->CompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts);
  size_t ls = 0;
  t_log.clear();
  if (N->GetProgramLogSize(prog, &ls) == 0 && ls > 1) {
    t_log.resize(ls);
    N->GetProgramLog(prog, &t_log[0]);
  }
  if (rc != 0) {
    t_log = std::string("nvrtcCompileProgram: ") + N->GetErrorString(rc) + "\n" + t_log;
    N->DestroyProgram(&prog);
    return MAB_ERR_JIT;
  }
  size_t cs = 0;
  if ((rc = N->GetCUBINSize(prog, &cs)) != 0 || cs == 0) {
    t_log = std::string("nvrtcGetCUBINSize: ") + N->GetErrorString(rc);
    N->DestroyProgram(&prog);
    return MAB_ERR_JIT;
  }
  cubin->resize(cs);
  rc = N->GetCUBIN(prog, cubin->data());
  N->DestroyProgram(&prog);
  if (rc != 0) { t_log = std::string("nvrtcGetCUBIN: ") + N->GetErrorString(rc); return MAB_ERR_JIT; }
  return 0;
}

int mab_jit_kernel(const std::string& key, std::string (*source)(const void* ctx), const void* ctx,
                   const MabJitHeader* headers, int nheaders, const char* kernel_name, cudaKernel_t* out) {
  std::lock_guard<std::mutex> g(g_cache_mutex);          // compilations are serialised; a hit costs one hash lookup
  auto it = g_cache.find(key);
  if (it != g_cache.end()) { *out = it->second.kernel; return 0; }
  std::vector<char> cubin;
  int rc = mab_jit_compile(source(ctx), headers, nheaders, &cubin);
  if (rc != 0) return rc;
  Entry e;
  cudaError_t ce = cudaLibraryLoadData(&e.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (ce != cudaSuccess) { t_log = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(ce); return (int)ce; }
  ce = cudaLibraryGetKernel(&e.kernel, e.lib, kernel_name);
  if (ce != cudaSuccess) {
    t_log = std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(ce);
    cudaLibraryUnload(e.lib);
    return (int)ce;
  }
  g_cache.emplace(key, e);
  *out = e.kernel;
  return 0;
}

void mab_jit_release(void) {
  std::lock_guard<std::mutex> g(g_cache_mutex);
  for (auto& kv : g_cache) cudaLibraryUnload(kv.second.lib);
  g_cache.clear();
}

extern "C" {
const char* mab_jit_log(void) { return t_log.c_str(); }
}
