// mab_jit.h -- run-time compilation of straight-line field programs (internal; see mab_<P>_modprog_jit in
// include/modarith_b200.h).  The reference's build model is "run the generator, compile what it printed"
// (pseudo.py:1694-1702, 1895-1903) and its consumers paste the generated functions into their own source
// (rfc7748.c:24-28, weierstrass.c:16-20): the JIT does the same for a call sequence handed over at run time --
// it prints a kernel that calls the generated functions on variables in machine registers, compiles it with
// NVRTC for sm_100a against the very headers the library was built from (embedded as text), and keeps the
// loaded kernel in a cache keyed by the program.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <string>
#include <vector>

struct MabJitHeader {
  const char* name;      // the name an #include directive uses
  const char* text;
};

// Compiles `source` (which may #include the given headers by name) to an sm_100a cubin.  Returns 0,
// MAB_ERR_NOJIT when NVRTC cannot be loaded, MAB_ERR_JIT when the compilation fails; the compiler's log is
// kept per thread for mab_jit_log().
int mab_jit_compile(const std::string& source, const MabJitHeader* headers, int nheaders, std::vector<char>* cubin);

// The loaded kernel `kernel_name` of the program identified by `key`: compiled on first use (by calling
// `source()`), then served from the cache.  Kernels are context-independent (cudaLibrary_t), so one entry serves
// every device.
int mab_jit_kernel(const std::string& key, std::string (*source)(const void* ctx), const void* ctx,
                   const MabJitHeader* headers, int nheaders, const char* kernel_name, cudaKernel_t* out);

void mab_jit_set_log(const std::string& s);
void mab_jit_release(void);            // unloads every cached program (mab_release_workspaces)
