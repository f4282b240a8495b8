// mab_runtime.cu -- modulus-independent part of the C ABI: version/params/error strings,
// algorithmic work counts, and the INT32 multiplier-pipe microbenchmark that supplies the
// roofline denominator ("in-repo microbenchmark of dependent-free
// IMAD.WIDE.U32 / IMAD+IMAD.HI streams on the same GPU").
#include <cuda_runtime.h>
#include <string.h>
#include <stdio.h>
#include "modarith_b200.h"
#include "gen/field_X25519.cuh"
#include "gen/field_X448.cuh"
#include "gen/field_NIST256.cuh"
#include "mab_probe.cuh"
#include "mab_workspace.h"
#include "mab_jit.h"
#include "mab_unsat29.cuh"
#include <mutex>

// ------------------------------------------------------------------------------------------
// Each thread runs NCH independent accumulator chains so the 4-5 cycle IMAD latency never
// limits issue; the multiplicands come from memory so nothing constant-folds.
#define NCH 8
template <int VARIANT> __global__ void __launch_bounds__(256) k_imad_peak(const uint32_t* seed, uint32_t* sink, int iters) {
  uint32_t x = seed[threadIdx.x & 31], y = seed[32 + (threadIdx.x & 31)];
  uint32_t lo[NCH], hi[NCH], al[NCH];
#pragma unroll
  for (int c = 0; c < NCH; c++) { lo[c] = x + c; hi[c] = y ^ c; al[c] = x * c; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (VARIANT == 0) {
#pragma unroll
        for (int c = 0; c < NCH; c++)
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(x), "r"(y));
      } else if (VARIANT == 1) {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          // the accumulator is also a multiplicand so that nothing is loop-invariant
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[c]) : "r"(x), "r"(y));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(hi[c]) : "r"(y), "r"(x));
        }
      } else if (VARIANT == 2) {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo[c]) : "r"(x), "r"(y));
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(hi[c]) : "r"(y), "r"(x));
        }
      } else if (VARIANT == 3) {
        // two carry chains of NCH/2 wide multiply-accumulates each, as in the field multiplier
        asm volatile(
            "mad.lo.cc.u32 %0, %16, %17, %0;\n\tmadc.hi.cc.u32 %1, %16, %17, %1;\n\t"
            "madc.lo.cc.u32 %2, %16, %17, %2;\n\tmadc.hi.cc.u32 %3, %16, %17, %3;\n\t"
            "madc.lo.cc.u32 %4, %16, %17, %4;\n\tmadc.hi.cc.u32 %5, %16, %17, %5;\n\t"
            "madc.lo.cc.u32 %6, %16, %17, %6;\n\tmadc.hi.u32 %7, %16, %17, %7;\n\t"
            "mad.lo.cc.u32 %8, %17, %16, %8;\n\tmadc.hi.cc.u32 %9, %17, %16, %9;\n\t"
            "madc.lo.cc.u32 %10, %17, %16, %10;\n\tmadc.hi.cc.u32 %11, %17, %16, %11;\n\t"
            "madc.lo.cc.u32 %12, %17, %16, %12;\n\tmadc.hi.cc.u32 %13, %17, %16, %13;\n\t"
            "madc.lo.cc.u32 %14, %17, %16, %14;\n\tmadc.hi.u32 %15, %17, %16, %15;"
            : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]),
              "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
            : "r"(x), "r"(y));
      } else if (VARIANT == 4 || VARIANT == 5) {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(x), "r"(y));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(al[c]) : "r"(x));
          if (VARIANT == 5) asm volatile("xor.b32 %0, %0, %1;" : "+r"(al[(c + 3) & (NCH - 1)]) : "r"(y));
        }
      } else {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[c]) : "r"(x));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(hi[c]) : "r"(y));
        }
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < NCH; c++) acc ^= lo[c] ^ hi[c] ^ al[c];
  if (acc == 0x12345678u) sink[0] = acc;        // practically never: keeps the chains live
}

template <int V> static void launch_probe(int v, int blocks, int threads, cudaStream_t st, const uint32_t* d, uint32_t* sink, int iters) {
  if (v == V) k_probe<V><<<blocks, threads, 0, st>>>(d, sink, iters);
  if constexpr (V + 1 < MAB_NPROBE) launch_probe<V + 1>(v, blocks, threads, st, d, sink, iters);
}

// ---- host-path workspace cache: one per device, guarded by a mutex held for the whole call ----
static MabWorkspace g_ws[MAB_WS_MAXDEV];
static std::mutex g_ws_mutex[MAB_WS_MAXDEV];

int mab_host_workspace_acquire(int device, size_t bytes, MabWorkspace** out) {
  if (device < 0 || device >= MAB_WS_MAXDEV) return MAB_ERR_BADARG;
  g_ws_mutex[device].lock();
  MabWorkspace* ws = &g_ws[device];
  cudaError_t e = cudaSuccess;
  if (!ws->ready) {
    for (int s = 0; s < MAB_WS_STREAMS && e == cudaSuccess; s++) {
      ws->buf[s] = nullptr;
      e = cudaStreamCreateWithFlags(&ws->stream[s], cudaStreamNonBlocking);
    }
    ws->bytes = 0;
    ws->device = device;
    ws->ready = (e == cudaSuccess);
  }
  if (e == cudaSuccess && ws->bytes < bytes) {
    for (int s = 0; s < MAB_WS_STREAMS && e == cudaSuccess; s++) {
      if (ws->buf[s]) cudaFree(ws->buf[s]);
      ws->buf[s] = nullptr;
      e = cudaMalloc((void**)&ws->buf[s], bytes);
    }
    ws->bytes = (e == cudaSuccess) ? bytes : 0;
  }
  if (e != cudaSuccess) {
    g_ws_mutex[device].unlock();
    return (int)e;
  }
  *out = ws;
  return 0;
}

void mab_host_workspace_release(MabWorkspace* ws) { g_ws_mutex[ws->device].unlock(); }

// ---- work counters of the persistent ladder kernels --------------------------------------------------
#define MAB_NSLOTS 256
struct MabCounterSlot {
  cudaEvent_t done;        // recorded behind the last kernel that used the slot
  bool used;
};
static unsigned long long* g_counters[MAB_WS_MAXDEV];
static MabCounterSlot* g_slots[MAB_WS_MAXDEV];
static unsigned g_counter_next[MAB_WS_MAXDEV];
static std::mutex g_counter_mutex;

int mab_queue_counters(cudaStream_t stream, unsigned nq, unsigned long long** out, int* slot) {
  if (nq == 0 || nq > MAB_QUEUE_MAX) return MAB_ERR_BADARG;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= MAB_WS_MAXDEV) return MAB_ERR_BADARG;
  unsigned id;
  {
    std::lock_guard<std::mutex> g(g_counter_mutex);
    if (!g_counters[dev]) {
      if ((e = cudaMalloc((void**)&g_counters[dev], (size_t)MAB_NSLOTS * MAB_QUEUE_MAX * sizeof(unsigned long long))) != cudaSuccess) return (int)e;
      g_slots[dev] = new MabCounterSlot[MAB_NSLOTS];
      for (int i = 0; i < MAB_NSLOTS; i++) {
        g_slots[dev][i].used = false;
        if ((e = cudaEventCreateWithFlags(&g_slots[dev][i].done, cudaEventDisableTiming)) != cudaSuccess) return (int)e;
      }
    }
    id = g_counter_next[dev]++ % MAB_NSLOTS;
    // the previous user of this slot (MAB_NSLOTS launches ago, possibly on another stream) must have finished
    if (g_slots[dev][id].used && (e = cudaStreamWaitEvent(stream, g_slots[dev][id].done, 0)) != cudaSuccess) return (int)e;
    g_slots[dev][id].used = true;
  }
  unsigned long long* p = g_counters[dev] + (size_t)id * MAB_QUEUE_MAX;
  if ((e = cudaMemsetAsync(p, 0, (size_t)nq * sizeof(unsigned long long), stream)) != cudaSuccess) return (int)e;
  *out = p;
  *slot = (int)id;
  return 0;
}

int mab_queue_counters_launched(int slot, cudaStream_t stream) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= MAB_WS_MAXDEV || slot < 0 || slot >= MAB_NSLOTS || !g_slots[dev]) return MAB_ERR_BADARG;
  std::lock_guard<std::mutex> g(g_counter_mutex);
  return (int)cudaEventRecord(g_slots[dev][slot].done, stream);
}

// ---- stream-ordered scratch pool (table slices of the scalar multiplication kernels) -------------------
static cudaMemPool_t g_pool[MAB_WS_MAXDEV];
static bool g_pool_ready[MAB_WS_MAXDEV];
static std::mutex g_pool_mutex;

int mab_scratch_alloc(void** out, size_t bytes, cudaStream_t stream) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= MAB_WS_MAXDEV) return MAB_ERR_BADARG;
  {
    std::lock_guard<std::mutex> g(g_pool_mutex);
    if (!g_pool_ready[dev]) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      if ((e = cudaMemPoolCreate(&g_pool[dev], &props)) != cudaSuccess) return (int)e;
      unsigned long long keep = ~0ull;
      if ((e = cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) return (int)e;
      g_pool_ready[dev] = true;
    }
  }
  return (int)cudaMallocFromPoolAsync(out, bytes, g_pool[dev], stream);
}

int mab_scratch_free(void* p, cudaStream_t stream) { return (int)cudaFreeAsync(p, stream); }

static const char* kVersion = "modarith_b200 0.1 (sm_100a)";

// W(modmul)=L^2, W(modsqr)=L(L+1)/2, W(modmli)=L with L=ceil(Nbits/32); chains
// counted with the S/M numbers of the chain this build actually emits.
template <class F> static long long products_of(const char* what) {
  const long long L = F::L, M = L * L, S = L * (L + 1) / 2, I = L;
  const long long pro = F::PRO_SQR * S + F::PRO_MUL * M;
  const long long inv_wrap = (F::PM1D2 - 1) * (S + M) + (F::PM1D2 + 1) * S + M;     // pseudo.py:802-810
  if (!strcmp(what, "modmul")) return M;
  if (!strcmp(what, "modsqr")) return S;
  if (!strcmp(what, "modmli")) return I;
  if (!strcmp(what, "modpro")) return pro;
  if (!strcmp(what, "modinv")) return pro + inv_wrap;
  if (!strcmp(what, "modsqrt")) {
    long long w = pro + M;
    if (F::PM1D2 > 1) {
      w += M;
      for (int k = F::PM1D2; k > 1; k--) w += (k - 2) * S + 2 * M + S;
    }
    return w;
  }
  if (!strcmp(what, "rfc7748") && F::HAS_CURVE)
    return (long long)F::NBITS * (5 * M + 4 * S + I) + pro + inv_wrap + M + (F::MONTGOMERY ? 4 * M : 0);
  return -1;
}

extern "C" {

const char* mab_version(void) { return kVersion; }

const char* mab_error_string(int code) {
  if (code == 0) return "ok";
  if (code == MAB_ERR_BADARG) return "modarith_b200: bad argument";
  if (code == MAB_ERR_NODEVICE) return "modarith_b200: no CUDA device";
  if (code == MAB_ERR_NOJIT) return "modarith_b200: NVRTC (libnvrtc.so.12) could not be loaded; see mab_jit_log()";
  if (code == MAB_ERR_JIT) return "modarith_b200: run-time compilation failed; see mab_jit_log()";
  return cudaGetErrorString((cudaError_t)code);
}

void mab_release_workspaces(void) {
  mab_jit_release();
  {
    std::lock_guard<std::mutex> g(g_counter_mutex);
    int prev = 0;
    cudaGetDevice(&prev);
    for (int d = 0; d < MAB_WS_MAXDEV; d++) {
      if (!g_counters[d]) continue;
      cudaSetDevice(d);
      cudaDeviceSynchronize();                       // no kernel may still be drawing from the counters
      for (int i = 0; i < MAB_NSLOTS; i++) cudaEventDestroy(g_slots[d][i].done);
      delete[] g_slots[d];
      g_slots[d] = nullptr;
      cudaFree(g_counters[d]);
      g_counters[d] = nullptr;
      g_counter_next[d] = 0;
    }
    cudaSetDevice(prev);
  }
  {
    std::lock_guard<std::mutex> g(g_pool_mutex);
    for (int d = 0; d < MAB_WS_MAXDEV; d++)
      if (g_pool_ready[d]) {
        cudaMemPoolDestroy(g_pool[d]);
        g_pool_ready[d] = false;
      }
  }
  for (int d = 0; d < MAB_WS_MAXDEV; d++) {
    std::lock_guard<std::mutex> g(g_ws_mutex[d]);
    MabWorkspace* ws = &g_ws[d];
    if (!ws->ready) continue;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d);
    for (int s = 0; s < MAB_WS_STREAMS; s++) {
      if (ws->buf[s]) cudaFree(ws->buf[s]);
      cudaStreamDestroy(ws->stream[s]);
      ws->buf[s] = nullptr;
    }
    ws->bytes = 0;
    ws->ready = false;
    cudaSetDevice(prev);
  }
}

int mab_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int mab_params(const char* prime, int* wordlength, int* nlimbs, int* radix, int* nbits, int* nbytes) {
  int L, nb, by;
  if (!strcmp(prime, "X25519")) { L = F_X25519::L; nb = F_X25519::NBITS; by = F_X25519::NBYTES; }
  else if (!strcmp(prime, "X448")) { L = F_X448::L; nb = F_X448::NBITS; by = F_X448::NBYTES; }
  else if (!strcmp(prime, "NIST256")) { L = F_NIST256::L; nb = F_NIST256::NBITS; by = F_NIST256::NBYTES; }
  else return MAB_ERR_BADARG;
  if (wordlength) *wordlength = 32;
  if (nlimbs) *nlimbs = L;
  if (radix) *radix = 32;
  if (nbits) *nbits = nb;
  if (nbytes) *nbytes = by;
  return 0;
}

long long mab_products(const char* prime, const char* what) {
  if (!strcmp(prime, "X25519")) return products_of<F_X25519>(what);
  if (!strcmp(prime, "X448")) return products_of<F_X448>(what);
  if (!strcmp(prime, "NIST256")) return products_of<F_NIST256>(what);
  return -1;
}

int mab_imad_peak(int variant, int iters, int blocks, int threads, float* ms, double* instructions, void* stream) {
  if (iters <= 0 || blocks <= 0 || threads <= 0 || threads > 256) return MAB_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t host[64];
  for (int i = 0; i < 64; i++) host[i] = 0x9e3779b9u * (i + 1) | 1u;
  uint32_t* d = nullptr;
  cudaError_t e = cudaMalloc((void**)&d, 65 * sizeof(uint32_t));
  if (e != cudaSuccess) return (int)e;
  cudaMemcpy(d, host, sizeof(host), cudaMemcpyHostToDevice);
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  for (int rep = 0; rep < 2; rep++) {          // rep 0 warms up
    cudaEventRecord(t0, st);
    switch (variant) {
      case 0: k_imad_peak<0><<<blocks, threads, 0, st>>>(d, d + 64, iters); break;
      case 1: k_imad_peak<1><<<blocks, threads, 0, st>>>(d, d + 64, iters); break;
      case 2: k_imad_peak<2><<<blocks, threads, 0, st>>>(d, d + 64, iters); break;
      case 3: k_imad_peak<3><<<blocks, threads, 0, st>>>(d, d + 64, iters); break;
      case 4: k_imad_peak<4><<<blocks, threads, 0, st>>>(d, d + 64, iters); break;
      case 5: k_imad_peak<5><<<blocks, threads, 0, st>>>(d, d + 64, iters); break;
      default: k_imad_peak<6><<<blocks, threads, 0, st>>>(d, d + 64, iters); break;
    }
    cudaEventRecord(t1, st);
    cudaEventSynchronize(t1);
  }
  e = cudaGetLastError();
  float t = 0.f;
  cudaEventElapsedTime(&t, t0, t1);
  if (ms) *ms = t;
  // multiply instructions per thread per iteration: 4 unrolls x NCH wide MADs (variants 0,3,4,5),
  // 4 x 2*NCH 32-bit MADs (variants 1,2), 4 x 2*NCH adds (variant 6)
  double per_iter = (variant == 1 || variant == 2 || variant >= 6) ? 4.0 * 2 * NCH : 4.0 * NCH;
  if (instructions) *instructions = per_iter * (double)iters * (double)blocks * (double)threads;
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(d);
  return (int)e;
}

int mab_probe_unsat29_modmul(const uint32_t* a, const uint32_t* b, uint32_t* c, unsigned int iters, size_t n, size_t stride, void* stream) {
  if (n == 0) return 0;
  if (stride < n) return MAB_ERR_BADARG;
  k_unsat29_mulchain<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, b, c, iters, n, stride);
  return (int)cudaGetLastError();
}

int mab_pipe_probe(int variant, int iters, int blocks, int threads, float* ms, const char** name, int* nwide, int* nalu, void* stream) {
  if (variant < 0 || variant >= MAB_NPROBE) return MAB_ERR_BADARG;
  if (iters <= 0 || blocks <= 0 || threads <= 0 || threads > 256) return MAB_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t host[64];
  for (int i = 0; i < 64; i++) host[i] = 0x9e3779b9u * (i + 1) | 1u;
  uint32_t* d = nullptr;
  cudaError_t e = cudaMalloc((void**)&d, 65 * sizeof(uint32_t));
  if (e != cudaSuccess) return (int)e;
  cudaMemcpy(d, host, sizeof(host), cudaMemcpyHostToDevice);
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(t0, st);
    launch_probe<0>(variant, blocks, threads, st, d, d + 64, iters);
    cudaEventRecord(t1, st);
    cudaEventSynchronize(t1);
  }
  e = cudaGetLastError();
  float t = 0.f;
  cudaEventElapsedTime(&t, t0, t1);
  if (ms) *ms = t;
  if (name) *name = kProbeNames[variant];
  if (nwide) *nwide = 0;
  if (nalu) *nalu = 0;
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(d);
  return (int)e;
}

}  // extern "C"
