// mab_workspace.h -- per-device staging area of the host-pointer entry points (internal).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#define MAB_WS_STREAMS 3
#define MAB_WS_MAXDEV 64
struct MabWorkspace {
  cudaStream_t stream[MAB_WS_STREAMS];
  char* buf[MAB_WS_STREAMS];
  size_t bytes;          // capacity of each buf
  int device;
  bool ready;
};
// Locks the device's workspace, growing each of its three buffers to at least `bytes`.
int mab_host_workspace_acquire(int device, size_t bytes, MabWorkspace** out);
void mab_host_workspace_release(MabWorkspace* ws);
// Zeroed 64-bit work counters (up to MAB_QUEUE_MAX) for one launch of a persistent kernel.  Slots of a
// per-device pool are handed out round-robin and cleared on `stream` ahead of the launch; every slot carries
// an event recorded after the launch that used it last (mab_queue_counters_launched), and a stream that gets
// the slot again first waits for that event -- so a launch can never clear counters another stream's kernel
// is still drawing from, however many launches are outstanding.
#define MAB_QUEUE_MAX 1024
int mab_queue_counters(cudaStream_t stream, unsigned nq, unsigned long long** out, int* slot);
int mab_queue_counters_launched(int slot, cudaStream_t stream);
// Stream-ordered scratch memory from a pool the library owns (one per device, created on first use, keeps
// what is freed so that repeated calls do not reach the driver; destroyed by mab_release_workspaces).
int mab_scratch_alloc(void** out, size_t bytes, cudaStream_t stream);
int mab_scratch_free(void* p, cudaStream_t stream);
