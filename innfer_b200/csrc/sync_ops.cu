// Stream-ordered cross-GPU signalling for the tile-sharded mode: monotonically increasing 32-bit
// counters in device memory that peers map through CUDA IPC.  A signal is a system-scope release store issued by a
// one-thread kernel AFTER everything enqueued earlier on the stream (e.g. the last conv's peer stores over NVLink) has
// completed; a wait is a kernel that spins with system-scope acquire loads until every watched counter has reached
// the value, so that work enqueued behind it on the stream sees the data the signaller published.  The reference has
// no equivalent (it is single-GPU: the tile loop of run.py:187-197 is serial); this replaces two host barriers per
// frame.  A wait that does not complete within `timeout_ns` raises the error word instead of hanging the GPU.
#include "sync_ops.cuh"

namespace innfer {

namespace {

struct FlagList {
  uint32_t* p[kMaxSyncFlags];
  int n;
};

__global__ void signal_kernel(const FlagList f, uint32_t value) {
  const int i = threadIdx.x;
  if (i >= f.n) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[i]), "r"(value) : "memory");
}

__global__ void wait_kernel(const FlagList f, uint32_t value, uint32_t* err, unsigned long long timeout_ns) {
  const int i = threadIdx.x;
  if (i >= f.n) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f.p[i]) : "memory");
    // counters only grow; the signed difference keeps the comparison valid across a wrap
    if ((int32_t)(v - value) >= 0) break;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) {
      if (err) atomicAdd(err, 1u);
      break;
    }
    __nanosleep(256);
  }
  __threadfence_system();
}

}  // namespace

int launch_signal(uint32_t* const* flags, int n, uint32_t value, cudaStream_t stream) {
  if (n < 1 || n > kMaxSyncFlags) return -1;
  FlagList f;
  f.n = n;
  for (int i = 0; i < n; ++i) f.p[i] = flags[i];
  signal_kernel<<<1, kMaxSyncFlags, 0, stream>>>(f, value);
  return (int)cudaGetLastError();
}

int launch_wait(uint32_t* const* flags, int n, uint32_t value, uint32_t* err, unsigned long long timeout_ns,
                cudaStream_t stream) {
  if (n < 1 || n > kMaxSyncFlags) return -1;
  FlagList f;
  f.n = n;
  for (int i = 0; i < n; ++i) f.p[i] = flags[i];
  wait_kernel<<<1, kMaxSyncFlags, 0, stream>>>(f, value, err, timeout_ns);
  return (int)cudaGetLastError();
}

}  // namespace innfer
