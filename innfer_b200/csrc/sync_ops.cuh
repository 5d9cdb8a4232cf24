// Stream-ordered device flags for the tile-sharded multi-GPU mode (sync_ops.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace innfer {

constexpr int kMaxSyncFlags = 32;

// store `value` (system-scope release) into each of the n counters, after all earlier work of `stream`
int launch_signal(uint32_t* const* flags, int n, uint32_t value, cudaStream_t stream);
// hold `stream` until every counter is >= value; on timeout *err is incremented and the stream continues
int launch_wait(uint32_t* const* flags, int n, uint32_t value, uint32_t* err, unsigned long long timeout_ns,
                cudaStream_t stream);

}  // namespace innfer
