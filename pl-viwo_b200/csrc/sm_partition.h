// SM partitioning with CUDA green contexts (driver API, CUDA 12.4+): the latency-critical tracking stream of every handle
// gets a partition of the device's SMs of its own, all the frame-independent work (equalise, pyramid, FAST, the line
// detector's connected components and chain walk) runs on the remaining SMs.  An LK launch then never waits for an SM
// slot behind the many long-running kernels of the frames that are being prepared ahead of the trackers.
// Opt-in: PLVIWO_LK_SMS = n (SMs of the tracking partition; 0 / unset = one context, priorities only).
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace plviwo {

enum class SmPart { Tracking, Rest };

// Creates a non-blocking stream with the given priority on the requested partition of `device` (the calling thread's
// current device).  Falls back to an ordinary stream when partitioning is off or unavailable; *partitioned (optional)
// tells which it was.  Returns a cudaError_t.
cudaError_t create_stream_on_partition(int device, SmPart part, int priority, cudaStream_t *out, bool *partitioned = nullptr);
// SMs of the two partitions on `device` (0, 0 when partitioning is off)
void partition_sm_counts(int device, int *tracking, int *rest);

}  // namespace plviwo
