#include "sm_partition.h"

#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace plviwo {

namespace {

struct DevicePartition {
  bool tried = false, ok = false;
  CUgreenCtx ctx[2] = {nullptr, nullptr};   // [Tracking, Rest]
  int sms[2] = {0, 0};
};
DevicePartition g_part[64];
std::mutex g_mu;

// driver entry points through the runtime: no link-time dependency on libcuda
template <class F>
bool entry(const char *name, F *fn) {
  cudaDriverEntryPointQueryResult q;
  void *p = nullptr;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}

bool build(int device, int want, DevicePartition &d) {
  decltype(&cuDeviceGet) f_dev = nullptr;
  decltype(&cuDeviceGetDevResource) f_res = nullptr;
  decltype(&cuDevSmResourceSplitByCount) f_split = nullptr;
  decltype(&cuDevResourceGenerateDesc) f_desc = nullptr;
  decltype(&cuGreenCtxCreate) f_create = nullptr;
  if (!entry("cuDeviceGet", &f_dev) || !entry("cuDeviceGetDevResource", &f_res) || !entry("cuDevSmResourceSplitByCount", &f_split) ||
      !entry("cuDevResourceGenerateDesc", &f_desc) || !entry("cuGreenCtxCreate", &f_create))
    return false;
  cudaFree(nullptr);   // the primary context must exist
  CUdevice dev;
  if (f_dev(&dev, device) != CUDA_SUCCESS) return false;
  CUdevResource all, group, rest;
  if (f_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
  unsigned nb = 1;
  if (f_split(&group, &nb, &all, &rest, 0, (unsigned)want) != CUDA_SUCCESS || nb != 1) return false;
  if (group.sm.smCount == 0 || rest.sm.smCount == 0) return false;
  CUdevResource parts[2] = {group, rest};
  for (int k = 0; k < 2; k++) {
    CUdevResourceDesc desc;
    if (f_desc(&desc, &parts[k], 1) != CUDA_SUCCESS) return false;
    if (f_create(&d.ctx[k], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    d.sms[k] = (int)parts[k].sm.smCount;
  }
  return true;
}

DevicePartition *get(int device) {
  static const int want = [] {
    const char *e = std::getenv("PLVIWO_LK_SMS");
    return e ? std::atoi(e) : 0;
  }();
  if (want <= 0 || device < 0 || device >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  DevicePartition &d = g_part[device];
  if (!d.tried) {
    d.tried = true;
    d.ok = build(device, want, d);
    if (!d.ok) std::fprintf(stderr, "plviwo: SM partitioning (PLVIWO_LK_SMS=%d) is not available on device %d, running unpartitioned\n", want, device);
  }
  return d.ok ? &d : nullptr;
}

}  // namespace

cudaError_t create_stream_on_partition(int device, SmPart part, int priority, cudaStream_t *out, bool *partitioned) {
  if (partitioned) *partitioned = false;
  DevicePartition *d = get(device);
  if (d) {
    decltype(&cuGreenCtxStreamCreate) f_stream = nullptr;
    if (entry("cuGreenCtxStreamCreate", &f_stream)) {
      CUstream st = nullptr;
      if (f_stream(&st, d->ctx[part == SmPart::Tracking ? 0 : 1], CU_STREAM_NON_BLOCKING, priority) == CUDA_SUCCESS) {
        *out = reinterpret_cast<cudaStream_t>(st);
        if (partitioned) *partitioned = true;
        return cudaSuccess;
      }
    }
  }
  return cudaStreamCreateWithPriority(out, cudaStreamNonBlocking, priority);
}

void partition_sm_counts(int device, int *tracking, int *rest) {
  DevicePartition *d = get(device);
  if (tracking) *tracking = d ? d->sms[0] : 0;
  if (rest) *rest = d ? d->sms[1] : 0;
}

}  // namespace plviwo
