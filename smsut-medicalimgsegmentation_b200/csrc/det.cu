// Deterministic accumulation: registry of fixed-point shadows and the resolve kernel (see common.cuh).
//
// The arithmetic it makes reproducible is the reference's own: sums over H*W for nn.InstanceNorm2d
// (network/blocks.py:22-23), weight gradients of aten::convolution_backward, the batch-wide Dice statistics
// (misc/loss.py:52-63) and the scalar loss means (trainer/uganConsisTrainer.py:129-177).
#include <mutex>
#include <vector>

#include "../../include/smsut_b200.h"
#include "common.cuh"

namespace smsut {

void count_launch();

struct DetRange {
  const char* base;
  size_t bytes;
  long long* shadow;
};
static std::mutex g_det_mu;
static std::vector<DetRange> g_det_ranges;

long long* det_shadow(const void* p) {
  if (p == nullptr) return nullptr;
  std::lock_guard<std::mutex> lk(g_det_mu);
  const char* c = reinterpret_cast<const char*>(p);
  for (const DetRange& r : g_det_ranges)
    if (c >= r.base && c < r.base + r.bytes) return r.shadow + (c - r.base) / 4;
  return nullptr;
}

// dst[i] += shadow[i] * 2^-32; shadow[i] = 0
__global__ void det_resolve_kernel(float* __restrict__ dst, long long* __restrict__ q, long long count) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    if (q[i] != 0) {
      // take-and-clear in one atomic: an add that lands concurrently stays pending instead of being lost
      const long long v = (long long)atomicExch(reinterpret_cast<unsigned long long*>(q + i), 0ull);
      dst[i] += (float)((double)v * (1.0 / 4294967296.0));
    }
  }
}

}  // namespace smsut

using namespace smsut;

extern "C" int smsut_det_register(const void* base, size_t bytes, void* shadow) {
  SMSUT_CHECK(base != nullptr && shadow != nullptr && bytes >= 4 && bytes % 4 == 0, -1, "det_register: bad range");
  SMSUT_CHECK((reinterpret_cast<uintptr_t>(shadow) & 7) == 0 && (reinterpret_cast<uintptr_t>(base) & 3) == 0, -1,
              "det_register: misaligned pointers");
  std::lock_guard<std::mutex> lk(g_det_mu);
  const char* b = reinterpret_cast<const char*>(base);
  for (size_t i = 0; i < g_det_ranges.size();) {       // a new registration replaces whatever it overlaps
    const DetRange& r = g_det_ranges[i];
    if (b < r.base + r.bytes && r.base < b + bytes) g_det_ranges.erase(g_det_ranges.begin() + i);
    else ++i;
  }
  g_det_ranges.push_back(DetRange{b, bytes, reinterpret_cast<long long*>(shadow)});
  return 0;
}

extern "C" int smsut_det_unregister(const void* base) {
  std::lock_guard<std::mutex> lk(g_det_mu);
  for (size_t i = 0; i < g_det_ranges.size(); ++i)
    if (g_det_ranges[i].base == reinterpret_cast<const char*>(base)) {
      g_det_ranges.erase(g_det_ranges.begin() + i);
      return 0;
    }
  return 0;
}

extern "C" int smsut_det_ranges(void) {
  std::lock_guard<std::mutex> lk(g_det_mu);
  return (int)g_det_ranges.size();
}

extern "C" void* smsut_det_shadow(const void* p) { return det_shadow(p); }

extern "C" int smsut_det_resolve(float* dst, int64_t count, smsut_stream_t st) {
  SMSUT_CHECK(dst != nullptr && count > 0, -1, "det_resolve: empty destination");
  long long* q = det_shadow(dst);
  if (q == nullptr) return 0;      // not a registered accumulator: its fp32 atomics landed in place
  SMSUT_CHECK(det_shadow(dst + (count - 1)) == q + (count - 1), -1, "det_resolve: range crosses a registration");
  long long blocks = (count + 255) / 256;
  const long long cap = 8LL * device_sm_count();
  if (blocks > cap) blocks = cap;
  launch_pdl(det_resolve_kernel, dim3((unsigned)blocks), 256, 0, (cudaStream_t)st, dst, q, (long long)count);
  count_launch();
  return launch_status("det_resolve_kernel");
}
