// host_buffers.cu -- page-locked, huge-page-backed host buffers (include/dynode_b200_host.h).
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/dynode_b200_host.h"

namespace dynode {
int fail_msg(const char* fmt, ...);  // capi.cu
}

namespace {
constexpr size_t kHuge = size_t(2) << 20;
size_t round_up(size_t b) { return (b + kHuge - 1) / kHuge * kHuge; }
}  // namespace

extern "C" int dynode_host_alloc(size_t bytes, uint32_t flags, int32_t threads, void** out) {
  if (!out) return dynode::fail_msg("dynode_host_alloc: out is null");
  *out = nullptr;
  if (bytes == 0) return dynode::fail_msg("dynode_host_alloc: zero bytes");
  const size_t len = round_up(bytes);
  // over-map by one huge page so the base can be aligned to 2 MiB (THP only collapses aligned extents)
  char* raw = (char*)mmap(nullptr, len + kHuge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (raw == (char*)MAP_FAILED) return dynode::fail_msg("dynode_host_alloc: mmap of %zu bytes failed", len);
  char* base = (char*)(((uintptr_t)raw + kHuge - 1) / kHuge * kHuge);
  if (base > raw) munmap(raw, base - raw);
  char* end = base + len;
  char* raw_end = raw + len + kHuge;
  if (raw_end > end) munmap(end, raw_end - end);
  if (flags & DYNODE_HOST_HUGEPAGES) madvise(base, len, MADV_HUGEPAGE);  // advisory: THP may be off, info() tells
  // first touch in parallel: faulting 7.6 GB in from one thread takes seconds
  int nt = threads > 0 ? threads : 1;
  if (nt > 64) nt = 64;
  const size_t per = round_up((len + nt - 1) / nt);
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t) {
    const size_t lo = (size_t)t * per, hi = lo + per < len ? lo + per : len;
    if (lo >= hi) break;
    pool.emplace_back([=] {
      for (size_t o = lo; o < hi; o += 4096) base[o] = 0;
    });
  }
  for (auto& th : pool) th.join();
  if (!(flags & DYNODE_HOST_NO_PIN)) {
    cudaError_t e = cudaHostRegister(base, len, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
      munmap(base, len);
      return dynode::fail_msg("dynode_host_alloc: cudaHostRegister(%zu bytes): %s", len, cudaGetErrorString(e));
    }
  }
  *out = base;
  return 0;
}

extern "C" int dynode_host_free(void* ptr, size_t bytes, uint32_t flags) {
  if (!ptr) return 0;
  if (!(flags & DYNODE_HOST_NO_PIN)) {
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess && e != cudaErrorCudartUnloading)
      return dynode::fail_msg("dynode_host_free: cudaHostUnregister: %s", cudaGetErrorString(e));
  }
  if (munmap(ptr, round_up(bytes)) != 0) return dynode::fail_msg("dynode_host_free: munmap failed");
  return 0;
}

extern "C" int64_t dynode_host_info(const void* ptr, size_t bytes) {
  // sum AnonHugePages of the smaps entries inside [ptr, ptr + bytes)
  FILE* f = fopen("/proc/self/smaps", "r");
  if (!f) return -1;
  const uintptr_t lo = (uintptr_t)ptr, hi = lo + round_up(bytes);
  char line[512];
  bool inside = false;
  int64_t kb = 0;
  while (fgets(line, sizeof(line), f)) {
    unsigned long a, b;
    // mapping headers start with a lower-case hex address ("7f12..-7f13.. rw-p"), attribute lines with a capital
    const bool header = (line[0] >= '0' && line[0] <= '9') || (line[0] >= 'a' && line[0] <= 'f');
    if (header) {
      inside = sscanf(line, "%lx-%lx ", &a, &b) == 2 && a < hi && b > lo;
      continue;
    }
    long v;
    if (inside && sscanf(line, "AnonHugePages: %ld kB", &v) == 1) kb += v;
  }
  fclose(f);
  return kb * 1024;
}
