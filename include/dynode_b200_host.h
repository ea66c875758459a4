/* dynode_b200_host.h -- page-locked host buffers for the ensemble's host<->device stream (C ABI).
 *
 * What this replaces: in the reference `Solution.ys` lands in host memory because the solve runs on the host
 * (src/dynode/simulation/odes.py:133-144 returns diffrax's arrays).  Here every saved value crosses PCIe, so the
 * buffer the copy engine writes into decides the end-to-end rate: it must be page-locked (DMA without a staging
 * copy), and backing it with 2 MiB transparent huge pages keeps the IOMMU / DMA page tables 512x smaller than
 * cudaHostAlloc's 4 KiB pages do.
 *
 *   dynode_host_alloc   anonymous mapping, madvise(MADV_HUGEPAGE) when asked, faulted in by `threads` host threads
 *                       (first touch), then cudaHostRegister(portable).  *out = base address, 2 MiB aligned.
 *   dynode_host_free    cudaHostUnregister + munmap; `bytes` as passed to alloc.
 *   dynode_host_info    how the mapping ended up: bytes backed by huge pages (AnonHugePages in smaps), -1 unknown.
 *
 * Return 0 on success; message via dynode_last_error().  No global state: the caller owns the buffer.
 */
#ifndef DYNODE_B200_HOST_H
#define DYNODE_B200_HOST_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DYNODE_HOST_HUGEPAGES 1u  /* madvise(MADV_HUGEPAGE) before the first touch */
#define DYNODE_HOST_NO_PIN 2u     /* do not cudaHostRegister (host-only tests) */

int dynode_host_alloc(size_t bytes, uint32_t flags, int32_t threads, void** out);
int dynode_host_free(void* ptr, size_t bytes, uint32_t flags);
int64_t dynode_host_info(const void* ptr, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif
