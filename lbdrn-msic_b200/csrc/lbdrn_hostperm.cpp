// lbdrn_hostperm.cpp -- the reference sampler's batch order on the host, faster.
//
// encode.py:69-70 draws every epoch's batch order with DataLoader(shuffle=True): RandomSampler calls
// torch.randperm(n, generator=g) on a CPU generator seeded from the default generator.  torch's CPU randperm for
// n < 2^32 / 20 is a forward Fisher-Yates shuffle driven by the generator's 32-bit outputs (aten/src/ATen/native/
// TensorFactories.cpp: randperm_cpu):
//     r[i] = i;   for i in [0, n-1):  z = random() % (n - i);  swap(r[i], r[i + z])
// with random() = one output of MT19937 seeded by the low 32 bits of the seed (ATen/core/MT19937RNGEngine.h; the standard
// generator, std::mt19937).  For the 67 M pixels of an 8192^2 scene that loop takes ~3.2 s in torch -- 20x the epoch it
// feeds on a B200 -- because every swap is a cache miss on a 537 MB array.  The draws do not depend on the array, so this
// restatement runs them LA iterations ahead of the swaps and prefetches the line each swap will touch; same outputs, the
// misses overlap.  Bit-exactness against torch.randperm is pinned by tests/test_host_logic.py (CPU test, several n and
// seeds); larger n (torch switches to a 64-bit inside-out variant there) are left to torch itself.
#include <stdint.h>
#include <stdlib.h>
#include <sys/mman.h>

#include <atomic>
#include <random>
#include <thread>
#include <vector>

#include "../../include/lbdrn.h"

namespace {

template <typename T>
int32_t host_randperm_t(int64_t n, uint64_t seed, T* out_host, int64_t* progress = nullptr) {
  if (n < 0 || (n > 0 && out_host == nullptr)) return LBDRN_E_INVALID;
  if (n >= (int64_t)(UINT32_MAX / 20)) return LBDRN_E_UNSUPPORTED;     // torch's other branch (random64, inside-out)
  // every swap is a TLB miss as well on 4 KB pages (65 k pages for an 8192^2 scene): ask for huge pages before the first
  // touch of the (usually fresh) buffer; advisory -- the call failing or the pages having been touched changes nothing
  {
    constexpr uintptr_t HP = (uintptr_t)2 << 20;
    const uintptr_t b = ((uintptr_t)out_host + HP - 1) & ~(HP - 1), e = ((uintptr_t)(out_host + n)) & ~(HP - 1);
    if (e > b && getenv("LBDRN_NO_THP") == nullptr) (void)madvise((void*)b, (size_t)(e - b), MADV_HUGEPAGE);
  }
  for (int64_t i = 0; i < n; ++i) out_host[i] = (T)i;
  if (n < 2) {
    if (progress) __atomic_store_n(progress, n, __ATOMIC_RELEASE);
    return LBDRN_OK;
  }
  std::mt19937 eng((uint32_t)(seed & 0xffffffffu));
  constexpr int64_t LA = 128;                    // look-ahead (iterations): ~LA independent misses in flight
  uint32_t ring[LA];
  const int64_t steps = n - 1;
  const int64_t warm = steps < LA ? steps : LA;
  for (int64_t i = 0; i < warm; ++i) {
    ring[i] = (uint32_t)eng() % (uint32_t)(n - i);
    __builtin_prefetch(out_host + i + ring[i], 0, 2);
  }
  for (int64_t i = 0; i < steps; ++i) {
    // entries [0, i) are final (later swaps touch positions >= i only): published every 2^18 steps for a consumer that
    // starts on the head of the order while the tail is still being shuffled
    if (progress && (i & 0x3ffff) == 0) __atomic_store_n(progress, i, __ATOMIC_RELEASE);
    const uint32_t z = ring[i & (LA - 1)];
    const int64_t j = i + LA;
    if (j < steps) {
      const uint32_t zj = (uint32_t)eng() % (uint32_t)(n - j);
      ring[i & (LA - 1)] = zj;
      __builtin_prefetch(out_host + j + zj, 0, 2);     // read intent into L2: measured 0.62 s vs 0.88 s (write, non-temporal) for 67 M swaps
    }
    const T sav = out_host[i];
    out_host[i] = out_host[i + z];
    out_host[i + z] = sav;
  }
  if (progress) __atomic_store_n(progress, n, __ATOMIC_RELEASE);
  return LBDRN_OK;
}

// The same shuffle with the draws on a second thread (the order nothing can hide -- epoch 1's -- is worth two cores): the
// generator and its n - i remainders run in a producer that stays up to kBlocks blocks ahead in a ring; this thread only
// swaps (and prefetches kLA swaps ahead inside the blocks already produced).  Same draws in the same order: same output.
template <typename T>
int32_t host_randperm_two_threads(int64_t n, uint64_t seed, T* out_host, int64_t* progress) {
  constexpr int64_t kBS = 1 << 16, kBlocks = 16, kMask = kBS * kBlocks - 1, kLA = 128;
  {
    constexpr uintptr_t HP = (uintptr_t)2 << 20;
    const uintptr_t b = ((uintptr_t)out_host + HP - 1) & ~(HP - 1), e = ((uintptr_t)(out_host + n)) & ~(HP - 1);
    if (e > b && getenv("LBDRN_NO_THP") == nullptr) (void)madvise((void*)b, (size_t)(e - b), MADV_HUGEPAGE);
  }
  const int64_t steps = n - 1, nblk = (steps + kBS - 1) / kBS;
  std::vector<uint32_t> ring((size_t)(kBS * kBlocks));
  std::atomic<int64_t> produced{0}, consumed{0};
  std::thread producer([&]() {
    std::mt19937 eng((uint32_t)(seed & 0xffffffffu));
    for (int64_t blk = 0; blk < nblk; ++blk) {
      while (blk - consumed.load(std::memory_order_acquire) >= kBlocks) std::this_thread::yield();
      const int64_t i0 = blk * kBS, i1 = i0 + kBS < steps ? i0 + kBS : steps;
      uint32_t* z = ring.data() + (i0 & kMask);
      for (int64_t i = i0; i < i1; ++i) z[i - i0] = (uint32_t)eng() % (uint32_t)(n - i);
      produced.store(blk + 1, std::memory_order_release);
    }
  });
  for (int64_t i = 0; i < n; ++i) out_host[i] = (T)i;        // under the producer's first blocks
  for (int64_t blk = 0; blk < nblk; ++blk) {
    const int64_t want = blk + 2 < nblk ? blk + 2 : nblk;     // this block and the one the look-ahead reaches into
    while (produced.load(std::memory_order_acquire) < want) std::this_thread::yield();
    const int64_t i0 = blk * kBS, i1 = i0 + kBS < steps ? i0 + kBS : steps;
    for (int64_t i = i0; i < i1; ++i) {
      if ((i & 0x3ffff) == 0) __atomic_store_n(progress, i, __ATOMIC_RELEASE);
      const int64_t j = i + kLA;
      if (j < steps) __builtin_prefetch(out_host + j + ring[(size_t)(j & kMask)], 0, 2);
      const uint32_t z = ring[(size_t)(i & kMask)];
      const T sav = out_host[i];
      out_host[i] = out_host[i + z];
      out_host[i + z] = sav;
    }
    consumed.store(blk + 1, std::memory_order_release);
  }
  producer.join();
  __atomic_store_n(progress, n, __ATOMIC_RELEASE);
  return LBDRN_OK;
}

}  // namespace

extern "C" int32_t lbdrn_host_randperm(int64_t n, uint64_t seed, int64_t* out_host) { return host_randperm_t(n, seed, out_host); }

// the same permutation as 32-bit indices (every n in range fits): half the memory traffic of the shuffle and of the upload
extern "C" int32_t lbdrn_host_randperm32(int64_t n, uint64_t seed, int32_t* out_host) { return host_randperm_t(n, seed, out_host); }

// the 32-bit order with its progress published: *progress_host = number of leading entries that are final (n at the end)
extern "C" int32_t lbdrn_host_randperm32_progress(int64_t n, uint64_t seed, int32_t* out_host, int64_t* progress_host) {
  if (progress_host == nullptr) return LBDRN_E_INVALID;
  __atomic_store_n(progress_host, (int64_t)0, __ATOMIC_RELEASE);
  if (n >= ((int64_t)1 << 22) && n < (int64_t)(UINT32_MAX / 20) && out_host != nullptr && getenv("LBDRN_PERM_ONE_THREAD") == nullptr &&
      std::thread::hardware_concurrency() >= 4)
    return host_randperm_two_threads(n, seed, out_host, progress_host);
  return host_randperm_t(n, seed, out_host, progress_host);
}
