/**
 * @file warp.cuh
 * @brief Warp/CTA building blocks of the sm_100a operators: shuffle scans and reductions, ballot
 * compaction, CTA-wide exclusive scan, and aggregated appends to a global queue. These replace the
 * cub::BlockScan / thrust scan / moderngpu calls on the reference's hot path
 * (advance/block_mapped.hxx:86, advance/helpers.hxx:67-76, filter/predicated.hxx:29-35).
 * All kernels in this tree use 1-D CTAs whose size is a multiple of 32.
 */
#pragma once

#include <cstdint>

namespace gunrock {
namespace b200 {

constexpr unsigned full_mask = 0xffffffffu;
using counter_t = unsigned long long;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ unsigned lanes_below(unsigned lane) { return (1u << lane) - 1u; }

template <typename T>
__device__ __forceinline__ T warp_inclusive_sum(T x) {
  const unsigned lane = lane_id();
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T y = __shfl_up_sync(full_mask, x, d);
    if (lane >= unsigned(d)) x += y;
  }
  return x;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T x) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(full_mask, x, d);
  return x;
}

template <typename T>
__device__ __forceinline__ T warp_max(T x) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    T y = __shfl_xor_sync(full_mask, x, d);
    x = y > x ? y : x;
  }
  return x;
}

/**
 * @brief CTA-wide exclusive sum. `warp_totals` is shared memory with THREADS/32 + 1 slots of T.
 * Every thread of the CTA must call it. Returns the exclusive prefix of `x`; `total` gets the CTA sum.
 * Ends with a barrier-protected read, so `warp_totals` may be reused after one more __syncthreads().
 */
template <int THREADS, typename T>
__device__ __forceinline__ T cta_exclusive_sum(T x, T& total, T* warp_totals) {
  constexpr int WARPS = THREADS / 32;
  const unsigned lane = lane_id(), warp = warp_id();
  T incl = warp_inclusive_sum(x);
  if (lane == 31) warp_totals[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = lane < WARPS ? warp_totals[lane] : T(0);
    T wi = warp_inclusive_sum(w);
    if (lane < WARPS) warp_totals[lane] = wi - w;
    if (lane == WARPS - 1) warp_totals[WARPS] = wi;
  }
  __syncthreads();
  total = warp_totals[WARPS];
  return warp_totals[warp] + incl - x;
}

/**
 * @brief Warp-aggregated append: the whole warp calls it (converged); lanes with `keep` get a unique slot
 * in the queue whose length lives at `counter`. One atomic per warp, none if nobody keeps.
 */
__device__ __forceinline__ counter_t warp_append_slot(bool keep, counter_t* counter) {
  const unsigned votes = __ballot_sync(full_mask, keep);
  if (votes == 0) return 0;
  const unsigned lane = lane_id();
  const int leader = __ffs(votes) - 1;
  counter_t base = 0;
  if (int(lane) == leader) base = atomicAdd(counter, counter_t(__popc(votes)));
  base = __shfl_sync(full_mask, base, leader);
  return base + __popc(votes & lanes_below(lane));
}

/**
 * @brief CTA-aggregated append of up to ITEMS kept values per thread: one global atomic per CTA call.
 * `keep_bits` bit i says whether vals[i] is kept. `smem` needs THREADS/32 + 2 counter_t slots.
 * Every thread of the CTA must call it. Returns how many the CTA appended (uniform).
 */
template <int THREADS, int ITEMS, typename T>
__device__ __forceinline__ unsigned cta_append(const T (&vals)[ITEMS], unsigned keep_bits, T* queue,
                                               counter_t* counter, counter_t capacity, counter_t* smem) {
  constexpr int WARPS = THREADS / 32;
  unsigned mine = __popc(keep_bits);
  unsigned total;
  unsigned* s = reinterpret_cast<unsigned*>(smem + 1);  // WARPS+1 unsigned slots after the base slot
  unsigned before = cta_exclusive_sum<THREADS, unsigned>(mine, total, s);
  if (total == 0) return 0;  // uniform
  if (threadIdx.x == 0) smem[0] = atomicAdd(counter, counter_t(total));
  __syncthreads();
  counter_t at = smem[0] + before;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i)
    if (keep_bits & (1u << i)) {
      if (at < capacity) queue[at] = vals[i];
      ++at;
    }
  __syncthreads();  // smem reusable by the caller
  (void)WARPS;
  return total;
}

/// Largest j in [0, count) with table[j] <= key (table ascending, table[0] <= key assumed).
template <typename T, typename K>
__device__ __forceinline__ int upper_segment(const T* table, int count, K key) {
  int lo = 0, hi = count;  // invariant: table[lo] <= key, table[hi] > key (virtual)
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (K(table[mid]) <= key)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

}  // namespace b200
}  // namespace gunrock
