/**
 * @file lookback.cuh
 * @brief Single-pass device-wide prefix sums by decoupled look-back over per-tile status words.
 * This is the scan under merge-path work preparation and the stable filters; it replaces
 * thrust::transform_exclusive_scan + a blocking D2H of the total (reference advance/helpers.hxx:67-95)
 * and thrust::copy_if's two-kernel select (reference filter/predicated.hxx:29-35): the input is read
 * once and totals stay on the device.
 *
 * Each scanned quantity owns one 64-bit word per tile: [63:62] status, [61:0] value, written/read with
 * single volatile 8-byte accesses, so a reader always sees a consistent (status, value) pair and no
 * fence is needed. Tiles take their index from an atomic ticket, which guarantees every predecessor a
 * tile waits on has already been scheduled (forward progress).
 */
#pragma once

#include <gunrock/b200/warp.cuh>

namespace gunrock {
namespace b200 {

using tile_word_t = unsigned long long;
constexpr tile_word_t tile_value_mask = (tile_word_t(1) << 62) - 1;
constexpr tile_word_t tile_has_aggregate = tile_word_t(1) << 62;
constexpr tile_word_t tile_has_inclusive = tile_word_t(2) << 62;

/**
 * @brief Called by ALL 32 lanes of one warp of the CTA that owns `tile`. Publishes the tile's aggregate,
 * resolves the exclusive prefix of the tile by inspecting up to 32 predecessors per step, publishes the
 * inclusive prefix and returns the exclusive prefix in every lane. `state` must be zeroed before launch.
 */
__device__ __forceinline__ tile_word_t lookback_exclusive(volatile tile_word_t* state, int tile,
                                                          tile_word_t aggregate) {
  const unsigned lane = lane_id();
  if (tile == 0) {
    if (lane == 0) state[0] = tile_has_inclusive | aggregate;
    return 0;
  }
  if (lane == 0) state[tile] = tile_has_aggregate | aggregate;
  tile_word_t exclusive = 0;
  int window_end = tile - 1;  // newest predecessor not yet accounted for
  for (;;) {
    const int idx = window_end - int(lane);
    tile_word_t word = tile_word_t(2) << 62;  // virtual tiles before 0 hold prefix 0
    if (idx >= 0) word = state[idx];  // virtual tiles before 0 hold prefix 0
    while (__any_sync(full_mask, (word >> 62) == 0)) {
      if ((word >> 62) == 0) word = state[idx];
    }
    const unsigned inclusive_at = __ballot_sync(full_mask, (word >> 62) == 2);
    const tile_word_t value = word & tile_value_mask;
    if (inclusive_at) {
      const unsigned first = __ffs(inclusive_at) - 1;  // nearest predecessor holding a full prefix
      exclusive += warp_sum(lane <= first ? value : tile_word_t(0));
      break;
    }
    exclusive += warp_sum(value);
    window_end -= 32;
  }
  if (lane == 0) state[tile] = tile_has_inclusive | ((exclusive + aggregate) & tile_value_mask);
  return exclusive;
}

}  // namespace b200
}  // namespace gunrock
