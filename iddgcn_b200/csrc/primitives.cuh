// primitives.cuh — device-wide exclusive scan and stable LSD radix sort, hand-written.
//
// Both are deterministic (no inter-warp atomics decide placement), which is what makes the
// CSR/CSC construction bit-reproducible.  The radix sort assigns every warp one contiguous
// chunk of the input; a pass is: per-warp digit histogram -> exclusive scan of the
// [digit][warp] matrix -> each warp re-walks its chunk in order and ranks the 32 keys of a
// step with __match_any_sync.  Stability follows from (chunk order) x (step order) x (lane order).
#pragma once
#include "common.cuh"

namespace iddgcn {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// scratch sized for scans up to `n` elements
size_t scan_scratch_elems(size_t n);
// out[i] = sum_{j<i} in[j]; out may alias in.  scratch: >= scan_scratch_elems(n) uint32.
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* scratch, cudaStream_t st);

struct RadixSortPlan {
  int num_warps = 0;
  size_t chunk = 0;       // keys per warp (multiple of 32)
  size_t hist_elems = 0;  // 256 * num_warps
  size_t scratch_elems = 0;
};
RadixSortPlan radix_sort_plan(size_t n);

// Sorts n 64-bit keys (optionally with 32-bit payloads) by bits [0, num_bits).
// keys_a/vals_a hold the input; keys_b/vals_b are same-sized ping-pong buffers.
// On return *keys_sorted / *vals_sorted point at whichever buffer holds the result.
// hist: >= plan.hist_elems uint32; scratch: >= plan.scratch_elems uint32.
int radix_sort_u64(uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, size_t n,
                   int num_bits, const RadixSortPlan& plan, uint32_t* hist, uint32_t* scratch,
                   cudaStream_t st, uint64_t** keys_sorted, uint32_t** vals_sorted);

}  // namespace iddgcn
