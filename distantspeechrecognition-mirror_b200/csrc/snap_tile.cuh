// snap_tile.cuh -- lane-group walk over stored snapshots [F][B][C] for the per-(frame, bin) kernels.
// One (frame, bin) item is C consecutive complex values.  A group of GS lanes (GS = C rounded up to a power of two, at
// most 32) owns one item at a time, lane lg of the group taking channels lg, lg + GS, ...: consecutive lanes read
// consecutive 8-byte words (groups of a warp sit on consecutive items, so a warp request covers one contiguous span
// when C is a power of two), there is no shared-memory staging and no CTA barrier, and per-item results are combined
// with group-wide shuffles.
#pragma once

#include "fb_core.cuh"

namespace btk {

#define SNAP_THREADS 256

__host__ __device__ inline int snap_group_size(int C) {
  int g = 1;
  while (g < C && g < 32) g <<= 1;
  return g;
}

// sum over the GS lanes of a group (GS a power of two <= 32); every lane of the group gets the total
template <int GS> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// exclusive prefix sum over the lanes of a group (lane lg gets the sum of lanes < lg)
template <int GS> __device__ __forceinline__ float group_exscan(float v, int lg) {
  float incl = v;
#pragma unroll
  for (int o = 1; o < GS; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o, GS);
    if (lg >= o) incl += t;
  }
  return incl - v;
}

inline int snap_grid(long long FB, int C) {
  const int per_cta = SNAP_THREADS / snap_group_size(C);
  long long blocks = (FB + per_cta - 1) / per_cta;
  if (blocks > 148 * 16) blocks = 148 * 16;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace btk
