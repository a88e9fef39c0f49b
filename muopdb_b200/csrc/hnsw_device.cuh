// hnsw_device.cuh -- graph view and register-resident sorted lists shared by the HNSW kernels (hnsw.cu, hnsw_spec.cu).
#pragma once
#include "internal.cuh"

struct HnswDev {
  const uint32_t *edges, *points, *upper_pid, *upper_pos;
  const int32_t *upper_dense;  // [(layer-1) * n + point] -> position in `points`, -1 if absent; may be null
  const uint32_t *edges0;      // fixed-stride layer-0 adjacency (deg0 entries per point, 0xFFFFFFFF padded); may be null
  uint32_t deg0, prefetch_rows;
  const uint64_t *edge_offsets, *level_offsets;
  const void *rows;
  const mgpu_u128 *doc_ids;
  const float *cb;
  uint32_t dim, qdim, num_layers, entry_point, m, K, dsub;
  uint64_t n, n_edge_offsets;
};


template <int EPL>
struct RegList {
  uint64_t v[EPL];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int e = 0; e < EPL; e++) v[e] = MGPU_EMPTY_KEY;
  }
  // element at global index g (warp-uniform g)
  __device__ __forceinline__ uint64_t get(int g) const {
    const int e = g % EPL;
    uint64_t x = v[0];
#pragma unroll
    for (int i = 1; i < EPL; i++) if (e == i) x = v[i];
    return shfl64(x, g / EPL);
  }
  // sorted insert of a warp-uniform key; the last entry falls off the end
  __device__ __forceinline__ void insert(uint64_t key) {
    const int lane = lane_id();
    int c = 0;
#pragma unroll
    for (int e = 0; e < EPL; e++) c += v[e] <= key ? 1 : 0;
    const int c_prev = __shfl_up_sync(0xffffffffu, c, 1);
    const uint64_t prev_last = shfl_up64(v[EPL - 1], 1);
    if (c == EPL) return;                                   // every entry of this lane stays in front of the key
    const bool ins_here = lane == 0 || c_prev == EPL;        // first lane with c < EPL
    uint64_t nv[EPL];
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      uint64_t x;
      if (e < c) x = v[e];
      else if (e == c) x = ins_here ? key : prev_last;       // c == 0 for every lane behind the insertion lane
      else x = v[e - 1];
      nv[e] = x;
    }
#pragma unroll
    for (int e = 0; e < EPL; e++) v[e] = nv[e];
  }
  // remove the first entry (shift left by one)
  __device__ __forceinline__ void pop_front() {
    const uint64_t next_first = shfl64(v[0], (lane_id() + 1) & 31);
#pragma unroll
    for (int e = 0; e + 1 < EPL; e++) v[e] = v[e + 1];
    v[EPL - 1] = lane_id() == 31 ? MGPU_EMPTY_KEY : next_first;
  }
  // number of entries among the first n whose distance key is <= fk (the array is sorted, so this is a prefix length)
  __device__ __forceinline__ int count_le(int n, uint32_t fk) const {
    const int lane = lane_id();
    int c = 0;
#pragma unroll
    for (int e = 0; e < EPL; e++) c += (lane * EPL + e < n && (uint32_t)(v[e] >> 32) <= fk) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    return c;
  }
};


int launch_hnsw_spec(mgpu_hnsw *h, const HnswDev &g, const HnswSearchArgs &a, uint32_t *err_flags, bool *launched);
