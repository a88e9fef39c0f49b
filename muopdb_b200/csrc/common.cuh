// common.cuh -- shared host/device helpers for the sm_100a search kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/muopdb_gpu.h"

#define MGPU_WARP 32
#define MGPU_CHUNK 32 /* rows per interleaved chunk: one row per lane */
#define MGPU_EMPTY_SLOT 0xFFFFFFFFu
#define MGPU_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

// ---- context -------------------------------------------------------------------------------
struct mgpu_ctx {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  size_t l2_bytes = 0;
  const char *last_kernel[16] = {};            // per kernel class: name of the kernel launched last (string literals)
  cudaStream_t stream = nullptr;
  cudaStream_t aux_stream = nullptr;           // side stream for work that is independent of the main chain (query encode)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool ext_codes_on_aux = false;               // sharded search: the gathered query codes are ready at ev_join (side stream)
  std::mutex mu;
  std::string err;
  // workspace (grown on demand)
  void *ws = nullptr; size_t ws_bytes = 0;
  void *pinned = nullptr; size_t pinned_bytes = 0;
  // timing
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  uint32_t profiling = 0;   // bit c set: launches of kernel class c are bracketed by events
  struct Prof { std::vector<cudaEvent_t> ev; uint64_t launches = 0; float done_ms = 0.f; };
  Prof prof[MGPU_K_COUNT];
  uint64_t launches = 0;
  // pipelined host-buffer searches (mgpu_ivf_search_submit / mgpu_search_wait): two slots alternate, each with its own
  // device staging for the queries and the results, so batch i+1's H2D copy and batch i-1's D2H copy overlap batch i's kernels
  struct Pipe {
    void *q = nullptr; size_t q_bytes = 0;       // staged queries
    void *out = nullptr; size_t out_bytes = 0;   // doc ids | scores | counts
    cudaEvent_t ev_h2d = nullptr, ev_done = nullptr, ev_out = nullptr;
    uint64_t seq = 0; bool pending = false, used = false;
  } pipe[2];
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  uint64_t pipe_seq = 0;
  void *shard_ws = nullptr; size_t shard_ws_bytes = 0;  // buffers of mgpu_shard_ivf_search (outlive the nested calls' workspace)
  // nccl
  void *nccl_lib = nullptr; void *nccl_comm = nullptr; int nranks = 1, rank = 0;
  // Result exchange of the sharded searches (all-gather of the per-shard top-k + merge kernel): runs on its own stream and,
  // when ncclCommSplit is available, on its own communicator, so that batch i's exchange overlaps batch i+1's coarse
  // scoring and scan (which stay on `stream`; the query-code all-gather of the split encode keeps `nccl_comm`).  Two slots of
  // per-batch buffers alternate.  ev_local: the shard's local top-k is ready; ev_xdone: the merged result is ready.
  void *nccl_comm_x = nullptr;
  cudaStream_t comm_stream = nullptr;
  struct ShardSlot { void *buf = nullptr; size_t bytes = 0; cudaEvent_t ev_local = nullptr, ev_xdone = nullptr; bool used = false; } shard_slot[2];
  uint64_t shard_seq = 0;
  bool shard_overlap = false;   // mgpu_shard_overlap: device-buffer sharded calls return with the exchange still in flight
};

int mgpu_fail(mgpu_ctx *ctx, int code, const char *fmt, ...);

#define CUDA_TRY(ctx, expr)                                                                          \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return mgpu_fail(ctx, _e == cudaErrorMemoryAllocation ? MGPU_ERR_OOM : MGPU_ERR_CUDA,           \
                       "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);   \
  } while (0)

#define MGPU_TRY(expr)            \
  do {                            \
    int _s = (expr);              \
    if (_s != MGPU_OK) return _s; \
  } while (0)

// Brackets one kernel launch for the per-class profiler and counts launches.
struct LaunchScope {
  mgpu_ctx *ctx; int cls; cudaStream_t st;
  // `name` (a string literal) is remembered as the class's most recent kernel: mgpu_last_kernel reports what actually ran
  LaunchScope(mgpu_ctx *c, int k, cudaStream_t s = nullptr, const char *name = nullptr) : ctx(c), cls(k), st(s ? s : c->stream) {
    ctx->launches++; ctx->prof[cls].launches++;
    if (name) ctx->last_kernel[cls] = name;
    if (ctx->profiling >> cls & 1u) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ctx->prof[cls].ev.push_back(e); }
  }
  ~LaunchScope() {
    if (ctx->profiling >> cls & 1u) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ctx->prof[cls].ev.push_back(e); }
  }
};

int mgpu_ws_reserve(mgpu_ctx *ctx, size_t bytes);     // device workspace
int mgpu_pinned_reserve(mgpu_ctx *ctx, size_t bytes); // pinned host staging

// bump allocator over the ctx workspace
struct WsAlloc {
  char *base; size_t off = 0, cap;
  WsAlloc(void *b, size_t c) : base((char *)b), cap(c) {}
  template <class T> T *get(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T *p = (T *)(base + off);
    off += n * sizeof(T);
    return p;
  }
};
static inline size_t ws_need(size_t acc, size_t bytes) { return ((acc + 255) & ~(size_t)255) + bytes; }

// ---- device helpers --------------------------------------------------------------------------
#ifdef __CUDACC__

// Order-preserving f32 -> u32 (ascending).  -0 is canonicalised to +0 (NotNan compares them equal,
// rs/index/src/utils.rs:71-76); NaNs sort last like IdWithScore::cmp (utils.rs:95-114).
__device__ __forceinline__ uint32_t f2key(float f) {
  f = f + 0.0f;
  uint32_t u = __float_as_uint(f);
  if (f != f) return 0xFFFFFFFEu;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float ordered_reduce(const float *acc, int lanes) {
  float s = -0.0f;  // std::simd reduce_sum == simd_reduce_add_ordered(self, -0.0)
#pragma unroll
  for (int i = 0; i < 16; i++)
    if (i < lanes) s = __fadd_rn(s, acc[i]);
  return s;
}

// Bit-faithful restatement of L2DistanceCalculator::calculate_squared (distance/l2.rs:30-68) and
// DotProductDistanceCalculator::calculate (distance/dot_product.rs:38-71) for arbitrary accessors:
// 16/8/4-lane accumulators, non-fused multiply/add, ordered lane reduction, scalar tail.
// ref_phase runs one lane-width phase starting at dimension p (returns the new p); ref_tail runs the 8-lane, 4-lane
// and scalar phases; ref_distance the whole cascade.  L2 results are SQUARED (caller applies sqrt); dot is NEGATED.
template <int METRIC, int L, class FA, class FB>
__device__ __forceinline__ int ref_phase(FA a, FB b, int p, int n, float &ret) {
  int rem = n - p;
  bool go = METRIC == MGPU_L2 ? (rem / L > 0) : (rem > L);
  if (!go) return p;
  float acc[L];
#pragma unroll
  for (int l = 0; l < L; l++) acc[l] = 0.0f;
  int chunks = rem / L;
  for (int c = 0; c < chunks; c++) {
#pragma unroll
    for (int l = 0; l < L; l++) {
      float x = a(p + c * L + l), y = b(p + c * L + l);
      if (METRIC == MGPU_L2) { float d = __fsub_rn(x, y); acc[l] = __fadd_rn(acc[l], __fmul_rn(d, d)); }
      else acc[l] = __fadd_rn(acc[l], __fmul_rn(x, y));
    }
  }
  float s = -0.0f;
#pragma unroll
  for (int l = 0; l < L; l++) s = __fadd_rn(s, acc[l]);
  ret = __fadd_rn(ret, s);
  return p + chunks * L;
}

template <int METRIC, class FA, class FB>
__device__ __forceinline__ float ref_tail(FA a, FB b, int p, int n, float ret) {
  p = ref_phase<METRIC, 8>(a, b, p, n, ret);
  p = ref_phase<METRIC, 4>(a, b, p, n, ret);
  for (; p < n; p++) {
    float x = a(p), y = b(p);
    if (METRIC == MGPU_L2) { float d = __fsub_rn(x, y); ret = __fadd_rn(ret, __fmul_rn(d, d)); }
    else ret = __fadd_rn(ret, __fmul_rn(x, y));
  }
  return ret;
}

template <int METRIC, class FA, class FB>
__device__ __forceinline__ float ref_distance(FA a, FB b, int n) {
  float ret = 0.0f;
  int p = ref_phase<METRIC, 16>(a, b, 0, n, ret);
  ret = ref_tail<METRIC>(a, b, p, n, ret);
  return METRIC == MGPU_L2 ? ret : -ret;
}

struct PtrAcc {
  const float *p;
  __device__ __forceinline__ float operator()(int i) const { return p[i]; }
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up64(uint64_t v, int d) {
  uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, d);
  uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), d);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor64(uint64_t v, int m) {
  uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m);
  uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
  return ((uint64_t)hi << 32) | lo;
}

// ---- warp-shuffle top-32 (ascending across lanes) -------------------------------------------
// Each warp keeps its 32 best (key, payload) pairs sorted across lanes: lane i holds the i-th smallest
// composite key.  Composite = (ordered score key << 32) | point_id, i.e. exactly the reference's
// (distance, point_id) ordering (rs/index/src/utils.rs:71-76).
struct WarpTop32 {
  uint64_t key; uint32_t pay;
  __device__ __forceinline__ void init() { key = MGPU_EMPTY_KEY; pay = MGPU_EMPTY_SLOT; }
  // insert one warp-uniform candidate
  __device__ __forceinline__ void insert(uint64_t ck, uint32_t cp) {
    bool gt = key > ck;
    uint64_t upk = shfl_up64(key, 1);
    uint32_t upp = __shfl_up_sync(0xffffffffu, pay, 1);
    bool upgt = (lane_id() > 0) && (upk > ck);
    if (gt) { if (upgt) { key = upk; pay = upp; } else { key = ck; pay = cp; } }
  }
  // offer per-lane candidates (pass = this lane has one); returns the warp's current worst score key
  __device__ __forceinline__ uint32_t offer(bool pass, uint64_t ck, uint32_t cp) {
    unsigned m = __ballot_sync(0xffffffffu, pass);
    while (m) {
      int src = __ffs(m) - 1;
      m &= m - 1;
      uint64_t k2 = shfl64(ck, src);
      uint32_t p2 = __shfl_sync(0xffffffffu, cp, src);
      insert(k2, p2);
    }
    return (uint32_t)(shfl64(key, 31) >> 32);
  }
  // bitonic sort of an arbitrary per-lane (key,pay) into ascending lane order
  __device__ __forceinline__ void sort() {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        uint64_t ok = shfl_xor64(key, j);
        uint32_t op = __shfl_xor_sync(0xffffffffu, pay, j);
        int l = lane_id();
        bool up = ((l & k) == 0);
        bool lower = ((l & j) == 0);
        bool take = (lower == up) ? (ok < key) : (ok > key);
        if (take) { key = ok; pay = op; }
      }
    }
  }
  // merge with another sorted list (given as this lane's element of it): keeps the 32 smallest, sorted
  __device__ __forceinline__ void merge(uint64_t okey, uint32_t opay) {
    // reverse the other list, elementwise min -> bitonic sequence holding the 32 smallest
    uint64_t rk = shfl64(okey, 31 - lane_id());
    uint32_t rp = __shfl_sync(0xffffffffu, opay, 31 - lane_id());
    if (rk < key) { key = rk; pay = rp; }
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
      uint64_t ok = shfl_xor64(key, j);
      uint32_t op = __shfl_xor_sync(0xffffffffu, pay, j);
      bool lower = ((lane_id() & j) == 0);
      bool take = lower ? (ok < key) : (ok > key);
      if (take) { key = ok; pay = op; }
    }
  }
};

#endif  // __CUDACC__
