// pq_device.cuh -- bit-exact device restatement of ProductQuantizer::distance (StreamingSIMD).
#pragma once
#include "common.cuh"

// ProductQuantizer::distance(a, b, StreamingSIMD) (pq/mod.rs:231-266), bit-exact: lane accumulators shared across
// subspaces, one ordered reduction at the end, `sum_1 =` assignment for the scalar tail, D::outermost_op.
template <int METRIC, class CodeA, class CodeB>
__device__ __forceinline__ float pq_distance_streaming(const float *__restrict__ cb, uint32_t m, uint32_t K,
                                                       uint32_t dsub, CodeA ca, CodeB cbk) {
  float s16[16], s8[8], s4[4], s1 = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; i++) s16[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; i++) s8[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; i++) s4[i] = 0.0f;
  if (dsub == 8) {
    // common case (reference default pq_subvector_dimension = 8): one 8-lane chunk per subspace, two 16-byte loads per
    // centroid; identical arithmetic to the general cascade below
    for (uint32_t s = 0; s < m; s++) {
      const float4 *av = (const float4 *)(cb + ((size_t)s * K + ca(s)) * 8);
      const float4 *bv = (const float4 *)(cb + ((size_t)s * K + cbk(s)) * 8);
      const float4 a0 = __ldg(av), a1 = __ldg(av + 1), b0 = __ldg(bv), b1 = __ldg(bv + 1);
      const float x[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float y[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int l = 0; l < 8; l++) {
        if (METRIC == MGPU_L2) { float d = __fsub_rn(x[l], y[l]); s8[l] = __fadd_rn(s8[l], __fmul_rn(d, d)); }
        else s8[l] = __fadd_rn(s8[l], __fmul_rn(x[l], y[l]));
      }
    }
    float r8 = __fadd_rn(ordered_reduce(s16, 16), ordered_reduce(s8, 8));
    r8 = __fadd_rn(r8, ordered_reduce(s4, 4));
    r8 = __fadd_rn(r8, s1);
    return METRIC == MGPU_L2 ? r8 : -r8;
  }
  for (uint32_t s = 0; s < m; s++) {
    const float *av = cb + ((size_t)s * K + ca(s)) * dsub;
    const float *bv = cb + ((size_t)s * K + cbk(s)) * dsub;
    uint32_t p = 0, rem = dsub;
    if (rem / 16 > 0) {
      uint32_t chunks = rem / 16;
      for (uint32_t c = 0; c < chunks; c++) {
#pragma unroll
        for (int l = 0; l < 16; l++) {
          float x = av[p + c * 16 + l], y = bv[p + c * 16 + l];
          if (METRIC == MGPU_L2) { float d = __fsub_rn(x, y); s16[l] = __fadd_rn(s16[l], __fmul_rn(d, d)); }
          else s16[l] = __fadd_rn(s16[l], __fmul_rn(x, y));
        }
      }
      p += chunks * 16; rem -= chunks * 16;
    }
    if (rem / 8 > 0) {
#pragma unroll
      for (int l = 0; l < 8; l++) {
        float x = av[p + l], y = bv[p + l];
        if (METRIC == MGPU_L2) { float d = __fsub_rn(x, y); s8[l] = __fadd_rn(s8[l], __fmul_rn(d, d)); }
        else s8[l] = __fadd_rn(s8[l], __fmul_rn(x, y));
      }
      p += 8; rem -= 8;
    }
    if (rem / 4 > 0) {
#pragma unroll
      for (int l = 0; l < 4; l++) {
        float x = av[p + l], y = bv[p + l];
        if (METRIC == MGPU_L2) { float d = __fsub_rn(x, y); s4[l] = __fadd_rn(s4[l], __fmul_rn(d, d)); }
        else s4[l] = __fadd_rn(s4[l], __fmul_rn(x, y));
      }
      p += 4; rem -= 4;
    }
    if (rem > 0) {
      float t = 0.0f;
      for (uint32_t i = 0; i < rem; i++) {
        float x = av[p + i], y = bv[p + i];
        if (METRIC == MGPU_L2) { float d = __fsub_rn(x, y); t = __fadd_rn(t, __fmul_rn(d, d)); }
        else t = __fadd_rn(t, __fmul_rn(x, y));
      }
      s1 = t;  // assignment, as in the reference (pq/mod.rs:260)
    }
  }
  float r = __fadd_rn(ordered_reduce(s16, 16), ordered_reduce(s8, 8));
  r = __fadd_rn(r, ordered_reduce(s4, 4));
  r = __fadd_rn(r, s1);
  return METRIC == MGPU_L2 ? r : -r;
}

struct RowMajorCode {
  const uint8_t *p;
  __device__ __forceinline__ uint32_t operator()(uint32_t s) const { return p[s]; }
};

