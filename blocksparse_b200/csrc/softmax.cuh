// Block-sparse (masked) softmax, its gradient and the partial-autoregressive mask.
//
// Replaces bst_masked_softmax / bst_masked_softmax_grad / bst_partial_autoregressive_mask
// (src/bst_softmax_op_gpu.cu:12-198, 200-310, 461-520).
//
// HBM-bound: algorithmic traffic is (s_in + s_out) bytes per element of the
// (batch, heads, blocks, bs, bs) tensor.  Work decomposition (differs from the
// reference's one-CTA-per-query-row): one WARP owns 64/bs consecutive query rows of
// one query block, so that every global access of the warp is a full, contiguous
// 128-byte line of one bs x bs block (2 x 16-bit elements per lane); row statistics
// are reduced with xor-shuffles inside the bs/2-lane group that shares a row.  The
// row's values are held in registers between the statistics pass and the write pass
// when the row has <= KEEP key blocks (the common case); longer rows re-read (L2 hits).
#pragma once
#include <float.h>
#include "common.cuh"

namespace bsmm {

struct SoftmaxParams {
  const int32_t* nn_lut;      // [lut_heads][ctx_blks_q + blocks][2]
  const int32_t* nt_lut;      // [lut_heads][blocks][2], only for autoregress
  long long nn_head_stride, nt_head_stride;
  const void* mask;           // uintBS [mask_heads][blocks][BS] or null
  long long mask_head_stride; // words
  int autoregress_at_key;     // <0: off
  const void* x; void* y;     // x = input (or dy for grad); y = output (or dx)
  const void* y_in;           // grad only: softmax output
  float scale;
  int batch, heads, blocks, ctx_blks_q;
};

template <int BS> struct MaskWord;
template <> struct MaskWord<8>  { using type = uint8_t;  };
template <> struct MaskWord<16> { using type = uint16_t; };
template <> struct MaskWord<32> { using type = uint32_t; };
template <> struct MaskWord<64> { using type = uint64_t; };

template <typename T> struct Pair;
template <> struct Pair<float>         { using type = float2; };
template <> struct Pair<__half>        { using type = __half2; };
template <> struct Pair<__nv_bfloat16> { using type = __nv_bfloat162; };

template <typename T> __device__ __forceinline__ float2 load2(const T* p) {
  typename Pair<T>::type v = *reinterpret_cast<const typename Pair<T>::type*>(p);
  return make_float2(to_f32<T>(v.x), to_f32<T>(v.y));
}
template <typename T> __device__ __forceinline__ void store2(T* p, float a, float b) {
  typename Pair<T>::type v;
  v.x = from_f32<T>(a); v.y = from_f32<T>(b);
  *reinterpret_cast<typename Pair<T>::type*>(p) = v;
}

template <int BS>
__device__ __forceinline__ uint64_t autoregress_word(uint64_t word, int ak, int k_blk, int q_row) {
  // blocksparse/transformer.py:264-274
  const int k0 = k_blk * BS;
  const int sa = BS - min(max(ak - k0, 0), BS);
  const int sb = min(max(BS - 1 + k0 - q_row, 0), BS);
  const int sh = min(sa, sb);
  const uint64_t ones = (BS == 64) ? ~0ull : ((1ull << BS) - 1ull);
  return sh >= BS ? 0ull : (word & (ones >> sh));
}

constexpr int SOFTMAX_WARPS = 4;
constexpr int SOFTMAX_KEEP = 16;    // key blocks of a row kept in registers

template <typename TX, typename TY, int BS>
__global__ void __launch_bounds__(SOFTMAX_WARPS * 32)
bst_softmax_kernel(const SoftmaxParams p) {
  using MT = typename MaskWord<BS>::type;
  constexpr int R = 64 / BS;               // rows per warp
  constexpr int GROUPS = BS / R;           // row groups per query block
  constexpr int HALF = BS / 2;             // lanes per row
  const int lane = threadIdx.x % 32;
  const long long gid = (long long)blockIdx.x * SOFTMAX_WARPS + threadIdx.x / 32;
  if (gid >= (long long)p.ctx_blks_q * GROUPS) return;
  const int q = (int)(gid / GROUPS);
  const int row = (int)(gid % GROUPS) * R + lane / HALF;
  const int col = (lane % HALF) * 2;
  const int h = blockIdx.y, b = blockIdx.z;

  const int hl = p.nn_head_stride ? h : 0;
  const int32_t* lut = p.nn_lut + (long long)hl * p.nn_head_stride;
  const int first = lut[2 * q], count = lut[2 * q + 1];
  if (count == 0) return;
  const long long zoff = ((long long)b * p.heads + h) * p.blocks;
  const TX* x = reinterpret_cast<const TX*>(p.x);
  TY* y = reinterpret_cast<TY*>(p.y);
  const MT* mask = reinterpret_cast<const MT*>(p.mask);
  if (mask) mask += (p.mask_head_stride ? (long long)h * p.mask_head_stride : 0);
  const int32_t* nt = p.nt_lut ? p.nt_lut + (long long)hl * p.nt_head_stride : nullptr;
  (void)nt;

  auto load_entry = [&](int e, float& v0, float& v1) {
    const int blk = lut[2 * (first + e)];
    const int kb = lut[2 * (first + e) + 1];
    float2 v = load2<TX>(x + (zoff + blk) * (BS * BS) + row * BS + col);
    v0 = v.x * p.scale; v1 = v.y * p.scale;
    if (mask) {
      uint64_t w = (uint64_t)mask[(long long)blk * BS + row];
      if (p.autoregress_at_key >= 0) w = autoregress_word<BS>(w, p.autoregress_at_key, kb, q * BS + row);
      if (!((w >> col) & 1ull)) v0 = -FLT_MAX;
      if (!((w >> (col + 1)) & 1ull)) v1 = -FLT_MAX;
    }
  };

  float keep0[SOFTMAX_KEEP], keep1[SOFTMAX_KEEP];
  float m = -FLT_MAX;
#pragma unroll
  for (int e = 0; e < SOFTMAX_KEEP; ++e) {
    keep0[e] = keep1[e] = -FLT_MAX;
    if (e < count) {
      load_entry(e, keep0[e], keep1[e]);
      m = fmaxf(m, fmaxf(keep0[e], keep1[e]));
    }
  }
  for (int e = SOFTMAX_KEEP; e < count; ++e) {
    float v0, v1; load_entry(e, v0, v1);
    m = fmaxf(m, fmaxf(v0, v1));
  }
#pragma unroll
  for (int o = HALF / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));

  constexpr float LOG2E = 1.4426950408889634f;
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < SOFTMAX_KEEP; ++e) {
    if (e < count) {
      keep0[e] = exp2f((keep0[e] - m) * LOG2E);
      keep1[e] = exp2f((keep1[e] - m) * LOG2E);
      s += keep0[e] + keep1[e];
    }
  }
  for (int e = SOFTMAX_KEEP; e < count; ++e) {
    float v0, v1; load_entry(e, v0, v1);
    s += exp2f((v0 - m) * LOG2E) + exp2f((v1 - m) * LOG2E);
  }
#pragma unroll
  for (int o = HALF / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.f / s;

#pragma unroll
  for (int e = 0; e < SOFTMAX_KEEP; ++e) {
    if (e < count) {
      const int blk = lut[2 * (first + e)];
      store2<TY>(y + (zoff + blk) * (BS * BS) + row * BS + col, keep0[e] * inv, keep1[e] * inv);
    }
  }
  for (int e = SOFTMAX_KEEP; e < count; ++e) {
    float v0, v1; load_entry(e, v0, v1);
    const int blk = lut[2 * (first + e)];
    store2<TY>(y + (zoff + blk) * (BS * BS) + row * BS + col,
               exp2f((v0 - m) * LOG2E) * inv, exp2f((v1 - m) * LOG2E) * inv);
  }
}

template <typename T, typename TD, int BS>
__global__ void __launch_bounds__(SOFTMAX_WARPS * 32)
bst_softmax_grad_kernel(const SoftmaxParams p) {
  constexpr int R = 64 / BS;
  constexpr int GROUPS = BS / R;
  constexpr int HALF = BS / 2;
  const int lane = threadIdx.x % 32;
  const long long gid = (long long)blockIdx.x * SOFTMAX_WARPS + threadIdx.x / 32;
  if (gid >= (long long)p.ctx_blks_q * GROUPS) return;
  const int q = (int)(gid / GROUPS);
  const int row = (int)(gid % GROUPS) * R + lane / HALF;
  const int col = (lane % HALF) * 2;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hl = p.nn_head_stride ? h : 0;
  const int32_t* lut = p.nn_lut + (long long)hl * p.nn_head_stride;
  const int first = lut[2 * q], count = lut[2 * q + 1];
  if (count == 0) return;
  const long long zoff = ((long long)b * p.heads + h) * p.blocks;
  const T* dy = reinterpret_cast<const T*>(p.x);
  const T* yv = reinterpret_cast<const T*>(p.y_in);
  TD* dx = reinterpret_cast<TD*>(p.y);

  float kd0[SOFTMAX_KEEP], kd1[SOFTMAX_KEEP], ky0[SOFTMAX_KEEP], ky1[SOFTMAX_KEEP];
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < SOFTMAX_KEEP; ++e) {
    kd0[e] = kd1[e] = ky0[e] = ky1[e] = 0.f;
    if (e < count) {
      const long long off = (zoff + lut[2 * (first + e)]) * (BS * BS) + row * BS + col;
      float2 d = load2<T>(dy + off), v = load2<T>(yv + off);
      kd0[e] = d.x; kd1[e] = d.y; ky0[e] = v.x; ky1[e] = v.y;
      s += d.x * v.x + d.y * v.y;
    }
  }
  for (int e = SOFTMAX_KEEP; e < count; ++e) {
    const long long off = (zoff + lut[2 * (first + e)]) * (BS * BS) + row * BS + col;
    float2 d = load2<T>(dy + off), v = load2<T>(yv + off);
    s += d.x * v.x + d.y * v.y;
  }
#pragma unroll
  for (int o = HALF / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
#pragma unroll
  for (int e = 0; e < SOFTMAX_KEEP; ++e) {
    if (e < count) {
      const long long off = (zoff + lut[2 * (first + e)]) * (BS * BS) + row * BS + col;
      store2<TD>(dx + off, (kd0[e] - s) * ky0[e] * p.scale, (kd1[e] - s) * ky1[e] * p.scale);
    }
  }
  for (int e = SOFTMAX_KEEP; e < count; ++e) {
    const long long off = (zoff + lut[2 * (first + e)]) * (BS * BS) + row * BS + col;
    float2 d = load2<T>(dy + off), v = load2<T>(yv + off);
    store2<TD>(dx + off, (d.x - s) * v.x * p.scale, (d.y - s) * v.y * p.scale);
  }
}

template <int BS>
__global__ void bst_autoregressive_mask_kernel(const int32_t* __restrict__ nt_lut, long long nt_head_stride,
                                               const void* __restrict__ mask_in, void* __restrict__ mask_out,
                                               int blocks, int ak) {
  using MT = typename MaskWord<BS>::type;
  const int hl = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;     // (blk, row)
  if (idx >= blocks * BS) return;
  const int blk = idx / BS, row = idx % BS;
  const int32_t* nt = nt_lut + (long long)hl * nt_head_stride;
  const int qb = nt[2 * blk], kb = nt[2 * blk + 1];
  const MT* in = reinterpret_cast<const MT*>(mask_in) + (long long)hl * blocks * BS;
  MT* out = reinterpret_cast<MT*>(mask_out) + (long long)hl * blocks * BS;
  out[idx] = (MT)autoregress_word<BS>((uint64_t)in[idx], ak, kb, qb * BS + row);
}

}  // namespace bsmm
