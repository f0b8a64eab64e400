// Block-sparse (masked) softmax, its gradient and the partial-autoregressive mask.
//
// Replaces bst_masked_softmax / bst_masked_softmax_grad / bst_partial_autoregressive_mask
// (src/bst_softmax_op_gpu.cu:12-198, 200-310, 461-520).
//
// HBM-bound: algorithmic traffic is (s_in + s_out) bytes per element of the
// (batch, heads, blocks, bs, bs) tensor.  Work decomposition (differs from the
// reference's one-CTA-per-query-row): one WARP owns 4 (bs 64) or 8 consecutive query
// rows of one query block and each lane 16 bytes of a row, so every warp-wide access is
// one contiguous 256..512-byte chunk of a block; the row's LUT entries are loaded once
// (one per lane) and broadcast by shuffle; row statistics are reduced with xor-shuffles
// inside the 4/8-lane group that shares a row.  Values stay in registers between the
// statistics pass and the write pass for rows of <= KEEP key blocks; longer rows re-read.
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

// Lane mapping: a warp covers RP consecutive query rows of one block per pass; LPR lanes share a row and each
// lane owns EPL consecutive keys (16 bytes of 16-bit data when bs >= 32), so one warp-wide load is a single
// contiguous RP*bs*sizeof(T) chunk of the block (512 B for bs 64).
template <int BS> struct SoftmaxMap {
  static constexpr int LPR = (BS == 64) ? 8 : 4;        // lanes per row
  static constexpr int EPL = BS / LPR;                  // elements per lane: 8, 8, 4, 2
  static constexpr int RP = 32 / LPR;                   // rows per pass: 4, 8, 8, 8
  static constexpr int GROUPS = BS / RP;                // passes per query block
  static constexpr int KEEP = (EPL >= 8) ? 4 : 8;       // key blocks of a row kept in registers (more costs occupancy: measured)
};

template <typename T, int EPL> __device__ __forceinline__ void load_vec(const T* p, float (&f)[EPL]) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int i = 0; i < EPL; i += 2) { const float2 v = *reinterpret_cast<const float2*>(p + i); f[i] = v.x; f[i + 1] = v.y; }
  } else if constexpr (EPL == 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { const typename Pair<T>::type h = *reinterpret_cast<const typename Pair<T>::type*>(&w[i]); f[2 * i] = to_f32<T>(h.x); f[2 * i + 1] = to_f32<T>(h.y); }
  } else if constexpr (EPL == 4) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const uint32_t w[2] = {v.x, v.y};
#pragma unroll
    for (int i = 0; i < 2; ++i) { const typename Pair<T>::type h = *reinterpret_cast<const typename Pair<T>::type*>(&w[i]); f[2 * i] = to_f32<T>(h.x); f[2 * i + 1] = to_f32<T>(h.y); }
  } else {
    const float2 v = load2<T>(p); f[0] = v.x; f[1] = v.y;
  }
}
template <typename T, int EPL> __device__ __forceinline__ void store_vec(T* p, const float (&f)[EPL]) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int i = 0; i < EPL; i += 2) *reinterpret_cast<float2*>(p + i) = make_float2(f[i], f[i + 1]);
  } else if constexpr (EPL == 8) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { typename Pair<T>::type h; h.x = from_f32<T>(f[2 * i]); h.y = from_f32<T>(f[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&h); }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  } else if constexpr (EPL == 4) {
    uint32_t w[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) { typename Pair<T>::type h; h.x = from_f32<T>(f[2 * i]); h.y = from_f32<T>(f[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&h); }
    *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
  } else {
    store2<T>(p, f[0], f[1]);
  }
}

template <typename TX, typename TY, int BS>
__global__ void __launch_bounds__(SOFTMAX_WARPS * 32)
bst_softmax_kernel(const SoftmaxParams p) {
  using MT = typename MaskWord<BS>::type;
  using M = SoftmaxMap<BS>;
  constexpr int LPR = M::LPR, EPL = M::EPL, RP = M::RP, GROUPS = M::GROUPS, KEEP = M::KEEP;
  const int lane = threadIdx.x % 32;
  const long long gid = (long long)blockIdx.x * SOFTMAX_WARPS + threadIdx.x / 32;
  if (gid >= (long long)p.ctx_blks_q * GROUPS) return;
  const int q = (int)(gid / GROUPS);
  const int row = (int)(gid % GROUPS) * RP + lane / LPR;
  const int col = (lane % LPR) * EPL;
  const int h = blockIdx.y, b = blockIdx.z;

  const int hl = p.nn_head_stride ? h : 0;
  const int32_t* lut = p.nn_lut + (long long)hl * p.nn_head_stride;
  const int first = lut[2 * q], count = lut[2 * q + 1];
  if (count == 0) return;
  const long long zoff = ((long long)b * p.heads + h) * p.blocks;
  const TX* x = reinterpret_cast<const TX*>(p.x) + (long long)row * BS + col;
  TY* y = reinterpret_cast<TY*>(p.y) + (long long)row * BS + col;
  const MT* mask = reinterpret_cast<const MT*>(p.mask);
  if (mask) mask += (p.mask_head_stride ? (long long)h * p.mask_head_stride : 0) + row;
  // the row's LUT entries, one per lane (rows longer than 32 key blocks reload per chunk of 32)
  const int2* ent = reinterpret_cast<const int2*>(lut) + first;
  int2 my = (lane < count) ? ent[lane] : make_int2(0, 0);

  auto entry = [&](int e, int& blk, int& kb) {
    if (e < 32) { blk = __shfl_sync(0xffffffffu, my.x, e); kb = __shfl_sync(0xffffffffu, my.y, e); }
    else { const int2 v = ent[e]; blk = v.x; kb = v.y; }
  };
  auto load_entry = [&](int e, float (&v)[EPL]) {
    int blk, kb;
    entry(e, blk, kb);
    load_vec<TX, EPL>(x + (zoff + blk) * (BS * BS), v);
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] *= p.scale;
    if (mask) {
      uint64_t w = (uint64_t)mask[(long long)blk * BS];
      if (p.autoregress_at_key >= 0) w = autoregress_word<BS>(w, p.autoregress_at_key, kb, q * BS + row);
      // most blocks are fully visible (only e.g. the diagonal ones carry a causal pattern): skip the bit tests there
      const uint64_t mine = (w >> col) & ((1ull << EPL) - 1ull);
      if (mine != ((1ull << EPL) - 1ull)) {
#pragma unroll
        for (int i = 0; i < EPL; ++i) if (!((mine >> i) & 1ull)) v[i] = -FLT_MAX;
      }
    }
  };

  // Pass A (one read of the row): online max / sum -- (m, s) with s = sum exp(v - m), rescaled whenever m grows.
  constexpr float LOG2E = 1.4426950408889634f;
  // exp2 is the other bound of this kernel (MUFU: 16/clk/SM, i.e. ~0.1 ms for cfg 3 if every element needed two):
  // kept entries are exponentiated ONCE, against the running max at that time (mref[e]), and rescaled by one
  // scalar exp2 per entry at the end.
  float keep[KEEP][EPL];
  float mref[KEEP];
  float m = -FLT_MAX, s = 0.f;
  auto absorb = [&](float (&v)[EPL]) -> float {
    float mv = v[0];
#pragma unroll
    for (int i = 1; i < EPL; ++i) mv = fmaxf(mv, v[i]);
    const float mn = fmaxf(m, mv);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) { v[i] = exp2f((v[i] - mn) * LOG2E); acc += v[i]; }
    s = s * exp2f((m - mn) * LOG2E) + acc;
    m = mn;
    return mn;
  };
#pragma unroll
  for (int e = 0; e < KEEP; ++e) {
    mref[e] = 0.f;
    if (e < count) { load_entry(e, keep[e]); mref[e] = absorb(keep[e]); }
  }
#pragma unroll 4
  for (int e = KEEP; e < count; ++e) {
    float v[EPL]; load_entry(e, v);
    absorb(v);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    const float mo = __shfl_xor_sync(0xffffffffu, m, o);
    const float so = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, mo);
    s = s * exp2f((m - mn) * LOG2E) + so * exp2f((mo - mn) * LOG2E);
    m = mn;
  }
  const float inv = 1.f / s;

  // Pass B: normalise and write (kept values from registers, the rest re-read -- L2 hits)
#pragma unroll
  for (int e = 0; e < KEEP; ++e) {
    if (e < count) {
      int blk, kb; entry(e, blk, kb);
      const float sc = exp2f((mref[e] - m) * LOG2E) * inv;
#pragma unroll
      for (int i = 0; i < EPL; ++i) keep[e][i] *= sc;
      store_vec<TY, EPL>(y + (zoff + blk) * (BS * BS), keep[e]);
    }
  }
#pragma unroll 4
  for (int e = KEEP; e < count; ++e) {
    float v[EPL]; load_entry(e, v);
    int blk, kb; entry(e, blk, kb);
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] = exp2f((v[i] - m) * LOG2E) * inv;
    store_vec<TY, EPL>(y + (zoff + blk) * (BS * BS), v);
  }
}

template <typename T, typename TD, int BS>
__global__ void __launch_bounds__(SOFTMAX_WARPS * 32)
bst_softmax_grad_kernel(const SoftmaxParams p) {
  using M = SoftmaxMap<BS>;
  constexpr int LPR = M::LPR, EPL = M::EPL, RP = M::RP, GROUPS = M::GROUPS, KEEP = M::KEEP / 2 < 4 ? 4 : M::KEEP / 2;
  const int lane = threadIdx.x % 32;
  const long long gid = (long long)blockIdx.x * SOFTMAX_WARPS + threadIdx.x / 32;
  if (gid >= (long long)p.ctx_blks_q * GROUPS) return;
  const int q = (int)(gid / GROUPS);
  const int row = (int)(gid % GROUPS) * RP + lane / LPR;
  const int col = (lane % LPR) * EPL;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hl = p.nn_head_stride ? h : 0;
  const int32_t* lut = p.nn_lut + (long long)hl * p.nn_head_stride;
  const int first = lut[2 * q], count = lut[2 * q + 1];
  if (count == 0) return;
  const long long zoff = ((long long)b * p.heads + h) * p.blocks;
  const long long roff = (long long)row * BS + col;
  const T* dy = reinterpret_cast<const T*>(p.x) + roff;
  const T* yv = reinterpret_cast<const T*>(p.y_in) + roff;
  TD* dx = reinterpret_cast<TD*>(p.y) + roff;
  const int2* ent = reinterpret_cast<const int2*>(lut) + first;
  const int my = (lane < count) ? ent[lane].x : 0;
  auto block_of = [&](int e) { return e < 32 ? __shfl_sync(0xffffffffu, my, e) : ent[e].x; };

  float kd[KEEP][EPL], ky[KEEP][EPL];
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < KEEP; ++e) {
    if (e < count) {
      const long long off = (zoff + block_of(e)) * (BS * BS);
      load_vec<T, EPL>(dy + off, kd[e]);
      load_vec<T, EPL>(yv + off, ky[e]);
#pragma unroll
      for (int i = 0; i < EPL; ++i) s += kd[e][i] * ky[e][i];
    }
  }
  for (int e = KEEP; e < count; ++e) {
    const long long off = (zoff + block_of(e)) * (BS * BS);
    float d[EPL], v[EPL];
    load_vec<T, EPL>(dy + off, d); load_vec<T, EPL>(yv + off, v);
#pragma unroll
    for (int i = 0; i < EPL; ++i) s += d[i] * v[i];
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
#pragma unroll
  for (int e = 0; e < KEEP; ++e) {
    if (e < count) {
      const long long off = (zoff + block_of(e)) * (BS * BS);
      float o[EPL];
#pragma unroll
      for (int i = 0; i < EPL; ++i) o[i] = (kd[e][i] - s) * ky[e][i] * p.scale;
      store_vec<TD, EPL>(dx + off, o);
    }
  }
  for (int e = KEEP; e < count; ++e) {
    const long long off = (zoff + block_of(e)) * (BS * BS);
    float d[EPL], v[EPL];
    load_vec<T, EPL>(dy + off, d); load_vec<T, EPL>(yv + off, v);
#pragma unroll
    for (int i = 0; i < EPL; ++i) d[i] = (d[i] - s) * v[i] * p.scale;
    store_vec<TD, EPL>(dx + off, d);
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
