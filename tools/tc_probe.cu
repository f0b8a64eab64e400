// tcgen05 probe: validates our shared-memory / instruction descriptor encodings against a CPU
// GEMM and measures issue throughput of small-N MMAs (the block-sparse case) on a B200.
// Every wait is bounded, so a protocol error prints TIMEOUT instead of hanging the GPU.
//
//   tools/tc_probe <test>     test in {ss_kk, ss_kmn, tma_kk, tma_mnk, ts_st, ts_cp, ss16_kk, ss16_kmn, bench}
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../blocksparse_b200/csrc/ptx.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

typedef __nv_bfloat16 bf16;

__host__ __device__ inline uint32_t swz64(uint32_t off) { return off ^ (((off >> 7) & 3u) << 4); }
__host__ __device__ inline uint32_t swz128(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

enum { T_SS_KK = 0, T_SS_KMN, T_TMA_KK, T_TMA_MNK, T_TS_ST, T_TS_CP };

struct Maps { CUtensorMap a, b; };

// One CTA, 128 threads.  D[128 x 32] = A[128 x 32] * B[32 x 32] with two K=16 MMAs.
//   Ag: A row-major [128][32] (K contiguous)     -- or for T_TMA_MNK the global is [32 k][128 m]
//   Bg: for *_kk   B^T row-major [n=32][k=32] (K contiguous);  for *_kmn  B row-major [k=32][n=32]
__global__ void __launch_bounds__(128) probe_gemm(int test, const bf16* Ag, const bf16* Bg, float* Dg, int* status,
                                                  const __grid_constant__ Maps maps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;              // 8 KB
  uint8_t* sB = smem + 8192;       // 2 KB
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;

  if (tid == 0) { ptx::mbar_init(&bar_tma, 1); ptx::mbar_init(&bar_mma, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base_s, 64); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  const bool use_tma = (test == T_TMA_KK || test == T_TMA_MNK);
  const bool a_mn = (test == T_TMA_MNK);
  const bool b_mn = (test == T_SS_KMN || test == T_TMA_MNK);
  if (!use_tma) {
    // hand-filled operands in the layout we believe the descriptors describe
    for (int i = tid; i < 128 * 32; i += 128) {
      int m = i / 32, k = i % 32;
      *reinterpret_cast<bf16*>(sA + swz64(m * 64 + k * 2)) = Ag[m * 32 + k];
    }
    for (int i = tid; i < 32 * 32; i += 128) {
      int r = i / 32, c = i % 32;      // row-major 32 x 32, 64-byte rows, both majors
      *reinterpret_cast<bf16*>(sB + swz64(r * 64 + c * 2)) = Bg[r * 32 + c];
    }
    ptx::fence_proxy_async();
    __syncthreads();
  } else {
    if (tid == 0) {
      ptx::mbar_expect_tx(&bar_tma, 8192 + 2048);
      if (!a_mn) {
        ptx::tma_load_2d(sA, &maps.a, &bar_tma, 0, 0);          // box {32 k, 128 m}, swizzle 64B
      } else {
        ptx::tma_load_2d(sA, &maps.a, &bar_tma, 0, 0);          // box {64 m, 32 k}, swizzle 128B
        ptx::tma_load_2d(sA + 4096, &maps.a, &bar_tma, 64, 0);
      }
      ptx::tma_load_2d(sB, &maps.b, &bar_tma, 0, 0);            // box {32, 32}, swizzle 64B
    }
    if (!ptx::mbar_wait(&bar_tma, 0)) { if (tid == 0) status[0] = 1; }
    __syncthreads();
  }

  const uint32_t idesc = ptx::make_idesc_f16(true, a_mn, b_mn, 128, 32);
  if (test == T_TS_ST) {
    // A straight from registers into TMEM columns [32, 48): lane = row, column j = (k = 2j, 2j+1)
    uint32_t r[8];
    for (int ks = 0; ks < 2; ++ks) {
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(Ag + tid * 32 + ks * 16 + 2 * j);
        r[j] = *reinterpret_cast<uint32_t*>(&v);
      }
      ptx::tmem_st_x8(tmem + ((uint32_t)(warp * 32) << 16) + 32 + ks * 8, r);
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t a0 = ptx::smem_u32(sA), b0 = ptx::smem_u32(sB);
    for (int ks = 0; ks < 2; ++ks) {
      uint64_t adesc, bdesc;
      if (!a_mn) adesc = ptx::make_smem_desc(a0 + ks * 32, 16, 512, ptx::SWZ_64B);
      else       adesc = ptx::make_smem_desc(a0 + ks * 2048, 4096, 1024, ptx::SWZ_128B);
      if (!b_mn) bdesc = ptx::make_smem_desc(b0 + ks * 32, 16, 512, ptx::SWZ_64B);
      else       bdesc = ptx::make_smem_desc(b0 + ks * 1024, 2048, 512, ptx::SWZ_64B);
      if (test == T_TS_CP) {
        ptx::tc_cp_128x256b(tmem + 32 + ks * 8, adesc);
        ptx::mma_ts(tmem, tmem + 32 + ks * 8, bdesc, idesc, ks);
      } else if (test == T_TS_ST) {
        ptx::mma_ts(tmem, tmem + 32 + ks * 8, bdesc, idesc, ks);
      } else {
        ptx::mma_ss(tmem, adesc, bdesc, idesc, ks);
      }
    }
    ptx::tc_commit(&bar_mma);
  }
  if (!ptx::mbar_wait(&bar_mma, 0)) { if (tid == 0) status[0] = 2; }
  ptx::tc_fence_after();
  uint32_t r[32];
  ptx::tmem_ld_x32(tmem + ((uint32_t)(warp * 32) << 16), r);
  ptx::tmem_ld_wait();
  for (int j = 0; j < 32; ++j) Dg[tid * 32 + j] = __uint_as_float(r[j]);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

// ----------------------------------------------------------------------------------------------
// 16 x 16 blocks (SWIZZLE_32B operands, 32-byte rows): D[128 x 32] = A[128 x 16] * [B0 | B1] where B0, B1 are two
// 16 x 16 blocks staged 512 bytes apart and issued as ONE N = 32 MMA (the merged-run case of the xprop kernel).
//   b_mn = 0: blocks stored [n][k] (K-major, bprop);  b_mn = 1: blocks stored [k][n] (MN-major, fprop).
__host__ __device__ inline uint32_t swz32(uint32_t off) { return off ^ (((off >> 7) & 1u) << 4); }

__global__ void __launch_bounds__(128) probe_gemm16(int b_mn, const bf16* Ag, const bf16* Bg, float* Dg, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;              // 128 rows x 32 B = 4 KB
  uint8_t* sB = smem + 4096;       // 2 blocks x 512 B
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  if (tid == 0) { ptx::mbar_init(&bar_mma, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base_s, 32); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  for (int i = tid; i < 128 * 16; i += 128) {
    const int m = i / 16, k = i % 16;
    *reinterpret_cast<bf16*>(sA + swz32(m * 32 + k * 2)) = Ag[m * 16 + k];
  }
  for (int i = tid; i < 2 * 16 * 16; i += 128) {
    const int blk = i / 256, r = (i % 256) / 16, c = i % 16;      // block-local row-major 16 x 16, 32-byte rows
    *reinterpret_cast<bf16*>(sB + blk * 512 + swz32(r * 32 + c * 2)) = Bg[i];
  }
  ptx::fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    const uint32_t idesc = ptx::make_idesc_f16(true, false, b_mn != 0, 128, 32);
    const uint64_t adesc = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 256, ptx::SWZ_32B);
    // K-major B: the second block continues the N rows (two more 8-row groups, SBO apart).  MN-major B: the second
    // block is the next 16-element N atom, LBO = 512 bytes away.
    const uint64_t bdesc = b_mn ? ptx::make_smem_desc(ptx::smem_u32(sB), 512, 256, ptx::SWZ_32B)
                                : ptx::make_smem_desc(ptx::smem_u32(sB), 16, 256, ptx::SWZ_32B);
    ptx::mma_ss(tmem, adesc, bdesc, idesc, 0);
    ptx::tc_commit(&bar_mma);
  }
  if (!ptx::mbar_wait(&bar_mma, 0)) { if (tid == 0) status[0] = 2; }
  ptx::tc_fence_after();
  uint32_t r[32];
  ptx::tmem_ld_x32(tmem + ((uint32_t)(warp * 32) << 16), r);
  ptx::tmem_ld_wait();
  for (int j = 0; j < 32; ++j) Dg[tid * 32 + j] = __uint_as_float(r[j]);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 32);
}

// ----------------------------------------------------------------------------------------------
// throughput: the elected lane of warp 1 issues `iters` groups of MMAs (warp-uniform control flow,
// descriptors advanced by adding constants, as a production issue loop does); reports SM cycles per MMA.
enum { B_SS = 0, B_SS_ROT_A, B_SS_COLLECT, B_TS, B_CP_ONLY, B_CP_TS };

template <int MODE, int N, int REUSE>
__global__ void __launch_bounds__(128) probe_bench(int iters, long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  uint8_t* sA = smem;                        // 8 x 8 KB A tiles (128 x 32, SW64)
  uint8_t* sB = smem + 65536;                // 4 x 16 KB B tiles (N x 32, SW64)
  for (int i = tid; i < (65536 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base_s, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr int NACC = (448 / N) > 0 ? (448 / N) : 1;     // accumulators in columns 64.., A staging in 0..63
  if (warp == 1) {
    const uint64_t a_base = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 512, ptx::SWZ_64B);
    const uint64_t b_base = ptx::make_smem_desc(ptx::smem_u32(sB), 16, 512, ptx::SWZ_64B);
    long long t0 = clock64();
    if (ptx::elect_one()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {                      // 8 groups per iteration, each REUSE MMAs on one A
          const uint64_t adesc = a_base + (uint64_t)(u * (8192 >> 4));
#pragma unroll
          for (int j = 0; j < REUSE; ++j) {
            const int q = u * REUSE + j;
            const uint32_t d = tmem + 64 + (uint32_t)((q % NACC) * N);
            const uint64_t bdesc = b_base + (uint64_t)((q & 3) * (16384 >> 4));
            if (MODE == B_SS) ptx::mma_ss(d, a_base, bdesc, idesc, 1);
            if (MODE == B_SS_ROT_A) ptx::mma_ss(d, a_base + (uint64_t)((q & 7) * (8192 >> 4)), bdesc, idesc, 1);
            if (MODE == B_SS_COLLECT) {
              if (REUSE == 1) ptx::mma_ss(d, adesc, bdesc, idesc, 1);
              else if (j == 0) ptx::mma_ss_a_fill(d, adesc, bdesc, idesc, 1);
              else if (j == REUSE - 1) ptx::mma_ss_a_lastuse(d, adesc, bdesc, idesc, 1);
              else ptx::mma_ss_a_use(d, adesc, bdesc, idesc, 1);
            }
            if (MODE == B_TS) ptx::mma_ts(d, tmem + u * 8, bdesc, idesc, 1);
            if (MODE == B_CP_ONLY) ptx::tc_cp_128x256b(tmem + (q & 7) * 8, adesc);
            if (MODE == B_CP_TS) {
              if (j == 0) ptx::tc_cp_128x256b(tmem + u * 8, adesc);
              ptx::mma_ts(d, tmem + u * 8, bdesc, idesc, 1);
            }
          }
        }
      }
      ptx::tc_commit(&bar);
    }
    __syncwarp();
    if (!ptx::mbar_wait(&bar, 0)) status[0] = 3;
    long long t1 = clock64();
    if (tid == 32) cycles[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
  return (EncodeFn)fn;
}

static CUtensorMap make_map(EncodeFn enc, void* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer,
                            CUtensorMapSwizzle swz) {
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(2); }
  return m;
}

static int run_gemm(int test, const char* name) {
  EncodeFn enc = get_encode();
  std::vector<float> A(128 * 32), B(32 * 32);            // A[m][k], B[k][n] logical
  srand(1234 + test);
  for (auto& v : A) v = (float)((rand() % 17) - 8) / 8.f;
  for (auto& v : B) v = (float)((rand() % 13) - 6) / 4.f;
  const bool a_mn = (test == T_TMA_MNK);
  const bool b_mn = (test == T_SS_KMN || test == T_TMA_MNK);
  std::vector<bf16> Ah(128 * 32), Bh(32 * 32);
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 32; ++k) {
      if (!a_mn) Ah[m * 32 + k] = __float2bfloat16(A[m * 32 + k]);
      else       Ah[k * 128 + m] = __float2bfloat16(A[m * 32 + k]);     // global [k][m], m contiguous
    }
  for (int k = 0; k < 32; ++k)
    for (int n = 0; n < 32; ++n) {
      if (b_mn) Bh[k * 32 + n] = __float2bfloat16(B[k * 32 + n]);        // [k][n], n contiguous
      else      Bh[n * 32 + k] = __float2bfloat16(B[k * 32 + n]);        // [n][k], k contiguous
    }
  bf16 *Ad, *Bd; float* Dd; int* st;
  CK(cudaMalloc(&Ad, Ah.size() * 2)); CK(cudaMalloc(&Bd, Bh.size() * 2)); CK(cudaMalloc(&Dd, 128 * 32 * 4)); CK(cudaMalloc(&st, 4));
  CK(cudaMemcpy(Ad, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(Bd, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(Dd, 0, 128 * 32 * 4)); CK(cudaMemset(st, 0, 4));
  Maps maps;
  if (!a_mn) maps.a = make_map(enc, Ad, 32, 128, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B);
  else       maps.a = make_map(enc, Ad, 128, 32, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  maps.b = make_map(enc, Bd, 32, 32, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  CK(cudaFuncSetAttribute(probe_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  probe_gemm<<<1, 128, 16384>>>(test, Ad, Bd, Dd, st, maps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-8s LAUNCH ERROR %s\n", name, cudaGetErrorString(e)); return 1; }
  std::vector<float> D(128 * 32); int status = 0;
  CK(cudaMemcpy(D.data(), Dd, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&status, st, 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 32; ++n) {
      double r = 0;
      for (int k = 0; k < 32; ++k) r += (double)A[m * 32 + k] * (double)B[k * 32 + n];
      maxerr = fmax(maxerr, fabs(r - D[m * 32 + n]));
      maxref = fmax(maxref, fabs(r));
    }
  printf("%-8s status=%d max_abs_err=%.4g (max |ref| %.3g) %s\n", name, status, maxerr, maxref,
         (status == 0 && maxerr < 1e-3) ? "PASS" : "FAIL");
  return (status == 0 && maxerr < 1e-3) ? 0 : 1;
}

static int run_gemm16(int b_mn, const char* name) {
  std::vector<float> A(128 * 16), B(16 * 32);             // A[m][k], B[k][n] logical with n = 0..31 over two blocks
  srand(4321 + b_mn);
  for (auto& v : A) v = (float)((rand() % 17) - 8) / 8.f;
  for (auto& v : B) v = (float)((rand() % 13) - 6) / 4.f;
  std::vector<bf16> Ah(128 * 16), Bh(2 * 256);
  for (int i = 0; i < 128 * 16; ++i) Ah[i] = __float2bfloat16(A[i]);
  for (int blk = 0; blk < 2; ++blk)
    for (int k = 0; k < 16; ++k)
      for (int n = 0; n < 16; ++n) {
        const float v = B[k * 32 + blk * 16 + n];
        if (b_mn) Bh[blk * 256 + k * 16 + n] = __float2bfloat16(v);      // [k][n]
        else      Bh[blk * 256 + n * 16 + k] = __float2bfloat16(v);      // [n][k]
      }
  bf16 *Ad, *Bd; float* Dd; int* st;
  CK(cudaMalloc(&Ad, Ah.size() * 2)); CK(cudaMalloc(&Bd, Bh.size() * 2)); CK(cudaMalloc(&Dd, 128 * 32 * 4)); CK(cudaMalloc(&st, 4));
  CK(cudaMemcpy(Ad, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(Bd, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(Dd, 0, 128 * 32 * 4)); CK(cudaMemset(st, 0, 4));
  probe_gemm16<<<1, 128, 8192>>>(b_mn, Ad, Bd, Dd, st);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-8s LAUNCH ERROR %s\n", name, cudaGetErrorString(e)); return 1; }
  std::vector<float> D(128 * 32); int status = 0;
  CK(cudaMemcpy(D.data(), Dd, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&status, st, 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 32; ++n) {
      double r = 0;
      for (int k = 0; k < 16; ++k) r += (double)A[m * 16 + k] * (double)B[k * 32 + n];
      maxerr = fmax(maxerr, fabs(r - D[m * 32 + n]));
      maxref = fmax(maxref, fabs(r));
    }
  printf("%-8s status=%d max_abs_err=%.4g (max |ref| %.3g) %s\n", name, status, maxerr, maxref,
         (status == 0 && maxerr < 1e-3) ? "PASS" : "FAIL");
  return (status == 0 && maxerr < 1e-3) ? 0 : 1;
}

template <int MODE, int N, int REUSE>
static void bench_one(const char* name, long long* cyc, int* st) {
  const int iters = 256;
  CK(cudaFuncSetAttribute(probe_bench<MODE, N, REUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
  for (int grid : {1, 148}) {
    CK(cudaMemset(cyc, 0, 148 * 8));
    probe_bench<MODE, N, REUSE><<<grid, 128, 131072>>>(iters, cyc, st);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("bench %-22s grid=%d LAUNCH ERROR %s\n", name, grid, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(148);
    CK(cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    int status; CK(cudaMemcpy(&status, st, 4, cudaMemcpyDeviceToHost));
    const double per = (double)mx / (iters * 8.0 * REUSE);
    const double ideal = MODE == B_CP_ONLY ? 0 : 128.0 * N / 256.0;
    printf("bench %-22s grid=%3d cycles/op=%7.2f  ideal_mma=%5.1f  eff=%5.1f%% status=%d\n", name, grid, per, ideal,
           ideal > 0 ? 100.0 * ideal / per : 0.0, status);
  }
}

static void run_bench() {
  long long* cyc; int* st;
  CK(cudaMalloc(&cyc, 148 * 8)); CK(cudaMalloc(&st, 4)); CK(cudaMemset(st, 0, 4));
  bench_one<B_SS, 32, 1>("ss same-A N=32", cyc, st);
  bench_one<B_SS, 64, 1>("ss same-A N=64", cyc, st);
  bench_one<B_SS, 128, 1>("ss same-A N=128", cyc, st);
  bench_one<B_SS, 256, 1>("ss same-A N=256", cyc, st);
  bench_one<B_SS_ROT_A, 32, 1>("ss rot-A N=32", cyc, st);
  bench_one<B_SS_ROT_A, 64, 1>("ss rot-A N=64", cyc, st);
  bench_one<B_SS_ROT_A, 128, 1>("ss rot-A N=128", cyc, st);
  bench_one<B_SS_ROT_A, 256, 1>("ss rot-A N=256", cyc, st);
  bench_one<B_SS_COLLECT, 32, 2>("ss collect r=2 N=32", cyc, st);
  bench_one<B_SS_COLLECT, 32, 4>("ss collect r=4 N=32", cyc, st);
  bench_one<B_SS_COLLECT, 32, 8>("ss collect r=8 N=32", cyc, st);
  bench_one<B_SS_COLLECT, 64, 4>("ss collect r=4 N=64", cyc, st);
  bench_one<B_TS, 32, 4>("ts N=32", cyc, st);
  bench_one<B_TS, 64, 4>("ts N=64", cyc, st);
  bench_one<B_TS, 128, 4>("ts N=128", cyc, st);
  bench_one<B_CP_ONLY, 32, 1>("cp 128x256b only", cyc, st);
  bench_one<B_CP_TS, 32, 1>("cp+ts r=1 N=32", cyc, st);
  bench_one<B_CP_TS, 32, 2>("cp+ts r=2 N=32", cyc, st);
  bench_one<B_CP_TS, 32, 4>("cp+ts r=4 N=32", cyc, st);
  bench_one<B_CP_TS, 32, 8>("cp+ts r=8 N=32", cyc, st);
  bench_one<B_CP_TS, 64, 4>("cp+ts r=4 N=64", cyc, st);
}

int main(int argc, char** argv) {
  const char* t = argc > 1 ? argv[1] : "all";
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d SMs=%d\n", p.name, p.major, p.minor, p.multiProcessorCount);
  if (!strcmp(t, "ss_kk")) return run_gemm(T_SS_KK, t);
  if (!strcmp(t, "ss_kmn")) return run_gemm(T_SS_KMN, t);
  if (!strcmp(t, "tma_kk")) return run_gemm(T_TMA_KK, t);
  if (!strcmp(t, "tma_mnk")) return run_gemm(T_TMA_MNK, t);
  if (!strcmp(t, "ts_st")) return run_gemm(T_TS_ST, t);
  if (!strcmp(t, "ts_cp")) return run_gemm(T_TS_CP, t);
  if (!strcmp(t, "ss16_kk")) return run_gemm16(0, t);       // 16x16 blocks, SWIZZLE_32B, K-major B (not yet run on a B200)
  if (!strcmp(t, "ss16_kmn")) return run_gemm16(1, t);      // 16x16 blocks, SWIZZLE_32B, MN-major B
  if (!strcmp(t, "bench")) { run_bench(); return 0; }
  printf("unknown test %s\n", t);
  return 2;
}
