// L2 -> shared-memory ingest probe for sm_100a: how fast can persistent CTAs pull tiles out of an L2-resident buffer
// with TMA, as a function of box shape (64-byte vs 128-byte rows, 2 KB vs 8 KB boxes), stages in flight, CTAs per SM and
// cluster multicast?  Both tcgen05 GEMM families of this repo sit at ~12 TB/s of TMA loads (profiles/r1_ncu_tc_kernels.txt);
// this tells whether that is the wall of the fabric, of the 64-byte rows, or of the bytes in flight.
// No compute, no tensor cores: producer lane issues loads into a ring, the same lane waits and re-arms.
// Every wait is wall-clock bounded (ptx::mbar_wait), so a protocol error prints TIMEOUT instead of hanging the GPU.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mem_probe tools/mem_probe.cu
//   tools/mem_probe            (prints one line per configuration)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../blocksparse_b200/csrc/ptx.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

namespace {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load delivered to every CTA of `mask` at the same CTA-relative shared-memory offset; each destination's mbarrier
// (same offset) receives the complete_tx for the bytes that landed there.
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* src, uint64_t* bar, uint32_t bytes) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(ptx::smem_u32(bar)) : "memory");
}

struct Params {
  const uint8_t* src;
  int box_bytes;        // bytes per TMA box
  int boxes_per_stage;  // boxes landing on one barrier
  int stages;           // ring depth
  int iters;            // stages filled per CTA
  int n_boxes;          // distinct boxes in the source (walk is strided so CTAs do not share lines)
  int rows_per_box;     // outer extent of a box (coordinate step)
  int multicast;        // 0 = every CTA loads its own tile; 1 = cluster, each CTA loads 1/G of the stage and multicasts it; 2 = cluster, every CTA loads everything (same tiles)
  int G;                // cluster size
  int bulk1d;           // 1 = plain cp.async.bulk of box_bytes contiguous bytes instead of a tensor-map box
};

// free-running ring: lane 0 of warp 0 keeps `stages` stages in flight
__global__ void __launch_bounds__(32) ingest_kernel(const __grid_constant__ CUtensorMap map, Params p, long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[16];
  __shared__ int abort_s;
  const int lane = threadIdx.x;
  if (lane == 0) {
    abort_s = 0;
    for (int i = 0; i < p.stages; ++i) ptx::mbar_init(&full[i], 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();
  if (p.multicast) cluster_sync();
  const uint32_t stage_bytes = (uint32_t)p.box_bytes * p.boxes_per_stage;
  const uint32_t rank = p.multicast ? cluster_ctarank() : 0;
  const long long t0 = clock64();
  if (lane == 0) {
    // the tile a CTA (or a cluster) pulls at step i: a different box every time, spread over the whole source
    const unsigned base = (p.multicast ? blockIdx.x / p.G : blockIdx.x) * 7919u;
    auto issue = [&](int i) {
      const int st = i % p.stages;
      uint8_t* dst = smem + st * stage_bytes;
      ptx::mbar_expect_tx(&full[st], stage_bytes);
      if (!p.multicast) {
        for (int b = 0; b < p.boxes_per_stage; ++b) {
          const unsigned box = (base + (unsigned)(i * p.boxes_per_stage + b) * 13u) % (unsigned)p.n_boxes;
          if (p.bulk1d) bulk_load_1d(dst + b * p.box_bytes, p.src + (size_t)box * p.box_bytes, &full[st], p.box_bytes);
          else ptx::tma_load_2d(dst + b * p.box_bytes, &map, &full[st], 0, (int)box * p.rows_per_box);
        }
      } else {
        // each CTA fetches every other box of the stage and delivers it to both CTAs
        for (int b = (int)rank; b < p.boxes_per_stage; b += p.G) {
          const unsigned box = (base + (unsigned)(i * p.boxes_per_stage + b) * 13u) % (unsigned)p.n_boxes;
          tma_load_2d_multicast(dst + b * p.box_bytes, &map, &full[st], 0, (int)box * p.rows_per_box, (uint16_t)((1u << p.G) - 1));
        }
      }
    };
    bool ok = true;
    if (!p.multicast) {
      for (int i = 0; i < p.stages && i < p.iters; ++i) issue(i);
      for (int i = 0; i < p.iters && ok; ++i) {
        ok = ptx::mbar_wait(&full[i % p.stages], (i / p.stages) & 1, &abort_s);
        if (ok && i + p.stages < p.iters) issue(i + p.stages);
      }
    }
    if (!ok) status[0] = 1;
  }
  if (p.multicast) {
    // Round-synchronous version for the cluster: arm all stages, cluster barrier (the peer's multicast must not
    // reach a barrier before it is armed), issue, wait for all stages, cluster barrier.  The same round structure
    // is used for the unicast reference (multicast == 2) so the two compare like for like.
    bool ok = true;
    for (int r0 = 0; r0 < p.iters && ok; r0 += p.stages) {
      cluster_sync();
      if (lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
          uint8_t* dst = smem + s * stage_bytes;
          ptx::mbar_expect_tx(&full[s], stage_bytes);
          const unsigned base = (blockIdx.x / p.G) * 7919u;
          for (int b = (p.multicast == 1 ? (int)rank : 0); b < p.boxes_per_stage; b += (p.multicast == 1 ? p.G : 1)) {
            const unsigned box = (base + (unsigned)((r0 + s) * p.boxes_per_stage + b) * 13u) % (unsigned)p.n_boxes;
            if (p.multicast == 1) tma_load_2d_multicast(dst + b * p.box_bytes, &map, &full[s], 0, (int)box * p.rows_per_box, (uint16_t)((1u << p.G) - 1));
            else                  ptx::tma_load_2d(dst + b * p.box_bytes, &map, &full[s], 0, (int)box * p.rows_per_box);
          }
        }
        for (int s = 0; s < p.stages && ok; ++s) ok = ptx::mbar_wait(&full[s], (r0 / p.stages) & 1, &abort_s);
        if (!ok) status[0] = 2;
      }
      __syncwarp();
    }
    cluster_sync();
  }
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
  return (EncodeFn)fn;
}

struct Shape { const char* name; int inner_elems, rows; CUtensorMapSwizzle swz; };

void run(EncodeFn enc, void* src, size_t src_bytes, const Shape& sh, int boxes_per_stage, int stages, int ctas_per_sm, int multicast,
         long long* cyc, int* st, int sm_count, double clock_ghz, int G = 2, int bulk1d = 0) {
  const int box_bytes = sh.inner_elems * 2 * sh.rows;
  // source viewed as [n_rows][row_elems]: row pitch 8 KB like the activation matrix of BASELINE cfg 2
  const uint64_t row_elems = 4096;
  const uint64_t n_rows = src_bytes / (row_elems * 2);
  CUtensorMap m;
  cuuint64_t dims[2] = {row_elems, n_rows};
  cuuint64_t strides[1] = {row_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)sh.inner_elems, (cuuint32_t)sh.rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sh.swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(2); }
  Params p;
  p.box_bytes = box_bytes; p.boxes_per_stage = boxes_per_stage; p.stages = stages;
  p.iters = 4096 / boxes_per_stage / stages * stages;      // a multiple of the ring depth
  p.rows_per_box = sh.rows; p.n_boxes = (int)(n_rows / sh.rows); p.multicast = multicast; p.G = multicast ? G : 1; p.bulk1d = bulk1d;
  p.src = (const uint8_t*)src;
  if (bulk1d) p.n_boxes = (int)(src_bytes / box_bytes);
  const size_t smem = (size_t)box_bytes * boxes_per_stage * stages;
  const int grid = multicast ? (sm_count / G) * G * ctas_per_sm : sm_count * ctas_per_sm;
  CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaMemset(st, 0, 4));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = multicast ? G : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {                        // first pass warms L2
    CK(cudaLaunchKernelEx(&cfg, ingest_kernel, m, p, cyc, st));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s LAUNCH ERROR %s\n", sh.name, cudaGetErrorString(e)); return; }
  }
  std::vector<long long> h(grid);
  CK(cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
  int status; CK(cudaMemcpy(&status, st, 4, cudaMemcpyDeviceToHost));
  long long mx = 0; for (long long v : h) mx = v > mx ? v : mx;
  const double bytes_per_cta = (double)p.iters * box_bytes * boxes_per_stage;       // bytes that land in each CTA
  const double per_sm = bytes_per_cta * ctas_per_sm / (double)mx;                    // bytes / cycle / SM
  printf("%-28s%s G%d boxes/stage %d stages %2d ctas/sm %d %-9s | %6.1f B/clk/SM  %6.2f TB/s at %.2f GHz  (%s)\n", sh.name, bulk1d ? " bulk1d" : "", p.G, boxes_per_stage, stages,
         ctas_per_sm, multicast == 1 ? "multicast" : multicast == 2 ? "rounds" : "ring", per_sm, per_sm * grid / ctas_per_sm * clock_ghz / 1e3, clock_ghz,
         status ? "TIMEOUT" : "ok");
}

}  // namespace

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs=%d L2=%d MB\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.l2CacheSize >> 20);
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double ghz = khz / 1e6;
  const size_t src_bytes = 64ull << 20;                      // L2-resident (126 MB L2)
  void* src; CK(cudaMalloc(&src, src_bytes)); CK(cudaMemset(src, 1, src_bytes));
  long long* cyc; int* st;
  CK(cudaMalloc(&cyc, 1024 * 8)); CK(cudaMalloc(&st, 4));
  EncodeFn enc = get_encode();
  const int sms = prop.multiProcessorCount;
  const Shape x64   = {"X tile 128 x 64B  (SW64)",  32, 128, CU_TENSOR_MAP_SWIZZLE_64B};     // what xprop loads for bs 32, axis 1
  const Shape x128  = {"X tile  64 x 128B (SW128)", 64,  64, CU_TENSOR_MAP_SWIZZLE_128B};    // same bytes, 128-byte rows
  const Shape w64   = {"W block 32 x 64B  (SW64)",  32,  32, CU_TENSOR_MAP_SWIZZLE_64B};     // 2 KB
  const Shape w32   = {"W block 16 x 32B  (SW32)",  16,  16, CU_TENSOR_MAP_SWIZZLE_32B};     // 512 B (bs 16)
  for (int ctas : {1, 2}) {
    for (int stages : {2, 4, 8, 12}) {
      if ((size_t)stages * 8192 * ctas > 200 * 1024) continue;
      run(enc, src, src_bytes, x64, 1, stages, ctas, 0, cyc, st, sms, ghz);
      run(enc, src, src_bytes, x128, 1, stages, ctas, 0, cyc, st, sms, ghz);
    }
    run(enc, src, src_bytes, w64, 4, 8, ctas, 0, cyc, st, sms, ghz);
    run(enc, src, src_bytes, w32, 8, 8, ctas, 0, cyc, st, sms, ghz);
  }
  // 1-D bulk copies of contiguous 2 KB (a W block) and 8 KB
  const Shape b2k = {"bulk 2 KB", 32, 32, CU_TENSOR_MAP_SWIZZLE_64B};
  const Shape b8k = {"bulk 8 KB", 32, 128, CU_TENSOR_MAP_SWIZZLE_64B};
  const Shape x128w = {"X tile 128 x 128B (SW128)", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B};   // two input blocks wide, 16 KB
  for (int ctas : {1, 2}) {
    run(enc, src, src_bytes, b2k, 4, 8, ctas, 0, cyc, st, sms, ghz, 1, 1);
    run(enc, src, src_bytes, b8k, 1, 8, ctas, 0, cyc, st, sms, ghz, 1, 1);
    run(enc, src, src_bytes, x128w, 1, 6, ctas, 0, cyc, st, sms, ghz);
  }
  // clusters of G CTAs pulling the SAME X tiles (8 KB boxes, 8 per stage): multicast vs everyone loads everything
  for (int G : {2, 4, 8}) {
    run(enc, src, src_bytes, x64, 8, 2, 1, 2, cyc, st, sms, ghz, G);
    run(enc, src, src_bytes, x64, 8, 2, 1, 1, cyc, st, sms, ghz, G);
    run(enc, src, src_bytes, x128, 8, 2, 1, 2, cyc, st, sms, ghz, G);
    run(enc, src, src_bytes, x128, 8, 2, 1, 1, cyc, st, sms, ghz, G);
  }
  // cluster of two CTAs on neighbouring SMs pulling the SAME tiles: every CTA fetches half and multicasts, against
  // both CTAs fetching everything, in the same round-synchronous structure
  for (int stages : {4, 8}) {
    run(enc, src, src_bytes, w64, 4, stages, 1, 2, cyc, st, sms, ghz);
    run(enc, src, src_bytes, w64, 4, stages, 1, 1, cyc, st, sms, ghz);
  }
  return 0;
}
