// tcgen05 (5th-gen tensor core) kernel families for sm_100a.
//
// tc_xprop: block-sparse fprop / bprop for 16-bit dtypes, feature_axis = 1, block size 32 / 64.
//   Replaces hgemm_blocksparse_64x64x64_nx_dsd / 64x32x32_nx_dsd
//   (reference src/blocksparse_hgemm_nc_op_gpu.cu:38-551).
//
// Formulation (see DESIGN.md "xprop on tcgen05"): the minibatch sits on the MMA M axis (128 rows per
// CTA tile), one W block is the B operand (N = K = block size), and an output tile covers up to
// 512/bs consecutive output blocks whose fp32 accumulators fill the CTA's tensor memory.  The tile
// schedule (blocksparse_b200/lut.py:build_tile_schedule) groups the LUT by INPUT block, so an
// activation tile is staged once by TMA and every block that consumes it is issued back to back with
// the A-operand collector hints (fill / use / lastuse): the B200 probe (profiles/r1_tc_probe.txt)
// shows an N=32 MMA is shared-memory bound at 40 cycles when A is re-read, 16 + 27/R with reuse R.
//
// Warp roles (persistent over tiles, (2*NP + 4) warps, see the pipeline protocol below):
//   warps 0..NP-1     TMA producers, one per pipeline (activation tile + W blocks of a group -> one mbarrier)
//   warps NP..2NP-1   MMA issuers, one per pipeline (one elected thread each); the first also owns TMEM alloc
//   last 4 warps      epilogue: tcgen05.ld -> 16-bit -> swizzled smem staging -> TMA store, then clear the lanes
#pragma once
#include <cuda.h>
#include <atomic>
#include "common.cuh"
#include "ptx.cuh"

namespace bsmm {
constexpr int TC_NOT_APPLICABLE = -1000;

// ---- TMA descriptor encoding via the driver entry point (no -lcuda link dependency) -----------
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = []() -> TmapEncodeFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    return (TmapEncodeFn)p;
  }();
  return fn;
}
// 2-D row-major 16-bit tensor [outer][inner]; box = box_outer rows x box_inner elements.
inline int make_tmap_2d(CUtensorMap* m, int dtype, const void* base, uint64_t inner, uint64_t outer, uint64_t row_pitch_elems,
                        uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz) {
  TmapEncodeFn enc = tmap_encoder();
  if (!enc) return fail(BSMM_E_NODEV, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = dtype == BSMM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(BSMM_E_ARG, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// Encoding a CUtensorMap costs a few microseconds of host time and the same (pointer, shape) tuples come back every
// training step, so the encoded descriptors are kept in a small thread-local cache (round-robin replacement).  The
// descriptors are still passed to the kernels by value (__grid_constant__), so launches stay CUDA-graph capturable.
struct TmapKey {
  const void* base; uint64_t inner, outer, pitch; uint32_t box_inner, box_outer; int dtype, swz, dev;
  bool operator==(const TmapKey& o) const {
    return base == o.base && inner == o.inner && outer == o.outer && pitch == o.pitch && box_inner == o.box_inner &&
           box_outer == o.box_outer && dtype == o.dtype && swz == o.swz && dev == o.dev;
  }
};
constexpr int TMAP_CACHE_SLOTS = 32;
struct TmapCache { TmapKey key[TMAP_CACHE_SLOTS]; CUtensorMap map[TMAP_CACHE_SLOTS]; int used = 0, next = 0; };
inline int cached_tmap_2d(CUtensorMap* m, int dtype, const void* base, uint64_t inner, uint64_t outer, uint64_t row_pitch_elems,
                          uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz) {
  static thread_local TmapCache cache;
  int dev = 0;
  cudaGetDevice(&dev);
  const TmapKey k = {base, inner, outer, row_pitch_elems, box_inner, box_outer, dtype, (int)swz, dev};
  for (int i = 0; i < cache.used; ++i)
    if (cache.key[i] == k) { *m = cache.map[i]; return 0; }
  if (int e = make_tmap_2d(m, dtype, base, inner, outer, row_pitch_elems, box_inner, box_outer, swz)) return e;
  const int slot = cache.next;
  cache.key[slot] = k; cache.map[slot] = *m;
  cache.next = (slot + 1) % TMAP_CACHE_SLOTS;
  if (cache.used < TMAP_CACHE_SLOTS) ++cache.used;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Pipeline protocol.  One pipeline stage == one schedule GROUP (lut.py:build_tile_schedule): the
// activation tile of one input block plus the <= WPS W blocks of the output tile that consume it, all
// landing on ONE mbarrier.  Every per-block decision (accumulator column, accumulate flag, A-collector
// hint, merging of adjacent blocks into one wider MMA) is precomputed on the host into a 128-byte
// group record, because the first profile (profiles/r1_xprop_v1_ncu.txt) showed both the producer and
// the single issuing thread instruction-bound at hundreds of cycles per block when they derived them.
//   producer warp p (groups gc with gc % NP == p): one coalesced 128-byte load per group, prefetched a
//     group ahead; lanes 4..11 each issue one W TMA load; lanes 12..19 write the run commands to smem
//   issuer warp p (same groups): one barrier wait per group, LDS.128 commands, then one tcgen05.mma per
//     run and K slice; the NP issuers take turns in group order (turn[] mbarriers) so results are deterministic
// Pipeline p owns the stages with index % NP == p: every mbarrier has exactly one waiter that walks its phases
// in order (with stages handed round-robin to whichever warp came next, a waiter two phases behind would pass
// a parity test early).
// OCC = CTAs per SM.  OCC 2 halves the output tile (256 TMEM columns) so that the epilogue of one CTA overlaps
// the main loop of the other.
// VAR 1 = "sparse" variant for layouts with ~1 W block per group (density <= ~12 %): 2 W slots per stage instead of
// 8, which doubles the number of stages in flight -- at 5-10 % density the kernel is bound by the TMA round trip.
template <int BS, int OCC, int VAR = 0> struct XpropCfg;
template <> struct XpropCfg<32, 2, 1> {
  static constexpr int XS = 6, WPS = 2, STG = 4, NP = 2, TCOLS = 256;
  static constexpr uint32_t SWZ = ptx::SWZ_64B, SBO = 512;
};
template <> struct XpropCfg<32, 2, 3> {
  static constexpr int XS = 6, WPS = 4, STG = 2, NP = 2, TCOLS = 256;
  static constexpr uint32_t SWZ = ptx::SWZ_64B, SBO = 512;
};
template <> struct XpropCfg<32, 1> {
  static constexpr int XS = 6, WPS = 8, STG = 8, NP = 3, TCOLS = 512;  // group stages, W slots per stage, staging buffers, TMEM columns
  static constexpr uint32_t SWZ = ptx::SWZ_64B, SBO = 512;     // 64-byte rows
};
// STG == 0 selects a direct epilogue (rows stored straight from registers, no staging).  Measured on B200
// (profiles/r1_xprop_tuning.txt): XS=6/WPS=4/STG=0 is ~7 % SLOWER at 25 % density and 23 % slower at 100 % than
// XS=3/WPS=8/STG=4, so the staged TMA-store epilogue stays.
template <> struct XpropCfg<32, 2> {
  static constexpr int XS = 4, WPS = 8, STG = 2, NP = 4, TCOLS = 256;
  static constexpr uint32_t SWZ = ptx::SWZ_64B, SBO = 512;
};
// 16 x 16 blocks: 16-block tiles (256 TMEM columns), 32-byte operand rows (SWIZZLE_32B), one K=16 slice per block
template <> struct XpropCfg<16, 2> {
  static constexpr int XS = 8, WPS = 8, STG = 4, NP = 4, TCOLS = 256;
  static constexpr uint32_t SWZ = ptx::SWZ_32B, SBO = 256;     // 8 rows of 32 bytes
};
template <> struct XpropCfg<64, 1> {
  static constexpr int XS = 4, WPS = 4, STG = 2, NP = 2, TCOLS = 512;
  static constexpr uint32_t SWZ = ptx::SWZ_128B, SBO = 1024;   // 128-byte rows
};
template <> struct XpropCfg<64, 2> {
  static constexpr int XS = 2, WPS = 2, STG = 2, NP = 2, TCOLS = 256;
  static constexpr uint32_t SWZ = ptx::SWZ_128B, SBO = 1024;
};
// Cfg::NP = number of pipelines: NP producer warps + NP MMA-issuing warps (pipeline p = producer p -> issuer p,
// groups dealt round-robin) + 4 epilogue warps.
template <class Cfg> constexpr int xprop_threads() { return (2 * Cfg::NP + 4) * 32; }

// sticky device-side error word: a kernel whose bounded wait timed out stores a non-zero code here
__device__ int g_tc_error = 0;

// ---- dynamic tile queue -------------------------------------------------------------------------
// The persistent grids used to deal tile t to CTA t mod grid.  That is only optimal when every CTA of the grid is
// resident from the first cycle: when a concurrent kernel (the NCCL all-reduce of the previous step's dW) holds a few
// SMs, the displaced CTAs start after a first-wave CTA retires and the kernel takes twice as long.  Now CTAs pull tile
// indices from a global counter (one atomicAdd per tile) in a host-chosen order (heaviest tiles first for skewed
// layouts): late CTAs simply take fewer tiles.  Results do not depend on which CTA computes a tile, so they stay
// bit-reproducible.  One counter pair per launch, taken round-robin from a pool; the last CTA to leave resets its slot.
constexpr int TILE_RING = 4;
constexpr int TILE_COUNTER_SLOTS = 64;
__device__ int g_tile_counters[TILE_COUNTER_SLOTS][2];
// BSMM_TILE_QUEUE=dynamic turns the global-counter queue on (blocksparse_b200.dist.reserve_sms_for_nccl does it for
// multi-GPU runs); the default is the static deal tile k of a CTA = blockIdx.x + k * gridDim.x, which measured 4-8 % faster
// when the grid has the GPU to itself (profiles/r2_tile_queue.txt).
inline bool static_tiles() {
  static const bool v = [] { const char* e = getenv("BSMM_TILE_QUEUE"); return !(e && (e[0] == 'd' || e[0] == 'D')); }();
  return v;
}
inline bool static_order() { static const bool v = [] { const char* e = getenv("BSMM_NATURAL_ORDER"); return e && atoi(e) != 0; }(); return v || static_tiles(); }
inline int* next_tile_counter() {
  static thread_local int* base[64] = {};              // per device: the symbol lives in every device's module image
  static std::atomic<unsigned> global_id{0};
  int dev = 0;
  cudaGetDevice(&dev);
  int*& b = base[dev & 63];
  if (!b) { void* ptr = nullptr; if (cudaGetSymbolAddress(&ptr, g_tile_counters) != cudaSuccess) return nullptr; b = (int*)ptr; }
  const unsigned id = global_id.fetch_add(1, std::memory_order_relaxed);
  return b + 2 * (id % TILE_COUNTER_SLOTS);
}
struct TileQueue {
  int ring[TILE_RING];
  uint64_t ready[TILE_RING], freed[TILE_RING];
};
__device__ __forceinline__ void tile_queue_init(TileQueue* q, int n_warps) {      // every warp of the CTA reads every slot
  for (int i = 0; i < TILE_RING; ++i) { ptx::mbar_init(&q->ready[i], 1); ptx::mbar_init(&q->freed[i], n_warps); }
}
// Reader side, called by every warp of every role with the same running tile count k: the k-th tile of this CTA, or
// -1 when the queue is drained (or a wait timed out).
__device__ __forceinline__ int tile_queue_next(TileQueue* q, uint32_t k, int lane, volatile int* abort_flag) {
  const uint32_t slot = k % TILE_RING;
  if (!ptx::mbar_wait(&q->ready[slot], (k / TILE_RING) & 1, abort_flag)) return -1;
  const int t = *reinterpret_cast<volatile int*>(&q->ring[slot]);
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(&q->freed[slot]);
  return t;
}
// Fetcher side (lane 0 of ONE warp of the CTA).  The index for tile k+1 is drawn (tile_queue_draw: one atomicAdd,
// result not consumed) when the fetcher starts working on tile k and published when it is done with it, so the two L2
// round trips (counter, order table) are off the critical path.
// The first tile of every CTA is its block index (no atomic: hundreds of CTAs hitting one address at launch serialise in
// the L2 atomic unit, ~4 us measured); later tiles are gridDim.x + a draw from the counter.
// counter == nullptr (BSMM_STATIC_TILES=1): the round-1 static deal, tile k of a CTA = blockIdx.x + k * gridDim.x.
__device__ __forceinline__ int tile_queue_draw(int* counter, uint32_t k_next) {
  return counter ? (int)gridDim.x + atomicAdd(counter, 1) : (int)(blockIdx.x + k_next * gridDim.x);
}
__device__ __forceinline__ void tile_queue_publish(TileQueue* q, uint32_t k, int drawn, const int32_t* order, int total, volatile int* abort_flag) {
  const uint32_t slot = k % TILE_RING;
  bool ok = true;
  if (k >= TILE_RING) ok = ptx::mbar_wait(&q->freed[slot], ((k / TILE_RING) - 1) & 1, abort_flag);   // every warp has read the slot's previous tile
  int t = -1;
  if (ok && drawn < total) t = order ? order[drawn] : drawn;
  q->ring[slot] = t;
  ptx::mbar_arrive(&q->ready[slot]);
}
// last CTA out resets the counter pair for its next user
__device__ __forceinline__ void tile_queue_retire(int* counter) {
  __threadfence();
  if (atomicAdd(counter + 1, 1) == (int)gridDim.x - 1) { counter[0] = 0; counter[1] = 0; __threadfence(); }
}

struct XpropTcParams {
  const int32_t* sched;      // tile schedule (lut.py:build_tile_schedule)
  int groups_off;            // int32 index of the first group record
  int n_ktiles;              // output tiles along the feature axis
  int n_ntiles;              // ceil(N / 128)
  int bprop;
  int axis0;                 // activations are (C, N): A operand is MN-major, output is stored transposed
  void* y;                   // output base, row pitch and row count (direct-store epilogue)
  long long y_pitch;         // elements
  int N;
  int* counter;              // tile queue: {next index, CTAs retired}
  const int32_t* order;      // tile order (heaviest first) or nullptr = natural order
};
struct XpropTmaps { CUtensorMap x, w, y; };

// CL = 2: thread-block clusters of two CTAs that work on NEIGHBOURING OUTPUT TILES (2P, 2P+1) of the same minibatch tile and
// walk one merged group list (lut.py:build_tile_schedule(pair_tiles=True)): the activation tile of group g is fetched by CTA
// g % 2 and MULTICAST to both, so the activation panel crosses the L2 -> SM fabric once per tile PAIR instead of once per
// tile -- the fabric (~12 TB/s for the chip) is what bounds this kernel (profiles/r2_xprop2_study.txt).  A stage may be
// refilled once BOTH CTAs have retired its MMAs: every issuer's tcgen05.commit arrives on the empty barrier of both CTAs.
template <int BS, bool BF16, int OCC, int VAR = 0, int CL = 1>
__global__ void __launch_bounds__((xprop_threads<XpropCfg<BS, OCC, VAR>>()), OCC)
tc_xprop_kernel(const XpropTcParams p, const __grid_constant__ XpropTmaps maps) {
  using Cfg = XpropCfg<BS, OCC, VAR>;
  constexpr int XS = Cfg::XS, WPS = Cfg::WPS, STG = Cfg::STG;
  constexpr int KS = BS / 16;                     // K=16 slices per block
  // Two independent in-order pipelines (producer warp p -> issuer warp p) share the ring: pipeline p owns the
  // stages with index % 2 == p, so every mbarrier has one waiter that walks its phases in order.
  constexpr int NP = Cfg::NP;
  static_assert(XS % NP == 0, "stages are split evenly between the pipelines");
  constexpr uint32_t HS = XS / NP;   // stages per pipeline
  constexpr uint32_t XBYTES = 128 * BS * 2, WBYTES = BS * BS * 2;
  constexpr uint32_t STAGE_BYTES = XBYTES + WPS * WBYTES;
  constexpr uint32_t ROW = BS * 2;                // bytes per smem row (== swizzle span)

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sStage = smem;                         // XS x (activation tile | WPS W blocks)
  uint8_t* sO = smem + XS * STAGE_BYTES;          // STG x XBYTES staging for the output tile
  __shared__ uint64_t full[XS], empty[XS], acc_full, acc_empty, turn[NP];
  __shared__ __align__(16) int4 cmd[XS][8];       // per run: (B descriptor low word for K slice 0, D tmem address, idesc, accumulate)
  __shared__ uint32_t tmem_base_s;
  __shared__ int abort_s;
  __shared__ TileQueue tq;
  volatile int* abort_flag = &abort_s;

  // the warp index is made provably warp-uniform (shuffle from lane 0) so that role dispatch, stage indices and descriptor arithmetic
  // can live in uniform registers
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid / 32, 0), lane = tid % 32;
  // CL == 2: the queue hands out tile PAIRS (index = minibatch tile * n_ktiles/2 + pair); this CTA takes output tile 2*pair + rank
  const uint32_t crank = CL == 2 ? ptx::cluster_ctarank() : 0u;
  const int kt_per = CL == 2 ? p.n_ktiles / 2 : p.n_ktiles;
  const int total_tiles = kt_per * p.n_ntiles;
  const int32_t* sched = p.sched;

  if (tid == 0) {
    abort_s = 0;
    tile_queue_init(&tq, (int)(blockDim.x / 32));
    for (int i = 0; i < XS; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], CL); }
    ptx::mbar_init(&acc_full, NP);
    ptx::mbar_init(&acc_empty, 1);
    for (int i = 0; i < NP; ++i) ptx::mbar_init(&turn[i], 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tensormap(&maps.x); ptx::prefetch_tensormap(&maps.w); ptx::prefetch_tensormap(&maps.y);
  }
  if (warp == NP) { ptx::tmem_alloc(&tmem_base_s, Cfg::TCOLS); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL == 2) ptx::cluster_sync();          // the peer's barriers are initialised before anything is multicast to them
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp < NP) {
    // ================================ TMA producers ================================
    // Producer `warp` owns the groups whose running index gc (over all tiles of this CTA) has
    // gc % NP == warp; lane i holds int i of the group record.
    uint32_t gbase = 0;                   // groups of earlier tiles
    bool alive = true;
    const uint32_t p_idesc0 = ptx::make_idesc_f16(BF16, p.axis0 != 0, !p.bprop, 128, 0);
    const uint32_t full0 = ptx::opaque(ptx::smem_u32(&full[0])), empty0 = ptx::opaque(ptx::smem_u32(&empty[0])),
                   cmd0 = ptx::opaque(ptx::smem_u32(&cmd[0][0])), stage0 = ptx::opaque(ptx::smem_u32(sStage));
    const uint32_t p_bdesc_lo = (uint32_t)ptx::make_smem_desc(ptx::smem_u32(sStage) + XBYTES, p.bprop ? 16u : WBYTES, Cfg::SBO, Cfg::SWZ);
    const bool fetcher = warp == 0 && lane == 0;
    int drawn = 0;
    if (fetcher) tile_queue_publish(&tq, 0, (int)(blockIdx.x / CL), p.order, total_tiles, abort_flag);
    for (uint32_t tk = 0; alive; ++tk) {
      const int t = tile_queue_next(&tq, tk, lane, abort_flag);
      if (t < 0) break;
      if (fetcher) drawn = CL == 2 ? (int)(blockIdx.x / 2 + (tk + 1) * (gridDim.x / 2))     // clusters use the static deal
                                   : tile_queue_draw(p.counter, tk + 1);   // for tile tk + 1; consumed after this tile's loads are issued
      const int nt = t / kt_per, kt = CL * (t % kt_per) + (int)crank;
      const int32_t* th = sched + 4 + 4 * kt;
      const int first_group = th[0], n_groups = th[1];
      const int32_t* grec = sched + p.groups_off + (size_t)first_group * 32;
      // first group of this tile owned by this warp
      int g = (int)((NP + warp - (gbase % NP)) % NP);
      int rec = (g < n_groups) ? grec[g * 32 + lane] : 0;
      for (; g < n_groups; g += NP) {
        const int cur = rec;
        const int gn = g + NP;
        if (gn < n_groups) rec = grec[gn * 32 + lane];          // prefetch the next record
        // pipeline `warp` (this producer + issuer warp `warp`) owns the stages st % 2 == warp and runs them in order
        const uint32_t gc = gbase + g;
        const uint32_t pj = gc / NP;                  // running group count of this pipeline
        const uint32_t st = (uint32_t)warp + NP * (pj % HS);
        const int in_blk = __shfl_sync(0xffffffffu, cur, 0);
        const int counts = __shfl_sync(0xffffffffu, cur, 1);
        const int n_w = counts & 0xff, n_runs = counts >> 8;
        if (!__all_sync(0xffffffffu, ptx::mbar_wait_a(empty0 + st * 8, ((pj / HS) & 1) ^ 1, abort_flag))) { g_tc_error = 1; alive = false; break; }
        // lanes 12..19 hold int0 of run (lane-12); int1 sits 8 lanes up.  They write the ready-to-issue
        // command so the issuing thread only moves registers.
        const uint32_t r1 = (uint32_t)__shfl_down_sync(0xffffffffu, cur, 8);
        if (lane >= 12 && (lane < 12 + n_runs || lane == 12)) {      // lane 12 always writes: it carries the run count (0 for an empty group)
          const uint32_t r0 = (uint32_t)cur;
          ptx::st_shared_v4(cmd0 + st * (8 * 16) + (lane - 12) * 16,
                            (int)(p_bdesc_lo + ((st * STAGE_BYTES) >> 4) + (r0 & 0xffffu)),
                            (int)(tmem + (r0 >> 16)), (int)(p_idesc0 | (r1 & ~1u)),
                            (int)((r1 & 1u) | (lane == 12 ? (uint32_t)n_runs << 8 : 0u)));
        }
        __syncwarp();
        const uint32_t stage = stage0 + st * STAGE_BYTES, fbar = full0 + st * 8;
        if (lane == 0) {
          ptx::mbar_expect_tx_a(fbar, XBYTES + (uint32_t)n_w * WBYTES);
          if (CL == 2) {
            // both CTAs expect the activation tile; the CTA whose rank equals the group's parity fetches it for both
            if ((g & 1) == (int)crank) {
              if (!p.axis0) {
                ptx::tma_load_2d_mc(stage, &maps.x, fbar, in_blk * BS, nt * 128, (uint16_t)3);
              } else {
                ptx::tma_load_2d_mc(stage, &maps.x, fbar, nt * 128, in_blk * BS, (uint16_t)3);
                ptx::tma_load_2d_mc(stage + XBYTES / 2, &maps.x, fbar, nt * 128 + 64, in_blk * BS, (uint16_t)3);
              }
            }
          } else if (!p.axis0) {
            ptx::tma_load_2d_a(stage, &maps.x, fbar, in_blk * BS, nt * 128);          // [128 n][bs c], K-major A
          } else {                                                                   // [bs c][128 n] as two 64-wide boxes, MN-major A
            ptx::tma_load_2d_a(stage, &maps.x, fbar, nt * 128, in_blk * BS);
            ptx::tma_load_2d_a(stage + XBYTES / 2, &maps.x, fbar, nt * 128 + 64, in_blk * BS);
          }
        }
        if (lane >= 4 && lane < 4 + n_w)
          ptx::tma_load_2d_a(stage + XBYTES + (lane - 4) * WBYTES, &maps.w, fbar, 0, cur * BS);
        __syncwarp();
      }
      gbase += n_groups;
      if (fetcher && alive) tile_queue_publish(&tq, tk + 1, drawn, p.order, total_tiles, abort_flag);
    }
  } else if (warp < 2 * NP) {
    // ================================ MMA issuers ================================
    // Two warps, each with one elected issuing thread, alternate groups (running group index gc % 2): a group
    // costs its issuing warp ~150 instructions (barrier wait, command fetch, descriptor moves to uniform
    // registers) for ~3 tcgen05.mma, and that instruction stream -- not the tensor pipe -- bounded the
    // single-issuer kernel (profiles/r1_xprop_tuning.txt).  Accumulators start from zero (the epilogue clears
    // them), so every MMA accumulates and the order in which the two warps' MMAs reach the pipe is irrelevant.
    // fprop: B = W[c][k] read as K x N with N contiguous (MN-major): K=16 slice = 16 rows, blocks LBO apart.
    // bprop: B = W[c][k] read as N x K with K contiguous (K-major):  K=16 slice = 32 bytes along the row.
    const uint32_t iw = (uint32_t)(warp - NP);
    const uint32_t b_kstep16 = (p.bprop ? 32u : 16u * ROW) >> 4;
    // axis 1: A = X[n][c] tile, K-major, rows of bs*2 bytes, K=16 slice = +32 B.
    // axis 0: A = X[c][n] tile, MN-major SW128: two [bs x 64] boxes (LBO = box), 8-row groups 1 KB apart, K=16 slice = 16 rows.
    const uint64_t a_desc0 = p.axis0 ? ptx::make_smem_desc(ptx::smem_u32(sStage), XBYTES / 2, 1024, ptx::SWZ_128B)
                                     : ptx::make_smem_desc(ptx::smem_u32(sStage), 16, Cfg::SBO, Cfg::SWZ);
    const uint32_t a_kstep16 = p.axis0 ? (16u * 128u) >> 4 : 2u;
    const uint32_t b_desc_hi = (uint32_t)(ptx::make_smem_desc(0, 16, Cfg::SBO, Cfg::SWZ) >> 32);
    const uint32_t full0 = ptx::opaque(ptx::smem_u32(&full[0])), empty0 = ptx::opaque(ptx::smem_u32(&empty[0])),
                   cmd0 = ptx::opaque(ptx::smem_u32(&cmd[0][0]));
    uint32_t tile_it = 0, gbase = 0;
    const uint32_t a_lo0 = (uint32_t)a_desc0, a_hi = (uint32_t)(a_desc0 >> 32);
    bool alive = true;
    for (; alive; ++tile_it) {
      const int t = tile_queue_next(&tq, tile_it, lane, abort_flag);
      if (t < 0) break;
      const int kt = CL * (t % kt_per) + (int)crank;
      const int n_groups = sched[4 + 4 * kt + 1];
      if (!__all_sync(0xffffffffu, ptx::mbar_wait(&acc_empty, tile_it & 1, abort_flag))) { g_tc_error = 3; break; }
      ptx::tc_fence_after();
      int g = (int)((NP + iw - (gbase % NP)) % NP);   // first group of this tile owned by this warp
      uint32_t pj = (gbase + g) / NP;                   // running group count of this pipeline
      uint32_t js = pj % HS, ph = (pj / HS) & 1;                   // slot within this pipeline's stages, phase bit
      for (; g < n_groups; g += NP) {
        const uint32_t st = iw + NP * js;
        if (!__all_sync(0xffffffffu, ptx::mbar_wait_a(full0 + st * 8, ph, abort_flag))) { g_tc_error = 4; alive = false; break; }
        // The two warps take turns in group order (group gc issues after group gc-1), so every accumulator sees
        // its MMAs in schedule order and results are bit-identical run to run.  Only the short issue section is
        // serialised; barrier waits, command fetches and descriptor set-up of the two warps still overlap.
        const bool my_turn = (iw == 0) ? (pj == 0 || ptx::mbar_wait(&turn[0], (pj - 1) & 1, abort_flag))
                                       : ptx::mbar_wait(&turn[iw], pj & 1, abort_flag);
        if (!__all_sync(0xffffffffu, my_turn)) { g_tc_error = 5; alive = false; break; }
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t a_lo = a_lo0 + st * (STAGE_BYTES >> 4);
          // Commands go to registers first (independent LDS.128s), then straight-line issue: run 0 fills the A
          // collector, runs 1.. reuse it.  Most groups have <= 4 runs, so only 4 commands are fetched eagerly;
          // the run count rides in the first command.
          int4 c[8];
          const uint32_t cq = cmd0 + st * (8 * 16);
#pragma unroll
          for (int r = 0; r < 4; ++r)
            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(c[r].x), "=r"(c[r].y), "=r"(c[r].z), "=r"(c[r].w) : "r"(cq + r * 16));
          const int n_runs = c[0].w >> 8;
          if (n_runs > 4) {
#pragma unroll
            for (int r = 4; r < 8; ++r)
              asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(c[r].x), "=r"(c[r].y), "=r"(c[r].z), "=r"(c[r].w) : "r"(cq + r * 16));
          }
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint32_t)(a_lo + ks * a_kstep16);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              if (r >= n_runs) break;              // a real (uniform) branch: skipped runs cost nothing
              const uint64_t bdesc = ((uint64_t)b_desc_hi << 32) | (uint32_t)((uint32_t)c[r].x + ks * b_kstep16);
              if (r == 0) ptx::mma_ss_a_fill((uint32_t)c[r].y, adesc, bdesc, (uint32_t)c[r].z, 1u);
              else        ptx::mma_ss_a_use((uint32_t)c[r].y, adesc, bdesc, (uint32_t)c[r].z, 1u);
            }
          }
          if (CL == 2) ptx::tc_commit_mc(empty0 + st * 8, (uint16_t)3);   // ... in BOTH CTAs: the peer may multicast into this stage
          else ptx::tc_commit_a(empty0 + st * 8);   // the stage is free once these MMAs retire
          ptx::tc_fence_before();
          ptx::mbar_arrive(&turn[iw + 1 == NP ? 0 : iw + 1]);     // hand the turn to the next issuer
        }
        __syncwarp();
        ++pj;
        if (++js == HS) { js = 0; ph ^= 1; }
      }
      if (ptx::elect_one()) ptx::tc_commit(&acc_full);   // arrives when this warp's MMAs of the tile have retired
      __syncwarp();
      gbase += (uint32_t)n_groups;
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;                  // row of the 128-row tile
    const int etid = (warp - 2 * NP) * 32 + lane;
    uint32_t tile_it = 0;
    // accumulators start from zero: clear this warp's lanes once, then after every read-out
#pragma unroll
    for (int c = 0; c < Cfg::TCOLS; c += 32) ptx::tmem_st_zero_x32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c);
    // columns handled per tcgen05.ld: 32, or the whole 16-column block of 16 x 16 blocks
    constexpr int CW = BS < 32 ? BS : 32, NH = BS / CW;
    auto ld_cols = [&](uint32_t col, uint32_t (&v)[32]) {
      if constexpr (CW == 32) ptx::tmem_ld_x32(tmem + ((uint32_t)(quad * 32) << 16) + col, v);
      else { uint32_t t[16]; ptx::tmem_ld_x16(tmem + ((uint32_t)(quad * 32) << 16) + col, t);
#pragma unroll
             for (int i = 0; i < 16; ++i) v[i] = t[i]; }
    };
    auto zero_cols = [&](uint32_t col) {
      if constexpr (CW == 32) ptx::tmem_st_zero_x32(tmem + ((uint32_t)(quad * 32) << 16) + col);
      else ptx::tmem_st_zero_x16(tmem + ((uint32_t)(quad * 32) << 16) + col);
    };
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (etid == 0) ptx::mbar_arrive(&acc_empty);
    for (;; ++tile_it) {
      const int t = tile_queue_next(&tq, tile_it, lane, abort_flag);
      if (t < 0) break;
      const int nt = t / kt_per, kt = CL * (t % kt_per) + (int)crank;
      const int32_t* th = sched + 4 + 4 * kt;
      const int first_out = th[2];
      const int n_out = th[3] & 0xff;
      const uint32_t mask = (uint32_t)th[3] >> 8;
      ptx::mbar_wait(&acc_full, tile_it & 1, abort_flag);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (*abort_flag) { g_tc_error = 6; break; }       // uniform across the 128 epilogue threads
      ptx::tc_fence_after();
      if (p.axis0) {
        // Y is (K, N): lane = minibatch column, register j = output feature -> for every j a warp writes 32
        // consecutive 16-bit values (one 64-byte segment); no staging needed.
        const long long gcol = (long long)nt * 128 + row;
        uint16_t* ycol = reinterpret_cast<uint16_t*>(p.y) + (long long)first_out * BS * p.y_pitch + gcol;
        for (int slot = 0; slot < n_out; ++slot) {
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            uint32_t v[32];
            if ((mask >> slot) & 1u) {
              ld_cols((uint32_t)(slot * BS + h * 32), v);
              ptx::tmem_ld_wait();
              zero_cols((uint32_t)(slot * BS + h * 32));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0u;
            }
            if (gcol < p.N) {
#pragma unroll
              for (int j = 0; j < CW; ++j) {
                uint16_t o;
                if (BF16) { __nv_bfloat16 q = __float2bfloat16_rn(__uint_as_float(v[j])); o = *reinterpret_cast<uint16_t*>(&q); }
                else      { __half q = __float2half_rn(__uint_as_float(v[j]));            o = *reinterpret_cast<uint16_t*>(&q); }
                ycol[(long long)(slot * BS + h * 32 + j) * p.y_pitch] = o;
              }
            }
          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (etid == 0) ptx::mbar_arrive(&acc_empty);
      } else if constexpr (STG == 0) {
        // direct epilogue: thread = one row of the tile; BS 16-bit outputs = one contiguous row segment per block
        const long long grow = (long long)nt * 128 + row;
        uint16_t* yrow = reinterpret_cast<uint16_t*>(p.y) + grow * p.y_pitch + (long long)first_out * BS;
        for (int slot = 0; slot < n_out; ++slot) {
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            uint32_t v[32];
            if ((mask >> slot) & 1u) {
              ld_cols((uint32_t)(slot * BS + h * 32), v);
              ptx::tmem_ld_wait();
              zero_cols((uint32_t)(slot * BS + h * 32));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0u;
            }
            if (grow < p.N) {
#pragma unroll
              for (int c = 0; c < CW / 8; ++c) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float a = __uint_as_float(v[c * 8 + 2 * e]), b = __uint_as_float(v[c * 8 + 2 * e + 1]);
                  if (BF16) { __nv_bfloat162 q = __floats2bfloat162_rn(a, b); pk[e] = *reinterpret_cast<uint32_t*>(&q); }
                  else      { __half2 q = __floats2half2_rn(a, b);           pk[e] = *reinterpret_cast<uint32_t*>(&q); }
                }
                *reinterpret_cast<uint4*>(yrow + slot * BS + h * 32 + c * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (etid == 0) ptx::mbar_arrive(&acc_empty);
      } else {
      for (int s0 = 0; s0 < n_out; s0 += STG) {
        // the staging buffers must have been drained by the TMA stores issued before
        if (etid == 0) ptx::tma_store_wait_read<0>();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int s1 = min(n_out, s0 + STG);
        for (int slot = s0; slot < s1; ++slot) {
          uint8_t* dst = sO + (slot - s0) * XBYTES + row * ROW;
#pragma unroll
          for (int h = 0; h < NH; ++h) {                 // 32 (16) fp32 columns at a time
            uint32_t v[32];
            if ((mask >> slot) & 1u) {
              ld_cols((uint32_t)(slot * BS + h * 32), v);
              ptx::tmem_ld_wait();
              zero_cols((uint32_t)(slot * BS + h * 32));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0u;    // output block with an empty LUT row
            }
#pragma unroll
            for (int c = 0; c < CW / 8; ++c) {           // 16-byte chunks (8 elements each)
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = __uint_as_float(v[c * 8 + 2 * e]), b = __uint_as_float(v[c * 8 + 2 * e + 1]);
                if (BF16) { __nv_bfloat162 q = __floats2bfloat162_rn(a, b); pk[e] = *reinterpret_cast<uint32_t*>(&q); }
                else      { __half2 q = __floats2half2_rn(a, b);           pk[e] = *reinterpret_cast<uint32_t*>(&q); }
              }
              const uint32_t chunk = h * 4 + c;                                  // 16-byte chunk index in the row
              // TMA swizzle of the staging tile: chunk ^= row bits (32B: (row/4)%2, 64B: (row/2)%4, 128B: row%8)
              const uint32_t swz = (BS == 16) ? (chunk ^ ((row >> 2) & 1)) : (BS == 32) ? (chunk ^ ((row >> 1) & 3)) : (chunk ^ (row & 7));
              *reinterpret_cast<uint4*>(dst + swz * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (etid == 0) {
          if (s1 == n_out) ptx::mbar_arrive(&acc_empty);     // accumulators are free for the next tile
          for (int slot = s0; slot < s1; ++slot)
            ptx::tma_store_2d(&maps.y, sO + (slot - s0) * XBYTES, (first_out + slot) * BS, nt * 128);
          ptx::tma_store_commit();
        }
      }
      }
    }
    if (etid == 0) ptx::tma_store_wait<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL == 2) ptx::cluster_sync();          // nobody leaves while the peer may still signal this CTA's barriers
  if (warp == NP) ptx::tmem_dealloc(tmem, Cfg::TCOLS);
  if (tid == 0 && p.counter) tile_queue_retire(p.counter);
}

template <int BS, int OCC, int VAR = 0>
constexpr size_t xprop_smem_bytes() {
  using Cfg = XpropCfg<BS, OCC, VAR>;
  return (size_t)Cfg::XS * (128 * BS * 2 + Cfg::WPS * BS * BS * 2) + (size_t)Cfg::STG * 128 * BS * 2;
}

template <int BS, bool BF16, int OCC, int VAR>
int launch_tc_xprop_pair(XpropTcParams p, const XpropTmaps& maps, int sm_count, cudaStream_t s) {
  auto kern = tc_xprop_kernel<BS, BF16, OCC, VAR, 2>;
  constexpr size_t smem = xprop_smem_bytes<BS, OCC, VAR>();
  static thread_local uint64_t configured = 0;
  if (int e = ensure_dyn_smem(kern, smem, configured)) return e;
  p.counter = nullptr; p.order = nullptr;                 // clusters use the static deal over tile pairs
  const int pairs = (p.n_ktiles / 2) * p.n_ntiles;
  int clusters = sm_count * OCC / 2;
  if (clusters > pairs) clusters = pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(xprop_threads<XpropCfg<BS, OCC, VAR>>()); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p, maps);
  if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "cluster launch: %s", cudaGetErrorString(e)); }
  return check_launch("tcgen05_xprop_bs32_pair");
}

template <int BS, bool BF16, int OCC, int VAR = 0>
int launch_tc_xprop(const XpropTcParams& p, const XpropTmaps& maps, int sm_count, cudaStream_t s) {
  auto kern = tc_xprop_kernel<BS, BF16, OCC, VAR>;
  constexpr size_t smem = xprop_smem_bytes<BS, OCC, VAR>();
  static thread_local uint64_t configured = 0;
  if (int e = ensure_dyn_smem(kern, smem, configured)) return e;
  const int total = p.n_ktiles * p.n_ntiles;
  const int grid = total < sm_count * OCC ? total : sm_count * OCC;
  kern<<<grid, xprop_threads<XpropCfg<BS, OCC, VAR>>(), smem, s>>>(p, maps);
  return check_launch(BS == 16 ? "tcgen05_xprop_bs16" : BS == 32 ? "tcgen05_xprop_bs32" : "tcgen05_xprop_bs64");
}

inline int tc_xprop(int dtype, int axis, int bsize, int bprop, const int32_t* lut, int n_out, int n_in, int blocks,
                    const void* x, const void* w, void* y, int N, const float* gate, const int32_t* sched, int sched_tiles, int sched_tile_blocks,
                    int sched_groups_off, int order_off, int order_ntiles, cudaStream_t s) {
  (void)lut;
  if (dtype != BSMM_F16 && dtype != BSMM_BF16) { fail(0, "fp32 runs on the FMA path"); return TC_NOT_APPLICABLE; }
  if (bsize != 16 && bsize != 32 && bsize != 64) { fail(0, "block size %d uses the CUDA-core path", bsize); return TC_NOT_APPLICABLE; }
  if (gate != nullptr) { fail(0, "gated xprop uses the CUDA-core path"); return TC_NOT_APPLICABLE; }
  if (axis == 0 && (N & 7)) { fail(0, "feature_axis 0 needs N %% 8 == 0 for TMA (row pitch multiple of 16 bytes)"); return TC_NOT_APPLICABLE; }
  if (sched == nullptr || sched_tiles <= 0) { fail(0, "no tile schedule supplied"); return TC_NOT_APPLICABLE; }
  if (order_off > 0 && order_ntiles != (N + 127) / 128)
    return fail(BSMM_E_ARG, "bsmm_xprop: tile order built for %d minibatch tiles, N=%d needs %d", order_ntiles, N, (N + 127) / 128);
  if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) { fail(0, "pointers must be 16-byte aligned for TMA"); return TC_NOT_APPLICABLE; }
  const DeviceInfo& dev = device_info();
  if (!dev.ok || dev.cc_major != 10) { fail(0, "tcgen05 needs an sm_100 device"); return TC_NOT_APPLICABLE; }

  // cuTensorMapEncodeTiled is a driver entry point: make sure this host thread (e.g. an autograd
  // worker) has the primary context bound before calling it
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  const uint64_t Cin = (uint64_t)n_in * bsize, Cout = (uint64_t)n_out * bsize;
  XpropTmaps maps;
  const CUtensorMapSwizzle swz = bsize == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : bsize == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  if (axis == 1) {
    if (int e = cached_tmap_2d(&maps.x, dtype, x, Cin, (uint64_t)N, Cin, bsize, 128, swz)) return e;
  } else {       // (C, N): inner dim = minibatch; box = 64 columns x bs feature rows, 128-byte rows
    if (int e = cached_tmap_2d(&maps.x, dtype, x, (uint64_t)N, Cin, (uint64_t)N, 64, bsize, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  if (int e = cached_tmap_2d(&maps.w, dtype, w, (uint64_t)bsize, (uint64_t)blocks * bsize, (uint64_t)bsize, bsize, bsize, swz)) return e;
  if (int e = cached_tmap_2d(&maps.y, dtype, y, Cout, (uint64_t)N, Cout, bsize, 128, swz)) return e;

  XpropTcParams p;
  p.sched = sched;
  p.n_ntiles = (N + 127) / 128;
  p.bprop = bprop;
  p.axis0 = axis == 0;
  p.y = y; p.y_pitch = axis == 0 ? (long long)N : (long long)Cout; p.N = N;
  // the schedule itself lives in device memory; its shape is passed by value
  p.n_ktiles = sched_tiles;
  p.groups_off = sched_groups_off;
  p.counter = static_tiles() ? nullptr : next_tile_counter();
  if (!p.counter && !static_tiles()) return fail(BSMM_E_NODEV, "bsmm_xprop: tile counters not available");
  p.order = (order_off > 0 && !static_order()) ? sched + order_off : nullptr;
  const int tile_blocks = sched_tile_blocks & 0xff;
  const int w_per_group = (sched_tile_blocks >> 8) & 0xf;  // 0 = the default of the tile width
  const bool pair_tiles = (sched_tile_blocks >> 12) & 1;   // schedule built with pair_tiles=True: cluster kernel
  const int occ = (tile_blocks * bsize <= 256) ? 2 : 1;      // half-width tiles run two CTAs per SM
  if (p.n_ktiles <= 0 || tile_blocks <= 0 || tile_blocks > 512 / bsize || (long long)p.n_ktiles * tile_blocks < n_out ||
      sched_groups_off < 4 + 4 * p.n_ktiles || (sched_groups_off & 31))
    return fail(BSMM_E_ARG, "bsmm_xprop: inconsistent tile schedule (n_tiles=%d, blocks_per_tile=%d, n_out=%d)",
                p.n_ktiles, tile_blocks, n_out);
  if (bsize == 16) {
    if (occ != 2 || (w_per_group != 0 && w_per_group != 8))
      return fail(BSMM_E_ARG, "bsmm_xprop: 16 x 16 blocks need 16-block tiles and 8 W blocks per group");
    return dtype == BSMM_BF16 ? launch_tc_xprop<16, true, 2>(p, maps, dev.sm_grid, s) : launch_tc_xprop<16, false, 2>(p, maps, dev.sm_grid, s);
  }
  if (pair_tiles) {
    if (!(occ == 2 && bsize == 32 && (w_per_group == 4 || w_per_group == 2)) || (p.n_ktiles & 1))
      return fail(BSMM_E_ARG, "bsmm_xprop: pair-tile schedules need 32 x 32 blocks, 8-block tiles, 2 or 4 W blocks per group and an even tile count");
    if (w_per_group == 4) return dtype == BSMM_BF16 ? launch_tc_xprop_pair<32, true, 2, 3>(p, maps, dev.sm_grid, s) : launch_tc_xprop_pair<32, false, 2, 3>(p, maps, dev.sm_grid, s);
    return dtype == BSMM_BF16 ? launch_tc_xprop_pair<32, true, 2, 1>(p, maps, dev.sm_grid, s) : launch_tc_xprop_pair<32, false, 2, 1>(p, maps, dev.sm_grid, s);
  }
  if (occ == 2 && bsize == 32 && w_per_group == 4) {
    return dtype == BSMM_BF16 ? launch_tc_xprop<32, true, 2, 3>(p, maps, dev.sm_grid, s)
                              : launch_tc_xprop<32, false, 2, 3>(p, maps, dev.sm_grid, s);
  }
  if (w_per_group != 0 && !(bsize == 32 && occ == 2 && w_per_group == 2) &&
      w_per_group != (bsize == 32 ? 8 : (occ == 2 ? 2 : 4)))
    return fail(BSMM_E_ARG, "bsmm_xprop: schedule built with %d W blocks per group, no kernel variant matches", w_per_group);
  if (occ == 2 && bsize == 32 && w_per_group == 2)
    return dtype == BSMM_BF16 ? launch_tc_xprop<32, true, 2, 1>(p, maps, dev.sm_grid, s)
                              : launch_tc_xprop<32, false, 2, 1>(p, maps, dev.sm_grid, s);
  if (occ == 2) {
    if (bsize == 32) return dtype == BSMM_BF16 ? launch_tc_xprop<32, true, 2>(p, maps, dev.sm_grid, s)
                                               : launch_tc_xprop<32, false, 2>(p, maps, dev.sm_grid, s);
    return dtype == BSMM_BF16 ? launch_tc_xprop<64, true, 2>(p, maps, dev.sm_grid, s)
                              : launch_tc_xprop<64, false, 2>(p, maps, dev.sm_grid, s);
  }
  if (bsize == 32) return dtype == BSMM_BF16 ? launch_tc_xprop<32, true, 1>(p, maps, dev.sm_grid, s)
                                             : launch_tc_xprop<32, false, 1>(p, maps, dev.sm_grid, s);
  return dtype == BSMM_BF16 ? launch_tc_xprop<64, true, 1>(p, maps, dev.sm_grid, s)
                            : launch_tc_xprop<64, false, 1>(p, maps, dev.sm_grid, s);
}

}  // namespace bsmm

#include "tc_xprop2.cuh"
#include "tc_updat.cuh"
#include "softmax.cuh"
#include "tc_bst.cuh"
