// Host side of the implicit-GEMM core: tensor-map construction, tile bookkeeping, dispatch over the
// (BN, KB, epilogue) instantiations, and the C-ABI ops that are thin descriptions on top of it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/srb.h"
#include "srb_common.h"
#include "srb_convgemm.cuh"

namespace srb {

// ----------------------------------------------------------------------------------------- error plumbing
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return -1;
}
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SRB_PDL");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}
bool l2_persist_reserve() {
  // 0 = not tried, 1 = reserved, -1 = unavailable (per device)
  static int state[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& st = state[dev & 63];
  if (st == 0) {
    int max_persist = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    // the fp32 residual stream of the largest single-GPU bucket (64 x 504 x 256 x 4 B = 33 MB) with headroom
    size_t want = 48u << 20;
    if ((size_t)max_persist < want) want = (size_t)max_persist;
    st = (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) ? 1 : -1;
    cudaGetLastError();
  }
  return st > 0;
}
int num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  return sms[dev];
}

// ----------------------------------------------------------------------------------------- tensor maps
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int kb) {
  return kb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// activation view (batch, rows, channels) bf16 -> 3-D map, box = (kb channels, 128 + tap-span rows, 1)
static int make_act_map(CUtensorMap* m, const void* ptr, int channels, int rows, int batch, long long row_stride,
                        long long batch_stride, int kb, int box_rows) {
  auto enc = get_encode();
  SRB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  SRB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "activation pointer not 16-byte aligned");
  SRB_REQUIRE((row_stride * 2) % 16 == 0 && (batch_stride * 2) % 16 == 0, "activation strides must be multiples of 16 bytes");
  cuuint64_t dims[3] = {(cuuint64_t)channels, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride * 2, (cuuint64_t)batch_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)kb, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kb), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: %d (C=%d rows=%d B=%d rs=%lld bs=%lld kb=%d)",
              (int)r, channels, rows, batch, row_stride, batch_stride, kb);
  return 0;
}

// packed weight [n_total][k_total] bf16 -> 2-D map, box = (kb, bn)
static int make_weight_map(CUtensorMap* m, const void* ptr, int k_total, int n_total, int kb, int bn) {
  auto enc = get_encode();
  SRB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  SRB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "weight pointer not 16-byte aligned");
  SRB_REQUIRE((k_total * 2) % 16 == 0, "weight row pitch must be a multiple of 16 bytes");
  cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)n_total};
  cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
  cuuint32_t box[2] = {(cuuint32_t)kb, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kb), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight) failed: %d (K=%d N=%d kb=%d bn=%d)", (int)r, k_total,
              n_total, kb, bn);
  return 0;
}

// epilogue I/O view (batch, rows, columns) of 2- or 4-byte elements -> 3-D map, box = (32 columns, 32 rows, 1): the block
// one epilogue warp stages.  The swizzle equals the row width (64 B for bf16, 128 B for fp32), which is exactly the
// XOR pattern of stage_slot<4> / stage_slot<8>.
static int make_io_map(CUtensorMap* m, const void* ptr, int elt_bytes, int columns, int rows, int batch, long long row_stride,
                       long long batch_stride, int box_cols = 32) {
  auto enc = get_encode();
  SRB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  SRB_REQUIRE(elt_bytes == 2 || elt_bytes == 4, "epilogue I/O maps are bf16 or fp32");
  SRB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "epilogue tensor not 16-byte aligned");
  SRB_REQUIRE((row_stride * elt_bytes) % 16 == 0 && (batch_stride * elt_bytes) % 16 == 0 && columns % box_cols == 0,
              "epilogue tensor strides must be multiples of 16 bytes and its width a multiple of the block width");
  const int row_bytes = box_cols * elt_bytes;
  SRB_REQUIRE(row_bytes == 32 || row_bytes == 64 || row_bytes == 128, "epilogue block rows are 32, 64 or 128 bytes");
  cuuint64_t dims[3] = {(cuuint64_t)columns, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride * elt_bytes, (cuuint64_t)(batch > 1 ? batch_stride : (long long)rows * row_stride) * elt_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(m, elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(epilogue I/O) failed: %d (cols=%d rows=%d B=%d rs=%lld bs=%lld)", (int)r,
              columns, rows, batch, row_stride, batch_stride);
  return 0;
}

// ----------------------------------------------------------------------------------------- generic description
struct ActView {
  const void* ptr = nullptr;
  int channels = 0, rows = 0, batch = 0;
  long long row_stride = 0, batch_stride = 0;
};

struct ConvGemmDesc {
  ActView src[kMaxSrc];
  int n_src = 1;
  const void* weight = nullptr;
  int n_total = 0;
  int block_n = 0, block_k = 0;
  int channels = 0;                         // input channels per tap (same for every source)
  int n_groups = 1;
  int group_tap_begin[kMaxGroups + 1] = {0};
  int group_rows[kMaxGroups] = {0};
  int group_row_add[kMaxGroups] = {0};
  int row_mul = 1;
  int tap_src[kMaxTaps] = {0};
  int tap_shift[kMaxTaps] = {0};
  int batch = 0;
  int epilogue = EPI_GENERIC;
  ConvGemmParams epi;                       // only the epilogue-operand fields are read from here
};

// cluster mode of the wide (BN = 256) tiles: 0 = single CTA, 1 = weight-slab multicast, 2 = CTA-pair MMA
// (cta_group::2).  SRB_CLUSTER_MODE overrides the default.
// Default 1 since round 2: three alternating same-box runs at config 2 (profiles/r02_cluster_mode_ab.txt) gave 28.68 /
// 28.65 / 28.58 ms per step with multicast against 28.68 / 29.14 / 28.79 without, and 29.02 +- 0.04 against 29.31 +- 0.02
// end to end -- 1 %, from the C = 256 vocoder convs and the FFN GEMM; the per-op table shows the K = 256 launches
// (q|k projection: 24.4 vs 22.6 us) paying for their cluster start-up, which is why mode 2 (pair MMA: 29.5 ms) loses.
// Per launch kind since the end of round 2: the same table (profiles/r02_ops_cluster_mode*.csv) and an isolated, L2-flushed
// timing of the FFN conv2 launch (52.2 us single CTA, 55.5 us multicast) show that only the GLU GEMM and the C = 256
// vocoder convs gain from multicast; the RESNORM launches lose 2-3 us each and the q|k projection loses its
// weight-stationary form (24.4 vs 22.6 us).  So: multicast for GLU / GENERIC, single CTA for RESNORM / QKV_ROPE.
// An explicit SRB_CLUSTER_MODE applies to every eligible launch as before.
static int cluster_mode(int epilogue) {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("SRB_CLUSTER_MODE");
    v = e ? atoi(e) : -1;
    if (v < -1 || v > 2) v = 0;
  }
  if (v >= 0) return v;
  return (epilogue == EPI_GLU || epilogue == EPI_GENERIC) ? 1 : 0;
}

// SRB_WEIGHT_STATIONARY is a bit mask for A/B measurements: 1 = transformer linears, 2 = vocoder convs (default 3)
static int weight_stationary_mask() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SRB_WEIGHT_STATIONARY");
    v = e ? (atoi(e) & 3) : 3;
  }
  return v;
}

template <int BN, int KB, int EPI, int MC = 0, int WS = 0>
static int launch_inst(ConvGemmParams& p, int total_tiles, cudaStream_t stream, int n_slabs = 0) {
  using L = StageLayout<BN, KB>;
  auto kernel = convgemm_kernel<BN, KB, EPI, MC, WS>;
  // two rings: activation boxes (one per K chunk, reused by all taps) and weight slabs (one per tap and K chunk)
  constexpr int threads = 64 + 32 * EpiWarps<BN, EPI>::value;
  const int a_bytes = p.a_box_bytes;
  int a_stages, w_stages;
  if (L::w_bytes >= 32 * 1024) { a_stages = 3; w_stages = 4; }        // wide tiles: one CTA per SM
  else if (L::w_bytes >= 8 * 1024) { a_stages = 2; w_stages = 4; }    // two CTAs per SM
  if (EPI == EPI_GENERIC && BN == 128) {
    // A/B knob: ring depth of the C = 128 vocoder convs ("a,w"), e.g. 2,3 so that two different launches fit one SM
    static const char* e = getenv("SRB_GEN128_STAGES");
    if (e) { a_stages = atoi(e); const char* c = strchr(e, ','); if (c) w_stages = atoi(c + 1); }
  }
  else { a_stages = 2; w_stages = 6; }
  {
    // experiment knobs (profiling only)
    const char* ea = getenv("SRB_A_STAGES");
    const char* ew = getenv("SRB_W_STAGES");
    if (ea && L::w_bytes >= 32 * 1024) a_stages = atoi(ea);
    if (ew && L::w_bytes >= 32 * 1024) w_stages = atoi(ew);
  }
  constexpr int w_stage_bytes = MC == 2 ? L::w_bytes / 2 : L::w_bytes;
  // RESNORM: two residual-stream buffers per epilogue warp when the K loop is short (the epilogue, not the MMAs,
  // paces those launches); one when the long K loop (FFN conv2) needs the shared memory for its weight ring
  if (EPI != EPI_GENERIC) p.res_bufs = (EPI == EPI_RESNORM && n_slabs <= 8) ? 2 : 1;   // (GENERIC: residual count, set by the caller)
  {
    // SRB_WEIGHTS_EARLY=0 requests the first weight slabs after the dependency wait instead of before it (A/B runs)
    static const int early = [] { const char* e = getenv("SRB_WEIGHTS_EARLY"); return e ? atoi(e) : 1; }();
    p.w_early = (early && !p.w_dynamic) ? 1 : 0;
  }
  // coalescing / streaming buffers of the epilogue warps + CTA-wide epilogue tables
  const int stage_smem = EpiWarps<BN, EPI>::value * EpiWarps<BN, EPI>::stage_bytes(p.res_bufs) + EpiWarps<BN, EPI>::extra_bytes;
  if (MC == 2) w_stages = 6;
  if (!WS && n_slabs > 0 && w_stages > n_slabs) w_stages = n_slabs < 2 ? 2 : n_slabs;
  if (WS) {
    // every slab resident; the activation ring takes what is left (at most 4 boxes)
    w_stages = n_slabs;
    a_stages = L::w_bytes >= 32 * 1024 ? 4 : 3;
    if (EPI == EPI_GENERIC) {
      static const int ws_a = [] { const char* e = getenv("SRB_WS_A_STAGES"); return e ? atoi(e) : 0; }();   // A/B knob
      if (ws_a >= 2) a_stages = ws_a;
    }
    while (a_stages > 2 && a_stages * a_bytes + w_stages * w_stage_bytes + stage_smem > 220 * 1024) --a_stages;
    // narrow tiles: prefer two resident CTAs (their epilogues overlap) over a third activation box
    if (a_stages == 3 && 3 * a_bytes + w_stages * w_stage_bytes + stage_smem > 108 * 1024 &&
        2 * a_bytes + w_stages * w_stage_bytes + stage_smem <= 108 * 1024)
      a_stages = 2;
    SRB_REQUIRE(a_stages * a_bytes + w_stages * w_stage_bytes + stage_smem <= 222 * 1024, "weight-stationary slabs do not fit");
  }
  while (a_stages * a_bytes + w_stages * w_stage_bytes > 208 * 1024 && w_stages > 2 && !WS) --w_stages;
  p.a_stages = a_stages;
  p.w_stages = w_stages;
  while (a_stages * a_bytes + w_stages * w_stage_bytes + stage_smem > 222 * 1024 && w_stages > 2 && !WS) --w_stages;
  p.w_stages = w_stages;
  // alignment slack + rings + 2 KB bookkeeping block (barriers, TMEM slot, norm exchange) + epilogue staging / tables
  SRB_REQUIRE(8 * (2 * (a_stages + w_stages) + 4) + 16 + 1024 + 384 <= 2048, "too many pipeline stages for the bookkeeping block");
  {
    // RESNORM with ONE residual buffer per warp (the long-K launches: FFN conv2): the last tile of a CTA takes its remaining
    // residual chunks at once into the idle weight ring (3 x 4 KB per epilogue warp) instead of paying one L2 latency per
    // chunk in its exposed epilogue.  Same-box A/B at config 2: ffn_out_norm 51.8 / 51.9 -> 49.7 / 49.6 us per launch
    // (857 -> 894 TFLOP/s); the two-buffer launches (attn_out_norm) gained nothing (28.5 / 28.9 vs 28.8 / 28.8) and keep
    // their stream.  SRB_RES_TAIL=0 switches it off, =2 also enables it for the two-buffer launches (A/B knob).
    static const int tail_on = [] { const char* e = getenv("SRB_RES_TAIL"); return e ? atoi(e) : 1; }();
    p.res_tail = (EPI == EPI_RESNORM && MC != 2 && tail_on && (p.res_bufs == 1 || tail_on == 2) && !p.w_dynamic &&
                  (long long)w_stages * w_stage_bytes >= (long long)EpiWarps<BN, EPI>::value * 3 * 4096) ? 1 : 0;
  }
  const int smem = 1024 + a_stages * a_bytes + w_stages * w_stage_bytes + 2048 + stage_smem;
  static int configured_smem[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_smem[dev & 63] < smem) {
    SRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem[dev & 63] = smem;
  }
  // resident CTAs per SM: shared memory, registers, threads and -- unknown to the occupancy API -- the 512 TMEM
  // columns every CTA carves its accumulators from
  static int num_regs[64] = {0};
  if (num_regs[dev & 63] == 0) {
    cudaFuncAttributes fa;
    SRB_CUDA(cudaFuncGetAttributes(&fa, kernel));
    num_regs[dev & 63] = fa.numRegs;
  }
  int occ = 232448 / (smem + 1024);
  const int by_regs = 65536 / (((num_regs[dev & 63] + 7) & ~7) * threads);
  const int by_threads = 2048 / threads;
  const int by_tmem = 512 / TmemCols<BN>::total;
  occ = occ < by_regs ? occ : by_regs;
  occ = occ < by_threads ? occ : by_threads;
  occ = occ < by_tmem ? occ : by_tmem;
  occ = occ < 1 ? 1 : occ;
  if (WS && EPI == EPI_GENERIC) {
    // A/B knob: cap the resident CTAs per SM of the weight-stationary vocoder convs (co-residency with the kernels of the
    // other resblock chains running as parallel graph branches)
    static const int cap = [] { const char* e = getenv("SRB_WS_OCC"); return e ? atoi(e) : 0; }();
    if (cap > 0 && occ > cap) occ = cap;
  }
  int grid = num_sms() * occ;
  if constexpr (MC) {
    // 2-CTA clusters; `total_tiles` counts PAIRS of row tiles here
    int clusters = grid / 2;
    if (clusters > total_tiles) clusters = total_tiles;
    if (clusters < 1) return 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    SRB_CUDA(cudaLaunchKernelEx(&cfg, kernel, p, total_tiles));
    return after_launch("convgemm_kernel(multicast)");
  }
  if constexpr (WS) {
    // CTA c keeps channel tile c % n_tiles: the grid is a multiple of n_tiles, no more row walkers than row tiles
    int walkers = grid / p.n_tiles;
    if (walkers > total_tiles) walkers = total_tiles;
    if (walkers < 1) walkers = 1;
    grid = walkers * p.n_tiles;
    SRB_CUDA(launch_pdl(kernel, dim3(grid), dim3(threads), smem, stream, p, total_tiles));
    return after_launch("convgemm_kernel(weight-stationary)");
  }
  if (grid > total_tiles) grid = total_tiles;
  if (grid < 1) return 0;
  L2Window win;
  if (EPI == EPI_RESNORM && p.l2_keep == 2) {
    // the residual stream is updated in place: keep it in the persisting part of L2 between its two uses per layer
    win.base = p.out1;
    win.bytes = (size_t)p.batch * (size_t)p.out_batch_stride * 4;
  }
  SRB_CUDA(launch_pdl_window(kernel, dim3(grid), dim3(threads), smem, stream, win, p, total_tiles));
  return after_launch("convgemm_kernel");
}

#ifdef SRB_TRACE
static unsigned long long* g_trace = nullptr;
extern "C" void srb_debug_set_trace(unsigned long long* buf) { g_trace = buf; }
unsigned long long* debug_trace_buffer() { return g_trace; }
#endif

static int launch_convgemm(const ConvGemmDesc& d, cudaStream_t stream) {
  ConvGemmParams p = d.epi;
#ifdef SRB_TRACE
  p.trace = g_trace;
#endif
  const int bn = d.block_n, kb = d.block_k;
  SRB_REQUIRE(d.n_total % bn == 0, "n_total %d not a multiple of block_n %d", d.n_total, bn);
  SRB_REQUIRE(d.n_groups >= 1 && d.n_groups <= kMaxGroups, "bad group count %d", d.n_groups);
  const int n_taps = d.group_tap_begin[d.n_groups];
  SRB_REQUIRE(n_taps >= 1 && n_taps <= kMaxTaps, "bad tap count %d", n_taps);
  p.batch = d.batch;
  p.kchunks = (d.channels * kSplit + kb - 1) / kb;
  p.n_groups = d.n_groups;
  p.n_tiles = d.n_total / bn;
  p.row_mul = d.row_mul;
  int tiles = 0;
  for (int g = 0; g < kMaxGroups; ++g) {
    p.tile_begin[g] = tiles;
    if (g < d.n_groups) {
      p.m_tiles[g] = (d.group_rows[g] + kTileM - 1) / kTileM;
      p.group_rows[g] = d.group_rows[g];
      p.group_row_add[g] = d.group_row_add[g];
      tiles += d.batch * p.m_tiles[g] * p.n_tiles;
    } else {
      p.m_tiles[g] = 1;
      p.group_rows[g] = 0;
      p.group_row_add[g] = 0;
    }
    p.group_tap_begin[g] = d.group_tap_begin[g < d.n_groups ? g : d.n_groups];
  }
  p.tile_begin[kMaxGroups] = tiles;
  p.group_tap_begin[kMaxGroups] = n_taps;
  for (int g = d.n_groups; g <= kMaxGroups; ++g) p.group_tap_begin[g] = n_taps;
  for (int t = 0; t < kMaxTaps; ++t) {
    p.tap_shift[t] = (short)(t < n_taps ? d.tap_shift[t] : 0);
    p.tap_src[t] = (signed char)(t < n_taps ? d.tap_src[t] : 0);
  }
  const int k_total = n_taps * p.kchunks * kb;
  // segments: maximal runs of taps with the same source inside a group; one A box (with halo) per segment and chunk
  int n_seg = 0, span = 0;
  for (int g = 0; g < d.n_groups; ++g) {
    p.group_seg_begin[g] = n_seg;
    int t = d.group_tap_begin[g];
    const int te = d.group_tap_begin[g + 1];
    while (t < te) {
      SRB_REQUIRE(n_seg < kMaxSegs, "too many tap segments");
      int u = t, lo = d.tap_shift[t], hi = d.tap_shift[t];
      while (u < te && d.tap_src[u] == d.tap_src[t]) {
        lo = d.tap_shift[u] < lo ? d.tap_shift[u] : lo;
        hi = d.tap_shift[u] > hi ? d.tap_shift[u] : hi;
        ++u;
      }
      p.seg_tap_begin[n_seg] = (short)t;
      p.seg_min_shift[n_seg] = (short)lo;
      p.seg_src[n_seg] = (signed char)d.tap_src[t];
      span = (hi - lo) > span ? (hi - lo) : span;
      ++n_seg;
      t = u;
    }
  }
  for (int g = d.n_groups; g <= kMaxGroups; ++g) p.group_seg_begin[g] = n_seg;
  for (int sgi = n_seg; sgi <= kMaxSegs; ++sgi) p.seg_tap_begin[sgi] = (short)n_taps;
  p.seg_tap_begin[n_seg] = (short)n_taps;
  SRB_REQUIRE(kTileM + span <= 256, "tap span %d too large for one TMA box", span);
  p.a_box_rows = kTileM + span;
  p.a_box_bytes = (p.a_box_rows * kb * 2 + 1023) & ~1023;
  for (int s = 0; s < kMaxSrc; ++s) {
    const ActView& a = d.src[s < d.n_src ? s : 0];
    int rc = make_act_map(&p.tmA[s], a.ptr, a.channels, a.rows, a.batch, a.row_stride, a.batch_stride, kb, p.a_box_rows);
    if (rc) return rc;
  }
  int rc = make_weight_map(&p.tmW, d.weight, k_total, d.n_total, kb, bn);
  if (rc) return rc;
  if (tiles == 0) return 0;
  if (d.epilogue == EPI_RESNORM || d.epilogue == EPI_QKV_ROPE) {
    // epilogue I/O through TMA (see epi_resnorm): all of these tensors are (batch, rows, row_stride columns)
    const int rows = d.group_rows[0];
    if (d.epilogue == EPI_RESNORM) {
      SRB_REQUIRE(p.res[0] != nullptr && p.out1 != nullptr, "RESNORM needs a residual and an fp32 output");
      rc = make_io_map(&p.tmR, p.res[0], 4, (int)p.res_row_stride, rows, d.batch, p.res_row_stride, p.res_batch_stride);
      if (rc) return rc;
      rc = make_io_map(&p.tmO1, p.out1, 4, (int)p.out_row_stride, rows, d.batch, p.out_row_stride, p.out_batch_stride);
      if (rc) return rc;
    }
    if (p.out0 != nullptr) {
      rc = make_io_map(&p.tmO0, p.out0, 2, (int)p.out_row_stride * kSplit, rows, d.batch, p.out_row_stride * kSplit,
                       p.out_batch_stride * kSplit);
      if (rc) return rc;
    }
  }
  if (d.epilogue == EPI_GENERIC && kSplit != 1) {
    // split build: always the register / LSU epilogue (it writes the three column blocks and reads residuals as hi + lo)
    p.n_res = 0;
    p.gen_lsu = 1;
    p.gen_nbuf = 1;
    p.res_bufs = -1;
  } else if (d.epilogue == EPI_GENERIC) {
    // residuals first (packed), then per tap group the raw / activated output views (row = q * row_mul + row_add)
    const int cw = bn < 32 ? bn : 32;
    int n_res = 0;
    for (int r = 0; r < 3; ++r) {
      if (p.res[r] == nullptr) continue;
      SRB_REQUIRE(d.row_mul == 1, "residuals are not supported with strided output rows");
      rc = make_io_map(&p.tmRes[n_res], p.res[r], 2, d.n_total, d.group_rows[0], d.batch, p.res_row_stride, p.res_batch_stride, cw);
      if (rc) return rc;
      ++n_res;
    }
    p.n_res = n_res;
    // Epilogue form.  Launches whose MMA phase dominates (wide tiles with many taps, the fused tails) keep the register /
    // LSU epilogue: it hides under the MMAs and its 2 KB of staging leaves the shared memory to the weight ring.  The
    // memory-bound ones (k = 3 convs, the C = 64 stage, the transposed convs, v^T) go through TMA, double buffered
    // where the tile is wide enough to afford it.  SRB_GENERIC_EPILOGUE = lsu | tma forces one form (A/B knob).
    const char* force = getenv("SRB_GENERIC_EPILOGUE");
    const int work = (d.group_tap_begin[1] - d.group_tap_begin[0]) * p.kchunks;   // K steps of a tile (first group)
    p.gen_lsu = (n_res >= 2 || (bn >= 128 && work > 16 && d.row_mul == 1)) ? 1 : 0;
    if (force && n_res < 2) p.gen_lsu = strcmp(force, "lsu") == 0;
    // (two staging blocks for the BN = 64 conv2 launches were measured at the end of round 2 -- same-box A/B: k = 11 272.9 ->
    // 268.8 us, k = 3 176.8 -> 174.7, step unchanged -- and not kept: their extra time over conv1 is not residual latency)
    p.gen_nbuf = bn >= 128 ? 2 : 1;
    p.res_bufs = p.gen_lsu ? -1 : p.gen_nbuf * (n_res + 1);
    void* outs[2] = {p.out1, p.out0};
    for (int k = 0; k < 2; ++k) {
      if (outs[k] == nullptr) continue;
      for (int g = 0; g < d.n_groups; ++g) {
        if (d.group_rows[g] <= 0) continue;
        const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(outs[k]) + (long long)d.group_row_add[g] * p.out_row_stride;
        rc = make_io_map(&p.tmOut[k][g], base, 2, d.n_total, d.group_rows[g], d.batch, p.out_row_stride * d.row_mul,
                         p.out_batch_stride, cw);
        if (rc) return rc;
      }
    }
  }
  // weight multicast across 2-CTA clusters for the wide tiles (needs an even number of row tiles)
  const int row_tiles = d.batch * p.m_tiles[0];
  const int mode = (bn == 256 && kb == 64 && d.n_groups == 1 && row_tiles % 2 == 0 && row_tiles >= 2 && d.epilogue != EPI_ARGMAX)
                       ? cluster_mode(d.epilogue) : 0;
  if (mode != 0) {
    rc = make_weight_map(&p.tmWh, d.weight, k_total, d.n_total, kb, bn / 2);
    if (rc) return rc;
    const int pair_tiles = (row_tiles / 2) * p.n_tiles;
    if (mode == 1) {
      if (d.epilogue == EPI_GENERIC) return launch_inst<256, 64, EPI_GENERIC, 1>(p, pair_tiles, stream, n_taps * p.kchunks);
      if (d.epilogue == EPI_GLU) return launch_inst<256, 64, EPI_GLU, 1>(p, pair_tiles, stream, n_taps * p.kchunks);
      if (d.epilogue == EPI_RESNORM) return launch_inst<256, 64, EPI_RESNORM, 1>(p, pair_tiles, stream, n_taps * p.kchunks);
      if (d.epilogue == EPI_QKV_ROPE) return launch_inst<256, 64, EPI_QKV_ROPE, 1>(p, pair_tiles, stream, n_taps * p.kchunks);
    } else {
      if (d.epilogue == EPI_GENERIC) return launch_inst<256, 64, EPI_GENERIC, 2>(p, pair_tiles, stream, n_taps * p.kchunks);
      if (d.epilogue == EPI_GLU) return launch_inst<256, 64, EPI_GLU, 2>(p, pair_tiles, stream, n_taps * p.kchunks);
      if (d.epilogue == EPI_RESNORM) return launch_inst<256, 64, EPI_RESNORM, 2>(p, pair_tiles, stream, n_taps * p.kchunks);
      if (d.epilogue == EPI_QKV_ROPE) return launch_inst<256, 64, EPI_QKV_ROPE, 2>(p, pair_tiles, stream, n_taps * p.kchunks);
    }
  } else {
    p.tmWh = p.tmW;
  }

  // weight-stationary form when all slabs of a channel tile fit beside the activation ring (see convgemm_kernel)
  {
    const int n_slabs = n_taps * p.kchunks;
    const long long w_res = (long long)n_slabs * ((bn * kb * 2 + 1023) & ~1023);
    const int ws_bit = d.epilogue == EPI_GENERIC ? 2 : 1;
    const bool fits = d.n_groups == 1 && w_res + 2 * p.a_box_bytes <= 184 * 1024 && (weight_stationary_mask() & ws_bit);
    if (fits) {
#define SRB_DISPATCH_WS(BN_, KB_, EPI_) \
  if (bn == BN_ && kb == KB_ && d.epilogue == EPI_) return launch_inst<BN_, KB_, EPI_, 0, 1>(p, row_tiles, stream, n_slabs);
      SRB_DISPATCH_WS(256, 64, EPI_QKV_ROPE)
      SRB_DISPATCH_WS(64, 64, EPI_GENERIC)
      SRB_DISPATCH_WS(32, 32, EPI_GENERIC)    // the C = 32 -> 2 x 16 row-group up-sampler: 6 KB of weights per 8 KB activation box
#undef SRB_DISPATCH_WS
    }
  }

#define SRB_DISPATCH(BN_, KB_, EPI_) \
  if (bn == BN_ && kb == KB_ && d.epilogue == EPI_) return launch_inst<BN_, KB_, EPI_>(p, tiles, stream, n_taps * p.kchunks);
  SRB_DISPATCH(256, 64, EPI_GENERIC)
  SRB_DISPATCH(128, 64, EPI_GENERIC)
  SRB_DISPATCH(64, 64, EPI_GENERIC)
  SRB_DISPATCH(32, 64, EPI_GENERIC)
  SRB_DISPATCH(32, 32, EPI_GENERIC)
  SRB_DISPATCH(16, 32, EPI_GENERIC)
  SRB_DISPATCH(16, 16, EPI_GENERIC)
  SRB_DISPATCH(256, 64, EPI_GLU)
  SRB_DISPATCH(256, 64, EPI_RESNORM)
  SRB_DISPATCH(256, 64, EPI_QKV_ROPE)
  SRB_DISPATCH(80, 64, EPI_EULER)
  SRB_DISPATCH(256, 64, EPI_ARGMAX)
#undef SRB_DISPATCH
  set_error("no convgemm instantiation for block_n=%d block_k=%d epilogue=%d", bn, kb, d.epilogue);
  return -4;
}

// residual stream residency (A/B knob SRB_L2_KEEP): 0 = none, 1 = evict_last policy on its TMA transfers (default),
// 2 = persisting L2 window as a launch attribute (costs the vocoder its share of L2)
static int l2_keep_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SRB_L2_KEEP");
    v = e ? atoi(e) : 1;
  }
  return v;
}

static ConvGemmParams empty_epi() {
  ConvGemmParams e;
  memset(&e, 0, sizeof(e));
  e.scale = 1.f;
  e.slope = 1.f;
  e.res_unact = 1.f;
  return e;
}

static ActView act(const void* ptr, int batch, int rows, int channels) {
  ActView a;
  // (split build: a bf16 activation tensor of logical width `channels` is stored as [hi | lo | hi], kSplit * channels wide)
  a.ptr = ptr;
  a.channels = channels * kSplit;
  a.rows = rows;
  a.batch = batch;
  a.row_stride = channels * kSplit;
  a.batch_stride = (long long)rows * channels * kSplit;
  return a;
}

static void same_conv_taps(ConvGemmDesc& d, int src, int kernel, int dilation, int& ntap) {
  for (int j = 0; j < kernel; ++j) {
    d.tap_src[ntap] = src;
    d.tap_shift[ntap] = (j - (kernel - 1) / 2) * dilation;
    ++ntap;
  }
}

// channel block the vocoder uses for a given input width
static int block_k_for(int c_in) { return c_in >= 64 ? 64 : (c_in >= 32 ? 32 : 16); }

}  // namespace srb

using namespace srb;

extern "C" {

int srb_version(void) { return SRB_VERSION; }
int srb_split_factor(void) { return kSplit; }
const char* srb_last_error(void) { return g_err; }

int srb_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  SRB_CUDA(cudaGetDevice(&dev));
  SRB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SRB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}

int srb_cfm_embed(const void* xt_bf16, const void* w_packed, const float* cond_proj, float* x0, int32_t batch,
                  int32_t frames, void* stream) {
  ConvGemmDesc d;
  d.src[0] = act(xt_bf16, batch, frames, 80);
  d.weight = w_packed;
  d.n_total = 256;
  d.block_n = 256;
  d.block_k = 64;
  d.channels = 80;
  d.group_tap_begin[1] = 1;
  d.group_rows[0] = frames;
  d.batch = batch;
  d.epilogue = EPI_RESNORM;
  d.epi = empty_epi();
  d.epi.norm_mode = 0;
  d.epi.res[0] = cond_proj;
  d.epi.res_row_stride = 256;
  d.epi.res_batch_stride = (long long)frames * 256;
  d.epi.out1 = x0;
  d.epi.out_row_stride = 256;
  d.epi.out_batch_stride = (long long)frames * 256;
  return launch_convgemm(d, (cudaStream_t)stream);
}

int srb_cfm_qkv_rope(const void* xn_bf16, const void* w_packed, const float* rot_cos, const float* rot_sin,
                     void* qkv_bf16, float* qk_norm2_max, float* qk_norm2_clear, int32_t batch, int32_t frames,
                     void* stream) {
  SRB_REQUIRE(qk_norm2_clear == nullptr || qk_norm2_clear != qk_norm2_max, "srb_cfm_qkv_rope: the buffer to clear must differ from the one to fill");
  ConvGemmDesc d;
  d.src[0] = act(xn_bf16, batch, frames, 256);
  d.weight = w_packed;
  d.n_total = 768;
  d.block_n = 256;
  d.block_k = 64;
  d.channels = 256;
  d.group_tap_begin[1] = 1;
  d.group_rows[0] = frames;
  d.batch = batch;
  d.epilogue = EPI_QKV_ROPE;
  d.epi = empty_epi();
  d.epi.vec0 = rot_cos;
  d.epi.vec1 = rot_sin;
  d.epi.out0 = qkv_bf16;
  d.epi.aux0 = qk_norm2_max;
  d.epi.aux1 = qk_norm2_clear;
  d.epi.out_row_stride = 768;
  d.epi.out_batch_stride = (long long)frames * 768;
  return launch_convgemm(d, (cudaStream_t)stream);
}

int srb_cfm_qk_rope_vt(const void* xn_bf16, const void* w_packed, const float* rot_cos, const float* rot_sin, void* qk_bf16,
                       void* vt_bf16, int64_t m_pad, float* qk_norm2_max, float* qk_norm2_clear, int32_t batch,
                       int32_t frames, void* stream) {
  SRB_REQUIRE(kSplit == 1, "srb_cfm_qk_rope_vt: not available in the tight-precision build");
  SRB_REQUIRE(qk_norm2_clear == nullptr || qk_norm2_clear != qk_norm2_max, "srb_cfm_qk_rope_vt: the buffer to clear must differ from the one to fill");
  SRB_REQUIRE(frames % 8 == 0 && m_pad % 8 == 0 && m_pad >= (int64_t)batch * frames,
              "srb_cfm_qk_rope_vt: frames and m_pad must be multiples of 8 and m_pad >= batch * frames");
  SRB_REQUIRE((reinterpret_cast<uintptr_t>(vt_bf16) & 15) == 0, "srb_cfm_qk_rope_vt: v^T pointer not 16-byte aligned");
  // the whole to_qkv GEMM (N = 768) in one launch: tiles 0 / 1 = q / k with rotary into the (B, N, 512) buffer, tile 2 = v
  // stored transposed
  ConvGemmDesc d;
  d.src[0] = act(xn_bf16, batch, frames, 256);
  d.weight = w_packed;
  d.n_total = 768;
  d.block_n = 256;
  d.block_k = 64;
  d.channels = 256;
  d.group_tap_begin[1] = 1;
  d.group_rows[0] = frames;
  d.batch = batch;
  d.epilogue = EPI_QKV_ROPE;
  d.epi = empty_epi();
  d.epi.vec0 = rot_cos;
  d.epi.vec1 = rot_sin;
  d.epi.out0 = qk_bf16;
  d.epi.aux0 = qk_norm2_max;
  d.epi.aux1 = qk_norm2_clear;
  d.epi.out_row_stride = 512;
  d.epi.out_batch_stride = (long long)frames * 512;
  d.epi.vt_out = vt_bf16;
  {
    auto enc = get_encode();
    SRB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)frames, (cuuint64_t)batch, 256};
    cuuint64_t strides[2] = {(cuuint64_t)frames * 2, (cuuint64_t)m_pad * 2};
    cuuint32_t box[3] = {32, 1, 32};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&d.epi.tmVt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, vt_bf16, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(v^T store) failed: %d (frames=%d batch=%d m_pad=%lld)", (int)r, frames,
                batch, (long long)m_pad);
  }
  return launch_convgemm(d, (cudaStream_t)stream);
}

int srb_cfm_qk_rope(const void* xn_bf16, const void* w_packed, const float* rot_cos, const float* rot_sin,
                    void* qk_bf16, float* qk_norm2_max, float* qk_norm2_clear, int32_t batch, int32_t frames, void* stream) {
  SRB_REQUIRE(kSplit == 1, "srb_cfm_qk_rope: not available in the tight-precision build");
  SRB_REQUIRE(qk_norm2_clear == nullptr || qk_norm2_clear != qk_norm2_max, "srb_cfm_qk_rope: the buffer to clear must differ from the one to fill");
  // q and k only (first 512 rows of to_qkv.weight); v is produced transposed by srb_cfm_v_transposed
  ConvGemmDesc d;
  d.src[0] = act(xn_bf16, batch, frames, 256);
  d.weight = w_packed;
  d.n_total = 512;
  d.block_n = 256;
  d.block_k = 64;
  d.channels = 256;
  d.group_tap_begin[1] = 1;
  d.group_rows[0] = frames;
  d.batch = batch;
  d.epilogue = EPI_QKV_ROPE;
  d.epi = empty_epi();
  d.epi.vec0 = rot_cos;
  d.epi.vec1 = rot_sin;
  d.epi.out0 = qk_bf16;
  d.epi.aux0 = qk_norm2_max;
  d.epi.aux1 = qk_norm2_clear;
  d.epi.out_row_stride = 512;
  d.epi.out_batch_stride = (long long)frames * 512;
  return launch_convgemm(d, (cudaStream_t)stream);
}

int srb_cfm_v_transposed(const void* xn_bf16, const void* wv_bf16, void* vt_bf16, int64_t m_pad, void* stream) {
  SRB_REQUIRE(kSplit == 1, "srb_cfm_v_transposed: not available in the tight-precision build");
  // V^T[d][row] = sum_c Wv[d][c] * xn[row][c]: the same GEMM core with the operand roles swapped (the weight
  // matrix is the "activation" tile, the activations are the K-major B operand), so V comes out keys-contiguous.
  SRB_REQUIRE(m_pad > 0 && m_pad % 256 == 0, "srb_cfm_v_transposed: m_pad must be a positive multiple of 256");
  ConvGemmDesc d;
  d.src[0] = act(wv_bf16, 1, 256, 256);
  d.weight = xn_bf16;
  d.n_total = (int)m_pad;
  d.block_n = 256;
  d.block_k = 64;
  d.channels = 256;
  d.group_tap_begin[1] = 1;
  d.group_rows[0] = 256;
  d.batch = 1;
  d.epilogue = EPI_GENERIC;
  d.epi = empty_epi();
  d.epi.w_dynamic = 1;   // the streamed operand is the previous kernel's output: never requested before the dependency wait
  d.epi.out1 = vt_bf16;
  d.epi.out_row_stride = m_pad;
  d.epi.out_batch_stride = 0;
  d.epi.res_row_stride = m_pad;
  d.epi.res_batch_stride = 0;
  return launch_convgemm(d, (cudaStream_t)stream);
}

int srb_cfm_attn_out_norm(const void* o_bf16, const void* w_packed, const float* g, const int32_t* lengths, float* x,
                          void* xn_bf16, int32_t batch, int32_t frames, void* stream) {
  ConvGemmDesc d;
  d.src[0] = act(o_bf16, batch, frames, 256);
  d.weight = w_packed;
  d.n_total = 256;
  d.block_n = 256;
  d.block_k = 64;
  d.channels = 256;
  d.group_tap_begin[1] = 1;
  d.group_rows[0] = frames;
  d.batch = batch;
  d.epilogue = EPI_RESNORM;
  d.epi = empty_epi();
  d.epi.norm_mode = 1;
  d.epi.l2_keep = l2_keep_enabled();
  d.epi.vec0 = g;
  d.epi.lengths = lengths;
  d.epi.res[0] = x;
  d.epi.res_row_stride = 256;
  d.epi.res_batch_stride = (long long)frames * 256;
  d.epi.out1 = x;
  d.epi.out0 = xn_bf16;
  d.epi.out_row_stride = 256;
  d.epi.out_batch_stride = (long long)frames * 256;
  return launch_convgemm(d, (cudaStream_t)stream);
}

int srb_cfm_ffn_glu(const void* xn_bf16, const void* w_packed, const float* bias_packed, const int32_t* lengths,
                    void* h_bf16, int32_t batch, int32_t frames, int32_t pad_separated, void* stream) {
  SRB_REQUIRE(lengths != nullptr, "srb_cfm_ffn_glu: lengths required");
  // pad_separated: every utterance ends in at least one pad row (lengths[b] < frames for all b).  Pad rows of xn are zero,
  // so the (B, frames) rows can then be convolved as ONE sequence of B * frames rows -- the zero row between two
  // utterances is exactly the conv's zero padding -- and row tiles need not stop at utterance ends: 252 x 7 tiles
  // instead of 64 x 4 x 7 = 1 792 at 64 x 504 frames, i.e. 12 rounds of the 148 SMs instead of 12.1 -> 13.
  static const int flat_on = [] { const char* e = getenv("SRB_GLU_FLAT"); return e ? atoi(e) : 1; }();   // A/B knob
  const bool flat = flat_on && pad_separated != 0 && batch > 1 && (long long)batch * frames < (1ll << 30);
  ConvGemmDesc d;
  d.src[0] = flat ? act(xn_bf16, 1, batch * frames, 256) : act(xn_bf16, batch, frames, 256);
  d.weight = w_packed;
  d.n_total = 1792;
  d.block_n = 256;
  d.block_k = 64;
  d.channels = 256;
  int ntap = 0;
  same_conv_taps(d, 0, 3, 1, ntap);
  d.group_tap_begin[1] = ntap;
  d.group_rows[0] = flat ? batch * frames : frames;
  d.batch = flat ? 1 : batch;
  d.epilogue = EPI_GLU;
  d.epi = empty_epi();
  d.epi.bias = bias_packed;
  d.epi.lengths = lengths;
  d.epi.flat_frames = flat ? frames : 0;
  d.epi.out0 = h_bf16;
  d.epi.out_row_stride = 896;
  d.epi.out_batch_stride = (long long)frames * 896;
  return launch_convgemm(d, (cudaStream_t)stream);
}

int srb_cfm_ffn_out_norm(const void* h_bf16, const void* w_packed, const float* bias, const float* g, int32_t norm_mode,
                         const int32_t* lengths, float* x, void* xn_bf16, int32_t batch, int32_t frames, void* stream) {
  SRB_REQUIRE(norm_mode == 1 || norm_mode == 2, "srb_cfm_ffn_out_norm: norm_mode must be 1 or 2");
  ConvGemmDesc d;
  d.src[0] = act(h_bf16, batch, frames, 896);
  d.weight = w_packed;
  d.n_total = 256;
  d.block_n = 256;
  d.block_k = 64;
  d.channels = 896;
  int ntap = 0;
  same_conv_taps(d, 0, 3, 1, ntap);
  d.group_tap_begin[1] = ntap;
  d.group_rows[0] = frames;
  d.batch = batch;
  d.epilogue = EPI_RESNORM;
  d.epi = empty_epi();
  d.epi.norm_mode = norm_mode;
  d.epi.l2_keep = l2_keep_enabled();
  d.epi.bias = bias;
  d.epi.vec0 = g;
  d.epi.lengths = lengths;
  d.epi.res[0] = x;
  d.epi.res_row_stride = 256;
  d.epi.res_batch_stride = (long long)frames * 256;
  d.epi.out1 = x;
  d.epi.out0 = xn_bf16;
  d.epi.out_row_stride = 256;
  d.epi.out_batch_stride = (long long)frames * 256;
  return launch_convgemm(d, (cudaStream_t)stream);
}

int srb_cfm_pred_euler(const void* xn_bf16, const void* w_packed, float dt, float* xt, void* xt_bf16, float* mel,
                       void* mel_bf16, int32_t mel_rows, float std, float mean, float pad_value, const int32_t* lengths,
                       int32_t batch, int32_t frames, void* stream) {
  SRB_REQUIRE((mel == nullptr) == (mel_bf16 == nullptr), "srb_cfm_pred_euler: mel and mel_bf16 go together");
  SRB_REQUIRE(mel == nullptr || (mel_rows > 0 && mel_rows <= frames), "srb_cfm_pred_euler: mel_rows must be in (0, frames]");
  ConvGemmDesc d;
  d.src[0] = act(xn_bf16, batch, frames, 256);
  d.weight = w_packed;
  d.n_total = 80;
  d.block_n = 80;
  d.block_k = 64;
  d.channels = 256;
  d.group_tap_begin[1] = 1;
  d.group_rows[0] = frames;
  d.batch = batch;
  d.epilogue = EPI_EULER;
  d.epi = empty_epi();
  d.epi.lengths = lengths;
  d.epi.out1 = xt;
  d.epi.out0 = xt_bf16;
  d.epi.out_row_stride = 80;
  d.epi.out_batch_stride = (long long)frames * 80;
  d.epi.f0 = dt;
  d.epi.f1 = std;
  d.epi.f2 = mean;
  d.epi.f3 = pad_value;
  d.epi.aux0 = mel;
  d.epi.aux1 = mel_bf16;
  d.epi.aux_rows = mel_rows;
  return launch_convgemm(d, (cudaStream_t)stream);
}

int srb_kmeans_scores_argmax(const void* feats_split_bf16, const void* centroids_packed, const float* neg_half_norm2,
                             void* keys_u64, int64_t rows, int32_t dim, int32_t n_padded, void* stream) {
  // x . c_j - |c_j|^2 / 2 over split bf16 operands ([xh | xl | xh] against [Ch | Ch | Cl]: fp32-grade products from the bf16
  // tensor-core loop, see srb_split_factor) with the argmax folded into the epilogue.  The operands are in split form in
  // BOTH libraries (the caller splits explicitly), so this view is built by hand instead of through act().
  SRB_REQUIRE(dim > 0 && dim % 8 == 0, "srb_kmeans_assign: the feature width must be a positive multiple of 8");
  SRB_REQUIRE(n_padded > 0 && n_padded % 256 == 0, "srb_kmeans_assign: the padded centroid count must be a multiple of 256");
  SRB_REQUIRE(rows > 0 && rows < (1ll << 31), "srb_kmeans_assign: bad row count");
  ConvGemmDesc d;
  ActView a;
  a.ptr = feats_split_bf16;
  a.channels = 3 * dim;
  a.rows = (int)rows;
  a.batch = 1;
  a.row_stride = 3 * dim;
  a.batch_stride = (long long)rows * 3 * dim;
  d.src[0] = a;
  d.weight = centroids_packed;
  d.n_total = n_padded;
  d.block_n = 256;
  d.block_k = 64;
  SRB_REQUIRE((3 * dim) % kSplit == 0, "srb_kmeans_assign: width not divisible by the library's split factor");
  d.channels = 3 * dim / kSplit;      // launch_convgemm multiplies by kSplit again
  d.group_tap_begin[1] = 1;
  d.group_rows[0] = (int)rows;
  d.batch = 1;
  d.epilogue = EPI_ARGMAX;
  d.epi = empty_epi();
  d.epi.bias = neg_half_norm2;
  d.epi.out1 = keys_u64;
  d.epi.out_row_stride = 1;
  d.epi.out_batch_stride = rows;
  return launch_convgemm(d, (cudaStream_t)stream);
}

int srb_kmeans_assign(const float* feats, const void* centroids_packed, const float* neg_half_norm2, void* split_ws,
                      uint64_t* keys_ws, int64_t* units, int64_t rows, int32_t dim, int32_t n_padded, int32_t id_offset,
                      const int32_t* lengths, int32_t frames, void* stream) {
  if (rows <= 0) return 0;
  int rc = srb_split_bf16(feats, split_ws, rows, dim, keys_ws, stream);
  if (rc) return rc;
  rc = srb_kmeans_scores_argmax(split_ws, centroids_packed, neg_half_norm2, keys_ws, rows, dim, n_padded, stream);
  if (rc) return rc;
  return srb_kmeans_decode(keys_ws, units, rows, id_offset, lengths, frames, stream);
}

int srb_hifigan_conv(const void* x0, const void* x1, const void* x2, int32_t n_src, const int32_t* kernel,
                     const int32_t* dilation, const void* w_packed, const float* bias, const void* res0,
                     const void* res1, const void* res2, void* out_raw, void* out_act, int32_t batch, int32_t rows,
                     int32_t c_in, int32_t c_out, float scale, float slope, void* stream) {
  return srb_hifigan_conv_res_act(x0, x1, x2, n_src, kernel, dilation, w_packed, bias, res0, res1, res2, 0.f, out_raw, out_act,
                                  batch, rows, c_in, c_out, scale, slope, stream);
}

int srb_hifigan_conv_res_act(const void* x0, const void* x1, const void* x2, int32_t n_src, const int32_t* kernel,
                             const int32_t* dilation, const void* w_packed, const float* bias, const void* res0,
                             const void* res1, const void* res2, float res_slope, void* out_raw, void* out_act, int32_t batch,
                             int32_t rows, int32_t c_in, int32_t c_out, float scale, float slope, void* stream) {
  SRB_REQUIRE(n_src >= 1 && n_src <= 3, "srb_hifigan_conv: n_src must be 1..3");
  SRB_REQUIRE(res_slope >= 0.f && (res_slope == 0.f || kSplit == 1), "srb_hifigan_conv_res_act: bad residual slope (the tight-precision build takes raw residuals only)");
  const void* xs[3] = {x0, x1, x2};
  ConvGemmDesc d;
  d.n_src = n_src;
  int ntap = 0;
  for (int s = 0; s < n_src; ++s) {
    SRB_REQUIRE(xs[s] != nullptr, "srb_hifigan_conv: source %d is NULL", s);
    SRB_REQUIRE(kernel[s] % 2 == 1 && ntap + kernel[s] <= kMaxTaps, "srb_hifigan_conv: bad kernel size");
    d.src[s] = act(xs[s], batch, rows, c_in);
    same_conv_taps(d, s, kernel[s], dilation[s], ntap);
  }
  d.weight = w_packed;
  d.n_total = c_out;
  d.block_n = c_out > 256 ? 256 : c_out;
  d.block_k = block_k_for(c_in);
  d.channels = c_in;
  d.group_tap_begin[1] = ntap;
  d.group_rows[0] = rows;
  d.batch = batch;
  d.epilogue = EPI_GENERIC;
  d.epi = empty_epi();
  d.epi.bias = bias;
  d.epi.res[0] = res0;
  d.epi.res[1] = res1;
  d.epi.res[2] = res2;
  d.epi.res_row_stride = c_out;
  d.epi.res_batch_stride = (long long)rows * c_out;
  d.epi.out1 = out_raw;
  d.epi.out0 = out_act;
  d.epi.out_row_stride = c_out;
  d.epi.out_batch_stride = (long long)rows * c_out;
  d.epi.scale = scale;
  d.epi.slope = slope;
  d.epi.res_unact = res_slope > 0.f ? 1.f / res_slope : 1.f;
  return launch_convgemm(d, (cudaStream_t)stream);
}

int srb_hifigan_upsample(const void* x, const void* w_packed, const float* bias, void* out_raw, void* out_act,
                         int32_t batch, int32_t rows_in, int32_t c_in, int32_t c_out, int32_t kernel, int32_t stride,
                         float slope, void* stream) {
  SRB_REQUIRE(stride >= 1 && stride <= kMaxGroups, "srb_hifigan_upsample: stride %d unsupported", stride);
  const int pad = (kernel - stride) / 2;
  const int rows_out = (rows_in - 1) * stride - 2 * pad + kernel;
  ConvGemmDesc d;
  d.src[0] = act(x, batch, rows_in, c_in);
  d.weight = w_packed;
  d.n_total = c_out;
  d.block_n = c_out > 256 ? 256 : c_out;
  d.block_k = block_k_for(c_in);
  d.channels = c_in;
  d.n_groups = stride;
  d.row_mul = stride;
  int ntap = 0;
  for (int r = 0; r < stride; ++r) {
    // out[q*stride + r] = sum_m W[:, :, j0 + m*stride]^T x[q + c_r - m],  j0 = (r+pad) % stride, c_r = (r+pad) / stride
    const int j0 = (r + pad) % stride, c_r = (r + pad) / stride;
    d.group_tap_begin[r] = ntap;
    for (int j = j0, m = 0; j < kernel; j += stride, ++m) {
      SRB_REQUIRE(ntap < kMaxTaps, "srb_hifigan_upsample: too many taps");
      d.tap_src[ntap] = 0;
      d.tap_shift[ntap] = c_r - m;
      ++ntap;
    }
    d.group_rows[r] = rows_out > r ? (rows_out - r + stride - 1) / stride : 0;
    d.group_row_add[r] = r;
  }
  d.group_tap_begin[stride] = ntap;
  d.batch = batch;
  d.epilogue = EPI_GENERIC;
  d.epi = empty_epi();
  d.epi.bias = bias;
  d.epi.out1 = out_raw;
  d.epi.out0 = out_act;
  d.epi.out_row_stride = c_out;
  d.epi.out_batch_stride = (long long)rows_out * c_out;
  d.epi.res_row_stride = c_out;
  d.epi.res_batch_stride = (long long)rows_out * c_out;
  d.epi.slope = slope;
  return launch_convgemm(d, (cudaStream_t)stream);
}

}  // extern "C"
