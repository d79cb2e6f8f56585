// One resblock pair of the HiFi-GAN generator as ONE kernel, for the C = 64 stage (HF:1359-1367):
//     x' = x + conv2( leaky_relu( conv1_{dil}( leaky_relu(x) ) ) ),  stored as leaky_relu(x')  (the single-copy form)
// Unfused this is two launches that run at the HBM roofline for the k = 3 / 7 resblocks: conv1 reads the activated tensor
// and writes t, conv2 reads t and the residual and writes the result -- five passes over 328 MB tensors at config 2.
// Here a CTA owns 128 - (k - 1) output rows: it loads the activated input rows with both halos ONCE (TMA, 128-byte
// swizzle), runs conv1 on a 128-row M tile straight from that box (a tap is the same box with the descriptor start moved
// by dil rows), turns the accumulator into the bf16 operand t in shared memory (bias + leaky_relu, zero outside the
// utterance: conv2 zero-pads there), runs conv2 on it (taps = row shifts of the t tile) and finishes with bias + residual
// -- recovered from the SAME box, un-activated (srb_convgemm.cuh: unact) -- + leaky_relu.  Two passes instead of five.
// Both convs' weights (2 k slabs of 64 x 64) stay resident; boxes, t tiles and both accumulators are double buffered and
// the MMA warp issues conv1 of tile i + 1 before conv2 of tile i, so the tensor core works while tile i's operand is built.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/srb.h"
#include "srb_common.h"
#include "srb_convgemm.cuh"   // EpiWarp / stage_slot / scatter_store / unact, PTX wrappers

namespace srb {

struct PairParams {
  CUtensorMap tmX;          // (64 channels, rows, batch) bf16, box (64, box_rows, 1), 128-byte swizzle
  CUtensorMap tmW1, tmW2;   // packed [64][k * 64] bf16, box (64, 64)
  const float* b1;
  const float* b2;
  __nv_bfloat16* out;       // (batch, rows, 64)
  int batch, rows, tiles_per_b, total_tiles;
  int dil, box_rows;
  float slope, res_unact;
  int debug;                // timing experiments only (SRB_PAIR_DEBUG): 1 = issue no MMAs, 2 = skip the output stores (wrong results)
};

template <int K>
struct PairLayout {
  static constexpr int H = (K - 1) / 2;
  static constexpr int R = 128 - (K - 1);                      // output rows of a tile
  static constexpr int slab = 64 * 64 * 2;                     // one tap of one conv
  static constexpr int box_bytes = ((128 + 2 * H * 5) * 128 + 1023) & ~1023;   // widest halo: dilation 5
  static constexpr int t_bytes = (128 + 8) * 128;              // 128 rows + the rows the last taps of conv2 touch (kept zero)
  // Activation boxes in flight.  A box is held from its request until the tile's LAST epilogue has read the residual rows
  // out of it, i.e. for the HBM latency plus the whole conv1 -> t -> conv2 chain (~5-6 us): with two boxes the kernel ran at
  // 3.0 us per tile (k = 3), slower than the two launches it replaces, whose 2-3 resident CTAs keep 70-100 KB in flight per
  // SM.  As many as fit beside the resident weights: six at k = 3, three at k = 7.
  static constexpr int NBOX = K == 3 ? 6 : 3;
  // conv2 of a tile is issued LAG tiles after its conv1 (LAG + 1 t tiles and accumulator pairs).  LAG = 2 (with five boxes)
  // was measured at k = 3 and is no faster than LAG = 1 (246-250 us against 232-243 us per launch at config 2): the operand
  // hand-over is not what paces the kernel.  What does, by the numbers: an M128 N64 K16 MMA reads 6 KB of operands from
  // shared memory, 48 clk of its 128 B/clk on its own, and the boxes arriving, the t tile being written and the residual /
  // staging traffic of the epilogues compete for the same port -- k = 7 costs exactly its 32 extra MMAs more per tile
  // (+1.2 us at 37.5 ns each), i.e. the MMAs run serially with ~0.85 us of other shared-memory work per tile.
  static constexpr int LAG = 1;
  static constexpr int NT = LAG + 1;
  static constexpr int off_w1 = 0;
  static constexpr int off_w2 = off_w1 + K * slab;
  static constexpr int off_box = off_w2 + K * slab;
  static constexpr int off_t = off_box + NBOX * box_bytes;
  static constexpr int off_stage = off_t + NT * t_bytes;        // 8 finishing warps x 2 KB (coalescing buffer of their 32 x 32 output block)
  static constexpr int off_bias = off_stage + 4 * 4096;
  static constexpr int off_bar = off_bias + 2 * 64 * 4;
  static constexpr int n_bars = 1 + 6 * NT + 2 * NBOX;
  static constexpr int off_tmem = off_bar + 8 * n_bars;
  static constexpr int total = off_tmem + 16;
  static_assert(t_bytes % 1024 == 0, "t tiles must keep the 1 KB swizzle alignment");
  static_assert(total + 1024 <= 232448, "pair kernel shared memory");
};

template <int K>
__global__ void __launch_bounds__(576, 1) pair_fused_kernel(const __grid_constant__ PairParams p) {
  using L = PairLayout<K>;
  constexpr int H = L::H, R = L::R;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, 64);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t s_w1 = sbase + L::off_w1, s_w2 = sbase + L::off_w2;
  auto s_box = [&](int s) { return sbase + L::off_box + s * L::box_bytes; };
  auto s_t = [&](int s) { return sbase + L::off_t + s * L::t_bytes; };
  const uint32_t bar0 = sbase + L::off_bar;
  constexpr int NBOX = L::NBOX, NT = L::NT, LAG = L::LAG;
  const uint32_t w_full = bar0;
  auto a1_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto a1_empty = [&](int s) { return bar0 + 8u * (1 + NT + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (1 + 2 * NT + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (1 + 3 * NT + s); };
  auto a2_full = [&](int s) { return bar0 + 8u * (1 + 4 * NT + s); };
  auto a2_empty = [&](int s) { return bar0 + 8u * (1 + 5 * NT + s); };
  auto box_full = [&](int s) { return bar0 + 8u * (1 + 6 * NT + s); };
  auto box_empty = [&](int s) { return bar0 + 8u * (1 + 6 * NT + NBOX + s); };
  const uint32_t tmem_slot = sbase + L::off_tmem;
  float* bias_s = reinterpret_cast<float*>(sgen + L::off_bias);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < NBOX; ++s) {
      mbar_init(box_full(s), 1);
      mbar_init(box_empty(s), 8);
    }
    for (int s = 0; s < NT; ++s) {
      mbar_init(a1_full(s), 1);
      mbar_init(a1_empty(s), 8);
      mbar_init(t_full(s), 8);
      mbar_init(t_empty(s), 1);
      mbar_init(a2_full(s), 1);
      mbar_init(a2_empty(s), 8);
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmW1);
    tma_prefetch_desc(&p.tmW2);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, NT == 2 ? 256 : 512);
    tmem_relinquish();
  }
  // the rows past a t tile's 128 stay zero for the whole kernel (conv2's last taps read them for discarded outputs)
  for (int i = threadIdx.x; i < NT * 8 * 8; i += blockDim.x) {
    const int s = i >> 6, r = 128 + ((i >> 3) & 7), piece = i & 7;
    *reinterpret_cast<uint4*>(sgen + L::off_t + s * L::t_bytes + r * 128 + piece * 16) = make_uint4(0, 0, 0, 0);
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) bias_s[i] = i < 64 ? p.b1[i] : p.b2[i - 64];
  pdl_launch_dependents();
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + L::off_tmem);
  auto acc1 = [&](int s) { return tmem_base + s * 64; };
  auto acc2 = [&](int s) { return tmem_base + NT * 64 + s * 64; };

  // weights never depend on a predecessor: requested before the dependency wait
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(w_full, 2 * K * L::slab);
    for (int j = 0; j < K; ++j) {
      tma_load_2d(s_w1 + j * L::slab, &p.tmW1, w_full, j * 64, 0);
      tma_load_2d(s_w2 + j * L::slab, &p.tmW2, w_full, j * 64, 0);
    }
  }
  pdl_wait();

  const int first = blockIdx.x, stride = gridDim.x;
  const int n_mine = first < p.total_tiles ? (p.total_tiles - first + stride - 1) / stride : 0;

  if (warp == 0) {
    // ---------------- producer: one activation box (with both halos) per tile
    if (lane == 0) {
      for (int i = 0; i < n_mine; ++i) {
        const int tile = first + i * stride, s = i % NBOX;
        const uint32_t ph = (i / NBOX) & 1;
        const int b = tile / p.tiles_per_b, r0 = (tile - b * p.tiles_per_b) * R;
        mbar_wait(box_empty(s), ph ^ 1u);
        mbar_expect_tx(box_full(s), static_cast<uint32_t>(p.box_rows) * 128u);
        tma_load_3d(s_box(s), &p.tmX, box_full(s), 0, r0 - H - H * p.dil, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA issuer: conv1 of tile i, then conv2 of tile i - 1
    mbar_wait(w_full, 0u);
    tc_fence_after();
    for (int step = 0; step < n_mine + LAG; ++step) {
      if (step < n_mine) {
        const int s = step % NT, sb = step % NBOX;
        const uint32_t ph = (step / NT) & 1;
        mbar_wait(box_full(sb), (step / NBOX) & 1);
        mbar_wait(a1_empty(s), ph ^ 1u);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < K; ++j) {
          const uint64_t adesc = umma_smem_desc<128>(s_box(sb) + static_cast<uint32_t>(j * p.dil) * 128u);
          const uint64_t wdesc = umma_smem_desc<128>(s_w1 + j * L::slab);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (!(p.debug & 1)) umma_bf16_pred(1u, acc1(s), adesc + 2 * k, wdesc + 2 * k, IDESC, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit_pred(1u, a1_full(s));
      }
      if (step >= LAG) {
        const int i = step - LAG, s = i % NT;
        const uint32_t ph = (i / NT) & 1;
        mbar_wait(t_full(s), ph);
        mbar_wait(a2_empty(s), ph ^ 1u);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < K; ++j) {
          const uint64_t adesc = umma_smem_desc<128>(s_t(s) + static_cast<uint32_t>(j) * 128u);
          const uint64_t wdesc = umma_smem_desc<128>(s_w2 + j * L::slab);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (!(p.debug & 1)) umma_bf16_pred(1u, acc2(s), adesc + 2 * k, wdesc + 2 * k, IDESC, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit_pred(1u, a2_full(s));
        umma_commit_pred(1u, t_empty(s));
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue warps: thread <-> (accumulator row, 32-column half).  Two groups of EIGHT warps (two per TMEM
    // lane quarter, one per column half) work concurrently: warps 2-9 build the operand t of tile i (E1) while warps 10-17
    // finish tile i - 1 (E2).  (One group of four doing both in turn: 3.0 / 4.4 us per tile at k = 3 / 7, slower than the two
    // launches this kernel replaces; two groups of four: no better, the box ring was the limit; see NBOX.)
    const int quarter = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;             // column half of this warp
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl = p.slope, ru = p.res_unact;
    if (warp < 10) {
      for (int i = 0; i < n_mine; ++i) {
        const int tile = first + i * stride, s = i % NT;
        const uint32_t ph = (i / NT) & 1;
        const int b = tile / p.tiles_per_b, r0 = (tile - b * p.tiles_per_b) * R;
        const int g = r0 - H + row;                      // global row of this thread's t row
        const bool inside = g >= 0 && g < p.rows;
        mbar_wait(a1_full(s), ph);
        tc_fence_after();
        mbar_wait(t_empty(s), ph ^ 1u);                  // conv2 of the tile two back has read this t buffer
        uint8_t* tg = sgen + L::off_t + s * L::t_bytes;
        uint32_t v[32];
        tmem_ld32(acc1(s) + lane_addr + half * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = half * 32 + 8 * j + 2 * e;
            const float a0 = lrelu(__uint_as_float(v[8 * j + 2 * e]) + bias_s[col], sl);
            const float a1 = lrelu(__uint_as_float(v[8 * j + 2 * e + 1]) + bias_s[col + 1], sl);
            o[e] = inside ? pack_bf16(a0, a1) : 0u;
          }
          *stage_slot<8>(tg, row, half * 4 + j) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(t_full(s));
          mbar_arrive(a1_empty(s));
        }
      }
    } else {
      EpiWarp ew;
      ew.stage = sgen + L::off_stage + (warp - 10) * 2048;
      ew.lane = lane;
      ew.row0 = quarter * 32;
      for (int i = 0; i < n_mine; ++i) {
        const int s = i % NT, sb = i % NBOX;
        const uint32_t ph = (i / NT) & 1;
        const int tile = first + i * stride;
        const int b = tile / p.tiles_per_b, r0 = (tile - b * p.tiles_per_b) * R;
        mbar_wait(a2_full(s), ph);
        tc_fence_after();
        // residual: the activated input row of this output row, still in the box (box row = row + H + H dil)
        const uint8_t* bg = sgen + L::off_box + sb * L::box_bytes;
        const int brow = row + H + H * p.dil;
        uint4 o[4];
        uint32_t v[32];
        tmem_ld32(acc2(s) + lane_addr + half * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 u = *stage_slot<8>(const_cast<uint8_t*>(bg), brow, half * 4 + j);
          const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
          uint32_t ow[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = half * 32 + 8 * j + 2 * e;
            const float y0 = __uint_as_float(v[8 * j + 2 * e]) + bias_s[64 + col] + unact(bf16_lo(uw[e]), ru);
            const float y1 = __uint_as_float(v[8 * j + 2 * e + 1]) + bias_s[64 + col + 1] + unact(bf16_hi(uw[e]), ru);
            ow[e] = pack_bf16(lrelu(y0, sl), lrelu(y1, sl));
          }
          o[j] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        // accumulator and box are in registers now: hand both back before the (slow) store
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(a2_empty(s));
          mbar_arrive(box_empty(sb));
        }
        int valid = (R < p.rows - r0 ? R : p.rows - r0) - quarter * 32;   // rows of this warp that exist and belong to the tile
        valid = valid < 0 ? 0 : (valid > 32 ? 32 : valid);
        // (every thread storing its own row's 64 bytes instead of staging through shared memory: 248 / 238 us against 252 / 242)
        if (!(p.debug & 2)) scatter_store<4>(ew, o, p.out + ((long long)b * p.rows + r0 + quarter * 32) * 64 + half * 32, 128, valid);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, NT == 2 ? 256 : 512);
}

static PFN_cuTensorMapEncodeTiled_v12000 pair_get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(q);
  }
  return fn;
}

template <int K>
static int launch_pair(PairParams& p, cudaStream_t stream) {
  using L = PairLayout<K>;
  p.tiles_per_b = (p.rows + L::R - 1) / L::R;
  p.total_tiles = p.tiles_per_b * p.batch;
  p.box_rows = 128 + 2 * L::H * p.dil;
  auto kernel = pair_fused_kernel<K>;
  const int smem = L::total + 1024;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    SRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev & 63] = true;
  }
  int grid = num_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (grid < 1) return 0;
  SRB_CUDA(launch_pdl(kernel, dim3(grid), dim3(576), smem, stream, p));
  return after_launch("pair_fused_kernel");
}

}  // namespace srb

using namespace srb;

extern "C" int srb_hifigan_pair_fused(const void* x_act, const void* w1_packed, const float* b1, const void* w2_packed,
                                      const float* b2, void* out_act, int32_t batch, int32_t rows, int32_t channels,
                                      int32_t kernel, int32_t dilation, float slope, void* stream) {
  SRB_REQUIRE(kSplit == 1, "srb_hifigan_pair_fused: not available in the tight-precision build");
  SRB_REQUIRE(channels == 64, "srb_hifigan_pair_fused: built for the C = 64 stage (got %d channels)", channels);
  SRB_REQUIRE(kernel == 3 || kernel == 7, "srb_hifigan_pair_fused: kernel size must be 3 or 7 (got %d)", kernel);
  SRB_REQUIRE(dilation >= 1 && dilation <= 5, "srb_hifigan_pair_fused: dilation must be 1..5");
  SRB_REQUIRE(slope > 0.f, "srb_hifigan_pair_fused: the single-copy form needs an invertible leaky_relu (slope > 0)");
  SRB_REQUIRE(x_act != out_act, "srb_hifigan_pair_fused: not an in-place operation (tiles read their neighbours' rows)");
  if (batch <= 0 || rows <= 0) return 0;
  auto enc = pair_get_encode();
  SRB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  PairParams p;
  memset(&p, 0, sizeof(p));
  p.b1 = b1;
  p.b2 = b2;
  p.out = static_cast<__nv_bfloat16*>(out_act);
  p.batch = batch;
  p.rows = rows;
  p.dil = dilation;
  p.slope = slope;
  p.res_unact = 1.f / slope;
  {
    static const int dbg = [] { const char* e = getenv("SRB_PAIR_DEBUG"); return e ? atoi(e) : 0; }();
    p.debug = dbg;
  }
  const int h = (kernel - 1) / 2;
  {
    cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {128, (cuuint64_t)rows * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)(128 + 2 * h * dilation), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&p.tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x_act), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(pair input) failed: %d", (int)r);
  }
  const void* ws[2] = {w1_packed, w2_packed};
  CUtensorMap* wm[2] = {&p.tmW1, &p.tmW2};
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[2] = {(cuuint64_t)kernel * 64, 64};
    cuuint64_t strides[1] = {(cuuint64_t)kernel * 64 * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(wm[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ws[i]), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(pair weights) failed: %d", (int)r);
  }
  return kernel == 3 ? launch_pair<3>(p, (cudaStream_t)stream) : launch_pair<7>(p, (cudaStream_t)stream);
}
