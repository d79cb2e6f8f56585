// Fused multi-receptive-field stage of the HiFi-GAN generator for the narrow stages (C = 16, 32):
//   out = leaky_relu( (resblock_3(u) + resblock_7(u) + resblock_11(u)) / 3 , slope_next )        HF:1308-1367, 1475-1480
// One CTA owns a window of R consecutive time rows of one utterance (R - 120 output rows + the 60-row receptive
// halo of the k = 11 resblock on each side) and runs all 18 convolutions on it without touching HBM:
//   * activations live in shared memory as tcgen05 K-major operands in the un-swizzled "interleaved" layout
//     [channel chunk of 8][row][8 x bf16]: rows are 16 bytes apart, so a conv tap is just the same buffer with the
//     descriptor start address moved by (tap shift) x 16 bytes -- no im2col, no reload;
//   * the raw residual stream x of the running resblock lives in TMEM as an fp32 accumulator that the conv2 MMAs
//     add into (it is initialised with u through an identity-matrix MMA), so residual adds are free and exact;
//   * epilogue warps turn accumulators into the next conv's operand (bias + leaky_relu + zero outside [0, L) +
//     bf16) and write it back to shared memory; two teams of four warps alternate over the 128-row M tiles so the
//     tensor core works on tile m+1 while tile m is being post-processed;
//   * per-tile mbarriers order MMA issue against the epilogue (a conv on tile m needs tiles m-1..m+1 of its input);
//   * weights stream through a double buffer with cp.async.bulk, pre-packed on the host in operand layout.
// HBM traffic: u read once (+ halo), out written once; everything else stays on chip.
//
// Phases (template parameter D).  An M = 128, K = 16 MMA costs 39 / 40 / 48 clk at N = 16 / 32 / 64 (measured,
// tools/probes/umma_rate_probe.cu), so at C = 16 this kernel is bound by its MMA COUNT.  With D = 2 the window is stored
// as two phase buffers (time row r -> phase r % D, q-row r / D) and an accumulator row holds D consecutive time rows
// (N = D C columns: column block d' = time row q D + d').  A dilation-1 conv then needs k - 1 + D MMAs of N = D C per
// 128 q-rows -- input phase e at q-shift s feeds output phase d' through tap c + s D + e - d', the host packs those
// D taps side by side ("phase matrices") -- instead of D k MMAs of N = C.  Dilated convs map every (tap, output phase)
// to one input phase and keep N = C MMAs into the column slice of the output phase.  D = 1 is the plain form.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/srb.h"
#include "srb_common.h"
#include "srb_ptx.cuh"

namespace srb {

struct MrfParams {
  const __nv_bfloat16* u;    // (B, L, C) raw up-sampler output
  __nv_bfloat16* out;        // (B, L, C)
  const __nv_bfloat16* w;    // packed weights, conv order (resblock j, pair q, conv1 then conv2), operand layout
  const float* bias;         // [18][C]
  int batch, rows, tiles_per_b, total_tiles;
  float slope, slope_next;
#ifdef SRB_TRACE
  unsigned long long* trace;   // debug build: CTA 0, roles (issuer 0, first epilogue warp), 256 ordered %globaltimer stamps each
#endif
};

// debug-only timeline (tools/trace_mrf.py): ordered wall-clock stamps per role; the epilogue warp also records clock64 so
// the SM clock during the launch can be read off (cycles / nanoseconds)
#ifdef SRB_TRACE
#define MRF_STAMP(role)                                                                      \
  do {                                                                                       \
    if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && trace_n < 256) {               \
      unsigned long long t_;                                                                 \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                 \
      p.trace[(role) * 256 + trace_n++] = t_;                                                \
      if ((role) == 1) p.trace[512 + trace_n - 1] = clock64();                               \
    }                                                                                        \
  } while (0)
#else
#define MRF_STAMP(role) do { } while (0)
#endif

constexpr int kMrfHalo = 60;    // 6 * (11 - 1): receptive half-width of the k = 11 resblock
constexpr int kMrfGuard = 32;   // zeroed rows on both sides of every operand buffer (largest tap shift is 25)

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

// un-swizzled K-major operand: 8-row groups 128 B apart (SBO), 16-byte K chunks `chunk_stride` bytes apart (LBO)
// (layout and the legality of row-granular start addresses verified on B200 by tools/probes/umma_probe.cu)
__device__ __forceinline__ uint64_t desc_interleaved(uint32_t addr, uint32_t chunk_stride) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((chunk_stride >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(128 >> 4) << 32;
  d |= 1ull << 46;
  return d;
}

__host__ __device__ constexpr int mrf_kernel_size(int j) { return j == 0 ? 3 : (j == 1 ? 7 : 11); }
__host__ __device__ constexpr int mrf_dilation(int q) { return q == 0 ? 1 : (q == 1 ? 3 : 5); }
// tap offset (in taps) of conv (j, q, which) inside the packed weight / its tap count is mrf_kernel_size(j)
__host__ __device__ constexpr int mrf_conv_tap_begin(int j, int q, int which) {
  int base = 0;
  for (int jj = 0; jj < j; ++jj) base += 6 * mrf_kernel_size(jj);
  return base + (2 * q + which) * mrf_kernel_size(j);
}

template <int C, int R, int D>
struct MrfLayout {
  static_assert(D == 1 || D == 2, "one or two phases");
  static constexpr int RQ = R / D;                    // q-rows per window (one accumulator row = D time rows)
  static constexpr int NM = RQ / 128;
  static constexpr int RP = RQ + 2 * kMrfGuard;       // q-rows of one (phase, chunk) operand column, with guards
  static constexpr int CH = C / 8;
  static constexpr int NPC = D * CH;                  // (phase, 8-channel chunk) columns of an operand buffer
  static constexpr int N = D * C;                     // accumulator columns of a tile
  static constexpr int buf_bytes = NPC * RP * 16;
  static constexpr int tap_bytes = C * C * 2;         // one tap matrix [C][C]
  static constexpr int pm_bytes = N * C * 2;          // one phase matrix [D C][C]
  static constexpr int wbuf_bytes = (10 + D) * pm_bytes;   // largest conv blob: k = 11, dilation 1
  // bytes of the packed weights of a conv (k taps, dilation dil) and the offset of conv (j, q, which)
  __host__ __device__ static constexpr int conv_bytes(int k, int dil) {
    return dil == 1 ? (k - 1 + D) * pm_bytes : k * tap_bytes;
  }
  __host__ __device__ static constexpr int conv_offset(int j, int cc) {
    int off = 0;
    for (int jj = 0; jj <= j; ++jj)
      for (int c2 = 0; c2 < (jj < j ? 6 : cc); ++c2)
        off += conv_bytes(mrf_kernel_size(jj), (c2 & 1) ? 1 : mrf_dilation(c2 >> 1));
    return off;
  }
  static constexpr int off_u = 0;
  static constexpr int off_a = buf_bytes;
  static constexpr int off_t = 2 * buf_bytes;
  static constexpr int off_w = 3 * buf_bytes;
  static constexpr int off_ident = off_w + 2 * wbuf_bytes;
  static constexpr int off_bias = off_ident + tap_bytes;
  static constexpr int off_bar = off_bias + 18 * C * 4;
  static constexpr int n_bars = NM + 4 + 1;
  static constexpr int off_cnt = off_bar + 8 * n_bars;     // uint32 written[NM]: monotonic arrival counters
  static constexpr int off_tmem = off_cnt + 4 * ((NM + 3) & ~3);
  static constexpr int total = off_tmem + 16;
  static constexpr int tout = R - 2 * kMrfHalo;
  static_assert(3 * NM * N <= 512, "accumulators exceed tensor memory");
  static_assert(N == 16 || N == 32, "the epilogue reads an accumulator row with one 16- or 32-column TMEM load");
};

// epilogue teams of four warps (one per TMEM lane quarter); team t post-processes the M tiles m = t (mod NTEAMS).
// Every (conv, tile) step is a latency chain (mbarrier wake-up, TMEM load, smem store, proxy fence, arrive), so the
// number of teams -- not arithmetic -- sets the epilogue rate.
// The MMA side is issue-bound too (measured: ~200 instructions of descriptor/predicate bookkeeping per (conv, tile)
// step on one thread), so NI issuer warps share the M tiles (issuer i owns tiles m = i (mod NI)); M tiles are
// independent accumulators, so no ordering between issuers is needed beyond the per-tile barriers.
template <int C>
struct MrfTeams {
#ifdef SRB_MRF_TEAMS
  static constexpr int value = C == 16 ? SRB_MRF_TEAMS : SRB_MRF_TEAMS32;
#else
  static constexpr int value = C == 16 ? 4 : 3;          // epilogue teams
#endif
#ifdef SRB_MRF_ISSUERS
  static constexpr int issuers = C == 16 ? SRB_MRF_ISSUERS : SRB_MRF_ISSUERS32;
#else
  // MMA issuer warps.  Measured at C = 16 with two phases (us per launch, config 2): 1 issuer 2907, 2: 2220, 3: 1995,
  // 4: 2190, 5: 2362 (four or five epilogue teams make no difference)
  // and at C = 32 (teams, issuers): (3,3) 2611, (4,3) 2664, (3,4) 2724, (2,5) 2713, (3,5) 2852, (4,5) 2912
  static constexpr int issuers = 3;
#endif
  static constexpr int threads = 32 * (1 + issuers) + 128 * value;
};

template <int C, int R, int D>
__global__ void __launch_bounds__(MrfTeams<C>::threads, 1) mrf_fused_kernel(const MrfParams p) {
  using L = MrfLayout<C, R, D>;
  constexpr int NM = L::NM, RP = L::RP, CH = L::CH, KS = C / 16, G = kMrfGuard, N = L::N, NPC = L::NPC;
  constexpr int NTEAMS = MrfTeams<C>::value, EPI_THREADS = 128 * NTEAMS, NI = MrfTeams<C>::issuers;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, C);      // one output phase (dilated convs, identity)
  constexpr uint32_t IDESC_N = umma_idesc_bf16(128, N);    // all output phases at once (dilation-1 convs)
  constexpr uint32_t TCOLS = 512;

  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t s_u = sbase + L::off_u, s_a = sbase + L::off_a, s_t = sbase + L::off_t;
  const uint32_t s_w = sbase + L::off_w, s_ident = sbase + L::off_ident;
  float* bias_s = reinterpret_cast<float*>(smem + L::off_bias);
  const uint32_t bar0 = sbase + L::off_bar;
  auto acc_ready = [&](int m) { return bar0 + 8u * m; };
  // "operand tile m has been (re)written" is signalled with MONOTONIC COUNTERS (4 warp arrivals per event), not
  // mbarrier phases: several issuers wait on a neighbour tile's events, a fast neighbour can run two events ahead
  // of a slow waiter at resblock / window boundaries, and a parity wait would then alias (and deadlock).
  const uint32_t cnt0 = sbase + L::off_cnt;
  auto written_cnt = [&](int m) { return cnt0 + 4u * m; };
  auto w_full = [&](int b) { return bar0 + 8u * (NM + b); };
  auto w_free = [&](int b) { return bar0 + 8u * (NM + 2 + b); };
  const uint32_t u_ready = bar0 + 8u * (NM + 4);
  const uint32_t tmem_slot = sbase + L::off_tmem;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform (keeps role code on the uniform datapath)
#ifdef SRB_TRACE
  int trace_n = 0;
#endif

  // ---- one-time setup: barriers, TMEM, zero guards, identity operand, biases
  if (tid == 0) {
    for (int m = 0; m < NM; ++m) {
      mbar_init(acc_ready(m), 1);
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(written_cnt(m)), "r"(0u) : "memory");
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(w_full(b), 1);
      mbar_init(w_free(b), NI);
    }
    mbar_init(u_ready, 4 * NTEAMS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TCOLS);
    tmem_relinquish();
  }
  {
    // zero all three operand buffers once (guards stay zero forever; interiors are rewritten per tile)
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < 3 * L::buf_bytes / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    // identity weight in operand layout: element (n, k) at chunk (k/8), row n, lane k%8
    __nv_bfloat16* id = reinterpret_cast<__nv_bfloat16*>(smem + L::off_ident);
    for (int i = tid; i < C * C; i += blockDim.x) {
      const int ch = i / (C * 8), n = (i / 8) % C, e = i % 8;
      id[i] = __float2bfloat16_rn((ch * 8 + e) == n ? 1.f : 0.f);
    }
    for (int i = tid; i < 18 * C; i += blockDim.x) bias_s[i] = p.bias[i];
  }
  pdl_launch_dependents();
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // setup above read only weights / biases; the up-sampler output is read from here on
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L::off_tmem);
  const uint32_t t_dt = tmem_base, t_dx = tmem_base + NM * N, t_f = tmem_base + 2 * NM * N;

  if (warp == 0) {
    // ================= weight producer =================
    if (lane == 0) {
      uint32_t n_loaded = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        for (int j = 0; j < 3; ++j) {
          const int k = mrf_kernel_size(j);
          for (int cc = 0; cc < 6; ++cc, ++n_loaded) {
            const int buf = n_loaded & 1;
            const uint32_t ph = (n_loaded >> 1) & 1;
            mbar_wait(w_free(buf), ph ^ 1u);
            const uint32_t bytes = L::conv_bytes(k, (cc & 1) ? 1 : mrf_dilation(cc >> 1));
            mbar_expect_tx(w_full(buf), bytes);
            bulk_g2s(s_w + buf * L::wbuf_bytes, reinterpret_cast<const uint8_t*>(p.w) + L::conv_offset(j, cc), bytes, w_full(buf));
          }
        }
      }
    }
    __syncwarp();
  } else if (warp <= NI) {
    // ================= MMA issuers =================
    // The whole warp runs the loop converged (uniform descriptor arithmetic); lane 0 is the one that issues.
    {
      const uint32_t leader = lane == 0 ? 1u : 0u;
      const int mi = warp - 1;   // this issuer owns tiles m = mi (mod NI)
      uint32_t n_conv = 0;       // convs issued so far (selects weight buffer / phase)
      uint32_t n_tiles_done = 0;
      // event e of tile m is complete once 4 * (e + 1) warp arrivals have been counted
      auto wait_written = [&](int m, uint32_t e) {
        const uint32_t target = 4u * (e + 1u);
        uint32_t spins = 0, v;
        do {
          asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(written_cnt(m)) : "memory");
          if (++spins > (1u << 26)) __trap();
        } while (v < target);
      };
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++n_tiles_done) {
        mbar_wait(u_ready, n_tiles_done & 1);
        tc_fence_after();
        if (mi == 0) MRF_STAMP(0);   // window's raw rows are in shared memory
        // events on a tile per window: per resblock 7 (E0 operand init, then conv1/conv2 of 3 pairs) => 21
        uint32_t ev = n_tiles_done * 21u;
        for (int j = 0; j < 3; ++j) {
          const int k = mrf_kernel_size(j);
          for (int cc = 0; cc < 6; ++cc, ++n_conv, ++ev) {
            // this conv consumes the operand produced by written-event `ev`
            const int which = cc & 1, q = cc >> 1;
            const int dil = which == 0 ? mrf_dilation(q) : 1;
            const uint32_t src = which == 0 ? s_a : s_t;
            const int buf = n_conv & 1;
            mbar_wait(w_full(buf), (n_conv >> 1) & 1);
            const uint32_t wb = s_w + buf * L::wbuf_bytes;
            // descriptors advance in 16-byte units: one q-row = +1, one K slice (two 8-channel chunks) = +2*RP (operand) or
            // +2*rows (weights), one phase of an operand = +CH*RP
            const uint64_t a_desc0 = desc_interleaved(src, RP * 16);
            const uint64_t u_desc0 = desc_interleaved(s_u, RP * 16);
            const uint64_t wn_desc0 = desc_interleaved(wb, N * 16);    // phase matrices [N][C]
            const uint64_t wc_desc0 = desc_interleaved(wb, C * 16);    // tap matrices [C][C]
            const uint64_t i_desc0 = desc_interleaved(s_ident, C * 16);
            const int c_tap = (k - 1) / 2;
            for (int m = mi; m < NM; m += NI) {
              if (m > 0) wait_written(m - 1, ev);
              wait_written(m, ev);
              if (m + 1 < NM) wait_written(m + 1, ev);
              tc_fence_after();
              const uint32_t dacc = (which == 0 ? t_dt : t_dx) + m * N;
              if (cc == 0) {
                // x := u (identity MMA, phase by phase) for this resblock, before the first conv2 accumulates into it
#pragma unroll
                for (int e = 0; e < D; ++e)
#pragma unroll
                  for (int s = 0; s < KS; ++s)
                    umma_bf16_pred(leader, t_dx + m * N + e * C, u_desc0 + (uint64_t)(e * CH * RP + m * 128 + G + s * 2 * RP),
                                   i_desc0 + (uint64_t)(s * 2 * C), IDESC, s != 0 ? 1u : 0u);
              }
              const uint64_t a_row = a_desc0 + (uint64_t)(m * 128 + G);
              if (dil == 1) {
                // phase matrix pi: input time offset o = pi - c, i.e. input phase e = o mod D at q-shift (o - e) / D
                for (int pi = 0; pi < k - 1 + D; ++pi) {
                  const int o = pi - c_tap;
                  const int e = ((o % D) + D) % D;
                  const int sh = (o - e) / D;
                  const uint64_t a_desc = a_row + (uint64_t)(int64_t)(e * CH * RP + sh);
                  const uint64_t w_desc = wn_desc0 + (uint64_t)(pi * (L::pm_bytes >> 4));
#pragma unroll
                  for (int s = 0; s < KS; ++s)
                    umma_bf16_pred(leader, dacc, a_desc + (uint64_t)(s * 2 * RP), w_desc + (uint64_t)(s * 2 * N), IDESC_N,
                                   (which == 1 || pi != 0 || s != 0) ? 1u : 0u);
                }
              } else {
                // dilated conv1: (tap t, output phase d') reads input phase e at q-shift sh into column slice d' C
                for (int t = 0; t < k; ++t) {
                  const uint64_t w_desc = wc_desc0 + (uint64_t)(t * (L::tap_bytes >> 4));
#pragma unroll
                  for (int dp = 0; dp < D; ++dp) {
                    const int o = dp + (t - c_tap) * dil;
                    const int e = ((o % D) + D) % D;
                    const int sh = (o - e) / D;
                    const uint64_t a_desc = a_row + (uint64_t)(int64_t)(e * CH * RP + sh);
#pragma unroll
                    for (int s = 0; s < KS; ++s)
                      umma_bf16_pred(leader, dacc + dp * C, a_desc + (uint64_t)(s * 2 * RP), w_desc + (uint64_t)(s * 2 * C), IDESC,
                                     (t != 0 || s != 0) ? 1u : 0u);
                  }
                }
              }
              umma_commit_pred(leader, acc_ready(m));
            }
            umma_commit_pred(leader, w_free(buf));
            if (mi == 0) MRF_STAMP(0);   // issuer 0 has issued its tiles of this conv
          }
          ++ev;  // the 7th event of the resblock (epilogue of its last conv) is not consumed by any conv
        }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue / operand-builder warps =================
    const int ew = warp - 1 - NI;
    const int team = ew >> 2;             // tiles m = team (mod NTEAMS)
    const int quarter = warp & 3;         // TMEM lane quarter
    const int etid = tid - 32 * (1 + NI);
    auto signal_written = [&](int m) {
      asm volatile("red.release.cta.shared.add.u32 [%0], %1;" ::"r"(written_cnt(m)), "r"(1u) : "memory");
    };
    uint32_t n_tiles_done = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++n_tiles_done) {
      const int b = tile / p.tiles_per_b;
      const int t0 = (tile % p.tiles_per_b) * L::tout - kMrfHalo;     // global row of window row 0
      const __nv_bfloat16* ub = p.u + (size_t)b * p.rows * C;
      if (ew == 0) MRF_STAMP(1);   // window start
      // ---- load the raw window into the U operand buffer (rows outside [0, L) are zero): time row r -> phase r % D,
      // q-row r / D
      for (int i = etid; i < R * CH; i += EPI_THREADS) {
        const int row = i / CH, ch = i % CH;
        const int gr = t0 + row;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (gr >= 0 && gr < p.rows) v = __ldg(reinterpret_cast<const uint4*>(ub + (size_t)gr * C) + ch);
        *reinterpret_cast<uint4*>(smem + L::off_u + (((row % D) * CH + ch) * RP + (row / D) + G) * 16) = v;
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // every epilogue warp finished writing U
      if (lane == 0) mbar_arrive(u_ready);
      if (ew == 0) MRF_STAMP(1);   // raw rows loaded

      uint32_t acc_ev = n_tiles_done * 18u;   // events on acc_ready[m]: one per conv

      for (int j = 0; j < 3; ++j) {
        // ---- E0: A = leaky_relu(U) for my tiles (all phases of my q-row)
        for (int m = team; m < NM; m += NTEAMS) {
          const int row = m * 128 + quarter * 32 + lane;
#pragma unroll
          for (int pc = 0; pc < NPC; ++pc) {
            const uint4 v = *reinterpret_cast<const uint4*>(smem + L::off_u + pc * RP * 16 + (row + G) * 16);
            uint4 o;
            o.x = pack_bf16(lrelu(bf16_lo(v.x), p.slope), lrelu(bf16_hi(v.x), p.slope));
            o.y = pack_bf16(lrelu(bf16_lo(v.y), p.slope), lrelu(bf16_hi(v.y), p.slope));
            o.z = pack_bf16(lrelu(bf16_lo(v.z), p.slope), lrelu(bf16_hi(v.z), p.slope));
            o.w = pack_bf16(lrelu(bf16_lo(v.w), p.slope), lrelu(bf16_hi(v.w), p.slope));
            *reinterpret_cast<uint4*>(smem + L::off_a + pc * RP * 16 + (row + G) * 16) = o;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) signal_written(m);
        }

        float bsum[C];
#pragma unroll
        for (int c = 0; c < C; ++c) bsum[c] = 0.f;
        for (int cc = 0; cc < 6; ++cc, ++acc_ev) {
          const int which = cc & 1, q = cc >> 1;
          const float* bias = bias_s + (j * 6 + cc) * C;
          if (which == 1) {
#pragma unroll
            for (int c = 0; c < C; ++c) bsum[c] += bias[c];
          }
          for (int m = team; m < NM; m += NTEAMS) {
            const int row = m * 128 + quarter * 32 + lane;     // q-row: accumulator column block d is time row row*D + d
            bool inside[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
              const int gr = t0 + row * D + d;
              inside[d] = gr >= 0 && gr < p.rows;
            }
            mbar_wait(acc_ready(m), acc_ev & 1);
            tc_fence_after();
            const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
            float y[N];
            if (which == 0) {
              // conv1 -> T = leaky_relu(acc + b1), zero outside the utterance (the next conv zero-pads there)
              tmem_ld_f<N>(t_dt + m * N + lane_addr, y);
#pragma unroll
              for (int pc = 0; pc < NPC; ++pc) {
                const int d = pc / CH, ch = pc % CH;
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float a0 = lrelu(y[d * C + ch * 8 + 2 * e] + bias[ch * 8 + 2 * e], p.slope);
                  const float a1 = lrelu(y[d * C + ch * 8 + 2 * e + 1] + bias[ch * 8 + 2 * e + 1], p.slope);
                  o[e] = inside[d] ? pack_bf16(a0, a1) : 0u;
                }
                *reinterpret_cast<uint4*>(smem + L::off_t + pc * RP * 16 + (row + G) * 16) = make_uint4(o[0], o[1], o[2], o[3]);
              }
              fence_proxy_async_smem();
            } else if (q < 2) {
              // conv2 of pairs 0,1: x = acc (u + all conv2 so far) + their biases; A = leaky_relu(x)
              tmem_ld_f<N>(t_dx + m * N + lane_addr, y);
#pragma unroll
              for (int pc = 0; pc < NPC; ++pc) {
                const int d = pc / CH, ch = pc % CH;
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float a0 = lrelu(y[d * C + ch * 8 + 2 * e] + bsum[ch * 8 + 2 * e], p.slope);
                  const float a1 = lrelu(y[d * C + ch * 8 + 2 * e + 1] + bsum[ch * 8 + 2 * e + 1], p.slope);
                  o[e] = inside[d] ? pack_bf16(a0, a1) : 0u;
                }
                *reinterpret_cast<uint4*>(smem + L::off_a + pc * RP * 16 + (row + G) * 16) = make_uint4(o[0], o[1], o[2], o[3]);
              }
              fence_proxy_async_smem();
            } else {
              // last conv2 of resblock j: x_final = acc + biases; F (+)= x_final; after the third resblock write out
              tmem_ld_f<N>(t_dx + m * N + lane_addr, y);
#pragma unroll
              for (int n = 0; n < N; ++n) y[n] += bsum[n % C];
              if (j > 0) {
                float f[N];
                tmem_ld_f<N>(t_f + m * N + lane_addr, f);
#pragma unroll
                for (int n = 0; n < N; ++n) y[n] += f[n];
              }
              if (j < 2) {
                if constexpr (N == 32) {
                  uint32_t v[32];
#pragma unroll
                  for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(y[c]);
                  tmem_st32(t_f + m * N + lane_addr, v);
                } else {
                  uint32_t v[16];
#pragma unroll
                  for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(y[c]);
                  tmem_st16(t_f + m * N + lane_addr, v);
                }
                tmem_st_wait();
              } else {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                  const int wr = row * D + d;          // window time row
                  if (inside[d] && wr >= kMrfHalo && wr < kMrfHalo + L::tout) {
                    __nv_bfloat16* ob = p.out + ((size_t)b * p.rows + (t0 + wr)) * C;
#pragma unroll
                    for (int ch = 0; ch < CH; ++ch) {
                      uint32_t o[4];
#pragma unroll
                      for (int e = 0; e < 4; ++e)
                        o[e] = pack_bf16(lrelu(y[d * C + ch * 8 + 2 * e] * (1.f / 3.f), p.slope_next),
                                         lrelu(y[d * C + ch * 8 + 2 * e + 1] * (1.f / 3.f), p.slope_next));
                      reinterpret_cast<uint4*>(ob)[ch] = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                  }
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) signal_written(m);
          }
          if (ew == 0) MRF_STAMP(1);   // team 0 finished the epilogues of conv (j, cc)
        }
      }
      // all eight warps must be done reading U / writing before the next window overwrites the buffers
      asm volatile("bar.sync 2, %0;" ::"n"(EPI_THREADS) : "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TCOLS);
}

template <int C, int R, int D>
static int launch_mrf(const MrfParams& p0, cudaStream_t stream) {
  using L = MrfLayout<C, R, D>;
  MrfParams p = p0;
  p.tiles_per_b = (p.rows + L::tout - 1) / L::tout;
  p.total_tiles = p.tiles_per_b * p.batch;
  auto kernel = mrf_fused_kernel<C, R, D>;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    SRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total));
    configured[dev & 63] = true;
  }
  int grid = num_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (grid < 1) return 0;
  SRB_CUDA(launch_pdl(kernel, dim3(grid), dim3(MrfTeams<C>::threads), L::total, stream, p));
  return after_launch("mrf_fused_kernel");
}

}  // namespace srb

using namespace srb;

extern "C" int srb_hifigan_mrf_phases(int32_t channels) { return channels == 16 ? 2 : 1; }

extern "C" int srb_hifigan_mrf_fused(const void* u_raw, const void* w_packed, const float* bias, void* out_act,
                                     int32_t batch, int32_t rows, int32_t channels, float slope, float slope_next,
                                     void* stream) {
  SRB_REQUIRE(kSplit == 1, "srb_hifigan_mrf_fused: not available in the tight-precision build");
  MrfParams p;
  p.u = static_cast<const __nv_bfloat16*>(u_raw);
  p.out = static_cast<__nv_bfloat16*>(out_act);
  p.w = static_cast<const __nv_bfloat16*>(w_packed);
  p.bias = bias;
  p.batch = batch;
  p.rows = rows;
  p.tiles_per_b = 0;
  p.total_tiles = 0;
  p.slope = slope;
  p.slope_next = slope_next;
#ifdef SRB_TRACE
  p.trace = debug_trace_buffer();
#endif
  if (batch <= 0 || rows <= 0) return 0;
  // weight layout: see srb_hifigan_mrf_phases() -- two phases at C = 16 (MMA-count bound), one at C = 32 (TMEM holds
  // only two M tiles of N = 64 accumulators, which would break the tile pipelining)
  if (channels == 16) return launch_mrf<16, 1280, 2>(p, (cudaStream_t)stream);
  if (channels == 32) return launch_mrf<32, 640, 1>(p, (cudaStream_t)stream);
  set_error("srb_hifigan_mrf_fused: channels must be 16 or 32 (got %d)", channels);
  return -2;
}
