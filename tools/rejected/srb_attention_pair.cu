// tcgen05 / TMEM attention for the velocity transformer (transformer.py:109-127): 2 heads x d_head 128, scale
// 1/sqrt(128), keys >= len_b masked, no dropout.  q | k come from the (B, N, ld) projection buffer, V from its transposed
// copy V^T ([256][utterance * frames], keys contiguous); o (B, N, 256).
//
// One work item = one (utterance, head, PAIR of 128-query tiles): the two query tiles share every K and V tile, which
// halves the K/V bytes an SM pulls per unit of work.  The first form of this kernel (one query tile per item,
// srb_attention_tc.cu) kept 64 KB of K/V loads in flight per SM against a ~1.2 us load latency and needed 64 KB per
// 128 x 128 tile step: 1.2 us per step, whatever the tensor core or the softmax warps could do.
//
//   S_t = Q_t K^T : tcgen05.mma M=128 (queries) x N=128 (keys) x K=128 (d); A = Q tile, B = K tile, both K-major
//                   (d contiguous) through TMA (128-byte swizzle); one fp32 S buffer per query tile t in {A, B};
//   P_t           : bf16, written by the softmax warps into ONE swizzled shared-memory tile used by A and B in turn;
//   O_t += P_t V  : A = P (K-major: keys contiguous), B = V^T tile (K-major as well).
// Operand forms were chosen by measurement (tools/probes/umma_n128_rate_probe.cu, B200): an M128 N128 K16 MMA with both
// operands K-major in shared memory issues every <= 100 clk, with the A operand in tensor memory ("TS" form) every
// 140 clk, with an MN-major B (V as stored, no transposed copy) every 150 clk.  A version of this kernel built on the TS
// form and MN-major V (P written in place over S, no P tile, no V^T) was correct but tensor-pipe bound at 2.1 us per
// tile pair (profiles/r02_trace_attention2_ts_mn_*.txt); this one spends the shared memory instead.
// The two query tiles ping-pong: while the softmax warps work on S_A(j) the tensor core runs P_B V(j-1) and S_B(j).
// O is never rescaled (see below), so there is no TMEM read-modify-write on the critical path.
//
// Softmax is exact.  Two forms, chosen per utterance and head:
//   * single pass (the common case).  softmax is shift invariant, and in floating point (fp32 sums, bf16 P: both carry
//     the fp32 exponent range) ANY shift works as long as nothing overflows or the row's largest term underflows.
//     The caller supplies max |q|^2 and max |k|^2 per (utterance, head) (the projection kernel records them in its
//     epilogue); by Cauchy-Schwarz every scaled logit lies in [-B, B], B = |q|max |k|max log2(e)/sqrt(128).  When
//     B <= 100 the kernel uses shift 0: P = exp2(s*scale) in [2^-100, 2^100], one sweep over K and V;
//   * two passes otherwise (or when no bounds are given): pass 1 runs Q K^T over all key tiles and keeps only the row
//     maxima, pass 2 recomputes S and forms P = exp2(s*scale - m) with the FINAL maximum.
//
// Shared memory: seven 32 KB tile slots = one Q pair (2), the P tile (1), a unified K/V ring (4 stages: the two tiles the
// tensor core is working on plus two in flight).
// Warp roles (704 threads): warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocator), warps 2-17 softmax (see there),
// warps 18-21 output (O / l -> bf16).  tcgen05.ld delivers 56 B / clk / SM whatever the number of warps
// (tools/probes/tmem_rate_probe.cu), so reading one fp32 S tile (64 KB) takes 0.62 us -- more than its MMAs or its 16 K
// exponentials at the rate eight warps get out of the MUFU pipe; hence sixteen softmax warps.
// Persistent: CTA c works on items c, c + gridDim.x, ...
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include "../../include/srb.h"
#include "srb_common.h"
#include "srb_convgemm.cuh"   // pair_barrier, PTX wrappers

namespace srb {

struct Attn3Params {
  CUtensorMap tm;      // 3-D (ld, frames, batch) view of the q | k buffer, box (64, 128, 1)
  CUtensorMap tm_vt;   // 2-D (m_pad, 256) view of V^T, box (64, 128)
  const int* lengths;
  const float* qk_norm2_max;   // (B, 2 [q|k], 2 [head], 2 [frequency half]) partial bounds of the squared row norms, or null
  __nv_bfloat16* out;  // (B, N, 256)
  int frames;
  int q_col, k_col;    // first columns of q and k in the buffer (0, 256)
  int q_tiles;         // 128-query tiles per utterance
  int q_pairs;         // ceil(q_tiles / 2)
  int n_items;         // batch * 2 heads * q_pairs
#ifdef SRB_TRACE
  unsigned long long* trace;   // debug build: per CTA (< 4) and role (producer, MMA, softmax warp 2) 256 time stamps in order
#endif
};

#ifdef SRB_TRACE
#define A3_STAMP(role)                                                                        \
  do {                                                                                        \
    if (p.trace != nullptr && blockIdx.x < 4 && lane == 0 && trace_n < 256) {                 \
      unsigned long long t_;                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                  \
      p.trace[(blockIdx.x * 3 + (role)) * 256 + trace_n++] = t_;                              \
    }                                                                                         \
  } while (0)
#else
#define A3_STAMP(role) do { } while (0)
#endif

constexpr int kA3Tile = 128;
constexpr int kA3Half = 128 * 128;            // one [128 rows][64 bf16] swizzled half tile = 16 KB
constexpr int kA3TileBytes = 2 * kA3Half;     // 32 KB
constexpr int kA3Slots = 7;

struct A3Smem {
  static constexpr int tiles = 0;                              // 7 x 32 KB: Q pair (2), P (1), K/V ring (4)
  static constexpr int red = tiles + kA3Slots * kA3TileBytes;  // [3][128] floats: row max / row sum exchange between column quarters
  static constexpr int lsum = red + 1536;                      // [2 query tiles][128] floats: row sums for the output warps
  static constexpr int items = lsum + 1024;                    // kA3ItemCache work descriptors of this CTA
  static constexpr int bars = items + 240;
  static constexpr int n_bars = 24;
  static constexpr int tmem = bars + 8 * n_bars;
  // no alignment slack: the dynamic shared-memory window of a kernel without static shared memory starts 1 KB aligned
  // (checked at kernel entry: the kernel traps otherwise)
  static constexpr int total = tmem + 16;
  static_assert(total <= 232448, "attention kernel shared memory");
};

// 128-thread named barrier of the four softmax warps that share a TMEM lane quarter (ids 1-4; id 0 is __syncthreads)
__device__ __forceinline__ void quad_barrier(int quarter) {
  switch (quarter) {
    case 0: asm volatile("bar.sync 1, 128;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 128;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 128;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 128;" ::: "memory"); break;
  }
}

constexpr float kA3ScaleLog2 = 0.08838834764831845f * 1.4426950408889634f;   // (1/sqrt(128)) * log2(e)

__device__ __forceinline__ float a3_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct A3Item {
  int b, h, q0, len, nkv;
  short two_pass, nq;
};
constexpr int kA3ItemCache = 10;   // 10 x 24 bytes <= 256

__device__ __forceinline__ A3Item a3_item(const Attn3Params& p, int item) {
  A3Item it;
  const int pr = item % p.q_pairs;
  const int bh = item / p.q_pairs;
  it.h = bh & 1;
  it.b = bh >> 1;
  it.q0 = pr * 2 * kA3Tile;
  it.nq = (pr * 2 + 1 < p.q_tiles) ? 2 : 1;
  int len = p.lengths[it.b];
  it.len = len < p.frames ? len : p.frames;
  it.nkv = (it.len + kA3Tile - 1) / kA3Tile;
  it.two_pass = 1;
  if (p.qk_norm2_max != nullptr) {
    // two partial maxima per head (one per rotary frequency half, see epi_qkv_rope); their sum bounds the row norm
    const float* nq = p.qk_norm2_max + ((it.b * 2 + 0) * 2 + it.h) * 2;
    const float* nk = p.qk_norm2_max + ((it.b * 2 + 1) * 2 + it.h) * 2;
    const float q2 = nq[0] + nq[1], k2 = nk[0] + nk[1];
    // 2 % slack covers the bf16 rounding of q and k after the norms were taken; NaN compares false -> two passes
    it.two_pass = (sqrtf(q2 * k2) * kA3ScaleLog2 * 1.02f <= 100.f) ? 0 : 1;
  }
  return it;
}

__global__ void __launch_bounds__(704, 1) attn3_kernel(const __grid_constant__ Attn3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0u) __trap();   // the swizzled tiles below need 1 KB alignment
  uint8_t* smem = smem_raw;
  constexpr int q_bufs = 1, kv_stages = 4;
  const uint32_t s_q = sbase + A3Smem::tiles;                               // query tile t at s_q + t * 32 KB
  const uint32_t s_p = s_q + 2 * kA3TileBytes;                              // P tile
  const uint32_t s_kv = s_p + kA3TileBytes;                                 // ring stage s: s_kv + s * 32 KB
  const uint32_t bar0 = sbase + A3Smem::bars;
  auto q_full = [&](int s) { return bar0 + 8u * s; };                       // 0..1
  auto q_empty = [&](int s) { return bar0 + 8u * (2 + s); };                // 2..3
  auto kv_full = [&](int s) { return bar0 + 8u * (4 + s); };                // 4..8
  auto kv_empty = [&](int s) { return bar0 + 8u * (9 + s); };               // 9..13
  auto s_full = [&](int t) { return bar0 + 8u * (14 + t); };                // 14..15
  auto s_done = [&](int t) { return bar0 + 8u * (16 + t); };                // 16..17
  const uint32_t o_full = bar0 + 8u * 18, o_empty = bar0 + 8u * 19;
  auto l_ready = [&](int t) { return bar0 + 8u * (20 + t); };             // 20..21
  const uint32_t p_empty = bar0 + 8u * 22;                                  // the P V product that read the P tile has completed
  auto s_read = [&](int t) { return bar0 + 8u * (21 + 2 * t); };            // 21, 23: the softmax warps have loaded S_t into registers
  const uint32_t tmem_slot = sbase + A3Smem::tmem;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
#ifdef SRB_TRACE
  int trace_n = 0;
#endif
  if (warp <= 2) A3_STAMP(warp);   // kernel entry
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
      mbar_init(s_full(s), 1);
      mbar_init(s_done(s), 16);
      mbar_init(s_read(s), 16);
    }
    for (int s = 0; s < 5; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4);
    mbar_init(p_empty, 1);
    mbar_init(l_ready(0), 16);
    fence_barrier_init();
    tma_prefetch_desc(&p.tm);
    tma_prefetch_desc(&p.tm_vt);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // q | k | v, lengths and the norm bounds are produced by the preceding kernels

  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + A3Smem::tmem);
  const uint32_t t_s0 = tmem_base, t_o0 = tmem_base + 256;   // S_A, S_B at +0, +128; O_A, O_B at +256, +384
  constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128);
  const int first = blockIdx.x, step = gridDim.x;
  // the descriptors of this CTA's first items (lengths, pass count) are resolved once, by one thread each: reading
  // them from global memory at every item start cost each role an L2 round trip per item
  A3Item* item_cache = reinterpret_cast<A3Item*>(smem + A3Smem::items);
  if (threadIdx.x < kA3ItemCache && first + (int)threadIdx.x * step < p.n_items)
    item_cache[threadIdx.x] = a3_item(p, first + threadIdx.x * step);
  __syncthreads();
  auto get_item = [&](int n, int item) { return n < kA3ItemCache ? item_cache[n] : a3_item(p, item); };

  if (warp == 0) {
    // ================= TMA producer =================
    int st = 0;
    uint32_t ph = 0;
    auto load_tile = [&](int col, int row, int b) {        // K tile: keys x d
      mbar_wait(kv_empty(st), ph ^ 1u);
      A3_STAMP(0);   // ring slot free, tile requested
      mbar_expect_tx_elect(kv_full(st), kA3TileBytes);
      tma_load_3d_elect(s_kv + st * kA3TileBytes, &p.tm, kv_full(st), col, row, b);
      tma_load_3d_elect(s_kv + st * kA3TileBytes + kA3Half, &p.tm, kv_full(st), col + 64, row, b);
      if (++st == kv_stages) { st = 0; ph ^= 1u; }
    };
    auto load_vt = [&](int col, int row) {                  // V^T tile: d x keys
      mbar_wait(kv_empty(st), ph ^ 1u);
      A3_STAMP(0);
      mbar_expect_tx_elect(kv_full(st), kA3TileBytes);
      tma_load_2d_elect(s_kv + st * kA3TileBytes, &p.tm_vt, kv_full(st), col, row);
      tma_load_2d_elect(s_kv + st * kA3TileBytes + kA3Half, &p.tm_vt, kv_full(st), col + 64, row);
      if (++st == kv_stages) { st = 0; ph ^= 1u; }
    };
    // Q of item nn lives in pair buffer nn % q_bufs
    auto load_q = [&](int nn, const A3Item& iq) {
      const int qb = nn % q_bufs, use = nn / q_bufs;
      mbar_wait(q_empty(qb), (use & 1) ^ 1u);   // every S tile of the item that used this buffer last has been issued and completed
      mbar_expect_tx_elect(q_full(qb), iq.nq * kA3TileBytes);
      for (int t = 0; t < iq.nq; ++t) {
        const uint32_t dst = s_q + (qb * 2 + t) * kA3TileBytes;
        tma_load_3d_elect(dst, &p.tm, q_full(qb), p.q_col + iq.h * 128, iq.q0 + t * kA3Tile, iq.b);
        tma_load_3d_elect(dst + kA3Half, &p.tm, q_full(qb), p.q_col + iq.h * 128 + 64, iq.q0 + t * kA3Tile, iq.b);
      }
    };
    int n = 0;
    for (int item = first; item < p.n_items; item += step, ++n) {
      const A3Item it = get_item(n, item);
      const bool has_next = item + step < p.n_items;
      if (n == 0) load_q(0, it);
      // two Q pairs: the next item's Q is requested before this item's K / V tiles, so it lands long before the tensor
      // core gets there; one pair: it can only follow this item's last S tile, i.e. after this item's loads
      if (q_bufs == 2 && has_next) load_q(n + 1, get_item(n + 1, item + step));
      const int kc = p.k_col + it.h * 128;
      if (it.two_pass)
        for (int j = 0; j < it.nkv; ++j) load_tile(kc, j * kA3Tile, it.b);
      for (int j = 0; j < it.nkv; ++j) {
        load_tile(kc, j * kA3Tile, it.b);
        load_vt(it.b * p.frames + j * kA3Tile, it.h * 128);
      }
      if (q_bufs == 1 && has_next) load_q(n + 1, get_item(n + 1, item + step));
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (converged warp, elected lane issues) =================
    int st = 0;
    uint32_t ph = 0;
    // per S buffer t: tiles issued into it, and how many of its s_read / s_done completions (one each per tile, in order)
    // this warp has consumed -- every completion is waited for exactly once, so a parity wait can never alias
    int uses[2] = {0, 0}, reads[2] = {0, 0}, dones[2] = {0, 0};
    auto ring_next = [&]() { if (++st == kv_stages) { st = 0; ph ^= 1u; } };
    auto ensure_read = [&](int t) {           // the softmax warps hold every tile issued into buffer t in registers: it may be overwritten
      while (reads[t] < uses[t]) {
        mbar_wait(s_read(t), reads[t] & 1);
        ++reads[t];
      }
      tc_fence_after();
    };
    auto wait_p = [&](int t) {                // the next P tile of query tile t is published (s_done completes once per main-sweep tile)
      mbar_wait(s_done(t), dones[t] & 1);
      ++dones[t];
      tc_fence_after();
    };
    auto issue_s = [&](int t, uint32_t qa, int k_slot) {
      ensure_read(t);
      const uint32_t kb = s_kv + k_slot * kA3TileBytes;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * kA3Half + (kk & 3) * 32;
        umma_bf16_pred(1u, t_s0 + t * 128, umma_smem_desc<128>(qa + off), umma_smem_desc<128>(kb + off), IDESC_S, kk != 0 ? 1u : 0u);
      }
      umma_commit_pred(1u, s_full(t));
      ++uses[t];
    };
    int n = 0;
    for (int item = first; item < p.n_items; item += step, ++n) {
      const A3Item it = get_item(n, item);
      const int nq = it.nq, nkv = it.nkv;
      const int qb = n % q_bufs;
      A3_STAMP(1);   // item: waiting for Q
      mbar_wait(q_full(qb), (n / q_bufs) & 1);
      tc_fence_after();
      A3_STAMP(1);   // item: Q landed
      const uint32_t qa0 = s_q + qb * 2 * kA3TileBytes;
      if (nkv == 0) {
        umma_commit_pred(1u, q_empty(qb));
        umma_commit_pred(1u, o_full);
        continue;
      }
      if (it.two_pass) {
        // maxima sweep: S only; a buffer takes its next tile as soon as the softmax warps have loaded the previous one
        for (int j = 0; j < nkv; ++j) {
          mbar_wait(kv_full(st), ph);
          tc_fence_after();
          for (int t = 0; t < nq; ++t) issue_s(t, qa0 + t * kA3TileBytes, st);
          umma_commit_pred(1u, kv_empty(st));
          ring_next();
        }
      }
      // main sweep, prologue: S_t(0)
      mbar_wait(kv_full(st), ph);
      tc_fence_after();
      A3_STAMP(1);   // K_0 landed
      for (int t = 0; t < nq; ++t) issue_s(t, qa0 + t * kA3TileBytes, st);
      umma_commit_pred(1u, kv_empty(st));
      ring_next();
      if (nkv == 1) umma_commit_pred(1u, q_empty(qb));
      for (int j = 0; j < nkv; ++j) {
        // ring: stage `st` holds V_j, the stage after it K_{j+1}
        const int v_st = st;
        const uint32_t v_ph = ph;
        ring_next();
        const int k_st = st;
        const uint32_t k_ph = ph;
        if (j + 1 < nkv) ring_next();
        for (int t = 0; t < nq; ++t) {
          if (j + 1 < nkv) {
            // S_t(j + 1) goes out as soon as the softmax warps have S_t(j) in registers -- ahead of P_t V(j), which has
            // to wait for their arithmetic: the next-but-one tile of the softmax stream is then ready in time
            if (t == 0) {
              mbar_wait(kv_full(k_st), k_ph);
              tc_fence_after();
            }
            issue_s(t, qa0 + t * kA3TileBytes, k_st);
            if (t == nq - 1) {
              umma_commit_pred(1u, kv_empty(k_st));
              if (j + 2 == nkv) umma_commit_pred(1u, q_empty(qb));   // the item's last S tiles are issued: Q may be replaced
            }
          }
          A3_STAMP(1);   // waiting for P_t(j)
          wait_p(t);                              // P_t(j) is in shared memory
          A3_STAMP(1);   // P_t(j) ready
          if (t == 0) {
            mbar_wait(kv_full(v_st), v_ph);
            if (j == 0) mbar_wait(o_empty, (n & 1) ^ 1u);   // the output warps have read the previous item's O
            tc_fence_after();
          }
          const uint32_t vb = s_kv + v_st * kA3TileBytes;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t off = (kk >> 2) * kA3Half + (kk & 3) * 32;
            umma_bf16_pred(1u, t_o0 + t * 128, umma_smem_desc<128>(s_p + off), umma_smem_desc<128>(vb + off), IDESC_S,
                           (j != 0 || kk != 0) ? 1u : 0u);
          }
          umma_commit_pred(1u, p_empty);      // the P tile may take the next tile of the stream
          if (t == nq - 1) umma_commit_pred(1u, kv_empty(v_st));
          A3_STAMP(1);   // P V issued
        }
      }
      umma_commit_pred(1u, o_full);
    }
    __syncwarp();
  } else if (warp >= 18) {
    // ================= output warps =================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float* lsum = reinterpret_cast<const float*>(smem + A3Smem::lsum);
    int n = 0;
    for (int item = first; item < p.n_items; item += step, ++n) {
      const A3Item it = get_item(n, item);
      mbar_wait(l_ready(0), n & 1);
      mbar_wait(o_full, n & 1);
      tc_fence_after();
      for (int t = 0; t < it.nq; ++t) {
        const float l = lsum[t * 128 + row];
        const bool ok = l > 0.f && it.nkv > 0;
        const float inv = ok ? 1.f / l : 0.f;
        const int q = it.q0 + t * kA3Tile + row;
        // thread <-> query row: the row's 128 output columns of this head are 256 contiguous bytes
        uint4* out = reinterpret_cast<uint4*>(p.out + ((long long)it.b * p.frames + q) * 256 + it.h * 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(t_o0 + t * 128 + lane_addr + c * 32, v);
          tmem_ld_wait();
          if (c == 3 && t == it.nq - 1) {
            // O and the row sums are in registers: the accumulators and lsum are free for the next item
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
          }
          if (q < p.frames) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              out[c * 4 + i] = ok ? make_uint4(pack_bf16(__uint_as_float(v[8 * i]) * inv, __uint_as_float(v[8 * i + 1]) * inv),
                                               pack_bf16(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv),
                                               pack_bf16(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv),
                                               pack_bf16(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv))
                                  : make_uint4(0, 0, 0, 0);
          }
        }
      }
    }
  } else {
    // ================= softmax warps (16) =================
    // All sixteen warps work on ONE S tile at a time, in the order the tensor core produces them (A(0), B(0), A(1), ...):
    // thread <-> (query row = TMEM lane, one of four 32-key column quarters).  Sixteen warps, not eight: an S tile costs
    // 16 K exponentials, the MUFU pipe delivers 16 per clock only with >= 16 resident warps (10 with 8:
    // tools/probes/mufu_rate_probe.cu), and four warps per SM sub-partition overlap one warp's tcgen05.ld with the
    // others' arithmetic without any software pipelining.
    const int quarter = warp & 3;                         // TMEM lane quarter this warp may access
    const int cq = (warp - 2) >> 2;                       // key-column quarter of S
    const int row = quarter * 32 + lane;                  // query row inside the tile = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = kA3ScaleLog2;
    float* red = reinterpret_cast<float*>(smem + A3Smem::red);   // [3][128]: partials of column quarters 1..3
    float* lsum = reinterpret_cast<float*>(smem + A3Smem::lsum);
    int usesA = 0, usesB = 0;   // S tiles consumed from buffers A / B
    int pv = 0;                 // P tiles written so far (the P tile's p_empty barrier completes once per P V product)
    // this thread's 32 keys are 64 bytes of its row of the [128 rows][64 keys] half tile (cq / 2) of P: 128-byte rows,
    // 16-byte pieces XOR-swizzled
    uint8_t* prow = smem + (s_p - sbase) + (cq >> 1) * kA3Half + row * 128;
    // combine a per-thread partial over the four column quarters of a row (max or sum); every thread gets the result
    auto combine = [&](float v, bool is_max) {
      if (cq != 0) red[(cq - 1) * 128 + row] = v;
      quad_barrier(quarter);
      if (cq == 0) {
        const float a = red[row], b = red[128 + row], c = red[256 + row];
        v = is_max ? fmaxf(fmaxf(v, a), fmaxf(b, c)) : (v + a) + (b + c);
        red[row] = v;
      }
      quad_barrier(quarter);
      v = red[row];
      quad_barrier(quarter);      // all four have read before `red` is reused
      return v;
    };
    int n = 0;
    for (int item = first; item < p.n_items; item += step, ++n) {
      const A3Item it = get_item(n, item);
      const int len = it.len, nkv = it.nkv, nq = it.nq;
      const int sweeps = (it.two_pass && nkv > 0) ? 2 : 1;
      float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;
      float m2A = 0.f, m2B = 0.f;   // single pass: shift 0 (see the header comment)
      for (int sw = 0; sw < sweeps; ++sw) {
        const bool maxima = sweeps == 2 && sw == 0;
        const int ntiles = nkv * nq;
        for (int k = 0; k < ntiles; ++k) {
          const int t = nq == 2 ? (k & 1) : 0;
          const int jj = nq == 2 ? (k >> 1) : k;
          if (warp == 2) A3_STAMP(2);   // waiting for S_t(j)
          mbar_wait(s_full(t), (t ? usesB : usesA) & 1);
          tc_fence_after();
          if (warp == 2) A3_STAMP(2);   // S_t(j) ready
          if (t) ++usesB; else ++usesA;
          uint32_t v[32];
          tmem_ld32(t_s0 + t * 128 + lane_addr + cq * 32, v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_read(t));   // S_t(j) is in registers: the tensor core may overwrite the buffer
          const int nvalid = len - jj * kA3Tile - cq * 32;   // valid keys among this thread's 32
          if (maxima) {
            float mt = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nvalid) mt = fmaxf(mt, __uint_as_float(v[i]));
            if (t) mB = fmaxf(mB, mt); else mA = fmaxf(mA, mt);
          } else {
            const float m2 = t ? m2B : m2A;
            uint32_t o[16];
            float lt;
            if (nvalid >= 32) {
              float l0 = 0.f, l1 = 0.f;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float p0 = a3_ex2(fmaf(__uint_as_float(v[2 * i]), sl2, -m2));
                const float p1 = a3_ex2(fmaf(__uint_as_float(v[2 * i + 1]), sl2, -m2));
                l0 += p0;
                l1 += p1;
                o[i] = pack_bf16(p0, p1);
              }
              lt = l0 + l1;
            } else {
              lt = 0.f;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float p0 = 2 * i < nvalid ? a3_ex2(fmaf(__uint_as_float(v[2 * i]), sl2, -m2)) : 0.f;
                const float p1 = 2 * i + 1 < nvalid ? a3_ex2(fmaf(__uint_as_float(v[2 * i + 1]), sl2, -m2)) : 0.f;
                lt += p0 + p1;
                o[i] = pack_bf16(p0, p1);
              }
            }
            if (t) lB += lt; else lA += lt;
            mbar_wait(p_empty, (pv & 1) ^ 1u);   // the product that read the previous P tile has completed
            ++pv;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int piece = (cq & 1) * 4 + i;
              *reinterpret_cast<uint4*>(prow + ((piece ^ (row & 7)) << 4)) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            }
            fence_proxy_async_smem();
          }
          if (!maxima) {
            __syncwarp();
            if (lane == 0) mbar_arrive(s_done(t));
          }
          if (warp == 2) A3_STAMP(2);   // P_t(j) published
        }
        if (maxima) {
          // row maxima over the four column quarters
          m2A = combine(mA, true) * sl2;       // finite: every utterance has at least one valid key
          if (nq == 2) m2B = combine(mB, true) * sl2;
        }
      }
      // row sums over the column quarters, handed to the output warps; lsum was last read by the output warps of the previous item
      mbar_wait(o_empty, (n & 1) ^ 1u);
      {
        const float tot = combine(lA, false);
        if (cq == 0) lsum[row] = tot;
      }
      if (nq == 2) {
        const float tot = combine(lB, false);
        if (cq == 0) lsum[128 + row] = tot;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(l_ready(0));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static PFN_cuTensorMapEncodeTiled_v12000 attn3_get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

}  // namespace srb

using namespace srb;

extern "C" int srb_cfm_attention_pair(const void* qk_bf16, int32_t ld, const void* vt_bf16, int64_t m_pad,
                                      const int32_t* lengths, const float* qk_norm2_max, void* o_bf16, int32_t batch,
                                      int32_t frames, void* stream) {
  if (batch <= 0 || frames <= 0) return 0;
  auto enc = attn3_get_encode();
  SRB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  SRB_REQUIRE(ld >= 512 && ld % 8 == 0 && m_pad % 8 == 0 && m_pad >= (int64_t)batch * frames, "srb_cfm_attention_pair: bad strides");
  SRB_REQUIRE(frames % 8 == 0, "srb_cfm_attention_pair: frames must be a multiple of 8 (TMA box origins in v^T must be 16-byte aligned)");
  Attn3Params p;
  {
    cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)frames, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)frames * ld * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qk_bf16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(q|k) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)m_pad, 256};
    cuuint64_t strides[1] = {(cuuint64_t)m_pad * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tm_vt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(vt_bf16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(v^T) failed: %d", (int)r);
  }
  p.lengths = lengths;
#ifdef SRB_TRACE
  p.trace = debug_trace_buffer();
#endif
  p.qk_norm2_max = qk_norm2_max;
  p.out = static_cast<__nv_bfloat16*>(o_bf16);
  p.frames = frames;
  p.q_col = 0;
  p.k_col = 256;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    SRB_CUDA(cudaFuncSetAttribute(attn3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A3Smem::total));
    configured[dev & 63] = true;
  }
  p.q_tiles = (frames + kA3Tile - 1) / kA3Tile;
  p.q_pairs = (p.q_tiles + 1) / 2;
  p.n_items = batch * 2 * p.q_pairs;
  int grid = num_sms();
  if (grid > p.n_items) grid = p.n_items;
  SRB_CUDA(launch_pdl(attn3_kernel, dim3(grid), dim3(704), A3Smem::total, (cudaStream_t)stream, p));
  return after_launch("attn3_kernel");
}
