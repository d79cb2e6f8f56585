// HBM-bound pieces of the path: embedding gather, conditioning tables, depthwise positional conv, prior
// preparation, the 16->1 output conv + tanh with its length-aware (cropping) store.  CUDA-core kernels, vectorised and
// coalesced; none of them is GEMM shaped.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <math.h>

#include "../../include/srb.h"
#include "srb_common.h"
#include "srb_ptx.cuh"

namespace srb {

// ------------------------------------------------------------------------------------------ embedding gather
// One warp per output row; 16-byte loads/stores; the table (<= 6 MB) stays L2 resident.  models.py:154
__global__ void __launch_bounds__(256) embed_gather_kernel(const float4* __restrict__ table, const int64_t* __restrict__ ids,
                                                           float4* __restrict__ out, long long m, int vocab_rows,
                                                           int dim4) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < m; row += warps) {
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= vocab_rows ? vocab_rows - 1 : id);
    const float4* src = table + id * dim4;
    float4* dst = out + row * dim4;
    for (int c = lane; c < dim4; c += 32) {
      float4 v = __ldg(src + c);
      __stcs(dst + c, v);  // streaming store: written once, read by later kernels from L2/HBM
    }
  }
}

// lengths[b] = number of non-zero ids; extents[b] (optional) = index of the last non-zero id + 1.  The path masks by
// prefix length, the reference by position (models.py:152): the two agree exactly when lengths == extents (right-padded
// rows), which the host checks.
__global__ void unit_lengths_kernel(const int64_t* __restrict__ ids, int* __restrict__ lengths, int* __restrict__ extents,
                                    int frames) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  int cnt = 0, ext = 0;
  for (int t = threadIdx.x; t < frames; t += blockDim.x) {
    const bool nz = ids[(long long)b * frames + t] != 0;
    cnt += nz;
    if (nz) ext = t + 1;
  }
  __shared__ int s[32], e[32];
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    ext = max(ext, __shfl_xor_sync(0xffffffffu, ext, o));
  }
  if ((threadIdx.x & 31) == 0) {
    s[threadIdx.x >> 5] = cnt;
    e[threadIdx.x >> 5] = ext;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0, mx = 0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) {
      tot += s[w];
      mx = max(mx, e[w]);
    }
    lengths[b] = tot;
    if (extents != nullptr) extents[b] = mx;
  }
}

// ------------------------------------------------------------------------------------------ time conditioning
// One block (256 threads) per ODE step.  fourier_embed.py:37-40, models.py:47-49, norm.py:42.
__global__ void __launch_bounds__(256) time_cond_kernel(const float* __restrict__ times, const float* __restrict__ four_w,
                                                        const float* __restrict__ lin_w, const float* __restrict__ lin_b,
                                                        const float* __restrict__ gamma_w, int n_norm,
                                                        float* __restrict__ time_emb, float* __restrict__ g) {
  __shared__ float four[257];
  __shared__ float c[256];
  const int s = blockIdx.x, tid = threadIdx.x;
  const float t = times[s];
  if (tid == 0) four[0] = t;
  if (tid < 128) {
    // freqs = ((t * w) * 2) * pi, evaluated in that order in fp32 (fourier_embed.py:38)
    float f = __fmul_rn(__fmul_rn(__fmul_rn(t, four_w[tid]), 2.0f), 3.14159265358979323846f);
    four[1 + tid] = sinf(f);
    four[129 + tid] = cosf(f);
  }
  __syncthreads();
  float acc = lin_b[tid];
  const float* wrow = lin_w + (long long)tid * 257;
  for (int k = 0; k < 257; ++k) acc = fmaf(wrow[k], four[k], acc);
  const float ce = acc / (1.f + expf(-acc));  // SiLU
  c[tid] = ce;
  time_emb[(long long)s * 256 + tid] = ce;
  __syncthreads();
  for (int j = 0; j < n_norm; ++j) {
    const float* w = gamma_w + ((long long)j * 256 + tid) * 256;
    float a = 0.f;
    for (int k = 0; k < 256; ++k) a = fmaf(w[k], c[k], a);
    g[((long long)s * n_norm + j) * 256 + tid] = 16.0f * (a + 1.0f);  // sqrt(256) * (gamma + 1)
  }
}

__global__ void rotary_table_kernel(const float* __restrict__ inv_freq, int rows, float* __restrict__ cs,
                                    float* __restrict__ sn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 64) return;
  const int pos = i >> 6, f = i & 63;
  const float a = __fmul_rn((float)pos, inv_freq[f]);  // transformer.py:61
  cs[i] = cosf(a);
  sn[i] = sinf(a);
}

// torch.clamp(x, -tv, tv) (models.py:169-170), including its min > max rule (the result is max) for a negative tv
__device__ __forceinline__ float clamp_tv(float v, float tv) { return fminf(fmaxf(v, -tv), tv); }

// bf16 copy of piece i (4 values) of an array of 80-value rows; split build: [hi | rest | hi] blocks of 80 columns
__device__ __forceinline__ void put_bf16x4(uint2* __restrict__ dst, long long i, float4 v) {
  const uint2 h = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  if constexpr (kSplit == 1) {
    dst[i] = h;
  } else {
    const long long row = i / 20;
    uint2* d = dst + row * (20 * kSplit) + (i - row * 20);
    d[0] = h;
    d[20] = make_uint2(pack_bf16_rest(v.x, v.y, h.x), pack_bf16_rest(v.z, v.w, h.y));
    d[40] = h;
  }
}

__global__ void prior_prepare_kernel(float4* __restrict__ xt, uint2* __restrict__ xtb, long long n4, float tv, int has_tv) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = xt[i];
    if (has_tv) {
      v.x = clamp_tv(v.x, tv);
      v.y = clamp_tv(v.y, tv);
      v.z = clamp_tv(v.z, tv);
      v.w = clamp_tv(v.w, tv);
      xt[i] = v;
    }
    put_bf16x4(xtb, i, v);
  }
}

// Per-call input staging (one launch ahead of the captured graph): the caller's units (B, n) and prior sample (B, n, 80)
// go into the plan's static buffers, whose rows are padded to n8 >= n: ids -> (B, n8) with zero (= pad) columns, prior ->
// fp32 ODE state (B, n8, 80) with zero pad rows, clamped (models.py:169-170), plus its bf16 copy; two small regions the
// loop relies on being zero (tail rows of the normalised activations, the attention norm bounds) are cleared.
__global__ void __launch_bounds__(256) stage_inputs_kernel(const int64_t* __restrict__ ids_in, const float4* __restrict__ noise,
                                                           int64_t* __restrict__ ids, float4* __restrict__ xt,
                                                           uint2* __restrict__ xtb, uint4* __restrict__ zero_a, long long za16,
                                                           uint4* __restrict__ zero_b, long long zb16, int batch, int n, int n8,
                                                           float tv, int has_tv) {
  pdl_launch_dependents();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_x = (long long)batch * n8 * 20;   // float4 pieces of the ODE state
  for (long long i = tid; i < n_x; i += stride) {
    const long long row = i / 20;
    const int piece = (int)(i - row * 20);
    const long long b = row / n8;
    const int t = (int)(row - b * n8);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < n) {
      v = __ldg(noise + (b * n + t) * 20 + piece);
      if (has_tv) {
        v.x = clamp_tv(v.x, tv);
        v.y = clamp_tv(v.y, tv);
        v.z = clamp_tv(v.z, tv);
        v.w = clamp_tv(v.w, tv);
      }
    }
    xt[i] = v;
    put_bf16x4(xtb, i, v);
  }
  const long long n_ids = (long long)batch * n8;
  for (long long i = tid; i < n_ids; i += stride) {
    const long long b = i / n8;
    const int t = (int)(i - b * n8);
    ids[i] = t < n ? ids_in[b * n + t] : 0;
  }
  for (long long i = tid; i < za16; i += stride) zero_a[i] = make_uint4(0, 0, 0, 0);
  for (long long i = tid; i < zb16; i += stride) zero_b[i] = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------ duration prediction
// Duration-prediction variant (models.py:157-164): d[b, n] = clamp(round(exp(conv_k3(E[ids])[b, n]) - 1), 0), pads 0
// (fastspeech/modules.py:87-107).  The Conv1d(768 -> 1, k 3) over embedding rows is linear in the rows, so it is
// three lookups in a (3, vocab + 1) table of per-unit dot products (dur_table[tap][u] = <W[0, :, tap], E[u]>, built in
// float64 at pack time) + bias.  One block per utterance; it also sums the durations (the expanded length).
__device__ __forceinline__ int64_t dur_clamp_id(int64_t id, int vocab_rows) {
  return id < 0 ? 0 : (id >= vocab_rows ? (int64_t)vocab_rows - 1 : id);
}
__global__ void __launch_bounds__(256) duration_predict_kernel(const int64_t* __restrict__ ids, const float* __restrict__ dur_table,
                                                               float bias, int* __restrict__ durations, int* __restrict__ totals,
                                                               int frames, int vocab_rows) {
  __shared__ int part[8];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int64_t* row = ids + (long long)b * frames;
  int sum = 0;
  for (int n = threadIdx.x; n < frames; n += blockDim.x) {
    const int64_t u = dur_clamp_id(row[n], vocab_rows);
    int d = 0;
    if (u != 0) {
      // zero padding of the conv; ids outside [0, vocab_rows) are clamped like the gather kernel does
      const int64_t um = n > 0 ? dur_clamp_id(row[n - 1], vocab_rows) : 0;
      const int64_t up = n + 1 < frames ? dur_clamp_id(row[n + 1], vocab_rows) : 0;
      const float x = (dur_table[um] + dur_table[vocab_rows + u]) + (dur_table[2 * vocab_rows + up] + bias);
      const float r = rintf(expf(x) - 1.f);   // torch.round: half to even
      d = r > 0.f ? (r < 1048576.f ? (int)r : 1048576) : 0;
    }
    durations[(long long)b * frames + n] = d;
    sum += d;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += part[w];
    totals[b] = t;
  }
}

// length_regulator (HF:88-134) on unit ids: out[b, pos] = ids[b, n] for pos in [cum[n], cum[n] + d[n]), 0 beyond the
// utterance's expanded length.  One block per utterance: chunked inclusive scan of the durations, then every thread
// writes its unit's run.  all_one != 0 applies HF:113-114 (every duration of the batch was 0 -> all become 1).
__global__ void __launch_bounds__(256) length_regulate_kernel(const int64_t* __restrict__ ids, const int* __restrict__ durations,
                                                              int64_t* __restrict__ out, int frames, int frames_out, int all_one) {
  __shared__ int warp_tot[8];
  __shared__ int carry_s;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t* row = ids + (long long)b * frames;
  int64_t* orow = out + (long long)b * frames_out;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < frames; base += blockDim.x) {
    const int n = base + threadIdx.x;
    const int d = n < frames ? (all_one ? 1 : durations[(long long)b * frames + n]) : 0;
    int incl = d;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int off = carry_s;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    const int start = off + incl - d;
    if (n < frames) {
      const int64_t u = row[n];
      for (int k = 0; k < d && start + k < frames_out; ++k) orow[start + k] = u;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = off + incl;
    __syncthreads();
  }
  for (int pos = carry_s + threadIdx.x; pos < frames_out; pos += blockDim.x) orow[pos] = 0;
}

// ------------------------------------------------------------------------------------------ log-mel front end
// mel_spectrogram of src/hifigan/data.py:17-53 (the reference's feature extractor / validation metric):
// hann(400) windowed frames, hop 320, center=False -> 201-bin |DFT| -> 80 slaney mel bands -> log(max(., 1e-5)).
// One block = 16 frames of one waveform.  The DFT is direct (n_fft = 400 is not a power of two and a frame is tiny):
// thread <-> frequency bin keeps 16 frames x (re, im) in registers and walks the 400 samples; the twiddle for (bin b,
// sample k) is entry (b k mod 400) of ONE 400-entry cos / sin table in shared memory, advanced incrementally; the 16
// frames' samples are stored [sample][frame] so a step reads them as four broadcast 16-byte loads.
// FLOPs: 201 x 400 x 2 FMA per frame (+ 80 x 201 for the mel matrix); bytes: 400 x 4 in, 80 x 4 out per frame.
constexpr int kMelFft = 400, kMelHop = 320, kMelBins = 201, kMelBands = 80, kMelFrames = 16;

__global__ void __launch_bounds__(224) log_mel_kernel(const float* __restrict__ wav, long long wav_stride, int samples,
                                                      const float* __restrict__ window, const float* __restrict__ tw_cos,
                                                      const float* __restrict__ tw_sin, const float* __restrict__ mel_basis,
                                                      float* __restrict__ out, int frames) {
  __shared__ __align__(16) float xw[kMelFft][kMelFrames];     // windowed samples, [sample][frame]
  __shared__ float cs[kMelFft], sn[kMelFft];
  __shared__ float spec[kMelFrames][kMelBins];                 // 201 is odd: frame-strided reads are conflict free
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, f0 = blockIdx.x * kMelFrames, tid = threadIdx.x;
  const float* y = wav + (long long)b * wav_stride;
  for (int i = tid; i < kMelFft; i += blockDim.x) {
    cs[i] = tw_cos[i];
    sn[i] = tw_sin[i];
  }
  for (int i = tid; i < kMelFft * kMelFrames; i += blockDim.x) {
    const int f = i / kMelFft, k = i - f * kMelFft;             // consecutive threads: consecutive samples of a frame
    const long long pos = (long long)(f0 + f) * kMelHop + k;
    xw[k][f] = (f0 + f < frames && pos < samples) ? y[pos] * window[k] : 0.f;
  }
  __syncthreads();
  if (tid < kMelBins) {
    float re[kMelFrames], im[kMelFrames];
#pragma unroll
    for (int f = 0; f < kMelFrames; ++f) re[f] = im[f] = 0.f;
    int idx = 0;
#pragma unroll 2
    for (int k = 0; k < kMelFft; ++k) {
      const float c = cs[idx], s = sn[idx];
      const float4* xr = reinterpret_cast<const float4*>(&xw[k][0]);
#pragma unroll
      for (int q = 0; q < kMelFrames / 4; ++q) {
        const float4 x = xr[q];
        re[4 * q + 0] = fmaf(x.x, c, re[4 * q + 0]); im[4 * q + 0] = fmaf(-x.x, s, im[4 * q + 0]);
        re[4 * q + 1] = fmaf(x.y, c, re[4 * q + 1]); im[4 * q + 1] = fmaf(-x.y, s, im[4 * q + 1]);
        re[4 * q + 2] = fmaf(x.z, c, re[4 * q + 2]); im[4 * q + 2] = fmaf(-x.z, s, im[4 * q + 2]);
        re[4 * q + 3] = fmaf(x.w, c, re[4 * q + 3]); im[4 * q + 3] = fmaf(-x.w, s, im[4 * q + 3]);
      }
      idx += tid;
      if (idx >= kMelFft) idx -= kMelFft;
    }
#pragma unroll
    for (int f = 0; f < kMelFrames; ++f) spec[f][tid] = sqrtf(re[f] * re[f] + im[f] * im[f]);   // spec.abs()
  }
  __syncthreads();
  for (int o = tid; o < kMelBands * kMelFrames; o += blockDim.x) {
    const int m = o / kMelFrames, f = o - m * kMelFrames;
    if (f0 + f >= frames) continue;
    const float* wrow = mel_basis + m * kMelBins;
    float acc = 0.f;
    for (int k = 0; k < kMelBins; ++k) acc = fmaf(__ldg(wrow + k), spec[f][k], acc);
    out[((long long)b * kMelBands + m) * frames + f0 + f] = logf(fmaxf(acc, 1e-5f));   // dynamic_range_compression_torch
  }
}

// ------------------------------------------------------------------------------------------ positional conv
// x = gelu(dwconv31(mask(x0)) + b) * mask + x0, then the first AdaptiveRMSNorm (transformer.py:84-96, models.py:177,
// norm.py:41-43).  CUDA-core work, instruction-issue bound in its first form (one thread per channel, 159 instructions
// per output, ncu: 0.62 IPC, 66 us): this form spends ~40.
//   * a thread owns TWO adjacent channels and ROWS output rows: the 31-tap window (ROWS + 30 rows) sits in registers
//     as float2 and every multiply-add is a packed fma.rn.f32x2 (two FMAs per issue slot on sm_100);
//   * taps are the outer loop (one float2 weight load per tap and thread, ROWS independent accumulator chains);
//   * exact-GELU's erf is Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, one EX2 + one RCP + 5 FMAs) instead of erff;
//   * row sums of squares go through shared memory once per block (a warp reduces four rows) instead of five
//     shuffles per thread and row.
// (A persistent variant that staged the window in shared memory with cp.async double buffering measured 57.7 us against
// 41.7 us: its 100 KB of staging leave 8 warps per SM, and this kernel needs the 20 warps of the register form.)
// Bytes per row and channel: 4 (x0) + 4 (x) + 2 (xn) = 10 -> 82 MB at 64 x 504 frames.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
// erf(x), |abs error| <= 1.5e-7 (Abramowitz & Stegun 7.1.26 carried to fp32 with the approximate reciprocal and
// exponential units: 1 ulp each, far below the formula's own error).  The IEEE reciprocal (__frcp_rn) this used before
// is a software routine -- a CALL per output, ~40 % of the kernel's instructions together with the unfused tail.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// exact (erf) GELU: 0.5 a (1 + erf(a / sqrt 2))
__device__ __forceinline__ float gelu_exact(float a) {
  const float z = a * 0.70710678118654752440f;
  const float az = fabsf(z);
  const float t = rcp_approx(fmaf(0.3275911f, az, 1.f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = ex2_approx(az * az * -1.4426950408889634f);   // exp(-z^2)
  const float erf_abs = fmaf(-(poly * t), e, 1.f);
  const float ha = 0.5f * a;
  return fmaf(ha, copysignf(erf_abs, z), ha);
}

template <int ROWS>
__global__ void __launch_bounds__(128) posconv_norm_kernel(const float* __restrict__ x0, const float* __restrict__ dw_w,
                                                           const float* __restrict__ dw_b, const float* __restrict__ g,
                                                           const int* __restrict__ lengths, float* __restrict__ x,
                                                           __nv_bfloat16* __restrict__ xn, int frames) {
  constexpr int K = 31, HALO = 15;
  static_assert(ROWS % 4 == 0, "four warps share the row reduction");
  __shared__ float sq_s[ROWS][128];
  __shared__ float inv_s[ROWS];
  pdl_launch_dependents();
  const int b = blockIdx.y, t0 = blockIdx.x * ROWS, cp = threadIdx.x;   // cp: channel pair (channels 2cp, 2cp + 1)
  const float2 bias = __ldg(reinterpret_cast<const float2*>(dw_b) + cp);
  const float2 gc = __ldg(reinterpret_cast<const float2*>(g) + cp);
  pdl_wait();   // x0 and lengths come from the preceding kernels
  const int len = min(lengths[b], frames);
  const float2* xb = reinterpret_cast<const float2*>(x0 + (long long)b * frames * 256) + cp;
  float2 win[ROWS + 2 * HALO];
#pragma unroll
  for (int i = 0; i < ROWS + 2 * HALO; ++i) {
    const int t = t0 - HALO + i;
    win[i] = (t >= 0 && t < len) ? __ldg(xb + (long long)t * 128) : make_float2(0.f, 0.f);   // masked input
  }
  float2 acc[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) acc[r] = bias;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const float2 wj = __ldg(reinterpret_cast<const float2*>(dw_w + j * 256) + cp);   // tap-major [31][256]
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = ffma2(wj, win[r + j], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int t = t0 + r;
    float2 v;
    if (t < len) {
      v.x = gelu_exact(acc[r].x) + win[r + HALO].x;
      v.y = gelu_exact(acc[r].y) + win[r + HALO].y;
    } else {
      // pad row: the conv output is masked, the residual is the unmasked x0
      v = t < frames ? __ldg(xb + (long long)t * 128) : make_float2(0.f, 0.f);
    }
    acc[r] = v;
    sq_s[r][cp] = v.x * v.x + v.y * v.y;
  }
  __syncthreads();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int rr = 0; rr < ROWS / 4; ++rr) {
      const int r = warp * (ROWS / 4) + rr;
      float s = (sq_s[r][lane] + sq_s[r][lane + 32]) + (sq_s[r][lane + 64] + sq_s[r][lane + 96]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) inv_s[r] = 1.f / fmaxf(sqrtf(s), 1e-12f);   // F.normalize (norm.py:41)
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int t = t0 + r;
    if (t >= frames) break;
    const long long o = ((long long)b * frames + t) * 128 + cp;
    reinterpret_cast<float2*>(x)[o] = acc[r];
    const float inv = inv_s[r];
    const float n0 = acc[r].x * inv * gc.x, n1 = acc[r].y * inv * gc.y;
    const uint32_t hi = t < len ? pack_bf16(n0, n1) : 0u;
    if constexpr (kSplit == 1) {
      reinterpret_cast<uint32_t*>(xn)[o] = hi;
    } else {
      uint32_t* d = reinterpret_cast<uint32_t*>(xn) + ((long long)b * frames + t) * (128 * kSplit) + cp;
      d[0] = hi;
      d[128] = t < len ? pack_bf16_rest(n0, n1, hi) : 0u;
      d[256] = hi;
    }
  }
}

// ------------------------------------------------------------------------------------------ conv_post + tanh
// wav[b, t] = tanh(bias + sum_{j<7, c<16} w[j][c] * x[b, t + j - 3, c]);  x already leaky_relu(0.01)'ed.  HF:1480-1482
// 128 threads x 2 consecutive outputs (256 outputs per block): the 8 input rows of an output pair are read once (16
// conflict-free 16-byte shared loads per pair instead of 28 two-way conflicting ones for two single outputs -- the
// one-output form was shared-memory-bandwidth bound at 232 us per 64 x 160 080 samples).  Row r of the window sits in
// slot r + r / 8 so that the stride-2 row pattern of a quarter warp covers eight different 16-byte bank groups.
constexpr int kPostOutputs = 256;
__device__ __forceinline__ int post_slot(int r) { return r + (r >> 3); }
// Ragged form (lengths != NULL): utterance b keeps its first n_b = min(rows, 320 * lengths[b] + 80) samples
// (_get_waveform_lengths, models.py:211-221) and they are stored back to back: wav[sum_{i<b} n_i + t] -- the reference's
// per-utterance crop loop (models.py:252-256) folded into the store.  Every block sums the (<= a few thousand) preceding
// lengths itself, so no offset array has to travel to the device.
__global__ void __launch_bounds__(128) post_tanh_kernel(const uint4* __restrict__ x, const float* __restrict__ w, float bias,
                                                        float* __restrict__ wav, int rows, const int* __restrict__ lengths) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int kRows = kPostOutputs + 6;
  constexpr int kSlots = kRows + kRows / 8 + 1;
  __shared__ uint4 tile[2][kSlots];
  __shared__ uint4 tile_lo[kSplit == 3 ? 2 : 1][kSplit == 3 ? kSlots : 1];   // split build only
  __shared__ __align__(16) float ws[7 * 16];
  __shared__ long long off_s[4];
  const int b = blockIdx.y, t0 = blockIdx.x * kPostOutputs;
  if (threadIdx.x < 112) ws[threadIdx.x] = w[threadIdx.x];
  int n_out = rows;
  if (lengths != nullptr) {
    n_out = min(rows, 320 * lengths[b] + 80);
    if (t0 >= n_out) return;   // block-uniform: the whole block is beyond the utterance's end
    long long part = 0;
    for (int i = threadIdx.x; i < b; i += 128) part += min(rows, 320 * lengths[i] + 80);
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) off_s[threadIdx.x >> 5] = part;
  }
  const uint4* xb = x + (long long)b * rows * 2 * kSplit;
  for (int i = threadIdx.x; i < kRows * 2; i += 128) {
    const int r = i >> 1, t = t0 - 3 + r;
    tile[i & 1][post_slot(r)] = (t >= 0 && t < rows) ? __ldg(xb + (long long)t * 2 * kSplit + (i & 1)) : make_uint4(0, 0, 0, 0);
    if constexpr (kSplit == 3)   // the rounding rests, 16 columns (two 16-byte pieces) further on
      tile_lo[i & 1][post_slot(r)] = (t >= 0 && t < rows) ? __ldg(xb + (long long)t * 2 * kSplit + 2 + (i & 1)) : make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  const int r0 = 2 * threadIdx.x;        // window row of tap 0 of the first output
  const int t = t0 + r0;
  if (t >= n_out) return;
  // channel pairs as packed fp32 FMAs: 8 fma.rn.f32x2 per (row, output) instead of 16 scalar ones; the weights of tap
  // k serve output 0 now and output 1 at the next row
  float2 a0 = make_float2(bias, 0.f), a1 = make_float2(bias, 0.f);
  float2 wprev[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint4 u0 = tile[0][post_slot(r0 + k)], u1 = tile[1][post_slot(r0 + k)];
    float2 f[8];
    f[0] = make_float2(bf16_lo(u0.x), bf16_hi(u0.x)); f[1] = make_float2(bf16_lo(u0.y), bf16_hi(u0.y));
    f[2] = make_float2(bf16_lo(u0.z), bf16_hi(u0.z)); f[3] = make_float2(bf16_lo(u0.w), bf16_hi(u0.w));
    f[4] = make_float2(bf16_lo(u1.x), bf16_hi(u1.x)); f[5] = make_float2(bf16_lo(u1.y), bf16_hi(u1.y));
    f[6] = make_float2(bf16_lo(u1.z), bf16_hi(u1.z)); f[7] = make_float2(bf16_lo(u1.w), bf16_hi(u1.w));
    if constexpr (kSplit == 3) {
      const uint4 l0 = tile_lo[0][post_slot(r0 + k)], l1 = tile_lo[1][post_slot(r0 + k)];
      f[0].x += bf16_lo(l0.x); f[0].y += bf16_hi(l0.x); f[1].x += bf16_lo(l0.y); f[1].y += bf16_hi(l0.y);
      f[2].x += bf16_lo(l0.z); f[2].y += bf16_hi(l0.z); f[3].x += bf16_lo(l0.w); f[3].y += bf16_hi(l0.w);
      f[4].x += bf16_lo(l1.x); f[4].y += bf16_hi(l1.x); f[5].x += bf16_lo(l1.y); f[5].y += bf16_hi(l1.y);
      f[6].x += bf16_lo(l1.z); f[6].y += bf16_hi(l1.z); f[7].x += bf16_lo(l1.w); f[7].y += bf16_hi(l1.w);
    }
    if (k > 0) {
#pragma unroll
      for (int c = 0; c < 8; ++c) a1 = ffma2(f[c], wprev[c], a1);
    }
    if (k < 7) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        wprev[c] = *reinterpret_cast<const float2*>(ws + k * 16 + 2 * c);
        a0 = ffma2(f[c], wprev[c], a0);
      }
    }
  }
  const float acc0 = a0.x + a0.y, acc1 = a1.x + a1.y;
  float* o = wav + (lengths != nullptr ? (off_s[0] + off_s[1]) + (off_s[2] + off_s[3]) : (long long)b * rows) + t;
  const float y0 = tanhf(acc0), y1 = tanhf(acc1);
  if (t + 1 < n_out && (reinterpret_cast<uintptr_t>(o) & 7) == 0) {
    *reinterpret_cast<float2*>(o) = make_float2(y0, y1);
  } else {
    o[0] = y0;
    if (t + 1 < n_out) o[1] = y1;
  }
}

// ------------------------------------------------------------------------------------------ reference-style ops
// fp32 CUDA-core forms used by the tight-precision build (and by tests): no tensor cores, no approximations beyond
// expf.  They read / write bf16 tensors in the build's activation format (plain bf16, or [hi | lo | hi] split).
__device__ __forceinline__ float load_act(const __nv_bfloat16* row, int logical_width, int c) {
  float v = __bfloat162float(row[c]);
  if constexpr (kSplit == 3) v += __bfloat162float(row[logical_width + c]);
  return v;
}
__device__ __forceinline__ void store_act(__nv_bfloat16* row, int logical_width, int c, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  row[c] = h;
  if constexpr (kSplit == 3) {
    row[logical_width + c] = __float2bfloat16_rn(v - __bfloat162float(h));
    row[2 * logical_width + c] = h;
  }
}

// Key-padding-masked softmax attention, 2 heads x 128 (transformer.py:115-127), online softmax in fp32.
// qkv (B, N, 768) = [q | k | v] after rotary, o (B, N, 256).  Block = 4 warps x 4 query rows of one (utterance, head);
// a lane owns one key of the current 32-key tile for the scores and four output dimensions for P V.
__global__ void __launch_bounds__(128) attention_simt_kernel(const __nv_bfloat16* __restrict__ qkv, const int* __restrict__ lengths,
                                                             __nv_bfloat16* __restrict__ o, int frames) {
  constexpr int D = 128, KT = 32, QR = 16;
  __shared__ float ks[KT][D + 1];
  __shared__ float vs[KT][D + 1];
  __shared__ float qs[QR][D];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z, h = blockIdx.y, t0 = blockIdx.x * QR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = min(lengths[b], frames);
  const long long pitch = 768 * kSplit;
  const __nv_bfloat16* base = qkv + (long long)b * frames * pitch;
  for (int i = threadIdx.x; i < QR * D; i += 128) {
    const int r = i / D, d = i % D, t = t0 + r;
    qs[r][d] = t < frames ? load_act(base + (long long)t * pitch, 768, h * D + d) : 0.f;
  }
  float m[4], l[4], acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;
  }
  const float scale = 0.08838834764831845f;   // 1 / sqrt(128)
  for (int j0 = 0; j0 < len; j0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * D; i += 128) {
      const int j = i / D, d = i % D, t = j0 + j;
      const bool ok = t < len;
      ks[j][d] = ok ? load_act(base + (long long)t * pitch, 768, 256 + h * D + d) : 0.f;
      vs[j][d] = ok ? load_act(base + (long long)t * pitch, 768, 512 + h * D + d) : 0.f;
    }
    __syncthreads();
    const bool key_ok = j0 + lane < len;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float* q = qs[warp * 4 + r];
      float sc = 0.f;
#pragma unroll 8
      for (int d = 0; d < D; ++d) sc = fmaf(q[d], ks[lane][d], sc);
      sc = key_ok ? sc * scale : -INFINITY;
      float mx = sc;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m[r], mx);        // finite: every tile holds at least one valid key
      const float corr = expf(m[r] - m_new);      // exp(-inf) = 0 on the first tile
      const float pr = key_ok ? expf(sc - m_new) : 0.f;
      float ps = pr;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l[r] = l[r] * corr + ps;
      m[r] = m_new;
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[r][i] *= corr;
      for (int j = 0; j < KT; ++j) {
        const float pj = __shfl_sync(0xffffffffu, pr, j);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[r][i] = fmaf(pj, vs[j][lane + 32 * i], acc[r][i]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int t = t0 + warp * 4 + r;
    if (t >= frames) continue;
    __nv_bfloat16* orow = o + ((long long)b * frames + t) * (256 * kSplit);
    const float inv = 1.f / l[r];
#pragma unroll
    for (int i = 0; i < 4; ++i) store_act(orow, 256, h * D + lane + 32 * i, acc[r][i] * inv);
  }
}

// out = leaky_relu((x0 + x1 + x2) * scale, slope): the MRF mean (HF:1475-1480) when the fused tail is not used
__global__ void __launch_bounds__(256) mean3_act_kernel(const __nv_bfloat16* __restrict__ x0, const __nv_bfloat16* __restrict__ x1,
                                                        const __nv_bfloat16* __restrict__ x2, __nv_bfloat16* __restrict__ out,
                                                        long long rows, int c, float scale, float slope) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n = rows * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / c;
    const int col = (int)(i - row * c);
    const long long off = row * c * kSplit;
    const float v = (load_act(x0 + off, c, col) + load_act(x1 + off, c, col) + load_act(x2 + off, c, col)) * scale;
    store_act(out + off, c, col, v > 0.f ? v : v * slope);
  }
}


// ------------------------------------------------------------------------------------------ unit quantiser (input side)
// Split copy of fp32 rows for the tight GEMM form: out row = [bf16(x) | bf16(x - bf16(x)) | bf16(x)], 3 * width bf16
// (include/srb.h: srb_split_factor); also clears the per-row argmax keys of srb_kmeans_assign.
__global__ void __launch_bounds__(256) split_rows_kernel(const float4* __restrict__ x, uint2* __restrict__ out, long long rows,
                                                         int w4, unsigned long long* __restrict__ keys) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n = rows * w4, stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = tid; i < n; i += stride) {
    const long long row = i / w4;
    const int piece = (int)(i - row * w4);
    const float4 v = __ldg(x + i);
    const uint2 h = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    uint2* d = out + row * (3ll * w4) + piece;
    d[0] = h;
    d[w4] = make_uint2(pack_bf16_rest(v.x, v.y, h.x), pack_bf16_rest(v.z, v.w, h.y));
    d[2 * w4] = h;
  }
  if (keys != nullptr)
    for (long long i = tid; i < rows; i += stride) keys[i] = 0ull;
}

// keys[row] = (ordered score bits << 32) | (0xFFFFFFFF - column)  ->  units[row] = column + id_offset
// (rows = batch * frames with per-utterance lengths: positions at or beyond lengths[b] become 0, the pad id)
__global__ void __launch_bounds__(256) kmeans_decode_kernel(const unsigned long long* __restrict__ keys, int64_t* __restrict__ units,
                                                            long long rows, int id_offset, const int* __restrict__ lengths, int frames) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
    const bool pad = lengths != nullptr && (int)(i % frames) >= lengths[i / frames];
    units[i] = pad ? 0 : (int64_t)(0xFFFFFFFFu - (uint32_t)(keys[i] & 0xFFFFFFFFull)) + id_offset;
  }
}

// torch.unique_consecutive(return_counts=True) per utterance (textless SpeechEncoder's `deduplicate`): runs of equal ids
// within the first lengths[b] entries of a row collapse to one id + its run length; outputs are right-padded with 0.
// One block per utterance: chunked scan of the "starts a run" flags (same scheme as length_regulate_kernel).
__global__ void __launch_bounds__(256) unique_consecutive_kernel(const int64_t* __restrict__ ids, const int* __restrict__ lengths,
                                                                 int64_t* __restrict__ out_ids, int* __restrict__ out_counts,
                                                                 int* __restrict__ out_len, int frames) {
  __shared__ int warp_tot[8];
  __shared__ int carry_s;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t* row = ids + (long long)b * frames;
  int64_t* orow = out_ids + (long long)b * frames;
  int* crow = out_counts + (long long)b * frames;
  const int len = lengths != nullptr ? min(lengths[b], frames) : frames;
  for (int n = threadIdx.x; n < frames; n += blockDim.x) {
    orow[n] = 0;
    crow[n] = 0;
  }
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < len; base += blockDim.x) {
    const int n = base + threadIdx.x;
    const int start = (n < len && (n == 0 || row[n] != row[n - 1])) ? 1 : 0;
    int incl = start;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int off = carry_s;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (start) {
      // this thread opens run number off + incl - 1: its length is the distance to the next run start (or to len)
      int e = n + 1;
      const int64_t u = row[n];
      while (e < len && row[e] == u) ++e;
      orow[off + incl - 1] = u;
      crow[off + incl - 1] = e - n;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) out_len[b] = carry_s;
}

}  // namespace srb

using namespace srb;

extern "C" {

int srb_embed_gather(const float* table, const int64_t* ids, float* out, int64_t m, int32_t vocab_rows, int32_t dim,
                     void* stream) {
  SRB_REQUIRE(dim % 4 == 0, "srb_embed_gather: dim must be a multiple of 4");
  SRB_REQUIRE(((uintptr_t)table & 15) == 0 && ((uintptr_t)out & 15) == 0, "srb_embed_gather: pointers must be 16-byte aligned");
  if (m <= 0) return 0;
  long long blocks = (m + 7) / 8;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  SRB_CUDA(launch_pdl(embed_gather_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(table), ids, reinterpret_cast<float4*>(out), (long long)m, (int)vocab_rows, (int)(dim / 4)));
  return after_launch("embed_gather_kernel");
}

int srb_unit_lengths(const int64_t* ids, int32_t* lengths, int32_t batch, int32_t frames, void* stream) {
  if (batch <= 0) return 0;
  SRB_CUDA(launch_pdl(unit_lengths_kernel, dim3(batch), dim3(256), 0, (cudaStream_t)stream, ids, lengths, (int*)nullptr, frames));
  return after_launch("unit_lengths_kernel");
}

int srb_unit_extents(const int64_t* ids, int32_t* lengths, int32_t* extents, int32_t batch, int32_t frames, void* stream) {
  if (batch <= 0) return 0;
  SRB_CUDA(launch_pdl(unit_lengths_kernel, dim3(batch), dim3(256), 0, (cudaStream_t)stream, ids, lengths, extents, frames));
  return after_launch("unit_lengths_kernel");
}

int srb_time_cond_table(const float* times, int32_t nfe, const float* four_w, const float* lin_w, const float* lin_b,
                        const float* gamma_w, int32_t n_norm, float* time_emb, float* g, void* stream) {
  if (nfe <= 0) return 0;
  time_cond_kernel<<<nfe, 256, 0, (cudaStream_t)stream>>>(times, four_w, lin_w, lin_b, gamma_w, n_norm, time_emb, g);
  return after_launch("time_cond_kernel");
}

int srb_rotary_table(const float* inv_freq, int32_t rows, float* cos_out, float* sin_out, void* stream) {
  if (rows <= 0) return 0;
  const int n = rows * 64;
  rotary_table_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(inv_freq, rows, cos_out, sin_out);
  return after_launch("rotary_table_kernel");
}

int srb_stage_inputs(const int64_t* ids_in, const float* prior_in, int64_t* ids, float* xt, void* xt_bf16, void* zero_a,
                     int64_t zero_a_bytes, void* zero_b, int64_t zero_b_bytes, int32_t batch, int32_t frames_in,
                     int32_t frames, float truncation, int32_t has_truncation, void* stream) {
  SRB_REQUIRE(frames >= frames_in && frames_in > 0 && batch > 0, "srb_stage_inputs: need 0 < frames_in <= frames and batch > 0");
  SRB_REQUIRE(((uintptr_t)prior_in & 15) == 0 && ((uintptr_t)xt & 15) == 0 && ((uintptr_t)xt_bf16 & 7) == 0,
              "srb_stage_inputs: prior / state pointers must be 16-byte aligned");
  SRB_REQUIRE(zero_a_bytes % 16 == 0 && zero_b_bytes % 16 == 0 && ((uintptr_t)zero_a & 15) == 0 && ((uintptr_t)zero_b & 15) == 0,
              "srb_stage_inputs: cleared regions must be 16-byte aligned multiples of 16 bytes");
  long long blocks = ((long long)batch * frames * 20 + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  SRB_CUDA(launch_pdl(stage_inputs_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, ids_in,
                      reinterpret_cast<const float4*>(prior_in), ids, reinterpret_cast<float4*>(xt),
                      reinterpret_cast<uint2*>(xt_bf16), static_cast<uint4*>(zero_a), (long long)(zero_a ? zero_a_bytes / 16 : 0),
                      static_cast<uint4*>(zero_b), (long long)(zero_b ? zero_b_bytes / 16 : 0), (int)batch, (int)frames_in,
                      (int)frames, truncation, (int)has_truncation));
  return after_launch("stage_inputs_kernel");
}

int srb_prior_prepare(float* xt, void* xt_bf16, int64_t n, float truncation, int32_t has_truncation, void* stream) {
  SRB_REQUIRE(n % 4 == 0, "srb_prior_prepare: element count must be a multiple of 4");
  if (n <= 0) return 0;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  SRB_CUDA(launch_pdl(prior_prepare_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
                      reinterpret_cast<float4*>(xt), reinterpret_cast<uint2*>(xt_bf16), (long long)(n / 4), truncation,
                      (int)has_truncation));
  return after_launch("prior_prepare_kernel");
}

int srb_duration_predict(const int64_t* ids, const float* dur_table, float bias, int32_t* durations, int32_t* totals,
                         int32_t batch, int32_t frames, int32_t vocab_rows, void* stream) {
  if (batch <= 0 || frames <= 0) return 0;
  SRB_CUDA(launch_pdl(duration_predict_kernel, dim3(batch), dim3(256), 0, (cudaStream_t)stream, ids, dur_table, bias, durations,
                      totals, frames, vocab_rows));
  return after_launch("duration_predict_kernel");
}

int srb_length_regulate(const int64_t* ids, const int32_t* durations, int64_t* out_ids, int32_t batch, int32_t frames,
                        int32_t frames_out, int32_t all_one, void* stream) {
  if (batch <= 0 || frames <= 0 || frames_out <= 0) return 0;
  SRB_CUDA(launch_pdl(length_regulate_kernel, dim3(batch), dim3(256), 0, (cudaStream_t)stream, ids, durations, out_ids, frames,
                      frames_out, all_one));
  return after_launch("length_regulate_kernel");
}

int srb_log_mel(const float* wav, int64_t wav_stride, int32_t batch, int32_t samples, const float* window, const float* tw_cos,
                const float* tw_sin, const float* mel_basis, float* out, int32_t frames, void* stream) {
  if (batch <= 0 || frames <= 0) return 0;
  SRB_REQUIRE(samples >= 400 && frames <= 1 + (samples - 400) / 320, "srb_log_mel: frames exceed 1 + (samples - 400) / 320");
  dim3 grid((frames + kMelFrames - 1) / kMelFrames, batch);
  SRB_CUDA(launch_pdl(log_mel_kernel, grid, dim3(224), 0, (cudaStream_t)stream, wav, (long long)wav_stride, (int)samples, window,
                      tw_cos, tw_sin, mel_basis, out, (int)frames));
  return after_launch("log_mel_kernel");
}

int srb_cfm_posconv_norm(const float* x0, const float* dw_w, const float* dw_b, const float* g, const int32_t* lengths,
                         float* x, void* xn_bf16, int32_t batch, int32_t frames, void* stream) {
  if (batch <= 0 || frames <= 0) return 0;
  // (A "marching" form -- one block sliding the register window down 16..128 rows, so x0 is read once instead of 4.75
  // times -- measured 47.7-50.3 us against 46.4 us for this one at 64 x 504 with a flushed L2: the launch is bound by
  // its instruction count (GELU), not by the window reloads.)
  dim3 grid((frames + 7) / 8, batch);
  SRB_CUDA(launch_pdl(posconv_norm_kernel<8>, grid, dim3(128), 0, (cudaStream_t)stream, x0, dw_w, dw_b, g, lengths, x,
                      static_cast<__nv_bfloat16*>(xn_bf16), frames));
  return after_launch("posconv_norm_kernel");
}

int srb_hifigan_post(const void* x_act, const float* w, float bias, float* wav, int32_t batch, int32_t rows,
                     const int32_t* lengths, void* stream) {
  if (batch <= 0 || rows <= 0) return 0;
  dim3 grid((rows + kPostOutputs - 1) / kPostOutputs, batch);
  SRB_CUDA(launch_pdl(post_tanh_kernel, grid, dim3(128), 0, (cudaStream_t)stream, static_cast<const uint4*>(x_act), w, bias, wav, rows,
                      lengths));
  return after_launch("post_tanh_kernel");
}

int srb_cfm_attention_simt(const void* qkv_bf16, const int32_t* lengths, void* o_bf16, int32_t batch, int32_t frames, void* stream) {
  if (batch <= 0 || frames <= 0) return 0;
  SRB_REQUIRE(lengths != nullptr, "srb_cfm_attention_simt: lengths required");
  dim3 grid((frames + 15) / 16, 2, batch);
  SRB_CUDA(launch_pdl(attention_simt_kernel, grid, dim3(128), 0, (cudaStream_t)stream, static_cast<const __nv_bfloat16*>(qkv_bf16),
                      lengths, static_cast<__nv_bfloat16*>(o_bf16), frames));
  return after_launch("attention_simt_kernel");
}

int srb_hifigan_mean3(const void* x0, const void* x1, const void* x2, void* out, int64_t rows, int32_t channels, float scale,
                      float slope, void* stream) {
  if (rows <= 0 || channels <= 0) return 0;
  long long blocks = (rows * channels + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  SRB_CUDA(launch_pdl(mean3_act_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
                      static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1),
                      static_cast<const __nv_bfloat16*>(x2), static_cast<__nv_bfloat16*>(out), (long long)rows, channels, scale, slope));
  return after_launch("mean3_act_kernel");
}

int srb_split_bf16(const float* x, void* out_bf16, int64_t rows, int32_t width, uint64_t* keys_to_clear, void* stream) {
  if (rows <= 0 || width <= 0) return 0;
  SRB_REQUIRE(width % 4 == 0, "srb_split_bf16: width must be a multiple of 4");
  long long blocks = (rows * (width / 4) + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  SRB_CUDA(launch_pdl(split_rows_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
                      reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(out_bf16), (long long)rows, (int)(width / 4),
                      reinterpret_cast<unsigned long long*>(keys_to_clear)));
  return after_launch("split_rows_kernel");
}

int srb_kmeans_decode(const uint64_t* keys, int64_t* units, int64_t rows, int32_t id_offset, const int32_t* lengths,
                      int32_t frames, void* stream) {
  if (rows <= 0) return 0;
  SRB_REQUIRE(lengths == nullptr || (frames > 0 && rows % frames == 0), "srb_kmeans_decode: rows must be batch * frames when lengths are given");
  long long blocks = (rows + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  SRB_CUDA(launch_pdl(kmeans_decode_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
                      reinterpret_cast<const unsigned long long*>(keys), units, (long long)rows, (int)id_offset, lengths,
                      (int)(lengths != nullptr ? frames : 1)));
  return after_launch("kmeans_decode_kernel");
}

int srb_unique_consecutive(const int64_t* ids, const int32_t* lengths, int64_t* out_ids, int32_t* out_counts, int32_t* out_lengths,
                           int32_t batch, int32_t frames, void* stream) {
  if (batch <= 0 || frames <= 0) return 0;
  SRB_CUDA(launch_pdl(unique_consecutive_kernel, dim3(batch), dim3(256), 0, (cudaStream_t)stream, ids, lengths, out_ids, out_counts,
                      out_lengths, (int)frames));
  return after_launch("unique_consecutive_kernel");
}

}  // extern "C"
