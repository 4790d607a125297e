// fft_mixer.cu — the global branch of LGMixer (models/common/LGT.py:149-180), a Fourier amplitude/phase
// mixer on the second half of the channels:
//     fre = rfft2(x); amp = |fre|; pha = angle(fre)                               LGT.py:166-169
//     amp' = conv_amp(amp), pha' = conv_pha(pha)     (depthwise 1x1 = per-channel affine)   LGT.py:171-172
//     out = | irfft2( complex(amp' cos pha' + 1e-8, amp' sin pha' + 1e-8) + 1e-8 ) |        LGT.py:174-178
// followed (when PROJ) by the rest of LGMixer + residual: proj(cat(local, global)) + x      LGT.py:212-217, :51
//
// True-fp32 FFTs written here (no cuFFT): the phase branch cut makes the result ill-conditioned in the
// spectrum, so twiddles are rounded from double and the four purely-real bins
// (ky in {0,H/2}, kx in {0,W/2}) get an exact +0.0 imaginary part like the CPU oracle's rfft2.
//
// Three HBM-bound passes per block (spectrum layout S[n][y][kx][ch] complex64, ch innermost):
//   rows_fwd : LN (optional) -> two real channels packed into one complex Stockham FFT along W -> split -> S
//   cols     : in-place radix-4 DIF FFT along H in shared memory (digit-reversed order), pointwise
//              amp/phase mixing, radix-4 DIT inverse (consumes digit-reversed order, no permutation pass)
//   rows_inv : Hermitian rebuild (C2R ignores Im of bins 0 and W/2) -> inverse Stockham -> |.|/(HW)
//              -> concat with the local branch -> proj 1x1 -> + residual
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace lg {

constexpr int kTwN = 1024;                               // largest supported FFT length
__device__ float2 g_tw[kTwN];                            // e^{-2 pi i k / 1024}, rounded from double

cudaError_t fft_init_tables(cudaStream_t s) {
  static float2 host[kTwN];
  for (int k = 0; k < kTwN; ++k) {
    double a = -2.0 * M_PI * (double)k / (double)kTwN;
    host[k] = make_float2((float)cos(a), (float)sin(a));
  }
  // exact values on the axes / diagonals
  host[0] = make_float2(1.f, 0.f);
  host[kTwN / 4] = make_float2(0.f, -1.f);
  host[kTwN / 2] = make_float2(-1.f, 0.f);
  host[3 * kTwN / 4] = make_float2(0.f, 1.f);
  cudaError_t e = cudaMemcpyToSymbolAsync(g_tw, host, sizeof(host), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
  return fft256_init_tables(s);
}

// tw[j] = e^{-2 pi i j / L} for the shared-memory passes: from the rounded-from-double table when L divides 1024 (every
// power of two), else computed in double on the spot (lengths with odd factors: any H, W that are multiples of 8)
__device__ __forceinline__ float2 tw_entry(int j, int L) {
  if (kTwN % L == 0) return g_tw[j * (kTwN / L)];
  double sn, cs;
  sincospi(-2.0 * (double)j / (double)L, &sn, &cs);
  return make_float2((float)cs, (float)sn);
}

size_t spectrum_floats(int N, int H, int W, int c2) { return (size_t)N * H * (W / 2 + 1) * c2 * 2; }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {     // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// ---- natural-order Stockham FFT of NF sequences of length L held in shared memory ----------------------
// a, b: ping-pong buffers [NF][LS], LS = L + 2 (kSeqPad): the pad keeps 128-bit alignment and moves consecutive sequences
// four banks apart, so the per-frequency gathers across sequences (pack / split phases) are conflict-free.
// tw[j] = e^{-2 pi i j / L}.  SIGN = -1 forward, +1 inverse (unnormalised).
// Returns the buffer holding the result.  All threads of the CTA must call it.
constexpr int kSeqPad = 2;
// ANY = false: L is a power of two (radix-4 / radix-2 stages only; the hot instantiations);  ANY = true: also lengths with odd
// factors (own kernel instantiations, so the generic stage's local array costs the power-of-two kernels nothing)
template <int SIGN, bool ANY = false>
__device__ float2* stockham(float2* a, float2* b, const float2* tw, int L, int NF, int tid, int nthreads) {
  const int LS = L + kSeqPad;
  for (int Ns = 1; Ns < L;) {
    const int rem = L / Ns;
    if ((rem & 3) == 0) {
      const int q = L >> 2;                        // items per sequence
      const int tws = L / (Ns * 4);                // twiddle stride: w_{4Ns}^k = tw[k * tws]
      for (int id = tid; id < NF * q; id += nthreads) {
        const int f = id / q, j = id - f * q;
        const int k = j & (Ns - 1);
        const float2* src = a + f * LS + j;
        float2 v0 = src[0], v1 = src[q], v2 = src[2 * q], v3 = src[3 * q];
        if (k) {
          float2 w1 = tw[k * tws], w2 = tw[2 * k * tws], w3 = tw[3 * k * tws];
          if (SIGN < 0) { v1 = cmul(v1, w1); v2 = cmul(v2, w2); v3 = cmul(v3, w3); }
          else          { v1 = cmul_conj(v1, w1); v2 = cmul_conj(v2, w2); v3 = cmul_conj(v3, w3); }
        }
        // radix-4 butterfly; forward: o1 = v0 - i v1 - v2 + i v3, inverse: o1 = v0 + i v1 - v2 - i v3
        float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
        float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
        float2 jd = (SIGN < 0) ? make_float2(d13.y, -d13.x) : make_float2(-d13.y, d13.x);   // (-/+ i) * d13
        float2* dst = b + f * LS + (j - k) * 4 + k;
        dst[0] = make_float2(s02.x + s13.x, s02.y + s13.y);
        dst[Ns] = make_float2(d02.x + jd.x, d02.y + jd.y);
        dst[2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
        dst[3 * Ns] = make_float2(d02.x - jd.x, d02.y - jd.y);
      }
      Ns <<= 2;
    } else if (ANY && (rem & 1)) {
      // odd factor (lengths that are not powers of two): one generic radix-R stage, R = smallest odd prime factor of rem;
      // the R-point DFT is evaluated directly (R <= 61 for lengths <= 1024 that are multiples of 16)
      int R = 3;
      while (rem % R) R += 2;
      const int q = L / R;
      const int tws = L / (Ns * R);                // twiddle stride: w_{R Ns}^{k r} = tw[k r tws]
      for (int id = tid; id < NF * q; id += nthreads) {
        const int f = id / q, j = id - f * q;
        const int k = j % Ns;
        const float2* src = a + f * LS + j;
        float2 v[64];
        for (int r = 0; r < R; ++r) {
          float2 t = src[r * q];
          if (k && r) {
            const float2 w1 = tw[k * r * tws];
            t = (SIGN < 0) ? cmul(t, w1) : cmul_conj(t, w1);
          }
          v[r] = t;
        }
        float2* dst = b + f * LS + (j - k) * R + k;
        for (int o = 0; o < R; ++o) {
          float2 acc = v[0];
          int e = 0;                               // (r * o) mod R
          for (int r = 1; r < R; ++r) {
            e += o;
            if (e >= R) e -= R;
            const float2 wr = tw[e * q];
            const float2 t = (SIGN < 0) ? cmul(v[r], wr) : cmul_conj(v[r], wr);
            acc.x += t.x;
            acc.y += t.y;
          }
          dst[o * Ns] = acc;
        }
      }
      Ns *= R;
    } else {
      const int q = L >> 1;
      const int tws = L / (Ns * 2);
      for (int id = tid; id < NF * q; id += nthreads) {
        const int f = id / q, j = id - f * q;
        const int k = j & (Ns - 1);
        const float2* src = a + f * LS + j;
        float2 v0 = src[0], v1 = src[q];
        if (k) {
          float2 w1 = tw[k * tws];
          v1 = (SIGN < 0) ? cmul(v1, w1) : cmul_conj(v1, w1);
        }
        float2* dst = b + f * LS + (j - k) * 2 + k;
        dst[0] = make_float2(v0.x + v1.x, v0.y + v1.y);
        dst[Ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
      }
      Ns <<= 1;
    }
    __syncthreads();
    float2* t = a; a = b; b = t;
  }
  return a;
}

// ---- compile-time specialised Stockham (sizes used by the headline shapes): strides and trip counts are constants,
// butterflies run on the packed fp32 pipe (one FADD2 / FFMA2 per complex add / half complex multiply) -------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
template <int SIGN>
__device__ __forceinline__ float2 ctw(float2 a, float2 w) {            // a * w (SIGN < 0) or a * conj(w) (SIGN > 0)
  const float2 t = __fmul2_rn(a, make_float2(w.x, w.x));
  return (SIGN < 0) ? __ffma2_rn(make_float2(-a.y, a.x), make_float2(w.y, w.y), t)
                    : __ffma2_rn(make_float2(a.y, -a.x), make_float2(w.y, w.y), t);
}

template <int L, int NF, int SIGN, int NT, int Ns>
__device__ __forceinline__ float2* stockham_ct(float2* a, float2* b, const float2* tw, int tid) {
  if constexpr (Ns >= L) {
    return a;
  } else {
    constexpr int R = (((L / Ns) & 3) == 0) ? 4 : 2;
    constexpr int q = L / R;                       // butterflies per sequence
    constexpr int tws = L / (Ns * R);
    constexpr int ITEMS = NF * q;
    constexpr int LS = L + kSeqPad;
#pragma unroll
    for (int base = 0; base < ITEMS; base += NT) {
      const int id = base + tid;
      if (ITEMS % NT == 0 || id < ITEMS) {
        const int f = id / q, j = id % q;
        const int k = j & (Ns - 1);
        const float2* src = a + f * LS + j;
        float2* dst = b + f * LS + (j - k) * R + k;
        if constexpr (R == 4) {
          float2 v0 = src[0], v1 = src[q], v2 = src[2 * q], v3 = src[3 * q];
          if constexpr (Ns > 1) {
            v1 = ctw<SIGN>(v1, tw[k * tws]);
            v2 = ctw<SIGN>(v2, tw[2 * k * tws]);
            v3 = ctw<SIGN>(v3, tw[3 * k * tws]);
          }
          const float2 s02 = cadd(v0, v2), d02 = csub(v0, v2), s13 = cadd(v1, v3), d13 = csub(v1, v3);
          const float2 jd = (SIGN < 0) ? make_float2(d13.y, -d13.x) : make_float2(-d13.y, d13.x);
          const float2 o0 = cadd(s02, s13), o1 = cadd(d02, jd), o2 = csub(s02, s13), o3 = csub(d02, jd);
          if constexpr (Ns == 1) {                   // the four outputs are contiguous: two conflict-free 128-bit stores
            reinterpret_cast<float4*>(dst)[0] = make_float4(o0.x, o0.y, o1.x, o1.y);
            reinterpret_cast<float4*>(dst)[1] = make_float4(o2.x, o2.y, o3.x, o3.y);
          } else {
            dst[0] = o0; dst[Ns] = o1; dst[2 * Ns] = o2; dst[3 * Ns] = o3;
          }
        } else {
          float2 v0 = src[0], v1 = src[q];
          if constexpr (Ns > 1) v1 = ctw<SIGN>(v1, tw[k * tws]);
          dst[0] = cadd(v0, v1);
          dst[Ns] = csub(v0, v1);
        }
      }
    }
    __syncthreads();
    return stockham_ct<L, NF, SIGN, NT, Ns * R>(b, a, tw, tid);
  }
}

constexpr int kFftThreads = 256;

// ---- pass 1: rows forward ----------------------------------------------------------------------------
// NT threads per CTA: 256 on the compile-time fast paths; long rows (W >= 512) hold one row of 64-128 KB per CTA, i.e.
// one CTA per SM, and use 512 / 1024 threads so that the SM still has 16 / 32 resident warps
template <int C2, bool PRE_LN, int WCT, int ROWS, int NT = kFftThreads, bool ANY = false>
__global__ void __launch_bounds__(NT) fft_rows_fwd_kernel(const float* __restrict__ x, float2* __restrict__ spec,
                                                                   BlockW w, int Wrt) {
  const int W = WCT ? WCT : Wrt;                          // WCT != 0: compile-time row length (fast path)
  constexpr int NF1 = C2 / 2;                             // complex sequences per image row (two real channels each)
  constexpr int NF = NF1 * ROWS;                          // ROWS consecutive image rows per CTA: more work per barrier
  constexpr int CIN = PRE_LN ? 2 * C2 : C2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tw = reinterpret_cast<float2*>(smem_raw);       // [W]
  const int LS = W + kSeqPad;                             // padded sequence stride
  float2* bufA = tw + W;                                  // [NF][LS]
  float2* bufB = bufA + NF * LS;
  const int tid = threadIdx.x;
  const size_t row0 = (size_t)blockIdx.x * ROWS;          // n*H + y of the first row
  for (int j = tid; j < W; j += NT) tw[j] = tw_entry(j, W);
  for (int p = tid; p < ROWS * W; p += NT) {
    const int rl = p / W, px = p - rl * W;
    const float* src = x + ((row0 + rl) * W + px) * CIN;
    float g[C2];
    if constexpr (PRE_LN) {
      float v[CIN];
      load_vec<CIN>(v, src);
      float mean = 0.f;
#pragma unroll
      for (int i = 0; i < CIN; ++i) mean += v[i];
      mean *= (1.0f / CIN);
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < CIN; ++i) { float d = v[i] - mean; var = fmaf(d, d, var); }
      float rstd = 1.0f / sqrtf(var * (1.0f / CIN) + kLnEps);
#pragma unroll
      for (int i = 0; i < C2; ++i)
        g[i] = (v[C2 + i] - mean) * rstd * __ldg(w.ln1_w + C2 + i) + __ldg(w.ln1_b + C2 + i);
    } else {
      load_vec<C2>(g, src);
    }
#pragma unroll
    for (int f = 0; f < NF1; ++f) bufA[(rl * NF1 + f) * LS + px] = make_float2(g[2 * f], g[2 * f + 1]);
  }
  __syncthreads();
  const float2* res;
  if constexpr (WCT != 0) res = stockham_ct<WCT, NF, -1, NT, 1>(bufA, bufB, tw, tid);
  else res = stockham<-1, ANY>(bufA, bufB, tw, W, NF, tid, NT);
  // split the packed transform: channel a = 2f (real part), b = 2f+1 (imag part)
  const int Wf = W / 2 + 1;
  float4* out = reinterpret_cast<float4*>(spec + row0 * Wf * C2);      // ROWS rows are contiguous in the spectrum
  for (int id = tid; id < ROWS * Wf * NF1; id += NT) {
    const int rl = id / (Wf * NF1), rem = id - rl * (Wf * NF1);
    const int k = rem / NF1, f = rl * NF1 + (rem - k * NF1);
    float2 z = res[f * LS + k], zm = res[f * LS + (k ? W - k : 0)];
    float4 o;
    o.x = 0.5f * (z.x + zm.x);        // Xa = (Z[k] + conj(Z[W-k])) / 2
    o.y = 0.5f * (z.y - zm.y);
    o.z = 0.5f * (z.y + zm.y);        // Xb = (Z[k] - conj(Z[W-k])) / (2i)
    o.w = 0.5f * (zm.x - z.x);
    out[id] = o;
  }
}

// ---- pass 2: columns (forward, pointwise, inverse) -----------------------------------------------------
// One CTA = Q adjacent complex lanes (kx*C2+ch) of one image, all H rows, in shared memory [H][Q].
template <int Q, int HCT>
__global__ void __launch_bounds__(kFftThreads) fft_cols_kernel(float2* __restrict__ spec, BlockW w, int Hrt, int W, int C2,
                                                               int lanes_per_row /* Wf*C2 */) {
  const int H = HCT ? HCT : Hrt;                          // HCT != 0: compile-time column length (fast path, loops unroll)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tw = reinterpret_cast<float2*>(smem_raw);       // [H]
  float2* d = tw + H;                                     // [H][Q]
  const int tid = threadIdx.x;
  const int l0 = blockIdx.x * Q;
  const int nl = min(Q, lanes_per_row - l0);
  float2* base = spec + (size_t)blockIdx.y * H * lanes_per_row + l0;
  for (int j = tid; j < H; j += kFftThreads) tw[j] = tw_entry(j, H);
#pragma unroll 4
  for (int id = tid; id < H * Q; id += kFftThreads) {
    const int r = id / Q, l = id - r * Q;
    d[id] = (l < nl) ? base[(size_t)r * lanes_per_row + l] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const bool odd = (__ffs(H) - 1) & 1;                    // log2(H) odd -> one trailing radix-2 stage
  // forward: decimation in frequency, natural in -> digit-reversed out
#pragma unroll
  for (int L = H; L >= 4; L >>= 2) {
    const int q = L >> 2, tws = H / L;
#pragma unroll
    for (int id = tid; id < (H >> 2) * Q; id += kFftThreads) {
      const int l = id % Q, bf = id / Q;
      const int blk = bf / q, j = bf - blk * q;
      float2* p = d + (size_t)(blk * L + j) * Q + l;
      const float2 a0 = p[0], a1 = p[(size_t)q * Q], a2 = p[(size_t)2 * q * Q], a3 = p[(size_t)3 * q * Q];
      const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
      const float2 mjd = make_float2(d13.y, -d13.x);      // -i * (a1 - a3)
      float2 y1 = cadd(d02, mjd), y2 = csub(s02, s13), y3 = csub(d02, mjd);
      if (q > 1) {                                         // j == 0 twiddles are exactly 1: multiplying is harmless
        y1 = ctw<-1>(y1, tw[j * tws]); y2 = ctw<-1>(y2, tw[2 * j * tws]); y3 = ctw<-1>(y3, tw[3 * j * tws]);
      }
      p[0] = cadd(s02, s13); p[(size_t)q * Q] = y1; p[(size_t)2 * q * Q] = y2; p[(size_t)3 * q * Q] = y3;
    }
    __syncthreads();
  }
  if (odd) {
#pragma unroll
    for (int id = tid; id < (H >> 1) * Q; id += kFftThreads) {
      const int l = id % Q, bf = id / Q;
      float2* p = d + (size_t)(2 * bf) * Q + l;
      const float2 a0 = p[0], a1 = p[Q];
      p[0] = cadd(a0, a1);
      p[Q] = csub(a0, a1);
    }
    __syncthreads();
  }
  // pointwise amplitude / phase mixing.  ky = 0 sits at position 0, ky = H/2 at position 1 (odd log2 H) or 2.
  const int pos_nyq = odd ? 1 : 2;
#pragma unroll 4
  for (int id = tid; id < H * Q; id += kFftThreads) {
    const int pos = id / Q, l = id - pos * Q;
    if (l >= nl) continue;
    const int lane = l0 + l;
    const int kx = lane / C2, ch = lane - kx * C2;
    float2 z = d[id];
    if ((pos == 0 || pos == pos_nyq) && (kx == 0 || kx == W / 2)) z.y = 0.0f;   // exactly-real bins: +0.0 (F7)
    // |z| by sqrt (no overflow risk at these magnitudes); the angle keeps the accurate atan2f because the phase weight
    // amplifies its error (F7); sin/cos of the mixed phase go to the SFU: |pha'| stays within a few radians, where
    // sin.approx / cos.approx are good to ~4e-7 absolute, i.e. below the fp32 rounding noise of the FFT itself
    float amp = sqrtf(fmaf(z.x, z.x, z.y * z.y));
    float pha = atan2f(z.y, z.x);
    amp = amp * __ldg(w.amp_w + ch) + __ldg(w.amp_b + ch);
    pha = pha * __ldg(w.pha_w + ch) + __ldg(w.pha_b + ch);
    const float sn = __sinf(pha), cs = __cosf(pha);
    float re = amp * cs + 1e-8f;
    float im = amp * sn + 1e-8f;
    re = re + 1e-8f;                                       // complex(real, imag) + 1e-8 adds to the real part
    d[id] = make_float2(re, im);
  }
  __syncthreads();
  // inverse: decimation in time mirror of the forward graph, digit-reversed in -> natural out (unnormalised)
  if (odd) {
#pragma unroll
    for (int id = tid; id < (H >> 1) * Q; id += kFftThreads) {
      const int l = id % Q, bf = id / Q;
      float2* p = d + (size_t)(2 * bf) * Q + l;
      const float2 a0 = p[0], a1 = p[Q];
      p[0] = cadd(a0, a1);
      p[Q] = csub(a0, a1);
    }
    __syncthreads();
  }
#pragma unroll
  for (int L = odd ? 8 : 4; L <= H; L <<= 2) {
    const int q = L >> 2, tws = H / L;
#pragma unroll
    for (int id = tid; id < (H >> 2) * Q; id += kFftThreads) {
      const int l = id % Q, bf = id / Q;
      const int blk = bf / q, j = bf - blk * q;
      float2* p = d + (size_t)(blk * L + j) * Q + l;
      float2 a0 = p[0], a1 = p[(size_t)q * Q], a2 = p[(size_t)2 * q * Q], a3 = p[(size_t)3 * q * Q];
      if (q > 1) {
        a1 = ctw<+1>(a1, tw[j * tws]); a2 = ctw<+1>(a2, tw[2 * j * tws]); a3 = ctw<+1>(a3, tw[3 * j * tws]);
      }
      const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
      const float2 jd = make_float2(-d13.y, d13.x);       // +i * (a1 - a3)
      p[0] = cadd(s02, s13);
      p[(size_t)q * Q] = cadd(d02, jd);
      p[(size_t)2 * q * Q] = csub(s02, s13);
      p[(size_t)3 * q * Q] = csub(d02, jd);
    }
    __syncthreads();
  }
#pragma unroll 4
  for (int id = tid; id < H * Q; id += kFftThreads) {
    const int r = id / Q, l = id - r * Q;
    if (l < nl) base[(size_t)r * lanes_per_row + l] = d[id];
  }
}

// ---- pass 2 for column lengths that are not powers of two: the same three steps with the natural-order Stockham
// transform (two shared-memory buffers, sequence = one complex lane's column) ------------------------------------------
template <int Q>
__global__ void __launch_bounds__(kFftThreads) fft_cols_any_kernel(float2* __restrict__ spec, BlockW w, int H, int W, int C2,
                                                                   int lanes_per_row /* Wf*C2 */) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tw = reinterpret_cast<float2*>(smem_raw);       // [H]
  const int LS = H + kSeqPad;
  float2* bufA = tw + H;                                  // [Q][LS]
  float2* bufB = bufA + Q * LS;
  const int tid = threadIdx.x;
  const int l0 = blockIdx.x * Q;
  const int nl = min(Q, lanes_per_row - l0);
  float2* base = spec + (size_t)blockIdx.y * H * lanes_per_row + l0;
  for (int j = tid; j < H; j += kFftThreads) tw[j] = tw_entry(j, H);
  for (int id = tid; id < H * Q; id += kFftThreads) {
    const int r = id / Q, l = id - r * Q;
    bufA[l * LS + r] = (l < nl) ? base[(size_t)r * lanes_per_row + l] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  float2* f = stockham<-1, true>(bufA, bufB, tw, H, Q, tid, kFftThreads);
  for (int id = tid; id < H * Q; id += kFftThreads) {
    const int l = id / H, ky = id - l * H;
    if (l >= nl) continue;
    const int lane = l0 + l;
    const int kx = lane / C2, ch = lane - kx * C2;
    float2 z = f[l * LS + ky];
    if ((ky == 0 || 2 * ky == H) && (kx == 0 || 2 * kx == W)) z.y = 0.0f;     // exactly-real bins: +0.0 (F7)
    float amp = sqrtf(fmaf(z.x, z.x, z.y * z.y));
    float pha = atan2f(z.y, z.x);
    amp = amp * __ldg(w.amp_w + ch) + __ldg(w.amp_b + ch);
    pha = pha * __ldg(w.pha_w + ch) + __ldg(w.pha_b + ch);
    const float sn = __sinf(pha), cs = __cosf(pha);
    float re = amp * cs + 1e-8f;
    float im = amp * sn + 1e-8f;
    re = re + 1e-8f;
    f[l * LS + ky] = make_float2(re, im);
  }
  __syncthreads();
  float2* g = stockham<+1, true>(f, f == bufA ? bufB : bufA, tw, H, Q, tid, kFftThreads);
  for (int id = tid; id < H * Q; id += kFftThreads) {
    const int r = id / Q, l = id - r * Q;
    if (l < nl) base[(size_t)r * lanes_per_row + l] = g[l * LS + r];
  }
}

// ---- pass 3: rows inverse (+ proj + residual) ------------------------------------------------------------
template <int C2, bool PROJ, int WCT, int ROWS, int NT = kFftThreads, bool ANY = false>
__global__ void __launch_bounds__(NT) fft_rows_inv_kernel(const float2* __restrict__ spec,
                                                                   const float* __restrict__ local,
                                                                   const float* __restrict__ xres, float* __restrict__ y,
                                                                   BlockW w, int Wrt, float scale) {
  const int W = WCT ? WCT : Wrt;
  constexpr int NF1 = C2 / 2;
  constexpr int NF = NF1 * ROWS;                          // ROWS consecutive image rows per CTA
  constexpr int C = 2 * C2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tw = reinterpret_cast<float2*>(smem_raw);       // [W]
  const int LS = W + kSeqPad;                             // padded sequence stride
  float2* bufA = tw + W;                                  // [NF][LS]
  float2* bufB = bufA + NF * LS;
  float* sW = reinterpret_cast<float*>(bufB + NF * LS);    // [C][C] proj weight (PROJ only)
  float* sBias = sW + C * C;                              // [C]
  const int tid = threadIdx.x;
  const size_t row0 = (size_t)blockIdx.x * ROWS;
  const int Wf = W / 2 + 1;
  for (int j = tid; j < W; j += NT) tw[j] = tw_entry(j, W);
  if constexpr (PROJ) {
    for (int i = tid; i < C * C; i += NT) sW[i] = __ldg(w.proj_w + i);
    for (int i = tid; i < C; i += NT) sBias[i] = __ldg(w.proj_b + i);
  }
  const float4* in = reinterpret_cast<const float4*>(spec + row0 * Wf * C2);
  for (int id = tid; id < ROWS * Wf * NF1; id += NT) {
    const int rl = id / (Wf * NF1), rem = id - rl * (Wf * NF1);
    const int k = rem / NF1, f = rl * NF1 + (rem - k * NF1);
    float4 v = __ldg(in + id);                            // (Xa.re, Xa.im, Xb.re, Xb.im)
    if (k == 0 || k == W / 2) {                           // C2R ignores Im of the DC and Nyquist bins
      bufA[f * LS + k] = make_float2(v.x, v.z);
    } else {
      bufA[f * LS + k] = make_float2(v.x - v.w, v.y + v.z);          // Xa + i Xb
      bufA[f * LS + W - k] = make_float2(v.x + v.w, v.z - v.y);      // conj(Xa) + i conj(Xb)
    }
  }
  __syncthreads();
  const float2* res;
  if constexpr (WCT != 0) res = stockham_ct<WCT, NF, +1, NT, 1>(bufA, bufB, tw, tid);
  else res = stockham<+1, ANY>(bufA, bufB, tw, W, NF, tid, NT);
  for (int p = tid; p < ROWS * W; p += NT) {
    const int rl = p / W, px = p - rl * W;
    const size_t row = row0 + rl;
    float g[C2];
#pragma unroll
    for (int f = 0; f < NF1; ++f) {
      float2 z = res[(rl * NF1 + f) * LS + px];
      g[2 * f] = fabsf(z.x * scale);
      g[2 * f + 1] = fabsf(z.y * scale);
    }
    if constexpr (!PROJ) {
      store_vec<C2>(y + (row * W + px) * C2, g);
    } else {
      float cat[C];
      {
        float t[C2];
        load_vec<C2>(t, local + (row * W + px) * C2);
#pragma unroll
        for (int i = 0; i < C2; ++i) { cat[i] = t[i]; cat[C2 + i] = g[i]; }
      }
      const float* xr = xres + (row * W + px) * C;
      float* dst = y + (row * W + px) * C;
#pragma unroll 1
      for (int o = 0; o < C; o += 4) {
        float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
        const float4* r0 = reinterpret_cast<const float4*>(sW + (o + 0) * C);
        const float4* r1 = reinterpret_cast<const float4*>(sW + (o + 1) * C);
        const float4* r2 = reinterpret_cast<const float4*>(sW + (o + 2) * C);
        const float4* r3 = reinterpret_cast<const float4*>(sW + (o + 3) * C);
#pragma unroll
        for (int k4 = 0; k4 < C / 4; ++k4) {
          const float2 va = make_float2(cat[4 * k4], cat[4 * k4 + 1]), vb = make_float2(cat[4 * k4 + 2], cat[4 * k4 + 3]);
          const float4 w0 = r0[k4], w1 = r1[k4], w2 = r2[k4], w3 = r3[k4];
          a0 = __ffma2_rn(make_float2(w0.x, w0.y), va, a0); a0 = __ffma2_rn(make_float2(w0.z, w0.w), vb, a0);
          a1 = __ffma2_rn(make_float2(w1.x, w1.y), va, a1); a1 = __ffma2_rn(make_float2(w1.z, w1.w), vb, a1);
          a2 = __ffma2_rn(make_float2(w2.x, w2.y), va, a2); a2 = __ffma2_rn(make_float2(w2.z, w2.w), vb, a2);
          a3 = __ffma2_rn(make_float2(w3.x, w3.y), va, a3); a3 = __ffma2_rn(make_float2(w3.z, w3.w), vb, a3);
        }
        float4 r = *reinterpret_cast<const float4*>(xr + o);
        *reinterpret_cast<float4*>(dst + o) =
            make_float4(((a0.x + a0.y) + sBias[o]) + r.x, ((a1.x + a1.y) + sBias[o + 1]) + r.y,
                        ((a2.x + a2.y) + sBias[o + 2]) + r.z, ((a3.x + a3.y) + sBias[o + 3]) + r.w);
      }
    }
  }
}

// ---- launchers ---------------------------------------------------------------------------------------------
static bool pow2_in_range(int v) { return v >= 8 && v <= kTwN && (v & (v - 1)) == 0; }
// every length the window partition allows: multiples of 8 (odd factors take the generic radix stages)
static bool len_ok(int v) { return v >= 8 && v <= kTwN && (v & 7) == 0; }

// rows per CTA on the fast path: 16 complex sequences per CTA (C2=8: 4 rows, C2=16: 2 rows, C2=32: 1 row)
template <int C2> constexpr int fast_rows() { return (16 / (C2 / 2)) > 0 ? 16 / (C2 / 2) : 1; }

template <int C2, bool PRE_LN, int WCT, int ROWS, int NT = kFftThreads, bool ANY = false>
static cudaError_t rows_fwd_launch(const BlockW& w, const float* x, float* spec, int N, int H, int W, cudaStream_t s) {
  size_t smem = (size_t)(W + 2 * ROWS * (C2 / 2) * (W + kSeqPad)) * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(fft_rows_fwd_kernel<C2, PRE_LN, WCT, ROWS, NT, ANY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fft_rows_fwd_kernel<C2, PRE_LN, WCT, ROWS, NT, ANY><<<N * H / ROWS, NT, smem, s>>>(x, reinterpret_cast<float2*>(spec), w, W);
  return cudaGetLastError();
}
template <int C2>
static cudaError_t rows_fwd_t(const BlockW& w, const float* x, float* spec, int pre_ln, int N, int H, int W, cudaStream_t s) {
  if (W & (W - 1)) {                                       // row length with odd factors: generic radix stages
    if (pre_ln) {
      if (W >= 512) return rows_fwd_launch<C2, true, 0, 1, 512, true>(w, x, spec, N, H, W, s);
      return rows_fwd_launch<C2, true, 0, 1, kFftThreads, true>(w, x, spec, N, H, W, s);
    }
    return rows_fwd_launch<C2, false, 0, 1, kFftThreads, true>(w, x, spec, N, H, W, s);
  }
  if (pre_ln) {
    if (W == 256 && H % fast_rows<C2>() == 0) return rows_fwd_launch<C2, true, 256, fast_rows<C2>()>(w, x, spec, N, H, W, s);
    if (W == 128 && H % fast_rows<C2>() == 0) return rows_fwd_launch<C2, true, 128, fast_rows<C2>()>(w, x, spec, N, H, W, s);
    if (W >= 1024) return rows_fwd_launch<C2, true, 0, 1, 1024>(w, x, spec, N, H, W, s);
    if (W >= 512) return rows_fwd_launch<C2, true, 0, 1, 512>(w, x, spec, N, H, W, s);
    return rows_fwd_launch<C2, true, 0, 1>(w, x, spec, N, H, W, s);
  }
  return rows_fwd_launch<C2, false, 0, 1>(w, x, spec, N, H, W, s);
}

static bool stockham_only() {                            // LGTEUN_FFT=stockham: A/B switch back to the shared-memory passes
  static const bool v = [] { const char* e = getenv("LGTEUN_FFT"); return e && e[0] == 's'; }();
  return v;
}

cudaError_t launch_fft_rows_fwd(const BlockW& w, int c, const float* x, float* spec, int pre_ln, int N, int H, int W,
                                cudaStream_t s) {
  if (!len_ok(W) || !len_ok(H)) return cudaErrorInvalidValue;
  if (W == 256 && pre_ln && H % 4 == 0 && !stockham_only()) return launch_fft_rows_fwd256(w, c, x, spec, N, H, s);
  if (W == 128 && pre_ln && H % 8 == 0 && !stockham_only()) return launch_fft_rows_fwd128(w, c, x, spec, N, H, s);
  switch (c) {
    case 16: return rows_fwd_t<8>(w, x, spec, pre_ln, N, H, W, s);
    case 32: return rows_fwd_t<16>(w, x, spec, pre_ln, N, H, W, s);
    case 64: return rows_fwd_t<32>(w, x, spec, pre_ln, N, H, W, s);
    default: return cudaErrorInvalidValue;
  }
}

template <int Q, int HCT>
static cudaError_t cols_t(const BlockW& w, int c2, float* spec, int N, int H, int W, cudaStream_t s) {
  const int lanes = (W / 2 + 1) * c2;
  size_t smem = (size_t)(H + H * Q) * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(fft_cols_kernel<Q, HCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((lanes + Q - 1) / Q, N);
  fft_cols_kernel<Q, HCT><<<grid, kFftThreads, smem, s>>>(reinterpret_cast<float2*>(spec), w, H, W, c2, lanes);
  return cudaGetLastError();
}

cudaError_t launch_fft_cols(const BlockW& w, int c, float* spec, int N, int H, int W, cudaStream_t s) {
  if (!len_ok(W) || !len_ok(H)) return cudaErrorInvalidValue;
  const int c2 = c / 2;
  if (!pow2_in_range(H)) {                                 // column length with odd factors
    const int lanes = (W / 2 + 1) * c2;
    constexpr int Q = 8;
    size_t smem = (size_t)(H + 2 * Q * (H + kSeqPad)) * sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(fft_cols_any_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((lanes + Q - 1) / Q, N);
    fft_cols_any_kernel<Q><<<grid, kFftThreads, smem, s>>>(reinterpret_cast<float2*>(spec), w, H, W, c2, lanes);
    return cudaGetLastError();
  }
  if (H == 128 && !stockham_only()) return launch_fft_cols128(w, c2, spec, N, W, s);
  if (H == 128) return cols_t<64, 128>(w, c2, spec, N, H, W, s);
  if (H < 128) return cols_t<64, 0>(w, c2, spec, N, H, W, s);
  if (H == 256 && !stockham_only()) return launch_fft_cols256(w, c2, spec, N, W, s);
  if (H == 256) return cols_t<32, 256>(w, c2, spec, N, H, W, s);
  if (H == 512) return cols_t<16, 0>(w, c2, spec, N, H, W, s);
  return cols_t<8, 0>(w, c2, spec, N, H, W, s);
}

template <int C2>
static cudaError_t rows_inv_t(const BlockW& w, const float* spec, const float* local, const float* xres, float* y, int proj,
                              int N, int H, int W, cudaStream_t s) {
  constexpr int C = 2 * C2;
  // measured: several rows per CTA do not pay here (the proj weights + 64 KB of buffers cut residency: 89 vs 67 us per
  // 16 pairs), so the inverse pass keeps one row per CTA
  const int fr = (proj && (W == 256 || W == 128) && fast_rows<C2>() >= 2 && H % 2 == 0) ? 2 : 1;
  size_t smem = (size_t)(W + 2 * fr * (C2 / 2) * (W + kSeqPad)) * sizeof(float2) + (proj ? (size_t)(C * C + C) * sizeof(float) : 0);
  const float scale = 1.0f / ((float)H * (float)W);
  auto go = [&](auto kern, int nt = kFftThreads) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<N * H / fr, nt, smem, s>>>(reinterpret_cast<const float2*>(spec), local, xres, y, w, W, scale);
    return cudaGetLastError();
  };
  if (W & (W - 1)) {                                       // row length with odd factors: generic radix stages
    if (proj && W >= 512) return go(fft_rows_inv_kernel<C2, true, 0, 1, 512, true>, 512);
    if (proj) return go(fft_rows_inv_kernel<C2, true, 0, 1, kFftThreads, true>);
    return go(fft_rows_inv_kernel<C2, false, 0, 1, kFftThreads, true>);
  }
  if (proj) {
    if (W == 256 && fr > 1) return go(fft_rows_inv_kernel<C2, true, 256, 2>);
    if (W == 128 && fr > 1) return go(fft_rows_inv_kernel<C2, true, 128, 2>);
    if (W == 256) return go(fft_rows_inv_kernel<C2, true, 256, 1>);
    if (W == 128) return go(fft_rows_inv_kernel<C2, true, 128, 1>);
    if (W >= 1024) return go(fft_rows_inv_kernel<C2, true, 0, 1, 1024>, 1024);
    if (W >= 512) return go(fft_rows_inv_kernel<C2, true, 0, 1, 512>, 512);
    return go(fft_rows_inv_kernel<C2, true, 0, 1>);
  }
  return go(fft_rows_inv_kernel<C2, false, 0, 1>);
}

cudaError_t launch_fft_rows_inv(const BlockW& w, int c, const float* spec, const float* local, const float* xres, float* y,
                                int proj, int N, int H, int W, cudaStream_t s) {
  if (!len_ok(W) || !len_ok(H)) return cudaErrorInvalidValue;
  if (W == 256 && proj && H % 4 == 0 && !stockham_only()) return launch_fft_rows_inv256(w, c, spec, local, xres, y, N, H, s);
  if (W == 128 && proj && H % 8 == 0 && !stockham_only()) return launch_fft_rows_inv128(w, c, spec, local, xres, y, N, H, s);
  switch (c) {
    case 16: return rows_inv_t<8>(w, spec, local, xres, y, proj, N, H, W, s);
    case 32: return rows_inv_t<16>(w, spec, local, xres, y, proj, N, H, W, s);
    case 64: return rows_inv_t<32>(w, spec, local, xres, y, proj, N, H, W, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace lg
