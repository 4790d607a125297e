// fft256.cu — register-resident 256-point FFT passes for the global mixer (models/common/LGT.py:162-180) at the
// headline resolution (H = 256 columns / W = 256 rows).  256 = 16 x 16: a thread holds 16 complex values, runs a
// 16-point FFT (two radix-4 levels, compile-time twiddles) entirely in registers, exchanges once through shared memory
// and runs the second 16-point FFT.  Compared with the shared-memory Stockham passes of fft_mixer.cu (8 stages, a barrier
// and runtime index arithmetic per stage) this needs ~3x fewer instructions and 2-3 barriers per transform, which is
// what bounded those kernels (issue / barrier latency, 14-36 % of HBM speed).
//
// Column pass (this file): forward FFT along H, amplitude/phase mixing on the registers that hold the spectrum,
// inverse FFT along H, in place on S[n][y][kx][ch].  The forward transform leaves thread k1 with X[k1 + 16 k2]
// (k2 = register index) — exactly the "fixed low index, stride-16" distribution the inverse transform consumes, so no
// reordering is needed between the two.
#include <math.h>
#include <stdlib.h>

#include <string>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace lg {

__device__ float2 g_tw256[256];                          // e^{-2 pi i k / 256}, rounded from double

cudaError_t fft256_init_tables(cudaStream_t s) {
  static float2 host[256];
  for (int k = 0; k < 256; ++k) {
    double a = -2.0 * M_PI * (double)k / 256.0;
    host[k] = make_float2((float)cos(a), (float)sin(a));
  }
  host[0] = make_float2(1.f, 0.f);
  host[64] = make_float2(0.f, -1.f);
  host[128] = make_float2(-1.f, 0.f);
  host[192] = make_float2(0.f, 1.f);
  cudaError_t e = cudaMemcpyToSymbolAsync(g_tw256, host, sizeof(host), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(s);
}

namespace f256 {
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a * w (SIGN < 0: forward twiddle) or a * conj(w) (SIGN > 0: inverse)
template <int SIGN>
__device__ __forceinline__ float2 ctw(float2 a, float2 w) {
  const float2 t = __fmul2_rn(a, make_float2(w.x, w.x));
  return (SIGN < 0) ? __ffma2_rn(make_float2(-a.y, a.x), make_float2(w.y, w.y), t)
                    : __ffma2_rn(make_float2(a.y, -a.x), make_float2(w.y, w.y), t);
}
// multiply by -i (SIGN < 0) or +i (SIGN > 0)
template <int SIGN>
__device__ __forceinline__ float2 rot(float2 a) { return (SIGN < 0) ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x); }

template <int SIGN>
__device__ __forceinline__ void radix4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = rot<SIGN>(csub(a1, a3));
  a0 = cadd(s02, s13);
  a1 = cadd(d02, d13);
  a2 = csub(s02, s13);
  a3 = csub(d02, d13);
}

// 16-point FFT in registers, natural order in and out (n = j + 4q, k = m + 4p).  Unnormalised; SIGN = -1 forward.
template <int SIGN>
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  // level 1: over q for each j -> t[j][m] stored at v[j + 4m]
#pragma unroll
  for (int j = 0; j < 4; ++j) radix4<SIGN>(v[j], v[j + 4], v[j + 8], v[j + 12]);
  // twiddles W16^{j m}, j,m in 1..3  (forward values; ctw conjugates them for the inverse)
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R = 0.70710678118654752f;
  v[1 + 4 * 1] = ctw<SIGN>(v[1 + 4 * 1], make_float2(C1, -S1));      // W^1
  v[1 + 4 * 2] = ctw<SIGN>(v[1 + 4 * 2], make_float2(R, -R));        // W^2
  v[1 + 4 * 3] = ctw<SIGN>(v[1 + 4 * 3], make_float2(S1, -C1));      // W^3
  v[2 + 4 * 1] = ctw<SIGN>(v[2 + 4 * 1], make_float2(R, -R));        // W^2
  v[2 + 4 * 2] = rot<SIGN>(v[2 + 4 * 2]);                            // W^4 = -i
  v[2 + 4 * 3] = ctw<SIGN>(v[2 + 4 * 3], make_float2(-R, -R));       // W^6
  v[3 + 4 * 1] = ctw<SIGN>(v[3 + 4 * 1], make_float2(S1, -C1));      // W^3
  v[3 + 4 * 2] = ctw<SIGN>(v[3 + 4 * 2], make_float2(-R, -R));       // W^6
  v[3 + 4 * 3] = ctw<SIGN>(v[3 + 4 * 3], make_float2(-C1, S1));      // W^9
  // level 2: over j for each m: inputs v[0+4m], v[1+4m], v[2+4m], v[3+4m] -> X[m], X[m+4], X[m+8], X[m+12]
  float2 o[16];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    float2 b0 = v[4 * m], b1 = v[4 * m + 1], b2 = v[4 * m + 2], b3 = v[4 * m + 3];
    radix4<SIGN>(b0, b1, b2, b3);
    o[m] = b0; o[m + 4] = b1; o[m + 8] = b2; o[m + 12] = b3;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = o[i];
}
// 8-point FFT in registers, natural order in and out: two radix-4 on the even / odd samples, W8 twiddles, radix-2
template <int SIGN>
__device__ __forceinline__ void fft8(float2* v) {
  constexpr float R = 0.70710678118654752f;
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  radix4<SIGN>(e0, e1, e2, e3);
  radix4<SIGN>(o0, o1, o2, o3);
  o1 = ctw<SIGN>(o1, make_float2(R, -R));                 // W8^1
  o2 = rot<SIGN>(o2);                                     // W8^2 = -i
  o3 = ctw<SIGN>(o3, make_float2(-R, -R));                // W8^3
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}
// M-point FFT on v[0..M), M in {8, 16}
template <int M, int SIGN>
__device__ __forceinline__ void fft_m(float2* v) {
  if constexpr (M == 16) fft16<SIGN>(*reinterpret_cast<float2(*)[16]>(v));
  else fft8<SIGN>(v);
}
}  // namespace f256

// atan2 without branches: octant reduction, t = min / max in [0, 1], atan(t) = t P(t^2) with a degree-8 P fitted on Chebyshev nodes
// (max |error| 1.1e-7 rad in fp32 evaluation, the level of atan2f's 2 ulp near pi/4), sign handling on the sign BITS so that
// (+0, x < 0) -> +pi, (-0, x < 0) -> -pi and (+0, -0) -> +pi as IEEE atan2 / torch.angle have it (the F7 rule feeds +0.0 here).
// atan2f's slow paths cost ~35 instructions and three divergent branches per bin; this is ~24 straight-line.
__device__ __forceinline__ float atan2_poly(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = mx > 0.f ? __fdividef(mn, mx) : 0.f;
  const float s = t * t;
  float p = 0.0028340641874819994f;
  p = fmaf(p, s, -0.016005029901862144f);
  p = fmaf(p, s, 0.042587608098983765f);
  p = fmaf(p, s, -0.07495445758104324f);
  p = fmaf(p, s, 0.10636754333972931f);
  p = fmaf(p, s, -0.14202570915222168f);
  p = fmaf(p, s, 0.19992484152317047f);
  p = fmaf(p, s, -0.3333306610584259f);
  p = fmaf(p, s, 1.0f);
  float r = p * t;
  r = ay > ax ? 1.5707963267948966f - r : r;
  r = (__float_as_uint(x) >> 31) ? 3.14159265358979323846f - r : r;
  return copysignf(r, y);
}
__device__ __forceinline__ float sqrt_fast(float v) {    // one MUFU, <= 1 ulp; no denormal / special-case branch
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// ---- column pass at H = 16 M (M = 16: 256, M = 8: 128) --------------------------------------------------------------------
// CTA = Q adjacent complex lanes (kx*C2 + ch) of one image; thread = (lane, low index); M * Q threads.
// H = 16 M is split as n = M n1 + n2: pass A is a 16-point FFT over n1 (one per thread, n2 = thread), twiddle W_H^{n2 k1},
// exchange, pass B an M-point FFT over n2 for each k1 (16/M of them per thread), leaving X[k1 + 16 k2] in registers.
// The inverse consumes exactly that distribution (m = m2 + 16 m1 with m2 = k1, m1 = k2): M-point over m1, twiddle,
// exchange, 16-point over m2, so no reordering is needed between the two transforms.
// MODE 0: forward + mixing + inverse (the global mixer).  MODE 1: forward only (MODE 3: the same without the real-bin rule,
// the adjoint of the inverse in the training backward), MODE 2: unnormalised inverse only — the plain
// column transforms of the companion operator SFIIN.Freprocess (companion_ops.cu), which mixes channels between the two.
template <int Q, int M, int MODE = 0>
__global__ void __launch_bounds__(M * Q, 3) fft_cols256_kernel(float2* __restrict__ spec, BlockW w, int W, int C2,
                                                                int lanes_per_row, int exact_math = 0) {
  using namespace f256;
  constexpr int H = 16 * M, KPT = 16 / M, TS = 16 / M;    // k1 values per thread in pass B; twiddle-table stride
  __shared__ float2 tw[256];
  extern __shared__ __align__(16) float2 ex[];            // exchange buffer [16][M][Q]
  const int tid = threadIdx.x;
  const int l = tid % Q, lo = tid / Q;                    // lane within the CTA, low index (n2 / j1)
  const int l0 = blockIdx.x * Q;
  const bool live = l0 + l < lanes_per_row;
  float2* base = spec + (size_t)blockIdx.y * H * lanes_per_row + l0 + l;
  for (int j = tid; j < 256; j += M * Q) tw[j] = g_tw256[j];

  float2 v[16];
  if constexpr (MODE == 2) {                              // the spectrum in the distribution the inverse consumes
#pragma unroll
    for (int i = 0; i < 16; ++i)
      v[i] = live ? base[(size_t)((lo + M * (i / M)) + 16 * (i % M)) * lanes_per_row] : make_float2(0.f, 0.f);
    __syncthreads();                                      // twiddle table staged
  } else {
  // load rows y = M n1 + lo
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1)
    v[n1] = live ? base[(size_t)(M * n1 + lo) * lanes_per_row] : make_float2(0.f, 0.f);
  __syncthreads();                                        // twiddle table staged
  fft16<-1>(v);
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) v[k1] = ctw<-1>(v[k1], tw[TS * lo * k1]);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) ex[(k1 * M + lo) * Q + l] = v[k1];
  __syncthreads();
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk) {
#pragma unroll
    for (int n2 = 0; n2 < M; ++n2) v[kk * M + n2] = ex[((lo + M * kk) * M + n2) * Q + l];
    fft_m<M, -1>(v + kk * M);                             // v[kk*M + k2] = X[ky = (lo + M kk) + 16 k2]
  }
  }
  if constexpr (MODE == 1 || MODE == 3) {
    const bool real_col = MODE == 1 && ((l0 + l) / C2 == 0 || (l0 + l) / C2 == W / 2);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int kk = i / M, k2 = i % M;
      if (real_col && lo == 0 && kk == 0 && (k2 == 0 || k2 == M / 2)) v[i].y = 0.0f;      // exactly-real bins (F7)
      if (live) base[(size_t)((lo + M * kk) + 16 * k2) * lanes_per_row] = v[i];
    }
  } else {
  // amplitude / phase mixing (LGT.py:168-177)
  if constexpr (MODE == 0) {
    const int lane = l0 + l;
    const int kx = lane / C2, ch = lane - kx * C2;
    const float aw = live ? __ldg(w.amp_w + ch) : 0.f, ab = live ? __ldg(w.amp_b + ch) : 0.f;
    const float pw = live ? __ldg(w.pha_w + ch) : 0.f, pb = live ? __ldg(w.pha_b + ch) : 0.f;
    const bool real_col = (kx == 0 || kx == W / 2);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int kk = i / M, k2 = i % M;
      float2 z = v[i];
      // ky in {0, H/2} <=> k1 = 0 and k2 in {0, M/2}: exactly-real bins get +0.0 (F7)
      if (real_col && lo == 0 && kk == 0 && (k2 == 0 || k2 == M / 2)) z.y = 0.0f;
      float amp = exact_math ? sqrtf(fmaf(z.x, z.x, z.y * z.y)) : sqrt_fast(fmaf(z.x, z.x, z.y * z.y));
      float pha = exact_math ? atan2f(z.y, z.x) : atan2_poly(z.y, z.x);
      amp = amp * aw + ab;
      pha = pha * pw + pb;
      const float sn = __sinf(pha), cs = __cosf(pha);
      float re = amp * cs + 1e-8f;
      const float im = amp * sn + 1e-8f;
      re = re + 1e-8f;                                    // complex(real, imag) + 1e-8 adds to the real part
      v[i] = make_float2(re, im);
    }
  }
  // inverse, pass A: M-point over m1 = k2 for each m2 = k1 held by this thread, twiddle conj W_H^{j1 m2}
  if constexpr (MODE == 0) __syncthreads();               // everyone has consumed the first exchange
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk) {
    const int k1 = lo + M * kk;
    fft_m<M, +1>(v + kk * M);
#pragma unroll
    for (int j1 = 1; j1 < M; ++j1) v[kk * M + j1] = ctw<+1>(v[kk * M + j1], tw[TS * k1 * j1]);
#pragma unroll
    for (int j1 = 0; j1 < M; ++j1) ex[(j1 * 16 + k1) * Q + l] = v[kk * M + j1];
  }
  __syncthreads();
  // pass B: 16-point over m2 for j1 = lo
#pragma unroll
  for (int m2 = 0; m2 < 16; ++m2) v[m2] = ex[(lo * 16 + m2) * Q + l];
  fft16<+1>(v);                                           // v[j2] = x[y = lo + M j2]  (unnormalised)
  if (live) {
#pragma unroll
    for (int j2 = 0; j2 < 16; ++j2) base[(size_t)(lo + M * j2) * lanes_per_row] = v[j2];
  }
  }   // MODE != 1
}

template <int M>
static cudaError_t cols_reg_t(const BlockW& w, int c2, float* spec, int N, int W, cudaStream_t s) {
  constexpr int Q = 256 / M;
  const int lanes = (W / 2 + 1) * c2;
  const size_t smem = (size_t)16 * M * Q * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(fft_cols256_kernel<Q, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((lanes + Q - 1) / Q, N);
  static const int exact = [] { const char* e = getenv("LGTEUN_SPEC_MATH"); return (e && std::string(e) == "libm") ? 1 : 0; }();   // A/B: atan2f / sqrtf
  fft_cols256_kernel<Q, M><<<grid, M * Q, smem, s>>>(reinterpret_cast<float2*>(spec), w, W, c2, lanes, exact);
  return cudaGetLastError();
}
cudaError_t launch_fft_cols256(const BlockW& w, int c2, float* spec, int N, int W, cudaStream_t s) {
  return cols_reg_t<16>(w, c2, spec, N, W, s);
}
cudaError_t launch_fft_cols128(const BlockW& w, int c2, float* spec, int N, int W, cudaStream_t s) {
  return cols_reg_t<8>(w, c2, spec, N, W, s);
}

}  // namespace lg

namespace lg {

// ---- row passes at W = 16 M (M = 16: 256, M = 8: 128) ----------------------------------------------------------------------
// A CTA owns 256 / M complex sequences (two real channels each) = 512 / (M C2) consecutive image rows; thread =
// (sequence, low index), M threads per sequence.  Same two-pass split as the column kernel (n = M n1 + n2).
// A sequence occupies 16 (M + 1) slots: inputs are stored with one pad slot per M samples (conflict-free stride-M reads),
// the exchange between the passes is [k1][M + 1], the result comes back in natural order — all in the same storage.
template <int M> __device__ __forceinline__ int padded_m(int i) { return i + i / M; }

// The transform itself: X holds 256 / M packed complex sequences ([seq][RP], one pad slot per M samples); runs the two
// register passes, splits the packed transforms and writes spec rows [row0, row0 + RW).  Starts with the barrier that
// publishes X.
template <int C2, int M>
__device__ __forceinline__ void rows_fwd_transform(float2* X, const float2* tw, float2* __restrict__ spec, size_t row0,
                                                   float scale = 0.5f, float wint = 1.f) {   // scale includes the 1/2 of the split
  using namespace f256;
  constexpr int W = 16 * M, Wf = W / 2 + 1, RP = 16 * (M + 1), KPT = 16 / M, TS = 16 / M;
  constexpr int NF1 = C2 / 2, RW = (256 / M) / NF1;
  float2* E = X;                                          // exchange between the two passes: the SAME storage (35 KB per
                                                          // CTA instead of 70: twice the resident CTAs)
  const int tid = threadIdx.x;
  __syncthreads();
  const int seq = tid / M, lo = tid % M;
  float2 v[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) v[n1] = X[seq * RP + (M + 1) * n1 + lo];       // padded(M n1 + lo)
  fft16<-1>(v);
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) v[k1] = ctw<-1>(v[k1], tw[TS * lo * k1]);
  // A sequence lives in M lanes of one warp and in its own region of the buffer: the hand-offs between the two passes
  // only need warp-level synchronisation
  __syncwarp();                                           // all inputs are in registers: the buffer becomes the exchange
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) E[seq * RP + k1 * (M + 1) + lo] = v[k1];
  __syncwarp();
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk) {
#pragma unroll
    for (int n2 = 0; n2 < M; ++n2) v[kk * M + n2] = E[seq * RP + (lo + M * kk) * (M + 1) + n2];
    fft_m<M, -1>(v + kk * M);                             // v[kk*M + k2] = Z[(lo + M kk) + 16 k2]
  }
  __syncwarp();
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk)
#pragma unroll
    for (int k2 = 0; k2 < M; ++k2) X[seq * RP + (lo + M * kk) + 16 * k2] = v[kk * M + k2];     // natural order, unpadded
  __syncthreads();
  // split the packed transforms: channel a = 2f (real input), b = 2f+1 (imaginary input)
  float4* out = reinterpret_cast<float4*>(spec + row0 * Wf * C2);
  for (int id = tid; id < RW * Wf * NF1; id += 256) {
    const int rl = id / (Wf * NF1), rem = id - rl * (Wf * NF1);
    const int k = rem / NF1, s = rl * NF1 + (rem - k * NF1);
    const float2 z = X[s * RP + k], zm = X[s * RP + ((W - k) & (W - 1))];
    const float f = (k > 0 && k < W / 2) ? scale * wint : scale;
    float4 o;
    o.x = f * (z.x + zm.x);           // Xa = (Z[k] + conj(Z[W-k])) / 2
    o.y = f * (z.y - zm.y);
    o.z = f * (z.y + zm.y);           // Xb = (Z[k] - conj(Z[W-k])) / (2i)
    o.w = f * (zm.x - z.x);
    out[id] = o;
  }
}

template <int C2, bool PRE_LN, int M>
__global__ void __launch_bounds__(256) fft_rows_fwd256_kernel(const float* __restrict__ x, float2* __restrict__ spec,
                                                               BlockW w) {
  using namespace f256;
  constexpr int W = 16 * M, RP = 16 * (M + 1);
  constexpr int NF1 = C2 / 2;                             // sequences per image row
  constexpr int RW = (256 / M) / NF1;                     // image rows per CTA
  static_assert(RW >= 1, "a CTA holds at least one image row");
  constexpr int CIN = PRE_LN ? 2 * C2 : C2;
  __shared__ float2 tw[256];
  extern __shared__ __align__(16) float2 smf[];
  float2* X = smf;                                        // [256/M][RP]  input sequences, later the natural-order spectrum
  const int tid = threadIdx.x;
  const size_t row0 = (size_t)blockIdx.x * RW;
  tw[tid] = g_tw256[tid];
  // phase 0: LayerNorm, pack channel pairs (2f, 2f+1) of the global half as complex samples
#pragma unroll
  for (int i = 0; i < RW * W / 256; ++i) {
    const int p = tid + 256 * i, rl = p / W, px = p % W;
    const float* src = x + ((row0 + rl) * W + px) * CIN;
    float g[C2];
    if constexpr (PRE_LN) {
      float v[CIN];
      load_vec<CIN>(v, src);
      float mean = 0.f;
#pragma unroll
      for (int c = 0; c < CIN; ++c) mean += v[c];
      mean *= (1.0f / CIN);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < CIN; ++c) { float d = v[c] - mean; var = fmaf(d, d, var); }
      const float rstd = 1.0f / sqrtf(var * (1.0f / CIN) + kLnEps);
#pragma unroll
      for (int c = 0; c < C2; ++c) g[c] = (v[C2 + c] - mean) * rstd * __ldg(w.ln1_w + C2 + c) + __ldg(w.ln1_b + C2 + c);
    } else {
      load_vec<C2>(g, src);
    }
#pragma unroll
    for (int f = 0; f < NF1; ++f) X[(rl * NF1 + f) * RP + padded_m<M>(px)] = make_float2(g[2 * f], g[2 * f + 1]);
  }
  rows_fwd_transform<C2, M>(X, tw, spec, row0);
}

// Forward row pass of Freprocess with its `pre1` / `pre2` 1x1 convs as the prologue (models/SFIIN.py:213-214,223-224): reads
// the two NCHW inputs, writes the spectrum of cat(pre1(msf) + 1e-8, pre2(panf) + 1e-8) — the pre-convolved map never
// exists in HBM.  C2 = 2 C channels per pixel.
template <int C2, int M>
__global__ void __launch_bounds__(256) fft_rows_fwd_pre_kernel(const float* __restrict__ msf, const float* __restrict__ panf,
                                                                const float* __restrict__ w1, const float* __restrict__ b1,
                                                                const float* __restrict__ w2, const float* __restrict__ b2,
                                                                float2* __restrict__ spec, int H) {
  using namespace f256;
  constexpr int W = 16 * M, RP = 16 * (M + 1), C = C2 / 2, NF1 = C2 / 2, RW = (256 / M) / NF1;
  static_assert(RW >= 1, "a CTA holds at least one image row");
  __shared__ float2 tw[256];
  __shared__ __align__(16) float sw[2][C][C];
  __shared__ float sb[2][C];
  extern __shared__ __align__(16) float2 smf[];
  float2* X = smf;
  const int tid = threadIdx.x;
  const size_t row0 = (size_t)blockIdx.x * RW;
  tw[tid] = g_tw256[tid];
  for (int i = tid; i < 2 * C * C; i += 256) sw[i / (C * C)][(i / C) % C][i % C] = __ldg((i < C * C ? w1 : w2) + i % (C * C));
  if (tid < 2 * C) sb[tid / C][tid % C] = __ldg((tid < C ? b1 : b2) + tid % C);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RW * W / 256; ++i) {
    const int p = tid + 256 * i, rl = p / W, px = p % W;
    const size_t grow = row0 + rl, n = grow / H, yy = grow - n * H;
    const size_t off = ((n * C) * (size_t)H + yy) * W + px;           // channel 0 of this pixel in an NCHW tensor
    float g[C2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const float* src = (t ? panf : msf) + off;
      float xin[C];
#pragma unroll
      for (int k = 0; k < C; ++k) xin[k] = __ldg(src + (size_t)k * H * W);
#pragma unroll
      for (int co = 0; co < C; ++co) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) a = fmaf(sw[t][co][k], xin[k], a);
        g[t * C + co] = (a + sb[t][co]) + 1e-8f;
      }
    }
#pragma unroll
    for (int f = 0; f < NF1; ++f) X[(rl * NF1 + f) * RP + padded_m<M>(px)] = make_float2(g[2 * f], g[2 * f + 1]);
  }
  rows_fwd_transform<C2, M>(X, tw, spec, row0);
}

// Inverse row pass + the rest of LGMixer: y = proj(cat(local, |irfft|)) + xres.  The c x c projection runs on the
// tensor cores (split-fp16 3xMMA, see tc_ptx.cuh): a 128-pixel tile of the concatenated map is written to shared memory
// as the A operand, the accumulator comes back from TMEM for the bias + residual epilogue.  (The CUDA-core matvec this
// replaces was 60 % of the kernel and stalled on shared-memory latency.)
template <int C2, int M>
__global__ void __launch_bounds__(256) fft_rows_inv256_kernel(const float2* __restrict__ spec, const float* __restrict__ local,
                                                               const float* __restrict__ xres, float* __restrict__ y,
                                                               BlockW w, float scale, int staged) {
  using namespace f256;
  using namespace tc;
  constexpr int W = 16 * M, Wf = W / 2 + 1, RP = 16 * (M + 1), NSEQ = 256 / M, KPT = 16 / M, TS = 16 / M;
  constexpr int NF1 = C2 / 2, RW = NSEQ / NF1, C = 2 * C2;
  static_assert(RW >= 1 && (RW * W) % 256 == 0, "a CTA holds whole rows and an even number of 128-pixel tiles");
  constexpr int TILES = RW * W / 128;                     // 128-pixel tiles per CTA, two in flight (warps 0-3 / 4-7)
  constexpr uint32_t TCOLS = (2 * C < 32) ? 32 : 2 * C;   // TMEM columns: two accumulators of C columns
  __shared__ float2 tw[256];
  __shared__ uint64_t mbar[2];                            // one MMA-completion barrier per tile in flight
  __shared__ uint32_t tmem_slot;
  extern __shared__ __align__(16) float2 smf[];
  float2* X = smf;                                        // [NSEQ][RP]
  float2* E = smf;                                        // exchange, same storage
  __half* wh = reinterpret_cast<__half*>(smf + NSEQ * RP);      // proj weight hi [C/8][C][8], then lo
  __half* wl = wh + C * C;
  __half* ah = wl + C * C;                                // A operand: [2 tiles][hi | lo][C/8][128][8]
  float* sBias = reinterpret_cast<float*>(ah + 2 * 2 * 128 * C);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t row0 = (size_t)blockIdx.x * RW;
  tw[tid] = g_tw256[tid];
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_slot, TCOLS);
  for (int i = tid; i < 2 * C * C * 2 / 16; i += 256) reinterpret_cast<uint4*>(wh)[i] = __ldg(reinterpret_cast<const uint4*>(w.proj_pack) + i);
  for (int i = tid; i < C; i += 256) sBias[i] = __ldg(w.proj_b + i);
  // phase 0: Hermitian rebuild of the packed spectra (C2R ignores Im of the DC and Nyquist bins)
  const float4* in = reinterpret_cast<const float4*>(spec + row0 * Wf * C2);
  {
    // all of this thread's spectrum loads are issued before the first one is consumed (the loop form waited a full HBM
    // round trip per iteration: 18 % of the kernel's stall samples)
    constexpr int TOTAL = RW * Wf * NF1, ITERS = (TOTAL + 255) / 256;
    float4 vb[ITERS];
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int id = tid + 256 * i;
      vb[i] = (id < TOTAL) ? __ldg(in + id) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int id = tid + 256 * i;
      if (id >= TOTAL) break;
      const int rl = id / (Wf * NF1), rem = id - rl * (Wf * NF1);
      const int k = rem / NF1, s = rl * NF1 + (rem - k * NF1);
      const float4 v = vb[i];                             // (Xa.re, Xa.im, Xb.re, Xb.im)
      if (k == 0 || k == W / 2) {
        X[s * RP + padded_m<M>(k)] = make_float2(v.x, v.z);
      } else {
        X[s * RP + padded_m<M>(k)] = make_float2(v.x - v.w, v.y + v.z);           // Xa + i Xb
        X[s * RP + padded_m<M>(W - k)] = make_float2(v.x + v.w, v.z - v.y);       // conj(Xa) + i conj(Xb)
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int seq = tid / M, lo = tid % M;
  float2 v[16];
#pragma unroll
  for (int m1 = 0; m1 < 16; ++m1) v[m1] = X[seq * RP + (M + 1) * m1 + lo];
  fft16<+1>(v);
#pragma unroll
  for (int j1 = 1; j1 < 16; ++j1) v[j1] = ctw<+1>(v[j1], tw[TS * lo * j1]);
  __syncwarp();                                           // a sequence = M lanes of one warp + its own region: warp-level hand-offs
#pragma unroll
  for (int j1 = 0; j1 < 16; ++j1) E[seq * RP + j1 * (M + 1) + lo] = v[j1];
  __syncwarp();
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk) {
#pragma unroll
    for (int m2 = 0; m2 < M; ++m2) v[kk * M + m2] = E[seq * RP + (lo + M * kk) * (M + 1) + m2];
    fft_m<M, +1>(v + kk * M);                             // v[kk*M + j2] = (xa + i xb)[(lo + M kk) + 16 j2], unnormalised
  }
  __syncwarp();
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk)
#pragma unroll
    for (int j2 = 0; j2 < M; ++j2)
      X[seq * RP + (lo + M * kk) + 16 * j2] = make_float2(fabsf(v[kk * M + j2].x * scale), fabsf(v[kk * M + j2].y * scale));
  __syncthreads();
  // phase 3: concat(local, global) -> proj 1x1 (tensor cores) -> + bias + residual
  const int t = warp >> 2;                                // which of the two tiles in flight
  const int trow = (warp & 3) * 32 + lane;                // row of the tile = TMEM lane
  __half* a_hi = ah + t * (2 * 128 * C);
  __half* a_lo = a_hi + 128 * C;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + t * C;
  uint32_t phase = 0;
#pragma unroll 1
  for (int round = 0; round < TILES / 2; ++round) {
    const int p = (round * 2 + t) * 128 + trow, rl = p / W, px = p % W;
    const size_t pix = (row0 + rl) * W + px;
    {
      const float4* lsrc = reinterpret_cast<const float4*>(local + pix * C2);
#pragma unroll
      for (int kc = 0; kc < C2 / 8; ++kc) {               // local half: channels [0, C2)
        const float4 t0 = __ldg(lsrc + 2 * kc), t1 = __ldg(lsrc + 2 * kc + 1);
        const float2 q[4] = {make_float2(t0.x, t0.y), make_float2(t0.z, t0.w), make_float2(t1.x, t1.y), make_float2(t1.z, t1.w)};
        uint4 hi, lo4;
        split8(q, hi, lo4);
        *reinterpret_cast<uint4*>(a_hi + (kc * 128 + trow) * 8) = hi;
        *reinterpret_cast<uint4*>(a_lo + (kc * 128 + trow) * 8) = lo4;
      }
#pragma unroll
      for (int kc = 0; kc < C2 / 8; ++kc) {               // global half: channels [C2, C), pairs (2f, 2f+1) per sequence
        float2 q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = X[(rl * NF1 + 4 * kc + i) * RP + px];
        uint4 hi, lo4;
        split8(q, hi, lo4);
        *reinterpret_cast<uint4*>(a_hi + ((C2 / 8 + kc) * 128 + trow) * 8) = hi;
        *reinterpret_cast<uint4*>(a_lo + ((C2 / 8 + kc) * 128 + trow) * 8) = lo4;
      }
    }
    // the two tiles in flight (warps 0-3 / 4-7) run free of each other: a 128-thread named barrier publishes a tile's A
    // operand, its own thread issues the MMAs and its own mbarrier reports completion
    fence_proxy_async();
    tc_fence_before();
    if (t == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
    if ((warp & 3) == 0 && elect_one()) {     // one lane of the tile's first warp (elect.sync: no divergence loop around the MMAs)
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc(C);
      const uint32_t ab = smem_u32(a_hi);
#pragma unroll
      for (int ks = 0; ks < C / 16; ++ks) {
        const uint64_t dah = umma_desc(ab + ks * 2 * 128 * 16, 128 * 16, 128);
        const uint64_t dal = umma_desc(ab + 128 * C * 2 + ks * 2 * 128 * 16, 128 * 16, 128);
        const uint64_t dbh = umma_desc(smem_u32(wh) + ks * 2 * C * 16, C * 16, 128);
        const uint64_t dbl = umma_desc(smem_u32(wl) + ks * 2 * C * 16, C * 16, 128);
        umma_f16(tmem + t * C, dah, dbh, idesc, ks > 0);
        umma_f16(tmem + t * C, dah, dbl, idesc, 1);
        umma_f16(tmem + t * C, dal, dbh, idesc, 1);
      }
      umma_commit(&mbar[t]);
    }
    mbar_wait(&mbar[t], phase);
    phase ^= 1;
    tc_fence_after();
    // Epilogue through shared memory: a TMEM lane is a pixel, so adding the residual and storing from here would make every
    // lane touch its own 64 - 256-byte row (16 - 32 L1 wavefronts per instruction; the L1 data pipe bounded the c = 32 form).
    // The accumulator + bias goes into the tile's own A-operand buffer (free once its MMAs have completed, and exactly
    // 128 x C floats), 16-byte pieces XOR-swizzled by the row so that the row-wise writes and the linear reads are both
    // conflict-free; the residual read and the store then run over the tile's contiguous 128 x C floats, lane = piece.
    if (!staged) {                                          // direct form: the TMEM lane's thread adds the residual and stores its own row
      const float* xr = xres + pix * C;
      float* dst = y + pix * C;
      const bool al32 = (reinterpret_cast<uintptr_t>(y) & 31) == 0;
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 8) {
        float2 acc[4];
        tmem_ld8(lane_addr + c0, acc);
        tmem_ld_wait();
        const float4 r0 = __ldg(reinterpret_cast<const float4*>(xr + c0)), r1 = __ldg(reinterpret_cast<const float4*>(xr + c0 + 4));
        const float4 b0 = *reinterpret_cast<const float4*>(sBias + c0), b1 = *reinterpret_cast<const float4*>(sBias + c0 + 4);
        const float4 y0 = make_float4((acc[0].x + b0.x) + r0.x, (acc[0].y + b0.y) + r0.y, (acc[1].x + b0.z) + r0.z, (acc[1].y + b0.w) + r0.w);
        const float4 y1 = make_float4((acc[2].x + b1.x) + r1.x, (acc[2].y + b1.y) + r1.y, (acc[3].x + b1.z) + r1.z, (acc[3].y + b1.w) + r1.w);
        if (al32) {                                 // one 32-byte store = one full L2 sector per thread and instruction
          asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + c0), "f"(y0.x), "f"(y0.y), "f"(y0.z), "f"(y0.w),
                       "f"(y1.x), "f"(y1.y), "f"(y1.z), "f"(y1.w) : "memory");
        } else {
          *reinterpret_cast<float4*>(dst + c0) = y0;
          *reinterpret_cast<float4*>(dst + c0 + 4) = y1;
        }
      }
    } else
    {
      float* stage = reinterpret_cast<float*>(a_hi);
      const int swz = (C == 16) ? ((trow >> 1) & 3) : (trow & 7);
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 8) {
        float2 acc[4];
        tmem_ld8(lane_addr + c0, acc);
        tmem_ld_wait();
        const float4 b0 = *reinterpret_cast<const float4*>(sBias + c0), b1 = *reinterpret_cast<const float4*>(sBias + c0 + 4);
        *reinterpret_cast<float4*>(stage + trow * C + (((c0 >> 2) ^ swz) << 2)) =
            make_float4(acc[0].x + b0.x, acc[0].y + b0.y, acc[1].x + b0.z, acc[1].y + b0.w);
        *reinterpret_cast<float4*>(stage + trow * C + ((((c0 >> 2) + 1) ^ swz) << 2)) =
            make_float4(acc[2].x + b1.x, acc[2].y + b1.y, acc[3].x + b1.z, acc[3].y + b1.w);
      }
      tc_fence_before();
      if (t == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      const size_t pix0 = (row0 + (size_t)((round * 2 + t) * 128) / W) * W + ((round * 2 + t) * 128) % W;   // the tile's first pixel
      const float4* xr4 = reinterpret_cast<const float4*>(xres + pix0 * C);
      float4* y4 = reinterpret_cast<float4*>(y + pix0 * C);
      constexpr int PPR = C / 4;                            // 16-byte pieces per pixel row
#pragma unroll
      for (int i = 0; i < PPR; ++i) {
        const int idx = i * 128 + trow, r = idx / PPR, pc = idx % PPR;
        const int sw = (C == 16) ? ((r >> 1) & 3) : (r & 7);
        const float4 a = *reinterpret_cast<const float4*>(stage + r * C + ((pc ^ sw) << 2));
        const float4 x4 = __ldg(xr4 + idx);
        y4[idx] = make_float4(a.x + x4.x, a.y + x4.y, a.z + x4.z, a.w + x4.w);
      }
      if (round + 1 < TILES / 2) {                          // the next round's A operand overwrites the staging tile
        if (t == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

template <int C2, int M>
static cudaError_t rows256_fwd_t(const BlockW& w, const float* x, float* spec, int N, int H, cudaStream_t s) {
  constexpr int RW = (256 / M) / (C2 / 2);
  if (H % RW) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(256 / M) * 16 * (M + 1) * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(fft_rows_fwd256_kernel<C2, true, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fft_rows_fwd256_kernel<C2, true, M><<<N * H / RW, 256, smem, s>>>(x, reinterpret_cast<float2*>(spec), w);
  return cudaGetLastError();
}
template <int C2, int M>
static cudaError_t rows256_inv_t(const BlockW& w, const float* spec, const float* local, const float* xres, float* y, int N,
                                 int H, cudaStream_t s) {
  constexpr int RW = (256 / M) / (C2 / 2), C = 2 * C2;
  if (H % RW) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(256 / M) * 16 * (M + 1) * sizeof(float2) + (size_t)(2 * C * C + 2 * 2 * 128 * C) * 2 + (size_t)C * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(fft_rows_inv256_kernel<C2, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // staged epilogue (coalesced residual read + store through the tile's A buffer) by default: forward 13.67 -> 13.49 ms (GF-2, 64 pairs),
  // 18.46 -> 17.69 ms (WV-3, 32 pairs); LGTEUN_ROWS_INV_EPI=direct keeps the lane-per-pixel form for A/B runs
  static const int epi = [] { const char* e = getenv("LGTEUN_ROWS_INV_EPI"); return !e ? -1 : (std::string(e) == "staged" ? 1 : 0); }();
  const int staged = epi >= 0 ? epi : 1;
  fft_rows_inv256_kernel<C2, M><<<N * H / RW, 256, smem, s>>>(reinterpret_cast<const float2*>(spec), local, xres, y, w,
                                                              1.0f / ((float)H * (float)(16 * M)), staged);
  return cudaGetLastError();
}

// ---- plain passes for the companion operator SFIIN.Freprocess (companion_ops.cu) -----------------------
template <int M, int MODE>
static cudaError_t cols_plain_t(int c2, float* spec, int N, int W, cudaStream_t s) {
  constexpr int Q = 256 / M;
  const int lanes = (W / 2 + 1) * c2;
  const size_t smem = (size_t)16 * M * Q * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(fft_cols256_kernel<Q, M, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((lanes + Q - 1) / Q, N);
  fft_cols256_kernel<Q, M, MODE><<<grid, M * Q, smem, s>>>(reinterpret_cast<float2*>(spec), BlockW{}, W, c2, lanes);
  return cudaGetLastError();
}
// in-place column FFT of spec[n][H][W/2+1][c2] for H in {128, 256}: dir < 0 forward (the four real bins get +0.0 imaginary
// parts), dir > 0 unnormalised inverse
cudaError_t launch_fft_cols_plain(int H, int c2, float* spec, int N, int W, int dir, cudaStream_t s, int fixreal) {
  if (H == 256) return dir > 0 ? cols_plain_t<16, 2>(c2, spec, N, W, s) : fixreal ? cols_plain_t<16, 1>(c2, spec, N, W, s) : cols_plain_t<16, 3>(c2, spec, N, W, s);
  if (H == 128) return dir > 0 ? cols_plain_t<8, 2>(c2, spec, N, W, s) : fixreal ? cols_plain_t<8, 1>(c2, spec, N, W, s) : cols_plain_t<8, 3>(c2, spec, N, W, s);
  return cudaErrorInvalidValue;
}
// The whole column stage of Freprocess in one kernel (models/SFIIN.py:223-234 between the row transforms): forward FFT
// along H of the 2C-channel spectrum S[n][ky][kx][2C], amp_fuse / pha_fuse per bin, inverse FFT along H of the C fused
// channels into G[n][ky][kx][C].  A CTA owns KX = Q / 2C adjacent kx (Q = 256 / M lanes, all 2C channels of a bin inside the
// CTA); the forward transform is the two-pass register scheme of fft_cols256_kernel, the spectrum then goes through shared
// memory ([ky][Q + 1], one pad slot per row) so that one thread sees the 2C channels of a bin, the fused C channels come back
// the same way and the threads of the first Q / 2 lanes run the inverse transform.
struct FreFuseW {
  const float *a0w, *a0b, *a2w, *a2b, *p0w, *p0b, *p2w, *p2b;
};
template <int C, int M>
__global__ void __launch_bounds__(256) fre_cols_fused_kernel(const float2* __restrict__ S, float2* __restrict__ G, FreFuseW fw,
                                                             int W) {
  using namespace f256;
  constexpr int H = 16 * M, Q = 256 / M, C2 = 2 * C, KX = Q / C2, KPT = 16 / M, TS = 16 / M, QP = Q + 1;
  static_assert(KX >= 1, "all 2C channels of a bin live in one CTA");
  __shared__ float2 tw[256];
  __shared__ __align__(16) float w0[2][C][C2];           // 16-byte rows: the broadcast weight reads become LDS.128
  __shared__ __align__(16) float w2[2][C][C];
  __shared__ float bb[2][2][C];
  extern __shared__ __align__(16) float2 ex[];            // [H][Q + 1]; the first H * Q slots double as the pass exchange
  const int tid = threadIdx.x;
  const int l = tid % Q, lo = tid / Q;
  const int Wh = W / 2 + 1, kx0 = blockIdx.x * KX;
  const bool live = kx0 + l / C2 < Wh;
  const size_t in_row = (size_t)Wh * C2, out_row = (size_t)Wh * C;
  const float2* base = S + (size_t)blockIdx.y * H * in_row + (size_t)kx0 * C2 + l;
  tw[tid] = g_tw256[tid];
  for (int i = tid; i < 2 * C * C2; i += 256) {
    const int t = i / (C * C2), r = i % (C * C2);
    w0[t][r / C2][r % C2] = __ldg((t ? fw.p0w : fw.a0w) + r);
  }
  for (int i = tid; i < 2 * C * C; i += 256) {
    const int t = i / (C * C), r = i % (C * C);
    w2[t][r / C][r % C] = __ldg((t ? fw.p2w : fw.a2w) + r);
  }
  for (int i = tid; i < 4 * C; i += 256) {
    const int t = i / (2 * C), ly = (i / C) & 1, k = i % C;
    bb[t][ly][k] = __ldg((t ? (ly ? fw.p2b : fw.p0b) : (ly ? fw.a2b : fw.a0b)) + k);
  }
  float2 v[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) v[n1] = live ? __ldg(base + (size_t)(M * n1 + lo) * in_row) : make_float2(0.f, 0.f);
  __syncthreads();                                        // tables staged
  fft16<-1>(v);
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) v[k1] = ctw<-1>(v[k1], tw[TS * lo * k1]);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) ex[(k1 * M + lo) * Q + l] = v[k1];
  __syncthreads();
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk) {
#pragma unroll
    for (int n2 = 0; n2 < M; ++n2) v[kk * M + n2] = ex[((lo + M * kk) * M + n2) * Q + l];
    fft_m<M, -1>(v + kk * M);                             // v[kk*M + k2] = X[ky = (lo + M kk) + 16 k2]
  }
  __syncthreads();                                        // everyone has consumed the pass exchange
  {
    const int kx = kx0 + l / C2;
    const bool real_col = (kx == 0 || kx == W / 2);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int kk = i / M, k2 = i % M;
      if (real_col && lo == 0 && kk == 0 && (k2 == 0 || k2 == M / 2)) v[i].y = 0.0f;      // exactly-real bins (F7)
      ex[((lo + M * kk) + 16 * k2) * QP + l] = v[i];
    }
  }
  __syncthreads();
  // amp_fuse / pha_fuse: thread = bin (ky, kx); the C fused values overwrite the first C slots of the bin's own 2C inputs
  for (int b = tid; b < H * KX; b += 256) {
    const int ky = b / KX, kxi = b % KX;
    float2* bin = ex + ky * QP + kxi * C2;
    float in[2][C2];
#pragma unroll
    for (int c = 0; c < C2; ++c) {
      const float2 z = bin[c];
      in[0][c] = sqrtf(z.x * z.x + z.y * z.y);
      in[1][c] = atan2f(z.y, z.x);
    }
    float out[2][C];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      float hdn[C];
#pragma unroll
      for (int co = 0; co < C; ++co) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < C2; ++k) a = fmaf(w0[t][co][k], in[t][k], a);
        a += bb[t][0][co];
        hdn[co] = a > 0.f ? a : 0.1f * a;                 // LeakyReLU(0.1)
      }
#pragma unroll
      for (int co = 0; co < C; ++co) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) a = fmaf(w2[t][co][k], hdn[k], a);
        out[t][co] = a + bb[t][1][co];
      }
    }
#pragma unroll
    for (int co = 0; co < C; ++co) {
      float sn, cs;
      sincosf(out[1][co], &sn, &cs);
      bin[co] = make_float2((out[0][co] * cs + 1e-8f) + 1e-8f, out[0][co] * sn + 1e-8f);
    }
  }
  __syncthreads();
  // inverse along H for the Q / 2 fused lanes (kx, co)
  const bool inv = l < Q / 2;
  const int kxi = l / C, co = l % C;
  const bool live_o = inv && (kx0 + kxi < Wh);
  if (inv) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = ex[((lo + M * (i / M)) + 16 * (i % M)) * QP + kxi * C2 + co];
  }
  __syncthreads();                                        // the bin buffer becomes the pass exchange again
  if (inv) {
#pragma unroll
    for (int kk = 0; kk < KPT; ++kk) {
      const int k1 = lo + M * kk;
      fft_m<M, +1>(v + kk * M);
#pragma unroll
      for (int j1 = 1; j1 < M; ++j1) v[kk * M + j1] = ctw<+1>(v[kk * M + j1], tw[TS * k1 * j1]);
#pragma unroll
      for (int j1 = 0; j1 < M; ++j1) ex[(j1 * 16 + k1) * Q + l] = v[kk * M + j1];
    }
  }
  __syncthreads();
  if (inv) {
#pragma unroll
    for (int m2 = 0; m2 < 16; ++m2) v[m2] = ex[(lo * 16 + m2) * Q + l];
    fft16<+1>(v);                                         // v[j2] = x[y = lo + M j2]  (unnormalised)
    if (live_o) {
      float2* dst = G + (size_t)blockIdx.y * H * out_row + (size_t)(kx0 + kxi) * C + co;
#pragma unroll
      for (int j2 = 0; j2 < 16; ++j2) dst[(size_t)(lo + M * j2) * out_row] = v[j2];
    }
  }
}
template <int C, int M>
static cudaError_t fre_cols_fused_t(const float* S, float* G, const FreFuseW& fw, int N, int W, cudaStream_t s) {
  constexpr int Q = 256 / M, KX = Q / (2 * C), H = 16 * M;
  const size_t smem = (size_t)H * (Q + 1) * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(fre_cols_fused_kernel<C, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((W / 2 + 1 + KX - 1) / KX, N);
  fre_cols_fused_kernel<C, M><<<grid, 256, smem, s>>>(reinterpret_cast<const float2*>(S), reinterpret_cast<float2*>(G), fw, W);
  return cudaGetLastError();
}
// returns cudaErrorNotSupported when (H, C) has no fused instantiation (the caller then runs the three separate passes)
cudaError_t launch_fre_cols_fused(int H, int C, const float* S, float* G, const float* const* fuse_w, int N, int W, cudaStream_t s) {
  const FreFuseW fw{fuse_w[0], fuse_w[1], fuse_w[2], fuse_w[3], fuse_w[4], fuse_w[5], fuse_w[6], fuse_w[7]};
  if (H == 256) {
    if (C == 4) return fre_cols_fused_t<4, 16>(S, G, fw, N, W, s);
    if (C == 8) return fre_cols_fused_t<8, 16>(S, G, fw, N, W, s);
  } else if (H == 128) {
    if (C == 4) return fre_cols_fused_t<4, 8>(S, G, fw, N, W, s);
    if (C == 8) return fre_cols_fused_t<8, 8>(S, G, fw, N, W, s);
    if (C == 16) return fre_cols_fused_t<16, 8>(S, G, fw, N, W, s);
  }
  return cudaErrorNotSupported;
}

template <int C2, int M>
static cudaError_t rows_fwd_pre_t(const float* msf, const float* panf, const float* const* pre_w, float* spec, int N, int H,
                                  cudaStream_t s) {
  constexpr int RW = (256 / M) / (C2 / 2);
  if (H % RW) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(256 / M) * 16 * (M + 1) * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(fft_rows_fwd_pre_kernel<C2, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fft_rows_fwd_pre_kernel<C2, M><<<N * H / RW, 256, smem, s>>>(msf, panf, pre_w[0], pre_w[1], pre_w[2], pre_w[3],
                                                               reinterpret_cast<float2*>(spec), H);
  return cudaGetLastError();
}
// pre1 / pre2 convs + row rFFT of Freprocess: NCHW msf, panf [N][C][H][W] -> spec[N][H][W/2+1][2C];
// pre_w = {pre1.weight, pre1.bias, pre2.weight, pre2.bias}; W in {128, 256}, C in {4, 8, 16}
cudaError_t launch_fft_rows_fwd_pre(int W, int C, const float* msf, const float* panf, const float* const* pre_w, float* spec, int N,
                                    int H, cudaStream_t s) {
  if (W == 256) {
    if (C == 4) return rows_fwd_pre_t<8, 16>(msf, panf, pre_w, spec, N, H, s);
    if (C == 8) return rows_fwd_pre_t<16, 16>(msf, panf, pre_w, spec, N, H, s);
    if (C == 16) return rows_fwd_pre_t<32, 16>(msf, panf, pre_w, spec, N, H, s);
  } else if (W == 128) {
    if (C == 4) return rows_fwd_pre_t<8, 8>(msf, panf, pre_w, spec, N, H, s);
    if (C == 8) return rows_fwd_pre_t<16, 8>(msf, panf, pre_w, spec, N, H, s);
    if (C == 16) return rows_fwd_pre_t<32, 8>(msf, panf, pre_w, spec, N, H, s);
  }
  return cudaErrorInvalidValue;
}
// Inverse row pass of Freprocess (models/SFIIN.py:235-236): C2R of spec[N][H][W/2+1][C] (the same register transform as
// fft_rows_inv256_kernel), |.| * scale, then the `post` 1x1 conv C -> C and the NCHW store — the |irfft2| map stays in
// shared memory.
template <int C, int M>
__global__ void __launch_bounds__(256) fft_rows_inv_post_kernel(const float2* __restrict__ spec, const float* __restrict__ pw,
                                                                 const float* __restrict__ pb, float* __restrict__ y, int H,
                                                                 float scale) {
  using namespace f256;
  constexpr int W = 16 * M, Wf = W / 2 + 1, RP = 16 * (M + 1), NSEQ = 256 / M, KPT = 16 / M, TS = 16 / M;
  constexpr int NF1 = C / 2, RW = NSEQ / NF1;
  static_assert(RW >= 1, "a CTA holds at least one image row");
  __shared__ float2 tw[256];
  __shared__ __align__(16) float sw[C][C];
  __shared__ float sb[C];
  extern __shared__ __align__(16) float2 smf[];
  float2* X = smf;                                        // [NSEQ][RP]
  float2* E = smf;
  const int tid = threadIdx.x;
  const size_t row0 = (size_t)blockIdx.x * RW;
  tw[tid] = g_tw256[tid];
  for (int i = tid; i < C * C; i += 256) sw[i / C][i % C] = __ldg(pw + i);
  if (tid < C) sb[tid] = __ldg(pb + tid);
  const float4* in = reinterpret_cast<const float4*>(spec + row0 * Wf * C);
  {
    constexpr int TOTAL = RW * Wf * NF1, ITERS = (TOTAL + 255) / 256;
    float4 vb[ITERS];
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int id = tid + 256 * i;
      vb[i] = (id < TOTAL) ? __ldg(in + id) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int id = tid + 256 * i;
      if (id >= TOTAL) break;
      const int rl = id / (Wf * NF1), rem = id - rl * (Wf * NF1);
      const int k = rem / NF1, sq = rl * NF1 + (rem - k * NF1);
      const float4 v = vb[i];                             // (Xa.re, Xa.im, Xb.re, Xb.im)
      if (k == 0 || k == W / 2) {
        X[sq * RP + padded_m<M>(k)] = make_float2(v.x, v.z);
      } else {
        X[sq * RP + padded_m<M>(k)] = make_float2(v.x - v.w, v.y + v.z);          // Xa + i Xb
        X[sq * RP + padded_m<M>(W - k)] = make_float2(v.x + v.w, v.z - v.y);      // conj(Xa) + i conj(Xb)
      }
    }
  }
  __syncthreads();
  const int seq = tid / M, lo = tid % M;
  float2 v[16];
#pragma unroll
  for (int m1 = 0; m1 < 16; ++m1) v[m1] = X[seq * RP + (M + 1) * m1 + lo];
  fft16<+1>(v);
#pragma unroll
  for (int j1 = 1; j1 < 16; ++j1) v[j1] = ctw<+1>(v[j1], tw[TS * lo * j1]);
  __syncwarp();
#pragma unroll
  for (int j1 = 0; j1 < 16; ++j1) E[seq * RP + j1 * (M + 1) + lo] = v[j1];
  __syncwarp();
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk) {
#pragma unroll
    for (int m2 = 0; m2 < M; ++m2) v[kk * M + m2] = E[seq * RP + (lo + M * kk) * (M + 1) + m2];
    fft_m<M, +1>(v + kk * M);
  }
  __syncwarp();
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk)
#pragma unroll
    for (int j2 = 0; j2 < M; ++j2)
      X[seq * RP + (lo + M * kk) + 16 * j2] = make_float2(fabsf(v[kk * M + j2].x * scale), fabsf(v[kk * M + j2].y * scale));
  __syncthreads();
  // post conv, thread = pixel; NCHW planes are written with consecutive threads on consecutive pixels
#pragma unroll
  for (int i = 0; i < RW * W / 256; ++i) {
    const int p = tid + 256 * i, rl = p / W, px = p % W;
    const size_t grow = row0 + rl, n = grow / H, yy = grow - n * H;
    float g[C];
#pragma unroll
    for (int f = 0; f < NF1; ++f) {
      const float2 t = X[(rl * NF1 + f) * RP + px];
      g[2 * f] = t.x;
      g[2 * f + 1] = t.y;
    }
    float* dst = y + ((n * C) * (size_t)H + yy) * W + px;
#pragma unroll
    for (int co = 0; co < C; ++co) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) a = fmaf(sw[co][k], g[k], a);
      dst[(size_t)co * H * W] = a + sb[co];
    }
  }
}
template <int C, int M>
static cudaError_t rows_inv_post_t(const float* spec, const float* pw, const float* pb, float* y, int N, int H, cudaStream_t s) {
  constexpr int RW = (256 / M) / (C / 2);
  if (H % RW) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(256 / M) * 16 * (M + 1) * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(fft_rows_inv_post_kernel<C, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fft_rows_inv_post_kernel<C, M><<<N * H / RW, 256, smem, s>>>(reinterpret_cast<const float2*>(spec), pw, pb, y, H,
                                                               1.0f / ((float)H * (float)(16 * M)));
  return cudaGetLastError();
}
cudaError_t launch_fft_rows_inv_post(int W, int C, const float* spec, const float* post_w, const float* post_b, float* y_nchw, int N,
                                     int H, cudaStream_t s) {
  if (W == 256) {
    if (C == 4) return rows_inv_post_t<4, 16>(spec, post_w, post_b, y_nchw, N, H, s);
    if (C == 8) return rows_inv_post_t<8, 16>(spec, post_w, post_b, y_nchw, N, H, s);
    if (C == 16) return rows_inv_post_t<16, 16>(spec, post_w, post_b, y_nchw, N, H, s);
  } else if (W == 128) {
    if (C == 4) return rows_inv_post_t<4, 8>(spec, post_w, post_b, y_nchw, N, H, s);
    if (C == 8) return rows_inv_post_t<8, 8>(spec, post_w, post_b, y_nchw, N, H, s);
    if (C == 16) return rows_inv_post_t<16, 8>(spec, post_w, post_b, y_nchw, N, H, s);
  }
  return cudaErrorInvalidValue;
}

// ---- plain row passes of the training step (train.cu; the adjoints are the same two transforms with other weights) -------------
// R2C: real rows x[pixel * ldx + c] (optionally times sign(sgn[pixel * C2 + c]): the backward of |.|) -> spec[row][k <= W/2][C2],
// times scale, interior bins (0 < k < W/2) also times wint.  wint = 2 makes it the adjoint of the C2R below.
template <int C2, int M>
__global__ void __launch_bounds__(256) fft_rows_r2c_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ sgn,
                                                            float2* __restrict__ spec, float scale, float wint) {
  using namespace f256;
  constexpr int W = 16 * M, RP = 16 * (M + 1), NF1 = C2 / 2, RW = (256 / M) / NF1;
  static_assert(RW >= 1, "a CTA holds at least one image row");
  __shared__ float2 tw[256];
  extern __shared__ __align__(16) float2 smf[];
  float2* X = smf;
  const int tid = threadIdx.x;
  const size_t row0 = (size_t)blockIdx.x * RW;
  tw[tid] = g_tw256[tid];
#pragma unroll
  for (int i = 0; i < RW * W / 256; ++i) {
    const int p = tid + 256 * i, rl = p / W, px = p % W;
    const size_t pix = (row0 + rl) * W + px;
    float g[C2];
    load_vec<C2>(g, x + pix * ldx);
    if (sgn) {
      float t[C2];
      load_vec<C2>(t, sgn + pix * C2);
#pragma unroll
      for (int c = 0; c < C2; ++c) g[c] = t[c] > 0.f ? g[c] : (t[c] < 0.f ? -g[c] : 0.f);
    }
#pragma unroll
    for (int f = 0; f < NF1; ++f) X[(rl * NF1 + f) * RP + padded_m<M>(px)] = make_float2(g[2 * f], g[2 * f + 1]);
  }
  rows_fwd_transform<C2, M>(X, tw, spec, row0, 0.5f * scale, wint);
}
// C2R: spec[row][k <= W/2][C2] (interior bins times wint; the imaginary parts of bins 0 and W/2 drop out) -> real rows times
// scale -> xout[pixel * ldo + c], and |.| of it -> xabs[pixel * ldabs + c].  wint = 1: irfft's C2R; wint = 1/2: the real part
// of the inverse transform of the zero-padded half spectrum (the adjoint of the plain R2C).
template <int C2, int M>
__global__ void __launch_bounds__(256) fft_rows_c2r_kernel(const float2* __restrict__ spec, float* __restrict__ xout, int ldo,
                                                            float* __restrict__ xabs, int ldabs, float scale, float wint) {
  using namespace f256;
  constexpr int W = 16 * M, Wf = W / 2 + 1, RP = 16 * (M + 1), NSEQ = 256 / M, KPT = 16 / M, TS = 16 / M;
  constexpr int NF1 = C2 / 2, RW = NSEQ / NF1;
  static_assert(RW >= 1, "a CTA holds at least one image row");
  __shared__ float2 tw[256];
  extern __shared__ __align__(16) float2 smf[];
  float2* X = smf;                                        // [NSEQ][RP]
  float2* E = smf;
  const int tid = threadIdx.x;
  const size_t row0 = (size_t)blockIdx.x * RW;
  tw[tid] = g_tw256[tid];
  const float4* in = reinterpret_cast<const float4*>(spec + row0 * Wf * C2);
  {
    constexpr int TOTAL = RW * Wf * NF1, ITERS = (TOTAL + 255) / 256;
    float4 vb[ITERS];
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int id = tid + 256 * i;
      vb[i] = (id < TOTAL) ? __ldg(in + id) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int id = tid + 256 * i;
      if (id >= TOTAL) break;
      const int rl = id / (Wf * NF1), rem = id - rl * (Wf * NF1);
      const int k = rem / NF1, sq = rl * NF1 + (rem - k * NF1);
      float4 v = vb[i];                                   // (Xa.re, Xa.im, Xb.re, Xb.im)
      if (k == 0 || k == W / 2) {
        X[sq * RP + padded_m<M>(k)] = make_float2(v.x, v.z);
      } else {
        v.x *= wint; v.y *= wint; v.z *= wint; v.w *= wint;
        X[sq * RP + padded_m<M>(k)] = make_float2(v.x - v.w, v.y + v.z);          // Xa + i Xb
        X[sq * RP + padded_m<M>(W - k)] = make_float2(v.x + v.w, v.z - v.y);      // conj(Xa) + i conj(Xb)
      }
    }
  }
  __syncthreads();
  const int seq = tid / M, lo = tid % M;
  float2 v[16];
#pragma unroll
  for (int m1 = 0; m1 < 16; ++m1) v[m1] = X[seq * RP + (M + 1) * m1 + lo];
  fft16<+1>(v);
#pragma unroll
  for (int j1 = 1; j1 < 16; ++j1) v[j1] = ctw<+1>(v[j1], tw[TS * lo * j1]);
  __syncwarp();
#pragma unroll
  for (int j1 = 0; j1 < 16; ++j1) E[seq * RP + j1 * (M + 1) + lo] = v[j1];
  __syncwarp();
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk) {
#pragma unroll
    for (int m2 = 0; m2 < M; ++m2) v[kk * M + m2] = E[seq * RP + (lo + M * kk) * (M + 1) + m2];
    fft_m<M, +1>(v + kk * M);
  }
  __syncwarp();
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk)
#pragma unroll
    for (int j2 = 0; j2 < M; ++j2)
      X[seq * RP + (lo + M * kk) + 16 * j2] = make_float2(v[kk * M + j2].x * scale, v[kk * M + j2].y * scale);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RW * W / 256; ++i) {                // thread = pixel
    const int p = tid + 256 * i, rl = p / W, px = p % W;
    const size_t pix = (row0 + rl) * W + px;
    float g[C2];
#pragma unroll
    for (int f = 0; f < NF1; ++f) {
      const float2 t = X[(rl * NF1 + f) * RP + px];
      g[2 * f] = t.x;
      g[2 * f + 1] = t.y;
    }
    store_vec<C2>(xout + pix * ldo, g);
    if (xabs) {
#pragma unroll
      for (int c = 0; c < C2; ++c) g[c] = fabsf(g[c]);
      store_vec<C2>(xabs + pix * ldabs, g);
    }
  }
}
template <int C2, int M>
static cudaError_t rows_r2c_t(const float* x, int ldx, const float* sgn, float* spec, size_t rows, float scale, float wint, cudaStream_t s) {
  constexpr int RW = (256 / M) / (C2 / 2);
  if (rows % RW) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(256 / M) * 16 * (M + 1) * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(fft_rows_r2c_kernel<C2, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fft_rows_r2c_kernel<C2, M><<<(unsigned)(rows / RW), 256, smem, s>>>(x, ldx, sgn, reinterpret_cast<float2*>(spec), scale, wint);
  return cudaGetLastError();
}
template <int C2, int M>
static cudaError_t rows_c2r_t(const float* spec, float* xout, int ldo, float* xabs, int ldabs, size_t rows, float scale, float wint,
                              cudaStream_t s) {
  constexpr int RW = (256 / M) / (C2 / 2);
  if (rows % RW) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(256 / M) * 16 * (M + 1) * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(fft_rows_c2r_kernel<C2, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fft_rows_c2r_kernel<C2, M><<<(unsigned)(rows / RW), 256, smem, s>>>(reinterpret_cast<const float2*>(spec), xout, ldo, xabs, ldabs,
                                                                       scale, wint);
  return cudaGetLastError();
}
bool fft_rows_plain_supported(int W, int c2, int ldx, const void* p0, const void* p1, const void* p2) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return (W == 128 || W == 256) && (c2 == 8 || c2 == 16 || c2 == 32) && ldx % 4 == 0 && al(p0) && al(p1) && al(p2);
}
cudaError_t launch_fft_rows_r2c(int W, int c2, const float* x, int ldx, const float* sgn, float* spec, size_t rows, float scale,
                                float wint, cudaStream_t s) {
#define LG_R2C(CC, MM) if (c2 == CC && W == 16 * MM) return rows_r2c_t<CC, MM>(x, ldx, sgn, spec, rows, scale, wint, s)
  LG_R2C(8, 16); LG_R2C(16, 16); LG_R2C(32, 16); LG_R2C(8, 8); LG_R2C(16, 8); LG_R2C(32, 8);
#undef LG_R2C
  return cudaErrorInvalidValue;
}
cudaError_t launch_fft_rows_c2r(int W, int c2, const float* spec, float* xout, int ldo, float* xabs, int ldabs, size_t rows,
                                float scale, float wint, cudaStream_t s) {
#define LG_C2R(CC, MM) if (c2 == CC && W == 16 * MM) return rows_c2r_t<CC, MM>(spec, xout, ldo, xabs, ldabs, rows, scale, wint, s)
  LG_C2R(8, 16); LG_C2R(16, 16); LG_C2R(32, 16); LG_C2R(8, 8); LG_C2R(16, 8); LG_C2R(32, 8);
#undef LG_C2R
  return cudaErrorInvalidValue;
}

// W == 256, LayerNorm prologue; H must be a multiple of 32 / C2 rows
cudaError_t launch_fft_rows_fwd256(const BlockW& w, int c, const float* x, float* spec, int N, int H, cudaStream_t s) {
  switch (c) {
    case 16: return rows256_fwd_t<8, 16>(w, x, spec, N, H, s);
    case 32: return rows256_fwd_t<16, 16>(w, x, spec, N, H, s);
    case 64: return rows256_fwd_t<32, 16>(w, x, spec, N, H, s);
    default: return cudaErrorInvalidValue;
  }
}
// W == 128
cudaError_t launch_fft_rows_fwd128(const BlockW& w, int c, const float* x, float* spec, int N, int H, cudaStream_t s) {
  switch (c) {
    case 16: return rows256_fwd_t<8, 8>(w, x, spec, N, H, s);
    case 32: return rows256_fwd_t<16, 8>(w, x, spec, N, H, s);
    case 64: return rows256_fwd_t<32, 8>(w, x, spec, N, H, s);
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_fft_rows_inv256(const BlockW& w, int c, const float* spec, const float* local, const float* xres, float* y,
                                   int N, int H, cudaStream_t s) {
  switch (c) {
    case 16: return rows256_inv_t<8, 16>(w, spec, local, xres, y, N, H, s);
    case 32: return rows256_inv_t<16, 16>(w, spec, local, xres, y, N, H, s);
    case 64: return rows256_inv_t<32, 16>(w, spec, local, xres, y, N, H, s);
    default: return cudaErrorInvalidValue;
  }
}
// W == 128
cudaError_t launch_fft_rows_inv128(const BlockW& w, int c, const float* spec, const float* local, const float* xres, float* y,
                                   int N, int H, cudaStream_t s) {
  switch (c) {
    case 16: return rows256_inv_t<8, 8>(w, spec, local, xres, y, N, H, s);
    case 32: return rows256_inv_t<16, 8>(w, spec, local, xres, y, N, H, s);
    case 64: return rows256_inv_t<32, 8>(w, spec, local, xres, y, N, H, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace lg
