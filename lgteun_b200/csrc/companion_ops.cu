// companion_ops.cu — SURVEY.md §8f rank 4: the FFT amplitude / phase operator of a compared network that the reference
// repository ships next to LGTEUN, built from the hot path's own building blocks.
//
//   SFIIN.Freprocess.forward(msf, panf)          models/SFIIN.py:210-236
//     msF  = rfft2(pre1(msf) + 1e-8), panF = rfft2(pre2(panf) + 1e-8)                                   :223-224
//     amp  = amp_fuse(cat(|msF|, |panF|)),  pha = pha_fuse(cat(angle msF, angle panF))                  :225-230
//            (1x1 conv 2C -> C, LeakyReLU(0.1), 1x1 conv C -> C, both evaluated per spectrum bin)
//     out  = post(|irfft2(complex(amp cos pha + 1e-8, amp sin pha + 1e-8) + 1e-8, s=(H, W))|)           :232-236
//
// Layout: the two pre-convolved maps are written NHWC side by side ([N,H,W,2C]: ms channels, then pan channels), so ONE
// real-to-complex transform produces both spectra S[n][ky][kx][2C]; the fusion kernel reads a bin's 2C complex values once,
// runs both two-layer perceptrons in registers with the weights in shared memory and writes the C fused complex values; the
// inverse row pass ends in |.|, the post conv and the NCHW store.  At 128 / 256 points the transforms are the register-
// resident radix-16 passes of the global mixer (fft256.cu: plain forward rows, forward-only / inverse-only columns, inverse
// rows fused with `post`), at other power-of-two sizes the shared-memory passes of train_kernels.cuh.  At 128 / 256 points
// the operator is three launches — pre convs + row rFFT, column FFT + fusion + inverse column FFT, inverse row FFT + |.| +
// post conv — and the two spectra are the only intermediates in HBM (the reference materialises 14 tensors).
// The four purely real bins get an exact +0.0 imaginary part before angle() (what rfft2 delivers).
#include <cuda_runtime.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <string>

#include "ctx.cuh"

namespace lgctx {   // train.cu
cudaError_t launch_fft_rows(cudaStream_t s, int mode, const float* xin, int ldx, float* spec, float* xout, int ldo, float* xabs,
                            int ldabs, int N, int H, int W, int c2, float scale, int weight2);
cudaError_t launch_fft_cols(cudaStream_t s, float* spec, int N, int H, int W, int c2, int dir, int fixreal);
}  // namespace lgctx

namespace lgcomp {

// pre1 / pre2 (SFIIN.py:213-214,223-224): y[n,p,0:C] = W1 msf[n,:,p] + b1 + 1e-8, y[n,p,C:2C] = W2 panf[n,:,p] + b2 + 1e-8.
// thread = pixel: the NCHW plane reads are coalesced across the warp, the 2C outputs leave as float4.
template <int C>
__global__ void __launch_bounds__(256) fre_pre_kernel(const float* __restrict__ msf, const float* __restrict__ panf,
                                                      const float* __restrict__ w1, const float* __restrict__ b1,
                                                      const float* __restrict__ w2, const float* __restrict__ b2,
                                                      float* __restrict__ y, int HW, size_t NP) {
  __shared__ float sw[2][C][C + 1], sb[2][C];
  for (int i = threadIdx.x; i < 2 * C * C; i += 256) {
    const int t = i / (C * C), r = i % (C * C);
    sw[t][r / C][r % C] = __ldg((t ? w2 : w1) + r);
  }
  for (int i = threadIdx.x; i < 2 * C; i += 256) sb[i / C][i % C] = __ldg((i / C ? b2 : b1) + i % C);
  __syncthreads();
  const size_t gp = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (gp >= NP) return;
  const size_t n = gp / HW, p = gp - n * HW;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const float* src = (t ? panf : msf) + n * C * (size_t)HW + p;
    float x[C], o[C];
#pragma unroll
    for (int k = 0; k < C; ++k) x[k] = __ldg(src + (size_t)k * HW);
#pragma unroll
    for (int co = 0; co < C; ++co) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) a = fmaf(sw[t][co][k], x[k], a);
      o[co] = (a + sb[t][co]) + 1e-8f;
    }
    float4* dst = reinterpret_cast<float4*>(y + gp * (2 * C) + t * C);
#pragma unroll
    for (int q = 0; q < C / 4; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  }
}

// amp_fuse / pha_fuse per spectrum bin (SFIIN.py:225-234): thread = bin.  Weights in shared memory:
// layer 0 [C][2C] + bias, layer 2 [C][C] + bias, for the amplitude and the phase perceptron.
template <int C>
__global__ void __launch_bounds__(128) fre_fuse_kernel(const float2* __restrict__ S, float2* __restrict__ G, size_t nbins,
                                                       const float* __restrict__ a0w, const float* __restrict__ a0b,
                                                       const float* __restrict__ a2w, const float* __restrict__ a2b,
                                                       const float* __restrict__ p0w, const float* __restrict__ p0b,
                                                       const float* __restrict__ p2w, const float* __restrict__ p2b) {
  __shared__ float w0[2][C][2 * C], w2[2][C][C], bb[2][2][C];
  for (int i = threadIdx.x; i < 2 * C * 2 * C; i += 128) {
    const int t = i / (2 * C * C), r = i % (2 * C * C);
    w0[t][r / (2 * C)][r % (2 * C)] = __ldg((t ? p0w : a0w) + r);
  }
  for (int i = threadIdx.x; i < 2 * C * C; i += 128) {
    const int t = i / (C * C), r = i % (C * C);
    w2[t][r / C][r % C] = __ldg((t ? p2w : a2w) + r);
  }
  for (int i = threadIdx.x; i < 4 * C; i += 128) {
    const int t = i / (2 * C), l = (i / C) & 1, k = i % C;
    bb[t][l][k] = __ldg((t ? (l ? p2b : p0b) : (l ? a2b : a0b)) + k);
  }
  __syncthreads();
  const size_t bin = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (bin >= nbins) return;
  float in[2][2 * C];     // [0]: amplitudes (ms, pan), [1]: phases
  const float4* src = reinterpret_cast<const float4*>(S + bin * (2 * C));
#pragma unroll
  for (int q = 0; q < C; ++q) {      // one float4 = two complex values
    const float4 v = __ldg(src + q);
    in[0][2 * q] = sqrtf(v.x * v.x + v.y * v.y);
    in[1][2 * q] = atan2f(v.y, v.x);
    in[0][2 * q + 1] = sqrtf(v.z * v.z + v.w * v.w);
    in[1][2 * q + 1] = atan2f(v.w, v.z);
  }
  float out[2][C];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    float hdn[C];
#pragma unroll
    for (int co = 0; co < C; ++co) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 2 * C; ++k) a = fmaf(w0[t][co][k], in[t][k], a);
      a += bb[t][0][co];
      hdn[co] = a > 0.f ? a : 0.1f * a;          // LeakyReLU(0.1)
    }
#pragma unroll
    for (int co = 0; co < C; ++co) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) a = fmaf(w2[t][co][k], hdn[k], a);
      out[t][co] = a + bb[t][1][co];
    }
  }
  float2* dst = G + bin * C;
#pragma unroll
  for (int co = 0; co < C; ++co) {
    float sn, cs;
    sincosf(out[1][co], &sn, &cs);
    dst[co] = make_float2((out[0][co] * cs + 1e-8f) + 1e-8f, out[0][co] * sn + 1e-8f);
  }
}

// post (SFIIN.py:219,236): NHWC |irfft2| map -> 1x1 conv C -> C -> NCHW.  thread = pixel.
template <int C>
__global__ void __launch_bounds__(256) fre_post_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, float* __restrict__ y, int HW, size_t NP) {
  __shared__ float sw[C][C + 1], sb[C];
  for (int i = threadIdx.x; i < C * C; i += 256) sw[i / C][i % C] = __ldg(w + i);
  for (int i = threadIdx.x; i < C; i += 256) sb[i] = __ldg(b + i);
  __syncthreads();
  const size_t gp = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (gp >= NP) return;
  const size_t n = gp / HW, p = gp - n * HW;
  float v[C];
  const float4* src = reinterpret_cast<const float4*>(x + gp * C);
#pragma unroll
  for (int q = 0; q < C / 4; ++q) {
    const float4 t = __ldg(src + q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
  float* dst = y + n * C * (size_t)HW + p;
#pragma unroll
  for (int co = 0; co < C; ++co) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) a = fmaf(sw[co][k], v[k], a);
    dst[(size_t)co * HW] = a + sb[co];
  }
}

struct FreW {
  const float *pre1_w, *pre1_b, *pre2_w, *pre2_b, *a0w, *a0b, *a2w, *a2b, *p0w, *p0b, *p2w, *p2b, *post_w, *post_b;
};

inline size_t align64(size_t f) { return (f + 63) & ~(size_t)63; }

template <int C>
cudaError_t run_freprocess(const float* msf, const float* panf, float* out, int N, int H, int W, const FreW& w, float* ws,
                           cudaStream_t s) {
  const int HW = H * W, Wh = W / 2 + 1;
  const size_t NP = (size_t)N * HW, nbins = (size_t)N * H * Wh;
  float* pre = ws;                                   // [N,H,W,2C]; reused for the |irfft2| map [N,H,W,C]
  float* S = pre + align64(NP * 2 * C);              // complex [N,H,Wh,2C]
  float* G = S + align64(nbins * 2 * C * 2);         // complex [N,H,Wh,C]
  // register-resident passes of fft256.cu; a CTA of the row kernels owns up to 16 whole image rows
  const bool fast_w = (W == 128 || W == 256) && H % 16 == 0, fast_h = (H == 128 || H == 256);
  cudaError_t e;
  if (fast_w) {                                           // pre1 / pre2 convs as the prologue of the row transform
    const float* pre_w[4] = {w.pre1_w, w.pre1_b, w.pre2_w, w.pre2_b};
    e = lg::launch_fft_rows_fwd_pre(W, C, msf, panf, pre_w, S, N, H, s);
  } else {
    fre_pre_kernel<C><<<(unsigned)((NP + 255) / 256), 256, 0, s>>>(msf, panf, w.pre1_w, w.pre1_b, w.pre2_w, w.pre2_b, pre, HW, NP);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    e = lgctx::launch_fft_rows(s, 0, pre, 2 * C, S, nullptr, 0, nullptr, 0, N, H, W, 2 * C, 1.f, 0);
  }
  if (e != cudaSuccess) return e;
  // column stage: one fused kernel (forward FFT, amp / pha fusion, inverse FFT) where it is built, else three passes
  static const bool unfused = [] { const char* v = getenv("LGTEUN_FRE_UNFUSED"); return v && v[0] == '1'; }();   // A/B aid
  const float* fuse_w[8] = {w.a0w, w.a0b, w.a2w, w.a2b, w.p0w, w.p0b, w.p2w, w.p2b};
  e = (fast_h && !unfused) ? lg::launch_fre_cols_fused(H, C, S, G, fuse_w, N, W, s) : cudaErrorNotSupported;
  if (e == cudaErrorNotSupported) {
    e = fast_h ? lg::launch_fft_cols_plain(H, 2 * C, S, N, W, -1, s) : lgctx::launch_fft_cols(s, S, N, H, W, 2 * C, -1, 1);
    if (e != cudaSuccess) return e;
    fre_fuse_kernel<C><<<(unsigned)((nbins + 127) / 128), 128, 0, s>>>(reinterpret_cast<const float2*>(S), reinterpret_cast<float2*>(G),
                                                                      nbins, w.a0w, w.a0b, w.a2w, w.a2b, w.p0w, w.p0b, w.p2w, w.p2b);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    e = fast_h ? lg::launch_fft_cols_plain(H, C, G, N, W, +1, s) : lgctx::launch_fft_cols(s, G, N, H, W, C, +1, 0);
  }
  if (e != cudaSuccess) return e;
  if (fast_w) return lg::launch_fft_rows_inv_post(W, C, G, w.post_w, w.post_b, out, N, H, s);   // C2R + |.| + post conv + NCHW
  if ((e = lgctx::launch_fft_rows(s, 1, nullptr, 0, G, pre, C, pre, C, N, H, W, C, 1.f / ((float)H * (float)W), 1)) != cudaSuccess)
    return e;
  fre_post_kernel<C><<<(unsigned)((NP + 255) / 256), 256, 0, s>>>(pre, w.post_w, w.post_b, out, HW, NP);
  return cudaGetLastError();
}

// the register-resident passes read their twiddles from a per-device table that lgteun_create() fills; this operator has no handle
cudaError_t ensure_tables(int device, cudaStream_t s) {
  static std::mutex mu;
  static bool done[64] = {};
  std::lock_guard<std::mutex> lock(mu);
  if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
  if (done[device]) return cudaSuccess;
  cudaError_t e = lg::fft256_init_tables(s);
  if (e == cudaSuccess) done[device] = true;
  return e;
}

}  // namespace lgcomp

using namespace lgctx;

extern "C" {

int64_t lgteun_op_freprocess_workspace_bytes(int N, int C, int H, int W) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  const size_t NP = (size_t)N * H * W, nbins = (size_t)N * H * (W / 2 + 1);
  return (int64_t)((lgcomp::align64(NP * 2 * C) + lgcomp::align64(nbins * 2 * C * 2) + lgcomp::align64(nbins * C * 2)) * sizeof(float));
}

int lgteun_op_freprocess(int device, const float* msf, const float* panf, float* out, int N, int C, int H, int W,
                         const float* const* weights, float* workspace, int64_t workspace_bytes, void* stream) {
  if (!msf || !panf || !out || !weights || !workspace) return fail(LGTEUN_EINVAL, "NULL argument");
  for (int i = 0; i < 14; ++i)
    if (!weights[i]) return fail(LGTEUN_ESTATE, "Freprocess: weight " + std::to_string(i) + " is NULL");
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  if (N <= 0 || !pow2(H) || !pow2(W) || H < 8 || W < 8 || H > 1024 || W > 1024)
    return fail(LGTEUN_EINVAL, "Freprocess: H and W must be powers of two in [8, 1024] (power-of-two FFT passes)");
  if (!(C == 4 || C == 8 || C == 16)) return fail(LGTEUN_EINVAL, "Freprocess: channels must be 4, 8 or 16 (SFIIN uses 8)");
  if (workspace_bytes < lgteun_op_freprocess_workspace_bytes(N, C, H, W)) return fail(LGTEUN_ENOMEM, "Freprocess: workspace too small");
  CK(cudaSetDevice(device));
  const lgcomp::FreW w{weights[0], weights[1], weights[2],  weights[3],  weights[4],  weights[5],  weights[6],
                       weights[7], weights[8], weights[9], weights[10], weights[11], weights[12], weights[13]};
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = lgcomp::ensure_tables(device, s);
  if (e != cudaSuccess) return fail_cuda(e, "Freprocess: twiddle tables");
  if (C == 4) e = lgcomp::run_freprocess<4>(msf, panf, out, N, H, W, w, workspace, s);
  else if (C == 8) e = lgcomp::run_freprocess<8>(msf, panf, out, N, H, W, w, workspace, s);
  else e = lgcomp::run_freprocess<16>(msf, panf, out, N, H, W, w, workspace, s);
  if (e != cudaSuccess) return fail_cuda(e, "Freprocess");
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------------
//   PanFormer's WindowAttention.forward(x, y=None)        models/common/modules.py:341-422
//     (cyclic shift by -ws/2 :388-391) -> to_qkv | to_q(y), to_kv(x) (bias-free Linear :397-402) -> windows of ws x ws
//     tokens, heads x head_dim channels (:407-410) -> dots = q k^T * scale (:416) + pos_embedding (relative table gathered by
//     get_relative_distances :336-339,418-421) + the two -inf masks on the last window row / column when shifted (:423-425,
//     create_mask :318-333) -> softmax -> attn v (:427-429) -> to_out (Linear with bias :433) -> cyclic shift back (:436-437).
// One CTA walks windows; the projection weights stay in shared memory (k-major, conflict-free columns), a window's tokens,
// q/k/v and the attention output never leave the SM.  x, y, out: NHWC [b, n_h, n_w, dim] as the reference passes them.
namespace lgcomp {

constexpr int WS = 4, T = WS * WS;      // PanFormer: win_size = 4 (models/panformer.py:22)

struct WinAttnArgs {
  const float *x, *y, *wq, *wkv, *wout, *bout, *pos, *ul_mask, *lr_mask;
  float* out;
  int b, n_h, n_w, dim, heads, shifted, relative;
  float scale;
};

template <int HD>
__global__ void __launch_bounds__(128) win_attn_kernel(WinAttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int dim = a.dim, inner = a.heads * HD, n3 = 3 * inner;
  float* wt = sm;                         // [dim][n3]   W^T of (q | k | v)
  float* wo = wt + dim * n3;              // [inner][dim] to_out.weight^T
  float* bo = wo + inner * dim;           // [dim]
  float* pos = bo + dim;                  // [T][T]
  float* ul = pos + T * T;                // [T][T] masks (zeros when not shifted)
  float* lr = ul + T * T;
  float* xs = lr + T * T;                 // [T][dim]  kv source
  float* ys = xs + T * dim;               // [T][dim]  q source (== xs for self attention)
  float* qkv = ys + T * dim;              // [T][n3]
  float* ao = qkv + T * n3;               // [T][inner]
  const int tid = threadIdx.x;
  for (int i = tid; i < dim * n3; i += 128) {
    const int k = i / n3, o = i - k * n3;
    wt[i] = o < inner ? __ldg(a.wq + (size_t)o * dim + k) : __ldg(a.wkv + (size_t)(o - inner) * dim + k);
  }
  for (int i = tid; i < inner * dim; i += 128) {
    const int k = i / dim, c = i - k * dim;
    wo[i] = __ldg(a.wout + (size_t)c * inner + k);
  }
  for (int i = tid; i < dim; i += 128) bo[i] = __ldg(a.bout + i);
  for (int i = tid; i < T * T; i += 128) {
    const int qi = i / T, kj = i % T;
    pos[i] = a.relative ? __ldg(a.pos + ((kj / WS - qi / WS) + WS - 1) * (2 * WS - 1) + (kj % WS - qi % WS) + WS - 1)
                        : __ldg(a.pos + i);
    ul[i] = a.shifted ? __ldg(a.ul_mask + i) : 0.f;
    lr[i] = a.shifted ? __ldg(a.lr_mask + i) : 0.f;
  }
  const int nw_h = a.n_h / WS, nw_w = a.n_w / WS, nwin = a.b * nw_h * nw_w, d = a.shifted ? WS / 2 : 0;
  const bool cross = a.y != nullptr;
  const int d4 = dim / 4;
  for (int g = blockIdx.x; g < nwin; g += gridDim.x) {
    const int n = g / (nw_h * nw_w), wy = (g / nw_w) % nw_h, wx = g % nw_w;
    __syncthreads();                      // weights staged / previous window done with the tiles
    for (int i = tid; i < T * d4; i += 128) {
      const int t = i / d4, c = i - t * d4;
      const int py = (wy * WS + t / WS + d) % a.n_h, px = (wx * WS + t % WS + d) % a.n_w;
      const size_t src = (((size_t)n * a.n_h + py) * a.n_w + px) * dim;
      reinterpret_cast<float4*>(xs + t * dim)[c] = __ldg(reinterpret_cast<const float4*>(a.x + src) + c);
      if (cross) reinterpret_cast<float4*>(ys + t * dim)[c] = __ldg(reinterpret_cast<const float4*>(a.y + src) + c);
    }
    __syncthreads();
    // projections: thread = output column, all 16 tokens
    for (int o = tid; o < n3; o += 128) {
      const float* in = (cross && o < inner) ? ys : xs;
      float acc[T];
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] = 0.f;
      for (int k = 0; k < dim; k += 4) {
        const float w0 = wt[k * n3 + o], w1 = wt[(k + 1) * n3 + o], w2 = wt[(k + 2) * n3 + o], w3 = wt[(k + 3) * n3 + o];
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const float4 xv = *reinterpret_cast<const float4*>(in + t * dim + k);
          acc[t] = fmaf(xv.x, w0, acc[t]);
          acc[t] = fmaf(xv.y, w1, acc[t]);
          acc[t] = fmaf(xv.z, w2, acc[t]);
          acc[t] = fmaf(xv.w, w3, acc[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < T; ++t) qkv[t * n3 + o] = acc[t];
    }
    __syncthreads();
    // attention: thread = (head, query)
    for (int r = tid; r < a.heads * T; r += 128) {
      const int h = r / T, qi = r % T;
      float q[HD], s[T];
#pragma unroll
      for (int e = 0; e < HD; ++e) q[e] = qkv[qi * n3 + h * HD + e];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < T; ++j) {
        const float* kr = qkv + j * n3 + inner + h * HD;
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < HD; ++e) dot = fmaf(q[e], kr[e], dot);
        float v = dot * a.scale + pos[qi * T + j];
        if (wy == nw_h - 1) v += ul[qi * T + j];
        if (wx == nw_w - 1) v += lr[qi * T + j];
        s[j] = v;
        mx = fmaxf(mx, v);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < T; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
      const float inv = 1.f / sum;
      float o[HD];
#pragma unroll
      for (int e = 0; e < HD; ++e) o[e] = 0.f;
#pragma unroll
      for (int j = 0; j < T; ++j) {
        const float* vr = qkv + j * n3 + 2 * inner + h * HD;
        const float p = s[j] * inv;
#pragma unroll
        for (int e = 0; e < HD; ++e) o[e] = fmaf(p, vr[e], o[e]);
      }
#pragma unroll
      for (int e = 0; e < HD; ++e) ao[qi * inner + h * HD + e] = o[e];
    }
    __syncthreads();
    // to_out + store at the un-shifted pixel: thread = (output channel, group of 4 tokens)
    for (int i = tid; i < dim * 4; i += 128) {
      const int c = i % dim, tg = i / dim;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < inner; k += 4) {
        const float w0 = wo[k * dim + c], w1 = wo[(k + 1) * dim + c], w2 = wo[(k + 2) * dim + c], w3 = wo[(k + 3) * dim + c];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float4 v = *reinterpret_cast<const float4*>(ao + (tg * 4 + t) * inner + k);
          acc[t] = fmaf(v.x, w0, acc[t]);
          acc[t] = fmaf(v.y, w1, acc[t]);
          acc[t] = fmaf(v.z, w2, acc[t]);
          acc[t] = fmaf(v.w, w3, acc[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int tok = tg * 4 + t;
        const int py = (wy * WS + tok / WS + d) % a.n_h, px = (wx * WS + tok % WS + d) % a.n_w;
        a.out[(((size_t)n * a.n_h + py) * a.n_w + px) * dim + c] = acc[t] + bo[c];
      }
    }
  }
}

// PanFormer's own configuration (dim 64, 4 heads of 16: models/panformer.py:22) as a register-tiled kernel.  The projections are
// 4 tokens x 6 (q/k/v) resp. 4 tokens x 2 (to_out) outputs per thread: a thread's columns are o = lane + 32 i, stored side by
// side in the k-major weight copy, so the weights arrive as conflict-free 64-bit loads and the tokens as warp-uniform 128-bit
// loads (12 + 4..8 shared-memory loads per 96 FMAs instead of 20 per 64).  The attention runs on all threads: two threads per
// (head, query) row take eight keys each and merge max / sum / output with one shuffle exchange.  A CTA of 256 threads works on
// two windows at a time that share the 64 KB of weights (16 warps per SM at two CTAs).
__device__ __forceinline__ float f4c(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

__global__ void __launch_bounds__(256, 2) win_attn_pf_kernel(WinAttnArgs a) {
  constexpr int DIM = 64, NH = 4, HD = 16, INNER = NH * HD, N3 = 3 * INNER, QS = N3 + 4, CPT = N3 / 32, DPT = DIM / 32;
  constexpr int PS = T + 1, AS = INNER + 4;          // padded row strides of the bias / mask tables and of the attention output
  constexpr int SLOT = 2 * T * DIM + T * QS;        // per-window tiles: x, y (later the attention output), q/k/v
  extern __shared__ __align__(16) float sm[];
  float* wt = sm;                         // [DIM][N3]     (q | k | v)^T, columns permuted: slot lane * CPT + i holds o = lane + 32 i
  float* wo = wt + DIM * N3;              // [INNER][DIM]  to_out.weight^T, slot lane * DPT + i holds c = lane + 32 i
  float* bo = wo + INNER * DIM;           // [DIM]
  float* pos = bo + DIM;                  // [T][PS]
  float* ul = pos + T * PS;
  float* lr = ul + T * PS;
  // A CTA works on two windows at a time (threads 0-127 / 128-255) that share the weights: 16 warps per SM instead of 8
  const int tid = threadIdx.x, slot = tid >> 7, lane = tid & 31, tg = (tid >> 5) & 3, t7 = tid & 127;
  float* xs = lr + T * PS + slot * SLOT;  // [T][DIM]
  float* ys = xs + T * DIM;
  float* qkv = ys + T * DIM;              // [T][QS]  (row stride padded: the per-query loads of the attention spread over banks)
  float* ao = xs;                         // [T][AS]  the attention output reuses the token tiles (dead after the projections)
  for (int i = tid; i < N3 * DIM; i += 256) {
    const int o = i / DIM, k = i - o * DIM;
    wt[k * N3 + (o & 31) * CPT + (o >> 5)] = o < INNER ? __ldg(a.wq + i) : __ldg(a.wkv + i - INNER * DIM);
  }
  for (int i = tid; i < DIM * INNER; i += 256) {
    const int c = i / INNER, k = i - c * INNER;
    wo[k * DIM + (c & 31) * DPT + (c >> 5)] = __ldg(a.wout + i);
  }
  for (int i = tid; i < DIM; i += 256) bo[i] = __ldg(a.bout + i);
  for (int i = tid; i < T * T; i += 256) {
    const int qi = i / T, kj = i % T, ps = qi * PS + kj;
    pos[ps] = a.relative ? __ldg(a.pos + ((kj / WS - qi / WS) + WS - 1) * (2 * WS - 1) + (kj % WS - qi % WS) + WS - 1)
                        : __ldg(a.pos + i);
    ul[ps] = a.shifted ? __ldg(a.ul_mask + i) : 0.f;
    lr[ps] = a.shifted ? __ldg(a.lr_mask + i) : 0.f;
  }
  const int nw_h = a.n_h / WS, nw_w = a.n_w / WS, nwin = a.b * nw_h * nw_w, d = a.shifted ? WS / 2 : 0;
  const bool cross = a.y != nullptr;
  for (int g0 = blockIdx.x * 2; g0 < nwin; g0 += gridDim.x * 2) {
    const int g = g0 + slot;
    const bool active = g < nwin;         // uniform over the four warps of a slot; every thread still meets the barriers
    const int n = g / (nw_h * nw_w), wy = (g / nw_w) % nw_h, wx = g % nw_w;
    __syncthreads();                      // weights staged / previous windows done with the tiles
    for (int i = t7; active && i < T * (DIM / 4); i += 128) {
      const int t = i / (DIM / 4), c = i - t * (DIM / 4);
      const int py = (wy * WS + t / WS + d) % a.n_h, px = (wx * WS + t % WS + d) % a.n_w;
      const size_t src = (((size_t)n * a.n_h + py) * a.n_w + px) * DIM;
      reinterpret_cast<float4*>(xs + t * DIM)[c] = __ldg(reinterpret_cast<const float4*>(a.x + src) + c);
      if (cross) reinterpret_cast<float4*>(ys + t * DIM)[c] = __ldg(reinterpret_cast<const float4*>(a.y + src) + c);
    }
    __syncthreads();
    if (active) {   // q / k / v projections: tokens 4 tg .. 4 tg + 3, columns lane + 32 i (i < 2: q, from y in a cross block)
      const float* qsrc = cross ? ys : xs;
      float acc[4][CPT];
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int i = 0; i < CPT; ++i) acc[t][i] = 0.f;
#pragma unroll 2
      for (int k = 0; k < DIM; k += 4) {
        float4 xv[4], qv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          xv[t] = *reinterpret_cast<const float4*>(xs + (tg * 4 + t) * DIM + k);
          qv[t] = *reinterpret_cast<const float4*>(qsrc + (tg * 4 + t) * DIM + k);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float* wr = wt + (k + kk) * N3 + lane * CPT;
          const float2 w01 = *reinterpret_cast<const float2*>(wr), w23 = *reinterpret_cast<const float2*>(wr + 2),
                       w45 = *reinterpret_cast<const float2*>(wr + 4);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float xin = f4c(xv[t], kk), qin = f4c(qv[t], kk);
            acc[t][0] = fmaf(qin, w01.x, acc[t][0]);
            acc[t][1] = fmaf(qin, w01.y, acc[t][1]);
            acc[t][2] = fmaf(xin, w23.x, acc[t][2]);
            acc[t][3] = fmaf(xin, w23.y, acc[t][3]);
            acc[t][4] = fmaf(xin, w45.x, acc[t][4]);
            acc[t][5] = fmaf(xin, w45.y, acc[t][5]);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int i = 0; i < CPT; ++i) qkv[(tg * 4 + t) * QS + lane + 32 * i] = acc[t][i];
    }
    __syncthreads();
    if (active) {   // attention: row = (head, query) = t7 / 2, this thread's eight keys = half * 8 ..
      const int r = t7 >> 1, half = t7 & 1, h = r >> 4, qi = r & 15;
      float q[HD], sc[8];
#pragma unroll
      for (int e = 0; e < HD; e += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(qkv + qi * QS + h * HD + e);
        q[e] = t4.x; q[e + 1] = t4.y; q[e + 2] = t4.z; q[e + 3] = t4.w;
      }
      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = half * 8 + jj;
        const float* kr = qkv + j * QS + INNER + h * HD;
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < HD; e += 4) {
          const float4 k4 = *reinterpret_cast<const float4*>(kr + e);
          dot = fmaf(q[e], k4.x, dot);
          dot = fmaf(q[e + 1], k4.y, dot);
          dot = fmaf(q[e + 2], k4.z, dot);
          dot = fmaf(q[e + 3], k4.w, dot);
        }
        float v = dot * a.scale + pos[qi * PS + j];
        if (wy == nw_h - 1) v += ul[qi * PS + j];
        if (wx == nw_w - 1) v += lr[qi * PS + j];
        sc[jj] = v;
        mx = fmaxf(mx, v);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));      // the row's diagonal key is never masked: mx is finite
      float sum = 0.f, o[HD];
#pragma unroll
      for (int e = 0; e < HD; ++e) o[e] = 0.f;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float p = expf(sc[jj] - mx);
        sum += p;
        const float* vr = qkv + (half * 8 + jj) * QS + 2 * INNER + h * HD;
#pragma unroll
        for (int e = 0; e < HD; e += 4) {
          const float4 v4 = *reinterpret_cast<const float4*>(vr + e);
          o[e] = fmaf(p, v4.x, o[e]);
          o[e + 1] = fmaf(p, v4.y, o[e + 1]);
          o[e + 2] = fmaf(p, v4.z, o[e + 2]);
          o[e + 3] = fmaf(p, v4.w, o[e + 3]);
        }
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
#pragma unroll
      for (int e = 0; e < HD; ++e) o[e] += __shfl_xor_sync(0xffffffffu, o[e], 1);
      const float inv = 1.f / sum;
#pragma unroll
      for (int e = 0; e < HD; ++e) o[e] *= inv;
      // ao aliases the token tiles: every thread is past the projections (barrier above), and q/k/v, which the other rows
      // still read, live in their own region
      float* dst = ao + qi * AS + h * HD + half * (HD / 2);
#pragma unroll
      for (int e = 0; e < HD / 2; ++e) dst[e] = half ? o[HD / 2 + e] : o[e];
    }
    __syncthreads();
    if (active) {   // to_out + store at the un-shifted pixel: tokens 4 tg .. 4 tg + 3, channels lane + 32 i
      float acc[4][DPT];
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int i = 0; i < DPT; ++i) acc[t][i] = 0.f;
#pragma unroll 4
      for (int k = 0; k < INNER; k += 4) {
        float4 av[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) av[t] = *reinterpret_cast<const float4*>(ao + (tg * 4 + t) * AS + k);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float2 w = *reinterpret_cast<const float2*>(wo + (k + kk) * DIM + lane * DPT);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float ain = f4c(av[t], kk);
            acc[t][0] = fmaf(ain, w.x, acc[t][0]);
            acc[t][1] = fmaf(ain, w.y, acc[t][1]);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int tok = tg * 4 + t;
        const int py = (wy * WS + tok / WS + d) % a.n_h, px = (wx * WS + tok % WS + d) % a.n_w;
        float* dst = a.out + (((size_t)n * a.n_h + py) * a.n_w + px) * DIM;
        dst[lane] = acc[t][0] + bo[lane];
        dst[lane + 32] = acc[t][1] + bo[lane + 32];
      }
    }
  }
}

inline size_t win_attn_smem(int dim, int inner) {
  return (size_t)(dim * 3 * inner + inner * dim + dim + 3 * T * T + 2 * T * dim + T * 3 * inner + T * inner) * sizeof(float);
}

cudaError_t launch_win_attn_pf(const WinAttnArgs& a, cudaStream_t s) {
  const size_t smem = (size_t)(64 * 192 + 64 * 64 + 64 + 3 * T * (T + 1) + 2 * (2 * T * 64 + T * 196)) * sizeof(float);   // 110.5 KB
  cudaError_t e = cudaFuncSetAttribute(win_attn_pf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int nwin = a.b * (a.n_h / WS) * (a.n_w / WS);
  win_attn_pf_kernel<<<(unsigned)std::min((nwin + 1) / 2, 148 * 2), 256, smem, s>>>(a);
  return cudaGetLastError();
}

template <int HD>
cudaError_t launch_win_attn(const WinAttnArgs& a, cudaStream_t s) {
  const size_t smem = win_attn_smem(a.dim, a.heads * HD);
  cudaError_t e = cudaFuncSetAttribute(win_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int nwin = a.b * (a.n_h / WS) * (a.n_w / WS);
  const int per_sm = smem > 110 * 1024 ? 1 : 2;
  win_attn_kernel<HD><<<(unsigned)std::min(nwin, 148 * per_sm), 128, smem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace lgcomp

extern "C" {

int lgteun_op_window_attention(int device, const float* x, const float* y, float* out, int b, int n_h, int n_w, int dim, int heads,
                               int head_dim, int window_size, int shifted, int relative_pos_embedding, float scale, const float* w_q,
                               const float* w_kv, const float* w_out, const float* b_out, const float* pos_embedding,
                               const float* upper_lower_mask, const float* left_right_mask, void* stream) {
  if (!x || !out || !w_q || !w_kv || !w_out || !b_out || !pos_embedding) return fail(LGTEUN_EINVAL, "NULL argument");
  if (shifted && (!upper_lower_mask || !left_right_mask)) return fail(LGTEUN_ESTATE, "WindowAttention: shifted block without masks");
  if (window_size != lgcomp::WS) return fail(LGTEUN_EINVAL, "WindowAttention: window_size must be 4 (PanFormer's win_size)");
  if (!(head_dim == 8 || head_dim == 16 || head_dim == 32)) return fail(LGTEUN_EINVAL, "WindowAttention: head_dim must be 8, 16 or 32");
  if (b <= 0 || n_h <= 0 || n_w <= 0 || n_h % window_size || n_w % window_size)
    return fail(LGTEUN_EINVAL, "WindowAttention: n_h and n_w must be positive multiples of the window size");
  if (dim <= 0 || dim % 4 || heads <= 0 || heads * head_dim > 128 || dim > 128 ||
      lgcomp::win_attn_smem(dim, heads * head_dim) > 227 * 1024)
    return fail(LGTEUN_EINVAL, "WindowAttention: dim must be a multiple of 4, dim and heads*head_dim at most 128, and the "
                               "projection weights (4 * dim * heads*head_dim floats) must fit 227 KB of shared memory");
  CK(cudaSetDevice(device));
  lgcomp::WinAttnArgs a{x, y, w_q, w_kv, w_out, b_out, pos_embedding, upper_lower_mask, left_right_mask, out, b, n_h, n_w, dim,
                        heads, shifted != 0, relative_pos_embedding != 0, scale};
  cudaStream_t s = (cudaStream_t)stream;
  static const bool generic_only = [] { const char* v = getenv("LGTEUN_WINATTN_GENERIC"); return v && v[0] == '1'; }();   // A/B aid
  cudaError_t e = (dim == 64 && heads == 4 && head_dim == 16 && !generic_only) ? lgcomp::launch_win_attn_pf(a, s)
                  : head_dim == 8 ? lgcomp::launch_win_attn<8>(a, s)
                  : head_dim == 16 ? lgcomp::launch_win_attn<16>(a, s) : lgcomp::launch_win_attn<32>(a, s);
  if (e != cudaSuccess) return fail_cuda(e, "WindowAttention");
  return 0;
}

}  // extern "C"
