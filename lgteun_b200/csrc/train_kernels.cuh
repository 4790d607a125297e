// train_kernels.cuh — fp32 CUDA-core kernels of the training step (forward with saved activations + backward).
//
// UnlgFormer.train_iter (models/unlg_former.py:87-113) = Pansharpening.forward in train() mode,
// L1 loss (models/base/losses.py:19-40), loss.backward(), Adam (models/base/base_model.py:116-131).  Round-1 scope of this
// file: correct, generic (any channel count / layout), one kernel per primitive and its adjoint; the fused tcgen05
// inference kernels are not reused here because the backward needs the intermediates they never write.
//
// Tensors are addressed through TV views: NHWC [pixels, ld] with a channel offset folded into the pointer (concats and
// channel splits are free), or NCHW planes (the data module and the module boundary).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace lgtrain {

struct TV {
  float* p;
  int ld;     // NHWC: floats per pixel
  int nchw;   // 1: p[((n*C + k)*P + pix)]
  int C, lp;  // NCHW only: channels, log2(pixels per image) (image sizes are powers of two)
};
__host__ __device__ __forceinline__ size_t tv_at(const TV& t, size_t gp, int k) {
  return t.nchw ? ((((gp >> t.lp) * t.C + k) << t.lp) + (gp & (((size_t)1 << t.lp) - 1))) : gp * (size_t)t.ld + k;
}
inline TV nhwc(const float* p, int ld) { return TV{const_cast<float*>(p), ld, 0, 0, 0}; }
inline TV nchw(const float* p, int C, int P) {
  int lp = 0;
  while ((1 << lp) < P) ++lp;
  return TV{const_cast<float*>(p), 0, 1, C, lp};
}

// exact-form (erf) GELU, 4.8e-7 absolute (common.cuh: one ex2 instead of erff; the inference kernels use the same form)
__device__ __forceinline__ float gelu_exact(float x) { return lg::gelu_fast(x); }
__device__ __forceinline__ float gelu_grad(float x) { return lg::gelu_grad_fast(x); }

// ---- pointwise (1x1) convolution:  y[p,co] = sum_ci W[co*wso + ci*wsi] * f(x[p,ci]) + b[co]  (+ add[p,co]) (* gelu'(gate)) -------
// Forward of bmu.point_conv (basic_module_unformer_v2.py:13-14) with (wso, wsi) = (Cin, 1); its data gradient with the
// transposed strides (1, Cout_fwd).  ACT = 1 applies the exact GELU to x while staging it (the conv-FFN's activations,
// LGT.py:97,99, are never stored).  Shared-memory tiled fp32 GEMM: a block owns TPX pixels x TCO output channels
// (TPX * TCO = 4096), a thread 4 x 4 of them; K (input channels) is streamed in chunks of 32.
template <int ACT, int TCO>
__global__ void __launch_bounds__(256) k_pw(TV x, int Cin, const float* __restrict__ W, int wso, int wsi,
                                            const float* __restrict__ b, TV y, int Cout, size_t NP, TV add, int use_add,
                                            TV gate, int use_gate) {
  constexpr int TPX = 4096 / TCO, KC = 32, NTX = TCO / 4;
  __shared__ float xs[TPX][KC + 1];
  __shared__ float ws[KC][TCO + 1];
  const int tx = threadIdx.x % NTX, ty = threadIdx.x / NTX;
  const size_t p0 = (size_t)blockIdx.x * TPX;
  const int co0 = blockIdx.y * TCO;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < Cin; k0 += KC) {
    for (int i = threadIdx.x; i < TPX * KC; i += 256) {
      int px, k;
      if (x.nchw) { px = i % TPX; k = i / TPX; } else { k = i % KC; px = i / KC; }
      const size_t gp = p0 + px;
      float v = 0.f;
      if (gp < NP && k0 + k < Cin) { v = x.p[tv_at(x, gp, k0 + k)]; if (ACT) v = gelu_exact(v); }
      xs[px][k] = v;
    }
    for (int i = threadIdx.x; i < TCO * KC; i += 256) {
      int co, k;
      if (wsi == 1) { k = i % KC; co = i / KC; } else { co = i % TCO; k = i / TCO; }
      ws[k][co] = (co0 + co < Cout && k0 + k < Cin) ? __ldg(W + (size_t)(co0 + co) * wso + (size_t)(k0 + k) * wsi) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < KC; ++k) {
      float xv[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = xs[ty * 4 + i][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const size_t gp = p0 + ty * 4 + i;
    if (gp >= NP) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= Cout) continue;
      float v = acc[i][j] + (b ? b[co] : 0.f);
      if (use_gate) v *= gelu_grad(gate.p[tv_at(gate, gp, co)]);
      if (use_add) v += add.p[tv_at(add, gp, co)];
      y.p[tv_at(y, gp, co)] = v;
    }
  }
}

// weight / bias gradient of the same conv:  dW[co*wso + ci*wsi] += sum_p dy[p,co] * f(x[p,ci]);  db[co] += sum_p dy[p,co].
// Split-K GEMM (K = pixels): blockIdx.y selects a 64 x 64 (co, ci) tile, blockIdx.x a pixel range; a thread accumulates
// 4 x 4 entries in registers over its block's range and issues one atomicAdd per entry.
template <int ACT>
__global__ void __launch_bounds__(256) k_pw_wgrad(TV x, int Cin, TV dy, int Cout, float* __restrict__ dW, int wso, int wsi,
                                                  float* __restrict__ db, size_t NP, int ci_tiles) {
  constexpr int TP = 32;
  __shared__ __align__(16) float xs[TP][64];
  __shared__ __align__(16) float dys[TP][64];
  const int ci0 = (blockIdx.y % ci_tiles) * 64, co0 = (blockIdx.y / ci_tiles) * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bacc = 0.f;
  const bool do_bias = db && ci0 == 0 && threadIdx.x < 64 && co0 + threadIdx.x < Cout;
  const size_t tiles = (NP + TP - 1) / TP;
  for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const size_t p0 = t * TP;
    for (int i = threadIdx.x; i < TP * 64; i += 256) {
      int px, c;
      if (x.nchw) { px = i % TP; c = i / TP; } else { c = i & 63; px = i >> 6; }
      const size_t gp = p0 + px;
      float v = 0.f;
      if (gp < NP && ci0 + c < Cin) { v = x.p[tv_at(x, gp, ci0 + c)]; if (ACT) v = gelu_exact(v); }
      xs[px][c] = v;
    }
    for (int i = threadIdx.x; i < TP * 64; i += 256) {
      int px, c;
      if (dy.nchw) { px = i % TP; c = i / TP; } else { c = i & 63; px = i >> 6; }
      const size_t gp = p0 + px;
      dys[px][c] = (gp < NP && co0 + c < Cout) ? dy.p[tv_at(dy, gp, co0 + c)] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int p = 0; p < TP; ++p) {
      const float4 xv = *reinterpret_cast<const float4*>(&xs[p][tx * 4]);
      const float4 dv = *reinterpret_cast<const float4*>(&dys[p][ty * 4]);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, da[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(da[i], xa[j], acc[i][j]);
    }
    if (do_bias) {
      float sacc = 0.f;
#pragma unroll 8
      for (int p = 0; p < TP; ++p) sacc += dys[p][threadIdx.x];
      bacc += sacc;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < Cin) atomicAdd(dW + (size_t)co * wso + (size_t)ci * wsi, acc[i][j]);
    }
  }
  if (do_bias) atomicAdd(db + co0 + threadIdx.x, bacc);
}

// ---- LayerNorm over channels (LGT.py:54-61; biased variance, eps 1e-5), one thread per pixel -------------------------------------
__global__ void __launch_bounds__(256) k_ln_fwd(TV x, int C, const float* __restrict__ g, const float* __restrict__ b, TV y,
                                                size_t NP) {
  const size_t gp = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (gp >= NP) return;
  float mean = 0.f;
  for (int k = 0; k < C; ++k) mean += x.p[tv_at(x, gp, k)];
  mean /= C;
  float var = 0.f;
  for (int k = 0; k < C; ++k) { const float d = x.p[tv_at(x, gp, k)] - mean; var = fmaf(d, d, var); }
  const float rstd = 1.f / sqrtf(var / C + 1e-5f);
  for (int k = 0; k < C; ++k) y.p[tv_at(y, gp, k)] = (x.p[tv_at(x, gp, k)] - mean) * rstd * g[k] + b[k];
}

// dx (+)= LN'(x)^T dy;  dgamma += sum dy * xhat;  dbeta += sum dy   (statistics recomputed from x)
__global__ void __launch_bounds__(256) k_ln_bwd(TV x, int C, const float* __restrict__ g, TV dy, TV dx, int accumulate,
                                                float* __restrict__ dgamma, float* __restrict__ dbeta, size_t NP) {
  extern __shared__ float sm[];   // [2*C]
  for (int i = threadIdx.x; i < 2 * C; i += 256) sm[i] = 0.f;
  __syncthreads();
  const size_t gp = (size_t)blockIdx.x * 256 + threadIdx.x;
  const bool live = gp < NP;
  float mean = 0.f, rstd = 0.f, m1 = 0.f, m2 = 0.f;
  if (live) {
    for (int k = 0; k < C; ++k) mean += x.p[tv_at(x, gp, k)];
    mean /= C;
    float var = 0.f;
    for (int k = 0; k < C; ++k) { const float d = x.p[tv_at(x, gp, k)] - mean; var = fmaf(d, d, var); }
    rstd = 1.f / sqrtf(var / C + 1e-5f);
    for (int k = 0; k < C; ++k) {
      const float xh = (x.p[tv_at(x, gp, k)] - mean) * rstd, gg = dy.p[tv_at(dy, gp, k)] * g[k];
      m1 += gg;
      m2 = fmaf(gg, xh, m2);
    }
    m1 /= C;
    m2 /= C;
  }
  for (int k = 0; k < C; ++k) {
    float xh = 0.f, d = 0.f;
    if (live) {
      xh = (x.p[tv_at(x, gp, k)] - mean) * rstd;
      d = dy.p[tv_at(dy, gp, k)];
      float v = rstd * (d * g[k] - m1 - xh * m2);
      const size_t o = tv_at(dx, gp, k);
      if (accumulate) v += dx.p[o];
      dx.p[o] = v;
    }
    float a = d * xh, bsum = d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(sm + k, a); atomicAdd(sm + C + k, bsum); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) { atomicAdd(dgamma + i, sm[i]); atomicAdd(dbeta + i, sm[C + i]); }
}

// NHWC-contiguous LayerNorm with lanes = channels (C in {16, 32, 64}): coalesced loads, statistics by warp shuffles,
// gamma / beta gradients accumulate per lane in registers over all the pixels a warp visits.  NP % (32 / min(C, 32)) == 0.
template <int C, int BWD>
__global__ void __launch_bounds__(256) k_ln_warp(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                                 const float* __restrict__ dy, float* __restrict__ out, int accumulate,
                                                 float* __restrict__ dgamma, float* __restrict__ dbeta, size_t NP) {
  constexpr int GROUP = C < 32 ? C : 32, VPL = C / GROUP, PPW = 32 / GROUP;
  __shared__ float red[2 * C];
  if (BWD) {
    for (int i = threadIdx.x; i < 2 * C; i += 256) red[i] = 0.f;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, cl = lane % GROUP, sub = lane / GROUP;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * 8;
  float gam[VPL], bet[VPL], ag[VPL], ab[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) { gam[t] = g[cl + 32 * t]; bet[t] = BWD ? 0.f : b[cl + 32 * t]; ag[t] = 0.f; ab[t] = 0.f; }
  for (size_t gp = warp * PPW + sub; gp < NP; gp += nwarps * PPW) {
    float v[VPL], s = 0.f;
#pragma unroll
    for (int t = 0; t < VPL; ++t) { v[t] = x[gp * C + cl + 32 * t]; s += v[t]; }
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / C);
    float var = 0.f;
#pragma unroll
    for (int t = 0; t < VPL; ++t) { v[t] -= mean; var = fmaf(v[t], v[t], var); }
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = 1.f / sqrtf(var * (1.f / C) + 1e-5f);
    if (!BWD) {
#pragma unroll
      for (int t = 0; t < VPL; ++t) out[gp * C + cl + 32 * t] = v[t] * rstd * gam[t] + bet[t];
    } else {
      float d[VPL], m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        v[t] *= rstd;
        d[t] = dy[gp * C + cl + 32 * t];
        const float gg = d[t] * gam[t];
        m1 += gg;
        m2 = fmaf(gg, v[t], m2);
        ag[t] = fmaf(d[t], v[t], ag[t]);
        ab[t] += d[t];
      }
#pragma unroll
      for (int o = GROUP / 2; o > 0; o >>= 1) {
        m1 += __shfl_xor_sync(0xffffffffu, m1, o);
        m2 += __shfl_xor_sync(0xffffffffu, m2, o);
      }
      m1 *= (1.f / C);
      m2 *= (1.f / C);
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        float r = rstd * (d[t] * gam[t] - m1 - v[t] * m2);
        const size_t o = gp * C + cl + 32 * t;
        if (accumulate) r += out[o];
        out[o] = r;
      }
    }
  }
  if (BWD) {
#pragma unroll
    for (int t = 0; t < VPL; ++t) { atomicAdd(red + cl + 32 * t, ag[t]); atomicAdd(red + C + cl + 32 * t, ab[t]); }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += 256) { atomicAdd(dgamma + i, red[i]); atomicAdd(dbeta + i, red[C + i]); }
  }
}

// The same with four channels per lane (C / 4 lanes per pixel, 128 / C pixels per warp and iteration): 16-byte accesses and
// log2(C / 4) shuffle steps per reduction instead of log2(min(C, 32)) on scalars — the scalar form spent its time in SHFL
// (20 per pixel in the backward).  NP % (128 / C) == 0.
template <int C, int BWD>
__global__ void __launch_bounds__(256) k_ln_v4(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                               const float* __restrict__ dy, float* __restrict__ out, int accumulate,
                                               float* __restrict__ dgamma, float* __restrict__ dbeta, size_t NP) {
  constexpr int GROUP = C / 4, PPW = 32 / GROUP;
  __shared__ float red[2 * C];
  if (BWD) {
    for (int i = threadIdx.x; i < 2 * C; i += 256) red[i] = 0.f;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, cl = lane % GROUP, sub = lane / GROUP;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * 8;
  const float4 gam = *reinterpret_cast<const float4*>(g + 4 * cl);
  const float4 bet = BWD ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(b + 4 * cl);
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* d4 = reinterpret_cast<const float4*>(dy);
  float4* o4 = reinterpret_cast<float4*>(out);
  for (size_t gp = warp * PPW + sub; gp < NP; gp += nwarps * PPW) {
    const size_t o = gp * GROUP + cl;
    float4 v = __ldg(x4 + o), d = make_float4(0.f, 0.f, 0.f, 0.f), acc = d;
    if (BWD) {
      d = __ldg(d4 + o);
      if (accumulate) acc = o4[o];
    }
    float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
    for (int w = GROUP / 2; w > 0; w >>= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
    const float mean = s * (1.f / C);
    v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
    float var = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
#pragma unroll
    for (int w = GROUP / 2; w > 0; w >>= 1) var += __shfl_xor_sync(0xffffffffu, var, w);
    const float rstd = 1.f / sqrtf(var * (1.f / C) + 1e-5f);
    v.x *= rstd; v.y *= rstd; v.z *= rstd; v.w *= rstd;
    if (!BWD) {
      o4[o] = make_float4(fmaf(v.x, gam.x, bet.x), fmaf(v.y, gam.y, bet.y), fmaf(v.z, gam.z, bet.z), fmaf(v.w, gam.w, bet.w));
    } else {
      const float4 gg = make_float4(d.x * gam.x, d.y * gam.y, d.z * gam.z, d.w * gam.w);
      float m1 = (gg.x + gg.y) + (gg.z + gg.w);
      float m2 = fmaf(gg.x, v.x, fmaf(gg.y, v.y, fmaf(gg.z, v.z, gg.w * v.w)));
      ag.x = fmaf(d.x, v.x, ag.x); ag.y = fmaf(d.y, v.y, ag.y); ag.z = fmaf(d.z, v.z, ag.z); ag.w = fmaf(d.w, v.w, ag.w);
      ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
#pragma unroll
      for (int w = GROUP / 2; w > 0; w >>= 1) {
        m1 += __shfl_xor_sync(0xffffffffu, m1, w);
        m2 += __shfl_xor_sync(0xffffffffu, m2, w);
      }
      m1 *= (1.f / C);
      m2 *= (1.f / C);
      o4[o] = make_float4(fmaf(rstd, gg.x - m1 - v.x * m2, acc.x), fmaf(rstd, gg.y - m1 - v.y * m2, acc.y),
                          fmaf(rstd, gg.z - m1 - v.z * m2, acc.z), fmaf(rstd, gg.w - m1 - v.w * m2, acc.w));
    }
  }
  if (BWD) {
    atomicAdd(red + 4 * cl, ag.x); atomicAdd(red + 4 * cl + 1, ag.y); atomicAdd(red + 4 * cl + 2, ag.z); atomicAdd(red + 4 * cl + 3, ag.w);
    atomicAdd(red + C + 4 * cl, ab.x); atomicAdd(red + C + 4 * cl + 1, ab.y); atomicAdd(red + C + 4 * cl + 2, ab.z); atomicAdd(red + C + 4 * cl + 3, ab.w);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += 256) { atomicAdd(dgamma + i, red[i]); atomicAdd(dbeta + i, red[C + i]); }
  }
}

// ---- depthwise KxK convolution with zero padding (bmu.dep_conv, basic_module_unformer_v2.py:17-18), K in {1, 3} ---------------
// flip = 1 uses the point-reflected taps: the data gradient of the same conv.  H, W, C are powers of two (lh, lw, lc).
template <int K>
__global__ void __launch_bounds__(256) k_dw(TV x, const float* __restrict__ w, const float* __restrict__ b, TV y, int N,
                                            int lh, int lw, int lc, int flip, TV add, float add_scale, int use_add) {
  const int H = 1 << lh, W = 1 << lw, C = 1 << lc;
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x, total = (size_t)N << (lh + lw + lc);
  if (idx >= total) return;
  int k;
  size_t gp;
  if (y.nchw) { const size_t r = idx >> (lh + lw); k = (int)(r & (C - 1)); gp = ((r >> lc) << (lh + lw)) + (idx & (((size_t)1 << (lh + lw)) - 1)); }
  else { k = (int)(idx & (C - 1)); gp = idx >> lc; }
  const int xx = (int)(gp & (W - 1)), yy = (int)((gp >> lw) & (H - 1));
  float acc = b ? b[k] : 0.f;
#pragma unroll
  for (int dy = 0; dy < K; ++dy)
#pragma unroll
    for (int dx = 0; dx < K; ++dx) {
      const int sy = yy + dy - K / 2, sx = xx + dx - K / 2;
      if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
      const int wi = flip ? K * K - 1 - (dy * K + dx) : dy * K + dx;
      acc = fmaf(w[k * K * K + wi], x.p[tv_at(x, gp + (size_t)(dy - K / 2) * W + (dx - K / 2), k)], acc);
    }
  if (use_add) acc = fmaf(add_scale, add.p[tv_at(add, gp, k)], acc);
  y.p[tv_at(y, gp, k)] = acc;
}
// 3x3, NHWC contiguous in and out (ld == C, C % 4 == 0, W % SEG == 0): thread = (SEG-pixel row segment, channel quad) with its
// 9 x 4 taps in registers; a 3x3 window of float4 slides along the row, so a pixel costs three 16-byte loads and one store
// instead of nine loads, nine shared-memory tap reads and one store (the L1 data pipe bounded that form: 110 us at
// 4 x 256 x 256 x 128; this one us).
template <int SEG>
__global__ void __launch_bounds__(256) k_dw3_v4(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                float* __restrict__ y, int N, int lh, int lw, int lc, int flip) {
  constexpr int LS = SEG == 16 ? 4 : 3;
  const int H = 1 << lh, W = 1 << lw, lq = lc - 2, CQ = 1 << lq, lseg = lw - LS;
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x, total = (size_t)N << (lh + lseg + lq);
  if (idx >= total) return;
  const int q = (int)(idx & (CQ - 1));
  const size_t s = idx >> lq;
  const int x0 = (int)(s & (((size_t)1 << lseg) - 1)) << LS, yy = (int)((s >> lseg) & (H - 1));
  const size_t row = (s >> lseg) << lw;          // pixel index of (n, yy, 0)
  const bool up = yy > 0, dn = yy + 1 < H;
  float4 wt[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int ts = flip ? 8 - t : t;
    wt[t] = make_float4(__ldg(w + (4 * q) * 9 + ts), __ldg(w + (4 * q + 1) * 9 + ts), __ldg(w + (4 * q + 2) * 9 + ts), __ldg(w + (4 * q + 3) * 9 + ts));
  }
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 bias = b ? make_float4(__ldg(b + 4 * q), __ldg(b + 4 * q + 1), __ldg(b + 4 * q + 2), __ldg(b + 4 * q + 3)) : zero;
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
  float4* __restrict__ y4 = reinterpret_cast<float4*>(y);
  float4 wa[3], wb[3], wc[3];
  auto column = [&](int xx, float4 (&c)[3]) {
    const bool in = xx >= 0 && xx < W;
    const size_t p = ((row + xx) << lq) + q;
    c[0] = (in && up) ? __ldg(x4 + p - ((size_t)W << lq)) : zero;
    c[1] = in ? __ldg(x4 + p) : zero;
    c[2] = (in && dn) ? __ldg(x4 + p + ((size_t)W << lq)) : zero;
  };
  column(x0 - 1, wa);
  column(x0, wb);
#pragma unroll 4
  for (int i = 0; i < SEG; ++i) {
    column(x0 + i + 1, wc);
    float4 acc = bias;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      acc.x = fmaf(wt[3 * r].x, wa[r].x, acc.x); acc.y = fmaf(wt[3 * r].y, wa[r].y, acc.y);
      acc.z = fmaf(wt[3 * r].z, wa[r].z, acc.z); acc.w = fmaf(wt[3 * r].w, wa[r].w, acc.w);
      acc.x = fmaf(wt[3 * r + 1].x, wb[r].x, acc.x); acc.y = fmaf(wt[3 * r + 1].y, wb[r].y, acc.y);
      acc.z = fmaf(wt[3 * r + 1].z, wb[r].z, acc.z); acc.w = fmaf(wt[3 * r + 1].w, wb[r].w, acc.w);
      acc.x = fmaf(wt[3 * r + 2].x, wc[r].x, acc.x); acc.y = fmaf(wt[3 * r + 2].y, wc[r].y, acc.y);
      acc.z = fmaf(wt[3 * r + 2].z, wc[r].z, acc.z); acc.w = fmaf(wt[3 * r + 2].w, wc[r].w, acc.w);
      wa[r] = wb[r];
      wb[r] = wc[r];
    }
    y4[((row + x0 + i) << lq) + q] = acc;
  }
}

// two channels per thread: half the registers of k_dw3_v4, twice the resident warps (the float4 form is latency-bound at 16 warps/SM)
template <int SEG>
__global__ void __launch_bounds__(256, 4) k_dw3_v2(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                float* __restrict__ y, int N, int lh, int lw, int lc, int flip) {
  constexpr int LS = SEG == 16 ? 4 : 3;
  const int H = 1 << lh, W = 1 << lw, lq = lc - 1, CQ = 1 << lq, lseg = lw - LS;
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x, total = (size_t)N << (lh + lseg + lq);
  if (idx >= total) return;
  const int q = (int)(idx & (CQ - 1));
  const size_t s = idx >> lq;
  const int x0 = (int)(s & (((size_t)1 << lseg) - 1)) << LS, yy = (int)((s >> lseg) & (H - 1));
  const size_t row = (s >> lseg) << lw;          // pixel index of (n, yy, 0)
  const bool up = yy > 0, dn = yy + 1 < H;
  float2 wt[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int ts = flip ? 8 - t : t;
    wt[t] = make_float2(__ldg(w + (2 * q) * 9 + ts), __ldg(w + (2 * q + 1) * 9 + ts));
  }
  const float2 zero = make_float2(0.f, 0.f);
  const float2 bias = b ? make_float2(__ldg(b + 2 * q), __ldg(b + 2 * q + 1)) : zero;
  const float2* __restrict__ x4 = reinterpret_cast<const float2*>(x);
  float2* __restrict__ y4 = reinterpret_cast<float2*>(y);
  float2 wa[3], wb[3], wc[3];
  auto column = [&](int xx, float2 (&c)[3]) {
    const bool in = xx >= 0 && xx < W;
    const size_t p = ((row + xx) << lq) + q;
    c[0] = (in && up) ? __ldg(x4 + p - ((size_t)W << lq)) : zero;
    c[1] = in ? __ldg(x4 + p) : zero;
    c[2] = (in && dn) ? __ldg(x4 + p + ((size_t)W << lq)) : zero;
  };
  column(x0 - 1, wa);
  column(x0, wb);
#pragma unroll 4
  for (int i = 0; i < SEG; ++i) {
    column(x0 + i + 1, wc);
    float2 acc = bias;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      acc.x = fmaf(wt[3 * r].x, wa[r].x, acc.x); acc.y = fmaf(wt[3 * r].y, wa[r].y, acc.y);
      acc.x = fmaf(wt[3 * r + 1].x, wb[r].x, acc.x); acc.y = fmaf(wt[3 * r + 1].y, wb[r].y, acc.y);
      acc.x = fmaf(wt[3 * r + 2].x, wc[r].x, acc.x); acc.y = fmaf(wt[3 * r + 2].y, wc[r].y, acc.y);
      wa[r] = wb[r];
      wb[r] = wc[r];
    }
    y4[((row + x0 + i) << lq) + q] = acc;
  }
}

// dw[k][tap] += sum dy[p] * x[p + tap];  db[k] += sum dy[p].  blockDim = 256, C (a power of two) divides 256;
// thread = (pixel slot, channel).
template <int K>
__global__ void __launch_bounds__(256) k_dw_wgrad(TV x, TV dy, float* __restrict__ dw, float* __restrict__ db, int N, int lh,
                                                  int lw, int lc) {
  extern __shared__ float sm[];   // [C][K*K+1]
  constexpr int T = K * K;
  const int H = 1 << lh, W = 1 << lw, C = 1 << lc;
  for (int i = threadIdx.x; i < C * (T + 1); i += 256) sm[i] = 0.f;
  __syncthreads();
  const int k = threadIdx.x & (C - 1), ppb = 256 >> lc;
  const size_t NP = (size_t)N << (lh + lw);
  float acc[T], bacc = 0.f;
#pragma unroll
  for (int i = 0; i < T; ++i) acc[i] = 0.f;
  for (size_t gp = (size_t)blockIdx.x * ppb + (threadIdx.x >> lc); gp < NP; gp += (size_t)gridDim.x * ppb) {
    const int xx = (int)(gp & (W - 1)), yy = (int)((gp >> lw) & (H - 1));
    const float g = dy.p[tv_at(dy, gp, k)];
    bacc += g;
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const int sy = yy + a - K / 2, sx = xx + c - K / 2;
        if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
        acc[a * K + c] = fmaf(g, x.p[tv_at(x, gp + (size_t)(a - K / 2) * W + (c - K / 2), k)], acc[a * K + c]);
      }
  }
#pragma unroll
  for (int i = 0; i < T; ++i) atomicAdd(sm + k * (T + 1) + i, acc[i]);
  atomicAdd(sm + k * (T + 1) + T, bacc);
  __syncthreads();
  for (int i = threadIdx.x; i < C * (T + 1); i += 256) {
    const int kk = i / (T + 1), t = i % (T + 1);
    if (t < T) atomicAdd(dw + kk * T + t, sm[i]);
    else if (db) atomicAdd(db + kk, sm[i]);
  }
}

// The conv-FFN's depthwise 3x3 (LGT.py:101, NHWC with ld == C, W >= 8): thread = (8-pixel row segment, channel quad); a
// 3x3 window of float4 slides along the row, so a pixel costs three x loads and one dy load of 128 bits for four channels.
__device__ __forceinline__ void fma4(float4& a, const float4 g, const float4 v) {
  a.x = fmaf(g.x, v.x, a.x); a.y = fmaf(g.y, v.y, a.y); a.z = fmaf(g.z, v.z, a.z); a.w = fmaf(g.w, v.w, a.w);
}
__global__ void __launch_bounds__(256) k_dw3_wgrad_v4(const float* __restrict__ x, const float* __restrict__ dy,
                                                      float* __restrict__ dw, float* __restrict__ db, int N, int lh, int lw,
                                                      int lc) {
  extern __shared__ float sm[];   // [C][10]
  const int H = 1 << lh, W = 1 << lw, C = 1 << lc, lq = lc - 2, CQ = 1 << lq;
  for (int i = threadIdx.x; i < C * 10; i += 256) sm[i] = 0.f;
  __syncthreads();
  const int q = threadIdx.x & (CQ - 1), slot = threadIdx.x >> lq, slots = 256 >> lq, lseg = lw - 3;
  const size_t nseg = (size_t)N << (lh + lseg);
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
  const float4* __restrict__ g4 = reinterpret_cast<const float4*>(dy);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc[9], bacc = zero;
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = zero;
  for (size_t s = (size_t)blockIdx.x * slots + slot; s < nseg; s += (size_t)gridDim.x * slots) {
    const int x0 = (int)(s & (((size_t)1 << lseg) - 1)) << 3, yy = (int)((s >> lseg) & (H - 1));
    const size_t row = (s >> lseg) << lw;        // pixel index of (n, yy, 0)
    const bool up = yy > 0, dn = yy + 1 < H;
    float4 wa[3], wb[3], wc[3];
    auto column = [&](int xx, float4 (&c)[3]) {
      const bool in = xx >= 0 && xx < W;
      const size_t p = ((row + xx) << lq) + q;
      c[0] = (in && up) ? __ldg(x4 + p - ((size_t)W << lq)) : zero;
      c[1] = in ? __ldg(x4 + p) : zero;
      c[2] = (in && dn) ? __ldg(x4 + p + ((size_t)W << lq)) : zero;
    };
    column(x0 - 1, wa);
    column(x0, wb);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int xx = x0 + i;
      column(xx + 1, wc);
      const float4 g = __ldg(g4 + ((row + xx) << lq) + q);
      bacc.x += g.x; bacc.y += g.y; bacc.z += g.z; bacc.w += g.w;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        fma4(acc[r * 3 + 0], g, wa[r]);
        fma4(acc[r * 3 + 1], g, wb[r]);
        fma4(acc[r * 3 + 2], g, wc[r]);
        wa[r] = wb[r];
        wb[r] = wc[r];
      }
    }
  }
  float* mine = sm + (q << 2) * 10;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    atomicAdd(mine + i, acc[i].x); atomicAdd(mine + 10 + i, acc[i].y);
    atomicAdd(mine + 20 + i, acc[i].z); atomicAdd(mine + 30 + i, acc[i].w);
  }
  atomicAdd(mine + 9, bacc.x); atomicAdd(mine + 19, bacc.y); atomicAdd(mine + 29, bacc.z); atomicAdd(mine + 39, bacc.w);
  __syncthreads();
  for (int i = threadIdx.x; i < C * 10; i += 256) {
    const int kk = i / 10, t = i % 10;
    if (t < 9) atomicAdd(dw + kk * 9 + t, sm[i]);
    else if (db) atomicAdd(db + kk, sm[i]);
  }
}

__device__ __forceinline__ void fma2v(float2& a, const float2 g, const float2 v) { a.x = fmaf(g.x, v.x, a.x); a.y = fmaf(g.y, v.y, a.y); }
// two channels per thread: half the registers, twice the resident warps (as k_dw3_v2)
__global__ void __launch_bounds__(256, 3) k_dw3_wgrad_v2(const float* __restrict__ x, const float* __restrict__ dy,
                                                      float* __restrict__ dw, float* __restrict__ db, int N, int lh, int lw,
                                                      int lc) {
  extern __shared__ float sm[];   // [C][10]
  const int H = 1 << lh, W = 1 << lw, C = 1 << lc, lq = lc - 1, CQ = 1 << lq;
  for (int i = threadIdx.x; i < C * 10; i += 256) sm[i] = 0.f;
  __syncthreads();
  const int q = threadIdx.x & (CQ - 1), slot = threadIdx.x >> lq, slots = 256 >> lq, lseg = lw - 3;
  const size_t nseg = (size_t)N << (lh + lseg);
  const float2* __restrict__ x4 = reinterpret_cast<const float2*>(x);
  const float2* __restrict__ g4 = reinterpret_cast<const float2*>(dy);
  const float2 zero = make_float2(0.f, 0.f);
  float2 acc[9], bacc = zero;
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = zero;
  for (size_t s = (size_t)blockIdx.x * slots + slot; s < nseg; s += (size_t)gridDim.x * slots) {
    const int x0 = (int)(s & (((size_t)1 << lseg) - 1)) << 3, yy = (int)((s >> lseg) & (H - 1));
    const size_t row = (s >> lseg) << lw;        // pixel index of (n, yy, 0)
    const bool up = yy > 0, dn = yy + 1 < H;
    float2 wa[3], wb[3], wc[3];
    auto column = [&](int xx, float2 (&c)[3]) {
      const bool in = xx >= 0 && xx < W;
      const size_t p = ((row + xx) << lq) + q;
      c[0] = (in && up) ? __ldg(x4 + p - ((size_t)W << lq)) : zero;
      c[1] = in ? __ldg(x4 + p) : zero;
      c[2] = (in && dn) ? __ldg(x4 + p + ((size_t)W << lq)) : zero;
    };
    column(x0 - 1, wa);
    column(x0, wb);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int xx = x0 + i;
      column(xx + 1, wc);
      const float2 g = __ldg(g4 + ((row + xx) << lq) + q);
      bacc.x += g.x; bacc.y += g.y;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        fma2v(acc[r * 3 + 0], g, wa[r]);
        fma2v(acc[r * 3 + 1], g, wb[r]);
        fma2v(acc[r * 3 + 2], g, wc[r]);
        wa[r] = wb[r];
        wb[r] = wc[r];
      }
    }
  }
  float* mine = sm + (q << 1) * 10;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    atomicAdd(mine + i, acc[i].x); atomicAdd(mine + 10 + i, acc[i].y);
  }
  atomicAdd(mine + 9, bacc.x); atomicAdd(mine + 19, bacc.y);
  __syncthreads();
  for (int i = threadIdx.x; i < C * 10; i += 256) {
    const int kk = i / 10, t = i % 10;
    if (t < 9) atomicAdd(dw + kk * 9 + t, sm[i]);
    else if (db) atomicAdd(db + kk, sm[i]);
  }
}

// ---- bicubic resize (bmu.sampling_ / sampling_unit_, basic_module_unformer_v2.py:21-34): Keys A = -0.75,
// src = (dst + .5) * rscale - .5, tap indices clamped.  adjoint = 1 scatters y (a gradient) into x with atomics. ----------------
__device__ __forceinline__ void cubic_coef(float t, float (&c)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x3 = 2.f - t, x2 = 1.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}
__global__ void __launch_bounds__(256) k_resize(TV x, int Hi, int Wi, TV y, int Ho, int Wo, int C, int N, float rscale,
                                                int adjoint) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x, Po = (size_t)Ho * Wo, total = (size_t)N * Po * C;
  if (idx >= total) return;
  int k;
  size_t gpo;
  if (y.nchw) { const size_t pix = idx % Po, r = idx / Po; k = (int)(r % C); gpo = (r / C) * Po + pix; }
  else { k = (int)(idx % C); gpo = idx / C; }
  const int ox = (int)(gpo % Wo), oy = (int)((gpo / Wo) % Ho);
  const size_t n = gpo / Po;
  const float sy = (oy + 0.5f) * rscale - 0.5f, sx = (ox + 0.5f) * rscale - 0.5f;
  const float fy = floorf(sy), fx = floorf(sx);
  float cy[4], cx[4];
  cubic_coef(sy - fy, cy);
  cubic_coef(sx - fx, cx);
  const int iy = (int)fy, ix = (int)fx;
  const size_t base = n * (size_t)Hi * Wi;
  if (!adjoint) {
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), Hi - 1);
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), Wi - 1);
        row = fmaf(cx[b], x.p[tv_at(x, base + (size_t)yy * Wi + xx, k)], row);
      }
      acc = fmaf(cy[a], row, acc);
    }
    y.p[tv_at(y, gpo, k)] = acc;
  } else {
    const float g = y.p[tv_at(y, gpo, k)];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), Hi - 1);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), Wi - 1);
        atomicAdd(x.p + tv_at(x, base + (size_t)yy * Wi + xx, k), cy[a] * cx[b] * g);
      }
    }
  }
}

// NCHW planes in and out (the data module): thread = output pixel of one image, looping over the C planes — the tap coefficients and
// the 16 clamped source indices are computed once instead of once per channel (they were ~100 of the ~150 instructions per output).
__global__ void __launch_bounds__(256) k_resize_nchw(const float* __restrict__ x, int Hi, int Wi, float* __restrict__ y, int Ho, int Wo,
                                                     int C, int N, float rscale, int adjoint) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x, Po = (size_t)Ho * Wo, Pi = (size_t)Hi * Wi;
  if (idx >= (size_t)N * Po) return;
  const size_t n = idx / Po, pix = idx - n * Po;
  const int ox = (int)(pix % Wo), oy = (int)(pix / Wo);
  const float sy = (oy + 0.5f) * rscale - 0.5f, sx = (ox + 0.5f) * rscale - 0.5f;
  const float fy = floorf(sy), fx = floorf(sx);
  float cy[4], cx[4];
  cubic_coef(sy - fy, cy);
  cubic_coef(sx - fx, cx);
  const int iy = (int)fy, ix = (int)fx;
  int ro[4], co[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    ro[a] = min(max(iy - 1 + a, 0), Hi - 1) * Wi;
    co[a] = min(max(ix - 1 + a, 0), Wi - 1);
  }
  const float* xp = x + n * C * Pi;
  float* yp = y + n * C * Po + pix;
  for (int k = 0; k < C; ++k, xp += Pi, yp += Po) {
    if (!adjoint) {
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float row = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) row = fmaf(cx[b], __ldg(xp + ro[a] + co[b]), row);
        acc = fmaf(cy[a], row, acc);
      }
      *yp = acc;
    } else {
      const float g = *yp;
      float* xo = const_cast<float*>(xp);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) atomicAdd(xo + ro[a] + co[b], cy[a] * cx[b] * g);
    }
  }
}

// NHWC maps with ld == C, C % 4 == 0: thread = (output pixel, channel quad), 128-bit taps; the adjoint scatters with the
// 128-bit vector atomics of sm_90+.
__global__ void __launch_bounds__(256) k_resize_v4(const float4* __restrict__ x, int Hi, int Wi, float4* __restrict__ y, int Ho, int Wo,
                                                   int CQ, int N, float rscale, int adjoint) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x, Po = (size_t)Ho * Wo, total = (size_t)N * Po * CQ;
  if (idx >= total) return;
  const int q = (int)(idx % CQ);
  const size_t gpo = idx / CQ;
  const int ox = (int)(gpo % Wo), oy = (int)((gpo / Wo) % Ho);
  const size_t n = gpo / Po;
  const float sy = (oy + 0.5f) * rscale - 0.5f, sx = (ox + 0.5f) * rscale - 0.5f;
  const float fy = floorf(sy), fx = floorf(sx);
  float cy[4], cx[4];
  cubic_coef(sy - fy, cy);
  cubic_coef(sx - fx, cx);
  const int iy = (int)fy, ix = (int)fx;
  const size_t base = n * (size_t)Hi * Wi;
  if (!adjoint) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), Hi - 1);
      float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), Wi - 1);
        const float4 v = __ldg(x + (base + (size_t)yy * Wi + xx) * CQ + q);
        row.x = fmaf(cx[b], v.x, row.x); row.y = fmaf(cx[b], v.y, row.y);
        row.z = fmaf(cx[b], v.z, row.z); row.w = fmaf(cx[b], v.w, row.w);
      }
      acc.x = fmaf(cy[a], row.x, acc.x); acc.y = fmaf(cy[a], row.y, acc.y);
      acc.z = fmaf(cy[a], row.z, acc.z); acc.w = fmaf(cy[a], row.w, acc.w);
    }
    y[gpo * CQ + q] = acc;
  } else {
    const float4 g = y[gpo * CQ + q];
    float4* xo = const_cast<float4*>(x);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), Hi - 1);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), Wi - 1);
        const float c = cy[a] * cx[b];
        atomicAdd(xo + (base + (size_t)yy * Wi + xx) * CQ + q, make_float4(c * g.x, c * g.y, c * g.z, c * g.w));
      }
    }
  }
}

// ---- window multi-head self-attention (local_mixer, LGT.py:130-146; window merge :207-208) ------------------------------------------
// qkv [NP, 6D] NHWC (q | k | v thirds, head-major inside a third), D = head dim, 8x8 windows, two heads.
// A 64-thread block owns one window, warp = head, lane = tokens {lane, lane + 32} (as queries, and as keys in the backward):
// every broadcast shared-memory read of a key / value / query row feeds two tokens, which halves the L1 data-pipe traffic that
// bounded the thread-per-query form (k_attn_fwd 210 -> us, k_attn_bwd 514 -> us at 4 x 256 x 256 x 32 channels).  The window's
// qkv rows (8 runs of 8 * 6D contiguous floats) move between HBM and shared memory as coalesced 16-byte pieces.  Scores are
// kept in base 2: q is scaled by D^-1/2 * log2(e) when staged and the positional bias comes pre-multiplied (k_attn_pos).
// The forward also writes the row statistic L = max + log2(sum) per (pixel, head); the backward needs nothing else from it
// (delta = sum_j P dP = dO . O, with O read from the tape), so it has no statistics sweep.

// pos [2][64 queries][64 keys] -> ppos[0 .. 8192): log2(e) * [head][key / 4][query][key % 4]  (a query lane fetches four keys with
// one 16-byte load, a warp reads 512 contiguous bytes);  ppos[8192 .. 16384): log2(e) * pos  (key lanes read consecutive words)
__global__ void k_attn_pos(const float* __restrict__ pos, float* __restrict__ ppos) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= 2 * 64 * 64) return;
  const int h = idx >> 12, jq = (idx >> 8) & 15, i = (idx >> 2) & 63, jr = idx & 3;
  ppos[idx] = pos[(h << 12) + (i << 6) + 4 * jq + jr] * 1.4426950408889634f;
  ppos[8192 + idx] = pos[idx] * 1.4426950408889634f;
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int D>
struct AttnGeo {
  static constexpr int LD = 6 * D;         // floats per pixel of qkv
  static constexpr int LDP = LD + 4;       // padded shared-memory row: 16-byte aligned, own-token reads hit distinct banks
  static constexpr int PPX = LD / 4;       // 16-byte pieces per pixel
  static constexpr int OLD = 2 * D + 4;    // padded row of the dO tile
  static constexpr int OPX = 2 * D / 4;
};
__device__ __forceinline__ size_t attn_pixel(int n, int win, int t, int H, int W) {
  const int nwx = W / 8;
  return ((size_t)n * H + (win / nwx) * 8 + (t >> 3)) * W + (win % nwx) * 8 + (t & 7);
}
template <int D>
__device__ __forceinline__ void ld_row(float (&r)[D], const float* src) {
#pragma unroll
  for (int d = 0; d < D; d += 4) {
    const float4 v = *reinterpret_cast<const float4*>(src + d);
    r[d] = v.x; r[d + 1] = v.y; r[d + 2] = v.z; r[d + 3] = v.w;
  }
}
template <int D>
__device__ __forceinline__ void st_row(float* dst, const float (&r)[D], float mul) {
#pragma unroll
  for (int d = 0; d < D; d += 4) *reinterpret_cast<float4*>(dst + d) = make_float4(r[d] * mul, r[d + 1] * mul, r[d + 2] * mul, r[d + 3] * mul);
}
// packed fp32 (FFMA2): a dot product keeps an even and an odd partial sum, which also lets two additive terms ride along
template <int D>
__device__ __forceinline__ float dot_row(const float (&a)[D], const float (&b)[D], float init_even, float init_odd) {
  float2 s = make_float2(init_even, init_odd);
#pragma unroll
  for (int d = 0; d < D; d += 2) s = __ffma2_rn(make_float2(a[d], a[d + 1]), make_float2(b[d], b[d + 1]), s);
  return s.x + s.y;
}
template <int D>
__device__ __forceinline__ void axpy_row(float (&y)[D], float a, const float (&x)[D]) {
  const float2 aa = make_float2(a, a);
#pragma unroll
  for (int d = 0; d < D; d += 2) {
    const float2 r = __ffma2_rn(aa, make_float2(x[d], x[d + 1]), make_float2(y[d], y[d + 1]));
    y[d] = r.x; y[d + 1] = r.y;
  }
}

template <int D>
__global__ void __launch_bounds__(64) k_attn_fwd(const float* __restrict__ qkv, const float* __restrict__ ppos, TV out,
                                                 float* __restrict__ lse, int N, int H, int W) {
  using G = AttnGeo<D>;
  __shared__ __align__(16) float tile[64 * G::LDP];
  const int nwin = (H / 8) * (W / 8);
  const int win = blockIdx.x % nwin, n = blockIdx.x / nwin;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float qs = rsqrtf((float)D) * 1.4426950408889634f;
  for (int f = threadIdx.x; f < 64 * G::PPX; f += 64) {
    const int t = f / G::PPX, q = f - t * G::PPX;
    float4 v = __ldg(reinterpret_cast<const float4*>(qkv + attn_pixel(n, win, t, H, W) * G::LD) + q);
    if (q < G::PPX / 3) { v.x *= qs; v.y *= qs; v.z *= qs; v.w *= qs; }
    *reinterpret_cast<float4*>(tile + t * G::LDP + 4 * q) = v;
  }
  __syncthreads();
  float q0[D], q1[D], o0[D], o1[D];
  ld_row<D>(q0, tile + lane * G::LDP + h * D);
  ld_row<D>(q1, tile + (lane + 32) * G::LDP + h * D);
#pragma unroll
  for (int d = 0; d < D; ++d) { o0[d] = 0.f; o1[d] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float4* pp = reinterpret_cast<const float4*>(ppos) + (size_t)h * 1024 + lane;   // [key / 4][query] float4
  const float* kb = tile + 2 * D + h * D;
  const float* vb = tile + 4 * D + h * D;
#pragma unroll 1
  for (int c = 0; c < 64; c += 8) {                       // 8 keys per softmax rescale
    float s0[8], s1[8];
#pragma unroll
    for (int jq = 0; jq < 2; ++jq) {
      const float4 b0 = __ldg(pp + (c / 4 + jq) * 64), b1 = __ldg(pp + (c / 4 + jq) * 64 + 32);
      const float bb0[4] = {b0.x, b0.y, b0.z, b0.w}, bb1[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int jr = 0; jr < 4; ++jr) {
        float k[D];
        ld_row<D>(k, kb + (c + 4 * jq + jr) * G::LDP);
        s0[4 * jq + jr] = dot_row<D>(q0, k, bb0[jr], 0.f);
        s1[4 * jq + jr] = dot_row<D>(q1, k, bb1[jr], 0.f);
      }
    }
    float c0 = s0[0], c1 = s1[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) { c0 = fmaxf(c0, s0[j]); c1 = fmaxf(c1, s1[j]); }
    const float n0 = fmaxf(m0, c0), n1 = fmaxf(m1, c1);
    const float r0 = ex2f(m0 - n0), r1 = ex2f(m1 - n1);
    m0 = n0; m1 = n1;
    l0 *= r0; l1 *= r1;
#pragma unroll
    for (int d = 0; d < D; ++d) { o0[d] *= r0; o1[d] *= r1; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v[D];
      ld_row<D>(v, vb + (c + j) * G::LDP);
      const float e0 = ex2f(s0[j] - m0), e1 = ex2f(s1[j] - m1);
      l0 += e0; l1 += e1;
      axpy_row<D>(o0, e0, v);
      axpy_row<D>(o1, e1, v);
    }
  }
  lse[attn_pixel(n, win, lane, H, W) * 2 + h] = m0 + log2f(l0);
  lse[attn_pixel(n, win, lane + 32, H, W) * 2 + h] = m1 + log2f(l1);
  // own q slots (no other lane reads them) carry the output to the coalesced store
  st_row<D>(tile + lane * G::LDP + h * D, o0, 1.f / l0);
  st_row<D>(tile + (lane + 32) * G::LDP + h * D, o1, 1.f / l1);
  __syncthreads();
  for (int f = threadIdx.x; f < 64 * G::OPX; f += 64) {
    const int t = f / G::OPX, q = f - t * G::OPX;
    *reinterpret_cast<float4*>(out.p + attn_pixel(n, win, t, H, W) * out.ld + 4 * q) =
        *reinterpret_cast<const float4*>(tile + t * G::LDP + 4 * q);
  }
}

// backward: persistent 64-thread blocks loop over windows.  Phase 1 (lane = two queries, loop over keys): P from the saved row
// statistic, dP = dO . v, dS = P (dP - delta), dq.  Phase 2 (lane = two keys, loop over queries): the same P and dS column-wise
// for dk, dv and the positional-bias gradient, which accumulates over the block's windows in shared memory (each lane owns its
// two key columns) and is flushed once with 16-byte atomics.  dqkv is written exactly once per element, coalesced.
template <int D>
struct AttnBwdSmem {
  using G = AttnGeo<D>;
  float dps[2 * 64 * 64];          // [head][query][key]
  float tile[64 * G::LDP];         // q (scaled) | k | v, later dq | dk | dv
  float dot[64 * G::OLD];          // dO, both heads
  float2 st[2 * 64];               // [head][query] (L, delta)
};
template <int D>
__global__ void __launch_bounds__(64) k_attn_bwd(const float* __restrict__ qkv, const float* __restrict__ ppos, TV dout, TV o,
                                                 const float* __restrict__ lse, float* __restrict__ dqkv,
                                                 float* __restrict__ dpos, int N, int H, int W) {
  using G = AttnGeo<D>;
  extern __shared__ __align__(16) unsigned char attn_smem_raw[];
  AttnBwdSmem<D>& S = *reinterpret_cast<AttnBwdSmem<D>*>(attn_smem_raw);
  const int nwin = (H / 8) * (W / 8);
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale = rsqrtf((float)D), qs = scale * 1.4426950408889634f;
  for (int f = threadIdx.x; f < 2 * 64 * 64 / 4; f += 64) reinterpret_cast<float4*>(S.dps)[f] = make_float4(0.f, 0.f, 0.f, 0.f);
  float* dps = S.dps + h * 4096 + lane;
  const float4* pp = reinterpret_cast<const float4*>(ppos) + (size_t)h * 1024 + lane;
  const float* pk = ppos + 8192 + h * 4096 + lane;         // log2(e) * pos[head][query][key]
  for (int g = blockIdx.x; g < N * nwin; g += gridDim.x) {
    const int win = g % nwin, n = g / nwin;
    __syncthreads();                                        // the previous window's store has read the tile
    for (int f = threadIdx.x; f < 64 * G::PPX; f += 64) {
      const int t = f / G::PPX, q = f - t * G::PPX;
      float4 v = __ldg(reinterpret_cast<const float4*>(qkv + attn_pixel(n, win, t, H, W) * G::LD) + q);
      if (q < G::PPX / 3) { v.x *= qs; v.y *= qs; v.z *= qs; v.w *= qs; }
      *reinterpret_cast<float4*>(S.tile + t * G::LDP + 4 * q) = v;
    }
    for (int f = threadIdx.x; f < 64 * G::OPX; f += 64) {  // dO tile and delta = dO . O (D / 4 neighbouring lanes per head)
      const int t = f / G::OPX, q = f - t * G::OPX;
      const size_t gp = attn_pixel(n, win, t, H, W);
      const float4 a = __ldg(reinterpret_cast<const float4*>(dout.p + gp * dout.ld) + q);
      const float4 b = __ldg(reinterpret_cast<const float4*>(o.p + gp * o.ld) + q);
      *reinterpret_cast<float4*>(S.dot + t * G::OLD + 4 * q) = a;
      float dl = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
#pragma unroll
      for (int w = 1; w < D / 4; w *= 2) dl += __shfl_xor_sync(0xffffffffu, dl, w);
      if (q % (D / 4) == 0) {
        const int hh = q / (D / 4);
        S.st[hh * 64 + t] = make_float2(__ldg(lse + gp * 2 + hh), dl);
      }
    }
    __syncthreads();
    float ga[D], gb[D];                                    // dq of the two queries
    // ---- phase 1 ---------------------------------------------------------------------------------------------------------------
    {
      float q0[D], q1[D], d0[D], d1[D];
      ld_row<D>(q0, S.tile + lane * G::LDP + h * D);
      ld_row<D>(q1, S.tile + (lane + 32) * G::LDP + h * D);
      ld_row<D>(d0, S.dot + lane * G::OLD + h * D);
      ld_row<D>(d1, S.dot + (lane + 32) * G::OLD + h * D);
      const float2 st0 = S.st[h * 64 + lane], st1 = S.st[h * 64 + lane + 32];
#pragma unroll
      for (int d = 0; d < D; ++d) { ga[d] = 0.f; gb[d] = 0.f; }
      const float* kb = S.tile + 2 * D + h * D;
      const float* vb = S.tile + 4 * D + h * D;
#pragma unroll 2
      for (int jq = 0; jq < 16; ++jq) {
        const float4 b0 = __ldg(pp + jq * 64), b1 = __ldg(pp + jq * 64 + 32);
        const float bb0[4] = {b0.x, b0.y, b0.z, b0.w}, bb1[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int jr = 0; jr < 4; ++jr) {
          float k[D], v[D];
          ld_row<D>(k, kb + (4 * jq + jr) * G::LDP);
          ld_row<D>(v, vb + (4 * jq + jr) * G::LDP);
          const float p0 = ex2f(dot_row<D>(q0, k, bb0[jr], -st0.x)), p1 = ex2f(dot_row<D>(q1, k, bb1[jr], -st1.x));
          const float s0 = p0 * dot_row<D>(d0, v, -st0.y, 0.f), s1 = p1 * dot_row<D>(d1, v, -st1.y, 0.f);
          axpy_row<D>(ga, s0, k);
          axpy_row<D>(gb, s1, k);
        }
      }
    }
    // ---- phase 2 ---------------------------------------------------------------------------------------------------------------
    float dk0[D], dk1[D], dv0[D], dv1[D];
    {
      float ra[D], rb[D], v0[D], v1[D];                    // k, v of the two keys
      ld_row<D>(ra, S.tile + lane * G::LDP + 2 * D + h * D);
      ld_row<D>(rb, S.tile + (lane + 32) * G::LDP + 2 * D + h * D);
      ld_row<D>(v0, S.tile + lane * G::LDP + 4 * D + h * D);
      ld_row<D>(v1, S.tile + (lane + 32) * G::LDP + 4 * D + h * D);
#pragma unroll
      for (int d = 0; d < D; ++d) { dk0[d] = 0.f; dk1[d] = 0.f; dv0[d] = 0.f; dv1[d] = 0.f; }
      const float* qb = S.tile + h * D;
      const float* ob = S.dot + h * D;
#pragma unroll 4
      for (int r = 0; r < 64; ++r) {
        float q[D], dO[D];
        ld_row<D>(q, qb + r * G::LDP);
        ld_row<D>(dO, ob + r * G::OLD);
        const float2 st = S.st[h * 64 + r];
        const float p0 = ex2f(dot_row<D>(q, ra, __ldg(pk + r * 64), -st.x)), p1 = ex2f(dot_row<D>(q, rb, __ldg(pk + r * 64 + 32), -st.x));
        const float s0 = p0 * dot_row<D>(dO, v0, -st.y, 0.f), s1 = p1 * dot_row<D>(dO, v1, -st.y, 0.f);
        dps[r * 64] += s0;
        dps[r * 64 + 32] += s1;
        axpy_row<D>(dk0, s0, q);
        axpy_row<D>(dk1, s1, q);
        axpy_row<D>(dv0, p0, dO);
        axpy_row<D>(dv1, p1, dO);
      }
    }
    __syncwarp();                                           // this head's slices of the tile are read by this warp only
    st_row<D>(S.tile + lane * G::LDP + h * D, ga, scale);
    st_row<D>(S.tile + (lane + 32) * G::LDP + h * D, gb, scale);
    st_row<D>(S.tile + lane * G::LDP + 2 * D + h * D, dk0, 0.6931471805599453f);      // q was staged with log2(e) folded in
    st_row<D>(S.tile + (lane + 32) * G::LDP + 2 * D + h * D, dk1, 0.6931471805599453f);
    st_row<D>(S.tile + lane * G::LDP + 4 * D + h * D, dv0, 1.f);
    st_row<D>(S.tile + (lane + 32) * G::LDP + 4 * D + h * D, dv1, 1.f);
    __syncthreads();
    for (int f = threadIdx.x; f < 64 * G::PPX; f += 64) {
      const int t = f / G::PPX, q = f - t * G::PPX;
      *(reinterpret_cast<float4*>(dqkv + attn_pixel(n, win, t, H, W) * G::LD) + q) =
          *reinterpret_cast<const float4*>(S.tile + t * G::LDP + 4 * q);
    }
  }
  __syncthreads();
  if ((reinterpret_cast<uintptr_t>(dpos) & 15) == 0) {
    for (int f = threadIdx.x; f < 2 * 64 * 64 / 4; f += 64) atomicAdd(reinterpret_cast<float4*>(dpos) + f, reinterpret_cast<const float4*>(S.dps)[f]);
  } else {
    for (int f = threadIdx.x; f < 2 * 64 * 64; f += 64) atomicAdd(dpos + f, S.dps[f]);
  }
}

// ---- FFT passes of the global mixer (LGT.py:162-180) and their adjoints ---------------------------------------------------------
// Shared-memory radix-2 transforms, `cpb` channels of one line per block.  Spectrum layout: complex [N][H][W/2+1][c2].
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__device__ __forceinline__ void fft_twiddles(float2* tw, int L, int dir) {
  for (int k = threadIdx.x; k < L / 2; k += blockDim.x) {
    float s, c;
    sincospif(2.f * k / L, &s, &c);
    tw[k] = make_float2(c, dir < 0 ? -s : s);
  }
}
// in-place transform of nch lines stored bit-reversed at buf[ch*ls + i]; ends with a __syncthreads().
// Two radix-2 levels per sweep (a thread carries four points through levels s and s + 1 in registers): the same butterflies
// and twiddles as the plain radix-2 schedule, hence the same bits, with half the barriers and shared-memory round trips.
__device__ __forceinline__ void fft_pow2(float2* buf, const float2* tw, int L, int logL, int nch, int ls) {
  __syncthreads();
  int s = 1;
  if (logL & 1) {
    for (int b = threadIdx.x; b < nch * (L / 2); b += blockDim.x) {
      const int ch = b / (L / 2), j = b - ch * (L / 2);
      float2* base = buf + ch * ls;
      const float2 a = base[2 * j], t = base[2 * j + 1];          // level 1: twiddle 1
      base[2 * j] = make_float2(a.x + t.x, a.y + t.y);
      base[2 * j + 1] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncthreads();
    s = 2;
  }
  for (; s < logL; s += 2) {
    const int half = 1 << (s - 1), t1 = L >> s, t2 = L >> (s + 1);
    for (int b = threadIdx.x; b < nch * (L / 4); b += blockDim.x) {
      const int ch = b / (L / 4), j = b - ch * (L / 4);
      const int pos = j & (half - 1), i0 = ((j >> (s - 1)) << (s + 1)) + pos;
      float2* base = buf + ch * ls;
      const float2 x0 = base[i0], x1 = base[i0 + half], x2 = base[i0 + 2 * half], x3 = base[i0 + 3 * half];
      const float2 w1 = tw[pos * t1], wa = tw[pos * t2], wb = tw[(pos + half) * t2];
      float2 t = cmul(w1, x1);
      const float2 y0 = make_float2(x0.x + t.x, x0.y + t.y), y1 = make_float2(x0.x - t.x, x0.y - t.y);
      t = cmul(w1, x3);
      const float2 y2 = make_float2(x2.x + t.x, x2.y + t.y), y3 = make_float2(x2.x - t.x, x2.y - t.y);
      t = cmul(wa, y2);
      base[i0] = make_float2(y0.x + t.x, y0.y + t.y);
      base[i0 + 2 * half] = make_float2(y0.x - t.x, y0.y - t.y);
      t = cmul(wb, y3);
      base[i0 + half] = make_float2(y1.x + t.x, y1.y + t.y);
      base[i0 + 3 * half] = make_float2(y1.x - t.x, y1.y - t.y);
    }
    __syncthreads();
  }
}

// mode 0 (R2C): real row (x, optionally times sign(sgn)) -> forward FFT -> spec[k], k <= W/2, times scale and (weight2 ? u_k : 1)
// mode 1 (HALF2R): spec[k] times (weight2 ? u_k : 1), zero padded -> inverse FFT -> real part times scale -> xout (and |.| -> xabs)
//   u_k = 1 for k in {0, W/2}, 2 otherwise: HALF2R with weight2 is the C2R transform of irfft2 (Im of bins 0 and W/2 drops out),
//   R2C with weight2 is its adjoint; HALF2R without weights is the adjoint of the plain R2C.
__global__ void __launch_bounds__(256) k_fft_rows(int mode, const float* __restrict__ xin, int ldx, const float* __restrict__ sgn,
                                                  float2* __restrict__ spec, float* __restrict__ xout, int ldo,
                                                  float* __restrict__ xabs, int ldabs, int H, int W, int logW, int c2, int cpb,
                                                  float scale, int weight2) {
  extern __shared__ float2 fsm[];
  float2* tw = fsm;
  float2* buf = fsm + W / 2;
  const int ls = W + 1, Wh = W / 2 + 1;
  const int y = blockIdx.x, n = blockIdx.y, ch0 = blockIdx.z * cpb;
  const size_t prow = ((size_t)n * H + y) * W, srow = ((size_t)n * H + y) * Wh;
  fft_twiddles(tw, W, mode == 0 ? -1 : 1);
  if (mode == 0) {
    for (int e = threadIdx.x; e < W * cpb; e += blockDim.x) {
      const int ch = e % cpb, i = e / cpb;
      float v = xin[(prow + i) * ldx + ch0 + ch];
      if (sgn) { const float t = sgn[(prow + i) * c2 + ch0 + ch]; v = t > 0.f ? v : (t < 0.f ? -v : 0.f); }
      buf[ch * ls + (__brev(i) >> (32 - logW))] = make_float2(v, 0.f);
    }
    fft_pow2(buf, tw, W, logW, cpb, ls);
    for (int e = threadIdx.x; e < Wh * cpb; e += blockDim.x) {
      const int ch = e % cpb, k = e / cpb;
      const float f = scale * ((weight2 && k > 0 && k < W / 2) ? 2.f : 1.f);
      const float2 v = buf[ch * ls + k];
      spec[(srow + k) * c2 + ch0 + ch] = make_float2(v.x * f, v.y * f);
    }
  } else {
    for (int e = threadIdx.x; e < W * cpb; e += blockDim.x) {
      const int ch = e % cpb, i = e / cpb;
      float2 v = make_float2(0.f, 0.f);
      if (i < Wh) {
        v = spec[(srow + i) * c2 + ch0 + ch];
        if (weight2 && i > 0 && i < W / 2) { v.x *= 2.f; v.y *= 2.f; }
      }
      buf[ch * ls + (__brev(i) >> (32 - logW))] = v;
    }
    fft_pow2(buf, tw, W, logW, cpb, ls);
    for (int e = threadIdx.x; e < W * cpb; e += blockDim.x) {
      const int ch = e % cpb, i = e / cpb;
      const float v = buf[ch * ls + i].x * scale;
      xout[(prow + i) * ldo + ch0 + ch] = v;
      if (xabs) xabs[(prow + i) * ldabs + ch0 + ch] = fabsf(v);
    }
  }
}

// in-place complex FFT along H of spec[n][:, kx, ch]; dir = -1 forward, +1 unnormalised inverse.
// fixreal: the four purely real bins get an exact +0.0 imaginary part (what rfft2 delivers).
__global__ void __launch_bounds__(256) k_fft_cols(float2* __restrict__ spec, int H, int logH, int W, int c2, int cpb, int dir,
                                                  int fixreal) {
  extern __shared__ float2 fsm[];
  float2* tw = fsm;
  float2* buf = fsm + H / 2;
  const int ls = H + 1, Wh = W / 2 + 1;
  const int kx = blockIdx.x, n = blockIdx.y, ch0 = blockIdx.z * cpb;
  fft_twiddles(tw, H, dir);
  for (int e = threadIdx.x; e < H * cpb; e += blockDim.x) {
    const int ch = e % cpb, i = e / cpb;
    buf[ch * ls + (__brev(i) >> (32 - logH))] = spec[(((size_t)n * H + i) * Wh + kx) * c2 + ch0 + ch];
  }
  fft_pow2(buf, tw, H, logH, cpb, ls);
  const bool realcol = fixreal && (kx == 0 || kx == W / 2);
  for (int e = threadIdx.x; e < H * cpb; e += blockDim.x) {
    const int ch = e % cpb, i = e / cpb;
    float2 v = buf[ch * ls + i];
    if (realcol && (i == 0 || i == H / 2)) v.y = 0.f;
    spec[(((size_t)n * H + i) * Wh + kx) * c2 + ch0 + ch] = v;
  }
}

// amplitude / phase mixing (LGT.py:168-177): G = complex(amp cos(pha) + 1e-8, amp sin(pha) + 1e-8) + 1e-8,
// amp = |F| wa + ba, pha = angle(F) wp + bp  (depthwise 1x1 convs = per-channel affine)
__global__ void __launch_bounds__(256) k_spec_mix(const float2* __restrict__ F, float2* __restrict__ G, size_t total, int c2,
                                                  const float* __restrict__ wa, const float* __restrict__ ba,
                                                  const float* __restrict__ wp, const float* __restrict__ bp) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % c2);
  const float2 f = F[idx];
  const float amp = sqrtf(f.x * f.x + f.y * f.y) * wa[ch] + ba[ch];
  const float pha = atan2f(f.y, f.x) * wp[ch] + bp[ch];
  float sn, cs;
  sincosf(pha, &sn, &cs);
  G[idx] = make_float2((amp * cs + 1e-8f) + 1e-8f, amp * sn + 1e-8f);
}
// backward: dG (in) -> dF (out, in place); per-channel parameter gradients accumulate.  gridDim.x * 256 is a multiple of c2.
__global__ void __launch_bounds__(256) k_spec_mix_bwd(const float2* __restrict__ F, float2* __restrict__ dG, size_t total, int c2,
                                                      const float* __restrict__ wa, const float* __restrict__ ba,
                                                      const float* __restrict__ wp, const float* __restrict__ bp,
                                                      float* __restrict__ dwa, float* __restrict__ dba,
                                                      float* __restrict__ dwp, float* __restrict__ dbp) {
  extern __shared__ float sm[];   // [4][c2]
  for (int i = threadIdx.x; i < 4 * c2; i += 256) sm[i] = 0.f;
  __syncthreads();
  const int ch = threadIdx.x % c2;
  const float a_w = wa[ch], a_b = ba[ch], p_w = wp[ch], p_b = bp[ch];
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
  for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
    const float2 f = F[idx], g = dG[idx];
    const float m2 = f.x * f.x + f.y * f.y, mag = sqrtf(m2), ang = atan2f(f.y, f.x);
    const float amp = mag * a_w + a_b, pha = ang * p_w + p_b;
    float sn, cs;
    sincosf(pha, &sn, &cs);
    const float d_amp = g.x * cs + g.y * sn, d_pha = amp * (g.y * cs - g.x * sn);
    g0 = fmaf(d_amp, mag, g0);
    g1 += d_amp;
    g2 = fmaf(d_pha, ang, g2);
    g3 += d_pha;
    const float dmag = d_amp * a_w, dang = d_pha * p_w;
    float2 o = make_float2(0.f, 0.f);
    if (mag > 0.f) {
      const float im = 1.f / mag, im2 = 1.f / m2;
      o.x = dmag * f.x * im - dang * f.y * im2;
      o.y = dmag * f.y * im + dang * f.x * im2;
    }
    dG[idx] = o;
  }
  atomicAdd(sm + ch, g0);
  atomicAdd(sm + c2 + ch, g1);
  atomicAdd(sm + 2 * c2 + ch, g2);
  atomicAdd(sm + 3 * c2 + ch, g3);
  __syncthreads();
  for (int i = threadIdx.x; i < c2; i += 256) {
    atomicAdd(dwa + i, sm[i]);
    atomicAdd(dba + i, sm[c2 + i]);
    atomicAdd(dwp + i, sm[2 * c2 + i]);
    atomicAdd(dbp + i, sm[3 * c2 + i]);
  }
}

// ---- dropout of the mixer projection (nn.Dropout(0.1), LGT.py:198,216) + the residual add (LGT.py:45-51) ------------------------
// The keep mask is a counter-based hash of (seed, layer, element index): the backward regenerates it instead of storing it.
__device__ __forceinline__ float drop_scale(uint64_t seed, int layer, size_t idx, float p) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull + (uint64_t)(layer + 1) * 0xD1B54A32D192ED03ull;
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.f / 16777216.f);
  return u >= p ? 1.f / (1.f - p) : 0.f;
}
// mode 0: y = a + drop(b);  mode 1: y = drop(b);  mode 2: y = the mask itself (tests).  ext != NULL: a recorded mask
// (values 0 or 1/(1-p)) replaces the generated one.
__global__ void __launch_bounds__(256) k_dropout(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                                                 size_t n, uint64_t seed, int layer, float p, int mode,
                                                 const float* __restrict__ ext, const uint64_t* __restrict__ seed_dev = nullptr) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  if (seed_dev) seed = *seed_dev;                      // a captured step reads the seed of the replay, not of the capture
  const float m = ext ? ext[i] : (p > 0.f ? drop_scale(seed, layer, i, p) : 1.f);
  y[i] = mode == 0 ? a[i] + m * b[i] : mode == 1 ? m * b[i] : m;
}

// four elements per thread (n % 4 == 0, 16-byte aligned pointers): the same mask, element by element
__global__ void __launch_bounds__(256) k_dropout_v4(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ y,
                                                    size_t n4, uint64_t seed, int layer, float p, int mode,
                                                    const float4* __restrict__ ext, const uint64_t* __restrict__ seed_dev) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  if (seed_dev) seed = *seed_dev;
  float4 m;
  if (ext) m = __ldg(ext + i);
  else if (p > 0.f) m = make_float4(drop_scale(seed, layer, 4 * i, p), drop_scale(seed, layer, 4 * i + 1, p),
                                    drop_scale(seed, layer, 4 * i + 2, p), drop_scale(seed, layer, 4 * i + 3, p));
  else m = make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 bv = __ldg(b + i);
  float4 r = make_float4(m.x * bv.x, m.y * bv.y, m.z * bv.z, m.w * bv.w);
  if (mode == 0) {
    const float4 av = __ldg(a + i);
    r.x += av.x; r.y += av.y; r.z += av.z; r.w += av.w;
  }
  y[i] = r;
}

// ---- data module (models/unlg_former.py:29-40,58-61), NCHW ------------------------------------------------------------------------
// r = R(Z) - pan
__global__ void __launch_bounds__(256) k_data_r(const float* __restrict__ Z, const float* __restrict__ pan,
                                                const float* __restrict__ rw, const float* __restrict__ rb, float* __restrict__ r,
                                                int N, int B, size_t P) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (size_t)N * P) return;
  const size_t n = idx / P, pix = idx % P;
  float acc = rb[0];
  for (int b = 0; b < B; ++b) acc = fmaf(rw[b], Z[(n * B + b) * P + pix], acc);
  r[idx] = acc - pan[idx];
}
// Zout = Z - eta * (T1 + RT(r))
__global__ void __launch_bounds__(256) k_data_update(const float* __restrict__ Z, const float* __restrict__ T1,
                                                     const float* __restrict__ r, const float* __restrict__ rtw,
                                                     const float* __restrict__ rtb, const float* __restrict__ eta,
                                                     float* __restrict__ Zout, int N, int B, size_t P) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (size_t)N * B * P) return;
  const size_t pix = idx % P, n = idx / (P * B);
  const int b = (int)((idx / P) % B);
  Zout[idx] = Z[idx] - eta[0] * (T1[idx] + fmaf(rtw[b], r[n * P + pix], rtb[b]));
}
// backward of the update: gz (dZout, in/out) -> s = -eta gz (gradient of T1), gz += R^T(RT^T s); parameter gradients.
// red: [1 + 2B + B + 1] accumulators = deta, drtw[B], drtb[B], drw[B], drb
__global__ void __launch_bounds__(256) k_data_update_bwd(float* __restrict__ gz, const float* __restrict__ Z,
                                                         const float* __restrict__ T1, const float* __restrict__ r,
                                                         const float* __restrict__ rw, const float* __restrict__ rtw,
                                                         const float* __restrict__ rtb, const float* __restrict__ eta,
                                                         float* __restrict__ s, float* __restrict__ deta, float* __restrict__ drtw,
                                                         float* __restrict__ drtb, float* __restrict__ drw, float* __restrict__ drb,
                                                         int N, int B, size_t P) {
  __shared__ float red[26];
  if (threadIdx.x < 26) red[threadIdx.x] = 0.f;
  __syncthreads();
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  const bool live = idx < (size_t)N * P;
  const size_t n = live ? idx / P : 0, pix = live ? idx % P : 0;
  const float e = eta[0], rv = live ? r[idx] : 0.f;
  float de = 0.f, dr = 0.f;
  for (int b = 0; b < B; ++b) {
    float g = 0.f, sb = 0.f;
    if (live) {
      const size_t o = (n * B + b) * P + pix;
      g = gz[o];
      de -= g * (T1[o] + fmaf(rtw[b], rv, rtb[b]));
      sb = -e * g;
      s[o] = sb;
      dr = fmaf(rtw[b], sb, dr);
    }
    float a = sb * rv, c = sb;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(red + 1 + b, a); atomicAdd(red + 1 + B + b, c); }
  }
  for (int b = 0; b < B; ++b) {
    float a = 0.f;
    if (live) {
      const size_t o = (n * B + b) * P + pix;
      a = dr * Z[o];
      gz[o] += rw[b] * dr;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(red + 1 + 2 * B + b, a);
  }
  float d2 = dr;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { de += __shfl_xor_sync(0xffffffffu, de, o); d2 += __shfl_xor_sync(0xffffffffu, d2, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(red, de); atomicAdd(red + 1 + 3 * B, d2); }
  __syncthreads();
  if (threadIdx.x == 0) { atomicAdd(deta, red[0]); atomicAdd(drb, red[1 + 3 * B]); }
  if (threadIdx.x < B) {
    atomicAdd(drtw + threadIdx.x, red[1 + threadIdx.x]);
    atomicAdd(drtb + threadIdx.x, red[1 + B + threadIdx.x]);
    atomicAdd(drw + threadIdx.x, red[1 + 2 * B + threadIdx.x]);
  }
}

// max |x| as the bit pattern of a non-negative float (monotone as unsigned), then the power of two 2^-ceil(log2 max)
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ x, size_t n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && m < INFINITY) atomicMax(out, __float_as_uint(m));
}
__global__ void k_pow2_scale(const unsigned* __restrict__ maxbits, float* __restrict__ scale) {
  const float m = __uint_as_float(*maxbits);
  *scale = m > 0.f ? exp2f(-ceilf(log2f(m))) : 1.f;
}

// ---- loss and optimiser ------------------------------------------------------------------------------------------------------------
// nn.L1Loss (mean) times weight (models/base/losses.py:29,39; loss_cfg rec_loss.w): loss += w/n sum |out - gt|; dout = w/n sign(.)
__global__ void __launch_bounds__(256) k_l1(const float* __restrict__ out, const float* __restrict__ gt, size_t n, float wn,
                                            float* __restrict__ loss, float* __restrict__ dout) {
  __shared__ float red[8];
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float d = out[i] - gt[i];
    acc += fabsf(d);
    if (dout) dout[i] = d > 0.f ? wn : (d < 0.f ? -wn : 0.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(loss, t * wn);
  }
}
// torch.optim.Adam (no weight decay, no amsgrad), one fused pass over the flat parameter buffer; g is pre-scaled by gscale
// (1 / world size after the gradient all-reduce).
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps,
                                              float bc1, float bc2_sqrt, float gscale) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * gscale;
  const float mi = b1 * m[i] + (1.f - b1) * gi, vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
}

}  // namespace lgtrain
