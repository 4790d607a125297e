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

namespace lgtrain {

struct TV {
  float* p;
  int ld;     // NHWC: floats per pixel
  int nchw;   // 1: p[((n*C + k)*P + pix)]
  int C, P;   // NCHW only: channels, pixels per image
};
__host__ __device__ __forceinline__ size_t tv_at(const TV& t, size_t gp, int k) {
  return t.nchw ? ((gp / t.P) * t.C + k) * (size_t)t.P + gp % t.P : gp * (size_t)t.ld + k;
}
inline TV nhwc(const float* p, int ld) { return TV{const_cast<float*>(p), ld, 0, 0, 0}; }
inline TV nchw(const float* p, int C, int P) { return TV{const_cast<float*>(p), 0, 1, C, P}; }

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// ---- pointwise (1x1) convolution:  y[p,co] = sum_ci W[co*wso + ci*wsi] * f(x[p,ci]) + b[co]  (+ add[p,co]) (* gelu'(gate)) -------
// Forward of bmu.point_conv (basic_module_unformer_v2.py:13-14) with (wso, wsi) = (Cin, 1); its data gradient with the
// transposed strides (1, Cout_fwd).  ACT = 1 applies the exact GELU to x while staging it (the conv-FFN's activations,
// LGT.py:97,99, are never stored).
template <int ACT>
__global__ void __launch_bounds__(256) k_pw(TV x, int Cin, const float* __restrict__ W, int wso, int wsi,
                                            const float* __restrict__ b, TV y, int Cout, size_t NP, TV add, int use_add,
                                            TV gate, int use_gate) {
  extern __shared__ float sm[];
  constexpr int TPX = 64;
  const size_t p0 = (size_t)blockIdx.x * TPX;
  const int ldx = Cin + 1;
  for (int i = threadIdx.x; i < TPX * Cin; i += 256) {
    int px, ci;
    if (x.nchw) { px = i % TPX; ci = i / TPX; } else { px = i / Cin; ci = i - px * Cin; }
    const size_t gp = p0 + px;
    float v = 0.f;
    if (gp < NP) { v = x.p[tv_at(x, gp, ci)]; if (ACT) v = gelu_exact(v); }
    sm[px * ldx + ci] = v;
  }
  __syncthreads();
  const int px = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const size_t gp = p0 + px;
  const float* xr = sm + px * ldx;
  for (int co0 = grp * 4; co0 < Cout; co0 += 16) {
    float acc[4];
    const float* wr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = min(co0 + j, Cout - 1);
      acc[j] = b ? b[co] : 0.f;
      wr[j] = W + (size_t)co * wso;
    }
    for (int ci = 0; ci < Cin; ++ci) {
      const float xv = xr[ci];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(xv, __ldg(wr[j] + (size_t)ci * wsi), acc[j]);
    }
    if (gp < NP) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = co0 + j;
        if (co < Cout) {
          float v = acc[j];
          if (use_gate) v *= gelu_grad(gate.p[tv_at(gate, gp, co)]);
          if (use_add) v += add.p[tv_at(add, gp, co)];
          y.p[tv_at(y, gp, co)] = v;
        }
      }
    }
  }
}

// weight / bias gradient of the same conv:  dW[co*wso + ci*wsi] += sum_p dy[p,co] * f(x[p,ci]);  db[co] += sum_p dy[p,co].
// grid = (pixel-tile groups, ceil(Cin*Cout / 4096)); every thread owns up to 16 (co, ci) pairs in registers over all the
// tiles of its block and issues one atomicAdd per pair at the end.
template <int ACT>
__global__ void __launch_bounds__(256) k_pw_wgrad(TV x, int Cin, TV dy, int Cout, float* __restrict__ dW, int wso, int wsi,
                                                  float* __restrict__ db, size_t NP) {
  extern __shared__ float sm[];
  constexpr int TPX = 32;
  const int ldx = Cin + 1, ldy = Cout + 1;
  float* xs = sm;
  float* dys = sm + TPX * ldx;
  const int pairs = Cin * Cout, chunk0 = blockIdx.y * 4096;
  int pco[16], pci[16];
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int pair = chunk0 + j * 256 + threadIdx.x;
    acc[j] = 0.f;
    if (pair < pairs) { pco[j] = pair / Cin; pci[j] = pair - pco[j] * Cin; } else { pco[j] = -1; pci[j] = 0; }
  }
  float bacc = 0.f;
  const bool do_bias = db && blockIdx.y == 0 && threadIdx.x < Cout;
  const size_t tiles = (NP + TPX - 1) / TPX;
  for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const size_t p0 = t * TPX;
    for (int i = threadIdx.x; i < TPX * Cin; i += 256) {
      int px, ci;
      if (x.nchw) { px = i % TPX; ci = i / TPX; } else { px = i / Cin; ci = i - px * Cin; }
      const size_t gp = p0 + px;
      float v = 0.f;
      if (gp < NP) { v = x.p[tv_at(x, gp, ci)]; if (ACT) v = gelu_exact(v); }
      xs[px * ldx + ci] = v;
    }
    for (int i = threadIdx.x; i < TPX * Cout; i += 256) {
      int px, co;
      if (dy.nchw) { px = i % TPX; co = i / TPX; } else { px = i / Cout; co = i - px * Cout; }
      const size_t gp = p0 + px;
      dys[px * ldy + co] = gp < NP ? dy.p[tv_at(dy, gp, co)] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (pco[j] >= 0) {
        const float* a = dys + pco[j];
        const float* bb = xs + pci[j];
        float s = 0.f;
#pragma unroll 8
        for (int p = 0; p < TPX; ++p) s = fmaf(a[p * ldy], bb[p * ldx], s);
        acc[j] += s;
      }
    }
    if (do_bias) {
      float s = 0.f;
      for (int p = 0; p < TPX; ++p) s += dys[p * ldy + threadIdx.x];
      bacc += s;
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (pco[j] >= 0) atomicAdd(dW + (size_t)pco[j] * wso + (size_t)pci[j] * wsi, acc[j]);
  if (do_bias) atomicAdd(db + threadIdx.x, bacc);
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

// ---- depthwise KxK convolution with zero padding (bmu.dep_conv, basic_module_unformer_v2.py:17-18), K in {1, 3} ---------------
// flip = 1 uses the point-reflected taps: the data gradient of the same conv.
template <int K>
__global__ void __launch_bounds__(256) k_dw(TV x, const float* __restrict__ w, const float* __restrict__ b, TV y, int N,
                                            int H, int W, int C, int flip, TV add, float add_scale, int use_add) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x, P = (size_t)H * W, total = (size_t)N * P * C;
  if (idx >= total) return;
  int k;
  size_t gp;
  if (y.nchw) { const size_t pix = idx % P, r = idx / P; k = (int)(r % C); gp = (r / C) * P + pix; }
  else { k = (int)(idx % C); gp = idx / C; }
  const int xx = (int)(gp % W), yy = (int)((gp / W) % H);
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

// dw[k][tap] += sum dy[p] * x[p + tap];  db[k] += sum dy[p].  blockDim = 256, C divides 256; thread = (pixel slot, channel).
template <int K>
__global__ void __launch_bounds__(256) k_dw_wgrad(TV x, TV dy, float* __restrict__ dw, float* __restrict__ db, int N, int H,
                                                  int W, int C) {
  extern __shared__ float sm[];   // [C][K*K+1]
  constexpr int T = K * K;
  for (int i = threadIdx.x; i < C * (T + 1); i += 256) sm[i] = 0.f;
  __syncthreads();
  const int k = threadIdx.x % C, ppb = 256 / C;
  const size_t NP = (size_t)N * H * W;
  float acc[T], bacc = 0.f;
#pragma unroll
  for (int i = 0; i < T; ++i) acc[i] = 0.f;
  for (size_t gp = (size_t)blockIdx.x * ppb + threadIdx.x / C; gp < NP; gp += (size_t)gridDim.x * ppb) {
    const int xx = (int)(gp % W), yy = (int)((gp / W) % H);
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

// ---- window multi-head self-attention (local_mixer, LGT.py:130-146; window merge :207-208) ------------------------------------------
// qkv [NP, 6D] NHWC (q | k | v thirds, head-major inside a third), D = head dim; one 64-thread block per (window, head),
// thread = query token i*8+j.
template <int D>
__global__ void __launch_bounds__(64) k_attn_fwd(const float* __restrict__ qkv, const float* __restrict__ pos, TV out, int N,
                                                 int H, int W) {
  __shared__ float ks[64][D], vs[64][D];
  const int nwx = W / 8, nwin = (H / 8) * nwx, c2 = 2 * D, ld = 3 * c2;
  const int head = blockIdx.y, win = blockIdx.x % nwin, n = blockIdx.x / nwin;
  const int i = threadIdx.x;
  const size_t gp = ((size_t)n * H + (win / nwx) * 8 + i / 8) * W + (win % nwx) * 8 + i % 8;
  const float* row = qkv + gp * ld + head * D;
  float q[D];
  const float scale = rsqrtf((float)D);
#pragma unroll
  for (int d = 0; d < D; ++d) { q[d] = row[d] * scale; ks[i][d] = row[c2 + d]; vs[i][d] = row[2 * c2 + d]; }
  __syncthreads();
  const float* pr = pos + ((size_t)head * 64 + i) * 64;
  float s[64], mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) a = fmaf(q[d], ks[j][d], a);
    s[j] = a + pr[j];
    mx = fmaxf(mx, s[j]);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) { s[j] = __expf(s[j] - mx); sum += s[j]; }
  const float inv = 1.f / sum;
  float o[D];
#pragma unroll
  for (int d = 0; d < D; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j)
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = fmaf(s[j], vs[j][d], o[d]);
#pragma unroll
  for (int d = 0; d < D; ++d) out.p[gp * out.ld + head * D + d] = o[d] * inv;
}

// backward: recomputes the probabilities; dqkv is written exactly once per element, dpos accumulates over windows in shared
// memory (the block loops over windows) and is flushed with one atomicAdd per entry.
template <int D>
__global__ void __launch_bounds__(64) k_attn_bwd(const float* __restrict__ qkv, const float* __restrict__ pos, TV dout,
                                                 float* __restrict__ dqkv, float* __restrict__ dpos, int N, int H, int W) {
  extern __shared__ float sm[];
  float* qs = sm;                 // [64][D]  unscaled q
  float* ks = qs + 64 * D;
  float* vs = ks + 64 * D;
  float* dos = vs + 64 * D;
  float* Ps = dos + 64 * D;       // [64][65]
  float* dSs = Ps + 64 * 65;
  float* dps = dSs + 64 * 65;
  const int nwx = W / 8, nwin = (H / 8) * nwx, c2 = 2 * D, ld = 3 * c2;
  const int head = blockIdx.y, i = threadIdx.x;
  const float scale = rsqrtf((float)D);
  const float* pr = pos + ((size_t)head * 64 + i) * 64;
  for (int j = 0; j < 64; ++j) dps[i * 65 + j] = 0.f;
  for (int g = blockIdx.x; g < N * nwin; g += gridDim.x) {
    const int win = g % nwin, n = g / nwin;
    const size_t gp = ((size_t)n * H + (win / nwx) * 8 + i / 8) * W + (win % nwx) * 8 + i % 8;
    const float* row = qkv + gp * ld + head * D;
    float q[D], dO[D];
    __syncthreads();              // previous window's phase 2 is done with the shared tiles
#pragma unroll
    for (int d = 0; d < D; ++d) {
      q[d] = row[d] * scale;
      qs[i * D + d] = row[d];
      ks[i * D + d] = row[c2 + d];
      vs[i * D + d] = row[2 * c2 + d];
      dO[d] = dout.p[gp * dout.ld + head * D + d];
      dos[i * D + d] = dO[d];
    }
    __syncthreads();
    float s[64], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) a = fmaf(q[d], ks[j * D + d], a);
      s[j] = a + pr[j];
      mx = fmaxf(mx, s[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) { s[j] = __expf(s[j] - mx); sum += s[j]; }
    const float inv = 1.f / sum;
    float delta = 0.f;
    float dp[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      s[j] *= inv;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) a = fmaf(dO[d], vs[j * D + d], a);
      dp[j] = a;
      delta = fmaf(s[j], a, delta);
    }
    float dq[D];
#pragma unroll
    for (int d = 0; d < D; ++d) dq[d] = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      const float ds = s[j] * (dp[j] - delta);
      Ps[i * 65 + j] = s[j];
      dSs[i * 65 + j] = ds;
      dps[i * 65 + j] += ds;
#pragma unroll
      for (int d = 0; d < D; ++d) dq[d] = fmaf(ds, ks[j * D + d], dq[d]);
    }
    __syncthreads();
    // phase 2: thread = key token
    float dk[D], dv[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
    for (int r = 0; r < 64; ++r) {
      const float ds = dSs[r * 65 + i], p = Ps[r * 65 + i];
#pragma unroll
      for (int d = 0; d < D; ++d) { dk[d] = fmaf(ds, qs[r * D + d], dk[d]); dv[d] = fmaf(p, dos[r * D + d], dv[d]); }
    }
    float* orow = dqkv + gp * ld + head * D;
#pragma unroll
    for (int d = 0; d < D; ++d) { orow[d] = dq[d] * scale; orow[c2 + d] = dk[d] * scale; orow[2 * c2 + d] = dv[d]; }
  }
  float* dpr = dpos + ((size_t)head * 64 + i) * 64;
  for (int j = 0; j < 64; ++j) atomicAdd(dpr + j, dps[i * 65 + j]);
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
// in-place transform of nch lines stored bit-reversed at buf[ch*ls + i]; ends with a __syncthreads()
__device__ __forceinline__ void fft_pow2(float2* buf, const float2* tw, int L, int logL, int nch, int ls) {
  __syncthreads();
  for (int s = 1; s <= logL; ++s) {
    const int half = 1 << (s - 1), tstep = L >> s;
    for (int b = threadIdx.x; b < nch * (L / 2); b += blockDim.x) {
      const int ch = b / (L / 2), j = b - ch * (L / 2);
      const int pos = j & (half - 1), i0 = ((j >> (s - 1)) << s) + pos, i1 = i0 + half;
      float2* base = buf + ch * ls;
      const float2 a = base[i0], t = cmul(tw[pos * tstep], base[i1]);
      base[i0] = make_float2(a.x + t.x, a.y + t.y);
      base[i1] = make_float2(a.x - t.x, a.y - t.y);
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
                                                 const float* __restrict__ ext) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float m = ext ? ext[i] : (p > 0.f ? drop_scale(seed, layer, i, p) : 1.f);
  y[i] = mode == 0 ? a[i] + m * b[i] : mode == 1 ? m * b[i] : m;
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
