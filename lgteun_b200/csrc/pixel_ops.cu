// pixel_ops.cu — the U-Net glue of the LGT prior (models/common/LGT.py), one thread per output pixel,
// NHWC inside, NCHW only at the prior's boundary.  All HBM-bound (<10 FLOP/B): one read of the inputs,
// one write of the output, weights broadcast from shared memory.
//   patch_embed : LGT.py:72-88   depthwise 1x1 -> 1x1 B->C -> LayerNorm(C)          NCHW -> NHWC
//   down        : LGT.py:280-281 bicubic 1/2 -> 1x1 C->2C                            NHWC -> NHWC (half res)
//   up_fuse     : LGT.py:294-295,336-338  bicubic x2 -> 1x1 2C->C ; cat(up, skip) -> 1x1 2C->C
//   tail        : LGT.py:302-303,342      bicubic x1 (identity taps [0,1,0,0]) -> 1x1 C->B ; + x   NHWC -> NCHW
#include "common.cuh"

namespace lg {

// ---- patch embedding -------------------------------------------------------------------------------
template <int B>
__global__ void __launch_bounds__(256) patch_embed_kernel(const float* __restrict__ x, float* __restrict__ y, PriorW w,
                                                           int HW, long long total) {
  constexpr int C = 4 * B;
  __shared__ float sW[C * B], sB[C], sG[C], sBeta[C], sDw[B], sDb[B];
  for (int i = threadIdx.x; i < C * B; i += 256) sW[i] = __ldg(w.pe_w + i);
  for (int i = threadIdx.x; i < C; i += 256) {
    sB[i] = __ldg(w.pe_b + i); sG[i] = __ldg(w.pe_ln_w + i); sBeta[i] = __ldg(w.pe_ln_b + i);
  }
  if (threadIdx.x < B) { sDw[threadIdx.x] = __ldg(w.pe_dw_w + threadIdx.x); sDb[threadIdx.x] = __ldg(w.pe_dw_b + threadIdx.x); }
  __syncthreads();
  long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= total) return;
  long long n = p / HW;
  int r = (int)(p - n * HW);
  float t[B];
#pragma unroll
  for (int b = 0; b < B; ++b) t[b] = fmaf(sDw[b], __ldg(x + (n * B + b) * HW + r), sDb[b]);
  float v[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int b = 0; b < B; ++b) acc = fmaf(sW[c * B + b], t[b], acc);
    v[c] = acc + sB[c];
  }
  layer_norm_inplace<C>(v, sG, sBeta);
  store_vec<C>(y + p * C, v);
}

cudaError_t launch_patch_embed(const PriorW& w, int B, const float* x, float* y, int N, int H, int W, cudaStream_t s) {
  long long total = (long long)N * H * W;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (B == 4) patch_embed_kernel<4><<<grid, 256, 0, s>>>(x, y, w, H * W, total);
  else if (B == 8) patch_embed_kernel<8><<<grid, 256, 0, s>>>(x, y, w, H * W, total);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ---- dense per-pixel matvec with weights [NOUT][K] in shared memory ---------------------------------
// out[o] = bias[o] + sum_k W[o][k] v[k]; four outputs at a time, weights as 128-bit broadcasts, dot products on the
// packed fp32 pipe (two k per FFMA2).  sW must be 16-byte aligned and K a multiple of 4.
template <int K, int NOUT>
__device__ __forceinline__ void matvec_store(const float (&v)[K], const float* __restrict__ sW,
                                             const float* __restrict__ sBias, float* __restrict__ dst) {
  static_assert(K % 4 == 0 && NOUT % 4 == 0, "matvec tiles are 4x4");
#pragma unroll 1
  for (int o = 0; o < NOUT; o += 4) {
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    const float4* r0 = reinterpret_cast<const float4*>(sW + (o + 0) * K);
    const float4* r1 = reinterpret_cast<const float4*>(sW + (o + 1) * K);
    const float4* r2 = reinterpret_cast<const float4*>(sW + (o + 2) * K);
    const float4* r3 = reinterpret_cast<const float4*>(sW + (o + 3) * K);
#pragma unroll
    for (int k4 = 0; k4 < K / 4; ++k4) {
      const float2 va = make_float2(v[4 * k4], v[4 * k4 + 1]), vb = make_float2(v[4 * k4 + 2], v[4 * k4 + 3]);
      const float4 w0 = r0[k4], w1 = r1[k4], w2 = r2[k4], w3 = r3[k4];
      a0 = __ffma2_rn(make_float2(w0.x, w0.y), va, a0); a0 = __ffma2_rn(make_float2(w0.z, w0.w), vb, a0);
      a1 = __ffma2_rn(make_float2(w1.x, w1.y), va, a1); a1 = __ffma2_rn(make_float2(w1.z, w1.w), vb, a1);
      a2 = __ffma2_rn(make_float2(w2.x, w2.y), va, a2); a2 = __ffma2_rn(make_float2(w2.z, w2.w), vb, a2);
      a3 = __ffma2_rn(make_float2(w3.x, w3.y), va, a3); a3 = __ffma2_rn(make_float2(w3.z, w3.w), vb, a3);
    }
    *reinterpret_cast<float4*>(dst + o) = make_float4((a0.x + a0.y) + sBias[o], (a1.x + a1.y) + sBias[o + 1],
                                                      (a2.x + a2.y) + sBias[o + 2], (a3.x + a3.y) + sBias[o + 3]);
  }
}

// ---- down unit -----------------------------------------------------------------------------------------
// bicubic 1/2 (taps [-3,19,19,-3]/32, clamped) then 1x1 C->2C.  C/4 lanes share one output pixel: a lane owns four
// channels during the resize (its 16 tap loads are 128-bit and coalesced with its neighbours') and eight output
// channels during the 1x1 conv (inputs exchanged through shared memory).
constexpr int kDownIters = 8;
template <int C>
__global__ void __launch_bounds__(128) down_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                    const float* __restrict__ wt, const float* __restrict__ bias,
                                                    int H, int W, long long total) {
  constexpr int LPP = C / 4;                       // lanes per pixel
  constexpr int PPB = 128 / LPP;                   // pixels per block
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                                // [C][2C]  (transposed at load time: PriorW::down_wt)
  float* sB = sW + 2 * C * C;                      // [2C]
  float* sV = sB + 2 * C;                          // [PPB][C + 4]  resized pixel vectors (padded rows)
  for (int i = threadIdx.x; i < 2 * C * C / 4; i += 128) reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(wt) + i);
  for (int i = threadIdx.x; i < 2 * C; i += 128) sB[i] = __ldg(bias + i);
  const int oh = H / 2, ow = W / 2;
  const int lp = threadIdx.x / LPP, q = threadIdx.x % LPP;
#pragma unroll 1
  for (int it = 0; it < kDownIters; ++it) {        // several pixel groups per CTA amortise the weight staging
  const long long p = ((long long)blockIdx.x * kDownIters + it) * PPB + lp;
  const bool live = p < total;
  __syncthreads();                                 // weights staged / previous group's vectors consumed
  if (live) {
    const int ox = (int)(p % ow);
    const long long t = p / ow;
    const int oy = (int)(t % oh);
    const long long n = t / oh;
    const float dn[4] = {-0.09375f, 0.59375f, 0.59375f, -0.09375f};
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int gy = clampi(2 * oy - 1 + a, 0, H - 1);
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int gx = clampi(2 * ox - 1 + b, 0, W - 1);
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((n * H + gy) * W + gx) * C) + q);
        r.x = fmaf(dn[b], v.x, r.x); r.y = fmaf(dn[b], v.y, r.y); r.z = fmaf(dn[b], v.z, r.z); r.w = fmaf(dn[b], v.w, r.w);
      }
      acc.x = fmaf(dn[a], r.x, acc.x); acc.y = fmaf(dn[a], r.y, acc.y);
      acc.z = fmaf(dn[a], r.z, acc.z); acc.w = fmaf(dn[a], r.w, acc.w);
    }
    *reinterpret_cast<float4*>(sV + lp * (C + 4) + 4 * q) = acc;
  }
  __syncthreads();
  if (!live) continue;
  // 2C outputs per pixel, 8 per lane (o = 8q .. 8q+7).  Weights are staged TRANSPOSED [k][2C]: for a fixed k the four
  // lanes of a pixel read one contiguous 128-byte row and all pixels of the warp read the same row (one wavefront).
  const float* v = sV + lp * (C + 4);
  float2 acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = make_float2(sB[8 * q + 2 * j], sB[8 * q + 2 * j + 1]);
#pragma unroll
  for (int k4 = 0; k4 < C / 4; ++k4) {
    const float4 v4 = *reinterpret_cast<const float4*>(v + 4 * k4);
    const float vk[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4* wr = reinterpret_cast<const float4*>(sW + (4 * k4 + kk) * (2 * C) + 8 * q);
      const float4 wa = wr[0], wb = wr[1];
      const float2 vv = make_float2(vk[kk], vk[kk]);
      acc[0] = __ffma2_rn(make_float2(wa.x, wa.y), vv, acc[0]);
      acc[1] = __ffma2_rn(make_float2(wa.z, wa.w), vv, acc[1]);
      acc[2] = __ffma2_rn(make_float2(wb.x, wb.y), vv, acc[2]);
      acc[3] = __ffma2_rn(make_float2(wb.z, wb.w), vv, acc[3]);
    }
  }
  float4* dst = reinterpret_cast<float4*>(y + p * (2 * C) + 8 * q);
  dst[0] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
  dst[1] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
  }
}

cudaError_t launch_down(const PriorW& w, int C, const float* x, float* y, int N, int H, int W, cudaStream_t s) {
  long long total = (long long)N * (H / 2) * (W / 2);
  if (C == 16) {
    constexpr int PPB = 128 / 4;
    size_t smem = (size_t)(2 * 16 * 16 + 2 * 16 + PPB * 20) * sizeof(float);
    down_kernel<16><<<(unsigned)((total + PPB * kDownIters - 1) / (PPB * kDownIters)), 128, smem, s>>>(x, y, w.down_wt, w.down_b, H, W, total);
  } else if (C == 32) {
    constexpr int PPB = 128 / 8;
    size_t smem = (size_t)(2 * 32 * 32 + 2 * 32 + PPB * 36) * sizeof(float);
    down_kernel<32><<<(unsigned)((total + PPB * kDownIters - 1) / (PPB * kDownIters)), 128, smem, s>>>(x, y, w.down_wt, w.down_b, H, W, total);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---- up unit + skip fusion ---------------------------------------------------------------------------
// reference (LGT.py:294-295,336-338): fea = up_conv(bicubic_x2(low)); y = fuse_conv(cat[fea, skip]).
// A per-pixel affine map commutes with the bicubic resize (its taps sum to 1 and index clamping is linear), so both
// the 2C->C up conv and the [:, :C] half of the fusion conv run at LOW resolution (4x fewer pixels; fp32 rounding
// differs by ~1e-7 relative):   T = Wf[:, :C] (Wu low + bu)          (low res, low_conv2_kernel)
//                               y = bicubic_x2(T) + Wf[:, C:] skip + bf   (full res, up_fuse_kernel)
// The full-res kernel gives each thread a 2x2 output block: the four pixels share a 5x5 low-res neighbourhood
// (25 shared-memory vector loads per block instead of 4 x 16).
template <int C>
__global__ void __launch_bounds__(128) low_conv2_kernel(const float* __restrict__ x, float* __restrict__ y, PriorW w,
                                                         long long total) {
  extern __shared__ __align__(16) float smem[];
  float* sU = smem;                 // [C][2C]   up conv
  float* sF = sU + 2 * C * C;       // [C][C]    left half of the fusion conv
  float* sUb = sF + C * C;          // [C]
  for (int i = threadIdx.x; i < 2 * C * C; i += 128) sU[i] = __ldg(w.up_w + i);
  for (int i = threadIdx.x; i < C * C; i += 128) sF[i] = __ldg(w.fuse_w + (i / C) * 2 * C + (i % C));
  for (int i = threadIdx.x; i < C; i += 128) sUb[i] = __ldg(w.up_b + i);
  __syncthreads();
  long long p = (long long)blockIdx.x * 128 + threadIdx.x;
  if (p >= total) return;
  float v[2 * C];
  load_vec<2 * C>(v, x + p * 2 * C);
  float t[C];
#pragma unroll
  for (int o = 0; o < C; o += 4) {
    float2 a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k4 = 0; k4 < 2 * C / 4; ++k4) {
      const float2 va = make_float2(v[4 * k4], v[4 * k4 + 1]), vb = make_float2(v[4 * k4 + 2], v[4 * k4 + 3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w4 = *reinterpret_cast<const float4*>(sU + (o + j) * 2 * C + 4 * k4);
        a[j] = __ffma2_rn(make_float2(w4.x, w4.y), va, a[j]);
        a[j] = __ffma2_rn(make_float2(w4.z, w4.w), vb, a[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) t[o + j] = (a[j].x + a[j].y) + sUb[o + j];
  }
  // second map (no bias: the fusion bias is added at full resolution)
#pragma unroll
  for (int o = 0; o < C; o += 4) {
    float2 a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k4 = 0; k4 < C / 4; ++k4) {
      const float2 va = make_float2(t[4 * k4], t[4 * k4 + 1]), vb = make_float2(t[4 * k4 + 2], t[4 * k4 + 3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w4 = *reinterpret_cast<const float4*>(sF + (o + j) * C + 4 * k4);
        a[j] = __ffma2_rn(make_float2(w4.x, w4.y), va, a[j]);
        a[j] = __ffma2_rn(make_float2(w4.z, w4.w), vb, a[j]);
      }
    }
    *reinterpret_cast<float4*>(y + p * C + o) = make_float4(a[0].x + a[0].y, a[1].x + a[1].y, a[2].x + a[2].y, a[3].x + a[3].y);
  }
}

constexpr int UFH = 8, UFW = 32;                  // output tile (rows x cols) = 64 blocks of 2x2 pixels
constexpr int UFPH = UFH / 2 + 4, UFPW = UFW / 2 + 4;   // low-res patch 12 x 20

// Work item = (2x2 output block, 4-channel quad): the low-res patch, the skip tile and the transposed skip half of the
// fusion conv sit in shared memory, every access is a 128-bit vector of one quad, all arithmetic is packed fp32.
// (The first version gave a thread a whole 2x2 block x 16 channels: 148 registers, 12 warps/SM, 27 % of HBM peak.)
template <int C>
__global__ void __launch_bounds__(256) up_fuse_kernel(const float* __restrict__ t_low, const float* __restrict__ skip,
                                                       float* __restrict__ y, PriorW w, int H, int W) {
  constexpr int PS = C + 4;                       // padded pixel stride: conflict-free 128-bit reads
  constexpr int Q = C / 4;                        // channel quads
  extern __shared__ __align__(16) float smem[];
  float* sP = smem;                               // [UFPH*UFPW][PS]   low-res patch (borders replicated)
  float* sS = sP + UFPH * UFPW * PS;              // [UFH*UFW][PS]     skip tile
  float* sWt = sS + UFH * UFW * PS;               // [C][C]  Wt[k][o] = right (skip) half of the fusion conv
  float* sFb = sWt + C * C;                       // [C]
  const int tid = threadIdx.x;
  const int lh = H / 2, lw = W / 2;
  const int n = blockIdx.z;
  const int Y0 = blockIdx.y * UFH, X0 = blockIdx.x * UFW;
  const int py0 = Y0 / 2 - 2, px0 = X0 / 2 - 2;   // low-res origin of the patch
  for (int i = tid; i < C * C; i += 256) sWt[i] = __ldg(w.fuse_w + (i % C) * 2 * C + C + (i / C));
  for (int i = tid; i < C; i += 256) sFb[i] = __ldg(w.fuse_b + i);
  for (int i = tid; i < UFPH * UFPW * Q; i += 256) {
    const int pix = i / Q, c4 = i - pix * Q;
    const int pr = pix / UFPW, pc = pix - pr * UFPW;
    const int gy = clampi(py0 + pr, 0, lh - 1), gx = clampi(px0 + pc, 0, lw - 1);
    *reinterpret_cast<float4*>(sP + pix * PS + 4 * c4) =
        __ldg(reinterpret_cast<const float4*>(t_low + (((size_t)n * lh + gy) * lw + gx) * C) + c4);
  }
  for (int i = tid; i < UFH * UFW * Q; i += 256) {
    const int pix = i / Q, c4 = i - pix * Q;
    const int oy = Y0 + pix / UFW, ox = X0 + pix % UFW;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (oy < H && ox < W) v = __ldg(reinterpret_cast<const float4*>(skip + (((size_t)n * H + oy) * W + ox) * C) + c4);
    *reinterpret_cast<float4*>(sS + pix * PS + 4 * c4) = v;
  }
  __syncthreads();
  // x2 taps: even dst 2q -> src q-2..q+1 (t=.75), odd dst 2q+1 -> src q-1..q+2 (t=.25); patch rows/cols by+r, bx+c
  // hold src q-2+r: even uses r = 0..3, odd uses r = 1..4
  const float te[4] = {-0.03515625f, 0.26171875f, 0.87890625f, -0.10546875f};
  const float to[4] = {-0.10546875f, 0.87890625f, 0.26171875f, -0.03515625f};
  struct F4 { float2 a, b; };                     // one quad on the packed pipe
  auto fma4 = [](float s, const float4& t, F4& acc) {
    const float2 s2 = make_float2(s, s);
    acc.a = __ffma2_rn(s2, make_float2(t.x, t.y), acc.a);
    acc.b = __ffma2_rn(s2, make_float2(t.z, t.w), acc.b);
  };
  auto fma4f = [](float s, const F4& t, F4& acc) {
    const float2 s2 = make_float2(s, s);
    acc.a = __ffma2_rn(s2, t.a, acc.a);
    acc.b = __ffma2_rn(s2, t.b, acc.b);
  };
#pragma unroll 1
  for (int item = tid; item < (UFH / 2) * (UFW / 2) * Q; item += 256) {
    const int blk = item / Q, oq = item - blk * Q;
    const int by = blk >> 4, bx = blk & 15;       // 2x2 block: low-res cell (Y0/2 + by, X0/2 + bx)
    F4 acc[2][2];
    {
      const float4 b4 = *reinterpret_cast<const float4*>(sFb + 4 * oq);
#pragma unroll
      for (int ey = 0; ey < 2; ++ey)
#pragma unroll
        for (int ex = 0; ex < 2; ++ex) { acc[ey][ex].a = make_float2(b4.x, b4.y); acc[ey][ex].b = make_float2(b4.z, b4.w); }
    }
    // skip half of the fusion conv: 4 outputs x C inputs for the four pixels of the block
#pragma unroll 1
    for (int k4 = 0; k4 < Q; ++k4) {
      float4 wq[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) wq[j] = *reinterpret_cast<const float4*>(sWt + (4 * k4 + j) * C + 4 * oq);
#pragma unroll
      for (int ey = 0; ey < 2; ++ey)
#pragma unroll
        for (int ex = 0; ex < 2; ++ex) {
          const float4 sk = *reinterpret_cast<const float4*>(sS + ((2 * by + ey) * UFW + 2 * bx + ex) * PS + 4 * k4);
          fma4(sk.x, wq[0], acc[ey][ex]);
          fma4(sk.y, wq[1], acc[ey][ex]);
          fma4(sk.z, wq[2], acc[ey][ex]);
          fma4(sk.w, wq[3], acc[ey][ex]);
        }
    }
    // bicubic x2 of the low-res map, separable inside the 5x5 neighbourhood the four pixels share
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      F4 rs_e, rs_o;
      rs_e.a = rs_e.b = rs_o.a = rs_o.b = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(sP + ((by + r) * UFPW + bx + c) * PS + 4 * oq);
        if (c < 4) fma4(te[c < 4 ? c : 0], t, rs_e);
        if (c >= 1) fma4(to[c >= 1 ? c - 1 : 0], t, rs_o);
      }
      if (r < 4) {
        fma4f(te[r < 4 ? r : 0], rs_e, acc[0][0]);
        fma4f(te[r < 4 ? r : 0], rs_o, acc[0][1]);
      }
      if (r >= 1) {
        fma4f(to[r >= 1 ? r - 1 : 0], rs_e, acc[1][0]);
        fma4f(to[r >= 1 ? r - 1 : 0], rs_o, acc[1][1]);
      }
    }
#pragma unroll
    for (int ey = 0; ey < 2; ++ey)
#pragma unroll
      for (int ex = 0; ex < 2; ++ex) {
        const int oy = Y0 + 2 * by + ey, ox = X0 + 2 * bx + ex;
        if (oy >= H || ox >= W) continue;
        *reinterpret_cast<float4*>(y + (((size_t)n * H + oy) * W + ox) * C + 4 * oq) =
            make_float4(acc[ey][ex].a.x, acc[ey][ex].a.y, acc[ey][ex].b.x, acc[ey][ex].b.y);
      }
  }
}

cudaError_t launch_up_fuse(const PriorW& w, int C, const float* low, const float* skip, float* t_low, float* y, int N, int H,
                           int W, cudaStream_t s) {
  const long long low_px = (long long)N * (H / 2) * (W / 2);
  const unsigned g0 = (unsigned)((low_px + 127) / 128);
  dim3 grid((W + UFW - 1) / UFW, (H + UFH - 1) / UFH, N);
  if (C == 16) {
    low_conv2_kernel<16><<<g0, 128, (size_t)(3 * 16 * 16 + 16) * sizeof(float), s>>>(low, t_low, w, low_px);
    size_t smem = (size_t)((UFPH * UFPW + UFH * UFW) * (16 + 4) + 16 * 16 + 16) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(up_fuse_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    up_fuse_kernel<16><<<grid, 256, smem, s>>>(t_low, skip, y, w, H, W);
  } else if (C == 32) {
    low_conv2_kernel<32><<<g0, 128, (size_t)(3 * 32 * 32 + 32) * sizeof(float), s>>>(low, t_low, w, low_px);
    size_t smem = (size_t)((UFPH * UFPW + UFH * UFW) * (32 + 4) + 32 * 32 + 32) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(up_fuse_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    up_fuse_kernel<32><<<grid, 256, smem, s>>>(t_low, skip, y, w, H, W);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---- tail + global residual ----------------------------------------------------------------------------
template <int B>
__global__ void __launch_bounds__(256) tail_kernel(const float* __restrict__ fea, const float* __restrict__ x,
                                                    float* __restrict__ y, const float* __restrict__ wt,
                                                    const float* __restrict__ bias, int HW, long long total) {
  constexpr int C = 4 * B;
  __shared__ float sW[B * C], sB[B];
  for (int i = threadIdx.x; i < B * C; i += 256) sW[i] = __ldg(wt + i);
  if (threadIdx.x < B) sB[threadIdx.x] = __ldg(bias + threadIdx.x);
  __syncthreads();
  long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= total) return;
  long long n = p / HW;
  int r = (int)(p - n * HW);
  float v[C];
  load_vec<C>(v, fea + p * C);
#pragma unroll
  for (int b = 0; b < B; ++b) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) acc = fmaf(sW[b * C + c], v[c], acc);
    size_t o = (size_t)(n * B + b) * HW + r;
    y[o] = (acc + sB[b]) + __ldg(x + o);
  }
}

cudaError_t launch_tail(const PriorW& w, int B, const float* fea, const float* x, float* y, int N, int H, int W,
                        cudaStream_t s) {
  long long total = (long long)N * H * W;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (B == 4) tail_kernel<4><<<grid, 256, 0, s>>>(fea, x, y, w.tail_w, w.tail_b, H * W, total);
  else if (B == 8) tail_kernel<8><<<grid, 256, 0, s>>>(fea, x, y, w.tail_w, w.tail_b, H * W, total);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// pos_emb [2][i][j] -> log2(e) * [2][j][i]: a warp of consecutive queries reads consecutive words, and the softmax
// of window_msa.cu runs in base 2 (ex2.approx)
__global__ void transpose_pos_kernel(const float* __restrict__ pos, float* __restrict__ pos_t) {
  int idx = blockIdx.x * 256 + threadIdx.x;       // over [2][64][64] of the destination
  if (idx >= 2 * 64 * 64) return;
  // destination layout [head][key/4][query][key%4]: one 128-bit load per query thread fetches four keys
  int h = idx >> 12, jq = (idx >> 8) & 15, i = (idx >> 2) & 63, jr = idx & 3;
  int j = 4 * jq + jr;
  pos_t[idx] = pos[(h << 12) + (i << 6) + j] * 1.4426950408889634f;
}
cudaError_t launch_transpose_pos(const float* pos, float* pos_t, cudaStream_t s) {
  transpose_pos_kernel<<<32, 256, 0, s>>>(pos, pos_t);
  return cudaGetLastError();
}

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= rows * cols) return;
  int c = idx / rows, r = idx - c * rows;          // idx enumerates dst [cols][rows]
  dst[idx] = src[(size_t)r * cols + c];
}
cudaError_t launch_transpose(const float* src, float* dst, int rows, int cols, cudaStream_t s) {
  transpose_kernel<<<(rows * cols + 255) / 256, 256, 0, s>>>(src, dst, rows, cols);
  return cudaGetLastError();
}

}  // namespace lg
