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
template <int C>
__global__ void __launch_bounds__(128) down_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                    const float* __restrict__ wt, const float* __restrict__ bias,
                                                    int H, int W, long long total) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                 // [2C][C]
  float* sB = smem + 2 * C * C;     // [2C]
  for (int i = threadIdx.x; i < 2 * C * C; i += 128) sW[i] = __ldg(wt + i);
  for (int i = threadIdx.x; i < 2 * C; i += 128) sB[i] = __ldg(bias + i);
  __syncthreads();
  const int oh = H / 2, ow = W / 2;
  long long p = (long long)blockIdx.x * 128 + threadIdx.x;
  if (p >= total) return;
  int ox = (int)(p % ow);
  long long q = p / ow;
  int oy = (int)(q % oh);
  long long n = q / oh;
  const float dn[4] = {-0.09375f, 0.59375f, 0.59375f, -0.09375f};
  float v[C];
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int gy = clampi(2 * oy - 1 + a, 0, H - 1);
    float r[C];
#pragma unroll
    for (int c = 0; c < C; ++c) r[c] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int gx = clampi(2 * ox - 1 + b, 0, W - 1);
      float t[C];
      load_vec<C>(t, x + ((n * H + gy) * W + gx) * C);
#pragma unroll
      for (int c = 0; c < C; ++c) r[c] = fmaf(dn[b], t[c], r[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = fmaf(dn[a], r[c], v[c]);
  }
  matvec_store<C, 2 * C>(v, sW, sB, y + p * (2 * C));
}

cudaError_t launch_down(const PriorW& w, int C, const float* x, float* y, int N, int H, int W, cudaStream_t s) {
  long long total = (long long)N * (H / 2) * (W / 2);
  unsigned grid = (unsigned)((total + 127) / 128);
  size_t smem = (size_t)(2 * C * C + 2 * C) * sizeof(float);
  if (C == 16) down_kernel<16><<<grid, 128, smem, s>>>(x, y, w.down_w, w.down_b, H, W, total);
  else if (C == 32) down_kernel<32><<<grid, 128, smem, s>>>(x, y, w.down_w, w.down_b, H, W, total);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ---- up unit + skip fusion ---------------------------------------------------------------------------
// reference (LGT.py:294-295,336-338): fea = up_conv(bicubic_x2(low)); y = fuse_conv(cat[fea, skip]).
// A per-pixel affine map commutes with the bicubic resize (its taps sum to 1 and index clamping is linear), so the
// 2C->C conv runs at LOW resolution (4x fewer pixels, fp32 rounding differs by ~1e-7 relative), then one kernel does
// bicubic x2 from a shared-memory patch, the concat with the skip map and the 2C->C fusion conv.
template <int K, int NOUT>
__global__ void __launch_bounds__(128) pw_conv_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                       const float* __restrict__ wt, const float* __restrict__ bias,
                                                       long long total) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                 // [NOUT][K]
  float* sB = smem + NOUT * K;
  for (int i = threadIdx.x; i < NOUT * K; i += 128) sW[i] = __ldg(wt + i);
  for (int i = threadIdx.x; i < NOUT; i += 128) sB[i] = __ldg(bias + i);
  __syncthreads();
  long long p = (long long)blockIdx.x * 128 + threadIdx.x;
  if (p >= total) return;
  float v[K];
  load_vec<K>(v, x + p * K);
  matvec_store<K, NOUT>(v, sW, sB, y + p * NOUT);
}

constexpr int UFH = 8, UFW = 32;                  // output tile (rows x cols), one thread per pixel
constexpr int UFPH = UFH / 2 + 4, UFPW = UFW / 2 + 4;   // low-res patch 8 x 20

template <int C>
__global__ void __launch_bounds__(256) up_fuse_kernel(const float* __restrict__ t_low, const float* __restrict__ skip,
                                                       float* __restrict__ y, PriorW w, int H, int W) {
  constexpr int PS = C + 4;                       // padded pixel stride: conflict-free 128-bit reads
  extern __shared__ __align__(16) float smem[];
  float* sP = smem;                               // [UFPH*UFPW][PS]
  float* sFu = sP + UFPH * UFPW * PS;             // [C][2C]
  float* sFb = sFu + 2 * C * C;                   // [C]
  const int tid = threadIdx.x;
  const int lh = H / 2, lw = W / 2;
  const int n = blockIdx.z;
  const int Y0 = blockIdx.y * UFH, X0 = blockIdx.x * UFW;
  const int py0 = Y0 / 2 - 2, px0 = X0 / 2 - 2;   // low-res origin of the patch (replicated borders)
  for (int i = tid; i < 2 * C * C; i += 256) sFu[i] = __ldg(w.fuse_w + i);
  for (int i = tid; i < C; i += 256) sFb[i] = __ldg(w.fuse_b + i);
  for (int i = tid; i < UFPH * UFPW * (C / 4); i += 256) {
    const int pix = i / (C / 4), c4 = i - pix * (C / 4);
    const int pr = pix / UFPW, pc = pix - pr * UFPW;
    const int gy = clampi(py0 + pr, 0, lh - 1), gx = clampi(px0 + pc, 0, lw - 1);
    *reinterpret_cast<float4*>(sP + pix * PS + 4 * c4) =
        __ldg(reinterpret_cast<const float4*>(t_low + (((size_t)n * lh + gy) * lw + gx) * C) + c4);
  }
  __syncthreads();
  const int ty = tid >> 5, tx = tid & 31;
  const int oy = Y0 + ty, ox = X0 + tx;
  if (oy >= H || ox >= W) return;
  // x2 taps: even dst 2q -> src q-2..q+1 (t=.75), odd dst 2q+1 -> src q-1..q+2 (t=.25)
  const float te[4] = {-0.03515625f, 0.26171875f, 0.87890625f, -0.10546875f};
  const float to[4] = {-0.10546875f, 0.87890625f, 0.26171875f, -0.03515625f};
  const int fy = (oy >> 1) - ((oy & 1) ? 1 : 2) - py0, fx = (ox >> 1) - ((ox & 1) ? 1 : 2) - px0;
  float cat[2 * C];                               // [upsampled | skip]
#pragma unroll
  for (int c = 0; c < C; ++c) cat[c] = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float wy = (oy & 1) ? to[a] : te[a];
    float r[C];
#pragma unroll
    for (int c = 0; c < C; ++c) r[c] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float wx = (ox & 1) ? to[b] : te[b];
      const float* src = sP + ((fy + a) * UFPW + fx + b) * PS;
#pragma unroll
      for (int c4 = 0; c4 < C; c4 += 4) {
        const float4 t = *reinterpret_cast<const float4*>(src + c4);
        r[c4] = fmaf(wx, t.x, r[c4]); r[c4 + 1] = fmaf(wx, t.y, r[c4 + 1]);
        r[c4 + 2] = fmaf(wx, t.z, r[c4 + 2]); r[c4 + 3] = fmaf(wx, t.w, r[c4 + 3]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) cat[c] = fmaf(wy, r[c], cat[c]);
  }
  const size_t p = ((size_t)n * H + oy) * W + ox;
  {
    float t[C];
    load_vec<C>(t, skip + p * C);
#pragma unroll
    for (int c = 0; c < C; ++c) cat[C + c] = t[c];
  }
  matvec_store<2 * C, C>(cat, sFu, sFb, y + p * C);
}

cudaError_t launch_up_fuse(const PriorW& w, int C, const float* low, const float* skip, float* t_low, float* y, int N, int H,
                           int W, cudaStream_t s) {
  const long long low_px = (long long)N * (H / 2) * (W / 2);
  const unsigned g0 = (unsigned)((low_px + 127) / 128);
  dim3 grid((W + UFW - 1) / UFW, (H + UFH - 1) / UFH, N);
  if (C == 16) {
    pw_conv_kernel<32, 16><<<g0, 128, (32 * 16 + 16) * sizeof(float), s>>>(low, t_low, w.up_w, w.up_b, low_px);
    size_t smem = (size_t)(UFPH * UFPW * (16 + 4) + 2 * 16 * 16 + 16) * sizeof(float);
    up_fuse_kernel<16><<<grid, 256, smem, s>>>(t_low, skip, y, w, H, W);
  } else if (C == 32) {
    pw_conv_kernel<64, 32><<<g0, 128, (64 * 32 + 32) * sizeof(float), s>>>(low, t_low, w.up_w, w.up_b, low_px);
    size_t smem = (size_t)(UFPH * UFPW * (32 + 4) + 2 * 32 * 32 + 32) * sizeof(float);
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
