// data_step.cu — the data module of one unfolding stage (models/unlg_former.py:29-40, 58-61):
//     Z' = Z - eta_i * ( DT(D(Z) - ms) + RT(R(Z) - pan) )
// D  = [bicubic 1/2 -> depthwise 3x3 (+bias, zero pad)] x2      (unlg_former.py:29-30)
// DT = [bicubic x2  -> depthwise 3x3 (+bias, zero pad)] x2      (unlg_former.py:32-33)
// R  = 1x1 B->1, RT = 1x1 1->B                                  (unlg_former.py:36-37)
// bicubic = F.interpolate(mode='bicubic', align_corners=False)  (basic_module_unformer_v2.py:21-34):
// Keys kernel A=-0.75, src = (dst+0.5)/scale-0.5, tap indices clamped to the image.
//
// HBM-bound.  Two launches per stage:
//   data_down : Z (full res)  -> resid = D(Z) - ms        (1/16 of the pixels; the four ops of D chained in
//               shared memory on an 82x82 -> 40 -> 38 -> 18 -> 16 halo pyramid, no intermediate in HBM)
//   data_up   : resid, Z, pan -> Z'                         (DT chain 14 -> 22 -> 20 -> 34 -> 32 in shared
//               memory per channel, fused with R/RT, the eta axpy and the store)
// Mixed border rules are kept exactly: a level consumed by a bicubic resize is stored with replicated
// (index-clamped) borders, a level consumed by a depthwise conv is stored with literal zeros outside.
#include "common.cuh"

namespace lg {

// Keys cubic convolution weights exactly as ATen evaluates them (fp32, A = -0.75).
__device__ __forceinline__ void cubic_weights(float t, float (&w)[4]) {
  const float A = -0.75f;
  float x0 = t + 1.0f, x1 = t, x2 = 1.0f - t, x3 = 2.0f - t;
  w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  w[1] = ((A + 2.0f) * x1 - (A + 3.0f)) * x1 * x1 + 1.0f;
  w[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

// ------------------------------------------------------------------------------------------------
// generic bicubic resize of NCHW planes by num/den in {4/1, 2/1, 1/2}  (bmu.sampling_)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bicubic_kernel(const float* __restrict__ x, float* __restrict__ y, int h, int w,
                                                       int oh, int ow, int num, int den) {
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
  const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= ow || oy >= oh) return;
  const float* xp = x + (size_t)blockIdx.z * h * w;
  const float inv = (float)den / (float)num;             // 1/scale_factor (exact for the supported scales)
  float sy = inv * (oy + 0.5f) - 0.5f, sx = inv * (ox + 0.5f) - 0.5f;
  float fy = floorf(sy), fx = floorf(sx);
  float wy[4], wx[4];
  cubic_weights(sy - fy, wy);
  cubic_weights(sx - fx, wx);
  int iy = (int)fy, ix = (int)fx;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float* row = xp + (size_t)clampi(iy - 1 + a, 0, h - 1) * w;
    float r = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) r = fmaf(wx[b], __ldg(row + clampi(ix - 1 + b, 0, w - 1)), r);
    acc = fmaf(wy[a], r, acc);
  }
  y[((size_t)blockIdx.z * oh + oy) * ow + ox] = acc;
}

// x4 (the Z0 initialisation, unlg_former.py:53): the four outputs 4i .. 4i+3 of one row read the five sources i-2 .. i+2
// (phases t = .625, .875 from i-1 and t = .125, .375 from i), so one thread produces them from 4 x 5 loads and stores one
// 128-bit vector.  Same weights and the same fma order as the generic kernel: bit-identical results, ~5x fewer
// instructions (the generic form was issue-bound at 77 %).
__global__ void __launch_bounds__(256) bicubic_x4_kernel(const float* __restrict__ x, float* __restrict__ y, int h, int w) {
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);    // source column
  const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (i >= w || oy >= 4 * h) return;
  const float* xp = x + (size_t)blockIdx.z * h * w;
  const float sy = 0.25f * (oy + 0.5f) - 0.5f, fy = floorf(sy);
  float wy[4];
  cubic_weights(sy - fy, wy);
  const int iy = (int)fy;
  float wx[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float sx = 0.25f * ((float)(4 * i + p) + 0.5f) - 0.5f;
    cubic_weights(sx - floorf(sx), wx[p]);
  }
  int cx[5];
#pragma unroll
  for (int b = 0; b < 5; ++b) cx[b] = clampi(i - 2 + b, 0, w - 1);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float* row = xp + (size_t)clampi(iy - 1 + a, 0, h - 1) * w;
    float v[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) v[b] = __ldg(row + cx[b]);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int o = (p < 2) ? 0 : 1;                      // outputs 0,1 start at source i-2, outputs 2,3 at i-1
      float r = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) r = fmaf(wx[p][b], v[o + b], r);
      acc[p] = fmaf(wy[a], r, acc[p]);
    }
  }
  *reinterpret_cast<float4*>(y + ((size_t)blockIdx.z * 4 * h + oy) * 4 * w + 4 * i) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

cudaError_t launch_bicubic(const float* x, float* y, int planes, int h, int w, int num, int den, cudaStream_t s) {
  if (num == 4 && den == 1) {
    dim3 grid((w + 31) / 32, (4 * h + 7) / 8, planes);
    bicubic_x4_kernel<<<grid, 256, 0, s>>>(x, y, h, w);
    return cudaGetLastError();
  }
  int oh = h * num / den, ow = w * num / den;
  dim3 grid((ow + 31) / 32, (oh + 7) / 8, planes);
  bicubic_kernel<<<grid, 256, 0, s>>>(x, y, h, w, oh, ow, num, den);
  return cudaGetLastError();
}

// taps of the two fixed-ratio resizes (exact dyadic constants)
#define LG_DN0 (-0.09375f)
#define LG_DN1 (0.59375f)
#define LG_UP_A (-0.03515625f)
#define LG_UP_B (0.26171875f)
#define LG_UP_C (0.87890625f)
#define LG_UP_D (-0.10546875f)

// ------------------------------------------------------------------------------------------------
// data_down: resid[n,b] = D(Z)[n,b] - ms[n,b]       one CTA = one 16x16 tile of one (n,b) plane at res h x w
// ------------------------------------------------------------------------------------------------
constexpr int DT_ = 16;                    // output tile (res 1/4)
constexpr int DZ = 4 * DT_ + 18;           // 82: Z patch, replicated borders
constexpr int DA1 = 2 * DT_ + 8;           // 40: bicubic-half(Z), zero outside
constexpr int DA1P = 2 * DT_ + 6;          // 38: dw3x3(D.1), replicated
constexpr int DA2 = DT_ + 2;               // 18: bicubic-half, zero outside

__global__ void __launch_bounds__(256) data_down_kernel(const float* __restrict__ z, const float* __restrict__ ms,
                                                         float* __restrict__ resid, const float* __restrict__ w1,
                                                         const float* __restrict__ b1, const float* __restrict__ w3,
                                                         const float* __restrict__ b3, int B, int h, int w) {
  __shared__ float sZ[DZ * DZ];
  __shared__ float sA1[DA1 * DA1];
  __shared__ float sA1p[DA1P * DA1P];
  __shared__ float sA2[DA2 * DA2];
  const int H = 4 * h, W = 4 * w, h2 = 2 * h, w2 = 2 * w;
  const int plane = blockIdx.z, b = plane % B;
  const int oy0 = blockIdx.y * DT_, ox0 = blockIdx.x * DT_;
  const int tid = threadIdx.x;
  const float* zp = z + (size_t)plane * H * W;

  for (int i = tid; i < DZ * DZ; i += 256) {
    int py = i / DZ, px = i - py * DZ;
    int gy = clampi(4 * oy0 - 9 + py, 0, H - 1), gx = clampi(4 * ox0 - 9 + px, 0, W - 1);
    sZ[i] = __ldg(zp + (size_t)gy * W + gx);
  }
  float k1[9], k3[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { k1[i] = __ldg(w1 + b * 9 + i); k3[i] = __ldg(w3 + b * 9 + i); }
  const float bias1 = __ldg(b1 + b), bias3 = __ldg(b3 + b);
  const float dn[4] = {LG_DN0, LG_DN1, LG_DN1, LG_DN0};
  __syncthreads();

  // a1 = bicubic 1/2 of Z at res 2h; abs coord j = 2*o0 - 4 + p ; taps Z[2j-1+k] -> patch index 2p+k
  for (int i = tid; i < DA1 * DA1; i += 256) {
    int py = i / DA1, px = i - py * DA1;
    int jy = 2 * oy0 - 4 + py, jx = 2 * ox0 - 4 + px;
    float acc = 0.f;
    if (jy >= 0 && jy < h2 && jx >= 0 && jx < w2) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float* row = sZ + (2 * py + a) * DZ + 2 * px;
        float r = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) r = fmaf(dn[c], row[c], r);
        acc = fmaf(dn[a], r, acc);
      }
    }
    sA1[i] = acc;
  }
  __syncthreads();
  // a1p = dw3x3(D.1)(a1) + bias, evaluated at the clamped coordinate (replicate for the next resize)
  for (int i = tid; i < DA1P * DA1P; i += 256) {
    int py = i / DA1P, px = i - py * DA1P;
    int cy = clampi(2 * oy0 - 3 + py, 0, h2 - 1) - (2 * oy0 - 4);   // index into sA1 of the centre
    int cx = clampi(2 * ox0 - 3 + px, 0, w2 - 1) - (2 * ox0 - 4);
    float acc = bias1;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc = fmaf(k1[a * 3 + c], sA1[(cy - 1 + a) * DA1 + cx - 1 + c], acc);
    sA1p[i] = acc;
  }
  __syncthreads();
  // a2 = bicubic 1/2 of a1p at res h; abs j = o0 - 1 + p ; taps a1p[2j-1+k] -> patch index 2p+k
  for (int i = tid; i < DA2 * DA2; i += 256) {
    int py = i / DA2, px = i - py * DA2;
    int jy = oy0 - 1 + py, jx = ox0 - 1 + px;
    float acc = 0.f;
    if (jy >= 0 && jy < h && jx >= 0 && jx < w) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float* row = sA1p + (2 * py + a) * DA1P + 2 * px;
        float r = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) r = fmaf(dn[c], row[c], r);
        acc = fmaf(dn[a], r, acc);
      }
    }
    sA2[i] = acc;
  }
  __syncthreads();
  {
    int py = tid >> 4, px = tid & 15;
    int oy = oy0 + py, ox = ox0 + px;
    if (oy < h && ox < w) {
      float acc = bias3;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc = fmaf(k3[a * 3 + c], sA2[(py + a) * DA2 + px + c], acc);
      size_t o = ((size_t)plane * h + oy) * w + ox;
      resid[o] = acc - __ldg(ms + o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// data_up: Z' = Z - eta*(DT(resid) + RT(R(Z) - pan))    one CTA = one 32x32 full-res tile, all B channels
// ------------------------------------------------------------------------------------------------
constexpr int UT = 32;
constexpr int UR = UT / 4 + 6;    // 14: resid patch, replicated
constexpr int UB2 = UT / 2 + 6;   // 22: bicubic x2, zero outside
constexpr int UB2P = UT / 2 + 4;  // 20: dw3x3(DT.1), replicated
constexpr int UB3 = UT + 2;       // 34: bicubic x2, zero outside

// x2 bicubic read: destination abs coord i, source patch s (stride ld) whose element 0 is abs coord s0.
__device__ __forceinline__ void up2_taps(int i, int s0, int& first, float (&t)[4]) {
  int q = i >> 1;                                   // i may be -1 -> q = -1 (arithmetic shift), odd
  if (i & 1) { first = q - 1 - s0; t[0] = LG_UP_D; t[1] = LG_UP_C; t[2] = LG_UP_B; t[3] = LG_UP_A; }
  else       { first = q - 2 - s0; t[0] = LG_UP_A; t[1] = LG_UP_B; t[2] = LG_UP_C; t[3] = LG_UP_D; }
}

template <int B>
__global__ void __launch_bounds__(256) data_up_kernel(const float* __restrict__ z, const float* __restrict__ resid,
                                                       const float* __restrict__ pan, float* __restrict__ zout,
                                                       DataW wts, int stage, int h, int w) {
  __shared__ float sR[UR * UR];
  __shared__ float sB2[UB2 * UB2];
  __shared__ float sB2p[UB2P * UB2P];
  __shared__ float sB3[UB3 * UB3];
  const int H = 4 * h, W = 4 * w, h2 = 2 * h, w2 = 2 * w;
  const int n = blockIdx.z;
  const int Y0 = blockIdx.y * UT, X0 = blockIdx.x * UT;
  const int y1 = Y0 / 2, x1 = X0 / 2, y2 = Y0 / 4, x2 = X0 / 4;
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const float eta = __ldg(wts.eta[stage]);

  // pan term needs R(Z) = sum_b r_w[b] Z[b] + r_b at the thread's 4 pixels
  float zv[B][4], u[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) u[k] = 0.f;
  const bool xin = (X0 + tx) < W;
#pragma unroll
  for (int b = 0; b < B; ++b) {
    const float rw = __ldg(wts.r_w + b);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int Y = Y0 + ty + 8 * k;
      zv[b][k] = (xin && Y < H) ? __ldg(z + (((size_t)n * B + b) * H + Y) * W + X0 + tx) : 0.f;
      u[k] = fmaf(rw, zv[b][k], u[k]);
    }
  }
  {
    const float rb = __ldg(wts.r_b);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int Y = Y0 + ty + 8 * k;
      float p = (xin && Y < H) ? __ldg(pan + ((size_t)n * H + Y) * W + X0 + tx) : 0.f;
      u[k] = (u[k] + rb) - p;
    }
  }

  for (int b = 0; b < B; ++b) {
    const float* rp = resid + ((size_t)n * B + b) * h * w;
    __syncthreads();                                  // previous channel finished reading the patches
    for (int i = tid; i < UR * UR; i += 256) {
      int py = i / UR, px = i - py * UR;
      sR[i] = __ldg(rp + (size_t)clampi(y2 - 3 + py, 0, h - 1) * w + clampi(x2 - 3 + px, 0, w - 1));
    }
    float k1[9], k3[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { k1[i] = __ldg(wts.dt1_w + b * 9 + i); k3[i] = __ldg(wts.dt3_w + b * 9 + i); }
    const float bias1 = __ldg(wts.dt1_b + b), bias3 = __ldg(wts.dt3_b + b);
    __syncthreads();
    // b2 = bicubic x2 of resid at res 2h, abs j = y1 - 3 + p, zero outside
    for (int i = tid; i < UB2 * UB2; i += 256) {
      int py = i / UB2, px = i - py * UB2;
      int jy = y1 - 3 + py, jx = x1 - 3 + px;
      float acc = 0.f;
      if (jy >= 0 && jy < h2 && jx >= 0 && jx < w2) {
        int fy, fx; float wy[4], wx[4];
        up2_taps(jy, y2 - 3, fy, wy);
        up2_taps(jx, x2 - 3, fx, wx);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float* row = sR + (fy + a) * UR + fx;
          float r = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) r = fmaf(wx[c], row[c], r);
          acc = fmaf(wy[a], r, acc);
        }
      }
      sB2[i] = acc;
    }
    __syncthreads();
    // b2p = dw3x3(DT.1)(b2) + bias at the clamped coordinate, abs i = y1 - 2 + p
    for (int i = tid; i < UB2P * UB2P; i += 256) {
      int py = i / UB2P, px = i - py * UB2P;
      int cy = clampi(y1 - 2 + py, 0, h2 - 1) - (y1 - 3);
      int cx = clampi(x1 - 2 + px, 0, w2 - 1) - (x1 - 3);
      float acc = bias1;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc = fmaf(k1[a * 3 + c], sB2[(cy - 1 + a) * UB2 + cx - 1 + c], acc);
      sB2p[i] = acc;
    }
    __syncthreads();
    // b3 = bicubic x2 of b2p at full res, abs i = Y0 - 1 + p, zero outside
    for (int i = tid; i < UB3 * UB3; i += 256) {
      int py = i / UB3, px = i - py * UB3;
      int iy = Y0 - 1 + py, ix = X0 - 1 + px;
      float acc = 0.f;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        int fy, fx; float wy[4], wx[4];
        up2_taps(iy, y1 - 2, fy, wy);
        up2_taps(ix, x1 - 2, fx, wx);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float* row = sB2p + (fy + a) * UB2P + fx;
          float r = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) r = fmaf(wx[c], row[c], r);
          acc = fmaf(wy[a], r, acc);
        }
      }
      sB3[i] = acc;
    }
    __syncthreads();
    const float rtw = __ldg(wts.rt_w + b), rtb = __ldg(wts.rt_b + b);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int py = ty + 8 * k, Y = Y0 + py;
      if (xin && Y < H) {
        float acc = bias3;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int c = 0; c < 3; ++c) acc = fmaf(k3[a * 3 + c], sB3[(py + a) * UB3 + tx + c], acc);
        float pan_term = fmaf(rtw, u[k], rtb);
        zout[(((size_t)n * B + b) * H + Y) * W + X0 + tx] = zv[b][k] - eta * (acc + pan_term);
      }
    }
  }
}

cudaError_t launch_data_step(const DataW& wts, int stage, int B, const float* z_in, const float* ms, const float* pan,
                             float* resid, float* z_out, int N, int h, int w, cudaStream_t s) {
  dim3 gd((w + DT_ - 1) / DT_, (h + DT_ - 1) / DT_, N * B);
  data_down_kernel<<<gd, 256, 0, s>>>(z_in, ms, resid, wts.d1_w, wts.d1_b, wts.d3_w, wts.d3_b, B, h, w);
  dim3 gu((4 * w + UT - 1) / UT, (4 * h + UT - 1) / UT, N);
  if (B == 4) data_up_kernel<4><<<gu, 256, 0, s>>>(z_in, resid, pan, z_out, wts, stage, h, w);
  else if (B == 8) data_up_kernel<8><<<gu, 256, 0, s>>>(z_in, resid, pan, z_out, wts, stage, h, w);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace lg
