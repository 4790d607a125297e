// ffn.cu — residual(pre_norm(feed_forward)) of one LGB block (models/common/LGT.py:91-109, 45-61;
// depthwise_conv: basic_module_unformer_v2.py:37-53):
//     y = x + W2 . GELU( dw3x3( W1 . GELU( W0 . LN(x) + b0 ) + b1 ) + bdw ) + b2
// 74 % of the network's FLOPs.  This file holds the fp32 CUDA-core path: two launches per block,
//   ffn_expand   : LN -> 1x1 c->4c -> GELU -> 1x1 4c->4c (+bias)           -> hidden [N,H,W,4c]
//   ffn_contract : depthwise 3x3 (zero pad, +bias) -> GELU -> 1x1 4c->c (+bias) -> + x
// Weights are pre-transposed ([k][out]) at load time and staged in shared memory so that a warp reads them as 128-bit
// broadcasts; activations are staged channel-major so per-pixel reads are conflict-free.
#include "common.cuh"

namespace lg {

size_t ffn_hidden_floats(int N, int H, int W, int c) { return (size_t)N * H * W * 4 * c; }

// ---- expand -------------------------------------------------------------------------------------------
constexpr int EP = 64;                  // pixels per CTA
constexpr int EKC = 32;                 // k-rows of W1 staged per chunk

template <int C>
__global__ void __launch_bounds__(256) ffn_expand_kernel(const float* __restrict__ x, float* __restrict__ hidden, BlockW w,
                                                          long long total) {
  constexpr int C4 = 4 * C;
  extern __shared__ __align__(16) float sm_e[];
  float* xs = sm_e;                     // [C][EP]
  float* h1 = xs + C * EP;              // [C4][EP]
  float* wt = h1 + C4 * EP;             // W0^T [C][C4]  then chunks of W1^T [EKC][C4]
  const int tid = threadIdx.x;
  const int px = tid & (EP - 1), og = tid >> 6;          // 4 output groups of C outputs each
  const long long p = (long long)blockIdx.x * EP + px;
  const bool live = p < total;

  for (int i = tid; i < C * C4 / 4; i += 256)            // W0^T [C][C4] (pre-transposed at load time)
    reinterpret_cast<float4*>(wt)[i] = __ldg(reinterpret_cast<const float4*>(w.f0_wt) + i);
  if (og == 0) {
    float v[C];
    if (live) {
      load_vec<C>(v, x + p * C);
      layer_norm_inplace<C>(v, w.ln2_w, w.ln2_b);
    } else {
#pragma unroll
      for (int i = 0; i < C; ++i) v[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < C; ++i) xs[i * EP + px] = v[i];
  }
  __syncthreads();
  {
    float acc[C];
#pragma unroll
    for (int o = 0; o < C; ++o) acc[o] = __ldg(w.f0_b + og * C + o);
#pragma unroll 4
    for (int k = 0; k < C; ++k) {
      const float xv = xs[k * EP + px];
      const float4* wr = reinterpret_cast<const float4*>(wt + k * C4 + og * C);
#pragma unroll
      for (int o4 = 0; o4 < C / 4; ++o4) {
        float4 t = wr[o4];
        acc[4 * o4] = fmaf(t.x, xv, acc[4 * o4]); acc[4 * o4 + 1] = fmaf(t.y, xv, acc[4 * o4 + 1]);
        acc[4 * o4 + 2] = fmaf(t.z, xv, acc[4 * o4 + 2]); acc[4 * o4 + 3] = fmaf(t.w, xv, acc[4 * o4 + 3]);
      }
    }
#pragma unroll
    for (int o = 0; o < C; ++o) h1[(og * C + o) * EP + px] = gelu_fast(acc[o]);
  }
  float acc[C];
#pragma unroll
  for (int o = 0; o < C; ++o) acc[o] = __ldg(w.f1_b + og * C + o);
  for (int k0 = 0; k0 < C4; k0 += EKC) {
    __syncthreads();                                      // h1 complete / previous chunk consumed
    for (int i = tid; i < EKC * C4 / 4; i += 256)         // rows k0..k0+EKC of W1^T [C4][C4]
      reinterpret_cast<float4*>(wt)[i] = __ldg(reinterpret_cast<const float4*>(w.f1_wt + (size_t)k0 * C4) + i);
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < EKC; ++kk) {
      const float hv = h1[(k0 + kk) * EP + px];
      const float4* wr = reinterpret_cast<const float4*>(wt + kk * C4 + og * C);
#pragma unroll
      for (int o4 = 0; o4 < C / 4; ++o4) {
        float4 t = wr[o4];
        acc[4 * o4] = fmaf(t.x, hv, acc[4 * o4]); acc[4 * o4 + 1] = fmaf(t.y, hv, acc[4 * o4 + 1]);
        acc[4 * o4 + 2] = fmaf(t.z, hv, acc[4 * o4 + 2]); acc[4 * o4 + 3] = fmaf(t.w, hv, acc[4 * o4 + 3]);
      }
    }
  }
  if (live) store_vec<C>(hidden + p * C4 + og * C, acc);
}

// ---- contract -----------------------------------------------------------------------------------------
constexpr int CTH = 8, CTW = 16;                 // pixel tile (rows x cols) = 128 pixels
constexpr int CHH = CTH + 2, CHW = CTW + 2;      // halo tile
constexpr int CKC = 32;                          // hidden channels per chunk
constexpr int CPAD = CKC + 4;                    // padded pixel stride (conflict-free float4 reads)

template <int C>
__global__ void __launch_bounds__(256) ffn_contract_kernel(const float* __restrict__ hidden, const float* __restrict__ x,
                                                            float* __restrict__ y, BlockW w, int H, int W) {
  constexpr int C4 = 4 * C;
  constexpr int CO = C / 2;                      // outputs per thread (two halves)
  extern __shared__ __align__(16) float sm_c[];
  float* w2t = sm_c;                             // W2^T [C4][C]
  float* hs = w2t + C4 * C;                      // [CHH*CHW][CPAD]
  float* act = hs + CHH * CHW * CPAD;            // [CKC][128]
  float* dws = act + CKC * 128;                  // [9][CKC] + [CKC] bias
  const int tid = threadIdx.x;
  const int px = tid & 127, half = tid >> 7;
  const int ty = px >> 4, tx = px & 15;
  const int tiles_x = (W + CTW - 1) / CTW;
  const int tile = blockIdx.x;
  const int n = blockIdx.y;
  const int Y0 = (tile / tiles_x) * CTH, X0 = (tile % tiles_x) * CTW;
  const int Y = Y0 + ty, X = X0 + tx;
  const bool live = (Y < H) && (X < W);

  for (int i = tid; i < C4 * C / 4; i += 256)    // W2^T [C4][C] (pre-transposed at load time)
    reinterpret_cast<float4*>(w2t)[i] = __ldg(reinterpret_cast<const float4*>(w.f2_wt) + i);
  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) acc[o] = __ldg(w.f2_b + half * CO + o);

  const int lane = tid & 31, warp = tid >> 5;
  for (int k0 = 0; k0 < C4; k0 += CKC) {
    __syncthreads();                             // previous chunk consumed (and w2t staged)
    for (int hp = warp; hp < CHH * CHW; hp += 8) {          // one warp loads one halo pixel's 32 channels
      int hy = hp / CHW, hx = hp - hy * CHW;
      int gy = Y0 - 1 + hy, gx = X0 - 1 + hx;
      float v = 0.f;                                          // zero padding of the dw conv (bmu:17-18)
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = __ldg(hidden + (((size_t)n * H + gy) * W + gx) * C4 + k0 + lane);
      hs[hp * CPAD + lane] = v;
    }
    for (int i = tid; i < 9 * CKC; i += 256) {                // dw weights [C4][9] -> [9][CKC]
      int ch = i / 9, t = i - ch * 9;
      dws[t * CKC + ch] = __ldg(w.dw_w + (size_t)(k0 + ch) * 9 + t);
    }
    if (tid < CKC) dws[9 * CKC + tid] = __ldg(w.dw_b + k0 + tid);
    __syncthreads();
    {
      // depthwise 3x3 + bias + GELU for 16 channels (half) of this thread's pixel
      float g[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) g[c] = dws[9 * CKC + half * 16 + c];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const float4* hv = reinterpret_cast<const float4*>(hs + ((ty + a) * CHW + tx + b) * CPAD + half * 16);
          const float4* wv = reinterpret_cast<const float4*>(dws + (a * 3 + b) * CKC + half * 16);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            float4 hh = hv[c4], ww = wv[c4];
            g[4 * c4] = fmaf(ww.x, hh.x, g[4 * c4]); g[4 * c4 + 1] = fmaf(ww.y, hh.y, g[4 * c4 + 1]);
            g[4 * c4 + 2] = fmaf(ww.z, hh.z, g[4 * c4 + 2]); g[4 * c4 + 3] = fmaf(ww.w, hh.w, g[4 * c4 + 3]);
          }
        }
#pragma unroll
      for (int c = 0; c < 16; ++c) act[(half * 16 + c) * 128 + px] = gelu_fast(g[c]);
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < CKC; ++kk) {
      const float av = act[kk * 128 + px];
      const float4* wr = reinterpret_cast<const float4*>(w2t + (k0 + kk) * C + half * CO);
#pragma unroll
      for (int o4 = 0; o4 < CO / 4; ++o4) {
        float4 t = wr[o4];
        acc[4 * o4] = fmaf(t.x, av, acc[4 * o4]); acc[4 * o4 + 1] = fmaf(t.y, av, acc[4 * o4 + 1]);
        acc[4 * o4 + 2] = fmaf(t.z, av, acc[4 * o4 + 2]); acc[4 * o4 + 3] = fmaf(t.w, av, acc[4 * o4 + 3]);
      }
    }
  }
  if (live) {
    const size_t o = (((size_t)n * H + Y) * W + X) * C + half * CO;
#pragma unroll
    for (int o4 = 0; o4 < CO / 4; ++o4) {
      float4 r = *reinterpret_cast<const float4*>(x + o + 4 * o4);
      *reinterpret_cast<float4*>(y + o + 4 * o4) =
          make_float4(acc[4 * o4] + r.x, acc[4 * o4 + 1] + r.y, acc[4 * o4 + 2] + r.z, acc[4 * o4 + 3] + r.w);
    }
  }
}

template <int C>
static cudaError_t ffn_t(const BlockW& w, const float* x, float* hidden, float* y, int N, int H, int W, cudaStream_t s) {
  constexpr int C4 = 4 * C;
  const long long total = (long long)N * H * W;
  size_t smem_e = (size_t)(C * EP + C4 * EP + (C > EKC ? C : EKC) * C4) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ffn_expand_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e);
  if (e != cudaSuccess) return e;
  ffn_expand_kernel<C><<<(unsigned)((total + EP - 1) / EP), 256, smem_e, s>>>(x, hidden, w, total);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  size_t smem_c = (size_t)(C4 * C + CHH * CHW * CPAD + CKC * 128 + 10 * CKC) * sizeof(float);
  e = cudaFuncSetAttribute(ffn_contract_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
  if (e != cudaSuccess) return e;
  dim3 grid(((W + CTW - 1) / CTW) * ((H + CTH - 1) / CTH), N);
  ffn_contract_kernel<C><<<grid, 256, smem_c, s>>>(hidden, x, y, w, H, W);
  return cudaGetLastError();
}

cudaError_t launch_ffn(const BlockW& w, int c, const float* x, float* hidden, float* y, int N, int H, int W,
                       cudaStream_t s) {
  switch (c) {
    case 16: return ffn_t<16>(w, x, hidden, y, N, H, W, s);
    case 32: return ffn_t<32>(w, x, hidden, y, N, H, W, s);
    case 64: return ffn_t<64>(w, x, hidden, y, N, H, W, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace lg
