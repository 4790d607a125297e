// window_msa.cu — the local branch of LGMixer: 8x8-window multi-head self-attention on the first half of
// the channels (models/common/LGT.py:112-146, window merge :207-208), fused with the pre-norm LayerNorm
// (LGT.py:54-61) and the to_qkv 1x1 conv.
//
//   per window (64 tokens, token = i*8+j, LGT.py:135) and head (2 heads, channel = head*d + c, LGT.py:138):
//     q,k,v = to_qkv(x_win) split in that order along out-channels (LGT.py:136)
//     out   = softmax(q*d^-0.5 . k^T + pos_emb[head]) . v                               (LGT.py:139-143)
//
// The window partition / reverse rearranges of the reference (≈40 copy kernels per stage) do not exist
// here: a window is addressed in place inside the NHWC map.  Logits never leave registers (the reference
// materialises a [N*nWin,2,64,64] tensor in HBM).  One thread owns one (window, head, query) row:
// 64 logits in registers, K/V rows broadcast from shared memory, pos_emb pre-transposed to [head][key][query]
// so the per-query reads are conflict-free.
#include "common.cuh"

namespace lg {

constexpr int kWinPerIter = 2;          // windows processed concurrently by one CTA (128 threads each)
constexpr int kMsaThreads = 128 * kWinPerIter;
constexpr bool kRecomputeLogits = false;   // two-pass softmax that recomputes q.k instead of keeping 64 logits in registers

template <int C2>
struct MsaSmem {
  static constexpr int D = C2 / kHeads;
  alignas(16) float pos_t[kHeads * 64 * 64];        // [h][j/4][i][j%4], pre-scaled by log2(e)
  alignas(16) float wqkv[3 * C2 * C2];              // [3*C2][C2]
  float bqkv[3 * C2];
  float xs[kWinPerIter][C2][64 + 1];                // LN'd local half, channel-major (conflict-free per-token reads)
  float ks[kWinPerIter][kHeads][D][64];                // channel-major: four keys per 128-bit broadcast
  float vs[kWinPerIter][kHeads][64][D];
};

template <int C2, bool PRE_LN>
__global__ void __launch_bounds__(kMsaThreads, (C2 <= 16) ? 2 : 1) window_msa_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                  BlockW w, int H, int W, int total_windows,
                                                                  int windows_per_cta) {
  constexpr int D = C2 / kHeads;
  constexpr int CIN = PRE_LN ? 2 * C2 : C2;         // channels per pixel of the input map
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MsaSmem<C2>& sm = *reinterpret_cast<MsaSmem<C2>*>(smem_raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < kHeads * 64 * 64; i += kMsaThreads) sm.pos_t[i] = __ldg(w.pos_t + i);
  for (int i = tid; i < 3 * C2 * C2; i += kMsaThreads) sm.wqkv[i] = __ldg(w.qkv_w + i);
  for (int i = tid; i < 3 * C2; i += kMsaThreads) sm.bqkv[i] = __ldg(w.qkv_b + i);

  const int nwx = W / kWin, nwy = H / kWin;
  const int slot = tid >> 7;                        // which of the concurrent windows
  const int lt = tid & 127;
  const int head = lt >> 6, tok = lt & 63;
  // head_channel ** -0.5 rounded to fp32 like the reference's python-float * tensor (LGT.py:119,139)
  const float scale = (D == 4) ? 0.5f : (D == 8) ? 0.35355339059327379f : (D == 16) ? 0.25f : 0.17677669529663689f;

  const int w_begin = blockIdx.x * windows_per_cta;
  const int w_end = min(w_begin + windows_per_cta, total_windows);
  for (int wbase = w_begin; wbase < w_end; wbase += kWinPerIter) {
    const int widx = wbase + slot;
    const bool active = widx < w_end;
    __syncthreads();                                // previous iteration's K/V/xs fully consumed; weights loaded
    int n = 0, wy = 0, wx = 0;
    if (active) {
      wx = widx % nwx;
      int q = widx / nwx;
      wy = q % nwy;
      n = q / nwy;
    }
    // 1) load (+ LayerNorm) : threads 0..63 of each slot own one token each
    if (active && lt < 64) {
      const int py = wy * kWin + (lt >> 3), px = wx * kWin + (lt & 7);
      const float* src = x + (((size_t)n * H + py) * W + px) * CIN;
      if constexpr (PRE_LN) {
        float v[CIN];
        load_vec<CIN>(v, src);
        // LayerNorm over all c channels, but only the local half is needed afterwards
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
          sm.xs[slot][i][lt] = (v[i] - mean) * rstd * __ldg(w.ln1_w + i) + __ldg(w.ln1_b + i);
      } else {
        float v[C2];
        load_vec<C2>(v, src);
#pragma unroll
        for (int i = 0; i < C2; ++i) sm.xs[slot][i][lt] = v[i];
      }
    }
    __syncthreads();
    // 2) q/k/v of this thread's (head, token)
    float q[D];
    if (active) {
      float kk[D], vv[D];
      float2 xv[C2 / 2];
#pragma unroll
      for (int k = 0; k < C2 / 2; ++k) xv[k] = make_float2(sm.xs[slot][2 * k][tok], sm.xs[slot][2 * k + 1][tok]);
      // one output channel = one weight row [C2] read as 128-bit broadcasts, dot product on the packed fp32 pipe
      auto dot = [&](int o) {
        const float4* wr = reinterpret_cast<const float4*>(&sm.wqkv[o * C2]);
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int k4 = 0; k4 < C2 / 4; ++k4) {
          const float4 w4 = wr[k4];
          acc = __ffma2_rn(make_float2(w4.x, w4.y), xv[2 * k4], acc);
          acc = __ffma2_rn(make_float2(w4.z, w4.w), xv[2 * k4 + 1], acc);
        }
        return (acc.x + acc.y) + sm.bqkv[o];
      };
#pragma unroll
      for (int j = 0; j < D; ++j) {
        q[j] = dot(head * D + j);
        kk[j] = dot(C2 + head * D + j);
        vv[j] = dot(2 * C2 + head * D + j);
      }
#pragma unroll
      for (int j = 0; j < D; ++j) {
        q[j] *= scale * 1.4426950408889634f;              // logits in base-2 units: softmax via ex2.approx
        sm.ks[slot][head][j][tok] = kk[j];
        sm.vs[slot][head][tok][j] = vv[j];
      }
    }
    __syncthreads();
    // 3) logits + softmax + PV, all in registers
    if (active) {
      // logits for 4 keys at a time on the packed fp32 pipe; pos_t already holds pos_emb * log2(e), transposed
      auto logits4 = [&](int j, float2& a01, float2& a23) {
        const float4 p4 = *reinterpret_cast<const float4*>(&sm.pos_t[((head * 16 + (j >> 2)) * 64 + tok) * 4]);
        a01 = make_float2(p4.x, p4.y);
        a23 = make_float2(p4.z, p4.w);
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const float4 kv = *reinterpret_cast<const float4*>(&sm.ks[slot][head][c][j]);
          const float2 qc = make_float2(q[c], q[c]);
          a01 = __ffma2_rn(qc, make_float2(kv.x, kv.y), a01);
          a23 = __ffma2_rn(qc, make_float2(kv.z, kv.w), a23);
        }
      };
      float2 sum2 = make_float2(0.f, 0.f);
      float2 o2[D / 2];
#pragma unroll
      for (int c = 0; c < D / 2; ++c) o2[c] = make_float2(0.f, 0.f);
      auto accumulate2 = [&](int j, float2 s, float2 nmx) {      // keys j, j+1
        const float2 d = __fadd2_rn(s, nmx);
        float2 p;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p.x) : "f"(d.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p.y) : "f"(d.y));
        sum2 = __fadd2_rn(sum2, p);
#pragma unroll
        for (int c4 = 0; c4 < D; c4 += 4) {
          const float4 v0 = *reinterpret_cast<const float4*>(&sm.vs[slot][head][j][c4]);
          const float4 v1 = *reinterpret_cast<const float4*>(&sm.vs[slot][head][j + 1][c4]);
          const float2 p0 = make_float2(p.x, p.x), p1 = make_float2(p.y, p.y);
          o2[c4 / 2] = __ffma2_rn(p0, make_float2(v0.x, v0.y), o2[c4 / 2]);
          o2[c4 / 2 + 1] = __ffma2_rn(p0, make_float2(v0.z, v0.w), o2[c4 / 2 + 1]);
          o2[c4 / 2] = __ffma2_rn(p1, make_float2(v1.x, v1.y), o2[c4 / 2]);
          o2[c4 / 2 + 1] = __ffma2_rn(p1, make_float2(v1.z, v1.w), o2[c4 / 2 + 1]);
        }
      };
      float mx = -INFINITY;
      if constexpr (kRecomputeLogits && D == 4) {
        // measured slower (175 vs 135 us / 16 pairs): the doubled K / pos_emb shared-memory traffic outweighs the
        // occupancy gained from ~64 registers per thread, so this variant is compiled out
#pragma unroll 4
        for (int j = 0; j < 64; j += 4) {
          float2 a01, a23;
          logits4(j, a01, a23);
          mx = fmaxf(mx, fmaxf(fmaxf(a01.x, a01.y), fmaxf(a23.x, a23.y)));
        }
        const float2 nmx = make_float2(-mx, -mx);
#pragma unroll 4
        for (int j = 0; j < 64; j += 4) {
          float2 a01, a23;
          logits4(j, a01, a23);
          accumulate2(j, a01, nmx);
          accumulate2(j + 2, a23, nmx);
        }
      } else {
        float2 s2[32];
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          logits4(j, s2[j / 2], s2[j / 2 + 1]);
          mx = fmaxf(mx, fmaxf(fmaxf(s2[j / 2].x, s2[j / 2].y), fmaxf(s2[j / 2 + 1].x, s2[j / 2 + 1].y)));
        }
        const float2 nmx = make_float2(-mx, -mx);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) accumulate2(2 * jj, s2[jj], nmx);
      }
      const float sum = sum2.x + sum2.y;
      float o[D];
#pragma unroll
      for (int c = 0; c < D / 2; ++c) { o[2 * c] = o2[c].x; o[2 * c + 1] = o2[c].y; }
      const float inv = 1.0f / sum;
      const int py = wy * kWin + (tok >> 3), px = wx * kWin + (tok & 7);
      float* dst = y + (((size_t)n * H + py) * W + px) * C2 + head * D;
#pragma unroll
      for (int c4 = 0; c4 < D; c4 += 4)
        *reinterpret_cast<float4*>(dst + c4) =
            make_float4(o[c4] * inv, o[c4 + 1] * inv, o[c4 + 2] * inv, o[c4 + 3] * inv);
    }
  }
}

template <int C2>
static cudaError_t launch_msa_t(const BlockW& w, const float* x, float* y, int pre_ln, int N, int H, int W,
                                cudaStream_t s) {
  const int total = N * (H / kWin) * (W / kWin);
  int per_cta = 16;                                  // amortise the 32 KB pos_emb + weight staging
  while (per_cta > kWinPerIter && (total + per_cta - 1) / per_cta < 2 * 148) per_cta /= 2;
  const int grid = (total + per_cta - 1) / per_cta;
  const size_t smem = sizeof(MsaSmem<C2>);
  cudaError_t e;
  if (pre_ln) {
    e = cudaFuncSetAttribute(window_msa_kernel<C2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    window_msa_kernel<C2, true><<<grid, kMsaThreads, smem, s>>>(x, y, w, H, W, total, per_cta);
  } else {
    e = cudaFuncSetAttribute(window_msa_kernel<C2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    window_msa_kernel<C2, false><<<grid, kMsaThreads, smem, s>>>(x, y, w, H, W, total, per_cta);
  }
  return cudaGetLastError();
}

cudaError_t launch_window_msa(const BlockW& w, int c, const float* x, float* y_half, int pre_ln, int N, int H, int W,
                              cudaStream_t s) {
  switch (c) {
    case 16: return launch_msa_t<8>(w, x, y_half, pre_ln, N, H, W, s);
    case 32: return launch_msa_t<16>(w, x, y_half, pre_ln, N, H, W, s);
    case 64: return launch_msa_t<32>(w, x, y_half, pre_ln, N, H, W, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace lg
