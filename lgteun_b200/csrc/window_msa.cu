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
// materialises a [N*nWin,2,64,64] tensor in HBM).
//
// Mapping: one warp = one (window, head); a lane owns TWO query rows (tokens lane and lane+32) so every K / V row
// fetched from shared memory (128-bit broadcasts) feeds two queries — the kernel is bound by the shared-memory
// instruction queue, not by math.  Softmax is computed online over blocks of 16 (d=4) or 8 keys in base 2
// (pos_emb and the q scale are pre-multiplied by log2 e; ex2.approx), all dot products on the packed fp32 pipe.
#include <stdlib.h>
#include "common.cuh"

namespace lg {

constexpr int kWinPerIter = 4;          // windows processed concurrently by one CTA (2 warps each)
constexpr int kMsaThreads = 64 * kWinPerIter;

template <int C2>
struct MsaSmem {
  static constexpr int D = C2 / kHeads;
  alignas(16) float pos_t[kHeads * 64 * 64];        // [h][j/4][i][j%4], pre-scaled by log2(e)
  alignas(16) float wqkv[3 * C2 * C2];              // [3*C2][C2]
  float bqkv[3 * C2];
  float lnw[C2], lnb[C2];                           // LayerNorm affine of the local half
  float xs[kWinPerIter][C2][64 + 1];                // LN'd local half, channel-major (conflict-free per-token reads)
  alignas(16) float ks[kWinPerIter][kHeads][D][64]; // channel-major: four keys per 128-bit broadcast
  alignas(16) float vs[kWinPerIter][kHeads][64][D];
  alignas(16) float qs[kWinPerIter][kHeads][64][D]; // scaled queries (staged so the projection loop need not be unrolled)
};

template <int C2, bool PRE_LN>
__global__ void __launch_bounds__(kMsaThreads, (C2 <= 8) ? 3 : (C2 <= 16) ? 2 : 1)
window_msa_kernel(const float* __restrict__ x, float* __restrict__ y, BlockW w, int H, int W, int total_windows,
                  int windows_per_cta) {
  constexpr int D = C2 / kHeads;
  constexpr int CIN = PRE_LN ? 2 * C2 : C2;         // channels per pixel of the input map
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MsaSmem<C2>& sm = *reinterpret_cast<MsaSmem<C2>*>(smem_raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < kHeads * 64 * 64; i += kMsaThreads) sm.pos_t[i] = __ldg(w.pos_t + i);
  for (int i = tid; i < 3 * C2 * C2; i += kMsaThreads) sm.wqkv[i] = __ldg(w.qkv_w + i);
  for (int i = tid; i < 3 * C2; i += kMsaThreads) sm.bqkv[i] = __ldg(w.qkv_b + i);
  if (PRE_LN)
    for (int i = tid; i < C2; i += kMsaThreads) { sm.lnw[i] = __ldg(w.ln1_w + i); sm.lnb[i] = __ldg(w.ln1_b + i); }

  // window coordinates: shifts when the window grid is a power of two in both directions (every forward shape), integer
  // divisions otherwise (the per-operator entry point accepts any multiple of 8)
  const int nwx = W / kWin, nwy = H / kWin;
  const bool grid_pow2 = ((nwx & (nwx - 1)) | (nwy & (nwy - 1))) == 0;
  const int lg_nwx = 31 - __clz(nwx), lg_nwy = 31 - __clz(nwy);
  auto window_of = [&](int widx, int& wx, int& wy, int& n) {
    if (grid_pow2) {
      wx = widx & (nwx - 1);
      const int t = widx >> lg_nwx;
      wy = t & (nwy - 1);
      n = t >> lg_nwy;
    } else {
      wx = widx % nwx;
      const int t = widx / nwx;
      wy = t % nwy;
      n = t / nwy;
    }
  };
  const int lane = tid & 31, warp = tid >> 5;
  const int slot = warp >> 1, head = warp & 1;      // attention phase: warp = (window slot, head)
  const int lslot = tid >> 6, ltok = tid & 63;      // load phase: thread = (window slot, token)
  // head_channel ** -0.5 rounded to fp32 like the reference's python-float * tensor (LGT.py:119,139), times log2(e)
  const float scale = ((D == 4) ? 0.5f : (D == 8) ? 0.35355339059327379f : (D == 16) ? 0.25f : 0.17677669529663689f) *
                      1.4426950408889634f;

  const int w_begin = blockIdx.x * windows_per_cta;
  const int w_end = min(w_begin + windows_per_cta, total_windows);
  __syncthreads();                                  // weights staged
  // A window slot is loaded, projected and attended by the SAME two warps (threads 64*slot .. 64*slot+63), so the
  // slots run free of each other: their hand-offs use a 64-thread named barrier, not a CTA-wide one.
  auto slot_sync = [&]() {
    switch (lslot) {                                // immediate barrier ids: the kernel reserves 5 barriers, not all 16
      case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
      case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
      case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
      default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    }
  };
  for (int wbase = w_begin; wbase < w_end; wbase += kWinPerIter) {
    slot_sync();                                    // previous iteration's xs of this slot fully consumed
    // 1) load (+ LayerNorm): one thread per (window, token)
    {
      const int widx = wbase + lslot;
      if (widx < w_end) {
        int wx, wy, n;
        window_of(widx, wx, wy, n);
        const int py = wy * kWin + (ltok >> 3), px = wx * kWin + (ltok & 7);
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
            sm.xs[lslot][i][ltok] = (v[i] - mean) * rstd * sm.lnw[i] + sm.lnb[i];
        } else {
          float v[C2];
          load_vec<C2>(v, src);
#pragma unroll
          for (int i = 0; i < C2; ++i) sm.xs[lslot][i][ltok] = v[i];
        }
      }
    }
    slot_sync();
    const int widx = wbase + slot;
    const bool active = widx < w_end;               // warp-uniform
    // 2) q/k/v of this lane's two tokens for this warp's head: every weight row (128-bit broadcasts) feeds both tokens,
    //    dot products on the packed fp32 pipe with independent accumulators
    if (active) {
      float2 xv[2][C2 / 2];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < C2 / 2; ++k)
          xv[r][k] = make_float2(sm.xs[slot][2 * k][lane + 32 * r], sm.xs[slot][2 * k + 1][lane + 32 * r]);
      auto dot2 = [&](int o, float& d0, float& d1) {
        const float4* wr = reinterpret_cast<const float4*>(&sm.wqkv[o * C2]);
        float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
        for (int k4 = 0; k4 < C2 / 4; ++k4) {
          const float4 w4 = wr[k4];
          const float2 wa = make_float2(w4.x, w4.y), wb = make_float2(w4.z, w4.w);
          a0 = __ffma2_rn(wa, xv[0][2 * k4], a0);
          a1 = __ffma2_rn(wa, xv[1][2 * k4], a1);
          a0 = __ffma2_rn(wb, xv[0][2 * k4 + 1], a0);
          a1 = __ffma2_rn(wb, xv[1][2 * k4 + 1], a1);
        }
        const float bias = sm.bqkv[o];
        d0 = (a0.x + a0.y) + bias;
        d1 = (a1.x + a1.y) + bias;
      };
#pragma unroll 4
      for (int j = 0; j < D; ++j) {
        float q0, q1, k0, k1, v0, v1;
        dot2(head * D + j, q0, q1);
        dot2(C2 + head * D + j, k0, k1);
        dot2(2 * C2 + head * D + j, v0, v1);
        sm.qs[slot][head][lane][j] = q0 * scale;
        sm.qs[slot][head][lane + 32][j] = q1 * scale;
        sm.ks[slot][head][j][lane] = k0;
        sm.ks[slot][head][j][lane + 32] = k1;
        sm.vs[slot][head][lane][j] = v0;
        sm.vs[slot][head][lane + 32][j] = v1;
      }
    }
    __syncwarp();                                   // K / V of a (window, head) are produced and consumed by one warp
    // 3) online softmax over blocks of KB keys, two queries per lane
    if (active) {
      float q[2][D];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c4 = 0; c4 < D; c4 += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&sm.qs[slot][head][lane + 32 * r][c4]);
          q[r][c4] = t.x; q[r][c4 + 1] = t.y; q[r][c4 + 2] = t.z; q[r][c4 + 3] = t.w;
        }
      float m[2] = {-INFINITY, -INFINITY};
      float2 sum2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      float2 o2[2][D / 2];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < D / 2; ++c) o2[r][c] = make_float2(0.f, 0.f);
      constexpr int KB = 8;          // keys per online-softmax block (register budget: 2*KB logits live)
#pragma unroll 1
      for (int jb = 0; jb < 64; jb += KB) {
        float2 s[2][KB / 2];
        float bm[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j4 = 0; j4 < KB / 4; ++j4) {
          const int j = jb + 4 * j4;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const float4 p4 =
                *reinterpret_cast<const float4*>(&sm.pos_t[((head * 16 + (j >> 2)) * 64 + lane + 32 * r) * 4]);
            s[r][2 * j4] = make_float2(p4.x, p4.y);
            s[r][2 * j4 + 1] = make_float2(p4.z, p4.w);
          }
#pragma unroll
          for (int c = 0; c < D; ++c) {
            const float4 kv = *reinterpret_cast<const float4*>(&sm.ks[slot][head][c][j]);
            const float2 k01 = make_float2(kv.x, kv.y), k23 = make_float2(kv.z, kv.w);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const float2 qc = make_float2(q[r][c], q[r][c]);
              s[r][2 * j4] = __ffma2_rn(qc, k01, s[r][2 * j4]);
              s[r][2 * j4 + 1] = __ffma2_rn(qc, k23, s[r][2 * j4 + 1]);
            }
          }
#pragma unroll
          for (int r = 0; r < 2; ++r)
            bm[r] = fmaxf(bm[r], fmaxf(fmaxf(s[r][2 * j4].x, s[r][2 * j4].y), fmaxf(s[r][2 * j4 + 1].x, s[r][2 * j4 + 1].y)));
        }
        float2 nm[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {                 // rescale the running sums to the new maximum
          const float mn = fmaxf(m[r], bm[r]);
          float corr;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(corr) : "f"(m[r] - mn));     // 0 on the first block (m = -inf)
          m[r] = mn;
          nm[r] = make_float2(-mn, -mn);
          const float2 c2 = make_float2(corr, corr);
          sum2[r] = __fmul2_rn(sum2[r], c2);
#pragma unroll
          for (int c = 0; c < D / 2; ++c) o2[r][c] = __fmul2_rn(o2[r][c], c2);
        }
#pragma unroll
        for (int jj = 0; jj < KB / 2; ++jj) {         // keys jb + 2jj, jb + 2jj + 1
          float2 p[2];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const float2 d = __fadd2_rn(s[r][jj], nm[r]);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[r].x) : "f"(d.x));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[r].y) : "f"(d.y));
            sum2[r] = __fadd2_rn(sum2[r], p[r]);
          }
#pragma unroll
          for (int c4 = 0; c4 < D; c4 += 4) {
            const float4 v0 = *reinterpret_cast<const float4*>(&sm.vs[slot][head][jb + 2 * jj][c4]);
            const float4 v1 = *reinterpret_cast<const float4*>(&sm.vs[slot][head][jb + 2 * jj + 1][c4]);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const float2 p0 = make_float2(p[r].x, p[r].x), p1 = make_float2(p[r].y, p[r].y);
              o2[r][c4 / 2] = __ffma2_rn(p0, make_float2(v0.x, v0.y), o2[r][c4 / 2]);
              o2[r][c4 / 2 + 1] = __ffma2_rn(p0, make_float2(v0.z, v0.w), o2[r][c4 / 2 + 1]);
              o2[r][c4 / 2] = __ffma2_rn(p1, make_float2(v1.x, v1.y), o2[r][c4 / 2]);
              o2[r][c4 / 2 + 1] = __ffma2_rn(p1, make_float2(v1.z, v1.w), o2[r][c4 / 2 + 1]);
            }
          }
        }
      }
      int wx, wy, n;
      window_of(widx, wx, wy, n);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int tok = lane + 32 * r;
        const float inv = 1.0f / (sum2[r].x + sum2[r].y);
        const int py = wy * kWin + (tok >> 3), px = wx * kWin + (tok & 7);
        float* dst = y + (((size_t)n * H + py) * W + px) * C2 + head * D;
#pragma unroll
        for (int c4 = 0; c4 < D; c4 += 4)
          *reinterpret_cast<float4*>(dst + c4) = make_float4(o2[r][c4 / 2].x * inv, o2[r][c4 / 2].y * inv,
                                                             o2[r][c4 / 2 + 1].x * inv, o2[r][c4 / 2 + 1].y * inv);
      }
    }
  }
}

// ---- four queries per lane (LGTEUN_MSA=q4; head dim 4 / 8) ------------------------------------------------------------------------
// One warp = one window, lanes 0-15 / 16-31 = the two heads, a lane owns the query rows li, li + 16, li + 32, li + 48 of its
// head: a K / V row fetched from shared memory (a 16-byte load whose two half-warps read the two heads' rows) now feeds four
// queries of each head, i.e. 40 % fewer L1 wavefronts per window than the two-queries form, whose L1 data pipe is 78 % busy.
// Price: a window's tiles belong to one warp instead of two (16 resident warps instead of 24) and ~110 registers.
constexpr int kQ4Slots = 8;                         // windows per CTA iteration = warps
constexpr int kQ4Threads = 32 * kQ4Slots;
template <int C2>
struct MsaQ4Smem {
  static constexpr int D = C2 / kHeads;
  alignas(16) float pos_t[kHeads * 64 * 64];        // [h][j/4][i][j%4], pre-scaled by log2(e)
  alignas(16) float wqkv[3 * C2 * C2];
  float bqkv[3 * C2];
  float lnw[C2], lnb[C2];
  float xs[kQ4Slots][C2][64 + 1];
  alignas(16) float ks[kQ4Slots][kHeads][D][64];
  alignas(16) float vs[kQ4Slots][kHeads][64][D];
};
template <int C2, bool PRE_LN>
__global__ void __launch_bounds__(kQ4Threads, 2)
window_msa_q4_kernel(const float* __restrict__ x, float* __restrict__ y, BlockW w, int H, int W, int total_windows,
                     int windows_per_cta) {
  constexpr int D = C2 / kHeads;
  constexpr int CIN = PRE_LN ? 2 * C2 : C2;
  constexpr int R = 4;                              // queries per lane
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MsaQ4Smem<C2>& sm = *reinterpret_cast<MsaQ4Smem<C2>*>(smem_raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < kHeads * 64 * 64; i += kQ4Threads) sm.pos_t[i] = __ldg(w.pos_t + i);
  for (int i = tid; i < 3 * C2 * C2; i += kQ4Threads) sm.wqkv[i] = __ldg(w.qkv_w + i);
  for (int i = tid; i < 3 * C2; i += kQ4Threads) sm.bqkv[i] = __ldg(w.qkv_b + i);
  if (PRE_LN)
    for (int i = tid; i < C2; i += kQ4Threads) { sm.lnw[i] = __ldg(w.ln1_w + i); sm.lnb[i] = __ldg(w.ln1_b + i); }
  const int nwx = W / kWin, nwy = H / kWin;
  auto window_of = [&](int widx, int& wx, int& wy, int& n) {
    wx = widx % nwx;
    const int t = widx / nwx;
    wy = t % nwy;
    n = t / nwy;
  };
  const int lane = tid & 31, slot = tid >> 5;
  const int head = lane >> 4, li = lane & 15;
  const float scale = ((D == 4) ? 0.5f : (D == 8) ? 0.35355339059327379f : (D == 16) ? 0.25f : 0.17677669529663689f) *
                      1.4426950408889634f;
  const int w_begin = blockIdx.x * windows_per_cta;
  const int w_end = min(w_begin + windows_per_cta, total_windows);
  __syncthreads();
  for (int wbase = w_begin; wbase < w_end; wbase += kQ4Slots) {
    const int widx = wbase + slot;
    if (widx >= w_end) continue;                    // warp-uniform; no CTA-wide barrier inside the loop
    int wx, wy, n;
    window_of(widx, wx, wy, n);
    __syncwarp();                                   // the previous window's xs / ks / vs of this warp are consumed
    // 1) load (+ LayerNorm): tokens lane and lane + 32
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int tok = lane + 32 * r;
      const int py = wy * kWin + (tok >> 3), px = wx * kWin + (tok & 7);
      const float* src = x + (((size_t)n * H + py) * W + px) * CIN;
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
        for (int i = 0; i < C2; ++i) sm.xs[slot][i][tok] = (v[i] - mean) * rstd * sm.lnw[i] + sm.lnb[i];
      } else {
        float v[C2];
        load_vec<C2>(v, src);
#pragma unroll
        for (int i = 0; i < C2; ++i) sm.xs[slot][i][tok] = v[i];
      }
    }
    __syncwarp();
    // 2) q / k / v of this lane's four tokens for its head
    float q[R][D];
    {
      float2 xv[R][C2 / 2];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < C2 / 2; ++k)
          xv[r][k] = make_float2(sm.xs[slot][2 * k][li + 16 * r], sm.xs[slot][2 * k + 1][li + 16 * r]);
      auto dot4 = [&](int o, float (&d)[R]) {
        const float4* wr = reinterpret_cast<const float4*>(&sm.wqkv[o * C2]);
        float2 a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k4 = 0; k4 < C2 / 4; ++k4) {
          const float4 w4 = wr[k4];
          const float2 wa = make_float2(w4.x, w4.y), wb = make_float2(w4.z, w4.w);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            a[r] = __ffma2_rn(wa, xv[r][2 * k4], a[r]);
            a[r] = __ffma2_rn(wb, xv[r][2 * k4 + 1], a[r]);
          }
        }
        const float bias = sm.bqkv[o];
#pragma unroll
        for (int r = 0; r < R; ++r) d[r] = (a[r].x + a[r].y) + bias;
      };
#pragma unroll
      for (int j = 0; j < D; ++j) {
        float qq[R], kk[R], vv[R];
        dot4(head * D + j, qq);
        dot4(C2 + head * D + j, kk);
        dot4(2 * C2 + head * D + j, vv);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          q[r][j] = qq[r] * scale;
          sm.ks[slot][head][j][li + 16 * r] = kk[r];
          sm.vs[slot][head][li + 16 * r][j] = vv[r];
        }
      }
    }
    __syncwarp();
    // 3) online softmax over blocks of KB keys
    float m[R];
    float2 sum2[R], o2[R][D / 2];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      m[r] = -INFINITY;
      sum2[r] = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < D / 2; ++c) o2[r][c] = make_float2(0.f, 0.f);
    }
    constexpr int KB = 4;
#pragma unroll 1
    for (int jb = 0; jb < 64; jb += KB) {
      float2 s[R][2];
      float2 nm[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 p4 = *reinterpret_cast<const float4*>(&sm.pos_t[((head * 16 + (jb >> 2)) * 64 + li + 16 * r) * 4]);
        s[r][0] = make_float2(p4.x, p4.y);
        s[r][1] = make_float2(p4.z, p4.w);
      }
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const float4 kv = *reinterpret_cast<const float4*>(&sm.ks[slot][head][c][jb]);
        const float2 k01 = make_float2(kv.x, kv.y), k23 = make_float2(kv.z, kv.w);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float2 qc = make_float2(q[r][c], q[r][c]);
          s[r][0] = __ffma2_rn(qc, k01, s[r][0]);
          s[r][1] = __ffma2_rn(qc, k23, s[r][1]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float bm = fmaxf(fmaxf(s[r][0].x, s[r][0].y), fmaxf(s[r][1].x, s[r][1].y));
        const float mn = fmaxf(m[r], bm);
        float corr;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(corr) : "f"(m[r] - mn));
        m[r] = mn;
        nm[r] = make_float2(-mn, -mn);
        const float2 c2 = make_float2(corr, corr);
        sum2[r] = __fmul2_rn(sum2[r], c2);
#pragma unroll
        for (int c = 0; c < D / 2; ++c) o2[r][c] = __fmul2_rn(o2[r][c], c2);
      }
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {              // keys jb + 2jj, jb + 2jj + 1
        float2 p[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float2 d = __fadd2_rn(s[r][jj], nm[r]);
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[r].x) : "f"(d.x));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[r].y) : "f"(d.y));
          sum2[r] = __fadd2_rn(sum2[r], p[r]);
        }
#pragma unroll
        for (int c4 = 0; c4 < D; c4 += 4) {
          const float4 v0 = *reinterpret_cast<const float4*>(&sm.vs[slot][head][jb + 2 * jj][c4]);
          const float4 v1 = *reinterpret_cast<const float4*>(&sm.vs[slot][head][jb + 2 * jj + 1][c4]);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float2 p0 = make_float2(p[r].x, p[r].x), p1 = make_float2(p[r].y, p[r].y);
            o2[r][c4 / 2] = __ffma2_rn(p0, make_float2(v0.x, v0.y), o2[r][c4 / 2]);
            o2[r][c4 / 2 + 1] = __ffma2_rn(p0, make_float2(v0.z, v0.w), o2[r][c4 / 2 + 1]);
            o2[r][c4 / 2] = __ffma2_rn(p1, make_float2(v1.x, v1.y), o2[r][c4 / 2]);
            o2[r][c4 / 2 + 1] = __ffma2_rn(p1, make_float2(v1.z, v1.w), o2[r][c4 / 2 + 1]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int tok = li + 16 * r;
      const float inv = 1.0f / (sum2[r].x + sum2[r].y);
      const int py = wy * kWin + (tok >> 3), px = wx * kWin + (tok & 7);
      float* dst = y + (((size_t)n * H + py) * W + px) * C2 + head * D;
#pragma unroll
      for (int c4 = 0; c4 < D; c4 += 4)
        *reinterpret_cast<float4*>(dst + c4) = make_float4(o2[r][c4 / 2].x * inv, o2[r][c4 / 2].y * inv,
                                                           o2[r][c4 / 2 + 1].x * inv, o2[r][c4 / 2 + 1].y * inv);
    }
  }
}
template <int C2>
static cudaError_t launch_msa_q4_t(const BlockW& w, const float* x, float* y, int pre_ln, int N, int H, int W, cudaStream_t s) {
  const int total = N * (H / kWin) * (W / kWin);
  int per_cta = 64;
  while (per_cta > kQ4Slots && (total + per_cta - 1) / per_cta < 2 * 148) per_cta /= 2;
  const int grid = (total + per_cta - 1) / per_cta;
  const size_t smem = sizeof(MsaQ4Smem<C2>);
  cudaError_t e;
  if (pre_ln) {
    e = cudaFuncSetAttribute(window_msa_q4_kernel<C2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    window_msa_q4_kernel<C2, true><<<grid, kQ4Threads, smem, s>>>(x, y, w, H, W, total, per_cta);
  } else {
    e = cudaFuncSetAttribute(window_msa_q4_kernel<C2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    window_msa_q4_kernel<C2, false><<<grid, kQ4Threads, smem, s>>>(x, y, w, H, W, total, per_cta);
  }
  return cudaGetLastError();
}

template <int C2>
static cudaError_t launch_msa_t(const BlockW& w, const float* x, float* y, int pre_ln, int N, int H, int W,
                                cudaStream_t s) {
  const int total = N * (H / kWin) * (W / kWin);
  int per_cta = 32;                                  // amortise the 32 KB pos_emb + weight staging
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
  // LGTEUN_MSA: unset = the fastest measured kernel per channel count (hybrid tcgen05 kernel at c = 32, this
  // CUDA-core kernel at c = 16 and 64); "simt" / "hybrid" / "tc" force one form wherever it is built (A/B measurement)
  static const int mode = [] {
    const char* e = getenv("LGTEUN_MSA");
    return !e ? 0 : e[0] == 's' ? 1 : e[0] == 'h' ? 2 : e[0] == 't' ? 3 : e[0] == 'q' ? 4 : 0;
  }();
  if (mode == 4 && c == 16) return launch_msa_q4_t<8>(w, x, y_half, pre_ln, N, H, W, s);
  if (mode == 4 && c == 32) return launch_msa_q4_t<16>(w, x, y_half, pre_ln, N, H, W, s);
  if (window_msa_tc_supported(c)) {
    if (mode == 3) return launch_window_msa_tc(w, c, x, y_half, pre_ln, N, H, W, 1, s);
    if (mode == 2 || (mode == 0 && c == 32)) return launch_window_msa_tc(w, c, x, y_half, pre_ln, N, H, W, 0, s);
  }
  switch (c) {
    case 16: return launch_msa_t<8>(w, x, y_half, pre_ln, N, H, W, s);
    case 32: return launch_msa_t<16>(w, x, y_half, pre_ln, N, H, W, s);
    case 64: return launch_msa_t<32>(w, x, y_half, pre_ln, N, H, W, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace lg
