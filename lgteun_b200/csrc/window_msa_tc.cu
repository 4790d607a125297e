// window_msa_tc.cu — the local branch of LGMixer (8x8-window multi-head self-attention on the first half of the
// channels, models/common/LGT.py:112-146, window merge :207-208) as a tcgen05 / TMEM kernel, fused with the pre-norm
// LayerNorm (LGT.py:54-61) and the to_qkv 1x1 conv.   c = 16 (C2 = 8, head dim 4) and c = 32 (C2 = 16, head dim 8).
//
//   per window (64 tokens, token = i*8+j, LGT.py:135) and head (2 heads, channel = head*d + c, LGT.py:138):
//     q,k,v = to_qkv(x_win) split in that order along out-channels (LGT.py:136)
//     out   = softmax(q*d^-0.5 . k^T + pos_emb[head]) . v                               (LGT.py:139-143)
//
// Mapping to the tensor pipe (all GEMMs M = 128, fp16 hi/lo split operands, fp32 accumulators in TMEM):
//   QKV  [128 tokens of TWO windows] x [3*C2 outputs], K = C2:                 D_qkv lane = token
//   S_w  [128 rows = (head, query) of ONE window] x [64 keys], K = 2*d: row (h,i) carries q_h[i] in the K slots of head h
//        and zeros in the other head's, the key operand is [k_0 | k_1]; the accumulator is PRE-LOADED with
//        pos_emb[h][i][:] * log2(e) through tcgen05.st, so the MMA adds the positional bias for free
//   O_w  [128 rows] x [v_0 | v_1 | 1], K = 64 keys: A = exp2(logit - rowmax) written back as fp16 hi/lo; the column of
//        ones delivers the softmax denominator from the same MMA
// A thread owns one TMEM lane = one (head, query) row for the whole kernel, so its 64 positional-bias values live in
// registers; the softmax is a row-local max (3-input FMNMX), ex2, and the hi/lo split of the probabilities — no
// shuffles, no shared-memory traffic for K / V, no dot-product FMAs on the CUDA cores.  The window partition / reverse
// of the reference are index arithmetic on the tile loads and stores.
// 256 threads = two windows in flight (warps 0-3: rows of window A, warps 4-7: window B) + one MMA-issuer warp; two CTAs
// per SM hide each other's MMA round trips.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace lg {

using namespace tc;

namespace msatc {

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float max3(float a, float b, float c) {
  float m;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
  return m;
}
__device__ __forceinline__ void split1(float v, __half& h, __half& l) {
  h = __float2half_rn(v);
  l = __float2half_rn(v - __half2float(h));
}

template <int C2>
struct Smem {
  static constexpr int D = C2 / kHeads;
  static constexpr int NQKV = (3 * C2 + 15) / 16 * 16;        // 32 (24 used) or 48
  static constexpr int NO = 16;                               // [v_0 | v_1 | 1 | 0...]; C2 = 16 fills all 16 rows with v
  static constexpr bool kSumInO = (2 * D < NO);               // room for the ones column inside the O operand
  uint64_t mma_qkv, mma_s[2], mma_o[2];                       // tcgen05.commit arrivals
  uint64_t ready_x, ready_s, ready_p[2];                      // operand-ready (thread arrivals)
  uint32_t tmem_base;
  uint32_t pad_[3];
  alignas(16) __half axh[2 * 128 * 8], axl[2 * 128 * 8];      // A of QKV  [K/8 = 2][128 tokens][8]
  alignas(16) __half wqh[2 * NQKV * 8], wql[2 * NQKV * 8];    // B of QKV  [2][NQKV][8]
  alignas(16) __half ash[2][2 * 128 * 8], asl[2][2 * 128 * 8];  // A of S_w [2][128 rows][8]
  alignas(16) __half bsh[2][2 * 64 * 8], bsl[2][2 * 64 * 8];    // B of S_w [2][64 keys][8]
  alignas(16) __half aoh[2][8 * 128 * 8], aol[2][8 * 128 * 8];  // A of O_w [K/8 = 8][128 rows][8]: probabilities
  alignas(16) __half boh[2][2][8 * NO * 8], bol[2][2][8 * NO * 8];  // B of O_w [pair parity][window][8][NO][8]: v^T (+ ones row)
  alignas(16) __half onesh[kSumInO ? 8 : 8 * 16 * 8];          // C2 = 16: separate B operand {1, 0, ...} for the denominators
  float bqkv[3 * C2];
  float lnw[C2], lnb[C2];
};

}  // namespace msatc

template <int C2, bool PRE_LN>
__global__ void __maxnreg__(112)
window_msa_tc_kernel(const float* __restrict__ x, float* __restrict__ y, BlockW w, int H, int W, int total_windows) {
  using namespace msatc;
  using SM = Smem<C2>;
  constexpr int D = SM::D, NQKV = SM::NQKV, NO = SM::NO;
  constexpr bool kSumInO = SM::kSumInO;
  constexpr int CIN = PRE_LN ? 2 * C2 : C2;
  // TMEM columns
  constexpr uint32_t QKV_COL = 0;
  constexpr uint32_t S_COL = (NQKV <= 32) ? 32 : 64;          // S_A, S_B: 64 columns each
  constexpr uint32_t O_COL = S_COL + 128;                     // per window: O (16) [+ SUM (16)]
  constexpr uint32_t O_STRIDE = kSumInO ? 16 : 32;
  constexpr uint32_t TMEM_COLS = 256;
  static_assert(O_COL + 2 * O_STRIDE <= TMEM_COLS, "TMEM plan");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q4 = warp & 3;                  // TMEM lane quarter of this warp
  const int nwx = W / kWin, nwy = H / kWin;

  // ---- one-time setup ---------------------------------------------------------------------------------------------------
  if (tid == 0) {
    mbar_init(&sm.mma_qkv, 1);
    mbar_init(&sm.mma_s[0], 1);
    mbar_init(&sm.mma_s[1], 1);
    mbar_init(&sm.mma_o[0], 1);
    mbar_init(&sm.mma_o[1], 1);
    mbar_init(&sm.ready_x, 256);
    mbar_init(&sm.ready_s, 256);
    mbar_init(&sm.ready_p[0], 128);
    mbar_init(&sm.ready_p[1], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, TMEM_COLS);
  {
    // zero every operand buffer once: the constant-zero K slots / rows are never written again
    uint4* z = reinterpret_cast<uint4*>(sm.axh);
    const int n16 = (int)((reinterpret_cast<unsigned char*>(sm.bqkv) - reinterpret_cast<unsigned char*>(sm.axh)) / 16);
    for (int i = tid; i < n16; i += 288) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  {
    // to_qkv weight [3*C2][C2] -> hi/lo fp16, K-major core-matrix layout [K/8][NQKV][8]
    for (int i = tid; i < 3 * C2 * C2; i += 288) {
      const int n = i / C2, k = i - n * C2;
      __half h, l;
      split1(__ldg(w.qkv_w + i), h, l);
      const int o = ((k >> 3) * NQKV + n) * 8 + (k & 7);
      sm.wqh[o] = h;
      sm.wql[o] = l;
    }
    for (int i = tid; i < 3 * C2; i += 288) sm.bqkv[i] = __ldg(w.qkv_b + i);
    if (PRE_LN)
      for (int i = tid; i < C2; i += 288) { sm.lnw[i] = __ldg(w.ln1_w + i); sm.lnb[i] = __ldg(w.ln1_b + i); }
    const __half one = __float2half_rn(1.0f);
    if (kSumInO) {            // ones row (index 2D) of both windows' O operands
      for (int i = tid; i < 4 * 8 * 8; i += 288) {
        const int wd = i / 64, e = i - wd * 64;
        sm.boh[wd >> 1][wd & 1][((e >> 3) * NO + 2 * D) * 8 + (e & 7)] = one;
      }
    } else {                  // separate operand: row 0 = ones
      for (int i = tid; i < 8 * 8; i += 288) sm.onesh[((i >> 3) * 16 + 0) * 8 + (i & 7)] = one;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
  const int pairs = (total_windows + 1) / 2;

  if (warp == 8) {
    // ---- MMA issuer (whole warp waits, one elected lane issues) -----------------------------------------------------------
    {
      const uint32_t axh = smem_u32(sm.axh), axl = smem_u32(sm.axl), wqh = smem_u32(sm.wqh), wql = smem_u32(sm.wql);
      uint32_t px = 0, ps = 0, pp[2] = {0, 0};
      auto issue_qkv = [&]() {
        mbar_wait_spin(&sm.ready_x, px); px ^= 1;
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc(NQKV);
          const uint64_t ah = umma_desc(axh, 128 * 16, 128), al = umma_desc(axl, 128 * 16, 128);
          const uint64_t bh = umma_desc(wqh, NQKV * 16, 128), bl = umma_desc(wql, NQKV * 16, 128);
          umma_f16(tmem + QKV_COL, ah, bh, idesc, 0);
          umma_f16(tmem + QKV_COL, ah, bl, idesc, 1);
          umma_f16(tmem + QKV_COL, al, bh, idesc, 1);
          umma_commit(&sm.mma_qkv);
        }
        __syncwarp();
      };
      auto issue_s = [&]() {                  // logits: accumulate onto the pre-loaded positional bias
        mbar_wait_spin(&sm.ready_s, ps); ps ^= 1;
        tc_fence_after();
        if (elect_one())
#pragma unroll
        for (int wd = 0; wd < 2; ++wd) {
          constexpr uint32_t idesc = umma_idesc(64);
          const uint64_t ah = umma_desc(smem_u32(sm.ash[wd]), 128 * 16, 128), al = umma_desc(smem_u32(sm.asl[wd]), 128 * 16, 128);
          const uint64_t bh = umma_desc(smem_u32(sm.bsh[wd]), 64 * 16, 128), bl = umma_desc(smem_u32(sm.bsl[wd]), 64 * 16, 128);
          const uint32_t d = tmem + S_COL + 64 * wd;
          umma_f16(d, ah, bh, idesc, 1);
          umma_f16(d, ah, bl, idesc, 1);
          umma_f16(d, al, bh, idesc, 1);
          umma_commit(&sm.mma_s[wd]);
        }
        __syncwarp();
      };
      auto issue_o = [&](int par) {
#pragma unroll
        for (int wd = 0; wd < 2; ++wd) {
          mbar_wait_spin(&sm.ready_p[wd], pp[wd]); pp[wd] ^= 1;
          tc_fence_after();
          if (elect_one()) {
          constexpr uint32_t idesc = umma_idesc(NO);
          const uint32_t aoh = smem_u32(sm.aoh[wd]), aol = smem_u32(sm.aol[wd]);
          const uint32_t boh = smem_u32(sm.boh[par][wd]), bol = smem_u32(sm.bol[par][wd]);
          const uint32_t d = tmem + O_COL + O_STRIDE * wd;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ah = umma_desc(aoh + ks * 2 * 128 * 16, 128 * 16, 128), al = umma_desc(aol + ks * 2 * 128 * 16, 128 * 16, 128);
            const uint64_t bh = umma_desc(boh + ks * 2 * NO * 16, NO * 16, 128), bl = umma_desc(bol + ks * 2 * NO * 16, NO * 16, 128);
            umma_f16(d, ah, bh, idesc, ks > 0);
            umma_f16(d, ah, bl, idesc, 1);
            umma_f16(d, al, bh, idesc, 1);
            if (!kSumInO) {
              const uint64_t oh = umma_desc(smem_u32(sm.onesh) + ks * 2 * 16 * 16, 16 * 16, 128);
              umma_f16(d + 16, ah, oh, umma_idesc(16), ks > 0);
              umma_f16(d + 16, al, oh, umma_idesc(16), 1);
            }
          }
          umma_commit(&sm.mma_o[wd]);
          }
          __syncwarp();
        }
      };
      // pair n: QKV(n) | S(n) | O(n);  issue order QKV(0) S(0) | QKV(n+1) O(n) S(n+1) | ... (mirrors the row warps' phases)
      int par = 0;
      int pr = blockIdx.x;
      if (pr < pairs) {
        issue_qkv();
        issue_s();
        for (; pr < pairs; pr += gridDim.x, par ^= 1) {
          const bool more = pr + (int)gridDim.x < pairs;
          if (more) issue_qkv();
          issue_o(par);
          if (more) issue_s();
        }
      }
    }
  } else {
    // ---- the 8 row warps --------------------------------------------------------------------------------------------
    const int wd3 = warp >> 2;              // window this thread's softmax row belongs to
    const int row = q4 * 32 + lane;         // TMEM lane: (head, query) in the S / O phases, token of the pair in QKV
    const int rh = row >> 6, ri = row & 63;
    // positional bias of this row, pre-scaled by log2(e): pos_t is [h][j/4][i][j%4]
    float pos[64];
#pragma unroll
    for (int j4 = 0; j4 < 16; ++j4) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(w.pos_t + ((size_t)(rh * 16 + j4) * 64 + ri) * 4));
      pos[4 * j4] = p.x; pos[4 * j4 + 1] = p.y; pos[4 * j4 + 2] = p.z; pos[4 * j4 + 3] = p.w;
    }
    // head_channel ** -0.5 rounded to fp32 like the reference's python-float * tensor (LGT.py:119,139), times log2(e)
    const float scale = ((D == 4) ? 0.5f : (D == 8) ? 0.35355339059327379f : 0.25f) * 1.4426950408889634f;
    uint32_t p_qkv = 0, p_s = 0, p_o0 = 0, p_o1 = 0;
    auto signal = [&](uint64_t* bar) {
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar);
    };
    auto window_origin = [&](int widx, int& n, int& py0, int& px0) {
      const int wx = widx % nwx, t = widx / nwx;
      py0 = (t % nwy) * kWin;
      px0 = wx * kWin;
      n = t / nwy;
    };

    // x of the pair P1 will process next, fetched one whole iteration ahead (warps 0-3: thread = token)
    float xr[CIN];
    bool xr_valid = false;
    auto prefetch = [&](int pr) {
      xr_valid = false;
      if (warp < 4 && pr < pairs) {
        const int widx = 2 * pr + (row >> 6);
        if (widx < total_windows) {
          int n, py0, px0;
          window_origin(widx, n, py0, px0);
          load_vec<CIN>(xr, x + (((size_t)n * H + py0 + (ri >> 3)) * W + px0 + (ri & 7)) * CIN);
          xr_valid = true;
        }
      }
    };
    auto phase1 = [&]() {
      // ---- P1: load (+ LayerNorm) the 128 tokens of the pair -> A operand of the QKV GEMM (warps 0-3: thread = token)
      if (warp < 4) {
        float v[C2];
        if (xr_valid) {
          if constexpr (PRE_LN) {
            float (&a)[CIN] = xr;
            float mean = 0.f;
#pragma unroll
            for (int i = 0; i < CIN; ++i) mean += a[i];
            mean *= (1.0f / CIN);
            float var = 0.f;
#pragma unroll
            for (int i = 0; i < CIN; ++i) { const float d = a[i] - mean; var = fmaf(d, d, var); }
            const float rstd = 1.0f / sqrtf(var * (1.0f / CIN) + kLnEps);
#pragma unroll
            for (int i = 0; i < C2; ++i) v[i] = (a[i] - mean) * rstd * sm.lnw[i] + sm.lnb[i];
          } else {
#pragma unroll
            for (int i = 0; i < C2; ++i) v[i] = xr[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < C2; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < C2 / 8; ++c) {
          const float2 t8[4] = {make_float2(v[8 * c], v[8 * c + 1]), make_float2(v[8 * c + 2], v[8 * c + 3]),
                                make_float2(v[8 * c + 4], v[8 * c + 5]), make_float2(v[8 * c + 6], v[8 * c + 7])};
          uint4 hi, lo;
          split8(t8, hi, lo);
          *reinterpret_cast<uint4*>(&sm.axh[(c * 128 + row) * 8]) = hi;
          *reinterpret_cast<uint4*>(&sm.axl[(c * 128 + row) * 8]) = lo;
        }
      }
      signal(&sm.ready_x);
    };
    auto phase2 = [&](int par) {
      // ---- P2: q / k / v of the pair's tokens -> operands of the S and O GEMMs; positional bias -> S accumulators ------
      // the previous pair's O GEMMs must have finished reading the v operands (both windows)
      // (their completion was already observed in P4 below; nothing to wait for here)
      {
        // pre-load this row's accumulator: S = pos (the S region is free: this thread read it in P3 of the previous pair)
        const uint32_t s_addr = lane_addr + S_COL + 64 * wd3;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) t[i] = pos[16 * c + i];
          tmem_st16(s_addr + 16 * c, t);
        }
      }
      mbar_wait_spin(&sm.mma_qkv, p_qkv);
      p_qkv ^= 1;
      tc_fence_after();
      {
        const int tw = row >> 6, ti = row & 63;            // QKV lanes are tokens: window of the pair, token index
        if (warp < 4) {
          // q (scaled) -> rows (0, ti) and (1, ti) of A_s[tw]; k -> row ti of B_s[tw]
          float qk[2 * C2];
#pragma unroll
          for (int c = 0; c < 2 * C2 / 16; ++c) {
            float t[16];
            tmem_ld16(lane_addr + QKV_COL + 16 * c, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) qk[16 * c + i] = t[i];
          }
#pragma unroll
          for (int i = 0; i < C2; ++i) qk[i] = (qk[i] + sm.bqkv[i]) * scale;
#pragma unroll
          for (int i = 0; i < C2; ++i) qk[C2 + i] += sm.bqkv[C2 + i];
          if constexpr (D == 4) {
            // K slots 0-3 = head 0, 4-7 = head 1 (one 8-wide chunk): the other head's slots of a row stay zero
            const float2 q0[4] = {make_float2(qk[0], qk[1]), make_float2(qk[2], qk[3]), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            const float2 q1[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(qk[4], qk[5]), make_float2(qk[6], qk[7])};
            const float2 kk[4] = {make_float2(qk[8], qk[9]), make_float2(qk[10], qk[11]), make_float2(qk[12], qk[13]), make_float2(qk[14], qk[15])};
            uint4 hi, lo;
            split8(q0, hi, lo);
            *reinterpret_cast<uint4*>(&sm.ash[tw][(0 * 128 + ti) * 8]) = hi;
            *reinterpret_cast<uint4*>(&sm.asl[tw][(0 * 128 + ti) * 8]) = lo;
            split8(q1, hi, lo);
            *reinterpret_cast<uint4*>(&sm.ash[tw][(0 * 128 + 64 + ti) * 8]) = hi;
            *reinterpret_cast<uint4*>(&sm.asl[tw][(0 * 128 + 64 + ti) * 8]) = lo;
            split8(kk, hi, lo);
            *reinterpret_cast<uint4*>(&sm.bsh[tw][(0 * 64 + ti) * 8]) = hi;
            *reinterpret_cast<uint4*>(&sm.bsl[tw][(0 * 64 + ti) * 8]) = lo;
          } else {
            // D == 8: K chunk 0 = head 0, chunk 1 = head 1; row (h, ti) writes chunk h only (the other chunk stays zero)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float2 qq[4] = {make_float2(qk[8 * h], qk[8 * h + 1]), make_float2(qk[8 * h + 2], qk[8 * h + 3]),
                                    make_float2(qk[8 * h + 4], qk[8 * h + 5]), make_float2(qk[8 * h + 6], qk[8 * h + 7])};
              const float2 kk[4] = {make_float2(qk[C2 + 8 * h], qk[C2 + 8 * h + 1]), make_float2(qk[C2 + 8 * h + 2], qk[C2 + 8 * h + 3]),
                                    make_float2(qk[C2 + 8 * h + 4], qk[C2 + 8 * h + 5]), make_float2(qk[C2 + 8 * h + 6], qk[C2 + 8 * h + 7])};
              uint4 hi, lo;
              split8(qq, hi, lo);
              *reinterpret_cast<uint4*>(&sm.ash[tw][(h * 128 + 64 * h + ti) * 8]) = hi;
              *reinterpret_cast<uint4*>(&sm.asl[tw][(h * 128 + 64 * h + ti) * 8]) = lo;
              split8(kk, hi, lo);
              *reinterpret_cast<uint4*>(&sm.bsh[tw][(h * 64 + ti) * 8]) = hi;
              *reinterpret_cast<uint4*>(&sm.bsl[tw][(h * 64 + ti) * 8]) = lo;
            }
          }
        } else {
          // v -> column ti (key) of rows 0 .. C2-1 of B_o[tw]  (K-major operand: keys are the K dimension)
          float vv[C2];
#pragma unroll
          for (int c = 0; c < C2 / 8; ++c) {
            float2 t[4];
            tmem_ld8(lane_addr + QKV_COL + 2 * C2 + 8 * c, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) { vv[8 * c + 2 * i] = t[i].x; vv[8 * c + 2 * i + 1] = t[i].y; }
          }
          __half* bh = &sm.boh[par][tw][((ti >> 3) * NO) * 8 + (ti & 7)];
          __half* bl = &sm.bol[par][tw][((ti >> 3) * NO) * 8 + (ti & 7)];
#pragma unroll
          for (int n = 0; n < C2; ++n) {
            __half h, l;
            split1(vv[n] + sm.bqkv[2 * C2 + n], h, l);
            bh[n * 8] = h;
            bl[n * 8] = l;
          }
        }
      }
      tmem_st_wait();
      signal(&sm.ready_s);
    };
    auto phase3 = [&]() {
      // ---- P3: softmax of this thread's row (head rh, query ri of window wd3) -> probabilities as A of the O GEMM -------
      mbar_wait_spin(&sm.mma_s[wd3], p_s);
      p_s ^= 1;
      tc_fence_after();
      {
        const uint32_t s_addr = lane_addr + S_COL + 64 * wd3;
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t[16];
          tmem_ld16(s_addr + 16 * c, t);
          tmem_ld_wait();
          const float a = max3(t[0], t[1], t[2]), b = max3(t[3], t[4], t[5]), cc = max3(t[6], t[7], t[8]);
          const float d = max3(t[9], t[10], t[11]), e = max3(t[12], t[13], t[14]);
          m = fmaxf(max3(max3(a, b, cc), max3(d, e, t[15]), m), m);
        }
        const float2 nm = make_float2(-m, -m);
        __half* aoh = sm.aoh[wd3];
        __half* aol = sm.aol[wd3];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t[16];
          tmem_ld16(s_addr + 16 * c, t);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float2 p[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 d = __fadd2_rn(make_float2(t[8 * g + 2 * i], t[8 * g + 2 * i + 1]), nm);
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[i].x) : "f"(d.x));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[i].y) : "f"(d.y));
            }
            uint4 hi, lo;
            split8(p, hi, lo);
            const int kc = 2 * c + g;                       // keys 8 kc .. 8 kc + 7
            *reinterpret_cast<uint4*>(&aoh[(kc * 128 + row) * 8]) = hi;
            *reinterpret_cast<uint4*>(&aol[(kc * 128 + row) * 8]) = lo;
          }
        }
      }
      signal(&sm.ready_p[wd3]);
    };
    auto phase4 = [&](int pr) {
      // ---- P4: normalise and store ----------------------------------------------------------------------------------------
      // every thread observes BOTH O GEMMs: the next pair's P2 overwrites the v operands of both windows
      mbar_wait_spin(&sm.mma_o[0], p_o0);
      p_o0 ^= 1;
      mbar_wait_spin(&sm.mma_o[1], p_o1);
      p_o1 ^= 1;
      tc_fence_after();
      {
        const uint32_t o_addr = lane_addr + O_COL + O_STRIDE * wd3;
        float o[16];
        tmem_ld16(o_addr, o);
        float den;
        if constexpr (kSumInO) {
          tmem_ld_wait();
          den = o[2 * D];
        } else {
          float2 t[4];
          tmem_ld8(o_addr + 16, t);
          tmem_ld_wait();
          den = t[0].x;
        }
        const int widx = 2 * pr + wd3;
        if (widx < total_windows) {
          int n, py0, px0;
          window_origin(widx, n, py0, px0);
          const float inv = 1.0f / den;
          float* dst = y + (((size_t)n * H + py0 + (ri >> 3)) * W + px0 + (ri & 7)) * C2 + rh * D;
#pragma unroll
          for (int c4 = 0; c4 < D; c4 += 4)
            *reinterpret_cast<float4*>(dst + c4) = make_float4(o[rh * D + c4] * inv, o[rh * D + c4 + 1] * inv,
                                                               o[rh * D + c4 + 2] * inv, o[rh * D + c4 + 3] * inv);
        }
      }
      tc_fence_before();      // TMEM reads of this pair are done before the next pair's accumulators are written
    };
    // software pipeline over the CTA's pairs: every MMA round trip flies under a compute phase of the neighbouring pair
    //   P1(n+1) [QKV(n+1)] | P3(n) [O(n)] | P2(n+1) [S(n+1)] | P4(n)
    {
      int pr = blockIdx.x, par = 0;
      if (pr < pairs) {
        prefetch(pr);
        phase1();
        prefetch(pr + (int)gridDim.x);
        phase2(par);
        for (; pr < pairs; pr += gridDim.x, par ^= 1) {
          const int nx = pr + (int)gridDim.x;
          if (nx < pairs) {
            phase1();
            prefetch(nx + (int)gridDim.x);
          }
          phase3();
          if (nx < pairs) phase2(par ^ 1);
          phase4(pr);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------------
// Hybrid form: QKV projection and Q.K^T (+ positional bias) on tcgen05 / TMEM, softmax and P.V on the CUDA cores.
// The probabilities never leave registers, so they need no fp16 hi/lo split (5 issue slots per pair in the full
// tensor-core form above, a third of its instructions) and no 64 KB operand buffer: a CTA is ONE window (128 threads =
// (head, query) rows + an issuer warp, 128 TMEM columns, ~25 KB of shared memory), several CTAs per SM hide the two MMA
// round trips of a window.  V stays fp32 in shared memory ([key][channel], warp-uniform 128-bit broadcasts).
// ------------------------------------------------------------------------------------------------------------------
namespace msaqk {
template <int C2>
struct Smem {
  static constexpr int D = C2 / kHeads;
  static constexpr int NQKV = (3 * C2 + 15) / 16 * 16;
  uint64_t mma_qkv, mma_s, ready_x, ready_s;
  uint32_t tmem_base;
  uint32_t pad_[3];
  alignas(16) __half axh[2 * 128 * 8], axl[2 * 128 * 8];      // A of QKV [2][128][8]: rows 0-63 tokens, rows 64-127 the same tokens again
  alignas(16) __half wqh[2 * NQKV * 8], wql[2 * NQKV * 8];
  alignas(16) __half ash[2 * 128 * 8], asl[2 * 128 * 8];      // A of S [2][128 rows = (head, query)][8]
  alignas(16) __half bsh[2 * 64 * 8], bsl[2 * 64 * 8];        // B of S [2][64 keys][8]
  alignas(16) float v[64 * C2];                               // V [key][channel] fp32
  float bqkv[3 * C2];
  float lnw[C2], lnb[C2];
};
}  // namespace msaqk

template <int C2, bool PRE_LN>
__global__ void __maxnreg__(96)
window_msa_qk_kernel(const float* __restrict__ x, float* __restrict__ y, BlockW w, int H, int W, int total_windows) {
  using namespace msatc;
  using SM = msaqk::Smem<C2>;
  constexpr int D = SM::D, NQKV = SM::NQKV;
  constexpr int CIN = PRE_LN ? 2 * C2 : C2;
  constexpr uint32_t QKV_COL = 0, S_COL = 64, TMEM_COLS = 128;
  static_assert(NQKV <= 64, "TMEM plan");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwx = W / kWin, nwy = H / kWin;

  if (tid == 0) {
    mbar_init(&sm.mma_qkv, 1);
    mbar_init(&sm.mma_s, 1);
    mbar_init(&sm.ready_x, 128);
    mbar_init(&sm.ready_s, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, TMEM_COLS);
  {
    uint4* z = reinterpret_cast<uint4*>(sm.axh);
    const int n16 = (int)((reinterpret_cast<unsigned char*>(sm.v) - reinterpret_cast<unsigned char*>(sm.axh)) / 16);
    for (int i = tid; i < n16; i += 160) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  for (int i = tid; i < 3 * C2 * C2; i += 160) {
    const int n = i / C2, k = i - n * C2;
    __half h, l;
    split1(__ldg(w.qkv_w + i), h, l);
    const int o = ((k >> 3) * NQKV + n) * 8 + (k & 7);
    sm.wqh[o] = h;
    sm.wql[o] = l;
  }
  for (int i = tid; i < 3 * C2; i += 160) sm.bqkv[i] = __ldg(w.qkv_b + i);
  if (PRE_LN)
    for (int i = tid; i < C2; i += 160) { sm.lnw[i] = __ldg(w.ln1_w + i); sm.lnb[i] = __ldg(w.ln1_b + i); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 4) {
    // MMA issuer: the whole warp waits, one elected lane issues (elect.sync lets ptxas keep the operands in uniform registers
    // without a divergence loop around every tcgen05.mma)
    const uint32_t axh = smem_u32(sm.axh), axl = smem_u32(sm.axl), wqh = smem_u32(sm.wqh), wql = smem_u32(sm.wql);
    const uint32_t ash = smem_u32(sm.ash), asl = smem_u32(sm.asl), bsh = smem_u32(sm.bsh), bsl = smem_u32(sm.bsl);
    uint32_t px = 0, ps = 0;
    for (int widx = blockIdx.x; widx < total_windows; widx += gridDim.x) {
      mbar_wait(&sm.ready_x, px); px ^= 1;
      tc_fence_after();
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc(NQKV);
        const uint64_t ah = umma_desc(axh, 128 * 16, 128), al = umma_desc(axl, 128 * 16, 128);
        const uint64_t bh = umma_desc(wqh, NQKV * 16, 128), bl = umma_desc(wql, NQKV * 16, 128);
        umma_f16(tmem + QKV_COL, ah, bh, idesc, 0);
        umma_f16(tmem + QKV_COL, ah, bl, idesc, 1);
        umma_f16(tmem + QKV_COL, al, bh, idesc, 1);
        umma_commit(&sm.mma_qkv);
      }
      __syncwarp();
      mbar_wait(&sm.ready_s, ps); ps ^= 1;
      tc_fence_after();
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc(64);
        const uint64_t ah = umma_desc(ash, 128 * 16, 128), al = umma_desc(asl, 128 * 16, 128);
        const uint64_t bh = umma_desc(bsh, 64 * 16, 128), bl = umma_desc(bsl, 64 * 16, 128);
        umma_f16(tmem + S_COL, ah, bh, idesc, 1);        // accumulate onto the pre-loaded positional bias
        umma_f16(tmem + S_COL, ah, bl, idesc, 1);
        umma_f16(tmem + S_COL, al, bh, idesc, 1);
        umma_commit(&sm.mma_s);
      }
      __syncwarp();
    }
  } else {
    const int row = tid;                       // TMEM lane: (head, query) in the S phase; token (twice) in the QKV phase
    const int rh = row >> 6, ri = row & 63;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    float pos[64];
#pragma unroll
    for (int j4 = 0; j4 < 16; ++j4) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(w.pos_t + ((size_t)(rh * 16 + j4) * 64 + ri) * 4));
      pos[4 * j4] = p.x; pos[4 * j4 + 1] = p.y; pos[4 * j4 + 2] = p.z; pos[4 * j4 + 3] = p.w;
    }
    const float scale = ((D == 4) ? 0.5f : (D == 8) ? 0.35355339059327379f : 0.25f) * 1.4426950408889634f;
    uint32_t p_qkv = 0, p_s = 0;
    auto signal = [&](uint64_t* bar) {
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar);
    };
    // window coordinates: shifts when the window grid is a power of two in both directions (every forward shape)
    const bool grid_pow2 = ((nwx & (nwx - 1)) | (nwy & (nwy - 1))) == 0;
    const int lg_nwx = 31 - __clz(nwx), lg_nwy = 31 - __clz(nwy);
    auto pixel_of = [&](int widx) {
      int wx, wy, n;
      if (grid_pow2) {
        wx = widx & (nwx - 1);
        const int wt = widx >> lg_nwx;
        wy = wt & (nwy - 1);
        n = wt >> lg_nwy;
      } else {
        wx = widx % nwx;
        const int wt = widx / nwx;
        wy = wt % nwy;
        n = wt / nwy;
      }
      return ((size_t)n * H + wy * kWin + (ri >> 3)) * W + wx * kWin + (ri & 7);
    };
    // optional register prefetch of the NEXT window's pixel by warps 0-1.  Measured (c = 16, 64 pairs): 404 us with it, 386 us
    // without (the extra live registers spill at the 96-register budget that four CTAs per SM need); a cp.async staging
    // buffer was slower as well (425 us).  Off.
    constexpr bool kPrefetch = false;
    float xr[CIN];
    auto fetch = [&](int widx) {
      if (kPrefetch && warp < 2 && widx < total_windows) load_vec<CIN>(xr, x + pixel_of(widx) * CIN);
    };
    fetch(blockIdx.x);
    for (int widx = blockIdx.x; widx < total_windows; widx += gridDim.x) {
      const size_t pix = pixel_of(widx);
      // ---- P1: token ri -> rows ri and 64 + ri of the QKV operand (warps 0-1 load / normalise, and store both copies)
      if (warp < 2) {
        float v[C2];
        const float* src = x + pix * CIN;
        if constexpr (PRE_LN) {
          float a[CIN];
          if constexpr (kPrefetch) {
#pragma unroll
            for (int i = 0; i < CIN; ++i) a[i] = xr[i];
          } else {
            load_vec<CIN>(a, src);
          }
          float mean = 0.f;
#pragma unroll
          for (int i = 0; i < CIN; ++i) mean += a[i];
          mean *= (1.0f / CIN);
          float var = 0.f;
#pragma unroll
          for (int i = 0; i < CIN; ++i) { const float d = a[i] - mean; var = fmaf(d, d, var); }
          const float rstd = 1.0f / sqrtf(var * (1.0f / CIN) + kLnEps);
#pragma unroll
          for (int i = 0; i < C2; ++i) v[i] = (a[i] - mean) * rstd * sm.lnw[i] + sm.lnb[i];
        } else {
          if constexpr (kPrefetch) {
#pragma unroll
            for (int i = 0; i < C2; ++i) v[i] = xr[i];
          } else {
            load_vec<C2>(v, src);
          }
        }
        fetch(widx + (int)gridDim.x);            // the next window's pixel flies under the rest of this window
#pragma unroll
        for (int c = 0; c < C2 / 8; ++c) {
          const float2 t8[4] = {make_float2(v[8 * c], v[8 * c + 1]), make_float2(v[8 * c + 2], v[8 * c + 3]),
                                make_float2(v[8 * c + 4], v[8 * c + 5]), make_float2(v[8 * c + 6], v[8 * c + 7])};
          uint4 hi, lo;
          split8(t8, hi, lo);
          *reinterpret_cast<uint4*>(&sm.axh[(c * 128 + row) * 8]) = hi;
          *reinterpret_cast<uint4*>(&sm.axl[(c * 128 + row) * 8]) = lo;
          *reinterpret_cast<uint4*>(&sm.axh[(c * 128 + 64 + row) * 8]) = hi;
          *reinterpret_cast<uint4*>(&sm.axl[(c * 128 + 64 + row) * 8]) = lo;
        }
      }
      signal(&sm.ready_x);
      // ---- P2: S accumulator = pos; q, k -> operands of the S GEMM (rows 0-63); v -> shared memory (rows 64-127)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float t[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = pos[16 * c + i];
        tmem_st16(lane_addr + S_COL + 16 * c, t);
      }
      mbar_wait(&sm.mma_qkv, p_qkv);
      p_qkv ^= 1;
      tc_fence_after();
      if (warp < 2) {
        float qk[2 * C2];
#pragma unroll
        for (int c = 0; c < 2 * C2 / 16; ++c) {
          float t[16];
          tmem_ld16(lane_addr + QKV_COL + 16 * c, t);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) qk[16 * c + i] = t[i];
        }
#pragma unroll
        for (int i = 0; i < C2; ++i) qk[i] = (qk[i] + sm.bqkv[i]) * scale;
#pragma unroll
        for (int i = 0; i < C2; ++i) qk[C2 + i] += sm.bqkv[C2 + i];
        if constexpr (D == 4) {
          const float2 q0[4] = {make_float2(qk[0], qk[1]), make_float2(qk[2], qk[3]), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
          const float2 q1[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(qk[4], qk[5]), make_float2(qk[6], qk[7])};
          const float2 kk[4] = {make_float2(qk[8], qk[9]), make_float2(qk[10], qk[11]), make_float2(qk[12], qk[13]), make_float2(qk[14], qk[15])};
          uint4 hi, lo;
          split8(q0, hi, lo);
          *reinterpret_cast<uint4*>(&sm.ash[(0 * 128 + ri) * 8]) = hi;
          *reinterpret_cast<uint4*>(&sm.asl[(0 * 128 + ri) * 8]) = lo;
          split8(q1, hi, lo);
          *reinterpret_cast<uint4*>(&sm.ash[(0 * 128 + 64 + ri) * 8]) = hi;
          *reinterpret_cast<uint4*>(&sm.asl[(0 * 128 + 64 + ri) * 8]) = lo;
          split8(kk, hi, lo);
          *reinterpret_cast<uint4*>(&sm.bsh[(0 * 64 + ri) * 8]) = hi;
          *reinterpret_cast<uint4*>(&sm.bsl[(0 * 64 + ri) * 8]) = lo;
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float2 qq[4] = {make_float2(qk[8 * h], qk[8 * h + 1]), make_float2(qk[8 * h + 2], qk[8 * h + 3]),
                                  make_float2(qk[8 * h + 4], qk[8 * h + 5]), make_float2(qk[8 * h + 6], qk[8 * h + 7])};
            const float2 kk[4] = {make_float2(qk[C2 + 8 * h], qk[C2 + 8 * h + 1]), make_float2(qk[C2 + 8 * h + 2], qk[C2 + 8 * h + 3]),
                                  make_float2(qk[C2 + 8 * h + 4], qk[C2 + 8 * h + 5]), make_float2(qk[C2 + 8 * h + 6], qk[C2 + 8 * h + 7])};
            uint4 hi, lo;
            split8(qq, hi, lo);
            *reinterpret_cast<uint4*>(&sm.ash[(h * 128 + 64 * h + ri) * 8]) = hi;
            *reinterpret_cast<uint4*>(&sm.asl[(h * 128 + 64 * h + ri) * 8]) = lo;
            split8(kk, hi, lo);
            *reinterpret_cast<uint4*>(&sm.bsh[(h * 64 + ri) * 8]) = hi;
            *reinterpret_cast<uint4*>(&sm.bsl[(h * 64 + ri) * 8]) = lo;
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < C2 / 8; ++c) {
          float2 t[4];
          tmem_ld8(lane_addr + QKV_COL + 2 * C2 + 8 * c, t);
          tmem_ld_wait();
          const float* bv = &sm.bqkv[2 * C2 + 8 * c];
          *reinterpret_cast<float4*>(&sm.v[ri * C2 + 8 * c]) = make_float4(t[0].x + bv[0], t[0].y + bv[1], t[1].x + bv[2], t[1].y + bv[3]);
          *reinterpret_cast<float4*>(&sm.v[ri * C2 + 8 * c + 4]) = make_float4(t[2].x + bv[4], t[2].y + bv[5], t[3].x + bv[6], t[3].y + bv[7]);
        }
      }
      tmem_st_wait();
      signal(&sm.ready_s);
      // ---- P3: softmax of row (rh, ri) and P.V on the CUDA cores ---------------------------------------------------
      mbar_wait(&sm.mma_s, p_s);             // (the arrivals on ready_s also ordered the v stores of warps 2-3)
      p_s ^= 1;
      tc_fence_after();
      asm volatile("bar.sync 1, 128;" ::: "memory");   // v of all 64 keys visible to the 128 row threads
      {
        const uint32_t s_addr = lane_addr + S_COL;
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t[16];
          tmem_ld16(s_addr + 16 * c, t);
          tmem_ld_wait();
          const float a = max3(t[0], t[1], t[2]), b = max3(t[3], t[4], t[5]), cc = max3(t[6], t[7], t[8]);
          const float d = max3(t[9], t[10], t[11]), e = max3(t[12], t[13], t[14]);
          m = max3(max3(a, b, cc), max3(d, e, t[15]), m);
        }
        const float2 nm = make_float2(-m, -m);
        float2 sum2 = make_float2(0.f, 0.f);
        float2 o2[D / 2];
#pragma unroll
        for (int i = 0; i < D / 2; ++i) o2[i] = make_float2(0.f, 0.f);
        const float* vb = &sm.v[rh * D];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t[16];
          tmem_ld16(s_addr + 16 * c, t);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 d = __fadd2_rn(make_float2(t[2 * i], t[2 * i + 1]), nm);
            float2 p;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p.x) : "f"(d.x));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p.y) : "f"(d.y));
            sum2 = __fadd2_rn(sum2, p);
            const int j = 16 * c + 2 * i;
#pragma unroll
            for (int c4 = 0; c4 < D; c4 += 4) {
              const float4 v0 = *reinterpret_cast<const float4*>(vb + j * C2 + c4);
              const float4 v1 = *reinterpret_cast<const float4*>(vb + (j + 1) * C2 + c4);
              const float2 p0 = make_float2(p.x, p.x), p1 = make_float2(p.y, p.y);
              o2[c4 / 2] = __ffma2_rn(p0, make_float2(v0.x, v0.y), o2[c4 / 2]);
              o2[c4 / 2 + 1] = __ffma2_rn(p0, make_float2(v0.z, v0.w), o2[c4 / 2 + 1]);
              o2[c4 / 2] = __ffma2_rn(p1, make_float2(v1.x, v1.y), o2[c4 / 2]);
              o2[c4 / 2 + 1] = __ffma2_rn(p1, make_float2(v1.z, v1.w), o2[c4 / 2 + 1]);
            }
          }
        }
        const float inv = 1.0f / (sum2.x + sum2.y);
        float* dst = y + pix * C2 + rh * D;
#pragma unroll
        for (int c4 = 0; c4 < D; c4 += 4)
          *reinterpret_cast<float4*>(dst + c4) = make_float4(o2[c4 / 2].x * inv, o2[c4 / 2].y * inv, o2[c4 / 2 + 1].x * inv, o2[c4 / 2 + 1].y * inv);
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");   // every row is done with v before the next window's P2 overwrites it
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

template <int C2>
static cudaError_t launch_msa_qk_t(const BlockW& w, const float* x, float* y, int pre_ln, int N, int H, int W, cudaStream_t s) {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  const int total = N * (H / kWin) * (W / kWin);
  const size_t smem = sizeof(msaqk::Smem<C2>) + 128;
  const int cap = 4 * sm_count;                              // 4 CTAs per SM: 4 x 128 TMEM columns
  const int grid = total < cap ? total : cap;
  cudaError_t e;
  if (pre_ln) {
    e = cudaFuncSetAttribute(window_msa_qk_kernel<C2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    window_msa_qk_kernel<C2, true><<<grid, 160, smem, s>>>(x, y, w, H, W, total);
  } else {
    e = cudaFuncSetAttribute(window_msa_qk_kernel<C2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    window_msa_qk_kernel<C2, false><<<grid, 160, smem, s>>>(x, y, w, H, W, total);
  }
  return cudaGetLastError();
}

template <int C2>
static cudaError_t launch_msa_tc_t(const BlockW& w, const float* x, float* y, int pre_ln, int N, int H, int W, cudaStream_t s) {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  const int total = N * (H / kWin) * (W / kWin);
  const int pairs = (total + 1) / 2;
  const size_t smem = sizeof(msatc::Smem<C2>) + 128;
  const int grid = pairs < 2 * sm_count ? pairs : 2 * sm_count;
  cudaError_t e;
  if (pre_ln) {
    e = cudaFuncSetAttribute(window_msa_tc_kernel<C2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    window_msa_tc_kernel<C2, true><<<grid, 288, smem, s>>>(x, y, w, H, W, total);
  } else {
    e = cudaFuncSetAttribute(window_msa_tc_kernel<C2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    window_msa_tc_kernel<C2, false><<<grid, 288, smem, s>>>(x, y, w, H, W, total);
  }
  return cudaGetLastError();
}

bool window_msa_tc_supported(int c) { return c == 16 || c == 32; }

// variant: 0 = hybrid (QKV + Q.K^T on the tensor pipe, softmax + P.V on the CUDA cores), 1 = all three GEMMs on the tensor pipe
cudaError_t launch_window_msa_tc(const BlockW& w, int c, const float* x, float* y_half, int pre_ln, int N, int H, int W,
                                 int variant, cudaStream_t s) {
  if (variant == 0) {
    switch (c) {
      case 16: return launch_msa_qk_t<8>(w, x, y_half, pre_ln, N, H, W, s);
      case 32: return launch_msa_qk_t<16>(w, x, y_half, pre_ln, N, H, W, s);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (c) {
    case 16: return launch_msa_tc_t<8>(w, x, y_half, pre_ln, N, H, W, s);
    case 32: return launch_msa_tc_t<16>(w, x, y_half, pre_ln, N, H, W, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace lg
