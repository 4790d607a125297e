// ffn_tc.cu — residual(pre_norm(feed_forward)) as ONE fused tcgen05/TMEM kernel (sm_100a).
//   reference: models/common/LGT.py:91-109 (feed_forward), :45-61 (residual/pre_norm),
//              models/common/basic_module_unformer_v2.py:37-53 (depthwise_conv = 1x1 then dw3x3, zero pad)
//     y = x + W2 . GELU( dw3x3( W1 . GELU( W0 . LN(x) + b0 ) + b1 ) + bdw ) + b2
//
// Mapping to the hardware
//   * The three 1x1 convs are GEMMs with M = pixels, N/K = channels, issued as tcgen05.mma kind::f16 by one thread,
//     accumulating in TMEM.  fp32 parity (max |delta| <= 1e-3 end to end) rules out plain bf16/tf32 operands
//     (SURVEY.md F8), so every operand is split x = hi + lo in fp16 (2 x 11-bit mantissas) and each GEMM issues
//     hi*hi + hi*lo + lo*hi into the same fp32 accumulator: the accuracy of 3xTF32 at twice the MMA rate.
//   * A 128-row MMA tile is ONE IMAGE ROW segment: TMEM lane = pixel column.  The hidden activation of the
//     depthwise 3x3 therefore never leaves TMEM for its vertical taps (three accumulator slots = rows y-1, y, y+1
//     in the same lane) and reaches its horizontal taps with two warp shuffles; there is no halo buffer in shared
//     memory and no 4c-channel hidden tensor in HBM (the reference round-trips 4 x 16.8 MB per pair per block).
//   * Each warp quarter of the tile is an independent 32-pixel strip (30 interior + 2 halo columns, recomputed);
//     a CTA streams R+2 rows of four such strips top to bottom.  Zero padding of the conv = masking lanes/rows
//     outside the image after the bias.
//   * Operands are written to shared memory by the epilogue threads in the canonical no-swizzle K-major core-matrix
//     layout [K/8][rows][8] (LBO = rows*16 B, SBO = 128 B): one 16-byte store per thread per K-chunk, conflict-free.
//   * 128*G threads: warp w owns TMEM lanes 32*(w%4).. and the (w/4)-th slice of the channels, so all CUDA cores
//     work on the GELU/conv epilogues, which bound this kernel (~70 instructions per hidden element; the tensor
//     pipe is <25 % busy).  C=16 runs two CTAs per SM (2 x 256 TMEM columns) so one CTA's MMAs hide under the
//     other's epilogue.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace lg {

constexpr int kG16 = 2;              // channel slices (thread groups of 128) per CTA at c = 16.  Measured: 4 (512 threads, 64-register
                                     // cap, ~0.5 KB of spills per thread) runs 2.6x slower than 2 (256 threads, 128 registers)
constexpr int kStripW = 30;          // interior pixels per warp strip (32 lanes - 2 halo lanes)
// output rows streamed by one CTA pass (band_rows) are chosen per launch: 64 when the grid is deep, fewer for small batches

using namespace tc;   // PTX wrappers, descriptors and the fp16 hi/lo split: tc_ptx.cuh

// ---- weight packing (load time): W [N][K] fp32 -> hi/lo fp16 in the UMMA K-major core-matrix layout [Kext/8][N][8] ---
// Kext = K + 8 appends one core-matrix chunk whose first column is the bias (the A operand carries a matching column of
// ones, so the GEMM adds the bias and a zeroed A row yields an exactly zero output row).
__global__ void pack_umma_f16_kernel(const float* __restrict__ w, const float* __restrict__ bias, __half* __restrict__ hi,
                                     __half* __restrict__ lo, int N, int K, int Kext) {
  int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= N * Kext) return;
  int n = idx / Kext, k = idx - n * Kext;
  float v = k < K ? w[n * K + k] : (k == K && bias) ? bias[n] : 0.f;
  __half h = __float2half_rn(v);
  __half l = __float2half_rn(v - __half2float(h));
  size_t o = ((size_t)(k >> 3) * N + n) * 8 + (k & 7);
  hi[o] = h;
  lo[o] = l;
}
cudaError_t launch_pack_umma_f16(const float* w, void* hi, void* lo, int N, int K, cudaStream_t s) {
  pack_umma_f16_kernel<<<(N * K + 255) / 256, 256, 0, s>>>(w, nullptr, (__half*)hi, (__half*)lo, N, K, K);
  return cudaGetLastError();
}
cudaError_t launch_pack_umma_f16_bias(const float* w, const float* bias, void* hi, void* lo, int N, int K, cudaStream_t s) {
  pack_umma_f16_kernel<<<(N * (K + 8) + 255) / 256, 256, 0, s>>>(w, bias, (__half*)hi, (__half*)lo, N, K, K + 8);
  return cudaGetLastError();
}

// ---- shared memory plan -------------------------------------------------------------------------------------------------
template <int C>
struct FfnTcSmem {
  static constexpr int C4 = 4 * C;
  uint64_t mbar[3];                                   // MMA-completion barriers of G1, G2, G3 (tcgen05.commit arrives)
  uint64_t ready[3];                                  // operand-ready barriers A1, A2, A3 (every epilogue thread arrives)
  uint32_t tmem_base;
  uint32_t pad_[3];
  // K of GEMM1 / GEMM2 is extended by one 16-wide step whose first column is the bias (weights) / the pixel-valid flag
  // (activations): the bias add and the zero padding of the conv ride in the MMA.  Only the first 8-column chunk of that
  // step is stored; its second chunk is the shared block `zero` (reached through the descriptor's leading-dimension offset)
  static constexpr bool kOwnA3 = (C == 16);                         // room for A3 next to A2: GEMM3 overlaps the next row's S_b
  alignas(16) __half w0h[C4 * (C + 8)], w0l[C4 * (C + 8)];          // [(C+8)/8][C4][8]
  alignas(16) __half w1h[C4 * (C4 + 8)], w1l[C4 * (C4 + 8)];        // [(C4+8)/8][C4][8]
  alignas(16) __half w2h[C * C4], w2l[C * C4];                      // [C4/8][C][8]
  alignas(16) __half a1h[128 * (C + 8)], a1l[128 * C];              // [K/8][128][8]; the flag step has no lo part
  alignas(16) __half a2h[128 * (C4 + 8)], a2l[128 * C4];            // A2 (and A3 when it has no buffer of its own)
  alignas(16) __half a3h[kOwnA3 ? 128 * C4 : 8], a3l[kOwnA3 ? 128 * C4 : 8];
  alignas(16) float dwb[C4], dww[9 * C4];                           // dww: [tap][channel]
  alignas(16) float b2[C], lng[C], lnb[C];
  alignas(16) float2 part[C / 8][128];                // LayerNorm partials (mean, M2) of each 8-channel slice of a pixel
  alignas(16) __half zero[128 * 8];                   // second chunk of the flag K-steps (must lie above every operand)
};

static_assert(sizeof(FfnTcSmem<16>) + 128 + 1024 <= 114 * 1024, "c = 16 must keep two CTAs per SM (228 KB, 1 KB reserved each)");
static_assert(sizeof(FfnTcSmem<32>) + 128 <= 227 * 1024, "c = 32: one CTA per SM");

// One CTA: four 30-pixel-wide strips (one per warp quarter), rows y0-1 .. y0+R streamed through the three GEMMs.
template <int C, int G>
__global__ void __launch_bounds__(128 * G + 32, (C == 16) ? 2 : 1)
ffn_tc_kernel(const float* __restrict__ xin, float* __restrict__ yout, BlockW w, const __half* __restrict__ wpack,
              int H, int W, int nws, int nbands, int band_rows, int total_units, int num_groups, int exp_mode) {
  constexpr int C4 = 4 * C;
  constexpr int NT = 128 * G;           // epilogue threads (warps 0 .. 4G-1); warp 4G only issues the MMAs
  constexpr int CH = C4 / G;            // hidden channels per thread
  static_assert(CH % 8 == 0, "channel slices are processed 8 columns at a time");
  constexpr int LG = C / 8;             // thread groups (cg < LG) that own an 8-channel slice of x / y (LayerNorm, residual, store)
  constexpr uint32_t D1_COL = 0;        // [0, C4): GEMM1 accumulator
  constexpr uint32_t D2_COL = C4;       // three slots of C4 columns: hidden rows y-1, y, y+1 (the dead slot hosts GEMM3's accumulator)
  constexpr uint32_t TMEM_COLS = 4 * C4;        // 256 (C=16) or 512 (C=32)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FfnTcSmem<C>& sm = *reinterpret_cast<FfnTcSmem<C>*>(smem_raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3;               // TMEM lane quarter owned by this warp
  const int cg = warp >> 2;             // channel slice
  const int row = q * 32 + lane;        // row of the 128-row MMA tile

  // ---- one-time setup: barrier, TMEM, weights -----------------------------------------------------------------------
  if (tid == 0) {
    mbar_init(&sm.mbar[0], 1);
    mbar_init(&sm.mbar[1], 1);
    mbar_init(&sm.mbar[2], 1);
    mbar_init(&sm.ready[0], NT);
    mbar_init(&sm.ready[1], NT);
    mbar_init(&sm.ready[2], NT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, TMEM_COLS);
  {
    // packed weights in global: w0h | w0l | w1h | w1l | w2h | w2l, contiguous, same order as the smem plan
    constexpr int n16 = (2 * (C4 * (C + 8) + C4 * (C4 + 8) + C * C4)) * 2 / 16;
    const uint4* src = reinterpret_cast<const uint4*>(wpack);
    uint4* dst = reinterpret_cast<uint4*>(sm.w0h);
    for (int i = tid; i < n16; i += NT + 32) dst[i] = __ldg(src + i);
    for (int i = tid; i < C4; i += NT + 32) {
      sm.dwb[i] = __ldg(w.dw_b + i);
#pragma unroll
      for (int t = 0; t < 9; ++t) sm.dww[t * C4 + i] = __ldg(w.dw_w + i * 9 + t);
    }
    for (int i = tid; i < C; i += NT + 32) {
      sm.b2[i] = __ldg(w.f2_b + i);
      sm.lng[i] = __ldg(w.ln2_w + i);
      sm.lnb[i] = __ldg(w.ln2_b + i);
    }
    for (int i = tid; i < 128; i += NT + 32) *reinterpret_cast<uint4*>(&sm.zero[i * 8]) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);

  const uint32_t a1h = smem_u32(sm.a1h), a1l = smem_u32(sm.a1l), a2h = smem_u32(sm.a2h), a2l = smem_u32(sm.a2l);
  const uint32_t w0h = smem_u32(sm.w0h), w0l = smem_u32(sm.w0l), w1h = smem_u32(sm.w1h), w1l = smem_u32(sm.w1l);
  const uint32_t w2h = smem_u32(sm.w2h), w2l = smem_u32(sm.w2l), zero = smem_u32(sm.zero);
  constexpr bool kOwnA3 = FfnTcSmem<C>::kOwnA3;
  const uint32_t a3h = kOwnA3 ? smem_u32(sm.a3h) : a2h, a3l = kOwnA3 ? smem_u32(sm.a3l) : a2l;

  const int iters = min(band_rows, H) + 2;                // H and band_rows are powers of two: every band has the same height

  // ---- the three GEMMs (issued by the dedicated warp) ----------------------------------------------------------------------
  // G1: D1 = [A1 | valid] . [W0 | b0]^T
  auto issue_g1 = [&]() {
    constexpr uint32_t idesc = umma_idesc(C4);
#pragma unroll
    for (int ks = 0; ks < C / 16; ++ks) {
      const uint64_t ah = umma_desc(a1h + ks * 2 * 128 * 16, 128 * 16, 128);
      const uint64_t al = umma_desc(a1l + ks * 2 * 128 * 16, 128 * 16, 128);
      const uint64_t bh = umma_desc(w0h + ks * 2 * C4 * 16, C4 * 16, 128);
      const uint64_t bl = umma_desc(w0l + ks * 2 * C4 * 16, C4 * 16, 128);
      umma_f16(tmem + D1_COL, ah, bh, idesc, ks > 0);
      umma_f16(tmem + D1_COL, ah, bl, idesc, 1);
      umma_f16(tmem + D1_COL, al, bh, idesc, 1);
    }
    {                                                     // + valid * b0 (second K chunk = the zero block)
      const uint32_t fa = a1h + (C / 8) * 128 * 16, fh = w0h + (C / 8) * C4 * 16, fl = w0l + (C / 8) * C4 * 16;
      const uint64_t ah = umma_desc(fa, zero - fa, 128);
      umma_f16(tmem + D1_COL, ah, umma_desc(fh, zero - fh, 128), idesc, 1);
      umma_f16(tmem + D1_COL, ah, umma_desc(fl, zero - fl, 128), idesc, 1);
    }
    umma_commit(&sm.mbar[0]);
  };
  // G2: D2[it % 3] = [A2 | valid] . [W1 | b1]^T: hidden rows / columns outside the image stay exactly 0
  auto issue_g2 = [&](int it) {
    constexpr uint32_t idesc = umma_idesc(C4);
    const uint32_t d = tmem + D2_COL + (uint32_t)(it % 3) * C4;
#pragma unroll
    for (int ks = 0; ks < C4 / 16; ++ks) {
      const uint64_t ah = umma_desc(a2h + ks * 2 * 128 * 16, 128 * 16, 128);
      const uint64_t al = umma_desc(a2l + ks * 2 * 128 * 16, 128 * 16, 128);
      const uint64_t bh = umma_desc(w1h + ks * 2 * C4 * 16, C4 * 16, 128);
      const uint64_t bl = umma_desc(w1l + ks * 2 * C4 * 16, C4 * 16, 128);
      umma_f16(d, ah, bh, idesc, ks > 0);
      umma_f16(d, ah, bl, idesc, 1);
      umma_f16(d, al, bh, idesc, 1);
    }
    {
      const uint32_t fa = a2h + (C4 / 8) * 128 * 16, fh = w1h + (C4 / 8) * C4 * 16, fl = w1l + (C4 / 8) * C4 * 16;
      const uint64_t ah = umma_desc(fa, zero - fa, 128);
      umma_f16(d, ah, umma_desc(fh, zero - fh, 128), idesc, 1);
      umma_f16(d, ah, umma_desc(fl, zero - fl, 128), idesc, 1);
    }
    umma_commit(&sm.mbar[1]);
  };
  // G3: D3 = A3 . W2^T, accumulated in the first C columns of the hidden-row slot that just died
  auto issue_g3 = [&](int it) {
    constexpr uint32_t idesc = umma_idesc(C);
    const uint32_t d = tmem + D2_COL + (uint32_t)((it + 1) % 3) * C4;
#pragma unroll
    for (int ks = 0; ks < C4 / 16; ++ks) {
      const uint64_t ah = umma_desc(a3h + ks * 2 * 128 * 16, 128 * 16, 128);
      const uint64_t al = umma_desc(a3l + ks * 2 * 128 * 16, 128 * 16, 128);
      const uint64_t bh = umma_desc(w2h + ks * 2 * C * 16, C * 16, 128);
      const uint64_t bl = umma_desc(w2l + ks * 2 * C * 16, C * 16, 128);
      umma_f16(d, ah, bh, idesc, ks > 0);
      umma_f16(d, ah, bl, idesc, 1);
      umma_f16(d, al, bh, idesc, 1);
    }
    umma_commit(&sm.mbar[2]);
  };

  if (warp == 4 * G) {
    // ---- MMA issuer: the whole warp waits for the operand-ready barriers in the order the epilogue warps raise them; one
    // elected lane feeds the tensor pipe (under elect.sync ptxas moves the operands to uniform registers without a
    // divergence loop), so no epilogue warp ever carries the issue sequence on its critical path
    uint32_t pa1 = 0, pa2 = 0, pa3 = 0;
    for (int grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
      mbar_wait(&sm.ready[0], pa1); pa1 ^= 1;
      tc_fence_after();
      if (elect_one()) issue_g1();
      __syncwarp();
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&sm.ready[1], pa2); pa2 ^= 1;
        tc_fence_after();
        if (elect_one()) issue_g2(it);
        __syncwarp();
        if (it + 1 < iters) {
          mbar_wait(&sm.ready[0], pa1); pa1 ^= 1;
          tc_fence_after();
          if (elect_one()) issue_g1();
          __syncwarp();
        }
        if (it >= 2) {
          mbar_wait(&sm.ready[2], pa3); pa3 ^= 1;
          tc_fence_after();
          if (elect_one()) issue_g3(it);
          __syncwarp();
        }
      }
    }
  } else {
  // ---- epilogue warps ------------------------------------------------------------------------------------------------------
  uint32_t ph1 = 0, ph2 = 0, ph3 = 0;   // phases of the three MMA-completion barriers (G1, G2, G3)
  // operands written (generic proxy) and TMEM reads finished: publish both, then count this thread in
  auto signal = [&](uint64_t* bar) {
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(bar);
  };
  // LayerNorm partials are exchanged between the G warps that own the same 32 pixels (warps q, q+4, ...): one named
  // barrier per TMEM lane quarter, 32*G threads each
  auto epi_sync = [&]() {
    switch (q) {
      case 0: asm volatile("bar.sync 1, %0;" ::"n"(32 * G) : "memory"); break;
      case 1: asm volatile("bar.sync 2, %0;" ::"n"(32 * G) : "memory"); break;
      case 2: asm volatile("bar.sync 3, %0;" ::"n"(32 * G) : "memory"); break;
      default: asm volatile("bar.sync 4, %0;" ::"n"(32 * G) : "memory"); break;
    }
  };

  for (int grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
    // this warp's strip
    const int unit = grp * 4 + q;
    const bool unit_ok = unit < total_units;
    int n = 0, band = 0, ws = 0;
    if (unit_ok) {
      ws = unit % nws;
      int t = unit / nws;
      band = t % nbands;
      n = t / nbands;
    }
    const int x = ws * kStripW + lane - 1;
    const bool x_ok = unit_ok && x >= 0 && x < W;
    const int y0 = band * band_rows;
    const int rows = min(band_rows, H - y0);              // output rows of this band (uniform over the image)
    const float* xrow0 = xin + (size_t)n * H * W * C;
    float* yrow0 = yout + (size_t)n * H * W * C;

    // ---- S_a: LayerNorm(x[row]) -> A1 (hi/lo fp16), spread over all warps -------------------------------------------
    // Thread (pixel row, cg) owns channels [8cg, 8cg+8): the global load is issued one row ahead (prefetch_x), the
    // statistics of the G slices are combined with Chan's formula through shared memory (publish_stats), and each
    // thread normalises / splits / stores only its own K-chunk of A1 (stage_a).
    float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xb = xa;
    bool xvalid = false;
    auto prefetch_x = [&](int it) {
      const int y = y0 + it - 1;
      xvalid = x_ok && y >= 0 && y < H && it < rows + 2;
      if (xvalid && cg < LG) {
        const float4* src = reinterpret_cast<const float4*>(xrow0 + ((size_t)y * W + x) * C + cg * 8);
        xa = __ldg(src);
        xb = __ldg(src + 1);
      }
    };
    auto publish_stats = [&]() {
      if (cg >= LG) return;
      const float mean = (((xa.x + xa.y) + (xa.z + xa.w)) + ((xb.x + xb.y) + (xb.z + xb.w))) * 0.125f;
      float m2 = 0.f, d;
      d = xa.x - mean; m2 = fmaf(d, d, m2); d = xa.y - mean; m2 = fmaf(d, d, m2);
      d = xa.z - mean; m2 = fmaf(d, d, m2); d = xa.w - mean; m2 = fmaf(d, d, m2);
      d = xb.x - mean; m2 = fmaf(d, d, m2); d = xb.y - mean; m2 = fmaf(d, d, m2);
      d = xb.z - mean; m2 = fmaf(d, d, m2); d = xb.w - mean; m2 = fmaf(d, d, m2);
      sm.part[cg][row] = make_float2(mean, m2);
    };
    const uint4 kFlagOn = make_uint4(0x00003C00u, 0u, 0u, 0u);   // fp16 {1, 0, 0, 0, 0, 0, 0, 0}
    const uint4 kFlagOff = make_uint4(0u, 0u, 0u, 0u);
    auto stage_a = [&]() {
      if (cg >= LG) return;
      if (cg == 0) *reinterpret_cast<uint4*>(&sm.a1h[((C / 8) * 128 + row) * 8]) = xvalid ? kFlagOn : kFlagOff;
      float2 t8[4];
      if (xvalid) {
        float2 pr[LG];
        float mean = 0.f;
#pragma unroll
        for (int g = 0; g < LG; ++g) { pr[g] = sm.part[g][row]; mean += pr[g].x; }
        mean *= (1.0f / LG);
        float m2 = 0.f;
#pragma unroll
        for (int g = 0; g < LG; ++g) { const float dm = pr[g].x - mean; m2 += fmaf(8.0f * dm, dm, pr[g].y); }
        const float rstd = 1.0f / sqrtf(m2 * (1.0f / C) + kLnEps);
        const float4 ga = *reinterpret_cast<const float4*>(&sm.lng[cg * 8]), gb = *reinterpret_cast<const float4*>(&sm.lng[cg * 8 + 4]);
        const float4 ba = *reinterpret_cast<const float4*>(&sm.lnb[cg * 8]), bb = *reinterpret_cast<const float4*>(&sm.lnb[cg * 8 + 4]);
        t8[0] = make_float2((xa.x - mean) * rstd * ga.x + ba.x, (xa.y - mean) * rstd * ga.y + ba.y);
        t8[1] = make_float2((xa.z - mean) * rstd * ga.z + ba.z, (xa.w - mean) * rstd * ga.w + ba.w);
        t8[2] = make_float2((xb.x - mean) * rstd * gb.x + bb.x, (xb.y - mean) * rstd * gb.y + bb.y);
        t8[3] = make_float2((xb.z - mean) * rstd * gb.z + bb.z, (xb.w - mean) * rstd * gb.w + bb.w);
      } else {
        t8[0] = t8[1] = t8[2] = t8[3] = make_float2(0.f, 0.f);
      }
      uint4 hi, lo;
      split8(t8, hi, lo);
      *reinterpret_cast<uint4*>(&sm.a1h[(cg * 128 + row) * 8]) = hi;
      *reinterpret_cast<uint4*>(&sm.a1l[(cg * 128 + row) * 8]) = lo;
    };
    // ---- S_d: y = D3 + b2 + x for the output row produced by iteration `itr` (interior lanes only) -----------------------
    __half* const a3h_p = kOwnA3 ? sm.a3h : sm.a2h;
    __half* const a3l_p = kOwnA3 ? sm.a3l : sm.a2l;
    auto row_store_ok = [&](int itr) { return cg < LG && x_ok && lane >= 1 && lane <= kStripW && itr >= 2 && itr - 2 < rows; };
    auto load_residual = [&](int itr, float4& r0, float4& r1) {
      r0 = make_float4(0.f, 0.f, 0.f, 0.f);
      r1 = r0;
      if (row_store_ok(itr)) {
        const float4* src = reinterpret_cast<const float4*>(xrow0 + ((size_t)(y0 + itr - 2) * W + x) * C + cg * 8);
        r0 = __ldg(src);
        r1 = __ldg(src + 1);
      }
    };
    auto store_row = [&](int itr, const float4& r0, const float4& r1) {
      const uint32_t d3 = D2_COL + (uint32_t)((itr + 1) % 3) * C4;   // GEMM3's accumulator: the hidden-row slot that died
      float2 v2[4];
      tmem_ld8(lane_addr + d3 + (cg < LG ? cg : 0) * 8, v2);
      tmem_ld_wait();
      if (row_store_ok(itr)) {
        float* dst = yrow0 + ((size_t)(y0 + itr - 2) * W + x) * C + cg * 8;
        const float4 ba = *reinterpret_cast<const float4*>(&sm.b2[cg * 8]);
        const float4 bb = *reinterpret_cast<const float4*>(&sm.b2[cg * 8 + 4]);
        *reinterpret_cast<float4*>(dst) = make_float4((v2[0].x + ba.x) + r0.x, (v2[0].y + ba.y) + r0.y,
                                                      (v2[1].x + ba.z) + r0.z, (v2[1].y + ba.w) + r0.w);
        *reinterpret_cast<float4*>(dst + 4) = make_float4((v2[2].x + bb.x) + r1.x, (v2[2].y + bb.y) + r1.y,
                                                          (v2[3].x + bb.z) + r1.z, (v2[3].y + bb.w) + r1.w);
      }
    };

    prefetch_x(0);
    publish_stats();
    epi_sync();
    stage_a();
    signal(&sm.ready[0]);

    for (int it = 0; it < iters; ++it) {
      const bool row_valid = xvalid;                      // this row's pixel lies inside the image
      if (it + 1 < iters) prefetch_x(it + 1);             // global load of the next row flies under the wait + S_b
      float4 pr0, pr1;                                    // residual of the previous output row (stored after S_b)
      if constexpr (kOwnA3) load_residual(it - 1, pr0, pr1);
      // ---- S_b: GELU(D1) -> A2 (hi/lo fp16); D1 already holds the bias, and is exactly 0 outside the image ------------
      if (exp_mode != 1) mbar_wait(&sm.mbar[0], ph1);
      ph1 ^= 1;
      tc_fence_after();
      {
        // all CH accumulator columns of this thread in one go (one wait), then CH/8 independent chunks the
        // scheduler can interleave: the epilogue is latency-bound, not issue-bound
        float2 acc[CH / 8][4];
#pragma unroll
        for (int c = 0; c < CH / 8; ++c) tmem_ld8(lane_addr + D1_COL + cg * CH + 8 * c, acc[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < CH / 8; ++c) {
          const int c0 = cg * CH + 8 * c;
          float2 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = gelu_pair(acc[c][i]);
          uint4 hi, lo;
          split8(v, hi, lo);
          *reinterpret_cast<uint4*>(&sm.a2h[((c0 >> 3) * 128 + row) * 8]) = hi;
          *reinterpret_cast<uint4*>(&sm.a2l[((c0 >> 3) * 128 + row) * 8]) = lo;
        }
        if (cg == 0) *reinterpret_cast<uint4*>(&sm.a2h[((C4 / 8) * 128 + row) * 8]) = row_valid ? kFlagOn : kFlagOff;
      }
      if constexpr (kOwnA3) {
        // the previous row's GEMM3 ran under this S_b; its accumulator sits in the slot GEMM2 of this row overwrites
        if (it >= 3) {
          if (exp_mode != 1) mbar_wait(&sm.mbar[2], ph3);
          ph3 ^= 1;
          tc_fence_after();
          store_row(it - 1, pr0, pr1);
        }
      }
      if (it + 1 < iters) publish_stats();
      // ---- G2 runs under S_a of the next row ------------------------------------------------------------------------------
      signal(&sm.ready[1]);
      // ---- S_a + G1 of the next row: G1 runs under the depthwise stage below -------------------------------------------
      if (it + 1 < iters) {
        epi_sync();
        stage_a();
        signal(&sm.ready[0]);
      }
      if (exp_mode != 1) mbar_wait(&sm.mbar[1], ph2);
      ph2 ^= 1;
      tc_fence_after();
      if (it < 2) continue;                               // uniform over the CTA
      // ---- S_c: depthwise 3x3 over hidden rows (it-2, it-1, it) + bias -> GELU -> A3 -------------------------------
      // zero padding of the conv: hidden rows / columns outside the image are exactly 0 in TMEM (flag column of GEMM2)
      uint32_t slot[3];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) slot[dy] = lane_addr + D2_COL + (uint32_t)((it - 2 + dy) % 3) * C4;
      float2 hn[3][4];
      tmem_ld8(slot[0] + cg * CH, hn[0]);
      tmem_ld8(slot[1] + cg * CH, hn[1]);
      tmem_ld8(slot[2] + cg * CH, hn[2]);
      tmem_ld_wait();
#pragma unroll
      for (int c0 = cg * CH; c0 < (cg + 1) * CH; c0 += 8) {
        float2 h[3][4];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int i = 0; i < 4; ++i) h[dy][i] = hn[dy][i];
        if (c0 + 8 < (cg + 1) * CH) {                     // next chunk in flight while this one computes
          tmem_ld8(slot[0] + c0 + 8, hn[0]);
          tmem_ld8(slot[1] + c0 + 8, hn[1]);
          tmem_ld8(slot[2] + c0 + 8, hn[2]);
        }
        float2 cl[4], cc[4], cr[4];                       // per-lane column sums for the left / centre / right taps
        const float4 da = *reinterpret_cast<const float4*>(&sm.dwb[c0]);
        const float4 db = *reinterpret_cast<const float4*>(&sm.dwb[c0 + 4]);
        const float2 dbv[4] = {make_float2(da.x, da.y), make_float2(da.z, da.w), make_float2(db.x, db.y), make_float2(db.z, db.w)};
        // the centre row initialises the sums (the conv bias rides in the centre tap)
#pragma unroll
        for (int dyi = 0; dyi < 3; ++dyi) {
          const int dy = (dyi == 0) ? 1 : (dyi == 1 ? 0 : 2);
          const float4* wl = reinterpret_cast<const float4*>(&sm.dww[(dy * 3 + 0) * C4 + c0]);
          const float4* wc = reinterpret_cast<const float4*>(&sm.dww[(dy * 3 + 1) * C4 + c0]);
          const float4* wr = reinterpret_cast<const float4*>(&sm.dww[(dy * 3 + 2) * C4 + c0]);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const float4 l4 = wl[hf], c4 = wc[hf], r4 = wr[hf];
            const float2 ha = h[dy][2 * hf], hb = h[dy][2 * hf + 1];
            // tap (dy,-1) of THIS pixel is consumed by lane+1, tap (dy,+1) by lane-1
            if (dy == 1) {
              cl[2 * hf] = __fmul2_rn(make_float2(l4.x, l4.y), ha);
              cl[2 * hf + 1] = __fmul2_rn(make_float2(l4.z, l4.w), hb);
              cc[2 * hf] = __ffma2_rn(make_float2(c4.x, c4.y), ha, dbv[2 * hf]);
              cc[2 * hf + 1] = __ffma2_rn(make_float2(c4.z, c4.w), hb, dbv[2 * hf + 1]);
              cr[2 * hf] = __fmul2_rn(make_float2(r4.x, r4.y), ha);
              cr[2 * hf + 1] = __fmul2_rn(make_float2(r4.z, r4.w), hb);
            } else {
              cl[2 * hf] = __ffma2_rn(make_float2(l4.x, l4.y), ha, cl[2 * hf]);
              cl[2 * hf + 1] = __ffma2_rn(make_float2(l4.z, l4.w), hb, cl[2 * hf + 1]);
              cc[2 * hf] = __ffma2_rn(make_float2(c4.x, c4.y), ha, cc[2 * hf]);
              cc[2 * hf + 1] = __ffma2_rn(make_float2(c4.z, c4.w), hb, cc[2 * hf + 1]);
              cr[2 * hf] = __ffma2_rn(make_float2(r4.x, r4.y), ha, cr[2 * hf]);
              cr[2 * hf + 1] = __ffma2_rn(make_float2(r4.z, r4.w), hb, cr[2 * hf + 1]);
            }
          }
        }
        float2 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 left, right;
          left.x = __shfl_up_sync(0xffffffffu, cl[i].x, 1);        // from pixel x-1
          left.y = __shfl_up_sync(0xffffffffu, cl[i].y, 1);
          right.x = __shfl_down_sync(0xffffffffu, cr[i].x, 1);     // from pixel x+1
          right.y = __shfl_down_sync(0xffffffffu, cr[i].y, 1);
          o[i] = gelu_pair(__fadd2_rn(__fadd2_rn(cc[i], left), right));
        }
        uint4 hi, lo;
        split8(o, hi, lo);
        *reinterpret_cast<uint4*>(&a3h_p[((c0 >> 3) * 128 + row) * 8]) = hi;
        *reinterpret_cast<uint4*>(&a3l_p[((c0 >> 3) * 128 + row) * 8]) = lo;
        tmem_ld_wait();
      }
      signal(&sm.ready[2]);                               // -> G3(it)
      if constexpr (!kOwnA3) {
        // A3 shares A2's buffer: GEMM3 has to finish before the next row's S_b may overwrite it
        float4 r0, r1;
        load_residual(it, r0, r1);
        mbar_wait(&sm.mbar[2], ph3);
        ph3 ^= 1;
        tc_fence_after();
        store_row(it, r0, r1);
      }
    }
    if constexpr (kOwnA3) {                               // drain: the last row's GEMM3
      float4 r0, r1;
      load_residual(iters - 1, r0, r1);
      mbar_wait(&sm.mbar[2], ph3);
      ph3 ^= 1;
      tc_fence_after();
      store_row(iters - 1, r0, r1);
    }
  }
  }   // role
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// halves of the packed operand block: the fused kernel (c = 16, 32) carries the bias K-steps of GEMM1 / GEMM2, the
// wide path (c = 64, pwgemm_tc.cu) the plain matrices
size_t ffn_tc_pack_halves(int c) {
  if (c == 64) return (size_t)2 * (4 * c * c + 16 * c * c + 4 * c * c);
  return (size_t)2 * (4 * c * (c + 8) + 4 * c * (4 * c + 8) + 4 * c * c);
}

template <int C, int G>
static cudaError_t ffn_tc_t(const BlockW& w, const float* x, float* y, int N, int H, int W, cudaStream_t s) {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  const int per_sm = (C == 16) ? 2 : 1;
  const int nws = (W + kStripW - 1) / kStripW;
  // taller bands amortise the two halo rows (130/128, 66/64, 34/32 ...) but leave fewer groups to spread over the
  // resident CTAs: pick the height with the best product of halo efficiency and last-wave fill
  int band_rows = 8;
  double best = 0.0;
  for (int r = 128; r >= 8; r >>= 1) {
    if (r > H) continue;
    const long long g = ((long long)N * ((H + r - 1) / r) * nws + 3) / 4;
    const long long cap = (long long)sm_count * per_sm;
    const long long waves = (g + cap - 1) / cap;
    const double score = ((double)r / (r + 2)) * ((double)g / (double)(waves * cap));
    if (score > best * 1.005) { best = score; band_rows = r; }
  }
  const int nbands = (H + band_rows - 1) / band_rows;
  const int units = N * nbands * nws;
  const int groups = (units + 3) / 4;
  const size_t smem = sizeof(FfnTcSmem<C>) + 128;
  cudaError_t e = cudaFuncSetAttribute(ffn_tc_kernel<C, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int grid = groups < sm_count * per_sm ? groups : sm_count * per_sm;
  static const int exp_mode = [] { const char* e = getenv("LGTEUN_FFN_EXP"); return e ? atoi(e) : 0; }();
  ffn_tc_kernel<C, G><<<grid, 128 * G + 32, smem, s>>>(x, y, w, reinterpret_cast<const __half*>(w.ffn_pack), H, W, nws, nbands,
                                                   band_rows, units, groups, exp_mode);
  return cudaGetLastError();
}

cudaError_t launch_ffn_tc(const BlockW& w, int c, const float* x, float* y, int N, int H, int W, cudaStream_t s) {
  switch (c) {
    case 16: return ffn_tc_t<16, kG16>(w, x, y, N, H, W, s);
    case 32: return ffn_tc_t<32, 4>(w, x, y, N, H, W, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace lg
