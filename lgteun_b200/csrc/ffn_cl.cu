// ffn_cl.cu — residual(pre_norm(feed_forward)) as one fused tcgen05/TMEM kernel, "channels on lanes" form (sm_100a).
//   reference: models/common/LGT.py:91-109 (feed_forward), :45-61 (residual/pre_norm),
//              models/common/basic_module_unformer_v2.py:37-53 (depthwise_conv = 1x1 then dw3x3, zero pad)
//     y = x + W2 . GELU( dw3x3( W1 . GELU( W0 . LN(x) + b0 ) + b1 ) + bdw ) + b2
//
// Same arithmetic as ffn_tc.cu (fp16 hi/lo split operands, hi*hi + hi*lo + lo*hi into fp32 TMEM accumulators, exact-form
// GELU), different mapping.  ffn_tc.cu puts PIXELS on the 128 TMEM lanes and the hidden channels in TMEM columns; its
// epilogue then needs the depthwise weights from shared memory (23 % of the instructions were LDS/SHFL, the L1 data pipe
// ran at 76 % of its peak) and a LayerNorm split over two threads per pixel.  Here the first two GEMMs are transposed:
//   * TMEM lane = HIDDEN CHANNEL, TMEM column = pixel of an image-row segment:  D1/D2[ch][px] = W[ch][k] . Act[px][k]^T.
//     The weights are the A operand (M = 4c = 64 or 128), the activations the B operand (N = 40 / 48 pixels).
//   * An epilogue thread owns ONE hidden channel: its nine depthwise taps and biases are registers, the horizontal taps
//     are neighbouring TMEM columns of its own lane (three shifted tcgen05.ld windows, no shuffles), the vertical taps
//     are three accumulator slots (hidden rows y-1, y, y+1) as before.  GELU outputs of 8 consecutive pixels are one
//     16-byte store into an MN-major (pixel-contiguous) operand tile, which GEMM2 reads as B and GEMM3 as A.
//   * GEMM3 goes back to pixels-on-lanes (M = 128 pixel columns, N = c), so the output row is stored by one warp with
//     thread = pixel, exactly as the NHWC map wants it.
//   * c = 16: the hidden layer has 64 channels, so two independent row segments (halves) share the 128 lanes through
//     the M = 64 "half sub-partition" accumulator layout (rows 16i..16i+15 -> lanes 32i..32i+15, second MMA at lane
//     offset 16): no block-diagonal weights, both halves use the same 28 KB of packed weights, two CTAs per SM.
//     c = 32: 128 hidden channels = the 128 lanes, two independent streams (own operands, barriers, 256 TMEM columns)
//     inside one CTA share the 104 KB of weights.
//   * Warp roles per stream: 8 epilogue warps (lane quarter x column half), 3 loader warps (coalesced 16-byte loads of
//     x, LayerNorm across the 4 / 8 lanes of a pixel, hi/lo split into the K-major B operand of GEMM1, validity flags),
//     one of which also stores the output rows (D3 + b2 + x), and one MMA-issuer thread.  Biases of GEMM1 / GEMM2 and the
//     zero padding of the conv ride in one extra K-step (bias_hi, bias_lo) x (valid, valid).
//   * Weights arrive with cp.async.bulk (TMA bulk copy) on an mbarrier.
// Probed on the hardware before the kernel was written (tools/tcprobe.cu): M = 64 with the accumulator at lane 0 / 16,
// MN-major B and A operands without swizzle, tcgen05.ld at unaligned column offsets, tcgen05.ld throughput.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace lg {

using namespace tc;

namespace cl {

constexpr int kWarpsPerStream = 12;

// Geometry of a row segment.  c = 16: 32 interior pixels (+ 2 halo) in 40 columns, M = 64 MMAs of N = 40.  c = 32: M = 128 MMAs need
// N = 48 anyway and cost the same ~44 clk for any N <= 64 (tools/tcprobe), so a segment carries 46 interior pixels.
template <int C>
struct Geo {
  static constexpr int KIN = (C == 16) ? 32 : 46;        // interior pixels of a row segment (output columns per half and iteration)
  static constexpr int NPS = (C == 16) ? 40 : 48;        // pixel columns of every operand tile and of a TMEM slot (= MMA N)
  static constexpr int CPW = NPS / 16;                   // 8-pixel chunks per epilogue warp and phase (two column halves)
  static constexpr bool EXTRA = (C == 16);               // columns 32, 33 (the halo side) are a separate 2-column piece
  static constexpr int B1S = NPS * 16 + 96;              // bytes between the 8-channel chunks of the K-major GEMM1 operand (+96: bank spread)
  static constexpr int MNK = (NPS / 8) * 128;            // bytes between 8-channel blocks of an MN-major tile
  static constexpr int A3B = (KIN + 7) / 8;              // 8-pixel blocks of one half in the A3 tile (interior pixels only)
  static constexpr int NB1 = (C == 16) ? 2 : 1;          // buffers of the GEMM1 operand (c = 32: shared memory is full, see Stream)
};

__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float2& v) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
  v = make_float2(__uint_as_float(r0), __uint_as_float(r1));
}
// shared-memory stores through 32-bit shared addresses (one STS.128 each; the generic-pointer form was split by the compiler)
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// packed weights (fp16), global and shared: K-major core-matrix tiles [K/8][rows][8]
//   w0h | w0l [C/8][C4][8]   w0b [2][C4][8] (col 0 = b0 hi, col 1 = b0 lo)
//   w1h | w1l [C4/8][C4][8]  w1b [2][C4][8]
//   w2      [C4/8][2C][8]   rows 0..C-1 = hi, C..2C-1 = lo (GEMM3 issues a . [hi;lo] and a_lo . hi: two MMAs per K-step)
template <int C>
struct Pack {
  static constexpr int C4 = 4 * C;
  static constexpr int w0 = C4 * C, wb = C4 * 16, w1 = C4 * C4, w2 = C * C4;
  static constexpr int o_w0h = 0, o_w0l = o_w0h + w0, o_w0b = o_w0l + w0, o_w1h = o_w0b + wb, o_w1l = o_w1h + w1,
                       o_w1b = o_w1l + w1, o_w2 = o_w1b + wb, halves = o_w2 + 2 * w2;
};

template <int C>
struct Stream {
  using G = Geo<C>;
  static constexpr int C4 = 4 * C;
  static constexpr int HV = 128 / C4;                         // row segments sharing the 128 lanes
  uint64_t g1, g2, g3[2], empty[2], ready_b1[2], ready_a2, ready_a3, d3_free[2];   // c = 32 uses g3[0] / d3_free[0] / ready_b1[0] only
  // GEMM1 B operand, K-major [C/8][NPS px][8], per buffer and half.  c = 32 has ONE buffer: the loaders rewrite it as soon as
  // GEMM1 of the previous row has completed (an iteration before it is needed), which frees 8 KB per stream
  alignas(128) unsigned char b1h[G::NB1][HV][(C / 8) * G::B1S];
  unsigned char b1l[G::NB1][HV][(C / 8) * G::B1S];
  unsigned char b1f[G::NB1][HV][2 * G::B1S];                  // flag K-step: chunk 0 cols 0,1 = valid, chunk 1 = 0
  unsigned char a2f[2][HV][2 * G::MNK];                       // flag K-step of GEMM2, MN-major: rows 0,1 = valid, rest 0
  unsigned char a2h[HV][C4 / 8 * G::MNK], a2l[HV][C4 / 8 * G::MNK];   // GELU(D1): GEMM2 B operand, MN-major [C4/8][NPS/8][8 ch][8 px]
  // GELU(dw): GEMM3 A operand (M = pixels), MN-major [C4/8][HV * A3B pixel blocks][8 ch][8 px]: the interior pixels of every half
  // side by side, so ONE M = 128 MMA covers all halves (c = 16: rows 0..31 = half 0, 32..63 = half 1; the rest aliases, unused)
  unsigned char a3h[C4 / 8 * HV * G::A3B * 128], a3l[C4 / 8 * HV * G::A3B * 128];
  unsigned char tail[1536];                                   // GEMM3 reads 16 pixel blocks per channel block: it runs <= 1280 B over
};

template <int C, int NS>
struct Smem {
  uint64_t wbar;
  uint32_t tmem_base;
  uint32_t pad_;
  alignas(16) float b2[C], lng[C], lnb[C];
  alignas(128) __half w[Pack<C>::halves];
  Stream<C> st[NS];
};

static_assert(sizeof(Smem<16, 1>) + 128 + 1024 <= 114 * 1024, "c = 16: two CTAs per SM");
static_assert(sizeof(Smem<32, 2>) + 128 <= 227 * 1024, "c = 32: one CTA per SM, two streams");

// rows = rows of one K-chunk of the destination tiles (N, or 2N when hi and lo share a tile: lo = hi + N * 8 halves)
__global__ void pack_kernel(const float* __restrict__ w, const float* __restrict__ bias, __half* __restrict__ hi,
                            __half* __restrict__ lo, __half* __restrict__ bb, int N, int K, int rows) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx < N * K) {
    const int n = idx / K, k = idx - n * K;
    const float v = w[idx];
    const __half h = __float2half_rn(v);
    const size_t o = ((size_t)(k >> 3) * rows + n) * 8 + (k & 7);
    hi[o] = h;
    lo[o] = __float2half_rn(v - __half2float(h));
  }
  if (bb && idx < N * 16) {
    const int n = idx >> 4, k = idx & 15;
    const float v = bias[n];
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    bb[((size_t)(k >> 3) * N + n) * 8 + (k & 7)] = k == 0 ? h : k == 1 ? l : __float2half_rn(0.f);
  }
}

template <int C, int NS>
__global__ void __launch_bounds__(NS * kWarpsPerStream * 32, NS == 1 ? 2 : 1)
ffn_cl_kernel(const float* __restrict__ xin, float* __restrict__ yout, BlockW w, const __half* __restrict__ wpack, int H, int W,
              int nws, int nbands, int band_rows, int total_units, int num_groups) {
  constexpr int C4 = 4 * C;
  constexpr int HV = 128 / C4;
  const bool in_al32 = (reinterpret_cast<uintptr_t>(xin) & 31) == 0;   // 32-byte loads of a pixel row (LDG.256)
  constexpr int NP = (C4 == 64) ? 40 : 48;           // MMA N of GEMM1 / GEMM2 (M = 128 needs a multiple of 16; columns >= 40 alias, unused)
  // Order of the epilogue phases.  LAG = 0: S_b(row) -> wait GEMM2(row) -> S_c(row).  LAG = 1 (c = 16): S_b(row) -> S_c(row - 1) -> wait
  // GEMM2(row): the latency of GEMM2 / GEMM1 hides under the depthwise stage of the previous row; needs a fourth hidden-row slot.
  constexpr int LAG = (C == 16) ? 1 : 0;
  constexpr int NSLOT = 3 + LAG;
  constexpr uint32_t SLOT = (C == 16) ? 40 : 48;     // TMEM columns of D1 and of one hidden-row slot (= MMA N)
  constexpr uint32_t D1_COL = 0, D2_COL = SLOT, D3_COL = SLOT * (1 + NSLOT);   // per stream: D1 | hidden-row slots | output accumulators
  // GEMM3 output: c = 16: two accumulators of 16 columns (hi.hi + hi.lo + lo.hi as three MMAs per K-step);
  //               c = 32: one of 64 columns ([0,32) + [32,64) = the row: a . [w_hi ; w_lo] and a_lo . w_hi, two MMAs per K-step)
  constexpr int NB3 = (C == 16) ? 2 : 1;
  constexpr bool CONCAT = (C == 32);
  constexpr uint32_t D3W = CONCAT ? 2 * C : C;
  static_assert(D3_COL + NB3 * D3W <= 256, "TMEM columns of one stream");
  using G = Geo<C>;
  constexpr int kIn = G::KIN, kB1Stride = G::B1S, kMnK = G::MNK, CPW = G::CPW, NB1 = G::NB1;
  constexpr uint32_t A3K = HV * G::A3B * 128;        // bytes between 8-channel blocks of the A3 tile
  using P = Pack<C>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem<C, NS>& sm = *reinterpret_cast<Smem<C, NS>*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = warp / kWarpsPerStream, lw = warp % kWarpsPerStream;
  Stream<C>& st = sm.st[s];

  // ---- one-time setup ---------------------------------------------------------------------------------------------------
  if (tid == 0) {
    mbar_init(&sm.wbar, 1);
    for (int i = 0; i < NS; ++i) {
      Stream<C>& t = sm.st[i];
      mbar_init(&t.g1, 1);
      mbar_init(&t.g2, 1);
      mbar_init(&t.ready_a2, 8);
      mbar_init(&t.ready_a3, 8);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&t.g3[b], 1);
        mbar_init(&t.empty[b], 1);
        mbar_init(&t.ready_b1[b], 3);
        mbar_init(&t.d3_free[b], 2);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    constexpr uint32_t bytes = P::halves * 2;
    mbar_expect_tx(&sm.wbar, bytes);
    constexpr uint32_t piece = 16384;
    for (uint32_t o = 0; o < bytes; o += piece)
      bulk_g2s(reinterpret_cast<unsigned char*>(sm.w) + o, reinterpret_cast<const unsigned char*>(wpack) + o,
               bytes - o < piece ? bytes - o : piece, &sm.wbar);
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 256 * NS);
  for (int i = tid; i < C; i += blockDim.x) {
    sm.b2[i] = __ldg(w.f2_b + i);
    sm.lng[i] = __ldg(w.ln2_w + i);
    sm.lnb[i] = __ldg(w.ln2_b + i);
  }
  {
    // the flag K-steps are zero except for the valid entries the loaders rewrite every row; clear the operand tiles too so that the
    // never-written pixel columns (34 .. 39) hold finite values
    uint4* z = reinterpret_cast<uint4*>(st.b1h);
    const int n16 = (int)((sizeof(Stream<C>) - offsetof(Stream<C>, b1h)) / 16);
    for (int i = lw * 32 + lane; i < n16; i += kWarpsPerStream * 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base + 256u * s;
  const int iters = min(band_rows, H) + 2;
  const int group0 = blockIdx.x * NS + s, gstep = gridDim.x * NS;

  if (lw == 10) {
    // ---- MMA issuer: the whole warp runs the loop, one elected lane issues (under elect.sync ptxas knows a single thread is
    // active and moves the operands to uniform registers without a divergence loop: ~3 instructions per MMA instead of ~10) ----------
    mbar_wait(&sm.wbar, 0);
    // descriptor = base (address >> 4, computed once) + compile-time offsets; the upper word is a per-layout constant
    const uint32_t wb = smem_u32(sm.w) >> 4, sb = smem_u32(&st) >> 4;
    constexpr uint32_t HI = (128u >> 4) | (1u << 14);                     // SBO = 128 B, descriptor version 1
    auto dsc = [&](uint32_t base16, uint32_t byte_off, uint32_t lbo) -> uint64_t {
      return (uint64_t)(base16 + (byte_off >> 4) + ((lbo >> 4) << 16)) | ((uint64_t)HI << 32);
    };
    constexpr uint32_t WK = C4 * 16;                                       // K-chunk stride of the weight tiles with C4 rows
    constexpr uint32_t o_b1h = offsetof(Stream<C>, b1h), o_b1l = offsetof(Stream<C>, b1l), o_b1f = offsetof(Stream<C>, b1f);
    constexpr uint32_t o_a2f = offsetof(Stream<C>, a2f), o_a2h = offsetof(Stream<C>, a2h), o_a2l = offsetof(Stream<C>, a2l);
    constexpr uint32_t o_a3h = offsetof(Stream<C>, a3h), o_a3l = offsetof(Stream<C>, a3l);
    constexpr uint32_t B1SZ = (C / 8) * kB1Stride, A2SZ = C4 / 8 * kMnK;   // bytes per (buffer, half) / per half
    uint32_t gi = 0, j3 = 0;          // running row index (B1 / flag buffer = gi & 1, hidden slot = gi % 3) and GEMM3 index (D3 buffer = j3 & 1)
    uint32_t ph_a2 = 0, ph_a3 = 0;
    auto issue_g1 = [&](uint32_t g) {
      const uint32_t b = NB1 == 2 ? (g & 1) : 0;
      mbar_wait(&st.ready_b1[b], NB1 == 2 ? (g >> 1) & 1 : g & 1);
      tc_fence_after();
      if (elect_one()) {
        constexpr uint32_t id = idesc(C4, NP, 0, 0);
        const uint32_t sbb = sb + ((b * HV * B1SZ) >> 4), sbf = sb + ((b * HV * 2 * kB1Stride) >> 4);
#pragma unroll
        for (int h = 0; h < HV; ++h) {
          const uint32_t d = tmem + D1_COL + ((uint32_t)(16 * h) << 16);
#pragma unroll
          for (int ks = 0; ks < C / 16; ++ks) {
            const uint64_t ah = dsc(wb, P::o_w0h * 2 + ks * 2 * WK, WK), al = dsc(wb, P::o_w0l * 2 + ks * 2 * WK, WK);
            const uint64_t xh = dsc(sbb, o_b1h + h * B1SZ + ks * 2 * kB1Stride, kB1Stride);
            const uint64_t xl = dsc(sbb, o_b1l + h * B1SZ + ks * 2 * kB1Stride, kB1Stride);
            umma_f16(d, ah, xh, id, ks > 0);
            umma_f16(d, ah, xl, id, 1);
            umma_f16(d, al, xh, id, 1);
          }
          umma_f16(d, dsc(wb, P::o_w0b * 2, WK), dsc(sbf, o_b1f + h * 2 * kB1Stride, kB1Stride), id, 1);
        }
        umma_commit(&st.g1);
        if (NB1 == 1) umma_commit(&st.empty[0]);      // single B1 buffer: free again once this GEMM1 (and everything before it) is done
      }
      __syncwarp();
    };
    auto issue_g2 = [&](uint32_t g) {
      if (elect_one()) {
        constexpr uint32_t id = idesc(C4, NP, 0, 1);
        const uint32_t b = g & 1;
        const uint32_t sbf = sb + ((b * HV * 2 * kMnK) >> 4);
        const uint32_t dcol = tmem + D2_COL + (g % NSLOT) * SLOT;
#pragma unroll
        for (int h = 0; h < HV; ++h) {
          const uint32_t d = dcol + ((uint32_t)(16 * h) << 16);
#pragma unroll
          for (int ks = 0; ks < C4 / 16; ++ks) {
            const uint64_t ah = dsc(wb, P::o_w1h * 2 + ks * 2 * WK, WK), al = dsc(wb, P::o_w1l * 2 + ks * 2 * WK, WK);
            const uint64_t xh = dsc(sb, o_a2h + h * A2SZ + ks * 2 * kMnK, kMnK), xl = dsc(sb, o_a2l + h * A2SZ + ks * 2 * kMnK, kMnK);
            umma_f16(d, ah, xh, id, ks > 0);
            umma_f16(d, ah, xl, id, 1);
            umma_f16(d, al, xh, id, 1);
          }
          umma_f16(d, dsc(wb, P::o_w1b * 2, WK), dsc(sbf, o_a2f + h * 2 * kMnK, kMnK), id, 1);
        }
        umma_commit(&st.g2);
        if (NB1 == 2) umma_commit(&st.empty[b]);      // B1 / flag buffers of this row may be rewritten
      }
      __syncwarp();
    };
    auto issue_g3 = [&](uint32_t j) {
      const uint32_t b = NB3 == 2 ? (j & 1) : 0;
      if (j >= NB3) {                   // the output warps have read the accumulator this GEMM overwrites
        mbar_wait(&st.d3_free[b], NB3 == 2 ? ((j - 2) >> 1) & 1 : (j - 1) & 1);
        tc_fence_after();
      }
      if (elect_one()) {
        const uint32_t d = tmem + D3_COL + b * D3W;
#pragma unroll
        for (int ks = 0; ks < C4 / 16; ++ks) {
          const uint64_t ah = dsc(sb, o_a3h + ks * 2 * A3K, A3K), al = dsc(sb, o_a3l + ks * 2 * A3K, A3K);
          const uint64_t bw = dsc(wb, P::o_w2 * 2 + ks * 2 * (2 * C * 16), 2 * C * 16);
          if (CONCAT) {
            umma_f16(d, ah, bw, idesc(128, 2 * C, 1, 0), ks > 0);      // [a_hi . w_hi | a_hi . w_lo]
            umma_f16(d, al, bw, idesc(128, C, 1, 0), 1);               //  a_lo . w_hi
          } else {
            const uint64_t bl = dsc(wb, P::o_w2 * 2 + ks * 2 * (2 * C * 16) + C * 16, 2 * C * 16);   // rows C .. 2C-1 of the tile
            umma_f16(d, ah, bw, idesc(128, C, 1, 0), ks > 0);
            umma_f16(d, ah, bl, idesc(128, C, 1, 0), 1);
            umma_f16(d, al, bw, idesc(128, C, 1, 0), 1);
          }
        }
        umma_commit(&st.g3[b]);
      }
      __syncwarp();
    };
    for (int grp = group0; grp < num_groups; grp += gstep) {
      issue_g1(gi);
      for (int it = 0; it < iters + LAG; ++it) {
        if (it < iters) {
          mbar_wait(&st.ready_a2, ph_a2); ph_a2 ^= 1;
          tc_fence_after();
          issue_g2(gi + it);
          if (it + 1 < iters) issue_g1(gi + it + 1);
        }
        if (it - LAG >= 2) {
          mbar_wait(&st.ready_a3, ph_a3); ph_a3 ^= 1;
          tc_fence_after();
          issue_g3(j3++);
        }
      }
      gi += iters;
    }
  } else {
    if (lw >= 8) {
    // ---- service warps 8, 9, 11: loaders; warps 8 / 9 (TMEM lane quarters 0 / 1 = D3 rows of half 0 / 1) also store the output --------
    //   c = 16 (5 pixel groups per half): warp 8: half 0 groups 0-2, warp 9: half 1 groups 0-2, warp 11: groups 3-4 of both halves
    //   c = 32 (9 pixel groups): three each
    // ---- service warps 8, 9, 11: loaders (a lane owns 16 channels of ONE pixel of the row: no cross-lane reduction at c = 16, one
    // shuffle at c = 32); warps 8 / 9 (TMEM lane quarters 0 / 1 = D3 rows of half 0 / 1) also store the output rows -------------------
    //   c = 16: warp 8 = half 0 pixel columns 0..31, warp 9 = half 1 columns 0..31, warp 11 lanes 0..3 = columns 32, 33 of both halves
    //   c = 32: two lanes per pixel: warp 8 = columns 0..15, warp 9 = 16..31, warp 11 lanes 0..3 = columns 32, 33
    const int li = lw == 8 ? 0 : lw == 9 ? 1 : 2;
    constexpr int LQ = C / 16;                       // lanes per pixel
    const int chalf = lane % LQ;                     // 16-channel half of this lane
    const int h = (HV == 2) ? (li < 2 ? li : (lane >> 1) & 1) : 0;
    //   c = 16: warps 8 / 9 = pixel columns 0..31 of half 0 / 1, warp 11 lanes 0..3 = columns 32, 33 of both halves
    //   c = 32: two lanes per pixel, 48 columns (46 interior + 2 halo): warp 8 = columns 0..15, warp 9 = 16..31, warp 11 = 32..47
    const int pc = (HV == 2) ? (li < 2 ? lane : 32 + (lane & 1)) : li * 16 + lane / LQ;
    const bool lane_on = HV == 1 || li < 2 || lane < 4;
    const bool outw = li < 2;                        // warps 8 / 9 store the output rows: c = 16 of half li, c = 32 pixels 32 li .. 32 li + 31
    uint32_t gi = 0, j3 = 0;
    // Output cursor (warps 8 / 9): rows are stored OUTLAG = NB1 + 1 + LAG loader iterations after their own.  The loaders run at
    // most NB1 rows ahead of S_b (empty[] gate), so by then the row's GEMM3 has been issued (short wait, no deadlock), and the
    // lag is small enough that the issuer's wait for a free output accumulator never depends on a loader iteration that
    // itself needs a later MMA.
    constexpr int OUTLAG = NB1 + 1 + LAG;
    int o_grp = group0, o_it = 2;
    uint32_t o_gi = 2;                                // running row index of the next output row
    size_t o_off = 0;
    bool o_col = false;
    int o_rows = 0;
    auto out_geometry = [&]() {
      const int unit = o_grp * HV + ((HV == 2) ? (li & 1) : 0);
      const bool uok = unit < total_units && o_grp < num_groups;
      const int ws = uok ? unit % nws : 0, t = uok ? unit / nws : 0;
      const int un = t / nbands, uy0 = (t % nbands) * band_rows, ux0 = ws * kIn;
      const int opx = (HV == 2) ? lane : 32 * (li & 1) + lane;     // interior pixel of the segment = D3 row
      o_col = uok && opx < min(kIn, W - ux0);
      o_rows = min(band_rows, H - uy0);
      o_off = (((size_t)un * H + uy0) * W + ux0 + opx) * C;
    };
    if (outw) out_geometry();
    const bool out_al32 = (reinterpret_cast<uintptr_t>(yout) & 31) == 0;
    auto out_step = [&]() {
      const uint32_t ob = NB3 == 2 ? (j3 & 1) : 0;
      const bool ok = o_col && o_it - 2 < o_rows;
      const size_t off = o_off + (size_t)(o_it - 2) * W * C;
      const float4* src = reinterpret_cast<const float4*>(xin + off);
      float4* dst = reinterpret_cast<float4*>(yout + off);
      float4 ra[C / 16], rb[C / 16];
      mbar_wait(&st.g3[ob], NB3 == 2 ? (j3 >> 1) & 1 : j3 & 1);
      ++j3;
      tc_fence_after();
      const uint32_t t3 = tmem + D3_COL + ob * D3W + ((uint32_t)(32 * li) << 16);
#pragma unroll
      for (int i0 = 0; i0 < C / 8; i0 += C / 16) {       // two passes of 16 (c = 16: 8) channels keep the register count down
        float2 o[CONCAT ? 2 : 1][C / 16][4];
#pragma unroll
        for (int i = 0; i < C / 16; ++i) {
          if (ok) {
            if (in_al32) lg::ldg256(src + 2 * (i0 + i), ra[i], rb[i]);
            else { ra[i] = __ldg(src + 2 * (i0 + i)); rb[i] = __ldg(src + 2 * (i0 + i) + 1); }
          }
          tmem_ld8(t3 + 8 * (i0 + i), o[0][i]);
          if (CONCAT) tmem_ld8(t3 + C + 8 * (i0 + i), o[CONCAT ? 1 : 0][i]);
        }
        tmem_ld_wait();
        if (i0 + C / 16 >= C / 8) {                       // last TMEM read of this accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&st.d3_free[ob]);
        }
        if (ok) {
#pragma unroll
          for (int i = 0; i < C / 16; ++i) {
            const float4 ba = *reinterpret_cast<const float4*>(&sm.b2[8 * (i0 + i)]), bb = *reinterpret_cast<const float4*>(&sm.b2[8 * (i0 + i) + 4]);
            float2 v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = CONCAT ? __fadd2_rn(o[0][i][e], o[CONCAT ? 1 : 0][i][e]) : o[0][i][e];
            const float4 y0 = make_float4((v[0].x + ba.x) + ra[i].x, (v[0].y + ba.y) + ra[i].y, (v[1].x + ba.z) + ra[i].z, (v[1].y + ba.w) + ra[i].w);
            const float4 y1 = make_float4((v[2].x + bb.x) + rb[i].x, (v[2].y + bb.y) + rb[i].y, (v[3].x + bb.z) + rb[i].z, (v[3].y + bb.w) + rb[i].w);
            if (out_al32) {                              // one 32-byte store = one full L2 sector per thread
              asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 2 * (i0 + i)), "f"(y0.x), "f"(y0.y), "f"(y0.z),
                           "f"(y0.w), "f"(y1.x), "f"(y1.y), "f"(y1.z), "f"(y1.w) : "memory");
            } else {
              dst[2 * (i0 + i)] = y0;
              dst[2 * (i0 + i) + 1] = y1;
            }
          }
        }
      }
      ++o_gi;
      if (++o_it == iters) {
        o_it = 2;
        o_gi += 2;
        o_grp += gstep;
        out_geometry();
      }
    };
    // shared-memory addresses of this lane's pixel for buffer 0 (buffer 1: + a constant)
    const uint32_t sa_x = smem_u32(&st.b1h[0][h][(2 * chalf) * kB1Stride + pc * 16]);
    const uint32_t sa_f1 = smem_u32(&st.b1f[0][h][pc * 16]);
    const uint32_t sa_f2 = smem_u32(&st.a2f[0][h][(pc >> 3) * 128 + (pc & 7) * 2]);
    constexpr uint32_t DL = offsetof(Stream<C>, b1l) - offsetof(Stream<C>, b1h);     // b1h -> b1l
    constexpr uint32_t BX = NB1 == 2 ? HV * (C / 8) * kB1Stride : 0, BF1 = NB1 == 2 ? HV * 2 * kB1Stride : 0;   // B1 buffer 0 -> 1
    constexpr uint32_t BF2 = HV * 2 * kMnK;                                                                       // a2f buffer 0 -> 1
    const float* lnw = &sm.lng[16 * chalf];
    const float* lnb = &sm.lnb[16 * chalf];
    for (int grp = group0; grp < num_groups; grp += gstep) {
      const int unit = grp * HV + h;
      const bool uok = unit < total_units && lane_on;
      const int ws = uok ? unit % nws : 0, t = uok ? unit / nws : 0;
      const int un = t / nbands, uy0 = (t % nbands) * band_rows, ux0 = ws * kIn;
      const int uwd = uok ? min(kIn, W - ux0) : 0, urows = min(band_rows, H - uy0);
      const int x = ux0 - 1 + pc;
      const bool colok = uok && pc < uwd + 2 && x >= 0 && x < W;
      const size_t rstride = (size_t)W * C;
      const float* src = xin + (((long long)un * H + uy0 - 1) * W + x) * C + 16 * chalf;   // row it = 0 (guarded)
      for (int it = 0; it < iters; ++it, ++gi, src += rstride) {
        const uint32_t b = gi & 1;
        const int y = uy0 - 1 + it;
        const bool ok = colok && y >= 0 && y < H && it < urows + 2;
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
          if (in_al32) {
            lg::ldg256(reinterpret_cast<const float4*>(src), v[0], v[1]);
            lg::ldg256(reinterpret_cast<const float4*>(src) + 2, v[2], v[3]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
          }
        }
        if (colok && y + 2 < H && it < urows) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + 2 * rstride));
        // LayerNorm over the C channels of the pixel
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        if (LQ == 2) sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        const float mean = sum * (1.0f / C);
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
          m2 = fmaf(v[i].x, v[i].x, m2); m2 = fmaf(v[i].y, v[i].y, m2); m2 = fmaf(v[i].z, v[i].z, m2); m2 = fmaf(v[i].w, v[i].w, m2);
        }
        if (LQ == 2) m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
        float rstd;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rstd) : "f"(fmaf(m2, 1.0f / C, kLnEps)));
        if (NB1 == 2) {
          if (gi >= 2) mbar_wait(&st.empty[b], ((gi - 2) >> 1) & 1);       // GEMM2 of row gi - 2 issued and done
        } else {
          if (gi >= 1) mbar_wait(&st.empty[0], (gi - 1) & 1);              // GEMM1 of row gi - 1 done (hence GEMM2 of row gi - 2 too)
        }
        if (lane_on) {
          const uint32_t msk = ok ? 0xffffffffu : 0u;           // pixels outside the image / segment: exact zeros
          const uint32_t ax = sa_x + b * BX;
#pragma unroll
          for (int c8 = 0; c8 < 2; ++c8) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float4 g = *reinterpret_cast<const float4*>(lnw + 8 * c8 + 4 * e), bb = *reinterpret_cast<const float4*>(lnb + 8 * c8 + 4 * e);
              const float4 u = v[2 * c8 + e];
              const float2 t0 = make_float2(fmaf(u.x * rstd, g.x, bb.x), fmaf(u.y * rstd, g.y, bb.y));
              const float2 t1 = make_float2(fmaf(u.z * rstd, g.z, bb.z), fmaf(u.w * rstd, g.w, bb.w));
              const uint32_t h0 = f2h2_sat(t0), h1 = f2h2_sat(t1);
              const float2 k0 = __half22float2(*reinterpret_cast<const __half2*>(&h0)), k1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
              hi[2 * e] = h0 & msk; hi[2 * e + 1] = h1 & msk;
              lo[2 * e] = f2h2_sat(make_float2(t0.x - k0.x, t0.y - k0.y)) & msk;
              lo[2 * e + 1] = f2h2_sat(make_float2(t1.x - k1.x, t1.y - k1.y)) & msk;
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ax + c8 * kB1Stride), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ax + c8 * kB1Stride + DL), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
          }
          if (chalf == 0) {
            const uint32_t f = 0x3C003C00u & msk;                                         // fp16 {1, 1}
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(sa_f1 + b * BF1), "r"(f) : "memory");
            asm volatile("st.shared.b16 [%0], %1;" ::"r"(sa_f2 + b * BF2), "h"((unsigned short)f) : "memory");
            asm volatile("st.shared.b16 [%0], %1;" ::"r"(sa_f2 + b * BF2 + 16), "h"((unsigned short)f) : "memory");
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&st.ready_b1[NB1 == 2 ? b : 0]);
        if (outw && o_grp < num_groups && o_gi + OUTLAG <= gi) out_step();
      }
    }
    if (outw)
      while (o_grp < num_groups) out_step();
    } else {
    // ---- epilogue warps: lane quarter q, column half ch; thread = one hidden channel (of one half) ---------------------------
    const int q = lw & 3, ch = lw >> 2;
    const int h = (HV == 2) ? (lane >> 4) : 0;
    const int k = (HV == 2) ? (16 * q + (lane & 15)) : (32 * q + lane);
    const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
    const uint32_t rowoff = (uint32_t)((k >> 3) * kMnK + (k & 7) * 16);
    const uint32_t a2h = smem_u32(st.a2h[h]) + rowoff + CPW * ch * 128;    // this warp's first 8-pixel block of the row
    const uint32_t a2l = smem_u32(st.a2l[h]) + rowoff + CPW * ch * 128;
    const uint32_t rowoff3 = (uint32_t)((k >> 3) * A3K + h * (G::A3B * 128) + (k & 7) * 16);
    const uint32_t a3h = smem_u32(st.a3h) + rowoff3 + CPW * ch * 128;
    const uint32_t a3l = smem_u32(st.a3l) + rowoff3 + CPW * ch * 128;
    float2 wt[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float v = __ldg(w.dw_w + k * 9 + t);
      wt[t] = make_float2(v, v);
    }
    const float dwb = __ldg(w.dw_b + k);
    uint32_t gi = 0, j3 = 0, ph1 = 0, ph2 = 0;
    auto signal = [&](uint64_t* bar) {      // one arrival per warp: every lane publishes its writes / TMEM reads first
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    const uint32_t d1a = tl + D1_COL + 8 * CPW * ch;       // this warp's 8 CPW columns of D1 / of a hidden-row slot
    for (int grp = group0; grp < num_groups; grp += gstep) {
      for (int it = 0; it < iters + LAG; ++it) {
        if (it < iters) {
          // ---- S_b: GELU(D1) -> A2 -----------------------------------------------------------------------------------------
          mbar_wait(&st.g1, ph1); ph1 ^= 1;
          tc_fence_after();
          float2 acc[CPW][4], ex = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < CPW; ++c) tmem_ld8(d1a + 8 * c, acc[c]);
          if (G::EXTRA && ch == 1) tmem_ld2(d1a + 8 * CPW, ex);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < CPW; ++c) {
            float2 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = gelu_pair(acc[c][i]);
            uint4 hi, lo;
            split8(v, hi, lo);
            sts128(a2h + c * 128, hi);
            sts128(a2l + c * 128, lo);
          }
          if (G::EXTRA && ch == 1) {                      // the two halo-side columns 32, 33
            const float2 v = gelu_pair(ex);
            const uint32_t hh = f2h2_sat(v);
            const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hh));
            sts32(a2h + CPW * 128, hh);                                          // (ch = 1: block 4 of the row)
            sts32(a2l + CPW * 128, f2h2_sat(make_float2(v.x - back.x, v.y - back.y)));
          }
          signal(&st.ready_a2);
          if (LAG == 0) {
            mbar_wait(&st.g2, ph2); ph2 ^= 1;
            tc_fence_after();
          }
        }
        if (it - LAG >= 2) {
          // ---- S_c: depthwise 3x3 over hidden rows r-2, r-1, r (r = gi + it - LAG) + bias -> GELU -> A3 ---------------------------
          if (j3 >= 1) mbar_wait(&st.g3[NB3 == 2 ? (j3 - 1) & 1 : 0], NB3 == 2 ? ((j3 - 1) >> 1) & 1 : (j3 - 1) & 1);     // GEMM3 of the previous row has read A3
          ++j3;
          const uint32_t r2 = gi + it - LAG - 2;
          uint32_t slot[3];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) slot[dy] = d1a + (D2_COL - D1_COL) + ((r2 + dy) % NSLOT) * SLOT;
#pragma unroll
          for (int c = 0; c < CPW; ++c) {
            // output columns 8 CPW ch + 8 c .. + 7 (pixel columns + 1), inputs .. + 9
            float2 acc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = make_float2(dwb, dwb);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              float2 L[4], M[4], R3;
              tmem_ld8(slot[dy] + 8 * c, L);
              tmem_ld8(slot[dy] + 8 * c + 1, M);
              tmem_ld2(slot[dy] + 8 * c + 8, R3);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                acc[i] = __ffma2_rn(wt[dy * 3 + 0], L[i], acc[i]);
                acc[i] = __ffma2_rn(wt[dy * 3 + 1], M[i], acc[i]);
                acc[i] = __ffma2_rn(wt[dy * 3 + 2], i < 3 ? L[i + 1] : R3, acc[i]);
              }
            }
            float2 o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = gelu_pair(acc[i]);
            uint4 hi, lo;
            split8(o, hi, lo);
            sts128(a3h + c * 128, hi);
            sts128(a3l + c * 128, lo);
          }
          signal(&st.ready_a3);
        }
        if (LAG == 1 && it < iters) {                     // A2 may be rewritten, hidden row gi + it is in its slot
          mbar_wait(&st.g2, ph2); ph2 ^= 1;
          tc_fence_after();
        }
      }
      gi += iters;
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(sm.tmem_base, 256 * NS);
}

template <int C, int NS>
static cudaError_t launch_t(const BlockW& w, const float* x, float* y, int N, int H, int W, cudaStream_t s) {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  constexpr int HV = 128 / (4 * C);
  const int streams = sm_count * 2;                     // resident streams on the device (2 CTAs x 1 or 1 CTA x 2 per SM)
  constexpr int kIn = Geo<C>::KIN;
  const int nws = (W + kIn - 1) / kIn;
  int band_rows = 8;
  double best = 0.0;
  for (int r = 256; r >= 8; r >>= 1) {
    if (r > H) continue;
    const long long g = ((long long)N * ((H + r - 1) / r) * nws + HV - 1) / HV;
    const long long waves = (g + streams - 1) / streams;
    const double score = ((double)r / (r + 2)) * ((double)g / (double)(waves * streams));
    if (score > best * 1.005) { best = score; band_rows = r; }
  }
  if (band_rows > H) band_rows = H;
  const int nbands = (H + band_rows - 1) / band_rows;
  const int units = N * nbands * nws;
  const int groups = (units + HV - 1) / HV;
  const size_t smem = sizeof(Smem<C, NS>) + 128;
  cudaError_t e = cudaFuncSetAttribute(ffn_cl_kernel<C, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int ctas_needed = (groups + NS - 1) / NS, cap = sm_count * (NS == 1 ? 2 : 1);
  const int grid = ctas_needed < cap ? ctas_needed : cap;
  ffn_cl_kernel<C, NS><<<grid, NS * kWarpsPerStream * 32, smem, s>>>(x, y, w, reinterpret_cast<const __half*>(w.ffn_cl_pack), H, W, nws,
                                                                    nbands, band_rows, units, groups);
  return cudaGetLastError();
}

}  // namespace cl

size_t ffn_cl_pack_halves(int c) { return c == 16 ? cl::Pack<16>::halves : c == 32 ? cl::Pack<32>::halves : 0; }

cudaError_t launch_ffn_cl_pack(const BlockW& b, int c, void* pack, cudaStream_t s) {
  __half* p = reinterpret_cast<__half*>(pack);
  const int c4 = 4 * c;
  auto run = [&](const float* w, const float* bias, int oh, int ol, int ob, int N, int K) {
    const int n = N * K > N * 16 ? N * K : N * 16;
    cl::pack_kernel<<<(n + 255) / 256, 256, 0, s>>>(w, bias, p + oh, p + ol, bias ? p + ob : nullptr, N, K, bias ? N : 2 * N);
  };
  if (c == 16) {
    using P = cl::Pack<16>;
    run(b.f0_w, b.f0_b, P::o_w0h, P::o_w0l, P::o_w0b, c4, c);
    run(b.f1_w, b.f1_b, P::o_w1h, P::o_w1l, P::o_w1b, c4, c4);
    run(b.f2_w, nullptr, P::o_w2, P::o_w2 + c * 8, 0, c, c4);
  } else if (c == 32) {
    using P = cl::Pack<32>;
    run(b.f0_w, b.f0_b, P::o_w0h, P::o_w0l, P::o_w0b, c4, c);
    run(b.f1_w, b.f1_b, P::o_w1h, P::o_w1l, P::o_w1b, c4, c4);
    run(b.f2_w, nullptr, P::o_w2, P::o_w2 + c * 8, 0, c, c4);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_ffn_cl(const BlockW& w, int c, const float* x, float* y, int N, int H, int W, cudaStream_t s) {
  switch (c) {
    case 16: return cl::launch_t<16, 1>(w, x, y, N, H, W, s);
    case 32: return cl::launch_t<32, 2>(w, x, y, N, H, W, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace lg
