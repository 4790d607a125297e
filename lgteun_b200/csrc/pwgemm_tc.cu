// pwgemm_tc.cu — per-pixel GEMM (1x1 convolution) on the tcgen05 tensor cores with split-fp16 operands, used for the
// conv-FFN of the widest level (c = 64, the WV-3 bottleneck: hidden = 256 channels), whose weights (384 KB as fp16
// hi/lo) and hidden rows (3 x 256 TMEM columns) do not fit the fused row-streaming kernel of ffn_tc.cu.
//
//   reference: feed_forward, models/common/LGT.py:95-109;  depthwise_conv, basic_module_unformer_v2.py:37-53
//   y = x + W2 . GELU( dw3x3( W1 . GELU( W0 . LN(x) + b0 ) + b1 ) + bdw ) + b2     as four launches:
//     pwgemm<K=64,  N=256, LN prologue,  bias+GELU epilogue>      x      -> h1
//     pwgemm<K=256, N=256, plain,        bias epilogue>           h1     -> hidden
//     dwconv_gelu (CUDA cores, HBM-bound)                         hidden -> act
//     pwgemm<K=256, N=64,  plain,        bias+residual epilogue>  act, x -> y
//   The 256-channel intermediates go through HBM/L2 in fp32; at the bottleneck resolution (1/4 of the pixels) that is
//   5 x 16.8 MB per pair, ~13 us at HBM speed, against ~160 us for the CUDA-core kernels this replaces.
//
// One CTA = 128 pixels (M = 128 = TMEM lanes).  A (activations) is loaded from global memory, optionally
// LayerNorm'ed, split x = hi + lo in fp16 and written to shared memory in the no-swizzle K-major core-matrix layout;
// B (weights, pre-packed hi/lo at load time) is streamed in slices of 64 output channels; every slice issues
// hi*hi + hi*lo + lo*hi into its 64 TMEM columns (fp32 accumulate).  See ffn_tc.cu for the descriptor formats.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace lg {

namespace pw = tc;

enum { PRO_PLAIN = 0, PRO_LN = 1, PRO_GELU = 2 };          // PRO_GELU: A = GELU(x) (the conv-FFN's activations are recomputed, not stored)
enum { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RESID = 2, EPI_GATE = 3 };   // EPI_GATE: out = acc * gelu'(gate) (training backward)

// d/dx of the exact (erf) GELU: Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_grad_exact(float x) { return gelu_grad_fast(x); }   // common.cuh: 3.3e-7 absolute
constexpr int kNS = 64;                 // output channels per weight slice
constexpr int kPwThreads = 256;

// wpack: hi [K/8][N][8] then lo [K/8][N][8] (fp16), as written by launch_pack_umma_f16
template <int K, int N, int PRO, int EPI>
__global__ void __launch_bounds__(kPwThreads, 1)
pwgemm_tc_kernel(const float* __restrict__ A, float* __restrict__ Out, const __half* __restrict__ wpack,
                 const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                 const float* __restrict__ resid, long long total_px, int num_tiles, const float* __restrict__ scale_dev) {
  using namespace pw;
  constexpr int kNS = N < lg::kNS ? N : lg::kNS;                                // narrow outputs (N = 16, 32) are one slice
  static_assert(K % 16 == 0 && N % kNS == 0 && N % 16 == 0 && N <= 256, "tile shape");
  constexpr int KC = K / 8;
  // training backward: the A operand (a gradient, ~1e-7) is multiplied by a power of two before the fp16 hi/lo split and
  // the accumulator divided by it, so that it sits in fp16's normal range like the O(1) activations of the forward
  const float a_scale = scale_dev ? __ldg(scale_dev) : 1.f, inv_scale = 1.f / a_scale;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 8);
  float* sbias = reinterpret_cast<float*>(smem_raw + 16);                       // [N]
  // A tile: [KC][128 rows][8] with the chunk stride padded by 16 bytes (LBO of the descriptor): the staging below stores 8-byte
  // pieces of 16 consecutive chunks from one half-warp, and 2048-byte strides would put them all on the same banks
  constexpr int LBO = 128 * 16 + 16;
  unsigned char* ah = smem_raw + 16 + N * 4 + ((16 - (N * 4) % 16) % 16);      // [KC][LBO]
  unsigned char* al = ah + KC * LBO;
  __half* wh = reinterpret_cast<__half*>(al + KC * LBO);                        // [KC][kNS][8]
  __half* wl = wh + kNS * K;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;                                     // TMEM lane quarter / column half
  const int row = q * 32 + lane;
  constexpr uint32_t TCOLS = (N <= 32) ? 32 : (N <= 64) ? 64 : (N <= 128 ? 128 : 256);

  if (tid == 0) {
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
  for (int i = tid; i < N; i += kPwThreads) sbias[i] = bias ? __ldg(bias + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
  const uint32_t a_h = smem_u32(ah), a_l = smem_u32(al), w_h = smem_u32(wh), w_l = smem_u32(wl);
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const long long p = (long long)tile * 128 + row;
    const bool live = p < total_px;
    // ---- prologue: the tile is ONE contiguous block of 128 x K floats; every warp instruction loads 512 contiguous bytes
    // (lane = 16-byte piece; the thread-per-pixel form touched 32 rows per instruction and was bound by the L1 data pipe),
    // converts its four channels to fp16 hi / lo and stores two 8-byte pieces into the K-major core-matrix layout ------------
    {
      constexpr int F4 = K / 4;                                                  // float4 pieces per pixel row
      constexpr int PER = 128 * F4 / kPwThreads;                                 // pieces per thread
      constexpr int UB = PER < 8 ? PER : 8;                                      // loads in flight per batch
      static_assert(128 * F4 % kPwThreads == 0 && PER % UB == 0, "tile / thread shape");
      static_assert(PRO != PRO_LN || F4 <= 32, "the LayerNorm prologue reduces over the lanes of one warp");
      const long long px0 = (long long)tile * 128;
      const int live_rows = (int)((total_px - px0 < 128) ? (total_px - px0) : 128);
      const float4* src = reinterpret_cast<const float4*>(A + px0 * K);
#pragma unroll 1
      for (int i0 = 0; i0 < PER; i0 += UB) {
        float4 t[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int i = (i0 + u) * kPwThreads + tid;
          t[u] = (i / F4 < live_rows) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int i = (i0 + u) * kPwThreads + tid;
          const int r = i / F4, c = i - r * F4;
          float2 v0 = make_float2(t[u].x, t[u].y), v1 = make_float2(t[u].z, t[u].w);
          if constexpr (PRO == PRO_LN) {                                         // the F4 lanes of a pixel are neighbours in the warp
            float sum = (v0.x + v0.y) + (v1.x + v1.y);
#pragma unroll
            for (int o = 1; o < F4; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum * (1.0f / K);
            const float d0 = v0.x - mean, d1 = v0.y - mean, d2 = v1.x - mean, d3 = v1.y - mean;
            float ss = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
#pragma unroll
            for (int o = 1; o < F4; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float rstd = 1.0f / sqrtf(ss * (1.0f / K) + kLnEps);
            const float4 g = __ldg(reinterpret_cast<const float4*>(ln_g) + c), bb = __ldg(reinterpret_cast<const float4*>(ln_b) + c);
            v0 = make_float2(d0 * rstd * g.x + bb.x, d1 * rstd * g.y + bb.y);
            v1 = make_float2(d2 * rstd * g.z + bb.z, d3 * rstd * g.w + bb.w);
            if (r >= live_rows) v0 = v1 = make_float2(0.f, 0.f);
          }
          if constexpr (PRO == PRO_GELU) { v0 = gelu_pair(v0); v1 = gelu_pair(v1); }
          v0 = make_float2(v0.x * a_scale, v0.y * a_scale);
          v1 = make_float2(v1.x * a_scale, v1.y * a_scale);
          const uint32_t h0 = f2h2_sat(v0), h1 = f2h2_sat(v1);
          const float2 k0 = __half22float2(*reinterpret_cast<const __half2*>(&h0)), k1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
          const uint32_t l0 = f2h2_sat(make_float2(v0.x - k0.x, v0.y - k0.y)), l1 = f2h2_sat(make_float2(v1.x - k1.x, v1.y - k1.y));
          const int off = (c >> 1) * LBO + r * 16 + (c & 1) * 8;
          *reinterpret_cast<uint2*>(ah + off) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(al + off) = make_uint2(l0, l1);
        }
      }
    }
    // ---- main loop over weight slices ------------------------------------------------------------------------------------
#pragma unroll 1
    for (int ns = 0; ns < N / kNS; ++ns) {
      // stage the slice: for every K-chunk, kNS rows of 16 bytes are contiguous in the packed weights
      {
        const uint4* gh = reinterpret_cast<const uint4*>(wpack);
        const uint4* gl = reinterpret_cast<const uint4*>(wpack + (size_t)N * K);
        uint4* sh = reinterpret_cast<uint4*>(wh);
        uint4* sl = reinterpret_cast<uint4*>(wl);
        for (int i = tid; i < KC * kNS; i += kPwThreads) {
          const int kc = i / kNS, n = i - kc * kNS;
          sh[i] = __ldg(gh + kc * N + ns * kNS + n);
          sl[i] = __ldg(gl + kc * N + ns * kNS + n);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (warp == 0 && elect_one()) {              // (elect.sync: ptxas emits the MMAs back to back, no divergence loop)
        tc_fence_after();
        constexpr uint32_t idesc = umma_idesc(kNS);
        const uint32_t d = tmem + ns * kNS;
#pragma unroll
        for (int ks = 0; ks < K / 16; ++ks) {
          const uint64_t dah = umma_desc(a_h + ks * 2 * LBO, LBO, 128);
          const uint64_t dal = umma_desc(a_l + ks * 2 * LBO, LBO, 128);
          const uint64_t dbh = umma_desc(w_h + ks * 2 * kNS * 16, kNS * 16, 128);
          const uint64_t dbl = umma_desc(w_l + ks * 2 * kNS * 16, kNS * 16, 128);
          umma_f16(d, dah, dbh, idesc, ks > 0);
          umma_f16(d, dah, dbl, idesc, 1);
          umma_f16(d, dal, dbh, idesc, 1);
        }
        umma_commit(mbar);
      }
      mbar_wait(mbar, phase);                      // the slice buffer (and, after the last slice, A) may be overwritten
      phase ^= 1;
    }
    tc_fence_after();
    // ---- epilogue: this thread owns columns [half*N/2, (half+1)*N/2) of its pixel ----------------------------------
    {
      float* dst = Out + p * N;
      const float* res = (EPI == EPI_BIAS_RESID || EPI == EPI_GATE) ? resid + p * N : nullptr;
#pragma unroll 2
      for (int c0 = half * (N / 2); c0 < (half + 1) * (N / 2); c0 += 8) {
        float2 v[4];
        tmem_ld8(lane_addr + c0, v);
        tmem_ld_wait();
        const float4 b0 = *reinterpret_cast<const float4*>(sbias + c0), b1 = *reinterpret_cast<const float4*>(sbias + c0 + 4);
        v[0] = __ffma2_rn(v[0], make_float2(inv_scale, inv_scale), make_float2(b0.x, b0.y));
        v[1] = __ffma2_rn(v[1], make_float2(inv_scale, inv_scale), make_float2(b0.z, b0.w));
        v[2] = __ffma2_rn(v[2], make_float2(inv_scale, inv_scale), make_float2(b1.x, b1.y));
        v[3] = __ffma2_rn(v[3], make_float2(inv_scale, inv_scale), make_float2(b1.z, b1.w));
        if constexpr (EPI == EPI_BIAS_GELU) {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = gelu_pair(v[i]);
        }
        if (live) {
          if constexpr (EPI == EPI_BIAS_RESID) {
            const float4 r0 = __ldg(reinterpret_cast<const float4*>(res + c0)), r1 = __ldg(reinterpret_cast<const float4*>(res + c0 + 4));
            v[0] = __fadd2_rn(v[0], make_float2(r0.x, r0.y)); v[1] = __fadd2_rn(v[1], make_float2(r0.z, r0.w));
            v[2] = __fadd2_rn(v[2], make_float2(r1.x, r1.y)); v[3] = __fadd2_rn(v[3], make_float2(r1.z, r1.w));
          }
          if constexpr (EPI == EPI_GATE) {
            const float4 r0 = __ldg(reinterpret_cast<const float4*>(res + c0)), r1 = __ldg(reinterpret_cast<const float4*>(res + c0 + 4));
            v[0] = make_float2(v[0].x * gelu_grad_exact(r0.x), v[0].y * gelu_grad_exact(r0.y));
            v[1] = make_float2(v[1].x * gelu_grad_exact(r0.z), v[1].y * gelu_grad_exact(r0.w));
            v[2] = make_float2(v[2].x * gelu_grad_exact(r1.x), v[2].y * gelu_grad_exact(r1.y));
            v[3] = make_float2(v[3].x * gelu_grad_exact(r1.z), v[3].y * gelu_grad_exact(r1.w));
          }
          *reinterpret_cast<float4*>(dst + c0) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
          *reinterpret_cast<float4*>(dst + c0 + 4) = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
        }
      }
    }
    tc_fence_before();
    __syncthreads();                               // TMEM columns and the A tile are free for the next tile
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// ---- the same GEMM as a warp-specialised pipeline ---------------------------------------------------------------------------
// pwgemm_tc_kernel above runs the phases of a tile one after the other (stage A, per slice: stage W, MMAs, wait; epilogue) on one
// CTA per SM: 7-11 % issue utilisation, 15-28 % of the DRAM peak.  Here the phases overlap:
//   8 loader warps   stage the A operand (coalesced 16-byte pieces -> fp16 hi / lo, K-major, padded chunk stride) in STAGES of
//                    KS = K (K <= 128) or K / 2 channels, two stage buffers: the next stage / tile loads under the MMAs
//   1 producer lane  streams the weight slice of every (K stage, 64 output columns) with cp.async.bulk into two buffers
//   1 issuer lane    (elect.sync) issues the MMAs of a (K stage, slice) as soon as both operands are there and commits the
//                    buffers back; the accumulator of a tile is one of TWO TMEM buffers of N columns
//   8 epilogue warps read the finished accumulator (thread = pixel row x column half), apply bias / GELU / residual / gate and store, under
//                    the next tile's MMAs
namespace pipe {
// 32-byte global store: a thread's 8 output columns are one full sector (two 16-byte stores are two L2 transactions)
__device__ __forceinline__ void stg256(float* p, const float2 (&w)[4]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(w[0].x), "f"(w[0].y), "f"(w[1].x), "f"(w[1].y),
               "f"(w[2].x), "f"(w[2].y), "f"(w[3].x), "f"(w[3].y)
               : "memory");
}
constexpr int kThreads = 576;
constexpr int kLoaders = 256;                    // warps 8 .. 15
constexpr int kLBO = 128 * 16 + 16;              // padded K-chunk stride of the A stages
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
               "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
template <int K, int N>
struct Plan {
  static constexpr int NSW = N < lg::kNS ? N : lg::kNS;         // output columns per weight slice
  static constexpr int KS = K > 128 ? K / 2 : K;                // channels per A / W stage
  static constexpr int NKH = K / KS;
  static constexpr int A_STAGE = 2 * (KS / 8) * kLBO;           // hi + lo
  static constexpr int W_STAGE = 2 * (KS / 8) * NSW * 16;       // hi + lo
  static constexpr int TCOLS = (N <= 32) ? 32 : (N <= 64) ? 64 : (N <= 128 ? 128 : 256);
  // weight ring: as deep as 227 KB allow (the weight stream is what bounds the pipeline: an L2 round trip per stage)
  static constexpr int NWS = (128 + N * 4 + 2 * A_STAGE + 4 * W_STAGE + 128 <= 232448) ? 4 : (128 + N * 4 + 2 * A_STAGE + 3 * W_STAGE + 128 <= 232448) ? 3 : 2;
  static constexpr size_t smem = 128 + (size_t)N * 4 + 2 * (size_t)A_STAGE + NWS * (size_t)W_STAGE + 128;
};
}  // namespace pipe

template <int K, int N, int PRO, int EPI>
__global__ void __launch_bounds__(pipe::kThreads, 1)
pwgemm_pipe_kernel(const float* __restrict__ A, float* __restrict__ Out, const __half* __restrict__ wpack,
                   const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                   const float* __restrict__ resid, long long total_px, int num_tiles, const float* __restrict__ scale_dev) {
  using namespace pw;
  using PL = pipe::Plan<K, N>;
  constexpr int NSW = PL::NSW, KS = PL::KS, NKH = PL::NKH, NSL = N / NSW, LBO = pipe::kLBO;
  static_assert(K % 16 == 0 && KS % 16 == 0 && N % NSW == 0 && N % 16 == 0 && N <= 256, "tile shape");
  static_assert(PRO != PRO_LN || (NKH == 1 && K / 4 <= 32), "the LayerNorm prologue needs the whole pixel row in one stage, within one warp");
  const float a_scale = scale_dev ? __ldg(scale_dev) : 1.f, inv_scale = 1.f / a_scale;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* const base = smem_raw;          // (no manual re-alignment: through an integer cast the compiler loses the shared address space and emits generic LD / ST)
  constexpr int NWS = PL::NWS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base);                          // a_full[2] a_empty[2] acc_full[2] acc_empty[2] w_full[4] w_empty[4]
  uint64_t *a_full = bars, *a_empty = bars + 2, *acc_full = bars + 4, *acc_empty = bars + 6, *w_full = bars + 8, *w_empty = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + 128 - 8);
  float* sbias = reinterpret_cast<float*>(base + 128);                         // [N]
  unsigned char* a_st = base + 128 + N * 4;                                    // two A stages: [hi | lo][KS/8][LBO]
  unsigned char* w_st = a_st + 2 * PL::A_STAGE;                                // NWS weight stages: [hi | lo][KS/8][NSW][16 B]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a_full[b], pipe::kLoaders / 32);
      mbar_init(&a_empty[b], 1);
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);
    }
    for (int b = 0; b < NWS; ++b) {
      mbar_init(&w_full[b], 1);
      mbar_init(&w_empty[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 2 * PL::TCOLS);
  for (int i = tid; i < N; i += pipe::kThreads) sbias[i] = bias ? __ldg(bias + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp >= 8 && warp < 16) {
    // ---- loaders: A stages ---------------------------------------------------------------------------------------------------
    const int ltid = tid - 256;
    constexpr int F4 = K / 4;                                                    // float4 pieces per pixel row
    constexpr int F4S = KS / 4;                                                  // ... of one stage
    static_assert(pipe::kLoaders % F4S == 0, "a loader step covers whole rows");
    constexpr int RSTEP = pipe::kLoaders / F4S;                                  // rows per step of all loader threads
    constexpr int NSTEP = (128 + RSTEP - 1) / RSTEP;                             // steps per stage
    constexpr int UB = NSTEP < 16 ? NSTEP : 16;                                  // 16-byte loads in flight per thread (48 KB per SM)
    // a thread keeps its piece column c and walks down the rows r0, r0 + RSTEP, ...: addresses are base + immediates
    const int r0 = ltid / F4S, c = ltid - r0 * F4S;
    const int off0 = (c >> 1) * LBO + r0 * 16 + (c & 1) * 8;
    uint32_t sa = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
      const long long px0 = ((long long)blockIdx.x + (long long)lt * gridDim.x) * 128;
      const int live_rows = (int)((total_px - px0 < 128) ? (total_px - px0) : 128);
      const float4* src = reinterpret_cast<const float4*>(A + px0 * K) + r0 * F4 + c;
      for (int kh = 0; kh < NKH; ++kh, ++sa) {
        const uint32_t buf = sa & 1;
        unsigned char* ah = a_st + buf * PL::A_STAGE + off0;
        unsigned char* al = ah + (KS / 8) * LBO;
        const float4* sk = src + kh * F4S;
#pragma unroll 1
        for (int n0 = 0; n0 < NSTEP; n0 += UB) {
          float4 t[UB];
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int r = r0 + (n0 + u) * RSTEP;
            t[u] = (r < live_rows) ? __ldg(sk + (n0 + u) * RSTEP * F4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          // the loads fly while the stage buffer is still being read by the MMAs of two stages ago
          if (n0 == 0 && sa >= 2) mbar_wait(&a_empty[buf], ((sa >> 1) - 1) & 1);
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int r = r0 + (n0 + u) * RSTEP;
            float2 v0 = make_float2(t[u].x, t[u].y), v1 = make_float2(t[u].z, t[u].w);
            if constexpr (PRO == PRO_LN) {                                       // the F4 lanes of a pixel are neighbours in the warp
              float sum = (v0.x + v0.y) + (v1.x + v1.y);
#pragma unroll
              for (int o = 1; o < F4; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
              const float mean = sum * (1.0f / K);
              const float d0 = v0.x - mean, d1 = v0.y - mean, d2 = v1.x - mean, d3 = v1.y - mean;
              float ss = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
#pragma unroll
              for (int o = 1; o < F4; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
              const float rstd = 1.0f / sqrtf(ss * (1.0f / K) + kLnEps);
              const float4 g = __ldg(reinterpret_cast<const float4*>(ln_g) + c), bb = __ldg(reinterpret_cast<const float4*>(ln_b) + c);
              v0 = make_float2(d0 * rstd * g.x + bb.x, d1 * rstd * g.y + bb.y);
              v1 = make_float2(d2 * rstd * g.z + bb.z, d3 * rstd * g.w + bb.w);
              if (r >= live_rows) v0 = v1 = make_float2(0.f, 0.f);
            }
            if constexpr (PRO == PRO_GELU) { v0 = gelu_pair(v0); v1 = gelu_pair(v1); }
            v0 = make_float2(v0.x * a_scale, v0.y * a_scale);
            v1 = make_float2(v1.x * a_scale, v1.y * a_scale);
            const uint32_t h0 = f2h2_sat(v0), h1 = f2h2_sat(v1);
            const float2 k0 = __half22float2(*reinterpret_cast<const __half2*>(&h0)), k1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
            const uint32_t l0 = f2h2_sat(make_float2(v0.x - k0.x, v0.y - k0.y)), l1 = f2h2_sat(make_float2(v1.x - k1.x, v1.y - k1.y));
            if (r < 128) {
              *reinterpret_cast<uint2*>(ah + (n0 + u) * RSTEP * 16) = make_uint2(h0, h1);
              *reinterpret_cast<uint2*>(al + (n0 + u) * RSTEP * 16) = make_uint2(l0, l1);
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[buf]);
      }
    }
  } else if (warp == 16) {
    // ---- weight producer: one lane, cp.async.bulk of 1 KB rows (64 output columns x 16 bytes of one K chunk) ------------------
    if (lane == 0) {
      uint32_t sw = 0;
      const unsigned char* gw = reinterpret_cast<const unsigned char*>(wpack);
      constexpr uint32_t row_bytes = NSW * 16, stage_bytes = 2 * (KS / 8) * row_bytes;
      for (int lt = 0; lt < my_tiles; ++lt)
        for (int kh = 0; kh < NKH; ++kh)
          for (int ns = 0; ns < NSL; ++ns, ++sw) {
            const uint32_t buf = sw % NWS;
            if (sw >= NWS) mbar_wait(&w_empty[buf], ((sw / NWS) - 1) & 1);
            const uint32_t dst = smem_u32(w_st + buf * PL::W_STAGE);
            pipe::mbar_expect_tx(&w_full[buf], stage_bytes);
#pragma unroll 1
            for (int c = 0; c < KS / 8; ++c) {
              const size_t goff = ((size_t)(kh * (KS / 8) + c) * N + (size_t)ns * NSW) * 16;
              pipe::bulk_g2s(dst + c * row_bytes, gw + goff, row_bytes, &w_full[buf]);
              pipe::bulk_g2s(dst + (KS / 8) * row_bytes + c * row_bytes, gw + (size_t)N * K * 2 + goff, row_bytes, &w_full[buf]);
            }
          }
    }
  } else if (warp == 17) {
    // ---- MMA issuer -------------------------------------------------------------------------------------------------------------
    uint32_t sa = 0, sw = 0;
    const uint32_t a0 = smem_u32(a_st) >> 4, w0 = smem_u32(w_st) >> 4;
    constexpr uint32_t HI = (128u >> 4) | (1u << 14);
    auto dsc = [&](uint32_t base16, uint32_t byte_off, uint32_t lbo) -> uint64_t {
      return (uint64_t)(base16 + (byte_off >> 4) + ((lbo >> 4) << 16)) | ((uint64_t)HI << 32);
    };
    for (int lt = 0; lt < my_tiles; ++lt) {
      const uint32_t ab = lt & 1;
      if (lt >= 2) {
        mbar_wait(&acc_empty[ab], ((lt >> 1) - 1) & 1);
        tc_fence_after();
      }
      for (int kh = 0; kh < NKH; ++kh, ++sa) {
        const uint32_t abuf = sa & 1;
        mbar_wait(&a_full[abuf], (sa >> 1) & 1);
        for (int ns = 0; ns < NSL; ++ns, ++sw) {
          const uint32_t wbuf = sw % NWS;
          mbar_wait(&w_full[wbuf], (sw / NWS) & 1);
          tc_fence_after();
          if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc(NSW);
            const uint32_t d = tmem + ab * PL::TCOLS + ns * NSW;
            const uint32_t ah = a0 + ((abuf * PL::A_STAGE) >> 4), al = ah + (((KS / 8) * LBO) >> 4);
            const uint32_t wh = w0 + ((wbuf * PL::W_STAGE) >> 4), wl = wh + (((KS / 8) * NSW * 16) >> 4);
#pragma unroll
            for (int ks = 0; ks < KS / 16; ++ks) {
              const uint64_t dah = dsc(ah, ks * 2 * LBO, LBO), dal = dsc(al, ks * 2 * LBO, LBO);
              const uint64_t dbh = dsc(wh, ks * 2 * NSW * 16, NSW * 16), dbl = dsc(wl, ks * 2 * NSW * 16, NSW * 16);
              umma_f16(d, dah, dbh, idesc, (kh > 0 || ks > 0) ? 1u : 0u);
              umma_f16(d, dah, dbl, idesc, 1);
              umma_f16(d, dal, dbh, idesc, 1);
            }
            umma_commit(&w_empty[wbuf]);
            if (ns == NSL - 1) {
              umma_commit(&a_empty[abuf]);
              if (kh == NKH - 1) umma_commit(&acc_full[ab]);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 8) {
    // ---- epilogue: thread = (pixel row, column half) of the finished accumulator ------------------------------------------------
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    constexpr int NH = N / 2;                      // columns per thread
    const bool al32 = (reinterpret_cast<uintptr_t>(Out) & 31) == 0;
    constexpr int CB = NH >= 32 ? 4 : NH / 8;      // 8-column chunks per TMEM batch
    for (int lt = 0; lt < my_tiles; ++lt) {
      const uint32_t ab = lt & 1;
      const long long p = ((long long)blockIdx.x + (long long)lt * gridDim.x) * 128 + row;
      const bool live = p < total_px;
      mbar_wait(&acc_full[ab], (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t lane_addr = tmem + ab * PL::TCOLS + ((uint32_t)(q * 32) << 16) + half * NH;
      float* dst = Out + p * N + half * NH;
      const float* res = (EPI == EPI_BIAS_RESID || EPI == EPI_GATE) ? resid + p * N + half * NH : nullptr;
      const float* sb = sbias + half * NH;
      const bool ral32 = (reinterpret_cast<uintptr_t>(resid) & 31) == 0;   // residual / gate rows with 32-byte loads
#pragma unroll 1
      for (int cb = 0; cb < NH; cb += 8 * CB) {
        float2 v[CB][4];
#pragma unroll
        for (int j = 0; j < CB; ++j) tmem_ld8(lane_addr + cb + 8 * j, v[j]);
        tmem_ld_wait();
        if (cb + 8 * CB >= NH) {                   // last TMEM read of this accumulator: hand it back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
#pragma unroll
        for (int j = 0; j < CB; ++j) {
          const int c0 = cb + 8 * j;
          const float4 b0 = *reinterpret_cast<const float4*>(sb + c0), b1 = *reinterpret_cast<const float4*>(sb + c0 + 4);
          float2* w = v[j];
          w[0] = __ffma2_rn(w[0], make_float2(inv_scale, inv_scale), make_float2(b0.x, b0.y));
          w[1] = __ffma2_rn(w[1], make_float2(inv_scale, inv_scale), make_float2(b0.z, b0.w));
          w[2] = __ffma2_rn(w[2], make_float2(inv_scale, inv_scale), make_float2(b1.x, b1.y));
          w[3] = __ffma2_rn(w[3], make_float2(inv_scale, inv_scale), make_float2(b1.z, b1.w));
          if constexpr (EPI == EPI_BIAS_GELU) {
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = gelu_pair(w[i]);
          }
          if (live) {
            if constexpr (EPI == EPI_BIAS_RESID) {
              float4 r0, r1;
              if (ral32) ldg256(reinterpret_cast<const float4*>(res + c0), r0, r1);
              else { r0 = __ldg(reinterpret_cast<const float4*>(res + c0)); r1 = __ldg(reinterpret_cast<const float4*>(res + c0 + 4)); }
              w[0] = __fadd2_rn(w[0], make_float2(r0.x, r0.y)); w[1] = __fadd2_rn(w[1], make_float2(r0.z, r0.w));
              w[2] = __fadd2_rn(w[2], make_float2(r1.x, r1.y)); w[3] = __fadd2_rn(w[3], make_float2(r1.z, r1.w));
            }
            if constexpr (EPI == EPI_GATE) {
              float4 r0, r1;
              if (ral32) ldg256(reinterpret_cast<const float4*>(res + c0), r0, r1);
              else { r0 = __ldg(reinterpret_cast<const float4*>(res + c0)); r1 = __ldg(reinterpret_cast<const float4*>(res + c0 + 4)); }
              w[0] = make_float2(w[0].x * gelu_grad_exact(r0.x), w[0].y * gelu_grad_exact(r0.y));
              w[1] = make_float2(w[1].x * gelu_grad_exact(r0.z), w[1].y * gelu_grad_exact(r0.w));
              w[2] = make_float2(w[2].x * gelu_grad_exact(r1.x), w[2].y * gelu_grad_exact(r1.y));
              w[3] = make_float2(w[3].x * gelu_grad_exact(r1.z), w[3].y * gelu_grad_exact(r1.w));
            }
            if (al32) {
              pipe::stg256(dst + c0, v[j]);
            } else {                                 // output base only 16-byte aligned
              *reinterpret_cast<float4*>(dst + c0) = make_float4(w[0].x, w[0].y, w[1].x, w[1].y);
              *reinterpret_cast<float4*>(dst + c0 + 4) = make_float4(w[2].x, w[2].y, w[3].x, w[3].y);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 2 * PL::TCOLS);
}

// ---- depthwise 3x3 (+bias, zero pad) + GELU on an NHWC map (HBM-bound) ------------------------------------------------------
// thread = (4 consecutive pixels of a row, 4 channels): a sliding 3 x 6 window of float4 feeds four outputs, i.e. 4.5 loads per
// output vector instead of 9 and one set of tap loads per four pixels (the one-pixel-per-thread form was bound by the L1 data
// pipe: 574 us per 134 M-element map)
constexpr int kDwPx = 4;
__global__ void __launch_bounds__(256) dwconv_gelu_kernel(const float* __restrict__ hid, float* __restrict__ act,
                                                           const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                           int H, int W, int C4, long long total_items) {
  extern __shared__ __align__(16) float s_dw[];     // [9][C4] taps then [C4] bias
  for (int i = threadIdx.x; i < C4; i += 256) {
#pragma unroll
    for (int t = 0; t < 9; ++t) s_dw[t * C4 + i] = __ldg(dw_w + i * 9 + t);
    s_dw[9 * C4 + i] = __ldg(dw_b + i);
  }
  __syncthreads();
  const long long v = (long long)blockIdx.x * 256 + threadIdx.x;
  if (v >= total_items) return;
  const int vecs = C4 / 4;
  const int c = (int)(v % vecs) * 4;
  const long long seg = v / vecs;                   // segment of kDwPx pixels
  const int segs = (W + kDwPx - 1) / kDwPx;
  const int x0 = (int)(seg % segs) * kDwPx;
  const long long t = seg / segs;
  const int y = (int)(t % H);
  const long long n = t / H;
  const float4 b = *reinterpret_cast<const float4*>(s_dw + 9 * C4 + c);
  float2 a0[kDwPx], a1[kDwPx];
#pragma unroll
  for (int i = 0; i < kDwPx; ++i) { a0[i] = make_float2(b.x, b.y); a1[i] = make_float2(b.z, b.w); }
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    const float* rowp = hid + ((n * H + yy) * W) * (long long)C4 + c;
    float4 wv[3];
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) wv[dx] = *reinterpret_cast<const float4*>(s_dw + ((dy + 1) * 3 + dx) * C4 + c);
#pragma unroll
    for (int j = 0; j < kDwPx + 2; ++j) {           // input column x0 - 1 + j feeds outputs j - 2 .. j
      const int xx = x0 - 1 + j;
      if (xx < 0 || xx >= W) continue;
      const float4 hv = __ldg(reinterpret_cast<const float4*>(rowp + (long long)xx * C4));
      const float2 h0 = make_float2(hv.x, hv.y), h1 = make_float2(hv.z, hv.w);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int o = j - dx;                       // output pixel x0 + o reads column x0 + o - 1 + dx
        if (o >= 0 && o < kDwPx) {
          a0[o] = __ffma2_rn(make_float2(wv[dx].x, wv[dx].y), h0, a0[o]);
          a1[o] = __ffma2_rn(make_float2(wv[dx].z, wv[dx].w), h1, a1[o]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kDwPx; ++i) {
    if (x0 + i >= W) continue;
    const float2 g0 = gelu_pair(a0[i]), g1 = gelu_pair(a1[i]);
    *reinterpret_cast<float4*>(act + (((n * H + y) * W) + x0 + i) * (long long)C4 + c) = make_float4(g0.x, g0.y, g1.x, g1.y);
  }
}

template <int K, int N, int PRO, int EPI>
static cudaError_t pwgemm_launch(const float* A, float* Out, const void* wpack, const float* bias, const float* ln_g,
                                 const float* ln_b, const float* resid, long long total_px, cudaStream_t s,
                                 const float* scale_dev = nullptr) {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles = (int)((total_px + 127) / 128);
  static const bool serial = [] { const char* e = getenv("LGTEUN_PWGEMM"); return e && e[0] == 's'; }();   // A/B: the one-phase-at-a-time kernel
  if (!serial) {
    using PL = pipe::Plan<K, N>;
    cudaError_t e = cudaFuncSetAttribute(pwgemm_pipe_kernel<K, N, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PL::smem);
    if (e != cudaSuccess) return e;
    const int grid = tiles < sm_count ? tiles : sm_count;
    pwgemm_pipe_kernel<K, N, PRO, EPI><<<grid, pipe::kThreads, PL::smem, s>>>(A, Out, reinterpret_cast<const __half*>(wpack), bias, ln_g,
                                                                         ln_b, resid, total_px, tiles, scale_dev);
    return cudaGetLastError();
  }
  constexpr int NS = N < kNS ? N : kNS;
  const size_t smem = 16 + (size_t)N * 4 + 16 + (size_t)2 * (K / 8) * (128 * 16 + 16) + (size_t)(2 * NS * K) * 2 + 128;
  cudaError_t e = cudaFuncSetAttribute(pwgemm_tc_kernel<K, N, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // one persistent CTA per SM (measured: two co-resident CTAs of the K = 64 variant are 5 % slower, the prologue is bound by the L1 data pipe)
  const int grid = tiles < sm_count ? tiles : sm_count;
  pwgemm_tc_kernel<K, N, PRO, EPI><<<grid, kPwThreads, smem, s>>>(A, Out, reinterpret_cast<const __half*>(wpack), bias, ln_g,
                                                                  ln_b, resid, total_px, tiles, scale_dev);
  return cudaGetLastError();
}

// ---- training step (train.cu): the conv-FFN's 1x1 convs and their data gradients on the same kernel -------------------------
//   Y[P,N] = f(s X[P,K]) . W^T / s (+ bias) (+ resid | * gelu'(gate)),  wpack = split-fp16 pack of W[N][K]
//   pro: 0 plain, 2 GELU on X;  epi: 0 bias, 2 bias + resid, 3 gate (aux = resid / gate tensor [P,N])
template <int K, int N>
static cudaError_t train_pw_dispatch(int pro, int epi, const float* X, float* Y, const void* wpack, const float* bias,
                                     const float* aux, long long px, const float* scale_dev, cudaStream_t s) {
  if (pro == PRO_PLAIN && epi == EPI_BIAS) return pwgemm_launch<K, N, PRO_PLAIN, EPI_BIAS>(X, Y, wpack, bias, nullptr, nullptr, nullptr, px, s, scale_dev);
  if (pro == PRO_GELU && epi == EPI_BIAS) return pwgemm_launch<K, N, PRO_GELU, EPI_BIAS>(X, Y, wpack, bias, nullptr, nullptr, nullptr, px, s, scale_dev);
  if (pro == PRO_GELU && epi == EPI_BIAS_RESID) return pwgemm_launch<K, N, PRO_GELU, EPI_BIAS_RESID>(X, Y, wpack, bias, nullptr, nullptr, aux, px, s, scale_dev);
  if (pro == PRO_PLAIN && epi == EPI_GATE) return pwgemm_launch<K, N, PRO_PLAIN, EPI_GATE>(X, Y, wpack, bias, nullptr, nullptr, aux, px, s, scale_dev);
  return cudaErrorInvalidValue;
}
bool train_pwgemm_supported(int K, int N) {
  const int c = K < N ? K : N, c4 = K < N ? N : K;
  return ((c == 16 || c == 32 || c == 64) && (c4 == 4 * c || (K == N && (K == 64 || K == 128 || K == 256)))) ||
         (K == N && (K == 16 || K == 32));       // the mixer's projection
}
cudaError_t launch_train_pwgemm(int K, int N, int pro, int epi, const float* X, float* Y, const void* wpack, const float* bias,
                                const float* aux, long long px, const float* scale_dev, cudaStream_t s) {
#define LG_TPW(KK, NN) if (K == KK && N == NN) return train_pw_dispatch<KK, NN>(pro, epi, X, Y, wpack, bias, aux, px, scale_dev, s)
  LG_TPW(16, 64); LG_TPW(32, 128); LG_TPW(64, 256);      // c -> 4c
  LG_TPW(64, 64); LG_TPW(128, 128); LG_TPW(256, 256);    // 4c -> 4c
  LG_TPW(16, 16); LG_TPW(32, 32);                        // c -> c
  LG_TPW(64, 16); LG_TPW(128, 32); LG_TPW(256, 64);      // 4c -> c
#undef LG_TPW
  return cudaErrorInvalidValue;
}
// W[n*wso + k*wsi] (fp32, any strides: transposed views are free) -> hi | lo fp16 in the UMMA layout [K/8][N][8]
__global__ void pack_umma_f16_strided_kernel(const float* __restrict__ w, int wso, int wsi, __half* __restrict__ hi,
                                             __half* __restrict__ lo, int N, int K) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= N * K) return;
  const int n = idx / K, k = idx - n * K;
  const float v = w[(size_t)n * wso + (size_t)k * wsi];
  const __half h = __float2half_rn(v);
  const size_t o = ((size_t)(k >> 3) * N + n) * 8 + (k & 7);
  hi[o] = h;
  lo[o] = __float2half_rn(v - __half2float(h));
}
cudaError_t launch_pack_umma_f16_strided(const float* w, int wso, int wsi, void* pack, int N, int K, cudaStream_t s) {
  __half* hi = reinterpret_cast<__half*>(pack);
  pack_umma_f16_strided_kernel<<<(N * K + 255) / 256, 256, 0, s>>>(w, wso, wsi, hi, hi + (size_t)N * K, N, K);
  return cudaGetLastError();
}


// ---- weight gradient of a 1x1 conv on the tensor cores (training backward) -------------------------------------------------------
//   dW[co*wso + ci*wsi] += sum_p dY[p,co] * f(X[p,ci]);   db[co] += sum_p dY[p,co]          (K of the GEMM = pixels)
// A CTA walks 128-pixel tiles: both operands are transposed while they are staged — a thread loads 8 consecutive pixels of
// one channel and writes them as one 16-byte K-run of the core-matrix layout — so that A = dY^T [co][pixel] (128 rows, rows
// >= Cout stay zero) and B = f(X)^T [ci][pixel] are K-major; every tile adds 8 x (hi*hi + hi*lo + lo*hi) MMAs into the same
// [128 x NCI] fp32 TMEM accumulator, which is read once at the end and added to dW with atomics (the CTAs split the pixels).
// dY is a gradient (~1e-7): it is multiplied by the step's power-of-two scale before the fp16 split, dW divided by it.
// Two groups of 8 warps work on alternate 64-pixel tiles, each with its own operand buffers, mbarrier and TMEM accumulator: one
// group's loads are in flight while the other converts or waits for its MMAs (a single group spent most of a tile waiting:
// 24k clocks per 128 pixels at 8 warps, 13k with batched loads, now ).  A thread issues every load of its share of a tile
// (up to 8 items = 64 values) before it converts any of them.  A is stored with only the live rows (K-group stride =
// mrows * 16 bytes): the M = 128 MMA then reads rows >= mrows out of the following K-groups / buffers, which only fills
// accumulator rows nobody reads.
constexpr int kGradThreads = 512;
template <int NCI, int ACT>
__global__ void __launch_bounds__(kGradThreads, 1)
pwgrad_tc_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ dY, int ldy, int Cout, float* __restrict__ dW,
                 int wso, int wsi, float* __restrict__ db, long long total_px, int num_tiles, const float* __restrict__ scale_dev) {
  using namespace pw;
  static_assert(NCI % 16 == 0 && NCI <= 256, "tile shape");
  constexpr int KC = 8, TP = 64;                                                // 64 pixels per tile = 8 K-chunks of 8
  constexpr int GT = 256;                                                       // threads per group
  constexpr int NX = KC * NCI, XI = (NX + GT - 1) / GT;                         // X items per tile, per thread
  constexpr int DI = 4;                                                         // dY items per thread (KC * 128 / GT)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);                       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = warp >> 3, gtid = tid & (GT - 1);
  const int co0 = blockIdx.y * 128;
  const int live = Cout - co0 < 128 ? Cout - co0 : 128;                         // output channels of this CTA
  int mrows = 16;                                                               // rows of A: the next power of two
  while (mrows < live) mrows *= 2;
  const int lm = __ffs(mrows) - 1;
  const uint32_t a_bytes = 2u * KC * mrows * 16, b_bytes = 2u * KC * NCI * 16;  // hi + lo of one group
  __half* ah = reinterpret_cast<__half*>(smem_raw + 128 + grp * a_bytes);       // [KC][mrows][8]
  __half* al = ah + KC * mrows * 8;
  __half* bh = reinterpret_cast<__half*>(smem_raw + 128 + 2 * a_bytes + grp * b_bytes);   // [KC][NCI][8]
  __half* bl = bh + KC * NCI * 8;
  constexpr uint32_t TCOLS = (NCI <= 16) ? 32 : (NCI <= 32) ? 64 : (NCI <= 64) ? 128 : (NCI <= 128 ? 256 : 512);
  const float a_scale = scale_dev ? __ldg(scale_dev) : 1.f, inv_scale = 1.f / a_scale;

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot, acc = tmem + grp * NCI;
  const uint32_t a_h = smem_u32(ah), a_l = smem_u32(al), b_h = smem_u32(bh), b_l = smem_u32(bl);
  const uint32_t lbo_a = (uint32_t)mrows * 16;
  uint32_t phase = 0;
  float bsum = 0.f;
  bool first = true;
  const int n_dy = KC * mrows;                                                  // <= DI * GT
  // tile-invariant part of every item: source offset (floats, relative to the tile's first pixel), destination, validity.
  // GT % mrows == 0: a thread keeps its channel co.
  const int co = gtid & (mrows - 1);
  int dyo[DI], dyd[DI];
#pragma unroll
  for (int u = 0; u < DI; ++u) {
    const int item = gtid + u * GT, g = item >> lm;
    dyo[u] = (item < n_dy && co < live) ? g * 8 * ldy + co0 + co : -1;          // rows in [live, mrows) are staged as zeros
    dyd[u] = item < n_dy ? ((g << lm) + co) * 8 : -1;
  }
  int xo[XI];
#pragma unroll
  for (int u = 0; u < XI; ++u) {
    const int item = gtid + u * GT, ci = item % NCI, g = item / NCI;
    xo[u] = item < NX ? g * 8 * ldx + ci : -1;
  }

  for (int tile = blockIdx.x * 2 + grp; tile < num_tiles; tile += 2 * gridDim.x) {   // total_px % 64 == 0 (launch check)
    const float* __restrict__ dyt = dY + (long long)tile * TP * ldy;
    const float* __restrict__ xt = X + (long long)tile * TP * ldx;
    constexpr int XB = XI < 4 ? XI : 4;                                         // X items whose loads are issued together
    float td[DI][8], tx[XB][8];
#pragma unroll
    for (int u = 0; u < DI; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) td[u][j] = dyo[u] >= 0 ? __ldg(dyt + (dyo[u] + j * ldy)) : 0.f;
#pragma unroll
    for (int u = 0; u < XB; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) tx[u][j] = xo[u] >= 0 ? __ldg(xt + (xo[u] + j * ldx)) : 0.f;
#pragma unroll
    for (int u = 0; u < DI; ++u) {
      if (dyd[u] >= 0) {
        float2 v[4];
#pragma unroll
        for (int j = 0; j < 8; ++j) bsum += td[u][j];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = make_float2(td[u][2 * j] * a_scale, td[u][2 * j + 1] * a_scale);
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4*>(ah + dyd[u]) = hi;
        *reinterpret_cast<uint4*>(al + dyd[u]) = lo;
      }
    }
#pragma unroll
    for (int xb = 0; xb < XI; xb += XB) {
      if (xb > 0) {
#pragma unroll
        for (int u = 0; u < XB; ++u)
#pragma unroll
          for (int j = 0; j < 8; ++j) tx[u][j] = xo[xb + u] >= 0 ? __ldg(xt + (xo[xb + u] + j * ldx)) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < XB; ++u) {
        if (xo[xb + u] >= 0) {
          const int item = gtid + (xb + u) * GT;                                // destination [g][ci][8] = item * 8
          float2 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[j] = make_float2(tx[u][2 * j], tx[u][2 * j + 1]);
            if constexpr (ACT == 1) v[j] = gelu_pair(v[j]);
          }
          uint4 hi, lo;
          split8(v, hi, lo);
          *reinterpret_cast<uint4*>(bh + item * 8) = hi;
          *reinterpret_cast<uint4*>(bl + item * 8) = lo;
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(GT) : "memory");
    if ((warp & 7) == 0 && elect_one()) {
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc(NCI);
#pragma unroll
      for (int ks = 0; ks < KC / 2; ++ks) {
        const uint64_t dah = umma_desc(a_h + ks * 2 * lbo_a, lbo_a, 128);
        const uint64_t dal = umma_desc(a_l + ks * 2 * lbo_a, lbo_a, 128);
        const uint64_t dbh = umma_desc(b_h + ks * 2 * NCI * 16, NCI * 16, 128);
        const uint64_t dbl = umma_desc(b_l + ks * 2 * NCI * 16, NCI * 16, 128);
        umma_f16(acc, dah, dbh, idesc, !(first && ks == 0));
        umma_f16(acc, dah, dbl, idesc, 1);
        umma_f16(acc, dal, dbh, idesc, 1);
      }
      umma_commit(&mbar[grp]);
    }
    first = false;
    mbar_wait(&mbar[grp], phase);                  // this group's operand tiles may be overwritten
    phase ^= 1;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  {
    const int q = warp & 3, part = warp >> 2;      // lane quarter, column slice
    const int row = q * 32 + lane;                 // output channel co0 + row
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const bool used0 = blockIdx.x * 2 < num_tiles, used1 = blockIdx.x * 2 + 1 < num_tiles;
#pragma unroll 1
    for (int c0 = part * 8; c0 < NCI && used0; c0 += 32) {
      float2 v[4], u[4];
      tmem_ld8(lane_addr + c0, v);
      if (used1) tmem_ld8(lane_addr + NCI + c0, u);
      tmem_ld_wait();
      if (row < live) {
        float* dst = dW + (size_t)(co0 + row) * wso;
        float r[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          r[2 * j] = (used1 ? v[j].x + u[j].x : v[j].x) * inv_scale;
          r[2 * j + 1] = (used1 ? v[j].y + u[j].y : v[j].y) * inv_scale;
        }
        if (wsi == 1 && wso % 4 == 0 && (reinterpret_cast<uintptr_t>(dW) & 15) == 0) {   // 16-byte atomics (see pwgrad_bulk_kernel)
          atomicAdd(reinterpret_cast<float4*>(dst + c0), make_float4(r[0], r[1], r[2], r[3]));
          atomicAdd(reinterpret_cast<float4*>(dst + c0 + 4), make_float4(r[4], r[5], r[6], r[7]));
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) atomicAdd(dst + (size_t)(c0 + j) * wsi, r[j]);
        }
      }
    }
    if (db && !first && gtid < n_dy && co < live) atomicAdd(db + co0 + co, bsum);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// ---- the same weight gradient as a bulk-copy pipeline (default) ---------------------------------------------------------------
// The direct kernel above spends its issue slots on 40 - 64 scalar global loads per thread and tile with 64-bit address
// arithmetic and then waits for them (long-scoreboard stalls 35 %).  Here a producer warp moves the raw fp32 pixel rows of a tile
// into shared memory with cp.async.bulk (dense operands: one 8 - 32 KB copy per tile and operand — per-row copies of 128 -
// 512 bytes were 2.5x slower than the direct kernel —, completion counted on an mbarrier), eight converter
// warps read them with immediate-offset LDS, apply the GELU / scale, split to fp16 hi | lo and write the K-major operand
// tiles, and one elected thread issues the MMAs; raw tiles and operand tiles are both double-buffered, so the copy of tile
// t + 2, the conversion of tile t + 1 and the MMAs of tile t overlap.  One TMEM accumulator for the CTA's whole pixel share.
namespace gradb {
constexpr int kConv = 512;                       // warps 0 .. 15 (8 warps left the SM's issue slots half idle: dependent LDS -> GELU -> split chains)
constexpr int kThreads = kConv + 64;             // then the producer warp and the MMA issuer warp
}
template <int NCI, int ACT>
__global__ void __launch_bounds__(gradb::kThreads, 1)
pwgrad_bulk_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ dY, int ldy, int Cout, float* __restrict__ dW,
                   int wso, int wsi, float* __restrict__ db, int num_tiles, int tp, int nraw, const float* __restrict__ scale_dev) {
  using namespace pw;
  static_assert(NCI % 16 == 0 && NCI <= 256, "tile shape");
  constexpr int GT = gradb::kConv;
  constexpr int XI = (8 * NCI + GT - 1) / GT;                                   // X items per thread at tp = 64
  constexpr int DI = 8 * 128 / GT;                                              // dY items per thread at tp = 64, 128 rows
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t *raw_full = bars, *raw_empty = bars + 4, *op_full = bars + 8, *op_empty = bars + 10, *done = bars + 12;   // raw: up to 4 stages
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 120);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kc = tp >> 3;                                                       // K-chunks of 8 pixels per tile (4 or 8)
  const int co0 = blockIdx.y * 128;
  const int live = Cout - co0 < 128 ? Cout - co0 : 128;                         // output channels of this CTA (a multiple of 4)
  int mrows = 16;                                                               // rows of A: the next power of two
  while (mrows < live) mrows *= 2;
  const int lm = __ffs(mrows) - 1;
  const uint32_t raw_bytes = (uint32_t)tp * (NCI + mrows) * 4;                  // X rows, then dY rows (live floats each)
  const uint32_t a_half = (uint32_t)kc * mrows * 16, b_half = (uint32_t)kc * NCI * 16;
  const uint32_t op_bytes = 2 * a_half + 2 * b_half;                            // A hi | A lo | B hi | B lo
  unsigned char* raw0 = smem_raw + 128;
  unsigned char* op0 = raw0 + nraw * raw_bytes;
  constexpr uint32_t TCOLS = (NCI <= 32) ? 32 : (NCI <= 64) ? 64 : (NCI <= 128 ? 128 : 256);
  const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    for (int b = 0; b < 4; ++b) {
      mbar_init(&raw_full[b], 1);
      mbar_init(&raw_empty[b], GT / 32);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&op_full[b], GT / 32);
      mbar_init(&op_empty[b], 1);
    }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == GT / 32) {
    // ---- producer ----------------------------------------------------------------------------------------------------------------
    for (int t = 0; t < my_tiles; ++t) {
      const int s = t % nraw, use = t / nraw;
      if (use >= 1) mbar_wait(&raw_empty[s], (use - 1) & 1);
      const long long p0 = ((long long)blockIdx.x + (long long)t * gridDim.x) * tp;
      if (lane == 0) pipe::mbar_expect_tx(&raw_full[s], (uint32_t)tp * (NCI + live) * 4);
      __syncwarp();
      const uint32_t xdst = smem_u32(raw0 + s * raw_bytes), ydst = xdst + (uint32_t)tp * NCI * 4;
      if (lane == 0) {                             // dense operands (launch check): a tile is one contiguous block of each
        pipe::bulk_g2s(xdst, X + p0 * NCI, (uint32_t)tp * NCI * 4, &raw_full[s]);
        pipe::bulk_g2s(ydst, dY + p0 * live, (uint32_t)tp * live * 4, &raw_full[s]);
      }
    }
  } else if (warp == GT / 32 + 1) {
    // ---- MMA issuer ---------------------------------------------------------------------------------------------------------------
    const uint32_t lbo_a = (uint32_t)mrows * 16;
    for (int t = 0; t < my_tiles; ++t) {
      const int s = t & 1;
      mbar_wait(&op_full[s], (t >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc(NCI);
        const uint32_t a_h = smem_u32(op0 + s * op_bytes), a_l = a_h + a_half, b_h = a_l + a_half, b_l = b_h + b_half;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (2 * ks < kc) {
            const uint64_t dah = umma_desc(a_h + ks * 2 * lbo_a, lbo_a, 128);
            const uint64_t dal = umma_desc(a_l + ks * 2 * lbo_a, lbo_a, 128);
            const uint64_t dbh = umma_desc(b_h + ks * 2 * NCI * 16, NCI * 16, 128);
            const uint64_t dbl = umma_desc(b_l + ks * 2 * NCI * 16, NCI * 16, 128);
            umma_f16(tmem, dah, dbh, idesc, !(t == 0 && ks == 0));
            umma_f16(tmem, dah, dbl, idesc, 1);
            umma_f16(tmem, dal, dbh, idesc, 1);
          }
        }
        umma_commit(&op_empty[s]);
        if (t == my_tiles - 1) umma_commit(done);
      }
      __syncwarp();
    }
  } else {
    // ---- converters ---------------------------------------------------------------------------------------------------------------
    const float a_scale = scale_dev ? __ldg(scale_dev) : 1.f, inv_scale = 1.f / a_scale;
    const int co = tid & (mrows - 1);                                           // GT % mrows == 0: a thread keeps its channel
    const int n_dy = kc * mrows, n_x = kc * NCI;
    float bsum = 0.f;
    for (int t = 0; t < my_tiles; ++t) {
      const int s = t & 1, sr = t % nraw;
      const float* xr = reinterpret_cast<const float*>(raw0 + sr * raw_bytes);
      const float* yr = xr + tp * NCI;
      __half* ah = reinterpret_cast<__half*>(op0 + s * op_bytes);
      __half* al = ah + kc * mrows * 8;
      __half* bh = al + kc * mrows * 8;
      __half* bl = bh + kc * NCI * 8;
      mbar_wait(&raw_full[sr], (t / nraw) & 1);
      if (t >= 2) mbar_wait(&op_empty[s], ((t >> 1) - 1) & 1);
#pragma unroll
      for (int u = 0; u < DI; ++u) {
        const int item = tid + u * GT;
        if (item < n_dy && co < live) {                                         // rows in [live, mrows) only feed accumulator rows nobody reads
          const int g = item >> lm;
          const float* src = yr + g * 8 * live + co;
          float2 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float e0 = src[(2 * j) * live], e1 = src[(2 * j + 1) * live];
            bsum += e0 + e1;
            v[j] = make_float2(e0 * a_scale, e1 * a_scale);
          }
          uint4 hi, lo;
          split8(v, hi, lo);
          *reinterpret_cast<uint4*>(ah + ((g << lm) + co) * 8) = hi;
          *reinterpret_cast<uint4*>(al + ((g << lm) + co) * 8) = lo;
        }
      }
#pragma unroll
      for (int u = 0; u < XI; ++u) {
        const int item = tid + u * GT;
        if (item < n_x) {
          const int ci = item % NCI, g = item / NCI;
          const float* src = xr + g * 8 * NCI + ci;
          float2 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[j] = make_float2(src[(2 * j) * NCI], src[(2 * j + 1) * NCI]);
            if constexpr (ACT == 1) v[j] = gelu_pair(v[j]);
          }
          uint4 hi, lo;
          split8(v, hi, lo);
          *reinterpret_cast<uint4*>(bh + item * 8) = hi;                        // [g][ci][8] = item * 8
          *reinterpret_cast<uint4*>(bl + item * 8) = lo;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&op_full[s]);
        mbar_arrive(&raw_empty[sr]);
      }
    }
    // ---- epilogue: accumulator -> dW (atomics: the CTAs split the pixels) ---------------------------------------------------------
    if (my_tiles > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
      const int q = warp & 3, part = warp >> 2;    // lane quarter, column slice
      const int row = q * 32 + lane;               // output channel co0 + row
      const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
      const bool vec = wsi == 1 && wso % 4 == 0 && (reinterpret_cast<uintptr_t>(dW) & 15) == 0;
#pragma unroll 1
      for (int c0 = part * 8; c0 < NCI; c0 += 8 * (GT / 128)) {
        float2 v[4];
        tmem_ld8(lane_addr + c0, v);
        tmem_ld_wait();
        if (row < live) {
          float* dst = dW + (size_t)(co0 + row) * wso;
          if (vec) {                               // 16-byte atomics: 148 CTAs add into the same few thousand words
            atomicAdd(reinterpret_cast<float4*>(dst + c0), make_float4(v[0].x * inv_scale, v[0].y * inv_scale, v[1].x * inv_scale, v[1].y * inv_scale));
            atomicAdd(reinterpret_cast<float4*>(dst + c0 + 4), make_float4(v[2].x * inv_scale, v[2].y * inv_scale, v[3].x * inv_scale, v[3].y * inv_scale));
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              atomicAdd(dst + (size_t)(c0 + 2 * j) * wsi, v[j].x * inv_scale);
              atomicAdd(dst + (size_t)(c0 + 2 * j + 1) * wsi, v[j].y * inv_scale);
            }
          }
        }
      }
      if (db && tid < n_dy && co < live) atomicAdd(db + co0 + co, bsum);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

template <int NCI>
static cudaError_t pwgrad_launch(int act, const float* X, int ldx, const float* dY, int ldy, int Cout, float* dW, int wso, int wsi,
                                 float* db, long long px, const float* scale_dev, cudaStream_t s) {
  if (px % 64) return cudaErrorInvalidValue;
  int mrows = 16;
  while (mrows < Cout && mrows < 128) mrows *= 2;
  static const bool direct = [] { const char* e = getenv("LGTEUN_PWGRAD"); return e && std::string(e) == "direct"; }();
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  cudaError_t e;
  if (!direct && al16(X) && al16(dY) && ldx == NCI && ldy == Cout && Cout <= 128 && Cout % 4 == 0) {
    // bulk-copy pipeline: two raw fp32 stages + two operand stages = 16 bytes per pixel and channel
    const int tp = (NCI + mrows) <= 192 ? 64 : 32;
    const int tiles = (int)(px / tp);
    const size_t per_stage = (size_t)4 * tp * (NCI + mrows);                     // one raw stage; an operand stage is the same size
    int nraw = 4;                                                              // as many raw stages in flight as 227 KB allow
    while (nraw > 2 && 128 + (nraw + 2) * per_stage + 128 * 16 > 232448) --nraw;
    const size_t smem = 128 + (nraw + 2) * per_stage + 128 * 16;
    const dim3 grid(tiles < 148 ? tiles : 148, (Cout + 127) / 128);
    if (act) {
      e = cudaFuncSetAttribute(pwgrad_bulk_kernel<NCI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      pwgrad_bulk_kernel<NCI, 1><<<grid, gradb::kThreads, smem, s>>>(X, ldx, dY, ldy, Cout, dW, wso, wsi, db, tiles, tp, nraw, scale_dev);
    } else {
      e = cudaFuncSetAttribute(pwgrad_bulk_kernel<NCI, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      pwgrad_bulk_kernel<NCI, 0><<<grid, gradb::kThreads, smem, s>>>(X, ldx, dY, ldy, Cout, dW, wso, wsi, db, tiles, tp, nraw, scale_dev);
    }
    return cudaGetLastError();
  }
  const int tiles = (int)(px / 64);
  // two groups x (A hi|lo [8][mrows][8] + B hi|lo [8][NCI][8]) + slack for the M = 128 read past the last live row
  const size_t smem = 128 + 2 * (size_t)(2 * 8 * mrows * 16 + 2 * 8 * NCI * 16) + 128 * 16;
  const dim3 grid(tiles < 2 * 148 ? (tiles + 1) / 2 : 148, (Cout + 127) / 128);
  if (act) {
    e = cudaFuncSetAttribute(pwgrad_tc_kernel<NCI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pwgrad_tc_kernel<NCI, 1><<<grid, kGradThreads, smem, s>>>(X, ldx, dY, ldy, Cout, dW, wso, wsi, db, px, tiles, scale_dev);
  } else {
    e = cudaFuncSetAttribute(pwgrad_tc_kernel<NCI, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pwgrad_tc_kernel<NCI, 0><<<grid, kGradThreads, smem, s>>>(X, ldx, dY, ldy, Cout, dW, wso, wsi, db, px, tiles, scale_dev);
  }
  return cudaGetLastError();
}
bool train_pwgrad_supported(int Cin, int Cout) {     // Cout: any count up to 256 (rows up to the next power of two are zeros)
  auto p2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  return p2(Cin) && Cin >= 16 && Cin <= 256 && Cout >= 1 && Cout <= 256;
}
cudaError_t launch_train_pwgrad(int Cin, int Cout, int act, const float* X, int ldx, const float* dY, int ldy, float* dW, int wso,
                                int wsi, float* db, long long px, const float* scale_dev, cudaStream_t s) {
  switch (Cin) {
    case 16: return pwgrad_launch<16>(act, X, ldx, dY, ldy, Cout, dW, wso, wsi, db, px, scale_dev, s);
    case 32: return pwgrad_launch<32>(act, X, ldx, dY, ldy, Cout, dW, wso, wsi, db, px, scale_dev, s);
    case 64: return pwgrad_launch<64>(act, X, ldx, dY, ldy, Cout, dW, wso, wsi, db, px, scale_dev, s);
    case 128: return pwgrad_launch<128>(act, X, ldx, dY, ldy, Cout, dW, wso, wsi, db, px, scale_dev, s);
    case 256: return pwgrad_launch<256>(act, X, ldx, dY, ldy, Cout, dW, wso, wsi, db, px, scale_dev, s);
  }
  return cudaErrorInvalidValue;
}

// conv-FFN of a c = 64 block.  scratch: two buffers of N*H*W*256 floats (h1 / act, hidden).
cudaError_t launch_ffn_wide_tc(const BlockW& w, const float* x, float* buf_a, float* buf_b, float* y, int N, int H, int W,
                               cudaStream_t s) {
  constexpr int C = 64, C4 = 256;
  const long long px = (long long)N * H * W;
  const char* pack = reinterpret_cast<const char*>(w.ffn_pack);
  const size_t s0 = (size_t)C4 * C * 2, s1 = (size_t)C4 * C4 * 2;       // bytes per half-tensor (hi or lo)
  cudaError_t e;
  e = pwgemm_launch<C, C4, PRO_LN, EPI_BIAS_GELU>(x, buf_a, pack, w.f0_b, w.ln2_w, w.ln2_b, nullptr, px, s);
  if (e != cudaSuccess) return e;
  e = pwgemm_launch<C4, C4, PRO_PLAIN, EPI_BIAS>(buf_a, buf_b, pack + 2 * s0, w.f1_b, nullptr, nullptr, nullptr, px, s);
  if (e != cudaSuccess) return e;
  const long long vec = (long long)N * H * ((W + kDwPx - 1) / kDwPx) * (C4 / 4);
  dwconv_gelu_kernel<<<(unsigned)((vec + 255) / 256), 256, 10 * C4 * sizeof(float), s>>>(buf_b, buf_a, w.dw_w, w.dw_b, H, W, C4, vec);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return pwgemm_launch<C4, C, PRO_PLAIN, EPI_BIAS_RESID>(buf_a, y, pack + 2 * s0 + 2 * s1, w.f2_b, nullptr, nullptr, x, px, s);
}

}  // namespace lg
