// common.cuh — shared declarations of the LGTEUN sm_100a kernels.
//
// Internal data layout: the data module works on NCHW planes exactly as the reference
// boundary delivers them; everything inside the Local-Global Transformer prior is NHWC fp32
// ([N,H,W,c], channels innermost = 64..256 B per pixel) so that a pixel's channel vector is one
// contiguous, vectorisable run and window / tile addressing is pure index arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lg {

constexpr int kWin = 8;              // window_size, models/unlg_former.py:47
constexpr int kHeads = 2;            // num_heads,   models/unlg_former.py:48
constexpr float kLnEps = 1e-5f;      // nn.LayerNorm default, LGT.py:58
constexpr int kMaxStages = 8;

// ---- weight views (pointers into the packed arena; native PyTorch layouts) ---------------------
struct BlockW {                       // one LGB block: LGT.py:231-239
  const float *ln1_w, *ln1_b;         // blocks.j.0.fn.norm
  const float *pos;                   // local_mixer.pos_emb        [1,2,64,64]
  const float *pos_t;                 // derived: pos transposed to [2][key j][query i]
  const float *qkv_w, *qkv_b;         // local_mixer.to_qkv         [3c/2, c/2]
  const float *amp_w, *amp_b;         // global_mixer.conv_amp.0    [c/2]
  const float *pha_w, *pha_b;         // global_mixer.conv_pha.0    [c/2]
  const float *proj_w, *proj_b;       // proj                       [c, c]
  const float *ln2_w, *ln2_b;         // blocks.j.1.fn.norm
  const float *f0_w, *f0_b;           // net.0                      [4c, c]
  const float *f1_w, *f1_b;           // net.2.point_conv           [4c, 4c]
  const float *dw_w, *dw_b;           // net.2.depth_conv           [4c, 3, 3]
  const float *f2_w, *f2_b;           // net.4                      [c, 4c]
  const float *f0_wt, *f1_wt, *f2_wt; // derived: the three FFN weights transposed to [in][out] (CUDA-core path)
  const float *ffn_pack;              // derived: fp16 hi/lo of W0,W1,W2 in UMMA K-major core-matrix layout (ffn_tc.cu)
  const float *ffn_cl_pack;           // derived: the same weights + bias K-steps in the order ffn_cl.cu copies to shared memory
  const float *proj_pack;             // derived: fp16 hi | lo of proj_w in the same layout (fft256.cu)
};

struct PriorW {                       // one LGT: LGT.py:251-303
  const float *pe_dw_w, *pe_dw_b;     // patch_embed.proj.0         [B]
  const float *pe_w, *pe_b;           // patch_embed.proj.1         [C, B]
  const float *pe_ln_w, *pe_ln_b;     // patch_embed.norm           [C]
  BlockW enc[2], bott[1], dec[2];
  const float *down_w, *down_b;       // encoder_layers.0.1.1       [2C, C]
  const float *down_wt;               // derived: down_w transposed to [C][2C]
  const float *up_w, *up_b;           // decoder_layers.0.0.1       [C, 2C]
  const float *fuse_w, *fuse_b;       // decoder_layers.0.1         [C, 2C]   input = [upsampled | skip]
  const float *tail_w, *tail_b;       // tail.1                     [B, C]
};

struct DataW {                        // models/unlg_former.py:29-40
  const float *d1_w, *d1_b, *d3_w, *d3_b;       // D.1, D.3     [B,3,3]
  const float *dt1_w, *dt1_b, *dt3_w, *dt3_b;   // DT.1, DT.3
  const float *r_w, *r_b;                       // R            [1,B]
  const float *rt_w, *rt_b;                     // RT           [B,1]
  const float *eta[kMaxStages];                 // eta.i        []
};

// ---- kernel launchers (one per .cu file) ---------------------------------------------------------
// data_step.cu
int data_step_launches();   // 1 (fused persistent kernel) or 2 (LGTEUN_DATA_STEP=split)
size_t data_step_scratch_floats(int N, int B, int h, int w);   // size of launch_data_step's `resid` scratch
cudaError_t launch_bicubic(const float* x, float* y, int planes, int h, int w, int num, int den, cudaStream_t s);
cudaError_t launch_data_step(const DataW& w, int stage, int B, const float* z_in, const float* ms, const float* pan,
                             float* resid /*scratch: data_step_scratch_floats()*/, float* z_out, int N, int h, int wd, cudaStream_t s);
// pixel_ops.cu
cudaError_t launch_patch_embed(const PriorW& w, int B, const float* x_nchw, float* y, int N, int H, int W, cudaStream_t s);
cudaError_t launch_down(const PriorW& w, int C, const float* x, float* y, int N, int H, int W, cudaStream_t s);
cudaError_t launch_up_fuse(const PriorW& w, int C, const float* low, const float* skip, float* t_low /*[N,H/2,W/2,C] scratch*/,
                           float* y, int N, int H, int W, cudaStream_t s);
cudaError_t launch_tail(const PriorW& w, int B, const float* fea, const float* x_nchw, float* y_nchw, int N, int H, int W,
                        cudaStream_t s);
// window_msa.cu
//  pre_ln = 1: x is the un-normalised [N,H,W,c] map, LN(blocks.j.0.fn.norm) is applied on the fly and the first
//  c/2 channels are used;  pre_ln = 0: x is already the [N,H,W,c/2] local half.
cudaError_t launch_window_msa(const BlockW& w, int c, const float* x, float* y_half, int pre_ln, int N, int H, int W,
                              cudaStream_t s);
// window_msa_tc.cu — the same operator on tcgen05 / TMEM (c = 16, 32).  variant 0 = hybrid (QKV projection and Q.K^T +
// positional bias on the tensor pipe, softmax and P.V on the CUDA cores), 1 = all three GEMMs on the tensor pipe.
// launch_window_msa picks per channel count (LGTEUN_MSA=simt|hybrid|tc overrides, for A/B measurement)
bool window_msa_tc_supported(int c);
cudaError_t launch_window_msa_tc(const BlockW& w, int c, const float* x, float* y_half, int pre_ln, int N, int H, int W,
                                 int variant, cudaStream_t s);
// fft_mixer.cu
size_t spectrum_floats(int N, int H, int W, int c2);
cudaError_t launch_fft_rows_fwd(const BlockW& w, int c, const float* x, float* spec, int pre_ln, int N, int H, int W,
                                cudaStream_t s);
cudaError_t launch_fft_cols(const BlockW& w, int c, float* spec, int N, int H, int W, cudaStream_t s);
//  proj = 1: y[N,H,W,c] = proj(cat(local, |irfft|)) + xres ;  proj = 0: y[N,H,W,c/2] = |irfft| only.
cudaError_t launch_fft_rows_inv(const BlockW& w, int c, const float* spec, const float* local, const float* xres,
                                float* y, int proj, int N, int H, int W, cudaStream_t s);
cudaError_t fft_init_tables(cudaStream_t s);
// fft256.cu — register-resident radix-16 passes for 256-point transforms
cudaError_t fft256_init_tables(cudaStream_t s);
cudaError_t launch_fft_cols256(const BlockW& w, int c2, float* spec, int N, int W, cudaStream_t s);
cudaError_t launch_fft_cols128(const BlockW& w, int c2, float* spec, int N, int W, cudaStream_t s);   // H = 128
cudaError_t launch_fft_rows_fwd256(const BlockW& w, int c, const float* x, float* spec, int N, int H, cudaStream_t s);
// plain register-resident passes (no LayerNorm / mixing / projection) for the companion operator of companion_ops.cu
cudaError_t launch_fft_cols_plain(int H, int c2, float* spec, int N, int W, int dir, cudaStream_t s, int fixreal = 1);                 // H in {128, 256}
// plain row transforms of the training step at W in {128, 256}, c2 in {8, 16, 32} (fft256.cu)
bool fft_rows_plain_supported(int W, int c2, int ldx, const void* p0, const void* p1, const void* p2);
cudaError_t launch_fft_rows_r2c(int W, int c2, const float* x, int ldx, const float* sgn, float* spec, size_t rows, float scale,
                                float wint, cudaStream_t s);
cudaError_t launch_fft_rows_c2r(int W, int c2, const float* spec, float* xout, int ldo, float* xabs, int ldabs, size_t rows,
                                float scale, float wint, cudaStream_t s);
// forward columns + amp/pha fusion + inverse columns of Freprocess in one launch; cudaErrorNotSupported if (H, C) is not built.
// fuse_w = {amp_fuse.0.weight, .0.bias, amp_fuse.2.weight, .2.bias, pha_fuse.0.weight, .0.bias, pha_fuse.2.weight, .2.bias}
cudaError_t launch_fre_cols_fused(int H, int C, const float* S, float* G, const float* const* fuse_w, int N, int W, cudaStream_t s);
cudaError_t launch_fft_rows_fwd_pre(int W, int C, const float* msf, const float* panf, const float* const* pre_w, float* spec, int N,
                                    int H, cudaStream_t s);
cudaError_t launch_fft_rows_inv_post(int W, int C, const float* spec, const float* post_w, const float* post_b, float* y_nchw, int N,
                                     int H, cudaStream_t s);
cudaError_t launch_fft_rows_fwd128(const BlockW& w, int c, const float* x, float* spec, int N, int H, cudaStream_t s);   // W = 128
cudaError_t launch_fft_rows_inv256(const BlockW& w, int c, const float* spec, const float* local, const float* xres, float* y,
                                   int N, int H, cudaStream_t s);
cudaError_t launch_fft_rows_inv128(const BlockW& w, int c, const float* spec, const float* local, const float* xres, float* y,
                                   int N, int H, cudaStream_t s);   // W = 128
// ffn.cu
size_t ffn_hidden_floats(int N, int H, int W, int c);
cudaError_t launch_ffn(const BlockW& w, int c, const float* x, float* hidden, float* y, int N, int H, int W,
                       cudaStream_t s);
// ffn_tc.cu — fused tcgen05/TMEM FFN (c = 16 or 32)
size_t ffn_tc_pack_halves(int c);
cudaError_t launch_pack_umma_f16(const float* w, void* hi, void* lo, int N, int K, cudaStream_t s);
cudaError_t launch_pack_umma_f16_bias(const float* w, const float* bias, void* hi, void* lo, int N, int K, cudaStream_t s);  // K+8 columns
cudaError_t launch_ffn_tc(const BlockW& w, int c, const float* x, float* y, int N, int H, int W, cudaStream_t s);
// ffn_cl.cu — the fused FFN with hidden channels on TMEM lanes and pixels in TMEM columns (c = 16 or 32); default
size_t ffn_cl_pack_halves(int c);
cudaError_t launch_ffn_cl_pack(const BlockW& b, int c, void* pack, cudaStream_t s);
cudaError_t launch_ffn_cl(const BlockW& w, int c, const float* x, float* y, int N, int H, int W, cudaStream_t s);
// pwgemm_tc.cu — conv-FFN of the widest level (c = 64) as tcgen05 pixel-GEMMs; buf_a / buf_b: N*H*W*256 floats each
cudaError_t launch_ffn_wide_tc(const BlockW& w, const float* x, float* buf_a, float* buf_b, float* y, int N, int H, int W,
                               cudaStream_t s);
// training-step GEMMs on the same tcgen05 kernel (train.cu): Y[P,N] = f(s X[P,K]) . W^T / s (+ bias) (+ resid | * gelu'(gate));
// pro: 0 plain, 2 GELU on X; epi: 0 bias (may be NULL), 2 bias + resid, 3 gate; wpack from launch_pack_umma_f16_strided
bool train_pwgemm_supported(int K, int N);
cudaError_t launch_train_pwgemm(int K, int N, int pro, int epi, const float* X, float* Y, const void* wpack, const float* bias,
                                const float* aux, long long px, const float* scale_dev, cudaStream_t s);
cudaError_t launch_pack_umma_f16_strided(const float* w, int wso, int wsi, void* pack /*N*K floats*/, int N, int K, cudaStream_t s);
// weight gradient of a 1x1 conv on tcgen05 (K = pixels): dW[co*wso + ci*wsi] += sum_p dY[p,co] f(X[p,ci]), db[co] += sum_p dY[p,co]
bool train_pwgrad_supported(int Cin, int Cout);
cudaError_t launch_train_pwgrad(int Cin, int Cout, int act, const float* X, int ldx, const float* dY, int ldy, float* dW, int wso,
                                int wsi, float* db, long long px, const float* scale_dev, cudaStream_t s);
// metrics.cu — PSNR / SAM / ERGAS per image in fp64 (acc: N*(2+2B) doubles scratch, out: N*3 doubles)
cudaError_t launch_metrics(const float* pred, const float* gt, double* acc, double* out, int N, int B, int H, int W,
                           float max_value, cudaStream_t s);
cudaError_t launch_normalize(const float* raw, float* out, size_t n, float max_value, cudaStream_t s);
cudaError_t launch_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, float scale, cudaStream_t s);
// misc
cudaError_t launch_transpose_pos(const float* pos, float* pos_t, cudaStream_t s);
cudaError_t launch_transpose(const float* src, float* dst, int rows, int cols, cudaStream_t s);   // dst[c][r] = src[r][c]

// ---- device helpers -----------------------------------------------------------------------------
#ifdef __CUDACC__
// Exact-form (erf) GELU x*Phi(x) on a PAIR of values with Blackwell's packed fp32 pipe.
//   gelu(x) = relu(x) - |x| * Phi(-|x|),   Phi(-a) = erfc(a / sqrt 2) / 2 = 2^(-Q(a))
// -log2 Phi(-a) is smooth (1 at 0, ~ a^2 log2(e)/2 for large a), so a degree-5 polynomial Q fitted with weight
// a * Phi(-a) reproduces GELU to 4.8e-7 absolute over the whole real line (6.4e-7 including fp32 rounding; torch's own
// fp32 erf-GELU is 1.2e-6 from the fp64 value).  The leading coefficient is positive: Q grows without bound, the
// correction underflows to 0 for large |x| and gelu(x) -> relu(x) exactly.  The polynomial is evaluated in n = -|x|
// (odd coefficients negated), which also is the factor of the correction term.
// Cost per pair: 2 LOP3 + 5 FFMA2 + 2 MUFU.EX2 + 2 FMNMX + 1 FFMA2 = 12 issue slots and ONE special-function op per
// element (the Abramowitz-Stegun 7.1.26 form used before needed 18 slots and two MUFU ops; erff() needs ~27/element).
// The conv-FFN is bound by these epilogues, not by the tensor pipe.
__device__ __forceinline__ float2 gelu_pair(float2 x) {
  const float2 n = make_float2(__int_as_float(__float_as_int(x.x) | 0x80000000), __int_as_float(__float_as_int(x.y) | 0x80000000));
  float2 q = make_float2(4.7329402712e-04f, 4.7329402712e-04f);
  q = __ffma2_rn(q, n, make_float2(7.0844612347e-03f, 7.0844612347e-03f));
  q = __ffma2_rn(q, n, make_float2(5.1827168585e-02f, 5.1827168585e-02f));
  q = __ffma2_rn(q, n, make_float2(-4.5999264548e-01f, -4.5999264548e-01f));
  q = __ffma2_rn(q, n, make_float2(1.1507877699e+00f, 1.1507877699e+00f));
  q = __ffma2_rn(q, n, make_float2(-1.0000376324e+00f, -1.0000376324e+00f));
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(q.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(q.y));
  return __ffma2_rn(n, e, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}
__device__ __forceinline__ float gelu_fast(float x) { return gelu_pair(make_float2(x, x)).x; }

// d/dx of the exact (erf) GELU: Phi(x) + x phi(x), for the backward kernels.  Phi(-a) = 2^(-Q(a)) with its own degree-5 Q
// (minimax in the absolute error of Phi: 3.3e-7 including fp32 evaluation, i.e. the accuracy of an erff-based fp32 form),
// phi(x) = 2^(-x^2 log2(e) / 2) / sqrt(2 pi): two ex2 and ~12 issue slots instead of erff + expf (~35).
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float a = fabsf(x);
  float q = 5.1941370432e-04f;
  q = fmaf(q, a, -7.3901742096e-03f);
  q = fmaf(q, a, 5.2543108209e-02f);
  q = fmaf(q, a, 4.59273491e-01f);
  q = fmaf(q, a, 1.1510838432e+00f);
  q = fmaf(q, a, 1.0000007771e+00f);
  float e, g;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-q));                       // Phi(-|x|)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(x * x * -0.72134752044f));  // exp(-x^2 / 2)
  const float Phi = x >= 0.f ? 1.f - e : e;
  return fmaf(x * 0.3989422804014327f, g, Phi);
}

// LayerNorm over C register-resident channels (biased variance, eps inside the sqrt).
template <int C>
__device__ __forceinline__ void layer_norm_inplace(float (&v)[C], const float* __restrict__ g,
                                                   const float* __restrict__ b) {
  float mean = 0.f;
#pragma unroll
  for (int i = 0; i < C; ++i) mean += v[i];
  mean *= (1.0f / C);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    float d = v[i] - mean;
    var = fmaf(d, d, var);
  }
  float rstd = 1.0f / sqrtf(var * (1.0f / C) + kLnEps);
#pragma unroll
  for (int i = 0; i < C; ++i) v[i] = (v[i] - mean) * rstd * g[i] + b[i];
}

// A thread reading its own pixel row: 32-byte loads (LDG.256) when the row allows it — with a lane stride of 64 - 256 bytes every
// load instruction costs one L1 wavefront per lane whatever its width, so halving the instruction count halves the L1 data-pipe
// traffic that bounds the thread-per-pixel prologues (LayerNorm of the FFT row pass, window MSA).
template <int C>
__device__ __forceinline__ void load_vec(float (&v)[C], const float* __restrict__ p) {
  static_assert(C % 4 == 0, "channel vectors are multiples of 4");
  if constexpr (C % 8 == 0) {
    if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
#pragma unroll
      for (int i = 0; i < C / 8; ++i)
        asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[8 * i]), "=f"(v[8 * i + 1]), "=f"(v[8 * i + 2]), "=f"(v[8 * i + 3]), "=f"(v[8 * i + 4]), "=f"(v[8 * i + 5]),
                       "=f"(v[8 * i + 6]), "=f"(v[8 * i + 7])
                     : "l"(p + 8 * i));
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < C / 4; ++i) {
    float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
// two adjacent float4 with one 32-byte load (p 32-byte aligned, read-only data)
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p));
}
template <int C>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&v)[C]) {
  if constexpr (C % 8 == 0) {
    if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {      // 32-byte stores: one full L2 sector per thread and instruction
#pragma unroll
      for (int i = 0; i < C / 8; ++i)
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p + 8 * i), "f"(v[8 * i]), "f"(v[8 * i + 1]),
                     "f"(v[8 * i + 2]), "f"(v[8 * i + 3]), "f"(v[8 * i + 4]), "f"(v[8 * i + 5]), "f"(v[8 * i + 6]), "f"(v[8 * i + 7])
                     : "memory");
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < C / 4; ++i)
    *reinterpret_cast<float4*>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
#endif

}  // namespace lg
