// train.cu — the training step behind include/lgteun.h:
//   UnlgFormer.train_iter  models/unlg_former.py:87-113 = forward in train() mode (Dropout(0.1) active, LGT.py:198,216),
//   nn.L1Loss (models/base/losses.py:19-40), loss.backward(), Adam (models/base/base_model.py:116-131).
// The forward records every tensor the backward needs on a bump-allocated tape; the backward walks the same structure in
// reverse and accumulates parameter gradients into one flat buffer whose layout equals the flat parameter layout
// (lgteun_weight_offset), so that data-parallel training needs ONE all-reduce over it.
// As in the reference only the last prior is live (unlg_former.py:63-67): priors 0..K-2 are not executed and their
// parameters receive no gradient (torch leaves their .grad at None; here the flat gradient is zero there).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "ctx.cuh"
#include "train_kernels.cuh"

using namespace lg;
using namespace lgctx;
using namespace lgtrain;

namespace lgctx {

struct BlockTape {
  float *Xin, *A, *qkv, *lse, *ppos, *F, *v, *cat, *Xmid, *A2, *h1, *h2, *h3, *Xout;
};
struct PriorTape {
  const float* zin;
  float *t0, *t1, *X0, *d0, *L0, *u0, *u1, *fz, *fea;
  BlockTape enc[2], bott, dec[2];
};
struct DataTape {
  const float* Zin;
  float *a1, *a2, *a3, *e, *u1, *u2, *u3, *T1, *r, *Zout;
};

struct TrainState {
  float* base = nullptr;
  size_t cap = 0;
  // tape of the last lgteun_train_forward
  bool valid = false;
  uint64_t generation = 0;      // bumped by every lgteun_train_forward: a backward names the forward it belongs to
  int N = 0, h = 0, w = 0;
  float p_drop = 0.f;
  uint64_t seed = 0;
  const uint64_t* seed_dev = nullptr;   // when set (lgteun_train_set_seed_ptr) the dropout kernels read the seed from device memory
  size_t fwd_end = 0;           // arena offset where the backward's scratch starts
  const float* flat_param = nullptr;
  const float *ms = nullptr, *pan = nullptr;
  float* gscale = nullptr;      // device scalar: power of two that brings max|dLoss/dOut| to [0.5, 1) (tensor-core backward GEMMs)
  DataTape data[kMaxStages];
  PriorTape prior;
  int launches_fwd = 0, launches_bwd = 0;
  const float* ext_mask[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // recorded dropout masks (lgteun_train_set_masks)
  bool use_ext = false;
};

void train_destroy(lgteun_ctx* c) {
  if (!c->train) return;
  if (c->train->base) cudaFree(c->train->base);
  delete c->train;
  c->train = nullptr;
}

}  // namespace lgctx

namespace {

struct Run {                    // one pass over the step: dry = size the arena only
  float* base;
  size_t off = 0;
  bool dry = false;
  cudaStream_t s = nullptr;
  cudaError_t err = cudaSuccess;
  int launches = 0;
  float* take(size_t floats) {
    float* p = base + off;
    off += (floats + 63) & ~(size_t)63;
    return p;
  }
  void check() {
    ++launches;
    if (err == cudaSuccess) err = cudaGetLastError();
  }
  void zero(float* p, size_t floats) {
    if (dry) return;
    if (err == cudaSuccess) err = cudaMemsetAsync(p, 0, floats * sizeof(float), s);
  }
};

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
unsigned blocks(size_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

template <typename K>
void optin_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// y = W f(x) + b (+ add) (* gelu'(gate))
void pw(Run& R, int act, TV x, int Cin, const float* W, int wso, int wsi, const float* b, TV y, int Cout, size_t NP,
        const TV* add = nullptr, const TV* gate = nullptr) {
  if (R.dry) return;
  const TV none{nullptr, 0, 0, 0, 0};
  const TV a = add ? *add : none, g = gate ? *gate : none;
  const int ua = add != nullptr, ug = gate != nullptr;
  if (Cout <= 16) {
    dim3 grid(blocks(NP, 256), 1);
    if (act) k_pw<1, 16><<<grid, 256, 0, R.s>>>(x, Cin, W, wso, wsi, b, y, Cout, NP, a, ua, g, ug);
    else k_pw<0, 16><<<grid, 256, 0, R.s>>>(x, Cin, W, wso, wsi, b, y, Cout, NP, a, ua, g, ug);
  } else if (Cout <= 32) {              // one block covers all output channels: x is read once
    dim3 grid(blocks(NP, 128), 1);
    if (act) k_pw<1, 32><<<grid, 256, 0, R.s>>>(x, Cin, W, wso, wsi, b, y, Cout, NP, a, ua, g, ug);
    else k_pw<0, 32><<<grid, 256, 0, R.s>>>(x, Cin, W, wso, wsi, b, y, Cout, NP, a, ua, g, ug);
  } else {
    dim3 grid(blocks(NP, 64), (Cout + 63) / 64);
    if (act) k_pw<1, 64><<<grid, 256, 0, R.s>>>(x, Cin, W, wso, wsi, b, y, Cout, NP, a, ua, g, ug);
    else k_pw<0, 64><<<grid, 256, 0, R.s>>>(x, Cin, W, wso, wsi, b, y, Cout, NP, a, ua, g, ug);
  }
  R.check();
}
bool use_tc_wgrad(int Cin, int Cout) {
  static const bool simt = [] { const char* e = getenv("LGTEUN_TRAIN_GEMM"); return e && std::string(e) == "simt"; }();
  return !simt && train_pwgrad_supported(Cin, Cout);
}
void pw_wgrad(Run& R, int act, TV x, int Cin, TV dy, int Cout, const float* dW, int wso, int wsi, const float* db, size_t NP,
              const float* gscale = nullptr) {
  if (R.dry) return;
  if (gscale && !x.nchw && !dy.nchw && NP % 64 == 0 && use_tc_wgrad(Cin, Cout)) {
    cudaError_t e = launch_train_pwgrad(Cin, Cout, act, x.p, x.ld, dy.p, dy.ld, const_cast<float*>(dW), wso, wsi,
                                        const_cast<float*>(db), (long long)NP, gscale, R.s);
    ++R.launches;
    if (R.err == cudaSuccess) R.err = e;
    return;
  }
  const int ci_tiles = (Cin + 63) / 64, co_tiles = (Cout + 63) / 64;
  const size_t tiles = (NP + 31) / 32;
  const unsigned gx = (unsigned)std::min<size_t>(tiles, std::max(1, 148 * 4 / (ci_tiles * co_tiles)));
  dim3 grid(gx, ci_tiles * co_tiles);
  if (act) k_pw_wgrad<1><<<grid, 256, 0, R.s>>>(x, Cin, dy, Cout, const_cast<float*>(dW), wso, wsi, const_cast<float*>(db), NP, ci_tiles);
  else k_pw_wgrad<0><<<grid, 256, 0, R.s>>>(x, Cin, dy, Cout, const_cast<float*>(dW), wso, wsi, const_cast<float*>(db), NP, ci_tiles);
  R.check();
}
// The conv-FFN's three 1x1 convs and their data gradients run on the tcgen05 pixel-GEMM of pwgemm_tc.cu (split-fp16 operands,
// fp32 TMEM accumulation) when the channel counts fit it; LGTEUN_TRAIN_GEMM=simt keeps them on the CUDA-core GEMM (A/B runs).
bool use_tc_gemm(int K, int N) {
  static const bool simt = [] { const char* e = getenv("LGTEUN_TRAIN_GEMM"); return e && std::string(e) == "simt"; }();
  return !simt && train_pwgemm_supported(K, N);
}
// Y[NP,N] = f(X[NP,K]) . W^T (+ bias) (+ aux | * gelu'(aux)); W[n*wso + k*wsi] is re-packed every step (the weights move)
void tc_pw(Run& R, int K, int N, int pro, int epi, const float* X, float* Y, const float* W, int wso, int wsi, const float* bias,
           const float* aux, size_t NP, const float* scale) {
  float* pack = R.take((size_t)N * K);
  if (R.dry) return;
  cudaError_t e = launch_pack_umma_f16_strided(W, wso, wsi, pack, N, K, R.s);
  ++R.launches;
  if (e == cudaSuccess) { e = launch_train_pwgemm(K, N, pro, epi, X, Y, pack, bias, aux, (long long)NP, scale, R.s); ++R.launches; }
  if (R.err == cudaSuccess) R.err = e;
}
bool ln_fast(TV a, TV b, int C, size_t NP) {
  return !a.nchw && !b.nchw && a.ld == C && b.ld == C && (C == 16 || C == 32 || C == 64) && NP % 2 == 0;
}
bool ln_v4(const void* a, const void* b, const void* c, const void* d, size_t NP) {   // 16-byte accesses, 8 pixels per warp at most
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al(a) && al(b) && al(c) && al(d) && NP % 8 == 0;
}
void ln_fwd(Run& R, TV x, int C, const float* g, const float* b, TV y, size_t NP) {
  if (R.dry) return;
  const unsigned gw = (unsigned)std::min<size_t>(148 * 8, (NP + 63) / 64);
  if (ln_fast(x, y, C, NP) && ln_v4(x.p, y.p, g, b, NP)) {
    if (C == 16) k_ln_v4<16, 0><<<gw, 256, 0, R.s>>>(x.p, g, b, nullptr, y.p, 0, nullptr, nullptr, NP);
    else if (C == 32) k_ln_v4<32, 0><<<gw, 256, 0, R.s>>>(x.p, g, b, nullptr, y.p, 0, nullptr, nullptr, NP);
    else k_ln_v4<64, 0><<<gw, 256, 0, R.s>>>(x.p, g, b, nullptr, y.p, 0, nullptr, nullptr, NP);
  } else if (ln_fast(x, y, C, NP)) {
    if (C == 16) k_ln_warp<16, 0><<<gw, 256, 0, R.s>>>(x.p, g, b, nullptr, y.p, 0, nullptr, nullptr, NP);
    else if (C == 32) k_ln_warp<32, 0><<<gw, 256, 0, R.s>>>(x.p, g, b, nullptr, y.p, 0, nullptr, nullptr, NP);
    else k_ln_warp<64, 0><<<gw, 256, 0, R.s>>>(x.p, g, b, nullptr, y.p, 0, nullptr, nullptr, NP);
  } else {
    k_ln_fwd<<<blocks(NP), 256, 0, R.s>>>(x, C, g, b, y, NP);
  }
  R.check();
}
void ln_bwd(Run& R, TV x, int C, const float* g, TV dy, TV dx, int accumulate, const float* dg, const float* db, size_t NP) {
  if (R.dry) return;
  float *pg = const_cast<float*>(dg), *pb = const_cast<float*>(db);
  const unsigned gw = (unsigned)std::min<size_t>(148 * 8, (NP + 63) / 64);
  if (ln_fast(x, dy, C, NP) && !dx.nchw && dx.ld == C && ln_v4(x.p, dy.p, dx.p, g, NP)) {
    if (C == 16) k_ln_v4<16, 1><<<gw, 256, 0, R.s>>>(x.p, g, nullptr, dy.p, dx.p, accumulate, pg, pb, NP);
    else if (C == 32) k_ln_v4<32, 1><<<gw, 256, 0, R.s>>>(x.p, g, nullptr, dy.p, dx.p, accumulate, pg, pb, NP);
    else k_ln_v4<64, 1><<<gw, 256, 0, R.s>>>(x.p, g, nullptr, dy.p, dx.p, accumulate, pg, pb, NP);
  } else if (ln_fast(x, dy, C, NP) && !dx.nchw && dx.ld == C) {
    if (C == 16) k_ln_warp<16, 1><<<gw, 256, 0, R.s>>>(x.p, g, nullptr, dy.p, dx.p, accumulate, pg, pb, NP);
    else if (C == 32) k_ln_warp<32, 1><<<gw, 256, 0, R.s>>>(x.p, g, nullptr, dy.p, dx.p, accumulate, pg, pb, NP);
    else k_ln_warp<64, 1><<<gw, 256, 0, R.s>>>(x.p, g, nullptr, dy.p, dx.p, accumulate, pg, pb, NP);
  } else {
    k_ln_bwd<<<blocks(NP), 256, 2 * C * sizeof(float), R.s>>>(x, C, g, dy, dx, accumulate, pg, pb, NP);
  }
  R.check();
}
void dwconv(Run& R, int K, TV x, const float* w, const float* b, TV y, int N, int H, int W, int C, int flip,
            const TV* add = nullptr, float add_scale = 1.f) {
  if (R.dry) return;
  const TV none{nullptr, 0, 0, 0, 0};
  const int lh = ilog2(H), lw = ilog2(W), lc = ilog2(C);
  if (K == 3 && !add && !x.nchw && !y.nchw && x.ld == C && y.ld == C && C >= 4 && W >= 8) {
    static const bool v4 = [] { const char* e = getenv("LGTEUN_DW3"); return e && std::string(e) == "v4"; }();
    if (W >= 16 && !v4) k_dw3_v2<16><<<blocks((size_t)N * H * (W / 16) * C / 2), 256, 0, R.s>>>(x.p, w, b, y.p, N, lh, lw, lc, flip);
    else if (W >= 16) k_dw3_v4<16><<<blocks((size_t)N * H * (W / 16) * C / 4), 256, 0, R.s>>>(x.p, w, b, y.p, N, lh, lw, lc, flip);
    else k_dw3_v4<8><<<blocks((size_t)N * H * (W / 8) * C / 4), 256, 0, R.s>>>(x.p, w, b, y.p, N, lh, lw, lc, flip);
    R.check();
    return;
  }
  const unsigned g = blocks((size_t)N * H * W * C);
  if (K == 3) k_dw<3><<<g, 256, 0, R.s>>>(x, w, b, y, N, lh, lw, lc, flip, add ? *add : none, add_scale, add != nullptr);
  else k_dw<1><<<g, 256, 0, R.s>>>(x, w, b, y, N, lh, lw, lc, flip, add ? *add : none, add_scale, add != nullptr);
  R.check();
}
void dw_wgrad(Run& R, int K, TV x, TV dy, const float* dw, const float* db, int N, int H, int W, int C) {
  if (R.dry) return;
  const size_t NP = (size_t)N * H * W;
  const int lh = ilog2(H), lw = ilog2(W), lc = ilog2(C);
  if (K == 3 && !x.nchw && !dy.nchw && x.ld == C && dy.ld == C && C >= 4 && C <= 512 && W >= 8) {
    const size_t per_block = (size_t)(1024 / C) * 8;   // pixels one block visits per sweep
    const unsigned gv = (unsigned)std::min<size_t>(148 * 8, (NP + per_block - 1) / per_block);
    static const bool v4 = [] { const char* e = getenv("LGTEUN_DW3"); return e && std::string(e) == "v4"; }();
    if (v4) {
      k_dw3_wgrad_v4<<<gv, 256, C * 10 * sizeof(float), R.s>>>(x.p, dy.p, const_cast<float*>(dw), const_cast<float*>(db), N, lh, lw, lc);
    } else {
      const size_t per_block2 = (size_t)(512 / C) * 8;
      const unsigned g2 = (unsigned)std::min<size_t>(148 * 3, (NP + per_block2 - 1) / per_block2);   // resident CTAs only: every CTA ends with C * 10 atomics
      k_dw3_wgrad_v2<<<g2, 256, C * 10 * sizeof(float), R.s>>>(x.p, dy.p, const_cast<float*>(dw), const_cast<float*>(db), N, lh, lw, lc);
    }
    R.check();
    return;
  }
  const unsigned g = (unsigned)std::min<size_t>(148 * 8, (NP + 256 / C - 1) / (256 / C));
  if (K == 3) k_dw_wgrad<3><<<g, 256, C * 10 * sizeof(float), R.s>>>(x, dy, const_cast<float*>(dw), const_cast<float*>(db), N, lh, lw, lc);
  else k_dw_wgrad<1><<<g, 256, C * 2 * sizeof(float), R.s>>>(x, dy, const_cast<float*>(dw), const_cast<float*>(db), N, lh, lw, lc);
  R.check();
}
// bicubic resize of [N,Hi,Wi,C] to [N,Ho,Wo,C]; adjoint: scatter-add of y (gradient) into x
void resize(Run& R, TV x, int Hi, int Wi, TV y, int Ho, int Wo, int C, int N, int adjoint) {
  if (R.dry) return;
  if (!x.nchw && !y.nchw && x.ld == C && y.ld == C && C % 4 == 0) {
    k_resize_v4<<<blocks((size_t)N * Ho * Wo * (C / 4)), 256, 0, R.s>>>(reinterpret_cast<const float4*>(x.p), Hi, Wi,
                                                                      reinterpret_cast<float4*>(y.p), Ho, Wo, C / 4, N,
                                                                      (float)Hi / (float)Ho, adjoint);
    R.check();
    return;
  }
  if (x.nchw && y.nchw && x.C == C && y.C == C && (1 << x.lp) == Hi * Wi && (1 << y.lp) == Ho * Wo) {
    k_resize_nchw<<<blocks((size_t)N * Ho * Wo), 256, 0, R.s>>>(x.p, Hi, Wi, y.p, Ho, Wo, C, N, (float)Hi / (float)Ho, adjoint);
    R.check();
    return;
  }
  k_resize<<<blocks((size_t)N * Ho * Wo * C), 256, 0, R.s>>>(x, Hi, Wi, y, Ho, Wo, C, N, (float)Hi / (float)Ho, adjoint);
  R.check();
}
// pos -> the two pre-scaled layouts the attention kernels read (k_attn_pos); lse: [NP][2] row statistic for the backward
void attn_fwd(Run& R, int D, const float* qkv, const float* pos, float* ppos, TV out, float* lse, int N, int H, int W) {
  if (R.dry) return;
  k_attn_pos<<<32, 256, 0, R.s>>>(pos, ppos);
  const unsigned grid = (unsigned)(N * (H / 8) * (W / 8));
  if (D == 4) k_attn_fwd<4><<<grid, 64, 0, R.s>>>(qkv, ppos, out, lse, N, H, W);
  else if (D == 8) k_attn_fwd<8><<<grid, 64, 0, R.s>>>(qkv, ppos, out, lse, N, H, W);
  else k_attn_fwd<16><<<grid, 64, 0, R.s>>>(qkv, ppos, out, lse, N, H, W);
  R.check();
}
template <int D>
void attn_bwd_launch(Run& R, const float* qkv, const float* ppos, TV dout, TV o, const float* lse, float* dqkv, float* dpos, int N,
                     int H, int W) {
  const size_t smem = sizeof(AttnBwdSmem<D>);
  optin_smem(k_attn_bwd<D>, smem);
  const int per_sm = (int)std::min<size_t>(8, (227 * 1024) / (smem + 1024));
  const unsigned grid = (unsigned)std::min(N * (H / 8) * (W / 8), 148 * per_sm);
  k_attn_bwd<D><<<grid, 64, smem, R.s>>>(qkv, ppos, dout, o, lse, dqkv, dpos, N, H, W);
}
void attn_bwd(Run& R, int D, const float* qkv, const float* ppos, TV dout, TV o, const float* lse, float* dqkv, const float* dpos,
              int N, int H, int W) {
  if (R.dry) return;
  float* dp = const_cast<float*>(dpos);
  if (D == 4) attn_bwd_launch<4>(R, qkv, ppos, dout, o, lse, dqkv, dp, N, H, W);
  else if (D == 8) attn_bwd_launch<8>(R, qkv, ppos, dout, o, lse, dqkv, dp, N, H, W);
  else attn_bwd_launch<16>(R, qkv, ppos, dout, o, lse, dqkv, dp, N, H, W);
  R.check();
}
int fft_cpb(int L, int c2) {    // channels of one line per block: at most 8192 complex points (64 KB) of shared memory
  int cpb = c2;
  while (cpb > 1 && (size_t)cpb * (L + 1) > 8192) cpb /= 2;
  return cpb;
}
void fft_rows(Run& R, int mode, const float* xin, int ldx, const float* sgn, float* spec, float* xout, int ldo, float* xabs,
              int ldabs, int N, int H, int W, int c2, float scale, int weight2) {
  if (R.dry) return;
  // W in {128, 256}: the register-resident row transforms of the inference path (two real channels per complex FFT, fft256.cu)
  static const bool smem_fft = [] { const char* e = getenv("LGTEUN_TRAIN_FFT"); return e && std::string(e) == "smem"; }();
  const size_t rows = (size_t)N * H;
  if (!smem_fft && rows % 4 == 0) {
    cudaError_t e = cudaErrorInvalidValue;
    if (mode == 0 && lg::fft_rows_plain_supported(W, c2, ldx, xin, sgn, spec))
      e = lg::launch_fft_rows_r2c(W, c2, xin, ldx, sgn, spec, rows, scale, weight2 ? 2.f : 1.f, R.s);
    else if (mode == 1 && (!xabs || ldabs % 4 == 0) && lg::fft_rows_plain_supported(W, c2, ldo, spec, xout, xabs))
      e = lg::launch_fft_rows_c2r(W, c2, spec, xout, ldo, xabs, ldabs, rows, scale, weight2 ? 1.f : 0.5f, R.s);
    if (e == cudaSuccess) { ++R.launches; return; }
    (void)cudaGetLastError();
  }
  const int cpb = fft_cpb(W, c2);
  const size_t smem = (W / 2 + (size_t)cpb * (W + 1)) * sizeof(float2);
  optin_smem(k_fft_rows, smem);
  k_fft_rows<<<dim3(H, N, c2 / cpb), 256, smem, R.s>>>(mode, xin, ldx, sgn, reinterpret_cast<float2*>(spec), xout, ldo, xabs,
                                                       ldabs, H, W, ilog2(W), c2, cpb, scale, weight2);
  R.check();
}
void fft_cols(Run& R, float* spec, int N, int H, int W, int c2, int dir, int fixreal) {
  if (R.dry) return;
  static const bool smem_fft = [] { const char* e = getenv("LGTEUN_TRAIN_FFT"); return e && std::string(e) == "smem"; }();
  if (!smem_fft && (H == 128 || H == 256)) {   // the register-resident two-pass column transform of the inference path (fft256.cu)
    cudaError_t e = lg::launch_fft_cols_plain(H, c2, spec, N, W, dir, R.s, fixreal);
    ++R.launches;
    if (R.err == cudaSuccess) R.err = e;
    return;
  }
  const int cpb = fft_cpb(H, c2);
  const size_t smem = (H / 2 + (size_t)cpb * (H + 1)) * sizeof(float2);
  optin_smem(k_fft_cols, smem);
  k_fft_cols<<<dim3(W / 2 + 1, N, c2 / cpb), 256, smem, R.s>>>(reinterpret_cast<float2*>(spec), H, ilog2(H), W, c2, cpb, dir,
                                                               fixreal);
  R.check();
}

struct Step {                   // everything one pass needs
  lgteun_ctx* c;
  TrainState* T;
  const WeightViews* w;         // parameters
  const WeightViews* g;         // gradients (backward only)
  int N;
};

// y = a + drop(b) (mode 0) or drop(b) (mode 1); 16-byte form when the pointers allow it
void dropout(Run& R, const float* a, const float* b, float* y, size_t n, const Step& S, int layer, int mode) {
  if (R.dry) return;
  const float* ext = S.T->use_ext ? S.T->ext_mask[layer] : nullptr;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (n % 4 == 0 && al(a) && al(b) && al(y) && al(ext))
    k_dropout_v4<<<blocks(n / 4), 256, 0, R.s>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                reinterpret_cast<float4*>(y), n / 4, S.T->seed, layer, S.T->p_drop, mode,
                                                reinterpret_cast<const float4*>(ext), S.T->seed_dev);
  else k_dropout<<<blocks(n), 256, 0, R.s>>>(a, b, y, n, S.T->seed, layer, S.T->p_drop, mode, ext, S.T->seed_dev);
  R.check();
}

// ---- one LGB block (LGT.py:231-247): x + drop(proj(cat(local, global)(LN x))), then x + FFN(LN x) -----------------------------
float* fwd_block(Run& R, const Step& S, const BlockW& w, int ch, BlockTape& t, float* Xin, int H, int W, int layer) {
  const int N = S.N, c2 = ch / 2, c4 = 4 * ch, Wh = W / 2 + 1;
  const size_t NP = (size_t)N * H * W, SP = (size_t)N * H * Wh * c2 * 2;
  t.Xin = Xin;
  t.A = R.take(NP * ch);
  ln_fwd(R, nhwc(Xin, ch), ch, w.ln1_w, w.ln1_b, nhwc(t.A, ch), NP);
  // local branch: to_qkv on the first channel half, window attention into cat[:, :c2]
  t.qkv = R.take(NP * 3 * c2);
  pw(R, 0, nhwc(t.A, ch), c2, w.qkv_w, c2, 1, w.qkv_b, nhwc(t.qkv, 3 * c2), 3 * c2, NP);
  t.cat = R.take(NP * ch);
  t.lse = R.take(NP * 2);
  t.ppos = R.take(2 * 2 * 64 * 64);
  attn_fwd(R, c2 / 2, t.qkv, w.pos, t.ppos, nhwc(t.cat, ch), t.lse, N, H, W);
  // global branch: rfft2 -> amplitude / phase mixing -> |irfft2| into cat[:, c2:]
  t.F = R.take(SP);
  fft_rows(R, 0, t.A + c2, ch, nullptr, t.F, nullptr, 0, nullptr, 0, N, H, W, c2, 1.f, 0);
  fft_cols(R, t.F, N, H, W, c2, -1, 1);
  float* G = R.take(SP);
  if (!R.dry) {
    k_spec_mix<<<blocks(SP / 2), 256, 0, R.s>>>(reinterpret_cast<const float2*>(t.F), reinterpret_cast<float2*>(G), SP / 2, c2,
                                                w.amp_w, w.amp_b, w.pha_w, w.pha_b);
    R.check();
  }
  fft_cols(R, G, N, H, W, c2, +1, 0);
  t.v = R.take(NP * c2);
  fft_rows(R, 1, nullptr, 0, nullptr, G, t.v, c2, t.cat + c2, ch, N, H, W, c2, 1.f / ((float)H * (float)W), 1);
  float* pr = R.take(NP * ch);
  if (use_tc_gemm(ch, ch)) tc_pw(R, ch, ch, 0, 0, t.cat, pr, w.proj_w, ch, 1, w.proj_b, nullptr, NP, nullptr);
  else pw(R, 0, nhwc(t.cat, ch), ch, w.proj_w, ch, 1, w.proj_b, nhwc(pr, ch), ch, NP);
  t.Xmid = R.take(NP * ch);
  dropout(R, Xin, pr, t.Xmid, NP * ch, S, layer, 0);
  // conv-FFN (LGT.py:95-109); the GELUs are applied while staging the next conv's input
  t.A2 = R.take(NP * ch);
  ln_fwd(R, nhwc(t.Xmid, ch), ch, w.ln2_w, w.ln2_b, nhwc(t.A2, ch), NP);
  const bool tc = use_tc_gemm(ch, c4);
  t.h1 = R.take(NP * c4);
  if (tc) tc_pw(R, ch, c4, 0, 0, t.A2, t.h1, w.f0_w, ch, 1, w.f0_b, nullptr, NP, nullptr);
  else pw(R, 0, nhwc(t.A2, ch), ch, w.f0_w, ch, 1, w.f0_b, nhwc(t.h1, c4), c4, NP);
  t.h2 = R.take(NP * c4);
  if (tc) tc_pw(R, c4, c4, 2, 0, t.h1, t.h2, w.f1_w, c4, 1, w.f1_b, nullptr, NP, nullptr);
  else pw(R, 1, nhwc(t.h1, c4), c4, w.f1_w, c4, 1, w.f1_b, nhwc(t.h2, c4), c4, NP);
  t.h3 = R.take(NP * c4);
  dwconv(R, 3, nhwc(t.h2, c4), w.dw_w, w.dw_b, nhwc(t.h3, c4), N, H, W, c4, 0);
  t.Xout = R.take(NP * ch);
  const TV res = nhwc(t.Xmid, ch);
  if (tc) tc_pw(R, c4, ch, 2, 2, t.h3, t.Xout, w.f2_w, c4, 1, w.f2_b, t.Xmid, NP, nullptr);
  else pw(R, 1, nhwc(t.h3, c4), c4, w.f2_w, c4, 1, w.f2_b, nhwc(t.Xout, ch), ch, NP, &res);
  return t.Xout;
}

// gX: gradient of the block output on entry, of the block input on return (in place)
void bwd_block(Run& R, const Step& S, const BlockW& w, const BlockW& g, int ch, const BlockTape& t, float* gX, int H, int W,
               int layer) {
  const int N = S.N, c2 = ch / 2, c4 = 4 * ch, Wh = W / 2 + 1;
  const size_t NP = (size_t)N * H * W, SP = (size_t)N * H * Wh * c2 * 2;
  const TV gx = nhwc(gX, ch);
  // FFN
  pw_wgrad(R, 1, nhwc(t.h3, c4), c4, gx, ch, g.f2_w, c4, 1, g.f2_b, NP, S.T->gscale);
  const bool tc = use_tc_gemm(ch, c4);
  const float* gs = S.T->gscale;
  float* dh3 = R.take(NP * c4);
  const TV gate3 = nhwc(t.h3, c4);
  if (tc) tc_pw(R, ch, c4, 0, 3, gX, dh3, w.f2_w, 1, c4, nullptr, t.h3, NP, gs);
  else pw(R, 0, gx, ch, w.f2_w, 1, c4, nullptr, nhwc(dh3, c4), c4, NP, nullptr, &gate3);
  dw_wgrad(R, 3, nhwc(t.h2, c4), nhwc(dh3, c4), g.dw_w, g.dw_b, N, H, W, c4);
  float* dh2 = R.take(NP * c4);
  dwconv(R, 3, nhwc(dh3, c4), w.dw_w, nullptr, nhwc(dh2, c4), N, H, W, c4, 1);
  pw_wgrad(R, 1, nhwc(t.h1, c4), c4, nhwc(dh2, c4), c4, g.f1_w, c4, 1, g.f1_b, NP, S.T->gscale);
  float* dh1 = R.take(NP * c4);
  const TV gate1 = nhwc(t.h1, c4);
  if (tc) tc_pw(R, c4, c4, 0, 3, dh2, dh1, w.f1_w, 1, c4, nullptr, t.h1, NP, gs);
  else pw(R, 0, nhwc(dh2, c4), c4, w.f1_w, 1, c4, nullptr, nhwc(dh1, c4), c4, NP, nullptr, &gate1);
  pw_wgrad(R, 0, nhwc(t.A2, ch), ch, nhwc(dh1, c4), c4, g.f0_w, ch, 1, g.f0_b, NP, S.T->gscale);
  float* dA2 = R.take(NP * ch);
  if (tc) tc_pw(R, c4, ch, 0, 0, dh1, dA2, w.f0_w, 1, ch, nullptr, nullptr, NP, gs);
  else pw(R, 0, nhwc(dh1, c4), c4, w.f0_w, 1, ch, nullptr, nhwc(dA2, ch), ch, NP);
  ln_bwd(R, nhwc(t.Xmid, ch), ch, w.ln2_w, nhwc(dA2, ch), gx, 1, g.ln2_w, g.ln2_b, NP);
  // mixer: gX is now the gradient of Xmid = Xin + drop(proj(cat))
  float* dpr = gX;
  if (S.T->p_drop > 0.f || S.T->use_ext) {
    dpr = R.take(NP * ch);
    dropout(R, nullptr, gX, dpr, NP * ch, S, layer, 1);
  }
  pw_wgrad(R, 0, nhwc(t.cat, ch), ch, nhwc(dpr, ch), ch, g.proj_w, ch, 1, g.proj_b, NP, S.T->gscale);
  float* dcat = R.take(NP * ch);
  if (use_tc_gemm(ch, ch)) tc_pw(R, ch, ch, 0, 0, dpr, dcat, w.proj_w, 1, ch, nullptr, nullptr, NP, S.T->gscale);
  else pw(R, 0, nhwc(dpr, ch), ch, w.proj_w, 1, ch, nullptr, nhwc(dcat, ch), ch, NP);
  float* dA = R.take(NP * ch);
  // local branch
  float* dqkv = R.take(NP * 3 * c2);
  attn_bwd(R, c2 / 2, t.qkv, t.ppos, nhwc(dcat, ch), nhwc(t.cat, ch), t.lse, dqkv, g.pos, N, H, W);
  pw_wgrad(R, 0, nhwc(t.A, ch), c2, nhwc(dqkv, 3 * c2), 3 * c2, g.qkv_w, c2, 1, g.qkv_b, NP, S.T->gscale);
  pw(R, 0, nhwc(dqkv, 3 * c2), 3 * c2, w.qkv_w, 1, c2, nullptr, nhwc(dA, ch), c2, NP);
  // global branch: |.| -> C2R rows -> inverse columns -> mixing -> forward columns -> R2C rows, all transposed
  float* dG = R.take(SP);
  fft_rows(R, 0, dcat + c2, ch, t.v, dG, nullptr, 0, nullptr, 0, N, H, W, c2, 1.f / ((float)H * (float)W), 1);
  fft_cols(R, dG, N, H, W, c2, -1, 0);
  if (!R.dry) {
    const unsigned gb = (unsigned)std::min<size_t>(148 * 8, (SP / 2 + 255) / 256);
    k_spec_mix_bwd<<<gb, 256, 4 * c2 * sizeof(float), R.s>>>(
        reinterpret_cast<const float2*>(t.F), reinterpret_cast<float2*>(dG), SP / 2, c2, w.amp_w, w.amp_b, w.pha_w, w.pha_b,
        const_cast<float*>(g.amp_w), const_cast<float*>(g.amp_b), const_cast<float*>(g.pha_w), const_cast<float*>(g.pha_b));
    R.check();
  }
  fft_cols(R, dG, N, H, W, c2, +1, 0);
  fft_rows(R, 1, nullptr, 0, nullptr, dG, dA + c2, ch, nullptr, 0, N, H, W, c2, 1.f, 0);
  ln_bwd(R, nhwc(t.Xin, ch), ch, w.ln1_w, nhwc(dA, ch), gx, 1, g.ln1_w, g.ln1_b, NP);
}

// ---- LGT (LGT.py:314-344) ---------------------------------------------------------------------------------------------------------
void fwd_prior(Run& R, const Step& S, const PriorW& w, PriorTape& t, const float* zin, float* out, int H, int W) {
  const int N = S.N, B = S.c->B, C = S.c->C, H2 = H / 2, W2 = W / 2;
  const size_t NP = (size_t)N * H * W, NP2 = NP / 4;
  const int P = H * W;
  t.zin = zin;
  t.t0 = R.take(NP * B);
  dwconv(R, 1, nchw(zin, B, P), w.pe_dw_w, w.pe_dw_b, nhwc(t.t0, B), N, H, W, B, 0);
  t.t1 = R.take(NP * C);
  pw(R, 0, nhwc(t.t0, B), B, w.pe_w, B, 1, w.pe_b, nhwc(t.t1, C), C, NP);
  t.X0 = R.take(NP * C);
  ln_fwd(R, nhwc(t.t1, C), C, w.pe_ln_w, w.pe_ln_b, nhwc(t.X0, C), NP);
  float* x = t.X0;
  for (int j = 0; j < 2; ++j) x = fwd_block(R, S, w.enc[j], C, t.enc[j], x, H, W, j);
  float* skip = x;
  t.d0 = R.take(NP2 * C);
  resize(R, nhwc(skip, C), H, W, nhwc(t.d0, C), H2, W2, C, N, 0);
  t.L0 = R.take(NP2 * 2 * C);
  pw(R, 0, nhwc(t.d0, C), C, w.down_w, C, 1, w.down_b, nhwc(t.L0, 2 * C), 2 * C, NP2);
  float* low = fwd_block(R, S, w.bott[0], 2 * C, t.bott, t.L0, H2, W2, 2);
  t.u0 = R.take(NP * 2 * C);
  resize(R, nhwc(low, 2 * C), H2, W2, nhwc(t.u0, 2 * C), H, W, 2 * C, N, 0);
  t.u1 = R.take(NP * C);
  pw(R, 0, nhwc(t.u0, 2 * C), 2 * C, w.up_w, 2 * C, 1, w.up_b, nhwc(t.u1, C), C, NP);
  t.fz = R.take(NP * C);                                   // fusion conv over cat([upsampled, skip]) as two half-convs
  pw(R, 0, nhwc(t.u1, C), C, w.fuse_w, 2 * C, 1, w.fuse_b, nhwc(t.fz, C), C, NP);
  const TV acc = nhwc(t.fz, C);
  pw(R, 0, nhwc(skip, C), C, w.fuse_w + C, 2 * C, 1, nullptr, nhwc(t.fz, C), C, NP, &acc);
  x = t.fz;
  for (int j = 0; j < 2; ++j) x = fwd_block(R, S, w.dec[j], C, t.dec[j], x, H, W, 3 + j);
  t.fea = x;
  const TV zres = nchw(zin, B, P);                         // tail: bicubic x1 is the identity; + global residual (LGT.py:342)
  pw(R, 0, nhwc(t.fea, C), C, w.tail_w, C, 1, w.tail_b, nchw(out, B, P), B, NP, &zres);
}

// dout: gradient of the prior output (NCHW); gZ (NCHW) receives += the gradient wrt the prior input, and must already hold
// whatever else flows into Z (nothing for the last stage: it is set to dout here, the residual path)
void bwd_prior(Run& R, const Step& S, const PriorW& w, const PriorW& g, const PriorTape& t, const float* dout, float* gZ, int H,
               int W) {
  const int N = S.N, B = S.c->B, C = S.c->C, H2 = H / 2, W2 = W / 2;
  const size_t NP = (size_t)N * H * W, NP2 = NP / 4;
  const int P = H * W;
  if (!R.dry && R.err == cudaSuccess)
    R.err = cudaMemcpyAsync(gZ, dout, NP * B * sizeof(float), cudaMemcpyDeviceToDevice, R.s);
  pw_wgrad(R, 0, nhwc(t.fea, C), C, nchw(dout, B, P), B, g.tail_w, C, 1, g.tail_b, NP);
  float* gX = R.take(NP * C);
  pw(R, 0, nchw(dout, B, P), B, w.tail_w, 1, C, nullptr, nhwc(gX, C), C, NP);
  for (int j = 1; j >= 0; --j) bwd_block(R, S, w.dec[j], g.dec[j], C, t.dec[j], gX, H, W, 3 + j);
  const float* skip = t.enc[1].Xout;
  pw_wgrad(R, 0, nhwc(t.u1, C), C, nhwc(gX, C), C, g.fuse_w, 2 * C, 1, g.fuse_b, NP, S.T->gscale);
  pw_wgrad(R, 0, nhwc(skip, C), C, nhwc(gX, C), C, g.fuse_w + C, 2 * C, 1, nullptr, NP, S.T->gscale);
  float* du1 = R.take(NP * C);
  pw(R, 0, nhwc(gX, C), C, w.fuse_w, 1, 2 * C, nullptr, nhwc(du1, C), C, NP);
  float* gSkip = R.take(NP * C);
  pw(R, 0, nhwc(gX, C), C, w.fuse_w + C, 1, 2 * C, nullptr, nhwc(gSkip, C), C, NP);
  pw_wgrad(R, 0, nhwc(t.u0, 2 * C), 2 * C, nhwc(du1, C), C, g.up_w, 2 * C, 1, g.up_b, NP, S.T->gscale);
  float* du0 = R.take(NP * 2 * C);
  pw(R, 0, nhwc(du1, C), C, w.up_w, 1, 2 * C, nullptr, nhwc(du0, 2 * C), 2 * C, NP);
  float* gL = R.take(NP2 * 2 * C);
  R.zero(gL, NP2 * 2 * C);
  resize(R, nhwc(gL, 2 * C), H2, W2, nhwc(du0, 2 * C), H, W, 2 * C, N, 1);
  bwd_block(R, S, w.bott[0], g.bott[0], 2 * C, t.bott, gL, H2, W2, 2);
  pw_wgrad(R, 0, nhwc(t.d0, C), C, nhwc(gL, 2 * C), 2 * C, g.down_w, C, 1, g.down_b, NP2, S.T->gscale);
  float* dd0 = R.take(NP2 * C);
  pw(R, 0, nhwc(gL, 2 * C), 2 * C, w.down_w, 1, C, nullptr, nhwc(dd0, C), C, NP2);
  resize(R, nhwc(gSkip, C), H, W, nhwc(dd0, C), H2, W2, C, N, 1);
  for (int j = 1; j >= 0; --j) bwd_block(R, S, w.enc[j], g.enc[j], C, t.enc[j], gSkip, H, W, j);
  float* dt1 = R.take(NP * C);
  ln_bwd(R, nhwc(t.t1, C), C, w.pe_ln_w, nhwc(gSkip, C), nhwc(dt1, C), 0, g.pe_ln_w, g.pe_ln_b, NP);
  pw_wgrad(R, 0, nhwc(t.t0, B), B, nhwc(dt1, C), C, g.pe_w, B, 1, g.pe_b, NP);
  float* dt0 = R.take(NP * B);
  pw(R, 0, nhwc(dt1, C), C, w.pe_w, 1, B, nullptr, nhwc(dt0, B), B, NP);
  dw_wgrad(R, 1, nchw(t.zin, B, P), nhwc(dt0, B), g.pe_dw_w, g.pe_dw_b, N, H, W, B);
  const TV acc = nchw(gZ, B, P);
  dwconv(R, 1, nhwc(dt0, B), w.pe_dw_w, nullptr, nchw(gZ, B, P), N, H, W, B, 0, &acc, 1.f);
}

// ---- data module step (unlg_former.py:58-61) -------------------------------------------------------------------------------------
float* fwd_data(Run& R, const Step& S, const DataW& w, int stage, DataTape& t, const float* Zin, const float* ms,
                const float* pan, int h, int wd) {
  const int N = S.N, B = S.c->B, H = 4 * h, W = 4 * wd, H2 = 2 * h, W2 = 2 * wd;
  const size_t P = (size_t)H * W, P2 = P / 4, P4 = P / 16;
  t.Zin = Zin;
  t.a1 = R.take(N * B * P2);
  resize(R, nchw(Zin, B, (int)P), H, W, nchw(t.a1, B, (int)P2), H2, W2, B, N, 0);
  t.a2 = R.take(N * B * P2);
  dwconv(R, 3, nchw(t.a1, B, (int)P2), w.d1_w, w.d1_b, nchw(t.a2, B, (int)P2), N, H2, W2, B, 0);
  t.a3 = R.take(N * B * P4);
  resize(R, nchw(t.a2, B, (int)P2), H2, W2, nchw(t.a3, B, (int)P4), h, wd, B, N, 0);
  t.e = R.take(N * B * P4);
  const TV msv = nchw(ms, B, (int)P4);
  dwconv(R, 3, nchw(t.a3, B, (int)P4), w.d3_w, w.d3_b, nchw(t.e, B, (int)P4), N, h, wd, B, 0, &msv, -1.f);
  t.u1 = R.take(N * B * P2);
  resize(R, nchw(t.e, B, (int)P4), h, wd, nchw(t.u1, B, (int)P2), H2, W2, B, N, 0);
  t.u2 = R.take(N * B * P2);
  dwconv(R, 3, nchw(t.u1, B, (int)P2), w.dt1_w, w.dt1_b, nchw(t.u2, B, (int)P2), N, H2, W2, B, 0);
  t.u3 = R.take(N * B * P);
  resize(R, nchw(t.u2, B, (int)P2), H2, W2, nchw(t.u3, B, (int)P), H, W, B, N, 0);
  t.T1 = R.take(N * B * P);
  dwconv(R, 3, nchw(t.u3, B, (int)P), w.dt3_w, w.dt3_b, nchw(t.T1, B, (int)P), N, H, W, B, 0);
  t.r = R.take(N * P);
  t.Zout = R.take(N * B * P);
  if (!R.dry) {
    k_data_r<<<blocks(N * P), 256, 0, R.s>>>(Zin, pan, w.r_w, w.r_b, t.r, N, B, P);
    R.check();
    k_data_update<<<blocks(N * B * P), 256, 0, R.s>>>(Zin, t.T1, t.r, w.rt_w, w.rt_b, w.eta[stage], t.Zout, N, B, P);
    R.check();
  }
  return t.Zout;
}
// gz: dZout on entry, dZin on return (in place)
void bwd_data(Run& R, const Step& S, const DataW& w, const DataW& g, int stage, const DataTape& t, float* gz, int h, int wd) {
  const int N = S.N, B = S.c->B, H = 4 * h, W = 4 * wd, H2 = 2 * h, W2 = 2 * wd;
  const size_t P = (size_t)H * W, P2 = P / 4, P4 = P / 16;
  float* s = R.take(N * B * P);
  if (!R.dry) {
    k_data_update_bwd<<<blocks(N * P), 256, 0, R.s>>>(gz, t.Zin, t.T1, t.r, w.r_w, w.rt_w, w.rt_b, w.eta[stage], s,
                                                      const_cast<float*>(g.eta[stage]), const_cast<float*>(g.rt_w),
                                                      const_cast<float*>(g.rt_b), const_cast<float*>(g.r_w),
                                                      const_cast<float*>(g.r_b), N, B, P);
    R.check();
  }
  dw_wgrad(R, 3, nchw(t.u3, B, (int)P), nchw(s, B, (int)P), g.dt3_w, g.dt3_b, N, H, W, B);
  float* du3 = R.take(N * B * P);
  dwconv(R, 3, nchw(s, B, (int)P), w.dt3_w, nullptr, nchw(du3, B, (int)P), N, H, W, B, 1);
  float* du2 = R.take(N * B * P2);
  R.zero(du2, N * B * P2);
  resize(R, nchw(du2, B, (int)P2), H2, W2, nchw(du3, B, (int)P), H, W, B, N, 1);
  dw_wgrad(R, 3, nchw(t.u1, B, (int)P2), nchw(du2, B, (int)P2), g.dt1_w, g.dt1_b, N, H2, W2, B);
  float* du1 = R.take(N * B * P2);
  dwconv(R, 3, nchw(du2, B, (int)P2), w.dt1_w, nullptr, nchw(du1, B, (int)P2), N, H2, W2, B, 1);
  float* de = R.take(N * B * P4);
  R.zero(de, N * B * P4);
  resize(R, nchw(de, B, (int)P4), h, wd, nchw(du1, B, (int)P2), H2, W2, B, N, 1);
  dw_wgrad(R, 3, nchw(t.a3, B, (int)P4), nchw(de, B, (int)P4), g.d3_w, g.d3_b, N, h, wd, B);
  float* da3 = R.take(N * B * P4);
  dwconv(R, 3, nchw(de, B, (int)P4), w.d3_w, nullptr, nchw(da3, B, (int)P4), N, h, wd, B, 1);
  float* da2 = R.take(N * B * P2);
  R.zero(da2, N * B * P2);
  resize(R, nchw(da2, B, (int)P2), H2, W2, nchw(da3, B, (int)P4), h, wd, B, N, 1);
  dw_wgrad(R, 3, nchw(t.a1, B, (int)P2), nchw(da2, B, (int)P2), g.d1_w, g.d1_b, N, H2, W2, B);
  float* da1 = R.take(N * B * P2);
  dwconv(R, 3, nchw(da2, B, (int)P2), w.d1_w, nullptr, nchw(da1, B, (int)P2), N, H2, W2, B, 1);
  resize(R, nchw(gz, B, (int)P), H, W, nchw(da1, B, (int)P2), H2, W2, B, N, 1);
}

// ---- the whole step ------------------------------------------------------------------------------------------------------------------
void run_fwd(Run& R, const Step& S, const float* ms, const float* pan, float* out, int h, int w) {
  const int N = S.N, B = S.c->B, K = S.c->K, H = 4 * h, W = 4 * w;
  float* z = R.take((size_t)N * B * H * W);                // Z0 = bicubic x4 (unlg_former.py:53)
  resize(R, nchw(ms, B, h * w), h, w, nchw(z, B, H * W), H, W, B, N, 0);
  for (int i = 0; i < K; ++i) z = fwd_data(R, S, S.w->dw, i, S.T->data[i], z, ms, pan, h, w);
  fwd_prior(R, S, S.w->prior[K - 1], S.T->prior, z, out, H, W);
}
void run_bwd(Run& R, const Step& S, const float* dout, int h, int w) {
  const int N = S.N, B = S.c->B, K = S.c->K, H = 4 * h, W = 4 * w;
  float* gZ = R.take((size_t)N * B * H * W);
  S.T->gscale = R.take(64);
  if (!R.dry) {
    R.zero(S.T->gscale, 64);
    k_absmax<<<(unsigned)std::min<size_t>(148 * 4, ((size_t)N * B * H * W + 255) / 256), 256, 0, R.s>>>(
        dout, (size_t)N * B * H * W, reinterpret_cast<unsigned*>(S.T->gscale + 1));
    R.check();
    k_pow2_scale<<<1, 1, 0, R.s>>>(reinterpret_cast<const unsigned*>(S.T->gscale + 1), S.T->gscale);
    R.check();
  }
  bwd_prior(R, S, S.w->prior[K - 1], S.g->prior[K - 1], S.T->prior, dout, gZ, H, W);
  for (int i = K - 1; i >= 0; --i) bwd_data(R, S, S.w->dw, S.g->dw, i, S.T->data[i], gZ, h, w);
}

void bind_views(const lgteun_ctx* c, const float* base, WeightViews* out) {
  memset(out, 0, sizeof(*out));
  for (const auto& s : c->slots) {
    const size_t member = reinterpret_cast<const char*>(s.slot) - reinterpret_cast<const char*>(&c->wv);
    *reinterpret_cast<const float**>(reinterpret_cast<char*>(out) + member) = base + s.offset;
  }
}

bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
int check_train_shape(int N, int h, int w) {
  if (N <= 0) return fail(LGTEUN_EINVAL, "batch must be positive");
  if (h < 4 || w < 4 || !pow2(h) || !pow2(w) || 4 * h > 1024 || 4 * w > 1024)
    return fail(LGTEUN_EINVAL, "unsupported shape: PAN height/width (4h, 4w) must be powers of two in [16, 1024]");
  return 0;
}

size_t plan_bytes(lgteun_ctx* c, TrainState* T, int N, int h, int w) {
  Run R;
  R.base = nullptr;
  R.dry = true;
  WeightViews wv;
  memset(&wv, 0, sizeof(wv));
  TrainState scratch = *T;
  Step S{c, &scratch, &wv, &wv, N};
  run_fwd(R, S, nullptr, nullptr, nullptr, h, w);
  run_bwd(R, S, nullptr, h, w);
  return R.off * sizeof(float);
}

}  // namespace

// The generic power-of-two FFT passes, exported for the companion operators of companion_ops.cu.
namespace lgctx {
cudaError_t launch_fft_rows(cudaStream_t s, int mode, const float* xin, int ldx, float* spec, float* xout, int ldo, float* xabs,
                            int ldabs, int N, int H, int W, int c2, float scale, int weight2) {
  Run R{nullptr};
  R.s = s;
  fft_rows(R, mode, xin, ldx, nullptr, spec, xout, ldo, xabs, ldabs, N, H, W, c2, scale, weight2);
  return R.err;
}
cudaError_t launch_fft_cols(cudaStream_t s, float* spec, int N, int H, int W, int c2, int dir, int fixreal) {
  Run R{nullptr};
  R.s = s;
  fft_cols(R, spec, N, H, W, c2, dir, fixreal);
  return R.err;
}
}  // namespace lgctx

extern "C" {

int64_t lgteun_flat_numel(const lgteun_t* c) { return c ? (int64_t)c->flat_floats : -1; }
int64_t lgteun_weight_offset(const lgteun_t* c, int i) {
  return (c && i >= 0 && i < (int)c->slots.size()) ? (int64_t)c->slots[i].offset : -1;
}

int64_t lgteun_train_workspace_bytes(lgteun_t* c, int N, int h, int w) {
  if (!c || check_train_shape(N, h, w)) return -1;
  TrainState T;
  T.p_drop = 0.1f;
  return (int64_t)plan_bytes(c, &T, N, h, w);
}

int lgteun_train_forward(lgteun_t* c, const float* flat_param, const float* ms, const float* pan, float* out, int N, int h,
                         int w, float dropout_p, uint64_t seed, void* stream) {
  if (!c || !flat_param || !ms || !pan || !out) return fail(LGTEUN_EINVAL, "NULL argument");
  if (!(dropout_p >= 0.f && dropout_p < 1.f)) return fail(LGTEUN_EINVAL, "dropout_p must be in [0, 1)");
  int rc = check_train_shape(N, h, w);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  if (!c->train) c->train = new TrainState();
  TrainState* T = c->train;
  T->valid = false;
  T->p_drop = dropout_p;
  const size_t need = plan_bytes(c, T, N, h, w);
  if (need > T->cap) {
    CK(cudaDeviceSynchronize());
    if (T->base) cudaFree(T->base);
    T->base = nullptr;
    T->cap = 0;
    cudaError_t e = cudaMalloc(&T->base, need);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(LGTEUN_ENOMEM, "training tape cudaMalloc of " + std::to_string(need) + " bytes failed: " + cudaGetErrorString(e));
    }
    T->cap = need;
  }
  T->N = N; T->h = h; T->w = w;
  T->p_drop = dropout_p;
  T->seed = seed;
  T->flat_param = flat_param;
  T->ms = ms; T->pan = pan;
  WeightViews wv;
  bind_views(c, flat_param, &wv);
  Run R;
  R.base = T->base;
  R.s = (cudaStream_t)stream;
  Step S{c, T, &wv, nullptr, N};
  run_fwd(R, S, ms, pan, out, h, w);
  if (R.err != cudaSuccess) return fail_cuda(R.err, "training forward launch");
  T->fwd_end = R.off;
  T->launches_fwd = R.launches;
  T->valid = true;
  ++T->generation;
  return 0;
}

uint64_t lgteun_train_generation(const lgteun_t* c) { return (c && c->train) ? c->train->generation : 0; }

int lgteun_train_backward_of(lgteun_t* c, uint64_t generation, const float* dout, float* flat_grad, void* stream) {
  if (!c) return fail(LGTEUN_EINVAL, "NULL argument");
  TrainState* T = c->train;
  if (!T || !T->valid || T->generation != generation)
    return fail(LGTEUN_ESTATE, "the activation tape belongs to another forward: a handle keeps ONE tape, so the forward of generation " +
                                   std::to_string(generation) + " was overwritten by a later lgteun_train_forward (or already consumed)");
  return lgteun_train_backward(c, dout, flat_grad, stream);
}

int lgteun_train_backward(lgteun_t* c, const float* dout, float* flat_grad, void* stream) {
  if (!c || !dout || !flat_grad) return fail(LGTEUN_EINVAL, "NULL argument");
  TrainState* T = c->train;
  if (!T || !T->valid) return fail(LGTEUN_ESTATE, "no tape: call lgteun_train_forward first (one backward per forward)");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemsetAsync(flat_grad, 0, c->flat_floats * sizeof(float), s));
  WeightViews wv, gv;
  bind_views(c, T->flat_param, &wv);
  bind_views(c, flat_grad, &gv);
  Run R;
  R.base = T->base;
  R.off = T->fwd_end;
  R.s = s;
  Step S{c, T, &wv, &gv, T->N};
  run_bwd(R, S, dout, T->h, T->w);
  T->valid = false;
  if (R.err != cudaSuccess) return fail_cuda(R.err, "training backward launch");
  T->launches_bwd = R.launches;
  return 0;
}

int lgteun_train_set_masks(lgteun_t* c, const float* const* masks) {
  if (!c) return fail(LGTEUN_EINVAL, "NULL handle");
  if (!c->train) c->train = new TrainState();
  c->train->use_ext = masks != nullptr;
  for (int i = 0; i < 5; ++i) {
    if (masks && !masks[i]) return fail(LGTEUN_EINVAL, "five mask pointers are required");
    c->train->ext_mask[i] = masks ? masks[i] : nullptr;
  }
  return 0;
}

// seed_dev != NULL: the dropout kernels of the following forwards / backwards read their seed from *seed_dev (device memory) when
// they run, so a CUDA graph of the step can be replayed with a new seed; NULL restores the by-value seed of lgteun_train_forward.
int lgteun_train_set_seed_ptr(lgteun_t* c, const uint64_t* seed_dev) {
  if (!c) return fail(LGTEUN_EINVAL, "NULL handle");
  if (!c->train) c->train = new TrainState();
  c->train->seed_dev = seed_dev;
  return 0;
}

int lgteun_train_launches(const lgteun_t* c) {
  return (c && c->train) ? c->train->launches_fwd + c->train->launches_bwd : 0;
}

int lgteun_l1_loss(lgteun_t* c, const float* out, const float* gt, int64_t n, float weight, float* loss_dev, float* dout,
                   void* stream) {
  if (!c || !out || !gt || !loss_dev) return fail(LGTEUN_EINVAL, "NULL argument");
  if (n <= 0) return fail(LGTEUN_EINVAL, "empty tensor");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemsetAsync(loss_dev, 0, sizeof(float), s));
  const unsigned g = (unsigned)std::min<int64_t>(148 * 8, (n + 255) / 256);
  k_l1<<<g, 256, 0, s>>>(out, gt, (size_t)n, weight / (float)n, loss_dev, dout);
  CK(cudaGetLastError());
  return 0;
}

int lgteun_adam_step(lgteun_t* c, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  if (!c || !param || !grad || !exp_avg || !exp_avg_sq) return fail(LGTEUN_EINVAL, "NULL argument");
  if (n <= 0 || step < 1) return fail(LGTEUN_EINVAL, "n must be positive and step >= 1");
  CK(cudaSetDevice(c->device));
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  k_adam<<<blocks((size_t)n), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, (size_t)n, lr, beta1, beta2, eps,
                                                             bc1, sqrtf(bc2), grad_scale);
  CK(cudaGetLastError());
  return 0;
}

int lgteun_dropout_mask(lgteun_t* c, uint64_t seed, int layer, float p, float* mask_out, int64_t n, void* stream) {
  if (!c || !mask_out || n <= 0) return fail(LGTEUN_EINVAL, "bad argument");
  CK(cudaSetDevice(c->device));
  k_dropout<<<blocks((size_t)n), 256, 0, (cudaStream_t)stream>>>(nullptr, nullptr, mask_out, (size_t)n, seed, layer, p, 2, nullptr);
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"
