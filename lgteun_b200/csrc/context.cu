// context.cu — the C ABI of include/lgteun.h: handle, weight arena, workspace, stage chaining
// (models/unlg_former.py:50-67 and models/common/LGT.py:314-344) and the CUDA-graph cache.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "ctx.cuh"

using namespace lg;

using namespace lgctx;

namespace lgctx {
std::string& err_slot() {
  thread_local std::string e;
  return e;
}
}  // namespace lgctx
#define g_err (lgctx::err_slot())

namespace {

size_t align4(size_t v) { return (v + 3) & ~(size_t)3; }

void add_slot(lgteun_ctx* c, const std::string& name, int64_t numel, const float** slot) {
  c->slots.push_back({name, numel, slot, 0});
}
void add_block(lgteun_ctx* c, const std::string& p, BlockW* b, int ch) {
  const int c2 = ch / 2, c4 = 4 * ch;
  add_slot(c, p + ".0.fn.norm.weight", ch, &b->ln1_w);
  add_slot(c, p + ".0.fn.norm.bias", ch, &b->ln1_b);
  add_slot(c, p + ".0.fn.fn.local_mixer.pos_emb", 2 * 64 * 64, &b->pos);
  add_slot(c, p + ".0.fn.fn.local_mixer.to_qkv.weight", 3 * c2 * c2, &b->qkv_w);
  add_slot(c, p + ".0.fn.fn.local_mixer.to_qkv.bias", 3 * c2, &b->qkv_b);
  add_slot(c, p + ".0.fn.fn.global_mixer.conv_amp.0.weight", c2, &b->amp_w);
  add_slot(c, p + ".0.fn.fn.global_mixer.conv_amp.0.bias", c2, &b->amp_b);
  add_slot(c, p + ".0.fn.fn.global_mixer.conv_pha.0.weight", c2, &b->pha_w);
  add_slot(c, p + ".0.fn.fn.global_mixer.conv_pha.0.bias", c2, &b->pha_b);
  add_slot(c, p + ".0.fn.fn.proj.weight", ch * ch, &b->proj_w);
  add_slot(c, p + ".0.fn.fn.proj.bias", ch, &b->proj_b);
  add_slot(c, p + ".1.fn.norm.weight", ch, &b->ln2_w);
  add_slot(c, p + ".1.fn.norm.bias", ch, &b->ln2_b);
  add_slot(c, p + ".1.fn.fn.net.0.weight", c4 * ch, &b->f0_w);
  add_slot(c, p + ".1.fn.fn.net.0.bias", c4, &b->f0_b);
  add_slot(c, p + ".1.fn.fn.net.2.point_conv.weight", c4 * c4, &b->f1_w);
  add_slot(c, p + ".1.fn.fn.net.2.point_conv.bias", c4, &b->f1_b);
  add_slot(c, p + ".1.fn.fn.net.2.depth_conv.weight", c4 * 9, &b->dw_w);
  add_slot(c, p + ".1.fn.fn.net.2.depth_conv.bias", c4, &b->dw_b);
  add_slot(c, p + ".1.fn.fn.net.4.weight", ch * c4, &b->f2_w);
  add_slot(c, p + ".1.fn.fn.net.4.bias", ch, &b->f2_b);
  c->derived.push_back({&b->pos, &b->pos_t, 0, 0, 0, nullptr});
  c->derived.push_back({&b->f0_w, &b->f0_wt, c4, ch, 0, nullptr});
  c->derived.push_back({&b->f1_w, &b->f1_wt, c4, c4, 0, nullptr});
  c->derived.push_back({&b->f2_w, &b->f2_wt, ch, c4, 0, nullptr});
  if (ch == 16 || ch == 32 || ch == 64) c->derived.push_back({&b->f0_w, &b->ffn_pack, -1, ch, 0, b});
  if (ch == 16 || ch == 32) c->derived.push_back({&b->f0_w, &b->ffn_cl_pack, -3, ch, 0, b});
  c->derived.push_back({&b->proj_w, &b->proj_pack, -2, ch, 0, b});      // proj [c][c] as fp16 hi | lo
}

// The weight ABI: reference state_dict key grammar.
void build_table(lgteun_ctx* c) {
  const int B = c->B, C = c->C;
  const char* dn[4] = {"D.1", "D.3", "DT.1", "DT.3"};
  const float** dwp[4] = {&c->wv.dw.d1_w, &c->wv.dw.d3_w, &c->wv.dw.dt1_w, &c->wv.dw.dt3_w};
  const float** dbp[4] = {&c->wv.dw.d1_b, &c->wv.dw.d3_b, &c->wv.dw.dt1_b, &c->wv.dw.dt3_b};
  for (int i = 0; i < 4; ++i) {
    add_slot(c, std::string(dn[i]) + ".weight", B * 9, dwp[i]);
    add_slot(c, std::string(dn[i]) + ".bias", B, dbp[i]);
  }
  add_slot(c, "R.weight", B, &c->wv.dw.r_w);
  add_slot(c, "R.bias", 1, &c->wv.dw.r_b);
  add_slot(c, "RT.weight", B, &c->wv.dw.rt_w);
  add_slot(c, "RT.bias", B, &c->wv.dw.rt_b);
  for (int i = 0; i < c->K; ++i) add_slot(c, "eta." + std::to_string(i), 1, &c->wv.dw.eta[i]);
  for (int i = 0; i < c->K; ++i) {
    PriorW* p = &c->wv.prior[i];
    const std::string pre = "prior_module." + std::to_string(i);
    add_slot(c, pre + ".patch_embed.proj.0.weight", B, &p->pe_dw_w);
    add_slot(c, pre + ".patch_embed.proj.0.bias", B, &p->pe_dw_b);
    add_slot(c, pre + ".patch_embed.proj.1.weight", C * B, &p->pe_w);
    add_slot(c, pre + ".patch_embed.proj.1.bias", C, &p->pe_b);
    add_slot(c, pre + ".patch_embed.norm.weight", C, &p->pe_ln_w);
    add_slot(c, pre + ".patch_embed.norm.bias", C, &p->pe_ln_b);
    for (int j = 0; j < 2; ++j) add_block(c, pre + ".encoder_layers.0.0.blocks." + std::to_string(j), &p->enc[j], C);
    add_slot(c, pre + ".encoder_layers.0.1.1.weight", 2 * C * C, &p->down_w);
    add_slot(c, pre + ".encoder_layers.0.1.1.bias", 2 * C, &p->down_b);
    c->derived.push_back({&p->down_w, &p->down_wt, 2 * C, C, 0, nullptr});
    add_block(c, pre + ".bottleneck.blocks.0", &p->bott[0], 2 * C);
    add_slot(c, pre + ".decoder_layers.0.0.1.weight", 2 * C * C, &p->up_w);
    add_slot(c, pre + ".decoder_layers.0.0.1.bias", C, &p->up_b);
    add_slot(c, pre + ".decoder_layers.0.1.weight", 2 * C * C, &p->fuse_w);
    add_slot(c, pre + ".decoder_layers.0.1.bias", C, &p->fuse_b);
    for (int j = 0; j < 2; ++j) add_block(c, pre + ".decoder_layers.0.2.blocks." + std::to_string(j), &p->dec[j], C);
    add_slot(c, pre + ".tail.1.weight", B * C, &p->tail_w);
    add_slot(c, pre + ".tail.1.bias", B, &p->tail_b);
  }
  size_t off = 0;
  for (auto& s : c->slots) { s.offset = off; off += align4((size_t)s.numel); }
  c->flat_floats = off;
  for (auto& d : c->derived) {
    d.offset = off;
    size_t floats = d.rows > 0 ? (size_t)d.rows * d.cols : d.rows == 0 ? 2 * 64 * 64
                    : d.rows == -1 ? ffn_tc_pack_halves(d.cols) / 2
                    : d.rows == -3 ? ffn_cl_pack_halves(d.cols) / 2 : (size_t)d.cols * d.cols;
    off += align4(floats);
  }
  c->arena_floats = off;
}

bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int check_shape(const lgteun_ctx* c, int N, int h, int w) {
  if (N <= 0) return fail(LGTEUN_EINVAL, "batch must be positive");
  // what the reference accepts: PAN height / width multiples of 16 (8x8 windows at full and at half resolution, LGT.py:135);
  // FFT lengths with odd factors take the generic radix stages of fft_mixer.cu
  if (h < 4 || w < 4 || (h & 3) || (w & 3) || 4 * h > 1024 || 4 * w > 1024)
    return fail(LGTEUN_EINVAL, "unsupported shape: PAN height/width (4h, 4w) must be multiples of 16 in [16, 1024] "
                               "(8x8 windows at two U-Net levels); got h=" +
                                   std::to_string(h) + " w=" + std::to_string(w));
  (void)c;
  return 0;
}

size_t ws_layout(const lgteun_ctx* c, int N, int h, int w, Workspace* out) {
  const size_t B = c->B, C = c->C, H = 4 * (size_t)h, W = 4 * (size_t)w, P = H * W;
  size_t off = 0;
  auto take = [&](size_t floats) { size_t o = off; off += (floats + 63) & ~(size_t)63; return o; };
  size_t o_ms = take(N * B * h * w), o_pan = take(N * P), o_out = take(N * B * P);
  size_t o_zA = take(N * B * P), o_zB = take(N * B * P), o_res = take(data_step_scratch_floats(N, (int)B, h, w));
  size_t o_X0 = take(N * P * C), o_X1 = take(N * P * C), o_X2 = take(N * P * C);
  size_t o_L0 = take(N * P / 4 * 2 * C), o_L1 = take(N * P / 4 * 2 * C);
  size_t o_loc = take(N * P * C / 2);
  size_t spec_full = spectrum_floats(N, (int)H, (int)W, (int)C / 2), spec_low = spectrum_floats(N, (int)H / 2, (int)W / 2, (int)C);
  size_t o_spec = take(spec_full > spec_low ? spec_full : spec_low);
  size_t o_hid = take(ffn_hidden_floats(N, (int)H, (int)W, (int)C));   // == bottleneck size (P/4 * 8C = 2PC) < 4PC
  if (out) {
    float* b = c->ws_base;
    *out = Workspace{b + o_ms, b + o_pan, b + o_out, b + o_zA, b + o_zB, b + o_res, b + o_X0, b + o_X1, b + o_X2,
                     b + o_L0, b + o_L1, b + o_loc, b + o_spec, b + o_hid};
  }
  return off * sizeof(float);
}

void drop_graphs(lgteun_ctx* c) {
  for (auto& g : c->graphs) { cudaGraphExecDestroy(g.exec); cudaGraphDestroy(g.graph); }
  c->graphs.clear();
}

int ensure_ws(lgteun_ctx* c, int N, int h, int w, Workspace* ws) {
  size_t need = ws_layout(c, N, h, w, nullptr);
  if (need > c->ws_bytes) {
    CK(cudaDeviceSynchronize());
    drop_graphs(c);
    if (c->ws_base) cudaFree(c->ws_base);
    c->ws_base = nullptr;
    c->ws_bytes = 0;
    cudaError_t e = cudaMalloc(&c->ws_base, need);
    if (e != cudaSuccess) {
      g_err = "workspace cudaMalloc of " + std::to_string(need) + " bytes failed: " + cudaGetErrorString(e);
      return LGTEUN_ENOMEM;
    }
    c->ws_bytes = need;
  }
  ws_layout(c, N, h, w, ws);
  return 0;
}

struct Launcher {       // counts launches and stops at the first error; optional per-launch CUDA-event timing
  cudaError_t err = cudaSuccess;
  int count = 0;
  cudaStream_t stream = nullptr;
  bool timing = false;    // LGTEUN_TIMING=1 with LGTEUN_NO_GRAPH: warm-cache device time per kernel family (stderr)
  struct Rec { std::string tag; cudaEvent_t e0, e1; };
  std::vector<Rec> recs;
  cudaEvent_t pending = nullptr;
  std::string pending_tag;
  void begin(const char* call, int ch = 0) {
    if (!timing) return;
    std::string t(call);
    t = t.substr(0, t.find('('));
    if (ch) t += " c=" + std::to_string(ch);
    pending_tag = t;
    cudaEventCreate(&pending);
    cudaEventRecord(pending, stream);
  }
  void operator()(cudaError_t e, int kernels = 1) {
    if (err == cudaSuccess) { err = e; count += kernels; }
    if (timing && pending) {
      cudaEvent_t e1;
      cudaEventCreate(&e1);
      cudaEventRecord(e1, stream);
      recs.push_back({pending_tag, pending, e1});
      pending = nullptr;
    }
  }
  bool ok() const { return err == cudaSuccess; }
  void report() {
    if (!timing || recs.empty()) return;
    cudaStreamSynchronize(stream);
    std::map<std::string, std::pair<int, float>> agg;
    float total = 0.f;
    for (auto& r : recs) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.e0, r.e1);
      agg[r.tag].first++;
      agg[r.tag].second += ms;
      total += ms;
      cudaEventDestroy(r.e0);
      cudaEventDestroy(r.e1);
    }
    fprintf(stderr, "[lgteun timing] %-34s %5s %10s %7s\n", "launch", "n", "total_us", "share");
    for (auto& kv : agg)
      fprintf(stderr, "[lgteun timing] %-34s %5d %10.1f %6.1f%%\n", kv.first.c_str(), kv.second.first, kv.second.second * 1e3f,
              100.f * kv.second.second / total);
    fprintf(stderr, "[lgteun timing] %-34s %5s %10.1f\n", "total", "", total * 1e3f);
    recs.clear();
  }
};
#define LG_L(L, ch, call, ...) do { (L).begin(#call, (ch)); (L)((call), ##__VA_ARGS__); } while (0)

// The conv-FFN runs as the fused tcgen05/TMEM kernel for c in {16, 32}; c = 64 (the WV-3 bottleneck) does not fit
// its weights / hidden rows on chip and runs as three tcgen05 pixel-GEMMs + one depthwise kernel (pwgemm_tc.cu).
// LGTEUN_FFN=simt selects the fp32 CUDA-core kernels of ffn.cu for every block (A/B measurement only).
bool ffn_simt() {
  static const bool simt = [] { const char* e = getenv("LGTEUN_FFN"); return e && std::string(e) == "simt"; }();
  return simt;
}
// LGTEUN_FFN=tc selects the pixels-on-lanes kernel of ffn_tc.cu instead of ffn_cl.cu (A/B measurement)
bool ffn_cl() {
  static const bool cl = [] { const char* e = getenv("LGTEUN_FFN"); return !(e && std::string(e) == "tc"); }();
  return cl;
}
bool use_tc_ffn(int ch) { return !ffn_simt() && (ch == 16 || ch == 32); }
bool use_wide_tc_ffn(int ch) { return !ffn_simt() && ch == 64; }
int ffn_launches(int ch) { return use_tc_ffn(ch) ? 1 : use_wide_tc_ffn(ch) ? 4 : 2; }
cudaError_t run_ffn(const BlockW& b, int ch, const float* x, float* y, float* hidden, int N, int H, int W, cudaStream_t s) {
  if (use_tc_ffn(ch)) return ffn_cl() ? launch_ffn_cl(b, ch, x, y, N, H, W, s) : launch_ffn_tc(b, ch, x, y, N, H, W, s);
  if (use_wide_tc_ffn(ch)) return launch_ffn_wide_tc(b, x, hidden, hidden + (size_t)N * H * W * 256, y, N, H, W, s);
  return launch_ffn(b, ch, x, hidden, y, N, H, W, s);
}

// x + LGMixer(LN(x)): window MSA on the first channel half || FFT mixer on the second, proj, residual.
void run_mixer(Launcher& L, const BlockW& b, int ch, const float* x, float* y, const Workspace& ws, int N, int H, int W,
               cudaStream_t s) {
  LG_L(L, ch, launch_window_msa(b, ch, x, ws.loc, 1, N, H, W, s));
  LG_L(L, ch, launch_fft_rows_fwd(b, ch, x, ws.spec, 1, N, H, W, s));
  LG_L(L, ch, launch_fft_cols(b, ch, ws.spec, N, H, W, s));
  LG_L(L, ch, launch_fft_rows_inv(b, ch, ws.spec, ws.loc, x, y, 1, N, H, W, s));
}
// one LGB block: a -> (mixer) -> t -> (ffn) -> a
void run_block(Launcher& L, const BlockW& b, int ch, float* a, float* t, const Workspace& ws, int N, int H, int W,
               cudaStream_t s) {
  run_mixer(L, b, ch, a, t, ws, N, H, W, s);
  LG_L(L, ch, run_ffn(b, ch, t, a, ws.hidden, N, H, W, s), ffn_launches(ch));
}
// LGT.forward (LGT.py:314-344)
void run_prior(Launcher& L, const lgteun_ctx* c, int i, const float* zin, float* zout, const Workspace& ws, int N, int H,
               int W, cudaStream_t s) {
  const PriorW& p = c->wv.prior[i];
  const int C = c->C;
  LG_L(L, 0, launch_patch_embed(p, c->B, zin, ws.X0, N, H, W, s));
  for (int j = 0; j < 2; ++j) run_block(L, p.enc[j], C, ws.X0, ws.X1, ws, N, H, W, s);       // skip = X0
  LG_L(L, 0, launch_down(p, C, ws.X0, ws.L0, N, H, W, s));
  run_block(L, p.bott[0], 2 * C, ws.L0, ws.L1, ws, N, H / 2, W / 2, s);
  LG_L(L, 0, launch_up_fuse(p, C, ws.L0, ws.X0, ws.L1, ws.X1, N, H, W, s), 2);
  for (int j = 0; j < 2; ++j) run_block(L, p.dec[j], C, ws.X1, ws.X2, ws, N, H, W, s);
  LG_L(L, 0, launch_tail(p, c->B, ws.X1, zin, zout, N, H, W, s));
}
// Pansharpening.forward (unlg_former.py:50-67)
cudaError_t run_forward(const lgteun_ctx* c, const float* ms, const float* pan, float* out, const Workspace& ws, int N,
                        int h, int w, int flags, cudaStream_t s, int* launches, bool timing = false) {
  Launcher L;
  L.stream = s;
  L.timing = timing;
  const int H = 4 * h, W = 4 * w;
  LG_L(L, 0, launch_bicubic(ms, ws.zA, N * c->B, h, w, 4, 1, s));
  float *za = ws.zA, *zb = ws.zB;
  for (int i = 0; i < c->K; ++i) {
    LG_L(L, 0, launch_data_step(c->wv.dw, i, c->B, za, ms, pan, ws.resid, zb, N, h, w, s), data_step_launches());
    float* t = za; za = zb; zb = t;
    const bool last = (i == c->K - 1);
    // the reference discards the priors of stages 0..K-2 (unlg_former.py:63-67); run them only on request
    if (last) run_prior(L, c, i, za, out, ws, N, H, W, s);
    else if (flags & LGTEUN_RUN_DEAD_PRIORS) run_prior(L, c, i, za, zb, ws, N, H, W, s);   // zb is scratch here
  }
  if (launches) *launches = L.count;
  L.report();
  return L.err;
}

}  // namespace

extern "C" {

int lgteun_abi_version(void) { return 1; }
const char* lgteun_last_error(void) { return g_err.c_str(); }

int lgteun_create(int device, int bands, int stages, lgteun_t** out) {
  if (!out) return fail(LGTEUN_EINVAL, "out is NULL");
  *out = nullptr;
  if (bands != 4 && bands != 8) return fail(LGTEUN_EINVAL, "bands (cfg.ms_chans) must be 4 or 8");
  if (stages < 1 || stages > kMaxStages) return fail(LGTEUN_EINVAL, "stages must be in [1, 8]");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(LGTEUN_EINVAL, "no such CUDA device");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(LGTEUN_EINVAL, std::string("this library is built for sm_100a (B200) only; device is sm_") +
                                   std::to_string(prop.major) + std::to_string(prop.minor));
  lgteun_ctx* c = new lgteun_ctx();
  c->device = device; c->B = bands; c->C = 4 * bands; c->K = stages;
  memset(&c->wv.dw, 0, sizeof(c->wv.dw));
  memset(c->wv.prior, 0, sizeof(c->wv.prior));
  build_table(c);
  cudaError_t e = cudaMalloc(&c->arena, c->arena_floats * sizeof(float));
  if (e != cudaSuccess) { delete c; return fail_cuda(e, "cudaMalloc(weight arena)"); }
  for (auto& s : c->slots) *s.slot = c->arena + s.offset;
  for (auto& d : c->derived) *d.dst = c->arena + d.offset;
  e = fft_init_tables(0);
  if (e != cudaSuccess) { cudaFree(c->arena); delete c; return fail_cuda(e, "fft_init_tables"); }
  *out = c;
  return 0;
}

void lgteun_destroy(lgteun_t* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  drop_graphs(c);
  lgctx::train_destroy(c);
  if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
  if (c->metric_acc) cudaFree(c->metric_acc);
  if (c->ws_base) cudaFree(c->ws_base);
  if (c->arena) cudaFree(c->arena);
  delete c;
}

int lgteun_num_weights(const lgteun_t* c) { return c ? (int)c->slots.size() : 0; }
const char* lgteun_weight_name(const lgteun_t* c, int i) {
  return (c && i >= 0 && i < (int)c->slots.size()) ? c->slots[i].name.c_str() : nullptr;
}
int64_t lgteun_weight_numel(const lgteun_t* c, int i) {
  return (c && i >= 0 && i < (int)c->slots.size()) ? c->slots[i].numel : -1;
}

int lgteun_load_weights(lgteun_t* c, const char* const* names, const float* const* ptrs, const int64_t* numels, int n,
                        void* stream) {
  if (!c || !names || !ptrs || !numels) return fail(LGTEUN_EINVAL, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaSetDevice(c->device));
  std::map<std::string, int> idx;
  for (int i = 0; i < n; ++i) idx[names[i]] = i;
  for (auto& sl : c->slots) {
    auto it = idx.find(sl.name);
    if (it == idx.end()) return fail(LGTEUN_ESTATE, "state_dict is missing key '" + sl.name + "'");
    if (numels[it->second] != sl.numel)
      return fail(LGTEUN_ESTATE, "size mismatch for '" + sl.name + "': expected " + std::to_string(sl.numel) +
                                     " elements, got " + std::to_string(numels[it->second]));
    if (!ptrs[it->second]) return fail(LGTEUN_EINVAL, "NULL pointer for '" + sl.name + "'");
  }
  for (auto& sl : c->slots)
    CK(cudaMemcpyAsync(c->arena + sl.offset, ptrs[idx[sl.name]], sl.numel * sizeof(float), cudaMemcpyDeviceToDevice, s));
  for (auto& d : c->derived) {
    float* dst = c->arena + d.offset;
    if (d.rows == 0) CK(launch_transpose_pos(*d.src, dst, s));
    else if (d.rows > 0) CK(launch_transpose(*d.src, dst, d.rows, d.cols, s));
    else if (d.rows == -2) {
      char* base = reinterpret_cast<char*>(dst);
      CK(launch_pack_umma_f16(*d.src, base, base + (size_t)d.cols * d.cols * 2, d.cols, d.cols, s));
    } else if (d.rows == -3) {
      CK(launch_ffn_cl_pack(*d.blk, d.cols, dst, s));
    } else {
      // fp16 hi/lo operands of the three FFN GEMMs: w0h | w0l | w1h | w1l | w2h | w2l (halves), see ffn_tc.cu
      const int ch = d.cols, c4 = 4 * ch;
      char* base = reinterpret_cast<char*>(dst);
      if (ch == 64) {              // wide path: plain matrices
        const size_t s0 = (size_t)c4 * ch * 2, s1 = (size_t)c4 * c4 * 2, s2 = (size_t)ch * c4 * 2;   // bytes per half-tensor
        CK(launch_pack_umma_f16(d.blk->f0_w, base, base + s0, c4, ch, s));
        CK(launch_pack_umma_f16(d.blk->f1_w, base + 2 * s0, base + 2 * s0 + s1, c4, c4, s));
        CK(launch_pack_umma_f16(d.blk->f2_w, base + 2 * s0 + 2 * s1, base + 2 * s0 + 2 * s1 + s2, ch, c4, s));
      } else {                     // fused kernel: GEMM1 / GEMM2 carry their bias as an extra K-step
        const size_t s0 = (size_t)c4 * (ch + 8) * 2, s1 = (size_t)c4 * (c4 + 8) * 2, s2 = (size_t)ch * c4 * 2;
        CK(launch_pack_umma_f16_bias(d.blk->f0_w, d.blk->f0_b, base, base + s0, c4, ch, s));
        CK(launch_pack_umma_f16_bias(d.blk->f1_w, d.blk->f1_b, base + 2 * s0, base + 2 * s0 + s1, c4, c4, s));
        CK(launch_pack_umma_f16(d.blk->f2_w, base + 2 * s0 + 2 * s1, base + 2 * s0 + 2 * s1 + s2, ch, c4, s));
      }
    }
  }
  c->loaded = true;
  return 0;
}

int64_t lgteun_workspace_bytes(const lgteun_t* c, int N, int h, int w) {
  if (!c || check_shape(c, N, h, w)) return -1;
  return (int64_t)ws_layout(c, N, h, w, nullptr);
}

int lgteun_forward_launches(lgteun_t* c, int N, int h, int w, int flags) {
  if (!c || check_shape(c, N, h, w)) return -1;
  // patch_embed, 5 blocks x (msa, 3 fft passes, ffn = 1 fused tcgen05 launch or 2 CUDA-core launches), down, up_fuse, tail
  const int per_prior = 1 + 4 * (4 + ffn_launches(c->C)) + (4 + ffn_launches(2 * c->C)) + 4;
  const int priors = (flags & LGTEUN_RUN_DEAD_PRIORS) ? c->K : 1;
  return 1 + data_step_launches() * c->K + priors * per_prior;
}

static int find_copy_nodes(cudaGraph_t g, const Workspace& ws, GraphEntry* ge) {
  size_t n = 0;
  CK(cudaGraphGetNodes(g, nullptr, &n));
  std::vector<cudaGraphNode_t> nodes(n);
  CK(cudaGraphGetNodes(g, nodes.data(), &n));
  ge->n_ms = ge->n_pan = ge->n_out = nullptr;
  for (auto nd : nodes) {
    cudaGraphNodeType t;
    CK(cudaGraphNodeGetType(nd, &t));
    if (t != cudaGraphNodeTypeMemcpy) continue;
    cudaMemcpy3DParms p;
    CK(cudaGraphMemcpyNodeGetParams(nd, &p));
    if (p.dstPtr.ptr == (void*)ws.ms) ge->n_ms = nd;
    else if (p.dstPtr.ptr == (void*)ws.pan) ge->n_pan = nd;
    else if (p.srcPtr.ptr == (void*)ws.out) ge->n_out = nd;
  }
  if (!ge->n_ms || !ge->n_pan || !ge->n_out) return fail(LGTEUN_ECUDA, "could not locate the I/O copy nodes of the graph");
  return 0;
}

int lgteun_forward(lgteun_t* c, const float* ms, const float* pan, float* out, int N, int h, int w, int flags,
                   void* stream) {
  if (!c || !ms || !pan || !out) return fail(LGTEUN_EINVAL, "NULL argument");
  if (!c->loaded) return fail(LGTEUN_ESTATE, "weights not loaded (call lgteun_load_weights first)");
  int rc = check_shape(c, N, h, w);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaSetDevice(c->device));
  Workspace ws;
  rc = ensure_ws(c, N, h, w, &ws);
  if (rc) return rc;
  const size_t B = c->B, P = 16 * (size_t)h * w;
  const size_t ms_bytes = N * B * h * w * sizeof(float), pan_bytes = N * P * sizeof(float), out_bytes = N * B * P * sizeof(float);

  if (flags & LGTEUN_NO_GRAPH) {
    static const bool timing = [] { const char* e = getenv("LGTEUN_TIMING"); return e && e[0] == '1'; }();
    cudaError_t e = run_forward(c, ms, pan, out, ws, N, h, w, flags, s, &c->last_launches, timing);
    if (e != cudaSuccess) return fail_cuda(e, "forward launch");
    return 0;
  }
  GraphEntry* ge = nullptr;
  for (size_t i = 0; i < c->graphs.size(); ++i) {
    GraphEntry& g = c->graphs[i];
    if (!(g.N == N && g.h == h && g.w == w && g.flags == flags)) continue;
    // replay with other caller buffers: retarget the three I/O copy nodes of the instantiated graph
    bool ok = true;
    if (g.ms != ms)
      ok = ok && cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.n_ms, ws.ms, ms, ms_bytes, cudaMemcpyDeviceToDevice) == cudaSuccess;
    if (ok && g.pan != pan)
      ok = ok && cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.n_pan, ws.pan, pan, pan_bytes, cudaMemcpyDeviceToDevice) == cudaSuccess;
    if (ok && g.out != out)
      ok = ok && cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.n_out, out, ws.out, out_bytes, cudaMemcpyDeviceToDevice) == cudaSuccess;
    if (ok) {
      g.ms = ms; g.pan = pan; g.out = out;
      ge = &g;
    } else {                       // could not retarget (driver refused): drop the entry and capture afresh
      cudaGetLastError();
      cudaGraphExecDestroy(g.exec);
      cudaGraphDestroy(g.graph);
      c->graphs.erase(c->graphs.begin() + i);
    }
    break;
  }
  if (!ge) {
    // Stage chaining: the whole K-stage forward is captured once per (N, h, w) into one CUDA graph.
    // I/O goes through fixed staging buffers so that replays with different caller pointers only
    // retarget three copy nodes.
    cudaGraph_t graph = nullptr;
    if (!c->cap_stream) CK(cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking));
    cudaStream_t cs = c->cap_stream;
    CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
    cudaError_t e = cudaMemcpyAsync(ws.ms, ms, ms_bytes, cudaMemcpyDeviceToDevice, cs);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ws.pan, pan, pan_bytes, cudaMemcpyDeviceToDevice, cs);
    int launches = 0;
    if (e == cudaSuccess) e = run_forward(c, ws.ms, ws.pan, ws.out, ws, N, h, w, flags, cs, &launches);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, ws.out, out_bytes, cudaMemcpyDeviceToDevice, cs);
    cudaError_t e2 = cudaStreamEndCapture(cs, &graph);
    if (e != cudaSuccess || e2 != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return fail_cuda(e != cudaSuccess ? e : e2, "graph capture of the forward");
    }
    GraphEntry g{N, h, w, flags, graph, nullptr, nullptr, nullptr, nullptr, ms, pan, out, launches};
    rc = find_copy_nodes(graph, ws, &g);
    if (rc) { cudaGraphDestroy(graph); return rc; }
    e = cudaGraphInstantiate(&g.exec, graph, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); return fail_cuda(e, "cudaGraphInstantiate"); }
    if (c->graphs.size() >= 8) {
      cudaGraphExecDestroy(c->graphs.front().exec);
      cudaGraphDestroy(c->graphs.front().graph);
      c->graphs.erase(c->graphs.begin());
    }
    c->graphs.push_back(g);
    ge = &c->graphs.back();
  }
  c->last_launches = ge->launches;
  CK(cudaGraphLaunch(ge->exec, s));
  return 0;
}

int lgteun_forward_host(lgteun_t* c, const float* ms_host, const float* pan_host, float* out_host, int N, int h, int w,
                        int flags, void* stream) {
  if (!c || !ms_host || !pan_host || !out_host) return fail(LGTEUN_EINVAL, "NULL argument");
  if (!c->loaded) return fail(LGTEUN_ESTATE, "weights not loaded (call lgteun_load_weights first)");
  int rc = check_shape(c, N, h, w);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaSetDevice(c->device));
  Workspace ws;
  rc = ensure_ws(c, N, h, w, &ws);
  if (rc) return rc;
  const size_t B = c->B, P = 16 * (size_t)h * w;
  // the staging buffers double as the device side of the host transfer; the forward then runs un-staged
  CK(cudaMemcpyAsync(ws.ms, ms_host, N * B * h * w * sizeof(float), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ws.pan, pan_host, N * P * sizeof(float), cudaMemcpyHostToDevice, s));
  cudaError_t e = run_forward(c, ws.ms, ws.pan, ws.out, ws, N, h, w, flags, s, &c->last_launches);
  if (e != cudaSuccess) return fail_cuda(e, "forward launch");
  CK(cudaMemcpyAsync(out_host, ws.out, N * B * P * sizeof(float), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

// ---- per-operator entry points ----------------------------------------------------------------------------
static const BlockW* pick_block(const lgteun_ctx* c, int prior, int lgb, int block, int* ch) {
  if (prior < 0 || prior >= c->K) return nullptr;
  const PriorW& p = c->wv.prior[prior];
  if (lgb == 0 && block >= 0 && block < 2) { *ch = c->C; return &p.enc[block]; }
  if (lgb == 1 && block == 0) { *ch = 2 * c->C; return &p.bott[0]; }
  if (lgb == 2 && block >= 0 && block < 2) { *ch = c->C; return &p.dec[block]; }
  return nullptr;
}
#define OP_PROLOGUE                                                                 \
  if (!c) return fail(LGTEUN_EINVAL, "NULL handle");                                \
  if (!c->loaded) return fail(LGTEUN_ESTATE, "weights not loaded");                 \
  cudaStream_t s = (cudaStream_t)stream;                                            \
  CK(cudaSetDevice(c->device));

static int op_shape(lgteun_ctx* c, int N, int H, int W, int min_side, Workspace* ws) {
  // powers of two (down to min_side) or multiples of 8
  if (N <= 0 || H < min_side || W < min_side || H > 1024 || W > 1024 || !(pow2(H) || (H & 7) == 0) || !(pow2(W) || (W & 7) == 0))
    return fail(LGTEUN_EINVAL, "unsupported operator shape");
  // a workspace for an (N, H/4, W/4) forward covers every operator at map size H x W (and smaller)
  int h = (H + 3) / 4 > 4 ? (H + 3) / 4 : 4, w = (W + 3) / 4 > 4 ? (W + 3) / 4 : 4;
  return ensure_ws(c, N, h, w, ws);
}

int lgteun_op_bicubic(lgteun_t* c, const float* x, float* y, int planes, int h, int w, int num, int den, void* stream) {
  if (!c) return fail(LGTEUN_EINVAL, "NULL handle");
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaSetDevice(c->device));
  if (!((num == 4 && den == 1) || (num == 2 && den == 1) || (num == 1 && den == 2) || (num == 1 && den == 1)))
    return fail(LGTEUN_EINVAL, "scale must be 4, 2, 1 or 1/2");
  if (planes <= 0 || h < 2 || w < 2 || (den == 2 && ((h | w) & 1))) return fail(LGTEUN_EINVAL, "bad plane shape");
  CK(launch_bicubic(x, y, planes, h, w, num, den, s));
  return 0;
}

int lgteun_op_data_step(lgteun_t* c, int stage, const float* z_in, const float* ms, const float* pan, float* z_out,
                        int N, int h, int w, void* stream) {
  OP_PROLOGUE
  if (stage < 0 || stage >= c->K) return fail(LGTEUN_EINVAL, "no such stage");
  if (z_in == z_out) return fail(LGTEUN_EINVAL, "z_in and z_out must not alias");
  int rc = check_shape(c, N, h, w);
  if (rc) return rc;
  Workspace ws;
  rc = ensure_ws(c, N, h, w, &ws);
  if (rc) return rc;
  CK(launch_data_step(c->wv.dw, stage, c->B, z_in, ms, pan, ws.resid, z_out, N, h, w, s));
  return 0;
}

int lgteun_op_patch_embed(lgteun_t* c, int prior, const float* x, float* y, int N, int H, int W, void* stream) {
  OP_PROLOGUE
  if (prior < 0 || prior >= c->K || N <= 0 || H <= 0 || W <= 0) return fail(LGTEUN_EINVAL, "bad argument");
  CK(launch_patch_embed(c->wv.prior[prior], c->B, x, y, N, H, W, s));
  return 0;
}

int lgteun_op_mixer(lgteun_t* c, int prior, int lgb, int block, const float* x, float* y, int N, int H, int W,
                    void* stream) {
  OP_PROLOGUE
  int ch = 0;
  const BlockW* b = pick_block(c, prior, lgb, block, &ch);
  if (!b) return fail(LGTEUN_EINVAL, "no such block");
  if (x == y) return fail(LGTEUN_EINVAL, "x and y must not alias");
  Workspace ws;
  int rc = op_shape(c, N * (lgb == 1 ? 4 : 1), H, W, 8, &ws);
  if (rc) return rc;
  Launcher L;
  run_mixer(L, *b, ch, x, y, ws, N, H, W, s);
  if (!L.ok()) return fail_cuda(L.err, "mixer launch");
  return 0;
}

int lgteun_op_local_mixer(lgteun_t* c, int prior, int lgb, int block, const float* x, float* y, int N, int H, int W,
                          void* stream) {
  OP_PROLOGUE
  int ch = 0;
  const BlockW* b = pick_block(c, prior, lgb, block, &ch);
  if (!b) return fail(LGTEUN_EINVAL, "no such block");
  if (N <= 0 || H < 8 || W < 8 || (H & 7) || (W & 7)) return fail(LGTEUN_EINVAL, "H and W must be multiples of 8");
  CK(launch_window_msa(*b, ch, x, y, 0, N, H, W, s));
  return 0;
}

int lgteun_op_global_mixer(lgteun_t* c, int prior, int lgb, int block, const float* x, float* y, int N, int H, int W,
                           void* stream) {
  OP_PROLOGUE
  int ch = 0;
  const BlockW* b = pick_block(c, prior, lgb, block, &ch);
  if (!b) return fail(LGTEUN_EINVAL, "no such block");
  Workspace ws;
  int rc = op_shape(c, N * (lgb == 1 ? 4 : 1), H, W, 8, &ws);
  if (rc) return rc;
  Launcher L;
  L(launch_fft_rows_fwd(*b, ch, x, ws.spec, 0, N, H, W, s));
  L(launch_fft_cols(*b, ch, ws.spec, N, H, W, s));
  L(launch_fft_rows_inv(*b, ch, ws.spec, nullptr, nullptr, y, 0, N, H, W, s));
  if (!L.ok()) return fail_cuda(L.err, "global mixer launch");
  return 0;
}

int lgteun_op_ffn(lgteun_t* c, int prior, int lgb, int block, const float* x, float* y, int N, int H, int W,
                  void* stream) {
  OP_PROLOGUE
  int ch = 0;
  const BlockW* b = pick_block(c, prior, lgb, block, &ch);
  if (!b) return fail(LGTEUN_EINVAL, "no such block");
  if (x == y) return fail(LGTEUN_EINVAL, "x and y must not alias");
  Workspace ws;
  int rc;
  if (use_tc_ffn(ch) && ffn_cl()) {
    // the channels-on-lanes kernel takes any map size (row segments and bands are clipped), and needs no scratch
    if (N <= 0 || H < 1 || W < 1 || H > 1024 || W > 1024) return fail(LGTEUN_EINVAL, "unsupported operator shape");
    rc = ensure_ws(c, 1, 4, 4, &ws);
  } else {
    rc = op_shape(c, N * (lgb == 1 ? 4 : 1), H, W, 1, &ws);
  }
  if (rc) return rc;
  CK(run_ffn(*b, ch, x, y, ws.hidden, N, H, W, s));
  return 0;
}

int lgteun_op_prior(lgteun_t* c, int prior, const float* x, float* y, int N, int H, int W, void* stream) {
  OP_PROLOGUE
  if (prior < 0 || prior >= c->K) return fail(LGTEUN_EINVAL, "no such prior");
  int rc = check_shape(c, N, H / 4, W / 4);
  if (rc) return rc;
  if ((H & 3) || (W & 3)) return fail(LGTEUN_EINVAL, "H and W must be multiples of 16");
  Workspace ws;
  rc = ensure_ws(c, N, H / 4, W / 4, &ws);
  if (rc) return rc;
  Launcher L;
  run_prior(L, c, prior, x, y, ws, N, H, W, s);
  if (!L.ok()) return fail_cuda(L.err, "prior launch");
  return 0;
}

int lgteun_op_metrics(lgteun_t* c, const float* pred, const float* gt, double* out_dev, int N, int H, int W, float max_value,
                      void* stream) {
  if (!c || !pred || !gt || !out_dev) return fail(LGTEUN_EINVAL, "NULL argument");
  if (N <= 0 || H <= 0 || W <= 0) return fail(LGTEUN_EINVAL, "bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaSetDevice(c->device));
  const size_t need = (size_t)N * (2 + 2 * c->B);
  if (need > c->metric_acc_doubles) {
    CK(cudaStreamSynchronize(s));
    if (c->metric_acc) cudaFree(c->metric_acc);
    c->metric_acc = nullptr;
    c->metric_acc_doubles = 0;
    CK(cudaMalloc(&c->metric_acc, need * sizeof(double)));
    c->metric_acc_doubles = need;
  }
  CK(launch_metrics(pred, gt, c->metric_acc, out_dev, N, c->B, H, W, max_value, s));
  return 0;
}

int lgteun_op_normalize(lgteun_t* c, const float* raw, float* out, int64_t n, float max_value, void* stream) {
  if (!c || !raw || !out) return fail(LGTEUN_EINVAL, "NULL argument");
  if (n <= 0 || !(max_value > 0.f)) return fail(LGTEUN_EINVAL, "bad size or max_value");
  CK(cudaSetDevice(c->device));
  CK(launch_normalize(raw, out, (size_t)n, max_value, (cudaStream_t)stream));
  return 0;
}

int lgteun_op_to_nhwc(lgteun_t* c, const float* nchw, float* nhwc, int N, int C, int H, int W, float scale, void* stream) {
  if (!c || !nchw || !nhwc) return fail(LGTEUN_EINVAL, "NULL argument");
  if (N <= 0 || H <= 0 || W <= 0 || !(C == 1 || C == 4 || C == 8)) return fail(LGTEUN_EINVAL, "bad shape (C must be 1, 4 or 8)");
  CK(cudaSetDevice(c->device));
  CK(launch_to_nhwc(nchw, nhwc, N, C, H, W, scale, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
