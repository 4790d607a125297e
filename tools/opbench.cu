// opbench — native (no Python, no torch) timing / A-B harness over the C ABI of include/lgteun.h.
//
//   tools/opbench [--bands 4|8] [--batch N] [--pan P] [--iters K] [--ops a,b,...] [--save DIR] [--ref DIR]
//
// Creates a handle with seeded random weights, runs each requested operator through its lgteun_op_* entry point (the
// same kernels the forward chains) on seeded random inputs, times it with CUDA events on the launch stream (3 warm-up
// launches, then K timed ones; inputs + outputs of every timed case are larger than L2 unless --batch is tiny) and
// prints one line per operator: microseconds per launch, algorithmic GB/s, and — with --ref DIR — the max |delta|
// against the outputs a previous run stored with --save DIR (A/B runs of two kernel variants selected by environment
// switches).  Starts in about a second on a fresh box, so one gpurun call can sweep many variants.
// Build: see tools/build_opbench.sh (nvcc, links lgteun_b200/_lgteun_cuda.so with an $ORIGIN rpath).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../include/lgteun.h"

#define CKC(x)                                                                                      \
  do {                                                                                              \
    cudaError_t e_ = (x);                                                                           \
    if (e_ != cudaSuccess) {                                                                        \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                                      \
    }                                                                                               \
  } while (0)
#define CKL(x)                                                                                      \
  do {                                                                                              \
    int rc_ = (x);                                                                                  \
    if (rc_ != 0) {                                                                                 \
      fprintf(stderr, "lgteun error %d (%s) at %s:%d\n", rc_, lgteun_last_error(), __FILE__, __LINE__); \
      exit(3);                                                                                      \
    }                                                                                               \
  } while (0)

static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static float urand() {   // xorshift64*, uniform in [0, 1)
  g_rng ^= g_rng >> 12;
  g_rng ^= g_rng << 25;
  g_rng ^= g_rng >> 27;
  return (float)((g_rng * 0x2545F4914F6CDD1Dull) >> 40) / 16777216.0f;
}
static float nrand() {   // approx normal
  float s = 0.f;
  for (int i = 0; i < 12; ++i) s += urand();
  return s - 6.0f;
}

static float* dev_random(size_t n, float lo, float hi, uint64_t seed) {
  std::vector<float> h(n);
  g_rng = seed * 0x9E3779B97F4A7C15ull + 12345;
  for (size_t i = 0; i < n; ++i) h[i] = lo + (hi - lo) * urand();
  float* d = nullptr;
  CKC(cudaMalloc(&d, n * sizeof(float)));
  CKC(cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}

static bool has(const std::string& s, const char* sub) { return s.find(sub) != std::string::npos; }

static void load_random_weights(lgteun_t* ctx) {
  const int n = lgteun_num_weights(ctx);
  std::vector<const char*> names(n);
  std::vector<const float*> ptrs(n);
  std::vector<int64_t> numels(n);
  std::vector<float*> owned;
  g_rng = 777;
  for (int i = 0; i < n; ++i) {
    const std::string name = lgteun_weight_name(ctx, i);
    const int64_t ne = lgteun_weight_numel(ctx, i);
    std::vector<float> h((size_t)ne);
    for (int64_t j = 0; j < ne; ++j) {
      float v;
      if (has(name, "norm.weight")) v = 1.0f + 0.1f * (urand() - 0.5f);
      else if (has(name, "norm.bias")) v = 0.1f * (urand() - 0.5f);
      else if (has(name, "pos_emb")) v = fmaxf(-2.f, fminf(2.f, nrand()));
      else if (has(name, "eta")) v = 0.1f;
      else if (has(name, "bias")) v = 0.2f * (urand() - 0.5f);
      else if (has(name, "depth_conv.weight") || has(name, "D.") || has(name, "DT.")) v = 0.6f * (urand() - 0.5f);
      else if (has(name, "conv_amp") || has(name, "conv_pha") || has(name, "proj.0.weight")) v = 1.0f + 0.5f * (urand() - 0.5f);
      else v = 0.5f * (urand() - 0.5f);
      h[(size_t)j] = v;
    }
    float* d = nullptr;
    CKC(cudaMalloc(&d, (size_t)ne * sizeof(float)));
    CKC(cudaMemcpy(d, h.data(), (size_t)ne * sizeof(float), cudaMemcpyHostToDevice));
    owned.push_back(d);
    names[i] = lgteun_weight_name(ctx, i);
    ptrs[i] = d;
    numels[i] = ne;
  }
  CKL(lgteun_load_weights(ctx, names.data(), ptrs.data(), numels.data(), n, nullptr));
  CKC(cudaDeviceSynchronize());
  for (float* d : owned) cudaFree(d);
}

struct Case {
  std::string name;
  size_t out_floats;
  double alg_bytes;      // algorithmic bytes per launch (inputs read once + outputs written once)
  float* out;
  int (*run)(void* self, cudaStream_t s);
  void* self;
};

static std::string g_save, g_ref;

static void report(const Case& c, float us, int iters) {
  std::vector<float> h(c.out_floats);
  CKC(cudaMemcpy(h.data(), c.out, c.out_floats * sizeof(float), cudaMemcpyDeviceToHost));
  double sum = 0.0, amax = 0.0;
  bool finite = true;
  for (float v : h) {
    if (!isfinite(v)) finite = false;
    sum += v;
    amax = fmax(amax, fabs((double)v));
  }
  char extra[256] = "";
  if (!g_ref.empty()) {
    std::string p = g_ref + "/" + c.name + ".bin";
    FILE* f = fopen(p.c_str(), "rb");
    if (f) {
      std::vector<float> r(c.out_floats);
      size_t got = fread(r.data(), sizeof(float), c.out_floats, f);
      fclose(f);
      double md = 0.0;
      if (got == c.out_floats)
        for (size_t i = 0; i < c.out_floats; ++i) md = fmax(md, fabs((double)h[i] - (double)r[i]));
      else md = -1.0;
      snprintf(extra, sizeof extra, " max|delta vs ref|=%.3e", md);
    } else {
      snprintf(extra, sizeof extra, " (no ref file)");
    }
  }
  if (!g_save.empty()) {
    std::string p = g_save + "/" + c.name + ".bin";
    FILE* f = fopen(p.c_str(), "wb");
    if (f) { fwrite(h.data(), sizeof(float), c.out_floats, f); fclose(f); }
  }
  printf("%-28s %10.1f us/launch  %8.1f GB/s(alg)  iters=%d  sum=%.6e max|y|=%.4e%s%s\n", c.name.c_str(), us,
         c.alg_bytes / (us * 1e-6) / 1e9, iters, sum, amax, finite ? "" : "  NON-FINITE", extra);
  fflush(stdout);
}

struct OpArgs {
  lgteun_t* ctx;
  int kind;   // 0 ffn, 1 local, 2 global, 3 mixer, 4 data_step, 5 bicubic4, 6 prior, 7 forward(dead priors), 8 forward(live), 9 patch_embed
  int lgb, N, H, W, bands;
  const float *x, *x2, *x3;
  float* y;
};
static int run_op(void* p, cudaStream_t s) {
  OpArgs* a = (OpArgs*)p;
  switch (a->kind) {
    case 0: return lgteun_op_ffn(a->ctx, 1, a->lgb, 0, a->x, a->y, a->N, a->H, a->W, s);
    case 1: return lgteun_op_local_mixer(a->ctx, 1, a->lgb, 0, a->x, a->y, a->N, a->H, a->W, s);
    case 2: return lgteun_op_global_mixer(a->ctx, 1, a->lgb, 0, a->x, a->y, a->N, a->H, a->W, s);
    case 3: return lgteun_op_mixer(a->ctx, 1, a->lgb, 0, a->x, a->y, a->N, a->H, a->W, s);
    case 4: return lgteun_op_data_step(a->ctx, 0, a->x, a->x2, a->x3, a->y, a->N, a->H / 4, a->W / 4, s);
    case 5: return lgteun_op_bicubic(a->ctx, a->x, a->y, a->N * a->bands, a->H / 4, a->W / 4, 4, 1, s);
    case 6: return lgteun_op_prior(a->ctx, 1, a->x, a->y, a->N, a->H, a->W, s);
    case 7: return lgteun_forward(a->ctx, a->x2, a->x3, a->y, a->N, a->H / 4, a->W / 4, LGTEUN_RUN_DEAD_PRIORS, s);
    case 8: return lgteun_forward(a->ctx, a->x2, a->x3, a->y, a->N, a->H / 4, a->W / 4, 0, s);
    case 10: return lgteun_forward(a->ctx, a->x2, a->x3, a->y, a->N, a->H / 4, a->W / 4, LGTEUN_RUN_DEAD_PRIORS | LGTEUN_NO_GRAPH, s);
    case 9: return lgteun_op_patch_embed(a->ctx, 1, a->x, a->y, a->N, a->H, a->W, s);
  }
  return -1;
}

int main(int argc, char** argv) {
  int bands = 4, batch = 64, pan = 256, iters = 10;
  std::string ops = "ffn,ffn_low,local,local_low,global,global_low,data_step,bicubic,patch_embed,forward";
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto next = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(1); } return std::string(argv[++i]); };
    if (a == "--bands") bands = atoi(next().c_str());
    else if (a == "--batch") batch = atoi(next().c_str());
    else if (a == "--pan") pan = atoi(next().c_str());
    else if (a == "--iters") iters = atoi(next().c_str());
    else if (a == "--ops") ops = next();
    else if (a == "--save") g_save = next();
    else if (a == "--ref") g_ref = next();
    else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 1; }
  }
  CKC(cudaSetDevice(0));
  lgteun_t* ctx = nullptr;
  CKL(lgteun_create(0, bands, 2, &ctx));
  load_random_weights(ctx);
  const int C = 4 * bands, H = pan, W = pan, N = batch;
  const size_t P = (size_t)H * W;
  cudaStream_t s;
  CKC(cudaStreamCreate(&s));
  printf("# opbench bands=%d batch=%d pan=%d iters=%d  (c=%d full res, c=%d half res)\n", bands, N, pan, iters, C, 2 * C);

  auto want = [&](const char* name) {
    std::string pat = std::string(",") + ops + ",";
    return pat.find(std::string(",") + name + ",") != std::string::npos;
  };
  auto timeit = [&](Case& c) {
    for (int i = 0; i < 3; ++i) CKL(c.run(c.self, s));
    CKC(cudaStreamSynchronize(s));
    cudaEvent_t e0, e1;
    CKC(cudaEventCreate(&e0));
    CKC(cudaEventCreate(&e1));
    CKC(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; ++i) CKL(c.run(c.self, s));
    CKC(cudaEventRecord(e1, s));
    CKC(cudaEventSynchronize(e1));
    float ms = 0.f;
    CKC(cudaEventElapsedTime(&ms, e0, e1));
    report(c, ms * 1e3f / iters, iters);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  };

  // feature maps: values like LN outputs / residual streams (O(1))
  struct Spec { const char* name; int kind, lgb, h, w, cin, cout; double bytes_per_px; };
  const Spec specs[] = {
      {"ffn", 0, 0, H, W, C, C, 8.0 * C},
      {"ffn_low", 0, 1, H / 2, W / 2, 2 * C, 2 * C, 16.0 * C},
      {"local", 1, 0, H, W, C / 2, C / 2, 4.0 * C},
      {"local_low", 1, 1, H / 2, W / 2, C, C, 8.0 * C},
      {"global", 2, 0, H, W, C / 2, C / 2, 4.0 * C},
      {"global_low", 2, 1, H / 2, W / 2, C, C, 8.0 * C},
      {"mixer", 3, 0, H, W, C, C, 8.0 * C},
      {"mixer_low", 3, 1, H / 2, W / 2, 2 * C, 2 * C, 16.0 * C},
  };
  for (const Spec& sp : specs) {
    if (!want(sp.name)) continue;
    const size_t px = (size_t)N * sp.h * sp.w;
    float* x = dev_random(px * sp.cin, -1.5f, 1.5f, 11 + sp.kind * 7 + sp.lgb);
    float* y = nullptr;
    CKC(cudaMalloc(&y, px * sp.cout * sizeof(float)));
    CKC(cudaMemset(y, 0, px * sp.cout * sizeof(float)));
    OpArgs a{ctx, sp.kind, sp.lgb, N, sp.h, sp.w, bands, x, nullptr, nullptr, y};
    Case c{sp.name, px * sp.cout, sp.bytes_per_px * (double)px, y, run_op, &a};
    timeit(c);
    cudaFree(x);
    cudaFree(y);
  }
  if (want("data_step") || want("bicubic") || want("forward") || want("forward_nograph") || want("forward_live") || want("prior") || want("patch_embed")) {
    float* ms = dev_random((size_t)N * bands * P / 16, 0.f, 1.f, 101);
    float* pn = dev_random((size_t)N * P, 0.f, 1.f, 102);
    float* z = dev_random((size_t)N * bands * P, 0.f, 1.f, 103);
    float* y = nullptr;
    CKC(cudaMalloc(&y, (size_t)N * bands * P * sizeof(float)));
    if (want("bicubic")) {
      OpArgs a{ctx, 5, 0, N, H, W, bands, ms, nullptr, nullptr, y};
      Case c{"bicubic", (size_t)N * bands * P, 4.0 * N * bands * P * (1.0 + 1.0 / 16), y, run_op, &a};
      timeit(c);
    }
    if (want("data_step")) {
      OpArgs a{ctx, 4, 0, N, H, W, bands, z, ms, pn, y};
      // Z read + Z' written + pan read + ms read
      Case c{"data_step", (size_t)N * bands * P, 4.0 * N * P * (2.0 * bands + 1.0 + bands / 16.0), y, run_op, &a};
      timeit(c);
    }
    if (want("patch_embed")) {
      float* e = nullptr;
      CKC(cudaMalloc(&e, (size_t)N * P * C * sizeof(float)));
      OpArgs a{ctx, 9, 0, N, H, W, bands, z, nullptr, nullptr, e};
      Case c{"patch_embed", (size_t)N * P * C, 4.0 * N * P * (bands + C), e, run_op, &a};
      timeit(c);
      cudaFree(e);
    }
    if (want("prior")) {
      OpArgs a{ctx, 6, 0, N, H, W, bands, z, nullptr, nullptr, y};
      Case c{"prior", (size_t)N * bands * P, 4.0 * N * P * 2.0 * bands, y, run_op, &a};
      timeit(c);
    }
    if (want("forward")) {
      OpArgs a{ctx, 7, 0, N, H, W, bands, nullptr, ms, pn, y};
      Case c{"forward", (size_t)N * bands * P, 4.0 * N * P * (bands + 1.0 + bands / 16.0), y, run_op, &a};
      timeit(c);
    }
    if (want("forward_nograph")) {     // direct launches; with LGTEUN_TIMING=1 the library prints per-kernel-family device times
      OpArgs a{ctx, 10, 0, N, H, W, bands, nullptr, ms, pn, y};
      Case c{"forward_nograph", (size_t)N * bands * P, 4.0 * N * P * (bands + 1.0 + bands / 16.0), y, run_op, &a};
      timeit(c);
    }
    if (want("forward_live")) {
      OpArgs a{ctx, 8, 0, N, H, W, bands, nullptr, ms, pn, y};
      Case c{"forward_live", (size_t)N * bands * P, 4.0 * N * P * (bands + 1.0 + bands / 16.0), y, run_op, &a};
      timeit(c);
    }
    cudaFree(ms); cudaFree(pn); cudaFree(z); cudaFree(y);
  }
  lgteun_destroy(ctx);
  return 0;
}
