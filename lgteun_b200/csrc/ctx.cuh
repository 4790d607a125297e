// ctx.cuh — the handle behind include/lgteun.h (shared by context.cu and train.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/lgteun.h"
#include "common.cuh"

namespace lgctx {

std::string& err_slot();   // thread-local message of the last failing call (context.cu)

inline int fail(int code, const std::string& msg) {
  err_slot() = msg;
  return code;
}
inline int fail_cuda(cudaError_t e, const char* what) {
  err_slot() = std::string(what) + ": " + cudaGetErrorString(e);
  return LGTEUN_ECUDA;
}
#define CK(call)                                                \
  do {                                                          \
    cudaError_t e_ = (call);                                    \
    if (e_ != cudaSuccess) return lgctx::fail_cuda(e_, #call);  \
  } while (0)

struct WeightSlot {
  std::string name;
  int64_t numel;
  const float** slot;   // where the arena pointer is published (a member of lgteun_ctx::wv)
  size_t offset;        // floats from the arena base == offset in the flat parameter / gradient layout
};

struct Derived {        // arena regions computed from loaded tensors
  const float* const* src;
  const float** dst;
  int rows, cols;       // transpose [rows][cols] -> [cols][rows]; rows == 0: pos_emb transpose; rows == -1: FFN fp16 pack
  size_t offset;
  const lg::BlockW* blk;    // FFN pack only: the block whose f0/f1/f2 weights are packed, cols = channels
};

struct Workspace {      // bump-allocated views for one problem size
  float *ms, *pan, *out;          // staging copies of the caller's tensors (graph replays use fixed addresses)
  float *zA, *zB, *resid;
  float *X0, *X1, *X2;            // full-res NHWC maps
  float *L0, *L1;                 // half-res NHWC maps (2C channels)
  float *loc, *spec, *hidden;
};

struct GraphEntry {
  int N, h, w, flags;
  cudaGraph_t graph;              // kept alive: the copy-node handles below belong to it
  cudaGraphExec_t exec;
  cudaGraphNode_t n_ms, n_pan, n_out;
  const float *ms, *pan;
  float* out;
  int launches;
};

struct WeightViews {    // pointers into one flat fp32 buffer, native PyTorch layouts
  lg::DataW dw;
  lg::PriorW prior[lg::kMaxStages];
};

struct TrainState;      // train.cu

}  // namespace lgctx

struct lgteun_ctx {
  int device, B, C, K;
  lgctx::WeightViews wv;
  std::vector<lgctx::WeightSlot> slots;
  std::vector<lgctx::Derived> derived;
  float* arena = nullptr;
  size_t arena_floats = 0;
  size_t flat_floats = 0;              // primary (state_dict) region of the arena: the flat parameter layout
  bool loaded = false;
  float* ws_base = nullptr;
  size_t ws_bytes = 0;
  std::vector<lgctx::GraphEntry> graphs;
  cudaStream_t cap_stream = nullptr;   // capture never runs on the caller's stream (it may be the legacy stream)
  double* metric_acc = nullptr;        // scratch of lgteun_op_metrics
  size_t metric_acc_doubles = 0;
  int last_launches = 0;
  lgctx::TrainState* train = nullptr;  // tape of the training step (train.cu), created on first use
};

namespace lgctx {
void train_destroy(lgteun_ctx* c);     // train.cu
}
