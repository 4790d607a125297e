// metrics.cu — the full-reference quality metrics of the reference's eval loop, on the device (SURVEY.md §8f rank 2:
// after the forward drops to ~0.3 ms per pair, the per-image numpy metrics of Base_model.test dominate an evaluation).
//   reference: models/base/metrics.py  psnr :39-48, sam :22-35, ergas :166-182; inputs are de-normalised first
//   (img * (2**bit_depth - .5) in float32, dataset/utils.py:252-263) and promoted to float64 like the reference does.
// Two launches: per-image partial sums accumulated with fp64 atomics, then one thread per image finalises.
#include <math.h>
#include "common.cuh"

namespace lg {

// acc layout per image: [0] sum (p-g)^2 over everything, [1] sum arccos, [2 .. 2+B) per-band sum (p-g)^2, [2+B .. 2+2B) per-band sum g
template <int B>
__global__ void __launch_bounds__(256) metrics_accumulate_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                                  double* __restrict__ acc, int HW, float max_value) {
  const int n = blockIdx.y;
  const float* p = pred + (size_t)n * B * HW;
  const float* g = gt + (size_t)n * B * HW;
  double se = 0.0, ang = 0.0, seb[B], sgb[B];
#pragma unroll
  for (int b = 0; b < B; ++b) seb[b] = sgb[b] = 0.0;
  const double eps = 2.220446049250313e-16;              // np.finfo(np.float64).eps
  for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) {
    double dot = 0.0, np2 = 0.0, ng2 = 0.0;
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const double pv = (double)(__ldg(p + (size_t)b * HW + i) * max_value);   // float32 de-normalisation, then float64
      const double gv = (double)(__ldg(g + (size_t)b * HW + i) * max_value);
      const double d = pv - gv;
      // explicit mul/add (no FMA contraction): per-pixel values round exactly like the reference's numpy expressions,
      // which matters for arccos next to 1 (nearly identical spectra)
      seb[b] = __dadd_rn(seb[b], __dmul_rn(d, d));
      sgb[b] += gv;
      dot = __dadd_rn(dot, __dmul_rn(pv, gv));
      np2 = __dadd_rn(np2, __dmul_rn(pv, pv));
      ng2 = __dadd_rn(ng2, __dmul_rn(gv, gv));
    }
    double c = dot / __dadd_rn(__dmul_rn(sqrt(np2), sqrt(ng2)), eps);
    c = fmin(fmax(c, 0.0), 1.0);
    ang += acos(c);
  }
#pragma unroll
  for (int b = 0; b < B; ++b) se += seb[b];
  // warp reduce, then one atomic per warp and quantity
  auto wsum = [](double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  double* a = acc + (size_t)n * (2 + 2 * B);
  se = wsum(se);
  ang = wsum(ang);
  if ((threadIdx.x & 31) == 0) { atomicAdd(a, se); atomicAdd(a + 1, ang); }
#pragma unroll
  for (int b = 0; b < B; ++b) {
    const double s1 = wsum(seb[b]), s2 = wsum(sgb[b]);
    if ((threadIdx.x & 31) == 0) { atomicAdd(a + 2 + b, s1); atomicAdd(a + 2 + B + b, s2); }
  }
}

__global__ void metrics_finalize_kernel(const double* __restrict__ acc, double* __restrict__ out, int N, int B, int HW,
                                        double dynamic_range) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double eps = 2.220446049250313e-16;
  const double* a = acc + (size_t)n * (2 + 2 * B);
  const double mse = a[0] / ((double)B * HW);
  out[3 * n + 0] = (mse <= 1e-10) ? INFINITY : 20.0 * log10(dynamic_range / (sqrt(mse) + eps));   // psnr
  out[3 * n + 1] = a[1] / HW;                                                                      // sam
  double e = 0.0;
  for (int b = 0; b < B; ++b) {
    const double mean_g = a[2 + B + b] / HW, mse_b = a[2 + b] / HW;
    e += mse_b / (mean_g * mean_g + eps);
  }
  out[3 * n + 2] = 100.0 / 4.0 * sqrt(e / B);                                                      // ergas (scale 4)
}

// data_normalize (dataset/utils.py:232-248): img / (2**bit_depth - .5), a true fp32 division like torch's
__global__ void normalize_kernel(const float* __restrict__ raw, float* __restrict__ out, size_t n, float max_value) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __fdiv_rn(__ldg(raw + i), max_value);
}
cudaError_t launch_normalize(const float* raw, float* out, size_t n, float max_value, cudaStream_t s) {
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  normalize_kernel<<<(unsigned)blocks, 256, 0, s>>>(raw, out, n, max_value);
  return cudaGetLastError();
}

// torch2np + data_denormalize (models/base/utils.py:28-39, dataset/utils.py:252-263): NCHW -> NHWC, times max_value.
// A 32-pixel x C tile goes through shared memory so both the NCHW reads and the NHWC writes are coalesced.
template <int C>
__global__ void __launch_bounds__(256) to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int HW, float scale) {
  __shared__ float tile[C][256 + 1];
  const int n = blockIdx.y, p0 = blockIdx.x * 256;
  const float* s = src + (size_t)n * C * HW;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int p = p0 + threadIdx.x;
    tile[c][threadIdx.x] = p < HW ? __ldg(s + (size_t)c * HW + p) * scale : 0.f;
  }
  __syncthreads();
  float* d = dst + ((size_t)n * HW + p0) * C;
  const int lim = (HW - p0 < 256 ? HW - p0 : 256) * C;
  for (int i = threadIdx.x; i < lim; i += 256) d[i] = tile[i % C][i / C];
}
cudaError_t launch_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, float scale, cudaStream_t s) {
  const int HW = H * W;
  dim3 grid((HW + 255) / 256, N);
  switch (C) {
    case 1: to_nhwc_kernel<1><<<grid, 256, 0, s>>>(src, dst, HW, scale); break;
    case 4: to_nhwc_kernel<4><<<grid, 256, 0, s>>>(src, dst, HW, scale); break;
    case 8: to_nhwc_kernel<8><<<grid, 256, 0, s>>>(src, dst, HW, scale); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_metrics(const float* pred, const float* gt, double* acc, double* out, int N, int B, int H, int W,
                           float max_value, cudaStream_t s) {
  const int HW = H * W;
  cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)N * (2 + 2 * B) * sizeof(double), s);
  if (e != cudaSuccess) return e;
  int bx = (HW + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid(bx, N);
  if (B == 4) metrics_accumulate_kernel<4><<<grid, 256, 0, s>>>(pred, gt, acc, HW, max_value);
  else if (B == 8) metrics_accumulate_kernel<8><<<grid, 256, 0, s>>>(pred, gt, acc, HW, max_value);
  else return cudaErrorInvalidValue;
  metrics_finalize_kernel<<<(N + 127) / 128, 128, 0, s>>>(acc, out, N, B, HW, (double)max_value);
  return cudaGetLastError();
}

}  // namespace lg
