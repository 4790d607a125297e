// tcprobe — hardware probes behind the design of ffn_cl.cu (channels on TMEM lanes, pixels in TMEM columns):
//   1. tcgen05.ld throughput per SM (32x32b.x8, 4 / 8 / 16 warps)
//   2. tcgen05.mma M=64 with A K-major and B MN-major (no swizzle), accumulator at TMEM lane offset 0 and 16
//      (the "half sub-partition" layout: row i -> lane 32*(i/16) + i%16 [+16])
//   3. tcgen05.mma M=128 with A MN-major (rows beyond the tile alias the next K block) and B K-major
//   4. tcgen05.ld at column offsets that are not multiples of the vector length
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tcprobe tools/tcprobe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../lgteun_b200/csrc/tc_ptx.cuh"
using namespace lg::tc;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

// ---- 1. LDTM throughput ---------------------------------------------------------------------------------------------
__global__ void ldtm_kernel(float* out, long long* cycles, int iters) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tbase, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long c0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float2 a[4], b[4], c[4], d[4];
    const uint32_t col = (uint32_t)((i * 32 + (warp >> 2) * 8) & 255);
    tmem_ld8(t + col, a);
    tmem_ld8(t + col + 64, b);
    tmem_ld8(t + col + 128, c);
    tmem_ld8(t + col + 192, d);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 4; ++k) acc += a[k].x + b[k].y + c[k].x + d[k].y;
  }
  const long long c1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = c1 - c0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// ---- 2-4. MMA layout probes -----------------------------------------------------------------------------------------
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr int K = 64, NB = 48;

// smem: A_k  K-major  [K/8][64][8]      (weights, M = 64)
//       B_mn MN-major [K/8][NB/8][8 k][8 n]
//       W_k  K-major  [K/8][16][8]      (N = 16 operand of the third probe)
__global__ void mma_probe_kernel(const __half* A, const __half* B, const __half* W, float* d0, float* d16, float* d3, float* dsh) {
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  __shared__ alignas(128) __half sA[K * 64], sB[K * NB + 8 * 128], sW[K * 16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tbase, 256);
  for (int i = tid; i < K * 64; i += blockDim.x) {   // A[m][k] -> [k/8][m][k%8]
    const int m = i / K, k = i % K;
    sA[((k >> 3) * 64 + m) * 8 + (k & 7)] = A[i];
  }
  for (int i = tid; i < K * NB; i += blockDim.x) {   // B[n][k] -> [k/8][n/8][k%8][n%8]
    const int n = i / K, k = i % K;
    sB[(((k >> 3) * (NB / 8) + (n >> 3)) * 8 + (k & 7)) * 8 + (n & 7)] = B[i];
  }
  for (int i = tid; i < 8 * 128; i += blockDim.x) sB[K * NB + i] = __float2half(0.f);
  for (int i = tid; i < K * 16; i += blockDim.x) {   // W[n][k] -> [k/8][n][k%8]
    const int n = i / K, k = i % K;
    sW[((k >> 3) * 16 + n) * 8 + (k & 7)] = W[i];
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tbase;
  if (tid == 0) {
    // probe 2: D[64][NB] = A . B^T, at lane offset 0 (cols 0..) and lane offset 16 (cols 64..)
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t ad = umma_desc(smem_u32(sA) + ks * 2 * 64 * 16, 64 * 16, 128);                   // K-major: LBO = K-half stride, SBO = 8-row stride
      const uint64_t bd = umma_desc(smem_u32(sB) + ks * 2 * (NB / 8) * 128, (NB / 8) * 128, 128);     // MN-major: LBO = K-block stride, SBO = MN-block stride
      umma_f16(tmem + 0, ad, bd, idesc_f16(64, NB, 0, 1), ks > 0);
      umma_f16(tmem + 64 + (16u << 16), ad, bd, idesc_f16(64, NB, 0, 1), ks > 0);
    }
    // probe 3: D3[128 (pixel cols; only < NB meaningful)][16] = Bt . W^T with A = sB read MN-major (M = 128)
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t ad = umma_desc(smem_u32(sB) + ks * 2 * (NB / 8) * 128, (NB / 8) * 128, 128);
      const uint64_t wd = umma_desc(smem_u32(sW) + ks * 2 * 16 * 16, 16 * 16, 128);
      umma_f16(tmem + 128, ad, wd, idesc_f16(128, 16, 1, 0), ks > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (warp < 4) {
    const uint32_t t = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < NB; c += 8) {
      float2 v[4];
      tmem_ld8(t + c, v);
      tmem_ld_wait();
      for (int i = 0; i < 4; ++i) { d0[(warp * 32 + lane) * NB + c + 2 * i] = v[i].x; d0[(warp * 32 + lane) * NB + c + 2 * i + 1] = v[i].y; }
      tmem_ld8(t + 64 + c, v);
      tmem_ld_wait();
      for (int i = 0; i < 4; ++i) { d16[(warp * 32 + lane) * NB + c + 2 * i] = v[i].x; d16[(warp * 32 + lane) * NB + c + 2 * i + 1] = v[i].y; }
    }
    for (int c = 0; c < 16; c += 8) {
      float2 v[4];
      tmem_ld8(t + 128 + c, v);
      tmem_ld_wait();
      for (int i = 0; i < 4; ++i) { d3[(warp * 32 + lane) * 16 + c + 2 * i] = v[i].x; d3[(warp * 32 + lane) * 16 + c + 2 * i + 1] = v[i].y; }
    }
    // probe 4: shifted windows (column offsets 1, 2, 3) of D at lane offset 0
    for (int sft = 1; sft <= 3; ++sft) {
      float2 v[4];
      tmem_ld8(t + sft, v);
      tmem_ld_wait();
      for (int i = 0; i < 4; ++i) { dsh[((sft - 1) * 128 + warp * 32 + lane) * 8 + 2 * i] = v[i].x; dsh[((sft - 1) * 128 + warp * 32 + lane) * 8 + 2 * i + 1] = v[i].y; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}


__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
// ---- 5. latency of accumulation chains of small MMAs ---------------------------------------------------------------------
// mode 0: n MMAs into one accumulator; 1: n/2 + n/2 into two accumulators (lane offset 0 / 16), one after the other;
// 2: the same two chains interleaved; 3: two chains in different TMEM columns, interleaved
template <int M, int N>
__global__ void chain_kernel(long long* cycles, int n, int mode) {
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  __shared__ alignas(128) __half sA[128 * 64], sB[64 * 64];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tbase, 512);
  for (int i = tid; i < 128 * 64; i += blockDim.x) sA[i] = __float2half(0.f);
  for (int i = tid; i < 64 * 64; i += blockDim.x) sB[i] = __float2half(0.f);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint64_t ad = umma_desc(smem_u32(sA), M * 16, 128), bd = umma_desc(smem_u32(sB), 64 * 16, 128);
    constexpr uint32_t id = idesc_f16(M, N, 0, 0);
    uint32_t ph = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long c0 = clock64();
      if (elect_one()) {
        if (mode == 0) {
#pragma unroll 2
          for (int i = 0; i < n; ++i) umma_f16(tbase, ad, bd, id, 1);
        } else if (mode == 1) {
          for (int i = 0; i < n / 2; ++i) umma_f16(tbase, ad, bd, id, 1);
          for (int i = 0; i < n / 2; ++i) umma_f16(tbase + (16u << 16), ad, bd, id, 1);
        } else if (mode == 2) {
          for (int i = 0; i < n / 2; ++i) {
            umma_f16(tbase, ad, bd, id, 1);
            umma_f16(tbase + (16u << 16), ad, bd, id, 1);
          }
        } else {
          for (int i = 0; i < n / 2; ++i) {
            umma_f16(tbase, ad, bd, id, 1);
            umma_f16(tbase + 256u, ad, bd, id, 1);
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait_spin(&bar, ph);
      ph ^= 1;
      const long long c1 = clock64();
      if (tid == 0) cycles[rep] = c1 - c0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}
template <int M, int N>
static void run_chain(const char* name, long long* cyc) {
  for (int mode = 0; mode < 4; ++mode) {
    if (M == 128 && (mode == 1 || mode == 2)) continue;
    for (int n : {2, 14, 26}) {
      chain_kernel<M, N><<<1, 128>>>(cyc, n, mode);
      CK(cudaDeviceSynchronize());
      long long h[3];
      CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
      printf("chain %s mode %d n=%2d: %lld clk (%.1f per MMA)\n", name, mode, n, h[2], (double)h[2] / n);
    }
  }
}

int main() {
  // ---- 1 ----
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float* out;
  long long* cyc;
  CK(cudaMalloc(&out, sizeof(float) * sms * 1024));
  CK(cudaMalloc(&cyc, sizeof(long long) * sms));
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 4096;
    ldtm_kernel<<<sms, warps * 32>>>(out, cyc, iters);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    const double bytes = (double)warps * iters * 4 * 32 * 8 * 4;
    printf("ldtm x8: %2d warps/SM  %.1f B/clk/SM  (%.1f clk per warp-instruction)\n", warps, bytes / (double)h[0],
           (double)h[0] / (iters * 4.0));
  }
  run_chain<64, 40>("M=64 N=40", cyc);
  run_chain<128, 48>("M=128 N=48", cyc);
  run_chain<128, 16>("M=128 N=16", cyc);
  run_chain<128, 64>("M=128 N=64", cyc);
  // ---- 2-4 ----
  std::vector<__half> hA(64 * K), hB(NB * K), hW(16 * K);
  std::vector<float> fA(64 * K), fB(NB * K), fW(16 * K);
  for (int m = 0; m < 64; ++m)
    for (int k = 0; k < K; ++k) { fA[m * K + k] = (float)((m * 3 + k) % 7 - 3); hA[m * K + k] = __float2half(fA[m * K + k]); }
  for (int n = 0; n < NB; ++n)
    for (int k = 0; k < K; ++k) { fB[n * K + k] = (float)((n + 2 * k + n * k) % 5 - 2); hB[n * K + k] = __float2half(fB[n * K + k]); }
  for (int n = 0; n < 16; ++n)
    for (int k = 0; k < K; ++k) { fW[n * K + k] = (float)((n * 5 + k) % 9 - 4); hW[n * K + k] = __float2half(fW[n * K + k]); }
  __half *dA, *dB, *dW;
  float *d0, *d16, *d3, *dsh;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dW, hW.size() * 2));
  CK(cudaMalloc(&d0, 128 * NB * 4)); CK(cudaMalloc(&d16, 128 * NB * 4)); CK(cudaMalloc(&d3, 128 * 16 * 4)); CK(cudaMalloc(&dsh, 3 * 128 * 8 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(d0, 0xff, 128 * NB * 4)); CK(cudaMemset(d16, 0xff, 128 * NB * 4));
  mma_probe_kernel<<<1, 128>>>(dA, dB, dW, d0, d16, d3, dsh);
  CK(cudaDeviceSynchronize());
  std::vector<float> r0(128 * NB), r16(128 * NB), r3(128 * 16), rsh(3 * 128 * 8);
  CK(cudaMemcpy(r0.data(), d0, r0.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r16.data(), d16, r16.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r3.data(), d3, r3.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(rsh.data(), dsh, rsh.size() * 4, cudaMemcpyDeviceToHost));
  int bad0 = 0, bad16 = 0, bad3 = 0, badsh = 0;
  for (int m = 0; m < 64; ++m)
    for (int n = 0; n < NB; ++n) {
      float ref = 0.f;
      for (int k = 0; k < K; ++k) ref += fA[m * K + k] * fB[n * K + k];
      const int l0 = 32 * (m / 16) + m % 16;
      if (r0[l0 * NB + n] != ref) { if (bad0 < 4) printf("  M64@0  m=%d n=%d got %g want %g\n", m, n, r0[l0 * NB + n], ref); ++bad0; }
      if (r16[(l0 + 16) * NB + n] != ref) { if (bad16 < 4) printf("  M64@16 m=%d n=%d got %g want %g\n", m, n, r16[(l0 + 16) * NB + n], ref); ++bad16; }
      for (int sft = 1; sft <= 3; ++sft)
        if (n >= sft && n < sft + 8 && rsh[((sft - 1) * 128 + l0) * 8 + n - sft] != ref) ++badsh;
    }
  for (int p = 0; p < NB; ++p)
    for (int n = 0; n < 16; ++n) {
      float ref = 0.f;
      for (int k = 0; k < K; ++k) ref += fB[p * K + k] * fW[n * K + k];
      if (r3[p * 16 + n] != ref) { if (bad3 < 4) printf("  M128 A-MN p=%d n=%d got %g want %g\n", p, n, r3[p * 16 + n], ref); ++bad3; }
    }
  printf("M=64 A K-major x B MN-major, D at lane 0: %s (%d bad)\n", bad0 ? "FAIL" : "ok", bad0);
  printf("M=64 A K-major x B MN-major, D at lane 16: %s (%d bad)\n", bad16 ? "FAIL" : "ok", bad16);
  printf("M=128 A MN-major x B K-major: %s (%d bad)\n", bad3 ? "FAIL" : "ok", bad3);
  printf("tcgen05.ld at column offsets 1..3: %s (%d bad)\n", badsh ? "FAIL" : "ok", badsh);
  return (bad0 || bad16 || bad3 || badsh) ? 1 : 0;
}
