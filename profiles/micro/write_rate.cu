// write_rate.cu -- ceiling of a pure HBM write stream shaped like get_comb_hij_fused's output:
//   n CTAs, CTA s writes rows [s*M, (s+1)*M) of two arrays of 8-byte elements (comb and Hmat), every warp store covering
//   32 consecutive rows.  Variants: (a) the same rows written by a flat grid-stride loop, (b) one CTA per sample with
//   1 KB per warp and step (the old enumerate loop), (c) one CTA per sample, warps walking separate ranges.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_rate write_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

__global__ void __launch_bounds__(256) flat(u64 *a, double *b, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    a[i] = i;
    b[i] = (double)i;
  }
}

__global__ void __launch_bounds__(256) per_sample_chunks(u64 *a, double *b, int M) {
  const size_t base = (size_t)blockIdx.x * M;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp * 128; c < M; c += 8 * 128) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = c + lane + 32 * j;
      if (m < M) {
        a[base + m] = m;
        b[base + m] = (double)m;
      }
    }
  }
}

__global__ void __launch_bounds__(256) per_sample_ranges(u64 *a, double *b, int M) {
  const size_t base = (size_t)blockIdx.x * M;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lo = (int)((long long)M * warp / 8), hi = (int)((long long)M * (warp + 1) / 8);
  for (int m = lo + lane; m < hi; m += 32) {
    a[base + m] = m;
    b[base + m] = (double)m;
  }
}

int main() {
  const int n = 32768, M = 7876;
  const size_t total = (size_t)n * M;
  u64 *a;
  double *b;
  cudaMalloc(&a, total * 8);
  cudaMalloc(&b, total * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int v = 0; v < 3; ++v) {
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      cudaEventRecord(e0);
      if (v == 0) flat<<<148 * 8, 256>>>(a, b, total);
      else if (v == 1) per_sample_chunks<<<n, 256>>>(a, b, M);
      else per_sample_ranges<<<n, 256>>>(a, b, M);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (it > 0 && ms < best) best = ms;
    }
    const char *names[] = {"flat grid-stride", "CTA per sample, 1 KB chunks per warp", "CTA per sample, one range per warp"};
    printf("%-40s %.4f ms  %.0f GB/s written (%s)\n", names[v], best, total * 16.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
