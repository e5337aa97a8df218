// test_rate.cu -- which formulation of "alpha strings a and k (equal weight) differ in exactly one orbital pair"
// sustains the most tests per clock.  All variants accumulate "any hit" over blocks of 8 keys like the block kernel.
//   0: d = a ^ k, popc(d) == 2                         (POPC pipe)
//   1: v = a & ~k, (v & (v - 1)) == 0                   (ALU + FMA pipes; v == 0 passes too: harmless false positive)
//   2: alternate 0 / 1 per key                          (spread over the pipes)
//   3: like 1 with the decrement forced onto the FMA pipe (mad.lo)
//   4: 5 keys by 1, 3 keys by 0
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o test_rate test_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIter = 4096;

__device__ __forceinline__ bool t_popc(unsigned a, unsigned k) { return __popc(a ^ k) == 2; }
__device__ __forceinline__ bool t_pow2(unsigned a, unsigned k) {
  const unsigned v = a & ~k;
  return (v & (v - 1u)) == 0u;
}
__device__ __forceinline__ bool t_pow2_mad(unsigned a, unsigned k) {
  const unsigned v = a & ~k;
  unsigned vm;
  asm("mad.lo.u32 %0, %1, 1, 0xffffffff;" : "=r"(vm) : "r"(v));
  return (v & vm) == 0u;
}

template <int MODE>
__global__ void __launch_bounds__(128) rate_kernel(const unsigned *__restrict__ keys, unsigned *out, int nkeys) {
  __shared__ unsigned q[64 * 128 / 8];
  __shared__ __align__(16) unsigned skeys[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) skeys[i] = keys[(i * 7 + blockIdx.x) % nkeys];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) q[i] = 0;
  __syncthreads();
  const unsigned a = keys[(blockIdx.x * blockDim.x + threadIdx.x) % nkeys];
  unsigned qa = threadIdx.x;
  for (int it = 0; it < kIter; ++it) {
    const uint4 k0 = reinterpret_cast<const uint4 *>(skeys)[(it * 2) & 255];
    const uint4 k1 = reinterpret_cast<const uint4 *>(skeys)[(it * 2 + 1) & 255];
    bool h;
    if (MODE == 0) h = t_popc(a, k0.x) | t_popc(a, k0.y) | t_popc(a, k0.z) | t_popc(a, k0.w) | t_popc(a, k1.x) | t_popc(a, k1.y) | t_popc(a, k1.z) | t_popc(a, k1.w);
    else if (MODE == 1) h = t_pow2(a, k0.x) | t_pow2(a, k0.y) | t_pow2(a, k0.z) | t_pow2(a, k0.w) | t_pow2(a, k1.x) | t_pow2(a, k1.y) | t_pow2(a, k1.z) | t_pow2(a, k1.w);
    else if (MODE == 2) h = t_popc(a, k0.x) | t_pow2(a, k0.y) | t_popc(a, k0.z) | t_pow2(a, k0.w) | t_popc(a, k1.x) | t_pow2(a, k1.y) | t_popc(a, k1.z) | t_pow2(a, k1.w);
    else if (MODE == 3) h = t_pow2_mad(a, k0.x) | t_pow2_mad(a, k0.y) | t_pow2_mad(a, k0.z) | t_pow2_mad(a, k0.w) | t_pow2_mad(a, k1.x) | t_pow2_mad(a, k1.y) | t_pow2_mad(a, k1.z) | t_pow2_mad(a, k1.w);
    else h = t_pow2_mad(a, k0.x) | t_pow2_mad(a, k0.y) | t_popc(a, k0.z) | t_pow2_mad(a, k0.w) | t_pow2_mad(a, k1.x) | t_popc(a, k1.y) | t_pow2_mad(a, k1.z) | t_popc(a, k1.w);
    if (h) {
      q[qa & 1023] = (unsigned)it;
      qa += 128;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = qa + q[threadIdx.x];
}

int main() {
  const int nkeys = 1 << 16;
  unsigned *h = new unsigned[nkeys];
  unsigned s = 12345u;
  for (int i = 0; i < nkeys; ++i) {  // random 15-of-20-bit strings, like folded Fe2S2 alpha strings
    unsigned v = 0;
    int c = 0;
    while (c < 15) {
      s = s * 1664525u + 1013904223u;
      const unsigned b = (s >> 8) % 20u;
      if (!((v >> b) & 1u)) { v |= 1u << b; ++c; }
    }
    h[i] = v;
  }
  unsigned *dk, *dout;
  const int blocks = 148 * 12;
  cudaMalloc(&dk, nkeys * 4);
  cudaMalloc(&dout, blocks * 128 * 4);
  cudaMemcpy(dk, h, nkeys * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char *names[5] = {"0 xor, popc == 2", "1 v = a & ~k, (v & (v-1)) == 0", "2 alternate 0 / 1", "3 like 1, decrement by mad.lo", "4 five by 3, three by 0"};
  for (int mode = 0; mode < 5; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) rate_kernel<0><<<blocks, 128>>>(dk, dout, nkeys);
      else if (mode == 1) rate_kernel<1><<<blocks, 128>>>(dk, dout, nkeys);
      else if (mode == 2) rate_kernel<2><<<blocks, 128>>>(dk, dout, nkeys);
      else if (mode == 3) rate_kernel<3><<<blocks, 128>>>(dk, dout, nkeys);
      else rate_kernel<4><<<blocks, 128>>>(dk, dout, nkeys);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double rows = (double)blocks * 4 * kIter * 8;  // warp-rows: one key against the 32 samples of a warp
    printf("%-34s %8.3f ms  %.2f cycles per key-row per scheduler at 1.965 GHz\n", names[mode], best,
           4.0 * 148 * 1.965e9 * (best * 1e-3) / rows);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
