// popc_rate.cu -- issue-rate microbenchmark behind the design of the block-scan kernel (profiles/micro/README):
// how many (key, sample) distance tests per clock an SM sustains with
//   A: XOR + POPC + ISETP + predicated counter update     (the test as written)
//   B: POPC-free "exactly two bits set" test               (d & (d-1), twice)
//   C: A with the hit recorded by a predicated shared-memory store (the kernel's inner loop)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o popc_rate popc_rate.cu ; run on one B200.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIter = 4096, kUnroll = 8;

template <int MODE>
__global__ void __launch_bounds__(256) rate_kernel(const unsigned *__restrict__ keys, unsigned *out, int nkeys) {
  __shared__ unsigned q[64 * 256 / 8];
  __shared__ unsigned skeys[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) skeys[i] = keys[i % nkeys];
  __syncthreads();
  const unsigned a = keys[(blockIdx.x * blockDim.x + threadIdx.x) % nkeys] | 1u;
  unsigned cnt = 0, qa = threadIdx.x;
  for (int it = 0; it < kIter; ++it) {
    const uint4 k0 = reinterpret_cast<const uint4 *>(skeys)[(it * 2) & 255];
    const uint4 k1 = reinterpret_cast<const uint4 *>(skeys)[(it * 2 + 1) & 255];
    const unsigned k[kUnroll] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const unsigned d = a ^ k[u];
      if (MODE == 0) {
        if (__popc(d) == 2) cnt++;
      } else if (MODE == 1) {
        const unsigned t = d & (d - 1u);
        if (t != 0u && (t & (t - 1u)) == 0u) cnt++;
      } else if (MODE == 2) {
        if (__popc(d) == 2) {
          q[qa & 2047] = (unsigned)(it * kUnroll + u);
          qa += 256;
        }
      } else if (MODE == 3) {
        const unsigned t = d & (d - 1u);
        if (t != 0u && (t & (t - 1u)) == 0u) {
          q[qa & 2047] = (unsigned)(it * kUnroll + u);
          qa += 256;
        }
      } else {  // bit mask of the hits of this iteration, no store
        if (__popc(d) == 2) cnt |= 1u << u;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (MODE == 2 || MODE == 3) ? cnt + qa + q[threadIdx.x] : cnt + qa;
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
  const int blocks = 148 * 8;
  cudaMalloc(&dk, nkeys * 4);
  cudaMalloc(&dout, blocks * 256 * 4);
  cudaMemcpy(dk, h, nkeys * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char *names[5] = {"A xor+popc+setp+@p add", "B popc-free two-bit test", "C xor+popc+setp+@p sts+@p add", "D popc-free + @p sts + @p add",
                          "E xor+popc+setp+@p or-mask"};
  for (int mode = 0; mode < 5; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) rate_kernel<0><<<blocks, 256>>>(dk, dout, nkeys);
      else if (mode == 1) rate_kernel<1><<<blocks, 256>>>(dk, dout, nkeys);
      else if (mode == 2) rate_kernel<2><<<blocks, 256>>>(dk, dout, nkeys);
      else if (mode == 3) rate_kernel<3><<<blocks, 256>>>(dk, dout, nkeys);
      else rate_kernel<4><<<blocks, 256>>>(dk, dout, nkeys);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double tests = (double)blocks * 256 * kIter * kUnroll;
    printf("%-34s %8.3f ms  %7.1f G tests/s  = %.2f warp-rows/clk/SM at 1.965 GHz (%.2f cycles per 32 tests per SMSP)\n", names[mode], best,
           tests / best / 1e6, tests / 32 / (best * 1e-3) / 148 / 1.965e9, 4.0 * 148 * 1.965e9 * (best * 1e-3) / (tests / 32));
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
