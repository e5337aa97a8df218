// convert.cu -- ONV <-> occupation-vector conversions feeding the ansatz.
//
// Replaces K7 (onv_to_tensor_kernel, cuda/kernel.cu:39-64; cpu_tensor.cpp:46-88) and K8
// (tensor_to_onv_kernel, cuda/kernel.cu:14-37; cpu_tensor.cpp:8-44).
#include "common.cuh"

namespace pynqs {

// one thread per output element; consecutive threads write consecutive elements
template <typename T>
__global__ void __launch_bounds__(256)
onv_to_tensor_kernel(const u64 *__restrict__ onv, T *__restrict__ out, long long n, int sorb, int L) {
  const long long total = n * sorb;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / sorb;
    const int k = (int)(i - s * sorb);
    const u64 w = __ldg(onv + s * L + (k >> 6));
    out[i] = ((w >> (k & 63)) & 1ull) ? (T)1.0 : (T)-1.0;
  }
}

// one thread per output byte: gathers 8 occupation flags (only the value 1 sets a bit, as in the
// reference's `== 1` test, cpu_tensor.cpp:36)
__global__ void __launch_bounds__(256)
tensor_to_onv_kernel(const unsigned char *__restrict__ states, unsigned char *__restrict__ onv, long long n, int sorb, int L) {
  const long long total = n * 8 * L;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / (8 * L);
    const int byte = (int)(i - s * 8 * L);
    unsigned v = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int k = byte * 8 + b;
      if (k < sorb && __ldg(states + s * sorb + k) == 1) v |= 1u << b;
    }
    onv[i] = (unsigned char)v;
  }
}

static inline unsigned grid_cap(long long total) {
  long long want = (total + 255) / 256;
  return (unsigned)(want < 148LL * 32 ? (want < 1 ? 1 : want) : 148LL * 32);
}

int launch_onv_to_tensor(const u64 *onv, void *out, int dtype, long long n, int sorb, cudaStream_t st) {
  if (n == 0) return 0;
  const int L = (sorb - 1) / 64 + 1;
  if (dtype == 1) onv_to_tensor_kernel<double><<<grid_cap(n * sorb), 256, 0, st>>>(onv, (double *)out, n, sorb, L);
  else onv_to_tensor_kernel<float><<<grid_cap(n * sorb), 256, 0, st>>>(onv, (float *)out, n, sorb, L);
  count_launch();
  return check_launch("onv_to_tensor_kernel");
}

int launch_tensor_to_onv(const unsigned char *states, unsigned char *onv, long long n, int sorb, cudaStream_t st) {
  if (n == 0) return 0;
  const int L = (sorb - 1) / 64 + 1;
  tensor_to_onv_kernel<<<grid_cap(n * 8 * L), 256, 0, st>>>(states, onv, n, sorb, L);
  count_launch();
  return check_launch("tensor_to_onv_kernel");
}

}  // namespace pynqs
