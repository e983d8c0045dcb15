// misc.cu -- ranking, buffer row gather/scatter, flat SGD (sm_100a).
#include <float.h>

#include "common.cuh"

namespace b200ocl {
namespace {

// ----------------------------------------------------------------------------- rank_desc
// One CTA, shared-memory bitonic sort of (order-preserving score bits, index) 64-bit keys.
// Descending score, ties lowest index first == ascending sort of (~orderable(score), index).
__device__ __forceinline__ unsigned int orderable(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(1024) rank_desc_kernel(const float* __restrict__ a, float sa,
                                                         const float* __restrict__ b, float sb, int n, int npad,
                                                         long long* __restrict__ idx_out, int n_out,
                                                         float* __restrict__ score_out) {
  extern __shared__ __align__(16) unsigned long long keys[];
  const int tid = threadIdx.x;
  for (int i = tid; i < npad; i += blockDim.x) {
    unsigned long long key = ~0ull;
    if (i < n) {
      float s = a[i] * sa;
      if (b) s += b[i] * sb;
      if (s == 0.f) s = 0.f;  // -0 -> +0 so that equal scores compare equal
      if (score_out) score_out[i] = s;
      key = (static_cast<unsigned long long>(~orderable(s)) << 32) | static_cast<unsigned int>(i);
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int k2 = 2; k2 <= npad; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < npad / 2; t += blockDim.x) {
        const int lo = ((t / j) * 2 * j) + (t % j);
        const int hi = lo + j;
        const bool up = ((lo & k2) == 0);
        const unsigned long long x = keys[lo], y = keys[hi];
        if (up ? (x > y) : (x < y)) {
          keys[lo] = y;
          keys[hi] = x;
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n_out; i += blockDim.x) idx_out[i] = (long long)(keys[i] & 0xffffffffull);
}

// ----------------------------------------------------------------------------- rows
// One CTA per row (grid-stride), 16-byte vectors when the row size and all bases allow it.
template <bool SCATTER, typename VecT>
__global__ void __launch_bounds__(256) move_rows_kernel(const unsigned char* __restrict__ src,
                                                        const long long* __restrict__ idx, int n_rows,
                                                        size_t row_bytes, unsigned char* __restrict__ dst) {
  const size_t nvec = row_bytes / sizeof(VecT);
  for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
    const long long ir = idx[r];
    const VecT* s = reinterpret_cast<const VecT*>(src + (SCATTER ? (size_t)r : (size_t)ir) * row_bytes);
    VecT* d = reinterpret_cast<VecT*>(dst + (SCATTER ? (size_t)ir : (size_t)r) * row_bytes);
    for (size_t v = threadIdx.x; v < nvec; v += blockDim.x) d[v] = s[v];
  }
}

template <bool SCATTER>
int move_rows(const void* src, const int64_t* idx, int n_rows, size_t row_bytes, void* dst, cudaStream_t stream) {
  if (n_rows == 0) return B200OCL_OK;
  const bool vec16 = (row_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  const int grid = n_rows < 8 * sm_count() ? n_rows : 8 * sm_count();
  B200OCL_PROF("move_rows", 2.0 * n_rows * (double)row_bytes, stream);
  if (vec16)
    move_rows_kernel<SCATTER, uint4><<<grid, 256, 0, stream>>>(static_cast<const unsigned char*>(src),
                                                               reinterpret_cast<const long long*>(idx), n_rows,
                                                               row_bytes, static_cast<unsigned char*>(dst));
  else
    move_rows_kernel<SCATTER, unsigned int><<<grid, 256, 0, stream>>>(static_cast<const unsigned char*>(src),
                                                                      reinterpret_cast<const long long*>(idx), n_rows,
                                                                      row_bytes, static_cast<unsigned char*>(dst));
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

// ----------------------------------------------------------------------------- SGD
__global__ void __launch_bounds__(256) sgd_kernel(const float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ out, size_t n, float lr, float wd) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float w = p[i];
    float gi = g[i];
    if (wd != 0.f) gi = fmaf(wd, w, gi);
    out[i] = w - lr * gi;
  }
}

}  // namespace
}  // namespace b200ocl

extern "C" {

int b200ocl_rank_desc(const float* a, float sa, const float* b, float sb, int n, int64_t* idx_out, int n_out,
                      float* score_out, void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200OCL_CHECK_ARG(a && idx_out, "null pointer");
  B200OCL_CHECK_ARG(n >= 0 && n_out >= 0 && n_out <= n, "need 0 <= n_out <= n");
  if (n > 4096) {
    set_error("b200ocl_rank_desc: n=%d exceeds the limit of 4096", n);
    return B200OCL_EUNSUPPORTED;
  }
  if (n == 0) return B200OCL_OK;
  int npad = 2;
  while (npad < n) npad <<= 1;
  const int threads = npad / 2 < 32 ? 32 : (npad / 2 > 1024 ? 1024 : npad / 2);
  B200OCL_PROF("rank_desc", 12.0 * n, stream);
  rank_desc_kernel<<<1, threads, (size_t)npad * sizeof(unsigned long long), stream>>>(
      a, sa, b, sb, n, npad, reinterpret_cast<long long*>(idx_out), n_out, score_out);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

int b200ocl_gather_rows(const void* src, const int64_t* idx, int n_rows, size_t row_bytes, void* dst, void* stream) {
  using namespace b200ocl;
  B200OCL_CHECK_ARG(n_rows >= 0 && row_bytes % 4 == 0, "need n_rows >= 0 and row_bytes % 4 == 0");
  B200OCL_CHECK_ARG(n_rows == 0 || (src && idx && dst), "null pointer");
  return move_rows<false>(src, idx, n_rows, row_bytes, dst, static_cast<cudaStream_t>(stream));
}

int b200ocl_scatter_rows(const void* src, const int64_t* idx, int n_rows, size_t row_bytes, void* dst, void* stream) {
  using namespace b200ocl;
  B200OCL_CHECK_ARG(n_rows >= 0 && row_bytes % 4 == 0, "need n_rows >= 0 and row_bytes % 4 == 0");
  B200OCL_CHECK_ARG(n_rows == 0 || (src && idx && dst), "null pointer");
  return move_rows<true>(src, idx, n_rows, row_bytes, dst, static_cast<cudaStream_t>(stream));
}

int b200ocl_sgd_step(const float* p, const float* g, float* out, size_t n, float lr, float wd, void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200OCL_CHECK_ARG(n == 0 || (p && g && out), "null pointer");
  if (n == 0) return B200OCL_OK;
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)8 * sm_count();
  if (blocks > cap) blocks = cap;
  B200OCL_PROF("sgd", 12.0 * n, stream);
  sgd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p, g, out, n, lr, wd);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

}  // extern "C"
