// tc_selftest.cu -- minimal tcgen05 (kind::tf32, SS operands, TMEM accumulator) GEMM used to validate
// the descriptor encodings of umma.cuh on the device before the tensor-core convolution relies on them.
//   D[128, N] = A[128, K] * B[N, K]^T     K multiple of 32, N multiple of 16 in [16, 256]
// mode 0: single TF32 pass (inputs truncated by the hardware); mode 1: 3xTF32 split
// (hi*hi + hi*lo + lo*hi), the numerics the convolution uses.
#include "common.cuh"
#include "umma.cuh"

namespace b200ocl {
namespace {

__global__ void __launch_bounds__(128) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ D, int N, int K, int mode,
                                                            int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // [A_hi | A_lo | B_hi | B_lo], each a swizzled K-major tile of 32 fp32 per row
  float* sAh = reinterpret_cast<float*>(smem_raw);
  float* sAl = sAh + 128 * 32;
  float* sBh = sAl + 128 * 32;
  float* sBl = sBh + 256 * 32;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) umma::tmem_alloc(&tmem_base_slot, ncols);
  if (tid == 0) {
    umma::mbar_init(&mbar, 1);
    umma::fence_mbar_init();
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t idesc = umma::make_idesc_tf32(128, N);
  uint32_t phase = 0;
  bool ok = true;

  for (int kb = 0; kb < K / 32; ++kb) {
    // stage the K block: 16-byte chunks into the swizzled layout, split into hi / lo
    for (int idx = tid; idx < 128 * 8; idx += 128) {
      const int r = idx >> 3, c = idx & 7;
      const float4 v = *reinterpret_cast<const float4*>(A + (size_t)r * K + kb * 32 + c * 4);
      float4 h, l;
      umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y);
      umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w);
      const int off = umma::sw128_offset_f32(r, c);
      *reinterpret_cast<float4*>(sAh + off) = (mode == 0) ? v : h;
      *reinterpret_cast<float4*>(sAl + off) = l;
    }
    for (int idx = tid; idx < N * 8; idx += 128) {
      const int r = idx >> 3, c = idx & 7;
      const float4 v = *reinterpret_cast<const float4*>(B + (size_t)r * K + kb * 32 + c * 4);
      float4 h, l;
      umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y);
      umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w);
      const int off = umma::sw128_offset_f32(r, c);
      *reinterpret_cast<float4*>(sBh + off) = (mode == 0) ? v : h;
      *reinterpret_cast<float4*>(sBl + off) = l;
    }
    umma::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_thread_sync();
      const uint64_t dAh = umma::make_smem_desc_sw128(umma::smem_u32(sAh));
      const uint64_t dAl = umma::make_smem_desc_sw128(umma::smem_u32(sAl));
      const uint64_t dBh = umma::make_smem_desc_sw128(umma::smem_u32(sBh));
      const uint64_t dBl = umma::make_smem_desc_sw128(umma::smem_u32(sBl));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)(k * 32 >> 4);   // 32 bytes per K step, in 16-byte units
        umma::mma_tf32_ss(tmem, dAh + adv, dBh + adv, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        if (mode == 1) {
          umma::mma_tf32_ss(tmem, dAh + adv, dBl + adv, idesc, 1u);
          umma::mma_tf32_ss(tmem, dAl + adv, dBh + adv, idesc, 1u);
        }
      }
      umma::mma_commit(&mbar);
    }
    // everybody waits for the MMAs before the staging buffers are overwritten
    if (!umma::mbar_wait(&mbar, phase)) ok = false;
    phase ^= 1;
    __syncthreads();
    if (!ok) break;
  }
  umma::fence_after_thread_sync();
  if (ok) {
    // warp w owns TMEM lanes 32w .. 32w+31 = output rows
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = v[j];
    }
  }
  if (tid == 0) *status = ok ? 0 : 1;
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

// Window test: the A operand is NOT a packed tile but a window into a larger swizzled buffer of 128-byte
// rows ("patch"): tile row 8g + r is patch row start_row + g * sbo_rows + r.  The patch is stored with the
// swizzle keyed on the absolute row index (row & 7), the start address is not 1024-byte aligned when
// start_row % 8 != 0, and the descriptor's base-offset field (bits 49..51) is set to start_row & 7 when
// base_off_mode == 1.  This is what lets a convolution feed the tensor core straight from a halo patch.
__global__ void __launch_bounds__(128) umma_window_kernel(const float* __restrict__ P, const float* __restrict__ B,
                                                          float* __restrict__ D, int rows, int start_row, int sbo_rows,
                                                          int base_off_mode, int N, int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* sP = reinterpret_cast<float*>(smem_raw);           // rows x 32 fp32
  float* sB = sP + (size_t)((rows + 7) / 8 * 8) * 32;       // 1024-byte aligned
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) umma::tmem_alloc(&tmem_base_slot, ncols);
  if (tid == 0) {
    umma::mbar_init(&mbar, 1);
    umma::fence_mbar_init();
  }
  for (int idx = tid; idx < rows * 8; idx += 128) {
    const int r = idx >> 3, c = idx & 7;
    *reinterpret_cast<float4*>(sP + umma::sw128_offset_f32(r, c)) = *reinterpret_cast<const float4*>(P + (size_t)r * 32 + c * 4);
  }
  for (int idx = tid; idx < N * 8; idx += 128) {
    const int r = idx >> 3, c = idx & 7;
    *reinterpret_cast<float4*>(sB + umma::sw128_offset_f32(r, c)) = *reinterpret_cast<const float4*>(B + (size_t)r * 32 + c * 4);
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t idesc = umma::make_idesc_tf32(128, N);
  if (tid == 0) {
    uint64_t dA = umma::make_smem_desc_sw128(umma::smem_u32(sP) + (uint32_t)start_row * 128u);
    dA &= ~((uint64_t)0x3FFF << 32);
    dA |= (uint64_t)(((uint32_t)sbo_rows * 128u >> 4) & 0x3FFF) << 32;
    if (base_off_mode == 1) dA |= (uint64_t)(start_row & 7) << 49;
    const uint64_t dB = umma::make_smem_desc_sw128(umma::smem_u32(sB));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t adv = (uint64_t)(k * 32 >> 4);
      umma::mma_tf32_ss(tmem, dA + adv, dB + adv, idesc, k > 0 ? 1u : 0u);
    }
    umma::mma_commit(&mbar);
  }
  const bool ok = umma::mbar_wait(&mbar, 0);
  __syncthreads();
  umma::fence_after_thread_sync();
  if (ok) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = v[j];
    }
  }
  if (tid == 0) *status = ok ? 0 : 1;
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

// MN-major test: both operands are windows into "strips" of 128-byte rows (32 fp32 per row, stored with the 32-byte-base
// swizzle keyed on the absolute row index, umma::sw128b32_offset_f32), the contraction runs over ROWS:
//   D[32 j + c][32 q + n] = sum_{g < 2 ksteps} sum_{t < 4} PA[a_row0 + j a_lbo_rows + g a_sbo_rows + t][c]
//                                                        * PB[b_row0 + q b_lbo_rows + g b_sbo_rows + t][n]
// for the 4 M blocks j and the N / 32 N blocks q (consecutive K steps advance both windows by 2 * sbo rows).
// a_lbo_rows = 1, a_sbo_rows = 4 is the weight-gradient case: M block j is the same strip shifted by j rows (the kernel
// column kw of a 3x3 tap), K runs over 8 consecutive rows.
__global__ void __launch_bounds__(128) umma_mn_kernel(const float* __restrict__ PA, const float* __restrict__ PB,
                                                      float* __restrict__ D, int rows_a, int rows_b, int a_row0,
                                                      int a_lbo_rows, int a_sbo_rows, int b_row0, int b_lbo_rows, int b_sbo_rows,
                                                      int ksteps, int N, int layout_type, int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* sA = reinterpret_cast<float*>(smem_raw);
  float* sB = sA + (size_t)((rows_a + 7) / 8 * 8) * 32;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) umma::tmem_alloc(&tmem_base_slot, ncols);
  if (tid == 0) {
    umma::mbar_init(&mbar, 1);
    umma::fence_mbar_init();
  }
  for (int idx = tid; idx < rows_a * 8; idx += 128) {
    const int r = idx >> 3, c = idx & 7;
    *reinterpret_cast<float4*>(sA + umma::sw128b32_offset_f32(r, c)) = *reinterpret_cast<const float4*>(PA + (size_t)r * 32 + c * 4);
  }
  for (int idx = tid; idx < rows_b * 8; idx += 128) {
    const int r = idx >> 3, c = idx & 7;
    *reinterpret_cast<float4*>(sB + umma::sw128b32_offset_f32(r, c)) = *reinterpret_cast<const float4*>(PB + (size_t)r * 32 + c * 4);
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t idesc = umma::make_idesc_tf32_major(128, N, 1, 1);
  if (tid == 0) {
    const uint64_t dA = umma::make_smem_desc_mn_b32(umma::smem_u32(sA) + (uint32_t)a_row0 * 128u, (uint32_t)a_lbo_rows * 128u,
                                                    (uint32_t)a_sbo_rows * 128u);
    const uint64_t dB = umma::make_smem_desc_mn_b32(umma::smem_u32(sB) + (uint32_t)b_row0 * 128u, (uint32_t)b_lbo_rows * 128u,
                                                    (uint32_t)b_sbo_rows * 128u);
    // layout_type 2 (the 16-byte-base SWIZZLE_128B of the K-major tiles) is kept reachable to document the measured
    // behaviour: the instruction completes and D is all zeros
    const uint64_t lt = (layout_type == 2) ? (((uint64_t)1 << 61) ^ ((uint64_t)2 << 61)) : 0;
    for (int k = 0; k < ksteps; ++k) {
      const uint64_t adv_a = (uint64_t)(k * 2 * a_sbo_rows * 128 >> 4), adv_b = (uint64_t)(k * 2 * b_sbo_rows * 128 >> 4);
      umma::mma_tf32_ss(tmem, (dA + adv_a) ^ lt, (dB + adv_b) ^ lt, idesc, k > 0 ? 1u : 0u);
    }
    umma::mma_commit(&mbar);
  }
  const bool ok = umma::mbar_wait(&mbar, 0);
  __syncthreads();
  umma::fence_after_thread_sync();
  if (ok) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = v[j];
    }
  }
  if (tid == 0) *status = ok ? 0 : 1;
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

}  // namespace
}  // namespace b200ocl

extern "C" int b200ocl_selftest_umma_mn(const float* PA, const float* PB, float* D, int rows_a, int rows_b, int a_row0,
                                        int a_lbo_rows, int a_sbo_rows, int b_row0, int b_lbo_rows, int b_sbo_rows, int ksteps,
                                        int N, int layout_type, int* status, void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200OCL_CHECK_ARG(PA && PB && D && status, "null pointer");
  B200OCL_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0 && ksteps >= 1 && ksteps <= 64 && (layout_type == 1 || layout_type == 2),
                    "need N in [16,256] %16, 1 <= ksteps <= 64, layout_type 1 or 2");
  B200OCL_CHECK_ARG(rows_a > 0 && rows_b > 0 && rows_a + rows_b <= 1600 && a_row0 >= 0 && b_row0 >= 0 && a_lbo_rows >= 1 &&
                        b_lbo_rows >= 1 && a_sbo_rows >= 1 && b_sbo_rows >= 1 &&
                        a_row0 + 3 * a_lbo_rows + (2 * ksteps - 1) * a_sbo_rows + 4 <= rows_a &&
                        b_row0 + ((N + 31) / 32 - 1) * b_lbo_rows + (2 * ksteps - 1) * b_sbo_rows + 4 <= rows_b,
                    "window exceeds the strips");
  const size_t smem = (size_t)((rows_a + 7) / 8 * 8 + (rows_b + 7) / 8 * 8) * 32 * sizeof(float) + 1024;
  B200OCL_CUDA(cudaFuncSetAttribute(umma_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_mn_kernel<<<1, 128, smem, stream>>>(PA, PB, D, rows_a, rows_b, a_row0, a_lbo_rows, a_sbo_rows, b_row0, b_lbo_rows, b_sbo_rows,
                                           ksteps, N, layout_type, status);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

extern "C" int b200ocl_selftest_umma_tf32(const float* A, const float* B, float* D, int N, int K, int mode, int* status,
                                          void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200OCL_CHECK_ARG(A && B && D && status, "null pointer");
  B200OCL_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0 && K >= 32 && K % 32 == 0, "need N in [16,256] %16, K %32");
  const size_t smem = (size_t)(2 * 128 * 32 + 2 * 256 * 32) * sizeof(float) + 1024;
  static bool configured_dev[B200OCL_MAX_DEVICES] = {};
  bool& configured = configured_dev[b200ocl::device_slot()];
  if (!configured) {
    B200OCL_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  umma_selftest_kernel<<<1, 128, smem, stream>>>(A, B, D, N, K, mode, status);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

extern "C" int b200ocl_selftest_umma_window(const float* P, const float* B, float* D, int rows, int start_row,
                                            int sbo_rows, int base_off_mode, int N, int* status, void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200OCL_CHECK_ARG(P && B && D && status, "null pointer");
  B200OCL_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0, "need N in [16,256] %16");
  B200OCL_CHECK_ARG(rows > 0 && rows <= 1024 && start_row >= 0 && sbo_rows > 0 &&
                        start_row + 15 * sbo_rows + 8 <= rows,
                    "window exceeds the patch");
  const size_t smem = (size_t)((rows + 7) / 8 * 8 + 256) * 32 * sizeof(float) + 1024;
  B200OCL_CUDA(cudaFuncSetAttribute(umma_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_window_kernel<<<1, 128, smem, stream>>>(P, B, D, rows, start_row, sbo_rows, base_off_mode, N, status);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}
