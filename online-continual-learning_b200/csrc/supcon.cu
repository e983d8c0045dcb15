// supcon.cu -- fused supervised-contrastive loss, forward + backward (sm_100a).
//
// Replaces SupConLoss.forward and its autograd backward (utils/loss.py:19-96; ~40 ATen
// kernels on [A,A] temporaries, A = V*B anchors) by two launches that never materialise the
// [A,A] logit matrix:
//   stats kernel  one warp per anchor i streams all contrast rows through shared memory,
//                 lane j of a 32-row tile owns logit l_ij = c_i.c_j / T; online max /
//                 exp-sum (diagonal in the max, out of the sum: loss.py:71-86), positive
//                 logit sum and count; writes lse_i = max_i + log Z_i, |P(i)|, and the loss
//                 through a fixed-order two-level reduction (last CTA finishes);
//   grad kernel   recomputes l_ij tile by tile and accumulates
//                 dc_i = (1/T) sum_j (G_ij + G_ji) c_j,  G_ij = (exp(l_ij - lse_i) - 1[j in P(i)]/|P(i)|)/A,
//                 using l_ji = l_ij (anchor set == contrast set in 'all' mode, loss.py:60-62).
// Anchors are in view-major order a = v*B + b (loss.py:56); features/grad are [B,V,d].
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace b200ocl {
namespace {

constexpr int SC_THREADS = 256;
constexpr int SC_WARPS = 8;

struct SupconParams {
  const float* feats;
  const long long* labels;
  int B, V, d, A, pitch;
  float T;
  float* lse;             // [A]
  float* npos;            // [A]
  float* part;            // [gridDim] loss partials
  unsigned int* counter;
  float* loss;
  float* dfeats;
};

__device__ __forceinline__ const float* anchor_row(const SupconParams& p, int a) {
  const int v = a / p.B, b = a - v * p.B;
  return p.feats + ((size_t)b * p.V + v) * p.d;
}

// Cooperative load of contrast rows [j0, j0+32) into sct (pitch floats per row, odd pitch).
__device__ __forceinline__ void load_tile(const SupconParams& p, int j0, float* sct, int tid) {
  for (int idx = tid; idx < 32 * p.d; idx += SC_THREADS) {
    const int r = idx / p.d, dd = idx - r * p.d;
    const int j = j0 + r;
    sct[r * p.pitch + dd] = (j < p.A) ? anchor_row(p, j)[dd] : 0.f;
  }
}

__device__ __forceinline__ float tile_dot(const float* sa_row, const float* sct_row, int d) {
  float acc0 = 0.f, acc1 = 0.f;
  int dd = 0;
  for (; dd + 1 < d; dd += 2) {
    acc0 = fmaf(sa_row[dd], sct_row[dd], acc0);
    acc1 = fmaf(sa_row[dd + 1], sct_row[dd + 1], acc1);
  }
  if (dd < d) acc0 = fmaf(sa_row[dd], sct_row[dd], acc0);
  return acc0 + acc1;
}

__global__ void __launch_bounds__(SC_THREADS) supcon_stats_kernel(SupconParams p) {
  extern __shared__ __align__(16) float smem[];
  float* sa = smem;                          // [8][d]
  float* sct = sa + SC_WARPS * p.d;          // [32][pitch]
  __shared__ float s_loss[SC_WARPS];
  __shared__ bool is_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a0 = blockIdx.x * SC_WARPS;
  const int i = a0 + warp;
  const bool valid = i < p.A;

  for (int idx = tid; idx < SC_WARPS * p.d; idx += SC_THREADS) {
    const int w = idx / p.d, dd = idx - w * p.d;
    sa[idx] = (a0 + w < p.A) ? anchor_row(p, a0 + w)[dd] : 0.f;
  }
  const long long yi = valid ? p.labels[i % p.B] : 0;

  float mx = -FLT_MAX, z = 0.f, ps = 0.f, np = 0.f;
  for (int j0 = 0; j0 < p.A; j0 += 32) {
    __syncthreads();
    load_tile(p, j0, sct, tid);
    __syncthreads();
    const int j = j0 + lane;
    if (valid && j < p.A) {
      const float l = __fdiv_rn(tile_dot(sa + warp * p.d, sct + lane * p.pitch, p.d), p.T);
      if (l > mx) {                  // rescale the running sum to the new max
        z *= expf(mx - l);
        mx = l;
      }
      if (j != i) {
        z += expf(l - mx);
        if (p.labels[j % p.B] == yi) {
          ps += l;
          np += 1.f;
        }
      }
    }
  }
  // combine lanes
  const float M = warp_max(mx);
  const float Z = warp_sum(z * expf(mx - M));   // lanes that saw nothing: z == 0
  const float PS = warp_sum(ps);
  const float NP = warp_sum(np);
  float loss_i = 0.f;
  if (valid) {
    const float lse = M + logf(Z);
    loss_i = -(PS - NP * lse) / NP;             // 0/0 -> NaN like loss.py:90
    if (lane == 0) {
      p.lse[i] = lse;
      p.npos[i] = NP;
    }
  }
  if (lane == 0) s_loss[warp] = loss_i;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < SC_WARPS; ++w) t += s_loss[w];
    p.part[blockIdx.x] = t;
    __threadfence();
    is_last = (atomicAdd(p.counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && tid == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) t += (double)__ldcg(p.part + b);
    *p.loss = (float)(t / (double)p.A);
  }
}

template <int DCH>
__global__ void __launch_bounds__(SC_THREADS) supcon_grad_kernel(SupconParams p) {
  extern __shared__ __align__(16) float smem[];
  float* sa = smem;                          // [8][d]
  float* sct = sa + SC_WARPS * p.d;          // [32][pitch]
  float* wbuf = sct + 32 * p.pitch;          // [8][32]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a0 = blockIdx.x * SC_WARPS;
  const int i = a0 + warp;
  const bool valid = i < p.A;

  for (int idx = tid; idx < SC_WARPS * p.d; idx += SC_THREADS) {
    const int w = idx / p.d, dd = idx - w * p.d;
    sa[idx] = (a0 + w < p.A) ? anchor_row(p, a0 + w)[dd] : 0.f;
  }
  const long long yi = valid ? p.labels[i % p.B] : 0;
  const float lse_i = valid ? p.lse[i] : 0.f;
  const float np_i = valid ? p.npos[i] : 1.f;
  const float invA = 1.f / (float)p.A;

  float acc[DCH];
#pragma unroll
  for (int c = 0; c < DCH; ++c) acc[c] = 0.f;

  for (int j0 = 0; j0 < p.A; j0 += 32) {
    __syncthreads();
    load_tile(p, j0, sct, tid);
    __syncthreads();
    const int j = j0 + lane;
    float w = 0.f;
    if (valid && j < p.A && j != i) {
      const float l = __fdiv_rn(tile_dot(sa + warp * p.d, sct + lane * p.pitch, p.d), p.T);
      const float pos = (p.labels[j % p.B] == yi) ? 1.f : 0.f;
      const float g_ij = expf(l - lse_i) - pos / np_i;
      const float g_ji = expf(l - p.lse[j]) - pos / p.npos[j];
      w = (g_ij + g_ji) * invA / p.T;
    }
    wbuf[warp * 32 + lane] = w;
    __syncwarp();
    const int jn = min(32, p.A - j0);
    for (int jj = 0; jj < jn; ++jj) {
      const float wj = wbuf[warp * 32 + jj];
      const float* row = sct + jj * p.pitch;
#pragma unroll
      for (int c = 0; c < DCH; ++c) {
        const int dd = lane + 32 * c;
        if (dd < p.d) acc[c] = fmaf(wj, row[dd], acc[c]);
      }
    }
    __syncwarp();
  }
  if (valid) {
    const int v = i / p.B, b = i - v * p.B;
    float* out = p.dfeats + ((size_t)b * p.V + v) * p.d;
#pragma unroll
    for (int c = 0; c < DCH; ++c) {
      const int dd = lane + 32 * c;
      if (dd < p.d) out[dd] = acc[c];
    }
  }
}

}  // namespace
}  // namespace b200ocl

extern "C" {

size_t b200ocl_supcon_workspace_bytes(int B, int V, int d) {
  (void)d;
  const size_t A = (size_t)(B > 0 ? B : 0) * (size_t)(V > 0 ? V : 0);
  const size_t blocks = (A + b200ocl::SC_WARPS - 1) / b200ocl::SC_WARPS;
  return 256 + b200ocl::align_up((2 * A + blocks) * sizeof(float), 256);
}

int b200ocl_supcon(const float* feats, const int64_t* labels, int B, int V, int d, float temperature, float* loss,
                   float* dfeats, void* workspace, size_t workspace_bytes, void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200OCL_CHECK_ARG(feats && labels && loss, "null pointer");
  B200OCL_CHECK_ARG(B >= 1 && V >= 1 && d >= 1, "need B,V,d >= 1");
  B200OCL_CHECK_ARG(temperature > 0.f, "temperature must be positive");
  if (d > 1024) {
    set_error("b200ocl_supcon: d=%d exceeds the kernel's limit of 1024", d);
    return B200OCL_EUNSUPPORTED;
  }
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255) ||
      workspace_bytes < b200ocl_supcon_workspace_bytes(B, V, d)) {
    set_error("b200ocl_supcon: workspace missing, misaligned or smaller than %zu bytes",
              b200ocl_supcon_workspace_bytes(B, V, d));
    return B200OCL_EWORKSPACE;
  }
  SupconParams p{};
  p.feats = feats;
  p.labels = reinterpret_cast<const long long*>(labels);
  p.B = B; p.V = V; p.d = d; p.A = B * V;
  p.pitch = (d % 2 == 0) ? d + 1 : d;
  p.T = temperature;
  p.counter = static_cast<unsigned int*>(workspace);
  float* ws = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + 256);
  p.lse = ws;
  p.npos = ws + p.A;
  p.part = ws + 2 * (size_t)p.A;
  p.loss = loss;
  p.dfeats = dfeats;
  const int grid = (p.A + SC_WARPS - 1) / SC_WARPS;
  const size_t smem_stats = (size_t)(SC_WARPS * d + 32 * p.pitch) * sizeof(float);
  const size_t smem_grad = smem_stats + SC_WARPS * 32 * sizeof(float);

  B200OCL_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), stream));
  static bool configured_dev[B200OCL_MAX_DEVICES] = {};
  bool& configured = configured_dev[b200ocl::device_slot()];
  if (!configured) {
    const int max_smem = 200 * 1024;  // d = 1024 needs 165 KB; static smem takes a little of the 227 KB
    B200OCL_CUDA(cudaFuncSetAttribute(supcon_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    B200OCL_CUDA(cudaFuncSetAttribute(supcon_grad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    B200OCL_CUDA(cudaFuncSetAttribute(supcon_grad_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    B200OCL_CUDA(cudaFuncSetAttribute(supcon_grad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    B200OCL_CUDA(cudaFuncSetAttribute(supcon_grad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    configured = true;
  }
  B200OCL_PROF("supcon", 4.0 * p.A * d + 8.0 * B + 8.0 * p.A, stream);
  supcon_stats_kernel<<<grid, SC_THREADS, smem_stats, stream>>>(p);
  B200OCL_LAUNCHED();
  if (!dfeats) return B200OCL_OK;
#define B200OCL_SC_GRAD(DCH) supcon_grad_kernel<DCH><<<grid, SC_THREADS, smem_grad, stream>>>(p)
  B200OCL_PROF("supcon", 8.0 * p.A * d, stream);
  if (d <= 128) B200OCL_SC_GRAD(4);
  else if (d <= 256) B200OCL_SC_GRAD(8);
  else if (d <= 512) B200OCL_SC_GRAD(16);
  else B200OCL_SC_GRAD(32);
#undef B200OCL_SC_GRAD
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

}  // extern "C"
