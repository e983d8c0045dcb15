// conv_tc.cu -- 3x3 stride-1 convolution on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Implicit GEMM  D[128 pixels, NT channels] += A[128, 32] * B[NT, 32]^T  per K block of 32 (tap, cin)
// values, kind::tf32 with the accumulator in tensor memory.  fp32 accuracy is kept with the 3xTF32
// split (A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, both halves exactly representable in TF32) AND by
// promoting every K block: the tensor core's own accumulation truncates (measured: the error of a
// long TMEM accumulation grows ~6e-9 * K, 9e-6 at K = 1440), so each K block accumulates its 12 MMAs
// into a fresh TMEM buffer that the threads then add into fp32 registers with round-to-nearest --
// two TMEM buffers ping-pong so that the MMAs of block kb overlap the read-out of block kb-1 and the
// operand staging of block kb+1.
//
//   A  activations, NHWC fp32 in HBM; thread t owns pixel t of the tile, gathers its 8 x 16-byte
//      chunks of the K block (zero outside the image / beyond K), splits them and writes the two
//      128-byte-swizzled K-major tiles the MMA reads (no im2col in HBM);
//   B  weights, pre-split and pre-swizzled tile images written by tc_pack_kernel (net_fwd.cu):
//      a byte copy with cp.async;
//   D  TMEM, 2 x NT fp32 columns; read with tcgen05.ld.32x32b (lane = pixel row = thread).
// Epilogues as in conv.cu: folded eval BN (+ReLU, +residual), raw store + deterministic fp64 batch
// statistics for training, raw / accumulate for the flipped stride-1 data gradient.
#include <stdlib.h>

#include "conv.cuh"
#include "umma.cuh"

namespace b200ocl {
namespace {

constexpr int TC_THREADS = 128;
constexpr int A_TILE = 128 * 32;  // floats

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned int s = static_cast<unsigned int>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}

template <int NT>
__device__ __forceinline__ void tmem_accumulate(uint32_t taddr, float (&acc)[NT]) {
#pragma unroll
  for (int c0 = 0; c0 < NT; c0 += 16) {
    float v[16];
    umma::tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[c0 + j] += v[j];
  }
}

template <int NT>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(ConvArgs a) {
  constexpr int B_TILE = NT * 32;
  constexpr int STAGE_F = 2 * A_TILE + 2 * B_TILE;
  constexpr uint32_t TMEM_COLS = (2 * NT <= 64) ? 64 : ((2 * NT <= 128) ? 128 : 256);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* stage0 = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t mma_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ bool is_last;
  __shared__ int s_fail;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int bn = a.tc_bn;
  const int n0 = blockIdx.y * bn;
  const int KB = a.tc_kb;
  const int ktot = 9 * a.CK;
  const int m = blockIdx.x * 128 + tid;
  const bool valid = m < a.M;
  const int hw = a.Hout * a.Wout;
  const int img = valid ? m / hw : 0;
  const int rem = valid ? m - img * hw : 0;
  const int py = rem / a.Wout, px = rem - py * a.Wout;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) {
    umma::mbar_init(&mma_bar[0], 1);
    umma::mbar_init(&mma_bar[1], 1);
    umma::fence_mbar_init();
    s_fail = 0;
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t my_lanes = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc = umma::make_idesc_tf32(128, NT);
  const float* wimg = a.w_tc + (size_t)blockIdx.y * KB * 2 * B_TILE;

  float acc[NT];
#pragma unroll
  for (int c = 0; c < NT; ++c) acc[c] = 0.f;

  // running (tap, channel) position of this thread's next 16-byte chunk and the tap's source pointer
  int tap = 0, ci = 0;
  const float* tap_src = nullptr;
  auto set_tap = [&]() {
    const int kh = tap / 3, kw = tap - kh * 3;
    const int iy = py + kh - 1, ix = px + kw - 1;
    const bool ok = valid && tap < 9 && (unsigned)iy < (unsigned)a.Hin && (unsigned)ix < (unsigned)a.Win;
    tap_src = ok ? a.in + ((size_t)(img * a.Hin + iy) * a.Win + ix) * a.CK : nullptr;
  };
  set_tap();
  bool ok_all = true;

  for (int kb = 0; kb < KB; ++kb) {
    const int s = kb & 1;
    float* sAh = stage0 + s * STAGE_F;
    float* sAl = sAh + A_TILE;
    float* sB = sAl + A_TILE;  // [hi | lo], contiguous like the image
    // ---- B: byte copy of the pre-swizzled hi/lo tiles of this K block
    const float* bsrc = wimg + (size_t)kb * 2 * B_TILE;
#pragma unroll
    for (int i = 0; i < (2 * B_TILE / 4) / TC_THREADS; ++i)
      cp_async16(sB + (tid + i * TC_THREADS) * 4, bsrc + (tid + i * TC_THREADS) * 4);
    asm volatile("cp.async.commit_group;\n" ::);
    // ---- A: gather this pixel's 8 chunks (all loads first, then split + store)
    float4 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tap_src != nullptr) v[c] = __ldg(reinterpret_cast<const float4*>(tap_src + ci));
      ci += 4;
      if (ci == a.CK) {
        ci = 0;
        ++tap;
        set_tap();
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 h, l;
      umma::split_tf32(v[c].x, h.x, l.x); umma::split_tf32(v[c].y, h.y, l.y);
      umma::split_tf32(v[c].z, h.z, l.z); umma::split_tf32(v[c].w, h.w, l.w);
      const int off = umma::sw128_offset_f32(tid, c);
      *reinterpret_cast<float4*>(sAh + off) = h;
      *reinterpret_cast<float4*>(sAl + off) = l;
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_thread_sync();
      const uint64_t dAh = umma::make_smem_desc_sw128(umma::smem_u32(sAh));
      const uint64_t dAl = umma::make_smem_desc_sw128(umma::smem_u32(sAl));
      const uint64_t dBh = umma::make_smem_desc_sw128(umma::smem_u32(sB));
      const uint64_t dBl = umma::make_smem_desc_sw128(umma::smem_u32(sB + B_TILE));
      const uint32_t dcol = tmem + (uint32_t)(s * NT);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)(k * 2);  // 32 bytes per K step, in 16-byte units
        umma::mma_tf32_ss(dcol, dAh + adv, dBh + adv, idesc, k > 0 ? 1u : 0u);
        umma::mma_tf32_ss(dcol, dAh + adv, dBl + adv, idesc, 1u);
        umma::mma_tf32_ss(dcol, dAl + adv, dBh + adv, idesc, 1u);
      }
      umma::mma_commit(&mma_bar[s]);
    }
    // ---- promote the previous K block while this one runs on the tensor core
    if (kb > 0) {
      if (!umma::mbar_wait(&mma_bar[s ^ 1], (uint32_t)(((kb - 1) >> 1) & 1))) ok_all = false;
      umma::fence_after_thread_sync();
      tmem_accumulate<NT>(my_lanes + (uint32_t)((s ^ 1) * NT), acc);
    }
  }
  {
    const int sl = (KB - 1) & 1;
    if (!umma::mbar_wait(&mma_bar[sl], (uint32_t)(((KB - 1) >> 1) & 1))) ok_all = false;
    umma::fence_after_thread_sync();
    tmem_accumulate<NT>(my_lanes + (uint32_t)(sl * NT), acc);
  }
  if (!ok_all) s_fail = 1;

  // ------------------------------------------------------------------ epilogue (thread = pixel row)
  if (a.mode == CONV_EVAL) {
    if (valid) {
      float* o = a.out + (size_t)m * a.CN + n0;
      const float* rs = a.residual ? a.residual + (size_t)m * a.CN + n0 : nullptr;
#pragma unroll
      for (int c0 = 0; c0 < NT; c0 += 4) {   // compile-time indices into acc[]; tiles are 20/40/80 wide
        if (c0 >= bn) break;
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = n0 + c0 + j;
          const float inv = 1.0f / sqrtf(a.rvar[c] + a.eps);
          r[j] = (acc[c0 + j] - a.rmean[c]) * (inv * a.gamma[c]) + a.beta[c];
        }
        if (rs) {
          const float4 r4 = *reinterpret_cast<const float4*>(rs + c0);
          r[0] += r4.x; r[1] += r4.y; r[2] += r4.z; r[3] += r4.w;
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 4; ++j) r[j] = fmaxf(r[j], 0.f);
        }
        *reinterpret_cast<float4*>(o + c0) = make_float4(r[0], r[1], r[2], r[3]);
      }
    }
  } else {
    if (valid) {
      float* o = a.out + (size_t)m * a.CN + n0;
#pragma unroll
      for (int c0 = 0; c0 < NT; c0 += 4) {
        if (c0 >= bn) break;
        float4 r = make_float4(acc[c0], acc[c0 + 1], acc[c0 + 2], acc[c0 + 3]);
        if (a.mode == CONV_ACCUM) {
          const float4 old = *reinterpret_cast<const float4*>(o + c0);
          r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
        }
        *reinterpret_cast<float4*>(o + c0) = r;
      }
    }
    if (a.mode == CONV_TRAIN) {
      // batch statistics: transpose through shared memory, one thread per channel sums its 128 rows in
      // fp64 in row order (rows beyond M hold exact zeros), then partial per CTA + last-CTA finalize
      __syncthreads();  // staging buffers are free
      float* s_t = stage0;  // [128][bn + 1]
#pragma unroll
      for (int c = 0; c < NT; ++c)
        if (c < bn) s_t[tid * (bn + 1) + c] = acc[c];
      __syncthreads();
      if (tid < bn) {
        double S = 0.0, Q = 0.0;
        for (int r = 0; r < 128; ++r) {
          const double x = (double)s_t[r * (bn + 1) + tid];
          S += x;
          Q += x * x;
        }
        double* dst = a.stat_part + ((size_t)blockIdx.x * a.CN + n0 + tid) * 2;
        dst[0] = S;
        dst[1] = Q;
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) is_last = (atomicAdd(a.counter + blockIdx.y, 1u) == gridDim.x - 1);
      __syncthreads();
      if (is_last) {
        __threadfence();
        const int groups = TC_THREADS / bn;          // 6, 3 or 1 partial-subsets per channel
        const int ch = tid % bn, grp = tid / bn;
        double* s_fin = reinterpret_cast<double*>(stage0);   // [groups][bn][2]
        __syncthreads();
        if (grp < groups) {
          double s4[4] = {0.0, 0.0, 0.0, 0.0}, q4[4] = {0.0, 0.0, 0.0, 0.0};
          unsigned int b = grp;
          for (; b + 3 * groups < gridDim.x; b += 4 * groups) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const double2 pv = __ldcg(reinterpret_cast<const double2*>(a.stat_part + ((size_t)(b + u * groups) * a.CN + n0 + ch) * 2));
              s4[u] += pv.x;
              q4[u] += pv.y;
            }
          }
          for (; b < gridDim.x; b += groups) {
            const double2 pv = __ldcg(reinterpret_cast<const double2*>(a.stat_part + ((size_t)b * a.CN + n0 + ch) * 2));
            s4[0] += pv.x;
            q4[0] += pv.y;
          }
          s_fin[(grp * bn + ch) * 2 + 0] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
          s_fin[(grp * bn + ch) * 2 + 1] = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        }
        __syncthreads();
        if (tid < bn) {
          double S = 0.0, Q = 0.0;
          for (int g = 0; g < groups; ++g) {
            S += s_fin[(g * bn + tid) * 2 + 0];
            Q += s_fin[(g * bn + tid) * 2 + 1];
          }
          const double cnt = (double)a.M;
          const double mean = S / cnt;
          double var = Q / cnt - mean * mean;
          if (var < 0.0) var = 0.0;
          const int c = n0 + tid;
          a.save_mean[c] = (float)mean;
          a.save_invstd[c] = (float)(1.0 / sqrt(var + (double)a.eps));
          const double unbiased = (a.M > 1) ? var * cnt / (cnt - 1.0) : var;
          a.run_mean[c] = (1.f - a.momentum) * a.run_mean[c] + a.momentum * (float)mean;
          a.run_var[c] = (1.f - a.momentum) * a.run_var[c] + a.momentum * (float)unbiased;
        }
      }
    }
  }
  // a timed-out MMA barrier (must never happen) poisons the output instead of hanging the GPU
  __syncthreads();
  if (s_fail && valid) a.out[(size_t)m * a.CN + n0] = __int_as_float(0x7fc00000);
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, TMEM_COLS);
}


// ------------------------------------------------------------------------------------------------
// Warp-specialised version (default).  The v1 kernel above runs gather -> MMA -> promote in lockstep
// and measured ~4400 cycles per K block with the tensor pipe ~10 % busy; here the three roles run
// concurrently through mbarrier rings:
//   warps 0-7   producers  two threads per pixel row (4 of its 8 chunks each): a lone warp per scheduler
//                          runs the convert/store chain at IPC ~0.3, so the staging work is spread over
//                          8 warps; the next block's LDG.128s are in flight while the current block is
//                          split (cvt.rna.tf32) and stored; B image copied alongside; arrive full[s]
//   warp  12    MMA issue  one thread: wait full[s] and tmem_empty[t], 12 x tcgen05.mma, commit ->
//                          empty[s] (smem stage reusable) and tmem_full[t]
//   warps 8-11  promote    wait tmem_full[t], tcgen05.ld their 32 lanes, add into fp32 registers,
//                          arrive tmem_empty[t]; then run the epilogue
template <int NT>
__global__ void __launch_bounds__(416) conv_tc_ws_kernel(ConvArgs a) {
  constexpr int B_TILE = NT * 32;
  constexpr int STAGE_F = 2 * A_TILE + 2 * B_TILE;
  constexpr int S = 3;
  constexpr uint32_t TMEM_COLS = (2 * NT <= 64) ? 64 : ((2 * NT <= 128) ? 128 : 256);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* stage0 = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t full_bar[S], empty_bar[S], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ bool is_last;
  __shared__ int s_fail;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int bn = a.tc_bn;
  const int n0 = blockIdx.y * bn;
  const int KB = a.tc_kb;

  if (warp == 12) umma::tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) {
    for (int i = 0; i < S; ++i) {
      umma::mbar_init(&full_bar[i], 256);
      umma::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(&tfull_bar[i], 1);
      umma::mbar_init(&tempty_bar[i], 128);
    }
    umma::fence_mbar_init();
    s_fail = 0;
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    // =========================================================== producers
    const int row = tid & 127, half = tid >> 7;      // this thread stages chunks [4*half, 4*half+4) of its row
    const int m = blockIdx.x * 128 + row;
    const bool valid = m < a.M;
    const int hw = a.Hout * a.Wout;
    const int img = valid ? m / hw : 0;
    const int rem = valid ? m - img * hw : 0;
    const int py = rem / a.Wout, px = rem - py * a.Wout;
    const float* wimg = a.w_tc + (size_t)blockIdx.y * KB * 2 * B_TILE;
    int tap = 0, ci = 0;
    const float* tap_src = nullptr;
    auto set_tap = [&]() {
      const int kh = tap / 3, kw = tap - kh * 3;
      const int iy = py + kh - 1, ix = px + kw - 1;
      const bool ok = valid && tap < 9 && (unsigned)iy < (unsigned)a.Hin && (unsigned)ix < (unsigned)a.Win;
      tap_src = ok ? a.in + ((size_t)(img * a.Hin + iy) * a.Win + ix) * a.CK : nullptr;
    };
    auto advance = [&]() {   // one 16-byte chunk further along (tap, channel)
      ci += 4;
      if (ci == a.CK) {
        ci = 0;
        ++tap;
        set_tap();
      }
    };
    auto load_block = [&](float4 (&v)[4]) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tap_src != nullptr) v[c] = __ldg(reinterpret_cast<const float4*>(tap_src + ci));
        advance();
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) advance();   // the partner thread's chunks
    };
    set_tap();
    if (half) {
#pragma unroll
      for (int c = 0; c < 4; ++c) advance();
    }
    float4 va[4], vb[4];
    load_block(va);
    constexpr int BW = (2 * B_TILE / 4) / 256;
    // weights of block kb are also fetched one iteration early, so no load is consumed in the
    // iteration that issued it
    float4 bwa[BW], bwb[BW];
    {
      const float4* bsrc0 = reinterpret_cast<const float4*>(wimg);
#pragma unroll
      for (int i = 0; i < BW; ++i) bwa[i] = __ldg(bsrc0 + tid + i * 256);
    }
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % S;
      if (kb + 1 < KB) {   // prefetch the next block's activations and weights before touching the current one
        const float4* bsrc = reinterpret_cast<const float4*>(wimg + (size_t)(kb + 1) * 2 * B_TILE);
        if (kb & 1) {
          load_block(va);
#pragma unroll
          for (int i = 0; i < BW; ++i) bwa[i] = __ldg(bsrc + tid + i * 256);
        } else {
          load_block(vb);
#pragma unroll
          for (int i = 0; i < BW; ++i) bwb[i] = __ldg(bsrc + tid + i * 256);
        }
      }
      if (!umma::mbar_wait(&empty_bar[s], (uint32_t)(((kb / S) & 1) ^ 1))) s_fail = 1;
      float* sAh = stage0 + s * STAGE_F;
      float* sAl = sAh + A_TILE;
      float* sB = sAl + A_TILE;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 x = (kb & 1) ? vb[c] : va[c];
        float4 h, l;
        umma::split_tf32(x.x, h.x, l.x); umma::split_tf32(x.y, h.y, l.y);
        umma::split_tf32(x.z, h.z, l.z); umma::split_tf32(x.w, h.w, l.w);
        const int off = umma::sw128_offset_f32(row, half * 4 + c);
        *reinterpret_cast<float4*>(sAh + off) = h;
        *reinterpret_cast<float4*>(sAl + off) = l;
      }
#pragma unroll
      for (int i = 0; i < BW; ++i) reinterpret_cast<float4*>(sB)[tid + i * 256] = (kb & 1) ? bwb[i] : bwa[i];
      umma::fence_proxy_async_smem();
      umma::mbar_arrive(&full_bar[s]);
    }
  } else if (warp == 12) {
    // =========================================================== MMA issuer
    if ((tid & 31) == 0) {
      const uint32_t idesc = umma::make_idesc_tf32(128, NT);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % S, t = kb & 1;
        if (!umma::mbar_wait(&full_bar[s], (uint32_t)((kb / S) & 1))) s_fail = 1;
        if (!umma::mbar_wait(&tempty_bar[t], (uint32_t)(((kb >> 1) & 1) ^ 1))) s_fail = 1;
        umma::fence_after_thread_sync();
        float* sAh = stage0 + s * STAGE_F;
        const uint64_t dAh = umma::make_smem_desc_sw128(umma::smem_u32(sAh));
        const uint64_t dAl = umma::make_smem_desc_sw128(umma::smem_u32(sAh + A_TILE));
        const uint64_t dBh = umma::make_smem_desc_sw128(umma::smem_u32(sAh + 2 * A_TILE));
        const uint64_t dBl = umma::make_smem_desc_sw128(umma::smem_u32(sAh + 2 * A_TILE + B_TILE));
        const uint32_t dcol = tmem + (uint32_t)(t * NT);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adv = (uint64_t)(k * 2);
          umma::mma_tf32_ss(dcol, dAh + adv, dBh + adv, idesc, k > 0 ? 1u : 0u);
          umma::mma_tf32_ss(dcol, dAh + adv, dBl + adv, idesc, 1u);
          umma::mma_tf32_ss(dcol, dAl + adv, dBh + adv, idesc, 1u);
        }
        umma::mma_commit(&empty_bar[s]);
        umma::mma_commit(&tfull_bar[t]);
      }
    }
  } else {
    // =========================================================== promotion + epilogue (warps 8-11)
    const int et = tid - 256;                 // 0..127 = pixel row = TMEM lane
    const int ew = warp - 8;                  // == warp % 4: the TMEM lane quarter this warp may read
    const uint32_t my_lanes = tmem + ((uint32_t)(ew * 32) << 16);
    const int m = blockIdx.x * 128 + et;
    const bool valid = m < a.M;
    float acc[NT];
#pragma unroll
    for (int c = 0; c < NT; ++c) acc[c] = 0.f;
    for (int kb = 0; kb < KB; ++kb) {
      const int t = kb & 1;
      if (!umma::mbar_wait(&tfull_bar[t], (uint32_t)((kb >> 1) & 1))) s_fail = 1;
      umma::fence_after_thread_sync();
      tmem_accumulate<NT>(my_lanes + (uint32_t)(t * NT), acc);
      umma::fence_before_thread_sync();
      umma::mbar_arrive(&tempty_bar[t]);
    }
    // ---- epilogue: identical arithmetic to the v1 kernel; barriers are the named barrier 1 (128 threads)
    auto esync = []() { asm volatile("bar.sync 1, 128;\n" ::); };
    if (a.mode == CONV_EVAL) {
      if (valid) {
        float* o = a.out + (size_t)m * a.CN + n0;
        const float* rs = a.residual ? a.residual + (size_t)m * a.CN + n0 : nullptr;
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 4) {
          if (c0 >= bn) break;
          float r[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = n0 + c0 + j;
            const float inv = 1.0f / sqrtf(a.rvar[c] + a.eps);
            r[j] = (acc[c0 + j] - a.rmean[c]) * (inv * a.gamma[c]) + a.beta[c];
          }
          if (rs) {
            const float4 r4 = *reinterpret_cast<const float4*>(rs + c0);
            r[0] += r4.x; r[1] += r4.y; r[2] += r4.z; r[3] += r4.w;
          }
          if (a.relu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = fmaxf(r[j], 0.f);
          }
          *reinterpret_cast<float4*>(o + c0) = make_float4(r[0], r[1], r[2], r[3]);
        }
      }
    } else {
      if (valid) {
        float* o = a.out + (size_t)m * a.CN + n0;
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 4) {
          if (c0 >= bn) break;
          float4 r = make_float4(acc[c0], acc[c0 + 1], acc[c0 + 2], acc[c0 + 3]);
          if (a.mode == CONV_ACCUM) {
            const float4 old = *reinterpret_cast<const float4*>(o + c0);
            r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
          }
          *reinterpret_cast<float4*>(o + c0) = r;
        }
      }
      if (a.mode == CONV_TRAIN) {
        // every MMA has completed (the last tfull was consumed), so the staging ring is free
        float* s_t = stage0;  // [128][bn + 1]
#pragma unroll
        for (int c = 0; c < NT; ++c)
          if (c < bn) s_t[et * (bn + 1) + c] = acc[c];
        esync();
        if (et < bn) {
          double Sm = 0.0, Q = 0.0;
          for (int r = 0; r < 128; ++r) {
            const double x = (double)s_t[r * (bn + 1) + et];
            Sm += x;
            Q += x * x;
          }
          double* dst = a.stat_part + ((size_t)blockIdx.x * a.CN + n0 + et) * 2;
          dst[0] = Sm;
          dst[1] = Q;
        }
        __threadfence();
        esync();
        if (et == 0) is_last = (atomicAdd(a.counter + blockIdx.y, 1u) == gridDim.x - 1);
        esync();
        if (is_last) {
          __threadfence();
          const int groups = 128 / bn;
          const int ch = et % bn, grp = et / bn;
          double* s_fin = reinterpret_cast<double*>(stage0);
          esync();
          if (grp < groups) {
            double s4[4] = {0.0, 0.0, 0.0, 0.0}, q4[4] = {0.0, 0.0, 0.0, 0.0};
            unsigned int b = grp;
            for (; b + 3 * groups < gridDim.x; b += 4 * groups) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const double2 pv = __ldcg(reinterpret_cast<const double2*>(a.stat_part + ((size_t)(b + u * groups) * a.CN + n0 + ch) * 2));
                s4[u] += pv.x;
                q4[u] += pv.y;
              }
            }
            for (; b < gridDim.x; b += groups) {
              const double2 pv = __ldcg(reinterpret_cast<const double2*>(a.stat_part + ((size_t)b * a.CN + n0 + ch) * 2));
              s4[0] += pv.x;
              q4[0] += pv.y;
            }
            s_fin[(grp * bn + ch) * 2 + 0] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            s_fin[(grp * bn + ch) * 2 + 1] = (q4[0] + q4[1]) + (q4[2] + q4[3]);
          }
          esync();
          if (et < bn) {
            double Sm = 0.0, Q = 0.0;
            for (int g = 0; g < groups; ++g) {
              Sm += s_fin[(g * bn + et) * 2 + 0];
              Q += s_fin[(g * bn + et) * 2 + 1];
            }
            const double cnt = (double)a.M;
            const double mean = Sm / cnt;
            double var = Q / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            const int c = n0 + et;
            a.save_mean[c] = (float)mean;
            a.save_invstd[c] = (float)(1.0 / sqrt(var + (double)a.eps));
            const double unbiased = (a.M > 1) ? var * cnt / (cnt - 1.0) : var;
            a.run_mean[c] = (1.f - a.momentum) * a.run_mean[c] + a.momentum * (float)mean;
            a.run_var[c] = (1.f - a.momentum) * a.run_var[c] + a.momentum * (float)unbiased;
          }
        }
      }
    }
    esync();
    if (s_fail && valid) a.out[(size_t)m * a.CN + n0] = __int_as_float(0x7fc00000);
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 12) umma::tmem_dealloc(tmem, TMEM_COLS);
}

template <int NT>
int launch_tc_ws(const ConvArgs& a, cudaStream_t stream) {
  constexpr size_t smem = (size_t)3 * (2 * A_TILE + 2 * NT * 32) * sizeof(float) + 1024;
  static bool configured_dev[B200OCL_MAX_DEVICES] = {};
  bool& configured = configured_dev[b200ocl::device_slot()];
  if (!configured) {
    B200OCL_CUDA(cudaFuncSetAttribute(conv_tc_ws_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid((a.M + 127) / 128, a.CN / a.tc_bn);
  B200OCL_PROF(a.flip ? "conv_tc_dgrad" : (a.mode == CONV_EVAL ? "conv_tc_eval" : "conv_tc_train"),
               2.0 * a.M * (double)a.CN * a.CK * 9.0, stream);
  conv_tc_ws_kernel<NT><<<grid, 416, smem, stream>>>(a);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

template <int NT>
int launch_tc(const ConvArgs& a, cudaStream_t stream) {
  constexpr size_t smem = (size_t)2 * (2 * A_TILE + 2 * NT * 32) * sizeof(float) + 1024;
  static bool configured_dev[B200OCL_MAX_DEVICES] = {};
  bool& configured = configured_dev[b200ocl::device_slot()];
  if (!configured) {
    B200OCL_CUDA(cudaFuncSetAttribute(conv_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid((a.M + 127) / 128, a.CN / a.tc_bn);
  B200OCL_PROF(a.flip ? "conv_tc_dgrad" : (a.mode == CONV_EVAL ? "conv_tc_eval" : "conv_tc_train"),
               2.0 * a.M * (double)a.CN * a.CK * 9.0, stream);
  conv_tc_kernel<NT><<<grid, TC_THREADS, smem, stream>>>(a);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

}  // namespace

bool conv_tc_eligible(const ConvArgs& a) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("B200OCL_TC");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled || !a.w_tc || a.ks != 3 || a.stride != 1 || a.transposed || a.CK % 20 != 0) return false;
  if (a.Hin != a.Hout || a.Win != a.Wout) return false;
  // enough 128-pixel tiles to occupy a good part of the machine; small problems stay on the fp32 kernels
  const long ctas = (long)((a.M + 127) / 128) * (a.CN / a.tc_bn);
  if (ctas < sm_count() / 4) return false;
  // 20 -> 20 channels (layer 1): K = 180 leaves the tensor-core kernel dominated by its per-tile gather;
  // measured (tools/net_layers.py) the fp32 patch kernel wins once there are >= ~2 tiles per SM
  // (N = 110: 39 vs 57 us, N = 210: 66 vs 80 us) and loses below (N = 20: 35 vs 21 us).
  if (a.CK == 20 && a.CN == 20 && ctas > 2L * sm_count()) return false;
  return true;
}

int launch_conv_tc(const ConvArgs& a, cudaStream_t stream) {
  static int version = 0;
  if (!version) {
    const char* e = getenv("B200OCL_TC");
    // B200OCL_TC=2: warp-specialised kernel; default: lockstep kernel (2 CTAs / SM).  Measured on B200
    // (tools/umma_latency.py, profiles/r01_v3_*): the MMAs themselves are cheap -- 8 extra
    // tcgen05.mma.kind::tf32 per K block cost ~80 ns at N = 32 (~19 cycles each, the documented floor) and
    // the tensor pipe is ~10 % active -- what a K block costs is its staging chain (gather, TF32 split,
    // swizzled stores, proxy fence, barrier: ~2 us single-CTA).  Two lockstep CTAs per SM overlap that chain
    // better than one warp-specialised CTA does; the lockstep kernel is the faster of the two by ~7 %.
    version = (e && e[0] == '2') ? 2 : 1;
  }
  const int nt = a.tc_bn <= 20 ? 32 : (a.tc_bn <= 40 ? 48 : 80);
  if (version == 1) {
    if (nt == 32) return launch_tc<32>(a, stream);
    if (nt == 48) return launch_tc<48>(a, stream);
    return launch_tc<80>(a, stream);
  }
  if (nt == 32) return launch_tc_ws<32>(a, stream);
  if (nt == 48) return launch_tc_ws<48>(a, stream);
  return launch_tc_ws<80>(a, stream);
}

}  // namespace b200ocl
