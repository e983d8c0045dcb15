// wgrad_tc.cu -- weight gradient of a 3x3 stride-1 convolution on tcgen05, both operands read IN PLACE from strips (sm_100a).
//
//   dW[co][ci][kh][kw] = sum over output pixels p of  dz[p][co] * x[p + (kh-1, kw-1)][ci]
//
// The contraction runs over pixel positions, so in the layout conv_tcp.cu already stages -- the zero-padded batch as a
// strip of positions, one 128-byte row of 32 channel slots per position -- BOTH operands are
// MN-major: the M / N index (channel) is contiguous inside a row, the K index (position) is the row (stored with the
// 32-byte-base 128-byte swizzle that kind::tf32 needs for MN-major operands, umma.cuh).  With
//   strip position s = (img*(H+2) + yp)*(W+2) + xp,  x_strip[s] = x[img, yp-1, xp-1] (zero on the halo),
//   dz_strip[s] = dz[img, yp, xp] when yp < H and xp < W, zero otherwise,
// the sum becomes  dW[co][ci][kh][kw] = sum_s dz_strip[s][co] * x_strip[s + kh*(W+2) + kw][ci].
// One tcgen05.mma.kind::tf32 contracts 8 positions.  Its A operand (M = 128) is the x strip through an MN-major
// descriptor whose 32-element M blocks are ONE ROW apart (leading byte offset 128): M block j = the strip shifted by j
// positions = kernel column kw = j (the fourth block is a by-product); kernel row kh moves the start address by W+2
// rows.  Its B operand (N = 32) is the dz strip of one 32-channel block.  So D[(kw, ci)][co] for one kh and 8 positions
// is a single instruction, 9 taps x 32 x 32 channels cost 3 accumulators, and nothing is transposed or gathered.
//
// fp32 grade as everywhere else: 3xTF32 (x_hi*dz_hi + x_hi*dz_lo + x_lo*dz_hi); the TMEM accumulation of the tensor
// core truncates (umma.cuh), so a chain is cut after `tpc` tiles (default 2 = 256 positions) and written out as one fp32
// partial; the existing deterministic finalize kernel sums the partials in fixed order.
//
// CTA = (range of chains, 32-channel slice of x, 32-channel block of dz), 11 warps:
//   warps 0-3   epilogue: TMEM lane quarter = kw block, lane = input channel; tcgen05.ld of the three kh accumulators,
//               rows [k = (kh*3+kw)*Cin + ci][32 co] of the partial
//   warp  4     MMA issue (one elected lane): per tile 16 K steps x 3 kh x 3 passes
//   warps 5-8   x loaders: coalesced LDG.128 of NHWC pixels with halo, cvt.rna.tf32 split, swizzled stores
//   warps 9-10  dz loaders: the tile's 128 positions of one channel block, zero rows at halo / by-product positions
// Two staging slots (x hi / lo + dz hi / lo), two TMEM accumulator sets.
#include "common.cuh"
#include "umma.cuh"
#include "wgrad_tc.cuh"

namespace b200ocl {
namespace {

constexpr int WT_THREADS = 32 * 11;
constexpr int WT_PS = 2;
constexpr int WT_LD_MAX = 13;             // x rows a loader thread stages per tile (16 rows per pass)
constexpr int WT_DZ_BYTES = 128 * 128;    // one dz half (hi or lo) of a tile

struct WtGeom {
  int wp, pp, prow, pbytes;
};
__host__ __device__ inline WtGeom wt_geom(int H, int W) {
  WtGeom g;
  g.wp = W + 2;
  g.pp = (H + 2) * (W + 2);
  g.prow = 128 + 2 * g.wp + 2;
  g.pbytes = (g.prow * 128 + 1023) / 1024 * 1024;
  return g;
}

__global__ void __launch_bounds__(WT_THREADS, 1) wgrad_tc_kernel(WgradTcArgs a, int tiles_m) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[WT_PS], empty[WT_PS], tfull[2], tempty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_fail;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const WtGeom G = wt_geom(a.H, a.W);
  const int stage_bytes = 2 * G.pbytes + 2 * WT_DZ_BYTES;
  const int sl = blockIdx.y, cb = blockIdx.z;
  const int chain0 = blockIdx.x * a.chains_per_cta;
  const int chain1 = min(a.chains, chain0 + a.chains_per_cta);
  const int k_total = 9 * a.Cin;

  if (warp == 4) umma::tmem_alloc(&tmem_slot, 256);
  if (tid == 0) {
    for (int i = 0; i < WT_PS; ++i) {
      umma::mbar_init(&full[i], 128 + 64);
      umma::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(&tfull[i], 1);
      umma::mbar_init(&tempty[i], 128);
    }
    umma::fence_mbar_init();
    s_fail = 0;
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tmem_slot;

  if (warp >= 9) {
    // =========================================================== dz loaders (64 threads)
    const int lt = tid - 32 * 9;
    const int ch = lt & 7, r0 = lt >> 3;                         // chunk, first row; rows r0 + 8 i
    const bool ch_real = cb * 32 + ch * 4 < a.Cout;
    const int hp = a.H + 2;
    int pc = 0;
    for (int chain = chain0; chain < chain1; ++chain) {
      const int t1 = min(tiles_m, (chain + 1) * a.tpc);
      for (int tile = chain * a.tpc; tile < t1; ++tile, ++pc) {
        float4 v[16];
        int img, yp, xp;
        {
          const int sp = tile * 128 + r0;
          img = sp / G.pp;
          const int rem = sp - img * G.pp;
          yp = rem / G.wp;
          xp = rem - yp * G.wp;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ch_real && img < a.N && yp < a.H && xp < a.W)
            v[i] = __ldg(reinterpret_cast<const float4*>(a.dz + ((size_t)(img * a.H + yp) * a.W + xp) * a.Cout + cb * 32) + ch);
          xp += 8;
          while (xp >= G.wp) {
            xp -= G.wp;
            if (++yp == hp) {
              yp = 0;
              ++img;
            }
          }
        }
        const int ps = pc % WT_PS;
        if (!umma::mbar_wait(&empty[ps], (uint32_t)(((pc / WT_PS) & 1) ^ 1))) s_fail = 1;
        float* zh = reinterpret_cast<float*>(smem_raw + (size_t)ps * stage_bytes + 2 * G.pbytes);
        float* zl = zh + WT_DZ_BYTES / 4;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float4 h, l;
          umma::split_tf32(v[i].x, h.x, l.x); umma::split_tf32(v[i].y, h.y, l.y);
          umma::split_tf32(v[i].z, h.z, l.z); umma::split_tf32(v[i].w, h.w, l.w);
          const int off = umma::sw128b32_offset_f32(r0 + 8 * i, ch);
          *reinterpret_cast<float4*>(zh + off) = h;
          *reinterpret_cast<float4*>(zl + off) = l;
        }
        umma::fence_proxy_async_smem();
        umma::mbar_arrive(&full[ps]);
      }
    }
  } else if (warp >= 5) {
    // =========================================================== x loaders (128 threads)
    const int lt = tid - 32 * 5;
    const int ch = lt & 7, r0 = lt >> 3;                         // rows r0 + 16 i
    const int ch_valid = min(32, a.Cin - sl * 32);
    const bool ch_real = ch * 4 < ch_valid;
    const int nrow = G.prow;
    const int hp = a.H + 2;
    int pc = 0;
    for (int chain = chain0; chain < chain1; ++chain) {
      const int t1 = min(tiles_m, (chain + 1) * a.tpc);
      for (int tile = chain * a.tpc; tile < t1; ++tile, ++pc) {
        float4 v[WT_LD_MAX];
        int img, yp, xp;
        {
          const int sp = tile * 128 + r0;
          img = sp / G.pp;
          const int rem = sp - img * G.pp;
          yp = rem / G.wp;
          xp = rem - yp * G.wp;
        }
#pragma unroll
        for (int i = 0; i < WT_LD_MAX; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int ri = r0 + 16 * i;
          if (ch_real && ri < nrow) {
            const int y = yp - 1, x = xp - 1;
            if (img < a.N && (unsigned)y < (unsigned)a.H && (unsigned)x < (unsigned)a.W)
              v[i] = __ldg(reinterpret_cast<const float4*>(a.x + ((size_t)(img * a.H + y) * a.W + x) * a.Cin + sl * 32) + ch);
          }
          xp += 16;
          while (xp >= G.wp) {
            xp -= G.wp;
            if (++yp == hp) {
              yp = 0;
              ++img;
            }
          }
        }
        const int ps = pc % WT_PS;
        if (!umma::mbar_wait(&empty[ps], (uint32_t)(((pc / WT_PS) & 1) ^ 1))) s_fail = 1;
        float* ph = reinterpret_cast<float*>(smem_raw + (size_t)ps * stage_bytes);
        float* pl = ph + G.pbytes / 4;
        if (ch_real) {
#pragma unroll
          for (int i = 0; i < WT_LD_MAX; ++i) {
            const int ri = r0 + 16 * i;
            if (ri < nrow) {
              float4 h, l;
              umma::split_tf32(v[i].x, h.x, l.x); umma::split_tf32(v[i].y, h.y, l.y);
              umma::split_tf32(v[i].z, h.z, l.z); umma::split_tf32(v[i].w, h.w, l.w);
              const int off = umma::sw128b32_offset_f32(ri, ch);
              *reinterpret_cast<float4*>(ph + off) = h;
              *reinterpret_cast<float4*>(pl + off) = l;
            }
          }
        }
        umma::fence_proxy_async_smem();
        umma::mbar_arrive(&full[ps]);
      }
    }
  } else if (warp == 4) {
    // =========================================================== MMA issue (whole warp waits, one elected lane issues)
    const uint32_t idesc = umma::make_idesc_tf32_major(128, 32, 1, 1);
    const uint64_t dX0 = umma::make_smem_desc_mn_b32(umma::smem_u32(smem_raw), 128u);            // M blocks one row apart
    const uint64_t dZ0 = umma::make_smem_desc_mn_b32(umma::smem_u32(smem_raw) + 2u * (uint32_t)G.pbytes, 4096u);
    const uint32_t X_LO = (uint32_t)G.pbytes >> 4;
    const uint32_t Z_LO = (uint32_t)WT_DZ_BYTES >> 4;
    const uint32_t STAGE = (uint32_t)stage_bytes >> 4;
    const uint32_t x_kh = (uint32_t)(G.wp * 8);            // one kernel row = W + 2 strip rows further (16-byte units)
    int pc = 0, cc = 0;
    for (int chain = chain0; chain < chain1; ++chain, ++cc) {
      const int t = cc & 1;
      if (!umma::mbar_wait(&tempty[t], (uint32_t)(((cc >> 1) & 1) ^ 1))) s_fail = 1;
      const int t0 = chain * a.tpc, t1 = min(tiles_m, (chain + 1) * a.tpc);
      for (int tile = t0; tile < t1; ++tile, ++pc) {
        const int ps = pc % WT_PS;
        if (!umma::mbar_wait(&full[ps], (uint32_t)((pc / WT_PS) & 1))) s_fail = 1;
        umma::fence_after_thread_sync();
        if (umma::elect_one_sync()) {
          const uint64_t dXs = dX0 + (uint64_t)(ps * STAGE);
          const uint64_t dZs = dZ0 + (uint64_t)(ps * STAGE);
#pragma unroll 1
          for (int ks = 0; ks < 16; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 64);      // 8 rows = 1024 bytes
            const uint64_t dZh = dZs + adv, dZl = dZh + Z_LO;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const uint64_t dXh = dXs + adv + (uint64_t)(kh * x_kh), dXl = dXh + X_LO;
              const uint32_t dcol = tmem + (uint32_t)(t * 96 + kh * 32);
              umma::mma_tf32_ss(dcol, dXh, dZh, idesc, (tile == t0 && ks == 0) ? 0u : 1u);
              umma::mma_tf32_ss(dcol, dXh, dZl, idesc, 1u);
              umma::mma_tf32_ss(dcol, dXl, dZh, idesc, 1u);
            }
          }
          umma::mma_commit(&empty[ps]);
          if (tile == t1 - 1) umma::mma_commit(&tfull[t]);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================================================== epilogue (warps 0-3): partial [k][co] rows
    const int j = warp, c = lane;                           // kernel column, input channel inside the slice
    const uint32_t my_lanes = tmem + ((uint32_t)(warp * 32) << 16);
    const bool row_live = j < 3 && sl * 32 + c < a.Cin;
    const int n_valid = min(32, a.Cout - cb * 32);
    int cc = 0;
    for (int chain = chain0; chain < chain1; ++chain, ++cc) {
      const int t = cc & 1;
      if (!umma::mbar_wait(&tfull[t], (uint32_t)((cc >> 1) & 1))) s_fail = 1;
      umma::fence_after_thread_sync();
      uint32_t r[96];
#pragma unroll
      for (int c0 = 0; c0 < 96; c0 += 16) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(r[c0 + 0]), "=r"(r[c0 + 1]), "=r"(r[c0 + 2]), "=r"(r[c0 + 3]), "=r"(r[c0 + 4]), "=r"(r[c0 + 5]),
              "=r"(r[c0 + 6]), "=r"(r[c0 + 7]), "=r"(r[c0 + 8]), "=r"(r[c0 + 9]), "=r"(r[c0 + 10]), "=r"(r[c0 + 11]),
              "=r"(r[c0 + 12]), "=r"(r[c0 + 13]), "=r"(r[c0 + 14]), "=r"(r[c0 + 15])
            : "r"(my_lanes + (uint32_t)(t * 96 + c0)));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      umma::fence_before_thread_sync();
      umma::mbar_arrive(&tempty[t]);
      if (row_live) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          float* dst = a.part + ((size_t)chain * k_total + (size_t)(kh * 3 + j) * a.Cin + sl * 32 + c) * a.Cout + cb * 32;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (4 * q < n_valid)
              *reinterpret_cast<float4*>(dst + 4 * q) =
                  make_float4(__uint_as_float(r[kh * 32 + 4 * q]), __uint_as_float(r[kh * 32 + 4 * q + 1]),
                              __uint_as_float(r[kh * 32 + 4 * q + 2]), __uint_as_float(r[kh * 32 + 4 * q + 3]));
        }
      }
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  // a timed-out barrier (must never happen) poisons the partials instead of hanging the GPU
  if (s_fail && tid == 0) a.part[0] = __int_as_float(0x7fc00000);
  if (warp == 4) umma::tmem_dealloc(tmem, 256);
}

// selftest helper: dW[co][ci][kh][kw] = sum over chains of part[chain][(kh*3+kw)*Cin + ci][co]
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ part, int chains, int Cin, int Cout, float* __restrict__ dw) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = 9 * Cin * Cout;
  if (e >= total) return;
  const int k = e / Cout, co = e - k * Cout;
  const int tap = k / Cin, ci = k - tap * Cin;
  double s = 0.0;
  for (int sp = 0; sp < chains; ++sp) s += (double)part[(size_t)sp * total + e];
  dw[((size_t)co * Cin + ci) * 9 + tap] = (float)s;
}

}  // namespace

int launch_wgrad_tc(const WgradTcArgs& a, const WgradTcCfg& g, cudaStream_t stream) {
  const WtGeom G = wt_geom(a.H, a.W);
  const size_t smem = (size_t)WT_PS * (2 * (size_t)G.pbytes + 2 * WT_DZ_BYTES) + 1024;
  static size_t configured_dev[B200OCL_MAX_DEVICES] = {};
  size_t& configured = configured_dev[device_slot()];
  if (smem > configured) {
    B200OCL_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const double M = (double)a.N * a.H * a.W;
  B200OCL_PROF("wgrad_tc", 2.0 * M * 9.0 * a.Cin * a.Cout, stream);
  wgrad_tc_kernel<<<dim3(g.ctas_x, g.slices, g.cout_blocks), WT_THREADS, smem, stream>>>(a, g.tiles);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

}  // namespace b200ocl

extern "C" size_t b200ocl_wgrad_tc_selftest_workspace_bytes(int N, int H, int W, int cin, int cout) {
  using namespace b200ocl;
  if (N < 1 || H < 1 || W < 1 || cin < 1 || cout < 1) return 0;
  const WgradTcCfg g = wgrad_tc_cfg(N, H, W, 3, 1, 1, cin, cout, sm_count());
  if (!g.eligible) return 0;
  return align_up((size_t)g.chains * 9 * cin * cout * sizeof(float), 256);
}

extern "C" int b200ocl_wgrad_tc_selftest(const float* x, const float* dz, float* dw_oihw, int N, int H, int W, int cin, int cout,
                                         void* workspace, size_t workspace_bytes, void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200OCL_CHECK_ARG(x && dz && dw_oihw && N >= 1, "null pointer / empty batch");
  const WgradTcCfg g = wgrad_tc_cfg(N, H, W, 3, 1, 1, cin, cout, sm_count());
  if (!g.eligible) {
    set_error("b200ocl_wgrad_tc_selftest: geometry not covered (3x3 stride 1, W <= 37, channels %% 4 == 0)");
    return B200OCL_EUNSUPPORTED;
  }
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255) ||
      workspace_bytes < b200ocl_wgrad_tc_selftest_workspace_bytes(N, H, W, cin, cout)) {
    set_error("b200ocl_wgrad_tc_selftest: workspace missing, misaligned or too small");
    return B200OCL_EWORKSPACE;
  }
  WgradTcArgs a{};
  a.x = x; a.dz = dz; a.part = static_cast<float*>(workspace);
  a.N = N; a.H = H; a.W = W; a.Cin = cin; a.Cout = cout;
  a.tpc = g.tpc; a.chains = g.chains; a.chains_per_cta = g.chains_per_cta;
  const int rc = launch_wgrad_tc(a, g, stream);
  if (rc) return rc;
  const int total = 9 * cin * cout;
  wgrad_tc_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(a.part, g.chains, cin, cout, dw_oihw);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}
