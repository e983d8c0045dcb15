// conv_tcp.cu -- 3x3 stride-1 convolution on tcgen05 fed IN PLACE from a halo patch (sm_100a).
//
// conv_tc.cu builds an im2col tile per K block: every output pixel re-gathers, re-splits and re-stores
// its input 9 times, and that staging chain -- not the tensor core -- is what a K block costs.  Here the
// input patch of a tile (with its halo) is staged ONCE per 32-channel slice as two 128-byte-swizzled
// arrays (TF32 hi / lo) of one 128-byte row per patch pixel, and the A operand of every tap is that
// same array read through a shifted shared-memory descriptor:
//
//   strip  = the zero-padded input of the whole batch, row-major: position s = (img*(H+2) + yp)*(W+2) + xp,
//            one 128-byte row (32 channel slots) per position, the halo positions hold zeros
//   tile   = 128 consecutive strip positions = the 128 rows of an MMA; row i is the output pixel
//            (img, y = yp, x = xp) when yp < H and xp < W, a discarded by-product otherwise
//   tap (kh, kw) of row i reads strip position  s_i + kh*(W+2) + kw  -> one descriptor per tap whose start
//            address is the patch base + (kh*(W+2) + kw) * 128 bytes; rows stay consecutive, so the 8-row
//            groups are the standard 1024 bytes apart.
// A stage therefore holds 128 + 2*(W+2) + 2 strip rows.  Any H, W works; the share of useful rows is
// H*W / ((H+2)*(W+2)): 89 % at 32x32, 79 % at 16x16, 64 % at 8x8, 44 % at 4x4 -- tensor-core time is not
// what bounds the kernel, and no im2col or per-shape tiling is needed.
//
// The hardware applies the 128-byte swizzle to absolute address bits, so a window that starts at any
// 128-byte row of a patch stored with "chunk ^= row & 7" reads back exactly (tests/test_gpu_umma.py,
// b200ocl_selftest_umma_window: every start row, base-offset field 0).
//
// Per tile and slice the tensor core runs 9 taps x ceil(channels/8) K steps x 3 MMAs (3xTF32:
// hi*hi + hi*lo + lo*hi); each tap accumulates into its own TMEM buffer which the promotion warps
// add into fp32 registers (a long TMEM accumulation truncates, see conv_tc.cu).  Roles:
//   warps 0-3  promotion + epilogue (TMEM lane quarter = warp), folded eval BN / raw + batch statistics
//   warps 4-6  MMA issue: K block q (tile, slice, tap) belongs to warp 4 + q % 3; one elected lane issues its
//              MMAs and commits -- a single issuing warp is latency-bound at ~1 us per K block, three keep
//              the tensor core fed; warp 4 also owns the TMEM allocation
//   warp  7    weight loader (one elected lane): cp.async.bulk of pre-split, pre-swizzled [NT x 32] tiles,
//              resident for the whole kernel when all 9 taps fit (cin <= 32), a 4-deep ring otherwise
//   warps 8-11 patch loaders: coalesced LDG.128 of NHWC pixels, cvt.rna.tf32 split, swizzled stores
// CTAs are persistent over pixel tiles (one CTA per SM); batch statistics are accumulated per CTA in
// fp64 in a fixed order and finalised by the last CTA (deterministic, as everywhere else).
#include <stdlib.h>

#include "conv.cuh"
#include "umma.cuh"

namespace b200ocl {
namespace {

constexpr int TP_MW = 3;                            // MMA-issuing warps: warp w issues the taps of kernel column kw = w
constexpr int TP_THREADS = 32 * (4 + TP_MW + 1 + 4);
constexpr int TP_PS_MAX = 3;                        // patch stages: 3 when shared memory allows, else 2
constexpr int TP_NB = 6;                            // TMEM accumulator buffers: two per MMA warp
constexpr int TP_BS_MAX = 6;                        // weight ring depth (streaming mode): 6 or 4
constexpr int TP_LD_MAX = 13;                       // patch rows a loader thread stages per tile (16 rows per pass)

using umma::mbar_expect_tx;
using umma::bulk_g2s;

// acc += the NT fp32 columns of this warp's 32 TMEM lanes: all loads issued, one wait
// comp: EXPERIMENTAL compensation of the TMEM accumulate's rounding (B200OCL_TCP_DEBIAS, default off).  Round-2 finding:
// tcgen05 kind::tf32 accumulation truncates at every accumulate step, so a read-out value behaves like z - c|z| for a
// small data-dependent c -- not a scale factor (a relative correction changes nothing, BatchNorm absorbs it) but a kink
// at zero that a train-mode forward differentiates.  v + comp * |v| with comp = n_mma * 2^-25 removes the whole effect in
// the A-GEM drop-in case (whole update vector 4.27e-5 off the live reference, exactly the fp32 kernels' 4.26e-5, against
// 1.0e-2 without) but over-corrects on zero-mean random data (rms against fp64 2.4e-7 -> 6.5e-7 .. 1.7e-6,
// tools/tcp_chain_accuracy.py) and breaks the golden-state gradient test: c depends on how the partial sums grow.  Until it
// is modelled properly the train-mode forward stays on the fp32 kernels (conv_tcp_mode_allowed below).  (Accumulating odd
// taps negated -- a_negate, idesc bit 13 -- and subtracting them at promotion gives bit-identical results: the rounding is
// sign-symmetric, i.e. a truncation of the magnitude, so alternating signs cannot cancel it.)
template <int NT>
__device__ __forceinline__ void tmem_accumulate(uint32_t taddr, float (&acc)[NT], float comp = 0.f) {
  uint32_t r[NT];
#pragma unroll
  for (int c0 = 0; c0 < NT; c0 += 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[c0 + 0]), "=r"(r[c0 + 1]), "=r"(r[c0 + 2]), "=r"(r[c0 + 3]), "=r"(r[c0 + 4]), "=r"(r[c0 + 5]),
          "=r"(r[c0 + 6]), "=r"(r[c0 + 7]), "=r"(r[c0 + 8]), "=r"(r[c0 + 9]), "=r"(r[c0 + 10]), "=r"(r[c0 + 11]),
          "=r"(r[c0 + 12]), "=r"(r[c0 + 13]), "=r"(r[c0 + 14]), "=r"(r[c0 + 15])
        : "r"(taddr + (uint32_t)c0));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int c = 0; c < NT; ++c) {
    const float v = __uint_as_float(r[c]);
    acc[c] += fmaf(comp, fabsf(v), v);
  }
}

struct TileGeom {
  int wp, pp;        // padded row length W + 2, padded image size (H + 2) * (W + 2)
  int tiles_m;       // tiles of 128 strip positions
  int prow;          // strip rows staged per tile: 128 + 2 * wp + 2
  int pbytes;        // bytes of one hi (or lo) patch, 1024-aligned
};
__host__ __device__ inline TileGeom tile_geom(int N, int H, int W) {
  TileGeom g;
  g.wp = W + 2;
  g.pp = (H + 2) * (W + 2);
  // the last useful position is the last real pixel of the last image
  const long last = (long)(N - 1) * g.pp + (long)(H - 1) * g.wp + (W - 1);
  g.tiles_m = (int)(last / 128) + 1;
  g.prow = 128 + 2 * g.wp + 2;
  g.pbytes = (g.prow * 128 + 1023) / 1024 * 1024;
  return g;
}

// KL: K steps (8 channels each) of the LAST 32-channel slice; every other slice has 4.  Compile-time so that the
// issue sequence of a K block is straight-line code with immediate descriptor offsets (KL = 0: run-time loop).
template <int KS>
__device__ __forceinline__ void issue_kblock(uint32_t dcol, uint64_t dAh, uint64_t dAl, uint64_t dBh, uint64_t dBl,
                                             uint32_t idesc, uint32_t acc0) {
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const uint64_t adv = (uint64_t)(k * 2);   // 32 bytes per K step, in 16-byte units
    umma::mma_tf32_ss(dcol, dAh + adv, dBh + adv, idesc, k > 0 ? 1u : acc0);
    umma::mma_tf32_ss(dcol, dAh + adv, dBl + adv, idesc, 1u);
    umma::mma_tf32_ss(dcol, dAl + adv, dBh + adv, idesc, 1u);
  }
}

template <int NT, int KL>
__global__ void __launch_bounds__(TP_THREADS, 1) conv_tcp_kernel(ConvArgs a) {
  constexpr uint32_t TMEM_COLS = (TP_NB * NT <= 128) ? 128 : ((TP_NB * NT <= 256) ? 256 : 512);
  constexpr int B_BLOCK = 2 * NT * 32;   // floats: hi tile then lo tile
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* patch0 = smem_raw;                                          // [PS][hi | lo][PROWS][128 B]
  __shared__ __align__(8) uint64_t pfull[TP_PS_MAX], pempty[TP_PS_MAX], bfull[9], bempty[TP_BS_MAX], tfull[TP_NB], tempty[TP_NB];
  __shared__ uint32_t tmem_slot;
  __shared__ bool is_last;
  __shared__ int s_fail;
  __shared__ float s_coef[3 * 80];   // eval BN: mean, scale, shift per channel of the tile

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bn = a.tp_bn;
  const int n0 = blockIdx.y * bn;
  const int slices = a.tp_slices;
  const TileGeom G = tile_geom(a.N, a.Hin, a.Win);
  const int PS = a.tp_ps, BS = a.tp_bs;            // patch stages, weight ring depth (launcher fits them to smem)
  const int taps_n = a.ks * a.ks;                  // 9, or 1 (1x1 convolution = the centre tap of the padded geometry)
  const bool one_tap = taps_n == 1;
  const bool chain3 = a.tp_chain == 3 && !one_tap;
  float* sB = reinterpret_cast<float*>(smem_raw + (size_t)PS * 2 * G.pbytes);   // weight blocks
  const bool resident = (slices == 1 && NT == 32);   // all 9 weight blocks stay in shared memory
  const int b_slots = resident ? 9 : BS;
  float* s_t = sB + (size_t)b_slots * B_BLOCK;  // [128][bn + 1] transpose scratch for the batch statistics

  if (warp == 4) umma::tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) {
    for (int i = 0; i < TP_PS_MAX; ++i) {
      umma::mbar_init(&pfull[i], 128);
      umma::mbar_init(&pempty[i], one_tap ? 1 : TP_MW);
    }
    for (int i = 0; i < 9; ++i) umma::mbar_init(&bfull[i], 1);
    for (int i = 0; i < TP_BS_MAX; ++i) umma::mbar_init(&bempty[i], 1);
    for (int i = 0; i < TP_NB; ++i) {
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
  const float* wimg = a.w_tp + (size_t)blockIdx.y * slices * taps_n * B_BLOCK;

  if (warp >= 5 + TP_MW) {
    // =========================================================== patch loaders (128 threads)
    const int lt = tid - 32 * (5 + TP_MW);
    int pc = 0;
    for (int tile = blockIdx.x; tile < G.tiles_m; tile += gridDim.x) {
      for (int sl = 0; sl < slices; ++sl, ++pc) {
        const int ch_valid = min(32, a.CK - sl * 32);       // real channels in this slice (multiple of 4)
        const int nch = 2 * ((ch_valid + 7) / 8);            // 16-byte chunks the MMAs will read per row
        // thread -> (patch row lt/8 + 16*i, chunk lt%8): lanes with chunk >= nch idle
        const int ch = lt & 7, r0 = lt >> 3;
        const int nrow = G.prow;
        const bool ch_live = ch < nch, ch_real = ch * 4 < ch_valid;
        float4 v[TP_LD_MAX];
        // strip position of this thread's first row, then 16 rows further per pass (no divisions in the loop)
        int img, yp, xp;
        {
          const int sp = tile * 128 + r0;
          img = sp / G.pp;
          const int rem = sp - img * G.pp;
          yp = rem / G.wp;
          xp = rem - yp * G.wp;
        }
        const int hp = a.Hin + 2;
#pragma unroll
        for (int i = 0; i < TP_LD_MAX; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int ri = r0 + 16 * i;
          if (ch_real && ri < nrow) {
            const int y = yp - 1, x = xp - 1;
            if (img < a.N && (unsigned)y < (unsigned)a.Hin && (unsigned)x < (unsigned)a.Win)
              v[i] = __ldg(reinterpret_cast<const float4*>(a.in + ((size_t)(img * a.Hin + y) * a.Win + x) * a.CK + sl * 32) + ch);
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
        const int ps = pc % PS;
        if (!umma::mbar_wait(&pempty[ps], (uint32_t)(((pc / PS) & 1) ^ 1))) s_fail = 1;
        float* ph = reinterpret_cast<float*>(patch0 + (size_t)ps * 2 * G.pbytes);
        float* pl = ph + G.pbytes / 4;
        if (ch_live) {
#pragma unroll
          for (int i = 0; i < TP_LD_MAX; ++i) {
            const int ri = r0 + 16 * i;
            if (ri < nrow) {
              float4 h, l;
              umma::split_tf32(v[i].x, h.x, l.x); umma::split_tf32(v[i].y, h.y, l.y);
              umma::split_tf32(v[i].z, h.z, l.z); umma::split_tf32(v[i].w, h.w, l.w);
              const int off = umma::sw128_offset_f32(ri, ch);
              *reinterpret_cast<float4*>(ph + off) = h;
              *reinterpret_cast<float4*>(pl + off) = l;
            }
          }
        }
        umma::fence_proxy_async_smem();
        umma::mbar_arrive(&pfull[ps]);
      }
    }
  } else if (warp == 4 + TP_MW) {
    // =========================================================== weight loader (one elected lane)
    const uint32_t bytes = (uint32_t)(B_BLOCK * sizeof(float));
    if (resident) {
      if (umma::elect_one_sync()) {
        for (int b = 0; b < taps_n; ++b) {
          mbar_expect_tx(&bfull[b], bytes);
          bulk_g2s(sB + (size_t)b * B_BLOCK, wimg + (size_t)b * B_BLOCK, bytes, &bfull[b]);
        }
      }
    } else {
      int q = 0;
      for (int tile = blockIdx.x; tile < G.tiles_m; tile += gridDim.x)
        for (int sl = 0; sl < slices; ++sl)
          for (int tap = 0; tap < taps_n; ++tap, ++q) {
            const int bs = q % BS;
            if (!umma::mbar_wait(&bempty[bs], (uint32_t)(((q / BS) & 1) ^ 1))) s_fail = 1;
            if (umma::elect_one_sync()) {
              mbar_expect_tx(&bfull[bs], bytes);
              bulk_g2s(sB + (size_t)bs * B_BLOCK, wimg + (size_t)(sl * taps_n + tap) * B_BLOCK, bytes, &bfull[bs]);
            }
            __syncwarp();
          }
    }
  } else if (warp >= 4) {
    // =========================================================== MMA issuers (whole warp waits, one elected lane issues)
    // Warp mw owns kernel column kw = mw (taps kh*3 + mw) and TMEM buffers 2*mw, 2*mw + 1, which it fills
    // alternately; the promotion warps drain the taps in tap order, so the fp32 sum order is fixed.
    const int mw = warp - 4;
    const uint32_t idesc = umma::make_idesc_tf32(128, NT);
    // descriptor templates: only the 14-bit start-address field (16-byte units) changes per stage / tap / K step
    const uint64_t dA0 = umma::make_smem_desc_sw128(umma::smem_u32(patch0)) + (uint64_t)(mw * 8);   // + kw rows of 128 B
    const uint64_t dB0 = umma::make_smem_desc_sw128(umma::smem_u32(sB));
    const uint32_t A_LO = (uint32_t)G.pbytes >> 4;            // hi -> lo patch, 16-byte units
    const uint32_t A_STAGE = (uint32_t)(2 * G.pbytes) >> 4;
    constexpr uint32_t B_LO = (NT * 128) >> 4;
    constexpr uint32_t B_SLOT = (B_BLOCK * 4) >> 4;
    const uint32_t a_kh = (uint32_t)(G.wp * 8);              // one kernel row = W + 2 strip rows further
    int cnt = 0, pc = 0;                                      // K blocks issued by this warp; (tile, slice) pairs
    bool b_ready = false;
    const bool warp_active = !one_tap || mw == 1;     // a 1x1 convolution only has the centre tap (kh = kw = 1)
    for (int tile = blockIdx.x; warp_active && tile < G.tiles_m; tile += gridDim.x) {
      for (int sl = 0; sl < slices; ++sl, ++pc) {
        const int ps = pc % PS;
        const int ksteps = (min(32, a.CK - sl * 32) + 7) / 8;
        if (!umma::mbar_wait(&pfull[ps], (uint32_t)((pc / PS) & 1))) s_fail = 1;
        const uint64_t dAs = dA0 + (uint64_t)(ps * A_STAGE);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          if (one_tap && kh != 1) continue;
          const int tap = one_tap ? 0 : kh * 3 + mw;   // index of the weight block inside the slice
          const int q = pc * taps_n + tap;             // position in the weight stream
          const int b = resident ? tap : q % BS;
          // chain == 1: every tap has its own TMEM buffer and is promoted separately; chain == 3: the three
          // taps of this warp's kernel column accumulate in one buffer (a TMEM chain of <= 96 products)
          const int cb = chain3 ? pc : cnt;
          const int t = 2 * mw + (cb & 1);
          if (!(resident && b_ready))
            if (!umma::mbar_wait(&bfull[b], resident ? 0u : (uint32_t)((q / BS) & 1))) s_fail = 1;
          if (!chain3 || kh == 0)
            if (!umma::mbar_wait(&tempty[t], (uint32_t)(((cb >> 1) & 1) ^ 1))) s_fail = 1;
          umma::fence_after_thread_sync();
          const uint64_t dAh = dAs + (uint64_t)(kh * a_kh);
          const uint64_t dAl = dAh + A_LO;
          const uint64_t dBh = dB0 + (uint64_t)(b * B_SLOT);
          const uint64_t dBl = dBh + B_LO;
          const uint32_t dcol = tmem + (uint32_t)(t * NT);
          const uint32_t acc0 = (chain3 && kh > 0) ? 1u : 0u;
          if (umma::elect_one_sync()) {
            if (KL == 0) {
              for (int k = 0; k < ksteps; ++k) {
                const uint64_t adv = (uint64_t)(k * 2);
                umma::mma_tf32_ss(dcol, dAh + adv, dBh + adv, idesc, k > 0 ? 1u : acc0);
                umma::mma_tf32_ss(dcol, dAh + adv, dBl + adv, idesc, 1u);
                umma::mma_tf32_ss(dcol, dAl + adv, dBh + adv, idesc, 1u);
              }
            } else if (sl == slices - 1) {
              issue_kblock<(KL > 0 ? KL : 1)>(dcol, dAh, dAl, dBh, dBl, idesc, acc0);
            } else {
              issue_kblock<4>(dcol, dAh, dAl, dBh, dBl, idesc, acc0);
            }
            if (!resident) umma::mma_commit(&bempty[b]);
            if (!chain3 || kh == 2) umma::mma_commit(&tfull[t]);
            if (kh == 2 || one_tap) umma::mma_commit(&pempty[ps]);
          }
          ++cnt;
          __syncwarp();
        }
        b_ready = true;
      }
    }
  } else {
    // =========================================================== promotion + epilogue (warps 0-3)
    const int et = tid;                        // 0..127 = MMA row = TMEM lane
    const uint32_t my_lanes = tmem + ((uint32_t)(warp * 32) << 16);
    auto esync = []() { asm volatile("bar.sync 1, 128;\n" ::); };
    if (a.mode == CONV_EVAL) {
      for (int c = et; c < bn; c += 128) {
        const float inv = 1.0f / sqrtf(a.rvar[n0 + c] + a.eps);
        s_coef[c] = a.rmean[n0 + c];
        s_coef[80 + c] = inv * a.gamma[n0 + c];
        s_coef[160 + c] = a.beta[n0 + c];
      }
      esync();
    }
    double statS = 0.0, statQ = 0.0;           // thread c < bn: running sums of channel c over this CTA's tiles
    int q = 0;
    for (int tile = blockIdx.x; tile < G.tiles_m; tile += gridDim.x) {
      const int sp = tile * 128 + et;                    // strip position of this MMA row
      const int img = sp / G.pp, rem = sp - img * G.pp;
      const int y = rem / G.wp, x = rem - y * G.wp;
      // halo positions and (stride 2) odd positions are by-products
      const bool valid = img < a.N && y < a.Hin && x < a.Win && (a.stride == 1 || ((y | x) & 1) == 0);
      const size_t m = ((size_t)img * a.Hout + (a.stride == 1 ? y : y >> 1)) * a.Wout + (a.stride == 1 ? x : x >> 1);
      // The accumulators start from what the epilogue would otherwise have to fetch after the last tap --
      // the residual (eval) or the gradient being accumulated into (data gradient) -- so that global
      // latency is paid while the tensor core works on the tile, not after it.
      float acc[NT], pre[NT];
#pragma unroll
      for (int c = 0; c < NT; ++c) {
        acc[c] = 0.f;
        pre[c] = 0.f;
      }
      {
        const float* src = nullptr;
        if (valid && a.mode == CONV_EVAL && a.residual) src = a.residual + m * a.CN + n0;
        if (valid && a.mode == CONV_ACCUM) src = a.out + m * a.CN + n0;
        if (src) {
#pragma unroll
          for (int c0 = 0; c0 < NT; c0 += 4) {
            if (c0 >= bn) break;
            const float4 v = *reinterpret_cast<const float4*>(src + c0);
            pre[c0] = v.x; pre[c0 + 1] = v.y; pre[c0 + 2] = v.z; pre[c0 + 3] = v.w;
          }
        }
      }
      for (int sl = 0; sl < slices; ++sl, ++q) {        // q counts (tile, slice) pairs here
#pragma unroll 1
        for (int tap = 0; tap < (chain3 ? 3 : taps_n); ++tap) {
          const int kh = (chain3 || one_tap) ? 0 : tap / 3, mwq = one_tap ? 1 : tap - 3 * kh;
          const int cq = (chain3 || one_tap) ? q : q * 3 + kh;   // buffers MMA warp mwq filled before this one
          const int t = 2 * mwq + (cq & 1);
          if (!umma::mbar_wait(&tfull[t], (uint32_t)((cq >> 1) & 1))) s_fail = 1;
          umma::fence_after_thread_sync();
          {
            // 3 MMAs per K step of this slice have accumulated into the buffer (x3 with 3-tap chains)
            const float n_mma = 3.f * (float)((min(32, a.CK - sl * 32) + 7) / 8) * (chain3 ? 3.f : 1.f);
            tmem_accumulate<NT>(my_lanes + (uint32_t)(t * NT), acc, a.tp_debias * n_mma * 2.9802322e-8f);   // 2^-25
          }
          umma::fence_before_thread_sync();
          umma::mbar_arrive(&tempty[t]);
        }
      }
      if (a.mode == CONV_EVAL) {
        if (valid) {
          float* o = a.out + m * a.CN + n0;
#pragma unroll
          for (int c0 = 0; c0 < NT; c0 += 4) {
            if (c0 >= bn) break;
            float rr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              rr[j] = (acc[c0 + j] - s_coef[c0 + j]) * s_coef[80 + c0 + j] + s_coef[160 + c0 + j];
              rr[j] += pre[c0 + j];                       // residual (0 when there is none)
              if (a.relu) rr[j] = fmaxf(rr[j], 0.f);
            }
            *reinterpret_cast<float4*>(o + c0) = make_float4(rr[0], rr[1], rr[2], rr[3]);
          }
        }
      } else {
        if (valid) {
          float* o = a.out + m * a.CN + n0;
#pragma unroll
          for (int c0 = 0; c0 < NT; c0 += 4) {
            if (c0 >= bn) break;
            // raw / train: pre == 0; accumulate: pre = previous contents
            *reinterpret_cast<float4*>(o + c0) = make_float4(acc[c0] + pre[c0], acc[c0 + 1] + pre[c0 + 1],
                                                             acc[c0 + 2] + pre[c0 + 2], acc[c0 + 3] + pre[c0 + 3]);
          }
        }
        if (a.mode == CONV_TRAIN) {
          // transpose through shared memory; the 128 rows of a channel are summed in fp64 by 128 / bn threads
          // (contiguous row ranges, combined in range order: the association is fixed)
#pragma unroll
          for (int c = 0; c < NT; ++c)
            if (c < bn) s_t[et * (bn + 1) + c] = valid ? acc[c] : 0.f;   // by-product rows do not count
          esync();
          const int parts = 128 / bn;                      // 6 (bn = 20) or 3 (bn = 40)
          const int rows_pp = (128 + parts - 1) / parts;
          const int ch = et % bn, part = et / bn;
          double S = 0.0, Q = 0.0;
          if (part < parts) {
            const int r1 = min(128, (part + 1) * rows_pp);
            for (int rr = part * rows_pp; rr < r1; ++rr) {
              const double xv = (double)s_t[rr * (bn + 1) + ch];
              S += xv;
              Q += xv * xv;
            }
          }
          esync();                                         // s_t fully read
          double* s_p = reinterpret_cast<double*>(s_t);    // [parts][bn][2] (<= 1920 B)
          if (part < parts) {
            s_p[(part * bn + ch) * 2 + 0] = S;
            s_p[(part * bn + ch) * 2 + 1] = Q;
          }
          esync();
          if (et < bn) {
            double Ss = 0.0, Qs = 0.0;
            for (int pp = 0; pp < parts; ++pp) {
              Ss += s_p[(pp * bn + et) * 2 + 0];
              Qs += s_p[(pp * bn + et) * 2 + 1];
            }
            statS += Ss;
            statQ += Qs;
          }
          esync();
        }
      }
    }
    if (a.mode == CONV_TRAIN) {
      if (et < bn) {
        double* dst = a.stat_part + ((size_t)blockIdx.x * a.CN + n0 + et) * 2;
        dst[0] = statS;
        dst[1] = statQ;
      }
      __threadfence();
      esync();
      if (et == 0) is_last = (atomicAdd(a.counter + blockIdx.y, 1u) == gridDim.x - 1);
      esync();
      if (is_last) {
        __threadfence();
        const int groups = 128 / bn;
        const int ch = et % bn, grp = et / bn;
        double* s_fin = reinterpret_cast<double*>(s_t);   // [groups][bn][2]
        if (grp < groups) {
          double S = 0.0, Q = 0.0;
          unsigned int b = grp;
          for (; b + 3 * groups < gridDim.x; b += 4 * groups) {
            double2 pv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              pv[u] = __ldcg(reinterpret_cast<const double2*>(a.stat_part + ((size_t)(b + u * groups) * a.CN + n0 + ch) * 2));
            S += (pv[0].x + pv[1].x) + (pv[2].x + pv[3].x);
            Q += (pv[0].y + pv[1].y) + (pv[2].y + pv[3].y);
          }
          for (; b < gridDim.x; b += groups) {
            const double2 pv = __ldcg(reinterpret_cast<const double2*>(a.stat_part + ((size_t)b * a.CN + n0 + ch) * 2));
            S += pv.x;
            Q += pv.y;
          }
          s_fin[(grp * bn + ch) * 2 + 0] = S;
          s_fin[(grp * bn + ch) * 2 + 1] = Q;
        }
        esync();
        if (et < bn) {
          double Sm = 0.0, Q = 0.0;
          for (int gq = 0; gq < groups; ++gq) {
            Sm += s_fin[(gq * bn + et) * 2 + 0];
            Q += s_fin[(gq * bn + et) * 2 + 1];
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
  umma::fence_before_thread_sync();
  __syncthreads();
  // a timed-out barrier (must never happen) poisons the output instead of hanging the GPU
  if (s_fail && tid == 0) a.out[(size_t)n0] = __int_as_float(0x7fc00000);
  if (warp == 4) umma::tmem_dealloc(tmem, TMEM_COLS);   // the allocating warp
}

template <int NT>
size_t tcp_smem_bytes(const TileGeom& G, int bn, int slices, int ps, int bs) {
  const int b_slots = (slices == 1 && NT == 32) ? 9 : bs;
  return (size_t)ps * 2 * G.pbytes + (size_t)b_slots * 2 * NT * 32 * sizeof(float) +
         (size_t)128 * (bn + 1) * sizeof(float) + 1024;
}

template <int NT, int KL>
int launch_tcp(ConvArgs a, cudaStream_t stream) {
  const TileGeom G0 = tile_geom(a.N, a.Hin, a.Win);
  // deepest pipeline that fits: 3 patch stages + 6 weight slots, 3 + 4, else 2 + 6
  const size_t limit = 227 * 1024 - 4096;     // static shared memory (barriers, coefficients) comes on top
  a.tp_ps = 3; a.tp_bs = 6;
  {
    // TMEM accumulation chains: every tap is promoted to fp32 registers separately (default).  Chaining the three
    // taps of a kernel column in TMEM (B200OCL_TCP_CHAIN=3) was measured again in round 2, for the train-mode and
    // data-gradient launches only: no change of the step time (the kernel is bound by shared-memory bandwidth --
    // three SS-mode MMAs re-read the 4 KB A window per 8 channels for N = 32..48 columns -- not by the number of
    // promotions) while the BN-weight gradient of layer2.0 moved from 2e-4 to 1.9e-3 off the reference: not taken.
    // Likewise a software-pipelined patch loader (loads of stage pc+1 issued before stage pc is converted): the
    // class went from 4.15 to 4.3 ms per step pair -- the loaders were not the bound either -- and was reverted.
    const char* e = getenv("B200OCL_TCP_CHAIN");
    a.tp_chain = (e && e[0] == '3') ? 3 : 1;
    const char* db = getenv("B200OCL_TCP_DEBIAS");      // multiplier of the experimental rounding compensation (default 0 = off)
    a.tp_debias = db ? (float)atof(db) : 0.f;
  }
  if (tcp_smem_bytes<NT>(G0, a.tp_bn, a.tp_slices, a.tp_ps, a.tp_bs) > limit) a.tp_bs = 4;
  if (tcp_smem_bytes<NT>(G0, a.tp_bn, a.tp_slices, a.tp_ps, a.tp_bs) > limit) { a.tp_ps = 2; a.tp_bs = 6; }
  const size_t smem = tcp_smem_bytes<NT>(G0, a.tp_bn, a.tp_slices, a.tp_ps, a.tp_bs);
  static size_t configured_dev[B200OCL_MAX_DEVICES] = {};
  size_t& configured = configured_dev[b200ocl::device_slot()];
  if (smem > configured) {
    B200OCL_CUDA(cudaFuncSetAttribute(conv_tcp_kernel<NT, KL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const TileGeom G = tile_geom(a.N, a.Hin, a.Win);
  const int n_tiles = a.CN / a.tp_bn;
  int gx = sm_count() / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > G.tiles_m) gx = G.tiles_m;
  // even out the tiles per CTA (e.g. 880 tiles on 148 CTAs = 6 rounds -> 147 CTAs of exactly 6)
  const int rounds = (G.tiles_m + gx - 1) / gx;
  gx = (G.tiles_m + rounds - 1) / rounds;
  // one kernel, three epilogues (eval / train / data gradient): profiled as one class
  B200OCL_PROF("conv_tcp",
               2.0 * a.M * (double)a.CN * a.CK * 9.0, stream);
  conv_tcp_kernel<NT, KL><<<dim3(gx, n_tiles), TP_THREADS, smem, stream>>>(a);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

}  // namespace

bool conv_tcp_mode_allowed(const ConvArgs& a) {
  {
    // Which launch kinds take the tensor-core path: bit 0 eval features, bit 1 train-mode forward, bit 2 data gradient
    // (B200OCL_TCP_MODES).  Default 5 = eval + data gradient.  Measured in round 2 against the live reference
    // (tests/test_gpu_dropin.py, A-GEM case, one batch-10 step from torch's default initialisation): with the TRAIN-MODE
    // FORWARD on this path the update of conv1.weight is 1.0e-2 off the reference (whose own one-ulp spread is 4e-6) and
    // only 8 % of the tensors are inside 1e-3; with it on the fp32 kernels every tensor is inside 5e-5 -- and it makes no
    // difference whether eval features and data gradients use the tensor path (modes 1 and 5 give the same 5e-5).  Cause:
    // the downward rounding of the TMEM accumulate, see tmem_accumulate().  Step pair: 8.59 ms (7) / 8.87 (5) / 9.26 (1).
    static int modes = -1;
    if (modes < 0) {
      const char* m = getenv("B200OCL_TCP_MODES");
      modes = (m && m[0] >= '0' && m[0] <= '7') ? (m[0] - '0') : 5;
    }
    const int bit = a.mode == CONV_EVAL ? 1 : (a.mode == CONV_TRAIN ? 2 : 4);
    return (modes & bit) != 0;
  }
}

bool conv_tcp_eligible(const ConvArgs& a) {
  // read per call (a getenv is negligible next to a launch) so that tools can switch paths in-process
  const char* e = getenv("B200OCL_TCP");
  const char* e2 = getenv("B200OCL_TC");
  const bool enabled = !((e && e[0] == '0') || (e2 && e2[0] == '0'));
  if (!enabled || !a.w_tp || a.transposed || a.CK % 4 != 0) return false;
  if (!((a.ks == 3 && a.pad == 1) || (a.ks == 1 && a.pad == 0))) return false;
  if (a.stride != 1 && (a.stride != 2 || a.flip)) return false;          // stride 2: forward only
  if (a.Hout != (a.Hin + 2 * a.pad - a.ks) / a.stride + 1 || a.Wout != (a.Win + 2 * a.pad - a.ks) / a.stride + 1) return false;
  if (128 + 2 * (a.Win + 2) + 2 > 16 * TP_LD_MAX) return false;   // strip rows one stage holds (W <= 37)
  if ((long)a.N * (a.Hin + 2) * (a.Win + 2) > 2000000000L) return false;
  if (a.tp_bn <= 0 || a.tp_bn > 40 || a.CN % a.tp_bn != 0) return false;
  // stat_part holds one row per persistent CTA; sized for conv_max_grid_m(M) >= tiles
  return true;
}

int launch_conv_tcp(const ConvArgs& a, cudaStream_t stream) {
  const int kl = (a.CK - 32 * (a.tp_slices - 1) + 7) / 8;   // K steps of the last slice
  if (a.tp_bn <= 20) {
    if (kl == 3 && a.tp_slices == 1) return launch_tcp<32, 3>(a, stream);   // 20 -> 20 channels
    return launch_tcp<32, 0>(a, stream);
  }
  // tp_bn <= 40 (conv_tcp_eligible)
  if (kl == 1) return launch_tcp<48, 1>(a, stream);   // 40 channels = 32 + 8
  if (kl == 2) return launch_tcp<48, 2>(a, stream);   // 80 = 32 + 32 + 16
  if (kl == 4) return launch_tcp<48, 4>(a, stream);   // 160
  return launch_tcp<48, 0>(a, stream);
}

}  // namespace b200ocl
