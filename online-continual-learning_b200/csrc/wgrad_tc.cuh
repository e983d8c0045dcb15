// wgrad_tc.cuh -- weight gradient of a 3x3 stride-1 convolution on tcgen05 (wgrad_tc.cu): configuration shared by the
// workspace sizing (net_ws.cuh), the launcher and the backward driver (net_bwd.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

namespace b200ocl {

struct WgradTcCfg {
  int eligible;        // geometry covered (3x3, stride 1, pad 1, W <= 37, channels % 4 == 0)
  int slices;          // ceil(cin / 32): one CTA column per 32-channel slice of the activation
  int cout_blocks;     // ceil(cout / 32): one CTA layer per 32-channel block of the output gradient
  int tiles;           // 128-position tiles of the zero-padded strip
  int tpc;             // tiles accumulated in TMEM before the sum is written out as one partial ("chain")
  int chains;          // = partials per (slice, block): ceil(tiles / tpc) -- the `splits` the finalize kernel sums
  int chains_per_cta;
  int ctas_x;
};

inline int wgrad_tc_tiles(int N, int H, int W) {
  const long pp = (long)(H + 2) * (W + 2);
  const long last = (long)(N - 1) * pp + (long)(H - 1) * (W + 2) + (W - 1);
  return (int)(last / 128) + 1;
}

// Which weight gradients take the tensor-core kernel: B200OCL_WGRAD_TC = 0 none, 1 every covered layer (default).
inline int wgrad_tc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("B200OCL_WGRAD_TC");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on;
}

inline WgradTcCfg wgrad_tc_cfg(int N, int H, int W, int ks, int stride, int pad, int cin, int cout, int sms) {
  WgradTcCfg g{};
  g.eligible = (ks == 3 && stride == 1 && pad == 1 && cin % 4 == 0 && cout % 4 == 0 && cin >= 4 && cout >= 4 &&
                128 + 2 * (W + 2) + 2 <= 16 * 13 && (long)N * (H + 2) * (W + 2) < 2000000000L) ? 1 : 0;
  if (!g.eligible) return g;
  g.slices = (cin + 31) / 32;
  g.cout_blocks = (cout + 31) / 32;
  g.tiles = wgrad_tc_tiles(N, H, W);
  static int tpc = 0;      // B200OCL_WGRAD_TPC: tiles per TMEM accumulation chain (default 2 = 256 positions, ~2e-6 relative)
  if (!tpc) {
    const char* e = getenv("B200OCL_WGRAD_TPC");
    tpc = (e && atoi(e) > 0 && atoi(e) <= 16) ? atoi(e) : 2;
  }
  g.tpc = tpc;
  g.chains = (g.tiles + g.tpc - 1) / g.tpc;
  int want = sms / (g.slices * g.cout_blocks);
  if (want < 1) want = 1;
  if (want > g.chains) want = g.chains;
  g.chains_per_cta = (g.chains + want - 1) / want;
  g.ctas_x = (g.chains + g.chains_per_cta - 1) / g.chains_per_cta;
  return g;
}

struct WgradTcArgs {
  const float* x;    // NHWC [N,H,W,Cin]   the convolution's input activation
  const float* dz;   // NHWC [N,H,W,Cout]  gradient of its raw output
  float* part;       // [chains][9 * Cin][Cout] partial sums, k = (kh * 3 + kw) * Cin + ci
  int N, H, W, Cin, Cout;
  int tpc, chains, chains_per_cta;
};

int launch_wgrad_tc(const WgradTcArgs& a, const WgradTcCfg& g, cudaStream_t stream);

}  // namespace b200ocl
