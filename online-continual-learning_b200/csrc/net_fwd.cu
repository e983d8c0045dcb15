// net_fwd.cu -- Reduced-ResNet18 / SupConResNet forward passes, weight packing, SGD, CE loss.
//
// Replaces model.features / model.forward of reference models/resnet.py:90-109,159-168 in
// eval mode (ASER deep features, utils/utils.py:45-90) and train mode (exp_replay.py:40,62,84;
// scr.py:55; mir_retrieve.py:24-25), torch.optim.SGD.step (setup_elements.py:73-75) and
// F.cross_entropy (agents/base.py:95,113; mir_retrieve.py:26-27).
#include <float.h>
#include <math.h>

#include "net_ws.cuh"
#include "umma.cuh"

namespace b200ocl {
namespace {

// ----------------------------------------------------------------------------- weight packing
struct PackTable {
  int n;
  struct {
    unsigned int w_off, pkf_off, pkd_off;
    int cin, cout, taps;
  } e[NET_MAX_CONV];
};

__global__ void __launch_bounds__(256) pack_kernel(PackTable t, const float* __restrict__ params,
                                                   float* __restrict__ packed) {
  const auto& L = t.e[blockIdx.y];
  const int total = L.cout * L.cin * L.taps;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int co = e / (L.cin * L.taps);
    const int rem = e - co * (L.cin * L.taps);
    const int ci = rem / L.taps, tap = rem - ci * L.taps;
    const float v = params[L.w_off + e];                              // OIHW
    packed[L.pkf_off + (tap * L.cin + ci) * L.cout + co] = v;         // [tap][cin][cout]
    packed[L.pkd_off + (tap * L.cout + co) * L.cin + ci] = v;         // [tap][cout][cin]
  }
}


// Tensor-core operand images: B[n][k] tiles, split into TF32 hi / lo, 128-byte swizzled.
//   forward        n = cout, k = tap*cin + ci            value W[co][ci][tap]
//   data gradient  n = cin,  k = tap*cout + co (flipped)  value W[co][ci][8 - tap]
struct TcPackTable {
  int n;
  struct {
    unsigned int w_off, img_off;
    int cin, cout, bn, nt, kb, dgrad;
  } e[2 * NET_MAX_CONV];
};

__global__ void __launch_bounds__(256) tc_pack_kernel(TcPackTable t, const float* __restrict__ params,
                                                      float* __restrict__ packed) {
  const auto& L = t.e[blockIdx.y];
  const int n_ch = L.dgrad ? L.cin : L.cout;      // output channels of this direction
  const int k_ch = L.dgrad ? L.cout : L.cin;      // channels contracted per tap
  const int ktot = 9 * k_ch;
  const int n_tiles = n_ch / L.bn;
  const int total = n_tiles * L.kb * L.nt * 8;    // (tile, kb, row, 16-byte chunk)
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int c = e & 7;
    int r = e >> 3;
    const int row = r % L.nt;
    r /= L.nt;
    const int kb = r % L.kb, tile = r / L.kb;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const int k0 = kb * 32 + c * 4;
    if (row < L.bn && k0 < ktot) {
      const int tap = k0 / k_ch, kc = k0 - tap * k_ch;   // 4 consecutive k share the tap (k_ch % 4 == 0)
      const int nch = tile * L.bn + row;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = L.dgrad ? kc + j : nch;
        const int ci = L.dgrad ? nch : kc + j;
        const int wt = L.dgrad ? 8 - tap : tap;
        v[j] = params[L.w_off + ((size_t)co * L.cin + ci) * 9 + wt];
      }
    }
    float4 h, l;
    umma::split_tf32(v[0], h.x, l.x); umma::split_tf32(v[1], h.y, l.y);
    umma::split_tf32(v[2], h.z, l.z); umma::split_tf32(v[3], h.w, l.w);
    float* base = packed + L.img_off + ((size_t)(tile * L.kb + kb) * 2) * L.nt * 32;
    const int off = umma::sw128_offset_f32(row, c);
    *reinterpret_cast<float4*>(base + off) = h;
    *reinterpret_cast<float4*>(base + L.nt * 32 + off) = l;
  }
}

// Halo-patch tensor-core images (conv_tcp.cu): blocks [channel tile][slice][tap][hi | lo][NT x 32], row n =
// output channel of the direction, K slot = channel slice*32 + k of the contracted side, zero padded;
// the data-gradient image stores tap t of the correlation = W[.][.][8 - t].
struct TpPackTable {
  int n;
  struct {
    unsigned int w_off, img_off;
    int cin, cout, bn, nt, slices, dgrad, taps;
  } e[2 * NET_MAX_CONV];
};

__global__ void __launch_bounds__(256) tp_pack_kernel(TpPackTable t, const float* __restrict__ params,
                                                      float* __restrict__ packed) {
  const auto& L = t.e[blockIdx.y];
  const int n_ch = L.dgrad ? L.cin : L.cout;
  const int k_ch = L.dgrad ? L.cout : L.cin;
  const int n_tiles = n_ch / L.bn;
  const int total = n_tiles * L.slices * L.taps * L.nt * 8;   // (tile, slice, tap, row, 16-byte chunk)
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int c = e & 7;
    int r = e >> 3;
    const int row = r % L.nt;
    r /= L.nt;
    const int tap = r % L.taps;
    r /= L.taps;
    const int sl = r % L.slices, tile = r / L.slices;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const int k0 = sl * 32 + c * 4;
    if (row < L.bn && k0 < k_ch) {
      const int nch = tile * L.bn + row;
      const int wt = L.dgrad ? L.taps - 1 - tap : tap;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = L.dgrad ? k0 + j : nch;
        const int ci = L.dgrad ? nch : k0 + j;
        v[j] = params[L.w_off + ((size_t)co * L.cin + ci) * L.taps + wt];
      }
    }
    float4 h, l;
    umma::split_tf32(v[0], h.x, l.x); umma::split_tf32(v[1], h.y, l.y);
    umma::split_tf32(v[2], h.z, l.z); umma::split_tf32(v[3], h.w, l.w);
    float* base = packed + L.img_off + ((size_t)((tile * L.slices + sl) * L.taps + tap) * 2) * L.nt * 32;
    const int off = umma::sw128_offset_f32(row, c);
    *reinterpret_cast<float4*>(base + off) = h;
    *reinterpret_cast<float4*>(base + L.nt * 32 + off) = l;
  }
}

int launch_pack(const NetPlan& p, const float* params, float* packed, cudaStream_t stream) {
  PackTable t{};
  t.n = p.n_conv;
  for (int i = 0; i < p.n_conv; ++i) {
    t.e[i].w_off = (unsigned)p.conv[i].w_off;
    t.e[i].pkf_off = (unsigned)p.conv[i].pkf_off;
    t.e[i].pkd_off = (unsigned)p.conv[i].pkd_off;
    t.e[i].cin = p.conv[i].cin;
    t.e[i].cout = p.conv[i].cout;
    t.e[i].taps = p.conv[i].ks * p.conv[i].ks;
  }
  B200OCL_PROF("pack", 12.0 * p.n_packed / 2, stream);
  pack_kernel<<<dim3(16, p.n_conv), 256, 0, stream>>>(t, params, packed);
  B200OCL_LAUNCHED();
  TcPackTable tc{};
  for (int i = 0; i < p.n_conv; ++i) {
    const ConvL& c = p.conv[i];
    if (!c.tc_kb_f) continue;
    for (int d = 0; d < 2; ++d) {
      auto& e = tc.e[tc.n++];
      e.w_off = (unsigned)c.w_off;
      e.img_off = (unsigned)(d ? c.tc_d_off : c.tc_f_off);
      e.cin = c.cin; e.cout = c.cout;
      e.bn = d ? c.tc_bn_d : c.tc_bn_f;
      e.nt = tc_nt(e.bn);
      e.kb = d ? c.tc_kb_d : c.tc_kb_f;
      e.dgrad = d;
    }
  }
  if (tc.n) {
    B200OCL_PROF("pack", 16.0 * p.n_packed / 2, stream);
    tc_pack_kernel<<<dim3(32, tc.n), 256, 0, stream>>>(tc, params, packed);
    B200OCL_LAUNCHED();
  }
  TpPackTable tp{};
  for (int i = 0; i < p.n_conv; ++i) {
    const ConvL& c = p.conv[i];
    for (int d = 0; d < 2; ++d) {
      if (!(d ? c.tp_sl_d : c.tp_sl_f)) continue;
      auto& e = tp.e[tp.n++];
      e.w_off = (unsigned)c.w_off;
      e.img_off = (unsigned)(d ? c.tp_d_off : c.tp_f_off);
      e.cin = c.cin; e.cout = c.cout;
      e.bn = d ? c.tp_bn_d : c.tp_bn_f;
      e.nt = tc_nt(e.bn);
      e.slices = d ? c.tp_sl_d : c.tp_sl_f;
      e.dgrad = d;
      e.taps = c.ks * c.ks;
    }
  }
  if (tp.n) {
    B200OCL_PROF("pack", 16.0 * p.n_packed / 2, stream);
    tp_pack_kernel<<<dim3(32, tp.n), 256, 0, stream>>>(tp, params, packed);
    B200OCL_LAUNCHED();
  }
  return B200OCL_OK;
}

// ----------------------------------------------------------------------------- SGD over the arena
__global__ void __launch_bounds__(256) net_sgd_kernel(const float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ out, size_t n, float lr, float wd,
                                                      size_t skip_lo, size_t skip_hi) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float w = p[i];
    if (i >= skip_lo && i < skip_hi) {  // tensors that never receive a gradient: torch skips them
      out[i] = w;
      continue;
    }
    float gi = g[i];
    if (wd != 0.f) gi = fmaf(wd, w, gi);
    out[i] = w - lr * gi;
  }
}

// ----------------------------------------------------------------------------- train-mode BN apply
struct BnApplyArgs {
  const float* z;
  float* a;
  size_t n_vec;  // float4 count
  int C;
  const float *gamma, *beta, *mean, *invstd;
  const float* res;                            // nullable
  const float *rgamma, *rbeta, *rmean, *rinvstd;  // when res is a raw conv output that needs its own BN
  int relu;
};

__global__ void __launch_bounds__(256) bn_apply_kernel(BnApplyArgs a) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float4* z4 = reinterpret_cast<const float4*>(a.z);
  const float4* r4 = reinterpret_cast<const float4*>(a.res);
  float4* o4 = reinterpret_cast<float4*>(a.a);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_vec; i += stride) {
    const int c = (int)((i * 4) % (size_t)a.C);
    const float4 g = *reinterpret_cast<const float4*>(a.gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(a.beta + c);
    const float4 mu = *reinterpret_cast<const float4*>(a.mean + c);
    const float4 is = *reinterpret_cast<const float4*>(a.invstd + c);
    const float4 x = z4[i];
    float4 v;
    v.x = (x.x - mu.x) * is.x * g.x + b.x;
    v.y = (x.y - mu.y) * is.y * g.y + b.y;
    v.z = (x.z - mu.z) * is.z * g.z + b.z;
    v.w = (x.w - mu.w) * is.w * g.w + b.w;
    if (a.res) {
      float4 r = r4[i];
      if (a.rgamma) {
        const float4 rg = *reinterpret_cast<const float4*>(a.rgamma + c);
        const float4 rb = *reinterpret_cast<const float4*>(a.rbeta + c);
        const float4 rm = *reinterpret_cast<const float4*>(a.rmean + c);
        const float4 ri = *reinterpret_cast<const float4*>(a.rinvstd + c);
        r.x = (r.x - rm.x) * ri.x * rg.x + rb.x;
        r.y = (r.y - rm.y) * ri.y * rg.y + rb.y;
        r.z = (r.z - rm.z) * ri.z * rg.z + rb.z;
        r.w = (r.w - rm.w) * ri.w * rg.w + rb.w;
      }
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (a.relu) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    o4[i] = v;
  }
}

int launch_bn_apply(const BnApplyArgs& a, cudaStream_t stream) {
  size_t blocks = (a.n_vec + 255) / 256;
  const size_t cap = (size_t)16 * sm_count();
  if (blocks > cap) blocks = cap;
  B200OCL_PROF("bn_apply", (a.res ? 48.0 : 32.0) * a.n_vec, stream);
  bn_apply_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

// ----------------------------------------------------------------------------- avg_pool2d(.,4) + flatten
// in NHWC [N,H,W,C] -> feat[n][c*PH*PW + ph*PW + pw]  (NCHW flatten order, resnet.py:97-98)
__global__ void __launch_bounds__(256) pool_kernel(const float* __restrict__ in, float* __restrict__ feat, int N,
                                                   int H, int W, int C, int PH, int PW) {
  const int total = N * PH * PW * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C;
    int t = i / C;
    const int pw = t % PW;
    t /= PW;
    const int ph = t % PH, n = t / PH;
    float s = 0.f;
#pragma unroll
    for (int dy = 0; dy < 4; ++dy)
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) s += in[((size_t)(n * H + ph * 4 + dy) * W + pw * 4 + dx) * C + c];
    feat[(size_t)n * (C * PH * PW) + (c * PH + ph) * PW + pw] = s * 0.0625f;
  }
}

// ----------------------------------------------------------------------------- linear forward
// y[n][o] = b[o] + sum_i x[n][i] * W[o][i]  (+ReLU).  One warp per output feature keeps its weight
// row in registers and walks the batch; in <= 1024.
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                         const float* __restrict__ b, float* __restrict__ y, int N,
                                                         int in, int out, int relu) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int o = blockIdx.x * 8 + warp;
  if (o >= out) return;
  float w[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int i = lane + 32 * j;
    w[j] = (i < in) ? W[(size_t)o * in + i] : 0.f;
  }
  const float bias = b[o];
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    const float* xr = x + (size_t)n * in;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int i = lane + 32 * j;
      if (i < in) s = fmaf(xr[i], w[j], s);
    }
    s = warp_sum(s);
    if (lane == 0) {
      s += bias;
      y[(size_t)n * out + o] = relu ? fmaxf(s, 0.f) : s;
    }
  }
}

int launch_linear_fwd(const float* x, const float* W, const float* b, float* y, int N, int in, int out, int relu,
                      cudaStream_t stream) {
  int gy = N < 32 ? N : 32;
  B200OCL_PROF("head", 4.0 * ((double)N * in + (double)in * out + (double)N * out), stream);
  linear_fwd_kernel<<<dim3((out + 7) / 8, gy), 256, 0, stream>>>(x, W, b, y, N, in, out, relu);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

// ----------------------------------------------------------------------------- F.normalize(dim=1)
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N,
                                                         int d) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  const float* xr = x + (size_t)n * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s = fmaf(xr[i], xr[i], s);
  s = warp_sum(s);
  const float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
  for (int i = lane; i < d; i += 32) y[(size_t)n * d + i] = xr[i] * inv;
}

// ----------------------------------------------------------------------------- cross-entropy
__global__ void __launch_bounds__(256) ce_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                 int N, int C, float* __restrict__ loss, float* __restrict__ per_sample,
                                                 float* __restrict__ dlogits, long long* __restrict__ n_correct) {
  __shared__ float s_loss[8];
  __shared__ int s_corr[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float lsum = 0.f;
  int corr = 0;
  for (int n = warp; n < N; n += 8) {   // fixed assignment of rows to warps: deterministic sum
    const float* lr = logits + (size_t)n * C;
    float mx = -FLT_MAX;
    int arg = 0;
    for (int c = lane; c < C; c += 32) {
      const float v = lr[c];
      if (v > mx) { mx = v; arg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(FULL_MASK, mx, o);
      const int oa = __shfl_xor_sync(FULL_MASK, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float z = 0.f;
    for (int c = lane; c < C; c += 32) z += expf(lr[c] - mx);
    z = warp_sum(z);
    const long long y = labels[n];
    const float lse = mx + logf(z);
    const float l = lse - lr[y];
    if (per_sample && lane == 0) per_sample[n] = l;
    if (dlogits) {
      const float invN = 1.f / (float)N;
      for (int c = lane; c < C; c += 32)
        dlogits[(size_t)n * C + c] = (expf(lr[c] - lse) - (c == y ? 1.f : 0.f)) * invN;
    }
    lsum += l;
    corr += (arg == (int)y) ? 1 : 0;
  }
  if (lane == 0) { s_loss[warp] = lsum; s_corr[warp] = corr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    int k = 0;
    for (int w = 0; w < 8; ++w) { t += s_loss[w]; k += s_corr[w]; }
    if (loss) *loss = t / (float)N;
    if (n_correct) *n_correct = k;
  }
}

// ----------------------------------------------------------------------------- orchestration
void fill_conv_common(ConvArgs& a, const ConvL& c, int N, const float* in, const float* packed, float* out) {
  a = ConvArgs{};
  a.in = in;
  a.w = packed + c.pkf_off;
  a.out = out;
  a.N = N;
  a.Hin = c.hin; a.Win = c.win; a.CK = c.cin;
  a.Hout = c.hout; a.Wout = c.wout; a.CN = c.cout;
  a.ks = c.ks; a.stride = c.stride; a.pad = c.pad;
  a.transposed = 0;
  a.M = N * c.hout * c.wout;
  a.eps = NET_BN_EPS;
  a.momentum = NET_BN_MOMENTUM;
  if (c.tc_kb_f) {
    a.w_tc = packed + c.tc_f_off;
    a.tc_kb = c.tc_kb_f;
    a.tc_bn = c.tc_bn_f;
  }
  if (c.tp_sl_f) {
    a.w_tp = packed + c.tp_f_off;
    a.tp_bn = c.tp_bn_f;
    a.tp_slices = c.tp_sl_f;
  }
}

int conv_eval(const NetPlan& p, const b200ocl_net_state& st, int ci, int N, const float* in, float* out,
              const float* residual, int relu, cudaStream_t stream) {
  ConvArgs a;
  fill_conv_common(a, p.conv[ci], N, in, st.packed, out);
  const BnL& b = p.bn[ci];
  a.mode = CONV_EVAL;
  a.gamma = st.params + b.g_off;
  a.beta = st.params + b.b_off;
  a.rmean = st.bn_stats + b.stat_off;
  a.rvar = st.bn_stats + b.stat_off + b.c;
  a.residual = residual;
  a.relu = relu;
  return ci == 0 ? launch_stem(a, stream) : launch_conv(a, stream);
}


// Diagnostic (B200OCL_RESTAT=1): recompute a layer's batch mean / invstd from the STORED raw output, two passes in
// fp64, one CTA per channel -- isolates "the statistics of the producing kernel" from "the values it stored".
__global__ void __launch_bounds__(256) bn_restat_kernel(const float* __restrict__ z, int M, int C, float eps,
                                                        float* __restrict__ mean_out, float* __restrict__ invstd_out) {
  __shared__ double s_a[256];
  const int c = blockIdx.x, tid = threadIdx.x;
  double s = 0.0;
  for (int m = tid; m < M; m += 256) s += (double)z[(size_t)m * C + c];
  s_a[tid] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s_a[tid] += s_a[tid + o];
    __syncthreads();
  }
  const double mean = s_a[0] / (double)M;
  __syncthreads();
  double q = 0.0;
  for (int m = tid; m < M; m += 256) {
    const double dlt = (double)z[(size_t)m * C + c] - mean;
    q += dlt * dlt;
  }
  s_a[tid] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s_a[tid] += s_a[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    mean_out[c] = (float)mean;
    invstd_out[c] = (float)(1.0 / sqrt(s_a[0] / (double)M + (double)eps));
  }
}

// Eval-statistics forward (GSS-greedy differentiates the network in eval mode, gss_greedy_update.py:16): the
// "saved" statistics that bn_apply and the backward use are the RUNNING ones; the batch statistics the train-mode
// convolution computes go to a throw-away buffer and the running statistics are left alone.
__global__ void bn_save_running_kernel(const float* __restrict__ rmean, const float* __restrict__ rvar, float eps, int C,
                                       float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    save_mean[c] = rmean[c];
    save_invstd[c] = 1.0f / sqrtf(rvar[c] + eps);
  }
}

// stats: where the BatchNorm statistics of a train-mode convolution go.
//   STATS_RUNNING  the reference's train mode: batch statistics normalise, the running statistics move (momentum 0.1)
//   STATS_EVAL     GSS-greedy's differentiable eval-mode pass: the running statistics normalise, nothing moves
//   STATS_DEFER    batch statistics normalise; the running statistics are NOT touched: the (mean, unbiased variance) pair
//                  of every BN is left in the workspace (momentum 1 into a zeroed buffer: 0 * 0 + 1 * s = s exactly) and
//                  applied later, in the caller's order, by b200ocl_net_apply_running_stats -- so that several train-mode
//                  passes of one step (exp_replay.py:40,62,84; scr.py:55) can run concurrently on different streams
enum { STATS_RUNNING = 0, STATS_EVAL = 1, STATS_DEFER = 2 };

int conv_train(const NetPlan& p, const b200ocl_net_state& st, const TrainWs& w, int ci, int N, const float* in,
               cudaStream_t stream, int stats = STATS_RUNNING) {
  const bool eval_stats = stats == STATS_EVAL;
  ConvArgs a;
  const ConvL& c = p.conv[ci];
  fill_conv_common(a, c, N, in, st.packed, w.z + (size_t)N * c.act_off);
  const BnL& b = p.bn[ci];
  a.mode = CONV_TRAIN;
  a.stat_part = w.stat_part;
  a.counter = w.counters + 8 * ci;   // up to 8 channel tiles per conv (160 = 8 x 20)
  a.save_mean = w.save + b.save_off;
  a.save_invstd = w.save + b.save_off + b.c;
  a.run_mean = eval_stats ? w.run_scratch : st.bn_stats + b.stat_off;
  a.run_var = eval_stats ? w.run_scratch + 1024 : st.bn_stats + b.stat_off + b.c;
  if (stats == STATS_DEFER) {
    a.run_mean = w.run_defer + b.stat_off;
    a.run_var = w.run_defer + b.stat_off + b.c;
    a.momentum = 1.0f;
  }
  const int rc = ci == 0 ? launch_stem(a, stream) : launch_conv(a, stream);
  if (rc == 0 && eval_stats) {
    bn_save_running_kernel<<<(b.c + 127) / 128, 128, 0, stream>>>(st.bn_stats + b.stat_off, st.bn_stats + b.stat_off + b.c, a.eps,
                                                                   b.c, a.save_mean, a.save_invstd);
    B200OCL_LAUNCHED();
  }
  static int restat = -1;
  if (restat < 0) {
    const char* e = getenv("B200OCL_RESTAT");
    restat = (e && e[0] == '1') ? 1 : 0;
  }
  if (rc == 0 && restat) {
    bn_restat_kernel<<<c.cout, 256, 0, stream>>>(a.out, N * c.hout * c.wout, c.cout, a.eps, a.save_mean, a.save_invstd);
    B200OCL_LAUNCHED();
  }
  return rc;
}

int bn_apply_train(const NetPlan& p, const b200ocl_net_state& st, const TrainWs& w, int ci, int N, int res_conv,
                   const float* res_plain, cudaStream_t stream) {
  const ConvL& c = p.conv[ci];
  const BnL& b = p.bn[ci];
  BnApplyArgs a{};
  a.z = w.z + (size_t)N * c.act_off;
  a.a = w.a + (size_t)N * c.act_off;
  a.n_vec = (size_t)N * c.hout * c.wout * c.cout / 4;
  a.C = c.cout;
  a.gamma = st.params + b.g_off;
  a.beta = st.params + b.b_off;
  a.mean = w.save + b.save_off;
  a.invstd = w.save + b.save_off + b.c;
  a.relu = 1;
  if (res_conv >= 0) {
    const BnL& rb = p.bn[res_conv];
    a.res = w.z + (size_t)N * p.conv[res_conv].act_off;
    a.rgamma = st.params + rb.g_off;
    a.rbeta = st.params + rb.b_off;
    a.rmean = w.save + rb.save_off;
    a.rinvstd = w.save + rb.save_off + rb.c;
  } else {
    a.res = res_plain;
  }
  return launch_bn_apply(a, stream);
}

int head_forward(const NetPlan& p, const b200ocl_net_state& st, const float* feat, int N, float* hid, float* proj,
                 float* out, cudaStream_t stream) {
  int rc;
  if (p.head == 0) {
    const LinL& l = p.lin[0];
    return launch_linear_fwd(feat, st.params + l.w_off, st.params + l.b_off, out, N, l.in, l.out, 0, stream);
  }
  const float* pre = feat;
  if (p.head == 1) {
    const LinL& l = p.lin[1];
    if ((rc = launch_linear_fwd(feat, st.params + l.w_off, st.params + l.b_off, proj, N, l.in, l.out, 0, stream))) return rc;
    pre = proj;
  } else if (p.head == 2) {
    const LinL& l1 = p.lin[1];
    const LinL& l2 = p.lin[2];
    if ((rc = launch_linear_fwd(feat, st.params + l1.w_off, st.params + l1.b_off, hid, N, l1.in, l1.out, 1, stream))) return rc;
    if ((rc = launch_linear_fwd(hid, st.params + l2.w_off, st.params + l2.b_off, proj, N, l2.in, l2.out, 0, stream))) return rc;
    pre = proj;
  }
  B200OCL_PROF("head", 8.0 * N * p.out_dim, stream);
  l2norm_fwd_kernel<<<(N + 7) / 8, 256, 0, stream>>>(pre, out, N, p.out_dim);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

// running = (1 - momentum) * running + momentum * s for every BN statistic (the expression the convolution epilogues
// use), plus num_batches_tracked += 1: one deferred train-mode pass applied to the running statistics
__global__ void bn_running_apply_kernel(float* __restrict__ run, const float* __restrict__ s, int n, float momentum,
                                        long long* __restrict__ tracked, int n_bn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) run[i] = (1.f - momentum) * run[i] + momentum * s[i];
  if (tracked && i < n_bn) tracked[i] += 1;
}

__global__ void bump_tracked_kernel(long long* t, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] += 1;
}

int check_state(const b200ocl_net_desc* desc, const b200ocl_net_state* st, NetPlan& p) {
  if (!desc || !st || !st->params || !st->packed || !st->bn_stats) {
    set_error("net: null descriptor/state pointer");
    return B200OCL_EINVAL;
  }
  const int rc = build_plan(*desc, p);
  if (rc) set_error("net: unsupported network description (nf must be 20, head in 0..3, dims <= 1024)");
  return rc;
}

}  // namespace
}  // namespace b200ocl

extern "C" {

int b200ocl_net_query(const b200ocl_net_desc* desc, b200ocl_net_info* info) {
  using namespace b200ocl;
  B200OCL_CHECK_ARG(desc && info, "null pointer");
  NetPlan p;
  const int rc = build_plan(*desc, p);
  if (rc) {
    set_error("b200ocl_net_query: unsupported network description");
    return rc;
  }
  info->n_params = p.n_params;
  info->n_packed = p.n_packed;
  info->n_bn_stats = p.n_stats;
  info->n_bn = p.n_conv;
  info->n_tensors = 3 * p.n_conv + 2 * p.n_lin;
  info->dim_in = p.dim_in;
  info->out_dim = p.out_dim;
  return B200OCL_OK;
}

int b200ocl_net_tensor(const b200ocl_net_desc* desc, int i, size_t* offset, size_t* numel, int* has_grad) {
  using namespace b200ocl;
  B200OCL_CHECK_ARG(desc && offset && numel && has_grad, "null pointer");
  NetPlan p;
  const int rc = build_plan(*desc, p);
  if (rc) return rc;
  const int n_conv_t = 3 * p.n_conv;
  *has_grad = 1;
  if (i < 0 || i >= n_conv_t + 2 * p.n_lin) {
    set_error("b200ocl_net_tensor: index out of range");
    return B200OCL_EINVAL;
  }
  if (i < n_conv_t) {
    const int c = i / 3, k = i % 3;
    if (k == 0) { *offset = p.conv[c].w_off; *numel = (size_t)p.conv[c].cout * p.conv[c].cin * p.conv[c].ks * p.conv[c].ks; }
    else if (k == 1) { *offset = p.bn[c].g_off; *numel = p.bn[c].c; }
    else { *offset = p.bn[c].b_off; *numel = p.bn[c].c; }
    return B200OCL_OK;
  }
  const int l = (i - n_conv_t) / 2, k = (i - n_conv_t) % 2;
  if (k == 0) { *offset = p.lin[l].w_off; *numel = (size_t)p.lin[l].in * p.lin[l].out; }
  else { *offset = p.lin[l].b_off; *numel = p.lin[l].out; }
  if (p.head != 0 && l == 0) *has_grad = 0;
  return B200OCL_OK;
}

int b200ocl_net_pack(const b200ocl_net_desc* desc, const b200ocl_net_state* st, void* stream) {
  using namespace b200ocl;
  NetPlan p;
  int rc = check_state(desc, st, p);
  if (rc) return rc;
  return launch_pack(p, st->params, st->packed, static_cast<cudaStream_t>(stream));
}

static void selftest_layer(b200ocl::ConvL& c, int cin, int cout, int H, int W, int ks, int stride, size_t& pk) {
  c = b200ocl::ConvL{};
  c.cin = cin; c.cout = cout; c.ks = ks; c.stride = stride; c.pad = (ks == 3) ? 1 : 0;
  c.hin = H; c.win = W;
  c.hout = b200ocl::conv_out(H, ks, stride, c.pad);
  c.wout = b200ocl::conv_out(W, ks, stride, c.pad);
  c.w_off = 0;
  pk = 0;
  b200ocl::conv_pack_layout(c, pk, true);
}

size_t b200ocl_conv_selftest_workspace_bytes(int N, int cin, int cout, int H, int W, int ks, int stride) {
  b200ocl::ConvL c;
  size_t pk = 0;
  selftest_layer(c, cin, cout, H, W, ks, stride, pk);
  const size_t M = (size_t)N * H * W;   // upper bound (stride 1)
  const size_t stat = (size_t)b200ocl::conv_max_grid_m((int)M) * (cin > cout ? cin : cout) * 2 * sizeof(double);
  return b200ocl::align_up(pk * sizeof(float), 256) + b200ocl::align_up(stat, 256) + 256 /* counters */ + 256;
}

int b200ocl_conv_selftest(const float* x, const float* w_oihw, float* out, int N, int H, int W, int cin, int cout,
                          int ks, int stride, int dgrad, int path, int mode, float* stats_out, void* workspace,
                          size_t workspace_bytes, void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200OCL_CHECK_ARG(x && w_oihw && out && workspace, "null pointer");
  B200OCL_CHECK_ARG(N > 0 && H > 0 && W > 0 && cin % 20 == 0 && cout % 20 == 0 && cin > 0 && cout > 0, "bad shape");
  B200OCL_CHECK_ARG(mode >= 0 && mode <= 2 && (mode != 2 || (stats_out && !dgrad)), "mode 2 (train) needs stats_out, forward only");
  B200OCL_CHECK_ARG((ks == 3 || ks == 1) && (stride == 1 || stride == 2) && (!dgrad || (ks == 3 && stride == 1)),
                    "3x3 or 1x1, stride 1 or 2; data gradient for 3x3 stride 1 only");
  B200OCL_CHECK_ARG(workspace_bytes >= b200ocl_conv_selftest_workspace_bytes(N, cin, cout, H, W, ks, stride), "workspace too small");
  NetPlan p;
  memset(&p, 0, sizeof(p));
  p.n_conv = 1;
  size_t pk = 0;
  selftest_layer(p.conv[0], cin, cout, H, W, ks, stride, pk);
  p.n_packed = pk;
  unsigned char* base = static_cast<unsigned char*>(workspace);
  float* packed = reinterpret_cast<float*>(base);
  double* stat_part = reinterpret_cast<double*>(base + align_up(pk * sizeof(float), 256));
  const size_t M = (size_t)N * H * W;
  const size_t stat = (size_t)conv_max_grid_m((int)M) * (cin > cout ? cin : cout) * 2 * sizeof(double);
  unsigned int* counters = reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(stat_part) + align_up(stat, 256));
  int rc = launch_pack(p, w_oihw, packed, stream);
  if (rc) return rc;
  const ConvL& c = p.conv[0];
  ConvArgs a{};
  a.in = x;
  a.out = out;
  a.N = N;
  a.Hin = H; a.Win = W;
  a.Hout = c.hout; a.Wout = c.wout;
  a.ks = ks; a.stride = stride; a.pad = c.pad;
  a.M = N * c.hout * c.wout;
  a.mode = mode == 2 ? CONV_TRAIN : (mode == 1 ? CONV_ACCUM : CONV_RAW);
  a.force_path = path;
  a.eps = NET_BN_EPS;
  a.momentum = NET_BN_MOMENTUM;
  if (mode == 2) {
    B200OCL_CUDA(cudaMemsetAsync(counters, 0, 64, stream));
    B200OCL_CUDA(cudaMemsetAsync(stats_out, 0, (size_t)4 * cout * sizeof(float), stream));
    a.stat_part = stat_part;
    a.counter = counters;
    a.save_mean = stats_out;
    a.save_invstd = stats_out + cout;
    a.run_mean = stats_out + 2 * cout;
    a.run_var = stats_out + 3 * cout;
  }
  if (!dgrad) {
    a.CK = cin; a.CN = cout;
    a.w = packed + c.pkf_off;
    if (c.tc_kb_f) { a.w_tc = packed + c.tc_f_off; a.tc_kb = c.tc_kb_f; a.tc_bn = c.tc_bn_f; }
    if (c.tp_sl_f) { a.w_tp = packed + c.tp_f_off; a.tp_bn = c.tp_bn_f; a.tp_slices = c.tp_sl_f; }
  } else {
    a.CK = cout; a.CN = cin;
    a.w = packed + c.pkd_off;
    a.flip = 1;
    if (c.tc_kb_d) { a.w_tc = packed + c.tc_d_off; a.tc_kb = c.tc_kb_d; a.tc_bn = c.tc_bn_d; }
    if (c.tp_sl_d) { a.w_tp = packed + c.tp_d_off; a.tp_bn = c.tp_bn_d; a.tp_slices = c.tp_sl_d; }
  }
  return launch_conv(a, stream);
}

int b200ocl_net_sgd_step(const b200ocl_net_desc* desc, const b200ocl_net_state* st, float lr, float weight_decay,
                         const b200ocl_net_state* dst, void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NetPlan p;
  int rc = check_state(desc, st, p);
  if (rc) return rc;
  B200OCL_CHECK_ARG(st->grads, "state has no gradient arena");
  float* out_params = dst ? dst->params : st->params;
  float* out_packed = dst ? dst->packed : st->packed;
  B200OCL_CHECK_ARG(out_params && out_packed, "destination state incomplete");
  size_t skip_lo = 0, skip_hi = 0;
  if (p.head != 0) {
    skip_lo = p.lin[0].w_off;
    skip_hi = p.lin[0].b_off + p.lin[0].out;
  }
  size_t blocks = (p.n_params + 255) / 256;
  const size_t cap = (size_t)8 * sm_count();
  if (blocks > cap) blocks = cap;
  B200OCL_PROF("sgd", 12.0 * p.n_params, stream);
  net_sgd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(st->params, st->grads, out_params, p.n_params, lr, weight_decay,
                                                       skip_lo, skip_hi);
  B200OCL_LAUNCHED();
  return launch_pack(p, out_params, out_packed, stream);
}

size_t b200ocl_net_eval_workspace_bytes(const b200ocl_net_desc* desc, int N) {
  using namespace b200ocl;
  NetPlan p;
  if (!desc || N < 0 || build_plan(*desc, p)) return 0;
  return eval_ws(p, N > 0 ? N : 1, nullptr).bytes;
}

int b200ocl_net_features_eval(const b200ocl_net_desc* desc, const b200ocl_net_state* st, const float* x, int N,
                              float* feat, void* workspace, size_t workspace_bytes, void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NetPlan p;
  int rc = check_state(desc, st, p);
  if (rc) return rc;
  B200OCL_CHECK_ARG(N >= 0, "negative batch");
  if (N == 0) return B200OCL_OK;
  B200OCL_CHECK_ARG(x && feat, "null pointer");
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255) ||
      workspace_bytes < b200ocl_net_eval_workspace_bytes(desc, N)) {
    set_error("b200ocl_net_features_eval: workspace missing, misaligned or too small");
    return B200OCL_EWORKSPACE;
  }
  EvalWs w = eval_ws(p, N, workspace);
  float* cur = w.buf[0];
  float* t1 = w.buf[1];
  float* t2 = w.buf[2];
  float* nxt = w.buf[3];
  if ((rc = conv_eval(p, *st, 0, N, x, cur, nullptr, 1, stream))) return rc;   // relu(bn1(conv1(x)))
  for (int b = 0; b < 8; ++b) {
    const BlockL& B = p.blk[b];
    if ((rc = conv_eval(p, *st, B.c1, N, cur, t1, nullptr, 1, stream))) return rc;
    const float* res = cur;
    if (B.sc >= 0) {
      if ((rc = conv_eval(p, *st, B.sc, N, cur, t2, nullptr, 0, stream))) return rc;
      res = t2;
    }
    if ((rc = conv_eval(p, *st, B.c2, N, t1, nxt, res, 1, stream))) return rc;
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  const int total = N * p.pooled_h * p.pooled_w * (desc->nf * 8);
  B200OCL_PROF("pool", 4.0 * 17 * total, stream);
  pool_kernel<<<(total + 255) / 256, 256, 0, stream>>>(cur, feat, N, p.final_h, p.final_w, desc->nf * 8, p.pooled_h,
                                                       p.pooled_w);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

size_t b200ocl_net_train_workspace_bytes(const b200ocl_net_desc* desc, int N) {
  using namespace b200ocl;
  NetPlan p;
  if (!desc || N < 0 || build_plan(*desc, p)) return 0;
  return train_ws(p, N > 0 ? N : 1, nullptr, sm_count()).bytes;
}

static int net_forward_impl(const b200ocl_net_desc* desc, const b200ocl_net_state* st, const float* x, int N, float* out,
                            void* workspace, size_t workspace_bytes, void* stream_, int ev) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NetPlan p;
  int rc = check_state(desc, st, p);
  if (rc) return rc;
  B200OCL_CHECK_ARG(N >= 1 && x && out, "need N >= 1 and non-null x/out");
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255) ||
      workspace_bytes < b200ocl_net_train_workspace_bytes(desc, N)) {
    set_error("b200ocl_net_forward_train: workspace missing, misaligned or too small");
    return B200OCL_EWORKSPACE;
  }
  TrainWs w = train_ws(p, N, workspace, sm_count());
  B200OCL_CUDA(cudaMemsetAsync(w.counters, 0, NET_COUNTERS * sizeof(unsigned int), stream));
  if (ev == STATS_DEFER) B200OCL_CUDA(cudaMemsetAsync(w.run_defer, 0, p.n_stats * sizeof(float), stream));
  if ((rc = conv_train(p, *st, w, 0, N, x, stream, ev))) return rc;
  if ((rc = bn_apply_train(p, *st, w, 0, N, -1, nullptr, stream))) return rc;
  const float* cur = w.a + (size_t)N * p.conv[0].act_off;
  for (int b = 0; b < 8; ++b) {
    const BlockL& B = p.blk[b];
    if ((rc = conv_train(p, *st, w, B.c1, N, cur, stream, ev))) return rc;
    if ((rc = bn_apply_train(p, *st, w, B.c1, N, -1, nullptr, stream))) return rc;
    const float* a1 = w.a + (size_t)N * p.conv[B.c1].act_off;
    if ((rc = conv_train(p, *st, w, B.c2, N, a1, stream, ev))) return rc;
    if (B.sc >= 0) {
      if ((rc = conv_train(p, *st, w, B.sc, N, cur, stream, ev))) return rc;
      if ((rc = bn_apply_train(p, *st, w, B.c2, N, B.sc, nullptr, stream))) return rc;
    } else {
      if ((rc = bn_apply_train(p, *st, w, B.c2, N, -1, cur, stream))) return rc;
    }
    cur = w.a + (size_t)N * p.conv[B.c2].act_off;
  }
  if (st->bn_tracked && ev == STATS_RUNNING) {
    B200OCL_PROF("misc", 16.0 * p.n_conv, stream);
    bump_tracked_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<long long*>(st->bn_tracked), p.n_conv);
    B200OCL_LAUNCHED();
  }
  const int total = N * p.pooled_h * p.pooled_w * (desc->nf * 8);
  B200OCL_PROF("pool", 4.0 * 17 * total, stream);
  pool_kernel<<<(total + 255) / 256, 256, 0, stream>>>(cur, w.feat, N, p.final_h, p.final_w, desc->nf * 8, p.pooled_h,
                                                       p.pooled_w);
  B200OCL_LAUNCHED();
  return head_forward(p, *st, w.feat, N, w.hid, w.proj, out, stream);
}

int b200ocl_net_forward_train(const b200ocl_net_desc* desc, const b200ocl_net_state* st, const float* x, int N,
                              float* out, void* workspace, size_t workspace_bytes, void* stream_) {
  return net_forward_impl(desc, st, x, N, out, workspace, workspace_bytes, stream_, b200ocl::STATS_RUNNING);
}

int b200ocl_net_forward_evalgrad(const b200ocl_net_desc* desc, const b200ocl_net_state* st, const float* x, int N,
                                 float* out, void* workspace, size_t workspace_bytes, void* stream_) {
  return net_forward_impl(desc, st, x, N, out, workspace, workspace_bytes, stream_, b200ocl::STATS_EVAL);
}

int b200ocl_net_forward_train_deferred(const b200ocl_net_desc* desc, const b200ocl_net_state* st, const float* x, int N,
                                       float* out, void* workspace, size_t workspace_bytes, void* stream_) {
  return net_forward_impl(desc, st, x, N, out, workspace, workspace_bytes, stream_, b200ocl::STATS_DEFER);
}

int b200ocl_net_apply_running_stats(const b200ocl_net_desc* desc, const b200ocl_net_state* st, int N, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NetPlan p;
  int rc = check_state(desc, st, p);
  if (rc) return rc;
  B200OCL_CHECK_ARG(N >= 1, "need N >= 1");
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255) ||
      workspace_bytes < b200ocl_net_train_workspace_bytes(desc, N)) {
    set_error("b200ocl_net_apply_running_stats: workspace missing, misaligned or too small");
    return B200OCL_EWORKSPACE;
  }
  TrainWs w = train_ws(p, N, workspace, sm_count());
  const int n = (int)p.n_stats;
  B200OCL_PROF("misc", 12.0 * n, stream);
  bn_running_apply_kernel<<<(n + 255) / 256, 256, 0, stream>>>(st->bn_stats, w.run_defer, n, NET_BN_MOMENTUM,
                                                               reinterpret_cast<long long*>(st->bn_tracked), p.n_conv);
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

int b200ocl_ce_loss(const float* logits, const int64_t* labels, int N, int C, float* loss, float* per_sample,
                    float* dlogits, int64_t* n_correct, void* stream_) {
  using namespace b200ocl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200OCL_CHECK_ARG(logits && labels && N >= 1 && C >= 1, "need logits, labels, N >= 1, C >= 1");
  B200OCL_PROF("ce_loss", 8.0 * N * C, stream);
  ce_kernel<<<1, 256, 0, stream>>>(logits, reinterpret_cast<const long long*>(labels), N, C, loss, per_sample, dlogits,
                                   reinterpret_cast<long long*>(n_correct));
  B200OCL_LAUNCHED();
  return B200OCL_OK;
}

}  // extern "C"
