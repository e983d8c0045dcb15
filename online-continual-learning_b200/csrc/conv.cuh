// conv.cuh -- argument block and launchers of the implicit-GEMM convolution kernels.
#pragma once
#include "common.cuh"

namespace b200ocl {

enum ConvMode {
  CONV_RAW = 0,    // out = acc
  CONV_EVAL = 1,   // out = relu?((acc - rmean) * gamma/sqrt(rvar+eps) + beta (+ residual))   (folded eval-mode BN)
  CONV_TRAIN = 2,  // out = acc, plus per-channel batch statistics (sum, sum of squares) -> mean/invstd,
                   // running-stat update by the last CTA of each channel tile
  CONV_ACCUM = 3   // out += acc (data-gradient merged into an existing gradient)
};

struct ConvArgs {
  const float* in;   // NHWC [N,Hin,Win,CK]   (stem: NCHW [N,3,Hin,Win])
  const float* w;    // packed [ks*ks][CK][CN]
  float* out;        // NHWC [N,Hout,Wout,CN]
  int N, Hin, Win, CK, Hout, Wout, CN;
  int ks, stride, pad;
  int transposed;    // 0: forward gather (hi = ho*stride - pad + kh); 1: data-gradient gather
                     //    (input pixel (t/stride) with t = ho + pad - kh, only when divisible)
  const float* w_tc; // tensor-core operand image of the same weights (net_plan.cuh), nullable
  int tc_kb, tc_bn;  // its K blocks and real channels per tile
  const float* w_tp; // halo-patch tensor-core image (conv_tcp.cu), nullable
  int tp_bn, tp_slices;
  int tp_ps, tp_bs;  // set by the launcher: patch stages, weight ring depth
  int tp_chain;      // taps accumulated in TMEM before a promotion: 1 (default) or 3 (B200OCL_TCP_CHAIN=3)
  float tp_debias;   // multiplier of the experimental TMEM-accumulate rounding compensation (default 0; conv_tcp.cu tmem_accumulate)
  int force_path;    // 0 automatic; 1 CUDA-core kernels only; 2 conv_tc; 3 conv_tcp (selftest: fails if not eligible)
  int parity_order;  // stride-2 data gradient only: pixels enumerated [parity class][n][h/2][w/2] so that a CTA
                     //    sees one class and skips the taps that cannot reach it (9 of 36 tap-pixel pairs are live)
  int flip;          // patch kernel only: use tap (ks*ks-1-tap) of the weights (stride-1 data gradient)
  int th, tw, ti;    // patch kernel only: spatial tile (rows, cols, images), set by the launcher
  int M;             // N*Hout*Wout output pixels
  int mode;
  // CONV_EVAL
  const float* gamma;
  const float* beta;
  const float* rmean;
  const float* rvar;
  const float* residual;  // NHWC like out, nullable
  int relu;
  float eps;
  // CONV_TRAIN
  double* stat_part;       // [gridDim.x][CN][2]
  unsigned int* counter;   // [gridDim.y], zeroed before the forward pass
  float* save_mean;
  float* save_invstd;
  float* run_mean;
  float* run_var;
  float momentum;
};

// Upper bound of gridDim.x over every tiling launch_conv may choose (sizes stat_part).
int conv_max_grid_m(int M);
int launch_conv(const ConvArgs& a, cudaStream_t stream);   // CK, CN multiples of 20
int launch_stem(const ConvArgs& a, cudaStream_t stream);   // CK == 3, CN == 20, ks == 3, NCHW input
int launch_conv_tc(const ConvArgs& a, cudaStream_t stream);  // conv_tc.cu: tcgen05 3xTF32 path
bool conv_tc_eligible(const ConvArgs& a);
int launch_conv_tcp(const ConvArgs& a, cudaStream_t stream);  // conv_tcp.cu: tcgen05 fed from a halo patch
bool conv_tcp_eligible(const ConvArgs& a);
bool conv_tcp_mode_allowed(const ConvArgs& a);   // launch kinds (eval / train / data gradient) the policy sends to conv_tcp

}  // namespace b200ocl
