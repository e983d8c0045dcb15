/*
 * b200ocl.h -- C ABI of libb200ocl.so: the replay-step hot path of
 * RaptorMai/online-continual-learning as hand-written sm_100a CUDA.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own (SURVEY.md
 * section 8b); each entry point below therefore names the reference *Python*
 * interface whose arithmetic it replaces (file:line under /root/reference).
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C: pointers, sizes, a cudaStream_t passed as void*; no torch types;
 *   - every pointer is a BORROWED DEVICE pointer, contiguous row-major, owned by
 *     the caller and kept alive until the stream work completes;
 *   - fp32 data, int64 labels/indices (the reference's dtypes, data_utils.py:41,
 *     buffer.py:23);
 *   - no hidden allocation: scratch is caller-provided, sized by the matching
 *     *_workspace_bytes() query (must be 256-byte aligned); kernel attributes are cached per device;
 *   - every function returns 0 on success or a B200OCL_E* code; the message is
 *     available from b200ocl_last_error() (thread-local); nothing throws;
 *   - launches are asynchronous on `stream`; no function synchronises.
 */
#ifndef B200OCL_H_
#define B200OCL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200OCL_OK 0
#define B200OCL_EINVAL 1      /* bad argument (null pointer, negative size, ...)   */
#define B200OCL_EUNSUPPORTED 2 /* shape outside what the kernels cover            */
#define B200OCL_EWORKSPACE 3  /* workspace too small or misaligned               */
#define B200OCL_ECUDA 4       /* a CUDA runtime call failed                      */

#define B200OCL_KNN_MAX_CAND 1024 /* candidates the fused (register-sorted) kNN-SV kernel takes */
#define B200OCL_KNN_MAX_CAND_LARGE 262144 /* beyond that: the scratch-line kernel (knn_sv_large.cu), same results */

const char* b200ocl_last_error(void);
int b200ocl_version(void);
/* Number of kernel launches issued through this library since load (bench.py reports it). */
uint64_t b200ocl_launch_count(void);
/* Optional per-launch device timing for bench.py's roofline: between begin and end every launch of
 * this library is bracketed by CUDA events on its launch stream and accumulated per kernel class
 * together with its algorithmic work (FLOPs for conv/wgrad classes, bytes otherwise).
 * profile_end synchronises the device and returns the number of classes. */
void b200ocl_profile_begin(void);
int b200ocl_profile_end(void);
int b200ocl_profile_get(int k, char* name, int name_len, double* ms, int* launches, double* work);

/* ---------------------------------------------------------------- kNN Shapley values
 * Replaces compute_knn_sv minus its network forward: sorted_cand_ind +
 * euclidean_distance + the Shapley recurrence + scatter (utils/buffer/aser_utils.py:29-59,
 * 94-116; utils/utils.py:93-95) and the row reductions its callers take
 * (aser_retrieve.py:79,82,86; aser_update.py:80), in ONE kernel.
 *   eval_f [E,d] f32, eval_y [E] i64, cand_f [C,d] f32, cand_y [C] i64, k neighbours.
 * Outputs (each nullable): sv [E,C] Shapley matrix in candidate order; col_sum /
 * col_max / col_min [C] reductions over eval rows.  Distance is the squared L2 in
 * direct-difference form; equal distances rank lowest candidate index first.
 * Reductions are deterministic (fixed-order, no float atomics).  C <= 262144 (one fused launch up to 1024
 * candidates, the scratch-line kernel beyond; the workspace query covers both). */
size_t b200ocl_knn_sv_workspace_bytes(int E, int C, int d);
int b200ocl_knn_sv(const float* eval_f, const int64_t* eval_y, const float* cand_f, const int64_t* cand_y,
                   int E, int C, int d, int k,
                   float* sv, float* col_sum, float* col_max, float* col_min,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- ranking
 * score[i] = a[i]*sa + (b ? b[i]*sb : 0); idx_out[0..n_out) = positions of the n_out
 * largest scores in descending order, ties lowest index first (the reference's
 * sv.argsort(descending=True)[:n], aser_retrieve.py:82-91, aser_update.py:80-93;
 * MIR's scores.sort(descending=True)[1][:n], mir_retrieve.py:28-29).  n <= 4096.
 * score_out [n] nullable. */
int b200ocl_rank_desc(const float* a, float sa, const float* b, float sb, int n,
                      int64_t* idx_out, int n_out, float* score_out, void* stream);

/* ---------------------------------------------------------------- SupCon loss
 * Replaces SupConLoss.forward (+ its autograd backward), all-views-anchor mode
 * (utils/loss.py:19-96): features [B,V,d] f32, labels [B] i64 -> loss[1] and, when
 * dfeats != NULL, dL/dfeatures [B,V,d].  d <= 1024. */
size_t b200ocl_supcon_workspace_bytes(int B, int V, int d);
int b200ocl_supcon(const float* feats, const int64_t* labels, int B, int V, int d, float temperature,
                   float* loss, float* dfeats, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- buffer rows
 * dst[i,:] = src[idx[i],:]  /  dst[idx[i],:] = src[i,:]   rows of row_bytes bytes
 * (buffer_img[indices], buffer_img[idx] = x: buffer_utils.py:19-21,115-116;
 * reservoir_update.py:59-60; aser_update.py:111-112).  row_bytes % 4 == 0.
 * Scatter with duplicate idx is undefined, as in the reference's index_put. */
int b200ocl_gather_rows(const void* src, const int64_t* idx, int n_rows, size_t row_bytes, void* dst, void* stream);
int b200ocl_scatter_rows(const void* src, const int64_t* idx, int n_rows, size_t row_bytes, void* dst, void* stream);

/* ---------------------------------------------------------------- evaluate() (agents/base.py:118-175)
 * Nearest-class-mean classification on encoder features.  class_means: for each class_ids[k] the normalised mean
 * of the normalised features of the samples with that label (base.py:121-141); counts[k] == 0 leaves means[k]
 * untouched (the reference draws a random vector there).  classify: pred[b] = class_ids[arg-min_k ||f_b/||f_b|| -
 * means[k]||^2] (first minimum), *n_correct += #(pred == truth) (truth / pred / n_correct nullable).
 * linear_argmax: pred[b] = arg-max_c (feats[b] . weight[c] + bias[c]) -- the classifier branch (base.py:172-175). */
int b200ocl_ncm_class_means(const float* feats, const int64_t* labels, int n, int d, const int64_t* class_ids, int K,
                            float* means, int* counts, void* stream);
int b200ocl_ncm_classify(const float* feats, int B, int d, const float* means, int K, const int64_t* class_ids,
                         const int64_t* truth, int64_t* pred, uint64_t* n_correct, void* stream);
int b200ocl_linear_argmax(const float* feats, int B, int d, const float* weight, const float* bias, int C,
                          const int64_t* truth, int64_t* pred, uint64_t* n_correct, void* stream);

/* A-GEM gradient projection (agents/agem.py:60-80) on flat gradient arenas: out = g - (g.g_ref / g_ref.g_ref) g_ref if
 * g.g_ref < 0, else out = g (out may alias g_ref or g).  dots_out (nullable, [2] f32) receives g.g_ref and g_ref.g_ref.
 * Deterministic: fp64 partials reduced in CTA order. */
size_t b200ocl_agem_project_workspace_bytes(void);
int b200ocl_agem_project(const float* g, const float* g_ref, float* out, size_t n, float* dots_out, void* workspace,
                         size_t workspace_bytes, void* stream);

/* GSS-greedy scores (utils/buffer/gss_greedy_update.py:84,120 with cosine_similarity of buffer_utils.py:50-55): cosine
 * similarity of the flat gradient g [n] with each of the K <= 64 stored gradients mem_grads [K,n] (cos_out [K], nullable)
 * and their maximum (max_out [1], nullable).  Deterministic fp64 partials. */
size_t b200ocl_grad_cosine_workspace_bytes(int K);
int b200ocl_grad_cosine(const float* mem_grads, const float* g, int K, size_t n, float* cos_out, float* max_out, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Stream feeder (continuum/data_utils.py:38-54: ToTensor on every sample + DataLoader shuffle): dst[i] = image
 * src[perm[i]] converted uint8 HWC -> fp32 CHW in [0,1] with an IEEE division by 255 (bit-identical to the
 * reference's CPU ToTensor).  perm may be NULL (identity).  h*w*3 % 4 == 0. */
int b200ocl_stream_prepare(const uint8_t* src_hwc, const int64_t* perm, int n, int h, int w, float* dst_chw, void* stream);

/* ASER memory replacement on the device (reference utils/buffer/aser_update.py:88-112: current samples ranked
 * inside the first n_cand_buf places of the descending SV ranking replace, pairwise in rank order, the buffered
 * candidates ranked below).  order[n_total] ranks positions of [buffered candidates | current batch];
 * cand_slot[n_cand_buf] are the candidates' buffer slots.  Moves image rows and labels; pairs_out (nullable,
 * [1 + 2*n_cur] int64 = count, src positions, dst slots, -1 padded) lets the host mirror follow asynchronously. */
int b200ocl_aser_replace(const int64_t* order, int n_total, int n_cand_buf, const int64_t* cand_slot, const void* cur_x,
                         const int64_t* cur_y, int n_cur, size_t row_bytes, void* buffer_img, int64_t* buffer_label,
                         int64_t* pairs_out, void* stream);

/* ---------------------------------------------------------------- SGD
 * p -= lr * (g + wd * p) over a flat arena of n floats: torch.optim.SGD without
 * momentum (utils/setup_elements.py:73-75) and MIR's virtual step
 * theta' = theta - lr*g when out != p (mir_retrieve.py:43-46).  out may alias p. */
int b200ocl_sgd_step(const float* p, const float* g, float* out, size_t n, float lr, float wd, void* stream);


/* ---------------------------------------------------------------- Reduced-ResNet18 engine
 * The only network on the replay-step path (models/resnet.py:14-37,69-116 BasicBlock ResNet
 * with nf=20; :140-168 SupConResNet; per-dataset classifier, utils/setup_elements.py:46-68).
 *   head: 0 classifier (logits [N,num_classes]); 1/2/3 SupConResNet with 'linear'/'mlp'/'None'
 *   head (L2-normalised projection [N,feat_dim]).
 * State = flat arenas owned by the caller:
 *   params / grads  every learnable tensor in torch parameters() order and torch layout
 *                   (a reference nn.Module can alias its Parameters onto it);
 *   packed          kernel-layout copies of the conv weights, refreshed by b200ocl_net_pack /
 *                   b200ocl_net_sgd_step;
 *   bn_stats        per BatchNorm2d running_mean[c], running_var[c] in module order;
 *   bn_tracked      num_batches_tracked per BatchNorm2d.
 * Images are fp32 NCHW [N,3,H,W] exactly as the reference feeds them (no normalisation,
 * setup_elements.py:29-43); activations are NHWC inside the engine. */
typedef struct {
  int in_h, in_w;      /* 32x32 CIFAR, 84x84 Mini-ImageNet (setup_elements.py:11-17) */
  int nf;              /* 20 (Reduced_ResNet18, resnet.py:112-116) */
  int num_classes;     /* classifier width when head == 0 */
  int head;            /* 0 classifier | 1 linear | 2 mlp | 3 none */
  int feat_dim;        /* SupCon projection size (128) */
} b200ocl_net_desc;

typedef struct {
  float* params;
  float* grads;
  float* packed;
  float* bn_stats;
  int64_t* bn_tracked;
} b200ocl_net_state;

/* Sizes (in elements) of the arenas and basic shape facts. */
typedef struct {
  size_t n_params, n_packed, n_bn_stats;
  int n_bn, n_tensors, dim_in, out_dim;
} b200ocl_net_info;
int b200ocl_net_query(const b200ocl_net_desc* desc, b200ocl_net_info* info);
/* i-th parameter tensor in parameters() order: offset/numel in the arena, has_grad = 0 for the
 * SupConResNet encoder classifier that never receives a gradient. */
int b200ocl_net_tensor(const b200ocl_net_desc* desc, int i, size_t* offset, size_t* numel, int* has_grad);

/* packed <- params (call after loading weights). */
int b200ocl_net_pack(const b200ocl_net_desc* desc, const b200ocl_net_state* st, void* stream);

/* model.eval(); model.features(x) under no_grad (utils/utils.py:45-90): feat [N,dim_in]. */
size_t b200ocl_net_eval_workspace_bytes(const b200ocl_net_desc* desc, int N);
int b200ocl_net_features_eval(const b200ocl_net_desc* desc, const b200ocl_net_state* st, const float* x, int N,
                              float* feat, void* workspace, size_t workspace_bytes, void* stream);

/* model.train(); out = model.forward(x): batch-statistics BN, running stats and
 * num_batches_tracked updated (momentum 0.1, unbiased variance); activations are kept in
 * `workspace` for b200ocl_net_backward.  out [N,out_dim]. */
size_t b200ocl_net_train_workspace_bytes(const b200ocl_net_desc* desc, int N);
int b200ocl_net_forward_train(const b200ocl_net_desc* desc, const b200ocl_net_state* st, const float* x, int N,
                              float* out, void* workspace, size_t workspace_bytes, void* stream);

/* model.eval(); out = model.forward(x) WITH activations kept for a backward pass (utils/buffer/gss_greedy_update.py:16,
 * 80-83, 118-120: GSS-greedy differentiates the network in eval mode): BatchNorm uses the running statistics, nothing is
 * updated.  Its backward is b200ocl_net_backward with bit 1 of `accumulate` set. */
int b200ocl_net_forward_evalgrad(const b200ocl_net_desc* desc, const b200ocl_net_state* st, const float* x, int N,
                                 float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Train-mode forward with DEFERRED running statistics: identical outputs and saved activations, but the BatchNorm running
 * statistics (and num_batches_tracked) are not touched; the batch statistics of every BN stay in `workspace` until
 * b200ocl_net_apply_running_stats applies them (running = 0.9 running + 0.1 batch, tracked += 1).  The train-mode passes of
 * one replay step (exp_replay.py:40,62,84; scr.py:55) only interact through those statistics, so a caller may run them
 * concurrently on different streams and then apply their statistics in the reference's order. */
int b200ocl_net_forward_train_deferred(const b200ocl_net_desc* desc, const b200ocl_net_state* st, const float* x, int N,
                                       float* out, void* workspace, size_t workspace_bytes, void* stream);
int b200ocl_net_apply_running_stats(const b200ocl_net_desc* desc, const b200ocl_net_state* st, int N, void* workspace,
                                    size_t workspace_bytes, void* stream);

/* loss.backward() for the forward kept in `workspace` (same x): dout [N,out_dim] -> st->grads
 * (overwritten, or added to when accumulate != 0 -- exp_replay.py:55,77 accumulate two
 * backward passes before one opt.step()).  accumulate bit 1 (value 2): the forward was b200ocl_net_forward_evalgrad.
 * The weight-gradient launches run on a helper stream that is forked from / joined into `stream` with events (captured
 * as parallel branches when `stream` is being captured into a CUDA graph); the helper stream and its four events (one set per
 * caller stream, at most four caller streams per device) are the one piece of state the library creates itself, on the first call (B200OCL_WG_ASYNC=0: everything on `stream`). */
int b200ocl_net_backward(const b200ocl_net_desc* desc, const b200ocl_net_state* st, const float* x, const float* dout,
                         int N, void* workspace, size_t workspace_bytes, int accumulate, void* stream);

/* opt.step() of torch.optim.SGD without momentum over every tensor that has a gradient
 * (setup_elements.py:73-75), then refresh `packed`.  With dst != NULL the updated weights go
 * to dst->params / dst->packed instead (MIR's virtual step theta - lr*grad on a copy,
 * mir_retrieve.py:34-47); tensors without gradient are copied. */
int b200ocl_net_sgd_step(const b200ocl_net_desc* desc, const b200ocl_net_state* st, float lr, float weight_decay,
                         const b200ocl_net_state* dst, void* stream);

/* Mean cross-entropy (agents/base.py:95,113) over logits [N,C], labels [N]:
 * loss[1]; per_sample [N] (F.cross_entropy(reduction='none'), mir_retrieve.py:26-27);
 * dlogits [N,C] = d(mean CE)/dlogits; n_correct[1] = #(argmax == label).  Each output nullable. */
int b200ocl_ce_loss(const float* logits, const int64_t* labels, int N, int C, float* loss, float* per_sample,
                    float* dlogits, int64_t* n_correct, void* stream);

/* ---------------------------------------------------------------- SCR augmentation
 * The second view of agents/scr.py:18-24,54 (kornia RandomResizedCrop -> RandomHorizontalFlip ->
 * ColorJitter -> RandomGrayscale) in one kernel over NCHW fp32 images in [0,1].
 * params [N,12] f32 (device), per sample: x0, y0, crop_w, crop_h, flip, jitter_on,
 * brightness delta, contrast factor, saturation factor, hue shift in turns, order code
 * (four 2-bit ids, 0 brightness 1 contrast 2 saturation 3 hue, first op in the low bits), gray.
 * The random draws are the caller's; kornia's exact arithmetic is parity-unpinned (absent). */
int b200ocl_scr_augment(const float* x, float* out, const float* params, int N, int H, int W, void* stream);

/* ---------------------------------------------------------------- tensor-core self test
 * D[128,N] = A[128,K] * B[N,K]^T on tcgen05 (kind::tf32, TMEM accumulator) with the descriptor
 * encodings the tensor-core convolution uses; mode 0 single TF32 pass, mode 1 the 3xTF32 split.
 * status[0] = 0 on completion, 1 if the MMA completion barrier timed out. */
int b200ocl_selftest_umma_tf32(const float* A, const float* B, float* D, int N, int K, int mode, int* status,
                               void* stream);

/* One convolution (ks = 3 pad 1, or ks = 1 pad 0; stride 1 or 2; dgrad = 0: out[N,Hout,Wout,cout] = x[N,H,W,cin] * w;
 * dgrad = 1, 3x3 stride 1 only: its data gradient, x = dz[N,H,W,cout] -> out[N,H,W,cin]) through a chosen kernel
 * family, NHWC fp32, weights OIHW: path 0 automatic, 1 CUDA-core kernels, 2 tcgen05 with im2col tiles
 * (conv_tc.cu), 3 tcgen05 fed from a halo strip (conv_tcp.cu).  mode 0 raw store, 1 accumulate into out, 2 train
 * (forward only): raw store plus stats_out[4*cout] = batch mean, 1/sqrt(var+eps), running mean, running var
 * updated from zero with momentum 0.1.  Exists so that tests can pin every convolution kernel against a reference
 * convolution; fails with B200OCL_EUNSUPPORTED when the path does not cover the shape. */
size_t b200ocl_conv_selftest_workspace_bytes(int N, int cin, int cout, int H, int W, int ks, int stride);
int b200ocl_conv_selftest(const float* x, const float* w_oihw, float* out, int N, int H, int W, int cin, int cout,
                          int ks, int stride, int dgrad, int path, int mode, float* stats_out, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Window variant: A is read in place from a larger swizzled buffer P[rows][32] of 128-byte rows (tile row
 * 8g + r = P row start_row + g * sbo_rows + r), start address unaligned to the swizzle repeat when
 * start_row % 8 != 0; base_off_mode 1 sets the descriptor's base-offset field to start_row & 7.  Validates the
 * addressing the halo-patch tensor-core convolution relies on.  D[128,N] = A_window * B[N,32]^T, single TF32 pass. */
int b200ocl_selftest_umma_window(const float* P, const float* B, float* D, int rows, int start_row, int sbo_rows,
                                 int base_off_mode, int N, int* status, void* stream);

/* MN-major variant: both operands are windows into strips of 128-byte rows (32 fp32 per row, stored with the 32-byte-base
 * 128-byte swizzle), the contraction runs over rows:
 *   D[32 j + c][32 q + n] = sum_{g < 2 ksteps} sum_{t < 4} PA[a_row0 + j a_lbo_rows + g a_sbo_rows + t][c]
 *                                                        * PB[b_row0 + q b_lbo_rows + g b_sbo_rows + t][n],   single TF32 pass.
 * layout_type 1 = SWIZZLE_128B_BASE32B, what kind::tf32 needs for MN-major operands; layout_type 2 = SWIZZLE_128B, kept
 * reachable because of what it does on B200: the instruction completes and D is all zeros (tests/test_gpu_umma.py).
 * a_lbo_rows = 1, a_sbo_rows = 4 (M block j = the strip shifted by j rows, K over 8 consecutive rows) is what the
 * tensor-core weight gradient relies on. */
int b200ocl_selftest_umma_mn(const float* PA, const float* PB, float* D, int rows_a, int rows_b, int a_row0, int a_lbo_rows,
                             int a_sbo_rows, int b_row0, int b_lbo_rows, int b_sbo_rows, int ksteps, int N, int layout_type,
                             int* status, void* stream);

/* Weight gradient of one 3x3 stride-1 pad-1 convolution through the tcgen05 kernel (wgrad_tc.cu), NHWC fp32 inputs,
 * dw in OIHW: the cuDNN weight-gradient call behind loss.backward() for nn.Conv2d (models/resnet.py:11-12).  Exists so
 * that tests can pin the kernel against a reference; B200OCL_EUNSUPPORTED when the geometry is not covered. */
size_t b200ocl_wgrad_tc_selftest_workspace_bytes(int N, int H, int W, int cin, int cout);
int b200ocl_wgrad_tc_selftest(const float* x, const float* dz, float* dw_oihw, int N, int H, int W, int cin, int cout,
                              void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200OCL_H_ */
