/*
 * mednet_b200.h -- C ABI of the B200-native (sm_100a) UNet3D hot path for torch-mednet.
 *
 * The reference (tobiashepp/torch-mednet) has no FFI: every FLOP of its hot path runs inside
 * third-party PyTorch/ATen ops (requirements.txt:5).  Each entry point below therefore replaces one
 * ATen call site of the reference; the call site is cited per function as `ref: file:line`
 * (paths relative to the reference root, midasmednet/ = mm/).
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller (PyTorch caching allocator on the Python
 *     side).  Launchers never allocate, free or synchronise; they enqueue on `stream` and return.
 *   - Activations are NDHWC ("channels-last-3d"), contiguous, dtype MEDNET_F32 or MEDNET_BF16.
 *     Logits / losses are NCDHW fp32 at the module boundary (the layout the reference returns).
 *   - Return value: 0 = success; negative = MEDNET_E* (unsupported shape / dtype / alignment -- the
 *     Python side raises, there is no fallback); positive = cudaError_t of the failed launch.
 *   - Re-entrant and stream ordered; the only global state is an immutable lazily built attribute /
 *     tensor-map encoder cache behind a mutex.
 *   - `workspace` is scratch of at least mednet_<op>_workspace_bytes(p) bytes, 256-byte aligned.
 */
#ifndef MEDNET_B200_H
#define MEDNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mednet_stream_t;

#define MEDNET_ABI_VERSION 3

/* dtypes */
#define MEDNET_F32  0
#define MEDNET_BF16 1
#define MEDNET_U8   2
#define MEDNET_I64  3

/* activations (ref: mm/unet/components.py:35-40) */
#define MEDNET_ACT_NONE  0
#define MEDNET_ACT_RELU  1
#define MEDNET_ACT_LEAKY 2   /* negative slope in act_param (reference uses 0.1) */
#define MEDNET_ACT_ELU   3   /* alpha = 1 */

/* convolution implementations */
#define MEDNET_IMPL_AUTO    0
#define MEDNET_IMPL_SIMT    1   /* CUDA-core fp32-accumulate implicit GEMM (fp32 validation mode, odd shapes) */
#define MEDNET_IMPL_TCGEN05 2   /* tcgen05/TMEM tensor-core implicit GEMM with TMA halo tiles (bf16) */

/* packed-weight layouts produced by mednet_conv3d_pack_weights */
#define MEDNET_WPACK_SIMT_FPROP    0   /* [Cout][27][Cin]            */
#define MEDNET_WPACK_SIMT_DGRAD    1   /* [Cin][27 flipped][Cout]     */
#define MEDNET_WPACK_TC_FPROP      2   /* [27][Cout][Cin]            */
#define MEDNET_WPACK_TC_DGRAD      3   /* [27 flipped][Cin][Cout]     */
#define MEDNET_WPACK_TC_CONVT_F    4   /* transposed conv fprop: [8 parity classes][27 window taps][Cout][Cin] (8*27*Cin*Cout) */
#define MEDNET_WPACK_TC_CONVT_B    5   /* transposed conv dgrad: [8 parity classes][27 window taps][Cin][Cout]                 */
#define MEDNET_WPACK_TC_UPCONV_F   6   /* conv over a nearest-upsampled input, fprop: [8 output parity classes][27 coarse window taps]
                                          [Cout][Cin], fine taps landing on the same coarse voxel summed (MEDNET_GATHER_UPCONV_F) */
#define MEDNET_WPACK_TC_UPCONV_B   7   /* same, data gradient w.r.t. the coarse input: [8 dY parity classes][27][Cin][Cout]          */

/* gather modes of the generic implicit GEMM */
#define MEDNET_GATHER_CONV3   0   /* 3x3x3, stride 1, pad 1               */
#define MEDNET_GATHER_CONVT_F 1   /* transposed k3 s2 p1 op1, output->input */
#define MEDNET_GATHER_CONVT_B 2   /* transposed k3 s2 p1 op1, input->output */
/* 3x3x3 conv (pad 1) of the NEAREST-UPSAMPLED (x2) input, computed on the coarse input itself -- the decoder join
 * F.interpolate(x, 'nearest') + Conv3d of mm/unet/components.py:277-280 then :8-9.  For output voxel v = 2u + p the fine tap
 * o in {-1,0,1} reads coarse voxel u + floor((p + o) / 2): 2 distinct coarse voxels per axis, so each of the 8 output parity
 * classes is a 2x2x2 convolution over the coarse grid with summed weights (8 taps instead of 27; zero padding of the coarse
 * grid == zero padding of the fine one).  UPCONV_F: x coarse [N,Di..] -> y fine [N,2Di..]; UPCONV_B: its data gradient, x = dY
 * fine -> y = d(coarse input).  Tensor-core implementation only (bf16, sizes exactly x2). */
#define MEDNET_GATHER_UPCONV_F 3
#define MEDNET_GATHER_UPCONV_B 4

/* error codes */
#define MEDNET_OK            0
#define MEDNET_EINVAL       (-1)
#define MEDNET_EUNSUPPORTED (-2)
#define MEDNET_EALIGN       (-3)
#define MEDNET_EWORKSPACE   (-4)
#define MEDNET_ENODRIVER    (-5)

int         mednet_abi_version(void);
const char* mednet_error_string(int code);
/* 1 when the current device is compute capability 10.x (tcgen05/TMEM present) */
int         mednet_device_has_tcgen05(void);
int         mednet_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * Layout / dtype conversion at the module boundary.
 * ref: the reference keeps NCDHW fp32 everywhere (mm/segmentation.py:59 `batch['data'].float()`).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* src; void* dst;
  int64_t N, C, S;            /* S = D*H*W */
  int32_t src_dtype, dst_dtype;
  int32_t to_channels_last;   /* 1: NCDHW -> NDHWC, 0: NDHWC -> NCDHW */
} mednet_layout_params;
int mednet_layout_convert(const mednet_layout_params* p, mednet_stream_t stream);

/* Channel padding / extraction of NDHWC rows: dst[r][c] = c < Cs ? src[r][c] : 0 (Cd >= Cs) or dst[r][c] = src[r][c]
 * for c < Cd (Cd < Cs).  Used to run the in_channels = 1 first layer (mm/segmentation.py:30-31) on the tensor cores:
 * the image is zero-padded to 16 channels once, the input gradient is cut back to the real channels. */
typedef struct {
  const void* src; void* dst; int64_t rows; int32_t Cs, Cd, dtype;
} mednet_chpad_params;
int mednet_channel_pad(const mednet_chpad_params* p, mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 3x3x3 convolution, stride 1, zero padding 1 (also hosts ConvTranspose3d through `gather`).
 * ref: mm/unet/components.py:8-9 (nn.Conv3d via create_conv :41-44); transposed: :259-264.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void*  w_oidhw;   /* fp32 PyTorch layout: conv (Cout,Cin,3,3,3); transposed conv (Cin,Cout,3,3,3) */
  void*        w_packed;
  int32_t Cin, Cout;      /* of the forward op */
  int32_t dtype;          /* dtype of w_packed */
  int32_t layout;         /* MEDNET_WPACK_* */
  int32_t transposed;     /* 1: w_oidhw is a ConvTranspose3d weight */
} mednet_wpack_params;
int mednet_conv3d_pack_weights(const mednet_wpack_params* p, mednet_stream_t stream);

typedef struct {
  const void*  x;         /* [N, Di,Hi,Wi, K]  rows gathered from here                     */
  const void*  w;         /* packed weights, see MEDNET_WPACK_* ([Nout][27][K] or [27][Nout][K]) */
  const float* bias;      /* [Nout] or NULL                                                 */
  const void*  addend;    /* optional tensor of y's shape added before the activation (skip sum, ref :284) */
  void*        y;         /* [N, Do,Ho,Wo, Nout]                                            */
  int32_t N, Di, Hi, Wi, Do, Ho, Wo;
  int32_t K, Nout;        /* reduction channels / output channels of THIS gemm              */
  int32_t dtype, act; float act_param;
  int32_t gather;         /* MEDNET_GATHER_*                                                */
  int32_t impl;           /* MEDNET_IMPL_*                                                  */
  /* fp32 partial sums between two launches that together form ONE convolution (tensor-core implementation only): the
   * first launch writes y as fp32 without bias/addend/activation (y_f32 = 1), the second takes it as its addend
   * (addend_f32 = 1) -- no rounding of the partial sum, so the pre-activation (and every ReLU' decision taken from it)
   * is the one a single launch over all input channels would produce. */
  int32_t y_f32, addend_f32;
} mednet_conv3d_params;
size_t mednet_conv3d_workspace_bytes(const mednet_conv3d_params* p);
/* Resolves p->impl (MEDNET_IMPL_AUTO included) to the implementation that will run, so the caller can
 * pack the weights in the matching layout; negative = MEDNET_E*. */
int    mednet_conv3d_select_impl(const mednet_conv3d_params* p);
/* y = act(gather-conv(x, w) + bias + addend).  fprop and dgrad are the same launcher with different
 * packed weights (dgrad of a stride-1 conv is the conv of dY with the flipped, transposed filter). */
int mednet_conv3d_fprop(const mednet_conv3d_params* p, void* workspace, size_t workspace_bytes,
                        mednet_stream_t stream);

typedef struct {
  const void* a;          /* [N, Da,Ha,Wa, Ca] "row" operand (conv: dY; transposed conv: x)          */
  const void* b;          /* [N, Db,Hb,Wb, Cb] gathered operand (conv: x; transposed conv: dY)       */
  float*      dw;         /* fp32 gradient in PyTorch layout (Ca,Cb,3,3,3)                           */
  float*      dbias;      /* optional bias gradient = column sums of the output-gradient operand
                             (a, [Ca], for GATHER_CONV3; b, [Cb], for GATHER_CONVT_B) or NULL        */
  int32_t N, Da, Ha, Wa, Db, Hb, Wb, Ca, Cb;
  int32_t dtype, gather, impl;
  int32_t accumulate;     /* 1: dw += result (gradient accumulation), 0: overwrite                   */
  /* Strided destination (tensor-core implementation): the (Ca,Cb,27) result is written into a weight-gradient tensor of
   * `dw_ld` channels per row, starting at channel `dw_c0`: dw[(ca * dw_ld + dw_c0 + cb) * 27 + tap], or with
   * dw_transposed dw[(cb * dw_ld + dw_c0 + ca) * 27 + tap].  dw_ld = 0: dense (Ca,Cb,3,3,3).  Lets the two halves of the
   * decoder-join weight gradient (skip channels / upsampled channels) land in the one Conv3d gradient tensor. */
  int32_t dw_ld, dw_c0, dw_transposed;
} mednet_wgrad_params;
size_t mednet_conv3d_wgrad_workspace_bytes(const mednet_wgrad_params* p);
int    mednet_conv3d_wgrad_select_impl(const mednet_wgrad_params* p);
int    mednet_conv3d_wgrad(const mednet_wgrad_params* p, void* workspace, size_t workspace_bytes,
                           mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Final 1x1x1 convolution with bias: NDHWC activations -> NCDHW fp32 logits.
 * ref: mm/unet/model.py:77,102 (UNet3D) and :179,207 (ResidualUNet3D).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void*  x;         /* [N*S, Cin] NDHWC                         */
  const float* w;         /* [Cout, Cin] fp32                         */
  const float* bias;      /* [Cout]                                   */
  float*       y;         /* [N, Cout, S] fp32                        */
  int64_t N, S; int32_t Cin, Cout, dtype;
} mednet_conv1_params;
int mednet_conv1x1_fwd(const mednet_conv1_params* p, mednet_stream_t stream);
typedef struct {
  const void*  x; const float* w; const float* dy;   /* dy [N, Cout, S] fp32 */
  void*        dx;        /* [N*S, Cin] NDHWC (dtype)                 */
  float*       dw;        /* [Cout, Cin]                              */
  float*       db;        /* [Cout]                                   */
  int64_t N, S; int32_t Cin, Cout, dtype; int32_t accumulate;
  int32_t in_act; float in_act_param;   /* deferred activation derivative: dx *= in_act'(x), x = in_act(pre) (see below) */
} mednet_conv1_bwd_params;
size_t mednet_conv1x1_bwd_workspace_bytes(const mednet_conv1_bwd_params* p);
int    mednet_conv1x1_bwd(const mednet_conv1_bwd_params* p, void* workspace, size_t workspace_bytes,
                          mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Deferred activation derivative (`in_act`).  A convolution whose epilogue applies an activation
 * (order 'gcr': conv -> ReLU) does not run a separate dpre = dy * act'(y) pass in backward.  Instead every
 * CONSUMER of its output y (the next GroupNorm, MaxPool3d, the upsample/concat, the final 1x1x1 conv) is
 * told `in_act` and multiplies the gradient it returns by act'(y), expressed through y itself -- its own
 * input, which it reads anyway (ReLU: y > 0; LeakyReLU: y > 0 ? 1 : slope; ELU: y > 0 ? 1 : y + 1).  The
 * mask is linear, so consumers that share y each apply it and autograd sums the results.
 * ref: the in-place ReLU/LeakyReLU/ELU modules after the conv, mm/unet/components.py:35-40.
 * ---------------------------------------------------------------------------------------------- */

/* ------------------------------------------------------------------------------------------------
 * GroupNorm (+ optional residual add + activation), eps 1e-5, biased variance.
 * ref: mm/unet/components.py:57 (nn.GroupNorm), :36-40 (activation after it), :177-178 (residual).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void*  x;         /* [N, S, C]                                                        */
  const float* gamma; const float* beta;      /* [C]                                          */
  const void*  residual;  /* optional [N,S,C] added before the activation, or NULL            */
  void*        y;         /* [N, S, C]                                                        */
  float*       mean; float* rstd;             /* [N, G] outputs (saved for backward)          */
  int64_t N, S; int32_t C, G, dtype, act; float act_param, eps;
} mednet_gn_fwd_params;
size_t mednet_groupnorm_fwd_workspace_bytes(const mednet_gn_fwd_params* p);
int    mednet_groupnorm_fwd(const mednet_gn_fwd_params* p, void* workspace, size_t workspace_bytes,
                            mednet_stream_t stream);
typedef struct {
  const void*  x; const void* y;              /* y = saved forward output (needed iff act != NONE) */
  const void*  dy;
  const float* gamma; const float* mean; const float* rstd;
  void*        dx;        /* [N,S,C]                                                          */
  void*        dresidual; /* optional [N,S,C]: gradient wrt the residual input, or NULL       */
  float*       dgamma; float* dbeta;          /* [C]                                          */
  int64_t N, S; int32_t C, G, dtype, act; float act_param; int32_t accumulate;
  int32_t in_act; float in_act_param;   /* deferred activation derivative of the PRODUCER of x: dx *= in_act'(x) */
} mednet_gn_bwd_params;
size_t mednet_groupnorm_bwd_workspace_bytes(const mednet_gn_bwd_params* p);
int    mednet_groupnorm_bwd(const mednet_gn_bwd_params* p, void* workspace, size_t workspace_bytes,
                            mednet_stream_t stream);

/* Stand-alone activation backward from the saved OUTPUT (conv epilogue activations):
 * dx = dy * act'(y).  ref: ReLU/LeakyReLU/ELU in-place modules, mm/unet/components.py:35-40. */
typedef struct {
  const void* y; const void* dy; void* dx; int64_t numel; int32_t dtype, act; float act_param;
} mednet_act_bwd_params;
int mednet_act_bwd(const mednet_act_bwd_params* p, mednet_stream_t stream);
/* Stand-alone activation forward (orders such as 'crg' where the activation is not fused). */
typedef struct {
  const void* x; void* y; int64_t numel; int32_t dtype, act; float act_param;
} mednet_act_fwd_params;
int mednet_act_fwd(const mednet_act_fwd_params* p, mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * MaxPool3d(kernel 2, stride 2, floor mode) with argmax.
 * ref: mm/unet/components.py:210 (nn.MaxPool3d), applied at :224.
 * Tie rule: first maximum in (d,h,w) scan order; NaN wins (ATen semantics, golden semantics.npz).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;          /* [N, D, H, W, C]                                                  */
  void*       y;          /* [N, D/2, H/2, W/2, C]                                            */
  uint8_t*    idx;        /* [N, D/2, H/2, W/2, C] window-local argmax code dz*4+dy*2+dx      */
  int32_t N, D, H, W, C, dtype;
} mednet_pool_params;
int mednet_maxpool3d_fwd(const mednet_pool_params* p, mednet_stream_t stream);
typedef struct {
  const void* dy; const uint8_t* idx; void* dx; int32_t N, D, H, W, C, dtype;
  const void* y;          /* pooled forward output (= x at the argmax); needed iff in_act != NONE       */
  int32_t in_act; float in_act_param;   /* deferred activation derivative of the producer of x         */
  const void* addend;     /* optional [N, D, H, W, C]: gradient of x's OTHER consumer (the skip connection),
                             added to dx in the same pass (replaces autograd's separate accumulation add)  */
} mednet_pool_bwd_params;
int mednet_maxpool3d_bwd(const mednet_pool_bwd_params* p, mednet_stream_t stream);
/* Expand the 3-bit codes to ATen's int64 flat D*H*W indices, NCDHW order (parity checks only). */
int mednet_maxpool3d_indices_i64(const uint8_t* idx, int64_t* out, int32_t N, int32_t D, int32_t H,
                                 int32_t W, int32_t C, mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Nearest-neighbour upsampling to the skip tensor's size fused with the channel concat
 * (encoder features first).  ref: mm/unet/components.py:277-280 (F.interpolate + torch.cat).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* skip;       /* [N, D, H, W, Cs]                                                 */
  const void* low;        /* [N, d, h, w, Cl]                                                 */
  void*       out;        /* [N, D, H, W, Cs+Cl]                                              */
  int32_t N, D, H, W, d, h, w, Cs, Cl, dtype;
} mednet_upcat_params;
int mednet_upsample_concat_fwd(const mednet_upcat_params* p, mednet_stream_t stream);
typedef struct {
  const void* dout; void* dskip; void* dlow; int32_t N, D, H, W, d, h, w, Cs, Cl, dtype;
  const void* skip; const void* low;    /* forward inputs; needed iff the matching *_act != NONE          */
  int32_t skip_act, low_act; float skip_act_param, low_act_param;   /* deferred activation derivatives     */
} mednet_upcat_bwd_params;
int mednet_upsample_concat_bwd(const mednet_upcat_bwd_params* p, mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm over the VIRTUAL concat (skip, nearest-upsampled low): the decoder's first layer in order
 * 'gcr' normalises cat((skip, up(low)), 1) (mm/unet/components.py:277-280 then :57).  The concat tensor is
 * never materialised: statistics come from per-channel sums over skip and low (exact 2x upsampling
 * replicates every low voxel 8 times), the apply pass reads skip/low and writes the normalised tensor the
 * convolution consumes, and the backward pass writes dskip / dlow (sum over the 8 children) directly, with
 * the deferred activation derivatives of the producers of skip / low applied.
 * Requires D = 2d, H = 2h, W = 2w (otherwise MEDNET_EUNSUPPORTED: use upsample_concat + groupnorm).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void*  skip;      /* [N, D, H, W, Cs]                                                   */
  const void*  low;       /* [N, d, h, w, Cl]                                                   */
  const float* gamma; const float* beta;      /* [Cs + Cl]                                      */
  void*        y;         /* [N, D, H, W, Cs + Cl] normalised concat                            */
  float*       mean; float* rstd;             /* [N, G]                                         */
  int32_t N, D, H, W, d, h, w, Cs, Cl, G, dtype; float eps;
} mednet_upcat_gn_fwd_params;
size_t mednet_upcat_groupnorm_fwd_workspace_bytes(const mednet_upcat_gn_fwd_params* p);
int    mednet_upcat_groupnorm_fwd(const mednet_upcat_gn_fwd_params* p, void* workspace, size_t workspace_bytes,
                                  mednet_stream_t stream);
/* Split variant for the upsample-aware decoder convolution (MEDNET_GATHER_UPCONV_*): same statistics over the virtual
 * concat, but the normalised skip part is written at full resolution and the normalised low part on the COARSE grid (a
 * per-channel affine commutes with nearest upsampling): the [N, D, H, W, Cs + Cl] tensor and its gradient never exist.
 * Backward takes the gradients of the two outputs (dy_low already summed over the 8 children of each coarse voxel). */
typedef struct {
  const void*  skip; const void* low;
  const float* gamma; const float* beta;
  void*        y_skip;    /* [N, D, H, W, Cs] */
  void*        y_low;     /* [N, d, h, w, Cl] */
  float*       mean; float* rstd;
  int32_t N, D, H, W, d, h, w, Cs, Cl, G, dtype; float eps;
} mednet_upcat_gn_split_fwd_params;
size_t mednet_upcat_groupnorm_split_fwd_workspace_bytes(const mednet_upcat_gn_split_fwd_params* p);
int    mednet_upcat_groupnorm_split_fwd(const mednet_upcat_gn_split_fwd_params* p, void* workspace, size_t workspace_bytes,
                                        mednet_stream_t stream);
typedef struct {
  const void*  skip; const void* low; const void* dy_skip; const void* dy_low;
  const float* gamma; const float* mean; const float* rstd;
  void*        dskip; void* dlow;
  float*       dgamma; float* dbeta;
  int32_t N, D, H, W, d, h, w, Cs, Cl, G, dtype, accumulate;
  int32_t skip_act, low_act; float skip_act_param, low_act_param;
} mednet_upcat_gn_split_bwd_params;
size_t mednet_upcat_groupnorm_split_bwd_workspace_bytes(const mednet_upcat_gn_split_bwd_params* p);
int    mednet_upcat_groupnorm_split_bwd(const mednet_upcat_gn_split_bwd_params* p, void* workspace, size_t workspace_bytes,
                                        mednet_stream_t stream);
typedef struct {
  const void*  skip; const void* low; const void* dy;      /* dy [N, D, H, W, Cs + Cl]           */
  const float* gamma; const float* mean; const float* rstd;
  void*        dskip;     /* [N, D, H, W, Cs]                                                   */
  void*        dlow;      /* [N, d, h, w, Cl]                                                   */
  float*       dgamma; float* dbeta;          /* [Cs + Cl]                                      */
  int32_t N, D, H, W, d, h, w, Cs, Cl, G, dtype, accumulate;
  int32_t skip_act, low_act; float skip_act_param, low_act_param;
} mednet_upcat_gn_bwd_params;
size_t mednet_upcat_groupnorm_bwd_workspace_bytes(const mednet_upcat_gn_bwd_params* p);
int    mednet_upcat_groupnorm_bwd(const mednet_upcat_gn_bwd_params* p, void* workspace, size_t workspace_bytes,
                                  mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Losses on NCDHW logits.  `logits` may be a channel slice of a wider tensor: element (n,c,s) lives
 * at logits[n*batch_stride + c*S + s].  Labels are MEDNET_I64 or MEDNET_U8 class maps [N,S].
 * All reductions are two-stage and deterministic.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void*  logits; const void* labels; const float* weight;   /* weight [C] or NULL */
  float*       sums;      /* [3][C]: intersect (unweighted), sum p, count -- saved for backward */
  float*       dice;      /* [C] per-channel dice (weighted on the intersect, ref mm/unet/loss.py:44-45) */
  float*       loss;      /* scalar: mean_c(1 - dice_c)                                          */
  int64_t N, S, batch_stride; int32_t C, logits_dtype, label_dtype, sigmoid; float epsilon;
} mednet_dice_params;
size_t mednet_dice_workspace_bytes(const mednet_dice_params* p);
/* ref: mm/unet/loss.py:114-130 (DiceLoss.forward) with :10-48, :58-88; dice_metric :51-55. */
int    mednet_dice_fwd(const mednet_dice_params* p, void* workspace, size_t workspace_bytes,
                       mednet_stream_t stream);
typedef struct {
  const void*  logits; const void* labels; const float* weight; const float* sums;
  const float* grad_out;  /* device scalar upstream gradient                                     */
  void*        dlogits;   /* same addressing as logits (batch_stride_out), dtype dlogits_dtype   */
  int64_t N, S, batch_stride, batch_stride_out; int32_t C, logits_dtype, label_dtype, dlogits_dtype, sigmoid;
  float epsilon;
} mednet_dice_bwd_params;
int    mednet_dice_bwd(const mednet_dice_bwd_params* p, mednet_stream_t stream);

typedef struct {
  const void*  logits; const void* labels; const float* weight;
  float*       sums;      /* [2]: sum w[y]*nll, sum w[y]  (saved)                                */
  float*       loss;      /* scalar                                                               */
  int64_t N, S, batch_stride; int32_t C, logits_dtype, label_dtype;
} mednet_ce_params;
size_t mednet_ce_workspace_bytes(const mednet_ce_params* p);
/* ref: mm/segmentation.py:49, mm/landmarks.py:49 (nn.CrossEntropyLoss(weight), weighted mean). */
int    mednet_ce_fwd(const mednet_ce_params* p, void* workspace, size_t workspace_bytes, mednet_stream_t stream);
typedef struct {
  const void*  logits; const void* labels; const float* weight; const float* sums; const float* grad_out;
  void*        dlogits;
  int64_t N, S, batch_stride, batch_stride_out; int32_t C, logits_dtype, label_dtype, dlogits_dtype;
} mednet_ce_bwd_params;
int    mednet_ce_bwd(const mednet_ce_bwd_params* p, mednet_stream_t stream);

typedef struct {
  const void*  pred;      /* channel slice, element (n,c,s) at pred[n*batch_stride + c*S + s]    */
  const void*  target;    /* [N, L, S] MEDNET_U8 or MEDNET_F32 heatmaps (0..255)                 */
  const float* weight;    /* [L] per-channel weights (ref mm/landmarks.py:128-132)                */
  float*       per_channel; /* [L] mean error per channel                                         */
  float*       loss;      /* scalar sum_c w_c * mean_c                                            */
  int64_t N, S, batch_stride; int32_t L, pred_dtype, target_dtype, l1;
} mednet_hmloss_params;
size_t mednet_heatmap_loss_workspace_bytes(const mednet_hmloss_params* p);
/* ref: mm/landmarks.py:53-55,125-134 (MSELoss / L1Loss per channel, python loop). */
int    mednet_heatmap_loss_fwd(const mednet_hmloss_params* p, void* workspace, size_t workspace_bytes,
                               mednet_stream_t stream);
typedef struct {
  const void*  pred; const void* target; const float* weight; const float* grad_out;
  void*        dpred;
  int64_t N, S, batch_stride, batch_stride_out; int32_t L, pred_dtype, target_dtype, dpred_dtype, l1;
} mednet_hmloss_bwd_params;
int    mednet_heatmap_loss_bwd(const mednet_hmloss_bwd_params* p, mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Inference epilogue: uint8 heatmaps (clip to [0,255], truncate) and uint8 argmax class map written
 * directly from logits.  ref: examples/predict.py:88-94.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* logits;     /* [N, L+K, S] */
  uint8_t*    out;        /* [N, L+1, S] */
  int64_t N, S; int32_t L, K, logits_dtype;
} mednet_predict_params;
int mednet_predict_epilogue(const mednet_predict_params* p, mednet_stream_t stream);

/* Test-time activation over the channel axis of NCDHW fp32 logits: softmax (sigmoid = 0) or sigmoid.
 * ref: mm/unet/model.py:79-82,107-108. */
int mednet_final_activation(const float* logits, float* out, int64_t N, int64_t S, int32_t C, int32_t sigmoid,
                            mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm affine -> 3x3x3 conv on a tensor that needs NO gradient (the image: first layer of the 'gcr' networks,
 * mm/unet/components.py:45-57 then :8-9).  The conv's input gradient would only feed the two per-channel sums of the
 * GroupNorm backward; the adjoint identity <dgrad(dpre), u> = <dpre, conv(u)> yields them from the weight gradient taken
 * against the normalised input xhat (ghat) and from border-class sums of dpre, so neither the dgrad nor the GroupNorm
 * backward of that layer is launched (csrc/input_affine.cu has the algebra).
 *   mednet_border_class_sums: bins[((a*3+b)*3+c)][C] = sum of x over the voxels of class a / b / c along d / h / w
 *     (0 interior, 1 first plane, 2 last plane), over the whole batch.  D, H, W >= 2.
 *   mednet_conv3d_affine_input_grads: dw = gamma*ghat + beta*B, dgamma = sum w*ghat, dbeta = sum w*B with
 *     B[co][t] = sum of the bins whose voxels see tap t inside the volume; `dtype` BF16 rounds w as the conv did.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x; float* bins; int32_t N, D, H, W, C, dtype;
} mednet_border_sums_params;
size_t mednet_border_class_sums_workspace_bytes(const mednet_border_sums_params* p);
int    mednet_border_class_sums(const mednet_border_sums_params* p, void* workspace, size_t workspace_bytes,
                                mednet_stream_t stream);
typedef struct {
  const float* w; const float* ghat; const float* bins; const float* gamma; const float* beta;
  float* dw; float* dgamma; float* dbeta; int32_t Cout, Cin, dtype;
} mednet_affine_input_params;
int mednet_conv3d_affine_input_grads(const mednet_affine_input_params* p, mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused Adam on a flat fp32 parameter bucket (PyTorch defaults: no weight decay, no amsgrad).
 * ref: mm/segmentation.py:119-120, mm/landmarks.py:176-177 (torch.optim.Adam(self.parameters(), lr)).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  int64_t numel; float lr, beta1, beta2, eps, grad_scale; int32_t step;
} mednet_adam_params;
int mednet_adam_step(const mednet_adam_params* p, mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Landmark path: Gaussian heatmap rendering and (soft-)argmax extraction.
 * ABSENT from the reference (heatmaps are read pre-rendered as uint8, mm/dataset.py:261-262);
 * builder specification in oracle/heatmaps.py.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* points;    /* [N, L, 3] voxel coordinates (d,h,w) */
  const float* sigmas;    /* [L]                                 */
  uint8_t*     out;       /* [N, L, D, H, W]                     */
  int32_t N, L, D, H, W;
} mednet_hmrender_params;
int mednet_heatmap_render(const mednet_hmrender_params* p, mednet_stream_t stream);
typedef struct {
  const void* heatmaps;   /* [N*L, S] */
  int64_t*    argmax;     /* [N*L, 3] (d,h,w), first maximal index        */
  float*      soft;       /* [N*L, 3] soft-argmax with temperature beta, or NULL */
  int64_t NL; int32_t D, H, W, dtype; float beta;
} mednet_landmark_params;
size_t mednet_landmark_workspace_bytes(const mednet_landmark_params* p);
int    mednet_landmark_extract(const mednet_landmark_params* p, void* workspace, size_t workspace_bytes,
                               mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Sliding-window helpers (tile gather with zero padding / centre-crop scatter), uint8 and float.
 * ref: mm/dataset.py:349-389 (grid_patch_generator), :444-474 (add_processed_batch).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* volume;     /* [C, X, Y, Z] source volume (NCDHW without N), dtype src_dtype */
  void*       tiles;      /* [B, P0,P1,P2, C] NDHWC tiles, dtype dst_dtype                  */
  const int32_t* origins; /* [B, 3] device array of tile origins in PADDED coordinates      */
  int32_t B, C, X, Y, Z, P0, P1, P2, O0, O1, O2, src_dtype, dst_dtype;
  int32_t ncdhw_out;      /* 0: tiles [B, P0,P1,P2, C] (NDHWC, network input); 1: tiles [B, C, P0,P1,P2] (label / heatmap
                             crops of the GPU-resident patch sampler, ref mm/dataset.py:315-336).  MEDNET_U8 -> MEDNET_U8 is
                             supported in addition to the float dtypes. */
} mednet_tile_gather_params;
int mednet_tile_gather(const mednet_tile_gather_params* p, mednet_stream_t stream);
/* Random-patch crop of a whole batch in one launch (ref mm/dataset.py:315-330 runs per sample on the host): every
 * sample names its own source volume, so subjects of different shape share the launch.  No padding: patches lie inside
 * their volume (get_random_patch_indices, dataset.py:54-88); voxels outside read as 0 all the same. */
typedef struct {
  const int64_t* table;   /* [B, 8] device array per sample: volume address ([C, X, Y, Z], src_dtype), X, Y, Z,
                             origin0, origin1, origin2, 0                                                    */
  void*          tiles;   /* ncdhw_out ? [B, C, P0,P1,P2] : [B, P0,P1,P2, C], dst_dtype                      */
  int64_t tile_stride;    /* elements between consecutive samples in `tiles`; 0 = dense (C*P0*P1*P2).  Lets the
                             heatmap and class-map crops land in their channel ranges of one label tensor   */
  int32_t B, C, P0, P1, P2, src_dtype, dst_dtype, ncdhw_out;   /* dtypes: f32/bf16 -> f32/bf16, or u8 -> u8 */
} mednet_patch_gather_params;
int mednet_patch_gather(const mednet_patch_gather_params* p, mednet_stream_t stream);
typedef struct {
  const uint8_t* tiles;   /* [B, Co, P0*P1*P2] epilogue output                                */
  uint8_t*       volume;  /* [Co, X, Y, Z]                                                     */
  const int32_t* origins; /* [B, 3]                                                            */
  int32_t B, Co, X, Y, Z, P0, P1, P2, O0, O1, O2;
} mednet_tile_scatter_params;
int mednet_tile_scatter(const mednet_tile_scatter_params* p, mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Intensity augmentation of a sampled patch batch (opt-in, `--data_augmentation`).
 * ref: examples/train_seg.py:82-86, train_ldmks.py:82-84 compose BrightnessTransform(mu=0, sigma=0.3),
 * GammaTransform(gamma_range=(0.7, 1.3)), ContrastAugmentationTransform(contrast_range=(0.3, 1.7)) from the
 * third-party package `batchgenerators` (requirements.txt:7, no version pinned; NOT vendored in the reference and
 * not installed here -- the published algorithm is restated in oracle/augment.py, parity unpinned).
 * The random decisions are drawn on the host and passed in `coef`, one row of 2 + 2C floats per sample:
 *   [0]        gamma exponent, <= 0: no gamma step for this sample
 *   [1]        1: apply the contrast step, 0: skip it
 *   [2 .. 2+C) additive brightness offset per channel (0: none)
 *   [2+C .. )  contrast factor per channel
 * v = x + offset;  g = ((v - min v) / (range v + 1e-7)) ** gamma * range v + min v   (min / range over the SAMPLE);
 * y = clip((g - mean_c g) * factor_c + mean_c g, min_c g, max_c g)                   (statistics per CHANNEL).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* x;         /* [B, V, C] fp32 patches, NDHWC (V = P0*P1*P2 voxels)                  */
  void*        y;         /* [B, V, C] dst_dtype (MEDNET_F32 / MEDNET_BF16); must not overlap x   */
  const float* coef;      /* [B, 2 + 2C] device array, see above                                  */
  int64_t V;
  int32_t B, C, dst_dtype;
} mednet_intensity_aug_params;
size_t mednet_intensity_augment_workspace_bytes(const mednet_intensity_aug_params* p);
int mednet_intensity_augment(const mednet_intensity_aug_params* p, void* workspace, size_t workspace_bytes,
                             mednet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * tcgen05 addressing calibration (see DESIGN.md "UMMA descriptor experiments").
 * The conv kernel reads its 27 taps as shifted windows of one TMA-written halo tile, i.e. with UMMA
 * descriptors whose start address is not aligned to the swizzle atom.  The probe runs D = A_window * I
 * for such a window (row_bytes = 128 / 64 / 32 selects the swizzle width; K = N = row_bytes / 2) and
 * returns the rows the tensor core actually fetched; the Python side calls it once per process, picks a
 * variant that addresses correctly and registers it with mednet_tcgen05_configure.  Until a variant is
 * registered for a swizzle width the tensor-core path reports MEDNET_EUNSUPPORTED for it.
 *   dense_halo = 1 -> 10-voxel halo row pitch, one TMA box per plane; 0 -> rows padded to 16 voxels.
 *   base_offset_mode: 0 -> descriptor base_offset 0; 1 -> (start >> 7) & 7; 2 -> (start / row_bytes) & 7.
 * ---------------------------------------------------------------------------------------------- */
int mednet_tcgen05_configure(int row_bytes, int enabled, int dense_halo, int base_offset_mode);
/* Tuning switches of the tensor-core kernels (A/B measurements; every default is the measured-faster setting):
 *   "dual_issue" 0|1             second MMA-issuing thread for output tiles of <= 96 channels (1)
 *   "kd_merge" 0|1               plain conv: the kd taps of one (kh,kw) window in ONE wide-N MMA (1)
 *   "class_merge" 0|1            the same for the parity classes of MEDNET_GATHER_UPCONV_F/B (1)
 *   "ntile_max" 128|256          widest output-channel tile (128)
 *   "wgrad_dual_issue" 0|1       second MMA-issuing thread of the weight-gradient kernel (1)
 *   "wgrad_class_merge" 0|1      parity-class wgrad passes: all tap groups in one role, needed kw windows only (1)
 *   "wgrad_reduce_s_fastest" 0|1 thread order of the split reduction (0)
 *   "wgrad_wt_fastest", "wgrad_pair_planes", "wgrad_d_fastest" 0|1   round-1 switches (1)
 *   "first_layer_mma" 0|1        Cin = 1 forward / weight gradient on mma.sync (1)
 *   "conv_profile", "wgrad_profile" 0|1   per-CTA wait-cycle counters written behind the workspace (0)
 * Unknown names return MEDNET_EINVAL. */
int mednet_tcgen05_set_option(const char* name, int value);
int mednet_tcgen05_probe(const void* a_bf16 /* [rows][row_bytes/2] */, int32_t row_bytes, int32_t rows,
                         int32_t row_shift, int32_t sbo_bytes, int32_t base_offset_mode,
                         float* out /* [128][row_bytes/2] */, mednet_stream_t stream);

/* UMMA descriptor laboratory (diagnostics / calibration, tests/test_tcgen05_gpu.py): a [rows][row_bytes] image of
 * 16-bit elements is written to shared memory by TMA (swizzle = row_bytes), then `ksteps` tcgen05.mma
 * (M x N x 16, fp32 accumulate) are issued with the shared-memory descriptors described here -- byte offsets into
 * the image, leading/stride byte offsets, K- or MN-major, per-step address advance, operand formats
 * (0 = f16, 1 = bf16) -- and the 128 x N accumulator is returned. */
typedef struct {
  const void* g; float* out;
  int32_t rows, row_bytes, M, N, ksteps;
  int32_t a_off, a_lbo, a_sbo, a_mn_major, a_kstep;
  int32_t b_off, b_lbo, b_sbo, b_mn_major, b_kstep;
  int32_t a_fmt, b_fmt;
  int32_t iters;          /* repeat the k-step sequence (timing); 0 = once */
  int64_t* cycles;        /* optional device scalar: SM clocks from first issue to completion */
  int32_t nacc;           /* timing: rotate over this many independent accumulators (0 = 1); out = the first */
} mednet_umma_lab_params;
int mednet_umma_lab(const mednet_umma_lab_params* p, mednet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MEDNET_B200_H */
