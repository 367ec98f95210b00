/* mcb200.h — C ABI of libmcb200.so: the B200 (sm_100a) hot path of neptune-ai/open-solution-mapping-challenge.
 *
 * The reference is 100 % Python; every FLOP of its hot path runs inside torch 0.3.1 / cuDNN / scipy / skimage /
 * pydensecrf calls.  This header is what a Python host binds (ctypes) in place of those library
 * calls.  Each entry point cites the reference call site it replaces as  file:line  under /root/reference.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MCB_ERR_* code on failure; mcb_last_error() returns a
 *     thread-local message (the Python layer raises RuntimeError with it, matching the reference's exceptions);
 *   - the caller owns all memory: device pointers allocated by the host framework; the library never allocates or
 *     frees device memory and keeps no pointer after returning;
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); no hidden synchronisation;
 *   - activations are NHWC bf16, dense; conv weights are bf16 [ky][kx][cout][cin] ("tap-major");
 *     weight gradients are fp32 in the same layout; vectors (bias, BN parameters, statistics) are fp32;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with MCB_ERR_CUDA.
 */
#ifndef MCB200_H_
#define MCB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB_OK 0
#define MCB_ERR_INVALID (-1)
#define MCB_ERR_CUDA (-2)
#define MCB_ERR_UNSUPPORTED (-3)

const char* mcb_last_error(void);
int mcb_version(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Convolutions as tcgen05 implicit GEMMs (TMA-fed, TMEM accumulators).
 * Replaces nn.Conv2d / nn.ConvTranspose2d forward + autograd backward at
 *   src/unet_models.py:21-34 (conv3x3, ConvRelu), :125-150 (DecoderBlockV2), :360-383 (encoder stages, final),
 *   torchvision/models/resnet.py BasicBlock / Bottleneck (used at src/unet_models.py:344-352).
 * ---------------------------------------------------------------------------------------------------------------- */

typedef struct {
  const void* x[2];   /* NHWC bf16 inputs; x[1] != NULL fuses torch.cat([x0, x1], 1) (src/unet_models.py:395-399) */
  int cin[2];         /* channels of x[0], x[1] (32, or multiples of 64) */
  int n, h, w;        /* input batch, height, width */
  const void* weight; /* bf16 [ksize*ksize][cout][cin0+cin1] */
  int cout;
  int ksize;          /* 1 or 3; padding = ksize / 2 */
  int stride;         /* 1 or 2 */
  const float* bias;  /* fp32 [cout] or NULL */
  int relu;           /* fuse ReLU (ConvRelu, src/unet_models.py:25-34) */
  float* stats;       /* fp32 [2*cout] or NULL: += per-channel sum / sum of squares of the stored (bf16) outputs:
                         the BatchNorm batch statistics of the layer that follows */
  void* y;            /* NHWC bf16 [n][h/stride][w/stride][cout] */
} mcb_conv_fwd_args;
int mcb_conv_fwd(const mcb_conv_fwd_args* a, void* stream);

typedef struct {
  const void* dy;        /* NHWC bf16 [n][h/stride][w/stride][cout] */
  int n, h, w;           /* dims of dx (the conv input) */
  const void* weight;    /* bf16 [ksize*ksize][cout][cin_total] */
  int cout, cin_total;
  int ci_off, cin;       /* dx covers weight input channels [ci_off, ci_off+cin) (one source of a fused concat) */
  int ksize, stride;
  void* dx;              /* NHWC bf16 [n][h][w][cin] */
  const void* relu_mask; /* NHWC bf16 like dx or NULL: dx is zeroed where relu_mask <= 0 (backward of the ReLU that
                            produced the conv input) */
  int accumulate;        /* dx += (TMA reduce-add) instead of dx = */
} mcb_conv_dgrad_args;
int mcb_conv_dgrad(const mcb_conv_dgrad_args* a, void* stream);

typedef struct {
  const void* dy;     /* NHWC bf16 [n][h/stride][w/stride][cout] */
  const void* x;      /* NHWC bf16 [n][h][w][cin] */
  int n, h, w;
  int cout, cin_total;
  int ci_off, cin;    /* x supplies weight input channels [ci_off, ci_off+cin) */
  int ksize, stride;
  float* dw;          /* fp32 [ksize*ksize][cout][cin_total], accumulated (+=); zero it first */
} mcb_conv_wgrad_args;
int mcb_conv_wgrad(const mcb_conv_wgrad_args* a, void* stream);

/* nn.ConvTranspose2d(kernel_size=4, stride=2, padding=1) (src/unet_models.py:138-139) as four sub-pixel phases */
typedef struct {
  const void* x;      /* NHWC bf16 [n][h][w][cin] */
  int n, h, w, cin;
  const void* weight; /* bf16 [16][cout][cin] */
  int cout;
  const float* bias;
  int relu;
  void* y;            /* NHWC bf16 [n][2h][2w][cout] */
} mcb_convt_fwd_args;
int mcb_convt_fwd(const mcb_convt_fwd_args* a, void* stream);

typedef struct {
  const void* dy;     /* NHWC bf16 [n][2h][2w][cout] */
  int n, h, w, cin;
  const void* weight; /* bf16 [16][cout][cin] */
  int cout;
  void* dx;           /* NHWC bf16 [n][h][w][cin] */
  const void* relu_mask;
  int accumulate;
} mcb_convt_dgrad_args;
int mcb_convt_dgrad(const mcb_convt_dgrad_args* a, void* stream);

typedef struct {
  const void* dy;     /* NHWC bf16 [n][2h][2w][cout] */
  const void* x;      /* NHWC bf16 [n][h][w][cin] */
  int n, h, w, cin, cout;
  float* dw;          /* fp32 [16][cout][cin], accumulated */
} mcb_convt_wgrad_args;
int mcb_convt_wgrad(const mcb_convt_wgrad_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCB200_H_ */
