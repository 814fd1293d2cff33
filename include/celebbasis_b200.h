/*
 * celebbasis_b200 — C-ABI of the B200-native (sm_100a) hot path of CelebBasis.
 *
 * The reference (ygtxr1997/CelebBasis) has no FFI: its hot path is a chain of stock ATen calls
 * made from Python.  This header is therefore the boundary a maintainer would
 * bind with ctypes from the reference's own modules; every entry point names the reference
 * code whose arithmetic it replaces (file:line relative to the reference tree).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name says `host`; the caller (PyTorch's caching
 *     allocator in our host mirror) owns all buffers including workspaces; nothing here allocates.
 *   - all work is enqueued on the cudaStream_t passed in (as void*); no hidden synchronisation;
 *     every entry point is CUDA-graph capturable and re-entrant.
 *   - return value: 0 = ok, negative = argument check failed, positive = cudaError_t.
 *     cb_last_error() returns a thread-local message for the last non-zero return.
 *   - activations are channels-last: images are NHWC ([N][H][W][C], C contiguous), token
 *     matrices are [rows][channels].  Module inputs/outputs in the Python mirror stay NCHW fp32
 *     exactly like the reference (ldm/models/diffusion/ddpm.py:344-350).
 */
#ifndef CELEBBASIS_B200_H_
#define CELEBBASIS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_ABI_VERSION 1

/* element types */
enum { CB_F16 = 0, CB_BF16 = 1, CB_F32 = 2 };
/* activation fused into the GEMM epilogue */
enum { CB_ACT_NONE = 0, CB_ACT_SILU = 1, CB_ACT_GELU = 2, CB_ACT_QUICK_GELU = 3 };
/* operand majorness: K-major = reduction dim contiguous; MN-major = M (or N) contiguous */
enum { CB_MAJOR_K = 0, CB_MAJOR_MN = 1 };

/* error codes (negative) */
enum {
    CB_OK = 0,
    CB_ERR_ARG = -1,       /* invalid argument / unsupported shape */
    CB_ERR_ALIGN = -2,     /* pointer or stride alignment */
    CB_ERR_DRIVER = -3,    /* CUDA driver entry point unavailable / tensor-map encode failed */
    CB_ERR_NO_DEVICE = -4, /* no sm_100 device */
};

int cb_abi_version(void);
const char* cb_last_error(void);
/* 1 if the current device is compute capability 10.x, else 0 (never throws). */
int cb_device_ok(void);

/* ---------------------------------------------------------------------------------------------
 * cb_gemm — tcgen05 tensor-core GEMM / implicit-GEMM convolution with TMA-staged operands.
 *
 *   D[b][m][n] = act(alpha * sum_{tap,k} A_tap[b][m][k] * B[b][tap][n][k] + bias[...]) + R[b][m][n]
 *
 * Replaces every torch.nn.Linear / Conv2d / einsum contraction on the path:
 *   ldm/modules/attention.py:170-193 (to_q/k/v/out, q.k^T, p.v), :37-64 (GEGLU / FF),
 *   ldm/modules/diffusionmodules/openaimodel.py:255-275 (ResBlock convs), :91-160 (up/down convs),
 *   ldm/modules/diffusionmodules/model.py:82-202 (VAE ResnetBlock/AttnBlock),
 *   ldm/modules/id_embedding/iresnet.py:26-64 (IBasicBlock convs),
 *   transformers CLIPAttention/CLIPMLP called from ldm/modules/encoders/modules.py:320-340.
 * and, with operand majors swapped, their activation-gradient (dgrad) counterparts that autograd
 * runs in the reference.
 *
 * Plain mode (conv == 0): A is [batch][M][K] (K-major, row stride lda) or, with
 * a_major == CB_MAJOR_MN, stored transposed as [batch][K][M] (row stride lda).  B likewise is
 * [batch][N][K] (K-major) or [batch][K][N] (MN-major).  Strides are in ELEMENTS.
 *
 * Conv mode (conv == 1): A is an NHWC image [img_n][img_h][img_w][K] (K = input channels, K-major)
 * and row m of the GEMM is output pixel (img, oh, ow) in raster order; tap (r,s) reads input pixel
 * (oh*stride + r - pad_top, ow*stride + s - pad_left); out-of-image taps read zeros (TMA OOB fill).
 * B rows for tap t start at row t*b_tap_rows: K-major B is [taps*b_tap_rows][K]; MN-major B is
 * [taps*b_tap_rows (k index)][N] (used for dgrad with the forward weight pack).  M is implied:
 * M = img_n*out_h*out_w.  tap order is r-major (t = r*kw + s); flip_taps reverses it (dgrad).
 * ------------------------------------------------------------------------------------------- */
typedef struct cb_gemm_desc {
    int32_t M, N, K;
    int32_t batch;
    int32_t ab_dtype; /* CB_F16 or CB_BF16 (both operands) */

    const void* A;
    int64_t lda;
    int64_t a_batch_stride;
    int32_t a_major;

    const void* B;
    int64_t ldb;
    int64_t b_batch_stride;
    int32_t b_major;

    /* conv mode */
    int32_t conv;
    int32_t img_n, img_h, img_w;
    int32_t out_h, out_w;
    int32_t kh, kw;
    int32_t stride;
    int32_t pad_top, pad_left;
    int32_t b_tap_rows;
    int32_t flip_taps;

    /* epilogue */
    void* D;
    int32_t d_dtype; /* CB_F16 / CB_BF16 / CB_F32 */
    int64_t ldd;
    int64_t d_batch_stride;
    int32_t d_transposed; /* write D[b][n][m] (row stride ldd) instead of D[b][m][n] */

    const float* bias;     /* fp32 [bias_rows][N] or NULL */
    int32_t bias_row_div;  /* bias row = (global row m) / bias_row_div; 0 => single row */
    int64_t ldbias;

    const void* R; /* residual added after the activation, same logical shape as D, or NULL */
    int32_t r_dtype;
    int64_t ldr;
    int64_t r_batch_stride;

    float alpha;
    int32_t act;
} cb_gemm_desc;

int cb_gemm(const cb_gemm_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CELEBBASIS_B200_H_ */
