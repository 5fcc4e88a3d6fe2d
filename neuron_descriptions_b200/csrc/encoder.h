// Launchers for the non-GEMM encoder kernels (encoder.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace milan {

// Per-image size of the 5-level mask pyramid: 112^2 + 56^2 + 28^2 + 14^2 + 7^2.
constexpr int kMaskPyramidSize = 12544 + 3136 + 784 + 196 + 49;
constexpr int kMaskLevelOffset[5] = {0, 12544, 15680, 16464, 16660};

// dtype: 0 = uint8, 1 = float32. images NCHW (n,3,224,224); masks (n,1,224,224).
// -> padded NHWC4 [n][232][232][4] (see conv_gemm.h build_stem_params)
// masks (same dtype, may be nullptr): multiply the normalised image by the mask (SpatialConvEncoder,
// src/milan/encoders.py:211); the pyramid encoder passes nullptr (it never masks the image, :298).
int launch_stem_pack(const void* images, int dtype, int n_images, __nv_bfloat16* p_hi, __nv_bfloat16* p_lo,
                     const float mean[3], const float stdv[3], int split, cudaStream_t stream,
                     const void* masks = nullptr);
// out[i] = hi[i] + lo[i] (lo may be nullptr): NHWC bf16 planes -> fp32 (SpatialConvEncoder output,
// src/milan/encoders.py:212-214: permute(0,2,3,1).reshape(n, 49, 512) is exactly the NHWC order).
int launch_planes_to_f32(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long n, float* out,
                         cudaStream_t stream);
int launch_mask_pyramid(const void* masks, int dtype, int n_images, float* wts, cudaStream_t stream);
int launch_masked_pool(const __nv_bfloat16* hi, const __nv_bfloat16* lo, const float* wts, int wts_stride,
                       int n_images, int P, int C, float* out, int out_stride, cudaStream_t stream);
int launch_bn_relu_maxpool(const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, const float* alpha,
                           const float* beta, int n_images, __nv_bfloat16* y_hi, __nv_bfloat16* y_lo,
                           cudaStream_t stream);


// ---- AlexNet pyramid (PyramidConvEncoder('alexnet'), src/milan/encoders.py:328-334)
constexpr int kAlexK0 = 384;  // 3*11*11 = 363 im2col columns of features.0, zero-padded to a multiple of 64
// Normalise + im2col of the 11x11 stride-4 pad-2 first convolution: A[n*55*55 + oh*55 + ow][(c*11 + r)*11 + s].
int launch_alexnet_im2col(const void* images, int dtype, int n_images, __nv_bfloat16* a_hi, __nv_bfloat16* a_lo,
                          const float mean[3], const float stdv[3], int split, cudaStream_t stream);
// Bilinear (align_corners=False, no antialias) resize of (n,1,224,224) masks to SxS for any S, then the
// per-image sum-normalisation with the all-zero exception (src/milan/encoders.py:303-314). out: [n][stride].
int launch_mask_resize(const void* masks, int dtype, int n_images, int S, float* out, int out_stride,
                       cudaStream_t stream);
// 3x3 stride-2 max-pool without padding on NHWC hi/lo planes (torchvision alexnet features.2 / .5); C % 8 == 0.
int launch_maxpool3x3s2(const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, int n_images, int H, int W, int C,
                        __nv_bfloat16* y_hi, __nv_bfloat16* y_lo, cudaStream_t stream);

}  // namespace milan
