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

}  // namespace milan
