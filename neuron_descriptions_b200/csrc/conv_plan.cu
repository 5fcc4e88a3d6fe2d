// Host-side planning for conv_gemm: tensor maps, M-tile boxes and filter-tap tables per convolution.
#include "conv_gemm.h"

#include <cstdio>
#include <cstring>

namespace milan {

// Chained tcgen05 accumulations per TMEM partial (see conv_gemm.cu): 4 k-blocks x 12 MMAs for the encoder convs,
// 2 k-blocks for the decoder / LM GEMMs whose log-probabilities are compared at 1e-3.
constexpr int kEncoderKbPerChunk = 4;
constexpr int kDecoderKbPerChunk = 2;

int build_conv_params(ConvGemmParams* p, const ConvDesc& d, const ConvIO& io, int split, int* block_n) {
  memset(p, 0, sizeof(*p));
  if (d.Cin % kGemmBlockK != 0) return -2;
  if (!(d.ksize == 1 || d.ksize == 3 || d.ksize == 5) || !(d.stride == 1 || d.stride == 2)) return -3;
  if (d.stride == 2 && ((d.H | d.W) & 1)) return -4;
  const int Ho = d.H / d.stride, Wo = d.W / d.stride;
  const int bn_tile = (d.Cout <= 64) ? 64 : 128;
  *block_n = bn_tile;
  const uint64_t esz = 2;
  const int nplanes_hi_lo = split ? 2 : 1;
  const __nv_bfloat16* in_planes[2] = {io.in_hi, io.in_lo};
  const __nv_bfloat16* w_planes[2] = {io.w_hi, io.w_lo};

  p->cin = d.Cin;
  p->cout = d.Cout;
  p->n_tiles = (d.Cout + bn_tile - 1) / bn_tile;
  p->bias = io.bias;
  p->res_hi = io.res_hi;
  p->res_lo = io.res_lo;
  p->out_hi = io.out_hi;
  p->out_lo = io.out_lo;
  p->out_f32 = io.out_f32;
  p->ldc = io.ldc > 0 ? io.ldc : d.Cout;
  p->relu = io.relu;

  int rc = 0;
  if (d.ksize == 1 && d.stride == 1) {
    // Flat GEMM: rows = all pixels.
    const long long M = static_cast<long long>(d.N) * d.H * d.W;
    p->box_w = kGemmBlockM; p->box_h = 1; p->box_n = 1;
    p->tiles_w = static_cast<int>((M + kGemmBlockM - 1) / kGemmBlockM);
    p->tiles_h = 1; p->tiles_n = 1;
    p->out_w = static_cast<int>(M); p->out_h = 1; p->out_n = 1;
    p->num_taps = 1;
    p->tap_plane[0] = 0; p->tap_dw[0] = 0; p->tap_dh[0] = 0;
    for (int hl = 0; hl < nplanes_hi_lo; ++hl) {
      const uint64_t pitch = static_cast<uint64_t>(d.Cin) * esz;
      rc = make_tmap_4d(&p->tmap_a[hl][0], in_planes[hl], d.Cin, M, 1, 1, pitch, pitch * M, pitch * M,
                        kGemmBlockM, 1, 1);
      if (rc) return rc;
      for (int pl = 1; pl < 4; ++pl) p->tmap_a[hl][pl] = p->tmap_a[hl][0];
    }
  } else {
    int bw, bh, bn;
    choose_box(Wo, Ho, d.N, &bw, &bh, &bn);
    p->box_w = bw; p->box_h = bh; p->box_n = bn;
    p->tiles_w = (Wo + bw - 1) / bw;
    p->tiles_h = (Ho + bh - 1) / bh;
    p->tiles_n = (d.N + bn - 1) / bn;
    p->out_w = Wo; p->out_h = Ho; p->out_n = d.N;
    const uint64_t cpitch = static_cast<uint64_t>(d.Cin) * esz;
    const uint64_t img = cpitch * d.W * d.H;
    if (d.stride == 1) {
      for (int hl = 0; hl < nplanes_hi_lo; ++hl) {
        rc = make_tmap_4d(&p->tmap_a[hl][0], in_planes[hl], d.Cin, d.W, d.H, d.N, cpitch, cpitch * d.W, img, bw, bh, bn);
        if (rc) return rc;
        for (int pl = 1; pl < 4; ++pl) p->tmap_a[hl][pl] = p->tmap_a[hl][0];
      }
    } else {
      // Parity planes: pixel (h, w) = (2*h2 + ph, 2*w2 + pw) lives in plane ph*2+pw at (h2, w2).
      for (int hl = 0; hl < nplanes_hi_lo; ++hl) {
        for (int ph = 0; ph < 2; ++ph) {
          for (int pw = 0; pw < 2; ++pw) {
            const uint8_t* base = reinterpret_cast<const uint8_t*>(in_planes[hl]) +
                                  (static_cast<uint64_t>(ph) * d.W + pw) * cpitch;
            rc = make_tmap_4d(&p->tmap_a[hl][ph * 2 + pw], base, d.Cin, d.W / 2, d.H / 2, d.N, 2 * cpitch,
                              2 * cpitch * d.W, img, bw, bh, bn);
            if (rc) return rc;
          }
        }
      }
    }
    int t = 0;
    for (int r = 0; r < d.ksize; ++r) {
      for (int s = 0; s < d.ksize; ++s, ++t) {
        const int oh = r - d.ksize / 2, ow = s - d.ksize / 2;  // input offset relative to stride*out
        if (d.stride == 1) {
          p->tap_plane[t] = 0; p->tap_dh[t] = static_cast<int8_t>(oh); p->tap_dw[t] = static_cast<int8_t>(ow);
        } else {
          // input row = 2*out + oh  ->  parity = oh & 1, plane row = out + floor(oh / 2)
          const int ph = oh & 1, pw = ow & 1;
          const int fh = (oh - ph) / 2, fw = (ow - pw) / 2;
          p->tap_plane[t] = static_cast<int8_t>(ph * 2 + pw);
          p->tap_dh[t] = static_cast<int8_t>(fh);
          p->tap_dw[t] = static_cast<int8_t>(fw);
        }
      }
    }
    p->num_taps = t;
  }
  p->a_box_bytes = static_cast<uint32_t>(p->box_w) * p->box_h * p->box_n * kGemmBlockK * 2;
  p->kb_per_chunk = split ? kEncoderKbPerChunk : 0;
  if (io.out_hi != nullptr) {
    // Output (and residual) tensors [N][Ho][Wo][Cout] addressed with the same M-tile boxes, 64 channels wide.
    const __nv_bfloat16* outs[2] = {io.out_hi, io.out_lo};
    const __nv_bfloat16* ress[2] = {io.res_hi, io.res_lo};
    const uint64_t opitch = static_cast<uint64_t>(p->ldc) * esz;
    p->has_res = io.res_hi != nullptr ? 1 : 0;
    for (int hl = 0; hl < nplanes_hi_lo; ++hl) {
      for (int which = 0; which < 2; ++which) {
        const __nv_bfloat16* base = which == 0 ? outs[hl] : ress[hl];
        if (base == nullptr) continue;
        CUtensorMap* tm = which == 0 ? &p->tmap_out[hl] : &p->tmap_res[hl];
        if (d.ksize == 1 && d.stride == 1) {
          const uint64_t M = static_cast<uint64_t>(p->out_w);
          rc = make_tmap_4d(tm, base, d.Cout, M, 1, 1, opitch, opitch * M, opitch * M, kGemmBlockM, 1, 1);
        } else {
          rc = make_tmap_4d(tm, base, d.Cout, p->out_w, p->out_h, p->out_n, opitch, opitch * p->out_w,
                            opitch * p->out_w * p->out_h, p->box_w, p->box_h, p->box_n);
        }
        if (rc) return rc;
      }
    }
    if (!split) {
      p->tmap_out[1] = p->tmap_out[0];
      p->tmap_res[1] = p->tmap_res[0];
    }
  }
  const uint64_t ktot = static_cast<uint64_t>(p->num_taps) * d.Cin;
  for (int hl = 0; hl < nplanes_hi_lo; ++hl) {
    rc = make_tmap_2d(&p->tmap_b[hl], w_planes[hl], ktot, d.Cout, ktot * esz, bn_tile);
    if (rc == 0 && bn_tile == 128) {
      rc = make_tmap_2d(&p->tmap_b_half[hl], w_planes[hl], ktot, d.Cout, ktot * esz, bn_tile / 2);
      p->has_b_half = 1;
    }
    if (rc) return rc;
  }
  if (!split) {
    p->tmap_b[1] = p->tmap_b[0];
    for (int pl = 0; pl < 4; ++pl) p->tmap_a[1][pl] = p->tmap_a[0][pl];
  }
  return 0;
}

int build_dual_1x1_params(ConvGemmParams* p, int N, int Ho, int Wo, int Ca, int Cb, int Cout, int stride,
                          const DualConvIO& io, int split, int* block_n) {
  memset(p, 0, sizeof(*p));
  if (Ca % kGemmBlockK != 0 || Cb % kGemmBlockK != 0) return -2;
  if (!(stride == 1 || stride == 2)) return -3;
  const int bn_tile = (Cout <= 64) ? 64 : 128;
  *block_n = bn_tile;
  const uint64_t esz = 2;
  const int np = split ? 2 : 1;
  const __nv_bfloat16* a_planes[2] = {io.a_hi, io.a_lo};
  const __nv_bfloat16* b_planes[2] = {io.b_hi, io.b_lo};
  const __nv_bfloat16* w_planes[2] = {io.w_hi, io.w_lo};
  const __nv_bfloat16* outs[2] = {io.out_hi, io.out_lo};
  p->cin = Ca;  // default depth; tap 1 overrides it through tap_cb
  p->cout = Cout;
  p->n_tiles = (Cout + bn_tile - 1) / bn_tile;
  p->bias = io.bias;
  p->out_hi = io.out_hi;
  p->out_lo = io.out_lo;
  p->ldc = Cout;
  p->relu = io.relu;
  p->num_taps = 2;
  p->tap_plane[0] = 0; p->tap_plane[1] = 1;
  p->tap_cb[0] = static_cast<int16_t>(Ca / kGemmBlockK);
  p->tap_cb[1] = static_cast<int16_t>(Cb / kGemmBlockK);
  p->kb_per_chunk = split ? kEncoderKbPerChunk : 0;
  int rc = 0;
  if (stride == 1) {  // both operands are flat [M][C] matrices
    const uint64_t M = static_cast<uint64_t>(N) * Ho * Wo;
    p->box_w = kGemmBlockM; p->box_h = 1; p->box_n = 1;
    p->tiles_w = static_cast<int>((M + kGemmBlockM - 1) / kGemmBlockM);
    p->tiles_h = 1; p->tiles_n = 1;
    p->out_w = static_cast<int>(M); p->out_h = 1; p->out_n = 1;
    for (int hl = 0; hl < np; ++hl) {
      const uint64_t pa = static_cast<uint64_t>(Ca) * esz, pb = static_cast<uint64_t>(Cb) * esz, po = static_cast<uint64_t>(Cout) * esz;
      if ((rc = make_tmap_4d(&p->tmap_a[hl][0], a_planes[hl], Ca, M, 1, 1, pa, pa * M, pa * M, kGemmBlockM, 1, 1))) return rc;
      if ((rc = make_tmap_4d(&p->tmap_a[hl][1], b_planes[hl], Cb, M, 1, 1, pb, pb * M, pb * M, kGemmBlockM, 1, 1))) return rc;
      if ((rc = make_tmap_4d(&p->tmap_out[hl], outs[hl], Cout, M, 1, 1, po, po * M, po * M, kGemmBlockM, 1, 1))) return rc;
    }
  } else {  // 4-D boxes over the output raster; b is read through its (even row, even column) parity plane
    int bw, bh, bn;
    choose_box(Wo, Ho, N, &bw, &bh, &bn);
    p->box_w = bw; p->box_h = bh; p->box_n = bn;
    p->tiles_w = (Wo + bw - 1) / bw;
    p->tiles_h = (Ho + bh - 1) / bh;
    p->tiles_n = (N + bn - 1) / bn;
    p->out_w = Wo; p->out_h = Ho; p->out_n = N;
    const int Hi = Ho * 2, Wi = Wo * 2;
    for (int hl = 0; hl < np; ++hl) {
      const uint64_t pa = static_cast<uint64_t>(Ca) * esz, pb = static_cast<uint64_t>(Cb) * esz, po = static_cast<uint64_t>(Cout) * esz;
      if ((rc = make_tmap_4d(&p->tmap_a[hl][0], a_planes[hl], Ca, Wo, Ho, N, pa, pa * Wo, pa * Wo * Ho, bw, bh, bn))) return rc;
      if ((rc = make_tmap_4d(&p->tmap_a[hl][1], b_planes[hl], Cb, Wo, Ho, N, 2 * pb, 2 * pb * Wi, pb * Wi * Hi, bw, bh, bn))) return rc;
      if ((rc = make_tmap_4d(&p->tmap_out[hl], outs[hl], Cout, Wo, Ho, N, po, po * Wo, po * Wo * Ho, bw, bh, bn))) return rc;
    }
  }
  for (int hl = 0; hl < np; ++hl) {
    p->tmap_a[hl][2] = p->tmap_a[hl][0];
    p->tmap_a[hl][3] = p->tmap_a[hl][0];
    const uint64_t ktot = static_cast<uint64_t>(Ca) + Cb;
    if ((rc = make_tmap_2d(&p->tmap_b[hl], w_planes[hl], ktot, Cout, ktot * esz, bn_tile))) return rc;
  }
  p->a_box_bytes = static_cast<uint32_t>(p->box_w) * p->box_h * p->box_n * kGemmBlockK * 2;
  if (!split) {
    p->tmap_b[1] = p->tmap_b[0];
    p->tmap_out[1] = p->tmap_out[0];
    for (int pl = 0; pl < 4; ++pl) p->tmap_a[1][pl] = p->tmap_a[0][pl];
  }
  p->tmap_res[0] = p->tmap_out[0];
  p->tmap_res[1] = p->tmap_out[1];
  return 0;
}

void pack_stem_weights(const float* w, float* packed) {
  // k = r*32 + sp*4 + c  <-  w[co][c][r][s = sp - 1]; zero for sp = 0 and c = 3.
  for (int co = 0; co < 64; ++co) {
    for (int k = 0; k < kStemKTotal; ++k) {
      const int r = k / kStemBlockK, sp = (k / 4) & 7, c = k & 3;
      const int s = sp - 1;
      float v = 0.f;
      if (s >= 0 && s < 7 && c < 3) v = w[((co * 3 + c) * 7 + r) * 7 + s];
      packed[co * kStemKTotal + k] = v;
    }
  }
}

int build_stem_params(ConvGemmParams* p, int N, const __nv_bfloat16* img_hi, const __nv_bfloat16* img_lo,
                      const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, __nv_bfloat16* out_hi,
                      __nv_bfloat16* out_lo, int split) {
  memset(p, 0, sizeof(*p));
  const int O = 112;
  p->stem_mode = 1;
  p->block_k = kStemBlockK;
  p->kb_per_chunk = split ? kEncoderKbPerChunk : 0;
  p->cin = kStemBlockK;   // one 64-byte k-block per filter row
  p->cout = 64;
  p->n_tiles = 1;
  p->ldc = 64;
  p->num_taps = 7;
  for (int r = 0; r < 7; ++r) {  // padded row 2*oh + r = 2*(oh + r/2) + (r & 1)
    p->tap_plane[r] = 0;
    p->tap_dw[r] = static_cast<int8_t>(r & 1);   // row parity coordinate
    p->tap_dh[r] = static_cast<int8_t>(r >> 1);  // row-pair offset
  }
  // 8 x 16 output pixels per tile: one 8-row MMA group per image row, so the raw row segments the filter row reads
  // (8 * 2 + 6 = 22 pixels = 176 B) sit at a constant pitch in shared memory (make_smem_desc_stem_rows).
  const int bw = 8, bh = 16, bn = 1;
  p->box_w = bw; p->box_h = bh; p->box_n = bn;
  p->tiles_w = (O + bw - 1) / bw; p->tiles_h = (O + bh - 1) / bh; p->tiles_n = (N + bn - 1) / bn;
  p->out_w = O; p->out_h = O; p->out_n = N;
  const uint32_t seg_elems = (2 * bw + 6) * 4;  // 88 bf16 = 176 B
  // one box per tile and plane: both parities of the bh + 3 row pairs the seven filter rows reach = padded rows
  // 2 h0 .. 2 h0 + 37 in order (the kernel's descriptors start filter row r at segment r, image rows 2 segments apart)
  const uint32_t row_pairs = bh + 3;
  p->a_box_bytes = seg_elems * 2 * 2 * row_pairs;
  const uint64_t P = static_cast<uint64_t>(kStemPadW) * 4 * 2;  // padded row pitch in bytes
  const __nv_bfloat16* imgs[2] = {img_hi, img_lo};
  const __nv_bfloat16* ws[2] = {w_hi, w_lo};
  __nv_bfloat16* outs[2] = {out_hi, out_lo};
  const int np = split ? 2 : 1;
  for (int hl = 0; hl < np; ++hl) {
    // dims: (element of a padded row: pixel * 4 + channel, row parity, row pair, n); no swizzle
    const uint64_t dims[4] = {static_cast<uint64_t>(kStemPadW) * 4, 2, static_cast<uint64_t>(kStemPadH / 2),
                              static_cast<uint64_t>(N)};
    const uint64_t strides[3] = {P, 2 * P, static_cast<uint64_t>(kStemPadH) * P};
    const uint32_t box[4] = {seg_elems, 2, row_pairs, static_cast<uint32_t>(bn)};
    int rc = make_tmap_nd(&p->tmap_a[hl][0], imgs[hl], 4, dims, strides, box, 0);
    if (rc) return rc;
    for (int pl = 1; pl < 4; ++pl) p->tmap_a[hl][pl] = p->tmap_a[hl][0];
    rc = make_tmap_2d(&p->tmap_b[hl], ws[hl], kStemKTotal, 64, kStemKTotal * 2, 64, kStemBlockK);
    if (rc) return rc;
    const uint64_t opitch = 64 * 2;
    rc = make_tmap_4d(&p->tmap_out[hl], outs[hl], 64, O, O, N, opitch, opitch * O, opitch * O * O, bw, bh, bn);
    if (rc) return rc;
  }
  if (!split) {
    p->tmap_b[1] = p->tmap_b[0];
    p->tmap_out[1] = p->tmap_out[0];
    for (int pl = 0; pl < 4; ++pl) p->tmap_a[1][pl] = p->tmap_a[0][pl];
  }
  p->out_hi = out_hi; p->out_lo = out_lo;
  return 0;
}

int build_gemm_params(ConvGemmParams* p, long long M, int K, int N, const __nv_bfloat16* a_hi,
                      const __nv_bfloat16* a_lo, long long a_pitch, const __nv_bfloat16* w_hi,
                      const __nv_bfloat16* w_lo, const float* bias, float* out, long long ldc, int split,
                      int fp16_operands) {
  memset(p, 0, sizeof(*p));
  if (K % kGemmBlockK != 0 || M <= 0) return -2;
  p->fp16_operands = fp16_operands;
  p->kb_per_chunk = split ? kDecoderKbPerChunk : 0;
  const int bn_tile = 128;
  p->cin = K;
  p->cout = N;
  p->n_tiles = (N + bn_tile - 1) / bn_tile;
  p->bias = bias;
  p->out_f32 = out;
  p->ldc = ldc;
  p->box_w = kGemmBlockM; p->box_h = 1; p->box_n = 1;
  p->tiles_w = static_cast<int>((M + kGemmBlockM - 1) / kGemmBlockM);
  p->tiles_h = 1; p->tiles_n = 1;
  p->out_w = static_cast<int>(M); p->out_h = 1; p->out_n = 1;
  p->num_taps = 1;
  p->a_box_bytes = kGemmBlockM * kGemmBlockK * 2;
  const __nv_bfloat16* a_planes[2] = {a_hi, a_lo};
  const __nv_bfloat16* w_planes[2] = {w_hi, w_lo};
  const int np = split ? 2 : 1;
  for (int hl = 0; hl < np; ++hl) {
    const uint64_t pitch = static_cast<uint64_t>(a_pitch) * 2;
    int rc = make_tmap_4d(&p->tmap_a[hl][0], a_planes[hl], K, M, 1, 1, pitch, pitch * M, pitch * M, kGemmBlockM, 1, 1);
    if (rc) return rc;
    for (int pl = 1; pl < 4; ++pl) p->tmap_a[hl][pl] = p->tmap_a[hl][0];
    rc = make_tmap_2d(&p->tmap_b[hl], w_planes[hl], K, N, static_cast<uint64_t>(K) * 2, bn_tile);
    if (rc) return rc;
  }
  if (!split) {
    p->tmap_b[1] = p->tmap_b[0];
    for (int pl = 0; pl < 4; ++pl) p->tmap_a[1][pl] = p->tmap_a[0][pl];
  }
  return 0;
}

int build_gemm2_params(ConvGemmParams* p, long long M, int K0, int K1, int N, const __nv_bfloat16* a0_hi,
                       const __nv_bfloat16* a0_lo, long long a0_pitch, const __nv_bfloat16* a1_hi,
                       const __nv_bfloat16* a1_lo, long long a1_pitch, const __nv_bfloat16* w_hi,
                       const __nv_bfloat16* w_lo, const float* bias, int split, int fp16_operands) {
  if (K1 % kGemmBlockK != 0) return -2;
  int rc = build_gemm_params(p, M, K0, N, a0_hi, a0_lo, a0_pitch, w_hi, w_lo, bias, nullptr, N, split, fp16_operands);
  if (rc) return rc;
  // two "taps" of different depth, each with its own tensor map (the mechanism of build_dual_1x1_params)
  p->num_taps = 2;
  p->tap_plane[0] = 0; p->tap_plane[1] = 1;
  p->tap_cb[0] = static_cast<int16_t>(K0 / kGemmBlockK);
  p->tap_cb[1] = static_cast<int16_t>(K1 / kGemmBlockK);
  const __nv_bfloat16* a1_planes[2] = {a1_hi, a1_lo};
  const __nv_bfloat16* w_planes[2] = {w_hi, w_lo};
  const int np = split ? 2 : 1;
  const uint64_t ktot = static_cast<uint64_t>(K0) + K1;
  for (int hl = 0; hl < np; ++hl) {
    const uint64_t pitch = static_cast<uint64_t>(a1_pitch) * 2;
    rc = make_tmap_4d(&p->tmap_a[hl][1], a1_planes[hl], K1, M, 1, 1, pitch, pitch * M, pitch * M, kGemmBlockM, 1, 1);
    if (rc) return rc;
    rc = make_tmap_2d(&p->tmap_b[hl], w_planes[hl], ktot, N, ktot * 2, 128);
    if (rc) return rc;
  }
  if (!split) {
    p->tmap_b[1] = p->tmap_b[0];
    p->tmap_a[1][1] = p->tmap_a[0][1];
  }
  return 0;
}

}  // namespace milan
