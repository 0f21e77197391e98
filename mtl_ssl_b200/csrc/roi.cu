// HBM-bound feature-map kernels (sm_100a): ROI bilinear gather/scatter (tf.image.crop_and_resize),
// position-sensitive ROI pooling, max / average pooling, stem im2col + preprocessing.
//
// Reference call sites (under /root/reference/):
//   crop_and_resize   object_detection/meta_architectures/faster_rcnn_meta_arch.py:1340-1348
//   PS-ROI            object_detection/utils/ops.py:462-609
//   max pool          slim/nets/resnet_v1.py:222, slim/nets/resnet_utils.py:59-74, fmA:1345-1348
//   spatial average   object_detection/core/box_predictor.py:470-472
//   preprocessing     object_detection/models/faster_rcnn_resnet_v1_feature_extractor.py:74-90
// The arithmetic of CropAndResize / MaxPool / ResizeBilinear is TensorFlow 1.7's (not vendored);
// it is restated in oracle/nn.py and these kernels follow the same formulas.
// All activations are NHWC; channel vectors are moved as 16-byte (8 x bf16) accesses, one
// warp per output pixel, so every global transaction is a full coalesced line.
#include "common.cuh"

namespace {

__device__ __forceinline__ void unpack8(const uint4 v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    f[2 * q] = __uint_as_float(w[q] << 16);
    f[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
    w[q] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------ crop_and_resize
struct CropGeom {
  bool valid;
  int top, bot, left, right;
  float yl, xl;
};

// TF 1.7 crop_and_resize_op.cc coordinate arithmetic for output pixel (i, j) of box (y1,x1,y2,x2).
__device__ __forceinline__ CropGeom crop_geom(const float4 bx, int i, int j, int ch, int cw, int H, int W) {
  CropGeom g;
  const float hs = ch > 1 ? (bx.z - bx.x) * (float)(H - 1) / (float)(ch - 1) : 0.0f;
  const float ws = cw > 1 ? (bx.w - bx.y) * (float)(W - 1) / (float)(cw - 1) : 0.0f;
  const float in_y = ch > 1 ? bx.x * (float)(H - 1) + (float)i * hs : 0.5f * (bx.x + bx.z) * (float)(H - 1);
  const float in_x = cw > 1 ? bx.y * (float)(W - 1) + (float)j * ws : 0.5f * (bx.y + bx.w) * (float)(W - 1);
  g.valid = !(in_y < 0.0f || in_y > (float)(H - 1) || in_x < 0.0f || in_x > (float)(W - 1));
  const float fy = floorf(in_y), fx = floorf(in_x);
  g.top = (int)fy; g.bot = (int)ceilf(in_y);
  g.left = (int)fx; g.right = (int)ceilf(in_x);
  g.yl = in_y - fy; g.xl = in_x - fx;
  return g;
}

// one warp per output pixel (r, i, j); lanes stride over 8-channel vectors
__global__ void __launch_bounds__(256)
crop_resize_fwd_kernel(const bf16* __restrict__ feat, int H, int W, int C, const float4* __restrict__ boxes,
                       const int* __restrict__ box_ind, int R, int ch, int cw, bf16* __restrict__ out) {
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)R * ch * cw;
  if (gw >= total) return;
  const int j = (int)(gw % cw);
  const int i = (int)((gw / cw) % ch);
  const int r = (int)(gw / ((long long)cw * ch));
  const float4 bx = boxes[r];
  const int bi = box_ind ? box_ind[r] : 0;
  const CropGeom g = crop_geom(bx, i, j, ch, cw, H, W);
  uint4* o = reinterpret_cast<uint4*>(out + gw * C);
  const int nvec = C >> 3;
  if (!g.valid) {
    for (int v = lane; v < nvec; v += 32) o[v] = make_uint4(0, 0, 0, 0);   // extrapolation_value = 0
    return;
  }
  const bf16* base = feat + (long long)bi * H * W * C;
  const uint4* tl = reinterpret_cast<const uint4*>(base + ((long long)g.top * W + g.left) * C);
  const uint4* tr = reinterpret_cast<const uint4*>(base + ((long long)g.top * W + g.right) * C);
  const uint4* bl = reinterpret_cast<const uint4*>(base + ((long long)g.bot * W + g.left) * C);
  const uint4* br = reinterpret_cast<const uint4*>(base + ((long long)g.bot * W + g.right) * C);
  for (int v = lane; v < nvec; v += 32) {
    float a[8], b[8], c[8], d[8], y[8];
    unpack8(__ldg(tl + v), a); unpack8(__ldg(tr + v), b);
    unpack8(__ldg(bl + v), c); unpack8(__ldg(br + v), d);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float top = a[q] + (b[q] - a[q]) * g.xl;
      const float bot = c[q] + (d[q] - c[q]) * g.xl;
      y[q] = top + (bot - top) * g.yl;
    }
    o[v] = pack8(y);
  }
}

// CropAndResizeGradImage: scatter-add of dcrops into an fp32 feature gradient (vector atomics)
__global__ void __launch_bounds__(256)
crop_resize_bwd_kernel(const bf16* __restrict__ dcrop, int H, int W, int C, const float4* __restrict__ boxes,
                       const int* __restrict__ box_ind, int R, int ch, int cw, float* __restrict__ dfeat) {
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)R * ch * cw;
  if (gw >= total) return;
  const int j = (int)(gw % cw);
  const int i = (int)((gw / cw) % ch);
  const int r = (int)(gw / ((long long)cw * ch));
  const float4 bx = boxes[r];
  const int bi = box_ind ? box_ind[r] : 0;
  const CropGeom g = crop_geom(bx, i, j, ch, cw, H, W);
  if (!g.valid) return;
  const uint4* src = reinterpret_cast<const uint4*>(dcrop + gw * C);
  float* base = dfeat + (long long)bi * H * W * C;
  float* tl = base + ((long long)g.top * W + g.left) * C;
  float* tr = base + ((long long)g.top * W + g.right) * C;
  float* bl = base + ((long long)g.bot * W + g.left) * C;
  float* br = base + ((long long)g.bot * W + g.right) * C;
  const float wtl = (1.0f - g.yl) * (1.0f - g.xl), wtr = (1.0f - g.yl) * g.xl;
  const float wbl = g.yl * (1.0f - g.xl), wbr = g.yl * g.xl;
  const int nvec = C >> 3;
  for (int v = lane; v < nvec; v += 32) {
    float d[8];
    unpack8(__ldg(src + v), d);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 x = make_float4(d[4 * h], d[4 * h + 1], d[4 * h + 2], d[4 * h + 3]);
      const int off = v * 8 + h * 4;
      atomicAdd(reinterpret_cast<float4*>(tl + off), make_float4(x.x * wtl, x.y * wtl, x.z * wtl, x.w * wtl));
      if (wtr != 0.0f)
        atomicAdd(reinterpret_cast<float4*>(tr + off), make_float4(x.x * wtr, x.y * wtr, x.z * wtr, x.w * wtr));
      if (wbl != 0.0f)
        atomicAdd(reinterpret_cast<float4*>(bl + off), make_float4(x.x * wbl, x.y * wbl, x.z * wbl, x.w * wbl));
      if (wbr != 0.0f)
        atomicAdd(reinterpret_cast<float4*>(br + off), make_float4(x.x * wbr, x.y * wbr, x.z * wbr, x.w * wbr));
    }
  }
}

// ------------------------------------------------------------------ max pool (NHWC bf16)
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const bf16* __restrict__ x, int N, int H, int W, int C, int k, int stride, int pad_h,
                   int pad_w, int P, int Q, bf16* __restrict__ y, long long ldy) {
  const int nvec = C >> 3;
  const long long total = (long long)N * P * Q * nvec;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(t % nvec);
    const long long pix = t / nvec;
    const int q = (int)(pix % Q);
    const int p = (int)((pix / Q) % P);
    const int n = (int)(pix / ((long long)P * Q));
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
    for (int r = 0; r < k; ++r) {
      const int h = p * stride - pad_h + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int w = q * stride - pad_w + s;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * C) + v), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], f[e]);
      }
    }
    reinterpret_cast<uint4*>(y + pix * ldy)[v] = pack8(m);
  }
}

// gather-form MaxPoolGrad: dx[h,w] = sum over windows containing (h,w) whose FIRST maximum
// (scan order r, s) is (h,w) of dy[window].  Deterministic, no atomics.
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, long long ldy, int N, int H, int W, int C,
                   int k, int stride, int pad_h, int pad_w, int P, int Q, bf16* __restrict__ dx) {
  const int nvec = C >> 3;
  const long long total = (long long)N * H * W * nvec;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(t % nvec);
    const long long pix = t / nvec;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)H * W));
    float mine[8], acc[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + pix * C) + v), mine);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
    // windows p with p*stride - pad_h <= h < p*stride - pad_h + k
    const int p_lo = max(0, (h + pad_h - k + stride) / stride);   // ceil((h+pad-k+1)/stride)
    const int p_hi = min(P - 1, (h + pad_h) / stride);
    const int q_lo = max(0, (w + pad_w - k + stride) / stride);
    const int q_hi = min(Q - 1, (w + pad_w) / stride);
    for (int p = p_lo; p <= p_hi; ++p) {
      for (int q = q_lo; q <= q_hi; ++q) {
        // is (h,w) the first maximum of window (p,q)?
        bool first[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) first[e] = true;
        for (int r = 0; r < k; ++r) {
          const int hh = p * stride - pad_h + r;
          if (hh < 0 || hh >= H) continue;
          for (int s = 0; s < k; ++s) {
            const int ww = q * stride - pad_w + s;
            if (ww < 0 || ww >= W) continue;
            if (hh == h && ww == w) continue;
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + hh) * W + ww) * C) + v), f);
            const bool before = (hh < h) || (hh == h && ww < w);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (before ? (f[e] >= mine[e]) : (f[e] > mine[e])) first[e] = false;
            }
          }
        }
        float g[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (((long long)n * P + p) * Q + q) * ldy) + v), g);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (first[e]) acc[e] += g[e];
      }
    }
    reinterpret_cast<uint4*>(dx + pix * C)[v] = pack8(acc);
  }
}

// ------------------------------------------------------------------ spatial average (ROI heads)
__global__ void __launch_bounds__(256)
avgpool_fwd_kernel(const bf16* __restrict__ x, int R, int HW, int C, bf16* __restrict__ y) {
  const int nvec = C >> 3;
  const long long total = (long long)R * nvec;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int v = (int)(t % nvec);
  const long long r = t / nvec;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
  const uint4* src = reinterpret_cast<const uint4*>(x + r * HW * C) + v;
  for (int p = 0; p < HW; ++p) {
    float f[8];
    unpack8(__ldg(src + (long long)p * nvec), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e];
  }
  const float inv = 1.0f / (float)HW;
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] *= inv;
  reinterpret_cast<uint4*>(y + r * C)[v] = pack8(acc);
}

// dx[r,p,c] = (mask[r,p,c] > 0 ? 1 : 0) * dy[r,c] / HW   (mask = the ReLU output being pooled)
template <typename TDY>
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const TDY* __restrict__ dy, long long ldy, const bf16* __restrict__ mask, float mask_hi, int R,
                   int HW, int C, bf16* __restrict__ dx) {
  const int nvec = C >> 3;
  const long long total = (long long)R * HW * nvec;
  const float inv = 1.0f / (float)HW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(t % nvec);
    const long long rp = t / nvec;
    const long long r = rp / HW;
    float g[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = to_f32<TDY>(dy[r * ldy + v * 8 + e]) * inv;
    if (mask) {
      float m[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(mask + rp * C) + v), m);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (!(m[e] > 0.0f) || (mask_hi > 0.0f && !(m[e] < mask_hi))) g[e] = 0.0f;
    }
    reinterpret_cast<uint4*>(dx + rp * C)[v] = pack8(g);
  }
}

// ------------------------------------------------------------------ stem im2col (+ preprocessing)
// rows[b,p,q, (r*S+s)*C + c] = (img[b, p*stride-pad+r, q*stride-pad+s, c] - mean[c]) * scale, zero
// outside the image and in the pad columns [R*S*C, ld).  One warp per output pixel.
__global__ void __launch_bounds__(256)
im2col_f32_kernel(const float* __restrict__ img, int B, int H, int W, int C, int R, int S, int stride, int pad_h,
                  int pad_w, int P, int Q, float m0, float m1, float m2, float m3, float scale,
                  bf16* __restrict__ out, int ld) {
  // one thread per (output pixel, 8-element vector): eight gathered taps -> one 16-byte store
  const int nvec = ld >> 3;
  const long long total = (long long)B * P * Q * nvec;
  const int kk = R * S * C;
  const float mean[4] = {m0, m1, m2, m3};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(t % nvec);
    const long long pix = t / nvec;
    const int q = (int)(pix % Q);
    const int p = (int)((pix / Q) % P);
    const int b = (int)(pix / ((long long)P * Q));
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = v * 8 + j;
      float val = 0.0f;
      if (e < kk) {
        const int c = e % C;
        const int rs = e / C;
        const int s = rs % S, r = rs / S;
        const int h = p * stride - pad_h + r, w = q * stride - pad_w + s;
        if (h >= 0 && h < H && w >= 0 && w < W)
          val = (__ldg(img + (((long long)b * H + h) * W + w) * C + c) - mean[c & 3]) * scale;
      }
      f[j] = val;
    }
    reinterpret_cast<uint4*>(out + pix * ld)[v] = pack8(f);
  }
}

// The same rows for up to 256 columns (every stem in the zoo: 7x7x3 -> 160, 3x3x3 -> 32): the
// (r, s, c) decomposition of a column is tabulated once per CTA, one warp walks output pixels and its
// lanes own the 8-column vectors, so the inner loop has no integer division at all.
__global__ void __launch_bounds__(256)
im2col_f32_tab_kernel(const float* __restrict__ img, int B, int H, int W, int C, int R, int S, int stride, int pad_h,
                      int pad_w, int P, int Q, float m0, float m1, float m2, float m3, float scale,
                      bf16* __restrict__ out, int ld) {
  __shared__ int t_off[256];
  __shared__ short t_rs[256];
  __shared__ float t_mean[256];
  const int kk = R * S * C;
  const float mean[4] = {m0, m1, m2, m3};
  for (int e = threadIdx.x; e < ld; e += blockDim.x) {
    if (e < kk) {
      const int c = e % C, rs = e / C;
      const int s_ = rs % S, r = rs / S;
      t_off[e] = (r * W + s_) * C + c;
      t_rs[e] = (short)((r << 8) | s_);
      t_mean[e] = mean[c & 3];
    } else {
      t_off[e] = 0; t_rs[e] = -1; t_mean[e] = 0.0f;
    }
  }
  __syncthreads();
  const int nvec = ld >> 3, lane = threadIdx.x & 31;
  const int npix = B * P * Q, nwarps = gridDim.x * (blockDim.x >> 5);
  for (int pix = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < npix; pix += nwarps) {
    const int q = pix % Q, bp = pix / Q;
    const int p = bp % P, b = bp / P;
    const int h0 = p * stride - pad_h, w0 = q * stride - pad_w;
    const long long idx0 = (((long long)b * H + h0) * W + w0) * C;
    for (int v = lane; v < nvec; v += 32) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = v * 8 + j;
        const int rs = t_rs[e];
        const int h = h0 + (rs >> 8), w = w0 + (rs & 255);
        float val = 0.0f;
        if (rs >= 0 && h >= 0 && h < H && w >= 0 && w < W) val = (__ldg(img + idx0 + t_off[e]) - t_mean[e]) * scale;
        f[j] = val;
      }
      reinterpret_cast<uint4*>(out + (long long)pix * ld)[v] = pack8(f);
    }
  }
}

// (x - mean[c]) * scale -> bf16 NHWC, for inspection / the oracle's view of the preprocessed image
__global__ void preprocess_kernel(const float* __restrict__ img, long long total, int C, float m0, float m1,
                                  float m2, float m3, float scale, float* __restrict__ out) {
  const float mean[4] = {m0, m1, m2, m3};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x)
    out[t] = (img[t] - mean[(int)(t % C) & 3]) * scale;
}

// ------------------------------------------------------------------ position-sensitive ROI pooling
// utils/ops.py:462-609 with global_pool=True: bin (by,bx) of the box crops channel group
// (by*nbx + bx) to a (ch/nby x cw/nbx) grid with crop_and_resize arithmetic; the output is the
// mean over bins of the mean over the grid.  feat [B,H,W,nb*D] -> out [R, D] fp32.
__global__ void __launch_bounds__(128)
psroi_fwd_kernel(const bf16* __restrict__ feat, int H, int W, int Ct, int c0, int D, int nby, int nbx, int gh, int gw_,
                 const float4* __restrict__ boxes, const int* __restrict__ box_ind, int R, float* __restrict__ out,
                 long long ldo, int ocol0) {
  const int r = blockIdx.x;
  const float4 bx = boxes[r];
  const int bi = box_ind ? box_ind[r] : 0;
  const bf16* base = feat + (long long)bi * H * W * Ct + c0;
  const float sy = (bx.z - bx.x) / (float)nby, sx = (bx.w - bx.y) / (float)nbx;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float total = 0.0f;
    for (int by = 0; by < nby; ++by)
      for (int bxi = 0; bxi < nbx; ++bxi) {
        const float4 bb = make_float4(bx.x + (float)by * sy, bx.y + (float)bxi * sx,
                                      bx.x + (float)(by + 1) * sy, bx.y + (float)(bxi + 1) * sx);
        const int c = (by * nbx + bxi) * D + d;
        float acc = 0.0f;
        for (int i = 0; i < gh; ++i)
          for (int j = 0; j < gw_; ++j) {
            const CropGeom g = crop_geom(bb, i, j, gh, gw_, H, W);
            if (!g.valid) continue;
            const float a = __bfloat162float(base[((long long)g.top * W + g.left) * Ct + c]);
            const float b = __bfloat162float(base[((long long)g.top * W + g.right) * Ct + c]);
            const float cc = __bfloat162float(base[((long long)g.bot * W + g.left) * Ct + c]);
            const float dd = __bfloat162float(base[((long long)g.bot * W + g.right) * Ct + c]);
            const float top = a + (b - a) * g.xl, bot = cc + (dd - cc) * g.xl;
            acc += top + (bot - top) * g.yl;
          }
        total += acc / (float)(gh * gw_);
      }
    out[(long long)r * ldo + ocol0 + d] = total / (float)(nby * nbx);
  }
}

__global__ void __launch_bounds__(128)
psroi_bwd_kernel(const float* __restrict__ dout, long long ldo, int ocol0, int H, int W, int Ct, int c0, int D, int nby,
                 int nbx, int gh, int gw_, const float4* __restrict__ boxes, const int* __restrict__ box_ind, int R,
                 float* __restrict__ dfeat) {
  const int r = blockIdx.x;
  const float4 bx = boxes[r];
  const int bi = box_ind ? box_ind[r] : 0;
  float* base = dfeat + (long long)bi * H * W * Ct + c0;
  const float sy = (bx.z - bx.x) / (float)nby, sx = (bx.w - bx.y) / (float)nbx;
  const float sc = 1.0f / (float)(nby * nbx) / (float)(gh * gw_);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float g0 = dout[(long long)r * ldo + ocol0 + d] * sc;
    if (g0 == 0.0f) continue;
    for (int by = 0; by < nby; ++by)
      for (int bxi = 0; bxi < nbx; ++bxi) {
        const float4 bb = make_float4(bx.x + (float)by * sy, bx.y + (float)bxi * sx,
                                      bx.x + (float)(by + 1) * sy, bx.y + (float)(bxi + 1) * sx);
        const int c = (by * nbx + bxi) * D + d;
        for (int i = 0; i < gh; ++i)
          for (int j = 0; j < gw_; ++j) {
            const CropGeom g = crop_geom(bb, i, j, gh, gw_, H, W);
            if (!g.valid) continue;
            atomicAdd(base + ((long long)g.top * W + g.left) * Ct + c, g0 * (1.0f - g.yl) * (1.0f - g.xl));
            atomicAdd(base + ((long long)g.top * W + g.right) * Ct + c, g0 * (1.0f - g.yl) * g.xl);
            atomicAdd(base + ((long long)g.bot * W + g.left) * Ct + c, g0 * g.yl * (1.0f - g.xl));
            atomicAdd(base + ((long long)g.bot * W + g.right) * Ct + c, g0 * g.yl * g.xl);
          }
      }
  }
}

// ------------------------------------------------------------------ image resize (preprocessing)
// tf.image.resize_images(BILINEAR, align_corners=False) of TF 1.7: src = dst * in/out,
// lower = floor(src), upper = min(lower + 1, in - 1)  (core/kernels/resize_bilinear_op.cc).
__global__ void resize_bilinear_f32_kernel(const float* __restrict__ x, int B, int H, int W, int C, int oh, int ow,
                                           float* __restrict__ y) {
  const long long total = (long long)B * oh * ow * C;
  const float sy = (float)H / (float)oh, sx = (float)W / (float)ow;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const long long pix = t / C;
    const int ox = (int)(pix % ow);
    const int oy = (int)((pix / ow) % oh);
    const int b = (int)(pix / ((long long)oh * ow));
    const float in_y = (float)oy * sy, in_x = (float)ox * sx;
    const int y0 = (int)floorf(in_y), x0 = (int)floorf(in_x);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float yl = in_y - (float)y0, xl = in_x - (float)x0;
    const float* base = x + (long long)b * H * W * C + c;
    const float tl = base[((long long)y0 * W + x0) * C], tr = base[((long long)y0 * W + x1) * C];
    const float bl = base[((long long)y1 * W + x0) * C], br = base[((long long)y1 * W + x1) * C];
    const float top = tl + (tr - tl) * xl, bot = bl + (br - bl) * xl;
    y[t] = top + (bot - top) * yl;
  }
}

}  // namespace

// ===================================================================================== C ABI
#define CHECK_VEC8(C, name) MTL_CHECK_ARG((C) > 0 && (C) % 8 == 0, name ": channels must be a multiple of 8 (C=%d)", (C))

extern "C" int mtl_crop_and_resize_fwd(const void* feat, int B, int H, int W, int C, const float* boxes,
                                       const int* box_ind, int R, int crop_h, int crop_w, void* out,
                                       cudaStream_t stream) {
  MTL_CHECK_ARG(feat && boxes && out, "mtl_crop_and_resize_fwd: null tensor");
  CHECK_VEC8(C, "mtl_crop_and_resize_fwd");
  if (R == 0) return MTL_OK;
  const long long warps = (long long)R * crop_h * crop_w;
  crop_resize_fwd_kernel<<<(unsigned)ceil_div_ll(warps, 8), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(feat), H, W, C, reinterpret_cast<const float4*>(boxes), box_ind, R, crop_h,
      crop_w, reinterpret_cast<bf16*>(out));
  MTL_CUDA_LAUNCH_CHECK("crop_resize_fwd_kernel");
  (void)B;
  return MTL_OK;
}

extern "C" int mtl_crop_and_resize_bwd(const void* dcrop, int B, int H, int W, int C, const float* boxes,
                                       const int* box_ind, int R, int crop_h, int crop_w, float* dfeat,
                                       cudaStream_t stream) {
  MTL_CHECK_ARG(dcrop && boxes && dfeat, "mtl_crop_and_resize_bwd: null tensor");
  CHECK_VEC8(C, "mtl_crop_and_resize_bwd");
  if (R == 0) return MTL_OK;
  const long long warps = (long long)R * crop_h * crop_w;
  crop_resize_bwd_kernel<<<(unsigned)ceil_div_ll(warps, 8), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(dcrop), H, W, C, reinterpret_cast<const float4*>(boxes), box_ind, R, crop_h,
      crop_w, dfeat);
  MTL_CUDA_LAUNCH_CHECK("crop_resize_bwd_kernel");
  (void)B;
  return MTL_OK;
}

extern "C" int mtl_maxpool_fwd(const void* x, int N, int H, int W, int C, int k, int stride, int pad_h, int pad_w,
                               int P, int Q, void* y, long long ldy, cudaStream_t stream) {
  if (ldy == 0) ldy = C;
  MTL_CHECK_ARG(x && y, "mtl_maxpool_fwd: null tensor");
  CHECK_VEC8(C, "mtl_maxpool_fwd");
  const long long total = (long long)N * P * Q * (C / 8);
  const int grid = (int)min(ceil_div_ll(total, 256), (long long)mtl_num_sms() * 16);
  maxpool_fwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), N, H, W, C, k, stride, pad_h,
                                               pad_w, P, Q, reinterpret_cast<bf16*>(y), ldy);
  MTL_CUDA_LAUNCH_CHECK("maxpool_fwd_kernel");
  return MTL_OK;
}

extern "C" int mtl_maxpool_bwd(const void* x, const void* dy, long long ldy, int N, int H, int W, int C, int k,
                               int stride, int pad_h, int pad_w, int P, int Q, void* dx, cudaStream_t stream) {
  if (ldy == 0) ldy = C;
  MTL_CHECK_ARG(x && dy && dx, "mtl_maxpool_bwd: null tensor");
  CHECK_VEC8(C, "mtl_maxpool_bwd");
  const long long total = (long long)N * H * W * (C / 8);
  const int grid = (int)min(ceil_div_ll(total, 256), (long long)mtl_num_sms() * 16);
  maxpool_bwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dy),
                                               ldy, N, H, W, C, k, stride, pad_h, pad_w, P, Q,
                                               reinterpret_cast<bf16*>(dx));
  MTL_CUDA_LAUNCH_CHECK("maxpool_bwd_kernel");
  return MTL_OK;
}

extern "C" int mtl_avgpool_fwd(const void* x, int R, int HW, int C, void* y, cudaStream_t stream) {
  MTL_CHECK_ARG(x && y, "mtl_avgpool_fwd: null tensor");
  CHECK_VEC8(C, "mtl_avgpool_fwd");
  if (R == 0) return MTL_OK;
  const long long total = (long long)R * (C / 8);
  avgpool_fwd_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), R, HW,
                                                                           C, reinterpret_cast<bf16*>(y));
  MTL_CUDA_LAUNCH_CHECK("avgpool_fwd_kernel");
  return MTL_OK;
}

extern "C" int mtl_avgpool_bwd(const void* dy, int dy_fp32, long long ldy, const void* relu_mask, float mask_hi,
                               int R, int HW, int C, void* dx, cudaStream_t stream) {
  MTL_CHECK_ARG(dy && dx, "mtl_avgpool_bwd: null tensor");
  CHECK_VEC8(C, "mtl_avgpool_bwd");
  if (R == 0) return MTL_OK;
  const long long total = (long long)R * HW * (C / 8);
  const int grid = (int)min(ceil_div_ll(total, 256), (long long)mtl_num_sms() * 16);
  if (dy_fp32)
    avgpool_bwd_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(dy), ldy,
                                                        reinterpret_cast<const bf16*>(relu_mask), mask_hi, R, HW, C,
                                                        reinterpret_cast<bf16*>(dx));
  else
    avgpool_bwd_kernel<bf16><<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(dy), ldy,
                                                       reinterpret_cast<const bf16*>(relu_mask), mask_hi, R, HW, C,
                                                       reinterpret_cast<bf16*>(dx));
  MTL_CUDA_LAUNCH_CHECK("avgpool_bwd_kernel");
  return MTL_OK;
}

extern "C" int mtl_im2col_f32(const float* img, int B, int H, int W, int C, int R, int S, int stride, int pad_h,
                              int pad_w, int P, int Q, const float* mean, float scale, void* out, int ld,
                              cudaStream_t stream) {
  MTL_CHECK_ARG(img && out, "mtl_im2col_f32: null tensor");
  MTL_CHECK_ARG(C >= 1 && C <= 4, "mtl_im2col_f32: 1..4 input channels (C=%d)", C);
  MTL_CHECK_ARG(ld >= R * S * C && ld % 8 == 0, "mtl_im2col_f32: ld must be >= R*S*C and a multiple of 8");
  float m[4] = {0, 0, 0, 0};
  if (mean) for (int c = 0; c < C; ++c) m[c] = mean[c];
  const long long total = (long long)B * P * Q * (ld / 8);
  if (ld <= 256 && R < 128 && S < 256 && (long long)B * P * Q < (1ll << 31)) {
    const int grid_t = (int)min(ceil_div_ll((long long)B * P * Q, 8), (long long)mtl_num_sms() * 16);
    im2col_f32_tab_kernel<<<grid_t, 256, 0, stream>>>(img, B, H, W, C, R, S, stride, pad_h, pad_w, P, Q, m[0], m[1],
                                                      m[2], m[3], scale, reinterpret_cast<bf16*>(out), ld);
    MTL_CUDA_LAUNCH_CHECK("im2col_f32_tab_kernel");
    return MTL_OK;
  }
  const int grid = (int)min(ceil_div_ll(total, 256), (long long)mtl_num_sms() * 32);
  im2col_f32_kernel<<<grid, 256, 0, stream>>>(img, B, H, W, C, R, S, stride, pad_h, pad_w, P, Q, m[0], m[1], m[2],
                                              m[3], scale, reinterpret_cast<bf16*>(out), ld);
  MTL_CUDA_LAUNCH_CHECK("im2col_f32_kernel");
  return MTL_OK;
}

extern "C" int mtl_preprocess(const float* img, long long total, int C, const float* mean, float scale, float* out,
                              cudaStream_t stream) {
  MTL_CHECK_ARG(img && out && C >= 1 && C <= 4, "mtl_preprocess: bad args");
  float m[4] = {0, 0, 0, 0};
  if (mean) for (int c = 0; c < C; ++c) m[c] = mean[c];
  const int grid = (int)min(ceil_div_ll(total, 256), (long long)mtl_num_sms() * 16);
  preprocess_kernel<<<grid, 256, 0, stream>>>(img, total, C, m[0], m[1], m[2], m[3], scale, out);
  MTL_CUDA_LAUNCH_CHECK("preprocess_kernel");
  return MTL_OK;
}

extern "C" int mtl_psroi_fwd(const void* feat, int B, int H, int W, int Ct, int c0, int D, int nby, int nbx,
                             int crop_h, int crop_w, const float* boxes, const int* box_ind, int R, float* out,
                             long long ldo, int ocol0, cudaStream_t stream) {
  MTL_CHECK_ARG(feat && boxes && out, "mtl_psroi_fwd: null tensor");
  MTL_CHECK_ARG(crop_h % nby == 0 && crop_w % nbx == 0, "mtl_psroi_fwd: crop size must be divisible by bins");
  MTL_CHECK_ARG(c0 + nby * nbx * D <= Ct, "mtl_psroi_fwd: position-sensitive channels exceed the map depth");
  if (R == 0) return MTL_OK;
  psroi_fwd_kernel<<<R, 128, 0, stream>>>(reinterpret_cast<const bf16*>(feat), H, W, Ct, c0, D, nby, nbx,
                                          crop_h / nby, crop_w / nbx, reinterpret_cast<const float4*>(boxes), box_ind,
                                          R, out, ldo, ocol0);
  MTL_CUDA_LAUNCH_CHECK("psroi_fwd_kernel");
  (void)B;
  return MTL_OK;
}

extern "C" int mtl_psroi_bwd(const float* dout, long long ldo, int ocol0, int B, int H, int W, int Ct, int c0, int D,
                             int nby, int nbx, int crop_h, int crop_w, const float* boxes, const int* box_ind, int R,
                             float* dfeat, cudaStream_t stream) {
  MTL_CHECK_ARG(dout && boxes && dfeat, "mtl_psroi_bwd: null tensor");
  MTL_CHECK_ARG(crop_h % nby == 0 && crop_w % nbx == 0, "mtl_psroi_bwd: crop size must be divisible by bins");
  if (R == 0) return MTL_OK;
  psroi_bwd_kernel<<<R, 128, 0, stream>>>(dout, ldo, ocol0, H, W, Ct, c0, D, nby, nbx, crop_h / nby, crop_w / nbx,
                                          reinterpret_cast<const float4*>(boxes), box_ind, R, dfeat);
  MTL_CUDA_LAUNCH_CHECK("psroi_bwd_kernel");
  (void)B;
  return MTL_OK;
}

// dst (fp32) += src (bf16): merges dense bf16 feature gradients into the fp32 accumulator
__global__ void add_bf16_to_f32_kernel(const bf16* __restrict__ src, long long n, float* __restrict__ dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] += __bfloat162float(src[i]);
}
extern "C" int mtl_add_bf16_to_f32(const void* src, long long n, float* dst, cudaStream_t stream) {
  MTL_CHECK_ARG(src && dst, "mtl_add_bf16_to_f32: null tensor");
  if (n == 0) return MTL_OK;
  const int grid = (int)min(ceil_div_ll(n, 256), (long long)mtl_num_sms() * 16);
  add_bf16_to_f32_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(src), n, dst);
  MTL_CUDA_LAUNCH_CHECK("add_bf16_to_f32_kernel");
  return MTL_OK;
}

extern "C" int mtl_resize_bilinear_f32(const float* x, int B, int H, int W, int C, int out_h, int out_w, float* y,
                                       cudaStream_t stream) {
  MTL_CHECK_ARG(x && y && out_h > 0 && out_w > 0, "mtl_resize_bilinear_f32: bad args");
  const long long total = (long long)B * out_h * out_w * C;
  const int grid = (int)min(ceil_div_ll(total, 256), (long long)mtl_num_sms() * 16);
  resize_bilinear_f32_kernel<<<grid, 256, 0, stream>>>(x, B, H, W, C, out_h, out_w, y);
  MTL_CUDA_LAUNCH_CHECK("resize_bilinear_f32_kernel");
  return MTL_OK;
}

// ===================================================================================== depthwise 3x3
// slim.separable_conv2d(depth_multiplier=1) depthwise stage (slim/nets/mobilenet_v1.py:230-238,
// object_detection/models/faster_rcnn_mobilenet_v1_feature_extractor.py:169-182): NHWC bf16, weights
// [C, 3, 3] bf16 (frozen batch-norm scale folded in), per-channel bias, optional ReLU6.  HBM-bound:
// one thread per (pixel, 8-channel vector), 16-byte accesses, taps served from L1/L2.
namespace {

__device__ __forceinline__ void unpack8b(const uint4 v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    f[2 * q] = __uint_as_float(w[q] << 16);
    f[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8b(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
    w[q] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256)
dwconv3x3_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const float* __restrict__ bias, int N,
                     int H, int W, int C, int stride, int pad_h, int pad_w, int P, int Q, int act,
                     bf16* __restrict__ y) {
  const int nvec = C >> 3;
  const long long total = (long long)N * P * Q * nvec;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(t % nvec);
    const long long pix = t / nvec;
    const int q = (int)(pix % Q);
    const int p = (int)((pix / Q) % P);
    const int n = (int)(pix / ((long long)P * Q));
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bias ? bias[v * 8 + e] : 0.0f;
    for (int r = 0; r < 3; ++r) {
      const int h = p * stride - pad_h + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int ww = q * stride - pad_w + s;
        if (ww < 0 || ww >= W) continue;
        float f[8];
        unpack8b(__ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + ww) * C) + v), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e] * __bfloat162float(w[(v * 8 + e) * 9 + r * 3 + s]);
      }
    }
    if (act) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaxf(acc[e], 0.0f);
      if (act == 2) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fminf(acc[e], 6.0f);
      }
    }
    reinterpret_cast<uint4*>(y + pix * C)[v] = pack8b(acc);
  }
}

// dx[n,h,w,c] = sum_{r,s} dy[n,(h+pad-r)/stride,(w+pad-s)/stride,c] * w[c,r,s], masked by the activation
// that produced x (mask > 0, and mask < mask_hi when mask_hi > 0)
__global__ void __launch_bounds__(256)
dwconv3x3_dgrad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ w, int N, int H, int W, int C,
                       int stride, int pad_h, int pad_w, int P, int Q, const bf16* __restrict__ mask,
                       float mask_hi, bf16* __restrict__ dx) {
  const int nvec = C >> 3;
  const long long total = (long long)N * H * W * nvec;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(t % nvec);
    const long long pix = t / nvec;
    const int ww = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)H * W));
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
    for (int r = 0; r < 3; ++r) {
      const int hy = h + pad_h - r;
      if (hy < 0 || hy % stride) continue;
      const int p = hy / stride;
      if (p >= P) continue;
      for (int s = 0; s < 3; ++s) {
        const int wy = ww + pad_w - s;
        if (wy < 0 || wy % stride) continue;
        const int q = wy / stride;
        if (q >= Q) continue;
        float g[8];
        unpack8b(__ldg(reinterpret_cast<const uint4*>(dy + (((long long)n * P + p) * Q + q) * C) + v), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += g[e] * __bfloat162float(w[(v * 8 + e) * 9 + r * 3 + s]);
      }
    }
    if (mask) {
      float m[8];
      unpack8b(__ldg(reinterpret_cast<const uint4*>(mask + pix * C) + v), m);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (!(m[e] > 0.0f) || (mask_hi > 0.0f && !(m[e] < mask_hi))) acc[e] = 0.0f;
    }
    reinterpret_cast<uint4*>(dx + pix * C)[v] = pack8b(acc);
  }
}

// dw[c,r,s] += scale[c] * sum_pixels dy[p,c] * x[p*stride - pad + (r,s), c]; block = 32 channels x 8 pixel lanes
__global__ void __launch_bounds__(256)
dwconv3x3_wgrad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, int N, int H, int W, int C,
                       int stride, int pad_h, int pad_w, int P, int Q, const float* __restrict__ scale,
                       float* __restrict__ dw) {
  __shared__ float part[8][32][9];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const long long npix = (long long)N * P * Q;
  const long long per = (npix + gridDim.y - 1) / gridDim.y;
  const long long p0 = blockIdx.y * per, p1 = min(p0 + per, npix);
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.0f;
  if (c < C) {
    for (long long pix = p0 + pl; pix < p1; pix += 8) {
      const int q = (int)(pix % Q);
      const int p = (int)((pix / Q) % P);
      const int n = (int)(pix / ((long long)P * Q));
      const float g = __bfloat162float(dy[pix * C + c]);
      if (g == 0.0f) continue;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int h = p * stride - pad_h + r;
        if (h < 0 || h >= H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ww = q * stride - pad_w + s;
          if (ww < 0 || ww >= W) continue;
          acc[r * 3 + s] += g * __bfloat162float(x[(((long long)n * H + h) * W + ww) * C + c]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) part[pl][cl][k] = acc[k];
  __syncthreads();
  if (pl == 0 && c < C) {
    const float sc = scale ? scale[c] : 1.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += part[i][cl][k];
      atomicAdd(dw + (long long)c * 9 + k, s * sc);
    }
  }
}

}  // namespace

extern "C" int mtl_dwconv3x3_fwd(const void* x, const void* w, const float* bias, int N, int H, int W, int C,
                                 int stride, int pad_h, int pad_w, int P, int Q, int activation, void* y,
                                 cudaStream_t stream) {
  MTL_CHECK_ARG(x && w && y, "mtl_dwconv3x3_fwd: null tensor");
  CHECK_VEC8(C, "mtl_dwconv3x3_fwd");
  const long long total = (long long)N * P * Q * (C / 8);
  const int grid = (int)min(ceil_div_ll(total, 256), (long long)mtl_num_sms() * 16);
  dwconv3x3_fwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(w),
                                                 bias, N, H, W, C, stride, pad_h, pad_w, P, Q, activation,
                                                 reinterpret_cast<bf16*>(y));
  MTL_CUDA_LAUNCH_CHECK("dwconv3x3_fwd_kernel");
  return MTL_OK;
}

extern "C" int mtl_dwconv3x3_dgrad(const void* dy, const void* w, int N, int H, int W, int C, int stride, int pad_h,
                                   int pad_w, int P, int Q, const void* mask, float mask_hi, void* dx,
                                   cudaStream_t stream) {
  MTL_CHECK_ARG(dy && w && dx, "mtl_dwconv3x3_dgrad: null tensor");
  CHECK_VEC8(C, "mtl_dwconv3x3_dgrad");
  const long long total = (long long)N * H * W * (C / 8);
  const int grid = (int)min(ceil_div_ll(total, 256), (long long)mtl_num_sms() * 16);
  dwconv3x3_dgrad_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(dy),
                                                   reinterpret_cast<const bf16*>(w), N, H, W, C, stride, pad_h, pad_w,
                                                   P, Q, reinterpret_cast<const bf16*>(mask), mask_hi,
                                                   reinterpret_cast<bf16*>(dx));
  MTL_CUDA_LAUNCH_CHECK("dwconv3x3_dgrad_kernel");
  return MTL_OK;
}

extern "C" int mtl_dwconv3x3_wgrad(const void* dy, const void* x, int N, int H, int W, int C, int stride, int pad_h,
                                   int pad_w, int P, int Q, const float* scale, float* dw, cudaStream_t stream) {
  MTL_CHECK_ARG(dy && x && dw, "mtl_dwconv3x3_wgrad: null tensor");
  const long long npix = (long long)N * P * Q;
  dim3 grid(ceil_div(C, 32), (unsigned)min((long long)256, ceil_div_ll(npix, 64)));
  dwconv3x3_wgrad_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(x),
                                                   N, H, W, C, stride, pad_h, pad_w, P, Q, scale, dw);
  MTL_CUDA_LAUNCH_CHECK("dwconv3x3_wgrad_kernel");
  return MTL_OK;
}

// ===================================================================================== avg pool 3x3/1 SAME
// slim.avg_pool2d(net, 3, stride=1, padding='SAME') of Mixed_5b (slim/nets/inception_resnet_v2.py:176-181):
// TF divides by the number of in-bounds taps.  bwd: dx[h,w] = sum over windows containing (h,w) of dy/count.
namespace {
__global__ void __launch_bounds__(256)
avgpool3x3_kernel(const bf16* __restrict__ x, int N, int H, int W, int C, int backward, bf16* __restrict__ y) {
  const int nvec = C >> 3;
  const long long total = (long long)N * H * W * nvec;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(t % nvec);
    const long long pix = t / nvec;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)H * W));
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
    int cnt = 0;
    for (int r = -1; r <= 1; ++r) {
      const int hh = h + r;
      if (hh < 0 || hh >= H) continue;
      for (int s = -1; s <= 1; ++s) {
        const int ww = w + s;
        if (ww < 0 || ww >= W) continue;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + hh) * W + ww) * C) + v), f);
        float sc = 1.0f;
        if (backward) {   // the neighbour's own window size normalises its gradient
          const int ch = min(hh + 1, H - 1) - max(hh - 1, 0) + 1, cw = min(ww + 1, W - 1) - max(ww - 1, 0) + 1;
          sc = 1.0f / (float)(ch * cw);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e] * sc;
        ++cnt;
      }
    }
    if (!backward) {
      const float inv = 1.0f / (float)cnt;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] *= inv;
    }
    reinterpret_cast<uint4*>(y + pix * C)[v] = pack8(acc);
  }
}
}  // namespace

extern "C" int mtl_avgpool3x3_same(const void* x, int N, int H, int W, int C, int backward, void* y,
                                   cudaStream_t stream) {
  MTL_CHECK_ARG(x && y, "mtl_avgpool3x3_same: null tensor");
  CHECK_VEC8(C, "mtl_avgpool3x3_same");
  const long long total = (long long)N * H * W * (C / 8);
  const int grid = (int)min(ceil_div_ll(total, 256), (long long)mtl_num_sms() * 16);
  avgpool3x3_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), N, H, W, C, backward,
                                              reinterpret_cast<bf16*>(y));
  MTL_CUDA_LAUNCH_CHECK("avgpool3x3_kernel");
  return MTL_OK;
}
