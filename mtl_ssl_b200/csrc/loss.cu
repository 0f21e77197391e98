// Fused loss epilogues (value + gradient in one pass) for the RPN, the box classifier and the
// three auxiliary heads, plus the small fp32 heads that are too narrow for tensor-core tiles
// (edge-mask 1x1 conv + tanh, MTL refiner FC).
//
// Reference (under /root/reference/object_detection/):
//   core/losses.py:169-196 (smooth L1), :285-352 (softmax CE, hard and soft labels)
//   meta_architectures/faster_rcnn_meta_arch.py:1591-1668 (_loss_rpn), :1670-1793
//   (_loss_box_classifier + closeness), :1795-1837 (refined), :1839-1858 (window),
//   :1860-1881 (edgemask), :764-846 (refiner), core/mask_predictor.py:90-119.
// Every kernel accumulates its scalar loss into losses[slot] with one atomicAdd per block
// (fp32; the summation order is not fixed, tolerance documented in the tests).
#include "common.cuh"

namespace {

__device__ __forceinline__ float block_sum(float v, float* red /*[32]*/) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    t = warp_sum(t);
  }
  return t;   // valid in warp 0
}

__device__ __forceinline__ float smooth_l1(float d, float sigma2, float* grad) {
  const float ad = fabsf(d);
  if (ad < 1.0f / sigma2) {
    *grad = sigma2 * d;
    return 0.5f * sigma2 * d * d;
  }
  *grad = d > 0.0f ? 1.0f : -1.0f;
  return ad - 0.5f / sigma2;
}

__device__ __forceinline__ float4 box_encode(const float4 g, const float4 a) {
  float wa = a.w - a.y, ha = a.z - a.x;
  const float yca = a.x + ha / 2.0f, xca = a.y + wa / 2.0f;
  float w = g.w - g.y, h = g.z - g.x;
  const float yc = g.x + h / 2.0f, xc = g.y + w / 2.0f;
  ha += 1e-8f; wa += 1e-8f; h += 1e-8f; w += 1e-8f;
  return make_float4((yc - yca) / ha * 10.0f, (xc - xca) / wa * 10.0f, logf(h / ha) * 5.0f, logf(w / wa) * 5.0f);
}

// ------------------------------------------------------------------ RPN loss (fmA:1591-1668)
// One thread per kept anchor; only sampled anchors contribute.  Regression targets are encoded
// on the fly from the match vector.  Gradients are scattered into the dense RPN head output
// gradient d_out[B,HW,ld] (bf16, pre-zeroed) at the anchor's box / objectness columns.
__global__ void __launch_bounds__(256)
rpn_loss_kernel(const float* __restrict__ rpn_out, long long ld, int box_col0, int cls_col0, int A, int HW,
                const int* __restrict__ keep_idx, const float4* __restrict__ anchors, int Nk,
                const float4* __restrict__ gt, int Gmax, const int* __restrict__ match,
                const unsigned char* __restrict__ sampled, const int* __restrict__ counts, int B, float loc_w,
                float obj_w, float sigma, float* __restrict__ losses, bf16* __restrict__ d_out) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  float l_loc = 0.0f, l_obj = 0.0f;
  if (j < Nk && sampled[(long long)b * Nk + j]) {
    const float norm = fmaxf((float)counts[b * 4 + 3], 1.0f) * (float)B;
    const int ai = keep_idx ? keep_idx[j] : j;
    const int loc = ai / A, a = ai - loc * A;
    const long long ro = ((long long)b * HW + loc) * ld;
    const float* row = rpn_out + ro;
    const int m = match[(long long)b * Nk + j];
    const float sigma2 = sigma * sigma;
    if (m >= 0) {
      const float4 t = box_encode(gt[(long long)b * Gmax + m], anchors[j]);
      const float tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float g;
        l_loc += smooth_l1(row[box_col0 + a * 4 + c] - tt[c], sigma2, &g);
        if (d_out) d_out[ro + box_col0 + a * 4 + c] = __float2bfloat16_rn(g * loc_w / norm);
      }
    }
    const float l0 = row[cls_col0 + a * 2], l1 = row[cls_col0 + a * 2 + 1];
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
    const float lse = mx + logf(e0 + e1);
    const int tgt = m >= 0 ? 1 : 0;
    l_obj = lse - (tgt ? l1 : l0);
    if (d_out) {
      const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
      d_out[ro + cls_col0 + a * 2] = __float2bfloat16_rn((p0 - (tgt ? 0.0f : 1.0f)) * obj_w / norm);
      d_out[ro + cls_col0 + a * 2 + 1] = __float2bfloat16_rn((p1 - (tgt ? 1.0f : 0.0f)) * obj_w / norm);
    }
    l_loc = l_loc * loc_w / norm;
    l_obj = l_obj * obj_w / norm;
  }
  const float s0 = block_sum(l_loc, red);
  const float s1 = block_sum(l_obj, red);
  if (threadIdx.x == 0) {
    if (s0 != 0.0f) atomicAdd(losses + 0, s0);
    if (s1 != 0.0f) atomicAdd(losses + 1, s1);
  }
}

// ------------------------------------------------------------------ box classifier loss (fmA:1670-1769)
// head_out[B*P, ld]: K*4 box codes at box_col0, K+1 class logits at cls_col0.  One warp per ROI.
__global__ void __launch_bounds__(256)
box_classifier_loss_kernel(const float* __restrict__ head, long long ld, int box_col0, int cls_col0, int K,
                           const int* __restrict__ cls_t, const float4* __restrict__ reg_t,
                           const float* __restrict__ reg_w, const float* __restrict__ cls_w,
                           const int* __restrict__ num_props, int B, int P, float loc_w, float clsl_w,
                           float* __restrict__ losses, float* __restrict__ d_head, long long ldd) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31;
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  float l_loc = 0.0f, l_cls = 0.0f;
  if (r < (long long)B * P) {
    const int b = (int)(r / P), p = (int)(r % P);
    const int np = num_props[b];
    const bool real = p < np;                                   // paddings_indicator
    const float norm = (float)max(np, 1) * (float)B;
    const float* row = head + r * ld;
    float* drow = d_head ? d_head + r * ldd : nullptr;
    const int c = cls_t[r];
    const int K1 = K + 1;
    // zero the whole gradient row first (box columns of other classes get no gradient)
    if (drow) {
      for (int k = lane; k < K * 4; k += 32) drow[box_col0 + k] = 0.0f;
    }
    __syncwarp();
    // localisation: the code of the target class (T12), sigma = 1
    if (lane < 4 && c > 0 && real) {
      const float4 t = reg_t[r];
      const float tt = lane == 0 ? t.x : (lane == 1 ? t.y : (lane == 2 ? t.z : t.w));
      float g;
      const float w = reg_w[r];
      l_loc = smooth_l1(row[box_col0 + (c - 1) * 4 + lane] - tt, 1.0f, &g) * w * loc_w / norm;
      if (drow) drow[box_col0 + (c - 1) * 4 + lane] = g * w * loc_w / norm;
    }
    // classification: softmax CE against one-hot(c)
    float mx = -INFINITY;
    for (int k = lane; k < K1; k += 32) mx = fmaxf(mx, row[cls_col0 + k]);
    mx = warp_max(mx);
    float se = 0.0f;
    for (int k = lane; k < K1; k += 32) se += expf(row[cls_col0 + k] - mx);
    se = warp_sum(se);
    const float w = real ? cls_w[r] * clsl_w / norm : 0.0f;
    if (lane == 0) l_cls = (mx + logf(se) - row[cls_col0 + c]) * w;
    if (drow)
      for (int k = lane; k < K1; k += 32)
        drow[cls_col0 + k] = (expf(row[cls_col0 + k] - mx) / se - (k == c ? 1.0f : 0.0f)) * w;
  }
  const float s0 = block_sum(l_loc, red);
  const float s1 = block_sum(l_cls, red);
  if (threadIdx.x == 0) {
    if (s0 != 0.0f) atomicAdd(losses + 0, s0);
    if (s1 != 0.0f) atomicAdd(losses + 1, s1);
  }
}

// ------------------------------------------------------------------ generic weighted softmax CE
// loss += scale * w[r] * ( -sum_k t[r,k] * log_softmax(z[r, col0:col0+C])_k ), soft (dense t)
// or hard (int class) targets; gradient written or ACCUMULATED into d[r, dcol0:dcol0+C].
// Padding rows (p >= num_props[b]) are skipped when num_props is given; with per_image_norm the
// weight is additionally divided by max(num_props[b],1).
__global__ void __launch_bounds__(256)
softmax_ce_kernel(const float* __restrict__ z, long long ldz, int col0, int C, const float* __restrict__ t_soft,
                  long long ldt, int tcol0, const int* __restrict__ t_hard, const float* __restrict__ w,
                  const int* __restrict__ num_props, int P, int per_image_norm, long long rows, float scale,
                  float* __restrict__ loss, float* __restrict__ d, long long ldd, int dcol0, int accumulate) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31;
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  float l = 0.0f;
  if (r < rows) {
    float wr = scale * (w ? w[r] : 1.0f);
    if (num_props) {
      const int b = (int)(r / P), p = (int)(r % P);
      const int np = num_props[b];
      if (p >= np) wr = 0.0f;
      if (per_image_norm) wr /= (float)max(np, 1);
    }
    const float* row = z + r * ldz + col0;
    float mx = -INFINITY;
    for (int k = lane; k < C; k += 32) mx = fmaxf(mx, row[k]);
    mx = warp_max(mx);
    float se = 0.0f;
    for (int k = lane; k < C; k += 32) se += expf(row[k] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    float tsum = 0.0f, dot = 0.0f;
    const int hc = t_hard ? t_hard[r] : -1;
    for (int k = lane; k < C; k += 32) {
      const float tk = t_hard ? (k == hc ? 1.0f : 0.0f) : t_soft[r * ldt + tcol0 + k];
      tsum += tk;
      dot += tk * (row[k] - lse);
    }
    tsum = warp_sum(tsum);
    dot = warp_sum(dot);
    if (lane == 0) l = -dot * wr;
    if (d) {
      float* drow = d + r * ldd + dcol0;
      for (int k = lane; k < C; k += 32) {
        const float tk = t_hard ? (k == hc ? 1.0f : 0.0f) : t_soft[r * ldt + tcol0 + k];
        const float g = (expf(row[k] - lse) * tsum - tk) * wr;
        drow[k] = accumulate ? drow[k] + g : g;
      }
    }
  }
  const float s = block_sum(l, red);
  if (threadIdx.x == 0 && s != 0.0f) atomicAdd(loss, s);
}

// ------------------------------------------------------------------ edge-mask head (mp:90-119, fmA:1860-1881)
// forward: z[b,px,k] = bias[k] + sum_c x[b,px,c] * w[k,c];  a = tanh(z).  One warp per pixel.
__global__ void __launch_bounds__(256)
edgemask_fwd_kernel(const bf16* __restrict__ x, long long npix, int C, const float* __restrict__ w,
                    const float* __restrict__ bias, float* __restrict__ act) {
  const long long px = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (px >= npix) return;
  float a0 = 0.0f, a1 = 0.0f;
  for (int c = lane * 2; c < C; c += 64) {
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(x + px * C + c);
    const float x0 = __low2float(v), x1 = __high2float(v);
    a0 += x0 * w[c] + x1 * w[c + 1];
    a1 += x0 * w[C + c] + x1 * w[C + c + 1];
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if (lane == 0) {
    act[px * 2] = tanhf(a0 + bias[0]);
    act[px * 2 + 1] = tanhf(a1 + bias[1]);
  }
}

// loss over the resized (TF ResizeBilinear, align_corners=False) predictions; gradient wrt the
// tanh activations scattered with atomics into d_act[b,px,2] (pre-zeroed).
__global__ void __launch_bounds__(256)
edgemask_loss_kernel(const float* __restrict__ act, int B, int H, int W, const float* __restrict__ gt /*[B,2,oh,ow]*/,
                     int oh, int ow, float scale, float* __restrict__ loss, float* __restrict__ d_act) {
  __shared__ float red[32];
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)B * oh * ow;
  float l = 0.0f;
  if (t < total) {
    const int ox = (int)(t % ow);
    const int oy = (int)((t / ow) % oh);
    const int b = (int)(t / ((long long)oh * ow));
    const float sy = (float)H / (float)oh, sx = (float)W / (float)ow;
    const float in_y = (float)oy * sy, in_x = (float)ox * sx;
    const int y0 = (int)floorf(in_y), x0 = (int)floorf(in_x);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float yl = in_y - (float)y0, xl = in_x - (float)x0;
    const long long i00 = ((long long)b * H + y0) * W + x0, i01 = ((long long)b * H + y0) * W + x1;
    const long long i10 = ((long long)b * H + y1) * W + x0, i11 = ((long long)b * H + y1) * W + x1;
    float p[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float top = act[i00 * 2 + k] + (act[i01 * 2 + k] - act[i00 * 2 + k]) * xl;
      const float bot = act[i10 * 2 + k] + (act[i11 * 2 + k] - act[i10 * 2 + k]) * xl;
      p[k] = top + (bot - top) * yl;
    }
    const float fg = gt[(((long long)b * 2 + 0) * oh + oy) * ow + ox];
    const float wt = gt[(((long long)b * 2 + 1) * oh + oy) * ow + ox] * scale;
    const float tg[2] = {1.0f - fg, fg};
    const float mx = fmaxf(p[0], p[1]);
    const float e0 = expf(p[0] - mx), e1 = expf(p[1] - mx);
    const float lse = mx + logf(e0 + e1);
    l = -(tg[0] * (p[0] - lse) + tg[1] * (p[1] - lse)) * wt;
    if (d_act && wt != 0.0f) {
      const float ts = tg[0] + tg[1];
      const float g[2] = {(e0 / (e0 + e1) * ts - tg[0]) * wt, (e1 / (e0 + e1) * ts - tg[1]) * wt};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        atomicAdd(d_act + i00 * 2 + k, g[k] * (1.0f - yl) * (1.0f - xl));
        atomicAdd(d_act + i01 * 2 + k, g[k] * (1.0f - yl) * xl);
        atomicAdd(d_act + i10 * 2 + k, g[k] * yl * (1.0f - xl));
        atomicAdd(d_act + i11 * 2 + k, g[k] * yl * xl);
      }
    }
  }
  const float s = block_sum(l, red);
  if (threadIdx.x == 0 && s != 0.0f) atomicAdd(loss, s);
}

// backward of the 1x1 conv + tanh: dz = d_act * (1 - act^2);  dfeat[px,c] += sum_k dz_k w[k,c];
// dw[k,c] += sum_px dz_k x[px,c];  dbias[k] += sum_px dz_k.  Block = 64 channels x 4 pixel lanes.
__global__ void __launch_bounds__(256)
edgemask_bwd_kernel(const bf16* __restrict__ x, long long npix, int C, const float* __restrict__ w,
                    const float* __restrict__ act, const float* __restrict__ d_act, float* __restrict__ dfeat,
                    float* __restrict__ dw, float* __restrict__ dbias) {
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int pl = threadIdx.x >> 6;                       // 0..3
  const long long px0 = (long long)blockIdx.y * 256;
  const long long px1 = min(px0 + 256, npix);
  if (c >= C) return;
  const float w0 = w[c], w1 = w[C + c];
  float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
  for (long long px = px0 + pl; px < px1; px += 4) {
    const float t0 = act[px * 2], t1 = act[px * 2 + 1];
    const float dz0 = d_act[px * 2] * (1.0f - t0 * t0), dz1 = d_act[px * 2 + 1] * (1.0f - t1 * t1);
    const float xv = __bfloat162float(x[px * C + c]);
    a0 += dz0 * xv;
    a1 += dz1 * xv;
    if (dfeat) dfeat[px * C + c] += dz0 * w0 + dz1 * w1;
    if (blockIdx.x == 0 && (threadIdx.x & 63) == 0) { b0 += dz0; b1 += dz1; }
  }
  atomicAdd(dw + c, a0);
  atomicAdd(dw + C + c, a1);
  if (blockIdx.x == 0 && (threadIdx.x & 63) == 0) {
    atomicAdd(dbias, b0);
    atomicAdd(dbias + 1, b1);
  }
}

// ------------------------------------------------------------------ small fp32 fully connected (refiner)
// y[m,n] = bias[n] + sum_k x[m,k] * w[n,k] (+ res[m,n])
__global__ void fc_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                              const float* __restrict__ bias, const float* __restrict__ res, long long ldr, int M,
                              int N, int K, float* __restrict__ y, long long ldy) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)M * N) return;
  const int n = (int)(t % N);
  const long long m = t / N;
  float acc = bias ? bias[n] : 0.0f;
  for (int k = 0; k < K; ++k) acc += x[m * ldx + k] * w[(long long)n * K + k];
  if (res) acc += res[m * ldr + n];
  y[m * ldy + n] = acc;
}
// dw[n,k] += sum_m dy[m,n] x[m,k];  db[n] += sum_m dy[m,n]
__global__ void fc_wgrad_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ dy,
                                long long ldy, int M, int N, int K, float* __restrict__ dw,
                                float* __restrict__ db) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)N * (K + 1)) return;
  const int k = (int)(t % (K + 1));
  const int n = (int)(t / (K + 1));
  float acc = 0.0f;
  if (k < K) {
    for (int m = 0; m < M; ++m) acc += dy[(long long)m * ldy + n] * x[(long long)m * ldx + k];
    dw[(long long)n * K + k] += acc;
  } else if (db) {
    for (int m = 0; m < M; ++m) acc += dy[(long long)m * ldy + n];
    db[n] += acc;
  }
}
// dx[m,k] = sum_n dy[m,n] w[n,k]
__global__ void fc_dgrad_kernel(const float* __restrict__ dy, long long ldy, const float* __restrict__ w, int M,
                                int N, int K, float* __restrict__ dx, long long ldx) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)M * K) return;
  const int k = (int)(t % K);
  const long long m = t / K;
  float acc = 0.0f;
  for (int n = 0; n < N; ++n) acc += dy[m * ldy + n] * w[(long long)n * K + k];
  dx[m * ldx + k] = acc;
}

// refiner input (fmA:764-833): [org(K1) | expanded-window logits, e-major inside the row (E*K1) |
// mean over ALL rows of the closeness logits (K1)].  One block per output row.
__global__ void refine_concat_kernel(const float* __restrict__ org, long long ldo, int ocol0,
                                     const float* __restrict__ win, long long ldw, int wcol0, int E,
                                     const float* __restrict__ close, long long ldc, int ccol0, int rows, int K1,
                                     float* __restrict__ out, long long ldout) {
  const int r = blockIdx.x;
  float* o = out + (long long)r * ldout;
  int col = 0;
  for (int k = threadIdx.x; k < K1; k += blockDim.x) o[col + k] = org[(long long)r * ldo + ocol0 + k];
  col += K1;
  if (win) {
    for (int t = threadIdx.x; t < E * K1; t += blockDim.x) {
      const int e = t / K1, k = t - e * K1;
      o[col + t] = win[((long long)e * rows + r) * ldw + wcol0 + k];
    }
    col += E * K1;
  }
  if (close) {
    for (int k = threadIdx.x; k < K1; k += blockDim.x) {
      float s = 0.0f;
      for (int m = 0; m < rows; ++m) s += close[(long long)m * ldc + ccol0 + k];
      o[col + k] = s / (float)rows;
    }
  }
}

// ------------------------------------------------------------------ column sums (bias gradients)
// db[n] += alpha * sum_m dy[m,n]
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ dy, long long ld, long long M, int N, float alpha, float* __restrict__ db) {
  // block handles 32 columns x a slab of rows; threads: 32 columns x 8 row lanes
  __shared__ float part[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + cl;
  const long long rows_per = (M + gridDim.y - 1) / gridDim.y;
  const long long m0 = blockIdx.y * rows_per, m1 = min(m0 + rows_per, M);
  float acc = 0.0f;
  if (n < N)
    for (long long m = m0 + rl; m < m1; m += 8) acc += to_f32<T>(dy[m * ld + n]);
  part[rl][cl] = acc;
  __syncthreads();
  if (rl == 0 && n < N) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][cl];
    atomicAdd(db + n, s * alpha);
  }
}

}  // namespace

// ===================================================================================== C ABI
extern "C" int mtl_rpn_loss(const float* rpn_out, long long ld, int box_col0, int cls_col0, int A, int HW,
                            const int* keep_idx, const float* anchors, int Nk, const float* gt, int Gmax,
                            const int* match, const unsigned char* sampled, const int* counts, int B,
                            float loc_weight, float obj_weight, float sigma, float* losses, void* d_rpn_out,
                            cudaStream_t stream) {
  MTL_CHECK_ARG(rpn_out && anchors && gt && match && sampled && counts && losses, "mtl_rpn_loss: null tensor");
  if (d_rpn_out) {
    cudaError_t e = cudaMemsetAsync(d_rpn_out, 0, sizeof(bf16) * (size_t)B * HW * ld, stream);
    if (e != cudaSuccess) { mtl_set_error("mtl_rpn_loss: memset: %s", cudaGetErrorString(e)); return MTL_ERR_CUDA; }
  }
  dim3 grid(ceil_div(Nk, 256), B);
  rpn_loss_kernel<<<grid, 256, 0, stream>>>(rpn_out, ld, box_col0, cls_col0, A, HW, keep_idx,
                                            reinterpret_cast<const float4*>(anchors), Nk,
                                            reinterpret_cast<const float4*>(gt), Gmax, match, sampled, counts, B,
                                            loc_weight, obj_weight, sigma, losses,
                                            reinterpret_cast<bf16*>(d_rpn_out));
  MTL_CUDA_LAUNCH_CHECK("rpn_loss_kernel");
  return MTL_OK;
}

extern "C" int mtl_box_classifier_loss(const float* head_out, long long ld, int box_col0, int cls_col0, int K,
                                       const int* cls_targets, const float* reg_targets, const float* reg_weights,
                                       const float* cls_weights, const int* num_proposals, int B, int P,
                                       float loc_weight, float cls_weight, float* losses, float* d_head,
                                       long long ldd, cudaStream_t stream) {
  MTL_CHECK_ARG(head_out && cls_targets && reg_targets && reg_weights && cls_weights && num_proposals && losses,
                "mtl_box_classifier_loss: null tensor");
  const long long warps = (long long)B * P;
  box_classifier_loss_kernel<<<(unsigned)ceil_div_ll(warps, 8), 256, 0, stream>>>(
      head_out, ld, box_col0, cls_col0, K, cls_targets, reinterpret_cast<const float4*>(reg_targets), reg_weights,
      cls_weights, num_proposals, B, P, loc_weight, cls_weight, losses, d_head, ldd);
  MTL_CUDA_LAUNCH_CHECK("box_classifier_loss_kernel");
  return MTL_OK;
}

extern "C" int mtl_softmax_ce(const float* logits, long long ld, int col0, int C, const float* soft_targets,
                              long long ldt, int tcol0, const int* hard_targets, const float* weights,
                              const int* num_proposals, int P, int per_image_norm, long long rows, float scale,
                              float* loss, float* d_logits, long long ldd, int dcol0, int accumulate,
                              cudaStream_t stream) {
  MTL_CHECK_ARG(logits && loss && (soft_targets || hard_targets), "mtl_softmax_ce: null tensor");
  if (rows == 0) return MTL_OK;
  softmax_ce_kernel<<<(unsigned)ceil_div_ll(rows, 8), 256, 0, stream>>>(
      logits, ld, col0, C, soft_targets, ldt, tcol0, hard_targets, weights, num_proposals, P, per_image_norm, rows,
      scale, loss, d_logits, ldd, dcol0, accumulate);
  MTL_CUDA_LAUNCH_CHECK("softmax_ce_kernel");
  return MTL_OK;
}

extern "C" int mtl_edgemask_fwd(const void* feat, long long npix, int C, const float* w, const float* bias,
                                float* act, cudaStream_t stream) {
  MTL_CHECK_ARG(feat && w && bias && act && C % 2 == 0, "mtl_edgemask_fwd: bad args");
  edgemask_fwd_kernel<<<(unsigned)ceil_div_ll(npix, 8), 256, 0, stream>>>(reinterpret_cast<const bf16*>(feat), npix,
                                                                         C, w, bias, act);
  MTL_CUDA_LAUNCH_CHECK("edgemask_fwd_kernel");
  return MTL_OK;
}

extern "C" int mtl_edgemask_loss(const float* act, int B, int H, int W, const float* gt, int oh, int ow,
                                 float loss_weight, float* loss, float* d_act, cudaStream_t stream) {
  MTL_CHECK_ARG(act && gt && loss, "mtl_edgemask_loss: null tensor");
  if (d_act) {
    cudaError_t e = cudaMemsetAsync(d_act, 0, sizeof(float) * (size_t)B * H * W * 2, stream);
    if (e != cudaSuccess) { mtl_set_error("mtl_edgemask_loss: memset: %s", cudaGetErrorString(e)); return MTL_ERR_CUDA; }
  }
  const long long total = (long long)B * oh * ow;
  const float scale = loss_weight / (float)total;          // reduce_mean over B*oh*ow
  edgemask_loss_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(act, B, H, W, gt, oh, ow, scale, loss,
                                                                             d_act);
  MTL_CUDA_LAUNCH_CHECK("edgemask_loss_kernel");
  return MTL_OK;
}

extern "C" int mtl_edgemask_bwd(const void* feat, long long npix, int C, const float* w, const float* act,
                                const float* d_act, float* dfeat, float* dw, float* dbias, cudaStream_t stream) {
  MTL_CHECK_ARG(feat && w && act && d_act && dw && dbias, "mtl_edgemask_bwd: null tensor");
  dim3 grid(ceil_div(C, 64), (unsigned)ceil_div_ll(npix, 256));
  edgemask_bwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(feat), npix, C, w, act, d_act, dfeat,
                                                dw, dbias);
  MTL_CUDA_LAUNCH_CHECK("edgemask_bwd_kernel");
  return MTL_OK;
}

extern "C" int mtl_fc_fwd(const float* x, long long ldx, const float* w, const float* bias, const float* res,
                          long long ldr, int M, int N, int K, float* y, long long ldy, cudaStream_t stream) {
  MTL_CHECK_ARG(x && w && y, "mtl_fc_fwd: null tensor");
  if (M == 0) return MTL_OK;
  fc_fwd_kernel<<<(unsigned)ceil_div_ll((long long)M * N, 256), 256, 0, stream>>>(x, ldx, w, bias, res, ldr, M, N, K,
                                                                                 y, ldy);
  MTL_CUDA_LAUNCH_CHECK("fc_fwd_kernel");
  return MTL_OK;
}

extern "C" int mtl_fc_bwd(const float* x, long long ldx, const float* w, const float* dy, long long ldy, int M,
                          int N, int K, float* dw, float* db, float* dx, long long lddx, cudaStream_t stream) {
  MTL_CHECK_ARG(x && dy && dw, "mtl_fc_bwd: null tensor");
  if (M == 0) return MTL_OK;
  fc_wgrad_kernel<<<(unsigned)ceil_div_ll((long long)N * (K + 1), 128), 128, 0, stream>>>(x, ldx, dy, ldy, M, N, K,
                                                                                         dw, db);
  MTL_CUDA_LAUNCH_CHECK("fc_wgrad_kernel");
  if (dx) {
    MTL_CHECK_ARG(w != nullptr, "mtl_fc_bwd: dx needs w");
    fc_dgrad_kernel<<<(unsigned)ceil_div_ll((long long)M * K, 256), 256, 0, stream>>>(dy, ldy, w, M, N, K, dx, lddx);
    MTL_CUDA_LAUNCH_CHECK("fc_dgrad_kernel");
  }
  return MTL_OK;
}

extern "C" int mtl_refine_concat(const float* org, long long ldo, int ocol0, const float* win, long long ldw,
                                 int wcol0, int E, const float* close, long long ldc, int ccol0, int rows, int K1,
                                 float* out, long long ldout, cudaStream_t stream) {
  MTL_CHECK_ARG(org && out && rows > 0, "mtl_refine_concat: bad args");
  refine_concat_kernel<<<rows, 128, 0, stream>>>(org, ldo, ocol0, win, ldw, wcol0, E, close, ldc, ccol0, rows, K1,
                                                 out, ldout);
  MTL_CUDA_LAUNCH_CHECK("refine_concat_kernel");
  return MTL_OK;
}

extern "C" int mtl_colsum(const void* dy, int is_fp32, long long ld, long long M, int N, float alpha, float* db,
                          cudaStream_t stream) {
  MTL_CHECK_ARG(dy && db, "mtl_colsum: null tensor");
  if (M == 0) return MTL_OK;
  int gy = (int)min((long long)64, ceil_div_ll(M, 64));
  dim3 grid(ceil_div(N, 32), gy);
  if (is_fp32)
    colsum_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(dy), ld, M, N, alpha, db);
  else
    colsum_kernel<bf16><<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(dy), ld, M, N, alpha, db);
  MTL_CUDA_LAUNCH_CHECK("colsum_kernel");
  return MTL_OK;
}
