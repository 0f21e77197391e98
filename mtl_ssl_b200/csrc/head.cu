// Fused second-stage heads of MaskRCNNBoxPredictor (one kernel per ROI batch and direction):
//   forward   spatial average over the ROI grid + fully connected layer(s)      pooled[R,C], logits[R,n]
//   backward  gradient of the logits -> bf16 copy for the weight-gradient GEMM, bias gradient, gradient of the pooled
//             features (dlogits x W) and its broadcast back over the ROI grid under the ReLU mask
// Reference: /root/reference/object_detection/core/box_predictor.py:470-500, :568-602 (tf.reduce_mean over [1, 2] with
// keep_dims, slim.flatten, slim.fully_connected for box encodings / class scores; `predict_class` for the auxiliary
// heads).  They replace avgpool_fwd + a tcgen05 GEMM with 8-456 output columns (forward) and cast + colsum + GEMM +
// avgpool_bwd (backward): six dependent launches per head on the critical path between the box-classifier tails and
// their backward pass.  The FC is far too narrow for a 128 x 256 tensor-core tile (R = 64..1280 rows, 24..456 columns,
// 0.03-0.5 GFLOP): CUDA cores, weights streamed from L2 (heads wider than 128 outputs, i.e. the 90-class box head, keep
// the tensor-core GEMM: core/box_predictor.py).
// Rounding points are those of the kernels replaced: pooled features and the pooled-feature gradient are bf16.
#include "common.cuh"

namespace {

constexpr int HG = 1;            // ROIs per block (1: as many blocks as the pooling kernel it replaces; the FC weights
                                 // are re-read from L2 by every block: R x n x C x 2 bytes, 0.1 GB for 256 x 104 x 2048)

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

// x [R,HW,C] bf16, w [n,C] bf16 (rows = outputs), bias [n] fp32 -> pooled [R,C] bf16, out [R,ldo] fp32 (columns 0..n-1)
// PART: instead of x, the partial row sums of the tail's last conv (mtl_conv_args.pool_out): part[(2g + s) * C + c] =
// sum over the rows of 32-row group g in its first / second ROI; a ROI (HW >= 32 rows) meets two or three groups,
// summed here in ascending order.
template <bool PART>
__global__ void __launch_bounds__(256)
head_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ part, int R, int HW, int C,
                const bf16* __restrict__ w, const float* __restrict__ bias, int n, bf16* __restrict__ pooled,
                float* __restrict__ out, long long ldo) {
  extern __shared__ float sp[];                     // [HG][C] pooled features (bf16-rounded values)
  const int nvec = C >> 3;
  const long long r0 = (long long)blockIdx.x * HG;
  const float inv = 1.0f / (float)HW;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
#pragma unroll
    for (int g = 0; g < HG; ++g) {
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
      if (r0 + g < R) {
        if (PART) {
          const long long row0 = (r0 + g) * HW;
          for (long long gi = row0 >> 5; gi <= (row0 + HW - 1) >> 5; ++gi) {
            const int slot = ((gi << 5) / HW == r0 + g) ? 0 : 1;      // the group starts inside this ROI or in the previous one
            const float4* src = reinterpret_cast<const float4*>(part + (2 * gi + slot) * C) + 2 * v;
            const float4 a = __ldg(src), b = __ldg(src + 1);
            acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
            acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
          }
        } else {
        const uint4* src = reinterpret_cast<const uint4*>(x + (r0 + g) * HW * C) + v;
        for (int p = 0; p < HW; ++p) {
          float f[8];
          unpack8(__ldg(src + (long long)p * nvec), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += f[e];
        }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] *= inv;
        const uint4 pk = pack8(acc);
        reinterpret_cast<uint4*>(pooled + (r0 + g) * C)[v] = pk;
        unpack8(pk, acc);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) sp[g * C + v * 8 + e] = acc[e];
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int o = warp; o < n; o += nwarps) {
    float acc[HG];
#pragma unroll
    for (int g = 0; g < HG; ++g) acc[g] = 0.0f;
    const uint4* wr = reinterpret_cast<const uint4*>(w + (long long)o * C);
    for (int v = lane; v < nvec; v += 32) {
      float f[8];
      unpack8(__ldg(wr + v), f);
#pragma unroll
      for (int g = 0; g < HG; ++g) {
        const float4 a = *reinterpret_cast<const float4*>(sp + g * C + v * 8);
        const float4 b = *reinterpret_cast<const float4*>(sp + g * C + v * 8 + 4);
        acc[g] += f[0] * a.x + f[1] * a.y + f[2] * a.z + f[3] * a.w + f[4] * b.x + f[5] * b.y + f[6] * b.z + f[7] * b.w;
      }
    }
#pragma unroll
    for (int g = 0; g < HG; ++g) {
      const float s = warp_sum(acc[g]);
      if (lane == 0 && r0 + g < R) out[(r0 + g) * ldo + o] = s + (bias ? bias[o] : 0.0f);
    }
  }
}

// d_out [R,ldd] fp32 (columns 0..n-1) -> dyb [R,n] bf16, db[n] += column sums of dyb, and (dx != null)
// dx[r,p,c] = mask(x[r,p,c]) * bf16(sum_o dyb[r,o] * w[o,c]) / HW
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ d_out, long long ldd, int n, const bf16* __restrict__ w,
                const bf16* __restrict__ x, float mask_hi, int R, int HW, int C, bf16* __restrict__ dyb,
                float* __restrict__ db, bf16* __restrict__ dx) {
  extern __shared__ float sd[];                     // [HG][n] logit gradients (bf16-rounded values)
  const long long r0 = (long long)blockIdx.x * HG;
  for (int i = threadIdx.x; i < HG * n; i += blockDim.x) {
    const int g = i / n, o = i - g * n;
    float v = 0.0f;
    if (r0 + g < R) {
      const bf16 b = __float2bfloat16_rn(d_out[(r0 + g) * ldd + o]);
      dyb[(r0 + g) * n + o] = b;
      v = __bfloat162float(b);
    }
    sd[i] = v;
  }
  __syncthreads();
  if (db) {
    for (int o = threadIdx.x; o < n; o += blockDim.x) {
      float s = 0.0f;
#pragma unroll
      for (int g = 0; g < HG; ++g) s += sd[g * n + o];
      if (s != 0.0f) atomicAdd(db + o, s);
    }
  }
  if (!dx) return;
  const int nvec = C >> 3;
  const float inv = 1.0f / (float)HW;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float acc[HG][8];
#pragma unroll
    for (int g = 0; g < HG; ++g)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[g][e] = 0.0f;
    for (int o = 0; o < n; ++o) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(w + (long long)o * C) + v), f);
#pragma unroll
      for (int g = 0; g < HG; ++g) {
        const float d = sd[g * n + o];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[g][e] += d * f[e];
      }
    }
#pragma unroll
    for (int g = 0; g < HG; ++g) {
      if (r0 + g >= R) continue;
      float dp[8];
      unpack8(pack8(acc[g]), dp);                   // the pooled-feature gradient is a bf16 tensor
#pragma unroll
      for (int e = 0; e < 8; ++e) dp[e] *= inv;
      const long long base = (r0 + g) * HW;
      for (int p = 0; p < HW; ++p) {
        float gg[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) gg[e] = dp[e];
        if (x) {
          float m[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(x + (base + p) * C) + v), m);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (!(m[e] > 0.0f) || (mask_hi > 0.0f && !(m[e] < mask_hi))) gg[e] = 0.0f;
        }
        reinterpret_cast<uint4*>(dx + (base + p) * C)[v] = pack8(gg);
      }
    }
  }
}

}  // namespace

extern "C" int mtl_head_fwd(const void* x, int R, int HW, int C, const void* w, const float* bias, int n, void* pooled,
                            float* out, long long ldo, cudaStream_t stream) {
  MTL_CHECK_ARG(x && w && pooled && out, "mtl_head_fwd: null tensor");
  MTL_CHECK_ARG(C % 8 == 0 && C > 0 && n > 0 && HW > 0 && ldo >= n, "mtl_head_fwd: bad geometry (C=%d n=%d)", C, n);
  if (R == 0) return MTL_OK;
  const size_t smem = sizeof(float) * HG * (size_t)C;
  MTL_CHECK_ARG(smem <= 48 * 1024, "mtl_head_fwd: C too large (%d)", C);
  head_fwd_kernel<false><<<(unsigned)ceil_div(R, HG), 256, smem, stream>>>(
      reinterpret_cast<const bf16*>(x), nullptr, R, HW, C, reinterpret_cast<const bf16*>(w), bias, n,
      reinterpret_cast<bf16*>(pooled), out, ldo);
  MTL_CUDA_LAUNCH_CHECK("head_fwd_kernel");
  return MTL_OK;
}

extern "C" int mtl_head_fwd_pooled(const float* part, int R, int HW, int C, const void* w, const float* bias, int n,
                                   void* pooled, float* out, long long ldo, cudaStream_t stream) {
  MTL_CHECK_ARG(part && w && pooled && out, "mtl_head_fwd_pooled: null tensor");
  MTL_CHECK_ARG(C % 8 == 0 && C > 0 && n > 0 && HW >= 32 && ldo >= n, "mtl_head_fwd_pooled: bad geometry (C=%d n=%d HW=%d)",
                C, n, HW);
  if (R == 0) return MTL_OK;
  const size_t smem = sizeof(float) * HG * (size_t)C;
  MTL_CHECK_ARG(smem <= 48 * 1024, "mtl_head_fwd_pooled: C too large (%d)", C);
  head_fwd_kernel<true><<<(unsigned)ceil_div(R, HG), 256, smem, stream>>>(
      nullptr, part, R, HW, C, reinterpret_cast<const bf16*>(w), bias, n, reinterpret_cast<bf16*>(pooled), out, ldo);
  MTL_CUDA_LAUNCH_CHECK("head_fwd_kernel<pooled>");
  return MTL_OK;
}

extern "C" int mtl_head_bwd(const float* d_out, long long ldd, int n, const void* w, const void* x, float mask_hi, int R,
                            int HW, int C, void* dyb, float* db, void* dx, cudaStream_t stream) {
  MTL_CHECK_ARG(d_out && w && dyb, "mtl_head_bwd: null tensor");
  MTL_CHECK_ARG(C % 8 == 0 && C > 0 && n > 0 && HW > 0 && ldd >= n, "mtl_head_bwd: bad geometry (C=%d n=%d)", C, n);
  if (R == 0) return MTL_OK;
  const size_t smem = sizeof(float) * HG * (size_t)n;
  MTL_CHECK_ARG(smem <= 48 * 1024, "mtl_head_bwd: too many outputs (%d)", n);
  head_bwd_kernel<<<(unsigned)ceil_div(R, HG), 256, smem, stream>>>(
      d_out, ldd, n, reinterpret_cast<const bf16*>(w), reinterpret_cast<const bf16*>(x), mask_hi, R, HW, C,
      reinterpret_cast<bf16*>(dyb), db, reinterpret_cast<bf16*>(dx));
  MTL_CUDA_LAUNCH_CHECK("head_bwd_kernel");
  return MTL_OK;
}
