// Fused multi-tensor optimizer over ONE flat parameter arena (sm_100a, HBM-bound).
//
// Reference semantics (under /root/reference/):
//   slim/deployment/model_deploy.py:198-307   total = sum(task losses)/num_clones + sum(L2 terms)
//   object_detection/trainer.py:389-419       gradient multipliers, per-tensor clip, apply_gradients
//   slim/learning.py:282-301                  clip_gradient_norms -> tf.clip_by_norm per tensor
//   object_detection/builders/optimizer_builder.py:49-53  MomentumOptimizer(lr, 0.9):
//        accum = momentum * accum + g ;  w -= lr * accum
//   slim l2_regularizer(weight)(w) = weight * 0.5 * sum(w^2)  ->  gradient weight * w
// Layout: params / grads / momentum are fp32 arenas of the same length; every tensor starts at
// a 64-element boundary.  A second arena holds the bf16 compute copy of every weight with the
// frozen batch-norm scale folded in (w_bf16[k, ...] = w[k, ...] * bn_scale[k]), which is what the
// tcgen05 conv engine consumes; it is refreshed by the same pass that updates the fp32 master.
// The arena is cut into fixed-size chunks; one CTA per chunk, so the grid is a multiple of the
// SM count for any realistic model (77 M parameters -> 18.8 k chunks).
#include "common.cuh"

struct mtl_tensor_desc {
  long long offset;      // element offset of the tensor in the arenas
  long long numel;
  long long row_len;     // elements per output channel (numel / K); 0 when there is no BN fold
  long long scale_off;   // offset of the per-output-channel fold scale in `fold_scales`, -1 if none
  float l2_weight;       // slim l2_regularizer weight (0 for biases)
  float grad_mult;       // gradient multiplier (trainer.py:389-402)
  int trainable;         // 0 = frozen (still counted in the L2 loss value, trap T4)
  int pad_;
};

struct mtl_chunk_desc {
  int tensor;
  int len;
  long long start;       // element offset inside the tensor
};

namespace {

constexpr int CHUNK = 4096;

__device__ __forceinline__ float block_sum256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    t = warp_sum(t);
  }
  return t;
}

// stats[t*2+0] += sum w^2 ; stats[t*2+1] += sum ((g*gscale + l2*w) * mult)^2
// (the reference multiplies the gradient of the TOTAL loss, regularisation included: trainer.py:387-402 acts on
//  grads_and_vars of optimize_clones, model_deploy.py:265-307)
__global__ void __launch_bounds__(256)
opt_stats_kernel(const mtl_tensor_desc* __restrict__ td, const mtl_chunk_desc* __restrict__ cd,
                 const float* __restrict__ params, const float* __restrict__ grads, float gscale,
                 float* __restrict__ stats, float* __restrict__ partials) {
  __shared__ float red[32];
  const mtl_chunk_desc c = cd[blockIdx.x];
  const mtl_tensor_desc t = td[c.tensor];
  const long long base = t.offset + c.start;
  float sw = 0.0f, sg = 0.0f;
  const float gm = gscale, tm = t.grad_mult, l2 = t.l2_weight;
  for (int i = threadIdx.x * 4; i < c.len; i += 256 * 4) {
    if (i + 4 <= c.len && (base & 3) == 0) {
      const float4 w = *reinterpret_cast<const float4*>(params + base + i);
      sw += w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w;
      if (t.trainable) {
        const float4 g = *reinterpret_cast<const float4*>(grads + base + i);
        const float a = (g.x * gm + l2 * w.x) * tm, b = (g.y * gm + l2 * w.y) * tm;
        const float cc = (g.z * gm + l2 * w.z) * tm, d = (g.w * gm + l2 * w.w) * tm;
        sg += a * a + b * b + cc * cc + d * d;
      }
    } else {
      for (int e = i; e < min(i + 4, c.len); ++e) {
        const float w = params[base + e];
        sw += w * w;
        if (t.trainable) {
          const float a = (grads[base + e] * gm + l2 * w) * tm;
          sg += a * a;
        }
      }
    }
  }
  const float tw = block_sum256(sw, red);
  const float tg = block_sum256(sg, red);
  if (threadIdx.x == 0) {
    if (partials) {          // deterministic mode: opt_stats_reduce_kernel sums the chunks in a fixed order
      partials[2 * blockIdx.x] = tw;
      partials[2 * blockIdx.x + 1] = t.trainable ? tg : 0.0f;
    } else {
      atomicAdd(stats + c.tensor * 2, tw);
      if (t.trainable) atomicAdd(stats + c.tensor * 2 + 1, tg);
    }
  }
}

// stats[t] = sum of the chunk partials of tensor t, one warp per tensor, fixed summation order: every
// data-parallel replica computes bit-identical clip factors from its (bit-identical) all-reduced gradients
__global__ void __launch_bounds__(256)
opt_stats_reduce_kernel(const int* __restrict__ chunk_start, int t0, int t1, int chunk0,
                        const float* __restrict__ partials, float* __restrict__ stats) {
  const int t = t0 + blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= t1) return;
  const int lane = threadIdx.x & 31;
  float a = 0.0f, b = 0.0f;
  for (int c = chunk_start[t] + lane; c < chunk_start[t + 1]; c += 32) {
    a += partials[2 * (c - chunk0)];
    b += partials[2 * (c - chunk0) + 1];
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) { stats[2 * t] = a; stats[2 * t + 1] = b; }
}

// reg_loss[0] = sum_t l2_t * 0.5 * sum w_t^2
__global__ void opt_reg_loss_kernel(const mtl_tensor_desc* __restrict__ td, int T, const float* __restrict__ stats,
                                    float* __restrict__ reg_loss) {
  __shared__ float red[32];
  float s = 0.0f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) s += td[t].l2_weight * 0.5f * stats[t * 2];
  const float tot = block_sum256(s, red);
  if (threadIdx.x == 0) reg_loss[0] = tot;
}

// hyper[0] = learning rate, hyper[1] = momentum, hyper[2] = clip norm (<= 0: off)
__global__ void __launch_bounds__(256)
opt_apply_kernel(const mtl_tensor_desc* __restrict__ td, const mtl_chunk_desc* __restrict__ cd,
                 float* __restrict__ params, float* __restrict__ grads, float* __restrict__ mom,
                 bf16* __restrict__ params_bf16, const float* __restrict__ fold_scales,
                 const float* __restrict__ stats, const float* __restrict__ hyper, float gscale,
                 float* __restrict__ wsq) {
  // wsq != null: also leave this chunk's sum of squares of the UPDATED weights in wsq[2 * block] (and 0 in the slot of the
  // gradient norm): what opt_stats_kernel would compute on the new weights, without a third pass over them
  __shared__ float red[32];
  const mtl_chunk_desc c = cd[blockIdx.x];
  const mtl_tensor_desc t = td[c.tensor];
  const long long base = t.offset + c.start;
  float sw = 0.0f;
  if (!t.trainable) {
    if (!wsq) return;
    for (int i = threadIdx.x; i < c.len; i += 256) { const float w = params[base + i]; sw += w * w; }
    const float tw = block_sum256(sw, red);
    if (threadIdx.x == 0) { wsq[2 * blockIdx.x] = tw; wsq[2 * blockIdx.x + 1] = 0.0f; }
    return;
  }
  const float lr = hyper[0], momentum = hyper[1], clip = hyper[2];
  float factor = 1.0f;
  if (clip > 0.0f) {
    const float norm = sqrtf(stats[c.tensor * 2 + 1]);
    factor = clip / fmaxf(norm, clip);                  // tf.clip_by_norm
  }
  const float gm = gscale, tm = t.grad_mult;
  if ((base & 3) == 0 && (c.len & 3) == 0 && (t.scale_off < 0 || (t.row_len & 3) == 0)) {
    // vector path: 4 parameters per thread, 16-byte loads / stores (8-byte for the bf16 copy)
    for (int i = threadIdx.x * 4; i < c.len; i += 256 * 4) {
      const long long o = base + i;
      const float4 w = *reinterpret_cast<const float4*>(params + o);
      const float4 g4 = *reinterpret_cast<const float4*>(grads + o);
      const float4 m4 = *reinterpret_cast<const float4*>(mom + o);
      const float sc = t.scale_off >= 0 ? fold_scales[t.scale_off + (c.start + i) / t.row_len] : 1.0f;
      float wv[4] = {w.x, w.y, w.z, w.w};
      const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
      float mv[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float g = (gv[e] * gm + t.l2_weight * wv[e]) * tm * factor;
        mv[e] = momentum * mv[e] + g;
        wv[e] = wv[e] - lr * mv[e];
      }
      *reinterpret_cast<float4*>(mom + o) = make_float4(mv[0], mv[1], mv[2], mv[3]);
      *reinterpret_cast<float4*>(params + o) = make_float4(wv[0], wv[1], wv[2], wv[3]);
      *reinterpret_cast<float4*>(grads + o) = make_float4(0.f, 0.f, 0.f, 0.f);
      sw += wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2] + wv[3] * wv[3];
      __nv_bfloat162 lo = __floats2bfloat162_rn(wv[0] * sc, wv[1] * sc);
      __nv_bfloat162 hi = __floats2bfloat162_rn(wv[2] * sc, wv[3] * sc);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(params_bf16 + o) = pk;
    }
  } else {
  for (int i = threadIdx.x; i < c.len; i += 256) {
    const long long o = base + i;
    const float w = params[o];
    const float g = (grads[o] * gm + t.l2_weight * w) * tm * factor;
    const float m = momentum * mom[o] + g;
    const float nw = w - lr * m;
    mom[o] = m;
    params[o] = nw;
    grads[o] = 0.0f;                                    // ready for the next step's += wgrad
    sw += nw * nw;
    float sc = 1.0f;
    if (t.scale_off >= 0) sc = fold_scales[t.scale_off + (c.start + i) / t.row_len];
    params_bf16[o] = __float2bfloat16_rn(nw * sc);
  }
  }
  if (wsq) {
    const float tw = block_sum256(sw, red);
    if (threadIdx.x == 0) { wsq[2 * blockIdx.x] = tw; wsq[2 * blockIdx.x + 1] = 0.0f; }
  }
}

// (re)build the bf16 compute copy of every tensor (after init / checkpoint load)
__global__ void __launch_bounds__(256)
opt_fold_kernel(const mtl_tensor_desc* __restrict__ td, const mtl_chunk_desc* __restrict__ cd,
                const float* __restrict__ params, bf16* __restrict__ params_bf16,
                const float* __restrict__ fold_scales) {
  const mtl_chunk_desc c = cd[blockIdx.x];
  const mtl_tensor_desc t = td[c.tensor];
  const long long base = t.offset + c.start;
  for (int i = threadIdx.x; i < c.len; i += 256) {
    float sc = 1.0f;
    if (t.scale_off >= 0) sc = fold_scales[t.scale_off + (c.start + i) / t.row_len];
    params_bf16[base + i] = __float2bfloat16_rn(params[base + i] * sc);
  }
}

// out = a * alpha  (fp32 -> bf16), used to hand the accumulated fp32 feature gradient to dgrad
__global__ void cast_f32_bf16_kernel(const float* __restrict__ a, long long n, float alpha, bf16* __restrict__ out) {
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < n;
       i += (long long)gridDim.x * blockDim.x * 4) {
    if (i + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(a + i);
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x * alpha, v.y * alpha);
      __nv_bfloat162 hi = __floats2bfloat162_rn(v.z * alpha, v.w * alpha);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(out + i) = pk;
    } else {
      for (long long e = i; e < n; ++e) out[e] = __float2bfloat16_rn(a[e] * alpha);
    }
  }
}

// out = mask > 0 ? (a [+ b]) : 0 ; fp32 accumulators + optional bf16 addend -> bf16
__global__ void relu_bwd_merge_kernel(const float* __restrict__ a, const bf16* __restrict__ b,
                                      const bf16* __restrict__ mask, long long n, bf16* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = a ? a[i] : 0.0f;
    if (b) v += __bfloat162float(b[i]);
    if (mask && !(__bfloat162float(mask[i]) > 0.0f)) v = 0.0f;
    out[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace

// ===================================================================================== C ABI
extern "C" int mtl_opt_chunk_size(void) { return CHUNK; }

extern "C" int mtl_opt_stats(const mtl_tensor_desc* tensors, int num_tensors, const mtl_chunk_desc* chunks,
                             int num_chunks, const float* params, const float* grads, float grad_scale,
                             float* stats, float* reg_loss, const int* chunk_start, float* partials,
                             cudaStream_t stream) {
  MTL_CHECK_ARG(tensors && chunks && params && grads && stats, "mtl_opt_stats: null tensor");
  MTL_CHECK_ARG(!partials || chunk_start, "mtl_opt_stats: partials need the chunk_start table");
  if (!partials) {
    cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(float) * 2 * num_tensors, stream);
    if (e != cudaSuccess) { mtl_set_error("mtl_opt_stats: memset: %s", cudaGetErrorString(e)); return MTL_ERR_CUDA; }
  }
  opt_stats_kernel<<<num_chunks, 256, 0, stream>>>(tensors, chunks, params, grads, grad_scale, stats, partials);
  MTL_CUDA_LAUNCH_CHECK("opt_stats_kernel");
  if (partials) {
    opt_stats_reduce_kernel<<<ceil_div(num_tensors, 8), 256, 0, stream>>>(chunk_start, 0, num_tensors, 0, partials, stats);
    MTL_CUDA_LAUNCH_CHECK("opt_stats_reduce_kernel");
  }
  if (reg_loss) {
    opt_reg_loss_kernel<<<1, 256, 0, stream>>>(tensors, num_tensors, stats, reg_loss);
    MTL_CUDA_LAUNCH_CHECK("opt_reg_loss_kernel");
  }
  return MTL_OK;
}

extern "C" int mtl_opt_stats_range(const mtl_tensor_desc* tensors, int t0, int t1, const mtl_chunk_desc* chunks,
                                   int num_chunks, const float* params, const float* grads, float grad_scale,
                                   float* stats, const int* chunk_start, int chunk0, float* partials,
                                   cudaStream_t stream) {
  MTL_CHECK_ARG(tensors && chunks && params && grads && stats && t0 >= 0 && t1 >= t0, "mtl_opt_stats_range: bad argument");
  MTL_CHECK_ARG(!partials || chunk_start, "mtl_opt_stats_range: partials need the chunk_start table");
  if (t1 == t0 || num_chunks == 0) return MTL_OK;
  if (!partials) {
    cudaError_t e = cudaMemsetAsync(stats + 2 * t0, 0, sizeof(float) * 2 * (t1 - t0), stream);
    if (e != cudaSuccess) { mtl_set_error("mtl_opt_stats_range: memset: %s", cudaGetErrorString(e)); return MTL_ERR_CUDA; }
  }
  opt_stats_kernel<<<num_chunks, 256, 0, stream>>>(tensors, chunks, params, grads, grad_scale, stats,
                                                   partials ? partials + 2 * (long long)chunk0 : nullptr);
  MTL_CUDA_LAUNCH_CHECK("opt_stats_kernel");
  if (partials) {
    opt_stats_reduce_kernel<<<ceil_div(t1 - t0, 8), 256, 0, stream>>>(chunk_start, t0, t1, 0, partials, stats);
    MTL_CUDA_LAUNCH_CHECK("opt_stats_reduce_kernel");
  }
  return MTL_OK;
}

extern "C" int mtl_opt_reg_loss(const mtl_tensor_desc* tensors, int num_tensors, const float* stats, float* reg_loss,
                                cudaStream_t stream) {
  MTL_CHECK_ARG(tensors && stats && reg_loss, "mtl_opt_reg_loss: null tensor");
  opt_reg_loss_kernel<<<1, 256, 0, stream>>>(tensors, num_tensors, stats, reg_loss);
  MTL_CUDA_LAUNCH_CHECK("opt_reg_loss_kernel");
  return MTL_OK;
}

extern "C" int mtl_opt_apply(const mtl_tensor_desc* tensors, const mtl_chunk_desc* chunks, int num_chunks,
                             float* params, float* grads, float* momentum, void* params_bf16,
                             const float* fold_scales, const float* stats, const float* hyper, float grad_scale,
                             cudaStream_t stream) {
  MTL_CHECK_ARG(tensors && chunks && params && grads && momentum && params_bf16 && stats && hyper,
                "mtl_opt_apply: null tensor");
  opt_apply_kernel<<<num_chunks, 256, 0, stream>>>(tensors, chunks, params, grads, momentum,
                                                   reinterpret_cast<bf16*>(params_bf16), fold_scales, stats, hyper,
                                                   grad_scale, nullptr);
  MTL_CUDA_LAUNCH_CHECK("opt_apply_kernel");
  return MTL_OK;
}

// mtl_opt_apply over the tensors [t0, t1) that also refreshes their squared weight norms (stats[2t], fixed summation
// order through `partials`) from the UPDATED weights; stats[2t + 1] (the gradient norm) is left at 0 until the next
// statistics pass.
extern "C" int mtl_opt_apply_norms(const mtl_tensor_desc* tensors, int t0, int t1, const mtl_chunk_desc* chunks,
                                   int num_chunks, float* params, float* grads, float* momentum, void* params_bf16,
                                   const float* fold_scales, float* stats, const float* hyper, float grad_scale,
                                   const int* chunk_start, int chunk0, float* partials, cudaStream_t stream) {
  MTL_CHECK_ARG(tensors && chunks && params && grads && momentum && params_bf16 && stats && hyper && chunk_start &&
                partials && t0 >= 0 && t1 >= t0, "mtl_opt_apply_norms: bad argument");
  if (t1 == t0 || num_chunks == 0) return MTL_OK;
  opt_apply_kernel<<<num_chunks, 256, 0, stream>>>(tensors, chunks, params, grads, momentum,
                                                   reinterpret_cast<bf16*>(params_bf16), fold_scales, stats, hyper,
                                                   grad_scale, partials + 2 * (long long)chunk0);
  MTL_CUDA_LAUNCH_CHECK("opt_apply_kernel");
  opt_stats_reduce_kernel<<<ceil_div(t1 - t0, 8), 256, 0, stream>>>(chunk_start, t0, t1, 0, partials, stats);
  MTL_CUDA_LAUNCH_CHECK("opt_stats_reduce_kernel");
  return MTL_OK;
}

extern "C" int mtl_opt_fold(const mtl_tensor_desc* tensors, const mtl_chunk_desc* chunks, int num_chunks,
                            const float* params, void* params_bf16, const float* fold_scales,
                            cudaStream_t stream) {
  MTL_CHECK_ARG(tensors && chunks && params && params_bf16, "mtl_opt_fold: null tensor");
  opt_fold_kernel<<<num_chunks, 256, 0, stream>>>(tensors, chunks, params, reinterpret_cast<bf16*>(params_bf16),
                                                  fold_scales);
  MTL_CUDA_LAUNCH_CHECK("opt_fold_kernel");
  return MTL_OK;
}

extern "C" int mtl_cast_f32_bf16(const float* a, long long n, float alpha, void* out, cudaStream_t stream) {
  MTL_CHECK_ARG(a && out, "mtl_cast_f32_bf16: null tensor");
  if (n == 0) return MTL_OK;
  const int grid = (int)min(ceil_div_ll(n, 1024), (long long)mtl_num_sms() * 16);
  cast_f32_bf16_kernel<<<grid, 256, 0, stream>>>(a, n, alpha, reinterpret_cast<bf16*>(out));
  MTL_CUDA_LAUNCH_CHECK("cast_f32_bf16_kernel");
  return MTL_OK;
}

extern "C" int mtl_relu_bwd_merge(const float* a, const void* b, const void* mask, long long n, void* out,
                                  cudaStream_t stream) {
  MTL_CHECK_ARG((a || b) && out, "mtl_relu_bwd_merge: null tensor");
  if (n == 0) return MTL_OK;
  const int grid = (int)min(ceil_div_ll(n, 256), (long long)mtl_num_sms() * 16);
  relu_bwd_merge_kernel<<<grid, 256, 0, stream>>>(a, reinterpret_cast<const bf16*>(b),
                                                  reinterpret_cast<const bf16*>(mask), n,
                                                  reinterpret_cast<bf16*>(out));
  MTL_CUDA_LAUNCH_CHECK("relu_bwd_merge_kernel");
  return MTL_OK;
}
