// Box-geometry kernels of the RPN / target-assignment path (sm_100a, compiled with -fmad=false
// so that every float op rounds exactly like the NumPy float32 restatement in oracle/).
//
// Reference semantics (all under /root/reference/object_detection/):
//   anchors            anchor_generators/grid_anchor_generator.py:96-214
//   prune              core/box_list_ops.py:140-169, meta_architectures/faster_rcnn_meta_arch.py:930-976
//   decode/clip/filter box_coders/faster_rcnn_box_coder.py:92-118, core/box_list_ops.py:102-137, 652-687
//   NMS                core/post_processing.py:25-164 -> tf.image.non_max_suppression (TF 1.7)
//   IoU / matcher      core/box_list_ops.py:201-272, matchers/argmax_matcher.py:102-175
//   sampler            core/balanced_positive_negative_sampler.py:50-91, core/minibatch_sampler.py:64-90
//   targets            core/target_assigner.py:99-213, 256-403
// There are no reference kernels (the reference runs these as chains of TF ops, several on
// the CPU); the kernels below are new: warp-ballot compaction, packed-key rank sort,
// chunked greedy NMS with a shared-memory kept list, packed atomicMax row arg-max.
#include "common.cuh"
#include <string.h>

namespace {

typedef unsigned long long u64;

// ------------------------------------------------------------------ ordered block compaction
// Returns the exclusive prefix of `flag` over the block (blockDim.x multiple of 32, <= 1024)
// and the block total in *total.  Uses one ballot per warp and a 32-entry scan.
__device__ __forceinline__ int block_excl_scan_flag(bool flag, int* total, int* warp_sums /*[33]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const unsigned b = __ballot_sync(0xffffffffu, flag);
  const int inwarp = __popc(b & ((1u << lane) - 1u));
  __syncthreads();   // protect warp_sums reuse across calls
  if (lane == 0) warp_sums[warp] = __popc(b);
  __syncthreads();
  if (warp == 0) {
    int v = lane < nwarps ? warp_sums[lane] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    warp_sums[lane] = incl - v;
    if (lane == 31) warp_sums[32] = incl;
  }
  __syncthreads();
  *total = warp_sums[32];
  return warp_sums[warp] + inwarp;
}

// ------------------------------------------------------------------ anchors
struct AnchorSpec {
  float scales[16];
  float ars[16];
  int ns, na;
  float base_h, base_w, stride_h, stride_w, off_h, off_w;
};

__global__ void grid_anchors_kernel(int Hf, int Wf, AnchorSpec s, float* __restrict__ out) {
  const int A = s.ns * s.na;
  const long long total = (long long)Hf * Wf * A;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(i % A);
    const long long loc = i / A;
    const int x = (int)(loc % Wf), y = (int)(loc / Wf);
    const int ai = a / s.ns, si = a - ai * s.ns;      // a = aspect_idx * len(scales) + scale_idx
    const float r = sqrtf(s.ars[ai]);
    const float h = s.scales[si] / r * s.base_h;
    const float w = s.scales[si] * r * s.base_w;
    const float cy = (float)y * s.stride_h + s.off_h;
    const float cx = (float)x * s.stride_w + s.off_w;
    float4 o;
    o.x = cy - 0.5f * h; o.y = cx - 0.5f * w; o.z = cy + 0.5f * h; o.w = cx + 0.5f * w;
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

// keep boxes fully inside the window; ordered compaction by ONE block (run once per shape)
__global__ void prune_outside_window_kernel(const float4* __restrict__ boxes, int N, float wy0, float wx0,
                                            float wy1, float wx1, int* __restrict__ keep_idx,
                                            float4* __restrict__ kept, int* __restrict__ num_keep) {
  __shared__ int wsum[33];
  int base = 0;
  for (int start = 0; start < N; start += blockDim.x) {
    const int i = start + threadIdx.x;
    bool ok = false;
    float4 b = make_float4(0, 0, 0, 0);
    if (i < N) {
      b = boxes[i];
      const bool viol = (b.x < wy0) || (b.y < wx0) || (b.z > wy1) || (b.w > wx1);
      ok = !viol;
    }
    int total;
    const int pos = block_excl_scan_flag(ok, &total, wsum);
    if (ok) {
      keep_idx[base + pos] = i;
      kept[base + pos] = b;
    }
    base += total;
  }
  if (threadIdx.x == 0) *num_keep = base;
}

// ------------------------------------------------------------------ RPN decode + score + clip + key
__global__ void rpn_decode_kernel(const float* __restrict__ rpn_out, long long ld, int box_col0, int cls_col0,
                                  int A, int HW, const int* __restrict__ keep_idx,
                                  const float4* __restrict__ anchors, int Nk, float img_h, float img_w,
                                  float score_thresh, float4* __restrict__ boxes, float* __restrict__ scores,
                                  u64* __restrict__ keys) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Nk) return;
  const int ai = keep_idx ? keep_idx[j] : j;
  const int loc = ai / A, a = ai - loc * A;
  const float* row = rpn_out + ((long long)b * HW + loc) * ld;
  const float4 an = anchors[j];
  // get_center_coordinates_and_sizes (box_list.py:172-183)
  const float wa = an.w - an.y, ha = an.z - an.x;
  const float yca = an.x + ha / 2.0f, xca = an.y + wa / 2.0f;
  const float ty = row[box_col0 + a * 4 + 0] / 10.0f;
  const float tx = row[box_col0 + a * 4 + 1] / 10.0f;
  const float th = row[box_col0 + a * 4 + 2] / 5.0f;
  const float tw = row[box_col0 + a * 4 + 3] / 5.0f;
  const float w = expf(tw) * wa;
  const float h = expf(th) * ha;
  const float yc = ty * ha + yca;
  const float xc = tx * wa + xca;
  float ymin = yc - h / 2.0f, xmin = xc - w / 2.0f, ymax = yc + h / 2.0f, xmax = xc + w / 2.0f;
  // softmax over (background, object), keep the object probability
  const float l0 = row[cls_col0 + a * 2], l1 = row[cls_col0 + a * 2 + 1];
  const float m = fmaxf(l0, l1);
  const float e0 = expf(l0 - m), e1 = expf(l1 - m);
  const float sc = e1 / (e0 + e1);
  // clip_to_window [0,0,H,W] (box_list_ops.py:102-137)
  ymin = fmaxf(fminf(ymin, img_h), 0.0f);
  ymax = fmaxf(fminf(ymax, img_h), 0.0f);
  xmin = fmaxf(fminf(xmin, img_w), 0.0f);
  xmax = fmaxf(fminf(xmax, img_w), 0.0f);
  const float area = (ymax - ymin) * (xmax - xmin);
  const bool valid = (sc > score_thresh) && (area > 0.0f);
  const long long o = (long long)b * Nk + j;
  boxes[o] = make_float4(ymin, xmin, ymax, xmax);
  scores[o] = sc;
  // descending score, ties -> lower index first; 0 marks "filtered out"
  keys[o] = valid ? (((u64)__float_as_uint(sc) << 32) | (u64)(0xffffffffu - (unsigned)j)) : 0ull;
}

// ------------------------------------------------------------------ inference-time proposal path
// fmA:586-590: anchors are clipped to the image window instead of pruned (box_list_ops.clip_to_window; every grid
// anchor has its centre inside the image, so none is dropped).  One thread per box, run once per shape.
__global__ void clip_boxes_kernel(const float4* __restrict__ boxes, int N, float wy0, float wx0, float wy1, float wx1,
                                  float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float4 b = boxes[i];
  out[i] = make_float4(fmaxf(fminf(b.x, wy1), wy0), fmaxf(fminf(b.y, wx1), wx0), fmaxf(fminf(b.z, wy1), wy0),
                       fmaxf(fminf(b.w, wx1), wx0));
}

// fmA:1111-1131 without the training-time minibatch sampling: the NMS output IS the proposal set (zero padded to
// max_proposals); normalise by the image size (to_normalized_coordinates, check_range=False).
__global__ void proposals_from_nms_kernel(const float4* __restrict__ nms_boxes, const float* __restrict__ nms_scores,
                                          const int* __restrict__ nms_num, int M, float img_h, float img_w,
                                          float4* __restrict__ out_abs, float4* __restrict__ out_norm,
                                          float* __restrict__ out_scores, int* __restrict__ num_out) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) num_out[b] = nms_num[b];
  if (j >= M) return;
  const long long o = (long long)b * M + j;
  // same arithmetic as the training path (gather_sampled): scale by the reciprocal, absolute = round trip
  // (box_list_ops.to_normalized_coordinates / ops.normalized_to_image_coordinates, fmA:682-683, :1124-1131)
  const float ys = 1.0f / img_h, xs = 1.0f / img_w;
  const float4 v = nms_boxes[o];
  const float4 nr = make_float4(ys * v.x, xs * v.y, ys * v.z, xs * v.w);
  out_norm[o] = nr;
  out_abs[o] = make_float4(img_h * nr.x, img_w * nr.y, img_h * nr.z, img_w * nr.w);
  if (out_scores) out_scores[o] = nms_scores[o];
}

// ------------------------------------------------------------------ second-stage detections
// meta_architectures/faster_rcnn_meta_arch.py:1387-1469 (_postprocess_box_classifier) up to the per-class NMS
// input: decode the per-class refined encodings against the proposals (box coder :92-118), convert the class
// logits (SOFTMAX / SIGMOID / IDENTITY), drop the background column, then per class (post_processing.py:106-143)
// filter score > threshold, clip to the image window, drop zero-area boxes, change to the window frame.
// Layout out: class-major [B, K, P] so that (image, class) pairs are the batch dimension of the NMS kernels.
__global__ void detection_decode_kernel(const float* __restrict__ enc, const float* __restrict__ logits,
                                        const float4* __restrict__ proposals, const int* __restrict__ num_props,
                                        int P, int K, float img_h, float img_w, float score_thresh, int score_mode,
                                        float4* __restrict__ boxes_n, float* __restrict__ scores,
                                        u64* __restrict__ keys, float4* __restrict__ decoded_abs) {
  const int b = blockIdx.z, k = blockIdx.y;
  const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (pidx >= P) return;
  const long long r = (long long)b * P + pidx;                 // row of the [B*P, ...] head outputs
  const float4 an = proposals[r];
  const float wa = an.w - an.y, ha = an.z - an.x;
  const float yca = an.x + ha / 2.0f, xca = an.y + wa / 2.0f;
  const float* e = enc + (r * K + k) * 4;
  const float ty = e[0] / 10.0f, tx = e[1] / 10.0f, th = e[2] / 5.0f, tw = e[3] / 5.0f;
  const float w = expf(tw) * wa, h = expf(th) * ha;
  const float yc = ty * ha + yca, xc = tx * wa + xca;
  float ymin = yc - h / 2.0f, xmin = xc - w / 2.0f, ymax = yc + h / 2.0f, xmax = xc + w / 2.0f;
  const long long o = ((long long)b * K + k) * P + pidx;
  if (decoded_abs) decoded_abs[o] = make_float4(ymin, xmin, ymax, xmax);
  const float* l = logits + r * (K + 1);
  float sc;
  if (score_mode == 1) {                 // SOFTMAX over background + classes
    float m = l[0];
    for (int c = 1; c <= K; ++c) m = fmaxf(m, l[c]);
    float sum = 0.0f;
    for (int c = 0; c <= K; ++c) sum += expf(l[c] - m);
    sc = expf(l[k + 1] - m) / sum;
  } else if (score_mode == 2) {          // SIGMOID
    sc = 1.0f / (1.0f + expf(-l[k + 1]));
  } else {
    sc = l[k + 1];
  }
  ymin = fmaxf(fminf(ymin, img_h), 0.0f);
  ymax = fmaxf(fminf(ymax, img_h), 0.0f);
  xmin = fmaxf(fminf(xmin, img_w), 0.0f);
  xmax = fmaxf(fminf(xmax, img_w), 0.0f);
  const float area = (ymax - ymin) * (xmax - xmin);
  const bool valid = pidx < num_props[b] && sc > score_thresh && area > 0.0f;
  boxes_n[o] = make_float4(ymin / img_h, xmin / img_w, ymax / img_h, xmax / img_w);
  scores[o] = sc;
  keys[o] = valid ? (((u64)__float_as_uint(sc) << 32) | (u64)(0xffffffffu - (unsigned)pidx)) : 0ull;
}

// concatenate the per-class NMS survivors in class order (post_processing.py:144-148): key of candidate
// (class k, rank r) = (score, position k*M + r) so that sort_by_field's ties keep the concatenation order
__global__ void detection_merge_keys_kernel(const float* __restrict__ cls_scores, const int* __restrict__ cls_num,
                                            int K, int M, u64* __restrict__ keys) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= K * M) return;
  const int k = j / M, r = j - k * M;
  const bool valid = r < cls_num[b * K + k];
  const float sc = cls_scores[((long long)b * K + k) * M + r];
  keys[(long long)b * K * M + j] = valid ? (((u64)__float_as_uint(sc) << 32) | (u64)(0xffffffffu - (unsigned)j)) : 0ull;
}

// top max_total of the merged list, zero padded (post_processing.py:149-164, :281-312)
__global__ void detection_gather_kernel(const float4* __restrict__ cls_boxes, const float* __restrict__ cls_scores,
                                        const int* __restrict__ order, const int* __restrict__ num_valid, int K,
                                        int M, int T, float4* __restrict__ det_boxes, float* __restrict__ det_scores,
                                        float* __restrict__ det_classes, float* __restrict__ num_det) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(num_valid[b], T);
  if (t == 0) num_det[b] = (float)n;
  if (t >= T) return;
  const long long o = (long long)b * T + t;
  if (t < n) {
    const int j = order[(long long)b * K * M + t];
    const long long src = (long long)b * K * M + j;
    det_boxes[o] = cls_boxes[src];
    det_scores[o] = cls_scores[src];
    det_classes[o] = (float)(j / M);
  } else {
    det_boxes[o] = make_float4(0, 0, 0, 0);
    det_scores[o] = 0.0f;
    det_classes[o] = 0.0f;
  }
}

// score/box inputs that are already decoded (generic NMS front end): key construction only
__global__ void make_keys_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, int N,
                                 float score_thresh, int require_area, u64* __restrict__ keys) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const long long o = (long long)b * N + j;
  const float4 bx = boxes[o];
  const float sc = scores[o];
  const float area = (bx.z - bx.x) * (bx.w - bx.y);
  const bool valid = (sc > score_thresh) && (!require_area || area > 0.0f);
  keys[o] = valid ? (((u64)__float_as_uint(sc) << 32) | (u64)(0xffffffffu - (unsigned)j)) : 0ull;
}

// ------------------------------------------------------------------ rank sort (descending, unique keys)
// rank_i = #{j : key_j > key_i}; order[rank_i] = i for valid (non-zero) keys.  O(N^2) compares
// spread over the whole GPU: 14 k keys -> 0.2 G compares, far cheaper than a multi-pass sort's launches.
constexpr int RS_THREADS = 256;
constexpr int RS_TILE = 2048;
constexpr int RS_SPLIT = 16;      // the comparison range is cut in RS_SPLIT slices (grid z): 14 k keys -> 880 blocks, six
                                  // per SM, so the dependent load-compare-add chains of 48 warps hide each other's latency
__global__ void __launch_bounds__(RS_THREADS)
rank_count_kernel(const u64* __restrict__ keys, int N, int* __restrict__ rank) {
  __shared__ __align__(16) u64 tile[RS_TILE];
  const int b = blockIdx.y;
  const u64* k = keys + (long long)b * N;
  const int i = blockIdx.x * RS_THREADS + threadIdx.x;
  const u64 mine = i < N ? k[i] : 0ull;
  const int per = (N + RS_SPLIT - 1) / RS_SPLIT;
  const int j0 = blockIdx.z * per, j1 = min(j0 + per, N);
  int r = 0;
  for (int t0 = j0; t0 < j1; t0 += RS_TILE) {
    const int n = min(RS_TILE, j1 - t0);
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += RS_THREADS) tile[t] = k[t0 + t];
    __syncthreads();
    if (mine != 0ull) {
      int t = 0, r1 = 0;
      for (; t + 8 <= n; t += 8) {        // two independent accumulators, 16-byte broadcast loads
        const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(tile + t);
        const ulonglong2 c = *reinterpret_cast<const ulonglong2*>(tile + t + 2);
        const ulonglong2 d = *reinterpret_cast<const ulonglong2*>(tile + t + 4);
        const ulonglong2 e = *reinterpret_cast<const ulonglong2*>(tile + t + 6);
        r += (a.x > mine) + (a.y > mine) + (c.x > mine) + (c.y > mine);
        r1 += (d.x > mine) + (d.y > mine) + (e.x > mine) + (e.y > mine);
      }
      r += r1;
      for (; t < n; ++t) r += (tile[t] > mine);
    }
  }
  if (mine != 0ull && r) atomicAdd(rank + (long long)b * N + i, r);
}

__global__ void rank_scatter_kernel(const u64* __restrict__ keys, const int* __restrict__ rank, int N,
                                    int* __restrict__ order, int* __restrict__ num_valid) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < N && keys[(long long)b * N + i] != 0ull;
  if (valid) order[(long long)b * N + rank[(long long)b * N + i]] = i;
  const unsigned bal = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(num_valid + b, __popc(bal));
}

// ------------------------------------------------------------------ greedy NMS
// TF 1.7 NonMaxSuppression: IoU with corner order normalised, 0 if either area <= 0.
__device__ __forceinline__ bool nms_iou_gt(const float4 a, float area_a, const float4 b, float area_b, float thr) {
  if (area_a <= 0.0f || area_b <= 0.0f) return false;
  const float iy0 = fmaxf(a.x, b.x), ix0 = fmaxf(a.y, b.y);
  const float iy1 = fminf(a.z, b.z), ix1 = fminf(a.w, b.w);
  const float inter = fmaxf(iy1 - iy0, 0.0f) * fmaxf(ix1 - ix0, 0.0f);
  const float uni = area_a + area_b - inter;        // >= max(area) > 0 up to rounding
  // inter / uni > thr, decided without the division unless the quotient lies within 1e-6 (relative) of the threshold:
  // the IEEE quotient and thr * uni each carry 6e-8 of rounding, so outside that band both tests agree exactly.
  // (NaNs fail both comparisons and take the division, as before.)
  const float t = thr * uni;
  if (inter > t * 1.000001f) return true;
  if (inter < t * 0.999999f) return false;
  return inter / uni > thr;
}

constexpr int NMS_CHUNK = 256;        // candidates per chunk
constexpr int NMS_THREADS = 1024;     // 4 threads per candidate
constexpr int NMS_MAX_OUT = 1024;
// One block per image.  Candidates are visited in score order in chunks of 256, four threads each:
//   1. the four threads of a candidate test it against interleaved quarters of the kept list of
//      earlier chunks (shared memory);
//   2. they compute the candidate's row of the chunk's upper-triangular suppression bit matrix
//      (8 x 32-bit words in shared memory, two words per thread);
//   3. ONE warp resolves the chunk in order with the live mask in registers (lane l owns word l):
//      the lowest live bit is kept and its matrix row is and-not'ed into the mask -- no block barriers
//      in the serial part and only kept boxes cost an iteration;
//   4. the newly kept boxes are appended to the kept list / outputs in parallel.
// Stops at max_out kept.  Same decisions, in the same order, as TF 1.7 NonMaxSuppression.
__global__ void __launch_bounds__(NMS_THREADS)
nms_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const int* __restrict__ order,
           const int* __restrict__ num_valid, int N, float thr, int max_out, float4* __restrict__ out_boxes,
           float* __restrict__ out_scores, int* __restrict__ out_idx, int* __restrict__ num_out) {
  __shared__ float4 kb[NMS_MAX_OUT];
  __shared__ float ka[NMS_MAX_OUT];
  __shared__ float4 cb[NMS_CHUNK];
  __shared__ float ca[NMS_CHUNK];
  __shared__ int cidx[NMS_CHUNK];
  __shared__ int dead[NMS_CHUNK];
  __shared__ unsigned mask[NMS_CHUNK][NMS_CHUNK / 32];
  __shared__ int newlist[NMS_CHUNK];
  __shared__ int s_new;
  const int b = blockIdx.x;
  const float4* bx = boxes + (long long)b * N;
  const int* ord = order + (long long)b * N;
  const int nv = min(num_valid[b], N);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int ci = t & (NMS_CHUNK - 1), part = t >> 8;       // candidate in chunk, quarter 0..3
  constexpr int W = NMS_CHUNK / 32;
  int nkept = 0;
  for (int pos = 0; pos < nv && nkept < max_out; pos += NMS_CHUNK) {
    const int c = pos + ci;
    const int nc = min(NMS_CHUNK, nv - pos);
    const bool valid = c < nv;
    float4 me = make_float4(0, 0, 0, 0);
    float area = 0.0f;
    int idx = -1;
    if (valid) {
      idx = ord[c];
      const float4 r = bx[idx];
      me = make_float4(fminf(r.x, r.z), fminf(r.y, r.w), fmaxf(r.x, r.z), fmaxf(r.y, r.w));
      area = (me.z - me.x) * (me.w - me.y);
    }
    if (part == 0) { cb[ci] = me; ca[ci] = area; cidx[ci] = idx; dead[ci] = valid ? 0 : 1; }
    __syncthreads();
    if (valid) {
      for (int k = nkept - 1 - part; k >= 0; k -= 4) {
        if (nms_iou_gt(me, area, kb[k], ka[k], thr)) { dead[ci] = 1; break; }
      }
    }
    __syncthreads();
    if (!dead[ci]) {
#pragma unroll 1
      for (int w = 2 * part; w < 2 * part + 2; ++w) {    // upper triangle only: columns j > ci
        unsigned word = 0;
        const int j0 = w * 32;
        if (j0 + 31 > ci) {
          for (int bit = 0; bit < 32; ++bit) {
            const int j = j0 + bit;
            if (j > ci && j < nc && nms_iou_gt(cb[j], ca[j], me, area, thr)) word |= 1u << bit;
          }
        }
        mask[ci][w] = word;
      }
    }
    __syncthreads();
    if (warp == 0) {
      unsigned aw = 0;
      if (lane < W) {
        for (int bit = 0; bit < 32; ++bit)
          if (!dead[lane * 32 + bit]) aw |= 1u << bit;
      }
      int nnew = 0;
      while (true) {
        const unsigned nz = __ballot_sync(0xffffffffu, aw != 0u);
        if (!nz) break;
        const int L = __ffs(nz) - 1;
        const unsigned word = __shfl_sync(0xffffffffu, aw, L);
        const int i = L * 32 + __ffs(word) - 1;
        if (lane == 0) newlist[nnew] = i;
        ++nnew;
        if (nkept + nnew >= max_out) break;
        if (lane == L) aw &= ~(1u << (i & 31));
        if (lane >= L && lane < W) aw &= ~mask[i][lane];
      }
      if (lane == 0) s_new = nnew;
    }
    __syncthreads();
    const int nnew = s_new;
    for (int k = t; k < nnew; k += NMS_THREADS) {
      const int i = newlist[k];
      kb[nkept + k] = cb[i]; ka[nkept + k] = ca[i];
      const long long o = (long long)b * max_out + nkept + k;
      out_boxes[o] = bx[cidx[i]];
      if (out_scores) out_scores[o] = scores[(long long)b * N + cidx[i]];
      if (out_idx) out_idx[o] = cidx[i];
    }
    nkept += nnew;
    __syncthreads();
  }
  // zero padding (post_processing.py:281-296)
  for (int k = nkept + t; k < max_out; k += NMS_THREADS) {
    const long long o = (long long)b * max_out + k;
    out_boxes[o] = make_float4(0, 0, 0, 0);
    if (out_scores) out_scores[o] = 0.0f;
    if (out_idx) out_idx[o] = -1;
  }
  if (t == 0) num_out[b] = nkept;
}

// ------------------------------------------------------------------ IoU + ArgMaxMatcher
// box_list_ops.iou: exactly 0 where the intersection is 0.
__device__ __forceinline__ float blo_iou(const float4 g, float area_g, const float4 a, float area_a) {
  const float ih = fmaxf(0.0f, fminf(g.z, a.z) - fmaxf(g.x, a.x));
  const float iw = fmaxf(0.0f, fminf(g.w, a.w) - fmaxf(g.y, a.y));
  const float inter = ih * iw;
  const float uni = area_g + area_a - inter;
  return inter == 0.0f ? 0.0f : inter / uni;
}

constexpr int MATCH_MAX_GT = 512;
__global__ void __launch_bounds__(256)
iou_match_kernel(const float4* __restrict__ gt, const int* __restrict__ num_gt, int Gmax,
                 const float4* __restrict__ boxes, long long box_batch_stride, const int* __restrict__ num_boxes,
                 int N, float matched_thr, float unmatched_thr, int use_thr, int force_match,
                 int* __restrict__ match, float* __restrict__ max_iou, u64* __restrict__ row_best) {
  __shared__ float4 sg[MATCH_MAX_GT];
  __shared__ float sa[MATCH_MAX_GT];
  const int b = blockIdx.y;
  const int G = min(num_gt[b], Gmax);
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float4 t = gt[(long long)b * Gmax + g];
    sg[g] = t;
    sa[g] = (t.z - t.x) * (t.w - t.y);
  }
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long o = (long long)b * N + n;
  if (num_boxes && n >= num_boxes[b]) { match[o] = -3; if (max_iou) max_iou[o] = 0.0f; return; }
  if (G == 0) { match[o] = -1; if (max_iou) max_iou[o] = 0.0f; return; }   // _match_when_rows_are_empty
  const float4 a = boxes[(long long)b * box_batch_stride + n];
  const float area_a = (a.z - a.x) * (a.w - a.y);
  float best = -1.0f;
  int arg = 0;
  for (int g = 0; g < G; ++g) {
    const float v = blo_iou(sg[g], sa[g], a, area_a);
    if (v > best) { best = v; arg = g; }           // first maximum wins (tf.argmax)
    if (force_match) {
      // per-row arg-max over columns: max IoU, ties -> lowest column
      const u64 key = ((u64)__float_as_uint(v) << 32) | (u64)(0xffffffffu - (unsigned)n);
      u64* slot = row_best + (long long)b * Gmax + g;
      if (key > *reinterpret_cast<volatile u64*>(slot)) atomicMax(slot, key);
    }
  }
  int m = arg;
  if (use_thr) {
    const bool below = unmatched_thr > best;
    const bool between = (best >= unmatched_thr) && (matched_thr > best);
    if (below) m = -1;
    if (between) m = -2;
  }
  match[o] = m;
  if (max_iou) max_iou[o] = best;
}

// force_match_for_each_row: later rows override earlier ones (dynamic_stitch order, am:158-169)
__global__ void force_match_kernel(const int* __restrict__ num_gt, int Gmax, int N,
                                   const u64* __restrict__ row_best, int* __restrict__ match) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  const int G = min(num_gt[b], Gmax);
  for (int g = 0; g < G; ++g) {
    const u64 k = row_best[(long long)b * Gmax + g];
    const unsigned col = 0xffffffffu - (unsigned)(k & 0xffffffffull);
    if (col < (unsigned)N) match[(long long)b * N + col] = g;
  }
}

// ------------------------------------------------------------------ balanced sampler on explicit keys
// code: 1 positive candidate (match >= 0), 0 negative candidate (match == -1), else excluded.
__device__ __forceinline__ int sample_code(int m) { return m >= 0 ? 1 : (m == -1 ? 0 : -1); }

// order-preserving float -> uint map (handles negative keys too)
__device__ __forceinline__ unsigned key_bits(float k) {
  const unsigned u = __float_as_uint(k);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// One block per image: exact selection of the `quota` smallest (key, index) pairs per class by a
// 2048-bin radix histogram on the key bits, a block scan to find the threshold bin, and an exact
// rank among the few candidates that fall INTO the threshold bin.  O(N) work instead of O(N^2).
constexpr int SB_BINS = 2048;
constexpr int SB_LIST = 2048;
__global__ void __launch_bounds__(1024)
sampler_kernel(const int* __restrict__ match, const float* __restrict__ keys, int N, int batch_size, int max_pos,
               int* __restrict__ counts, unsigned char* __restrict__ sampled) {
  __shared__ int hist[2][SB_BINS];
  __shared__ int list[2][SB_LIST];
  __shared__ int wsum[33];
  __shared__ int s_tot[2], s_thr[2], s_before[2], s_nlist[2];
  const int b = blockIdx.x;
  const int* mt = match + (long long)b * N;
  const float* ky = keys + (long long)b * N;
  unsigned char* out = sampled + (long long)b * N;
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * SB_BINS; i += blockDim.x) (&hist[0][0])[i] = 0;
  if (tid < 2) { s_tot[tid] = 0; s_thr[tid] = -1; s_before[tid] = 0; s_nlist[tid] = 0; }
  __syncthreads();
  for (int i = tid; i < N; i += blockDim.x) {
    const int c = sample_code(mt[i]);
    if (c >= 0) atomicAdd(&hist[c][key_bits(ky[i]) >> 21], 1);
  }
  __syncthreads();
  // class totals
  for (int c = 0; c < 2; ++c) {
    int v = 0;
    for (int i = tid; i < SB_BINS; i += blockDim.x) v += hist[c][i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0 && v) atomicAdd(&s_tot[c], v);
  }
  __syncthreads();
  const int npos = s_tot[1], nneg = s_tot[0];
  const int take_pos = min(npos, max_pos);
  const int take_neg = min(nneg, batch_size - take_pos);
  // threshold bin per class: exclusive prefix < quota <= inclusive prefix (2 bins per thread)
  for (int c = 0; c < 2; ++c) {
    const int quota = c == 1 ? take_pos : take_neg;
    const int h0 = hist[c][2 * tid], h1 = hist[c][2 * tid + 1];
    const int mine = h0 + h1;
    int incl = mine;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int v = wsum[lane];
      int wi = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      wsum[lane] = wi - v;
    }
    __syncthreads();
    const int excl = wsum[warp] + incl - mine;
    if (quota > 0) {
      if (excl < quota && quota <= excl + h0) { s_thr[c] = 2 * tid; s_before[c] = excl; }
      else if (excl + h0 < quota && quota <= excl + mine) { s_thr[c] = 2 * tid + 1; s_before[c] = excl + h0; }
    }
  }
  __syncthreads();
  // decide everything outside the threshold bins, collect the candidates inside them
  for (int i = tid; i < N; i += blockDim.x) {
    const int c = sample_code(mt[i]);
    unsigned char sel = 0;
    if (c >= 0) {
      const int bin = (int)(key_bits(ky[i]) >> 21);
      if (bin < s_thr[c]) sel = 1;
      else if (bin == s_thr[c]) {
        const int pos = atomicAdd(&s_nlist[c], 1);
        if (pos < SB_LIST) list[c][pos] = i;
      }
    }
    out[i] = sel;
  }
  __syncthreads();
  for (int c = 0; c < 2; ++c) {
    const int quota = (c == 1 ? take_pos : take_neg) - s_before[c];
    const int m = s_nlist[c];
    if (m <= SB_LIST) {
      for (int e = tid; e < m; e += blockDim.x) {
        const int i = list[c][e];
        const float ki = ky[i];
        int rank = 0;
        for (int f = 0; f < m; ++f) {
          const int j = list[c][f];
          const float kj = ky[j];
          rank += (kj < ki) || (kj == ki && j < i);
        }
        out[i] = rank < quota ? 1 : 0;
      }
    } else {
      // pathological tie mass (more than SB_LIST equal-bin keys): exact but O(m N)
      const int thr = s_thr[c];
      for (int i = tid; i < N; i += blockDim.x) {
        if (sample_code(mt[i]) != c || (int)(key_bits(ky[i]) >> 21) != thr) continue;
        const float ki = ky[i];
        int rank = 0;
        for (int j = 0; j < N; ++j) {
          if (sample_code(mt[j]) != c || (int)(key_bits(ky[j]) >> 21) != thr) continue;
          const float kj = ky[j];
          rank += (kj < ki) || (kj == ki && j < i);
        }
        out[i] = rank < quota ? 1 : 0;
      }
    }
  }
  if (tid == 0) {
    counts[b * 4 + 0] = npos;
    counts[b * 4 + 1] = nneg;
    counts[b * 4 + 2] = take_pos;
    counts[b * 4 + 3] = take_pos + take_neg;     // number of sampled entries (the RPN normaliser)
  }
}

// ------------------------------------------------------------------ ordered gather of sampled proposals
// boolean_mask (keeps score order) + pad_or_clip_box_list(P) + to_normalized_coordinates +
// normalized_to_image_coordinates (fmA:1134-1216, 1126-1131, 682-683).  One block per image.
__global__ void gather_sampled_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores,
                                      const unsigned char* __restrict__ sampled, int N, int P, float img_h,
                                      float img_w, float4* __restrict__ out_abs, float4* __restrict__ out_norm,
                                      float* __restrict__ out_scores, int* __restrict__ num_out) {
  __shared__ int wsum[33];
  const int b = blockIdx.x;
  const float ys = 1.0f / img_h, xs = 1.0f / img_w;
  int base = 0;
  for (int start = 0; start < N; start += blockDim.x) {
    const int i = start + threadIdx.x;
    const bool ok = i < N && sampled[(long long)b * N + i] != 0;
    int total;
    const int pos = base + block_excl_scan_flag(ok, &total, wsum);
    if (ok && pos < P) {
      const float4 r = boxes[(long long)b * N + i];
      const float4 nr = make_float4(ys * r.x, xs * r.y, ys * r.z, xs * r.w);
      out_norm[(long long)b * P + pos] = nr;
      out_abs[(long long)b * P + pos] = make_float4(img_h * nr.x, img_w * nr.y, img_h * nr.z, img_w * nr.w);
      if (out_scores) out_scores[(long long)b * P + pos] = scores[(long long)b * N + i];
    }
    base += total;
  }
  const int n = min(base, P);
  for (int k = n + threadIdx.x; k < P; k += blockDim.x) {
    out_norm[(long long)b * P + k] = make_float4(0, 0, 0, 0);
    out_abs[(long long)b * P + k] = make_float4(0, 0, 0, 0);
    if (out_scores) out_scores[(long long)b * P + k] = 0.0f;
  }
  if (threadIdx.x == 0) num_out[b] = n;
}

// ------------------------------------------------------------------ target creation
// FasterRcnnBoxCoder._encode (coder:60-90), EPS = 1e-8, scale factors 10,10,5,5.
__device__ __forceinline__ float4 box_encode(const float4 g, const float4 a) {
  float wa = a.w - a.y, ha = a.z - a.x;
  const float yca = a.x + ha / 2.0f, xca = a.y + wa / 2.0f;
  float w = g.w - g.y, h = g.z - g.x;
  const float yc = g.x + h / 2.0f, xc = g.y + w / 2.0f;
  ha += 1e-8f; wa += 1e-8f; h += 1e-8f; w += 1e-8f;
  float4 t;
  t.x = (yc - yca) / ha * 10.0f;
  t.y = (xc - xca) / wa * 10.0f;
  t.z = logf(h / ha) * 5.0f;
  t.w = logf(w / wa) * 5.0f;
  return t;
}

// detection-stage targets for one padded proposal batch; one block per image (P <= 1024 looped).
// Also emits the closeness row weight  reg_w / max(1, sum_p reg_w) * sum_{k>=1} closeness_target
// (fmA:1771-1789) so that the loss kernel needs no second pass.
__global__ void detection_targets_kernel(const int* __restrict__ match, const float4* __restrict__ props,
                                         const float4* __restrict__ gt, const int* __restrict__ gt_cls,
                                         const float* __restrict__ gt_close, int Gmax, int P, int K1,
                                         int* __restrict__ cls_t, float4* __restrict__ reg_t,
                                         float* __restrict__ reg_w, float* __restrict__ cls_w,
                                         float* __restrict__ close_t, float* __restrict__ close_w) {
  __shared__ float red[32];
  __shared__ float s_total;
  const int b = blockIdx.x;
  float local = 0.0f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) local += (match[(long long)b * P + p] >= 0) ? 1.0f : 0.0f;
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    v = warp_sum(v);
    if (threadIdx.x == 0) s_total = v;
  }
  __syncthreads();
  const float norm_reg = fmaxf(1.0f, s_total);
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const long long o = (long long)b * P + p;
    const int m = match[o];
    const bool matched = m >= 0;
    float4 rt = make_float4(0, 0, 0, 0);
    int c = 0;
    if (matched) {
      rt = box_encode(gt[(long long)b * Gmax + m], props[o]);
      c = gt_cls[(long long)b * Gmax + m];
    }
    cls_t[o] = c;
    reg_t[o] = rt;
    reg_w[o] = matched ? 1.0f : 0.0f;
    cls_w[o] = (matched || m == -1) ? 1.0f : 0.0f;
    if (close_t) {
      float s = 0.0f;
      for (int k = 0; k < K1; ++k) {
        const float v = matched ? gt_close[((long long)b * Gmax + m) * K1 + k] : 0.0f;
        close_t[o * K1 + k] = v;
        if (k >= 1) s += v;
      }
      close_w[o] = (matched ? 1.0f : 0.0f) / norm_reg * s;
    }
  }
}

// RPN targets (dense, for inspection / parity): cls target 0/1, weights, encoded regression target.
__global__ void rpn_targets_kernel(const int* __restrict__ match, const float4* __restrict__ anchors,
                                   const float4* __restrict__ gt, int Gmax, int N, float* __restrict__ cls_t,
                                   float* __restrict__ cls_w, float4* __restrict__ reg_t,
                                   float* __restrict__ reg_w) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long o = (long long)b * N + n;
  const int m = match[o];
  const bool matched = m >= 0;
  cls_t[o] = matched ? 1.0f : 0.0f;
  cls_w[o] = (matched || m == -1) ? 1.0f : 0.0f;
  reg_w[o] = matched ? 1.0f : 0.0f;
  reg_t[o] = matched ? box_encode(gt[(long long)b * Gmax + m], anchors[n]) : make_float4(0, 0, 0, 0);
}

// proposals expanded linearly toward the full image for the refine head (fmA:783-803):
// out[e, b, p] = (ymin - e*ymin/4, xmin - e*xmin/4, ymax + e*(1-ymax)/4, xmax + e*(1-xmax)/4)
__global__ void expand_windows_kernel(const float4* __restrict__ props, int BP, int P, int n_expand,
                                      float4* __restrict__ out, int* __restrict__ box_ind) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BP) return;
  const float4 r = props[i];
  const float div = (float)n_expand;
  const float dyn = r.x / div, dxn = r.y / div, dyp = (1.0f - r.z) / div, dxp = (1.0f - r.w) / div;
  for (int e = 0; e <= n_expand; ++e) {
    const float fe = (float)e;
    out[(long long)e * BP + i] = make_float4(r.x - dyn * fe, r.y - dxn * fe, r.z + dyp * fe, r.w + dxp * fe);
    if (box_ind) box_ind[(long long)e * BP + i] = i / P;
  }
}

}  // namespace

// ===================================================================================== C ABI
extern "C" int mtl_grid_anchors(int Hf, int Wf, const float* scales, int ns, const float* ars, int na,
                                float base_h, float base_w, float stride_h, float stride_w, float off_h,
                                float off_w, float* anchors, cudaStream_t stream) {
  MTL_CHECK_ARG(Hf > 0 && Wf > 0 && anchors, "mtl_grid_anchors: bad geometry");
  MTL_CHECK_ARG(ns > 0 && ns <= 16 && na > 0 && na <= 16, "mtl_grid_anchors: 1..16 scales/aspect ratios");
  AnchorSpec s;
  memset(&s, 0, sizeof(s));
  for (int i = 0; i < ns; ++i) s.scales[i] = scales[i];
  for (int i = 0; i < na; ++i) s.ars[i] = ars[i];
  s.ns = ns; s.na = na; s.base_h = base_h; s.base_w = base_w; s.stride_h = stride_h; s.stride_w = stride_w;
  s.off_h = off_h; s.off_w = off_w;
  const long long total = (long long)Hf * Wf * ns * na;
  const int grid = (int)min((long long)mtl_num_sms() * 8, ceil_div_ll(total, 256));
  grid_anchors_kernel<<<grid, 256, 0, stream>>>(Hf, Wf, s, anchors);
  MTL_CUDA_LAUNCH_CHECK("grid_anchors_kernel");
  return MTL_OK;
}

extern "C" int mtl_prune_outside_window(const float* boxes, int N, float wy0, float wx0, float wy1, float wx1,
                                        int* keep_idx, float* kept_boxes, int* num_keep, cudaStream_t stream) {
  MTL_CHECK_ARG(N >= 0 && keep_idx && kept_boxes && num_keep, "mtl_prune_outside_window: null output");
  prune_outside_window_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const float4*>(boxes), N, wy0, wx0, wy1,
                                                      wx1, keep_idx, reinterpret_cast<float4*>(kept_boxes),
                                                      num_keep);
  MTL_CUDA_LAUNCH_CHECK("prune_outside_window_kernel");
  return MTL_OK;
}

extern "C" int mtl_rpn_decode(const float* rpn_out, long long ld, int box_col0, int cls_col0, int A, int HW,
                              const int* keep_idx, const float* anchors, int Nk, int B, float img_h,
                              float img_w, float score_thresh, float* boxes, float* scores,
                              unsigned long long* keys, cudaStream_t stream) {
  MTL_CHECK_ARG(rpn_out && anchors && boxes && scores && keys, "mtl_rpn_decode: null tensor");
  MTL_CHECK_ARG(Nk > 0 && B > 0 && A > 0, "mtl_rpn_decode: empty problem");
  dim3 grid(ceil_div(Nk, 256), B);
  rpn_decode_kernel<<<grid, 256, 0, stream>>>(rpn_out, ld, box_col0, cls_col0, A, HW, keep_idx,
                                              reinterpret_cast<const float4*>(anchors), Nk, img_h, img_w,
                                              score_thresh, reinterpret_cast<float4*>(boxes), scores, keys);
  MTL_CUDA_LAUNCH_CHECK("rpn_decode_kernel");
  return MTL_OK;
}

extern "C" int mtl_nms_make_keys(const float* boxes, const float* scores, int B, int N, float score_thresh,
                                 int require_area, unsigned long long* keys, cudaStream_t stream) {
  MTL_CHECK_ARG(boxes && scores && keys && B > 0 && N > 0, "mtl_nms_make_keys: bad args");
  dim3 grid(ceil_div(N, 256), B);
  make_keys_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(boxes), scores, N, score_thresh,
                                             require_area, keys);
  MTL_CUDA_LAUNCH_CHECK("make_keys_kernel");
  return MTL_OK;
}

extern "C" int mtl_rank_sort_desc(const unsigned long long* keys, int B, int N, int* order, int* num_valid,
                                  int* rank_ws, cudaStream_t stream) {
  MTL_CHECK_ARG(keys && order && num_valid && rank_ws && B > 0 && N > 0, "mtl_rank_sort_desc: bad args");
  cudaError_t e = cudaMemsetAsync(num_valid, 0, sizeof(int) * B, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(rank_ws, 0, sizeof(int) * (size_t)B * N, stream);
  if (e != cudaSuccess) { mtl_set_error("mtl_rank_sort_desc: memset: %s", cudaGetErrorString(e)); return MTL_ERR_CUDA; }
  dim3 grid(ceil_div(N, RS_THREADS), B, RS_SPLIT);
  rank_count_kernel<<<grid, RS_THREADS, 0, stream>>>(keys, N, rank_ws);
  MTL_CUDA_LAUNCH_CHECK("rank_count_kernel");
  dim3 g2(ceil_div(N, 256), B);
  rank_scatter_kernel<<<g2, 256, 0, stream>>>(keys, rank_ws, N, order, num_valid);
  MTL_CUDA_LAUNCH_CHECK("rank_scatter_kernel");
  return MTL_OK;
}

extern "C" int mtl_nms(const float* boxes, const float* scores, const int* order, const int* num_valid, int B,
                       int N, float iou_thresh, int max_out, float* out_boxes, float* out_scores, int* out_idx,
                       int* num_out, cudaStream_t stream) {
  MTL_CHECK_ARG(boxes && order && num_valid && out_boxes && num_out, "mtl_nms: null tensor");
  MTL_CHECK_ARG(max_out > 0 && max_out <= NMS_MAX_OUT, "mtl_nms: max_out must be in 1..%d", NMS_MAX_OUT);
  nms_kernel<<<B, NMS_THREADS, 0, stream>>>(reinterpret_cast<const float4*>(boxes), scores, order, num_valid, N,
                                            iou_thresh, max_out, reinterpret_cast<float4*>(out_boxes), out_scores,
                                            out_idx, num_out);
  MTL_CUDA_LAUNCH_CHECK("nms_kernel");
  return MTL_OK;
}

extern "C" int mtl_clip_boxes(const float* boxes, int N, float wy0, float wx0, float wy1, float wx1, float* out,
                              cudaStream_t stream) {
  MTL_CHECK_ARG(boxes && out && N >= 0, "mtl_clip_boxes: bad args");
  if (N == 0) return MTL_OK;
  clip_boxes_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(boxes), N, wy0, wx0, wy1,
                                                         wx1, reinterpret_cast<float4*>(out));
  MTL_CUDA_LAUNCH_CHECK("clip_boxes_kernel");
  return MTL_OK;
}

extern "C" int mtl_proposals_from_nms(const float* nms_boxes, const float* nms_scores, const int* nms_num, int B,
                                      int M, float img_h, float img_w, float* out_abs, float* out_norm,
                                      float* out_scores, int* num_out, cudaStream_t stream) {
  MTL_CHECK_ARG(nms_boxes && nms_scores && nms_num && out_abs && out_norm && num_out && B > 0 && M > 0,
                "mtl_proposals_from_nms: bad args");
  dim3 grid(ceil_div(M, 128), B);
  proposals_from_nms_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const float4*>(nms_boxes), nms_scores, nms_num,
                                                      M, img_h, img_w, reinterpret_cast<float4*>(out_abs),
                                                      reinterpret_cast<float4*>(out_norm), out_scores, num_out);
  MTL_CUDA_LAUNCH_CHECK("proposals_from_nms_kernel");
  return MTL_OK;
}

extern "C" int mtl_detection_decode(const float* box_encodings, const float* class_logits, const float* proposals,
                                    const int* num_proposals, int B, int P, int K, float img_h, float img_w,
                                    float score_thresh, int score_mode, float* boxes_norm, float* scores,
                                    unsigned long long* keys, float* decoded_abs, cudaStream_t stream) {
  MTL_CHECK_ARG(box_encodings && class_logits && proposals && num_proposals && boxes_norm && scores && keys,
                "mtl_detection_decode: null tensor");
  MTL_CHECK_ARG(B > 0 && P > 0 && K > 0 && K < 65536 && B < 65536, "mtl_detection_decode: bad shape");
  MTL_CHECK_ARG(score_mode >= 0 && score_mode <= 2, "mtl_detection_decode: score_mode 0 identity, 1 softmax, 2 sigmoid");
  dim3 grid(ceil_div(P, 128), K, B);
  detection_decode_kernel<<<grid, 128, 0, stream>>>(box_encodings, class_logits,
                                                    reinterpret_cast<const float4*>(proposals), num_proposals, P, K,
                                                    img_h, img_w, score_thresh, score_mode,
                                                    reinterpret_cast<float4*>(boxes_norm), scores, keys,
                                                    reinterpret_cast<float4*>(decoded_abs));
  MTL_CUDA_LAUNCH_CHECK("detection_decode_kernel");
  return MTL_OK;
}

extern "C" int mtl_detection_merge_keys(const float* cls_scores, const int* cls_num, int B, int K, int M,
                                        unsigned long long* keys, cudaStream_t stream) {
  MTL_CHECK_ARG(cls_scores && cls_num && keys && B > 0 && K > 0 && M > 0, "mtl_detection_merge_keys: bad args");
  dim3 grid(ceil_div(K * M, 256), B);
  detection_merge_keys_kernel<<<grid, 256, 0, stream>>>(cls_scores, cls_num, K, M, keys);
  MTL_CUDA_LAUNCH_CHECK("detection_merge_keys_kernel");
  return MTL_OK;
}

extern "C" int mtl_detection_gather(const float* cls_boxes, const float* cls_scores, const int* order,
                                    const int* num_valid, int B, int K, int M, int T, float* det_boxes,
                                    float* det_scores, float* det_classes, float* num_detections,
                                    cudaStream_t stream) {
  MTL_CHECK_ARG(cls_boxes && cls_scores && order && num_valid && det_boxes && det_scores && det_classes &&
                num_detections && B > 0 && K > 0 && M > 0 && T > 0, "mtl_detection_gather: bad args");
  dim3 grid(ceil_div(T, 128), B);
  detection_gather_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const float4*>(cls_boxes), cls_scores, order,
                                                    num_valid, K, M, T, reinterpret_cast<float4*>(det_boxes),
                                                    det_scores, det_classes, num_detections);
  MTL_CUDA_LAUNCH_CHECK("detection_gather_kernel");
  return MTL_OK;
}

extern "C" int mtl_iou_match(const float* gt, const int* num_gt, int Gmax, const float* boxes,
                             long long box_batch_stride, const int* num_boxes, int B, int N, float matched_thr,
                             float unmatched_thr, int use_thresholds, int force_match, int* match,
                             float* max_iou, unsigned long long* row_best, cudaStream_t stream) {
  MTL_CHECK_ARG(gt && num_gt && boxes && match, "mtl_iou_match: null tensor");
  MTL_CHECK_ARG(Gmax > 0 && Gmax <= MATCH_MAX_GT, "mtl_iou_match: Gmax must be in 1..%d", MATCH_MAX_GT);
  MTL_CHECK_ARG(!force_match || row_best, "mtl_iou_match: force_match needs the row_best workspace");
  if (force_match) {
    cudaError_t e = cudaMemsetAsync(row_best, 0, sizeof(u64) * (size_t)B * Gmax, stream);
    if (e != cudaSuccess) { mtl_set_error("mtl_iou_match: memset: %s", cudaGetErrorString(e)); return MTL_ERR_CUDA; }
  }
  dim3 grid(ceil_div(N, 256), B);
  iou_match_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(gt), num_gt, Gmax,
                                             reinterpret_cast<const float4*>(boxes), box_batch_stride, num_boxes,
                                             N, matched_thr, unmatched_thr, use_thresholds, force_match, match,
                                             max_iou, row_best);
  MTL_CUDA_LAUNCH_CHECK("iou_match_kernel");
  if (force_match) {
    force_match_kernel<<<B, 32, 0, stream>>>(num_gt, Gmax, N, row_best, match);
    MTL_CUDA_LAUNCH_CHECK("force_match_kernel");
  }
  return MTL_OK;
}

extern "C" int mtl_balanced_sample(const int* match, const float* keys, int B, int N, int batch_size,
                                   float positive_fraction, unsigned char* sampled, int* counts,
                                   cudaStream_t stream) {
  MTL_CHECK_ARG(match && keys && sampled && counts && B > 0 && N > 0, "mtl_balanced_sample: bad args");
  const int max_pos = (int)(positive_fraction * (float)batch_size);   // int(frac * batch) (bpns:72)
  sampler_kernel<<<B, 1024, 0, stream>>>(match, keys, N, batch_size, max_pos, counts, sampled);
  MTL_CUDA_LAUNCH_CHECK("sampler_kernel");
  return MTL_OK;
}

extern "C" int mtl_gather_sampled(const float* boxes, const float* scores, const unsigned char* sampled, int B,
                                  int N, int P, float img_h, float img_w, float* out_abs, float* out_norm,
                                  float* out_scores, int* num_out, cudaStream_t stream) {
  MTL_CHECK_ARG(boxes && sampled && out_abs && out_norm && num_out, "mtl_gather_sampled: null tensor");
  gather_sampled_kernel<<<B, 1024, 0, stream>>>(reinterpret_cast<const float4*>(boxes), scores, sampled, N, P,
                                                img_h, img_w, reinterpret_cast<float4*>(out_abs),
                                                reinterpret_cast<float4*>(out_norm), out_scores, num_out);
  MTL_CUDA_LAUNCH_CHECK("gather_sampled_kernel");
  return MTL_OK;
}

extern "C" int mtl_detection_targets(const int* match, const float* proposals, const float* gt,
                                     const int* gt_classes, const float* gt_closeness, int B, int Gmax, int P,
                                     int K1, int* cls_targets, float* reg_targets, float* reg_weights,
                                     float* cls_weights, float* closeness_targets, float* closeness_weights,
                                     cudaStream_t stream) {
  MTL_CHECK_ARG(match && proposals && gt && gt_classes && cls_targets && reg_targets && reg_weights && cls_weights,
                "mtl_detection_targets: null tensor");
  MTL_CHECK_ARG((closeness_targets == nullptr) == (gt_closeness == nullptr) &&
                (closeness_targets == nullptr) == (closeness_weights == nullptr),
                "mtl_detection_targets: closeness inputs/outputs must come together");
  detection_targets_kernel<<<B, 256, 0, stream>>>(match, reinterpret_cast<const float4*>(proposals),
                                                  reinterpret_cast<const float4*>(gt), gt_classes, gt_closeness,
                                                  Gmax, P, K1, cls_targets, reinterpret_cast<float4*>(reg_targets),
                                                  reg_weights, cls_weights, closeness_targets, closeness_weights);
  MTL_CUDA_LAUNCH_CHECK("detection_targets_kernel");
  return MTL_OK;
}

extern "C" int mtl_rpn_targets(const int* match, const float* anchors, const float* gt, int B, int Gmax, int N,
                               float* cls_targets, float* cls_weights, float* reg_targets, float* reg_weights,
                               cudaStream_t stream) {
  MTL_CHECK_ARG(match && anchors && gt && cls_targets && cls_weights && reg_targets && reg_weights,
                "mtl_rpn_targets: null tensor");
  dim3 grid(ceil_div(N, 256), B);
  rpn_targets_kernel<<<grid, 256, 0, stream>>>(match, reinterpret_cast<const float4*>(anchors),
                                               reinterpret_cast<const float4*>(gt), Gmax, N, cls_targets,
                                               cls_weights, reinterpret_cast<float4*>(reg_targets), reg_weights);
  MTL_CUDA_LAUNCH_CHECK("rpn_targets_kernel");
  return MTL_OK;
}

extern "C" int mtl_expand_windows(const float* proposals_norm, int B, int P, int n_expand, float* out,
                                  int* box_ind, cudaStream_t stream) {
  MTL_CHECK_ARG(proposals_norm && out && n_expand > 0, "mtl_expand_windows: bad args");
  const int BP = B * P;
  expand_windows_kernel<<<ceil_div(BP, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(proposals_norm), BP,
                                                               P, n_expand, reinterpret_cast<float4*>(out), box_ind);
  MTL_CUDA_LAUNCH_CHECK("expand_windows_kernel");
  return MTL_OK;
}
