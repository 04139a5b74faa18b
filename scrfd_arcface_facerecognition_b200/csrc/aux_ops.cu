// CUDA-core kernels around the tensor-core path (all HBM-bound):
//   stem conv (cin <= 4), depthwise conv, pooling       -- parts of the ONNX graphs behind
//                                                           reference models/scrfd.py:83, models/arcface.py:51
//   L2 normalisation, pairwise cosine, top-k merge       -- reference utils/helpers.py:110-123, main.py:136-142,
//                                                           qdrant_manager.py:138-188, duplicate.py:1491-1496
//   greedy duplicate-merge resolve                        -- reference duplicate.py:2726-2797
#include "b2f_common.cuh"
#include "../../include/b2f.h"

#include <atomic>

namespace b2f {
extern std::atomic<long long> g_launches;

__device__ __forceinline__ float h2f(uint16_t v, int is_bf16) {
  if (is_bf16) return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&v));
  return __half2float(*reinterpret_cast<__half*>(&v));
}
__device__ __forceinline__ uint16_t f2h(float v, int is_bf16) {
  if (is_bf16) {
    __nv_bfloat16 t = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&t);
  }
  __half t = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&t);
}
__device__ __forceinline__ float act_f(float v, int act, float slope) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v >= 0.f ? v : v * slope;
  if (act == 3) return 1.f / (1.f + expf(-v));
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& q, int is_bf16, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = h2f((uint16_t)(w[i] & 0xFFFF), is_bf16);
    f[2 * i + 1] = h2f((uint16_t)(w[i] >> 16), is_bf16);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int is_bf16) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = (uint32_t)f2h(f[2 * i], is_bf16) | ((uint32_t)f2h(f[2 * i + 1], is_bf16) << 16);
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------------------------------
// stem: 3x3 pad-1 conv over a 4-channel (RGB0) NHWC input, 16 output channels per thread
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stem_conv_kernel(const uint16_t* __restrict__ in, int n, int h, int w, int stride, int ho, int wo,
                 const float* __restrict__ weight /*[9][4][cout_p]*/, const float* __restrict__ bias,
                 const float* __restrict__ slope, int act, int cout_p, int is_bf16, uint16_t* __restrict__ out) {
  // each thread: 16 output channels x 2 horizontally adjacent output pixels, weights read as float4 from smem
  extern __shared__ __align__(16) float s_w[];  // [9*4][cout_p] + bias[cout_p] + slope[cout_p]
  const int wsize = 36 * cout_p;
  for (int i = threadIdx.x; i < wsize; i += blockDim.x) s_w[i] = weight[i];
  for (int i = threadIdx.x; i < cout_p; i += blockDim.x) {
    s_w[wsize + i] = bias[i];
    s_w[wsize + cout_p + i] = slope ? slope[i] : 0.f;
  }
  __syncthreads();
  const int groups = cout_p >> 4;
  const int pairs = (wo + 1) >> 1;
  const long long total = (long long)n * ho * pairs * groups;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(t % groups);
    const long long pr = t / groups;
    const int px = (int)(pr % pairs), oy = (int)((pr / pairs) % ho), b = (int)(pr / ((long long)pairs * ho));
    const int ox = px * 2;
    float acc0[16], acc1[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc0[i] = acc1[i] = s_w[wsize + cg * 16 + i];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy * stride + r - 1;
      if (iy < 0 || iy >= h) continue;
      const uint16_t* row = in + ((size_t)b * h + iy) * w * 4;
      float v[5][3];                       // up to stride + 3 input columns
      const int ix0 = ox * stride - 1;
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        const int ix = ix0 + c;
        uint2 q = make_uint2(0u, 0u);
        if (c < stride + 3 && ix >= 0 && ix < w) q = *reinterpret_cast<const uint2*>(row + (size_t)ix * 4);
        v[c][0] = h2f((uint16_t)(q.x & 0xFFFF), is_bf16);
        v[c][1] = h2f((uint16_t)(q.x >> 16), is_bf16);
        v[c][2] = h2f((uint16_t)(q.y & 0xFFFF), is_bf16);
      }
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const float* wp = s_w + ((r * 3 + s) * 4) * cout_p + cg * 16;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float a0 = v[s][ci];
          const float a1 = stride == 1 ? v[s + 1][ci] : v[s + 2][ci];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wp + ci * cout_p + q4 * 4);
            acc0[q4 * 4 + 0] = fmaf(a0, w4.x, acc0[q4 * 4 + 0]);
            acc0[q4 * 4 + 1] = fmaf(a0, w4.y, acc0[q4 * 4 + 1]);
            acc0[q4 * 4 + 2] = fmaf(a0, w4.z, acc0[q4 * 4 + 2]);
            acc0[q4 * 4 + 3] = fmaf(a0, w4.w, acc0[q4 * 4 + 3]);
            acc1[q4 * 4 + 0] = fmaf(a1, w4.x, acc1[q4 * 4 + 0]);
            acc1[q4 * 4 + 1] = fmaf(a1, w4.y, acc1[q4 * 4 + 1]);
            acc1[q4 * 4 + 2] = fmaf(a1, w4.z, acc1[q4 * 4 + 2]);
            acc1[q4 * 4 + 3] = fmaf(a1, w4.w, acc1[q4 * 4 + 3]);
          }
        }
      }
    }
    const float* sl = s_w + wsize + cout_p + cg * 16;
    float o8[8];
    uint16_t* op = out + (((size_t)b * ho + oy) * wo + ox) * cout_p + cg * 16;
#pragma unroll
    for (int half_i = 0; half_i < 2; ++half_i) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o8[i] = act_f(acc0[half_i * 8 + i], act, sl[half_i * 8 + i]);
      *reinterpret_cast<uint4*>(op + half_i * 8) = pack8(o8, is_bf16);
    }
    if (ox + 1 < wo) {
#pragma unroll
      for (int half_i = 0; half_i < 2; ++half_i) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o8[i] = act_f(acc1[half_i * 8 + i], act, sl[half_i * 8 + i]);
        *reinterpret_cast<uint4*>(op + cout_p + half_i * 8) = pack8(o8, is_bf16);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------
// 3x3 pad-1 patch extraction for the first layer: [n][h][w][4] -> [n][ho][wo][32], k = tap*3 + channel
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const uint16_t* __restrict__ in, int n, int h, int w, int stride, int ho, int wo,
                 uint16_t* __restrict__ out) {
  const long long total = (long long)n * ho * wo;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(t % wo), oy = (int)((t / wo) % ho), b = (int)(t / ((long long)wo * ho));
    uint16_t v[32];
#pragma unroll
    for (int i = 27; i < 32; ++i) v[i] = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy * stride + r - 1;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = ox * stride + s - 1;
        uint2 q = make_uint2(0u, 0u);
        if (iy >= 0 && iy < h && ix >= 0 && ix < w)
          q = __ldg(reinterpret_cast<const uint2*>(in + (((size_t)b * h + iy) * w + ix) * 4));
        const int k = (r * 3 + s) * 3;
        v[k] = (uint16_t)(q.x & 0xFFFF), v[k + 1] = (uint16_t)(q.x >> 16), v[k + 2] = (uint16_t)(q.y & 0xFFFF);
      }
    }
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)t * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = make_uint4(v[8 * j] | ((uint32_t)v[8 * j + 1] << 16), v[8 * j + 2] | ((uint32_t)v[8 * j + 3] << 16),
                        v[8 * j + 4] | ((uint32_t)v[8 * j + 5] << 16), v[8 * j + 6] | ((uint32_t)v[8 * j + 7] << 16));
  }
}

// ------------------------------------------------------------------------------------------
// depthwise k x k conv, 8 channels per thread
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dwconv_kernel(const uint16_t* __restrict__ in, int n, int h, int w, int c_p, int k, int stride, int pad, int ho, int wo,
              const float* __restrict__ weight /*[k*k][c_p]*/, const float* __restrict__ bias,
              const float* __restrict__ slope, int act, int is_bf16, uint16_t* __restrict__ out) {
  const int groups = c_p >> 3;
  const long long total = (long long)n * ho * wo * groups;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(t % groups);
    const long long pix = t / groups;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), b = (int)(pix / ((long long)wo * ho));
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = __ldg(bias + cg * 8 + i);
    for (int r = 0; r < k; ++r) {
      const int iy = oy * stride + r - pad;
      if (iy < 0 || iy >= h) continue;
      for (int s = 0; s < k; ++s) {
        const int ix = ox * stride + s - pad;
        if (ix < 0 || ix >= w) continue;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * h + iy) * w + ix) * c_p + cg * 8));
        float v[8];
        unpack8(q, is_bf16, v);
        const float* wp = weight + (size_t)(r * k + s) * c_p + cg * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(v[i], __ldg(wp + i), acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = act_f(acc[i], act, (act == 2) ? __ldg(slope + cg * 8 + i) : 0.f);
    *reinterpret_cast<uint4*>(out + (size_t)pix * c_p + cg * 8) = pack8(acc, is_bf16);
  }
}

// ------------------------------------------------------------------------------------------
// pooling: max (ONNX MaxPool with pads) or average (ceil_mode, count_include_pad = 0)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pool_kernel(const uint16_t* __restrict__ in, int n, int h, int w, int c_p, int k, int stride, int pad, int mode, int ho,
            int wo, int is_bf16, uint16_t* __restrict__ out) {
  const int groups = c_p >> 3;
  const long long total = (long long)n * ho * wo * groups;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(t % groups);
    const long long pix = t / groups;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), b = (int)(pix / ((long long)wo * ho));
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = mode == 0 ? -INFINITY : 0.f;
    int cnt = 0;
    for (int r = 0; r < k; ++r) {
      const int iy = oy * stride + r - pad;
      if (iy < 0 || iy >= h) continue;
      for (int s = 0; s < k; ++s) {
        const int ix = ox * stride + s - pad;
        if (ix < 0 || ix >= w) continue;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * h + iy) * w + ix) * c_p + cg * 8));
        float v[8];
        unpack8(q, is_bf16, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = mode == 0 ? fmaxf(acc[i], v[i]) : acc[i] + v[i];
        ++cnt;
      }
    }
    if (mode == 1) {
      const float inv = cnt > 0 ? 1.f / (float)cnt : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] *= inv;
    }
    *reinterpret_cast<uint4*>(out + (size_t)pix * c_p + cg * 8) = pack8(acc, is_bf16);
  }
}



// max-pool k x k on packed 16-bit pairs (max is exact in any format), one thread per (output pixel, 8 channels);
// grid (pieces of one output row, output rows, batch): 32-bit index arithmetic only
constexpr int kPoolRows = 8;
template <bool BF16>
__global__ void __launch_bounds__(256)
maxpool_rows_kernel(const uint16_t* __restrict__ in, int h, int w, int c_p, int k, int stride, int pad, int ho, int wo,
                    uint16_t* __restrict__ out) {
  const int groups = c_p >> 3;
  const int b = blockIdx.z;
  // a CTA walks kPoolRows consecutive output rows: the input rows two neighbouring output rows share are re-read
  // from L1 / L2 by the same CTA instead of from DRAM by another one
  for (int oy = blockIdx.y * kPoolRows; oy < ho && oy < (blockIdx.y + 1) * kPoolRows; ++oy)
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < wo * groups; t += gridDim.x * blockDim.x) {
    const int ox = t / groups, cg = t - ox * groups;
    uint4 acc;
    bool first = true;
    for (int r = 0; r < k; ++r) {
      const int iy = oy * stride + r - pad;
      if (iy < 0 || iy >= h) continue;
      const uint16_t* row = in + ((size_t)(b * h + iy) * w) * c_p + cg * 8;
      for (int s = 0; s < k; ++s) {
        const int ix = ox * stride + s - pad;
        if (ix < 0 || ix >= w) continue;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(row + (size_t)ix * c_p));
        if (first) {
          acc = q;
          first = false;
        } else if (BF16) {
          __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&acc);
          const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = __hmax2(a[i], v[i]);
        } else {
          __half2* a = reinterpret_cast<__half2*>(&acc);
          const __half2* v = reinterpret_cast<const __half2*>(&q);
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = __hmax2(a[i], v[i]);
        }
      }
    }
    if (first) acc = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(out + ((size_t)(b * ho + oy) * wo + ox) * c_p + cg * 8) = acc;
  }
}

// ------------------------------------------------------------------------------------------
// generic per-channel affine + add + activation (fallback for graph nodes no conv absorbed)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
eltwise_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ b, long long pixels, int c_p,
               const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ slope,
               int act, int is_bf16, uint16_t* __restrict__ out) {
  const int groups = c_p >> 3;
  const long long total = pixels * groups;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(t % groups);
    const size_t off = (size_t)(t / groups) * c_p + cg * 8;
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(a + off)), is_bf16, v);
    if (scale) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], __ldg(scale + cg * 8 + i), __ldg(shift + cg * 8 + i));
    }
    if (b) {
      float w[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(b + off)), is_bf16, w);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += w[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = act_f(v[i], act, (act == 2) ? __ldg(slope + cg * 8 + i) : 0.f);
    *reinterpret_cast<uint4*>(out + off) = pack8(v, is_bf16);
  }
}

// ------------------------------------------------------------------------------------------
// L2 normalisation / cosine pieces: one warp per row
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
l2norm_kernel(const float* __restrict__ x, long long rows, int dim, float* __restrict__ out_f32,
              uint16_t* __restrict__ out_16, int is_bf16, float* __restrict__ norms) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* xr = x + r * dim;
    float ss = 0.f;
    for (int i = lane; i < dim; i += 32) {
      const float v = xr[i];
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float inv = nrm > 0.f ? 1.f / nrm : 0.f;
    for (int i = lane; i < dim; i += 32) {
      const float v = xr[i] * inv;
      if (out_f32) out_f32[r * dim + i] = v;
      if (out_16) out_16[r * dim + i] = f2h(v, is_bf16);
    }
    if (norms && lane == 0) norms[r] = nrm;
  }
}

__global__ void __launch_bounds__(256)
cosine_pairs_kernel(const float* __restrict__ a, const float* __restrict__ b, int pairs, int dim,
                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= pairs) return;
  const float* pa = a + (size_t)warp * dim;
  const float* pb = b + (size_t)warp * dim;
  float ab = 0.f, aa = 0.f, bb = 0.f;
  for (int i = lane; i < dim; i += 32) {
    const float u = pa[i], v = pb[i];
    ab = fmaf(u, v, ab), aa = fmaf(u, u, aa), bb = fmaf(v, v, bb);
  }
  ab = warp_sum(ab), aa = warp_sum(aa), bb = warp_sum(bb);
  if (lane == 0) out[warp] = ab / (sqrtf(aa) * sqrtf(bb));
}

// ------------------------------------------------------------------------------------------
// top-k merge: pick the best R coarse candidates, re-score them exactly in fp32, sort, threshold
// ------------------------------------------------------------------------------------------
constexpr int kRescore = 8;

__global__ void __launch_bounds__(128)
match_merge_kernel(const float* __restrict__ part_score, const int* __restrict__ part_idx, int q, int n_cand,
                   const float* __restrict__ q_f32, const float* __restrict__ g_f32, int dim, int topk, float threshold,
                   int strict_gt, long long idx_base, float* __restrict__ out_score, long long* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= q) return;
  const float* ps = part_score + (size_t)row * n_cand;
  const int* pi = part_idx + (size_t)row * n_cand;
  float sel_s[kRescore];
  int sel_i[kRescore];
  // R rounds of warp arg-max over (score desc, index asc); picked entries are excluded by (score,idx) order
  float last_s = INFINITY;
  int last_i = -1;
  int nsel = 0;
  for (int r = 0; r < kRescore; ++r) {
    float bs = -INFINITY;
    int bi = 0x7FFFFFFF;
    for (int c = lane; c < n_cand; c += 32) {
      const float s = ps[c];
      const int i = pi[c];
      if (i < 0) continue;
      const bool after_last = (s < last_s) || (s == last_s && i > last_i);
      if (!after_last) continue;
      if (s > bs || (s == bs && i < bi)) bs = s, bi = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xFFFFFFFFu, bs, o);
      const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
      if (os > bs || (os == bs && oi < bi)) bs = os, bi = oi;
    }
    if (bi == 0x7FFFFFFF) break;
    sel_s[r] = bs, sel_i[r] = bi;
    last_s = bs, last_i = bi;
    nsel = r + 1;
  }
  // exact fp32 cosine of unit rows for the selected candidates
  if (g_f32 && q_f32) {
    const float* qr = q_f32 + (size_t)row * dim;
    for (int r = 0; r < nsel; ++r) {
      const float* gr = g_f32 + (size_t)sel_i[r] * dim;
      float acc = 0.f;
      for (int i = lane; i < dim; i += 32) acc = fmaf(qr[i], gr[i], acc);
      sel_s[r] = warp_sum(acc);
    }
    // insertion sort by (score desc, index asc); every lane holds the same values
    for (int a = 1; a < nsel; ++a) {
      const float s = sel_s[a];
      const int i = sel_i[a];
      int b = a - 1;
      while (b >= 0 && (sel_s[b] < s || (sel_s[b] == s && sel_i[b] > i))) {
        sel_s[b + 1] = sel_s[b], sel_i[b + 1] = sel_i[b];
        --b;
      }
      sel_s[b + 1] = s, sel_i[b + 1] = i;
    }
  }
  if (lane == 0) {
    for (int t = 0; t < topk; ++t) {
      bool ok = t < nsel;
      if (ok) ok = strict_gt ? (sel_s[t] > threshold) : (sel_s[t] >= threshold);
      out_score[(size_t)row * topk + t] = ok ? sel_s[t] : 0.f;
      out_idx[(size_t)row * topk + t] = ok ? (long long)sel_i[t] + idx_base : -1LL;
    }
  }
}

// ------------------------------------------------------------------------------------------
// duplicate-merge resolve: lexicographically-first maximal independent set over the >=thr graph.
// state: 0 undecided, 1 leader (alive), 2 merged.  A node becomes a leader once every smaller
// neighbour is merged; it is merged as soon as one smaller neighbour is a leader (its lowest such
// neighbour is the leader the reference's ascending-id sweep would have merged it into).
// ------------------------------------------------------------------------------------------
__global__ void cluster_round_kernel(const long long* __restrict__ pairs, long long n_pairs, int n,
                                     const int* __restrict__ state_in, int* __restrict__ blocked,
                                     int* __restrict__ leader) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_pairs;
       t += (long long)gridDim.x * blockDim.x) {
    const long long pr = pairs[t];
    const int i = (int)(pr >> 32), j = (int)(pr & 0xFFFFFFFFll);  // i < j
    if (state_in[j] == 1) continue;             // leaders need nothing; merged nodes keep refining their leader
    const int si = state_in[i];
    if (si == 1) atomicMin(&leader[j], i);       // merged into the lowest leader neighbour
    else if (si == 0 && state_in[j] == 0) blocked[j] = 1;   // cannot decide j yet
  }
}
__global__ void cluster_commit_kernel(int n, int* __restrict__ state, int* __restrict__ blocked,
                                      const int* __restrict__ leader, int* __restrict__ undecided) {
  int local = 0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    if (state[v] != 0) continue;
    if (leader[v] < v) state[v] = 2;
    else if (!blocked[v]) state[v] = 1;
    else local = 1;
    blocked[v] = 0;
  }
  if (local) atomicOr(undecided, 1);
}
__global__ void cluster_init_kernel(int n, int* state, int* blocked, int* leader) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    state[v] = 0, blocked[v] = 0, leader[v] = v;
  }
}

static int grid_for(long long total, int block) {
  long long blocks = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace b2f

using namespace b2f;

extern "C" int b2f_stem_conv3x3(const void* in, int n, int h, int w, int cin_s, int stride, const float* weight,
                                const float* bias, const float* slope, int act, int cout_p, int dtype, void* out,
                                void* stream) {
  B2F_REQUIRE(cin_s == 4, "b2f_stem_conv3x3: input must be stored with 4 channels (got %d)", cin_s);
  B2F_REQUIRE(cout_p % 16 == 0 && cout_p <= 256, "b2f_stem_conv3x3: cout_p %d unsupported", cout_p);
  B2F_REQUIRE(act != 2 || slope != nullptr, "b2f_stem_conv3x3: PReLU needs slope");
  const int ho = (h + 2 - 3) / stride + 1, wo = (w + 2 - 3) / stride + 1;
  const long long total = (long long)n * ho * ((wo + 1) / 2) * (cout_p / 16);
  const size_t smem = (size_t)(36 + 2) * cout_p * sizeof(float);
  stem_conv_kernel<<<grid_for(total, 256), 256, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint16_t*>(in), n, h, w, stride, ho, wo, weight, bias, slope, act, cout_p,
      dtype == B2F_BF16, reinterpret_cast<uint16_t*>(out));
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}


extern "C" int b2f_im2col3x3(const void* in, int n, int h, int w, int stride, int ho, int wo, int dtype, void* out,
                             void* stream) {
  (void)dtype;  // a pure 16-bit move: fp16 and bf16 are handled alike
  B2F_REQUIRE(ho == (h + 2 - 3) / stride + 1 && wo == (w + 2 - 3) / stride + 1, "b2f_im2col3x3: output size mismatch");
  const long long total = (long long)n * ho * wo;
  if (total <= 0) return 0;
  im2col3x3_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint16_t*>(in), n, h, w, stride, ho, wo, reinterpret_cast<uint16_t*>(out));
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2f_dwconv(const void* in, int n, int h, int w, int c_p, int k, int stride, int pad, const float* weight,
                          const float* bias, const float* slope, int act, int dtype, void* out, void* stream) {
  B2F_REQUIRE(c_p % 8 == 0, "b2f_dwconv: channels must be padded to 8");
  B2F_REQUIRE(act != 2 || slope != nullptr, "b2f_dwconv: PReLU needs slope");
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  const long long total = (long long)n * ho * wo * (c_p / 8);
  dwconv_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint16_t*>(in), n, h, w, c_p, k, stride, pad, ho, wo, weight, bias, slope, act,
      dtype == B2F_BF16, reinterpret_cast<uint16_t*>(out));
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2f_pool(const void* in, int n, int h, int w, int c_p, int k, int stride, int pad, int mode, int ho,
                        int wo, int dtype, void* out, void* stream) {
  B2F_REQUIRE(c_p % 8 == 0, "b2f_pool: channels must be padded to 8");
  const long long total = (long long)n * ho * wo * (c_p / 8);
  if (mode == 0 && ho <= 65535 && n <= 65535 && n > 0) {
    const dim3 grid((wo * (c_p / 8) + 255) / 256, (ho + kPoolRows - 1) / kPoolRows, n);
    if (dtype == B2F_BF16)
      maxpool_rows_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint16_t*>(in), h, w, c_p, k,
                                                                       stride, pad, ho, wo, reinterpret_cast<uint16_t*>(out));
    else
      maxpool_rows_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint16_t*>(in), h, w, c_p, k,
                                                                        stride, pad, ho, wo, reinterpret_cast<uint16_t*>(out));
    g_launches.fetch_add(1);
    B2F_LAUNCH_CHECK();
    return 0;
  }
  pool_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint16_t*>(in), n, h, w,
                                                                     c_p, k, stride, pad, mode, ho, wo,
                                                                     dtype == B2F_BF16, reinterpret_cast<uint16_t*>(out));
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}


extern "C" int b2f_eltwise(const void* a, const void* b, long long pixels, int c_p, const float* scale,
                           const float* shift, const float* slope, int act, int dtype, void* out, void* stream) {
  B2F_REQUIRE(c_p % 8 == 0, "b2f_eltwise: channels must be padded to 8");
  B2F_REQUIRE((scale == nullptr) == (shift == nullptr), "b2f_eltwise: scale and shift come together");
  B2F_REQUIRE(act != 2 || slope != nullptr, "b2f_eltwise: PReLU needs slope");
  if (pixels <= 0) return 0;
  eltwise_kernel<<<grid_for(pixels * (c_p / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint16_t*>(a), reinterpret_cast<const uint16_t*>(b), pixels, c_p, scale, shift, slope, act,
      dtype == B2F_BF16, reinterpret_cast<uint16_t*>(out));
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2f_l2norm_rows(const float* x, long long rows, int dim, float* out_f32, void* out_16, int dtype,
                               float* norms, void* stream) {
  if (rows <= 0) return 0;
  l2norm_kernel<<<grid_for(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      x, rows, dim, out_f32, reinterpret_cast<uint16_t*>(out_16), dtype == B2F_BF16, norms);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

// one unit query against every stored unit row: out[r] = <rows[r], query>, a warp per row (HBM-bound: rows * dim * 4 bytes)
__global__ void __launch_bounds__(256)
rows_dot_kernel(const float* __restrict__ rows, long long n, int dim, const float* __restrict__ query, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* pr = rows + (size_t)row * dim;
  float acc = 0.f;
  for (int i = lane; i < dim; i += 32) acc = fmaf(pr[i], __ldg(query + i), acc);
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}
extern "C" int b2f_rows_dot(const float* rows, long long n, int dim, const float* query, float* out, void* stream) {
  if (n <= 0) return 0;
  rows_dot_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, n, dim, query, out);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2f_cosine_pairs(const float* a, const float* b, int pairs, int dim, float* out, void* stream) {
  if (pairs <= 0) return 0;
  cosine_pairs_kernel<<<(pairs * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a, b, pairs, dim, out);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2f_match_merge(const float* part_score, const int* part_idx, int q, int n_cand, const float* q_f32,
                               const float* g_f32, int dim, int topk, float threshold, int strict_gt,
                               long long idx_base, float* out_score, long long* out_idx, void* stream) {
  B2F_REQUIRE(topk >= 1 && topk <= kRescore, "b2f_match_merge: topk must be in [1,%d]", kRescore);
  if (q <= 0) return 0;
  match_merge_kernel<<<(q * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      part_score, part_idx, q, n_cand, q_f32, g_f32, dim, topk, threshold, strict_gt, idx_base, out_score, out_idx);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

// ---- sharded top-1 exchange: (score, global index) <-> one 64-bit key whose SIGNED integer maximum is the winner by
// (score descending, index ascending), so the cross-shard merge of reference-order top-1 lists is a single MAX reduction
// (int64 is what both NCCL and gloo reduce; the top bit is flipped so signed order == the unsigned order of the fields)
__device__ __forceinline__ uint32_t score_order_bits(float s) {
  const uint32_t u = __float_as_uint(s);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__global__ void topk_pack_keys_kernel(const float* __restrict__ score, const long long* __restrict__ idx, long long n,
                                      unsigned long long* __restrict__ keys) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long g = idx[i];
  const unsigned long long k = g < 0 ? 0ull : (((unsigned long long)score_order_bits(score[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)g));
  keys[i] = k ^ 0x8000000000000000ull;
}
__global__ void topk_unpack_keys_kernel(const unsigned long long* __restrict__ keys, long long n, float* __restrict__ score,
                                        long long* __restrict__ idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i] ^ 0x8000000000000000ull;
  if (k == 0ull) {
    score[i] = 0.f, idx[i] = -1;
    return;
  }
  const uint32_t ob = (uint32_t)(k >> 32);
  score[i] = __uint_as_float(ob ^ ((ob >> 31) ? 0x80000000u : 0xFFFFFFFFu));
  idx[i] = (long long)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
}
extern "C" int b2f_topk_pack_keys(const float* score, const long long* idx, long long n, long long* keys, void* stream) {
  if (n <= 0) return 0;
  topk_pack_keys_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(score, idx, n, reinterpret_cast<unsigned long long*>(keys));
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}
extern "C" int b2f_topk_unpack_keys(const long long* keys, long long n, float* score, long long* idx, void* stream) {
  if (n <= 0) return 0;
  topk_unpack_keys_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(keys), n, score, idx);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

// leader[] doubles as output; workspace (3 ints per node + 1) is carved from the tail of `leader`'s
// caller-provided scratch: state = leader + n, blocked = leader + 2n, flag = leader + 3n.
extern "C" int b2f_cluster_resolve(const long long* pairs, long long n_pairs, int n, int* leader, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n <= 0) return 0;
  int* state = leader + n;
  int* blocked = leader + 2 * (size_t)n;
  int* flag = leader + 3 * (size_t)n;
  cluster_init_kernel<<<grid_for(n, 256), 256, 0, stream>>>(n, state, blocked, leader);
  g_launches.fetch_add(1);
  for (int round = 0; round < n + 1; ++round) {
    B2F_CHECK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), stream));
    if (n_pairs > 0) {
      cluster_round_kernel<<<grid_for(n_pairs, 256), 256, 0, stream>>>(pairs, n_pairs, n, state, blocked, leader);
      g_launches.fetch_add(1);
    }
    cluster_commit_kernel<<<grid_for(n, 256), 256, 0, stream>>>(n, state, blocked, leader, flag);
    g_launches.fetch_add(1);
    int h_flag = 0;
    B2F_CHECK_CUDA(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    B2F_CHECK_CUDA(cudaStreamSynchronize(stream));
    if (!h_flag) break;
  }
  if (n_pairs > 0) {  // leaders decided in the last round still have to claim their merged neighbours
    cluster_round_kernel<<<grid_for(n_pairs, 256), 256, 0, stream>>>(pairs, n_pairs, n, state, blocked, leader);
    g_launches.fetch_add(1);
  }
  B2F_LAUNCH_CHECK();
  return 0;
}
