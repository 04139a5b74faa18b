// Integer / float32-exact pre- and post-processing kernels (HBM- and latency-bound; CUDA cores).
//   letterbox / preprocess / blob   : reference models/scrfd.py:122-138, 76-82; models/arcface.py:44-50
//   decode + threshold + sort + NMS : reference models/scrfd.py:89-119,142-177,180-207; utils/helpers.py:62-107
//   estimate_norm / warpAffine      : reference utils/helpers.py:18-59
// Every arithmetic step that the reference performs in numpy float32 / cv2 fixed point is spelled
// with round-to-nearest intrinsics so the compiler cannot contract it into FMAs.
#include "b2f_common.cuh"
#include "../../include/b2f.h"

#include <atomic>

namespace b2f {
extern std::atomic<long long> g_launches;

// ============================================================================================
// resize + letterbox
// ============================================================================================
struct ResizeGeom {
  int H, W, new_w, new_h, in_w, in_h;
  int mode;  // 0 copy, 1 exact-2x area average, 2 fixed-point bilinear, 3 odd integer ratio (bilinear weights collapse)
  int ratio;
  double scale_x, scale_y;
};

__device__ __forceinline__ void linear_coeff(int d, double scale, int src, bool horizontal, int& s0, int& s1, int& w0,
                                             int& w1) {
  const double fd = __dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  float f = __double2float_rn(fd);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (horizontal) {
    if (s < 0) {
      f = 0.f;
      s = 0;
    }
    if (s >= src - 1) {
      f = 0.f;
      s = src - 1;
    }
  }
  w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  s0 = min(max(s, 0), src - 1);
  s1 = min(max(s + 1, 0), src - 1);
}

// one letterboxed BGR pixel (uint8 values as int)
__device__ __forceinline__ void letterbox_pixel(const uint8_t* __restrict__ img, const ResizeGeom& g, int x, int y,
                                                int (&bgr)[3]) {
  if (x >= g.new_w || y >= g.new_h) {
    bgr[0] = bgr[1] = bgr[2] = 0;
    return;
  }
  if (g.mode == 0) {
    const uint8_t* p = img + ((size_t)y * g.W + x) * 3;
    bgr[0] = p[0], bgr[1] = p[1], bgr[2] = p[2];
  } else if (g.mode == 1) {
    const uint8_t* p0 = img + ((size_t)(2 * y) * g.W + 2 * x) * 3;
    const uint8_t* p1 = p0 + (size_t)g.W * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) bgr[c] = (p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2;
  } else if (g.mode == 3) {
    // odd integer ratio r: the source coordinate (d + 0.5) r - 0.5 = r d + (r - 1) / 2 is an integer, the fixed-point
    // weights are (2048, 0) and cv2's INTER_LINEAR returns that pixel exactly (1080p -> 640 x 360 is img[1::3, 1::3])
    const int o = (g.ratio - 1) >> 1;
    const uint8_t* p = img + ((size_t)(y * g.ratio + o) * g.W + (x * g.ratio + o)) * 3;
    bgr[0] = p[0], bgr[1] = p[1], bgr[2] = p[2];
  } else {
    int x0, x1, a0, a1, y0, y1, b0, b1;
    linear_coeff(x, g.scale_x, g.W, true, x0, x1, a0, a1);
    linear_coeff(y, g.scale_y, g.H, false, y0, y1, b0, b1);
    const uint8_t* r0 = img + (size_t)y0 * g.W * 3;
    const uint8_t* r1 = img + (size_t)y1 * g.W * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int h0 = r0[x0 * 3 + c] * a0 + r0[x1 * 3 + c] * a1;
      const int h1 = r1[x0 * 3 + c] * a0 + r1[x1 * 3 + c] * a1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      bgr[c] = min(max(v, 0), 255);
    }
  }
}

template <int OUT>  // 0: uint8 BGR canvas, 1: normalised 16-bit NHWC RGB (channel-padded)
__global__ void letterbox_kernel(const uint8_t* __restrict__ frames, ResizeGeom g, int batch, float mean, float scale,
                                 void* __restrict__ out, int c_pad, int is_bf16) {
  const long long total = (long long)batch * g.in_h * g.in_w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % g.in_w);
    const int y = (int)((i / g.in_w) % g.in_h);
    const int b = (int)(i / ((long long)g.in_w * g.in_h));
    int bgr[3];
    letterbox_pixel(frames + (size_t)b * g.H * g.W * 3, g, x, y, bgr);
    if (OUT == 0) {
      uint8_t* o = reinterpret_cast<uint8_t*>(out) + (size_t)i * 3;
      o[0] = (uint8_t)bgr[0], o[1] = (uint8_t)bgr[1], o[2] = (uint8_t)bgr[2];
    } else {
      float v[4];
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = __fmul_rn(__fsub_rn((float)bgr[2 - c], mean), scale);
      v[3] = 0.f;
      uint16_t* o = reinterpret_cast<uint16_t*>(out) + (size_t)i * c_pad;
      uint16_t h[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (is_bf16) {
          __nv_bfloat16 t = __float2bfloat16_rn(v[c]);
          h[c] = *reinterpret_cast<uint16_t*>(&t);
        } else {
          __half t = __float2half_rn(v[c]);
          h[c] = *reinterpret_cast<uint16_t*>(&t);
        }
      }
      *reinterpret_cast<uint2*>(o) = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
      for (int c = 4; c < c_pad; c += 4) *reinterpret_cast<uint2*>(o + c) = make_uint2(0u, 0u);
    }
  }
}


// letterbox + normalise + 3x3 / pad 1 patch extraction in one pass: out [b][ho][wo][32], k = tap*3 + rgb (27 used).
// A CTA owns a 32 x 8 tile of output pixels: it evaluates the letterboxed source pixels the tile touches ONCE into
// shared memory (uint8 BGR + an inside-the-canvas flag), then every thread assembles the 27 values of its pixel.
// Taps outside the in_h x in_w canvas are the convolution's zero padding; the letterbox pad inside it is a real
// (normalised) zero pixel, as in the reference blob (models/scrfd.py:76-82, 135-138).
constexpr int kLpW = 32, kLpH = 8;
__device__ __forceinline__ uint16_t norm16(int u8v, float mean, float scale, int is_bf16) {
  const float f = __fmul_rn(__fsub_rn((float)u8v, mean), scale);
  if (is_bf16) {
    __nv_bfloat16 t = __float2bfloat16_rn(f);
    return *reinterpret_cast<uint16_t*>(&t);
  }
  __half t = __float2half_rn(f);
  return *reinterpret_cast<uint16_t*>(&t);
}

__global__ void __launch_bounds__(kLpW * kLpH)
letterbox_patches_kernel(const uint8_t* __restrict__ frames, ResizeGeom g, int stride, int ho, int wo, float mean,
                         float scale, uint16_t* __restrict__ out, int is_bf16) {
  constexpr int kMaxIw = kLpW * 2 + 1, kMaxIh = kLpH * 2 + 1;
  __shared__ uint8_t tile[kMaxIh * kMaxIw * 4];                      // b, g, r, inside flag
  const int b = blockIdx.z, ox0 = blockIdx.x * kLpW, oy0 = blockIdx.y * kLpH;
  const int iw = (kLpW - 1) * stride + 3, ih = (kLpH - 1) * stride + 3;
  const int ix0 = ox0 * stride - 1, iy0 = oy0 * stride - 1;
  const uint8_t* img = frames + (size_t)b * g.H * g.W * 3;
  for (int t = threadIdx.x; t < iw * ih; t += blockDim.x) {
    const int ty = t / iw, tx = t - ty * iw;
    const int iy = iy0 + ty, ix = ix0 + tx;
    int bgr[3] = {0, 0, 0};
    const bool inside = iy >= 0 && iy < g.in_h && ix >= 0 && ix < g.in_w;
    if (inside) letterbox_pixel(img, g, ix, iy, bgr);
    *reinterpret_cast<uchar4*>(tile + t * 4) = make_uchar4((unsigned char)bgr[0], (unsigned char)bgr[1],
                                                           (unsigned char)bgr[2], inside ? 1 : 0);
  }
  __syncthreads();
  // emission: one thread per output pixel (27 values from the tile; measured faster than piece-per-lane coalescing)
  const int lx = threadIdx.x % kLpW, ly = threadIdx.x / kLpW;
  const int ox = ox0 + lx, oy = oy0 + ly;
  if (ox >= wo || oy >= ho) return;
  uint16_t v[32];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uchar4 px = *reinterpret_cast<const uchar4*>(tile + ((ly * stride + tap / 3) * iw + lx * stride + tap % 3) * 4);
    v[tap * 3 + 0] = px.w ? norm16(px.z, mean, scale, is_bf16) : (uint16_t)0;      // R
    v[tap * 3 + 1] = px.w ? norm16(px.y, mean, scale, is_bf16) : (uint16_t)0;      // G
    v[tap * 3 + 2] = px.w ? norm16(px.x, mean, scale, is_bf16) : (uint16_t)0;      // B
  }
#pragma unroll
  for (int k = 27; k < 32; ++k) v[k] = 0;
  uint4* o = reinterpret_cast<uint4*>(out + (((size_t)b * ho + oy) * wo + ox) * 32);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o[j] = make_uint4(v[8 * j] | ((uint32_t)v[8 * j + 1] << 16), v[8 * j + 2] | ((uint32_t)v[8 * j + 3] << 16),
                      v[8 * j + 4] | ((uint32_t)v[8 * j + 5] << 16), v[8 * j + 6] | ((uint32_t)v[8 * j + 7] << 16));
}

// ---- letterbox + normalise + the detector's FIRST CONVOLUTION (3x3 / stride 2 / pad 1, 3 -> cout_p <= 32) in one pass ----
// The stem layer has K = 27: its arithmetic is 0.6 % of the detector's and it is bound by writing its 320 x 320 x 32
// output, so it is computed where the pixels are produced instead of materialising a 419 MB patch tensor for the
// convolution kernel (one HBM round trip and one launch less).  A CTA owns a 32 x 8 tile of output pixels:
//   1. the letterboxed source pixels the tile touches are evaluated once into shared memory (uint8 BGR + an
//      inside-the-canvas flag: taps outside the canvas are the convolution's zero padding, the letterbox pad inside it a
//      real normalised zero pixel, reference models/scrfd.py:76-82, 135-138);
//   2. every thread assembles the 27 normalised values of its pixel (k = tap * 3 + rgb, the order of the weight layout
//      [cout_p][32]) and writes them as one 64-byte row of the A operand -- two 128-row K-major tiles in the canonical
//      64-byte-swizzled layout a TMA load would have produced; the weights go to shared memory the same way;
//   3. one thread issues four tcgen05.mma (2 tiles x K = 32 in two steps, M = 128, N = cout_p) into 2 x cout_p TMEM columns;
//   4. each warp reads its TMEM lane quarter back (tcgen05.ld), adds the bias, applies ReLU and stores the pixel's 64
//      bytes (a warp covers 2 KB of contiguous output).
// The CTAs are persistent (four per SM): TMEM, the barrier, the normalisation table, weights and bias are set up once,
// and the source pixels of the NEXT tile are fetched into registers (one packed register per pixel) while the tensor
// core and the epilogue work on the current one, so the strided gather's latency is off the critical path.
template <bool BF16, int COUT>
__global__ void __launch_bounds__(kLpW * kLpH, 4)
letterbox_conv1_kernel(const uint8_t* __restrict__ frames, ResizeGeom g, int ho, int wo, float mean, float scale,
                       const uint16_t* __restrict__ weight, const float* __restrict__ bias, int act,
                       uint16_t* __restrict__ out, int tiles_x, int tiles_y, int n_tiles) {
  constexpr int IW = kLpW * 2 + 1, IH = kLpH * 2 + 1;
  constexpr int kSweeps = (IW * IH + kLpW * kLpH - 1) / (kLpW * kLpH);
  constexpr uint32_t kCols = 2 * COUT < 32 ? 32 : 2 * COUT;          // TMEM columns: two accumulators of COUT
  __shared__ __align__(1024) uint8_t a_tile[2 * 128 * 64];            // [m tile][row][32 k] 16-bit, SWIZZLE_64B
  __shared__ __align__(1024) uint8_t b_tile[COUT * 64];               // [cout][32 k] 16-bit, SWIZZLE_64B
  __shared__ __align__(8) uint16_t tile[IH * IW * 4];                 // normalised R, G, B, 0 (all zero outside the canvas)
  __shared__ uint16_t lut[256];
  __shared__ __align__(16) float s_bias[COUT];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // ---- once per CTA (the CTA is persistent: it walks tiles blockIdx.x, + gridDim.x, ...) ----
  if (warp == 0) {
    tmem_alloc(&tmem_slot, kCols);
    tmem_relinquish();
  }
  if (tid == 32) {
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  lut[tid] = norm16(tid, mean, scale, BF16);                          // blockDim.x == 256
  if (tid < COUT) s_bias[tid] = bias[tid];
  for (int c = tid; c < COUT * 4; c += blockDim.x) {                  // weights: 16-byte chunks into the swizzled rows
    const uint32_t o = (uint32_t)c * 16;
    *reinterpret_cast<uint4*>(b_tile + (o ^ (((o >> 7) & 3u) << 4))) = __ldg(reinterpret_cast<const uint4*>(weight) + c);
  }
  fence_proxy_async();                                                // the weight tile is read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const bool direct = g.mode == 0 || g.mode == 3;
  const int ratio = g.mode == 3 ? g.ratio : 1, roff = (ratio - 1) >> 1;
  const int lx = tid & (kLpW - 1), ly = tid / kLpW;
  const int tiles_xy = tiles_x * tiles_y;

  // source pixels of a tile as packed b | g << 8 | r << 16 | inside << 24, one register per sweep.  The odd-ratio and
  // copy modes (1080p -> 640 x 360 is img[1::3, 1::3]) take a byte-offset fast path, the general bilinear / 2x-area modes
  // go through letterbox_pixel.  All loads of a thread are issued before the first use.
  auto fetch = [&](int t_idx, uint32_t (&px)[kSweeps]) {
    const int b = t_idx / tiles_xy, rem = t_idx - b * tiles_xy;
    const int ty0 = rem / tiles_x, tx0 = rem - ty0 * tiles_x;
    const int ix0 = tx0 * kLpW * 2 - 1, iy0 = ty0 * kLpH * 2 - 1;
    const uint8_t* img = frames + (size_t)b * g.H * g.W * 3;
#pragma unroll
    for (int i = 0; i < kSweeps; ++i) {
      const int t = tid + i * (kLpW * kLpH);
      const int ty = t / IW, tx = t - ty * IW;
      const int iy = iy0 + ty, ix = ix0 + tx;
      uint32_t v = 0;
      if (t < IW * IH && ix >= 0 && ix < g.in_w && iy >= 0 && iy < g.in_h) {
        int bgr[3] = {0, 0, 0};
        if (direct) {
          if (ix < g.new_w && iy < g.new_h) {
            const uint8_t* p = img + ((size_t)(iy * ratio + roff) * g.W + (ix * ratio + roff)) * 3;
            bgr[0] = p[0], bgr[1] = p[1], bgr[2] = p[2];
          }
        } else {
          letterbox_pixel(img, g, ix, iy, bgr);
        }
        v = (uint32_t)bgr[0] | ((uint32_t)bgr[1] << 8) | ((uint32_t)bgr[2] << 16) | (1u << 24);
      }
      px[i] = v;
    }
  };

  uint32_t px[kSweeps];
  int t_idx = blockIdx.x;
  if (t_idx < n_tiles) fetch(t_idx, px);
  uint32_t phase = 0;
  for (; t_idx < n_tiles; t_idx += gridDim.x) {
    const int b = t_idx / tiles_xy, rem = t_idx - b * tiles_xy;
    const int ty0 = rem / tiles_x, tx0 = rem - ty0 * tiles_x;
    const int ox0 = tx0 * kLpW, oy0 = ty0 * kLpH;
    // 1. the tile's source pixels, normalised once (taps outside the canvas are the convolution's zero padding)
#pragma unroll
    for (int i = 0; i < kSweeps; ++i) {
      const int t = tid + i * (kLpW * kLpH);
      uint2 v = make_uint2(0u, 0u);
      if (px[i] >> 24) {
        v.x = lut[(px[i] >> 16) & 0xFF] | ((uint32_t)lut[(px[i] >> 8) & 0xFF] << 16);      // R, G
        v.y = lut[px[i] & 0xFF];                                                            // B, 0
      }
      if (t < IW * IH) *reinterpret_cast<uint2*>(tile + t * 4) = v;
    }
    __syncthreads();
    // 2. A operand: thread (ly, lx) = row (tid & 127) of M tile (tid >> 7); k = tap * 3 + rgb, 27 used
    {
      uint16_t v[32];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint2 q = *reinterpret_cast<const uint2*>(tile + ((ly * 2 + tap / 3) * IW + lx * 2 + tap % 3) * 4);
        v[tap * 3 + 0] = (uint16_t)(q.x & 0xFFFFu);           // R
        v[tap * 3 + 1] = (uint16_t)(q.x >> 16);               // G
        v[tap * 3 + 2] = (uint16_t)(q.y & 0xFFFFu);           // B
      }
#pragma unroll
      for (int k = 27; k < 32; ++k) v[k] = 0;
      uint8_t* row_base = a_tile + (tid >> 7) * (128 * 64);
      const uint32_t row_off = (uint32_t)(tid & 127) * 64u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t o = row_off + (uint32_t)j * 16;
        *reinterpret_cast<uint4*>(row_base + (o ^ (((o >> 7) & 3u) << 4))) =
            make_uint4(v[8 * j] | ((uint32_t)v[8 * j + 1] << 16), v[8 * j + 2] | ((uint32_t)v[8 * j + 3] << 16),
                       v[8 * j + 4] | ((uint32_t)v[8 * j + 5] << 16), v[8 * j + 6] | ((uint32_t)v[8 * j + 7] << 16));
      }
    }
    fence_proxy_async();                                        // generic-proxy writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // 3. four MMAs (two M tiles x K = 32 in two steps)
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t idesc = umma_idesc(128, COUT, BF16 ? 1u : 0u);
        const uint64_t bd = umma_smem_desc(smem_u32(b_tile), 64);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint64_t ad = umma_smem_desc(smem_u32(a_tile) + (uint32_t)mt * 128u * 64u, 64);
          umma_f16(tmem_base + (uint32_t)mt * COUT, ad, bd, idesc, 0u);
          umma_f16(tmem_base + (uint32_t)mt * COUT, ad + 2, bd + 2, idesc, 1u);
        }
        umma_commit(&done_bar);
      }
      __syncwarp();
    }
    // the next tile's source pixels travel while the tensor core works and the epilogue stores
    if (t_idx + (int)gridDim.x < n_tiles) fetch(t_idx + gridDim.x, px);
    mbar_wait(&done_bar, phase);
    phase ^= 1u;
    tc_fence_after();
    // 4. epilogue: warp w reads lanes 32 (w & 3) .. +31 of accumulator w >> 2, i.e. output row ly == w, pixel lx == lane
    {
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * COUT;
      uint32_t r[COUT];
      if (COUT == 32) {
        uint32_t (&r32)[32] = reinterpret_cast<uint32_t (&)[32]>(r);
        tmem_ld32(taddr, r32);
      } else {
        uint32_t (&r16)[16] = reinterpret_cast<uint32_t (&)[16]>(r);
        tmem_ld16(taddr, r16);
      }
      tmem_ld_wait();
      const int oy = oy0 + ly, ox = ox0 + lx;
      if (oy < ho && ox < wo) {
        uint4* dst = reinterpret_cast<uint4*>(out + (((size_t)b * ho + oy) * wo + ox) * COUT);
#pragma unroll
        for (int j = 0; j < COUT / 8; ++j) {
          uint32_t w4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float v0 = __uint_as_float(r[8 * j + 2 * i]) + s_bias[8 * j + 2 * i];
            float v1 = __uint_as_float(r[8 * j + 2 * i + 1]) + s_bias[8 * j + 2 * i + 1];
            if (act == 1) v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f);
            if (BF16) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
              w4[i] = *reinterpret_cast<uint32_t*>(&h2);
            } else {
              __half2 h2 = __floats2half2_rn(v0, v1);
              w4[i] = *reinterpret_cast<uint32_t*>(&h2);
            }
          }
          dst[j] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
    }
    // the accumulators are read (next tile's MMAs overwrite them), and every thread is past its reads of `tile`
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem_base, kCols);
}

static int make_geom(ResizeGeom* g, int h, int w, int new_w, int new_h, int in_w, int in_h) {
  B2F_REQUIRE(h > 0 && w > 0 && new_w > 0 && new_h > 0 && new_w <= in_w && new_h <= in_h,
              "letterbox: bad geometry %dx%d -> %dx%d in %dx%d", w, h, new_w, new_h, in_w, in_h);
  g->H = h, g->W = w, g->new_w = new_w, g->new_h = new_h, g->in_w = in_w, g->in_h = in_h;
  g->ratio = 1;
  if (new_w == w && new_h == h)
    g->mode = 0;
  else if (w == 2 * new_w && h == 2 * new_h)
    g->mode = 1;
  else if (w % new_w == 0 && h % new_h == 0 && w / new_w == h / new_h && ((w / new_w) & 1) && w / new_w <= 15)
    g->mode = 3, g->ratio = w / new_w;
  else
    g->mode = 2;
  g->scale_x = 1.0 / ((double)new_w / (double)w);
  g->scale_y = 1.0 / ((double)new_h / (double)h);
  return 0;
}

static int grid_for(long long total, int block) {
  long long blocks = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

__global__ void blob_kernel(const uint8_t* __restrict__ img, int batch, int h, int w, float mean, float scale,
                            float* __restrict__ out) {
  const long long plane = (long long)h * w;
  const long long total = (long long)batch * 3 * plane;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % plane;
    const int c = (int)((i / plane) % 3);
    const long long b = i / (3 * plane);
    const float v = (float)img[(b * plane + pix) * 3 + (2 - c)];
    out[i] = __fmul_rn(__fsub_rn(v, mean), scale);
  }
}

// ============================================================================================
// decode + threshold + sort + greedy NMS + max_num
// ============================================================================================
struct DecodeParams {
  b2f_det_levels lv;
  int batch, in_h, in_w;
  const float* det_scale;
  const int* image_hw;
  float conf, iou;
  int max_num, metric, max_cand, cap2, max_det;
  int forward_mode;  // 1: no NMS, anchor order, unscaled (the reference's SCRFD.forward view)
  float* det;
  float* kps;
  int* keep_idx;
  int* counts;
  uint8_t* ws;
  long long ws_per_frame;
};

constexpr int kSmemKeys = 4096;
constexpr int kAliveWords = 1024;  // up to 32768 candidates

__device__ __forceinline__ uint32_t float_order_bits(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_bits(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

__device__ void bitonic_sort_desc(uint64_t* k, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t a = k[lo], b = k[hi];
        if (desc ? (a < b) : (a > b)) {
          k[lo] = b;
          k[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

struct AnchorRef {
  int level, pix, anchor, stride, ws;
};
__device__ __forceinline__ AnchorRef anchor_ref(int a, int in_h, int in_w) {
  AnchorRef r;
  int base = 0;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const int s = 8 << l;
    const int cnt = (in_h / s) * (in_w / s) * 2;
    if (a < base + cnt || l == 2) {
      r.level = l, r.stride = s, r.ws = in_w / s;
      r.pix = (a - base) >> 1, r.anchor = (a - base) & 1;
      return r;
    }
    base += cnt;
  }
  return r;
}

__global__ void __launch_bounds__(1024, 1) decode_nms_kernel(DecodeParams p) {
  extern __shared__ uint64_t smem_keys[];
  __shared__ uint32_t alive[kAliveWords];
  __shared__ int word_prefix[kAliveWords];
  __shared__ int s_count, s_keep;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;
  uint8_t* ws = p.ws + (size_t)b * p.ws_per_frame;
  uint64_t* keys_g = reinterpret_cast<uint64_t*>(ws);
  uint64_t* keys2 = keys_g + p.cap2;
  float4* boxes = reinterpret_cast<float4*>(keys2 + p.cap2);
  int* kept_pos = reinterpret_cast<int*>(boxes + p.max_cand);
  int* sel = kept_pos + p.max_cand;

  int lvl_cnt[3], lvl_base[3];
  int total = 0;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const int s = 8 << l;
    lvl_cnt[l] = (p.in_h / s) * (p.in_w / s) * 2;
    lvl_base[l] = total;
    total += lvl_cnt[l];
  }
  if (tid == 0) s_count = 0;
  __syncthreads();

  // ---- A. threshold + ordered stream compaction (warp ballots + a block-level prefix per 1024-anchor sweep) ------
  // Candidates take slots in ANCHOR ORDER, so when a frame has more than max_cand of them the ones kept are always the
  // first max_cand by anchor index (deterministic; counts[b][3] flags the overflow).  The sort below restores the
  // reference's visiting order.  Keys live in shared memory when they fit, else in the global workspace.
  uint64_t* keys = (p.cap2 <= kSmemKeys) ? smem_keys : keys_g;
  {
    const int lane_a = tid & 31, warp_a = tid >> 5, nwarps_a = nthreads >> 5;
    int base = 0;
    for (int a0 = 0; a0 < total; a0 += nthreads) {
      const int a = a0 + tid;
      bool cand = false;
      float sc = 0.f;
      if (a < total) {
        const int l = a < lvl_base[1] ? 0 : (a < lvl_base[2] ? 1 : 2);
        const int local = a - lvl_base[l];
        const size_t npix = (size_t)(lvl_cnt[l] >> 1);
        sc = p.lv.score[l][((size_t)b * npix + (local >> 1)) * p.lv.score_ps[l] + (local & 1)];
        cand = sc >= p.conf;
      }
      const uint32_t bal = __ballot_sync(0xFFFFFFFFu, cand);
      if (lane_a == 0) word_prefix[warp_a] = __popc(bal);          // (word_prefix is free until stage E)
      __syncthreads();
      int before = 0, sweep = 0;
      for (int w = 0; w < nwarps_a; ++w) {
        const int c = word_prefix[w];
        before += w < warp_a ? c : 0;
        sweep += c;
      }
      if (cand) {
        const int slot = base + before + __popc(bal & ((1u << lane_a) - 1u));
        if (slot < p.max_cand)
          keys[slot] = (p.forward_mode ? 0ull : ((uint64_t)float_order_bits(sc) << 32)) |
                       (uint64_t)(0xFFFFFFFFu - (uint32_t)a) | (p.forward_mode ? (1ull << 63) : 0ull);
      }
      base += sweep;
      __syncthreads();
    }
    if (tid == 0) s_count = base;
  }
  __syncthreads();
  const int n_found = s_count;
  const int n = min(n_found, p.max_cand);
  int P = 32;
  while (P < n) P <<= 1;
  for (int i = n + tid; i < P; i += nthreads) keys[i] = 0ull;
  // ---- B. sort: (score desc, anchor asc) == the order the reference's NMS visits candidates ------
  bitonic_sort_desc(keys, P);

  // ---- C. decode boxes for the sorted candidates ------------------------------------------------
  const float dscale = p.det_scale[b];
  for (int i = tid; i < n; i += nthreads) {
    const int a = (int)(0xFFFFFFFFu - (uint32_t)(keys[i] & 0xFFFFFFFFull));
    const AnchorRef r = anchor_ref(a, p.in_h, p.in_w);
    const size_t npix = (size_t)(lvl_cnt[r.level] >> 1);
    const float* bp = p.lv.bbox[r.level] + ((size_t)b * npix + r.pix) * p.lv.bbox_ps[r.level] + r.anchor * 4;
    const float fs = (float)r.stride;
    const float cx = (float)((r.pix % r.ws) * r.stride), cy = (float)((r.pix / r.ws) * r.stride);
    float4 bx;
    bx.x = __fdiv_rn(__fsub_rn(cx, __fmul_rn(bp[0], fs)), dscale);
    bx.y = __fdiv_rn(__fsub_rn(cy, __fmul_rn(bp[1], fs)), dscale);
    bx.z = __fdiv_rn(__fadd_rn(cx, __fmul_rn(bp[2], fs)), dscale);
    bx.w = __fdiv_rn(__fadd_rn(cy, __fmul_rn(bp[3], fs)), dscale);
    boxes[i] = bx;
  }
  const int words = (n + 31) >> 5;
  for (int w = tid; w < words; w += nthreads) {
    const int rem = n - w * 32;
    alive[w] = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
  }
  __syncthreads();

  // ---- D. greedy NMS over a shared alive bitmask: one warp owns one 32-candidate word at a time ----
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  for (int i = 0; i < (p.forward_mode ? 0 : n); ++i) {
    if (!((alive[i >> 5] >> (i & 31)) & 1u)) continue;  // block-uniform: bits only ever clear, and only for j > i
    const float4 bi = boxes[i];
    const float ai = __fmul_rn(__fadd_rn(__fsub_rn(bi.z, bi.x), 1.f), __fadd_rn(__fsub_rn(bi.w, bi.y), 1.f));
    for (int w = ((i + 1) >> 5) + warp; w < words; w += nwarps) {
      const int j = w * 32 + lane;
      bool suppress = false;
      if (j > i && j < n) {
        const float4 bj = boxes[j];
        const float aj = __fmul_rn(__fadd_rn(__fsub_rn(bj.z, bj.x), 1.f), __fadd_rn(__fsub_rn(bj.w, bj.y), 1.f));
        // np.maximum(0.0, x) propagates a NaN x (fmaxf would drop it): keep it, so a NaN overlap suppresses as in numpy
        const float w0 = __fadd_rn(__fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), 1.f);
        const float h0 = __fadd_rn(__fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)), 1.f);
        const float ww = w0 < 0.f ? 0.f : w0, hh = h0 < 0.f ? 0.f : h0;
        const float inter = __fmul_rn(ww, hh);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, aj), inter));
        suppress = !(ovr <= p.iou);  // NaN (0/0 on degenerate boxes) suppresses, as np.where(ovr <= thr) does
      }
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, suppress);
      if (lane == 0 && m) alive[w] &= ~m;
    }
    __syncthreads();
  }

  // ---- E. rank the survivors ----------------------------------------------------------------------
  if (tid == 0) {
    int acc = 0;
    for (int w = 0; w < words; ++w) {
      word_prefix[w] = acc;
      acc += __popc(alive[w]);
    }
    s_keep = acc;
  }
  __syncthreads();
  const int n_keep = s_keep;
  for (int i = tid; i < n; i += nthreads) {
    const uint32_t wbits = alive[i >> 5];
    if ((wbits >> (i & 31)) & 1u) kept_pos[word_prefix[i >> 5] + __popc(wbits & ((1u << (i & 31)) - 1u))] = i;
  }
  __syncthreads();

  // ---- F. optional max_num selection by area (or centre-weighted area) -------------------------------
  int n_out = n_keep;
  const bool select = p.max_num > 0 && p.max_num < n_keep;
  if (select) {
    int P2 = 32;
    while (P2 < n_keep) P2 <<= 1;
    uint64_t* k2 = (P2 <= kSmemKeys && keys != smem_keys) ? smem_keys : keys2;
    float cx = 0.f, cy = 0.f;
    if (p.metric != 0 && p.image_hw) {
      cy = (float)(p.image_hw[2 * b] / 2);
      cx = (float)(p.image_hw[2 * b + 1] / 2);
    }
    for (int r = tid; r < P2; r += nthreads) {
      uint64_t key = 0ull;
      if (r < n_keep) {
        const float4 bx = boxes[kept_pos[r]];
        float v = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
        if (p.metric != 0) {
          const float ox = __fsub_rn(__fdiv_rn(__fadd_rn(bx.x, bx.z), 2.f), cx);
          const float oy = __fsub_rn(__fdiv_rn(__fadd_rn(bx.y, bx.w), 2.f), cy);
          v = __fsub_rn(v, __fmul_rn(__fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy)), 2.f));
        }
        key = ((uint64_t)float_order_bits(v) << 32) | (uint32_t)r;
      }
      k2[r] = key;
    }
    bitonic_sort_desc(k2, P2);
    n_out = p.max_num;
    for (int t = tid; t < n_out; t += nthreads) sel[t] = (int)(k2[t] & 0xFFFFFFFFull);
    __syncthreads();
  }
  const int n_write = min(n_out, p.max_det);

  // ---- G. emit rows ------------------------------------------------------------------------------------
  for (int t = tid; t < n_write; t += nthreads) {
    const int r = select ? sel[t] : t;
    const int i = kept_pos[r];
    const uint64_t key = keys[i];
    const int a = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
    const AnchorRef ar = anchor_ref(a, p.in_h, p.in_w);
    const float score = p.lv.score[ar.level][((size_t)b * (size_t)(lvl_cnt[ar.level] >> 1) + ar.pix) *
                                                 p.lv.score_ps[ar.level] + ar.anchor];
    const float4 bx = boxes[i];
    float* d = p.det + ((size_t)b * p.max_det + t) * 5;
    d[0] = bx.x, d[1] = bx.y, d[2] = bx.z, d[3] = bx.w, d[4] = score;
    const size_t npix = (size_t)(lvl_cnt[ar.level] >> 1);
    const float* kp = p.lv.kps[ar.level] + ((size_t)b * npix + ar.pix) * p.lv.kps_ps[ar.level] + ar.anchor * 10;
    const float fs = (float)ar.stride;
    const float cx = (float)((ar.pix % ar.ws) * ar.stride), cy = (float)((ar.pix / ar.ws) * ar.stride);
    float* ko = p.kps + ((size_t)b * p.max_det + t) * 10;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      ko[2 * j] = __fdiv_rn(__fadd_rn(cx, __fmul_rn(kp[2 * j], fs)), dscale);
      ko[2 * j + 1] = __fdiv_rn(__fadd_rn(cy, __fmul_rn(kp[2 * j + 1], fs)), dscale);
    }
    if (p.keep_idx && p.forward_mode) {
      p.keep_idx[(size_t)b * p.max_det + t] = a;
    } else if (p.keep_idx) {
      // index into the reference's pre_det order (score desc, anchor DESC): mirror i inside its tie run
      const uint32_t sb = (uint32_t)(key >> 32);
      int rs = i, re = i + 1;
      while (rs > 0 && (uint32_t)(keys[rs - 1] >> 32) == sb) --rs;
      while (re < n && (uint32_t)(keys[re] >> 32) == sb) ++re;
      p.keep_idx[(size_t)b * p.max_det + t] = rs + (re - 1 - i);
    }
  }
  if (tid == 0) {
    int* c = p.counts + 4 * b;
    c[0] = n_write;
    c[1] = n_found;
    c[2] = n_keep;
    c[3] = (n_found > p.max_cand ? 1 : 0) | (n_out > p.max_det ? 2 : 0);
  }
}


// ============================================================================================
// stand-alone greedy NMS over an arbitrary (K,5) array (reference SCRFD.nms, models/scrfd.py:180-207)
// and the two anchor-decode helpers (reference utils/helpers.py:62-107)
// ============================================================================================
__global__ void __launch_bounds__(1024, 1)
nms_kernel(const float* __restrict__ dets, int n, float iou, uint64_t* __restrict__ keys_g, int P,
           int* __restrict__ keep, int* __restrict__ n_keep) {
  extern __shared__ uint64_t smem_keys[];
  __shared__ uint32_t alive[kAliveWords];
  const int tid = threadIdx.x, nthreads = blockDim.x;
  uint64_t* keys = (P <= kSmemKeys) ? smem_keys : keys_g;
  // scores.argsort()[::-1] with a stable sort: score descending, equal scores by DEscending index
  for (int i = tid; i < P; i += nthreads)
    keys[i] = i < n ? (((uint64_t)float_order_bits(dets[5 * i + 4]) << 32) | (uint32_t)i) : 0ull;
  bitonic_sort_desc(keys, P);
  const int words = (n + 31) >> 5;
  for (int w = tid; w < words; w += nthreads) {
    const int rem = n - w * 32;
    alive[w] = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  for (int i = 0; i < n; ++i) {
    if (!((alive[i >> 5] >> (i & 31)) & 1u)) continue;
    const float* di = dets + 5 * (size_t)(keys[i] & 0xFFFFFFFFull);
    const float bx1 = di[0], by1 = di[1], bx2 = di[2], by2 = di[3];
    const float ai = __fmul_rn(__fadd_rn(__fsub_rn(bx2, bx1), 1.f), __fadd_rn(__fsub_rn(by2, by1), 1.f));
    for (int w = ((i + 1) >> 5) + warp; w < words; w += nwarps) {
      const int j = w * 32 + lane;
      bool suppress = false;
      if (j > i && j < n) {
        const float* dj = dets + 5 * (size_t)(keys[j] & 0xFFFFFFFFull);
        const float aj = __fmul_rn(__fadd_rn(__fsub_rn(dj[2], dj[0]), 1.f), __fadd_rn(__fsub_rn(dj[3], dj[1]), 1.f));
        const float w0 = __fadd_rn(__fsub_rn(fminf(bx2, dj[2]), fmaxf(bx1, dj[0])), 1.f);
        const float h0 = __fadd_rn(__fsub_rn(fminf(by2, dj[3]), fmaxf(by1, dj[1])), 1.f);
        const float ww = w0 < 0.f ? 0.f : w0, hh = h0 < 0.f ? 0.f : h0;   // np.maximum(0.0, x) keeps a NaN x
        const float inter = __fmul_rn(ww, hh);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, aj), inter));
        suppress = !(ovr <= iou);
      }
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, suppress);
      if (lane == 0 && m) alive[w] &= ~m;
    }
    __syncthreads();
  }
  if (tid == 0) {
    int c = 0;
    for (int i = 0; i < n; ++i)
      if ((alive[i >> 5] >> (i & 31)) & 1u) keep[c++] = (int)(keys[i] & 0xFFFFFFFFull);
    *n_keep = c;
  }
}

__global__ void distance2bbox_kernel(const float* __restrict__ pts, const float* __restrict__ d, int n,
                                     float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pts[2 * i], y = pts[2 * i + 1];
  out[4 * i + 0] = __fsub_rn(x, d[4 * i + 0]);
  out[4 * i + 1] = __fsub_rn(y, d[4 * i + 1]);
  out[4 * i + 2] = __fadd_rn(x, d[4 * i + 2]);
  out[4 * i + 3] = __fadd_rn(y, d[4 * i + 3]);
}
__global__ void distance2kps_kernel(const float* __restrict__ pts, const float* __restrict__ d, int n, int k2,
                                    float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * k2) return;
  const int i = t / k2, j = t % k2;
  out[t] = __fadd_rn(pts[2 * i + (j & 1)], d[t]);
}

// ============================================================================================
// five-point similarity + warpAffine
// ============================================================================================
__constant__ float c_template[10] = {38.2946f, 51.6963f, 73.5318f, 51.5014f, 56.0252f,
                                     71.7366f, 41.5493f, 92.3655f, 70.7299f, 92.2041f};

__device__ void estimate_norm_dev(const float* __restrict__ lm, int image_size, double* M) {
  double sx[5], sy[5], dx[5], dy[5];
  const float ratio = (float)((double)image_size / 112.0);
  double msx = 0, msy = 0, mdx = 0, mdy = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    sx[i] = (double)lm[2 * i], sy[i] = (double)lm[2 * i + 1];
    const float tx = image_size == 112 ? c_template[2 * i] : __fmul_rn(ratio, c_template[2 * i]);
    const float ty = image_size == 112 ? c_template[2 * i + 1] : __fmul_rn(ratio, c_template[2 * i + 1]);
    dx[i] = (double)tx, dy[i] = (double)ty;
    msx += sx[i], msy += sy[i], mdx += dx[i], mdy += dy[i];
  }
  msx /= 5.0, msy /= 5.0, mdx /= 5.0, mdy /= 5.0;
  double a = 0, bq = 0, v = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const double px = sx[i] - msx, py = sy[i] - msy, qx = dx[i] - mdx, qy = dy[i] - mdy;
    a += px * qx + py * qy;
    bq += px * qy - py * qx;
    v += px * px + py * py;
  }
  const double pp = a / v, qq = bq / v;
  M[0] = pp, M[1] = -qq, M[2] = mdx - (pp * msx - qq * msy);
  M[3] = qq, M[4] = pp, M[5] = mdy - (qq * msx + pp * msy);
}

__global__ void estimate_norm_kernel(const float* __restrict__ lm, int faces, int image_size, double* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < faces) estimate_norm_dev(lm + (size_t)f * 10, image_size, out + (size_t)f * 6);
}

struct WarpParams {
  const uint8_t* frames;
  int h, w;
  const int* frame_idx;
  const float* landmarks;  // used when m == nullptr
  const double* m;
  int faces, size;
  float mean, scale;
  void* out_nhwc;
  int c_pad, is_bf16;
  uint8_t* crop_u8;
  double* m_out;
};


// Four bilinear taps (2 x 2 pixels x BGR) of cv2.warpAffine's fixed-point interpolation.  Interior pixels read the six
// contiguous bytes of each source row through two aligned 8-byte loads instead of six byte loads (the gather is
// bound by LSU wavefronts); pixels on the frame border take the guarded byte path (out-of-image taps add 0).
__device__ __forceinline__ void warp_taps(const uint8_t* __restrict__ img, int h, int w, int sx, int sy, int w00, int w01,
                                          int w10, int w11, int (&bgr)[3]) {
  const bool interior = sx >= 0 && sx + 1 < w && sy >= 0 && sy + 1 < h;
  const uint8_t* r0 = img + ((size_t)sy * w + sx) * 3;
  const uint8_t* r1 = r0 + (size_t)w * 3;
  const uint8_t* end = img + (size_t)h * w * 3;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(r0) & ~(uintptr_t)7, a1 = reinterpret_cast<uintptr_t>(r1) & ~(uintptr_t)7;
  if (interior && reinterpret_cast<const uint8_t*>(a1) + 16 <= end) {
    const uint2 p0 = __ldg(reinterpret_cast<const uint2*>(a0)), p1 = __ldg(reinterpret_cast<const uint2*>(a0) + 1);
    const uint2 q0 = __ldg(reinterpret_cast<const uint2*>(a1)), q1 = __ldg(reinterpret_cast<const uint2*>(a1) + 1);
    const int s0 = (int)(reinterpret_cast<uintptr_t>(r0) & 7) * 8, s1 = (int)(reinterpret_cast<uintptr_t>(r1) & 7) * 8;
    const unsigned long long lo0 = ((unsigned long long)p0.y << 32) | p0.x, hi0 = ((unsigned long long)p1.y << 32) | p1.x;
    const unsigned long long lo1 = ((unsigned long long)q0.y << 32) | q0.x, hi1 = ((unsigned long long)q1.y << 32) | q1.x;
    const unsigned long long t0 = s0 ? (lo0 >> s0) | (hi0 << (64 - s0)) : lo0;      // six bytes: b g r b g r
    const unsigned long long t1 = s1 ? (lo1 >> s1) | (hi1 << (64 - s1)) : lo1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int acc = (int)((t0 >> (8 * c)) & 0xFF) * w00 + (int)((t0 >> (8 * (c + 3))) & 0xFF) * w01 +
                      (int)((t1 >> (8 * c)) & 0xFF) * w10 + (int)((t1 >> (8 * (c + 3))) & 0xFF) * w11;
      bgr[c] = (acc + 16384) >> 15;
    }
    return;
  }
  const bool x0ok = sx >= 0 && sx < w, x1ok = sx + 1 >= 0 && sx + 1 < w;
  const bool y0ok = sy >= 0 && sy < h, y1ok = sy + 1 >= 0 && sy + 1 < h;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int acc = 0;
    if (y0ok && x0ok) acc += r0[c] * w00;
    if (y0ok && x1ok) acc += r0[3 + c] * w01;
    if (y1ok && x0ok) acc += r1[c] * w10;
    if (y1ok && x1ok) acc += r1[3 + c] * w11;
    bgr[c] = (acc + 16384) >> 15;
  }
}

__global__ void __launch_bounds__(256) warp_affine_kernel(WarpParams p) {
  const int f = blockIdx.y;
  __shared__ double sM[6];
  if (threadIdx.x == 0) {
    if (p.m) {
      for (int i = 0; i < 6; ++i) sM[i] = p.m[(size_t)f * 6 + i];
    } else {
      estimate_norm_dev(p.landmarks + (size_t)f * 10, p.size, sM);
    }
    if (p.m_out && blockIdx.x == 0)
      for (int i = 0; i < 6; ++i) p.m_out[(size_t)f * 6 + i] = sM[i];
  }
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.size * p.size) return;
  const int x = idx % p.size, y = idx / p.size;
  // invert the forward transform exactly as cv2.warpAffine does (float64, no fused multiply-add)
  double D = __dsub_rn(__dmul_rn(sM[0], sM[4]), __dmul_rn(sM[1], sM[3]));
  D = D != 0.0 ? 1.0 / D : 0.0;
  const double i00 = __dmul_rn(sM[4], D), i11 = __dmul_rn(sM[0], D);
  const double i01 = __dmul_rn(sM[1], -D), i10 = __dmul_rn(sM[3], -D);
  const double i02 = __dsub_rn(__dmul_rn(-i00, sM[2]), __dmul_rn(i01, sM[5]));
  const double i12 = __dsub_rn(__dmul_rn(-i10, sM[2]), __dmul_rn(i11, sM[5]));
  const int adelta = (int)__double2ll_rn(__dmul_rn(__dmul_rn(i00, (double)x), 1024.0));
  const int bdelta = (int)__double2ll_rn(__dmul_rn(__dmul_rn(i10, (double)x), 1024.0));
  const int X0 = (int)__double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(i01, (double)y), i02), 1024.0)) + 16;
  const int Y0 = (int)__double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(i11, (double)y), i12), 1024.0)) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
  const int fx = X & 31, fy = Y & 31;
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  const uint8_t* img = p.frames + (size_t)p.frame_idx[f] * p.h * p.w * 3;
  int bgr[3];
  warp_taps(img, p.h, p.w, sx, sy, w00, w01, w10, w11, bgr);
  const size_t opix = (size_t)f * p.size * p.size + idx;
  if (p.crop_u8) {
    uint8_t* o = p.crop_u8 + opix * 3;
    o[0] = (uint8_t)bgr[0], o[1] = (uint8_t)bgr[1], o[2] = (uint8_t)bgr[2];
  }
  if (p.out_nhwc) {
    uint16_t h[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float v = c < 3 ? __fmul_rn(__fsub_rn((float)bgr[2 - c], p.mean), p.scale) : 0.f;
      if (p.is_bf16) {
        __nv_bfloat16 t = __float2bfloat16_rn(v);
        h[c] = *reinterpret_cast<uint16_t*>(&t);
      } else {
        __half t = __float2half_rn(v);
        h[c] = *reinterpret_cast<uint16_t*>(&t);
      }
    }
    uint16_t* o = reinterpret_cast<uint16_t*>(p.out_nhwc) + opix * p.c_pad;
    *reinterpret_cast<uint2*>(o) = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
    for (int c = 4; c < p.c_pad; c += 4) *reinterpret_cast<uint2*>(o + c) = make_uint2(0u, 0u);
  }
}


// norm_crop + normalise + 3x3 / pad 1 / stride 1 patch extraction: one CTA per face.  The aligned crop is built once in
// shared memory (uint8, exactly the cv2.warpAffine arithmetic of warp_affine_kernel), then every thread emits
// 16-byte pieces of the [size][size][32] patch tensor the first ArcFace convolution consumes as a 1x1 conv.
constexpr int kPatchCrop = 112;
template <bool IMG8>   // IMG8: emit the crop itself as 16-byte pixels (R, G, B, 0 x 5) for the stem-form convolution
__global__ void __launch_bounds__(512, 2) warp_patches_kernel(WarpParams p, uint16_t* __restrict__ out) {
  __shared__ uint8_t crop[kPatchCrop * kPatchCrop * 3];
  __shared__ uint16_t lut[256];
  __shared__ double sM[6];
  const int f = blockIdx.x, size = p.size;
  if (threadIdx.x == 0) estimate_norm_dev(p.landmarks + (size_t)f * 10, size, sM);
  if (threadIdx.x < 256) lut[threadIdx.x] = norm16((int)threadIdx.x, p.mean, p.scale, p.is_bf16);
  __syncthreads();
  // inverse transform, float64 without fused multiply-add (cv2.warpAffine)
  double D = __dsub_rn(__dmul_rn(sM[0], sM[4]), __dmul_rn(sM[1], sM[3]));
  D = D != 0.0 ? 1.0 / D : 0.0;
  const double i00 = __dmul_rn(sM[4], D), i11 = __dmul_rn(sM[0], D);
  const double i01 = __dmul_rn(sM[1], -D), i10 = __dmul_rn(sM[3], -D);
  const double i02 = __dsub_rn(__dmul_rn(-i00, sM[2]), __dmul_rn(i01, sM[5]));
  const double i12 = __dsub_rn(__dmul_rn(-i10, sM[2]), __dmul_rn(i11, sM[5]));
  const uint8_t* img = p.frames + (size_t)p.frame_idx[f] * p.h * p.w * 3;
#pragma unroll 2
  for (int idx = threadIdx.x; idx < size * size; idx += blockDim.x) {
    const int x = idx % size, y = idx / size;
    const int adelta = (int)__double2ll_rn(__dmul_rn(__dmul_rn(i00, (double)x), 1024.0));
    const int bdelta = (int)__double2ll_rn(__dmul_rn(__dmul_rn(i10, (double)x), 1024.0));
    const int X0 = (int)__double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(i01, (double)y), i02), 1024.0)) + 16;
    const int Y0 = (int)__double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(i11, (double)y), i12), 1024.0)) + 16;
    const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
    const int fx = X & 31, fy = Y & 31;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    int bgr[3];
    warp_taps(img, p.h, p.w, sx, sy, w00, w01, w10, w11, bgr);
    crop[idx * 3 + 0] = (uint8_t)bgr[0], crop[idx * 3 + 1] = (uint8_t)bgr[1], crop[idx * 3 + 2] = (uint8_t)bgr[2];
  }
  __syncthreads();
  if (IMG8) {
    uint4* o8 = reinterpret_cast<uint4*>(out) + (size_t)f * size * size;
    for (int pix = threadIdx.x; pix < size * size; pix += blockDim.x) {
      const uint8_t* c = crop + pix * 3;
      o8[pix] = make_uint4(lut[c[2]] | ((uint32_t)lut[c[1]] << 16), lut[c[0]], 0u, 0u);
    }
    return;
  }
  uint16_t* o = out + (size_t)f * size * size * 32;
  // emission: one thread per crop pixel, values through a 256-entry table (measured faster than recomputing them
  // and than piece-per-lane coalescing)
  for (int pix = threadIdx.x; pix < size * size; pix += blockDim.x) {
    const int x = pix % size, y = pix / size;
    uint16_t v[32];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = y + tap / 3 - 1, ix = x + tap % 3 - 1;
      const bool ok = iy >= 0 && iy < size && ix >= 0 && ix < size;
      const uint8_t* c = crop + (iy * size + ix) * 3;
      v[tap * 3 + 0] = ok ? lut[c[2]] : (uint16_t)0;       // R
      v[tap * 3 + 1] = ok ? lut[c[1]] : (uint16_t)0;       // G
      v[tap * 3 + 2] = ok ? lut[c[0]] : (uint16_t)0;       // B
    }
#pragma unroll
    for (int k = 27; k < 32; ++k) v[k] = 0;
    uint4* q = reinterpret_cast<uint4*>(o + (size_t)pix * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      q[j] = make_uint4(v[8 * j] | ((uint32_t)v[8 * j + 1] << 16), v[8 * j + 2] | ((uint32_t)v[8 * j + 3] << 16),
                        v[8 * j + 4] | ((uint32_t)v[8 * j + 5] << 16), v[8 * j + 6] | ((uint32_t)v[8 * j + 7] << 16));
  }
}

}  // namespace b2f

using namespace b2f;

extern "C" int b2f_letterbox_u8(const uint8_t* frames, int batch, int h, int w, int new_w, int new_h, int in_w,
                                int in_h, uint8_t* out, void* stream) {
  ResizeGeom g;
  int rc = make_geom(&g, h, w, new_w, new_h, in_w, in_h);
  if (rc) return rc;
  const long long total = (long long)batch * in_h * in_w;
  letterbox_kernel<0><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(frames, g, batch, 0.f, 1.f, out, 3, 0);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2f_preprocess(const uint8_t* frames, int batch, int h, int w, int new_w, int new_h, int in_w, int in_h,
                              float mean, float scale, void* out_nhwc, int c_pad, int dtype, void* stream) {
  B2F_REQUIRE(c_pad >= 4 && c_pad % 4 == 0, "b2f_preprocess: c_pad must be a multiple of 4 (got %d)", c_pad);
  B2F_REQUIRE(dtype == B2F_F16 || dtype == B2F_BF16, "b2f_preprocess: dtype must be f16 or bf16");
  ResizeGeom g;
  int rc = make_geom(&g, h, w, new_w, new_h, in_w, in_h);
  if (rc) return rc;
  const long long total = (long long)batch * in_h * in_w;
  letterbox_kernel<1><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(frames, g, batch, mean, scale, out_nhwc,
                                                                            c_pad, dtype == B2F_BF16);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}


extern "C" int b2f_preprocess_patches(const uint8_t* frames, int batch, int h, int w, int new_w, int new_h, int in_w,
                                      int in_h, int stride, float mean, float scale, void* out_patches, int dtype,
                                      void* stream) {
  B2F_REQUIRE(dtype == B2F_F16 || dtype == B2F_BF16, "b2f_preprocess_patches: dtype must be f16 or bf16");
  B2F_REQUIRE(stride == 1 || stride == 2, "b2f_preprocess_patches: stride must be 1 or 2");
  ResizeGeom g;
  int rc = make_geom(&g, h, w, new_w, new_h, in_w, in_h);
  if (rc) return rc;
  const int ho = (in_h + 2 - 3) / stride + 1, wo = (in_w + 2 - 3) / stride + 1;
  if (batch <= 0) return 0;
  B2F_REQUIRE(batch <= 65535, "b2f_preprocess_patches: batch too large");
  letterbox_patches_kernel<<<dim3((wo + kLpW - 1) / kLpW, (ho + kLpH - 1) / kLpH, batch), kLpW * kLpH, 0,
                             (cudaStream_t)stream>>>(frames, g, stride, ho, wo, mean, scale,
                                                     reinterpret_cast<uint16_t*>(out_patches), dtype == B2F_BF16);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2f_preprocess_conv1(const uint8_t* frames, int batch, int h, int w, int new_w, int new_h, int in_w, int in_h,
                                    float mean, float scale, const void* weight, const float* bias, int cout_p, int act,
                                    void* out, int dtype, void* stream) {
  B2F_REQUIRE(dtype == B2F_F16 || dtype == B2F_BF16, "b2f_preprocess_conv1: dtype must be f16 or bf16");
  B2F_REQUIRE(cout_p == 16 || cout_p == 32, "b2f_preprocess_conv1: cout_p must be 16 or 32 (got %d)", cout_p);
  B2F_REQUIRE(act == B2F_ACT_NONE || act == B2F_ACT_RELU, "b2f_preprocess_conv1: activation must be none or ReLU");
  B2F_REQUIRE(weight && bias && out, "b2f_preprocess_conv1: null argument");
  ResizeGeom g;
  int rc = make_geom(&g, h, w, new_w, new_h, in_w, in_h);
  if (rc) return rc;
  const int ho = (in_h + 2 - 3) / 2 + 1, wo = (in_w + 2 - 3) / 2 + 1;
  if (batch <= 0) return 0;
  // persistent CTAs, four per SM (29 KB of shared memory, 64 TMEM columns and <= 64 registers per thread each)
  const int tiles_x = (wo + kLpW - 1) / kLpW, tiles_y = (ho + kLpH - 1) / kLpH;
  const long long n_tiles_ll = (long long)tiles_x * tiles_y * batch;
  B2F_REQUIRE(n_tiles_ll < (1LL << 31), "b2f_preprocess_conv1: too many tiles");
  const int n_tiles = (int)n_tiles_ll;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    B2F_CHECK_CUDA(cudaGetDevice(&dev));
    B2F_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = n_tiles < sms * 4 ? n_tiles : sms * 4;
  const uint16_t* wt = reinterpret_cast<const uint16_t*>(weight);
  uint16_t* o = reinterpret_cast<uint16_t*>(out);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B2F_BF16) {
    if (cout_p == 32) letterbox_conv1_kernel<true, 32><<<grid, kLpW * kLpH, 0, st>>>(frames, g, ho, wo, mean, scale, wt, bias, act, o, tiles_x, tiles_y, n_tiles);
    else letterbox_conv1_kernel<true, 16><<<grid, kLpW * kLpH, 0, st>>>(frames, g, ho, wo, mean, scale, wt, bias, act, o, tiles_x, tiles_y, n_tiles);
  } else {
    if (cout_p == 32) letterbox_conv1_kernel<false, 32><<<grid, kLpW * kLpH, 0, st>>>(frames, g, ho, wo, mean, scale, wt, bias, act, o, tiles_x, tiles_y, n_tiles);
    else letterbox_conv1_kernel<false, 16><<<grid, kLpW * kLpH, 0, st>>>(frames, g, ho, wo, mean, scale, wt, bias, act, o, tiles_x, tiles_y, n_tiles);
  }
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2f_blob_nchw_f32(const uint8_t* images, int batch, int h, int w, float mean, float scale, float* out,
                                 void* stream) {
  const long long total = (long long)batch * 3 * h * w;
  blob_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(images, batch, h, w, mean, scale, out);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

static int cap_pow2(int n) {
  int p = 32;
  while (p < n) p <<= 1;
  return p;
}

extern "C" long long b2f_decode_nms_workspace(int batch, int max_cand) {
  const long long cap2 = cap_pow2(max_cand);
  const long long per = cap2 * 8 * 2 + (long long)max_cand * 16 + (long long)max_cand * 8;
  return ((per + 255) & ~255LL) * batch;
}

extern "C" int b2f_decode_nms(const b2f_det_levels* lv, int batch, int in_h, int in_w, const float* det_scale,
                              const int* image_hw, float conf_thres, float iou_thres, int max_num, int metric,
                              int max_cand, int max_det, float* det, float* kps, int* keep_idx, int* counts,
                              void* workspace, long long workspace_bytes, void* stream) {
  B2F_REQUIRE(lv && det_scale && det && kps && counts && workspace, "b2f_decode_nms: null argument");
  B2F_REQUIRE(in_h % 32 == 0 && in_w % 32 == 0, "b2f_decode_nms: input size must be a multiple of 32");
  const int total = (in_h / 8) * (in_w / 8) * 2 + (in_h / 16) * (in_w / 16) * 2 + (in_h / 32) * (in_w / 32) * 2;
  if (max_cand > total) max_cand = total;
  B2F_REQUIRE(max_cand >= 1 && max_cand <= kAliveWords * 32, "b2f_decode_nms: max_cand %d out of range", max_cand);
  B2F_REQUIRE(max_det >= 1, "b2f_decode_nms: max_det must be positive");
  B2F_REQUIRE(workspace_bytes >= b2f_decode_nms_workspace(batch, max_cand), "b2f_decode_nms: workspace too small");
  DecodeParams p;
  p.lv = *lv;
  p.batch = batch, p.in_h = in_h, p.in_w = in_w;
  p.det_scale = det_scale, p.image_hw = image_hw;
  p.conf = conf_thres, p.iou = iou_thres;
  p.forward_mode = iou_thres < 0.f ? 1 : 0;   // negative IoU threshold selects the SCRFD.forward() view
  p.max_num = p.forward_mode ? 0 : max_num, p.metric = metric, p.max_cand = max_cand, p.cap2 = cap_pow2(max_cand);
  p.max_det = max_det;
  p.det = det, p.kps = kps, p.keep_idx = keep_idx, p.counts = counts;
  p.ws = reinterpret_cast<uint8_t*>(workspace);
  p.ws_per_frame = b2f_decode_nms_workspace(1, max_cand);
  decode_nms_kernel<<<batch, 1024, kSmemKeys * 8, (cudaStream_t)stream>>>(p);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}


extern "C" int b2f_nms(const float* dets, int n, float iou_thres, int* keep, int* n_keep, void* workspace,
                       long long workspace_bytes, void* stream) {
  B2F_REQUIRE(n >= 0 && n <= kAliveWords * 32, "b2f_nms: at most %d boxes (got %d)", kAliveWords * 32, n);
  const int P = cap_pow2(n);
  B2F_REQUIRE(workspace_bytes >= (long long)P * 8, "b2f_nms: workspace must hold %d keys", P);
  nms_kernel<<<1, 1024, kSmemKeys * 8, (cudaStream_t)stream>>>(dets, n, iou_thres,
                                                              reinterpret_cast<uint64_t*>(workspace), P, keep, n_keep);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2f_distance2bbox(const float* points, const float* distance, int n, float* out, void* stream) {
  if (n <= 0) return 0;
  distance2bbox_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, distance, n, out);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}
extern "C" int b2f_distance2kps(const float* points, const float* distance, int n, int k2, float* out, void* stream) {
  if (n <= 0) return 0;
  B2F_REQUIRE(k2 > 0 && k2 % 2 == 0, "b2f_distance2kps: distance must have an even number of columns");
  distance2kps_kernel<<<(n * k2 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, distance, n, k2, out);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2f_estimate_norm(const float* landmarks, int faces, int image_size, double* m_out, void* stream) {
  if (faces <= 0) return 0;
  estimate_norm_kernel<<<(faces + 127) / 128, 128, 0, (cudaStream_t)stream>>>(landmarks, faces, image_size, m_out);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

static int launch_warp(WarpParams& p, void* stream) {
  if (p.faces <= 0) return 0;
  B2F_REQUIRE(p.faces <= 65535, "norm_crop: at most 65535 faces per call (got %d)", p.faces);
  dim3 grid((p.size * p.size + 255) / 256, p.faces);
  warp_affine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2f_warp_affine_u8(const uint8_t* frames, int h, int w, const int* frame_idx, const double* m, int faces,
                                  int size, uint8_t* out, void* stream) {
  WarpParams p;
  memset(&p, 0, sizeof(p));
  p.frames = frames, p.h = h, p.w = w, p.frame_idx = frame_idx, p.m = m, p.faces = faces, p.size = size;
  p.crop_u8 = out;
  return launch_warp(p, stream);
}

extern "C" int b2f_norm_crop(const uint8_t* frames, int h, int w, const int* frame_idx, const float* landmarks,
                             int faces, int size, float mean, float scale, void* out_nhwc, int c_pad, int dtype,
                             uint8_t* crop_u8, double* m_out, void* stream) {
  B2F_REQUIRE(out_nhwc == nullptr || (c_pad >= 4 && c_pad % 4 == 0), "b2f_norm_crop: c_pad must be a multiple of 4");
  WarpParams p;
  memset(&p, 0, sizeof(p));
  p.frames = frames, p.h = h, p.w = w, p.frame_idx = frame_idx, p.landmarks = landmarks, p.faces = faces;
  p.size = size, p.mean = mean, p.scale = scale, p.out_nhwc = out_nhwc, p.c_pad = c_pad, p.is_bf16 = dtype == B2F_BF16;
  p.crop_u8 = crop_u8, p.m_out = m_out;
  return launch_warp(p, stream);
}

extern "C" int b2f_norm_crop_patches(const uint8_t* frames, int h, int w, const int* frame_idx, const float* landmarks,
                                     int faces, int size, float mean, float scale, void* out_patches, int dtype,
                                     void* stream) {
  B2F_REQUIRE(size == kPatchCrop, "b2f_norm_crop_patches: crop size must be %d (got %d)", kPatchCrop, size);
  B2F_REQUIRE(dtype == B2F_F16 || dtype == B2F_BF16, "b2f_norm_crop_patches: dtype must be f16 or bf16");
  if (faces <= 0) return 0;
  WarpParams p;
  memset(&p, 0, sizeof(p));
  p.frames = frames, p.h = h, p.w = w, p.frame_idx = frame_idx, p.landmarks = landmarks, p.faces = faces;
  p.size = size, p.mean = mean, p.scale = scale, p.is_bf16 = dtype == B2F_BF16;
  warp_patches_kernel<false><<<faces, 512, 0, (cudaStream_t)stream>>>(p, reinterpret_cast<uint16_t*>(out_patches));
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

extern "C" int b2f_norm_crop_image8(const uint8_t* frames, int h, int w, const int* frame_idx, const float* landmarks,
                                    int faces, int size, float mean, float scale, void* out_image8, int dtype, void* stream) {
  B2F_REQUIRE(size == kPatchCrop, "b2f_norm_crop_image8: crop size must be %d (got %d)", kPatchCrop, size);
  B2F_REQUIRE(dtype == B2F_F16 || dtype == B2F_BF16, "b2f_norm_crop_image8: dtype must be f16 or bf16");
  if (faces <= 0) return 0;
  WarpParams p;
  memset(&p, 0, sizeof(p));
  p.frames = frames, p.h = h, p.w = w, p.frame_idx = frame_idx, p.landmarks = landmarks, p.faces = faces;
  p.size = size, p.mean = mean, p.scale = scale, p.is_bf16 = dtype == B2F_BF16;
  warp_patches_kernel<true><<<faces, 512, 0, (cudaStream_t)stream>>>(p, reinterpret_cast<uint16_t*>(out_image8));
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}
