// tcgen05 / TMEM / TMA implicit-GEMM kernel for sm_100a.
//
// One warp-specialised kernel serves three callers:
//   * conv2d (3x3 / 1x1 / kxk, stride 1|2) on NHWC fp16|bf16 activations  -- replaces the ONNX Runtime
//     Conv/BN/Relu/PRelu/Add nodes behind reference models/scrfd.py:83 and models/arcface.py:51
//   * fully-connected (a kxk "valid" conv over a kxk map)                 -- the ArcFace Gemm node
//   * cosine top-k matching (1x1 "conv" whose weights are the gallery)    -- reference main.py:136-142,
//     qdrant_manager.py:164-170, duplicate.py:2726-2797
//
// Data path per CTA: TMA (4-D tiled map over the NHWC input, one box per filter tap, OOB zero fill
// = conv padding, elementStrides = conv stride) -> 128B/64B/32B-swizzled smem ring -> tcgen05.mma
// (M=128 output pixels x N<=256 output channels, fp32 accumulators in TMEM) -> tcgen05.ld epilogue
// (bias table / residual / ReLU|PReLU|sigmoid, or running top-k) -> global.
#include "umma_shared.cuh"
#include "../../include/b2f.h"

#include <atomic>
#include <mutex>

namespace b2f {

extern std::atomic<long long> g_launches;

// ------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------
constexpr int kMaxStages = 8;
constexpr int kATileBytes = 128 * 128;  // 128 rows x 128 B (largest swizzle span)
constexpr int kTopKMax = 8;

enum { EPI_STORE = 0, EPI_TOPK = 1, EPI_PAIRS = 2 };

struct UmmaParams {
  // problem geometry (output space)
  int N, Ho, Wo;            // batch, output height/width
  int H, W;                 // input height/width (border classes)
  int cout_p;               // padded output channels (row pitch of `out`)
  int kh, kw, stride, pad;
  int cchunks;              // Cin_p / kchunk
  int kchunk;               // K elements per pipeline stage: 64 / 32 / 16
  int tw, th, tn;           // M-tile geometry: tw*th*tn <= 128 output pixels
  int tiles_x, tiles_y;     // tiles along Wo, Ho
  int block_n;              // UMMA N
  int n_tiles;              // ceil(cout / block_n)
  int n_per_cta;            // N tiles looped by one CTA (grid.y = ceil(n_tiles / n_per_cta))
  int stages, b_tile_bytes, tmem_cols;
  int total_tiles, n_acc, acc_stride;   // persistent kernel: tiles = M tiles x N tiles, TMEM accumulator ring
  int a_stage_bytes, stages_a, stages_b, b_resident;   // vertical-halo kernel
  int a_res;                // umma_conv_kernel: the M tile stays resident, only N tiles stream
  int halo2;                // 1: one (th+2) x (tw+2) box per channel chunk serves all nine taps (tw == 8)
  int is_bf16;
  int debug;                // bottleneck isolation (b2f_set_tuning key 4): 1 no stores, 2 no residual, 4 no MMA, 8 no A loads, 16 no epilogue math
  // EPI_STORE
  void* out;
  int out_dtype;            // 0 f16, 1 bf16, 2 f32
  const float* bias;        // [bias_classes][cout_p]
  int bias_classes;         // 1 or 9 (3x3 border classes: which taps fall inside the image)
  const float* slope;       // PReLU slope [cout_p] (act == 2)
  int act;                  // 0 none, 1 relu, 2 prelu, 3 sigmoid
  int sig_hi;               // sigmoid only on channels [0, sig_hi); 0 = all
  const void* residual;     // same dtype as activations
  int res_mode;             // 0 none, 1 same-size, 2 nearest 2x upsample of a (res_h,res_w) map
  int res_round;            // in-place block output without activation: round the conv result to 16 bits before the add,
                            // which is what conv_tile_kernel's TMA reduce-store computes (same bits from either kernel)
  int res_h, res_w;
  // EPI_TOPK
  int topk;                 // <= kTopKMax
  long long n_valid;        // gallery rows
  const float* row_scale;   // per query (M) multiplier or null
  const float* col_scale;   // per gallery row (N) multiplier or null
  float* part_score;        // [Q][gridDim.y][2 column halves][topk]
  int* part_idx;
  // EPI_PAIRS (rows row_begin.. of the same matrix against all rows; only j > i is reported)
  int row_begin;
  int diag_skip;            // 1: skip column tiles that lie entirely at or below the diagonal
  float thr_coarse, thr_exact;
  const float* exact_rows;  // unit fp32 rows for the exact re-check (may be null)
  int exact_dim;
  long long* pairs;
  long long max_pairs;
  unsigned long long* pair_count;
};

// ------------------------------------------------------------------------------------------
// the kernel: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue
// ------------------------------------------------------------------------------------------
constexpr int kV1Threads = 64 + 2 * 128;   // producer warp, MMA warp, two epilogue groups of four warps

template <int EPI>
__global__ void __launch_bounds__(kV1Threads, 1)
umma_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);

  // a_res: the CTA's M tile (all K chunks of it) stays resident in shared memory while the CTA walks its N tiles --
  // the matching / clustering GEMMs re-use one 128 x K query tile against thousands of gallery tiles, and re-streaming
  // it per N tile made them L2->SM bound; only the gallery tiles go through the ring then
  const int a_res_bytes = p.a_res ? p.kh * p.kw * p.cchunks * kATileBytes : 0;
  const int stage_bytes = (p.a_res ? 0 : kATileBytes) + p.b_tile_bytes;
  uint8_t* ring = smem + a_res_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;
  uint64_t* ares_bar = bars + 2 * kMaxStages + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates: one M tile per CTA, a contiguous range of N tiles looped with two TMEM accumulators
  const int m_tile = blockIdx.x;
  const int tx = m_tile % p.tiles_x;
  const int ty = (m_tile / p.tiles_x) % p.tiles_y;
  const int tz = m_tile / (p.tiles_x * p.tiles_y);
  const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tz * p.tn;
  int nt_begin = blockIdx.y * p.n_per_cta;
  const int nt_end = min(nt_begin + p.n_per_cta, p.n_tiles);
  if (EPI == EPI_PAIRS && p.diag_skip) nt_begin = max(nt_begin, (p.row_begin + n0) / p.block_n);
  const int k_iters = p.kh * p.kw * p.cchunks;
  const uint32_t row_bytes = p.kchunk * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);
    }
    mbar_init(ares_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      const int kh = p.kh, kw = p.kw, cchunks = p.cchunks, kchunk = p.kchunk, block_n = p.block_n, stages = p.stages;
      const int ax = x0 * p.stride - p.pad, ay = y0 * p.stride - p.pad;
      const bool a_res = p.a_res != 0;
      const uint32_t a_bytes = (uint32_t)(p.tw * p.th * p.tn) * row_bytes;
      const uint32_t tx_bytes = (a_res ? 0u : a_bytes) + (uint32_t)block_n * row_bytes;
      if (a_res && nt_begin < nt_end) {
        mbar_arrive_expect_tx(ares_bar, a_bytes * (uint32_t)(kh * kw * cchunks));
        uint8_t* dst = smem;
        for (int r = 0; r < kh; ++r)
          for (int sx = 0; sx < kw; ++sx)
            for (int cc = 0; cc < cchunks; ++cc, dst += kATileBytes)
              tma_load_4d(dst, &tmA, ares_bar, cc * kchunk, ax + sx, ay + r, n0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int nt = nt_begin; nt < nt_end; ++nt) {
        int tap = 0;
        for (int r = 0; r < kh; ++r) {
          for (int sx = 0; sx < kw; ++sx, ++tap) {
            for (int cc = 0; cc < cchunks; ++cc) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* a_dst = ring + stage * stage_bytes;
              mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
              if (!a_res) tma_load_4d(a_dst, &tmA, &full_bar[stage], cc * kchunk, ax + sx, ay + r, n0);
              tma_load_3d(a_dst + (a_res ? 0 : kATileBytes), &tmB, &full_bar[stage], cc * kchunk, nt * block_n, tap);
              if (++stage == stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const uint32_t idesc = umma_idesc(128, (uint32_t)p.block_n, (uint32_t)p.is_bf16);
    const int ksteps = p.kchunk >> 4, stages = p.stages;
    const uint32_t block_n = (uint32_t)p.block_n;
    const bool a_res = p.a_res != 0;
    const uint64_t a_desc0 = umma_smem_desc(smem_u32(a_res ? smem : ring), row_bytes);
    const uint64_t b_desc0 = umma_smem_desc(smem_u32(ring) + (a_res ? 0 : kATileBytes), row_bytes);
    const uint64_t stage_inc = (uint64_t)(stage_bytes >> 4), a_res_inc = (uint64_t)(kATileBytes >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (a_res && nt_begin < nt_end) {
      mbar_wait(ares_bar, 0);
      tc_fence_after();
    }
    for (int nt = nt_begin; nt < nt_end; ++nt, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * block_n;
      for (int kk = 0; kk < k_iters; ++kk) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t off = stage_inc * (uint64_t)stage;
          issue_stage_rt(ksteps, d_tmem, a_desc0 + (a_res ? a_res_inc * (uint64_t)kk : off), b_desc0 + off, idesc, kk != 0);
          umma_commit(&empty_bar[stage]);           // frees the smem slot when these MMAs retire
          if (kk == k_iters - 1) umma_commit(&tfull_bar[acc]);
        }
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ================================ epilogue: two groups of four warps split the columns ================
    const int group = (warp - 2) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may touch
    const int m = q * 32 + lane;                  // accumulator row == output pixel / query within the tile
    const int lx = m % p.tw;
    const int ly = (m / p.tw) % p.th;
    const int lz = m / (p.tw * p.th);
    const int ox = x0 + lx, oy = y0 + ly, on = n0 + lz;
    const bool valid = (lz < p.tn) && ox < p.Wo && oy < p.Ho && on < p.N;
    const int half = ((p.block_n >> 1) + 15) & ~15;
    const int c_begin = group == 0 ? 0 : half;
    const int c_end = group == 0 ? half : p.block_n;

    if (EPI == EPI_STORE) {
      int cls = 0;
      if (p.bias_classes == 9) {
        const int iy = oy * p.stride - p.pad, ix = ox * p.stride - p.pad;
        const int cy = iy < 0 ? 0 : (iy + p.kh - 1 >= p.H ? 2 : 1);
        const int cx = ix < 0 ? 0 : (ix + p.kw - 1 >= p.W ? 2 : 1);
        cls = cy * 3 + cx;
      }
      const float* bias_row = p.bias + (size_t)cls * p.cout_p;
      const size_t pix = ((size_t)on * p.Ho + oy) * p.Wo + ox;
      const int esz = p.out_dtype == 2 ? 4 : 2;
      size_t res_pix = pix;
      if (p.res_mode == 2) {
        const int ry = min(oy >> 1, p.res_h - 1), rx = min(ox >> 1, p.res_w - 1);
        res_pix = ((size_t)on * p.res_h + ry) * p.res_w + rx;
      }
      int it = 0;
      for (int nt = nt_begin; nt < nt_end; ++nt, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n);
        const int cbase = nt * p.block_n;
        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
          uint32_t r[16];
          __syncwarp();                               // the TMEM load is warp-collective: reconverge first
          tmem_ld16(t_addr + (uint32_t)c0, r);
          tmem_ld_wait();
          const int c = cbase + c0;
          if (valid && c < p.cout_p) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(r[i]) + __ldg(bias_row + c + i);
            if (p.res_mode) {
              float rs[16];
              load16_as_float(reinterpret_cast<const uint8_t*>(p.residual) + (res_pix * p.cout_p + c) * 2, p.is_bf16,
                              rs);
              if (p.res_round) {
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = round16(f[i], p.is_bf16);
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += rs[i];
            }
            if (p.act) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (p.act != 3 || p.sig_hi == 0 || c + i < p.sig_hi)
                  f[i] = act_apply(f[i], p.act, p.act == 2 ? __ldg(p.slope + c + i) : 0.f);
            }
            store16(reinterpret_cast<uint8_t*>(p.out) + (pix * p.cout_p + c) * esz, p.out_dtype, f);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
    } else if (EPI == EPI_PAIRS) {
      const long long gi = (long long)p.row_begin + on;   // global row of this thread
      int it = 0;
      for (int nt = nt_begin; nt < nt_end; ++nt, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n);
        const long long cbase = (long long)nt * p.block_n;
        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
          uint32_t r[16];
          __syncwarp();                               // the TMEM load is warp-collective: reconverge first
          tmem_ld16(t_addr + (uint32_t)c0, r);
          tmem_ld_wait();
          float mx = __uint_as_float(r[0]);
#pragma unroll
          for (int i = 1; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
          if (!(valid && mx >= p.thr_coarse)) continue;           // nothing in this chunk can pass
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const long long gj = cbase + c0 + i;
            const float v = __uint_as_float(r[i]);
            if (gj > gi && gj < p.n_valid && v >= p.thr_coarse) {
              bool hit = true;
              if (p.exact_rows) {
                const float* a = p.exact_rows + (size_t)gi * p.exact_dim;
                const float* b = p.exact_rows + (size_t)gj * p.exact_dim;
                float d = 0.f;
                for (int k = 0; k < p.exact_dim; ++k) d = fmaf(a[k], b[k], d);
                hit = d >= p.thr_exact;
              }
              if (hit) {
                const unsigned long long slot = atomicAdd(p.pair_count, 1ull);
                if ((long long)slot < p.max_pairs) p.pairs[slot] = (gi << 32) | gj;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
    } else {
      // running top-k over this CTA's slice of gallery rows; one query row per thread and column half
      float best_s[kTopKMax];
      int best_i[kTopKMax];
#pragma unroll
      for (int t = 0; t < kTopKMax; ++t) {
        best_s[t] = -INFINITY;
        best_i[t] = -1;
      }
      const bool scaled = p.row_scale != nullptr || p.col_scale != nullptr;
      const float rsc = (valid && p.row_scale) ? __ldg(p.row_scale + on) : 1.f;
      float worst = -INFINITY;   // current k-th best: most chunks fail one max test and skip the insertion
      int it = 0;
      for (int nt = nt_begin; nt < nt_end; ++nt, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n);
        const long long cbase = (long long)nt * p.block_n;
        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
          uint32_t r[16];
          __syncwarp();                               // the TMEM load is warp-collective: reconverge first
          tmem_ld16(t_addr + (uint32_t)c0, r);
          tmem_ld_wait();
          const long long g0 = cbase + c0;
          const bool full_chunk = g0 + 16 <= p.n_valid;
          if (full_chunk && !scaled) {
            float mx = __uint_as_float(r[0]);
#pragma unroll
            for (int i = 1; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
            if (!(mx > worst)) continue;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const long long g = g0 + i;
            if (full_chunk || g < p.n_valid) {
              float v = __uint_as_float(r[i]);
              if (scaled) {
                v *= rsc;
                if (p.col_scale) v *= __ldg(p.col_scale + g);
              }
              if (v > worst) {
                // insertion into the descending list; strict '>' keeps the lowest index among equals
                float cs = v;
                int ci = (int)g;
#pragma unroll
                for (int t = 0; t < kTopKMax; ++t) {
                  if (t < p.topk && cs > best_s[t]) {
                    const float ts = best_s[t];
                    const int ti = best_i[t];
                    best_s[t] = cs;
                    best_i[t] = ci;
                    cs = ts;
                    ci = ti;
                  }
                }
#pragma unroll
                for (int t = 0; t < kTopKMax; ++t)
                  if (t == p.topk - 1) worst = best_s[t];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
      if (valid) {
        const size_t o = (((size_t)on * gridDim.y + blockIdx.y) * 2 + group) * p.topk;
#pragma unroll
        for (int t = 0; t < kTopKMax; ++t) {
          if (t < p.topk) {
            p.part_score[o + t] = best_s[t];
            p.part_idx[o + t] = best_i[t];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// persistent convolution kernel (EPI_STORE only): one CTA per SM loops over output tiles.
//   warp 0 = TMA producer (never drains between tiles), warp 1 = MMA issuer over a ring of TMEM
//   accumulators, warps 2..9 = two epilogue groups that alternate tiles, so the epilogue of tile i
//   overlaps the MMAs of tile i+1 and barrier / TMEM set-up is paid once per SM instead of per tile.
// ------------------------------------------------------------------------------------------
constexpr int kMaxAcc = 4;
constexpr int kEpiGroups = 4;
constexpr int kPersistThreads = 64 + kEpiGroups * 128;   // producer warp, MMA warp, 16 epilogue warps
constexpr int kEpiSmemBytes = 24 * 1024;                 // bias table [<=9][cout_p] + slopes [cout_p], cout_p <= 512

__device__ __forceinline__ void epilogue_store16(const UmmaParams& p, const uint32_t* r, const float* bias_row,
                                                 const float* s_slope, int c, size_t pix, size_t res_pix, int esz) {
  float f[16];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const float4 b4 = *(reinterpret_cast<const float4*>(bias_row + c) + v);     // shared memory
    f[4 * v + 0] = __uint_as_float(r[4 * v + 0]) + b4.x;
    f[4 * v + 1] = __uint_as_float(r[4 * v + 1]) + b4.y;
    f[4 * v + 2] = __uint_as_float(r[4 * v + 2]) + b4.z;
    f[4 * v + 3] = __uint_as_float(r[4 * v + 3]) + b4.w;
  }
  if (p.res_mode && !(p.debug & 2)) {
    float rs[16];
    load16_as_float(reinterpret_cast<const uint8_t*>(p.residual) + (res_pix * p.cout_p + c) * 2, p.is_bf16, rs);
    if (p.res_round) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = round16(f[i], p.is_bf16);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] += rs[i];
  }
  if (p.act == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
  } else if (p.act == 2) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float4 s4 = *(reinterpret_cast<const float4*>(s_slope + c) + v);        // shared memory
      f[4 * v + 0] = f[4 * v + 0] >= 0.f ? f[4 * v + 0] : f[4 * v + 0] * s4.x;
      f[4 * v + 1] = f[4 * v + 1] >= 0.f ? f[4 * v + 1] : f[4 * v + 1] * s4.y;
      f[4 * v + 2] = f[4 * v + 2] >= 0.f ? f[4 * v + 2] : f[4 * v + 2] * s4.z;
      f[4 * v + 3] = f[4 * v + 3] >= 0.f ? f[4 * v + 3] : f[4 * v + 3] * s4.w;
    }
  } else if (p.act == 3) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (p.sig_hi == 0 || c + i < p.sig_hi) f[i] = 1.f / (1.f + expf(-f[i]));
  }
  if ((p.debug & 1) && f[0] != 12345.678f) return;
  store16(reinterpret_cast<uint8_t*>(p.out) + (pix * p.cout_p + c) * esz, p.out_dtype, f);
}


// epilogue of the persistent kernels: four groups of four warps.  With four TMEM accumulators each group owns
// every fourth tile; with two (N > 128) a pair of groups shares a tile and splits its columns.  Bias table and
// PReLU slopes are staged in shared memory once per CTA.  Each thread owns one output pixel (TMEM lane).
__device__ __forceinline__ void persistent_epilogue(const UmmaParams& p, uint32_t tmem_base, uint64_t* tfull_bar,
                                                    uint64_t* tempty_bar, const float* s_bias, const float* s_slope,
                                                    int warp, int lane) {
  const int tiles_xy = p.tiles_x * p.tiles_y;
  const int group = (warp - 2) >> 2;                 // 0..3
  const int split = kEpiGroups / p.n_acc;            // groups per tile: 1 or 2
  const int my_acc_slot = group / split;             // which tile residue (it % n_acc) this group serves
  const int col_part = group % split;
  const int q = warp & 3;
  const int m = q * 32 + lane;
  const int lx = m % p.tw;
  const int ly = (m / p.tw) % p.th;
  const int lz = m / (p.tw * p.th);
  const int esz = p.out_dtype == 2 ? 4 : 2;
  int it = 0;
  for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
    const int acc = it % p.n_acc;
    if (acc != my_acc_slot) continue;
    const int nt = t % p.n_tiles;
    const int m_tile = t / p.n_tiles;
    const int ox = (m_tile % p.tiles_x) * p.tw + lx;
    const int oy = ((m_tile / p.tiles_x) % p.tiles_y) * p.th + ly;
    const int on = (m_tile / tiles_xy) * p.tn + lz;
    const bool valid = (lz < p.tn) && ox < p.Wo && oy < p.Ho && on < p.N;
    int cls = 0;
    if (p.bias_classes == 9) {
      const int iy = oy * p.stride - p.pad, ix = ox * p.stride - p.pad;
      const int cy = iy < 0 ? 0 : (iy + p.kh - 1 >= p.H ? 2 : 1);
      const int cx = ix < 0 ? 0 : (ix + p.kw - 1 >= p.W ? 2 : 1);
      cls = cy * 3 + cx;
    }
    const float* bias_row = s_bias + (size_t)cls * p.cout_p;
    const size_t pix = ((size_t)on * p.Ho + oy) * p.Wo + ox;
    size_t res_pix = pix;
    if (p.res_mode == 2) {
      const int ry = min(oy >> 1, p.res_h - 1), rx = min(ox >> 1, p.res_w - 1);
      res_pix = ((size_t)on * p.res_h + ry) * p.res_w + rx;
    }
    const uint32_t acc_phase = (uint32_t)(it / p.n_acc) & 1u;
    mbar_wait(&tfull_bar[acc], acc_phase);
    tc_fence_after();
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
    const int cbase = nt * p.block_n;
    const int c_begin = col_part * 128;                                   // split == 2 only when block_n > 128
    const int c_end = split == 1 ? p.block_n : min(p.block_n, c_begin + 128);
    for (int c0 = c_begin; c0 < c_end && !(p.debug & 16); c0 += 32) {
      uint32_t r[32];
      tmem_ld32(t_addr + (uint32_t)c0, r);
      tmem_ld_wait();
      if (valid) {
        epilogue_store16(p, r, bias_row, s_slope, cbase + c0, pix, res_pix, esz);
        if (c0 + 16 < c_end) epilogue_store16(p, r + 16, bias_row, s_slope, cbase + c0 + 16, pix, res_pix, esz);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
  }
}

__global__ void __launch_bounds__(kPersistThreads, 1)
umma_conv_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);

  const int stage_bytes = kATileBytes + p.b_tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;
  uint64_t* tempty_bar = bars + 2 * kMaxStages + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2 * kMaxAcc);
  uint8_t* epi_smem = reinterpret_cast<uint8_t*>(bars) + 512;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k_iters = p.kh * p.kw * p.cchunks;
  const uint32_t row_bytes = p.kchunk * 2;
  const int tiles_xy = p.tiles_x * p.tiles_y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kMaxAcc; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4 * (kEpiGroups / p.n_acc));
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // bias table and PReLU slopes -> shared memory (read by every epilogue thread for every tile)
  float* s_bias = reinterpret_cast<float*>(epi_smem);
  float* s_slope = s_bias + p.bias_classes * p.cout_p;
  for (int i = threadIdx.x; i < p.bias_classes * p.cout_p; i += blockDim.x) s_bias[i] = p.bias[i];
  for (int i = threadIdx.x; i < p.cout_p; i += blockDim.x) s_slope[i] = p.act == 2 ? p.slope[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      // every parameter the loop needs lives in a register: the single issuing thread is a critical path
      const int tw = p.tw, th = p.th, tn = p.tn, tiles_x = p.tiles_x, tiles_y = p.tiles_y, n_tiles = p.n_tiles;
      const int kh = p.kh, kw = p.kw, cchunks = p.cchunks, kchunk = p.kchunk, stride = p.stride, pad = p.pad;
      const int block_n = p.block_n, stages = p.stages, total = p.total_tiles;
      const uint32_t tx_bytes = (uint32_t)(tw * th * tn) * row_bytes + (uint32_t)block_n * row_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int nt = t % n_tiles;
        const int m_tile = t / n_tiles;
        const int x0 = (m_tile % tiles_x) * tw * stride - pad;
        const int y0 = ((m_tile / tiles_x) % tiles_y) * th * stride - pad;
        const int n0 = (m_tile / tiles_xy) * tn;
        const int nrow = nt * block_n;
        int tap = 0;
        for (int r = 0; r < kh; ++r) {
          for (int sx = 0; sx < kw; ++sx, ++tap) {
            for (int cc = 0; cc < cchunks; ++cc) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* a_dst = smem + stage * stage_bytes;
              if (p.debug & 8) {
                mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)block_n * row_bytes);
              } else {
                mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                tma_load_4d(a_dst, &tmA, &full_bar[stage], cc * kchunk, x0 + sx, y0 + r, n0);
              }
              tma_load_3d(a_dst + kATileBytes, &tmB, &full_bar[stage], cc * kchunk, nrow, tap);
              if (++stage == stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const uint32_t idesc = umma_idesc(128, (uint32_t)p.block_n, (uint32_t)p.is_bf16);
    const int ksteps = p.kchunk >> 4, stages = p.stages, total = p.total_tiles, n_acc = p.n_acc;
    const uint32_t acc_stride = (uint32_t)p.acc_stride;
    const uint64_t a_desc0 = umma_smem_desc(smem_u32(smem), row_bytes);
    const uint64_t b_desc0 = umma_smem_desc(smem_u32(smem) + kATileBytes, row_bytes);
    const uint64_t stage_inc = (uint64_t)(stage_bytes >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
      for (int kk = 0; kk < k_iters; ++kk) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t off = stage_inc * (uint64_t)stage;
          if (!(p.debug & 4)) issue_stage_rt(ksteps, d_tmem, a_desc0 + off, b_desc0 + off, idesc, kk != 0);
          umma_commit(&empty_bar[stage]);
          if (kk == k_iters - 1) umma_commit(&tfull_bar[acc]);
        }
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == n_acc) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    persistent_epilogue(p, tmem_base, tfull_bar, tempty_bar, s_bias, s_slope, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// persistent 3x3 / stride-1 kernel with vertical-halo reuse of the A operand and optional resident weights.
//   For every input-channel chunk and horizontal tap s, ONE box of (th+2) x tw pixels is fetched; the three
//   vertical taps r read it at row offsets r*tw (tw % 8 == 0 keeps every 8-row swizzle atom aligned, so the
//   UMMA descriptor only changes its start address).  A traffic drops from 9 to 3*(th+2)/th tile loads.
//   When all 9*cchunks weight tiles fit in shared memory they are loaded once per CTA (b_resident).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPersistThreads, 1)
umma_conv_vhalo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);

  uint8_t* a_ring = smem;
  uint8_t* b_base = smem + p.stages_a * p.a_stage_bytes;
  const int b_slots = p.b_resident ? 9 * p.cchunks : p.stages_b;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)b_slots * p.b_tile_bytes);
  uint64_t* fullA = bars;
  uint64_t* emptyA = bars + kMaxStages;
  uint64_t* fullB = bars + 2 * kMaxStages;
  uint64_t* emptyB = bars + 3 * kMaxStages;
  uint64_t* tfull_bar = bars + 4 * kMaxStages;
  uint64_t* tempty_bar = bars + 4 * kMaxStages + kMaxAcc;
  uint64_t* bres_bar = bars + 4 * kMaxStages + 2 * kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);
  uint8_t* epi_smem = reinterpret_cast<uint8_t*>(bars) + 512;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t row_bytes = p.kchunk * 2;
  const int tiles_xy = p.tiles_x * p.tiles_y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&fullA[s], 1);
      mbar_init(&emptyA[s], 1);
      mbar_init(&fullB[s], 1);
      mbar_init(&emptyB[s], 1);
    }
    for (int s = 0; s < kMaxAcc; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4 * (kEpiGroups / p.n_acc));
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // bias table and PReLU slopes -> shared memory (read by every epilogue thread for every tile)
  float* s_bias = reinterpret_cast<float*>(epi_smem);
  float* s_slope = s_bias + p.bias_classes * p.cout_p;
  for (int i = threadIdx.x; i < p.bias_classes * p.cout_p; i += blockDim.x) s_bias[i] = p.bias[i];
  for (int i = threadIdx.x; i < p.cout_p; i += blockDim.x) s_slope[i] = p.act == 2 ? p.slope[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      const int tw = p.tw, th = p.th, tiles_x = p.tiles_x, tiles_y = p.tiles_y, n_tiles = p.n_tiles;
      const int cchunks = p.cchunks, kchunk = p.kchunk, block_n = p.block_n, total = p.total_tiles;
      const int stages_a = p.stages_a, stages_b = p.stages_b, a_stage_bytes = p.a_stage_bytes;
      const int b_tile_bytes = p.b_tile_bytes;
      const bool resident = p.b_resident != 0;
      const uint32_t a_bytes = (uint32_t)((th + 2) * (p.halo2 ? tw + 2 : tw)) * row_bytes;
      const uint32_t b_bytes = (uint32_t)block_n * row_bytes;
      if (resident) {
        mbar_arrive_expect_tx(bres_bar, b_bytes * 9u * (uint32_t)cchunks);
        uint8_t* dst = b_base;
        for (int cc = 0; cc < cchunks; ++cc)
          for (int sx = 0; sx < 3; ++sx)
            for (int r = 0; r < 3; ++r, dst += b_tile_bytes)
              tma_load_3d(dst, &tmB, bres_bar, cc * kchunk, 0, r * 3 + sx);
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      const int n_boxes = p.halo2 ? 1 : 3, sx_per_box = p.halo2 ? 3 : 1;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int nt = t % n_tiles;
        const int m_tile = t / n_tiles;
        const int x0 = (m_tile % tiles_x) * tw - 1;
        const int y0 = ((m_tile / tiles_x) % tiles_y) * th - 1;
        const int n0 = m_tile / tiles_xy;
        const int nrow = nt * block_n;
        for (int cc = 0; cc < cchunks; ++cc) {
          for (int bx = 0; bx < n_boxes; ++bx) {
            mbar_wait(&emptyA[sa], pa ^ 1);
            if (p.debug & 8) {
              mbar_arrive(&fullA[sa]);
            } else {
              mbar_arrive_expect_tx(&fullA[sa], a_bytes);
              tma_load_4d(a_ring + (size_t)sa * a_stage_bytes, &tmA, &fullA[sa], cc * kchunk, x0 + bx, y0, n0);
            }
            if (++sa == stages_a) sa = 0, pa ^= 1;
            if (!resident) {
              for (int sxi = 0; sxi < sx_per_box; ++sxi) {
                for (int r = 0; r < 3; ++r) {
                  mbar_wait(&emptyB[sb], pb ^ 1);
                  mbar_arrive_expect_tx(&fullB[sb], b_bytes);
                  tma_load_3d(b_base + (size_t)sb * b_tile_bytes, &tmB, &fullB[sb], cc * kchunk, nrow, r * 3 + bx + sxi);
                  if (++sb == stages_b) sb = 0, pb ^= 1;
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const uint32_t idesc = umma_idesc(128, (uint32_t)p.block_n, (uint32_t)p.is_bf16);
    const int ksteps = p.kchunk >> 4, total = p.total_tiles, n_acc = p.n_acc;
    const int groups = (p.halo2 ? 1 : 3) * p.cchunks, sx_per_box = p.halo2 ? 3 : 1;
    const int stages_a = p.stages_a, stages_b = p.stages_b;
    const bool resident = p.b_resident != 0;
    const uint32_t acc_stride = (uint32_t)p.acc_stride;
    const int box_w = p.halo2 ? p.tw + 2 : p.tw;
    // 8-row groups of the A tile are 8 consecutive pixels of one box row: group stride = one box row when the
    // box is wider than the tile (halo2), else the canonical 8 rows
    const uint32_t sbo = p.halo2 ? (uint32_t)box_w * row_bytes : 8u * row_bytes;
    const uint64_t a_desc0 = umma_smem_desc_sbo(smem_u32(a_ring), row_bytes, sbo);
    const uint64_t b_desc0 = umma_smem_desc(smem_u32(b_base), row_bytes);
    const uint64_t a_inc = (uint64_t)(p.a_stage_bytes >> 4), b_inc = (uint64_t)(p.b_tile_bytes >> 4);
    const uint64_t r_inc = (uint64_t)(((uint32_t)box_w * row_bytes) >> 4);       // one row of the halo box
    const uint64_t s_inc = (uint64_t)(row_bytes >> 4);                             // one pixel
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (resident) {
      mbar_wait(bres_bar, 0);
      tc_fence_after();
    }
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
      for (int g = 0; g < groups; ++g) {                 // one A box: (channel chunk, horizontal tap) or a whole chunk
        mbar_wait(&fullA[sa], pa);
        tc_fence_after();
        const uint64_t ad = a_desc0 + a_inc * (uint64_t)sa;
        if (resident) {
          if (elect_one()) {
            if (!(p.debug & 4)) {
              uint64_t bd = b_desc0 + b_inc * (uint64_t)(3 * sx_per_box * g);
              for (int sxi = 0; sxi < sx_per_box; ++sxi, bd += 3 * b_inc) {
                const uint64_t as = ad + s_inc * (uint64_t)sxi;
                issue_stage_rt(ksteps, d_tmem, as, bd, idesc, (g | sxi) != 0);
                issue_stage_rt(ksteps, d_tmem, as + r_inc, bd + b_inc, idesc, 1u);
                issue_stage_rt(ksteps, d_tmem, as + 2 * r_inc, bd + 2 * b_inc, idesc, 1u);
              }
            }
            umma_commit(&emptyA[sa]);
            if (g == groups - 1) umma_commit(&tfull_bar[acc]);
          }
        } else {
          for (int sxi = 0; sxi < sx_per_box; ++sxi) {
            for (int r = 0; r < 3; ++r) {
              mbar_wait(&fullB[sb], pb);
              tc_fence_after();
              if (elect_one()) {
                if (!(p.debug & 4))
                  issue_stage_rt(ksteps, d_tmem, ad + s_inc * (uint64_t)sxi + r_inc * (uint64_t)r,
                                 b_desc0 + b_inc * (uint64_t)sb, idesc, (g | sxi | r) != 0);
                umma_commit(&emptyB[sb]);
                if (r == 2 && sxi == sx_per_box - 1) {
                  umma_commit(&emptyA[sa]);
                  if (g == groups - 1) umma_commit(&tfull_bar[acc]);
                }
              }
              if (++sb == stages_b) sb = 0, pb ^= 1;
            }
          }
        }
        if (++sa == stages_a) sa = 0, pa ^= 1;
      }
      if (++acc == n_acc) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    persistent_epilogue(p, tmem_base, tfull_bar, tempty_bar, s_bias, s_slope, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// host-side planning
// ------------------------------------------------------------------------------------------
static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

extern int g_tile_reduce;   // conv_tile.cu
void pick_m_tile(int N, int Ho, int Wo, int stride, int* tw, int* th, int* tn) {
  long long best = -1;
  int bw = 1, bh = 1, bn = 1;
  const int max_w = Wo < 128 ? Wo : 128;
  for (int w = 1; w <= max_w; ++w) {
    if (w * stride > 256) break;
    const int max_h = (128 / w) < Ho ? (128 / w) : Ho;
    for (int h = 1; h <= max_h; ++h) {
      if (h * stride > 256) break;
      int n = 128 / (w * h);
      if (n > N) n = N;
      if (n < 1) continue;
      const long long tiles = (long long)((Wo + w - 1) / w) * ((Ho + h - 1) / h) * ((N + n - 1) / n);
      // fewer tiles first; then wider rows (longer contiguous runs in NHWC)
      if (best < 0 || tiles < best || (tiles == best && w > bw)) {
        best = tiles;
        bw = w, bh = h, bn = n;
      }
    }
  }
  *tw = bw, *th = bh, *tn = bn;
}

int g_smem_budget_single = 100 * 1024;  // lets two CTAs share an SM when each owns one N tile
int g_smem_budget_loop = 200 * 1024;
int g_max_block_n = 256;
int g_a_res = 1;

template <int EPI>
static int launch_umma(const CUtensorMap& tmA, const CUtensorMap& tmB, UmmaParams& p, int m_tiles, int grid_y,
                       cudaStream_t stream) {
  const int k_iters = p.kh * p.kw * p.cchunks;
  const int budget = p.n_per_cta > 1 ? g_smem_budget_loop : g_smem_budget_single;
  // resident M tile when a CTA walks many N tiles and the tile plus three weight stages fit
  const int a_res_bytes = k_iters * kATileBytes;
  p.a_res = (g_a_res && p.n_per_cta >= 4 && a_res_bytes + 3 * p.b_tile_bytes + 2048 <= 227 * 1024) ? 1 : 0;
  const int stage_bytes = (p.a_res ? 0 : kATileBytes) + p.b_tile_bytes;
  int stages = p.a_res ? (227 * 1024 - 2048 - a_res_bytes) / stage_bytes : (budget - 2048) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  if (stages > k_iters * p.n_per_cta && k_iters * p.n_per_cta >= 2) stages = k_iters * p.n_per_cta;
  p.stages = stages;
  const size_t smem = (size_t)(p.a_res ? a_res_bytes : 0) + (size_t)stages * stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] {
    attr_rc = cudaFuncSetAttribute(umma_conv_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  B2F_CHECK_CUDA(attr_rc);
  B2F_REQUIRE(smem <= 227 * 1024, "umma kernel: %zu bytes of shared memory requested", smem);
  umma_conv_kernel<EPI><<<dim3(m_tiles, grid_y), kV1Threads, smem, stream>>>(tmA, tmB, p);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

int g_persistent = 2;     // 0: one CTA per tile, 1: first persistent kernels, 2: conv_tile_kernel (conv_tile.cu)
int conv_tile_launch(const b2f_conv_desc* d, int kchunk, cudaStream_t stream, int optional);
int g_tile_max_n = 256;
int g_num_sms = 0;

static int launch_persistent(const CUtensorMap& tmA, const CUtensorMap& tmB, UmmaParams& p, int m_tiles,
                             cudaStream_t stream) {
  if (g_num_sms == 0) {
    int dev = 0;
    B2F_CHECK_CUDA(cudaGetDevice(&dev));
    B2F_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int stage_bytes = kATileBytes + p.b_tile_bytes;
  int stages = (226 * 1024 - 1024 - 512 - kEpiSmemBytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  p.total_tiles = m_tiles * p.n_tiles;
  p.acc_stride = (p.block_n + 31) & ~31;
  p.n_acc = (512 / p.acc_stride) >= 4 ? 4 : 2;
  size_t smem = (size_t)stages * stage_bytes + 1024 + 512 + kEpiSmemBytes;
  if (smem < 120 * 1024) smem = 120 * 1024;     // one CTA per SM: it owns all 512 TMEM columns
  B2F_REQUIRE((size_t)(p.bias_classes + 1) * p.cout_p * 4 <= (size_t)kEpiSmemBytes,
              "conv: bias table of %d x %d floats does not fit the epilogue staging area", p.bias_classes + 1, p.cout_p);
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] {
    attr_rc = cudaFuncSetAttribute(umma_conv_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024);
  });
  B2F_CHECK_CUDA(attr_rc);
  const int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  umma_conv_persistent_kernel<<<grid, kPersistThreads, smem, stream>>>(tmA, tmB, p);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

int g_vhalo = 1;
int g_debug = 0;

static int launch_vhalo(const CUtensorMap& tmA, const CUtensorMap& tmB, UmmaParams& p, int m_tiles,
                        cudaStream_t stream) {
  if (g_num_sms == 0) {
    int dev = 0;
    B2F_CHECK_CUDA(cudaGetDevice(&dev));
    B2F_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int row_bytes = p.kchunk * 2;
  p.a_stage_bytes = (((p.th + 2) * (p.halo2 ? p.tw + 2 : p.tw) * row_bytes) + 1023) & ~1023;
  const int budget = 226 * 1024 - 1024 - 512 - kEpiSmemBytes;
  int b_bytes_total;
  if (p.b_resident) {
    b_bytes_total = 9 * p.cchunks * p.b_tile_bytes;
    p.stages_b = 0;
  } else {
    p.stages_b = 6;
    while (p.stages_b > 3 && p.stages_b * p.b_tile_bytes > budget / 2) --p.stages_b;
    b_bytes_total = p.stages_b * p.b_tile_bytes;
  }
  p.stages_a = (budget - b_bytes_total) / p.a_stage_bytes;
  if (p.stages_a > kMaxStages) p.stages_a = kMaxStages;
  B2F_REQUIRE(p.stages_a >= 2, "vhalo conv: not enough shared memory for two A stages");
  p.total_tiles = m_tiles * p.n_tiles;
  p.acc_stride = (p.block_n + 31) & ~31;
  p.n_acc = (512 / p.acc_stride) >= 4 ? 4 : 2;
  size_t smem = (size_t)p.stages_a * p.a_stage_bytes + b_bytes_total + 1024 + 512 + kEpiSmemBytes;
  if (smem < 120 * 1024) smem = 120 * 1024;     // keep one CTA per SM: every CTA owns all 512 TMEM columns
  B2F_REQUIRE((size_t)(p.bias_classes + 1) * p.cout_p * 4 <= (size_t)kEpiSmemBytes,
              "conv: bias table of %d x %d floats does not fit the epilogue staging area", p.bias_classes + 1, p.cout_p);
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] {
    attr_rc = cudaFuncSetAttribute(umma_conv_vhalo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  B2F_CHECK_CUDA(attr_rc);
  B2F_REQUIRE(smem <= 227 * 1024, "vhalo conv: %zu bytes of shared memory requested", smem);
  const int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  umma_conv_vhalo_kernel<<<grid, kPersistThreads, smem, stream>>>(tmA, tmB, p);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

static int pick_kchunk(int cin_p) {
  if (cin_p % 64 == 0) return 64;
  if (cin_p % 32 == 0) return 32;
  return 16;
}

}  // namespace b2f

using namespace b2f;

extern "C" int b2f_conv2d(const b2f_conv_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B2F_REQUIRE(d != nullptr, "b2f_conv2d: null descriptor");
  B2F_REQUIRE((d->cin_p % 16 == 0 || d->cin_p == 8) && d->cout_p % 16 == 0,
              "b2f_conv2d: channels must be padded to 16, or cin_p == 8 for the stem form (got %d, %d)", d->cin_p, d->cout_p);
  B2F_REQUIRE(d->stride == 1 || d->stride == 2, "b2f_conv2d: stride %d unsupported", d->stride);
  B2F_REQUIRE(d->bias != nullptr && d->in != nullptr && d->weight != nullptr && d->out != nullptr,
              "b2f_conv2d: null tensor");
  B2F_REQUIRE(d->bias_classes == 1 || (d->bias_classes == 9 && d->kh <= 3 && d->kw <= 3),
              "b2f_conv2d: bias_classes must be 1 or 9 (3x3)");
  const int Ho = (d->h + 2 * d->pad - d->kh) / d->stride + 1;
  const int Wo = (d->w + 2 * d->pad - d->kw) / d->stride + 1;
  B2F_REQUIRE(Ho == d->ho && Wo == d->wo, "b2f_conv2d: output size mismatch (%dx%d vs %dx%d)", d->ho, d->wo, Ho, Wo);

  if (d->cin_p == 8) {
    // stem form: input [n][h][w][8] (16-byte pixels), weight [10][cout_p][8] with slot = filter tap and slot 9 zero;
    // only conv_tile_kernel has this operand layout (A mode 3)
    B2F_REQUIRE(d->stride == 1 || d->stride == 2, "b2f_conv2d: stride %d unsupported", d->stride);
    return conv_tile_launch(d, 8, stream, 0);
  }
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->n, p.Ho = Ho, p.Wo = Wo, p.H = d->h, p.W = d->w;
  p.cout_p = d->cout_p;
  p.kh = d->kh, p.kw = d->kw, p.stride = d->stride, p.pad = d->pad;
  p.kchunk = d->force_kchunk ? d->force_kchunk : pick_kchunk(d->cin_p);
  B2F_REQUIRE(d->cin_p % p.kchunk == 0, "b2f_conv2d: kchunk %d does not divide cin_p %d", p.kchunk, d->cin_p);
  B2F_REQUIRE(d->act != 2 || d->slope != nullptr, "b2f_conv2d: PReLU needs a slope vector");
  if (g_persistent >= 2) {
    // conv_tile_kernel (halo boxes, shared weight tiles, TMA-store epilogue) unless it declines a wide-tile layer
    // with too few tiles per SM; g_persistent == 3 forces it
    const int rc = conv_tile_launch(d, p.kchunk, stream, g_persistent == 2);
    if (rc != -1000) return rc;
  }
  B2F_REQUIRE(d->sc_in == nullptr, "b2f_conv2d: the fused shortcut needs conv_tile_kernel (tuning key 2 >= 2)");
  p.cchunks = d->cin_p / p.kchunk;
  pick_m_tile(d->n, Ho, Wo, d->stride, &p.tw, &p.th, &p.tn);
  p.tiles_x = (Wo + p.tw - 1) / p.tw;
  p.tiles_y = (Ho + p.th - 1) / p.th;
  const int tiles_z = (d->n + p.tn - 1) / p.tn;
  // N tiling: largest block_n <= 256 that divides cout_p evenly into equal 16-multiples
  int n_tiles = (d->cout_p + g_max_block_n - 1) / g_max_block_n;
  while ((d->cout_p % n_tiles) != 0 || ((d->cout_p / n_tiles) % 16) != 0) ++n_tiles;
  p.n_tiles = n_tiles;
  p.block_n = d->cout_p / n_tiles;
  p.n_per_cta = 1;
  p.b_tile_bytes = ((p.block_n * p.kchunk * 2) + 1023) & ~1023;
  p.tmem_cols = pow2_cols(p.block_n);
  p.is_bf16 = d->dtype == 1;
  p.out = d->out, p.out_dtype = d->out_dtype;
  p.bias = d->bias, p.bias_classes = d->bias_classes;
  p.slope = d->slope, p.act = d->act, p.sig_hi = d->sig_hi;
  B2F_REQUIRE(d->act != 2 || d->slope != nullptr, "b2f_conv2d: PReLU needs a slope vector");
  p.residual = d->residual, p.res_mode = d->residual ? d->res_mode : 0;
  p.res_round = (g_tile_reduce && p.res_mode == 1 && d->residual == d->out && d->act == 0 && d->out_dtype != 2) ? 1 : 0;
  p.res_h = d->res_h, p.res_w = d->res_w;
  p.debug = g_debug;

  // ---- vertical-halo variant: 3x3 / stride 1 / pad 1 on maps wide enough for 8-pixel-aligned tiles ----------
  bool vhalo = false;
  // under the default dispatch only wide tiles arrive here, and they keep the tap-major order of the persistent
  // kernel (conv_tile_kernel's mode 0 uses the same one), so the vertical-halo variant is for g_persistent == 1
  if (g_persistent == 1 && g_vhalo && d->kh == 3 && d->kw == 3 && d->stride == 1 && d->pad == 1 && Wo >= 8 && Ho >= 2) {
    // per-tile time model (cycles): tensor pipe = N/2 per 128xNx16 MMA; L2->SM fabric ~37 B/clk/SM (measured:
    // ~10.4 TB/s over 148 SMs on the 256-channel layers, which is what bounds them)
    const double kFabric = 37.0;
    const double row_b = p.kchunk * 2.0;
    const double mma = 9.0 * p.cchunks * (p.kchunk / 16) * (p.block_n / 2.0);
    const int n_plan = d->n < 128 ? 128 : d->n;      // batch-independent choice (fixes the accumulation order)
    int ptw, pth, ptn;
    pick_m_tile(n_plan, Ho, Wo, d->stride, &ptw, &pth, &ptn);
    const double def_tiles = (double)((Wo + ptw - 1) / ptw) * ((Ho + pth - 1) / pth) * ((n_plan + ptn - 1) / ptn) * p.n_tiles;
    const double def_bytes = 9.0 * p.cchunks * (ptw * pth * ptn + p.block_n) * row_b;
    const double def_time = def_tiles * (mma > def_bytes / kFabric ? mma : def_bytes / kFabric);
    const bool can_res = p.n_tiles == 1 && 9 * p.cchunks * p.b_tile_bytes <= 100 * 1024;
    double best = -1;
    int btw = 0;
    for (int tw = 8; tw <= 128; tw <<= 1) {
      const int th = 128 / tw;
      if (tw > ((Wo + 7) & ~7)) continue;
      const double tiles = (double)((Wo + tw - 1) / tw) * ((Ho + th - 1) / th) * n_plan * p.n_tiles;
      const double bytes = p.cchunks * (3.0 * (th + 2) * tw + (can_res ? 0.0 : 9.0 * p.block_n)) * row_b;
      const double t = tiles * (mma > bytes / kFabric ? mma : bytes / kFabric);
      if (best < 0 || t < best) best = t, btw = tw;
    }
    if (g_vhalo == 2) {          // full-halo variant: 8 x 16 tiles, one (10 x 18) box per channel chunk
      vhalo = true;
      p.halo2 = 1;
      p.tw = 8, p.th = 16, p.tn = 1;
      p.tiles_x = (Wo + p.tw - 1) / p.tw;
      p.tiles_y = (Ho + p.th - 1) / p.th;
      p.b_resident = can_res ? 1 : 0;
    } else if (best >= 0 && best < 0.9 * def_time) {
      vhalo = true;
      p.tw = btw, p.th = 128 / btw, p.tn = 1;
      p.tiles_x = (Wo + p.tw - 1) / p.tw;
      p.tiles_y = (Ho + p.th - 1) / p.th;
      p.b_resident = can_res ? 1 : 0;
    }
  }
  const int tiles_z_final = vhalo ? d->n : tiles_z;

  CUtensorMap tmA, tmB;
  if (vhalo) {
    uint64_t dims[4] = {(uint64_t)d->cin_p, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->n};
    uint64_t str[3] = {(uint64_t)d->cin_p * 2, (uint64_t)d->w * d->cin_p * 2, (uint64_t)d->h * d->w * d->cin_p * 2};
    uint32_t box[4] = {(uint32_t)p.kchunk, (uint32_t)(p.halo2 ? p.tw + 2 : p.tw), (uint32_t)(p.th + 2), 1};
    uint32_t es[4] = {1, 1, 1, 1};
    int rc = make_tmap(&tmA, d->in, 4, dims, str, box, es, p.kchunk * 2, p.is_bf16);
    if (rc) return rc;
  } else {
    uint64_t dims[4] = {(uint64_t)d->cin_p, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->n};
    uint64_t str[3] = {(uint64_t)d->cin_p * 2, (uint64_t)d->w * d->cin_p * 2, (uint64_t)d->h * d->w * d->cin_p * 2};
    uint32_t box[4] = {(uint32_t)p.kchunk, (uint32_t)(p.tw * d->stride), (uint32_t)(p.th * d->stride), (uint32_t)p.tn};
    uint32_t es[4] = {1, (uint32_t)d->stride, (uint32_t)d->stride, 1};
    int rc = make_tmap(&tmA, d->in, 4, dims, str, box, es, p.kchunk * 2, p.is_bf16);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)d->cin_p, (uint64_t)d->cout_p, (uint64_t)(d->kh * d->kw)};
    uint64_t str[2] = {(uint64_t)d->cin_p * 2, (uint64_t)d->cout_p * d->cin_p * 2};
    uint32_t box[3] = {(uint32_t)p.kchunk, (uint32_t)p.block_n, 1};
    uint32_t es[3] = {1, 1, 1};
    int rc = make_tmap(&tmB, d->weight, 3, dims, str, box, es, p.kchunk * 2, p.is_bf16);
    if (rc) return rc;
  }
  const int m_tiles = p.tiles_x * p.tiles_y * tiles_z_final;
  if (vhalo) return launch_vhalo(tmA, tmB, p, m_tiles, stream);
  if (g_persistent) return launch_persistent(tmA, tmB, p, m_tiles, stream);
  return launch_umma<EPI_STORE>(tmA, tmB, p, m_tiles, p.n_tiles, stream);
}

// partial top-k of Q x G cosine scores: queries [Q][D], gallery [G][D] (both fp16 or bf16, K-major)
namespace b2f {
int match_pair_topk(const void* queries, int q, const void* gallery, long long g, int dim, int dtype, int topk, int keep,
                    int n_splits, float* part_score, int* part_idx, cudaStream_t stream, int causal, long long causal_base);
int match_pair_pairs(const void* emb16, int n, int dim, int dtype, int row_begin, int row_end, float thr_coarse, float thr_exact,
                     const float* emb_f32, long long* pairs, long long max_pairs, unsigned long long* pair_count,
                     cudaStream_t stream);
}  // namespace b2f
using b2f::match_pair_pairs;
using b2f::match_pair_topk;

extern "C" int b2f_match_partial(const void* queries, int q, const void* gallery, long long g, int dim, int dtype,
                                 const float* row_scale, const float* col_scale, int topk, int n_splits,
                                 float* part_score, int* part_idx, void* stream_) {
  return b2f_match_partial_keep(queries, q, gallery, g, dim, dtype, row_scale, col_scale, topk, topk, n_splits, part_score,
                                part_idx, stream_);
}

extern "C" int b2f_match_partial_keep(const void* queries, int q, const void* gallery, long long g, int dim, int dtype,
                                      const float* row_scale, const float* col_scale, int topk, int keep, int n_splits,
                                      float* part_score, int* part_idx, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B2F_REQUIRE(keep >= 1 && keep <= topk, "b2f_match_partial_keep: keep must be in [1, topk]");
  B2F_REQUIRE(topk >= 1 && topk <= kTopKMax, "b2f_match_partial: topk must be in [1,%d]", kTopKMax);
  B2F_REQUIRE(dim % 64 == 0, "b2f_match_partial: dim must be a multiple of 64");
  B2F_REQUIRE(q > 0 && g > 0 && n_splits >= 1, "b2f_match_partial: empty problem");
  if (row_scale == nullptr && col_scale == nullptr) {       // more than 128 queries: persistent CTA pairs (match_pair.cu)
    const int rc = match_pair_topk(queries, q, gallery, g, dim, dtype, topk, keep, n_splits, part_score, part_idx, stream, 0, 0);
    if (rc != -1000) return rc;
  }
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.N = q, p.Ho = 1, p.Wo = 1, p.H = 1, p.W = 1;
  p.kh = p.kw = 1, p.stride = 1, p.pad = 0;
  p.kchunk = 64, p.cchunks = dim / 64;
  p.tw = 1, p.th = 1, p.tn = 128;
  p.tiles_x = p.tiles_y = 1;
  p.block_n = 256;
  p.n_tiles = (int)((g + 255) / 256);
  if (n_splits > p.n_tiles) n_splits = p.n_tiles;
  p.n_per_cta = (p.n_tiles + n_splits - 1) / n_splits;
  const int grid_y = (p.n_tiles + p.n_per_cta - 1) / p.n_per_cta;
  B2F_REQUIRE(grid_y == n_splits, "b2f_match_partial: n_splits %d does not tile %d column tiles evenly (use %d)",
              n_splits, p.n_tiles, grid_y);
  p.b_tile_bytes = 256 * 64 * 2;
  p.tmem_cols = p.n_per_cta > 1 ? 512 : 256;
  p.is_bf16 = dtype == 1;
  p.topk = topk, p.n_valid = g;
  p.row_scale = row_scale, p.col_scale = col_scale;
  p.part_score = part_score, p.part_idx = part_idx;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)dim, 1, 1, (uint64_t)q};
    uint64_t str[3] = {(uint64_t)dim * 2, (uint64_t)dim * 2, (uint64_t)dim * 2};
    uint32_t box[4] = {64, 1, 1, 128};
    uint32_t es[4] = {1, 1, 1, 1};
    int rc = make_tmap(&tmA, queries, 4, dims, str, box, es, 128, p.is_bf16);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)dim, (uint64_t)g, 1};
    uint64_t str[2] = {(uint64_t)dim * 2, (uint64_t)g * dim * 2};
    uint32_t box[3] = {64, 256, 1};
    uint32_t es[3] = {1, 1, 1};
    int rc = make_tmap(&tmB, gallery, 3, dims, str, box, es, 128, p.is_bf16);
    if (rc) return rc;
  }
  const int m_tiles = (q + 127) / 128;
  return launch_umma<EPI_TOPK>(tmA, tmB, p, m_tiles, grid_y, stream);
}


// prefix search: query row i (global index causal_base + i) against the gallery rows BEFORE it only -- the online loop's
// "best earlier person" (reference duplicate.py:1853-1855) for a whole batch of rows in one tensor-core pass
extern "C" int b2f_match_partial_causal(const void* queries, int q, const void* gallery, long long g, int dim, int dtype,
                                        int topk, int keep, int n_splits, long long causal_base, float* part_score,
                                        int* part_idx, void* stream_) {
  B2F_REQUIRE(keep >= 1 && keep <= topk && topk <= kTopKMax, "b2f_match_partial_causal: need 1 <= keep <= topk <= %d", kTopKMax);
  B2F_REQUIRE(q > 0 && g > 0 && n_splits >= 1 && dim % 64 == 0, "b2f_match_partial_causal: bad problem size");
  const int rc = match_pair_topk(queries, q, gallery, g, dim, dtype, topk, keep, n_splits, part_score, part_idx,
                                 reinterpret_cast<cudaStream_t>(stream_), 1, causal_base);
  B2F_REQUIRE(rc != -1000, "b2f_match_partial_causal: n_splits %d does not tile the gallery evenly (use b2f_match_plan)", n_splits);
  return rc;
}

// all-pairs cosine >= threshold among unit rows: rows [row_begin,row_end) against every row j > i
extern "C" int b2f_pairs_threshold(const void* emb16, int n, int dim, int dtype, int row_begin, int row_end,
                                   float threshold, const float* emb_f32, long long* pairs, long long max_pairs,
                                   unsigned long long* pair_count, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B2F_REQUIRE(dim % 64 == 0, "b2f_pairs_threshold: dim must be a multiple of 64");
  B2F_REQUIRE(0 <= row_begin && row_begin < row_end && row_end <= n, "b2f_pairs_threshold: bad row range");
  const int rows = row_end - row_begin;
  {                                                           // more than 128 rows: persistent CTA pairs (match_pair.cu)
    // the 16-bit coarse score may sit a little under the exact one: widen the gate, re-check in fp32
    const int rc = match_pair_pairs(emb16, n, dim, dtype, row_begin, row_end, emb_f32 ? threshold - 0.02f : threshold, threshold,
                                    emb_f32, pairs, max_pairs, pair_count, stream);
    if (rc != -1000) return rc;
  }
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.N = rows, p.Ho = 1, p.Wo = 1, p.H = 1, p.W = 1;
  p.kh = p.kw = 1, p.stride = 1, p.pad = 0;
  p.kchunk = 64, p.cchunks = dim / 64;
  p.tw = 1, p.th = 1, p.tn = 128;
  p.tiles_x = p.tiles_y = 1;
  p.block_n = 256;
  p.n_tiles = (n + 255) / 256;
  const int m_tiles = (rows + 127) / 128;
  int n_splits = (296 + m_tiles - 1) / m_tiles;
  if (n_splits > p.n_tiles) n_splits = p.n_tiles;
  if (n_splits < 1) n_splits = 1;
  p.n_per_cta = (p.n_tiles + n_splits - 1) / n_splits;
  const int grid_y = (p.n_tiles + p.n_per_cta - 1) / p.n_per_cta;
  p.b_tile_bytes = 256 * 64 * 2;
  p.tmem_cols = 512;
  p.is_bf16 = dtype == 1;
  p.n_valid = n;
  p.row_begin = row_begin, p.diag_skip = 1;
  // the 16-bit coarse score may sit a little under the exact one: widen the gate, re-check in fp32
  p.thr_coarse = emb_f32 ? threshold - 0.02f : threshold;
  p.thr_exact = threshold;
  p.exact_rows = emb_f32, p.exact_dim = dim;
  p.pairs = pairs, p.max_pairs = max_pairs, p.pair_count = pair_count;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)dim, 1, 1, (uint64_t)rows};
    uint64_t str[3] = {(uint64_t)dim * 2, (uint64_t)dim * 2, (uint64_t)dim * 2};
    uint32_t box[4] = {64, 1, 1, 128};
    uint32_t es[4] = {1, 1, 1, 1};
    int rc = make_tmap(&tmA, reinterpret_cast<const uint8_t*>(emb16) + (size_t)row_begin * dim * 2, 4, dims, str, box,
                       es, 128, p.is_bf16);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)dim, (uint64_t)n, 1};
    uint64_t str[2] = {(uint64_t)dim * 2, (uint64_t)n * dim * 2};
    uint32_t box[3] = {64, 256, 1};
    uint32_t es[3] = {1, 1, 1};
    int rc = make_tmap(&tmB, emb16, 3, dims, str, box, es, 128, p.is_bf16);
    if (rc) return rc;
  }
  return launch_umma<EPI_PAIRS>(tmA, tmB, p, m_tiles, grid_y, stream);
}

extern "C" int b2f_match_splits(long long g, int want) {
  int n_tiles = (int)((g + 255) / 256);
  if (want > n_tiles) want = n_tiles;
  if (want < 1) want = 1;
  int per = (n_tiles + want - 1) / want;
  return (n_tiles + per - 1) / per;
}

// ------------------------------------------------------------------------------------------
// debug: load one TMA box and dump the raw shared-memory image (validates maps / swizzle / strides)
// ------------------------------------------------------------------------------------------
namespace b2f {
__global__ void tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3,
                                 uint32_t bytes, uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bar;
  for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) smem[i] = 0xAB;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, bytes);
    tma_load_4d(smem, &tm, &bar, c0, c1, c2, c3);
  }
  mbar_wait(&bar, 0);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}
}  // namespace b2f

extern "C" int b2f_debug_tma_probe(const void* src, const long long* dims4, const int* box4, const int* estr4,
                                   int swizzle_bytes, const int* coords4, unsigned char* out_dev, int out_bytes,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  uint64_t dims[4], str[3];
  uint32_t box[4], es[4];
  for (int i = 0; i < 4; ++i) dims[i] = (uint64_t)dims4[i], box[i] = (uint32_t)box4[i], es[i] = (uint32_t)estr4[i];
  str[0] = dims[0] * 2, str[1] = str[0] * dims[1], str[2] = str[1] * dims[2];
  CUtensorMap tm;
  int rc = make_tmap(&tm, src, 4, dims, str, box, es, swizzle_bytes, 0);
  if (rc) return rc;
  uint32_t bytes = 2;
  for (int i = 0; i < 4; ++i) bytes *= (uint32_t)((box4[i] + estr4[i] - 1) / estr4[i]);
  B2F_REQUIRE((int)bytes <= out_bytes && bytes <= 64 * 1024, "tma probe: box of %u bytes too large", bytes);
  tma_probe_kernel<<<1, 128, bytes + 1024, stream>>>(tm, coords4[0], coords4[1], coords4[2], coords4[3], bytes, out_dev);
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}
