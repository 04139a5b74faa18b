// Persistent tcgen05 convolution kernel for sm_100a (the conv / FC layers of SCRFD and ArcFace; replaces the
// ONNX Runtime Conv/BN/Relu/PRelu/Add/Gemm nodes behind reference models/scrfd.py:83 and models/arcface.py:51).
//
// One CTA per SM walks a list of work items.  Roles:
//   warp 0      TMA producer (one thread): activation boxes into the A ring, weight tiles into the B ring
//               (or all of them once, when they fit: "resident" weights)
//   warp 1      tcgen05.mma issuer: M = 128 output pixels x N <= 256 output channels, fp32 accumulators in a ring
//               of TMEM accumulators, so the epilogue of one tile overlaps the MMAs of the next ones
//   warps 2..   epilogue groups of four warps (one TMEM lane quarter each), alternating tiles
//
// What bounds these layers on B200 is not the tensor pipe but (a) the L2 -> SM fabric (~40 B/clk/SM) when every
// filter tap re-fetches its activation box and every tile re-fetches the weights, and (b) the LSU when 128
// threads store 16-byte pieces 128+ bytes apart.  Hence:
//   * A operand, three modes: one box per filter tap (any conv); one (th+2) x tw box per horizontal tap (3x3
//     stride 1: the three vertical taps read it at row offsets); one (th+2) x (tw+2) box per channel chunk
//     (tw == 8: all nine taps read it; the UMMA descriptor's 8-row group stride becomes one box row -- the
//     hardware applies the 128B/64B/32B swizzle to absolute shared-memory address bits, so a tap is only a
//     different start address)
//   * B operand: resident, or streamed once per PAIR of M tiles (mt == 2: two accumulators share each weight tile)
//   * epilogue: TMEM -> registers -> bias / residual / activation -> swizzled staging tile in shared memory ->
//     one TMA tensor store per 16/32/64-channel slice (bounds clipping for free); the residual arrives through TMA
//     into the same staging buffers two slices ahead
#include "umma_shared.cuh"
#include "../../include/b2f.h"

#include <atomic>
#include <stdlib.h>

namespace b2f {

extern std::atomic<long long> g_launches;
extern int g_max_block_n;
extern int g_debug;
extern int g_vhalo;
int g_tile_groups = 0;   // 0 = auto, else 2 or 4 epilogue groups
int g_tile_mt = 0;       // 0 = auto, else 1 or 2 M tiles per weight tile
int g_tile_amode = -1;   // -1 = auto, else force A mode 0 / 1 / 2 where legal
int g_tile_epi = -1;     // -1 = auto, 0 = direct global stores, 1 = TMA stores
int g_tile_cg2 = 1;      // CTA pairs (cta_group::2): 0 = never, 1 = tiles wider than g_tile_cg2_min_n, 2 = wherever legal
int g_tile_cg2_min_n = 128;
int g_tile_pdl = 1;      // programmatic dependent launch between consecutive layers
int g_tile_wide_res = 2;
int g_tile_reduce = 1;   // in-place residual layers add through the TMA reduce-store
int g_tile_big_res = 1;  // pairs keep up to 150 KB of weights resident (lean staging ring) // wide tiles on CTA pairs: TMA-fed residual through the staging ring

constexpr int kTStages = 16;
constexpr int kTAcc = 4;
constexpr int kTGroups = 4;
constexpr int kStgBufs = 3;
constexpr int kSmemMax = 227 * 1024;
constexpr int kTileDeclined = -1000;

// Division by a launch-invariant divisor: q = (x * mul) >> shift, exact for 0 <= x < 2^31 (mul = ceil(2^shift / d),
// shift = 31 + ceil(log2 d)).  The producer thread turns a work-item index into tile coordinates once per tile; with
// hardware-emulated `/` and `%` (eight of them, ~200 dependent clocks each) that chain alone took ~1500 clocks per tile
// and bounded every layer whose tile has fewer MMA clocks than that (the 32- and 64-channel layers).
struct FastDiv {
  uint32_t mul, shift, d;
  __host__ void set(int div) {
    d = (uint32_t)div;
    uint32_t s = 0;
    while ((1u << s) < d) ++s;
    shift = 31 + s;
    mul = (uint32_t)((((unsigned long long)1 << shift) + d - 1) / d);
  }
  __device__ __forceinline__ int div(int x) const { return (int)(((unsigned long long)(uint32_t)x * mul) >> shift); }
  __device__ __forceinline__ void divmod(int x, int& q, int& r) const { q = div(x), r = x - q * (int)d; }
};

struct TileParams {
  int N, Ho, Wo, H, W, cout_p;
  int kh, kw, stride, pad;
  int cchunks, kchunk;
  int tw, th, tn, tiles_x, tiles_y, m_tiles;
  int block_n, n_tiles, items, mt;
  int a_mode, boxes_per_chunk, taps_per_box;
  int a_bytes, a_box_bytes, a_stage_bytes, stages_a;
  int b_tile_bytes, stages_b, b_resident;
  int b_stride;             // bytes between weight stages; mode 0 streamed: the stage also holds its activation box(es)
  int combined;             // 1: one barrier pair per tap covers the weight tile and the activation boxes (mode 0, streamed)
  int pdl;                  // launched with programmatic stream serialization
  int sc_cchunks, sc_stride; // fused projection shortcut: extra K chunks from a second tensor (mode 0 only), its stride
  int cg2;                  // 1: CTA pair (cluster of 2, tcgen05 cta_group::2): M = 256 over two SMs, each loads half of B
  int n_acc_log2, acc_stride;
  int groups;
  int pool, pool_h, pool_w, off_pool;   // fused 3x3 / s2 / p1 max-pool: pooled map size, per-group pooled staging
  int res_reduce;   // residual == out and no activation: the epilogue reduce-adds into the output instead of loading it
  int is_bf16, debug;       // debug (b2f_set_tuning key 4) = the bits of umma_conv.cu plus, in the TMA-store epilogue: 64 polling wait
                            // on the accumulator, 128 no proxy fence, 256 no TMEM read, 512 no group barrier (timing experiments only)
  int epi_tma, res_smem, res_global, ochunk, n_sub, stg_bytes, stg_box_bytes, stg_bufs;
  void* out;
  int out_dtype;
  const float* bias;
  int bias_classes;
  const float* slope;
  int act, sig_hi;
  const void* residual;
  int res_mode, res_h, res_w;
  int off_b, off_stg, off_bar, off_tab;
  FastDiv fd_ntiles, fd_tx, fd_ty, fd_mt, fd_ksplit;   // n_tiles, tiles_x, tiles_y, mt, ksplit
  int ksplit;               // split-K: an item covers taps [split * taps / ksplit, +taps / ksplit) and stores fp32 partial sums
  long long split_stride;   // floats between the partial outputs of consecutive splits (p.out is then the workspace)
};

__device__ __forceinline__ void unpack8(const uint4& v, int is_bf16, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (is_bf16) {
      __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      f[2 * i] = __bfloat162float(h.x);
      f[2 * i + 1] = __bfloat162float(h.y);
    } else {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      f[2 * i] = __half2float(h.x);
      f[2 * i + 1] = __half2float(h.y);
    }
  }
}
__device__ __forceinline__ uint4 pack8(const float* f, int is_bf16) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (is_bf16) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    } else {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// bias (+ residual) + activation on 16 accumulator columns; the caller decides where the result goes
__device__ __forceinline__ void epi_math16(const TileParams& p, const uint32_t* r, const float* bias_row,
                                           const float* s_slope, int c, const float* res, float* f) {
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const float4 b4 = *(reinterpret_cast<const float4*>(bias_row + c) + v);     // shared memory, broadcast
    f[4 * v + 0] = __uint_as_float(r[4 * v + 0]) + b4.x;
    f[4 * v + 1] = __uint_as_float(r[4 * v + 1]) + b4.y;
    f[4 * v + 2] = __uint_as_float(r[4 * v + 2]) + b4.z;
    f[4 * v + 3] = __uint_as_float(r[4 * v + 3]) + b4.w;
  }
  if (res) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] += res[i];
  }
  if (p.act == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
  } else if (p.act == 2) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float4 s4 = *(reinterpret_cast<const float4*>(s_slope + c) + v);
      f[4 * v + 0] = fmaxf(f[4 * v + 0], 0.f) + s4.x * fminf(f[4 * v + 0], 0.f);
      f[4 * v + 1] = fmaxf(f[4 * v + 1], 0.f) + s4.y * fminf(f[4 * v + 1], 0.f);
      f[4 * v + 2] = fmaxf(f[4 * v + 2], 0.f) + s4.z * fminf(f[4 * v + 2], 0.f);
      f[4 * v + 3] = fmaxf(f[4 * v + 3], 0.f) + s4.w * fminf(f[4 * v + 3], 0.f);
    }
  } else if (p.act == 3) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (p.sig_hi == 0 || c + i < p.sig_hi) f[i] = 1.f / (1.f + expf(-f[i]));
  }
}

// ---- specialised epilogue arithmetic: straight-line code per (activation, dtype, residual) ------------------
template <bool BF16>
__device__ __forceinline__ void unpack8_t(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (BF16) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    } else {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x, f[2 * i + 1] = t.y;
    }
  }
}
template <bool BF16>
__device__ __forceinline__ uint4 pack8_t(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (BF16) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    } else {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// 16 accumulator columns -> bias (+ residual read from the staging row) -> activation -> packed into the staging row
template <int ACT, bool BF16, int RES>
__device__ __forceinline__ void epi_slice16(const uint32_t* r, const float* bias_c, const float* slope_c, uint4* p0,
                                            uint4* p1, int c, int sig_hi, const uint4* gres) {
  float f[16];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const float4 b4 = *(reinterpret_cast<const float4*>(bias_c) + v);           // shared memory, broadcast
    f[4 * v + 0] = __uint_as_float(r[4 * v + 0]) + b4.x;
    f[4 * v + 1] = __uint_as_float(r[4 * v + 1]) + b4.y;
    f[4 * v + 2] = __uint_as_float(r[4 * v + 2]) + b4.z;
    f[4 * v + 3] = __uint_as_float(r[4 * v + 3]) + b4.w;
  }
  if (RES == 1) {                                   // residual slice was TMA-loaded into this staging row
    float rs[16];
    unpack8_t<BF16>(*p0, rs);
    unpack8_t<BF16>(*p1, rs + 8);
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] += rs[i];
  } else if (RES == 2) {                            // residual straight from global memory (wide tiles)
    if (gres != nullptr) {
      float rs[16];
      unpack8_t<BF16>(__ldg(gres), rs);
      unpack8_t<BF16>(__ldg(gres + 1), rs + 8);
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] += rs[i];
    }
  }
  if (ACT == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
  } else if (ACT == 2) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float4 s4 = *(reinterpret_cast<const float4*>(slope_c) + v);
      f[4 * v + 0] = fmaxf(f[4 * v + 0], 0.f) + s4.x * fminf(f[4 * v + 0], 0.f);
      f[4 * v + 1] = fmaxf(f[4 * v + 1], 0.f) + s4.y * fminf(f[4 * v + 1], 0.f);
      f[4 * v + 2] = fmaxf(f[4 * v + 2], 0.f) + s4.z * fminf(f[4 * v + 2], 0.f);
      f[4 * v + 3] = fmaxf(f[4 * v + 3], 0.f) + s4.w * fminf(f[4 * v + 3], 0.f);
    }
  } else if (ACT == 3) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (sig_hi == 0 || c + i < sig_hi) f[i] = 1.f / (1.f + expf(-f[i]));
  }
  *p0 = pack8_t<BF16>(f);
  *p1 = pack8_t<BF16>(f + 8);
}

struct EpiCtx {
  uint32_t tmem_base;
  uint64_t *tfull, *tempty, *rbar;
  uint8_t* stg;
  const float *s_bias, *s_slope;
  const CUtensorMap *tmO, *tmR;
  int tiles_cta, group, q, lane;
  int first, stride, rank;        // work items of this CTA (or CTA pair): first, first + stride, ...; rank in the pair
  uint64_t* peer_tempty;          // cta_group::2: the pair's second CTA releases accumulators on the leader's barrier
  bool leader;
};

struct TileCoord {
  int x0, y0, n0, cbase, m_tile, split;
};

__device__ __forceinline__ TileCoord tile_of(const TileParams& p, int first, int stride, int rank, int seq) {
  // an item = `pair` consecutive M tiles of one N tile: mt tiles of one CTA, or one tile for each CTA of a pair
  int l, u, i, nt, mp, tx, ty, tyx, tn;
  TileCoord t;
  p.fd_mt.divmod(seq, l, u);
  p.fd_ksplit.divmod(first + l * stride, i, t.split);
  p.fd_ntiles.divmod(i, mp, nt);
  const int m_tile = p.cg2 ? (mp * 2 + rank) * p.mt + u : mp * p.mt + u;
  p.fd_tx.divmod(m_tile, tyx, tx);
  p.fd_ty.divmod(tyx, tn, ty);
  t.m_tile = m_tile;
  t.x0 = tx * p.tw;
  t.y0 = ty * p.th;
  t.n0 = tn * p.tn;
  t.cbase = nt * p.block_n;
  return t;
}

// TMA-store epilogue of one group of four warps: TMEM -> registers -> arithmetic -> swizzled staging slice ->
// cp.async.bulk.tensor store.  Staging buffers form a ring of kStgBufs per group; with a residual, the slice that
// will be processed two steps later is TMA-loaded into its buffer first, and each thread reads / overwrites only
// its own 16-byte pieces of it.
template <int ACT, bool BF16, int RES, bool CG2>
__device__ __forceinline__ void epilogue_tma(const TileParams& p, const EpiCtx& c) {
  // RES: 0 none, 1 residual through the staging ring, 2 residual from global memory, 3 the output buffer already holds
  // the residual (in-place block output, no activation after the add): plain arithmetic, then a TMA reduce-add store
  constexpr int MR = RES == 3 ? 0 : RES;
  const int G = p.groups, n_sub = p.n_sub, och = p.ochunk, group = c.group;
  const int n_acc_mask = (1 << p.n_acc_log2) - 1;
  const uint32_t orow = (uint32_t)och * 2;                           // staging row bytes: 64 / 32
  const uint32_t swz_mask = orow == 64 ? 3u : 1u;
  const int m = c.q * 32 + c.lane;
  const int lx = m % p.tw, ly = (m / p.tw) % p.th, lz = m / (p.tw * p.th);
  const int NB = p.stg_bufs;
  const int my_tiles = c.tiles_cta > group ? (c.tiles_cta - group + G - 1) / G : 0;
  const int total_sub = my_tiles * n_sub;
  const uint32_t row_off = (uint32_t)m * orow;
  // swizzled byte offsets of this thread's four 16-byte pieces (two per 16 channels)
  uint32_t off[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t o = row_off + (uint32_t)i * 16;
    off[i] = o ^ (((o >> 7) & swz_mask) << 4);
  }
  const bool skip_math = (p.debug & 16) != 0, skip_store = (p.debug & 1) != 0;
  const int stg_bytes = p.stg_bytes, sig_hi = p.sig_hi, cout_p = p.cout_p;
  // running (tile, slice, buffer) of the residual prefetch, NB - 1 slices ahead of the arithmetic: no `/` or `%` by
  // launch-time values on the per-slice path (each is ~200 dependent clocks)
  int rs_tile = 0, rs_sub = 0, rs_buf = 0, rs_s = 0;
  auto issue_res = [&]() {                                           // leader only: residual slice rs_s
    const TileCoord t = tile_of(p, c.first, c.stride, c.rank, group + rs_tile * G);
    mbar_arrive_expect_tx(&c.rbar[rs_buf], (uint32_t)p.stg_box_bytes);
    tma_load_4d(c.stg + (size_t)rs_buf * stg_bytes, c.tmR, &c.rbar[rs_buf], t.cbase + rs_sub * och, t.x0, t.y0, t.n0);
    ++rs_s;
    if (++rs_sub == n_sub) rs_sub = 0, ++rs_tile;
    if (++rs_buf == NB) rs_buf = 0;
  };
  if (RES == 1 && c.leader)
    for (int s = 0; s < NB - 1 && s < total_sub; ++s) issue_res();
  int k = 0, buf = 0;
  uint32_t buf_phase = 0;                                            // (k / NB) & 1
  for (int tl = 0; tl < my_tiles; ++tl) {
    const int seq = group + tl * G;
    const TileCoord t = tile_of(p, c.first, c.stride, c.rank, seq);
    int cls = 0;
    if (p.bias_classes == 9) {
      const int ox = t.x0 + lx, oy = t.y0 + ly;
      const int iy = oy * p.stride - p.pad, ix = ox * p.stride - p.pad;
      const int cy = iy < 0 ? 0 : (iy + p.kh - 1 >= p.H ? 2 : 1);
      const int cx = ix < 0 ? 0 : (ix + p.kw - 1 >= p.W ? 2 : 1);
      cls = cy * 3 + cx;
    }
    const float* bias_row = c.s_bias + cls * cout_p + t.cbase;
    const float* slope_row = c.s_slope + t.cbase;
    const uint8_t* gres_row = nullptr;
    if (RES == 2) {
      const int ox = t.x0 + lx, oy = t.y0 + ly, on = t.n0 + lz;
      if (lz < p.tn && ox < p.Wo && oy < p.Ho && on < p.N)
        gres_row = reinterpret_cast<const uint8_t*>(p.residual) + ((((size_t)on * p.Ho + oy) * p.Wo + ox) * cout_p + t.cbase) * 2;
    }
    const int acc = seq & n_acc_mask;
    const uint32_t ph = (uint32_t)(seq >> p.n_acc_log2) & 1u;
    if (p.debug & 64) mbar_wait_poll(&c.tfull[acc], ph); else mbar_wait(&c.tfull[acc], ph);
    tc_fence_after();
    const uint32_t t_addr = c.tmem_base + ((uint32_t)(c.q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
    for (int j = 0; j < n_sub; ++j, ++k) {
      uint8_t* bufp = c.stg + (size_t)buf * stg_bytes;
      uint32_t r[32];
      if (och == 16) {
        uint32_t r16[16];
        tmem_ld16(t_addr + (uint32_t)(j * och), r16);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = r16[i];
      } else if (!(p.debug & 256)) {
        tmem_ld32(t_addr + (uint32_t)(j * och), r);
        tmem_ld_wait();
      }
      if (j == n_sub - 1) {                              // accumulator fully read: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (c.lane == 0) {
          if (CG2 && c.rank) mbar_arrive_remote(&c.peer_tempty[acc], 0);
          else mbar_arrive(&c.tempty[acc]);
        }
      }
      if (RES == 1) mbar_wait(&c.rbar[buf], buf_phase);
      if (!skip_math) {
        const int cl = j * och;
        const uint4* g0 = (RES == 2 && gres_row) ? reinterpret_cast<const uint4*>(gres_row + cl * 2) : nullptr;
        epi_slice16<ACT, BF16, MR>(r, bias_row + cl, slope_row + cl, reinterpret_cast<uint4*>(bufp + off[0]),
                                   reinterpret_cast<uint4*>(bufp + off[1]), t.cbase + cl, sig_hi, g0);
        if (och == 32)
          epi_slice16<ACT, BF16, MR>(r + 16, bias_row + cl + 16, slope_row + cl + 16, reinterpret_cast<uint4*>(bufp + off[2]),
                                      reinterpret_cast<uint4*>(bufp + off[3]), t.cbase + cl + 16, sig_hi, g0 ? g0 + 2 : nullptr);
      }
      if (!(p.debug & 128)) fence_proxy_async();
      if (c.leader) {
        // the store issued one slice ago has had this slice's arithmetic to leave its buffer; once it has, the
        // residual of the slice two steps ahead may land there
        bulk_wait_read0();
        if (RES == 1 && rs_s < total_sub) issue_res();
      }
      if (!(p.debug & 512)) bar_sync_named(1 + group, 128);
      if (c.leader && !skip_store) {
        if (RES == 3) tma_reduce_add_4d(c.tmO, bufp, t.cbase + j * och, t.x0, t.y0, t.n0);
        else tma_store_4d(c.tmO, bufp, t.cbase + j * och, t.x0, t.y0, t.n0);
        bulk_commit();
      }
      if (++buf == NB) buf = 0, buf_phase ^= 1u;
    }
  }
  if (c.leader) bulk_wait_all();
}

// Epilogue with a fused 3x3 / stride 2 / pad 1 max-pool (tiles of 8 x 16 pixels, ReLU, no residual).  The conv result
// never leaves the registers: a warp owns a 8 x 4 strip of the tile (one TMEM lane quarter), a 32-channel slice of a
// pixel is 16 packed 16-bit pairs per thread, and the window maxima are built with warp shuffles -- horizontally over
// lanes +-1, vertically over lanes +-8.  A strip contributes to 3 x 5 pooled pixels (windows that straddle a strip or
// tile edge are partial); the 15 lanes that hold them write 64 bytes each into the warp's own 1 KB buffer and ONE TMA
// reduce-store per warp max-merges them into the zero-initialised pooled map (ReLU outputs are >= 0, so 0 is the
// identity of the max and partial maxima of neighbouring strips / tiles combine in L2).  No staging tile, no block-level
// barrier, no shared-memory reads.
constexpr int kPoolW = 5, kPoolH = 3, kPoolBufBytes = 1024;      // per warp and buffer; two buffers per warp
template <bool BF16>
__device__ __forceinline__ uint32_t hmax2u(uint32_t a, uint32_t b) {
  if (BF16) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <bool BF16>
__device__ __forceinline__ void epilogue_pool(const TileParams& p, const EpiCtx& c, uint8_t* pbuf_warp) {
  const int G = p.groups, n_sub = p.n_sub, group = c.group;
  const int n_acc_mask = (1 << p.n_acc_log2) - 1;
  const int lane = c.lane, lx = lane & 7, lr = lane >> 3;       // pixel of the strip: column lx, row lr of rows 4q .. 4q+3
  const int ly = c.q * 4 + lr;
  const int my_tiles = c.tiles_cta > group ? (c.tiles_cta - group + G - 1) / G : 0;
  const int cout_p = p.cout_p;
  // which pooled pixel of the strip's 3 x 5 this lane ends up holding (if any)
  const int prow = lr == 0 ? 0 : (lr == 2 ? 1 : (lr == 3 ? 2 : -1));
  const int pcol = lx == 7 ? 4 : ((lx & 1) ? -1 : (lx >> 1));
  const bool holder = prow >= 0 && pcol >= 0;
  const uint32_t dst_off = holder ? (uint32_t)((prow * kPoolW + pcol) * 64) : 0u;
  // which neighbours a lane's window takes in (see the loop below): none to the left of column 0, none at all for
  // column 7; the row above only for row 2, the row below for rows 0 and 2
  const uint32_t m_left = (lx > 0 && lx < 7) ? 0xFFFFFFFFu : 0u, m_right = lx < 7 ? 0xFFFFFFFFu : 0u;
  const uint32_t m_up = lr == 2 ? 0xFFFFFFFFu : 0u, m_down = (lr == 0 || lr == 2) ? 0xFFFFFFFFu : 0u;
  int k = 0;
  for (int tl = 0; tl < my_tiles; ++tl) {
    const int seq = group + tl * G;
    const TileCoord t = tile_of(p, c.first, c.stride, c.rank, seq);
    int cls = 0;
    if (p.bias_classes == 9) {
      const int iy = t.y0 + ly - 1, ix = t.x0 + lx - 1;
      const int cy = iy < 0 ? 0 : (iy + 2 >= p.H ? 2 : 1);
      const int cx = ix < 0 ? 0 : (ix + 2 >= p.W ? 2 : 1);
      cls = cy * 3 + cx;
    }
    const float* bias_row = c.s_bias + cls * cout_p + t.cbase;
    const bool in_image = t.y0 + ly < p.Ho && t.x0 + lx < p.Wo;     // a tile may overhang the right / bottom edge
    const int acc = seq & n_acc_mask;
    const uint32_t ph = (uint32_t)(seq >> p.n_acc_log2) & 1u;
    mbar_wait(&c.tfull[acc], ph);
    tc_fence_after();
    const uint32_t t_addr = c.tmem_base + ((uint32_t)(c.q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
    for (int j = 0; j < n_sub; ++j, ++k) {
      uint32_t r[32];
      tmem_ld32(t_addr + (uint32_t)(j * 32), r);
      tmem_ld_wait();
      if (j == n_sub - 1) {                              // accumulator fully read: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&c.tempty[acc]);
      }
      const int cl = j * 32;
      uint32_t v[16];
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const float4 b4 = *(reinterpret_cast<const float4*>(bias_row + cl) + q4);          // shared memory, broadcast
        const float f0 = fmaxf(__uint_as_float(r[4 * q4 + 0]) + b4.x, 0.f), f1 = fmaxf(__uint_as_float(r[4 * q4 + 1]) + b4.y, 0.f);
        const float f2 = fmaxf(__uint_as_float(r[4 * q4 + 2]) + b4.z, 0.f), f3 = fmaxf(__uint_as_float(r[4 * q4 + 3]) + b4.w, 0.f);
        if (BF16) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(f0, f1), h1 = __floats2bfloat162_rn(f2, f3);
          v[2 * q4] = *reinterpret_cast<uint32_t*>(&h0), v[2 * q4 + 1] = *reinterpret_cast<uint32_t*>(&h1);
        } else {
          __half2 h0 = __floats2half2_rn(f0, f1), h1 = __floats2half2_rn(f2, f3);
          v[2 * q4] = *reinterpret_cast<uint32_t*>(&h0), v[2 * q4 + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
      }
      // Branch-free window maxima.  Everything is a ReLU output (>= +0), so a neighbour that must not take part is
      // replaced by +0, the identity of the maximum, with lane masks -- lane-dependent `if`s around the shuffles compile
      // to divergent branches with a reconvergence barrier per register (ncu: the vertical shuffles stalled on
      // branch resolving for a quarter of all samples).
      const uint32_t m_img = in_image ? 0xFFFFFFFFu : 0u;           // a tile may overhang the right / bottom edge
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t x = v[i] & m_img;
        // horizontal: columns lx-1, lx, lx+1 of this row, clipped to the strip; column 7 also serves, alone, the window
        // centred on column 8 (which belongs to the next tile)
        const uint32_t left = __shfl_up_sync(0xFFFFFFFFu, x, 1) & m_left, right = __shfl_down_sync(0xFFFFFFFFu, x, 1) & m_right;
        const uint32_t hs = hmax2u<BF16>(hmax2u<BF16>(x, left), right);
        // vertical: rows lr-1, lr, lr+1 of the strip.  Row 0 centres the window whose upper row lies in the strip above,
        // row 2 a complete window, row 3 alone serves the window centred on the row below the strip
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, hs, 8) & m_up, down = __shfl_down_sync(0xFFFFFFFFu, hs, 8) & m_down;
        v[i] = hmax2u<BF16>(hmax2u<BF16>(up, hs), down);
      }
      uint8_t* buf = pbuf_warp + (k & 1) * kPoolBufBytes;
      if (lane == 0) bulk_wait_read1();                  // the reduce-store issued two slices ago has left this buffer
      __syncwarp();
      if (holder) {
        uint4* d4 = reinterpret_cast<uint4*>(buf + dst_off);
#pragma unroll
        for (int i = 0; i < 4; ++i) d4[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && !(p.debug & 1)) {
        tma_reduce_max_4d(c.tmO, buf, t.cbase + cl, t.x0 >> 1, (t.y0 >> 1) + 2 * c.q, t.n0);
        bulk_commit();
      }
    }
  }
  if (lane == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------
// MMA issue loop, specialised on (A mode, M tiles per weight tile, resident weights, K steps per stage) so the
// nine-tap body is straight-line code: the issuing thread's instruction stream is the critical path of narrow
// tiles (every extra instruction per tcgen05.mma is tensor-pipe idle time)
// ------------------------------------------------------------------------------------------
struct MmaCtx {
  int items_cta, rank;
  uint32_t tmem_base, a_ring, b_base;
  uint64_t *fullA, *emptyA, *fullB, *emptyB, *tfull, *tempty, *bres, *peerB, *peer_tempty;
};

template <int KSTEPS, bool CG2>
__device__ __forceinline__ void issue_k(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t first_acc) {
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k) {
    if (CG2) umma_f16_cg2(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k == 0 ? first_acc : 1u);
    else umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k == 0 ? first_acc : 1u);
  }
}
template <bool CG2>
__device__ __forceinline__ void commit_to(uint64_t* bar) {
  if (CG2) umma_commit_cg2(bar);       // the barrier at this offset in both CTAs of the pair
  else umma_commit(bar);
}

// CG2: only the pair's leader (rank 0) issues -- M = 256 MMAs over both CTAs' operands, commits multicast to both CTAs
template <int MODE, int MT, bool RES, int KSTEPS, bool CG2>
__device__ __forceinline__ void mma_issuer(const TileParams& p, const MmaCtx& c) {
  if (CG2 && c.rank != 0) return;
  constexpr int TPB = MODE == 0 ? 1 : (MODE == 1 ? 3 : 9);
  const uint32_t row_bytes = p.kchunk * 2;
  const uint32_t idesc = umma_idesc(CG2 ? 256 : 128, (uint32_t)p.block_n, (uint32_t)p.is_bf16);
  const int groups_per_item = p.cchunks * p.boxes_per_chunk / p.ksplit + p.sc_cchunks;   // shortcut chunks ride as extra taps
  const int stages_a = p.stages_a, stages_b = p.stages_b, items_cta = c.items_cta;
  const int acc_mask = (1 << p.n_acc_log2) - 1, acc_log2 = p.n_acc_log2;
  const uint32_t acc_stride = (uint32_t)p.acc_stride, tmem_base = c.tmem_base;
  const int box_w = MODE == 2 ? p.tw + 2 : p.tw;
  // mode 2: an 8-row group of the A tile = 8 consecutive pixels of one box row, groups one box row apart
  const uint32_t sbo = MODE == 2 ? (uint32_t)box_w * row_bytes : 8u * row_bytes;
  const uint64_t a_desc0 = umma_smem_desc_sbo(c.a_ring, row_bytes, sbo);
  const uint64_t b_desc0 = umma_smem_desc(c.b_base, row_bytes);
  const uint64_t a_inc = (uint64_t)(p.a_stage_bytes >> 4), b_inc = (uint64_t)(p.b_stride >> 4);
  const uint64_t u_inc = (uint64_t)(p.a_box_bytes >> 4);
  const uint64_t a_in_b = umma_smem_desc(c.b_base + (uint32_t)p.b_tile_bytes, row_bytes);   // combined stages
  const uint64_t r_inc = (uint64_t)(((uint32_t)box_w * row_bytes) >> 4);       // one row of the box
  const uint64_t s_inc = (uint64_t)(row_bytes >> 4);                             // one pixel
  const bool do_mma = !(p.debug & 4);
  uint64_t* const fullA = c.fullA;
  uint64_t* const emptyA = c.emptyA;
  uint64_t* const fullB = c.fullB;
  uint64_t* const emptyB = c.emptyB;
  uint64_t* const tfull = c.tfull;
  uint64_t* const tempty = c.tempty;
  int sa = 0, sb = 0;
  uint32_t pa = 0, pb = 0;
  if (RES) {
    mbar_wait(c.bres, 0);
    tc_fence_after();
  }
  int seq = 0;
  for (int l = 0; l < items_cta; ++l, seq += MT) {
    const int acc0 = seq & acc_mask, acc1 = (seq + 1) & acc_mask;
    mbar_wait(&tempty[acc0], ((uint32_t)(seq >> acc_log2) & 1u) ^ 1u);
    if (MT == 2) mbar_wait(&tempty[acc1], ((uint32_t)((seq + 1) >> acc_log2) & 1u) ^ 1u);
    if (CG2) {                                   // the other CTA's epilogue released its copy of the accumulator(s)
      mbar_wait_cluster(&c.peer_tempty[acc0], ((uint32_t)(seq >> acc_log2) & 1u) ^ 1u);
      if (MT == 2) mbar_wait_cluster(&c.peer_tempty[acc1], ((uint32_t)((seq + 1) >> acc_log2) & 1u) ^ 1u);
    }
    tc_fence_after();
    const uint32_t d0 = tmem_base + (uint32_t)acc0 * acc_stride, d1 = tmem_base + (uint32_t)acc1 * acc_stride;
    uint64_t bd = b_desc0;                               // resident weights: consumed in load order
    for (int g = 0; g < groups_per_item; ++g) {
      if (MODE == 0 && !RES) {
        // combined stage: weight tile + activation box(es) behind one barrier pair
        mbar_wait(&fullB[sb], pb);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t off = b_inc * (uint64_t)sb;
          if (do_mma) {
            issue_k<KSTEPS, CG2>(d0, a_in_b + off, b_desc0 + off, idesc, (uint32_t)(g != 0));
            if (MT == 2) issue_k<KSTEPS, CG2>(d1, a_in_b + off + u_inc, b_desc0 + off, idesc, (uint32_t)(g != 0));
          }
          commit_to<CG2>(&emptyB[sb]);
          if (g == groups_per_item - 1) {
            commit_to<CG2>(&tfull[acc0]);
            if (MT == 2) commit_to<CG2>(&tfull[acc1]);
          }
        }
        if (++sb == stages_b) sb = 0, pb ^= 1;
        continue;
      }
      mbar_wait(&fullA[sa], pa);
      tc_fence_after();
      const uint64_t ad = a_desc0 + a_inc * (uint64_t)sa;
      if (RES) {
        if (elect_one()) {
          if (do_mma) {
#pragma unroll
            for (int j = 0; j < TPB; ++j) {
              const uint64_t aoff = MODE == 0 ? 0ull : (MODE == 1 ? r_inc * (uint64_t)j
                                                                  : s_inc * (uint64_t)(j / 3) + r_inc * (uint64_t)(j % 3));
              issue_k<KSTEPS, CG2>(d0, ad + aoff, bd + b_inc * (uint64_t)j, idesc, j == 0 ? (uint32_t)(g != 0) : 1u);
              // two tiles per item: their accumulation chains are independent, so alternating them hides the
              // MMA -> MMA accumulate latency that bounds narrow tiles
              if (MT == 2)
                issue_k<KSTEPS, CG2>(d1, ad + aoff + u_inc, bd + b_inc * (uint64_t)j, idesc, j == 0 ? (uint32_t)(g != 0) : 1u);
            }
          }
          commit_to<CG2>(&emptyA[sa]);
          if (g == groups_per_item - 1) {
            commit_to<CG2>(&tfull[acc0]);
            if (MT == 2) commit_to<CG2>(&tfull[acc1]);
          }
        }
        bd += b_inc * (uint64_t)TPB;
      } else {
#pragma unroll
        for (int j = 0; j < TPB; ++j) {
          const uint64_t aoff = MODE == 0 ? 0ull : (MODE == 1 ? r_inc * (uint64_t)j
                                                              : s_inc * (uint64_t)(j / 3) + r_inc * (uint64_t)(j % 3));
          mbar_wait(&fullB[sb], pb);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t bdj = b_desc0 + b_inc * (uint64_t)sb;
            const uint32_t accum = j == 0 ? (uint32_t)(g != 0) : 1u;
            if (do_mma) {
              issue_k<KSTEPS, CG2>(d0, ad + aoff, bdj, idesc, accum);
              if (MT == 2) issue_k<KSTEPS, CG2>(d1, ad + aoff + u_inc, bdj, idesc, accum);
            }
            commit_to<CG2>(&emptyB[sb]);
            if (j == TPB - 1) {
              commit_to<CG2>(&emptyA[sa]);
              if (g == groups_per_item - 1) {
                commit_to<CG2>(&tfull[acc0]);
                if (MT == 2) commit_to<CG2>(&tfull[acc1]);
              }
            }
          }
          if (++sb == stages_b) sb = 0, pb ^= 1;
        }
      }
      if (++sa == stages_a) sa = 0, pa ^= 1;
    }
  }
}

// A mode 3, the 8-channel stem (3x3 over an image of 16-byte pixels): a stage holds ten 2 KB slots, slot t = the
// {8 ch, tw, th, tn} box of filter tap t (slot 9 stays zero), laid out as no-swizzle core matrices, so one K = 16 step
// spans two taps: five MMAs per tile against ten resident weight slots [slot][N rows][8 ch]
constexpr int kStemSlots = 10, kStemSlotBytes = 128 * 16;
__device__ __forceinline__ void mma_issuer_stem(const TileParams& p, const MmaCtx& c) {
  const uint32_t idesc = umma_idesc(128, (uint32_t)p.block_n, (uint32_t)p.is_bf16);
  const uint32_t b_slot = (uint32_t)p.b_tile_bytes;
  const int acc_mask = (1 << p.n_acc_log2) - 1, acc_log2 = p.n_acc_log2, stages_a = p.stages_a;
  const bool do_mma = !(p.debug & 4);
  mbar_wait(c.bres, 0);
  tc_fence_after();
  int sa = 0;
  uint32_t pa = 0;
  for (int seq = 0; seq < c.items_cta; ++seq) {
    const int acc = seq & acc_mask;
    mbar_wait(&c.tempty[acc], ((uint32_t)(seq >> acc_log2) & 1u) ^ 1u);
    mbar_wait(&c.fullA[sa], pa);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t a0 = c.a_ring + (uint32_t)(sa * p.a_stage_bytes);
      const uint32_t d = c.tmem_base + (uint32_t)acc * (uint32_t)p.acc_stride;
      if (do_mma) {
#pragma unroll
        for (int s = 0; s < kStemSlots / 2; ++s)
          umma_f16(d, umma_smem_desc_noswz(a0 + (uint32_t)(s * 2 * kStemSlotBytes), kStemSlotBytes, 128),
                   umma_smem_desc_noswz(c.b_base + (uint32_t)s * 2u * b_slot, b_slot, 128), idesc, s ? 1u : 0u);
      }
      umma_commit(&c.emptyA[sa]);
      umma_commit(&c.tfull[acc]);
    }
    if (++sa == stages_a) sa = 0, pa ^= 1;
  }
}

// A mode 4, the stride-1 stem: ONE box per 8 x 16 tile -- (16 + 2) rows of (8 + 2) 16-byte pixels, 160 B per row, fetched
// through a tensor map that merges the pixel and channel dimensions -- serves all nine taps: a tap's operand starts
// (ky * 10 + kx) pixels into the box, its 8-row groups are one box row (160 B) apart, and the second K core matrix of a
// step is simply the next tap's start (16 B, or 128 B across a row of taps).  The last step pairs tap 8 with zero weights.
constexpr int kStemBoxRow = 10 * 16, kStemBoxBytes = 18 * kStemBoxRow, kStemBoxStage = 3072;
__device__ __forceinline__ void mma_issuer_stem_halo(const TileParams& p, const MmaCtx& c) {
  const uint32_t idesc = umma_idesc(128, (uint32_t)p.block_n, (uint32_t)p.is_bf16);
  const uint32_t b_slot = (uint32_t)p.b_tile_bytes;
  const int acc_mask = (1 << p.n_acc_log2) - 1, acc_log2 = p.n_acc_log2, stages_a = p.stages_a;
  const bool do_mma = !(p.debug & 4);
  mbar_wait(c.bres, 0);
  tc_fence_after();
  int sa = 0;
  uint32_t pa = 0;
  for (int seq = 0; seq < c.items_cta; ++seq) {
    const int acc = seq & acc_mask;
    mbar_wait(&c.tempty[acc], ((uint32_t)(seq >> acc_log2) & 1u) ^ 1u);
    mbar_wait(&c.fullA[sa], pa);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t a0 = c.a_ring + (uint32_t)(sa * kStemBoxStage);
      const uint32_t d = c.tmem_base + (uint32_t)acc * (uint32_t)p.acc_stride;
      if (do_mma) {
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          const int t0 = 2 * s, t1 = t0 + 1;
          const uint32_t o0 = (uint32_t)((t0 / 3) * 10 + t0 % 3) * 16u;
          const uint32_t o1 = s == 4 ? o0 + 16u : (uint32_t)((t1 / 3) * 10 + t1 % 3) * 16u;
          umma_f16(d, umma_smem_desc_noswz(a0 + o0, o1 - o0, kStemBoxRow),
                   umma_smem_desc_noswz(c.b_base + (uint32_t)s * 2u * b_slot, b_slot, 128), idesc, s ? 1u : 0u);
        }
      }
      umma_commit(&c.emptyA[sa]);
      umma_commit(&c.tfull[acc]);
    }
    if (++sa == stages_a) sa = 0, pa ^= 1;
  }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <bool CG2>
__global__ void __launch_bounds__(64 + 128 * kTGroups, 1)
conv_tile_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const TileParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);

  uint8_t* a_ring = smem;
  uint8_t* b_base = smem + p.off_b;
  uint8_t* stg_base = smem + p.off_stg;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* fullA = bars;
  uint64_t* emptyA = bars + kTStages;
  uint64_t* fullB = bars + 2 * kTStages;
  uint64_t* emptyB = bars + 3 * kTStages;
  uint64_t* tfull = bars + 4 * kTStages;
  uint64_t* tempty = tfull + kTAcc;
  uint64_t* bres = tempty + kTAcc;
  uint64_t* rbar = bres + 1;                         // [kTGroups][kStgBufs]
  uint64_t* peerB = rbar + kTGroups * kStgBufs;      // cta_group::2: "the other CTA's stage has landed" (leader side)
  uint64_t* peer_tempty = peerB + kTStages;          // cta_group::2: the other CTA's epilogue released the accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_tempty + kTAcc);
  float* s_bias = reinterpret_cast<float*>(smem + p.off_tab);
  float* s_slope = s_bias + p.bias_classes * p.cout_p;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t row_bytes = p.kchunk * 2;
  const int n_acc = 1 << p.n_acc_log2;
  // work items of this CTA (or CTA pair): first, first + stride, ...; an item = mt consecutive M tiles of one N tile
  // (cta_group::2: two tiles, one per CTA of the pair)
  const int cg2 = CG2 ? 1 : 0;        // compile-time: a kernel that contains cta_group::2 code must be launched as a cluster
  int cta_rank = 0;
  if constexpr (CG2) cta_rank = (int)cluster_ctarank();
  const int first = cg2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int stride_items = cg2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int items_cta = (p.items - first + stride_items - 1) / stride_items;
  const int tiles_cta = items_cta * p.mt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.epi_tma) tma_prefetch_desc(&tmO);
    if (p.res_smem) tma_prefetch_desc(&tmR);
    for (int s = 0; s < kTStages; ++s) {
      mbar_init(&fullA[s], 1);
      mbar_init(&emptyA[s], 1);
      mbar_init(&fullB[s], 1);
      mbar_init(&emptyB[s], 1);
    }
    for (int s = 0; s < kTAcc; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 4);
    }
    mbar_init(bres, 1);
    for (int s = 0; s < kTGroups * kStgBufs; ++s) mbar_init(&rbar[s], 1);
    for (int s = 0; s < kTStages; ++s) mbar_init(&peerB[s], 1);
    for (int s = 0; s < kTAcc; ++s) mbar_init(&peer_tempty[s], 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG2) {
      tmem_alloc2(tmem_slot, 512);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  if (p.a_mode == 3) {                         // the tenth slot of every stage is the zero half of the last K step
    for (int i = threadIdx.x; i < p.stages_a * 128; i += blockDim.x)
      *reinterpret_cast<uint4*>(a_ring + (size_t)(i >> 7) * p.a_stage_bytes + 9 * kStemSlotBytes + (i & 127) * 16) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }
  if (p.a_mode == 4) {                         // the slack after each box is read by the zero-weight half of the last K step
    constexpr int kPad16 = (kStemBoxStage - kStemBoxBytes) / 16;
    for (int i = threadIdx.x; i < p.stages_a * kPad16; i += blockDim.x)
      *reinterpret_cast<uint4*>(a_ring + (size_t)(i / kPad16) * kStemBoxStage + kStemBoxBytes + (i % kPad16) * 16) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }
  for (int i = threadIdx.x; i < p.bias_classes * p.cout_p; i += blockDim.x) s_bias[i] = p.bias[i];
  for (int i = threadIdx.x; i < p.cout_p; i += blockDim.x) s_slope[i] = p.act == 2 ? p.slope[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  if constexpr (CG2) cluster_sync_all();  // the peer's barriers and TMEM exist before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched only weights-side constants (bias table, slopes, tensor maps): it may overlap the tail of
  // the previous layer; activations and residuals are read below
  if (p.pdl) {
    pdl_launch_dependents();
    pdl_wait_primary();
  }

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      const int mode = p.a_mode, mt = p.mt, n_tiles = p.n_tiles, cchunks = p.cchunks, kchunk = p.kchunk;
      const int boxes = p.boxes_per_chunk, tpb = p.taps_per_box, kw = p.kw, block_n = p.block_n;
      const int stages_a = p.stages_a, stages_b = p.stages_b;
      const int a_stage_bytes = p.a_stage_bytes, a_box_bytes = p.a_box_bytes, b_tile_bytes = p.b_tile_bytes;
      const int tiles_x = p.tiles_x, tiles_y = p.tiles_y;
      const int sx_scale = p.tw * p.stride, sy_scale = p.th * p.stride, org = mode == 0 ? p.pad : 1;
      const bool resident = p.b_resident != 0;
      const uint32_t a_tx = (uint32_t)(p.a_bytes * mt);
      // cta_group::2: this CTA loads rows [rank * N/2, +N/2) of each weight tile
      const uint32_t b_bytes = (uint32_t)(cg2 ? block_n / 2 : block_n) * row_bytes;
      const int b_row_off = cg2 ? cta_rank * (block_n / 2) : 0;
      // CTA pair: every load lands in this CTA's shared memory but is counted on the LEADER's barrier, which the
      // leader's producer arms for both CTAs (so its MMA warp waits on one barrier per stage)
      auto arm = [&](uint64_t* bar, uint32_t bytes) {
        if (!CG2) mbar_arrive_expect_tx(bar, bytes);
        else if (cta_rank == 0) mbar_arrive_expect_tx(bar, 2u * bytes);
      };
      auto load_a_from = [&](const CUtensorMap* tm, void* dst, uint64_t* bar, int c0, int x, int y, int n) {
        if (CG2) tma_load_4d_cg2(dst, tm, mapa_u32(bar, 0), c0, x, y, n);
        else tma_load_4d(dst, tm, bar, c0, x, y, n);
      };
      auto load_b_from = [&](const CUtensorMap* tm, void* dst, uint64_t* bar, int c0, int row, int tap) {
        if (CG2) tma_load_3d_cg2(dst, tm, mapa_u32(bar, 0), c0, row, tap);
        else tma_load_3d(dst, tm, bar, c0, row, tap);
      };
      auto load_a = [&](void* dst, uint64_t* bar, int c0, int x, int y, int n) { load_a_from(&tmA, dst, bar, c0, x, y, n); };
      auto load_b = [&](void* dst, uint64_t* bar, int c0, int row, int tap) { load_b_from(&tmB, dst, bar, c0, row, tap); };
      const int sc_cchunks = p.sc_cchunks;
      if (mode == 4) {
        if constexpr (!CG2) {
          mbar_arrive_expect_tx(bres, (uint32_t)(kStemSlots * block_n * 16));
          for (int j = 0; j < kStemSlots; ++j) tma_load_3d(b_base + (size_t)j * b_tile_bytes, &tmB, bres, 0, 0, j);
          int sa = 0;
          uint32_t pa = 0;
          for (int i = first; i < p.items; i += stride_items) {
            int tyx, tx, ty, an;
            p.fd_tx.divmod(i, tyx, tx);
            p.fd_ty.divmod(tyx, an, ty);
            const int ax = tx * p.tw - 1, ay = ty * p.th - 1;
            mbar_wait(&emptyA[sa], pa ^ 1);
            mbar_arrive_expect_tx(&fullA[sa], (uint32_t)kStemBoxBytes);
            tma_load_3d(a_ring + (size_t)sa * kStemBoxStage, &tmA, &fullA[sa], ax * 8, ay, an);
            if (++sa == stages_a) sa = 0, pa ^= 1;
          }
        }
      } else if (mode == 3) {
        if constexpr (!CG2) {
          mbar_arrive_expect_tx(bres, (uint32_t)(kStemSlots * block_n * 16));      // the bytes that land; slots are b_tile_bytes apart
          for (int j = 0; j < kStemSlots; ++j) tma_load_3d(b_base + (size_t)j * b_tile_bytes, &tmB, bres, 0, 0, j);
          const int sxs = p.tw * p.stride, sys = p.th * p.stride, kw3 = p.kw;
          const uint32_t tx = (uint32_t)p.a_bytes * (uint32_t)(p.kh * p.kw);
          int sa = 0;
          uint32_t pa = 0;
          for (int i = first; i < p.items; i += stride_items) {
            int tile_yz, tile_x, tile_y, tile_z;
            p.fd_tx.divmod(i, tile_yz, tile_x);
            p.fd_ty.divmod(tile_yz, tile_z, tile_y);
            const int ax = tile_x * sxs - p.pad, ay = tile_y * sys - p.pad, an = tile_z * p.tn;
            mbar_wait(&emptyA[sa], pa ^ 1);
            mbar_arrive_expect_tx(&fullA[sa], tx);
            uint8_t* dst = a_ring + (size_t)sa * a_stage_bytes;
            for (int ky = 0; ky < p.kh; ++ky)
              for (int kx = 0; kx < kw3; ++kx)
                tma_load_4d(dst + (ky * kw3 + kx) * kStemSlotBytes, &tmA, &fullA[sa], 0, ax + kx, ay + ky, an);
            if (++sa == stages_a) sa = 0, pa ^= 1;
          }
        }
      } else {
      if (resident) {
        arm(bres, b_bytes * (uint32_t)(cchunks * boxes * tpb + sc_cchunks));
        uint8_t* dst = b_base;
        if (mode == 0) {
          for (int b = 0; b < boxes; ++b)                          // consumption order: tap-major, channel chunks inner
            for (int cc = 0; cc < cchunks; ++cc, dst += b_tile_bytes) load_b(dst, bres, cc * kchunk, b_row_off, b);
          for (int cc = 0; cc < sc_cchunks; ++cc, dst += b_tile_bytes) load_b_from(&tmB2, dst, bres, cc * kchunk, b_row_off, 0);
        } else {
          for (int cc = 0; cc < cchunks; ++cc)
            for (int b = 0; b < boxes; ++b)
              for (int j = 0; j < tpb; ++j, dst += b_tile_bytes) {
                const int tap = mode == 1 ? j * 3 + b : (j % 3) * 3 + j / 3;
                load_b(dst, bres, cc * kchunk, b_row_off, tap);
              }
        }
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      // boxes of one channel chunk as a (dy, dx) grid, weight taps of one box as (sxi, r): no division per stage
      const int nbx = mode == 0 ? kw : (mode == 1 ? 3 : 1), nby = mode == 0 ? p.kh : 1;
      const int nsx = mode == 2 ? 3 : 1, nr = mode == 0 ? 1 : 3;
      const bool skip_a = (p.debug & 8) != 0, combined = p.combined != 0;
      const int b_stride = p.b_stride;
      // mode 0 walks taps outer / channel chunks inner (the accumulation order of the first persistent kernel, so
      // either kernel yields the same bits); the halo modes walk chunks outer / boxes inner
      const int n_outer = (mode == 0 ? nby * nbx : cchunks) / p.ksplit, n_inner = mode == 0 ? cchunks : nbx;
      for (int is = first; is < p.items; is += stride_items) {
        int i, split, nt, mp, tyx, tx, ty, tz;                         // split-K (mode 0 only): this item's share of the taps
        p.fd_ksplit.divmod(is, i, split);
        p.fd_ntiles.divmod(i, mp, nt);
        const int nrow = nt * block_n + b_row_off;
        const int m0 = cg2 ? (mp * 2 + cta_rank) * mt : mp * mt, m1 = m0 + 1;
        p.fd_tx.divmod(m0, tyx, tx);
        p.fd_ty.divmod(tyx, tz, ty);
        const int ax0 = tx * sx_scale - org, ay0 = ty * sy_scale - org, an0 = tz * p.tn;
        // the second tile of an mt = 2 item is the next tile in raster order
        if (++tx == tiles_x) { tx = 0; if (++ty == tiles_y) ty = 0, ++tz; }
        const int ax1 = tx * sx_scale - org, ay1 = ty * sy_scale - org, an1 = tz * p.tn;
        int dy = 0, dx = 0;
        if (split) dy = (split * n_outer) / nbx, dx = (split * n_outer) % nbx;
        for (int o = 0; o < n_outer; ++o) {
          for (int in = 0; in < n_inner; ++in) {
            const int c0 = (mode == 0 ? in : o) * kchunk;
            if (mode != 0) dx = in;
            {
              if (combined) {
                // mode 0, streamed weights: one stage = weight tile + activation box(es), one barrier pair per tap
                mbar_wait(&emptyB[sb], pb ^ 1);
                uint8_t* dst = b_base + (size_t)sb * b_stride;
                arm(&fullB[sb], (skip_a && !CG2) ? b_bytes : b_bytes + a_tx);
                if (!skip_a || CG2) {
                  load_a(dst + b_tile_bytes, &fullB[sb], c0, ax0 + dx, ay0 + dy, an0);
                  if (mt == 2) load_a(dst + b_tile_bytes + a_box_bytes, &fullB[sb], c0, ax1 + dx, ay1 + dy, an1);
                }
                load_b(dst, &fullB[sb], c0, nrow, dy * kw + dx);
                if (++sb == stages_b) sb = 0, pb ^= 1;
                continue;
              }
              mbar_wait(&emptyA[sa], pa ^ 1);
              if (skip_a && !CG2) {
                mbar_arrive(&fullA[sa]);
              } else {
                arm(&fullA[sa], a_tx);
                uint8_t* dst = a_ring + (size_t)sa * a_stage_bytes;
                load_a(dst, &fullA[sa], c0, ax0 + dx, ay0 + dy, an0);
                if (mt == 2) load_a(dst + a_box_bytes, &fullA[sa], c0, ax1 + dx, ay1 + dy, an1);
              }
              if (++sa == stages_a) sa = 0, pa ^= 1;
              if (!resident) {
                for (int sxi = 0; sxi < nsx; ++sxi) {
                  for (int r = 0; r < nr; ++r) {
                    const int tap = r * 3 + dx + sxi;            // halo modes only (mode 0 streamed is combined)
                    mbar_wait(&emptyB[sb], pb ^ 1);
                    arm(&fullB[sb], b_bytes);
                    load_b(b_base + (size_t)sb * b_stride, &fullB[sb], c0, nrow, tap);
                    if (++sb == stages_b) sb = 0, pb ^= 1;
                  }
                }
              }
            }
          }
          if (mode == 0 && ++dx == nbx) dx = 0, ++dy;
        }
        // fused projection shortcut: extra K chunks whose activation boxes come from the block input (1x1, no padding)
        for (int cc = 0; cc < sc_cchunks; ++cc) {
          const int c0 = cc * kchunk;
          const int sx0 = (ax0 + org) / p.stride * p.sc_stride, sy0 = (ay0 + org) / p.stride * p.sc_stride;
          const int sx1 = (ax1 + org) / p.stride * p.sc_stride, sy1 = (ay1 + org) / p.stride * p.sc_stride;
          if (combined) {
            mbar_wait(&emptyB[sb], pb ^ 1);
            uint8_t* dst = b_base + (size_t)sb * b_stride;
            arm(&fullB[sb], b_bytes + a_tx);
            load_a_from(&tmA2, dst + b_tile_bytes, &fullB[sb], c0, sx0, sy0, an0);
            if (mt == 2) load_a_from(&tmA2, dst + b_tile_bytes + a_box_bytes, &fullB[sb], c0, sx1, sy1, an1);
            load_b_from(&tmB2, dst, &fullB[sb], c0, nrow, 0);
            if (++sb == stages_b) sb = 0, pb ^= 1;
          } else {                                   // resident weights: only the activation box streams
            mbar_wait(&emptyA[sa], pa ^ 1);
            arm(&fullA[sa], a_tx);
            uint8_t* dst = a_ring + (size_t)sa * a_stage_bytes;
            load_a_from(&tmA2, dst, &fullA[sa], c0, sx0, sy0, an0);
            if (mt == 2) load_a_from(&tmA2, dst + a_box_bytes, &fullA[sa], c0, sx1, sy1, an1);
            if (++sa == stages_a) sa = 0, pa ^= 1;
          }
        }
      }
      }   // mode != 3
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    MmaCtx c;
    c.items_cta = items_cta;
    c.tmem_base = tmem_base;
    c.a_ring = smem_u32(a_ring), c.b_base = smem_u32(b_base);
    c.fullA = fullA, c.emptyA = emptyA, c.fullB = fullB, c.emptyB = emptyB, c.tfull = tfull, c.tempty = tempty, c.bres = bres;
    const int ks = p.kchunk >> 4;
    c.peerB = peerB, c.peer_tempty = peer_tempty, c.rank = cta_rank;
    if (p.a_mode == 3) {
      if constexpr (!CG2) mma_issuer_stem(p, c);
    } else if (p.a_mode == 4) {
      if constexpr (!CG2) mma_issuer_stem_halo(p, c);
    } else {
    const int variant = p.a_mode * 4 + (p.b_resident ? (p.mt == 2 ? 3 : 0) : p.mt);   // (mode, {res, stream mt1, stream mt2, res mt2})
#define B2F_MMA_CASE(MODE, V, MT, RES)                                         \
    case MODE * 4 + V:                                                           \
      if (ks == 4) mma_issuer<MODE, MT, RES, 4, CG2>(p, c);                      \
      else if (ks == 2) mma_issuer<MODE, MT, RES, 2, CG2>(p, c);                 \
      else mma_issuer<MODE, MT, RES, 1, CG2>(p, c);                              \
      break;
    switch (variant) {
      B2F_MMA_CASE(0, 0, 1, true)
      B2F_MMA_CASE(0, 1, 1, false)
      B2F_MMA_CASE(0, 2, 2, false)
      B2F_MMA_CASE(0, 3, 2, true)
      B2F_MMA_CASE(1, 0, 1, true)
      B2F_MMA_CASE(1, 1, 1, false)
      B2F_MMA_CASE(1, 2, 2, false)
      B2F_MMA_CASE(1, 3, 2, true)
      B2F_MMA_CASE(2, 0, 1, true)
      B2F_MMA_CASE(2, 1, 1, false)
      B2F_MMA_CASE(2, 2, 2, false)
      B2F_MMA_CASE(2, 3, 2, true)
      default: break;
    }
    }
#undef B2F_MMA_CASE
  } else if (warp < 2 + 4 * p.groups) {
    // ================================ epilogue ================================
    const int group = (warp - 2) >> 2;
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;                     // accumulator row == pixel of the tile
    const int lx = m % p.tw;
    const int ly = (m / p.tw) % p.th;
    const int lz = m / (p.tw * p.th);
    const bool leader = ((warp - 2) & 3) == 0 && lane == 0;
    const int G = p.groups;

    if (p.epi_tma) {
      EpiCtx c;
      c.tmem_base = tmem_base, c.tfull = tfull, c.tempty = tempty, c.rbar = rbar + group * kStgBufs;
      c.stg = stg_base + (size_t)group * p.stg_bufs * p.stg_bytes;
      c.s_bias = s_bias, c.s_slope = s_slope, c.tmO = &tmO, c.tmR = &tmR;
      c.tiles_cta = tiles_cta, c.group = group, c.q = q, c.lane = lane, c.leader = leader;
      c.first = first, c.stride = stride_items, c.rank = cta_rank, c.peer_tempty = peer_tempty;
      const int variant = p.pool ? 26 + (p.is_bf16 ? 1 : 0)
                          : p.res_reduce ? 24 + (p.is_bf16 ? 1 : 0)
                                         : p.act * 6 + (p.is_bf16 ? 3 : 0) + (p.res_smem ? 1 : (p.res_global ? 2 : 0));
#define B2F_EPI_CASE(ACT)                                                   \
      case ACT * 6 + 0: epilogue_tma<ACT, false, 0, CG2>(p, c); break;             \
      case ACT * 6 + 1: epilogue_tma<ACT, false, 1, CG2>(p, c); break;             \
      case ACT * 6 + 2: epilogue_tma<ACT, false, 2, CG2>(p, c); break;             \
      case ACT * 6 + 3: epilogue_tma<ACT, true, 0, CG2>(p, c); break;              \
      case ACT * 6 + 4: epilogue_tma<ACT, true, 1, CG2>(p, c); break;              \
      case ACT * 6 + 5: epilogue_tma<ACT, true, 2, CG2>(p, c); break;
      switch (variant) {
        B2F_EPI_CASE(0)
        B2F_EPI_CASE(1)
        B2F_EPI_CASE(2)
        B2F_EPI_CASE(3)
        case 24: epilogue_tma<0, false, 3, CG2>(p, c); break;
        case 25: epilogue_tma<0, true, 3, CG2>(p, c); break;
        case 26: if constexpr (!CG2) epilogue_pool<false>(p, c, smem + p.off_pool + (warp - 2) * 2 * kPoolBufBytes); break;
        case 27: if constexpr (!CG2) epilogue_pool<true>(p, c, smem + p.off_pool + (warp - 2) * 2 * kPoolBufBytes); break;
        default: break;
      }
#undef B2F_EPI_CASE
    } else {
      // direct stores (fp32 outputs, up-sampled residual): each thread writes its pixel's channels
      const int esz = p.out_dtype == 2 ? 4 : 2;
      for (int seq = group; seq < tiles_cta; seq += G) {
        const TileCoord t = tile_of(p, first, stride_items, cta_rank, seq);
        const int ox = t.x0 + lx, oy = t.y0 + ly, on = t.n0 + lz;
        const bool valid = (lz < p.tn) && ox < p.Wo && oy < p.Ho && on < p.N;
        int cls = 0;
        if (p.bias_classes == 9) {
          const int iy = oy * p.stride - p.pad, ix = ox * p.stride - p.pad;
          const int cy = iy < 0 ? 0 : (iy + p.kh - 1 >= p.H ? 2 : 1);
          const int cx = ix < 0 ? 0 : (ix + p.kw - 1 >= p.W ? 2 : 1);
          cls = cy * 3 + cx;
        }
        // split-K: the first split carries the bias (the slope table is all zeros without PReLU), each split its own plane
        const float* bias_row = t.split ? s_slope : s_bias + cls * p.cout_p;
        uint8_t* out_base = reinterpret_cast<uint8_t*>(p.out) + (size_t)t.split * (size_t)p.split_stride * 4;
        const size_t pix = ((size_t)on * p.Ho + oy) * p.Wo + ox;
        size_t res_pix = pix;
        if (p.res_mode == 2) {
          const int ry = min(oy >> 1, p.res_h - 1), rx = min(ox >> 1, p.res_w - 1);
          res_pix = ((size_t)on * p.res_h + ry) * p.res_w + rx;
        }
        const int acc = seq & (n_acc - 1);
        const uint32_t ph = (uint32_t)(seq >> p.n_acc_log2) & 1u;
        mbar_wait(&tfull[acc], ph);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
        for (int c0 = 0; c0 < p.block_n && !(p.debug & 16); c0 += 32) {
          uint32_t r[32];
          tmem_ld32(t_addr + (uint32_t)c0, r);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (c0 + h * 16 < p.block_n) {
                const int c = t.cbase + c0 + h * 16;
                float rs[16], f[16];
                if (p.res_mode)
                  load16_as_float(reinterpret_cast<const uint8_t*>(p.residual) + (res_pix * p.cout_p + c) * 2, p.is_bf16, rs);
                epi_math16(p, r + h * 16, bias_row, s_slope, c, p.res_mode ? rs : nullptr, f);
                if (!(p.debug & 1))
                  store16(out_base + (pix * p.cout_p + c) * esz, p.out_dtype, f);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG2 && cta_rank) mbar_arrive_remote(&peer_tempty[acc], 0);
          else mbar_arrive(&tempty[acc]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG2) cluster_sync_all();  // the leader's MMAs read this CTA's shared memory until the very end
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG2) tmem_dealloc2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// sum of the split-K partial planes, split 0 first (fixed order: the result is reproducible bit for bit)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float4* __restrict__ ws, int ksplit, long long n4,
                                                            float4* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = ws[i];
    for (int s = 1; s < ksplit; ++s) {
      const float4 b = ws[(long long)s * n4 + i];
      a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    }
    out[i] = a;
  }
}

// ------------------------------------------------------------------------------------------
// host: plan + launch
// ------------------------------------------------------------------------------------------
static int g_sms = 0;

static inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

void pick_m_tile(int N, int Ho, int Wo, int stride, int* tw, int* th, int* tn);   // umma_conv.cu

struct TilePlan {
  int mode, tw, th, tn, mt, groups, resident, stages_a, stages_b;
  long long m_tiles_plan;
  double cost;
};

// returns kTileDeclined (nothing launched) when `optional` and the first persistent kernel is the better fit
static int conv_tile_plan_launch(const b2f_conv_desc* d, int kchunk, cudaStream_t stream, int optional, int pair);

int conv_tile_launch(const b2f_conv_desc* d, int kchunk, cudaStream_t stream, int optional) {
  return conv_tile_plan_launch(d, kchunk, stream, optional, -1);
}

// pair: -1 = decide here (plan for single CTAs first, re-plan for CTA pairs when the layer qualifies), 0 / 1 = fixed
static int conv_tile_plan_launch(const b2f_conv_desc* d, int kchunk, cudaStream_t stream, int optional, int pair) {
  if (g_sms == 0) {
    int dev = 0;
    B2F_CHECK_CUDA(cudaGetDevice(&dev));
    B2F_CHECK_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int Ho = d->ho, Wo = d->wo;
  TileParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->n, p.Ho = Ho, p.Wo = Wo, p.H = d->h, p.W = d->w, p.cout_p = d->cout_p;
  p.kh = d->kh, p.kw = d->kw, p.stride = d->stride, p.pad = d->pad;
  p.kchunk = kchunk, p.cchunks = d->cin_p / kchunk;
  if (d->sc_in != nullptr) {
    B2F_REQUIRE(d->sc_weight != nullptr && d->sc_cin_p > 0 && d->sc_cin_p % kchunk == 0 && d->sc_stride >= 1,
                "b2f_conv2d: fused shortcut needs sc_weight and sc_cin_p (%d) divisible by the K chunk (%d)", d->sc_cin_p, kchunk);
    B2F_REQUIRE((d->sc_h - 1) / d->sc_stride + 1 == d->ho && (d->sc_w - 1) / d->sc_stride + 1 == d->wo,
                "b2f_conv2d: fused shortcut output size mismatch");
    p.sc_cchunks = d->sc_cin_p / kchunk, p.sc_stride = d->sc_stride;
  }
  const int row_bytes = kchunk * 2, ksteps = kchunk / 16, taps = d->kh * d->kw;
  int n_tiles = (d->cout_p + g_max_block_n - 1) / g_max_block_n;
  while ((d->cout_p % n_tiles) != 0 || ((d->cout_p / n_tiles) % 16) != 0) ++n_tiles;
  p.n_tiles = n_tiles;
  p.block_n = d->cout_p / n_tiles;
  // CTA pairs (cta_group::2): the two CTAs of a cluster run M = 256 MMAs; each holds half of the N rows of every weight
  // tile, so weight traffic, resident-weight footprint and B operand reads per SM halve
  // the 8-channel stem (A mode 3): cin_p == 8, weights [10][cout_p][8] (slot = filter tap, slot 9 zero)
  const bool stem8 = kchunk == 8;
  if (stem8)
    B2F_REQUIRE(d->cin_p == 8 && d->kh == 3 && d->kw == 3 && n_tiles == 1 && d->sc_in == nullptr && d->residual == nullptr,
                "b2f_conv2d: the 8-channel stem form needs a 3x3 kernel, cout_p <= %d and no residual / shortcut", g_max_block_n);
  p.pool = d->pool ? 1 : 0;
  if (p.pool) {
    B2F_REQUIRE(d->kh == 3 && d->kw == 3 && d->stride == 1 && d->pad == 1 && d->act == 1 && d->residual == nullptr &&
                    d->sc_in == nullptr && d->out_dtype == d->dtype && p.block_n % 32 == 0 && p.block_n <= 128 && !stem8,
                "b2f_conv2d: the fused max-pool needs a 3x3 / stride 1 / pad 1 convolution with ReLU, a 16-bit output, "
                "at most 128 channels per tile in multiples of 32, and no residual / shortcut");
    p.pool_h = (Ho - 1) / 2 + 1, p.pool_w = (Wo - 1) / 2 + 1;
  }
  const bool pair_legal = g_tile_cg2 != 0 && p.block_n % 32 == 0 && !stem8 && !p.pool;
  p.cg2 = (pair_legal && (pair == 1 || (pair < 0 && g_tile_cg2 == 2))) ? 1 : 0;
  p.b_tile_bytes = round_up(p.block_n / (p.cg2 ? 2 : 1) * row_bytes, 1024);
  p.acc_stride = round_up(p.block_n, 32);
  p.n_acc_log2 = p.acc_stride <= 128 ? 2 : 1;
  p.is_bf16 = d->dtype == 1;
  p.debug = g_debug;
  p.out = d->out, p.out_dtype = d->out_dtype;
  p.bias = d->bias, p.bias_classes = d->bias_classes;
  p.slope = d->slope, p.act = d->act, p.sig_hi = d->sig_hi;
  p.residual = d->residual, p.res_mode = d->residual ? d->res_mode : 0;
  p.res_h = d->res_h, p.res_w = d->res_w;

  // ---- epilogue flavour -------------------------------------------------------------------------------
  // TMA-store epilogue whenever the output has the activation dtype.  Narrow tiles: 32-channel slices, a ring of
  // three per group, residual prefetched into the ring by TMA.  Wide tiles (N > 128) need the shared memory for
  // four operand stages: 16-channel slices, a ring of two, residual read straight from global memory.
  p.epi_tma = (d->out_dtype == d->dtype && p.res_mode != 2) ? 1 : 0;
  if (g_tile_epi >= 0) p.epi_tma = p.epi_tma && g_tile_epi;
  // A pair's half of all weight tiles of a 128-channel 3x3 layer is 147 KB: it stays resident next to two activation
  // boxes when the staging ring shrinks to 16-channel slices for two epilogue groups (g_tile_big_res = 0 streams it).
  // Layers with a residual need the four-group 32-channel ring more than the resident weights (326 vs 267 us).
  const int b_all_bytes = (taps * p.cchunks + p.sc_cchunks) * p.b_tile_bytes;
  // in-place block output (engine.py aliases the residual's buffer to the output when nothing else reads it): the add
  // happens in L2, as a 16-bit add of the rounded conv result and the stored value.  The first-generation kernels, which
  // still take wide layers with few tiles, round the same way (`res_round` in umma_conv.cu), so the dispatch between
  // the kernels does not change a bit.
  p.res_reduce = (g_tile_reduce && p.epi_tma && p.res_mode == 1 && d->residual == d->out && d->act == 0) ? 1 : 0;
  const int res_eff = p.res_reduce ? 0 : p.res_mode;                 // what the epilogue still has to read
  const bool big_res = g_tile_big_res && p.cg2 && n_tiles == 1 && p.epi_tma && res_eff == 0 && b_all_bytes > 100 * 1024 &&
                       b_all_bytes <= 150 * 1024;
  if (p.epi_tma) {
    // wide tiles on CTA pairs stream half-size weight tiles, which leaves room for the narrow-tile epilogue (32-channel
    // slices, ring of three, TMA-fed residual) next to five operand stages; g_tile_wide_res = 0 keeps the lean one
    const bool wide = p.block_n > 128 &&
                      !(p.cg2 && (res_eff == 1 || g_tile_wide_res == 2) && g_tile_wide_res && p.block_n % 32 == 0);
    p.ochunk = (!wide && !big_res && p.block_n % 32 == 0) ? 32 : 16;
    p.n_sub = p.block_n / p.ochunk;
    p.stg_bytes = round_up(128 * p.ochunk * 2, 1024);
    p.stg_bufs = wide ? 2 : 3;
    p.res_smem = (res_eff == 1 && !wide) ? 1 : 0;
    p.res_global = (res_eff == 1 && wide) ? 1 : 0;
  }
  const int tab_bytes = round_up((p.bias_classes + 1) * p.cout_p * 4, 256);
  const int pool_bytes = p.pool ? kTGroups * 4 * 2 * kPoolBufBytes : 0;      // two 1 KB buffers per epilogue warp
  const int fixed = 1024 /*align*/ + 1024 /*barriers*/ + tab_bytes + pool_bytes;

  // ---- choose A mode, tile geometry, weight sharing and epilogue groups with a per-tile cycle model -----------
  const double kFabric = 64.0;                               // L2 -> SM bytes per clock per SM the rings can pull (measured 50-65)
  const double opnd = (128 + p.block_n / (p.cg2 ? 2.0 : 1.0)) / 4.0;      // operand reads per MMA, clocks at 128 B/clk
  const double mma_tap = ksteps * ((p.block_n / 2.0) > opnd ? (p.block_n / 2.0) : opnd);
  // halo modes only for narrow tiles: wide ones keep one accumulation order across both kernel generations
  const bool halo_ok = g_vhalo && d->kh == 3 && d->kw == 3 && d->stride == 1 && d->pad == 1 && p.block_n <= 128 &&
                       p.sc_cchunks == 0;
  const int b_all = (stem8 ? kStemSlots : taps * p.cchunks + p.sc_cchunks) * p.b_tile_bytes;
  // The plan (A mode, weight sharing) fixes the order in which taps are accumulated; it is chosen for a batch of
  // at least 128 images so that an image's result does not depend on how many others share its launch.
  const int n_plan = d->n < 128 ? 128 : d->n;
  TilePlan best;
  best.cost = -1;
  if (stem8) {
    // one plan: a stage = ten 2 KB tap slots, weights resident, as many stages and epilogue groups as fit
    for (int groups = 4; groups >= 2 && best.cost < 0; groups -= 2) {
      const int staging = p.epi_tma ? groups * p.stg_bufs * p.stg_bytes : 0;
      const bool halo = d->stride == 1 && d->pad == 1 && g_vhalo;        // one box per tile instead of nine
      int stages_a = (kSmemMax - fixed - staging - b_all) / (halo ? kStemBoxStage : kStemSlots * kStemSlotBytes);
      if (stages_a > kTStages) stages_a = kTStages;
      if (stages_a < 2) continue;
      best.cost = 1.0, best.mode = halo ? 4 : 3, best.mt = 1, best.groups = groups, best.resident = 1, best.stages_a = stages_a, best.stages_b = 0;
    }
  }
  for (int relax = 0; relax < 2 && best.cost < 0 && !stem8; ++relax) {      // forced knobs that cannot fit are dropped
  const int f_groups = relax ? 0 : g_tile_groups, f_mt = relax ? 0 : g_tile_mt;
  for (int mode = 0; mode <= (halo_ok ? 2 : 0); ++mode) {
    if (g_tile_amode >= 0 && halo_ok && mode != g_tile_amode && !p.pool) continue;
    if (p.pool && mode != 2) continue;                         // the pooling epilogue is written for 8 x 16 tiles
    for (int tw = (mode == 0 ? 0 : 8); tw <= (mode == 1 ? 128 : (mode == 2 ? 8 : 0)); tw = tw ? tw * 2 : 1) {
      int gw, gh, gn;
      if (mode == 0) pick_m_tile(n_plan, Ho, Wo, d->stride, &gw, &gh, &gn);
      else gw = tw, gh = 128 / tw, gn = 1;
      if (mode == 1 && gw > round_up(Wo, 8)) break;
      const long long m_tiles = (long long)((Wo + gw - 1) / gw) * ((Ho + gh - 1) / gh) * ((n_plan + gn - 1) / gn);
      const int boxes = mode == 0 ? taps : (mode == 1 ? 3 : 1);
      const int a_bytes = (mode == 0 ? gw * gh * gn : (mode == 1 ? gw * (gh + 2) : (gw + 2) * (gh + 2))) * row_bytes;
      const int a_box = round_up(a_bytes, 1024);
      for (int mt = 1; mt <= 2; ++mt) {
        const int resident = (n_tiles == 1 && (b_all <= 100 * 1024 || big_res)) ? 1 : 0;
        if (f_mt && mt != f_mt && !(mt == 1 && (p.n_acc_log2 < 2 || m_tiles < 2))) continue;
        if (mt == 2 && (p.n_acc_log2 < 2 || m_tiles < 2)) continue;
        // two interleaved tiles over resident weights hide the MMA -> MMA accumulate latency of short-K layers (1x1,
        // per-tap boxes, N = 32, Cin <= 32); on 64-channel halo-box layers the doubled activation stage costs more than it hides
        if (mt == 2 && resident && !f_mt && !(mode == 0 || p.block_n <= 32 || kchunk <= 32)) continue;
        if (p.cg2 && m_tiles < 2 * mt) continue;
        for (int groups = 4; groups >= 2; groups -= 2) {
          if (groups > (1 << p.n_acc_log2)) continue;          // a group must never be a whole accumulator phase ahead
          if (f_groups && groups != f_groups && f_groups <= (1 << p.n_acc_log2)) continue;
          const int staging = (p.epi_tma && !p.pool) ? groups * p.stg_bufs * p.stg_bytes : 0;   // pooled layers stage nothing
          int avail = kSmemMax - fixed - staging;
          int stages_a = 0, stages_b = 0;
          const int tpb = taps / boxes;
          if (resident) {
            stages_a = (avail - b_all) / (mt * a_box);
            if (stages_a > kTStages) stages_a = kTStages;
          } else if (mode == 0) {
            // combined stages: weight tile + activation box(es) per tap behind one barrier pair
            stages_b = avail / (p.b_tile_bytes + mt * a_box);
            if (stages_b > kTStages) stages_b = kTStages;
            stages_a = stages_b;
          } else {
            // s weight tiles in flight, and the activation boxes that feed them (one box serves tpb taps)
            for (int sb = kTStages; sb >= 2 && !stages_b; --sb) {
              int sa = (sb + tpb - 1) / tpb + 1;
              if (sa < 2) sa = 2;
              if (sa > kTStages) sa = kTStages;
              if (sa * mt * a_box + sb * p.b_tile_bytes <= avail) stages_a = sa, stages_b = sb;
            }
          }
          if (stages_a < 2) continue;
          // cycles per M tile: tensor pipe (operand reads from shared memory included), L2 -> SM traffic at the
          // rate the bytes in flight can sustain (Little: ~2000 cycles from issue to a reusable slot), epilogue
          const double mma = (taps * p.cchunks + p.sc_cchunks) * mma_tap;
          const double b_bytes = (double)p.block_n * row_bytes / (p.cg2 ? 2.0 : 1.0);
          const double bytes = (double)(p.cchunks * boxes + p.sc_cchunks) * a_bytes +
                               (resident ? 0.0 : (double)(taps * p.cchunks + p.sc_cchunks) * b_bytes / mt);
          const double inflight = (double)(stages_a - 1) * mt * a_bytes + (resident ? 0.0 : (double)(stages_b - 1) * b_bytes);
          double rate = inflight / 2000.0;
          if (rate > kFabric) rate = kFabric;
          const double fab = bytes / rate;
          const double epi = (p.epi_tma ? (p.block_n / 32.0) * 900.0 : (p.block_n / 32.0) * 1500.0) / groups;
          // TMA writes share the shared-memory port with the MMA operand reads (128 B/clk)
          double t = mma + bytes / 128.0;
          if (fab > t) t = fab;
          if (epi > t) t = epi;
          // hand-shake overhead of the single issuing threads: per activation box, and per streamed weight tile
          t += 150.0 * p.cchunks * boxes / mt + (resident ? 0.0 : 60.0 * p.cchunks * taps / mt);
          const double cost = (double)m_tiles * n_tiles * t;
          if (best.cost < 0 || cost < best.cost) {
            best.cost = cost, best.mode = mode, best.tw = gw, best.th = gh, best.tn = gn, best.mt = mt;
            best.groups = groups, best.resident = resident, best.stages_a = stages_a, best.stages_b = stages_b;
            best.m_tiles_plan = m_tiles;
          }
        }
      }
      if (mode != 1) break;
    }
  }
  }
  if (best.cost < 0 && p.cg2)                     // a one-tile problem cannot be paired: plan for single CTAs
    return conv_tile_plan_launch(d, kchunk, stream, optional, 0);
  if (pair < 0 && g_tile_cg2 == 1 && pair_legal && !p.cg2 && best.cost >= 0) {
    // CTA pairs pay when the tile is MMA-heavy rather than epilogue- or hand-shake-bound (1x1 and other short-K layers
    // stay single: the pair's one issuing thread would carry twice the per-tile hand-shakes); measured in
    // profiles/r01_conv_sweep_cta_pairs.log
    const double tile_cyc = (double)(taps * p.cchunks + p.sc_cchunks) * mma_tap;
    if (p.block_n > g_tile_cg2_min_n || tile_cyc >= 1200.0)
      return conv_tile_plan_launch(d, kchunk, stream, optional, 1);
  }
  B2F_REQUIRE(best.cost >= 0, "conv: no tile plan fits in shared memory (cin_p %d cout_p %d k %d)", d->cin_p, d->cout_p, d->kh);
  if (best.mode == 0 || best.mode == 3) pick_m_tile(d->n, Ho, Wo, d->stride, &best.tw, &best.th, &best.tn);   // geometry for the real batch
  if (best.mode == 4) best.tw = 8, best.th = 16, best.tn = 1;
  p.a_mode = best.mode, p.tw = best.tw, p.th = best.th, p.tn = best.tn, p.mt = best.mt, p.groups = best.groups;
  p.b_resident = best.resident, p.stages_a = best.stages_a, p.stages_b = best.stages_b;
  p.tiles_x = (Wo + p.tw - 1) / p.tw;
  p.tiles_y = (Ho + p.th - 1) / p.th;
  p.m_tiles = p.tiles_x * p.tiles_y * ((d->n + p.tn - 1) / p.tn);
  p.items = n_tiles * ((p.m_tiles + p.mt - 1) / p.mt);
  // wide tiles have two accumulators and two epilogue groups: with only a few tiles per SM the exposed epilogue of
  // the last tile costs more than the first persistent kernel's column-split epilogue
  // (both kernels accumulate wide tiles in the same order, so this batch-dependent choice does not change results)
  // -- except when a tile is hundreds of K chunks long (the 7x7x512 embedding layer): there the epilogue is noise and
  // the deeper operand ring wins (138 -> 118 us sustained)
  if (optional && p.block_n > 128 && p.items < 4 * g_sms && p.sc_cchunks == 0 && taps * p.cchunks < 100) return kTileDeclined;
  const bool per_tap = p.a_mode == 0 || p.a_mode == 3;
  p.boxes_per_chunk = per_tap ? taps : (p.a_mode == 1 ? 3 : 1);
  p.taps_per_box = taps / p.boxes_per_chunk;
  const int box_w = per_tap ? p.tw * d->stride : (p.a_mode == 1 ? p.tw : p.tw + 2);
  const int box_h = per_tap ? p.th * d->stride : p.th + 2;
  p.a_bytes = (per_tap ? p.tw * p.th * p.tn : box_w * box_h) * row_bytes;
  // the slot always holds the 128 rows an MMA reads, also when the whole problem has fewer output pixels than one tile
  // (a short box would let the last stage's operand read run past the end of the shared-memory allocation)
  p.a_box_bytes = round_up(p.a_mode == 0 && p.a_bytes < 128 * row_bytes ? 128 * row_bytes : p.a_bytes, 1024);
  p.a_stage_bytes = p.a_box_bytes * p.mt;
  if (p.a_mode == 3) p.a_box_bytes = kStemSlotBytes, p.a_stage_bytes = kStemSlots * kStemSlotBytes;
  if (p.a_mode == 4) p.a_bytes = kStemBoxBytes, p.a_box_bytes = kStemBoxStage, p.a_stage_bytes = kStemBoxStage;
  p.stg_box_bytes = p.tw * p.th * p.tn * p.ochunk * 2;
  p.combined = (p.a_mode == 0 && !p.b_resident) ? 1 : 0;
  if (p.cg2) p.items = n_tiles * ((p.m_tiles + 2 * p.mt - 1) / (2 * p.mt));      // an item = 2 x mt M tiles, mt per CTA
  // split-K for long-K layers with a handful of work items (the 7 x 7 x 512 -> 512 embedding layer: 16 items of 392 K
  // chunks for 148 SMs, bound by what one SM can pull from L2): the taps are dealt to `ksplit` items per tile, which store
  // fp32 partial sums into the caller's workspace; splitk_reduce_kernel adds them in a fixed order.  The choice depends on
  // the layer and the workspace only, never on the batch, so an image's result does not depend on its batch mates.
  p.ksplit = 1;
  if (d->splitk_ws != nullptr && p.combined && !p.epi_tma && d->out_dtype == 2 && p.sc_cchunks == 0 && p.res_mode == 0 &&
      d->act == 0 && d->bias_classes == 1 && taps * p.cchunks >= 128) {
    int ks = 8;
    while (ks > 1 && taps % ks != 0) --ks;
    const long long plane = (long long)d->n * Ho * Wo * d->cout_p;
    if (ks > 1 && d->splitk_ws_bytes >= (long long)ks * plane * 4) {
      p.ksplit = ks, p.split_stride = plane, p.items *= ks;
      p.out = d->splitk_ws;
    }
  }
  p.fd_ntiles.set(p.n_tiles), p.fd_tx.set(p.tiles_x), p.fd_ty.set(p.tiles_y), p.fd_mt.set(p.mt), p.fd_ksplit.set(p.ksplit);
  p.b_stride = p.combined ? p.b_tile_bytes + p.a_stage_bytes : p.b_tile_bytes;
  if (p.combined) p.stages_a = 0;                      // activation boxes live inside the weight stages
  p.off_b = p.stages_a * p.a_stage_bytes;
  p.off_stg = p.off_b + (p.b_resident ? b_all : p.stages_b * p.b_stride);
  p.off_bar = p.off_stg + ((p.epi_tma && !p.pool) ? p.groups * p.stg_bufs * p.stg_bytes : 0);
  p.off_tab = p.off_bar + 1024;
  p.off_pool = p.off_tab + tab_bytes;
  size_t smem = (size_t)p.off_pool + pool_bytes + 1024;
  B2F_REQUIRE(smem <= (size_t)kSmemMax, "conv tile kernel: %zu bytes of shared memory requested", smem);
  if (smem < 120 * 1024) smem = 120 * 1024;      // one CTA per SM: it owns all 512 TMEM columns

  CUtensorMap tmA, tmB, tmO, tmR, tmA2, tmB2;
  if (p.a_mode == 4) {
    // pixels and channels merged into one dimension: a box row is (8 + 2) whole 16-byte pixels
    uint64_t dims[3] = {(uint64_t)d->w * 8, (uint64_t)d->h, (uint64_t)d->n};
    uint64_t str[2] = {(uint64_t)d->w * 16, (uint64_t)d->h * d->w * 16};
    uint32_t box[3] = {80, 18, 1};
    uint32_t es[3] = {1, 1, 1};
    int rc = make_tmap(&tmA, d->in, 3, dims, str, box, es, 0, p.is_bf16);
    if (rc) return rc;
  } else {
    uint64_t dims[4] = {(uint64_t)d->cin_p, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->n};
    uint64_t str[3] = {(uint64_t)d->cin_p * 2, (uint64_t)d->w * d->cin_p * 2, (uint64_t)d->h * d->w * d->cin_p * 2};
    uint32_t box[4] = {(uint32_t)kchunk, (uint32_t)box_w, (uint32_t)box_h, (uint32_t)p.tn};
    uint32_t es[4] = {1, 1, 1, 1};
    if (per_tap) es[1] = es[2] = (uint32_t)d->stride;
    int rc = make_tmap(&tmA, d->in, 4, dims, str, box, es, row_bytes, p.is_bf16);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)d->cin_p, (uint64_t)d->cout_p, (uint64_t)(stem8 ? kStemSlots : taps)};
    uint64_t str[2] = {(uint64_t)d->cin_p * 2, (uint64_t)d->cout_p * d->cin_p * 2};
    uint32_t box[3] = {(uint32_t)kchunk, (uint32_t)(p.cg2 ? p.block_n / 2 : p.block_n), 1};
    uint32_t es[3] = {1, 1, 1};
    int rc = make_tmap(&tmB, d->weight, 3, dims, str, box, es, row_bytes, p.is_bf16);
    if (rc) return rc;
  }
  tmA2 = tmA, tmB2 = tmB;
  if (p.sc_cchunks) {
    uint64_t dims[4] = {(uint64_t)d->sc_cin_p, (uint64_t)d->sc_w, (uint64_t)d->sc_h, (uint64_t)d->n};
    uint64_t str[3] = {(uint64_t)d->sc_cin_p * 2, (uint64_t)d->sc_w * d->sc_cin_p * 2, (uint64_t)d->sc_h * d->sc_w * d->sc_cin_p * 2};
    uint32_t box[4] = {(uint32_t)kchunk, (uint32_t)(p.tw * d->sc_stride), (uint32_t)(p.th * d->sc_stride), (uint32_t)p.tn};
    uint32_t es[4] = {1, (uint32_t)d->sc_stride, (uint32_t)d->sc_stride, 1};
    int rc = make_tmap(&tmA2, d->sc_in, 4, dims, str, box, es, row_bytes, p.is_bf16);
    if (rc) return rc;
    uint64_t dimsb[3] = {(uint64_t)d->sc_cin_p, (uint64_t)d->cout_p, 1};
    uint64_t strb[2] = {(uint64_t)d->sc_cin_p * 2, (uint64_t)d->cout_p * d->sc_cin_p * 2};
    uint32_t boxb[3] = {(uint32_t)kchunk, (uint32_t)(p.cg2 ? p.block_n / 2 : p.block_n), 1};
    uint32_t esb[3] = {1, 1, 1};
    rc = make_tmap(&tmB2, d->sc_weight, 3, dimsb, strb, boxb, esb, row_bytes, p.is_bf16);
    if (rc) return rc;
  }
  tmO = tmA, tmR = tmA;
  if (p.epi_tma) {
    uint64_t dims[4] = {(uint64_t)d->cout_p, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)d->n};
    uint64_t str[3] = {(uint64_t)d->cout_p * 2, (uint64_t)Wo * d->cout_p * 2, (uint64_t)Ho * Wo * d->cout_p * 2};
    uint32_t box[4] = {(uint32_t)p.ochunk, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
    uint32_t es[4] = {1, 1, 1, 1};
    int rc = 0;
    if (p.pool) {
      B2F_REQUIRE(p.tw == 8 && p.th == 16 && p.tn == 1 && p.ochunk == 32, "conv tile kernel: fused max-pool needs 8 x 16 tiles");
      uint64_t pdims[4] = {(uint64_t)d->cout_p, (uint64_t)p.pool_w, (uint64_t)p.pool_h, (uint64_t)d->n};
      uint64_t pstr[3] = {(uint64_t)d->cout_p * 2, (uint64_t)p.pool_w * d->cout_p * 2, (uint64_t)p.pool_h * p.pool_w * d->cout_p * 2};
      uint32_t pbox[4] = {32, (uint32_t)kPoolW, (uint32_t)kPoolH, 1};
      rc = make_tmap(&tmO, d->out, 4, pdims, pstr, pbox, es, 0, d->out_dtype == 1);
      if (rc) return rc;
      // 0 is the identity of the max over ReLU outputs: the tiles' partial window maxima are max-reduced into the map
      B2F_CHECK_CUDA(cudaMemsetAsync(d->out, 0, (size_t)d->n * p.pool_h * p.pool_w * d->cout_p * 2, stream));
    } else {
      rc = make_tmap(&tmO, d->out, 4, dims, str, box, es, p.ochunk * 2, d->out_dtype == 1);
    }
    if (rc) return rc;
    if (p.res_smem) {
      rc = make_tmap(&tmR, d->residual, 4, dims, str, box, es, p.ochunk * 2, p.is_bf16);
      if (rc) return rc;
    }
  }
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] {
    attr_rc = cudaFuncSetAttribute(conv_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (attr_rc == cudaSuccess)
      attr_rc = cudaFuncSetAttribute(conv_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
  });
  B2F_CHECK_CUDA(attr_rc);
  static const bool trace = getenv("B2F_PLAN_TRACE") != nullptr;
  if (trace)
    fprintf(stderr, "[b2f plan] n%d %dx%d %d->%d k%d s%d: mode %d tile %dx%dx%d mt %d groups %d resident %d stagesA %d stagesB %d "
            "block_n %d kchunk %d epi_tma %d res_smem %d ochunk %d bufs %d cg2 %d smem %zu items %d\n", d->n, d->h, d->w, d->cin_p, d->cout_p, d->kh, d->stride,
            p.a_mode, p.tw, p.th, p.tn, p.mt, p.groups, p.b_resident, p.stages_a, p.stages_b, p.block_n, kchunk, p.epi_tma,
            p.res_smem, p.ochunk, p.stg_bufs, p.cg2, smem, p.items);
  p.pdl = g_tile_pdl ? 1 : 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (p.pdl) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.blockDim = dim3(64 + 128 * p.groups), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  if (p.cg2) {
    const int clusters = p.items < g_sms / 2 ? p.items : g_sms / 2;
    cfg.gridDim = dim3(2 * clusters);
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = 2, attr[n_attr].val.clusterDim.y = 1, attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
    cfg.attrs = attr, cfg.numAttrs = n_attr;
    B2F_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tile_kernel<true>, tmA, tmB, tmO, tmR, tmA2, tmB2, p));
  } else {
    cfg.gridDim = dim3(p.items < g_sms ? p.items : g_sms);
    cfg.attrs = attr, cfg.numAttrs = n_attr;
    B2F_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tile_kernel<false>, tmA, tmB, tmO, tmR, tmA2, tmB2, p));
  }
  g_launches.fetch_add(1);
  if (p.ksplit > 1) {
    const long long n4 = p.split_stride / 4;              // cout_p is a multiple of 16
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(d->splitk_ws), p.ksplit, n4,
                                                      reinterpret_cast<float4*>(d->out));
    g_launches.fetch_add(1);
    B2F_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace b2f
