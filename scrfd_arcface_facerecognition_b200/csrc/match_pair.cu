// Second-generation cosine matching / all-pairs kernel for sm_100a: persistent CTA PAIRS (tcgen05 cta_group::2).
//
// Replaces, for more than 128 query rows, the first-generation `umma_conv_kernel<EPI_TOPK|EPI_PAIRS>` behind
// b2f_match_partial (reference main.py:136-142 scan, qdrant_manager.py:164-170 search) and b2f_pairs_threshold
// (reference duplicate.py:2726-2797: N searches of N).
//
// Why pairs: the query tile is resident and only gallery tiles stream, so the kernel is bound by the L2 -> SM fabric
// (~40 B/clk/SM): a single CTA doing 128 x 256 x 512 per gallery tile needs 256 KB per 4096 MMA clocks = 64 B/clk.
// Two CTAs of a cluster run ONE M = 256 MMA per K step: each holds its own 128 query rows and loads only HALF of the
// rows of every gallery tile (32 B/clk/SM), the tensor core reads the other half from the peer's shared memory.
// Why persistent: one launch = a list of items (query-row pair tile x gallery range), dealt round-robin to the 74
// pairs, so neither a third partial wave (304 CTAs on 148 SMs) nor the ragged upper triangle of the all-pairs case
// leaves SMs idle.
//
// Per CTA:  warp 0 = TMA producer (query tile of the item, then its half of every gallery tile through a 5-stage ring),
//           warp 1 = MMA issuer (pair leader only) into two 256-column TMEM accumulators,
//           warps 2..9 = epilogue, two groups of four warps splitting the 256 columns: running top-k per query row
//           (strict '>' keeps the lowest index among equals, as the reference scan does) or thresholded pair emission
//           with an exact fp32 re-check.
// Accumulation order over K is the same as in the first-generation kernel (eight 64-wide chunks, four K = 16 steps
// each), so either kernel yields the same coarse scores.
#include "umma_shared.cuh"
#include "../../include/b2f.h"

#include <atomic>
#include <mutex>

namespace b2f {

extern std::atomic<long long> g_launches;
int g_match_pair = 1;          // b2f_set_tuning key 17: 0 = always the first-generation kernel

constexpr int kMpStages = 5;
constexpr int kMpTile = 128 * 128;         // 128 rows x 128 B (one 64-element K chunk, SWIZZLE_128B): 16 KB
constexpr int kMpThreads = 64 + 2 * 128;
constexpr int kMpTopK = 8;
enum { MP_TOPK = 0, MP_PAIRS = 1 };

struct MatchPairParams {
  int q;                    // query rows (A)
  long long g;              // gallery rows (B)
  int kchunks;              // dim / 64
  int m_pairs, n_tiles, splits, items;
  int is_bf16;
  // MP_TOPK
  int causal;               // 1: query row i only sees gallery rows with index < causal_base + i (a prefix search)
  long long causal_base;
  int topk;
  float* part_score;        // [q][splits][2 column halves][topk]
  int* part_idx;
  // MP_PAIRS: rows row_begin.. of the same matrix against all rows; only j > i is reported
  int row_begin;
  float thr_coarse, thr_exact;
  const float* exact_rows;
  int exact_dim;
  long long* pairs;
  long long max_pairs;
  unsigned long long* pair_count;
};

// gallery tile range [begin, end) of item `it`.  MP_PAIRS: tiles that lie entirely at or below the diagonal of the
// pair tile's first row are skipped, and the remaining range is what gets split, so the items of one pair tile are even
// The items of a launch and the order the pairs take them in.  MP_TOPK: split-major (it = split * m_pairs + m_pair), so
// the pairs running at the same time walk the SAME gallery range for different query tiles and each gallery tile is
// read from DRAM once, then from L2.  MP_PAIRS: query tiles in row order (the cost of a tile falls linearly with its row:
// only column tiles above the diagonal run), dealt to the pairs in snake order so every pair gets the same total.
template <int EPI>
__device__ __forceinline__ int item_of(const MatchPairParams& p, int pair_id, int n_pairs, int round) {
  if (EPI == MP_PAIRS && (round & 1)) return round * n_pairs + (n_pairs - 1 - pair_id);
  return round * n_pairs + pair_id;
}
// gallery tile range [begin, end) of item `it`.  MP_PAIRS: tiles that lie entirely at or below the diagonal of the
// pair tile's first row are skipped, and the remaining range is what gets split, so the items of one pair tile are even
template <int EPI>
__device__ __forceinline__ void item_range(const MatchPairParams& p, int it, int& m_pair, int& split, int& nt_begin, int& nt_end) {
  if (EPI == MP_TOPK) {
    split = it / p.m_pairs;
    m_pair = it - split * p.m_pairs;
  } else {
    m_pair = it / p.splits;
    split = it - m_pair * p.splits;
  }
  int first = 0, last = p.n_tiles;
  if (EPI == MP_PAIRS) first = min((p.row_begin + m_pair * 256) / 256, p.n_tiles);
  if (EPI == MP_TOPK && p.causal) {      // tiles past the last column any row of this pair tile may see are skipped
    const long long hi = p.causal_base + (long long)m_pair * 256 + 255;          // largest limit among the tile's rows
    last = (int)min((long long)p.n_tiles, hi <= 0 ? 0ll : (hi - 1) / 256 + 1);
  }
  const int span = max(last - first, 0);
  const int per = (span + p.splits - 1) / p.splits;
  nt_begin = min(first + split * per, last);
  nt_end = min(nt_begin + per, last);
}

// running top-K of one query row (one thread) over the gallery tiles of an item, K a compile-time list length.
// Per 32-column chunk: four 8-column group maxima -> chunk maximum; the chunk is skipped unless it beats the current
// K-th best (`worst`), then only the groups that beat it are walked.  The insertion is strict '>' from the top of the
// descending list, so among equal scores the lowest gallery index stays in front (the order of the reference scan).
template <int K>
__device__ __forceinline__ void topk_item(const MatchPairParams& p, uint32_t tmem_base, uint64_t* tfull, uint64_t* tempty,
                                          uint64_t* peer_tempty, int rank, int qd, int lane, int c_begin, int nt_begin, int nt_end,
                                          int& t, bool valid, int on, int split, int group) {
  float best_s[K];
  int best_i[K];
#pragma unroll
  for (int k = 0; k < K; ++k) best_s[k] = -INFINITY, best_i[k] = -1;
  float worst = -INFINITY;
  // columns this row may see: the whole gallery, or (causal) only the rows before its own global index
  const long long lim = p.causal ? max(0ll, min(p.g, p.causal_base + (long long)on)) : p.g;
  for (int nt = nt_begin; nt < nt_end; ++nt, ++t) {
    const int acc = t & 1;
    mbar_wait(&tfull[acc], (uint32_t)(t >> 1) & 1u);
    tc_fence_after();
    const uint32_t t_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * 256);
    const long long cbase = (long long)nt * 256;
    const bool full_tile = cbase + 256 <= lim;
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_begin + 128; c0 += 32) {
      uint32_t r[32];
      __syncwarp();                               // the TMEM load is warp-collective: reconverge first
      tmem_ld32(t_addr + (uint32_t)c0, r);
      tmem_ld_wait();
      const int g0 = (int)cbase + c0;
      if (!full_tile) {                           // ragged last tile / the causal limit: columns past it can never win
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (g0 + i >= lim) r[i] = 0xff800000u;  // -inf
      }
      float gm[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        gm[j] = __uint_as_float(r[8 * j]);
#pragma unroll
        for (int i = 1; i < 8; ++i) gm[j] = fmaxf(gm[j], __uint_as_float(r[8 * j + i]));
      }
      if (!(fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3])) > worst)) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!(gm[j] > worst)) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v = __uint_as_float(r[8 * j + i]);
          if (v > worst) {
            float cs = v;
            int ci = g0 + 8 * j + i;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              if (cs > best_s[k]) {
                const float ts = best_s[k];
                const int ti = best_i[k];
                best_s[k] = cs, best_i[k] = ci;
                cs = ts, ci = ti;
              }
            }
            worst = best_s[K - 1];
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (rank) mbar_arrive_remote_relaxed(&peer_tempty[acc], 0);
      else mbar_arrive(&tempty[acc]);
    }
  }
  if (valid) {
    // the caller's lists are p.topk long; slots beyond K stay empty
    const size_t o = (((size_t)on * p.splits + split) * 2 + group) * p.topk;
#pragma unroll
    for (int k = 0; k < kMpTopK; ++k) {
      if (k < p.topk) {
        p.part_score[o + k] = k < K ? best_s[k < K ? k : 0] : -INFINITY;
        p.part_idx[o + k] = k < K ? best_i[k < K ? k : 0] : -1;
      }
    }
  }
}

template <int EPI, int K>
__global__ void __launch_bounds__(kMpThreads, 1)
match_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const MatchPairParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* a_res = smem;                                     // kchunks tiles of 16 KB
  uint8_t* ring = smem + (size_t)p.kchunks * kMpTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kMpStages * kMpTile);
  uint64_t* fullB = bars;                                    // leader's: both CTAs' bytes of the stage
  uint64_t* emptyB = bars + kMpStages;                       // per CTA: multicast commit
  uint64_t* tfull = bars + 2 * kMpStages;                    // [2] per CTA: multicast commit
  uint64_t* tempty = tfull + 2;                              // [2] leader: its own epilogue warps
  uint64_t* peer_tempty = tempty + 2;                        // [2] leader: the other CTA's epilogue warps (remote arrives)
  uint64_t* afull = peer_tempty + 2;                         // leader's: both query tiles of the item have landed
  uint64_t* afree = afull + 1;                               // per CTA: the item's MMAs are done with the query tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afree + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair_id = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kMpStages; ++s) {
      mbar_init(&fullB[s], 1);
      mbar_init(&emptyB[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 8);
      mbar_init(&peer_tempty[s], 8);
    }
    mbar_init(afull, 1);
    mbar_init(afree, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // the peer's barriers and TMEM exist before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      const uint32_t a_bytes = (uint32_t)p.kchunks * kMpTile;
      const uint32_t afull_leader = mapa_u32(afull, 0);
      int stage = 0, j = 0;
      uint32_t phase = 0;
      for (int round = 0, it; (it = item_of<EPI>(p, pair_id, n_pairs, round)) < p.items || round * n_pairs < p.items; ++round) {
        if (it >= p.items) continue;
        int m_pair, split, nt_begin, nt_end;
        item_range<EPI>(p, it, m_pair, split, nt_begin, nt_end);
        if (nt_begin >= nt_end) continue;                    // j counts the items that run (an empty one touches no barrier)
        if (j > 0) mbar_wait(afree, (uint32_t)(j - 1) & 1u); // the previous item's MMAs no longer read the query tile
        ++j;
        if (rank == 0) mbar_arrive_expect_tx(afull, 2u * a_bytes);
        const int row0 = m_pair * 256 + rank * 128;
        for (int cc = 0; cc < p.kchunks; ++cc)
          tma_load_4d_cg2(a_res + (size_t)cc * kMpTile, &tmA, afull_leader, cc * 64, 0, 0, row0);
        for (int nt = nt_begin; nt < nt_end; ++nt) {
          for (int cc = 0; cc < p.kchunks; ++cc) {
            mbar_wait(&emptyB[stage], phase ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&fullB[stage], 2u * kMpTile);
            tma_load_3d_cg2(ring + (size_t)stage * kMpTile, &tmB, mapa_u32(&fullB[stage], 0), cc * 64, nt * 256 + rank * 128, 0);
            if (++stage == kMpStages) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (pair leader) ================================
    if (rank == 0) {
      const uint32_t idesc = umma_idesc(256, 256, (uint32_t)p.is_bf16);
      const uint64_t a_desc0 = umma_smem_desc(smem_u32(a_res), 128);
      const uint64_t b_desc0 = umma_smem_desc(smem_u32(ring), 128);
      const uint64_t tile_inc = (uint64_t)(kMpTile >> 4);
      int stage = 0, j = 0, t = 0;
      uint32_t phase = 0;
      for (int round = 0, it; (it = item_of<EPI>(p, pair_id, n_pairs, round)) < p.items || round * n_pairs < p.items; ++round) {
        if (it >= p.items) continue;
        int m_pair, split, nt_begin, nt_end;
        item_range<EPI>(p, it, m_pair, split, nt_begin, nt_end);
        if (nt_begin >= nt_end) continue;
        mbar_wait(afull, (uint32_t)j & 1u);
        ++j;
        tc_fence_after();
        for (int nt = nt_begin; nt < nt_end; ++nt, ++t) {
          const int acc = t & 1;
          const uint32_t acc_phase = ((uint32_t)(t >> 1) & 1u) ^ 1u;
          mbar_wait(&tempty[acc], acc_phase);
          mbar_wait_cluster(&peer_tempty[acc], acc_phase);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
          for (int cc = 0; cc < p.kchunks; ++cc) {
            mbar_wait(&fullB[stage], phase);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da = a_desc0 + tile_inc * (uint64_t)cc, db = b_desc0 + tile_inc * (uint64_t)stage;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_cg2(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (cc | k) ? 1u : 0u);
              umma_commit_cg2(&emptyB[stage]);               // frees the stage in both CTAs when these MMAs retire
              if (cc == p.kchunks - 1) umma_commit_cg2(&tfull[acc]);
            }
            __syncwarp();
            if (++stage == kMpStages) stage = 0, phase ^= 1;
          }
        }
        if (elect_one()) umma_commit_cg2(afree);              // both producers may overwrite their query tiles
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue: two groups of four warps split the 256 columns ================
    const int group = (warp - 2) >> 2;
    const int qd = warp & 3;                      // TMEM lane quarter this warp may touch
    const int m = qd * 32 + lane;                 // accumulator row of this CTA == query row within its 128
    const int c_begin = group * 128;
    int t = 0;
    for (int round = 0, it; (it = item_of<EPI>(p, pair_id, n_pairs, round)) < p.items || round * n_pairs < p.items; ++round) {
      if (it >= p.items) continue;
      int m_pair, split, nt_begin, nt_end;
      item_range<EPI>(p, it, m_pair, split, nt_begin, nt_end);
      const int on = m_pair * 256 + rank * 128 + m;          // query row (local to the launch)
      const bool valid = on < p.q;
      if (EPI == MP_TOPK) {
        topk_item<K>(p, tmem_base, tfull, tempty, peer_tempty, rank, qd, lane, c_begin, nt_begin, nt_end, t, valid, on, split, group);
      } else {
        const long long gi = (long long)p.row_begin + on;   // global row of this thread
        for (int nt = nt_begin; nt < nt_end; ++nt, ++t) {
          const int acc = t & 1;
          mbar_wait(&tfull[acc], (uint32_t)(t >> 1) & 1u);
          tc_fence_after();
          const uint32_t t_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * 256);
          const long long cbase = (long long)nt * 256;
#pragma unroll 1
          for (int c0 = c_begin; c0 < c_begin + 128; c0 += 32) {
            uint32_t r[32];
            __syncwarp();
            tmem_ld32(t_addr + (uint32_t)c0, r);
            tmem_ld_wait();
            float gm[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              gm[j] = __uint_as_float(r[8 * j]);
#pragma unroll
              for (int i = 1; i < 8; ++i) gm[j] = fmaxf(gm[j], __uint_as_float(r[8 * j + i]));
            }
            const float mx = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
            // warp-uniform gates (votes): a candidate's exact re-check is done by the WHOLE warp, so every lane must
            // reach it; chunks / groups in which no lane has a coarse hit cost one vote each
            if (!__any_sync(0xFFFFFFFFu, valid && mx >= p.thr_coarse)) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (!__any_sync(0xFFFFFFFFu, valid && gm[j] >= p.thr_coarse)) continue;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const long long gj = cbase + c0 + 8 * j + i;           // same column for every lane
                const float v = __uint_as_float(r[8 * j + i]);
                uint32_t cand = __ballot_sync(0xFFFFFFFFu, valid && gj > gi && gj < p.g && v >= p.thr_coarse);
                while (cand) {
                  const int src = __ffs(cand) - 1;
                  cand &= cand - 1;
                  const long long gi_src = gi - lane + src;             // lanes hold consecutive rows
                  bool hit = true;
                  if (p.exact_rows) {
                    // exact fp32 dot of the two unit rows: 16-byte pieces interleaved over the lanes, then a butterfly sum
                    const float4* a = reinterpret_cast<const float4*>(p.exact_rows + (size_t)gi_src * p.exact_dim);
                    const float4* b = reinterpret_cast<const float4*>(p.exact_rows + (size_t)gj * p.exact_dim);
                    float d = 0.f;
                    for (int k = lane; k < (p.exact_dim >> 2); k += 32) {
                      const float4 x = __ldg(a + k), y = __ldg(b + k);
                      d = fmaf(x.x, y.x, d), d = fmaf(x.y, y.y, d), d = fmaf(x.z, y.z, d), d = fmaf(x.w, y.w, d);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, o);
                    hit = d >= p.thr_exact;
                  }
                  if (hit && lane == src) {
                    const unsigned long long slot = atomicAdd(p.pair_count, 1ull);
                    if ((long long)slot < p.max_pairs) p.pairs[slot] = (gi << 32) | gj;
                  }
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank) mbar_arrive_remote_relaxed(&peer_tempty[acc], 0);
            else mbar_arrive(&tempty[acc]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // no CTA leaves while its peer may still signal into its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int num_pairs_hw() {
  static int pairs = 0;
  if (!pairs) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    pairs = sms / 2;
  }
  return pairs;
}

// number of gallery ranges per query pair tile that keeps the pairs evenly busy: the candidate with the lowest
// (waves x tiles per item + a per-item overhead of about one tile for the query reload)
int match_pair_splits(int q, long long g) {
  const int m_pairs = (q + 255) / 256, n_tiles = (int)((g + 255) / 256), hw = num_pairs_hw();
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 64 && s <= n_tiles; ++s) {
    const int per = (n_tiles + s - 1) / s;
    if (per < 4 && s > 1) break;
    const int items = m_pairs * ((n_tiles + per - 1) / per);
    const double waves = (double)((items + hw - 1) / hw);
    const double cost = waves * (per + 1.0);
    if (cost < best_cost - 1e-9) best_cost = cost, best = s;
  }
  const int per = (n_tiles + best - 1) / best;
  return (n_tiles + per - 1) / per;
}

template <int EPI, int K>
static int launch_match_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, MatchPairParams& p, cudaStream_t stream) {
  const size_t smem = (size_t)(p.kchunks + kMpStages) * kMpTile + 1024 /*align*/ + 256 /*barriers*/;
  B2F_REQUIRE(smem <= 227 * 1024, "match_pair_kernel: dim too large for a resident query tile (%zu bytes of shared memory)", smem);
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] {
    attr_rc = cudaFuncSetAttribute(match_pair_kernel<EPI, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  B2F_CHECK_CUDA(attr_rc);
  int pairs = num_pairs_hw();
  if (pairs > p.items) pairs = p.items;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs), cfg.blockDim = dim3(kMpThreads), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  B2F_CHECK_CUDA(cudaLaunchKernelEx(&cfg, match_pair_kernel<EPI, K>, tmA, tmB, p));
  g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}

static int make_maps(CUtensorMap* tmA, CUtensorMap* tmB, const void* a, int rows_a, const void* b, long long rows_b, int dim,
                     int is_bf16) {
  {
    uint64_t dims[4] = {(uint64_t)dim, 1, 1, (uint64_t)rows_a};
    uint64_t str[3] = {(uint64_t)dim * 2, (uint64_t)dim * 2, (uint64_t)dim * 2};
    uint32_t box[4] = {64, 1, 1, 128};
    uint32_t es[4] = {1, 1, 1, 1};
    int rc = make_tmap(tmA, a, 4, dims, str, box, es, 128, is_bf16);
    if (rc) return rc;
  }
  uint64_t dims[3] = {(uint64_t)dim, (uint64_t)rows_b, 1};
  uint64_t str[2] = {(uint64_t)dim * 2, (uint64_t)rows_b * dim * 2};
  uint32_t box[3] = {64, 128, 1};
  uint32_t es[3] = {1, 1, 1};
  return make_tmap(tmB, b, 3, dims, str, box, es, 128, is_bf16);
}

// returns kPairDeclined when the first-generation kernel should take the call
constexpr int kPairDeclined = -1000;

// `keep` = how many candidates per list the caller will look at (<= topk, the stride of the lists): the running list in
// registers is that long, which is what the epilogue costs -- a top-1 query with a re-score margin of two extra
// candidates keeps 3, a top-5 search keeps all 8
int match_pair_topk(const void* queries, int q, const void* gallery, long long g, int dim, int dtype, int topk, int keep,
                    int n_splits, float* part_score, int* part_idx, cudaStream_t stream, int causal, long long causal_base) {
  if (((!g_match_pair || q <= 128) && !causal) || dim % 64 != 0 || (dim / 64 + kMpStages) * kMpTile + 1280 > 227 * 1024) return kPairDeclined;
  MatchPairParams p;
  memset(&p, 0, sizeof(p));
  p.q = q, p.g = g, p.kchunks = dim / 64, p.is_bf16 = dtype == 1;
  p.m_pairs = (q + 255) / 256, p.n_tiles = (int)((g + 255) / 256);
  const int per = (p.n_tiles + n_splits - 1) / n_splits;
  if ((p.n_tiles + per - 1) / per != n_splits) return kPairDeclined;     // the caller's partial buffers assume n_splits ranges
  p.splits = n_splits, p.items = p.m_pairs * n_splits;
  p.topk = topk, p.part_score = part_score, p.part_idx = part_idx;
  p.causal = causal ? 1 : 0, p.causal_base = causal_base;
  CUtensorMap tmA, tmB;
  int rc = make_maps(&tmA, &tmB, queries, q, gallery, g, dim, p.is_bf16);
  if (rc) return rc;
  if (keep <= 1) return launch_match_pair<MP_TOPK, 1>(tmA, tmB, p, stream);
  if (keep <= 3) return launch_match_pair<MP_TOPK, 3>(tmA, tmB, p, stream);
  return launch_match_pair<MP_TOPK, 8>(tmA, tmB, p, stream);
}

int match_pair_pairs(const void* emb16, int n, int dim, int dtype, int row_begin, int row_end, float thr_coarse, float thr_exact,
                     const float* emb_f32, long long* pairs, long long max_pairs, unsigned long long* pair_count,
                     cudaStream_t stream) {
  const int rows = row_end - row_begin;
  if (!g_match_pair || rows <= 128 || dim % 64 != 0 || (dim / 64 + kMpStages) * kMpTile + 1280 > 227 * 1024) return kPairDeclined;
  // the diagonal skip works in whole 256-row pair tiles: a row block that does not start on one would shift the tile
  // grid of its rows against the column tiles -- still correct (the j > i test is per element), only the skip is coarser
  MatchPairParams p;
  memset(&p, 0, sizeof(p));
  p.q = rows, p.g = n, p.kchunks = dim / 64, p.is_bf16 = dtype == 1;
  p.m_pairs = (rows + 255) / 256, p.n_tiles = (n + 255) / 256;
  // enough items for several rounds per pair: costs fall linearly with the row index (upper triangle), round-robin evens them out
  const int hw = num_pairs_hw();
  int splits = (8 * hw + p.m_pairs - 1) / p.m_pairs;
  if (splits > 16) splits = 16;
  if (splits > p.n_tiles) splits = p.n_tiles;
  if (splits < 1) splits = 1;
  p.splits = splits, p.items = p.m_pairs * splits;
  p.row_begin = row_begin, p.thr_coarse = thr_coarse, p.thr_exact = thr_exact;
  p.exact_rows = emb_f32, p.exact_dim = dim;
  p.pairs = pairs, p.max_pairs = max_pairs, p.pair_count = pair_count;
  CUtensorMap tmA, tmB;
  int rc = make_maps(&tmA, &tmB, reinterpret_cast<const uint8_t*>(emb16) + (size_t)row_begin * dim * 2, rows, emb16, n, dim, p.is_bf16);
  if (rc) return rc;
  return launch_match_pair<MP_PAIRS, 1>(tmA, tmB, p, stream);
}

}  // namespace b2f

// recommended number of gallery ranges for b2f_match_partial with q query rows against g gallery rows
extern "C" int b2f_match_plan(int q, long long g) {
  if (q <= 0 || g <= 0) return 1;
  if (!b2f::g_match_pair || q <= 128) {
    const int m_tiles = (q + 127) / 128;
    int want = (2 * 148 + m_tiles - 1) / m_tiles;
    if (want < 1) want = 1;
    return b2f_match_splits(g, want);
  }
  return b2f::match_pair_splits(q, g);
}
