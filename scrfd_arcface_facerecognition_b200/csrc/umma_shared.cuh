// Pieces shared by the tcgen05 kernels of libb2f (umma_conv.cu, conv_tile.cu): tensor-map encoding on the host,
// epilogue arithmetic and the unrolled MMA issue helpers on the device.
#pragma once
#include "b2f_common.cuh"

#include <mutex>

namespace b2f {

// ------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// rank-R map over 2-byte elements; dims/strides innermost first; strides in bytes for dims 1..R-1
static inline int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                     const uint32_t* box, const uint32_t* estr, int swizzle_bytes, int is_bf16) {
  EncodeTiledFn fn = get_encode_fn();
  B2F_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = estr[i];
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_b[i];
  CUresult r = fn(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank,
                  const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2F_REQUIRE(r == CUDA_SUCCESS,
              "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u] estr [%u %u %u %u] sw %d",
              (int)r, rank, (unsigned long long)gd[0], (unsigned long long)gd[1],
              (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0), bx[0], bx[1],
              rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0, es[0], es[1], rank > 2 ? es[2] : 0, rank > 3 ? es[3] : 0,
              swizzle_bytes);
  return 0;
}

// ------------------------------------------------------------------------------------------
// epilogue helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float v, int act, float slope) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v >= 0.f ? v : v * slope;
  if (act == 3) return 1.f / (1.f + expf(-v));
  return v;
}

__device__ __forceinline__ void load16_as_float(const void* p, int is_bf16, float (&f)[16]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (is_bf16) {
      __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&w[i]);
      f[2 * i] = __bfloat162float(h.x);
      f[2 * i + 1] = __bfloat162float(h.y);
    } else {
      __half2 h = *reinterpret_cast<__half2*>(&w[i]);
      f[2 * i] = __half2float(h.x);
      f[2 * i + 1] = __half2float(h.y);
    }
  }
}

// fp32 -> the 16-bit activation type and back (round to nearest even)
__device__ __forceinline__ float round16(float v, int is_bf16) {
  return is_bf16 ? __bfloat162float(__float2bfloat16_rn(v)) : __half2float(__float2half_rn(v));
}

__device__ __forceinline__ void store16(void* p, int dtype, const float (&f)[16]) {
  if (dtype == 2) {
    float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    return;
  }
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (dtype == 1) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    } else {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(w[0], w[1], w[2], w[3]);
  q[1] = make_uint4(w[4], w[5], w[6], w[7]);
}


// ------------------------------------------------------------------------------------------
// PTX: TMA tensor store, bulk async groups, named barriers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// TMA reduce-store: global[tile] += shared[tile], element-wise in the tensor map's data type (fp16 / bf16), done in L2
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// TMA reduce-store: global[tile] = max(global[tile], shared[tile]) element-wise (fp16 / bf16), done in L2
__device__ __forceinline__ void tma_reduce_max_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.max.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_named(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) pieces: the two CTAs of a cluster run one M = 256 MMA; each holds its own 128 rows of A and
// half of the N rows of B, so the weight traffic and the B operand reads per SM halve
// ------------------------------------------------------------------------------------------
// programmatic dependent launch: let the next kernel in the stream start its prologue while this one drains, and
// hold this kernel's first read of the previous kernel's output until that kernel has completed and flushed
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// same without the release fence (which costs a MEMBAR.ALL.GPU per arrive): for hand-offs that order nothing but
// "I am done reading TMEM", which tcgen05.wait::ld + tcgen05.fence::before_thread_sync already guarantee
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 23)) __trap();
  }
}
__device__ __forceinline__ uint32_t mapa_u32(const void* local, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
  return remote;
}
// TMA loads of a CTA pair: the data lands in this CTA's shared memory, the bytes are counted on `mbar_cluster`
// (a shared::cluster address -- the leader CTA's barrier), so the leader's MMA warp waits on ONE barrier per stage
__device__ __forceinline__ void tma_load_4d_cg2(void* dst, const CUtensorMap* m, uint32_t mbar_cluster, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_cg2(void* dst, const CUtensorMap* m, uint32_t mbar_cluster, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}


// K-steps of one pipeline stage, fully unrolled (the issuing thread's instruction stream is the critical path
// for narrow tiles: every extra instruction per tcgen05.mma shows up as tensor-pipe idle time)
template <int KSTEPS>
__device__ __forceinline__ void issue_stage(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t first_acc) {
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k)
    umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k == 0 ? first_acc : 1u);
}
__device__ __forceinline__ void issue_stage_rt(int ksteps, uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc,
                                               uint32_t first_acc) {
  if (ksteps == 4) issue_stage<4>(d_tmem, da, db, idesc, first_acc);
  else if (ksteps == 2) issue_stage<2>(d_tmem, da, db, idesc, first_acc);
  else issue_stage<1>(d_tmem, da, db, idesc, first_acc);
}

}  // namespace b2f
