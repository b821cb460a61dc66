// TMA (cp.async.bulk.tensor) + mbarrier helpers for sm_100a, and host-side tensor-map construction.
// The driver entry point cuTensorMapEncodeTiled is resolved through the runtime (no -lcuda link dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// orders prior generic-proxy smem accesses before subsequent async-proxy (TMA / tcgen05) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// 5-D tiled TMA load: coordinates innermost first
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16)
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*da_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

da_encode_tiled_fn da_get_encode_tiled();  // defined in conv3d_tma.cu; nullptr if the driver lacks the entry point

// fp32 planar [N][C][D][H][W] volume as a rank-5 map {W,H,D,C,N}; box {bx,by,bz,bc,1}.  Out-of-range box elements
// (spatial halo, channel padding) read as zero.  Requires W % 4 == 0 and a 16-byte aligned base.
inline int da_make_volume_map(CUtensorMap* map, const float* base, int N, int C, int D, int H, int W, int bx, int by, int bz,
                              int bc) {
  da_encode_tiled_fn enc = da_get_encode_tiled();
  if (!enc) { da_set_error("cuTensorMapEncodeTiled unavailable"); return DA_ERR_UNSUPPORTED; }
  cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)C, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * 4, (cuuint64_t)W * H * D * C * 4};
  cuuint32_t box[5] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, (cuuint32_t)bc, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { da_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return DA_ERR_UNSUPPORTED; }
  return DA_OK;
}

// The same volume with the channel dimension second: map dims {W, C, H, D, N}, box {bx, bc, by, 1, 1}, coordinates
// (x, c, y, z, n).  The box lands in shared memory as [by rows][bc channels][bx] -- threads whose lanes run along
// channels then read it with a pitch of bx floats instead of by*bx.
inline int da_make_volume_map_xcy(CUtensorMap* map, const float* base, int N, int C, int D, int H, int W, int bx, int bc, int by) {
  da_encode_tiled_fn enc = da_get_encode_tiled();
  if (!enc) { da_set_error("cuTensorMapEncodeTiled unavailable"); return DA_ERR_UNSUPPORTED; }
  cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)W * H * D * 4, (cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * C * 4};
  cuuint32_t box[5] = {(cuuint32_t)bx, (cuuint32_t)bc, (cuuint32_t)by, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { da_set_error("cuTensorMapEncodeTiled (x, c, y order) failed (%d)", (int)r); return DA_ERR_UNSUPPORTED; }
  return DA_OK;
}
