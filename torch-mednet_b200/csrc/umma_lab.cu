// UMMA descriptor laboratory: ONE CTA, a TMA-written shared-memory image, a caller-described sequence of
// tcgen05.mma instructions, accumulator dumped to global memory.  The tensor-core kernels of this library
// read their operands as shifted / chained windows of TMA-written tiles (27 taps of one halo for fprop,
// voxel-major "MN-major" windows for wgrad); those addressing modes are not spelled out by the PTX ISA
// text available offline, so every mode the kernels rely on is verified here on the device first
// (tests/test_tcgen05_gpu.py, DESIGN.md "UMMA descriptor experiments").
#include "common.cuh"
#include "tc_common.cuh"

namespace mednet {

struct LabArgs {
  int rows, row_bytes, M, N, ksteps;
  int a_off, a_lbo, a_sbo, a_mn, a_kstep;
  int b_off, b_lbo, b_sbo, b_mn, b_kstep;
  int a_fmt, b_fmt, box_rows, iters, nacc;
  float* out;
  long long* cycles;
};

__global__ void __launch_bounds__(128, 1) umma_lab_kernel(const __grid_constant__ CUtensorMap map, const LabArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const size_t img = (size_t)p.rows * p.row_bytes;
  uint64_t* bar = (uint64_t*)(smem + ((img + 1023) & ~(size_t)1023));
  uint32_t* slot = (uint32_t*)(bar + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int cols = 32;
  while (cols < p.N * (p.nacc < 0 ? 2 : p.nacc)) cols <<= 1;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::mbar_init(&bar[2], 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) {
    tc::tmem_alloc(slot, (uint32_t)cols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    tc::mbar_arrive_expect_tx(&bar[0], (uint32_t)img);
    for (int r = 0; r < p.rows; r += p.box_rows) tc::tma_load_2d(smem + (size_t)r * p.row_bytes, &map, &bar[0], 0, r);
    tc::mbar_wait(&bar[0], 0);
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_16(p.M, p.N, p.a_mn, p.b_mn, p.a_fmt, p.b_fmt);
    const uint32_t layout = p.row_bytes == 128 ? tc::SWZ_128B : (p.row_bytes == 64 ? tc::SWZ_64B : tc::SWZ_32B);
    const uint32_t base = tc::smem_u32(smem);
    const long long t0 = clock64();
    int rr = 0;
    if (p.nacc < 0) {
      // lean issue loop (timing only): descriptors built once, 4 k-steps unrolled, address advance = one 64-bit add;
      // -nacc independent accumulators are rotated through when nacc < -1
      const uint64_t da0 = tc::make_smem_desc(base + p.a_off, (uint32_t)p.a_lbo, (uint32_t)p.a_sbo, 0, layout);
      const uint64_t db0 = tc::make_smem_desc(base + p.b_off, (uint32_t)p.b_lbo, (uint32_t)p.b_sbo, 0, layout);
      const uint64_t ak = (uint64_t)(p.a_kstep >> 4), bk = (uint64_t)(p.b_kstep >> 4);
      const uint32_t dstep = p.nacc < -1 ? (uint32_t)p.N : 0u;
      for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::umma_bf16(tmem + (k & 1) * dstep, da0 + k * ak, db0 + k * bk, idesc, 1u);
      }
    } else
    for (int it = 0; it < p.iters; ++it) {
      for (int k = 0; k < p.ksteps; ++k, ++rr) {
        const uint64_t da = tc::make_smem_desc(base + p.a_off + k * p.a_kstep, (uint32_t)p.a_lbo, (uint32_t)p.a_sbo, 0, layout);
        const uint64_t db = tc::make_smem_desc(base + p.b_off + k * p.b_kstep, (uint32_t)p.b_lbo, (uint32_t)p.b_sbo, 0, layout);
        tc::umma_bf16(tmem + (uint32_t)((rr % p.nacc) * p.N), da, db, idesc, rr >= p.nacc);
      }
    }
    tc::umma_commit(&bar[1]);
    tc::mbar_wait(&bar[1], 0);
    if (p.cycles) { p.cycles[0] = clock64() - t0; p.cycles[2] = t0; p.cycles[3] = clock64(); }
  }
  if (p.nacc == -3 && threadIdx.x == 32) {
    // second issuer (timing only): same operands, its own accumulator, its own completion barrier
    tc::mbar_wait(&bar[0], 0);
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_16(p.M, p.N, p.a_mn, p.b_mn, p.a_fmt, p.b_fmt);
    const uint32_t layout = p.row_bytes == 128 ? tc::SWZ_128B : (p.row_bytes == 64 ? tc::SWZ_64B : tc::SWZ_32B);
    const uint32_t base = tc::smem_u32(smem);
    const uint64_t da0 = tc::make_smem_desc(base + p.a_off, (uint32_t)p.a_lbo, (uint32_t)p.a_sbo, 0, layout);
    const uint64_t db0 = tc::make_smem_desc(base + p.b_off, (uint32_t)p.b_lbo, (uint32_t)p.b_sbo, 0, layout);
    const uint64_t ak = (uint64_t)(p.a_kstep >> 4), bk = (uint64_t)(p.b_kstep >> 4);
    for (int it = 0; it < p.iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) tc::umma_bf16(tmem + (uint32_t)p.N, da0 + k * ak, db0 + k * bk, idesc, 1u);
    }
    tc::umma_commit(&bar[2]);
    tc::mbar_wait(&bar[2], 0);
    if (p.cycles) p.cycles[1] = clock64();
  }
  __syncwarp();
  tc::mbar_wait(&bar[1], 0);
  tc::tc_fence_after();
  const int m = warp * 32 + lane;
  for (int j = 0; j < p.N; j += 8) {
    uint32_t r[8];
    tc::tmem_ld_x8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)j, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) p.out[(size_t)m * p.N + j + i] = __uint_as_float(r[i]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)cols);
}

}  // namespace mednet

using namespace mednet;

extern "C" int mednet_umma_lab(const mednet_umma_lab_params* q, mednet_stream_t stream) {
  MEDNET_REQUIRE(q && q->g && q->out, MEDNET_EINVAL);
  MEDNET_REQUIRE(q->row_bytes == 128 || q->row_bytes == 64 || q->row_bytes == 32, MEDNET_EINVAL);
  MEDNET_REQUIRE(q->rows > 0 && (int64_t)q->rows * q->row_bytes <= 200 * 1024, MEDNET_EINVAL);
  MEDNET_REQUIRE((q->M == 128 || q->M == 64) && q->N >= 8 && q->N <= 256 && (q->N % 8) == 0 && q->ksteps > 0, MEDNET_EINVAL);
  if (!mednet_device_has_tcgen05()) return MEDNET_EUNSUPPORTED;
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return MEDNET_ENODRIVER;
  LabArgs p;
  p.rows = q->rows; p.row_bytes = q->row_bytes; p.M = q->M; p.N = q->N; p.ksteps = q->ksteps;
  p.a_off = q->a_off; p.a_lbo = q->a_lbo; p.a_sbo = q->a_sbo; p.a_mn = q->a_mn_major; p.a_kstep = q->a_kstep;
  p.b_off = q->b_off; p.b_lbo = q->b_lbo; p.b_sbo = q->b_sbo; p.b_mn = q->b_mn_major; p.b_kstep = q->b_kstep;
  p.a_fmt = q->a_fmt; p.b_fmt = q->b_fmt; p.out = q->out;
  p.iters = q->iters > 0 ? q->iters : 1; p.nacc = q->nacc != 0 ? q->nacc : 1;
  MEDNET_REQUIRE((p.nacc < 0 ? 2 : p.nacc) * p.N <= 512, MEDNET_EINVAL); p.cycles = (long long*)q->cycles;
  p.box_rows = 1;
  for (int b = 256; b >= 1; b >>= 1)
    if (q->rows % b == 0) { p.box_rows = b; break; }
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)(q->row_bytes / 2), (cuuint64_t)q->rows};
  cuuint64_t strides[1] = {(cuuint64_t)q->row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(q->row_bytes / 2), (cuuint32_t)p.box_rows};
  cuuint32_t estr[2] = {1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(q->g), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(q->row_bytes), CU_TENSOR_MAP_L2_PROMOTION_NONE,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return MEDNET_EUNSUPPORTED;
  const size_t smem = 1024 + (((size_t)q->rows * q->row_bytes + 1023) & ~(size_t)1023) + 64;
  cudaError_t e = cudaFuncSetAttribute(umma_lab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  umma_lab_kernel<<<1, 128, smem, stream>>>(map, p);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}
