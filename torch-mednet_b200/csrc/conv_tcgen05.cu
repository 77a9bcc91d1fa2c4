// 3x3x3 convolution (stride 1, pad 1) as an NDHWC implicit GEMM on the 5th-generation tensor cores.
// ref: midasmednet/unet/components.py:8-9 (nn.Conv3d) -- fprop, and dgrad through flipped/transposed
// packed weights (MEDNET_WPACK_TC_DGRAD).
//
// Mapping.  One CTA computes an output brick of TD x 16 x 8 voxels (d,h,w) by Ntile output channels:
// each d-plane of the brick is one UMMA accumulator (M = 128 voxels, N = Ntile) living in TMEM.
//   A operand : for every 64/32/16-channel chunk the (TD+2) x 18 x 10 HALO of the brick is staged ONCE
//               in shared memory by TMA (zero fill outside the volume = the conv padding).  The 27 taps
//               are 27 shifted windows of that halo: same bytes, different UMMA descriptor start address
//               (kh -> halo row pitch, kw -> one 128-byte voxel row, kd -> plane).  No im2col, no
//               27x re-read from L2.
//   B operand : weights [27][Nout][K], one [Ntile x chunk] K-major tile per tap streamed through a ring.
//   D         : TD accumulators x (1 or 2) stages in TMEM; the epilogue (bias, addend, activation,
//               bf16 pack, NDHWC store) of tile i overlaps the MMAs of tile i+1 (persistent CTAs).
// Warp roles: 0 = halo TMA producer, 1 = weight TMA producer, 2 = MMA issuer (+ TMEM alloc), 3..6 =
// epilogue (one TMEM lane quadrant each).  All hand-offs are mbarriers; halo planes are released as soon
// as their last kd phase has been issued so the next chunk's planes stream in under the current MMAs.
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "conv_impl.h"
#include "tc_common.cuh"

namespace mednet {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)ptr;
  }
  return fn;
}

// runtime-selectable A-operand addressing per swizzle width (index: 0 = 128 B rows, 1 = 64 B, 2 = 32 B);
// filled in by mednet_tcgen05_configure after the descriptor probe has run (see DESIGN.md
// "UMMA descriptor experiments").  Until then the tensor-core path reports "unsupported".
static int g_enabled[3] = {0, 0, 0};
static int g_dense_halo[3] = {0, 0, 0};        // 0: halo rows padded to a 16-voxel pitch, one TMA per (d,h) row
                                               // 1: dense 10-voxel pitch, one TMA box per halo plane
static int g_base_offset_mode[3] = {1, 1, 1};  // 0: base_offset = 0; 1: (addr >> 7) & 7; 2: (addr / row_bytes) & 7
static int g_dual_issue = 1;                  // mednet_tcgen05_set_option("dual_issue", 0|1)
static int g_kd_merge = 1;                    // mednet_tcgen05_set_option("kd_merge", 0|1): kd-merged wide-N MMAs (see TcConv::mt)
static int g_class_merge = 1;                 // mednet_tcgen05_set_option("class_merge", 0|1): the same for the parity classes of the
                                              // upsample conv (see TcConv::cmerge)
static inline int rb_class(int rb) { return rb == 128 ? 0 : rb == 64 ? 1 : 2; }

constexpr int HALO_H = 18, HALO_W = 10, TILE_H = 16, TILE_W = 8;
constexpr int MAX_PLANES = 6, MAX_BSTAGES = 8;
constexpr int NUM_THREADS = 256;     // warps: 0 halo TMA, 1 weight TMA, 2 MMA issuer (+TMEM alloc), 3..6 epilogue, 7 second MMA issuer

struct TcConv {
  int N, D, H, W, K, Nout;
  int tiles_w, tiles_h, tiles_d, tiles_n;
  int64_t num_tiles;
  int TD, Ntile, nchunks, RB, pitch, per_row, bo_mode;
  int plane_bytes, b_bytes, BS, acc_stages, tmem_cols;
  int dual;                    // 1: two MMA-issuing threads, each owning half of the brick's d-planes
  // kd-merged MMAs (plain conv, output tiles <= 128 channels).  The window (kh,kw) of halo plane p is the A operand of
  // up to three taps -- kd = 0,1,2 feeding the accumulators of d-planes p, p-1, p-2 -- so with the accumulators laid out
  // in TMEM in DECREASING d order and the three kd weight tiles of one (kh,kw) adjacent in shared memory, ONE MMA with
  // N = mt*Ntile computes mt taps while reading the 4 KB A window once instead of mt times: the 64-channel-output layers
  // go from shared-memory bound (A 4 KB + B 2 KB = 48 clk for 32 clk of math per tap) to math bound (10 KB = 80 clk for
  // 96 clk), and one issuing thread suffices (one instruction per >= 64 clk of math).  mt = taps per MMA (1 = off).
  int mt;
  int b_tile;                  // bytes of one [Ntile x chunk] weight tile (a B stage holds 3 of them in kd-merged mode)
  // Class-merged mode (parity classes of the conv over a nearest-upsampled input, MEDNET_GATHER_UPCONV_*): every class has
  // exactly 2 x 2 x 2 taps, kd in {kd_lo, kd_lo + 1}.  mt = 2: halo plane kd_lo + j feeds accumulators j (tap kd_lo) and
  // j - 1 (tap kd_lo + 1) with ONE MMA of N = 2 * Ntile; only the TD + 1 planes a class needs are staged (slot j = plane
  // kd_lo + j) and only its four (kh,kw) windows are visited.  The plain class path (one MMA per tap and plane, all TD + 2
  // planes, planes released per kd phase) ran at 57 % of its shared-memory bound: with a single halo set 4 of the 6 planes
  // of the next chunk could not be requested before the current chunk's last MMA.
  int cmerge, nplanes;
  int a_stages;                // halo sets in shared memory.  The kd-merged order needs every plane until the last (kh,kw)
                               // window, so it works on 32-channel chunks with TWO halo sets: the next chunk streams in
                               // under the current chunk's MMAs (with one set the tensor pipe idles for a halo load per chunk)
  // Parity classes of the stride-2 transposed convolution (components.py:259-264).  A plain conv has one class with all
  // 27 taps.  Transposed fprop: 8 OUTPUT classes (output voxel = 2j + parity), each a sub-conv over the input grid with
  // 1/2/4/8 taps; transposed dgrad: 8 INPUT classes (A read at 2j + parity through a stride-2 TMA map) accumulated into
  // the same TMEM tile.  Window tap index = offset + 1 per axis, exactly as for the plain conv.
  int ncls_in, ncls_out, in_scale, out_scale;
  int OD, OH, OW;              // dims of the output tensor (= out_scale * grid)
  uint32_t tapmask[8];
  int act;
  float act_param;
  const float* bias;
  const bf16* addend;
  bf16* y;
  int f32_io;                  // 0: bf16 output / addend; 1: y is fp32 (unrounded partial sum); 2: addend is fp32 (see mednet_conv3d_params)
};

struct TileCoord {
  int n, d0, h0, w0, n0, cls;
};
__device__ __forceinline__ TileCoord decode_tile(const TcConv& p, int64_t t) {
  TileCoord c;
  c.n0 = (int)(t % p.tiles_n) * p.Ntile; t /= p.tiles_n;
  c.cls = (int)(t % p.ncls_out); t /= p.ncls_out;     // the output classes of one brick run together: they share the halo in L2
  c.w0 = (int)(t % p.tiles_w) * TILE_W; t /= p.tiles_w;
  c.h0 = (int)(t % p.tiles_h) * TILE_H; t /= p.tiles_h;
  c.d0 = (int)(t % p.tiles_d) * p.TD;
  c.n = (int)(t / p.tiles_d);
  return c;
}

// mbarrier wait that, in the profiling instantiation of the kernel, adds the cycles spent waiting to `acc`
template <bool PROF>
__device__ __forceinline__ void pwait(uint64_t* bar, uint32_t parity, long long& acc) {
  if (PROF) {
    const long long t0 = clock64();
    tc::mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    tc::mbar_wait(bar, parity);
  }
}

// descriptor-address offset (16-byte units) of window (kh, kw) = g / 3, g % 3 inside a halo plane
__device__ __forceinline__ uint32_t tapoff9(int g, uint32_t tap_h, uint32_t tap_w) {
  const uint32_t kh = (uint32_t)g / 3u, kw = (uint32_t)g - kh * 3u;
  return kh * tap_h + kw * tap_w;
}

template <int ACT>
__device__ __forceinline__ float act_apply_t(float x, float a) {
  if (ACT == MEDNET_ACT_RELU) return x > 0.f ? x : 0.f;
  if (ACT == MEDNET_ACT_LEAKY) return x > 0.f ? x : a * x;
  if (ACT == MEDNET_ACT_ELU) return x > 0.f ? x : expm1f(x);
  return x;
}

// Epilogue of one CTA (warps 3..6, one TMEM lane quadrant each): TMEM -> registers -> (+bias, +addend, activation) -> bf16
// NDHWC rows.  Accumulator row m = voxel (h, w) of the brick plane.
// F32IO: 1 = the output is written as fp32 (a partial sum another launch completes: no rounding in between),
//        2 = the addend is such an fp32 partial sum.
template <int ACT, bool PROF, int F32IO = 0>
__device__ __forceinline__ void epilogue_loop(const TcConv& p, uint32_t tmem_base, uint64_t* acc_full, uint64_t* acc_empty,
                                              int warp, int lane, long long* prof) {
  long long w_full = 0;
  const long long t_begin = PROF ? clock64() : 0;
  const int q = warp & 3;                 // TMEM lane quadrant this warp may access
  const int m = q * 32 + lane;
  const int hh = m >> 3, ww = m & 7;
  uint32_t tcount = 0;
  for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tcount) {
    const TileCoord tc_ = decode_tile(p, t);
    const uint32_t as = tcount % (uint32_t)p.acc_stages, aph = (tcount / (uint32_t)p.acc_stages) & 1u;
    pwait<PROF>(&acc_full[as], aph, w_full);
    tc::tc_fence_after();
    const int h = tc_.h0 + hh, w = tc_.w0 + ww;
    const int oh = p.out_scale * h + ((tc_.cls >> 1) & 1), ow = p.out_scale * w + (tc_.cls & 1);
    for (int dz = 0; dz < p.TD; ++dz) {
      const int d = tc_.d0 + dz;
      const bool inb = d < p.D && h < p.H && w < p.W;
      const int od = p.out_scale * d + (tc_.cls >> 2);
      const int64_t vox = (((int64_t)tc_.n * p.OD + od) * p.OH + oh) * p.OW + ow;
      bf16* yrow = p.y + vox * p.Nout + tc_.n0;
      const bf16* arow = p.addend ? p.addend + vox * p.Nout + tc_.n0 : nullptr;
      const uint32_t dcol = p.mt > 1 ? (uint32_t)(p.TD - 1 - dz) : (uint32_t)dz;     // kd-merged mode: decreasing d order
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (as * (uint32_t)p.TD + dcol) * (uint32_t)p.Ntile;
      for (int jb = 0; jb < p.Ntile; jb += 64) {
      // fp32 partial-sum addend: all loads of a 64-channel block are issued before the first TMEM read (16 x 16 bytes in
      // flight per thread; one load per use made the epilogue latency-bound and the MMA issuer wait for it)
      float4 pre[4][4];
      if (F32IO == 2 && inb) {
        const float4* aq = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.addend) + vox * p.Nout + tc_.n0 + jb);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (jb + 16 * g < p.Ntile) {
#pragma unroll
            for (int i = 0; i < 4; ++i) pre[g][i] = aq[4 * g + i];
          }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int j = jb + 16 * g;
        if (j >= p.Ntile) break;
        uint32_t r[16];
        tc::tmem_ld_x16(taddr + (uint32_t)j, r);
        tc::tmem_ld_wait();
        if (p.mt > 1) tc::tmem_st_x16_zero(taddr + (uint32_t)j);     // kd-merged mode: every MMA accumulates
        if (inb) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
          if (p.bias != nullptr) {
            const float4* bq = reinterpret_cast<const float4*>(p.bias + tc_.n0 + j);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = __ldg(bq + i);
              v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
            }
          }
          if (F32IO == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 a4 = pre[g][i];
              v[4 * i] += a4.x; v[4 * i + 1] += a4.y; v[4 * i + 2] += a4.z; v[4 * i + 3] += a4.w;
            }
          } else if (arow != nullptr) {
            float a0[8], a1[8];
            load_vec<bf16, 8>(arow + j, a0);
            load_vec<bf16, 8>(arow + j + 8, a1);
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] += a0[i]; v[8 + i] += a1[i]; }
          }
          if (F32IO == 1) {
            float4* yq = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + vox * p.Nout + tc_.n0 + j);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              yq[i] = make_float4(act_apply_t<ACT>(v[4 * i], p.act_param), act_apply_t<ACT>(v[4 * i + 1], p.act_param),
                                  act_apply_t<ACT>(v[4 * i + 2], p.act_param), act_apply_t<ACT>(v[4 * i + 3], p.act_param));
          } else {
            float o0[8], o1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              o0[i] = act_apply_t<ACT>(v[i], p.act_param);
              o1[i] = act_apply_t<ACT>(v[8 + i], p.act_param);
            }
            store_vec<bf16, 8>(yrow + j, o0);
            store_vec<bf16, 8>(yrow + j + 8, o1);
          }
        }
      }
      }
    }
    if (p.mt > 1) tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&acc_empty[as]);
  }
  if (PROF && warp == 3 && lane == 0) { prof[4] = clock64() - t_begin; prof[5] = w_full; }
}

// PROF: per-CTA cycle counters written to prof_out[blockIdx.x * 8 + i] (mednet_tcgen05_set_option("conv_profile", 1)):
//   0 MMA issuer total, 1 its wait for halo planes, 2 for weight stages, 3 for a free accumulator stage,
//   4 epilogue warp total, 5 its wait for a finished accumulator, 6 halo producer wait for free planes, 7 weight producer
//   wait for free stages
template <bool PROF>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcConv p,
                long long* __restrict__ prof_out) {
  long long* prof = PROF ? prof_out + (size_t)blockIdx.x * 8 : nullptr;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;
  uint8_t* b_base = a_base + (size_t)p.a_stages * p.nplanes * p.plane_bytes;
  uint64_t* bars = (uint64_t*)(b_base + (size_t)p.BS * p.b_bytes);
  uint64_t* a_full = bars;                       // [a_stages][MAX_PLANES]
  uint64_t* a_empty = bars + 2 * MAX_PLANES;
  uint64_t* b_full = bars + 4 * MAX_PLANES;
  uint64_t* b_empty = b_full + MAX_BSTAGES;
  uint64_t* acc_full = b_empty + MAX_BSTAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nplanes = p.nplanes;

  if (threadIdx.x == 0) {
    const uint32_t nissue = p.dual ? 2u : 1u;         // every issuer commits to the barriers the MMAs release
    for (int st = 0; st < p.a_stages; ++st)
      for (int i = 0; i < nplanes; ++i) {
        tc::mbar_init(&a_full[st * MAX_PLANES + i], 1);
        tc::mbar_init(&a_empty[st * MAX_PLANES + i], nissue);
      }
    for (int i = 0; i < p.BS; ++i) { tc::mbar_init(&b_full[i], 1); tc::mbar_init(&b_empty[i], nissue); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], nissue); tc::mbar_init(&acc_empty[i], 4); }
    tc::fence_barrier_init();
    tc::tma_prefetch_desc(&map_x);
    tc::tma_prefetch_desc(&map_w);
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int KC = p.RB / 2;
  if (p.mt > 1) {
    // kd-merged mode: accumulators start at zero (the epilogue re-zeroes what it reads)
    if (warp >= 3 && warp <= 6) {
      for (int col = 0; col < p.tmem_cols; col += 16) tc::tmem_st_x16_zero(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col);
      tc::tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  }

  if (warp == 0) {
    // ===================== halo (A) producer =====================
    if (tc::elect_one()) {
      uint32_t ait = 0;
      long long w_empty = 0;
      for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const TileCoord tc_ = decode_tile(p, t);
        for (int ci = 0; ci < p.ncls_in; ++ci) {
          const int s = p.in_scale;
          // class-merged mode: slot j holds plane kd_lo + j of the class (the other plane of the TD + 2 halo is never read)
          const int kd_lo = (p.cmerge && !(p.tapmask[p.ncls_out > 1 ? tc_.cls : ci] & 0x1ffu)) ? 1 : 0;
          const int cw = s * (tc_.w0 - 1) + (ci & 1), ch = s * (tc_.h0 - 1) + ((ci >> 1) & 1),
                    cd0 = s * (tc_.d0 - 1 + kd_lo) + (ci >> 2);
          for (int c = 0; c < p.nchunks; ++c, ++ait) {
            const uint32_t ast = ait % (uint32_t)p.a_stages, aph = (ait / (uint32_t)p.a_stages) & 1u;
            uint64_t* full = a_full + ast * MAX_PLANES;
            uint64_t* empty = a_empty + ast * MAX_PLANES;
            for (int pl = 0; pl < nplanes; ++pl) {
              pwait<PROF>(&empty[pl], aph ^ 1u, w_empty);
              tc::mbar_arrive_expect_tx(&full[pl], (uint32_t)(HALO_H * HALO_W * p.RB));
              uint8_t* dst = a_base + (size_t)(ast * nplanes + pl) * p.plane_bytes;
              if (p.per_row) {
                for (int ph = 0; ph < HALO_H; ++ph)
                  tc::tma_load_5d(dst + (size_t)ph * p.pitch * p.RB, &map_x, &full[pl], c * KC, cw, ch + s * ph,
                                  cd0 + s * pl, tc_.n);
              } else {
                tc::tma_load_5d(dst, &map_x, &full[pl], c * KC, cw, ch, cd0 + s * pl, tc_.n);
              }
            }
          }
        }
      }
      if (PROF) prof[6] = w_empty;
    }
  } else if (warp == 1) {
    // ===================== weight (B) producer =====================
    if (tc::elect_one()) {
      uint32_t bit = 0;
      long long w_empty = 0;
      for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const TileCoord tc_ = decode_tile(p, t);
        for (int ci = 0; ci < p.ncls_in; ++ci) {
          const int wcls = p.ncls_out > 1 ? tc_.cls : ci;          // weight set / tap mask of this (tile, input class)
          const uint32_t mask = p.tapmask[wcls];
          for (int c = 0; c < p.nchunks; ++c) {
            if (p.cmerge) {
              // class-merged mode: one stage = the two kd tiles of one of the class's four (kh,kw) windows
              const int kd_lo = (mask & 0x1ffu) ? 0 : 1;
              const uint32_t gm = (mask >> (kd_lo * 9)) & 0x1ffu;
              for (int g = 0; g < 9; ++g) {
                if (!((gm >> g) & 1u)) continue;
                const uint32_t st = bit % (uint32_t)p.BS, ph = (bit / (uint32_t)p.BS) & 1u;
                ++bit;
                pwait<PROF>(&b_empty[st], ph ^ 1u, w_empty);
                tc::mbar_arrive_expect_tx(&b_full[st], (uint32_t)(2 * p.b_tile));
                for (int k2 = 0; k2 < 2; ++k2)
                  tc::tma_load_2d(b_base + (size_t)st * p.b_bytes + (size_t)k2 * p.b_tile, &map_w, &b_full[st], c * KC,
                                  (wcls * 27 + (kd_lo + k2) * 9 + g) * p.Nout + tc_.n0);
              }
              continue;
            }
            if (p.mt > 1) {
              // kd-merged mode: one stage = the three kd tiles of one (kh,kw), adjacent in shared memory
              for (int g = 0; g < 9; ++g) {
                const uint32_t st = bit % (uint32_t)p.BS, ph = (bit / (uint32_t)p.BS) & 1u;
                ++bit;
                pwait<PROF>(&b_empty[st], ph ^ 1u, w_empty);
                tc::mbar_arrive_expect_tx(&b_full[st], (uint32_t)(3 * p.b_tile));
                for (int kd = 0; kd < 3; ++kd)
                  tc::tma_load_2d(b_base + (size_t)st * p.b_bytes + (size_t)kd * p.b_tile, &map_w, &b_full[st], c * KC,
                                  (kd * 9 + g) * p.Nout + tc_.n0);
              }
              continue;
            }
            for (int tap = 0; tap < 27; ++tap) {
              if (!((mask >> tap) & 1u)) continue;
              const uint32_t st = bit % (uint32_t)p.BS, ph = (bit / (uint32_t)p.BS) & 1u;
              ++bit;
              pwait<PROF>(&b_empty[st], ph ^ 1u, w_empty);
              tc::mbar_arrive_expect_tx(&b_full[st], (uint32_t)(p.Ntile * p.RB));
              tc::tma_load_2d(b_base + (size_t)st * p.b_bytes, &map_w, &b_full[st], c * KC,
                              (wcls * 27 + tap) * p.Nout + tc_.n0);
            }
          }
        }
      }
      if (PROF && warp == 1) prof[7] = w_empty;
    }
  } else if (warp == 2 || warp == 7) {
    // ===================== MMA issuer(s) =====================
    // A single issuing thread sustains one M=128 MMA per ~55 clk whatever N is (measured, csrc/umma_lab.cu), while the
    // tensor pipe needs only N/2 clk and shared memory (4 KB + 32 N B)/128 clk per MMA.  For N <= 96 a second issuer
    // (warp 7) takes half of the brick's d-planes (its own accumulators), which moves the bound from the issue rate to
    // the shared-memory read rate (48 clk at N = 64).
    // The issuing thread is a serial instruction stream: every instruction between two tcgen05.mma costs
    // ~4-5 clk of dependent latency, and an M=128 MMA only occupies the tensor pipe for N/2 clk (32 clk at
    // N=64).  So the descriptors are built ONCE; per MMA only the 32-bit address word is advanced by
    // register-resident offsets (measured: 111 clk/MMA with per-MMA descriptor construction, 55 clk lean).
    const int issuer = warp == 2 ? 0 : 1;
    const int TDI = p.dual ? p.TD / 2 : p.TD;           // planes per issuer (kernel-uniform loop bound)
    if ((issuer == 0 || p.dual) && tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_bf16(128, p.Ntile, 0, 0);
      const uint32_t layout = p.RB == 128 ? tc::SWZ_128B : (p.RB == 64 ? tc::SWZ_64B : tc::SWZ_32B);
      const uint64_t da_base = tc::make_smem_desc(tc::smem_u32(a_base), 16, (uint32_t)(p.pitch * p.RB), 0, layout);
      const uint64_t db_base = tc::make_smem_desc(tc::smem_u32(b_base), 16, (uint32_t)(8 * p.RB), 0, layout);
      const uint32_t a_hi = (uint32_t)(da_base >> 32), b_hi = (uint32_t)(db_base >> 32);
      const uint32_t plane16 = (uint32_t)p.plane_bytes >> 4, b16 = (uint32_t)p.b_bytes >> 4;
      const uint32_t a_lo0 = (uint32_t)da_base + (uint32_t)(issuer * TDI) * plane16, b_lo0 = (uint32_t)db_base;
      const uint32_t d_issuer = (uint32_t)(issuer * TDI * p.Ntile);
      uint32_t tapoff[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) tapoff[i] = (uint32_t)(((i / 3) * p.pitch + (i % 3)) * p.RB) >> 4;
      const int ksteps = p.RB / 32;
      // kd-merged mode constants
      const uint32_t ntile = (uint32_t)p.Ntile, tile16 = (uint32_t)p.b_tile >> 4, dcol_top = (uint32_t)((p.TD - 1) * p.Ntile);
      const uint32_t id1 = idesc, id2 = tc::make_idesc_bf16(128, 2 * p.Ntile <= 256 ? 2 * p.Ntile : 8, 0, 0),
                     id3 = tc::make_idesc_bf16(128, 3 * p.Ntile <= 256 ? 3 * p.Ntile : 8, 0, 0);
      const uint32_t tap_h = (uint32_t)(p.pitch * p.RB) >> 4, tap_w = (uint32_t)p.RB >> 4;
      uint32_t ait = 0, bit = 0, tcount = 0;
      uint32_t bst = 0, bph = 0;                       // B ring position (stage, parity) without div/mod
      long long w_a = 0, w_b = 0, w_acc = 0;
      const long long t_begin = PROF ? clock64() : 0;
      for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tcount) {
        const uint32_t as = tcount % (uint32_t)p.acc_stages, aph = (tcount / (uint32_t)p.acc_stages) & 1u;
        pwait<PROF>(&acc_empty[as], aph ^ 1u, w_acc);
        tc::tc_fence_after();
        const uint32_t d_tile = tmem_base + as * (uint32_t)(p.TD * p.Ntile) + d_issuer;
        const int out_cls = p.ncls_out > 1 ? decode_tile(p, t).cls : -1;
        uint32_t fresh = 1u;                              // the first MMA into each accumulator overwrites it
        for (int ci = 0; ci < p.ncls_in; ++ci) {
          const uint32_t mask = p.tapmask[out_cls >= 0 ? out_cls : ci];
          for (int c = 0; c < p.nchunks; ++c, ++ait) {
            const uint32_t ast = ait % (uint32_t)p.a_stages, aph = (ait / (uint32_t)p.a_stages) & 1u;
            uint64_t* a_full_s = a_full + ast * MAX_PLANES;
            uint64_t* a_empty_s = a_empty + ast * MAX_PLANES;
            const uint32_t a_lo_s = a_lo0 + ast * (uint32_t)nplanes * plane16;
#define KDM_PLANE(PL, DCOL, BLO, IDESC)                                                          \
                {                                                                                \
                  if (first_g) { pwait<PROF>(&a_full_s[PL], aph, w_a); tc::tc_fence_after(); }   \
                  const uint32_t d_ = d_tile + (DCOL);                                           \
                  if (ksteps == 4) {                                                             \
                    tc::umma_bf16_lohi(d_, a, a_hi, (BLO), b_hi, (IDESC), 1u);                   \
                    tc::umma_bf16_lohi(d_, a + 2, a_hi, (BLO) + 2, b_hi, (IDESC), 1u);           \
                    tc::umma_bf16_lohi(d_, a + 4, a_hi, (BLO) + 4, b_hi, (IDESC), 1u);           \
                    tc::umma_bf16_lohi(d_, a + 6, a_hi, (BLO) + 6, b_hi, (IDESC), 1u);           \
                  } else if (ksteps == 2) {                                                      \
                    tc::umma_bf16_lohi(d_, a, a_hi, (BLO), b_hi, (IDESC), 1u);                   \
                    tc::umma_bf16_lohi(d_, a + 2, a_hi, (BLO) + 2, b_hi, (IDESC), 1u);           \
                  } else {                                                                       \
                    tc::umma_bf16_lohi(d_, a, a_hi, (BLO), b_hi, (IDESC), 1u);                   \
                  }                                                                              \
                  if (last_g) tc::umma_commit(&a_empty_s[PL]);                                   \
                  a += plane16;                                                                  \
                }
            if (p.cmerge) {
              // ---- class-merged issue order (see TcConv::cmerge): the class's four (kh,kw) windows outer, its TD + 1 halo
              // planes inner.  Slot j = plane kd_lo + j: tap kd_lo -> accumulator j, tap kd_lo + 1 -> accumulator j - 1, i.e.
              // columns (TD-1-j)*Ntile and the next Ntile (decreasing-d layout); all MMAs accumulate (zeroed accumulators)
              const uint32_t kd_lo = (mask & 0x1ffu) ? 0u : 1u;
              const uint32_t gm = (mask >> (kd_lo * 9u)) & 0x1ffu;
              const int g_first = __ffs((int)gm) - 1, g_last = 31 - __clz((int)gm);
              for (int g = g_first; g <= g_last; ++g) {
                if (!((gm >> g) & 1u)) continue;
                pwait<PROF>(&b_full[bst], bph, w_b);
                tc::tc_fence_after();
                const uint32_t b0 = b_lo0 + bst * b16, b1 = b0 + tile16;
                uint32_t a = a_lo_s + tapoff9(g, tap_h, tap_w);
                const bool first_g = g == g_first, last_g = g == g_last;
                KDM_PLANE(0, dcol_top, b0, id1)
                for (int j = 1; j < p.TD; ++j) KDM_PLANE(j, dcol_top - (uint32_t)j * ntile, b0, id2)
                KDM_PLANE(p.TD, 0u, b1, id1)
                tc::umma_commit(&b_empty[bst]);
                if (++bst == (uint32_t)p.BS) { bst = 0; bph ^= 1u; }
              }
              continue;
            }
            if (p.mt > 1) {
              // ---- kd-merged issue order: (kh,kw) outer, halo plane inner; see TcConv::mt.  Every accumulator column is
              // zero when its tile starts (the epilogue clears what it has read), so all MMAs accumulate.  Planes 0 / 1
              // reach 1 / 2 d-planes, the interior planes 3, planes TD / TD+1 again 2 / 1:
              //   plane pl, taps kd_lo..kd_hi -> accumulators dz = pl-kd_lo .. pl-kd_hi = columns (TD-1-dz)*Ntile ascending
              for (int g = 0; g < 9; ++g) {
                pwait<PROF>(&b_full[bst], bph, w_b);
                tc::tc_fence_after();
                const uint32_t b0 = b_lo0 + bst * b16, b1 = b0 + tile16, b2 = b1 + tile16;
                uint32_t a = a_lo_s + tapoff9(g, tap_h, tap_w);
                const bool first_g = g == 0, last_g = g == 8;
                KDM_PLANE(0, dcol_top, b0, id1)
                KDM_PLANE(1, dcol_top - ntile, b0, id2)
                for (int pl = 2; pl < p.TD; ++pl) KDM_PLANE(pl, dcol_top - (uint32_t)pl * ntile, b0, id3)
                KDM_PLANE(p.TD, 0u, b1, id2)
                KDM_PLANE(p.TD + 1, 0u, b2, id1)
                tc::umma_commit(&b_empty[bst]);
                if (++bst == (uint32_t)p.BS) { bst = 0; bph ^= 1u; }
              }
              continue;
            }
            for (int kd = 0; kd < 3; ++kd) {
              if (kd == 0) {
                for (int pl = 0; pl < p.TD; ++pl) pwait<PROF>(&a_full_s[pl], aph, w_a);
              } else {
                pwait<PROF>(&a_full_s[p.TD - 1 + kd], aph, w_a);
              }
              tc::tc_fence_after();
              const uint32_t a_kd = a_lo_s + (uint32_t)kd * plane16;
              const uint32_t mask_kd = (mask >> (kd * 9)) & 0x1ffu;
#pragma unroll
              for (int khw = 0; khw < 9; ++khw) {
                if (!((mask_kd >> khw) & 1u)) continue;
                pwait<PROF>(&b_full[bst], bph, w_b);
                tc::tc_fence_after();
                const uint32_t b_lo = b_lo0 + bst * b16;
                uint32_t a_lo = a_kd + tapoff[khw];
                uint32_t d_tmem = d_tile;
                const uint32_t first = fresh ? 0u : 1u;
                fresh = 0u;
                for (int dz = 0; dz < TDI; ++dz) {
                  if (ksteps == 4) {
                    tc::umma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, first);
                    tc::umma_bf16_lohi(d_tmem, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
                    tc::umma_bf16_lohi(d_tmem, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
                    tc::umma_bf16_lohi(d_tmem, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
                  } else if (ksteps == 2) {
                    tc::umma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, first);
                    tc::umma_bf16_lohi(d_tmem, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
                  } else {
                    tc::umma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, first);
                  }
                  a_lo += plane16;
                  d_tmem += (uint32_t)p.Ntile;
                }
                tc::umma_commit(&b_empty[bst]);
                if (++bst == (uint32_t)p.BS) { bst = 0; bph ^= 1u; }
              }
              // planes whose last tap phase is this kd can be refilled for the next chunk
              if (kd < 2) {
                tc::umma_commit(&a_empty_s[kd]);
              } else {
                for (int pl = 2; pl < nplanes; ++pl) tc::umma_commit(&a_empty_s[pl]);
              }
            }
          }
        }
        tc::umma_commit(&acc_full[as]);
      }
#undef KDM_PLANE
      (void)bit;
      if (PROF && issuer == 0) { prof[0] = clock64() - t_begin; prof[1] = w_a; prof[2] = w_b; prof[3] = w_acc; }
    }
  } else {
    // ===================== epilogue =====================
    // the activation is a template parameter: a run-time switch inside the per-element code is compiled to either
    // selects or an indirect branch per element depending on unrelated details of the kernel (measured: 2x slower
    // epilogue-bound layers), so it is dispatched once per CTA here
    if (p.f32_io == 1) {           // fp32 partial sum out: by contract without activation
      epilogue_loop<MEDNET_ACT_NONE, PROF, 1>(p, tmem_base, acc_full, acc_empty, warp, lane, prof);
    } else if (p.f32_io == 2) {    // fp32 partial sum in: the layers that use it end in ReLU or nothing
      if (p.act == MEDNET_ACT_RELU) epilogue_loop<MEDNET_ACT_RELU, PROF, 2>(p, tmem_base, acc_full, acc_empty, warp, lane, prof);
      else if (p.act == MEDNET_ACT_LEAKY) epilogue_loop<MEDNET_ACT_LEAKY, PROF, 2>(p, tmem_base, acc_full, acc_empty, warp, lane, prof);
      else if (p.act == MEDNET_ACT_ELU) epilogue_loop<MEDNET_ACT_ELU, PROF, 2>(p, tmem_base, acc_full, acc_empty, warp, lane, prof);
      else epilogue_loop<MEDNET_ACT_NONE, PROF, 2>(p, tmem_base, acc_full, acc_empty, warp, lane, prof);
    } else
    switch (p.act) {
      case MEDNET_ACT_RELU: epilogue_loop<MEDNET_ACT_RELU, PROF>(p, tmem_base, acc_full, acc_empty, warp, lane, prof); break;
      case MEDNET_ACT_LEAKY: epilogue_loop<MEDNET_ACT_LEAKY, PROF>(p, tmem_base, acc_full, acc_empty, warp, lane, prof); break;
      case MEDNET_ACT_ELU: epilogue_loop<MEDNET_ACT_ELU, PROF>(p, tmem_base, acc_full, acc_empty, warp, lane, prof); break;
      default: epilogue_loop<MEDNET_ACT_NONE, PROF>(p, tmem_base, acc_full, acc_empty, warp, lane, prof); break;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host
static int g_ntile_max = 128;                 // mednet_tcgen05_set_option("ntile_max", 128|256)
// Output-channel tile.  Tiles of <= 128 channels keep TWO accumulator stages in TMEM (2 x TD x Ntile <= 512 columns), so
// the epilogue of one tile overlaps the MMAs of the next (with 192 / 256-channel tiles the MMA issuer waited for the
// epilogue for up to 29 % of the kernel, tools/conv_profile.py), and they take the kd-merged path (N = 2 * Ntile).
static int pick_ntile(int Nout, int K) {
  if (Nout == 256 && K >= 512) return 256;     // long K loops hide the epilogue; the wider MMA wins (768 -> 256: +5 %)
  if (Nout <= g_ntile_max) return Nout;
  for (int t = g_ntile_max; t >= 16; t -= 16)
    if (Nout % t == 0) return t;
  return 0;
}
static int pick_row_bytes(int K) { return (K % 64 == 0) ? 128 : (K % 32 == 0) ? 64 : (K % 16 == 0) ? 32 : 0; }

static bool plan_tc(const mednet_conv3d_params* q, TcConv* out) {
  TcConv p;
  // the tile grid is the SMALLER of the two volumes for the transposed conv (its 8 parity classes tile the other one)
  // "tf": 8 OUTPUT parity classes on the 2x grid (transposed conv fprop, upsample-conv fprop); "tb": 8 INPUT parity classes
  // read through a stride-2 map (their data gradients); "up": the nearest-upsample conv (MEDNET_GATHER_UPCONV_*)
  const bool up = q->gather == MEDNET_GATHER_UPCONV_F || q->gather == MEDNET_GATHER_UPCONV_B;
  const bool tf = q->gather == MEDNET_GATHER_CONVT_F || q->gather == MEDNET_GATHER_UPCONV_F;
  const bool tb = q->gather == MEDNET_GATHER_CONVT_B || q->gather == MEDNET_GATHER_UPCONV_B;
  p.N = q->N; p.K = q->K; p.Nout = q->Nout;
  p.D = tf ? q->Di : q->Do; p.H = tf ? q->Hi : q->Ho; p.W = tf ? q->Wi : q->Wo;
  p.OD = q->Do; p.OH = q->Ho; p.OW = q->Wo;
  p.ncls_in = tb ? 8 : 1; p.ncls_out = tf ? 8 : 1;
  p.in_scale = tb ? 2 : 1; p.out_scale = tf ? 2 : 1;
  if (tf && (q->Do != 2 * q->Di || q->Ho != 2 * q->Hi || q->Wo != 2 * q->Wi)) return false;
  if (tb && (q->Di != 2 * q->Do || q->Hi != 2 * q->Ho || q->Wi != 2 * q->Wo)) return false;
  for (int cls = 0; cls < 8; ++cls) {
    // window tap index = offset + 1 per axis.  plain conv: offsets -1..1.  transposed fprop, output parity 1: offsets
    // {0, +1}; transposed dgrad, input parity 1: offsets {-1, 0}; parity 0: offset 0 only (components.py:259-264:
    // k3 s2 p1 op1 -> o = 2i - 1 + k)
    uint32_t m = 0;
    for (int tap = 0; tap < 27; ++tap) {
      const int kk[3] = {tap / 9, (tap / 3) % 3, tap % 3}, par[3] = {cls >> 2, (cls >> 1) & 1, cls & 1};
      bool on = true;
      for (int a = 0; a < 3; ++a) {
        if (up && tf) on = on && (par[a] ? kk[a] >= 1 : kk[a] <= 1);        // coarse offsets {0,+1} / {-1,0}
        else if (up) on = on && (par[a] ? kk[a] <= 1 : kk[a] >= 1);         // gradient: mirrored
        else if (tf) on = on && (par[a] ? kk[a] >= 1 : kk[a] == 1);
        else if (tb) on = on && (par[a] ? kk[a] <= 1 : kk[a] == 1);
      }
      if (on) m |= 1u << tap;
    }
    p.tapmask[cls] = m;
  }
  p.RB = pick_row_bytes(q->K);
  p.Ntile = pick_ntile(q->Nout, q->K);
  if (up && g_class_merge && p.Ntile > 128 && q->Nout % 128 == 0) p.Ntile = 128;     // class-merged MMAs: N = 2 * Ntile <= 256
  if (p.RB == 0 || p.Ntile == 0 || (q->Nout % 16) != 0) return false;
  // more planes per brick = fewer halo re-reads and weight-tile loads per voxel; TMEM holds TD * Ntile columns per stage
  p.TD = (p.Ntile <= 64 && p.D >= 4) ? 4 : (p.D >= 2 ? 2 : 1);
  p.cmerge = (up && g_class_merge && p.D >= 2 && 2 * p.Ntile <= 256 && ((p.Ntile * p.RB) % 1024) == 0) ? 1 : 0;
  // class-merged 128-channel tiles: four planes per brick with ONE accumulator stage (4 x 128 = all 512 TMEM columns) --
  // with TD = 2 a staged weight pair (32 KB) feeds 12 MMAs and the launch is bound by L2 (48 B/clk per SM); the epilogue it
  // no longer overlaps is ~5 % of a tile that runs 8 classes (dgrad) or >= 4 chunks (fprop)
  if (p.cmerge && p.Ntile == 128 && p.D >= 4) p.TD = 4;
  p.nplanes = p.cmerge ? p.TD + 1 : p.TD + 2;
  // kd-merged MMAs: plain conv, up to 128-channel output tiles, 32-channel chunks (64-byte rows) so that two halo sets fit
  p.mt = 1;
  p.a_stages = 1;
  if (g_kd_merge && !tf && !tb && p.TD >= 2 && ((p.Ntile * p.RB) % 1024) == 0)
    p.mt = 3 * p.Ntile <= 256 ? 3 : ((2 * p.Ntile <= 256 && p.TD == 2) ? 2 : 1);
  if (p.cmerge) p.mt = 2;
  p.nchunks = q->K / (p.RB / 2);
  const int rc = rb_class(p.RB);
  if (!g_enabled[rc] || g_base_offset_mode[rc] != 0) return false;   // the kernel issues base_offset = 0 descriptors
  p.per_row = g_dense_halo[rc] ? 0 : 1;
  p.pitch = g_dense_halo[rc] ? HALO_W : 16;
  p.bo_mode = g_base_offset_mode[rc];
  p.plane_bytes = (int)align_up((size_t)HALO_H * p.pitch * p.RB, 1024);
  p.b_tile = p.Ntile * p.RB;                     // a multiple of 1024 in kd-merged mode (Ntile % 16 == 0, 64-byte rows)
  p.b_bytes = p.cmerge ? 2 * p.b_tile : (p.mt > 1 ? 3 * p.b_tile : (int)align_up((size_t)p.b_tile, 1024));
  const int budget = 227 * 1024 - 1024 - 1024;   // alignment slack + barrier block
  if (p.mt > 1 && 2 * p.nplanes * p.plane_bytes + 3 * p.b_bytes <= budget) p.a_stages = 2;
  int bs = (budget - p.a_stages * p.nplanes * p.plane_bytes) / p.b_bytes;
  if (bs > MAX_BSTAGES) bs = MAX_BSTAGES;
  if (p.mt > 1 && bs > 4) bs = 4;
  if (bs < 2) return false;
  p.BS = bs;
  p.acc_stages = (2 * p.TD * p.Ntile <= 512) ? 2 : 1;
  p.dual = (p.mt == 1 && g_dual_issue && p.Ntile <= 96 && p.TD >= 2 && (p.TD % 2) == 0) ? 1 : 0;
  int cols = p.acc_stages * p.TD * p.Ntile, pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  if (pow2 > 512) return false;
  p.tmem_cols = pow2;
  p.tiles_w = ceil_div(p.W, TILE_W);
  p.tiles_h = ceil_div(p.H, TILE_H);
  p.tiles_d = ceil_div(p.D, p.TD);
  p.tiles_n = q->Nout / p.Ntile;
  p.num_tiles = (int64_t)p.N * p.tiles_d * p.tiles_h * p.tiles_w * p.tiles_n * p.ncls_out;
  if ((tf || tb) && p.per_row) return false;          // strided halo boxes are implemented for the dense halo only
  p.act = q->act; p.act_param = q->act_param;
  p.bias = q->bias; p.addend = (const bf16*)q->addend; p.y = (bf16*)q->y;
  p.f32_io = q->y_f32 ? 1 : (q->addend_f32 && q->addend ? 2 : 0);
  if (q->y_f32 && (q->act != MEDNET_ACT_NONE || q->addend != nullptr)) return false;
  *out = p;
  return true;
}

bool tc_fprop_supported(const mednet_conv3d_params* q) {
  if (q->dtype != MEDNET_BF16) return false;
  if (q->gather < MEDNET_GATHER_CONV3 || q->gather > MEDNET_GATHER_UPCONV_B) return false;
  if (!mednet_device_has_tcgen05()) return false;
  if (((uintptr_t)q->x | (uintptr_t)q->w | (uintptr_t)q->y | (uintptr_t)q->addend | (uintptr_t)q->bias) & 15) return false;
  TcConv p;
  return plan_tc(q, &p);
}

static int g_conv_profile = 0;                // mednet_tcgen05_set_option("conv_profile", 0|1): see conv3_tc_kernel<PROF>

int tc_fprop(const mednet_conv3d_params* q, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  TcConv p;
  if (!plan_tc(q, &p)) return MEDNET_EUNSUPPORTED;
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return MEDNET_ENODRIVER;
  const int KC = p.RB / 2;
  CUtensorMap map_x, map_w;
  {
    // A is read from the INPUT tensor (q->Di..); transposed dgrad reads every second voxel (traversal stride 2: the box
    // spans 2 * halo elements and lands as halo elements in shared memory), the parity is in the start coordinate
    const cuuint64_t XD = (cuuint64_t)q->Di, XH = (cuuint64_t)q->Hi, XW = (cuuint64_t)q->Wi;
    const cuuint32_t es = (cuuint32_t)p.in_scale;
    cuuint64_t dims[5] = {(cuuint64_t)p.K, XW, XH, XD, (cuuint64_t)p.N};
    cuuint64_t strides[4] = {(cuuint64_t)p.K * 2, XW * p.K * 2, XH * XW * p.K * 2, XD * XH * XW * p.K * 2};
    cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)HALO_W * es, (cuuint32_t)(p.per_row ? 1 : HALO_H * es), 1, 1};
    cuuint32_t estr[5] = {1, es, p.per_row ? 1u : es, 1, 1};
    if (enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(q->x), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(p.RB), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MEDNET_EUNSUPPORTED;
  }
  {
    const int wsets = p.ncls_in > 1 ? p.ncls_in : p.ncls_out;       // one [27][Nout][K] weight set per parity class
    cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)27 * p.Nout * wsets};
    cuuint64_t strides[1] = {(cuuint64_t)p.K * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)p.Ntile};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(q->w), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(p.RB), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MEDNET_EUNSUPPORTED;
  }
  const size_t smem = 1024 + (size_t)p.a_stages * p.nplanes * p.plane_bytes + (size_t)p.BS * p.b_bytes + 1024;
  static std::mutex mu;
  static size_t configured = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (smem > configured) {
      cudaError_t e = cudaFuncSetAttribute(conv3_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(conv3_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      configured = smem;
    }
  }
  int64_t grid = p.num_tiles < sm_count_cached() ? p.num_tiles : sm_count_cached();
  if (g_conv_profile && workspace != nullptr && workspace_bytes >= (size_t)grid * 8 * sizeof(long long))
    conv3_tc_kernel<true><<<(unsigned)grid, NUM_THREADS, smem, st>>>(map_x, map_w, p, (long long*)workspace);
  else
    conv3_tc_kernel<false><<<(unsigned)grid, NUM_THREADS, smem, st>>>(map_x, map_w, p, nullptr);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

// ------------------------------------------------------------------------------------------------
// descriptor probe: D = A_window * I with A written by TMA (swizzle = row bytes), B = identity written
// by hand in the canonical K-major swizzled layout.  K = row_bytes / 2 elements, N = K.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, int rb, int rows, int row_shift, int sbo_bytes, int bo_mode,
             float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;                               // rows x rb bytes (<= 32 KB)
  uint8_t* b_s = smem + 32768;                       // K x rb bytes identity
  uint64_t* bar = (uint64_t*)(b_s + 8192);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = rb / 2;
  const int chunks = rb / 16;                        // 16-byte chunks per row: 8 / 4 / 2
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::fence_barrier_init();
  }
  // identity B[n][k]: row n at n*rb, 16-byte chunk index XORed with the row's swizzle phase
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const int row_off = n * rb;
    const int phase = (row_off >> 7) & (chunks - 1);
    const int off = row_off + ((((k >> 3) ^ phase) & (chunks - 1)) << 4) + (k & 7) * 2;
    *reinterpret_cast<bf16*>(b_s + off) = __float2bfloat16_rn(n == k ? 1.f : 0.f);
  }
  tc::fence_proxy_async();
  if (warp == 0) {
    tc::tmem_alloc(slot, 64);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    tc::mbar_arrive_expect_tx(&bar[0], (uint32_t)(rows * rb));
    tc::tma_load_2d(a_s, &map_a, &bar[0], 0, 0);
    tc::mbar_wait(&bar[0], 0);
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_bf16(128, K, 0, 0);
    const uint32_t layout = rb == 128 ? tc::SWZ_128B : (rb == 64 ? tc::SWZ_64B : tc::SWZ_32B);
    const uint32_t a_addr = tc::smem_u32(a_s) + (uint32_t)(row_shift * rb);
    const uint32_t bo = bo_mode == 0 ? 0u : (bo_mode == 1 ? ((a_addr >> 7) & 7u) : ((a_addr / (uint32_t)rb) & 7u));
    for (int k = 0; k < rb / 32; ++k) {
      const uint64_t da = tc::make_smem_desc(a_addr + k * 32, 16, (uint32_t)sbo_bytes, bo, layout);
      const uint64_t db = tc::make_smem_desc(tc::smem_u32(b_s) + k * 32, 16, (uint32_t)(8 * rb), 0, layout);
      tc::umma_bf16(tmem, da, db, idesc, k != 0);
    }
    tc::umma_commit(&bar[1]);
  }
  __syncwarp();
  tc::mbar_wait(&bar[1], 0);
  tc::tc_fence_after();
  const int m = warp * 32 + lane;
  for (int j = 0; j < K; j += 16) {
    uint32_t r[16];
    tc::tmem_ld_x16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)j, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) out[m * K + j + i] = __uint_as_float(r[i]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

}  // namespace mednet

using namespace mednet;

extern "C" int mednet_tcgen05_configure(int row_bytes, int enabled, int dense_halo, int base_offset_mode) {
  MEDNET_REQUIRE(row_bytes == 128 || row_bytes == 64 || row_bytes == 32, MEDNET_EINVAL);
  MEDNET_REQUIRE(base_offset_mode >= 0 && base_offset_mode <= 2, MEDNET_EINVAL);
  const int rc = rb_class(row_bytes);
  g_enabled[rc] = enabled ? 1 : 0;
  g_dense_halo[rc] = dense_halo ? 1 : 0;
  g_base_offset_mode[rc] = base_offset_mode;
  return MEDNET_OK;
}

extern "C" int mednet_tcgen05_set_option(const char* name, int value) {
  MEDNET_REQUIRE(name != nullptr, MEDNET_EINVAL);
  if (strcmp(name, "dual_issue") == 0) { g_dual_issue = value ? 1 : 0; return MEDNET_OK; }
  if (strcmp(name, "kd_merge") == 0) { g_kd_merge = value ? 1 : 0; return MEDNET_OK; }
  if (strcmp(name, "class_merge") == 0) { g_class_merge = value ? 1 : 0; return MEDNET_OK; }
  if (strcmp(name, "conv_profile") == 0) { g_conv_profile = value ? 1 : 0; return MEDNET_OK; }
  if (strcmp(name, "ntile_max") == 0) { g_ntile_max = value >= 256 ? 256 : 128; return MEDNET_OK; }
  if (strcmp(name, "wgrad_wt_fastest") == 0) { tc_wgrad_set_wt_fastest(value); return MEDNET_OK; }
  if (strcmp(name, "wgrad_pair_planes") == 0) { tc_wgrad_set_pair_planes(value); return MEDNET_OK; }
  if (strcmp(name, "wgrad_d_fastest") == 0) { tc_wgrad_set_d_fastest(value); return MEDNET_OK; }
  if (strcmp(name, "wgrad_profile") == 0) { tc_wgrad_set_profile(value); return MEDNET_OK; }
  if (strcmp(name, "wgrad_dual_issue") == 0) { tc_wgrad_set_dual(value); return MEDNET_OK; }
  if (strcmp(name, "wgrad_class_merge") == 0) { tc_wgrad_set_class_merge(value); return MEDNET_OK; }
  if (strcmp(name, "wgrad_reduce_s_fastest") == 0) { tc_wgrad_set_reduce_s_fastest(value); return MEDNET_OK; }
  if (strcmp(name, "first_layer_mma") == 0) { in1_mma_set_enabled(value); return MEDNET_OK; }
  return MEDNET_EINVAL;
}

extern "C" int mednet_tcgen05_probe(const void* a_bf16, int32_t row_bytes, int32_t rows, int32_t row_shift,
                                    int32_t sbo_bytes, int32_t base_offset_mode, float* out, mednet_stream_t stream) {
  MEDNET_REQUIRE(a_bf16 && out && rows > 0 && rows <= 256 && row_shift >= 0 && sbo_bytes > 0 && (sbo_bytes % 16) == 0,
                 MEDNET_EINVAL);
  MEDNET_REQUIRE(row_bytes == 128 || row_bytes == 64 || row_bytes == 32, MEDNET_EINVAL);
  MEDNET_REQUIRE(row_shift * row_bytes + 15 * sbo_bytes + 8 * row_bytes <= rows * row_bytes, MEDNET_EINVAL);
  if (!mednet_device_has_tcgen05()) return MEDNET_EUNSUPPORTED;
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return MEDNET_ENODRIVER;
  CUtensorMap map_a;
  cuuint64_t dims[2] = {(cuuint64_t)(row_bytes / 2), (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(row_bytes / 2), (cuuint32_t)rows};
  cuuint32_t estr[2] = {1, 1};
  if (enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a_bf16), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_NONE,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return MEDNET_EUNSUPPORTED;
  const size_t smem = 1024 + 32768 + 8192 + 64;
  cudaError_t e = cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  probe_kernel<<<1, 128, smem, stream>>>(map_a, row_bytes, rows, row_shift, sbo_bytes, base_offset_mode, out);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}
