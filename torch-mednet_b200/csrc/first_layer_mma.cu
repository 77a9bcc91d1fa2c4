// First layer of the reference networks (in_channels = 1, midasmednet/segmentation.py:30-31) on the warp-level tensor
// cores: forward and weight gradient of the 1 -> Cout 3x3x3 convolution (components.py:8-9).
//
// With one input channel the layer is HBM-bound in principle (read 2 B, write 2*Cout B per voxel) but its 27 * Cout FMAs per
// voxel made the CUDA-core stencils FMA/LDS-bound (0.09-0.15 of the HBM rate).  Here the 27 taps are the K dimension of an
// m16n8k16 `mma.sync` (K = 27 padded to 32 = two k-steps): the im2col fragment is built in REGISTERS from a shared-memory
// halo of the image, no im2col tile is written anywhere.  tcgen05 is the wrong tool at this size: an M = 128 MMA with
// K = 32, N = 32 is issue- and latency-bound (measured in round 1, profiles/r01j_first_layer_pad.txt).
//   fprop : M = 16 voxels (along w), N = Cout (n-tiles of 8), K = taps;  A = im2col(x) from the halo, B = weights (registers)
//   wgrad : M = Cout (m-tiles of 16), N = taps (4 n-tiles), K = voxels;  A = dY^T via ldmatrix.trans, B = im2col(x)
#include "common.cuh"
#include "conv_impl.h"

namespace mednet {

namespace {

constexpr int TH = 8, TW = 64;              // block tile: 8 rows (one per warp) x 64 voxels of one (n, d) plane
constexpr int XP = TW + 4;                  // halo row pitch in elements (66 used)
constexpr int XS_ELEMS = 3 * (TH + 2) * XP;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack2(uint16_t lo, uint16_t hi) { return (uint32_t)lo | ((uint32_t)hi << 16); }
__device__ __forceinline__ uint16_t bits(bf16 v) { return __bfloat16_as_ushort(v); }

// shared-memory halo of the image: xs[p][r][c] = x[d + p - 1][h0 + r - 1][w0 + c - 1], zero outside the volume
__device__ __forceinline__ void load_halo(uint16_t* xs, const bf16* __restrict__ x, int n, int d, int h0, int w0, int D, int H, int W) {
  for (int i = threadIdx.x; i < 3 * (TH + 2) * (TW + 2); i += blockDim.x) {
    const int c = i % (TW + 2), r = (i / (TW + 2)) % (TH + 2), p = i / ((TW + 2) * (TH + 2));
    const int id = d + p - 1, ih = h0 + r - 1, iw = w0 + c - 1;
    uint16_t v = 0;
    if (id >= 0 && id < D && ih >= 0 && ih < H && iw >= 0 && iw < W) v = bits(x[(((int64_t)n * D + id) * H + ih) * W + iw]);
    xs[(p * (TH + 2) + r) * XP + c] = v;
  }
}
// halo offset of tap k (relative to the output voxel's (row, col) in tile coordinates); taps >= 27 alias tap 26 (their
// weights / results are zero / unused)
__device__ __forceinline__ int tap_off(int k) {
  if (k > 26) k = 26;
  return ((k / 9) * (TH + 2) + (k / 3) % 3) * XP + k % 3;
}

// ------------------------------------------------------------------------------------------------ forward
// w: bf16 [Cout][27] (MEDNET_WPACK_SIMT_FPROP with Cin = 1).  NT = Cout / 8.
template <int NT>
__global__ void __launch_bounds__(256, 2) in1_mma_fprop_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                            const float* __restrict__ bias, bf16* __restrict__ y, int D, int H,
                                                            int W, int tiles_w, int act, float act_param) {
  __shared__ uint16_t xs[XS_ELEMS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int n = blockIdx.y / D, d = blockIdx.y - n * D;
  const int h0 = (blockIdx.x / tiles_w) * TH, w0 = (blockIdx.x % tiles_w) * TW;
  constexpr int Nout = NT * 8;
  load_halo(xs, x, n, d, h0, w0, D, H, W);
  // B fragments (weights): element (k, n) = w[n][k], k = ks * 16 + 2q + {0, 1} (+ 8), n = nt * 8 + g
  uint32_t bfr[NT][2][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int k = ks * 16 + r * 8 + 2 * q;
        const bf16* wr = w + (nt * 8 + g) * 27;
        bfr[nt][ks][r] = pack2(k < 27 ? bits(wr[k]) : (uint16_t)0, k + 1 < 27 ? bits(wr[k + 1]) : (uint16_t)0);
      }
  int koff[2][2][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int e = 0; e < 2; ++e) koff[ks][r][e] = tap_off(ks * 16 + r * 8 + 2 * q + e);
  __syncthreads();
  const int h = h0 + warp;
  if (h >= H) return;
  const uint16_t* row = xs + warp * XP;               // tile row `warp` <-> halo rows warp + kh
#pragma unroll 1
  for (int mt = 0; mt < TW / 16; ++mt) {
    if (w0 + mt * 16 >= W) break;
    float c[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float b0 = bias ? bias[nt * 8 + 2 * q] : 0.f, b1 = bias ? bias[nt * 8 + 2 * q + 1] : 0.f;
      c[nt][0] = b0; c[nt][1] = b1; c[nt][2] = b0; c[nt][3] = b1;
    }
    const uint16_t* p0 = row + mt * 16 + g;           // voxel of fragment row g; row g + 8 is 8 voxels further
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t a[4];
      a[0] = pack2(p0[koff[ks][0][0]], p0[koff[ks][0][1]]);
      a[1] = pack2(p0[8 + koff[ks][0][0]], p0[8 + koff[ks][0][1]]);
      a[2] = pack2(p0[koff[ks][1][0]], p0[koff[ks][1][1]]);
      a[3] = pack2(p0[8 + koff[ks][1][0]], p0[8 + koff[ks][1][1]]);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(c[nt], a, bfr[nt][ks]);
    }
    const int wa = w0 + mt * 16 + g, wb = wa + 8;
    bf16* ya = y + ((((int64_t)n * D + d) * H + h) * W + wa) * Nout + 2 * q;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (wa < W)
        *reinterpret_cast<__nv_bfloat162*>(ya + nt * 8) =
            __floats2bfloat162_rn(act_apply(c[nt][0], act, act_param), act_apply(c[nt][1], act, act_param));
      if (wb < W)
        *reinterpret_cast<__nv_bfloat162*>(ya + (int64_t)8 * Nout + nt * 8) =
            __floats2bfloat162_rn(act_apply(c[nt][2], act, act_param), act_apply(c[nt][3], act, act_param));
    }
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// partial[block][co][27] = sum over the block's tiles of dY[v][co] * x[v + tap].  MT = Ca / 16.
template <int MT>
__global__ void __launch_bounds__(256) in1_mma_wgrad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                            float* __restrict__ partial, int N, int D, int H, int W,
                                                            int tiles_h, int tiles_w) {
  constexpr int Ca = MT * 16;
  constexpr int VP = Ca + 8;                         // voxel pitch of the dY tile in elements (+16 B: conflict-free ldmatrix)
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint16_t* xs = reinterpret_cast<uint16_t*>(smem_raw);                                    // halo of x
  uint16_t* dys = xs + ((XS_ELEMS + 7) & ~7);                                                // [TH][TW][VP] bf16
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  float acc[MT][4][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  int toff[4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) toff[nt] = tap_off(nt * 8 + g);
  const int64_t tiles = (int64_t)N * D * tiles_h * tiles_w;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    int64_t r = t;
    const int w0 = (int)(r % tiles_w) * TW; r /= tiles_w;
    const int h0 = (int)(r % tiles_h) * TH; r /= tiles_h;
    const int d = (int)(r % D);
    const int n = (int)(r / D);
    __syncthreads();                                  // previous tile fully consumed
    load_halo(xs, x, n, d, h0, w0, D, H, W);
    // dY tile: TH x TW voxels x Ca channels in 16-byte pieces, zero outside the volume
    constexpr int PIECES = Ca / 8;
    for (int i = threadIdx.x; i < TH * TW * PIECES; i += blockDim.x) {
      const int pc = i % PIECES, v = (i / PIECES) % TW, hr = i / (PIECES * TW);
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (h0 + hr < H && w0 + v < W)
        val = *reinterpret_cast<const uint4*>(dy + ((((int64_t)n * D + d) * H + h0 + hr) * W + w0 + v) * Ca + pc * 8);
      *reinterpret_cast<uint4*>(dys + ((size_t)hr * TW + v) * VP + pc * 8) = val;
    }
    __syncthreads();
    const uint16_t* xrow = xs + warp * XP;
    const uint16_t* drow = dys + (size_t)warp * TW * VP;
#pragma unroll 1
    for (int ks = 0; ks < TW / 16; ++ks) {
      // B fragments: element (k = voxel 2q + {0,1} (+8), n = tap nt*8 + g) = x[voxel + tap offset]
      uint32_t b[4][2];
      const uint16_t* xv = xrow + ks * 16 + 2 * q;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        b[nt][0] = pack2(xv[toff[nt]], xv[toff[nt] + 1]);
        b[nt][1] = pack2(xv[toff[nt] + 8], xv[toff[nt] + 9]);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        // A = dY^T (row = channel, col = voxel) through ldmatrix.trans of the [voxel][channel] tile
        uint32_t a[4];
        const int j = lane >> 3;                      // matrix this lane addresses: voxel block j >> 1, channel block j & 1
        const uint16_t* src = drow + (size_t)(ks * 16 + (j >> 1) * 8 + (lane & 7)) * VP + mt * 16 + (j & 1) * 8;
        const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(src);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                     : "r"(saddr));
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], a, b[nt]);
      }
    }
  }
  // combine the 8 warps (fixed order) -> partial[block][co][tap]
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem_raw);    // [8 warps][Ca][32]
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float* o = red + ((size_t)warp * Ca + mt * 16 + g) * 32 + nt * 8 + 2 * q;
      o[0] = acc[mt][nt][0];
      o[1] = acc[mt][nt][1];
      o[8 * 32] = acc[mt][nt][2];
      o[8 * 32 + 1] = acc[mt][nt][3];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < Ca * 27; i += blockDim.x) {
    const int co = i / 27, tap = i - co * 27;
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[((size_t)wv * Ca + co) * 32 + tap];
    partial[(int64_t)blockIdx.x * Ca * 27 + i] = s;
  }
}

int g_enabled = 1;           // mednet_tcgen05_set_option("first_layer_mma", 0|1)

}  // namespace

void in1_mma_set_enabled(int v) { g_enabled = v ? 1 : 0; }

bool in1_mma_fprop_ok(const mednet_conv3d_params* p) {
  return g_enabled && p->dtype == MEDNET_BF16 && p->gather == MEDNET_GATHER_CONV3 && p->K == 1 && p->addend == nullptr &&
         (p->Nout == 8 || p->Nout == 16 || p->Nout == 32 || p->Nout == 64) && (int64_t)p->N * p->Do <= 65535;
}

int in1_mma_fprop(const mednet_conv3d_params* p, cudaStream_t st) {
  const int tiles_w = ceil_div(p->Wo, TW), tiles_h = ceil_div(p->Ho, TH);
  dim3 grid((unsigned)(tiles_w * tiles_h), (unsigned)(p->N * p->Do));
#define IN1_F(NTV)                                                                                                      \
  in1_mma_fprop_kernel<NTV><<<grid, 256, 0, st>>>((const bf16*)p->x, (const bf16*)p->w, p->bias, (bf16*)p->y, p->Do, p->Ho, \
                                                  p->Wo, tiles_w, p->act, p->act_param)
  switch (p->Nout) {
    case 8: IN1_F(1); break;
    case 16: IN1_F(2); break;
    case 32: IN1_F(4); break;
    default: IN1_F(8); break;
  }
#undef IN1_F
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

bool in1_mma_wgrad_ok(const mednet_wgrad_params* p) {
  return g_enabled && p->dtype == MEDNET_BF16 && p->gather == MEDNET_GATHER_CONV3 && p->Cb == 1 &&
         (p->Ca == 16 || p->Ca == 32 || p->Ca == 64);
}

// writes partial[blocks][Ca][27]; `blocks` <= the count the caller sized its workspace for
int in1_mma_wgrad(const mednet_wgrad_params* p, float* partial, int max_blocks, int* blocks_out, cudaStream_t st) {
  const int tiles_w = ceil_div(p->Wa, TW), tiles_h = ceil_div(p->Ha, TH);
  const int64_t tiles = (int64_t)p->N * p->Da * tiles_h * tiles_w;
  int64_t blocks = (int64_t)sm_count_cached() * 3;
  if (blocks > tiles) blocks = tiles;
  if (blocks > max_blocks) blocks = max_blocks;
  *blocks_out = (int)blocks;
  const size_t xs_bytes = (size_t)((XS_ELEMS + 7) & ~7) * 2;
  auto smem_for = [&](int Ca) {
    const size_t tile = xs_bytes + (size_t)TH * TW * (Ca + 8) * 2, red = (size_t)8 * Ca * 32 * sizeof(float);
    return tile > red ? tile : red;
  };
  const size_t smem = smem_for(p->Ca);
#define IN1_W(MTV)                                                                                                        \
  do {                                                                                                                    \
    cudaError_t e = cudaFuncSetAttribute(in1_mma_wgrad_kernel<MTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return (int)e;                                                                                  \
    in1_mma_wgrad_kernel<MTV><<<(unsigned)blocks, 256, smem, st>>>((const bf16*)p->a, (const bf16*)p->b, partial, p->N,    \
                                                                   p->Da, p->Ha, p->Wa, tiles_h, tiles_w);               \
  } while (0)
  if (p->Ca == 16) IN1_W(1);
  else if (p->Ca == 32) IN1_W(2);
  else IN1_W(4);
#undef IN1_W
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

}  // namespace mednet
