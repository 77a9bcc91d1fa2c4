// Final 1x1x1 convolution with bias: NDHWC activations -> NCDHW fp32 logits, and its backward.
// ref: midasmednet/unet/model.py:77,102 (UNet3D) and :179,207 (ResidualUNet3D).
// Memory-bound (Cout = classes + heatmaps is tiny): one pass over the widest activation of the net.
#include "common.cuh"

namespace mednet {

constexpr int CO_TILE = 8;

// thread per voxel; x row read with 16-byte vectors; weights staged in shared memory
template <typename T, int V>
__global__ void conv1_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                 float* __restrict__ y, int64_t N, int64_t S, int Cin, int Cout) {
  extern __shared__ float sw[];  // [Cout][Cin]
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int64_t total = N * S;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = r / S, s = r - n * S;
    const T* xr = x + r * Cin;
    for (int c0 = 0; c0 < Cout; c0 += CO_TILE) {
      float acc[CO_TILE];
#pragma unroll
      for (int j = 0; j < CO_TILE; ++j) acc[j] = (c0 + j < Cout) ? bias[c0 + j] : 0.f;
      for (int k = 0; k < Cin; k += V) {
        float xv[V];
        load_vec<T, V>(xr + k, xv);
#pragma unroll
        for (int j = 0; j < CO_TILE; ++j) {
          if (c0 + j < Cout) {
            const float* wr = sw + (c0 + j) * Cin + k;
#pragma unroll
            for (int i = 0; i < V; ++i) acc[j] = fmaf(xv[i], wr[i], acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < CO_TILE; ++j)
        if (c0 + j < Cout) y[(n * Cout + (c0 + j)) * S + s] = acc[j];
    }
  }
}


// One voxel per thread: the thread issues ALL 16-byte loads of its channel row first (CIN / 8 of them in flight), takes the
// CO_T x CIN weights from shared memory (every lane reads the same address: broadcast) and writes CO_T NCDHW values that are
// contiguous across the warp.  No shuffles, no staging tile, no block barrier per 256 voxels; a warp's row loads hit
// each 128-byte line once per instruction and L1 serves the other seven.  bf16, CIN in {32, 64, 128}.
template <int CIN, int CO_T>
__global__ void __launch_bounds__(256) conv1_fwd_row_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y,
                                                            int64_t total, int64_t S, int Cout) {
  __shared__ float ws[CO_T * CIN];
  __shared__ float bs[CO_T];
  for (int co0 = 0; co0 < Cout; co0 += CO_T) {
    const int nco = min(CO_T, Cout - co0);
    __syncthreads();
    for (int i = threadIdx.x; i < CO_T * CIN; i += blockDim.x) ws[i] = (i / CIN) < nco ? w[(co0 + i / CIN) * CIN + i % CIN] : 0.f;
    if (threadIdx.x < CO_T) bs[threadIdx.x] = threadIdx.x < nco ? bias[co0 + threadIdx.x] : 0.f;
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
      asm volatile("" ::: "memory");      // keep the weight reads inside the loop (hoisted, they cost CO_T * CIN registers)
      uint4 raw[CIN / 8];
      const uint4* row = reinterpret_cast<const uint4*>(x + r * CIN);
#pragma unroll
      for (int i = 0; i < CIN / 8; ++i) raw[i] = row[i];
      float acc[CO_T];
#pragma unroll
      for (int j = 0; j < CO_T; ++j) acc[j] = bs[j];
#pragma unroll
      for (int i = 0; i < CIN / 8; ++i) {
        float xv[8];
        cvt_raw<bf16, 8>(raw[i], xv);
#pragma unroll
        for (int j = 0; j < CO_T; ++j) {
          const float4 w0 = *reinterpret_cast<const float4*>(ws + j * CIN + i * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(ws + j * CIN + i * 8 + 4);
          acc[j] = fmaf(xv[0], w0.x, fmaf(xv[1], w0.y, fmaf(xv[2], w0.z, fmaf(xv[3], w0.w, acc[j]))));
          acc[j] = fmaf(xv[4], w1.x, fmaf(xv[5], w1.y, fmaf(xv[6], w1.z, fmaf(xv[7], w1.w, acc[j]))));
        }
      }
      const int64_t n = r / S, sidx = r - n * S;
#pragma unroll
      for (int j = 0; j < CO_T; ++j)
        if (j < nco) y[(n * Cout + co0 + j) * S + sidx] = acc[j];
    }
  }
}

// Coalesced forward for power-of-two Cin / V <= 32 (16-byte vectors): `ncol` adjacent lanes share one voxel row
// (one fully used 16 B x ncol line per row), partial dot products are combined with warp shuffles, the block's
// 256 voxels x CO_T results are staged in shared memory and written as contiguous NCDHW runs.
template <typename T, int V, int CO_T>
__global__ void __launch_bounds__(256) conv1_fwd_vec_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y,
                                                            int64_t total, int64_t S, int Cin, int Cout) {
  __shared__ float tile[CO_T][256];
  const int ncol = Cin / V, vpp = 256 / ncol;                  // voxels per pass
  const int col = threadIdx.x % ncol, vl = threadIdx.x / ncol;
  for (int co0 = 0; co0 < Cout; co0 += CO_T) {
    const int nco = min(CO_T, Cout - co0);
    float wr[CO_T][V];
#pragma unroll
    for (int j = 0; j < CO_T; ++j)
#pragma unroll
      for (int i = 0; i < V; ++i) wr[j][i] = j < nco ? w[(co0 + j) * Cin + col * V + i] : 0.f;
    for (int64_t base = (int64_t)blockIdx.x * 256; base < total; base += (int64_t)gridDim.x * 256) {
      for (int ps0 = 0; ps0 < ncol; ps0 += 8) {
        // up to 8 passes' rows are loaded before any arithmetic (memory-level parallelism)
        typename RawVec<sizeof(T) * V>::type raw[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t r = base + (ps0 + u) * vpp + vl;
          if (ps0 + u < ncol && r < total) raw[u] = load_raw<T, V>(x + r * Cin + col * V);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (ps0 + u < ncol) {
            const int vb = (ps0 + u) * vpp + vl;                 // voxel slot inside the block's 256
            const int64_t r = base + vb;
            float acc[CO_T];
#pragma unroll
            for (int j = 0; j < CO_T; ++j) acc[j] = 0.f;
            if (r < total) {
              float xv[V];
              cvt_raw<T, V>(raw[u], xv);
#pragma unroll
              for (int j = 0; j < CO_T; ++j)
#pragma unroll
                for (int i = 0; i < V; ++i) acc[j] = fmaf(xv[i], wr[j][i], acc[j]);
            }
            for (int off = 1; off < ncol; off <<= 1)
#pragma unroll
              for (int j = 0; j < CO_T; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
            if (col == 0) {
#pragma unroll
              for (int j = 0; j < CO_T; ++j) tile[j][vb] = acc[j];
            }
          }
        }
      }
      __syncthreads();
      const int64_t r = base + threadIdx.x;
      if (r < total) {
        const int64_t n = r / S, sidx = r - n * S;
#pragma unroll
        for (int j = 0; j < CO_T; ++j)
          if (j < nco) y[(n * Cout + co0 + j) * S + sidx] = tile[j][threadIdx.x] + bias[co0 + j];
      }
      __syncthreads();
    }
  }
}

// dx[r][ci] = sum_co dy[n][co][s] * w[co][ci]  (times in_act'(x) when the producer's activation is deferred)
template <typename T, int V>
__global__ void conv1_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx,
                                int64_t N, int64_t S, int Cin, int Cout, const T* __restrict__ x, int in_act,
                                float in_act_param) {
  extern __shared__ float sw[];  // [Cout][Cin]
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int64_t total = N * S;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = r / S, s = r - n * S;
    T* out = dx + r * Cin;
    for (int k = 0; k < Cin; k += V) {
      float acc[V];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = 0.f;
      for (int co = 0; co < Cout; ++co) {
        const float g = dy[(n * Cout + co) * S + s];
        const float* wr = sw + co * Cin + k;
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = fmaf(g, wr[i], acc[i]);
      }
      if (in_act != MEDNET_ACT_NONE) {
        float xv[V];
        load_vec<T, V>(x + r * Cin + k, xv);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] *= act_grad_from_out(xv[i], in_act, in_act_param);
      }
      store_vec<T, V>(out + k, acc);
    }
  }
}

// partial[b][co][ci] = sum over the block's rows of dy[.,co] * x[.,ci]
// block = (Cin_t, R): thread (tx,ty) owns input channel tx and rows ty, ty+R, ...
constexpr int DW_MAX_CO = 16;
constexpr int DW_STAGES = 6;   // rows in flight per thread in the vectorised dw/dx kernel
template <typename T>
__global__ void conv1_dw_partial_kernel(const T* __restrict__ x, const float* __restrict__ dy,
                                        float* __restrict__ partial, int64_t N, int64_t S, int Cin, int Cout,
                                        int co0, int64_t rows_per_block) {
  extern __shared__ float sm[];
  const int ci = blockIdx.y * blockDim.x + threadIdx.x;
  const int64_t total = N * S;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > total) r1 = total;
  const int nco = min(DW_MAX_CO, Cout - co0);
  float acc[DW_MAX_CO];
#pragma unroll
  for (int j = 0; j < DW_MAX_CO; ++j) acc[j] = 0.f;
  const bool active = ci < Cin;
  for (int64_t r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    const int64_t n = r / S, s = r - n * S;
    const float xv = active ? to_f32<T>(x[r * Cin + ci]) : 0.f;
#pragma unroll
    for (int j = 0; j < DW_MAX_CO; ++j) {
      if (j < nco) {
        const float g = dy[(n * Cout + co0 + j) * S + s];
        acc[j] = fmaf(g, xv, acc[j]);
      }
    }
  }
  // reduce across threadIdx.y
  const int bx = blockDim.x, R = blockDim.y, tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int j = 0; j < DW_MAX_CO; ++j) sm[(ty * bx + tx) * DW_MAX_CO + j] = acc[j];
  __syncthreads();
  for (int j = ty; j < nco; j += R) {
    float a = 0.f;
    for (int t = 0; t < R; ++t) a += sm[(t * bx + tx) * DW_MAX_CO + j];
    if (active) partial[((int64_t)blockIdx.x * Cout + co0 + j) * Cin + ci] = a;
  }
}

// Vectorised weight gradient for power-of-two Cin / V <= 32: thread (tx, ty) owns V input channels and walks rows
// ty, ty+R, ... of one sample's slab (no 64-bit division in the loop: blockIdx.y = sample); CO_T output channels per
// pass.  Row-lanes are combined with warp shuffles, warps through shared memory.
// partial[(n*nblk + b)][co][ci]
template <typename T, int V, int CO_T>
__global__ void __launch_bounds__(256) conv1_dw_vec_kernel(const T* __restrict__ x, const float* __restrict__ dy,
                                                           float* __restrict__ partial, int64_t S, int Cin, int Cout, int co0,
                                                           int64_t rows_per_block, const float* __restrict__ w,
                                                           T* __restrict__ dx, int in_act, float in_act_param) {
  extern __shared__ float sm[];                                  // [8 warps][ncol][CO_T * V]
  const int ncol = Cin / V, R = 256 / ncol;
  const int tx = threadIdx.x % ncol, ty = threadIdx.x / ncol;
  const int n = blockIdx.y;
  const int64_t s0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t s1 = s0 + rows_per_block;
  if (s1 > S) s1 = S;
  const int nco = min(CO_T, Cout - co0);
  float acc[CO_T][V];
#pragma unroll
  for (int j = 0; j < CO_T; ++j)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[j][i] = 0.f;
  const T* xb = x + (int64_t)n * S * Cin + tx * V;
  const float* dyb = dy + ((int64_t)n * Cout + co0) * S;
  // Private LDGSTS pipeline: the x vector (16 B) and the CO_T dy scalars of this thread's next DW_STAGES - 1 rows are
  // in flight in its own shared-memory slots (the accumulators keep the register file full, so loads cannot be
  // batched in registers).
  uint8_t* pipe = reinterpret_cast<uint8_t*>(sm) + (size_t)8 * ncol * CO_T * V * sizeof(float);
  uint8_t* xslot = pipe + (size_t)threadIdx.x * 16;                                    // [stage][256] x 16 B
  float* gslot = reinterpret_cast<float*>(pipe + (size_t)DW_STAGES * 256 * 16) + threadIdx.x * CO_T;   // [stage][256][CO_T]
  const int64_t nrows = s1 > s0 + ty ? (s1 - s0 - ty + R - 1) / R : 0;
  auto issue = [&](int64_t k) {
    if (k < nrows) {
      const int64_t srow = s0 + ty + k * R;
      const int st = (int)(k % DW_STAGES);
      cp_async16(xslot + (size_t)st * 256 * 16, xb + srow * Cin);
#pragma unroll
      for (int j = 0; j < CO_T; ++j)
        if (j < nco) cp_async4(gslot + (size_t)st * 256 * CO_T + j, dyb + (int64_t)j * S + srow);
    }
    cp_async_commit();
  };
  for (int k = 0; k < DW_STAGES - 1; ++k) issue(k);
  float wr[CO_T][V];
  T* dxb = nullptr;
  if (dx != nullptr) {
    // fused input gradient (single pass: Cout <= CO_T): x and dy are read ONCE for dx, dw and db;
    // dx = (sum_co dy[co] * w[co][ci]) * in_act'(x)
#pragma unroll
    for (int j = 0; j < CO_T; ++j)
#pragma unroll
      for (int i = 0; i < V; ++i) wr[j][i] = j < nco ? w[(co0 + j) * Cin + tx * V + i] : 0.f;
    dxb = dx + (int64_t)n * S * Cin + tx * V;
  }
  for (int64_t k = 0; k < nrows; ++k) {
    issue(k + DW_STAGES - 1);
    cp_async_wait<DW_STAGES - 1>();
    const int st = (int)(k % DW_STAGES);
    const int64_t srow = s0 + ty + k * R;
    float xv[V], g[CO_T];
    load_vec<T, V>(reinterpret_cast<const T*>(xslot + (size_t)st * 256 * 16), xv);
#pragma unroll
    for (int j = 0; j < CO_T; ++j) g[j] = j < nco ? gslot[(size_t)st * 256 * CO_T + j] : 0.f;
#pragma unroll
    for (int j = 0; j < CO_T; ++j)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[j][i] = fmaf(g[j], xv[i], acc[j][i]);
    if (dx != nullptr) {
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = 0.f;
#pragma unroll
      for (int j = 0; j < CO_T; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = fmaf(g[j], wr[j][i], o[i]);
      if (in_act != MEDNET_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] *= act_grad_from_out(xv[i], in_act, in_act_param);
      }
      store_vec<T, V>(dxb + srow * Cin, o);
    }
  }
  cp_async_wait<0>();
  // rows that share a warp (lane = (ty % rpw) * ncol + tx when ncol < 32)
  if (ncol < 32) {
    for (int off = ncol; off < 32; off <<= 1)
#pragma unroll
      for (int j = 0; j < CO_T; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[j][i] += __shfl_xor_sync(0xffffffffu, acc[j][i], off);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cols_in_warp = ncol < 32 ? ncol : 32;                // distinct tx values a warp holds
  const int colw = ncol < 32 ? tx : lane;                        // this thread's column slot inside its warp
  const int warps_per_colset = ncol < 32 ? 8 : 8 / (ncol / 32);  // warps holding the same columns
  if (ncol >= 32 || lane < ncol) {
#pragma unroll
    for (int j = 0; j < CO_T; ++j)
#pragma unroll
      for (int i = 0; i < V; ++i) sm[((size_t)warp * cols_in_warp + colw) * (CO_T * V) + j * V + i] = acc[j][i];
  }
  __syncthreads();
  // final: thread t < ncol * CO_T * V sums over the warps that hold its column
  float* out = partial + ((int64_t)n * gridDim.x + blockIdx.x) * (int64_t)Cout * Cin;
  for (int t = threadIdx.x; t < ncol * CO_T * V; t += 256) {
    const int col = t / (CO_T * V), e = t % (CO_T * V), j = e / V, i = e % V;
    if (j >= nco) continue;
    float a = 0.f;
    if (ncol < 32) {
      for (int wp = 0; wp < 8; ++wp) a += sm[((size_t)wp * cols_in_warp + col) * (CO_T * V) + e];
    } else {
      const int wset = col / 32, cw = col % 32, nset = ncol / 32;
      for (int k = 0; k < warps_per_colset; ++k) a += sm[((size_t)(k * nset + wset) * 32 + cw) * (CO_T * V) + e];
    }
    out[(int64_t)(co0 + j) * Cin + col * V + i] = a;
  }
}

// db partial, per-sample blocks (blockIdx.y = sample): block sums dy[n, co, its rows]
__global__ void conv1_db_partial_kernel(const float* __restrict__ dy, float* __restrict__ partial_b, int64_t S, int Cout,
                                        int64_t rows_per_block) {
  __shared__ float scratch[32];
  const int n = blockIdx.y;
  const int64_t s0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t s1 = s0 + rows_per_block;
  if (s1 > S) s1 = S;
  for (int co = 0; co < Cout; ++co) {
    const float* src = dy + ((int64_t)n * Cout + co) * S;
    float a = 0.f;
    for (int64_t s = s0 + threadIdx.x; s < s1; s += blockDim.x) a += src[s];
    a = block_sum(a, scratch);
    if (threadIdx.x == 0) partial_b[((int64_t)n * gridDim.x + blockIdx.x) * Cout + co] = a;
  }
}

// db partial over flat row ranges (scalar path)
__global__ void conv1_db_partial_rows_kernel(const float* __restrict__ dy, float* __restrict__ partial_b, int64_t N,
                                             int64_t S, int Cout, int64_t rows_per_block) {
  __shared__ float scratch[32];
  const int64_t total = N * S;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > total) r1 = total;
  for (int co = 0; co < Cout; ++co) {
    float a = 0.f;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
      const int64_t n = r / S, s = r - n * S;
      a += dy[(n * Cout + co) * S + s];
    }
    a = block_sum(a, scratch);
    if (threadIdx.x == 0) partial_b[(int64_t)blockIdx.x * Cout + co] = a;
  }
}

__global__ void conv1_dw_final_kernel(const float* __restrict__ partial, const float* __restrict__ partial_b,
                                      float* __restrict__ dw, float* __restrict__ db, int Cin, int Cout, int nblocks,
                                      int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nw = Cout * Cin;
  if (i < nw) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += (double)partial[(int64_t)b * nw + i];
    dw[i] = accumulate ? dw[i] + (float)s : (float)s;
  } else if (i < nw + Cout) {
    const int co = i - nw;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += (double)partial_b[(int64_t)b * Cout + co];
    db[co] = accumulate ? db[co] + (float)s : (float)s;
  }
}

struct Conv1Plan {
  int nblocks, cin_t, cin_tiles, R;     // scalar kernel: nblocks over all rows
  int64_t rows_per_block;
  int vec, V, ncol, co_t, nblk_per_sample;   // vectorised kernel: nblk_per_sample blocks per sample
  int64_t rows_per_block_vec;
};
static Conv1Plan conv1_plan(int64_t N, int64_t S, int Cin, int Cout, int elem_bytes) {
  Conv1Plan pl;
  const int64_t rows = N * S;
  pl.cin_t = Cin < 256 ? Cin : 256;
  pl.cin_tiles = ceil_div(Cin, pl.cin_t);
  pl.R = 256 / pl.cin_t;
  if (pl.R < 1) pl.R = 1;
  int64_t nb = (int64_t)sm_count_cached() * 4 / pl.cin_tiles;
  const int64_t maxb = rows / ((int64_t)pl.R * 8);
  if (nb > maxb) nb = maxb;
  if (nb < 1) nb = 1;
  pl.rows_per_block = ceil_div64(rows, nb);
  pl.nblocks = (int)ceil_div64(rows, pl.rows_per_block);
  // vectorised variant
  pl.V = pick_vec(Cin, elem_bytes);
  pl.ncol = Cin / pl.V;
  pl.co_t = Cout <= 4 ? 4 : (Cout <= 8 ? 8 : 16);
  pl.vec = (pl.ncol <= 16 && (pl.ncol & (pl.ncol - 1)) == 0 && pl.V * elem_bytes == 16 && N <= 65535 &&
            (size_t)8 * pl.ncol * pl.co_t * pl.V * sizeof(float) + (size_t)DW_STAGES * 256 * (16 + 4 * pl.co_t) <= 100 * 1024)
               ? 1 : 0;
  int64_t per = ((int64_t)sm_count_cached() * 4 + N - 1) / N;
  const int64_t maxper = S / ((256 / pl.ncol) * 8);
  if (per > maxper) per = maxper;
  if (per < 1) per = 1;
  pl.rows_per_block_vec = ceil_div64(S, per);
  pl.nblk_per_sample = (int)ceil_div64(S, pl.rows_per_block_vec);
  if (pl.vec) pl.nblocks = (int)(N * pl.nblk_per_sample);
  return pl;
}

}  // namespace mednet

using namespace mednet;

extern "C" int mednet_conv1x1_fwd(const mednet_conv1_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->x && p->w && p->bias && p->y && p->N > 0 && p->S > 0 && p->Cin > 0 && p->Cout > 0,
                 MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  const size_t smem = (size_t)p->Cin * p->Cout * sizeof(float);
  MEDNET_REQUIRE(smem <= 48 * 1024, MEDNET_EUNSUPPORTED);
  const int V = pick_vec(p->Cin, dtype_bytes(p->dtype));
  const int ncol = p->Cin / V;
  if (p->dtype == MEDNET_BF16 && (p->Cin == 32 || p->Cin == 64 || p->Cin == 128) && (((uintptr_t)p->x) & 15) == 0) {
    const int64_t total = p->N * p->S;
    const int grid = grid_for(total, 256, 8);
#define MEDNET_C1R(CI)                                                                                                   \
    do {                                                                                                                 \
      if (p->Cout <= 4) conv1_fwd_row_kernel<CI, 4><<<grid, 256, 0, stream>>>((const bf16*)p->x, p->w, p->bias, p->y, total, p->S, p->Cout); \
      else conv1_fwd_row_kernel<CI, 8><<<grid, 256, 0, stream>>>((const bf16*)p->x, p->w, p->bias, p->y, total, p->S, p->Cout);              \
    } while (0)
    if (p->Cin == 32) MEDNET_C1R(32); else if (p->Cin == 64) MEDNET_C1R(64); else MEDNET_C1R(128);
#undef MEDNET_C1R
    MEDNET_LAUNCH_CHECK();
    return MEDNET_OK;
  }
  if (V * dtype_bytes(p->dtype) == 16 && ncol <= 32 && (ncol & (ncol - 1)) == 0) {
    const int64_t total = p->N * p->S;
    const int grid = grid_for(ceil_div64(total, 256) * 256, 256, 8);
#define MEDNET_C1F(TT, VV, CT) \
    conv1_fwd_vec_kernel<TT, VV, CT><<<grid, 256, 0, stream>>>((const TT*)p->x, p->w, p->bias, p->y, total, p->S, p->Cin, p->Cout)
    if (p->dtype == MEDNET_F32) { if (p->Cout <= 4) MEDNET_C1F(float, 4, 4); else MEDNET_C1F(float, 4, 8); }
    else { if (p->Cout <= 4) MEDNET_C1F(bf16, 8, 4); else MEDNET_C1F(bf16, 8, 8); }
#undef MEDNET_C1F
    MEDNET_LAUNCH_CHECK();
    return MEDNET_OK;
  }
  MEDNET_DISPATCH_TV(p->dtype, V, {
    conv1_fwd_kernel<T, VV><<<grid_for(p->N * p->S, 128, 16), 128, smem, stream>>>((const T*)p->x, p->w, p->bias, p->y,
                                                                                  p->N, p->S, p->Cin, p->Cout);
  });
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" size_t mednet_conv1x1_bwd_workspace_bytes(const mednet_conv1_bwd_params* p) {
  if (!p || p->Cin <= 0 || p->Cout <= 0) return 0;
  Conv1Plan pl = conv1_plan(p->N, p->S, p->Cin, p->Cout, dtype_bytes(p->dtype));
  return align_up((size_t)pl.nblocks * p->Cout * p->Cin * sizeof(float), 256) +
         align_up((size_t)pl.nblocks * p->Cout * sizeof(float), 256);
}

extern "C" int mednet_conv1x1_bwd(const mednet_conv1_bwd_params* p, void* workspace, size_t workspace_bytes,
                                  mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->x && p->w && p->dy && p->dw && p->db && p->N > 0 && p->S > 0 && p->Cin > 0 && p->Cout > 0,
                 MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_conv1x1_bwd_workspace_bytes(p), MEDNET_EWORKSPACE);
  const size_t smem = (size_t)p->Cin * p->Cout * sizeof(float);
  MEDNET_REQUIRE(smem <= 48 * 1024, MEDNET_EUNSUPPORTED);
  const int V = pick_vec(p->Cin, dtype_bytes(p->dtype));
  Conv1Plan pl = conv1_plan(p->N, p->S, p->Cin, p->Cout, dtype_bytes(p->dtype));
  const bool fuse_dx = p->dx != nullptr && pl.vec && p->Cout <= pl.co_t;     // one pass over x and dy for dx, dw, db
  if (p->dx != nullptr && !fuse_dx) {
    MEDNET_DISPATCH_TV(p->dtype, V, {
      conv1_dx_kernel<T, VV><<<grid_for(p->N * p->S, 128, 16), 128, smem, stream>>>(p->dy, p->w, (T*)p->dx, p->N, p->S,
                                                                                   p->Cin, p->Cout, (const T*)p->x, p->in_act,
                                                                                   p->in_act_param);
    });
    MEDNET_LAUNCH_CHECK();
  }
  float* partial = (float*)workspace;
  float* partial_b = (float*)((char*)workspace + align_up((size_t)pl.nblocks * p->Cout * p->Cin * sizeof(float), 256));
  if (pl.vec) {
    dim3 grid(pl.nblk_per_sample, (unsigned)p->N);
    const size_t smv = (size_t)8 * pl.ncol * pl.co_t * pl.V * sizeof(float) + (size_t)DW_STAGES * 256 * (16 + 4 * pl.co_t);
    for (int co0 = 0; co0 < p->Cout; co0 += pl.co_t) {
#define MEDNET_DWV(TT, VV, CT)                                                                                         \
      do {                                                                                                             \
        cudaError_t ea = cudaFuncSetAttribute(conv1_dw_vec_kernel<TT, VV, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                              (int)smv);                                                               \
        if (ea != cudaSuccess) return (int)ea;                                                                         \
        conv1_dw_vec_kernel<TT, VV, CT><<<grid, 256, smv, stream>>>((const TT*)p->x, p->dy, partial, p->S, p->Cin, p->Cout, \
                                                                    co0, pl.rows_per_block_vec, p->w,                  \
                                                                    fuse_dx ? (TT*)p->dx : nullptr, p->in_act, p->in_act_param); \
      } while (0)
      if (p->dtype == MEDNET_F32) {
        if (pl.co_t == 4) MEDNET_DWV(float, 4, 4); else if (pl.co_t == 8) MEDNET_DWV(float, 4, 8); else MEDNET_DWV(float, 4, 16);
      } else {
        if (pl.co_t == 4) MEDNET_DWV(bf16, 8, 4); else if (pl.co_t == 8) MEDNET_DWV(bf16, 8, 8); else MEDNET_DWV(bf16, 8, 16);
      }
#undef MEDNET_DWV
      MEDNET_LAUNCH_CHECK();
    }
    conv1_db_partial_kernel<<<grid, 256, 0, stream>>>(p->dy, partial_b, p->S, p->Cout, pl.rows_per_block_vec);
    MEDNET_LAUNCH_CHECK();
  } else {
    dim3 grid(pl.nblocks, pl.cin_tiles), block(pl.cin_t, pl.R);
    const size_t sm2 = (size_t)pl.cin_t * pl.R * DW_MAX_CO * sizeof(float);
    for (int co0 = 0; co0 < p->Cout; co0 += DW_MAX_CO) {
      if (p->dtype == MEDNET_F32)
        conv1_dw_partial_kernel<float><<<grid, block, sm2, stream>>>((const float*)p->x, p->dy, partial, p->N, p->S, p->Cin,
                                                                    p->Cout, co0, pl.rows_per_block);
      else
        conv1_dw_partial_kernel<bf16><<<grid, block, sm2, stream>>>((const bf16*)p->x, p->dy, partial, p->N, p->S, p->Cin,
                                                                   p->Cout, co0, pl.rows_per_block);
      MEDNET_LAUNCH_CHECK();
    }
    conv1_db_partial_rows_kernel<<<pl.nblocks, 256, 0, stream>>>(p->dy, partial_b, p->N, p->S, p->Cout, pl.rows_per_block);
    MEDNET_LAUNCH_CHECK();
  }
  const int tot = p->Cout * p->Cin + p->Cout;
  conv1_dw_final_kernel<<<ceil_div(tot, 128), 128, 0, stream>>>(partial, partial_b, p->dw, p->db, p->Cin, p->Cout,
                                                               pl.nblocks, p->accumulate);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}
