// CUDA-core (fp32 accumulate) gather implicit GEMM for the 3x3x3 convolutions: the fp32 validation
// path of the hot path and the route for shapes the tcgen05 kernels do not take (tiny channel counts,
// the transposed convolution).  One code path serves
//   * Conv3d k3 s1 p1 fprop and dgrad                     (ref: midasmednet/unet/components.py:8-9)
//   * ConvTranspose3d k3 s2 p1 op1 fprop and dgrad        (ref: midasmednet/unet/components.py:259-264)
//   * both weight gradients (deterministic split-K)
// through a `gather` rule mapping (row voxel, tap) -> source voxel.
#include "common.cuh"

namespace mednet {

struct Geom {
  int N, Di, Hi, Wi;   // gathered tensor
  int Do, Ho, Wo;      // row space
  int mode;
};

// source coordinate along one axis, or -1
__device__ __forceinline__ int gather_axis(int mode, int o, int k, int in_size) {
  int i;
  if (mode == MEDNET_GATHER_CONV3) {
    i = o + k - 1;
  } else if (mode == MEDNET_GATHER_CONVT_F) {
    const int t = o + 1 - k;
    if (t & 1) return -1;
    i = t >> 1;
  } else {
    i = 2 * o - 1 + k;
  }
  return (i >= 0 && i < in_size) ? i : -1;
}

__device__ __forceinline__ int64_t gather_src(const Geom& g, int n, int od, int oh, int ow, int tap) {
  const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
  const int id = gather_axis(g.mode, od, kd, g.Di);
  const int ih = gather_axis(g.mode, oh, kh, g.Hi);
  const int iw = gather_axis(g.mode, ow, kw, g.Wi);
  if ((id | ih | iw) < 0) return -1;
  return (((int64_t)n * g.Di + id) * g.Hi + ih) * g.Wi + iw;
}

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

// y[m][n] = act( sum_{tap,k} x[src(m,tap)][k] * w[n][tap][k] + bias[n] + addend[m][n] )
template <typename T>
__global__ void __launch_bounds__(256) igemm_fprop_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                          const float* __restrict__ bias,
                                                          const T* __restrict__ addend, T* __restrict__ y, Geom g,
                                                          int K, int Nout, int act, float act_param) {
  __shared__ float As[BK][BM + PAD];
  __shared__ float Bs[BK][BN + PAD];
  __shared__ int rinfo[BM][4];

  const int tid = threadIdx.x;
  const int64_t M = (int64_t)g.N * g.Do * g.Ho * g.Wo;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Kflat = 27 * K;

  if (tid < BM) {
    const int64_t m = m0 + tid;
    int n = -1, od = 0, oh = 0, ow = 0;
    if (m < M) {
      int64_t t = m;
      ow = (int)(t % g.Wo); t /= g.Wo;
      oh = (int)(t % g.Ho); t /= g.Ho;
      od = (int)(t % g.Do);
      n = (int)(t / g.Do);
    }
    rinfo[tid][0] = n; rinfo[tid][1] = od; rinfo[tid][2] = oh; rinfo[tid][3] = ow;
  }
  __syncthreads();

  const int lrow = tid >> 2, kq = (tid & 3) * 4;
  const int tx = tid & 15, ty = tid >> 4;
  const bool vec_ok = (K % 4) == 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int rn = rinfo[lrow][0], rd = rinfo[lrow][1], rh = rinfo[lrow][2], rw = rinfo[lrow][3];
  const int bn = n0 + lrow;

  for (int k0 = 0; k0 < Kflat; k0 += BK) {
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    const int kk = k0 + kq;
    if (vec_ok) {
      if (kk < Kflat) {
        const int tap = kk / K, ci = kk - tap * K;
        if (rn >= 0) {
          const int64_t s = gather_src(g, rn, rd, rh, rw, tap);
          if (s >= 0) load_vec<T, 4>(x + s * K + ci, a);
        }
        if (bn < Nout) load_vec<T, 4>(w + (int64_t)bn * Kflat + kk, b);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = kk + j;
        if (k < Kflat) {
          const int tap = k / K, ci = k - tap * K;
          if (rn >= 0) {
            const int64_t s = gather_src(g, rn, rd, rh, rw, tap);
            if (s >= 0) a[j] = to_f32<T>(x[s * K + ci]);
          }
          if (bn < Nout) b[j] = to_f32<T>(w[(int64_t)bn * Kflat + k]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[kq + j][lrow] = a[j];
      Bs[kq + j][lrow] = b[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= Nout) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += bias[n];
      if (addend != nullptr) v += to_f32<T>(addend[m * Nout + n]);
      y[m * Nout + n] = from_f32<T>(act_apply(v, act, act_param));
    }
  }
}

// partial[z][a][tap][b] = sum over the z-th slice of rows m of A[m][a] * Bg[src(m,tap)][b]
template <typename T>
__global__ void __launch_bounds__(256) igemm_wgrad_kernel(const T* __restrict__ A, const T* __restrict__ Bg,
                                                          float* __restrict__ partial, Geom g, int Ca, int Cb,
                                                          int64_t rows_per_split) {
  __shared__ float As[BK][BM + PAD];
  __shared__ float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)g.N * g.Do * g.Ho * g.Wo;
  const int a0 = blockIdx.x * BM;
  const int j0 = blockIdx.y * BN;
  const int J = 27 * Cb;
  const int64_t mbeg = (int64_t)blockIdx.z * rows_per_split;
  int64_t mend = mbeg + rows_per_split;
  if (mend > M) mend = M;

  const int mk = tid >> 4, q = (tid & 15) * 4;
  const int tx = tid & 15, ty = tid >> 4;
  const bool vec_a = (Ca % 4) == 0, vec_b = (Cb % 4) == 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // column decomposition of this thread's 4 gathered columns is loop invariant
  const int jj = j0 + q;
  int tapv[4], bv_[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = jj + j;
    tapv[j] = col < J ? col / Cb : -1;
    bv_[j] = col < J ? col - (col / Cb) * Cb : 0;
  }

  for (int64_t mb = mbeg; mb < mend; mb += BK) {
    const int64_t m = mb + mk;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < mend) {
      int64_t t = m;
      const int ow = (int)(t % g.Wo); t /= g.Wo;
      const int oh = (int)(t % g.Ho); t /= g.Ho;
      const int od = (int)(t % g.Do);
      const int n = (int)(t / g.Do);
      const int ac = a0 + q;
      if (vec_a && ac + 3 < Ca) {
        load_vec<T, 4>(A + m * Ca + ac, a);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ac + j < Ca) a[j] = to_f32<T>(A[m * Ca + ac + j]);
      }
      if (vec_b && tapv[0] >= 0) {   // Cb % 4 == 0 -> the 4 columns share one tap
        const int64_t s = gather_src(g, n, od, oh, ow, tapv[0]);
        if (s >= 0) load_vec<T, 4>(Bg + s * Cb + bv_[0], b);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (tapv[j] >= 0) {
            const int64_t s = gather_src(g, n, od, oh, ow, tapv[j]);
            if (s >= 0) b[j] = to_f32<T>(Bg[s * Cb + bv_[j]]);
          }
        }
      }
    }
    *reinterpret_cast<float4*>(&As[mk][q]) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(&Bs[mk][q]) = make_float4(b[0], b[1], b[2], b[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bw = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bw.x, bw.y, bw.z, bw.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* out = partial + (int64_t)blockIdx.z * Ca * J;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + ty * 4 + i;
    if (a >= Ca) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = j0 + tx * 4 + j;
      if (col < J) out[(int64_t)a * J + col] = acc[i][j];
    }
  }
}

// dw[a][b][tap] (PyTorch layout) (+)= sum_z partial[z][a][tap][b]   (fixed order -> deterministic)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int Ca, int Cb,
                                    int splits, int accumulate) {
  const int64_t total = (int64_t)Ca * Cb * 27;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 27);
    const int b = (int)((i / 27) % Cb);
    const int a = (int)(i / (27 * (int64_t)Cb));
    const int64_t src = ((int64_t)a * 27 + tap) * Cb + b;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(int64_t)z * total + src];
    dw[i] = accumulate ? dw[i] + s : s;
  }
}

// column sums of A[M][C] (bias gradients), two-stage deterministic
template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ A, float* __restrict__ partial, int64_t M, int C,
                                      int64_t rows_per_block) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += to_f32<T>(A[r * C + c]);
    partial[(int64_t)blockIdx.x * C + c] = s;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int C, int nblocks,
                                    int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[(int64_t)b * C + c];
  out[c] = accumulate ? out[c] + s : s;
}

template <typename T>
__global__ void pack_weights_kernel(const float* __restrict__ src, T* __restrict__ dst, int Cin, int Cout,
                                    int transposed, int dgrad, int tapmajor) {
  const int Nout = dgrad ? Cin : Cout, K = dgrad ? Cout : Cin;
  const int64_t total = (int64_t)Nout * 27 * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int nout, tap, k;
    if (tapmajor) {
      k = (int)(i % K);
      nout = (int)((i / K) % Nout);
      tap = (int)(i / ((int64_t)K * Nout));
    } else {
      k = (int)(i % K);
      tap = (int)((i / K) % 27);
      nout = (int)(i / ((int64_t)K * 27));
    }
    const int co = dgrad ? k : nout, ci = dgrad ? nout : k;
    const int ksrc = (dgrad && !transposed) ? 26 - tap : tap;
    const float v = transposed ? src[((int64_t)ci * Cout + co) * 27 + ksrc] : src[((int64_t)co * Cin + ci) * 27 + ksrc];
    dst[i] = from_f32<T>(v);
  }
}

int colsum_bias(const void* a, int dtype, int64_t M, int C, float* dbias, int accumulate, void* workspace,
                cudaStream_t st);

struct WgradPlan {
  int splits, colsum_blocks;
  int64_t rows_per_split, colsum_rows;
};

static WgradPlan wgrad_plan(const mednet_wgrad_params* p) {
  WgradPlan pl;
  const int64_t M = (int64_t)p->N * p->Da * p->Ha * p->Wa;
  const int64_t tiles = (int64_t)ceil_div(p->Ca, BM) * ceil_div(27 * p->Cb, BN);
  int64_t splits = ((int64_t)sm_count_cached() * 4) / tiles;
  const int64_t max_by_rows = M / (BK * 16);
  if (splits > max_by_rows) splits = max_by_rows;
  const int64_t per_split_bytes = (int64_t)p->Ca * 27 * p->Cb * 4;
  const int64_t max_by_mem = ((int64_t)256 << 20) / per_split_bytes;
  if (splits > max_by_mem) splits = max_by_mem;
  if (splits < 1) splits = 1;
  pl.rows_per_split = ceil_div64(ceil_div64(M, splits), BK) * BK;
  pl.splits = (int)ceil_div64(M, pl.rows_per_split);
  int64_t cb = M / 512;
  if (cb > 1024) cb = 1024;
  if (cb < 1) cb = 1;
  pl.colsum_rows = ceil_div64(M, cb);
  pl.colsum_blocks = (int)ceil_div64(M, pl.colsum_rows);
  return pl;
}

size_t colsum_workspace_bytes(const mednet_wgrad_params* p) {
  const int c = p->Ca > p->Cb ? p->Ca : p->Cb;
  return align_up((size_t)1024 * c * sizeof(float), 256);
}

size_t simt_wgrad_workspace_bytes(const mednet_wgrad_params* p) {
  WgradPlan pl = wgrad_plan(p);
  return align_up((size_t)pl.splits * p->Ca * 27 * p->Cb * sizeof(float), 256) + colsum_workspace_bytes(p);
}

template <typename T>
static int simt_wgrad_t(const mednet_wgrad_params* p, void* workspace, cudaStream_t st) {
  WgradPlan pl = wgrad_plan(p);
  Geom g{p->N, p->Db, p->Hb, p->Wb, p->Da, p->Ha, p->Wa, p->gather};
  float* partial = (float*)workspace;
  float* cpart = (float*)((char*)workspace + align_up((size_t)pl.splits * p->Ca * 27 * p->Cb * sizeof(float), 256));
  dim3 grid(ceil_div(p->Ca, BM), ceil_div(27 * p->Cb, BN), pl.splits);
  igemm_wgrad_kernel<T><<<grid, 256, 0, st>>>((const T*)p->a, (const T*)p->b, partial, g, p->Ca, p->Cb,
                                              pl.rows_per_split);
  MEDNET_LAUNCH_CHECK();
  const int64_t total = (int64_t)p->Ca * p->Cb * 27;
  wgrad_reduce_kernel<<<grid_for(total, 256), 256, 0, st>>>(partial, p->dw, p->Ca, p->Cb, pl.splits, p->accumulate);
  MEDNET_LAUNCH_CHECK();
  if (p->dbias != nullptr) {
    // bias gradient = column sums of the output-gradient operand (a for conv, b for transposed conv)
    if (p->gather == MEDNET_GATHER_CONV3)
      return colsum_bias(p->a, p->dtype, (int64_t)p->N * p->Da * p->Ha * p->Wa, p->Ca, p->dbias, p->accumulate, cpart, st);
    return colsum_bias(p->b, p->dtype, (int64_t)p->N * p->Db * p->Hb * p->Wb, p->Cb, p->dbias, p->accumulate, cpart, st);
  }
  return MEDNET_OK;
}

int simt_wgrad(const mednet_wgrad_params* p, void* workspace, cudaStream_t st) {
  return p->dtype == MEDNET_F32 ? simt_wgrad_t<float>(p, workspace, st) : simt_wgrad_t<bf16>(p, workspace, st);
}

// bias gradient helper shared with the tensor-core wgrad
int colsum_bias(const void* a, int dtype, int64_t M, int C, float* dbias, int accumulate, void* workspace,
                cudaStream_t st) {
  int64_t cb = M / 512;
  if (cb > 1024) cb = 1024;
  if (cb < 1) cb = 1;
  const int64_t rows = ceil_div64(M, cb);
  const int blocks = (int)ceil_div64(M, rows);
  float* cpart = (float*)workspace;
  if (dtype == MEDNET_F32)
    colsum_partial_kernel<float><<<blocks, 128, 0, st>>>((const float*)a, cpart, M, C, rows);
  else
    colsum_partial_kernel<bf16><<<blocks, 128, 0, st>>>((const bf16*)a, cpart, M, C, rows);
  MEDNET_LAUNCH_CHECK();
  colsum_final_kernel<<<ceil_div(C, 128), 128, 0, st>>>(cpart, dbias, C, blocks, accumulate);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

template <typename T>
static int simt_fprop_t(const mednet_conv3d_params* p, cudaStream_t st) {
  Geom g{p->N, p->Di, p->Hi, p->Wi, p->Do, p->Ho, p->Wo, p->gather};
  const int64_t M = (int64_t)p->N * p->Do * p->Ho * p->Wo;
  MEDNET_REQUIRE(ceil_div(p->Nout, BN) <= 65535, MEDNET_EUNSUPPORTED);
  dim3 grid((unsigned)ceil_div64(M, BM), ceil_div(p->Nout, BN));
  igemm_fprop_kernel<T><<<grid, 256, 0, st>>>((const T*)p->x, (const T*)p->w, p->bias, (const T*)p->addend, (T*)p->y, g,
                                              p->K, p->Nout, p->act, p->act_param);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

int simt_fprop(const mednet_conv3d_params* p, cudaStream_t st) {
  return p->dtype == MEDNET_F32 ? simt_fprop_t<float>(p, st) : simt_fprop_t<bf16>(p, st);
}

}  // namespace mednet

using namespace mednet;

extern "C" int mednet_conv3d_pack_weights(const mednet_wpack_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->w_oidhw && p->w_packed && p->Cin > 0 && p->Cout > 0, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(p->layout >= 0 && p->layout <= 3, MEDNET_EINVAL);
  const int dgrad = p->layout & 1, tapmajor = p->layout >> 1;
  const int64_t total = (int64_t)p->Cin * p->Cout * 27;
  if (p->dtype == MEDNET_F32)
    pack_weights_kernel<float><<<grid_for(total, 256), 256, 0, stream>>>((const float*)p->w_oidhw, (float*)p->w_packed,
                                                                        p->Cin, p->Cout, p->transposed, dgrad, tapmajor);
  else
    pack_weights_kernel<bf16><<<grid_for(total, 256), 256, 0, stream>>>((const float*)p->w_oidhw, (bf16*)p->w_packed,
                                                                       p->Cin, p->Cout, p->transposed, dgrad, tapmajor);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}
