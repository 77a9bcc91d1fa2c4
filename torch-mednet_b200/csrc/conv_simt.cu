// CUDA-core (fp32 accumulate) gather implicit GEMM for the 3x3x3 convolutions: the fp32 validation
// path of the hot path and the route for shapes the tcgen05 kernels do not take (tiny channel counts,
// the transposed convolution).  One code path serves
//   * Conv3d k3 s1 p1 fprop and dgrad                     (ref: midasmednet/unet/components.py:8-9)
//   * ConvTranspose3d k3 s2 p1 op1 fprop and dgrad        (ref: midasmednet/unet/components.py:259-264)
//   * both weight gradients (deterministic split-K)
// through a `gather` rule mapping (row voxel, tap) -> source voxel.
#include "common.cuh"
#include "conv_impl.h"

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
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;      // interleaved partial sums: independent loads in flight, fixed order
    int z = 0;
    for (; z + 3 < splits; z += 4) {
      s0 += partial[(int64_t)z * total + src];
      s1 += partial[(int64_t)(z + 1) * total + src];
      s2 += partial[(int64_t)(z + 2) * total + src];
      s3 += partial[(int64_t)(z + 3) * total + src];
    }
    for (; z < splits; ++z) s0 += partial[(int64_t)z * total + src];
    const float s = (s0 + s1) + (s2 + s3);
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

// Packed weights of the transposed convolution for the tensor-core kernel: one [27][Nout][K] set per parity class,
// indexed by WINDOW tap (offset + 1 per axis); inactive taps are zero (never loaded).  src is the ConvTranspose3d
// weight (Cin, Cout, 3, 3, 3).  k3 s2 p1 op1: o = 2i - 1 + k.
//   fprop  (dgrad = 0): y[2j + par] += x[j + off] W[k]:  par 0: (off 0, k 1);  par 1: (off 0, k 2), (off +1, k 0)
//   dgrad  (dgrad = 1): dx[i] += dy[2(i + off) + par] W[k]^T:  par 0: (off 0, k 1);  par 1: (off -1, k 0), (off 0, k 2)
template <typename T>
__global__ void pack_weights_convt_kernel(const float* __restrict__ src, T* __restrict__ dst, int Cin, int Cout, int dgrad) {
  const int Nout = dgrad ? Cin : Cout, K = dgrad ? Cout : Cin;
  const int64_t total = (int64_t)8 * 27 * Nout * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int nout = (int)((i / K) % Nout);
    const int tap = (int)((i / ((int64_t)K * Nout)) % 27);
    const int cls = (int)(i / ((int64_t)K * Nout * 27));
    const int kk[3] = {tap / 9, (tap / 3) % 3, tap % 3}, par[3] = {cls >> 2, (cls >> 1) & 1, cls & 1};
    int ksrc[3];
    bool on = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int off = kk[a] - 1;
      if (par[a] == 0) { on = on && off == 0; ksrc[a] = 1; }
      else if (!dgrad) { on = on && off >= 0; ksrc[a] = off == 0 ? 2 : 0; }
      else { on = on && off <= 0; ksrc[a] = off == 0 ? 2 : 0; }
    }
    if (!on) continue;           // inactive taps of a parity class are never loaded by the kernel (tap mask): not written
    const int ci = dgrad ? nout : k, co = dgrad ? k : nout;
    const float v = src[((int64_t)ci * Cout + co) * 27 + (ksrc[0] * 3 + ksrc[1]) * 3 + ksrc[2]];
    dst[i] = from_f32<T>(v);
  }
}

// Packed weights of the conv over a nearest-upsampled input (MEDNET_GATHER_UPCONV_*): one [27][Nout][K] set per parity
// class, indexed by COARSE window tap (offset + 1 per axis).  For v = 2u + p the fine tap o reads coarse voxel
// u + c(p, o), c(0, .) = (-1, 0, 0), c(1, .) = (0, 0, +1) for o = (-1, 0, +1): the fine taps that land on the same coarse
// voxel are summed.  src is the Conv3d weight (Cout, Cin, 3, 3, 3) of the upsampled input channels.
//   fprop (dgrad = 0): window offset = c;   data gradient (dgrad = 1): dX[u] += W_p[c]^T dY[2 (u - c) + p]: offset = -c
template <typename T>
__global__ void pack_weights_upconv_kernel(const float* __restrict__ src, T* __restrict__ dst, int Cin, int Cout, int dgrad) {
  const int Nout = dgrad ? Cin : Cout, K = dgrad ? Cout : Cin;
  // only the 2x2x2 window of each class is enumerated (8 of 27 taps): taps outside it are never loaded by the kernel
  // (tap mask) and are not written
  const int64_t total = (int64_t)8 * 8 * Nout * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int nout = (int)((i / K) % Nout);
    const int t8 = (int)((i / ((int64_t)K * Nout)) % 8);
    const int cls = (int)(i / ((int64_t)K * Nout * 8));
    const int bit[3] = {t8 >> 2, (t8 >> 1) & 1, t8 & 1}, par[3] = {cls >> 2, (cls >> 1) & 1, cls & 1};
    // per axis: coarse offset c = bit - 1 + par (par 0: -1, 0;  par 1: 0, +1), the fine taps (indices 0..2) that map to
    // it, and the window tap index it is stored under
    int lo[3], hi[3], kk[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int c = bit[a] - 1 + par[a];
      if (par[a] == 0) {
        if (c == -1) { lo[a] = 0; hi[a] = 0; }          // o = -1
        else { lo[a] = 1; hi[a] = 2; }                   // o = 0, +1
      } else {
        if (c == 0) { lo[a] = 0; hi[a] = 1; }            // o = -1, 0
        else { lo[a] = 2; hi[a] = 2; }                   // o = +1
      }
      kk[a] = dgrad ? 1 - c : c + 1;
    }
    float v = 0.f;
    {
      const int ci = dgrad ? nout : k, co = dgrad ? k : nout;
      const float* w = src + ((int64_t)co * Cin + ci) * 27;
      for (int od = lo[0]; od <= hi[0]; ++od)
        for (int oh = lo[1]; oh <= hi[1]; ++oh)
          for (int ow = lo[2]; ow <= hi[2]; ++ow) v += w[(od * 3 + oh) * 3 + ow];
    }
    const int tap = (kk[0] * 3 + kk[1]) * 3 + kk[2];
    dst[(((int64_t)cls * 27 + tap) * Nout + nout) * K + k] = from_f32<T>(v);
  }
}

int colsum_bias(const void* a, int dtype, int64_t M, int C, float* dbias, int accumulate, void* workspace,
                cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// Small-channel specialisations of the stride-1 3x3x3 convolution.  The first layer of every reference
// network has in_channels = 1 (midasmednet/segmentation.py:30-31, landmarks.py:30-31): its forward pass
// is a 27-point stencil producing 8..64 channels, its dgrad a 27 x C -> 1 reduction and its wgrad
// 27 x C accumulators over all voxels.  None of them is GEMM-shaped (K or N = 1), all three are bound by
// the one wide activation they read or write; the generic 64x64 implicit-GEMM tile wastes 63/64 of its
// work on them (measured 38 ms per launch at 8 x 128^3).
// ------------------------------------------------------------------------------------------------

// few INPUT channels (K <= 4): one thread per output voxel, NT output channels per thread
template <typename T, int NT>
__global__ void __launch_bounds__(128) conv3_fewin_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                          const float* __restrict__ bias, const T* __restrict__ addend,
                                                          T* __restrict__ y, int D, int H, int W, int K, int Nout, int act,
                                                          float act_param) {
  extern __shared__ float sw[];                       // [27*K][NT] for output channels n0 .. n0+NT
  const int n0 = blockIdx.z * NT, TK = 27 * K;
  for (int i = threadIdx.x; i < TK * NT; i += blockDim.x) {
    const int j = i % NT, tk = i / NT;
    sw[i] = to_f32<T>(w[(int64_t)(n0 + j) * TK + tk]);
  }
  __syncthreads();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= H * W) return;
  const int h = q / W, wq = q - h * W;
  const int n = blockIdx.y / D, d = blockIdx.y - n * D;
  float acc[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j] = bias != nullptr ? bias[n0 + j] : 0.f;
  for (int kd = 0; kd < 3; ++kd) {
    const int id = d + kd - 1;
    if (id < 0 || id >= D) continue;
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = h + kh - 1;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = wq + kw - 1;
        if (iw < 0 || iw >= W) continue;
        const T* xp = x + ((((int64_t)n * D + id) * H + ih) * W + iw) * K;
        const float* wr = sw + (size_t)((kd * 3 + kh) * 3 + kw) * K * NT;
        for (int k = 0; k < K; ++k) {
          const float xv = to_f32<T>(xp[k]);
#pragma unroll
          for (int j = 0; j < NT; j += 4) {
            const float4 wv = *reinterpret_cast<const float4*>(wr + k * NT + j);
            acc[j] = fmaf(xv, wv.x, acc[j]);
            acc[j + 1] = fmaf(xv, wv.y, acc[j + 1]);
            acc[j + 2] = fmaf(xv, wv.z, acc[j + 2]);
            acc[j + 3] = fmaf(xv, wv.w, acc[j + 3]);
          }
        }
      }
    }
  }
  const int64_t vox = (((int64_t)n * D + d) * H + h) * W + wq;
#pragma unroll
  for (int j = 0; j < NT; j += 8) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = acc[j + i];
    if (addend != nullptr) {
      float a8[8];
      load_vec<T, 8>(addend + vox * Nout + n0 + j, a8);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += a8[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = act_apply(o[i], act, act_param);
    store_vec<T, 8>(y + vox * Nout + n0 + j, o);
  }
}

// Register-tiled variant of the few-input-channel stencil: one thread computes 16 output channels of FOUR consecutive
// w-voxels, so every weight vector read from shared memory (LDS.128 = 4 output channels) feeds 16 FMAs instead of 4
// and the six input voxels of a (kd, kh) row are loaded once.  Grid: (groups of 4 voxels per (h) row) x (n, d) x Nout/16.
template <typename T>
__global__ void __launch_bounds__(128) conv3_fewin4_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                           const float* __restrict__ bias, const T* __restrict__ addend,
                                                           T* __restrict__ y, int D, int H, int W, int K, int Nout, int act,
                                                           float act_param) {
  constexpr int NT = 16;
  extern __shared__ float sw[];                       // [27*K][NT] for output channels n0 .. n0+NT
  const int n0 = blockIdx.z * NT, TK = 27 * K;
  for (int i = threadIdx.x; i < TK * NT; i += blockDim.x) {
    const int j = i % NT, tk = i / NT;
    sw[i] = to_f32<T>(w[(int64_t)(n0 + j) * TK + tk]);
  }
  __syncthreads();
  const int WG = (W + 3) >> 2;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= H * WG) return;
  const int h = q / WG, w0 = (q - h * WG) * 4;
  const int n = blockIdx.y / D, d = blockIdx.y - n * D;
  float acc[4][NT];
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[v][j] = bias != nullptr ? bias[n0 + j] : 0.f;
  for (int kd = 0; kd < 3; ++kd) {
    const int id = d + kd - 1;
    if (id < 0 || id >= D) continue;
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = h + kh - 1;
      if (ih < 0 || ih >= H) continue;
      const T* xrow = x + (((int64_t)n * D + id) * H + ih) * (int64_t)W * K;
      for (int k = 0; k < K; ++k) {
        float xv[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int iw = w0 - 1 + i;
          xv[i] = (iw >= 0 && iw < W) ? to_f32<T>(xrow[(int64_t)iw * K + k]) : 0.f;
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float* wr = sw + (size_t)(((kd * 3 + kh) * 3 + kw) * K + k) * NT;
#pragma unroll
          for (int j = 0; j < NT; j += 4) {
            const float4 wv = *reinterpret_cast<const float4*>(wr + j);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              acc[v][j] = fmaf(xv[v + kw], wv.x, acc[v][j]);
              acc[v][j + 1] = fmaf(xv[v + kw], wv.y, acc[v][j + 1]);
              acc[v][j + 2] = fmaf(xv[v + kw], wv.z, acc[v][j + 2]);
              acc[v][j + 3] = fmaf(xv[v + kw], wv.w, acc[v][j + 3]);
            }
          }
        }
      }
    }
  }
  const int64_t vox0 = (((int64_t)n * D + d) * H + h) * W + w0;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    if (w0 + v >= W) break;
#pragma unroll
    for (int j = 0; j < NT; j += 8) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = acc[v][j + i];
      if (addend != nullptr) {
        float a8[8];
        load_vec<T, 8>(addend + (vox0 + v) * Nout + n0 + j, a8);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += a8[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = act_apply(o[i], act, act_param);
      store_vec<T, 8>(y + (vox0 + v) * Nout + n0 + j, o);
    }
  }
}

// few OUTPUT channels (Nout <= 4, K = 8 * TPV): TPV threads per voxel, each owns 8 input channels (one 16-byte
// load per tap: a warp reads whole cache lines), partial dot products combined with warp shuffles
template <typename T, int TPV>
__global__ void __launch_bounds__(128) conv3_fewout_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                           const float* __restrict__ bias, const T* __restrict__ addend,
                                                           T* __restrict__ y, int D, int H, int W, int Nout, int act,
                                                           float act_param) {
  constexpr int K = 8 * TPV;
  extern __shared__ float sw[];                       // [Nout][27][K]
  for (int i = threadIdx.x; i < Nout * 27 * K; i += blockDim.x) sw[i] = to_f32<T>(w[i]);
  __syncthreads();
  const int sub = threadIdx.x % TPV;
  const int q = blockIdx.x * (128 / TPV) + threadIdx.x / TPV;
  const bool valid = q < H * W;
  const int h = valid ? q / W : 0, wq = valid ? q - h * W : 0;
  const int n = blockIdx.y / D, d = blockIdx.y - n * D;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (valid) {
    for (int kd = 0; kd < 3; ++kd) {
      const int id = d + kd - 1;
      if (id < 0 || id >= D) continue;
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = h + kh - 1;
        if (ih < 0 || ih >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = wq + kw - 1;
          if (iw < 0 || iw >= W) continue;
          float xv[8];
          load_vec<T, 8>(x + ((((int64_t)n * D + id) * H + ih) * W + iw) * K + sub * 8, xv);
          const int tap = (kd * 3 + kh) * 3 + kw;
          for (int o = 0; o < Nout; ++o) {
            const float* wr = sw + ((size_t)o * 27 + tap) * K + sub * 8;
            const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
            acc[o] = fmaf(xv[0], w0.x, fmaf(xv[1], w0.y, fmaf(xv[2], w0.z, fmaf(xv[3], w0.w, acc[o]))));
            acc[o] = fmaf(xv[4], w1.x, fmaf(xv[5], w1.y, fmaf(xv[6], w1.z, fmaf(xv[7], w1.w, acc[o]))));
          }
        }
      }
    }
  }
#pragma unroll
  for (int off = TPV / 2; off > 0; off >>= 1)
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
  if (valid && sub == 0) {
    const int64_t vox = (((int64_t)n * D + d) * H + h) * W + wq;
    for (int o = 0; o < Nout; ++o) {
      float v = acc[o];
      if (bias != nullptr) v += bias[o];
      if (addend != nullptr) v += to_f32<T>(addend[vox * Nout + o]);
      y[vox * Nout + o] = from_f32<T>(act_apply(v, act, act_param));
    }
  }
}


// few OUTPUT channels, register-tiled along w: TPV threads per group of 4 consecutive output voxels, each thread owns 8
// input channels.  Per (kd, kh) the 6 input voxels of the group are loaded once (16 B each) and every weight vector
// (2 x LDS.128) is reused by the 4 outputs -> 96 * NO FMAs for 6 global + 6 * NO shared loads.
template <typename T, int TPV, int NO>
__global__ void __launch_bounds__(128) conv3_fewout4_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                            const float* __restrict__ bias, const T* __restrict__ addend,
                                                            T* __restrict__ y, int D, int H, int W, int act,
                                                            float act_param) {
  constexpr int K = 8 * TPV;
  extern __shared__ float sw[];                       // [NO][27][K]
  for (int i = threadIdx.x; i < NO * 27 * K; i += blockDim.x) sw[i] = to_f32<T>(w[i]);
  __syncthreads();
  const int sub = threadIdx.x % TPV;
  const int WG = (W + 3) >> 2;
  const int g = blockIdx.x * (128 / TPV) + threadIdx.x / TPV;
  const bool valid = g < H * WG;
  const int h = valid ? g / WG : 0, w0 = valid ? (g - h * WG) * 4 : 0;
  const int n = blockIdx.y / D, d = blockIdx.y - n * D;
  float acc[4][NO];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int o = 0; o < NO; ++o) acc[j][o] = 0.f;
  if (valid) {
    for (int kd = 0; kd < 3; ++kd) {
      const int id = d + kd - 1;
      if (id < 0 || id >= D) continue;
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = h + kh - 1;
        if (ih < 0 || ih >= H) continue;
        const T* xrow = x + ((((int64_t)n * D + id) * H + ih) * W) * K + sub * 8;
        float xv[6][8];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int iw = w0 - 1 + i;
          if (iw >= 0 && iw < W) {
            load_vec<T, 8>(xrow + (int64_t)iw * K, xv[i]);
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) xv[i][c] = 0.f;
          }
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int tap = (kd * 3 + kh) * 3 + kw;
#pragma unroll
          for (int o = 0; o < NO; ++o) {
            const float* wr = sw + ((size_t)o * 27 + tap) * K + sub * 8;
            const float4 w0v = *reinterpret_cast<const float4*>(wr), w1v = *reinterpret_cast<const float4*>(wr + 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float* xx = xv[j + kw];
              float a = acc[j][o];
              a = fmaf(xx[0], w0v.x, a); a = fmaf(xx[1], w0v.y, a); a = fmaf(xx[2], w0v.z, a); a = fmaf(xx[3], w0v.w, a);
              a = fmaf(xx[4], w1v.x, a); a = fmaf(xx[5], w1v.y, a); a = fmaf(xx[6], w1v.z, a); a = fmaf(xx[7], w1v.w, a);
              acc[j][o] = a;
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int off = TPV / 2; off > 0; off >>= 1)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int o = 0; o < NO; ++o) acc[j][o] += __shfl_xor_sync(0xffffffffu, acc[j][o], off);
  if (valid && sub == 0) {
    const int64_t vox0 = (((int64_t)n * D + d) * H + h) * W + w0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (w0 + j < W) {
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          float v = acc[j][o];
          if (bias != nullptr) v += bias[o];
          if (addend != nullptr) v += to_f32<T>(addend[(vox0 + j) * NO + o]);
          y[(vox0 + j) * NO + o] = from_f32<T>(act_apply(v, act, act_param));
        }
      }
    }
  }
}

// weight gradient with ONE gathered channel (the in_channels = 1 first layer), register-tiled: a block stages a
// 64-voxel w-segment of dY transposed ([Ca][64], fp32) and the 9 (kd, kh) halo rows of x; thread (ca, vg) owns output
// channel ca and the 4*PIECES consecutive voxels of group vg, and for each halo row slides a 3-quad window over x so one
// LDS.128 of x and one of dY feed 12 FMAs.  partial[block][ca][27].
constexpr int WSEG = 64;
constexpr int XS_PITCH = WSEG + 8, DYS_PITCH = WSEG + 4;
template <typename T, int PIECES>
__global__ void __launch_bounds__(256) wgrad_fewin1_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                           float* __restrict__ partial, int N, int D, int H, int W, int Ca) {
  extern __shared__ float sm[];
  float* dys = sm;                                    // [Ca][DYS_PITCH]
  float* xs = sm + Ca * DYS_PITCH;                    // [9][XS_PITCH]: element i holds x at iw = w0 + i - 4
  const int ca = threadIdx.x % Ca, vg = threadIdx.x / Ca;
  const int v0 = vg * 4 * PIECES;
  const int wchunks = (W + WSEG - 1) / WSEG;
  const int64_t segs = (int64_t)N * D * H * wchunks;
  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  const int c8n = Ca / 8;
  for (int64_t seg = blockIdx.x; seg < segs; seg += gridDim.x) {
    int64_t t = seg;
    const int w0 = (int)(t % wchunks) * WSEG; t /= wchunks;
    const int h = (int)(t % H); t /= H;
    const int d = (int)(t % D);
    const int n = (int)(t / D);
    const int64_t rowbase = (((int64_t)n * D + d) * H + h) * W + w0;
    for (int i = threadIdx.x; i < WSEG * c8n; i += 256) {
      const int v = i / c8n, c8 = i - v * c8n;
      float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (w0 + v < W) load_vec<T, 8>(dy + (rowbase + v) * Ca + c8 * 8, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) dys[(c8 * 8 + k) * DYS_PITCH + v] = g[k];
    }
    for (int i = threadIdx.x; i < 9 * XS_PITCH; i += 256) {
      const int r = i / XS_PITCH, wv = i - r * XS_PITCH;
      const int id = d + r / 3 - 1, ih = h + r % 3 - 1, iw = w0 + wv - 4;
      float v = 0.f;
      if (id >= 0 && id < D && ih >= 0 && ih < H && iw >= 0 && iw < W)
        v = to_f32<T>(x[(((int64_t)n * D + id) * H + ih) * W + iw]);
      xs[i] = v;
    }
    __syncthreads();
    const float4* gq = reinterpret_cast<const float4*>(dys + ca * DYS_PITCH + v0);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const float4* xq = reinterpret_cast<const float4*>(xs + r * XS_PITCH + v0);      // quad 0 = iw w0+v0-4 .. -1
      float4 qp = xq[0], qc = xq[1];
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int p = 0; p < PIECES; ++p) {
        const float4 qn = xq[p + 2];
        const float4 g = gq[p];
        a0 = fmaf(g.x, qp.w, fmaf(g.y, qc.x, fmaf(g.z, qc.y, fmaf(g.w, qc.z, a0))));
        a1 = fmaf(g.x, qc.x, fmaf(g.y, qc.y, fmaf(g.z, qc.z, fmaf(g.w, qc.w, a1))));
        a2 = fmaf(g.x, qc.y, fmaf(g.y, qc.z, fmaf(g.z, qc.w, fmaf(g.w, qn.x, a2))));
        qp = qc;
        qc = qn;
      }
      acc[r * 3 + 0] += a0;
      acc[r * 3 + 1] += a1;
      acc[r * 3 + 2] += a2;
    }
    __syncthreads();
  }
  // combine the voxel groups: sm reused as [TG][Ca][27]
  const int TG = 256 / Ca;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 27; ++i) sm[((size_t)vg * Ca + ca) * 27 + i] = acc[i];
  __syncthreads();
  for (int i = threadIdx.x; i < Ca * 27; i += 256) {
    float a = 0.f;
    for (int tg = 0; tg < TG; ++tg) a += sm[(size_t)tg * Ca * 27 + i];
    partial[(int64_t)blockIdx.x * Ca * 27 + i] = a;
  }
}

// weight gradient with few GATHERED channels (Cb <= 4): partial[block][ca][tap][cb]; a block walks 64-voxel
// w-segments, thread (ca, tg) owns output channel ca and taps tg, tg+TG, ... (TG = 256 / Ca)
constexpr int FEWIN_MAX_OWN = 7;
template <typename T, int CB>
__global__ void __launch_bounds__(256) wgrad_fewin_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                          float* __restrict__ partial, int N, int D, int H, int W, int Ca) {
  extern __shared__ float sm[];
  float* dys = sm;                                    // [WSEG][Ca]
  float* xs = sm + WSEG * Ca;                         // [9][WSEG + 2][CB]
  const int TG = 256 / Ca;
  const int ca = threadIdx.x % Ca, tg = threadIdx.x / Ca;
  const int wchunks = (W + WSEG - 1) / WSEG;
  const int64_t segs = (int64_t)N * D * H * wchunks;
  float acc[FEWIN_MAX_OWN][CB];
  int xoff[FEWIN_MAX_OWN];
#pragma unroll
  for (int i = 0; i < FEWIN_MAX_OWN; ++i) {
    const int t = tg + i * TG;
    xoff[i] = t < 27 ? ((t / 3) * (WSEG + 2) + (t % 3)) * CB : -1;
#pragma unroll
    for (int c = 0; c < CB; ++c) acc[i][c] = 0.f;
  }
  const int c8n = Ca / 8;
  for (int64_t seg = blockIdx.x; seg < segs; seg += gridDim.x) {
    int64_t t = seg;
    const int w0 = (int)(t % wchunks) * WSEG; t /= wchunks;
    const int h = (int)(t % H); t /= H;
    const int d = (int)(t % D);
    const int n = (int)(t / D);
    const int64_t rowbase = (((int64_t)n * D + d) * H + h) * W + w0;
    for (int i = threadIdx.x; i < WSEG * c8n; i += 256) {
      const int v = i / c8n, c8 = i - v * c8n;
      float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (w0 + v < W) load_vec<T, 8>(dy + (rowbase + v) * Ca + c8 * 8, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) dys[v * Ca + c8 * 8 + k] = g[k];
    }
    for (int i = threadIdx.x; i < 9 * (WSEG + 2) * CB; i += 256) {
      const int c = i % CB, wv = (i / CB) % (WSEG + 2), r = i / (CB * (WSEG + 2));
      const int id = d + r / 3 - 1, ih = h + r % 3 - 1, iw = w0 + wv - 1;
      float v = 0.f;
      if (id >= 0 && id < D && ih >= 0 && ih < H && iw >= 0 && iw < W)
        v = to_f32<T>(x[((((int64_t)n * D + id) * H + ih) * W + iw) * CB + c]);
      xs[i] = v;
    }
    __syncthreads();
    for (int v = 0; v < WSEG; ++v) {
      const float g = dys[v * Ca + ca];
#pragma unroll
      for (int i = 0; i < FEWIN_MAX_OWN; ++i) {
        if (xoff[i] >= 0) {
#pragma unroll
          for (int c = 0; c < CB; ++c) acc[i][c] = fmaf(g, xs[xoff[i] + v * CB + c], acc[i][c]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < FEWIN_MAX_OWN; ++i) {
    const int tp = tg + i * TG;
    if (tp < 27) {
#pragma unroll
      for (int c = 0; c < CB; ++c) partial[(((int64_t)blockIdx.x * Ca + ca) * 27 + tp) * CB + c] = acc[i][c];
    }
  }
}

static bool fewin_fprop_ok(const mednet_conv3d_params* p) {
  return p->gather == MEDNET_GATHER_CONV3 && p->K <= 4 && p->Nout % 8 == 0 && p->Nout <= 256 &&
         (int64_t)p->N * p->Do <= 65535;
}
static bool fewout_fprop_ok(const mednet_conv3d_params* p) {
  return p->gather == MEDNET_GATHER_CONV3 && p->Nout <= 4 && (p->K == 8 || p->K == 16 || p->K == 32 || p->K == 64) &&
         (int64_t)p->N * p->Do <= 65535;
}
static bool fewin_wgrad_ok(const mednet_wgrad_params* p) {
  return p->gather == MEDNET_GATHER_CONV3 && (p->Cb == 1 || p->Cb == 2 || p->Cb == 4) &&
         (p->Ca == 8 || p->Ca == 16 || p->Ca == 32 || p->Ca == 64);
}
static int fewin_wgrad_blocks(const mednet_wgrad_params* p) {
  const int64_t segs = (int64_t)p->N * p->Da * p->Ha * ceil_div(p->Wa, WSEG);
  const int64_t cap = (int64_t)sm_count_cached() * 4;
  return (int)(segs < cap ? segs : cap);
}


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
  if (fewin_wgrad_ok(p))
    return align_up((size_t)fewin_wgrad_blocks(p) * p->Ca * 27 * p->Cb * sizeof(float), 256) + colsum_workspace_bytes(p);
  WgradPlan pl = wgrad_plan(p);
  return align_up((size_t)pl.splits * p->Ca * 27 * p->Cb * sizeof(float), 256) + colsum_workspace_bytes(p);
}

template <typename T>
static int fewin_wgrad_t(const mednet_wgrad_params* p, void* workspace, cudaStream_t st) {
  const int blocks = fewin_wgrad_blocks(p);
  float* partial = (float*)workspace;
  const size_t pbytes = align_up((size_t)blocks * p->Ca * 27 * p->Cb * sizeof(float), 256);
  const size_t smem = ((size_t)WSEG * p->Ca + 9 * (WSEG + 2) * p->Cb) * sizeof(float);
  if (in1_mma_wgrad_ok(p)) {                                  // bf16, one gathered channel: warp-level tensor cores
    int nb = blocks;
    const int r = in1_mma_wgrad(p, partial, blocks, &nb, st);
    if (r != MEDNET_OK) return r;
    const int64_t tot = (int64_t)p->Ca * 27;
    wgrad_reduce_kernel<<<grid_for(tot, 256), 256, 0, st>>>(partial, p->dw, p->Ca, p->Cb, nb, p->accumulate);
    MEDNET_LAUNCH_CHECK();
    if (p->dbias != nullptr)
      return colsum_bias(p->a, p->dtype, (int64_t)p->N * p->Da * p->Ha * p->Wa, p->Ca, p->dbias, p->accumulate,
                         (char*)workspace + pbytes, st);
    return MEDNET_OK;
  }
  if (p->Cb == 1 && p->Ca >= 16) {
    // chunk of WSEG / (256 / Ca) = Ca / 4 voxels per thread = PIECES quads
    size_t sm1 = ((size_t)p->Ca * DYS_PITCH + 9 * XS_PITCH) * sizeof(float);
    const size_t sm_red = (size_t)256 * 27 * sizeof(float);
    if (sm1 < sm_red) sm1 = sm_red;
    if (p->Ca == 16)
      wgrad_fewin1_kernel<T, 1><<<blocks, 256, sm1, st>>>((const T*)p->a, (const T*)p->b, partial, p->N, p->Da, p->Ha, p->Wa, p->Ca);
    else if (p->Ca == 32)
      wgrad_fewin1_kernel<T, 2><<<blocks, 256, sm1, st>>>((const T*)p->a, (const T*)p->b, partial, p->N, p->Da, p->Ha, p->Wa, p->Ca);
    else
      wgrad_fewin1_kernel<T, 4><<<blocks, 256, sm1, st>>>((const T*)p->a, (const T*)p->b, partial, p->N, p->Da, p->Ha, p->Wa, p->Ca);
  } else if (p->Cb == 1)
    wgrad_fewin_kernel<T, 1><<<blocks, 256, smem, st>>>((const T*)p->a, (const T*)p->b, partial, p->N, p->Da, p->Ha, p->Wa, p->Ca);
  else if (p->Cb == 2)
    wgrad_fewin_kernel<T, 2><<<blocks, 256, smem, st>>>((const T*)p->a, (const T*)p->b, partial, p->N, p->Da, p->Ha, p->Wa, p->Ca);
  else
    wgrad_fewin_kernel<T, 4><<<blocks, 256, smem, st>>>((const T*)p->a, (const T*)p->b, partial, p->N, p->Da, p->Ha, p->Wa, p->Ca);
  MEDNET_LAUNCH_CHECK();
  const int64_t total = (int64_t)p->Ca * p->Cb * 27;
  wgrad_reduce_kernel<<<grid_for(total, 256), 256, 0, st>>>(partial, p->dw, p->Ca, p->Cb, blocks, p->accumulate);
  MEDNET_LAUNCH_CHECK();
  if (p->dbias != nullptr)
    return colsum_bias(p->a, p->dtype, (int64_t)p->N * p->Da * p->Ha * p->Wa, p->Ca, p->dbias, p->accumulate,
                       (char*)workspace + pbytes, st);
  return MEDNET_OK;
}

template <typename T>
static int simt_wgrad_t(const mednet_wgrad_params* p, void* workspace, cudaStream_t st) {
  if (fewin_wgrad_ok(p)) return fewin_wgrad_t<T>(p, workspace, st);
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
static int small_fprop_t(const mednet_conv3d_params* p, cudaStream_t st) {
  if (in1_mma_fprop_ok(p)) return in1_mma_fprop(p, st);      // bf16, one input channel: warp-level tensor cores
  const int HW = p->Ho * p->Wo;
  if (fewin_fprop_ok(p) && p->Nout % 16 == 0 && p->Wo >= 4) {
    dim3 grid(ceil_div(p->Ho * ceil_div(p->Wo, 4), 128), (unsigned)(p->N * p->Do), p->Nout / 16);
    const size_t smem = (size_t)27 * p->K * 16 * sizeof(float);
    conv3_fewin4_kernel<T><<<grid, 128, smem, st>>>((const T*)p->x, (const T*)p->w, p->bias, (const T*)p->addend, (T*)p->y,
                                                    p->Do, p->Ho, p->Wo, p->K, p->Nout, p->act, p->act_param);
  } else if (fewin_fprop_ok(p)) {
    const int NT = p->Nout % 32 == 0 ? 32 : (p->Nout % 16 == 0 ? 16 : 8);
    dim3 grid(ceil_div(HW, 128), (unsigned)(p->N * p->Do), p->Nout / NT);
    const size_t smem = (size_t)27 * p->K * NT * sizeof(float);
#define MEDNET_FEWIN(NTV)                                                                                            \
    conv3_fewin_kernel<T, NTV><<<grid, 128, smem, st>>>((const T*)p->x, (const T*)p->w, p->bias, (const T*)p->addend, \
                                                        (T*)p->y, p->Do, p->Ho, p->Wo, p->K, p->Nout, p->act, p->act_param)
    if (NT == 32) MEDNET_FEWIN(32); else if (NT == 16) MEDNET_FEWIN(16); else MEDNET_FEWIN(8);
#undef MEDNET_FEWIN
  } else {
    const int TPV = p->K / 8;
    const size_t smem = (size_t)p->Nout * 27 * p->K * sizeof(float);
    if (p->Wo >= 4) {
      dim3 grid4(ceil_div(p->Ho * ceil_div(p->Wo, 4), 128 / TPV), (unsigned)(p->N * p->Do));
#define MEDNET_FEWOUT4(TP, NO)                                                                                              \
      conv3_fewout4_kernel<T, TP, NO><<<grid4, 128, smem, st>>>((const T*)p->x, (const T*)p->w, p->bias, (const T*)p->addend, \
                                                                (T*)p->y, p->Do, p->Ho, p->Wo, p->act, p->act_param)
#define MEDNET_FEWOUT4_NO(TP)                                                                          \
      do {                                                                                             \
        if (p->Nout == 1) MEDNET_FEWOUT4(TP, 1); else if (p->Nout == 2) MEDNET_FEWOUT4(TP, 2);         \
        else if (p->Nout == 3) MEDNET_FEWOUT4(TP, 3); else MEDNET_FEWOUT4(TP, 4);                      \
      } while (0)
      if (TPV == 1) MEDNET_FEWOUT4_NO(1); else if (TPV == 2) MEDNET_FEWOUT4_NO(2); else if (TPV == 4) MEDNET_FEWOUT4_NO(4);
      else MEDNET_FEWOUT4_NO(8);
#undef MEDNET_FEWOUT4_NO
#undef MEDNET_FEWOUT4
      MEDNET_LAUNCH_CHECK();
      return MEDNET_OK;
    }
    dim3 grid(ceil_div(HW, 128 / TPV), (unsigned)(p->N * p->Do));
#define MEDNET_FEWOUT(TP)                                                                                             \
    conv3_fewout_kernel<T, TP><<<grid, 128, smem, st>>>((const T*)p->x, (const T*)p->w, p->bias, (const T*)p->addend, \
                                                        (T*)p->y, p->Do, p->Ho, p->Wo, p->Nout, p->act, p->act_param)
    if (TPV == 1) MEDNET_FEWOUT(1); else if (TPV == 2) MEDNET_FEWOUT(2); else if (TPV == 4) MEDNET_FEWOUT(4); else MEDNET_FEWOUT(8);
#undef MEDNET_FEWOUT
  }
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

template <typename T>
static int simt_fprop_t(const mednet_conv3d_params* p, cudaStream_t st) {
  if (fewin_fprop_ok(p) || fewout_fprop_ok(p)) return small_fprop_t<T>(p, st);
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
  MEDNET_REQUIRE(p->layout >= 0 && p->layout <= MEDNET_WPACK_TC_UPCONV_B, MEDNET_EINVAL);
  if (p->layout == MEDNET_WPACK_TC_UPCONV_F || p->layout == MEDNET_WPACK_TC_UPCONV_B) {
    MEDNET_REQUIRE(!p->transposed, MEDNET_EINVAL);
    const int dg = p->layout == MEDNET_WPACK_TC_UPCONV_B ? 1 : 0;
    const int64_t tot = (int64_t)8 * 8 * p->Cin * p->Cout;
    if (p->dtype == MEDNET_F32)
      pack_weights_upconv_kernel<float><<<grid_for(tot, 256), 256, 0, stream>>>((const float*)p->w_oidhw, (float*)p->w_packed,
                                                                               p->Cin, p->Cout, dg);
    else
      pack_weights_upconv_kernel<bf16><<<grid_for(tot, 256), 256, 0, stream>>>((const float*)p->w_oidhw, (bf16*)p->w_packed,
                                                                              p->Cin, p->Cout, dg);
    MEDNET_LAUNCH_CHECK();
    return MEDNET_OK;
  }
  if (p->layout >= MEDNET_WPACK_TC_CONVT_F) {
    MEDNET_REQUIRE(p->transposed, MEDNET_EINVAL);
    const int dg = p->layout == MEDNET_WPACK_TC_CONVT_B ? 1 : 0;
    const int64_t tot = (int64_t)8 * 27 * p->Cin * p->Cout;
    if (p->dtype == MEDNET_F32)
      pack_weights_convt_kernel<float><<<grid_for(tot, 256), 256, 0, stream>>>((const float*)p->w_oidhw, (float*)p->w_packed,
                                                                              p->Cin, p->Cout, dg);
    else
      pack_weights_convt_kernel<bf16><<<grid_for(tot, 256), 256, 0, stream>>>((const float*)p->w_oidhw, (bf16*)p->w_packed,
                                                                             p->Cin, p->Cout, dg);
    MEDNET_LAUNCH_CHECK();
    return MEDNET_OK;
  }
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
