// GroupNorm (+ residual + activation) forward/backward and stand-alone activations, NDHWC.
// Replaces nn.GroupNorm / ReLU / LeakyReLU / ELU call sites of midasmednet/unet/components.py:35-57
// and the residual add + non-linearity of :177-178.  All kernels are HBM-bound: 16-byte vector
// accesses along the channel axis, per-thread register partials, deterministic two-stage reductions.
#include "common.cuh"

namespace mednet {

constexpr int UNR = 4;    // rows in flight per thread in the concat-grid kernels (few, fat blocks: memory-level parallelism)
constexpr int UAPP = 1;   // plain GroupNorm apply: one row per iteration
constexpr int USTD = 4;   // plain GroupNorm kernels: 4 rows in flight in the statistics pass, USTD / 2 per stream in the backward passes

// ------------------------------------------------------------------------------------------------
// launch plan shared by stats / apply / backward kernels: block = (ncol_t, R) threads, thread (tx,ty)
// owns channel vector `tx` and walks rows ty, ty+R, ... of its slab -> a block reads R consecutive
// NDHWC rows per step (fully coalesced) and every thread always sees the same channels.
// ------------------------------------------------------------------------------------------------
struct SlabPlan {
  int V, ncol, ncol_t, coltiles, R, nslab;
  int64_t rows_per_slab;
};

static SlabPlan make_plan(int64_t N, int64_t S, int C, int elem_bytes, int force_v = 0, int slabs_per_sm = 6) {
  SlabPlan p;
  p.V = force_v > 0 ? force_v : pick_vec(C, elem_bytes);
  p.ncol = C / p.V;
  if (p.ncol <= 256) {
    p.ncol_t = p.ncol;
    p.coltiles = 1;
    p.R = 256 / p.ncol;
    if (p.R < 1) p.R = 1;
  } else {
    p.ncol_t = 256;
    p.coltiles = ceil_div(p.ncol, 256);
    p.R = 1;
  }
  // The slab count -- and with it the order in which a sample's statistics are summed -- depends on (S, C) only, NOT on
  // the batch size: a sample gives bit-identical results whatever batch it is evaluated in (the sliding-window predictor
  // batches tiles freely, examples/predict.py).  Sized for a batch of two filling the GPU; larger batches just launch
  // more (still >= 4 row steps long) slabs.
  (void)N;
  int64_t target = (int64_t)sm_count_cached() * slabs_per_sm;
  int64_t nslab = target / (2 * p.coltiles);
  int64_t max_slab = S / ((int64_t)p.R * 4);
  if (nslab > max_slab) nslab = max_slab;
  if (nslab < 1) nslab = 1;
  p.rows_per_slab = ceil_div64(S, nslab);
  p.nslab = (int)ceil_div64(S, p.rows_per_slab);
  return p;
}

// reduce NV per-thread floats across threadIdx.y; the result for value i lands in the thread with
// (i % RED_CHUNK % R == ty) which calls emit(i, total).  Done in chunks of RED_CHUNK values so that the scratch is
// blockDim * RED_CHUNK floats (4 KB for 256 threads) whatever NV is: the statistics kernels then fit beside a resident
// tensor-core CTA (224 KB of the SM's 228 KB) and can run UNDER the asynchronous weight gradients instead of waiting for
// an SM to drain.  Fixed summation order: deterministic.
constexpr int RED_CHUNK = 4;
template <int NV, typename Emit>
__device__ __forceinline__ void reduce_over_y(const float (&vals)[NV], float* sm, Emit emit) {
  const int bx = blockDim.x, R = blockDim.y, tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int c0 = 0; c0 < NV; c0 += RED_CHUNK) {
    if (c0 > 0) __syncthreads();
#pragma unroll
    for (int j = 0; j < RED_CHUNK; ++j)
      if (c0 + j < NV) sm[(ty * bx + tx) * RED_CHUNK + j] = vals[c0 + j];
    __syncthreads();
    for (int j = ty; j < RED_CHUNK && c0 + j < NV; j += R) {
      float acc = 0.f;
      for (int t = 0; t < R; ++t) acc += sm[(t * bx + tx) * RED_CHUNK + j];
      emit(c0 + j, acc);
    }
  }
}

// Statistics are accumulated as SHIFTED sums around a per-(n, c) pivot p = x[n, voxel 0, c]:
//   partial[((n*(nslab+1) + slab)*2 + which)*C + c] : which 0 = sum (x - p), 1 = sum (x - p)^2 ; row slab = nslab holds p.
// E[x^2] - mean^2 on raw sums cancels catastrophically in fp32 when |mean| >> std (un-normalised CT intensities reach the
// first GroupNorm of the 'gcr' order, components.py:45-57); with the pivot the sums stay at the scale of the spread and
// the result matches torch's Welford-based GroupNorm.
template <typename T, int V>
__global__ void gn_partial_kernel(const T* __restrict__ x, float* __restrict__ partial, int64_t S, int C,
                                  int ncol, int64_t rows_per_slab, int nslab) {
  extern __shared__ float sm[];
  const int col = blockIdx.z * blockDim.x + threadIdx.x;
  const int n = blockIdx.y, slab = blockIdx.x;
  const int64_t r0 = (int64_t)slab * rows_per_slab;
  int64_t r1 = r0 + rows_per_slab;
  if (r1 > S) r1 = S;
  float acc[2 * V];
#pragma unroll
  for (int i = 0; i < 2 * V; ++i) acc[i] = 0.f;
  const bool active = col < ncol;
  float* out = partial + ((int64_t)n * (nslab + 1) + slab) * 2 * C;
  if (active) {
    const T* base = x + (int64_t)n * S * C + (int64_t)col * V;
    float pv[V];
    load_vec<T, V>(base, pv);
    if (slab == 0 && threadIdx.y == 0) {
      float* prow = partial + ((int64_t)n * (nslab + 1) + nslab) * 2 * C + col * V;
#pragma unroll
      for (int i = 0; i < V; ++i) prow[i] = pv[i];
    }
    const int64_t R = blockDim.y;
    for (int64_t r = r0 + threadIdx.y; r < r1; r += USTD * R) {      // USTD independent 16-byte loads in flight per thread
      typename RawVec<sizeof(T) * V>::type raw[USTD];
#pragma unroll
      for (int u = 0; u < USTD; ++u)
        if (r + u * R < r1) raw[u] = load_raw<T, V>(base + (r + u * R) * C);
#pragma unroll
      for (int u = 0; u < USTD; ++u)
        if (r + u * R < r1) {
          float v[V];
          cvt_raw<T, V>(raw[u], v);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float d = v[i] - pv[i];
            acc[i] += d;
            acc[V + i] += d * d;
          }
        }
    }
  }
  reduce_over_y<2 * V>(acc, sm, [&](int i, float total) {
    if (active) out[(i / V) * C + col * V + (i % V)] = total;
  });
}

// Group statistics from per-channel shifted sums (fixed order, double): channel c with pivot p_c, weight w (voxels each
// partial element stands for), s_c = sum (x - p_c), q_c = sum (x - p_c)^2 over its S values:
//   mean_c = p_c + s_c / S,  M2_c = q_c - s_c^2 / S;   group: mu = avg_c mean_c,  var = avg_c [M2_c / S + (mean_c - mu)^2]
// One warp per channel sums the slabs; sh[0..cpg) = mean_c, sh[cpg..2cpg) = M2_c / S.  Returns (mu, var) in thread 0.
struct ChanSrc {
  const float* partial;   // [n][nslab + 1][2][C_src]
  int nslab, C_src, c0;   // group channels [c0_group, ...) map to source channels c - c0
  double weight;
};
__device__ __forceinline__ void channel_moments(const ChanSrc& src, int n, int c_src, double S, int lane, double& mean_c,
                                                double& var_c) {
  const float* base = src.partial + (int64_t)n * (src.nslab + 1) * 2 * src.C_src;
  double s = 0.0, q = 0.0;
  for (int slab = lane; slab < src.nslab; slab += 32) {
    s += (double)base[(int64_t)slab * 2 * src.C_src + c_src];
    q += (double)base[(int64_t)slab * 2 * src.C_src + src.C_src + c_src];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  s *= src.weight;
  q *= src.weight;
  const double pivot = (double)base[(int64_t)src.nslab * 2 * src.C_src + c_src];
  mean_c = pivot + s / S;
  var_c = q / S - (s / S) * (s / S);
  if (var_c < 0.0) var_c = 0.0;
}
__device__ __forceinline__ void group_stats_finish(double* sh, int cpg, double* scratch, double& mu, double& var) {
  __syncthreads();
  double a = 0.0;
  for (int i = threadIdx.x; i < cpg; i += blockDim.x) a += sh[i];
  a = block_sum(a, scratch);
  __shared__ double s_mu;
  if (threadIdx.x == 0) s_mu = a / (double)cpg;
  __syncthreads();
  mu = s_mu;
  double b = 0.0;
  for (int i = threadIdx.x; i < cpg; i += blockDim.x) b += sh[cpg + i] + (sh[i] - mu) * (sh[i] - mu);
  b = block_sum(b, scratch);
  var = b / (double)cpg;
}

// one block per (n, g): mean / rstd and the per-(n,c) affine table ab[n][0][c] = a, ab[n][1][c] = b
__global__ void gn_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ ab, int64_t S, int C, int G,
                                   int nslab, float eps) {
  extern __shared__ double sh_gn[];                  // [2 * cpg]
  __shared__ double scratch[32];
  __shared__ float s_mean, s_rstd;
  const int n = blockIdx.x / G, g = blockIdx.x % G;
  const int cpg = C / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const ChanSrc src{partial, nslab, C, 0, 1.0};
  for (int i = warp; i < cpg; i += nwarps) {
    double m, v;
    channel_moments(src, n, g * cpg + i, (double)S, lane, m, v);
    if (lane == 0) { sh_gn[i] = m; sh_gn[cpg + i] = v; }
  }
  double mu, var;
  group_stats_finish(sh_gn, cpg, scratch, mu, var);
  if (threadIdx.x == 0) {
    s_mean = (float)mu;
    s_rstd = (float)(1.0 / sqrt(var + (double)eps));
    mean[n * G + g] = s_mean;
    rstd[n * G + g] = s_rstd;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cpg; i += blockDim.x) {
    const int c = g * cpg + i;
    const float a = gamma[c] * s_rstd;
    ab[((int64_t)n * 2 + 0) * C + c] = a;
    ab[((int64_t)n * 2 + 1) * C + c] = beta[c] - s_mean * a;
  }
}

template <typename T, int V>
__global__ void gn_apply_kernel(const T* __restrict__ x, const T* __restrict__ residual, T* __restrict__ y,
                                const float* __restrict__ ab, int64_t S, int C, int ncol,
                                int64_t rows_per_slab, int act, float act_param) {
  const int col = blockIdx.z * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const int n = blockIdx.y, slab = blockIdx.x;
  const int64_t r0 = (int64_t)slab * rows_per_slab;
  int64_t r1 = r0 + rows_per_slab;
  if (r1 > S) r1 = S;
  float a[V], b[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    a[i] = ab[((int64_t)n * 2 + 0) * C + col * V + i];
    b[i] = ab[((int64_t)n * 2 + 1) * C + col * V + i];
  }
  const int64_t base = (int64_t)n * S * C + (int64_t)col * V;
  const int64_t R = blockDim.y;
  for (int64_t r = r0 + threadIdx.y; r < r1; r += UAPP * R) {
    typename RawVec<sizeof(T) * V>::type rx[UAPP], rres[UAPP];
#pragma unroll
    for (int u = 0; u < UAPP; ++u)
      if (r + u * R < r1) {
        rx[u] = load_raw<T, V>(x + base + (r + u * R) * C);
        if (residual != nullptr) rres[u] = load_raw<T, V>(residual + base + (r + u * R) * C);
      }
#pragma unroll
    for (int u = 0; u < UAPP; ++u)
      if (r + u * R < r1) {
        float v[V];
        cvt_raw<T, V>(rx[u], v);
        if (residual != nullptr) {
          float rr[V];
          cvt_raw<T, V>(rres[u], rr);
#pragma unroll
          for (int i = 0; i < V; ++i) v[i] = act_apply(fmaf(v[i], a[i], b[i]) + rr[i], act, act_param);
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) v[i] = act_apply(fmaf(v[i], a[i], b[i]), act, act_param);
        }
        store_vec<T, V>(y + base + (r + u * R) * C, v);
      }
  }
}

// backward stage 1: per (n, slab, c): S1 = sum dyh, S2x = sum dyh * x, dyh = dy * act'(y)
template <typename T, int V>
__global__ void gn_bwd_partial_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                      const T* __restrict__ dy, float* __restrict__ partial, int64_t S, int C,
                                      int ncol, int64_t rows_per_slab, int nslab, int act, float act_param) {
  extern __shared__ float sm[];
  const int col = blockIdx.z * blockDim.x + threadIdx.x;
  const int n = blockIdx.y, slab = blockIdx.x;
  const int64_t r0 = (int64_t)slab * rows_per_slab;
  int64_t r1 = r0 + rows_per_slab;
  if (r1 > S) r1 = S;
  float acc[2 * V];
#pragma unroll
  for (int i = 0; i < 2 * V; ++i) acc[i] = 0.f;
  const bool active = col < ncol;
  if (active) {
    const int64_t base = (int64_t)n * S * C + (int64_t)col * V;
    const int64_t R = blockDim.y;
    constexpr int U2 = USTD / 2;
    for (int64_t r = r0 + threadIdx.y; r < r1; r += U2 * R) {
      typename RawVec<sizeof(T) * V>::type rx[U2], rg[U2], ry[U2];
#pragma unroll
      for (int u = 0; u < U2; ++u)
        if (r + u * R < r1) {
          rx[u] = load_raw<T, V>(x + base + (r + u * R) * C);
          rg[u] = load_raw<T, V>(dy + base + (r + u * R) * C);
          if (act != MEDNET_ACT_NONE) ry[u] = load_raw<T, V>(y + base + (r + u * R) * C);
        }
#pragma unroll
      for (int u = 0; u < U2; ++u)
        if (r + u * R < r1) {
          float xv[V], gv[V];
          cvt_raw<T, V>(rx[u], xv);
          cvt_raw<T, V>(rg[u], gv);
          if (act != MEDNET_ACT_NONE) {
            float yv[V];
            cvt_raw<T, V>(ry[u], yv);
#pragma unroll
            for (int i = 0; i < V; ++i) gv[i] *= act_grad_from_out(yv[i], act, act_param);
          }
#pragma unroll
          for (int i = 0; i < V; ++i) {
            acc[i] += gv[i];
            acc[V + i] += gv[i] * xv[i];
          }
        }
    }
  }
  float* out = partial + ((int64_t)n * nslab + slab) * 2 * C;
  reduce_over_y<2 * V>(acc, sm, [&](int i, float total) {
    if (active) out[(i / V) * C + col * V + (i % V)] = total;
  });
}

// backward stage 2, one block per (n,g): coefficient table coef[n][{A,B,Cc}][c] with
//   dx = A*dyh + B*x + Cc,   A = rstd*gamma, B = -rstd^2*DS/m, Cc = rstd^2*DS*mu/m - rstd*DB/m
// and the per-sample parameter-gradient contributions dgb[n][0][c] = sum dyh*xhat, dgb[n][1][c] = sum dyh.
__global__ void gn_bwd_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ gamma,
                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                       float* __restrict__ coef, float* __restrict__ dgb, int64_t S, int C,
                                       int G, int nslab) {
  __shared__ double scratch[32];
  __shared__ double s_ds, s_db;
  const int n = blockIdx.x / G, g = blockIdx.x % G;
  const int cpg = C / G;
  const double mu = (double)mean[n * G + g], rs = (double)rstd[n * G + g];
  double ds = 0.0, db = 0.0;
  // one warp per channel of the group, lanes stride over the slabs (fixed order: deterministic)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = warp; i < cpg; i += nw) {
    const int c = g * cpg + i;
    double s1 = 0.0, s2x = 0.0;
    for (int slab = lane; slab < nslab; slab += 32) {
      const float* p = partial + ((int64_t)n * nslab + slab) * 2 * C;
      s1 += (double)p[c];
      s2x += (double)p[C + c];
    }
    s1 = warp_sum(s1);
    s2x = warp_sum(s2x);
    if (lane == 0) {
      const double s2 = rs * (s2x - mu * s1);
      dgb[((int64_t)n * 2 + 0) * C + c] = (float)s2;
      dgb[((int64_t)n * 2 + 1) * C + c] = (float)s1;
      ds += (double)gamma[c] * s2;
      db += (double)gamma[c] * s1;
    }
  }
  ds = block_sum(ds, scratch);
  db = block_sum(db, scratch);
  if (threadIdx.x == 0) {
    s_ds = ds;
    s_db = db;
  }
  __syncthreads();
  const double m = (double)cpg * (double)S;
  const double B = -rs * rs * s_ds / m;
  const double Cc = rs * rs * s_ds * mu / m - rs * s_db / m;
  for (int i = threadIdx.x; i < cpg; i += blockDim.x) {
    const int c = g * cpg + i;
    coef[((int64_t)n * 3 + 0) * C + c] = (float)(rs * (double)gamma[c]);
    coef[((int64_t)n * 3 + 1) * C + c] = (float)B;
    coef[((int64_t)n * 3 + 2) * C + c] = (float)Cc;
  }
}

__global__ void gn_bwd_param_kernel(const float* __restrict__ dgb, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int N, int C, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float dg = 0.f, dbv = 0.f;
  for (int n = 0; n < N; ++n) {
    dg += dgb[((int64_t)n * 2 + 0) * C + c];
    dbv += dgb[((int64_t)n * 2 + 1) * C + c];
  }
  if (accumulate) {
    dgamma[c] += dg;
    dbeta[c] += dbv;
  } else {
    dgamma[c] = dg;
    dbeta[c] = dbv;
  }
}

// ACT = false: no activation after the norm and no residual branch (every GroupNorm of the 'gcr' blocks): x and dy are the
// only streams, four rows of each in flight per thread (the ACT = true body keeps two: it also holds y)
template <typename T, int V, bool ACT>
__global__ void __launch_bounds__(256, 2) gn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ dy,
                                    const float* __restrict__ coef, T* __restrict__ dx,
                                    T* __restrict__ dresidual, int64_t S, int C, int ncol,
                                    int64_t rows_per_slab, int act, float act_param, int in_act, float in_act_param) {
  const int col = blockIdx.z * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const int n = blockIdx.y, slab = blockIdx.x;
  const int64_t r0 = (int64_t)slab * rows_per_slab;
  int64_t r1 = r0 + rows_per_slab;
  if (r1 > S) r1 = S;
  float A[V], B[V], Cc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    A[i] = coef[((int64_t)n * 3 + 0) * C + col * V + i];
    B[i] = coef[((int64_t)n * 3 + 1) * C + col * V + i];
    Cc[i] = coef[((int64_t)n * 3 + 2) * C + col * V + i];
  }
  const int64_t base = (int64_t)n * S * C + (int64_t)col * V;
  const int64_t R = blockDim.y;
  constexpr int U2 = ACT ? USTD / 2 : USTD;
  for (int64_t r = r0 + threadIdx.y; r < r1; r += U2 * R) {
    typename RawVec<sizeof(T) * V>::type rx[U2], rg[U2], ry[U2];
#pragma unroll
    for (int u = 0; u < U2; ++u)
      if (r + u * R < r1) {
        rx[u] = load_raw<T, V>(x + base + (r + u * R) * C);
        rg[u] = load_raw<T, V>(dy + base + (r + u * R) * C);
        if (ACT && act != MEDNET_ACT_NONE) ry[u] = load_raw<T, V>(y + base + (r + u * R) * C);
      }
#pragma unroll
    for (int u = 0; u < U2; ++u)
      if (r + u * R < r1) {
        float xv[V], gv[V];
        cvt_raw<T, V>(rx[u], xv);
        cvt_raw<T, V>(rg[u], gv);
        if (ACT && act != MEDNET_ACT_NONE) {
          float yv[V];
          cvt_raw<T, V>(ry[u], yv);
#pragma unroll
          for (int i = 0; i < V; ++i) gv[i] *= act_grad_from_out(yv[i], act, act_param);
        }
        if (ACT && dresidual != nullptr) store_vec<T, V>(dresidual + base + (r + u * R) * C, gv);
        if (in_act != MEDNET_ACT_NONE) {            // deferred derivative of the activation that produced x
#pragma unroll
          for (int i = 0; i < V; ++i)
            xv[i] = fmaf(A[i], gv[i], fmaf(B[i], xv[i], Cc[i])) * act_grad_from_out(xv[i], in_act, in_act_param);
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) xv[i] = fmaf(A[i], gv[i], fmaf(B[i], xv[i], Cc[i]));
        }
        store_vec<T, V>(dx + base + (r + u * R) * C, xv);
      }
  }
}

template <typename T, int V>
__global__ void act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t nvec, int act, float a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float v[V];
    load_vec<T, V>(x + i * V, v);
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = act_apply(v[k], act, a);
    store_vec<T, V>(y + i * V, v);
  }
}

template <typename T, int V>
__global__ void act_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx,
                               int64_t nvec, int act, float a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float yv[V], gv[V];
    load_vec<T, V>(y + i * V, yv);
    load_vec<T, V>(dy + i * V, gv);
#pragma unroll
    for (int k = 0; k < V; ++k) gv[k] *= act_grad_from_out(yv[k], act, a);
    store_vec<T, V>(dx + i * V, gv);
  }
}


// ------------------------------------------------------------------------------------------------
// GroupNorm over the VIRTUAL concat cat((skip, nearest_up2(low)), channel): the concat tensor is never
// materialised (mednet_upcat_groupnorm_fwd / _bwd).  Exact 2x upsampling: every low voxel has 8 children.
// ------------------------------------------------------------------------------------------------
// one block per (n, g): statistics from the per-channel partial sums of skip (pa) and low (pb, weight 8)
__global__ void upcat_gn_finalize_kernel(const float* __restrict__ pa, const float* __restrict__ pb,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ ab,
                                         int64_t S, int Cs, int Cl, int G, int nslab_a, int nslab_b, float eps) {
  extern __shared__ double sh_gn[];                  // [2 * cpg]
  __shared__ double scratch[32];
  __shared__ float s_mean, s_rstd;
  const int n = blockIdx.x / G, g = blockIdx.x % G;
  const int C = Cs + Cl, cpg = C / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const ChanSrc sa{pa, nslab_a, Cs, 0, 1.0};
  const ChanSrc sb{pb, nslab_b, Cl, Cs, 8.0};       // every low-resolution voxel appears 8 times in the upsampled concat
  for (int i = warp; i < cpg; i += nwarps) {
    const int c = g * cpg + i;
    double m, v;
    if (c < Cs) channel_moments(sa, n, c, (double)S, lane, m, v);
    else channel_moments(sb, n, c - Cs, (double)S, lane, m, v);
    if (lane == 0) { sh_gn[i] = m; sh_gn[cpg + i] = v; }
  }
  double mu, var;
  group_stats_finish(sh_gn, cpg, scratch, mu, var);
  if (threadIdx.x == 0) {
    s_mean = (float)mu;
    s_rstd = (float)(1.0 / sqrt(var + (double)eps));
    mean[n * G + g] = s_mean;
    rstd[n * G + g] = s_rstd;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cpg; i += blockDim.x) {
    const int c = g * cpg + i;
    const float a = gamma[c] * s_rstd;
    ab[((int64_t)n * 2 + 0) * C + c] = a;
    ab[((int64_t)n * 2 + 1) * C + c] = beta[c] - s_mean * a;
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256, 4) upcat_gn_apply_kernel(const T* __restrict__ skip, const T* __restrict__ low, T* __restrict__ y,
                                      const float* __restrict__ ab, int D, int H, int W, int Cs, int Cl, int ncol,
                                      int lines_per_slab) {
  // a "line" is one (z, y) row of W voxels; thread (tx, ty) owns channel vector tx and voxels x = ty, ty + R, ...
  const int col = blockIdx.z * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const int n = blockIdx.y, slab = blockIdx.x;
  const int C = Cs + Cl, ncs = Cs / V, nlines = D * H;
  const int h = H >> 1, w = W >> 1;
  const int l0 = slab * lines_per_slab;
  const int l1 = min(l0 + lines_per_slab, nlines);
  float a[V], b[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    a[i] = ab[((int64_t)n * 2 + 0) * C + col * V + i];
    b[i] = ab[((int64_t)n * 2 + 1) * C + col * V + i];
  }
  const bool from_skip = col < ncs;
  const T* sbase = skip + (int64_t)n * nlines * W * Cs + col * V;
  const T* lbase = low + (int64_t)n * (nlines >> 2) * w * Cl + (col - ncs) * V;
  T* ybase = y + (int64_t)n * nlines * W * C + col * V;
  for (int l = l0; l < l1; ++l) {
    const int z = l / H, yy = l - z * H;
    const int64_t r0 = (int64_t)l * W;
    const int64_t rl0 = (int64_t)((z >> 1) * h + (yy >> 1)) * w;
    const int R = blockDim.y;
    for (int x0 = threadIdx.y; x0 < W; x0 += UNR * R) {
      typename RawVec<sizeof(T) * V>::type raw[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int x = x0 + u * R;
        if (x < W) raw[u] = from_skip ? load_raw<T, V>(sbase + (r0 + x) * Cs) : load_raw<T, V>(lbase + (rl0 + (x >> 1)) * Cl);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int x = x0 + u * R;
        if (x < W) {
          float v[V];
          cvt_raw<T, V>(raw[u], v);
#pragma unroll
          for (int i = 0; i < V; ++i) v[i] = fmaf(v[i], a[i], b[i]);
          store_vec<T, V>(ybase + (r0 + x) * C, v);
        }
      }
    }
  }
}

// backward stage 1 on the virtual concat: partial[n][slab][{sum dy, sum dy*x}][C]
template <typename T, int V>
__global__ void __launch_bounds__(256, 4) upcat_gn_bwd_partial_kernel(const T* __restrict__ skip, const T* __restrict__ low,
                                            const T* __restrict__ dy, float* __restrict__ partial, int D, int H, int W,
                                            int Cs, int Cl, int ncol, int lines_per_slab, int nslab) {
  extern __shared__ float sm[];
  const int col = blockIdx.z * blockDim.x + threadIdx.x;
  const int n = blockIdx.y, slab = blockIdx.x;
  const int C = Cs + Cl, ncs = Cs / V, nlines = D * H;
  const int h = H >> 1, w = W >> 1;
  const int l0 = slab * lines_per_slab;
  const int l1 = min(l0 + lines_per_slab, nlines);
  float acc[2 * V];
#pragma unroll
  for (int i = 0; i < 2 * V; ++i) acc[i] = 0.f;
  const bool active = col < ncol;
  if (active) {
    const bool from_skip = col < ncs;
    const T* sbase = skip + (int64_t)n * nlines * W * Cs + col * V;
    const T* lbase = low + (int64_t)n * (nlines >> 2) * w * Cl + (col - ncs) * V;
    const T* gbase = dy + (int64_t)n * nlines * W * C + col * V;
    for (int l = l0; l < l1; ++l) {
      const int z = l / H, yy = l - z * H;
      const int64_t r0 = (int64_t)l * W;
      const int64_t rl0 = (int64_t)((z >> 1) * h + (yy >> 1)) * w;
      const int R = blockDim.y;
      for (int x0 = threadIdx.y; x0 < W; x0 += UNR * R) {
        typename RawVec<sizeof(T) * V>::type rx[UNR], rg[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int x = x0 + u * R;
          if (x < W) {
            rx[u] = from_skip ? load_raw<T, V>(sbase + (r0 + x) * Cs) : load_raw<T, V>(lbase + (rl0 + (x >> 1)) * Cl);
            rg[u] = load_raw<T, V>(gbase + (r0 + x) * C);
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          if (x0 + u * R < W) {
            float xv[V], gv[V];
            cvt_raw<T, V>(rx[u], xv);
            cvt_raw<T, V>(rg[u], gv);
#pragma unroll
            for (int i = 0; i < V; ++i) {
              acc[i] += gv[i];
              acc[V + i] += gv[i] * xv[i];
            }
          }
        }
      }
    }
  }
  float* out = partial + ((int64_t)n * nslab + slab) * 2 * C;
  reduce_over_y<2 * V>(acc, sm, [&](int i, float total) {
    if (active) out[(i / V) * C + col * V + (i % V)] = total;
  });
}

// backward stage 3a: dskip = (A*dy + B*skip + Cc) * skip_act'(skip) for the first Cs concat channels
template <typename T, int V>
__global__ void __launch_bounds__(256, 3) upcat_gn_bwd_skip_kernel(const T* __restrict__ skip, const T* __restrict__ dy,
                                         const float* __restrict__ coef, T* __restrict__ dskip, int64_t S, int Cs, int C,
                                         int ncol, int64_t rows_per_slab, int act, float act_param) {
  const int col = blockIdx.z * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const int n = blockIdx.y, slab = blockIdx.x;
  const int64_t r0 = (int64_t)slab * rows_per_slab;
  int64_t r1 = r0 + rows_per_slab;
  if (r1 > S) r1 = S;
  float A[V], B[V], Cc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    A[i] = coef[((int64_t)n * 3 + 0) * C + col * V + i];
    B[i] = coef[((int64_t)n * 3 + 1) * C + col * V + i];
    Cc[i] = coef[((int64_t)n * 3 + 2) * C + col * V + i];
  }
  const T* xb = skip + (int64_t)n * S * Cs + col * V;
  const T* gb = dy + (int64_t)n * S * C + col * V;
  T* ob = dskip + (int64_t)n * S * Cs + col * V;
  const int64_t R = blockDim.y;
  for (int64_t r = r0 + threadIdx.y; r < r1; r += UNR * R) {
    typename RawVec<sizeof(T) * V>::type rx[UNR], rg[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (r + u * R < r1) {
        rx[u] = load_raw<T, V>(xb + (r + u * R) * Cs);
        rg[u] = load_raw<T, V>(gb + (r + u * R) * C);
      }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (r + u * R < r1) {
        float xv[V], gv[V];
        cvt_raw<T, V>(rx[u], xv);
        cvt_raw<T, V>(rg[u], gv);
#pragma unroll
        for (int i = 0; i < V; ++i)
          xv[i] = fmaf(A[i], gv[i], fmaf(B[i], xv[i], Cc[i])) * act_grad_from_out(xv[i], act, act_param);
        store_vec<T, V>(ob + (r + u * R) * Cs, xv);
      }
  }
}

// backward stage 3b: dlow = (A * sum_8 dy + 8 * (B*low + Cc)) * low_act'(low) for the last Cl concat channels
template <typename T, int V>
__global__ void upcat_gn_bwd_low_kernel(const T* __restrict__ low, const T* __restrict__ dy,
                                        const float* __restrict__ coef, T* __restrict__ dlow, int N, int D, int H, int W,
                                        int Cs, int Cl, int act, float act_param) {
  const int C = Cs + Cl, ncl = Cl / V;
  const int d = D >> 1, h = H >> 1, w = W >> 1;
  const int64_t total = (int64_t)N * d * h * w * ncl;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % ncl);
    int64_t t = i / ncl;
    const int x = (int)(t % w); t /= w;
    const int yy = (int)(t % h); t /= h;
    const int z = (int)(t % d);
    const int n = (int)(t / d);
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t row = (((int64_t)n * D + 2 * z + (k >> 2)) * H + 2 * yy + ((k >> 1) & 1)) * W + 2 * x + (k & 1);
      float g[V];
      load_vec<T, V>(dy + row * C + Cs + cv * V, g);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += g[j];
    }
    float xv[V];
    load_vec<T, V>(low + i * V, xv);
    const float* cf = coef + (int64_t)n * 3 * C + Cs + cv * V;
#pragma unroll
    for (int j = 0; j < V; ++j)
      xv[j] = fmaf(cf[j], acc[j], 8.f * fmaf(cf[C + j], xv[j], cf[2 * C + j])) * act_grad_from_out(xv[j], act, act_param);
    store_vec<T, V>(dlow + i * V, xv);
  }
}


// ------------------------------------------------------------------------------------------------
// GroupNorm over the virtual concat, SPLIT outputs (decoder join feeding the upsample-aware convolution):
// the normalised skip part stays at full resolution, the normalised low part stays on the COARSE grid
// (a per-channel affine commutes with nearest upsampling).  Forward reuses gn_partial / upcat_gn_finalize / gn_apply
// with per-part affine tables; backward reuses gn_bwd_partial / gn_bwd_apply with per-part coefficient tables.
// ------------------------------------------------------------------------------------------------
// ab[n][2][Cs + Cl] -> ab_a[n][2][Cs], ab_b[n][2][Cl]
__global__ void split_ab_kernel(const float* __restrict__ ab, float* __restrict__ ab_a, float* __restrict__ ab_b, int N, int Cs,
                                int Cl) {
  const int C = Cs + Cl;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * 2 * C) return;
  const int c = i % C, w = (i / C) % 2, n = i / (2 * C);
  if (c < Cs) ab_a[((int64_t)n * 2 + w) * Cs + c] = ab[i];
  else ab_b[((int64_t)n * 2 + w) * Cl + c - Cs] = ab[i];
}

// one block per (n, g) of the virtual concat.  pa / pb: [n][nslab][2][Cs | Cl] sums of dy and dy*x over the skip part
// (full resolution) and over the low part (coarse grid, dy already summed over the 8 children of each coarse voxel).
//   skip:  dx = A dy + B x + Cc            low:  dx = A dy + 8 (B x + Cc)      (8 children share x)
__global__ void split_gn_bwd_finalize_kernel(const float* __restrict__ pa, const float* __restrict__ pb,
                                             const float* __restrict__ gamma, const float* __restrict__ mean,
                                             const float* __restrict__ rstd, float* __restrict__ coef_a,
                                             float* __restrict__ coef_b, float* __restrict__ dgb, int64_t S, int Cs, int Cl,
                                             int G, int nslab_a, int nslab_b) {
  __shared__ double scratch[32];
  __shared__ double s_ds, s_db;
  const int n = blockIdx.x / G, g = blockIdx.x % G;
  const int C = Cs + Cl, cpg = C / G;
  const double mu = (double)mean[n * G + g], rs = (double)rstd[n * G + g];
  double ds = 0.0, db = 0.0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = warp; i < cpg; i += nw) {
    const int c = g * cpg + i;
    const bool lo = c >= Cs;
    const float* base = lo ? pb + (int64_t)n * nslab_b * 2 * Cl : pa + (int64_t)n * nslab_a * 2 * Cs;
    const int Cp = lo ? Cl : Cs, cp = lo ? c - Cs : c, ns = lo ? nslab_b : nslab_a;
    double s1 = 0.0, s2x = 0.0;
    for (int slab = lane; slab < ns; slab += 32) {
      s1 += (double)base[(int64_t)slab * 2 * Cp + cp];
      s2x += (double)base[(int64_t)slab * 2 * Cp + Cp + cp];
    }
    s1 = warp_sum(s1);
    s2x = warp_sum(s2x);
    if (lane == 0) {
      const double s2 = rs * (s2x - mu * s1);
      dgb[((int64_t)n * 2 + 0) * C + c] = (float)s2;
      dgb[((int64_t)n * 2 + 1) * C + c] = (float)s1;
      ds += (double)gamma[c] * s2;
      db += (double)gamma[c] * s1;
    }
  }
  ds = block_sum(ds, scratch);
  db = block_sum(db, scratch);
  if (threadIdx.x == 0) {
    s_ds = ds;
    s_db = db;
  }
  __syncthreads();
  const double m = (double)cpg * (double)S;
  const double B = -rs * rs * s_ds / m;
  const double Cc = rs * rs * s_ds * mu / m - rs * s_db / m;
  for (int i = threadIdx.x; i < cpg; i += blockDim.x) {
    const int c = g * cpg + i;
    const float A = (float)(rs * (double)gamma[c]);
    if (c < Cs) {
      coef_a[((int64_t)n * 3 + 0) * Cs + c] = A;
      coef_a[((int64_t)n * 3 + 1) * Cs + c] = (float)B;
      coef_a[((int64_t)n * 3 + 2) * Cs + c] = (float)Cc;
    } else {
      coef_b[((int64_t)n * 3 + 0) * Cl + c - Cs] = A;
      coef_b[((int64_t)n * 3 + 1) * Cl + c - Cs] = (float)(8.0 * B);
      coef_b[((int64_t)n * 3 + 2) * Cl + c - Cs] = (float)(8.0 * Cc);
    }
  }
}

}  // namespace mednet

using namespace mednet;

extern "C" size_t mednet_groupnorm_fwd_workspace_bytes(const mednet_gn_fwd_params* p) {
  if (!p || !dtype_ok(p->dtype) || p->C <= 0) return 0;
  SlabPlan pl = make_plan(p->N, p->S, p->C, dtype_bytes(p->dtype));
  return align_up((size_t)p->N * (pl.nslab + 1) * 2 * p->C * sizeof(float), 256) +
         align_up((size_t)p->N * 2 * p->C * sizeof(float), 256);
}

extern "C" int mednet_groupnorm_fwd(const mednet_gn_fwd_params* p, void* workspace, size_t workspace_bytes,
                                    mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->x && p->y && p->gamma && p->beta && p->mean && p->rstd, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(p->N > 0 && p->S > 0 && p->C > 0 && p->G > 0 && p->C % p->G == 0, MEDNET_EINVAL);
  MEDNET_REQUIRE(p->N <= 65535, MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_groupnorm_fwd_workspace_bytes(p), MEDNET_EWORKSPACE);
  SlabPlan pl = make_plan(p->N, p->S, p->C, dtype_bytes(p->dtype));
  float* partial = (float*)workspace;
  float* ab = (float*)((char*)workspace + align_up((size_t)p->N * (pl.nslab + 1) * 2 * p->C * sizeof(float), 256));
  dim3 grid(pl.nslab, (unsigned)p->N, pl.coltiles), block(pl.ncol_t, pl.R);
  MEDNET_DISPATCH_TV(p->dtype, pl.V, {
    size_t smem = (size_t)pl.ncol_t * pl.R * RED_CHUNK * sizeof(float);
    gn_partial_kernel<T, VV><<<grid, block, smem, stream>>>((const T*)p->x, partial, p->S, p->C, pl.ncol,
                                                            pl.rows_per_slab, pl.nslab);
  });
  MEDNET_LAUNCH_CHECK();
  gn_finalize_kernel<<<(unsigned)(p->N * p->G), 128, 2 * (p->C / p->G) * sizeof(double), stream>>>(
      partial, p->gamma, p->beta, p->mean, p->rstd, ab, p->S, p->C, p->G, pl.nslab, p->eps);
  MEDNET_LAUNCH_CHECK();
  MEDNET_DISPATCH_TV(p->dtype, pl.V, {
    gn_apply_kernel<T, VV><<<grid, block, 0, stream>>>((const T*)p->x, (const T*)p->residual, (T*)p->y, ab, p->S,
                                                       p->C, pl.ncol, pl.rows_per_slab, p->act, p->act_param);
  });
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" size_t mednet_groupnorm_bwd_workspace_bytes(const mednet_gn_bwd_params* p) {
  if (!p || !dtype_ok(p->dtype) || p->C <= 0) return 0;
  SlabPlan pl = make_plan(p->N, p->S, p->C, dtype_bytes(p->dtype));
  return align_up((size_t)p->N * pl.nslab * 2 * p->C * sizeof(float), 256) +
         align_up((size_t)p->N * 3 * p->C * sizeof(float), 256) + align_up((size_t)p->N * 2 * p->C * sizeof(float), 256);
}

extern "C" int mednet_groupnorm_bwd(const mednet_gn_bwd_params* p, void* workspace, size_t workspace_bytes,
                                    mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->x && p->dy && p->gamma && p->mean && p->rstd && p->dx && p->dgamma && p->dbeta,
                 MEDNET_EINVAL);
  MEDNET_REQUIRE(p->act == MEDNET_ACT_NONE || p->y != nullptr, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(p->N > 0 && p->S > 0 && p->C > 0 && p->G > 0 && p->C % p->G == 0, MEDNET_EINVAL);
  MEDNET_REQUIRE(p->N <= 65535, MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_groupnorm_bwd_workspace_bytes(p), MEDNET_EWORKSPACE);
  SlabPlan pl = make_plan(p->N, p->S, p->C, dtype_bytes(p->dtype));
  char* ws = (char*)workspace;
  float* partial = (float*)ws;
  ws += align_up((size_t)p->N * pl.nslab * 2 * p->C * sizeof(float), 256);
  float* coef = (float*)ws;
  ws += align_up((size_t)p->N * 3 * p->C * sizeof(float), 256);
  float* dgb = (float*)ws;
  dim3 grid(pl.nslab, (unsigned)p->N, pl.coltiles), block(pl.ncol_t, pl.R);
  MEDNET_DISPATCH_TV(p->dtype, pl.V, {
    size_t smem = (size_t)pl.ncol_t * pl.R * RED_CHUNK * sizeof(float);
    gn_bwd_partial_kernel<T, VV><<<grid, block, smem, stream>>>((const T*)p->x, (const T*)p->y, (const T*)p->dy,
                                                                partial, p->S, p->C, pl.ncol, pl.rows_per_slab,
                                                                pl.nslab, p->act, p->act_param);
  });
  MEDNET_LAUNCH_CHECK();
  gn_bwd_finalize_kernel<<<(unsigned)(p->N * p->G), 128, 0, stream>>>(partial, p->gamma, p->mean, p->rstd, coef, dgb,
                                                                      p->S, p->C, p->G, pl.nslab);
  MEDNET_LAUNCH_CHECK();
  gn_bwd_param_kernel<<<ceil_div(p->C, 128), 128, 0, stream>>>(dgb, p->dgamma, p->dbeta, (int)p->N, p->C,
                                                               p->accumulate);
  MEDNET_LAUNCH_CHECK();
  if (p->act == MEDNET_ACT_NONE && p->dresidual == nullptr) {
    MEDNET_DISPATCH_TV(p->dtype, pl.V, {
      gn_bwd_apply_kernel<T, VV, false><<<grid, block, 0, stream>>>((const T*)p->x, (const T*)p->y, (const T*)p->dy, coef,
                                                                    (T*)p->dx, (T*)nullptr, p->S, p->C, pl.ncol,
                                                                    pl.rows_per_slab, p->act, p->act_param, p->in_act,
                                                                    p->in_act_param);
    });
  } else {
    MEDNET_DISPATCH_TV(p->dtype, pl.V, {
      gn_bwd_apply_kernel<T, VV, true><<<grid, block, 0, stream>>>((const T*)p->x, (const T*)p->y, (const T*)p->dy, coef,
                                                                   (T*)p->dx, (T*)p->dresidual, p->S, p->C, pl.ncol,
                                                                   pl.rows_per_slab, p->act, p->act_param, p->in_act,
                                                                   p->in_act_param);
    });
  }
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_act_fwd(const mednet_act_fwd_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->x && p->y && p->numel > 0, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  const int V = pick_vec(p->numel, dtype_bytes(p->dtype));
  const int64_t nvec = p->numel / V;
  MEDNET_DISPATCH_TV(p->dtype, V, {
    act_fwd_kernel<T, VV><<<grid_for(nvec, 256), 256, 0, stream>>>((const T*)p->x, (T*)p->y, nvec, p->act, p->act_param);
  });
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_act_bwd(const mednet_act_bwd_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->y && p->dy && p->dx && p->numel > 0, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  const int V = pick_vec(p->numel, dtype_bytes(p->dtype));
  const int64_t nvec = p->numel / V;
  MEDNET_DISPATCH_TV(p->dtype, V, {
    act_bwd_kernel<T, VV><<<grid_for(nvec, 256), 256, 0, stream>>>((const T*)p->y, (const T*)p->dy, (T*)p->dx, nvec,
                                                                   p->act, p->act_param);
  });
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm over the virtual concat
// ------------------------------------------------------------------------------------------------
namespace {
struct UpcatPlan {
  int V;
  SlabPlan pa, pb, pc;       // stats over skip, stats over low, apply/backward over the concat grid
  size_t pa_bytes, pb_bytes;
};
bool upcat_geometry_ok(int N, int D, int H, int W, int d, int h, int w, int Cs, int Cl, int G) {
  return N > 0 && N <= 65535 && Cs > 0 && Cl > 0 && G > 0 && (Cs + Cl) % G == 0 && d > 0 && h > 0 && w > 0 &&
         D == 2 * d && H == 2 * h && W == 2 * w && (int64_t)D * H * W < ((int64_t)1 << 31);
}
UpcatPlan upcat_plan(int N, int D, int H, int W, int Cs, int Cl, int dtype) {
  UpcatPlan u;
  const int eb = dtype_bytes(dtype);
  const int va = pick_vec(Cs, eb), vb = pick_vec(Cl, eb);
  u.V = va < vb ? va : vb;
  const int64_t S = (int64_t)D * H * W;
  u.pa = make_plan(N, S, Cs, eb);
  u.pb = make_plan(N, S / 8, Cl, eb);
  u.pc = make_plan(N, (int64_t)D * H, Cs + Cl, eb, u.V, 12);   // slabs of (z, y) LINES of W voxels (concat-grid kernels)
  u.pa_bytes = align_up((size_t)N * (u.pa.nslab + 1) * 2 * Cs * sizeof(float), 256);
  u.pb_bytes = align_up((size_t)N * (u.pb.nslab + 1) * 2 * Cl * sizeof(float), 256);
  return u;
}
}  // namespace

extern "C" size_t mednet_upcat_groupnorm_fwd_workspace_bytes(const mednet_upcat_gn_fwd_params* p) {
  if (!p || !dtype_ok(p->dtype) || !upcat_geometry_ok(p->N, p->D, p->H, p->W, p->d, p->h, p->w, p->Cs, p->Cl, p->G)) return 0;
  UpcatPlan u = upcat_plan(p->N, p->D, p->H, p->W, p->Cs, p->Cl, p->dtype);
  return u.pa_bytes + u.pb_bytes + align_up((size_t)p->N * 2 * (p->Cs + p->Cl) * sizeof(float), 256);
}

extern "C" int mednet_upcat_groupnorm_fwd(const mednet_upcat_gn_fwd_params* p, void* workspace, size_t workspace_bytes,
                                          mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->skip && p->low && p->gamma && p->beta && p->y && p->mean && p->rstd, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(upcat_geometry_ok(p->N, p->D, p->H, p->W, p->d, p->h, p->w, p->Cs, p->Cl, p->G), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_upcat_groupnorm_fwd_workspace_bytes(p), MEDNET_EWORKSPACE);
  UpcatPlan u = upcat_plan(p->N, p->D, p->H, p->W, p->Cs, p->Cl, p->dtype);
  const int64_t S = (int64_t)p->D * p->H * p->W;
  const int C = p->Cs + p->Cl;
  float* pa = (float*)workspace;
  float* pb = (float*)((char*)workspace + u.pa_bytes);
  float* ab = (float*)((char*)workspace + u.pa_bytes + u.pb_bytes);
  {
    const SlabPlan& pl = u.pa;
    dim3 grid(pl.nslab, (unsigned)p->N, pl.coltiles), block(pl.ncol_t, pl.R);
    MEDNET_DISPATCH_TV(p->dtype, pl.V, {
      size_t smem = (size_t)pl.ncol_t * pl.R * RED_CHUNK * sizeof(float);
      gn_partial_kernel<T, VV><<<grid, block, smem, stream>>>((const T*)p->skip, pa, S, p->Cs, pl.ncol, pl.rows_per_slab,
                                                              pl.nslab);
    });
    MEDNET_LAUNCH_CHECK();
  }
  {
    const SlabPlan& pl = u.pb;
    dim3 grid(pl.nslab, (unsigned)p->N, pl.coltiles), block(pl.ncol_t, pl.R);
    MEDNET_DISPATCH_TV(p->dtype, pl.V, {
      size_t smem = (size_t)pl.ncol_t * pl.R * RED_CHUNK * sizeof(float);
      gn_partial_kernel<T, VV><<<grid, block, smem, stream>>>((const T*)p->low, pb, S / 8, p->Cl, pl.ncol, pl.rows_per_slab,
                                                              pl.nslab);
    });
    MEDNET_LAUNCH_CHECK();
  }
  upcat_gn_finalize_kernel<<<(unsigned)(p->N * p->G), 128, 2 * ((p->Cs + p->Cl) / p->G) * sizeof(double), stream>>>(
      pa, pb, p->gamma, p->beta, p->mean, p->rstd, ab, S, p->Cs, p->Cl, p->G, u.pa.nslab, u.pb.nslab, p->eps);
  MEDNET_LAUNCH_CHECK();
  {
    const SlabPlan& pl = u.pc;
    dim3 grid(pl.nslab, (unsigned)p->N, pl.coltiles), block(pl.ncol_t, pl.R);
    MEDNET_DISPATCH_TV(p->dtype, pl.V, {
      upcat_gn_apply_kernel<T, VV><<<grid, block, 0, stream>>>((const T*)p->skip, (const T*)p->low, (T*)p->y, ab, p->D, p->H,
                                                               p->W, p->Cs, p->Cl, pl.ncol, (int)pl.rows_per_slab);
    });
    MEDNET_LAUNCH_CHECK();
  }
  (void)C;
  return MEDNET_OK;
}

extern "C" size_t mednet_upcat_groupnorm_bwd_workspace_bytes(const mednet_upcat_gn_bwd_params* p) {
  if (!p || !dtype_ok(p->dtype) || !upcat_geometry_ok(p->N, p->D, p->H, p->W, p->d, p->h, p->w, p->Cs, p->Cl, p->G)) return 0;
  UpcatPlan u = upcat_plan(p->N, p->D, p->H, p->W, p->Cs, p->Cl, p->dtype);
  const int C = p->Cs + p->Cl;
  return align_up((size_t)p->N * u.pc.nslab * 2 * C * sizeof(float), 256) + align_up((size_t)p->N * 3 * C * sizeof(float), 256) +
         align_up((size_t)p->N * 2 * C * sizeof(float), 256);
}

extern "C" int mednet_upcat_groupnorm_bwd(const mednet_upcat_gn_bwd_params* p, void* workspace, size_t workspace_bytes,
                                          mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->skip && p->low && p->dy && p->gamma && p->mean && p->rstd && p->dskip && p->dlow && p->dgamma &&
                     p->dbeta, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(upcat_geometry_ok(p->N, p->D, p->H, p->W, p->d, p->h, p->w, p->Cs, p->Cl, p->G), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_upcat_groupnorm_bwd_workspace_bytes(p), MEDNET_EWORKSPACE);
  UpcatPlan u = upcat_plan(p->N, p->D, p->H, p->W, p->Cs, p->Cl, p->dtype);
  const int64_t S = (int64_t)p->D * p->H * p->W;
  const int C = p->Cs + p->Cl;
  char* ws = (char*)workspace;
  float* partial = (float*)ws;
  ws += align_up((size_t)p->N * u.pc.nslab * 2 * C * sizeof(float), 256);
  float* coef = (float*)ws;
  ws += align_up((size_t)p->N * 3 * C * sizeof(float), 256);
  float* dgb = (float*)ws;
  {
    const SlabPlan& pl = u.pc;
    dim3 grid(pl.nslab, (unsigned)p->N, pl.coltiles), block(pl.ncol_t, pl.R);
    MEDNET_DISPATCH_TV(p->dtype, pl.V, {
      size_t smem = (size_t)pl.ncol_t * pl.R * RED_CHUNK * sizeof(float);
      upcat_gn_bwd_partial_kernel<T, VV><<<grid, block, smem, stream>>>((const T*)p->skip, (const T*)p->low, (const T*)p->dy,
                                                                        partial, p->D, p->H, p->W, p->Cs, p->Cl, pl.ncol,
                                                                        (int)pl.rows_per_slab, pl.nslab);
    });
    MEDNET_LAUNCH_CHECK();
  }
  gn_bwd_finalize_kernel<<<(unsigned)(p->N * p->G), 128, 0, stream>>>(partial, p->gamma, p->mean, p->rstd, coef, dgb, S, C,
                                                                      p->G, u.pc.nslab);
  MEDNET_LAUNCH_CHECK();
  gn_bwd_param_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(dgb, p->dgamma, p->dbeta, p->N, C, p->accumulate);
  MEDNET_LAUNCH_CHECK();
  {
    SlabPlan pl = make_plan(p->N, S, p->Cs, dtype_bytes(p->dtype), u.V);
    dim3 grid(pl.nslab, (unsigned)p->N, pl.coltiles), block(pl.ncol_t, pl.R);
    MEDNET_DISPATCH_TV(p->dtype, pl.V, {
      upcat_gn_bwd_skip_kernel<T, VV><<<grid, block, 0, stream>>>((const T*)p->skip, (const T*)p->dy, coef, (T*)p->dskip, S,
                                                                  p->Cs, C, pl.ncol, pl.rows_per_slab, p->skip_act,
                                                                  p->skip_act_param);
    });
    MEDNET_LAUNCH_CHECK();
  }
  {
    const int64_t total = (int64_t)p->N * (S / 8) * (p->Cl / u.V);
    MEDNET_DISPATCH_TV(p->dtype, u.V, {
      upcat_gn_bwd_low_kernel<T, VV><<<grid_for(total, 256), 256, 0, stream>>>((const T*)p->low, (const T*)p->dy, coef,
                                                                               (T*)p->dlow, p->N, p->D, p->H, p->W, p->Cs,
                                                                               p->Cl, p->low_act, p->low_act_param);
    });
    MEDNET_LAUNCH_CHECK();
  }
  return MEDNET_OK;
}


// ---------------------------------------------------------------------------------------------- split variants
namespace {
struct SplitWs { size_t pa, pb, ab, ab_a, ab_b, total; };
SplitWs split_fwd_ws(const UpcatPlan& u, int N, int Cs, int Cl) {
  SplitWs w;
  w.pa = 0;
  w.pb = u.pa_bytes;
  w.ab = w.pb + u.pb_bytes;
  w.ab_a = w.ab + align_up((size_t)N * 2 * (Cs + Cl) * sizeof(float), 256);
  w.ab_b = w.ab_a + align_up((size_t)N * 2 * Cs * sizeof(float), 256);
  w.total = w.ab_b + align_up((size_t)N * 2 * Cl * sizeof(float), 256);
  return w;
}
}  // namespace

extern "C" size_t mednet_upcat_groupnorm_split_fwd_workspace_bytes(const mednet_upcat_gn_split_fwd_params* p) {
  if (!p || !dtype_ok(p->dtype) || !upcat_geometry_ok(p->N, p->D, p->H, p->W, p->d, p->h, p->w, p->Cs, p->Cl, p->G)) return 0;
  UpcatPlan u = upcat_plan(p->N, p->D, p->H, p->W, p->Cs, p->Cl, p->dtype);
  return split_fwd_ws(u, p->N, p->Cs, p->Cl).total;
}

extern "C" int mednet_upcat_groupnorm_split_fwd(const mednet_upcat_gn_split_fwd_params* p, void* workspace,
                                                size_t workspace_bytes, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->skip && p->low && p->gamma && p->beta && p->y_skip && p->y_low && p->mean && p->rstd, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(upcat_geometry_ok(p->N, p->D, p->H, p->W, p->d, p->h, p->w, p->Cs, p->Cl, p->G), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_upcat_groupnorm_split_fwd_workspace_bytes(p), MEDNET_EWORKSPACE);
  UpcatPlan u = upcat_plan(p->N, p->D, p->H, p->W, p->Cs, p->Cl, p->dtype);
  const SplitWs w = split_fwd_ws(u, p->N, p->Cs, p->Cl);
  const int64_t S = (int64_t)p->D * p->H * p->W;
  char* ws = (char*)workspace;
  float* pa = (float*)(ws + w.pa);
  float* pb = (float*)(ws + w.pb);
  float* ab = (float*)(ws + w.ab);
  float* ab_a = (float*)(ws + w.ab_a);
  float* ab_b = (float*)(ws + w.ab_b);
  for (int part = 0; part < 2; ++part) {
    const SlabPlan& pl = part == 0 ? u.pa : u.pb;
    const void* src = part == 0 ? p->skip : p->low;
    const int Cp = part == 0 ? p->Cs : p->Cl;
    const int64_t Sp = part == 0 ? S : S / 8;
    dim3 grid(pl.nslab, (unsigned)p->N, pl.coltiles), block(pl.ncol_t, pl.R);
    MEDNET_DISPATCH_TV(p->dtype, pl.V, {
      size_t smem = (size_t)pl.ncol_t * pl.R * RED_CHUNK * sizeof(float);
      gn_partial_kernel<T, VV><<<grid, block, smem, stream>>>((const T*)src, part == 0 ? pa : pb, Sp, Cp, pl.ncol,
                                                              pl.rows_per_slab, pl.nslab);
    });
    MEDNET_LAUNCH_CHECK();
  }
  upcat_gn_finalize_kernel<<<(unsigned)(p->N * p->G), 128, 2 * ((p->Cs + p->Cl) / p->G) * sizeof(double), stream>>>(
      pa, pb, p->gamma, p->beta, p->mean, p->rstd, ab, S, p->Cs, p->Cl, p->G, u.pa.nslab, u.pb.nslab, p->eps);
  MEDNET_LAUNCH_CHECK();
  split_ab_kernel<<<ceil_div(p->N * 2 * (p->Cs + p->Cl), 256), 256, 0, stream>>>(ab, ab_a, ab_b, p->N, p->Cs, p->Cl);
  MEDNET_LAUNCH_CHECK();
  for (int part = 0; part < 2; ++part) {
    const SlabPlan& pl = part == 0 ? u.pa : u.pb;
    const int Cp = part == 0 ? p->Cs : p->Cl;
    const int64_t Sp = part == 0 ? S : S / 8;
    dim3 grid(pl.nslab, (unsigned)p->N, pl.coltiles), block(pl.ncol_t, pl.R);
    MEDNET_DISPATCH_TV(p->dtype, pl.V, {
      gn_apply_kernel<T, VV><<<grid, block, 0, stream>>>((const T*)(part == 0 ? p->skip : p->low), (const T*)nullptr,
                                                         (T*)(part == 0 ? p->y_skip : p->y_low), part == 0 ? ab_a : ab_b, Sp, Cp,
                                                         pl.ncol, pl.rows_per_slab, MEDNET_ACT_NONE, 0.f);
    });
    MEDNET_LAUNCH_CHECK();
  }
  return MEDNET_OK;
}

namespace {
struct SplitBwdWs { size_t pa, pb, coef_a, coef_b, dgb, total; };
SplitBwdWs split_bwd_ws(const UpcatPlan& u, int N, int Cs, int Cl) {
  SplitBwdWs w;
  w.pa = 0;
  w.pb = align_up((size_t)N * u.pa.nslab * 2 * Cs * sizeof(float), 256);
  w.coef_a = w.pb + align_up((size_t)N * u.pb.nslab * 2 * Cl * sizeof(float), 256);
  w.coef_b = w.coef_a + align_up((size_t)N * 3 * Cs * sizeof(float), 256);
  w.dgb = w.coef_b + align_up((size_t)N * 3 * Cl * sizeof(float), 256);
  w.total = w.dgb + align_up((size_t)N * 2 * (Cs + Cl) * sizeof(float), 256);
  return w;
}
}  // namespace

extern "C" size_t mednet_upcat_groupnorm_split_bwd_workspace_bytes(const mednet_upcat_gn_split_bwd_params* p) {
  if (!p || !dtype_ok(p->dtype) || !upcat_geometry_ok(p->N, p->D, p->H, p->W, p->d, p->h, p->w, p->Cs, p->Cl, p->G)) return 0;
  UpcatPlan u = upcat_plan(p->N, p->D, p->H, p->W, p->Cs, p->Cl, p->dtype);
  return split_bwd_ws(u, p->N, p->Cs, p->Cl).total;
}

extern "C" int mednet_upcat_groupnorm_split_bwd(const mednet_upcat_gn_split_bwd_params* p, void* workspace,
                                                size_t workspace_bytes, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->skip && p->low && p->dy_skip && p->dy_low && p->gamma && p->mean && p->rstd && p->dskip && p->dlow &&
                     p->dgamma && p->dbeta, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(upcat_geometry_ok(p->N, p->D, p->H, p->W, p->d, p->h, p->w, p->Cs, p->Cl, p->G), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_upcat_groupnorm_split_bwd_workspace_bytes(p), MEDNET_EWORKSPACE);
  UpcatPlan u = upcat_plan(p->N, p->D, p->H, p->W, p->Cs, p->Cl, p->dtype);
  const SplitBwdWs w = split_bwd_ws(u, p->N, p->Cs, p->Cl);
  const int64_t S = (int64_t)p->D * p->H * p->W;
  const int C = p->Cs + p->Cl;
  char* ws = (char*)workspace;
  float* pa = (float*)(ws + w.pa);
  float* pb = (float*)(ws + w.pb);
  float* coef_a = (float*)(ws + w.coef_a);
  float* coef_b = (float*)(ws + w.coef_b);
  float* dgb = (float*)(ws + w.dgb);
  for (int part = 0; part < 2; ++part) {
    const SlabPlan& pl = part == 0 ? u.pa : u.pb;
    const int Cp = part == 0 ? p->Cs : p->Cl;
    const int64_t Sp = part == 0 ? S : S / 8;
    dim3 grid(pl.nslab, (unsigned)p->N, pl.coltiles), block(pl.ncol_t, pl.R);
    MEDNET_DISPATCH_TV(p->dtype, pl.V, {
      size_t smem = (size_t)pl.ncol_t * pl.R * RED_CHUNK * sizeof(float);
      gn_bwd_partial_kernel<T, VV><<<grid, block, smem, stream>>>(
          (const T*)(part == 0 ? p->skip : p->low), (const T*)nullptr, (const T*)(part == 0 ? p->dy_skip : p->dy_low),
          part == 0 ? pa : pb, Sp, Cp, pl.ncol, pl.rows_per_slab, pl.nslab, MEDNET_ACT_NONE, 0.f);
    });
    MEDNET_LAUNCH_CHECK();
  }
  split_gn_bwd_finalize_kernel<<<(unsigned)(p->N * p->G), 128, 0, stream>>>(pa, pb, p->gamma, p->mean, p->rstd, coef_a, coef_b,
                                                                            dgb, S, p->Cs, p->Cl, p->G, u.pa.nslab, u.pb.nslab);
  MEDNET_LAUNCH_CHECK();
  gn_bwd_param_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(dgb, p->dgamma, p->dbeta, p->N, C, p->accumulate);
  MEDNET_LAUNCH_CHECK();
  for (int part = 0; part < 2; ++part) {
    const SlabPlan& pl = part == 0 ? u.pa : u.pb;
    const int Cp = part == 0 ? p->Cs : p->Cl;
    const int64_t Sp = part == 0 ? S : S / 8;
    dim3 grid(pl.nslab, (unsigned)p->N, pl.coltiles), block(pl.ncol_t, pl.R);
    MEDNET_DISPATCH_TV(p->dtype, pl.V, {
      gn_bwd_apply_kernel<T, VV, false><<<grid, block, 0, stream>>>(
          (const T*)(part == 0 ? p->skip : p->low), (const T*)nullptr, (const T*)(part == 0 ? p->dy_skip : p->dy_low),
          part == 0 ? coef_a : coef_b, (T*)(part == 0 ? p->dskip : p->dlow), (T*)nullptr, Sp, Cp, pl.ncol, pl.rows_per_slab,
          MEDNET_ACT_NONE, 0.f, part == 0 ? p->skip_act : p->low_act, part == 0 ? p->skip_act_param : p->low_act_param);
    });
    MEDNET_LAUNCH_CHECK();
  }
  return MEDNET_OK;
}
