// Encoder/decoder glue, NDHWC, all HBM-bound with 16-byte vector accesses along the channel axis:
//   * layout/dtype conversion at the module boundary
//   * MaxPool3d(2) with window-local argmax codes            (ref: midasmednet/unet/components.py:210,224)
//   * nearest upsample fused with the skip concat            (ref: midasmednet/unet/components.py:277-280)
#include "common.cuh"

namespace mednet {

// ------------------------------------------------------------------------------------------------
// layout conversion: NCDHW <-> NDHWC with dtype cast.  32x32 smem tile transpose over (C, S).
// ------------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void transpose_cs_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t R, int64_t Ccols) {
  // src is [n][R][Ccols] (row-major), dst is [n][Ccols][R]
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  const TS* s = src + n * R * Ccols;
  TD* d = dst + n * R * Ccols;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Ccols) tile[i][threadIdx.x] = to_f32<TS>(s[r * Ccols + c]);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Ccols) d[c * R + r] = from_f32<TD>(tile[threadIdx.x][i]);
  }
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = from_f32<TD>(to_f32<TS>(src[i]));
}

template <typename TS, typename TD>
static int launch_layout(const mednet_layout_params* p, cudaStream_t st) {
  if (p->C == 1) {  // both layouts coincide
    cast_kernel<TS, TD><<<grid_for(p->N * p->S, 256), 256, 0, st>>>((const TS*)p->src, (TD*)p->dst, p->N * p->S);
  } else {
    const int64_t R = p->to_channels_last ? p->C : p->S;
    const int64_t Cc = p->to_channels_last ? p->S : p->C;
    MEDNET_REQUIRE(ceil_div64(R, 32) <= 65535 && p->N <= 65535, MEDNET_EUNSUPPORTED);
    dim3 grid((unsigned)ceil_div64(Cc, 32), (unsigned)ceil_div64(R, 32), (unsigned)p->N), block(32, 8);
    transpose_cs_kernel<TS, TD><<<grid, block, 0, st>>>((const TS*)p->src, (TD*)p->dst, R, Cc);
  }
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

// ------------------------------------------------------------------------------------------------
// max pool 2x2x2 / stride 2 / floor mode
// ------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ idx, int N,
                                   int D, int H, int W, int C) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2, ncol = C / V;
  const int64_t total = (int64_t)N * Do * Ho * Wo * ncol;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % ncol);
    int64_t t = i / ncol;
    const int ow = (int)(t % Wo); t /= Wo;
    const int oh = (int)(t % Ho); t /= Ho;
    const int od = (int)(t % Do);
    const int n = (int)(t / Do);
    float best[V];
    int code[V];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
      const int64_t off = ((((int64_t)n * D + (2 * od + dz)) * H + (2 * oh + dy)) * W + (2 * ow + dx)) * C + cv * V;
      float v[V];
      load_vec<T, V>(x + off, v);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        // ATen rule: (val > max) || isnan(val) -> first maximum wins, NaN wins
        if (k == 0 || v[j] > best[j] || v[j] != v[j]) {
          best[j] = v[j];
          code[j] = k;
        }
      }
    }
    const int64_t o = ((((int64_t)n * Do + od) * Ho + oh) * Wo + ow) * C + cv * V;
    store_vec<T, V>(y + o, best);
    union { typename RawVec<V>::type r; uint8_t b[V]; } u;
#pragma unroll
    for (int j = 0; j < V; ++j) u.b[j] = (uint8_t)code[j];
    *reinterpret_cast<typename RawVec<V>::type*>(idx + o) = u.r;
  }
}

// One block per input LINE (n, d, h): thread (tx, ty) owns channel vector tx and walks w = ty, ty + R, ...  Reads of the
// skip-path gradient (addend) and writes of dx are fully coalesced rows; the pooled gradient / argmax code / pooled value
// of the parent voxel are read 8 times (by its 8 children) out of L1/L2.  No per-element index division.
template <typename T, int V>
__global__ void maxpool_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx, T* __restrict__ dx,
                                   int N, int D, int H, int W, int C, const T* __restrict__ y, int in_act,
                                   float in_act_param, const T* __restrict__ addend) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2, ncol = C / V;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= ncol) return;
  const int line = blockIdx.x;                       // (n * D + d) * H + h
  const int h = line % H, nd = line / H, d = nd % D, n = nd / D;
  const int od = d >> 1, oh = h >> 1;
  const bool line_in = od < Do && oh < Ho;
  const int dh_code = ((d & 1) << 2) | ((h & 1) << 1);
  const int64_t lbase = (int64_t)line * W * C + cv * V;
  const int64_t pbase = (((int64_t)n * Do + od) * Ho + oh) * (int64_t)Wo * C + cv * V;
#pragma unroll 2                                     // two voxels' loads in flight per thread
  for (int w = threadIdx.y; w < W; w += blockDim.y) {
    float g[V];
#pragma unroll
    for (int j = 0; j < V; ++j) g[j] = 0.f;
    const int ow = w >> 1;
    if (line_in && ow < Wo) {
      const int64_t o = pbase + (int64_t)ow * C;
      const int mine = dh_code | (w & 1);
      float gy[V];
      load_vec<T, V>(dy + o, gy);
      union { typename RawVec<V>::type r; uint8_t b[V]; } u;
      u.r = *reinterpret_cast<const typename RawVec<V>::type*>(idx + o);
      if (in_act != MEDNET_ACT_NONE) {     // deferred derivative of the producer's activation: x[argmax] == y
        float yv[V];
        load_vec<T, V>(y + o, yv);
#pragma unroll
        for (int j = 0; j < V; ++j) gy[j] *= act_grad_from_out(yv[j], in_act, in_act_param);
      }
#pragma unroll
      for (int j = 0; j < V; ++j) g[j] = (u.b[j] == mine) ? gy[j] : 0.f;
    }
    const int64_t off = lbase + (int64_t)w * C;
    if (addend != nullptr) {
      float a[V];
      load_vec<T, V>(addend + off, a);
#pragma unroll
      for (int j = 0; j < V; ++j) g[j] += a[j];
    }
    store_vec<T, V>(dx + off, g);
  }
}

__global__ void pool_idx_i64_kernel(const uint8_t* __restrict__ idx, int64_t* __restrict__ out, int N, int D, int H,
                                    int W, int C) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const int64_t total = (int64_t)N * C * Do * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i enumerates the NCDHW output
    int64_t t = i;
    const int ow = (int)(t % Wo); t /= Wo;
    const int oh = (int)(t % Ho); t /= Ho;
    const int od = (int)(t % Do); t /= Do;
    const int c = (int)(t % C);
    const int n = (int)(t / C);
    const int k = idx[((((int64_t)n * Do + od) * Ho + oh) * Wo + ow) * C + c];
    out[i] = ((int64_t)(2 * od + (k >> 2)) * H + (2 * oh + ((k >> 1) & 1))) * W + (2 * ow + (k & 1));
  }
}

// ------------------------------------------------------------------------------------------------
// nearest upsample (+ concat).  ATen's legacy 'nearest' source index, computed in fp32 exactly as
// upstream: src = min((int)floorf(dst * ((float)in / out)), in - 1).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  const int s = (int)floorf(__fmul_rn((float)dst, scale));
  return s < in_size - 1 ? s : in_size - 1;
}

template <typename T, int V>
__global__ void upcat_fwd_kernel(const T* __restrict__ skip, const T* __restrict__ low, T* __restrict__ out, int N,
                                 int D, int H, int W, int d, int h, int w, int Cs, int Cl) {
  const int C = Cs + Cl, ncol = C / V, ncs = Cs / V;
  const float sd = (float)d / (float)D, sh = (float)h / (float)H, sw = (float)w / (float)W;
  const int64_t total = (int64_t)N * D * H * W * ncol;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % ncol);
    const int64_t vox = i / ncol;
    typedef typename RawVec<sizeof(T) * V>::type R;
    R val;
    if (cv < ncs) {
      val = *reinterpret_cast<const R*>(skip + vox * Cs + cv * V);
    } else {
      int64_t t = vox;
      const int x = (int)(t % W); t /= W;
      const int y = (int)(t % H); t /= H;
      const int z = (int)(t % D);
      const int n = (int)(t / D);
      const int64_t src = (((int64_t)n * d + nearest_src(z, sd, d)) * h + nearest_src(y, sh, h)) * w + nearest_src(x, sw, w);
      val = *reinterpret_cast<const R*>(low + src * Cl + (cv - ncs) * V);
    }
    *reinterpret_cast<R*>(out + i * V) = val;
  }
}

// first destination index whose source is >= s (the mapping is monotone non-decreasing)
__device__ __forceinline__ int nearest_first_dst(int s, float scale, int in_size, int out_size) {
  int g = (int)(((int64_t)s * out_size) / in_size);
  if (g > out_size) g = out_size;
  while (g > 0 && nearest_src(g - 1, scale, in_size) >= s) --g;
  while (g < out_size && nearest_src(g, scale, in_size) < s) ++g;
  return g;
}

template <typename T, int V>
__global__ void upcat_bwd_skip_kernel(const T* __restrict__ dout, T* __restrict__ dskip, int64_t vox, int Cs, int Cl,
                                      const T* __restrict__ skip, int act, float act_param) {
  const int ncs = Cs / V;
  const int64_t total = vox * ncs;
  const int C = Cs + Cl;
  typedef typename RawVec<sizeof(T) * V>::type R;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % ncs);
    const int64_t v = i / ncs;
    if (act == MEDNET_ACT_NONE) {
      *reinterpret_cast<R*>(dskip + i * V) = *reinterpret_cast<const R*>(dout + v * C + cv * V);
    } else {
      float g[V], xv[V];
      load_vec<T, V>(dout + v * C + cv * V, g);
      load_vec<T, V>(skip + i * V, xv);
#pragma unroll
      for (int j = 0; j < V; ++j) g[j] *= act_grad_from_out(xv[j], act, act_param);
      store_vec<T, V>(dskip + i * V, g);
    }
  }
}

template <typename T, int V>
__global__ void upcat_bwd_low_kernel(const T* __restrict__ dout, T* __restrict__ dlow, int N, int D, int H, int W,
                                     int d, int h, int w, int Cs, int Cl, const T* __restrict__ low, int act,
                                     float act_param) {
  const int C = Cs + Cl, ncl = Cl / V;
  const float sd = (float)d / (float)D, sh = (float)h / (float)H, sw = (float)w / (float)W;
  const int64_t total = (int64_t)N * d * h * w * ncl;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % ncl);
    int64_t t = i / ncl;
    const int x = (int)(t % w); t /= w;
    const int y = (int)(t % h); t /= h;
    const int z = (int)(t % d);
    const int n = (int)(t / d);
    const int z0 = nearest_first_dst(z, sd, d, D), z1 = nearest_first_dst(z + 1, sd, d, D);
    const int y0 = nearest_first_dst(y, sh, h, H), y1 = nearest_first_dst(y + 1, sh, h, H);
    const int x0 = nearest_first_dst(x, sw, w, W), x1 = nearest_first_dst(x + 1, sw, w, W);
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    for (int zz = z0; zz < z1; ++zz)
      for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) {
          float g[V];
          load_vec<T, V>(dout + ((((int64_t)n * D + zz) * H + yy) * W + xx) * C + Cs + cv * V, g);
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] += g[j];
        }
    if (act != MEDNET_ACT_NONE) {
      float xv[V];
      load_vec<T, V>(low + i * V, xv);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] *= act_grad_from_out(xv[j], act, act_param);
    }
    store_vec<T, V>(dlow + i * V, acc);
  }
}

// dst[r][0..Cd) from src[r][0..Cs): zero fill past Cs (pad) or truncation (extract); one thread per destination row
template <typename T>
__global__ void channel_pad_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t rows, int Cs, int Cd) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    if (Cd == 16 && sizeof(T) == 2) {                       // the first-layer case: one 32-byte row per thread
      union { U32B v; T t[16]; } u;
#pragma unroll
      for (int c = 0; c < 16; ++c) u.t[c] = c < Cs ? src[r * Cs + c] : from_f32<T>(0.f);
      *reinterpret_cast<U32B*>(dst + r * 16) = u.v;
    } else {
      for (int c = 0; c < Cd; ++c) dst[r * Cd + c] = c < Cs ? src[r * Cs + c] : from_f32<T>(0.f);
    }
  }
}

}  // namespace mednet

using namespace mednet;

extern "C" int mednet_channel_pad(const mednet_chpad_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->src && p->dst && p->rows > 0 && p->Cs > 0 && p->Cd > 0, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  if (p->dtype == MEDNET_F32)
    channel_pad_kernel<float><<<grid_for(p->rows, 256), 256, 0, stream>>>((const float*)p->src, (float*)p->dst, p->rows, p->Cs, p->Cd);
  else
    channel_pad_kernel<bf16><<<grid_for(p->rows, 256), 256, 0, stream>>>((const bf16*)p->src, (bf16*)p->dst, p->rows, p->Cs, p->Cd);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_layout_convert(const mednet_layout_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->src && p->dst && p->N > 0 && p->C > 0 && p->S > 0, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->src_dtype) && dtype_ok(p->dst_dtype), MEDNET_EUNSUPPORTED);
  if (p->src_dtype == MEDNET_F32 && p->dst_dtype == MEDNET_F32) return launch_layout<float, float>(p, stream);
  if (p->src_dtype == MEDNET_F32 && p->dst_dtype == MEDNET_BF16) return launch_layout<float, bf16>(p, stream);
  if (p->src_dtype == MEDNET_BF16 && p->dst_dtype == MEDNET_F32) return launch_layout<bf16, float>(p, stream);
  return launch_layout<bf16, bf16>(p, stream);
}

extern "C" int mednet_maxpool3d_fwd(const mednet_pool_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->x && p->y && p->idx, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(p->N > 0 && p->C > 0 && p->D >= 2 && p->H >= 2 && p->W >= 2, MEDNET_EINVAL);
  const int V = pick_vec(p->C, dtype_bytes(p->dtype));
  const int64_t total = (int64_t)p->N * (p->D / 2) * (p->H / 2) * (p->W / 2) * (p->C / V);
  MEDNET_DISPATCH_TV(p->dtype, V, {
    maxpool_fwd_kernel<T, VV><<<grid_for(total, 256), 256, 0, stream>>>((const T*)p->x, (T*)p->y, p->idx, p->N, p->D,
                                                                       p->H, p->W, p->C);
  });
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_maxpool3d_bwd(const mednet_pool_bwd_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->dy && p->idx && p->dx, MEDNET_EINVAL);
  MEDNET_REQUIRE(p->in_act == MEDNET_ACT_NONE || p->y != nullptr, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(p->N > 0 && p->C > 0 && p->D >= 2 && p->H >= 2 && p->W >= 2, MEDNET_EINVAL);
  const int V = pick_vec(p->C, dtype_bytes(p->dtype));
  const int ncol = p->C / V;
  const int ncol_t = ncol < 256 ? ncol : 256;
  int R = 256 / ncol_t;
  if (R > p->W) R = p->W;
  if (R < 1) R = 1;
  const int64_t lines = (int64_t)p->N * p->D * p->H;
  MEDNET_REQUIRE(lines < ((int64_t)1 << 31) && ceil_div(ncol, ncol_t) <= 65535, MEDNET_EUNSUPPORTED);
  dim3 grid((unsigned)lines, (unsigned)ceil_div(ncol, ncol_t)), block(ncol_t, R);
  MEDNET_DISPATCH_TV(p->dtype, V, {
    maxpool_bwd_kernel<T, VV><<<grid, block, 0, stream>>>((const T*)p->dy, p->idx, (T*)p->dx, p->N, p->D, p->H, p->W, p->C,
                                                          (const T*)p->y, p->in_act, p->in_act_param, (const T*)p->addend);
  });
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_maxpool3d_indices_i64(const uint8_t* idx, int64_t* out, int32_t N, int32_t D, int32_t H,
                                            int32_t W, int32_t C, mednet_stream_t stream) {
  MEDNET_REQUIRE(idx && out && N > 0 && C > 0 && D >= 2 && H >= 2 && W >= 2, MEDNET_EINVAL);
  const int64_t total = (int64_t)N * C * (D / 2) * (H / 2) * (W / 2);
  pool_idx_i64_kernel<<<grid_for(total, 256), 256, 0, stream>>>(idx, out, N, D, H, W, C);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_upsample_concat_fwd(const mednet_upcat_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->low && p->out && (p->skip || p->Cs == 0), MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(p->N > 0 && p->D > 0 && p->H > 0 && p->W > 0 && p->d > 0 && p->h > 0 && p->w > 0 && p->Cs >= 0 &&
                     p->Cl > 0, MEDNET_EINVAL);
  int V = pick_vec(p->Cl, dtype_bytes(p->dtype));
  if (p->Cs > 0) { const int v2 = pick_vec(p->Cs, dtype_bytes(p->dtype)); if (v2 < V) V = v2; }
  const int64_t total = (int64_t)p->N * p->D * p->H * p->W * ((p->Cs + p->Cl) / V);
  MEDNET_DISPATCH_TV(p->dtype, V, {
    upcat_fwd_kernel<T, VV><<<grid_for(total, 256), 256, 0, stream>>>((const T*)p->skip, (const T*)p->low, (T*)p->out,
                                                                     p->N, p->D, p->H, p->W, p->d, p->h, p->w, p->Cs,
                                                                     p->Cl);
  });
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_upsample_concat_bwd(const mednet_upcat_bwd_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->dout && p->dlow && (p->dskip || p->Cs == 0), MEDNET_EINVAL);
  MEDNET_REQUIRE((p->skip_act == MEDNET_ACT_NONE || p->Cs == 0 || p->skip) && (p->low_act == MEDNET_ACT_NONE || p->low),
                 MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(p->N > 0 && p->D > 0 && p->H > 0 && p->W > 0 && p->d > 0 && p->h > 0 && p->w > 0 && p->Cs >= 0 &&
                     p->Cl > 0, MEDNET_EINVAL);
  int V = pick_vec(p->Cl, dtype_bytes(p->dtype));
  if (p->Cs > 0) { const int v2 = pick_vec(p->Cs, dtype_bytes(p->dtype)); if (v2 < V) V = v2; }
  const int64_t vox = (int64_t)p->N * p->D * p->H * p->W;
  const int64_t tl = (int64_t)p->N * p->d * p->h * p->w * (p->Cl / V);
  MEDNET_DISPATCH_TV(p->dtype, V, {
    if (p->Cs > 0) {
      upcat_bwd_skip_kernel<T, VV><<<grid_for(vox * (p->Cs / VV), 256), 256, 0, stream>>>(
          (const T*)p->dout, (T*)p->dskip, vox, p->Cs, p->Cl, (const T*)p->skip, p->skip_act, p->skip_act_param);
    }
    upcat_bwd_low_kernel<T, VV><<<grid_for(tl, 256), 256, 0, stream>>>((const T*)p->dout, (T*)p->dlow, p->N, p->D, p->H,
                                                                      p->W, p->d, p->h, p->w, p->Cs, p->Cl,
                                                                      (const T*)p->low, p->low_act, p->low_act_param);
  });
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}
