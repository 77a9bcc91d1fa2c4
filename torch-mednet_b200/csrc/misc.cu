// Optimiser, landmark utilities, sliding-window tile movers and library-level entry points.
#include <mutex>

#include "common.cuh"

namespace mednet {

int sm_count_cached() {
  static int cached = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;  // B200; not cached so a later call can still query a live device
  }
  return cached;
}

// ------------------------------------------------------------------------------------------------
// fused Adam (ref: midasmednet/segmentation.py:119-120 -> torch.optim.Adam defaults)
// ------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                            float grad_scale, float bc1, float bc2_sqrt) {
  const float step_size = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
}

// ------------------------------------------------------------------------------------------------
// heatmap rendering: out[n,l,d,h,w] = (uint8) 255 * exp(-r^2 / (2 sigma^2))   (spec: oracle/heatmaps.py)
// ------------------------------------------------------------------------------------------------
__global__ void heatmap_render_kernel(const float* __restrict__ points, const float* __restrict__ sigmas,
                                      uint8_t* __restrict__ out, int N, int L, int D, int H, int W) {
  const int64_t S = (int64_t)D * H * W, total = (int64_t)N * L * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t nl = i / S;
    int64_t s = i - nl * S;
    const int w = (int)(s % W); s /= W;
    const int h = (int)(s % H);
    const int d = (int)(s / H);
    const int l = (int)(nl % L);
    const float* p = points + nl * 3;
    const float dd = (float)d - p[0], dh = (float)h - p[1], dw = (float)w - p[2];
    const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(dh, dh)), __fmul_rn(dw, dw));
    const float sg = sigmas[l];
    const float inv = __fdiv_rn(1.f, __fmul_rn(__fmul_rn(2.f, sg), sg));
    const float v = __fmul_rn(255.f, expf(-__fmul_rn(r2, inv)));
    out[i] = (uint8_t)(int)v;
  }
}

// ------------------------------------------------------------------------------------------------
// landmark extraction: per row (n,l) first-max argmax, optional soft-argmax (two-stage)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void lm_argmax_partial_kernel(const T* __restrict__ hm, float* __restrict__ pval, int64_t* __restrict__ pidx,
                                         int64_t S, int64_t per_chunk) {
  __shared__ float sv[32];
  __shared__ int64_t si[32];
  const int64_t row = blockIdx.y;
  const T* p = hm + row * S;
  const int64_t s0 = (int64_t)blockIdx.x * per_chunk;
  int64_t s1 = s0 + per_chunk;
  if (s1 > S) s1 = S;
  float best = -INFINITY;
  int64_t arg = INT64_MAX;
  for (int64_t s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
    const float v = to_f32<T>(p[s]);
    if (v > best || arg == INT64_MAX) {   // strict >: earliest index among equals within a thread
      best = v;
      arg = s;
    }
  }
  // lexicographic (value desc, index asc) reduction
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = best; si[warp] = arg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < arg)) { best = sv[w]; arg = si[w]; }
    pval[row * gridDim.x + blockIdx.x] = best;
    pidx[row * gridDim.x + blockIdx.x] = arg;
  }
}

__global__ void lm_argmax_final_kernel(const float* __restrict__ pval, const int64_t* __restrict__ pidx,
                                       int64_t* __restrict__ argmax, float* __restrict__ rowmax, int64_t NL, int chunks,
                                       int H, int W) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= NL) return;
  float best = pval[row * chunks];
  int64_t arg = pidx[row * chunks];
  for (int k = 1; k < chunks; ++k) {
    const float v = pval[row * chunks + k];
    const int64_t i = pidx[row * chunks + k];
    if (v > best || (v == best && i < arg)) { best = v; arg = i; }
  }
  rowmax[row] = best;
  argmax[row * 3 + 0] = arg / ((int64_t)H * W);
  argmax[row * 3 + 1] = (arg / W) % H;
  argmax[row * 3 + 2] = arg % W;
}

template <typename T>
__global__ void lm_soft_partial_kernel(const T* __restrict__ hm, const float* __restrict__ rowmax,
                                       float* __restrict__ psum, int64_t S, int64_t per_chunk, int H, int W, float beta) {
  __shared__ float scratch[32];
  const int64_t row = blockIdx.y;
  const T* p = hm + row * S;
  const float mx = rowmax[row];
  const int64_t s0 = (int64_t)blockIdx.x * per_chunk;
  int64_t s1 = s0 + per_chunk;
  if (s1 > S) s1 = S;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int64_t s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
    const float e = expf(beta * (to_f32<T>(p[s]) - mx));
    const int w = (int)(s % W), h = (int)((s / W) % H), d = (int)(s / ((int64_t)W * H));
    a0 += e; a1 += e * (float)d; a2 += e * (float)h; a3 += e * (float)w;
  }
  a0 = block_sum(a0, scratch); a1 = block_sum(a1, scratch); a2 = block_sum(a2, scratch); a3 = block_sum(a3, scratch);
  if (threadIdx.x == 0) {
    float* o = psum + (row * gridDim.x + blockIdx.x) * 4;
    o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
  }
}

__global__ void lm_soft_final_kernel(const float* __restrict__ psum, float* __restrict__ soft, int64_t NL, int chunks) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= NL) return;
  double a[4] = {0, 0, 0, 0};
  for (int k = 0; k < chunks; ++k)
    for (int j = 0; j < 4; ++j) a[j] += (double)psum[(row * chunks + k) * 4 + j];
  for (int j = 0; j < 3; ++j) soft[row * 3 + j] = (float)(a[j + 1] / a[0]);
}

static int lm_chunks(int64_t NL, int64_t S, int64_t* per_chunk) {
  int64_t chunks = ((int64_t)sm_count_cached() * 4) / NL;
  const int64_t maxc = S / 4096;
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  *per_chunk = ceil_div64(S, chunks);
  return (int)ceil_div64(S, *per_chunk);
}

// ------------------------------------------------------------------------------------------------
// sliding-window tile movers (ref: midasmednet/dataset.py:349-389 and :444-474)
// ------------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void tile_gather_kernel(const TS* __restrict__ vol, TD* __restrict__ tiles, const int32_t* __restrict__ org,
                                   mednet_tile_gather_params p) {
  const int64_t per_tile = (int64_t)p.P0 * p.P1 * p.P2 * p.C;
  const int64_t total = per_tile * p.B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_tile);
    int64_t t = i - (int64_t)b * per_tile;
    int c, x, y, z;
    if (p.ncdhw_out) {                         // [B, C, P0, P1, P2]
      z = (int)(t % p.P2); t /= p.P2;
      y = (int)(t % p.P1); t /= p.P1;
      x = (int)(t % p.P0);
      c = (int)(t / p.P0);
    } else {                                   // [B, P0, P1, P2, C]
      c = (int)(t % p.C); t /= p.C;
      z = (int)(t % p.P2); t /= p.P2;
      y = (int)(t % p.P1);
      x = (int)(t / p.P1);
    }
    // padded coordinate -> original coordinate (low pad = overlap); outside -> constant 0 (predict.py:68)
    const int gx = org[b * 3 + 0] + x - p.O0, gy = org[b * 3 + 1] + y - p.O1, gz = org[b * 3 + 2] + z - p.O2;
    float v = 0.f;
    if (gx >= 0 && gx < p.X && gy >= 0 && gy < p.Y && gz >= 0 && gz < p.Z)
      v = to_f32<TS>(vol[(((int64_t)c * p.X + gx) * p.Y + gy) * p.Z + gz]);
    tiles[i] = from_f32<TD>(v);
  }
}

// Row variant for channel-major tiles (ncdhw_out) and for C == 1, where an output row along axis 2 is one contiguous run
// of the source volume: a warp per (tile, channel, x, y) row, lanes along z -- coalesced, one index decomposition per
// row instead of five 64-bit divisions per element.
template <typename TS, typename TD>
__global__ void tile_gather_rows_kernel(const TS* __restrict__ vol, TD* __restrict__ tiles, const int32_t* __restrict__ org,
                                        mednet_tile_gather_params p) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int64_t rows = (int64_t)p.B * p.C * p.P0 * p.P1;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    int64_t t = r;
    const int y = (int)(t % p.P1); t /= p.P1;
    const int x = (int)(t % p.P0); t /= p.P0;
    const int c = (int)(t % p.C);
    const int b = (int)(t / p.C);
    const int gx = org[b * 3 + 0] + x - p.O0, gy = org[b * 3 + 1] + y - p.O1, gz0 = org[b * 3 + 2] - p.O2;
    const bool inside = gx >= 0 && gx < p.X && gy >= 0 && gy < p.Y;
    const TS* src = vol + (((int64_t)c * p.X + (inside ? gx : 0)) * p.Y + (inside ? gy : 0)) * p.Z;
    TD* dst = tiles + r * p.P2;
    for (int z = lane; z < p.P2; z += 32) {
      const int gz = gz0 + z;
      float v = 0.f;
      if (inside && gz >= 0 && gz < p.Z) v = to_f32<TS>(src[gz]);
      dst[z] = from_f32<TD>(v);
    }
  }
}

// Batch crop with a per-sample source table: a warp per (sample, channel, x, y) row, lanes along z.
template <typename TS, typename TD>
__global__ void patch_gather_kernel(const int64_t* __restrict__ table, TD* __restrict__ tiles, mednet_patch_gather_params p) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int64_t rows = (int64_t)p.B * p.C * p.P0 * p.P1;
  const int64_t stride = p.tile_stride ? p.tile_stride : (int64_t)p.C * p.P0 * p.P1 * p.P2;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    int64_t t = r;
    const int y = (int)(t % p.P1); t /= p.P1;
    const int x = (int)(t % p.P0); t /= p.P0;
    const int c = (int)(t % p.C);
    const int b = (int)(t / p.C);
    const int64_t* e = table + (int64_t)b * 8;
    const TS* vol = reinterpret_cast<const TS*>(e[0]);
    const int X = (int)e[1], Y = (int)e[2], Z = (int)e[3];
    const int gx = (int)e[4] + x, gy = (int)e[5] + y, gz0 = (int)e[6];
    const bool inside = gx >= 0 && gx < X && gy >= 0 && gy < Y;
    const TS* src = vol + (((int64_t)c * X + (inside ? gx : 0)) * Y + (inside ? gy : 0)) * Z;
    TD* dst;
    int zs;
    TD* tile = tiles + (int64_t)b * stride;
    if (p.ncdhw_out) {
      dst = tile + (((int64_t)c * p.P0 + x) * p.P1 + y) * p.P2, zs = 1;
    } else {
      dst = tile + (((int64_t)x * p.P1 + y) * p.P2) * p.C + c, zs = p.C;
    }
    for (int z = lane; z < p.P2; z += 32) {
      const int gz = gz0 + z;
      float v = 0.f;
      if (inside && gz >= 0 && gz < Z) v = to_f32<TS>(src[gz]);
      dst[(int64_t)z * zs] = from_f32<TD>(v);
    }
  }
}

__global__ void tile_scatter_kernel(const uint8_t* __restrict__ tiles, uint8_t* __restrict__ vol,
                                    const int32_t* __restrict__ org, mednet_tile_scatter_params p) {
  const int c0 = p.P0 - 2 * p.O0, c1 = p.P1 - 2 * p.O1, c2 = p.P2 - 2 * p.O2;
  const int64_t per_tile = (int64_t)p.Co * c0 * c1 * c2;
  const int64_t total = per_tile * p.B;
  const int64_t PS = (int64_t)p.P0 * p.P1 * p.P2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_tile);
    int64_t t = i - (int64_t)b * per_tile;
    const int z = (int)(t % c2); t /= c2;
    const int y = (int)(t % c1); t /= c1;
    const int x = (int)(t % c0);
    const int c = (int)(t / c0);
    const int gx = org[b * 3 + 0] + x, gy = org[b * 3 + 1] + y, gz = org[b * 3 + 2] + z;   // crop start == tile origin
    if (gx < p.X && gy < p.Y && gz < p.Z) {
      const int64_t src = ((int64_t)b * p.Co + c) * PS + (((int64_t)(x + p.O0) * p.P1) + (y + p.O1)) * p.P2 + (z + p.O2);
      vol[(((int64_t)c * p.X + gx) * p.Y + gy) * p.Z + gz] = tiles[src];
    }
  }
}

}  // namespace mednet

using namespace mednet;

extern "C" int mednet_abi_version(void) { return MEDNET_ABI_VERSION; }

extern "C" const char* mednet_error_string(int code) {
  switch (code) {
    case MEDNET_OK: return "success";
    case MEDNET_EINVAL: return "invalid argument";
    case MEDNET_EUNSUPPORTED: return "unsupported shape or dtype (no fallback by design)";
    case MEDNET_EALIGN: return "pointer or channel count violates the alignment the kernel needs";
    case MEDNET_EWORKSPACE: return "workspace missing or too small";
    case MEDNET_ENODRIVER: return "CUDA driver entry point unavailable";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown mednet error";
  }
}

extern "C" int mednet_device_has_tcgen05(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" int mednet_sm_count(void) { return sm_count_cached(); }

extern "C" int mednet_adam_step(const mednet_adam_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->param && p->grad && p->exp_avg && p->exp_avg_sq && p->numel > 0 && p->step >= 1, MEDNET_EINVAL);
  const double bc1 = 1.0 - pow((double)p->beta1, (double)p->step);
  const double bc2 = 1.0 - pow((double)p->beta2, (double)p->step);
  adam_kernel<<<grid_for(p->numel, 256), 256, 0, stream>>>(p->param, p->grad, p->exp_avg, p->exp_avg_sq, p->numel, p->lr,
                                                          p->beta1, p->beta2, p->eps, p->grad_scale, (float)bc1,
                                                          (float)sqrt(bc2));
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_heatmap_render(const mednet_hmrender_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->points && p->sigmas && p->out && p->N > 0 && p->L > 0 && p->D > 0 && p->H > 0 && p->W > 0,
                 MEDNET_EINVAL);
  const int64_t total = (int64_t)p->N * p->L * p->D * p->H * p->W;
  heatmap_render_kernel<<<grid_for(total, 256), 256, 0, stream>>>(p->points, p->sigmas, p->out, p->N, p->L, p->D, p->H,
                                                                 p->W);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" size_t mednet_landmark_workspace_bytes(const mednet_landmark_params* p) {
  if (!p || p->NL <= 0) return 0;
  int64_t pc;
  const int chunks = lm_chunks(p->NL, (int64_t)p->D * p->H * p->W, &pc);
  return align_up((size_t)p->NL * chunks * sizeof(float), 256) + align_up((size_t)p->NL * chunks * sizeof(int64_t), 256) +
         align_up((size_t)p->NL * sizeof(float), 256) + align_up((size_t)p->NL * chunks * 4 * sizeof(float), 256);
}

extern "C" int mednet_landmark_extract(const mednet_landmark_params* p, void* workspace, size_t workspace_bytes,
                                       mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->heatmaps && p->argmax && p->NL > 0 && p->NL <= 65535 && p->D > 0 && p->H > 0 && p->W > 0,
                 MEDNET_EINVAL);
  MEDNET_REQUIRE(p->dtype == MEDNET_F32 || p->dtype == MEDNET_BF16 || p->dtype == MEDNET_U8, MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_landmark_workspace_bytes(p), MEDNET_EWORKSPACE);
  const int64_t S = (int64_t)p->D * p->H * p->W;
  int64_t pc;
  const int chunks = lm_chunks(p->NL, S, &pc);
  char* ws = (char*)workspace;
  float* pval = (float*)ws; ws += align_up((size_t)p->NL * chunks * sizeof(float), 256);
  int64_t* pidx = (int64_t*)ws; ws += align_up((size_t)p->NL * chunks * sizeof(int64_t), 256);
  float* rowmax = (float*)ws; ws += align_up((size_t)p->NL * sizeof(float), 256);
  float* psum = (float*)ws;
  dim3 grid(chunks, (unsigned)p->NL);
  if (p->dtype == MEDNET_F32) lm_argmax_partial_kernel<float><<<grid, 256, 0, stream>>>((const float*)p->heatmaps, pval, pidx, S, pc);
  else if (p->dtype == MEDNET_BF16) lm_argmax_partial_kernel<bf16><<<grid, 256, 0, stream>>>((const bf16*)p->heatmaps, pval, pidx, S, pc);
  else lm_argmax_partial_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)p->heatmaps, pval, pidx, S, pc);
  MEDNET_LAUNCH_CHECK();
  lm_argmax_final_kernel<<<(unsigned)ceil_div64(p->NL, 128), 128, 0, stream>>>(pval, pidx, p->argmax, rowmax, p->NL, chunks,
                                                                             p->H, p->W);
  MEDNET_LAUNCH_CHECK();
  if (p->soft != nullptr) {
    if (p->dtype == MEDNET_F32) lm_soft_partial_kernel<float><<<grid, 256, 0, stream>>>((const float*)p->heatmaps, rowmax, psum, S, pc, p->H, p->W, p->beta);
    else if (p->dtype == MEDNET_BF16) lm_soft_partial_kernel<bf16><<<grid, 256, 0, stream>>>((const bf16*)p->heatmaps, rowmax, psum, S, pc, p->H, p->W, p->beta);
    else lm_soft_partial_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)p->heatmaps, rowmax, psum, S, pc, p->H, p->W, p->beta);
    MEDNET_LAUNCH_CHECK();
    lm_soft_final_kernel<<<(unsigned)ceil_div64(p->NL, 128), 128, 0, stream>>>(psum, p->soft, p->NL, chunks);
    MEDNET_LAUNCH_CHECK();
  }
  return MEDNET_OK;
}

extern "C" int mednet_tile_gather(const mednet_tile_gather_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->volume && p->tiles && p->origins && p->B > 0 && p->C > 0, MEDNET_EINVAL);
  const int64_t total = (int64_t)p->B * p->C * p->P0 * p->P1 * p->P2;
  const bool rows = p->ncdhw_out || p->C == 1;                 // for C == 1 the two tile layouts coincide
  const int nb = rows ? grid_for((int64_t)p->B * p->C * p->P0 * p->P1 * 32, 256) : grid_for(total, 256);
#define MEDNET_GATHER(TS, TD)                                                                                         \
  do {                                                                                                                \
    if (rows)                                                                                                         \
      tile_gather_rows_kernel<TS, TD><<<nb, 256, 0, stream>>>((const TS*)p->volume, (TD*)p->tiles, p->origins, *p);   \
    else                                                                                                              \
      tile_gather_kernel<TS, TD><<<nb, 256, 0, stream>>>((const TS*)p->volume, (TD*)p->tiles, p->origins, *p);        \
  } while (0)
  if (p->src_dtype == MEDNET_U8 && p->dst_dtype == MEDNET_U8) {
    MEDNET_GATHER(uint8_t, uint8_t);
    MEDNET_LAUNCH_CHECK();
    return MEDNET_OK;
  }
  MEDNET_REQUIRE(dtype_ok(p->src_dtype) && dtype_ok(p->dst_dtype), MEDNET_EUNSUPPORTED);
  if (p->src_dtype == MEDNET_F32 && p->dst_dtype == MEDNET_F32)
    MEDNET_GATHER(float, float);
  else if (p->src_dtype == MEDNET_F32)
    MEDNET_GATHER(float, bf16);
  else if (p->dst_dtype == MEDNET_F32)
    MEDNET_GATHER(bf16, float);
  else
    MEDNET_GATHER(bf16, bf16);
#undef MEDNET_GATHER
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_patch_gather(const mednet_patch_gather_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->table && p->tiles && p->B > 0 && p->C > 0 && p->P0 > 0 && p->P1 > 0 && p->P2 > 0, MEDNET_EINVAL);
  MEDNET_REQUIRE(p->tile_stride == 0 || p->tile_stride >= (int64_t)p->C * p->P0 * p->P1 * p->P2, MEDNET_EINVAL);
  const int nb = grid_for((int64_t)p->B * p->C * p->P0 * p->P1 * 32, 256);
  if (p->src_dtype == MEDNET_U8 && p->dst_dtype == MEDNET_U8)
    patch_gather_kernel<uint8_t, uint8_t><<<nb, 256, 0, stream>>>(p->table, (uint8_t*)p->tiles, *p);
  else {
    MEDNET_REQUIRE(dtype_ok(p->src_dtype) && dtype_ok(p->dst_dtype), MEDNET_EUNSUPPORTED);
    if (p->src_dtype == MEDNET_F32 && p->dst_dtype == MEDNET_F32)
      patch_gather_kernel<float, float><<<nb, 256, 0, stream>>>(p->table, (float*)p->tiles, *p);
    else if (p->src_dtype == MEDNET_F32)
      patch_gather_kernel<float, bf16><<<nb, 256, 0, stream>>>(p->table, (bf16*)p->tiles, *p);
    else if (p->dst_dtype == MEDNET_F32)
      patch_gather_kernel<bf16, float><<<nb, 256, 0, stream>>>(p->table, (float*)p->tiles, *p);
    else
      patch_gather_kernel<bf16, bf16><<<nb, 256, 0, stream>>>(p->table, (bf16*)p->tiles, *p);
  }
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_tile_scatter(const mednet_tile_scatter_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->tiles && p->volume && p->origins && p->B > 0 && p->Co > 0, MEDNET_EINVAL);
  MEDNET_REQUIRE(p->P0 > 2 * p->O0 && p->P1 > 2 * p->O1 && p->P2 > 2 * p->O2, MEDNET_EINVAL);
  const int64_t total = (int64_t)p->B * p->Co * (p->P0 - 2 * p->O0) * (p->P1 - 2 * p->O1) * (p->P2 - 2 * p->O2);
  tile_scatter_kernel<<<grid_for(total, 256), 256, 0, stream>>>(p->tiles, p->volume, p->origins, *p);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}
