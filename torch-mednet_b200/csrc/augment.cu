// Intensity augmentation of sampled patches on the device: the chain the reference's training scripts compose
// (examples/train_seg.py:82-86, train_ldmks.py:82-84) -- additive brightness per channel, gamma over the whole
// sample, contrast per channel with range preservation -- with the random DECISIONS drawn on the host (same order as
// the CPU library) and handed in as coefficients; the kernels only evaluate them.
//
// Three launches over a patch batch that is small next to a training step (2 x 128^3 x C floats):
//   stats<0>  per (sample, channel, chunk): min / max of v = x + offset
//   stats<1>  per (sample, channel, chunk): sum / min / max of g(v), g = gamma curve between the sample's min and max
//   apply     y = clip((g - mean) * factor + mean, min g, max g)   -> fp32 or bf16, same NDHWC layout
// Partials are combined in a fixed order by the consumer (no float atomics): results are run-to-run identical.
#include "common.cuh"

namespace mednet {

constexpr int AUG_MAXC = 8;       // channels per patch handled by one thread
constexpr int AUG_THREADS = 256;
constexpr float AUG_EPS = 1e-7f;  // epsilon of the gamma curve's denominator

struct AugCoef {                  // view of one sample's row of `coef`
  float gamma, contrast;
  const float* add;
  const float* factor;
};
__device__ __forceinline__ AugCoef coef_of(const float* coef, int b, int C) {
  const float* r = coef + (size_t)b * (2 + 2 * C);
  return {r[0], r[1], r + 2, r + 2 + C};
}

// partials layout: [stage][b][chunk][c][3] = (min, max, sum)
__device__ __forceinline__ float* part_at(float* ws, int stage, int b, int chunk, int c, int B, int nch, int C) {
  return ws + ((((size_t)stage * B + b) * nch + chunk) * C + c) * 3;
}

// gamma window of sample b: minimum and range over ALL channels of v (the CPU code takes them over the whole sample)
__device__ __forceinline__ void gamma_window(const float* ws, int b, int B, int nch, int C, float& minm, float& rnge) {
  float lo = INFINITY, hi = -INFINITY;
  for (int ch = 0; ch < nch; ++ch)
    for (int c = 0; c < C; ++c) {
      const float* q = part_at(const_cast<float*>(ws), 0, b, ch, c, B, nch, C);
      lo = fminf(lo, q[0]);
      hi = fmaxf(hi, q[1]);
    }
  minm = lo;
  rnge = __fsub_rn(hi, lo);
}

// one rounding per operation, as the array library evaluates it (no fused multiply-add)
__device__ __forceinline__ float gamma_curve(float v, float gamma, float minm, float rnge) {
  const float t = __fdiv_rn(__fsub_rn(v, minm), __fadd_rn(rnge, AUG_EPS));
  return __fadd_rn(__fmul_rn(powf(t, gamma), rnge), minm);
}

template <int STAGE>
__global__ void __launch_bounds__(AUG_THREADS) aug_stats_kernel(const float* __restrict__ x, const float* __restrict__ coef,
                                                                float* __restrict__ ws, int B, int C, int64_t V, int nch) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const AugCoef k = coef_of(coef, b, C);
  __shared__ float s_win[2];
  __shared__ float s_red[AUG_THREADS / 32][AUG_MAXC][3];
  if (STAGE == 1) {
    if (threadIdx.x == 0) gamma_window(ws, b, B, nch, C, s_win[0], s_win[1]);
    __syncthreads();
  }
  const bool use_gamma = STAGE == 1 && k.gamma > 0.f;
  float add[AUG_MAXC], lo[AUG_MAXC], hi[AUG_MAXC], sum[AUG_MAXC];
#pragma unroll
  for (int c = 0; c < AUG_MAXC; ++c) {
    add[c] = c < C ? k.add[c] : 0.f;
    lo[c] = INFINITY, hi[c] = -INFINITY, sum[c] = 0.f;
  }
  const int64_t v0 = V * chunk / nch, v1 = V * (chunk + 1) / nch;
  const float* xb = x + (size_t)b * V * C;
  for (int64_t i = v0 + threadIdx.x; i < v1; i += AUG_THREADS) {
#pragma unroll
    for (int c = 0; c < AUG_MAXC; ++c)
      if (c < C) {
        float v = __fadd_rn(xb[i * C + c], add[c]);
        if (use_gamma) v = gamma_curve(v, k.gamma, s_win[0], s_win[1]);
        lo[c] = fminf(lo[c], v);
        hi[c] = fmaxf(hi[c], v);
        sum[c] += v;
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < AUG_MAXC; ++c)
    if (c < C) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
        hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], o);
      }
      if (lane == 0) s_red[warp][c][0] = lo[c], s_red[warp][c][1] = hi[c], s_red[warp][c][2] = sum[c];
    }
  __syncthreads();
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    float l = INFINITY, h = -INFINITY, s = 0.f;
    for (int w = 0; w < AUG_THREADS / 32; ++w) l = fminf(l, s_red[w][c][0]), h = fmaxf(h, s_red[w][c][1]), s += s_red[w][c][2];
    float* q = part_at(ws, STAGE, b, chunk, c, B, nch, C);
    q[0] = l, q[1] = h, q[2] = s;
  }
}

template <typename TD>
__global__ void __launch_bounds__(AUG_THREADS) aug_apply_kernel(const float* __restrict__ x, TD* __restrict__ y,
                                                                const float* __restrict__ coef, const float* __restrict__ ws,
                                                                int B, int C, int64_t V, int nch) {
  const int b = blockIdx.y;
  const AugCoef k = coef_of(coef, b, C);
  __shared__ float s_win[2];
  __shared__ float s_ch[AUG_MAXC][3];                        // min, max, mean of g per channel
  if (threadIdx.x == 0) gamma_window(ws, b, B, nch, C, s_win[0], s_win[1]);
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    float l = INFINITY, h = -INFINITY, s = 0.f;
    for (int ch = 0; ch < nch; ++ch) {
      const float* q = part_at(const_cast<float*>(ws), 1, b, ch, c, B, nch, C);
      l = fminf(l, q[0]), h = fmaxf(h, q[1]), s += q[2];
    }
    s_ch[c][0] = l, s_ch[c][1] = h, s_ch[c][2] = __fdiv_rn(s, (float)V);
  }
  __syncthreads();
  const bool use_gamma = k.gamma > 0.f, use_contrast = k.contrast != 0.f;
  const float minm = s_win[0], rnge = s_win[1];
  const float* xb = x + (size_t)b * V * C;
  TD* yb = y + (size_t)b * V * C;
  const int64_t total = V * C;
  for (int64_t i = (int64_t)blockIdx.x * AUG_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * AUG_THREADS) {
    const int c = (int)(i % C);
    float v = __fadd_rn(xb[i], k.add[c]);
    if (use_gamma) v = gamma_curve(v, k.gamma, minm, rnge);
    if (use_contrast) {
      const float mn = s_ch[c][2];
      v = __fadd_rn(__fmul_rn(__fsub_rn(v, mn), k.factor[c]), mn);
      v = fminf(fmaxf(v, s_ch[c][0]), s_ch[c][1]);
    }
    yb[i] = from_f32<TD>(v);
  }
}

static int aug_chunks(int64_t V) {
  int64_t n = V / 2048;
  return (int)(n < 1 ? 1 : n > 64 ? 64 : n);
}

}  // namespace mednet

using namespace mednet;

extern "C" size_t mednet_intensity_augment_workspace_bytes(const mednet_intensity_aug_params* p) {
  if (!p || p->B <= 0 || p->C <= 0 || p->V <= 0) return 0;
  return (size_t)2 * p->B * aug_chunks(p->V) * p->C * 3 * sizeof(float);
}

extern "C" int mednet_intensity_augment(const mednet_intensity_aug_params* p, void* workspace, size_t workspace_bytes,
                                        mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->x && p->y && p->coef && workspace && p->B > 0 && p->C > 0 && p->V > 0, MEDNET_EINVAL);
  MEDNET_REQUIRE(workspace_bytes >= mednet_intensity_augment_workspace_bytes(p), MEDNET_EWORKSPACE);
  float* ws = (float*)workspace;
  MEDNET_REQUIRE(p->C <= AUG_MAXC, MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(p->dst_dtype == MEDNET_F32 || p->dst_dtype == MEDNET_BF16, MEDNET_EUNSUPPORTED);
  const int nch = aug_chunks(p->V);
  const dim3 sgrid(nch, p->B);
  aug_stats_kernel<0><<<sgrid, AUG_THREADS, 0, stream>>>(p->x, p->coef, ws, p->B, p->C, p->V, nch);
  MEDNET_LAUNCH_CHECK();
  aug_stats_kernel<1><<<sgrid, AUG_THREADS, 0, stream>>>(p->x, p->coef, ws, p->B, p->C, p->V, nch);
  MEDNET_LAUNCH_CHECK();
  const dim3 agrid(grid_for(p->V * p->C, AUG_THREADS, 4), p->B);
  if (p->dst_dtype == MEDNET_F32)
    aug_apply_kernel<float><<<agrid, AUG_THREADS, 0, stream>>>(p->x, (float*)p->y, p->coef, ws, p->B, p->C, p->V, nch);
  else
    aug_apply_kernel<bf16><<<agrid, AUG_THREADS, 0, stream>>>(p->x, (bf16*)p->y, p->coef, ws, p->B, p->C, p->V, nch);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}
