// Parameter gradients of a per-channel affine that feeds a 3x3x3 convolution WITHOUT the convolution's data gradient.
// ref: the first layer of the 'gcr' networks, midasmednet/unet/components.py:45-57 (GroupNorm on the image) followed by
// :8-9 (Conv3d): the image needs no gradient, so the only consumers of the conv's input gradient dxn are the two
// per-channel sums of the GroupNorm backward, dbeta = sum dxn and dgamma = sum dxn * xhat.  With xn = gamma * xhat + beta,
// Wb the weights as the conv used them and Ghat = wgrad(dpre, xhat), B[co][t] = sum of dpre over the voxels whose tap t
// lies inside the volume (= wgrad(dpre, 1)), the adjoint identity <dgrad(dpre), u> = <dpre, conv(u)> gives, exactly,
//     dgamma[ci] = sum_{co,t} Wb[co][ci][t] * Ghat[co][ci][t]        dbeta[ci] = sum_{co,t} Wb[co][ci][t] * B[co][t]
//     dW[co][ci][t] = gamma[ci] * Ghat[co][ci][t] + beta[ci] * B[co][t]
// so the 32 -> 1 channel dgrad (a full read of the largest gradient tensor into a CUDA-core stencil, 1.7 ms per cfg-3
// step) and the GroupNorm backward passes of that layer disappear; B comes from one streaming pass that bins dpre by
// border class (first / interior / last plane per axis, 27 classes).
#include "common.cuh"

namespace mednet {

// partial[(n * D + d)][b * 3 + c][C]: sums of the (n, d) plane over rows of h-class b and voxels of w-class c
// (class 0 interior, 1 first, 2 last).  Block (C / V, R): thread (tx, ty) owns channel vector tx and voxels w = ty, ty+R, ..
template <typename T, int V>
__global__ void border_class_partial_kernel(const T* __restrict__ x, float* __restrict__ partial, int D, int H, int W, int C) {
  extern __shared__ float sm[];
  const int n = blockIdx.y, d = blockIdx.x;
  const int tx = threadIdx.x, ty = threadIdx.y, R = blockDim.y, bx = blockDim.x;
  const T* plane = x + (((int64_t)n * D + d) * H) * (int64_t)W * C + (int64_t)tx * V;
  float acc[3][3][V];
#pragma unroll
  for (int b = 0; b < 3; ++b)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[b][c][i] = 0.f;
  // UR rows per iteration: all their 16-byte loads are issued before the first add (memory-level parallelism)
  constexpr int UR = 4;
  auto rows = [&](int h0, int h1, float (&a)[3][V]) {
    for (int h = h0; h < h1; h += UR) {
      for (int w = ty; w < W; w += R) {
        typename RawVec<sizeof(T) * V>::type raw[UR];
#pragma unroll
        for (int u = 0; u < UR; ++u)
          if (h + u < h1) raw[u] = load_raw<T, V>(plane + ((int64_t)(h + u) * W + w) * C);
        const int cls = w == 0 ? 1 : (w == W - 1 ? 2 : 0);
#pragma unroll
        for (int u = 0; u < UR; ++u)
          if (h + u < h1) {
            float v[V];
            cvt_raw<T, V>(raw[u], v);
            if (cls == 0) {
#pragma unroll
              for (int i = 0; i < V; ++i) a[0][i] += v[i];
            } else if (cls == 1) {
#pragma unroll
              for (int i = 0; i < V; ++i) a[1][i] += v[i];
            } else {
#pragma unroll
              for (int i = 0; i < V; ++i) a[2][i] += v[i];
            }
          }
      }
    }
  };
  rows(0, 1, acc[1]);
  rows(1, H - 1, acc[0]);
  rows(H - 1, H, acc[2]);
  float* out = partial + ((int64_t)n * D + d) * 9 * C;
  // reduce over ty through shared memory, one (b, c) class at a time (fixed order: deterministic)
#pragma unroll
  for (int b = 0; b < 3; ++b)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < V; ++i) sm[(ty * bx + tx) * V + i] = acc[b][c][i];
      __syncthreads();
      for (int i = ty; i < V; i += R) {
        float s = 0.f;
        for (int t = 0; t < R; ++t) s += sm[(t * bx + tx) * V + i];
        out[(b * 3 + c) * C + tx * V + i] = s;
      }
    }
}

// bins[(a * 3 + b) * 3 + c][C] = sum over n and over the d-planes of class a
__global__ void border_class_final_kernel(const float* __restrict__ partial, float* __restrict__ bins, int N, int D, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 27 * C) return;
  const int ch = i % C, bc = (i / C) % 9, a = i / (9 * C);
  const int d0 = a == 0 ? 1 : (a == 1 ? 0 : D - 1), d1 = a == 0 ? D - 1 : d0 + 1;
  double s = 0.0;
  for (int n = 0; n < N; ++n)
    for (int d = d0; d < d1; ++d) s += (double)partial[(((int64_t)n * D + d) * 9 + bc) * C + ch];
  bins[i] = (float)s;
}

// one block per input channel ci
__global__ void affine_input_grads_kernel(const float* __restrict__ w, const float* __restrict__ ghat,
                                          const float* __restrict__ bins, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, float* __restrict__ dw, float* __restrict__ dgamma,
                                          float* __restrict__ dbeta, int Cout, int Cin, int round_bf16) {
  __shared__ double scratch[32];
  const int ci = blockIdx.x;
  const float g = gamma[ci], bt = beta[ci];
  double sg = 0.0, sb = 0.0;
  for (int i = threadIdx.x; i < Cout * 27; i += blockDim.x) {
    const int co = i / 27, t = i - co * 27;
    const int td = t / 9, th = (t / 3) % 3, tw = t % 3;
    // tap offset -1 (index 0) needs v' >= 1: the FIRST plane (class 1) drops out; offset +1: the LAST plane (class 2)
    float B = 0.f;
    for (int a = 0; a < 3; ++a) {
      if ((td == 0 && a == 1) || (td == 2 && a == 2)) continue;
      for (int b = 0; b < 3; ++b) {
        if ((th == 0 && b == 1) || (th == 2 && b == 2)) continue;
        for (int c = 0; c < 3; ++c) {
          if ((tw == 0 && c == 1) || (tw == 2 && c == 2)) continue;
          B += bins[((a * 3 + b) * 3 + c) * Cout + co];
        }
      }
    }
    const int64_t idx = ((int64_t)co * Cin + ci) * 27 + t;
    float wv = w[idx];
    if (round_bf16) wv = __bfloat162float(__float2bfloat16_rn(wv));      // the weights as the bf16 convolution used them
    const float gh = ghat[idx];
    dw[idx] = fmaf(g, gh, bt * B);
    sg += (double)wv * (double)gh;
    sb += (double)wv * (double)B;
  }
  sg = block_sum(sg, scratch);
  sb = block_sum(sb, scratch);
  if (threadIdx.x == 0) {
    dgamma[ci] = (float)sg;
    dbeta[ci] = (float)sb;
  }
}

static bool border_plan(const mednet_border_sums_params* p, int* V, int* ncol, int* R) {
  if (!p || !dtype_ok(p->dtype) || p->N <= 0 || p->C <= 0 || p->D < 2 || p->H < 2 || p->W < 2) return false;
  *V = pick_vec(p->C, dtype_bytes(p->dtype));
  *ncol = p->C / *V;
  if (*ncol > 256 || p->N > 65535) return false;
  *R = 256 / *ncol;
  return true;
}

}  // namespace mednet

using namespace mednet;

extern "C" size_t mednet_border_class_sums_workspace_bytes(const mednet_border_sums_params* p) {
  int V, ncol, R;
  if (!border_plan(p, &V, &ncol, &R)) return 0;
  return align_up((size_t)p->N * p->D * 9 * p->C * sizeof(float), 256);
}

extern "C" int mednet_border_class_sums(const mednet_border_sums_params* p, void* workspace, size_t workspace_bytes,
                                        mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->x && p->bins, MEDNET_EINVAL);
  int V, ncol, R;
  MEDNET_REQUIRE(border_plan(p, &V, &ncol, &R), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_border_class_sums_workspace_bytes(p), MEDNET_EWORKSPACE);
  float* partial = (float*)workspace;
  dim3 grid(p->D, p->N), block(ncol, R);
  MEDNET_DISPATCH_TV(p->dtype, V, {
    const size_t smem = (size_t)ncol * R * VV * sizeof(float);
    border_class_partial_kernel<T, VV><<<grid, block, smem, stream>>>((const T*)p->x, partial, p->D, p->H, p->W, p->C);
  });
  MEDNET_LAUNCH_CHECK();
  border_class_final_kernel<<<ceil_div(27 * p->C, 128), 128, 0, stream>>>(partial, p->bins, p->N, p->D, p->C);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_conv3d_affine_input_grads(const mednet_affine_input_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->w && p->ghat && p->bins && p->gamma && p->beta && p->dw && p->dgamma && p->dbeta, MEDNET_EINVAL);
  MEDNET_REQUIRE(p->Cout > 0 && p->Cin > 0 && dtype_ok(p->dtype), MEDNET_EINVAL);
  affine_input_grads_kernel<<<p->Cin, 256, 0, stream>>>(p->w, p->ghat, p->bins, p->gamma, p->beta, p->dw, p->dgamma,
                                                       p->dbeta, p->Cout, p->Cin, p->dtype == MEDNET_BF16 ? 1 : 0);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}
