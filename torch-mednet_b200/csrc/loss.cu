// Fused losses on NCDHW logits: softmax/sigmoid + Dice, weighted cross entropy, heatmap MSE/L1,
// each with its gradient, plus the uint8 inference epilogue.
//   Dice .......... ref: midasmednet/unet/loss.py:114-130 (+ :10-48 flatten/per-channel dice, :58-88 one-hot)
//   CE ............ ref: midasmednet/segmentation.py:49, landmarks.py:49
//   heatmap loss .. ref: midasmednet/landmarks.py:53-55,125-134
//   epilogue ...... ref: examples/predict.py:88-94
// The one-hot tensor, the (C, N*DHW) transposed copies and the python loop over heatmap channels of the
// reference never materialise: one pass over the logits forward, one pass backward.  Per-voxel threads
// read channel planes (stride S) so every warp access is coalesced.  Reductions: warp shuffle -> block ->
// per-block partials -> fixed-order final sum (deterministic).
#include "common.cuh"

namespace mednet {

constexpr int MAXC = 32;

template <typename TL>
__device__ __forceinline__ int load_label(const void* labels, int64_t i);
template <> __device__ __forceinline__ int load_label<int64_t>(const void* l, int64_t i) { return (int)((const int64_t*)l)[i]; }
template <> __device__ __forceinline__ int load_label<uint8_t>(const void* l, int64_t i) { return (int)((const uint8_t*)l)[i]; }

// probabilities of one voxel into p[0..C)
template <typename T, int CM>
__device__ __forceinline__ void voxel_probs(const T* __restrict__ base, int64_t S, int C, int sigmoid, float (&p)[CM]) {
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < CM; ++c) {
    if (c < C) {
      p[c] = to_f32<T>(base[(int64_t)c * S]);
      mx = fmaxf(mx, p[c]);
    }
  }
  if (sigmoid) {
#pragma unroll
    for (int c = 0; c < CM; ++c)
      if (c < C) p[c] = 1.f / (1.f + expf(-p[c]));
  } else {
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CM; ++c)
      if (c < C) {
        p[c] = expf(p[c] - mx);
        sum += p[c];
      }
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < CM; ++c)
      if (c < C) p[c] *= inv;
  }
}

// same, from logits already held in registers (lets a kernel batch the loads of several voxels first)
template <int CM>
__device__ __forceinline__ void probs_inplace(int C, int sigmoid, float (&p)[CM]) {
  if (sigmoid) {
#pragma unroll
    for (int c = 0; c < CM; ++c)
      if (c < C) p[c] = 1.f / (1.f + expf(-p[c]));
    return;
  }
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < CM; ++c)
    if (c < C) mx = fmaxf(mx, p[c]);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < CM; ++c)
    if (c < C) {
      p[c] = expf(p[c] - mx);
      sum += p[c];
    }
  const float inv = 1.f / sum;
#pragma unroll
  for (int c = 0; c < CM; ++c)
    if (c < C) p[c] *= inv;
}

// voxel index -> (sample, offset) without a 64-bit division when the tensor is small enough
__device__ __forceinline__ void split_index(int64_t i, int64_t S, int64_t& n, int64_t& s) {
  if (i < 0x7fffffffLL && S < 0x7fffffffLL) {
    const uint32_t q = (uint32_t)i / (uint32_t)S;
    n = q;
    s = (int64_t)((uint32_t)i - q * (uint32_t)S);
  } else {
    n = i / S;
    s = i - n * S;
  }
}

constexpr int LOSS_UNR = 4;   // voxels per thread and iteration: their logits / labels are loaded before any arithmetic

// block-reduce NV values; thread 0 writes them to out[0..NV)
template <int NV>
__device__ __forceinline__ void block_reduce_store(float (&v)[NV], int nvalid, float* out) {
  __shared__ float sm[32][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (i < nvalid) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nvalid) sm[warp][i] = v[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nvalid; i += blockDim.x) {
    float a = 0.f;
    for (int w = 0; w < nw; ++w) a += sm[w][i];
    out[i] = a;
  }
}

// ---------------------------------------------------------------- Dice
template <typename T, typename TL, int CM>
__global__ void dice_partial_kernel(const T* __restrict__ logits, const void* __restrict__ labels,
                                    float* __restrict__ partial, int64_t N, int64_t S, int64_t bstride, int C,
                                    int sigmoid) {
  float acc[3 * CM];
#pragma unroll
  for (int i = 0; i < 3 * CM; ++i) acc[i] = 0.f;
  const int64_t total = N * S, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += LOSS_UNR * stride) {
    float p[LOSS_UNR][CM];
    int y[LOSS_UNR];
#pragma unroll
    for (int u = 0; u < LOSS_UNR; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total) {
        int64_t n, sv;
        split_index(i, S, n, sv);
        const T* base = logits + n * bstride + sv;
#pragma unroll
        for (int c = 0; c < CM; ++c)
          if (c < C) p[u][c] = to_f32<T>(base[(int64_t)c * S]);
        y[u] = load_label<TL>(labels, i);
      }
    }
#pragma unroll
    for (int u = 0; u < LOSS_UNR; ++u) {
      if (i0 + u * stride < total) {
        probs_inplace<CM>(C, sigmoid, p[u]);
#pragma unroll
        for (int c = 0; c < CM; ++c) {
          if (c < C) {
            const float t = (y[u] == c) ? 1.f : 0.f;
            acc[c] += p[u][c] * t;
            acc[CM + c] += p[u][c];
            acc[2 * CM + c] += t;
          }
        }
      }
    }
  }
  block_reduce_store<3 * CM>(acc, 3 * CM, partial + (int64_t)blockIdx.x * 3 * CM);
}

// sums[3][C], dice[C], loss
__global__ void dice_final_kernel(const float* __restrict__ partial, const float* __restrict__ weight,
                                  float* __restrict__ sums, float* __restrict__ dice, float* __restrict__ loss,
                                  int C, int CM, int nblocks, float eps) {
  __shared__ float s_d[MAXC];
  const int c = threadIdx.x;
  if (c < C) {
    double v[3];
    for (int k = 0; k < 3; ++k) {
      double a = 0.0;
      for (int b = 0; b < nblocks; ++b) a += (double)partial[((int64_t)b * 3 + k) * CM + c];
      v[k] = a;
      sums[k * C + c] = (float)a;
    }
    const float w = weight ? weight[c] : 1.f;
    const float den = fmaxf((float)(v[1] + v[2]), eps);
    const float d = 2.f * (w * (float)v[0]) / den;
    dice[c] = d;
    s_d[c] = 1.f - d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int k = 0; k < C; ++k) a += s_d[k];
    loss[0] = a / (float)C;
  }
}

template <typename T, typename TL, typename TO, int CM>
__global__ void dice_bwd_kernel(const T* __restrict__ logits, const void* __restrict__ labels,
                                const float* __restrict__ weight, const float* __restrict__ sums,
                                const float* __restrict__ grad_out, TO* __restrict__ dlogits, int64_t N, int64_t S,
                                int64_t bstride, int64_t bstride_out, int C, int sigmoid, float eps) {
  // dL/dp_c(v) = a_c * t_c(v) + b_c with
  //   a_c = -(1/C) * 2 w_c / U_c,  b_c = (1/C) * 2 w_c I_c / U_c^2 if the clamp is inactive, else 0
  __shared__ float sa[MAXC], sb[MAXC];
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    const float w = weight ? weight[c] : 1.f;
    const float den = sums[C + c] + sums[2 * C + c];
    const float U = fmaxf(den, eps);
    const float go = grad_out[0] / (float)C;
    sa[c] = -go * 2.f * w / U;
    sb[c] = (den >= eps) ? go * 2.f * w * sums[c] / (U * U) : 0.f;
  }
  __syncthreads();
  const int64_t total = N * S, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += LOSS_UNR * stride) {
    float p[LOSS_UNR][CM];
    int y[LOSS_UNR];
    int64_t off_out[LOSS_UNR];
#pragma unroll
    for (int u = 0; u < LOSS_UNR; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total) {
        int64_t n, sv;
        split_index(i, S, n, sv);
        const T* base = logits + n * bstride + sv;
        off_out[u] = n * bstride_out + sv;
#pragma unroll
        for (int c = 0; c < CM; ++c)
          if (c < C) p[u][c] = to_f32<T>(base[(int64_t)c * S]);
        y[u] = load_label<TL>(labels, i);
      }
    }
#pragma unroll
    for (int u = 0; u < LOSS_UNR; ++u) {
      if (i0 + u * stride < total) {
        probs_inplace<CM>(C, sigmoid, p[u]);
        float g[CM];
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < CM; ++c) {
          if (c < C) {
            g[c] = ((y[u] == c) ? sa[c] : 0.f) + sb[c];
            dot += p[u][c] * g[c];
          }
        }
        TO* out = dlogits + off_out[u];
#pragma unroll
        for (int c = 0; c < CM; ++c) {
          if (c < C) {
            const float dz = sigmoid ? g[c] * p[u][c] * (1.f - p[u][c]) : p[u][c] * (g[c] - dot);
            out[(int64_t)c * S] = from_f32<TO>(dz);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------- weighted cross entropy
template <typename T, typename TL, int CM>
__global__ void ce_partial_kernel(const T* __restrict__ logits, const void* __restrict__ labels,
                                  const float* __restrict__ weight, float* __restrict__ partial, int64_t N, int64_t S,
                                  int64_t bstride, int C) {
  float acc[2] = {0.f, 0.f};
  const int64_t total = N * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / S, s = i - n * S;
    const T* base = logits + n * bstride + s;
    const int y = load_label<TL>(labels, i);
    float mx = -INFINITY, zy = 0.f;
    float z[CM];
#pragma unroll
    for (int c = 0; c < CM; ++c)
      if (c < C) {
        z[c] = to_f32<T>(base[(int64_t)c * S]);
        mx = fmaxf(mx, z[c]);
        if (c == y) zy = z[c];
      }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CM; ++c)
      if (c < C) sum += expf(z[c] - mx);
    const float nll = logf(sum) + mx - zy;
    // a label outside [0, C) is a caller error (nn.CrossEntropyLoss device-asserts): never index weight[] with it, and
    // poison the loss so that it cannot pass silently
    const float w = ((unsigned)y < (unsigned)C) ? (weight ? weight[y] : 1.f) : NAN;
    acc[0] += w * nll;
    acc[1] += w;
  }
  block_reduce_store<2>(acc, 2, partial + (int64_t)blockIdx.x * 2);
}

__global__ void ce_final_kernel(const float* __restrict__ partial, float* __restrict__ sums, float* __restrict__ loss,
                                int nblocks) {
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < nblocks; ++i) {
      a += (double)partial[2 * i];
      b += (double)partial[2 * i + 1];
    }
    sums[0] = (float)a;
    sums[1] = (float)b;
    loss[0] = (float)(a / b);
  }
}

template <typename T, typename TL, typename TO, int CM>
__global__ void ce_bwd_kernel(const T* __restrict__ logits, const void* __restrict__ labels,
                              const float* __restrict__ weight, const float* __restrict__ sums,
                              const float* __restrict__ grad_out, TO* __restrict__ dlogits, int64_t N, int64_t S,
                              int64_t bstride, int64_t bstride_out, int C) {
  const float scale = grad_out[0] / sums[1];
  const int64_t total = N * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / S, s = i - n * S;
    float p[CM];
    voxel_probs<T, CM>(logits + n * bstride + s, S, C, 0, p);
    const int y = load_label<TL>(labels, i);
    const float w = (((unsigned)y < (unsigned)C) ? (weight ? weight[y] : 1.f) : NAN) * scale;
    TO* out = dlogits + n * bstride_out + s;
#pragma unroll
    for (int c = 0; c < CM; ++c)
      if (c < C) out[(int64_t)c * S] = from_f32<TO>(w * (p[c] - ((c == y) ? 1.f : 0.f)));
  }
}

// ---------------------------------------------------------------- heatmap MSE / L1
// grid (chunks, L, N): partial[(n*L + c)*chunks + chunk] = sum of per-element error over the chunk
template <typename T, typename TT>
__global__ void hm_partial_kernel(const T* __restrict__ pred, const TT* __restrict__ target,
                                  float* __restrict__ partial, int64_t S, int64_t bstride, int L, int l1,
                                  int64_t per_chunk) {
  __shared__ float scratch[32];
  const int c = blockIdx.y, n = blockIdx.z;
  const T* p = pred + (int64_t)n * bstride + (int64_t)c * S;
  const TT* t = target + ((int64_t)n * L + c) * S;
  const int64_t s0 = (int64_t)blockIdx.x * per_chunk;
  int64_t s1 = s0 + per_chunk;
  if (s1 > S) s1 = S;
  float acc = 0.f;
  for (int64_t s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
    const float d = to_f32<T>(p[s]) - to_f32<TT>(t[s]);
    acc += l1 ? fabsf(d) : d * d;
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) partial[((int64_t)n * L + c) * gridDim.x + blockIdx.x] = acc;
}

__global__ void hm_final_kernel(const float* __restrict__ partial, const float* __restrict__ weight,
                                float* __restrict__ per_channel, float* __restrict__ loss, int64_t N, int64_t S,
                                int L, int chunks) {
  __shared__ float s_v[64];
  const int c = threadIdx.x;
  if (c < L) {
    double a = 0.0;
    for (int64_t n = 0; n < N; ++n)
      for (int k = 0; k < chunks; ++k) a += (double)partial[(n * L + c) * chunks + k];
    const float m = (float)(a / ((double)N * (double)S));
    per_channel[c] = m;
    s_v[c] = weight[c] * m;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;   // sequential += in channel order, as the reference's python loop (landmarks.py:128-132)
    for (int k = 0; k < L; ++k) a += s_v[k];
    loss[0] = a;
  }
}

template <typename T, typename TT, typename TO>
__global__ void hm_bwd_kernel(const T* __restrict__ pred, const TT* __restrict__ target,
                              const float* __restrict__ weight, const float* __restrict__ grad_out,
                              TO* __restrict__ dpred, int64_t N, int64_t S, int64_t bstride, int64_t bstride_out,
                              int L, int l1) {
  const int c = blockIdx.y, n = blockIdx.z;
  const T* p = pred + (int64_t)n * bstride + (int64_t)c * S;
  const TT* t = target + ((int64_t)n * L + c) * S;
  TO* o = dpred + (int64_t)n * bstride_out + (int64_t)c * S;
  const float k = grad_out[0] * weight[c] / ((float)N * (float)S);
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
    const float d = to_f32<T>(p[s]) - to_f32<TT>(t[s]);
    const float g = l1 ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : 2.f * d;
    o[s] = from_f32<TO>(k * g);
  }
}

// ---------------------------------------------------------------- inference epilogue
template <typename T>
__global__ void predict_epilogue_kernel(const T* __restrict__ logits, uint8_t* __restrict__ out, int64_t N, int64_t S,
                                        int L, int K) {
  const int64_t total = N * S;
  const int C = L + K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / S, s = i - n * S;
    const T* base = logits + n * C * S + s;
    uint8_t* o = out + n * (L + 1) * S + s;
    for (int c = 0; c < L; ++c) {
      float v = to_f32<T>(base[(int64_t)c * S]);
      v = fminf(fmaxf(v, 0.f), 255.f);          // np.clip; NaN -> 0 (fmaxf drops the NaN operand)
      o[(int64_t)c * S] = (uint8_t)(int)v;      // astype(uint8): truncation toward zero
    }
    float best = to_f32<T>(base[(int64_t)L * S]);
    int arg = 0;
    for (int c = 1; c < K; ++c) {
      const float v = to_f32<T>(base[(int64_t)(L + c) * S]);
      if (v > best || (v != v && best == best)) {   // first maximum wins; NaN is maximal (torch.argmax)
        best = v;
        arg = c;
      }
    }
    o[(int64_t)L * S] = (uint8_t)arg;
  }
}

// ---------------------------------------------------------------- test-time activation (model.py:107-108)
template <int CM>
__global__ void final_activation_kernel(const float* __restrict__ logits, float* __restrict__ out, int64_t N, int64_t S,
                                        int C, int sigmoid) {
  const int64_t total = N * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / S, s = i - n * S;
    float p[CM];
    voxel_probs<float, CM>(logits + n * C * S + s, S, C, sigmoid, p);
#pragma unroll
    for (int c = 0; c < CM; ++c)
      if (c < C) out[(n * C + c) * S + s] = p[c];
  }
}

static int loss_blocks(int64_t total) {
  int b = grid_for(total, 256, 4);
  return b;
}

#define MEDNET_DISPATCH_CM(C, ...)                         \
  do {                                                     \
    if ((C) <= 4) { constexpr int CM = 4; __VA_ARGS__; }   \
    else if ((C) <= 8) { constexpr int CM = 8; __VA_ARGS__; } \
    else if ((C) <= 16) { constexpr int CM = 16; __VA_ARGS__; } \
    else { constexpr int CM = 32; __VA_ARGS__; }           \
  } while (0)

#define MEDNET_DISPATCH_LOGITS(dt, ...)                         \
  do {                                                          \
    if ((dt) == MEDNET_F32) { typedef float T; __VA_ARGS__; }   \
    else { typedef bf16 T; __VA_ARGS__; }                       \
  } while (0)

#define MEDNET_DISPATCH_LABEL(dt, ...)                              \
  do {                                                              \
    if ((dt) == MEDNET_I64) { typedef int64_t TL; __VA_ARGS__; }    \
    else { typedef uint8_t TL; __VA_ARGS__; }                       \
  } while (0)

static inline bool label_ok(int dt) { return dt == MEDNET_I64 || dt == MEDNET_U8; }
static inline int cm_of(int C) { return C <= 4 ? 4 : C <= 8 ? 8 : C <= 16 ? 16 : 32; }

}  // namespace mednet

using namespace mednet;

extern "C" size_t mednet_dice_workspace_bytes(const mednet_dice_params* p) {
  if (!p) return 0;
  return align_up((size_t)loss_blocks(p->N * p->S) * 3 * cm_of(p->C) * sizeof(float), 256);
}

extern "C" int mednet_dice_fwd(const mednet_dice_params* p, void* workspace, size_t workspace_bytes,
                               mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->logits && p->labels && p->sums && p->dice && p->loss, MEDNET_EINVAL);
  MEDNET_REQUIRE(p->N > 0 && p->S > 0 && p->C > 0 && p->C <= MAXC, MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(dtype_ok(p->logits_dtype) && label_ok(p->label_dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_dice_workspace_bytes(p), MEDNET_EWORKSPACE);
  const int nb = loss_blocks(p->N * p->S);
  float* partial = (float*)workspace;
  MEDNET_DISPATCH_LOGITS(p->logits_dtype, MEDNET_DISPATCH_LABEL(p->label_dtype, MEDNET_DISPATCH_CM(p->C, {
    dice_partial_kernel<T, TL, CM><<<nb, 256, 0, stream>>>((const T*)p->logits, p->labels, partial, p->N, p->S,
                                                          p->batch_stride, p->C, p->sigmoid);
  })));
  MEDNET_LAUNCH_CHECK();
  dice_final_kernel<<<1, MAXC, 0, stream>>>(partial, p->weight, p->sums, p->dice, p->loss, p->C, cm_of(p->C), nb,
                                           p->epsilon);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

#define MEDNET_DISPATCH_OUT(dt, ...)                             \
  do {                                                           \
    if ((dt) == MEDNET_F32) { typedef float TO; __VA_ARGS__; }   \
    else { typedef bf16 TO; __VA_ARGS__; }                       \
  } while (0)

extern "C" int mednet_dice_bwd(const mednet_dice_bwd_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->logits && p->labels && p->sums && p->grad_out && p->dlogits, MEDNET_EINVAL);
  MEDNET_REQUIRE(p->N > 0 && p->S > 0 && p->C > 0 && p->C <= MAXC, MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(dtype_ok(p->logits_dtype) && dtype_ok(p->dlogits_dtype) && label_ok(p->label_dtype),
                 MEDNET_EUNSUPPORTED);
  const int nb = grid_for(p->N * p->S, 256, 8);
  MEDNET_DISPATCH_LOGITS(p->logits_dtype, MEDNET_DISPATCH_LABEL(p->label_dtype, MEDNET_DISPATCH_OUT(p->dlogits_dtype,
    MEDNET_DISPATCH_CM(p->C, {
      dice_bwd_kernel<T, TL, TO, CM><<<nb, 256, 0, stream>>>((const T*)p->logits, p->labels, p->weight, p->sums,
                                                            p->grad_out, (TO*)p->dlogits, p->N, p->S, p->batch_stride,
                                                            p->batch_stride_out, p->C, p->sigmoid, p->epsilon);
    }))));
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" size_t mednet_ce_workspace_bytes(const mednet_ce_params* p) {
  if (!p) return 0;
  return align_up((size_t)loss_blocks(p->N * p->S) * 2 * sizeof(float), 256);
}

extern "C" int mednet_ce_fwd(const mednet_ce_params* p, void* workspace, size_t workspace_bytes, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->logits && p->labels && p->sums && p->loss, MEDNET_EINVAL);
  MEDNET_REQUIRE(p->N > 0 && p->S > 0 && p->C > 0 && p->C <= MAXC, MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(dtype_ok(p->logits_dtype) && label_ok(p->label_dtype), MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_ce_workspace_bytes(p), MEDNET_EWORKSPACE);
  const int nb = loss_blocks(p->N * p->S);
  float* partial = (float*)workspace;
  MEDNET_DISPATCH_LOGITS(p->logits_dtype, MEDNET_DISPATCH_LABEL(p->label_dtype, MEDNET_DISPATCH_CM(p->C, {
    ce_partial_kernel<T, TL, CM><<<nb, 256, 0, stream>>>((const T*)p->logits, p->labels, p->weight, partial, p->N, p->S,
                                                        p->batch_stride, p->C);
  })));
  MEDNET_LAUNCH_CHECK();
  ce_final_kernel<<<1, 32, 0, stream>>>(partial, p->sums, p->loss, nb);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_ce_bwd(const mednet_ce_bwd_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->logits && p->labels && p->sums && p->grad_out && p->dlogits, MEDNET_EINVAL);
  MEDNET_REQUIRE(p->N > 0 && p->S > 0 && p->C > 0 && p->C <= MAXC, MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(dtype_ok(p->logits_dtype) && dtype_ok(p->dlogits_dtype) && label_ok(p->label_dtype),
                 MEDNET_EUNSUPPORTED);
  const int nb = grid_for(p->N * p->S, 256, 8);
  MEDNET_DISPATCH_LOGITS(p->logits_dtype, MEDNET_DISPATCH_LABEL(p->label_dtype, MEDNET_DISPATCH_OUT(p->dlogits_dtype,
    MEDNET_DISPATCH_CM(p->C, {
      ce_bwd_kernel<T, TL, TO, CM><<<nb, 256, 0, stream>>>((const T*)p->logits, p->labels, p->weight, p->sums,
                                                          p->grad_out, (TO*)p->dlogits, p->N, p->S, p->batch_stride,
                                                          p->batch_stride_out, p->C);
    }))));
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

static int hm_chunks(int64_t N, int64_t S, int L, int64_t* per_chunk) {
  int64_t chunks = ((int64_t)sm_count_cached() * 4) / (N * L);
  const int64_t maxc = S / 2048;
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  *per_chunk = ceil_div64(S, chunks);
  return (int)ceil_div64(S, *per_chunk);
}

extern "C" size_t mednet_heatmap_loss_workspace_bytes(const mednet_hmloss_params* p) {
  if (!p || p->L <= 0) return 0;
  int64_t pc;
  const int chunks = hm_chunks(p->N, p->S, p->L, &pc);
  return align_up((size_t)p->N * p->L * chunks * sizeof(float), 256);
}

extern "C" int mednet_heatmap_loss_fwd(const mednet_hmloss_params* p, void* workspace, size_t workspace_bytes,
                                       mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->pred && p->target && p->weight && p->per_channel && p->loss, MEDNET_EINVAL);
  MEDNET_REQUIRE(p->N > 0 && p->N <= 65535 && p->S > 0 && p->L > 0 && p->L <= 64, MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(dtype_ok(p->pred_dtype) && (p->target_dtype == MEDNET_U8 || p->target_dtype == MEDNET_F32),
                 MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_heatmap_loss_workspace_bytes(p), MEDNET_EWORKSPACE);
  int64_t pc;
  const int chunks = hm_chunks(p->N, p->S, p->L, &pc);
  dim3 grid(chunks, p->L, (unsigned)p->N);
  float* partial = (float*)workspace;
  MEDNET_DISPATCH_LOGITS(p->pred_dtype, {
    if (p->target_dtype == MEDNET_U8)
      hm_partial_kernel<T, uint8_t><<<grid, 256, 0, stream>>>((const T*)p->pred, (const uint8_t*)p->target, partial, p->S,
                                                             p->batch_stride, p->L, p->l1, pc);
    else
      hm_partial_kernel<T, float><<<grid, 256, 0, stream>>>((const T*)p->pred, (const float*)p->target, partial, p->S,
                                                           p->batch_stride, p->L, p->l1, pc);
  });
  MEDNET_LAUNCH_CHECK();
  hm_final_kernel<<<1, 64, 0, stream>>>(partial, p->weight, p->per_channel, p->loss, p->N, p->S, p->L, chunks);
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_heatmap_loss_bwd(const mednet_hmloss_bwd_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->pred && p->target && p->weight && p->grad_out && p->dpred, MEDNET_EINVAL);
  MEDNET_REQUIRE(p->N > 0 && p->N <= 65535 && p->S > 0 && p->L > 0 && p->L <= 64, MEDNET_EUNSUPPORTED);
  MEDNET_REQUIRE(dtype_ok(p->pred_dtype) && dtype_ok(p->dpred_dtype) &&
                     (p->target_dtype == MEDNET_U8 || p->target_dtype == MEDNET_F32), MEDNET_EUNSUPPORTED);
  int gx = grid_for(p->S, 256, 4);
  const int64_t rest = (int64_t)p->N * p->L;
  if (gx > 1 && rest > 1) { gx = (int)ceil_div64(gx, rest) + 1; }
  dim3 grid(gx, p->L, (unsigned)p->N);
  MEDNET_DISPATCH_LOGITS(p->pred_dtype, MEDNET_DISPATCH_OUT(p->dpred_dtype, {
    if (p->target_dtype == MEDNET_U8)
      hm_bwd_kernel<T, uint8_t, TO><<<grid, 256, 0, stream>>>((const T*)p->pred, (const uint8_t*)p->target, p->weight,
                                                             p->grad_out, (TO*)p->dpred, p->N, p->S, p->batch_stride,
                                                             p->batch_stride_out, p->L, p->l1);
    else
      hm_bwd_kernel<T, float, TO><<<grid, 256, 0, stream>>>((const T*)p->pred, (const float*)p->target, p->weight,
                                                           p->grad_out, (TO*)p->dpred, p->N, p->S, p->batch_stride,
                                                           p->batch_stride_out, p->L, p->l1);
  }));
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_predict_epilogue(const mednet_predict_params* p, mednet_stream_t stream) {
  MEDNET_REQUIRE(p && p->logits && p->out && p->N > 0 && p->S > 0 && p->L >= 0 && p->K >= 1, MEDNET_EINVAL);
  MEDNET_REQUIRE(dtype_ok(p->logits_dtype), MEDNET_EUNSUPPORTED);
  const int nb = grid_for(p->N * p->S, 256, 8);
  MEDNET_DISPATCH_LOGITS(p->logits_dtype, {
    predict_epilogue_kernel<T><<<nb, 256, 0, stream>>>((const T*)p->logits, p->out, p->N, p->S, p->L, p->K);
  });
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}

extern "C" int mednet_final_activation(const float* logits, float* out, int64_t N, int64_t S, int32_t C,
                                       int32_t sigmoid, mednet_stream_t stream) {
  MEDNET_REQUIRE(logits && out && N > 0 && S > 0 && C > 0 && C <= MAXC, MEDNET_EUNSUPPORTED);
  const int nb = grid_for(N * S, 256, 8);
  MEDNET_DISPATCH_CM(C, { final_activation_kernel<CM><<<nb, 256, 0, stream>>>(logits, out, N, S, C, sigmoid); });
  MEDNET_LAUNCH_CHECK();
  return MEDNET_OK;
}
