// Shared device helpers for the mednet_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mednet_b200.h"

#define MEDNET_LAUNCH_CHECK()                          \
  do {                                                 \
    cudaError_t e__ = cudaGetLastError();              \
    if (e__ != cudaSuccess) return (int)e__;           \
  } while (0)

#define MEDNET_REQUIRE(cond, code) \
  do {                             \
    if (!(cond)) return (code);    \
  } while (0)

namespace mednet {

typedef __nv_bfloat16 bf16;

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

int sm_count_cached();

// ------------------------------------------------------------------------------------------------
// scalar conversions
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<uint8_t>(uint8_t v) { return (float)v; }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ uint8_t from_f32<uint8_t>(float v) { return (uint8_t)v; }

// ------------------------------------------------------------------------------------------------
// V-wide vector load/store (V * sizeof(T) in {2,4,8,16} bytes, naturally aligned)
// ------------------------------------------------------------------------------------------------
template <int BYTES> struct RawVec;
struct alignas(32) U32B { uint4 lo, hi; };
template <> struct RawVec<32> { typedef U32B type; };
template <> struct RawVec<16> { typedef uint4 type; };
template <> struct RawVec<8> { typedef uint2 type; };
template <> struct RawVec<4> { typedef uint32_t type; };
template <> struct RawVec<2> { typedef uint16_t type; };
template <> struct RawVec<1> { typedef uint8_t type; };

template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, float (&out)[V]) {
  typedef typename RawVec<sizeof(T) * V>::type R;
  union { R r; T t[V]; } u;
  u.r = *reinterpret_cast<const R*>(p);
#pragma unroll
  for (int i = 0; i < V; ++i) out[i] = to_f32<T>(u.t[i]);
}

// split form: issue the raw load now, convert later (keeps only sizeof(T)*V bytes of registers live per pending load)
template <typename T, int V>
__device__ __forceinline__ typename RawVec<sizeof(T) * V>::type load_raw(const T* __restrict__ p) {
  return *reinterpret_cast<const typename RawVec<sizeof(T) * V>::type*>(p);
}
template <typename T, int V>
__device__ __forceinline__ void cvt_raw(const typename RawVec<sizeof(T) * V>::type& r, float (&out)[V]) {
  union { typename RawVec<sizeof(T) * V>::type r; T t[V]; } u;
  u.r = r;
#pragma unroll
  for (int i = 0; i < V; ++i) out[i] = to_f32<T>(u.t[i]);
}

template <typename T, int V>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const float (&in)[V]) {
  typedef typename RawVec<sizeof(T) * V>::type R;
  union { R r; T t[V]; } u;
#pragma unroll
  for (int i = 0; i < V; ++i) u.t[i] = from_f32<T>(in[i]);
  *reinterpret_cast<R*>(p) = u.r;
}

// largest vector width (elements) with V | c and V*size <= 16 bytes
static inline int pick_vec(int64_t c, int elem_bytes) {
  int v = 16 / elem_bytes;
  while (v > 1 && (c % v) != 0) v >>= 1;
  return v;
}

// ------------------------------------------------------------------------------------------------
// per-thread asynchronous global->shared copies (LDGSTS).  Used as a PRIVATE software pipeline: a thread copies
// the operands of its next few loop iterations into its own shared-memory slots, so many loads are in flight
// without holding registers, and reads back only what it wrote itself (no block barrier needed).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------
// activations (ref: midasmednet/unet/components.py:35-40)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float x, int act, float a) {
  switch (act) {
    case MEDNET_ACT_RELU: return x > 0.f ? x : 0.f;
    case MEDNET_ACT_LEAKY: return x > 0.f ? x : a * x;
    case MEDNET_ACT_ELU: return x > 0.f ? x : expm1f(x);
    default: return x;
  }
}
// derivative expressed through the OUTPUT y (what the in-place reference modules keep)
__device__ __forceinline__ float act_grad_from_out(float y, int act, float a) {
  switch (act) {
    case MEDNET_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case MEDNET_ACT_LEAKY: return y > 0.f ? 1.f : a;
    case MEDNET_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    default: return 1.f;
  }
}

// ------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the block; result valid in thread 0.  `scratch` must hold >= 32 values.
template <typename F>
__device__ __forceinline__ F block_sum(F v, F* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  F r = (threadIdx.x < nw) ? scratch[threadIdx.x] : (F)0;
  if (warp == 0) r = warp_sum(r);
  return r;
}

// grid size for a grid-stride loop over `work` thread-items
static inline int grid_for(int64_t work, int block, int waves = 8) {
  int64_t blocks = ceil_div64(work, block);
  int64_t cap = (int64_t)sm_count_cached() * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// dispatch helpers ---------------------------------------------------------------------------------
#define MEDNET_DISPATCH_TV(dtype, V, ...)                                   \
  do {                                                                      \
    if ((dtype) == MEDNET_F32) {                                            \
      typedef float T;                                                      \
      switch (V) {                                                          \
        case 4: { constexpr int VV = 4; __VA_ARGS__; } break;               \
        case 2: { constexpr int VV = 2; __VA_ARGS__; } break;               \
        default: { constexpr int VV = 1; __VA_ARGS__; } break;              \
      }                                                                     \
    } else {                                                                \
      typedef bf16 T;                                                       \
      switch (V) {                                                          \
        case 8: { constexpr int VV = 8; __VA_ARGS__; } break;               \
        case 4: { constexpr int VV = 4; __VA_ARGS__; } break;               \
        case 2: { constexpr int VV = 2; __VA_ARGS__; } break;               \
        default: { constexpr int VV = 1; __VA_ARGS__; } break;              \
      }                                                                     \
    }                                                                       \
  } while (0)

static inline bool dtype_ok(int dtype) { return dtype == MEDNET_F32 || dtype == MEDNET_BF16; }
static inline int dtype_bytes(int dtype) { return dtype == MEDNET_F32 ? 4 : 2; }


}  // namespace mednet
