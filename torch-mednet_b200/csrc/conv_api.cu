// Dispatch of the 3x3x3 convolution entry points between the tcgen05 tensor-core kernels
// (conv_tcgen05.cu) and the CUDA-core gather implicit GEMM (conv_simt.cu).  There is no fallback in the
// silent sense: MEDNET_IMPL_AUTO resolves deterministically from shape/dtype/device, an explicit impl
// that cannot take the problem returns MEDNET_EUNSUPPORTED.
#include "common.cuh"
#include "conv_impl.h"

using namespace mednet;

static bool conv_args_ok(const mednet_conv3d_params* p) {
  return p && p->x && p->w && p->y && p->N > 0 && p->Di > 0 && p->Hi > 0 && p->Wi > 0 && p->Do > 0 && p->Ho > 0 &&
         p->Wo > 0 && p->K > 0 && p->Nout > 0 && dtype_ok(p->dtype) && p->gather >= 0 && p->gather <= MEDNET_GATHER_UPCONV_B;
}

extern "C" int mednet_conv3d_select_impl(const mednet_conv3d_params* p) {
  if (!p) return MEDNET_EINVAL;
  if (p->impl == MEDNET_IMPL_SIMT) return p->gather >= MEDNET_GATHER_UPCONV_F ? MEDNET_EUNSUPPORTED : MEDNET_IMPL_SIMT;
  const bool tc = tc_fprop_supported(p);
  if (p->impl == MEDNET_IMPL_TCGEN05) return tc ? MEDNET_IMPL_TCGEN05 : MEDNET_EUNSUPPORTED;
  if (!tc && p->gather >= MEDNET_GATHER_UPCONV_F) return MEDNET_EUNSUPPORTED;   // tensor-core kernel only (callers use the
  return tc ? MEDNET_IMPL_TCGEN05 : MEDNET_IMPL_SIMT;                           // materialised-concat path otherwise)
}

extern "C" size_t mednet_conv3d_workspace_bytes(const mednet_conv3d_params* p) {
  (void)p;
  return 256;
}

extern "C" int mednet_conv3d_fprop(const mednet_conv3d_params* p, void* workspace, size_t workspace_bytes,
                                   mednet_stream_t stream) {
  MEDNET_REQUIRE(conv_args_ok(p), MEDNET_EINVAL);
  if (p->gather == MEDNET_GATHER_CONV3)
    MEDNET_REQUIRE(p->Di == p->Do && p->Hi == p->Ho && p->Wi == p->Wo, MEDNET_EINVAL);
  else if (p->gather == MEDNET_GATHER_CONVT_F || p->gather == MEDNET_GATHER_UPCONV_F)
    MEDNET_REQUIRE(p->Do == 2 * p->Di && p->Ho == 2 * p->Hi && p->Wo == 2 * p->Wi, MEDNET_EINVAL);
  else
    MEDNET_REQUIRE(p->Di == 2 * p->Do && p->Hi == 2 * p->Ho && p->Wi == 2 * p->Wo, MEDNET_EINVAL);
  const int impl = mednet_conv3d_select_impl(p);
  if (impl < 0) return impl;
  if ((p->y_f32 || p->addend_f32) && impl != MEDNET_IMPL_TCGEN05) return MEDNET_EUNSUPPORTED;
  if (impl == MEDNET_IMPL_TCGEN05) return tc_fprop(p, workspace, workspace_bytes, stream);   // workspace: profiling counters only
  return simt_fprop(p, stream);
}

static bool wgrad_args_ok(const mednet_wgrad_params* p) {
  return p && p->a && p->b && p->dw && p->N > 0 && p->Da > 0 && p->Ha > 0 && p->Wa > 0 && p->Db > 0 && p->Hb > 0 &&
         p->Wb > 0 && p->Ca > 0 && p->Cb > 0 && dtype_ok(p->dtype) &&
         (p->gather == MEDNET_GATHER_CONV3 || p->gather == MEDNET_GATHER_CONVT_B || p->gather == MEDNET_GATHER_UPCONV_B);
}

extern "C" int mednet_conv3d_wgrad_select_impl(const mednet_wgrad_params* p) {
  if (!p) return MEDNET_EINVAL;
  const bool upc = p->gather == MEDNET_GATHER_UPCONV_B;        // tensor-core kernel only
  if (p->impl == MEDNET_IMPL_SIMT) return upc ? MEDNET_EUNSUPPORTED : MEDNET_IMPL_SIMT;
  const bool tc = tc_wgrad_supported(p);
  if (p->impl == MEDNET_IMPL_TCGEN05 || upc) return tc ? MEDNET_IMPL_TCGEN05 : MEDNET_EUNSUPPORTED;
  return tc ? MEDNET_IMPL_TCGEN05 : MEDNET_IMPL_SIMT;
}

extern "C" size_t mednet_conv3d_wgrad_workspace_bytes(const mednet_wgrad_params* p) {
  if (!wgrad_args_ok(p)) return 0;
  const int impl = mednet_conv3d_wgrad_select_impl(p);
  if (impl < 0) return 0;
  if (impl == MEDNET_IMPL_TCGEN05) return tc_wgrad_workspace_bytes(p);
  return simt_wgrad_workspace_bytes(p);
}

extern "C" int mednet_conv3d_wgrad(const mednet_wgrad_params* p, void* workspace, size_t workspace_bytes,
                                   mednet_stream_t stream) {
  MEDNET_REQUIRE(wgrad_args_ok(p), MEDNET_EINVAL);
  if (p->gather == MEDNET_GATHER_CONV3)
    MEDNET_REQUIRE(p->Da == p->Db && p->Ha == p->Hb && p->Wa == p->Wb, MEDNET_EINVAL);
  else
    MEDNET_REQUIRE(p->Db == 2 * p->Da && p->Hb == 2 * p->Ha && p->Wb == 2 * p->Wa, MEDNET_EINVAL);
  MEDNET_REQUIRE(workspace && workspace_bytes >= mednet_conv3d_wgrad_workspace_bytes(p), MEDNET_EWORKSPACE);
  const int impl = mednet_conv3d_wgrad_select_impl(p);
  if (impl < 0) return impl;
  if (impl == MEDNET_IMPL_TCGEN05) return tc_wgrad(p, workspace, stream);
  MEDNET_REQUIRE(p->dw_ld == 0, MEDNET_EUNSUPPORTED);          // strided destinations: tensor-core implementation only
  return simt_wgrad(p, workspace, stream);
}
