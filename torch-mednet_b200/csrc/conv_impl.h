// Internal interfaces between the convolution dispatcher and its two implementations.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "mednet_b200.h"

namespace mednet {
int simt_fprop(const mednet_conv3d_params* p, cudaStream_t st);
int simt_wgrad(const mednet_wgrad_params* p, void* workspace, cudaStream_t st);
size_t simt_wgrad_workspace_bytes(const mednet_wgrad_params* p);
size_t colsum_workspace_bytes(const mednet_wgrad_params* p);
int colsum_bias(const void* a, int dtype, int64_t M, int C, float* dbias, int accumulate, void* workspace,
                cudaStream_t st);

// first layer (in_channels = 1) on warp-level mma.sync (first_layer_mma.cu)
void in1_mma_set_enabled(int v);
bool in1_mma_fprop_ok(const mednet_conv3d_params* p);
int in1_mma_fprop(const mednet_conv3d_params* p, cudaStream_t st);
bool in1_mma_wgrad_ok(const mednet_wgrad_params* p);
int in1_mma_wgrad(const mednet_wgrad_params* p, float* partial, int max_blocks, int* blocks_out, cudaStream_t st);

bool tc_fprop_supported(const mednet_conv3d_params* p);
int tc_fprop(const mednet_conv3d_params* p, void* workspace, size_t workspace_bytes, cudaStream_t st);
bool tc_wgrad_supported(const mednet_wgrad_params* p);
void tc_wgrad_set_wt_fastest(int v);
void tc_wgrad_set_pair_planes(int v);
void tc_wgrad_set_d_fastest(int v);
void tc_wgrad_set_profile(int v);
void tc_wgrad_set_dual(int v);
void tc_wgrad_set_class_merge(int v);
void tc_wgrad_set_reduce_s_fastest(int v);
size_t tc_wgrad_workspace_bytes(const mednet_wgrad_params* p);
int tc_wgrad(const mednet_wgrad_params* p, void* workspace, cudaStream_t st);
}  // namespace mednet
