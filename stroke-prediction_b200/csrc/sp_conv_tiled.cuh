// sp_conv_tiled.cuh — shared-memory tiled fast paths for the FFMA-bound 3x3x3 stride-1 layers.
// (placeholder: fast paths are enabled once the generic kernels are parity-green on the GPU)
#pragma once
#include "sp_common.cuh"

static inline bool sp_tiled_corr_supported(const SpConvDesc*) { return false; }
static inline int sp_tiled_corr_launch(const SpConvDesc*, int, const float*, const float*, int, const float*, const float*,
                                       const float*, float*, cudaStream_t) { return -1; }
static inline bool sp_tiled_wgrad_supported(const SpConvDesc*) { return false; }
static inline size_t sp_tiled_wgrad_workspace_bytes(const SpConvDesc*) { return 0; }
static inline int sp_tiled_wgrad_launch(const SpConvDesc*, int, const float*, const float*, const float*, const float*,
                                        float*, float, float*, cudaStream_t) { return -1; }
