// conv_args.h -- kernel argument block shared by the CUDA-core (conv.cu) and tcgen05 (conv_tc.cu) paths.
#pragma once
#include "kernels.h"

namespace b2s {

struct ConvKArgs {
    ConvGeom g;
    const float* act[kMaxPairs];   // gather source (forward: x-like; dgrad: ybar-like; wgrad: x-like)
    const float* wt[kMaxPairs];    // weights (forward/dgrad) or adjoint (wgrad)
    float scale[kMaxPairs];
    int npairs;
    const float* bias;
    const float* relu_ref;
    int relu_mode;
    float* out;
    int accumulate;
    int k_chunk;                   // wgrad: K elements per blockIdx.z
};

enum { MODE_FWD = 0, MODE_DGRAD = 1 };

// tcgen05 / TMEM path for wide layers (conv_tc.cu). mode: MODE_FWD or MODE_DGRAD.
// Returns 1 when the kernel was launched, 0 when the shape is not eligible, <0 on error.
int try_launch_conv_tc(int mode, cudaStream_t st, const ConvKArgs& a);
// 0 = never, 1 = automatic (default), 2 = whenever legal (tests)
void set_tc_mode(int mode);
int get_tc_mode();

}  // namespace b2s
