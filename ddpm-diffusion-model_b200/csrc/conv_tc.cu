// tcgen05 / TMEM implicit-GEMM convolution (bf16).  Placeholder until the tensor-core kernels land:
// reports "unsupported" so ddpm_conv dispatches every shape to conv_simt.cu.
#include "common.cuh"
int conv_tc_supported(const ddpm_conv_args*) { return 0; }
int conv_tc_launch(const ddpm_conv_args*, cudaStream_t) { return DDPM_E_ARG; }
int wgrad_tc_supported(const ddpm_wgrad_args*) { return 0; }
int wgrad_tc_launch(const ddpm_wgrad_args*, cudaStream_t) { return DDPM_E_ARG; }
