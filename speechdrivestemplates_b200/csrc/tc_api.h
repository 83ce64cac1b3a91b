// Internal interface between the FFMA dispatchers (conv_gemm.cu) and the tcgen05 kernels (tc_conv.cu, tc_wgrad.cu).
#pragma once
#include <cuda_runtime.h>

#include "sdt_b200.h"

bool sdt_tc_conv_eligible(const sdt_conv_desc* d);
int sdt_tc_conv_launch(const sdt_conv_desc* d, int row_tiles, cudaStream_t st);
bool sdt_tc_wgrad_eligible(const sdt_conv_desc* d);
int sdt_tc_wgrad_launch(const sdt_conv_desc* d, cudaStream_t st);
bool sdt_tc_conv_tma_eligible(const sdt_conv_desc* d);
int sdt_tc_conv_tma_row_tiles(const sdt_conv_desc* d);
int sdt_tc_conv_tma_launch(const sdt_conv_desc* d, cudaStream_t st);
bool sdt_tc_conv_tma_rownorm_ok(const sdt_conv_desc* d);      // fused row-norm epilogue (cluster of N / 64 = 4 CTAs per row tile)
bool sdt_tc_conv_ytap_eligible(const sdt_conv_desc* d);
bool sdt_tc_conv_ytap_shape_ok(const sdt_conv_desc* d);
int sdt_tc_conv_ytap_row_tiles(const sdt_conv_desc* d);
int sdt_tc_conv_ytap_describe(const sdt_conv_desc* d, int32_t* out10);
int sdt_tc_conv_ytap_launch(const sdt_conv_desc* d, cudaStream_t st);
// several problems (stride-parity classes of one data gradient) as one persistent launch
bool sdt_tc_conv_ytap_multi_ok(const sdt_conv_desc* ds, int n);
int sdt_tc_conv_ytap_launch_multi(const sdt_conv_desc* ds, int n, cudaStream_t st);
bool sdt_tc_conv_pair_eligible(const sdt_conv_desc* d);
bool sdt_tc_conv_pair_shape_ok(const sdt_conv_desc* d);
int sdt_tc_conv_pair_row_tiles(const sdt_conv_desc* d);
int sdt_tc_conv_pair_describe(const sdt_conv_desc* d, int32_t* out10);
int sdt_tc_conv_pair_launch(const sdt_conv_desc* d, cudaStream_t st);
bool sdt_tc_wgrad_tma_eligible(const sdt_conv_desc* d);
int sdt_tc_wgrad_tma_launch(const sdt_conv_desc* d, cudaStream_t st);
bool sdt_tc_wgrad_ytap_eligible(const sdt_conv_desc* d);
int sdt_tc_wgrad_ytap_launch(const sdt_conv_desc* d, cudaStream_t st);
void sdt_note_tc_launch();
