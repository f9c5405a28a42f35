// tensor_pass.h -- interface between the C ABI host code and the tcgen05 /
// TMEM / TMA contraction kernel (tensor_pass.cu).
#pragma once
#include "common.cuh"

bool tp_supported(const nbmf_handle *h);      // K, sizes and device fit the tensor path
int tp_prepare(nbmf_handle *h);               // allocate f16 operand buffers + tensor maps
int tp_convert_operands(nbmf_handle *h);      // A, BB (fp64) -> scaled f16 hi/lo
int tp_pass(nbmf_handle *h, int owner_rows, double *sumlogD);
void tp_free(nbmf_handle *h);
