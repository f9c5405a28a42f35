// tensor_pass.h -- interface between the C ABI host code and the tcgen05 /
// TMEM / TMA contraction kernel (tensor_pass.cu).
#pragma once
#include "common.cuh"

bool tp_supported(const nbmf_handle *h, int K);   // K, sizes and device fit the tensor path
// AUTO precision: the fastest nbmf_path whose error growth stays inside the 1e-5 gate for `horizon`
// updates at this size (NBMF_PATH_FP64 if no tensor variant does)
int tp_auto_mode(const nbmf_handle *h, int K, int horizon);
int tp_prepare(nbmf_handle *h);               // allocate f16 operand buffers + tensor maps
int tp_convert_operands(nbmf_handle *h, bool rows_sliced = false);   // rows_sliced: own rows of A only, then all-gather of the f16 parts      // A, BB (fp64) -> scaled f16 hi/lo
int tp_pass(nbmf_handle *h, int owner_rows, double *sumlogD);
int tp_log_correction_rows_slice(nbmf_handle *h, double *sumlogD);   // row-sliced sessions: after the exchange, own rows
// both passes, column block by column block behind the events of a pipelined upload
int tp_passes_pipelined(nbmf_handle *h, int n_blocks, const int64_t *n0, cudaEvent_t *ev,
                        int (*prepare_block)(nbmf_handle *, int64_t, int64_t), double *sumlogD);
int tp_log_correction(nbmf_handle *h, int owner_rows, double *sumlogD); // call after each pass, before the exchange
int tp_finalize(nbmf_handle *h);            // statistics = U (.) G with exact-count row normalisation
void tp_invalidate_data(nbmf_handle *h);     // V changed under the same shape: recount on next use
void tp_free(nbmf_handle *h);
int tp_check_error(nbmf_handle *h);         // reports (and clears) a kernel-side barrier timeout
