// tensor_pass.cu -- placeholder until the tcgen05 kernel lands (same commit series)
#include "tensor_pass.h"
bool tp_supported(const nbmf_handle *) { return false; }
int tp_prepare(nbmf_handle *h) { h->err = "tensor path not built"; return NBMF_ERR_UNSUPPORTED; }
int tp_convert_operands(nbmf_handle *h) { h->err = "tensor path not built"; return NBMF_ERR_UNSUPPORTED; }
int tp_pass(nbmf_handle *h, int, double *) { h->err = "tensor path not built"; return NBMF_ERR_UNSUPPORTED; }
void tp_free(nbmf_handle *) {}
