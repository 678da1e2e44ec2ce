// qft_fused.cu -- fused tile sweeps (placeholder until the tile kernels land).
#include "qcs_internal.h"

int qcs_fused_qft(qcs_register *, unsigned, unsigned, bool) { return QCS_UNKNOWN_ERROR; }
int qcs_fused_modexp(qcs_register *, unsigned, const unsigned *, unsigned) { return QCS_UNKNOWN_ERROR; }
