// tcgen05 (kind::tf32) dense-layer variant -- placeholder until the TMEM/TMA kernel lands: nothing is eligible,
// so FBSNN_PREC_TF32 currently runs the fp32 SIMT kernels.
#pragma once
#include "gemm_simt.cuh"

namespace fbsnn {
template <bool A_KC, bool B_KC>
inline bool tc_eligible(const GemmArgs&, int) { return false; }
template <bool A_KC, bool B_KC, class Epi>
inline cudaError_t launch_gemm_tc(const GemmArgs&, const Epi&, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace fbsnn
