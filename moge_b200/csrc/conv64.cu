// conv64_kernel instantiations + launcher.
#include "conv64_kernel.cuh"
#include "host_api.h"

namespace mg {

template <int BN, int EPI, bool BF16>
static int launch_inst(const CUtensorMap& a, const CUtensorMap& aux, const CUtensorMap& w, const UmmaParams& p, int num_sms, cudaStream_t st) {
    using Cfg = Conv64Cfg<BN>;
    auto kern = conv64_kernel<BN, EPI, BF16>;
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        attr_set = true;
    }
    const int nnt = p.num_n_tiles;
    int grid = (num_sms / nnt) * nnt;
    if (p.num_m_tiles * nnt < grid) grid = p.num_m_tiles * nnt;
    if (grid <= 0) return 0;
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(a, aux, w, p);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_conv64(int bn, int epi, bool bf16, const CUtensorMap& a, const CUtensorMap& aux, const CUtensorMap& w, const UmmaParams& p,
                  int num_sms, cudaStream_t st) {
    if (p.ntaps != 9 || p.kb_main != 1 || p.kb_aux > 1) return set_error("conv64: needs a 3x3 conv with C_in = 64");
    if (p.N % bn) return set_error("conv64: N=%d not a multiple of %d", p.N, bn);
#define INST(BN, EPI)                                                                                         \
    if (bn == BN && epi == EPI)                                                                               \
        return bf16 ? launch_inst<BN, EPI, true>(a, aux, w, p, num_sms, st) : launch_inst<BN, EPI, false>(a, aux, w, p, num_sms, st);
    INST(64, EPI_DEC)
    INST(16, EPI_HEADOUT)
#undef INST
    return set_error("no conv64 instantiation for bn=%d epi=%d", bn, epi);
}

}  // namespace mg
