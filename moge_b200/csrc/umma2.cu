// umma2_kernel (cta_group::2) instantiations + launcher.
#include "umma2_kernel.cuh"
#include "host_api.h"

namespace mg {

template <int EPI, bool BF16>
static int launch_inst(const CUtensorMap& a, const CUtensorMap& b, const UmmaParams& p, int num_sms, cudaStream_t st) {
    auto kern = umma2_kernel<EPI, BF16>;
    MG_SET_SMEM_ONCE(kern, Umma2Cfg::kSmemBytes);
    const int total = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    if (total <= 0) return 0;
    const int pairs = total < num_sms / 2 ? total : num_sms / 2;
    CUDA_TRY(launch_pdl(kern, dim3(2 * pairs), dim3(Umma2Cfg::kThreads), Umma2Cfg::kSmemBytes, st, a, b, p));
    return 0;
}

// a: box {64,128}; b: box {64,128} (each CTA of the pair stages half of the 256-column tile); p.num_n_tiles = N / 256.
int launch_umma2(int epi, bool bf16, const CUtensorMap& a, const CUtensorMap& b, const UmmaParams& p, int num_sms, cudaStream_t st) {
    if (p.N % 256) return set_error("umma2: N=%d must be a multiple of 256", p.N);
#define INST(EPI)                                                                                     \
    if (epi == EPI) return bf16 ? launch_inst<EPI, true>(a, b, p, num_sms, st) : launch_inst<EPI, false>(a, b, p, num_sms, st);
    INST(EPI_STORE16) INST(EPI_GELU16) INST(EPI_RESID)
#undef INST
    return set_error("no umma2 instantiation for epi=%d", epi);
}

}  // namespace mg
