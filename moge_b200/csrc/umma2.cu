// umma2_kernel (cta_group::2) instantiations + launcher.
#include "umma2_kernel.cuh"
#include "host_api.h"

namespace mg {

template <int EPI, bool BF16, int RS = 0>
static int launch_inst(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& r, const UmmaParams& p, int num_sms, cudaStream_t st) {
    auto kern = umma2_kernel<EPI, BF16, RS>;
    MG_SET_SMEM_ONCE(kern, Umma2CfgT<RS>::kSmemBytes);
    const int total = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    if (total <= 0) return 0;
    const int pairs = total < num_sms / 2 ? total : num_sms / 2;
    CUDA_TRY(launch_pdl(kern, dim3(2 * pairs), dim3(Umma2CfgT<RS>::kThreads), Umma2CfgT<RS>::kSmemBytes, st, a, b, r, p));
    return 0;
}

// a: box {64,128}; b: box {64,128} (each CTA of the pair stages half of the 256-column tile); p.num_n_tiles = N / 256.
int launch_umma2(int epi, bool bf16, const CUtensorMap& a, const CUtensorMap& b, const UmmaParams& p, int num_sms, cudaStream_t st,
                 const CUtensorMap* resid, int resid_bufs) {
    if (p.N % 256) return set_error("umma2: N=%d must be a multiple of 256", p.N);
    if (epi == EPI_RESID && resid != nullptr && resid_bufs == 2)
        return bf16 ? launch_inst<EPI_RESID, true, 2>(a, b, *resid, p, num_sms, st) : launch_inst<EPI_RESID, false, 2>(a, b, *resid, p, num_sms, st);
    if (epi == EPI_RESID && resid != nullptr && resid_bufs == 1)
        return bf16 ? launch_inst<EPI_RESID, true, 1>(a, b, *resid, p, num_sms, st) : launch_inst<EPI_RESID, false, 1>(a, b, *resid, p, num_sms, st);
#define INST(EPI)                                                                                     \
    if (epi == EPI) return bf16 ? launch_inst<EPI, true>(a, b, a, p, num_sms, st) : launch_inst<EPI, false>(a, b, a, p, num_sms, st);
    INST(EPI_STORE16) INST(EPI_GELU16) INST(EPI_RESID)
#undef INST
    return set_error("no umma2 instantiation for epi=%d", epi);
}

}  // namespace mg
