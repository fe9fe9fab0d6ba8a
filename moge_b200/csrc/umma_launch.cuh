// Host-side launch helper shared by the two instantiation units of umma_kernel.
#pragma once
#include "umma_kernel.cuh"
#include "host_api.h"

namespace mg {

template <int BN, int AMODE, int EPI, bool BF16, int DF = -1>
int launch_umma_inst(const CUtensorMap& a, const CUtensorMap& aux, const CUtensorMap& b, const UmmaParams& p, int num_sms,
                     cudaStream_t st) {
    using Cfg = UmmaCfg<BN>;
    auto kern = umma_kernel<BN, AMODE, EPI, BF16, DF>;
    MG_SET_SMEM_ONCE(kern, Cfg::kSmemBytes);
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    if (tiles <= 0) return 0;
    const int grid = tiles < num_sms ? tiles : num_sms;
    CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, a, aux, b, p));
    return 0;
}

int launch_umma_rows(int bn, int epi, bool bf16, const CUtensorMap& a, const CUtensorMap& aux, const CUtensorMap& b,
                     const UmmaParams& p, int num_sms, cudaStream_t st);
int launch_umma_tiles(int bn, int epi, bool bf16, const CUtensorMap& a, const CUtensorMap& aux, const CUtensorMap& b,
                      const UmmaParams& p, int num_sms, cudaStream_t st);

}  // namespace mg
