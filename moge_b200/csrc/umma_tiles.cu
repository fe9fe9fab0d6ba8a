// umma_kernel instantiations with a pixel-tile A operand (decoder implicit-GEMM convolutions).
#include "umma_launch.cuh"

namespace mg {

#define INST(BN, EPI)                                                                                       \
    if (bn == BN && epi == EPI)                                                                             \
        return bf16 ? launch_umma_inst<BN, AMODE_TILES, EPI, true>(a, aux, b, p, num_sms, st)               \
                    : launch_umma_inst<BN, AMODE_TILES, EPI, false>(a, aux, b, p, num_sms, st);

int launch_umma_tiles(int bn, int epi, bool bf16, const CUtensorMap& a, const CUtensorMap& aux, const CUtensorMap& b,
                      const UmmaParams& p, int num_sms, cudaStream_t st) {
    INST(256, EPI_DEC) INST(128, EPI_DEC) INST(64, EPI_DEC) INST(32, EPI_DEC)
    INST(16, EPI_HEADOUT)
    return set_error("no umma_tiles instantiation for bn=%d epi=%d", bn, epi);
}
#undef INST

}  // namespace mg
