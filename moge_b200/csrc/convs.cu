// convs_kernel instantiations + launcher.
#include "convs_kernel.cuh"
#include "host_api.h"

namespace mg {

template <int EPI, bool BF16, int DF>
static int launch_inst(const CUtensorMap& a, const CUtensorMap& aux, const CUtensorMap& w, const UmmaParams& p, int num_sms, cudaStream_t st) {
    auto kern = convs_kernel<EPI, BF16, DF>;
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvSCfg::kSmemBytes));
        attr_set = true;
    }
    const int nnt = p.num_n_tiles;
    int grid = (num_sms / nnt) * nnt;
    if (p.num_m_tiles * nnt < grid) grid = p.num_m_tiles * nnt;
    if (grid <= 0) return 0;
    kern<<<grid, ConvSCfg::kThreads, ConvSCfg::kSmemBytes, st>>>(a, aux, w, p);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// a: box {64,16,18}; aux: box {64,16,16}; w: box {64, 64}.  p.tiles_* / num_m_tiles count 16x16-pixel tiles, num_n_tiles = N / 64.
int launch_convs(int epi, bool bf16, const CUtensorMap& a, const CUtensorMap& aux, const CUtensorMap& w, const UmmaParams& p, int num_sms,
                 cudaStream_t st) {
    if (p.ntaps != 9 || p.kb_main != 1 || p.kb_aux > 1 || epi != EPI_DEC) return set_error("convs: needs a 3x3 EPI_DEC conv with C_in = 64");
    int df = (p.out0 ? DF_RAW : 0) | (p.out1 ? DF_RELU : 0) | (p.skip ? DF_SKIP : 0) | (p.vec1 ? DF_UV : 0) | (p.shuffle ? DF_SHUFFLE : 0);
#define INST(DFV)                                                                                                \
    if (df == (DFV))                                                                                             \
        return bf16 ? launch_inst<EPI_DEC, true, DFV>(a, aux, w, p, num_sms, st) : launch_inst<EPI_DEC, false, DFV>(a, aux, w, p, num_sms, st);
    INST(DF_RELU)
    INST(DF_RAW | DF_SKIP)
    INST(DF_RAW | DF_RELU | DF_SKIP)
    INST(DF_RAW | DF_RELU)
    INST(DF_RAW | DF_RELU | DF_UV)
    INST(DF_RAW | DF_UV | DF_SHUFFLE)
    df = -1;
    INST(-1)
#undef INST
    return set_error("convs: unreachable");
}

}  // namespace mg
