// pass_c64.cu -- instantiations of the line-pass kernel for float precision.
#include "pass_dispatch.h"
#include "pass_kernel.cuh"

namespace paosb {

// 512^2: five 128-thread CTAs per SM (96 registers; measured +3-4 % over four at 128 registers, six spill and lose 8 %)
#ifndef PAOS_EXP_512_MINB
#define PAOS_EXP_512_MINB 5
#endif
#define PAOS_ROW_512 PAOS_CASE(512, 8, 2, 4, PAOS_EXP_512_MINB, PAOS_EXP_512_MINB)
//                 N    E  Wrow Wcol minb(row) minb(col)
#define PAOS_TILE_TABLE \
    PAOS_CASE(64, 8, 16, 16, 1, 1) \
    PAOS_CASE(128, 8, 8, 8, 1, 1) \
    PAOS_CASE(256, 16, 4, 4, 2, 2) \
    PAOS_ROW_512 \
    PAOS_CASE(1024, 16, 2, 8, 4, 2) \
    PAOS_CASE(2048, 16, 1, 4, 4, 2) \
    PAOS_CASE(4096, 16, 1, 4, 2, 1)

#define PAOS_CASE(N, E, WR, WC, MR, MC)                                                              \
    case N:                                                                                          \
        return col ? launch_pass_t<float, N, E, WC, true, MC>(Ps, nb, tw1, tw2, st, device)                 \
                   : launch_pass_t<float, N, E, WR, false, MR>(Ps, nb, tw1, tw2, st, device);

cudaError_t launch_pass_c64(int n, bool col, bool /*wide*/, const PassParams* const* Ps, int nb, const void* tw1, const void* tw2,
                             cudaStream_t st, int device) {
    switch (n) {
        PAOS_TILE_TABLE
        default: return cudaErrorInvalidValue;
    }
}
#undef PAOS_CASE

#define PAOS_CASE(N, E, WR, WC, MR, MC) \
    case N:                             \
        return col ? WC : WR;
int tile_width_c64(int n, bool col, bool /*wide*/) {
    switch (n) {
        PAOS_TILE_TABLE
        default: return 1;
    }
}
#undef PAOS_CASE
}  // namespace paosb
