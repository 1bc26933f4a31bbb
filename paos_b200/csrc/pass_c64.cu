// pass_c64.cu -- instantiations of the line-pass kernel for float precision.
#include "pass_dispatch.h"
#include "pass_kernel.cuh"

namespace paosb {

#define PAOS_CASE(N, E, WR, WC, MR, MC)                                                              \
    case N:                                                                                          \
        return col ? launch_pass_t<float, N, E, WC, true, MC>(P, tw1, tw2, st, device)                 \
                   : launch_pass_t<float, N, E, WR, false, MR>(P, tw1, tw2, st, device);

cudaError_t launch_pass_c64(int n, bool col, const PassParams& P, const void* tw1, const void* tw2,
                             cudaStream_t st, int device) {
    switch (n) {
        //        N    E  Wrow Wcol minb   (complex64: W*8 B contiguous per row in a column tile, so twice the columns of the c128 table)
        PAOS_CASE(64, 8, 16, 16, 1, 1)
        PAOS_CASE(128, 8, 8, 8, 1, 1)
        PAOS_CASE(256, 16, 4, 4, 2, 2)
        PAOS_CASE(512, 8, 2, 4, 4, 4)
        PAOS_CASE(1024, 16, 2, 8, 4, 2)
        PAOS_CASE(2048, 16, 1, 4, 4, 2)
        PAOS_CASE(4096, 16, 1, 4, 2, 1)
        default: return cudaErrorInvalidValue;
    }
}
#undef PAOS_CASE
}  // namespace paosb
