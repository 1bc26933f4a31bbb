// pass_dispatch.h -- host entry points of the pass kernels (one translation unit per precision).
#pragma once
#include "device_types.h"

namespace paosb {
// points per thread (= first/last radix) used for grid size n
inline int geom_E(int n) { return (n == 64 || n == 128 || n == 512) ? 8 : 16; }
// lines per CTA (tile width) of the row / column pass for grid size n and precision dtype (0 = c128, 1 = c64)
int tile_width_c128(int n, bool col);
int tile_width_c64(int n, bool col);
inline int tile_width(int n, int dtype, bool col) { return dtype == 0 ? tile_width_c128(n, col) : tile_width_c64(n, col); }
cudaError_t launch_pass_c128(int n, bool col, const PassParams* const* Ps, int nb, const void* tw1, const void* tw2,
                             cudaStream_t st, int device);
cudaError_t launch_pass_c64(int n, bool col, const PassParams* const* Ps, int nb, const void* tw1, const void* tw2,
                            cudaStream_t st, int device);
}  // namespace paosb
