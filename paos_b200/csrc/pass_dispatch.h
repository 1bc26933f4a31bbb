// pass_dispatch.h -- host entry points of the pass kernels (one translation unit per precision).
#pragma once
#include "device_types.h"

namespace paosb {
// points per thread (= first/last radix) used for grid size n
inline int geom_E(int n) { return (n == 64 || n == 128 || n == 512) ? 8 : 16; }
// lines per CTA (tile width) of the row / column pass for grid size n and precision dtype (0 = c128, 1 = c64).
// Column passes of some sizes come in two widths: the default one and a *wide* one for passes with a fused read-out, whose
// real-valued tile needs W * sizeof(real) >= 16 bytes per row to leave through the TMA unit (complex128 2048^2: one column
// per CTA and four CTAs per SM by default, two columns per CTA for the read-out pass).  wide is ignored where there is one width.
int tile_width_c128(int n, bool col, bool wide);
int tile_width_c64(int n, bool col, bool wide);
inline int tile_width(int n, int dtype, bool col, bool wide = false) {
    return dtype == 0 ? tile_width_c128(n, col, wide) : tile_width_c64(n, col, wide);
}
inline bool has_wide_tiles(int n, int dtype) { return tile_width(n, dtype, true, true) != tile_width(n, dtype, true, false); }
cudaError_t launch_pass_c128(int n, bool col, bool wide, const PassParams* const* Ps, int nb, const void* tw1, const void* tw2,
                             cudaStream_t st, int device);
cudaError_t launch_pass_c64(int n, bool col, bool wide, const PassParams* const* Ps, int nb, const void* tw1, const void* tw2,
                            cudaStream_t st, int device);
}  // namespace paosb
