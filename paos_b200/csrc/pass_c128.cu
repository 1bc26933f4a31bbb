// pass_c128.cu -- instantiations of the line-pass kernel for double precision.
#include "pass_dispatch.h"
#include "pass_kernel.cuh"

namespace paosb {

// 2048^2 column passes come in two tile widths (pass_dispatch.h).  Default: ONE column per CTA, 128 threads, four CTAs per SM
// like the row kernel -- the field moves through 16-byte-wide TMA boxes, a CTA that waits for its tile or drains its store
// idles a quarter of the SM instead of half, and a barrier couples four warps instead of eight (measured, batches of 8:
// column x4 261 -> 241 us, x6 314 -> 292, x8 351 -> 327).  Wide: TWO columns per CTA, two CTAs per SM, phase tables staged by
// the bulk-copy engine -- the pass with the fused read-out, whose tile of reals must be 16 bytes per row for the TMA unit
// (written straight from registers by single-column CTAs that pass takes 561 instead of 435 us).
// -DPAOS_EXP_COLW_2048=2 builds the round's earlier layout (two columns everywhere) for the A/B measurement.
#ifndef PAOS_EXP_COLW_2048
#define PAOS_EXP_COLW_2048 1
#endif
#if PAOS_EXP_COLW_2048 == 1
#define PAOS_COL_2048_NARROW launch_pass_t<double, 2048, 16, 1, true, 4>
#else
#define PAOS_COL_2048_NARROW launch_pass_t<double, 2048, 16, 2, true, 2>
#endif
#define PAOS_COL_2048_WIDE launch_pass_t<double, 2048, 16, 2, true, 2>
// 4096^2: the same two widths (256 or 512 threads per CTA, two CTAs or one per SM; AIRS-CH0 4096^2: 417 -> 443 PSF/s)
#ifndef PAOS_EXP_COLW_4096
#define PAOS_EXP_COLW_4096 1
#endif
#if PAOS_EXP_COLW_4096 == 1
#define PAOS_COL_4096_NARROW launch_pass_t<double, 4096, 16, 1, true, 2>
#else
#define PAOS_COL_4096_NARROW launch_pass_t<double, 4096, 16, 2, true, 1>
#endif
#define PAOS_COL_4096_WIDE launch_pass_t<double, 4096, 16, 2, true, 1>
#ifdef PAOS_EXP_ROW2  // experiment: two rows per CTA (256 threads, 2 CTAs/SM) like the column kernel
#define PAOS_ROW_2048 launch_pass_t<double, 2048, 16, 2, false, 2>
#define PAOS_ROW_2048_W 2
#else  // (five row CTAs per SM would need 96 registers: 1.7 KB of spills per thread, measured slower in round 1)
#define PAOS_ROW_2048 launch_pass_t<double, 2048, 16, 1, false, 4>
#define PAOS_ROW_2048_W 1
#endif
// 512^2: five 128-thread CTAs per SM (96 registers; measured +3-4 % over four at 128 registers, six spill and lose 8 %)
#ifndef PAOS_EXP_512_MINB
#define PAOS_EXP_512_MINB 5
#endif
// (1024^2 column kernel at two columns per CTA and four CTAs per SM, with or without TMA tiles: within +-3 % of four columns
// and two CTAs on AIRS / TA-Ground, -7 % on Hubble; TMA tiles at 512^2 and 1024^2: -1..-4 % -- profiles/README.md)
#define PAOS_ROW_512 PAOS_CASE(512, 8, 2, 2, PAOS_EXP_512_MINB, PAOS_EXP_512_MINB)
//                 N    E  Wrow Wcol minb(row) minb(col)
#define PAOS_TILE_TABLE \
    PAOS_CASE(64, 8, 16, 16, 1, 1) \
    PAOS_CASE(128, 8, 8, 8, 1, 1) \
    PAOS_CASE(256, 16, 4, 4, 2, 2) \
    PAOS_ROW_512 \
    PAOS_CASE(1024, 16, 2, 4, 4, 2)

#define PAOS_CASE(N, E, WR, WC, MR, MC)                                                              \
    case N:                                                                                          \
        return col ? launch_pass_t<double, N, E, WC, true, MC>(Ps, nb, tw1, tw2, st, device)                 \
                   : launch_pass_t<double, N, E, WR, false, MR>(Ps, nb, tw1, tw2, st, device);

cudaError_t launch_pass_c128(int n, bool col, bool wide, const PassParams* const* Ps, int nb, const void* tw1, const void* tw2,
                             cudaStream_t st, int device) {
    switch (n) {
        PAOS_TILE_TABLE
        case 2048:
            if (!col) return PAOS_ROW_2048(Ps, nb, tw1, tw2, st, device);
            return wide ? PAOS_COL_2048_WIDE(Ps, nb, tw1, tw2, st, device) : PAOS_COL_2048_NARROW(Ps, nb, tw1, tw2, st, device);
        case 4096:
            if (!col) return launch_pass_t<double, 4096, 16, 1, false, 2>(Ps, nb, tw1, tw2, st, device);
            return wide ? PAOS_COL_4096_WIDE(Ps, nb, tw1, tw2, st, device) : PAOS_COL_4096_NARROW(Ps, nb, tw1, tw2, st, device);
        default: return cudaErrorInvalidValue;
    }
}
#undef PAOS_CASE

#define PAOS_CASE(N, E, WR, WC, MR, MC) \
    case N:                             \
        return col ? WC : WR;
int tile_width_c128(int n, bool col, bool wide) {
    switch (n) {
        PAOS_TILE_TABLE
        case 2048: return !col ? PAOS_ROW_2048_W : (wide ? 2 : PAOS_EXP_COLW_2048);
        case 4096: return !col ? 1 : (wide ? 2 : PAOS_EXP_COLW_4096);
        default: return 1;
    }
}
#undef PAOS_CASE
}  // namespace paosb
