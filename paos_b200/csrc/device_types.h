// device_types.h -- plain structs shared by the host planner and the CUDA kernels.
#pragma once
#include <cuda.h>  // CUtensorMap (type only: the encoder is fetched from the driver at run time, nothing links libcuda)
#include <cuda_runtime.h>
#include <stdint.h>

// TMA tiles for the field loads / stores of the column passes (pass_kernel.cuh: use_tma_field): on by default for grids of
// PAOS_TMA_FIELD_MIN_N and up; -DPAOS_TMA_FIELD=0 builds the direct-access variant for the A/B measurement.
#ifndef PAOS_TMA_FIELD
#define PAOS_TMA_FIELD 1
#endif
#ifndef PAOS_TMA_FIELD_MIN_N
#define PAOS_TMA_FIELD_MIN_N 2048
#endif

namespace paosb {

constexpr int KMAX = 16;  // line FFTs chained in one pass kernel (= PAOS_MAX_CHAINED_FFTS of the public header)
constexpr int GMAX = 12;  // general (non-separable) factors applied by one pass kernel
constexpr int TERM_MAX = 8;
constexpr int TB_MAX = 64;  // tables built by one launch of the table builder (kernel parameters may be 32 KB on sm_70+)

// general elementwise factors (real masks, phase screens, device scalars)
enum GenKind : int {
    GEN_NONE = 0,
    GEN_ELLIPSE = 1,    // exact pixel/ellipse overlap (theta = 0): p0=xc p1=yc p2=1/a p3=1/b p4=a*b; ptr0 = edge table or null,
                        // valid for lines p7..p8
    GEN_RECT = 2,       // 32x32 sub-pixel rectangle: ptr0 = x counts, ptr1 = y counts (double[N])
    GEN_SCREEN = 3,     // exp(i*(2*pi*w)/wl): ptr0 = w (double[N*N]), p0 = wl
    GEN_SCALE_DEV = 4,  // multiply by *ptr0 (double in device memory)
    GEN_PSD = 5,        // PSD amplitude filter in frequency space (psd.py:121-129), see aux_kernels.cu
    GEN_ELLIPSE_TILT = 6,  // tilted ellipse: p0..p6 as GEN_ELLIPSE, p7 = cos(theta), p8 = sin(theta)
    GEN_RECT_TILT = 7      // tilted rectangle, 32x32 sub-pixel count per pixel: p0=xc p1=yc p2=w/2 p3=h/2 p7=cos p8=sin
};

struct GenOp {
    int kind;
    int pos;   // position in the pass program: 0 = before the first FFT, k = after the k-th FFT
    int flag;  // ellipse / rect: 1 = obscuration (use 1 - mask)
    int pad;
    double p0, p1, p2, p3, p4, p5, p6, p7, p8;
    const void* ptr0;
    const void* ptr1;
};

// Edge table of one elliptical mask on one pass (built by build_edge_tables_kernel, read by the pass kernel).  Along a line
// the pixels that are neither certainly inside nor certainly outside the ellipse form at most two short runs, where the
// line crosses the rim; their exact overlap fractions are evaluated once, one lane per pixel, instead of by a lane or two
// of the warp that meets them in the pass (a ~400-instruction dependent FP64 chain on the critical path of the whole line).
//   header[line] = {start0, start1, len0 | len1 << 8 | flag << 16, 0};  factor[(line * 2 + side) * EDGE_CAP + k]
// flag = 1: this line's runs do not fit (lines within a pixel or so of the ellipse's poles): the pass kernel computes them.
constexpr int EDGE_CAP = 12;
__host__ __device__ constexpr size_t edge_table_bytes(int n) { return (size_t)n * 16 + (size_t)n * 2 * EDGE_CAP * sizeof(double); }

struct PassParams {
    const void* src;  // null = field of ones (never materialised)
    void* dst;
    int nfft;
    int ngen;
    int dir[KMAX];             // +1 forward, -1 inverse
    const void* tab[KMAX + 1]; // along-line complex tables (null = none), applied after gen ops of that position
    double scl[KMAX + 1];      // real scale used when tab[k] is null (1.0 = nothing)
    const void* ctab_in;       // cross-axis complex table, one value per line, applied at position 0 (null = none)
    const void* ctab_out;      // same, applied after the last position
    unsigned genmask;          // bit p set: some general factor sits at position p
    unsigned sgnmask;          // bit p set: position p carries the sign (-1)^index (no table)
    int tile_lo, tile_hi;      // tiles outside this range end the pass exactly zero (aperture bounding box, or zero on input):
                               // they are not touched at all -- the planner remembers the band (virtual zeros)
    int in_lo, in_hi;          // along-line index range outside which the input is (virtually) zero: not loaded
    int readout;               // 0: store the complex field; PAOS_READ_* (1..3): store a real read-out into dst_real instead
    int zero_fill;             // 1: blank tiles store zeros into the field (diagnostic mode PAOS_ZERO_FILL=1)
    int tile_base;             // set by the launcher: tile of CTA 0 (blank tiles are not launched unless they have to store)
    int out_lo, out_hi;        // along-line index range the NEXT pass of the plan will read (its tile range: it runs along the
                               // other axis and touches nothing else); elements outside are not stored -- they stay stale
    int wide;                  // 1: this column pass runs on the wide-tile kernel (pass_dispatch.h) and its tiles count in that width
    const void* tmap_host;     // host pointer to the CUtensorMap of `dst` (column tiles of W complex x 256 rows), or null; the
                               // launcher copies it into BatchParams::tmap (a tensor map must sit in kernel parameter space)
    const void* tmap_real_host;  // same for `dst_real` (tiles of W reals x 256 rows) when a read-out is fused into the pass
    void* dst_real;
    GenOp gen[GMAX];
};

// One launch of the pass kernel covers the same-axis passes of up to BMAX independent wavefronts (a batch of wavelengths,
// fields or Monte-Carlo realizations that are at the same point of their chains): CTA c belongs to item b with
// start[b] <= c < start[b+1] and works on tile  c - start[b] + p[b].tile_base  of that item's field.  The whole block
// travels as a __grid_constant__ kernel parameter (up to 32764 bytes on sm_70+).
// Two capacities are instantiated: CAP = 1 (a single wavefront: 1.7 KB of parameters) and CAP = BMAX.
constexpr int BMAX = 16;
template <int CAP> struct BatchParams {
    int nb;
    int use_tmap;  // bit 0: every item carries a tensor map of its field (column kernels: TMA tile loads / stores);
                   // bit 1: every item carries one of its read-out destination as well
    int start[CAP + 2];
    PassParams p[CAP];
    CUtensorMap tmap[CAP];
    CUtensorMap tmap_real[CAP];
};
static_assert(sizeof(BatchParams<BMAX>) <= 32764, "BatchParams must fit the kernel parameter space");

// one separable phase term: exp(i * c1*c2 * u^2), u = (k - n/2) * d  (QSPACE)  or  (k - n/2) * (1/(n*d))  (QFREQ)
// TERM_COUNT: real factor count(k)/32, count = sub-pixel centres of pixel k inside a rectangle side (c1 = centre, c2 = full side)
enum TermKind : int { TERM_QSPACE = 1, TERM_QFREQ = 2, TERM_COUNT = 3 };
struct TableTerm {
    int kind;
    int pad;
    double c1, c2, d;
};

enum TableKind : int { TABLE_PHASE = 0, TABLE_COUNT = 1 };
struct TableSpec {
    void* out;
    int kind;      // TABLE_PHASE: complex table; TABLE_COUNT: double table of sub-pixel counts
    int nterms;
    int sign;      // multiply by (-1)^k
    int pad;
    double scale;  // real scale folded in
    double cnt_c, cnt_full;  // TABLE_COUNT: aperture centre (pixels) and full side (pixels)
    TableTerm terms[TERM_MAX];
};
struct TableBlock {
    int ntab;
    int n;
    int dtype;  // 0: complex128 tables, 1: complex64 tables (count tables are always double)
    int pad;
    TableSpec spec[TB_MAX];
};

// one edge table to build: the mask, the axis of the pass that applies it and the threads per line of that pass
struct EdgeSpec {
    GenOp g;
    void* out;
    int col;  // 1: lines are columns (line = ix, along = iy)
    int T;    // threads per line of the pass kernel: element idx = t + j * T, classified as the kernel does
    int line_lo, line_hi;  // lines that can hold rim pixels (the ellipse's extent across the lines, two lines of slack); the
                           // same bounds travel in g.p7, g.p8 of the GenOp the pass kernel gets, which looks up nothing outside
};
constexpr int EB_MAX = 64;
struct EdgeBlock {
    int n;
    int nspec;
    EdgeSpec spec[EB_MAX];
};

constexpr int ZERN_MAX = 64;
struct ZernParams {
    int K;
    int origin;  // 0 = 'x', 1 = 'y'
    int n;
    int accumulate;  // 1: add to existing out
    double radius, dx, dy, cos_off, sin_off;
    int m[ZERN_MAX];
    int nn[ZERN_MAX];
    double coef[ZERN_MAX];   // Z[k]*norm[k]*binom(k_r+|m|, k_r)*(-1)^k_r
};

}  // namespace paosb
