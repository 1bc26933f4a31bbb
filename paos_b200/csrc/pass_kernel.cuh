// pass_kernel.cuh -- the line-pass kernel: one kernel = one sweep of the wavefront through HBM.
//
// A pass loads every line (row or column) of the N x N complex field once, runs a small *program* on it
// in registers -- diagonal factors and up to KMAX line FFTs -- and stores it once.  Because the Fresnel
// chirps, the lens phase and the fftshift signs are separable (exp(i c (x^2+y^2)) = exp(i c x^2) *
// exp(i c y^2)), a run of FFT2s with only separable factors between them factorises into ONE row pass
// and ONE column pass; non-separable factors (aperture masks, phase screens, the stop scalar) are applied
// as "general ops" between two chained line transforms.  The host planner (runtime.cu) builds the program.
//
// Replaces the array arithmetic of paos/classes/wfo.py:200-201 (make_stop scale), :273-276 (aperture),
// :359-366 (lens), :462-472 (ptp), :493-509 (stw), :530-545 (wts), :650-652/:867-869/:947 (phase screens).
#pragma once
#include <cstdlib>
#include <cstring>
#include <memory>

#include "device_types.h"
#include "fft_core.cuh"

namespace paosb {

// ---- exact area of the unit disk inside an axis-aligned rectangle ----------------------------------
// First-quadrant piece 0 <= x0 <= x1, 0 <= y0 <= y1 :  integral over x of clamp(sqrt(1-x^2), y0, y1) - y0.
__device__ __forceinline__ double disk_rect_q1(double x0, double x1, double y0, double y1) {
    if (x1 <= x0 || y1 <= y0) return 0.0;
    if (x0 * x0 + y0 * y0 >= 1.0) return 0.0;
    if (x1 * x1 + y1 * y1 <= 1.0) return (x1 - x0) * (y1 - y0);
    const double xa = (y1 < 1.0) ? sqrt(fmax(0.0, 1.0 - y1 * y1)) : 0.0;  // arc crosses the top edge
    const double xb = (y0 < 1.0) ? sqrt(fmax(0.0, 1.0 - y0 * y0)) : 0.0;  // arc crosses the bottom edge
    const double L = fmin(fmax(xa, x0), x1);
    const double U = fmin(fmax(xb, x0), x1);
    double area = (y1 - y0) * (L - x0);
    if (U > L) {
        const double sL = sqrt(fmax(0.0, 1.0 - L * L));
        const double sU = sqrt(fmax(0.0, 1.0 - U * U));
        // area under the arc = trapezoid under the chord + circular segment
        const double sn = U * sL - L * sU;  // sin of the angle th between the two radius vectors
        const double cs = L * U + sL * sU;
        double seg;  // th - sin(th)
        if (sn < 0.05 && cs > 0.0) {
            // A pixel subtends a small angle (1/a rad for a semi-axis of a pixels), so this is the branch an edge pixel
            // takes: th = asin(sn), th - sn = sn^3/6 + 3 sn^5/40 + 15 sn^7/336 + 105 sn^9/3456 + 945 sn^11/42240 + ...
            // (next term < 1e-14 of the sum at sn = 0.05): no cancellation and no atan2, whose ~150 dependent FP64
            // instructions on one or two lanes of a warp used to stand on the critical path of the whole line.
            const double s2 = sn * sn;
            seg = sn * s2 * (1.0 / 6.0 + s2 * (3.0 / 40.0 + s2 * (15.0 / 336.0 + s2 * (105.0 / 3456.0 + s2 * (945.0 / 42240.0)))));
        } else {
            seg = atan2(sn, cs) - sn;
        }
        area += (0.5 * (sL + sU) - y0) * (U - L) + 0.5 * seg;
    }
    return area;
}

static __device__ __noinline__ double disk_rect_area(double u0, double u1, double v0, double v1) {
    // split at the axes and fold every piece into the first quadrant
    double xa0 = fmax(u0, 0.0), xa1 = u1;            // x >= 0 part
    double xb0 = fmax(-u1, 0.0), xb1 = -u0;          // x <= 0 part, mirrored
    double ya0 = fmax(v0, 0.0), ya1 = v1;
    double yb0 = fmax(-v1, 0.0), yb1 = -v0;
    double a = 0.0;
    if (xa1 > xa0) {
        if (ya1 > ya0) a += disk_rect_q1(xa0, xa1, ya0, ya1);
        if (yb1 > yb0) a += disk_rect_q1(xa0, xa1, yb0, yb1);
    }
    if (xb1 > xb0) {
        if (ya1 > ya0) a += disk_rect_q1(xb0, xb1, ya0, ya1);
        if (yb1 > yb0) a += disk_rect_q1(xb0, xb1, yb0, yb1);
    }
    return a;
}

// area fraction of the unit pixel centred on (ix, iy) inside the ellipse (theta = 0)
__device__ __forceinline__ double ellipse_fraction(const GenOp& g, double ix, double iy) {
    const double px = ix - g.p0, py = iy - g.p1;
    const double u0 = (px - 0.5) * g.p2, u1 = (px + 0.5) * g.p2;
    const double v0 = (py - 0.5) * g.p3, v1 = (py + 0.5) * g.p3;
    const double uf = fmax(fabs(u0), fabs(u1)), vf = fmax(fabs(v0), fabs(v1));
    if (uf * uf + vf * vf <= 1.0) return 1.0;
    const double un = (u0 <= 0.0 && u1 >= 0.0) ? 0.0 : fmin(fabs(u0), fabs(u1));
    const double vn = (v0 <= 0.0 && v1 >= 0.0) ? 0.0 : fmin(fabs(v0), fabs(v1));
    if (un * un + vn * vn >= 1.0) return 0.0;
    double f = disk_rect_area(u0, u1, v0, v1) * g.p4;
    return fmin(fmax(f, 0.0), 1.0);
}

// FP32 classification of the element idx = t + j*T of a line against an elliptical mask, shared by the pass kernel and the
// builder of the edge tables (explicit fused operations: both must round identically).  Returns r^2 in the frame where the
// ellipse is the unit circle; a0 = (along-coordinate of element t minus the centre) * scale, da = T * scale, b2 = (cross
// coordinate of the line minus the centre)^2 * scale^2.
__device__ __forceinline__ float ellipse_a0(int t, double c_along, float s_along) { return __fmul_rn((float)((double)t - c_along), s_along); }
__device__ __forceinline__ float ellipse_b2(int line, double c_cross, float s_cross) {
    const float bq = __fmul_rn((float)((double)line - c_cross), s_cross);
    return __fmul_rn(bq, bq);
}
__device__ __forceinline__ float ellipse_r2(float a0, float da, int j, float b2) {
    const float aq = __fmaf_rn((float)j, da, a0);
    return __fmaf_rn(aq, aq, b2);
}

// PSD amplitude filter sqrt(psd2d)*sqrt(N*N), zero outside [fmin, fmax] (psd.py:118-129, wfo.py:913-918).
// p0=A p1=B p2=C p3=fknee p4=fmin p5=fmax p6=valx p7=valy p8=N ; frequencies in unshifted (fftfreq) order
static __device__ __noinline__ double psd_filter(const GenOp& g, int ix, int iy) {
    const int n = (int)g.p8;
    const double kx = (double)(ix < n / 2 ? ix : ix - n), ky = (double)(iy < n / 2 ? iy : iy - n);
    const double fx = kx * g.p6, fy = ky * g.p7;
    double f = sqrt(fx * fx + fy * fy);
    if (f == 0.0) f = 1e-100;
    if (f < g.p4 || f > g.p5) return 0.0;
    const double dfx = 2.0 * g.p6 - 1.0 * g.p6, dfy = 2.0 * g.p7 - 1.0 * g.p7;  // f[0,2]-f[0,1], f[2,0]-f[1,0]
    const double psd2d = g.p0 / (g.p1 + pow(f / g.p3, g.p2)) / (2.0 * 3.141592653589793 * f) * (dfx * dfy);
    return sqrt(psd2d) * sqrt((double)n * (double)n);
}

// |.|, angle or |.|^2 of one element (what = PAOS_READ_AMPLITUDE / _PHASE / _PSF)
template <typename R> __device__ __noinline__ R readout_value(C<R> v, int what) {
    if (what == 1) return (R)hypot((double)v.x, (double)v.y);
    if (what == 2) return (R)atan2((double)v.y, (double)v.x);
    return v.x * v.x + v.y * v.y;
}

// ---- tilted shapes (wfo.py:243-268 with tilt != None; never produced by paos.core.run) ------------------------
// signed area of (triangle O,p,q) intersected with the unit disc; the sum over the edges of a counter-clockwise
// polygon is the area of polygon x disc
static __device__ __noinline__ double edge_disk_area(double px, double py, double qx, double qy) {
    const double dx = qx - px, dy = qy - py;
    const double a = dx * dx + dy * dy, b = px * dx + py * dy, c = px * px + py * py - 1.0;
    const double disc = b * b - a * c;
    const double full = 0.5 * atan2(px * qy - py * qx, px * qx + py * qy);  // pure sector
    if (!(disc > 0.0) || !(a > 0.0)) return full;
    const double sq = sqrt(disc), t1 = (-b - sq) / a, t2 = (-b + sq) / a;
    if (!(t2 > 0.0 && t1 < 1.0)) return full;
    const double ta = fmin(fmax(t1, 0.0), 1.0), tb = fmin(fmax(t2, 0.0), 1.0);
    const double ax = px + ta * dx, ay = py + ta * dy, bx = px + tb * dx, by = py + tb * dy;
    double area = 0.5 * (ax * by - ay * bx);
    if (t1 > 0.0) area += 0.5 * atan2(px * ay - py * ax, px * ax + py * ay);
    if (t2 < 1.0) area += 0.5 * atan2(bx * qy - by * qx, bx * qx + by * qy);
    return area;
}

static __device__ __noinline__ double tilted_ellipse_fraction(const GenOp& g, double ix, double iy) {
    const double ct = g.p7, st = -g.p8;  // rotate by -theta into the ellipse frame
    const double cx = ix - g.p0, cy = iy - g.p1;
    const double u0 = (cx * ct - cy * st) * g.p2, v0 = (cx * st + cy * ct) * g.p3;
    const double r2 = u0 * u0 + v0 * v0;
    if (r2 <= g.p5) return 1.0;
    if (r2 >= g.p6) return 0.0;
    double ux[4], uy[4];
    const double ox[4] = {-0.5, 0.5, 0.5, -0.5}, oy[4] = {-0.5, -0.5, 0.5, 0.5};
    for (int k = 0; k < 4; ++k) {
        const double x = cx + ox[k], y = cy + oy[k];
        ux[k] = (x * ct - y * st) * g.p2;
        uy[k] = (x * st + y * ct) * g.p3;
    }
    double area = 0.0;
    for (int k = 0; k < 4; ++k) area += edge_disk_area(ux[k], uy[k], ux[(k + 1) & 3], uy[(k + 1) & 3]);
    double f = area * g.p4;
    if (f < 1e-14) f = 0.0;  // a pixel that only grazes the ellipse: the published routine returns exactly 0 / 1
    if (f > 1.0 - 1e-14) f = 1.0;
    return f;
}

static __device__ __noinline__ double tilted_rect_fraction(const GenOp& g, double ix, double iy) {
    const double ct = g.p7, st = g.p8;
    const double x0 = (ix - 0.5) - g.p0, y0 = (iy - 0.5) - g.p1;
    // quick decision from the pixel centre: the pixel spans at most h in either rectangle axis
    const double xm = x0 + 0.5, ym = y0 + 0.5, h = 0.5 * (fabs(ct) + fabs(st));
    const double xt0 = fabs(ym * st + xm * ct), yt0 = fabs(ym * ct - xm * st);
    if (xt0 + h < g.p2 && yt0 + h < g.p3) return 1.0;
    if (xt0 - h >= g.p2 || yt0 - h >= g.p3) return 0.0;
    // 32 x 32 sub-pixel centres, accumulated like the published routine (x = x0 - d/2; x += d; same for y)
    const double d = 1.0 / 32.0;
    int cnt = 0;
    double x = x0 - 0.5 * d;
    for (int i = 0; i < 32; ++i) {
        x += d;
        double y = y0 - 0.5 * d;
        for (int j = 0; j < 32; ++j) {
            y += d;
            const double xt = y * st + x * ct, yt = y * ct - x * st;
            if (fabs(xt) < g.p2 && fabs(yt) < g.p3) ++cnt;
        }
    }
    return (double)cnt / 1024.0;
}

// Slow path of one general factor at pixel (ix, iy): exact edge-pixel overlap, phase screens, the PSD filter.
// Out of line on purpose: this is cold code next to the line FFT, and inlining it per register element made
// the kernel 250 KB of SASS.  Returns the complex factor (fr, fi).
static __device__ __noinline__ void gen_factor_slow(const GenOp& g, int ix, int iy, int n, double& fr, double& fi) {
    double re = 1.0, im = 0.0;
    switch (g.kind) {
        case GEN_ELLIPSE: {
            double m = ellipse_fraction(g, (double)ix, (double)iy);
            re = g.flag ? 1.0 - m : m;
        } break;
        case GEN_SCREEN: {
            const double w = __ldg((const double*)g.ptr0 + (size_t)iy * n + ix);
            if (w != 0.0) sincos((6.283185307179586 * w) / g.p0, &im, &re);
        } break;
        case GEN_PSD: re = psd_filter(g, ix, iy); break;
        default: break;
    }
    fr = re;
    fi = im;
}

// tilted shapes, kept apart so that the edge-pixel path above stays small
static __device__ __noinline__ double gen_factor_tilt(const GenOp& g, int ix, int iy) {
    const double m = g.kind == GEN_ELLIPSE_TILT ? tilted_ellipse_fraction(g, (double)ix, (double)iy)
                                                : tilted_rect_fraction(g, (double)ix, (double)iy);
    return g.flag ? 1.0 - m : m;
}

// single-op version used by the stop reduction (aux_kernels.cu)
template <typename R>
__device__ __forceinline__ void apply_gen(C<R>& v, const GenOp& g, int ix, int iy, int n) {
    double re = 1.0, im = 0.0;
    switch (g.kind) {
        case GEN_ELLIPSE: {
            // same FP32 interior / exterior classification as the pass kernel; exact routine only in the edge band
            // centre subtracted in double: (float)centre alone would carry up to 1.2e-4 px at n = 4096, more than the margin
            // allows for semi-axes below ~25 px
            const float uq = (float)((double)ix - g.p0) * (float)g.p2, wq = (float)((double)iy - g.p1) * (float)g.p3;
            const float r2 = uq * uq + wq * wq;
            double m;
            bool from_table = false;
            if (r2 <= (float)g.p5 - 1e-5f) m = 1.0;
            else if (r2 >= (float)g.p6 + 1e-5f) m = 0.0;
            else {
                m = 0.0;
                if (g.ptr0 && iy >= (int)g.p7 && iy <= (int)g.p8) {  // row-axis edge table built for this reduction (EdgeSpec); holds the final factor
                    const int4 hdr = __ldg(reinterpret_cast<const int4*>(g.ptr0) + iy);
                    const double* fac = reinterpret_cast<const double*>(reinterpret_cast<const char*>(g.ptr0) + (size_t)n * 16) +
                                        (size_t)iy * 2 * EDGE_CAP;
                    const unsigned k0 = (unsigned)(ix - hdr.x), k1 = (unsigned)(ix - hdr.y);
                    if (k0 < (unsigned)(hdr.z & 0xff)) {
                        m = __ldg(fac + k0);
                        from_table = true;
                    } else if (k1 < (unsigned)((hdr.z >> 8) & 0xff)) {
                        m = __ldg(fac + EDGE_CAP + k1);
                        from_table = true;
                    }
                }
                if (!from_table) m = ellipse_fraction(g, (double)ix, (double)iy);
            }
            re = from_table ? m : (g.flag ? 1.0 - m : m);
        } break;
        case GEN_RECT: {
            const double cx = __ldg((const double*)g.ptr0 + ix), cy = __ldg((const double*)g.ptr1 + iy);
            double m = (cy * cx) / 1024.0;
            re = g.flag ? 1.0 - m : m;
        } break;
        case GEN_SCALE_DEV: re = __ldg((const double*)g.ptr0); break;
        case GEN_ELLIPSE_TILT:
        case GEN_RECT_TILT: re = gen_factor_tilt(g, ix, iy); break;
        default: break;
    }
    (void)im;
    v = v * (R)re;
}

// ---- TMA staging of the along-line phase tables (experiment -DPAOS_TMA_TABLES) ----------------------------------
// One elected thread asks the bulk-copy engine (cp.async.bulk, 1-D TMA) for the N-entry table of the next position
// while the current line FFT runs; completion is signalled on an mbarrier in shared memory, and the threads then read
// their 16 entries with LDS.128 instead of 16 LDG.128 through L1/L2.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, void* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Where it is used: the column kernels of the 2048^2 and 4096^2 grids, whose two CTAs per SM have the 32-64 KB to spare
// (measured on AIRS-CH0 2048^2, batches of 8: column passes 2-5 % faster, 521 -> 495 us for the dense final pass).  The row
// kernel (one line per CTA, four CTAs per SM) would drop to three CTAs per SM and loses what it gains (profiles/README.md).
// -DPAOS_TMA_TABLES=1 / =0 forces it on / off everywhere for the A/B measurement.
// A single-column CTA (W = 1: the 2048^2 column passes without a read-out, four CTAs per SM like the row kernel) has no room
// for the staging buffer either.
template <int N, bool COL, int W> __host__ __device__ constexpr bool use_tma_tables() {
#ifdef PAOS_TMA_TABLES
    return PAOS_TMA_TABLES != 0;
#else
    return COL && N >= 2048 && W >= 2;
#endif
}

// ---- TMA tiles for the column passes' field stores -------------------------------------------------------------
// A column pass owns W adjacent columns: in global memory that is a tile of W complex values (32 B) per row.  Stored
// straight from registers, a warp's 16-byte stores touch 16 different 128-byte lines per instruction (16 data-pipe
// wavefronts instead of the 4 of a row pass).  With a tensor map of the field the CTA instead lays the finished tile out in
// the (now idle) exchange buffer -- conflict-free 16-byte shared stores -- and one thread hands it to the TMA unit
// (cp.async.bulk.tensor.2d, boxes of W x 256 rows), which writes it while the CTA retires; on entry the live rows of the
// tile come in the same way (boxes landing in the exchange buffer behind an mbarrier, then conflict-free shared loads).
// Measured on AIRS-CH0 2048^2, batches of 8: column passes 4-7 % faster (column x4: 302 -> 280 us), sweep 1 884 -> 1 954 PSF/s.
template <int N, bool COL> __host__ __device__ constexpr bool use_tma_field() {
    return COL && PAOS_TMA_FIELD != 0 && N >= PAOS_TMA_FIELD_MIN_N;
}
constexpr int TMA_BOX_ROWS = 256;
#ifndef PAOS_TMA_FIELD_LOAD  // A/B switches: tile loads / tile stores of the complex field through the TMA unit
#define PAOS_TMA_FIELD_LOAD 1
#endif
#ifndef PAOS_TMA_FIELD_STORE
#define PAOS_TMA_FIELD_STORE 1
#endif
__device__ __forceinline__ void tma_load_tile(void* smem_dst, const CUtensorMap* tm, int c0, int c1, void* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1),
                 "r"(smem_u32(smem_src))
                 : "memory");
}

// -x by its sign bit (bit-identical to x * -1, but not an instruction of the FP64 pipe)
__device__ __forceinline__ double flip_sign(double x) { return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x)); }
__device__ __forceinline__ float flip_sign(float x) { return __int_as_float(__float_as_int(x) ^ (int)0x80000000); }

// ---- the pass kernel ---------------------------------------------------------------------------------
// R: real type; N: line length; E: points per thread; W: lines per CTA; COL: lines are columns.
// zero store of one tile (and of its read-out); out of line so that it does not share registers with the main path
template <typename R, int N, int E, int W, bool COL>
__device__ __noinline__ void store_zero_tile(const PassParams& P, int tile) {
    constexpr int T = N / E;
    const int tid = threadIdx.x;
    const int w = COL ? (tid % W) : (tid / T);
    const int t = COL ? (tid / W) : (tid % T);
    const int line = tile * W + w;
    C<R>* dst = reinterpret_cast<C<R>*>(P.dst);
    R* out = reinterpret_cast<R*>(P.dst_real);
    const C<R> zero((R)0, (R)0);
    for (int j = 0; j < E; ++j) {
        const int idx = t + j * T;
        const size_t ga = COL ? ((size_t)idx * N + line) : ((size_t)line * N + idx);
        if (P.zero_fill && dst) stc_stream(dst + ga, zero);
        if (P.readout) out[ga] = (R)0;  // |0|, angle(0), |0|^2
    }
}

template <typename R, int N, int E, int W, bool COL, int MINB, int CAP>
__global__ void __launch_bounds__(W*(N / E), MINB)
    pass_kernel(const __grid_constant__ BatchParams<CAP> BP, const C<R>* __restrict__ tw1, const C<R>* __restrict__ tw2) {
    using G = LineGeom<N, E>;
    constexpr int T = G::T;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C<R>* smem = reinterpret_cast<C<R>*>(smem_raw);

    // which wavefront of the batch this CTA works for (CTA-uniform; the parameter block sits in the constant bank)
    int b = 0;
    if constexpr (CAP > 1) {
#pragma unroll 1
        while (b + 1 < BP.nb && (int)blockIdx.x >= BP.start[b + 1]) ++b;
    }
    const PassParams& P = BP.p[b];

    const int tid = threadIdx.x;
    const int w = COL ? (tid % W) : (tid / T);
    const int t = COL ? (tid / W) : (tid % T);
    const int tile = (int)blockIdx.x - BP.start[b] + P.tile_base;
    const int line = tile * W + w;
    C<R>* sm = smem + w * G::line_stride(COL ? W : 1, (int)sizeof(C<R>));
    auto sync = [] { __syncthreads(); };
    constexpr bool kTmaTables = use_tma_tables<N, COL, W>();
    // TMA variant: [exchange buffers of the W lines][one table of N entries][mbarrier]
    C<R>* tabbuf = smem + W * G::line_stride(COL ? W : 1, (int)sizeof(C<R>));
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(kTmaTables ? tabbuf + N : tabbuf);
    unsigned tab_phase = 0;
    constexpr bool kTmaField = use_tma_field<N, COL>();
    constexpr bool kInplace = PAOS_INPLACE_EXCHANGE != 0 && G::R2 > 1;  // two barriers per transform (fft_core.cuh)
    int flip = 0;
    if constexpr (kTmaTables || kTmaField) {
        if (tid == 0) {
            mbar_init(mbar, 1);      // phase table of the next position
            mbar_init(mbar + 1, 1);  // tile of the field on entry
        }
        __syncthreads();
    }

    const C<R>* src = reinterpret_cast<const C<R>*>(P.src);
    C<R>* dst = reinterpret_cast<C<R>*>(P.dst);

    // Programmatic dependent launch: this grid may have been started while its predecessor in the stream (the previous
    // pass, which wrote the field; the table builder; the stop reduction) was still draining.  Everything above is
    // index arithmetic; nothing below may run before the predecessors have completed and flushed their writes.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // Tiles outside [tile_lo, tile_hi] are blanked by an elliptical aperture somewhere in this pass (the planner
    // works the range out from the apertures' bounding boxes) or are zero on input: such a tile is zero at the end of
    // the pass whatever happens before the mask, so it is neither loaded, transformed nor stored -- the planner
    // remembers the band outside which the field is zero ("virtual zeros") and hands it to the next consumer as
    // [in_lo, in_hi] (other axis) or folds it into its tile range (same axis).  Only a fused read-out needs the zeros.
    if (tile < P.tile_lo || tile > P.tile_hi) {
        if (P.zero_fill | P.readout) store_zero_tile<R, N, E, W, COL>(P, tile);
        return;
    }

    if constexpr (kTmaTables) {
        if (tid == 0 && P.tab[0]) bulk_load(tabbuf, P.tab[0], (unsigned)(N * sizeof(C<R>)), mbar);
    }
    C<R> v[E];
    bool loaded = false;
    if constexpr (kTmaField) {
        if (src && (BP.use_tmap & 1) && PAOS_TMA_FIELD_LOAD) {
            // the live rows [in_lo, in_hi] of this CTA's W columns arrive as TMA boxes of W x 256 rows in the exchange buffer
            // (idle until the first transform); the threads then pick their elements with conflict-free shared loads
            const int base = P.in_lo, rows = P.in_hi < N ? P.in_hi - base + 1 : 0;  // empty band: in_lo = in_hi = N
            if (rows > 0) {
                const int boxes = (rows + TMA_BOX_ROWS - 1) / TMA_BOX_ROWS;
                if (tid == 0) {
                    constexpr int REALS = (int)(sizeof(C<R>) / sizeof(R));
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar + 1)),
                                 "r"((unsigned)(boxes * TMA_BOX_ROWS * W * (int)sizeof(C<R>)))
                                 : "memory");
#pragma unroll 1
                    for (int k = 0; k < boxes; ++k)
                        tma_load_tile(smem + (size_t)k * TMA_BOX_ROWS * W, &BP.tmap[b], tile * W * REALS, base + k * TMA_BOX_ROWS, mbar + 1);
                }
                mbar_wait(mbar + 1, 0);
            }
            const unsigned span = (unsigned)(P.in_hi - P.in_lo);
#pragma unroll
            for (int j = 0; j < E; ++j) {
                const int idx = t + j * T;
                v[j] = ((unsigned)(idx - P.in_lo) <= span && rows > 0) ? ldc(smem + (size_t)(idx - base) * W + w) : C<R>((R)0, (R)0);
            }
            loaded = true;
            // the in-place transforms write the exchange buffer without a barrier of their own in front
            if constexpr (kInplace) __syncthreads();
        }
    }
    if (loaded) {
    } else if (src) {
        // memory outside [in_lo, in_hi] is stale: those elements are zeros that were never written
        const unsigned span = (unsigned)(P.in_hi - P.in_lo);
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int idx = t + j * T;
            const size_t ga = COL ? ((size_t)idx * N + line) : ((size_t)line * N + idx);
            v[j] = ((unsigned)(idx - P.in_lo) <= span) ? ldc_stream(src + ga) : C<R>((R)0, (R)0);
        }
    } else {
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] = C<R>((R)1, (R)0);
    }

    if (P.ctab_in) {
        const C<R> c = ldc_ro(reinterpret_cast<const C<R>*>(P.ctab_in) + line);
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] = v[j] * c;
    }
    for (int pos = 0;; ++pos) {
        // diagonal factors of this position: general (masks, screens, stop scalar), then the along-line table
#ifndef PAOS_EXP_NO_GEN
        if (P.genmask >> pos & 1) {
            for (int gi = 0; gi < P.ngen; ++gi) {
                const GenOp& g = P.gen[gi];
                if (g.pos != pos) continue;
                if (g.kind == GEN_ELLIPSE) {
                    // Interior / exterior pixels are classified in FP32 (its pipe is idle next to the FP64 butterflies):
                    // p5, p6 = squared radii, in the frame where the ellipse is the unit circle, inside / outside which
                    // a whole pixel is certainly inside / outside; the 1e-5 margins cover the float rounding (the centre
                    // is subtracted in double: (float)centre alone would carry up to 1.2e-4 px at n = 4096).  The pixels
                    // of the thin band between them take their factor from the edge table built for this pass (one
                    // load), or the exact (double) routine on the few lines the table does not cover.
                    const float s_al = COL ? (float)g.p3 : (float)g.p2, s_cr = COL ? (float)g.p2 : (float)g.p3;
                    const float in5 = (float)g.p5 - 1e-5f, out6 = (float)g.p6 + 1e-5f;
                    const float a0 = ellipse_a0(t, COL ? g.p1 : g.p0, s_al);
                    const float da = __fmul_rn((float)T, s_al);
                    const float b2 = ellipse_b2(line, COL ? g.p0 : g.p1, s_cr);
                    const bool obsc = g.flag != 0;
                    unsigned emask = 0;
#pragma unroll
                    for (int j = 0; j < E; ++j) {
                        const float r2 = ellipse_r2(a0, da, j, b2);
                        if (r2 <= in5) {
                            if (obsc) v[j] = C<R>((R)0, (R)0);
                        } else if (r2 >= out6) {
                            if (!obsc) v[j] = C<R>((R)0, (R)0);
                        } else {
                            emask |= 1u << j;
                        }
                    }
#ifndef PAOS_EXP_NO_EDGE
                    if (emask) {
                        int4 hdr = make_int4(0, 0, 1 << 16, 0);
                        const double* fac = nullptr;
                        if (g.ptr0 && line >= (int)g.p7 && line <= (int)g.p8) {
                            hdr = __ldg(reinterpret_cast<const int4*>(g.ptr0) + line);
                            fac = reinterpret_cast<const double*>(reinterpret_cast<const char*>(g.ptr0) + (size_t)N * 16) +
                                  (size_t)line * 2 * EDGE_CAP;
                        }
                        while (emask) {
                            const int je = __ffs((int)emask) - 1;
                            emask &= emask - 1;
                            const int idx = t + je * T;
                            double m;
                            const unsigned k0 = (unsigned)(idx - hdr.x), k1 = (unsigned)(idx - hdr.y);
                            if (k0 < (unsigned)(hdr.z & 0xff)) m = __ldg(fac + k0);
                            else if (k1 < (unsigned)((hdr.z >> 8) & 0xff)) m = __ldg(fac + EDGE_CAP + k1);
                            else {
                                double fi;
                                gen_factor_slow(g, COL ? line : idx, COL ? idx : line, N, m, fi);
                            }
#pragma unroll
                            for (int j = 0; j < E; ++j)
                                if (j == je) v[j] = v[j] * (R)m;
                        }
                    }
#endif
                } else if (g.kind == GEN_RECT) {
#pragma unroll
                    for (int j = 0; j < E; ++j) {
                        const int idx = t + j * T;
                        const double cx = __ldg((const double*)g.ptr0 + (COL ? line : idx));
                        const double cy = __ldg((const double*)g.ptr1 + (COL ? idx : line));
                        double m = (cy * cx) / 1024.0;
                        if (g.flag) m = 1.0 - m;
                        v[j] = v[j] * (R)m;
                    }
                } else if (g.kind == GEN_SCALE_DEV) {
                    const R m = (R)__ldg((const double*)g.ptr0);
#pragma unroll
                    for (int j = 0; j < E; ++j) v[j] = v[j] * m;
                } else if (g.kind == GEN_ELLIPSE_TILT || g.kind == GEN_RECT_TILT) {
#pragma unroll
                    for (int j = 0; j < E; ++j) {
                        const int idx = t + j * T;
                        v[j] = v[j] * (R)gen_factor_tilt(g, COL ? line : idx, COL ? idx : line);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < E; ++j) {
                        const int idx = t + j * T;
                        double fr, fi;
                        gen_factor_slow(g, COL ? line : idx, COL ? idx : line, N, fr, fi);
                        v[j] = v[j] * C<R>((R)fr, (R)fi);
                    }
                }
            }
        }
#endif
#ifdef PAOS_EXP_NO_TAB
        const C<R>* tab = nullptr;
#else
        const C<R>* tab = reinterpret_cast<const C<R>*>(P.tab[pos]);
#endif
        if (tab) {
            if constexpr (kTmaTables) {
                mbar_wait(mbar, tab_phase);
                tab_phase ^= 1u;
#pragma unroll
                for (int j = 0; j < E; ++j) v[j] = v[j] * ldc(tabbuf + t + j * T);
            } else {
#pragma unroll
                for (int j = 0; j < E; ++j) v[j] = v[j] * ldc_ro(tab + t + j * T);
            }
        } else {
            // real scale (the planner moves it into a table of the pass whenever there is one), and the (-1)^index sign of an
            // fftshift when flagged (index parity = t parity: T is even) as a flip of the sign bits on the integer pipe
            const R s = (R)P.scl[pos];
            if (s != (R)1) {
#pragma unroll
                for (int j = 0; j < E; ++j) v[j] = v[j] * s;
            }
            if ((P.sgnmask >> pos & 1) && (t & 1)) {
#pragma unroll
                for (int j = 0; j < E; ++j) v[j] = C<R>(flip_sign(v[j].x), flip_sign(v[j].y));
            }
        }
        if (pos == P.nfft) break;
        // one forward line transform of the registers (inlined into both direction branches below)
        auto transform = [&] {
            if constexpr (kTmaTables) {
                // the copy engine fetches the table of the next position while this transform runs; the buffer is free once
                // every thread has consumed the staged table, i.e. after any barrier that follows the multiply above
                auto next_table = [&] {
                    if (tid == 0 && P.tab[pos + 1]) bulk_load(tabbuf, P.tab[pos + 1], (unsigned)(N * sizeof(C<R>)), mbar);
                };
                if constexpr (kInplace) {
                    line_fft_fwd<G, R, decltype(sync), false, decltype(next_table)>(v, t, sm, tw1, tw2, sync, flip, next_table);
                } else {
                    __syncthreads();
                    next_table();
                    line_fft_fwd<G, R, decltype(sync), false>(v, t, sm, tw1, tw2, sync);
                }
            } else {
                // pull the next position's phase table into L1 while this transform runs (one 128-byte line per
                // prefetch; the table is N complex values, shared by every line of the pass)
                const char* nxt = reinterpret_cast<const char*>(P.tab[pos + 1]);
                if (nxt) {
                    constexpr int LINES = N * (int)sizeof(C<R>) / 128;
#pragma unroll
                    for (int i = t; i < LINES; i += T) asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt + (size_t)i * 128));
                }
                line_fft_fwd<G, R>(v, t, sm, tw1, tw2, sync, flip);
            }
        };
        // the inverse transform is the forward one on swapped (re, im); with the direction known at compile time
        // inside each branch the swaps are register renaming, not moves
        if (P.dir[pos] < 0) {
#pragma unroll
            for (int j = 0; j < E; ++j) v[j] = C<R>(v[j].y, v[j].x);
            transform();
#pragma unroll
            for (int j = 0; j < E; ++j) v[j] = C<R>(v[j].y, v[j].x);
        } else {
            transform();
        }
        flip ^= 1;  // consecutive transforms alternate between the two layouts of the exchange buffer (fft_core.cuh)
    }
    if (P.ctab_out) {
        const C<R> c = ldc_ro(reinterpret_cast<const C<R>*>(P.ctab_out) + line);
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] = v[j] * c;
    }

    // let the next kernel of the stream get its CTAs scheduled while this grid stores (it waits for our completion above)
    asm volatile("griddepcontrol.launch_dependents;");
    if constexpr (use_tma_field<N, COL>()) {
        if (dst && (BP.use_tmap & 1) && !P.readout && PAOS_TMA_FIELD_STORE) {
            // tile [row][W] in the exchange buffer, then W x 256-row boxes to the TMA unit
            // only the boxes that hold rows [out_lo, out_hi] leave: the next pass reads nothing else (device_types.h)
            const int box_lo = P.out_lo / TMA_BOX_ROWS, box_hi = P.out_hi < N ? P.out_hi / TMA_BOX_ROWS : -1;
            __syncthreads();  // every thread is done with the exchange data of the last transform
#pragma unroll
            for (int j = 0; j < E; ++j) {
                const int bx = (t + j * T) / TMA_BOX_ROWS;
                if (bx >= box_lo && bx <= box_hi) stc(smem + (size_t)(t + j * T) * W + w, v[j]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                constexpr int REALS = (int)(sizeof(C<R>) / sizeof(R));  // tensor-map elements per complex value
#pragma unroll 1
                for (int r0 = box_lo * TMA_BOX_ROWS; r0 <= box_hi * TMA_BOX_ROWS; r0 += TMA_BOX_ROWS)
                    tma_store_tile(&BP.tmap[b], smem + (size_t)r0 * W, tile * W * REALS, r0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile has left shared memory
            }
            return;
        }
    }
    if (dst) {  // null: a final read-out, the field itself is not needed any more
        const unsigned ospan = (unsigned)(P.out_hi - P.out_lo);  // the stretch the next pass reads (device_types.h)
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int idx = t + j * T;
            const size_t ga = COL ? ((size_t)idx * N + line) : ((size_t)line * N + idx);
            if ((unsigned)(idx - P.out_lo) <= ospan) stc_stream(dst + ga, v[j]);
        }
    }
    if constexpr (use_tma_field<N, COL>()) {
        if (P.readout && (BP.use_tmap & 2)) {
            // fused read-out through a TMA tile of W reals per row, like the field above
            R* rt = reinterpret_cast<R*>(smem_raw);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < E; ++j)
                rt[(size_t)(t + j * T) * W + w] = (P.readout == 3) ? v[j].x * v[j].x + v[j].y * v[j].y : readout_value<R>(v[j], P.readout);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
#pragma unroll 1
                for (int r0 = 0; r0 < N; r0 += TMA_BOX_ROWS) tma_store_tile(&BP.tmap_real[b], rt + (size_t)r0 * W, tile * W, r0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            return;
        }
    }
    if (P.readout) {
        // fused read-out (wfo.py:167-172, plot.py:125-130) so a snapshot costs no extra sweep
        R* out = reinterpret_cast<R*>(P.dst_real);
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int idx = t + j * T;
            const size_t ga = COL ? ((size_t)idx * N + line) : ((size_t)line * N + idx);
            // |.|^2 inline (the common sweep read-out); |.| and angle go through the out-of-line libm path
            out[ga] = (P.readout == 3) ? v[j].x * v[j].x + v[j].y * v[j].y : readout_value<R>(v[j], P.readout);
        }
    }
}

// host-side launcher table -------------------------------------------------------------------------
// One launch for the same-axis passes of nb <= BMAX wavefronts (nb = 1: the plain single-wavefront pass).
template <typename R, int N, int E, int W, bool COL, int MINB, int CAP>
cudaError_t launch_pass_cap(const PassParams* const* Ps, int nb, const void* tw1, const void* tw2, cudaStream_t st, int device) {
    using G = LineGeom<N, E>;
    constexpr int threads = W * G::T;
    const size_t smem = (size_t)W * G::line_stride(COL ? W : 1, (int)sizeof(C<R>)) * sizeof(C<R>) +
                        (use_tma_tables<N, COL, W>() ? (size_t)N * sizeof(C<R>) + 16 : (use_tma_field<N, COL>() ? 16 : 0));
    auto kern = pass_kernel<R, N, E, W, COL, MINB, CAP>;
    static bool configured[64] = {};  // per instantiation and device
    if (!configured[device & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[device & 63] = true;
    }
    // the parameter block is 1.8 KB (CAP = 1) or 28 KB (CAP = 16): kept on the heap behind a thread-local pointer, so that threads
    // that never launch (the planner threads of a batch) do not pay for ~1 MB of thread-local storage per instantiation set
    static thread_local std::unique_ptr<BatchParams<CAP>> holder;
    if (!holder) holder.reset(new BatchParams<CAP>());
    BatchParams<CAP>& BP = *holder;
    int total = 0, used = 0;
    bool all_maps = use_tma_field<N, COL>(), all_real = use_tma_field<N, COL>();
    for (int i = 0; i < nb; ++i) {
        // blank tiles have nothing to do unless they must store zeros (fused read-out, diagnostic zero fill): launch the rest
        PassParams& P = BP.p[used];
        P = *Ps[i];
        int tiles = N / W;
        P.tile_base = 0;
        if (!P.zero_fill && !P.readout) {
            P.tile_base = P.tile_lo > 0 ? P.tile_lo : 0;
            tiles = (P.tile_hi < N / W - 1 ? P.tile_hi : N / W - 1) - P.tile_base + 1;
            if (tiles <= 0) continue;  // the whole field is (virtually) zero after this pass
        }
        if (all_maps && P.tmap_host) std::memcpy(&BP.tmap[used], P.tmap_host, sizeof(CUtensorMap));
        else all_maps = false;
        if (all_real && P.readout && P.tmap_real_host) std::memcpy(&BP.tmap_real[used], P.tmap_real_host, sizeof(CUtensorMap));
        else all_real = false;
        BP.start[used] = total;
        total += tiles;
        ++used;
    }
    BP.use_tmap = (all_maps ? 1 : 0) | (all_real ? 2 : 0);
    if (used == 0) return cudaSuccess;
    BP.nb = used;
    BP.start[used] = total;
    static const bool pdl = getenv("PAOS_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)total);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, BP, reinterpret_cast<const C<R>*>(tw1), reinterpret_cast<const C<R>*>(tw2));
}

// One launch for the same-axis passes of nb <= BMAX wavefronts (nb = 1: the plain single-wavefront pass).
template <typename R, int N, int E, int W, bool COL, int MINB>
cudaError_t launch_pass_t(const PassParams* const* Ps, int nb, const void* tw1, const void* tw2, cudaStream_t st, int device) {
    if (nb < 1 || nb > BMAX) return cudaErrorInvalidValue;
    if (nb == 1) return launch_pass_cap<R, N, E, W, COL, MINB, 1>(Ps, nb, tw1, tw2, st, device);
    return launch_pass_cap<R, N, E, W, COL, MINB, BMAX>(Ps, nb, tw1, tw2, st, device);
}

}  // namespace paosb
