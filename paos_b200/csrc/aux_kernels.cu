// aux_kernels.cu -- the small kernels around the line passes: separable phase tables, the stop
// reduction, read-out (|.|, angle, |.|^2), the Zernike screen, PSD helpers.
#include <memory>

#include "aux_kernels.h"
#include "pass_kernel.cuh"

namespace paosb {

// ---- separable phase tables ------------------------------------------------------------------------
// exp(i * c1*c2 * u^2) with the phase formed in double-double and reduced mod 2*pi before sincos, so a
// table entry is accurate to ~1 ulp even when the phase is 1e7 rad.  u is rounded exactly as numpy does:
// x = (k - n//2) * dx   (wfo.py:359-360, :530-531)   fx = k' * (1.0/(n*d))   (numpy fftfreq, wfo.py:464-465)
__device__ __forceinline__ void phase_term(const TableTerm& tm, int k, int n, double& cs, double& sn) {
    const double kk = (double)(k - n / 2);
    double u;
    if (tm.kind == TERM_QFREQ) {
        const double val = 1.0 / ((double)n * tm.d);
        u = kk * val;
    } else {
        u = kk * tm.d;
    }
    const double u2 = u * u;  // rounded square, as numpy's xx**2
    // c = c1*c2 in double-double
    const double ch = tm.c1 * tm.c2;
    const double cl = fma(tm.c1, tm.c2, -ch);
    // p = c*u2 in double-double
    const double ph = ch * u2;
    const double pl = fma(ch, u2, -ph) + cl * u2;
    // reduce mod 2*pi (three-term constant)
    const double TWO_PI_H = 6.283185307179586232e+00;
    const double TWO_PI_M = 2.449293598294706414e-16;
    const double TWO_PI_L = -5.989539619436679332e-33;
    const double q = rint(ph * 0.15915494309189534561);
    double r = fma(-q, TWO_PI_H, ph);          // exact (cancellation)
    double rl = fma(-q, TWO_PI_M, pl);
    rl = fma(-q, TWO_PI_L, rl);
    const double rs = r + rl;
    const double re = rl - (rs - r);           // residual of the sum
    double s, c;
    sincos(rs, &s, &c);
    cs = fma(-s, re, c);
    sn = fma(c, re, s);
}

// number of the 32 sub-pixel centres of pixel k strictly inside (-full/2, full/2) around the centre;
// accumulation order of the published photutils routine (x = x0 - d/2; x += d)
__device__ __forceinline__ int subpixel_count(int k, double centre, double full) {
    const double half = full / 2.0;
    const double x0 = ((double)k - 0.5) - centre;
    const double x1 = x0 + 1.0;
    const double d = (x1 - x0) / 32.0;
    double x = x0 - 0.5 * d;
    int cnt = 0;
    for (int s = 0; s < 32; ++s) {
        x += d;
        if (fabs(x) < half) ++cnt;
    }
    return cnt;
}

template <typename R>
__global__ void build_tables_kernel(const __grid_constant__ TableBlock B) {
    const TableSpec& sp = B.spec[blockIdx.y];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = B.n;
    if (k >= n) return;
    if (sp.kind == TABLE_COUNT) {
        const int cnt = subpixel_count(k, sp.cnt_c, sp.cnt_full);
        reinterpret_cast<double*>(sp.out)[k] = (double)cnt;
        return;
    }
    double re = 1.0, im = 0.0;
    for (int i = 0; i < sp.nterms; ++i) {
        double c, s;
        if (sp.terms[i].kind == TERM_COUNT) {
            // separable rectangular aperture: (cx/32)*(cy/32) == (cy*cx)/1024 exactly (dyadic rationals)
            c = (double)subpixel_count(k, sp.terms[i].c1, sp.terms[i].c2) / 32.0;
            s = 0.0;
        } else {
            phase_term(sp.terms[i], k, n, c, s);
        }
        const double nr = re * c - im * s, ni = re * s + im * c;
        re = nr;
        im = ni;
    }
    double sc = sp.scale;
    if (sp.sign && (k & 1)) sc = -sc;
    re *= sc;
    im *= sc;
    stc(reinterpret_cast<C<R>*>(sp.out) + k, C<R>((R)re, (R)im));
}

cudaError_t launch_build_tables(const TableBlock& B, cudaStream_t st) {
    dim3 grid((B.n + 127) / 128, B.ntab);
    if (B.dtype == 0) build_tables_kernel<double><<<grid, 128, 0, st>>>(B);
    else build_tables_kernel<float><<<grid, 128, 0, st>>>(B);
    return cudaGetLastError();
}

// ---- edge tables of the elliptical masks (device_types.h: EdgeSpec) ---------------------------------------------
// One warp per line.  The half-warps look at 16 consecutive elements around the left and the right crossing of the line
// with the ellipse's rim, classify them exactly as the pass kernel will (ellipse_r2), and the lanes that fall into the edge
// band evaluate the exact overlap in parallel.  Lines whose band is wider than the window (within a pixel or so of the
// poles) are flagged and left to the pass kernel.
__global__ void __launch_bounds__(256) build_edge_tables_kernel(const __grid_constant__ EdgeBlock B) {
    const EdgeSpec& sp = B.spec[blockIdx.y];
    const GenOp& g = sp.g;
    const int n = B.n, lane = threadIdx.x & 31, line = sp.line_lo + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (line >= n || line > sp.line_hi) return;  // lines outside [line_lo, line_hi] hold no rim pixel and are never looked up
    const bool col = sp.col != 0;
    const int T = sp.T;
    const double c_al = col ? g.p1 : g.p0, c_cr = col ? g.p0 : g.p1;
    const float s_al = col ? (float)g.p3 : (float)g.p2, s_cr = col ? (float)g.p2 : (float)g.p3;
    const float in5 = (float)g.p5 - 1e-5f, out6 = (float)g.p6 + 1e-5f;
    const float da = __fmul_rn((float)T, s_al);
    const float b2 = ellipse_b2(line, c_cr, s_cr);
    int4* hdr = reinterpret_cast<int4*>(sp.out) + line;
    double* fac = reinterpret_cast<double*>(reinterpret_cast<char*>(sp.out) + (size_t)n * 16) + (size_t)line * 2 * EDGE_CAP;
    // geometry of the band along this line, in pixels (double; the windows keep a pixel of slack against the FP32 rounding)
    const double ro2 = (double)out6 - (double)b2, ri2 = (double)in5 - (double)b2;
    if (!(ro2 > 0.0)) {  // the whole line is outside the band: no edge pixel
        if (lane == 0) *hdr = make_int4(0, 0, 0, 0);
        return;
    }
    const double ro = sqrt(ro2) / (double)s_al, ri = ri2 > 0.0 ? sqrt(ri2) / (double)s_al : 0.0;
    bool flag = (ro - ri) + 3.0 > (double)EDGE_CAP || !(ri2 > 0.0) || 2.0 * ri < 34.0;  // wide band, or the two windows could meet
    const int side = lane >> 4, k = lane & 15;
    const int start = side == 0 ? (int)floor(c_al - ro) - 2 : (int)ceil(c_al + ro) + 2 - 15;
    const int idx = start + k;
    bool edge = false;
    if (!flag && idx >= 0 && idx < n) {
        const float r2 = ellipse_r2(ellipse_a0(idx % T, c_al, s_al), da, idx / T, b2);
        edge = !(r2 <= in5) && !(r2 >= out6);
    }
    const unsigned bits = __ballot_sync(0xffffffffu, edge);
    const unsigned half[2] = {bits & 0xffffu, bits >> 16};
    int first[2], len[2];
    for (int s2 = 0; s2 < 2; ++s2) {
        first[s2] = half[s2] ? __ffs((int)half[s2]) - 1 : 0;
        len[s2] = __popc(half[s2]);
        const unsigned run = half[s2] >> first[s2];
        // the run must be contiguous, fit the table and stay clear of both ends of the window (else it may continue outside)
        if (half[s2] && ((run & (run + 1)) != 0 || len[s2] > EDGE_CAP || (half[s2] & 0x8001u))) flag = true;
    }
    flag = __any_sync(0xffffffffu, flag);
    if (flag) {
        if (lane == 0) *hdr = make_int4(0, 0, 1 << 16, 0);
        return;
    }
    if (edge) {
        const double f = ellipse_fraction(g, (double)(col ? line : idx), (double)(col ? idx : line));
        fac[side * EDGE_CAP + (k - first[side])] = g.flag ? 1.0 - f : f;
    }
    if (lane == 0) {
        const int s0 = start + first[0] - 0;  // lane 0 is on side 0: its `start` is the left window's
        const int right_start = (int)ceil(c_al + ro) + 2 - 15;
        *hdr = make_int4(s0, right_start + first[1], len[0] | (len[1] << 8), 0);
    }
}

cudaError_t launch_build_edge_tables(const EdgeBlock& B, cudaStream_t st) {
    if (B.nspec < 1) return cudaSuccess;
    int lines = 1;
    for (int i = 0; i < B.nspec; ++i) lines = max(lines, B.spec[i].line_hi - B.spec[i].line_lo + 1);
    dim3 grid((lines + 7) / 8, B.nspec);
    build_edge_tables_kernel<<<grid, 256, 0, st>>>(B);
    return cudaGetLastError();
}

// ---- stop: sum |field * pending real factors|^2 (wfo.py:200) ----------------------------------------
struct Norm2Params {
    const void* src;
    int n;
    int ngen;
    GenOp gen[GMAX];
};
// the stop reductions of up to BMAX wavefronts in one launch: blockIdx.y = item
struct Norm2Batch {
    int nb;
    int pad;
    double* partials[BMAX];
    double* out[BMAX];
    Norm2Params p[BMAX];
};
static_assert(sizeof(Norm2Batch) <= 32764, "Norm2Batch must fit the kernel parameter space");

// (256, 4): without the bound the inlined exact-overlap routine takes 203 registers, one CTA per SM, and the 9 472 mostly
// empty CTAs of a batched reduction run in 64 waves (116 us); its rare path may spill
template <typename R>
__global__ void __launch_bounds__(256, 4) norm2_partial_kernel(const __grid_constant__ Norm2Batch B) {
    const Norm2Params& P = B.p[blockIdx.y];
    double* __restrict__ partials = B.partials[blockIdx.y];
    const int n = P.n;
    const C<R>* src = reinterpret_cast<const C<R>*>(P.src);
    double acc = 0.0;
    bool any_row = false;  // CTA-uniform
    // columns outside the bounding box of an elliptical aperture contribute exact zeros as well: a zoom-4 pupil covers a
    // quarter of the row (same FP32 test, in the along-row coordinate, as the one that blanks rows below)
    int x_lo = 0, x_hi = n - 1;
    for (int g = 0; g < P.ngen; ++g) {
        const GenOp& gg = P.gen[g];
        if (gg.kind == GEN_ELLIPSE && !gg.flag) {
            const double half = sqrt(gg.p6 + 2e-5) / gg.p2 + 1.0;  // a pixel further out certainly fails r^2 < p6 + 1e-5
            x_lo = max(x_lo, (int)floor(gg.p0 - half));
            x_hi = min(x_hi, (int)ceil(gg.p0 + half));
        }
    }
    // one row per CTA step (32-bit index math; the 64-bit div/mod of a flat index cost more than the mask itself)
    for (int iy = blockIdx.x; iy < n; iy += gridDim.x) {
        // a row that misses the bounding box of an elliptical aperture contributes nothing
        bool blank = false;
        for (int g = 0; g < P.ngen; ++g) {
            const GenOp& gg = P.gen[g];
            if (gg.kind == GEN_ELLIPSE && !gg.flag) {
                const float bq = (float)((double)iy - gg.p1) * (float)gg.p3;
                if (bq * bq >= (float)gg.p6 + 1e-5f) blank = true;
            }
        }
        if (blank) continue;
        any_row = true;
        for (int ix = x_lo + (int)threadIdx.x; ix <= x_hi; ix += blockDim.x) {
            C<R> v = src ? ldc(src + (size_t)iy * n + ix) : C<R>((R)1, (R)0);
            for (int g = 0; g < P.ngen; ++g) apply_gen(v, P.gen[g], ix, iy, n);
            acc += (double)v.x * (double)v.x + (double)v.y * (double)v.y;
        }
    }
    if (!any_row) {  // every row of this CTA is blanked by an aperture: nothing to reduce
        if (threadIdx.x == 0) partials[blockIdx.x] = 0.0;
        return;
    }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}

__global__ void norm2_final_kernel(const __grid_constant__ Norm2Batch B, int np) {
    const double* __restrict__ partials = B.partials[blockIdx.x];
    double* __restrict__ out = B.out[blockIdx.x];
    __shared__ double red[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < np; i += 256) acc += partials[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = 1.0 / sqrt(red[0]);
        out[1] = red[0];
    }
}

cudaError_t launch_norm2_batch(const Norm2Item* items, int nb, int n, int dtype, int npartials, cudaStream_t st) {
    if (nb < 1 || nb > BMAX) return cudaErrorInvalidValue;
    static thread_local std::unique_ptr<Norm2Batch> holder;  // 21 KB: heap, not thread-local storage
    if (!holder) holder.reset(new Norm2Batch());
    Norm2Batch& B = *holder;
    B.nb = nb;
    for (int b = 0; b < nb; ++b) {
        B.partials[b] = items[b].partials;
        B.out[b] = items[b].out_slot;
        B.p[b].src = items[b].src;
        B.p[b].n = n;
        B.p[b].ngen = items[b].ngen;
        for (int i = 0; i < items[b].ngen; ++i) B.p[b].gen[i] = items[b].gen[i];
    }
    const dim3 grid((unsigned)npartials, (unsigned)nb);
    if (dtype == 0) norm2_partial_kernel<double><<<grid, 256, 0, st>>>(B);
    else norm2_partial_kernel<float><<<grid, 256, 0, st>>>(B);
    norm2_final_kernel<<<nb, 256, 0, st>>>(B, npartials);
    return cudaGetLastError();
}

cudaError_t launch_norm2(const void* src, int n, int dtype, const GenOp* gen, int ngen, double* partials,
                         int npartials, double* out_slot, cudaStream_t st) {
    Norm2Item it{};
    it.src = src;
    it.ngen = ngen;
    for (int i = 0; i < ngen; ++i) it.gen[i] = gen[i];
    it.partials = partials;
    it.out_slot = out_slot;
    return launch_norm2_batch(&it, 1, n, dtype, npartials, st);
}

// ---- virtual zeros made real -----------------------------------------------------------------------------
// A blanked pass leaves the lines outside a band untouched (pass_kernel.cuh); before anything other than a pass kernel
// reads the field (a copy of the complex array, the stand-alone read-out, a stop reduction on a stored field) the
// zeros are written once.  axis 0: the band is a range of rows, axis 1: a range of columns.
template <typename R>
__global__ void __launch_bounds__(256) zero_outside_band_kernel(C<R>* __restrict__ f, int n, int axis, int lo, int hi) {
    const int y = blockIdx.y;
    if (axis == 0 && y >= lo && y <= hi) return;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
        if (axis == 1 && x >= lo && x <= hi) continue;
        stc(f + (size_t)y * n + x, C<R>((R)0, (R)0));
    }
}

cudaError_t launch_zero_outside_band(void* field, int n, int dtype, int axis, int lo, int hi, cudaStream_t st) {
    const dim3 grid((n + 255) / 256 < 4 ? (n + 255) / 256 : 4, n);
    if (dtype == 0) zero_outside_band_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<C<double>*>(field), n, axis, lo, hi);
    else zero_outside_band_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<C<float>*>(field), n, axis, lo, hi);
    return cudaGetLastError();
}

// ---- read-out (wfo.py:163-172, plot.py:125-130) ------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256) readout_kernel(const C<R>* __restrict__ src, size_t total, int what, R* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const C<R> v = ldc(src + i);
        out[i] = readout_value<R>(v, what);
    }
}

cudaError_t launch_readout(const void* src, int n, int dtype, int what, void* out, cudaStream_t st) {
    const size_t total = (size_t)n * n;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    if (dtype == 0) readout_kernel<double><<<blocks, 256, 0, st>>>(reinterpret_cast<const C<double>*>(src), total, what, reinterpret_cast<double*>(out));
    else readout_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const C<float>*>(src), total, what, reinterpret_cast<float*>(out));
    return cudaGetLastError();
}

// ---- Zernike wavefront-error screen (wfo.py:620-647, zernike.py:77-109, :245-247) --------------------
// Radial part through the Jacobi recurrence that scipy.special.eval_jacobi uses for integer order;
// the binomial prefactor and (-1)^k are folded into coef[] by the host.
__device__ __forceinline__ double jacobi_m0(int k, int m, double x) {
    // P_k^{(m,0)}(x) / binom(k+m, k)
    if (k == 0) return 1.0;
    const double alpha = (double)m;
    double d = (alpha + 2.0) * (x - 1.0) / (2.0 * (alpha + 1.0));
    double p = d + 1.0;
    for (int kk = 0; kk < k - 1; ++kk) {
        const double kf = kk + 1.0;
        const double t = 2.0 * kf + alpha;
        d = ((t * (t + 1.0) * (t + 2.0)) * (x - 1.0) * p + 2.0 * kf * kf * (t + 2.0) * d) /
            (2.0 * (kf + alpha + 1.0) * (kf + alpha + 1.0) * t);
        p = d + p;
    }
    return p;
}

// polar unit vector and rho of one pixel; returns false outside the unit disc (rho > 1 is masked: zernike.py:85)
__device__ __forceinline__ bool zern_polar(const ZernParams& Z, int ix, int iy, double& rho, double& c1, double& s1) {
    const int n = Z.n;
    const double x = (double)(ix - n / 2) * Z.dx, y = (double)(iy - n / 2) * Z.dy;
    const double r = sqrt(x * x + y * y);
    rho = r / Z.radius;
    if (!(rho <= 1.0)) return false;
    // unit vector of the polar angle (origin x: atan2(y, x); origin y: atan2(x, y)), rotated by the offset
    double c0, s0;
    if (r > 0.0) {
        c0 = (Z.origin == 0 ? x : y) / r;
        s0 = (Z.origin == 0 ? y : x) / r;
    } else {
        c0 = 1.0;
        s0 = 0.0;
    }
    c1 = c0 * Z.cos_off - s0 * Z.sin_off;
    s1 = s0 * Z.cos_off + c0 * Z.sin_off;
    return true;
}

// k-th polynomial times coef[k] (coef carries Z[k]*norm[k]*binom*(-1)^k, or norm*binom*(-1)^k for the covariance)
__device__ __forceinline__ double zern_term(const ZernParams& Z, int k, double rho, double c1, double s1, double xj) {
    const int m = Z.m[k], am = m < 0 ? -m : m, kr = (Z.nn[k] - am) / 2;
    double rp = 1.0, cm = 1.0, sm = 0.0;  // rho^|m| and cos/sin(|m| phi) by repeated multiplication
    for (int q = 0; q < am; ++q) {
        rp *= rho;
        const double nc = cm * c1 - sm * s1;
        sm = sm * c1 + cm * s1;
        cm = nc;
    }
    const double ang = (m > 0) ? cm : ((m < 0) ? sm : 1.0);
    return Z.coef[k] * (rp * jacobi_m0(kr, am, xj)) * ang;
}

__global__ void __launch_bounds__(256) zernike_kernel(const __grid_constant__ ZernParams Z, const unsigned char* __restrict__ mask,
                                                      double* __restrict__ out) {
    const int n = Z.n;
    const size_t total = (size_t)n * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int iy = (int)(i / n), ix = (int)(i % n);
        double rho, c1, s1, wfe = 0.0;
        if (zern_polar(Z, ix, iy, rho, c1, s1) && !(mask && mask[i])) {
            const double xj = 1.0 - 2.0 * (rho * rho);
            for (int k = 0; k < Z.K; ++k) wfe += zern_term(Z, k, rho, c1, s1, xj);
        }
        if (Z.accumulate) out[i] += wfe;
        else out[i] = wfe;
    }
}

cudaError_t launch_zernike(const ZernParams& Z, const unsigned char* mask, double* out, cudaStream_t st) {
    const size_t total = (size_t)Z.n * Z.n;
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    zernike_kernel<<<blocks, 256, 0, st>>>(Z, mask, out);
    return cudaGetLastError();
}

// ---- Zernike polynomials at arbitrary points (the stand-alone class, zernike.py:63-109) ---------------------------------
// stack[k][p] = norm_k * R_k(rho_p) * ang_k(phi_p); 0 where the point is masked (rho > 1 or mask[p]).  Z.coef carries
// norm * binom * (-1)^k as for the screen kernel; cos/sin(|m| phi) by repeated rotation of (cos phi, sin phi).
__global__ void __launch_bounds__(256) zernike_points_kernel(const __grid_constant__ ZernParams Z, const double* __restrict__ rho,
                                                             const double* __restrict__ phi, const unsigned char* __restrict__ mask,
                                                             size_t npoints, double* __restrict__ out) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npoints; p += (size_t)gridDim.x * blockDim.x) {
        const double r = rho[p];
        if (r > 1.0 || (mask && mask[p])) {
            for (int k = 0; k < Z.K; ++k) out[(size_t)k * npoints + p] = 0.0;
            continue;
        }
        double s1, c1;
        sincos(phi[p], &s1, &c1);
        const double xj = 1.0 - 2.0 * (r * r);
        for (int k = 0; k < Z.K; ++k) out[(size_t)k * npoints + p] = zern_term(Z, k, r, c1, s1, xj);
    }
}

// sums of Z_i * Z_j over all points (masked points hold zeros), one CTA per pair i <= j; CTA `npairs` counts the unmasked
// points.  Deterministic tree reduction.  out[pair], out[npairs] = count.
__global__ void __launch_bounds__(256) stack_cov_kernel(const double* __restrict__ stack, const double* __restrict__ rho,
                                                        const unsigned char* __restrict__ mask, int K, size_t npoints,
                                                        double* __restrict__ out) {
    const int npairs = K * (K + 1) / 2;
    int i = 0, rem = blockIdx.x;
    const bool counting = (int)blockIdx.x == npairs;
    if (!counting)
        while (rem >= K - i) {
            rem -= K - i;
            ++i;
        }
    const int j = i + rem;
    double acc = 0.0;
    for (size_t p = threadIdx.x; p < npoints; p += blockDim.x) {
        if (counting) acc += (rho[p] > 1.0 || (mask && mask[p])) ? 0.0 : 1.0;
        else acc += stack[(size_t)i * npoints + p] * stack[(size_t)j * npoints + p];
    }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

// U[i][p] = sum_j mat[i][j] * Z[j][p] (zernike.py:398-401, the Gram-Schmidt matrix applied to the stack)
__global__ void __launch_bounds__(256) stack_transform_kernel(const double* __restrict__ stack, const double* __restrict__ mat, int K,
                                                              size_t npoints, double* __restrict__ out) {
    extern __shared__ double msh[];  // K*K
    for (int q = threadIdx.x; q < K * K; q += blockDim.x) msh[q] = mat[q];
    __syncthreads();
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npoints; p += (size_t)gridDim.x * blockDim.x) {
        double z[ZERN_MAX];
        for (int j = 0; j < K; ++j) z[j] = stack[(size_t)j * npoints + p];
        for (int i = 0; i < K; ++i) {
            double acc = 0.0;
            for (int j = 0; j < K; ++j) acc += msh[i * K + j] * z[j];
            out[(size_t)i * npoints + p] = acc;
        }
    }
}

cudaError_t launch_zernike_points(const ZernParams& Z, const double* rho, const double* phi, const unsigned char* mask,
                                  size_t npoints, double* out, cudaStream_t st) {
    const int blocks = (int)((npoints + 255) / 256 < 148 * 8 ? (npoints + 255) / 256 : 148 * 8);
    zernike_points_kernel<<<blocks, 256, 0, st>>>(Z, rho, phi, mask, npoints, out);
    return cudaGetLastError();
}
cudaError_t launch_stack_cov(const double* stack, const double* rho, const unsigned char* mask, int K, size_t npoints, double* out,
                             cudaStream_t st) {
    stack_cov_kernel<<<K * (K + 1) / 2 + 1, 256, 0, st>>>(stack, rho, mask, K, npoints, out);
    return cudaGetLastError();
}
cudaError_t launch_stack_transform(const double* stack, const double* mat, int K, size_t npoints, double* out, cudaStream_t st) {
    const int blocks = (int)((npoints + 255) / 256 < 148 * 8 ? (npoints + 255) / 256 : 148 * 8);
    stack_transform_kernel<<<blocks, 256, (size_t)K * K * sizeof(double), st>>>(stack, mat, K, npoints, out);
    return cudaGetLastError();
}

// ---- Zernike covariance (zernike.py:293-317): sums of Z_i*Z_j over the unmasked pixels ---------------------
// One CTA walks over tiles of COV_TILE pixels: every polynomial of every pixel of the tile goes to shared memory,
// then each thread accumulates the dot products of the (i <= j) pairs it owns.  Per-CTA partial sums (and the
// unmasked pixel count in slot K*K) go to `partial[block][K*K+1]`; the host adds the blocks up.
constexpr int COV_TILE = 128;
__global__ void __launch_bounds__(256) zernike_cov_kernel(const __grid_constant__ ZernParams Z, const unsigned char* __restrict__ mask,
                                                          double* __restrict__ partial) {
    extern __shared__ double zs[];  // [K][COV_TILE + 1]
    constexpr int LD = COV_TILE + 1;
    const int K = Z.K, n = Z.n, npairs = K * (K + 1) / 2;
    const size_t total = (size_t)n * n;
    constexpr int PMAX = 9;  // pairs per thread: K <= 64 -> 2080 pairs / 256 threads
    double acc[PMAX];
    int pi_[PMAX], pj_[PMAX];
#pragma unroll
    for (int q = 0; q < PMAX; ++q) {
        acc[q] = 0.0;
        const int pidx = threadIdx.x + q * 256;
        int i = 0, rem = pidx;  // unrank (i <= j) from the pair index
        if (pidx < npairs) {
            while (rem >= K - i) {
                rem -= K - i;
                ++i;
            }
        }
        pi_[q] = i;
        pj_[q] = i + rem;
    }
    double count = 0.0;
    for (size_t base = (size_t)blockIdx.x * COV_TILE; base < total; base += (size_t)gridDim.x * COV_TILE) {
        __syncthreads();
        if (threadIdx.x < COV_TILE) {
            const size_t i = base + threadIdx.x;
            double rho = 2.0, c1 = 1.0, s1 = 0.0;
            bool ok = i < total;
            if (ok) ok = zern_polar(Z, (int)(i % n), (int)(i / n), rho, c1, s1) && !(mask && mask[i]);
            const double xj = 1.0 - 2.0 * (rho * rho);
            for (int k = 0; k < K; ++k) zs[k * LD + threadIdx.x] = ok ? zern_term(Z, k, rho, c1, s1, xj) : 0.0;
            if (ok) count += 1.0;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PMAX; ++q) {
            if (threadIdx.x + q * 256 < npairs) {
                const double* a = zs + pi_[q] * LD;
                const double* b = zs + pj_[q] * LD;
                double s = 0.0;
                for (int p2 = 0; p2 < COV_TILE; ++p2) s += a[p2] * b[p2];
                acc[q] += s;
            }
        }
    }
    double* mine = partial + (size_t)blockIdx.x * (K * K + 1);
#pragma unroll
    for (int q = 0; q < PMAX; ++q)
        if (threadIdx.x + q * 256 < npairs) {
            mine[pi_[q] * K + pj_[q]] = acc[q];
            mine[pj_[q] * K + pi_[q]] = acc[q];
        }
    // unmasked pixel count of this CTA
    __shared__ double red[256];
    red[threadIdx.x] = count;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) mine[K * K] = red[0];
}

cudaError_t launch_zernike_cov(const ZernParams& Z, const unsigned char* mask, double* partial, int blocks, cudaStream_t st) {
    const size_t smem = (size_t)Z.K * (COV_TILE + 1) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(zernike_cov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    zernike_cov_kernel<<<blocks, 256, smem, st>>>(Z, mask, partial);
    return cudaGetLastError();
}

// ---- PSD helpers (psd.py:113-148) --------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256) real_to_complex_kernel(const double* __restrict__ src, size_t total, C<R>* __restrict__ dst) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        stc(dst + i, C<R>((R)src[i], (R)0));
}
template <typename R>
__global__ void __launch_bounds__(256) psd_finalize_kernel(const C<R>* __restrict__ f, const double* __restrict__ noise2, double SR,
                                                           double unit, double inv_nn, size_t total, double* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        double w = (double)ldc(f + i).x * inv_nn;     // ifft2 default normalisation 1/(Nx*Ny)
        if (noise2) w += SR * noise2[i];
        w *= 2.0;
        w *= unit;
        out[i] = w;
    }
}

cudaError_t launch_real_to_complex(const double* src, int n, int dtype, void* dst, cudaStream_t st) {
    const size_t total = (size_t)n * n;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    if (dtype == 0) real_to_complex_kernel<double><<<blocks, 256, 0, st>>>(src, total, reinterpret_cast<C<double>*>(dst));
    else real_to_complex_kernel<float><<<blocks, 256, 0, st>>>(src, total, reinterpret_cast<C<float>*>(dst));
    return cudaGetLastError();
}

cudaError_t launch_psd_finalize(const void* f, int n, int dtype, const double* noise2, double SR, double unit, double* out,
                                cudaStream_t st) {
    const size_t total = (size_t)n * n;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    const double inv_nn = 1.0 / ((double)n * (double)n);
    if (dtype == 0) psd_finalize_kernel<double><<<blocks, 256, 0, st>>>(reinterpret_cast<const C<double>*>(f), noise2, SR, unit, inv_nn, total, out);
    else psd_finalize_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const C<float>*>(f), noise2, SR, unit, inv_nn, total, out);
    return cudaGetLastError();
}

// standard normal field: Philox-4x32-10 counter + Box-Muller (fast mode of paos_wfo_psd; statistical parity only)
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* o) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
__global__ void __launch_bounds__(256) normal_kernel(uint64_t seed, uint32_t stream_id, size_t total, double* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (total + 1) / 2; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t o[4];
        philox4x32((uint32_t)i, (uint32_t)(i >> 32), stream_id, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
        const double u1 = ((double)(((uint64_t)o[0] << 21) ^ (o[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
        const double u2 = ((double)(((uint64_t)o[2] << 21) ^ (o[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
        const double rad = sqrt(-2.0 * log(u1));
        double s, c;
        sincospi(2.0 * u2, &s, &c);
        out[2 * i] = rad * c;
        if (2 * i + 1 < total) out[2 * i + 1] = rad * s;
    }
}
cudaError_t launch_normal(uint64_t seed, uint32_t stream_id, int n, double* out, cudaStream_t st) {
    const size_t total = (size_t)n * n;
    const int blocks = (int)((total / 2 + 255) / 256 < 148 * 8 ? (total / 2 + 255) / 256 : 148 * 8);
    normal_kernel<<<blocks, 256, 0, st>>>(seed, stream_id, total, out);
    return cudaGetLastError();
}

// ---- encircled energy (docs/source/user/aberration/index.rst:47-67; no code in the reference) ---------------------------
// f(R) = sum of the PSF over pixels with r <= R, r = sqrt(x^2 + y^2) / (F# * lambda) in image-space-normalised units,
// x = (ix - xc) * dx.  Radial histogram in shared memory (bins of width dR, one overflow bin), merged with one atomic
// per bin and CTA; ee_finish turns the histogram into the cumulative, normalised curve.
template <typename R>
__global__ void __launch_bounds__(256) ee_hist_kernel(const R* __restrict__ psf, int n, double dx, double dy, double xc, double yc,
                                                      double inv_bin, int nbins, double* __restrict__ hist) {
    extern __shared__ double sh[];
    for (int b = threadIdx.x; b <= nbins; b += blockDim.x) sh[b] = 0.0;
    __syncthreads();
    for (int iy = blockIdx.x; iy < n; iy += gridDim.x) {
        const double y = ((double)iy - yc) * dy;
        for (int ix = threadIdx.x; ix < n; ix += blockDim.x) {
            const double x = ((double)ix - xc) * dx;
            const double q = sqrt(x * x + y * y) * inv_bin;
            const int b = q < (double)nbins ? (int)q : nbins;
            const double v = (double)psf[(size_t)iy * n + ix];
            // lanes of a warp that hit the same bin (coarse bins: all 32 of them) add up in registers first; the
            // shared-memory double atomic is a compare-and-swap loop and serialises on equal addresses
            const unsigned peers = __match_any_sync(0xffffffffu, b);
            double tot = 0.0;
            for (unsigned rem = peers; rem; rem &= rem - 1) tot += __shfl_sync(peers, v, __ffs(rem) - 1);
            if ((int)(threadIdx.x & 31) == __ffs(peers) - 1 && tot != 0.0) atomicAdd(&sh[b], tot);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b <= nbins; b += blockDim.x)
        if (sh[b] != 0.0) atomicAdd(&hist[b], sh[b]);
}

// Centred grid (xc = yc = n/2, the WFO's own origin): the pixels (ix, iy), (n-ix, iy), (ix, n-iy), (n-ix, n-iy) share one
// radius, so one thread adds the four and issues one histogram update: 4x fewer square roots and shared-memory atomics
// (which are compare-and-swap loops for fp64 and bound this kernel).  Column 0 and row 0 have no mirror image inside the
// grid and are taken singly; x = 0 and y = 0 are their own images.
template <typename R>
__global__ void __launch_bounds__(256) ee_hist_folded_kernel(const R* __restrict__ psf, int n, double dx, double dy, double inv_bin,
                                                             int nbins, double* __restrict__ hist) {
    extern __shared__ double sh[];
    for (int b = threadIdx.x; b <= nbins; b += blockDim.x) sh[b] = 0.0;
    __syncthreads();
    const int h = n / 2;
    // quadrant x >= 0, y >= 0 plus (as "row h" / "column h" of the loop) the unpaired row 0 and column 0
    for (int qy = blockIdx.x; qy <= h; qy += gridDim.x) {
        const bool row0 = qy == h;            // extra iteration: the grid's row 0 (y = -h*dy)
        const int iy = row0 ? 0 : h + qy;
        const double y = row0 ? -(double)h * dy : (double)qy * dy;
        for (int qx = threadIdx.x; qx < 256 * ((h + 1 + 255) / 256); qx += blockDim.x) {
            double v = 0.0, q = 0.0;
            if (qx <= h) {
                const bool col0 = qx == h;    // extra iteration: the grid's column 0
                const int ix = col0 ? 0 : h + qx;
                const double x = col0 ? -(double)h * dx : (double)qx * dx;
                q = sqrt(x * x + y * y) * inv_bin;
                v = (double)psf[(size_t)iy * n + ix];
                const bool mx = !col0 && qx > 0, my = !row0 && qy > 0;  // mirror images exist and are distinct
                if (mx) v += (double)psf[(size_t)iy * n + (n - ix)];
                if (my) v += (double)psf[(size_t)(n - iy) * n + ix];
                if (mx && my) v += (double)psf[(size_t)(n - iy) * n + (n - ix)];
            }
            const int b = q < (double)nbins ? (int)q : nbins;
            const unsigned peers = __match_any_sync(0xffffffffu, b);
            double tot = 0.0;
            for (unsigned rem = peers; rem; rem &= rem - 1) tot += __shfl_sync(peers, v, __ffs(rem) - 1);
            if ((int)(threadIdx.x & 31) == __ffs(peers) - 1 && tot != 0.0) atomicAdd(&sh[b], tot);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b <= nbins; b += blockDim.x)
        if (sh[b] != 0.0) atomicAdd(&hist[b], sh[b]);
}

// one CTA: ee[k] = (hist[0] + ... + hist[k]) / total, total includes the overflow bin; ee[nbins] receives the total.
// Block-wide scan (5 consecutive bins per thread, 1024 threads >= 4097 bins): a serial loop over global memory cost
// more than the whole propagation.
__global__ void __launch_bounds__(1024) ee_finish_kernel(double* __restrict__ hist, int nbins, double* __restrict__ ee) {
    constexpr int PER = 5;
    __shared__ double warp_sum[32];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    double v[PER], mine = 0.0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int b = t * PER + k;
        v[k] = b <= nbins ? hist[b] : 0.0;
        if (b <= nbins) hist[b] = 0.0;  // leave the histogram clean for the next call on this handle (stream order)
        mine += v[k];
    }
    double incl = mine;  // inclusive scan over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        double w = warp_sum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += up;
        }
        warp_sum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const double total = warp_sum[31];
    double acc = (wid ? warp_sum[wid - 1] : 0.0) + (incl - mine);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int b = t * PER + k;
        acc += v[k];
        if (b < nbins) ee[b] = total != 0.0 ? acc / total : 0.0;
    }
    if (t == 0) ee[nbins] = total;
}

cudaError_t launch_encircled_energy(const void* psf, int n, int real_is_float, double dx, double dy, double xc, double yc,
                                    double inv_bin, int nbins, double* hist, double* ee, cudaStream_t st) {
    // `hist` is all zeros on entry: zeroed at allocation and by ee_finish_kernel after every use
    const int blocks = n < 148 * 4 ? n : 148 * 4;
    const size_t smem = (size_t)(nbins + 1) * sizeof(double);
    if (xc == 0.5 * n && yc == 0.5 * n) {
        const int fb = n / 2 + 1 < 148 * 4 ? n / 2 + 1 : 148 * 4;
        if (real_is_float) ee_hist_folded_kernel<float><<<fb, 256, smem, st>>>(reinterpret_cast<const float*>(psf), n, dx, dy, inv_bin, nbins, hist);
        else ee_hist_folded_kernel<double><<<fb, 256, smem, st>>>(reinterpret_cast<const double*>(psf), n, dx, dy, inv_bin, nbins, hist);
    } else if (real_is_float) {
        ee_hist_kernel<float><<<blocks, 256, smem, st>>>(reinterpret_cast<const float*>(psf), n, dx, dy, xc, yc, inv_bin, nbins, hist);
    } else {
        ee_hist_kernel<double><<<blocks, 256, smem, st>>>(reinterpret_cast<const double*>(psf), n, dx, dy, xc, yc, inv_bin, nbins, hist);
    }
    ee_finish_kernel<<<1, 1024, 0, st>>>(hist, nbins, ee);
    return cudaGetLastError();
}

// ---- Strehl ratio helpers (docs/source/user/aberration/index.rst:27-45) -----------------------------------------
// (a) peak and centre value of a PSF (exact Strehl = centre irradiance of the aberrated PSF / that of the ideal one);
// (b) mean and variance of a wavefront-error screen over the pupil rho <= 1 (Marechal: Strehl ~ 1 - k^2 sigma_W^2).
// Single CTA of 1024 threads: both are O(N^2) reads of data that sits in L2 right after the pass that produced it, and a
// fixed summation order keeps the result reproducible run to run.
template <typename R>
__global__ void __launch_bounds__(1024) psf_peak_kernel(const R* __restrict__ psf, int n, double* __restrict__ out) {
    __shared__ double red[1024];
    const size_t total = (size_t)n * n;
    double mx = 0.0;
    for (size_t i = threadIdx.x; i < total; i += 1024) mx = fmax(mx, (double)psf[i]);
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = (double)psf[(size_t)(n / 2) * n + n / 2];  // x = y = 0 sits at pixel (n/2, n/2)
        out[1] = red[0];
    }
}

__global__ void __launch_bounds__(1024) screen_stats_kernel(const double* __restrict__ screen, int n, double inv_r2, double dx,
                                                            double dy, double* __restrict__ out) {
    __shared__ double r0[1024], r1[1024], r2[1024];
    double s = 0.0, ss = 0.0, cnt = 0.0;
    for (int iy = threadIdx.x >> 5; iy < n; iy += 32) {
        const double y = (double)(iy - n / 2) * dy;
        for (int ix = threadIdx.x & 31; ix < n; ix += 32) {
            const double x = (double)(ix - n / 2) * dx;
            if ((x * x + y * y) * inv_r2 <= 1.0) {
                const double w = screen[(size_t)iy * n + ix];
                s += w;
                ss += w * w;
                cnt += 1.0;
            }
        }
    }
    r0[threadIdx.x] = s;
    r1[threadIdx.x] = ss;
    r2[threadIdx.x] = cnt;
    __syncthreads();
    for (int k = 512; k > 0; k >>= 1) {
        if (threadIdx.x < k) {
            r0[threadIdx.x] += r0[threadIdx.x + k];
            r1[threadIdx.x] += r1[threadIdx.x + k];
            r2[threadIdx.x] += r2[threadIdx.x + k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double c = r2[0], mean = c > 0 ? r0[0] / c : 0.0;
        out[0] = mean;
        out[1] = c > 0 ? fmax(r1[0] / c - mean * mean, 0.0) : 0.0;  // variance sigma_W^2 over the pupil
        out[2] = c;
    }
}

cudaError_t launch_psf_peak(const void* psf, int n, int real_is_float, double* out, cudaStream_t st) {
    if (real_is_float) psf_peak_kernel<float><<<1, 1024, 0, st>>>(reinterpret_cast<const float*>(psf), n, out);
    else psf_peak_kernel<double><<<1, 1024, 0, st>>>(reinterpret_cast<const double*>(psf), n, out);
    return cudaGetLastError();
}

cudaError_t launch_screen_stats(const double* screen, int n, double radius, double dx, double dy, double* out, cudaStream_t st) {
    screen_stats_kernel<<<1, 1024, 0, st>>>(screen, n, 1.0 / (radius * radius), dx, dy, out);
    return cudaGetLastError();
}

// ---- reduced host product: centred (or any) window of a real read-out, optionally narrowed to float ------------------
template <typename S, typename D>
__global__ void __launch_bounds__(256) crop_convert_kernel(const S* __restrict__ src, int n, int x0, int y0, int nx, int ny, D* __restrict__ dst) {
    const size_t total = (size_t)nx * ny;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int iy = (int)(i / nx), ix = (int)(i - (size_t)iy * nx);
        dst[i] = (D)src[(size_t)(y0 + iy) * n + (x0 + ix)];
    }
}

cudaError_t launch_crop_convert(const void* src, int n, int src_is_float, int x0, int y0, int nx, int ny, int dst_is_float, void* dst,
                                cudaStream_t st) {
    const size_t total = (size_t)nx * ny;
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    if (src_is_float && dst_is_float) crop_convert_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)src, n, x0, y0, nx, ny, (float*)dst);
    else if (src_is_float) crop_convert_kernel<float, double><<<blocks, 256, 0, st>>>((const float*)src, n, x0, y0, nx, ny, (double*)dst);
    else if (dst_is_float) crop_convert_kernel<double, float><<<blocks, 256, 0, st>>>((const double*)src, n, x0, y0, nx, ny, (float*)dst);
    else crop_convert_kernel<double, double><<<blocks, 256, 0, st>>>((const double*)src, n, x0, y0, nx, ny, (double*)dst);
    return cudaGetLastError();
}

}  // namespace paosb
