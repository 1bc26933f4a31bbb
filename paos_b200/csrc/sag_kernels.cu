// sag_kernels.cu -- device preparation of a Grid Sag map (reference paos/classes/wfo.py:696-862): sub-pixel recentring by a
// Fourier shift, pad / crop to the extent of the WFO grid, cubic-spline rescale / resize with Gaussian anti-aliasing.
//
// The reference does this on the host with scipy.ndimage.fourier_shift and scikit-image 0.24's rescale / resize
// (order 3, explicit anti_aliasing, mode 'reflect' = scipy 'mirror', clip to the input range).  Every step is separable, so
// the device version is a handful of per-axis kernels on real (rows x cols) arrays of arbitrary shape:
//   * circular convolution along an axis with a dense kernel = the Fourier shift of a periodic signal (the map is not a
//     power of two, e.g. 838 x 1158, so this is the O(n^2)-per-line form; h = ifft(shift multiplier) is built on the host);
//   * symmetric FIR along an axis with whole-sample-symmetric (mirror) boundaries = the Gaussian pre-filter;
//   * the recursive cubic B-spline pre-filter along an axis with exact mirror initialisation (one thread per line);
//   * the 4-tap B-spline evaluation at the pixel-centre-aligned positions (o + 1/2) * n_in/n_out - 1/2;
//   * pad / crop, min / max, clip, and the final mask > 0.1 test.
// The host logic (which steps run, with which sizes) follows wfo.py:753-862 decision for decision and lives in
// runtime.cu: paos_wfo_grid_sag.  Parity: held to the oracle's flow (oracle/paos_np.py: grid_sag over the scipy restatement
// of the skimage calls, itself unpinned against scikit-image, which this image lacks) in tests/test_gpu_sag.py.
#include <cuda_runtime.h>

#include <cmath>

#include "sag_kernels.h"

namespace paosb {

namespace {

__device__ __forceinline__ int mirror_index(long i, int n) {
    if (n == 1) return 0;
    const long period = 2L * (n - 1);
    long m = i % period;
    if (m < 0) m += period;
    return (int)(m >= n ? period - m : m);
}

// element (r, c) of a rows x cols array; `axis` 0 runs along rows (index r), 1 along columns
__global__ void __launch_bounds__(256) conv_circ_kernel(const double* __restrict__ a, int rows, int cols, int axis,
                                                        const double* __restrict__ h, double sign, int accumulate,
                                                        double* __restrict__ out) {
    const size_t total = (size_t)rows * cols;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / cols), c = (int)(e - (size_t)r * cols);
        const int n = axis == 0 ? rows : cols, i = axis == 0 ? r : c;
        const size_t stride = axis == 0 ? (size_t)cols : 1, base = axis == 0 ? (size_t)c : (size_t)r * cols;
        double acc = 0.0;
        int k = i;  // h index (i - m) mod n for m = 0, 1, ...
        for (int m = 0; m < n; ++m) {
            acc += h[k] * a[base + (size_t)m * stride];
            k = k == 0 ? n - 1 : k - 1;
        }
        out[e] = accumulate ? out[e] + sign * acc : sign * acc;
    }
}

__global__ void __launch_bounds__(256) fir_mirror_kernel(const double* __restrict__ a, int rows, int cols, int axis,
                                                         const double* __restrict__ w, int radius, double* __restrict__ out) {
    const size_t total = (size_t)rows * cols;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / cols), c = (int)(e - (size_t)r * cols);
        const int n = axis == 0 ? rows : cols, i = axis == 0 ? r : c;
        const size_t stride = axis == 0 ? (size_t)cols : 1, base = axis == 0 ? (size_t)c : (size_t)r * cols;
        double acc = 0.0;
        for (int k = -radius; k <= radius; ++k) acc += w[k + radius] * a[base + (size_t)mirror_index((long)i + k, n) * stride];
        out[e] = acc;
    }
}

// in place; one thread per line
__global__ void __launch_bounds__(128) bspline_prefilter_kernel(double* __restrict__ a, int rows, int cols, int axis) {
    const int lines = axis == 0 ? cols : rows, n = axis == 0 ? rows : cols;
    const int line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= lines || n == 1) return;
    const size_t stride = axis == 0 ? (size_t)cols : 1, base = axis == 0 ? (size_t)line : (size_t)line * cols;
    double* c = a + base;
    const double z = -0.26794919243112270647;  // sqrt(3) - 2, pole of the cubic B-spline
    const double zn1 = pow(z, (double)(n - 1));
    // causal initialisation: sum over one period of the mirror extension, closed over all periods
    double head = 6.0 * c[0] + zn1 * (6.0 * c[(size_t)(n - 1) * stride]);
    double zk = z;
    for (int k = 1; k < n - 1; ++k) {
        head += zk * (6.0 * c[(size_t)k * stride] + zn1 * (6.0 * c[(size_t)(n - 1 - k) * stride]));
        zk *= z;
        if (fabs(zk) < 1e-300) break;
    }
    double prev = head / (1.0 - zn1 * zn1);
    double last2 = 0.0;  // c[n-2] after the causal pass
    c[0] = prev;
    for (int k = 1; k < n; ++k) {
        const double v = 6.0 * c[(size_t)k * stride] + z * prev;
        c[(size_t)k * stride] = v;
        if (k == n - 2) last2 = v;
        prev = v;
    }
    if (n == 2) last2 = c[0];
    double nxt = (z / (z * z - 1.0)) * (prev + z * last2);
    c[(size_t)(n - 1) * stride] = nxt;
    for (int k = n - 2; k >= 0; --k) {
        const double v = z * (nxt - c[(size_t)k * stride]);
        c[(size_t)k * stride] = v;
        nxt = v;
    }
}

__global__ void __launch_bounds__(256) bspline_interp_kernel(const double* __restrict__ cf, int rows, int cols, int axis, int n_out,
                                                             double* __restrict__ out) {
    const int orows = axis == 0 ? n_out : rows, ocols = axis == 1 ? n_out : cols;
    const size_t total = (size_t)orows * ocols;
    const int n_in = axis == 0 ? rows : cols;
    const double ratio = (double)n_in / (double)n_out;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / ocols), c = (int)(e - (size_t)r * ocols);
        const int o = axis == 0 ? r : c;
        const size_t stride = axis == 0 ? (size_t)cols : 1, base = axis == 0 ? (size_t)c : (size_t)r * cols;
        const double x = ((double)o + 0.5) * ratio - 0.5;
        const double fl = floor(x);
        const long i0 = (long)fl;
        const double t = x - fl;
        const double omt = 1.0 - t;
        const double w0 = omt * omt * omt / 6.0;
        const double w1 = (3.0 * t * t * t - 6.0 * t * t + 4.0) / 6.0;
        const double w2 = (-3.0 * t * t * t + 3.0 * t * t + 3.0 * t + 1.0) / 6.0;
        const double w3 = t * t * t / 6.0;
        double acc = 0.0;
        acc += w0 * cf[base + (size_t)mirror_index(i0 - 1, n_in) * stride];
        acc += w1 * cf[base + (size_t)mirror_index(i0, n_in) * stride];
        acc += w2 * cf[base + (size_t)mirror_index(i0 + 1, n_in) * stride];
        acc += w3 * cf[base + (size_t)mirror_index(i0 + 2, n_in) * stride];
        out[e] = acc;
    }
}

__global__ void __launch_bounds__(1024) minmax_kernel(const double* __restrict__ a, size_t total, double* __restrict__ lohi) {
    __shared__ double lo[1024], hi[1024];
    double l = INFINITY, h = -INFINITY;
    for (size_t i = threadIdx.x; i < total; i += 1024) {
        const double v = a[i];
        if (v == v) {  // nanmin / nanmax when the map holds NaNs
            l = fmin(l, v);
            h = fmax(h, v);
        }
    }
    lo[threadIdx.x] = l;
    hi[threadIdx.x] = h;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            lo[threadIdx.x] = fmin(lo[threadIdx.x], lo[threadIdx.x + s]);
            hi[threadIdx.x] = fmax(hi[threadIdx.x], hi[threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        lohi[0] = lo[0];
        lohi[1] = hi[0];
    }
}

__global__ void __launch_bounds__(256) clip_kernel(double* __restrict__ a, size_t total, const double* __restrict__ lohi) {
    const double lo = lohi[0], hi = lohi[1];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        a[i] = fmin(fmax(a[i], lo), hi);
}

__global__ void __launch_bounds__(256) padcrop_kernel(const double* __restrict__ a, int rows, int cols, int row_off, int col_off,
                                                      double fill, int orows, int ocols, double* __restrict__ out) {
    const size_t total = (size_t)orows * ocols;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / ocols) + row_off, c = (int)(e % ocols) + col_off;
        out[e] = (r >= 0 && r < rows && c >= 0 && c < cols) ? a[(size_t)r * cols + c] : fill;
    }
}

// raw map -> (sag with masked samples zeroed, mask as 0/1 doubles); given: optional explicit mask bytes (non-zero = masked)
__global__ void __launch_bounds__(256) sag_split_kernel(const double* __restrict__ raw, const unsigned char* __restrict__ given, size_t total,
                                                        double* __restrict__ sag, double* __restrict__ mask) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const double v = raw[i];
        const bool m = given ? given[i] != 0 : (!isfinite(v) || v == 0.0);
        sag[i] = m ? 0.0 : v;
        mask[i] = m ? 1.0 : 0.0;
    }
}

__global__ void __launch_bounds__(256) sag_finish_kernel(const double* __restrict__ sag, const double* __restrict__ mask, size_t total,
                                                         double* __restrict__ screen, unsigned char* __restrict__ mask_out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const bool m = mask[i] > 0.1;
        screen[i] = m ? 0.0 : sag[i];
        if (mask_out) mask_out[i] = m ? 1 : 0;
    }
}

inline int blocks_for(size_t total) { return (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8); }

}  // namespace

cudaError_t sag_conv_circ(const double* a, int rows, int cols, int axis, const double* h_dev, double sign, int accumulate, double* out,
                          cudaStream_t st) {
    conv_circ_kernel<<<blocks_for((size_t)rows * cols), 256, 0, st>>>(a, rows, cols, axis, h_dev, sign, accumulate, out);
    return cudaGetLastError();
}
cudaError_t sag_fir_mirror(const double* a, int rows, int cols, int axis, const double* w_dev, int radius, double* out, cudaStream_t st) {
    fir_mirror_kernel<<<blocks_for((size_t)rows * cols), 256, 0, st>>>(a, rows, cols, axis, w_dev, radius, out);
    return cudaGetLastError();
}
cudaError_t sag_bspline_prefilter(double* a, int rows, int cols, int axis, cudaStream_t st) {
    const int lines = axis == 0 ? cols : rows;
    bspline_prefilter_kernel<<<(lines + 127) / 128, 128, 0, st>>>(a, rows, cols, axis);
    return cudaGetLastError();
}
cudaError_t sag_bspline_interp(const double* cf, int rows, int cols, int axis, int n_out, double* out, cudaStream_t st) {
    const size_t total = (size_t)(axis == 0 ? n_out : rows) * (axis == 1 ? n_out : cols);
    bspline_interp_kernel<<<blocks_for(total), 256, 0, st>>>(cf, rows, cols, axis, n_out, out);
    return cudaGetLastError();
}
cudaError_t sag_minmax(const double* a, size_t total, double* lohi_dev, cudaStream_t st) {
    minmax_kernel<<<1, 1024, 0, st>>>(a, total, lohi_dev);
    return cudaGetLastError();
}
cudaError_t sag_clip(double* a, size_t total, const double* lohi_dev, cudaStream_t st) {
    clip_kernel<<<blocks_for(total), 256, 0, st>>>(a, total, lohi_dev);
    return cudaGetLastError();
}
cudaError_t sag_padcrop(const double* a, int rows, int cols, int row_off, int col_off, double fill, int orows, int ocols, double* out,
                        cudaStream_t st) {
    padcrop_kernel<<<blocks_for((size_t)orows * ocols), 256, 0, st>>>(a, rows, cols, row_off, col_off, fill, orows, ocols, out);
    return cudaGetLastError();
}
cudaError_t sag_split(const double* raw, const unsigned char* given, size_t total, double* sag, double* mask, cudaStream_t st) {
    sag_split_kernel<<<blocks_for(total), 256, 0, st>>>(raw, given, total, sag, mask);
    return cudaGetLastError();
}
cudaError_t sag_finish(const double* sag, const double* mask, size_t total, double* screen, unsigned char* mask_out, cudaStream_t st) {
    sag_finish_kernel<<<blocks_for(total), 256, 0, st>>>(sag, mask, total, screen, mask_out);
    return cudaGetLastError();
}

}  // namespace paosb
