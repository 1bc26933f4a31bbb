// sag_kernels.h -- host entry points of sag_kernels.cu (device preparation of a Grid Sag map)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace paosb {
// all arrays: rows x cols doubles, row-major; axis 0 runs along rows, 1 along columns
cudaError_t sag_conv_circ(const double* a, int rows, int cols, int axis, const double* h_dev, double sign, int accumulate, double* out,
                          cudaStream_t st);
cudaError_t sag_fir_mirror(const double* a, int rows, int cols, int axis, const double* w_dev, int radius, double* out, cudaStream_t st);
cudaError_t sag_bspline_prefilter(double* a, int rows, int cols, int axis, cudaStream_t st);
cudaError_t sag_bspline_interp(const double* cf, int rows, int cols, int axis, int n_out, double* out, cudaStream_t st);
cudaError_t sag_minmax(const double* a, size_t total, double* lohi_dev, cudaStream_t st);
cudaError_t sag_clip(double* a, size_t total, const double* lohi_dev, cudaStream_t st);
cudaError_t sag_padcrop(const double* a, int rows, int cols, int row_off, int col_off, double fill, int orows, int ocols, double* out,
                        cudaStream_t st);
cudaError_t sag_split(const double* raw, const unsigned char* given, size_t total, double* sag, double* mask, cudaStream_t st);
cudaError_t sag_finish(const double* sag, const double* mask, size_t total, double* screen, unsigned char* mask_out, cudaStream_t st);
}  // namespace paosb
