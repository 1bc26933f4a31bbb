// aux_kernels.h -- host entry points of aux_kernels.cu
#pragma once
#include "device_types.h"

namespace paosb {
cudaError_t launch_build_tables(const TableBlock& B, cudaStream_t st);
cudaError_t launch_build_edge_tables(const EdgeBlock& B, cudaStream_t st);
// one stop reduction (paos/classes/wfo.py:200): sum |src * masks|^2 -> out_slot = (1/sqrt(sum), sum)
struct Norm2Item {
    const void* src;  // null: the analytic field of ones
    int ngen;
    GenOp gen[GMAX];
    double* partials;
    double* out_slot;
};
cudaError_t launch_norm2_batch(const Norm2Item* items, int nb, int n, int dtype, int npartials, cudaStream_t st);
cudaError_t launch_norm2(const void* src, int n, int dtype, const GenOp* gen, int ngen, double* partials,
                         int npartials, double* out_slot, cudaStream_t st);
cudaError_t launch_zero_outside_band(void* field, int n, int dtype, int axis, int lo, int hi, cudaStream_t st);
cudaError_t launch_readout(const void* src, int n, int dtype, int what, void* out, cudaStream_t st);
cudaError_t launch_zernike(const ZernParams& Z, const unsigned char* mask, double* out, cudaStream_t st);
cudaError_t launch_zernike_cov(const ZernParams& Z, const unsigned char* mask, double* partial, int blocks, cudaStream_t st);
cudaError_t launch_zernike_points(const ZernParams& Z, const double* rho, const double* phi, const unsigned char* mask,
                                  size_t npoints, double* out, cudaStream_t st);
cudaError_t launch_stack_cov(const double* stack, const double* rho, const unsigned char* mask, int K, size_t npoints, double* out,
                             cudaStream_t st);
cudaError_t launch_stack_transform(const double* stack, const double* mat, int K, size_t npoints, double* out, cudaStream_t st);
cudaError_t launch_real_to_complex(const double* src, int n, int dtype, void* dst, cudaStream_t st);
cudaError_t launch_psd_finalize(const void* f, int n, int dtype, const double* noise2, double SR, double unit,
                                double* out, cudaStream_t st);
cudaError_t launch_encircled_energy(const void* psf, int n, int real_is_float, double dx, double dy, double xc, double yc,
                                    double inv_bin, int nbins, double* hist, double* ee, cudaStream_t st);
cudaError_t launch_psf_peak(const void* psf, int n, int real_is_float, double* out, cudaStream_t st);
cudaError_t launch_screen_stats(const double* screen, int n, double radius, double dx, double dy, double* out, cudaStream_t st);
cudaError_t launch_crop_convert(const void* src, int n, int src_is_float, int x0, int y0, int nx, int ny, int dst_is_float, void* dst,
                                cudaStream_t st);
cudaError_t launch_normal(uint64_t seed, uint32_t stream_id, int n, double* out, cudaStream_t st);
}  // namespace paosb
