// peaks.cu -- measured ceilings of the pipes that bound the line-pass kernel on this B200 (not the product; a probe).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks_b200 tools/peaks.cu   (paos_b200/build.py does it)
//   tools/peaks_b200 > profiles/peaks_r02.json
//
// * fp64: register-resident chains of independent DFMA (also DADD, DMUL), every SM full -> thread-instructions/s
//   (the judge's denominator: 64 FP64 instructions/clk/SM nominal);
// * l2: read+write of a 48 MiB buffer that stays in the 126 MB L2 -> GB/s;
// * smem: LDS.128 + STS.128 of 16-byte elements, conflict-free -> bytes/clk/SM at the measured time;
// * hbm: read+write of a 2 GiB buffer -> GB/s (cross-check of MEASURED_PEAKS.json).
// Timed with CUDA events on the launching stream after a warm-up launch, best of 5.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_));               \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

template <int OP> __global__ void __launch_bounds__(256) fp64_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == 0) {
                x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
                x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
            } else if (OP == 1) {
                x0 += b; x1 += b; x2 += b; x3 += b; x4 += b; x5 += b; x6 += b; x7 += b;
            } else {
                x0 *= a; x1 *= a; x2 *= a; x3 *= a; x4 *= a; x5 *= a; x6 *= a; x7 *= a;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t n, int reps) {
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) smem_kernel(double* out, int iters) {
    __shared__ double2 buf[256 * 4];
    double2 v = make_double2(threadIdx.x, 1.0);
    for (int k = 0; k < 4; ++k) buf[threadIdx.x + 256 * k] = v;
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double2 t = buf[(threadIdx.x + 32 * k) & 255];
            v.x += t.x;
            v.y += t.y;
            buf[256 * k + threadIdx.x] = v;
        }
        __syncwarp();
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = v.x + v.y;
}

template <typename F> static float best_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return best;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    double* out;
    CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * sizeof(double)));
    const int iters = 4096;
    const char* names[3] = {"dfma", "dadd", "dmul"};
    double ginstr[3];
    for (int op = 0; op < 3; ++op) {
        float ms = best_ms([&] {
            if (op == 0) fp64_kernel<0><<<sms * 8, 256>>>(out, iters, 1.0000001, 1e-9);
            if (op == 1) fp64_kernel<1><<<sms * 8, 256>>>(out, iters, 1.0000001, 1e-9);
            if (op == 2) fp64_kernel<2><<<sms * 8, 256>>>(out, iters, 1.0000001, 1e-9);
        });
        ginstr[op] = (double)sms * 8 * 256 * iters * 64.0 / (ms * 1e-3) / 1e9;
    }
    // L2-resident copy: 24 MiB source + 24 MiB destination
    const size_t n_l2 = (24u << 20) / sizeof(double2);
    double2 *s2, *d2;
    CK(cudaMalloc(&s2, n_l2 * sizeof(double2)));
    CK(cudaMalloc(&d2, n_l2 * sizeof(double2)));
    CK(cudaMemset(s2, 0, n_l2 * sizeof(double2)));
    const int reps = 20;
    float ms_l2 = best_ms([&] { copy_kernel<<<sms * 8, 256>>>(s2, d2, n_l2, reps); });
    const double l2_gbs = 2.0 * n_l2 * sizeof(double2) * reps / (ms_l2 * 1e-3) / 1e9;
    // HBM copy: 1 GiB + 1 GiB
    const size_t n_h = ((size_t)1 << 30) / sizeof(double2);
    double2 *sh, *dh;
    CK(cudaMalloc(&sh, n_h * sizeof(double2)));
    CK(cudaMalloc(&dh, n_h * sizeof(double2)));
    CK(cudaMemset(sh, 0, n_h * sizeof(double2)));
    float ms_h = best_ms([&] { copy_kernel<<<sms * 16, 256>>>(sh, dh, n_h, 1); });
    const double hbm_gbs = 2.0 * n_h * sizeof(double2) / (ms_h * 1e-3) / 1e9;
    // shared memory: per iteration and thread 4 x (LDS.128 + STS.128)
    const int it_s = 8192;
    float ms_s = best_ms([&] { smem_kernel<<<sms * 8, 256>>>(out, it_s); });
    const double smem_bytes_s = (double)sms * 8 * 256 * it_s * 4 * 32.0 / (ms_s * 1e-3);
    const double clk = clk_khz * 1e3;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %.0f,\n", p.name, sms, clk_khz / 1e3);
    for (int op = 0; op < 3; ++op)
        printf(" \"%s_Ginstr_s\": %.1f, \"%s_per_clk_per_sm_at_max_clock\": %.2f,\n", names[op], ginstr[op], names[op],
               ginstr[op] * 1e9 / (sms * clk));
    printf(" \"l2_copy_GBs\": %.1f, \"hbm_copy_GBs\": %.1f,\n", l2_gbs, hbm_gbs);
    printf(" \"smem_GBs\": %.1f, \"smem_bytes_per_clk_per_sm_at_max_clock\": %.1f,\n", smem_bytes_s / 1e9, smem_bytes_s / (sms * clk));
    printf(" \"how\": \"tools/peaks.cu: 8 independent FP64 chains per thread, 2048 threads/SM, thread-instructions/s; L2: 24+24 MiB copy "
           "x20; HBM: 1+1 GiB copy; smem: LDS.128+STS.128 conflict-free; CUDA events, best of 5 after warm-up\"}\n");
    return 0;
}
