// fft_core.cuh -- register-resident power-of-two line FFT for sm_100a.
//
// One line of N complex points is held by T = N/E threads, E points per thread, in the "strided"
// distribution  thread t  <->  x[t + j*T], j = 0..E-1.  The transform is a three-stage decimation in
// frequency  N = E * R2 * E  (R2 = 1 for the two-stage sizes): radix-E butterflies in registers, a
// shared-memory exchange, radix-R2 butterflies, a second exchange, radix-E butterflies.  The result comes
// out in natural order in the SAME distribution (thread p <-> X[p + j*T]), so several transforms and
// diagonal factors can be chained in one kernel without touching global memory (pass_kernels.cu).
//
// Replaces numpy.fft (pocketfft) as used by paos/classes/wfo.py:462-472, :493-509, :535-545 and
// paos/classes/psd.py:116,:133.  Forward = exp(-2*pi*i*jk/N), unnormalised; the inverse is obtained by
// swapping real and imaginary parts around the forward transform.
#pragma once
#include <cuda_runtime.h>

namespace paosb {

template <typename R> struct cx2;
template <> struct cx2<double> { typedef double2 type; };
template <> struct cx2<float> { typedef float2 type; };

template <typename R> struct C {
    R x, y;
    __device__ __forceinline__ C() {}
    __device__ __forceinline__ C(R a, R b) : x(a), y(b) {}
};

template <typename R> __device__ __forceinline__ C<R> operator+(C<R> a, C<R> b) { return C<R>(a.x + b.x, a.y + b.y); }
template <typename R> __device__ __forceinline__ C<R> operator-(C<R> a, C<R> b) { return C<R>(a.x - b.x, a.y - b.y); }
template <typename R> __device__ __forceinline__ C<R> operator*(C<R> a, C<R> b) {
    return C<R>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename R> __device__ __forceinline__ C<R> operator*(C<R> a, R s) { return C<R>(a.x * s, a.y * s); }
// multiply by -i
template <typename R> __device__ __forceinline__ C<R> mul_mi(C<R> a) { return C<R>(a.y, -a.x); }

template <typename R> __device__ __forceinline__ C<R> ldc(const C<R>* p) {
    typedef typename cx2<R>::type V;
    V v = *reinterpret_cast<const V*>(p);
    return C<R>(v.x, v.y);
}
template <typename R> __device__ __forceinline__ C<R> ldc_ro(const C<R>* p) {
    typedef typename cx2<R>::type V;
    V v = __ldg(reinterpret_cast<const V*>(p));
    return C<R>(v.x, v.y);
}
// streaming (evict-first) access for the wavefront itself: it is touched once per pass and must not push
// the twiddle and phase tables out of L1/L2
template <typename R> __device__ __forceinline__ C<R> ldc_stream(const C<R>* p);
template <> __device__ __forceinline__ C<double> ldc_stream<double>(const C<double>* p) {
    double x, y;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p));
    return C<double>(x, y);
}
template <> __device__ __forceinline__ C<float> ldc_stream<float>(const C<float>* p) {
    float x, y;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "l"(p));
    return C<float>(x, y);
}
template <typename R> __device__ __forceinline__ void stc_stream(C<R>* p, C<R> a) {
    typedef typename cx2<R>::type V;
    V v;
    v.x = a.x;
    v.y = a.y;
    __stcs(reinterpret_cast<V*>(p), v);
}
template <typename R> __device__ __forceinline__ void stc(C<R>* p, C<R> a) {
    typedef typename cx2<R>::type V;
    V v;
    v.x = a.x;
    v.y = a.y;
    *reinterpret_cast<V*>(p) = v;
}

// ---- small DFTs on registers (forward, natural order in / natural order out) -------------------

template <typename R> __device__ __forceinline__ void dft2(C<R>& a, C<R>& b) {
    C<R> t = a;
    a = t + b;
    b = t - b;
}

template <typename R> __device__ __forceinline__ void dft4(C<R>& a0, C<R>& a1, C<R>& a2, C<R>& a3) {
    C<R> t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = mul_mi(a1 - a3);
    a0 = t0 + t2;
    a2 = t0 - t2;
    a1 = t1 + t3;
    a3 = t1 - t3;
}

template <typename R> __device__ __forceinline__ void dft8(C<R>* v) {
    const R h = (R)0.70710678118654752440084436210485;
    C<R> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    C<R> o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    o1 = C<R>((o1.x + o1.y) * h, (o1.y - o1.x) * h);    // * W8^1
    o2 = mul_mi(o2);                                    // * W8^2
    o3 = C<R>((o3.y - o3.x) * h, -(o3.x + o3.y) * h);   // * W8^3
    v[0] = e0 + o0;
    v[4] = e0 - o0;
    v[1] = e1 + o1;
    v[5] = e1 - o1;
    v[2] = e2 + o2;
    v[6] = e2 - o2;
    v[3] = e3 + o3;
    v[7] = e3 - o3;
}

template <typename R> __device__ __forceinline__ void dft16(C<R>* v) {
    const R h = (R)0.70710678118654752440084436210485;
    const R c1 = (R)0.92387953251128675612818318939679;  // cos(pi/8)
    const R s1 = (R)0.38268343236508977172845998403040;  // sin(pi/8)
    C<R> e[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        e[k] = v[2 * k];
        o[k] = v[2 * k + 1];
    }
    dft8(e);
    dft8(o);
    o[1] = o[1] * C<R>(c1, -s1);
    o[2] = C<R>((o[2].x + o[2].y) * h, (o[2].y - o[2].x) * h);
    o[3] = o[3] * C<R>(s1, -c1);
    o[4] = mul_mi(o[4]);
    o[5] = o[5] * C<R>(-s1, -c1);
    o[6] = C<R>((o[6].y - o[6].x) * h, -(o[6].x + o[6].y) * h);
    o[7] = o[7] * C<R>(-c1, -s1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = e[k] + o[k];
        v[k + 8] = e[k] - o[k];
    }
}

template <int RADIX, typename R> __device__ __forceinline__ void dft(C<R>* v) {
#ifdef PAOS_EXP_NO_DP
    return;
#endif
    if constexpr (RADIX == 2) dft2(v[0], v[1]);
    else if constexpr (RADIX == 4) dft4(v[0], v[1], v[2], v[3]);
    else if constexpr (RADIX == 8) dft8(v);
    else if constexpr (RADIX == 16) dft16(v);
}

// ---- line geometry -------------------------------------------------------------------------------

template <int N_, int E_> struct LineGeom {
    static constexpr int N = N_;
    static constexpr int E = E_;              // points per thread = first and last radix
    static constexpr int T = N / E;           // threads per line
    static constexpr int R2 = N / (E * E);    // middle radix (1 = two-stage transform)
    static constexpr int TP = T + 1;          // padded row length of the exchange buffer (odd)
    static_assert(N % (E * E) == 0, "N must be E*R2*E");
    static_assert(R2 == 1 || R2 == 2 || R2 == 4 || R2 == 8 || R2 == 16, "unsupported middle radix");
    static_assert(R2 <= E, "middle radix must divide E");
    static constexpr int TW1 = (E - 1) * T;                 // stage-1 twiddles  [k1-1][t] = W_N^(k1*t)
    static constexpr int TW2 = (R2 > 1) ? (R2 - 1) * E : 0; // stage-2 twiddles  [k2-1][n3] = W_T^(k2*n3)
    // Line stride in the exchange buffer (complex elements) so that W interleaved lines are bank-conflict free:
    // a shared-memory wavefront covers 128 bytes = SLOTS complex elements (8 for complex128, 16 for complex64);
    // with the column-minor mapping a wavefront holds SLOTS/W consecutive elements of each of the W lines, so the
    // lines must sit SLOTS/W slots apart modulo SLOTS.
    __host__ __device__ static constexpr int line_stride(int W, int elem_bytes) {
        int slots = 128 / elem_bytes;
        int base = E * TP;
        int want = (W >= slots) ? 1 : (W <= 1 ? 0 : slots / W);
        int pad = ((want - (base % slots)) % slots + slots) % slots;
        return base + pad;
    }
};

// Forward transform of one line.  v: E registers in strided distribution.  t: thread index inside the
// line (0..T-1).  sm: this line's exchange buffer.  tw1/tw2: twiddle tables.  SYNC: barrier functor
// covering all threads of the line.
// FIRST_SYNC = false: the caller has already passed a barrier since the last read of the exchange buffer.
//
// In-place exchanges (three-stage sizes, PAOS_INPLACE_EXCHANGE): a barrier is needed after every write of the exchange
// buffer (the readers are other threads), but the two barriers *before* the writes only guard against overwriting
// elements somebody still has to read -- and they disappear if every thread writes exactly the addresses it has just
// read itself.  The middle stage does so by construction (its R2-point butterflies put their results back where the
// operands came from); the first write of the NEXT transform does so if it uses the layout the last read of THIS
// transform left, so consecutive transforms of a pass alternate between two layouts of the buffer (flip = 0 / 1):
//
//   flip 0: stage-1 write  sm[k1][t]              middle read+write  sm[q + R2 c][n2 E + n3]    last read  sm[n3][q E + m]
//   flip 1: stage-1 write  sm[n3][q E + k1]       middle read+write  sm[n3][n2 E + q + R2 c]    last read  sm[m][t]
//
// with t = q E + n3.  Rows are TP = T + 1 (odd) elements long, so the eight lanes of a 16-byte shared-memory phase
// (consecutive t, hence consecutive n3 or consecutive columns) hit eight different bank groups in all six patterns.
// Two barriers per transform instead of four; the arithmetic and its order are unchanged (results are bit-identical).
// tests/test_fft_layouts.py restates the six address maps in numpy and checks the transform, the in-place property and the
// bank groups for every three-stage size.
// MID is called by every thread right after the first barrier (the column kernels start the bulk copy of the next
// phase table there: every thread has consumed the current one before its stage-1 butterfly).
#ifndef PAOS_INPLACE_EXCHANGE
#define PAOS_INPLACE_EXCHANGE 1
#endif
struct NoMid {
    __device__ __forceinline__ void operator()() const {}
};
template <typename G, typename R, typename SYNC, bool FIRST_SYNC = true, typename MID = NoMid>
__device__ __forceinline__ void line_fft_fwd(C<R>* v, int t, C<R>* sm, const C<R>* __restrict__ tw1,
                                             const C<R>* __restrict__ tw2, SYNC sync, int flip = 0, MID mid = MID()) {
    constexpr int E = G::E, T = G::T, R2 = G::R2, TP = G::TP;
    constexpr bool INPLACE = PAOS_INPLACE_EXCHANGE != 0 && R2 > 1;
    // stage 1: radix-E over j, twiddle W_N^(k1*t)
#ifndef PAOS_TABLE_TWIDDLES
    // load only the power-of-two twiddles (issued before the butterfly so their latency hides behind it) and
    // form the others by one or two products (error <= ~3 ulp), instead of E-1 dependent table loads
    C<R> wp[E];
#pragma unroll
    for (int b = 1; b < E; b <<= 1) wp[b] = ldc_ro(tw1 + (b - 1) * T + t);
    dft<E>(v);
#pragma unroll
    for (int k1 = 1; k1 < E; ++k1) {
        const int hb = (k1 >= 8) ? 8 : (k1 >= 4) ? 4 : (k1 >= 2) ? 2 : 1;  // highest set bit
        if (k1 != hb) wp[k1] = wp[hb] * wp[k1 - hb];
#ifndef PAOS_EXP_NO_DP
        v[k1] = v[k1] * wp[k1];
#endif
    }
#else
    dft<E>(v);
#pragma unroll
    for (int k1 = 1; k1 < E; ++k1) v[k1] = v[k1] * ldc_ro(tw1 + (k1 - 1) * T + t);
#endif
    if constexpr (INPLACE) {
        const int n3 = t % E, q = t / E;
        C<R>* const a0 = sm + t;                 // flip 0: stage-1 write, flip 1: last read   (stride TP)
        C<R>* const a1 = sm + n3 * TP + q * E;   // flip 1: stage-1 write, flip 0: last read   (stride 1)
        constexpr int NB = E / R2;               // butterflies per thread
#ifndef PAOS_EXP_NO_SMEM
        if (!flip) {
#pragma unroll
            for (int k1 = 0; k1 < E; ++k1) stc(a0 + k1 * TP, v[k1]);
        } else {
#pragma unroll
            for (int k1 = 0; k1 < E; ++k1) stc(a1 + k1, v[k1]);
        }
        sync();
        mid();
        C<R>* const m0 = sm + q * TP + n3;  // flip 0: + c R2 TP + n2 E
        C<R>* const m1 = sm + n3 * TP + q;  // flip 1: + c R2    + n2 E
        if (!flip) {
#pragma unroll
            for (int c = 0; c < NB; ++c)
#pragma unroll
                for (int n2 = 0; n2 < R2; ++n2) v[c * R2 + n2] = ldc(m0 + c * R2 * TP + n2 * E);
        } else {
#pragma unroll
            for (int c = 0; c < NB; ++c)
#pragma unroll
                for (int n2 = 0; n2 < R2; ++n2) v[c * R2 + n2] = ldc(m1 + c * R2 + n2 * E);
        }
#endif
#pragma unroll
        for (int c = 0; c < NB; ++c) dft<R2>(v + c * R2);
#pragma unroll
        for (int k2 = 1; k2 < R2; ++k2) {
            C<R> w = ldc_ro(tw2 + (k2 - 1) * E + n3);
#pragma unroll
#ifndef PAOS_EXP_NO_DP
            for (int c = 0; c < NB; ++c) v[c * R2 + k2] = v[c * R2 + k2] * w;
#endif
        }
#ifndef PAOS_EXP_NO_SMEM
        if (!flip) {
#pragma unroll
            for (int c = 0; c < NB; ++c)
#pragma unroll
                for (int k2 = 0; k2 < R2; ++k2) stc(m0 + c * R2 * TP + k2 * E, v[c * R2 + k2]);
        } else {
#pragma unroll
            for (int c = 0; c < NB; ++c)
#pragma unroll
                for (int k2 = 0; k2 < R2; ++k2) stc(m1 + c * R2 + k2 * E, v[c * R2 + k2]);
        }
        sync();
        if (!flip) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = ldc(a1 + m);
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = ldc(a0 + m * TP);
        }
#endif
        dft<E>(v);
        return;
    }
#ifndef PAOS_EXP_NO_SMEM
    if constexpr (FIRST_SYNC) sync();  // previous readers of the buffer are done
#pragma unroll
    for (int k1 = 0; k1 < E; ++k1) stc(sm + k1 * TP + t, v[k1]);
    sync();
#endif
    mid();
    if constexpr (R2 == 1) {
        // two-stage: thread p = k1 gathers A[p][n3]
#pragma unroll
#ifndef PAOS_EXP_NO_SMEM
        for (int n3 = 0; n3 < E; ++n3) v[n3] = ldc(sm + t * TP + n3);
#endif
    } else {
        const int n3 = t % E, q = t / E;
        constexpr int NB = E / R2;  // butterflies per thread
#ifndef PAOS_EXP_NO_SMEM
#pragma unroll
        for (int c = 0; c < NB; ++c)
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) v[c * R2 + n2] = ldc(sm + (q + R2 * c) * TP + n2 * E + n3);
#endif
#pragma unroll
        for (int c = 0; c < NB; ++c) dft<R2>(v + c * R2);
#pragma unroll
        for (int k2 = 1; k2 < R2; ++k2) {
            C<R> w = ldc_ro(tw2 + (k2 - 1) * E + n3);
#pragma unroll
#ifndef PAOS_EXP_NO_DP
            for (int c = 0; c < NB; ++c) v[c * R2 + k2] = v[c * R2 + k2] * w;
#endif
        }
#ifndef PAOS_EXP_NO_SMEM
        sync();
#pragma unroll
        for (int c = 0; c < NB; ++c)
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) stc(sm + n3 * TP + k2 * E + (q + R2 * c), v[c * R2 + k2]);
        sync();
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = ldc(sm + m * TP + t);
#endif
    }
    // stage 3: radix-E over n3 -> X[t + k3*T]
    dft<E>(v);
}

}  // namespace paosb
