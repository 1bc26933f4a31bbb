// runtime.cu -- host side of libpaos_b200.so: the wavefront handle, the operation queue, the planner that
// turns a run of recorded operations into line passes, and the C ABI of include/paos_b200.h.
//
// Planner in one paragraph.  Every recorded operation is either an FFT2, a *separable* diagonal factor
// (fftshift signs, Fresnel chirps, lens phase, the ortho scales: f(x)*g(y)), or a *general* pointwise factor
// (aperture masks, phase screens, the stop scalar).  An FFT2 is FFTx * FFTy and a separable factor is
// Dx * Dy, and everything that acts along x commutes with everything that acts along y.  The queue is
// therefore split into an x-thread and a y-thread of 1-D items that only have to meet at the general
// factors ("barriers").  A row pass executes the x-thread up to a barrier the y-thread has not reached
// yet, a column pass does the same for the y-thread, and the pass that arrives second applies the general
// factor in registers and carries on.  A chain of K FFT2s with G general factors between them costs about
// G + 2 sweeps of the wavefront through HBM instead of 2K (plus ~10 elementwise sweeps per primitive in
// the reference, paos/classes/wfo.py:462-472).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/paos_b200.h"
#include "aux_kernels.h"
#include "device_types.h"
#include "pass_dispatch.h"
#include "sag_kernels.h"

using namespace paosb;

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU(expr)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(PAOS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// twiddle tables (one set per device, grid size and precision; shared by all handles)
// ---------------------------------------------------------------------------------------------------
struct Twiddles {
    void* tw1 = nullptr;
    void* tw2 = nullptr;
};
static std::mutex g_tw_mutex;
static std::map<long, Twiddles> g_tw;

static void unit_root(long k, long n, double& c, double& s) {
    // exp(-2*pi*i*k/n) with the argument reduced to the first octant in exact integer arithmetic
    k %= n;
    if (k < 0) k += n;
    const long double PI = 3.14159265358979323846264338327950288L;
    long oct = (8 * k) / n;               // octant 0..7
    long double cc, ss;
    long r8 = 8 * k - oct * n;            // position inside the octant, in units of 2*pi/(8n)
    long double a = (2.0L * PI * (long double)r8) / (8.0L * (long double)n);
    long double b = (2.0L * PI * (long double)(n - r8)) / (8.0L * (long double)n);  // pi/4 - a
    switch (oct) {
        case 0: cc = cosl(a); ss = sinl(a); break;
        case 1: cc = sinl(b); ss = cosl(b); break;
        case 2: cc = -sinl(a); ss = cosl(a); break;
        case 3: cc = -cosl(b); ss = sinl(b); break;
        case 4: cc = -cosl(a); ss = -sinl(a); break;
        case 5: cc = -sinl(b); ss = -cosl(b); break;
        case 6: cc = sinl(a); ss = -cosl(a); break;
        default: cc = cosl(b); ss = -sinl(b); break;
    }
    c = (double)cc;
    s = (double)(-ss);
}

static int get_twiddles(int device, int n, int dtype, Twiddles& out) {
    std::lock_guard<std::mutex> lock(g_tw_mutex);
    const long key = ((long)device << 32) | ((long)n << 2) | dtype;
    auto it = g_tw.find(key);
    if (it != g_tw.end()) {
        out = it->second;
        return PAOS_OK;
    }
    const int E = geom_E(n), T = n / E, R2 = n / (E * E);
    const size_t n1 = (size_t)(E - 1) * T, n2 = (R2 > 1) ? (size_t)(R2 - 1) * E : 1;
    std::vector<double> h1(2 * n1), h2(2 * n2, 0.0);
    for (int k1 = 1; k1 < E; ++k1)
        for (int t = 0; t < T; ++t) unit_root((long)k1 * t, n, h1[2 * ((size_t)(k1 - 1) * T + t)], h1[2 * ((size_t)(k1 - 1) * T + t) + 1]);
    if (R2 > 1)
        for (int k2 = 1; k2 < R2; ++k2)
            for (int n3 = 0; n3 < E; ++n3) unit_root((long)k2 * n3, T, h2[2 * ((size_t)(k2 - 1) * E + n3)], h2[2 * ((size_t)(k2 - 1) * E + n3) + 1]);
    Twiddles tw;
    const size_t es = dtype == PAOS_C128 ? sizeof(double) : sizeof(float);
    CU(cudaMalloc(&tw.tw1, 2 * n1 * es));
    CU(cudaMalloc(&tw.tw2, 2 * n2 * es));
    if (dtype == PAOS_C128) {
        CU(cudaMemcpy(tw.tw1, h1.data(), 2 * n1 * es, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(tw.tw2, h2.data(), 2 * n2 * es, cudaMemcpyHostToDevice));
    } else {
        std::vector<float> f1(h1.begin(), h1.end()), f2(h2.begin(), h2.end());
        CU(cudaMemcpy(tw.tw1, f1.data(), 2 * n1 * es, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(tw.tw2, f2.data(), 2 * n2 * es, cudaMemcpyHostToDevice));
    }
    g_tw[key] = tw;
    out = tw;
    return PAOS_OK;
}

// ---------------------------------------------------------------------------------------------------
// the operation queue
// ---------------------------------------------------------------------------------------------------
enum OpKind { OP_FFT = 1, OP_SIGN = 2, OP_PHASE = 3, OP_GEN = 4, OP_FFT_RAW = 5 };

struct Op {
    int kind;
    int dir;            // OP_FFT / OP_FFT_RAW: +1 forward, -1 inverse
    double raw_scale;   // OP_FFT_RAW: real scale applied with the x half
    TableTerm tx, ty;   // OP_PHASE: the x and y halves
    GenOp gen;          // OP_GEN
};

// items of one axis thread
enum ItemKind { IT_TERM = 1, IT_SIGN = 2, IT_SCALE = 3, IT_FFT = 4, IT_BARRIER = 5 };
struct Item {
    int kind;
    int dir;
    int gen;  // index into the gen list for barriers
    double scale;
    TableTerm term;
};

struct PosAcc {  // what accumulates at one position of a pass
    int nterms = 0;
    bool sign = false;
    double scale = 1.0;
    TableTerm terms[TERM_MAX];
    bool trivial() const { return nterms == 0; }  // a real scale and/or the (-1)^k sign: no table needed
};

// Virtual zeros.  A pass whose tiles outside [tile_lo, tile_hi] end up exactly zero does not store those zeros: the
// memory of those lines is stale and the planner remembers that the field is zero outside a band of rows (axis 0, left
// by a row pass) or of columns (axis 1).  The next pass folds the band into its tile range (same axis) or skips the
// loads outside it (other axis); any other reader of the field has the zeros written first (materialize_band).
struct ZeroBand {
    bool valid = false;
    int axis = 0;
    int lo = 0, hi = -1;  // lines [lo, hi] may be non-zero (lo > hi: the whole field is zero)
};

struct TimedLaunch {
    cudaEvent_t a, b;
    int nfft;   // line FFTs chained, summed over the wavefronts of a batched launch
    int col;
    int items;  // wavefronts served by the launch
};

// Deferred execution.  While a handle is *recording*, every device launch the library would make for it is appended to
// its program instead; paos_batch_execute then walks the programs of up to BMAX handles in lockstep on one stream and
// issues ONE launch for the heads that are of the same kind (the same-axis passes of several wavelengths, their table
// builds, their stop reductions).  Inside one handle the order of the records is the order of the stream, so everything
// that relies on stream order (table pool reuse per flush, screen recycling, stop scalars) holds unchanged.
enum RecKind { REC_TABLES = 1, REC_PASS = 2, REC_NORM2 = 3, REC_FN = 4 };
struct Rec {
    int kind = 0;
    bool col = false;
    PassParams P;                         // REC_PASS
    std::vector<TableSpec> specs;         // REC_TABLES
    std::vector<EdgeSpec> edges;          // REC_TABLES: edge tables of the elliptical masks, built with the phase tables
    Norm2Item norm;                       // REC_NORM2
    std::function<int(cudaStream_t)> fn;  // REC_FN: anything else (screens, uploads, zero fill), run per handle
};

struct paos_wfo {
    int n = 0, dtype = 0, device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    void* field = nullptr;
    bool own_field = false;
    bool materialized = false;  // false: the field is all ones and lives nowhere yet
    ZeroBand band;              // virtual zeros of `field` (see ZeroBand)
    bool dropped = false;       // a final read-out skipped the field store: only reset / upload make the handle usable again
    size_t elem = 16;           // bytes per complex element
    Twiddles tw;
    std::vector<Op> ops;

    // along-line table pool (bump allocated per flush; stream order makes reuse safe)
    char* tab_pool = nullptr;
    size_t tab_cap = 0, tab_used = 0;
    // N*N double buffers (phase screens, read-out staging, PSD scratch)
    std::vector<double*> screens_free, screens_busy;
    void* scratch_field = nullptr;  // complex scratch for the PSD screen
    double* partials = nullptr;
    double* slots = nullptr;  // stop scalars: pairs (1/sqrt(sum), sum)
    double* ee_hist = nullptr;  // radial histogram of paos_encircled_energy
    int slot_next = 0;
    static constexpr int NSLOTS = 256;
    static constexpr int NPARTIALS = 148 * 8;

    bool recording = false;
    std::vector<Rec> program;
    std::vector<void*> retired_pools;  // table pools outgrown while recording: still referenced by the program
    std::map<std::pair<const void*, int>, CUtensorMap*> tmaps;  // tensor maps of the field buffers (column tiles), built on first use
    std::vector<CUtensorMap*> retired_tmaps;                    // dropped from the cache while recorded passes may still use them

    paos_stats stats{};
    bool timing = false;
    std::vector<TimedLaunch> timed;
    std::vector<cudaEvent_t> event_pool;
    double timed_ms[2][KMAX + 1] = {};
    uint64_t timed_n[2][KMAX + 1] = {};
    double timed_total_ms = 0.0;       // since the last paos_wfo_timing_totals(reset)
    uint64_t timed_total_launches = 0, timed_total_sweeps = 0, timed_total_items = 0;
};

static int set_device(paos_wfo* w) {
    CU(cudaSetDevice(w->device));
    return PAOS_OK;
}

// Tensor map of an n x n complex field for the column passes' TMA tile stores: a 2-D tensor of reals (2n per row), boxes of
// 2W reals x 256 rows.  cuTensorMapEncodeTiled is fetched from the driver at run time (no link against libcuda); null when
// the variant is not compiled in or the driver refuses.
static const CUtensorMap* field_tmap(paos_wfo* w, const void* field, bool real_readout = false, bool wide = false) {
#if PAOS_TMA_FIELD
    if (w->n < PAOS_TMA_FIELD_MIN_N || !field) return nullptr;
    const std::pair<const void*, int> key(field, (real_readout ? 1 : 0) | (wide ? 2 : 0));
    auto it = w->tmaps.find(key);
    if (it != w->tmaps.end()) return it->second;
    if (w->tmaps.size() > 4096) {  // read-out destinations come and go: keep the cache bounded
        // passes that are planned or recorded but not yet launched still point at these maps: retire them, free them at the
        // next paos_wfo_sync / destroy (128 bytes each)
        for (auto& kv : w->tmaps)
            if (kv.second) w->retired_tmaps.push_back(kv.second);
        w->tmaps.clear();
    }
    typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Encode encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        cudaGetLastError();
        return (Encode)fn;
    }();
    CUtensorMap* tm = nullptr;
    if (encode) {
        const int W = tile_width(w->n, w->dtype, true, wide);
        const bool c128 = w->dtype == PAOS_C128;
        const int per = real_readout ? 1 : 2;  // reals per pixel: a read-out (|.|, angle, |.|^2) or the complex field
        const cuuint64_t gdim[2] = {(cuuint64_t)per * w->n, (cuuint64_t)w->n};
        const cuuint64_t gstride[1] = {(cuuint64_t)w->n * (w->elem / 2) * per};
        const cuuint32_t box[2] = {(cuuint32_t)(per * W), 256u};
        const cuuint32_t estr[2] = {1u, 1u};
        tm = new CUtensorMap;
        const size_t row_bytes = (size_t)per * W * (w->elem / 2);  // the TMA unit moves rows of 16 bytes or more
        CUresult r = row_bytes < 16 ? CUDA_ERROR_INVALID_VALUE : encode(tm, c128 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(field), gdim, gstride, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            delete tm;
            tm = nullptr;
        }
    }
    w->tmaps[key] = tm;  // null is remembered too: the direct stores are used
    return tm;
#else
    (void)w;
    (void)field;
    return nullptr;
#endif
}

static int alloc_table(paos_wfo* w, size_t bytes, void** out) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (w->tab_used + bytes > w->tab_cap) return fail(PAOS_ERR_STATE, "table pool exhausted");
    *out = w->tab_pool + w->tab_used;
    w->tab_used += bytes;
    return PAOS_OK;
}

static int get_screen(paos_wfo* w, double** out) {
    if (!w->screens_free.empty()) {
        *out = w->screens_free.back();
        w->screens_free.pop_back();
    } else {
        CU(cudaMalloc((void**)out, (size_t)w->n * w->n * sizeof(double)));
    }
    w->screens_busy.push_back(*out);
    return PAOS_OK;
}

// A screen that is written immediately (not through a deferred record) while the handle is recording must not come from the
// recycled pool: passes recorded earlier, not yet executed, may still read those buffers.  Such a screen gets memory of its
// own, released at the next paos_wfo_sync / destroy.
static int get_screen_for_immediate_write(paos_wfo* w, double** out) {
    if (!w->recording) return get_screen(w, out);
    CU(cudaMalloc((void**)out, (size_t)w->n * w->n * sizeof(double)));
    w->retired_pools.push_back(*out);
    return PAOS_OK;
}

static void recycle_screens(paos_wfo* w) {
    // every consumer of a busy screen has been enqueued; later writers are ordered behind them on the stream
    for (double* p : w->screens_busy) w->screens_free.push_back(p);
    w->screens_busy.clear();
}

// ---------------------------------------------------------------------------------------------------
// planner
// ---------------------------------------------------------------------------------------------------
static void split_ops(const paos_wfo* w, const std::vector<Op>& ops, std::vector<Item>& X, std::vector<Item>& Y,
                      std::vector<GenOp>& gens) {
    const double ortho = 1.0 / std::sqrt((double)w->n);
    for (const Op& op : ops) {
        Item ix{}, iy{};
        switch (op.kind) {
            case OP_FFT:
                ix.kind = iy.kind = IT_FFT;
                ix.dir = iy.dir = op.dir;
                X.push_back(ix);
                Y.push_back(iy);
                ix.kind = iy.kind = IT_SCALE;
                ix.scale = iy.scale = ortho;
                X.push_back(ix);
                Y.push_back(iy);
                break;
            case OP_FFT_RAW:
                ix.kind = iy.kind = IT_FFT;
                ix.dir = iy.dir = op.dir;
                X.push_back(ix);
                Y.push_back(iy);
                if (op.raw_scale != 1.0) {
                    ix.kind = IT_SCALE;
                    ix.scale = op.raw_scale;
                    X.push_back(ix);
                }
                break;
            case OP_SIGN:
                ix.kind = iy.kind = IT_SIGN;
                X.push_back(ix);
                Y.push_back(iy);
                break;
            case OP_PHASE:
                ix.kind = iy.kind = IT_TERM;
                ix.term = op.tx;
                iy.term = op.ty;
                X.push_back(ix);
                Y.push_back(iy);
                break;
            case OP_GEN:
                ix.kind = iy.kind = IT_BARRIER;
                ix.gen = iy.gen = (int)gens.size();
                gens.push_back(op.gen);
                X.push_back(ix);
                Y.push_back(iy);
                break;
        }
    }
}

static bool is_diag(const Item& it) { return it.kind == IT_TERM || it.kind == IT_SIGN || it.kind == IT_SCALE; }

static bool acc_item(PosAcc& a, const Item& it) {
    if (it.kind == IT_SIGN) a.sign = !a.sign;
    else if (it.kind == IT_SCALE) a.scale *= it.scale;
    else {
        if (a.nterms == TERM_MAX) return false;
        a.terms[a.nterms++] = it.term;
    }
    return true;
}

struct PlannedPass {
    bool col;
    PassParams P;
};

struct Plan {
    std::vector<PlannedPass> passes;
    std::vector<TableSpec> specs;
    std::vector<EdgeSpec> edges;
};

// edge table of elliptical mask g for a pass along `axis` (0 rows, 1 columns): where it goes and which lines it covers
static EdgeSpec make_edge_spec(const paos_wfo* w, GenOp& g, void* mem, int axis) {
    const double c_cr = axis == 1 ? g.p0 : g.p1, s_cr = axis == 1 ? g.p2 : g.p3;
    const double reach = (std::sqrt(g.p6) + 1e-3) / s_cr + 2.0;
    int lo = 0, hi = w->n - 1;
    if (std::isfinite(reach) && std::isfinite(c_cr)) {
        lo = (int)std::max(0.0, std::min((double)w->n, std::floor(c_cr - reach)));
        hi = (int)std::min((double)(w->n - 1), std::max(-1.0, std::ceil(c_cr + reach)));
    }
    g.ptr0 = mem;
    g.p7 = (double)lo;
    g.p8 = (double)hi;
    EdgeSpec es{};
    es.g = g;
    es.out = mem;
    es.col = axis == 1 ? 1 : 0;
    es.T = w->n / geom_E(w->n);
    es.line_lo = lo;
    es.line_hi = hi;
    return es;
}

static int spec_from_acc(paos_wfo* w, const PosAcc& a, Plan& plan, const void** tab_out) {
    void* mem;
    int rc = alloc_table(w, (size_t)w->n * w->elem, &mem);
    if (rc) return rc;
    TableSpec sp{};
    sp.out = mem;
    sp.kind = TABLE_PHASE;
    sp.nterms = a.nterms;
    sp.sign = a.sign ? 1 : 0;
    sp.scale = a.scale;
    for (int i = 0; i < a.nterms; ++i) sp.terms[i] = a.terms[i];
    plan.specs.push_back(sp);
    *tab_out = mem;
    return PAOS_OK;
}

// Build the passes for the queued ops.  readout/dst_real: fused read-out for the final pass (0 = none).
// band: virtual zeros of `field` on entry, updated to the state after the last planned pass.
static int build_plan(paos_wfo* w, const std::vector<Op>& ops, Plan& plan, int readout, void* dst_real, void* field, ZeroBand& band) {
    static const bool zero_fill = getenv("PAOS_ZERO_FILL") != nullptr;  // diagnostic: store the zeros, keep no band
    // fold the band into a pass on `axis` whose own apertures leave lines [line_lo, line_hi], and advance the band
    auto finish_pass = [&](PassParams& P, int axis, int line_lo, int line_hi) {
        const int n = w->n;
        P.in_lo = 0;
        P.in_hi = n - 1;
        P.out_lo = 0;
        P.out_hi = n - 1;
        if (band.valid && P.src) {
            if (band.axis == axis) {  // lines outside the band are zero on input, hence on output
                line_lo = std::max(line_lo, band.lo);
                line_hi = std::min(line_hi, band.hi);
            } else {  // every line crosses the band: only that stretch is loaded
                P.in_lo = band.hi >= band.lo ? band.lo : n;  // empty band: [n, n] matches no index (the kernel tests
                P.in_hi = band.hi >= band.lo ? band.hi : n;  // (unsigned)(idx - in_lo) <= (unsigned)(in_hi - in_lo))
            }
        }
        P.tile_lo = 0;
        P.tile_hi = 0x7fffffff;
        P.zero_fill = zero_fill ? 1 : 0;
        band.valid = false;
        if (line_lo > 0 || line_hi < n - 1) {
            const int W = tile_width(n, w->dtype, axis == 1);
            P.tile_lo = line_lo / W;
            P.tile_hi = line_hi >= line_lo ? line_hi / W : -1;  // empty range: everything is blank
            if (!zero_fill) {
                band.valid = true;
                band.axis = axis;
                band.lo = P.tile_lo * W;
                band.hi = P.tile_hi < 0 ? -1 : std::min(n - 1, P.tile_hi * W + W - 1);
            }
        }
    };
    std::vector<Item> th[2];
    std::vector<GenOp> gens;
    split_ops(w, ops, th[0], th[1], gens);
    size_t cur[2] = {0, 0};
    bool have_src = w->materialized;
    int axis = 0;
    int idle = 0;
    while (cur[0] < th[0].size() || cur[1] < th[1].size()) {
        std::vector<Item>& mine = th[axis];
        std::vector<Item>& other = th[1 - axis];
        size_t& c = cur[axis];
        size_t& oc = cur[1 - axis];
        PosAcc acc[KMAX + 1];
        PassParams P{};
        P.src = have_src ? field : nullptr;
        P.dst = field;
        int pos = 0;
        bool any = false;
        while (c < mine.size()) {
            const Item& it = mine[c];
            if (is_diag(it)) {
                if (!acc_item(acc[pos], it)) break;  // table full: end the pass here
                ++c;
                any = true;
            } else if (it.kind == IT_FFT) {
                if (P.nfft == KMAX) break;
                P.dir[P.nfft++] = it.dir;
                ++pos;
                ++c;
                any = true;
            } else {  // barrier: the other thread must have nothing but diagonal items before the same barrier
                size_t j = oc;
                while (j < other.size() && is_diag(other[j])) ++j;
                if (j < other.size() && other[j].kind == IT_BARRIER && other[j].gen == it.gen) {
                    if (P.ngen == GMAX) break;
                    GenOp g = gens[it.gen];
                    g.pos = pos;
                    P.gen[P.ngen++] = g;
                    P.genmask |= 1u << pos;
                    other.erase(other.begin() + j);  // diagonal items before it commute past the barrier
                    ++c;
                    any = true;
                } else {
                    break;  // the other axis has to catch up first
                }
            }
        }
        if (!any) {
            if (++idle > 2) return fail(PAOS_ERR_STATE, "planner made no progress");
            axis = 1 - axis;
            continue;
        }
        idle = 0;
        // if this thread is finished and the other one has only diagonal items left, fold them in as a
        // per-line factor so that no extra sweep is needed
        if (c == mine.size()) {
            size_t j = oc;
            while (j < other.size() && is_diag(other[j])) ++j;
            if (j == other.size() && oc < other.size()) {
                PosAcc cross;
                bool ok = true;
                for (size_t k = oc; k < other.size() && ok; ++k) ok = acc_item(cross, other[k]);
                if (ok) {
                    int rc = spec_from_acc(w, cross, plan, &P.ctab_out);
                    if (rc) return rc;
                    oc = other.size();
                }
            }
        }
        // lines that miss the bounding box of an elliptical *aperture* of this pass end up exactly zero: a pixel whose
        // centre lies at normalised distance >= 1 + d from the centre (d = half the pixel diagonal in the frame where
        // the ellipse is the unit circle, sqrt(p6) = 1 + d) is entirely outside, and every pixel of such a line is
        int line_lo = 0, line_hi = w->n - 1;
        for (int gi = 0; gi < P.ngen; ++gi) {
            const GenOp& g = P.gen[gi];
            if (g.kind != GEN_ELLIPSE || g.flag) continue;
            const double c0 = axis == 1 ? g.p0 : g.p1, s = axis == 1 ? g.p2 : g.p3;  // column pass: line = ix
            const double reach = (std::sqrt(g.p6) + 1e-9) / s;
            if (!std::isfinite(reach) || !std::isfinite(c0)) continue;
            line_lo = std::max(line_lo, (int)std::ceil(c0 - reach));
            line_hi = std::min(line_hi, (int)std::floor(c0 + reach));
        }
        finish_pass(P, axis, line_lo, line_hi);
        // edge tables: the exact overlap of the rim pixels of every elliptical mask of this pass, evaluated once by a
        // small kernel instead of on the critical path of the lines that meet them (device_types.h: EdgeSpec)
        static const bool edge_tables = getenv("PAOS_NO_EDGE_TABLES") == nullptr;
        for (int gi = 0; gi < P.ngen && edge_tables; ++gi) {
            GenOp& g = P.gen[gi];
            if (g.kind != GEN_ELLIPSE) continue;
            void* mem;
            if (alloc_table(w, edge_table_bytes(w->n), &mem) != PAOS_OK) {
                g_last_error.clear();
                continue;  // pool exhausted: the pass kernel evaluates these pixels itself
            }
            plan.edges.push_back(make_edge_spec(w, g, mem, axis));
        }
        // A real scale commutes with everything a pass does (line transforms and diagonal factors are linear), so a position
        // that carries nothing but a scale -- the 1/sqrt(N) of an orthonormal transform with no phase factor next to it --
        // hands it to a position of the same pass that has a table anyway (tables carry their position's scale already) and
        // saves E real multiplications per thread.  The product of the two scales is rounded once more: ~1e-16 relative.
        static const bool fold_scales = getenv("PAOS_NO_SCALE_FOLD") == nullptr;
        int host = -1;
        for (int p = 0; p <= P.nfft && host < 0; ++p)
            if (!acc[p].trivial()) host = p;
        for (int p = 0; p <= P.nfft && fold_scales && host >= 0; ++p) {
            if (!acc[p].trivial() || acc[p].scale == 1.0) continue;
            acc[host].scale *= acc[p].scale;
            acc[p].scale = 1.0;
        }
        for (int p = 0; p <= P.nfft; ++p) {
            P.scl[p] = 1.0;
            P.tab[p] = nullptr;
            if (acc[p].trivial()) {
                P.scl[p] = acc[p].scale;
                if (acc[p].sign) P.sgnmask |= 1u << p;
            } else {
                int rc = spec_from_acc(w, acc[p], plan, &P.tab[p]);
                if (rc) return rc;
            }
        }
        if (axis == 1) P.tmap_host = field_tmap(w, field);
        PlannedPass pp;
        pp.col = axis == 1;
        pp.P = P;
        plan.passes.push_back(pp);
        have_src = true;
        axis = 1 - axis;
    }
    if (plan.passes.empty() && (!w->materialized || readout)) {
        // nothing queued: materialise the field of ones / run a pure read-out sweep
        PassParams P{};
        P.src = w->materialized ? field : nullptr;
        P.dst = field;
        P.scl[0] = 1.0;
        finish_pass(P, 0, 0, w->n - 1);
        PlannedPass pp;
        pp.col = false;
        pp.P = P;
        plan.passes.push_back(pp);
    }
    if (readout && !plan.passes.empty()) {
        PlannedPass& last = plan.passes.back();
        last.P.readout = readout;
        last.P.dst_real = dst_real;
        if (last.col && has_wide_tiles(w->n, w->dtype)) {
            // a column pass with a fused read-out runs on the wide-tile kernel (pass_dispatch.h): recount its tiles, and the
            // band it leaves, in that width (a superset of the lines: the extra ones are transformed and come out as zeros)
            const int n = w->n, Wn = tile_width(n, w->dtype, true, false), Ww = tile_width(n, w->dtype, true, true);
            PassParams& P = last.P;
            P.wide = 1;
            if (P.tile_lo > 0 || P.tile_hi != 0x7fffffff) {
                const long long line_lo = (long long)std::max(0, P.tile_lo) * Wn;
                const long long line_hi = P.tile_hi < 0 ? -1 : std::min<long long>(n - 1, (long long)P.tile_hi * Wn + Wn - 1);
                P.tile_lo = (int)(line_lo / Ww);
                P.tile_hi = line_hi >= line_lo ? (int)(line_hi / Ww) : -1;
                if (band.valid) {
                    band.lo = P.tile_lo * Ww;
                    band.hi = P.tile_hi < 0 ? -1 : std::min(n - 1, P.tile_hi * Ww + Ww - 1);
                }
            }
            P.tmap_host = field_tmap(w, field, false, true);
        }
        if (last.col) last.P.tmap_real_host = field_tmap(w, dst_real, true, last.P.wide != 0);
    }
    // Look ahead: a pass followed, inside this plan, by a pass along the other axis is read by that pass only, and that
    // pass touches nothing outside its tile range (its blank tiles are neither loaded nor stored).  Along the lines of the
    // first pass that range is a stretch of indices: only that stretch is stored (a zoom-4 beam: a quarter to 40 % of
    // every line).  The last pass of a plan stores everything: whoever reads the field next is not known here.
    static const bool store_ahead = getenv("PAOS_NO_STORE_RANGE") == nullptr;
    for (size_t i = 0; store_ahead && !zero_fill && i + 1 < plan.passes.size(); ++i) {
        PlannedPass& a = plan.passes[i];
        const PlannedPass& b = plan.passes[i + 1];
        if (a.col == b.col || !a.P.dst || b.P.src != a.P.dst || b.P.zero_fill) continue;
        if (b.P.tile_lo <= 0 && b.P.tile_hi == 0x7fffffff) continue;
        const int n = w->n, Wb = tile_width(n, w->dtype, b.col, b.P.wide != 0);
        const int lo = std::max(0, b.P.tile_lo) * Wb;
        const int hi = b.P.tile_hi < 0 ? -1 : (int)std::min<long long>(n - 1, (long long)b.P.tile_hi * Wb + Wb - 1);
        a.P.out_lo = hi >= lo ? lo : n;  // empty: [n, n] matches no index
        a.P.out_hi = hi >= lo ? hi : n;
    }
    static const bool debug_plan = getenv("PAOS_DEBUG_PLAN") != nullptr;
    if (debug_plan) {
        for (const PlannedPass& pp : plan.passes) {
            int ntab = 0;
            for (int p2 = 0; p2 <= pp.P.nfft; ++p2) ntab += pp.P.tab[p2] != nullptr;
            fprintf(stderr, "[plan] %s nfft=%d tables=%d gens=%d (", pp.col ? "col" : "row", pp.P.nfft, ntab, pp.P.ngen);
            for (int g2 = 0; g2 < pp.P.ngen; ++g2) fprintf(stderr, "%d@%d ", pp.P.gen[g2].kind, pp.P.gen[g2].pos);
            fprintf(stderr, ") pos[");
            for (int p2 = 0; p2 <= pp.P.nfft; ++p2) {
                if (pp.P.tab[p2]) fprintf(stderr, "T ");
                else fprintf(stderr, "%s%g ", (pp.P.sgnmask >> p2 & 1) ? "s" : "", pp.P.scl[p2]);
            }
            fprintf(stderr, "] ctab=%d%d src=%d tiles=[%d,%d] in=[%d,%d] out=[%d,%d]\n", pp.P.ctab_in != nullptr, pp.P.ctab_out != nullptr,
                    pp.P.src != nullptr, pp.P.tile_lo, pp.P.tile_hi == 0x7fffffff ? -1 : pp.P.tile_hi, pp.P.in_lo, pp.P.in_hi,
                    pp.P.out_lo, pp.P.out_hi);
        }
    }
    return PAOS_OK;
}

static void account_pass(paos_wfo* w, bool col, const PassParams& P) {
    w->stats.passes_planned++;
    w->stats.line_ffts_run += (uint64_t)P.nfft;
    const int W = tile_width(w->n, w->dtype, col, P.wide != 0), tiles = w->n / W;
    const int active = std::max(0, std::min(P.tile_hi, tiles - 1) - std::max(P.tile_lo, 0) + 1);
    w->stats.lines_transformed += (uint64_t)P.nfft * (uint64_t)active * (uint64_t)W;
    int tabs = (P.ctab_in ? 1 : 0) + (P.ctab_out ? 1 : 0);
    for (int p = 0; p <= P.nfft; ++p) tabs += P.tab[p] != nullptr;
    w->stats.lines_tabled += (uint64_t)tabs * (uint64_t)active * (uint64_t)W;
    w->stats.lines_swept += (uint64_t)active * (uint64_t)W;
}

// one launch of the pass kernel for the same-axis passes Ps[0..nb) of handles that share grid size, precision, device and
// stream; launch counts and timing are booked on `lead`
static int launch_pass_group(paos_wfo* lead, bool col, const PassParams* const* Ps, int nb, cudaStream_t st) {
    cudaEvent_t ea = nullptr, eb = nullptr;
    if (lead->timing) {
        for (cudaEvent_t* e : {&ea, &eb}) {
            if (!lead->event_pool.empty()) {
                *e = lead->event_pool.back();
                lead->event_pool.pop_back();
            } else {
                CU(cudaEventCreate(e));
            }
        }
        CU(cudaEventRecord(ea, st));
    }
    const bool wide = Ps[0]->wide != 0;  // the caller groups passes of one width
    cudaError_t e = (lead->dtype == PAOS_C128) ? launch_pass_c128(lead->n, col, wide, Ps, nb, lead->tw.tw1, lead->tw.tw2, st, lead->device)
                                               : launch_pass_c64(lead->n, col, wide, Ps, nb, lead->tw.tw1, lead->tw.tw2, st, lead->device);
    if (e != cudaSuccess) return fail(PAOS_ERR_CUDA, "pass kernel launch failed: %s", cudaGetErrorString(e));
    if (lead->timing) {
        CU(cudaEventRecord(eb, st));
        int nfft = 0;
        for (int i = 0; i < nb; ++i) nfft += Ps[i]->nfft;
        lead->timed.push_back(TimedLaunch{ea, eb, nfft, col ? 1 : 0, nb});
    }
    lead->stats.kernel_launches++;
    lead->stats.pass_launches++;
    return PAOS_OK;
}

static int launch_edge_tables(paos_wfo* lead, const std::vector<EdgeSpec>& edges, cudaStream_t st) {
    for (size_t s = 0; s < edges.size(); s += EB_MAX) {
        EdgeBlock B{};
        B.n = lead->n;
        B.nspec = (int)std::min<size_t>(EB_MAX, edges.size() - s);
        for (int i = 0; i < B.nspec; ++i) B.spec[i] = edges[s + i];
        cudaError_t e = launch_build_edge_tables(B, st);
        if (e != cudaSuccess) return fail(PAOS_ERR_CUDA, "edge-table builder launch failed: %s", cudaGetErrorString(e));
        lead->stats.kernel_launches++;
    }
    return PAOS_OK;
}

static int launch_tables(paos_wfo* lead, const std::vector<TableSpec>& specs, cudaStream_t st) {
    for (size_t s = 0; s < specs.size(); s += TB_MAX) {
        TableBlock B{};
        B.n = lead->n;
        B.dtype = lead->dtype;
        B.ntab = (int)std::min<size_t>(TB_MAX, specs.size() - s);
        for (int i = 0; i < B.ntab; ++i) B.spec[i] = specs[s + i];
        cudaError_t e = launch_build_tables(B, st);
        if (e != cudaSuccess) return fail(PAOS_ERR_CUDA, "table builder launch failed: %s", cudaGetErrorString(e));
        lead->stats.kernel_launches++;
    }
    return PAOS_OK;
}

static int launch_norm2_group(paos_wfo* lead, const Norm2Item* items, int nb, cudaStream_t st) {
    cudaError_t e = launch_norm2_batch(items, nb, lead->n, lead->dtype, paos_wfo::NPARTIALS, st);
    if (e != cudaSuccess) return fail(PAOS_ERR_CUDA, "stop reduction launch failed: %s", cudaGetErrorString(e));
    lead->stats.kernel_launches += 2;
    return PAOS_OK;
}

// record (deferred mode) or execute right away
static int do_pass(paos_wfo* w, bool col, const PassParams& P) {
    account_pass(w, col, P);
    if (w->recording) {
        w->program.emplace_back();
        Rec& r = w->program.back();
        r.kind = REC_PASS;
        r.col = col;
        r.P = P;
        return PAOS_OK;
    }
    const PassParams* one = &P;
    return launch_pass_group(w, col, &one, 1, w->stream);
}

static int do_tables(paos_wfo* w, std::vector<TableSpec>& specs, const std::vector<EdgeSpec>* edges = nullptr) {
    if (specs.empty() && (!edges || edges->empty())) return PAOS_OK;
    if (w->recording) {
        w->program.emplace_back();
        Rec& r = w->program.back();
        r.kind = REC_TABLES;
        r.specs = specs;
        if (edges) r.edges = *edges;
        return PAOS_OK;
    }
    int rc = launch_tables(w, specs, w->stream);
    if (!rc && edges) rc = launch_edge_tables(w, *edges, w->stream);
    return rc;
}

static int do_norm2(paos_wfo* w, const Norm2Item& item) {
    if (w->recording) {
        w->program.emplace_back();
        Rec& r = w->program.back();
        r.kind = REC_NORM2;
        r.norm = item;
        return PAOS_OK;
    }
    return launch_norm2_group(w, &item, 1, w->stream);
}

// anything else that touches the device: fn(stream) returns a PAOS_* status; launches counts kernels for the statistics
static int do_fn(paos_wfo* w, int launches, std::function<int(cudaStream_t)> fn) {
    w->stats.kernel_launches += (uint64_t)launches;
    if (w->recording) {
        w->program.emplace_back();
        Rec& r = w->program.back();
        r.kind = REC_FN;
        r.fn = std::move(fn);
        return PAOS_OK;
    }
    return fn(w->stream);
}

// Execute the recorded programs of nb handles in lockstep (see Rec).  All handles share n, dtype, device and stream.
static int execute_programs(paos_wfo** ws, int nb) {
    paos_wfo* lead = ws[0];
    cudaStream_t st = lead->stream;
    std::vector<size_t> at((size_t)nb, 0);
    std::vector<TableSpec> specs;
    std::vector<EdgeSpec> edges;
    std::vector<Norm2Item> norms;
    const PassParams* group[BMAX];
    for (;;) {
        int n_fn = 0, n_tab = 0, n_norm = 0, n_row = 0, n_col = 0, live = 0;
        for (int b = 0; b < nb; ++b) {
            if (at[b] >= ws[b]->program.size()) continue;
            ++live;
            const Rec& r = ws[b]->program[at[b]];
            if (r.kind == REC_FN) ++n_fn;
            else if (r.kind == REC_TABLES) ++n_tab;
            else if (r.kind == REC_NORM2) ++n_norm;
            else if (r.col) ++n_col;
            else ++n_row;
        }
        if (!live) break;
        int rc = PAOS_OK;
        if (n_fn) {
            for (int b = 0; b < nb; ++b)
                while (at[b] < ws[b]->program.size() && ws[b]->program[at[b]].kind == REC_FN) {
                    if ((rc = ws[b]->program[at[b]].fn(st))) return rc;
                    ++at[b];
                }
        } else if (n_tab) {
            specs.clear();
            edges.clear();
            for (int b = 0; b < nb; ++b)
                if (at[b] < ws[b]->program.size() && ws[b]->program[at[b]].kind == REC_TABLES) {
                    const Rec& r = ws[b]->program[at[b]++];
                    specs.insert(specs.end(), r.specs.begin(), r.specs.end());
                    edges.insert(edges.end(), r.edges.begin(), r.edges.end());
                }
            if ((rc = launch_tables(lead, specs, st)) || (rc = launch_edge_tables(lead, edges, st))) return rc;
        } else if (n_norm) {
            norms.clear();
            for (int b = 0; b < nb; ++b)
                if (at[b] < ws[b]->program.size() && ws[b]->program[at[b]].kind == REC_NORM2) norms.push_back(ws[b]->program[at[b]++].norm);
            if ((rc = launch_norm2_group(lead, norms.data(), (int)norms.size(), st))) return rc;
        } else {
            const bool col = n_col > n_row;
            int k = 0, wide = -1;  // one launch = one kernel: same axis and same tile width (the first head decides)
            for (int b = 0; b < nb; ++b)
                if (at[b] < ws[b]->program.size() && ws[b]->program[at[b]].kind == REC_PASS && ws[b]->program[at[b]].col == col) {
                    const int wd = ws[b]->program[at[b]].P.wide;
                    if (wide < 0) wide = wd;
                    if (wd != wide) continue;
                    group[k++] = &ws[b]->program[at[b]++].P;
                }
            if ((rc = launch_pass_group(lead, col, group, k, st))) return rc;
        }
    }
    for (int b = 0; b < nb; ++b) ws[b]->program.clear();
    return PAOS_OK;
}

static int ensure_pool(paos_wfo* w, size_t need_tables) {
    const size_t per = (((size_t)w->n * w->elem) + 255) & ~(size_t)255;
    const size_t need = need_tables * per;
    if (need <= w->tab_cap) return PAOS_OK;
    if (w->tab_pool) {
        if (w->recording) {
            w->retired_pools.push_back(w->tab_pool);  // recorded passes still point into it
        } else {
            CU(cudaStreamSynchronize(w->stream));
            CU(cudaFree(w->tab_pool));
        }
        w->tab_pool = nullptr;
    }
    size_t cap = need * 2;
    CU(cudaMalloc((void**)&w->tab_pool, cap));
    w->tab_cap = cap;
    return PAOS_OK;
}

static int run_plan(paos_wfo* w, Plan& plan) {
    // tables first (one or more launches of the builder), then the passes
    int rc = do_tables(w, plan.specs, &plan.edges);
    if (rc) return rc;
    for (PlannedPass& pp : plan.passes) {
        if ((rc = do_pass(w, pp.col, pp.P))) return rc;
    }
    return PAOS_OK;
}

// flush `ops` (a prefix of the queue or all of it)
static int flush_ops(paos_wfo* w, std::vector<Op>& ops, int readout, void* dst_real, bool discard = false) {
    if (ops.empty() && w->materialized && !readout) return PAOS_OK;
    if (w->dropped) return fail(PAOS_ERR_STATE, "the field was discarded by a final read-out: reset the handle before re-using it");
    int rc = set_device(w);
    if (rc) return rc;
    // worst case: every op opens two tables per axis; every elliptical mask adds an edge table
    size_t ellipses = 0;
    for (const Op& op : ops) ellipses += op.kind == OP_GEN && op.gen.kind == GEN_ELLIPSE;
    const size_t per_table = (((size_t)w->n * w->elem) + 255) & ~(size_t)255;
    // (+ room for the edge tables of a stop reduction that may follow this flush, so that it never has to grow the pool)
    rc = ensure_pool(w, 4 * ops.size() + 8 + (ellipses + (size_t)GMAX) * (edge_table_bytes(w->n) / per_table + 2));
    if (rc) return rc;
    w->tab_used = 0;
    Plan plan;
    // rectangle count tables are TableSpecs too: they were registered when the op was recorded and live in
    // screens (see paos_wfo_aperture)
    ZeroBand band = w->band;
    rc = build_plan(w, ops, plan, readout, dst_real, w->field, band);
    if (rc) return rc;
    if (discard && readout && !plan.passes.empty()) {
        plan.passes.back().P.dst = nullptr;  // the last pass writes the read-out only
        w->dropped = true;
    }
    rc = run_plan(w, plan);
    if (rc) return rc;
    ops.clear();
    w->materialized = true;
    w->band = band;
    return PAOS_OK;
}

// write the virtual zeros of `field` (see ZeroBand) before something other than a pass kernel reads it
static int materialize_band(paos_wfo* w, void* field, ZeroBand& band) {
    if (!band.valid) return PAOS_OK;
    const int n = w->n, dtype = w->dtype, axis = band.axis, lo = band.lo, hi = band.hi;
    band.valid = false;
    return do_fn(w, 1, [=](cudaStream_t st) {
        cudaError_t e = launch_zero_outside_band(field, n, dtype, axis, lo, hi, st);
        return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "zero fill launch failed: %s", cudaGetErrorString(e));
    });
}

static int flush_all(paos_wfo* w, int readout = 0, void* dst_real = nullptr, bool discard = false) {
    int rc = flush_ops(w, w->ops, readout, dst_real, discard);
    if (rc) return rc;
    recycle_screens(w);
    return PAOS_OK;
}

static int resolve_timing(paos_wfo* w) {
    for (TimedLaunch& t : w->timed) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, t.a, t.b));
        const int bucket = std::min(KMAX, (t.nfft + t.items / 2) / std::max(t.items, 1));  // average chain length of the launch
        w->timed_ms[t.col][bucket] += ms;
        w->timed_n[t.col][bucket] += 1;
        w->timed_total_ms += ms;
        w->timed_total_launches += 1;
        w->timed_total_sweeps += (uint64_t)t.nfft;
        w->timed_total_items += (uint64_t)t.items;
        w->event_pool.push_back(t.a);
        w->event_pool.push_back(t.b);
    }
    w->timed.clear();
    return PAOS_OK;
}

// RAII for the temporaries of the stateless entry points (paos_zernike_points)
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};


// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int paos_abi_version(void) { return PAOS_ABI_VERSION; }
const char* paos_last_error(void) { return g_last_error.c_str(); }
#ifndef PAOS_SOURCE_HASH
#define PAOS_SOURCE_HASH "unrecorded"
#endif
const char* paos_build_info(void) { return "libpaos_b200 sm_100a source " PAOS_SOURCE_HASH " compiled " __DATE__ " " __TIME__; }
long paos_abi_struct_size(int which) {
    switch (which) {
        case 0: return (long)sizeof(paos_surface);
        case 1: return (long)sizeof(paos_snapshot);
        case 2: return (long)sizeof(paos_stats);
        case 3: return (long)sizeof(paos_chain_args);
        default: return -1;
    }
}

int paos_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

int paos_wfo_create(paos_wfo** out, int n, int dtype, int device, void* stream, void* borrowed) {
    if (!out) return fail(PAOS_ERR_ARG, "out is null");
    *out = nullptr;
    if (n < 64 || n > 4096 || (n & (n - 1))) return fail(PAOS_ERR_ARG, "grid size %d is not a power of two in [64, 4096]", n);
    if (dtype != PAOS_C128 && dtype != PAOS_C64) return fail(PAOS_ERR_ARG, "unknown dtype %d", dtype);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(PAOS_ERR_CUDA, "no CUDA device: libpaos_b200 has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(PAOS_ERR_ARG, "device %d out of range (%d devices)", device, count);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(PAOS_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    CU(cudaSetDevice(device));
    paos_wfo* w = new paos_wfo();
    w->n = n;
    w->dtype = dtype;
    w->device = device;
    w->elem = dtype == PAOS_C128 ? 16 : 8;
    int rc = get_twiddles(device, n, dtype, w->tw);
    if (rc) {
        delete w;
        return rc;
    }
    auto bail = [&](cudaError_t e, const char* what) {
        int r = fail(PAOS_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        paos_wfo_destroy(w);
        return r;
    };
    cudaError_t e;
    if (stream) {
        w->stream = (cudaStream_t)stream;
    } else {
        if ((e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
        w->own_stream = true;
    }
    if (borrowed) {
        w->field = borrowed;
    } else {
        if ((e = cudaMalloc(&w->field, (size_t)n * n * w->elem)) != cudaSuccess) return bail(e, "cudaMalloc(field)");
        w->own_field = true;
    }
    if ((e = cudaMalloc((void**)&w->partials, paos_wfo::NPARTIALS * sizeof(double))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)&w->slots, paos_wfo::NSLOTS * 2 * sizeof(double))) != cudaSuccess) return bail(e, "cudaMalloc");
    *out = w;
    return PAOS_OK;
}

int paos_wfo_destroy(paos_wfo* w) {
    if (!w) return PAOS_OK;
    cudaSetDevice(w->device);
    if (w->stream) cudaStreamSynchronize(w->stream);
    for (TimedLaunch& t : w->timed) {
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    for (cudaEvent_t e : w->event_pool) cudaEventDestroy(e);
    for (double* p : w->screens_free) cudaFree(p);
    for (double* p : w->screens_busy) cudaFree(p);
    if (w->scratch_field) cudaFree(w->scratch_field);
    if (w->tab_pool) cudaFree(w->tab_pool);
    for (void* p : w->retired_pools) cudaFree(p);
    for (auto& kv : w->tmaps) delete kv.second;
    for (CUtensorMap* tm : w->retired_tmaps) delete tm;
    if (w->partials) cudaFree(w->partials);
    if (w->slots) cudaFree(w->slots);
    if (w->own_field && w->field) cudaFree(w->field);
    if (w->ee_hist) cudaFree(w->ee_hist);
    if (w->own_stream && w->stream) cudaStreamDestroy(w->stream);
    cudaGetLastError();
    delete w;
    return PAOS_OK;
}

int paos_wfo_reset(paos_wfo* w) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    w->ops.clear();
    recycle_screens(w);
    w->materialized = false;
    w->band.valid = false;
    w->dropped = false;
    return PAOS_OK;
}

int paos_wfo_fill_ones(paos_wfo* w) { return paos_wfo_reset(w); }

int paos_wfo_flush(paos_wfo* w) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    return flush_all(w);
}

int paos_wfo_materialize(paos_wfo* w) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    int rc = flush_all(w);
    if (rc) return rc;
    return materialize_band(w, w->field, w->band);
}

int paos_wfo_sync(paos_wfo* w) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    if (w->recording) return fail(PAOS_ERR_STATE, "the handle is recording: run paos_batch_execute first");
    int rc = flush_all(w);
    if (rc) return rc;
    CU(cudaStreamSynchronize(w->stream));
    for (void* p : w->retired_pools) cudaFree(p);
    w->retired_pools.clear();
    for (CUtensorMap* tm : w->retired_tmaps) delete tm;
    w->retired_tmaps.clear();
    return resolve_timing(w);
}

// ---- deferred execution of several handles as one batch -----------------------------------------------
int paos_wfo_begin_record(paos_wfo* w) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    if (w->recording) return fail(PAOS_ERR_STATE, "the handle is recording already");
    w->recording = true;
    w->program.clear();
    return PAOS_OK;
}

static void abandon_records(paos_wfo* const* ws, int nb) {
    for (int b = 0; b < nb; ++b)
        if (ws[b]) {
            ws[b]->recording = false;
            ws[b]->program.clear();
            ws[b]->ops.clear();
        }
}

int paos_batch_execute(paos_wfo* const* ws, int nb) {
    if (!ws || nb < 1) return fail(PAOS_ERR_ARG, "no handles");
    if (nb > BMAX) return fail(PAOS_ERR_ARG, "a batch holds at most %d wavefronts", BMAX);
    for (int b = 0; b < nb; ++b) {
        if (!ws[b]) return fail(PAOS_ERR_ARG, "null handle in the batch");
        if (ws[b]->n != ws[0]->n || ws[b]->dtype != ws[0]->dtype || ws[b]->device != ws[0]->device || ws[b]->stream != ws[0]->stream) {
            abandon_records(ws, nb);
            return fail(PAOS_ERR_ARG, "the handles of a batch must share grid size, precision, device and stream");
        }
        for (int c = 0; c < b; ++c)
            if (ws[c] == ws[b]) {
                abandon_records(ws, nb);
                return fail(PAOS_ERR_ARG, "a handle appears twice in the batch");
            }
    }
    int rc = set_device(ws[0]);
    std::vector<paos_wfo*> hs(ws, ws + nb);
    for (int b = 0; b < nb && !rc; ++b) {
        // what is still queued behind the last flush (nothing after a final read-out) is planned now, still deferred
        if (hs[b]->recording && !hs[b]->dropped) rc = flush_all(hs[b]);
    }
    for (int b = 0; b < nb; ++b) hs[b]->recording = false;
    if (!rc) rc = execute_programs(hs.data(), nb);
    if (rc) abandon_records(ws, nb);
    return rc;
}

int paos_batch_chain_run(paos_wfo* const* ws, int nb, const paos_chain_args* args) {
    if (!ws || !args || nb < 1) return fail(PAOS_ERR_ARG, "null argument");
    if (nb > BMAX) return fail(PAOS_ERR_ARG, "a batch holds at most %d wavefronts", BMAX);
    const auto t0 = std::chrono::steady_clock::now();
    struct Book {  // host time of planning + launching, booked on the first handle
        paos_wfo* w;
        std::chrono::steady_clock::time_point t0;
        ~Book() {
            if (w) w->stats.host_plan_us += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        }
    } book{ws[0], t0};
    // The chains of a batch are independent until they are executed, so their surfaces are walked and their passes planned
    // by a few host threads (PAOS_PLAN_THREADS, default 4 for batches of 8 and more): at 512^2 the device needs ~25 us
    // per PSF and one thread's planning would be the bottleneck.
    static const int max_threads = [] {
        const char* e = getenv("PAOS_PLAN_THREADS");
        const int v = e ? atoi(e) : 4;
        return v < 1 ? 1 : (v > 16 ? 16 : v);
    }();
    for (int b = 0; b < nb; ++b)
        if (!ws[b]) return fail(PAOS_ERR_ARG, "null handle in the batch");
    std::vector<int> rcs((size_t)nb, PAOS_OK);
    std::vector<std::string> msgs((size_t)nb);
    auto record = [&](int b) {
        int rc = paos_wfo_begin_record(ws[b]);
        if (!rc) {
            const paos_chain_args& a = args[b];
            rc = paos_chain_run(ws[b], a.pupil_diameter, a.wavelength, a.zoom, a.us, a.ut, a.surfaces, a.n_surfaces, a.snapshots,
                                a.max_snapshots, a.n_snapshots, a.final_state);
        }
        rcs[(size_t)b] = rc;
        if (rc) msgs[(size_t)b] = g_last_error;  // thread-local: carry it to the calling thread
    };
    const int nthreads = nb >= 8 ? std::min(max_threads, nb) : 1;
    if (nthreads <= 1) {
        for (int b = 0; b < nb; ++b) record(b);
    } else {
        std::vector<std::thread> pool;
        for (int t = 1; t < nthreads; ++t)
            pool.emplace_back([&, t] {
                for (int b = t; b < nb; b += nthreads) record(b);
            });
        for (int b = 0; b < nb; b += nthreads) record(b);
        for (std::thread& th : pool) th.join();
    }
    for (int b = 0; b < nb; ++b)
        if (rcs[(size_t)b]) {
            abandon_records(ws, nb);
            g_last_error = msgs[(size_t)b];
            return rcs[(size_t)b];
        }
    return paos_batch_execute(ws, nb);
}

int paos_batch_capacity(void) { return BMAX; }

int paos_wfo_upload(paos_wfo* w, const void* host_src) {
    if (!w || !host_src) return fail(PAOS_ERR_ARG, "null argument");
    if (w->recording) return fail(PAOS_ERR_STATE, "the handle is recording");
    int rc = set_device(w);
    if (rc) return rc;
    w->ops.clear();
    recycle_screens(w);
    CU(cudaMemcpyAsync(w->field, host_src, (size_t)w->n * w->n * w->elem, cudaMemcpyHostToDevice, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    w->materialized = true;
    w->band.valid = false;
    w->dropped = false;
    return PAOS_OK;
}

int paos_wfo_upload_device(paos_wfo* w, const void* dev_src) {
    if (!w || !dev_src) return fail(PAOS_ERR_ARG, "null argument");
    int rc = set_device(w);
    if (rc) return rc;
    w->ops.clear();
    recycle_screens(w);
    if (dev_src != w->field)
        CU(cudaMemcpyAsync(w->field, dev_src, (size_t)w->n * w->n * w->elem, cudaMemcpyDeviceToDevice, w->stream));
    w->materialized = true;
    w->band.valid = false;
    w->dropped = false;
    return PAOS_OK;
}

static int read_impl(paos_wfo* w, int what, void* dst, bool to_host, bool discard = false) {
    if (!w || !dst) return fail(PAOS_ERR_ARG, "null argument");
    if (w->recording && to_host) return fail(PAOS_ERR_STATE, "a recording handle cannot be read to the host: execute the batch first");
    if (what < PAOS_READ_WFO || what > PAOS_READ_PSF) return fail(PAOS_ERR_ARG, "unknown read-out %d", what);
    if (w->dropped) return fail(PAOS_ERR_STATE, "the field was discarded by a final read-out: reset the handle before re-using it");
    int rc = set_device(w);
    if (rc) return rc;
    const size_t nn = (size_t)w->n * w->n;
    if (what == PAOS_READ_WFO) {
        rc = flush_all(w);
        if (rc) return rc;
        rc = materialize_band(w, w->field, w->band);
        if (rc) return rc;
        {
            const void* src = w->field;
            const size_t bytes = nn * w->elem;
            rc = do_fn(w, 0, [=](cudaStream_t st) {
                cudaError_t e = cudaMemcpyAsync(dst, src, bytes, to_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st);
                return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "field copy failed: %s", cudaGetErrorString(e));
            });
            if (rc) return rc;
        }
    } else {
        const size_t rbytes = nn * (w->elem / 2);
        void* dev_out = dst;
        double* staging = nullptr;
        if (to_host) {
            rc = get_screen(w, &staging);
            if (rc) return rc;
            dev_out = staging;
        }
        if (!w->ops.empty() || !w->materialized) {
            rc = flush_all(w, what, dev_out, discard && !to_host);  // read-out fused into the last pass
            if (rc) return rc;
        } else {
            rc = materialize_band(w, w->field, w->band);
            if (rc) return rc;
            const void* src = w->field;
            const int n = w->n, dtype = w->dtype;
            rc = do_fn(w, 1, [=](cudaStream_t st) {
                cudaError_t e = launch_readout(src, n, dtype, what, dev_out, st);
                return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "read-out launch failed: %s", cudaGetErrorString(e));
            });
            if (rc) return rc;
        }
        if (to_host) {
            CU(cudaMemcpyAsync(dst, dev_out, rbytes, cudaMemcpyDeviceToHost, w->stream));
            // flush_all recycled the staging buffer already when it ran; make sure it is not busy-listed twice
            recycle_screens(w);
        }
    }
    if (to_host) {
        CU(cudaStreamSynchronize(w->stream));
        return resolve_timing(w);
    }
    return PAOS_OK;
}

int paos_wfo_read(paos_wfo* w, int what, void* host_dst) { return read_impl(w, what, host_dst, true); }
int paos_wfo_read_device(paos_wfo* w, int what, void* dev_dst) { return read_impl(w, what, dev_dst, false); }
int paos_wfo_read_device_final(paos_wfo* w, int what, void* dev_dst) {
    if (what == PAOS_READ_WFO) return fail(PAOS_ERR_ARG, "a final read-out is |.|, angle or |.|^2");
    return read_impl(w, what, dev_dst, false, true);
}

// ---- elementwise operators -----------------------------------------------------------------------
static void push_gen(paos_wfo* w, const GenOp& g) {
    Op op{};
    op.kind = OP_GEN;
    op.gen = g;
    w->ops.push_back(op);
}

int paos_wfo_aperture(paos_wfo* w, int shape, double ixc, double iyc, double ihx, double ihy, double theta, int obscuration) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    if (!(ihx > 0.0) || !(ihy > 0.0) || !std::isfinite(ixc) || !std::isfinite(iyc)) return fail(PAOS_ERR_ARG, "bad aperture geometry");
    GenOp g{};
    g.flag = obscuration ? 1 : 0;
    if (theta != 0.0 && (shape == PAOS_SHAPE_ELLIPSE || shape == PAOS_SHAPE_RECT)) {
        // tilted shapes (never produced by paos.core.run): general factors evaluated per pixel, quick interior / exterior
        // decision first, exact overlap (ellipse) or 32x32 sub-pixel count (rectangle) at the edge
        if (!std::isfinite(theta)) return fail(PAOS_ERR_ARG, "bad aperture tilt");
        g.p0 = ixc;
        g.p1 = iyc;
        g.p7 = std::cos(theta);
        g.p8 = std::sin(theta);
        if (shape == PAOS_SHAPE_ELLIPSE) {
            g.kind = GEN_ELLIPSE_TILT;
            g.p2 = 1.0 / ihx;
            g.p3 = 1.0 / ihy;
            g.p4 = ihx * ihy;
            const double d = std::sqrt(0.5) * std::max(g.p2, g.p3) * (1.0 + 1e-9) + 1e-12;  // pixel half-diagonal, worst axis
            g.p5 = d < 1.0 ? (1.0 - d) * (1.0 - d) : -1.0;
            g.p6 = (1.0 + d) * (1.0 + d);
        } else {
            g.kind = GEN_RECT_TILT;
            g.p2 = ihx / 2.0;
            g.p3 = ihy / 2.0;
        }
        push_gen(w, g);
        return PAOS_OK;
    }
    if (shape == PAOS_SHAPE_ELLIPSE) {
        g.kind = GEN_ELLIPSE;
        g.p0 = ixc;
        g.p1 = iyc;
        g.p2 = 1.0 / ihx;
        g.p3 = 1.0 / ihy;
        g.p4 = ihx * ihy;
        // a unit pixel lies within d = half its diagonal (in the frame where the ellipse is the unit circle) of its
        // centre: certainly inside when r <= 1 - d, certainly outside when r >= 1 + d (plus a rounding margin)
        const double d = 0.5 * std::sqrt(g.p2 * g.p2 + g.p3 * g.p3) * (1.0 + 1e-9) + 1e-12;
        g.p5 = d < 1.0 ? (1.0 - d) * (1.0 - d) : -1.0;
        g.p6 = (1.0 + d) * (1.0 + d);
    } else if (shape == PAOS_SHAPE_RECT && !obscuration) {
        // a rectangular *aperture* is separable, mask = (cx[ix]/32) * (cy[iy]/32): it rides in the phase tables like a
        // chirp and costs no pass boundary (an obscuration, 1 - mask, is not separable: general factor below)
        Op op{};
        op.kind = OP_PHASE;
        op.tx.kind = op.ty.kind = TERM_COUNT;
        op.tx.c1 = ixc;
        op.tx.c2 = ihx;
        op.ty.c1 = iyc;
        op.ty.c2 = ihy;
        w->ops.push_back(op);
        return PAOS_OK;
    } else if (shape == PAOS_SHAPE_RECT) {
        // separable 32-sub-pixel counts, built right away into a screen buffer (2*n doubles)
        int rc = set_device(w);
        if (rc) return rc;
        double* buf;
        rc = get_screen(w, &buf);
        if (rc) return rc;
        std::vector<TableSpec> specs(2);
        specs[0].out = buf;
        specs[0].kind = TABLE_COUNT;
        specs[0].cnt_c = ixc;
        specs[0].cnt_full = ihx;
        specs[1].out = buf + w->n;
        specs[1].kind = TABLE_COUNT;
        specs[1].cnt_c = iyc;
        specs[1].cnt_full = ihy;
        if ((rc = do_tables(w, specs))) return rc;
        g.kind = GEN_RECT;
        g.ptr0 = buf;
        g.ptr1 = buf + w->n;
    } else {
        return fail(PAOS_ERR_ARG, "unknown aperture shape %d", shape);
    }
    push_gen(w, g);
    return PAOS_OK;
}

int paos_wfo_make_stop(paos_wfo* w) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    int rc = set_device(w);
    if (rc) return rc;
    // trailing pointwise operations: unit-modulus ones (signs, phases, screens) do not change the energy, real
    // masks and earlier stop scalars are folded into the reduction; everything before them is flushed
    size_t k = w->ops.size();
    std::vector<GenOp> gens;
    while (k > 0) {
        const Op& op = w->ops[k - 1];
        if (op.kind == OP_SIGN || (op.kind == OP_PHASE && op.tx.kind != TERM_COUNT)) {
            --k;
        } else if (op.kind == OP_GEN && op.gen.kind == GEN_SCREEN) {
            --k;
        } else if (op.kind == OP_GEN && op.gen.kind != GEN_PSD && gens.size() < (size_t)GMAX) {
            gens.push_back(op.gen);
            --k;
        } else {
            break;
        }
    }
    std::vector<Op> tail(w->ops.begin() + k, w->ops.end());
    w->ops.resize(k);
    if (k > 0) {
        rc = flush_ops(w, w->ops, 0, nullptr);
        if (rc) return rc;
    }
    w->ops = tail;
    if (w->materialized && (rc = materialize_band(w, w->field, w->band))) return rc;
    // The reduction multiplies every pixel by the folded masks; the rim pixels of the elliptical ones take their exact overlap
    // from row-axis edge tables built right before it (one lane per rim pixel) instead of one serial evaluation per mask
    // and rim pixel inside the reduction (which made it 124 us per batch of 8 wavefronts).  The tables sit behind those of
    // the flush above in the pool; the next flush reuses the memory in stream order.
    static const bool edge_tables = getenv("PAOS_NO_EDGE_TABLES") == nullptr;
    if (edge_tables) {
        std::vector<EdgeSpec> edges;
        const size_t per_table = (((size_t)w->n * w->elem) + 255) & ~(size_t)255;
        size_t ellipses = 0;
        for (const GenOp& g : gens) ellipses += g.kind == GEN_ELLIPSE;
        if (ellipses && ensure_pool(w, w->tab_used / per_table + 1 + ellipses * (edge_table_bytes(w->n) / per_table + 2)) == PAOS_OK) {
            for (GenOp& g : gens) {
                if (g.kind != GEN_ELLIPSE) continue;
                void* mem;
                if (alloc_table(w, edge_table_bytes(w->n), &mem) != PAOS_OK) break;
                edges.push_back(make_edge_spec(w, g, mem, 0));
            }
            std::vector<TableSpec> none;
            if ((rc = do_tables(w, none, &edges))) return rc;
        }
    }
    double* slot = w->slots + 2 * (w->slot_next++ % paos_wfo::NSLOTS);
    Norm2Item item{};
    item.src = w->materialized ? w->field : nullptr;
    item.ngen = (int)gens.size();
    for (int i = 0; i < item.ngen; ++i) item.gen[i] = gens[i];
    item.partials = w->partials;
    item.out_slot = slot;
    if ((rc = do_norm2(w, item))) return rc;
    GenOp g{};
    g.kind = GEN_SCALE_DEV;
    g.ptr0 = slot;
    push_gen(w, g);
    return PAOS_OK;
}

int paos_wfo_quadphase(paos_wfo* w, double c1, double c2, double dx, double dy) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    Op op{};
    op.kind = OP_PHASE;
    op.tx.kind = op.ty.kind = TERM_QSPACE;
    op.tx.c1 = op.ty.c1 = c1;
    op.tx.c2 = op.ty.c2 = c2;
    op.tx.d = dx;
    op.ty.d = dy;
    w->ops.push_back(op);
    return PAOS_OK;
}

int paos_wfo_phase_screen_device(paos_wfo* w, const double* dev_screen, double wl) {
    if (!w || !dev_screen) return fail(PAOS_ERR_ARG, "null argument");
    GenOp g{};
    g.kind = GEN_SCREEN;
    g.ptr0 = dev_screen;
    g.p0 = wl;
    push_gen(w, g);
    return PAOS_OK;
}

int paos_wfo_phase_screen(paos_wfo* w, const double* host_screen, double wl) {
    if (!w || !host_screen) return fail(PAOS_ERR_ARG, "null argument");
    int rc = set_device(w);
    if (rc) return rc;
    double* buf;
    rc = get_screen(w, &buf);
    if (rc) return rc;
    const size_t bytes = (size_t)w->n * w->n * sizeof(double);
    rc = do_fn(w, 0, [=](cudaStream_t st) {
        cudaError_t e = cudaMemcpyAsync(buf, host_screen, bytes, cudaMemcpyHostToDevice, st);
        return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "screen upload failed: %s", cudaGetErrorString(e));
    });
    if (rc) return rc;
    return paos_wfo_phase_screen_device(w, buf, wl);
}

static double binom(int n, int k) {
    double b = 1.0;
    for (int i = 1; i <= k; ++i) b = b * (double)(n - k + i) / (double)i;
    return b;
}

// fill the kernel parameters of one chunk of <= ZERN_MAX terms; coef_in may be NULL (coefficient 1)
static int fill_zern(int grid_n, ZernParams& Z, int s, int nterms, const int* m, const int* n, const double* coef_in,
                     const double* norm_in, double radius, double dx, double dy, double offset, int origin) {
    Z = ZernParams{};
    Z.K = std::min(ZERN_MAX, nterms - s);
    Z.origin = origin;
    Z.n = grid_n;
    Z.accumulate = s > 0;
    Z.radius = radius;
    Z.dx = dx;
    Z.dy = dy;
    Z.cos_off = std::cos(offset);
    Z.sin_off = std::sin(offset);
    for (int k = 0; k < Z.K; ++k) {
        const int mm = m[s + k], nn = n[s + k], am = mm < 0 ? -mm : mm;
        if (nn < am || ((nn - am) & 1)) return fail(PAOS_ERR_ARG, "invalid Zernike (m, n) = (%d, %d)", mm, nn);
        const int kr = (nn - am) / 2;
        Z.m[k] = mm;
        Z.nn[k] = nn;
        // scipy's eval_jacobi returns binom(k+alpha, k) * recurrence; (-1)^k from zernike.py:245-247
        Z.coef[k] = (coef_in ? coef_in[s + k] : 1.0) * (norm_in ? norm_in[s + k] : 1.0) * binom(kr + am, kr) * ((kr & 1) ? -1.0 : 1.0);
    }
    return PAOS_OK;
}

// pixel mask (numpy masked-array convention: non-zero = masked) to the device; NULL stays NULL
static int upload_mask(paos_wfo* w, const unsigned char* host_mask, unsigned char** dev) {
    *dev = nullptr;
    if (!host_mask) return PAOS_OK;
    double* buf;
    int rc = get_screen(w, &buf);  // n*n doubles are more than enough for n*n bytes
    if (rc) return rc;
    const size_t bytes = (size_t)w->n * w->n;
    rc = do_fn(w, 0, [=](cudaStream_t st) {
        cudaError_t e = cudaMemcpyAsync(buf, host_mask, bytes, cudaMemcpyHostToDevice, st);
        return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "mask upload failed: %s", cudaGetErrorString(e));
    });
    if (rc) return rc;
    *dev = reinterpret_cast<unsigned char*>(buf);
    return PAOS_OK;
}

int paos_wfo_zernike_masked(paos_wfo* w, int nterms, const int* m, const int* n, const double* coef, double radius, double dx,
                            double dy, double offset, int origin, double wl, const unsigned char* host_mask, double* wfe_host_out) {
    if (!w || !m || !n || !coef) return fail(PAOS_ERR_ARG, "null argument");
    if (nterms < 1) return fail(PAOS_ERR_ARG, "need at least one Zernike term");
    if (origin != 0 && origin != 1) return fail(PAOS_ERR_ARG, "origin must be 0 ('x') or 1 ('y')");
    if (!(radius > 0.0)) return fail(PAOS_ERR_ARG, "radius must be positive");
    if (w->recording && wfe_host_out) return fail(PAOS_ERR_STATE, "a recording handle cannot return the screen to the host");
    int rc = set_device(w);
    if (rc) return rc;
    double* screen;
    rc = get_screen(w, &screen);
    if (rc) return rc;
    unsigned char* dmask;
    if ((rc = upload_mask(w, host_mask, &dmask))) return rc;
    for (int s = 0; s < nterms; s += ZERN_MAX) {
        ZernParams Z;
        if ((rc = fill_zern(w->n, Z, s, nterms, m, n, coef, nullptr, radius, dx, dy, offset, origin))) return rc;
        rc = do_fn(w, 1, [=](cudaStream_t st) {
            cudaError_t e = launch_zernike(Z, dmask, screen, st);
            return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "zernike launch failed: %s", cudaGetErrorString(e));
        });
        if (rc) return rc;
    }
    if (wfe_host_out) {
        CU(cudaMemcpyAsync(wfe_host_out, screen, (size_t)w->n * w->n * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
        CU(cudaStreamSynchronize(w->stream));
    }
    return paos_wfo_phase_screen_device(w, screen, wl);
}

int paos_wfo_zernike(paos_wfo* w, int nterms, const int* m, const int* n, const double* coef, double radius, double dx,
                     double dy, double offset, int origin, double wl, double* wfe_host_out) {
    return paos_wfo_zernike_masked(w, nterms, m, n, coef, radius, dx, dy, offset, origin, wl, nullptr, wfe_host_out);
}

int paos_zernike_cov(paos_wfo* w, int nterms, const int* m, const int* n, const double* norm, double radius, double dx, double dy,
                     double offset, int origin, const unsigned char* host_mask, double* cov_host_out) {
    if (!w || !m || !n || !cov_host_out) return fail(PAOS_ERR_ARG, "null argument");
    if (nterms < 1 || nterms > ZERN_MAX) return fail(PAOS_ERR_ARG, "covariance supports 1..%d polynomials", ZERN_MAX);
    if (origin != 0 && origin != 1) return fail(PAOS_ERR_ARG, "origin must be 0 ('x') or 1 ('y')");
    if (!(radius > 0.0)) return fail(PAOS_ERR_ARG, "radius must be positive");
    if (w->recording) return fail(PAOS_ERR_STATE, "paos_zernike_cov is synchronous: not available on a recording handle");
    int rc = set_device(w);
    if (rc) return rc;
    ZernParams Z;
    if ((rc = fill_zern(w->n, Z, 0, nterms, m, n, nullptr, norm, radius, dx, dy, offset, origin))) return rc;
    unsigned char* dmask;
    if ((rc = upload_mask(w, host_mask, &dmask))) return rc;
    const int K = nterms, per = K * K + 1;
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>(148 * 2, (size_t)w->n * w->n / per));
    double* partial;
    if ((rc = get_screen(w, &partial))) return rc;  // an n*n double buffer holds blocks * (K*K+1) partial sums
    if ((size_t)blocks * per > (size_t)w->n * w->n) return fail(PAOS_ERR_UNSUPPORTED, "grid too small for %d covariance terms", K);
    cudaError_t e = launch_zernike_cov(Z, dmask, partial, blocks, w->stream);
    if (e != cudaSuccess) return fail(PAOS_ERR_CUDA, "covariance launch failed: %s", cudaGetErrorString(e));
    w->stats.kernel_launches++;
    std::vector<double> host((size_t)blocks * per);
    CU(cudaMemcpyAsync(host.data(), partial, host.size() * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    std::vector<long double> sum(per, 0.0L);
    for (int bk = 0; bk < blocks; ++bk)
        for (int q = 0; q < per; ++q) sum[q] += host[(size_t)bk * per + q];
    const long double count = sum[K * K];
    if (!(count > 0)) return fail(PAOS_ERR_STATE, "every pixel is masked: the covariance is undefined");
    for (int q = 0; q < K * K; ++q) cov_host_out[q] = (double)(sum[q] / count);  // np.ma.mean over the unmasked pixels
    return PAOS_OK;
}

// run a private op list on the scratch field (PSD synthesis) without touching the handle's queue
static int run_on_scratch(paos_wfo* w, std::vector<Op>& ops) {
    int rc = ensure_pool(w, 4 * (ops.size() + w->ops.size()) + 16);
    if (rc) return rc;
    // the handle's own queue has not been planned yet, so the pool can be used from the start; the tables of
    // this private plan are consumed before any later flush rewrites them (stream order)
    w->tab_used = 0;
    Plan plan;
    const bool mat = w->materialized;
    w->materialized = true;  // the scratch field has data
    ZeroBand band;  // the scratch field arrives fully written
    rc = build_plan(w, ops, plan, 0, nullptr, w->scratch_field, band);
    w->materialized = mat;
    if (rc) return rc;
    rc = run_plan(w, plan);
    if (rc) return rc;
    return materialize_band(w, w->scratch_field, band);  // its readers are plain kernels
}

int paos_wfo_psd(paos_wfo* w, double A, double B, double C, double fknee, double fmin, double fmax, double SR,
                 double unit_scale, double dx, double dy, double wl, const double* noise1, const double* noise2,
                 uint64_t seed, double* wfe_host_out) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    if ((noise1 == nullptr) != (noise2 == nullptr)) return fail(PAOS_ERR_ARG, "pass both noise arrays or neither");
    int rc = set_device(w);
    if (rc) return rc;
    if (w->recording && wfe_host_out) return fail(PAOS_ERR_STATE, "a recording handle cannot return the screen to the host");
    const size_t nn = (size_t)w->n * w->n;
    const int n = w->n, dtype = w->dtype;
    double *d1, *d2, *screen;
    if ((rc = get_screen(w, &d1)) || (rc = get_screen(w, &d2)) || (rc = get_screen(w, &screen))) return rc;
    if (!w->scratch_field) CU(cudaMalloc(&w->scratch_field, nn * w->elem));
    void* scratch = w->scratch_field;
    rc = do_fn(w, noise1 ? 1 : 3, [=](cudaStream_t st) {
        cudaError_t e;
        if (noise1) {
            e = cudaMemcpyAsync(d1, noise1, nn * sizeof(double), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d2, noise2, nn * sizeof(double), cudaMemcpyHostToDevice, st);
        } else {
            e = launch_normal(seed, 1u, n, d1, st);
            if (e == cudaSuccess) e = launch_normal(seed, 2u, n, d2, st);
        }
        if (e == cudaSuccess) e = launch_real_to_complex(d1, n, dtype, scratch, st);
        return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "PSD noise stage failed: %s", cudaGetErrorString(e));
    });
    if (rc) return rc;
    // F = fft2(noise) * filter ; wfe = real(ifft2(F)) (numpy default normalisation: 1/(N*N) on the inverse)
    std::vector<Op> ops;
    Op f{};
    f.kind = OP_FFT_RAW;
    f.dir = +1;
    f.raw_scale = 1.0;
    ops.push_back(f);
    Op g{};
    g.kind = OP_GEN;
    g.gen.kind = GEN_PSD;
    g.gen.p0 = A;
    g.gen.p1 = B;
    g.gen.p2 = C;
    g.gen.p3 = fknee;
    g.gen.p4 = fmin;
    g.gen.p5 = fmax;
    g.gen.p6 = 1.0 / ((double)w->n * dx);
    g.gen.p7 = 1.0 / ((double)w->n * dy);
    g.gen.p8 = (double)w->n;
    ops.push_back(g);
    f.dir = -1;
    ops.push_back(f);
    rc = run_on_scratch(w, ops);
    if (rc) return rc;
    rc = do_fn(w, 1, [=](cudaStream_t st) {
        cudaError_t e = launch_psd_finalize(scratch, n, dtype, d2, SR, unit_scale, screen, st);
        return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "launch failed: %s", cudaGetErrorString(e));
    });
    if (rc) return rc;
    w->stats.fft2_recorded += 2;
    if (wfe_host_out) {
        CU(cudaMemcpyAsync(wfe_host_out, screen, nn * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
        CU(cudaStreamSynchronize(w->stream));
    }
    return paos_wfo_phase_screen_device(w, screen, wl);
}

// ---- Grid Sag: device preparation of the map (wfo.py:696-862) ---------------------------------------------------
namespace {

// Impulse response of scipy.ndimage.fourier_shift along one axis of length n for a complex transform (the reference
// multiplies fft2(sag) by it and takes ifft2(...).real, wfo.py:768-773): multiplier exp(-2*pi*i*shift*f/n) with f = k for
// 2k < n, else k - n, hence h[m] = (1/n) * sum_f exp(2*pi*i*f*(m - shift)/n), a Dirichlet kernel in closed form.
void fourier_shift_kernel(int n, double shift, std::vector<double>& re, std::vector<double>& im) {
    re.assign(n, 0.0);
    im.assign(n, 0.0);
    const long double PI = 3.14159265358979323846264338327950288L;
    const bool even = (n % 2) == 0;
    for (int m = 0; m < n; ++m) {
        const long double d = (long double)m - (long double)shift;  // u/2 = pi*d/n
        const long double half = PI * d / (long double)n;
        const long double den = sinl(half);
        long double mag;
        if (fabsl(den) < 1e-18L) {
            mag = (long double)n * cosl(PI * d) / cosl(half);  // limit of sin(n x)/sin(x) at a multiple of pi
        } else {
            mag = sinl(PI * d) / den;
        }
        // odd n: f runs over -(n-1)/2 .. (n-1)/2 (real kernel); even n: -n/2 .. n/2 - 1, phase factor exp(-i*u/2)
        long double cr = 1.0L, ci = 0.0L;
        if (even) {
            cr = cosl(half);
            ci = -sinl(half);
        }
        re[m] = (double)(mag * cr / (long double)n);
        im[m] = (double)(mag * ci / (long double)n);
    }
}

struct SagBuf {  // a rows x cols device array
    double* p = nullptr;
    int rows = 0, cols = 0;
    size_t count() const { return (size_t)rows * cols; }
};

struct SagWork {  // temporaries of one preparation, freed on exit
    std::vector<void*> owned;
    ~SagWork() {
        for (void* p : owned) cudaFree(p);
    }
    cudaError_t alloc(SagBuf& b, int rows, int cols) {
        b.rows = rows;
        b.cols = cols;
        cudaError_t e = cudaMalloc((void**)&b.p, std::max<size_t>(1, b.count()) * sizeof(double));
        if (e == cudaSuccess) owned.push_back(b.p);
        return e;
    }
    cudaError_t alloc_raw(void** p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, std::max<size_t>(1, bytes));
        if (e == cudaSuccess) owned.push_back(*p);
        return e;
    }
};

#define SAG(expr)                                                                                     \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(PAOS_ERR_CUDA, "grid sag: %s failed: %s", #expr, cudaGetErrorString(e__));    \
    } while (0)

// skimage.transform.resize(order=3) of one map: Gaussian pre-filter per axis when anti-aliasing, B-spline pre-filter and
// evaluation per axis, clip to the input range (paos_b200/resample.py is the host statement of the same operators)
int sag_resize(SagWork& wk, SagBuf& a, int out_rows, int out_cols, bool anti_aliasing, cudaStream_t st) {
    double* lohi;
    SAG(wk.alloc_raw((void**)&lohi, 2 * sizeof(double)));
    SAG(sag_minmax(a.p, a.count(), lohi, st));
    const int out_shape[2] = {out_rows, out_cols};
    for (int axis = 0; axis < 2; ++axis) {
        const int n_in = axis == 0 ? a.rows : a.cols, n_out = out_shape[axis];
        const double sigma = anti_aliasing ? std::max(0.0, ((double)n_in / (double)n_out - 1.0) / 2.0) : 0.0;
        if (sigma > 1e-15) {
            const int radius = (int)(4.0 * sigma + 0.5);
            std::vector<double> w(2 * radius + 1);
            double sum = 0.0;
            for (int k = -radius; k <= radius; ++k) sum += (w[k + radius] = std::exp(-0.5 / (sigma * sigma) * (double)k * (double)k));
            for (double& v : w) v /= sum;
            double* wd;
            SAG(wk.alloc_raw((void**)&wd, w.size() * sizeof(double)));
            SAG(cudaMemcpyAsync(wd, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice, st));
            SAG(cudaStreamSynchronize(st));  // w is a stack object
            SagBuf out;
            SAG(wk.alloc(out, a.rows, a.cols));
            SAG(sag_fir_mirror(a.p, a.rows, a.cols, axis, wd, radius, out.p, st));
            a = out;
        }
    }
    for (int axis = 0; axis < 2; ++axis) {
        SagBuf cf;
        SAG(wk.alloc(cf, a.rows, a.cols));
        SAG(cudaMemcpyAsync(cf.p, a.p, a.count() * sizeof(double), cudaMemcpyDeviceToDevice, st));
        SAG(sag_bspline_prefilter(cf.p, cf.rows, cf.cols, axis, st));
        SagBuf out;
        SAG(wk.alloc(out, axis == 0 ? out_shape[0] : a.rows, axis == 1 ? out_shape[1] : a.cols));
        SAG(sag_bspline_interp(cf.p, cf.rows, cf.cols, axis, out_shape[axis], out.p, st));
        a = out;
    }
    SAG(sag_clip(a.p, a.count(), lohi, st));
    return PAOS_OK;
}

int sag_rescale(SagWork& wk, SagBuf& a, double scale_y, double scale_x, bool anti_aliasing, cudaStream_t st) {
    const int out_rows = (int)std::max(std::nearbyint(scale_y * (double)a.rows), 1.0);
    const int out_cols = (int)std::max(std::nearbyint(scale_x * (double)a.cols), 1.0);
    return sag_resize(wk, a, out_rows, out_cols, anti_aliasing, st);
}

int sag_shift(SagWork& wk, SagBuf& a, double shift0, double shift1, cudaStream_t st) {
    std::vector<double> re0, im0, re1, im1;
    fourier_shift_kernel(a.rows, shift0, re0, im0);
    fourier_shift_kernel(a.cols, shift1, re1, im1);
    double *d_re0, *d_im0, *d_re1, *d_im1;
    SAG(wk.alloc_raw((void**)&d_re0, re0.size() * sizeof(double)));
    SAG(wk.alloc_raw((void**)&d_im0, im0.size() * sizeof(double)));
    SAG(wk.alloc_raw((void**)&d_re1, re1.size() * sizeof(double)));
    SAG(wk.alloc_raw((void**)&d_im1, im1.size() * sizeof(double)));
    SAG(cudaMemcpyAsync(d_re0, re0.data(), re0.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    SAG(cudaMemcpyAsync(d_im0, im0.data(), im0.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    SAG(cudaMemcpyAsync(d_re1, re1.data(), re1.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    SAG(cudaMemcpyAsync(d_im1, im1.data(), im1.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    SAG(cudaStreamSynchronize(st));
    // Re[(h0 (*)_0)(h1 (*)_1) a] = Re h0 (*) Re h1 (*) a  -  Im h0 (*) Im h1 (*) a
    SagBuf t, out;
    SAG(wk.alloc(t, a.rows, a.cols));
    SAG(wk.alloc(out, a.rows, a.cols));
    SAG(sag_conv_circ(a.p, a.rows, a.cols, 1, d_re1, 1.0, 0, t.p, st));
    SAG(sag_conv_circ(t.p, a.rows, a.cols, 0, d_re0, 1.0, 0, out.p, st));
    bool any_im = false;
    for (double v : im0) any_im = any_im || v != 0.0;
    bool any_im1 = false;
    for (double v : im1) any_im1 = any_im1 || v != 0.0;
    if (any_im && any_im1) {
        SAG(sag_conv_circ(a.p, a.rows, a.cols, 1, d_im1, 1.0, 0, t.p, st));
        SAG(sag_conv_circ(t.p, a.rows, a.cols, 0, d_im0, -1.0, 1, out.p, st));
    }
    a = out;
    return PAOS_OK;
}

int sag_fit(SagWork& wk, SagBuf& a, int diff, int axis, double fill, cudaStream_t st) {
    if (diff == 0) return PAOS_OK;
    int rows = a.rows, cols = a.cols, row_off = 0, col_off = 0;
    int& n = axis == 0 ? rows : cols;
    int& off = axis == 0 ? row_off : col_off;
    if (diff < 0) {
        const int before = (-diff) / 2;
        off = -before;
        n = n - diff;
    } else {
        const int lo = diff / 2;
        off = lo;
        n = std::max(0, n - diff);
    }
    SagBuf out;
    SAG(wk.alloc(out, rows, cols));
    SAG(sag_padcrop(a.p, a.rows, a.cols, row_off, col_off, fill, rows, cols, out.p, st));
    a = out;
    return PAOS_OK;
}

}  // namespace

extern "C" int paos_fourier_shift_kernel(int n, double shift, double* re_out, double* im_out) {
    if (n < 1 || !re_out || !im_out) return fail(PAOS_ERR_ARG, "bad argument");
    std::vector<double> re, im;
    fourier_shift_kernel(n, shift, re, im);
    std::memcpy(re_out, re.data(), re.size() * sizeof(double));
    std::memcpy(im_out, im.data(), im.size() * sizeof(double));
    return PAOS_OK;
}

// prepare the n x n screen (and, if asked, its mask) of a raw map on the handle's stream; blocks until it is ready
static int prepare_grid_sag(paos_wfo* w, const double* host_sag, const unsigned char* host_mask, int nx, int ny, double delx, double dely,
                            double xdec, double ydec, double dx, double dy, double* screen, unsigned char* dmask) {
    int rc;
    const int n = w->n;
    cudaStream_t st = w->stream;
    // sizes first (wfo.py:776-814), so that an absurd pitch is refused before anything is allocated
    long cols = nx, rows = ny;
    int width_diff = (int)std::floor(((double)cols * delx - (double)n * dx) / delx);
    int height_diff = (int)std::floor(((double)rows * dely - (double)n * dy) / dely);
    {
        const double wide = (double)std::max<long>(cols - 2L * std::min(width_diff, 0), 1);
        const double tall = (double)std::max<long>(rows - 2L * std::min(height_diff, 0), 1);
        if (wide * tall > (double)(1L << 28))
            return fail(PAOS_ERR_ARG, "grid_sag: padding the %d x %d map (pitch %g x %g m) to the WFO extent (%g x %g m) needs more than 2^28 samples",
                        ny, nx, delx, dely, n * dx, n * dy);
    }
    SagWork wk;
    const size_t total = (size_t)nx * ny;
    double* raw;
    unsigned char* given = nullptr;
    SAG(wk.alloc_raw((void**)&raw, total * sizeof(double)));
    SAG(cudaMemcpyAsync(raw, host_sag, total * sizeof(double), cudaMemcpyHostToDevice, st));
    if (host_mask) {
        SAG(wk.alloc_raw((void**)&given, total));
        SAG(cudaMemcpyAsync(given, host_mask, total, cudaMemcpyHostToDevice, st));
    }
    SagBuf sag, mask;
    SAG(wk.alloc(sag, ny, nx));
    SAG(wk.alloc(mask, ny, nx));
    SAG(sag_split(raw, given, total, sag.p, mask.p, st));
    if (xdec != 0.0 || ydec != 0.0) {  // step 1 (wfo.py:768-773): fourier_shift(..., shift=(-xdec, -ydec)) acts on (axis 0, axis 1)
        if ((rc = sag_shift(wk, sag, -xdec, -ydec, st)) || (rc = sag_shift(wk, mask, -xdec, -ydec, st))) return rc;
    }
    // step 2: pad or crop to the extent of the WFO grid; an odd difference is made even by sampling twice as finely
    auto is_odd = [](int v) { return ((v % 2) + 2) % 2 == 1; };
    double sx = 1.0, sy = 1.0;
    if (is_odd(width_diff)) {
        sx = 2.0;
        delx /= 2.0;
        width_diff *= 2;
    }
    if (is_odd(height_diff)) {
        sy = 2.0;
        dely /= 2.0;
        height_diff *= 2;
    }
    if (sx != 1.0 || sy != 1.0) {
        const bool aa = sx < 1.0 || sy < 1.0;
        if ((rc = sag_rescale(wk, sag, sy, sx, aa, st)) || (rc = sag_rescale(wk, mask, sy, sx, aa, st))) return rc;
    }
    if ((rc = sag_fit(wk, sag, width_diff, 1, 0.0, st)) || (rc = sag_fit(wk, mask, width_diff, 1, 1.0, st))) return rc;
    if ((rc = sag_fit(wk, sag, height_diff, 0, 0.0, st)) || (rc = sag_fit(wk, mask, height_diff, 0, 1.0, st))) return rc;
    if (sag.rows < 1 || sag.cols < 1) return fail(PAOS_ERR_ARG, "grid_sag: the map does not overlap the WFO grid");
    // step 3: bring the map to the WFO pixel pitch; step 4: force the exact grid shape (can be one pixel off)
    sx = delx / dx;
    sy = dely / dy;
    if (sx != 1.0 || sy != 1.0) {
        const bool aa = sx < 1.0 || sy < 1.0;
        if ((rc = sag_rescale(wk, sag, sy, sx, aa, st)) || (rc = sag_rescale(wk, mask, sy, sx, aa, st))) return rc;
    }
    if (sag.rows != n || sag.cols != n) {
        const bool aa = (double)n / sag.cols < 1.0 || (double)n / sag.rows < 1.0;
        if ((rc = sag_resize(wk, sag, n, n, aa, st)) || (rc = sag_resize(wk, mask, n, n, aa, st))) return rc;
    }
    SAG(sag_finish(sag.p, mask.p, (size_t)n * n, screen, dmask, st));
    w->stats.kernel_launches += 8;
    SAG(cudaStreamSynchronize(st));  // the temporaries are freed when `wk` goes out of scope
    return PAOS_OK;
}

static int check_grid_sag_args(paos_wfo* w, const double* host_sag, int nx, int ny, double delx, double dely, double xdec, double ydec,
                               double dx, double dy, double wl) {
    if (!w || !host_sag) return fail(PAOS_ERR_ARG, "null argument");
    if (nx < 1 || ny < 1 || !(delx > 0) || !(dely > 0) || !(dx > 0) || !(dy > 0) || !(wl > 0) || !std::isfinite(xdec) || !std::isfinite(ydec))
        return fail(PAOS_ERR_ARG, "bad grid-sag geometry");
    return PAOS_OK;
}

extern "C" int paos_wfo_grid_sag(paos_wfo* w, const double* host_sag, const unsigned char* host_mask, int nx, int ny, double delx,
                                 double dely, double xdec, double ydec, double dx, double dy, double wl, double* screen_host_out,
                                 unsigned char* mask_host_out) {
    int rc = check_grid_sag_args(w, host_sag, nx, ny, delx, dely, xdec, ydec, dx, dy, wl);
    if (rc) return rc;
    if (w->recording && (screen_host_out || mask_host_out)) return fail(PAOS_ERR_STATE, "a recording handle cannot return the screen to the host");
    if ((rc = set_device(w))) return rc;
    const size_t nn = (size_t)w->n * w->n;
    double* screen;
    if ((rc = get_screen_for_immediate_write(w, &screen))) return rc;
    DevBuf dmask;
    if (mask_host_out) CU(dmask.alloc(nn));
    if ((rc = prepare_grid_sag(w, host_sag, host_mask, nx, ny, delx, dely, xdec, ydec, dx, dy, screen, dmask.as<unsigned char>()))) return rc;
    if (screen_host_out) CU(cudaMemcpyAsync(screen_host_out, screen, nn * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
    if (mask_host_out) CU(cudaMemcpyAsync(mask_host_out, dmask.p, nn, cudaMemcpyDeviceToHost, w->stream));
    if (screen_host_out || mask_host_out) CU(cudaStreamSynchronize(w->stream));
    return paos_wfo_phase_screen_device(w, screen, wl);
}

// Prepared screens shared between the jobs of a sweep (one map, hundreds of wavelengths): keyed by the caller's content
// key and every number the preparation depends on.  Entries live until paos_grid_sag_cache_clear (at most 32 are kept).
namespace {
struct SagCacheEntry {
    double* screen = nullptr;
    cudaEvent_t ready = nullptr;
};
std::mutex g_sag_mutex;
std::map<std::string, SagCacheEntry> g_sag_cache;
}  // namespace

extern "C" int paos_grid_sag_cache_clear(void) {
    std::lock_guard<std::mutex> lock(g_sag_mutex);
    for (auto& kv : g_sag_cache) {
        if (kv.second.ready) cudaEventDestroy(kv.second.ready);
        if (kv.second.screen) cudaFree(kv.second.screen);
    }
    g_sag_cache.clear();
    cudaGetLastError();
    return PAOS_OK;
}

// chain runner: a Grid Sag surface at whatever pitch the beam has there
static int chain_grid_sag(paos_wfo* w, const paos_surface& s, double dx, double dy, double wl) {
    int rc = check_grid_sag_args(w, s.sag, s.sag_nx, s.sag_ny, s.sag_delx, s.sag_dely, s.sag_xdec, s.sag_ydec, dx, dy, wl);
    if (rc) return rc;
    if ((rc = set_device(w))) return rc;
    if (s.sag_key == 0) {
        double* screen;
        if ((rc = get_screen_for_immediate_write(w, &screen))) return rc;
        if ((rc = prepare_grid_sag(w, s.sag, s.sag_mask, s.sag_nx, s.sag_ny, s.sag_delx, s.sag_dely, s.sag_xdec, s.sag_ydec, dx, dy, screen, nullptr)))
            return rc;
        return paos_wfo_phase_screen_device(w, screen, wl);
    }
    char key[512];
    snprintf(key, sizeof key, "%d|%llu|%d|%d|%d|%a|%a|%a|%a|%a|%a", w->device, (unsigned long long)s.sag_key, w->n, s.sag_nx, s.sag_ny,
             s.sag_delx, s.sag_dely, s.sag_xdec, s.sag_ydec, dx, dy);
    SagCacheEntry entry;
    {
        std::lock_guard<std::mutex> lock(g_sag_mutex);
        auto it = g_sag_cache.find(key);
        if (it != g_sag_cache.end()) entry = it->second;
    }
    if (!entry.screen) {
        double* screen = nullptr;
        CU(cudaMalloc((void**)&screen, (size_t)w->n * w->n * sizeof(double)));
        rc = prepare_grid_sag(w, s.sag, s.sag_mask, s.sag_nx, s.sag_ny, s.sag_delx, s.sag_dely, s.sag_xdec, s.sag_ydec, dx, dy, screen, nullptr);
        if (rc) {
            cudaFree(screen);
            return rc;
        }
        entry.screen = screen;
        CU(cudaEventCreateWithFlags(&entry.ready, cudaEventDisableTiming));
        CU(cudaEventRecord(entry.ready, w->stream));
        std::lock_guard<std::mutex> lock(g_sag_mutex);
        auto it = g_sag_cache.find(key);
        if (it != g_sag_cache.end()) {  // another thread prepared the same map meanwhile: keep one
            cudaEventDestroy(entry.ready);
            cudaFree(entry.screen);
            entry = it->second;
        } else if (g_sag_cache.size() < 32) {
            g_sag_cache[key] = entry;
        } else {
            // cache full: this one screen is kept alive for the life of the process rather than freed under a queued pass
        }
    }
    cudaEvent_t ready = entry.ready;
    rc = do_fn(w, 0, [=](cudaStream_t st) {
        cudaError_t e = cudaStreamWaitEvent(st, ready, 0);
        return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
    });
    if (rc) return rc;
    return paos_wfo_phase_screen_device(w, entry.screen, wl);
}

// ---- propagators ---------------------------------------------------------------------------------
static void push_sign(paos_wfo* w) {
    Op op{};
    op.kind = OP_SIGN;
    w->ops.push_back(op);
}
static void push_fft(paos_wfo* w, int dir) {
    Op op{};
    op.kind = OP_FFT;
    op.dir = dir;
    w->ops.push_back(op);
    w->stats.fft2_recorded++;
}
static void push_phase(paos_wfo* w, int kind, double c1, double c2, double dx, double dy) {
    Op op{};
    op.kind = OP_PHASE;
    op.tx.kind = op.ty.kind = kind;
    op.tx.c1 = op.ty.c1 = c1;
    op.tx.c2 = op.ty.c2 = c2;
    op.tx.d = dx;
    op.ty.d = dy;
    w->ops.push_back(op);
}

static int check_prop(paos_wfo* w, double wl, double dz, double dx, double dy) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    if (!(wl > 0.0) || !(dx > 0.0) || !(dy > 0.0) || !std::isfinite(dz) || dz == 0.0) return fail(PAOS_ERR_ARG, "bad propagation scalars");
    return PAOS_OK;
}

int paos_wfo_ptp(paos_wfo* w, double wl, double dz, double dx, double dy) {
    int rc = check_prop(w, wl, dz, dx, dy);
    if (rc) return rc;
    const double c = (M_PI * wl) * dz;  // numpy: (np.pi * wl * dz), left to right
    push_sign(w);
    push_fft(w, +1);
    push_phase(w, TERM_QFREQ, -1.0, c, dx, dy);
    push_fft(w, -1);
    push_sign(w);
    return PAOS_OK;
}

int paos_wfo_stw(paos_wfo* w, double wl, double dz, double dx, double dy) {
    int rc = check_prop(w, wl, dz, dx, dy);
    if (rc) return rc;
    const double c = (M_PI * wl) * dz;
    push_sign(w);
    push_fft(w, dz >= 0 ? +1 : -1);
    push_phase(w, TERM_QFREQ, 1.0, c, dx, dy);
    push_sign(w);
    return PAOS_OK;
}

int paos_wfo_wts(paos_wfo* w, double wl, double dz, double dx, double dy) {
    int rc = check_prop(w, wl, dz, dx, dy);
    if (rc) return rc;
    const double c = M_PI / (dz * wl);
    push_phase(w, TERM_QSPACE, 1.0, c, dx, dy);
    push_sign(w);
    push_fft(w, dz >= 0 ? +1 : -1);
    push_sign(w);
    return PAOS_OK;
}

int paos_wfo_fft2(paos_wfo* w, int inverse) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    push_sign(w);
    push_fft(w, inverse ? -1 : +1);
    push_sign(w);
    return PAOS_OK;
}

// ---- statistics ----------------------------------------------------------------------------------
int paos_zernike_points(int device, int nterms, const int* m, const int* n, const double* norm, const double* mat, const double* rho,
                        const double* phi, const unsigned char* mask, size_t npoints, double* out, double* cov_out) {
    if (!m || !n || !rho || !phi) return fail(PAOS_ERR_ARG, "null argument");
    if (!out && !cov_out) return fail(PAOS_ERR_ARG, "nothing to compute: out and cov_out are both null");
    if (nterms < 1 || nterms > ZERN_MAX) return fail(PAOS_ERR_ARG, "1..%d polynomials per call", ZERN_MAX);
    if (npoints == 0) return PAOS_OK;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(PAOS_ERR_CUDA, "no CUDA device: libpaos_b200 has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(PAOS_ERR_ARG, "device %d out of range (%d devices)", device, count);
    CU(cudaSetDevice(device));
    ZernParams Z;
    int rc = fill_zern(0, Z, 0, nterms, m, n, nullptr, norm, 1.0, 1.0, 1.0, 0.0, 0);
    if (rc) return rc;
    const size_t K = (size_t)nterms, vec = npoints * sizeof(double);
    DevBuf d_rho, d_phi, d_mask, d_stack, d_out, d_mat, d_cov;
    CU(d_rho.alloc(vec));
    CU(d_phi.alloc(vec));
    CU(d_stack.alloc(K * vec));
    CU(cudaMemcpy(d_rho.p, rho, vec, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_phi.p, phi, vec, cudaMemcpyHostToDevice));
    if (mask) {
        CU(d_mask.alloc(npoints));
        CU(cudaMemcpy(d_mask.p, mask, npoints, cudaMemcpyHostToDevice));
    }
    cudaError_t e = launch_zernike_points(Z, d_rho.as<double>(), d_phi.as<double>(), mask ? d_mask.as<unsigned char>() : nullptr, npoints,
                                          d_stack.as<double>(), nullptr);
    if (e != cudaSuccess) return fail(PAOS_ERR_CUDA, "zernike launch failed: %s", cudaGetErrorString(e));
    if (cov_out) {
        const int npairs = nterms * (nterms + 1) / 2;
        CU(d_cov.alloc((size_t)(npairs + 1) * sizeof(double)));
        e = launch_stack_cov(d_stack.as<double>(), d_rho.as<double>(), mask ? d_mask.as<unsigned char>() : nullptr, nterms, npoints,
                             d_cov.as<double>(), nullptr);
        if (e != cudaSuccess) return fail(PAOS_ERR_CUDA, "covariance launch failed: %s", cudaGetErrorString(e));
        std::vector<double> sums((size_t)npairs + 1);
        CU(cudaMemcpy(sums.data(), d_cov.p, sums.size() * sizeof(double), cudaMemcpyDeviceToHost));
        const double cnt = sums[(size_t)npairs];
        if (!(cnt > 0)) return fail(PAOS_ERR_STATE, "every point is masked: the covariance is undefined");
        size_t q = 0;
        for (int i = 0; i < nterms; ++i)
            for (int j = i; j < nterms; ++j, ++q) cov_out[(size_t)i * nterms + j] = cov_out[(size_t)j * nterms + i] = sums[q] / cnt;
    }
    if (out) {
        const double* src = d_stack.as<double>();
        if (mat) {
            CU(d_mat.alloc(K * K * sizeof(double)));
            CU(d_out.alloc(K * vec));
            CU(cudaMemcpy(d_mat.p, mat, K * K * sizeof(double), cudaMemcpyHostToDevice));
            e = launch_stack_transform(d_stack.as<double>(), d_mat.as<double>(), nterms, npoints, d_out.as<double>(), nullptr);
            if (e != cudaSuccess) return fail(PAOS_ERR_CUDA, "stack transform launch failed: %s", cudaGetErrorString(e));
            src = d_out.as<double>();
        }
        CU(cudaMemcpy(out, src, K * vec, cudaMemcpyDeviceToHost));
    }
    CU(cudaDeviceSynchronize());
    return PAOS_OK;
}

int paos_encircled_energy(paos_wfo* w, const void* psf_dev, double dx, double dy, double xc, double yc, double r_unit,
                          double r_max, int nbins, double* ee_dev_out) {
    if (!w || !psf_dev || !ee_dev_out) return fail(PAOS_ERR_ARG, "null argument");
    if (nbins < 1 || nbins > 4096) return fail(PAOS_ERR_ARG, "nbins %d outside [1, 4096]", nbins);
    if (!(r_unit > 0) || !(r_max > 0) || !(dx > 0) || !(dy > 0)) return fail(PAOS_ERR_ARG, "dx, dy, r_unit and r_max must be positive");
    int rc = set_device(w);
    if (rc) return rc;
    if (!w->ee_hist) {
        CU(cudaMalloc((void**)&w->ee_hist, (4096 + 2) * sizeof(double)));
        CU(cudaMemsetAsync(w->ee_hist, 0, (4096 + 2) * sizeof(double), w->stream));
    }
    const double inv_bin = (double)nbins / (r_unit * r_max);
    double* hist = w->ee_hist;
    const int n = w->n, is_float = w->dtype == PAOS_C128 ? 0 : 1;
    return do_fn(w, 2, [=](cudaStream_t st) {
        cudaError_t e = launch_encircled_energy(psf_dev, n, is_float, dx, dy, xc, yc, inv_bin, nbins, hist, ee_dev_out, st);
        return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "encircled-energy launch failed: %s", cudaGetErrorString(e));
    });
}

int paos_psf_peak(paos_wfo* w, const void* psf_dev, double* out_dev) {
    if (!w || !psf_dev || !out_dev) return fail(PAOS_ERR_ARG, "null argument");
    int rc = set_device(w);
    if (rc) return rc;
    const int n = w->n, is_float = w->dtype == PAOS_C128 ? 0 : 1;
    return do_fn(w, 1, [=](cudaStream_t st) {
        cudaError_t e = launch_psf_peak(psf_dev, n, is_float, out_dev, st);
        return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "peak launch failed: %s", cudaGetErrorString(e));
    });
}

int paos_screen_stats(paos_wfo* w, const double* screen_dev, double radius, double dx, double dy, double* out_dev) {
    if (!w || !screen_dev || !out_dev) return fail(PAOS_ERR_ARG, "null argument");
    if (!(radius > 0) || !(dx > 0) || !(dy > 0)) return fail(PAOS_ERR_ARG, "radius, dx and dy must be positive");
    int rc = set_device(w);
    if (rc) return rc;
    const int n = w->n;
    return do_fn(w, 1, [=](cudaStream_t st) {
        cudaError_t e = launch_screen_stats(screen_dev, n, radius, dx, dy, out_dev, st);
        return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "screen statistics launch failed: %s", cudaGetErrorString(e));
    });
}

int paos_crop_convert(paos_wfo* w, const void* src_dev, int x0, int y0, int nx, int ny, int to_float, void* dst_dev) {
    if (!w || !src_dev || !dst_dev) return fail(PAOS_ERR_ARG, "null argument");
    if (x0 < 0 || y0 < 0 || nx < 1 || ny < 1 || x0 + nx > w->n || y0 + ny > w->n) return fail(PAOS_ERR_ARG, "window outside the grid");
    int rc = set_device(w);
    if (rc) return rc;
    const int n = w->n, src_float = w->dtype == PAOS_C128 ? 0 : 1;
    return do_fn(w, 1, [=](cudaStream_t st) {
        cudaError_t e = launch_crop_convert(src_dev, n, src_float, x0, y0, nx, ny, to_float ? 1 : 0, dst_dev, st);
        return e == cudaSuccess ? PAOS_OK : fail(PAOS_ERR_CUDA, "crop launch failed: %s", cudaGetErrorString(e));
    });
}

int paos_wfo_stats(paos_wfo* w, paos_stats* out) {
    if (!w || !out) return fail(PAOS_ERR_ARG, "null argument");
    *out = w->stats;
    return PAOS_OK;
}

int paos_wfo_enable_timing(paos_wfo* w, int enable) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    w->timing = enable != 0;
    return PAOS_OK;
}

int paos_wfo_timing(paos_wfo* w, double* pass_ms, uint64_t* pass_launches) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    double ms = 0.0;
    uint64_t n = 0;
    for (int c = 0; c < 2; ++c)
        for (int k = 0; k <= KMAX; ++k) {
            ms += w->timed_ms[c][k];
            n += w->timed_n[c][k];
        }
    if (pass_ms) *pass_ms = ms;
    if (pass_launches) *pass_launches = n;
    return PAOS_OK;
}

int paos_wfo_timing_totals(paos_wfo* w, double* ms, uint64_t* launches, uint64_t* line_fft_sweeps, uint64_t* wavefronts, int reset) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    if (ms) *ms = w->timed_total_ms;
    if (launches) *launches = w->timed_total_launches;
    if (line_fft_sweeps) *line_fft_sweeps = w->timed_total_sweeps;
    if (wavefronts) *wavefronts = w->timed_total_items;
    if (reset) {
        w->timed_total_ms = 0.0;
        w->timed_total_launches = w->timed_total_sweeps = w->timed_total_items = 0;
    }
    return PAOS_OK;
}

int paos_wfo_timing_detail(paos_wfo* w, int col, int nfft, double* ms, uint64_t* launches, int reset) {
    if (!w) return fail(PAOS_ERR_ARG, "null handle");
    if (col < 0 || col > 1 || nfft < 0 || nfft > KMAX) return fail(PAOS_ERR_ARG, "bad timing bucket");
    if (ms) *ms = w->timed_ms[col][nfft];
    if (launches) *launches = w->timed_n[col][nfft];
    if (reset) {
        w->timed_ms[col][nfft] = 0.0;
        w->timed_n[col][nfft] = 0;
    }
    return PAOS_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// whole-chain execution: the per-surface loop of paos/core/run.py:30-228 with the pilot-beam scalar state
// machine of paos/classes/wfo.py:280-443, :547-572 in C++ doubles (same expressions, same order)
// ---------------------------------------------------------------------------------------------------
namespace {

struct Beam {
    double wl, z, w0, zw0, zr, rf, dx, dy, C, fratio;
    int n;
    char prop[4];
};

inline double sq(double x) { return x * x; }
inline double beam_wz(const Beam& b) { return b.w0 * std::sqrt(1.0 + sq((b.z - b.zw0) / b.zr)); }
inline char beam_io(const Beam& b, double z) { return std::fabs(z - b.zw0) < b.rf * b.zr ? 'I' : 'O'; }

int beam_lens(paos_wfo* w, Beam& b, double fl) {
    const double wz = beam_wz(b);
    double delta_z = b.z - b.zw0;
    const char before = beam_io(b, b.z);
    const double gCobj = delta_z / (sq(delta_z) + sq(b.zr));
    const double gCima = gCobj - 1.0 / fl;
    b.w0 = wz / std::sqrt(1.0 + sq(M_PI * sq(wz) * gCima / b.wl));
    b.zw0 = -gCima / (sq(gCima) + sq(b.wl / (M_PI * sq(wz)))) + b.z;
    b.zr = M_PI * sq(b.w0) / b.wl;
    const char after = beam_io(b, b.z);
    const double Cobj = (before == 'I' || b.C == 0.0) ? 0.0 : 1.0 / delta_z;
    delta_z = b.z - b.zw0;
    const double Cima = after == 'I' ? 0.0 : 1.0 / delta_z;
    b.C = Cima;
    double lens_phase = 1.0 / fl;
    if (before == 'O') lens_phase = lens_phase - Cobj;
    if (after == 'O') lens_phase = lens_phase + Cima;
    b.fratio = std::fabs(delta_z) / (2 * wz);
    return paos_wfo_quadphase(w, -2.0 * M_PI, 0.5 * lens_phase / b.wl, b.dx, b.dy);
}

int beam_magnify(Beam& b, double My, double Mx) {
    if (!(Mx > 0.0) || !(My > 0.0)) return fail(PAOS_ERR_ARG, "Negative magnification not implemented yet.");
    b.dx *= Mx;
    b.dy *= My;
    if (std::fabs(Mx - 1.0) < 1.0e-8) return PAOS_OK;
    double delta_z = b.z - b.zw0;
    double wz = beam_wz(b);
    delta_z *= sq(Mx);
    wz *= Mx;
    b.w0 *= Mx;
    b.zr *= sq(Mx);
    b.zw0 = b.z - delta_z;
    b.fratio = std::fabs(delta_z) / (2 * wz);
    return PAOS_OK;
}

void beam_medium(Beam& b, double n1n2) {
    double delta_z = b.z - b.zw0;
    delta_z /= n1n2;
    b.zr /= n1n2;
    b.wl *= n1n2;
    b.zw0 = b.z - delta_z;
    b.fratio /= n1n2;
}

int beam_ptp(paos_wfo* w, Beam& b, double dz) {
    if (std::fabs(dz) < 0.001 * b.wl) return PAOS_OK;
    if (b.C != 0) return fail(PAOS_ERR_STATE, "PTP wavefront should be planar");
    int rc = paos_wfo_ptp(w, b.wl, dz, b.dx, b.dy);
    b.z = b.z + dz;
    return rc;
}
int beam_stw(paos_wfo* w, Beam& b, double dz) {
    if (std::fabs(dz) < 0.001 * b.wl) return PAOS_OK;
    if (b.C == 0.0) return fail(PAOS_ERR_STATE, "STW wavefront should not be planar");
    int rc = paos_wfo_stw(w, b.wl, dz, b.dx, b.dy);
    const double fx1 = 1 * (1.0 / (b.n * b.dx)), fy1 = 1 * (1.0 / (b.n * b.dy));
    b.z = b.z + dz;
    b.C = 0.0;
    b.dx = (fx1 - 0.0) * b.wl * std::fabs(dz);
    b.dy = (fy1 - 0.0) * b.wl * std::fabs(dz);
    return rc;
}
int beam_wts(paos_wfo* w, Beam& b, double dz) {
    if (std::fabs(dz) < 0.001 * b.wl) return PAOS_OK;
    if (b.C != 0.0) return fail(PAOS_ERR_STATE, "WTS wavefront should be planar");
    int rc = paos_wfo_wts(w, b.wl, dz, b.dx, b.dy);
    b.z = b.z + dz;
    b.C = 1 / (b.z - b.zw0);
    b.dx = b.wl * std::fabs(dz) / (b.n * b.dx);
    b.dy = b.wl * std::fabs(dz) / (b.n * b.dy);
    return rc;
}
int beam_propagate(paos_wfo* w, Beam& b, double dz) {
    const double z1 = b.z, z2 = b.z + dz;
    const char p0 = beam_io(b, b.z), p1 = beam_io(b, z2);
    int rc = PAOS_OK;
    if (p0 == 'O') rc = beam_stw(w, b, b.zw0 - z1);
    else rc = (p1 == 'I') ? beam_ptp(w, b, dz) : beam_ptp(w, b, b.zw0 - z1);
    if (rc) return rc;
    if (p1 == 'O') rc = beam_wts(w, b, z2 - b.zw0);
    else if (p0 == 'O') rc = beam_ptp(w, b, z2 - b.zw0);
    b.prop[0] = p0;
    b.prop[1] = p1;
    b.prop[2] = 0;
    return rc;
}

// paos/core/coordinateBreak.py:7-72 : decenter, rotate into the new frame (scipy 'xyz' = rotations about the
// fixed x, then y, then z axes: R = Rz*Ry*Rx; the new frame sees R^T v), re-intersect with z = 0
void chain_coordinate_break(double vt[2], double vs[2], double xdec, double ydec, double xrot, double yrot, double zrot) {
    auto fin = [](double v) { return std::isfinite(v) ? v : 0.0; };
    xdec = fin(xdec); ydec = fin(ydec);
    const double a = fin(xrot) * M_PI / 180.0, bb = fin(yrot) * M_PI / 180.0, c = fin(zrot) * M_PI / 180.0;
    const double ca = std::cos(a), sa = std::sin(a), cb = std::cos(bb), sb = std::sin(bb), cc = std::cos(c), sc = std::sin(c);
    // R = Rz(c) * Ry(b) * Rx(a)
    const double R[3][3] = {{cc * cb, cc * sb * sa - sc * ca, cc * sb * ca + sc * sa},
                            {sc * cb, sc * sb * sa + cc * ca, sc * sb * ca - cc * sa},
                            {-sb, cb * sa, cb * ca}};
    auto applyT = [&](const double v[3], double o[3]) {
        for (int i = 0; i < 3; ++i) o[i] = R[0][i] * v[0] + R[1][i] * v[1] + R[2][i] * v[2];
    };
    const double r0[3] = {vs[0] - xdec, vt[0] - ydec, 0.0}, n0[3] = {vs[1], vt[1], 1.0};
    double n1[3], r1l[3];
    applyT(n0, n1);
    const double nz = n1[2];
    for (double& v : n1) v /= nz;
    applyT(r0, r1l);
    double r1[3];
    for (int i = 0; i < 3; ++i) r1[i] = r1l[i] - n1[i] * r1l[2] / n1[2];
    vt[0] = r1[1];
    vt[1] = n1[1];
    vs[0] = r1[0];
    vs[1] = n1[0];
}

void fill_snapshot(paos_snapshot& s, int index, const Beam& b, const double vt[2], const double vs[2]) {
    s.surface = index;
    std::memcpy(s.propagator, b.prop, 4);
    s.wl = b.wl; s.z = b.z; s.w0 = b.w0; s.zw0 = b.zw0; s.zr = b.zr; s.dx = b.dx; s.dy = b.dy; s.C = b.C; s.fratio = b.fratio;
    s.wz = beam_wz(b);
    s.distancetofocus = b.zw0 - b.z;
    s.vt[0] = vt[0]; s.vt[1] = vt[1]; s.vs[0] = vs[0]; s.vs[1] = vs[1];
}

}  // namespace

extern "C" int paos_chain_run(paos_wfo* w, double pupil_diameter, double wavelength, double zoom, double us, double ut,
                              const paos_surface* surfaces, int n_surfaces, paos_snapshot* snapshots, int max_snapshots,
                              int* n_snapshots, paos_snapshot* final_state) {
    if (!w || (!surfaces && n_surfaces > 0)) return fail(PAOS_ERR_ARG, "null argument");
    struct Book {  // host time of an unbatched chain (a recording handle is timed by paos_batch_chain_run)
        paos_wfo* w;
        std::chrono::steady_clock::time_point t0;
        ~Book() {
            if (w) w->stats.host_plan_us += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        }
    } book{w->recording ? nullptr : w, std::chrono::steady_clock::now()};
    if (!(zoom > 0) || !(pupil_diameter > 0) || !(wavelength > 0)) return fail(PAOS_ERR_ARG, "zoom, beam diameter and wavelength must be positive");
    int rc = paos_wfo_reset(w);
    if (rc) return rc;
    Beam b{};
    b.n = w->n;
    b.wl = wavelength;
    b.z = 0.0;
    b.w0 = pupil_diameter / 2.0;
    b.zw0 = 0.0;
    b.zr = M_PI * sq(b.w0) / wavelength;
    b.rf = 2.0;
    b.dx = pupil_diameter * zoom / b.n;
    b.dy = pupil_diameter * zoom / b.n;
    b.C = 0.0;
    b.fratio = INFINITY;
    b.prop[0] = 0;
    double vt[2] = {0.0, ut}, vs[2] = {0.0, us};
    int nsnap = 0;
    for (int i = 0; i < n_surfaces; ++i) {
        const paos_surface& s = surfaces[i];
        if (s.type == PAOS_SURF_COORDBREAK) chain_coordinate_break(vt, vs, s.xdec, s.ydec, s.xrot, s.yrot, 0.0);
        if (s.has_aperture) {
            const double xc = std::isfinite(s.ap_xc) ? s.ap_xc : vs[0], yc = std::isfinite(s.ap_yc) ? s.ap_yc : vt[0];
            const double xrad = s.ap_xrad * std::sqrt(1 / (sq(vs[1]) + 1)), yrad = s.ap_yrad * std::sqrt(1 / (sq(vt[1]) + 1));
            if (std::isfinite(xrad) && std::isfinite(yrad)) {
                const double ixc = (xc - vs[0]) / b.dx + b.n / 2.0, iyc = (yc - vt[0]) / b.dy + b.n / 2.0;
                rc = paos_wfo_aperture(w, s.ap_shape, ixc, iyc, xrad / b.dx, yrad / b.dy, 0.0, s.ap_obscuration);
                if (rc) return rc;
            }
        }
        if (s.is_stop && (rc = paos_wfo_make_stop(w))) return rc;
        if (s.type == PAOS_SURF_ZERNIKE) {
            const double radius = std::isfinite(s.zernike_radius) ? s.zernike_radius : beam_wz(b);
            rc = paos_wfo_zernike(w, s.zernike_terms, s.zernike_m, s.zernike_n, s.zernike_coef, radius, b.dx, b.dy, 0.0,
                                  s.zernike_origin, b.wl, nullptr);
            if (rc) return rc;
        } else if (s.type == PAOS_SURF_SCREEN) {
            // the map was resampled on the host for one pixel pitch (wfo.py:848-862 uses the pitch at the surface)
            if (std::fabs(b.dx - s.screen_dx) > 1e-12 * std::fabs(b.dx) || std::fabs(b.dy - s.screen_dy) > 1e-12 * std::fabs(b.dy))
                return fail(PAOS_ERR_UNSUPPORTED, "grid-sag screen of surface %d was prepared for pitch (%g, %g) but the beam is sampled at (%g, %g)",
                            i, s.screen_dx, s.screen_dy, b.dx, b.dy);
            rc = s.screen_on_device ? paos_wfo_phase_screen_device(w, s.screen, b.wl) : paos_wfo_phase_screen(w, s.screen, b.wl);
            if (rc) return rc;
        } else if (s.type == PAOS_SURF_GRIDSAG) {
            if ((rc = chain_grid_sag(w, s, b.dx, b.dy, b.wl))) return rc;
        } else if (s.type == PAOS_SURF_PSD) {
            const double f_nyq = 0.5 * std::sqrt(1.0 / sq(b.dx) + 1.0 / sq(b.dy));
            if (!(s.psd[5] <= f_nyq)) return fail(PAOS_ERR_ARG, "fmax must be less than or equal to f_Nyq (%g)", f_nyq);
            rc = paos_wfo_psd(w, s.psd[0], s.psd[1], s.psd[2], s.psd[3], s.psd[4], s.psd[5], s.psd[6], s.psd[7], b.dx, b.dy, b.wl,
                              s.psd_noise1, s.psd_noise2, s.psd_seed, nullptr);
            if (rc) return rc;
        }
        if (s.save) {
            if (s.read_what >= 0 && s.read_dst) {
                // the read-out of the last surface may drop the field: nothing can observe it before the next reset
                const bool fin = s.read_discard && i == n_surfaces - 1 && s.read_what != PAOS_READ_WFO;
                rc = fin ? paos_wfo_read_device_final(w, s.read_what, s.read_dst) : paos_wfo_read_device(w, s.read_what, s.read_dst);
                if (rc) return rc;
            }
            if (snapshots && nsnap < max_snapshots) fill_snapshot(snapshots[nsnap], i, b, vt, vs);
            ++nsnap;
        }
        const double At = s.abcd_t[0], Bt = s.abcd_t[1], Ct = s.abcd_t[2], Dt = s.abcd_t[3];
        const double As = s.abcd_s[0], Bs = s.abcd_s[1], Cs = s.abcd_s[2], Ds = s.abcd_s[3];
        const double Mt = (At * Dt - Bt * Ct) / Dt, Ms = (As * Ds - Bs * Cs) / Ds;
        const double power = -Ct / Mt;
        const double fl = power == 0 ? INFINITY : s.cout_t / power;
        const double T = s.cout_t * (Bt / Dt);
        const double n1n2 = Dt * Mt;
        if (Mt != 1.0 || Ms != 1.0) {
            if ((rc = beam_magnify(b, Mt, Ms))) return rc;
        }
        if (std::fabs(n1n2) != 1.0) beam_medium(b, n1n2);
        if (std::isfinite(fl) && (rc = beam_lens(w, b, fl))) return rc;
        if (std::isfinite(T) && std::fabs(T) > 1e-10 && (rc = beam_propagate(w, b, T))) return rc;
        const double vt0 = At * vt[0] + Bt * vt[1], vt1 = Ct * vt[0] + Dt * vt[1];
        const double vs0 = As * vs[0] + Bs * vs[1], vs1 = Cs * vs[0] + Ds * vs[1];
        vt[0] = vt0; vt[1] = vt1; vs[0] = vs0; vs[1] = vs1;
    }
    if (w->dropped) w->ops.clear();  // what the last surface queued behind its final read-out can never be observed
    if (n_snapshots) *n_snapshots = nsnap;
    if (final_state) fill_snapshot(*final_state, n_surfaces, b, vt, vs);
    return PAOS_OK;
}
