/*
 * paos_b200.h -- C ABI of libpaos_b200.so, the B200 (sm_100a) replacement for the array arithmetic of
 * PAOS's Fresnel-propagation hot path.
 *
 * The reference (arielmission-space/PAOS v1.2.12) has no FFI boundary of its own: its operator API for
 * this path is the Python class paos.WFO (paos/classes/wfo.py) and the chain driver paos.core.run.run
 * (paos/core/run.py).  This header is the boundary a drop-in inserts *underneath* that class: the Python
 * host keeps the Gaussian pilot-beam scalars (wl, z, w0, zw0, zr, dx, dy, C, fratio) exactly as the
 * reference does and calls one entry point per array operation.  Each entry point cites the reference
 * lines whose array arithmetic it replaces.
 *
 * Conventions
 *  - plain C types only; every function returns 0 on success and a negative code on error, with a
 *    human readable message available from paos_last_error() (thread local);
 *  - a wavefront handle owns (or borrows) one N x N complex array, row-major [y][x], resident in HBM for
 *    the whole surface chain; dtype PAOS_C128 matches numpy complex128, PAOS_C64 is the stated
 *    single-precision mode;
 *  - operations are *recorded* and executed lazily: the library fuses the elementwise factors (fftshift
 *    signs, Fresnel chirps, lens phase, aperture masks, phase screens, the stop normalisation) into the
 *    first/last stage of the FFT passes when the queue is flushed by a read, an explicit
 *    paos_wfo_flush, or paos_wfo_sync.  No call except *_read, *_sync and paos_wfo_stats blocks the host;
 *  - all device work is enqueued on the handle's CUDA stream;
 *  - there is no CPU fallback: every entry point fails with PAOS_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef PAOS_B200_H
#define PAOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAOS_ABI_VERSION 3
#define PAOS_MAX_CHAINED_FFTS 16 /* line FFTs one pass kernel can chain */

enum {
    PAOS_OK = 0,
    PAOS_ERR_ARG = -1,     /* bad argument (grid not 2^n in 64..4096, null pointer, ...) */
    PAOS_ERR_CUDA = -2,    /* CUDA runtime error or no usable device */
    PAOS_ERR_STATE = -3,   /* operation not valid in the current state */
    PAOS_ERR_UNSUPPORTED = -4
};

enum { PAOS_C128 = 0, PAOS_C64 = 1 };

/* what paos_wfo_read materialises */
enum {
    PAOS_READ_WFO = 0,       /* complex field (wfo.py:163-164)          -> 2*N*N reals       */
    PAOS_READ_AMPLITUDE = 1, /* |wfo|, numpy abs (wfo.py:167-168)       -> N*N reals         */
    PAOS_READ_PHASE = 2,     /* angle(wfo) (wfo.py:171-172)             -> N*N reals         */
    PAOS_READ_PSF = 3        /* |wfo|^2 (paos/core/plot.py:125-130)     -> N*N reals         */
};

enum { PAOS_SHAPE_ELLIPSE = 0, PAOS_SHAPE_RECT = 1 };

typedef struct paos_wfo paos_wfo; /* opaque */

/* ---- library ------------------------------------------------------------------------------- */
int paos_abi_version(void);
const char *paos_last_error(void);
/* "libpaos_b200 sm_100a source <sha1 of csrc/ + include/ + flags> compiled <date>": which tree this binary was made from */
const char *paos_build_info(void);
/* sizeof of a public struct as this library was compiled (0: paos_surface, 1: paos_snapshot, 2: paos_stats, 3: paos_chain_args; -1 for
 * anything else): lets a foreign-function binding check its own mirror of the layout */
long paos_abi_struct_size(int which);
/* number of usable sm_100 devices (0 when none); never fails */
int paos_device_count(void);

/* ---- lifetime (wfo.py:99-120: _wfo = ones((N, N), complex128)) ------------------------------ */
/* n: grid size, power of two in [64, 4096].  device: CUDA ordinal.  stream: a cudaStream_t cast to
 * void* (NULL = the library creates a non-blocking stream for the handle).  borrowed: optional device
 * pointer to n*n complex elements owned by the caller (e.g. a torch CUDA tensor); NULL = the library
 * allocates.  The field starts as all ones and is not materialised until first needed. */
int paos_wfo_create(paos_wfo **out, int n, int dtype, int device, void *stream, void *borrowed);
int paos_wfo_destroy(paos_wfo *w);
/* reset to the initial all-ones field (re-use of a handle for the next chain) */
int paos_wfo_reset(paos_wfo *w);
/* same thing under the name the reference's constructor suggests (wfo.py:118: np.ones) */
int paos_wfo_fill_ones(paos_wfo *w);
/* run everything recorded so far (asynchronous) */
int paos_wfo_flush(paos_wfo *w);
/* flush and wait for the stream */
int paos_wfo_sync(paos_wfo *w);
/* Flush, then write out the zeros that blanked passes left unwritten ("virtual zeros": lines an aperture blanks are
 * neither loaded nor stored, the planner only remembers the band outside which the field is zero), so that a
 * *borrowed* field buffer holds the complete complex array (wfo.py:163-164).  paos_wfo_read* do this themselves. */
int paos_wfo_materialize(paos_wfo *w);

/* ---- data movement --------------------------------------------------------------------------- */
/* replace the field by host data (n*n complex, dtype of the handle); discards recorded operations */
int paos_wfo_upload(paos_wfo *w, const void *host_src);
/* flush, compute `what` on the device and copy it to host memory (blocking).  dst holds doubles for
 * PAOS_C128 handles and floats for PAOS_C64 handles. */
int paos_wfo_read(paos_wfo *w, int what, void *host_dst);
/* same, into device memory (asynchronous on the handle's stream) */
int paos_wfo_read_device(paos_wfo *w, int what, void *dev_dst);
/* Same for |.|, angle or |.|^2 when this read-out is the last use of the wavefront (the IMAGE_PLANE PSF of a sweep,
 * pipeline.py:112-115 `light_output`): the pass that produces it does not store the complex field.  Afterwards
 * the handle only accepts paos_wfo_reset / _upload / _destroy (anything else: PAOS_ERR_STATE). */
int paos_wfo_read_device_final(paos_wfo *w, int what, void *dev_dst);
/* replace the field by n*n complex elements already in device memory (asynchronous; a no-op copy when
 * dev_src is the handle's own buffer, i.e. the caller wrote into the borrowed tensor) */
int paos_wfo_upload_device(paos_wfo *w, const void *dev_src);

/* ---- elementwise operators ------------------------------------------------------------------- */
/* wfo.py:203-278 (aperture): multiply by the mask (or 1-mask when obscuration != 0).  All lengths in
 * pixels as the reference passes them to photutils: centre (ixc, iyc) = (xc/dx + N/2, yc/dy + N/2);
 * ellipse: semi-axes (ihx, ihy), exact pixel/ellipse overlap; rectangle: full sides (ihx, ihy), 32x32
 * sub-pixel sampling.  theta in radians, counter-clockwise from +x (photutils convention); run.py:114-121 never
 * tilts, so theta != 0 takes a slower per-pixel path. */
int paos_wfo_aperture(paos_wfo *w, int shape, double ixc, double iyc, double ihx, double ihy,
                      double theta, int obscuration);
/* wfo.py:195-201 (make_stop): divide by sqrt(sum |wfo|^2) */
int paos_wfo_make_stop(paos_wfo *w);
/* wfo.py:359-366 (lens): multiply by exp(i * c1*c2 * (x^2 + y^2)), x = (j - N/2)*dx, y = (i - N/2)*dy.
 * The product c1*c2 is formed in double-double so the host can pass the reference's two rounded
 * factors (c1 = -2*pi, c2 = 0.5*lens_phase/wl) unchanged. */
int paos_wfo_quadphase(paos_wfo *w, double c1, double c2, double dx, double dy);
/* wfo.py:650-652, :867-869, :947: multiply by exp(2*pi*i * screen / wl); screen = n*n host doubles
 * (metres, 0 where masked).  The library copies it to the device. */
int paos_wfo_phase_screen(paos_wfo *w, const double *host_screen, double wl);
/* same with a device pointer that must stay valid until the next flush */
int paos_wfo_phase_screen_device(paos_wfo *w, const double *dev_screen, double wl);
/* wfo.py:574-654 + zernike.py:63-109,:210-247: wfe = sum_k coef[k] * Z_k(rho, phi) inside rho <= 1
 * (0 outside), then multiply by exp(2*pi*i*wfe/wl).  m[k], n[k]: azimuthal / radial numbers from
 * Zernike.j2mn; coef[k] = Z[k]*norm[k] (host applies the normalisation).  origin: 0 = 'x'
 * (phi = atan2(y, x) + offset), 1 = 'y' (phi = atan2(x, y) + offset); offset in radians.
 * wfe_host_out: optional n*n doubles receiving the screen (blocking when non-NULL). */
int paos_wfo_zernike(paos_wfo *w, int nterms, const int *m, const int *n, const double *coef,
                     double radius, double dx, double dy, double offset, int origin, double wl,
                     double *wfe_host_out);
/* the same with a pupil mask (wfo.py:624-628: numpy masked-array convention, n*n host bytes, non-zero = masked
 * pixel -> wfe 0 there); host_mask may be NULL */
int paos_wfo_zernike_masked(paos_wfo *w, int nterms, const int *m, const int *n, const double *coef,
                            double radius, double dx, double dy, double offset, int origin, double wl,
                            const unsigned char *host_mask, double *wfe_host_out);
/* zernike.py:293-317 (Zernike.cov), the reduction behind PolyOrthoNorm (zernike.py:388-402):
 * cov[i*K+j] = mean over the unmasked pixels (rho <= 1 and mask == 0) of norm[i]*Z_i * norm[j]*Z_j, K = nterms <= 64.
 * Blocking (returns the K x K matrix in host memory); does not touch the wavefront.  norm may be NULL (ones). */
int paos_zernike_cov(paos_wfo *w, int nterms, const int *m, const int *n, const double *norm, double radius,
                     double dx, double dy, double offset, int origin, const unsigned char *host_mask,
                     double *cov_host_out);
/* wfo.py:656-871 (grid_sag) in full: the raw map (ny x nx host doubles at pitch delx x dely, decentred by xdec, ydec pixels;
 * host_mask: optional ny*nx bytes, non-zero = masked, NULL = mask the non-finite and the zero samples like the reference)
 * is masked, recentred by a Fourier shift, padded / cropped to the WFO extent and resampled to the pitch (dx, dy) of the
 * wavefront with cubic B-splines and Gaussian anti-aliasing -- all on the device (csrc/sag_kernels.cu) -- then applied as
 * exp(2*pi*i*sag/wl) with masked pixels at 0.  screen_host_out / mask_host_out: optional n*n doubles / bytes receiving the
 * resampled map and its mask.  Blocks the host until the map is ready (input preparation, not the per-wavelength path). */
int paos_wfo_grid_sag(paos_wfo *w, const double *host_sag, const unsigned char *host_mask, int nx, int ny, double delx,
                      double dely, double xdec, double ydec, double dx, double dy, double wl, double *screen_host_out,
                      unsigned char *mask_host_out);
int paos_grid_sag_cache_clear(void);
/* impulse response of scipy.ndimage.fourier_shift along an axis of n samples (what paos_wfo_grid_sag convolves with);
 * host-only helper exported for the CPU tests */
int paos_fourier_shift_kernel(int n, double shift, double *re_out, double *im_out);
/* wfo.py:873-949 + psd.py:113-148: surface-error screen with power spectrum A/(B+(f/fknee)^C) between
 * fmin and fmax plus white roughness SR, times 2*unit_scale, applied as a phase screen.
 * noise1/noise2: n*n host doubles replacing the reference's two np.random.randn draws (bit-parity
 * mode); when both are NULL the library draws them on the device from `seed` (Philox + Box-Muller,
 * statistical parity only).  wfe_host_out optional as above. */
int paos_wfo_psd(paos_wfo *w, double A, double B, double C, double fknee, double fmin, double fmax,
                 double SR, double unit_scale, double dx, double dy, double wl, const double *noise1,
                 const double *noise2, uint64_t seed, double *wfe_host_out);

/* ---- propagators ------------------------------------------------------------------------------ */
/* wfo.py:462-472 (ptp): ifftshift, fft2(ortho), * exp(-i*pi*wl*dz*(fx^2+fy^2)), ifft2(ortho), fftshift */
int paos_wfo_ptp(paos_wfo *w, double wl, double dz, double dx, double dy);
/* wfo.py:493-509 (stw): ifftshift, fft2|ifft2 by sign of dz, * exp(+i*pi*wl*dz*(fx^2+fy^2)), fftshift */
int paos_wfo_stw(paos_wfo *w, double wl, double dz, double dx, double dy);
/* wfo.py:530-545 (wts): * exp(+i*pi/(dz*wl)*(x^2+y^2)), ifftshift, fft2|ifft2 by sign of dz, fftshift */
int paos_wfo_wts(paos_wfo *w, double wl, double dz, double dx, double dy);
/* plain shifted transform: fftshift(fft2|ifft2(ifftshift(wfo), norm="ortho")); used by tests */
int paos_wfo_fft2(paos_wfo *w, int inverse);

/* ---- whole-chain execution (paos/core/run.py:30-228) ------------------------------------------------
 * The reference's per-surface loop, including the Gaussian pilot-beam scalar state machine of
 * wfo.py:304-443 and :547-572, run natively so that a sweep over many wavelengths is not limited by the
 * Python interpreter.  One record per *non-ignored* surface after INIT, in chain order.  Every step calls
 * the same internal operators as the per-operation entry points above, so results are those of
 * paos_b200.run. */
enum {
    PAOS_SURF_GENERIC = 0,   /* Standard, Paraxial Lens, ABCD: aperture / stop / ABCD only        */
    PAOS_SURF_COORDBREAK = 1,
    PAOS_SURF_ZERNIKE = 2,
    PAOS_SURF_SCREEN = 3,    /* Grid Sag already on the WFO grid (metres, 0 where masked)           */
    PAOS_SURF_PSD = 4,
    PAOS_SURF_GRIDSAG = 5    /* Grid Sag as the lens file gives it: raw map, resampled on the device at the surface's pitch */
};
typedef struct paos_surface {
    int type;
    int is_stop;
    int save;              /* take a snapshot (paos_snapshot) before magnification/lens/propagation */
    int has_aperture;
    int ap_shape;          /* PAOS_SHAPE_ELLIPSE ("elliptical") or PAOS_SHAPE_RECT ("rectangular") */
    int ap_obscuration;
    int read_what;         /* PAOS_READ_* to materialise at this surface when save != 0, -1 = none */
    int zernike_terms;
    int zernike_origin;    /* 0 = 'x', 1 = 'y' */
    int screen_on_device;  /* PAOS_SURF_SCREEN: `screen` is a device pointer that stays valid until the chain has run */
    int read_discard;      /* last surface only: its read-out is final (paos_wfo_read_device_final)                  */
    int reserved;
    double ap_xrad, ap_yrad, ap_xc, ap_yc;      /* as in opt_chain[..]["aperture"]; NaN centre = follow the chief ray */
    double abcd_t[4], abcd_s[4];                /* row-major A, B, C, D of item["ABCDt"], item["ABCDs"]   */
    double cout_t;                              /* item["ABCDt"].cout (+1 / -1)                         */
    double xdec, ydec, xrot, yrot;              /* Coordinate Break (NaN = 0)                            */
    double zernike_radius;                      /* NaN = use wz (run.py:131)                             */
    double screen_dx, screen_dy;                /* pixel pitch `screen` was resampled to (wfo.py:848-862); the chain stops
                                                   with PAOS_ERR_UNSUPPORTED if the beam's pitch differs at the surface */
    double psd[8];                              /* A, B, C, fknee, fmin, fmax, SR, unit_scale            */
    uint64_t psd_seed;
    const int *zernike_m, *zernike_n;           /* host arrays [zernike_terms]                           */
    const double *zernike_coef;                 /* Z[k] * norm[k]                                        */
    const double *screen;                       /* PAOS_SURF_SCREEN: n*n doubles (host, or device if flagged) */
    const double *psd_noise1, *psd_noise2;      /* host arrays or NULL (device RNG from psd_seed)        */
    void *read_dst;                             /* device destination of the read-out                    */
    /* PAOS_SURF_GRIDSAG (wfo.py:656-871): sag = sag_ny x sag_nx host doubles in metres, sag_mask = optional bytes (non-zero =
     * masked; NULL = mask non-finite and zero samples), pitch sag_delx x sag_dely, decentre sag_xdec, sag_ydec pixels.
     * sag_key != 0 identifies the map's content: jobs that carry the same key (one map, many wavelengths) share the
     * prepared screen whenever the pitch at the surface is the same (paos_grid_sag_cache_clear releases them). */
    const double *sag;
    const unsigned char *sag_mask;
    int sag_nx, sag_ny;
    double sag_delx, sag_dely, sag_xdec, sag_ydec;
    uint64_t sag_key;
} paos_surface;
typedef struct paos_snapshot {
    int surface;           /* index into the surface array */
    char propagator[4];    /* "", "II", "IO", "OI", "OO" (of the previous propagation, wfo.py:572) */
    double wl, z, w0, zw0, zr, dx, dy, C, fratio, wz, distancetofocus;
    double vt[2], vs[2];   /* chief-ray vectors at the surface */
} paos_snapshot;
/* Resets the handle, runs the chain for one wavelength / field and fills one paos_snapshot per saved
 * surface (at most max_snapshots; *n_snapshots receives the count) plus the final beam state in
 * final_state (may be NULL).  Asynchronous like every other call. */
int paos_chain_run(paos_wfo *w, double pupil_diameter, double wavelength, double zoom, double us, double ut,
                   const paos_surface *surfaces, int n_surfaces, paos_snapshot *snapshots, int max_snapshots,
                   int *n_snapshots, paos_snapshot *final_state);

/* ---- batched execution: the `batch` of SURVEY.md section 8(b) --------------------------------------------------
 * Independent propagations (wavelengths, field points, Monte-Carlo realizations: the joblib fan-out of
 * paos/core/pipeline.py:140-150) that are at the same point of their chains share their kernel launches: ONE grid runs the
 * same-axis FFT pass of every wavefront of the batch (a batch index on the grid; each item keeps its own phase tables,
 * masks, blank-line range and output pointer), one launch builds all their phase tables, one reduces all their stops.
 * A pass of a single 2048^2 wavefront whose aperture blanks most lines is a wave or two of CTAs and is bound by the latency
 * of one CTA's chain; the batch fills the machine and amortises the launch.  Results are bit-identical to running the
 * handles one by one.
 *
 * paos_wfo_begin_record: from now on the handle records its device work instead of launching it (host-blocking calls --
 * paos_wfo_read, paos_wfo_sync, paos_zernike_cov, the *_host_out arguments -- fail with PAOS_ERR_STATE meanwhile).  Host
 * arrays handed to a recording handle (screens, masks, PSD noise) are read when the batch executes: keep them alive until
 * paos_batch_execute has returned.
 * paos_batch_execute: plans what is still queued on each handle and executes the recorded programs of nb <=
 * paos_batch_capacity() handles in lockstep on their common stream (same grid size, precision, device and stream
 * required); the handles leave recording mode.  Asynchronous.  On error every handle of the batch is reset to a
 * non-recording, empty state.
 * paos_batch_chain_run: begin_record + paos_chain_run on every handle (args[b]) + paos_batch_execute. */
typedef struct paos_chain_args {
    double pupil_diameter, wavelength, zoom, us, ut;
    const paos_surface *surfaces;
    int n_surfaces;
    int max_snapshots;
    paos_snapshot *snapshots;
    int *n_snapshots;
    paos_snapshot *final_state;
} paos_chain_args;
int paos_batch_capacity(void);
int paos_wfo_begin_record(paos_wfo *w);
int paos_batch_execute(paos_wfo *const *ws, int nb);
int paos_batch_chain_run(paos_wfo *const *ws, int nb, const paos_chain_args *args);

/* ---- the single collective: gather of the per-GPU result stacks ---------------------------
 * One process per GPU; propagation needs no communication, only the final stack (PSFs, or the much smaller encircled-energy
 * curves) travels to one rank over NVLink.  NCCL is bound at run time (dlopen), so the library loads without it;
 * PAOS_ERR_UNSUPPORTED when it is missing.  Messages of these four calls: paos_comm_last_error().
 *   paos_comm_unique_id: 128 bytes identifying a new communicator (call on one rank, hand the bytes to the others by any
 *     means, e.g. torch.distributed.broadcast_object_list);
 *   paos_comm_create: collective over all ranks of the communicator;
 *   paos_gather_psf: rank q contributes bytes_per_rank[q] bytes at local_dev; on `root` the block of rank q lands at
 *     dst_dev + dst_offsets[q] (dst_offsets NULL: back to back in rank order).  A sweep gathers chunk by chunk -- the PSFs
 *     of the batch that has just finished go to their final rows of the [sum of counts, N, N] stack while later
 *     wavelengths still propagate.  bytes_per_rank (one entry per rank) must be the same on every rank.  Asynchronous on
 *     `stream` (a cudaStream_t); calls on one communicator must be issued in the same order on every rank. */
typedef struct paos_comm paos_comm;
const char *paos_comm_last_error(void);
int paos_comm_unique_id(void *id128);
int paos_comm_create(paos_comm **out, const void *id128, int rank, int world, int device);
int paos_comm_destroy(paos_comm *c);
int paos_gather_psf(paos_comm *c, const void *local_dev, const size_t *bytes_per_rank, const size_t *dst_offsets, void *dst_dev,
                    int root, void *stream);

/* ---- the stand-alone polynomial classes (paos/classes/zernike.py:63-109 `Zernike`, :293-317 `cov`, :388-402 `PolyOrthoNorm`) ----
 * Polynomials at arbitrary points: rho, phi, mask (numpy masked-array convention, may be NULL) are host arrays of npoints
 * entries; points with rho > 1 or mask != 0 give 0.  norm may be NULL (ones).  nterms <= 64.
 *   cov_out (nterms x nterms, may be NULL): mean over the unmasked points of Z_i * Z_j (before any transform);
 *   out (nterms x npoints, may be NULL): the stack, multiplied from the left by the row-major nterms x nterms matrix
 *   mat when that is not NULL (the Gram-Schmidt matrix of PolyOrthoNorm).
 * Stateless and synchronous (its own device buffers on `device`); an input-preparation call, not part of the
 * per-wavelength path. */
int paos_zernike_points(int device, int nterms, const int *m, const int *n, const double *norm, const double *mat,
                        const double *rho, const double *phi, const unsigned char *mask, size_t npoints, double *out,
                        double *cov_out);

/* ---- encircled energy (docs/source/user/aberration/index.rst:47-67; the reference documents it but has no code) ---
 * psf_dev: n*n reals on the handle's device (double for a complex128 handle, float for complex64), e.g. the
 * destination of paos_wfo_read_device(PAOS_READ_PSF).  Pixel (ix, iy) sits at x = (ix - xc)*dx, y = (iy - yc)*dy
 * (the WFO grid has xc = yc = n/2); r_unit = F# * lambda converts metres to the normalised radius R_f; the curve is
 * sampled at R_f = (k + 1) * r_max / nbins, k = 0 .. nbins-1.  ee_dev_out receives nbins + 1 doubles on the device:
 * the fraction of the total energy inside each radius, then the total energy itself.  nbins <= 4096.  Asynchronous
 * on the handle's stream; the curve is 8*(nbins+1) bytes instead of the 8*n*n of the PSF (what a sweep gathers). */
int paos_encircled_energy(paos_wfo *w, const void *psf_dev, double dx, double dy, double xc, double yc, double r_unit,
                          double r_max, int nbins, double *ee_dev_out);

/* ---- Strehl ratio (docs/source/user/aberration/index.rst:27-45; documented by the reference, no code there) -------
 * paos_psf_peak: out_dev[0] = value of the PSF at the optical axis (pixel (n/2, n/2)), out_dev[1] = its maximum; the Strehl
 * ratio is the quotient of out_dev[0] for the aberrated and the ideal system.  psf_dev as for paos_encircled_energy.
 * paos_screen_stats: mean, variance sigma_W^2 and pixel count of a wavefront-error screen (n*n device doubles, metres) over
 * the pupil (x^2 + y^2)/radius^2 <= 1, x = (ix - n/2)*dx: the Marechal estimate is 1 - (2*pi/wl)^2 * sigma_W^2.
 * Both asynchronous on the handle's stream, results in device memory (2 / 3 doubles). */
int paos_psf_peak(paos_wfo *w, const void *psf_dev, double *out_dev);
int paos_screen_stats(paos_wfo *w, const double *screen_dev, double radius, double dx, double dy, double *out_dev);
/* Reduced host product of a sweep: copy the window [y0, y0+ny) x [x0, x0+nx) of an n*n real read-out (double for a
 * complex128 handle, float for complex64) into a dense nx*ny device array, narrowed to float when to_float != 0.  A
 * centred 512^2 float window of a 2048^2 PSF is 1 MiB instead of 32 MiB on the PCIe link.  Asynchronous. */
int paos_crop_convert(paos_wfo *w, const void *src_dev, int x0, int y0, int nx, int ny, int to_float, void *dst_dev);

/* ---- statistics -------------------------------------------------------------------------------- */
typedef struct paos_stats {
    uint64_t kernel_launches;   /* kernels of this library launched for the handle so far (a batched launch is booked
                                   once, on the first handle of the batch: the sum over handles is the true count) */
    uint64_t pass_launches;     /* of which FFT line-pass kernels */
    uint64_t passes_planned;    /* line passes (sweeps of the field) of this wavefront, however they were launched */
    uint64_t fft2_recorded;     /* FFT2s requested through the API (algorithmic count) */
    uint64_t line_ffts_run;     /* 1-D line-FFT batches executed (2 per FFT2) */
    uint64_t lines_transformed; /* single lines actually transformed (a batch is n lines unless an aperture blanks some) */
    uint64_t lines_tabled;      /* along-line phase-table multiplies, in lines (tables of a pass x lines it really processes) */
    uint64_t lines_swept;       /* lines loaded and stored by the passes (one per pass and live line) */
    uint64_t host_plan_us;      /* host microseconds spent inside paos_chain_run / paos_batch_chain_run: walking the surfaces,
                                   planning the passes and issuing the launches (booked on the first handle of a batch) */
    double last_flush_ms;       /* device time of the most recent flushed batch (CUDA events), 0 if untimed */
} paos_stats;
int paos_wfo_stats(paos_wfo *w, paos_stats *out);
/* enable CUDA-event timing of every pass kernel (adds two event records per launch); the per-kernel
 * totals are reported by paos_wfo_timing */
int paos_wfo_enable_timing(paos_wfo *w, int enable);
/* sum of pass-kernel device times (ms) and number of timed launches since the last call */
int paos_wfo_timing(paos_wfo *w, double *pass_ms, uint64_t *pass_launches);
/* totals since the last reset: device time of the timed pass launches, their number, the line-FFT sweeps they carried
 * (chained FFTs summed over the wavefronts of each launch: algorithmic bytes = sweeps * 2 * sizeof(complex) * N^2 in the
 * convention of SURVEY.md 8d) and the wavefront-passes they served */
int paos_wfo_timing_totals(paos_wfo *w, double *ms, uint64_t *launches, uint64_t *line_fft_sweeps, uint64_t *wavefronts,
                           int reset);
/* the same split by pass kind: col = 0 row pass / 1 column pass, nfft = line FFTs chained in the pass
 * (0..PAOS_MAX_CHAINED_FFTS); reset != 0 clears the bucket after reading it */
int paos_wfo_timing_detail(paos_wfo *w, int col, int nfft, double *ms, uint64_t *launches, int reset);

#ifdef __cplusplus
}
#endif
#endif /* PAOS_B200_H */
