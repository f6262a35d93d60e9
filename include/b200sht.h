/* b200sht.h -- C ABI of libb200sht.so: the B200-native spherical-harmonic-transform engine that
 * stands in for the ducc0 / cmisc / fft-engine calls on pixell's curvedsky hot path.
 *
 * Each entry point names the reference interface it replaces (paths relative to the pixell tree).
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is available from
 *     b2_last_error() (thread-local).  Nothing throws across this boundary.
 *   - all buffers are caller-owned.  `mem` says where they live: B2_MEM_HOST (pageable or pinned
 *     host memory; the library stages them through the device on its stream) or B2_MEM_DEVICE
 *     (device pointers, e.g. torch tensors; zero-copy).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Calls are
 *     asynchronous with respect to the host only for B2_MEM_DEVICE; host-memory calls return
 *     after the result is in the caller's buffer.
 *   - plans own their device scratch and tables and are not thread-safe (one plan per thread).
 *   - complex numbers are interleaved (re, im); `dtype` B2_F64 means float64/complex128 buffers,
 *     B2_F32 float32/complex64 (arithmetic is always fp64 on the device).
 */
#ifndef B200SHT_H
#define B200SHT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B2_MEM_HOST   0
#define B2_MEM_DEVICE 1
#define B2_F64 0
#define B2_F32 1
#define B2_MODE_STANDARD 0
#define B2_MODE_DERIV1   1   /* spin-1 gradient mode: alm[1,nalm] <-> map[2,npix] (ducc mode="DERIV1") */

typedef struct b2_sht_plan b2_sht_plan;
typedef struct b2_fft_plan b2_fft_plan;

/* ---- runtime -------------------------------------------------------------------------------- */
int         b2_init(int device);              /* select the CUDA device for this thread (cudaSetDevice) */
const char *b2_last_error(void);
int         b2_version(void);
int         b2_device_synchronize(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t     b2_launch_count(void);
/* measured peak FP64 FMA rate of the current device in GFLOP/s (microbenchmark; the roofline
 * denominator for the Legendre kernels) */
int         b2_dfma_peak_gflops(double *out);

/* ---- SHT plans --------------------------------------------------------------------------------
 * b2_sht_plan_rings replaces the ring description pixell hands to ducc0.sht.experimental.synthesis /
 * adjoint_synthesis (pixell/curvedsky.py:936-960, 1068-1084; built by get_ring_info :1170-1190):
 *   ring r has colatitude theta[r], nphi pixels per full circle (equal for all rings: CAR), the
 *   first stored pixel sits at azimuth phi0 and pixel j at phi0 + xdir*2*pi*j/nphi, xdir = +1 or -1;
 *   only the first npix_ring <= nphi pixels of each ring are stored (cut-sky rows), at element
 *   offset ringstart[r] + j inside one map component.  weight (nullable) multiplies ring r in
 *   adjoint_synthesis (quadrature / pixel-area weights, curvedsky.py:852-861).
 *   alm layout (curvedsky.alm_info, curvedsky.py:409-447): index(l,m) = mstart[m] + l*lstride.
 * Rings may come in any order; the plan pairs theta with pi-theta itself.  y/x flips of the
 * caller's array are expressed through ringstart and xdir, so no map2buffer/buffer2map copies
 * (curvedsky.py:1384-1411) are needed. */
int b2_sht_plan_rings(b2_sht_plan **out, int nring, const double *theta, int64_t nphi, double phi0,
                      int xdir, int64_t npix_ring, const int64_t *ringstart, const double *weight,
                      int lmax, int mmax, const int64_t *mstart, int64_t lstride);
/* Ring sets whose rings differ in nphi / phi0 (HEALPix: pixell/curvedsky.py:1192-1222 get_ring_info_healpix, used by
 * alm2map_healpix :312-353 and map2alm_healpix :355-405; single-pixel rings: get_ring_info_radial :1224-1234).
 * nphi[nring], phi0[nring], ringstart[nring]; pixel j of ring r is element ringstart[r] + j.  Rings are grouped by
 * nphi internally (one FFT table per distinct length, any length: mixed radix or Bluestein). */
int b2_sht_plan_rings_general(b2_sht_plan **out, int nring, const double *theta, const int64_t *nphi, const double *phi0,
                              const int64_t *ringstart, const double *weight /* nullable */,
                              int lmax, int mmax, const int64_t *mstart, int64_t lstride);

/* b2_sht_plan_2d replaces the geometry arguments of ducc0.sht.experimental.synthesis_2d /
 * adjoint_synthesis_2d / analysis_2d / adjoint_analysis_2d (curvedsky.py:907-924, 1032-1046):
 * geometry in {"CC","F1","MW","MWflip","DH","F2"}, map component = [ntheta][nphi] C-contiguous.
 * flip_y / flip_x say that the caller's array is stored south-first / with phi decreasing in x
 * (what pixell fixes by copying in map2buffer); phi0 is the azimuth of the caller's pixel x=0. */
int b2_sht_plan_2d(b2_sht_plan **out, const char *geometry, int ntheta, int64_t nphi, double phi0,
                   int flip_y, int flip_x, int lmax, int mmax, const int64_t *mstart, int64_t lstride);
void b2_sht_plan_destroy(b2_sht_plan *plan);
/* bytes of device memory held by the plan (tables + scratch) */
int64_t b2_sht_plan_bytes(const b2_sht_plan *plan);

/* ---- transforms --------------------------------------------------------------------------------
 * spin 0: alm[1] <-> map[1]; spin>0: alm[2] (E,B) <-> map[2] (Q,U); DERIV1: alm[1] <-> map[2].
 * Component c of alm starts at alm + c*alm_cstride (complex elements), of map at map + c*map_cstride
 * (real elements).  nbatch independent transforms are laid out with strides alm_bstride / map_bstride.
 *   b2_synthesis          ducc0 synthesis / synthesis_2d               (curvedsky.py:908, 937)
 *   b2_adjoint_synthesis  ducc0 adjoint_synthesis / adjoint_synthesis_2d (curvedsky.py:907, 936)
 *   b2_analysis_2d        ducc0 analysis_2d (exact quadrature; 2d plans only) (curvedsky.py:1033)
 *   b2_adjoint_analysis_2d ducc0 adjoint_analysis_2d                   (curvedsky.py:1032)
 * Entries of alm outside m<=mmax, max(m,spin)<=l<=lmax are left untouched by the *->alm calls,
 * except l<spin entries inside the triangle, which are set to zero. */
int b2_synthesis(b2_sht_plan *plan, int spin, int mode, int dtype, int nbatch,
                 const void *alm, int64_t alm_cstride, int64_t alm_bstride,
                 void *map, int64_t map_cstride, int64_t map_bstride, int mem, void *stream);
int b2_adjoint_synthesis(b2_sht_plan *plan, int spin, int mode, int dtype, int nbatch,
                 void *alm, int64_t alm_cstride, int64_t alm_bstride,
                 const void *map, int64_t map_cstride, int64_t map_bstride, int mem, void *stream);
int b2_analysis_2d(b2_sht_plan *plan, int spin, int dtype, int nbatch,
                 void *alm, int64_t alm_cstride, int64_t alm_bstride,
                 const void *map, int64_t map_cstride, int64_t map_bstride, int mem, void *stream);
int b2_adjoint_analysis_2d(b2_sht_plan *plan, int spin, int dtype, int nbatch,
                 const void *alm, int64_t alm_cstride, int64_t alm_bstride,
                 void *map, int64_t map_cstride, int64_t map_bstride, int mem, void *stream);
/* Several spin groups of one map (e.g. T with spin 0 and Q,U with spin 2: what pixell's curvedsky loops over at
 * curvedsky.py:922, 1064) in one call.  op: 0 synthesis, 1 adjoint_synthesis, 2 analysis_2d, 3 adjoint_analysis_2d.
 * Group g transforms alm[g] (component stride alm_cstride[g]) <-> map[g] (component stride map_cstride[g]) with
 * spin spins[g].  With B2_MEM_HOST the groups are pipelined over three streams: the host-to-device copies of
 * group g+1 and the device-to-host copies of group g-1 overlap the kernels of group g (pinned host memory is
 * needed for the overlap, not for correctness).  At most 8 groups. */
int b2_sht_execute_groups(b2_sht_plan *plan, int op, int ngroups, const int *spins, int mode, int dtype,
                 void *const *alm, const int64_t *alm_cstride, void *const *map, const int64_t *map_cstride,
                 int mem, void *stream);
/* per-stage device times (ms) of the last transform executed on this plan:
 * out[0] = Legendre, out[1] = ring FFT, out[2] = theta resampling, out[3] = staging copies */
int b2_sht_last_timing(b2_sht_plan *plan, double out[4]);
/* test hooks: run only the Legendre stage on device-resident leg[ncomp][mmax+1][nring] (ring order = plan order) */
int b2_alm2leg(b2_sht_plan *plan, int spin, int mode, const void *alm_dev, int64_t alm_cstride, void *leg_dev, void *stream);
int b2_leg2alm(b2_sht_plan *plan, int spin, int mode, void *alm_dev, int64_t alm_cstride, const void *leg_dev, void *stream);
/* test hook: the theta-weighting operator of analysis_2d (adjoint = 0) or its conjugate transpose (adjoint = 1) alone,
 * in place on device-resident leg[ncomp][mmax+1][nring_pad] (nring_pad = rings rounded up to a multiple of 32) */
int b2_theta_weighting(b2_sht_plan *plan, int spin, int ncomp, int adjoint, void *leg_dev, void *stream);
/* tuning hook: choose the kernel variant (launch shape) of Legendre kernel `which` (0 synth spin 0, 1 adjoint spin 0,
 * 2 synth spin>0, 3 adjoint spin>0); results are identical across variants */
int b2_set_leg_variant(int which, int variant);

/* ducc0.sht.experimental.get_gridweights(name, ntheta) (curvedsky.py:501, 531, 855): ring
 * quadrature weights, sum = 4*pi.  Host computation into out[ntheta]. */
int b2_gridweights(const char *geometry, int ntheta, double *out);

/* ---- alm helpers: replace pixell.cmisc / cython/cmisc_core.c ------------------------------------
 *   b2_alm2cl        cmisc_core.c:16-110  (alm2cl_sp / _sp_to_dp / _dp); cl_dtype may differ from dtype
 *   b2_lmul          cmisc_core.c:159-182 (lmul_dp / lmul_sp), in place
 *   b2_lmatmul       cmisc_core.c:185-274 (lmatmul_dp / _sp): oalm[r] = sum_c lmat[r][c][l] alm[c]; in-place safe
 *   b2_transpose_alm cmisc_core.c:116-156
 *   b2_transfer_alm  cmisc.pyx:131-150 / cmisc_core.c:278-304
 */
int b2_alm2cl(int lmax, int mmax, const int64_t *mstart, int dtype, const void *alm1, const void *alm2,
              int cl_dtype, void *cl, int mem, void *stream);
int b2_lmul(int lmax, int mmax, const int64_t *mstart, int dtype, void *alm, int lfmax, const void *lfun,
            int mem, void *stream);
int b2_lmatmul(int N, int M, int lmax, int mmax, const int64_t *mstart, int dtype,
               const void *alm, int64_t alm_cstride, int lfmax, const void *lmat /* [N][M][lfmax+1] */,
               void *oalm, int64_t oalm_cstride, int mem, void *stream);
int b2_transpose_alm(int lmax, int mmax, const int64_t *mstart, int dtype, const void *ialm, void *oalm,
                     int mem, void *stream);
int b2_transfer_alm(int lmax1, int mmax1, const int64_t *mstart1, int64_t lstride1, const void *alm1,
                    int lmax2, int mmax2, const int64_t *mstart2, int64_t lstride2, void *alm2,
                    int dtype, int mem, void *stream);

/* b2_rand_alm: curvedsky.rand_alm (pixell/curvedsky.py:61-77) on the device: white unit normals in the reference's fill
 * order (rand_alm_white / fill_gauss, :602-628: memory order of the l-major array, component after component) from a
 * counter-based Philox4x32-10 stream keyed by `seed`, coloured with ps12[r][c][l]/sqrt(2) (ps12 = multi_pow(ps, 0.5),
 * [ncomp][ncomp][lmax+1]; NULL: white alm as rand_alm_white returns them), m = 0 made real with the sqrt(2) restored.
 * Component r of the result starts at alm + r*alm_cstride (complex elements).  The numbers differ from numpy's MT19937
 * stream (which pixell_b200.curvedsky.rand_alm reproduces on the host); the statistics and the fill order are the same. */
int b2_rand_alm(int lmax, int mmax, const int64_t *mstart, int ncomp, uint64_t seed, const double *ps12,
                int dtype, void *alm, int64_t alm_cstride, int mem, void *stream);

/* ---- FFT engine: replaces the engines[...] .FFTW plan objects of pixell/fft.py:8-113 -------------
 * Batched multi-dimensional DFT over up to 2 axes of a strided array (what enmap.fft / fft.rfft /
 * fft.irfft need, pixell/fft.py:133-209, enmap.py:1307-1337).
 *   kind: B2_FFT_C2C, B2_FFT_R2C (last transformed axis halved+1 in the output), B2_FFT_C2R
 *   shape/istride/ostride describe the full ndim-dimensional arrays (strides in elements of the
 *   respective dtype); axes lists the transformed axes (last listed = the r2c/c2r axis).
 *   Backward transforms are unnormalised, like FFTW (fft.py:180); `scale` multiplies the output. */
#define B2_FFT_C2C 0
#define B2_FFT_R2C 1
#define B2_FFT_C2R 2
int b2_fft_plan_create(b2_fft_plan **out, int ndim, const int64_t *shape, const int64_t *istride,
                       const int64_t *ostride, int naxes, const int *axes, int kind, int dtype);
int b2_fft_execute(b2_fft_plan *plan, const void *in, void *out, int forward, double scale,
                   int mem, void *stream);
void b2_fft_plan_destroy(b2_fft_plan *plan);

/* QU <-> EB rotation of flat-sky Fourier maps: enmap.queb_rotmat + enmap.map_mul (pixell/enmap.py:1391-1400,
 * 1418-1427) as used by map2harm / harm2map (:1358-1383).  Rotates, in place, nbatch pairs of complex [ny][nx]
 * components (pair k: data + k*batch_stride and data + k*batch_stride + comp_stride, strides in complex elements)
 * by the angle spin*atan2(sign*lx[x], ly[y]); ly[ny], lx[nx] are HOST arrays (enmap.laxes).  sign = +1: map2harm
 * (HEALPix convention), -1: harm2map or iau=True (the product of the two flips, enmap.py:1394-1396). */
int b2_queb_rotate(void *data, int64_t comp_stride, int64_t nbatch, int64_t batch_stride, int ny, int nx,
                   const double *ly, const double *lx, int spin, int sign, int dtype, int mem, void *stream);

/* Fourier-space filter of complex [nbatch][ny][nx] maps, in place, DEVICE pointers: data[b][y][x] *= fy[y]*fx[x] (both given)
 * or *= f2[y][x] (f2 given): what enmap.smooth_gauss / apply_window do between enmap.fft and enmap.ifft with numpy
 * broadcasting (pixell/enmap.py:1429-1460), as one pass over the array.  Strides in complex elements; the filters have the
 * real dtype of the data. */
int b2_fourier_filter(void *data, int64_t nbatch, int64_t batch_stride, int64_t row_stride, int ny, int nx,
                      const void *fy, const void *fx, const void *f2, int dtype, void *stream);

/* ---- synthesis at arbitrary positions: the steps of ducc0.sht.experimental.synthesis_general (call site
 * pixell/curvedsky.py:993-1016) around the Legendre stage (b2_alm2leg on a Clenshaw-Curtis plan) and the FFT
 * engine (b2_fft_*).  All pointers are DEVICE pointers, complex128 / float64.
 *   b2_general_extend   leg[ncomp][nm][nring_pad] on nt CC rings -> ext[ncomp][nm][N], N = 2 (nt - 1): the
 *                       (-1)^(m+spin)-symmetric continuation to the full theta circle
 *   b2_general_scatter  coef[ncomp][nm][N] (theta-FFT of ext, scaled 1/N) -> grid[ncomp][M][M/2+1]: modes |k| <= lmax,
 *                       m < nm, each times corr[|k|] corr[m] (kernel deconvolution, corr on the device, lmax+1
 *                       entries), zero elsewhere, m = 0 column Hermitian
 *   b2_general_interp   fine[ncomp][M][M] (inverse FFT of grid) -> out[c*out_comp_stride + i], i < npos, at
 *                       loc[i] = (theta, phi) radians: W x W "exponential of semicircle" interpolation */
int b2_general_extend(const void *leg_dev, void *ext_dev, int ncomp, int nm, int nt, int64_t nring_pad, int spin, void *stream);
int b2_general_scatter(const void *coef_dev, void *grid_dev, int ncomp, int lmax, int nm, int N, int M,
                       const double *corr_dev, void *stream);
int b2_general_interp(const void *fine_dev, int ncomp, int M, const double *loc_dev, int64_t npos, int W, double beta,
                      void *out_dev, int64_t out_comp_stride, void *stream);
/* the transposed steps (ducc0.sht.experimental.adjoint_synthesis_general, same call site with adjoint=True):
 *   b2_general_spread  val[c*val_comp_stride + i] -> fine[ncomp][M][M] (zeroed first; atomic adds)
 *   b2_general_gather  grid[ncomp][M][M/2+1] (forward r2c FFT of fine) -> coef[ncomp][nm][N], modes |k| <= lmax
 *                      times corr[|k|] corr[m], zero elsewhere
 *   b2_general_fold    ext[ncomp][nm][N] (unnormalised inverse theta-FFT of coef, scaled 1/N) -> leg[ncomp][nm][nring_pad]
 *                      for b2_leg2alm on the same Clenshaw-Curtis plan */
int b2_general_spread(void *fine_dev, int ncomp, int M, const double *loc_dev, int64_t npos, int W, double beta,
                      const void *val_dev, int64_t val_comp_stride, void *stream);
int b2_general_gather(void *coef_dev, const void *grid_dev, int ncomp, int lmax, int nm, int N, int M,
                      const double *corr_dev, void *stream);
int b2_general_fold(void *leg_dev, const void *ext_dev, int ncomp, int nm, int nt, int64_t nring_pad, int spin, void *stream);

#ifdef __cplusplus
}
#endif
#endif
