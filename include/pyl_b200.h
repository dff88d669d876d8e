/* pyl_b200.h -- C ABI of the B200-native density-field / power-spectrum hot path.
 *
 * This is the drop-in boundary for the path
 *     MAS_library.MA -> FFT3Dr_f -> Pk_library.Pk / XPk
 * of Pylians3.  Every entry point is `extern "C"`, takes plain pointers and sizes, returns
 * an int status (0 = PYL_OK, negative = error; never throws, never exits) and is ordered
 * on the CUDA stream it is given.  No torch types appear here.
 *
 * Reference interfaces replaced (paths relative to the Pylians3 tree):
 *   library/MAS_library/MAS_c.h:3-10          void NGP|CIC|TSC|PCS(FLOAT *pos, FLOAT *number,
 *                                             FLOAT *W, long particles, int dims, int axes,
 *                                             FLOAT BoxSize, int threads)
 *   library/MAS_library/MAS_library.pyx:72-80 MA()'s dispatch to NGP/CIC/TSC/PCS(+W)
 *   library/MAS_library/MAS_library.pyx:558-599 CIC_interp(density, BoxSize, pos, den)
 *   library/redshift_space_library/redshift_space_library.pyx:29-46 pos_redshift_space(...)
 *   library/Pk_library/Pk_library.pyx:117-130 FFT3Dr_f(a, threads)
 *   library/Pk_library/Pk_library.pyx:311-378 Pk.__init__ hot loop   (no native ABI exists)
 *   library/Pk_library/Pk_library.pyx:623-732 XPk.__init__ hot loop  (no native ABI exists)
 *   library/Pk_library/Pk_library.pyx:149-163 IFFT3Dr_f(a, threads)
 *   library/Pk_library/Pk_library.pyx:470-499, 905-1016, 1151-1200, 1273-1316, 1386-1432, 1515-1568,
 *       1909-1929, 2004-2037, 2198-2267, 2335-2412   the loops of Pk_plane, XPk_imag, XPk_plane, Pk_theta,
 *       XPk_dv, XPk_vv, correct_MAS, expected_Pk, Xi, XXi (section 5; no native ABI exists)
 *   library/smoothing_library/smoothing_library.pyx:227-232 field_smoothing's mode loop; :37-114, :141-203 the
 *       filter loops of FT_filter / FT_filter_2D; :255-258 field_smoothing_2D's loop
 *   library/Pk_library/Pk_library.pyx:213-226 IFFT2Dr_f(a, threads)
 *
 * Ownership: all device pointers are BORROWED.  The device-pointer entry points never
 * allocate: scratch memory is passed in (`ws`, sized by the matching *_workspace_bytes
 * query).  Only the host-pointer convenience entry points (section 4) own a grow-only
 * device arena.
 */
#ifndef PYL_B200_H
#define PYL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cudaStream_t without dragging cuda_runtime.h into C / Cython / cgo consumers */
typedef struct CUstream_st *pyl_stream_t;

/* ---- status codes ------------------------------------------------------------------ */
#define PYL_OK 0
#define PYL_ERR_ARG (-1)       /* bad argument (NULL pointer, unknown scheme, bad axes ...) */
#define PYL_ERR_CUDA (-2)      /* a CUDA runtime call or kernel launch failed               */
#define PYL_ERR_CUFFT (-3)     /* a cuFFT call failed                                       */
#define PYL_ERR_WORKSPACE (-4) /* workspace missing or smaller than *_workspace_bytes()     */
#define PYL_ERR_DOMAIN (-5)    /* slab deposit: a particle's stencil left the local planes  */

const char *pyl_error_string(int status);
/* detail of the last failure on the calling thread ("" if none) */
const char *pyl_last_error(void);
/* "pyl_b200 <version> sm_100a" */
const char *pyl_version(void);
/* number of kernels of THIS library launched by the process so far (cuFFT's are not counted) */
unsigned long long pyl_kernel_launches(void);

/* ---- 1. mass assignment ------------------------------------------------------------ */
/* scheme ids; the MAS window exponent of Pk_library.pyx:72-78 is (id + 1) */
#define PYL_MAS_NGP 0
#define PYL_MAS_CIC 1
#define PYL_MAS_TSC 2
#define PYL_MAS_PCS 3

/* deposit algorithms */
#define PYL_MODE_AUTO (-1)         /* pick by particle density and scheme                  */
#define PYL_MODE_ATOMIC 0          /* one red.global.add.f32 per stencil cell              */
#define PYL_MODE_TILED 1           /* bucket by tile, accumulate in shared memory, flush   */
#define PYL_MODE_DETERMINISTIC 2   /* sort by cell, fixed-order segmented sums: bit-reproducible */

/* Scratch bytes pyl_deposit needs for this problem (0 for PYL_MODE_ATOMIC). */
size_t pyl_deposit_workspace_bytes(int mas, int64_t particles, int dims, int axes, int mode);

/* number[dims^axes] (float32, C order, DEVICE) += deposit of `particles` particles.
 *   pos : DEVICE float32 [particles][axes], C-contiguous, 0 <= pos <= BoxSize
 *   W   : DEVICE float32 [particles] or NULL
 *   axes: 3, or 2 (plane; like the reference each particle then lands 1/2/3/4 times, i.e.
 *         the plane is 1x/2x/3x/4x the 2D deposit -- MAS_library.pyx:84-110 divides after)
 * Arithmetic follows MAS_library.pyx:123-166,273-292,369-404,463-497 (+W variants): the
 * cell coordinate is the ROUNDED float32 product pos*(dims/BoxSize).
 * Replaces MAS_c.h:3-10 / MAS_library.pyx:72-80. */
int pyl_deposit(int mas, const float *pos, float *number, const float *W, int64_t particles,
                int dims, int axes, float BoxSize, int mode, void *ws, size_t ws_bytes,
                pyl_stream_t stream);

/* Slab variant for one rank of a multi-GPU run (x-slab decomposition).
 * `number` holds `x_planes` consecutive x-planes of the global grid starting at global plane
 * `x_origin` (periodic: plane indices are taken modulo dims): the rank's `x_own` own planes followed
 * by x_planes - x_own upward ghost planes.  Particles must have been routed by
 * pyl_stencil_base_plane (first stencil plane among the own planes).  A stencil cell whose plane
 * falls outside the window is NOT deposited and is counted in *dropped (DEVICE int64, may be NULL;
 * the caller checks that it stays 0).  3D only.  With a workspace of
 * pyl_deposit_slab_workspace_bytes() bytes the tiled kernel is used, otherwise (ws NULL) the atomic. */
size_t pyl_deposit_slab_workspace_bytes(int mas, int64_t particles, int dims, int x_own);
int pyl_deposit_slab(int mas, const float *pos, float *number, const float *W,
                     int64_t particles, int dims, float BoxSize, int x_origin, int x_own, int x_planes,
                     int64_t *dropped, void *ws, size_t ws_bytes, pyl_stream_t stream);

/* plane[i] = wrapped x index of the FIRST stencil cell of particle i (NGP: the cell itself; CIC:
 * floor(dist); TSC: floor(dist-1.5)+1; PCS: floor(dist-2)+1), computed with the deposit's own
 * arithmetic.  The multi-GPU path routes a particle to the rank owning that plane; its stencil then
 * covers planes plane .. plane+S-1 (S = 1,2,3,4), i.e. only upward ghost planes are needed.
 * pos: DEVICE float32 [particles][3]; plane: DEVICE int32 [particles]. */
int pyl_stencil_base_plane(int mas, const float *pos, int64_t particles, int dims, float BoxSize,
                           int32_t *plane, pyl_stream_t stream);

/* Grid -> particle CIC interpolation: den[i] = sum of the 8 cells of `density` (dims^3 float32, DEVICE) around
 * pos[i] times the CIC weights; `den` (DEVICE float32 [particles]) is overwritten.
 * Replaces MAS_library.pyx:558-599 (CIC_interp). */
int pyl_cic_interp(const float *density, int dims, float BoxSize, const float *pos, int64_t particles,
                   float *den, pyl_stream_t stream);

/* Real -> redshift space IN PLACE: pos[i][axis] += vel[i][axis]*(1+redshift)/Hubble, wrapped into [0,BoxSize].
 * pos, vel: DEVICE float32 [particles][3].  Replaces redshift_space_library.pyx:29-46 (pos_redshift_space). */
int pyl_pos_redshift_space(float *pos, const float *vel, int64_t particles, float BoxSize, float Hubble,
                           float redshift, int axis, pyl_stream_t stream);

/* x[i] /= divisor, IEEE float32 division: the 2D renormalisation `number2 /= 2.0|3.0|4.0`
 * of MAS_library.pyx:90-107 */
int pyl_divide_inplace(float *x, int64_t n, float divisor, pyl_stream_t stream);
/* x[i] *= factor */
int pyl_scale_inplace(float *x, int64_t n, float factor, pyl_stream_t stream);
/* x[i] = x[i]*a + b */
int pyl_affine_inplace(float *x, int64_t n, float a, float b, pyl_stream_t stream);
/* out[0] = sum_i x[i] accumulated in float64 (DEVICE double; zeroed by the call) */
int pyl_sum_f64(const float *x, int64_t n, double *out, pyl_stream_t stream);
/* x[i] = float(double(x[i]) / (sum[0]/count)) - 1 : delta = n/<n> - 1 exactly as the reference's
 * callers compute it in NumPy (docs/source/construction.rst:50, Pk_snapshot.py:88).  `sum` is a
 * DEVICE double (e.g. from pyl_sum_f64, all-reduced across ranks), so no host sync is needed. */
int pyl_overdensity_inplace(float *x, int64_t n, const double *sum, double count, pyl_stream_t stream);
/* x[i] = -c with c = float32(numerator[0] / cells), numerator a DEVICE double; c_out[0] (DEVICE double) = c.
 * The start value of a deposit whose spectrum is taken with pyl_pk_density_scale(offset = c_out): the grid then
 * holds n - c and the float32 transform does not carry the whole mass in its DC mode. */
int pyl_fill_negative(float *x, int64_t n, const double *numerator, double cells, double *c_out,
                      pyl_stream_t stream);
/* out[i] += in[i] (ghost-plane merge after the halo exchange) */
int pyl_add_inplace(float *out, const float *in, int64_t n, pyl_stream_t stream);

/* ---- 2. FFT stage (cuFFT; replaces FFT3Dr_f, Pk_library.pyx:117-130) ---------------- */
/* Unnormalised forward r2c, (dims,dims,dims) float32 -> (dims,dims,dims/2+1) complex64,
 * C order, out of place; `delta` is not modified.  Plans are cached per (dims, device). */
size_t pyl_fft_r2c_workspace_bytes(int dims);
int pyl_fft_r2c(const float *delta, float *delta_k /* interleaved re,im */, int dims, void *ws,
                size_t ws_bytes, pyl_stream_t stream);

/* Pieces of the slab-decomposed distributed transform:
 *   stage 1: `nx` local x-planes, batched 2D r2c over (y,z): (nx,dims,dims) -> (nx,dims,nz)
 *   stage 2: after the all-to-all transpose, 1D c2c along x IN PLACE on (dims, nky, nz)   */
size_t pyl_fft_slab_workspace_bytes(int dims, int nx, int nky);
int pyl_fft_slab_yz(const float *slab, float *slab_k, int dims, int nx, void *ws, size_t ws_bytes,
                    pyl_stream_t stream);
int pyl_fft_slab_x(float *cols_k, int dims, int nky, void *ws, size_t ws_bytes,
                   pyl_stream_t stream);
/* Transpose of the distributed transform as one kernel over peer memory (multi-GPU, one node): every row
 * (ix, ky, :) of the local (nx, dims, dims/2+1) complex64 output of stage 1 is stored directly into the receive
 * buffer of the rank owning ky -- peer_recv[r] (HOST array of `nranks` DEVICE pointers, peer-mapped; shape
 * (dims, nky_of_rank[r], dims/2+1)) at [x0+ix][ky_row[ky]][:].  ky_owner / ky_row: DEVICE int32 [dims];
 * nky_of_rank: HOST int [nranks].  ky_order: DEVICE int32 [dims] or NULL -- the order in which the rows are
 * issued (a permutation of 0..dims-1); interleave the owners, starting with a different one on every rank, so
 * that the ranks do not all write to the same receiver at the same time.  The caller synchronises the ranks
 * before (buffers free) and after (rows landed).  Replaces "pack + NCCL all-to-all". */
int pyl_transpose_scatter(const float *slab_k, void *const *peer_recv, const int *nky_of_rank,
                          const int *ky_owner, const int *ky_row, const int *ky_order, int dims, int nx, int x0,
                          int nranks, pyl_stream_t stream);
/* The same transpose into receive buffers laid out (nky[r], dims, dims/2+1), and the 1D transforms along x of such a
 * buffer (one strided cuFFT call per ky plane).  With x in the middle the x transforms stay inside one
 * (dims, dims/2+1) plane per ky; with x outermost they touch a different 2 MB page per element at 4096^3. */
int pyl_transpose_scatter_kymajor(const float *slab_k, void *const *peer_recv, const int *nky_of_rank,
                                  const int *ky_owner, const int *ky_row, const int *ky_order, int dims, int nx,
                                  int x0, int nranks, pyl_stream_t stream);
size_t pyl_fft_slab_x_kymajor_workspace_bytes(int dims);
int pyl_fft_slab_x_kymajor(float *cols_k, int dims, int nky, void *ws, size_t ws_bytes, pyl_stream_t stream);
/* Unnormalised inverse c2r, (dims,dims,dims/2+1) complex64 -> (dims,dims,dims) float32, out of place; cuFFT may
 * OVERWRITE delta_k.  The reference's IFFT3Dr_f (Pk_library.pyx:149-163) returns the NORMALISED inverse (pyfftw
 * scales inverse transforms by 1/N^3 by default): callers follow with pyl_scale_inplace(delta, N^3, 1/N^3) or
 * pass that factor to the consumer (pyl_shell_bin's `scale`). */
size_t pyl_fft_c2r_workspace_bytes(int dims);
int pyl_fft_c2r(float *delta_k /* interleaved re,im */, float *delta, int dims, void *ws, size_t ws_bytes,
                pyl_stream_t stream);
/* Same for one (dims,dims/2+1) complex64 image -> (dims,dims) float32 (IFFT2Dr_f, Pk_library.pyx:213-226). */
size_t pyl_fft2d_c2r_workspace_bytes(int dims);
int pyl_fft2d_c2r(float *image_k, float *image, int dims, void *ws, size_t ws_bytes, pyl_stream_t stream);
/* release every cached cuFFT plan of the calling thread's current device */
int pyl_fft_clear_plans(void);

/* ---- 3. |delta_k|^2 binning (replaces Pk_library.pyx:311-378 and :623-732) ---------- */
/* Offsets (in 8-byte words) of each accumulator inside the single output buffer.
 * F fields, X = F(F-1)/2 pairs (i<j lexicographic).  Nm* are uint64 counts, the rest are
 * float64 RAW sums (no units, DC bin included), exactly the accumulators of the reference
 * loop before Pk_library.pyx:384 / :735:
 *   k3D[kmax+1]  Nm3D[kmax+1]  Pk3D[kmax+1][3][F]  PkX3D[kmax+1][3][X]  phase[kmax+1]
 *   Nm1D[kmax_par+1]  Pk1D[kmax_par+1][F]  PkX1D[kmax_par+1][X]       (k1D = k_par*Nm1D)
 *   Nm2D[n2d]  Pk2D[n2d][F]  PkX2D[n2d][X],   n2d = (kmax_par+1)*(kmax_per+1)           */
typedef struct pyl_pk_layout {
    int32_t dims, fields, xfields, kmax_par, kmax_per, kmax;
    int64_t n2d;
    int64_t k3D, Nm3D, Pk3D, PkX3D, phase;
    int64_t Nm1D, Pk1D, PkX1D;
    int64_t Nm2D, Pk2D, PkX2D;
    int64_t total_words; /* size of the output buffer in 8-byte words */
} pyl_pk_layout_t;

/* frequencies() of Pk_library.pyx:56-61 plus the buffer layout */
int pyl_pk_layout(int dims, int fields, pyl_pk_layout_t *layout);

size_t pyl_pk_bin_workspace_bytes(int dims, int fields);

/* Bin `fields` (1..PYL_MAX_FIELDS) half-spectra in ONE pass.
 *   delta_k  : HOST array of `fields` DEVICE pointers, each (dims, nky, dims/2+1) complex64
 *              holding global ky rows [ky_lo, ky_lo+nky) of the r2c output (single GPU:
 *              ky_lo = 0, nky = dims).  Not modified (the reference's in-place
 *              deconvolution, :351-352, is an internal temporary).
 *   mas_index: HOST array, per field 0..4 = None,NGP,CIC,TSC,PCS (Pk_library.pyx:72-78)
 *   axis     : line of sight 0|1|2
 *   want_phase: flag word.  PYL_PK_PHASE: accumulate Pkphase of field 0 (Pk only; Pk_library.pyx:358,377).
 *              PYL_PK_CROSS_IMAG: the cross terms are imag_i*real_j - real_i*imag_j (XPk_imag,
 *              Pk_library.pyx:1000-1001) instead of real_i*real_j + imag_i*imag_j (XPk, :716-717)
 *   out      : DEVICE buffer of layout.total_words 8-byte words; zeroed by the call
 * Applies the Hermitian-duplicate skip rule (:324-327), the float32 window factor (:351)
 * and the float64 accumulation of the reference. */
#define PYL_MAX_FIELDS 4
#define PYL_PK_PHASE 1
#define PYL_PK_CROSS_IMAG 2
/* PYL_PK_KY_MAJOR: the fields are laid out (nky, dims, dims/2+1) -- stored ky row outermost, x in the middle -- as
 * pyl_transpose_scatter_kymajor / pyl_fft_slab_x_kymajor produce them (slab-distributed spectra only). */
#define PYL_PK_KY_MAJOR 4
int pyl_pk_bin(const float *const *delta_k, int fields, const int *mas_index, int dims,
               int ky_lo, int nky, int axis, int want_phase, void *out, void *ws,
               size_t ws_bytes, pyl_stream_t stream);

/* pyl_deposit_slab for particles whose COUNT lives on the device (the output of pyl_route_scatter): `capacity` is the
 * size of the pos/W arrays, *count (device uint32) the number of particles in them.  No reference counterpart. */
int pyl_deposit_slab_counted(int mas, const float *pos, float *number, const float *W, int64_t capacity,
                             const uint32_t *count, int dims, float BoxSize, int x_origin, int x_own, int x_planes,
                             int64_t *dropped, void *ws, size_t ws_bytes, pyl_stream_t stream);

/* Particle routing to the x-slab owners over peer memory (no reference counterpart: the reference is one process).
 * Every rank owns the planes [x_offsets[r], x_offsets[r+1]) and a receive block of pyl_route_block_bytes(capacity)
 * bytes, allocated symmetrically and mapped into every peer (peer_blocks[r] = this process's pointer to rank r's
 * block).  Block layout: byte 0: uint32 cursor = particles received; byte 256: positions, (capacity, 3) float32;
 * then, 256-byte aligned, `capacity` float32 weights.  The kernel sends each of the caller's particles to the
 * owner of the x-plane of its first stencil cell (pyl_stencil_base_plane's rule): one system-scope atomic per
 * (4096-particle chunk, destination) reserves a run behind the destination's cursor, consecutive lanes store the
 * run into the destination's arrays.  The caller zeroes its cursor, synchronises the ranks, launches, synchronises
 * again, and deposits [block + 256] with pyl_deposit_slab_counted(count = block).  *lost (device int64) counts
 * particles that found a block full; it must stay 0. */
size_t pyl_route_block_bytes(int64_t capacity);
int pyl_route_scatter(int mas, const float *pos, const float *W, int64_t particles, int dims, float BoxSize,
                      int nranks, const int *x_offsets, void *const *peer_blocks, int64_t capacity, int64_t *lost,
                      pyl_stream_t stream);

/* pyl_pk_bin keeps, per (device, dims, axis, ky window), a device copy of the field-independent sections of its
 * result (mode counts, sum of |k| per shell) after the first call and skips their accumulation afterwards -- the
 * counterpart of a cached FFT plan.  This frees the copies of the current device. */
int pyl_pk_clear_cache(void);

/* Slab-distributed spectra (no reference counterpart: the reference is one process): converts the
 * uint64 mode counts of an accumulator block (layout of pyl_pk_layout) to float64 in place -- exact below 2^53 -- so
 * that a single float64 SUM all-reduce covers the whole block; follow with pyl_pk_finalize(counts_are_f64 = 1). */
int pyl_pk_counts_to_f64(void *acc, int dims, int fields, pyl_stream_t stream);

/* Finalisation of the accumulators IN PLACE (Pk_library.pyx:384-418 / :735-791): afterwards every word of `acc`
 * is the float64 value the reference stores for that slot -- k3D = <k> kF, Pk3D = P_l (2l+1)/Nmodes (BoxSize/
 * dims^2)^3, phase likewise, Pk1D with its perpendicular-area weight, Pk2D = P/Nmodes2D * units; counts are
 * converted to float64 (counts_are_f64 != 0: they already are, e.g. after a float64 all-reduce).  The DC bins
 * keep their slots (the reference drops bin 0 of the 3D and 1D arrays; the caller slices).  Operation order
 * follows the reference's expressions, so results equal a host finalisation bit for bit.
 * kpar/kper (DEVICE double [n2d] each, or both NULL) receive the bin-centre coordinates of the 2D array. */
int pyl_pk_finalize(void *acc, int dims, int fields, double BoxSize, int counts_are_f64, double *kpar,
                    double *kper, pyl_stream_t stream);

/* Spectra of delta = n/<n> - 1 from the transform of the DENSITY n (fuses Pk_snapshot.py:88-89, the caller's
 * `delta /= mean; delta -= 1`, into the spectrum): FFT(delta) = FFT(n)/<n> away from k = 0, and <n> =
 * Re FFT(n)[0] / dims^3.  pyl_pk_take_dc stores the DC mode of every field in dc[fields] (DEVICE float64) and
 * zeroes it in the spectrum (holds_dc = 0: this rank's rows do not contain k = 0; dc is set to 0 so that a SUM
 * all-reduce delivers it everywhere).  pyl_pk_density_scale multiplies the RAW sums of pyl_pk_bin (call it
 * before pyl_pk_finalize) by dims^6 / (dc_i dc_j): autos by 1/<n>_i^2, crosses (pairs i<j in lexicographic
 * order) by 1/(<n>_i <n>_j); counts, k sums and the phase sums are untouched.
 * offset (DEVICE float64 [fields], or NULL = zeros): the grids held n_i - offset_i, i.e. the deposit started from
 * -offset_i instead of 0; <n>_i = offset_i + dc_i / dims^3.  Any constant near <n> keeps the DC mode, and with
 * it the float32 rounding noise a large DC mode leaves on the axes through k = 0, small. */
int pyl_pk_take_dc(float *const *delta_k, int fields, int holds_dc, double *dc, pyl_stream_t stream);
int pyl_pk_density_scale(void *acc, int dims, int fields, const double *dc, const double *offset,
                         pyl_stream_t stream);

/* Mirrored slab form (multi-GPU): the rank holds the rows |ky| in [ky_lo, ky_lo+ny_lo) of the half-range
 * 0..dims/2 AND their mirrors N-ky, as (dims, nky, dims/2+1) complex64 with the ky axis ordered: first the
 * ny_lo lower rows (ascending ky), then the mirrors that exist as separate modes (ky != 0, ky != Nyquist) in
 * ascending order of their global index.  pyl_pk_mirrored_rows returns nky (and the global index of the first
 * upper row).  Holding both signs of ky lets the kernel share geometry between the four modes (+-kx, +-ky, kz)
 * exactly as on a whole grid, and walk along an axis that is not the line of sight. */
int pyl_pk_mirrored_rows(int dims, int ky_lo, int ny_lo, int *first_upper);
int pyl_pk_bin_mirrored(const float *const *delta_k, int fields, const int *mas_index, int dims,
                        int ky_lo, int ny_lo, int axis, int want_phase, void *out, void *ws,
                        size_t ws_bytes, pyl_stream_t stream);

/* ---- 5. sibling estimators of Pk_library on |k| shells, and half-spectrum passes ------- */
/* One kernel family reduces every independent mode (or every real-space cell) into bins of one fundamental
 * frequency, k_index = (int)|k|, with a per-mode functional selected by `kind`:
 *   kind            replaces (Pk_library.pyx)  fields (DEVICE, in this order)              values per bin
 *   PYL_SHELL_THETA    Pk_theta    :1273-1316  Vx_k, Vy_k, Vz_k                             |i k.V|^2
 *   PYL_SHELL_DV       XPk_dv      :1386-1432  delta_k, Vx_k, Vy_k, Vz_k                    |d|^2, |k.V|^2, cross
 *   PYL_SHELL_VV       XPk_vv      :1515-1568  Vx1_k, Vy1_k, Vz1_k, Vx2_k, Vy2_k, Vz2_k     |k.V1|^2, |k.V2|^2, cross
 *   PYL_SHELL_EXPECTED expected_Pk :2004-2037  none (interpolation table)                   P_interp(k)
 *   PYL_SHELL_PLANE    Pk_plane    :470-499    one (dims, dims/2+1) image transform         |d|^2
 *   PYL_SHELL_XPLANE   XPk_plane   :1151-1200  two image transforms (mas_index[0], [1])     |d1|^2, |d2|^2, cross
 *   PYL_SHELL_XI       Xi / XXi    :2233-2267  one REAL (dims,dims,dims) float32 grid       xi*L_0, xi*L_2, xi*L_4
 * Complex fields are (dims,dims,dims/2+1) complex64 half-spectra (not modified); k.V is formed in float32 and
 * the window factor is a float32 product exactly as in the reference; all sums are float64.
 * `out` (DEVICE, zeroed by the call) receives (2 + values) * bins 8-byte words:
 *   sum_k[bins] (float64)   Nmodes[bins] (uint64)   value_j[bins] (float64), j = 0..values-1
 * RAW sums, DC bin included; units and averaging stay with the caller (they are O(bins)).
 * `scale` multiplies the grid values of PYL_SHELL_XI (float32 product: the 1/dims^3 of the normalised inverse
 * transform); ignored otherwise.  `axis` is the line of sight of the PYL_SHELL_XI multipoles. */
#define PYL_SHELL_THETA 0
#define PYL_SHELL_DV 1
#define PYL_SHELL_VV 2
#define PYL_SHELL_EXPECTED 3
#define PYL_SHELL_PLANE 4
#define PYL_SHELL_XPLANE 5
#define PYL_SHELL_XI 6
/* expected_Pk's log-spaced interpolation table (Pk_library.pyx:1983-1993), DEVICE float32 arrays of n entries */
typedef struct pyl_shell_table {
    const float *k;
    const float *P;
    int32_t n;
    float kF;          /* fundamental frequency, float32 like the reference's `cdef float kF` */
    double log10_kmin; /* log10(k_in[0]) */
    double deltak;     /* float32 spacing in log10 k, promoted */
} pyl_shell_table_t;
int pyl_shell_layout(int kind, int dims, int *bins, int *values);
size_t pyl_shell_bin_workspace_bytes(int kind, int dims);
int pyl_shell_bin(int kind, const float *const *fields, int nfields, const int *mas_index, int dims, int axis,
                  float scale, const pyl_shell_table_t *table, void *out, void *ws, size_t ws_bytes,
                  pyl_stream_t stream);

/* Elementwise passes over a (dims,dims,dims/2+1) complex64 half-spectrum, IN PLACE on a_k / delta_k:
 *   pyl_modes_deconvolve  correct_MAS's loop (Pk_library.pyx:1909-1929): multiply by the float32 window.  The
 *       reference corrects only the modes it counts as independent, leaving their Hermitian duplicates on the
 *       kz = 0 / Nyquist planes untouched; its c2r transform then sees the Hermitian part of those planes, i.e.
 *       (1 + w)/2 on both members of a pair.  That factor is applied explicitly here (planes stay Hermitian).
 *   pyl_modes_power       Xi / XXi's loop (:2198-2218, :2335-2362): a_k <- (re_a*re_b + im_a*im_b, 0) of the
 *       deconvolved modes in float32; b_k NULL = auto-correlation.  b_k is not modified. */
size_t pyl_modes_workspace_bytes(int dims);
int pyl_modes_deconvolve(float *delta_k, int dims, int mas_index, void *ws, size_t ws_bytes, pyl_stream_t stream);
int pyl_modes_power(float *a_k, const float *b_k, int dims, int mas_a, int mas_b, void *ws, size_t ws_bytes,
                    pyl_stream_t stream);
/* XXi_projected (Pk_library.pyx:2684-2789), the 2D forms: pyl_modes_power_2d deconvolves EVERY stored mode of two
 * (dims, dims/2+1) half-spectra (the reference's 2D loop has no duplicate-mode rule, :2735-2758) and leaves
 * a_k = (re_a*re_b + im_a*im_b, 0) in float32; pyl_radial_bin_2d bins every cell of a (dims, dims) float32 image
 * (times `scale`, a float32 product) by int(sqrt(kx^2+ky^2)) (:2776-2787) into out = [sum |k| (float64) |
 * count (uint64) | sum value (float64)], each kmax+1 = int((dims/2)*sqrt(2))+1 words. */
int pyl_modes_power_2d(float *a_k, const float *b_k, int dims, int mas_a, int mas_b, void *ws, size_t ws_bytes,
                       pyl_stream_t stream);
int pyl_radial_bin_2d(const float *image, int dims, float scale, void *out, pyl_stream_t stream);
/* a_k[i] *= b_k[i], complex64 (smoothing_library.pyx:227-232, field_k * filter_k) */
int pyl_cmul_inplace(float *a_k, const float *b_k, int64_t n_complex, pyl_stream_t stream);
/* Smoothing filters placed on a grid (smoothing_library.pyx:37-100 FT_filter, :141-191 FT_filter_2D), axes = 3 | 2:
 *   kind 0 Top-Hat   : out = float32 (dims,)*axes, 1 where d2 <= R2 (d2 = squared distance in cells, periodic)
 *   kind 1 Gaussian  : out = float32 (dims,)*axes, exp(-d2 / (2 R2)) evaluated in double
 *   kind 2 Top-Hat-k : out = complex64 (dims,)*(axes-1) x (dims/2+1), 1 where kmin <= kF*sqrt(d2) < kmax, DC kept
 * R2, kF, kmin, kmax are float32 like the reference's C locals.  Normalisation (sum -> pyl_sum_f64, division ->
 * pyl_divide_by_f64) and the transforms are separate calls. */
int pyl_filter_fill(int kind, void *out, int dims, int axes, float R2, float kF, float kmin, float kmax,
                    pyl_stream_t stream);
/* x[i] = float(double(x[i]) / divisor[0]), divisor a DEVICE double (smoothing_library.pyx:110-114) */
int pyl_divide_by_f64(float *x, int64_t n, const double *divisor, pyl_stream_t stream);
/* v[i] *= (1 + delta[i]), float32: the momentum fields of XPk_dv / XPk_vv (Pk_library.pyx:1367, :1491-1492) */
int pyl_mul_one_plus(float *v, const float *delta, int64_t n, pyl_stream_t stream);

/* ---- 4. host-pointer entry points with the reference's exact C signature ------------ */
/* Same argument list as MAS_c.h:3-10 (HOST pointers; `threads` is accepted and ignored).
 * They copy pos/W/number to the GPU, run pyl_deposit, and copy number back; a maintainer
 * can point MAS_c.pxd at them unchanged.  Return PYL_* status. */
int pyl_NGP(float *pos, float *number, float *W, long particles, int dims, int axes,
            float BoxSize, int threads);
int pyl_CIC(float *pos, float *number, float *W, long particles, int dims, int axes,
            float BoxSize, int threads);
int pyl_TSC(float *pos, float *number, float *W, long particles, int dims, int axes,
            float BoxSize, int threads);
int pyl_PCS(float *pos, float *number, float *W, long particles, int dims, int axes,
            float BoxSize, int threads);
/* free the arena owned by the entry points above */
int pyl_host_arena_release(void);

#ifdef __cplusplus
}
#endif
#endif /* PYL_B200_H */
