/* libpymd_b200 -- C ABI of the B200-native generalized-Langevin stepping path.
 *
 * The reference (phyjtlu/pymd) has no FFI for this path: it is the Python classes
 * md / phbath / ebath.  pymd_b200/{md,phbath,ebath}.py keep
 * those class interfaces and call the entry points below through ctypes; each entry
 * point cites the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (pymd_last_error() explains);
 *     the reference's "print + sys.exit()" convention becomes a Python exception.
 *   - "host" pointers are plain C arrays in the reference's own shapes (row-major);
 *     the library copies/pads what it needs.  "dev" pointers are device memory
 *     BORROWED for the lifetime of the binding (torch tensors on the Python side).
 *   - ensemble layout: every per-trajectory array has the trajectory index FASTEST,
 *     e.g. p is [nd][ldb] doubles.  ldb (the padded ensemble size) is a multiple of 32.
 *   - all launches go to the cudaStream_t passed as `stream` (void* here); no hidden
 *     synchronisation.  One ctx per process per GPU; not re-entrant.
 */
#ifndef PYMD_B200_H
#define PYMD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pymd_ctx pymd_ctx;

#define PYMD_MAX_BATHS 8
#define PYMD_MAX_CONSTRAINTS 16

const char* pymd_last_error(void);
int pymd_version(void);

/* md.__init__ (md.py:59-125): nd = md.nph, ldb = padded ensemble size, dt, nmd. */
int pymd_ctx_create(int nd, int ldb, int nbath, double dt, int nmd, int device, pymd_ctx** out);
int pymd_ctx_destroy(pymd_ctx* ctx);

/* md.setDyn result (md.py:211-250): host dyn[nd][nd] as md.dyn holds it after the
 * eigen clean-up; used as f = -dyn.q (md.potforce, md.py:483-484). */
int pymd_set_dyn(pymd_ctx* ctx, const double* dyn_host);

/* md.constraint (md.py:66, ApplyConstraint md.py:739-762): host c[n][nd]; rows are
 * normalised as the reference does on every call; zero rows stay zero. */
int pymd_set_constraints(pymd_ctx* ctx, const double* constr_host, int n);

/* md.AddBath (md.py:160-176) + the operands of bath.bforce:
 *   phbath.bforce (phbath.py:192-202): kernel[ml][nc][nc], friction scaled by dt when ml>1
 *                                       (dt_scaled=1), unscaled for the local Debye bath.
 *   ebath.bforce  (ebath.py:180-199):  ml=1, kernel[0] = efric + bias*zeta2 (acts on p),
 *                                       aq = bias*(exim - zeta1) (acts on q), dt_scaled=0.
 * cids_host[nc] are the connected dofs (phbath.py:52, ebath.py:33).  aq_host may be NULL. */
int pymd_bath_set(pymd_ctx* ctx, int b, int nc, const int* cids_host, int ml,
                  const double* kernel_host, const double* aq_host, int dt_scaled);

/* md.p, md.q (md.py:105-106): dev [nd][ldb]. */
int pymd_bind_state(pymd_ctx* ctx, double* p_dev, double* q_dev);

/* md.phis restricted to the bath's dofs (md.py:292-301, functions.rpadleft): a ring
 * buffer dev [ring_len][ncp][ldb], ncp = pymd_bath_ncp(nc), ring_len >= ml-1+block
 * (see pymd_ring_len).  Only non-local baths (ml>1) need one. */
int pymd_bath_ncp(int nc);
int pymd_ring_len(int ml, int block);
int pymd_bind_history(pymd_ctx* ctx, int b, double* ring_dev, int ring_len);

/* bath.noise (phbath.py:157, ebath.py:147): dev [nmd][nc][ldb]. */
int pymd_bind_noise(pymd_ctx* ctx, int b, const double* noise_dev);

/* md.ps, md.qs (md.py:146-147, written at md.py:322-324; NULL when savepq is False),
 * md.etot (md.py:119,328): dev [nmd][nd][ldb], [nmd][nd][ldb], [nmd][ldb]. */
int pymd_bind_outputs(pymd_ctx* ctx, double* ps_dev, double* qs_dev, double* etot_dev);

/* Overlaid records (optional): saved trajectories and noise of a large ensemble in ONE buffer.  Row t of ps / qs is
 * written exactly when noise row t has been consumed, so the rows of [ps | qs] (2 nd ldb doubles each) may overlay the
 * END-aligned noise rows (sum_b nc_b ldb doubles each, <= 2 nd ldb) of the same buffer: 129 GB instead of 218 GB for
 * 4096 trajectories of config 2, which lets them step as one ensemble with savepq.  The rows are then strided:
 * row_stride = doubles between consecutive time rows (ps and qs share it; every bath's noise shares the noise row).
 * The noise is no longer periodic in place: row `nmd` (a copy of row 0, pymd_set_noise_wrap_row(ctx, nmd)) is what the
 * last step of a segment reads (phbath.py:196), and a pymd_step call must not run past the end of the noise table.
 * Only the two-unit on-chip step kernel (fused3.cu) supports strided records; pymd_step fails otherwise. */
int pymd_bind_outputs_strided(pymd_ctx* ctx, double* ps_dev, double* qs_dev, double* etot_dev, long long row_stride);
int pymd_bind_noise_strided(pymd_ctx* ctx, int b, const double* noise_dev, long long row_stride);
int pymd_set_noise_wrap_row(pymd_ctx* ctx, int row);

/* bath.cur (md.py:342): dev [nmd][ldb];  md.fhis[b] restricted to the bath's dofs
 * (md.py:343; NULL = not recorded): dev [nmd][nc][ldb]. */
int pymd_bind_bath_outputs(pymd_ctx* ctx, int b, double* cur_dev, double* fhis_dev);

/* md.ResetHis (md.py:292-301): zero every ring, forget cached forces. */
int pymd_reset_history(pymd_ctx* ctx, void* stream);
/* Tell the library that p/q were modified from outside (md.initialise, a resume):
 * the cached potential force (md.potforce's q0/f0, md.py:456-457) is dropped. */
int pymd_invalidate_cache(pymd_ctx* ctx);

/* History access for checkpoint/resume (md.dump writes phis, md.py:700-703): copies the
 * newest `nrows` rows (newest first, like md.phis) of bath b to/from dev [nrows][nc][ldb]. */
int pymd_history_get(pymd_ctx* ctx, int b, double* out_dev, int nrows, void* stream);
int pymd_history_set(pymd_ctx* ctx, int b, const double* in_dev, int nrows, void* stream);

/* nsteps calls of md.vv (md.py:315-358) starting at md.t = t0, for the whole ensemble.
 * mode: 0 = auto, 1 = per-step GEMM pipeline (any configuration; memories of 64 steps and more through the
 * frequency-domain delay line of fdl.cu), 2 = on-chip step kernels (fused3.cu / fused.cu for nd <= 64 with the FFT
 * far field of far.cu when every memory bath has ml >= 24 and nc <= 16; local.cu for 64 < nd <= 96 with one
 * memory-less bath); mode 2 fails if the configuration fits neither.  Auto prefers mode 2 where it applies, except
 * for memories of 2000 steps and more. */
int pymd_step(pymd_ctx* ctx, long long t0, int nsteps, int mode, void* stream);
/* number of kernel launches issued by this ctx so far */
long long pymd_launch_count(const pymd_ctx* ctx);
/* Per-kernel-class device timing with CUDA events on the launching stream (measurement aid for
 * bench.py's roofline; adds an event pair per launch while enabled).  Classes: 0 = memory-kernel
 * tail GEMM, 1 = fused step kernel, 2 = other tensor-core GEMMs, 3 = elementwise stage kernels,
 * 4 = FFT far field of the memory kernel, 5..7 unused.
 * pymd_profile_read synchronises, writes the accumulated milliseconds and launch counts of the 8
 * classes and clears them. */
#define PYMD_PROFILE_CLASSES 8
int pymd_profile_enable(pymd_ctx* ctx, int on);
int pymd_profile_read(pymd_ctx* ctx, double* ms8, long long* count8);

/* ---- noise generation: noise.phnoise / noise.enoise (noise.py:38-88, 135-190) --------
 * factors: L[f] = U_f diag(sqrt(max(lambda_f,0))) for f = 0..nmd/2, dev [nf][nc][nc]
 * (real part; imaginary part NULL for phonon baths).  xi: standard normals dev
 * [nf][nc][ldb].  Computes X_f = L_f xi_f, the Hermitian mirror (noise.py:74-81), the
 * w->t transform myfft.iFourier1D (myfft.py:35-52: fft/(nmd*dt)) and keeps the real part.
 * out: dev [nmd][nc][ldb]; one call produces all component rows of the trajectories n0 <= n < n1
 * (multiples of 8), so that the FFT scratch (pymd_noise_work_bytes(nmd, nc, n1-n0)) stays small for
 * large ensembles and every white-noise row is read exactly once. */
size_t pymd_noise_work_bytes(int nmd, int nc, int ncols);
int pymd_noise_generate(const double* lre_dev, const double* lim_dev, const double* xi_dev,
                        int nmd, int nc, int ldb, int n0, int n1, double dt, double* out_dev,
                        void* work_dev, void* stream);

/* The same for white noise held per trajectory chunk: xi_chunk dev [nf][nc][cw] belongs to the trajectories
 * n0 <= n < n0 + cw of out (dev [nmd][nc][ldb]); cw a multiple of 8.  With pymd_philox_normal called per chunk
 * (traj0 + n0) the white noise of a large ensemble never exists as a whole (noise.py:282 draws it value by value). */
int pymd_noise_generate_chunk(const double* lre_dev, const double* lim_dev, const double* xi_chunk_dev, int cw,
                              int nmd, int nc, int ldb, int n0, double dt, double* out_dev, void* work_dev, void* stream);

/* ... with the rows of out `out_row_stride` doubles apart (overlaid records, see pymd_bind_noise_strided) */
int pymd_noise_generate_chunk_ld(const double* lre_dev, const double* lim_dev, const double* xi_chunk_dev, int cw,
                                 int nmd, int nc, int ldb, int n0, double dt, double* out_dev, long long out_row_stride,
                                 void* work_dev, void* stream);

/* N.random.normal replacement (noise.py:282): counter-based Philox4x32-10 + Box-Muller using both
 * outputs (rows 2r and 2r+1 share the counter r).  Element (row r, trajectory n) depends only on
 * (seed, stream_id, r, traj0+n): results do not depend on how the ensemble is sharded over GPUs.
 * out: dev [rows][ldb]. */
int pymd_philox_normal(double* out_dev, long long rows, int ldb, uint64_t seed, uint32_t stream_id,
                       long long traj0, void* stream);
/* N.random.rand replacement (md.py:280). */
int pymd_philox_uniform(double* out_dev, long long rows, int ldb, uint64_t seed, uint32_t stream_id,
                        long long traj0, void* stream);

/* ---- md.initialise (md.py:253-290): q = sum_i U[:,i] am_i cos(2 pi r_i),
 * p = -sum_i hw_i U[:,i] am_i sin(2 pi r_i), constraints applied after every mode.
 * U_host[nd][nd] (columns = modes), am_host[nd], hw_host[nd]; r: dev [nd][ldb]. */
int pymd_init_state(pymd_ctx* ctx, const double* U_host, const double* hw_host,
                    const double* am_host, const double* r_dev, void* stream);

/* ---- spectrum.powerspec / powerspec2 (spectrum.py:8-45) summed over the ensemble:
 * traj: dev [nmd][nd][ldb] (md.qs or md.ps); out: dev [nmd] receives
 * sum_{traj<nvalid} sum_dof |Fourier1D(traj)|^2 / (dt*nmd)  (times w_i^2 when wsq != 0).
 * The nd*ldb columns are transformed chunk_cols at a time (0 = all at once; multiple of 32);
 * work: dev scratch of pymd_powerspec_work_bytes().  The sum is deterministic. */
size_t pymd_powerspec_work_bytes(int nmd, int nd, int ldb, long long chunk_cols);
int pymd_powerspec(const double* traj_dev, int nmd, int nd, int ldb, int nvalid, double dt, int wsq,
                   double* out_dev, void* work_dev, long long chunk_cols, void* stream);

/* pymd_powerspec for rows `row_stride` doubles apart (overlaid records) */
int pymd_powerspec_ld(const double* traj_dev, long long row_stride, int nmd, int nd, int ldb, int nvalid, double dt, int wsq,
                      double* out_dev, void* work_dev, long long chunk_cols, void* stream);

/* Time means of per-trajectory records (the kappa.*.dat numbers of md.Run, md.py:619: mean over the nmd rows of
 * bath.cur; likewise md.etot): out[n] = scale * sum_r a[r][n] for a dev [rows][ldb], deterministic (fixed chunking).
 * work: dev scratch of pymd_column_sums_work_bytes(ldb). */
size_t pymd_column_sums_work_bytes(int ldb);
int pymd_column_sums(const double* a_dev, long long rows, int ldb, double scale, double* out_dev, void* work_dev, void* stream);

/* Batched complex FFT along the slowest axis: data dev [n][cols] complex128 in place,
 * sign = -1 (numpy fft) or +1 (numpy ifft, unscaled). */
int pymd_fft_columns(double* data_dev, int n, long long cols, int sign, void* work_dev, void* stream);
/* bytes of scratch pymd_fft_columns needs */
size_t pymd_fft_work_bytes(int n, long long cols);
/* sweeps over HBM (kernel launches) of a length-n transform: the plan combines fused radix-128/64/16
 * passes and radix-8/4/2 passes for the fewest sweeps (1 up to n = 128 except 32, 2 up to 2^14, 3 up to 2^17); -1 if n is not a power of two */
int pymd_fft_num_passes(int n);

/* General tensor-core product used by tests and by setup-on-device:
 * Y[M][ldb] = alpha * A_host-uploaded... (device pointers, A row-major [M][lda], lda % 32 == 0) */
int pymd_gemm(const double* a_dev, int lda, const double* x_dev, double* y_dev, int M, int K, int ldb,
              double alpha, int accumulate, void* stream);

/* ---- set-up tables on the device ------------------------
 * phbath.gamt (phbath.py:221-254): out[k][e] = scale * sum_i coef(tl[k], w[i]) g[i][e] with
 * coef = 2 cos(t w) for eta_ad == 0, else the damped form
 * exp(-eta t) [w/(w - i eta) e^{-iwt} + w/(w + i eta) e^{+iwt}].  phbath.gmem passes
 * scale = w[nw-1]/(pi nw) and g = flinterp(w, gwl, gamma) flattened to ne = nc*nc columns; the
 * eta_ad re-derivation of gamma (phbath.py:175-189) is the same transform with the roles of the two
 * grids exchanged.  All pointers dev: tl[ml], w[nw], g[nw][ne], out[ml][ne]. */
int pymd_gamt(const double* tl_dev, int ml, const double* w_dev, int nw, const double* g_dev, int ne,
              double eta_ad, double scale, double* out_dev, void* stream);

/* harmonic.Drs / DDos / UU (harmonic.py:37-84) for nw frequencies at once:
 * D = [(w + i eta)^2 - dyn - se(w)]^-1 (n x n complex, inverted in shared memory, one CTA per
 * frequency), ddos[f] = -2 w Tr Im D, uu[f] = Tr[D pi(w) D^dagger] (PP = w^2 uu).
 * dev pointers: wl[nw], dyn[n][n], se and pi [nw][n][n] complex128 (interleaved re, im), ddos[nw],
 * uu[nw]; pi and uu may both be NULL.  n <= pymd_harmonic_max_n(). */
int pymd_harmonic_max_n(void);
int pymd_harmonic(const double* wl_dev, int nw, const double* dyn_dev, const double* se_dev, const double* pi_dev,
                  int n, double eta, double* ddos_dev, double* uu_dev, void* stream);

/* numpy.linalg.eigh of the noise covariances (noise.py:69, 172) for nf frequencies at once, by cyclic
 * Jacobi rotations in shared memory (one CTA per matrix).  c: dev [nf][n][n], real part and (Hermitian
 * case; NULL for real symmetric) imaginary part.  Outputs, eigenvalues ascending as eigh orders them:
 * lam dev [nf][n]; l = U diag(sqrt(max(lam, 0))) dev [nf][n][n] (re, im) -- the factor
 * pymd_noise_generate consumes; u (optional, may be NULL) the eigenvectors; sweeps (optional) dev [nf]
 * ints, Jacobi sweeps used.  Eigenvectors are defined up to a phase, so L differs from the LAPACK one
 * by a unitary diagonal: L L^H is the same.  n <= pymd_eigh_max_n(complex?).
 * With ncomb > 0 the matrices are not passed but assembled in the kernel from a small table:
 * c is then dev [ntab][n][n] and C_f = herm(sum_{m<ncomb} coef[f][m] * c[idx[f][m]]), herm(M) = (M + M^H)/2,
 * idx dev [nf][ncomb] int32, coef dev [nf][ncomb] -- three fixed matrices with frequency-dependent weights
 * for the electron bath (noise.py:156-170), two neighbouring gamma-table entries with the flinterp weights
 * for a phonon bath (noise.py:61-66).  ncomb == 0: idx and coef are ignored (may be NULL).
 * Matrices too large to hold twice in shared memory (n > 83 complex, > 118 real) keep the eigenvector
 * accumulator in a global scratch: work dev of pymd_eigh_work_bytes(nf, n, complex?) bytes (0: not needed,
 * work may be NULL). */
int pymd_eigh_max_n(int cplx);
size_t pymd_eigh_work_bytes(int nf, int n, int cplx);
int pymd_eigh_factors(const double* cre_dev, const double* cim_dev, const int* idx_dev, const double* coef_dev, int ncomb,
                      int nf, int n, double* lam_dev, double* lre_dev, double* lim_dev, double* ure_dev, double* uim_dev,
                      int* sweeps_dev, void* work_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PYMD_B200_H */
