/* Thin C ABI between the C11 host library and the sm_100a CUDA kernels.
 *
 * Plain pointers, ints and doubles only; every function returns 0 on success and a negative code on
 * failure (sxs_cuda_last_error() gives the text).  The host library (libfmftsaxs_b200/csrc/host) turns
 * failures into the reference's behaviour — message on stderr and exit (src/common.h:26-36).
 * There is no CPU implementation behind any of these entry points.
 *
 * Reference interfaces replaced:
 *   sxs_cuda_expand            atom_grp2spf_inplace            src/pdb2spf.c:24-152      (kernel K1)
 *   sxs_cuda_plan_*            sxs_compute_saxs_scores         src/fftsaxs.c:608-986     (kernels K2-K4)
 *     K2 rotate/translate      compute_sum1, fill_t_matrix     src/fftsaxs.c:182-333
 *     K3 angular DFT + cross   compute_sum2, simple_ft/fftw,   src/fftsaxs.c:416-606, 52-108
 *                              fill_const, fill_var
 *     K4 fused fit             sxs_fit_params, sxs_lbfgs_fitting + setulb   src/min_saxs.c:153-259
 *   sxs_cuda_fit_profiles      sxs_fit_params / sxs_lbfgs_fitting on explicit cross terms
 *   sxs_cuda_profile_from_spf  sxs_profile_from_spf            src/profile.c:186-233
 *
 * Flat layouts:
 *   coef  [((c*qnum + q)*(L+1)^2 + lm)*2 + {re,im}]   c = 0:V 1:D 2:W, lm = l(l+1)+m
 *   cross [(p*6 + k)*qnum + q]                        k = VV,VD,VW,DD,DW,WW (host-facing fit entry)
 */
#ifndef FMFTSAXS_SXS_CUDA_H
#define FMFTSAXS_SXS_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sxs_cuda_plan sxs_cuda_plan;

int sxs_cuda_device_count(void);
/* Measured FP64 FMA throughput of the device (TFLOP/s): the roofline denominator of the FP64-bound kernels. */
int sxs_cuda_fp64_peak(int device, double *tflops);
const char *sxs_cuda_last_error(void);

/* K1.  Per-atom inputs are the reference's own intermediates, computed on the host with libm so the
 * angles match bit for bit (cart2sph, src/pdb2spf.c:9-22): r, cos(theta), phi, and the three
 * q-independent form factors (vacuum, dummy, h2o*SASA fraction).  ynorm = generate_spherical_norm(L+1),
 * inv_dfact[l] = 1/(2l+1)!!, four_pi = 4*SXS_PI.  coef is a host buffer in the flat layout. */
int sxs_cuda_expand(int device, int natoms, const double *r, const double *cos_theta, const double *phi,
                    const double *ff_vacuum, const double *ff_dummy, const double *ff_water,
                    const double *qvals, int qnum, int L, const double *ynorm, const double *inv_dfact,
                    double four_pi, double *coef);

/* I(q) of one molecule from its coefficients, G frozen at qvals[0] (src/profile.c:214). */
int sxs_cuda_profile_from_spf(int device, const double *coef, int qnum, int L, double mult, const double *qvals,
                              double c1, double c2, double *intensity);

/* Six self cross terms of one molecule, out[k*qnum + q], VD/VW/DW doubled (sxs_spf2cross_terms,
 * src/min_saxs.c:437-499). */
int sxs_cuda_self_terms(int device, const double *coef, int qnum, int L, double *out);

/* Scoring plan: device-resident tables for one (L, q grid). */
sxs_cuda_plan *sxs_cuda_plan_create(int device, int L, int qnum, const double *qvals, const double *dsymb,
                                    const double *dwig, const double *twiddle);
void sxs_cuda_plan_destroy(sxs_cuda_plan *plan);

/* Uploads both molecules, computes the self terms (comp_const_int, src/fftsaxs.c:27-50) and the
 * beta/gamma-rotated coefficient tables of receptor (A) and ligand (B). */
int sxs_cuda_plan_set_molecules(sxs_cuda_plan *plan, const double *coefA, const double *coefB);
/* Compressed experiment a[6*qnum] (scoring_helper), mult and peak (src/min_saxs.c:116-124). */
int sxs_cuda_plan_set_experiment(sxs_cuda_plan *plan, const double *a, double mult, double peak);
/* Translation steps: bessel[(zi*qnum + q)*(2L+1) + p] = j_p(q z_zi) from the host's reference-exact series. */
int sxs_cuda_plan_set_translations(sxs_cuda_plan *plan, const double *bessel, int znum);

/* Score poses given as flat grid indices (host buffers; H2D/D2H inside).  Only z digits in
 * [z_lo, z_hi) are scored — the caller's shard; others and out-of-range digits keep their values. */
int sxs_cuda_plan_score_i32(sxs_cuda_plan *plan, const int *index, long long nout, int z_lo, int z_hi,
                            double *scores, double *c1, double *c2);
int sxs_cuda_plan_score_i64(sxs_cuda_plan *plan, const long long *index, long long nout, int z_lo, int z_hi,
                            double *scores, double *c1, double *c2);

/* Same with every buffer already in device memory; work is enqueued on `stream` (a cudaStream_t,
 * NULL = default stream) and the call returns after the last kernel is queued, except for one small
 * synchronising read of the sorted list's offsets. */
int sxs_cuda_plan_score_dev_i32(sxs_cuda_plan *plan, const int *d_index, long long nout, int z_lo, int z_hi,
                                double *d_scores, double *d_c1, double *d_c2, void *stream);
int sxs_cuda_plan_score_dev_i64(sxs_cuda_plan *plan, const long long *d_index, long long nout, int z_lo, int z_hi,
                                double *d_scores, double *d_c1, double *d_c2, void *stream);

/* Counters of the last score call: [0] distinct grid points fitted, [1] (z,beta2) slabs translated,
 * [2] kernel launches, [3] K3 threads that took more than one point (points differing only in a2 share a thread;
 * 0 = one point per thread), [4] z groups. */
int sxs_cuda_plan_stats(const sxs_cuda_plan *plan, long long *stats5);

/* Histogram of objective evaluations per fit over the distinct points of the last score call
 * (hist64[n] = fits that took n evaluations, bin 63 = 63 or more; the reference's isave[33]). */
int sxs_cuda_plan_fit_evaluations(sxs_cuda_plan *plan, long long *hist64);

/* Device-side timing of the kernel classes of subsequent score calls, with CUDA events recorded on the
 * launching stream: ms5/launches5 = {0 key sort + distinct points, 1 T-matrix + translation, 2 cross terms
 * (K3), 3 fit (K4), 4 scatter}.  kernel_times() waits for the recorded events and resets the counters. */
int sxs_cuda_plan_set_profiling(sxs_cuda_plan *plan, int on);
int sxs_cuda_plan_kernel_times(sxs_cuda_plan *plan, double *ms5, long long *launches5);

/* Pinned host memory owned by the plan (grow-only, freed with it); NULL on failure.  The host layer stages the compact
 * index list and the result columns of a z shard in it. */
void *sxs_cuda_plan_host_buffer(sxs_cuda_plan *plan, size_t bytes);

/* Dense scan (the reference's skip = 0 mode of src/fftsaxs.c:867-872 computes every point of every cell
 * but reports only listed rows): every grid point (b1, b2, a2, g1, g2) of the z steps [z_lo, z_hi) is scored and the
 * k points of lowest chi are returned, chi ascending: flat 64-bit index (-1 = empty slot), chi, c1, c2.  Host arrays
 * of length k. */
int sxs_cuda_plan_scan_topk(sxs_cuda_plan *plan, int z_lo, int z_hi, int k, long long *index, double *scores,
                            double *c1, double *c2);

/* Debug/stage access for parity tests: cross terms of the listed poses, cross[(row*6 + k)*qnum + q],
 * before the peak rescale (what fill_const/fill_var leave in the profile, src/fftsaxs.c:52-108). */
int sxs_cuda_plan_cross_terms_i32(sxs_cuda_plan *plan, const int *index, long long nout, double *cross);

/* K4 alone on explicit cross terms: out[p*4 + {0:chi,1:c1,2:c2,3:evaluations}].  rescale != 0 applies the
 * peak rescale of sxs_fit_params first (src/min_saxs.c:170-188). */
int sxs_cuda_fit_profiles(int device, const double *cross, long long npts, const double *a, const double *qvals,
                          int qnum, double mult, double peak, int rescale, double *out);
/* Objective pieces at a given (c1, c2) for one profile: out[0] = best scale k, out[1] = f, out[2..3] = gradient. */
int sxs_cuda_fit_eval(int device, const double *cross, const double *a, const double *qvals, int qnum, double mult,
                      double c1, double c2, double *out4);

/* exp() as the objective evaluates it on the device (exp_glibc.h: the reference libm's algorithm), y[i] = exp(x[i]);
 * host arrays.  Stage access for the parity tests. */
int sxs_cuda_exp_array(int device, const double *x, long long n, double *y);

/* ft rows -> flat grid indices on the device: what sxs_ft_rows_to_indices64 (index.h) computes on host
 * threads — sxs_ft2euler (src/index.c:38-75), the three-decimal text round trip of the Euler file (src/index.c:114,
 * tools/correlate.c:214), the z table lookup and the snapping of tools/correlate.c:219-247 — one thread per row.
 * rots: nrot row-major 3x3 matrices (struct mol_matrix3 is nine doubles); ref_lig: three host doubles.
 * index[i] = flat 64-bit index of row i, -1 when its z is not on zvals (the tool drops the row) or its geometry is
 * degenerate (a 0 / 0 angle: the tool's index is built from (int)round(nan) there), -2 when rot_id[i] is outside the
 * table.  Host arrays; the _dev form takes device arrays and a stream (ref_lig stays a host pointer). */
int sxs_cuda_ft_rows_to_indices(int device, const int *rot_id, const double *trans, long long n, const double *rots,
                                long long nrot, const double *ref_lig, const double *zvals, int znum, int L,
                                long long *index);
int sxs_cuda_ft_rows_to_indices_dev(const int *d_rot_id, const double *d_trans, long long n, const double *d_rots,
                                    long long nrot, const double *ref_lig, const double *d_zvals, int znum, int L,
                                    long long *d_index, void *stream);

#ifdef __cplusplus
}
#endif

#endif
