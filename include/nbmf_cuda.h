/*
 * nbmf_cuda.h -- C ABI of libnbmf_cuda, the B200 (sm_100a) engine for the
 * data-parallel hot path of alumbreras/NBMF (R package rMMLEDirBer).
 *
 * The library replaces the *bodies* of the reference's Rcpp hot functions; the
 * R entry point BernoulliNMF() (R/BernoulliNMF.R:12-106), the generated glue
 * (src/RcppExports.cpp:259-344, R/RcppExports.R:4-86) and the C++ signatures
 * stay as they are.  INTEGRATION.md shows the marshalling bodies.
 *
 * Conventions (the reference's own):
 *   - host matrices are column-major; V and Z are int32, NA = INT_MIN
 *     (R's NA_INTEGER; tested as V(f,n) != NA_INTEGER at
 *     collapsed_gibbs_z_VB.cpp:56,124,449); Z labels are 0-based, NA where V
 *     is NA; every output is host fp64 column-major, caller-owned.
 *   - V is F x N.  W is the Dirichlet factor (F x K, rows sum to 1), H the
 *     Beta factor (N x K), as VB_Aspect_Rcpp returns them
 *     (collapsed_gibbs_z_VB.cpp:507-525).
 *   - every function returns 0 on success or a negative nbmf_status;
 *     nbmf_last_error() gives the message the Rcpp body passes to Rcpp::stop().
 *     The library never calls R, never throws, never prints.
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with NBMF_ERR_CUDA.
 */
#ifndef NBMF_CUDA_H
#define NBMF_CUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBMF_NA_INTEGER INT32_MIN

typedef enum nbmf_status {
    NBMF_OK = 0,
    NBMF_ERR_ARG = -1,     /* bad argument (the reference would index out of bounds) */
    NBMF_ERR_CUDA = -2,    /* CUDA runtime / driver error, no device, wrong arch     */
    NBMF_ERR_STATE = -3,   /* call order (no data, no statistics, ...)               */
    NBMF_ERR_NEGATIVE = -4,/* the reference's "Negative values in ..." stop()        */
    NBMF_ERR_OOM = -5,
    NBMF_ERR_UNSUPPORTED = -6
} nbmf_status;

/* which arithmetic the contraction kernel uses */
typedef enum nbmf_precision {
    NBMF_PREC_AUTO = 0,        /* the fastest variant whose measured error growth stays inside the
                                  1e-5 gate on W and H for the planned number of iterations at
                                  this size (nbmf_opts.planned_iters):
                                  FP64 for K <= 32 or small F, N, else one of the tensor variants */
    NBMF_PREC_FP64 = 1,        /* FP64: DMMA tensor pipe for Kp <= 128, FMA + shuffles above       */
    NBMF_PREC_TENSOR = 2,      /* tcgen05 f16: stage 2 with the streamed operand split hi+lo,
                                  fp32 accumulate in TMEM                                          */
    NBMF_PREC_TENSOR_FAST = 3, /* same, hi only in stage 2: 1/3 fewer MMAs; its error falls
                                  like ~1/sqrt(min(F,N)) -- short runs at very large sizes         */
    NBMF_PREC_TENSOR_LO8 = 4,  /* the hi+lo form with the 2^-11-sized lo term of stage 2 computed through
                                  kind::f8f6f4 on e5m2 copies of R and W_lo: the accuracy of TENSOR at
                                  5/4 (not 6/4) of the FAST variant's tensor time                   */
    NBMF_PREC_FP32 = 5         /* FP32 FMA + warp shuffles (small K, short runs)    */
} nbmf_precision;

/* contraction actually used by the last VB session (nbmf_timing.path) */
typedef enum nbmf_path {
    NBMF_PATH_FP64 = 0,
    NBMF_PATH_TENSOR = 1,      /* f16, stage 2 hi+lo          */
    NBMF_PATH_TENSOR_FAST = 2, /* f16, hi only                */
    NBMF_PATH_TENSOR_LO8 = 3,  /* f16 hi + e5m2 lo in stage 2 */
    NBMF_PATH_FP32 = 4
} nbmf_path;

/* VB_DirBer_Rcpp modes */
typedef enum nbmf_dirber_mode {
    NBMF_DIRBER_EXACT_WAVEFRONT = 0, /* bit-identical schedule to collapsed_gibbs_z_VB.cpp:119-141 */
    NBMF_DIRBER_CONTRACTION = 1      /* self-term-free batch form        */
} nbmf_dirber_mode;

typedef struct nbmf_opts {
    int device;          /* CUDA ordinal                                                    */
    int precision;       /* nbmf_precision                                                  */
    int flush_interval;  /* tensor path: steps between TMEM accumulator flushes; 0 = default */
    int pipe_block_cols; /* nbmf_vb_aspect_bits: columns per upload block; 0 = about N/8         */
    int64_t n_global;    /* column sharding: total number of columns (0 = not sharded)      */
    int64_t n_offset;    /* first global column this handle owns                            */
    int planned_iters;   /* AUTO precision: VB updates the caller will run from this start
                            (nbmf_vb_begin / nbmf_vb_step sessions; nbmf_vb_aspect* use their own
                            iter).  0 = unknown = 100, VB_Aspect_Rcpp's default
                            (collapsed_gibbs_z_VB.cpp:376)                                   */
    int reserved;
} nbmf_opts;

typedef struct nbmf_handle nbmf_handle;

/* collective hook: called on the exchange step of each iteration with a device
 * buffer the caller must sum across ranks in place (count elements of fp64),
 * enqueued on `stream` (a cudaStream_t) -- the handle's own stream, or, on the
 * tensor path, a priority side stream so that the collective overlaps the second
 * pass; the hook must order its work on the stream it is given and not
 * synchronise the device.  Return 0 on success.  This is the only data-path
 * collective: the F x K Dirichlet-side statistic.               */
typedef int (*nbmf_allreduce_fn)(void *ctx, void *device_buf, size_t count, void *stream);

/* timing of the last VB / Gibbs call, CUDA events on the handle's stream */
typedef struct nbmf_timing {
    double total_ms;        /* iteration loop only (no H2D / D2H / packing)  */
    double params_ms;
    double pass_rows_ms;    /* owner = rows f (Dirichlet-side statistic)     */
    double pass_cols_ms;    /* owner = columns n (Beta-side statistics)      */
    double exchange_ms;
    int64_t launches;       /* kernels launched inside the loop              */
    int64_t iters;
    int64_t path;           /* nbmf_path of the session */
} nbmf_timing;

const char *nbmf_version(void);
int nbmf_device_count(void);

int nbmf_create(nbmf_handle **out, const nbmf_opts *opts);
void nbmf_destroy(nbmf_handle *h);
const char *nbmf_last_error(const nbmf_handle *h);
void *nbmf_stream(nbmf_handle *h); /* cudaStream_t the handle launches on */
int nbmf_synchronize(nbmf_handle *h);
int nbmf_set_allreduce_hook(nbmf_handle *h, nbmf_allreduce_fn fn, void *ctx);
int nbmf_get_timing(const nbmf_handle *h, nbmf_timing *out);

/* ---- data ---------------------------------------------------------------
 * replaces: `const arma::imat& V` argument of VB_Aspect_Rcpp / VB_DirBer_Rcpp /
 * Gibbs_DirBer_finite_Rcpp (collapsed_gibbs_z_VB.cpp:373,97;
 * collapsed_gibbs_z.cpp:363).  Packs V into a value-bit plane and a mask-bit
 * plane on the device (2 bits per entry), in both column- and row-packed form. */
int nbmf_set_data_i32(nbmf_handle *h, const int32_t *V_colmajor, int64_t F, int64_t N);
/* already-packed input: bit f of column n is bit (f & 31) of word
 * vbits[n * ceil(F/32) + (f >> 5)]; mbits == NULL means no NA.               */
int nbmf_set_data_bits(nbmf_handle *h, const uint32_t *vbits, const uint32_t *mbits, int64_t F,
                       int64_t N);
/* synthetic DirBer matrix generated on the device (R/utils.R:7-16 sample_V,
 * experiments/xp_paper_basic_example.R:22-30): W_f ~ Dir(a 1), H_nk ~ Beta(a,b),
 * V_fn ~ Bernoulli(W_f . H_n); Philox4x32-10 keyed (seed, f, global n) so every
 * sharding produces the same matrix.  N = local columns.                      */
int nbmf_generate_synthetic(nbmf_handle *h, int64_t F, int64_t N, int Ktrue, double a, double b,
                            uint64_t seed);
int nbmf_get_dims(const nbmf_handle *h, int64_t *F, int64_t *N, int64_t *nobs);
/* column-packed planes back to the host (for tests and checkpoints) */
int nbmf_get_data_bits(nbmf_handle *h, uint32_t *vbits, uint32_t *mbits);

/* ---- initial statistics ---------------------------------------------------
 * replaces vb_matrix_to_tensor + compute_vbcounts (commons.cpp:29-42,
 * collapsed_gibbs_z_VB.cpp:42-66) and compute_counts (collapsed_gibbs_z.cpp:
 * 38-60): a histogram of the labels, no F x K x N tensor.                     */
int nbmf_init_from_labels(nbmf_handle *h, const int32_t *Z_colmajor, int Kc);
/* Z_fn = Philox(seed, f, global n) mod Kc on the device (BernoulliNMF.R:23-33
 * draws sample(K_init,1) per observed entry)                                  */
int nbmf_init_random_labels(nbmf_handle *h, int Kc, uint64_t seed);
/* n_total F x Kc, f_ones N x Kc, f_zeros N x Kc (vbcounts,
 * collapsed_gibbs_z_VB.cpp:14-35); also the checkpoint / resume state.        */
int nbmf_set_stats(nbmf_handle *h, int Kc, const double *n_total, const double *f_ones,
                   const double *f_zeros);
int nbmf_get_stats(nbmf_handle *h, int Kc, double *n_total, double *f_ones, double *f_zeros);

/* ---- VB_Aspect_Rcpp (collapsed_gibbs_z_VB.cpp:373-526) ---------------------
 * iter batch iterations; E_W F x K, E_H N x K from the parameters at the top of
 * the last iteration (:507-514); lowerbounds[iter] (0 when !checklb, :474-477);
 * E_Z (F x K x N, arma::cube layout) only if non-NULL.
 * Requires data + initial statistics with Kc == K.                            */
int nbmf_vb_aspect(nbmf_handle *h, int K, double alpha, double beta, double gamma, int iter,
                   int checklb, double *E_W, double *E_H, double *lowerbounds, double *E_Z);
/* The same call with V and the initial statistics taken from HOST memory, i.e. the whole of
 * VB_Aspect_Rcpp(V, Z, ...) (collapsed_gibbs_z_VB.cpp:373-526) behind one entry point: vbits is
 * the column-packed bit plane of nbmf_set_data_bits (no NA), n_total / f_ones / f_zeros the
 * vbcounts of nbmf_set_stats.  The upload is cut into column blocks and the first iteration
 * consumes them as they arrive, so the host->device transfer overlaps the contraction; pinned
 * host memory is needed for that overlap (pageable memory still works, serially).  E_W and
 * E_H are the parameters at the top of the last iteration (:507-514), so they start their
 * way back beside that iteration.  With a communicator (nbmf_comm_init_rank, or inside a
 * group) the row-side arrays belong to rank 0: n_total is read there and broadcast over
 * NVLink (other ranks may pass NULL), and E_W may be NULL on the other ranks.            */
int nbmf_vb_aspect_bits(nbmf_handle *h, const uint32_t *vbits, int64_t F, int64_t N, int K,
                        double alpha, double beta, double gamma, int iter, int checklb,
                        const double *n_total, const double *f_ones, const double *f_zeros,
                        double *E_W, double *E_H, double *lowerbounds);
/* the same, one iteration at a time (what bench.py times):
 *   begin -> step x iter -> end                                               */
int nbmf_vb_begin(nbmf_handle *h, int K, double alpha, double beta, double gamma, int mean_params);
int nbmf_vb_step(nbmf_handle *h, int checklb);
int nbmf_vb_end(nbmf_handle *h, double *E_W, double *E_H, double *lowerbounds, int n_lb,
                double *E_Z);
/* ELBO split of the last VB session, 3 doubles per iteration: sum_obs log D over
 * the handle's columns, pW - qW (rows: replicated under column sharding),
 * pH - qH (the handle's columns).  lowerbounds[i] of nbmf_vb_end is their sum. */
int nbmf_vb_get_lb_terms(nbmf_handle *h, double *terms3, int n_iter);

/* ---- which.max hand-off (R/BernoulliNMF.R:45-52) ------------------------------
 * Z_fn <- argmax_k E_Z(f,k,n) from the operands of the last VB iteration, computed on
 * the device without the E_Z cube; the labels stay on the device for a following
 * nbmf_gibbs_dirber and are copied to Z_out (F x N, 0-based, NA where V is NA) if
 * it is non-NULL.                                                               */
int nbmf_vb_argmax_labels(nbmf_handle *h, int32_t *Z_out);

/* ---- lower_bound_aspect (collapsed_gibbs_z_VB.cpp:339-369) -----------------
 * ELBO of the current statistics' parameters with q(Z) at its optimum, and its
 * seven-term split (pW,pZ+pV-qZ,pH,0,qW,0,qH) for diagnostics.                */
int nbmf_lower_bound(nbmf_handle *h, int K, double alpha, double beta, double gamma, double *lb,
                     double *terms7);

/* ---- VB_DirBer_Rcpp (collapsed_gibbs_z_VB.cpp:97-168) ----------------------
 * Kc = K+1 columns in E_W (F x (K+1)) and E_H (N x (K+1)); iter-1 sweeps.
 * EXACT_WAVEFRONT needs the labels (call nbmf_init_from_labels with Kc = K+1
 * first; the engine keeps them) and materialises E_Z on the device.           */
int nbmf_vb_dirber(nbmf_handle *h, int K, double alpha, double beta, double gamma, int iter,
                   int mode, double *E_W, double *E_H, double *E_Z);

/* ---- Gibbs (collapsed_gibbs_z.cpp:363-437 + :565-605; R/MCEM - sample_gibbs.R
 * :40-77) -- uncollapsed blocked sweep  Z|W,H -> counts -> W|Z, H|Z,V.
 * iter-1 sweeps, kept when i >= floor(iter*burnin).  E_W / E_H are the
 * Rao-Blackwell posterior means of expectation.GibbsDirBer
 * (R/generic_Gibbs_DirBer.R:54-101).  test_idx: n_test linear indices
 * f + F*n of held-out entries; E_V_test[n_test] receives mean_j (E_W E_H^T)
 * there.  Z_last (F x N) receives the last assignment.                        */
int nbmf_gibbs_dirber(nbmf_handle *h, int K, double alpha, double beta, double gamma, int iter,
                      double burnin, uint64_t seed, double *E_W, double *E_H,
                      const int64_t *test_idx, int64_t n_test, double *E_V_test, int32_t *Z_last);

/* ---- Gibbs_DirBer_finite_Rcpp with its own (collapsed) schedule -------------
 * The sequential sweep of collapsed_gibbs_z.cpp:392-427 run as an anti-diagonal
 * wavefront (entries on a diagonal touch disjoint count rows, and the order of
 * every conflicting pair is the reference's).  R's global stream is replaced by
 * a uniform keyed (seed; f, n, sweep), so the result does not depend on the
 * traversal order: labels and Z_samples equal the reference loop run with the
 * same key bit for bit (oracle_gibbs_dirber_finite_keyed).  Needs the labels
 * (nbmf_init_from_labels with Kc = K).  Z_samples: F x N x (iter -
 * floor(iter*burnin)) int32, zero-filled, slices written as in :423-426, or
 * NULL; Z_last F x N or NULL.  Replicas only; sized for configs the reference
 * can run (labels and Z_samples live on the device).                          */
int nbmf_gibbs_dirber_collapsed(nbmf_handle *h, int K, double alpha, double beta, double gamma, int iter,
                                double burnin, uint64_t seed, int32_t *Z_samples, int *n_samples_written,
                                int32_t *Z_last);

/* ---- Gibbs_DirBer_DP_Rcpp (collapsed_gibbs_z.cpp:275-359) and Gibbs_DirDir_finite_Rcpp
 * (:441-557), the sibling collapsed samplers, same wavefront + keyed-uniform scheme and the same
 * bit-exact parity bar (oracle_gibbs_dirber_dp_keyed / oracle_gibbs_dirdir_keyed).
 * DP: Kmax = 200 components (:282) -- install the labels with nbmf_init_from_labels(Z, 200);
 * alpha and beta are ignored as in the reference (its Beta prior is epsilon = 1e-8, :281);
 * V must be fully observed (the reference has no NA test there, :308-310).
 * Dir-Dir: a second assignment matrix C (initialised to Z, :465); C_samples / C_last like the
 * Z outputs; K >= 2.                                                              */
int nbmf_gibbs_dirber_dp_collapsed(nbmf_handle *h, double alpha, double beta, double gamma, int iter,
                                   double burnin, uint64_t seed, int32_t *Z_samples, int *n_samples_written,
                                   int32_t *Z_last);
int nbmf_gibbs_dirdir_collapsed(nbmf_handle *h, int K, double alpha, double gamma, int iter, double burnin,
                                uint64_t seed, int32_t *Z_samples, int32_t *C_samples, int *n_samples_written,
                                int32_t *Z_last, int32_t *C_last);

/* one column's sweep given a fixed dictionary (sample_gibbs.cpp:22-80,
 * R sample_gibbs_ZH): v_n[F] in {0,1}, W F x K (paper orientation: Beta means),
 * returns the mean one-hot assignment C_mean F x K over kept sweeps.          */
int nbmf_gibbs_sweep_given_W(nbmf_handle *h, const int32_t *v_n, const double *W, int64_t F,
                             int K, double alpha, int iter, double burnin, uint64_t seed,
                             double *C_mean);

/* ---- loglikelihood.Bernoulli (R/commons.R:17-33) ---------------------------
 * sum over non-NA entries of V_test of log(v p + (1-v)(1-p)), p = E_W E_H^T,
 * without forming F x N.                                                      */
int nbmf_predictive_ll(nbmf_handle *h, const double *E_W, const double *E_H, int K,
                       const int32_t *V_test_colmajor, int64_t F, int64_t N, double *ll,
                       int64_t *n_test);

/* ---- posterior-mean reducers of the collapsed samplers ------------------------
 * compute_expectation_W_Rcpp (collapsed_gibbs_z.cpp:178-197) and
 * compute_expectation_H_Rcpp (:230-272) on the device.  Z_samples: F x N x J int32
 * (arma::icube layout), NA = INT_MIN; Kq = max label + 1 (the reference sizes its
 * outputs by Z_samples.max(), :180,236); E_W F x Kq, E_H N x Kq with the last column
 * left at 0 as the reference's k < K loop leaves it (:236-242).  Bit-identical to the
 * reference (same order of the sum over samples).                                  */
int nbmf_compute_expectation_W(nbmf_handle *h, const int32_t *Z_samples, int64_t F, int64_t N, int J,
                               double gamma, int Kq, double *E_W);
int nbmf_compute_expectation_H(nbmf_handle *h, const int32_t *Z_samples, const int32_t *V_colmajor,
                               int64_t F, int64_t N, int J, double alpha, double beta, int Kq,
                               double *E_H);

/* ---- samples_VWH_DirBer (collapsed_gibbs_z.cpp:565-625) ------------------------
 * for every kept label matrix: W_f ~ Dir(gamma + n_total_f) (normalised Gammas),
 * H_nk ~ Beta(alpha + ones, beta + zeros), V ~ Bernoulli(W H^T); R's stream is
 * replaced by Philox keyed (seed; sample, row / column, component).  K = max label + 1
 * (:573).  W_samples F x K x J, H_samples N x K x J, V_samples F x N x J; any may be
 * NULL.                                                                            */
int nbmf_samples_VWH_DirBer(nbmf_handle *h, const int32_t *V_colmajor, const int32_t *Z_samples,
                            int64_t F, int64_t N, int J, int K, double gamma, double alpha,
                            double beta, uint64_t seed, double *W_samples, double *H_samples,
                            int32_t *V_samples);

/* ---- lower_bound_aspect with its eleven arguments (collapsed_gibbs_z_VB.cpp:339-369)
 * for sizes where the E_Z cube (F x K x N, arma::cube layout) exists: the seven terms
 * as the reference states them (:173-294).  terms7 = pW, pZ, pH, pV, qW, qZ, qH (may
 * be NULL); *lb = pW + pZ + pH + pV - qW - qZ - qH.  qZ is NaN when E_Z holds an exact
 * zero, as in the reference (:234).                                                 */
int nbmf_lower_bound_aspect_full(nbmf_handle *h, const double *E_Z, const double *E_log_W,
                                 const double *E_log_H, const double *E_log_1_H,
                                 const double *gamma_vb, const double *alpha_vb,
                                 const double *beta_vb, double alpha, double beta, double gamma,
                                 const int32_t *V_colmajor, int64_t F, int64_t N, int K, double *lb,
                                 double *terms7);

/* ---- sample_gibbs_cpp (sample_gibbs.cpp:22-80) ---------------------------------
 * one column's sweep over the one-hot matrix C (F x K int32, column-major) given W
 * (paper orientation), h collapsed.  as_coded != 0: the reference as written -- its
 * draw is never stored (:60-69), so the result is C / (number of kept states);
 * as_coded == 0: the draw is stored (R sample_gibbs_Z, R/MCEM - sample_gibbs.R:1-37).
 * C_mean F x K.                                                                     */
int nbmf_sample_gibbs_cpp(nbmf_handle *h, const int32_t *v_n, const double *W, const int32_t *C,
                          int64_t F, int K, double alpha, int iter, double burnin, uint64_t seed,
                          int as_coded, double *C_mean);

/* ---- collectives inside the library ("multi-GPU is internal to the
 * handle") -------------------------------------------------------------------------
 * One process per GPU: rank 0 fills id128 (128 bytes) with nbmf_comm_unique_id, ships it
 * to the other ranks, every rank calls nbmf_comm_init_rank on its handle (created with
 * nbmf_opts.n_global / n_offset describing its column block).  From then on the
 * F x K Dirichlet-side statistic (int32 row counts for Gibbs) is summed over the ranks
 * by the library itself (NCCL over NVLink), on a priority side stream beside the cols
 * pass; the hook of nbmf_set_allreduce_hook is not consulted.                        */
int nbmf_comm_unique_id(void *id128);
int nbmf_comm_init_rank(nbmf_handle *h, const void *id128, int world, int rank);
int nbmf_comm_destroy(nbmf_handle *h);
int nbmf_comm_info(const nbmf_handle *h, int *rank, int *world);

/* One process, several GPUs -- the reference's single R process
 * (src/RcppExports.cpp:259-274): one handle, one NCCL rank and one host thread per
 * device; the columns of V are cut into contiguous blocks (boundaries at multiples of
 * 64; N >= 64 * n_devices).  opts->device is ignored, everything else applies to every
 * handle.  The nbmf_group_* calls take and return WHOLE matrices.                    */
typedef struct nbmf_group nbmf_group;
int nbmf_group_create(nbmf_group **out, const nbmf_opts *opts, const int *devices, int n_devices);
void nbmf_group_destroy(nbmf_group *g);
int nbmf_group_size(const nbmf_group *g);
nbmf_handle *nbmf_group_handle(nbmf_group *g, int rank);
const char *nbmf_group_last_error(const nbmf_group *g);
int nbmf_group_columns(const nbmf_group *g, int rank, int64_t *n_lo, int64_t *n_hi);
int nbmf_group_set_data_i32(nbmf_group *g, const int32_t *V_colmajor, int64_t F, int64_t N);
int nbmf_group_generate_synthetic(nbmf_group *g, int64_t F, int64_t N, int Ktrue, double a, double b,
                                  uint64_t seed);
int nbmf_group_init_from_labels(nbmf_group *g, const int32_t *Z_colmajor, int Kc);
int nbmf_group_init_random_labels(nbmf_group *g, int Kc, uint64_t seed);
int nbmf_group_vb_begin(nbmf_group *g, int K, double alpha, double beta, double gamma, int mean_params);
int nbmf_group_vb_step(nbmf_group *g, int checklb);
int nbmf_group_vb_end(nbmf_group *g, double *E_W, double *E_H, double *lowerbounds, int n_lb);
int nbmf_group_synchronize(nbmf_group *g);
/* VB_Aspect_Rcpp(V, Z, K, alpha, beta, gamma, iter, checklb) over all devices, one call */
int nbmf_group_vb_aspect(nbmf_group *g, const int32_t *V_colmajor, const int32_t *Z_colmajor, int64_t F,
                         int64_t N, int K, double alpha, double beta, double gamma, int iter,
                         int checklb, double *E_W, double *E_H, double *lowerbounds);
/* the same from host bit planes + host vbcounts (see nbmf_vb_aspect_bits) */
int nbmf_group_vb_aspect_bits(nbmf_group *g, const uint32_t *vbits, int64_t F, int64_t N, int K,
                              double alpha, double beta, double gamma, int iter, int checklb,
                              const double *n_total, const double *f_ones, const double *f_zeros,
                              double *E_W, double *E_H, double *lowerbounds);
int nbmf_group_gibbs_dirber(nbmf_group *g, int K, double alpha, double beta, double gamma, int iter,
                            double burnin, uint64_t seed, double *E_W, double *E_H);

#ifdef __cplusplus
}
#endif
#endif /* NBMF_CUDA_H */
