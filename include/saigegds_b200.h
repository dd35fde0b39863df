/* ===========================================================================
 * saigegds_b200.h -- C-ABI of the B200-native null-model hot path of SAIGEgds.
 *
 * This is the drop-in boundary: plain pointers and sizes, no R, Rcpp, torch or CUDA
 * types.  Every entry point names the reference routine it replaces
 * (file:line relative to the SAIGEgds v1.12.5 source tree).  An Rcpp shim with the
 * reference's exact .Call signatures sits on top of it; the Python
 * package `saigegds_b200` binds the same symbols through ctypes.
 *
 * Conventions
 *   - all matrices are column-major (R / Armadillo layout);
 *   - host pointers unless the name ends in `_device`;
 *   - every function returns SGB_OK (0) or an error code; sgb_last_error() gives the text.
 *     SGB_ERR_INVALID  <-> std::invalid_argument, SGB_ERR_OVERFLOW <-> std::overflow_error
 *     in the reference (both become an R stop() through BEGIN_RCPP/END_RCPP);
 *   - one context == one GPU == one shard of variants.  State persists between calls
 *     exactly like the file-scope statics of saige_fitnull.cpp:122-131; a second
 *     sgb_store_2b_geno() on the same context replaces the first.
 *   - there is no CPU fallback: without a CUDA device every call fails with SGB_ERR_CUDA.
 * =========================================================================== */
#ifndef SAIGEGDS_B200_H
#define SAIGEGDS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgb_context sgb_context;

enum {
    SGB_OK = 0,
    SGB_ERR_INVALID = 1,   /* std::invalid_argument in the reference */
    SGB_ERR_CUDA = 2,      /* CUDA runtime / no device */
    SGB_ERR_OVERFLOW = 3,  /* std::overflow_error: "Large variance estimate ...", "Sigma_E = 0 ..." */
    SGB_ERR_COMM = 4,      /* NCCL, or a rank that did not reach a collective within SGB_WAIT_TIMEOUT_MS */
    SGB_ERR_STATE = 5      /* call order (e.g. fit before store) */
};

enum { SGB_FAMILY_BINOMIAL = 0, SGB_FAMILY_GAUSSIAN = 1 };

/* Product kernels selectable at run time (debug / measurement); default SGB_KERNEL_AUTO. */
enum { SGB_KERNEL_AUTO = 0, SGB_KERNEL_SIMT = 1, SGB_KERNEL_IMMA = 2 /* fused single pass if possible */,
       SGB_KERNEL_IMMA_TWOPASS = 3, SGB_KERNEL_UMMA = 4 /* batched tcgen05 path even for one column */ };

/* The `param` list built at R/saige_main.r:442-453 and read at saige_fitnull.cpp:954-966. */
typedef struct {
    double tol;
    double tolPCG;
    int seed;
    int maxiter;
    int maxiterPCG;
    int no_iteration;
    int nrun;
    int num_marker;
    double traceCVcutoff;
    double ratioCVcutoff;
    int verbose;
    const char *indent; /* may be NULL */
} sgb_param;

/* The fields of R's glm object `fit0` that the native code reads (saige_fitnull.cpp:968-984). */
typedef struct {
    int64_t n;                       /* number of samples */
    int p;                           /* number of fixed-effect columns */
    const double *y;                 /* fit0$y */
    const double *offset;            /* fit0$offset or NULL */
    const double *linear_predictors; /* fit0$linear.predictors */
    const double *fitted_values;     /* fit0$fitted.values */
    const double *coefficients;      /* fit0$coefficients [p] */
    int family;                      /* SGB_FAMILY_*: fit0$family (binomial/logit or gaussian/identity) */
} sgb_fit0;

/* The list returned by saige_fit_AI_PCG_binary/_quant (saige_fitnull.cpp:1089-1096). Caller allocates. */
typedef struct {
    double *coefficients;      /* [p] */
    double tau[2];             /* Sigma_E, Sigma_G */
    double *linear_predictors; /* [n] */
    double *fitted_values;     /* [n] */
    double *residuals;         /* [n] */
    double *cov;               /* [p*p] */
    int converged;
} sgb_glmm;

/* obj.noK members read by saige_calc_var_ratio_* (saige_fitnull.cpp:1286-1291). */
typedef struct {
    int p;
    const double *X1;       /* n x p */
    const double *XV;       /* p x n */
    const double *XXVX_inv; /* n x p */
} sgb_noK;

/* The data.frame returned by saige_calc_var_ratio_* (saige_fitnull.cpp:1357-1359). Caller allocates
 * `capacity` rows; `n` rows are filled. */
typedef struct {
    int capacity;
    int n;
    int *id;
    double *maf, *mac, *var1, *var2, *ratio;
} sgb_var_ratio;

/* The model list of saige_score_test_init (src/saige_main.cpp:101-155), i.e. the arrays .init_nullmod builds
 * (R/assoc_single.r:17-67).  "K x n" matrices are in R's column-major order, so sample i's K values are contiguous. */
typedef struct {
    int trait;                  /* 0 = binary (saige_score_test_bin), 1 = quantitative (saige_score_test_quant) */
    int64_t n;                  /* samples */
    int K;                      /* fixed-effect columns, <= 32 */
    const double *tau;          /* [2] */
    const double *y, *mu, *y_mu, *mu2;                 /* [n] */
    const double *t_XXVX_inv, *XV;                     /* K x n; accepted for interface parity, not read: the dense
                                                          branch they serve (:252-262) is folded into the other one */
    const double *t_XVX_inv_XV, *t_X;                  /* K x n */
    const double *XVX;                                 /* K x K */
    const double *S_a;                                 /* [K] */
    double var_ratio;
} sgb_score_model;

/* ---- life cycle ------------------------------------------------------------------------- */
int sgb_ctx_create(sgb_context **out, int device_ordinal);
int sgb_ctx_destroy(sgb_context *ctx);
const char *sgb_last_error(void);
/* Route verbose output (Rprintf in the reference) and R RNG draws through the host language.
 * print_fn == NULL -> stdout.  rademacher_fn == NULL -> built-in restatement of R's
 * set.seed()/rbinom(n,1,0.5) under RNGkind("Mersenne-Twister","Inversion","Rounding")
 * (saige_fitnull.cpp:109-114, :649).  rademacher_fn(user, reseed, seed, n, out) must fill out[n]
 * with 0/1 draws, calling set.seed(seed) first when reseed != 0. */
int sgb_set_callbacks(sgb_context *ctx, void (*print_fn)(const char *),
                      void (*rademacher_fn)(void *user, int reseed, int seed, int64_t n, int8_t *out), void *user);
int sgb_set_kernel(sgb_context *ctx, int kernel);

/* ---- multi-GPU: one process per GPU, variants sharded, N-vectors replicated --------------- */
/* 128-byte NCCL unique id created on rank 0 and distributed by the host language
 * (torch.distributed / MPI / files).  Replaces nothing in the reference (it is single-process);
 * the analogue is NumThreads at saige_fitnull.cpp:174-177. */
int sgb_comm_unique_id(unsigned char id[128]);
int sgb_comm_init(sgb_context *ctx, const unsigned char id[128], int rank, int world_size);

/* ---- saige_store_2b_geno (saige_fitnull.cpp:159-230) --------------------------------------- */
/* packed: this rank's variants, one column of n_bytes_per_variant = ceil(n_samp/4) bytes per variant,
 * sample 4j+k in bits 2k..2k+1 of byte j, 0/1/2 = dosage, 3 = missing.  The matrix is copied to the
 * device (the reference keeps a raw pointer into R memory).  buf_std_geno [4*n_variant_local] and
 * buf_diag_grm [n_samp] receive the same values the reference writes into r_buf_geno / r_buf_sigma
 * (either may be NULL).  n_variant_total / variant_offset describe the shard (== n_variant_local / 0
 * on one GPU).  r_buf_crossprod has no equivalent: its only role was to carry num.thread. */
int sgb_store_2b_geno(sgb_context *ctx, const uint8_t *packed, int64_t n_samp, int64_t n_bytes_per_variant,
                      int64_t n_variant_local, int64_t n_variant_total, int64_t variant_offset,
                      double *buf_std_geno, double *buf_diag_grm);
/* Same, but `packed_device` already lives in device memory (synthetic benchmarks generate it there).
 * take_ownership != 0: the context adopts the allocation (must come from cudaMalloc). */
int sgb_store_2b_geno_device(sgb_context *ctx, uint8_t *packed_device, int take_ownership, int64_t n_samp,
                             int64_t n_bytes_per_variant, int64_t n_variant_local, int64_t n_variant_total,
                             int64_t variant_offset, double *buf_std_geno, double *buf_diag_grm);

/* ---- sparse genotypes: the reference's default geno.sparse=TRUE ----------------------------- */
#define SGB_GENO_RAW 0  /* RAWSXP: one byte per sample, 0/1/2, anything else = missing */
#define SGB_GENO_INT 1  /* INTSXP: 0/1/2, anything else (NA_integer_) = missing        */
#define SGB_GENO_REAL 2 /* REALSXP: dosage rounded to 0/1/2, non-finite = missing     */
/* saige_get_sparse (saige_fitnull.cpp:252-320): one variant's genotypes -> the integer vector
 * (n1, n2, n3, indices of 1s, indices of 2s, indices of missing), 0-based, counted on the minor allele
 * (codes are flipped 0<->2 when the coded allele is the major one, :295-303).  `out` needs n_samp + 3
 * ints -- the buffer saige_init_sparse (:244-249) registers; *out_len receives the used length.
 * Host only: no context, no device.  Unlike the reference it does not overwrite `geno`. */
int sgb_get_sparse(const void *geno, int geno_type, int64_t n_samp, int32_t *out, int64_t *out_len);
/* saige_store_sp_geno (saige_fitnull.cpp:324-388).  The R list of integer vectors arrives flattened:
 * variant j's vector is sp_data[sp_offsets[j] .. sp_offsets[j+1]).  The lists are packed to the 2-bit layout on
 * the host (slab-wise, pinned staging) and stored like sgb_store_2b_geno, so every later call (products, fits,
 * variance ratio, sgb_get_geno_ds) is the same device path.  buf_std_geno receives the reference's SPARSE table
 * (entries 1..3 relative to entry 0, :358), bit-identical; buf_diag_grm as :363-385.  Malformed vectors
 * (counts not matching the length, index outside [0, n_samp)) -> SGB_ERR_INVALID (the reference reads out of bounds). */
int sgb_store_sp_geno(sgb_context *ctx, const int32_t *sp_data, const int64_t *sp_offsets, int64_t n_samp,
                      int64_t n_variant_local, int64_t n_variant_total, int64_t variant_offset,
                      double *buf_std_geno, double *buf_diag_grm);
/* The host packing step of sgb_store_sp_geno on its own: packed [n_variant][ceil(n_samp/4)], pad samples = 3. */
int sgb_sparse_to_packed(const int32_t *sp_data, const int64_t *sp_offsets, int64_t n_samp, int64_t n_variant,
                         uint8_t *packed);
/* Integer results of the LUT pass (n_valid, sum at saige_fitnull.cpp:187-192), bit-exact. */
int sgb_allele_counts(sgb_context *ctx, int32_t *n_valid, int32_t *sum);
/* get_geno_ds (saige_fitnull.cpp:394-427): dosage of local variant snp_idx, missing -> NaN. */
int sgb_get_geno_ds(sgb_context *ctx, int64_t snp_idx, double *ds);

/* ---- get_crossprod_b_grm (saige_fitnull.cpp:435-536) ---------------------------------------- */
/* out[:,c] = (1/M) G_std G_std' b[:,c] for k right-hand sides (n_samp x k, column-major). */
int sgb_grm_mv(sgb_context *ctx, const double *b, double *out, int k);
int sgb_grm_mv_device(sgb_context *ctx, const double *b_device, double *out_device, int k);

/* ---- get_diag_sigma / PCG_diag_sigma (saige_fitnull.cpp:542-614) ----------------------------- */
int sgb_diag_sigma(sgb_context *ctx, const double *w, const double tau[2], double *out);
/* k independent solves Sigma x = b in lock-step; each column keeps its own alpha/beta and stops on
 * its own sum(r*r) <= tolPCG test, so every column's iterates equal a stand-alone solve. iters[k]. */
int sgb_pcg(sgb_context *ctx, const double *w, const double tau[2], const double *b, int k, int maxiterPCG,
            double tolPCG, double *x, int *iters);

/* ---- saige_fit_AI_PCG_binary / _quant (saige_fitnull.cpp:949-1099, :1103-1248) -------------- */
int sgb_fit_AI_PCG_binary(sgb_context *ctx, const sgb_fit0 *fit0, const double *X, const double tau[2],
                          const sgb_param *param, sgb_glmm *out);
int sgb_fit_AI_PCG_quant(sgb_context *ctx, const sgb_fit0 *fit0, const double *X, const double tau[2],
                         const sgb_param *param, sgb_glmm *out);

/* ---- saige_calc_var_ratio_binary / _quant (saige_fitnull.cpp:1255-1362, :1366-1474) --------- */
/* marker_list: 1-based indices into the GLOBAL variant order (sample.int(n_var, n_var) in R). */
int sgb_calc_var_ratio_binary(sgb_context *ctx, const sgb_fit0 *fit0, const double tau[2], const sgb_noK *noK,
                              const sgb_param *param, const int32_t *marker_list, int64_t n_marker,
                              sgb_var_ratio *out);
int sgb_calc_var_ratio_quant(sgb_context *ctx, const sgb_fit0 *fit0, const double tau[2], const sgb_noK *noK,
                             const sgb_param *param, const int32_t *marker_list, int64_t n_marker,
                             sgb_var_ratio *out);

/* ---- genotype ingestion from a GDS genotype node (device half) -------------------------------------- */
/* Replaces SeqArray:::.seqGet2bGeno (call site R/saige_main.r:420) and the seqSetFilterCond(maf=, missing.rate=) call at
 * :319 for integer genotypes.  allele_bits: the decompressed bytes of SeqArray's `genotype/data` node -- bit2 allele
 * indices [variant][sample][ploidy = 2], four per byte, no padding between variants: sample s of variant v is nibble
 * v * n_samp_file + s (allele 1 in bits 0-1, allele 2 in bits 2-3; 0 = reference, 1 / 2 = alternative, 3 = missing).
 * sample_sel (NULL = all): ascending indices of the n_samp samples of the model.  On the device each nibble becomes a 2-bit
 * alt-allele dosage (missing if either allele is), alleles are counted per variant (integer, order independent), variants
 * with MAF >= maf and missing rate <= missing_rate are kept (NaN: no filter) and stored like sgb_store_2b_geno.
 * Outputs: variant_sel[n_variant_file] 0/1, *n_variant = number kept, optional allele counts [n_variant_file],
 * buf_std_geno [4 * n_variant_file is always enough] and buf_diag_grm [n_samp] as in sgb_store_2b_geno.  One GPU. */
int sgb_store_gds_geno(sgb_context *ctx, const uint8_t *allele_bits, int64_t n_samp_file, int64_t n_variant_file,
                       const int32_t *sample_sel, int64_t n_samp, double maf, double missing_rate, int32_t *variant_sel,
                       int64_t *n_variant, int32_t *n_valid_alleles, int32_t *n_alt_alleles, double *buf_std_geno,
                       double *buf_diag_grm);

/* ---- saige_GxG_snp_bin (saige_fitnull.cpp:1480-1558): the interaction-term test of seqGLMM_GxG_spa -------------- */
typedef struct {   /* the one-row data.frame returned at :1550-1555 */
    double beta, SE, pval, p_norm, tau_G;
    int64_t n_nonzero;
    int converged;
} sgb_gxg;
/* fit0: the glm object (y, linear.predictors, fitted.values, family); tau: glmm$tau; inter_term[n]: the interaction
 * term; noK: obj.noK (X1, XV, XXVX_inv); param: tolPCG / maxiterPCG are read.  Runs on the stored genotypes. */
int sgb_GxG_snp_bin(sgb_context *ctx, const sgb_fit0 *fit0, const double tau[2], const double *inter_term,
                    const sgb_noK *noK, const sgb_param *param, int verbose, sgb_gxg *out);

/* ---- single-variant score test with saddle-point approximation (seqAssocGLMM_SPA's inner loop) ---------------- */
/* saige_score_test_init (saige_main.cpp:101-155): copies the model to the device.  maf / mac / missing / spa_pval are
 * M$maf, M$mac, M$missing, M$spa.pval; a non-finite value disables that filter (spa_pval: 0.05), as at :106-113. */
int sgb_score_test_init(sgb_context *ctx, const sgb_score_model *model, double maf, double mac, double missing,
                        double spa_pval);
/* saige_score_test_bin / saige_score_test_quant (:288-407 / :188-285) over a batch of variants -- the reference is called
 * once per variant by seqApply.  out[n_variant][8] = AF, mac, num, beta, SE, pval, p.norm, converged (the vector returned at
 * :398-406; for quantitative traits p.norm == pval and converged == 1); valid[v] == 0 where the reference returns NULL
 * (filtered variant; its out row is NaN).  Three genotype sources:
 *   _packed  2-bit codes [n_variant][ceil(n/4)] on the host (integer genotypes: 16x less PCIe traffic than doubles),
 *   _dosage  doubles [n_variant][n] on the host, non-finite = missing (REALSXP dosages, get_ds at :166-186),
 *   _stored  variants [first, first + n_variant) of the matrix already stored by sgb_store_2b_geno / _sp_geno; no copy,
 *            *kernel_ms (may be NULL) receives the CUDA-event time of the kernel. */
int sgb_score_test_packed(sgb_context *ctx, const uint8_t *packed, int64_t n_bytes_per_variant, int64_t n_variant,
                          double *out, int32_t *valid);
/* Kernel choice (after sgb_score_test_init).  SGB_SCORE_TENSOR (default for 2-bit packed genotypes): the score statistics of a
 * block of variants are three integer GEMMs on the tcgen05 tensor cores (class sums of the model columns, exact fixed point) and a
 * per-variant finishing kernel; SGB_SCORE_TILED: the shared-memory-tiled CUDA-core kernel (always used for dosages);
 * either way only the saddle-point candidates go through the per-variant kernel.  SGB_SCORE_PER_VARIANT runs every variant
 * through the latter.  Same results to rounding. */
#define SGB_SCORE_TILED 0
#define SGB_SCORE_PER_VARIANT 1
#define SGB_SCORE_TENSOR 2
int sgb_score_test_set_path(sgb_context *ctx, int path);
int sgb_score_test_dosage(sgb_context *ctx, const double *dosage, int64_t n_variant, double *out, int32_t *valid);
int sgb_score_test_stored(sgb_context *ctx, int64_t first, int64_t n_variant, double *out, int32_t *valid,
                          float *kernel_ms);

/* ---- R RNG restatement, exposed so the host can draw sample.int(n_var, n_var) (R/saige_main.r:509-511) */
int sgb_r_set_seed(sgb_context *ctx, uint32_t seed);
int sgb_r_unif_rand(sgb_context *ctx, int64_t n, double *out);
int sgb_r_sample_int(sgb_context *ctx, int32_t n, int32_t *out);

/* ---- instrumentation ------------------------------------------------------------------------ */
typedef struct {
    int64_t n_products;        /* single-RHS GRM products executed (a k-RHS call counts k) */
    int64_t n_product_launches;/* product kernel launches */
    int64_t n_kernel_launches; /* all kernels launched by this library */
    int64_t n_pcg_solves;
    int64_t n_pcg_iterations;
    double last_product_ms;    /* CUDA-event time of the most recent sgb_grm_mv*_ call, device part only */
    int64_t n_host_syncs;      /* cudaStreamSynchronize calls on the library's stream */
    double host_wait_s;        /* wall-clock seconds the host spent blocked in them (GPU busy); entry-point wall time minus this
                                  is host-side work with the GPU idle or running ahead */
} sgb_stats;
int sgb_get_stats(sgb_context *ctx, sgb_stats *out);
int sgb_reset_stats(sgb_context *ctx);
/* Generate a synthetic packed matrix directly in device memory (bench.py):
 * maf_j ~ U(0.005, 0.5), genotype ~ Binomial(2, maf_j), missing_rate as code 3; counter-based, so a
 * shard [variant_offset, variant_offset + n_variant_local) is identical however the job is split.
 * Returns a cudaMalloc'ed pointer in *packed_device. */
int sgb_synth_geno_device(sgb_context *ctx, int64_t n_samp, int64_t n_variant_local, int64_t variant_offset,
                          uint64_t seed, double missing_rate, uint8_t **packed_device);
int sgb_copy_from_device(sgb_context *ctx, void *dst_host, const void *src_device, int64_t bytes);
int sgb_free_device(sgb_context *ctx, void *ptr_device);
/* Time `reps` back-to-back device-resident products on the context's stream with CUDA events
 * (bench.py's `value` leg). b/out: device, n_samp x k. Returns total milliseconds. */
int sgb_time_products_device(sgb_context *ctx, const double *b_device, double *out_device, int k, int reps,
                             float *total_ms);
int sgb_malloc_device(sgb_context *ctx, int64_t bytes, void **ptr_device);
/* Page-locked host memory for the vectors handed to the host-pointer entry points (sgb_grm_mv, sgb_pcg, ...): the copies
 * then run at PCIe speed instead of being staged by the driver.  Any host pointer is accepted by those entry points;
 * this only makes them faster (an Rcpp shim can keep a pinned mirror of the vectors it passes per PCG iteration). */
int sgb_malloc_host(sgb_context *ctx, int64_t bytes, void **ptr_host);
int sgb_free_host(sgb_context *ctx, void *ptr_host);
/* Per-kernel CUDA-event timing of the product kernels (serialises the stream; measurement aid only).
 * sgb_kernel_times writes lines "name total_ms launches\n" into buf. */
int sgb_set_profiling(sgb_context *ctx, int on);
int sgb_kernel_times(sgb_context *ctx, char *buf, int64_t buf_size);
int sgb_copy_to_device(sgb_context *ctx, void *dst_device, const void *src_host, int64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* SAIGEGDS_B200_H */
