// extern "C" entry points of include/saigegds_b200.h: argument checks, exception -> status code.
#include <cmath>
#include <cstring>

#include "sparse_host.h"
#include "vecops.cuh"

namespace sgb {
void fit_AI_PCG(Context &c, bool quant, const sgb_fit0 *f, const double *hX, const double tau_in[2], const sgb_param *Pin,
                sgb_glmm *out);
void calc_var_ratio(Context &c, bool quant, const sgb_fit0 *f, const double tau_in[2], const sgb_noK *noK,
                    const sgb_param *Pin, const int32_t *marker_list, int64_t n_marker, sgb_var_ratio *out);
}  // namespace sgb

struct sgb_context : sgb::Context {};

namespace {

thread_local std::string g_last_error;

template <typename F>
int guarded(sgb_context *ctx, F &&fn, bool need_ctx = true) {
    try {
        if (need_ctx) {
            if (!ctx) throw sgb::Error(SGB_ERR_INVALID, "context is NULL");
            SGB_CUDA(cudaSetDevice(ctx->dev));
        }
        try {
            fn();
        } catch (const sgb::Error &e) {
            // A kernel with cross-CTA waits (the fused single-pass product) ran out of its wall-clock bound -- a time-sliced,
            // throttled or instrumented GPU.  The entry points are functions of their arguments, so redo the call once on the
            // kernels that have no such waits (two HBM passes); report an error only if that fails too.  Not with several ranks:
            // they could no longer agree on the sequence of collectives.
            if (e.code != sgb::SGB_INTERNAL_RETRY_NO_WAIT_KERNELS || !ctx || ctx->world > 1 || ctx->fused_disabled) throw;
            ctx->fused_disabled = true;
            ctx->printf("note: %s; redoing the call with the two-pass kernels (used from now on)\n", e.what());
            cudaStreamSynchronize(ctx->stream);
            fn();
        }
        return SGB_OK;
    } catch (const sgb::Error &e) {
        g_last_error = e.what();
        return e.code == sgb::SGB_INTERNAL_RETRY_NO_WAIT_KERNELS ? SGB_ERR_CUDA : e.code;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return SGB_ERR_INVALID;
    }
}

void store_common(sgb::Context &c, int64_t n_samp, int64_t nb, int64_t m_local, int64_t m_total, int64_t offset) {
    if (n_samp < 1 || m_local < 1) throw sgb::Error(SGB_ERR_INVALID, "empty genotype matrix");
    if (nb != (n_samp + 3) / 4) throw sgb::Error(SGB_ERR_INVALID, "n_bytes_per_variant must equal ceil(n_samp/4)");
    if (m_total < m_local || offset < 0 || offset + m_local > m_total)
        throw sgb::Error(SGB_ERR_INVALID, "inconsistent variant shard description");
    if (c.world == 1 && m_total != m_local) throw sgb::Error(SGB_ERR_INVALID, "sharded store needs sgb_comm_init first");
    sgb::imma_release(c);
    c.stored = false;
    c.N = n_samp; c.NB = nb; c.M = m_local; c.M_total = m_total; c.var_offset = offset;
}

void store_outputs(sgb::Context &c, double *buf_std_geno, double *buf_diag_grm) {
    if (buf_std_geno) c.d2h(buf_std_geno, c.lut.get(), sizeof(double) * 4 * c.M);
    if (buf_diag_grm) c.d2h(buf_diag_grm, c.diag.get(), sizeof(double) * c.N);
    c.sync();
    c.stored = true;
    sgb::imma_prepare(c);
}

}  // namespace

extern "C" {

const char *sgb_last_error(void) { return g_last_error.c_str(); }

int sgb_ctx_create(sgb_context **out, int device_ordinal) {
    return guarded(nullptr, [&] {
        if (!out) throw sgb::Error(SGB_ERR_INVALID, "out is NULL");
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw sgb::Error(SGB_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                               "); this library has no CPU fallback");
        if (device_ordinal < 0 || device_ordinal >= n) throw sgb::Error(SGB_ERR_INVALID, "device ordinal out of range");
        SGB_CUDA(cudaSetDevice(device_ordinal));
        sgb_context *c = new sgb_context();
        c->dev = device_ordinal;
        cudaDeviceProp prop;
        SGB_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
        c->sm_count = prop.multiProcessorCount;
        SGB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        SGB_CUDA(cudaEventCreate(&c->ev0));
        SGB_CUDA(cudaEventCreate(&c->ev1));
        SGB_CUDA(cudaEventCreate(&c->pev0));
        SGB_CUDA(cudaEventCreate(&c->pev1));
        *out = c;
    }, false);
}

int sgb_ctx_destroy(sgb_context *ctx) {
    return guarded(ctx, [&] {
        cudaStreamSynchronize(ctx->stream);
        sgb::imma_release(*ctx);
        sgb::score_release(*ctx);
        sgb::solver_release(*ctx);
        sgb::comm_destroy(*ctx);
        cudaEventDestroy(ctx->ev0);
        cudaEventDestroy(ctx->ev1);
        cudaEventDestroy(ctx->pev0);
        cudaEventDestroy(ctx->pev1);
        cudaStreamDestroy(ctx->stream);
        delete ctx;
    });
}

int sgb_set_callbacks(sgb_context *ctx, void (*print_fn)(const char *),
                      void (*rademacher_fn)(void *, int, int, int64_t, int8_t *), void *user) {
    return guarded(ctx, [&] { ctx->print_fn = print_fn; ctx->rademacher_fn = rademacher_fn; ctx->cb_user = user; });
}

int sgb_set_kernel(sgb_context *ctx, int kernel) {
    return guarded(ctx, [&] {
        if (kernel < SGB_KERNEL_AUTO || kernel > SGB_KERNEL_UMMA) throw sgb::Error(SGB_ERR_INVALID, "unknown kernel id");
        ctx->kernel = kernel;
    });
}

int sgb_comm_unique_id(unsigned char id[128]) {
    return guarded(nullptr, [&] { sgb::comm_unique_id(id); }, false);
}

int sgb_comm_init(sgb_context *ctx, const unsigned char id[128], int rank, int world_size) {
    return guarded(ctx, [&] { sgb::comm_init(*ctx, id, rank, world_size); });
}

int sgb_store_2b_geno(sgb_context *ctx, const uint8_t *packed, int64_t n_samp, int64_t n_bytes_per_variant,
                      int64_t n_variant_local, int64_t n_variant_total, int64_t variant_offset, double *buf_std_geno,
                      double *buf_diag_grm) {
    return guarded(ctx, [&] {
        if (!packed) throw sgb::Error(SGB_ERR_INVALID, "packed is NULL");
        store_common(*ctx, n_samp, n_bytes_per_variant, n_variant_local, n_variant_total, variant_offset);
        sgb::DevBuf<uint8_t> raw;
        const size_t bytes = (size_t)n_bytes_per_variant * n_variant_local;
        raw.ensure(bytes);
        ctx->h2d(raw.get(), packed, bytes);
        sgb::store_device_layout(*ctx, raw.get(), (size_t)n_bytes_per_variant);
        store_outputs(*ctx, buf_std_geno, buf_diag_grm);
    });
}

int sgb_store_2b_geno_device(sgb_context *ctx, uint8_t *packed_device, int take_ownership, int64_t n_samp,
                             int64_t n_bytes_per_variant, int64_t n_variant_local, int64_t n_variant_total,
                             int64_t variant_offset, double *buf_std_geno, double *buf_diag_grm) {
    return guarded(ctx, [&] {
        if (!packed_device) throw sgb::Error(SGB_ERR_INVALID, "packed_device is NULL");
        // with take_ownership the buffer is released on every path, also when the shard description is rejected or the store fails
        struct Guard { uint8_t *p; ~Guard() { if (p) cudaFree(p); } } guard{take_ownership ? packed_device : nullptr};
        store_common(*ctx, n_samp, n_bytes_per_variant, n_variant_local, n_variant_total, variant_offset);
        sgb::store_device_layout(*ctx, packed_device, (size_t)n_bytes_per_variant);
        if (take_ownership) { guard.p = nullptr; SGB_CUDA(cudaFree(packed_device)); }
        store_outputs(*ctx, buf_std_geno, buf_diag_grm);
    });
}

int sgb_get_sparse(const void *geno, int geno_type, int64_t n_samp, int32_t *out, int64_t *out_len) {
    return guarded(nullptr, [&] {
        if (geno_type < SGB_GENO_RAW || geno_type > SGB_GENO_REAL) throw sgb::Error(SGB_ERR_INVALID, "Invalid data type.");
        if (!geno || !out || !out_len || n_samp < 0) throw sgb::Error(SGB_ERR_INVALID, "invalid arguments");
        *out_len = sgb::get_sparse_host(geno, geno_type, n_samp, out);
    }, false);
}

int sgb_sparse_to_packed(const int32_t *sp_data, const int64_t *sp_offsets, int64_t n_samp, int64_t n_variant,
                         uint8_t *packed) {
    return guarded(nullptr, [&] {
        if (!sp_data || !sp_offsets || !packed || n_samp < 1 || n_variant < 0) throw sgb::Error(SGB_ERR_INVALID, "invalid arguments");
        try {
            sgb::sparse_to_packed_host(sp_data, sp_offsets, 0, n_variant, n_samp, packed);
        } catch (const std::string &e) {
            throw sgb::Error(SGB_ERR_INVALID, e);
        }
    }, false);
}

int sgb_store_sp_geno(sgb_context *ctx, const int32_t *sp_data, const int64_t *sp_offsets, int64_t n_samp,
                      int64_t n_variant_local, int64_t n_variant_total, int64_t variant_offset, double *buf_std_geno,
                      double *buf_diag_grm) {
    return guarded(ctx, [&] {
        if (!sp_data || !sp_offsets) throw sgb::Error(SGB_ERR_INVALID, "sparse genotype list is NULL");
        const int64_t nb = (n_samp + 3) / 4;
        store_common(*ctx, n_samp, nb, n_variant_local, n_variant_total, variant_offset);
        // pack on the host in slabs of ~64 MB (two pinned stages: one is packed while the other is copied)
        sgb::DevBuf<uint8_t> raw;
        raw.ensure((size_t)nb * n_variant_local);
        const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(n_variant_local, ((int64_t)64 << 20) / nb));
        sgb::PinBuf<uint8_t> stage[2];
        cudaEvent_t done[2] = {nullptr, nullptr};
        SGB_CUDA(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
        SGB_CUDA(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
        try {
            stage[0].ensure((size_t)slab * nb);
            if (slab < n_variant_local) stage[1].ensure((size_t)slab * nb);
            int s = 0;
            for (int64_t j0 = 0; j0 < n_variant_local; j0 += slab, s ^= 1) {
                const int64_t j1 = std::min(n_variant_local, j0 + slab);
                SGB_CUDA(cudaEventSynchronize(done[s]));   // the copy that last used this stage has finished
                try {
                    sgb::sparse_to_packed_host(sp_data, sp_offsets, j0, j1, n_samp, stage[s].p);
                } catch (const std::string &e) {
                    throw sgb::Error(SGB_ERR_INVALID, e);
                }
                ctx->h2d(raw.get() + (size_t)j0 * nb, stage[s].p, (size_t)(j1 - j0) * nb);
                SGB_CUDA(cudaEventRecord(done[s], ctx->stream));
            }
            ctx->sync();
        } catch (...) {
            cudaStreamSynchronize(ctx->stream);
            cudaEventDestroy(done[0]);
            cudaEventDestroy(done[1]);
            throw;
        }
        cudaEventDestroy(done[0]);
        cudaEventDestroy(done[1]);
        sgb::store_device_layout(*ctx, raw.get(), (size_t)nb);
        store_outputs(*ctx, buf_std_geno, buf_diag_grm);
        // r_buf_geno in the reference's sparse layout: entries 1..3 relative to entry 0 (saige_fitnull.cpp:358)
        if (buf_std_geno)
            for (int64_t j = 0; j < n_variant_local; j++) {
                double *p = buf_std_geno + 4 * j;
                p[1] -= p[0]; p[2] -= p[0]; p[3] -= p[0];
            }
    });
}

int sgb_store_gds_geno(sgb_context *ctx, const uint8_t *allele_bits, int64_t n_samp_file, int64_t n_variant_file,
                       const int32_t *sample_sel, int64_t n_samp, double maf, double missing_rate, int32_t *variant_sel,
                       int64_t *n_variant, int32_t *n_valid_alleles, int32_t *n_alt_alleles, double *buf_std_geno,
                       double *buf_diag_grm) {
    return guarded(ctx, [&] {
        if (!allele_bits || !variant_sel || !n_variant) throw sgb::Error(SGB_ERR_INVALID, "NULL argument");
        if (n_samp_file < 1 || n_variant_file < 1) throw sgb::Error(SGB_ERR_INVALID, "empty genotype node");
        if (!sample_sel) n_samp = n_samp_file;
        if (n_samp < 1 || n_samp > n_samp_file) throw sgb::Error(SGB_ERR_INVALID, "invalid number of selected samples");
        if (sample_sel)
            for (int64_t i = 0; i < n_samp; i++)
                if (sample_sel[i] < 0 || sample_sel[i] >= n_samp_file) throw sgb::Error(SGB_ERR_INVALID, "sample index out of range");
        if (ctx->world > 1) throw sgb::Error(SGB_ERR_INVALID, "sgb_store_gds_geno stores an unsharded matrix (one GPU)");
        const int64_t nb = (n_samp + 3) / 4, mf = n_variant_file;
        sgb::DevBuf<uint8_t> all;
        sgb::DevBuf<int32_t> d_valid, d_alt;
        all.ensure((size_t)nb * mf); d_valid.ensure(mf); d_alt.ensure(mf);
        sgb::gds_to_dosage(*ctx, allele_bits, n_samp_file, mf, sample_sel, n_samp, all.get(), d_valid.get(), d_alt.get());
        std::vector<int32_t> hv(mf), ha(mf);
        ctx->d2h(hv.data(), d_valid.get(), sizeof(int32_t) * mf);
        ctx->d2h(ha.data(), d_alt.get(), sizeof(int32_t) * mf);
        ctx->sync();
        // seqSetFilterCond(maf=, missing.rate=) as called at R/saige_main.r:319 -- allele based, integer counts
        std::vector<int64_t> rows;
        for (int64_t v = 0; v < mf; v++) {
            const double af = hv[v] > 0 ? (double)ha[v] / hv[v] : NAN;
            const double mf_v = std::min(af, 1 - af), miss = 1.0 - (double)hv[v] / (2.0 * n_samp);
            bool keep = true;
            if (std::isfinite(maf)) keep = keep && (mf_v >= maf);
            if (std::isfinite(missing_rate)) keep = keep && (miss <= missing_rate);
            variant_sel[v] = keep ? 1 : 0;
            if (keep) rows.push_back(v);
        }
        if (n_valid_alleles) memcpy(n_valid_alleles, hv.data(), sizeof(int32_t) * mf);
        if (n_alt_alleles) memcpy(n_alt_alleles, ha.data(), sizeof(int32_t) * mf);
        *n_variant = (int64_t)rows.size();
        if (rows.empty()) throw sgb::Error(SGB_ERR_INVALID, "no variant passes the MAF / missing-rate filter");
        const int64_t m = (int64_t)rows.size();
        store_common(*ctx, n_samp, nb, m, m, 0);
        sgb::DevBuf<uint8_t> raw;
        raw.ensure((size_t)nb * m);
        sgb::gather_rows(*ctx, all.get(), rows.data(), m, nb, raw.get());
        all.release();
        sgb::store_device_layout(*ctx, raw.get(), (size_t)nb);
        store_outputs(*ctx, buf_std_geno, buf_diag_grm);
    });
}

int sgb_allele_counts(sgb_context *ctx, int32_t *n_valid, int32_t *sum) {
    return guarded(ctx, [&] {
        ctx->require_stored();
        if (n_valid) ctx->d2h(n_valid, ctx->n_valid.get(), sizeof(int32_t) * ctx->M);
        if (sum) ctx->d2h(sum, ctx->sum.get(), sizeof(int32_t) * ctx->M);
        ctx->sync();
    });
}

int sgb_get_geno_ds(sgb_context *ctx, int64_t snp_idx, double *ds) {
    return guarded(ctx, [&] {
        ctx->require_stored();
        if (snp_idx < 0 || snp_idx >= ctx->M || !ds) throw sgb::Error(SGB_ERR_INVALID, "variant index out of range");
        ctx->ws_vec.ensure(ctx->N);
        sgb::decode_variant(*ctx, snp_idx, ctx->ws_vec.get());
        ctx->d2h(ds, ctx->ws_vec.get(), sizeof(double) * ctx->N);
        ctx->sync();
    });
}

int sgb_grm_mv_device(sgb_context *ctx, const double *b_device, double *out_device, int k) {
    return guarded(ctx, [&] {
        if (!b_device || !out_device || k < 1) throw sgb::Error(SGB_ERR_INVALID, "invalid arguments");
        SGB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        sgb::grm_mv_device(*ctx, b_device, out_device, k);
        SGB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        ctx->sync();
        float ms = 0;
        SGB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->stats.last_product_ms = ms;
    });
}

int sgb_grm_mv(sgb_context *ctx, const double *b, double *out, int k) {
    return guarded(ctx, [&] {
        ctx->require_stored();
        if (!b || !out || k < 1) throw sgb::Error(SGB_ERR_INVALID, "invalid arguments");
        const size_t n = (size_t)ctx->N * k;
        sgb::DevBuf<double> &db = ctx->io_in, &dout = ctx->io_out;
        db.ensure(n); dout.ensure(n);
        ctx->h2d(db.get(), b, sizeof(double) * n);
        SGB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        sgb::grm_mv_device(*ctx, db.get(), dout.get(), k);
        SGB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        ctx->d2h(out, dout.get(), sizeof(double) * n);
        ctx->sync();
        float ms = 0;
        SGB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->stats.last_product_ms = ms;
    });
}

int sgb_diag_sigma(sgb_context *ctx, const double *w, const double tau[2], double *out) {
    return guarded(ctx, [&] {
        if (!w || !tau || !out) throw sgb::Error(SGB_ERR_INVALID, "w, tau or out is NULL");
        ctx->require_stored();
        sgb::DevBuf<double> dw, dout;
        dw.ensure(ctx->N); dout.ensure(ctx->N);
        ctx->h2d(dw.get(), w, sizeof(double) * ctx->N);
        sgb::diag_sigma(*ctx, dw.get(), tau[0], tau[1], dout.get(), false);
        ctx->d2h(out, dout.get(), sizeof(double) * ctx->N);
        ctx->sync();
    });
}

int sgb_pcg(sgb_context *ctx, const double *w, const double tau[2], const double *b, int k, int maxiterPCG, double tolPCG,
            double *x, int *iters) {
    return guarded(ctx, [&] {
        ctx->require_stored();
        if (!w || !tau || !b || !x || k < 1) throw sgb::Error(SGB_ERR_INVALID, "invalid arguments");
        const size_t n = (size_t)ctx->N * k;
        sgb::DevBuf<double> dw, db, dx;
        dw.ensure(ctx->N); db.ensure(n); dx.ensure(n);
        ctx->h2d(dw.get(), w, sizeof(double) * ctx->N);
        ctx->h2d(db.get(), b, sizeof(double) * n);
        sgb::PcgWork ws;
        sgb::pcg_solve(*ctx, ws, dw.get(), tau[0], tau[1], db.get(), k, maxiterPCG, tolPCG, dx.get(), iters);
        ctx->d2h(x, dx.get(), sizeof(double) * n);
        ctx->sync();
    });
}

int sgb_fit_AI_PCG_binary(sgb_context *ctx, const sgb_fit0 *fit0, const double *X, const double tau[2], const sgb_param *param,
                          sgb_glmm *out) {
    return guarded(ctx, [&] { sgb::fit_AI_PCG(*ctx, false, fit0, X, tau, param, out); });
}

int sgb_fit_AI_PCG_quant(sgb_context *ctx, const sgb_fit0 *fit0, const double *X, const double tau[2], const sgb_param *param,
                         sgb_glmm *out) {
    return guarded(ctx, [&] { sgb::fit_AI_PCG(*ctx, true, fit0, X, tau, param, out); });
}

int sgb_calc_var_ratio_binary(sgb_context *ctx, const sgb_fit0 *fit0, const double tau[2], const sgb_noK *noK,
                              const sgb_param *param, const int32_t *marker_list, int64_t n_marker, sgb_var_ratio *out) {
    return guarded(ctx, [&] { sgb::calc_var_ratio(*ctx, false, fit0, tau, noK, param, marker_list, n_marker, out); });
}

int sgb_calc_var_ratio_quant(sgb_context *ctx, const sgb_fit0 *fit0, const double tau[2], const sgb_noK *noK,
                             const sgb_param *param, const int32_t *marker_list, int64_t n_marker, sgb_var_ratio *out) {
    return guarded(ctx, [&] { sgb::calc_var_ratio(*ctx, true, fit0, tau, noK, param, marker_list, n_marker, out); });
}

int sgb_GxG_snp_bin(sgb_context *ctx, const sgb_fit0 *fit0, const double tau[2], const double *inter_term, const sgb_noK *noK,
                    const sgb_param *param, int verbose, sgb_gxg *out) {
    return guarded(ctx, [&] { sgb::gxg_snp_bin(*ctx, fit0, tau, inter_term, noK, param, verbose, out); });
}

int sgb_score_test_init(sgb_context *ctx, const sgb_score_model *model, double maf, double mac, double missing,
                        double spa_pval) {
    return guarded(ctx, [&] { sgb::score_init(*ctx, model, maf, mac, missing, spa_pval); });
}
int sgb_score_test_set_path(sgb_context *ctx, int path) {
    return guarded(ctx, [&] { sgb::score_set_path(*ctx, path); });
}
int sgb_score_test_packed(sgb_context *ctx, const uint8_t *packed, int64_t n_bytes_per_variant, int64_t n_variant, double *out,
                          int32_t *valid) {
    return guarded(ctx, [&] { sgb::score_test_packed(*ctx, packed, n_bytes_per_variant, n_variant, out, valid); });
}
int sgb_score_test_dosage(sgb_context *ctx, const double *dosage, int64_t n_variant, double *out, int32_t *valid) {
    return guarded(ctx, [&] { sgb::score_test_dosage(*ctx, dosage, n_variant, out, valid); });
}
int sgb_score_test_stored(sgb_context *ctx, int64_t first, int64_t n_variant, double *out, int32_t *valid, float *kernel_ms) {
    return guarded(ctx, [&] { sgb::score_test_stored(*ctx, first, n_variant, out, valid, kernel_ms); });
}

int sgb_r_set_seed(sgb_context *ctx, uint32_t seed) {
    return guarded(ctx, [&] { ctx->rng.set_seed(seed); });
}
int sgb_r_unif_rand(sgb_context *ctx, int64_t n, double *out) {
    return guarded(ctx, [&] {
        if (!out || n < 0) throw sgb::Error(SGB_ERR_INVALID, "out is NULL or n < 0");
        for (int64_t i = 0; i < n; i++) out[i] = ctx->rng.unif_rand();
    });
}
int sgb_r_sample_int(sgb_context *ctx, int32_t n, int32_t *out) {
    return guarded(ctx, [&] {
        if (!out || n < 0) throw sgb::Error(SGB_ERR_INVALID, "out is NULL or n < 0");
        ctx->rng.sample_int(n, out);
    });
}

int sgb_get_stats(sgb_context *ctx, sgb_stats *out) {
    return guarded(ctx, [&] {
        if (!out) throw sgb::Error(SGB_ERR_INVALID, "out is NULL");
        *out = ctx->stats;
    });
}
int sgb_reset_stats(sgb_context *ctx) {
    return guarded(ctx, [&] { ctx->stats = sgb_stats{}; });
}

int sgb_synth_geno_device(sgb_context *ctx, int64_t n_samp, int64_t n_variant_local, int64_t variant_offset, uint64_t seed,
                          double missing_rate, uint8_t **packed_device) {
    return guarded(ctx, [&] {
        if (!packed_device || n_samp < 1 || n_variant_local < 1) throw sgb::Error(SGB_ERR_INVALID, "invalid arguments");
        const size_t bytes = (size_t)((n_samp + 3) / 4) * n_variant_local;
        uint8_t *p = nullptr;
        SGB_CUDA(cudaMalloc((void **)&p, bytes));
        try {
            sgb::synth_geno(*ctx, n_samp, n_variant_local, variant_offset, seed, missing_rate, p);
        } catch (...) {
            cudaFree(p);
            throw;
        }
        *packed_device = p;
    });
}

int sgb_set_profiling(sgb_context *ctx, int on) {
    return guarded(ctx, [&] { ctx->profiling = on != 0; if (on) ctx->ktimes.clear(); });
}
int sgb_kernel_times(sgb_context *ctx, char *buf, int64_t buf_size) {
    return guarded(ctx, [&] {
        std::string s;
        for (auto &kv : ctx->ktimes) {
            char line[256];
            snprintf(line, sizeof(line), "%s %.6f %lld\n", kv.first.c_str(), kv.second.first, (long long)kv.second.second);
            s += line;
        }
        if ((int64_t)s.size() + 1 > buf_size) throw sgb::Error(SGB_ERR_INVALID, "buffer too small");
        memcpy(buf, s.c_str(), s.size() + 1);
    });
}

int sgb_malloc_device(sgb_context *ctx, int64_t bytes, void **ptr_device) {
    return guarded(ctx, [&] { SGB_CUDA(cudaMalloc(ptr_device, (size_t)bytes)); });
}
int sgb_free_device(sgb_context *ctx, void *ptr_device) {
    return guarded(ctx, [&] { SGB_CUDA(cudaFree(ptr_device)); });
}
int sgb_malloc_host(sgb_context *ctx, int64_t bytes, void **ptr_host) {
    return guarded(ctx, [&] { SGB_CUDA(cudaMallocHost(ptr_host, (size_t)bytes)); });
}
int sgb_free_host(sgb_context *ctx, void *ptr_host) {
    return guarded(ctx, [&] { SGB_CUDA(cudaFreeHost(ptr_host)); });
}
int sgb_copy_to_device(sgb_context *ctx, void *dst_device, const void *src_host, int64_t bytes) {
    return guarded(ctx, [&] { ctx->h2d(dst_device, src_host, (size_t)bytes); ctx->sync(); });
}
int sgb_copy_from_device(sgb_context *ctx, void *dst_host, const void *src_device, int64_t bytes) {
    return guarded(ctx, [&] { ctx->d2h(dst_host, src_device, (size_t)bytes); ctx->sync(); });
}

int sgb_time_products_device(sgb_context *ctx, const double *b_device, double *out_device, int k, int reps, float *total_ms) {
    return guarded(ctx, [&] {
        if (!b_device || !out_device || k < 1 || reps < 1 || !total_ms) throw sgb::Error(SGB_ERR_INVALID, "invalid arguments");
        ctx->sync();
        SGB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        for (int r = 0; r < reps; r++) sgb::grm_mv_device(*ctx, b_device, out_device, k);
        SGB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        ctx->sync();
        SGB_CUDA(cudaEventElapsedTime(total_ms, ctx->ev0, ctx->ev1));
    });
}

}  // extern "C"
