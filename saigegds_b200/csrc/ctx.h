// Internal context and helpers of libsaigegds_b200 (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <chrono>
#include <cstdarg>
#include <map>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/saigegds_b200.h"
#include "rrng.h"

namespace sgb {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define SGB_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            throw sgb::Error(SGB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +    \
                                               __FILE__ + ":" + std::to_string(__LINE__) + ")");          \
    } while (0)

#define SGB_CHECK_LAUNCH() SGB_CUDA(cudaGetLastError())

// internal: a kernel with cross-CTA waits timed out; the C-ABI layer redoes the call on kernels without such waits (api.cu)
constexpr int SGB_INTERNAL_RETRY_NO_WAIT_KERNELS = -1000;

// host time spent in cudaMalloc / cudaFree by DevBuf (a debugging aid: SGB_FIT_TIMING prints it per fit)
inline double g_alloc_seconds = 0;
inline long g_alloc_calls = 0;

// Owning device buffer (cudaMalloc / cudaFree), resizable without preserving contents.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) {
            const auto t0 = std::chrono::steady_clock::now();
            cudaFree(p);
            g_alloc_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            g_alloc_calls++;
        }
        p = nullptr;
        n = 0;
    }
    void ensure(size_t count) {
        if (count <= n) return;
        release();
        const auto t0 = std::chrono::steady_clock::now();
        SGB_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
        g_alloc_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        g_alloc_calls++;
        n = count;
    }
    T *get() const { return p; }
};

// Pinned host staging buffer
template <typename T>
struct PinBuf {
    T *p = nullptr;
    size_t n = 0;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    void ensure(size_t count) {
        if (count <= n) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        SGB_CUDA(cudaMallocHost((void **)&p, count * sizeof(T)));
        n = count;
    }
};

struct Comm;  // comm.cu (NCCL through dlopen)

// Tiled int8-tensor-core layout of the genotype shard (grm_imma.cu)
struct ImmaPlan;
// Model and workspaces of the single-variant score test (score.cu)
struct ScoreState;
// Device workspaces of the null-model fits, kept between calls (solver.cu)
struct SolverWs;

struct Context {
    int dev = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // ---- genotype shard (state of saige_fitnull.cpp:122-131) ----
    int64_t N = 0;        // Geno_NumSamp
    int64_t NB = 0;       // Geno_PackedNumSamp = ceil(N/4)
    int64_t M = 0;        // local variants
    int64_t M_total = 0;  // Geno_NumVariant across all ranks
    int64_t var_offset = 0;
    size_t pitch = 0;     // bytes per variant row on the device (NB rounded up to 256)
    DevBuf<uint8_t> packed;   // [M][pitch], pad samples and pitch padding = code 3
    DevBuf<double> lut;       // buf_std_geno [M][4]
    DevBuf<double> diag;      // buf_diag_grm [N]
    DevBuf<int32_t> n_valid, sum;   // reference allele counts (all NB bytes, incl. pad codes)
    DevBuf<int32_t> cnt_num, cnt_sum;  // counts over samples < N only (get_geno_ds / f64_af_ac_impute semantics)
    std::vector<int32_t> h_cnt_num, h_cnt_sum;
    bool stored = false;

    // ---- product workspaces ----
    int kernel = SGB_KERNEL_AUTO;
    DevBuf<double> ws_partial;  // SIMT: [n_chunk][M] partial dots
    DevBuf<double> ws_tab;      // [M][4] per-variant apply table
    DevBuf<double> ws_vec;      // scratch N x k
    ImmaPlan *imma = nullptr;
    ScoreState *score = nullptr;
    SolverWs *solver_ws = nullptr;

    // ---- generic reduction workspace ----
    DevBuf<double> red_partial;
    DevBuf<double> red_out;
    DevBuf<unsigned int> red_counter;
    PinBuf<double> h_scalars;
    PinBuf<double> h_stage;    // pinned staging for host<->device vector copies

    // ---- multi-GPU ----
    Comm *comm = nullptr;
    int rank = 0, world = 1;

    // ---- host callbacks ----
    void (*print_fn)(const char *) = nullptr;
    void (*rademacher_fn)(void *, int, int, int64_t, int8_t *) = nullptr;
    void *cb_user = nullptr;
    RRng rng;

    sgb_stats stats{};
    DevBuf<double> io_in, io_out;   // persistent device buffers of the host-pointer entry points

    // per-kernel CUDA-event timing (sgb_set_profiling): serialises the stream, use outside timed regions
    bool profiling = false;
    std::map<std::string, std::pair<double, int64_t>> ktimes;
    cudaEvent_t pev0 = nullptr, pev1 = nullptr;
    void prof_begin() { if (profiling) cudaEventRecord(pev0, stream); }
    void prof_end(const char *name) {
        if (!profiling) return;
        cudaEventRecord(pev1, stream);
        cudaEventSynchronize(pev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, pev0, pev1);
        auto &e = ktimes[name];
        e.first += ms;
        e.second += 1;
    }

    void printf(const char *fmt, ...) {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        if (print_fn) print_fn(buf);
        else { fputs(buf, stdout); fflush(stdout); }
    }
    void require_stored() const {
        if (!stored) throw Error(SGB_ERR_STATE, "no genotypes stored: call sgb_store_2b_geno first");
    }
    bool fused_disabled = false;         // set after a time-out of the fused kernel: the two-pass kernels take over
    volatile int *async_err = nullptr;   // pinned flag copied back after every fused-kernel launch
    int *async_err_dev = nullptr;        // its device-side source (cleared after an error was reported)
    bool product_reduced = false;        // the product kernels already summed `out` over the ranks (combine + all-reduce in one kernel)
    volatile int *comm_err = nullptr;    // pinned flag of the peer-memory all-reduce kernel (comm.cu): a rank did not arrive in time
    void sync() {
        const auto t_sync0 = std::chrono::steady_clock::now();
        SGB_CUDA(cudaStreamSynchronize(stream));
        stats.n_host_syncs++;
        stats.host_wait_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_sync0).count();
        if (comm_err && *comm_err) {
            *comm_err = 0;
            throw Error(SGB_ERR_COMM, "the peer-memory all-reduce timed out waiting for another rank (SGB_WAIT_TIMEOUT_MS); "
                                      "SGB_NO_PEER_ALLREDUCE=1 selects ncclAllReduce for every collective");
        }
        if (async_err && *async_err) {
            const int code = *async_err;
            *async_err = 0;
            if (async_err_dev) cudaMemsetAsync(async_err_dev, 0, sizeof(int), stream);
            if (code == 2) throw Error(SGB_ERR_CUDA, "the fused GRM kernel found |e_j| above its a-priori bound (internal error)");
            throw Error(SGB_INTERNAL_RETRY_NO_WAIT_KERNELS, "a GRM kernel timed out waiting for another CTA or warp (wall-clock bound, "
                                                             "SGB_WAIT_TIMEOUT_MS)");
        }
    }
    void h2d(void *dst, const void *src, size_t bytes) {
        SGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    }
    void d2h(void *dst, const void *src, size_t bytes) {
        SGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    }
};

// ---- store.cu ----
void store_device_layout(Context &c, const uint8_t *src_device, size_t src_pitch);  // fills packed/lut/diag/counts
void decode_variant(Context &c, int64_t local_idx, double *out_device);             // get_geno_ds, missing -> NaN
void synth_geno(Context &c, int64_t n_samp, int64_t m_local, int64_t var_offset, uint64_t seed, double miss,
                uint8_t *out_device);
// GDS genotype/data block (bit2 allele pairs) -> 2-bit dosage rows [m_file][ceil(N/4)] + per-variant allele counts
void gds_to_dosage(Context &c, const uint8_t *bits_host, int64_t n_file, int64_t m_file, const int32_t *sample_sel_host,
                   int64_t N, uint8_t *packed_device, int32_t *n_valid_alleles_device, int32_t *n_alt_alleles_device);
void gather_rows(Context &c, const uint8_t *src_device, const int64_t *rows_host, int64_t n_rows, int64_t NB,
                 uint8_t *dst_device);
// ---- grm_simt.cu ----
void simt_table_apply(Context &c, const double *tab_device, double *out_device, double scale);
void simt_grm_mv(Context &c, const double *b_device, double *out_device);  // local shard, scaled by 1/M_total
// ---- grm_imma.cu ----
void imma_prepare(Context &c);
void imma_release(Context &c);
bool imma_available(const Context &c);
void imma_grm_mv(Context &c, const double *b_device, double *out_device, int k);
// class sums of a packed block against fixed model columns on the tcgen05 pair kernel (the dense part of the score scan)
constexpr int kClassMaxCols = 32, kClassDigitRows = 192, kClassScal = 8;   // columns per group, rows of a digit matrix, scalars per column
void umma_class_digits(Context &c, const double *cols_device, int64_t n, int ncols, int64_t cpad, int8_t *digits_device,
                       double *scal_device, long long *tot_device);
void umma_class_sums(Context &c, const uint8_t *packed_device, size_t pitch, int64_t rows, int64_t n, const int8_t *digits_device,
                     int64_t cpad, int ncols, int amode, unsigned long long *out_lo, unsigned long long *out_hi, int *err_device);
void solver_release(Context &c);   // solver.cu: frees the kept workspaces of the fits
// ---- product dispatch (solver.cu) ----
void grm_mv_device(Context &c, const double *b_device, double *out_device, int k);
// ---- score.cu: single-variant score test + SPA (saige_main.cpp:101-407, SPATest.cpp) ----
void score_init(Context &c, const sgb_score_model *m, double maf, double mac, double missing, double spa_pval);
void score_release(Context &c);
void score_set_path(Context &c, int path);
void saddle_prob_dense(Context &c, const double *g_device, const double *mu_device, int64_t n, double q, double m1, double var1,
                       double cutoff, double *pval, double *p_noadj, bool *converged);   // Saddle_Prob, SPATest.cpp:232-296
double qnorm_host(double p);   // R's qnorm5 (AS 241)
// ---- solver.cu: interaction-term test, saige_GxG_snp_bin (saige_fitnull.cpp:1480-1558) ----
void gxg_snp_bin(Context &c, const sgb_fit0 *f, const double tau[2], const double *inter_term, const sgb_noK *noK,
                 const sgb_param *P, int verbose, sgb_gxg *out);
void score_test_packed(Context &c, const uint8_t *packed_host, int64_t nb, int64_t n_var, double *out, int32_t *valid);
void score_test_dosage(Context &c, const double *dosage_host, int64_t n_var, double *out, int32_t *valid);
void score_test_stored(Context &c, int64_t first, int64_t n_var, double *out, int32_t *valid, float *kernel_ms);
// ---- comm.cu ----
void comm_unique_id(unsigned char id[128]);
void comm_init(Context &c, const unsigned char id[128], int rank, int world);
void comm_destroy(Context &c);
void comm_allreduce_sum(Context &c, double *buf_device, size_t count);
// parts of a fused single-RHS product before its last addition: out_n = rout_n + sum(h_part) - sum_t cpart[t][n]
struct CombineSrc {
    const double *rout = nullptr, *cpart = nullptr, *h_part = nullptr;
    int n_ctiles = 0, n_hpart = 0;
};
bool comm_combine_allreduce(Context &c, const CombineSrc &src, double *out_device);

}  // namespace sgb
