// See vecops.cuh.
#include <cfloat>

#include "vecops.cuh"

namespace sgb {

namespace {

constexpr int kThreads = 256;

inline int red_grid(const Context &c, int64_t n) {
    int64_t g = (n + kThreads * 4 - 1) / (kThreads * 4);
    int64_t cap = 2 * (int64_t)c.sm_count;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// Sum NS values over the block in a fixed tree; result valid in thread 0.
template <int NS>
__device__ __forceinline__ void block_sum(double (&v)[NS], double *smem /* [NS][8] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; s++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[s] += __shfl_xor_sync(0xffffffffu, v[s], o);
        if (lane == 0) smem[s * 8 + warp] = v[s];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            double t = 0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; w++) t += smem[s * 8 + w];
            v[s] = t;
        }
    }
    __syncthreads();
}

// Publish this block's partial sums; the last block of the group (blockIdx.y) adds all partials in
// block order and writes out[s].  partial layout: [group][NS][gridDim.x].
template <int NS>
__device__ __forceinline__ void finalize_sums(double (&v)[NS], double *partial, unsigned int *counter,
                                              double *const (&out)[NS]) {
    __shared__ bool is_last;
    const int G = gridDim.x, grp = blockIdx.y;
    double *pp = partial + (size_t)grp * NS * G;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) pp[s * G + blockIdx.x] = v[s];
        __threadfence();
        unsigned int t = atomicInc(&counter[grp], (unsigned int)(G - 1));
        is_last = (t == (unsigned int)(G - 1));
    }
    __syncthreads();
    if (is_last && threadIdx.x < NS) {
        __threadfence();
        const volatile double *q = pp + threadIdx.x * G;
        double t = 0;
        for (int i = 0; i < G; i++) t += q[i];
        *out[threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kThreads) dot_pairs_kernel(DotArgs args, int64_t N, double *partial, unsigned int *counter,
                                                             double *out) {
    __shared__ double smem[8];
    const int q = blockIdx.y;
    const double *a = args.a[q], *b = args.b[q];
    double v[1] = {0};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kThreads) v[0] += a[i] * b[i];
    block_sum<1>(v, smem);
    double *const o[1] = {out + q};
    finalize_sums<1>(v, partial, counter, o);
}

__global__ void lincomb_kernel(double *out, double s, const double *base, const double *cols, int64_t ld, LinArgs la, int64_t N) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        double t = 0;
        for (int k = 0; k < la.n; k++) t += la.coef[k] * cols[i + (int64_t)k * ld];
        out[i] = (base ? s * base[i] : 0.0) + t;
    }
}

__device__ __forceinline__ double linkinv1(int fam, double eta) {
    if (fam == SGB_FAMILY_GAUSSIAN) return eta;
    double t = (eta < -30) ? DBL_EPSILON : ((eta > 30) ? 1 / DBL_EPSILON : exp(eta));
    return t / (1 + t);
}
__device__ __forceinline__ double mu_eta1(int fam, double eta) {
    if (fam == SGB_FAMILY_GAUSSIAN) return 1;
    double e = exp(eta), op = 1 + e;
    return (eta > 30 || eta < -30) ? DBL_EPSILON : e / (op * op);
}
__device__ __forceinline__ double variance1(int fam, double mu) { return fam == SGB_FAMILY_GAUSSIAN ? 1 : mu * (1 - mu); }

__global__ void family_update_kernel(int fam, double *eta, const double *offset, const double *y, double *mu, double *Y,
                                     double *W, int add_offset, int64_t N) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double off = offset ? offset[i] : 0.0;
        double e = eta[i];
        if (add_offset) { e += off; eta[i] = e; }
        const double m = linkinv1(fam, e), me = mu_eta1(fam, e);
        mu[i] = m;
        Y[i] = e - off + (y[i] - m) / me;
        W[i] = (me * me) / variance1(fam, m);
    }
}

__global__ void family_weights_kernel(int fam, const double *eta, const double *mu, double *W, int64_t N) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double me = mu_eta1(fam, eta[i]);
        W[i] = (me * me) / variance1(fam, mu[i]);
    }
}

__global__ void eta_update_kernel(double *eta, const double *Y, const double *SiY, const double *SiX, int64_t ld, LinArgs la,
                                  double tau0, const double *w, int64_t N) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        double t = 0;
        for (int k = 0; k < la.n; k++) t += SiX[i + (int64_t)k * ld] * la.coef[k];
        eta[i] = Y[i] - tau0 * (SiY[i] - t) / w[i];
    }
}

__global__ void diag_sigma_kernel(const double *w, const double *diag, double tau0, double tau1, double *out, int invert, int64_t N) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        double v = tau0 / w[i] + tau1 * diag[i];
        if (v < 1e-4) v = 1e-4;
        out[i] = invert ? 1 / v : v;
    }
}

__global__ void rademacher_kernel(const int8_t *bits, double *out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = 2.0 * (double)bits[i] - 1;
}

__global__ void __launch_bounds__(kThreads) weighted_sumsq_kernel(const double *wgt, const double *G, int64_t ld, int64_t N,
                                                                  double *partial, unsigned int *counter, double *out) {
    __shared__ double smem[8];
    const double *g = G + (int64_t)blockIdx.y * ld;
    double v[1] = {0};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kThreads) {
        double x = g[i];
        v[0] += (wgt ? wgt[i] : 1.0) * x * x;
    }
    block_sum<1>(v, smem);
    double *const o[1] = {out + blockIdx.y};
    finalize_sums<1>(v, partial, counter, o);
}

__global__ void impute_flip_kernel(double *G0, double imp, int flip, int64_t N) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        double v = G0[i];
        if (!isfinite(v)) v = imp;
        if (flip) v = 2 - v;
        G0[i] = v;
    }
}

// ---------------- PCG kernels: grid (G, n_active); column = cols[blockIdx.y] ----------------
__global__ void __launch_bounds__(kThreads) pcg_init_kernel(const double *b, const double *minv, double *r, double *z, double *p,
                                                            double *x, int64_t N, int K, double *scal, double *partial,
                                                            unsigned int *counter) {
    __shared__ double smem[16];
    const int col = blockIdx.y;
    const int64_t o = (int64_t)col * N;
    double v[2] = {0, 0};  // rz, rr
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kThreads) {
        const double ri = b[o + i], zi = minv[i] * ri;
        r[o + i] = ri; z[o + i] = zi; p[o + i] = zi; x[o + i] = 0;
        v[0] += ri * zi; v[1] += ri * ri;
    }
    block_sum<2>(v, smem);
    double *const out[2] = {scal + 0 * K + col, scal + 3 * K + col};
    finalize_sums<2>(v, partial, counter, out);
}

// pack active columns of p into a contiguous buffer for the product
__global__ void pcg_pack_kernel(const double *p, const int *cols, double *packed, int64_t N) {
    const int col = cols[blockIdx.y];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        packed[(int64_t)blockIdx.y * N + i] = p[(int64_t)col * N + i];
}

// Ap = tau0 * (p * (1/w)) + tau1 * GRMp   (get_crossprod :564-576);  pAp = sum(p * Ap)
__global__ void __launch_bounds__(kThreads) pcg_ap_kernel(const double *p, const double *w, const double *gp, double tau0,
                                                          double tau1, double *Ap, const int *cols, int64_t N, int K,
                                                          double *scal, double *partial, unsigned int *counter) {
    __shared__ double smem[8];
    const int col = cols[blockIdx.y];
    const int64_t o = (int64_t)col * N, og = (int64_t)blockIdx.y * N;
    double v[1] = {0};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kThreads) {
        const double pi = p[o + i];
        double a = tau0 * (pi * (1 / w[i]));
        if (gp) a += tau1 * gp[og + i];
        Ap[o + i] = a;
        v[0] += pi * a;
    }
    block_sum<1>(v, smem);
    double *const out[1] = {scal + 2 * K + col};
    finalize_sums<1>(v, partial, counter, out);
}

// a = rz/pAp; x += a p; r1 = r - a Ap; z1 = minv r1; rz_new = sum(z1 r1); rr = sum(r1 r1)   (:599-607)
__global__ void __launch_bounds__(kThreads) pcg_update_kernel(double *x, double *r, double *z, const double *p, const double *Ap,
                                                              const double *minv, const int *cols, int64_t N, int K, int rz_cur,
                                                              double *scal, double *partial, unsigned int *counter) {
    __shared__ double smem[16];
    const int col = cols[blockIdx.y];
    const int64_t o = (int64_t)col * N;
    const double a = scal[rz_cur * K + col] / scal[2 * K + col];
    double v[2] = {0, 0};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kThreads) {
        x[o + i] += a * p[o + i];
        const double ri = r[o + i] - a * Ap[o + i];
        const double zi = minv[i] * ri;
        r[o + i] = ri; z[o + i] = zi;
        v[0] += zi * ri; v[1] += ri * ri;
    }
    block_sum<2>(v, smem);
    double *const out[2] = {scal + (1 - rz_cur) * K + col, scal + 3 * K + col};
    finalize_sums<2>(v, partial, counter, out);
}

// bet = rz_new / rz; p = z1 + bet p
__global__ void pcg_p_kernel(double *p, const double *z, const int *cols, int64_t N, int K, int rz_cur, const double *scal) {
    const int col = cols[blockIdx.y];
    const int64_t o = (int64_t)col * N;
    const double bet = scal[(1 - rz_cur) * K + col] / scal[rz_cur * K + col];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        p[o + i] = z[o + i] + bet * p[o + i];
}

inline int ew_grid(const Context &c, int64_t n) {
    int64_t g = (n + 255) / 256;
    int64_t cap = 8 * (int64_t)c.sm_count;
    return (int)std::max<int64_t>(1, std::min(g, cap));
}

void ensure_red(Context &c, size_t groups, int ns, int G) {
    c.red_partial.ensure(groups * ns * G);
    if (c.red_counter.n < groups) {
        c.red_counter.ensure(std::max<size_t>(groups, 256));
        SGB_CUDA(cudaMemsetAsync(c.red_counter.get(), 0, sizeof(unsigned int) * c.red_counter.n, c.stream));
    }
    c.red_out.ensure(std::max<size_t>(groups * ns, 256));
    c.h_scalars.ensure(std::max<size_t>(groups * ns, 256));
}

}  // namespace

void dot_pairs(Context &c, const std::vector<const double *> &a, const std::vector<const double *> &b, double *out_host) {
    const int G = red_grid(c, c.N);
    size_t done = 0;
    while (done < a.size()) {
        DotArgs args;
        args.q = (int)std::min<size_t>(kMaxPairs, a.size() - done);
        for (int i = 0; i < args.q; i++) { args.a[i] = a[done + i]; args.b[i] = b[done + i]; }
        ensure_red(c, args.q, 1, G);
        dot_pairs_kernel<<<dim3(G, args.q), kThreads, 0, c.stream>>>(args, c.N, c.red_partial.get(), c.red_counter.get(),
                                                                     c.red_out.get());
        SGB_CHECK_LAUNCH();
        c.stats.n_kernel_launches++;
        c.d2h(c.h_scalars.p, c.red_out.get(), sizeof(double) * args.q);
        c.sync();
        for (int i = 0; i < args.q; i++) out_host[done + i] = c.h_scalars.p[i];
        done += args.q;
    }
}

void lincomb(Context &c, double *out, double s, const double *base, const double *cols, int64_t ld,
             const std::vector<double> &coef) {
    if ((int)coef.size() > kMaxCoef) throw Error(SGB_ERR_INVALID, "too many fixed-effect columns (max 40)");
    LinArgs la;
    la.n = (int)coef.size();
    for (int i = 0; i < la.n; i++) la.coef[i] = coef[i];
    lincomb_kernel<<<ew_grid(c, c.N), 256, 0, c.stream>>>(out, s, base, cols, ld, la, c.N);
    SGB_CHECK_LAUNCH();
    c.stats.n_kernel_launches++;
}

void family_update(Context &c, int family, double *eta, const double *offset, const double *y, double *mu, double *Y, double *W,
                   bool add_offset) {
    family_update_kernel<<<ew_grid(c, c.N), 256, 0, c.stream>>>(family, eta, offset, y, mu, Y, W, add_offset ? 1 : 0, c.N);
    SGB_CHECK_LAUNCH();
    c.stats.n_kernel_launches++;
}

void family_weights(Context &c, int family, const double *eta, const double *mu, double *W) {
    family_weights_kernel<<<ew_grid(c, c.N), 256, 0, c.stream>>>(family, eta, mu, W, c.N);
    SGB_CHECK_LAUNCH();
    c.stats.n_kernel_launches++;
}

void eta_update(Context &c, double *eta, const double *Y, const double *Sigma_iY, const double *Sigma_iX, int64_t ld,
                const std::vector<double> &alpha, double tau0, const double *w) {
    if ((int)alpha.size() > kMaxCoef) throw Error(SGB_ERR_INVALID, "too many fixed-effect columns (max 40)");
    LinArgs la;
    la.n = (int)alpha.size();
    for (int i = 0; i < la.n; i++) la.coef[i] = alpha[i];
    eta_update_kernel<<<ew_grid(c, c.N), 256, 0, c.stream>>>(eta, Y, Sigma_iY, Sigma_iX, ld, la, tau0, w, c.N);
    SGB_CHECK_LAUNCH();
    c.stats.n_kernel_launches++;
}

void diag_sigma(Context &c, const double *w, double tau0, double tau1, double *out, bool invert) {
    diag_sigma_kernel<<<ew_grid(c, c.N), 256, 0, c.stream>>>(w, c.diag.get(), tau0, tau1, out, invert ? 1 : 0, c.N);
    SGB_CHECK_LAUNCH();
    c.stats.n_kernel_launches++;
}

void expand_rademacher(Context &c, const int8_t *bits_device, double *out, int64_t count) {
    rademacher_kernel<<<ew_grid(c, count), 256, 0, c.stream>>>(bits_device, out, count);
    SGB_CHECK_LAUNCH();
    c.stats.n_kernel_launches++;
}

void weighted_sumsq(Context &c, const double *wgt, const double *G, int64_t ld, int k, double *out_host) {
    const int Gd = red_grid(c, c.N);
    ensure_red(c, k, 1, Gd);
    weighted_sumsq_kernel<<<dim3(Gd, k), kThreads, 0, c.stream>>>(wgt, G, ld, c.N, c.red_partial.get(), c.red_counter.get(),
                                                                  c.red_out.get());
    SGB_CHECK_LAUNCH();
    c.stats.n_kernel_launches++;
    c.d2h(c.h_scalars.p, c.red_out.get(), sizeof(double) * k);
    c.sync();
    for (int i = 0; i < k; i++) out_host[i] = c.h_scalars.p[i];
}

void impute_flip(Context &c, double *G0, double impute_value, bool flip) {
    impute_flip_kernel<<<ew_grid(c, c.N), 256, 0, c.stream>>>(G0, impute_value, flip ? 1 : 0, c.N);
    SGB_CHECK_LAUNCH();
    c.stats.n_kernel_launches++;
}

void pcg_solve(Context &c, PcgWork &ws, const double *w, double tau0, double tau1, const double *b, int K, int maxiter,
               double tol, double *x, int *iters_host) {
    const int64_t N = c.N;
    const int G = red_grid(c, N);
    const size_t NK = (size_t)N * K;
    ws.minv.ensure(N); ws.r.ensure(NK); ws.z.ensure(NK); ws.p.ensure(NK); ws.Ap.ensure(NK); ws.gp.ensure(NK);
    ws.scal.ensure((size_t)4 * K); ws.cols.ensure(K); ws.partial.ensure((size_t)K * 2 * G);
    if (ws.counter.n < (size_t)K) {
        ws.counter.ensure(std::max(K, 64));
        SGB_CUDA(cudaMemsetAsync(ws.counter.get(), 0, sizeof(unsigned int) * ws.counter.n, c.stream));
    }
    c.h_scalars.ensure(std::max<size_t>(4 * (size_t)K, 256));
    diag_sigma(c, w, tau0, tau1, ws.minv.get(), true);
    pcg_init_kernel<<<dim3(G, K), kThreads, 0, c.stream>>>(b, ws.minv.get(), ws.r.get(), ws.z.get(), ws.p.get(), x, N, K,
                                                           ws.scal.get(), ws.partial.get(), ws.counter.get());
    SGB_CHECK_LAUNCH();
    c.stats.n_kernel_launches++;
    std::vector<int> active(K), iters(K, 0);
    std::vector<double> rr(K);
    c.d2h(c.h_scalars.p, ws.scal.get() + 3 * (size_t)K, sizeof(double) * K);
    c.sync();
    for (int k = 0; k < K; k++) rr[k] = c.h_scalars.p[k];
    int rz_cur = 0;
    bool warned = false;
    for (;;) {
        int na = 0;
        for (int k = 0; k < K; k++)
            if (iters[k] < maxiter && rr[k] > tol) active[na++] = k;   // :595
        if (na == 0) break;
        c.h2d(ws.cols.get(), active.data(), sizeof(int) * na);
        const double *gp = nullptr;
        if (tau1 != 0) {   // :568 skips the GRM product when tau[1] == 0
            const double *pin = ws.p.get();
            if (na != K) {
                // scratch: Ap of inactive columns is dead, but keep it simple and use a dedicated pack buffer
                c.ws_vec.ensure((size_t)N * K);
                pcg_pack_kernel<<<dim3(ew_grid(c, N), na), 256, 0, c.stream>>>(ws.p.get(), ws.cols.get(), c.ws_vec.get(), N);
                SGB_CHECK_LAUNCH();
                c.stats.n_kernel_launches++;
                pin = c.ws_vec.get();
            }
            grm_mv_device(c, pin, ws.gp.get(), na);
            gp = ws.gp.get();
        }
        pcg_ap_kernel<<<dim3(G, na), kThreads, 0, c.stream>>>(ws.p.get(), w, gp, tau0, tau1, ws.Ap.get(), ws.cols.get(), N, K,
                                                              ws.scal.get(), ws.partial.get(), ws.counter.get());
        SGB_CHECK_LAUNCH();
        pcg_update_kernel<<<dim3(G, na), kThreads, 0, c.stream>>>(x, ws.r.get(), ws.z.get(), ws.p.get(), ws.Ap.get(),
                                                                  ws.minv.get(), ws.cols.get(), N, K, rz_cur, ws.scal.get(),
                                                                  ws.partial.get(), ws.counter.get());
        SGB_CHECK_LAUNCH();
        pcg_p_kernel<<<dim3(ew_grid(c, N), na), 256, 0, c.stream>>>(ws.p.get(), ws.z.get(), ws.cols.get(), N, K, rz_cur,
                                                                    ws.scal.get());
        SGB_CHECK_LAUNCH();
        c.stats.n_kernel_launches += 3;
        c.d2h(c.h_scalars.p, ws.scal.get() + 3 * (size_t)K, sizeof(double) * K);
        c.sync();
        for (int i = 0; i < na; i++) {
            const int k = active[i];
            rr[k] = c.h_scalars.p[k];
            iters[k]++;
            c.stats.n_pcg_iterations++;
        }
        // rz slot alternates, but only for the columns that took this step: inactive columns never read it again
        rz_cur = 1 - rz_cur;
    }
    for (int k = 0; k < K; k++) {
        if (iters[k] >= maxiter && !warned) {
            c.printf("PCG does not converge (may need to increase 'maxiter').\n");   // :610-611
            warned = true;
        }
        if (iters_host) iters_host[k] = iters[k];
    }
    c.stats.n_pcg_solves += K;
}

}  // namespace sgb
