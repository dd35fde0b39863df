// Device-resident null-model drivers: IRLS (get_coeff), Hutchinson trace, AI score, tau update,
// AI-REML outer loops and variance-ratio estimation.  Host code orchestrates kernels; all N-vectors
// stay in HBM, only p x p algebra (p <= 40) and scalars live on the host.
//
// Differences from the reference that do not change results beyond FP64 rounding:
//   * the (1+p) solves of get_coeff_w (:744-752), the nrun solves of get_trace (:646-654) and the
//     marker solves of the variance-ratio step (:1321) are mutually independent and run as
//     multi-right-hand-side PCG batches; every column keeps its own alpha/beta/stop test;
//   * the Rademacher vectors u_i are identical in every get_trace call because the reference re-seeds
//     R's RNG each time (:631, :676), so u_i and GRM*u_i (:652) are generated once per fit and cached.
#include <chrono>
#include <future>
#include <map>
#include <algorithm>
#include <cmath>

#include "vecops.cuh"

namespace sgb {

void grm_mv_device(Context &c, const double *b, double *out, int k) {
    c.require_stored();
    c.product_reduced = false;
    const bool use_imma = (c.kernel == SGB_KERNEL_IMMA) || (c.kernel == SGB_KERNEL_IMMA_TWOPASS) || (c.kernel == SGB_KERNEL_UMMA) ||
                          (c.kernel == SGB_KERNEL_AUTO && imma_available(c));
    if (use_imma) {
        imma_grm_mv(c, b, out, k);
    } else {
        for (int i = 0; i < k; i++) simt_grm_mv(c, b + (size_t)i * c.N, out + (size_t)i * c.N);
        c.stats.n_product_launches += k;
    }
    if (c.product_reduced) {
        c.product_reduced = false;
    } else if (c.world > 1) {
        c.prof_begin();
        comm_allreduce_sum(c, out, (size_t)c.N * k);
        c.prof_end(k == 1 ? "ncclAllReduce (N doubles)" : "ncclAllReduce (N x k doubles)");
    }
    if (c.profiling) {
        char nm[48];
        snprintf(nm, sizeof(nm), "[calls] grm_mv k=%02d", k);
        c.ktimes[nm].second += 1;
    }
    c.stats.n_products += k;
}

namespace {

typedef std::vector<double> hvec;

// wall-clock phase timer of a fit (env SGB_FIT_TIMING: printed once at the end; a debugging aid for host-side overheads)
struct PhaseTimer {
    std::map<std::string, std::pair<double, double>> t;   // wall seconds, of which waiting for the GPU
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    const sgb_stats *st = nullptr;
    double w0 = 0;
    void lap(const char *name) {
        const auto now = std::chrono::steady_clock::now();
        auto &e = t[name];
        e.first += std::chrono::duration<double>(now - t0).count();
        if (st) { e.second += st->host_wait_s - w0; w0 = st->host_wait_s; }
        t0 = now;
    }
};
PhaseTimer *g_phase = nullptr;
inline void lap(const char *name) { if (g_phase) g_phase->lap(name); }

struct hmat {  // column-major host matrix
    int nr = 0, nc = 0;
    hvec a;
    hmat() {}
    hmat(int r, int cc) : nr(r), nc(cc), a((size_t)r * cc, 0.0) {}
    double &operator()(int i, int j) { return a[i + (size_t)j * nr]; }
    double operator()(int i, int j) const { return a[i + (size_t)j * nr]; }
};

hvec matvec(const hmat &A, const hvec &v) {
    hvec o(A.nr, 0.0);
    for (int j = 0; j < A.nc; j++) for (int i = 0; i < A.nr; i++) o[i] += A(i, j) * v[j];
    return o;
}

hmat general_inverse(hmat a) {
    const int n = a.nr;
    hmat inv(n, n);
    for (int i = 0; i < n; i++) inv(i, i) = 1;
    for (int col = 0; col < n; col++) {
        int piv = col;
        for (int r = col + 1; r < n; r++) if (fabs(a(r, col)) > fabs(a(piv, col))) piv = r;
        if (a(piv, col) == 0) throw Error(SGB_ERR_INVALID, "inv(): matrix is singular");
        if (piv != col) for (int j = 0; j < n; j++) { std::swap(a(col, j), a(piv, j)); std::swap(inv(col, j), inv(piv, j)); }
        const double d = 1 / a(col, col);
        for (int j = 0; j < n; j++) { a(col, j) *= d; inv(col, j) *= d; }
        for (int r = 0; r < n; r++) {
            if (r == col) continue;
            const double f = a(r, col);
            if (f != 0) for (int j = 0; j < n; j++) { a(r, j) -= f * a(col, j); inv(r, j) -= f * inv(col, j); }
        }
    }
    return inv;
}

// mat_inv (:722-733): inverse of symmatu(m) through Cholesky, general inverse as the fallback
hmat mat_inv(Context &c, const hmat &m) {
    const int n = m.nr;
    hmat s = m;
    for (int j = 0; j < n; j++) for (int i = j + 1; i < n; i++) s(i, j) = m(j, i);
    hmat L(n, n);
    bool pd = true;
    for (int j = 0; j < n && pd; j++) {
        double d = s(j, j);
        for (int k = 0; k < j; k++) d -= L(j, k) * L(j, k);
        if (!(d > 0)) { pd = false; break; }
        L(j, j) = sqrt(d);
        for (int i = j + 1; i < n; i++) {
            double t = s(i, j);
            for (int k = 0; k < j; k++) t -= L(i, k) * L(j, k);
            L(i, j) = t / L(j, j);
        }
    }
    if (!pd) {
        c.printf("Warning: arma::inv_sympd(), matrix is singular or not positive definite, use arma::inv() instead.\n");
        return general_inverse(s);
    }
    hmat Li(n, n);
    for (int j = 0; j < n; j++) {
        Li(j, j) = 1 / L(j, j);
        for (int i = j + 1; i < n; i++) {
            double t = 0;
            for (int k = j; k < i; k++) t -= L(i, k) * Li(k, j);
            Li(i, j) = t / L(i, i);
        }
    }
    hmat r(n, n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double t = 0;
            for (int k = i; k < n; k++) t += Li(k, i) * Li(k, j);
            r(i, j) = r(j, i) = t;
        }
    return r;
}

double calcCV(const hvec &x) {  // :618-623
    const size_t n = x.size();
    double m = 0;
    for (double v : x) m += v;
    m /= n;
    double ss = 0;
    for (double v : x) ss += (v - m) * (v - m);
    return sqrt(ss / (n - 1)) / (m * (int)n);
}

void print_vec(Context &c, const char *indent, const char *s, const double *x, int n, bool nl = true) {
    std::string line = std::string(indent) + s + "(";
    char buf[64];
    for (int i = 0; i < n; i++) {
        if (i) line += ", ";
        snprintf(buf, sizeof(buf), "%0.7g", x[i]);
        line += buf;
    }
    line += nl ? ")\n" : ")";
    c.printf("%s", line.c_str());
}

}  // namespace

// Device workspaces of the fits, owned by the context and kept between calls (they only grow): a fit makes ~50 cudaMalloc /
// cudaFree calls otherwise, 10-70 ms -- nothing next to 1.2 s on one GPU, a fifth of the 0.3 s fit on eight.
struct SolverWs {
    DevBuf<double> X, y, offset, W, Y, mu, eta, rhs, sol;
    DevBuf<double> eta_acc, mu_acc;
    DevBuf<double> PY, APY, PAPY1, PAPY, PA0PY1, PA0PY;
    DevBuf<double> U, AU, SiU, PU;
    DevBuf<double> eta0;
    DevBuf<double> vr[10];   // variance-ratio step / interaction test: eta, mu, XVt, XXVX_inv, SiX1, G0, G, SiG, adj, wgt
    DevBuf<int8_t> bits;
    PcgWork pcg;
};
void solver_release(Context &c) {
    delete c.solver_ws;
    c.solver_ws = nullptr;
}

namespace {

struct Solver {
    Context &c;
    const int64_t N;
    const int p;
    const int family;
    const sgb_param P;
    SolverWs &ws;
    DevBuf<double> &X = ws.X, &y = ws.y, &offset = ws.offset, &W = ws.W, &Y = ws.Y, &mu = ws.mu, &eta = ws.eta, &rhs = ws.rhs,
                   &sol = ws.sol;   // rhs/sol: N x (1+p)  [Y | X] -> [Sigma_iY | Sigma_iX]
    DevBuf<double> &eta_acc = ws.eta_acc, &mu_acc = ws.mu_acc;
    DevBuf<double> &PY = ws.PY, &APY = ws.APY, &PAPY1 = ws.PAPY1, &PAPY = ws.PAPY, &PA0PY1 = ws.PA0PY1, &PA0PY = ws.PA0PY;
    DevBuf<double> &U = ws.U, &AU = ws.AU, &SiU = ws.SiU, &PU = ws.PU;
    DevBuf<int8_t> &bits = ws.bits;
    int n_u = 0, cap_u = 0;
    RRng trace_rng;
    std::future<std::vector<int8_t>> pre_draws;   // the first nrun Rademacher vectors, drawn beside get_coeff (prefetch_draws)
    bool has_offset = false;
    PcgWork &pcg = ws.pcg;
    hmat cov;
    hvec alpha;

    static SolverWs &workspaces(Context &ctx) {
        if (!ctx.solver_ws) ctx.solver_ws = new SolverWs();
        return *ctx.solver_ws;
    }
    Solver(Context &ctx, int64_t n, int pp, int fam, const sgb_param &par)
        : c(ctx), N(n), p(pp), family(fam), P(par), ws(workspaces(ctx)) {
        cap_u = (int)std::min<size_t>(U.n, AU.n) / (int)std::max<int64_t>(1, n);   // columns the kept trace buffers already hold
    }

    double *SiY() { return sol.get(); }
    double *SiX() { return sol.get() + N; }
    const double *off() const { return has_offset ? offset.get() : nullptr; }

    void upload(DevBuf<double> &d, const double *h, size_t n) {
        d.ensure(n);
        c.h2d(d.get(), h, sizeof(double) * n);
    }
    void download(double *h, const double *d, size_t n) { c.d2h(h, d, sizeof(double) * n); }
    void copy(double *dst, const double *src, size_t n) {
        SGB_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToDevice, c.stream));
    }

    void init(const double *hX, const double *hy, const double *hoff) {
        upload(X, hX, (size_t)N * p);
        lap("set-up: upload X");
        upload(y, hy, N);
        has_offset = hoff != nullptr;
        if (has_offset) upload(offset, hoff, N);
        for (DevBuf<double> *b : {&W, &Y, &mu, &eta, &eta_acc, &mu_acc, &PY, &APY, &PAPY1, &PAPY}) b->ensure(N);
        rhs.ensure((size_t)N * (1 + p));
        sol.ensure((size_t)N * (1 + p));
        copy(rhs.get() + N, X.get(), (size_t)N * p);
        trace_rng.set_seed((uint32_t)P.seed);
    }

    // R's RNG is serial (3.8 ns per draw, 50 ms for 30 vectors of 430K): take the first nrun vectors on a host thread while the
    // GPU runs the first get_coeff.  Same stream of draws as drawing them in ensure_u (trace_rng is not touched until the join).
    void prefetch_draws() {
        if (c.rademacher_fn || n_u != 0 || pre_draws.valid()) return;
        const size_t cnt = (size_t)N * P.nrun;
        pre_draws = std::async(std::launch::async, [this, cnt] {
            std::vector<int8_t> h(cnt);
            for (size_t i = 0; i < cnt; i++) h[i] = (int8_t)trace_rng.bernoulli_half();
            return h;
        });
    }

    // t(A) %*% v for the p columns of A (ld N)
    hvec cols_dot(const double *A, int ncol, const double *v) {
        std::vector<const double *> a(ncol), b(ncol);
        for (int i = 0; i < ncol; i++) { a[i] = A + (size_t)i * N; b[i] = v; }
        hvec o(ncol);
        dot_pairs(c, a, b, o.data());
        return o;
    }
    double dot1(const double *a, const double *b) {
        double o;
        dot_pairs(c, {a}, {b}, &o);
        return o;
    }
    // out = v - Sigma_iX (cov (Sigma_iX' rhs))     (:651, :823, :831)
    void project(double *out, const double *v, const double *rhs_vec) {
        hvec t = matvec(cov, cols_dot(SiX(), p, rhs_vec));
        for (double &x : t) x = -x;
        lincomb(c, out, 1.0, v, SiX(), N, t);
    }

    // get_coeff_w, :739-758
    void get_coeff_w(const double tau[2]) {
        copy(rhs.get(), Y.get(), N);
        lap("other");
        pcg_solve(c, pcg, W.get(), tau[0], tau[1], rhs.get(), 1 + p, P.maxiterPCG, P.tolPCG, sol.get(), nullptr);
        lap("get_coeff_w: PCG of the 1+p columns");
        std::vector<const double *> a, b;
        for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) { a.push_back(X.get() + (size_t)i * N); b.push_back(SiX() + (size_t)j * N); }
        for (int i = 0; i < p; i++) { a.push_back(SiX() + (size_t)i * N); b.push_back(Y.get()); }
        hvec d(a.size());
        dot_pairs(c, a, b, d.data());
        hmat XtSiX(p, p);
        for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) XtSiX(i, j) = d[(size_t)j * p + i];
        cov = mat_inv(c, XtSiX);
        hvec t(d.begin() + (size_t)p * p, d.end());
        alpha = matvec(cov, t);
        eta_update(c, eta.get(), Y.get(), SiY(), SiX(), N, alpha, tau[0], W.get());
    }

    // get_coeff, :778-813.  eta0 is a device vector; results land in Y, mu, eta, W, cov, alpha, sol.
    void get_coeff(const double tau[2], const hvec &alpha0, const double *eta0) {
        const double tol_coef = 0.1;
        copy(eta.get(), eta0, N);
        family_update(c, family, eta.get(), off(), y.get(), mu.get(), Y.get(), W.get(), false);
        hvec a0 = alpha0;
        for (int it = 0; it < P.maxiter; it++) {
            get_coeff_w(tau);
            family_update(c, family, eta.get(), off(), y.get(), mu.get(), Y.get(), W.get(), true);
            double mx = 0;
            for (int k = 0; k < p; k++) mx = std::max(mx, fabs(alpha[k] - a0[k]) / (fabs(alpha[k]) + fabs(a0[k]) + tol_coef));
            if (mx < tol_coef) break;
            a0 = alpha;
        }
    }

    // Make sure Rademacher vectors [0, upto) and their GRM products are cached.
    void ensure_u(int upto) {
        if (upto <= n_u) return;
        if (upto > cap_u) {
            int new_cap = std::max(upto, cap_u + 10);
            DevBuf<double> nU, nAU;
            nU.ensure((size_t)N * new_cap); nAU.ensure((size_t)N * new_cap);
            if (n_u > 0) {
                SGB_CUDA(cudaMemcpyAsync(nU.get(), U.get(), sizeof(double) * (size_t)N * n_u, cudaMemcpyDeviceToDevice, c.stream));
                SGB_CUDA(cudaMemcpyAsync(nAU.get(), AU.get(), sizeof(double) * (size_t)N * n_u, cudaMemcpyDeviceToDevice, c.stream));
            }
            c.sync();
            std::swap(U.p, nU.p); std::swap(U.n, nU.n);
            std::swap(AU.p, nAU.p); std::swap(AU.n, nAU.n);
            cap_u = new_cap;
        }
        const int cnt = upto - n_u;
        lap("other");
        std::vector<int8_t> h;
        if (pre_draws.valid()) {
            h = pre_draws.get();
            if (n_u != 0 || h.size() != (size_t)N * cnt) {   // not the request the prefetch was made for: redo it in order
                trace_rng.set_seed((uint32_t)P.seed);
                h.clear();
            }
        }
        if (h.empty()) {
            h.resize((size_t)N * cnt);
            if (c.rademacher_fn) c.rademacher_fn(c.cb_user, n_u == 0 ? 1 : 0, P.seed, (int64_t)N * cnt, h.data());
            else for (size_t i = 0; i < h.size(); i++) h[i] = (int8_t)trace_rng.bernoulli_half();
        }
        lap("rademacher draws on the host (R's RNG)");
        bits.ensure(h.size());
        c.h2d(bits.get(), h.data(), h.size());
        expand_rademacher(c, bits.get(), U.get() + (size_t)N * n_u, (int64_t)N * cnt);
        c.sync();  // h goes out of scope
        grm_mv_device(c, U.get() + (size_t)N * n_u, AU.get() + (size_t)N * n_u, cnt);   // :652 / :698
        c.sync();
        lap("GRM u_i (cached once per fit)");
        n_u = upto;
    }

    // get_trace (:627-668) and get_trace_q (:672-718)
    void get_trace(const double tau[2], bool quant, double &trace0, double &trace1) {
        int nrunStart = 0, nrunEnd = P.nrun;
        double traceCV = P.traceCVcutoff + 0.1, traceCV0 = P.traceCVcutoff + 0.1;
        hvec buf(P.nrun, 0.0), buf0(P.nrun, 0.0);
        while (traceCV > P.traceCVcutoff || (quant && traceCV0 > P.traceCVcutoff)) {
            const int K = nrunEnd - nrunStart;
            ensure_u(nrunEnd);
            SiU.ensure((size_t)N * K); PU.ensure((size_t)N * K);
            const double *u = U.get() + (size_t)N * nrunStart, *au = AU.get() + (size_t)N * nrunStart;
            lap("other");
            pcg_solve(c, pcg, W.get(), tau[0], tau[1], u, K, P.maxiterPCG, P.tolPCG, SiU.get(), nullptr);
            lap("trace: PCG of the nrun columns");
            std::vector<const double *> a, b;
            for (int i = 0; i < K; i++) for (int j = 0; j < p; j++) { a.push_back(SiX() + (size_t)j * N); b.push_back(u + (size_t)i * N); }
            hvec d(a.size());
            dot_pairs(c, a, b, d.data());
            for (int i = 0; i < K; i++) {
                hvec t = matvec(cov, hvec(d.begin() + (size_t)i * p, d.begin() + (size_t)(i + 1) * p));
                for (double &x : t) x = -x;
                lincomb(c, PU.get() + (size_t)i * N, 1.0, SiU.get() + (size_t)i * N, SiX(), N, t);
            }
            a.clear(); b.clear();
            for (int i = 0; i < K; i++) { a.push_back(au + (size_t)i * N); b.push_back(PU.get() + (size_t)i * N); }
            for (int i = 0; i < K; i++) { a.push_back(u + (size_t)i * N); b.push_back(PU.get() + (size_t)i * N); }
            d.resize(a.size());
            dot_pairs(c, a, b, d.data());
            for (int i = 0; i < K; i++) { buf[nrunStart + i] = d[i]; buf0[nrunStart + i] = d[K + i]; }
            traceCV = calcCV(buf);
            traceCV0 = quant ? calcCV(buf0) : 0;
            if (traceCV > P.traceCVcutoff || (quant && traceCV0 > P.traceCVcutoff)) {
                nrunStart = nrunEnd;
                nrunEnd += 10;
                buf.resize(nrunEnd, 0.0); buf0.resize(nrunEnd, 0.0);
                c.printf("CV for trace random estimator using %d runs is %g > %g\n", P.nrun, traceCV, P.traceCVcutoff);
                c.printf("try %d runs ...\n", nrunEnd);
            }
        }
        double m1 = 0, m0 = 0;
        for (double v : buf) m1 += v;
        for (double v : buf0) m0 += v;
        trace1 = m1 / buf.size();
        trace0 = m0 / buf0.size();
    }

    // get_AI_score, :817-833
    void get_AI_score(const double tau[2], double &YPAPY, double &Trace, double &AI) {
        project(PY.get(), SiY(), Y.get());
        grm_mv_device(c, PY.get(), APY.get(), 1);
        YPAPY = dot1(PY.get(), APY.get());
        double t0;
        get_trace(tau, false, t0, Trace);
        pcg_solve(c, pcg, W.get(), tau[0], tau[1], APY.get(), 1, P.maxiterPCG, P.tolPCG, PAPY1.get(), nullptr);
        project(PAPY.get(), PAPY1.get(), PAPY1.get());
        AI = dot1(APY.get(), PAPY.get());
    }

    // get_AI_score_q, :836-862
    void get_AI_score_q(const double tau[2], double YPAPY[2], double Trace[2], double AI[4]) {
        PA0PY1.ensure(N); PA0PY.ensure(N);
        project(PY.get(), SiY(), Y.get());            // A0PY == PY
        grm_mv_device(c, PY.get(), APY.get(), 1);
        YPAPY[0] = dot1(PY.get(), APY.get());
        YPAPY[1] = dot1(PY.get(), PY.get());
        get_trace(tau, true, Trace[0], Trace[1]);
        // the two solves of :854 and :857 are independent: one 2-RHS batch
        DevBuf<double> &two = PU;  // reuse as N x 2 scratch: [A0PY | APY] -> solutions
        two.ensure((size_t)N * 2); SiU.ensure((size_t)N * 2);
        copy(two.get(), PY.get(), N);
        copy(two.get() + N, APY.get(), N);
        pcg_solve(c, pcg, W.get(), tau[0], tau[1], two.get(), 2, P.maxiterPCG, P.tolPCG, SiU.get(), nullptr);
        copy(PA0PY1.get(), SiU.get(), N);
        copy(PAPY1.get(), SiU.get() + N, N);
        project(PA0PY.get(), PA0PY1.get(), PA0PY1.get());
        project(PAPY.get(), PAPY1.get(), PAPY1.get());
        double d[3];
        dot_pairs(c, {PY.get(), APY.get(), PY.get()}, {PA0PY.get(), PAPY.get(), PAPY.get()}, d);
        AI[0] = d[0]; AI[3] = d[1]; AI[1] = AI[2] = d[2];
    }

    // fitglmmaiRPCG, :866-895
    void fit_tau_binary(const double in_tau[2], double tau[2]) {
        double YPAPY, Trace, AI;
        get_AI_score(in_tau, YPAPY, Trace, AI);
        const double Dtau = (YPAPY - Trace) / AI;
        tau[0] = in_tau[0];
        tau[1] = in_tau[1] + Dtau;
        for (int i = 0; i < 2; i++) if (tau[i] < P.tol) tau[i] = 0;
        double step = 1.0;
        while (tau[1] < 0.0) { step *= 0.5; tau[1] = in_tau[1] + step * Dtau; }
        for (int i = 0; i < 2; i++) if (tau[i] < P.tol) tau[i] = 0;
    }

    // fitglmmaiRPCG_q, :898-928
    void fit_tau_quant(const double in_tau[2], double tau[2]) {
        const bool zero_v[2] = {in_tau[0] < P.tol, in_tau[1] < P.tol};
        double YPAPY[2], Trace[2], AI[4];
        get_AI_score_q(in_tau, YPAPY, Trace, AI);
        double s0 = YPAPY[1] - Trace[0], s1 = YPAPY[0] - Trace[1];
        double a = AI[0], b = AI[2], cc = AI[1], d = AI[3];   // solve(AI, score), 2x2 with partial pivoting
        if (fabs(cc) > fabs(a)) { std::swap(a, cc); std::swap(b, d); std::swap(s0, s1); }
        const double l = cc / a;
        double Dtau[2];
        Dtau[1] = (s1 - l * s0) / (d - l * b);
        Dtau[0] = (s0 - b * Dtau[1]) / a;
        for (int i = 0; i < 2; i++) { tau[i] = in_tau[i] + Dtau[i]; if (zero_v[i] && tau[i] < P.tol) tau[i] = 0; }
        double step = 1.0;
        while (tau[0] < 0.0 || tau[1] < 0.0) {
            step *= 0.5;
            for (int i = 0; i < 2; i++) { tau[i] = in_tau[i] + step * Dtau[i]; if (zero_v[i] && tau[i] < P.tol) tau[i] = 0; }
        }
        for (int i = 0; i < 2; i++) if (tau[i] < P.tol) tau[i] = 0;
    }
};

sgb_param checked_param(const sgb_param *P) {
    if (!P) throw Error(SGB_ERR_INVALID, "param is NULL");
    if (P->nrun < 2 || P->maxiter < 1 || P->maxiterPCG < 1) throw Error(SGB_ERR_INVALID, "invalid param values");
    return *P;
}

}  // namespace

// saige_fit_AI_PCG_binary (:949-1099) / saige_fit_AI_PCG_quant (:1103-1248)
void fit_AI_PCG(Context &c, bool quant, const sgb_fit0 *f, const double *hX, const double tau_in[2], const sgb_param *Pin,
                sgb_glmm *out) {
    c.require_stored();
    if (!f || !hX || !tau_in || !out) throw Error(SGB_ERR_INVALID, "NULL argument");
    if (f->n != c.N) throw Error(SGB_ERR_INVALID, "fit0$y length differs from the number of stored samples");
    if (f->p < 1 || f->p > kMaxCoef - 1) throw Error(SGB_ERR_INVALID, "unsupported number of fixed-effect columns");
    if (f->family != SGB_FAMILY_BINOMIAL && f->family != SGB_FAMILY_GAUSSIAN) throw Error(SGB_ERR_INVALID, "unsupported family");
    const sgb_param P = checked_param(Pin);
    const char *indent = P.indent ? P.indent : "";
    const int64_t N = c.N;
    const int p = f->p;
    const double tol = P.tol, tol_inv_2 = 1 / (tol * tol);
    PhaseTimer timer;
    const double alloc_s0 = g_alloc_seconds;
    const long alloc_n0 = g_alloc_calls;
    timer.st = &c.stats;
    timer.w0 = c.stats.host_wait_s;
    g_phase = getenv("SGB_FIT_TIMING") ? &timer : nullptr;
    Solver S(c, N, p, f->family, P);
    lap("set-up: construct");
    S.init(hX, f->y, f->offset);
    lap("set-up: init (uploads of X, y; workspaces)");
    // eta, mu of the glm fit; Y at :983/:1138 is recomputed inside get_coeff, so only eta is needed here
    c.h2d(S.eta_acc.get(), f->linear_predictors, sizeof(double) * N);
    c.h2d(S.mu_acc.get(), f->fitted_values, sizeof(double) * N);
    hvec alpha0(f->coefficients, f->coefficients + p), alpha = alpha0;
    hmat cov(p, p);
    double tau[2] = {tau_in[0], tau_in[1]}, tau0[2] = {tau_in[0], tau_in[1]};
    const bool no_iteration = !quant && P.no_iteration;
    if (P.verbose && !no_iteration) {
        c.printf("%sInitial variance component estimates, tau:\n", indent);
        c.printf("%s    Sigma_E: %g, Sigma_G: %g\n", indent, tau[0], tau[1]);
    }
    DevBuf<double> &eta0 = S.ws.eta0;
    eta0.ensure(N);
    S.copy(eta0.get(), S.eta_acc.get(), N);
    c.sync();
    lap("set-up and uploads");
    if (!no_iteration) S.prefetch_draws();
    S.get_coeff(tau, alpha0, eta0.get());
    c.sync();
    lap("get_coeff");
    int iter = 1;
    bool converged;
    if (no_iteration) {   // :1004-1014
        alpha = S.alpha; cov = S.cov;
        S.copy(S.eta_acc.get(), S.eta.get(), N);
        S.copy(S.mu_acc.get(), S.mu.get(), N);
        converged = true;
    } else {
        if (!quant) {
            double YPAPY, Trace, AI;
            S.get_AI_score(tau, YPAPY, Trace, AI);
            c.sync();
            lap("AI score + trace");
            tau[1] = std::max(0.0, tau0[1] + tau0[1] * tau0[1] * (YPAPY - Trace) / (double)N);   // :1024
        } else {
            double YPAPY[2], Trace[2], AI[4];
            S.get_AI_score_q(tau, YPAPY, Trace, AI);
            tau[0] = std::max(0.0, tau0[0] + tau0[0] * tau0[0] * (YPAPY[1] - Trace[0]) / (double)N);   // :1166-1167
            tau[1] = std::max(0.0, tau0[1] + tau0[1] * tau0[1] * (YPAPY[0] - Trace[1]) / (double)N);
        }
        for (; iter <= P.maxiter; iter++) {
            if (P.verbose) {
                c.printf("%sIteration %d:\n", indent, iter);
                print_vec(c, indent, "    tau: ", tau, 2);
                print_vec(c, indent, "    fixed coeff: ", alpha.data(), p);
            }
            alpha0 = S.alpha;
            tau0[0] = tau[0]; tau0[1] = tau[1];
            S.copy(eta0.get(), S.eta_acc.get(), N);   // eta0 = eta  (:1036)
            for (int itry = 1; itry <= 11; itry++) {
                S.get_coeff(tau0, alpha0, eta0.get());
                if (g_phase) { c.sync(); lap("get_coeff"); }
                if (!quant) S.fit_tau_binary(tau0, tau); else S.fit_tau_quant(tau0, tau);
                if (g_phase) { c.sync(); lap("AI score + trace"); }
                if (std::max(tau[0], tau[1]) > tol_inv_2) {
                    if (itry <= 10) {
                        if (quant && P.verbose) print_vec(c, indent, "tau: ", tau, 2, false);
                        if (!quant && P.verbose) print_vec(c, indent, "    tau: ", tau, 2, false);
                        tau0[1] *= 0.5;
                        if (P.verbose) {
                            c.printf(", large variance estimate observed, retry (%d) ...\n", itry);
                            print_vec(c, indent, "    set new tau: ", tau0, 2);
                        }
                        continue;
                    }
                    if (P.verbose) print_vec(c, indent, "tau: ", tau, 2);
                    throw Error(SGB_ERR_OVERFLOW, "Large variance estimate observed in the iterations, model not converged!");
                }
                break;
            }
            cov = S.cov; alpha = S.alpha;
            S.copy(S.eta_acc.get(), S.eta.get(), N);
            S.copy(S.mu_acc.get(), S.mu.get(), N);
            if (!quant) {
                if (tau[1] == 0) break;
            } else if (tau[0] <= 0) {
                print_vec(c, indent, "    tau: ", tau, 2);
                throw Error(SGB_ERR_OVERFLOW, "Sigma_E = 0, model not converged!");
            }
            double mx = 0;
            for (int k = 0; k < 2; k++) mx = std::max(mx, fabs(tau[k] - tau0[k]) / (fabs(tau[k]) + fabs(tau0[k]) + tol));
            if (mx < tol) break;
        }
        S.get_coeff(tau, alpha0, eta0.get());   // :1075 / :1224
        cov = S.cov; alpha = S.alpha;
        S.copy(S.eta_acc.get(), S.eta.get(), N);
        S.copy(S.mu_acc.get(), S.mu.get(), N);
        converged = iter <= P.maxiter;
    }
    if (P.verbose && !no_iteration) {
        print_vec(c, indent, "Final tau: ", tau, 2);
        print_vec(c, indent, "    fixed coeff: ", alpha.data(), p);
    }
    for (int i = 0; i < p; i++) out->coefficients[i] = alpha[i];
    out->tau[0] = tau[0]; out->tau[1] = tau[1];
    S.download(out->linear_predictors, S.eta_acc.get(), N);
    S.download(out->fitted_values, S.mu_acc.get(), N);
    c.sync();
    for (int64_t i = 0; i < N; i++) out->residuals[i] = f->y[i] - out->fitted_values[i];
    for (int i = 0; i < p * p; i++) out->cov[i] = cov.a[i];
    out->converged = converged ? 1 : 0;
    if (g_phase) {
        lap("download");
        std::string line = "fit timing (wall s, of which waiting for the GPU):";
        for (auto &kv : timer.t) {
            char b[128];
            snprintf(b, sizeof(b), " %s %.3f (%.3f);", kv.first.c_str(), kv.second.first, kv.second.second);
            line += b;
        }
        c.printf("%s\n", line.c_str());
        c.printf("fit timing: %ld cudaMalloc / cudaFree calls in this fit, %.3f s (the workspaces are kept by the context between fits)\n",
                 g_alloc_calls - alloc_n0, g_alloc_seconds - alloc_s0);
        g_phase = nullptr;
    }
}

// saige_calc_var_ratio_binary (:1255-1362) / _quant (:1366-1474)
void calc_var_ratio(Context &c, bool quant, const sgb_fit0 *f, const double tau_in[2], const sgb_noK *noK,
                    const sgb_param *Pin, const int32_t *marker_list, int64_t n_marker, sgb_var_ratio *out) {
    c.require_stored();
    if (!f || !tau_in || !noK || !marker_list || !out) throw Error(SGB_ERR_INVALID, "NULL argument");
    if (f->n != c.N) throw Error(SGB_ERR_INVALID, "fit0$y length differs from the number of stored samples");
    const sgb_param P = checked_param(Pin);
    const int64_t N = c.N;
    const int p = noK->p;
    if (p < 1 || p > kMaxCoef - 1) throw Error(SGB_ERR_INVALID, "unsupported number of columns in obj.noK$X1");
    const double tau[2] = {tau_in[0], tau_in[1]};
    int num_marker = P.num_marker;
    Solver S(c, N, p, f->family, P);
    S.init(noK->X1, f->y, nullptr);
    DevBuf<double> (&vr)[10] = S.ws.vr;
    DevBuf<double> &eta = vr[0], &mu = vr[1], &XVt = vr[2], &XXVX_inv = vr[3], &SiX1 = vr[4], &G0 = vr[5], &G = vr[6], &SiG = vr[7], &adj = vr[8];
    S.upload(eta, f->linear_predictors, N);
    S.upload(mu, f->fitted_values, N);
    family_weights(c, f->family, eta.get(), mu.get(), S.W.get());   // W from the *glm* fit (:1281-1284)
    // XV is p x N: keep its transpose (N x p) so that XV %*% G0 is p column dots
    {
        std::vector<double> t((size_t)N * p);
        for (int64_t i = 0; i < N; i++) for (int j = 0; j < p; j++) t[(size_t)j * N + i] = noK->XV[(size_t)i * p + j];
        S.upload(XVt, t.data(), t.size());
        c.sync();
    }
    S.upload(XXVX_inv, noK->XXVX_inv, (size_t)N * p);
    // Sigma_iX = get_sigma_X(W, tau, X1)  (:1287)
    SiX1.ensure((size_t)N * p);
    pcg_solve(c, S.pcg, S.W.get(), tau[0], tau[1], S.X.get(), p, P.maxiterPCG, P.tolPCG, SiX1.get(), nullptr);
    // mat_inv(X1' Sigma_iX) does not depend on the marker (:1322)
    hmat Minv;
    {
        std::vector<const double *> a, b;
        for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) { a.push_back(S.X.get() + (size_t)i * N); b.push_back(SiX1.get() + (size_t)j * N); }
        hvec d(a.size());
        dot_pairs(c, a, b, d.data());
        hmat XtS(p, p);
        for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) XtS(i, j) = d[(size_t)j * p + i];
        Minv = mat_inv(c, XtS);
    }
    // var2 weights: mu (1 - mu) for binary (:1325), 1 for quantitative (:1436)
    DevBuf<double> &wgt = vr[9];
    if (!quant) {
        wgt.ensure(N);
        std::vector<double> h(N);
        for (int64_t i = 0; i < N; i++) h[i] = f->fitted_values[i] * (1 - f->fitted_values[i]);
        c.h2d(wgt.get(), h.data(), sizeof(double) * N);
        c.sync();
    }
    double ratioCV = P.ratioCVcutoff + 0.1;
    int num_tested = 0;
    int64_t snp_idx = 0;
    hvec lst_ratio;
    struct Cand { int32_t id; double AF, AC; bool flip; double imp; int64_t local; };
    while (ratioCV > P.ratioCVcutoff && snp_idx < n_marker) {
        // pick the next (num_marker - num_tested) markers that pass the integer AC > 20 test, in rand_index order
        std::vector<Cand> batch;
        while (num_tested + (int)batch.size() < num_marker && snp_idx < n_marker) {
            const int32_t i_snp = marker_list[snp_idx++];
            if (i_snp < 1 || i_snp > c.M_total) throw Error(SGB_ERR_INVALID, "marker index out of range");
            // f64_af_ac_impute (vectorization.cpp:186-205) on integer counts over samples < N
            int32_t cnt[2] = {0, 0};
            const int64_t local = (int64_t)i_snp - 1 - c.var_offset;
            if (local >= 0 && local < c.M) { cnt[0] = c.h_cnt_num[local]; cnt[1] = c.h_cnt_sum[local]; }
            if (c.world > 1) {
                double t[2] = {(double)cnt[0], (double)cnt[1]};
                c.red_out.ensure(256);
                c.h2d(c.red_out.get(), t, sizeof(t));
                comm_allreduce_sum(c, c.red_out.get(), 2);
                c.d2h(t, c.red_out.get(), sizeof(t));
                c.sync();
                cnt[0] = (int32_t)t[0]; cnt[1] = (int32_t)t[1];
            }
            const int Num = cnt[0];
            double AC = (double)cnt[1];
            double AF = (Num > 0) ? (AC / (2 * Num)) : NAN;
            const double imp = AF * 2;
            bool flip = false;
            if (AF > 0.5) { flip = true; AC = 2 * Num - AC; AF = 1 - AF; }
            if (AC <= 20) continue;   // :1316
            batch.push_back(Cand{i_snp, AF, AC, flip, imp, local});
        }
        const int K = (int)batch.size();
        if (K > 0) {
            if (num_tested + K > out->capacity) throw Error(SGB_ERR_INVALID, "var.ratio output capacity exceeded");
            G0.ensure((size_t)N * K); G.ensure((size_t)N * K); SiG.ensure((size_t)N * K); adj.ensure(N);
            for (int i = 0; i < K; i++) {
                double *g0 = G0.get() + (size_t)i * N;
                if (batch[i].local >= 0 && batch[i].local < c.M) {
                    decode_variant(c, batch[i].local, g0);   // get_geno_ds (:1305)
                    impute_flip(c, g0, batch[i].imp, batch[i].flip);
                } else {
                    SGB_CUDA(cudaMemsetAsync(g0, 0, sizeof(double) * N, c.stream));
                }
            }
            if (c.world > 1) comm_allreduce_sum(c, G0.get(), (size_t)N * K);
            // G = G0 - XXVX_inv (XV G0)   (:1319)
            std::vector<const double *> a, b;
            for (int i = 0; i < K; i++) for (int j = 0; j < p; j++) { a.push_back(XVt.get() + (size_t)j * N); b.push_back(G0.get() + (size_t)i * N); }
            hvec d(a.size());
            dot_pairs(c, a, b, d.data());
            for (int i = 0; i < K; i++) {
                hvec t(d.begin() + (size_t)i * p, d.begin() + (size_t)(i + 1) * p);
                for (double &x : t) x = -x;
                lincomb(c, G.get() + (size_t)i * N, 1.0, G0.get() + (size_t)i * N, XXVX_inv.get(), N, t);
            }
            pcg_solve(c, S.pcg, S.W.get(), tau[0], tau[1], G.get(), K, P.maxiterPCG, P.tolPCG, SiG.get(), nullptr);   // :1321
            // X1' Sigma_iG (p dots per marker) and G' Sigma_iG
            a.clear(); b.clear();
            for (int i = 0; i < K; i++) for (int j = 0; j < p; j++) { a.push_back(S.X.get() + (size_t)j * N); b.push_back(SiG.get() + (size_t)i * N); }
            for (int i = 0; i < K; i++) { a.push_back(G.get() + (size_t)i * N); b.push_back(SiG.get() + (size_t)i * N); }
            d.resize(a.size());
            dot_pairs(c, a, b, d.data());
            hvec var2(K);
            weighted_sumsq(c, quant ? nullptr : wgt.get(), G.get(), N, K, var2.data());
            for (int i = 0; i < K; i++) {
                hvec t = matvec(Minv, hvec(d.begin() + (size_t)i * p, d.begin() + (size_t)(i + 1) * p));
                lincomb(c, adj.get(), 0.0, nullptr, SiX1.get(), N, t);   // adj = Sigma_iX Minv X1' Sigma_iG
                const double g_adj = S.dot1(G.get() + (size_t)i * N, adj.get());
                const double AC = batch[i].AC;
                const double var1 = (d[(size_t)K * p + i] - g_adj) / AC;
                const double v2 = var2[i] / AC;   // g = G / sqrt(AC): sum(w g g) = sum(w G G) / AC
                const double ratio = var1 / v2;
                const int r = num_tested++;
                out->id[r] = batch[i].id; out->maf[r] = batch[i].AF; out->mac[r] = AC;
                out->var1[r] = var1; out->var2[r] = v2; out->ratio[r] = ratio;
                lst_ratio.push_back(ratio);
                if (P.verbose)
                    c.printf("%6d, maf: %0.4f, mac: %g,\tratio: %0.4f (var1: %.3g, var2: %.3g)\n", num_tested, batch[i].AF, AC,
                             ratio, var1, v2);
            }
        }
        if (lst_ratio.size() < 2) break;
        ratioCV = calcCV(lst_ratio);
        if (ratioCV > P.ratioCVcutoff) {
            if (P.verbose)
                c.printf("CV for variance ratio estimate using %d markers is %g > ratioCVcutoff (%g), try more markers ...\n",
                         num_marker, ratioCV, P.ratioCVcutoff);
            num_marker += 10;
        }
    }
    out->n = num_tested;
}


// saige_GxG_snp_bin, saige_fitnull.cpp:1480-1558: score test of one interaction term under the fitted mixed model, with the
// full saddle-point approximation.  Same building blocks as the variance-ratio markers (W, Sigma^-1 X, covariate
// adjustment, one PCG solve, the adj term); the saddle-point step runs on the device vectors (saddle_prob_dense).
void gxg_snp_bin(Context &c, const sgb_fit0 *f, const double tau_in[2], const double *inter_term, const sgb_noK *noK,
                 const sgb_param *Pin, int verbose, sgb_gxg *out) {
    c.require_stored();
    if (!f || !tau_in || !inter_term || !noK || !out) throw Error(SGB_ERR_INVALID, "NULL argument");
    if (f->n != c.N) throw Error(SGB_ERR_INVALID, "fit0$y length differs from the number of stored samples");
    const sgb_param P = checked_param(Pin);
    const int64_t N = c.N;
    const int p = noK->p;
    if (p < 1 || p > kMaxCoef - 1) throw Error(SGB_ERR_INVALID, "unsupported number of columns in obj.noK$X1");
    const double tau[2] = {tau_in[0], tau_in[1]};
    Solver S(c, N, p, f->family, P);
    S.init(noK->X1, f->y, nullptr);
    DevBuf<double> (&vr)[10] = S.ws.vr;
    DevBuf<double> &eta = vr[0], &mu = vr[1], &XVt = vr[2], &XXVX_inv = vr[3], &SiX1 = vr[4], &G0 = vr[5], &G = vr[6], &SiG = vr[7], &adj = vr[8], &wgt = vr[9];
    S.upload(eta, f->linear_predictors, N);
    S.upload(mu, f->fitted_values, N);
    family_weights(c, f->family, eta.get(), mu.get(), S.W.get());   // :1499-1502
    std::vector<double> t((size_t)N * p), h(N);
    for (int64_t i = 0; i < N; i++) for (int j = 0; j < p; j++) t[(size_t)j * N + i] = noK->XV[(size_t)i * p + j];
    S.upload(XVt, t.data(), t.size());
    S.upload(XXVX_inv, noK->XXVX_inv, (size_t)N * p);
    int64_t n_nonzero = 0;
    for (int64_t i = 0; i < N; i++) {
        h[i] = f->fitted_values[i] * (1 - f->fitted_values[i]);
        if (inter_term[i] != 0) n_nonzero++;
    }
    S.upload(wgt, h.data(), N);
    S.upload(G0, inter_term, N);
    c.sync();
    // Sigma_iX = get_sigma_X(W, tau, X1)  (:1507)
    SiX1.ensure((size_t)N * p);
    pcg_solve(c, S.pcg, S.W.get(), tau[0], tau[1], S.X.get(), p, P.maxiterPCG, P.tolPCG, SiX1.get(), nullptr);
    // G = G0 - XXVX_inv (XV G0)  (:1521)
    G.ensure(N); SiG.ensure(N); adj.ensure(N);
    {
        hvec d = S.cols_dot(XVt.get(), p, G0.get());
        for (double &x : d) x = -x;
        lincomb(c, G.get(), 1.0, G0.get(), XXVX_inv.get(), N, d);
    }
    pcg_solve(c, S.pcg, S.W.get(), tau[0], tau[1], G.get(), 1, P.maxiterPCG, P.tolPCG, SiG.get(), nullptr);   // :1525
    // adj = Sigma_iX mat_inv(X1' Sigma_iX) X1' Sigma_iG  (:1527)
    {
        std::vector<const double *> a, b;
        for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) { a.push_back(S.X.get() + (size_t)i * N); b.push_back(SiX1.get() + (size_t)j * N); }
        hvec d(a.size());
        dot_pairs(c, a, b, d.data());
        hmat XtS(p, p);
        for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) XtS(i, j) = d[(size_t)j * p + i];
        hvec coef = matvec(mat_inv(c, XtS), S.cols_dot(S.X.get(), p, SiG.get()));
        lincomb(c, adj.get(), 0.0, nullptr, SiX1.get(), N, coef);
    }
    // :1529-1536
    double d[4];
    dot_pairs(c, {S.y.get(), mu.get(), G.get(), G.get()}, {G.get(), G.get(), SiG.get(), adj.get()}, d);
    double var2 = 0;
    weighted_sumsq(c, wgt.get(), G.get(), N, 1, &var2);
    const double q = d[0], m1 = d[1], Tstat = q - m1;
    const double var1 = d[2] - d[3];
    const double beta = Tstat / var1;
    const double qtilde = Tstat / sqrt(var1) * sqrt(var2) + m1;
    double pval, pnorm;
    bool converged;
    saddle_prob_dense(c, G.get(), mu.get(), N, qtilde, m1, var2, 2.0, &pval, &pnorm, &converged);   // :1539-1542
    out->beta = beta;
    out->SE = fabs(beta / qnorm_host(pval / 2));
    out->n_nonzero = n_nonzero;
    out->pval = pval;
    out->p_norm = pnorm;
    out->converged = converged ? 1 : 0;
    out->tau_G = tau[1];
    if (verbose)
        c.printf("    Nonzero #: %lld(%.3g%%), beta: %.6g, SE: %.6g, pval: %.6g, pnorm: %.6g, tau_G: %.5g\n", (long long)n_nonzero,
                 100.0 * n_nonzero / N, out->beta, out->SE, pval, pnorm, tau[1]);
}

}  // namespace sgb
