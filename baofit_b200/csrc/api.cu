// C ABI of libbaofit_b200.so (include/baofit_b200.h): object lifetime, one-time setup, and the
// per-batch kernel pipelines. No CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "host_setup.h"
#include "kernels.h"

namespace bf {
static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
bool pdl_enabled() {
    static const bool on = std::getenv("BAOFIT_NO_PDL") == nullptr;
    return on;
}
}  // namespace bf

using namespace bf;

#define BF_FAIL(code, msg)  \
    do {                    \
        bf::set_error(msg); \
        return (code);      \
    } while (0)

extern "C" const char *bf_last_error(void) { return bf::g_error.c_str(); }
extern "C" const char *bf_version(void) { return "baofit_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
extern "C" int bf_ctx_create(int device, bf_ctx **out) {
    if (!out) BF_FAIL(BF_ERR_OPTIONS, "bf_ctx_create: null output pointer");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        BF_FAIL(BF_ERR_ANALYSIS, std::string("bf_ctx_create: no usable CUDA device (") + cudaGetErrorString(e) +
                                     "); this library has no CPU fallback");
    if (device < 0 || device >= ndev) BF_FAIL(BF_ERR_OPTIONS, "bf_ctx_create: invalid device index");
    BF_CUDA_OK(cudaSetDevice(device));
    bf_ctx *c = new bf_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
        (e = cudaMallocHost((void **)&c->pinned_flag, 256)) != cudaSuccess) {
        bf_ctx_destroy(c);
        BF_FAIL(BF_ERR_ANALYSIS, std::string("bf_ctx_create: ") + cudaGetErrorString(e));
    }
    c->sm_count = prop.multiProcessorCount;
    c->no_graphs = std::getenv("BAOFIT_NO_GRAPHS") != nullptr;
    for (auto &sl : c->async_slot)
        if ((e = cudaEventCreateWithFlags(&sl.uploaded, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess) {
            bf_ctx_destroy(c);
            BF_FAIL(BF_ERR_ANALYSIS, std::string("bf_ctx_create: ") + cudaGetErrorString(e));
        }
    {   // keep stream-ordered allocations (toy tables) cached in the device pool between studies
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    *out = c;
    return BF_OK;
}

static void purge_graphs(bf_ctx *ctx, unsigned long long model_serial, const void *data);

extern "C" int bf_ctx_destroy(bf_ctx *c) {
    if (!c) return BF_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto &r : c->prof_pending) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    c->prof_pending.clear();
    if (c->ws) cudaFree(c->ws);
    purge_graphs(c, 0, nullptr);
    if (c->io) cudaFree(c->io);
    if (c->stage) cudaFreeHost(c->stage);
    if (c->aux) cudaFree(c->aux);
    if (c->mp) cudaFree(c->mp);
    if (c->chi2_part) cudaFree(c->chi2_part);
    if (c->pinned_flag) cudaFreeHost(c->pinned_flag);
    for (auto &sl : c->async_slot) {
        if (sl.uploaded) cudaEventDestroy(sl.uploaded);
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.buf) cudaFree(sl.buf);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return BF_OK;
}

extern "C" long long bf_ctx_launch_count(const bf_ctx *c) { return c ? c->launches : 0; }
extern "C" void *bf_ctx_stream(bf_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int bf_ctx_synchronize(bf_ctx *c) {
    if (!c) BF_FAIL(BF_ERR_OPTIONS, "null context");
    std::lock_guard<std::recursive_mutex> lock(c->api_mu);
    BF_CUDA_OK(cudaStreamSynchronize(c->stream));
    return BF_OK;
}

extern "C" int bf_ctx_set_profiling(bf_ctx *c, int enable) {
    if (!c) BF_FAIL(BF_ERR_OPTIONS, "null context");
    std::lock_guard<std::recursive_mutex> lock(c->api_mu);
    BF_CUDA_OK(cudaStreamSynchronize(c->stream));
    for (auto &r : c->prof_pending) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    c->prof_pending.clear();
    c->prof_acc.clear();
    c->profiling = enable != 0;
    return BF_OK;
}

static int prof_drain(bf_ctx *c) {
    BF_CUDA_OK(cudaStreamSynchronize(c->stream));
    for (auto &r : c->prof_pending) {
        float ms = 0;
        BF_CUDA_OK(cudaEventElapsedTime(&ms, r.e0, r.e1));
        auto &acc = c->prof_acc[r.name];
        acc.first += ms;
        acc.second += 1;
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    c->prof_pending.clear();
    return BF_OK;
}

extern "C" int bf_ctx_get_profile(bf_ctx *c, int max_entries, const char **names, double *ms_total,
                                  long long *launches) {
    if (!c) return 0;
    std::lock_guard<std::recursive_mutex> lock(c->api_mu);
    if (prof_drain(c)) return 0;
    int k = 0;
    for (auto &kv : c->prof_acc) {
        if (k < max_entries) {
            if (names) names[k] = kv.first.c_str();
            if (ms_total) ms_total[k] = kv.second.first;
            if (launches) launches[k] = kv.second.second;
        }
        ++k;
    }
    return k;
}

namespace {
struct ProfScope {
    bf_ctx *c;
    bf_ctx::ProfRec rec{};
    ProfScope(bf_ctx *ctx, const char *name) : c(ctx) {
        c->launches++;
        if (!c->profiling) return;
        rec.name = name;
        cudaEventCreate(&rec.e0);
        cudaEventCreate(&rec.e1);
        cudaEventRecord(rec.e0, c->stream);
    }
    ~ProfScope() {
        if (!c->profiling) return;
        cudaEventRecord(rec.e1, c->stream);
        c->prof_pending.push_back(rec);
    }
};
}  // namespace
#define BF_KERNEL(ctx, name, call) \
    do {                           \
        ProfScope _ps(ctx, name);  \
        call;                      \
    } while (0)

static int ensure(void **p, size_t *have, size_t need) {
    if (*have >= need) return BF_OK;
    if (*p) BF_CUDA_OK(cudaFree(*p));
    *p = nullptr;
    *have = 0;
    size_t want = need + need / 8;
    BF_CUDA_OK(cudaMalloc(p, want));
    *have = want;
    return BF_OK;
}

template <class T> static int upload(std::vector<void *> &owner, const T *host, size_t count, T **dev) {
    *dev = nullptr;
    if (count == 0) return BF_OK;
    BF_CUDA_OK(cudaMalloc((void **)dev, count * sizeof(T)));
    owner.push_back(*dev);
    BF_CUDA_OK(cudaMemcpy(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice));
    return BF_OK;
}

// ------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------
static void copy_bb(const bf_broadband_spec &s, DevBB &d) {
    d.enabled = s.enabled; d.idx_base = s.idx_base;
    d.r_min = s.r_min; d.r_max = s.r_max; d.r_step = s.r_step; d.r_denom = s.r_denom;
    d.mu_min = s.mu_min; d.mu_max = s.mu_max; d.mu_step = s.mu_step;
    d.rp_min = s.rp_min; d.rp_max = s.rp_max; d.rp_step = s.rp_step;
    d.rt_min = s.rt_min; d.rt_max = s.rt_max; d.rt_step = s.rt_step;
    d.z_min = s.z_min; d.z_max = s.z_max; d.z_step = s.z_step;
    d.inv_r0 = s.enabled ? 1.0 / s.r0 : 0.0;
    d.z0 = s.z0;
    d.inv_1pz0 = 1.0 / (1.0 + s.z0);
    d.plan_r = axis_plan(s.r_min, s.r_max, s.r_step);
    d.plan_rp = axis_plan(s.rp_min, s.rp_max, s.rp_step);
    d.plan_rt = axis_plan(s.rt_min, s.rt_max, s.rt_step);
}

static int bb_ncoef(const bf_broadband_spec &s) {
    auto cnt = [](int lo, int hi, int st) { return st > 0 && hi >= lo ? (hi - lo) / st + 1 : 0; };
    return cnt(s.z_min, s.z_max, s.z_step) * cnt(s.mu_min, s.mu_max, s.mu_step) * cnt(s.r_min, s.r_max, s.r_step) *
           cnt(s.rp_min, s.rp_max, s.rp_step) * cnt(s.rt_min, s.rt_max, s.rt_step);
}

static int make_spline_set(bf_model *m, int n, const double *x, const double *y0, const double *y2,
                           const double *y4, DevSplineSet &out) {
    if (n < 3 || !x || !y0 || !y2 || !y4) BF_FAIL(BF_ERR_MODEL, "BaoCorrelationModel: error while reading model interpolation data.");
    for (int j = 0; j + 1 < n; ++j)
        if (!(x[j + 1] > x[j])) BF_FAIL(BF_ERR_MODEL, "BaoCorrelationModel: interpolation knots must increase");
    std::vector<double> coef((size_t)3 * (n - 1) * 4);
    const double *ys[3] = {y0, y2, y4};
    for (int l = 0; l < 3; ++l) {
        cspline_interval_coefs(n, x, ys[l], coef.data() + (size_t)l * (n - 1) * 4);
        out.ylo[l] = ys[l][0];
        out.yhi[l] = ys[l][n - 1];
    }
    out.n = n;
    out.x0 = x[0];
    out.xlast = x[n - 1];
    double dx = (x[n - 1] - x[0]) / (n - 1);
    out.inv_dx = 1.0 / dx;
    out.uniform = 1;
    for (int j = 0; j < n; ++j)
        if (std::fabs(x[j] - (x[0] + j * dx)) > 1e-9 * dx) out.uniform = 0;
    double *dxp = nullptr, *dcoef = nullptr;
    int rc = upload(m->allocs, x, n, &dxp);
    if (rc) return rc;
    rc = upload(m->allocs, coef.data(), coef.size(), &dcoef);
    if (rc) return rc;
    out.x = dxp;
    out.coef = dcoef;
    return BF_OK;
}

extern "C" int bf_model_destroy(bf_model *m) {
    if (!m) return BF_OK;
    std::lock_guard<std::recursive_mutex> lock(m->ctx->api_mu);
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    purge_graphs(m->ctx, m->serial, nullptr);
    // the (model, data) plans built for this model die with it
    for (bf_data *d : m->ctx->live_data) {
        auto it = d->plans.find(m->serial);
        if (it == d->plans.end()) continue;
        for (void *p : it->second.allocs) cudaFree(p);
        d->plans.erase(it);
    }
    for (void *p : m->allocs) cudaFree(p);
    delete m;
    return BF_OK;
}

extern "C" int bf_model_npar(const bf_model *m) { return m ? m->dev.npar : 0; }

extern "C" int bf_model_create(bf_ctx *ctx, const bf_model_desc *d, bf_model **out) {
    if (!ctx || !d || !out) BF_FAIL(BF_ERR_OPTIONS, "bf_model_create: null argument");
    if (d->kind != BF_MODEL_RSPACE && d->kind != BF_MODEL_KSPACE && d->kind != BF_MODEL_KSPACE_FFT)
        BF_FAIL(BF_ERR_MODEL, "bf_model_create: unknown model kind");
    if (d->npar <= 0 || d->npar > 256) BF_FAIL(BF_ERR_MODEL, "bf_model_create: bad parameter count");
    auto chk = [&](int idx, int span) { return idx >= 0 && idx + span <= d->npar; };
    if (!chk(d->idx_beta, 1) || !chk(d->idx_bb, 1) || !chk(d->idx_gamma_bias, 1) || !chk(d->idx_gamma_beta, 1) ||
        !chk(d->idx_bao, 5))
        BF_FAIL(BF_ERR_MODEL, "AbsCorrelationModel: no linear bias parameters defined.");
    const bool cross = d->flags & BF_FLAG_CROSS_CORRELATION;
    if (cross && (!chk(d->idx_delta_v, 1) || !chk(d->idx_bias2, 1) || !chk(d->idx_bb2, 1)))
        BF_FAIL(BF_ERR_MODEL, "bf_model_create: cross-correlation needs delta-v, bias2, beta2*bias2");
    if ((d->flags & BF_FLAG_COMBINED_BIAS) && !chk(d->idx_beta_bias, 1)) BF_FAIL(BF_ERR_MODEL, "bf_model_create: combined-bias needs beta*bias");
    if ((d->flags & BF_FLAG_COMBINED_SCALE) && !chk(d->idx_comb_scale, 1)) BF_FAIL(BF_ERR_MODEL, "bf_model_create: combined-scale needs BAO aperp/apar");
    if (d->zref < 0) BF_FAIL(BF_ERR_MODEL, "AbsCorrelationModel: expected zref >= 0.");
    if (d->omega_matter < 0) BF_FAIL(BF_ERR_MODEL, "AbsCorrelationModel: expected OmegaMatter >= 0.");
    for (int s = 0; s < 4; ++s)
        if (d->bb[s].enabled) {
            const bf_broadband_spec &b = d->bb[s];
            if (b.r_max < b.r_min || b.r_step <= 0 || b.r_denom <= 0) BF_FAIL(BF_ERR_MODEL, "BroadbandModel: illegal r-parameter specification.");
            if (b.mu_max < b.mu_min || b.mu_step <= 0) BF_FAIL(BF_ERR_MODEL, "BroadbandModel: illegal mu-parameter specification.");
            if (b.mu_min < 0 || b.mu_max > 8) BF_FAIL(BF_ERR_MODEL, "BroadbandModel: only multipoles 0-8 are allowed in mu-parameter specification.");
            if (b.rp_max < b.rp_min || b.rp_step <= 0) BF_FAIL(BF_ERR_MODEL, "BroadbandModel: illegal rP-parameter specification.");
            if (b.rt_max < b.rt_min || b.rt_step <= 0) BF_FAIL(BF_ERR_MODEL, "BroadbandModel: illegal rT-parameter specification.");
            if (b.z_max < b.z_min || b.z_step <= 0) BF_FAIL(BF_ERR_MODEL, "BroadbandModel: illegal z-parameter specification.");
            if (!chk(b.idx_base, bb_ncoef(b))) BF_FAIL(BF_ERR_MODEL, "BroadbandModel: coefficients outside the parameter vector");
            auto big = [](int lo, int hi) { return lo < -31 || hi > 31; };
            if (big(b.r_min, b.r_max) || big(b.rp_min, b.rp_max) || big(b.rt_min, b.rt_max) || big(b.z_min, b.z_max) ||
                b.r_step > 31 || b.rp_step > 31 || b.rt_step > 31)
                BF_FAIL(BF_ERR_MODEL, "BroadbandModel: exponents beyond +-31 are not supported");
            if (!(b.r0 > 0)) BF_FAIL(BF_ERR_MODEL, "BroadbandModel: expected r0 > 0");
        }
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    bf_model *m = new bf_model();
    m->ctx = ctx;
    m->dzmin = d->dzmin;
    static std::atomic<unsigned long long> next_serial{1};
    m->serial = next_serial.fetch_add(1);
    DevModel &dm = m->dev;
    std::memset(&dm, 0, sizeof(dm));
    dm.kind = d->kind; dm.flags = d->flags; dm.npar = d->npar; dm.zref = d->zref; dm.omega_matter = d->omega_matter;
    dm.idx_beta = d->idx_beta; dm.idx_bb = d->idx_bb; dm.idx_gamma_bias = d->idx_gamma_bias;
    dm.idx_gamma_beta = d->idx_gamma_beta; dm.idx_delta_v = cross ? d->idx_delta_v : -1; dm.idx_bias2 = d->idx_bias2;
    dm.idx_bb2 = d->idx_bb2; dm.idx_beta_bias = (d->flags & BF_FLAG_COMBINED_BIAS) ? d->idx_beta_bias : -1;
    dm.idx_bao = d->idx_bao; dm.idx_comb_scale = d->idx_comb_scale; dm.idx_rad = d->idx_rad; dm.idx_nl = d->idx_nl;
    dm.idx_cont = d->idx_cont; dm.idx_bs = d->idx_bs; dm.idx_hcd = d->idx_hcd; dm.idx_uv = d->idx_uv;
    dm.idx_smgaus = d->idx_smgaus; dm.idx_smlor = d->idx_smlor; dm.idx_nlcorr = d->idx_nlcorr;
    for (int s = 0; s < 4; ++s) copy_bb(d->bb[s], dm.bb[s]);
    m->zref = d->zref;
    int rc = BF_OK;
    auto fail = [&](int code) { bf_model_destroy(m); return code; };

    if (d->kind == BF_MODEL_RSPACE) {
        if (d->idx_rad >= 0 && !chk(d->idx_rad, 4)) { bf::set_error("bf_model_create: bad radiation block"); return fail(BF_ERR_MODEL); }
        if ((rc = make_spline_set(m, d->n_fid, d->fid_r, d->fid_xi0, d->fid_xi2, d->fid_xi4, dm.fid))) return fail(rc);
        if ((rc = make_spline_set(m, d->n_nw, d->nw_r, d->nw_xi0, d->nw_xi2, d->nw_xi4, dm.nw))) return fail(rc);
        // Fast pass: sum_l L_l n_l xi_l(r) with the Kaiser factors n_0 = 1 + (2/3) ba + (1/5) bp, n_2 = (4/3) ba + (4/7) bp,
        // n_4 = (8/35) bp (AbsCorrelationModel.cc:196-204) is a polynomial in c1 = 2 ba and c2 = bp, i.e. the
        // basis-table form X_l0 + c1 X_l1 + c2 X_l2 of the k-space fast pass. Component 0 holds fiducial - nowiggles.
        static const bool generic_only = std::getenv("BAOFIT_RSPACE_GENERIC") != nullptr;
        const bool same_grid = dm.fid.uniform && dm.nw.uniform && d->n_fid == d->n_nw && d->fid_r[0] == d->nw_r[0] &&
                               d->fid_r[d->n_fid - 1] == d->nw_r[d->n_nw - 1];
        // (not with a multiplicative broadband: the radiation term is added after it, BaoCorrelationModel.cc:150-176,
        // while the fast pass adds it with the cosmological part)
        if (same_grid && !generic_only && d->metal_kind == BF_METAL_NONE && !d->dist_matrix && !(d->flags & BF_FLAG_RADIATION_MODEL) &&
            !d->bb[BF_BB_DIST_MUL].enabled &&
            (size_t)2 * d->n_fid * 9 * sizeof(double2) <= 200 * 1024) {
            const int n = d->n_fid;
            std::vector<double> cf((size_t)(n - 1) * 4);
            std::vector<double2> tabs((size_t)2 * n * 9, make_double2(0.0, 0.0));
            const double *fy[3] = {d->fid_xi0, d->fid_xi2, d->fid_xi4}, *ny[3] = {d->nw_xi0, d->nw_xi2, d->nw_xi4};
            // weights of (1, c1, c2) in n_l
            const double wgt[3][3] = {{1.0, 1.0 / 3.0, 1.0 / 5.0}, {0.0, 2.0 / 3.0, 4.0 / 7.0}, {0.0, 0.0, 8.0 / 35.0}};
            for (int l = 0; l < 3; ++l) {
                std::vector<double> yc[2] = {std::vector<double>(n), std::vector<double>(n)}, cc[2] = {std::vector<double>(n, 0.0), std::vector<double>(n, 0.0)};
                for (int which = 0; which < 2; ++which) {
                    const double *y = which == 0 ? fy[l] : ny[l];
                    cspline_interval_coefs(n, which == 0 ? d->fid_r : d->nw_r, y, cf.data());
                    for (int j = 0; j < n; ++j) { yc[which][j] = y[j]; cc[which][j] = j < n - 1 ? cf[(size_t)j * 4 + 2] : 0.0; }
                }
                for (int j = 0; j < n; ++j)
                    for (int q = 0; q < 3; ++q) {
                        tabs[((size_t)0 * n + j) * 9 + 3 * l + q] = make_double2(wgt[l][q] * (yc[0][j] - yc[1][j]), wgt[l][q] * (cc[0][j] - cc[1][j]));
                        tabs[((size_t)1 * n + j) * 9 + 3 * l + q] = make_double2(wgt[l][q] * yc[1][j], wgt[l][q] * cc[1][j]);
                    }
            }
            if ((rc = upload(m->allocs, tabs.data(), tabs.size(), &m->d_rs_tabs))) return fail(rc);
            dm.nr = n;
            dm.tr_r0 = d->fid_r[0];
            dm.tr_dr = (d->fid_r[n - 1] - d->fid_r[0]) / (n - 1);
            dm.tr_inv_dr = 1.0 / dm.tr_dr;
            dm.dilmin = 0.0;     // the r-space model has no dilation limits
            dm.dilmax = 1e300;
        }
    } else if (d->kind == BF_MODEL_KSPACE_FFT) {
        // baofit/BaoKSpaceFftCorrelationModel.cc:36-113
        if (!chk(d->idx_nl, 2) || !chk(d->idx_cont, 2)) { bf::set_error("bf_model_create: FFT model needs SigmaNL-perp, 1+f, cont-kc, cont-pc"); return fail(BF_ERR_MODEL); }
        if ((d->flags & BF_FLAG_FIT_NL_CORRECTION) && !chk(d->idx_nlcorr, 2)) { bf::set_error("bf_model_create: bad nlcorr block"); return fail(BF_ERR_MODEL); }
        if (d->n_pk < 8 || !d->pk_k || !d->pk_fid || !d->pk_nw) {
            bf::set_error("BaoKSpaceFftCorrelationModel: error while reading tabulated P(k) data.");
            return fail(BF_ERR_MODEL);
        }
        if (!(d->fft_spacing > 0) || d->fft_nx < 8 || d->fft_ny < 8 || d->fft_nz < 8 || d->fft_nx > 4096 || d->fft_ny > 4096 || d->fft_nz > 4096) {
            bf::set_error("BaoKSpaceFftCorrelationModel: bad 3D FFT grid (gridspacing > 0, 8 <= ngrid <= 4096)");
            return fail(BF_ERR_MODEL);
        }
        if (d->metal_kind != BF_METAL_NONE || d->dist_matrix) { bf::set_error("BaoKSpaceFftCorrelationModel has no metal model / distortion matrix"); return fail(BF_ERR_MODEL); }
        DevFft &f = dm.fft;
        f.nx = d->fft_nx; f.ny = d->fft_ny; f.nz = d->fft_nz;
        f.mx = f.nx / 2 + 1; f.my = f.ny / 2 + 1; f.mz = f.nz / 2 + 1;
        f.my_pad = round_up(f.my, 16);
        // fft_plane_kernel keeps two sets of k_par tables (4 vectors each) next to >= 2 pipeline stages in shared memory
        if ((size_t)2 * 4 * f.my_pad * sizeof(double) + 2 * (size_t)(kFftStageA + kFftStageT) * sizeof(double) > (size_t)226 * 1024) {
            bf::set_error("BaoKSpaceFftCorrelationModel: ngridy too large for the shared-memory pipeline (ngridy <= 4600)");
            return fail(BF_ERR_MODEL);
        }
        f.delta = d->fft_spacing; f.inv_delta = 1.0 / d->fft_spacing;
        f.norm = 1.0 / ((double)f.nx * f.ny * f.nz * f.delta * f.delta * f.delta);
        f.zcorr0 = d->zcorr0; f.zcorr1 = d->zcorr1; f.zcorr2 = d->zcorr2;
        const double plane_rmax = d->fft_rmax > 0 ? d->fft_rmax : d->rmax * d->dilmax;
        if (!(plane_rmax > 0)) { bf::set_error("BaoKSpaceFftCorrelationModel: need fft_rmax > 0 or rmax * dilmax > 0"); return fail(BF_ERR_MODEL); }
        const int np = (int)std::floor(plane_rmax / f.delta) + 3;
        f.npx = std::min(np, f.mx); f.npy = std::min(np, f.my);
        f.npx_pad = round_up(f.npx, 8); f.npy_pad = round_up(f.npy, 8);
        dm.sigma8 = d->sigma8;
        // P(k): natural cubic splines in ln k, power-law continuation through the end points
        const int n = d->n_pk;
        for (int j = 0; j + 1 < n; ++j)
            if (!(d->pk_k[j + 1] > d->pk_k[j]) || !(d->pk_k[j] > 0)) { bf::set_error("BaoKSpaceFftCorrelationModel: P(k) nodes must be positive and increasing"); return fail(BF_ERR_MODEL); }
        std::vector<double> lnk(n), coef((size_t)2 * (n - 1) * 4);
        for (int j = 0; j < n; ++j) lnk[j] = std::log(d->pk_k[j]);
        const double *tabs[2] = {d->pk_fid, d->pk_nw};
        auto slope = [](double k0, double p0, double k1, double p1) { return p0 * p1 <= 0 ? 0.0 : std::log(p1 / p0) / std::log(k1 / k0); };
        for (int w = 0; w < 2; ++w) {
            cspline_interval_coefs(n, lnk.data(), tabs[w], coef.data() + (size_t)w * (n - 1) * 4);
            f.pk_end[w][0] = tabs[w][0]; f.pk_end[w][1] = tabs[w][n - 1];
            f.pk_slope[w][0] = slope(d->pk_k[0], tabs[w][0], d->pk_k[1], tabs[w][1]);
            f.pk_slope[w][1] = slope(d->pk_k[n - 2], tabs[w][n - 2], d->pk_k[n - 1], tabs[w][n - 1]);
        }
        f.n_pk = n; f.pk_k0 = d->pk_k[0]; f.pk_k1 = d->pk_k[n - 1];
        double *p = nullptr;
        if ((rc = upload(m->allocs, lnk.data(), lnk.size(), &p))) return fail(rc);
        f.lnk = p;
        if ((rc = upload(m->allocs, coef.data(), coef.size(), &p))) return fail(rc);
        f.pk_coef = p;
        // cosine tables with exact argument reduction
        const long double two_pi = 6.283185307179586476925286766559005768L;
        const int nrf = f.npy_pad / 8, ncfr = f.npx_pad / 8;
        f.rtiles = (nrf + 8) / 9; f.ctiles = (ncfr + 8) / 9;
        f.wr = (nrf + f.rtiles - 1) / f.rtiles; f.ncf = (ncfr + f.ctiles - 1) / f.ctiles;
        const int nchunks = f.my_pad / kFftKC;
        std::vector<double> cosx(f.nx), cy((size_t)f.rtiles * nchunks * kFftStageA, 0.0);
        for (int t = 0; t < f.nx; ++t) cosx[t] = (double)cosl(two_pi * t / f.nx);
        for (int iy = 0; iy < f.npy; ++iy)
            for (int my = 0; my < f.my; ++my) {
                const int rt = iy / (f.wr * 8), row = iy - rt * f.wr * 8, ch = my / kFftKC, k = my - ch * kFftKC;
                cy[((size_t)(rt * nchunks + ch) * 72 + row) * kFftLdA + k] = (double)cosl(two_pi * (((long long)my * iy) % f.ny) / f.ny);
            }
        if ((rc = upload(m->allocs, cosx.data(), cosx.size(), &p))) return fail(rc);
        f.cosx = p;
        if ((rc = upload(m->allocs, cy.data(), cy.size(), &p))) return fail(rc);
        f.Cy = p;
        auto dev_alloc = [&](double **ptr, size_t count) {
            cudaError_t e = cudaMalloc((void **)ptr, count * sizeof(double));
            if (e != cudaSuccess) { bf::set_error(std::string("bf_model_create: cudaMalloc: ") + cudaGetErrorString(e)); return BF_ERR_ANALYSIS; }
            m->allocs.push_back(*ptr);
            return BF_OK;
        };
        if ((rc = dev_alloc(&m->d_fft_T, (size_t)2 * f.ctiles * nchunks * kFftStageT))) return fail(rc);
        if ((rc = dev_alloc(&m->d_fft_H, (size_t)6 * f.my * f.mx))) return fail(rc);
    } else {
        if (!chk(d->idx_nl, 2)) { bf::set_error("bf_model_create: k-space model needs SigmaNL-perp, 1+f"); return fail(BF_ERR_MODEL); }
        if ((d->flags & BF_FLAG_RADIATION_MODEL) && !chk(d->idx_rad, 4)) { bf::set_error("bf_model_create: bad radiation block"); return fail(BF_ERR_MODEL); }
        if ((d->flags & (BF_FLAG_BIN_SMOOTH | BF_FLAG_BIN_SMOOTH_ALT)) && !chk(d->idx_bs, 2)) { bf::set_error("bf_model_create: bad bin-smooth block"); return fail(BF_ERR_MODEL); }
        if ((d->flags & (BF_FLAG_HCD_MODEL | BF_FLAG_HCD_MODEL_ALT)) && !chk(d->idx_hcd, 3)) { bf::set_error("bf_model_create: bad HCD block"); return fail(BF_ERR_MODEL); }
        if ((d->flags & BF_FLAG_UVFLUCTUATION) && !chk(d->idx_uv, 3)) { bf::set_error("bf_model_create: bad UV block"); return fail(BF_ERR_MODEL); }
        if ((d->flags & BF_FLAG_SMOOTH_GAUSS) && !chk(d->idx_smgaus, 1)) { bf::set_error("bf_model_create: bad smooth-gauss block"); return fail(BF_ERR_MODEL); }
        if ((d->flags & BF_FLAG_SMOOTH_LORENTZ) && !chk(d->idx_smlor, 1)) { bf::set_error("bf_model_create: bad smooth-lorentz block"); return fail(BF_ERR_MODEL); }
        if ((d->flags & BF_FLAG_FIT_NL_CORRECTION) && !chk(d->idx_nlcorr, 2)) { bf::set_error("bf_model_create: bad nlcorr block"); return fail(BF_ERR_MODEL); }
        // baofit/BaoKSpaceCorrelationModel.cc:156-159
        if (d->rmin >= d->rmax) { bf::set_error("BaoKSpaceCorrelationModel: expected rmin < rmax."); return fail(BF_ERR_MODEL); }
        if (d->rmin <= 0) { bf::set_error("BaoKSpaceCorrelationModel: expected rmin > 0."); return fail(BF_ERR_MODEL); }
        if (d->dilmin > d->dilmax) { bf::set_error("BaoKSpaceCorrelationModel: expected dilmin <= dilmax."); return fail(BF_ERR_MODEL); }
        if (d->dilmin <= 0) { bf::set_error("BaoKSpaceCorrelationModel: expected dilmin > 0."); return fail(BF_ERR_MODEL); }
        if (d->n_pk < 8 || !d->pk_k || !d->pk_fid || !d->pk_nw || d->samples_per_decade < 1) {
            bf::set_error("BaoKSpaceCorrelationModel: error while reading model interpolation data.");
            return fail(BF_ERR_MODEL);
        }
        if (d->ell_max != 0 && d->ell_max != 2 && d->ell_max != 4) { bf::set_error("bf_model_create: ell-max must be 0, 2 or 4"); return fail(BF_ERR_MODEL); }
        HankelSetup hs;
        build_hankel_weights(d->n_pk, d->pk_k, d->pk_fid, d->pk_nw, d->rmin, d->rmax, d->dilmin, d->dilmax,
                             d->samples_per_decade, kFineSub, hs);
        if (hs.nr < 4 || hs.nk < 8) { bf::set_error("bf_model_create: k-space grids too small"); return fail(BF_ERR_MODEL); }
        dm.nk = hs.nk; dm.nk_pad = round_up(hs.nk, kGemmBK); dm.nr = hs.nr; dm.nr_pad = round_up(hs.nr, kGemmBM);
        dm.ell_max = d->ell_max; dm.tr_r0 = hs.r0; dm.tr_dr = hs.dr; dm.tr_inv_dr = 1.0 / hs.dr; dm.rlo = hs.rlo; dm.rhi = hs.rhi;
        dm.dilmin = d->dilmin; dm.dilmax = d->dilmax; dm.zeff = d->zeff; dm.sigma8 = d->sigma8;
        std::vector<double> wpad((size_t)6 * dm.nr_pad * dm.nk_pad, 0.0);
        for (int t = 0; t < 6; ++t)
            for (int j = 0; j < hs.nr; ++j)
                std::memcpy(&wpad[((size_t)t * dm.nr_pad + j) * dm.nk_pad], &hs.weights[((size_t)t * hs.nr + j) * hs.nk],
                            sizeof(double) * hs.nk);
        double *p = nullptr;
        if ((rc = upload(m->allocs, wpad.data(), wpad.size(), &p))) return fail(rc);
        dm.hankel = p;
        if ((rc = upload(m->allocs, hs.kc.data(), hs.kc.size(), &p))) return fail(rc);
        dm.kc = p;
        if ((rc = upload(m->allocs, hs.pk_kc.data(), hs.pk_kc.size(), &p))) return fail(rc);
        dm.pk_kc = p;
        gauss_legendre_01(kMuNodes, dm.mu_x, dm.mu_w);
        // constant (h, 4h, h) tridiagonal factors of the natural spline on the radial nodes
        int msz = hs.nr - 2;
        std::vector<double> thw(msz, 0.0), thd(msz, 0.0);
        double h = hs.dr, diag = 4.0 * h;
        thd[0] = 1.0 / diag;
        for (int i = 1; i < msz; ++i) {
            double w = h / diag;
            thw[i] = w;
            diag = 4.0 * h - w * h;
            thd[i] = 1.0 / diag;
        }
        if ((rc = upload(m->allocs, thw.data(), thw.size(), &p))) return fail(rc);
        dm.th_w = p;
        if ((rc = upload(m->allocs, thd.data(), thd.size(), &p))) return fail(rc);
        dm.th_inv_diag = p;
        {   // shared-transform basis tables (filled on first use, kept while the key parameters are unchanged), the key
            // they were built for and the two ints the device-side decision is published in
            cudaError_t e = cudaMalloc((void **)&m->d_tabs_cache, (size_t)2 * dm.nr * 9 * sizeof(double2));
            if (e == cudaSuccess) { m->allocs.push_back(m->d_tabs_cache); e = cudaMalloc((void **)&m->d_keystate, 32 * sizeof(double)); }
            if (e == cudaSuccess) { m->allocs.push_back(m->d_keystate); e = cudaMalloc((void **)&m->d_gate, 4 * sizeof(int)); }
            if (e == cudaSuccess) { m->allocs.push_back(m->d_gate); e = cudaMemset(m->d_keystate, 0, 32 * sizeof(double)); }
            if (e == cudaSuccess) e = cudaMemset(m->d_gate, 0, 4 * sizeof(int));
            if (e != cudaSuccess) { bf::set_error(std::string("bf_model_create: cudaMalloc: ") + cudaGetErrorString(e)); return fail(BF_ERR_ANALYSIS); }
        }
    }
    // metals
    dm.metal_kind = d->metal_kind;
    dm.idx_metal = d->idx_metal;
    if (d->metal_kind == BF_METAL_TOY) {
        if (!chk(d->idx_metal, 5)) { bf::set_error("MetalCorrelationModel: bad toy parameter block"); return fail(BF_ERR_MODEL); }
        if (cross) { bf::set_error("MetalCorrelationModel: illegal option for cross-correlation."); return fail(BF_ERR_MODEL); }
    } else if (d->metal_kind == BF_METAL_AUTO_TABLE || d->metal_kind == BF_METAL_CROSS_BICUBIC) {
        const bool mcross = d->metal_kind == BF_METAL_CROSS_BICUBIC;
        // baofit/MetalCorrelationModel.cc:35-37
        if (mcross != cross) { bf::set_error("MetalCorrelationModel: illegal option for cross-correlation."); return fail(BF_ERR_MODEL); }
        if (d->metal_ncomb < 1 || d->metal_ncomb > 20 || d->metal_nmet < 2 || d->metal_nmet > 6 || !d->metal_p1 || !d->metal_p2 ||
            !d->metal_zgrid || d->metal_ngrid < 1 || !chk(d->idx_metal, 2 * (d->metal_nmet - 1))) {
            bf::set_error("MetalCorrelationModel: bad metal description");
            return fail(BF_ERR_MODEL);
        }
        dm.metal_ncomb = d->metal_ncomb; dm.metal_nmet = d->metal_nmet; dm.metal_ngrid = d->metal_ngrid;
        for (int c = 0; c < d->metal_ncomb; ++c) {
            dm.metal_p1[c] = d->metal_p1[c];
            dm.metal_p2[c] = d->metal_p2[c];
            if (d->metal_p1[c] < 0 || d->metal_p1[c] >= d->metal_nmet || d->metal_p2[c] < 0 || d->metal_p2[c] >= d->metal_nmet) {
                bf::set_error("MetalCorrelationModel: bad combination index");
                return fail(BF_ERR_MODEL);
            }
        }
        m->metal_zgrid.assign(d->metal_zgrid, d->metal_zgrid + (size_t)d->metal_ncomb * d->metal_ngrid);
        double *p = nullptr;
        if (mcross) {
            if (d->metal_nx < 2 || d->metal_ny < 2 || !d->metal_cross || !(d->metal_dx > 0) || !(d->metal_dy > 0)) {
                bf::set_error("MetalCorrelationModel: error while reading metal model interpolation data.");
                return fail(BF_ERR_MODEL);
            }
            if (d->metal_ncomb > 5) { bf::set_error("MetalCorrelationModel: at most 5 cross combinations"); return fail(BF_ERR_MODEL); }
            int nt = d->metal_ncomb * 3, nx = d->metal_nx, ny = d->metal_ny;
            std::vector<double> il((size_t)nt * nx * ny);
            for (int t = 0; t < nt; ++t)
                for (int iy = 0; iy < ny; ++iy)
                    for (int ix = 0; ix < nx; ++ix)
                        il[((size_t)iy * nx + ix) * nt + t] = d->metal_cross[((size_t)t * ny + iy) * nx + ix];
            if ((rc = upload(m->allocs, il.data(), il.size(), &p))) return fail(rc);
            dm.metal_cross = p;
            m->metal_cross_h.assign(d->metal_cross, d->metal_cross + (size_t)nt * nx * ny);
            m->metal_dx = d->metal_dx;
            m->metal_dy = d->metal_dy;
            dm.metal_nx = nx; dm.metal_ny = ny; dm.metal_x0 = d->metal_x0; dm.metal_y0 = d->metal_y0;
            dm.metal_inv_dx = 1.0 / d->metal_dx; dm.metal_inv_dy = 1.0 / d->metal_dy;
            dm.metal_xmax = d->metal_x0 + (nx - 1) * d->metal_dx;
            dm.metal_ymax = d->metal_y0 + (ny - 1) * d->metal_dy;
        } else {
            if (!d->metal_auto) { bf::set_error("MetalCorrelationModel: missing templates"); return fail(BF_ERR_MODEL); }
            if ((rc = upload(m->allocs, d->metal_auto, (size_t)d->metal_ncomb * 3 * d->metal_ngrid, &p))) return fail(rc);
            dm.metal_auto = p;
        }
    } else if (d->metal_kind != BF_METAL_NONE) {
        bf::set_error("MetalCorrelationModel: illegal option specification.");
        return fail(BF_ERR_MODEL);
    }
    // distortion matrix (k-space model only, baofit/BaoKSpaceCorrelationModel.cc:188-200)
    if (d->dist_matrix) {
        if (d->kind != BF_MODEL_KSPACE) { bf::set_error("bf_model_create: the distortion matrix needs the k-space model"); return fail(BF_ERR_MODEL); }
        if (d->dist_matrix_order <= 0) { bf::set_error("DistortionMatrix: expected distortion matrix order > 0."); return fail(BF_ERR_MODEL); }
        dm.dm_order = d->dist_matrix_order;
        m->dist_matrix.assign(d->dist_matrix, d->dist_matrix + (size_t)d->dist_matrix_order * d->dist_matrix_order);
    }
    *out = m;
    return BF_OK;
}

extern "C" int bf_model_transform_nr(const bf_model *m) { return m ? m->dev.nr : 0; }
extern "C" double bf_model_transform_r0(const bf_model *m) { return m ? m->dev.tr_r0 : 0; }
extern "C" double bf_model_transform_dr(const bf_model *m) { return m ? m->dev.tr_dr : 0; }

// ------------------------------------------------------------------------------------------
// data
// ------------------------------------------------------------------------------------------
extern "C" int bf_data_destroy(bf_data *d) {
    if (!d) return BF_OK;
    std::lock_guard<std::recursive_mutex> lock(d->ctx->api_mu);
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    {
        auto &live = d->ctx->live_data;
        live.erase(std::remove(live.begin(), live.end(), d), live.end());
    }
    purge_graphs(d->ctx, 0, d);
    if (d->expanded) bf_data_destroy(d->expanded);
    for (void *p : d->allocs) cudaFree(p);
    for (auto &kv : d->plans)
        for (void *p : kv.second.allocs) cudaFree(p);
    delete d;
    return BF_OK;
}

extern "C" int bf_data_ndata(const bf_data *d) { return d ? d->n : 0; }

template <class T> static int upload_padded(std::vector<void *> &owner, const T *host, int n, int n_pad, T **dev) {
    std::vector<T> tmp(n_pad, T(0));
    if (host) std::copy(host, host + n, tmp.begin());
    return upload(owner, tmp.data(), (size_t)n_pad, dev);
}

extern "C" int bf_data_create(bf_ctx *ctx, const bf_data_desc *d, bf_data **out) {
    if (!ctx || !d || !out) BF_FAIL(BF_ERR_OPTIONS, "bf_data_create: null argument");
    if (d->n_data <= 0) BF_FAIL(BF_ERR_DATA, "CorrelationFitter: need some data to fit.");
    if (!d->r || !d->mu || !d->z || !d->index || !d->data || !d->cov) BF_FAIL(BF_ERR_DATA, "bf_data_create: missing arrays");
    if (d->n_grid < 0 || (d->n_grid > 0 && (!d->grid_r || !d->grid_mu || !d->grid_z)))
        BF_FAIL(BF_ERR_DATA, "AbsCorrelationModel::setCoordinates: coordinate vectors not the same size.");
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    const int n = d->n_data, n_pad = round_up(n, kPadM);
    // factor the covariance once (host): W lower triangular with W^T W = C^-1
    std::vector<double> a((size_t)n * n), W((size_t)n * n), L;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) a[(size_t)i * n + j] = 0.5 * (d->cov[(size_t)i * n + j] + d->cov[(size_t)j * n + i]);
    if (!d->cov_is_inverse) {
        if (!cholesky_lower(n, a.data())) BF_FAIL(BF_ERR_DATA, "bf_data_create: covariance is not positive definite");
        L = a;
        invert_lower(n, L.data(), W.data());
    } else {
        // C^-1 = U U^T with U upper triangular, from the Cholesky factor of the index-reversed
        // matrix; W = U^T
        std::vector<double> f((size_t)n * n);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) f[(size_t)i * n + j] = a[(size_t)(n - 1 - i) * n + (n - 1 - j)];
        if (!cholesky_lower(n, f.data())) BF_FAIL(BF_ERR_DATA, "bf_data_create: inverse covariance is not positive definite");
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) W[(size_t)i * n + j] = (i >= j) ? f[(size_t)(n - 1 - j) * n + (n - 1 - i)] : 0.0;
    }
    bf_data *o = new bf_data();
    o->ctx = ctx;
    ctx->live_data.push_back(o);
    o->n = n;
    o->n_pad = n_pad;
    o->n_grid = d->n_grid;
    o->n_grid_pad = round_up(std::max(d->n_grid, 1), kPadM);
    o->h_r.assign(d->r, d->r + n);
    o->h_mu.assign(d->mu, d->mu + n);
    o->h_z.assign(d->z, d->z + n);
    o->h_index.assign(d->index, d->index + n);
    if (d->n_grid > 0) {
        o->h_grid_r.assign(d->grid_r, d->grid_r + d->n_grid);
        o->h_grid_mu.assign(d->grid_mu, d->grid_mu + d->n_grid);
        o->h_grid_z.assign(d->grid_z, d->grid_z + d->n_grid);
    }
    int rc;
    auto fail = [&](int code) { bf_data_destroy(o); return code; };
    if ((rc = upload_padded(o->allocs, d->r, n, n_pad, &o->d_r))) return fail(rc);
    if ((rc = upload_padded(o->allocs, d->mu, n, n_pad, &o->d_mu))) return fail(rc);
    if ((rc = upload_padded(o->allocs, d->z, n, n_pad, &o->d_z))) return fail(rc);
    if ((rc = upload_padded(o->allocs, d->index, n, n_pad, &o->d_index))) return fail(rc);
    if ((rc = upload_padded(o->allocs, d->data, n, n_pad, &o->d_data))) return fail(rc);
    if (d->n_grid > 0) {
        if ((rc = upload_padded(o->allocs, d->grid_r, d->n_grid, o->n_grid_pad, &o->d_grid_r))) return fail(rc);
        if ((rc = upload_padded(o->allocs, d->grid_mu, d->n_grid, o->n_grid_pad, &o->d_grid_mu))) return fail(rc);
        if ((rc = upload_padded(o->allocs, d->grid_z, d->n_grid, o->n_grid_pad, &o->d_grid_z))) return fail(rc);
    }
    auto pad_square = [&](const std::vector<double> &src, double **dev) {
        std::vector<double> p((size_t)n_pad * n_pad, 0.0);
        for (int i = 0; i < n; ++i) std::memcpy(&p[(size_t)i * n_pad], &src[(size_t)i * n], sizeof(double) * n);
        return upload(o->allocs, p.data(), p.size(), dev);
    };
    o->h_W = W;
    o->h_data.assign(d->data, d->data + n);
    if ((rc = pad_square(W, &o->d_W))) return fail(rc);
    if (!L.empty() && (rc = pad_square(L, &o->d_L))) return fail(rc);
    if (d->binning_type == BF_BINNING_MULTIPOLE) {
        // CorrelationFitter.cc:66-69: every bin becomes nmu points (r_i, mu_q, z_i) of a Gauss-Legendre rule
        // on [-1,1]; the prediction is sum_q (2 ell_i + 1)/2 w_q L_ell_i(mu_q) xi(r_i, mu_q, z_i)
        if (d->n_grid > 0) return fail((bf::set_error("bf_data_create: multipole data cannot carry a distortion-matrix grid"), BF_ERR_DATA));
        const int nmu = d->nmu_project > 0 ? d->nmu_project : 16;
        if (nmu > 64) return fail((bf::set_error("bf_data_create: nmu_project <= 64"), BF_ERR_DATA));
        std::vector<double> x(nmu), w(nmu), proj((size_t)n * nmu), er((size_t)n * nmu), emu((size_t)n * nmu), ez((size_t)n * nmu);
        std::vector<int> eidx((size_t)n * nmu);
        gauss_legendre_01(nmu, x.data(), w.data());
        auto legendre = [](int ell, double mu) {
            double m2 = mu * mu;
            return ell == 0 ? 1.0 : ell == 2 ? (-1 + 3 * m2) / 2. : (3 + m2 * (-30 + 35 * m2)) / 8.;
        };
        for (int i = 0; i < n; ++i) {
            const int ell = (int)std::floor(d->mu[i] + 0.5);  // ComovingCorrelationData.cc:73-75
            if (ell != 0 && ell != 2 && ell != 4) return fail((bf::set_error("bf_data_create: multipole bins must have ell = 0, 2 or 4"), BF_ERR_DATA));
            for (int q = 0; q < nmu; ++q) {
                const double mu = 2 * x[q] - 1;
                const size_t k = (size_t)i * nmu + q;
                er[k] = d->r[i]; emu[k] = mu; ez[k] = d->z[i]; eidx[k] = d->index[i];
                proj[k] = 0.5 * (2 * ell + 1) * 2 * w[q] * legendre(ell, mu);
            }
        }
        bf_data *e = new bf_data();
        e->ctx = ctx;
        ctx->live_data.push_back(e);
        e->points_only = true;
        e->trigger_mu_zero = true;
        e->n = n * nmu;
        e->n_pad = round_up(e->n, kPadM);
        e->n_grid = 0;
        e->n_grid_pad = kPadM;
        e->h_r = er; e->h_mu = emu; e->h_z = ez; e->h_index = eidx;
        o->expanded = e;
        o->binning_type = BF_BINNING_MULTIPOLE;
        o->nmu = nmu;
        if ((rc = upload_padded(e->allocs, er.data(), e->n, e->n_pad, &e->d_r))) return fail(rc);
        if ((rc = upload_padded(e->allocs, emu.data(), e->n, e->n_pad, &e->d_mu))) return fail(rc);
        if ((rc = upload_padded(e->allocs, ez.data(), e->n, e->n_pad, &e->d_z))) return fail(rc);
        if ((rc = upload_padded(e->allocs, eidx.data(), e->n, e->n_pad, &e->d_index))) return fail(rc);
        if ((rc = upload(o->allocs, proj.data(), proj.size(), &o->d_proj))) return fail(rc);
    }
    *out = o;
    return BF_OK;
}

// ------------------------------------------------------------------------------------------
// plan = (model, data) binding: redshift classes, kept rows of the distortion matrix
// ------------------------------------------------------------------------------------------
static int get_plan(const bf_model *m, const bf_data *dc, const bf_data::Plan **out) {
    bf_data *d = const_cast<bf_data *>(dc);  // the plan cache is the one mutable part of a data object (guarded by the context lock)
    auto it = d->plans.find(m->serial);
    if (it != d->plans.end()) { *out = &it->second; return BF_OK; }
    const DevModel &dm = m->dev;
    bf_data::Plan p;
    struct FreeOnFailure {  // every early return below leaves through here
        bf_data::Plan &p;
        bool keep = false;
        ~FreeOnFailure() { if (!keep) for (void *q : p.allocs) cudaFree(q); }
    } guard{p};
    std::map<double, int> cls;
    auto classify = [&](double z) {
        auto f = cls.find(z);
        if (f != cls.end()) return f->second;
        int id = (int)p.zvals.size();
        cls[z] = id;
        p.zvals.push_back(z);
        return id;
    };
    const int n = d->n, n_pad = d->n_pad, ng = d->n_grid, ng_pad = d->n_grid_pad;
    std::vector<int> zc_data(n_pad, 0), zc_grid(ng_pad, 0);
    for (int i = 0; i < n; ++i) zc_data[i] = classify(d->h_z[i]);
    p.zc_trigger = zc_data[0];  // the first kept bin triggers the transform
    const bool use_dm = dm.dm_order > 0;
    if (use_dm) {
        // baofit/src/baofit.cc:591-594 order check
        if (ng != dm.dm_order) BF_FAIL(BF_ERR_DATA, "distortion matrix order does not match the number of grid bins");
        for (int g = 0; g < ng; ++g) zc_grid[g] = classify(d->h_grid_z[g]);
        for (int i = 0; i < n; ++i)
            if (d->h_index[i] < 0 || d->h_index[i] >= ng) BF_FAIL(BF_ERR_DATA, "DistortionMatrix::getDistortion: invalid indices.");
    }
    std::vector<int> zcm_data, zcm_grid;
    if (dm.metal_kind == BF_METAL_AUTO_TABLE || dm.metal_kind == BF_METAL_CROSS_BICUBIC) {
        zcm_data.assign((size_t)dm.metal_ncomb * n_pad, 0);
        for (int c = 0; c < dm.metal_ncomb; ++c)
            for (int i = 0; i < n; ++i) {
                int idx = d->h_index[i];
                if (idx < 0 || idx >= dm.metal_ngrid) BF_FAIL(BF_ERR_DATA, "MetalCorrelationModel::_evaluate: invalid index.");
                zcm_data[(size_t)c * n_pad + i] = classify(m->metal_zgrid[(size_t)c * dm.metal_ngrid + idx]);
            }
        if (use_dm) {
            if (dm.metal_ngrid < ng) BF_FAIL(BF_ERR_DATA, "MetalCorrelationModel: grid shorter than the distortion matrix order");
            zcm_grid.assign((size_t)dm.metal_ncomb * ng_pad, 0);
            for (int c = 0; c < dm.metal_ncomb; ++c)
                for (int g = 0; g < ng; ++g) zcm_grid[(size_t)c * ng_pad + g] = classify(m->metal_zgrid[(size_t)c * dm.metal_ngrid + g]);
        }
    }
    p.nz = (int)p.zvals.size();
    // The reference re-runs its k-space transforms in the middle of a sweep over the bins when the redshift of two
    // consecutive bins differs by more than dzmin (BaoKSpaceCorrelationModel.cc:317-319,333), with beta(z) -- and z_eff
    // when none is set -- of the bin it happens to be at: a result that depends on the bin order. Here the transform
    // uses the redshift of the first kept bin. The two agree unless such a jump exists AND something in D(k, mu) depends
    // on z: gamma-beta != 0 (flagged per vector, BF_STATUS_Z_RETRIGGER) or a non-linear correction without z_eff (refused).
    p.z_retrigger = false;
    if (dm.kind == BF_MODEL_KSPACE) {
        const std::vector<double> &zs = use_dm ? d->h_grid_z : d->h_z;
        const int nzs = use_dm ? ng : n;
        for (int i = 1; i < nzs; ++i)
            if (std::fabs(zs[i] - zs[i - 1]) > m->dzmin) p.z_retrigger = true;
        if (p.z_retrigger && dm.zeff <= 0 && (dm.flags & (BF_FLAG_NL_CORRECTION | BF_FLAG_NL_CORRECTION_ALT | BF_FLAG_FIT_NL_CORRECTION)))
            BF_FAIL(BF_ERR_OPTIONS, "k-space model with a non-linear correction and no zeff on data whose consecutive bins jump by more than "
                                    "dzmin in redshift: the reference re-transforms per bin here (BaoKSpaceCorrelationModel.cc:315-333), "
                                    "which is not reproduced; set zeff");
    }
    std::vector<double> vfac(p.nz);
    for (int c = 0; c < p.nz; ++c) {
        double zp1 = 1.0 + p.zvals[c];
        vfac[c] = (1.0 / 100.) * zp1 / std::sqrt(1.0 - dm.omega_matter + dm.omega_matter * zp1 * zp1 * zp1);
    }
    BF_CUDA_OK(cudaSetDevice(d->ctx->device));
    int rc;
    if ((rc = upload(p.allocs, p.zvals.data(), p.zvals.size(), &p.d_zvals))) return rc;
    if ((rc = upload(p.allocs, vfac.data(), vfac.size(), &p.d_vfac))) return rc;
    if ((rc = upload(p.allocs, zc_data.data(), zc_data.size(), &p.d_zc_data))) return rc;
    if ((rc = upload(p.allocs, zc_grid.data(), zc_grid.size(), &p.d_zc_grid))) return rc;
    if (!zcm_data.empty() && (rc = upload(p.allocs, zcm_data.data(), zcm_data.size(), &p.d_zc_metal_data))) return rc;
    if (!zcm_grid.empty() && (rc = upload(p.allocs, zcm_grid.data(), zcm_grid.size(), &p.d_zc_metal_grid))) return rc;
    if (use_dm) {
        std::vector<double> dk((size_t)n_pad * ng_pad, 0.0);
        for (int i = 0; i < n; ++i)
            std::memcpy(&dk[(size_t)i * ng_pad], &m->dist_matrix[(size_t)d->h_index[i] * ng], sizeof(double) * ng);
        if ((rc = upload(p.allocs, dk.data(), dk.size(), &p.d_Dkept))) return rc;
    }
    // bin-static tables of the cartesian-form kernels
    auto build_tab = [&](const std::vector<double> &r, const std::vector<double> &mu, const std::vector<double> &z,
                         int npts, int npts_pad, const int *slots, int nslots, bool want_mx,
                         bf_data::Plan::PointTabDev &t) -> int {
        std::vector<double> rpar(npts_pad, 0.0), rperp(npts_pad, 0.0);
        for (int i = 0; i < npts; ++i) {
            rpar[i] = r[i] * mu[i];
            rperp[i] = r[i] * std::sqrt(1 - mu[i] * mu[i]);
        }
        int rc2;
        if ((rc2 = upload(p.allocs, rpar.data(), rpar.size(), &t.rpar0))) return rc2;
        if ((rc2 = upload(p.allocs, rperp.data(), rperp.size(), &t.rperp0))) return rc2;
        for (int q = 0; q < nslots; ++q) {
            int s = slots[q];
            const DevBB &b = dm.bb[s];
            if (!b.enabled) continue;
            int n_rt = (b.rt_max - b.rt_min) / b.rt_step + 1, n_z = (b.z_max - b.z_min) / b.z_step + 1;
            std::vector<double> rt((size_t)npts * n_rt), zp((size_t)npts * n_z);
            for (int i = 0; i < npts; ++i) {
                // BroadbandModel.cc:201-213: rrT = r*sqrt(1-mu^2)/r0, zz = (1+z)/(1+z0), "x-1 if exponent > 0"
                double rrT = r[i] * std::sqrt(1 - mu[i] * mu[i]) * b.inv_r0, zz = (1 + z[i]) / (1 + b.z0);
                int k = 0;
                for (int ti = b.rt_min; ti <= b.rt_max; ti += b.rt_step) rt[(size_t)i * n_rt + k++] = std::pow(ti > 0 ? rrT - 1 : rrT, ti);
                k = 0;
                for (int zi = b.z_min; zi <= b.z_max; zi += b.z_step) zp[(size_t)i * n_z + k++] = std::pow(zz, zi);
            }
            t.n_rt[s] = n_rt;
            t.n_z[s] = n_z;
            if ((rc2 = upload(p.allocs, rt.data(), rt.size(), &t.bb_rt[s]))) return rc2;
            if ((rc2 = upload(p.allocs, zp.data(), zp.size(), &t.bb_z[s]))) return rc2;
        }
        if (want_mx && dm.metal_kind == BF_METAL_CROSS_BICUBIC) {
            // collapse the bicubic patch along r_perp, which the velocity shift leaves unchanged
            const int nt = dm.metal_ncomb * 3, nx = dm.metal_nx, ny = dm.metal_ny;
            std::vector<double> mx((size_t)npts * ny * nt);
            auto wrap = [](int i, int n) { i %= n; return i < 0 ? i + n : i; };
            for (int i = 0; i < npts; ++i) {
                double x = std::min(std::max(rperp[i], dm.metal_x0), dm.metal_xmax);
                double fx = (x - dm.metal_x0) * dm.metal_inv_dx, flx = std::floor(fx), tt = fx - flx;
                int ix = (int)flx;
                double t2 = tt * tt, t3 = t2 * tt;
                double wx[4] = {0.5 * (-tt + 2.0 * t2 - t3), 0.5 * (2.0 - 5.0 * t2 + 3.0 * t3), 0.5 * (tt + 4.0 * t2 - 3.0 * t3),
                                0.5 * (-t2 + t3)};
                for (int row = 0; row < ny; ++row)
                    for (int tpl = 0; tpl < nt; ++tpl) {
                        const double *g = &m->metal_cross_h[((size_t)tpl * ny + row) * nx];
                        double v = 0;
                        for (int jx = 0; jx < 4; ++jx) v += wx[jx] * g[wrap(ix - 1 + jx, nx)];
                        mx[((size_t)i * ny + row) * nt + tpl] = v;
                    }
            }
            if ((rc2 = upload(p.allocs, mx.data(), mx.size(), &t.mx))) return rc2;
            // the same rows with the periodic wrap of the Catmull-Rom stencil materialised (halo row -1 in front, rows
            // ny and ny+1 behind): the fused kernel stages windows of consecutive rows and never wraps an index
            const int nyh = ny + 3;
            std::vector<double> mxh((size_t)npts * nyh * nt);
            for (int i = 0; i < npts; ++i)
                for (int row = 0; row < nyh; ++row)
                    std::memcpy(&mxh[((size_t)i * nyh + row) * nt], &mx[((size_t)i * ny + wrap(row - 1, ny)) * nt], sizeof(double) * nt);
            if ((rc2 = upload(p.allocs, mxh.data(), mxh.size(), &t.mxh))) return rc2;
        }
        return BF_OK;
    };
    {
        const int post_slots[2] = {BF_BB_DIST_ADD, BF_BB_DIST_MUL}, pre_slots[2] = {BF_BB_DM_DIST_ADD, BF_BB_DM_DIST_MUL};
        if ((rc = build_tab(d->h_r, d->h_mu, d->h_z, n, n_pad, post_slots, 2, !use_dm, p.tab_data))) return rc;
        if (use_dm && (rc = build_tab(d->h_grid_r, d->h_grid_mu, d->h_grid_z, ng, ng_pad, pre_slots, 2, true, p.tab_grid))) return rc;
    }
    // ---- fused tail ----
    // Distortion-matrix path: the post-matrix broadband and "- d" become extra K columns of [D | basis | -d] and W is
    // folded in. Without a distortion matrix (data_mode) the same trick removes the second model pass of lean
    // models: A = [W | W.basis | -W.d], the panel is [xi_cosmology ; coefficient rows ; 1].
    auto build_fold = [&](bool data_mode) -> int {
        const DevBB &ba = dm.bb[BF_BB_DIST_ADD];
        bool single_class = true;
        for (int i = 1; i < n; ++i) single_class &= zc_data[i] == zc_data[0];
        const bool shift = dm.idx_delta_v >= 0;
        int fcase = -1;
        if (!dm.bb[BF_BB_DIST_MUL].enabled && single_class) {
            if (!ba.enabled) fcase = 0;
            else if (!shift) fcase = 1;
            else if (ba.r_min == 0 && ba.r_max == 0 && ba.mu_min == 0 && ba.mu_max == 0 && ba.rp_min >= 0 && ba.rp_max <= 8) fcase = 2;
        }
        // several redshift classes among the data bins (e.g. the three z bins of BOSSDR9LyaF): the evolution factor
        // of the broadband differs per class, so every class gets its own block of basis columns / coefficient rows
        int ncls = 1;
        if (fcase < 0 && data_mode && !dm.bb[BF_BB_DIST_MUL].enabled && !single_class && !shift) {
            fcase = ba.enabled ? 1 : 0;
            ncls = ba.enabled ? p.nz : 1;
        }
        if (fcase < 0) return BF_OK;
        auto cnt = [](int lo, int hi, int st) { return (hi - lo) / st + 1; };
        // (the coefficient rows must fit the 64 spare rows below the panel; checked below)
        int n_rt = ba.enabled ? cnt(ba.rt_min, ba.rt_max, ba.rt_step) : 0, n_z = ba.enabled ? cnt(ba.z_min, ba.z_max, ba.z_step) : 0;
        int nb = 0;
        if (fcase == 1)
            nb = n_z * cnt(ba.mu_min, ba.mu_max, ba.mu_step) * cnt(ba.r_min, ba.r_max, ba.r_step) * cnt(ba.rp_min, ba.rp_max, ba.rp_step) * n_rt;
        else if (fcase == 2) nb = n_z * (ba.rp_max + 1) * n_rt;
        const int nb1 = nb;  // coefficients of one class
        nb *= ncls;
        const int kext = round_up(nb + 1, kGemmBK), kleft = data_mode ? n_pad : ng_pad, ktot = kleft + kext;
        if (kext > 64) return BF_OK;
        auto legendre = [](int ell, double mu) {
            double m2 = mu * mu;
            switch (ell) {
            case 0: return 1.0;
            case 1: return mu;
            case 2: return (-1 + 3 * m2) / 2.;
            case 3: return mu * (-3 + 5 * m2) / 2.;
            case 4: return (3 + m2 * (-30 + 35 * m2)) / 8.;
            case 5: return mu * (15 + m2 * (-70 + 63 * m2)) / 8.;
            case 6: return (-5 + m2 * (105 + m2 * (-315 + 231 * m2))) / 16.;
            case 7: return mu * (-35 + m2 * (315 + m2 * (-693 + 429 * m2))) / 16.;
            default: return (35 + m2 * (-1260 + m2 * (6930 + m2 * (-12012 + 6435 * m2)))) / 128.;
            }
        };
        // tail columns of data bin i: the broadband basis values, then -d_i
        auto tail_row = [&](int i, double *bs) {
            if (ba.enabled) {
                const double r = d->h_r[i], mu = d->h_mu[i], z = d->h_z[i];
                const double rr = r * ba.inv_r0, rrP = r * mu * ba.inv_r0, rrT = r * std::sqrt(1 - mu * mu) * ba.inv_r0;
                const double zz = (1 + z) / (1 + ba.z0);
                int k = ncls > 1 ? zc_data[i] * nb1 : 0;
                if (fcase == 1) {  // BroadbandModel.cc:204-213, one basis value per coefficient
                    for (int zi = ba.z_min; zi <= ba.z_max; zi += ba.z_step)
                        for (int mi = ba.mu_min; mi <= ba.mu_max; mi += ba.mu_step)
                            for (int ri = ba.r_min; ri <= ba.r_max; ri += ba.r_step)
                                for (int pi = ba.rp_min; pi <= ba.rp_max; pi += ba.rp_step)
                                    for (int ti = ba.rt_min; ti <= ba.rt_max; ti += ba.rt_step)
                                        bs[k++] = std::pow(zz, zi) * legendre(mi, mu) *
                                                  std::pow(ri > 0 ? rr - 1 : rr, (double)ri / ba.r_denom) *
                                                  std::pow(pi > 0 ? rrP - 1 : rrP, pi) * std::pow(ti > 0 ? rrT - 1 : rrT, ti);
                } else {
                    const double u = rrP - 1;
                    for (int zi = ba.z_min; zi <= ba.z_max; zi += ba.z_step)
                        for (int mm = 0; mm <= ba.rp_max; ++mm)
                            for (int ti = ba.rt_min; ti <= ba.rt_max; ti += ba.rt_step)
                                bs[k++] = std::pow(zz, zi) * std::pow(u, mm) * std::pow(ti > 0 ? rrT - 1 : rrT, ti);
                }
            }
            bs[nb] = -d->h_data[i];
        };
        std::vector<double> wa((size_t)n_pad * ktot, 0.0);
        const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        if (!data_mode) {
            std::vector<double> aext((size_t)n_pad * ktot, 0.0);
            for (int i = 0; i < n; ++i) {
                double *row = &aext[(size_t)i * ktot];
                std::memcpy(row, &m->dist_matrix[(size_t)d->h_index[i] * ng], sizeof(double) * ng);
                tail_row(i, row + ng_pad);
            }
            // WA = W . A_ext (W lower triangular), accumulated in long double
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < nt; ++t)
                pool.emplace_back([&, t] {
                    std::vector<long double> acc(ktot);
                    for (int i = (int)t; i < n; i += (int)nt) {
                        std::fill(acc.begin(), acc.end(), 0.0L);
                        for (int j = 0; j <= i; ++j) {
                            const long double wij = d->h_W[(size_t)i * n + j];
                            if (wij == 0) continue;
                            const double *arow = &aext[(size_t)j * ktot];
                            for (int k = 0; k < ktot; ++k) acc[k] += wij * arow[k];
                        }
                        for (int k = 0; k < ktot; ++k) wa[(size_t)i * ktot + k] = (double)acc[k];
                    }
                });
            for (auto &th : pool) th.join();
            if ((rc = upload(p.allocs, aext.data(), aext.size(), &p.d_Aext))) return rc;
        } else {
            // [W | W.tail]: the left block is W itself, only the kext tail columns need a product
            std::vector<double> tails((size_t)n * kext, 0.0);
            for (int i = 0; i < n; ++i) tail_row(i, &tails[(size_t)i * kext]);
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < nt; ++t)
                pool.emplace_back([&, t] {
                    std::vector<long double> acc(kext);
                    for (int i = (int)t; i < n; i += (int)nt) {
                        double *row = &wa[(size_t)i * ktot];
                        std::fill(acc.begin(), acc.end(), 0.0L);
                        for (int j = 0; j <= i; ++j) {
                            const double wij = d->h_W[(size_t)i * n + j];
                            row[j] = wij;
                            if (wij == 0) continue;
                            const double *trow = &tails[(size_t)j * kext];
                            for (int k = 0; k < kext; ++k) acc[k] += (long double)wij * trow[k];
                        }
                        for (int k = 0; k < kext; ++k) row[n_pad + k] = (double)acc[k];
                    }
                });
            for (auto &th : pool) th.join();
        }
        if ((rc = upload(p.allocs, wa.data(), wa.size(), &p.d_WA))) return rc;
        if (!data_mode && n_pad <= kFullmMaxM && n_pad % 32 == 0 && ktot % kFullmKC == 0) {
            std::vector<double> wat((size_t)(ktot / kFullmKC) * n_pad * kFullmLdA);
            gemm_fullm_tile_A(wa.data(), n_pad, ktot, ktot, wat.data());
            if ((rc = upload(p.allocs, wat.data(), wat.size(), &p.d_WAt))) return rc;
        }
        if (!data_mode && ktot % 8 == 0) {
            std::vector<double> waf((size_t)n_pad * ktot);
            gemm_fused_tile_A(wa.data(), n_pad, ktot, ktot, waf.data());
            if ((rc = upload(p.allocs, waf.data(), waf.size(), &p.d_WAf))) return rc;
            if (ktot % 16 == 0 && n_pad == 448) {
                // chi2 = || WA x ||^2 = || R x ||^2 with WA = Q R (Householder, Q orthonormal): R is upper trapezoidal, row i
                // starts at column i, so the fused kernel can skip every DMMA fragment whose rows have not started yet at
                // the K chunk it is working on -- 43 % of the tensor work of a 440 x 528 matrix. The row order of R is free
                // (a permutation does not change the norm): fragment f (rows 8f .. 8f+7) goes to consumer warp f % 4, half
                // (f / 4) % 2, slot f / 8, so that all warps and both halves fill up at the same pace.
                static const bool no_tri = std::getenv("BAOFIT_FUSED_DENSE") != nullptr;  // A/B measurements
                if (!no_tri && n <= ktot) {
                    std::vector<double> R(wa.begin(), wa.begin() + (size_t)n * ktot), wrow(ktot), v(n);
                    for (int j = 0; j < n; ++j) {
                        double nrm = 0;
                        for (int i = j; i < n; ++i) nrm += R[(size_t)i * ktot + j] * R[(size_t)i * ktot + j];
                        nrm = std::sqrt(nrm);
                        if (nrm == 0) continue;
                        const double alpha = R[(size_t)j * ktot + j] > 0 ? -nrm : nrm;
                        double vn2 = 0;
                        for (int i = j; i < n; ++i) { v[i] = R[(size_t)i * ktot + j] - (i == j ? alpha : 0.0); vn2 += v[i] * v[i]; }
                        if (vn2 == 0) continue;
                        std::fill(wrow.begin() + j, wrow.end(), 0.0);
                        for (int i = j; i < n; ++i) {
                            const double vi = v[i];
                            const double *row = &R[(size_t)i * ktot];
                            for (int cidx = j; cidx < ktot; ++cidx) wrow[cidx] += vi * row[cidx];
                        }
                        const double s2 = 2.0 / vn2;
                        for (int i = j; i < n; ++i) {
                            const double f = s2 * v[i];
                            double *row = &R[(size_t)i * ktot];
                            for (int cidx = j; cidx < ktot; ++cidx) row[cidx] -= f * wrow[cidx];
                        }
                        R[(size_t)j * ktot + j] = alpha;
                        for (int i = j + 1; i < n; ++i) R[(size_t)i * ktot + j] = 0.0;
                    }
                    std::vector<double> perm((size_t)n_pad * ktot, 0.0);
                    for (int f = 0; f < n_pad / 8; ++f) {
                        const int w_ = f % 4, h_ = (f / 4) % 2, mi = f / 8;
                        for (int jr = 0; jr < 8; ++jr) {
                            const int src = 8 * f + jr, dst = h_ * (n_pad / 2) + w_ * (n_pad / 8) + mi * 8 + jr;
                            if (src < n) std::memcpy(&perm[(size_t)dst * ktot], &R[(size_t)src * ktot], sizeof(double) * ktot);
                        }
                    }
                    gemm_fused_tile_A3(perm.data(), n_pad, ktot, ktot, waf.data());
                    p.fused_tri = true;
                } else
                    gemm_fused_tile_A3(wa.data(), n_pad, ktot, ktot, waf.data());
                if ((rc = upload(p.allocs, waf.data(), waf.size(), &p.d_WAf3))) return rc;
            }
        }
        p.fold_ok = true;
        p.fold_data_mode = data_mode;
        p.fold_case = fcase;
        p.fold_nb = nb;
        p.fold_ncls = ncls;
        p.fold_kext = kext;
        p.fold_pmax = ba.enabled ? ba.rp_max : 0;
        p.fold_nrt = n_rt;
        p.fold_nz = n_z;
        return BF_OK;
    };
    if (use_dm) {
        if ((rc = build_fold(false))) return rc;
    } else if ((dm.kind == BF_MODEL_KSPACE || (dm.kind == BF_MODEL_RSPACE && m->d_rs_tabs)) && dm.metal_kind == BF_METAL_NONE &&
               !(dm.flags & BF_FLAG_RADIATION_MODEL) && d->h_W.size() == (size_t)n * n) {
        if ((rc = build_fold(true))) return rc;
    }
    guard.keep = true;
    auto ins = d->plans.emplace(m->serial, std::move(p));
    *out = &ins.first->second;
    return BF_OK;
}

// ------------------------------------------------------------------------------------------
// pipelines
// ------------------------------------------------------------------------------------------
namespace {
struct Workspace {
    double *pT, *S, *G, *T, *C2, *XiU, *Y, *Z;
    int *status;
    // fixed-size region of the shared-transform path (independent of the batch size)
    double *Gb, *Tb, *Cb;
    double2 *tabs;
    double *flag;  // [kKeyMax] key-check answer
    int *sort_hist, *sort_base, *rank;  // coherence sort
    double *chi2_sorted;
    double *fft_c12, *fft_cont, *fft_planes;  // FFT model: [2][ldb], [my_pad][ldb], [ldb][2][npy_pad][npx_pad]
    double *fft_keys;                         // [kFftKeyMax-1][ldb]
};

size_t ws_fixed_doubles(const DevModel &dm) {
    size_t sort = kSortTableInts;  // two int tables of 2048 = 2048 doubles
    if (dm.kind != BF_MODEL_KSPACE) return 16 + sort;
    return (size_t)6 * dm.nk_pad * kPadN + (size_t)12 * dm.nr_pad * kPadN + (size_t)2 * dm.nr * 9 * 2 + 32 + sort;
}

size_t ws_doubles_per_column(const DevModel &dm, const bf_data *d, int nz) {
    size_t c = (size_t)dm.npar + (size_t)nz * 3 + 1 /*status*/ + 2 /*rank, sorted chi2*/ + d->n_pad + 64 /*Z + fused-tail rows*/;
    if (dm.kind == BF_MODEL_KSPACE) c += (size_t)6 * dm.nk_pad + (size_t)12 * dm.nr_pad;
    if (dm.dm_order > 0) c += (size_t)d->n_grid_pad + 64 + d->n_pad;  // xi_u panel (+ fused-tail rows) and Y
    if (dm.kind == BF_MODEL_KSPACE_FFT) c += 2 + kFftKeyMax + (size_t)dm.fft.my_pad + (size_t)2 * dm.fft.npy_pad * dm.fft.npx_pad;
    return c;
}

void carve(void *base, const DevModel &dm, const bf_data *d, int nz, int ldb, Workspace &w) {
    double *p = (double *)base;
    w.flag = p;
    p += 16;
    w.sort_hist = (int *)p;
    w.sort_base = w.sort_hist + kSortTableInts;
    p += kSortTableInts;
    w.Gb = w.Tb = w.Cb = nullptr;
    w.tabs = nullptr;
    if (dm.kind == BF_MODEL_KSPACE) {
        w.tabs = (double2 *)p;
        p += (size_t)2 * dm.nr * 9 * 2 + 16;
        w.Gb = p;
        p += (size_t)6 * dm.nk_pad * kPadN;
        w.Tb = p;
        p += (size_t)6 * dm.nr_pad * kPadN;
        w.Cb = p;
        p += (size_t)6 * dm.nr_pad * kPadN;
    }
    auto take = [&](size_t rows) { double *q = p; p += rows * (size_t)ldb; return q; };
    w.pT = take(dm.npar);
    w.S = take((size_t)nz * 3);
    w.status = (int *)take(1);
    w.rank = (int *)take(1);
    w.chi2_sorted = take(1);
    w.Z = take((size_t)d->n_pad + 64);
    w.G = w.T = w.C2 = w.XiU = w.Y = nullptr;
    if (dm.kind == BF_MODEL_KSPACE) {
        w.G = take((size_t)6 * dm.nk_pad);
        w.T = take((size_t)6 * dm.nr_pad);
        w.C2 = take((size_t)6 * dm.nr_pad);
    }
    if (dm.dm_order > 0) {
        w.XiU = take((size_t)d->n_grid_pad + 64);
        w.Y = take(d->n_pad);
    }
    w.fft_c12 = w.fft_cont = w.fft_planes = nullptr;
    if (dm.kind == BF_MODEL_KSPACE_FFT) {
        w.fft_c12 = take(2);
        w.fft_keys = take(kFftKeyMax);
        w.fft_cont = take(dm.fft.my_pad);
        w.fft_planes = take((size_t)2 * dm.fft.npy_pad * dm.fft.npx_pad);
    }
}

void nl_table(const DevModel &dm, double zeff, double out[5], double *pk_scale) {
    // baofit/NonLinearCorrectionModel.cc:22-27, 55-64: linear interpolation in z, end-point clamp
    static const double zA[] = {2.2, 2.4, 2.6, 2.8, 3.0};
    static const double tab[5][5] = {{0.8670, 0.8510, 0.7810, 0.7730, 0.7920},
                                     {1.1200, 1.1122, 1.2570, 1.2765, 1.2928},
                                     {0.5140, 0.5480, 0.6110, 0.6080, 0.5780},
                                     {1.6000, 1.6100, 1.6400, 1.6500, 1.6300},
                                     {19.400, 19.500, 21.100, 19.200, 17.100}};
    for (int q = 0; q < 5; ++q) {
        const double *y = tab[q];
        if (zeff <= zA[0]) out[q] = y[0];
        else if (zeff >= zA[4]) out[q] = y[4];
        else {
            int i = 0;
            while (i < 3 && zeff >= zA[i + 1]) ++i;
            out[q] = y[i] + (y[i + 1] - y[i]) * (zeff - zA[i]) / (zA[i + 1] - zA[i]);
        }
    }
    double s = 0.8338 / dm.sigma8;
    *pk_scale = s * s * std::pow((1.0 + zeff) / (1.0 + dm.zref), -2.0);
}

// Small batches run the chi-square GEMM with the row tiles split over CTAs; this provides its scratch.
static int gemm_chi2_scratch(bf_ctx *ctx, cudaStream_t s, GemmArgs &g) {
    g.chi2_part = nullptr;
    g.part_ld = 0;
    if (g.N > gemm_chi2_split_columns(g.M, ctx->sm_count)) return BF_OK;
    const int ld = round_up(g.N, 64);
    const size_t need = (size_t)(g.M / 32) * ld * sizeof(double);  // row tiles of 32 when the batch is very small
    if (need > ctx->chi2_part_bytes) {
        BF_CUDA_OK(cudaStreamSynchronize(s));
        void *p = ctx->chi2_part;
        if (int rc = ensure(&p, &ctx->chi2_part_bytes, need)) return rc;
        ctx->chi2_part = (double *)p;
    }
    g.chi2_part = ctx->chi2_part;
    g.part_ld = ld;
    ctx->launches++;  // the kernel that sums the partials
    return BF_OK;
}

// Below this many vectors a persistent CTA per 32-column panel leaves most SMs idle (one panel is a serial chain of 66
// stages, ~80 us): small batches keep the separate passes, whose chi2 GEMM splits its row tiles over CTAs.
constexpr int kFusedMinBatch = 1024;

enum Output { OUT_PRED, OUT_CHI2, OUT_XIU, OUT_TRANSFORM, OUT_WHITENED /* y = L^-1 (pred - d), [n_pad][ldo] */ };

// Runs one chunk (B <= chunk capacity) entirely on the context's stream.
int run_chunk(bf_ctx *ctx, const bf_model *m, const bf_data *d, const bf_data::Plan *plan, const double *d_params,
              int B, Output what, double *d_out, int ldo, const double *d_override, int *d_status_out,
              double z_trigger_override, bool use_ztrig_override) {
    const DevModel &dm = m->dev;
    cudaStream_t s = ctx->stream;
    const int ldb = round_up(B, kPadN);
    const int nz = plan ? plan->nz : 1;
    // transform-only calls have no data: a single class holding the trigger redshift
    Workspace w;
    // fake data dims for transform-only
    static bf_data empty_data;
    const bf_data *dd = d ? d : &empty_data;
    size_t need = (ws_fixed_doubles(dm) + ws_doubles_per_column(dm, dd, nz) * (size_t)ldb) * sizeof(double);
    int rc = ensure(&ctx->ws, &ctx->ws_bytes, need);
    if (rc) return rc;
    carve(ctx->ws, dm, dd, nz, ldb, w);

    const double *d_zvals = plan ? plan->d_zvals : nullptr;
    if (!plan) {
        // OUT_TRANSFORM without data: one class = z_trigger, kept in a spare slot of the flag block
        BF_CUDA_OK(cudaMemcpyAsync(w.flag + 15, &z_trigger_override, sizeof(double), cudaMemcpyHostToDevice, s));
        d_zvals = w.flag + 15;
    }
    // coherence sort (chi2 only; outputs that are indexed by the original column stay unsorted)
    const bool sorted = what == OUT_CHI2 && !d_override && B >= 2048 && dm.kind != BF_MODEL_KSPACE_FFT;
    if (sorted) BF_KERNEL(ctx, "coherence_sort_kernels", launch_coherence_sort(s, dm, d_params, B, w.sort_hist, w.sort_base, w.rank));
    BF_KERNEL(ctx, "prep_kernel", launch_prep(s, dm, d_params, B, ldb, nz, d_zvals, sorted ? w.rank : nullptr, w.pT, w.S, w.status, plan && plan->z_retrigger, w.flag));
    double *chi2_dst = sorted ? w.chi2_sorted : d_out;

    // ---- k-space model: which chain runs is decided on the device (no host round trip inside a call) ----
    const Gate ungated{nullptr, 0};
    bool can_share = false;  // shared-transform fast path possible: polynomial tracer factor (no UV / HCD terms)
    bool host_shared = false, host_current = false;  // KeyHint: see below
    ProjArgs pa;
    std::memset(&pa, 0, sizeof(pa));
    // per-vector transform chain: multipoles of D(k,mu) -> Hankel GEMM -> spline second derivatives
    auto emit_transform = [&](Gate gate) -> int {
        pa.G = w.G;
        pa.gate = gate;
        BF_KERNEL(ctx, "kspace_project_kernel", launch_kspace_project(s, pa, ctx->sm_count));
        GemmArgs g{};
        g.gate = gate;
        g.A = dm.hankel; g.Bm = w.G; g.C = w.T;
        g.M = dm.nr_pad; g.N = B; g.K = dm.nk_pad;
        g.lda = dm.nk_pad; g.ldb = ldb; g.ldc = ldb;
        g.strideA = (long long)dm.nr_pad * dm.nk_pad; g.strideB = (long long)dm.nk_pad * ldb; g.strideC = (long long)dm.nr_pad * ldb;
        BF_KERNEL(ctx, "dgemm_store_kernel[hankel]", launch_gemm_store(s, g, 6, ctx->sm_count));
        if (what == OUT_TRANSFORM) {
            // out[((comp*3+l)*nr + j)*ldo + b]
            for (int t = 0; t < 6; ++t)
                BF_CUDA_OK(cudaMemcpy2DAsync(d_out + (size_t)t * dm.nr * ldo, (size_t)ldo * sizeof(double),
                                             w.T + (size_t)t * dm.nr_pad * ldb, (size_t)ldb * sizeof(double),
                                             (size_t)B * sizeof(double), dm.nr, cudaMemcpyDeviceToDevice, s));
            return BF_OK;
        }
        BF_KERNEL(ctx, "kspace_thomas_kernel",
                  launch_kspace_thomas(s, B, ldb, dm.nr, dm.nr_pad, dm.tr_dr, dm.th_w, dm.th_inv_diag, w.T, w.C2, gate));
        return BF_OK;
    };
    if (dm.kind == BF_MODEL_KSPACE) {
        pa.m = dm; pa.pT = w.pT; pa.S = w.S; pa.ldb = ldb; pa.B = B;
        pa.zc_trigger = plan ? plan->zc_trigger : 0;
        double ztrig = plan ? plan->zvals[plan->zc_trigger] : z_trigger_override;
        (void)use_ztrig_override;
        double zeff = dm.zeff > 0 ? dm.zeff : ztrig;
        nl_table(dm, zeff, pa.nl_tab, &pa.nl_pk_scale);
        if (what == OUT_TRANSFORM) return emit_transform(ungated);
        // Shared-transform fast path: possible when the tracer factor of D(k,mu) is polynomial (no UV / HCD terms) and
        // every vector of the chunk has the same key parameters. kspace_key_check_kernel compares the vectors,
        // kspace_key_decide_kernel compares their key with the one the cached basis tables were built for and sets
        // gate[0] (shared) / gate[1] (rebuild the tables); everything downstream is predicated on those two ints.
        can_share = !(dm.flags & (BF_FLAG_UVFLUCTUATION | BF_FLAG_HCD_MODEL | BF_FLAG_HCD_MODEL_ALT));
        if (can_share) {
            KeyExtras ex;
            ex.v[0] = zeff;
            ex.v[1] = pa.nl_pk_scale;
            for (int q = 0; q < 5; ++q) ex.v[2 + q] = pa.nl_tab[q];
            // Host-buffer calls of modest size have compared the key parameters of their vectors on the host (KeyHint): then
            // the per-vector chain is not enqueued at all, and when the key is also the one the basis tables on the device were
            // last built for (host mirror of d_keystate, valid as long as every call since went through a hint) nothing is
            // decided on the device: a fit's single-vector evaluations cost 6 launches instead of 25.
            const KeyHint *hint = ctx->key_hint;
            double hk[32] = {};
            if (hint) {
                for (int k = 0; k < hint->nkey; ++k) hk[k] = hint->vals[k];
                for (int k = 0; k < 7; ++k) hk[kKeyMax - 1 + k] = ex.v[k];
                host_shared = true;
                host_current = m->h_key_valid;
                for (int k = 0; k < 32 && host_current; ++k) host_current = hk[k] == m->h_key[k];
                std::memcpy(ctx->last_hk, hk, sizeof(hk));
                ctx->last_host_current = host_current;
            }
            if (!host_current) {
                BF_KERNEL(ctx, "kspace_key_check_kernel", launch_kspace_key_check(s, dm, w.pT, ldb, B, w.flag));
                BF_KERNEL(ctx, "kspace_key_decide_kernel", launch_kspace_key_decide(s, w.flag, m->d_keystate, m->d_gate, ex));
                const Gate rebuild{m->d_gate + 1, 1};
                pa.G = w.Gb;
                pa.gate = rebuild;
                BF_KERNEL(ctx, "kspace_basis_project_kernel", launch_kspace_basis_project(s, pa));
                GemmArgs g{};
                g.gate = rebuild;
                g.A = dm.hankel; g.Bm = w.Gb; g.C = w.Tb;
                g.M = dm.nr_pad; g.N = 3; g.K = dm.nk_pad;
                g.lda = dm.nk_pad; g.ldb = kPadN; g.ldc = kPadN;
                g.strideA = (long long)dm.nr_pad * dm.nk_pad; g.strideB = (long long)dm.nk_pad * kPadN; g.strideC = (long long)dm.nr_pad * kPadN;
                BF_KERNEL(ctx, "dgemm_store_kernel[hankel-basis]", launch_gemm_store(s, g, 6, ctx->sm_count));
                BF_KERNEL(ctx, "kspace_basis_finish_kernel", launch_kspace_basis_finish(s, dm, w.Tb, m->d_tabs_cache, rebuild));
                // after a call whose vectors share a key the tables are those of that key; otherwise the host cannot know
                m->h_key_valid = host_shared;
                if (host_shared) std::memcpy(m->h_key, hk, sizeof(hk));
            }
        }
    }

    if (dm.kind == BF_MODEL_KSPACE_FFT) {
        const DevFft &f = dm.fft;
        FftVecArgs va;
        std::memset(&va, 0, sizeof(va));
        va.m = dm; va.pT = w.pT; va.ldb = ldb; va.B = B;
        if (d) { va.r0 = d->h_r[0]; va.mu0 = d->trigger_mu_zero ? 0.0 : d->h_mu[0]; va.z0 = d->h_z[0]; }
        va.use_ztrig = d ? 0 : 1;
        va.ztrig = z_trigger_override;
        va.keyflag = w.flag; va.keys = w.fft_keys; va.c12 = w.fft_c12; va.cont = w.fft_cont; va.status = w.status;
        BF_KERNEL(ctx, "fft_vec_kernel", launch_fft_vec(s, va));
        BF_CUDA_OK(cudaMemcpyAsync(ctx->pinned_flag, w.flag, sizeof(double) * kFftKeyMax, cudaMemcpyDeviceToHost, s));
        BF_CUDA_OK(cudaStreamSynchronize(s));
        // key = SigmaNL-perp, 1+f, nlcorr qnl, nlcorr kp, z of the triggering bin: what the per-key
        // tables read. The reference re-runs its 3-D FFTs when one of these changes (:191-214); here
        // the tables are rebuilt, once per distinct key of the chunk.
        constexpr int NK = 5;
        auto tables_for = [&](const std::vector<double> &key) {
            if (m->cache_valid && key == m->cached_key) return;
            FftKeyState ks;
            const double snl_perp = key[0], snl_par = key[0] * key[1], zeff = key[4];
            ks.snl_perp2 = snl_perp * snl_perp;
            ks.snl_par2 = snl_par * snl_par;
            nl_table(dm, zeff, ks.nl_tab, &ks.nl_pk_scale);
            if (dm.flags & BF_FLAG_FIT_NL_CORRECTION) { ks.nl_tab[0] = key[2]; ks.nl_tab[4] = key[3]; }
            ProfScope ps1(ctx, "fft_hsum_kernel+fft_xtrans_kernel");
            ctx->launches++;  // two kernels
            launch_fft_tables(s, dm, ks, m->d_fft_H, m->d_fft_T);
            m->cached_key = key;
            m->cache_valid = true;
        };
        if (ctx->pinned_flag[0] == 0.0) {
            tables_for(std::vector<double>(ctx->pinned_flag + 1, ctx->pinned_flag + 1 + NK));
            BF_KERNEL(ctx, "fft_plane_kernel",
                      launch_fft_planes(s, dm, m->d_fft_T, w.fft_c12, w.fft_cont, ldb, nullptr, B, w.fft_planes));
        } else {
            // several keys in one chunk (e.g. the finite-difference probes of a fit that floats
            // SigmaNL): group the columns by key, one table build + one plane launch per group
            std::vector<double> hk((size_t)NK * ldb);
            BF_CUDA_OK(cudaMemcpyAsync(hk.data(), w.fft_keys, hk.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
            BF_CUDA_OK(cudaStreamSynchronize(s));
            std::map<std::vector<double>, std::vector<int>> groups;
            std::vector<double> key(NK);
            for (int b = 0; b < B; ++b) {
                for (int q = 0; q < NK; ++q) key[q] = hk[(size_t)q * ldb + b];
                groups[key].push_back(b);
            }
            std::vector<int> order;
            order.reserve(B);
            for (auto &kv : groups) order.insert(order.end(), kv.second.begin(), kv.second.end());
            BF_CUDA_OK(cudaMemcpyAsync(w.rank, order.data(), sizeof(int) * B, cudaMemcpyHostToDevice, s));
            BF_CUDA_OK(cudaStreamSynchronize(s));  // `order` is pageable host memory
            int first = 0;
            for (auto &kv : groups) {
                tables_for(kv.first);
                BF_KERNEL(ctx, "fft_plane_kernel", launch_fft_planes(s, dm, m->d_fft_T, w.fft_c12, w.fft_cont, ldb, w.rank + first,
                                                                     (int)kv.second.size(), w.fft_planes));
                first += (int)kv.second.size();
            }
        }
        if (what == OUT_TRANSFORM) {
            // out[((b*2 + comp)*npy + iy)*npx + ix]
            BF_CUDA_OK(cudaMemcpy2DAsync(d_out, (size_t)f.npx * sizeof(double), w.fft_planes, (size_t)f.npx_pad * sizeof(double),
                                         (size_t)f.npx * sizeof(double), (size_t)B * 2 * f.npy_pad, cudaMemcpyDeviceToDevice, s));
            BF_CUDA_OK(cudaGetLastError());
            return BF_OK;
        }
        if (what == OUT_XIU) BF_FAIL(BF_ERR_ANALYSIS, "bf_eval_undistorted_batch: the model has no distortion matrix");
        const bool to_user = what == OUT_PRED;
        FftEvalArgs ea;
        std::memset(&ea, 0, sizeof(ea));
        ea.m = dm; ea.pT = w.pT; ea.ldb = ldb; ea.B = B; ea.npts = d->n;
        ea.r = d->d_r; ea.mu = d->d_mu; ea.z = d->d_z; ea.planes = w.fft_planes;
        ea.out = to_user ? d_out : w.Z; ea.ldo = to_user ? ldo : ldb;
        ea.dsub = to_user ? nullptr : d->d_data; ea.dov = to_user ? nullptr : d_override;
        ea.status = w.status;
        if (!to_user && d->n_pad > d->n)
            BF_CUDA_OK(cudaMemsetAsync(w.Z + (size_t)d->n * ldb, 0, (size_t)(d->n_pad - d->n) * ldb * sizeof(double), s));
        BF_KERNEL(ctx, "fft_eval_kernel", launch_fft_eval(s, ea));
        if (what == OUT_CHI2 || what == OUT_WHITENED) {
            GemmArgs g{};
            g.A = d->d_W; g.Bm = w.Z; g.C = nullptr;
            g.M = d->n_pad; g.N = B; g.K = d->n_pad;
            g.lda = d->n_pad; g.ldb = ldb; g.ldc = 0;
            g.lower_triangular = 1;
            if (what == OUT_CHI2) {
                g.chi2 = d_out; g.status = w.status;
                if (int rc = gemm_chi2_scratch(ctx, s, g)) return rc;
                BF_KERNEL(ctx, "dgemm_chi2_kernel", launch_gemm_chi2(s, g, ctx->sm_count));
            } else {
                g.C = d_out; g.ldc = ldo;
                BF_KERNEL(ctx, "dgemm_store_kernel[whiten]", launch_gemm_store(s, g, 1, ctx->sm_count));
            }
        }
        if (d_status_out) BF_CUDA_OK(cudaMemcpyAsync(d_status_out, w.status, sizeof(int) * B, cudaMemcpyDeviceToDevice, s));
        BF_CUDA_OK(cudaGetLastError());
        return BF_OK;
    }

    // Everything from the model evaluation to the chi-square / prediction, for one of the two table sources.
    auto emit = [&](bool shared_tables, Gate gate) -> int {
        EvalArgs a;
        std::memset(&a, 0, sizeof(a));
        a.m = dm; a.pT = w.pT; a.S = w.S; a.ldb = ldb; a.B = B;
        a.vfac = plan->d_vfac; a.T = w.T; a.C2 = w.C2; a.status = w.status; a.gate = gate;
        a.tabs = m->d_tabs_cache; a.zc_trigger = plan->zc_trigger;
        const bool use_dm = dm.dm_order > 0;
        const bool to_user = what == OUT_PRED;  // final rows go straight to the caller's buffer

        auto fill_tab = [&](const bf_data::Plan::PointTabDev &t, bool grid, PointTab &pt) {
            pt.npts = grid ? d->n_grid : d->n;
            pt.npts_pad = grid ? d->n_grid_pad : d->n_pad;
            pt.rpar0 = t.rpar0; pt.rperp0 = t.rperp0;
            pt.z = grid ? d->d_grid_z : d->d_z;
            pt.zc = grid ? plan->d_zc_grid : plan->d_zc_data;
            pt.gidx = grid ? nullptr : d->d_index;
            pt.zc_metal = grid ? plan->d_zc_metal_grid : plan->d_zc_metal_data;
            pt.zc_stride = grid ? d->n_grid_pad : d->n_pad;
            for (int q = 0; q < 4; ++q) { pt.bb_rt[q] = t.bb_rt[q]; pt.bb_z[q] = t.bb_z[q]; pt.n_rt[q] = t.n_rt[q]; pt.n_z[q] = t.n_z[q]; }
            pt.mx = t.mx; pt.mxh = t.mxh; pt.mx_ny = dm.metal_ny; pt.mx_nt = dm.metal_ncomb * 3;
        };
        FastArgs fa;
        std::memset(&fa, 0, sizeof(fa));
        fa.m = dm; fa.pT = w.pT; fa.S = w.S; fa.vfac = plan->d_vfac; fa.tabs = m->d_tabs_cache; fa.ldb = ldb; fa.B = B;
        fa.nz = plan->nz; fa.zc_trigger = plan->zc_trigger; fa.status = w.status; fa.gate = gate;
        if (dm.kind == BF_MODEL_RSPACE && m->d_rs_tabs) {
            // r-space model through the fast pass (tables are static)
            fa.tabs = m->d_rs_tabs;
            fa.rspace = 1;
        }
        bool finished = false;
        if (use_dm) {
            // undistorted xi on the full grid
            a.npts = d->n_grid; a.npts_pad = d->n_grid_pad;
            a.r = d->d_grid_r; a.mu = d->d_grid_mu; a.z = d->d_grid_z; a.zc = plan->d_zc_grid; a.gidx = nullptr;
            a.zc_metal = plan->d_zc_metal_grid; a.zc_stride = d->n_grid_pad;
            a.mode = 1;
            a.out = w.XiU; a.ldo = ldb; a.dsub = nullptr; a.dov = nullptr;
            if (what == OUT_XIU) { a.out = d_out; a.ldo = ldo; a.npts_pad = a.npts; }
            // Headline path: the model pass is the producer of the chi2 GEMM's right-hand operand (kernels_fused.cu);
            // xi_u is never written to HBM. BAOFIT_UNFUSED=1 keeps the separate passes (A/B measurements).
            static const bool no_fused = std::getenv("BAOFIT_UNFUSED") != nullptr;
            if (what == OUT_CHI2 && shared_tables && !d_override && plan->fold_ok && plan->d_WAf && plan->nz == 1 && !no_fused && B >= kFusedMinBatch) {
                FusedArgs fu;
                std::memset(&fu, 0, sizeof(fu));
                fu.gate = gate;
                fu.m = dm;
                fill_tab(plan->tab_grid, true, fu.pt);
                fu.pT = w.pT; fu.S = w.S; fu.vfac = plan->d_vfac; fu.tabs = m->d_tabs_cache; fu.ldb = ldb; fu.B = B;
                fu.zc_trigger = plan->zc_trigger;
                fu.At = plan->d_WAf; fu.At3 = plan->d_WAf3; fu.tri = plan->fused_tri ? 1 : 0; fu.M = d->n_pad; fu.kgrid = d->n_grid_pad; fu.K = d->n_grid_pad + plan->fold_kext;
                fu.rows = w.XiU + (size_t)d->n_grid_pad * ldb;
                fu.chi2 = chi2_dst; fu.status = w.status;
                if (fused_chi2_supported(fu)) {
                    FoldArgs fo;
                    std::memset(&fo, 0, sizeof(fo));
                    fo.gate = gate;
                    fo.m = dm; fo.pT = w.pT; fo.S = w.S; fo.vfac = plan->d_vfac; fo.ldb = ldb; fo.B = B;
                    fo.zc_data = plan->zc_trigger;
                    fo.fold_case = plan->fold_case; fo.nb = plan->fold_nb; fo.kext = plan->fold_kext; fo.pmax = plan->fold_pmax;
                    fo.nrt = plan->fold_nrt; fo.nz = plan->fold_nz;
                    fo.minus_d_row_value = 1;
                    fo.rows = w.XiU + (size_t)d->n_grid_pad * ldb;
                    fo.npts = d->n; fo.rpar0 = plan->tab_data.rpar0; fo.rperp0 = plan->tab_data.rperp0; fo.status = w.status;
                    BF_KERNEL(ctx, "fold_coef_kernel", launch_fold_coef(s, fo));
                    BF_KERNEL(ctx, "fused_model_chi2_kernel", launch_fused_model_chi2(s, fu, ctx->sm_count));
                    finished = true;
                }
            }
            if (finished) {
                // nothing left to do for this chunk
            } else if (shared_tables) {
                fill_tab(plan->tab_grid, true, fa.pt);
                fa.mode = 1; fa.out = a.out; fa.ldo = a.ldo;
                if (what == OUT_XIU) fa.pt.npts_pad = fa.pt.npts;
                if (fast_eval_is_lean(fa)) {
                    BF_KERNEL(ctx, "kspace_fast_eval_kernel<cosmology>[grid]", launch_kspace_fast_eval(s, fa, ctx->sm_count, 1));
                    if (fast_eval_has_extras(fa))
                        BF_KERNEL(ctx, "kspace_fast_eval_kernel<extras>[grid]", launch_kspace_fast_eval(s, fa, ctx->sm_count, 2));
                } else
                    BF_KERNEL(ctx, "kspace_fast_eval_kernel<all>[grid]", launch_kspace_fast_eval(s, fa, ctx->sm_count, 0));
            } else
                BF_KERNEL(ctx, "kspace_eval_kernel[grid]", launch_kspace_eval(s, a, ctx->sm_count));
            if (what == OUT_XIU) finished = true;
            if (!finished && plan->fold_ok && !d_override) {
                // fused tail: the post-matrix broadband and "- d" are extra K rows of the panel, and for
                // chi2 the whitening W is folded into the left matrix: one GEMM, squares in the epilogue
                FoldArgs fo;
                std::memset(&fo, 0, sizeof(fo));
                    fo.gate = gate;
                fo.m = dm; fo.pT = w.pT; fo.S = w.S; fo.vfac = plan->d_vfac; fo.ldb = ldb; fo.B = B;
                fo.zc_data = plan->zc_trigger;  // single class of the data bins (= class of the first kept bin)
                fo.fold_case = plan->fold_case; fo.nb = plan->fold_nb; fo.kext = plan->fold_kext; fo.pmax = plan->fold_pmax;
                fo.nrt = plan->fold_nrt; fo.nz = plan->fold_nz;
                fo.minus_d_row_value = (what == OUT_CHI2 || what == OUT_WHITENED) ? 1 : 0;
                fo.rows = w.XiU + (size_t)d->n_grid_pad * ldb;
                fo.npts = d->n; fo.rpar0 = plan->tab_data.rpar0; fo.rperp0 = plan->tab_data.rperp0; fo.status = w.status;
                BF_KERNEL(ctx, "fold_coef_kernel", launch_fold_coef(s, fo));
                GemmArgs g{};
            g.gate = gate;
                g.Bm = w.XiU;
                g.M = d->n_pad; g.N = B; g.K = d->n_grid_pad + plan->fold_kext;
                g.lda = g.K; g.ldb = ldb;
                if (what == OUT_CHI2) {
                    g.A = plan->d_WA; g.C = nullptr; g.ldc = 0;
                    g.chi2 = chi2_dst; g.status = w.status;
                    // opt-in: reads the panel from HBM exactly once, but streams the left matrix from L2 twice as
                    // often (32-column panels) and measures slower on B200 (0.63 vs 0.56 ms)
                    static const bool use_fullm = std::getenv("BAOFIT_FULLM_GEMM") != nullptr;
                    if (plan->d_WAt && use_fullm) {
                        g.A = plan->d_WAt;
                        BF_KERNEL(ctx, "dgemm_chi2_kernel[fused tail]", launch_gemm_chi2_fullm(s, g, ctx->sm_count));
                    } else {
                        if (int rc = gemm_chi2_scratch(ctx, s, g)) return rc;
                        BF_KERNEL(ctx, "dgemm_chi2_kernel[fused tail]", launch_gemm_chi2(s, g, ctx->sm_count));
                    }
                } else if (what == OUT_WHITENED) {
                    // whitened residuals for the normal-equation probes: the same folded left matrix, store epilogue
                    g.A = plan->d_WA; g.C = d_out; g.ldc = ldo;
                    BF_KERNEL(ctx, "dgemm_store_kernel[whiten, fused tail]", launch_gemm_store(s, g, 1, ctx->sm_count));
                } else {
                    g.A = plan->d_Aext; g.C = w.Z; g.ldc = ldb;
                    BF_KERNEL(ctx, "dgemm_store_kernel[distortion+broadband]", launch_gemm_store(s, g, 1, ctx->sm_count));
                    BF_CUDA_OK(cudaMemcpy2DAsync(d_out, (size_t)ldo * sizeof(double), w.Z, (size_t)ldb * sizeof(double),
                                                 (size_t)B * sizeof(double), d->n, cudaMemcpyDeviceToDevice, s));
                }
                finished = true;
            }
            if (!finished) {
                GemmArgs g{};
            g.gate = gate;
                g.A = plan->d_Dkept; g.Bm = w.XiU; g.C = w.Y;
                g.M = d->n_pad; g.N = B; g.K = d->n_grid_pad;
                g.lda = d->n_grid_pad; g.ldb = ldb; g.ldc = ldb;
                BF_KERNEL(ctx, "dgemm_store_kernel[distortion]", launch_gemm_store(s, g, 1, ctx->sm_count));
                // tail on the data bins: post-matrix broadband, dilation check, residual.
                // Y (ld = ldb) -> Z (ld = ldb); predictions are copied out row by row afterwards.
                a.npts = d->n; a.npts_pad = d->n_pad;
                a.r = d->d_r; a.mu = d->d_mu; a.z = d->d_z; a.zc = plan->d_zc_data; a.gidx = d->d_index;
                a.in = w.Y; a.out = w.Z; a.ldo = ldb;
                a.dsub = to_user ? nullptr : d->d_data; a.dov = to_user ? nullptr : d_override;
                fill_tab(plan->tab_data, false, fa.pt);
                fa.mode = 0; fa.in = w.Y; fa.out = w.Z; fa.ldo = ldb; fa.dsub = a.dsub; fa.dov = a.dov;
                BF_KERNEL(ctx, "fast_post_kernel", launch_fast_post(s, fa, ctx->sm_count));
                if (to_user)
                    BF_CUDA_OK(cudaMemcpy2DAsync(d_out, (size_t)ldo * sizeof(double), w.Z, (size_t)ldb * sizeof(double),
                                                 (size_t)B * sizeof(double), d->n, cudaMemcpyDeviceToDevice, s));
            }
        } else {
            if (what == OUT_XIU) BF_FAIL(BF_ERR_ANALYSIS, "bf_eval_undistorted_batch: the model has no distortion matrix");
            a.npts = d->n; a.npts_pad = to_user ? d->n : d->n_pad;
            a.r = d->d_r; a.mu = d->d_mu; a.z = d->d_z; a.zc = plan->d_zc_data; a.gidx = d->d_index;
            a.zc_metal = plan->d_zc_metal_data; a.zc_stride = d->n_pad;
            a.mode = 0;
            a.out = to_user ? d_out : w.Z; a.ldo = to_user ? ldo : ldb;
            a.dsub = to_user ? nullptr : d->d_data; a.dov = to_user ? nullptr : d_override;
            if (dm.kind == BF_MODEL_RSPACE && !shared_tables) BF_KERNEL(ctx, "rspace_eval_kernel", launch_rspace_eval(s, a, ctx->sm_count));
            else if (shared_tables && plan->fold_ok && plan->fold_data_mode && (what == OUT_CHI2 || what == OUT_WHITENED)) {
                // fused data tail: cosmology pass into Z, broadband coefficients and the "- d" switch as extra rows,
                // one GEMM against [W | W.basis | -W.d] (triangular block + dense tail columns)
                fill_tab(plan->tab_data, false, fa.pt);
                fa.mode = 0; fa.out = w.Z; fa.ldo = ldb; fa.dsub = nullptr; fa.dov = d_override;
                fa.sub_dov_in_part1 = d_override ? 1 : 0;
                BF_KERNEL(ctx, "kspace_fast_eval_kernel<cosmology>[data]", launch_kspace_fast_eval(s, fa, ctx->sm_count, 1));
                FoldArgs fo;
                std::memset(&fo, 0, sizeof(fo));
                    fo.gate = gate;
                fo.m = dm; fo.pT = w.pT; fo.S = w.S; fo.vfac = plan->d_vfac; fo.ldb = ldb; fo.B = B;
                fo.zc_data = plan->zc_trigger;
                fo.fold_case = plan->fold_case; fo.nb = plan->fold_nb; fo.kext = plan->fold_kext; fo.pmax = plan->fold_pmax;
                fo.nrt = plan->fold_nrt; fo.nz = plan->fold_nz; fo.ncls = plan->fold_ncls;
                fo.minus_d_row_value = d_override ? 0 : 1;  // with an override the data were subtracted column by column
                fo.rows = w.Z + (size_t)d->n_pad * ldb;
                fo.npts = d->n; fo.rpar0 = plan->tab_data.rpar0; fo.rperp0 = plan->tab_data.rperp0; fo.status = w.status;
                BF_KERNEL(ctx, "fold_coef_kernel", launch_fold_coef(s, fo));
                GemmArgs g{};
            g.gate = gate;
                g.A = plan->d_WA; g.Bm = w.Z;
                g.M = d->n_pad; g.N = B; g.K = d->n_pad + plan->fold_kext;
                g.lda = g.K; g.ldb = ldb;
                g.lower_triangular = 1; g.k_tri = d->n_pad;
                if (what == OUT_CHI2) {
                    g.C = nullptr; g.ldc = 0; g.chi2 = chi2_dst; g.status = w.status;
                    if (int rc = gemm_chi2_scratch(ctx, s, g)) return rc;
                    BF_KERNEL(ctx, "dgemm_chi2_kernel[fused data tail]", launch_gemm_chi2(s, g, ctx->sm_count));
                } else {
                    g.C = d_out; g.ldc = ldo;
                    BF_KERNEL(ctx, "dgemm_store_kernel[whiten, fused data tail]", launch_gemm_store(s, g, 1, ctx->sm_count));
                }
                finished = true;
            }
            else if (shared_tables) {
                fill_tab(plan->tab_data, false, fa.pt);
                fa.pt.npts_pad = a.npts_pad;
                fa.mode = 0; fa.out = a.out; fa.ldo = a.ldo; fa.dsub = a.dsub; fa.dov = a.dov;
                if (fast_eval_is_lean(fa)) {
                    BF_KERNEL(ctx, "kspace_fast_eval_kernel<cosmology>[data]", launch_kspace_fast_eval(s, fa, ctx->sm_count, 1));
                    if (fast_eval_has_extras(fa))
                        BF_KERNEL(ctx, "kspace_fast_eval_kernel<extras>[data]", launch_kspace_fast_eval(s, fa, ctx->sm_count, 2));
                } else
                    BF_KERNEL(ctx, "kspace_fast_eval_kernel<all>[data]", launch_kspace_fast_eval(s, fa, ctx->sm_count, 0));
            }
            else BF_KERNEL(ctx, "kspace_eval_kernel[data]", launch_kspace_eval(s, a, ctx->sm_count));
        }
        if (what == OUT_CHI2 && !finished) {
            GemmArgs g{};
            g.gate = gate;
            g.A = d->d_W; g.Bm = w.Z; g.C = nullptr;
            g.M = d->n_pad; g.N = B; g.K = d->n_pad;
            g.lda = d->n_pad; g.ldb = ldb; g.ldc = 0;
            g.lower_triangular = 1;
            g.chi2 = chi2_dst; g.status = w.status;
            if (int rc = gemm_chi2_scratch(ctx, s, g)) return rc;
            BF_KERNEL(ctx, "dgemm_chi2_kernel", launch_gemm_chi2(s, g, ctx->sm_count));
        }
        if (what == OUT_WHITENED && !finished) {
            GemmArgs g{};
            g.gate = gate;
            g.A = d->d_W; g.Bm = w.Z; g.C = d_out;
            g.M = d->n_pad; g.N = B; g.K = d->n_pad;
            g.lda = d->n_pad; g.ldb = ldb; g.ldc = ldo;
            g.lower_triangular = 1;
            BF_KERNEL(ctx, "dgemm_store_kernel[whiten]", launch_gemm_store(s, g, 1, ctx->sm_count));
        }
        return BF_OK;
    };
    if (dm.kind == BF_MODEL_KSPACE && can_share) {
        // both chains are enqueued; gate[0], written by kspace_key_decide_kernel, lets exactly one of them run
        if ((rc = emit(true, host_current ? ungated : Gate{m->d_gate, 1}))) return rc;
        if (!host_shared) {
            if ((rc = emit_transform(Gate{m->d_gate, 0}))) return rc;
            if ((rc = emit(false, Gate{m->d_gate, 0}))) return rc;
        }
    } else if (dm.kind == BF_MODEL_KSPACE) {
        if ((rc = emit_transform(ungated))) return rc;
        if ((rc = emit(false, ungated))) return rc;
    } else {
        if ((rc = emit(dm.kind == BF_MODEL_RSPACE && m->d_rs_tabs != nullptr, ungated))) return rc;
    }
    if (sorted) {
        BF_KERNEL(ctx, "unpermute_kernel", launch_unpermute(s, B, w.rank, w.chi2_sorted, w.status, d_out, d_status_out));
    } else if (d_status_out)
        BF_CUDA_OK(cudaMemcpyAsync(d_status_out, w.status, sizeof(int) * B, cudaMemcpyDeviceToDevice, s));
    BF_CUDA_OK(cudaGetLastError());
    return BF_OK;
}

int chunk_capacity(const bf_model *m, const bf_data *d, int nz) {
    static bf_data empty_data;
    size_t per_col = ws_doubles_per_column(m->dev, d ? d : &empty_data, nz) * sizeof(double);
    size_t budget = ((size_t)8 << 30) - ws_fixed_doubles(m->dev) * sizeof(double);  // 8 GiB of workspace at most
    size_t cap = budget / per_col;
    cap = std::min<size_t>(cap, 65536);
    cap = std::max<size_t>(cap / kPadN * kPadN, kPadN);
    return (int)cap;
}

int check_args(bf_ctx *ctx, const bf_model *m, const bf_data *d, const void *params, int B) {
    if (!ctx || !m || !d || !params) BF_FAIL(BF_ERR_OPTIONS, "null argument");
    if (B <= 0) BF_FAIL(BF_ERR_OPTIONS, "batch size must be positive");
    if (m->ctx != ctx || d->ctx != ctx) BF_FAIL(BF_ERR_OPTIONS, "model/data belong to a different context");
    return BF_OK;
}
}  // namespace

static int run_device(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *d_params, int B, Output what,
                      double *d_out, int ldo, const double *d_override, int *d_status);

// Multipole-binned data: evaluate the model at the expanded (r, mu_q, z) points, project, then the
// usual chi-square GEMM on the residual panel.
static int run_device_multipole(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *d_params, int B, Output what,
                                double *d_out, int ldo, const double *d_override, int *d_status) {
    if (what != OUT_PRED && what != OUT_CHI2 && what != OUT_WHITENED) BF_FAIL(BF_ERR_ANALYSIS, "multipole data: only predictions and chi-square are defined");
    if (m->dev.dm_order > 0) BF_FAIL(BF_ERR_DATA, "multipole data cannot be combined with a distortion matrix");
    if (m->dev.idx_delta_v >= 0)
        BF_FAIL(BF_ERR_DATA, "multipole data with a cross-correlation model is not supported (the reference's multipole overload "
                             "applies no velocity shift, AbsCorrelationModel.cc:43-49)");
    const bf_data *e = d->expanded;
    const int cap = 8192;  // columns per pass: bounds the panel of expanded predictions
    cudaStream_t s = ctx->stream;
    for (int b0 = 0; b0 < B; b0 += cap) {
        const int bc = std::min(cap, B - b0), ldb = round_up(bc, kPadN);
        const size_t need = ((size_t)e->n + d->n_pad) * ldb * sizeof(double);
        int rc = ensure(&ctx->mp, &ctx->mp_bytes, need);
        if (rc) return rc;
        double *X = (double *)ctx->mp, *Z = X + (size_t)e->n * ldb;
        rc = run_device(ctx, m, e, d_params + (size_t)b0 * m->dev.npar, bc, OUT_PRED, X, ldb, nullptr, d_status ? d_status + b0 : nullptr);
        if (rc) return rc;
        if (what == OUT_PRED) {
            BF_KERNEL(ctx, "multipole_project_kernel",
                      launch_multipole_project(s, X, ldb, d->d_proj, d->n, d->nmu, bc, d_out + b0, ldo, nullptr, nullptr));
        } else {
            if (d->n_pad > d->n) BF_CUDA_OK(cudaMemsetAsync(Z + (size_t)d->n * ldb, 0, (size_t)(d->n_pad - d->n) * ldb * sizeof(double), s));
            BF_KERNEL(ctx, "multipole_project_kernel",
                      launch_multipole_project(s, X, ldb, d->d_proj, d->n, d->nmu, bc, Z, ldb, d->d_data,
                                               d_override ? d_override + (size_t)b0 * d->n : nullptr));
            GemmArgs g{};
            g.A = d->d_W; g.Bm = Z; g.C = nullptr;
            g.M = d->n_pad; g.N = bc; g.K = d->n_pad;
            g.lda = d->n_pad; g.ldb = ldb; g.ldc = 0;
            g.lower_triangular = 1;
            if (what == OUT_CHI2) {
                g.chi2 = d_out + b0; g.status = d_status ? d_status + b0 : nullptr;
                if (int rc = gemm_chi2_scratch(ctx, s, g)) return rc;
                BF_KERNEL(ctx, "dgemm_chi2_kernel", launch_gemm_chi2(s, g, ctx->sm_count));
            } else {
                g.C = d_out + b0; g.ldc = ldo;
                BF_KERNEL(ctx, "dgemm_store_kernel[whiten]", launch_gemm_store(s, g, 1, ctx->sm_count));
            }
        }
    }
    BF_CUDA_OK(cudaGetLastError());
    return BF_OK;
}

static int run_device(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *d_params, int B, Output what,
                      double *d_out, int ldo, const double *d_override, int *d_status) {
    int rc = check_args(ctx, m, d, d_params, B);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    if (d->binning_type == BF_BINNING_MULTIPOLE)
        return run_device_multipole(ctx, m, d, d_params, B, what, d_out, ldo, d_override, d_status);
    const bf_data::Plan *plan = nullptr;
    if ((rc = get_plan(m, d, &plan))) return rc;
    int cap = chunk_capacity(m, d, plan->nz);
    // equal-sized chunks (a small trailing chunk would under-fill the GPU)
    int nchunks = (B + cap - 1) / cap;
    cap = round_up((B + nchunks - 1) / nchunks, kPadN);
    for (int b0 = 0; b0 < B; b0 += cap) {
        int bc = std::min(cap, B - b0);
        double *o = d_out + b0;  // chi2[b] or column offset in bin-major outputs
        const double *ov = d_override ? d_override + (size_t)b0 * d->n : nullptr;
        rc = run_chunk(ctx, m, d, plan, d_params + (size_t)b0 * m->dev.npar, bc, what, o, ldo, ov,
                       d_status ? d_status + b0 : nullptr, 0.0, false);
        if (rc) return rc;
    }
    return BF_OK;
}

extern "C" int bf_eval_batch_dev(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *d_params, int B,
                                 double *d_pred, int ldp, int *d_status) {
    if (!d_pred || ldp < B) BF_FAIL(BF_ERR_OPTIONS, "bf_eval_batch_dev: bad output buffer");
    return run_device(ctx, m, d, d_params, B, OUT_PRED, d_pred, ldp, nullptr, d_status);
}

extern "C" int bf_chi2_batch_dev(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *d_params, int B,
                                 const double *d_override, double *d_chi2, int *d_status) {
    if (!d_chi2) BF_FAIL(BF_ERR_OPTIONS, "bf_chi2_batch_dev: null output");
    return run_device(ctx, m, d, d_params, B, OUT_CHI2, d_chi2, 0, d_override, d_status);
}

// graphs of small-call kernel chains hold device pointers of a model and a data set: dropped with either
// (model_serial = 0 and data = null: drop all)
static void purge_graphs(bf_ctx *ctx, unsigned long long model_serial, const void *data) {
    for (auto it = ctx->graphs.begin(); it != ctx->graphs.end();) {
        const bool hit = (!model_serial && !data) || (model_serial && it->first.model_serial == model_serial) || (data && it->first.data == data);
        if (hit) {
            if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
            it = ctx->graphs.erase(it);
        } else
            ++it;
    }
}

// kspace_key_check_kernel on the host, for host-buffer calls small enough that looking at the parameters costs less than
// the launches it saves: true when every vector carries the same key parameters (same order and comparison as the kernel).
static bool host_key_check(const DevModel &m, const double *params, int B, KeyHint *hint, const double *steps) {
    constexpr int kHostKeyCheckMax = 4096;
    if (m.kind != BF_MODEL_KSPACE || B > kHostKeyCheckMax) return false;
    if (m.flags & (BF_FLAG_UVFLUCTUATION | BF_FLAG_HCD_MODEL | BF_FLAG_HCD_MODEL_ALT)) return false;
    int idx[kKeyMax], n = 0;
    idx[n++] = m.idx_nl; idx[n++] = m.idx_nl + 1;
    if (m.flags & (BF_FLAG_BIN_SMOOTH | BF_FLAG_BIN_SMOOTH_ALT)) { idx[n++] = m.idx_bs; idx[n++] = m.idx_bs + 1; }
    if (m.flags & BF_FLAG_SMOOTH_GAUSS) idx[n++] = m.idx_smgaus;
    if (m.flags & BF_FLAG_SMOOTH_LORENTZ) idx[n++] = m.idx_smlor;
    if (m.flags & BF_FLAG_FIT_NL_CORRECTION) { idx[n++] = m.idx_nlcorr; idx[n++] = m.idx_nlcorr + 1; }
    hint->nkey = n;
    for (int k = 0; k < n; ++k) hint->vals[k] = params[idx[k]];
    for (int b = 1; b < B; ++b) {
        const double *p = params + (size_t)b * m.npar;
        for (int k = 0; k < n; ++k)
            if (p[idx[k]] != hint->vals[k]) return false;
    }
    for (int k = 0; k < n; ++k)
        if (hint->vals[k] != hint->vals[k]) return false;  // NaN never compares equal on the device either
    if (steps)  // probe groups expanded on the device from (centre, steps): no key parameter may be probed
        for (int b = 0; b < B; ++b)
            for (int k = 0; k < n; ++k)
                if (steps[(size_t)b * m.npar + idx[k]] != 0.0) return false;
    return true;
}

// RAII: the hint is visible to run_chunk for the duration of one run_device call
struct KeyHintScope {
    bf_ctx *ctx;
    KeyHint hint;
    KeyHintScope(bf_ctx *c, const DevModel &m, const double *params, int B, const double *steps = nullptr) : ctx(c) {
        ctx->key_hint = host_key_check(m, params, B, &hint, steps) ? &hint : nullptr;
    }
    ~KeyHintScope() { ctx->key_hint = nullptr; }
};

// host-buffer entry points: stage through device memory owned by the context
static int run_host(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *params, int B, Output what,
                    double *out, int ldo, int out_rows, const double *override_h, int *status) {
    int rc = check_args(ctx, m, d, params, B);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    const int npar = m->dev.npar;
    size_t nb_params = (size_t)B * npar * sizeof(double);
    size_t nb_out = what == OUT_CHI2 ? (size_t)B * sizeof(double) : (size_t)out_rows * B * sizeof(double);
    size_t nb_ov = override_h ? (size_t)B * d->n * sizeof(double) : 0;
    size_t nb_st = (size_t)B * sizeof(int);
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t total = al(nb_params) + al(nb_out) + al(nb_ov) + al(nb_st);
    if ((rc = ensure(&ctx->io, &ctx->io_bytes, total))) return rc;
    char *base = (char *)ctx->io;
    double *d_params = (double *)base;
    double *d_out = (double *)(base + al(nb_params));
    double *d_ov = override_h ? (double *)(base + al(nb_params) + al(nb_out)) : nullptr;
    int *d_st = (int *)(base + al(nb_params) + al(nb_out) + al(nb_ov));
    cudaStream_t s = ctx->stream;
    // Small chi-square calls (the rounds of a fit) go through page-locked staging owned by the context: one upload, ONE
    // download of the adjacent chi2 | status block and one wait, instead of three pageable copies that each synchronise.
    constexpr size_t kStageBytes = 1u << 20;
    const size_t nb_back = al(nb_out) + nb_st;
    const bool staged = what == OUT_CHI2 && !override_h && nb_params <= kStageBytes && nb_back <= kStageBytes;
    if (staged && !ctx->stage) {
        if (cudaHostAlloc((void **)&ctx->stage, 2 * kStageBytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); ctx->stage = nullptr; }
    }
    if (staged && ctx->stage) {
        std::memcpy(ctx->stage, params, nb_params);
        BF_CUDA_OK(cudaMemcpyAsync(d_params, ctx->stage, nb_params, cudaMemcpyHostToDevice, s));
        // The kernel chain of a small call is a fixed function of (model, data, B, buffers) once the host has established that
        // the vectors share their key and that the device tables are current for it (KeyHint): such a call is seen once, then
        // captured into a CUDA graph on its second occurrence and replayed from the third on -- one launch instead of six.
        // (Repeated small calls are what a Minuit-style fit makes: B = 1, 2 n + 1, a handful of line-search points.)
        KeyHint hint;
        const bool hinted = host_key_check(m->dev, params, B, &hint, nullptr);
        const bool graphable = B <= bf_ctx::kGraphMaxB && !ctx->profiling && !ctx->no_graphs && m->dev.kind != BF_MODEL_KSPACE_FFT &&
                               d->binning_type != BF_BINNING_MULTIPOLE && (m->dev.kind != BF_MODEL_KSPACE || hinted);
        bf_ctx::GraphEntry *ge = nullptr;
        const bf_ctx::GraphKey key{m->serial, (const void *)d, B, ctx->io, ctx->ws, ctx->ws_bytes, ctx->chi2_part};
        if (graphable) {
            if (ctx->graphs.size() >= 64 && !ctx->graphs.count(key)) purge_graphs(ctx, 0, nullptr);  // bounded cache
            ge = &ctx->graphs[key];
            bool current = true;  // k-space: the tables on the device must be those the chain was captured for
            if (m->dev.kind == BF_MODEL_KSPACE) {
                current = m->h_key_valid && ge->nkey == hint.nkey;
                for (int k = 0; k < hint.nkey && current; ++k) current = hint.vals[k] == ge->hk[k];
                for (int k = 0; k < 32 && current; ++k) current = ge->hk[k] == m->h_key[k];
            }
            if (ge->exec && current) {
                BF_CUDA_OK(cudaGraphLaunch(ge->exec, s));
                ctx->launches += ge->launches;
                goto staged_download;
            }
            if (ge->exec) { cudaGraphExecDestroy(ge->exec); ge->exec = nullptr; ge->seen = 0; }
        }
        {
            // capture only what will run as the fixed chain: for the k-space model the host mirror must already hold this key
            bool mirror_ok = m->dev.kind != BF_MODEL_KSPACE || m->h_key_valid;
            for (int k = 0; k < hint.nkey && mirror_ok && m->dev.kind == BF_MODEL_KSPACE; ++k) mirror_ok = hint.vals[k] == m->h_key[k];
            bool capturing = ge && ge->seen >= 1 && mirror_ok;
            const long long l0 = ctx->launches;
            if (capturing && cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); capturing = false; ge->seen = -1000000; }
            {
                KeyHintScope hs(ctx, m->dev, params, B);
                rc = run_device(ctx, m, d, d_params, B, what, d_out, B, nullptr, d_st);
            }
            if (capturing) {
                cudaGraph_t graph = nullptr;
                const cudaError_t e = cudaStreamEndCapture(s, &graph);
                const bool same_buffers = ctx->io == key.io && ctx->ws == key.ws && ctx->ws_bytes == key.ws_bytes && ctx->chi2_part == key.chi2_part;
                if (rc == BF_OK && e == cudaSuccess && graph && same_buffers && (m->dev.kind != BF_MODEL_KSPACE || ctx->last_host_current) &&
                    cudaGraphInstantiate(&ge->exec, graph, 0) == cudaSuccess) {
                    ge->launches = ctx->launches - l0;
                    ge->nkey = hint.nkey;
                    std::memcpy(ge->hk, ctx->last_hk, sizeof(ge->hk));
                } else {
                    cudaGetLastError();
                    ge->exec = nullptr;
                    ge->seen = -1000000;  // do not try again for this key
                }
                if (graph) cudaGraphDestroy(graph);
                if (rc) return rc;
                if (ge->exec) BF_CUDA_OK(cudaGraphLaunch(ge->exec, s));
                else {  // nothing ran during the capture: run the chain the ordinary way
                    m->h_key_valid = false;  // the host mirror may have been advanced by the pass that did not execute
                    KeyHintScope hs(ctx, m->dev, params, B);
                    if ((rc = run_device(ctx, m, d, d_params, B, what, d_out, B, nullptr, d_st))) return rc;
                }
            } else {
                if (rc) return rc;
                if (ge) ge->seen++;
            }
        }
    staged_download:
        char *back = ctx->stage + kStageBytes;
        BF_CUDA_OK(cudaMemcpyAsync(back, d_out, nb_back, cudaMemcpyDeviceToHost, s));
        BF_CUDA_OK(cudaStreamSynchronize(s));
        std::memcpy(out, back, nb_out);
        if (status) std::memcpy(status, back + al(nb_out), nb_st);
        return BF_OK;
    }
    BF_CUDA_OK(cudaMemcpyAsync(d_params, params, nb_params, cudaMemcpyHostToDevice, s));
    if (override_h) BF_CUDA_OK(cudaMemcpyAsync(d_ov, override_h, nb_ov, cudaMemcpyHostToDevice, s));
    {
        KeyHintScope hs(ctx, m->dev, params, B);
        rc = run_device(ctx, m, d, d_params, B, what, d_out, B, d_ov, d_st);
    }
    if (rc) return rc;
    if (what == OUT_CHI2)
        BF_CUDA_OK(cudaMemcpyAsync(out, d_out, nb_out, cudaMemcpyDeviceToHost, s));
    else
        BF_CUDA_OK(cudaMemcpy2DAsync(out, (size_t)ldo * sizeof(double), d_out, (size_t)B * sizeof(double),
                                     (size_t)B * sizeof(double), out_rows, cudaMemcpyDeviceToHost, s));
    std::vector<int> st_local;
    int *st = status;
    if (!st && what != OUT_CHI2) { st_local.resize(B); st = st_local.data(); }
    if (st) BF_CUDA_OK(cudaMemcpyAsync(st, d_st, nb_st, cudaMemcpyDeviceToHost, s));
    BF_CUDA_OK(cudaStreamSynchronize(s));
    if (what != OUT_CHI2 && st)
        for (int b = 0; b < B; ++b)
            if (st[b])
                for (int i = 0; i < out_rows; ++i) out[(size_t)i * ldo + b] = NAN;
    return BF_OK;
}

extern "C" int bf_eval_batch(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *params, int B,
                             double *pred, int ldp, int *status) {
    if (!pred || ldp < B) BF_FAIL(BF_ERR_OPTIONS, "bf_eval_batch: bad output buffer");
    return run_host(ctx, m, d, params, B, OUT_PRED, pred, ldp, d ? d->n : 0, nullptr, status);
}

extern "C" int bf_chi2_batch(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *params, int B,
                             const double *data_override, double *chi2, int *status) {
    if (!chi2) BF_FAIL(BF_ERR_OPTIONS, "bf_chi2_batch: null output");
    return run_host(ctx, m, d, params, B, OUT_CHI2, chi2, 0, 0, data_override, status);
}

// Asynchronous host-buffer chi-square: the upload of call n + 1 crosses the bus on the copy stream while call n is
// being evaluated. Two staging slots; a ticket names the slot and its generation.
extern "C" int bf_chi2_batch_submit(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *params, int B,
                                    double *chi2, int *status, int *ticket) {
    if (!chi2 || !ticket) BF_FAIL(BF_ERR_OPTIONS, "bf_chi2_batch_submit: null output");
    int rc = check_args(ctx, m, d, params, B);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    const int slot = (int)(ctx->submit_serial & 1);
    bf_ctx::AsyncSlot &sl = ctx->async_slot[slot];
    if (sl.pending) BF_FAIL(BF_ERR_OPTIONS, "bf_chi2_batch_submit: two calls are already in flight, wait for the older one first");
    const int npar = m->dev.npar;
    const size_t nb_params = (size_t)B * npar * sizeof(double), nb_out = (size_t)B * sizeof(double), nb_st = (size_t)B * sizeof(int);
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    if ((rc = ensure(&sl.buf, &sl.bytes, al(nb_params) + al(nb_out) + al(nb_st)))) return rc;
    double *d_params = (double *)sl.buf;
    double *d_out = (double *)((char *)sl.buf + al(nb_params));
    int *d_st = (int *)((char *)sl.buf + al(nb_params) + al(nb_out));
    BF_CUDA_OK(cudaMemcpyAsync(d_params, params, nb_params, cudaMemcpyHostToDevice, ctx->copy_stream));
    BF_CUDA_OK(cudaEventRecord(sl.uploaded, ctx->copy_stream));
    BF_CUDA_OK(cudaStreamWaitEvent(ctx->stream, sl.uploaded, 0));
    {
        KeyHintScope hs(ctx, m->dev, params, B);
        rc = run_device(ctx, m, d, d_params, B, OUT_CHI2, d_out, B, nullptr, d_st);
    }
    if (rc) return rc;
    BF_CUDA_OK(cudaMemcpyAsync(chi2, d_out, nb_out, cudaMemcpyDeviceToHost, ctx->stream));
    if (status) BF_CUDA_OK(cudaMemcpyAsync(status, d_st, nb_st, cudaMemcpyDeviceToHost, ctx->stream));
    BF_CUDA_OK(cudaEventRecord(sl.done, ctx->stream));
    sl.pending = true;
    sl.serial = ctx->submit_serial;
    *ticket = (int)(ctx->submit_serial & 0x7fffffff);
    ctx->submit_serial++;
    return BF_OK;
}

extern "C" int bf_chi2_batch_wait(bf_ctx *ctx, int ticket) {
    if (!ctx) BF_FAIL(BF_ERR_OPTIONS, "bf_chi2_batch_wait: null context");
    cudaEvent_t ev;
    {
        std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
        bf_ctx::AsyncSlot &sl = ctx->async_slot[ticket & 1];
        if (!sl.pending || (int)(sl.serial & 0x7fffffff) != ticket) BF_FAIL(BF_ERR_OPTIONS, "bf_chi2_batch_wait: unknown ticket");
        ev = sl.done;
    }
    BF_CUDA_OK(cudaEventSynchronize(ev));  // outside the lock: another thread may submit meanwhile
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    ctx->async_slot[ticket & 1].pending = false;
    return BF_OK;
}

extern "C" int bf_eval_undistorted_batch(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *params,
                                         int B, double *xiu, int ldp, int *status) {
    if (!xiu || ldp < B) BF_FAIL(BF_ERR_OPTIONS, "bf_eval_undistorted_batch: bad output buffer");
    if (!m || m->dev.dm_order <= 0) BF_FAIL(BF_ERR_ANALYSIS, "bf_eval_undistorted_batch: the model has no distortion matrix");
    return run_host(ctx, m, d, params, B, OUT_XIU, xiu, ldp, d ? d->n_grid : 0, nullptr, status);
}

extern "C" int bf_transform_batch(bf_ctx *ctx, const bf_model *m, const double *params, int B, double z_trigger,
                                  double *out, int ldb_out) {
    if (!ctx || !m || !params || !out) BF_FAIL(BF_ERR_OPTIONS, "null argument");
    if (m->dev.kind != BF_MODEL_KSPACE) BF_FAIL(BF_ERR_ANALYSIS, "bf_transform_batch: not a k-space model");
    if (B <= 0 || ldb_out < B) BF_FAIL(BF_ERR_OPTIONS, "bf_transform_batch: bad batch size / leading dimension");
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    const int npar = m->dev.npar, nr = m->dev.nr;
    int cap = chunk_capacity(m, nullptr, 1);
    size_t nb_params = (size_t)std::min(B, cap) * npar * sizeof(double);
    size_t nb_out = (size_t)6 * nr * std::min(B, cap) * sizeof(double);
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    int rc = ensure(&ctx->io, &ctx->io_bytes, al(nb_params) + al(nb_out));
    if (rc) return rc;
    double *d_params = (double *)ctx->io;
    double *d_out = (double *)((char *)ctx->io + al(nb_params));
    for (int b0 = 0; b0 < B; b0 += cap) {
        int bc = std::min(cap, B - b0);
        BF_CUDA_OK(cudaMemcpyAsync(d_params, params + (size_t)b0 * npar, (size_t)bc * npar * sizeof(double),
                                   cudaMemcpyHostToDevice, ctx->stream));
        rc = run_chunk(ctx, m, nullptr, nullptr, d_params, bc, OUT_TRANSFORM, d_out, bc, nullptr, nullptr, z_trigger, true);
        if (rc) return rc;
        BF_CUDA_OK(cudaMemcpy2DAsync(out + b0, (size_t)ldb_out * sizeof(double), d_out, (size_t)bc * sizeof(double),
                                     (size_t)bc * sizeof(double), (size_t)6 * nr, cudaMemcpyDeviceToHost, ctx->stream));
        BF_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    }
    return BF_OK;
}

extern "C" int bf_kspace_transform_weights(const bf_model_desc *d, int *nr, int *nk, double *weights) {
    if (!d || !nr || !nk) BF_FAIL(BF_ERR_OPTIONS, "bf_kspace_transform_weights: null argument");
    if (d->kind != BF_MODEL_KSPACE || d->n_pk < 8 || !d->pk_k || !d->pk_fid || !d->pk_nw || d->samples_per_decade < 1 ||
        !(d->rmin > 0) || !(d->rmin < d->rmax) || !(d->dilmin > 0) || d->dilmin > d->dilmax)
        BF_FAIL(BF_ERR_MODEL, "bf_kspace_transform_weights: not a valid k-space model description");
    HankelSetup hs;
    build_hankel_weights(d->n_pk, d->pk_k, d->pk_fid, d->pk_nw, d->rmin, d->rmax, d->dilmin, d->dilmax, d->samples_per_decade, kFineSub, hs);
    *nr = hs.nr;
    *nk = hs.nk;
    if (weights) std::memcpy(weights, hs.weights.data(), hs.weights.size() * sizeof(double));
    return BF_OK;
}

extern "C" int bf_model_fft_plane_dims(const bf_model *m, int *npx, int *npy) {
    if (!m || m->dev.kind != BF_MODEL_KSPACE_FFT) BF_FAIL(BF_ERR_OPTIONS, "bf_model_fft_plane_dims: not an FFT model");
    if (npx) *npx = m->dev.fft.npx;
    if (npy) *npy = m->dev.fft.npy;
    return BF_OK;
}

extern "C" int bf_fft_planes_batch(bf_ctx *ctx, const bf_model *m, const double *params, int B, double z_trigger,
                                   double *out) {
    if (!ctx || !m || !params || !out) BF_FAIL(BF_ERR_OPTIONS, "null argument");
    if (m->dev.kind != BF_MODEL_KSPACE_FFT) BF_FAIL(BF_ERR_ANALYSIS, "bf_fft_planes_batch: not an FFT model");
    if (B <= 0) BF_FAIL(BF_ERR_OPTIONS, "bf_fft_planes_batch: bad batch size");
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    const DevFft &f = m->dev.fft;
    const int npar = m->dev.npar;
    const size_t per_b = (size_t)2 * f.npy_pad * f.npx;  // rows are compacted to npx, all npy_pad rows copied
    int cap = std::min(chunk_capacity(m, nullptr, 1), 4096);
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t nb_params = (size_t)std::min(B, cap) * npar * sizeof(double);
    size_t nb_out = per_b * std::min(B, cap) * sizeof(double);
    int rc = ensure(&ctx->io, &ctx->io_bytes, al(nb_params) + al(nb_out));
    if (rc) return rc;
    double *d_params = (double *)ctx->io;
    double *d_out = (double *)((char *)ctx->io + al(nb_params));
    for (int b0 = 0; b0 < B; b0 += cap) {
        int bc = std::min(cap, B - b0);
        BF_CUDA_OK(cudaMemcpyAsync(d_params, params + (size_t)b0 * npar, (size_t)bc * npar * sizeof(double),
                                   cudaMemcpyHostToDevice, ctx->stream));
        rc = run_chunk(ctx, m, nullptr, nullptr, d_params, bc, OUT_TRANSFORM, d_out, 0, nullptr, nullptr, z_trigger, true);
        if (rc) return rc;
        // drop the padded rows: out[((b*2+comp)*npy + iy)*npx + ix]
        BF_CUDA_OK(cudaMemcpy2DAsync(out + (size_t)b0 * 2 * f.npy * f.npx, (size_t)f.npy * f.npx * sizeof(double), d_out,
                                     (size_t)f.npy_pad * f.npx * sizeof(double), (size_t)f.npy * f.npx * sizeof(double),
                                     (size_t)bc * 2, cudaMemcpyDeviceToHost, ctx->stream));
        BF_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    }
    return BF_OK;
}

extern "C" int bf_sample_toys(bf_ctx *ctx, const bf_data *d, const double *truth, unsigned long long seed,
                              long long first_toy, int ntoy, double scale, double *out) {
    if (!ctx || !d || !truth || !out || ntoy <= 0) BF_FAIL(BF_ERR_OPTIONS, "bf_sample_toys: bad argument");
    if (!d->d_L) BF_FAIL(BF_ERR_ANALYSIS, "bf_sample_toys: needs a covariance (not an inverse covariance)");
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t nb_truth = (size_t)d->n * sizeof(double), nb_out = (size_t)ntoy * d->n * sizeof(double);
    int rc = ensure(&ctx->io, &ctx->io_bytes, al(nb_truth) + al(nb_out));
    if (rc) return rc;
    double *d_truth = (double *)ctx->io, *d_out = (double *)((char *)ctx->io + al(nb_truth));
    if ((rc = ensure(&ctx->ws, &ctx->ws_bytes, toy_noise_work_doubles(d->n_pad, ntoy) * sizeof(double)))) return rc;
    BF_CUDA_OK(cudaMemcpyAsync(d_truth, truth, nb_truth, cudaMemcpyHostToDevice, ctx->stream));
    BF_KERNEL(ctx, "toy_noise[deviates + dgemm_store_kernel + transpose]",
              launch_toy_noise(ctx->stream, d->d_L, d->n, d->n_pad, d_truth, seed, first_toy, ntoy, scale, d_out, (double *)ctx->ws));
    ctx->launches += 3 * ((ntoy + kToyPassCols - 1) / kToyPassCols) - 1;
    BF_CUDA_OK(cudaMemcpyAsync(out, d_out, nb_out, cudaMemcpyDeviceToHost, ctx->stream));
    BF_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return BF_OK;
}

// ------------------------------------------------------------------------------------------
// device-resident toy tables
// ------------------------------------------------------------------------------------------
extern "C" int bf_toys_create(bf_ctx *ctx, const bf_data *d, const double *truth, unsigned long long seed,
                              long long first_toy, int ntoy, double scale, bf_toys **out) {
    if (!ctx || !d || !truth || !out || ntoy <= 0) BF_FAIL(BF_ERR_OPTIONS, "bf_toys_create: bad argument");
    if (!d->d_L) BF_FAIL(BF_ERR_ANALYSIS, "bf_toys_create: needs a covariance (not an inverse covariance)");
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    if (int rc = ensure(&ctx->ws, &ctx->ws_bytes, toy_noise_work_doubles(d->n_pad, ntoy) * sizeof(double))) return rc;
    bf_toys *t = new bf_toys();
    t->ctx = ctx; t->ntoy = ntoy; t->n = d->n;
    double *d_truth = nullptr;
    // stream-ordered allocations from the device's pool (its release threshold is raised in bf_ctx_create): a toy
    // study is a 30 ms burst, a cudaMalloc / cudaFree pair per study would be a visible part of it
    cudaError_t e = cudaMallocAsync((void **)&t->d_table, (size_t)ntoy * d->n * sizeof(double), ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_truth, (size_t)d->n * sizeof(double), ctx->stream);
    if (e != cudaSuccess) { if (t->d_table) cudaFreeAsync(t->d_table, ctx->stream); delete t; BF_FAIL(BF_ERR_ANALYSIS, std::string("bf_toys_create: ") + cudaGetErrorString(e)); }
    cudaMemcpyAsync(d_truth, truth, (size_t)d->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    BF_KERNEL(ctx, "toy_noise[deviates + dgemm_store_kernel + transpose]",
              launch_toy_noise(ctx->stream, d->d_L, d->n, d->n_pad, d_truth, seed, first_toy, ntoy, scale, t->d_table, (double *)ctx->ws));
    ctx->launches += 3 * ((ntoy + kToyPassCols - 1) / kToyPassCols) - 1;
    cudaFreeAsync(d_truth, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFreeAsync(t->d_table, ctx->stream); delete t; BF_FAIL(BF_ERR_ANALYSIS, std::string("bf_toys_create: ") + cudaGetErrorString(e)); }
    *out = t;
    return BF_OK;
}

extern "C" int bf_toys_destroy(bf_toys *t) {
    if (!t) return BF_OK;
    std::lock_guard<std::recursive_mutex> lock(t->ctx->api_mu);
    cudaSetDevice(t->ctx->device);
    cudaFreeAsync(t->d_table, t->ctx->stream);
    delete t;
    return BF_OK;
}

extern "C" int bf_toys_download(bf_ctx *ctx, const bf_toys *t, double *out) {
    if (!ctx || !t || !out) BF_FAIL(BF_ERR_OPTIONS, "bf_toys_download: null argument");
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    BF_CUDA_OK(cudaMemcpyAsync(out, t->d_table, (size_t)t->ntoy * t->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    BF_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return BF_OK;
}

extern "C" int bf_chi2_batch_toys(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *params, int B,
                                  const bf_toys *toys, const int *toy_index, double *chi2, int *status) {
    if (!chi2 || !toys || !toy_index) BF_FAIL(BF_ERR_OPTIONS, "bf_chi2_batch_toys: null argument");
    int rc = check_args(ctx, m, d, params, B);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    if (toys->ctx != ctx || toys->n != d->n) BF_FAIL(BF_ERR_OPTIONS, "bf_chi2_batch_toys: the toy table belongs to another context / data set");
    for (int b = 0; b < B; ++b)
        if (toy_index[b] < 0 || toy_index[b] >= toys->ntoy) BF_FAIL(BF_ERR_OPTIONS, "bf_chi2_batch_toys: toy index out of range");
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    const int npar = m->dev.npar, n = d->n;
    cudaStream_t s = ctx->stream;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    // passes of at most 16384 points bound the gathered override panel (B x n doubles)
    const int cap = 16384;
    const int pass = std::min(B, cap);
    size_t nb_params = al((size_t)pass * npar * sizeof(double)), nb_idx = al((size_t)pass * sizeof(int)), nb_chi2 = al((size_t)pass * sizeof(double)),
           nb_st = al((size_t)pass * sizeof(int)), nb_ov = al((size_t)pass * n * sizeof(double));
    if ((rc = ensure(&ctx->io, &ctx->io_bytes, nb_params + nb_idx + nb_chi2 + nb_st + nb_ov))) return rc;
    char *base = (char *)ctx->io;
    double *d_params = (double *)base;
    int *d_idx = (int *)(base + nb_params);
    double *d_chi2 = (double *)(base + nb_params + nb_idx);
    int *d_st = (int *)(base + nb_params + nb_idx + nb_chi2);
    double *d_ov = (double *)(base + nb_params + nb_idx + nb_chi2 + nb_st);
    for (int b0 = 0; b0 < B; b0 += cap) {
        const int bc = std::min(cap, B - b0);
        BF_CUDA_OK(cudaMemcpyAsync(d_params, params + (size_t)b0 * npar, (size_t)bc * npar * sizeof(double), cudaMemcpyHostToDevice, s));
        BF_CUDA_OK(cudaMemcpyAsync(d_idx, toy_index + b0, (size_t)bc * sizeof(int), cudaMemcpyHostToDevice, s));
        BF_KERNEL(ctx, "gather_rows_kernel", launch_gather_rows(s, toys->d_table, d_idx, bc, n, d_ov));
        {
            KeyHintScope hs(ctx, m->dev, params + (size_t)b0 * npar, bc);
            rc = run_device(ctx, m, d, d_params, bc, OUT_CHI2, d_chi2, 0, d_ov, d_st);
        }
        if (rc) return rc;
        BF_CUDA_OK(cudaMemcpyAsync(chi2 + b0, d_chi2, (size_t)bc * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (status) BF_CUDA_OK(cudaMemcpyAsync(status + b0, d_st, (size_t)bc * sizeof(int), cudaMemcpyDeviceToHost, s));
        BF_CUDA_OK(cudaStreamSynchronize(s));
    }
    return BF_OK;
}

// ------------------------------------------------------------------------------------------
// Gauss-Newton normal equations
// ------------------------------------------------------------------------------------------
extern "C" int bf_gram_batch(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *params, int nfit, int G,
                             const bf_toys *toys, const int *toy_index, double *gram, int *status) {
    if (!gram || nfit <= 0 || G < 1 || G > 96) BF_FAIL(BF_ERR_OPTIONS, "bf_gram_batch: bad argument (1 <= G <= 96)");
    int rc = check_args(ctx, m, d, params, nfit * G);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    if ((toys == nullptr) != (toy_index == nullptr)) BF_FAIL(BF_ERR_OPTIONS, "bf_gram_batch: toys and toy_index go together");
    if (toys) {
        if (toys->ctx != ctx || toys->n != d->n) BF_FAIL(BF_ERR_OPTIONS, "bf_gram_batch: the toy table belongs to another context / data set");
        for (int f = 0; f < nfit; ++f)
            if (toy_index[f] < 0 || toy_index[f] >= toys->ntoy) BF_FAIL(BF_ERR_OPTIONS, "bf_gram_batch: toy index out of range");
    }
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    const int npar = m->dev.npar, n = d->n;
    cudaStream_t s = ctx->stream;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    // passes of whole fits, at most ~16384 columns each
    const int fits_per_pass = std::max(1, std::min(nfit, 65536 / G));
    const int colcap = fits_per_pass * G, ldy = round_up(colcap, kPadN);
    size_t nb_params = al((size_t)colcap * npar * sizeof(double)), nb_idx = al((size_t)colcap * sizeof(int)),
           nb_st = al((size_t)colcap * sizeof(int)), nb_gram = al((size_t)fits_per_pass * G * G * sizeof(double)),
           nb_ov = toys ? al((size_t)colcap * n * sizeof(double)) : 0, nb_y = al((size_t)d->n_pad * ldy * sizeof(double));
    if ((rc = ensure(&ctx->io, &ctx->io_bytes, nb_params + nb_idx + nb_st + nb_gram + nb_ov + nb_y))) return rc;
    char *base = (char *)ctx->io;
    double *d_params = (double *)base;
    int *d_idx = (int *)(base + nb_params);
    int *d_st = (int *)(base + nb_params + nb_idx);
    double *d_gram = (double *)(base + nb_params + nb_idx + nb_st);
    double *d_ov = toys ? (double *)(base + nb_params + nb_idx + nb_st + nb_gram) : nullptr;
    double *d_y = (double *)(base + nb_params + nb_idx + nb_st + nb_gram + nb_ov);
    std::vector<int> colidx;
    for (int f0 = 0; f0 < nfit; f0 += fits_per_pass) {
        const int nf = std::min(fits_per_pass, nfit - f0), bc = nf * G;
        BF_CUDA_OK(cudaMemcpyAsync(d_params, params + (size_t)f0 * G * npar, (size_t)bc * npar * sizeof(double), cudaMemcpyHostToDevice, s));
        if (toys) {
            colidx.resize(bc);
            for (int c = 0; c < bc; ++c) colidx[c] = toy_index[f0 + c / G];
            BF_CUDA_OK(cudaMemcpyAsync(d_idx, colidx.data(), (size_t)bc * sizeof(int), cudaMemcpyHostToDevice, s));
            BF_KERNEL(ctx, "gather_rows_kernel", launch_gather_rows(s, toys->d_table, d_idx, bc, n, d_ov));
        }
        {
            KeyHintScope hs(ctx, m->dev, params + (size_t)f0 * G * npar, bc);
            rc = run_device(ctx, m, d, d_params, bc, OUT_WHITENED, d_y, ldy, d_ov, d_st);
        }
        if (rc) return rc;
        BF_KERNEL(ctx, "gram_kernel", launch_gram(s, d_y, ldy, n, G, nf, d_gram));
        BF_CUDA_OK(cudaMemcpyAsync(gram + (size_t)f0 * G * G, d_gram, (size_t)nf * G * G * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (status) BF_CUDA_OK(cudaMemcpyAsync(status + (size_t)f0 * G, d_st, (size_t)bc * sizeof(int), cudaMemcpyDeviceToHost, s));
        BF_CUDA_OK(cudaStreamSynchronize(s));  // colidx is reused by the next pass
    }
    BF_CUDA_OK(cudaGetLastError());
    return BF_OK;
}

static int gram_probes(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *centres, const double *steps, int nfit,
                       int nprobe, const bf_toys *toys, const int *toy_index, double *gram, int *status, bool normal);

extern "C" int bf_gram_probes(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *centres, const double *steps,
                              int nfit, int nprobe, const bf_toys *toys, const int *toy_index, double *gram, int *status) {
    return gram_probes(ctx, m, d, centres, steps, nfit, nprobe, toys, toy_index, gram, status, false);
}

extern "C" int bf_normal_probes(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *centres, const double *steps,
                                int nfit, int nprobe, const bf_toys *toys, const int *toy_index, double *normal, int *status) {
    return gram_probes(ctx, m, d, centres, steps, nfit, nprobe, toys, toy_index, normal, status, true);
}

extern "C" int bf_pinned_alloc(size_t bytes, void **ptr) {
    if (!ptr) BF_FAIL(BF_ERR_OPTIONS, "bf_pinned_alloc: null argument");
    BF_CUDA_OK(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return BF_OK;
}

extern "C" int bf_pinned_free(void *ptr) {
    // may run from a thread-local destructor at process exit, after the CUDA runtime has gone: errors are ignored
    if (ptr && cudaFreeHost(ptr) != cudaSuccess) cudaGetLastError();
    return BF_OK;
}

static int gram_probes(bf_ctx *ctx, const bf_model *m, const bf_data *d, const double *centres, const double *steps, int nfit,
                       int nprobe, const bf_toys *toys, const int *toy_index, double *gram, int *status, bool normal) {
    const int G = 2 * nprobe + 1;
    const size_t per_fit = normal ? (size_t)(nprobe + 1) * (nprobe + 1) + nprobe : (size_t)G * G;
    if (!gram || !steps || nfit <= 0 || nprobe < 0 || G > 96) BF_FAIL(BF_ERR_OPTIONS, "bf_gram_probes: bad argument (2 nprobe + 1 <= 96)");
    int rc = check_args(ctx, m, d, centres, nfit);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lock(ctx->api_mu);
    if ((toys == nullptr) != (toy_index == nullptr)) BF_FAIL(BF_ERR_OPTIONS, "bf_gram_probes: toys and toy_index go together");
    const int npar = m->dev.npar, n = d->n;
    for (int f = 0; f < nfit; ++f) {
        int cnt = 0;
        for (int j = 0; j < npar; ++j) cnt += steps[(size_t)f * npar + j] != 0.0;
        if (cnt != nprobe) BF_FAIL(BF_ERR_OPTIONS, "bf_gram_probes: every fit must probe exactly nprobe parameters");
        if (toys && (toy_index[f] < 0 || toy_index[f] >= toys->ntoy)) BF_FAIL(BF_ERR_OPTIONS, "bf_gram_probes: toy index out of range");
    }
    if (toys && (toys->ctx != ctx || toys->n != n)) BF_FAIL(BF_ERR_OPTIONS, "bf_gram_probes: the toy table belongs to another context / data set");
    BF_CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const int fits_per_pass = std::max(1, std::min(nfit, 65536 / G));
    const int colcap = fits_per_pass * G, ldy = round_up(colcap, kPadN);
    size_t nb_cs = al((size_t)fits_per_pass * npar * sizeof(double)), nb_params = al((size_t)colcap * npar * sizeof(double)),
           nb_idx = al((size_t)colcap * sizeof(int)), nb_st = al((size_t)colcap * sizeof(int)), nb_sf = al((size_t)fits_per_pass * sizeof(int)),
           nb_gram = al((size_t)fits_per_pass * per_fit * sizeof(double)), nb_ov = toys ? al((size_t)colcap * n * sizeof(double)) : 0,
           nb_y = al((size_t)d->n_pad * ldy * sizeof(double)), nb_part = normal ? al(normal_probes_part_doubles(nprobe, fits_per_pass) * sizeof(double)) : 0;
    if ((rc = ensure(&ctx->io, &ctx->io_bytes, 2 * nb_cs + nb_params + nb_idx + nb_st + nb_sf + nb_gram + nb_ov + nb_y + nb_part))) return rc;
    char *p = (char *)ctx->io;
    auto take = [&](size_t bytes) { char *q = p; p += bytes; return q; };
    double *d_c = (double *)take(nb_cs), *d_h = (double *)take(nb_cs), *d_params = (double *)take(nb_params);
    int *d_idx = (int *)take(nb_idx), *d_st = (int *)take(nb_st), *d_sf = (int *)take(nb_sf);
    double *d_gram = (double *)take(nb_gram), *d_ov = toys ? (double *)take(nb_ov) : nullptr, *d_y = (double *)take(nb_y);
    double *d_part = normal ? (double *)take(nb_part) : nullptr;
    // Probed parameters that enter the whitened residual linearly with a bin-static basis -- the additive broadband
    // coefficients of a model without velocity shift, the structure the fused data tail is built on -- need no
    // evaluations: their difference columns are synthesised from W . basis (normal_probes_lin_kernel), exactly.
    // BOSSDR11LyaF_k.ini floats 2 non-linear and 9 such parameters per grid fit: 5 evaluations per fit instead of 23.
    std::vector<int> lin(npar, -1);
    int n_nl = nprobe;
    const bf_data::Plan *plan = nullptr;
    const bool no_lin = std::getenv("BAOFIT_NO_LINEAR_PROBES") != nullptr;  // A/B tests
    if (normal && !no_lin && d->binning_type != BF_BINNING_MULTIPOLE && m->dev.bb[BF_BB_DIST_ADD].enabled) {
        if ((rc = get_plan(m, d, &plan))) return rc;
        const DevBB &ba = m->dev.bb[BF_BB_DIST_ADD];
        const int n_rp = (ba.rp_max - ba.rp_min) / ba.rp_step + 1;
        if (plan->fold_ok && plan->fold_data_mode && plan->fold_ncls <= 16 &&
            (plan->fold_case == 1 || (plan->fold_case == 2 && plan->fold_ncls == 1 && plan->fold_pmax < 16))) {
            // case 1: one static column per coefficient (and redshift class); case 2 (velocity shift): the coefficient
            // (z, p, t) is a combination of the static columns (z, m <= p, t) with weights C(p, m) v^(p - m) of the fit
            const int ncoef = plan->fold_case == 1 ? plan->fold_nb / plan->fold_ncls : plan->fold_nz * n_rp * plan->fold_nrt;
            const int base = ba.idx_base;
            for (int k = 0; k < ncoef; ++k) lin[base + k] = k;
            for (int f = 0; f < nfit; ++f) {
                int cnt = 0;
                for (int j = 0; j < npar; ++j) cnt += steps[(size_t)f * npar + j] != 0.0 && lin[j] < 0;
                if (f == 0) n_nl = cnt;
                else if (cnt != n_nl) { n_nl = nprobe; break; }  // fits probe different numbers of non-linear parameters
            }
        }
    }
    const bool use_lin = n_nl < nprobe;
    const int Ge = use_lin ? 2 * n_nl + 1 : G;  // evaluated columns per fit
    int *d_lin = nullptr;
    if (use_lin) {
        if ((rc = ensure(&ctx->aux, &ctx->aux_bytes, al((size_t)npar * sizeof(int))))) return rc;
        d_lin = (int *)ctx->aux;
        // (pageable source: the call returns once the bytes are in the driver's staging buffer)
        BF_CUDA_OK(cudaMemcpyAsync(d_lin, lin.data(), (size_t)npar * sizeof(int), cudaMemcpyHostToDevice, s));
    }
    std::vector<int> colidx;
    for (int f0 = 0; f0 < nfit; f0 += fits_per_pass) {
        const int nf = std::min(fits_per_pass, nfit - f0), bc = nf * Ge;
        BF_CUDA_OK(cudaMemcpyAsync(d_c, centres + (size_t)f0 * npar, (size_t)nf * npar * sizeof(double), cudaMemcpyHostToDevice, s));
        BF_CUDA_OK(cudaMemcpyAsync(d_h, steps + (size_t)f0 * npar, (size_t)nf * npar * sizeof(double), cudaMemcpyHostToDevice, s));
        BF_KERNEL(ctx, "expand_probes_kernel", launch_expand_probes(s, d_c, d_h, nf, npar, use_lin ? n_nl : nprobe, d_params, d_lin));
        if (toys) {
            colidx.resize(bc);
            for (int c = 0; c < bc; ++c) colidx[c] = toy_index[f0 + c / Ge];
            BF_CUDA_OK(cudaMemcpyAsync(d_idx, colidx.data(), (size_t)bc * sizeof(int), cudaMemcpyHostToDevice, s));
            BF_KERNEL(ctx, "gather_rows_kernel", launch_gather_rows(s, toys->d_table, d_idx, bc, n, d_ov));
        }
        {
            KeyHintScope hs(ctx, m->dev, centres + (size_t)f0 * npar, nf, steps + (size_t)f0 * npar);
            rc = run_device(ctx, m, d, d_params, bc, OUT_WHITENED, d_y, ldy, d_ov, d_st);
        }
        if (rc) return rc;
        if (use_lin) {
            NormalLinArgs na;
            std::memset(&na, 0, sizeof(na));
            na.Y = d_y; na.ldy = ldy; na.nrows = n; na.n = nprobe; na.n_nl = n_nl; na.nfit = nf; na.npar = npar;
            na.centres = d_c; na.steps = d_h; na.lin = d_lin;
            na.WA = plan->d_WA; na.lda = d->n_pad + plan->fold_kext; na.kleft = d->n_pad;
            na.nb1 = plan->fold_nb / plan->fold_ncls; na.ncls = plan->fold_ncls;
            {
                const DevBB &ba = m->dev.bb[BF_BB_DIST_ADD];
                na.fold_case = plan->fold_case; na.pmax = plan->fold_pmax; na.nrt = plan->fold_nrt;
                na.n_rp = (ba.rp_max - ba.rp_min) / ba.rp_step + 1; na.rp_min = ba.rp_min; na.rp_step = ba.rp_step;
                na.idx_delta_v = m->dev.idx_delta_v; na.inv_r0 = ba.inv_r0; na.vfac = plan->d_vfac;
            }
            na.zvals = plan->d_zvals; na.zc0 = plan->zc_trigger; na.zref = m->dev.zref; na.idx_gamma_bias = m->dev.idx_gamma_bias;
            na.out = d_gram; na.part = d_part;
            BF_KERNEL(ctx, "normal_probes_lin_kernel", launch_normal_probes_lin(s, na));
        } else if (normal) BF_KERNEL(ctx, "normal_probes_kernel", launch_normal_probes(s, d_y, ldy, n, nprobe, nf, d_gram, d_part));
        else BF_KERNEL(ctx, "gram_kernel", launch_gram(s, d_y, ldy, n, G, nf, d_gram));
        BF_CUDA_OK(cudaMemcpyAsync(gram + (size_t)f0 * per_fit, d_gram, (size_t)nf * per_fit * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (status) {
            BF_KERNEL(ctx, "reduce_status_kernel", launch_reduce_status(s, d_st, nf, Ge, d_sf));
            BF_CUDA_OK(cudaMemcpyAsync(status + f0, d_sf, (size_t)nf * sizeof(int), cudaMemcpyDeviceToHost, s));
        }
        BF_CUDA_OK(cudaStreamSynchronize(s));
    }
    BF_CUDA_OK(cudaGetLastError());
    return BF_OK;
}
