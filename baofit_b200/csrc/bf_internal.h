// Internal structures of libbaofit_b200.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/baofit_b200.h"

namespace bf {

void set_error(const std::string &msg);

#define BF_CUDA_OK(expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            bf::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                   \
            return BF_ERR_ANALYSIS;                                                              \
        }                                                                                        \
    } while (0)

// ---- tile geometry of the DMMA GEMM kernels (kernels_gemm.cu) ----
constexpr int kGemmBM = 64;   // rows of C per CTA tile
constexpr int kGemmBN = 64;   // columns (parameter vectors) per CTA tile
constexpr int kGemmBK = 16;   // k depth of one pipeline stage
constexpr int kPadN = 64;     // batch (b) leading dimensions are multiples of this
constexpr int kPadM = 64;     // row counts of GEMM operands are padded to this
constexpr int kMuNodes = 20;  // Gauss-Legendre nodes in mu for the multipoles of D(k,mu)
constexpr int kFineSub = 4;   // fine k nodes per native P(k) interval

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Exponent run lo, lo+step, ..., <= hi of one broadband axis, split for Horner evaluation: the first nneg
// exponents are <= 0 (powers of x, the last of them is -elast), the remaining n - nneg are > 0 (powers of
// x - 1, the first of them is efirst); baofit/BroadbandModel.cc:209-213.
struct AxisPlan {
    int n, nneg, elast, efirst, step;
};
inline AxisPlan axis_plan(int lo, int hi, int step) {
    AxisPlan p{0, 0, 0, 0, step};
    if (step <= 0 || hi < lo) return p;
    p.n = (hi - lo) / step + 1;
    for (int k = 0; k < p.n; ++k)
        if (lo + k * step <= 0) p.nneg = k + 1;
    if (p.nneg > 0) p.elast = -(lo + (p.nneg - 1) * step);
    if (p.nneg < p.n) p.efirst = lo + p.nneg * step;
    return p;
}

struct DevBB {
    int enabled, idx_base;
    AxisPlan plan_r, plan_rp, plan_rt;
    int r_min, r_max, r_step, r_denom;
    int mu_min, mu_max, mu_step;
    int rp_min, rp_max, rp_step;
    int rt_min, rt_max, rt_step;
    int z_min, z_max, z_step;
    double inv_r0, z0, inv_1pz0;
};

// One tabulated multipole family on shared knots: per-interval cubic coefficients
// coef[(l*(n-1) + j)*4 + {0,1,2,3}] = y_j, b_j, c_j, d_j  (gsl cspline evaluation form)
struct DevSplineSet {
    int n;          // knots
    int uniform;    // 1: x_j = x0 + j*dx
    double x0, inv_dx, xlast;
    const double *x;     // knots (device)
    const double *coef;  // [3][n-1][4] (device)
    double ylo[3], yhi[3];
};

// 3-D FFT k-space model (BaoKSpaceFftCorrelationModel.cc): grid geometry and static tables
struct DevFft {
    int nx, ny, nz;        // cells per axis (y = line of sight)
    int mx, my, mz;        // non-negative frequencies per axis, n/2 + 1
    int my_pad;            // K of the plane GEMM (multiple of 16)
    int npx, npy;          // extent of the kept (r_perp, r_par) plane
    int npx_pad, npy_pad;  // multiples of 8 (DMMA fragment)
    double delta, inv_delta, norm;  // cell size, 1/(nx ny nz delta^3)
    double zcorr0, zcorr1, zcorr2;
    int n_pk;
    const double *lnk;      // [n_pk] ln k knots of the tabulated P(k)
    const double *pk_coef;  // [2][n_pk-1][4] cubic coefficients in ln k (fiducial, no-wiggles)
    double pk_k0, pk_k1;    // table range; power-law continuation outside
    double pk_end[2][2], pk_slope[2][2];
    const double *cosx;     // [nx] cos(2 pi t / nx)
    // tiling of the plane for fft_plane_kernel: rtiles x ctiles tiles of wr x ncf 8x8 fragments
    int rtiles, ctiles, wr, ncf;
    // Cy = cos(2 pi (my iy mod ny) / ny), stored in the kernel's shared-memory stage layout
    // [rtile][chunk of 16 my][72 rows][20] (zero padded) so that a stage is one contiguous bulk copy
    const double *Cy;
};
constexpr int kFftKC = 16;                 // k_par rows per pipeline stage of fft_plane_kernel
constexpr int kFftLdA = kFftKC + 4;        // row stride of a Cy stage
constexpr int kFftLdT = 9 * 8 + 4;         // row stride of a T stage (4 mod 8 doubles: conflict-free)
constexpr int kFftStageA = 9 * 8 * kFftLdA;        // doubles
constexpr int kFftStageT = 3 * kFftKC * kFftLdT;   // doubles

// Immutable model description passed by value to kernels.
struct DevModel {
    int kind;
    unsigned flags;
    int npar;
    double zref, omega_matter;
    int idx_beta, idx_bb, idx_gamma_bias, idx_gamma_beta, idx_delta_v, idx_bias2, idx_bb2, idx_beta_bias;
    int idx_bao, idx_comb_scale, idx_rad, idx_nl, idx_cont, idx_bs, idx_hcd, idx_uv, idx_smgaus, idx_smlor,
        idx_nlcorr;
    DevBB bb[4];
    // r-space
    DevSplineSet fid, nw;
    // k-space
    int nk, nk_pad, nr, nr_pad, ell_max;
    double tr_r0, tr_dr, tr_inv_dr, rlo, rhi, dilmin, dilmax, zeff, sigma8;
    const double *kc;      // [nk] coarse nodes
    const double *pk_kc;   // [2][nk] P_comp at coarse nodes (argument of the non-linear correction)
    const double *hankel;  // [2][3][nr_pad][nk_pad] Filon weights
    const double *th_w, *th_inv_diag;  // [nr-2] factors of the constant (1,4,1) spline system
    double mu_x[kMuNodes], mu_w[kMuNodes];
    // metals
    int metal_kind, idx_metal, metal_ncomb, metal_nmet, metal_ngrid;
    int metal_p1[20], metal_p2[20];
    const double *metal_auto;   // [ncomb*3][ngrid]
    const double *metal_cross;  // interleaved [ny][nx][ncomb*3]
    int metal_nx, metal_ny;
    double metal_x0, metal_y0, metal_inv_dx, metal_inv_dy, metal_xmax, metal_ymax;
    int dm_order;
    DevFft fft;
};

struct bf_ctx_t;
}  // namespace bf

struct bf_data;

// All vectors of a host-buffer call carry these key parameters (compared on the host, host_key_check in api.cu)
struct KeyHint {
    int nkey = 0;
    double vals[16] = {};
};

struct bf_ctx {
    int device = 0;
    // Every compute entry point holds this lock for the duration of the call: two host threads may share one
    // context (their calls are serialised on its stream); independent contexts never contend.
    std::recursive_mutex api_mu;
    std::vector<bf_data *> live_data;  // data objects of this context (bf_model_destroy purges their plans)
    cudaStream_t stream = nullptr;
    // bf_chi2_batch_submit / _wait: uploads run on a second stream into one of two staging slots, so that the
    // parameters of call n + 1 cross the bus while call n is being evaluated
    cudaStream_t copy_stream = nullptr;
    struct AsyncSlot {
        void *buf = nullptr;
        size_t bytes = 0;
        cudaEvent_t uploaded = nullptr, done = nullptr;
        bool pending = false;
        long long serial = 0;
    } async_slot[2];
    long long submit_serial = 0;
    const KeyHint *key_hint = nullptr;  // set by the host-buffer entry point around its run_device call
    // what run_chunk made of the last hinted k-space call (read by the graph cache of the host-buffer entry point)
    double last_hk[32] = {};
    bool last_host_current = false;
    // CUDA graphs of the kernel chains of small host-buffer chi-square calls (run_host)
    static constexpr int kGraphMaxB = 256;
    struct GraphKey {
        unsigned long long model_serial;
        const void *data;
        int B;
        const void *io, *ws;
        size_t ws_bytes;
        const void *chi2_part;
        bool operator<(const GraphKey &o) const {
            return std::tie(model_serial, data, B, io, ws, ws_bytes, chi2_part) < std::tie(o.model_serial, o.data, o.B, o.io, o.ws, o.ws_bytes, o.chi2_part);
        }
    };
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        int seen = 0;
        long long launches = 0;
        int nkey = 0;
        double hk[32] = {};
    };
    std::map<GraphKey, GraphEntry> graphs;
    bool no_graphs = false;  // BAOFIT_NO_GRAPHS
    long long launches = 0;
    // growable workspaces (device)
    void *ws = nullptr;
    size_t ws_bytes = 0;
    void *io = nullptr;  // staging for host-buffer entry points
    char *stage = nullptr;  // page-locked staging of small host-buffer chi-square calls (2 x 1 MB: in, out)
    void *aux = nullptr;  // small side buffers of the probe entry points (linear-parameter map, masked steps)
    size_t aux_bytes = 0;
    size_t io_bytes = 0;
    void *mp = nullptr;  // multipole-binned data: predictions at the expanded points + residual panel
    size_t mp_bytes = 0;
    double *chi2_part = nullptr;  // per-row-tile chi2 partial sums of small batches (launch_gemm_chi2)
    size_t chi2_part_bytes = 0;
    double *pinned_flag = nullptr;  // small pinned buffer for tiny device->host answers (flag + key values)
    int sm_count = 148;
    // optional per-kernel timing (bf_ctx_set_profiling): event pairs around every launch
    bool profiling = false;
    struct ProfRec { const char *name; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof_pending;
    std::map<std::string, std::pair<double, long long>> prof_acc;  // name -> (ms, launches)
};

struct bf_model {
    bf_ctx *ctx = nullptr;
    unsigned long long serial = 0;  // unique per created model: key of the (model, data) plans
    double dzmin = 0.5;             // bf_model_desc::dzmin (plan-time check of the reference's mid-sweep re-transform)
    bf::DevModel dev;
    std::vector<void *> allocs;  // device allocations owned by the model
    // host copies needed at plan time
    std::vector<double> metal_zgrid;  // [ncomb][ngrid]
    std::vector<double> dist_matrix;  // [order][order]
    std::vector<double> metal_cross_h;  // [ncomb*3][ny][nx] as given (for the plan tables)
    // shared-transform basis tables, kept across calls while the "key" parameters are unchanged
    // (the reference likewise re-transforms only when those parameters change,
    // baofit/BaoKSpaceCorrelationModel.cc:333-380)
    mutable double2 *d_tabs_cache = nullptr;
    double *d_keystate = nullptr;  // [32] key of the cached tables + valid marker (kspace_key_decide_kernel)
    int *d_gate = nullptr;         // [4] device-side decision of the current call: shared?, rebuild?
    // host mirror of d_keystate, maintained by calls that carry a KeyHint (all calls of a model run under its context's lock)
    mutable bool h_key_valid = false;
    mutable double h_key[32] = {};
    mutable std::vector<double> cached_key;
    mutable bool cache_valid = false;
    double metal_dx = 0, metal_dy = 0;
    double zref = 0;
    // FFT model: per-key tables T (x-transformed, weights folded in), stored in the plane kernel's
    // stage layout [2 comps][ctiles][chunks][3 powers of mu^2][16][76], and the scratch
    // H[2][3][my][mx] they are built from
    mutable double *d_fft_T = nullptr, *d_fft_H = nullptr;
    // r-space model: basis tables [2 comps][n knots][9] of (y, c) built from the tabulated multipoles when both
    // tabulations share one uniform grid; lets the model run through the k-space fast pass (kernels_fast.cu, kRS)
    double2 *d_rs_tabs = nullptr;
};

struct bf_data {
    bf_ctx *ctx = nullptr;
    int n = 0, n_pad = 0, n_grid = 0, n_grid_pad = 0;
    std::vector<double> h_r, h_mu, h_z, h_grid_r, h_grid_mu, h_grid_z;
    std::vector<int> h_index;
    std::vector<void *> allocs;
    double *d_r = nullptr, *d_mu = nullptr, *d_z = nullptr;  // [n_pad]
    int *d_index = nullptr;
    double *d_grid_r = nullptr, *d_grid_mu = nullptr, *d_grid_z = nullptr;  // [n_grid_pad]
    double *d_data = nullptr;  // [n_pad]
    double *d_W = nullptr;     // [n_pad][n_pad] lower triangular, W^T W = C^-1
    double *d_L = nullptr;     // [n_pad][n_pad] lower Cholesky factor of C (toy noise); may be null
    // multipole-binned data (BF_BINNING_MULTIPOLE): every bin is expanded into nmu Gauss-Legendre points
    // in mu, held by `expanded` (a points-only data object); the prediction is the weighted sum
    int binning_type = 0, nmu = 0;
    bf_data *expanded = nullptr;
    double *d_proj = nullptr;  // [n][nmu] weights (2 ell + 1)/2 w_q L_ell(mu_q)
    bool points_only = false;  // internal object without data vector / covariance
    bool trigger_mu_zero = false;  // the multipole overload triggers transforms with mu = 0 (AbsCorrelationModel.cc:73)
    std::vector<double> h_W;   // host copy [n][n] (folded into the fused tail matrices at plan time)
    std::vector<double> h_data;
    // plans: per model, the z-class table and the gathered distortion matrix rows
    struct Plan {
        int nz = 0;
        std::vector<double> zvals;
        double *d_zvals = nullptr;   // [nz]
        double *d_vfac = nullptr;    // [nz] (1/100)(1+z)/sqrt(1-Om+Om(1+z)^3)
        int *d_zc_data = nullptr;    // [n_pad]
        int *d_zc_grid = nullptr;    // [n_grid_pad]
        int *d_zc_metal_data = nullptr;  // [ncomb][n_pad]    class of metal zgrid[c][index_i]
        int *d_zc_metal_grid = nullptr;  // [ncomb][n_grid_pad]
        int zc_trigger = 0;          // class of the z that triggers the k-space transform
        bool z_retrigger = false;    // consecutive bins jump by more than dzmin in z (see get_plan)
        double *d_Dkept = nullptr;   // [n_pad][n_grid_pad] rows of the distortion matrix of kept bins
        // bin-static tables of the cartesian-form kernels (kernels_fast.cu)
        struct PointTabDev {
            double *rpar0 = nullptr, *rperp0 = nullptr;
            double *bb_rt[4] = {nullptr, nullptr, nullptr, nullptr}, *bb_z[4] = {nullptr, nullptr, nullptr, nullptr};
            int n_rt[4] = {0, 0, 0, 0}, n_z[4] = {0, 0, 0, 0};
            double *mx = nullptr, *mxh = nullptr;
        } tab_data, tab_grid;
        // Fused tail of the distortion-matrix path: chi2 = || WA . [xi_u ; coef(b) ; 1] ||^2 with
        // WA = W . [D_kept | broadband basis | -d]  (see DESIGN.md "fused tail")
        bool fold_ok = false;
        int fold_ncls = 1;            // redshift classes with their own block of broadband rows (data mode)
        bool fold_data_mode = false;  // fused tail without a distortion matrix: d_WA = [W | W.basis | -W.d]
        int fold_case = 0;            // 0: no post broadband, 1: static basis (no velocity shift), 2: binomial in dpi
        int fold_nb = 0, fold_kext = 0;  // number of basis columns; extra K rows (multiple of 16)
        int fold_pmax = 0, fold_nrt = 0, fold_nz = 0;
        double *d_Aext = nullptr;     // [n_pad][n_grid_pad + fold_kext]
        double *d_WA = nullptr;       // [n_pad][n_grid_pad + fold_kext]
        double *d_WAt = nullptr;      // the same matrix in the stage layout of dgemm_chi2_fullm_kernel (n_pad <= 512)
        double *d_WAf = nullptr;      // the same matrix tiled [K/4][M][4] for fused_model_chi2_kernel
        double *d_WAf3 = nullptr;     // ... and in the streaming order of fused3_model_chi2_kernel (gemm_fused_tile_A3)
        bool fused_tri = false;       // d_WAf3 holds the trapezoidal R factor of WA in interleaved fragment order (get_plan)
        std::vector<void *> allocs;
    };
    std::map<unsigned long long, Plan> plans;  // keyed by bf_model::serial (never by address: addresses are reused)
};

struct bf_toys {
    bf_ctx *ctx = nullptr;
    int ntoy = 0, n = 0;
    double *d_table = nullptr;  // [ntoy][n]
};
