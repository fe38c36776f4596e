// Kernel argument blocks and launchers (implemented in kernels_model.cu / kernels_gemm.cu).
#pragma once
#include <atomic>
#include <utility>
#include "bf_internal.h"

namespace bf {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: remembers, per device, how much has been
// configured for one kernel (a process may hold contexts on several GPUs; launchers may run on several host threads).
struct SmemOptIn {
    static constexpr int kMaxDevices = 64;
    std::atomic<size_t> configured[kMaxDevices];
    SmemOptIn() { for (auto &c : configured) c.store(0); }
    template <class F> void ensure(F func, size_t bytes) {
        int dev = 0;
        cudaGetDevice(&dev);
        const bool tracked = dev >= 0 && dev < kMaxDevices;
        if (tracked && configured[dev].load(std::memory_order_acquire) >= bytes) return;
        cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (tracked) configured[dev].store(bytes, std::memory_order_release);
    }
};


// Programmatic dependent launch. A call is a chain of ~20 small dependent kernels on one stream, and a plain launch
// boundary costs 3-4 us on a B200; with the stream-serialization attribute the next grid is set up while its predecessor
// drains: 1.7 us per boundary (tools/pdl_probe.cu, profiles/pdl_probe_r02.json). EVERY kernel of the library starts with
// pdl_sync() = griddepcontrol.wait: nothing of a kernel runs before all the work before it in the stream is complete and
// visible, so the order of memory effects is that of plain launches. No kernel triggers its successor early
// (griddepcontrol.launch_dependents): the probe shows no gain over the implicit trigger at grid completion, and parked
// CTAs of a successor next to a running multi-wave grid cost the FFT model and the lock-step scans 25 % when it was tried.
// BAOFIT_NO_PDL=1 switches the attribute off (A/B measurements).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

bool pdl_enabled();  // api.cu

struct Launch {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr;
    Launch(dim3 grid, dim3 block, size_t smem, cudaStream_t s) {
        cfg = cudaLaunchConfig_t{};
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = pdl_enabled() ? 1 : 0;
    }
    template <class... K, class... A> void operator()(void (*kernel)(K...), A &&...args) {
        cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
    }
};

// Device-side predication: a kernel whose gate is set returns at once unless *flag == want. The decision which of the
// two k-space chains runs (shared basis tables / per-vector transforms) and whether the basis tables must be rebuilt is
// taken by kspace_key_decide_kernel on the GPU, so no entry point waits for the device in the middle of a call.
struct Gate {
    const int *flag;
    int want;
};
#define BF_GATE(g) do { if ((g).flag && *(g).flag != (g).want) return; } while (0)

struct EvalArgs {
    Gate gate;
    DevModel m;
    const double *pT;  // transposed parameters [npar][ldb]
    const double *S;   // evolution factors [nz*3][ldb]
    int ldb, B;
    // evaluation points (data bins or grid bins)
    int npts, npts_pad, pts_per_block;
    const double *r, *mu, *z;
    const int *zc;        // redshift class of each point
    const int *gidx;      // global grid index of each point (null: the point index itself)
    const int *zc_metal;  // [ncomb][zc_stride] class of the metal grid redshift at each point
    int zc_stride;
    const double *vfac;   // [nz] velocity-shift factor of each class
    // k-space per-vector multipole tables and their spline second derivatives
    const double *T, *C2;
    // shared-transform path: basis tables {y, y''/2} [2 comps][nr][9 = 3 ell x 3 powers of mu^2]
    const double2 *tabs;
    int zc_trigger;
    int mode;
    // output: out[i*ldo + b] = value - (dov ? dov[b*npts+i] : dsub ? dsub[i] : 0)
    const double *in;
    double *out;
    int ldo;
    const double *dsub;
    const double *dov;
    int *status;
};

// Bin-static tables of one point set (data bins or grid bins), built once per (model, data) plan.
struct PointTab {
    int npts, npts_pad;
    const double *rpar0, *rperp0, *z;  // [npts_pad] r*mu, r*sqrt(1-mu^2), redshift
    const int *zc;                     // redshift class
    const int *gidx;                   // global grid index (null: the point index itself)
    const int *zc_metal;               // [ncomb][zc_stride]
    int zc_stride;
    // r_T- and z-axis factors of the broadband polynomials, [npts][n_rt] / [npts][n_z] per slot
    const double *bb_rt[4], *bb_z[4];
    int n_rt[4], n_z[4];
    // cross metal templates collapsed along r_perp: mx[(i*mx_ny + row)*mx_nt + t]
    const double *mx;
    const double *mxh;                 // the same with halo rows: mxh[(i*(mx_ny+3) + row+1)*mx_nt + t], row = -1 .. mx_ny+1 (wrapped copies)
    int mx_ny, mx_nt;
};

struct FastArgs {
    Gate gate;
    DevModel m;
    PointTab pt;
    const double *pT, *S, *vfac;
    const double2 *tabs;
    int ldb, B, nz, zc_trigger, mode, pts_per_block;
    const double *in;
    double *out;
    int ldo;
    const double *dsub, *dov;
    int *status;
    int rspace;            // the r-space model through the cosmology pass (tables built at model creation)
    int sub_dov_in_part1;  // fused data tail: the cosmology pass subtracts the per-column data override itself
};

// Fused model + distortion matrix + chi-square kernel (kernels_fused.cu): the model evaluation on the grid bins
// is the producer of the right-hand operand of chi2 = || WA . [xi_u ; coefficient rows] ||^2.
struct FusedArgs {
    Gate gate;
    DevModel m;
    PointTab pt;                 // grid bins
    const double *pT, *S, *vfac;
    const double2 *tabs;         // shared-transform basis tables
    int ldb, B, zc_trigger;
    const double *At;            // WA pre-tiled as [K/4][M][4] (gemm_fused_tile_A)
    const double *At3;           // WA in the streaming order of the TMEM-parking kernel (gemm_fused_tile_A3), or null
    int tri;                     // At3 is the upper-trapezoidal R factor of WA, fragment f at (warp f % 4, half (f / 4) % 2, slot f / 8):
                                 // fragments whose first row 8 f lies beyond the current K columns are skipped
    int M, K, kgrid;             // rows of WA; K = kgrid + kext; rows 0..kgrid-1 of the panel come from the model
    const double *rows;          // [kext][ldb] coefficient rows (fold_coef_kernel)
    double *chi2;
    const int *status;
    int debug;                   // timing experiments only (BAOFIT_FUSED_DEBUG): 2 = consumers skip the DMMAs (producer-bound time)
};
void gemm_fused_tile_A(const double *A, int M, int K, int lda, double *At /* [K/4][M][4] */);
void gemm_fused_tile_A3(const double *A, int M, int K, int lda, double *At /* [K/16][2][4][M/2][4], K % 16 == 0, M even */);
bool fused_chi2_supported(const FusedArgs &a);
void launch_fused_model_chi2(cudaStream_t s, const FusedArgs &a, int sm_count);

// Coefficient rows of the fused tail + dilation-limit check of the data bins.
struct FoldArgs {
    Gate gate;
    DevModel m;
    const double *pT, *S, *vfac;
    int ldb, B, zc_data;         // single redshift class of the data bins
    int fold_case, nb, kext, pmax, nrt, nz;
    int ncls;                    // > 1: nb = ncls blocks of nb/ncls coefficient rows, block c scaled by the bias evolution of class c
    int minus_d_row_value;       // 1: chi2 mode (row multiplies the -d column), 0: prediction mode
    double *rows;                // [kext][ldb] appended below the xi_u panel
    // data bins for the dilation check
    int npts;
    const double *rpar0, *rperp0;
    int *status;
};

struct ProjArgs {
    Gate gate;
    DevModel m;
    const double *pT, *S;
    int ldb, B, zc_trigger, k_per_block;
    double nl_tab[5];    // qnl, kv, av, bv, kp interpolated at z_eff (NonLinearCorrectionModel.cc:22-27)
    double nl_pk_scale;  // (sigma8Sim/sigma8)^2 ((1+zeff)/(1+zref))^-2   (:62-64)
    double *G;           // [2][3][nk_pad][ldb]
};

// C[M x N] = A[M x K] . B[K x N]; A row-major lda, B row-major ldb (N = parameter vectors),
// batched over blockIdx.z with the given strides. Operands are padded: M, K multiples of the
// tile sizes with zero fill in A.
struct GemmArgs {
    Gate gate;
    const double *A;
    const double *Bm;
    double *C;
    int M, N, K;  // padded M and K; N = number of valid columns
    int lda, ldb, ldc;
    long long strideA, strideB, strideC;
    int lower_triangular;  // A is lower triangular: skip k tiles above the diagonal
    int k_tri;             // with lower_triangular: only the first k_tri columns of A are triangular, columns
                           // k_tri .. K-1 are dense for every row (0 = all K columns are triangular)
    // chi-square epilogue: chi2[n] = sum_m C[m][n]^2 (C is not stored)
    double *chi2;
    const int *status;
    // small batches: one CTA per (column panel, row tile) writes chi2_part[row tile][part_ld], summed in fixed
    // order by a second kernel (launch_gemm_chi2 decides; scratch of (M/64) * part_ld doubles or null)
    double *chi2_part;
    int part_ld;
};

void launch_prep(cudaStream_t s, const DevModel &m, const double *params, int B, int ldb, int nz,
                 const double *zvals, const int *rank, double *pT, double *S, int *status, bool z_retrigger,
                 double *key_flag /* key_flag[0] = 0: the key checks that follow raise it */);
// counting sort of the vectors on their quantised dilation parameters; hist/base: 2048 ints each
void launch_coherence_sort(cudaStream_t s, const DevModel &m, const double *params, int B, int *hist, int *base, int *rank);
void launch_unpermute(cudaStream_t s, int B, const int *rank, const double *chi2_sorted, const int *status_sorted,
                      double *chi2, int *status);
constexpr int kSortTableInts = 2048;
void launch_rspace_eval(cudaStream_t s, EvalArgs &a, int sm_count);
void launch_kspace_project(cudaStream_t s, ProjArgs &a, int sm_count);
void launch_kspace_thomas(cudaStream_t s, int B, int ldb, int nr, int nr_pad, double h, const double *th_w,
                          const double *th_inv_diag, const double *T, double *C2, Gate gate = Gate{nullptr, 0});
void launch_kspace_eval(cudaStream_t s, EvalArgs &a, int sm_count);
// shared-transform fast path (all vectors of the batch share the non-polynomial part of D(k,mu))
// keyout[0] = 1 if any vector's key differs from vector 0's, keyout[1..] = key values of vector 0
void launch_kspace_key_check(cudaStream_t s, const DevModel &m, const double *pT, int ldb, int B, double *keyout);
constexpr int kKeyMax = 12;
void launch_kspace_basis_project(cudaStream_t s, ProjArgs &a);  // a.G: [6][nk_pad][kPadN], columns 0..2
void launch_kspace_basis_finish(cudaStream_t s, const DevModel &m, const double *T, double2 *tabs, Gate gate = Gate{nullptr, 0});
// Device-side decision after kspace_key_check_kernel: gate[0] = 1 if all vectors share their key parameters, gate[1] = 1
// if they do and the key differs from the one the cached basis tables were built for (state[32], kept in the model)
struct KeyExtras { double v[7]; };  // z_eff, P(k) rescale, the five interpolated nlcorr coefficients
void launch_kspace_key_decide(cudaStream_t s, const double *flag, double *state, int *gate, const KeyExtras &ex);
void launch_kspace_eval_shared(cudaStream_t s, EvalArgs &a, int sm_count);
void launch_dist_post(cudaStream_t s, EvalArgs &a, int sm_count);
void launch_kspace_fast_eval(cudaStream_t s, FastArgs &a, int sm_count, int part);  // part 0 all, 1 cosmology, 2 extras
bool fast_eval_is_lean(const FastArgs &a);
bool fast_eval_has_extras(const FastArgs &a);
void launch_fast_post(cudaStream_t s, FastArgs &a, int sm_count);
void launch_fold_coef(cudaStream_t s, const FoldArgs &a);

// ---- 3-D FFT k-space model (kernels_fft.cu) ----
constexpr int kFftKeyMax = 8;
struct FftKeyState {        // what the per-key tables depend on (values of vector 0 of the chunk)
    double snl_perp2, snl_par2;  // SigmaNL^2 of the peak component (no-wiggles: 0 unless nl-broadband)
    double nl_tab[5];            // qnl, kv, av, bv, kp at z_eff (NonLinearCorrectionModel.cc:55-61)
    double nl_pk_scale;
};
struct FftVecArgs {
    DevModel m;
    const double *pT;
    int ldb, B;
    // bin that triggers the transform (first kept bin); use_ztrig: take ztrig as given instead
    double r0, mu0, z0, ztrig;
    int use_ztrig;
    double *keyflag;  // [0] = 1 if a vector's key differs from vector 0's; [1..] key of vector 0
    double *keys;     // [kFftKeyMax-1][ldb] key of every vector (read back only when they differ)
    double *c12;      // [2][ldb]  c1 = beta_z + beta2_z, c2 = beta_z beta2_z
    double *cont;     // [my_pad][ldb] continuum-fitting distortion at k_par = 2 pi my / (ny delta)
    int *status;
};
void launch_fft_vec(cudaStream_t s, const FftVecArgs &a);
void launch_fft_tables(cudaStream_t s, const DevModel &m, const FftKeyState &k, double *H, double *T);
// planes[((b*2 + comp)*npy_pad + iy)*npx_pad + ix] for the `count` columns b = cols[i] (cols null: b = i)
void launch_fft_planes(cudaStream_t s, const DevModel &m, const double *T, const double *c12, const double *cont,
                       int ldb, const int *cols, int count, double *planes);
struct FftEvalArgs {
    DevModel m;
    const double *pT;
    int ldb, B;
    int npts;
    const double *r, *mu, *z;  // [npts] data bins
    const double *planes;
    double *out;               // out[i*ldo + b] = xi - (dov ? dov[b*npts+i] : dsub ? dsub[i] : 0)
    int ldo;
    const double *dsub, *dov;
    int *status;
};
void launch_fft_eval(cudaStream_t s, const FftEvalArgs &a);

// multipole projection: out[i*ldo + b] = sum_q proj[i*nmu + q] X[(i*nmu + q)*ldx + b] - (dov ? dov[b*n + i] : dsub ? dsub[i] : 0)
void launch_multipole_project(cudaStream_t s, const double *X, int ldx, const double *proj, int n, int nmu, int B,
                              double *out, int ldo, const double *dsub, const double *dov);

// per fit f: Gram matrix of [y0, y1-y0, ..., y_{G-1}-y0] over the columns f*G .. f*G+G-1 of Y (ld = ldy)
void launch_gram(cudaStream_t s, const double *Y, int ldy, int nrows, int G, int nfit, double *out);
// normal equations of central probes: out[f] = (n+1)^2 matrix [y0 a]^T [y0 a] then n values y0 . s (see the kernel)
// part: normal_probes_part_doubles(n, nfit) doubles of scratch (partial sums of the row slices)
size_t normal_probes_part_doubles(int n, int nfit);
void launch_normal_probes(cudaStream_t s, const double *Y, int ldy, int nrows, int n, int nfit, double *out, double *part);
// The same normal equations when some probed parameters enter the whitened residual LINEARLY with a bin-static basis
// (the additive broadband coefficients of a model without velocity shift): their columns a_i = 2 h_i dy/dx_i are
// synthesised from the static matrix WA (tail columns = W . basis, bf_data::Plan) instead of being evaluated, so a
// fit costs 2 n_nl + 1 model evaluations instead of 2 n + 1. lin[j] = coefficient index of parameter j, or -1.
struct NormalLinArgs {
    const double *Y;            // whitened residuals of the centre and the 2 n_nl non-linear probes, [nrows][ldy]
    int ldy, nrows, n, n_nl, nfit, npar;
    const double *centres, *steps;  // [nfit][npar] (device)
    const int *lin;                 // [npar] (device)
    const double *WA;               // [n_pad][lda], basis columns start at kleft
    int lda, kleft, nb1, ncls;
    // fold_case 2 (velocity shift, fold_coef_kernel): coefficient (z, p, t) feeds the rows (z, m <= p, t) with C(p, m) v^(p-m),
    // v = delta-v . vfac[zc0] / r0 of the fit's centre
    int fold_case, pmax, nrt, n_rp, rp_min, rp_step, idx_delta_v;
    double inv_r0;
    const double *vfac;
    const double *zvals;            // [ncls] redshift of every class (single class: the class of the data bins)
    int zc0;                        // first class index into zvals when ncls == 1
    double zref;
    int idx_gamma_bias;
    double *out;                    // [nfit][(n+1)^2 + n]
    double *part;                   // scratch, normal_probes_part_doubles(n, nfit)
};
void launch_normal_probes_lin(cudaStream_t s, const NormalLinArgs &a);

// params[(f*G + g)*npar + j] of the central-difference probe groups (bf_gram_probes); status_fit[f] = OR of status[f*G..]
void launch_expand_probes(cudaStream_t s, const double *centres, const double *steps, int nfit, int npar, int nprobe, double *params,
                          const int *skip /* null, or [npar]: parameters with skip[j] >= 0 are not probed */);
void launch_reduce_status(cudaStream_t s, const int *status, int nfit, int G, int *status_fit);
// override panel of the toy path: out[b][i] = table[index[b]][i]
void launch_gather_rows(cudaStream_t s, const double *table, const int *index, int B, int n, double *out);

// Full-M chi-square GEMM (M <= 512): one CTA owns all rows of a 32-column panel, so the panel is read from
// HBM exactly once; the left matrix streams from L2 in the pre-tiled stage layout At[K/8][M][12]
// (see gemm_fullm_tile_A). chi2[n] = sum_m (A.B)[m][n]^2.
constexpr int kFullmKC = 8, kFullmLdA = kFullmKC + 4, kFullmBN = 32, kFullmMaxM = 512;
void gemm_fullm_tile_A(const double *A, int M, int K, int lda, double *At /* [K/8][M][12] */);
void launch_gemm_chi2_fullm(cudaStream_t s, const GemmArgs &g /* g.A = At */, int sm_count);

void launch_gemm_store(cudaStream_t s, const GemmArgs &g, int batch, int sm_count = 0);  // sm_count > 0: small batches may take 32 x 32 tiles
void launch_gemm_chi2(cudaStream_t s, const GemmArgs &g, int sm_count = 0);
// columns up to which launch_gemm_chi2 splits the row tiles over CTAs (scratch sizing)
int gemm_chi2_split_columns(int M, int sm_count);

// toy noise: out[t][i] = truth_i + scale * sum_j L[i][j] g_{t,j} (deviate panel, lower-triangular DMMA GEMM, transpose);
// work: toy_noise_work_doubles(n_pad, ntoy) doubles of scratch
constexpr int kToyPassCols = 4096;  // toys per pass
size_t toy_noise_work_doubles(int n_pad, int ntoy);
void launch_toy_noise(cudaStream_t s, const double *L, int n, int n_pad, const double *truth,
                      unsigned long long seed, long long first_toy, int ntoy, double scale, double *out, double *work);

}  // namespace bf
