// Host-side C++ mirror of the reference's hot-path interfaces (namespace baofit).
//
// Class names, constructor argument order, parameter names / order and error messages follow
// the reference so that these remain drop-in plugins configured by the same .ini files:
//   AbsCorrelationModel        baofit/AbsCorrelationModel.{h,cc}
//   BaoCorrelationModel        baofit/BaoCorrelationModel.{h,cc}
//   BaoKSpaceCorrelationModel  baofit/BaoKSpaceCorrelationModel.{h,cc}
//   BaoKSpaceFftCorrelationModel  baofit/BaoKSpaceFftCorrelationModel.{h,cc}
//   BroadbandModel             baofit/BroadbandModel.{h,cc}
//   MetalCorrelationModel      baofit/MetalCorrelationModel.{h,cc}
//   NonLinearCorrectionModel   baofit/NonLinearCorrectionModel.{h,cc}
//   DistortionMatrix           baofit/DistortionMatrix.{h,cc}
//   AbsCorrelationData, ComovingCorrelationData, loadCorrelationData   baofit/*Data.{h,cc}
//   CorrelationFitter          baofit/CorrelationFitter.{h,cc}
//   CorrelationAnalyzer        baofit/CorrelationAnalyzer.{h,cc} (fitSample, parameterScan, toy MC)
// What differs: the models do not evaluate anything themselves. Their constructors lay out the
// parameters and describe the model in a bf_model_desc; every number is computed by the CUDA
// library through the C ABI (include/baofit_b200.h), batched over parameter vectors.
#pragma once

#include <array>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../../include/baofit_b200.h"
#include "likely_lite.h"

namespace baofit {

// ---- device binding ---------------------------------------------------------------------------
// One CUDA context per process (one process per GPU); created on first use.
bf_ctx *deviceContext();
void setDevice(int device);  // call before the first use to pick a GPU (default 0)
void releaseDeviceContext();

// ---- models -----------------------------------------------------------------------------------
class AbsCorrelationModel {
public:
    explicit AbsCorrelationModel(std::string const &name);
    virtual ~AbsCorrelationModel();
    // likely::FitModel interface (subset)
    int defineParameter(std::string const &name, double value, double error);
    int getNParameters() const { return (int)_parameters.size(); }
    void configureFitParameters(std::string const &script);
    likely::FitParameters const &getFitParameters() const { return _parameters; }
    int findParameter(std::string const &name) const;  // -1 if absent
    std::string const &getName() const { return _name; }
    // Sets the grid coordinates to use for the distortion matrix (AbsCorrelationModel.cc:51-63).
    void setCoordinates(std::vector<double> rbin, std::vector<double> mubin, std::vector<double> zbin);
    std::vector<double> const &gridR() const { return _rbin; }
    std::vector<double> const &gridMu() const { return _mubin; }
    std::vector<double> const &gridZ() const { return _zbin; }
    // Single-point evaluation kept for interface parity (AbsCorrelationModel.cc:21-41). It runs
    // a batch of one through the CUDA path; use CorrelationFitter for anything hot.
    double evaluate(double r, double mu, double z, likely::Parameters const &params, int index);
    // The device-resident model (created on first use, after all parameters are defined).
    bf_model *deviceModel();
    bf_model_desc const &description() { finalizeDescription(); return _desc; }
    virtual void printToStream(std::ostream &out) const;

protected:
    friend class BroadbandModel;
    friend class MetalCorrelationModel;
    friend class NonLinearCorrectionModel;
    friend class DistortionMatrix;
    // Defines the standard linear bias parameters (AbsCorrelationModel.cc:91-130); returns the
    // index of the last parameter defined.
    int _defineLinearBiasParameters(double zref, bool crossCorrelation = false, bool combinedBias = false);
    void _setZRef(double zref);
    void _setOmegaMatter(double OmegaMatter);
    virtual void finalizeDescription() {}
    bf_model_desc _desc;
    // owned copies of every array the descriptor points at
    std::vector<std::vector<double>> _arrays;
    std::vector<std::vector<int>> _iarrays;
    const double *_keep(std::vector<double> const &v);
    const int *_keep(std::vector<int> const &v);

private:
    std::string _name;
    likely::FitParameters _parameters;
    std::vector<double> _rbin, _mubin, _zbin;
    bf_model *_device = nullptr;
};
typedef std::shared_ptr<AbsCorrelationModel> AbsCorrelationModelPtr;

// Parses the broadband specification (grammar of BroadbandModel.cc:21-90) and registers the
// coefficients on the parent model (BroadbandModel.cc:141-166).
class BroadbandModel {
public:
    BroadbandModel(std::string const &name, std::string const &tag, std::string const &paramSpec, double r0,
                   double z0, AbsCorrelationModel *base);
    bf_broadband_spec const &spec() const { return _spec; }
    int _getIndexBase() const { return _spec.idx_base; }

private:
    bf_broadband_spec _spec;
};

// Reads the dense matrix from "<name>.dmat" lines "i j value" (DistortionMatrix.cc:22-73).
class DistortionMatrix {
public:
    DistortionMatrix(std::string const &distMatrixName, int distMatrixOrder, bool verbose = false);
    DistortionMatrix(std::vector<double> const &dense, int order);  // in-memory variant
    double getDistortion(int index1, int index2) const;             // DistortionMatrix.cc:77-83
    int order() const { return _nbins; }
    std::vector<double> const &dense() const { return _dist; }

private:
    int _nbins;
    std::vector<double> _dist;
};

// Registers the metal-line parameters on the parent and loads the templates
// (MetalCorrelationModel.cc:27-161).
class MetalCorrelationModel {
public:
    MetalCorrelationModel(std::string const &metalModelName, bool metalModel, bool metalModelInterpolate,
                          bool metalCIV, bool toyMetal, bool crossCorrelation, AbsCorrelationModel *base);
    void fill(bf_model_desc &d, AbsCorrelationModel &owner) const;

private:
    int _kind, _indexBase, _ncomb = 0, _nmet = 0, _ngrid = 0, _nx = 0, _ny = 0;
    double _x0 = 0, _y0 = 0, _dx = 0, _dy = 0;
    std::vector<int> _p1, _p2;
    std::vector<double> _zgrid, _auto, _cross;
    friend class AbsCorrelationModel;
};

class BaoCorrelationModel : public AbsCorrelationModel {
public:
    // Same argument list as baofit/BaoCorrelationModel.h (BaoCorrelationModel.cc:18-23)
    BaoCorrelationModel(std::string const &modelrootName, std::string const &fiducialName,
                        std::string const &nowigglesName, std::string const &metalModelName,
                        std::string const &distAdd, std::string const &distMul, double distR0, double zref,
                        double OmegaMatter, bool anisotropic = false, bool decoupled = false,
                        bool metalModel = false, bool metalModelInterpolate = false, bool metalCIV = false,
                        bool toyMetal = false, bool combinedBias = false, bool combinedScale = false,
                        bool crossCorrelation = false);

private:
    std::shared_ptr<MetalCorrelationModel> _metalCorr;
    std::shared_ptr<BroadbandModel> _distortAdd, _distortMul;
};

class BaoKSpaceCorrelationModel : public AbsCorrelationModel {
public:
    // Same argument list as baofit/BaoKSpaceCorrelationModel.h (BaoKSpaceCorrelationModel.cc:26-39)
    BaoKSpaceCorrelationModel(std::string const &modelrootName, std::string const &fiducialName,
                              std::string const &nowigglesName, std::string const &distMatrixName,
                              std::string const &metalModelName, double zref, double OmegaMatter, double rmin,
                              double rmax, double dilmin, double dilmax, double relerr, double abserr, int ellMax,
                              int samplesPerDecade, std::string const &distAdd, std::string const &distMul,
                              double distR0, double zeff, double sigma8, double dzmin, int distMatrixOrder,
                              std::string const &distMatrixDistAdd, std::string const &distMatrixDistMul,
                              bool anisotropic = false, bool decoupled = false, bool nlBroadband = false,
                              bool nlCorrection = false, bool fitNLCorrection = false, bool nlCorrectionAlt = false,
                              bool binSmooth = false, bool binSmoothAlt = false, bool hcdModel = false,
                              bool hcdModelAlt = false, bool uvfluctuation = false, bool radiationModel = false,
                              bool smoothGauss = false, bool smoothLorentz = false, bool distMatrix = false,
                              bool metalModel = false, bool metalModelInterpolate = false, bool metalCIV = false,
                              bool toyMetal = false, bool combinedBias = false, bool combinedScale = false,
                              bool crossCorrelation = false, bool verbose = false);
    // in-memory distortion matrix (synthetic workloads, tests)
    void setDistortionMatrix(std::shared_ptr<DistortionMatrix> dm);

private:
    std::shared_ptr<MetalCorrelationModel> _metalCorr;
    std::shared_ptr<DistortionMatrix> _distMat;
    std::shared_ptr<BroadbandModel> _distortAdd, _distortMul, _distMatDistortAdd, _distMatDistortMul;
};

class BaoKSpaceFftCorrelationModel : public AbsCorrelationModel {
public:
    // Same argument list as baofit/BaoKSpaceFftCorrelationModel.h (BaoKSpaceFftCorrelationModel.cc:21-28);
    // planeRMax (extra, default rmax*dilmax of the k-space options) bounds the tabulated plane.
    BaoKSpaceFftCorrelationModel(std::string const &modelrootName, std::string const &fiducialName,
                                 std::string const &nowigglesName, double zref, double OmegaMatter, double spacing,
                                 int nx, int ny, int nz, std::string const &distAdd, std::string const &distMul,
                                 double distR0, double zcorr0, double zcorr1, double zcorr2, double sigma8,
                                 bool anisotropic = false, bool decoupled = false, bool nlBroadband = false,
                                 bool nlCorrection = false, bool fitNLCorrection = false, bool nlCorrectionAlt = false,
                                 bool distortionAlt = false, bool noDistortion = false, bool crossCorrelation = false,
                                 bool verbose = false, double planeRMax = 300.);

private:
    std::shared_ptr<BroadbandModel> _distortAdd, _distortMul;
};

// ---- data -------------------------------------------------------------------------------------
class BinnedGrid {
public:
    BinnedGrid() {}
    BinnedGrid(std::vector<double> a1, std::vector<double> a2, std::vector<double> a3);
    int getNBinsTotal() const { return (int)(_axis[0].size() * _axis[1].size() * _axis[2].size()); }
    // bin centres of global index (axis 1 slowest, axis 3 fastest)
    void getBinCenters(int index, double centers[3]) const;
    void getBinWidths(int index, double widths[3]) const;
    void setWidths(std::vector<double> w1, std::vector<double> w2, std::vector<double> w3);

private:
    std::vector<double> _axis[3], _width[3];
};
// eigenmodes of a dense symmetric matrix: values ascending, vectors[k*n + i] = component i of mode k
void symmetricEigenModes(int n, std::vector<double> a, std::vector<double> &values, std::vector<double> &vectors);
// regularised upper incomplete gamma function Q(a, x) = 1 - P(a, x) (boost::math::gamma_p in the reference, CorrelationAnalyzer.cc:117)
double gammaQ(double a, double x);
class AbsCorrelationData;
// data vector and covariance (never the inverse) over the bins with data (finalized = false) or over the kept bins of a
// finalized data set
void dataAndCovariance(AbsCorrelationData const &data, bool finalized, std::vector<double> &d, std::vector<double> &cov);
BinnedGrid createCorrelationGrid(std::string const &axis1Bins, std::string const &axis2Bins,
                                 std::string const &axis3Bins);  // AbsCorrelationData.cc:96-150

class AbsCorrelationData {
public:
    enum TransverseBinningType { Coordinate, Multipole };
    AbsCorrelationData(BinnedGrid grid, TransverseBinningType type);
    virtual ~AbsCorrelationData();
    virtual double getRadius(int index) const = 0;
    virtual double getCosAngle(int index) const { return 0; }
    virtual int getMultipole(int index) const { (void)index; return 0; }
    TransverseBinningType getTransverseBinningType() const { return _type; }
    virtual double getRedshift(int index) const = 0;
    // BinnedData subset
    // weighted = the value is an element of Cinv.d (".wdata" files); such data is un-weighted with the full
    // covariance over the bins with data when the data set is finalized
    void setData(int index, double value, bool weighted = false);
    bool hasData(int index) const { return _values.count(index) > 0; }
    // custom bin centres (".grid" files, AbsCorrelationData.cc:243-274): replace the grid's centres once every
    // bin of the grid has been given one
    void setCustomBinCenters(int index, double c1, double c2, double c3);
    int getNCustomBins() const { return (int)_custom.size(); }
    bool useCustomGrid() const { return _useCustom; }
    void enableCustomGrid();
    void copyCustomGridFrom(AbsCorrelationData const &other);
    bool isWeighted() const { return _weighted; }
    void getBinCenters(int index, double centers[3]) const;
    void setCovariance(int i, int j, double value);
    void setInverseCovariance(int i, int j, double value);
    int getNBinsWithData() const { return (int)_values.size(); }
    BinnedGrid const &getGrid() const { return _grid; }
    std::vector<int> const &keptIndices() const { return _kept; }
    double getData(int index) const;
    double getCovarianceElement(int i, int j) const;  // as loaded (covariance or inverse); 0 if absent
    // AbsCorrelationData.cc:34-54
    void setFinalCuts(double rMin, double rMax, double rVetoMin, double rVetoMax, double muMin, double muMax,
                      double rperpMin, double rperpMax, double rparMin, double rparMax, int lMin, int lMax,
                      double zMin, double zMax);
    // applies the final cuts (AbsCorrelationData.cc:67-94), prunes, and freezes the data set
    virtual void finalize();
    bool isFinalized() const { return _finalized; }
    // "project-modes-keep" (src/baofit.cc:194,597-603): before finalize(), replaces the data vector by its projection onto the
    // nkeep largest- (nkeep > 0) or smallest-variance (nkeep < 0) eigenmodes of the covariance over the bins with data
    void projectOntoModes(int nkeep);
    // "fix-mode-scales" (src/baofit.cc:192,570-572): multiplies the eigenvalues of the covariance over the bins with data by
    // modeScales (ascending eigenvalue order, all > 0); before finalize()
    void rescaleEigenvalues(std::vector<double> const &modeScales);
    // coordinates of every grid bin, for AbsCorrelationModel::setCoordinates
    void gridCoordinates(std::vector<double> &r, std::vector<double> &mu, std::vector<double> &z) const;
    // device-resident data (kept bins, d, Cholesky factors); gridFor: model whose distortion
    // matrix needs the full grid (may be null)
    bf_data *deviceData(bool withGrid);
    // all bins with data (before final cuts), ascending index; dense covariance / inverse over them
    std::vector<int> binsWithData() const;
    std::vector<double> denseOverBinsWithData() const;
    std::vector<double> dataVector() const;           // kept bins, iteration order
    std::vector<double> covarianceMatrix() const;     // dense kept x kept (cov or icov as loaded)
    bool isInverse() const { return _isInverse; }
    // "save.data" / "save.icov" of src/baofit.cc:612-626: the finalized data set in the loader's own text formats
    // ("index value", "i j Cinv_ij" for j <= i), %.17g so that reloading with load-icov reproduces the fit
    void saveData(std::ostream &out) const;
    void saveInverseCovariance(std::ostream &out) const;

protected:
    BinnedGrid _grid;
    TransverseBinningType _type;
    // format-specific cut applied to bins that passed the final cuts (true = drop the bin);
    // QuasarCorrelationData::finalize, QuasarCorrelationData.cc:112-127
    virtual bool extraCut(int index) const { (void)index; return false; }

private:
    std::map<int, double> _values;
    std::map<int, std::array<double, 3>> _custom;
    bool _useCustom = false, _weighted = false;
    std::map<std::pair<int, int>, double> _cov;
    bool _isInverse = false, _haveCov = false, _haveFinalCuts = false, _finalized = false;
    double _rMin, _rMax, _rVetoMin, _rVetoMax, _muMin, _muMax, _rperpMin, _rperpMax, _rparMin, _rparMax, _zMin, _zMax;
    int _lMin, _lMax;
    std::vector<int> _kept;
    bf_data *_device = nullptr;
    bool _deviceHasGrid = false;
};
typedef std::shared_ptr<AbsCorrelationData> AbsCorrelationDataPtr;

class ComovingCorrelationData : public AbsCorrelationData {
public:
    enum CoordinateSystem { PolarCoordinates, CartesianCoordinates, MultipoleCoordinates };
    ComovingCorrelationData(BinnedGrid grid, CoordinateSystem coordinateSystem);
    double getRadius(int index) const override;    // ComovingCorrelationData.cc:46-55
    double getCosAngle(int index) const override;  // ComovingCorrelationData.cc:57-69
    int getMultipole(int index) const override;    // ComovingCorrelationData.cc:71-79
    double getRedshift(int index) const override;  // ComovingCorrelationData.cc:81-84

private:
    CoordinateSystem _coordinateSystem;
};

// Stand-in for cosmo::LambdaCdmRadiationUniverse (src/baofit.cc:413; the dependency is not in the
// tree): flat or curved LCDM with photons + massless neutrinos,
//   E(z)^2 = Or (1+z)^4 + Om (1+z)^3 + Ok (1+z)^2 + (1 - Om - Ok - Or),
//   Or h^2 = 2.469e-5 (Tcmb/2.725)^4 (1 + 0.2271 Nnu)            UNVERIFIED (dep)
// Distances in Mpc/h.
class LambdaCdmRadiationUniverse {
public:
    LambdaCdmRadiationUniverse(double OmegaMatter, double OmegaK, double hubbleConstant, double Tcmb = 2.725,
                               double Nnu = 3.04);
    double getHubbleFunction(double z) const;                  // E(z)
    double getLineOfSightComovingDistance(double z) const;     // Mpc/h
    double getTransverseComovingScale(double z) const;         // Mpc/h per radian
    double omegaRadiation() const { return _OmegaRadiation; }
    void setOmegaRadiation(double v) { _OmegaRadiation = v; }

private:
    double _OmegaMatter, _OmegaK, _h, _OmegaRadiation;
};
typedef std::shared_ptr<LambdaCdmRadiationUniverse> AbsHomogeneousUniversePtr;

// baofit/QuasarCorrelationData.{h,cc}: observing coordinates (log(lam2/lam1), angular separation in
// arcmin, mean redshift) converted to (r, mu) through the cosmology (:139-152), with the extra
// ll / separation cuts of finalize() (:102-137).
class QuasarCorrelationData : public AbsCorrelationData {
public:
    QuasarCorrelationData(BinnedGrid grid, double llMin, double llMax, double sepMin, double sepMax, bool fixCov,
                          AbsHomogeneousUniversePtr cosmology);
    double getRadius(int index) const override;
    double getCosAngle(int index) const override;
    double getRedshift(int index) const override;
    void transform(double ll, double sep, double dsep, double z, double &r, double &mu) const;

protected:
    bool extraCut(int index) const override;

private:
    double _llMin, _llMax, _sepMin, _sepMax, _arcminToRad;
    AbsHomogeneousUniversePtr _cosmology;
    mutable int _lastIndex = -1;
    mutable double _rLast = 0, _muLast = 0, _zLast = 0;
    void _setIndex(int index) const;
};

// "<name>.data" / ".cov" / ".icov" text loaders (AbsCorrelationData.cc:156-277); ".cov.gz" is not
// handled here (tests decompress fixtures first).
void loadCorrelationData(std::string const &dataName, AbsCorrelationData &data, bool icov = false, bool weighted = false,
                         bool customGrid = false);

// Stand-in for likely::BinnedDataResampler as baofit uses it (CorrelationAnalyzer.cc:42-75 addData/getCombined,
// :213-260 the jackknife / bootstrap / each samplers): keeps the individual observations of a platelist and
// builds inverse-covariance weighted combinations of subsets of them,
//   Cinv = sum_k n_k Cinv_k,   Cinv.d = sum_k n_k Cinv_k.d_k     (n_k = how often observation k was drawn).
// With fixCovariance the covariance of a bootstrap sample accounts for repeated observations,
//   C_fixed = C (sum_k n_k^2 Cinv_k) C   (the variance of the weighted mean above)        UNVERIFIED (dep).
// Every sample is a fresh data set from the prototype factory (grid, final cuts), not finalized.
class BinnedDataResampler {
public:
    typedef std::function<AbsCorrelationDataPtr()> Factory;
    explicit BinnedDataResampler(Factory prototype) : _prototype(prototype) {}
    void addObservation(AbsCorrelationDataPtr observation);  // not finalized; congruent with the first one
    int getNObservations() const { return (int)_icov.size(); }
    AbsCorrelationDataPtr combined() const;
    // observations left after dropping the seqno-th ndrop-subset (subsets in lexicographic order of the dropped
    // indices); null once seqno is past the last subset
    AbsCorrelationDataPtr jackknife(int ndrop, long seqno) const;
    static long jackknifeSize(int nobs, int ndrop);
    // size draws with replacement (0: as many as there are observations) from a counter-based generator keyed by
    // (seed, trial), so that trials can be sharded over ranks
    AbsCorrelationDataPtr bootstrap(int size, bool fixCovariance, unsigned long long seed, long trial) const;
    std::vector<int> bootstrapCounts(int size, unsigned long long seed, long trial) const;
    AbsCorrelationDataPtr getObservationCopy(int index) const;
    AbsCorrelationDataPtr combine(std::vector<int> const &counts, bool fixCovariance) const;

private:
    Factory _prototype;
    std::vector<int> _bins;
    AbsCorrelationDataPtr _first;
    std::vector<std::vector<double>> _icov, _icovData;  // per observation: dense Cinv over the bins with data, Cinv.d
};
typedef std::shared_ptr<BinnedDataResampler> BinnedDataResamplerPtr;

// ---- fitter -----------------------------------------------------------------------------------
class CorrelationFitter {
public:
    // CorrelationFitter.cc:20-42
    CorrelationFitter(AbsCorrelationDataPtr data, AbsCorrelationModelPtr model, int covSampleSize = 0);
    void setErrorScale(double scale);
    // CorrelationFitter.cc:53-72
    void getPrediction(likely::Parameters const &params, std::vector<double> &prediction) const;
    // CorrelationFitter.cc:74-86: (0.5*icovScale*chi2 + priors)/errorScale
    double operator()(likely::Parameters const &params) const;
    // The batched entry point: fval[b] for B parameter vectors in one CUDA pass; dataOverride is
    // empty or B alternative data vectors (toy MC).
    void evaluateBatch(std::vector<likely::Parameters> const &params, std::vector<double> &fval,
                       std::vector<std::vector<double>> const *dataOverride = nullptr) const;
    // Same against device-resident toy realisations (bf_toys): point b is compared with toy toyIndex[b].
    void evaluateBatchToys(std::vector<likely::Parameters> const &params, std::vector<double> &fval, bf_toys const *toys,
                           std::vector<int> const &toyIndex) const;
    // chi2 of B flat parameter vectors (params[B*npar]); NaN where the reference would have thrown.
    // toys / toyIndex (per point) optional.
    void chiSquareFlat(std::vector<double> const &params, int B, std::vector<double> &chi2, bf_toys const *toys = nullptr,
                       const int *toyIndex = nullptr) const;
    // Gauss-Newton normal equations of nfit probe groups of G points (bf_gram_batch); bad[f] = some point invalid
    void gramFlat(std::vector<double> const &params, int nfit, int G, std::vector<double> &gram, std::vector<char> &bad,
                  bf_toys const *toys = nullptr, const int *toyIndexPerFit = nullptr) const;
    // the same with the probe columns generated on the GPU from centres / steps (bf_gram_probes)
    void gramProbes(std::vector<double> const &centres, std::vector<double> const &steps, int nfit, int nprobe, std::vector<double> &gram,
                    std::vector<char> &bad, bf_toys const *toys = nullptr, const int *toyIndexPerFit = nullptr) const;
    // the same probes reduced on the device to the normal equations (bf_normal_probes); the returned buffer is
    // page-locked, owned by the calling thread and valid until its next call: [(nprobe+1)^2 + nprobe] per fit
    const double *normalProbes(std::vector<double> const &centres, std::vector<double> const &steps, int nfit, int nprobe,
                               std::vector<char> &bad, bf_toys const *toys = nullptr, const int *toyIndexPerFit = nullptr) const;
    // Runs many LM fits in lock step (Jacobian rounds through gramFlat, trial rounds through chiSquareFlat);
    // toyOfFit (optional) gives the realisation each fit is taken against.
    void runLockStepLM(std::vector<likely::LMFit> &fits, bf_toys const *toys = nullptr, std::vector<int> const *toyOfFit = nullptr) const;
    double errorScale() const { return _errorScale; }
    void predictBatch(std::vector<likely::Parameters> const &params, std::vector<double> &pred /*[n][B]*/) const;
    // CorrelationFitter.cc:88-92; method is accepted for interface parity ("mn2::vmetric")
    likely::FunctionMinimumPtr fit(std::string const &methodName = "mn2::vmetric", std::string const &config = "") const;
    // CorrelationFitter.cc:109-123 (likely::MarkovChainEngine "saunter" is one Metropolis chain; here nwalkers
    // independent Metropolis chains advance in lock step, one chi-square batch per step): appends nchain
    // samples of npar values + fval, one every `interval` steps of a walker, proposals drawn from the
    // covariance of fminStart scaled by 2.38/sqrt(nfloat).
    void mcmc(likely::FunctionMinimumPtr fminStart, int nchain, int interval, std::vector<double> &samples, int nwalkers = 256,
              unsigned long long seed = 1966) const;
    AbsCorrelationModelPtr model() const { return _model; }
    AbsCorrelationDataPtr data() const { return _data; }
    double icovScale() const { return _icovScale; }
    // the covariance of the data times `factor` (toy studies with toymc-scale): chi2 terms are weighted by icovScale / factor
    void scaleCovariance(double factor) { _icovScale /= factor; }

private:
    AbsCorrelationDataPtr _data;
    AbsCorrelationModelPtr _model;
    double _errorScale, _icovScale;
};

// ---- analyzer (the callers either side of the hot path) ----------------------------------------
struct ScanResult {
    likely::FunctionMinimumPtr best;          // unconstrained fit
    std::vector<std::string> axisNames;       // scanned parameters, in parameter order
    std::vector<std::vector<double>> points;  // [npoints][npar] best-fit parameters at each grid point
    std::vector<double> chi2;                 // 2 * fval at each grid point (0 for failed fits)
    std::vector<int> status;
};

class CorrelationAnalyzer {
public:
    CorrelationAnalyzer(std::string const &method = "mn2::vmetric", double rmin = 0, double rmax = 200,
                        int covSampleSize = 0, bool verbose = false);
    void addData(AbsCorrelationDataPtr data) { _data = data; }
    void setModel(AbsCorrelationModelPtr model) { _model = model; }
    // CorrelationAnalyzer.cc:169-190
    likely::FunctionMinimumPtr fitSample(AbsCorrelationDataPtr sample, std::string const &config = "") const;
    // CorrelationAnalyzer.cc:573-605: refits at every point of the parameter binning grid, last
    // axis fastest; grid points [first, first+count) only (count < 0: all) so that ranks can
    // shard the grid. All fits of the shard run in lock step on the GPU.
    ScanResult parameterScan(AbsCorrelationDataPtr sample, long first = 0, long count = -1) const;
    long scanSize() const;
    // writes the reference's scan.dat layout (CorrelationAnalyzer.cc:379-443)
    static void writeScan(std::string const &path, ScanResult const &scan);
    // CorrelationAnalyzer.cc:337-373 + 452-532: toy MC samples [first, first+count) drawn on the GPU
    // from the truth prediction + L g, each refitted; returns best-fit vectors and 2*fval
    // CorrelationAnalyzer.cc:607-676: one line per kept bin "index c1 c2 c3 r mu|ell z predicted data error
    // [d pred / d p_j ...]"; the gradients are central differences over 0.1 * error_j, all 1 + 2 npar
    // predictions in one batch on the GPU.
    void dumpResiduals(std::ostream &out, likely::FunctionMinimumPtr fmin, std::string const &script = "", bool dumpGradients = false) const;
    // CorrelationAnalyzer.cc:678-716: "r mono quad hexa" at ndump radii between rmin and rmax at redshift zdump
    // (< 0: z of the first kept bin), from AbsCorrelationModel::evaluate(r, multipole, z, ...).
    void dumpModel(std::ostream &out, likely::FitParameters parameters, int ndump, double zdump = -1, std::string const &script = "",
                   bool oneLine = false) const;
    // the same multipoles for many parameter vectors in one GPU batch: [points][3 * ndump]
    std::vector<double> modelMultipoles(std::vector<likely::Parameters> const &points, int ndump, double zdump = -1) const;
    // SamplingOutput (CorrelationAnalyzer.cc:375-450): the file toy-MC, bootstrap, jackknife and scan analyses save
    void writeSamples(std::string const &path, likely::FunctionMinimumPtr fmin, std::vector<std::vector<double>> const &fitted,
                      std::vector<double> const &chi2, int nsave = 0, double zsave = -1,
                      likely::FunctionMinimumPtr fmin2 = likely::FunctionMinimumPtr(),
                      std::vector<std::vector<double>> const *refitted = nullptr, std::vector<double> const *rechi2 = nullptr) const;
    // refitConfig (the reference's refit-config, CorrelationAnalyzer.cc:491-503): every sample is fitted a second time
    // with this script applied to the initial parameters; results go to refitted / rechi2 / restatus.
    // "toymc-scale" (src/baofit.cc:294, CorrelationAnalyzer.cc:349-357): the covariance of the toy prototype times this factor, for the
    // noise and, as the reference's code has it, for the fits of the toys (their chi2 is 1/scale of the unscaled one)
    void setToyMCScale(double varianceScale);
    void doToyMCSampling(long first, long count, unsigned long long seed, std::vector<std::vector<double>> &fitted,
                         std::vector<double> &chi2, std::vector<int> &status, std::string const &toymcConfig = "",
                         std::string const &refitConfig = "", std::vector<std::vector<double>> *refitted = nullptr,
                         std::vector<double> *rechi2 = nullptr, std::vector<int> *restatus = nullptr,
                         likely::FitParameters const *truthBase = nullptr, std::string const &saveFirstName = "") const;
    // truthBase: the parameters the truth vector is predicted from before toymcConfig is applied -- the reference takes those of the
    // fit of the combined sample (fmin, CorrelationAnalyzer.cc:361-369); null: the model's initial parameters (what fmin is with
    // no-initial-fit). saveFirstName ("toymc-save", :281-286): the first realisation of the study (toy 0) in saveData's format.
    // CorrelationAnalyzer.cc:301-335 (doJackknifeAnalysis, doBootstrapAnalysis, fitEach) + the loop of
    // doSamplingAnalysis :452-532: every sample is a new data vector AND covariance, refitted from the model's
    // initial parameters. kind = "jackknife" (n = observations dropped per sample), "bootstrap" (n = trials,
    // size = draws per trial, 0: number of observations), "each" (one sample per observation). Samples
    // [first, first+count) of the sequence, count < 0: to its end. Returns the number of samples fitted.
    void setResampler(BinnedDataResamplerPtr resampler) { _resampler = resampler; }
    BinnedDataResamplerPtr resampler() const { return _resampler; }
    // CorrelationAnalyzer::compareEach (CorrelationAnalyzer.cc:77-125; "compare-each" / "compare-each-final" of src/baofit.cc:636-645):
    // every observation against the combined data, before (finalized = false) or after the final cuts. Per observation:
    // log|C_k| / n, chi2 = sum over the eigenmodes of C_k - C_ref of (v_m . (d_k - d_ref))^2 / lambda_m, and its chi-square
    // probability for n degrees of freedom; the per-mode contributions go to the file (empty name: no file),
    // "<k> <log|C|/n> <chi2> <mode 0> ... <mode n-1>" with modes in ascending eigenvalue order.
    void compareEach(std::string const &saveName, bool finalized, std::vector<double> &prob, std::vector<double> &chi2,
                     std::vector<double> &logDet) const;
    // CorrelationAnalyzer::estimateCombinedCovariance (CorrelationAnalyzer.cc:733-746; "bootstrap-cov-trials", src/baofit.cc:736-756):
    // sample covariance, over nSamples bootstrap draws of the observations, of the combined (unfinalized) data vector; dense over
    // the bins with data (returned in bins). With icovName, its inverse is written in the loader's "i j value" format when the
    // estimate is positive definite ("bs.icov"). Returns whether it is. Draws: bootstrapCounts(0, seed, trial).
    bool estimateCombinedCovariance(int nSamples, unsigned long long seed, std::string const &icovName, std::vector<int> &bins,
                                    std::vector<double> &cov) const;
    long resamplingSize(std::string const &kind, int n) const;
    AbsCorrelationDataPtr resample(std::string const &kind, int n, int size, bool fixCovariance, unsigned long long seed, long seqno) const;
    long doResampling(std::string const &kind, int n, int size, bool fixCovariance, unsigned long long seed, long first, long count,
                      std::vector<std::vector<double>> &fitted, std::vector<double> &chi2, std::vector<int> &status,
                      std::string const &refitConfig = "", std::vector<std::vector<double>> *refitted = nullptr,
                      std::vector<double> *rechi2 = nullptr, std::vector<int> *restatus = nullptr) const;

private:
    std::string _method;
    double _rmin, _rmax;
    int _covSampleSize;
    double _toyScale = 1.0;
    bool _verbose;
    AbsCorrelationDataPtr _data;
    AbsCorrelationModelPtr _model;
    BinnedDataResamplerPtr _resampler;
};

// ---- .ini driven construction -----------------------------------------------------------------
struct Options {
    std::map<std::string, std::string> values;       // last value wins
    std::vector<std::string> modelConfig;             // composing, in order
    std::set<std::string> flags;                      // bool switches that are on
    bool has(std::string const &k) const { return values.count(k) > 0; }
    std::string str(std::string const &k, std::string const &def = "") const;
    double num(std::string const &k, double def) const;
    bool flag(std::string const &k) const { return flags.count(k) > 0; }
};
// Parses an INI text (key = value, '#' comments; "yes/true/1" switches; model-config composes).
Options parseIni(std::string const &text);
// keys outside the reference's option list, and options that select what this build does not have, throw (called by parseIni)
void checkOptions(Options const &options);
// src/baofit.cc:415-464: builds the model the options describe (r-space or --kspace).
AbsCorrelationModelPtr createModel(Options const &opt, std::string const &baseDir);
// src/baofit.cc:475-628 (comoving formats): builds, loads, cuts and finalizes the data set.
// With plateroot/platelist the individual observations are also handed to *resampler (if given).
AbsCorrelationDataPtr createData(Options const &opt, std::string const &baseDir, BinnedDataResamplerPtr *resampler = nullptr);

}  // namespace baofit
