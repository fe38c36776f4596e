// Host-side model classes: parameter layout + bf_model_desc; no arithmetic of the hot path.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "baofit_host.h"

namespace baofit {

// ------------------------------------------------------------------------------------------
// device binding
// ------------------------------------------------------------------------------------------
namespace {
bf_ctx *g_ctx = nullptr;
int g_device = 0;
void check(int rc, const char *what) {
    if (rc != BF_OK) throw RuntimeError(std::string(what) + ": " + bf_last_error());
}
}  // namespace

void setDevice(int device) {
    if (g_ctx && device != g_device) throw RuntimeError("setDevice: the device context already exists");
    g_device = device;
}
bf_ctx *deviceContext() {
    if (!g_ctx) check(bf_ctx_create(g_device, &g_ctx), "bf_ctx_create");
    return g_ctx;
}
void releaseDeviceContext() {
    if (g_ctx) bf_ctx_destroy(g_ctx);
    g_ctx = nullptr;
}

// ------------------------------------------------------------------------------------------
// file readers
// ------------------------------------------------------------------------------------------
static void readTwoColumns(std::string const &path, std::vector<double> &x, std::vector<double> &y) {
    std::ifstream in(path.c_str());
    if (!in.good()) throw RuntimeError("Unable to open " + path);
    std::string line;
    x.clear();
    y.clear();
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        double a, b;
        if (ss >> a >> b) { x.push_back(a); y.push_back(b); }
    }
    if (x.size() < 3) throw RuntimeError("too few rows in " + path);
}

// ------------------------------------------------------------------------------------------
// AbsCorrelationModel
// ------------------------------------------------------------------------------------------
AbsCorrelationModel::AbsCorrelationModel(std::string const &name) : _name(name) {
    std::memset(&_desc, 0, sizeof(_desc));
    _desc.idx_beta = _desc.idx_bb = _desc.idx_gamma_bias = _desc.idx_gamma_beta = -1;
    _desc.idx_delta_v = _desc.idx_bias2 = _desc.idx_bb2 = _desc.idx_beta_bias = -1;
    _desc.idx_bao = _desc.idx_comb_scale = _desc.idx_rad = _desc.idx_nl = _desc.idx_cont = _desc.idx_bs = -1;
    _desc.idx_hcd = _desc.idx_uv = _desc.idx_smgaus = _desc.idx_smlor = _desc.idx_nlcorr = _desc.idx_metal = -1;
    _desc.omega_matter = 0.27;  // AbsCorrelationModel.cc:16
}

AbsCorrelationModel::~AbsCorrelationModel() {
    if (_device) bf_model_destroy(_device);
}

int AbsCorrelationModel::defineParameter(std::string const &name, double value, double error) {
    if (_device) throw RuntimeError("AbsCorrelationModel: cannot define parameters after the model is in use.");
    likely::FitParameter p;
    p.name = name;
    p.value = value;
    p.error = error;
    _parameters.push_back(p);
    return (int)_parameters.size() - 1;
}

void AbsCorrelationModel::configureFitParameters(std::string const &script) {
    try {
        likely::modifyFitParameters(_parameters, script);
    } catch (likely::RuntimeError const &e) {
        throw RuntimeError(std::string("configureFitParameters: ") + e.what());
    }
}

int AbsCorrelationModel::findParameter(std::string const &name) const {
    for (size_t i = 0; i < _parameters.size(); ++i)
        if (_parameters[i].name == name) return (int)i;
    return -1;
}

void AbsCorrelationModel::setCoordinates(std::vector<double> rbin, std::vector<double> mubin, std::vector<double> zbin) {
    if (rbin.size() != mubin.size() || rbin.size() != zbin.size())
        throw RuntimeError("AbsCorrelationModel::setCoordinates: coordinate vectors not the same size.");
    _rbin = rbin;
    _mubin = mubin;
    _zbin = zbin;
}

void AbsCorrelationModel::_setZRef(double zref) {
    if (zref < 0) throw RuntimeError("AbsCorrelationModel: expected zref >= 0.");
    _desc.zref = zref;
}
void AbsCorrelationModel::_setOmegaMatter(double OmegaMatter) {
    if (OmegaMatter < 0) throw RuntimeError("AbsCorrelationModel: expected OmegaMatter >= 0.");
    _desc.omega_matter = OmegaMatter;
}

int AbsCorrelationModel::_defineLinearBiasParameters(double zref, bool crossCorrelation, bool combinedBias) {
    if (_desc.idx_beta >= 0) throw RuntimeError("AbsCorrelationModel: linear bias parameters already defined.");
    _setZRef(zref);
    _desc.idx_beta = defineParameter("beta", 1.4, 0.1);
    _desc.idx_bb = defineParameter("(1+beta)*bias", -0.336, 0.03);
    _desc.idx_gamma_bias = defineParameter("gamma-bias", 3.8, 0.3);
    int last = _desc.idx_gamma_beta = defineParameter("gamma-beta", 0, 0.1);
    if (crossCorrelation) {
        _desc.flags |= BF_FLAG_CROSS_CORRELATION;
        _desc.idx_delta_v = defineParameter("delta-v", 0, 10);
        _desc.idx_bias2 = defineParameter("bias2", 3.6, 0.1);
        last = _desc.idx_bb2 = defineParameter("beta2*bias2", 1, 0.05);
    }
    if (combinedBias) {
        _desc.flags |= BF_FLAG_COMBINED_BIAS;
        last = _desc.idx_beta_bias = defineParameter("beta*bias", -0.196, 0.02);
        configureFitParameters("fix[(1+beta)*bias]=0");
    }
    return last;
}

const double *AbsCorrelationModel::_keep(std::vector<double> const &v) {
    _arrays.push_back(v);
    return _arrays.back().data();
}
const int *AbsCorrelationModel::_keep(std::vector<int> const &v) {
    _iarrays.push_back(v);
    return _iarrays.back().data();
}

bf_model *AbsCorrelationModel::deviceModel() {
    if (!_device) {
        finalizeDescription();
        _desc.npar = getNParameters();
        check(bf_model_create(deviceContext(), &_desc, &_device), "bf_model_create");
    }
    return _device;
}

double AbsCorrelationModel::evaluate(double r, double mu, double z, likely::Parameters const &params, int index) {
    if ((int)params.size() != getNParameters()) throw RuntimeError("AbsCorrelationModel::evaluate: wrong number of parameters.");
    bf_data_desc dd;
    std::memset(&dd, 0, sizeof(dd));
    double one = 1.0, zero = 0.0;
    int idx = index < 0 ? 0 : index;
    dd.n_data = 1;
    dd.r = &r; dd.mu = &mu; dd.z = &z; dd.index = &idx; dd.data = &zero; dd.cov = &one;
    if (_desc.dist_matrix && !_rbin.empty()) {
        dd.n_grid = (int)_rbin.size();
        dd.grid_r = _rbin.data(); dd.grid_mu = _mubin.data(); dd.grid_z = _zbin.data();
    }
    bf_data *d = nullptr;
    check(bf_data_create(deviceContext(), &dd, &d), "bf_data_create");
    double pred = 0;
    int status = 0;
    int rc = bf_eval_batch(deviceContext(), deviceModel(), d, params.data(), 1, &pred, 1, &status);
    bf_data_destroy(d);
    check(rc, "bf_eval_batch");
    if (status & BF_STATUS_DILATION_MIN) throw RuntimeError("BaoKSpaceCorrelationModel: hit min dilation limit.");
    if (status & BF_STATUS_DILATION_MAX) throw RuntimeError("BaoKSpaceCorrelationModel: hit max dilation limit.");
    return pred;
}

void AbsCorrelationModel::printToStream(std::ostream &out) const {
    out << _name << " with " << _parameters.size() << " parameters:" << std::endl;
    for (size_t i = 0; i < _parameters.size(); ++i) {
        auto const &p = _parameters[i];
        char buf[160];
        std::snprintf(buf, sizeof(buf), "  [%2d] %-28s = %12.6f +/- %10.6f%s", (int)i, p.name.c_str(), p.value, p.error,
                      p.floating ? "" : " (fixed)");
        out << buf << std::endl;
    }
    out << "Reference redshift = " << _desc.zref << std::endl;
}

// ------------------------------------------------------------------------------------------
// BroadbandModel (grammar of BroadbandModel.cc:21-90)
// ------------------------------------------------------------------------------------------
static bool parseAxis(std::string const &s, int out[3]) {
    std::vector<int> v;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ':')) {
        if (tok.empty()) return false;
        char *end = nullptr;
        long x = std::strtol(tok.c_str(), &end, 10);
        if (*end) return false;
        v.push_back((int)x);
    }
    if (s.empty() || s.back() == ':' || v.empty() || v.size() > 3) return false;
    if (v.size() == 1) { v.push_back(v[0]); v.push_back(1); }
    else if (v.size() == 2) v.push_back(1);
    out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
    return true;
}

BroadbandModel::BroadbandModel(std::string const &name, std::string const &tag, std::string const &paramSpec,
                               double r0, double z0, AbsCorrelationModel *base) {
    (void)name;
    if (!base) throw RuntimeError("BroadbandModel: needs a parent model.");
    int ax[5][3] = {{0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}};  // r, mu, rP, rT, z
    struct Form { const char *tag; std::vector<int> axes; };
    const Form forms[] = {{"r,mu,rP,rT,z=", {0, 1, 2, 3, 4}}, {"rP,rT,z=", {2, 3, 4}}, {"rP,rT=", {2, 3}},
                          {"r,mu,z=", {0, 1, 4}}, {"r,mu=", {0, 1}}};
    std::string body = paramSpec;
    std::vector<int> axes = {0, 1, 4};  // the r,mu,z tag is optional
    for (auto const &f : forms)
        if (paramSpec.compare(0, std::strlen(f.tag), f.tag) == 0) { body = paramSpec.substr(std::strlen(f.tag)); axes = f.axes; break; }
    std::vector<std::string> parts;
    {
        std::stringstream ss(body);
        std::string tok;
        while (std::getline(ss, tok, ',')) parts.push_back(tok);
        if (!body.empty() && body.back() == ',') parts.push_back("");
    }
    bool ok = parts.size() == axes.size();
    for (size_t i = 0; ok && i < parts.size(); ++i) ok = parseAxis(parts[i], ax[axes[i]]);
    if (!ok) throw RuntimeError("BroadbandModel: badly formatted parameter specification: " + paramSpec);
    std::memset(&_spec, 0, sizeof(_spec));
    _spec.enabled = 1;
    _spec.r0 = r0;
    _spec.z0 = z0;
    int rmin = ax[0][0], rmax = ax[0][1], rstep = ax[0][2];
    if (rmax < rmin || rstep == 0) throw RuntimeError("BroadbandModel: illegal r-parameter specification.");
    if (rstep < 0) {
        _spec.r_denom = -rstep;
        rmin *= _spec.r_denom;
        rmax *= _spec.r_denom;
        rstep = 1;
    } else
        _spec.r_denom = 1;
    _spec.r_min = rmin; _spec.r_max = rmax; _spec.r_step = rstep;
    _spec.mu_min = ax[1][0]; _spec.mu_max = ax[1][1]; _spec.mu_step = ax[1][2];
    if (_spec.mu_max < _spec.mu_min || _spec.mu_step <= 0) throw RuntimeError("BroadbandModel: illegal mu-parameter specification.");
    if (_spec.mu_min < 0 || _spec.mu_max > 8)
        throw RuntimeError("BroadbandModel: only multipoles 0-8 are allowed in mu-parameter specification.");
    _spec.rp_min = ax[2][0]; _spec.rp_max = ax[2][1]; _spec.rp_step = ax[2][2];
    if (_spec.rp_max < _spec.rp_min || _spec.rp_step <= 0) throw RuntimeError("BroadbandModel: illegal rP-parameter specification.");
    _spec.rt_min = ax[3][0]; _spec.rt_max = ax[3][1]; _spec.rt_step = ax[3][2];
    if (_spec.rt_max < _spec.rt_min || _spec.rt_step <= 0) throw RuntimeError("BroadbandModel: illegal rT-parameter specification.");
    _spec.z_min = ax[4][0]; _spec.z_max = ax[4][1]; _spec.z_step = ax[4][2];
    if (_spec.z_max < _spec.z_min || _spec.z_step <= 0) throw RuntimeError("BroadbandModel: illegal z-parameter specification.");
    bool first = true;
    for (int zi = _spec.z_min; zi <= _spec.z_max; zi += _spec.z_step)
        for (int mi = _spec.mu_min; mi <= _spec.mu_max; mi += _spec.mu_step)
            for (int ri = _spec.r_min; ri <= _spec.r_max; ri += _spec.r_step)
                for (int pi = _spec.rp_min; pi <= _spec.rp_max; pi += _spec.rp_step)
                    for (int ti = _spec.rt_min; ti <= _spec.rt_max; ti += _spec.rt_step) {
                        char pname[200];  // BroadbandModel.cc:150
                        std::snprintf(pname, sizeof(pname), "%s z%d mu%d r%+d rP%d rT%d", tag.c_str(), zi, mi, ri, pi, ti);
                        int index = base->defineParameter(pname, 0, 1e-3);
                        if (first) { _spec.idx_base = index; first = false; }
                    }
}

// ------------------------------------------------------------------------------------------
// DistortionMatrix
// ------------------------------------------------------------------------------------------
DistortionMatrix::DistortionMatrix(std::string const &distMatrixName, int distMatrixOrder, bool verbose)
    : _nbins(distMatrixOrder) {
    if (_nbins <= 0) throw RuntimeError("DistortionMatrix: expected distortion matrix order > 0.");
    _dist.assign((size_t)_nbins * _nbins, 0.0);
    std::string distName = distMatrixName + ".dmat";
    std::ifstream in(distName.c_str());
    if (!in.good()) throw RuntimeError("DistortionMatrix: Unable to open " + distName);
    std::string line;
    int lines = 0;
    while (std::getline(in, line)) {
        lines++;
        std::istringstream ss(line);
        int i1, i2;
        double value;
        if (!(ss >> i1 >> i2 >> value))
            throw RuntimeError("DistortionMatrix: error reading line " + std::to_string(lines) + " of " + distName);
        if (i1 < 0 || i2 < 0 || i1 >= _nbins || i2 >= _nbins)
            throw RuntimeError("DistortionMatrix: invalid indices on line " + std::to_string(lines) + " of " + distName);
        _dist[(size_t)i1 * _nbins + i2] = value;
    }
    if (verbose)
        std::cout << "Read " << lines << " of " << (size_t)_nbins * _nbins << " distortion matrix values from " << distName << std::endl;
}

DistortionMatrix::DistortionMatrix(std::vector<double> const &dense, int order) : _nbins(order), _dist(dense) {
    if (_nbins <= 0) throw RuntimeError("DistortionMatrix: expected distortion matrix order > 0.");
    if (_dist.size() != (size_t)order * order) throw RuntimeError("DistortionMatrix: wrong matrix size.");
}

double DistortionMatrix::getDistortion(int index1, int index2) const {
    if (index1 < 0 || index2 < 0 || index1 >= _nbins || index2 >= _nbins)
        throw RuntimeError("DistortionMatrix::getDistortion: invalid indices.");
    return _dist[(size_t)index1 * _nbins + index2];
}

// ------------------------------------------------------------------------------------------
// MetalCorrelationModel
// ------------------------------------------------------------------------------------------
static int readIndexValueColumn(std::string const &path, std::vector<double> &out, int column, int ncols) {
    std::ifstream in(path.c_str());
    if (!in.good()) throw RuntimeError("MetalCorrelationModel: Unable to open " + path);
    std::string line;
    int lines = 0;
    out.clear();
    while (std::getline(in, line)) {
        lines++;
        std::istringstream ss(line);
        int index;
        std::vector<double> v(ncols);
        bool ok = bool(ss >> index);
        for (int c = 0; ok && c < ncols; ++c) ok = bool(ss >> v[c]);
        if (!ok) throw RuntimeError("MetalCorrelationModel: error reading line " + std::to_string(lines) + " of " + path);
        out.push_back(v[column]);
    }
    return lines;
}

// Bicubic grid file. likely::createBiCubicInterpolator's on-disk layout is not in the tree
// (UNVERIFIED dep); ours: first line "nx ny x0 y0 dx dy", then ny*nx values, x fastest.
static void readBicubicGrid(std::string const &path, int &nx, int &ny, double &x0, double &y0, double &dx, double &dy,
                            std::vector<double> &values) {
    std::ifstream in(path.c_str());
    if (!in.good()) throw RuntimeError("MetalCorrelationModel: error while reading metal model interpolation data.");
    if (!(in >> nx >> ny >> x0 >> y0 >> dx >> dy) || nx < 2 || ny < 2)
        throw RuntimeError("MetalCorrelationModel: error while reading metal model interpolation data.");
    values.resize((size_t)nx * ny);
    for (auto &v : values)
        if (!(in >> v)) throw RuntimeError("MetalCorrelationModel: error while reading metal model interpolation data.");
}

MetalCorrelationModel::MetalCorrelationModel(std::string const &metalModelName, bool metalModel,
                                             bool metalModelInterpolate, bool metalCIV, bool toyMetal,
                                             bool crossCorrelation, AbsCorrelationModel *base)
    : _kind(BF_METAL_NONE), _indexBase(-1) {
    if (!base) throw RuntimeError("MetalCorrelationModel: needs a parent model.");
    if ((metalModel && metalModelInterpolate) || (metalModel && toyMetal) || (metalModelInterpolate && toyMetal))
        throw RuntimeError("MetalCorrelationModel: illegal option specification.");
    if ((metalModel && crossCorrelation) || (metalModelInterpolate && !crossCorrelation) || (toyMetal && crossCorrelation))
        throw RuntimeError("MetalCorrelationModel: illegal option for cross-correlation.");
    if (metalModel || metalModelInterpolate) {
        _indexBase = base->defineParameter("beta Si2a", 1, 0.1);
        base->defineParameter("bias Si2a", -0.01, 0.001);
        base->defineParameter("beta Si2b", 1, 0.1);
        base->defineParameter("bias Si2b", -0.01, 0.001);
        base->defineParameter("beta Si2c", 1, 0.1);
        base->defineParameter("bias Si2c", -0.01, 0.001);
        base->defineParameter("beta Si3", 1, 0.1);
        base->defineParameter("bias Si3", -0.01, 0.001);
        if (metalCIV) {
            base->defineParameter("beta CIV", 1, 0.1);
            base->defineParameter("bias CIV", -0.01, 0.001);
        }
        std::vector<std::string> l1, l2;
        if (metalModel) {
            _kind = BF_METAL_AUTO_TABLE;
            _nmet = 5;
            l1 = {"_Lya", "_Lya", "_Lya", "_Lya", "_Si2a", "_Si2a", "_Si2a", "_Si2b", "_Si2b", "_Si2c", "_Si3", "_Si3", "_Si3", "_Si3"};
            l2 = {"_Si2a", "_Si2b", "_Si2c", "_Si3", "_Si2a", "_Si2b", "_Si2c", "_Si2b", "_Si2c", "_Si2c", "_Si2a", "_Si2b", "_Si2c", "_Si3"};
            _p1 = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 4, 4, 4, 4};
            _p2 = {1, 2, 3, 4, 1, 2, 3, 2, 3, 3, 1, 2, 3, 4};
            if (metalCIV) {
                _nmet += 1;
                for (const char *s : {"_Lya", "_Si2a", "_Si2b", "_Si2c", "_Si3", "_CIVa"}) { l1.push_back(s); l2.push_back("_CIVa"); }
                for (int i = 0; i < 6; ++i) { _p1.push_back(i); _p2.push_back(5); }
            }
        } else {
            _kind = BF_METAL_CROSS_BICUBIC;
            _nmet = 5;
            l1 = {"_QSO", "_QSO", "_QSO", "_QSO"};
            l2 = {"_Si2a", "_Si2b", "_Si2c", "_Si3"};
            _p1 = {0, 0, 0, 0};
            _p2 = {1, 2, 3, 4};
            if (metalCIV) { _nmet += 1; l1.push_back("_QSO"); l2.push_back("_CIVa"); _p1.push_back(0); _p2.push_back(5); }
        }
        _ncomb = (int)_p1.size();
        int lastLines = -1;
        for (int i = 0; i < _ncomb; ++i) {
            for (int ell = 0; ell <= 4; ell += 2) {
                std::string fn = metalModelName + l1[i] + l2[i] + "." + std::to_string(ell) + ".dat";
                if (_kind == BF_METAL_AUTO_TABLE) {
                    std::vector<double> col;
                    int lines = readIndexValueColumn(fn, col, 0, 1);
                    if (lastLines >= 0 && lastLines != lines) throw RuntimeError("MetalCorrelationModel: deviating number of lines in " + fn);
                    lastLines = lines;
                    _auto.insert(_auto.end(), col.begin(), col.end());
                } else {
                    int nx, ny;
                    double x0, y0, dx, dy;
                    std::vector<double> g;
                    readBicubicGrid(fn, nx, ny, x0, y0, dx, dy, g);
                    if (_cross.empty()) { _nx = nx; _ny = ny; _x0 = x0; _y0 = y0; _dx = dx; _dy = dy; }
                    else if (nx != _nx || ny != _ny) throw RuntimeError("MetalCorrelationModel: error while reading metal model interpolation data.");
                    _cross.insert(_cross.end(), g.begin(), g.end());
                }
            }
            std::vector<double> zg;
            int lines = readIndexValueColumn(metalModelName + l1[i] + l2[i] + ".grid", zg, 2, 3);
            if (lastLines >= 0 && lastLines != lines)
                throw RuntimeError("MetalCorrelationModel: deviating number of lines in " + metalModelName + l1[i] + l2[i] + ".grid");
            lastLines = lines;
            _zgrid.insert(_zgrid.end(), zg.begin(), zg.end());
        }
        _ngrid = lastLines;
    } else if (toyMetal) {
        _kind = BF_METAL_TOY;
        _indexBase = base->defineParameter("metal ampl0", 1, 0.1);
        base->defineParameter("metal ampl1", 1, 0.1);
        base->defineParameter("metal ampl2", 1, 0.1);
        base->defineParameter("metal ampl3", 1, 0.1);
        base->defineParameter("metal width0", 6, 0.6);
    }
}

void MetalCorrelationModel::fill(bf_model_desc &d, AbsCorrelationModel &owner) const {
    d.metal_kind = _kind;
    d.idx_metal = _indexBase;
    if (_kind == BF_METAL_AUTO_TABLE || _kind == BF_METAL_CROSS_BICUBIC) {
        d.metal_ncomb = _ncomb;
        d.metal_nmet = _nmet;
        d.metal_p1 = owner._keep(_p1);
        d.metal_p2 = owner._keep(_p2);
        d.metal_ngrid = _ngrid;
        d.metal_zgrid = owner._keep(_zgrid);
        if (_kind == BF_METAL_AUTO_TABLE) d.metal_auto = owner._keep(_auto);
        else {
            d.metal_nx = _nx; d.metal_ny = _ny; d.metal_x0 = _x0; d.metal_y0 = _y0; d.metal_dx = _dx; d.metal_dy = _dy;
            d.metal_cross = owner._keep(_cross);
        }
    }
}

// ------------------------------------------------------------------------------------------
// BaoCorrelationModel (BaoCorrelationModel.cc:18-86)
// ------------------------------------------------------------------------------------------
static std::string withSlash(std::string root) {
    if (!root.empty() && root.back() != '/') root += '/';
    return root;
}

BaoCorrelationModel::BaoCorrelationModel(std::string const &modelrootName, std::string const &fiducialName,
                                         std::string const &nowigglesName, std::string const &metalModelName,
                                         std::string const &distAdd, std::string const &distMul, double distR0,
                                         double zref, double OmegaMatter, bool anisotropic, bool decoupled,
                                         bool metalModel, bool metalModelInterpolate, bool metalCIV, bool toyMetal,
                                         bool combinedBias, bool combinedScale, bool crossCorrelation)
    : AbsCorrelationModel("BAO Correlation Model") {
    _desc.kind = BF_MODEL_RSPACE;
    _setOmegaMatter(OmegaMatter);
    if (anisotropic) _desc.flags |= BF_FLAG_ANISOTROPIC;
    if (decoupled) _desc.flags |= BF_FLAG_DECOUPLED;
    _defineLinearBiasParameters(zref, crossCorrelation, combinedBias);
    _desc.idx_bao = defineParameter("BAO amplitude", 1, 0.15);
    defineParameter("BAO alpha-iso", 1, 0.02);
    defineParameter("BAO alpha-parallel", 1, 0.1);
    defineParameter("BAO alpha-perp", 1, 0.1);
    defineParameter("gamma-scale", 0, 0.5);
    _desc.idx_rad = defineParameter("Rad strength", 0., 0.1);
    defineParameter("Rad anisotropy", 0., 0.1);
    defineParameter("Rad mean free path", 200., 10.);
    defineParameter("Rad quasar lifetime", 10., 0.1);
    configureFitParameters("fix[Rad*]");
    if (combinedScale) {
        _desc.flags |= BF_FLAG_COMBINED_SCALE;
        _desc.idx_comb_scale = defineParameter("BAO aperp/apar", 1, 0.1);
        configureFitParameters("fix[BAO alpha-perp]");
    }
    std::string root = withSlash(modelrootName);
    try {
        std::vector<double> r0, r2, r4, y0, y2, y4;
        readTwoColumns(root + fiducialName + ".0.dat", r0, y0);
        readTwoColumns(root + fiducialName + ".2.dat", r2, y2);
        readTwoColumns(root + fiducialName + ".4.dat", r4, y4);
        if (r0 != r2 || r0 != r4) throw RuntimeError("multipole tables must share knots");
        _desc.n_fid = (int)r0.size();
        _desc.fid_r = _keep(r0); _desc.fid_xi0 = _keep(y0); _desc.fid_xi2 = _keep(y2); _desc.fid_xi4 = _keep(y4);
        readTwoColumns(root + nowigglesName + ".0.dat", r0, y0);
        readTwoColumns(root + nowigglesName + ".2.dat", r2, y2);
        readTwoColumns(root + nowigglesName + ".4.dat", r4, y4);
        if (r0 != r2 || r0 != r4) throw RuntimeError("multipole tables must share knots");
        _desc.n_nw = (int)r0.size();
        _desc.nw_r = _keep(r0); _desc.nw_xi0 = _keep(y0); _desc.nw_xi2 = _keep(y2); _desc.nw_xi4 = _keep(y4);
    } catch (RuntimeError const &) {
        throw RuntimeError("BaoCorrelationModel: error while reading model interpolation data.");
    }
    if (metalModel || metalModelInterpolate || toyMetal) {
        _metalCorr.reset(new MetalCorrelationModel(metalModelName, metalModel, metalModelInterpolate, metalCIV, toyMetal,
                                                   crossCorrelation, this));
        _metalCorr->fill(_desc, *this);
    }
    if (!distAdd.empty()) {
        _distortAdd.reset(new BroadbandModel("Additive broadband distortion", "dist add", distAdd, distR0, zref, this));
        _desc.bb[BF_BB_DIST_ADD] = _distortAdd->spec();
    }
    if (!distMul.empty()) {
        _distortMul.reset(new BroadbandModel("Multiplicative broadband distortion", "dist mul", distMul, distR0, zref, this));
        _desc.bb[BF_BB_DIST_MUL] = _distortMul->spec();
    }
}

// ------------------------------------------------------------------------------------------
// BaoKSpaceCorrelationModel (BaoKSpaceCorrelationModel.cc:26-211)
// ------------------------------------------------------------------------------------------
BaoKSpaceCorrelationModel::BaoKSpaceCorrelationModel(
    std::string const &modelrootName, std::string const &fiducialName, std::string const &nowigglesName,
    std::string const &distMatrixName, std::string const &metalModelName, double zref, double OmegaMatter, double rmin,
    double rmax, double dilmin, double dilmax, double relerr, double abserr, int ellMax, int samplesPerDecade,
    std::string const &distAdd, std::string const &distMul, double distR0, double zeff, double sigma8, double dzmin,
    int distMatrixOrder, std::string const &distMatrixDistAdd, std::string const &distMatrixDistMul, bool anisotropic,
    bool decoupled, bool nlBroadband, bool nlCorrection, bool fitNLCorrection, bool nlCorrectionAlt, bool binSmooth,
    bool binSmoothAlt, bool hcdModel, bool hcdModelAlt, bool uvfluctuation, bool radiationModel, bool smoothGauss,
    bool smoothLorentz, bool distMatrix, bool metalModel, bool metalModelInterpolate, bool metalCIV, bool toyMetal,
    bool combinedBias, bool combinedScale, bool crossCorrelation, bool verbose)
    : AbsCorrelationModel("BAO k-Space Correlation Model") {
    (void)relerr;
    (void)abserr;  // the Filon quadrature has no tolerance knobs
    _desc.kind = BF_MODEL_KSPACE;
    _setZRef(zref);
    _setOmegaMatter(OmegaMatter);
    auto on = [&](bool b, unsigned f) { if (b) _desc.flags |= f; };
    on(anisotropic, BF_FLAG_ANISOTROPIC); on(decoupled, BF_FLAG_DECOUPLED); on(nlBroadband, BF_FLAG_NL_BROADBAND);
    on(nlCorrection, BF_FLAG_NL_CORRECTION); on(fitNLCorrection, BF_FLAG_FIT_NL_CORRECTION);
    on(nlCorrectionAlt, BF_FLAG_NL_CORRECTION_ALT); on(binSmooth, BF_FLAG_BIN_SMOOTH); on(binSmoothAlt, BF_FLAG_BIN_SMOOTH_ALT);
    on(hcdModel, BF_FLAG_HCD_MODEL); on(hcdModelAlt, BF_FLAG_HCD_MODEL_ALT); on(uvfluctuation, BF_FLAG_UVFLUCTUATION);
    on(radiationModel, BF_FLAG_RADIATION_MODEL); on(smoothGauss, BF_FLAG_SMOOTH_GAUSS); on(smoothLorentz, BF_FLAG_SMOOTH_LORENTZ);
    on(combinedBias, BF_FLAG_COMBINED_BIAS); on(combinedScale, BF_FLAG_COMBINED_SCALE); on(crossCorrelation, BF_FLAG_CROSS_CORRELATION);
    _desc.idx_beta = defineParameter("beta", 1.4, 0.1);
    _desc.idx_bb = defineParameter("(1+beta)*bias", -0.336, 0.03);
    _desc.idx_gamma_bias = defineParameter("gamma-bias", 3.8, 0.3);
    _desc.idx_gamma_beta = defineParameter("gamma-beta", 0, 0.1);
    if (crossCorrelation) {
        _desc.idx_delta_v = defineParameter("delta-v", 0, 10);
        _desc.idx_bias2 = defineParameter("bias2", 3.6, 0.1);
        _desc.idx_bb2 = defineParameter("beta2*bias2", 1, 0.05);
    }
    _desc.idx_nl = defineParameter("SigmaNL-perp", 3.26, 0.3);
    defineParameter("1+f", 2, 0.1);
    _desc.idx_bao = defineParameter("BAO amplitude", 1, 0.15);
    defineParameter("BAO alpha-iso", 1, 0.02);
    defineParameter("BAO alpha-parallel", 1, 0.1);
    defineParameter("BAO alpha-perp", 1, 0.1);
    defineParameter("gamma-scale", 0, 0.5);
    if (binSmooth || binSmoothAlt) { _desc.idx_bs = defineParameter("bin scale-par", 4, 0.4); defineParameter("bin scale-perp", 4, 0.4); }
    if (hcdModel || hcdModelAlt) {
        _desc.idx_hcd = defineParameter("HCD beta", 1, 0.1);
        defineParameter("HCD rel bias", 0.1, 0.01);
        defineParameter("HCD scale", 20, 2);
    }
    if (uvfluctuation) {
        _desc.idx_uv = defineParameter("UV rel bias", -1, 0.1);
        defineParameter("UV abs resp bias", -0.667, 0.06);
        defineParameter("UV mean free path", 300, 30);
    }
    if (radiationModel) {
        _desc.idx_rad = defineParameter("Rad strength", 0, 0.1);
        defineParameter("Rad anisotropy", 0, 0.1);
        defineParameter("Rad mean free path", 200, 10);
        defineParameter("Rad quasar lifetime", 10, 0.1);
    }
    if (smoothGauss) _desc.idx_smgaus = defineParameter("smooth scale gauss", 1, 0.1);
    if (smoothLorentz) _desc.idx_smlor = defineParameter("smooth scale lorentz", 1, 0.1);
    if (combinedBias) {
        _desc.idx_beta_bias = defineParameter("beta*bias", -0.196, 0.02);
        configureFitParameters("fix[(1+beta)*bias]=0");
    }
    if (combinedScale) {
        _desc.idx_comb_scale = defineParameter("BAO aperp/apar", 1, 0.1);
        configureFitParameters("fix[BAO alpha-perp]");
    }
    std::string root = withSlash(modelrootName);
    try {
        std::vector<double> k1, p1, k2, p2;
        readTwoColumns(root + fiducialName + "_matterpower.dat", k1, p1);
        readTwoColumns(root + nowigglesName + "_matterpower.dat", k2, p2);
        if (k1 != k2) throw RuntimeError("P(k) tables must share k nodes");
        _desc.n_pk = (int)k1.size();
        _desc.pk_k = _keep(k1); _desc.pk_fid = _keep(p1); _desc.pk_nw = _keep(p2);
    } catch (RuntimeError const &) {
        throw RuntimeError("BaoKSpaceCorrelationModel: error while reading model interpolation data.");
    }
    if (rmin >= rmax) throw RuntimeError("BaoKSpaceCorrelationModel: expected rmin < rmax.");
    if (rmin <= 0) throw RuntimeError("BaoKSpaceCorrelationModel: expected rmin > 0.");
    if (dilmin > dilmax) throw RuntimeError("BaoKSpaceCorrelationModel: expected dilmin <= dilmax.");
    if (dilmin <= 0) throw RuntimeError("BaoKSpaceCorrelationModel: expected dilmin > 0.");
    _desc.rmin = rmin; _desc.rmax = rmax; _desc.dilmin = dilmin; _desc.dilmax = dilmax;
    _desc.samples_per_decade = samplesPerDecade; _desc.ell_max = ellMax; _desc.zeff = zeff; _desc.sigma8 = sigma8; _desc.dzmin = dzmin;
    // NonLinearCorrectionModel registers its parameters here (NonLinearCorrectionModel.cc:41-44)
    if (fitNLCorrection) { _desc.idx_nlcorr = defineParameter("nlcorr qnl", 0.867, 0.08); defineParameter("nlcorr kp", 19.4, 1); }
    if (metalModel || metalModelInterpolate || toyMetal) {
        _metalCorr.reset(new MetalCorrelationModel(metalModelName, metalModel, metalModelInterpolate, metalCIV, toyMetal,
                                                   crossCorrelation, this));
        _metalCorr->fill(_desc, *this);
    }
    if (distMatrix) {
        if (!distMatrixName.empty()) {
            _distMat.reset(new DistortionMatrix(distMatrixName, distMatrixOrder, verbose));
            _desc.dist_matrix_order = _distMat->order();
            _desc.dist_matrix = _distMat->dense().data();
        }
        if (!distMatrixDistAdd.empty()) {
            _distMatDistortAdd.reset(new BroadbandModel("Additive broadband distortion before distortion matrix",
                                                        "dist Matrix dist add", distMatrixDistAdd, distR0, zref, this));
            _desc.bb[BF_BB_DM_DIST_ADD] = _distMatDistortAdd->spec();
        }
        if (!distMatrixDistMul.empty()) {
            _distMatDistortMul.reset(new BroadbandModel("Multiplicative broadband distortion before distortion matrix",
                                                        "dist Matrix dist mul", distMatrixDistMul, distR0, zref, this));
            _desc.bb[BF_BB_DM_DIST_MUL] = _distMatDistortMul->spec();
        }
    }
    if (!distAdd.empty()) {
        _distortAdd.reset(new BroadbandModel("Additive broadband distortion", "dist add", distAdd, distR0, zref, this));
        _desc.bb[BF_BB_DIST_ADD] = _distortAdd->spec();
    }
    if (!distMul.empty()) {
        _distortMul.reset(new BroadbandModel("Multiplicative broadband distortion", "dist mul", distMul, distR0, zref, this));
        _desc.bb[BF_BB_DIST_MUL] = _distortMul->spec();
    }
}

// baofit/BaoKSpaceFftCorrelationModel.cc:21-113
BaoKSpaceFftCorrelationModel::BaoKSpaceFftCorrelationModel(
    std::string const &modelrootName, std::string const &fiducialName, std::string const &nowigglesName, double zref,
    double OmegaMatter, double spacing, int nx, int ny, int nz, std::string const &distAdd, std::string const &distMul,
    double distR0, double zcorr0, double zcorr1, double zcorr2, double sigma8, bool anisotropic, bool decoupled,
    bool nlBroadband, bool nlCorrection, bool fitNLCorrection, bool nlCorrectionAlt, bool distortionAlt,
    bool noDistortion, bool crossCorrelation, bool verbose, double planeRMax)
    : AbsCorrelationModel("BAO k-Space FFT Correlation Model") {
    _desc.kind = BF_MODEL_KSPACE_FFT;
    _setZRef(zref);
    _setOmegaMatter(OmegaMatter);
    auto on = [&](bool b, unsigned f) { if (b) _desc.flags |= f; };
    on(anisotropic, BF_FLAG_ANISOTROPIC); on(decoupled, BF_FLAG_DECOUPLED); on(nlBroadband, BF_FLAG_NL_BROADBAND);
    on(nlCorrection, BF_FLAG_NL_CORRECTION); on(fitNLCorrection, BF_FLAG_FIT_NL_CORRECTION);
    on(nlCorrectionAlt, BF_FLAG_NL_CORRECTION_ALT); on(distortionAlt, BF_FLAG_DISTORTION_ALT);
    on(noDistortion, BF_FLAG_NO_DISTORTION); on(crossCorrelation, BF_FLAG_CROSS_CORRELATION);
    // :39-62
    _desc.idx_beta = defineParameter("beta", 1.4, 0.1);
    _desc.idx_bb = defineParameter("(1+beta)*bias", -0.336, 0.03);
    _desc.idx_gamma_bias = defineParameter("gamma-bias", 3.8, 0.3);
    _desc.idx_gamma_beta = defineParameter("gamma-beta", 0, 0.1);
    if (crossCorrelation) {
        _desc.idx_delta_v = defineParameter("delta-v", 0, 10);
        _desc.idx_bias2 = defineParameter("bias2", 3.6, 0.1);
        _desc.idx_bb2 = defineParameter("beta2*bias2", 1, 0.05);
    }
    _desc.idx_nl = defineParameter("SigmaNL-perp", 3.26, 0.3);
    defineParameter("1+f", 2, 0.1);
    _desc.idx_cont = defineParameter("cont-kc", 0.02, 0.002);
    defineParameter("cont-pc", 1, 0.1);
    _desc.idx_bao = defineParameter("BAO amplitude", 1, 0.15);
    defineParameter("BAO alpha-iso", 1, 0.02);
    defineParameter("BAO alpha-parallel", 1, 0.1);
    defineParameter("BAO alpha-perp", 1, 0.1);
    defineParameter("gamma-scale", 0, 0.5);
    std::string root = withSlash(modelrootName);
    try {
        std::vector<double> k1, p1, k2, p2;
        readTwoColumns(root + fiducialName + "_matterpower.dat", k1, p1);
        readTwoColumns(root + nowigglesName + "_matterpower.dat", k2, p2);
        if (k1 != k2) throw RuntimeError("P(k) tables must share k nodes");
        _desc.n_pk = (int)k1.size();
        _desc.pk_k = _keep(k1); _desc.pk_fid = _keep(p1); _desc.pk_nw = _keep(p2);
    } catch (RuntimeError const &) {
        throw RuntimeError("BaoKSpaceFftCorrelationModel: error while reading tabulated P(k) data.");
    }
    _desc.fft_spacing = spacing; _desc.fft_nx = nx; _desc.fft_ny = ny; _desc.fft_nz = nz; _desc.fft_rmax = planeRMax;
    _desc.zcorr0 = zcorr0; _desc.zcorr1 = zcorr1; _desc.zcorr2 = zcorr2; _desc.sigma8 = sigma8;
    if (verbose) {
        // the reference prints the size of its 3-D grids here (:93-96); ours are never allocated
        std::cout << "3D FFT memory size = " << (double)nx * ny * nz * 16 / 1048576. << " Mb (not allocated: the transform is"
                  << " evaluated on the z = 0 plane only)" << std::endl;
    }
    if (!distAdd.empty()) {
        _distortAdd.reset(new BroadbandModel("Additive broadband distortion", "dist add", distAdd, distR0, zref, this));
        _desc.bb[BF_BB_DIST_ADD] = _distortAdd->spec();
    }
    if (!distMul.empty()) {
        _distortMul.reset(new BroadbandModel("Multiplicative broadband distortion", "dist mul", distMul, distR0, zref, this));
        _desc.bb[BF_BB_DIST_MUL] = _distortMul->spec();
    }
    // :110-112 the non-linear correction model is constructed last
    if (fitNLCorrection) { _desc.idx_nlcorr = defineParameter("nlcorr qnl", 0.867, 0.08); defineParameter("nlcorr kp", 19.4, 1); }
}

void BaoKSpaceCorrelationModel::setDistortionMatrix(std::shared_ptr<DistortionMatrix> dm) {
    _distMat = dm;
    _desc.dist_matrix_order = dm->order();
    _desc.dist_matrix = dm->dense().data();
}

}  // namespace baofit
