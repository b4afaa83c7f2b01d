// host_capi.cpp -- flat C entry points around the C++ host mirror (jaicov_host.hpp) so that it can be driven from the
// parity tests (ctypes) and from non-C++ callers: a scene arrives as plain arrays, the network is built through the mirror's
// own class API (Camera, Image.add, DistortionModel.add, ScaleBar, DirectlyObservedParameterGroup, BundleAdjustment.add), and
// the integer bookkeeping / the adjustment results come back as arrays.  Built into libjaicov_host.so next to
// libjaicov_b200.so (bundle-adjustment_b200/build.sh); nothing here computes on the CPU.
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "jaicov_host.hpp"

using namespace jaicov::host;

namespace {

struct Network {
    std::vector<std::unique_ptr<ObjectCoordinate>> points;
    std::vector<std::unique_ptr<Camera>> cameras;
    std::vector<std::vector<UnknownParameter *>> coefs;   // per camera, in the order the scene lists them
    std::vector<Image *> images;                          // camera -> image order
    std::vector<std::unique_ptr<ScaleBar>> bars;
    std::vector<std::unique_ptr<ObservationParameter>> observations;
    std::vector<std::unique_ptr<DirectlyObservedParameterGroup>> groups;
    BundleAdjustment adjustment;
    std::unique_ptr<FlatProblem> flat;                    // jhost_get_flat: the problem is indexed once (like the reference, a second
                                                          // prepareUnknownParameters on the same graph finds every column already set)
    std::string error;
};

DistortionModel::Type model_of(int parameter_type) {
    switch (parameter_type) {
        case 141: case 142: return DistortionModel::Type::AFFINITY_AND_SHEAR;
        case 131: case 132: case 133: return DistortionModel::Type::TANGENTIAL_DISTORTION;
        case 121: return DistortionModel::Type::RADIAL_DISTORTION;
        case 151: return DistortionModel::Type::DISTANCE_DISTORTION;
        case 161: return DistortionModel::Type::ZERNIKE_X;
        case 162: return DistortionModel::Type::ZERNIKE_Y;
        case 163: return DistortionModel::Type::ZERNIKE_GRADIENT;
    }
    throw std::invalid_argument("unknown distortion parameter type " + std::to_string(parameter_type));
}

}  // namespace

extern "C" {

void *jhost_create(void) { return new Network(); }
void jhost_destroy(void *n) { delete static_cast<Network *>(n); }
const char *jhost_last_error(void *n) { return static_cast<Network *>(n)->error.c_str(); }

#define JH_BEGIN Network &N = *static_cast<Network *>(n); try {
#define JH_END } catch (const std::exception &e) { N.error = e.what(); return -1; } return 0;

// object points: xyz[3 n], fixed[3 n] (0 / 1), datum[n]
int jhost_add_points(void *n, int count, const double *xyz, const uint8_t *fixed, const uint8_t *datum) {
    JH_BEGIN
    for (int i = 0; i < count; i++) {
        N.points.emplace_back(new ObjectCoordinate(std::to_string(N.points.size()), xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
        ObjectCoordinate &oc = *N.points.back();
        for (int c = 0; c < 3; c++)
            if (fixed && fixed[3 * i + c]) oc.component(c).setColumn(COL_FIXED);
        oc.setDatum(datum ? datum[i] != 0 : true);
    }
    JH_END
}

// one camera with its coefficients (ParameterType id, order, value, fixed) in any order; returns the camera index through *index
int jhost_add_camera(void *n, double r0, const double *io_val, const uint8_t *io_fixed, int n_coef, const int32_t *coef_type,
                     const int32_t *coef_order, const double *coef_val, const uint8_t *coef_fixed, int *index) {
    JH_BEGIN
    std::vector<DistortionModel::Type> types;
    for (int k = 0; k < n_coef; k++) {
        const DistortionModel::Type t = model_of(coef_type[k]);
        if (std::find(types.begin(), types.end(), t) == types.end()) types.push_back(t);
    }
    N.cameras.emplace_back(new Camera((int)N.cameras.size() + 1, r0, types));
    Camera &cam = *N.cameras.back();
    for (int i = 0; i < 3; i++) {
        cam.getInteriorOrientation().at(i).setValue(io_val[i]);
        cam.getInteriorOrientation().at(i).setColumn(io_fixed[i] ? COL_FIXED : COL_UNSET);
    }
    std::vector<UnknownParameter *> listed;
    for (int k = 0; k < n_coef; k++) {
        DistortionModel *m = cam.getDistortionModel(model_of(coef_type[k]));
        UnknownParameter *p;
        switch (coef_type[k]) {
            case 141: p = &m->getCx(); break;
            case 142: p = &m->getCy(); break;
            case 132: p = &m->getBx(); break;
            case 133: p = &m->getBy(); break;
            default: p = &m->add(coef_order[k]);
        }
        p->setValue(coef_val[k]);
        p->setColumn(coef_fixed[k] ? COL_FIXED : COL_UNSET);
        listed.push_back(p);
    }
    N.coefs.push_back(listed);
    N.adjustment.add(&cam);
    if (index) *index = (int)N.cameras.size() - 1;
    JH_END
}

// one image of camera `camera` with m observations of points obj[m] (indices into the point list)
int jhost_add_image(void *n, int camera, const double *eo_val, const uint8_t *eo_fixed, int64_t m, const int32_t *obj, const double *xy,
                    const double *sigma, const double *rho) {
    JH_BEGIN
    Image &img = N.cameras.at(camera)->add((int)N.images.size() + 1);
    N.images.push_back(&img);
    for (int i = 0; i < 6; i++) {
        img.getExteriorOrientation().at(i).setValue(eo_val[i]);
        img.getExteriorOrientation().at(i).setColumn(eo_fixed[i] ? COL_FIXED : COL_UNSET);
    }
    for (int64_t j = 0; j < m; j++) img.add(N.points.at(obj[j]).get(), xy[2 * j], xy[2 * j + 1], sigma[2 * j], sigma[2 * j + 1], rho ? rho[j] : 0.0);
    JH_END
}

int jhost_add_scale_bar(void *n, int a, int b, double length, double sigma) {
    JH_BEGIN
    N.bars.emplace_back(new ScaleBar(N.points.at(a).get(), N.points.at(b).get(), length, sigma));
    N.adjustment.add(N.bars.back().get());
    JH_END
}

// one directly observed group: kind 0 = point (index, comp), 1 = interior orientation (camera, comp), 2 = coefficient (camera,
// position in the camera's LISTED coefficients), 3 = exterior orientation (image in camera -> image order, comp);
// var (may be NULL when dispersion_packed is given), dispersion_packed (may be NULL)
int jhost_add_group(void *n, int r, const int32_t *kind, const int32_t *index, const int32_t *comp, const double *obs, const double *var,
                    const double *dispersion_packed) {
    JH_BEGIN
    std::vector<ObservationParameter *> ops;
    for (int i = 0; i < r; i++) {
        UnknownParameter *ref;
        if (kind[i] == 0) ref = &N.points.at(index[i])->component(comp[i]);
        else if (kind[i] == 1) ref = &N.cameras.at(index[i])->getInteriorOrientation().at(comp[i]);
        else if (kind[i] == 2) ref = N.coefs.at(index[i]).at(comp[i]);
        else ref = &N.images.at(index[i])->getExteriorOrientation().at(comp[i]);
        N.observations.emplace_back(new ObservationParameter(ref));
        N.observations.back()->setValue(obs[i]);
        if (var) N.observations.back()->setVariance(var[i]);
        ops.push_back(N.observations.back().get());
    }
    if (dispersion_packed)
        N.groups.emplace_back(new DirectlyObservedParameterGroup(ops, std::vector<double>(dispersion_packed, dispersion_packed + (size_t)r * (r + 1) / 2)));
    else
        N.groups.emplace_back(new DirectlyObservedParameterGroup(ops));
    N.adjustment.add(N.groups.back().get());
    JH_END
}

int jhost_configure(void *n, int invert_mode, int estimation_type, int max_iterations, int use_centroid, int apply_aposteriori, double damping,
                    int device, int solver) {
    JH_BEGIN
    N.adjustment.setInvertNormalEquation((MatrixInversion)invert_mode);
    N.adjustment.setEstimationType(estimation_type == 1 ? EstimationType::SIMULATION : EstimationType::L2NORM);
    N.adjustment.setMaximalNumberOfIterations(max_iterations);
    N.adjustment.useCentroidedCoordinates(use_centroid != 0);
    N.adjustment.applyAposterioriVarianceOfUnitWeight(apply_aposteriori != 0);
    N.adjustment.setLevenbergMarquardtDampingValue(damping);
    N.adjustment.setDevice(device);
    N.adjustment.setSolver(solver);
    JH_END
}

// prepareUnknownParameters + detectRankDefect.  counts[6] = observations, unknowns, unset interior-orientation parameters, unset
// distortion parameters, datum defect d, object points in the adjustment; flags[7]; *sigma2 = a-priori variance factor
int jhost_prepare(void *n, int32_t *counts, int32_t *flags, double *sigma2) {
    JH_BEGIN
    N.adjustment.prepareUnknownParameters();
    const BundleAdjustment &a = N.adjustment;
    counts[0] = a.getNumberOfObservations(); counts[1] = a.getNumberOfUnknownParameters(); counts[2] = a.getNumberOfInteriorOrientationParameters();
    counts[3] = a.getNumberOfDistortionParameters(); counts[4] = a.getNumberOfDatumConditions(); counts[5] = (int32_t)a.getObjectCoordinates().size();
    a.getRankDefect().flags(flags);
    *sigma2 = a.getVarianceFactorApriori();
    JH_END
}

// The flattened problem exactly as estimateModel() hands it to the C ABI (test access): sizes[8] = cameras, coefficients, images,
// image points, object points, scale bars, groups, unknowns; call once with NULL arrays for the sizes, then with buffers.
// point_of_scene[i] (may be NULL) = position of scene point i in the flat point list, -1 if the point is not part of the problem.
int jhost_get_flat(void *n, int64_t *sizes, double *io_val, int32_t *io_col, int32_t *coef_ptr, int32_t *coef_type, int32_t *coef_order,
                   double *coef_val, int32_t *coef_col, int32_t *cam_of_img, double *eo_val, int32_t *eo_col, int64_t *pt_ptr, int32_t *obj_idx,
                   double *xy, double *var, double *rho, double *xyz, int32_t *pt_col, uint8_t *is_datum, int32_t *bar_a, int32_t *bar_b,
                   int32_t *free_flags, int32_t *point_of_scene) {
    JH_BEGIN
    if (!N.flat) N.flat.reset(new FlatProblem(N.adjustment.prepareUnknownParameters()));
    const FlatProblem &f = *N.flat;
    sizes[0] = (int64_t)f.r0.size(); sizes[1] = (int64_t)f.coef_type.size(); sizes[2] = (int64_t)f.cam_of_img.size();
    sizes[3] = (int64_t)f.obj_idx.size(); sizes[4] = (int64_t)f.xyz.size() / 3; sizes[5] = (int64_t)f.bar_a.size();
    sizes[6] = (int64_t)f.groups.size(); sizes[7] = f.n_unknowns;
    auto put = [](auto *dst, const auto &v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    put(io_val, f.io_val); put(io_col, f.io_col); put(coef_ptr, f.coef_ptr); put(coef_type, f.coef_type); put(coef_order, f.coef_order);
    put(coef_val, f.coef_val); put(coef_col, f.coef_col); put(cam_of_img, f.cam_of_img); put(eo_val, f.eo_val); put(eo_col, f.eo_col);
    put(pt_ptr, f.pt_ptr); put(obj_idx, f.obj_idx); put(xy, f.xy); put(var, f.var); put(rho, f.rho); put(xyz, f.xyz); put(pt_col, f.pt_col);
    put(is_datum, f.is_datum); put(bar_a, f.bar_a); put(bar_b, f.bar_b);
    if (free_flags) std::memcpy(free_flags, f.free_flags, sizeof f.free_flags);
    if (point_of_scene) {
        // flat point p holds the coordinates of exactly one scene point: match by address through the values' owners
        std::unordered_map<const ObjectCoordinate *, int32_t> pos;
        for (size_t p = 0; p < N.adjustment.flatPoints().size(); p++) pos[N.adjustment.flatPoints()[p]] = (int32_t)p;
        for (size_t i = 0; i < N.points.size(); i++) {
            auto it = pos.find(N.points[i].get());
            point_of_scene[i] = it == pos.end() ? -1 : it->second;
        }
    }
    JH_END
}

// group g of the flattened problem (after jhost_get_flat): *r = rows, *has_sigma = 1 if a packed dispersion travels instead of var;
// kind / index / comp / obs / var (r entries each, may be NULL)
int jhost_get_flat_group(void *n, int g, int32_t *r, int32_t *has_sigma, int32_t *kind, int32_t *index, int32_t *comp, double *obs, double *var) {
    JH_BEGIN
    if (!N.flat) throw std::invalid_argument("call jhost_get_flat first");
    const FlatProblem::Group &G = N.flat->groups.at(g);
    *r = (int32_t)G.obs.size();
    *has_sigma = G.sigma.empty() ? 0 : 1;
    auto put = [](auto *dst, const auto &v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    put(kind, G.kind); put(index, G.index); put(comp, G.comp); put(obs, G.obs); put(var, G.var);
    JH_END
}

// runs estimateModel(); returns the EstimationStateType id through *state
int jhost_estimate(void *n, int *state) {
    JH_BEGIN
    *state = (int)N.adjustment.estimateModel();
    if (*state == (int)EstimationStateType::NOT_INITIALISED) N.error = N.adjustment.getLastError();
    JH_END
}

// columns (and values) as they stand in the object graph: points in scene order, per camera x0 y0 c, the coefficients in the
// camera's EVALUATION order (what the C ABI receives), exterior orientations in camera -> image order.  Any pointer may be NULL.
int jhost_get_columns(void *n, int64_t *pt_col, int64_t *io_col, int64_t *coef_col, int64_t *eo_col, double *xyz, double *io_val,
                      double *coef_val, double *eo_val) {
    JH_BEGIN
    for (size_t p = 0; p < N.points.size(); p++)
        for (int c = 0; c < 3; c++) {
            if (pt_col) pt_col[3 * p + c] = N.points[p]->component(c).getColumn();
            if (xyz) xyz[3 * p + c] = N.points[p]->component(c).getValue();
        }
    size_t kc = 0;
    for (size_t ci = 0; ci < N.cameras.size(); ci++) {
        for (int i = 0; i < 3; i++) {
            if (io_col) io_col[3 * ci + i] = N.cameras[ci]->getInteriorOrientation().at(i).getColumn();
            if (io_val) io_val[3 * ci + i] = N.cameras[ci]->getInteriorOrientation().at(i).getValue();
        }
        for (DistortionModel *m : N.cameras[ci]->getDistortionModels())
            for (auto &p : m->parameters()) {
                if (coef_col) coef_col[kc] = p->getColumn();
                if (coef_val) coef_val[kc] = p->getValue();
                kc++;
            }
    }
    for (size_t ii = 0; ii < N.images.size(); ii++)
        for (int i = 0; i < 6; i++) {
            if (eo_col) eo_col[6 * ii + i] = N.images[ii]->getExteriorOrientation().at(i).getColumn();
            if (eo_val) eo_val[6 * ii + i] = N.images[ii]->getExteriorOrientation().at(i).getValue();
        }
    JH_END
}

// results of the last estimateModel(): stats[6] = omega, sigma0^2 a posteriori, sigma0^2 a priori, degrees of freedom, passes, solver used;
// qxx (may be NULL): packed cofactor matrix, (u + d)(u + d + 1) / 2 doubles; returns 1 through *has_qxx if it exists
int jhost_get_results(void *n, double *stats, double *qxx, int *has_qxx) {
    JH_BEGIN
    const BundleAdjustment &a = N.adjustment;
    stats[0] = a.getStatistics().omega; stats[1] = a.getVarianceFactorAposteriori(); stats[2] = a.getVarianceFactorApriori();
    stats[3] = a.getDegreeOfFreedom(); stats[4] = a.getStatistics().iterations; stats[5] = a.getStatistics().solver_used;
    const UpperSymmPackMatrix *Q = a.getCofactorMatrix();
    if (has_qxx) *has_qxx = Q ? 1 : 0;
    if (Q && qxx) std::memcpy(qxx, Q->getData().data(), Q->getData().size() * sizeof(double));
    JH_END
}

// ---- callers either side of the path -----------------------------------------------------------------------------------------
// DirectLinearTransformation over ALL images of the network with the object points known_idx[n_known] as homologous points.
// run == 0: only the gathering of DirectLinearTransformation.java:78-94 (pt_ptr[n_img + 1], xy, xyz, io as handed to jaicov_dlt_batch;
// xy / xyz sized for every observation); run != 0: adjustAll() on the device, out20[20 n_img] = coefficient values in the reference's
// insertion order, ok[n_img].
int jhost_dlt(void *n, int n_known, const int32_t *known_idx, int n_restr, const int32_t *restr, int run, int64_t *pt_ptr, double *xy, double *xyz,
              double *io, double *out20, uint8_t *ok) {
    JH_BEGIN
    std::map<std::string, ObjectCoordinate *> known;
    for (int i = 0; i < n_known; i++) known[N.points.at(known_idx[i])->getName()] = N.points.at(known_idx[i]).get();
    std::vector<std::unique_ptr<DLTCoefficients>> own;
    std::vector<DLTCoefficients *> coefs;
    for (Image *img : N.images) { own.emplace_back(new DLTCoefficients(img)); coefs.push_back(own.back().get()); }
    if (!run) {
        DirectLinearTransformation::Gathered g = DirectLinearTransformation::gather(coefs, known);
        std::memcpy(pt_ptr, g.pt_ptr.data(), g.pt_ptr.size() * sizeof(int64_t));
        if (!g.xy.empty()) std::memcpy(xy, g.xy.data(), g.xy.size() * sizeof(double));
        if (!g.xyz.empty()) std::memcpy(xyz, g.xyz.data(), g.xyz.size() * sizeof(double));
        if (!g.io.empty()) std::memcpy(io, g.io.data(), g.io.size() * sizeof(double));
    } else {
        std::vector<DirectLinearTransformation::RestrictionType> r;
        for (int i = 0; i < n_restr; i++) r.push_back((DirectLinearTransformation::RestrictionType)restr[i]);
        std::vector<bool> res = DirectLinearTransformation::adjustAll(coefs, known, r, 0);
        for (size_t i = 0; i < coefs.size(); i++) {
            ok[i] = res[i] ? 1 : 0;
            for (int k = 0; k < 20; k++) out20[20 * i + k] = coefs[i]->getValue(k);
        }
    }
    JH_END
}

// CoordinateTransformationExteriorOrientation.transform after estimateModel(): the object points pts[n_pts] seen in the images
// imgs[n_imgs] (camera -> image order) are carried into the frame of image `reference`.  *n_out = number of triples; xyz (3 per triple)
// and cov_packed (may be NULL) are filled when cap >= *n_out.
int jhost_transform(void *n, int n_pts, const int32_t *pts, int reference, int n_imgs, const int32_t *imgs, double sigma2, int cap, int *n_out,
                    double *xyz, double *cov_packed) {
    JH_BEGIN
    std::vector<ObjectCoordinate *> oc;
    for (int i = 0; i < n_pts; i++) oc.push_back(N.points.at(pts[i]).get());
    std::vector<Image *> list;
    for (int i = 0; i < n_imgs; i++) list.push_back(N.images.at(imgs[i]));
    CoordinateTransformationExteriorOrientation t;
    t.transform(oc, {{N.images.at(reference), list}}, sigma2, N.adjustment);
    const int r = (int)t.getTransformedCoordinates().size();
    *n_out = r;
    if (cap >= r) {
        for (int k = 0; k < r; k++) {
            xyz[3 * k] = t.getTransformedCoordinates()[k].x; xyz[3 * k + 1] = t.getTransformedCoordinates()[k].y; xyz[3 * k + 2] = t.getTransformedCoordinates()[k].z;
        }
        if (cov_packed && r) std::memcpy(cov_packed, t.getCovarianceMatrix()->getData().data(), t.getCovarianceMatrix()->getData().size() * sizeof(double));
    }
    JH_END
}

// DefaultResultWriter on the network as it stands (after jhost_prepare: <base>.info; after jhost_estimate: also <base>.cxx)
int jhost_export_default(void *n, const char *base) {
    JH_BEGIN
    DefaultResultWriter w(base);
    w.exportResults(N.adjustment);
    JH_END
}

// test hook: java_format_f of jaicov_host.hpp (the number format of the result writers); returns the length written
int jhost_java_format_f(double v, int width, int precision, int plus, char *out, int cap) {
    const std::string s = java_format_f(v, width, precision, plus != 0);
    if (!out || cap <= (int)s.size()) return -1;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

}  // extern "C"
