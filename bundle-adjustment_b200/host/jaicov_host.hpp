// jaicov_host.hpp -- native (C++17, header-only) host side of the adjustment path above the C ABI.
//
// The reference is compiled Java and no JVM exists in this image, so what stays on the Java side in a real integration
//  is restated here in C++ with the reference's own class and method names, argument meaning and error
// behaviour: the object graph, the integer bookkeeping of prepareUnknownParameters / detectRankDefect, the flattening into
// include/jaicov_b200.h and the getters the result writers read.  All floating-point work of the adjustment happens in
// libjaicov_b200.so on the GPU; nothing here computes a Jacobian, a normal equation or a solve.
// (bundle-adjustment_b200/host.py is the same mirror in Python with bulk array entry points; the parity tests drive both
// against the reference's executed bookkeeping, tests/golden/reference_bookkeeping.npz.)
//
// Reference files (relative to JAICOV/src/org/applied_geodesy/adjustment/):
//   bundle/parameter/ParameterType.java:27-110          ParameterType ids
//   bundle/parameter/UnknownParameter.java:26-53        value + column (-1 unset, Integer.MAX_VALUE fixed)
//   bundle/parameter/ObservationParameter.java:26-64    value, variance (> 0), row
//   bundle/ObjectCoordinate.java:32-110, bundle/ScaleBar.java:30-39
//   bundle/camera/orientation/{Interior,Exterior}Orientation.java   iteration orders :70-79 / :39-45
//   bundle/camera/distortion/*.java                     model types in enum-ordinal order, default-fixed Bx, By, Cx, Cy
//   bundle/camera/{Camera,Image,ImageCoordinate}.java   Camera.java:38-139, Image.java:32-90, ImageCoordinate.java:40-54
//   bundle/parameter/DirectlyObservedParameterGroup.java:37-105
//   defect/RankDefect.java:35-130
//   bundle/BundleAdjustment.java                        add :652-665, prepareUnknownParameters :667-782, detectRankDefect
//                                                       :836-1042, estimateModel :203-387, getters :1048-1118, :1177
#pragma once
#include <algorithm>
#include <array>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/jaicov_b200.h"

namespace jaicov {
namespace host {

constexpr int COL_UNSET = -1;
constexpr int COL_FIXED = 2147483647;   // Integer.MAX_VALUE, parameter/UnknownParameter.java:27

enum class ParameterType : int {
    PRINCIPAL_POINT_X = 111, PRINCIPAL_POINT_Y = 112, PRINCIPAL_DISTANCE = 113,
    RADIAL_POLYNOMIAL_A = 121,
    TANGENTIAL_POLYNOMIAL_B = 131, TANGENTIAL_DISTORTION_Bx = 132, TANGENTIAL_DISTORTION_By = 133,
    AFFINITY_AND_SHEAR_Cx = 141, AFFINITY_AND_SHEAR_Cy = 142,
    DISTANCE_POLYNOMIAL_D = 151,
    ZERNIKE_POLYNOMIAL_X = 161, ZERNIKE_POLYNOMIAL_Y = 162, ZERNIKE_POLYNOMIAL_Z = 163,
    CAMERA_COORDINATE_X = 251, CAMERA_COORDINATE_Y = 252, CAMERA_COORDINATE_Z = 253,
    CAMERA_OMEGA = 261, CAMERA_PHI = 262, CAMERA_KAPPA = 263,
    OBJECT_COORDINATE_X = 311, OBJECT_COORDINATE_Y = 312, OBJECT_COORDINATE_Z = 313,
    IMAGE_COORDINATE_X = 411, IMAGE_COORDINATE_Y = 412,
    SCALE_BAR_LENGTH = 511
};

enum class EstimationStateType : int {   // EstimationStateType.java:24-42
    ERROR_FREE_ESTIMATION = 1, BUSY = 0, INTERRUPT = -1, SINGULAR_MATRIX = -2, ROBUST_ESTIMATION_FAILED = -3, NO_CONVERGENCE = -4,
    NOT_INITIALISED = -5, EXPORT_ADJUSTMENT_RESULTS_FAILED = -6, OUT_OF_MEMORY = -7
};
enum class EstimationType : int { L1NORM = 1, L2NORM = 2, SIMULATION = 3, MODIFIED_UNSCENTED_TRANSFORMATION = 4, SPHERICAL_SIMPLEX_UNSCENTED_TRANSFORMATION = 5 };
enum class MatrixInversion : int { NONE = 0, FULL = 1, PRE_ELIMINATION = 2, REDUCED = 3 };   // BundleAdjustment.java:65-70

class ObjectCoordinate;

class UnknownParameter {
public:
    UnknownParameter(ParameterType type, void *reference = nullptr, double value = 0.0) : type_(type), reference_(reference), value_(value) {}
    virtual ~UnknownParameter() = default;
    ParameterType getParameterType() const { return type_; }
    void *getReference() const { return reference_; }
    double getValue() const { return value_; }
    void setValue(double v) { value_ = v; }
    int getColumn() const { return column_; }
    void setColumn(int c) { column_ = c; }
    virtual int getOrder() const { return 0; }

private:
    ParameterType type_;
    void *reference_;
    double value_;
    int column_ = COL_UNSET;
};

class PolynomialCoefficient : public UnknownParameter {
public:
    PolynomialCoefficient(ParameterType type, void *reference, int order) : UnknownParameter(type, reference), order_(order) {}
    int getOrder() const override { return order_; }

private:
    int order_;
};

class ObservationParameter {
public:
    // observation of `parameter`; value defaults to the parameter's current value (ObservationParameter.java:36-41)
    explicit ObservationParameter(UnknownParameter *parameter) : ref_(parameter), type_(parameter->getParameterType()), value_(parameter->getValue()) {}
    ObservationParameter(UnknownParameter *parameter, double value, double variance) : ObservationParameter(parameter) {
        value_ = value;
        setVariance(variance);
    }
    ObservationParameter(ParameterType type, double value, double variance) : ref_(nullptr), type_(type), value_(value) { setVariance(variance); }
    ParameterType getParameterType() const { return type_; }
    UnknownParameter *getReference() const { return ref_; }
    double getValue() const { return value_; }
    void setValue(double v) { value_ = v; }
    double getVariance() const { return variance_; }
    bool hasVariance() const { return variance_ > 0; }
    void setVariance(double variance) {
        if (!(variance > 0)) throw std::invalid_argument("Error, variance must be positive: " + std::to_string(variance));
        variance_ = variance;
    }
    int getRow() const { return row_; }
    void setRow(int r) { row_ = r; }

private:
    UnknownParameter *ref_;
    ParameterType type_;
    double value_;
    double variance_ = -1.0;
    int row_ = -1;
};

class ObjectCoordinate {
public:
    ObjectCoordinate(std::string name, double x, double y, double z)
        : name_(std::move(name)), X_(ParameterType::OBJECT_COORDINATE_X, this, x), Y_(ParameterType::OBJECT_COORDINATE_Y, this, y),
          Z_(ParameterType::OBJECT_COORDINATE_Z, this, z) {}
    ObjectCoordinate(const ObjectCoordinate &) = delete;
    const std::string &getName() const { return name_; }
    UnknownParameter &getX() { return X_; }
    UnknownParameter &getY() { return Y_; }
    UnknownParameter &getZ() { return Z_; }
    UnknownParameter &component(int c) { return c == 0 ? X_ : (c == 1 ? Y_ : Z_); }
    bool isDatum() const { return datum_; }
    void setDatum(bool d) { datum_ = d; }

private:
    std::string name_;
    UnknownParameter X_, Y_, Z_;
    bool datum_ = true;   // ObjectCoordinate.java:34
};

class ScaleBar {   // ScaleBar.java:34-39: variance = sigma^2
public:
    ScaleBar(ObjectCoordinate *a, ObjectCoordinate *b, double value, double sigma)
        : a_(a), b_(b), length_(ParameterType::SCALE_BAR_LENGTH, value, sigma * sigma) {}
    ObservationParameter &getLength() { return length_; }
    ObjectCoordinate *getObjectCoordinateA() const { return a_; }
    ObjectCoordinate *getObjectCoordinateB() const { return b_; }

private:
    ObjectCoordinate *a_, *b_;
    ObservationParameter length_;
};

class Camera;
class Image;

class InteriorOrientation {   // iterator order x0, y0, c (InteriorOrientation.java:70-79)
public:
    explicit InteriorOrientation(Camera *camera)
        : camera_(camera), x0_(ParameterType::PRINCIPAL_POINT_X, this), y0_(ParameterType::PRINCIPAL_POINT_Y, this),
          c_(ParameterType::PRINCIPAL_DISTANCE, this) {}
    UnknownParameter &getPrinciplePointX() { return x0_; }
    UnknownParameter &getPrinciplePointY() { return y0_; }
    UnknownParameter &getPrincipleDistance() { return c_; }
    Camera *getReference() const { return camera_; }
    UnknownParameter &at(int i) { return i == 0 ? x0_ : (i == 1 ? y0_ : c_); }

private:
    Camera *camera_;
    UnknownParameter x0_, y0_, c_;
};

class ExteriorOrientation {   // X0, Y0, Z0, omega, phi, kappa (ExteriorOrientation.java:39-45)
public:
    explicit ExteriorOrientation(Image *image) : image_(image) {
        static const ParameterType order[6] = {ParameterType::CAMERA_COORDINATE_X, ParameterType::CAMERA_COORDINATE_Y, ParameterType::CAMERA_COORDINATE_Z,
                                               ParameterType::CAMERA_OMEGA, ParameterType::CAMERA_PHI, ParameterType::CAMERA_KAPPA};
        for (int i = 0; i < 6; i++) p_.emplace_back(new UnknownParameter(order[i], this));
    }
    UnknownParameter &get(ParameterType t) {
        for (auto &p : p_)
            if (p->getParameterType() == t) return *p;
        throw std::invalid_argument("not an exterior orientation parameter");
    }
    UnknownParameter &at(int i) { return *p_[i]; }
    Image *getReference() const { return image_; }

private:
    Image *image_;
    std::vector<std::unique_ptr<UnknownParameter>> p_;
};

class DistortionModel {
public:
    enum class Type : int {   // enum ordinal order, camera/distortion/DistortionModel.java:29-37
        AFFINITY_AND_SHEAR = 0, TANGENTIAL_DISTORTION = 1, RADIAL_DISTORTION = 2, DISTANCE_DISTORTION = 3, ZERNIKE_X = 4, ZERNIKE_Y = 5,
        ZERNIKE_GRADIENT = 6
    };
    DistortionModel(Camera *camera, Type type, double r0) : camera_(camera), type_(type), r0_(r0) {
        if (type == Type::AFFINITY_AND_SHEAR) {          // AffinityShearDistortionModel.java:39-40: Cx, Cy, fixed by default
            fixedPair(ParameterType::AFFINITY_AND_SHEAR_Cx, ParameterType::AFFINITY_AND_SHEAR_Cy);
        } else if (type == Type::TANGENTIAL_DISTORTION) {   // TangentialDistortionModel.java:34-42: Bx, By, fixed by default
            fixedPair(ParameterType::TANGENTIAL_DISTORTION_Bx, ParameterType::TANGENTIAL_DISTORTION_By);
        }
    }
    Type getType() const { return type_; }
    Camera *getReference() const { return camera_; }
    double getR0() const { return r0_; }
    int getNumberOfParameters() const { return (int)params_.size(); }
    // the two named parameters of the affinity (Cx, Cy) and tangential (Bx, By) models
    UnknownParameter &getCx() { return named(Type::AFFINITY_AND_SHEAR, 0); }
    UnknownParameter &getCy() { return named(Type::AFFINITY_AND_SHEAR, 1); }
    UnknownParameter &getBx() { return named(Type::TANGENTIAL_DISTORTION, 0); }
    UnknownParameter &getBy() { return named(Type::TANGENTIAL_DISTORTION, 1); }
    // polynomial coefficient of the given order (A_i, B_i, D_i; Zernike: the single index j)
    PolynomialCoefficient &add(int order) {
        if (type_ == Type::AFFINITY_AND_SHEAR) throw std::invalid_argument("Error, the affinity and shear model has no polynomial coefficients");
        if (order <= 0) throw std::invalid_argument("Error, polynomial coefficient order must be a real positive integer. " + std::to_string(order));
        if (orders_.count(order)) throw std::invalid_argument("Error, polynomial coefficient order already exists. " + std::to_string(order));
        ParameterType pt = ParameterType::RADIAL_POLYNOMIAL_A;
        switch (type_) {
            case Type::TANGENTIAL_DISTORTION: pt = ParameterType::TANGENTIAL_POLYNOMIAL_B; break;
            case Type::RADIAL_DISTORTION: pt = ParameterType::RADIAL_POLYNOMIAL_A; break;
            case Type::DISTANCE_DISTORTION: pt = ParameterType::DISTANCE_POLYNOMIAL_D; break;
            case Type::ZERNIKE_X: pt = ParameterType::ZERNIKE_POLYNOMIAL_X; break;
            case Type::ZERNIKE_Y: pt = ParameterType::ZERNIKE_POLYNOMIAL_Y; break;
            default: pt = ParameterType::ZERNIKE_POLYNOMIAL_Z; break;
        }
        orders_.insert(order);
        auto *c = new PolynomialCoefficient(pt, this, order);
        params_.emplace_back(c);
        return *c;
    }
    PolynomialCoefficient *get(int order) {
        for (auto &p : params_)
            if (p->getOrder() == order && dynamic_cast<PolynomialCoefficient *>(p.get())) return static_cast<PolynomialCoefficient *>(p.get());
        return nullptr;
    }
    // insertion order: what the reference's iterator yields
    const std::vector<std::unique_ptr<UnknownParameter>> &parameters() const { return params_; }

private:
    void fixedPair(ParameterType a, ParameterType b) {
        params_.emplace_back(new UnknownParameter(a, this));
        params_.emplace_back(new UnknownParameter(b, this));
        params_[0]->setColumn(COL_FIXED);
        params_[1]->setColumn(COL_FIXED);
    }
    UnknownParameter &named(Type t, int i) {
        if (type_ != t) throw std::invalid_argument("parameter does not belong to this distortion model");
        return *params_[i];
    }
    Camera *camera_;
    Type type_;
    double r0_;
    std::vector<std::unique_ptr<UnknownParameter>> params_;
    std::unordered_set<int> orders_;
};

class ImageCoordinate {   // camera/ImageCoordinate.java:40-54: x, y, sigma^2, rho in the open interval (-1, 1)
public:
    ImageCoordinate(ObjectCoordinate *oc, double xp, double yp, double sigmax, double sigmay, double corrCoefXY)
        : oc_(oc), x_(xp), y_(yp), varx_(sigmax * sigmax), vary_(sigmay * sigmay), rho_(corrCoefXY) {
        if (std::fabs(corrCoefXY) >= 1)
            throw std::invalid_argument("Error, correlation coefficient rho(x,y) must be in the open interval (-1 1): " + std::to_string(corrCoefXY));
        if (!(varx_ > 0) || !(vary_ > 0)) throw std::invalid_argument("Error, variance must be positive");
    }
    ObjectCoordinate *getObjectCoordinate() const { return oc_; }
    double getX() const { return x_; }
    double getY() const { return y_; }
    double getVarianceX() const { return varx_; }
    double getVarianceY() const { return vary_; }
    double getCorrelationCoefficient() const { return rho_; }
    int rowX = -1, rowY = -1;

private:
    ObjectCoordinate *oc_;
    double x_, y_, varx_, vary_, rho_;
};

class Image {   // camera/Image.java:32-90
public:
    Image(int id, Camera *camera) : id_(id), camera_(camera), eo_(this) {}
    int getId() const { return id_; }
    Camera *getReference() const { return camera_; }
    ExteriorOrientation &getExteriorOrientation() { return eo_; }
    int getNumberOfImageCoordinates() const { return (int)coords_.size(); }
    // a second observation of the same object point is ignored (Image.java:56-58)
    ImageCoordinate *add(ObjectCoordinate *oc, double xp, double yp, double sigmax, double sigmay, double corrCoefXY = 0.0) {
        auto it = index_.find(oc);
        if (it != index_.end()) return &coords_[it->second];
        coords_.emplace_back(oc, xp, yp, sigmax, sigmay, corrCoefXY);
        index_[oc] = coords_.size() - 1;
        return &coords_.back();
    }
    ImageCoordinate *get(ObjectCoordinate *oc) {   // Image.java:77-79
        auto it = index_.find(oc);
        return it == index_.end() ? nullptr : &coords_[it->second];
    }
    std::deque<ImageCoordinate> &coordinates() { return coords_; }   // deque: pointers handed out by add() / get() stay valid

private:
    int id_;
    Camera *camera_;
    ExteriorOrientation eo_;
    std::deque<ImageCoordinate> coords_;
    std::unordered_map<ObjectCoordinate *, size_t> index_;
};

class Camera {   // camera/Camera.java:38-139; models are kept in enum-ordinal order (Arrays.sort, :50)
public:
    Camera(int id, double r0, std::vector<DistortionModel::Type> types = {}) : id_(id), r0_(r0), io_(this) {
        std::sort(types.begin(), types.end());
        for (auto t : types) {
            if (models_.count(t)) throw std::invalid_argument("Error, duplicate type of distortion model detected.");
            models_[t].reset(new DistortionModel(this, t, (t == DistortionModel::Type::AFFINITY_AND_SHEAR || t == DistortionModel::Type::TANGENTIAL_DISTORTION) ? 0.0 : r0));
        }
    }
    Camera(const Camera &) = delete;
    int getId() const { return id_; }
    double getR0() const { return r0_; }
    InteriorOrientation &getInteriorOrientation() { return io_; }
    int getNumberOfImages() const { return (int)images_.size(); }
    DistortionModel *getDistortionModel(DistortionModel::Type t) {
        auto it = models_.find(t);
        return it == models_.end() ? nullptr : it->second.get();
    }
    std::vector<DistortionModel *> getDistortionModels() {
        std::vector<DistortionModel *> v;
        for (auto &kv : models_) v.push_back(kv.second.get());   // std::map: ascending enum ordinal
        return v;
    }
    Image &add(int imageId) {   // Camera.java:92-98: one Image per id, insertion order kept
        auto it = byId_.find(imageId);
        if (it != byId_.end()) return *images_[it->second];
        images_.emplace_back(new Image(imageId, this));
        byId_[imageId] = images_.size() - 1;
        return *images_.back();
    }
    const std::vector<std::unique_ptr<Image>> &images() const { return images_; }

private:
    int id_;
    double r0_;
    InteriorOrientation io_;
    std::map<DistortionModel::Type, std::unique_ptr<DistortionModel>> models_;
    std::vector<std::unique_ptr<Image>> images_;
    std::unordered_map<int, size_t> byId_;
};

class DirectlyObservedParameterGroup {   // parameter/DirectlyObservedParameterGroup.java:37-105
public:
    // diagonal stochastic model: every ObservationParameter carries its variance
    explicit DirectlyObservedParameterGroup(std::vector<ObservationParameter *> observedParameters) : obs_(std::move(observedParameters)) { check(); }
    // fully populated dispersion: packed upper, column-major (MTJ UpperSPDPackMatrix data), r (r + 1) / 2 doubles (:49-60)
    DirectlyObservedParameterGroup(std::vector<ObservationParameter *> observedParameters, std::vector<double> dispersionPackedUpper)
        : obs_(std::move(observedParameters)), packed_(std::move(dispersionPackedUpper)) {
        check();
        const size_t r = obs_.size();
        if (packed_.size() != r * (r + 1) / 2)
            throw std::invalid_argument("Error, number of observations and number of rows/columns in dispersion matrix are unequal");
        for (size_t i = 0; i < r; i++) obs_[i]->setVariance(packed_[i + i * (i + 1) / 2]);
    }
    // the reference's argument order (dispersionMatrix first, :49)
    DirectlyObservedParameterGroup(std::vector<double> dispersionPackedUpper, std::vector<ObservationParameter *> observedParameters)
        : DirectlyObservedParameterGroup(std::move(observedParameters), std::move(dispersionPackedUpper)) {}
    bool hasFullyPopulatedWeightMatrix() const { return !packed_.empty(); }
    int getNumberOfParameters() const { return (int)obs_.size(); }
    const std::vector<ObservationParameter *> &observations() const { return obs_; }
    const std::vector<double> &dispersionPacked() const { return packed_; }

private:
    void check() {
        std::unordered_set<ObservationParameter *> s(obs_.begin(), obs_.end());
        if (s.size() != obs_.size()) throw std::invalid_argument("Error, array contains duplicate observations.");
    }
    std::vector<ObservationParameter *> obs_;
    std::vector<double> packed_;
};

class RankDefect {   // defect/RankDefect.java:35-130: which of tx, ty, tz, rx, ry, rz, scale stay FREE
public:
    bool tx = true, ty = true, tz = true, rx = true, ry = true, rz = true, scale = true;
    int getDefect() const { return (int)tx + (int)ty + (int)tz + (int)rx + (int)ry + (int)rz + (int)scale; }
    bool noneFree() const { return getDefect() == 0; }
    void flags(int32_t out[7]) const {
        const bool f[7] = {tx, ty, tz, rx, ry, rz, scale};
        for (int i = 0; i < 7; i++) out[i] = f[i] ? 1 : 0;
    }
};

// Qxx as handed out by BundleAdjustment.getCofactorMatrix(): MTJ UpperSymmPackMatrix layout, element (r, c), r <= c, at r + c (c + 1) / 2
class UpperSymmPackMatrix {
public:
    UpperSymmPackMatrix(int n, std::vector<double> data) : n_(n), data_(std::move(data)) {}
    int numRows() const { return n_; }
    int numColumns() const { return n_; }
    const std::vector<double> &getData() const { return data_; }
    double get(int r, int c) const {
        if (r > c) std::swap(r, c);
        return data_[(size_t)r + (size_t)c * ((size_t)c + 1) / 2];
    }

private:
    int n_;
    std::vector<double> data_;
};

// The flattened problem: exactly the arrays of the set_* calls of include/jaicov_b200.h
struct FlatProblem {
    std::vector<double> io_val, r0, coef_val, eo_val, xy, var, rho, xyz, bar_len, bar_var;
    std::vector<int32_t> io_col, coef_ptr{0}, coef_type, coef_order, coef_col, cam_of_img, eo_col, obj_idx, pt_col, bar_a, bar_b;
    std::vector<int64_t> pt_ptr{0};
    std::vector<uint8_t> is_datum;
    struct Group {
        std::vector<int32_t> kind, index, comp;
        std::vector<double> obs, var, sigma;
    };
    std::vector<Group> groups;
    int32_t free_flags[7] = {0, 0, 0, 0, 0, 0, 0};
    int32_t n_unknowns = 0, n_observations = 0;
};

class BundleAdjustment;

// util/io/writer/AdjustmentResultWritable.java:36: export(BundleAdjustment), called by exportAdjustmentResults() (BA:1164-1171)
class AdjustmentResultWritable {
public:
    virtual ~AdjustmentResultWritable() = default;
    virtual void exportResults(BundleAdjustment &bundleAdjustment) = 0;   // "export" is a C++ keyword
};

class BundleAdjustment {
public:
    using PropertyChangeListener = std::function<void(int state, double oldValue, double newValue)>;
    void setAdjustmentResultWriter(AdjustmentResultWritable *writer) { writer_ = writer; }   // :1123-1125

    // ---- BundleAdjustment.java:652-665 ------------------------------------------------------------------------------------
    void add(Camera *camera) { if (std::find(cameras_.begin(), cameras_.end(), camera) == cameras_.end()) cameras_.push_back(camera); }
    void add(ScaleBar *bar) { if (std::find(scaleBars_.begin(), scaleBars_.end(), bar) == scaleBars_.end()) scaleBars_.push_back(bar); }
    void add(DirectlyObservedParameterGroup *g) { if (std::find(groups_.begin(), groups_.end(), g) == groups_.end()) groups_.push_back(g); }

    // ---- setters :1123-1195 -----------------------------------------------------------------------------------------------
    void setEstimationType(EstimationType t) {   // only L2NORM and SIMULATION are supported (:1132-1137)
        if (t != EstimationType::L2NORM && t != EstimationType::SIMULATION)
            throw std::invalid_argument("BundleAdjustment Error, this estimation type is not supported!");
        estimationType_ = t;
    }
    void setInvertNormalEquation(MatrixInversion m) { invert_ = m; }
    MatrixInversion getInvertNormalEquation() const { return invert_; }
    void useCentroidedCoordinates(bool f) { useCentroid_ = f; }
    void applyAposterioriVarianceOfUnitWeight(bool f) { applyAposteriori_ = f; }
    void setLevenbergMarquardtDampingValue(double lambda) { damping_ = std::fabs(lambda); }
    double getLevenbergMarquardtDampingValue() const { return damping_; }
    void setMaximalNumberOfIterations(int n) { maxIter_ = n; }
    void setDevice(int device) { device_ = device; }      // not in the reference: CUDA device ordinal
    void setSolver(int solver) { solver_ = solver; }      // not in the reference: JAICOV_SOLVER_*
    void addPropertyChangeListener(PropertyChangeListener l) { listeners_.push_back(std::move(l)); }   // :1459-1461
    void interrupt() { interruptFlag_ = 1; }              // :1455-1457, polled by the library twice per pass

    // ---- getters :1048-1118 -----------------------------------------------------------------------------------------------
    int getNumberOfObservations() const { return numObs_; }
    int getNumberOfUnknownParameters() const { return numUnknown_; }
    int getNumberOfDatumConditions() const { return rankDefect_.getDefect(); }
    int getDegreeOfFreedom() const { return numObs_ - numUnknown_ + rankDefect_.getDefect(); }   // :1080-1082
    double getVarianceFactorApriori() const { return sigma2apriori_; }
    double getVarianceFactorAposteriori() const {   // :1090-1093
        const int dof = getDegreeOfFreedom();
        if (dof > 0 && omega_ > 0 && estimationType_ != EstimationType::SIMULATION && applyAposteriori_) return std::fabs(omega_ / (double)dof);
        return sigma2apriori_;
    }
    const RankDefect &getRankDefect() const { return rankDefect_; }
    const std::vector<ObjectCoordinate *> &getObjectCoordinates() const { return objectCoordinates_; }
    const std::vector<Camera *> &getCameras() const { return cameras_; }
    const std::vector<ScaleBar *> &getScaleBars() const { return scaleBars_; }
    int getNumberOfInteriorOrientationParameters() const { return numIO_; }
    int getNumberOfDistortionParameters() const { return numDist_; }
    const jaicov_stats &getStatistics() const { return stats_; }
    // object points in the order of the flattened problem of the last prepareUnknownParameters() (index = obj_idx of the C ABI)
    const std::vector<ObjectCoordinate *> &flatPoints() const { return flatPoints_; }
    const std::string &getLastError() const { return lastError_; }
    // getCofactorMatrix(), :1177-1179; nullptr if the inversion was switched off or no adjustment has run
    const UpperSymmPackMatrix *getCofactorMatrix() const { return Qxx_.get(); }

    // ---- prepareUnknownParameters (:667-782) + detectRankDefect (:836-1042) + flattening ----------------------------------
    // Integer bookkeeping only; writes rows and columns into the object graph exactly as the reference does.
    FlatProblem prepareUnknownParameters() {
        FlatProblem f;
        double sigma2 = sigma2apriori_;
        int counter = 0;
        numObs_ = 0;
        objectCoordinates_.clear();
        std::unordered_map<ObjectCoordinate *, int32_t> pointIndex;   // object point -> index in the flat point list
        std::vector<ObjectCoordinate *> points;                      // flat point list (first appearance anywhere)
        std::unordered_set<ObjectCoordinate *> inAdjustment;
        auto pidx = [&](ObjectCoordinate *oc) {
            auto it = pointIndex.find(oc);
            if (it != pointIndex.end()) return it->second;
            const int32_t i = (int32_t)points.size();
            pointIndex[oc] = i;
            points.push_back(oc);
            return i;
        };
        auto addObjectCoordinate = [&](ObjectCoordinate *oc) {
            if (inAdjustment.insert(oc).second) objectCoordinates_.push_back(oc);
        };
        std::vector<UnknownParameter *> unknownParameters;   // the parameters indexed by THIS call: only they are renumbered (:776-781)
        auto addUnknown = [&](UnknownParameter &p) {   // addUnknownParameter, :645-650
            if (p.getColumn() == COL_UNSET) {
                p.setColumn(counter++);
                unknownParameters.push_back(&p);
            }
        };
        // image observations: rows 2j, 2j + 1 in camera -> image -> point order; object coordinates in order of first appearance (:670-693)
        std::vector<Image *> images;
        for (size_t ci = 0; ci < cameras_.size(); ci++)
            for (auto &img : cameras_[ci]->images()) {
                images.push_back(img.get());
                f.cam_of_img.push_back((int32_t)ci);
                for (ImageCoordinate &ic : img->coordinates()) {
                    ic.rowX = numObs_++;
                    ic.rowY = numObs_++;
                    sigma2 = std::min(sigma2, std::min(ic.getVarianceX(), ic.getVarianceY()));
                    ObjectCoordinate *oc = ic.getObjectCoordinate();
                    f.obj_idx.push_back(pidx(oc));
                    f.xy.push_back(ic.getX()); f.xy.push_back(ic.getY());
                    f.var.push_back(ic.getVarianceX()); f.var.push_back(ic.getVarianceY());
                    f.rho.push_back(ic.getCorrelationCoefficient());
                    if (inAdjustment.insert(oc).second) {
                        objectCoordinates_.push_back(oc);
                        addUnknown(oc->getX()); addUnknown(oc->getY()); addUnknown(oc->getZ());
                    }
                }
                f.pt_ptr.push_back((int64_t)f.obj_idx.size());
            }
        // interior orientation and distortion parameters of every camera, then all exterior orientations (:695-722)
        numIO_ = numDist_ = 0;
        std::vector<std::vector<UnknownParameter *>> camParams;
        for (Camera *cam : cameras_) {
            std::vector<UnknownParameter *> pl;
            for (int i = 0; i < 3; i++) {
                UnknownParameter &p = cam->getInteriorOrientation().at(i);
                if (p.getColumn() == COL_UNSET) numIO_++;
                pl.push_back(&p);
            }
            for (DistortionModel *m : cam->getDistortionModels())
                for (auto &p : m->parameters()) {
                    if (p->getColumn() == COL_UNSET) numDist_++;
                    pl.push_back(p.get());
                }
            camParams.push_back(std::move(pl));
        }
        for (auto &pl : camParams)
            for (UnknownParameter *p : pl) addUnknown(*p);
        for (Image *img : images)
            for (int i = 0; i < 6; i++) addUnknown(img->getExteriorOrientation().at(i));
        // scale bars (:724-745)
        for (ScaleBar *bar : scaleBars_) {
            bar->getLength().setRow(numObs_++);
            ObjectCoordinate *a = bar->getObjectCoordinateA(), *b = bar->getObjectCoordinateB();
            f.bar_a.push_back(pidx(a)); f.bar_b.push_back(pidx(b));
            addObjectCoordinate(a); addObjectCoordinate(b);
            for (ObjectCoordinate *oc : {a, b}) { addUnknown(oc->getX()); addUnknown(oc->getY()); addUnknown(oc->getZ()); }
            f.bar_len.push_back(bar->getLength().getValue()); f.bar_var.push_back(bar->getLength().getVariance());
            sigma2 = std::min(sigma2, bar->getLength().getVariance());
        }
        // directly observed parameters (:747-771)
        std::unordered_map<UnknownParameter *, std::array<int32_t, 3>> slot;   // parameter -> (kind, index, comp) of the C ABI
        {
            int32_t coefBase = 0;
            for (size_t ci = 0; ci < camParams.size(); ci++) {
                for (size_t k = 0; k < camParams[ci].size(); k++)
                    slot[camParams[ci][k]] = k < 3 ? std::array<int32_t, 3>{1, (int32_t)ci, (int32_t)k}
                                                   : std::array<int32_t, 3>{2, coefBase + (int32_t)(k - 3), 0};   // global coefficient list
                coefBase += (int32_t)camParams[ci].size() - 3;
            }
            for (size_t ii = 0; ii < images.size(); ii++)
                for (int k = 0; k < 6; k++) slot[&images[ii]->getExteriorOrientation().at(k)] = {3, (int32_t)ii, k};
        }
        std::vector<std::vector<ParameterType>> groupTypes;
        for (DirectlyObservedParameterGroup *grp : groups_) {
            FlatProblem::Group g;
            std::vector<ParameterType> types;
            for (ObservationParameter *op : grp->observations()) {
                UnknownParameter *ref = op->getReference();
                if (!ref) throw std::invalid_argument("observation without a reference parameter");
                const ParameterType t = ref->getParameterType();
                if (t == ParameterType::OBJECT_COORDINATE_X || t == ParameterType::OBJECT_COORDINATE_Y || t == ParameterType::OBJECT_COORDINATE_Z) {
                    ObjectCoordinate *oc = static_cast<ObjectCoordinate *>(ref->getReference());
                    const int32_t gi = pidx(oc);
                    addObjectCoordinate(oc);
                    addUnknown(*ref);           // only the observed component gets its column here
                    g.kind.push_back(0); g.index.push_back(gi); g.comp.push_back((int)t - (int)ParameterType::OBJECT_COORDINATE_X);
                } else {
                    auto it = slot.find(ref);
                    if (it == slot.end()) throw std::invalid_argument("observed parameter does not belong to a camera or image of this adjustment");
                    addUnknown(*ref);
                    g.kind.push_back(it->second[0]); g.index.push_back(it->second[1]); g.comp.push_back(it->second[2]);
                }
                op->setRow(numObs_++);
                g.obs.push_back(op->getValue());
                if (!op->hasVariance()) throw std::invalid_argument("Error, variance must be positive");
                g.var.push_back(op->getVariance());
                sigma2 = std::min(sigma2, op->getVariance());
                types.push_back(op->getParameterType());
            }
            if (grp->hasFullyPopulatedWeightMatrix()) { g.sigma = grp->dispersionPacked(); g.var.clear(); }
            f.groups.push_back(std::move(g));
            groupTypes.push_back(std::move(types));
        }
        sigma2apriori_ = sigma2 > 0 ? sigma2 : 1.0;   // :221
        numUnknown_ = counter;
        // ---- detectRankDefect, :836-1042 ----------------------------------------------------------------------------------
        const bool hasBars = !scaleBars_.empty();
        RankDefect rd;
        rd.scale = !hasBars;                          // :841-849
        int cnt[3] = {0, 0, 0};
        auto rules = [&]() {                          // :912-937
            if (rd.tx && cnt[0] > 0) rd.tx = false;
            if (rd.ty && cnt[1] > 0) rd.ty = false;
            if (rd.tz && cnt[2] > 0) rd.tz = false;
            if (!hasBars && (cnt[0] >= 2 || cnt[1] >= 2 || cnt[2] >= 2)) rd.scale = false;
            if (rd.rx && cnt[1] >= 2 && cnt[2] >= 2) rd.rx = false;
            if (rd.ry && cnt[0] >= 2 && cnt[2] >= 2) rd.ry = false;
            if (rd.rz && cnt[0] >= 2 && cnt[1] >= 2) rd.rz = false;
            if (cnt[0] > 0 && cnt[1] > 0 && cnt[2] > 0 && cnt[0] + cnt[1] + cnt[2] >= (hasBars ? 6 : 7)) rd.rx = rd.ry = rd.rz = false;
        };
        auto angle = [&](ParameterType t) {
            if (t == ParameterType::CAMERA_OMEGA) rd.rx = false;
            else if (t == ParameterType::CAMERA_PHI) rd.ry = false;
            else if (t == ParameterType::CAMERA_KAPPA) rd.rz = false;
        };
        for (auto &types : groupTypes)                // observed angles fix rotations (:860-881)
            for (ParameterType t : types) {
                angle(t);
                if (!rd.rx && !rd.ry && !rd.rz) break;
            }
        for (auto &types : groupTypes)                // observed coordinates (:883-944)
            for (ParameterType t : types) {
                if (t == ParameterType::CAMERA_COORDINATE_X || t == ParameterType::OBJECT_COORDINATE_X) cnt[0]++;
                else if (t == ParameterType::CAMERA_COORDINATE_Y || t == ParameterType::OBJECT_COORDINATE_Y) cnt[1]++;
                else if (t == ParameterType::CAMERA_COORDINATE_Z || t == ParameterType::OBJECT_COORDINATE_Z) cnt[2]++;
                else angle(t);
                rules();
                if (rd.noneFree()) break;
            }
        bool done = false;
        for (ObjectCoordinate *oc : objectCoordinates_) {   // fixed object components (:946-983)
            for (int c = 0; c < 3; c++) cnt[c] += oc->component(c).getColumn() == COL_FIXED ? 1 : 0;
            rules();
            if (rd.noneFree()) { done = true; break; }
        }
        if (!done && !rd.noneFree())
            for (Image *img : images) {                     // fixed exterior orientations (:990-1041)
                ExteriorOrientation &eo = img->getExteriorOrientation();
                if (rd.rx && eo.at(3).getColumn() == COL_FIXED) rd.rx = false;
                if (rd.ry && eo.at(4).getColumn() == COL_FIXED) rd.ry = false;
                if (rd.rz && eo.at(5).getColumn() == COL_FIXED) rd.rz = false;
                for (int c = 0; c < 3; c++) cnt[c] += eo.at(c).getColumn() == COL_FIXED ? 1 : 0;
                rules();
                if (rd.noneFree()) break;
            }
        rankDefect_ = rd;
        const int d = rd.getDefect();
        // ---- every column += d (:776-781) ---------------------------------------------------------------------------------
        if (d > 0)
            for (UnknownParameter *p : unknownParameters) p->setColumn(p->getColumn() + d);
        // ---- flatten -------------------------------------------------------------------------------------------------------
        for (size_t ci = 0; ci < cameras_.size(); ci++) {
            auto &pl = camParams[ci];
            for (int i = 0; i < 3; i++) { f.io_val.push_back(pl[i]->getValue()); f.io_col.push_back(pl[i]->getColumn()); }
            f.r0.push_back(cameras_[ci]->getR0());
            for (size_t k = 3; k < pl.size(); k++) {
                f.coef_type.push_back((int32_t)pl[k]->getParameterType()); f.coef_order.push_back(pl[k]->getOrder());
                f.coef_val.push_back(pl[k]->getValue()); f.coef_col.push_back(pl[k]->getColumn());
            }
            f.coef_ptr.push_back((int32_t)f.coef_type.size());
        }
        for (Image *img : images)
            for (int i = 0; i < 6; i++) {
                f.eo_val.push_back(img->getExteriorOrientation().at(i).getValue());
                f.eo_col.push_back(img->getExteriorOrientation().at(i).getColumn());
            }
        for (ObjectCoordinate *oc : points) {
            for (int c = 0; c < 3; c++) { f.xyz.push_back(oc->component(c).getValue()); f.pt_col.push_back(oc->component(c).getColumn()); }
            f.is_datum.push_back(oc->isDatum() && inAdjustment.count(oc) ? 1 : 0);
        }
        rd.flags(f.free_flags);
        f.n_unknowns = numUnknown_;
        f.n_observations = numObs_;
        flatPoints_ = points;
        flatCamParams_ = camParams;
        flatImages_ = images;
        flatPointPos_.clear();
        flatImagePos_.clear();
        for (size_t p = 0; p < points.size(); p++) flatPointPos_[points[p]] = (int)p;
        for (size_t i = 0; i < images.size(); i++) flatImagePos_[images[i]] = (int)i;
        return f;
    }

    // ---- estimateModel, :203-387: everything between "parameters are indexed" and "results are exported" runs in the library ----
    EstimationStateType estimateModel() {
        FlatProblem f = prepareUnknownParameters();
        Qxx_.reset();
        omega_ = 0.0;
        jaicov_options o;
        jaicov_default_options(&o);
        o.invert_mode = (int32_t)invert_;
        o.estimation_type = estimationType_ == EstimationType::SIMULATION ? JAICOV_SIMULATION : JAICOV_L2NORM;
        o.max_iterations = maxIter_;
        o.use_centroid = useCentroid_ ? 1 : 0;
        o.apply_aposteriori = applyAposteriori_ ? 1 : 0;
        o.device = device_;
        o.solver = solver_;
        o.sigma2apriori = sigma2apriori_;
        o.damping_value = damping_;
        release();                                     // the previous adjustment's device state, if any
        jaicov_handle *h = nullptr;
        int rc = jaicov_create(&o, &h);
        if (rc != JAICOV_OK) { lastError_ = "jaicov_create failed"; return EstimationStateType::NOT_INITIALISED; }
        auto fail = [&](int code) {
            lastError_ = jaicov_last_error(h);
            jaicov_destroy(h);
            return code == JAICOV_OUT_OF_MEMORY ? EstimationStateType::OUT_OF_MEMORY : EstimationStateType::NOT_INITIALISED;
        };
        const int32_t nCam = (int32_t)cameras_.size(), nImg = (int32_t)f.cam_of_img.size(), nPt = (int32_t)(f.xyz.size() / 3);
        if ((rc = jaicov_set_cameras(h, nCam, f.io_val.data(), f.io_col.data(), f.r0.data(), f.coef_ptr.data(), f.coef_type.data(),
                                     f.coef_order.data(), f.coef_val.data(), f.coef_col.data())) != JAICOV_OK) return fail(rc);
        if ((rc = jaicov_set_images(h, nImg, f.cam_of_img.data(), f.eo_val.data(), f.eo_col.data(), f.pt_ptr.data())) != JAICOV_OK) return fail(rc);
        if ((rc = jaicov_set_image_points(h, (int64_t)f.obj_idx.size(), f.obj_idx.data(), f.xy.data(), f.var.data(), f.rho.data())) != JAICOV_OK)
            return fail(rc);
        if ((rc = jaicov_set_object_points(h, nPt, f.xyz.data(), f.pt_col.data(), f.is_datum.data())) != JAICOV_OK) return fail(rc);
        if (!f.bar_a.empty() &&
            (rc = jaicov_set_scale_bars(h, (int32_t)f.bar_a.size(), f.bar_a.data(), f.bar_b.data(), f.bar_len.data(), f.bar_var.data())) != JAICOV_OK)
            return fail(rc);
        for (auto &g : f.groups)
            if ((rc = jaicov_add_observed_group(h, (int32_t)g.obs.size(), g.kind.data(), g.index.data(), g.comp.data(), g.obs.data(),
                                                g.sigma.empty() ? g.var.data() : nullptr, g.sigma.empty() ? nullptr : g.sigma.data())) != JAICOV_OK)
                return fail(rc);
        if ((rc = jaicov_set_datum(h, f.free_flags, f.n_unknowns, f.n_observations)) != JAICOV_OK) return fail(rc);
        if (invert_ == MatrixInversion::REDUCED || invert_ == MatrixInversion::PRE_ELIMINATION)   // numRows of the reduced system, :262
            jaicov_set_reduced_rows(h, numIO_ + numDist_ + 3 * (int)objectCoordinates_.size() + rankDefect_.getDefect());
        const int id = jaicov_estimate(h, listeners_.empty() ? nullptr : &BundleAdjustment::fire, this, &interruptFlag_);
        interruptFlag_ = 0;
        if (id == JAICOV_ILLEGAL_ARGUMENT || id == JAICOV_NOT_INITIALISED) {
            // conditions for which the reference throws (too few datum points :515-516, ...) or no usable device: the text is kept
            EstimationStateType s = fail(id);
            if (id == JAICOV_ILLEGAL_ARGUMENT) throw std::invalid_argument(lastError_);
            return s;
        }
        jaicov_get_stats(h, &stats_);
        omega_ = stats_.omega;
        // write the adjusted values back into the object graph (what updateUnknownParameters did, :450-462)
        std::vector<double> xyz(f.xyz.size()), io(f.io_val.size()), coef(f.coef_val.size()), eo(f.eo_val.size());
        jaicov_get_values(h, xyz.data(), io.data(), coef.data(), eo.data());
        for (size_t p = 0; p < flatPoints_.size(); p++)
            for (int c = 0; c < 3; c++) flatPoints_[p]->component(c).setValue(xyz[3 * p + c]);
        size_t ki = 0, kc = 0, ke = 0;
        for (auto &pl : flatCamParams_) {
            for (int i = 0; i < 3; i++) pl[i]->setValue(io[ki++]);
            for (size_t k = 3; k < pl.size(); k++) pl[k]->setValue(coef[kc++]);
        }
        for (Image *img : flatImages_)
            for (int i = 0; i < 6; i++) img->getExteriorOrientation().at(i).setValue(eo[ke++]);
        if (id == JAICOV_ERROR_FREE_ESTIMATION && invert_ != MatrixInversion::NONE) {
            const int64_t n = (int64_t)numUnknown_ + rankDefect_.getDefect();
            std::vector<double> q((size_t)(n * (n + 1) / 2), 0.0);   // REDUCED modes fill the leading block only
            if (jaicov_get_qxx_packed(h, q.data()) == JAICOV_OK) Qxx_.reset(new UpperSymmPackMatrix((int)n, std::move(q)));
        }
        handle_ = h;                                   // Qxx stays on the device for the consumers next to the path (transform, writers)
        if ((id == JAICOV_ERROR_FREE_ESTIMATION || id == JAICOV_NO_CONVERGENCE) && writer_) {   // exportAdjustmentResults, :360-368
            try {
                writer_->exportResults(*this);
            } catch (const std::exception &e) {
                lastError_ = e.what();
                return EstimationStateType::EXPORT_ADJUSTMENT_RESULTS_FAILED;
            }
        }
        return (EstimationStateType)id;
    }

    ~BundleAdjustment() { release(); }
    BundleAdjustment() = default;
    BundleAdjustment(const BundleAdjustment &) = delete;
    BundleAdjustment &operator=(const BundleAdjustment &) = delete;
    // frees the device state of the last estimateModel() (the host copy handed out by getCofactorMatrix() stays valid)
    void release() {
        if (handle_) jaicov_destroy(handle_);
        handle_ = nullptr;
    }
    // the library handle of the last estimateModel() and the numbering of its flattened problem (for the callers next to the path)
    jaicov_handle *handle() const { return handle_; }
    int flatPointIndex(const ObjectCoordinate *oc) const {
        auto it = flatPointPos_.find(oc);
        return it == flatPointPos_.end() ? -1 : it->second;
    }
    int flatImageIndex(const Image *img) const {
        auto it = flatImagePos_.find(img);
        return it == flatImagePos_.end() ? -1 : it->second;
    }

private:
    static void fire(void *user, int32_t state, double oldValue, double newValue) {
        for (auto &l : static_cast<BundleAdjustment *>(user)->listeners_) l(state, oldValue, newValue);
    }
    std::vector<Camera *> cameras_;
    std::vector<ScaleBar *> scaleBars_;
    std::vector<DirectlyObservedParameterGroup *> groups_;
    std::vector<ObjectCoordinate *> objectCoordinates_;
    std::vector<PropertyChangeListener> listeners_;
    EstimationType estimationType_ = EstimationType::L2NORM;
    MatrixInversion invert_ = MatrixInversion::FULL;   // :91
    int maxIter_ = 5000;                               // DefaultValue.java:25
    bool applyAposteriori_ = true, useCentroid_ = true;   // :86-87
    double damping_ = 0.0, sigma2apriori_ = 1.0, omega_ = 0.0;   // :96, :98
    int device_ = 0, solver_ = JAICOV_SOLVER_AUTO;
    int numObs_ = 0, numUnknown_ = 0, numIO_ = 0, numDist_ = 0;
    RankDefect rankDefect_;
    volatile int32_t interruptFlag_ = 0;
    jaicov_stats stats_{};
    std::string lastError_;
    std::unique_ptr<UpperSymmPackMatrix> Qxx_;
    std::vector<ObjectCoordinate *> flatPoints_;
    std::vector<std::vector<UnknownParameter *>> flatCamParams_;
    std::vector<Image *> flatImages_;
    AdjustmentResultWritable *writer_ = nullptr;
    std::unordered_map<const ObjectCoordinate *, int> flatPointPos_;
    std::unordered_map<const Image *, int> flatImagePos_;
    jaicov_handle *handle_ = nullptr;
};

// ---- callers either side of the path ----------------------------------------------------------------------

// tranformation/CoordinateTransformationExteriorOrientation.java:49-121: object points carried into the frame of a reference image,
// X_trg = X0_trg + R_trg R_src' (X - X0_src), with the propagated covariance sigma2 * J Qxx J'.  The visibility loops (:57-98) stay
// here; the contraction with Qxx runs on the device that holds it (jaicov_propagate_eo_transform), so transform() takes the
// adjustment whose estimateModel() produced the cofactor matrix instead of a host copy of it.
class CoordinateTransformationExteriorOrientation {
public:
    struct TransformedCoordinate {
        std::string name;           // "<point> <image> <reference image>", :103
        double x, y, z;
    };
    // imagesToAlign: reference image -> images whose object points are carried into its frame (insertion order of the caller)
    void transform(const std::vector<ObjectCoordinate *> &objectCoordinatesToTransform,
                   const std::vector<std::pair<Image *, std::vector<Image *>>> &imagesToAlign, double sigma2, BundleAdjustment &adjustment) {
        if (!adjustment.handle()) throw std::invalid_argument("the adjustment holds no cofactor matrix on the device (run estimateModel() first)");
        point_.clear(); src_.clear(); trg_.clear(); transformed_.clear();
        for (auto &entry : imagesToAlign)                                      // :81-105
            for (Image *image : entry.second)
                for (ObjectCoordinate *oc : objectCoordinatesToTransform) {
                    if (!image->get(oc)) continue;                             // not visible in the current image, :91-95
                    const int p = adjustment.flatPointIndex(oc), si = adjustment.flatImageIndex(image), ti = adjustment.flatImageIndex(entry.first);
                    if (p < 0 || si < 0 || ti < 0) throw std::invalid_argument("point or image is not part of the adjustment");
                    point_.push_back(p); src_.push_back(si); trg_.push_back(ti);
                    transformed_.push_back({oc->getName() + " " + std::to_string(image->getId()) + " " + std::to_string(entry.first->getId()), 0, 0, 0});
                }
        const size_t n = point_.size();
        std::vector<double> xyz(3 * n);
        std::vector<double> cov(3 * n * (3 * n + 1) / 2);
        const int rc = jaicov_propagate_eo_transform(adjustment.handle(), (int32_t)n, point_.data(), src_.data(), trg_.data(), sigma2,
                                                     n ? xyz.data() : nullptr, n ? cov.data() : nullptr);
        if (rc != JAICOV_OK) throw std::runtime_error(std::string("jaicov_propagate_eo_transform: ") + jaicov_last_error(adjustment.handle()));
        for (size_t k = 0; k < n; k++) { transformed_[k].x = xyz[3 * k]; transformed_[k].y = xyz[3 * k + 1]; transformed_[k].z = xyz[3 * k + 2]; }
        covariance_.reset(new UpperSymmPackMatrix((int)(3 * n), std::move(cov)));   // rows 3 k .. 3 k + 2 belong to transformed point k (:146-148)
    }
    const UpperSymmPackMatrix *getCovarianceMatrix() const { return covariance_.get(); }
    const std::vector<TransformedCoordinate> &getTransformedCoordinates() const { return transformed_; }
    const std::vector<int32_t> &points() const { return point_; }
    const std::vector<int32_t> &sourceImages() const { return src_; }
    const std::vector<int32_t> &targetImages() const { return trg_; }

private:
    std::vector<int32_t> point_, src_, trg_;
    std::vector<TransformedCoordinate> transformed_;
    std::unique_ptr<UpperSymmPackMatrix> covariance_;
};

// String.format(Locale.ENGLISH, "%[+]<width>.<precision>f", double) as java.util.Formatter prints it (the reference needs JDK 25): the
// digits of Double.toString -- the SHORTEST decimal that round-trips -- rounded HALF_UP to the precision and padded with zeros beyond
// them.  printf expands the exact binary value and rounds half-even instead ("%.20f" of 0.1: 0.10000000000000000000 in Java,
// 0.10000000000000000555 in C; "%.1f" of 0.15: 0.2 vs 0.1); the writer's "%35.15f" carries up to 18 significant digits, so the last
// digits of the .info / .cxx files depend on it.  Same function as writers.py: java_format_f (compared on random doubles in the tests).
inline std::string java_format_f(double v, int width, int precision, bool plus) {
    std::string s;
    if (std::isnan(v)) s = "NaN";
    else if (std::isinf(v)) s = std::string(v < 0 ? "-" : (plus ? "+" : "")) + "Infinity";
    else {
        char buf[64];
        const auto r = std::to_chars(buf, buf + sizeof buf - 1, std::fabs(v), std::chars_format::scientific);   // shortest round-trip digits
        *r.ptr = 0;
        std::string digs;
        const char *p = buf;
        for (; *p && *p != 'e'; ++p)
            if (*p != '.') digs.push_back(*p);
        const long point = (*p ? std::atol(p + 1) : 0) + 1;       // |v| = 0.d1d2d3... x 10^point
        const long keep = point + precision;                      // digits in front of the rounding position
        std::string kept;                                         // round(|v| 10^precision) as a digit string
        if (keep >= 0) {
            kept = digs.substr(0, (size_t)std::min<long>(keep, (long)digs.size()));
            const bool up = keep < (long)digs.size() && digs[(size_t)keep] >= '5';
            if (keep > (long)digs.size()) kept.append((size_t)(keep - (long)digs.size()), '0');
            if (up) {
                long i = (long)kept.size() - 1;
                for (; i >= 0; i--) {
                    if (kept[(size_t)i] == '9') kept[(size_t)i] = '0';
                    else { kept[(size_t)i]++; break; }
                }
                if (i < 0) kept.insert(kept.begin(), '1');
            }
        }
        if ((long)kept.size() < precision + 1) kept.insert(0, (size_t)(precision + 1 - (long)kept.size()), '0');
        if (std::signbit(v)) s = "-";
        else if (plus) s = "+";
        s += kept.substr(0, kept.size() - (size_t)precision);
        if (precision > 0) { s += '.'; s += kept.substr(kept.size() - (size_t)precision); }
    }
    if ((long)s.size() < width) s.insert(0, (size_t)(width - (long)s.size()), ' ');
    return s;
}

// util/io/writer/DefaultResultWriter.java:46-155: <base>.info (name, component, coordinate, row / column of the exported matrix;
// format :67) and <base>.cxx (sigma0^2 a posteriori * Qxx of the object coordinates; format :142).  Only the exported sub-matrix
// leaves the device (jaicov_get_qxx_submatrix); the full cofactor matrix never travels to the host for this.
class DefaultResultWriter : public AdjustmentResultWritable {
public:
    explicit DefaultResultWriter(std::string exportPathAndFileBaseName) : base_(std::move(exportPathAndFileBaseName)) {}
    const std::string &getExportPathAndFileBaseName() const { return base_; }
    void exportResults(BundleAdjustment &adj) override {
        if (base_.empty()) throw std::invalid_argument("Error, export path cannot be null!");
        std::vector<int32_t> indices;
        FILE *f = std::fopen((base_ + ".info").c_str(), "w");
        if (!f) throw std::runtime_error("cannot write " + base_ + ".info");
        int k = 0;
        for (ObjectCoordinate *oc : adj.getObjectCoordinates())
            for (int c = 0; c < 3; c++) {
                UnknownParameter &p = oc->component(c);
                int ci = -1;
                if (p.getColumn() >= 0 && p.getColumn() < COL_FIXED) { indices.push_back(p.getColumn()); ci = k++; }
                std::fprintf(f, "%25s\t%5s\t%s\t%10d\n", oc->getName().c_str(), c == 0 ? "X" : (c == 1 ? "Y" : "Z"), java_format_f(p.getValue(), 35, 15, false).c_str(), ci);
            }
        std::fclose(f);
        if (adj.getInvertNormalEquation() == MatrixInversion::NONE || !adj.handle()) return;   // no cofactor matrix: .info only
        const size_t n = indices.size();
        std::vector<double> C(n * n);
        if (n && jaicov_get_qxx_submatrix(adj.handle(), (int32_t)n, indices.data(), adj.getVarianceFactorAposteriori(), C.data()) != JAICOV_OK)
            throw std::runtime_error(std::string("jaicov_get_qxx_submatrix: ") + jaicov_last_error(adj.handle()));
        f = std::fopen((base_ + ".cxx").c_str(), "w");
        if (!f) throw std::runtime_error("cannot write " + base_ + ".cxx");
        for (size_t r = 0; r < n; r++) {
            for (size_t c = 0; c < n; c++) { std::fputs(java_format_f(C[r * n + c], 35, 15, true).c_str(), f); std::fputs("  ", f); }
            std::fprintf(f, "\n");
        }
        std::fclose(f);
    }

private:
    std::string base_;
};

// dlt/DLTCoefficients.java:33-84 and dlt/DirectLinearTransformation.java:49-184: initial interior / exterior orientation of images
// from homologous points.  adjust() keeps the reference's per-image meaning; adjustAll() is the batched form (one kernel launch for
// all images, jaicov_dlt_batch).
class DLTCoefficients {
public:
    explicit DLTCoefficients(Image *image) : image_(image) {
        static const int order[20] = {611, 612, 613, 614, 621, 622, 623, 624, 631, 632, 633, 111, 112, 113, 251, 252, 253, 261, 262, 263};
        for (int t : order) { type_.push_back(t); value_.push_back(0.0); fixed_.push_back(false); }
    }
    Image *getReference() const { return image_; }
    int size() const { return 20; }
    int typeId(int i) const { return type_[i]; }             // ParameterType ids in the reference's insertion order
    double getValue(int i) const { return value_[i]; }
    void setValue(int i, double v) { value_[i] = v; }
    bool isFixed(int i) const { return fixed_[i]; }
    void setFixed(int i, bool f) { fixed_[i] = f; }
    int indexOf(int typeId) const {
        for (int i = 0; i < 20; i++)
            if (type_[i] == typeId) return i;
        return -1;
    }

private:
    Image *image_;
    std::vector<int> type_;
    std::vector<double> value_;
    std::vector<bool> fixed_;
};

class DirectLinearTransformation {
public:
    enum class RestrictionType : int {   // ordinal order, DirectLinearTransformation.java:50-57
        IDENTICAL_PRINCIPLE_DISTANCE = 0, ROTATION_WITHOUT_SHEAR = 1, FIXED_PRINCIPLE_DISTANCE_X = 2, FIXED_PRINCIPLE_DISTANCE_Y = 3,
        FIXED_PRINCIPAL_POINT_X = 4, FIXED_PRINCIPAL_POINT_Y = 5
    };
    static constexpr int maximalNumberOfIterations = 5000;   // DefaultValue.getMaximalNumberOfIterations(), :63

    // the homologous points of every image, gathered the way :78-94 does (image point + object coordinates matched by name)
    struct Gathered {
        std::vector<int64_t> pt_ptr{0};
        std::vector<double> xy, xyz, io;
    };
    static Gathered gather(std::vector<DLTCoefficients *> &coefficients, const std::map<std::string, ObjectCoordinate *> &objectCoordinates) {
        Gathered g;
        for (DLTCoefficients *coef : coefficients) {
            Image *image = coef->getReference();
            InteriorOrientation &inner = image->getReference()->getInteriorOrientation();
            // prepareUnknwonParameters, :279-314: values reset, the camera's interior orientation taken over (fixed stays fixed)
            for (int i = 0; i < 20; i++) coef->setValue(i, 0.0);
            UnknownParameter *ioP[3] = {&inner.getPrincipleDistance(), &inner.getPrinciplePointX(), &inner.getPrinciplePointY()};
            for (UnknownParameter *p : ioP) {
                const int k = coef->indexOf((int)p->getParameterType());
                coef->setValue(k, p->getValue());
                if (p->getColumn() == COL_FIXED) coef->setFixed(k, true);
            }
            for (ImageCoordinate &ic : image->coordinates()) {
                auto it = objectCoordinates.find(ic.getObjectCoordinate()->getName());
                if (it == objectCoordinates.end()) continue;
                g.xy.push_back(ic.getX()); g.xy.push_back(ic.getY());
                g.xyz.push_back(it->second->getX().getValue()); g.xyz.push_back(it->second->getY().getValue()); g.xyz.push_back(it->second->getZ().getValue());
            }
            g.pt_ptr.push_back((int64_t)g.xy.size() / 2);
            g.io.push_back(coef->getValue(coef->indexOf(113))); g.io.push_back(coef->getValue(coef->indexOf(111))); g.io.push_back(coef->getValue(coef->indexOf(112)));
        }
        return g;
    }
    // one bool per image: what adjust() returns in the reference
    static std::vector<bool> adjustAll(std::vector<DLTCoefficients *> &coefficients, const std::map<std::string, ObjectCoordinate *> &objectCoordinates,
                                       const std::vector<RestrictionType> &restrictions = {}, int device = 0) {
        Gathered g = gather(coefficients, objectCoordinates);
        const int32_t nImg = (int32_t)coefficients.size();
        std::vector<int32_t> restr;
        for (RestrictionType r : restrictions) restr.push_back((int32_t)r);
        std::vector<double> out((size_t)20 * std::max(nImg, 1));
        std::vector<int32_t> status(std::max(nImg, 1)), passes(std::max(nImg, 1));
        const int rc = jaicov_dlt_batch(device, nImg, g.pt_ptr.data(), g.xy.data(), g.xyz.data(), g.io.data(), (int32_t)restr.size(), restr.data(),
                                        maximalNumberOfIterations, out.data(), status.data(), passes.data());
        if (rc != JAICOV_OK) throw std::runtime_error("jaicov_dlt_batch failed (no sm_100 device, or illegal argument): " + std::to_string(rc));
        std::vector<bool> ok;
        for (int32_t i = 0; i < nImg; i++) {
            DLTCoefficients &c = *coefficients[i];
            if (status[i] < 0) { ok.push_back(false); continue; }
            const double *o = out.data() + (size_t)20 * i;
            for (int k = 0; k < 11; k++) c.setValue(k, o[k]);
            // :248-265: fixed interior orientation values are kept; out = ..., c, x0, y0, X0, Y0, Z0, omega, phi, kappa
            const int ioIdx[3] = {c.indexOf(113), c.indexOf(111), c.indexOf(112)};
            for (int k = 0; k < 3; k++)
                if (!c.isFixed(ioIdx[k])) c.setValue(ioIdx[k], o[11 + k]);
            for (int k = 14; k < 20; k++) c.setValue(k, o[k]);
            ok.push_back(status[i] == 1);
        }
        return ok;
    }
    static bool adjust(DLTCoefficients &coefficients, const std::map<std::string, ObjectCoordinate *> &objectCoordinates,
                       const std::vector<RestrictionType> &restrictions = {}, int device = 0) {
        std::vector<DLTCoefficients *> one{&coefficients};
        return adjustAll(one, objectCoordinates, restrictions, device)[0];
    }
};

}  // namespace host
}  // namespace jaicov
