// example_adjustment.cpp -- the call sequence of the reference's example programs (bundle/example/ExampleFlatFiles.java:
// build cameras / images / object points, configure, estimateModel(), read the results back through the getters) written
// against the native host mirror bundle-adjustment_b200/host/jaicov_host.hpp.  The network is a small synthetic one generated
// here (two rings of camera stations around a box of targets, a pin-hole camera with one radial coefficient), so the program
// needs no input files; everything numerical happens in libjaicov_b200.so on the GPU.
//
//   g++ -std=c++17 -O2 examples/example_adjustment.cpp -Lbundle-adjustment_b200 -ljaicov_b200 -Wl,-rpath,$PWD/bundle-adjustment_b200 -o /tmp/example_adjustment
#include <cmath>
#include <cstdio>
#include <memory>
#include <random>
#include <vector>

#include "../bundle-adjustment_b200/host/jaicov_host.hpp"

using namespace jaicov::host;

int main() {
    std::mt19937_64 rng(20260000);
    std::normal_distribution<double> gauss(0.0, 1.0);
    std::uniform_real_distribution<double> uni(-1.0, 1.0);
    const int nPoints = 80, nImages = 12;
    const double c = 28.8, a1 = -1.1e-4, r0 = 13.488, sigma = 0.0005;

    std::vector<std::unique_ptr<ObjectCoordinate>> points;
    std::vector<std::array<double, 3>> truth;
    for (int i = 0; i < nPoints; i++) {
        truth.push_back({1000 * uni(rng), 750 * uni(rng), 250 * uni(rng)});
        points.emplace_back(new ObjectCoordinate(std::to_string(i + 1), truth[i][0] + 0.5 * gauss(rng), truth[i][1] + 0.5 * gauss(rng),
                                                 truth[i][2] + 0.5 * gauss(rng)));
    }
    Camera camera(1, r0, {DistortionModel::Type::RADIAL_DISTORTION});
    camera.getInteriorOrientation().getPrincipleDistance().setValue(c * 1.01);
    camera.getInteriorOrientation().getPrinciplePointX().setColumn(COL_FIXED);   // no roll diversity in this toy network: keep the
    camera.getInteriorOrientation().getPrinciplePointY().setColumn(COL_FIXED);   // principal point fixed (UnknownParameter.java:27)
    camera.getDistortionModel(DistortionModel::Type::RADIAL_DISTORTION)->add(1).setValue(0.0);

    for (int i = 0; i < nImages; i++) {
        // station on a ring, camera z axis pointing away from the scene centre (the reference's convention: N < 0 in front)
        const double az = 2 * M_PI * i / nImages, el = (i % 2 ? 0.35 : 0.8), rad = 3000;
        const double X0[3] = {rad * std::cos(el) * std::cos(az), rad * std::cos(el) * std::sin(az), rad * std::sin(el)};
        double z[3] = {X0[0], X0[1], X0[2]}, nz = std::sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
        for (double &v : z) v /= nz;
        double x[3] = {-z[1], z[0], 0.0}, nx = std::sqrt(x[0] * x[0] + x[1] * x[1]);
        for (double &v : x) v /= nx;
        const double y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
        // R = [x y z] (columns) -> omega, phi, kappa of PartialDerivativeFactory.java:125-135
        const double phi = std::asin(z[0]), omega = std::atan2(-z[1], z[2]), kappa = std::atan2(-y[0], x[0]);
        Image &image = camera.add(i + 1);
        ExteriorOrientation &eo = image.getExteriorOrientation();
        const double eoTrue[6] = {X0[0], X0[1], X0[2], omega, phi, kappa};
        for (int k = 0; k < 6; k++) eo.at(k).setValue(eoTrue[k] + (k < 3 ? 0.5 : 1e-4) * gauss(rng));
        const double so = std::sin(omega), co = std::cos(omega), sp = std::sin(phi), cp = std::cos(phi), sk = std::sin(kappa), ck = std::cos(kappa);
        const double R[3][3] = {{cp * ck, -cp * sk, sp}, {co * sk + so * sp * ck, co * ck - so * sp * sk, -so * cp}, {so * sk - co * sp * ck, so * ck + co * sp * sk, co * cp}};
        for (int p = 0; p < nPoints; p++) {
            const double d[3] = {truth[p][0] - X0[0], truth[p][1] - X0[1], truth[p][2] - X0[2]};
            const double kx = R[0][0] * d[0] + R[1][0] * d[1] + R[2][0] * d[2], ky = R[0][1] * d[0] + R[1][1] * d[1] + R[2][1] * d[2];
            const double N = R[0][2] * d[0] + R[1][2] * d[1] + R[2][2] * d[2];
            const double xs = -c * kx / N, ys = -c * ky / N, rr = xs * xs + ys * ys, dr = a1 * (rr - r0 * r0);
            image.add(points[p].get(), xs + xs * dr + sigma * gauss(rng), ys + ys * dr + sigma * gauss(rng), sigma, sigma);
        }
    }

    BundleAdjustment adjustment;
    adjustment.add(&camera);
    adjustment.setInvertNormalEquation(MatrixInversion::FULL);
    adjustment.addPropertyChangeListener([](int state, double, double value) {
        if (state == JAICOV_STATE_ITERATE) std::printf("  pass, max|dx| = %.3e\n", value);
    });
    const EstimationStateType state = adjustment.estimateModel();
    std::printf("observations %d, unknowns %d, datum defect %d, redundancy %d\n", adjustment.getNumberOfObservations(),
                adjustment.getNumberOfUnknownParameters(), adjustment.getNumberOfDatumConditions(), adjustment.getDegreeOfFreedom());
    if (state != EstimationStateType::ERROR_FREE_ESTIMATION) {
        std::printf("estimateModel() -> %d (%s)\n", (int)state, adjustment.getLastError().c_str());
        return state == EstimationStateType::NOT_INITIALISED ? 0 : 1;   // no sm_100 device: nothing to show, not a failure of the example
    }
    const double s2 = adjustment.getVarianceFactorAposteriori();
    std::printf("sigma0 a posteriori / a priori = %.4f, c = %.6f, A1 = %.6e\n", std::sqrt(s2 / adjustment.getVarianceFactorApriori()),
                camera.getInteriorOrientation().getPrincipleDistance().getValue(),
                camera.getDistortionModel(DistortionModel::Type::RADIAL_DISTORTION)->get(1)->getValue());
    const UpperSymmPackMatrix *Q = adjustment.getCofactorMatrix();
    const int col = camera.getInteriorOrientation().getPrincipleDistance().getColumn();
    std::printf("sigma(c) = %.3e (from the complete cofactor matrix, %d x %d)\n", std::sqrt(s2 * Q->get(col, col)), Q->numRows(), Q->numColumns());
    return 0;
}
