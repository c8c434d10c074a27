// camera.hpp — PinholeCamera as far as the hot path's callers use it (reference source/vision/camera.hpp:13-55,
// camera.cpp:14-18,55-142): intrinsics + cached inverse, normalize_point(s), the text file format
// "fx fy shear px py" / six se3 numbers, and CameraManager, the process-wide camera ImagePair reads its K from
// (reference source/front-end/camera-manager.hpp:6-14, camera-manager.cpp:10-38).  Host code only.
#pragma once
#include <fstream>
#include <mutex>
#include <string>
#include <vector>

#include "types.hpp"

namespace mvSLAM {

using CameraExtrinsics = SE3;

namespace b200 {
/** Eigen's fixed 3x3 inverse (cofactor^T / det) in the operation order the device uses (csrc/api.cu h_inverse3). */
inline Matrix3Type inverse3(const Matrix3Type &K)
{
    const double c00 = K(1, 1) * K(2, 2) - K(1, 2) * K(2, 1), c01 = K(1, 2) * K(2, 0) - K(1, 0) * K(2, 2), c02 = K(1, 0) * K(2, 1) - K(1, 1) * K(2, 0);
    const double c10 = K(0, 2) * K(2, 1) - K(0, 1) * K(2, 2), c11 = K(0, 0) * K(2, 2) - K(0, 2) * K(2, 0), c12 = K(0, 1) * K(2, 0) - K(0, 0) * K(2, 1);
    const double c20 = K(0, 1) * K(1, 2) - K(0, 2) * K(1, 1), c21 = K(0, 2) * K(1, 0) - K(0, 0) * K(1, 2), c22 = K(0, 0) * K(1, 1) - K(0, 1) * K(1, 0);
    const double id = 1.0 / (K(0, 0) * c00 + K(0, 1) * c01 + K(0, 2) * c02);
    Matrix3Type I;
    I(0, 0) = c00 * id; I(0, 1) = c10 * id; I(0, 2) = c20 * id;
    I(1, 0) = c01 * id; I(1, 1) = c11 * id; I(1, 2) = c21 * id;
    I(2, 0) = c02 * id; I(2, 1) = c12 * id; I(2, 2) = c22 * id;
    return I;
}
/** SE3::exp (source/math/lie-group.hpp:268-297), translation part first. */
inline SE3 se3_exp(const double v[6])
{
    const double w[3] = {v[3], v[4], v[5]};
    const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double A, B, C;
    if (theta < 1e-5) { A = 1.0 - theta * theta / 6.0; B = 0.5 - theta * theta / 24.0; C = 1.0 / 6.0 - theta * theta / 120.0; }
    else { A = std::sin(theta) / theta; B = (1.0 - std::cos(theta)) / (theta * theta); C = (1.0 - A) / (theta * theta); }
    const double Kx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double KK[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) KK[i * 3 + j] = Kx[i * 3] * Kx[j] + Kx[i * 3 + 1] * Kx[3 + j] + Kx[i * 3 + 2] * Kx[6 + j];
    Matrix3Type R;
    double t[3];
    for (int i = 0; i < 3; ++i) {
        t[i] = 0;
        for (int j = 0; j < 3; ++j) {
            R(i, j) = (i == j ? 1.0 : 0.0) + A * Kx[i * 3 + j] + B * KK[i * 3 + j];
            t[i] += ((i == j ? 1.0 : 0.0) + B * Kx[i * 3 + j] + C * KK[i * 3 + j]) * v[j];
        }
    }
    return SE3(SO3(R), Vector3Type(t[0], t[1], t[2]));
}
}  // namespace b200

class PinholeCamera {
public:
    explicit PinholeCamera(const std::string &filename)
    {
        if (!load_from_file(filename)) throw b200::Error(MVS_E_BAD_ARG, "PinholeCamera: cannot read " + filename);   // camera.cpp:10-11 assert
    }
    PinholeCamera(const CameraIntrinsics &K_, const CameraExtrinsics &P_) : K(K_), K_inv(b200::inverse3(K_)), P(P_) {}

    /** camera.cpp:55-79: K^-1 (u, v, 1) */
    IdealCameraImagePoint normalize_point(const ImagePoint &p) const
    {
        IdealCameraImagePoint r;
        for (int i = 0; i < 3; ++i) r[i] = (K_inv(i, 0) * p.x + K_inv(i, 1) * p.y) + K_inv(i, 2) * 1.0;
        return r;
    }
    std::vector<IdealCameraImagePoint> normalize_points(const std::vector<ImagePoint> &pts) const
    {
        std::vector<IdealCameraImagePoint> r;
        r.reserve(pts.size());
        for (const auto &p : pts) r.push_back(normalize_point(p));
        return r;
    }
    const CameraIntrinsics &get_intrinsics() const { return K; }
    const Matrix3Type &get_intrinsics_inverse() const { return K_inv; }
    const CameraExtrinsics &get_extrinsics() const { return P; }

    /** camera.cpp:105-124: "fx fy shear px py" then six se3 numbers (translation first) */
    bool load_from_file(const std::string &filename)
    {
        std::ifstream in(filename);
        double fx, fy, sh, px, py, se3[6];
        if (!(in >> fx >> fy >> sh >> px >> py)) return false;
        for (double &x : se3) if (!(in >> x)) return false;
        K = Matrix3Type::Zero();
        K(0, 0) = fx; K(1, 1) = fy; K(0, 1) = sh; K(0, 2) = px; K(1, 2) = py; K(2, 2) = 1;
        K_inv = b200::inverse3(K);
        P = b200::se3_exp(se3);
        for (int i = 0; i < 6; ++i) m_se3[i] = se3[i];
        return true;
    }
    /** camera.cpp:126-142 (the tangent vector read from the file is written back unchanged) */
    bool save_to_file(const std::string &filename) const
    {
        std::ofstream out(filename);
        out << K(0, 0) << " " << K(1, 1) << " " << K(0, 1) << " " << K(0, 2) << " " << K(1, 2) << std::endl;
        out << m_se3[0] << " " << m_se3[1] << " " << m_se3[2] << " " << m_se3[3] << " " << m_se3[4] << " " << m_se3[5] << std::endl;
        return bool(out);
    }

private:
    CameraIntrinsics K;
    Matrix3Type K_inv;
    CameraExtrinsics P;
    double m_se3[6] = {0, 0, 0, 0, 0, 0};
};

/** camera-manager.hpp:6-14: THE camera; an ideal camera (K = I) until load_from_file. */
class CameraManager {
public:
    static const PinholeCamera &get_camera()
    {
        std::lock_guard<std::mutex> lock(mutex());
        return camera();
    }
    static void load_from_file(const std::string &filename)
    {
        std::lock_guard<std::mutex> lock(mutex());
        if (!camera().load_from_file(filename)) throw b200::Error(MVS_E_BAD_ARG, "CameraManager: cannot read " + filename);   // :29 assert
    }
    static void save_to_file(const std::string &filename)
    {
        std::lock_guard<std::mutex> lock(mutex());
        if (!camera().save_to_file(filename)) throw b200::Error(MVS_E_BAD_ARG, "CameraManager: cannot write " + filename);
    }
    /** not in the reference: sets THE camera from values already in memory (tests, tools) */
    static void set_camera(const PinholeCamera &c)
    {
        std::lock_guard<std::mutex> lock(mutex());
        camera() = c;
    }
private:
    static PinholeCamera &camera() { static PinholeCamera c(Matrix3Type::Identity(), SE3()); return c; }
    static std::mutex &mutex() { static std::mutex m; return m; }
};

}  // namespace mvSLAM
