// util/precompiled.hpp — Qt-free stand-in for the reference's PCH (util/precompiled.hpp:21-66):
// std headers, the FORWARD_DECLARE macro (:60-63) and a small 3-vector / 3x3 / 3x4 algebra with
// the subset of Eigen's interface the stereo/ and project/ class API uses.  If the real Eigen
// has been included first (EIGEN_CORE_H) the stand-ins are skipped.
#ifndef SR_UTIL_PRECOMPILED_HPP
#define SR_UTIL_PRECOMPILED_HPP
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#define FORWARD_DECLARE(cls)               \
    class cls;                             \
    typedef std::shared_ptr<cls> cls##Ptr; \
    typedef std::weak_ptr<cls> cls##WeakPtr

#ifndef EIGEN_CORE_H
namespace Eigen {
struct Vector3d {
    double v[3];
    Vector3d() : v{0, 0, 0} {}
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    static Vector3d Zero() { return Vector3d(); }
    double &operator[](int i) { return v[i]; }
    const double &operator[](int i) const { return v[i]; }
    double &operator()(int i) { return v[i]; }
    const double &operator()(int i) const { return v[i]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
    double dot(const Vector3d &o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    double squaredNorm() const { return dot(*this); }
    double norm() const { return std::sqrt(squaredNorm()); }
    Vector3d normalized() const { double n = norm(); return Vector3d(v[0] / n, v[1] / n, v[2] / n); }
    void normalize() { *this = normalized(); }
    Vector3d &operator+=(const Vector3d &o) { for (int i = 0; i < 3; ++i) v[i] += o.v[i]; return *this; }
    Vector3d &operator-=(const Vector3d &o) { for (int i = 0; i < 3; ++i) v[i] -= o.v[i]; return *this; }
    Vector3d &operator*=(double s) { for (int i = 0; i < 3; ++i) v[i] *= s; return *this; }
    Vector3d operator-() const { return Vector3d(-v[0], -v[1], -v[2]); }
};
inline Vector3d operator+(const Vector3d &a, const Vector3d &b) { return Vector3d(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Vector3d operator-(const Vector3d &a, const Vector3d &b) { return Vector3d(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline Vector3d operator*(double s, const Vector3d &a) { return Vector3d(s * a[0], s * a[1], s * a[2]); }
inline Vector3d operator*(const Vector3d &a, double s) { return s * a; }
inline Vector3d operator/(const Vector3d &a, double s) { return Vector3d(a[0] / s, a[1] / s, a[2] / s); }

struct Matrix3d {
    double m[9];  // row-major
    Matrix3d() : m{0, 0, 0, 0, 0, 0, 0, 0, 0} {}
    static Matrix3d Identity() { Matrix3d r; r.m[0] = r.m[4] = r.m[8] = 1; return r; }
    static Matrix3d Zero() { return Matrix3d(); }
    double &operator()(int i, int j) { return m[3 * i + j]; }
    const double &operator()(int i, int j) const { return m[3 * i + j]; }
    Vector3d col(int j) const { return Vector3d(m[j], m[3 + j], m[6 + j]); }
    Vector3d row(int i) const { return Vector3d(m[3 * i], m[3 * i + 1], m[3 * i + 2]); }
    void setCol(int j, const Vector3d &c) { m[j] = c[0]; m[3 + j] = c[1]; m[6 + j] = c[2]; }
    void setRow(int i, const Vector3d &r) { m[3 * i] = r[0]; m[3 * i + 1] = r[1]; m[3 * i + 2] = r[2]; }
    Matrix3d transpose() const { Matrix3d r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = (*this)(j, i); return r; }
    Matrix3d inverse() const {  // cofactor formula, as Eigen's fixed-size 3x3 inverse
        const Matrix3d &a = *this;
        Matrix3d c;
        c(0, 0) = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1);
        c(0, 1) = -(a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0));
        c(0, 2) = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
        c(1, 0) = -(a(0, 1) * a(2, 2) - a(0, 2) * a(2, 1));
        c(1, 1) = a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0);
        c(1, 2) = -(a(0, 0) * a(2, 1) - a(0, 1) * a(2, 0));
        c(2, 0) = a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1);
        c(2, 1) = -(a(0, 0) * a(1, 2) - a(0, 2) * a(1, 0));
        c(2, 2) = a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0);
        const double det = a(0, 0) * c(0, 0) + a(0, 1) * c(0, 1) + a(0, 2) * c(0, 2);
        Matrix3d r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = c(j, i) / det;
        return r;
    }
};
inline Vector3d operator*(const Matrix3d &M, const Vector3d &v) {
    return Vector3d(M.m[0] * v[0] + M.m[1] * v[1] + M.m[2] * v[2], M.m[3] * v[0] + M.m[4] * v[1] + M.m[5] * v[2],
                    M.m[6] * v[0] + M.m[7] * v[1] + M.m[8] * v[2]);
}
inline Matrix3d operator*(const Matrix3d &A, const Matrix3d &B) {
    Matrix3d r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j);
    return r;
}
}  // namespace Eigen
#endif  // EIGEN_CORE_H

// 3x4 projection matrix, row-major (typedef Eigen::Matrix<double,3,4> ProjMat in project/camera.hpp:44)
struct ProjMat {
    double m[12];
    ProjMat() { for (double &x : m) x = 0; }
    static ProjMat Zero() { return ProjMat(); }
    double &operator()(int i, int j) { return m[4 * i + j]; }
    const double &operator()(int i, int j) const { return m[4 * i + j]; }
};
#endif
