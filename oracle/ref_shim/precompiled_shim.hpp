// precompiled_shim.hpp — force-included (-include) when compiling the reference's own leaf
// sources from /root/reference for oracle/_ref (TEST INFRASTRUCTURE).  Stands in for the
// reference's PCH util/precompiled.hpp, which pulls Boost/Eigen/GL/OpenCV/Qt that this image
// lacks.  Only what util/{lineiter,ray,plane,vectorimage} and stereo/{adaptive,geodesic}weight
// touch is provided: std headers and a 3-vector with Eigen::Vector3d's interface.
#ifndef SR_REF_PRECOMPILED_SHIM_HPP
#define SR_REF_PRECOMPILED_SHIM_HPP
#if defined(__cplusplus)
#define _USE_MATH_DEFINES
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <functional>
#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <vector>

using std::fabs;
using std::sqrt;

namespace Eigen {
struct Vector3d {
    double v[3];
    Vector3d() { v[0] = v[1] = v[2] = 0.0; }
    Vector3d(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; }
    static Vector3d Zero() { return Vector3d(0, 0, 0); }
    double &operator[](int i) { return v[i]; }
    const double &operator[](int i) const { return v[i]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
    double dot(const Vector3d &o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    double squaredNorm() const { return dot(*this); }
    double norm() const { return std::sqrt(squaredNorm()); }
    Vector3d normalized() const { double n = norm(); return Vector3d(v[0] / n, v[1] / n, v[2] / n); }
    void normalize() { double n = norm(); v[0] /= n; v[1] /= n; v[2] /= n; }
    Vector3d &operator+=(const Vector3d &o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
    Vector3d &operator-=(const Vector3d &o) { v[0] -= o.v[0]; v[1] -= o.v[1]; v[2] -= o.v[2]; return *this; }
    Vector3d &operator*=(double s) { v[0] *= s; v[1] *= s; v[2] *= s; return *this; }
    Vector3d operator-() const { return Vector3d(-v[0], -v[1], -v[2]); }
};
inline Vector3d operator+(const Vector3d &a, const Vector3d &b) { return Vector3d(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Vector3d operator-(const Vector3d &a, const Vector3d &b) { return Vector3d(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline Vector3d operator*(double s, const Vector3d &a) { return Vector3d(s * a[0], s * a[1], s * a[2]); }
inline Vector3d operator*(const Vector3d &a, double s) { return Vector3d(s * a[0], s * a[1], s * a[2]); }
inline Vector3d operator/(const Vector3d &a, double s) { return Vector3d(a[0] / s, a[1] / s, a[2] / s); }
}  // namespace Eigen

#define FORWARD_DECLARE(cls) \
    class cls;               \
    typedef std::shared_ptr<cls> cls##Ptr; \
    typedef std::weak_ptr<cls> cls##WeakPtr
#endif
#endif
