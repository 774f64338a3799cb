// precompiled_shim.hpp — force-included (-include) when compiling the reference's own leaf
// sources from /root/reference for oracle/_ref (TEST INFRASTRUCTURE).  Stands in for the
// reference's PCH util/precompiled.hpp, which pulls Boost/Eigen/GL/OpenCV/Qt that this image
// lacks.  Only what util/{lineiter,ray,plane,vectorimage,linalg}, stereo/{adaptive,geodesic}weight
// and project/camera touch is provided: std headers and a small fixed-size matrix with Eigen's interface.
#ifndef SR_REF_PRECOMPILED_SHIM_HPP
#define SR_REF_PRECOMPILED_SHIM_HPP
#if defined(__cplusplus)
#define _USE_MATH_DEFINES
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <limits>
#include <memory>
#include <sstream>
#include <cstdint>
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

using std::fabs;
using std::sqrt;

// The reference was written on OS X (README.md), whose C++ library declares the floating-point
// overloads of abs in the GLOBAL namespace; glibc + libstdc++ declare only `int abs(int)` there, and
// the unqualified abs(sum1) of stereo/twoviewstereo.cpp:976 would silently truncate its double
// argument.  Provide what the author's toolchain provided.
inline double abs(double x) { return std::fabs(x); }
inline float abs(float x) { return std::fabs(x); }

// glibc's <math.h> declares a one-argument ::iszero template under _GNU_SOURCE (which g++ defines),
// ambiguous with the two-argument template of project/camera.cpp:50-53.  All standard headers are
// included above; from here on the identifier names the reference's own function.
#define iszero ref_iszero

// A fixed-size dense matrix with the part of Eigen's interface the reference's leaf sources and
// project/camera.cpp use: element access, vector algebra, products, transpose, the closed-form 3x3
// inverse (cofactors / determinant, as Eigen's compute_inverse<3>), column / row / corner views,
// the comma initialiser and a Householder QR.  Row-major storage; nothing here is reference code.
namespace Eigen {
enum { Upper = 1 };
template <class T, int R, int C> struct Matrix;

template <class T, int R, int C> struct CommaInit {
    Matrix<T, R, C> &M;
    int idx;
    CommaInit &operator,(T v) { M.m[idx++] = v; return *this; }
};
template <class T, int R, int C> struct ColView {  // M.col(c) of a non-const matrix
    Matrix<T, R, C> &M;
    int c;
    operator Matrix<T, R, 1>() const { Matrix<T, R, 1> v; for (int i = 0; i < R; ++i) v.m[i] = M.m[i * C + c]; return v; }
    ColView &operator=(const Matrix<T, R, 1> &v) { for (int i = 0; i < R; ++i) M.m[i * C + c] = v.m[i]; return *this; }
    ColView &operator-=(const Matrix<T, R, 1> &v) { for (int i = 0; i < R; ++i) M.m[i * C + c] -= v.m[i]; return *this; }
    void normalize() { Matrix<T, R, 1> v = *this; v.normalize(); *this = v; }
    T dot(const Matrix<T, R, 1> &o) const { return Matrix<T, R, 1>(*this).dot(o); }
};
template <class T, int R, int C> struct RowView {  // M.row(r)
    Matrix<T, R, C> &M;
    int r;
    operator Matrix<T, 1, C>() const { Matrix<T, 1, C> v; for (int i = 0; i < C; ++i) v.m[i] = M.m[r * C + i]; return v; }
    RowView &operator=(const Matrix<T, 1, C> &v) { for (int i = 0; i < C; ++i) M.m[r * C + i] = v.m[i]; return *this; }
    Matrix<T, 1, C> operator-() const { Matrix<T, 1, C> v; for (int i = 0; i < C; ++i) v.m[i] = -M.m[r * C + i]; return v; }
    template <int N> Matrix<T, N, 1> head() const { Matrix<T, N, 1> v; for (int i = 0; i < N; ++i) v.m[i] = M.m[r * C + i]; return v; }
};
template <class T, int R, int C, int BR, int BC> struct CornerView {  // M.topLeftCorner<BR,BC>()
    Matrix<T, R, C> &M;
    operator Matrix<T, BR, BC>() const { Matrix<T, BR, BC> v; for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) v.m[i * BC + j] = M.m[i * C + j]; return v; }
    CornerView &operator=(const Matrix<T, BR, BC> &v) { for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) M.m[i * C + j] = v.m[i * BC + j]; return *this; }
};

template <class T, int R, int C> struct Matrix {
    T m[R * C];
    Matrix() { for (int i = 0; i < R * C; ++i) m[i] = T(0); }
    Matrix(T x, T y, T z) { static_assert(R * C == 3, "3-vector constructor"); m[0] = x; m[1] = y; m[2] = z; }
    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() { Matrix I; for (int i = 0; i < (R < C ? R : C); ++i) I.m[i * C + i] = T(1); return I; }
    T &operator()(int r, int c) { return m[r * C + c]; }
    const T &operator()(int r, int c) const { return m[r * C + c]; }
    T &operator[](int i) { return m[i]; }
    const T &operator[](int i) const { return m[i]; }
    T x() const { return m[0]; }
    T y() const { return m[1]; }
    T z() const { return m[2]; }
    T dot(const Matrix &o) const { T s = m[0] * o.m[0]; for (int i = 1; i < R * C; ++i) s += m[i] * o.m[i]; return s; }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(squaredNorm()); }
    Matrix normalized() const { const T n = norm(); Matrix r; for (int i = 0; i < R * C; ++i) r.m[i] = m[i] / n; return r; }
    void normalize() { const T n = norm(); for (int i = 0; i < R * C; ++i) m[i] /= n; }
    Matrix &operator+=(const Matrix &o) { for (int i = 0; i < R * C; ++i) m[i] += o.m[i]; return *this; }
    Matrix &operator-=(const Matrix &o) { for (int i = 0; i < R * C; ++i) m[i] -= o.m[i]; return *this; }
    Matrix &operator*=(T s) { for (int i = 0; i < R * C; ++i) m[i] *= s; return *this; }
    Matrix &operator/=(T s) { for (int i = 0; i < R * C; ++i) m[i] /= s; return *this; }
    Matrix operator-() const { Matrix r; for (int i = 0; i < R * C; ++i) r.m[i] = -m[i]; return r; }
    bool operator==(const Matrix &o) const { for (int i = 0; i < R * C; ++i) if (!(m[i] == o.m[i])) return false; return true; }
    bool operator!=(const Matrix &o) const { return !(*this == o); }
    Matrix<T, C, R> transpose() const { Matrix<T, C, R> t; for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) t.m[j * R + i] = m[i * C + j]; return t; }
    Matrix inverse() const {  // 3x3: cofactor matrix transposed, times 1/det (Eigen's compute_inverse_size3)
        static_assert(R == 3 && C == 3, "closed-form inverse is 3x3 only");
        const Matrix &a = *this;
        auto cof = [&](int i, int j) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            return a(i1, j1) * a(i2, j2) - a(i1, j2) * a(i2, j1);
        };
        const T c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
        const T det = c00 * a(0, 0) + c10 * a(1, 0) + c20 * a(2, 0);
        const T invdet = T(1) / det;
        Matrix r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r(j, i) = cof(i, j) * invdet;
        return r;
    }
    ColView<T, R, C> col(int c) { return ColView<T, R, C>{*this, c}; }
    Matrix<T, R, 1> col(int c) const { Matrix<T, R, 1> v; for (int i = 0; i < R; ++i) v.m[i] = m[i * C + c]; return v; }
    RowView<T, R, C> row(int r) { return RowView<T, R, C>{*this, r}; }
    template <int BR, int BC> CornerView<T, R, C, BR, BC> topLeftCorner() { return CornerView<T, R, C, BR, BC>{*this}; }
    template <int BR, int BC> Matrix<T, BR, BC> topLeftCorner() const {
        Matrix<T, BR, BC> v;
        for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) v.m[i * BC + j] = m[i * C + j];
        return v;
    }
    CommaInit<T, R, C> operator<<(T v) { m[0] = v; return CommaInit<T, R, C>{*this, 1}; }
    template <int Mode> Matrix triangularView() const {  // Upper
        Matrix r;
        for (int i = 0; i < R; ++i) for (int j = i; j < C; ++j) r.m[i * C + j] = m[i * C + j];
        return r;
    }
};
template <class T, int R, int C> Matrix<T, R, C> operator+(const Matrix<T, R, C> &a, const Matrix<T, R, C> &b) { Matrix<T, R, C> r = a; r += b; return r; }
template <class T, int R, int C> Matrix<T, R, C> operator-(const Matrix<T, R, C> &a, const Matrix<T, R, C> &b) { Matrix<T, R, C> r = a; r -= b; return r; }
template <class T, int R, int C> Matrix<T, R, C> operator*(typename std::common_type<T>::type s, const Matrix<T, R, C> &a) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.m[i] = s * a.m[i]; return r; }
template <class T, int R, int C> Matrix<T, R, C> operator*(const Matrix<T, R, C> &a, typename std::common_type<T>::type s) { return s * a; }
template <class T, int R, int C> Matrix<T, R, C> operator/(const Matrix<T, R, C> &a, typename std::common_type<T>::type s) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.m[i] = a.m[i] / s; return r; }
template <class T, int R, int K, int C> Matrix<T, R, C> operator*(const Matrix<T, R, K> &a, const Matrix<T, K, C> &b) {
    Matrix<T, R, C> r;
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) {
            T s = a.m[i * K] * b.m[j];
            for (int k = 1; k < K; ++k) s += a.m[i * K + k] * b.m[k * C + j];
            r.m[i * C + j] = s;
        }
    return r;
}
template <class T, int R, int C> Matrix<T, R, 1> operator*(const Matrix<T, R, C> &a, const ColView<T, C, 4> &v) { return a * Matrix<T, C, 1>(v); }

// Householder QR (Golub & Van Loan 5.2.1); only Camera::setP's RQ factorisation uses it, and the
// glue configures cameras with Camera::set(K, R, t).
template <class M> struct HouseholderQR {
    M qr, q;
    explicit HouseholderQR(const M &A) : qr(A), q(M::Identity()) {
        const int n = 3;
        for (int k = 0; k < n; ++k) {
            double nrm = 0;
            for (int i = k; i < n; ++i) nrm += qr(i, k) * qr(i, k);
            nrm = std::sqrt(nrm);
            if (nrm == 0) continue;
            const double alpha = qr(k, k) > 0 ? -nrm : nrm;
            double v[3] = {0, 0, 0};
            for (int i = k; i < n; ++i) v[i] = qr(i, k);
            v[k] -= alpha;
            double vv = 0;
            for (int i = k; i < n; ++i) vv += v[i] * v[i];
            if (vv == 0) continue;
            for (int j = 0; j < n; ++j) {  // qr = H qr,  q = q H
                double s = 0, t = 0;
                for (int i = k; i < n; ++i) { s += v[i] * qr(i, j); t += q(j, i) * v[i]; }
                for (int i = k; i < n; ++i) { qr(i, j) -= 2 * s / vv * v[i]; q(j, i) -= 2 * t / vv * v[i]; }
            }
        }
    }
    M householderQ() const { return q; }
    const M &matrixQR() const { return qr; }
};

typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 3, 3> Matrix3d;
}  // namespace Eigen

#define FORWARD_DECLARE(cls) \
    class cls;               \
    typedef std::shared_ptr<cls> cls##Ptr; \
    typedef std::weak_ptr<cls> cls##WeakPtr
#endif
#endif
