// project/camera.hpp — Camera with the reference's interface (project/camera.hpp:39-186):
// pinhole + OpenCV 5-coefficient distortion + planar refractive interface.  This is the input
// data model of the dense-matching path; toPod() snapshots it into the sr_camera the C ABI takes.
// Qt-free: the reference's signals (intrinsicParametersChanged, ...) are dropped.
#ifndef SR_PROJECT_CAMERA_HPP
#define SR_PROJECT_CAMERA_HPP
#include "util/precompiled.hpp"
#include "util/ray.hpp"
#include "sr_b200.h"
#include <cstring>

FORWARD_DECLARE(Camera);
typedef std::array<double, 5> LensDistortions;

class Camera {
public:
    explicit Camera(const std::string &id = std::string(), const std::string &name = std::string())
        : id_(id), name_(name.empty() ? "<no name>" : name), R_(Eigen::Matrix3d::Identity()), K_(Eigen::Matrix3d::Identity()),
          Rinv_(Eigen::Matrix3d::Identity()), Kinv_(Eigen::Matrix3d::Identity()), refractiveIndex_(1.0), isRefractive_(false),
          isDistorted_(false) {
        lensDistortion_.fill(0.0);
        P_(0, 0) = P_(1, 1) = P_(2, 2) = 1.0;
    }
    const std::string &id() const { return id_; }
    const std::string &name() const { return name_; }
    void setName(const std::string &n) { name_ = n; }

    const ProjMat &P() const { return P_; }
    const Eigen::Matrix3d &K() const { return K_; }
    const Eigen::Matrix3d &Kinv() const { return Kinv_; }
    const Eigen::Matrix3d &R() const { return R_; }
    const Eigen::Matrix3d &Rinv() const { return Rinv_; }
    const Eigen::Vector3d &t() const { return t_; }
    const Eigen::Vector3d &C() const { return C_; }
    const Ray3d &principleRay() const { return principleRay_; }
    const Plane3d &plane() const { return plane_; }
    double refractiveIndex() const { return refractiveIndex_; }
    bool isRefractive() const { return isRefractive_; }
    bool isDistorted() const { return isDistorted_; }
    const LensDistortions &lensDistortion() const { return lensDistortion_; }

    // project/camera.cpp:205-222
    void set(const Eigen::Matrix3d &K, const Eigen::Matrix3d &R, const Eigen::Vector3d &t) {
        K_ = K; R_ = R; t_ = t;
        orthonormalize(R_);
        Kinv_ = K_.inverse();
        Rinv_ = R_.transpose();
        C_ = Rinv_ * (-t);
        updateProjection();
    }
    // project/camera.cpp:169-174 + updateOthers :251-288 (RQ factorisation of the 3x4 matrix)
    void setP(const ProjMat &P) {
        P_ = P;
        updateOthers();
    }
    void setK(const Eigen::Matrix3d &K) { K_ = K; Kinv_ = K_.inverse(); updateProjection(); }
    void setR(const Eigen::Matrix3d &R) { R_ = R; orthonormalize(R_); Rinv_ = R_.transpose(); updateProjection(); }
    void sett(const Eigen::Vector3d &t) { t_ = t; C_ = Rinv_ * (-t); updateProjection(); }
    void setC(const Eigen::Vector3d &C) { C_ = C; t_ = R_ * (-C); updateProjection(); }
    // project/camera.cpp:302-312
    void setLensDistortion(const LensDistortions &d) {
        lensDistortion_ = d;
        isDistorted_ = false;
        for (double v : d) isDistorted_ = isDistorted_ || !iszero(v);
    }
    // project/camera.cpp:326-342
    void setPlane(const Plane3d &plane) {
        plane_ = plane;
        isRefractive_ = (!iszero(refractiveIndex_ - 1) && !iszero(plane_.distance()));
    }
    void setRefractiveIndex(double n) {
        refractiveIndex_ = n;
        isRefractive_ = (!iszero(refractiveIndex_ - 1) && !iszero(plane_.distance()));
    }

    Eigen::Vector3d fromGlobalToLocal(const Eigen::Vector3d &p) const { return R_ * p + t_; }
    Eigen::Vector3d fromLocalToGlobal(const Eigen::Vector3d &p) const { return Rinv_ * (p - t_); }
    Ray3d fromLocalToGlobal(const Ray3d &r) const { return Ray3d(fromLocalToGlobal(r.source()), Rinv_ * r.direction()); }
    Ray3d fromGlobalToLocal(const Ray3d &r) const { return Ray3d(fromGlobalToLocal(r.source()), R_ * r.direction()); }

    // project/camera.cpp:423-459 (host; the bulk version is sr_unproject_grid)
    Ray3d unproject(double px, double py) const { return unproject(Eigen::Vector3d(px, py, 1.0)); }
    Ray3d unproject(const Eigen::Vector3d &p) const {
        Eigen::Vector3d pp = p;
        if (isDistorted_) {
            const double cx = K_(0, 2), cy = K_(1, 2), ifx = 1.0 / K_(0, 0), ify = 1.0 / K_(1, 1);
            double x = (pp[0] - cx) * ifx, y = (pp[1] - cy) * ify;
            const double x0 = x, y0 = y;
            const LensDistortions &k = lensDistortion_;
            for (int j = 0; j < 5; j++) {
                const double r2 = x * x + y * y;
                const double icdist = 1.0 / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
                const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
                const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
                x = (x0 - deltaX) * icdist;
                y = (y0 - deltaY) * icdist;
            }
            x /= ifx; y /= ify;
            pp[0] = x + cx; pp[1] = y + cy;
        }
        Ray3d ray(Ray3d::Point::Zero(), Kinv_ * pp);
        if (isRefractive_) refract(ray, plane_, refractiveIndex_, ray);
        return fromLocalToGlobal(ray);
    }
    // project/camera.cpp:380-419 (host; the bulk version is sr_project_points).  The refractive
    // branch solves the un-squared Snell equation, whose unique root in [0,r] is the root the
    // reference's quartic + acceptance test selects.
    bool project(Eigen::Vector3d &p) const {
        Eigen::Vector3d point = fromGlobalToLocal(p);
        if (isRefractive_ && !projectRefraction(point)) {
            p = Eigen::Vector3d(NAN, NAN, NAN);
            return false;
        }
        p = K_ * point;
        p = p / p.z();
        if (isDistorted_) {
            const double cx = K_(0, 2), cy = K_(1, 2), fx = K_(0, 0), fy = K_(1, 1);
            double x = (p[0] - cx) / fx, y = (p[1] - cy) / fy;
            const LensDistortions &k = lensDistortion_;
            const double r2 = x * x + y * y;
            const double cdist = 1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2;
            const double xo = x, yo = y;
            x = xo * cdist + 2 * k[2] * xo * yo + k[3] * (r2 + 2 * xo * xo);
            y = yo * cdist + k[2] * (r2 + 2 * yo * yo) + 2 * k[3] * x * yo;  // uses the distorted x (:411-412)
            p[0] = fx * x + cx;
            p[1] = fy * y + cy;
        }
        return true;
    }
    bool project(const Eigen::Vector3d &p, double &x, double &y) const {
        Eigen::Vector3d q = p;
        const bool ok = project(q);
        x = q[0];
        y = q[1];
        return ok;
    }

    //! Snapshot for the C ABI (include/sr_b200.h::sr_camera).
    sr_camera toPod() const {
        sr_camera c;
        std::memset(&c, 0, sizeof(c));
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                c.K[3 * i + j] = K_(i, j); c.Kinv[3 * i + j] = Kinv_(i, j);
                c.R[3 * i + j] = R_(i, j); c.Rinv[3 * i + j] = Rinv_(i, j);
            }
        for (int i = 0; i < 3; ++i) {
            c.t[i] = t_[i]; c.C[i] = C_[i];
            c.plane_n[i] = plane_.normal()[i];
            c.prin_dir[i] = principleRay_.direction()[i];
        }
        for (int i = 0; i < 5; ++i) c.dist[i] = lensDistortion_[i];
        c.plane_d = plane_.distance();
        c.n = refractiveIndex_;
        c.is_refractive = isRefractive_;
        c.is_distorted = isDistorted_;
        return c;
    }

private:
    static bool iszero(double x, double eps = 1e-10) { return (x <= eps && x >= -eps); }
    // Gram-Schmidt on columns, project/camera.cpp:143-165
    static void orthonormalize(Eigen::Matrix3d &mat) {
        for (int i = 0; i < 3; ++i) {
            Eigen::Vector3d accum = Eigen::Vector3d::Zero();
            for (int j = 0; j < i; ++j) {
                Eigen::Vector3d vi = mat.col(i), vj = mat.col(j);
                accum += vj * (vi.dot(vj) / vj.squaredNorm());
            }
            Eigen::Vector3d c = mat.col(i) - accum;
            c.normalize();
            for (int r = 0; r < 3; ++r) mat(r, i) = c[r];
        }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                if (-1e-10 < mat(i, j) && mat(i, j) < 1e-10) mat(i, j) = 0.0;
    }
    void updateProjection() {  // project/camera.cpp:244-249
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) P_(i, j) = K_(i, 0) * R_(0, j) + K_(i, 1) * R_(1, j) + K_(i, 2) * R_(2, j);
            P_(i, 3) = K_(i, 0) * t_[0] + K_(i, 1) * t_[1] + K_(i, 2) * t_[2];
        }
        updatePrincipleRay();
    }
    void updatePrincipleRay() {  // project/camera.cpp:292-298
        const Eigen::Vector3d tcol = K_.col(2);
        const Eigen::Vector3d dir = Kinv_ * (tcol / tcol[2]);
        principleRay_.setSource(C_);
        principleRay_.setDirection(Rinv_ * dir.normalized());
    }
    // project/camera.cpp:251-288: P /= |P.row(2).head<3>()|^2; RQ by Householder QR of the
    // row-reversed transpose; sign fix-ups; orthonormalise; t = Kinv P.col(3); C = -Rinv t.
    void updateOthers() {
        double sn = P_(2, 0) * P_(2, 0) + P_(2, 1) * P_(2, 1) + P_(2, 2) * P_(2, 2);
        for (double &x : P_.m) x /= sn;
        // A = (reverseRows * M)^T, columns a0,a1,a2;  Householder QR: A = Q Rq
        double A[3][3], Q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) A[i][j] = P_(2 - j, i);
        for (int k = 0; k < 2; ++k) {  // reflections, beta = -sign(x0)*|x| as Eigen/LAPACK
            double nrm = 0;
            for (int i = k; i < 3; ++i) nrm += A[i][k] * A[i][k];
            nrm = std::sqrt(nrm);
            double tail = 0;
            for (int i = k + 1; i < 3; ++i) tail += A[i][k] * A[i][k];
            if (tail == 0.0) continue;
            const double beta = (A[k][k] >= 0) ? -nrm : nrm;
            double v[3] = {0, 0, 0};
            v[k] = A[k][k] - beta;
            for (int i = k + 1; i < 3; ++i) v[i] = A[i][k];
            double vv = 0;
            for (int i = k; i < 3; ++i) vv += v[i] * v[i];
            for (int j = 0; j < 3; ++j) {  // A <- (I - 2 v v^T / vv) A
                double s = 0;
                for (int i = k; i < 3; ++i) s += v[i] * A[i][j];
                for (int i = k; i < 3; ++i) A[i][j] -= 2 * v[i] * s / vv;
            }
            for (int i = 0; i < 3; ++i) {  // Q <- Q (I - 2 v v^T / vv)
                double s = 0;
                for (int j = k; j < 3; ++j) s += Q[i][j] * v[j];
                for (int j = k; j < 3; ++j) Q[i][j] -= 2 * s * v[j] / vv;
            }
        }
        // R_ = reverseRows * Q^T ; K_ = reverseRows * Rq^T * reverseRows
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                R_(i, j) = Q[j][2 - i];
                K_(i, j) = (2 - j <= 2 - i) ? A[2 - j][2 - i] : 0.0;  // Rq upper triangular
            }
        for (int axis = 2; axis >= 0; --axis) {
            if (K_(axis, axis) < 0) {
                K_(axis, axis) = -K_(axis, axis);
                for (int j = 0; j < 3; ++j) R_(axis, j) = -R_(axis, j);
            }
            if (K_(axis, 2) < 0) K_(axis, 2) = -K_(axis, 2);
        }
        orthonormalize(R_);
        Kinv_ = K_.inverse();
        Rinv_ = R_.transpose();
        t_ = Kinv_ * Eigen::Vector3d(P_(0, 3), P_(1, 3), P_(2, 3));
        C_ = -(Rinv_ * t_);
        updatePrincipleRay();
    }
    // project/camera.cpp:95-138 with the monotone solve in place of the GSL quartic.
    bool projectRefraction(Eigen::Vector3d &p) const {
        const Eigen::Vector3d N = plane_.normal();
        const Eigen::Vector3d proj = N.dot(p) * N;
        const Eigen::Vector3d radv = p - proj;
        const double z = proj.norm(), r = radv.norm(), d = plane_.distance(), n = refractiveIndex_;
        if (!(r > 0)) return false;
        const double h = z - d, dd = d * d, hh = h * h;
        double lo = 0, hi = r, x = n * std::fabs(d) * r / (std::fabs(h) + n * std::fabs(d) + 1e-300);
        if (!(x >= lo && x <= hi)) x = 0.5 * r;
        for (int it = 0; it < 100; ++it) {
            const double rx = r - x, ia = 1.0 / std::sqrt(x * x + dd), ib = 1.0 / std::sqrt(rx * rx + hh);
            const double g = x * ia - n * rx * ib;
            if (g == 0) break;
            if (g < 0) lo = x; else hi = x;
            const double step = g / (dd * ia * ia * ia + n * hh * ib * ib * ib);
            double xn = x - step;
            if (std::fabs(step) <= 4e-16 * r) { x = xn; break; }
            if (!(xn >= lo && xn <= hi)) xn = 0.5 * (lo + hi);
            x = xn;
            if (hi - lo <= 4e-16 * r) break;
        }
        p = (x / r) * radv + plane_.x0();
        return true;
    }

    std::string id_, name_;
    ProjMat P_;
    Eigen::Vector3d t_, C_;
    Eigen::Matrix3d R_, K_, Rinv_, Kinv_;
    LensDistortions lensDistortion_;
    Plane3d plane_;
    double refractiveIndex_;
    bool isRefractive_, isDistorted_;
    Ray3d principleRay_;
};
#endif
