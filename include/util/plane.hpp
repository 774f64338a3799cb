// util/plane.hpp — Plane3d with the reference's interface (util/plane.hpp:26-47).
#ifndef SR_UTIL_PLANE_HPP
#define SR_UTIL_PLANE_HPP
#include "util/precompiled.hpp"
class Plane3d {
public:
    typedef Eigen::Vector3d Point;
    typedef Eigen::Vector3d Vector;
    Plane3d() : normal_(0, 0, 1), dist_(0) {}
    Plane3d(const Vector &normal, double d) : normal_(normal.normalized()), dist_(d) {}
    Plane3d(const Vector &normal, const Point &x0) : normal_(normal.normalized()), dist_(normal_.dot(x0)) {}
    void setNormal(const Vector &normal) { normal_ = normal.normalized(); }
    void setDistance(double dist) { dist_ = dist; }
    const Vector &normal() const { return normal_; }
    double distance() const { return dist_; }
    Point x0() const { return dist_ * normal_; }
private:
    Vector normal_;
    double dist_;
};
inline bool operator==(const Plane3d &a, const Plane3d &b) {
    return (a.normal() - b.normal()).squaredNorm() + std::fabs(a.distance() - b.distance()) < 1e-10;
}
inline bool operator!=(const Plane3d &a, const Plane3d &b) { return !(a == b); }
#endif
