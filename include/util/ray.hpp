// util/ray.hpp — Ray3d, intersect, refract, midpoint with the reference's interface
// (util/ray.hpp:30-98).  Input data model of the dense-matching path (host side).
#ifndef SR_UTIL_RAY_HPP
#define SR_UTIL_RAY_HPP
#include "util/plane.hpp"
class Ray3d {
public:
    typedef Eigen::Vector3d Point;
    typedef Eigen::Vector3d Vector;
    Ray3d() : source_(0, 0, 0), dir_(0, 0, 1) {}
    Ray3d(const Point &source, const Vector &dir) : source_(source), dir_(dir.normalized()) {}
    void setSource(const Point &p) { source_ = p; }
    void setDirection(const Vector &v) { dir_ = v.normalized(); }
    const Point &source() const { return source_; }
    const Vector &direction() const { return dir_; }
    Point point(double dist) const { return source_ + dist * dir_; }
    // util/ray.cpp:53-74
    void closestPoints(const Ray3d &ray, Point &p1, Point &p2) const {
        Point w0 = source() - ray.source();
        double a = direction().dot(direction()), b = direction().dot(ray.direction());
        double c = ray.direction().dot(ray.direction()), d = direction().dot(w0), e = ray.direction().dot(w0);
        double den = 1.0 / (a * c - b * b);
        double tl = (b * e - c * d) * den, tr = (a * e - b * d) * den;
        p1 = source();
        p2 = ray.source();
        if (tl > 0) p1 += tl * direction();
        if (tr > 0) p2 += tr * ray.direction();
    }
    Point closestPoint(const Ray3d &ray) const { Point p, s; closestPoints(ray, p, s); return p; }
    double distance(const Ray3d &ray) const { Point a, b; closestPoints(ray, a, b); return (a - b).norm(); }
private:
    Point source_;
    Vector dir_;
};
// util/ray.cpp:78-88
inline bool intersect(const Ray3d &R, const Plane3d &P, Ray3d::Point &p) {
    double nd = P.normal().dot(R.direction());
    if (std::fabs(nd) < 1e-10) return false;
    double t = P.normal().dot(P.x0() - R.source()) / nd;
    if (t < 1e-10) return false;
    p = R.point(t);
    return true;
}
// util/ray.cpp:92-106
inline bool refract(const Ray3d &R, const Plane3d &P, double n, Ray3d &Rout) {
    Ray3d::Point p;
    if (intersect(R, P, p)) {
        double cosI = -(P.normal().dot(R.direction()));
        double cosT2 = 1.0 - (1.0 - cosI * cosI) / (n * n);
        if (cosT2 > 0.0) {
            double sign = (cosI > 0.0 ? -1.0 : 1.0);
            Ray3d::Vector d = R.direction() + (cosI + n * sign * std::sqrt(cosT2)) * P.normal();
            Rout.setSource(p);
            Rout.setDirection(d);
            return true;
        }
    }
    return false;
}
inline Ray3d::Point midpoint(const Ray3d &R1, const Ray3d &R2) {
    Ray3d::Point a, b;
    R1.closestPoints(R2, a, b);
    return (a + b) / 2;
}
#endif
