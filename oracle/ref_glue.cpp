// ref_glue.cpp — extern "C" entry points over the REFERENCE'S OWN classes, compiled together
// with the reference's sources read from /root/reference (see Makefile target _ref/libref.so).
// TEST INFRASTRUCTURE: lets tests/test_oracle_vs_ref.py pin oracle/oracle.cpp against the real
// LineIterator/clipLine, Ray3d/intersect/refract/closestPoints, VectorImage::pixel/sample,
// AdaptiveWeight, GeodesicWeight and Camera (project / projectRefraction / unproject / set).  This
// file contains no reference code, only calls.
#include "util/lineiter.hpp"
#include "util/ray.hpp"
#include "util/vectorimage.hpp"
#include "stereo/adaptiveweight.hpp"
#include "stereo/geodesicweight.hpp"
#include "project/camera.hpp"
#include <cstdint>

typedef Eigen::Vector3d V3;

extern "C" {

int ref_line(int x0, int y0, int x1, int y1, int clip, int w, int h, int32_t *out_xy, int max_pts) {
    LineIterator it = clip ? LineIterator(x0, y0, x1, y1, w, h) : LineIterator(x0, y0, x1, y1);
    int n = 0;
    while (it.hasNext()) {
        int tx, ty;
        it.current(tx, ty);
        if (n < max_pts) { out_xy[2 * n] = tx; out_xy[2 * n + 1] = ty; }
        ++n;
        ++it;
        if (n > (1 << 24)) break;
    }
    return n;
}
int ref_clip_line(int32_t *xyxy, int w, int h) {
    int x0 = xyxy[0], y0 = xyxy[1], x1 = xyxy[2], y1 = xyxy[3];
    bool ok = clipLine(x0, y0, x1, y1, w, h);
    xyxy[0] = x0; xyxy[1] = y0; xyxy[2] = x1; xyxy[3] = y1;
    return ok;
}
int ref_intersect(const double *src, const double *dir, const double *pn, double pd, double *out3) {
    Ray3d R(V3(src[0], src[1], src[2]), V3(dir[0], dir[1], dir[2]));
    Plane3d P(V3(pn[0], pn[1], pn[2]), pd);
    V3 p(NAN, NAN, NAN);
    bool ok = intersect(R, P, p);
    out3[0] = p[0]; out3[1] = p[1]; out3[2] = p[2];
    return ok;
}
int ref_refract(const double *src, const double *dir, const double *pn, double pd, double n, double *out6) {
    Ray3d R(V3(src[0], src[1], src[2]), V3(dir[0], dir[1], dir[2]));
    Plane3d P(V3(pn[0], pn[1], pn[2]), pd);
    Ray3d O = R;
    bool ok = refract(R, P, n, O);
    out6[0] = O.source()[0]; out6[1] = O.source()[1]; out6[2] = O.source()[2];
    out6[3] = O.direction()[0]; out6[4] = O.direction()[1]; out6[5] = O.direction()[2];
    return ok;
}
void ref_closest_points(const double *s1, const double *d1, const double *s2, const double *d2, double *out6) {
    Ray3d A(V3(s1[0], s1[1], s1[2]), V3(d1[0], d1[1], d1[2]));
    Ray3d B(V3(s2[0], s2[1], s2[2]), V3(d2[0], d2[1], d2[2]));
    V3 p1, p2;
    A.closestPoints(B, p1, p2);
    out6[0] = p1[0]; out6[1] = p1[1]; out6[2] = p1[2];
    out6[3] = p2[0]; out6[4] = p2[1]; out6[5] = p2[2];
}

// images: RGBA8 (R,G,B,A bytes) -> shim QImage (ARGB32 words) -> the reference's fromQImage
struct ref_image { VectorImage img; };
ref_image *ref_image_create(const uint8_t *rgba8, int w, int h) {
    QImage q(w, h, QImage::Format_ARGB32);
    for (int y = 0; y < h; ++y) {
        QRgb *line = reinterpret_cast<QRgb *>(q.scanLine(y));
        for (int x = 0; x < w; ++x) {
            const uint8_t *p = rgba8 + 4 * ((size_t)y * w + x);
            line[x] = qRgba(p[0], p[1], p[2], p[3]);
        }
    }
    ref_image *r = new ref_image;
    r->img = VectorImage::fromQImage(q);
    return r;
}
void ref_image_destroy(ref_image *r) { delete r; }
void ref_pixel(const ref_image *r, int x, int y, double *out4) {
    const RGBA &c = r->img.pixel(x, y);
    out4[0] = c.r; out4[1] = c.g; out4[2] = c.b; out4[3] = c.a;
}
void ref_sample(const ref_image *r, double x, double y, double *out4) {
    RGBA c = r->img.sample(x, y);
    out4[0] = c.r; out4[1] = c.g; out4[2] = c.b; out4[3] = c.a;
}
double ref_to_gray(const ref_image *r, int x, int y) { return r->img.pixel(x, y).toGray(); }
int ref_is_white(const ref_image *r, int x, int y) { return r->img.pixel(x, y) == WHITE; }

// kind 0: AdaptiveWeight, 1: GeodesicWeight; out = n*(2r+1)^2 doubles [row+r][col+r]
void ref_weights(const ref_image *r, int kind, int radius, int n, const int32_t *cx, const int32_t *cy, double *out) {
    const int wn = 2 * radius + 1;
    AdaptiveWeight aw(radius);
    GeodesicWeight gw(radius);
    for (int i = 0; i < n; ++i) {
        if (kind == 0) aw.init_weights(r->img, cx[i], cy[i]);
        else gw.init_weights(r->img, cx[i], cy[i]);
        for (int row = -radius; row <= radius; ++row)
            for (int col = -radius; col <= radius; ++col)
                out[((size_t)i * wn + (row + radius)) * wn + (col + radius)] = (kind == 0) ? aw(row, col) : gw(row, col);
    }
}

// ---- project/camera.cpp: the reference's own Camera, configured from the oracle's camera POD ----
// POD layout (oracle.cpp: struct Camera == include/sr_b200.h: sr_camera): K[9] Kinv[9] R[9] Rinv[9]
// t[3] C[3] dist[5] plane_n[3] plane_d n prin_dir[3] is_refractive is_distorted.
struct ref_cam_pod {
    double K[9], Kinv[9], R[9], Rinv[9], t[3], C[3], dist[5], plane_n[3], plane_d, n, prin_dir[3];
    int32_t is_refractive, is_distorted;
};
static void ref_configure(Camera &cam, const ref_cam_pod *p) {
    Eigen::Matrix3d K, R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { K(i, j) = p->K[3 * i + j]; R(i, j) = p->R[3 * i + j]; }
    cam.set(K, R, V3(p->t[0], p->t[1], p->t[2]));
    LensDistortions d;
    for (int i = 0; i < 5; ++i) d[i] = p->dist[i];
    cam.setLensDistortion(d);
    cam.setRefractiveIndex(p->n);
    cam.setPlane(Plane3d(V3(p->plane_n[0], p->plane_n[1], p->plane_n[2]), p->plane_d));
}
// What Camera::set / setLensDistortion / setPlane / setRefractiveIndex derive: the same POD back.
void ref_camera_derived(const ref_cam_pod *in, ref_cam_pod *out) {
    Camera cam("ref");
    ref_configure(cam, in);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            out->K[3 * i + j] = cam.K()(i, j); out->Kinv[3 * i + j] = cam.Kinv()(i, j);
            out->R[3 * i + j] = cam.R()(i, j); out->Rinv[3 * i + j] = cam.Rinv()(i, j);
        }
    for (int i = 0; i < 3; ++i) {
        out->t[i] = cam.t()[i]; out->C[i] = cam.C()[i];
        out->plane_n[i] = cam.plane().normal()[i];
        out->prin_dir[i] = cam.principleRay().direction()[i];
    }
    for (int i = 0; i < 5; ++i) out->dist[i] = cam.lensDistortion()[i];
    out->plane_d = cam.plane().distance();
    out->n = cam.refractiveIndex();
    out->is_refractive = cam.isRefractive();
    out->is_distorted = cam.isDistorted();
}
// Camera::project (camera.cpp:380-419) of n global points
void ref_camera_project(const ref_cam_pod *pod, int n, const double *xyz, double *out_xy, int32_t *out_ok) {
    Camera cam("ref");
    ref_configure(cam, pod);
    for (int i = 0; i < n; ++i) {
        V3 p(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        out_ok[i] = cam.project(p) ? 1 : 0;
        out_xy[2 * i] = p[0];
        out_xy[2 * i + 1] = p[1];
    }
}
// Camera::unproject (camera.cpp:423-459) of n pixels: out = n * (source xyz, direction xyz)
void ref_camera_unproject(const ref_cam_pod *pod, int n, const double *xy, double *out6) {
    Camera cam("ref");
    ref_configure(cam, pod);
    for (int i = 0; i < n; ++i) {
        Ray3d r = cam.unproject(xy[2 * i], xy[2 * i + 1]);
        for (int k = 0; k < 3; ++k) { out6[6 * i + k] = r.source()[k]; out6[6 * i + 3 + k] = r.direction()[k]; }
    }
}
// Camera::setP -> updateOthers (camera.cpp:251-288): K, R, t back out of a 3x4 matrix (row-major in)
void ref_camera_from_P(const double *P12, ref_cam_pod *out) {
    Camera cam("ref");
    ProjMat P;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) P(i, j) = P12[4 * i + j];
    cam.setP(P);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            out->K[3 * i + j] = cam.K()(i, j); out->Kinv[3 * i + j] = cam.Kinv()(i, j);
            out->R[3 * i + j] = cam.R()(i, j); out->Rinv[3 * i + j] = cam.Rinv()(i, j);
        }
    for (int i = 0; i < 3; ++i) { out->t[i] = cam.t()[i]; out->C[i] = cam.C()[i]; out->prin_dir[i] = cam.principleRay().direction()[i]; }
}
}

// What moc would generate for Camera's signals (project/camera.hpp:146-152): nobody is connected.
void Camera::nameChanged(QString) {}
void Camera::intrinsicParametersChanged(const Eigen::Matrix3d &) {}
void Camera::extrinsicParametersChanged(const Eigen::Matrix3d &, const Eigen::Vector3d &) {}
void Camera::lensDistortionChanged(const LensDistortions &) {}
void Camera::responseChanged(const Responses &) {}
void Camera::refractiveParametersChanged(const Plane3d &, double) {}
