// ref_glue_two.cpp — the reference's own TwoViewStereo (stereo/twoviewstereo.cpp), compiled where it
// lies and driven headless: constructor -> computeDepthMaps() (both directions of the live
// rasterised-curve search with its second-best test, then the cross-check), through the stand-ins
// of ref_shim/.  TEST INFRASTRUCTURE, see ref_glue_mvs.cpp; this file contains no reference code.
#define private public
#define protected public
#include "stereo/twoviewstereo.cpp"
#undef private
#undef protected

namespace {
struct cam_pod {  // == oracle.cpp: struct Camera == include/sr_b200.h: sr_camera
    double K[9], Kinv[9], R[9], Rinv[9], t[3], C[3], dist[5], plane_n[3], plane_d, n, prin_dir[3];
    int32_t is_refractive, is_distorted;
};
CameraPtr make_camera(const cam_pod *p, const char *id) {
    CameraPtr cam(new Camera(QString(id)));
    Eigen::Matrix3d K, R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { K(i, j) = p->K[3 * i + j]; R(i, j) = p->R[3 * i + j]; }
    cam->set(K, R, Eigen::Vector3d(p->t[0], p->t[1], p->t[2]));
    LensDistortions d;
    for (int i = 0; i < 5; ++i) d[i] = p->dist[i];
    cam->setLensDistortion(d);
    cam->setRefractiveIndex(p->n);
    cam->setPlane(Plane3d(Eigen::Vector3d(p->plane_n[0], p->plane_n[1], p->plane_n[2]), p->plane_d));
    return cam;
}
QImage image_of(const uint8_t *rgba8, int w, int h) {
    QImage q(w, h, QImage::Format_ARGB32);
    for (int y = 0; y < h; ++y) {
        QRgb *line = reinterpret_cast<QRgb *>(q.scanLine(y));
        for (int x = 0; x < w; ++x) {
            const uint8_t *p = rgba8 + 4 * ((size_t)y * w + x);
            line[x] = qRgba(p[0], p[1], p[2], p[3]);
        }
    }
    return q;
}
QImage mask_of(const uint8_t *mask8, int w, int h) {  // 255 -> WHITE, anything else -> BLACK; null -> null image
    if (!mask8) return QImage();
    QImage q(w, h, QImage::Format_ARGB32);
    for (int y = 0; y < h; ++y) {
        QRgb *line = reinterpret_cast<QRgb *>(q.scanLine(y));
        for (int x = 0; x < w; ++x) {
            const int v = mask8[(size_t)y * w + x] == 255 ? 255 : 0;
            line[x] = qRgba(v, v, v, 255);
        }
    }
    return q;
}
}  // namespace

struct ref_two {
    TwoViewStereo *task;
    int w, h;
};

extern "C" {
// Images are handed over at their final size (the constructor's scaledToWidth is the identity at
// scale 1); imageScale, the factor between calibrated and image pixels, is set afterwards.
ref_two *ref_two_create(const void *left_cam, const void *right_cam, const uint8_t *left_rgba, const uint8_t *right_rgba,
                        const uint8_t *left_mask, const uint8_t *right_mask, int w, int h, double minDepth, double maxDepth,
                        int numDepthLevels, double imageScale) {
    ref_two *t = new ref_two;
    t->w = w;
    t->h = h;
    t->task = new TwoViewStereo(make_camera(static_cast<const cam_pod *>(left_cam), "left"), image_of(left_rgba, w, h),
                                mask_of(left_mask, w, h), make_camera(static_cast<const cam_pod *>(right_cam), "right"),
                                image_of(right_rgba, w, h), mask_of(right_mask, w, h), minDepth, maxDepth, numDepthLevels, 1.0);
    t->task->imageScale = imageScale;
    t->task->cancelled = false;  // Task leaves it uninitialised (gui/task.hpp:100)
    return t;
}
void ref_two_destroy(ref_two *t) {
    delete t->task;
    delete t;
}
// TwoViewStereo::computeCostVolumes (twoviewstereo.cpp:233-500): both directions of the search,
// depths before the cross-check (NaN: nothing evaluated, +INF: second-best test failed)
void ref_two_search(ref_two *t, double *left_depth, double *right_depth) {
    t->task->computeCostVolumes(t->task->leftView, t->task->rightView);
    const size_t n = (size_t)t->w * t->h;
    std::memcpy(left_depth, t->task->computedDepthLeft.data(), n * sizeof(double));
    std::memcpy(right_depth, t->task->computedDepthRight.data(), n * sizeof(double));
}
// TwoViewStereo::computeDepthMaps (:150-227): search + crossCheck (:596-672)
void ref_two_run(ref_two *t, double *left_depth, double *right_depth) {
    t->task->computeDepthMaps();
    const size_t n = (size_t)t->w * t->h;
    std::memcpy(left_depth, t->task->computedDepthLeft.data(), n * sizeof(double));
    std::memcpy(right_depth, t->task->computedDepthRight.data(), n * sizeof(double));
}
// cost_ncc (:909-977) / cost_sad (:864-905) of the left pixel (x1,y1) against the right pixel (x2,y2)
// (dir 0) or the other way round (dir 1), GeodesicWeight r = 5
double ref_two_cost(ref_two *t, int sad, int dir, int x1, int y1, int x2, int y2) {
    TwoViewStereo &s = *t->task;
    const VectorImage &a = dir ? s.right : s.left, &b = dir ? s.left : s.right;
    const VectorImage &am = dir ? s.rightMask : s.leftMask, &bm = dir ? s.leftMask : s.rightMask;
    weightFuncs[0].init_weights(a, x1, y1);
    return sad ? s.cost_sad(a, b, am, bm, x1, y1, x2, y2) : s.cost_ncc(a, b, am, bm, x1, y1, x2, y2);
}
}
