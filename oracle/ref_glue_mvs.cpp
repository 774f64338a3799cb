// ref_glue_mvs.cpp — the reference's own MultiViewStereo (stereo/multiviewstereo.cpp), compiled where
// it lies and driven headless: initialize() -> runTask() exactly as the GUI thread does
// (gui/widgets/stereowidget.cpp:985-1000), through stand-ins for Qt, Project/ImageSet and Eigen
// (ref_shim/).  TEST INFRASTRUCTURE: gives tests/test_oracle_vs_ref.py the reference's END-TO-END
// result — neighbour selection, rasterised epipolar curves, weighted NCC, the K = 9 peak lists, the
// selection rule and the cross-check — to pin oracle/oracle.cpp against.  This file contains no
// reference code: the translation unit includes the reference's .cpp so that its file-local
// functions (cost_ncc) and the class's protected members can be called from here.  Built with
// -DUSE_TBB: the reference's own tbb::parallel_for over image rows (multiviewstereo.cpp:548,675) is what
// parallelises it (ref_shim/tbb/tbb.h stands in for the library).
#define private public
#define protected public
#include "stereo/multiviewstereo.cpp"
#undef private
#undef protected
#include <cstdint>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REF_MVS_ADAPTIVE_TU  // (ref_glue_mvs_ada.cpp compiles this file a second time, see there)
// What moc would generate for Task's signals (gui/task.hpp:87-97): nobody is connected.
void Task::started(const Task *) {}
void Task::finished(const Task *) {}
void Task::progressUpdate(int) {}
void Task::stageUpdate(QString) {}
#endif

namespace {
struct cam_pod {  // == oracle.cpp: struct Camera == include/sr_b200.h: sr_camera
    double K[9], Kinv[9], R[9], Rinv[9], t[3], C[3], dist[5], plane_n[3], plane_d, n, prin_dir[3];
    int32_t is_refractive, is_distorted;
};
CameraPtr make_camera(const cam_pod *p, int index) {
    CameraPtr cam(new Camera(QString(std::string("cam") + std::to_string(index))));
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
}  // namespace

struct ref_mvs {
    MultiViewStereo task;
    std::vector<CameraPtr> cams;
};

extern "C" {

// V views: camera PODs + RGBA8 images whose alpha byte is the mask (255 = WHITE), as the reference
// reads a PNG (multiviewstereo.cpp:216-240).  The images are handed over at their final size:
// initialize() runs with imageScale 1 (its QImage::scaledToWidth is the identity) and imageScale —
// the factor between calibrated and image pixels, (x + 0.5) / imageScale — is set afterwards
// ("identically pre-scaled inputs").
ref_mvs *ref_mvs_create(int V, const void *cams_pod, const uint8_t *const *rgba8, int w, int h, double minDepth,
                        double maxDepth, int numDepthLevels, double crossCheckThreshold, double imageScale) {
    ref_mvs *m = new ref_mvs;
    ImageSetPtr set(new ImageSet);
    const cam_pod *pods = static_cast<const cam_pod *>(cams_pod);
    for (int v = 0; v < V; ++v) {
        m->cams.push_back(make_camera(pods + v, v));
        const std::string name = "mem://view" + std::to_string(v) + "@" + std::to_string((size_t)(void *)m);
        QImage q(w, h, QImage::Format_ARGB32);
        for (int y = 0; y < h; ++y) {
            QRgb *line = reinterpret_cast<QRgb *>(q.scanLine(y));
            for (int x = 0; x < w; ++x) {
                const uint8_t *p = rgba8[v] + 4 * ((size_t)y * w + x);
                line[x] = qRgba(p[0], p[1], p[2], p[3]);
            }
        }
        QImage::registry()[name] = q;
        set->setDefaultImage(m->cams[v], ProjectImagePtr(new ProjectImage(QString(name))));
    }
    m->task.cancelled = false;  // Task leaves it uninitialised (gui/task.hpp:100); the GUI never cancels before run
    m->task.initialize(ProjectPtr(new Project), set, m->cams, minDepth, maxDepth, numDepthLevels, crossCheckThreshold, 1.0);
    m->task.imageScale = imageScale;
    return m;
}
void ref_mvs_destroy(ref_mvs *m) {
    for (size_t v = 0; v < m->cams.size(); ++v)
        QImage::registry().erase("mem://view" + std::to_string(v) + "@" + std::to_string((size_t)(void *)m));
    delete m;
}
int ref_mvs_num_views(const ref_mvs *m) { return (int)m->task.views.size(); }

// MultiViewStereo::runTask (multiviewstereo.cpp:325-475): neighbours, initial estimates, cross-check.
// depth_after: V * h*w doubles (the class's computedDepths); neighbours: V * 3 indices, -1 padded.
void ref_mvs_run(ref_mvs *m, double *depth_after, int32_t *neighbours) {
    m->task.runTask();
    const size_t V = m->task.views.size();
    for (size_t v = 0; v < V; ++v) {
        const std::vector<double> &d = m->task.computedDepths[v];
        std::memcpy(depth_after + v * d.size(), d.data(), d.size() * sizeof(double));
        for (int k = 0; k < 3; ++k)
            neighbours[3 * v + k] = k < (int)m->task.neighbours[v].size() ? (int32_t)m->task.neighbours[v][k] : -1;
    }
}
// MultiViewStereo::computeInitialEstimate (:524-662) of one view, after ref_mvs_run has selected the
// neighbours: the depths before the cross-check and the K = 9 (ncc, depth) peaks of every pixel.
void ref_mvs_initial_estimate(ref_mvs *m, int view, double *depth, double *peaks /* h*w*9*2 or null */) {
    m->task.computeInitialEstimate((size_t)view);
    const std::vector<double> &d = m->task.computedDepths[view];
    std::memcpy(depth, d.data(), d.size() * sizeof(double));
    if (peaks) {
        const int h = (int)CostFunction::peakPairs.size(), w = h ? (int)CostFunction::peakPairs[0].size() : 0;
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const std::vector<PeakPair> &pk = CostFunction::peakPairs[y][x];
                for (int k = 0; k < K; ++k) {
                    double *o = peaks + (((size_t)y * w + x) * K + k) * 2;
                    o[0] = k < (int)pk.size() ? pk[k].first : 0.0;
                    o[1] = k < (int)pk.size() ? pk[k].second : -1.0;
                }
            }
    }
}
// bench.py --impl reference: time computeInitialEstimate on a bounded sample without running the
// whole task.  The neighbour lists are handed over (runTask's rule, :335-360, is not separately
// callable; the caller's lists are checked equal to it in tests/test_oracle_vs_ref.py), and the
// reference view's mask is reduced to a row band — the class skips pixels outside its mask (:565).
void ref_mvs_set_neighbours(ref_mvs *m, int view, const int32_t *nbrs, int n) {
    m->task.neighbours[view].assign(nbrs, nbrs + n);
}
void ref_mvs_mask_rows(ref_mvs *m, int view, int row_begin, int row_end) {
    VectorImage &mask = m->task.masks[view];
    for (int y = 0; y < mask.height(); ++y)
        if (y < row_begin || y >= row_end)
            for (int x = 0; x < mask.width(); ++x) mask.setPixel(x, y, BLACK);
}
int ref_mvs_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// cost_ncc (:113-189) of view a's pixel (x1,y1) against view b's pixel (x2,y2), GeodesicWeight r = 2
double ref_mvs_cost_ncc(ref_mvs *m, int a, int b, int x1, int y1, int x2, int y2) {
    WeightFunc wf(WINDOW_RADIUS);
    wf.init_weights(m->task.images[a], x1, y1);
    return cost_ncc(m->task.images[a], m->task.images[b], m->task.masks[a], m->task.masks[b], x1, y1, x2, y2, wf);
}
// MultiViewStereo::epipolarCurve (:754-810) of view a's pixel towards view b: returns the point count
int ref_mvs_curve(ref_mvs *m, int a, int b, int x, int y, int32_t *out_xy, int max_pts) {
    const CameraPtr &view = m->task.views[a];
    Ray3d ray = view->unproject((x + 0.5) / m->task.imageScale, (y + 0.5) / m->task.imageScale);
    std::vector<Eigen::Vector3d> c = m->task.epipolarCurve(ray, view->C(), view->principleRay().direction(), m->task.masks[b], m->task.views[b]);
    for (size_t i = 0; i < c.size() && (int)i < max_pts; ++i) { out_xy[2 * i] = (int32_t)c[i][0]; out_xy[2 * i + 1] = (int32_t)c[i][1]; }
    return (int)c.size();
}
}
