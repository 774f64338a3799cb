// stereo/twoviewstereo.hpp — TwoViewStereo with the reference's public interface
// (stereo/twoviewstereo.hpp:37-70) on top of the B200 C ABI (include/sr_b200.h).
//
// The class owns copies of its inputs (as the reference does, twoviewstereo.cpp:97-113), snapshots
// the two cameras into sr_camera PODs at construction, and computes ONLY through the C ABI:
//   computeDepthMaps()  = label-mode cost volumes + WTA in both directions (sr_run_view, the
//                         compiled-out label sweep twoviewstereo.cpp:308-329,477-497 with the live
//                         WTA margins :293-305), crossCheck (:596-672, sr_cross_check), then the
//                         HSV colourisation (:128-146) of the device-computed depths.
// Reference file-scope constants (twoviewstereo.cpp:64-80) are the defaults of params(); a caller
// may change radius / weight / cost before run() — an extension, the reference fixes them by typedef.
// Extension accessors leftDepths()/rightDepths()/leftIndices()/rightIndices() expose the raw maps
// (the reference keeps them private; parity tests need them).
#ifndef SR_STEREO_TWOVIEWSTEREO_HPP
#define SR_STEREO_TWOVIEWSTEREO_HPP
#include "gui/task.hpp"
#include "project/camera.hpp"
#include "stereo/sr_session.hpp"
#include "util/plane.hpp"
#include "util/ray.hpp"
#include "util/vectorimage.hpp"

class TwoViewStereo : public Task {
public:
    typedef std::vector<double> DepthMap;

    TwoViewStereo(CameraPtr leftView, QImage left, QImage leftMask, CameraPtr rightView, QImage right, QImage rightMask,
                  double minDepth, double maxDepth, int numDepthLevels, double imageScale = 1.0)
        : leftView(leftView), rightView(rightView),
          left(VectorImage::fromQImage(left.scaledToWidth((int)(left.width() * imageScale)))),
          right(VectorImage::fromQImage(right.scaledToWidth((int)(right.width() * imageScale)))),
          minDepth(minDepth), maxDepth(maxDepth), numDepthLevels(numDepthLevels), imageScale(imageScale),
          crossCheckThreshold_(1.0) /* INCONSISTENCY_THRESH, twoviewstereo.cpp:80 */ {
        // a null mask means "every pixel" (twoviewstereo.cpp:105-113)
        this->leftMask = leftMask.isNull() ? VectorImage(this->left.width(), this->left.height(), WHITE)
                                           : VectorImage::fromQImage(leftMask.scaledToWidth((int)(leftMask.width() * imageScale)));
        this->rightMask = rightMask.isNull() ? VectorImage(this->right.width(), this->right.height(), WHITE)
                                             : VectorImage::fromQImage(rightMask.scaledToWidth((int)(rightMask.width() * imageScale)));
        resultLeft = VectorImage(this->left.width(), this->left.height());
        resultRight = VectorImage(this->right.width(), this->right.height());
        const double nan = std::numeric_limits<double>::quiet_NaN();
        computedDepthLeft.assign((size_t)this->left.width() * this->left.height(), nan);
        computedDepthRight.assign((size_t)this->left.width() * this->left.height(), nan);
        sr_params_default(&params_, 0);
        params_.min_depth = minDepth;
        params_.max_depth = maxDepth;
        params_.num_levels = numDepthLevels;
        params_.image_scale = imageScale;
    }

    // Task implementation
    std::string title() const { return "Two-View Stereo"; }
    int numSteps() const { return 8; }
    void runTask() { computeDepthMaps(); }

    void computeDepthMaps() {
        resultLeft.fill(INVALID);
        resultRight.fill(INVALID);
        if (!leftView || !rightView || left.isNull() || right.isNull()) return;
        if (left.width() != right.width() || left.height() != right.height())
            throw std::runtime_error("TwoViewStereo: the B200 path needs equally sized views");
        sr_host::Session s(device_);
        const int w = left.width(), h = left.height();
        const sr_camera cams[2] = {leftView->toPod(), rightView->toPod()};
        const std::vector<uint8_t> il = left.toRGBA8(), ir = right.toRGBA8(), ml = leftMask.toMask8(), mr = rightMask.toMask8();
        const uint8_t *imgs[2] = {il.data(), ir.data()}, *msks[2] = {ml.data(), mr.data()};
        s.check(sr_set_views(s.get(), 2, cams, imgs, msks, w, h), "sr_set_views");
        s.check(sr_set_params(s.get(), &params_), "sr_set_params");
        const int32_t one = 1, zero = 0;
        stageUpdate("Computing cost volume (left to right)");
        progressUpdate(0);
        if (isCancelled()) return;
        s.check(curveMode_ ? sr_run_view_curve(s.get(), 0, &one, 1) : sr_run_view(s.get(), 0, &one, 1), "sr_run_view");
        stageUpdate("Computing cost volume (right to left)");
        progressUpdate(3);
        if (isCancelled()) return;
        s.check(curveMode_ ? sr_run_view_curve(s.get(), 1, &zero, 1) : sr_run_view(s.get(), 1, &zero, 1), "sr_run_view");
        fetch(s, w, h);  // the reference colourises once before cross-checking (:160-180)
        if (isCancelled()) return;
        stageUpdate("Cross-checking");
        progressUpdate(6);
        s.check(sr_cross_check(s.get(), 1, crossCheckThreshold_), "sr_cross_check");
        fetch(s, w, h);
        progressUpdate(8);
        stageUpdate("Finished!");
    }

    QImage leftDepthMap() const { return VectorImage::toQImage(resultLeft); }
    QImage rightDepthMap() const { return VectorImage::toQImage(resultRight); }

    //! The refractive epipolar curve of `ray` in `view` as integer pixels (twoviewstereo.cpp:999-1054).
    //! One curve is interactive-preview work (gui/widgets/stereowidget.cpp:621-672), evaluated
    //! with the host Camera model; the per-pixel dense search runs on the GPU.
    std::vector<Eigen::Vector3d> epipolarCurve(const Ray3d &ray, const Eigen::Vector3d &cameraOffset,
                                               const Eigen::Vector3d &depthPlaneNormal, const VectorImage &mask,
                                               CameraPtr view) const {
        std::vector<Eigen::Vector3d> pts;
        bool have = false;
        Eigen::Vector3d last;
        for (int d = 0; d < numDepthLevels; ++d) {
            Ray3d::Point p;
            const double depth = depthFromLabel(d);
            if (!intersect(ray, Plane3d(depthPlaneNormal, cameraOffset + depthPlaneNormal * depth), p)) continue;
            if (!view->project(p)) continue;
            p *= imageScale;
            if (!have) {
                have = true;
                last = p;
                continue;
            }
            const double dx = p[0] - last[0], dy = p[1] - last[1];
            if (!(dx * dx + dy * dy >= 1)) continue;
            appendSegment(pts, last, p, mask);
            last = p;
        }
        return pts;
    }

    // ---- extensions ------------------------------------------------------------------------
    sr_params &params() { return params_; }
    void setDevice(int device) { device_ = device; }
    //! true (default): the reference's live search along the rasterised epipolar curve (twoviewstereo.cpp:285-305);
    //! false: the depth-label cost volume + WTA (the compiled-out branch, :308-329).
    void setCurveMode(bool on) { curveMode_ = on; }
    void setCrossCheckThreshold(double t) { crossCheckThreshold_ = t; }
    const DepthMap &leftDepths() const { return computedDepthLeft; }
    const DepthMap &rightDepths() const { return computedDepthRight; }
    const std::vector<int32_t> &leftIndices() const { return indexLeft; }
    const std::vector<int32_t> &rightIndices() const { return indexRight; }

protected:
    double depthFromLabel(int label) const {  // twoviewstereo.cpp:981-985
        double t = label / (numDepthLevels - 1.0);
        t /= (5 - 4 * t);
        return minDepth * (1 - t) + maxDepth * t;
    }

    RGBA colorFromDepth(double depth) const {  // twoviewstereo.cpp:128-146
        if (std::isnan(depth) || std::isinf(depth)) return BLACK;
        const double t = (depth - minDepth) / (maxDepth - minDepth);
        if (t < 1e-5) return BLACK;
        if (t > 1.1) return WHITE;
        // QColor::fromHsvF(2t/3, 1, 1): hue in [0,1) scaled to 16-bit, s = v = 1, then to 8-bit RGB
        double hh = 2.0 * t / 3.0;
        hh -= std::floor(hh);
        const double h6 = hh * 6.0;
        const int sector = (int)h6;
        const double f = h6 - sector;
        double r = 0, g = 0, b = 0;
        switch (sector % 6) {
            case 0: r = 1; g = f; b = 0; break;
            case 1: r = 1 - f; g = 1; b = 0; break;
            case 2: r = 0; g = 1; b = f; break;
            case 3: r = 0; g = 1 - f; b = 1; break;
            case 4: r = f; g = 0; b = 1; break;
            default: r = 1; g = 0; b = 1 - f; break;
        }
        return RGBA(std::floor(r * 255 + 0.5), std::floor(g * 255 + 0.5), std::floor(b * 255 + 0.5));
    }

private:
    void fetch(sr_host::Session &s, int w, int h) {
        indexLeft.resize((size_t)w * h);
        indexRight.resize((size_t)w * h);
        s.check(sr_get_depth(s.get(), 0, computedDepthLeft.data()), "sr_get_depth");
        s.check(sr_get_depth(s.get(), 1, computedDepthRight.data()), "sr_get_depth");
        s.check(sr_get_depth_index(s.get(), 0, indexLeft.data()), "sr_get_depth_index");
        s.check(sr_get_depth_index(s.get(), 1, indexRight.data()), "sr_get_depth_index");
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                resultLeft.setPixel(x, y, colorFromDepth(computedDepthLeft[(size_t)y * w + x]));
                resultRight.setPixel(x, y, colorFromDepth(computedDepthRight[(size_t)y * w + x]));
            }
    }

    // Bresenham between the truncated end points, emitted low-x first, WHITE-mask pixels only
    // (util/lineiter.hpp:32-118 semantics; no clipping in the two-view class, twoviewstereo.cpp:1028)
    static void appendSegment(std::vector<Eigen::Vector3d> &out, const Eigen::Vector3d &a, const Eigen::Vector3d &b,
                              const VectorImage &mask) {
        int x0 = (int)a[0], y0 = (int)a[1], x1 = (int)b[0], y1 = (int)b[1];
        const bool steep = std::abs(y1 - y0) > std::abs(x1 - x0);
        if (steep) { std::swap(x0, y0); std::swap(x1, y1); }
        if (x0 > x1) { std::swap(x0, x1); std::swap(y0, y1); }
        const int dx = x1 - x0, dy = std::abs(y1 - y0), ystep = (y0 < y1) ? 1 : -1;
        int err = dx / 2, y = y0;
        for (int x = x0; x <= x1; ++x) {
            const int px = steep ? y : x, py = steep ? x : y;
            if (mask.isNull() || mask.pixel(px, py) == WHITE) out.push_back(Eigen::Vector3d(px, py, 1));
            err -= dy;
            if (err < 0) { y += ystep; err += dx; }
        }
    }

    CameraPtr leftView, rightView;
    VectorImage left, right, leftMask, rightMask, resultLeft, resultRight;
    DepthMap computedDepthLeft, computedDepthRight;
    std::vector<int32_t> indexLeft, indexRight;
    double minDepth, maxDepth;
    int numDepthLevels;
    double imageScale;
    double crossCheckThreshold_;
    sr_params params_;
    int device_ = 0;
    bool curveMode_ = true;   // the reference's live formulation; setCurveMode(false) = disparity/depth-label cost volume
};
#endif
