// stereo/adaptiveweight.hpp — AdaptiveWeight with the reference's interface
// (stereo/adaptiveweight.hpp:28-47, stereo/adaptiveweight.cpp:33-79).  init_weights() computes the window on the GPU through
// sr_compute_weights; the image is uploaded once per VectorImage (identity camera) and cached.
#ifndef SR_STEREO_ADAPTIVEWEIGHT_HPP
#define SR_STEREO_ADAPTIVEWEIGHT_HPP
#include "stereo/sr_session.hpp"
#include "util/vectorimage.hpp"
class AdaptiveWeight : public std::binary_function<int, int, double> {
public:
    AdaptiveWeight() : radius(0), bound_(nullptr), bw_(0), bh_(0) {}
    AdaptiveWeight(int radius) : radius(0), bound_(nullptr), bw_(0), bh_(0) { initialize(radius); }
    void initialize(int r) {
        radius = r;
        weights.assign((size_t)(2 * r + 1) * (2 * r + 1), 0.0);
    }
    void init_weights(const VectorImage &img, int x, int y) {
        bind(img);
        const int32_t cx = x, cy = y;
        session_->check(sr_compute_weights(session_->get(), 0, SR_WEIGHT_ADAPTIVE, radius, 1, &cx, &cy, weights.data()), "sr_compute_weights");
    }
    //! Batched form (extension): n window centres in one launch, out = n*(2r+1)^2 doubles.
    void init_weights_batch(const VectorImage &img, int n, const int32_t *xs, const int32_t *ys, double *out) {
        bind(img);
        session_->check(sr_compute_weights(session_->get(), 0, SR_WEIGHT_ADAPTIVE, radius, n, xs, ys, out), "sr_compute_weights");
    }
    double operator()(int row, int col) const { return weights[(size_t)(row + radius) * (2 * radius + 1) + (col + radius)]; }
private:
    void bind(const VectorImage &img) {
        if (!session_) session_.reset(new sr_host::Session(0));
        if (bound_ == &img && bw_ == img.width() && bh_ == img.height()) return;
        const std::vector<uint8_t> rgba = img.toRGBA8();
        sr_camera cam;
        std::memset(&cam, 0, sizeof(cam));
        cam.K[0] = cam.K[4] = cam.K[8] = cam.Kinv[0] = cam.Kinv[4] = cam.Kinv[8] = 1;
        cam.R[0] = cam.R[4] = cam.R[8] = cam.Rinv[0] = cam.Rinv[4] = cam.Rinv[8] = 1;
        cam.plane_n[2] = cam.prin_dir[2] = 1;
        cam.n = 1;
        const uint8_t *ip = rgba.data();
        session_->check(sr_set_views(session_->get(), 1, &cam, &ip, nullptr, img.width(), img.height()), "sr_set_views");
        session_->check(sr_synchronize(session_->get()), "sr_synchronize");
        bound_ = &img;
        bw_ = img.width();
        bh_ = img.height();
    }
    int radius;
    std::vector<double> weights;
    std::shared_ptr<sr_host::Session> session_;
    const VectorImage *bound_;
    int bw_, bh_;
};
#endif
