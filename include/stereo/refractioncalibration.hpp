// stereo/refractioncalibration.hpp — RefractionCalibration with the reference's interface
// (stereo/refractioncalibration.hpp:40-101): calibrates the planar refractive interface of every
// view (normal + distance) and the refractive index by Levenberg-Marquardt on the ray/ray distance
// of feature correspondences (stereo/refractioncalibration.cpp:127-253,289-404).
//
// Every residual is computed on the GPU: one LM step evaluates the current model and both
// finite-difference perturbations of every free parameter for all correspondences with ONE
// sr_calibration_residuals_batch launch (the reference calls diff() 2 x points x parameters^2
// times per step).  The LM loop itself (include/util/lm.hpp) is host code, as SURVEY §8f rank 4
// prescribes.  No CPU fallback: construction of the session throws without a CUDA device.
//
// Model layout (RefractiveCalibrationFunction::update, :238-251): model[0] = refractive index,
// model[3v+1], model[3v+2] = the pixel whose back-projection K^-1 (px, py, 1) is view v's interface
// normal, model[3v+3] = its distance; size 1 + 3V.  (The reference's calibrate() insists on size 3V,
// :291, which leaves the last view's distance one element past the end of the vector; the caller
// in gui/widgets/stereowidget.cpp:573-592 writes that element too.  1 + 3V is the layout both of
// them index.)
//
// Qt-free input: the reference collects the correspondences from the project's FeatureDatabase
// (:339-361, out of scope); here they are handed over with setCorrespondences().
#ifndef SR_STEREO_REFRACTIONCALIBRATION_HPP
#define SR_STEREO_REFRACTIONCALIBRATION_HPP
#include "project/camera.hpp"
#include "stereo/sr_session.hpp"
#include "util/lm.hpp"
#include <limits>

typedef std::pair<int, int> IntPair;

//! Ray/ray distance metric of the interface calibration (stereo/refractioncalibration.cpp:127-253).
class RefractiveCalibrationFunction : public LevenbergMarquardt::Function {
public:
    typedef LevenbergMarquardt::Point Point;
    typedef LevenbergMarquardt::PointPair PointPair;
    typedef LevenbergMarquardt::Points Points;
    typedef LevenbergMarquardt::Model Model;
    typedef LevenbergMarquardt::FixedParams FixedParams;

    RefractiveCalibrationFunction(sr_host::Session &session, const std::vector<CameraPtr> &views,
                                  const std::vector<IntPair> &point2cams = std::vector<IntPair>())
        : s(session), views(views), point2cams(point2cams), exactAttribution_(true) {}
    ~RefractiveCalibrationFunction() {  // (:151-154)
        for (size_t v = 0; v < views.size() && v < original_P.size(); ++v) views[v]->setPlane(original_P[v]);
    }
    //! Which correspondences a parameter's finite difference is evaluated for.  Default: parameter
    //! p > 0 belongs to view (p - 1) / 3 and the refractive index to every point.  true: the
    //! reference's rule (:205-207), zero unless paramIndex / 3 is one of the point's two camera
    //! indices — with the layout update() indexes that ties view v's distance to view v+1, leaves the
    //! last view's distance without any gradient, and H singular: no step is ever accepted.
    void setReferenceAttribution(bool on) { exactAttribution_ = !on; }

    void initialize() {  // (:157-162)
        original_P.resize(views.size());
        for (size_t v = 0; v < views.size(); ++v) original_P[v] = views[v]->plane();
    }

    Point diff(const PointPair &pp, int point_index, const Model &model) {  // (:170-173)
        const IntPair &p2c = point2cams[point_index];
        return diff(pp, views[p2c.first], views[p2c.second], model);
    }
    //! the residual of one correspondence between two cameras in their CURRENT state (:175-201)
    Point diff(const PointPair &pp, CameraPtr view1, CameraPtr view2, const Model &) {
        const sr_camera cams[2] = {view1->toPod(), view2->toPod()};
        const int32_t pair[2] = {0, 1};
        const double pix[4] = {pp.first[0], pp.first[1], pp.second[0], pp.second[1]};
        Point out(1);
        s.check(sr_calibration_residuals(s.get(), 2, cams, 1, pair, pix, &out[0]), "sr_calibration_residuals");
        return out;
    }

    Point gradient(const PointPair &pp, int point_index, const Model &model, int paramIndex) {  // (:203-236)
        if (!touches(paramIndex, point2cams[point_index])) return Point(1, 0.0);
        Model m1, m2;
        perturb(model, paramIndex, m1, m2);
        update(m1);
        const Point val1 = diff(pp, point_index, m1);
        update(m2);
        const Point val2 = diff(pp, point_index, m2);
        update(model);
        return Point(1, (val2[0] - val1[0]) / (m2[paramIndex] - m1[paramIndex]));
    }

    bool update(const Model &model) {  // (:238-251)
        for (size_t v = 0; v < views.size(); ++v)
            if (model[3 * v + 2] < 1e-4) return false;
        for (size_t v = 0; v < views.size(); ++v) {
            const CameraPtr &view = views[v];
            Eigen::Vector3d normal = view->Kinv() * Eigen::Vector3d(model[3 * v + 1], model[3 * v + 2], 1);
            Plane3d P;
            P.setDistance(model[3 * v + 3]);
            P.setNormal(normal);  // normalises
            if (std::fabs(model[0] - view->refractiveIndex()) > 1e-10) view->setRefractiveIndex(model[0]);  // camera.cpp:337
            if (view->plane() != P) view->setPlane(P);                                                      // camera.cpp:327
        }
        return true;
    }

    // ---- bulk evaluation: one launch per LM step / per error evaluation ---------------------
    bool evaluateAll(const Points &pts, const Model &model, const FixedParams &fixed, std::vector<Point> &diffs,
                     std::vector<std::vector<Point> > &grads) {
        const int n = (int)pts.size(), nparms = (int)model.size(), V = (int)views.size();
        std::vector<int> free_;
        for (int p = 0; p < nparms; ++p)
            if (!fixed[p]) free_.push_back(p);
        const int M = 1 + 2 * (int)free_.size();
        std::vector<sr_camera> cams((size_t)M * V);
        std::vector<double> h(free_.size());
        update(model);
        snapshot(&cams[0]);
        for (size_t a = 0; a < free_.size(); ++a) {
            Model m1, m2;
            perturb(model, free_[a], m1, m2);
            h[a] = m2[free_[a]] - m1[free_[a]];
            update(m1);  // a rejected model leaves the cameras at `model`, as in the reference's sequence
            snapshot(&cams[(size_t)(1 + 2 * a) * V]);
            update(model);
            update(m2);
            snapshot(&cams[(size_t)(2 + 2 * a) * V]);
            update(model);
        }
        std::vector<double> res((size_t)M * n);
        evaluate(M, cams, pts, res);
        diffs.assign(n, Point(1, 0.0));
        for (int i = 0; i < n; ++i) diffs[i][0] = res[i];
        grads.assign(nparms, std::vector<Point>());
        for (size_t a = 0; a < free_.size(); ++a) {
            std::vector<Point> &gr = grads[free_[a]];
            gr.assign(n, Point(1, 0.0));
            const double *v1 = &res[(size_t)(1 + 2 * a) * n], *v2 = &res[(size_t)(2 + 2 * a) * n];
            for (int i = 0; i < n; ++i)
                if (touches(free_[a], point2cams[i])) gr[i][0] = (v2[i] - v1[i]) / h[a];
        }
        return true;
    }
    bool chiSquaredAll(const Points &pts, const Model &, double &sum) {
        const int n = (int)pts.size(), V = (int)views.size();
        std::vector<sr_camera> cams(V);
        snapshot(&cams[0]);
        std::vector<double> res(n);
        evaluate(1, cams, pts, res);
        sum = 0.0;
        for (int i = 0; i < n; ++i) sum += res[i] * res[i];
        return true;
    }

private:
    bool touches(int paramIndex, const IntPair &p2c) const {
        if (exactAttribution_) {
            if (paramIndex == 0) return true;
            const int v = (paramIndex - 1) / 3;
            return v == p2c.first || v == p2c.second;
        }
        return paramIndex / 3 == p2c.first || paramIndex / 3 == p2c.second;  // (:205-207)
    }
    // finite-difference stencils (:210-226)
    static void perturb(const Model &model, int paramIndex, Model &m1, Model &m2) {
        m1 = model;
        m2 = model;
        if (paramIndex == 0) {  // refractive index
            m1[paramIndex] = model[paramIndex] - 0.01;
            m2[paramIndex] = model[paramIndex] + 0.01;
        } else if ((paramIndex - 1) % 3 == 0) {  // pixel x of the interface normal
            m1[paramIndex] = model[paramIndex] - 0.5;
            m2[paramIndex] = model[paramIndex] + 0.5;
        } else if ((paramIndex - 1) % 3 == 1) {  // pixel y
            m1[paramIndex] = model[paramIndex] - 0.1;
            m2[paramIndex] = model[paramIndex] + 0.1;
        } else {  // distance: one-sided
            m1[paramIndex] = model[paramIndex];
            m2[paramIndex] = model[paramIndex] + 0.0001;
        }
    }
    void snapshot(sr_camera *out) const {
        for (size_t v = 0; v < views.size(); ++v) out[v] = views[v]->toPod();
    }
    void evaluate(int M, const std::vector<sr_camera> &cams, const Points &pts, std::vector<double> &res) {
        const int n = (int)pts.size();
        std::vector<int32_t> pairs((size_t)2 * n);
        std::vector<double> pix((size_t)4 * n);
        for (int i = 0; i < n; ++i) {
            pairs[2 * i] = point2cams[i].first;
            pairs[2 * i + 1] = point2cams[i].second;
            pix[4 * i] = pts[i].first[0];
            pix[4 * i + 1] = pts[i].first[1];
            pix[4 * i + 2] = pts[i].second[0];
            pix[4 * i + 3] = pts[i].second[1];
        }
        s.check(sr_calibration_residuals_batch(s.get(), M, (int)views.size(), cams.data(), n, pairs.data(), pix.data(), res.data()),
                "sr_calibration_residuals_batch");
    }

    sr_host::Session &s;
    std::vector<CameraPtr> views;
    std::vector<IntPair> point2cams;
    std::vector<Plane3d> original_P;
    bool exactAttribution_;
};

class RefractionCalibration {
public:
    typedef LevenbergMarquardt::Model Model;
    typedef LevenbergMarquardt::FixedParams FixedParams;
    typedef LevenbergMarquardt::Points Points;

    RefractionCalibration() : model_(1, 0.0), fixed(1, false), device_(0), referenceAttribution_(false), literalSolveCheck_(false), iterations_(0) {}

    void setModel(const Model &model) { setModel(model, FixedParams(model.size(), false)); }
    void setModel(const Model &model, const FixedParams &fixed) {
        this->model_ = model;
        this->fixed = fixed;
    }
    Model model() const { return model_; }
    void setViews(const std::vector<CameraPtr> &views) { this->views = views; }
    //! Extension replacing setProject()/setImageSets(): the correspondences themselves — pixel pairs
    //! (first in view point2cams[i].first, second in view point2cams[i].second).
    void setCorrespondences(const Points &points, const std::vector<IntPair> &point2cams) {
        this->points = points;
        this->point2cams = point2cams;
    }
    void setDevice(int device) { device_ = device; }
    void setReferenceAttribution(bool on) { referenceAttribution_ = on; }  // see RefractiveCalibrationFunction
    void setLiteralSolveCheck(bool on) { literalSolveCheck_ = on; }  // see util/lm.hpp
    int iterations() const { return iterations_; }

    // stereo/refractioncalibration.cpp:289-404 (the non-bundle-adjustment branch, the one compiled)
    bool calibrate() {
        if (views.size() < 2 || points.empty() || points.size() != point2cams.size() || model_.size() != 1 + 3 * views.size() ||
            fixed.size() != model_.size())
            return false;
        sr_host::Session s(device_);
        RefractiveCalibrationFunction func(s, views, point2cams);
        func.setReferenceAttribution(referenceAttribution_);
        func.initialize();
        func.update(model_);
        initialError_ = LevenbergMarquardt::chiSquared(func, points, model_);
        LevenbergMarquardt lm(100, 1);  // (:372)
        lm.setLiteralSolveCheck(literalSolveCheck_);
        lm.optimize(func, points, model_, fixed);
        iterations_ = lm.iterations();
        func.update(model_);
        finalError_ = LevenbergMarquardt::chiSquared(func, points, model_);
        for (size_t i = 0; i < model_.size(); ++i)  // NaN = bad optimisation (:399-401)
            if (std::isnan(model_[i])) return false;
        return true;
    }
    //! chi^2 before / after the last calibrate() (the reference prints them with qDebug, :366-378)
    double initialError() const { return initialError_; }
    double finalError() const { return finalError_; }

    // (:408-447) sum of squared residuals of all correspondences under the current model
    double totalError(double *average = nullptr) const {
        if (views.size() < 2 || model_.size() != 1 + 3 * views.size()) return std::numeric_limits<double>::quiet_NaN();
        sr_host::Session s(device_);
        RefractiveCalibrationFunction func(s, views, point2cams);
        func.initialize();
        func.update(model_);
        double total = 0.0;
        func.chiSquaredAll(points, model_, total);
        if (average) *average = total / (double)points.size();
        return total;
    }
    // (:451-465)
    double error(const LevenbergMarquardt::PointPair &c, CameraPtr view1, CameraPtr view2) const {
        if (!view1 || !view2) return std::numeric_limits<double>::quiet_NaN();
        sr_host::Session s(device_);
        RefractiveCalibrationFunction func(s, views, point2cams);
        func.initialize();
        func.update(model_);
        return std::fabs(func.diff(c, view1, view2, model_)[0]);
    }

private:
    std::vector<CameraPtr> views;
    Points points;
    std::vector<IntPair> point2cams;
    Model model_;
    FixedParams fixed;
    int device_;
    bool referenceAttribution_, literalSolveCheck_;
    int iterations_;
    double initialError_ = std::numeric_limits<double>::quiet_NaN(), finalError_ = std::numeric_limits<double>::quiet_NaN();
};
#endif
