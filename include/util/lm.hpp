// util/lm.hpp — LevenbergMarquardt with the reference's interface (util/lm.hpp:28-92) and loop
// (util/lm.cpp:59-150): the outer loop of the refractive-interface calibration, which SURVEY §8f
// rank 4 keeps on the host.  The residuals it consumes come from the GPU when the Function
// implements evaluateAll()/chiSquaredAll() (include/stereo/refractioncalibration.hpp does, through
// sr_calibration_residuals_batch); a Function that only has diff()/gradient() is driven point by
// point exactly as the reference drives it.
//
// Same as the reference: lambda starts at 1, the diagonal of H = J^T J is boosted by (1 + lambda),
// H d = -g is solved by LU with partial pivoting, a rejected step multiplies lambda by 10 and an
// accepted one by 0.1, and the loop ends after maxIterations or 5 consecutive iterations whose
// error changed by at most epsilon (or that were rejected outright).
//
// Two documented deviations from the reference text:
//  * fixed parameters are removed from the system.  The reference leaves their rows and columns of
//    H at zero (its own TODO, util/lm.cpp:22-25), the LU solve of that singular matrix yields NaN
//    and every step is rejected; the GUI always fixes the refractive index
//    (gui/widgets/stereowidget.cpp:577-578), so the reference never moves with its own caller.
//  * util/lm.cpp:104-107 skips the step when H*d IS approximately -g (the sense of the test was
//    lost when the Eigen-2 `solve` that returned a success flag was replaced), i.e. whenever the
//    solve succeeded.  The default here is the evident intent (skip when the solve FAILED);
//    setLiteralSolveCheck(true) restores the text as written, under which a well-conditioned
//    problem ends after 5 iterations with the model it started from.
//
// Point / Model are std::vector<double> (Eigen::VectorXd in the reference; size() and operator[]
// are what the callers use).
#ifndef SR_UTIL_LM_HPP
#define SR_UTIL_LM_HPP
#include <cmath>
#include <cstddef>
#include <utility>
#include <vector>

class LevenbergMarquardt {
public:
    typedef std::vector<double> Point;
    typedef std::pair<Point, Point> PointPair;
    typedef std::vector<PointPair> Points;
    typedef Point Model;
    typedef std::vector<bool> FixedParams;

    class Function {
    public:
        virtual ~Function() {}
        virtual void initialize() {}
        //! f(p.first; p.second, model)
        virtual Point diff(const PointPair &p, int point_index, const Model &model) = 0;
        //! d f / d model[paramIndex] at the given point
        virtual Point gradient(const PointPair &p, int point_index, const Model &model, int paramIndex) = 0;
        //! false: the model is outside the function's domain (the optimizer backs off)
        virtual bool update(const Model &) { return true; }

        // ---- bulk extensions (return false: not provided, the optimizer loops over the points) ----
        //! diffs[i] = diff(pts[i], i, model) and grads[r][i] = gradient(pts[i], i, model, r) for every
        //! free parameter r (grads[r] stays empty for fixed ones), in one go.  Must leave the function
        //! updated to `model`, as a sequence of gradient() calls does.
        virtual bool evaluateAll(const Points &, const Model &, const FixedParams &, std::vector<Point> &,
                                 std::vector<std::vector<Point> > &) { return false; }
        //! sum_i |diff(pts[i], i, model)|^2 for the model the function was last updated to
        virtual bool chiSquaredAll(const Points &, const Model &, double &) { return false; }
    };

    LevenbergMarquardt(int maxIterations = 1000, double epsilon = 1e-10)
        : maxIterations(maxIterations), epsilon(epsilon), literalSolveCheck_(false), iterations_(0) {}

    void setLiteralSolveCheck(bool on) { literalSolveCheck_ = on; }
    int iterations() const { return iterations_; }

    // util/lm.cpp:50-55
    static double chiSquared(Function &f, const Points &pts, const Model &model) {
        double sum = 0.0;
        if (f.chiSquaredAll(pts, model, sum)) return sum;
        for (size_t i = 0; i < pts.size(); ++i) {
            const Point d = f.diff(pts[i], (int)i, model);
            sum += dot(d, d);
        }
        return sum;
    }

    void optimize(Function &f, const Points &pts, Model &model) { optimize(f, pts, model, FixedParams(model.size(), false)); }

    // util/lm.cpp:64-150
    void optimize(Function &f, const Points &pts, Model &model, const FixedParams &fixed) {
        const int npts = (int)pts.size(), nparms = (int)model.size();
        iterations_ = 0;
        if (nparms <= 0 || npts <= 0) return;
        std::vector<int> free_;  // the reduced system's unknowns
        for (int p = 0; p < nparms; ++p)
            if (!fixed[p]) free_.push_back(p);
        const int nf = (int)free_.size();
        if (nf == 0) return;

        f.initialize();
        f.update(model);
        double e0 = chiSquared(f, pts, model);
        double lambda = 1;
        std::vector<double> H((size_t)nf * nf), g(nf), step(nf);
        std::vector<Point> diffs;
        std::vector<std::vector<Point> > grads;

        int iter = 0, term = 0;
        do {
            // H = J^T J and g = J^T f, accumulated point by point in the reference's order (:80-94)
            std::fill(H.begin(), H.end(), 0.0);
            std::fill(g.begin(), g.end(), 0.0);
            diffs.clear();
            grads.assign(nparms, std::vector<Point>());
            if (!f.evaluateAll(pts, model, fixed, diffs, grads)) {
                diffs.resize(npts);
                for (int r : free_) grads[r].resize(npts);
                for (int i = 0; i < npts; ++i) {
                    diffs[i] = f.diff(pts[i], i, model);
                    for (int r : free_) grads[r][i] = f.gradient(pts[i], i, model, r);
                }
            }
            for (int i = 0; i < npts; ++i)
                for (int a = 0; a < nf; ++a) {
                    const Point &gradr = grads[free_[a]][i];
                    for (int b = 0; b < nf; ++b) H[(size_t)a * nf + b] += dot(gradr, grads[free_[b]][i]);
                    g[a] += dot(diffs[i], gradr);
                }
            for (int a = 0; a < nf; ++a) H[(size_t)a * nf + a] *= 1.0 + lambda;  // towards gradient descent (:97-98)

            // Solve H d = -g (:101-107)
            std::vector<double> rhs(nf);
            for (int a = 0; a < nf; ++a) rhs[a] = -g[a];
            luSolve(H, rhs, step, nf);
            const bool solved = isApprox(H, step, rhs, nf, 1e-10);
            if (literalSolveCheck_ ? solved : !solved) {
                ++term;
                continue;
            }
            bool bad_model = false;  // (:110-119)
            for (int p = 0; p < nparms; ++p)
                if (std::isnan(model[p])) {
                    bad_model = true;
                    lambda *= 10.0;
                    ++term;
                }
            if (bad_model) continue;

            Model new_model = model;
            for (int a = 0; a < nf; ++a) new_model[free_[a]] += step[a];
            if (!f.update(new_model)) {  // (:123-128)
                f.update(model);
                lambda *= 10.0;
                ++term;
                continue;
            }
            const double e1 = chiSquared(f, pts, new_model);  // (:134-139)
            if (std::fabs(e1 - e0) > epsilon) term = 0;
            else ++term;
            const bool worse = (e0 - e1 < 0);  // (:142-150)
            if (worse || std::isnan(e1)) {
                lambda *= 10.0;
                f.update(model);
            } else {
                lambda *= 0.1;
                e0 = e1;
                model = new_model;
            }
        } while (++iter < maxIterations && term < 5);
        iterations_ = iter;
    }

private:
    static double dot(const Point &a, const Point &b) {
        double s = 0.0;
        for (size_t i = 0; i < a.size() && i < b.size(); ++i) s += a[i] * b[i];
        return s;
    }
    // Gaussian elimination with partial pivoting (what Eigen's PartialPivLU does; an exactly zero pivot
    // column is skipped as there, and the back substitution then divides by zero -> inf/NaN -> the
    // caller's residual test fails).
    static void luSolve(const std::vector<double> &A, const std::vector<double> &b, std::vector<double> &x, int n) {
        std::vector<double> M(A), y(b);
        for (int k = 0; k < n; ++k) {
            int piv = k;
            double big = std::fabs(M[(size_t)k * n + k]);
            for (int r = k + 1; r < n; ++r)
                if (std::fabs(M[(size_t)r * n + k]) > big) {
                    big = std::fabs(M[(size_t)r * n + k]);
                    piv = r;
                }
            if (big == 0.0) continue;
            if (piv != k) {
                for (int c = 0; c < n; ++c) std::swap(M[(size_t)k * n + c], M[(size_t)piv * n + c]);
                std::swap(y[k], y[piv]);
            }
            for (int r = k + 1; r < n; ++r) {
                const double l = M[(size_t)r * n + k] / M[(size_t)k * n + k];
                M[(size_t)r * n + k] = l;
                for (int c = k + 1; c < n; ++c) M[(size_t)r * n + c] -= l * M[(size_t)k * n + c];
                y[r] -= l * y[k];
            }
        }
        x.assign(n, 0.0);
        for (int k = n - 1; k >= 0; --k) {
            double s = y[k];
            for (int c = k + 1; c < n; ++c) s -= M[(size_t)k * n + c] * x[c];
            x[k] = s / M[(size_t)k * n + k];
        }
    }
    // Eigen's isApprox for vectors: |a - b|^2 <= prec^2 * min(|a|^2, |b|^2), a = H x
    static bool isApprox(const std::vector<double> &H, const std::vector<double> &x, const std::vector<double> &b, int n,
                         double prec) {
        double d2 = 0.0, a2 = 0.0, b2 = 0.0;
        for (int r = 0; r < n; ++r) {
            double a = 0.0;
            for (int c = 0; c < n; ++c) a += H[(size_t)r * n + c] * x[c];
            d2 += (a - b[r]) * (a - b[r]);
            a2 += a * a;
            b2 += b[r] * b[r];
        }
        return d2 <= prec * prec * (a2 < b2 ? a2 : b2);  // false for NaN
    }

    int maxIterations;
    double epsilon;
    bool literalSolveCheck_;
    int iterations_;
};
#endif
