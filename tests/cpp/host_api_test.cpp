// host_api_test — drives the stereo/ class API (include/stereo, include/project, include/util)
// the way gui/widgets/stereowidget.cpp:974-1002 does, headless: loads a project XML + image set,
// runs MultiViewStereo over all cameras and TwoViewStereo over the first two, and dumps raw
// results for tests/test_host_api.py to compare against the C ABI called directly and the oracle.
//
//   host_api_test <project.xml> <imageSetId> <outdir> <minDepth> <maxDepth> <levels> <crossCheck>
//   host_api_test --cameras <project.xml> <outdir>
//       no GPU needed: loads the project and dumps every camera's sr_camera POD (Camera::setP ->
//       updateOthers, lens distortion, interface) to <outdir>/cams.bin, ordered by camera id
//   host_api_test --calibrate <project.xml> <problem.bin> <outdir>
//       RefractionCalibration over the project's cameras; problem.bin = int32 n, int32 flags
//       (1: reference gradient attribution, 2: literal solve check), int32 pairs[2n],
//       double pixels[4n], double model[1+3V], uint8 fixed[1+3V]
#include <cstdio>
#include <fstream>
#include <iostream>

#include "project/project.hpp"
#include "stereo/adaptiveweight.hpp"
#include "stereo/geodesicweight.hpp"
#include "stereo/multiviewstereo.hpp"
#include "stereo/refractioncalibration.hpp"
#include "stereo/twoviewstereo.hpp"

template <typename T>
static void dump(const std::string &path, const T *data, size_t n) {
    std::ofstream f(path.c_str(), std::ios::binary);
    f.write(reinterpret_cast<const char *>(data), (std::streamsize)(n * sizeof(T)));
}

static int calibrate_main(const char *xml, const char *problem, const std::string &out) {
    ProjectPtr project(new Project(xml));
    std::vector<CameraPtr> views;
    for (const auto &kv : project->cameras()) views.push_back(kv.second);
    const size_t V = views.size(), np_ = 1 + 3 * V;
    std::ifstream f(problem, std::ios::binary);
    int32_t n = 0, flags = 0;
    f.read(reinterpret_cast<char *>(&n), 4);
    f.read(reinterpret_cast<char *>(&flags), 4);
    std::vector<int32_t> pairs((size_t)2 * n);
    std::vector<double> pix((size_t)4 * n), model(np_);
    std::vector<uint8_t> fx(np_);
    f.read(reinterpret_cast<char *>(pairs.data()), (std::streamsize)(pairs.size() * 4));
    f.read(reinterpret_cast<char *>(pix.data()), (std::streamsize)(pix.size() * 8));
    f.read(reinterpret_cast<char *>(model.data()), (std::streamsize)(np_ * 8));
    f.read(reinterpret_cast<char *>(fx.data()), (std::streamsize)np_);
    if (!f) throw std::runtime_error("short calibration problem file");
    LevenbergMarquardt::Points points;
    std::vector<IntPair> p2c;
    for (int i = 0; i < n; ++i) {
        LevenbergMarquardt::Point a(2), b(2);
        a[0] = pix[4 * i]; a[1] = pix[4 * i + 1]; b[0] = pix[4 * i + 2]; b[1] = pix[4 * i + 3];
        points.push_back(LevenbergMarquardt::PointPair(a, b));
        p2c.push_back(IntPair(pairs[2 * i], pairs[2 * i + 1]));
    }
    RefractionCalibration calib;
    calib.setViews(views);
    calib.setCorrespondences(points, p2c);
    calib.setModel(model, LevenbergMarquardt::FixedParams(fx.begin(), fx.end()));
    calib.setReferenceAttribution((flags & 1) != 0);
    calib.setLiteralSolveCheck((flags & 2) != 0);
    const Plane3d before = views[0]->plane();
    const bool ok = calib.calibrate();
    double avg = 0.0;
    const double total = calib.totalError(&avg);
    std::vector<double> res = calib.model();
    res.push_back((double)calib.iterations());
    res.push_back(calib.initialError());
    res.push_back(calib.finalError());
    res.push_back(total);
    res.push_back(avg);
    res.push_back(calib.error(points[0], views[p2c[0].first], views[p2c[0].second]));
    res.push_back(ok ? 1.0 : 0.0);
    // the function's destructor puts back the planes it saw at its LAST initialize() (:151-162) — LM calls
    // initialize() again after the caller's update(model), so that is the start model's plane, as in the reference
    res.push_back(views[0]->plane() == before ? 1.0 : 0.0);
    dump(out + "/calib.bin", res.data(), res.size());
    std::printf("calibrate: %d correspondences, %zu views, %d iterations, chi2 %.6g -> %.6g, ok=%d\n", n, V, calib.iterations(),
                calib.initialError(), calib.finalError(), (int)ok);
    return 0;
}

int main(int argc, char **argv) {
    if (argc == 4 && std::string(argv[1]) == "--cameras") {
        try {
            ProjectPtr project(new Project(argv[2]));
            std::vector<sr_camera> pods;
            for (const auto &kv : project->cameras()) pods.push_back(kv.second->toPod());
            dump(std::string(argv[3]) + "/cams.bin", pods.data(), pods.size());
            std::printf("cameras: %zu\n", pods.size());
            return 0;
        } catch (const std::exception &e) {
            std::fprintf(stderr, "host_api_test: %s\n", e.what());
            return 1;
        }
    }
    if (argc == 5 && std::string(argv[1]) == "--calibrate") {
        try {
            return calibrate_main(argv[2], argv[3], argv[4]);
        } catch (const std::exception &e) {
            std::fprintf(stderr, "host_api_test: %s\n", e.what());
            return 1;
        }
    }
    if (argc < 8) {
        std::fprintf(stderr, "usage: %s project.xml imageSetId outdir minDepth maxDepth levels crossCheck\n", argv[0]);
        return 2;
    }
    const std::string out = argv[3];
    const double minDepth = std::atof(argv[4]), maxDepth = std::atof(argv[5]), crossCheck = std::atof(argv[7]);
    const int levels = std::atoi(argv[6]);
    try {
        ProjectPtr project(new Project(argv[1]));
        ImageSetPtr set = project->imageSet(argv[2]);
        if (!set) throw std::runtime_error("image set not found");
        std::vector<CameraPtr> views;
        for (const auto &kv : project->cameras()) views.push_back(kv.second);  // stereowidget.cpp:985-987
        std::printf("project: %zu cameras, image set '%s'\n", views.size(), set->name().c_str());

        // ---- MultiViewStereo, as the GUI runs it
        MultiViewStereo mvs;
        int lastProgress = -1;
        std::string lastStage;
        mvs.onProgressUpdate = [&](int v) { lastProgress = v; };
        mvs.onStageUpdate = [&](const std::string &s) { lastStage = s; };
        mvs.initialize(project, set, views, minDepth, maxDepth, levels, crossCheck, 1.0);
        mvs.setCurveMode(false);  // first the depth-label volume (the class default is the reference's curve search)
        std::printf("mvs: %zu views loaded, numSteps=%d, title=%s\n", mvs.numViews(), mvs.numSteps(), mvs.title().c_str());
        mvs.run();
        std::vector<sr_camera> pods;
        for (const CameraPtr &c : views) pods.push_back(c->toPod());
        dump(out + "/cams.bin", pods.data(), pods.size());
        for (size_t v = 0; v < mvs.numViews(); ++v) {
            const std::string tag = out + "/mvs_v" + std::to_string(v);
            dump(tag + "_depth.bin", mvs.depths(v).data(), mvs.depths(v).size());
            dump(tag + "_index.bin", mvs.indices(v).data(), mvs.indices(v).size());
            QImage img = mvs.depthMap(views[v]);
            dump(tag + "_image.bin", img.bits(), (size_t)img.width() * img.height() * 4);
            std::vector<int32_t> nb(mvs.selectedNeighbours()[v].begin(), mvs.selectedNeighbours()[v].end());
            dump(tag + "_nbrs.bin", nb.data(), nb.size());
            std::printf("  view %zu (%s): %.1f%% of in-mask pixels have depth before, %.1f%% after cross-check\n", v,
                        views[v]->name().c_str(), 100 * mvs.coverageBeforeCrossCheck()[v], 100 * mvs.coverageAfterCrossCheck()[v]);
        }
        std::printf("mvs: last stage '%s', last progress %d, unknown view -> null image: %d\n", lastStage.c_str(), lastProgress,
                    (int)mvs.depthMap(CameraPtr(new Camera("nope", "nope"))).isNull());
        {   // the reference's live formulation: curve-mode search, same task object re-run
            mvs.setCurveMode(true);
            mvs.run();
            for (size_t v = 0; v < mvs.numViews(); ++v)
                dump(out + "/mvs_curve_v" + std::to_string(v) + "_depth.bin", mvs.depths(v).data(), mvs.depths(v).size());
            // ... and once more keeping the K = 9 peak lists (CostFunction::peakPairs)
            mvs.setKeepPeaks(true);
            mvs.run();
            dump(out + "/mvs_curve_v0_peaks.bin", mvs.peakPairs(0).data(), mvs.peakPairs(0).size());
            mvs.setKeepPeaks(false);
            std::printf("mvs (curve mode): %.1f%% of view 0 has depth after cross-check\n", 100 * mvs.coverageAfterCrossCheck()[0]);
            mvs.setCurveMode(false);
            mvs.run();
        }
        const std::vector<PLYPoint> cloud = mvs.pointCloud();
        outputPLYFile(out + "/cloud.ply", cloud);
        std::printf("mvs: point cloud of %zu points written\n", cloud.size());

        // ---- TwoViewStereo on the first two views (library API only in the reference)
        if (views.size() >= 2) {
            QImage l(set->defaultImageForCamera(views[0])->file()), r(set->defaultImageForCamera(views[1])->file());
            TwoViewStereo two(views[0], l, QImage(), views[1], r, QImage(), minDepth, maxDepth, levels, 1.0);
            two.params().radius = 2;               // keep the harness quick; the reference constant is 5
            two.setCrossCheckThreshold(30.0);
            two.setCurveMode(false);  // the label volume (compared with the oracle's label mode)
            two.run();
            dump(out + "/two_left_depth.bin", two.leftDepths().data(), two.leftDepths().size());
            dump(out + "/two_right_depth.bin", two.rightDepths().data(), two.rightDepths().size());
            dump(out + "/two_left_index.bin", two.leftIndices().data(), two.leftIndices().size());
            QImage li = two.leftDepthMap();
            dump(out + "/two_left_image.bin", li.bits(), (size_t)li.width() * li.height() * 4);
            // the public curve preview
            const Ray3d ray = views[0]->unproject((l.width() / 2 + 0.5), (l.height() / 2 + 0.5));
            const auto curve = two.epipolarCurve(ray, views[0]->C(), views[0]->principleRay().direction(), VectorImage(), views[1]);
            std::vector<int32_t> cp;
            for (const auto &p : curve) { cp.push_back((int32_t)p[0]); cp.push_back((int32_t)p[1]); }
            dump(out + "/two_curve.bin", cp.data(), cp.size());
            std::printf("two-view: done, centre-pixel epipolar curve has %zu pixels\n", curve.size());
        }

        // ---- weight functors, reference call shape: init_weights(img, x, y) then operator()(row, col)
        {
            VectorImage img = VectorImage::fromFile(set->defaultImageForCamera(views[0])->file());
            GeodesicWeight gw(2);
            AdaptiveWeight aw(2);
            const int cx = img.width() / 3, cy = img.height() / 2;
            gw.init_weights(img, cx, cy);
            aw.init_weights(img, cx, cy);
            std::vector<double> wv;
            for (int row = -2; row <= 2; ++row)
                for (int col = -2; col <= 2; ++col) wv.push_back(gw(row, col));
            for (int row = -2; row <= 2; ++row)
                for (int col = -2; col <= 2; ++col) wv.push_back(aw(row, col));
            dump(out + "/weights.bin", wv.data(), wv.size());
            std::printf("weights: geodesic centre %.3f, adaptive centre %.3f\n", gw(0, 0), aw(0, 0));
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "host_api_test: %s\n", e.what());
        return 1;
    }
    std::printf("OK\n");
    return 0;
}
