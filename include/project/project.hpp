// project/project.hpp — Project (project/project.hpp:50-120): cameras + image sets loaded from
// the project XML (project/project.cpp:74-226).  Qt-free: a small attribute scanner replaces
// QDomDocument; only the elements the dense-matching path reads are interpreted
// (<camera id name>, <projectionMatrix m11..m34>, <lensDistortion k1 k2 p1 p2 k3>,
//  <refractiveInterface px py dist refractiveRatio>, <imageSet id name root>, <image for file>).
#ifndef SR_PROJECT_PROJECT_HPP
#define SR_PROJECT_PROJECT_HPP
#include "project/camera.hpp"
#include "project/imageset.hpp"
#include <fstream>
#include <sstream>
FORWARD_DECLARE(Project);
class Project {
public:
    Project() {}
    explicit Project(const std::string &path) { load(path); }
    const std::string &projectPath() const { return path_; }
    const std::map<std::string, CameraPtr> &cameras() const { return cameras_; }
    const std::map<std::string, ImageSetPtr> &imageSets() const { return imageSets_; }
    CameraPtr camera(const std::string &id) const { auto it = cameras_.find(id); return it == cameras_.end() ? CameraPtr() : it->second; }
    ImageSetPtr imageSet(const std::string &id) const { auto it = imageSets_.find(id); return it == imageSets_.end() ? ImageSetPtr() : it->second; }
    void addCamera(CameraPtr c) { cameras_[c->id()] = c; }
    void addImageSet(ImageSetPtr s) { imageSets_[s->id()] = s; }

    void load(const std::string &path) {
        std::ifstream in(path.c_str());
        if (!in) throw std::runtime_error("Failed to open project file");  // project/project.cpp:97
        std::stringstream ss;
        ss << in.rdbuf();
        path_ = path;
        parse(ss.str());
    }
    //! Parses project XML text (exposed for tests).
    void parse(const std::string &xml) {
        cameras_.clear();
        imageSets_.clear();
        std::string dir = path_;
        const size_t slash = dir.find_last_of('/');
        dir = (slash == std::string::npos) ? std::string(".") : dir.substr(0, slash);
        CameraPtr cam;
        ImageSetPtr set;
        size_t pos = 0;
        while ((pos = xml.find('<', pos)) != std::string::npos) {
            const size_t end = xml.find('>', pos);
            if (end == std::string::npos) throw std::runtime_error("Failed to set XML content");  // :101
            const std::string tag = xml.substr(pos + 1, end - pos - 1);
            pos = end + 1;
            if (tag.empty() || tag[0] == '?' || tag[0] == '!') continue;
            if (tag[0] == '/') {
                const std::string name = tag.substr(1);
                if (name == "camera" && cam) { cameras_[cam->id()] = cam; cam.reset(); }
                if (name == "imageSet" && set) { if (!set->images().empty()) imageSets_[set->id()] = set; set.reset(); }
                continue;
            }
            const std::string name = tag.substr(0, tag.find_first_of(" \t\r\n/"));
            const std::map<std::string, std::string> at = attributes(tag);
            auto num = [&](const char *k, double def) { auto it = at.find(k); return it == at.end() ? def : std::atof(it->second.c_str()); };
            auto str = [&](const char *k, const std::string &def) { auto it = at.find(k); return it == at.end() ? def : it->second; };
            if (name == "camera") {
                cam.reset(new Camera(str("id", ""), str("name", str("id", ""))));
                if (tag.back() == '/') { cameras_[cam->id()] = cam; cam.reset(); }
            } else if (name == "projectionMatrix" && cam) {  // project/project.cpp:120-137
                ProjMat P;
                const char *keys[12] = {"m11", "m12", "m13", "m14", "m21", "m22", "m23", "m24", "m31", "m32", "m33", "m34"};
                for (int i = 0; i < 12; ++i) P.m[i] = num(keys[i], 0.0);
                cam->setP(P);
            } else if (name == "lensDistortion" && cam) {  // :140-149
                LensDistortions d = {{num("k1", 0), num("k2", 0), num("p1", 0), num("p2", 0), num("k3", 0)}};
                cam->setLensDistortion(d);
            } else if (name == "refractiveInterface" && cam) {  // :172-181
                cam->setRefractiveIndex(num("refractiveRatio", 1.0));
                cam->setPlane(Plane3d(cam->K().inverse() * Eigen::Vector3d(num("px", 0), num("py", 0), 1), num("dist", 0)));
            } else if (name == "imageSet") {  // :193-222
                set.reset(new ImageSet(str("id", "")));
                set->setName(str("name", set->id()));
                std::string root = str("root", "");
                set->setRoot(root.empty() ? dir : (root[0] == '/' ? root : dir + "/" + root));
            } else if (name == "image" && set) {
                ProjectImagePtr img(new ProjectImage(set->root() + "/" + str("file", "")));
                img->setExposure(num("exposure", -1.0));
                auto it = cameras_.find(str("for", ""));
                if (it != cameras_.end()) set->addImageForCamera(it->second, img);
            }
        }
    }
private:
    static std::map<std::string, std::string> attributes(const std::string &tag) {
        std::map<std::string, std::string> out;
        size_t p = tag.find_first_of(" \t\r\n");
        while (p != std::string::npos && p < tag.size()) {
            while (p < tag.size() && (tag[p] == ' ' || tag[p] == '\t' || tag[p] == '\r' || tag[p] == '\n')) ++p;
            const size_t eq = tag.find('=', p);
            if (eq == std::string::npos) break;
            const std::string key = tag.substr(p, eq - p);
            const size_t q0 = tag.find_first_of("\"'", eq);
            if (q0 == std::string::npos) break;
            const size_t q1 = tag.find(tag[q0], q0 + 1);
            if (q1 == std::string::npos) break;
            out[key] = tag.substr(q0 + 1, q1 - q0 - 1);
            p = q1 + 1;
        }
        return out;
    }
    std::string path_;
    std::map<std::string, CameraPtr> cameras_;
    std::map<std::string, ImageSetPtr> imageSets_;
};
#endif
