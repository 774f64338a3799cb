// project/imageset.hpp — ImageSet (project/imageset.hpp:40-100): the images taken by the
// project's cameras at one instant; defaultImageForCamera (project/imageset.cpp:74-78) is what
// MultiViewStereo::initialize reads.
#ifndef SR_PROJECT_IMAGESET_HPP
#define SR_PROJECT_IMAGESET_HPP
#include "project/projectimage.hpp"
FORWARD_DECLARE(ImageSet);
class ImageSet {
public:
    explicit ImageSet(const std::string &id = std::string()) : id_(id), name_(id) {}
    const std::string &id() const { return id_; }
    const std::string &name() const { return name_; }
    void setName(const std::string &n) { name_ = n; }
    const std::string &root() const { return root_; }
    void setRoot(const std::string &r) { root_ = r; }
    const std::vector<ProjectImagePtr> &images() const { return images_; }
    void addImageForCamera(CameraPtr cam, ProjectImagePtr image) {
        image->setCamera(cam);
        images_.push_back(image);
        byCamera_[cam.get()].push_back(image);
    }
    std::vector<ProjectImagePtr> imagesForCamera(CameraPtr cam) const {
        auto it = byCamera_.find(cam.get());
        return it == byCamera_.end() ? std::vector<ProjectImagePtr>() : it->second;
    }
    ProjectImagePtr defaultImageForCamera(CameraPtr cam) const {
        auto it = byCamera_.find(cam.get());
        return (it == byCamera_.end() || it->second.empty()) ? ProjectImagePtr() : it->second.front();
    }
private:
    std::string id_, name_, root_;
    std::vector<ProjectImagePtr> images_;
    std::map<const Camera *, std::vector<ProjectImagePtr>> byCamera_;
};
#endif
