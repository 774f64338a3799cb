// project/projectimage.hpp — ProjectImage (project/projectimage.hpp:30-60): an image file of an
// image set, with its exposure.
#ifndef SR_PROJECT_PROJECTIMAGE_HPP
#define SR_PROJECT_PROJECTIMAGE_HPP
#include "util/precompiled.hpp"
FORWARD_DECLARE(ProjectImage);
FORWARD_DECLARE(Camera);
class ProjectImage {
public:
    explicit ProjectImage(const std::string &file) : file_(file), exposure_(-1.0) {}
    const std::string &file() const { return file_; }
    double exposure() const { return exposure_; }
    void setExposure(double e) { exposure_ = e; }
    CameraPtr camera() const { return camera_.lock(); }
    void setCamera(CameraPtr c) { camera_ = c; }
private:
    std::string file_;
    double exposure_;
    CameraWeakPtr camera_;
};
#endif
