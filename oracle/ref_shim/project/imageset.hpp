// Stand-in for project/imageset.hpp: camera -> default image, as MultiViewStereo::initialize asks.
#ifndef SR_REF_SHIM_IMAGESET
#define SR_REF_SHIM_IMAGESET
#include <map>
#include "project/projectimage.hpp"
FORWARD_DECLARE(Camera);
FORWARD_DECLARE(ImageSet);
class ImageSet {
public:
    void setDefaultImage(CameraPtr cam, ProjectImagePtr img) { images_[cam.get()] = img; }
    ProjectImagePtr defaultImageForCamera(CameraPtr cam) const {
        std::map<const Camera *, ProjectImagePtr>::const_iterator it = images_.find(cam.get());
        return it == images_.end() ? ProjectImagePtr() : it->second;
    }
private:
    std::map<const Camera *, ProjectImagePtr> images_;
};
#endif
