// Stand-in for project/projectimage.hpp: an image of an image set is a (registered) file name.
#ifndef SR_REF_SHIM_PROJECTIMAGE
#define SR_REF_SHIM_PROJECTIMAGE
#include <QObject>
FORWARD_DECLARE(ProjectImage);
class ProjectImage {
public:
    explicit ProjectImage(const QString &file) : file_(file) {}
    QString file() const { return file_; }
private:
    QString file_;
};
#endif
