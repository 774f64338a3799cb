// Stand-in for the reference's project/project.hpp (Qt XML document model, out of scope): the
// stereo classes only hold the pointer.
#ifndef SR_REF_SHIM_PROJECT
#define SR_REF_SHIM_PROJECT
class Project {};
#endif
