// GSL stand-in (see gsl_poly.h).
#ifndef SR_REF_SHIM_GSL_ERRNO
#define SR_REF_SHIM_GSL_ERRNO
typedef void gsl_error_handler_t(const char *reason, const char *file, int line, int gsl_errno);
inline gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *h) { return h; }
#endif
