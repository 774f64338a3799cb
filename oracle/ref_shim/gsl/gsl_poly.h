// GSL 1.14 is not in this image.  gsl_poly_complex_solve (the one GSL call of the reference's hot
// path, project/camera.cpp:77-80) is answered by the oracle's restatement of its published algorithm
// (balanced companion matrix + Hessenberg QR, oracle.cpp: poly_roots4), linked from liboracle.so.
// Everything around that call — the quartic's coefficients, the root acceptance test, the point
// on the interface — is the reference's own code.
#ifndef SR_REF_SHIM_GSL_POLY
#define SR_REF_SHIM_GSL_POLY
#include <cstddef>
extern "C" int orc_poly_roots4(const double *coeffs5, double *re, double *im);
struct gsl_poly_complex_workspace { int n; };
inline gsl_poly_complex_workspace *gsl_poly_complex_workspace_alloc(size_t n) { return new gsl_poly_complex_workspace{(int)n}; }
inline void gsl_poly_complex_workspace_free(gsl_poly_complex_workspace *w) { delete w; }
// z = packed (re, im) pairs; on failure GSL leaves z untouched (the reference zero-initialises it)
inline int gsl_poly_complex_solve(const double *a, size_t n, gsl_poly_complex_workspace *, double *z) {
    if (n != 5) return 1;
    double re[4], im[4];
    if (!orc_poly_roots4(a, re, im)) return 1;
    for (int i = 0; i < 4; ++i) { z[2 * i] = re[i]; z[2 * i + 1] = im[i]; }
    return 0;
}
#endif
