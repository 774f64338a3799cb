// oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A line-by-line restatement, in plain double-precision C++ (+OpenMP over image rows, as
// the reference parallelises), of the dense-matching hot path of thegedge/StereoReconstruction.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library, and only as the checker / the timed CPU baseline.  Nothing under
// stereoreconstruction_b200/ links, imports or calls it.
//
// PARITY STATUS: PINNED against the reference's own code, compiled here from /root/reference where
// it lies (oracle/_ref/libref.so; oracle/Makefile, ref_glue*.cpp, ref_shim/): util/lineiter.cpp,
// util/ray.cpp, util/vectorimage.cpp, stereo/adaptiveweight.cpp, stereo/geodesicweight.cpp,
// project/camera.cpp, stereo/multiviewstereo.cpp and stereo/twoviewstereo.cpp — the whole path of
// SURVEY section 8(a), driven as the GUI drives it (initialize() -> runTask(); constructor ->
// computeDepthMaps()).  tests/test_oracle_vs_ref.py: this restatement reproduces the reference BIT FOR
// BIT end to end (neighbour rule, rasterised curves, both NCCs and SAD, K = 9 peak lists, selection
// rules, both cross-checks) and the camera model to rounding (1e-12; Eigen is a stand-in there).
// tests/golden/ref_*.npz hold the reference's outputs for machines without /root/reference.
// Not the reference's own: Qt, Project/ImageSet, Eigen's fixed-size matrices and the one GSL call are
// stand-ins (ref_shim/); the reference's label-mode branch is compiled out in the reference itself
// (twoviewstereo.cpp:283,308-329), so label mode is pinned by composing it, in the tests, from the
// reference's own compiled pieces (unproject, intersect, project, cost_ncc) for a pixel sample:
// index, depth and winning cost equal this restatement's label mode exactly.
//
// Third-party arithmetic not in /root/reference: GSL 1.14 gsl_poly_complex_solve
// (project/camera.cpp:77-80): eigenvalues of the balanced companion matrix by Hessenberg QR.
// Restated below (poly_roots4) from the published algorithm (EISPACK balanc + hqr).  Root
// ORDER is implementation-defined in GSL and only matters in a measure-zero band (SURVEY §8a
// G4); the oracle counts how often the reference-style "first acceptable root" differs from
// the unique physical root (orc_stats).
//
// Deviations from the literal reference, each deliberate and documented:
//  * TwoViewStereo::cost_ncc uses unqualified abs() on a double (twoviewstereo.cpp:976); on
//    the author's libc++ toolchain that is the floating overload.  We use std::fabs.
//  * label mode stores cost volumes indexed by label d, not by a counter that only advances
//    on success (twoviewstereo.cpp:318, a bug that shifts labels).
//  * double->int conversions of projected coordinates (lineiter.hpp:34 ctor args) are UB when
//    out of range; we fix the x86 behaviour (cvttsd2si: INT_MIN) so CPU and GPU agree.
//  * the OpenMP loop uses `continue`-free bodies; cancellation is not modelled.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double NaN = std::numeric_limits<double>::quiet_NaN();
const double INF = std::numeric_limits<double>::infinity();

// ---------------------------------------------------------------------------------
// minimal vector algebra (stands in for Eigen::Vector3d / Matrix3d, closed-form 3x3 only)
struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) {
    double n = norm(a);
    return {a.x / n, a.y / n, a.z / n};
}
inline V3 mul(const double *M, V3 v) {  // row-major 3x3
    return {M[0] * v.x + M[1] * v.y + M[2] * v.z, M[3] * v.x + M[4] * v.y + M[5] * v.z,
            M[6] * v.x + M[7] * v.y + M[8] * v.z};
}

// ---------------------------------------------------------------------------------
// util/vectorimage.hpp:28-74  RGBA, util/vectorimage.cpp:115-155 pixel()/sample()
struct RGBA {
    double r, g, b, a;
    bool isValid() const { return !(std::isnan(r) || std::isnan(g) || std::isnan(b)); }
    double toGray() const { return (0.11 * r + 0.59 * g + 0.3 * b); }  // vectorimage.hpp:60-62
};
const RGBA INVALID = {NaN, NaN, NaN, 255.0};  // vectorimage.cpp:35-38

struct Image {
    int w = 0, h = 0;
    std::vector<RGBA> data;
    void fromRGBA8(const uint8_t *p, int w_, int h_) {  // vectorimage.cpp:48-68
        w = w_;
        h = h_;
        data.resize((size_t)w * h);
        for (size_t i = 0; i < (size_t)w * h; ++i)
            data[i] = {(double)p[4 * i], (double)p[4 * i + 1], (double)p[4 * i + 2],
                       (double)p[4 * i + 3]};
    }
    const RGBA &pixel(int x, int y) const {  // vectorimage.cpp:115-119
        if (x < 0 || y < 0 || x >= w || y >= h) return INVALID;
        return data[(size_t)y * w + x];
    }
    RGBA sample(double x, double y) const {  // vectorimage.cpp:129-155
        RGBA r = INVALID;
        if (x >= 0 && y >= 0 && x + 1 < w && y + 1 < h) {
            int ix = (int)x, iy = (int)y;
            double dx = x - ix, dy = y - iy;
            r.r = r.g = r.b = 0.0;
            auto acc = [&](const RGBA &t, double s) {
                r.r += t.r * s;
                r.g += t.g * s;
                r.b += t.b * s;
            };
            acc(data[ix + (size_t)iy * w], (1 - dx) * (1 - dy));
            acc(data[ix + (size_t)(iy + 1) * w], (1 - dx) * dy);
            acc(data[ix + 1 + (size_t)iy * w], dx * (1 - dy));
            acc(data[ix + 1 + (size_t)(iy + 1) * w], dx * dy);
        }
        return r;
    }
    // sample() at integer coordinates: bilinear weights are exactly (1,0,0,0), so the value is
    // the pixel itself; only the validity rule differs from pixel() (last row/col invalid).
    // tests/test_oracle.py checks this against sample().
    const RGBA &sampleInt(int x, int y) const {
        if (x >= 0 && y >= 0 && x + 1 < w && y + 1 < h) return data[(size_t)y * w + x];
        return INVALID;
    }
};

// Masks are VectorImages compared against WHITE=(255,255,255,255) (vectorimage.cpp:26,
// operator== vectorimage.hpp:64-69 includes alpha).  The harness reduces a mask to one byte
// per pixel: 255 <=> that comparison holds.  Out of bounds pixel() is INVALID => != WHITE.
struct Mask {
    int w = 0, h = 0;
    std::vector<uint8_t> data;
    void set(const uint8_t *p, int w_, int h_) {
        w = w_;
        h = h_;
        data.assign((size_t)w * h, 255);
        if (p) std::memcpy(data.data(), p, (size_t)w * h);
    }
    bool white(int x, int y) const {
        if (x < 0 || y < 0 || x >= w || y >= h) return false;
        return data[(size_t)y * w + x] == 255;
    }
};

// ---------------------------------------------------------------------------------
// util/ray.hpp:30-72, util/plane.hpp:26-47
struct Ray {
    V3 src{0, 0, 0}, dir{0, 0, 1};
    Ray() {}
    Ray(V3 s, V3 d) : src(s), dir(normalized(d)) {}  // ray.cpp:30-33
    void setDirection(V3 v) { dir = normalized(v); }  // ray.hpp:41
    V3 point(double t) const { return src + t * dir; }
};
struct Plane {
    V3 n{0, 0, 1};
    double d = 0;
    Plane() {}
    Plane(V3 normal, double dist) : n(normalized(normal)), d(dist) {}            // plane.hpp:33
    Plane(V3 normal, V3 x0) : n(normalized(normal)) { d = dot(n, x0); }         // plane.hpp:34
    V3 x0() const { return d * n; }                                               // plane.hpp:42
};

// util/ray.cpp:78-88
bool intersect(const Ray &R, const Plane &P, V3 &p) {
    double nd = dot(P.n, R.dir);
    if (std::fabs(nd) < 1e-10) return false;
    double t = dot(P.n, P.x0() - R.src) / nd;
    if (t < 1e-10) return false;
    p = R.point(t);
    return true;
}

// util/ray.cpp:92-106
bool refract(const Ray &R, const Plane &P, double n, Ray &Rout) {
    V3 p;
    if (intersect(R, P, p)) {
        double cosI = -(dot(P.n, R.dir));
        double cosT2 = 1.0 - (1.0 - cosI * cosI) / (n * n);
        if (cosT2 > 0.0) {
            double sign = (cosI > 0.0 ? -1.0 : 1.0);
            V3 d = R.dir + (cosI + n * sign * std::sqrt(cosT2)) * P.n;
            Rout.src = p;
            Rout.setDirection(d);
            return true;
        }
    }
    return false;
}

// util/ray.cpp:53-74
void closestPoints(const Ray &A, const Ray &B, V3 &p1, V3 &p2) {
    V3 w0 = A.src - B.src;
    double a = dot(A.dir, A.dir);
    double b = dot(A.dir, B.dir);
    double c = dot(B.dir, B.dir);
    double d = dot(A.dir, w0);
    double e = dot(B.dir, w0);
    double den = 1.0 / (a * c - b * b);
    double tl = (b * e - c * d) * den;
    double tr = (a * e - b * d) * den;
    p1 = A.src;
    p2 = B.src;
    if (tl > 0) p1 = p1 + tl * A.dir;
    if (tr > 0) p2 = p2 + tr * B.dir;
}

// ---------------------------------------------------------------------------------
// GSL gsl_poly_complex_solve restated: balanced companion matrix + Hessenberg QR (hqr).
// Input a[0..4] = coefficients, a[4] the leading one (camera.cpp:73 ordering {e,d,c,b,a}).
// Output re[4], im[4].  Returns false on non-convergence (GSL would report an error that the
// reference's no-op handler swallows, camera.cpp:63-65,77; roots stay 0 there).
bool poly_roots4(const double *a, double *re, double *im) {
    const int n = 4;
    double m[n][n];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) m[i][j] = 0.0;
    for (int i = 1; i < n; ++i) m[i][i - 1] = 1.0;
    for (int i = 0; i < n; ++i) m[i][n - 1] = -a[i] / a[n];
    // balance (Parlett-Reinsch, radix 2, no permutation)
    const double RADIX = 2.0, RADIX2 = 4.0;
    bool notconv = true;
    while (notconv) {
        notconv = false;
        for (int i = 0; i < n; ++i) {
            double cn = 0, rn = 0;
            for (int j = 0; j < n; ++j)
                if (j != i) {
                    cn += std::fabs(m[j][i]);
                    rn += std::fabs(m[i][j]);
                }
            if (cn == 0 || rn == 0) continue;
            double g = rn / RADIX, f = 1, s = cn + rn;
            while (cn < g) {
                f *= RADIX;
                cn *= RADIX2;
            }
            g = rn * RADIX;
            while (cn > g) {
                f /= RADIX;
                cn /= RADIX2;
            }
            if ((rn + cn) < 0.95 * s * f) {
                notconv = true;
                g = 1 / f;
                for (int j = 0; j < n; ++j) m[i][j] *= g;
                for (int j = 0; j < n; ++j) m[j][i] *= f;
            }
        }
    }
    // hqr (EISPACK), matrix already upper Hessenberg
    auto SIGN = [](double a_, double b_) { return b_ >= 0 ? std::fabs(a_) : -std::fabs(a_); };
    double anorm = 0;
    for (int i = 0; i < n; ++i)
        for (int j = std::max(i - 1, 0); j < n; ++j) anorm += std::fabs(m[i][j]);
    int nn = n - 1;
    double t = 0, p = 0, q = 0, r = 0, s = 0, x, y, z, w, u, v;
    while (nn >= 0) {
        int its = 0, l;
        do {
            for (l = nn; l >= 1; --l) {
                s = std::fabs(m[l - 1][l - 1]) + std::fabs(m[l][l]);
                if (s == 0) s = anorm;
                if (std::fabs(m[l][l - 1]) + s == s) {
                    m[l][l - 1] = 0;
                    break;
                }
            }
            x = m[nn][nn];
            if (l == nn) {
                re[nn] = x + t;
                im[nn--] = 0;
            } else {
                y = m[nn - 1][nn - 1];
                w = m[nn][nn - 1] * m[nn - 1][nn];
                if (l == nn - 1) {
                    p = 0.5 * (y - x);
                    q = p * p + w;
                    z = std::sqrt(std::fabs(q));
                    x += t;
                    if (q >= 0) {
                        z = p + SIGN(z, p);
                        re[nn - 1] = re[nn] = x + z;
                        if (z != 0) re[nn] = x - w / z;
                        im[nn - 1] = im[nn] = 0;
                    } else {
                        re[nn - 1] = re[nn] = x + p;
                        im[nn - 1] = -(im[nn] = z);
                    }
                    nn -= 2;
                } else {
                    if (its == 60) return false;
                    if (its == 10 || its == 20) {
                        t += x;
                        for (int i = 0; i <= nn; ++i) m[i][i] -= x;
                        s = std::fabs(m[nn][nn - 1]) + std::fabs(m[nn - 1][nn - 2]);
                        y = x = 0.75 * s;
                        w = -0.4375 * s * s;
                    }
                    ++its;
                    int mm;
                    for (mm = nn - 2; mm >= l; --mm) {
                        z = m[mm][mm];
                        r = x - z;
                        s = y - z;
                        p = (r * s - w) / m[mm + 1][mm] + m[mm][mm + 1];
                        q = m[mm + 1][mm + 1] - z - r - s;
                        r = m[mm + 2][mm + 1];
                        s = std::fabs(p) + std::fabs(q) + std::fabs(r);
                        p /= s;
                        q /= s;
                        r /= s;
                        if (mm == l) break;
                        u = std::fabs(m[mm][mm - 1]) * (std::fabs(q) + std::fabs(r));
                        v = std::fabs(p) * (std::fabs(m[mm - 1][mm - 1]) + std::fabs(z) +
                                            std::fabs(m[mm + 1][mm + 1]));
                        if (u + v == v) break;
                    }
                    for (int i = mm + 2; i <= nn; ++i) {
                        m[i][i - 2] = 0;
                        if (i != mm + 2) m[i][i - 3] = 0;
                    }
                    for (int k = mm; k <= nn - 1; ++k) {
                        if (k != mm) {
                            p = m[k][k - 1];
                            q = m[k + 1][k - 1];
                            r = 0;
                            if (k != nn - 1) r = m[k + 2][k - 1];
                            if ((x = std::fabs(p) + std::fabs(q) + std::fabs(r)) != 0) {
                                p /= x;
                                q /= x;
                                r /= x;
                            }
                        }
                        if ((s = SIGN(std::sqrt(p * p + q * q + r * r), p)) != 0) {
                            if (k == mm) {
                                if (l != mm) m[k][k - 1] = -m[k][k - 1];
                            } else
                                m[k][k - 1] = -s * x;
                            p += s;
                            x = p / s;
                            y = q / s;
                            z = r / s;
                            q /= p;
                            r /= p;
                            for (int j = k; j <= nn; ++j) {
                                p = m[k][j] + q * m[k + 1][j];
                                if (k != nn - 1) {
                                    p += r * m[k + 2][j];
                                    m[k + 2][j] -= p * z;
                                }
                                m[k + 1][j] -= p * y;
                                m[k][j] -= p * x;
                            }
                            int mmin = nn < k + 3 ? nn : k + 3;
                            for (int i = l; i <= mmin; ++i) {
                                p = x * m[i][k] + y * m[i][k + 1];
                                if (k != nn - 1) {
                                    p += z * m[i][k + 2];
                                    m[i][k + 2] -= p * r;
                                }
                                m[i][k + 1] -= p * q;
                                m[i][k] -= p;
                            }
                        }
                    }
                }
            }
        } while (l < nn - 1);
    }
    return true;
}

// statistics about the refractive root selection (see header)
struct Stats {
    long long project_calls = 0;       // refractive projections attempted
    long long quartic_fail = 0;        // QR did not converge
    long long root_mismatch = 0;       // reference-style root differs from physical root > 1e-9*max(1,r)
    long long no_root = 0;             // reference-style selection found no acceptable root
    double max_root_diff = 0;
};
Stats g_stats;

inline bool iszero(double x, double eps = 1e-10) { return (x <= eps && x >= -eps); }  // camera.cpp:52-53

// project/camera.cpp:68-86
void findRoots(double a, double b, double c, double d, double e, double &r1, double &r2,
               double &r3, double &r4, bool &ok) {
    const double coeffs[] = {e, d, c, b, a};
    double re[4] = {0, 0, 0, 0}, im[4] = {0, 0, 0, 0};
    ok = poly_roots4(coeffs, re, im);
    r1 = (iszero(im[0]) ? re[0] : NaN);
    r2 = (iszero(im[1]) ? re[1] : NaN);
    r3 = (iszero(im[2]) ? re[2] : NaN);
    r4 = (iszero(im[3]) ? re[3] : NaN);
}

// The unique root in [0,r] of the un-squared Snell equation
//   g(x) = x/sqrt(x^2+d^2) - n (r-x)/sqrt((r-x)^2 + h^2),  h = z-d      
// g is strictly increasing with g(0) < 0 < g(r): safeguarded Newton (bisection fallback).
double snellRoot(double r, double d, double h, double n) {
    const double dd = d * d, hh = h * h;
    double lo = 0, hi = r;
    double x = n * std::fabs(d) * r / (std::fabs(h) + n * std::fabs(d) + 1e-300);  // paraxial start
    if (!(x > lo && x < hi)) x = 0.5 * r;
    for (int it = 0; it < 200; ++it) {
        double a = x * x + dd, b = (r - x) * (r - x) + hh;
        double ia = 1.0 / std::sqrt(a), ib = 1.0 / std::sqrt(b);
        double g = x * ia - n * (r - x) * ib;
        if (g == 0) break;
        if (g < 0)
            lo = x;
        else
            hi = x;
        double gp = dd * ia * ia * ia + n * hh * ib * ib * ib;
        double xn = x - g / gp;
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        if (std::fabs(xn - x) <= 4e-16 * r || hi - lo <= 4e-16 * r) {
            x = xn;
            break;
        }
        x = xn;
    }
    return x;
}

enum RootMode { ROOT_QUARTIC = 0, ROOT_MONOTONE = 1, ROOT_BOTH = 2 };

// project/camera.cpp:95-138
bool projectRefraction(V3 &p, const Plane &P, double n, int mode) {
    const V3 bn = normalized(P.n);  // linalg.hpp:34
    const V3 proj = dot(bn, p) * bn;
    const V3 radv = p - proj;
    const double y = radv.y;
    const double z = norm(proj);
    const double r = norm(radv);
    const double d = P.d;
    const double rr = r * r, nn = n * n, dd = d * d;
    V3 dir = {radv.x / r, radv.y / r, radv.z / r};  // Eigen normalize(): r==0 -> NaN

    ++g_stats.project_calls;
    double roots[4] = {NaN, NaN, NaN, NaN};
    double xm = NaN;
    if (mode != ROOT_QUARTIC) xm = (r > 0) ? snellRoot(r, d, z - d, n) : NaN;
    if (mode != ROOT_MONOTONE) {
        bool ok;
        findRoots(nn - 1, -2 * r * (nn - 1), rr * (nn - 1) + dd * nn - (z - d) * (z - d),
                  -2 * dd * nn * r, dd * nn * rr, roots[0], roots[1], roots[2], roots[3], ok);
        if (!ok) ++g_stats.quartic_fail;
    } else {
        roots[0] = xm;
    }
    // Find the root that makes sense (camera.cpp:118-135)
    for (int index = 0; index < 4; ++index) {
        if (!std::isnan(roots[index])) {
            const V3 pp = roots[index] * dir;
            const double py = pp.y;
            bool accept = false;
            if (py > -1e-3 && y > -1e-3) {
                if (py < y + 1e-3) accept = true;
            } else if (py < 1e-3 && y < 1e-3) {
                if (y < py + 1e-3) accept = true;
            }
            if (accept) {
                if (mode == ROOT_BOTH) {
                    double diff = std::fabs(roots[index] - xm);
                    if (diff > g_stats.max_root_diff) g_stats.max_root_diff = diff;
                    if (!(diff <= 1e-9 * std::max(1.0, r))) ++g_stats.root_mismatch;
                }
                p = pp + P.x0();
                return true;
            }
        }
    }
    ++g_stats.no_root;
    return false;
}

// ---------------------------------------------------------------------------------
// Camera POD, identical layout to include/sr_b200.h::sr_camera (kept separate on purpose:
// the oracle does not include product headers).
struct Camera {
    double K[9], Kinv[9], R[9], Rinv[9], t[3], C[3];
    double dist[5];
    double plane_n[3];
    double plane_d;
    double n;
    double prin_dir[3];
    int32_t is_refractive, is_distorted;

    V3 tv() const { return {t[0], t[1], t[2]}; }
    V3 Cv() const { return {C[0], C[1], C[2]}; }
    V3 prin() const { return {prin_dir[0], prin_dir[1], prin_dir[2]}; }
    Plane plane() const {
        Plane P;
        P.n = {plane_n[0], plane_n[1], plane_n[2]};
        P.d = plane_d;
        return P;
    }
    V3 fromGlobalToLocal(V3 p) const { return mul(R, p) + tv(); }  // camera.cpp:346-348
    Ray fromLocalToGlobal(const Ray &r) const {                     // camera.cpp:372-376
        V3 direction = mul(Rinv, r.dir);
        V3 source = mul(Rinv, r.src - tv());
        return Ray(source, direction);
    }

    // project/camera.cpp:380-419.  p: global point in, pixel (x,y,1) out.
    bool project(V3 &p, int mode) const {
        V3 point = fromGlobalToLocal(p);
        if (is_refractive) {
            if (!projectRefraction(point, plane(), n, mode)) {
                p = {NaN, NaN, NaN};
                return false;
            }
        }
        p = mul(K, point);
        p = {p.x / p.z, p.y / p.z, p.z / p.z};
        if (is_distorted) {
            const double cx = K[2], cy = K[5], fx = K[0], fy = K[4];
            double x = (p.x - cx) / fx;
            double y = (p.y - cy) / fy;
            {
                const double *k = dist;
                const double r2 = x * x + y * y;
                const double cdist = 1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2;
                const double xo = x, yo = y;  // the reference updates x then uses the NEW x for y:
                x = xo * cdist + 2 * k[2] * xo * yo + k[3] * (r2 + 2 * xo * xo);
                // camera.cpp:411-412: y is computed AFTER x was overwritten (x is a reference
                // into p), so the tangential terms of y see the distorted x.
                y = yo * cdist + k[2] * (r2 + 2 * yo * yo) + 2 * k[3] * x * yo;
            }
            p.x = fx * x + cx;
            p.y = fy * y + cy;
        }
        return true;
    }

    // project/camera.cpp:423-459
    Ray unproject(double px, double py) const {
        double x = px, y = py;
        if (is_distorted) {
            const double cx = K[2], cy = K[5];
            const double ifx = 1.0 / K[0], ify = 1.0 / K[4];
            const double x0 = x = (x - cx) * ifx;
            const double y0 = y = (y - cy) * ify;
            const double *k = dist;
            for (int j = 0; j < 5; j++) {
                const double r2 = x * x + y * y;
                const double icdist = 1.0 / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
                const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
                const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
                x = (x0 - deltaX) * icdist;
                y = (y0 - deltaY) * icdist;
            }
            x /= ifx;
            y /= ify;
            x += cx;
            y += cy;
        }
        Ray ray(V3{0, 0, 0}, mul(Kinv, V3{x, y, 1.0}));
        if (is_refractive) refract(ray, plane(), n, ray);  // return value ignored, camera.cpp:455-456
        return fromLocalToGlobal(ray);
    }
};

// ---------------------------------------------------------------------------------
// Support weights.  weights[(row+r)*(2r+1) + (col+r)] == operator()(row,col).
struct WeightFunc {
    int kind = 1, radius = 0;
    std::vector<double> distance_weights, weights;
    void initialize(int kind_, int r) {
        kind = kind_;
        radius = r;
        distance_weights.resize(r + 1);
        for (int ind = 0; ind <= r; ++ind)
            distance_weights[ind] = std::exp(-ind / (1.0 * r));  // adaptiveweight.cpp:36-38
        weights.assign((size_t)(2 * r + 1) * (2 * r + 1), 0.0);
    }
    double operator()(int row, int col) const {
        return weights[(size_t)(row + radius) * (2 * radius + 1) + (col + radius)];
    }
    double &at(int row, int col) {
        return weights[(size_t)(row + radius) * (2 * radius + 1) + (col + radius)];
    }
    void init_weights(const Image &img, int cx, int cy) {
        if (kind == 0)
            init_adaptive(img, cx, cy);
        else
            init_geodesic(img, cx, cy);
    }
    // stereo/adaptiveweight.cpp:47-79
    void init_adaptive(const Image &img, int cx, int cy) {
        const double COLOR_SIGMA = 10.0;
        const RGBA crgb = img.pixel(cx, cy);
        for (int row = -radius; row <= radius; ++row)
            for (int col = -radius; col <= radius; ++col) {
                double weight = 0.0;
                RGBA rgb = img.pixel(cx + col, cy + row);
                if (rgb.isValid()) {
                    rgb.r -= crgb.r;
                    rgb.g -= crgb.g;
                    rgb.b -= crgb.b;
                    const double diff = std::sqrt(rgb.r * rgb.r + rgb.g * rgb.g + rgb.b * rgb.b);
                    const double w1 = distance_weights[std::abs(row)] * distance_weights[std::abs(col)];
                    const double w2 = std::exp(-diff / COLOR_SIGMA);
                    weight = w1 * w2;
                    if (std::isnan(weight)) weight = 0.0;
                }
                at(row, col) = weight;
            }
    }
    // stereo/geodesicweight.cpp:59-131
    void init_geodesic(const Image &img, int cx, int cy) {
        const double GEODESIC_SIGMA = 50.0;
        const int NUM_ITERS = 3;
        const int KERNEL_SIZE = 8;
        static const double K1[] = {-1, -1, 0, -1, 1, -1, -1, 0};
        static const double K2[] = {-1, 1, 0, 1, 1, 1, 1, 0};
        std::fill(weights.begin(), weights.end(), 1000000.0);
        at(0, 0) = 0.0;
        for (int iter = 0; iter < NUM_ITERS; ++iter) {
            for (int y = -radius; y <= radius; ++y) {
                for (int x = -radius; x <= radius; ++x) {
                    const RGBA &rgb1 = img.pixel(cx + x, cy + y);
                    if (!rgb1.isValid()) continue;
                    double &weight = at(y, x);
                    for (int ind = 0; ind < KERNEL_SIZE; ind += 2) {
                        int dx = (int)K1[ind + 0], dy = (int)K1[ind + 1];
                        if (x + dx > radius || y + dy > radius || x + dx < -radius || y + dy < -radius)
                            continue;
                        RGBA rgb2 = img.pixel(cx + x + dx, cy + y + dy);
                        if (rgb2.isValid()) {
                            rgb2.r -= rgb1.r;
                            rgb2.g -= rgb1.g;
                            rgb2.b -= rgb1.b;
                            double diff = std::sqrt(rgb2.r * rgb2.r + rgb2.g * rgb2.g + rgb2.b * rgb2.b);
                            double cost = at(y + dy, x + dx);
                            weight = std::min(weight, cost + diff);
                        }
                    }
                }
            }
            for (int y = radius; y >= -radius; --y) {
                for (int x = radius; x >= -radius; --x) {
                    const RGBA &rgb1 = img.pixel(cx + x, cy + y);
                    if (!rgb1.isValid()) continue;
                    double &weight = at(y, x);
                    for (int ind = 0; ind < KERNEL_SIZE; ind += 2) {
                        int dx = (int)K2[ind + 0], dy = (int)K2[ind + 1];
                        if (x + dx > radius || y + dy > radius || x + dx < -radius || y + dy < -radius)
                            continue;
                        RGBA rgb2 = img.pixel(cx + x + dx, cy + y + dy);
                        if (rgb2.isValid()) {
                            rgb2.r -= rgb1.r;
                            rgb2.g -= rgb1.g;
                            rgb2.b -= rgb1.b;
                            double diff = std::sqrt(rgb2.r * rgb2.r + rgb2.g * rgb2.g + rgb2.b * rgb2.b);
                            double cost = at(y + dy, x + dx);
                            weight = std::min(weight, cost + diff);
                        }
                    }
                }
            }
        }
        for (double &w : weights) w = std::exp(-w / GEODESIC_SIGMA);
    }
};

// ---------------------------------------------------------------------------------
// Matching costs.
// stereo/multiviewstereo.cpp:113-189 (higher is better; pixel() taps; masks compiled out)
double cost_ncc_mvs(const Image &img1, const Image &img2, int x1, int y1, int x2, int y2,
                    const WeightFunc &weightFunc) {
    const int R = weightFunc.radius;
    double meanL = 0, meanR = 0, totalWeight = 0.0;
    for (int row = -R; row <= R; ++row)
        for (int col = -R; col <= R; ++col) {
            const RGBA &lrgb = img1.pixel(x1 + col, y1 + row);
            if (!lrgb.isValid()) continue;
            const RGBA &rrgb = img2.pixel(x2 + col, y2 + row);
            if (!rrgb.isValid()) continue;
            const double weight = weightFunc(row, col);
            if (weight > 1e-10) {
                meanL += weight * lrgb.toGray();
                meanR += weight * rrgb.toGray();
                totalWeight += weight;
            }
        }
    if (totalWeight < 1e-10) return 0;
    meanL /= totalWeight;
    meanR /= totalWeight;
    double sum1 = 0, sum2 = 0, sum3 = 0;
    for (int row = -R; row <= R; ++row)
        for (int col = -R; col <= R; ++col) {
            const RGBA &lrgb = img1.pixel(x1 + col, y1 + row);
            if (!lrgb.isValid()) continue;
            const RGBA &rrgb = img2.pixel(x2 + col, y2 + row);
            if (!rrgb.isValid()) continue;
            const double weight = weightFunc(row, col);
            if (weight > 1e-10) {
                const double pixel_gray_l = weight * lrgb.toGray();
                const double pixel_gray_r = weight * rrgb.toGray();
                sum1 += (pixel_gray_l - meanL) * (pixel_gray_r - meanR);
                sum2 += (pixel_gray_l - meanL) * (pixel_gray_l - meanL);
                sum3 += (pixel_gray_r - meanR) * (pixel_gray_r - meanR);
            }
        }
    if (sum2 * sum3 < 1e-10) return 0;
    return sum1 / std::sqrt(sum2 * sum3);
}

const int BAD_RET = 1000;             // twoviewstereo.cpp:65
const double MAX_COLOR_DIFF = 120;    // twoviewstereo.cpp:74

// stereo/twoviewstereo.cpp:909-977 (lower is better; sample() taps; masks enabled)
double cost_ncc_two(const Image &left, const Image &right, const Mask &leftMask,
                    const Mask &rightMask, int x1, int y1, int x2, int y2,
                    const WeightFunc &weightFunc) {
    const int R = weightFunc.radius;
    double meanL = 0, meanR = 0, totalWeight = 0.0;
    for (int row = -R; row <= R; ++row)
        for (int col = -R; col <= R; ++col) {
            if (!leftMask.white(x1 + col, y1 + row)) continue;
            if (!rightMask.white(x2 + col, y2 + row)) continue;
            const RGBA &lrgb = left.sampleInt(x1 + col, y1 + row);
            if (!lrgb.isValid()) continue;
            const RGBA &rrgb = right.sampleInt(x2 + col, y2 + row);
            if (!rrgb.isValid()) continue;
            const double weight = weightFunc(row, col);
            if (weight > 1e-10) {
                meanL += weight * lrgb.toGray();
                meanR += weight * rrgb.toGray();
                totalWeight += weight;
            }
        }
    if (totalWeight < 1e-10) return BAD_RET;
    meanL /= totalWeight;
    meanR /= totalWeight;
    double sum1 = 0, sum2 = 0, sum3 = 0;
    for (int row = -R; row <= R; ++row)
        for (int col = -R; col <= R; ++col) {
            const RGBA &lrgb = left.sampleInt(x1 + col, y1 + row);
            const RGBA &rrgb = right.sampleInt(x2 + col, y2 + row);
            if (!leftMask.white(x1 + col, y1 + row)) continue;
            if (!rightMask.white(x2 + col, y2 + row)) continue;
            if (!lrgb.isValid()) continue;
            if (!rrgb.isValid()) continue;
            const double weight = weightFunc(row, col);
            if (weight > 1e-10) {
                const double pixel_gray_l = weight * lrgb.toGray();
                const double pixel_gray_r = weight * rrgb.toGray();
                sum1 += (pixel_gray_l - meanL) * (pixel_gray_r - meanR);
                sum2 += (pixel_gray_l - meanL) * (pixel_gray_l - meanL);
                sum3 += (pixel_gray_r - meanR) * (pixel_gray_r - meanR);
            }
        }
    // std::min(120.0, NaN) returns 120 (first argument) — keep that behaviour.
    const double v = 255 * (1.0 - std::fabs(sum1) / std::sqrt(sum2 * sum3));
    return (v < MAX_COLOR_DIFF) ? v : MAX_COLOR_DIFF;
}

// stereo/twoviewstereo.cpp:864-905 (defined, never called by the reference)
double cost_sad_two(const Image &left, const Image &right, const Mask &leftMask,
                    const Mask &rightMask, int x1, int y1, int x2, int y2,
                    const WeightFunc &weightFunc) {
    const int R = weightFunc.radius;
    int numPixels = 0;
    double sum = 0.0, totalWeight = 0.0;
    for (int row = -R; row <= R; ++row)
        for (int col = -R; col <= R; ++col) {
            if (!leftMask.white(x1 + col, y1 + row)) continue;
            if (!rightMask.white(x2 + col, y2 + row)) continue;
            const RGBA &lrgb = left.sampleInt(x1 + col, y1 + row);
            if (!lrgb.isValid()) continue;
            const RGBA &rrgb = right.pixel(x2 + col, y2 + row);
            if (!rrgb.isValid()) continue;
            double weight = weightFunc(row, col);
            if (weight > 1e-10) {
                double diff = std::fabs(lrgb.toGray() - rrgb.toGray());
                sum += weight * std::min(MAX_COLOR_DIFF, diff);
                totalWeight += weight;
                ++numPixels;
            }
        }
    if (numPixels <= 4 || totalWeight <= 1e-10) return BAD_RET;
    return (sum / totalWeight);
}

// ---------------------------------------------------------------------------------
// util/lineiter.hpp:32-118 + util/lineiter.cpp:35-88
inline int to_int_x86(double v) {  // implicit double->int at lineiter.hpp:34 / cost_ncc args
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
    return (int)v;
}
inline int wrap_mul(int a, int b) { return (int)(uint32_t)((int64_t)a * (int64_t)b); }
inline int wrap_sub(int a, int b) { return (int)((uint32_t)a - (uint32_t)b); }
inline int wrap_add(int a, int b) { return (int)((uint32_t)a + (uint32_t)b); }
inline int wrap_div(int a, int b) {
    if (b == 0 || (a == INT32_MIN && b == -1)) return 0;  // x86 would trap; never reached for finite clips
    return a / b;
}
const int CS_LEFT = 1, CS_RIGHT = 2, CS_BOTTOM = 4, CS_TOP = 8;
int outCode(int x, int y, int w, int h) {
    int code = 0;
    if (x < 0) code |= CS_LEFT;
    else if (x > w) code |= CS_RIGHT;
    if (y < 0) code |= CS_BOTTOM;
    else if (y > h) code |= CS_TOP;
    return code;
}
bool clipLine(int &x0, int &y0, int &x1, int &y1, int w, int h) {
    w--;
    h--;
    int outcode0 = outCode(x0, y0, w, h), outcode1 = outCode(x1, y1, w, h);
    bool accept = false;
    int guard = 0;
    while (true) {
        if (!(outcode0 | outcode1)) {
            accept = true;
            break;
        } else if (outcode0 & outcode1) {
            break;
        } else {
            if (++guard > 64) break;  // wrapped arithmetic could in principle cycle; never seen
            int x = 0, y = 0;
            int outcodeOut = outcode0 ? outcode0 : outcode1;
            if (outcodeOut & CS_TOP) {
                x = wrap_add(x0, wrap_div(wrap_mul(wrap_sub(x1, x0), wrap_sub(h, y0)), wrap_sub(y1, y0)));
                y = h;
            } else if (outcodeOut & CS_BOTTOM) {
                x = wrap_add(x0, wrap_div(wrap_mul(wrap_sub(x1, x0), wrap_sub(0, y0)), wrap_sub(y1, y0)));
                y = 0;
            } else if (outcodeOut & CS_RIGHT) {
                y = wrap_add(y0, wrap_div(wrap_mul(wrap_sub(y1, y0), wrap_sub(w, x0)), wrap_sub(x1, x0)));
                x = w;
            } else if (outcodeOut & CS_LEFT) {
                y = wrap_add(y0, wrap_div(wrap_mul(wrap_sub(y1, y0), wrap_sub(0, x0)), wrap_sub(x1, x0)));
                x = 0;
            }
            if (outcodeOut == outcode0) {
                x0 = x;
                y0 = y;
                outcode0 = outCode(x0, y0, w, h);
            } else {
                x1 = x;
                y1 = y;
                outcode1 = outCode(x1, y1, w, h);
            }
        }
    }
    return accept;
}

struct LineIterator {
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0, x = 1, y = 1;
    int error = 0, ystep = 0, deltax = 0, deltay = 0;
    bool steep = false;
    LineIterator(int ax0, int ay0, int ax1, int ay1) {
        initialize(ax0, ay0, ax1, ay1);
        x0 = ax0; y0 = ay0; x1 = ax1; y1 = ay1;
        reset();
    }
    LineIterator(int ax0, int ay0, int ax1, int ay1, int w, int h) {
        if (clipLine(ax0, ay0, ax1, ay1, w, h)) {
            initialize(ax0, ay0, ax1, ay1);
            x0 = ax0; y0 = ay0; x1 = ax1; y1 = ay1;
            reset();
        } else {
            // lineiter.hpp:58-61 assigns the ctor PARAMETER x1 (not the member): members keep
            // their initialisers x1=0, x=1 => hasNext() is false.
        }
    }
    bool hasNext() const { return (x <= x1); }
    void current(int &ox, int &oy) const {
        if (steep) { ox = y; oy = x; } else { ox = x; oy = y; }
    }
    void next() {
        ++x;
        error -= deltay;
        if (error < 0) {
            y += ystep;
            error += deltax;
        }
    }
    void reset() {
        error = deltax / 2;
        x = x0;
        y = y0;
    }
    void initialize(int &ax0, int &ay0, int &ax1, int &ay1) {
        steep = std::abs((long long)ay1 - ay0) > std::abs((long long)ax1 - ax0);
        if (steep) { std::swap(ax0, ay0); std::swap(ax1, ay1); }
        if (ax0 > ax1) { std::swap(ax0, ax1); std::swap(ay0, ay1); }
        deltax = wrap_sub(ax1, ax0);
        deltay = std::abs(wrap_sub(ay1, ay0));
        ystep = (ay0 < ay1 ? 1 : -1);
    }
};

// ---------------------------------------------------------------------------------
struct Params {
    double min_depth, max_depth;
    int32_t num_levels;
    double image_scale;
    int32_t radius, weight_kind, cost_kind, depth_kind, select_kind;
    double second_best_factor, ncc_threshold;
    int32_t keep_cost_volume, row_begin, row_end;
};

double depthFromLabel(const Params &P, int label) {
    double t = label / (P.num_levels - 1.0);
    if (P.depth_kind == 1) t /= (5 - 4 * t);  // twoviewstereo.cpp:981-985
    return P.min_depth * (1 - t) + P.max_depth * t;  // multiviewstereo.cpp:733-736
}

// multiviewstereo.cpp:740-750 / twoviewstereo.cpp:987-995
bool pointFromDepth(const Ray &ray, V3 normal, double depth, V3 &p) {
    Plane plane(normal, p + normal * depth);
    return intersect(ray, plane, p);
}

struct View {
    Camera cam;
    Image img;
    Mask mask;
};

struct IPoint {
    int x, y;
};

// stereo/multiviewstereo.cpp:754-810 (mvs=true: clipped iterator + consecutive-duplicate removal)
// stereo/twoviewstereo.cpp:999-1054 (mvs=false: unclipped, duplicates kept)
void epipolarCurve(const Params &P, const Ray &ray, V3 cameraOffset, V3 depthPlaneNormal,
                   const Mask &mask, const Camera &view, bool mvs, int rootMode,
                   std::vector<IPoint> &curve) {
    curve.clear();
    double x1 = NaN, y1 = NaN;
    for (int d = 0; d < P.num_levels; ++d) {
        V3 point = cameraOffset;
        const double depth = depthFromLabel(P, d);
        if (pointFromDepth(ray, depthPlaneNormal, depth, point)) {
            if (view.project(point, rootMode)) {
                const double x2 = point.x * P.image_scale;
                const double y2 = point.y * P.image_scale;
                if (std::isnan(x1)) {
                    x1 = x2;
                    y1 = y2;
                } else {
                    const double dx = x2 - x1, dy = y2 - y1;
                    if (dx * dx + dy * dy >= 1) {
                        LineIterator iter = mvs ? LineIterator(to_int_x86(x1), to_int_x86(y1), to_int_x86(x2),
                                                               to_int_x86(y2), mask.w, mask.h)
                                                : LineIterator(to_int_x86(x1), to_int_x86(y1), to_int_x86(x2),
                                                               to_int_x86(y2));
                        long long guard = 0;
                        while (iter.hasNext()) {
                            int tx, ty;
                            iter.current(tx, ty);
                            if (mask.white(tx, ty)) curve.push_back({tx, ty});
                            iter.next();
                            if (++guard > (1LL << 22)) break;  // absurd unclipped segments
                        }
                        x1 = x2;
                        y1 = y2;
                    }
                }
            }
        }
    }
    if (mvs) {  // multiviewstereo.cpp:800-807: remove consecutive duplicates
        curve.erase(std::unique(curve.begin(), curve.end(),
                                [](const IPoint &a, const IPoint &b) { return a.x == b.x && a.y == b.y; }),
                    curve.end());
    }
}

int rowBegin(const Params &P) { return P.row_begin > 0 ? P.row_begin : 0; }
int rowEnd(const Params &P, int h) { return (P.row_end > 0 && P.row_end < h) ? P.row_end : h; }

const int IDX_NONE = -1, IDX_MASKED = -2, IDX_REJECTED = -3;

double evalCost(const Params &P, const View &A, const View &B, int x, int y, int x2, int y2,
                const WeightFunc &wf) {
    switch (P.cost_kind) {
        case 0: return cost_ncc_two(A.img, B.img, A.mask, B.mask, x, y, x2, y2, wf);
        case 1: return cost_ncc_mvs(A.img, B.img, x, y, x2, y2, wf);
        default: return cost_sad_two(A.img, B.img, A.mask, B.mask, x, y, x2, y2, wf);
    }
}

// Two-view, one direction, LABEL mode: stereo/twoviewstereo.cpp:265-276 (per-pixel setup) +
// :308-329 (label sweep) with the WTA rule of :320-325 and the ratio test of :304-305.
// volume (optional) is [row-row_begin][x][d], NaN where the label could not be evaluated.
void twoviewLabel(const Params &P, const View &A, const View &B, int rootMode, double *depthOut,
                  int32_t *indexOut, double *bestOut, double *volume) {
    const int w = A.img.w, h = A.img.h, D = P.num_levels;
    const V3 cameraC = A.cam.Cv();
    const V3 depthPlaneNormal = A.cam.prin();
    const int r0 = rowBegin(P), r1 = rowEnd(P, h);
#pragma omp parallel
    {
        WeightFunc wf;
        wf.initialize(P.weight_kind, P.radius);
#pragma omp for schedule(static)
        for (int y = r0; y < r1; ++y) {
            for (int x = 0; x < w; ++x) {
                const size_t pv = (size_t)y * w + x;
                depthOut[pv] = NaN;
                indexOut[pv] = IDX_MASKED;
                if (bestOut) bestOut[pv] = NaN;
                double *vol = volume ? volume + ((size_t)(y - r0) * w + x) * D : nullptr;
                if (vol)
                    for (int d = 0; d < D; ++d) vol[d] = NaN;
                if (!A.mask.white(x, y)) continue;
                indexOut[pv] = IDX_NONE;
                wf.init_weights(A.img, x, y);
                Ray ray = A.cam.unproject((x + 0.5) / P.image_scale, (y + 0.5) / P.image_scale);
                double secondBestCost = INF, minCost = INF;
                for (int d = 0; d < D; ++d) {
                    V3 point = cameraC;
                    const double depth = depthFromLabel(P, d);
                    if (pointFromDepth(ray, depthPlaneNormal, depth, point)) {
                        if (B.cam.project(point, rootMode)) {
                            double x2 = point.x * P.image_scale - 0.5;
                            double y2 = point.y * P.image_scale - 0.5;
                            double cost = evalCost(P, A, B, x, y, to_int_x86(x2), to_int_x86(y2), wf);
                            if (vol) vol[d] = cost;
                            if (cost + 1e-10 < minCost) {
                                secondBestCost = minCost;
                                minCost = cost;
                                depthOut[pv] = depth;
                                indexOut[pv] = d;
                            }
                        }
                    }
                }
                if (bestOut) bestOut[pv] = minCost;
                if (P.second_best_factor > 0 && minCost > P.second_best_factor * secondBestCost) {
                    depthOut[pv] = INF;
                    indexOut[pv] = IDX_REJECTED;
                }
            }
        }
    }
}

// Two-view, one direction, CURVE mode (the reference's live path):
// stereo/twoviewstereo.cpp:265-305.
void twoviewCurve(const Params &P, const View &A, const View &B, int rootMode, double *depthOut,
                  double *bestOut, int32_t *countOut) {
    const int w = A.img.w, h = A.img.h;
    const V3 cameraC = A.cam.Cv();
    const V3 depthPlaneNormal = A.cam.prin();
    const int r0 = rowBegin(P), r1 = rowEnd(P, h);
#pragma omp parallel
    {
        WeightFunc wf;
        wf.initialize(P.weight_kind, P.radius);
        std::vector<IPoint> curve;
#pragma omp for schedule(static)
        for (int y = r0; y < r1; ++y) {
            for (int x = 0; x < w; ++x) {
                const size_t pv = (size_t)y * w + x;
                depthOut[pv] = NaN;
                if (bestOut) bestOut[pv] = NaN;
                if (countOut) countOut[pv] = 0;
                if (!A.mask.white(x, y)) continue;
                wf.init_weights(A.img, x, y);
                Ray ray = A.cam.unproject((x + 0.5) / P.image_scale, (y + 0.5) / P.image_scale);
                double secondBestCost = INF, minCost = INF;
                epipolarCurve(P, ray, cameraC, depthPlaneNormal, B.mask, B.cam, false, rootMode, curve);
                if (countOut) countOut[pv] = (int32_t)curve.size();
                for (const IPoint &p : curve) {
                    Ray ray2 = B.cam.unproject((p.x + 0.5) / P.image_scale, (p.y + 0.5) / P.image_scale);
                    V3 p1, p2;
                    closestPoints(ray, ray2, p1, p2);
                    const double cost = evalCost(P, A, B, x, y, p.x, p.y, wf);
                    if (cost + 1e-10 < minCost) {
                        p1 = p1 + p2;
                        p1 = p1 * 0.5;
                        p1 = A.cam.fromGlobalToLocal(p1);
                        secondBestCost = minCost;
                        minCost = cost;
                        depthOut[pv] = p1.z;
                    }
                }
                if (bestOut) bestOut[pv] = minCost;
                if (P.second_best_factor > 0 && minCost > P.second_best_factor * secondBestCost)
                    depthOut[pv] = INF;
            }
        }
    }
}

typedef std::pair<double, double> PeakPair;  // <ncc, depth>, multiviewstereo.cpp:478

// Multi-view, one reference view.  curve=true: stereo/multiviewstereo.cpp:543-604 + WTA :654-660.
// curve=false ("label mode", our restatement over depth labels, SURVEY §8a S4 applied to S2):
// the candidates are the label projections instead of the rasterised curve pixels:
//   tap = trunc(project(point_d) * scale)  (pixel convention of :773-775,:787),
//   skipped unless the neighbour mask is WHITE there (:787), depth = depthFromLabel(d).
// volume (optional, label mode) is [nbr][row-row_begin][x][d] (NaN = not evaluated).
void mvsView(const Params &P, const std::vector<View> &views, int ref, const int32_t *nbrs, int nn,
             bool curveMode, int rootMode, double *depthOut, int32_t *indexOut, double *bestOut,
             double *volume, double *peaksOut /* [h][w][9][2] or null */) {
    const int K = 9;
    const View &A = views[ref];
    const int w = A.img.w, h = A.img.h, D = P.num_levels;
    const V3 cameraC = A.cam.Cv();
    const V3 depthPlaneNormal = A.cam.prin();
    const int r0 = rowBegin(P), r1 = rowEnd(P, h);
    const size_t volStride = (size_t)(r1 - r0) * w * D;
#pragma omp parallel
    {
        WeightFunc wf;
        wf.initialize(P.weight_kind, P.radius);
        std::vector<IPoint> curve;
        std::vector<PeakPair> peaks;
        std::vector<int> peakLabel;
#pragma omp for schedule(static)
        for (int y = r0; y < r1; ++y) {
            for (int x = 0; x < w; ++x) {
                const size_t pv = (size_t)y * w + x;
                depthOut[pv] = INF;  // multiviewstereo.cpp:559
                if (indexOut) indexOut[pv] = IDX_MASKED;
                if (bestOut) bestOut[pv] = NaN;
                if (volume)
                    for (int j = 0; j < nn; ++j) {
                        double *vol = volume + j * volStride + ((size_t)(y - r0) * w + x) * D;
                        for (int d = 0; d < D; ++d) vol[d] = NaN;
                    }
                peaks.assign(K, PeakPair(0, -1));
                if (!A.mask.white(x, y)) continue;
                wf.init_weights(A.img, x, y);
                Ray ray = A.cam.unproject((x + 0.5) / P.image_scale, (y + 0.5) / P.image_scale);
                int bestLabel = IDX_NONE;
                PeakPair bestPair(0, -1);
                for (int j = 0; j < nn; ++j) {
                    const View &B = views[nbrs[j]];
                    if (curveMode) {
                        epipolarCurve(P, ray, cameraC, depthPlaneNormal, B.mask, B.cam, true, rootMode, curve);
                        for (const IPoint &p : curve) {
                            Ray ray2 = B.cam.unproject((p.x + 0.5) / P.image_scale, (p.y + 0.5) / P.image_scale);
                            V3 p1, p2;
                            closestPoints(ray, ray2, p1, p2);
                            const double cost = cost_ncc_mvs(A.img, B.img, x, y, p.x, p.y, wf);
                            if (cost > P.ncc_threshold) {
                                p1 = p1 + p2;
                                p1 = p1 * 0.5;
                                p1 = A.cam.fromGlobalToLocal(p1);
                                peaks.push_back(PeakPair(cost, p1.z));
                            }
                        }
                    } else {
                        double *vol = volume ? volume + j * volStride + ((size_t)(y - r0) * w + x) * D : nullptr;
                        for (int d = 0; d < D; ++d) {
                            V3 point = cameraC;
                            const double depth = depthFromLabel(P, d);
                            if (pointFromDepth(ray, depthPlaneNormal, depth, point)) {
                                if (B.cam.project(point, rootMode)) {
                                    const int tx = to_int_x86(point.x * P.image_scale);
                                    const int ty = to_int_x86(point.y * P.image_scale);
                                    if (!B.mask.white(tx, ty)) continue;
                                    const double cost = cost_ncc_mvs(A.img, B.img, x, y, tx, ty, wf);
                                    if (vol) vol[d] = cost;
                                    if (cost > P.ncc_threshold) {
                                        PeakPair pp(cost, depth);
                                        peaks.push_back(pp);
                                        if (pp > bestPair || bestLabel == IDX_NONE) {
                                            bestPair = pp;
                                            bestLabel = d;
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
                std::sort(peaks.begin(), peaks.end());
                peaks = std::vector<PeakPair>(peaks.end() - K, peaks.end());
                if (peaksOut)
                    for (int k = 0; k < K; ++k) {
                        peaksOut[(pv * K + k) * 2 + 0] = peaks[k].first;
                        peaksOut[(pv * K + k) * 2 + 1] = peaks[k].second;
                    }
                depthOut[pv] = peaks.back().second;  // multiviewstereo.cpp:654-660
                if (bestOut) bestOut[pv] = peaks.back().first;
                if (indexOut) indexOut[pv] = bestLabel;
            }
        }
    }
}

// stereo/twoviewstereo.cpp:596-672, one direction (A checked against B).  Reads depthB as it is
// at call time: the reference runs left first (in place) and then right against the UPDATED left.
void crossCheckTwoDir(const Params &P, const View &A, const View &B, double *depthA,
                      const double *depthB, double thresh, int rootMode) {
    const int w = A.img.w, h = A.img.h, w2 = B.img.w, h2 = B.img.h;
    const V3 nA = A.cam.prin(), nB = B.cam.prin();
    const double s = P.image_scale;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            double &depth = depthA[(size_t)y * w + x];
            if (!std::isfinite(depth)) continue;
            Ray ray = A.cam.unproject((x + 0.5) / s, (y + 0.5) / s);
            V3 p1 = A.cam.Cv();
            if (pointFromDepth(ray, nA, depth, p1)) {
                V3 q = p1;
                if (B.cam.project(q, rootMode)) {
                    double x2 = q.x * s, y2 = q.y * s;
                    if (x2 >= 0 && y2 >= 0 && x2 < w2 && y2 < h2) {
                        const double odepth = depthB[(size_t)((int)y2) * w2 + (int)x2];
                        if (std::isfinite(odepth)) {
                            Ray ray2 = B.cam.unproject((x2 + 0.5) / s, (y2 + 0.5) / s);
                            V3 p2 = B.cam.Cv();
                            if (pointFromDepth(ray2, nB, odepth, p2)) {
                                const double nrm = norm(p1 - p2);
                                if (!std::isfinite(nrm) || nrm > thresh) depth = INF;
                            } else depth = INF;
                        } else depth = INF;
                    } else depth = INF;
                } else depth = INF;
            }
        }
}

// stereo/multiviewstereo.cpp:666-729 for one view, in place; reads the other views' CURRENT depths
// (the reference cross-checks views in index order, so later views see earlier views' NaNs).
void crossCheckMvs(const Params &P, const std::vector<View> &views, std::vector<double *> &depths,
                   int vi, double thresh, int rootMode) {
    const View &A = views[vi];
    const int w = A.img.w, h = A.img.h;
    const V3 viewNormal = A.cam.prin();
    const double s = P.image_scale;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            double &depth = depths[vi][(size_t)y * w + x];
            if (!std::isfinite(depth)) continue;
            Ray ray = A.cam.unproject((x + 0.5) / s, (y + 0.5) / s);
            V3 p1 = A.cam.Cv();
            if (pointFromDepth(ray, viewNormal, depth, p1)) {
                bool found = false;
                for (size_t v2 = 0; v2 < views.size(); ++v2) {
                    if ((int)v2 == vi) continue;
                    const View &B = views[v2];
                    V3 q = p1;
                    if (B.cam.project(q, rootMode)) {
                        double x2 = q.x * s, y2 = q.y * s;
                        if (x2 >= 0 && y2 >= 0 && x2 < B.img.w && y2 < B.img.h) {
                            const double odepth = depths[v2][(size_t)((int)y2) * B.img.w + (int)x2];
                            if (std::isfinite(odepth)) {
                                Ray ray2 = B.cam.unproject((x2 + 0.5) / s, (y2 + 0.5) / s);
                                V3 p2 = B.cam.Cv();
                                if (pointFromDepth(ray2, B.cam.prin(), odepth, p2)) {
                                    const double nrm = norm(p1 - p2);
                                    if (std::isfinite(nrm) && nrm < thresh) {
                                        found = true;
                                        break;
                                    }
                                }
                            }
                        }
                    }
                }
                if (!found) depth = NaN;
            }
        }
}

void makeViews(int V, const Camera *cams, const uint8_t *const *rgba8, const uint8_t *const *mask8,
               int w, int h, std::vector<View> &views) {
    views.resize(V);
    for (int i = 0; i < V; ++i) {
        views[i].cam = cams[i];
        views[i].img.fromRGBA8(rgba8[i], w, h);
        views[i].mask.set(mask8 ? mask8[i] : nullptr, w, h);
    }
}

}  // namespace

// =================================================================================
// C ABI for ctypes (tests/, bench.py cpu_baseline).  Prefix orc_.
extern "C" {

struct orc_scene {
    std::vector<View> views;
};

orc_scene *orc_scene_create(int V, const Camera *cams, const uint8_t *const *rgba8,
                            const uint8_t *const *mask8, int w, int h) {
    orc_scene *s = new orc_scene;
    makeViews(V, cams, rgba8, mask8, w, h, s->views);
    return s;
}
void orc_scene_destroy(orc_scene *s) { delete s; }

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void orc_stats_reset() { g_stats = Stats(); }
void orc_stats_get(long long *out5, double *maxdiff) {
    out5[0] = g_stats.project_calls;
    out5[1] = g_stats.quartic_fail;
    out5[2] = g_stats.root_mismatch;
    out5[3] = g_stats.no_root;
    out5[4] = 0;
    *maxdiff = g_stats.max_root_diff;
}

// geometry leaves --------------------------------------------------------------
void orc_unproject(const Camera *cam, double px, double py, double *out6) {
    Ray r = cam->unproject(px, py);
    out6[0] = r.src.x; out6[1] = r.src.y; out6[2] = r.src.z;
    out6[3] = r.dir.x; out6[4] = r.dir.y; out6[5] = r.dir.z;
}
void orc_unproject_grid(const Camera *cam, int w, int h, double scale, double *out) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            orc_unproject(cam, (x + 0.5) / scale, (y + 0.5) / scale, out + ((size_t)y * w + x) * 6);
}
// stereo/refractioncalibration.cpp:175-201 (RefractiveCalibrationFunction::diff): distance of the
// two unprojected rays, scaled by 0.5*fx/z in each view to approximate an image-space distance.
void orc_calibration_residuals(const Camera *cams, int n, const int32_t *pairs, const double *pix, double *out) {
    for (int i = 0; i < n; ++i) {
        const Camera &v1 = cams[pairs[2 * i]], &v2 = cams[pairs[2 * i + 1]];
        Ray R1 = v1.unproject(pix[4 * i], pix[4 * i + 1]);
        Ray R2 = v2.unproject(pix[4 * i + 2], pix[4 * i + 3]);
        V3 p1, p2;
        closestPoints(R1, R2, p1, p2);
        V3 df = p1 - p2;
        const double dist = std::sqrt(dot(df, df));
        V3 mid = (p1 + p2) * 0.5;
        V3 mid1 = v1.fromGlobalToLocal(mid), mid2 = v2.fromGlobalToLocal(mid);
        const double e1 = (0.5 * v1.K[0] * dist) / mid1.z;
        const double e2 = (0.5 * v2.K[0] * dist) / mid2.z;
        out[i] = e1 + e2;
    }
}
// ---- interface calibration: util/lm.cpp:59-150 around stereo/refractioncalibration.cpp:127-253 ----
// A literal restatement of the reference's loops (diff and gradient are called point by point and
// parameter by parameter, the cameras are re-configured by update() for every finite difference),
// with the two deviations documented in include/util/lm.hpp: fixed parameters are removed from
// the linear system, and the sense of the solve test (:104) is selectable (literalCheck != 0 is
// the text as written).  Model layout: [n, (px, py, dist) x V].
namespace {
struct CalibFunction {
    Camera *views;
    int V;
    const int32_t *p2c;
    const double *pix;
    int exactAttribution;

    bool update(const double *model) {  // refractioncalibration.cpp:238-251
        for (int v = 0; v < V; ++v)
            if (model[3 * v + 2] < 1e-4) return false;
        for (int v = 0; v < V; ++v) {
            Camera &view = views[v];
            V3 normal = mul(view.Kinv, V3{model[3 * v + 1], model[3 * v + 2], 1.0});
            normal = normalized(normal);
            if (std::fabs(model[0] - view.n) > 1e-10) {  // Camera::setRefractiveIndex, camera.cpp:337-342
                view.n = model[0];
                view.is_refractive = (!iszero(view.n - 1) && !iszero(view.plane_d));
            }
            const V3 dn = normal - V3{view.plane_n[0], view.plane_n[1], view.plane_n[2]};
            if (!(dot(dn, dn) + std::fabs(model[3 * v + 3] - view.plane_d) < 1e-10)) {  // plane != P: setPlane, camera.cpp:326-332
                view.plane_n[0] = normal.x;
                view.plane_n[1] = normal.y;
                view.plane_n[2] = normal.z;
                view.plane_d = model[3 * v + 3];
                view.is_refractive = (!iszero(view.n - 1) && !iszero(view.plane_d));
            }
        }
        return true;
    }
    double diff(int i) const {  // :170-201
        double out;
        const int32_t pair[2] = {p2c[2 * i], p2c[2 * i + 1]};
        orc_calibration_residuals(views, 1, pair, pix + 4 * i, &out);
        return out;
    }
    bool touches(int paramIndex, int i) const {
        if (exactAttribution) {
            if (paramIndex == 0) return true;
            const int v = (paramIndex - 1) / 3;
            return v == p2c[2 * i] || v == p2c[2 * i + 1];
        }
        return paramIndex / 3 == p2c[2 * i] || paramIndex / 3 == p2c[2 * i + 1];  // :205-207
    }
    double gradient(int i, const std::vector<double> &model, int paramIndex) {  // :203-236
        if (!touches(paramIndex, i)) return 0.0;
        std::vector<double> m1 = model, m2 = model;
        if (paramIndex == 0) {
            m1[paramIndex] = model[paramIndex] - 0.01;
            m2[paramIndex] = model[paramIndex] + 0.01;
        } else if ((paramIndex - 1) % 3 == 0) {
            m1[paramIndex] = model[paramIndex] - 0.5;
            m2[paramIndex] = model[paramIndex] + 0.5;
        } else if ((paramIndex - 1) % 3 == 1) {
            m1[paramIndex] = model[paramIndex] - 0.1;
            m2[paramIndex] = model[paramIndex] + 0.1;
        } else {
            m1[paramIndex] = model[paramIndex];
            m2[paramIndex] = model[paramIndex] + 0.0001;
        }
        update(m1.data());
        const double val1 = diff(i);
        update(m2.data());
        const double val2 = diff(i);
        update(model.data());
        return (val2 - val1) / (m2[paramIndex] - m1[paramIndex]);
    }
    double chiSquared(int n) const {  // lm.cpp:50-55
        double sum = 0.0;
        for (int i = 0; i < n; ++i) {
            const double d = diff(i);
            sum += d * d;
        }
        return sum;
    }
};

// LU with partial pivoting + substitution (Eigen's PartialPivLU in the reference, lm.cpp:101)
void luSolveDense(std::vector<double> M, std::vector<double> y, std::vector<double> &x, int n) {
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
}  // namespace

// Returns the number of iterations; model (1 + 3V) is updated in place, chi2[0] / chi2[1] = the error
// before / after.  cams is modified while running and restored (planes AND index) before returning.
int orc_calibration_lm(const Camera *cams_in, int V, int n, const int32_t *pairs, const double *pix, double *model_io,
                       const uint8_t *fixed, int maxIterations, double epsilon, int literalCheck, int exactAttribution,
                       double *chi2) {
    std::vector<Camera> cams(cams_in, cams_in + V);
    CalibFunction f{cams.data(), V, pairs, pix, exactAttribution};
    const int nparms = 1 + 3 * V;
    std::vector<double> model(model_io, model_io + nparms);
    std::vector<int> free_;
    for (int p = 0; p < nparms; ++p)
        if (!fixed[p]) free_.push_back(p);
    const int nf = (int)free_.size();
    f.update(model.data());
    double e0 = f.chiSquared(n);
    chi2[0] = chi2[1] = e0;
    if (nf == 0 || n <= 0) return 0;
    double lambda = 1;
    int iter = 0, term = 0;
    std::vector<double> H((size_t)nf * nf), g(nf), step;
    do {
        std::fill(H.begin(), H.end(), 0.0);
        std::fill(g.begin(), g.end(), 0.0);
        for (int i = 0; i < n; ++i) {  // lm.cpp:83-94
            const double diff = f.diff(i);
            for (int a = 0; a < nf; ++a) {
                const double gradr = f.gradient(i, model, free_[a]);
                for (int b = 0; b < nf; ++b) {
                    const double gradc = f.gradient(i, model, free_[b]);
                    H[(size_t)a * nf + b] += gradr * gradc;
                }
                g[a] += diff * gradr;
            }
        }
        for (int a = 0; a < nf; ++a) H[(size_t)a * nf + a] *= 1.0 + lambda;
        std::vector<double> rhs(nf);
        for (int a = 0; a < nf; ++a) rhs[a] = -g[a];
        luSolveDense(H, rhs, step, nf);
        double d2 = 0.0, a2 = 0.0, b2 = 0.0;  // (H*new_model).isApprox(-g, 1e-10)
        for (int r = 0; r < nf; ++r) {
            double acc = 0.0;
            for (int c = 0; c < nf; ++c) acc += H[(size_t)r * nf + c] * step[c];
            d2 += (acc - rhs[r]) * (acc - rhs[r]);
            a2 += acc * acc;
            b2 += rhs[r] * rhs[r];
        }
        const bool solved = d2 <= 1e-20 * std::min(a2, b2);
        if (literalCheck ? solved : !solved) {
            ++term;
            continue;
        }
        bool bad_model = false;
        for (int p = 0; p < nparms; ++p)
            if (std::isnan(model[p])) {
                bad_model = true;
                lambda *= 10.0;
                ++term;
            }
        if (bad_model) continue;
        std::vector<double> new_model = model;
        for (int a = 0; a < nf; ++a) new_model[free_[a]] += step[a];
        if (!f.update(new_model.data())) {
            f.update(model.data());
            lambda *= 10.0;
            ++term;
            continue;
        }
        const double e1 = f.chiSquared(n);
        if (std::fabs(e1 - e0) > epsilon) term = 0;
        else ++term;
        const bool worse = (e0 - e1 < 0);
        if (worse || std::isnan(e1)) {
            lambda *= 10.0;
            f.update(model.data());
        } else {
            lambda *= 0.1;
            e0 = e1;
            model = new_model;
        }
    } while (++iter < maxIterations && term < 5);
    f.update(model.data());
    chi2[1] = f.chiSquared(n);
    for (int p = 0; p < nparms; ++p) model_io[p] = model[p];
    return iter;
}
int orc_project(const Camera *cam, const double *xyz, int rootMode, double *out2) {
    V3 p = {xyz[0], xyz[1], xyz[2]};
    bool ok = cam->project(p, rootMode);
    out2[0] = p.x;
    out2[1] = p.y;
    return ok ? 1 : 0;
}
void orc_project_points(const Camera *cam, int n, const double *xyz, int rootMode, double *out_xy,
                        int32_t *out_ok) {
    for (int i = 0; i < n; ++i) out_ok[i] = orc_project(cam, xyz + 3 * i, rootMode, out_xy + 2 * i);
}
int orc_poly_roots4(const double *coeffs5, double *re, double *im) {
    return poly_roots4(coeffs5, re, im) ? 1 : 0;
}
double orc_snell_root(double r, double d, double h, double n) { return snellRoot(r, d, h, n); }
int orc_intersect(const double *src, const double *dir, const double *pn, double pd, double *out3) {
    Ray R(V3{src[0], src[1], src[2]}, V3{dir[0], dir[1], dir[2]});
    Plane P(V3{pn[0], pn[1], pn[2]}, pd);
    V3 p{NaN, NaN, NaN};
    bool ok = intersect(R, P, p);
    out3[0] = p.x; out3[1] = p.y; out3[2] = p.z;
    return ok;
}
int orc_refract(const double *src, const double *dir, const double *pn, double pd, double n, double *out6) {
    Ray R(V3{src[0], src[1], src[2]}, V3{dir[0], dir[1], dir[2]});
    Plane P(V3{pn[0], pn[1], pn[2]}, pd);
    Ray O = R;
    bool ok = refract(R, P, n, O);
    out6[0] = O.src.x; out6[1] = O.src.y; out6[2] = O.src.z;
    out6[3] = O.dir.x; out6[4] = O.dir.y; out6[5] = O.dir.z;
    return ok;
}
void orc_closest_points(const double *s1, const double *d1, const double *s2, const double *d2, double *out6) {
    Ray A(V3{s1[0], s1[1], s1[2]}, V3{d1[0], d1[1], d1[2]});
    Ray B(V3{s2[0], s2[1], s2[2]}, V3{d2[0], d2[1], d2[2]});
    V3 p1, p2;
    closestPoints(A, B, p1, p2);
    out6[0] = p1.x; out6[1] = p1.y; out6[2] = p1.z;
    out6[3] = p2.x; out6[4] = p2.y; out6[5] = p2.z;
}

// images / weights / costs ------------------------------------------------------
void orc_sample(const orc_scene *s, int view, double x, double y, double *out4) {
    RGBA r = s->views[view].img.sample(x, y);
    out4[0] = r.r; out4[1] = r.g; out4[2] = r.b; out4[3] = r.a;
}
void orc_sample_int(const orc_scene *s, int view, int x, int y, double *out4) {
    RGBA r = s->views[view].img.sampleInt(x, y);
    out4[0] = r.r; out4[1] = r.g; out4[2] = r.b; out4[3] = r.a;
}
void orc_weights(const orc_scene *s, int view, int kind, int radius, int n, const int32_t *cx,
                 const int32_t *cy, double *out) {
    const size_t wn = (size_t)(2 * radius + 1) * (2 * radius + 1);
#pragma omp parallel
    {
        WeightFunc wf;
        wf.initialize(kind, radius);
#pragma omp for schedule(static)
        for (int i = 0; i < n; ++i) {
            wf.init_weights(s->views[view].img, cx[i], cy[i]);
            std::memcpy(out + i * wn, wf.weights.data(), wn * sizeof(double));
        }
    }
}
double orc_cost(const orc_scene *s, const Params *P, int va, int vb, int x1, int y1, int x2, int y2) {
    WeightFunc wf;
    wf.initialize(P->weight_kind, P->radius);
    wf.init_weights(s->views[va].img, x1, y1);
    return evalCost(*P, s->views[va], s->views[vb], x1, y1, x2, y2, wf);
}
double orc_depth_from_label(const Params *P, int label) { return depthFromLabel(*P, label); }

// rasteriser ---------------------------------------------------------------------
int orc_line(int x0, int y0, int x1, int y1, int clip, int w, int h, int32_t *out_xy, int max_pts) {
    LineIterator it = clip ? LineIterator(x0, y0, x1, y1, w, h) : LineIterator(x0, y0, x1, y1);
    int n = 0;
    while (it.hasNext()) {
        int tx, ty;
        it.current(tx, ty);
        if (n < max_pts) {
            out_xy[2 * n] = tx;
            out_xy[2 * n + 1] = ty;
        }
        ++n;
        it.next();
        if (n > (1 << 24)) break;
    }
    return n;
}
int orc_clip_line(int32_t *xyxy, int w, int h) {
    int x0 = xyxy[0], y0 = xyxy[1], x1 = xyxy[2], y1 = xyxy[3];
    bool ok = clipLine(x0, y0, x1, y1, w, h);
    xyxy[0] = x0; xyxy[1] = y0; xyxy[2] = x1; xyxy[3] = y1;
    return ok;
}
int orc_epipolar_curve(const orc_scene *s, const Params *P, int ref, int nbr, int x, int y, int mvs,
                       int rootMode, int32_t *out_xy, int max_pts) {
    const View &A = s->views[ref];
    const View &B = s->views[nbr];
    Ray ray = A.cam.unproject((x + 0.5) / P->image_scale, (y + 0.5) / P->image_scale);
    std::vector<IPoint> curve;
    epipolarCurve(*P, ray, A.cam.Cv(), A.cam.prin(), B.mask, B.cam, mvs != 0, rootMode, curve);
    for (size_t i = 0; i < curve.size() && (int)i < max_pts; ++i) {
        out_xy[2 * i] = curve[i].x;
        out_xy[2 * i + 1] = curve[i].y;
    }
    return (int)curve.size();
}

// the path ------------------------------------------------------------------------
void orc_twoview_label(const orc_scene *s, const Params *P, int a, int b, int rootMode, double *depth,
                       int32_t *index, double *best, double *volume) {
    twoviewLabel(*P, s->views[a], s->views[b], rootMode, depth, index, best, volume);
}
void orc_twoview_curve(const orc_scene *s, const Params *P, int a, int b, int rootMode, double *depth,
                       double *best, int32_t *count) {
    twoviewCurve(*P, s->views[a], s->views[b], rootMode, depth, best, count);
}
void orc_mvs_view(const orc_scene *s, const Params *P, int ref, const int32_t *nbrs, int nn, int curveMode,
                  int rootMode, double *depth, int32_t *index, double *best, double *volume, double *peaks) {
    mvsView(*P, s->views, ref, nbrs, nn, curveMode != 0, rootMode, depth, index, best, volume, peaks);
}
// stereo/multiviewstereo.cpp:335-360
void orc_select_neighbours(const orc_scene *s, int maxN, int32_t *out, int32_t *counts) {
    const int V = (int)s->views.size();
    for (int i = 0; i < V; ++i) {
        std::vector<std::pair<double, size_t>> nearViews;
        const Camera &c1 = s->views[i].cam;
        for (int j = 0; j < V; ++j)
            if (i != j) {
                const Camera &c2 = s->views[j].cam;
                if (std::fabs(dot(c1.prin(), c2.prin())) > 0.2) {
                    V3 dc = c1.Cv() - c2.Cv();
                    nearViews.push_back(std::make_pair(dot(dc, dc), (size_t)j));
                }
            }
        size_t end = nearViews.size();
        if ((size_t)maxN < nearViews.size()) {
            std::sort(nearViews.begin(), nearViews.end());
            end = maxN;
        }
        counts[i] = (int32_t)end;
        for (size_t k = 0; k < end; ++k) out[i * maxN + k] = (int32_t)nearViews[k].second;
    }
}
// two-view cross-check, both directions, in place (twoviewstereo.cpp:596-672)
void orc_crosscheck_two(const orc_scene *s, const Params *P, int l, int r, double *depthL, double *depthR,
                        double thresh, int rootMode) {
    crossCheckTwoDir(*P, s->views[l], s->views[r], depthL, depthR, thresh, rootMode);
    crossCheckTwoDir(*P, s->views[r], s->views[l], depthR, depthL, thresh, rootMode);
}
// MVS cross-check, all views in index order, in place (multiviewstereo.cpp:427-431,666-729).
// snapshot != 0: every view is checked against the PRE-cross-check depths of the others (an
// order-independent variant kept for comparison; the GPU path uses the in-order semantics).
void orc_crosscheck_mvs(const orc_scene *s, const Params *P, double *const *depths, double thresh,
                        int rootMode, int snapshot) {
    const int V = (int)s->views.size();
    std::vector<double *> cur(V);
    std::vector<std::vector<double>> snap;
    if (snapshot) {
        snap.resize(V);
        for (int v = 0; v < V; ++v) {
            const size_t n = (size_t)s->views[v].img.w * s->views[v].img.h;
            snap[v].assign(depths[v], depths[v] + n);
        }
    }
    for (int v = 0; v < V; ++v) {
        for (int u = 0; u < V; ++u) cur[u] = (snapshot && u != v) ? snap[u].data() : depths[u];
        crossCheckMvs(*P, s->views, cur, v, thresh, rootMode);
    }
}

}  // extern "C"
