// util/vectorimage.hpp — RGBA / VectorImage with the reference's interface
// (util/vectorimage.hpp:28-124) and Image8, the Qt-free stand-in for QImage (w, h, RGBA8 rows;
// PNG / PPM loading through zlib).  Define HAVE_QT to use the real QImage instead.
#ifndef SR_UTIL_VECTORIMAGE_HPP
#define SR_UTIL_VECTORIMAGE_HPP
#include "util/precompiled.hpp"
#include <cstdio>
#include <cstring>
#ifdef HAVE_QT
#include <QImage>
#else
#include <zlib.h>
#endif

struct RGBA {
    double r, g, b, a;
    RGBA() : r(0), g(0), b(0), a(255) {}
    RGBA(double c) : r(c), g(c), b(c), a(255) {}
    RGBA(double r, double g, double b) : r(r), g(g), b(b), a(255) {}
    RGBA(double r, double g, double b, double a) : r(r), g(g), b(b), a(a) {}
    RGBA &operator+=(const RGBA &o) { r += o.r; g += o.g; b += o.b; return *this; }
    RGBA &operator-=(const RGBA &o) { r -= o.r; g -= o.g; b -= o.b; return *this; }
    RGBA &operator*=(double s) { r *= s; g *= s; b *= s; return *this; }
    bool isValid() const { return !(std::isnan(r) || std::isnan(g) || std::isnan(b)); }
    double toGray() const { return (0.11 * r + 0.59 * g + 0.3 * b); }
    bool operator==(const RGBA &o) const {
        return (std::fabs(r - o.r) < 1e-10 && std::fabs(g - o.g) < 1e-10 && std::fabs(b - o.b) < 1e-10 &&
                std::fabs(a - o.a) < 1e-10);
    }
    bool operator!=(const RGBA &o) const { return !operator==(o); }
};
static const RGBA BLACK(0, 0, 0);
static const RGBA WHITE(255, 255, 255);
static const RGBA INVALID(std::numeric_limits<double>::quiet_NaN(), std::numeric_limits<double>::quiet_NaN(),
                          std::numeric_limits<double>::quiet_NaN());

#ifndef HAVE_QT
// Image8: what the class API needs from QImage.  Pixels are R,G,B,A bytes.
class Image8 {
public:
    Image8() : w_(0), h_(0), alpha_(false) {}
    Image8(int w, int h) : w_(w), h_(h), alpha_(true), px_((size_t)w * h * 4, 255) {}
    explicit Image8(const std::string &file) : w_(0), h_(0), alpha_(false) { load(file); }
    Image8(const uint8_t *rgba8, int w, int h, bool hasAlpha = true)
        : w_(w), h_(h), alpha_(hasAlpha), px_(rgba8, rgba8 + (size_t)w * h * 4) {}
    bool isNull() const { return w_ <= 0 || h_ <= 0; }
    int width() const { return w_; }
    int height() const { return h_; }
    bool hasAlphaChannel() const { return alpha_; }
    const uint8_t *bits() const { return px_.data(); }
    uint8_t *bits() { return px_.data(); }
    const uint8_t *pixel(int x, int y) const { return &px_[((size_t)y * w_ + x) * 4]; }
    void setPixel(int x, int y, uint8_t r, uint8_t g, uint8_t b, uint8_t a = 255) {
        uint8_t *p = &px_[((size_t)y * w_ + x) * 4];
        p[0] = r; p[1] = g; p[2] = b; p[3] = a;
    }
    // QImage::scaledToWidth(w, Qt::SmoothTransformation) stand-in: area-weighted (box) resampling.
    // Qt's own filter is not reproducible outside Qt (SURVEY §7 hard part 6): parity runs use
    // imageScale == 1, where this is an identity copy exactly like Qt's.
    Image8 scaledToWidth(int nw) const {
        if (isNull() || nw == w_ || nw <= 0) return *this;
        const int nh = std::max(1, (int)((long long)h_ * nw / w_));
        Image8 out(nw, nh);
        out.alpha_ = alpha_;
        const double sx = (double)w_ / nw, sy = (double)h_ / nh;
        for (int y = 0; y < nh; ++y)
            for (int x = 0; x < nw; ++x) {
                double acc[4] = {0, 0, 0, 0}, wsum = 0;
                const double x0 = x * sx, x1 = (x + 1) * sx, y0 = y * sy, y1 = (y + 1) * sy;
                for (int yy = (int)y0; yy < std::min(h_, (int)std::ceil(y1)); ++yy)
                    for (int xx = (int)x0; xx < std::min(w_, (int)std::ceil(x1)); ++xx) {
                        const double wx = std::min(x1, xx + 1.0) - std::max(x0, (double)xx);
                        const double wy = std::min(y1, yy + 1.0) - std::max(y0, (double)yy);
                        const double ww = wx * wy;
                        const uint8_t *p = pixel(xx, yy);
                        for (int c = 0; c < 4; ++c) acc[c] += ww * p[c];
                        wsum += ww;
                    }
                uint8_t *q = &out.px_[((size_t)y * nw + x) * 4];
                for (int c = 0; c < 4; ++c) q[c] = (uint8_t)std::min(255.0, std::floor(acc[c] / wsum + 0.5));
            }
        return out;
    }
    // 8-bit gray / RGB / RGBA, non-interlaced PNG; binary PPM (P6).  Returns false otherwise.
    bool load(const std::string &file) {
        FILE *f = std::fopen(file.c_str(), "rb");
        if (!f) return false;
        std::vector<uint8_t> buf;
        uint8_t tmp[65536];
        size_t n;
        while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
        std::fclose(f);
        if (buf.size() > 8 && !std::memcmp(buf.data(), "\x89PNG\r\n\x1a\n", 8)) return decodePng(buf);
        if (buf.size() > 2 && buf[0] == 'P' && buf[1] == '6') return decodePpm(buf);
        return false;
    }
private:
    static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }
    bool decodePpm(const std::vector<uint8_t> &b) {
        size_t pos = 2;
        int vals[3], k = 0;
        while (k < 3 && pos < b.size()) {
            while (pos < b.size() && (b[pos] == ' ' || b[pos] == '\n' || b[pos] == '\r' || b[pos] == '\t')) ++pos;
            if (pos < b.size() && b[pos] == '#') { while (pos < b.size() && b[pos] != '\n') ++pos; continue; }
            int v = 0;
            while (pos < b.size() && b[pos] >= '0' && b[pos] <= '9') v = v * 10 + (b[pos++] - '0');
            vals[k++] = v;
        }
        ++pos;
        if (k < 3 || vals[2] != 255 || b.size() < pos + (size_t)vals[0] * vals[1] * 3) return false;
        w_ = vals[0]; h_ = vals[1]; alpha_ = false;
        px_.resize((size_t)w_ * h_ * 4);
        for (size_t i = 0; i < (size_t)w_ * h_; ++i) {
            px_[4 * i] = b[pos + 3 * i]; px_[4 * i + 1] = b[pos + 3 * i + 1]; px_[4 * i + 2] = b[pos + 3 * i + 2]; px_[4 * i + 3] = 255;
        }
        return true;
    }
    bool decodePng(const std::vector<uint8_t> &b) {
        size_t pos = 8;
        int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
        std::vector<uint8_t> idat;
        while (pos + 12 <= b.size()) {
            const uint32_t len = be32(&b[pos]);
            const char *type = (const char *)&b[pos + 4];
            if (pos + 12 + len > b.size()) return false;
            if (!std::memcmp(type, "IHDR", 4)) {
                w = (int)be32(&b[pos + 8]); h = (int)be32(&b[pos + 12]);
                depth = b[pos + 16]; ctype = b[pos + 17]; interlace = b[pos + 20];
            } else if (!std::memcmp(type, "IDAT", 4)) {
                idat.insert(idat.end(), b.begin() + pos + 8, b.begin() + pos + 8 + len);
            } else if (!std::memcmp(type, "IEND", 4)) break;
            pos += 12 + len;
        }
        const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
        if (w <= 0 || h <= 0 || depth != 8 || ch == 0 || interlace) return false;
        const size_t stride = (size_t)w * ch;
        std::vector<uint8_t> raw((stride + 1) * h);
        uLongf rawlen = raw.size();
        if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) return false;
        std::vector<uint8_t> img(stride * h);
        for (int y = 0; y < h; ++y) {  // PNG filters (None, Sub, Up, Average, Paeth)
            const uint8_t ft = raw[(stride + 1) * y];
            const uint8_t *src = &raw[(stride + 1) * y + 1];
            uint8_t *dst = &img[stride * y];
            const uint8_t *up = y ? &img[stride * (y - 1)] : nullptr;
            for (size_t i = 0; i < stride; ++i) {
                const int a = i >= (size_t)ch ? dst[i - ch] : 0, bb = up ? up[i] : 0, c = (up && i >= (size_t)ch) ? up[i - ch] : 0;
                int pred = 0;
                if (ft == 1) pred = a;
                else if (ft == 2) pred = bb;
                else if (ft == 3) pred = (a + bb) / 2;
                else if (ft == 4) {
                    const int p = a + bb - c, pa = std::abs(p - a), pb = std::abs(p - bb), pc = std::abs(p - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? bb : c);
                }
                dst[i] = (uint8_t)(src[i] + pred);
            }
        }
        w_ = w; h_ = h; alpha_ = (ch == 2 || ch == 4);
        px_.resize((size_t)w * h * 4);
        for (size_t i = 0; i < (size_t)w * h; ++i) {
            const uint8_t *p = &img[i * ch];
            uint8_t *q = &px_[4 * i];
            if (ch >= 3) { q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; q[3] = ch == 4 ? p[3] : 255; }
            else { q[0] = q[1] = q[2] = p[0]; q[3] = ch == 2 ? p[1] : 255; }
        }
        return true;
    }
    int w_, h_;
    bool alpha_;
    std::vector<uint8_t> px_;
};
typedef Image8 QImage;
#endif  // !HAVE_QT

//! A simple image class that uses a 1D vector to store pixels (util/vectorimage.hpp:86-124)
class VectorImage {
public:
    VectorImage() : w(0), h(0) {}
    VectorImage(int w, int h, const RGBA &fillVal = WHITE) : w(w), h(h), data((size_t)w * h, fillVal) {}
#ifndef HAVE_QT
    static VectorImage fromFile(const std::string &file) { return fromQImage(QImage(file)); }
    static VectorImage fromQImage(const QImage &img) {  // util/vectorimage.cpp:48-68
        VectorImage v;
        if (!img.isNull()) {
            v.w = img.width(); v.h = img.height();
            v.data.resize((size_t)v.w * v.h);
            const uint8_t *p = img.bits();
            for (size_t i = 0; i < v.data.size(); ++i) v.data[i] = RGBA(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
        }
        return v;
    }
    static QImage toQImage(const VectorImage &img, bool includeAlpha = true) {  // util/vectorimage.cpp:72-91
        if (img.w <= 0 || img.h <= 0) return QImage();
        QImage ret(img.w, img.h);
        for (int y = 0; y < img.h; ++y)
            for (int x = 0; x < img.w; ++x) {
                const RGBA &c = img.pixel(x, y);
                ret.setPixel(x, y, (uint8_t)(int)c.r, (uint8_t)(int)c.g, (uint8_t)(int)c.b, includeAlpha ? (uint8_t)(int)c.a : 255);
            }
        return ret;
    }
#endif
    bool isNull() const { return (w <= 0 || h <= 0 || data.size() == 0); }
    int width() const { return w; }
    int height() const { return h; }
    VectorImage &fill(const RGBA &rgb) { std::fill(data.begin(), data.end(), rgb); return *this; }
    void setPixel(int x, int y, const RGBA &rgb) {
        if (x < 0 || y < 0 || x >= w || y >= h) return;
        data[(size_t)y * w + x] = rgb;
    }
    const RGBA &pixel(int x, int y) const {  // util/vectorimage.cpp:115-119
        if (x < 0 || y < 0 || x >= w || y >= h) return INVALID;
        return data[(size_t)y * w + x];
    }
    RGBA sample(double x, double y) const {  // util/vectorimage.cpp:129-155
        RGBA r = INVALID;
        if (x >= 0 && y >= 0 && x + 1 < w && y + 1 < h) {
            int ix = (int)x, iy = (int)y;
            double dx = x - ix, dy = y - iy;
            r.r = r.g = r.b = 0.0;
            RGBA t = data[ix + (size_t)iy * w];           t *= (1 - dx) * (1 - dy); r += t;
            t = data[ix + (size_t)(iy + 1) * w];          t *= (1 - dx) * dy;       r += t;
            t = data[ix + 1 + (size_t)iy * w];            t *= dx * (1 - dy);       r += t;
            t = data[ix + 1 + (size_t)(iy + 1) * w];      t *= dx * dy;             r += t;
        }
        return r;
    }
    // ---- extensions used by the GPU binding -------------------------------------------------
    //! RGBA8 bytes (R,G,B,A) of the image, the layout sr_set_views takes.
    std::vector<uint8_t> toRGBA8() const {
        std::vector<uint8_t> out((size_t)w * h * 4);
        for (size_t i = 0; i < data.size(); ++i) {
            out[4 * i] = (uint8_t)(int)data[i].r; out[4 * i + 1] = (uint8_t)(int)data[i].g;
            out[4 * i + 2] = (uint8_t)(int)data[i].b; out[4 * i + 3] = (uint8_t)(int)data[i].a;
        }
        return out;
    }
    //! One byte per pixel: 255 where pixel == WHITE (the comparison every mask test uses).
    std::vector<uint8_t> toMask8() const {
        std::vector<uint8_t> out((size_t)w * h);
        for (size_t i = 0; i < data.size(); ++i) out[i] = (data[i] == WHITE) ? 255 : 0;
        return out;
    }
private:
    int w, h;
    std::vector<RGBA> data;
};
#endif
