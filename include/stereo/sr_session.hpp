// stereo/sr_session.hpp — RAII wrapper of an sr_ctx shared by the stereo class API.
// Every numeric result of the classes in include/stereo comes from the C ABI (CUDA); this file
// only checks status codes.  No CPU fallback: construction throws when no CUDA device exists.
#ifndef SR_STEREO_SESSION_HPP
#define SR_STEREO_SESSION_HPP
#include "util/precompiled.hpp"
#include "sr_b200.h"
namespace sr_host {
class Session {
public:
    explicit Session(int device = 0) : ctx_(nullptr) {
        if (sr_ctx_create(device, &ctx_) != SR_OK) throw std::runtime_error(std::string("sr_ctx_create: ") + sr_last_error(nullptr));
    }
    ~Session() { sr_ctx_destroy(ctx_); }
    Session(const Session &) = delete;
    Session &operator=(const Session &) = delete;
    sr_ctx *get() const { return ctx_; }
    void check(int rc, const char *what) const {
        if (rc != SR_OK) throw std::runtime_error(std::string(what) + ": " + sr_last_error(ctx_));
    }
private:
    sr_ctx *ctx_;
};
}  // namespace sr_host
#endif
