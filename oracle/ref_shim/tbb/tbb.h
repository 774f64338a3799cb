// Intel TBB is not in this image.  The two calls the reference makes (stereo/multiviewstereo.cpp:548,
// 675: tbb::parallel_for over a tbb::blocked_range<int> of image rows) are answered with an OpenMP
// loop that hands the body one row at a time, dynamically scheduled — the closest stand-in for TBB's
// work-stealing auto partitioner.  (The reference's USE_OPENMP alternative does not compile: it
// returns out of the structured block, :557.)
#ifndef SR_REF_SHIM_TBB
#define SR_REF_SHIM_TBB
namespace tbb {
template <class T> class blocked_range {
public:
    blocked_range(T b, T e) : b_(b), e_(e) {}
    T begin() const { return b_; }
    T end() const { return e_; }
private:
    T b_, e_;
};
template <class Range, class Body> void parallel_for(const Range &r, const Body &body) {
    const int b = (int)r.begin(), e = (int)r.end();
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = b; i < e; ++i) body(Range(i, i + 1));
}
}  // namespace tbb
#endif
