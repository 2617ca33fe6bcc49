// overlap_policy.hpp -- host-side bookkeeping of caf_b200_set_overlap (include/caf_b200.h): may a single-pair surface
// launch skip the wait for the grid before it?  Plain C++ (no CUDA), so tests/cpp/test_overlap_policy.cpp exercises it on
// a machine without a GPU.
//
// A launch is described by the byte ranges it reads (needle, haystack, doppler grid) and writes (surface, row peaks,
// peak, packed words).  It is independent of an earlier launch when no range it WRITES meets a range the earlier one
// reads or writes, and no range it READS meets a range the earlier one writes (read-read sharing is free: every step of
// a doppler sweep reads the same grid).  The history holds the last kHist overlappable launches, newest first; a launch
// may overlap only if it is independent of all of them, directly follows the newest one on the handle (nothing else
// launched in between), and the history is not empty.  Anything else makes it wait and restarts the history with it.
#pragma once
#include <cstddef>

namespace caf_host {

struct Range { const char* p; size_t n; };

inline bool ranges_meet(const Range& x, const Range& y) {
    return x.n && y.n && x.p < y.p + y.n && y.p < x.p + x.n;
}

template <int kHist, int kIn = 3, int kOut = 6>
struct OverlapHistory {
    Range in[kHist][kIn] = {}, out[kHist][kOut] = {};
    int valid = 0;                               // launches in the history
    unsigned long long launch_no = ~0ull;        // the handle's launch count right after the newest of them

    void reset() { valid = 0; launch_no = ~0ull; }

    // Decide for a launch with these buffers, issued when the handle's launch count is `launches_now`.  Does not record it.
    bool independent(const Range (&cur_in)[kIn], const Range (&cur_out)[kOut], unsigned long long launches_now) {
        if (launch_no != launches_now) valid = 0;                // something else ran in between
        if (valid == 0) return false;
        for (int q = 0; q < valid; ++q) {
            for (int i = 0; i < kOut; ++i) {
                for (int j = 0; j < kOut; ++j) if (ranges_meet(cur_out[i], out[q][j])) return false;     // write after write
                for (int j = 0; j < kIn; ++j) if (ranges_meet(cur_out[i], in[q][j])) return false;       // write after read
            }
            for (int i = 0; i < kIn; ++i)
                for (int j = 0; j < kOut; ++j) if (ranges_meet(cur_in[i], out[q][j])) return false;      // read after write
        }
        return true;
    }
    // Record an overlappable launch that has just been issued; `waited` = it was NOT independent (it waited for everything
    // before it, so the history restarts with it).
    void push(const Range (&cur_in)[kIn], const Range (&cur_out)[kOut], unsigned long long launches_after, bool waited) {
        if (waited) valid = 0;
        for (int q = kHist - 1; q > 0; --q) {
            for (int i = 0; i < kIn; ++i) in[q][i] = in[q - 1][i];
            for (int i = 0; i < kOut; ++i) out[q][i] = out[q - 1][i];
        }
        for (int i = 0; i < kIn; ++i) in[0][i] = cur_in[i];
        for (int i = 0; i < kOut; ++i) out[0][i] = cur_out[i];
        valid = valid < kHist ? valid + 1 : kHist;
        launch_no = launches_after;
    }
};

// CTAs of an overlapped launch: mode 1 = one per SM, mode n >= 2 = ceil(full / n) so that about n launches share the GPU.
inline long long overlapped_grid(long long full, int mode) { return mode >= 2 ? (full + mode - 1) / mode : full; }

}  // namespace caf_host
