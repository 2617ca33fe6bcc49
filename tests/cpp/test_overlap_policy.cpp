// CPU test of the launch-overlap bookkeeping (caf_cookoff_b200/csrc/overlap_policy.hpp): which launches may skip the wait
// for the grid before them.  Built and run by tests/test_overlap_policy.py; exits non-zero on the first failed check.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../caf_cookoff_b200/csrc/overlap_policy.hpp"

using caf_host::Range;
static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED line %d: %s\n", __LINE__, #c); ++failures; } } while (0)

struct Launch { Range in[3]; Range out[6]; };
static char arena[1 << 20];
static Range R(size_t off, size_t n) { return Range{arena + off, n}; }
// a surface launch: inputs at in_off (needle 100, haystack 100, grid 50 shared by everybody at 900000), outputs at out_off
static Launch mk(size_t in_off, size_t out_off, bool with_surface = true) {
    Launch l{};
    l.in[0] = R(in_off, 100); l.in[1] = R(in_off + 100, 100); l.in[2] = R(900000, 50);
    l.out[0] = with_surface ? R(out_off, 1000) : Range{nullptr, 0};
    l.out[1] = R(out_off + 1000, 40); l.out[2] = R(out_off + 1040, 40); l.out[3] = R(out_off + 1080, 32);
    return l;
}

int main() {
    constexpr int H = 7;
    caf_host::OverlapHistory<H> h;
    unsigned long long launches = 10;
    auto issue = [&](const Launch& l, bool expect_indep) {
        const bool indep = h.independent(l.in, l.out, launches);
        CHECK(indep == expect_indep);
        ++launches;                                    // the launch itself
        h.push(l.in, l.out, launches, !indep);
        return indep;
    };
    // the first launch has nothing to be independent of
    issue(mk(0, 10000), false);
    CHECK(h.valid == 1);
    // disjoint buffers, shared read-only grid: independent, as long as the history holds
    for (int k = 1; k <= 12; ++k) issue(mk(1000 * k, 10000 + 2000 * k), true);
    CHECK(h.valid == H);
    // write after write: the surface of the launch 3 back
    issue(mk(50000, 10000 + 2000 * 10), false);
    CHECK(h.valid == 1);                               // it waited: the history restarts with it
    issue(mk(51000, 200000), true);
    // write after read: output over an earlier launch's needle
    { Launch l = mk(52000, 210000); l.out[0] = R(51000 + 50, 10); issue(l, false); }
    issue(mk(53000, 220000), true);
    // read after write: needle inside an earlier launch's surface
    { Launch l = mk(220000 + 10, 230000); issue(l, false); }
    issue(mk(54000, 240000), true);
    // peaks only (no surface) still conflict on the peak slot
    { Launch l = mk(55000, 240000, false); issue(l, false); }
    issue(mk(56000, 250000, false), true);
    // something else was launched by the library in between: the history is void
    ++launches;
    issue(mk(57000, 260000), false);
    issue(mk(58000, 270000), true);
    // an alias older than the history is not seen (the CTA-counting argument of DESIGN.md covers it): 8 launches back
    const size_t old_out = 300000;
    issue(mk(60000, old_out), true);
    for (int k = 1; k <= H; ++k) issue(mk(60000 + 1000 * k, old_out + 3000 * k), true);
    issue(mk(70000, old_out), true);                   // 8 back: outside the 7 compared
    issue(mk(71000, old_out + 3000 * 3), false);       // 5 back: inside
    // empty ranges never meet anything
    CHECK(!caf_host::ranges_meet(Range{arena, 0}, Range{arena, 10}));
    CHECK(caf_host::ranges_meet(Range{arena, 10}, Range{arena + 9, 1}));
    CHECK(!caf_host::ranges_meet(Range{arena, 10}, Range{arena + 10, 1}));
    // reset
    h.reset();
    issue(mk(80000, 400000), false);
    // grids of overlapped launches
    CHECK(caf_host::overlapped_grid(148, 1) == 148 && caf_host::overlapped_grid(148, 2) == 74 &&
          caf_host::overlapped_grid(148, 3) == 50 && caf_host::overlapped_grid(148, 4) == 37);
    if (failures) { std::printf("%d check(s) failed\n", failures); return 1; }
    std::printf("overlap policy: all checks passed\n");
    return 0;
}
