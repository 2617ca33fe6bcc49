// C++ restatement of caf_rust/tests/test.rs (all 13 tests) on the C++ mirror of the reference API.
// Built and run by tests/test_gpu_cpp.py on the GPU box:  ./test_rs <data dir>
#include <algorithm>
#include <cstdio>
#include <string>
#include "caf_b200.hpp"

using namespace caf;
static int failures = 0;

#define ASSERT_EQ(a, b) do { if (!((a) == (b))) { std::printf("  FAILED %s == %s  (line %d)\n", #a, #b, __LINE__); ++failures; } } while (0)

template <class Strategy>
static void run_case(const char* name, const std::string& dir, const char* raw, const char* hayfile,
                     double start, double end, double step, double want_freq, std::size_t want_idx) {
    auto needle = read_file_c64(dir + "/" + raw);
    auto haystack = read_file_c64(dir + "/" + hayfile);
    haystack.resize(needle.size());                                    // &haystack[..needle.len()]
    auto shifts = gen_float_shifts(start, end, step);
    auto surface = Strategy::caf_surface(needle, haystack, shifts, 48000);
    auto [freq, samp_idx] = Strategy::find_peak(std::move(surface));
    std::printf("%s: freq %.2f idx %zu\n", name, freq, samp_idx);
    ASSERT_EQ(freq, want_freq);
    ASSERT_EQ(samp_idx, want_idx);
}

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "../data";
    try {
        const char* c0 = "chirp_0_T+202samp_F+69.25Hz.c64";
        run_case<CafRustFFT>("test_rustfft_chirp0", dir, "chirp_0_raw.c64", c0, -100.0, 100.0, 0.25, 69.25, 202);
        run_case<CafRustFFTIter>("test_rustfft_iter_chirp0", dir, "chirp_0_raw.c64", c0, -100.0, 100.0, 0.25, 69.25, 202);
        run_case<CafRustFFTRayon>("test_rustfft_rayon_chirp0", dir, "chirp_0_raw.c64", c0, -100.0, 100.0, 0.25, 69.25, 202);
        run_case<CafRustFFTIterRayon>("test_rustfft_iter_rayon_chirp0", dir, "chirp_0_raw.c64", c0, -100.0, 100.0, 0.25, 69.25, 202);
        run_case<CafRustFFTThreads>("test_rustfft_threads_chirp0", dir, "chirp_0_raw.c64", c0, -100.0, 100.0, 0.25, 69.25, 202);
        run_case<CafRustFFTThreadpool>("test_rustfft_threadpool_chirp0", dir, "chirp_0_raw.c64", c0, -100.0, 100.0, 0.25, 69.25, 202);
        run_case<CafFFTW>("test_fftw_chirp0", dir, "chirp_0_raw.c64", c0, -100.0, 100.0, 0.25, 69.25, 202);
        run_case<CafRustFFTThreads>("chirp1", dir, "chirp_1_raw.c64", "chirp_1_T+78samp_F+35.99Hz.c64", -50.0, 50.0, 1.0, 36.0, 78);
        run_case<CafRustFFTThreads>("chirp2", dir, "chirp_2_raw.c64", "chirp_2_T+169samp_F+32.16Hz.c64", 30.0, 35.0, 0.05, 32.15, 169);
        run_case<CafRustFFTThreads>("chirp3", dir, "chirp_3_raw.c64", "chirp_3_T+151samp_F-76.22Hz.c64", -100.0, 100.0, 0.25, -76.25, 151);
        run_case<CafRustFFTThreads>("chirp4", dir, "chirp_4_raw.c64", "chirp_4_T+70samp_F+82.89Hz.c64", 80.0, 100.0, 0.1, 82.9, 70);
        run_case<CafRustFFTThreads>("chirp5", dir, "chirp_5_raw.c64", "chirp_5_T+177samp_F-92.72Hz.c64", -100.0, 100.0, 0.25, -92.75, 177);
        run_case<CafRustFFTThreads>("chirp6", dir, "chirp_6_raw.c64", "chirp_6_T+15samp_F-49.69Hz.c64", -100.0, 100.0, 0.25, -49.75, 15);
        run_case<CafRustFFTThreads>("chirp7", dir, "chirp_7_raw.c64", "chirp_7_T+84samp_F+68.26Hz.c64", -100.0, 100.0, 0.25, 68.25, 84);
        run_case<CafRustFFTThreads>("chirp8", dir, "chirp_8_raw.c64", "chirp_8_T+80samp_F-46.28Hz.c64", -100.0, 100.0, 0.25, -46.25, 80);
        run_case<CafRustFFTThreads>("chirp9", dir, "chirp_9_raw.c64", "chirp_9_T+176samp_F+61.49Hz.c64", -100.0, 100.0, 0.5, 61.5, 176);
        // length mismatch panics like xcor_rustfft.rs:54-55
        bool panicked = false;
        try { CafRustFFT::caf_surface(std::vector<Complex64>(8), std::vector<Complex64>(9), {0.0}, 48000); } catch (const Panic&) { panicked = true; }
        ASSERT_EQ(panicked, true);
        // main.rs:10-32 equivalent: peak-only call prints the two lines of the CLI
        auto needle = read_file_c64(dir + "/chirp_0_raw.c64");
        auto hay = read_file_c64(dir + "/" + c0); hay.resize(needle.size());
        std::vector<double> shifts; for (int m = -100000; m < 100000; m += 500) shifts.push_back(m / 1e3);
        auto [f, idx] = CafB200::caf_peak(needle, hay, shifts, 48000);
        std::printf("Frequency offset: %.1fHz\nTime offset: %zu samples (%.3fms)\n", f, idx, (double)idx / 48.0);
        ASSERT_EQ(f, 69.0); ASSERT_EQ(idx, (std::size_t)202);
        // lazy rows: caf_surface leaves the surface on the GPU; a row's xcor_mag crosses PCIe only when asked for, and is
        // the same bits the host-surface entry point delivers; find_peak on a reordered / truncated vector scans the rows
        {
            auto rows = CafRustFFT::caf_surface(needle, hay, shifts, 48000);
            ASSERT_EQ(rows.size(), shifts.size());
            ASSERT_EQ(rows[0].surface()->cells_per_row(), (std::size_t)8192);
            std::vector<double> dense(shifts.size() * 8192), pv(shifts.size());
            std::vector<uint64_t> pi(shifts.size());
            check(caf_b200_surface_f64(thread_handle(), reinterpret_cast<const caf_c128*>(needle.data()),
                                       reinterpret_cast<const caf_c128*>(hay.data()), needle.size(), shifts.data(), shifts.size(), 48000,
                                       dense.data(), pv.data(), pi.data(), nullptr));
            for (std::size_t r : {std::size_t(0), std::size_t(338), shifts.size() - 1}) {
                auto mag = rows[r].xcor_mag();
                ASSERT_EQ(mag.size(), (std::size_t)8192);
                ASSERT_EQ(std::equal(mag.begin(), mag.end(), dense.begin() + (std::ptrdiff_t)(r * 8192)), true);
                ASSERT_EQ(rows[r].xcor_peak_val(), pv[r]); ASSERT_EQ(rows[r].xcor_peak_idx(), (std::size_t)pi[r]); ASSERT_EQ(rows[r].freq(), shifts[r]);
                ASSERT_EQ(mag[rows[r].xcor_peak_idx()], rows[r].xcor_peak_val());
            }
            auto whole = CafRustFFT::find_peak(rows);                              // untouched: the kernel's fused answer
            ASSERT_EQ(whole.first, 69.0); ASSERT_EQ(whole.second, (std::size_t)202);
            std::vector<CafSurfaceRow> rev(rows.rbegin(), rows.rend());           // reordered: scanned in vector order
            auto r2 = CafRustFFT::find_peak(rev);
            ASSERT_EQ(r2.first, 69.0); ASSERT_EQ(r2.second, (std::size_t)202);
            std::vector<CafSurfaceRow> head(rows.begin(), rows.begin() + 100);   // truncated: the best of rows 0..99
            auto r3 = CafRustFFT::find_peak(head);
            double bv = 0.0, bf = 0.0; std::size_t bi = 0;
            for (std::size_t r = 0; r < 100; ++r) if (pv[r] > bv) { bv = pv[r]; bf = shifts[r]; bi = (std::size_t)pi[r]; }
            ASSERT_EQ(r3.first, bf); ASSERT_EQ(r3.second, bi);
            rows.clear(); rev.clear();                                            // `head` keeps the surface alive
            ASSERT_EQ(head[5].xcor_mag().size(), (std::size_t)8192);
        }
        // main.go:13-35 equivalent on the chirp_4 pair: "caf result: 70 samples 83 hz"
        {
            auto apple = read_file_c64(dir + "/chirp_4_raw.c64");
            auto banana = read_file_c64(dir + "/chirp_4_T+70samp_F+82.89Hz.c64"); banana.resize(4096);
            std::vector<double> fz; for (double x = 80.0; x < 86.0; x += .5) fz.push_back(x);
            auto surf = GoSibling::amb_surf(apple, banana, fz, 48000);
            auto pk = GoSibling::find_2d_peak(surf);
            std::printf("caf result: %d samples %g hz @ amb = %g\n", (int)apple.size() - pk.tdx, fz[pk.fdx], pk.max);
            ASSERT_EQ((int)apple.size() - pk.tdx, 70); ASSERT_EQ(fz[pk.fdx], 83.0);
        }
    } catch (const std::exception& e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
    std::printf(failures ? "%d FAILURES\n" : "all tests passed\n", failures);
    return failures ? 1 : 0;
}
