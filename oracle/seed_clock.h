/* Pre-included (g++ -include) when the reference DRIVERS are compiled for oracle/_ref.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference drivers seed std::mt19937 from the wall clock
 * (FRIES_bin/frisys_mol.cpp:104, frifull_mol.cpp:63, frisys_hh.cpp:66).  To make reference runs
 * reproducible WITHOUT editing or copying the reference sources, the recipe renames the clock they
 * read to one whose "now" is the integer in $FRIES_SEED (default 1).  Loop time is measured by the
 * harness as the wall-clock difference of two runs with different --max_iter.
 */
#ifndef FRIES_B200_ORACLE_SEED_CLOCK_H
#define FRIES_B200_ORACLE_SEED_CLOCK_H
#ifdef __cplusplus
#include <chrono>
#include <cstdlib>

namespace fries_seed {
struct clock {
    typedef std::chrono::nanoseconds duration;
    typedef duration::rep rep;
    typedef duration::period period;
    typedef std::chrono::time_point<clock, duration> time_point;
    static constexpr bool is_steady = true;
    static time_point now() noexcept {
        const char *s = std::getenv("FRIES_SEED");
        long long v = s ? std::atoll(s) : 1;
        return time_point(duration(v));
    }
};
}  // namespace fries_seed
namespace std { namespace chrono { using fries_seed_clock = ::fries_seed::clock; } }
#define high_resolution_clock fries_seed_clock
#endif
#endif
