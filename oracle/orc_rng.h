/* orc_rng.h -- TEST INFRASTRUCTURE (oracle side).
 *
 * Restates the random-number path of the reference:
 *   - engine: pcg32 (setseq XSH-RR 64/32) from pcg-cpp, which is NOT vendored
 *     in /root/reference (CMakeLists.txt:98-105 fetches HunterBelanger/pcg-cpp,
 *     branch feature/cmake, unpinned).  Published algorithm (O'Neill 2014):
 *     state' = state*6364136223846793005 + inc, output on the pre-advance state,
 *     x = ((s>>18)^s)>>27 rotated right by s>>59; seed(s): state=(s+inc)*mult+inc;
 *     advance(d): O(log d) LCG jump.  Default increment 1442695040888963407.
 *   - distributions: libstdc++ 13 <random> semantics used through
 *     include/utils/rng.hpp:41-96 (uniform_real, exponential, discrete):
 *     generate_canonical<double,53> = 2 engine calls, (lo + hi*2^32)/2^64
 *     (bits/random.tcc:3349-3381); exponential = -log(1-xi)/lambda
 *     (bits/random.h:4904) with the reference's lambda==0 -> INF guard
 *     (rng.hpp:75); discrete: <2 weights -> 0 without drawing, else
 *     normalise, partial sums, last=1.0, lower_bound (random.tcc:2655-2713).
 * Pinned by tests/test_rng_kat.py against vectors generated with the
 * pcg header vendored inside pyarrow + libstdc++ itself (tests/golden/).
 */
#ifndef ORC_RNG_H
#define ORC_RNG_H
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace orc {

constexpr double INF = std::numeric_limits<double>::max();  // constants.hpp:46
constexpr double PI = 3.14159265358979323846264338327950288;  // constants.hpp:48

struct MathFns {
  double (*log)(double);
  double (*sin)(double);
  double (*cos)(double);
  double (*exp)(double);
  double (*acos)(double);
};
extern MathFns g_math;  // selected by orc_set_math()

struct Pcg32 {
  static constexpr uint64_t MULT = 6364136223846793005ULL;
  static constexpr uint64_t INC = 1442695040888963407ULL;
  uint64_t state = 0;
  uint64_t ndraw = 0;  // engine outputs produced (instrumentation only)

  void seed(uint64_t s) {
    state = (s + INC) * MULT + INC;
    ndraw = 0;
  }
  uint32_t next() {
    uint64_t old = state;
    state = old * MULT + INC;
    ndraw++;
    uint32_t xorshifted = static_cast<uint32_t>(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = static_cast<uint32_t>(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
  }
  void advance(uint64_t delta) {
    uint64_t acc_mult = 1, acc_plus = 0, cur_mult = MULT, cur_plus = INC;
    while (delta > 0) {
      if (delta & 1) {
        acc_mult *= cur_mult;
        acc_plus = acc_plus * cur_mult + cur_plus;
      }
      cur_plus = (cur_mult + 1) * cur_plus;
      cur_mult *= cur_mult;
      delta >>= 1;
    }
    state = acc_mult * state + acc_plus;
  }
};

// settings::rng (settings.cpp:59,116-119): the one global engine, seeded with rng_seed on the default stream and then switched to
// stream 2 (pcg-cpp set_stream: inc = (2 << 1) | 1, the state is kept).  Regional cancellation and the branchless comb draw from
// it; it also is the UniformRandomBitGenerator handed to std::shuffle (branchless_power_iterator.cpp:619-650), hence the typedefs.
struct Pcg32Stream {
  using result_type = uint32_t;
  static constexpr result_type min() { return 0u; }
  static constexpr result_type max() { return 0xffffffffu; }
  uint64_t state = 0, inc = Pcg32::INC;
  void seed_global(uint64_t s) {  // initialize_global_rng()
    state = (s + Pcg32::INC) * Pcg32::MULT + Pcg32::INC;
    inc = (2ULL << 1) | 1ULL;
  }
  uint32_t next() {
    const uint64_t old = state;
    state = old * Pcg32::MULT + inc;
    const uint32_t xorshifted = static_cast<uint32_t>(((old >> 18u) ^ old) >> 27u);
    const uint32_t rot = static_cast<uint32_t>(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
  }
  result_type operator()() { return next(); }
  void advance(uint64_t delta) {
    uint64_t acc_mult = 1, acc_plus = 0, cur_mult = Pcg32::MULT, cur_plus = inc;
    while (delta > 0) {
      if (delta & 1) {
        acc_mult *= cur_mult;
        acc_plus = acc_plus * cur_mult + cur_plus;
      }
      cur_plus = (cur_mult + 1) * cur_plus;
      cur_mult *= cur_mult;
      delta >>= 1;
    }
    state = acc_mult * state + acc_plus;
  }
};

// RNG::rand  (rng.hpp:41) == libstdc++ generate_canonical<double,53>(pcg32)
template <class Engine>
inline double rng_rand(Engine& g) {
  double sum = 0.0, tmp = 1.0;
  sum += static_cast<double>(g.next()) * tmp;
  tmp *= 4294967296.0;
  sum += static_cast<double>(g.next()) * tmp;
  tmp *= 4294967296.0;
  double ret = sum / tmp;
  if (ret >= 1.0) ret = std::nextafter(1.0, 0.0);
  return ret;
}

// RNG::exponential (rng.hpp:74-79)
inline double rng_exponential(Pcg32& g, double lambda) {
  if (lambda == 0.) return INF;
  return -g_math.log(1.0 - rng_rand(g)) / lambda;
}

// std::discrete_distribution cumulative table, built once per weight row.
// Returns an EMPTY table for <2 weights (no draw is consumed then).
inline std::vector<double> discrete_table(const double* w, size_t n) {
  std::vector<double> cp;
  if (n < 2) return cp;
  double sum = 0.0;
  for (size_t i = 0; i < n; i++) sum += w[i];
  std::vector<double> p(n);
  for (size_t i = 0; i < n; i++) p[i] = w[i] / sum;
  cp.resize(n);
  double acc = p[0];
  cp[0] = acc;
  for (size_t i = 1; i < n; i++) {
    acc = acc + p[i];
    cp[i] = acc;
  }
  cp[n - 1] = 1.0;
  return cp;
}

// RNG::discrete (rng.hpp:88-96) with the table above
template <class Engine>
inline int rng_discrete(Engine& g, const std::vector<double>& cp) {
  if (cp.empty()) return 0;
  const double p = rng_rand(g);
  size_t lo = 0, len = cp.size();  // std::lower_bound
  while (len > 0) {
    size_t half = len >> 1;
    if (cp[lo + half] < p) {
      lo = lo + half + 1;
      len = len - half - 1;
    } else {
      len = half;
    }
  }
  return static_cast<int>(lo);
}

}  // namespace orc
#endif
