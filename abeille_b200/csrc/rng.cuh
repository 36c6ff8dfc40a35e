/* rng.cuh -- per-history random number streams.
 *
 * Reproduces the reference's stream layout bit for bit:
 *   engine   pcg32 = PCG setseq XSH-RR 64/32 with the default increment (pcg-cpp, un-vendored
 *            dependency of the reference, CMakeLists.txt:98-105); each history starts at
 *            seed(s); advance(stride * history_id)      (include/simulation/particle.hpp:188-193)
 *   rand     libstdc++ generate_canonical<double,53>: two engine outputs, (lo + hi*2^32) / 2^64,
 *            clamped below 1                            (include/utils/rng.hpp:41)
 *   exponential  -log(1 - xi) / lambda, INF without a draw when lambda == 0   (rng.hpp:74-79)
 *   discrete     libstdc++ discrete_distribution: fewer than 2 weights -> 0 and NO draw,
 *                otherwise lower_bound over the partial sums built on the host    (rng.hpp:88-96)
 * The O(log d) advance is replaced by a table of the 64 power-of-two LCG jumps (one multiply-add
 * per set bit of the distance), built once per problem on the host.
 */
#pragma once
#include <stdint.h>

#include "detmath.cuh"

namespace abl {

#define ABL_PCG_MULT 6364136223846793005ULL
#define ABL_PCG_INC 1442695040888963407ULL
#define ABL_INF 1.7976931348623157e308 /* std::numeric_limits<double>::max(), constants.hpp:46 */
#define ABL_PI 3.14159265358979323846264338327950288

struct JumpTable {  // state' = mult[k]*state + plus[k] advances the LCG by 2^k steps
  uint64_t mult[64];
  uint64_t plus[64];
};

__host__ __device__ inline uint64_t pcg_seed_state(uint64_t seed) { return (seed + ABL_PCG_INC) * ABL_PCG_MULT + ABL_PCG_INC; }

__device__ __forceinline__ uint64_t pcg_advance(uint64_t state, uint64_t delta, const JumpTable* __restrict__ jt) {
  while (delta) {
    const int k = __ffsll((long long)delta) - 1;
    state = ldt(&jt->mult[k]) * state + ldt(&jt->plus[k]);
    delta &= delta - 1;
  }
  return state;
}

__device__ __forceinline__ uint32_t pcg_next(uint64_t& state) {
  const uint64_t old = state;
  state = old * ABL_PCG_MULT + ABL_PCG_INC;
  const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
  const uint32_t rot = (uint32_t)(old >> 59u);
  return __funnelshift_r(xorshifted, xorshifted, rot);
}

__device__ __forceinline__ double rng_rand(uint64_t& state) {
  const uint32_t lo = pcg_next(state);
  const uint32_t hi = pcg_next(state);
  // (double(lo) + double(hi)*2^32) / 2^64 ; the sum is rounded exactly as in libstdc++
  double sum = (double)lo;
  sum += (double)hi * 4294967296.0;
  double ret = sum / 18446744073709551616.0;
  if (ret >= 1.0) ret = 0.99999999999999988897769753748;  // nextafter(1, 0)
  return ret;
}

// ---- math policies ------------------------------------------------------------------------------------------
// The same IEEE operations either expanded inline (InlineMath) or issued as calls to ONE shared copy per
// kernel (CallMath).  fp64 division, sqrt and the pcg32 -> double conversion expand to 15-40 instructions each;
// the delta-tracking loop uses dozens of them, and ncu showed the kernel bound by SM instruction-cache misses
// (sm__icc hit rate 57 %), so the hot loop trades a call for the footprint.  Results are bit-identical.
struct RandOut {
  double v;
  uint64_t state;
};
static __device__ __noinline__ RandOut rng_rand_call(uint64_t state) {
  RandOut o;
  o.v = rng_rand(state);
  o.state = state;
  return o;
}
static __device__ __noinline__ double ddiv_call(double a, double b) { return a / b; }
static __device__ __noinline__ double dsqrt_call(double a) { return sqrt(a); }

struct InlineMath {
  static __device__ __forceinline__ double rand(uint64_t& s) { return rng_rand(s); }
  static __device__ __forceinline__ double div(double a, double b) { return a / b; }
  static __device__ __forceinline__ double sqrt_(double a) { return sqrt(a); }
};
struct CallMath {
  static __device__ __forceinline__ double rand(uint64_t& s) {
    const RandOut o = rng_rand_call(s);
    s = o.state;
    return o.v;
  }
  static __device__ __forceinline__ double div(double a, double b) { return ddiv_call(a, b); }
  static __device__ __forceinline__ double sqrt_(double a) { return dsqrt_call(a); }
};

template <class M = InlineMath>
__device__ __forceinline__ double rng_exponential(uint64_t& state, double lambda) {
  if (lambda == 0.) return ABL_INF;
  return M::div(-det_log(1.0 - M::rand(state)), lambda);
}

// lower_bound over a cumulative table of n >= 2 entries (the caller handles n < 2: no draw).  The table is
// non-decreasing, so the lower bound is the number of entries below p: up to 8 entries are counted without a
// branch (multigroup decks have a handful of groups), longer tables use the binary search.
template <class M = InlineMath>
__device__ __forceinline__ int rng_discrete(uint64_t& state, const double* __restrict__ cp, int n) {
  const double p = M::rand(state);
  if (n <= 8) {
    int lo = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) lo += (i < n && ldt(&cp[i < n ? i : 0]) < p) ? 1 : 0;
    return lo;
  }
  int lo = 0, len = n;
  while (len > 0) {
    const int half = len >> 1;
    if (ldt(&cp[lo + half]) < p) {
      lo = lo + half + 1;
      len = len - half - 1;
    } else {
      len = half;
    }
  }
  return lo;
}

// a / b for a finite b > 0.  A zero numerator (nu*Sigma_f in a non-fissile material, wgt2 outside noise mode)
// would send CUDA's division through its slow-path subroutine; 0 / b is 0 with the sign of a, exactly.
template <class M = InlineMath>
__device__ __forceinline__ double ddiv_pos(double a, double b) { return (a == 0.) ? a : M::div(a, b); }

}  // namespace abl
