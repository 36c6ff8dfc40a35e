// Generates tests/golden/rng_kat.json from the REAL third-party code the reference uses for random
// numbers: the pcg32 engine (pcg-cpp; here the byte-identical copy vendored inside pyarrow, namespace
// arrow_vendored::pcg32) driven through libstdc++ 13's uniform_real / exponential / discrete
// distributions exactly as include/utils/rng.hpp:41-96 and include/simulation/particle.hpp:188-193 do.
// Build + run (in the build container only; the JSON is committed):
//   g++ -O2 -std=c++17 -I$(python -c 'import pyarrow,os;print(os.path.join(os.path.dirname(pyarrow.__file__),"include"))') \
//       tests/golden/gen_rng_kat.cpp -o /tmp/gen_rng_kat && /tmp/gen_rng_kat > tests/golden/rng_kat.json
#include <arrow/vendored/pcg/pcg_random.hpp>

#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

using pcg32 = arrow_vendored::pcg32;

int main() {
  const uint64_t seed = 19073486328125ULL, stride = 152917ULL;  // src/settings.cpp:57-58
  const uint64_t ids[] = {0, 1, 2, 7, 100000, 123456789, 4000000000ULL, 18446744073709ULL};
  std::printf("{\n \"seed\": %llu, \"stride\": %llu,\n \"histories\": [\n", (unsigned long long)seed, (unsigned long long)stride);
  bool first = true;
  for (uint64_t id : ids) {
    pcg32 a;
    a.seed(seed);
    a.advance(stride * id);
    std::printf("%s  {\"id\": %llu, \"u32\": [", first ? "" : ",\n", (unsigned long long)id);
    first = false;
    for (int i = 0; i < 8; i++) std::printf("%s%u", i ? ", " : "", (unsigned)a());
    std::printf("], \"rand\": [");
    pcg32 b;
    b.seed(seed);
    b.advance(stride * id);
    std::uniform_real_distribution<double> unit(0., 1.);
    for (int i = 0; i < 6; i++) std::printf("%s%.17g", i ? ", " : "", unit(b));
    std::printf("], \"exponential_lambda_0.5\": ");
    pcg32 c;
    c.seed(seed);
    c.advance(stride * id);
    std::exponential_distribution<double> ex(0.5);
    std::printf("%.17g", ex(c));
    std::printf(", \"discrete\": [");
    pcg32 d;
    d.seed(seed);
    d.advance(stride * id);
    const std::vector<double> w{0.58791, 0.41176, 3.3906e-4, 1.1761e-7, 0., 0., 0.};
    for (int i = 0; i < 12; i++) {
      std::discrete_distribution<int> dist(w.begin(), w.end());
      std::printf("%s%d", i ? ", " : "", dist(d));
    }
    std::printf("], \"discrete_single_draws\": ");
    pcg32 e;
    e.seed(seed);
    e.advance(stride * id);
    const std::vector<double> one{1.0};
    std::discrete_distribution<int> d1(one.begin(), one.end());
    const int r1 = d1(e);
    pcg32 e2;
    e2.seed(seed);
    e2.advance(stride * id);
    std::printf("%d, \"discrete_single_consumed\": %s}", r1, (e == e2) ? "false" : "true");
  }
  std::printf("\n ],\n \"discrete_weights\": [0.58791, 0.41176, 3.3906e-4, 1.1761e-7, 0.0, 0.0, 0.0]\n}\n");
  return 0;
}
