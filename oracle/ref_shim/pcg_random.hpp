// <pcg_random.hpp> for oracle/_ref only: the reference takes pcg32 from pcg-cpp (fetched from the network by its CMake).
// pyarrow vendors a byte-identical copy of that header in namespace arrow_vendored; expose it under the name the
// reference's include/utils/rng.hpp expects.
#pragma once
#include <arrow/vendored/pcg/pcg_random.hpp>
using pcg32 = arrow_vendored::pcg32;
