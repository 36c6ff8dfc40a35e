// oracle/_ref only: <plotting/plotter.hpp> of the reference includes <plotting/slice_plot.hpp>, which drags the whole
// simulation layer (boost, HighFive, NDArray: not available offline) into src/cell.cpp for the sake of two colour maps
// used by the YAML cell factory.  This empty header takes its place on the include path; the plotter is not built.
#pragma once
