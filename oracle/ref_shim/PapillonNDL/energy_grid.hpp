// Stand-in for pndl::EnergyGrid (PapillonNDL is not available offline), for oracle/_ref only.  The delta and carter
// trackers look the majorant up with get_lower_index(E) and CrossSection::evaluate(E, i).  In multigroup mode the grid
// is [b0, b1, b1, b2, b2, ...] with the same majorant on both points of a group (src/majorant.cpp:132-176) and E is a
// group mid-point, so ANY bracketing search followed by linear interpolation between two equal values returns that
// value exactly; this stand-in does exactly that.
#pragma once
#include <algorithm>
#include <cstddef>
#include <vector>
namespace pndl {
class EnergyGrid {
 public:
  explicit EnergyGrid(const std::vector<double>& grid) : grid_(grid) {}
  std::size_t get_lower_index(double E) const {
    if (E <= grid_.front()) return 0;
    if (E >= grid_.back()) return grid_.size() - 2;
    return static_cast<std::size_t>(std::upper_bound(grid_.begin(), grid_.end(), E) - grid_.begin()) - 1;
  }
  double operator[](std::size_t i) const { return grid_[i]; }
  std::size_t size() const { return grid_.size(); }
  const std::vector<double>& grid() const { return grid_; }

 private:
  std::vector<double> grid_;
};
}  // namespace pndl
