// Stand-in for pndl::CrossSection, for oracle/_ref only: see energy_grid.hpp in this directory.
#pragma once
#include <PapillonNDL/energy_grid.hpp>

#include <memory>
#include <vector>
namespace pndl {
class CrossSection {
 public:
  CrossSection(const std::vector<double>& xs, std::shared_ptr<EnergyGrid> grid, std::size_t index)
      : xs_(xs), grid_(grid), index_(index) {}
  double evaluate(double E, std::size_t i) const {
    const double x0 = (*grid_)[i], x1 = (*grid_)[i + 1], y0 = xs_[i - index_], y1 = xs_[i - index_ + 1];
    if (x1 == x0) return y0;
    return y0 + (E - x0) / (x1 - x0) * (y1 - y0);
  }
  double operator()(double E) const { return evaluate(E, grid_->get_lower_index(E)); }

 private:
  std::vector<double> xs_;
  std::shared_ptr<EnergyGrid> grid_;
  std::size_t index_;
};
}  // namespace pndl
