// Stand-in for the NDArray header (github.com/HunterBelanger/ndarray, downloaded by the reference's CMake; not available
// offline), for oracle/_ref only.  A dense row-major (C-order) array with the members the mesh tallies and the delta
// tracker use: reallocate, fill, operator()(i...), operator[], size, shape, data_vector.
#pragma once
#include <cstddef>
#include <initializer_list>
#include <stdexcept>
#include <vector>
template <class T>
class NDArray {
 public:
  NDArray() = default;
  NDArray(std::initializer_list<std::size_t> shape) { reallocate(std::vector<std::size_t>(shape)); }
  explicit NDArray(const std::vector<std::size_t>& shape) { reallocate(shape); }
  void reallocate(const std::vector<std::size_t>& shape) {
    shape_ = shape;
    std::size_t n = 1;
    for (auto s : shape_) n *= s;
    data_.assign(n, T());
  }
  void reallocate(std::initializer_list<std::size_t> shape) { reallocate(std::vector<std::size_t>(shape)); }
  void fill(const T& v) { data_.assign(data_.size(), v); }
  std::size_t size() const { return data_.size(); }
  const std::vector<std::size_t>& shape() const { return shape_; }
  std::vector<T>& data_vector() { return data_; }
  const std::vector<T>& data_vector() const { return data_; }
  T& operator[](std::size_t i) { return data_[i]; }
  const T& operator[](std::size_t i) const { return data_[i]; }
  template <class... I>
  T& operator()(I... idx) { return data_[offset({static_cast<std::size_t>(idx)...})]; }
  template <class... I>
  const T& operator()(I... idx) const { return data_[offset({static_cast<std::size_t>(idx)...})]; }

 private:
  std::size_t offset(std::initializer_list<std::size_t> idx) const {
    if (idx.size() != shape_.size()) throw std::out_of_range("NDArray stand-in: rank mismatch");
    std::size_t o = 0, d = 0;
    for (auto i : idx) o = o * shape_[d++] + i;
    return o;
  }
  std::vector<std::size_t> shape_;
  std::vector<T> data_;
};
