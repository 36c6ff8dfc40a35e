// Stand-in for Boost.Unordered (not available offline), for oracle/_ref only: MaterialHelper keeps two lookup caches in
// boost::unordered_flat_map and builds them in its per-history constructor (src/material_helper.cpp:28-40).  This is an
// open-addressing table in one contiguous allocation like Boost's (power-of-two capacity, linear probing, mixed hash, no
// erase), so that the CPU arm of bench.py pays what the reference pays for these maps -- a node-based std::unordered_map
// allocates per element and was the slower stand-in of round 1.  The interface is what the reference uses: default / copy
// construction and assignment, range-for over pairs, find / end, operator[], size, clear.  No result depends on iteration order.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>
namespace boost {
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>, class A = std::allocator<std::pair<const K, V>>>
class unordered_flat_map {
 public:
  using key_type = K;
  using mapped_type = V;
  using value_type = std::pair<K, V>;

  template <bool Const>
  class iter {
   public:
    using map_t = std::conditional_t<Const, const unordered_flat_map, unordered_flat_map>;
    using ref_t = std::conditional_t<Const, const value_type&, value_type&>;
    using ptr_t = std::conditional_t<Const, const value_type*, value_type*>;
    iter() = default;
    iter(map_t* m, std::size_t i) : m_(m), i_(i) { skip(); }
    ref_t operator*() const { return m_->slots_[i_]; }
    ptr_t operator->() const { return &m_->slots_[i_]; }
    iter& operator++() {
      ++i_;
      skip();
      return *this;
    }
    template <bool C2>
    bool operator==(const iter<C2>& o) const { return i_ == o.i_; }
    template <bool C2>
    bool operator!=(const iter<C2>& o) const { return i_ != o.i_; }
    std::size_t i_ = 0;

   private:
    void skip() {
      while (i_ < m_->used_.size() && !m_->used_[i_]) ++i_;
    }
    map_t* m_ = nullptr;
  };
  using iterator = iter<false>;
  using const_iterator = iter<true>;

  unordered_flat_map() = default;

  iterator begin() { return iterator(this, 0); }
  iterator end() { return iterator(this, used_.size()); }
  const_iterator begin() const { return const_iterator(this, 0); }
  const_iterator end() const { return const_iterator(this, used_.size()); }
  std::size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  void clear() {
    slots_.clear();
    used_.clear();
    n_ = 0;
  }

  iterator find(const K& k) {
    const std::size_t i = locate(k);
    return i == npos ? end() : iterator(this, i);
  }
  const_iterator find(const K& k) const {
    const std::size_t i = locate(k);
    return i == npos ? end() : const_iterator(this, i);
  }
  V& operator[](const K& k) {
    std::size_t i = locate(k);
    if (i != npos) return slots_[i].second;
    if (2 * (n_ + 1) > used_.size()) grow();
    i = free_slot(k);
    slots_[i].first = k;
    slots_[i].second = V();
    used_[i] = 1;
    n_++;
    return slots_[i].second;
  }

 private:
  static constexpr std::size_t npos = static_cast<std::size_t>(-1);
  std::size_t bucket(const K& k) const {
    std::uint64_t h = static_cast<std::uint64_t>(H()(k));
    h *= 0x9E3779B97F4A7C15ull;  // std::hash of a pointer or an integer is the identity: mix it
    return static_cast<std::size_t>(h >> 32) & (used_.size() - 1);
  }
  std::size_t locate(const K& k) const {
    if (used_.empty()) return npos;
    for (std::size_t i = bucket(k);; i = (i + 1) & (used_.size() - 1)) {
      if (!used_[i]) return npos;
      if (E()(slots_[i].first, k)) return i;
    }
  }
  std::size_t free_slot(const K& k) const {
    std::size_t i = bucket(k);
    while (used_[i]) i = (i + 1) & (used_.size() - 1);
    return i;
  }
  void grow() {
    std::vector<value_type> old_slots;
    std::vector<unsigned char> old_used;
    old_slots.swap(slots_);
    old_used.swap(used_);
    const std::size_t cap = old_used.empty() ? 16 : 2 * old_used.size();
    slots_.resize(cap);
    used_.assign(cap, 0);
    for (std::size_t j = 0; j < old_used.size(); j++)
      if (old_used[j]) {
        const std::size_t i = free_slot(old_slots[j].first);
        slots_[i] = std::move(old_slots[j]);
        used_[i] = 1;
      }
  }
  std::vector<value_type> slots_;
  std::vector<unsigned char> used_;
  std::size_t n_ = 0;
};
}  // namespace boost
