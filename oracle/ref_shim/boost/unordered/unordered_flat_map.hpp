// Stand-in for Boost.Unordered (not available offline), for oracle/_ref only: MaterialHelper keeps two lookup caches in
// boost::unordered_flat_map; std::unordered_map has the interface it uses (find / operator[] / clear), and no result
// depends on iteration order.
#pragma once
#include <functional>
#include <memory>
#include <unordered_map>
namespace boost {
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>, class A = std::allocator<std::pair<const K, V>>>
using unordered_flat_map = std::unordered_map<K, V, H, E, A>;
}  // namespace boost
