#pragma once
#include <iterator>
namespace boost {
template <class R> auto begin(R const& r) -> decltype(std::begin(r)) { return std::begin(r); }
template <class R> auto end(R const& r) -> decltype(std::end(r)) { return std::end(r); }
template <class R> size_t size(R const& r) { return size_t(std::distance(std::begin(r), std::end(r))); }
template <class R> struct range_const_iterator { typedef decltype(std::begin(std::declval<R const&>())) type; };
}
