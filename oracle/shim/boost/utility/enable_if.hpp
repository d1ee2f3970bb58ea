#pragma once
#include <type_traits>
namespace boost { template <class C, class T = void> struct enable_if : std::enable_if<C::value, T> {}; template <class C, class T = void> struct disable_if : std::enable_if<!C::value, T> {}; }
