#pragma once
#include <type_traits>
namespace boost { template <class T> struct is_pod : std::integral_constant<bool, std::is_standard_layout<T>::value && std::is_trivial<T>::value> {}; }
