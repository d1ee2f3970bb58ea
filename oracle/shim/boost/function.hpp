#pragma once
#include <functional>
namespace boost { template <class S> using function = std::function<S>; }
