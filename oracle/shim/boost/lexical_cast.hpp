#pragma once
#include <sstream>
namespace boost { template <class T, class S> T lexical_cast(S const& s) { std::istringstream is(s); T t; is >> t; return t; } }
