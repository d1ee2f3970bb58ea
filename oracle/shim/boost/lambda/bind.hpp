#pragma once
namespace boost { namespace lambda { template <class F, class A> auto bind(F f, A a) { return [=]() { f(a); }; } } }
