// stand-in for boost::math::round (half away from zero for finite arguments, which is std::round)
#pragma once
#include <cmath>
namespace boost { namespace math { template <typename T> inline T round(T const& v) { return std::round(v); } } }
