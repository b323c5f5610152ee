// TEST INFRASTRUCTURE (oracle/_ref build only): the one Boost facility the narrow-phase headers use.
#pragma once
namespace boost { namespace math { namespace constants {
template <class T> inline constexpr T pi() { return static_cast<T>(3.141592653589793238462643383279502884L); }
}}}
