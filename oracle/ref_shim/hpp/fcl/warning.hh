// TEST INFRASTRUCTURE (oracle/_ref build only): stand-in for the generated warning.hh.
#ifndef HPP_FCL_WARNING_HH
#define HPP_FCL_WARNING_HH
#define HPP_FCL_WARN_STRINGISE_IMPL(x) #x
#define HPP_FCL_WARN_STRINGISE(x) HPP_FCL_WARN_STRINGISE_IMPL(x)
#define HPP_FCL_WARN(exp) ("WARNING: " exp)
#endif
