// TEST INFRASTRUCTURE (oracle/_ref build only): stand-in for the generated deprecated.hh.
#ifndef HPP_FCL_DEPRECATED_HH
#define HPP_FCL_DEPRECATED_HH
#define HPP_FCL_DEPRECATED [[deprecated]]
#define HPP_FCL_DEPRECATED_MESSAGE(message) [[deprecated(#message)]]
#ifndef HPP_FCL_PRAGMA
#define HPP_FCL_PRAGMA(X) _Pragma(#X)
#endif
#define HPP_FCL_DEPRECATED_HEADER(MSG) HPP_FCL_PRAGMA(GCC warning MSG)
#endif
