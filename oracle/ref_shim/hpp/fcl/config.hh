// TEST INFRASTRUCTURE (oracle/_ref build only).  Hand-written stand-in for the header that the
// reference's build system generates from cmake/config.hh.cmake: version macros and symbol
// visibility, nothing else.  Version numbers: package.xml of /root/reference at the time of writing.
#ifndef HPP_FCL_CONFIG_HH
#define HPP_FCL_CONFIG_HH
#define HPP_FCL_VERSION_UNKNOWN_TAG 0
#define HPP_FCL_VERSION "3.0.0"
#define HPP_FCL_MAJOR_VERSION 3
#define HPP_FCL_MINOR_VERSION 0
#define HPP_FCL_PATCH_VERSION 0
#define HPP_FCL_VERSION_AT_LEAST(major, minor, patch)                                   \
  (HPP_FCL_MAJOR_VERSION > major ||                                                       \
   (HPP_FCL_MAJOR_VERSION >= major &&                                                     \
    (HPP_FCL_MINOR_VERSION > minor || (HPP_FCL_MINOR_VERSION >= minor && HPP_FCL_PATCH_VERSION >= patch))))
#define HPP_FCL_VERSION_AT_MOST(major, minor, patch)                                    \
  (HPP_FCL_MAJOR_VERSION < major ||                                                       \
   (HPP_FCL_MAJOR_VERSION <= major &&                                                     \
    (HPP_FCL_MINOR_VERSION < minor || (HPP_FCL_MINOR_VERSION <= minor && HPP_FCL_PATCH_VERSION <= patch))))
#define HPP_FCL_DLLIMPORT __attribute__((visibility("default")))
#define HPP_FCL_DLLEXPORT __attribute__((visibility("default")))
#define HPP_FCL_DLLLOCAL __attribute__((visibility("hidden")))
#define HPP_FCL_EXPLICIT_INSTANTIATION_DECLARATION_DLLIMPORT __attribute__((visibility("default")))
#define HPP_FCL_EXPLICIT_INSTANTIATION_DECLARATION_DLLEXPORT __attribute__((visibility("default")))
#define HPP_FCL_EXPLICIT_INSTANTIATION_DEFINITION_DLLIMPORT
#define HPP_FCL_EXPLICIT_INSTANTIATION_DEFINITION_DLLEXPORT
#define HPP_FCL_DLLAPI HPP_FCL_DLLEXPORT
#define HPP_FCL_LOCAL HPP_FCL_DLLLOCAL
#define HPP_FCL_EXPLICIT_INSTANTIATION_DECLARATION_DLLAPI HPP_FCL_EXPLICIT_INSTANTIATION_DECLARATION_DLLEXPORT
#define HPP_FCL_EXPLICIT_INSTANTIATION_DEFINITION_DLLAPI HPP_FCL_EXPLICIT_INSTANTIATION_DEFINITION_DLLEXPORT
#endif
