// Compile-time lattice size, as in the reference (CMakeLists.txt:17-20 -> include/config.h.in):
// pass -DNS=<Nx> -DNT=<Nt>; the executable is named SM_${NS}x${NT}.
#ifndef SM_HOST_CONFIG_H
#define SM_HOST_CONFIG_H
#ifndef NS
#define NS 64
#endif
#ifndef NT
#define NT 64
#endif
#endif
