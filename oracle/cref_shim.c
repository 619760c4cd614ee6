/* cref_shim.c -- compiles the UNMODIFIED reference C header (found through -I,
 * it carries non-static function definitions, c_superintervals.h:363-1089) into
 * oracle/_ref/libsi_cref.so so the same ctypes test driver can be pointed at
 * the reference's C ABI and at libsuperintervals_b200.so. TEST INFRASTRUCTURE. */
#include "c_superintervals.h"
