/* Hand-written stand-in for the autoconf-generated sphinx_config.h
 * (template: sphinxbase/include/sphinx_config.h.in).  Float build
 * (no FIXED_POINT), 64-bit little-endian Linux.  Test infrastructure only. */
#define AD_BACKEND_NONE 1
#define SIZEOF_LONG 8
#define HAVE_LONG_LONG 1
#define SIZEOF_LONG_LONG 8
