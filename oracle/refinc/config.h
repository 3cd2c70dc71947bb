/* Hand-written stand-in for the autoconf-generated config.h of
 * sphinxbase / pocketsphinx.  Test infrastructure only. */
#define HAVE_LONG_LONG 1
#define SIZEOF_LONG 8
#define SIZEOF_LONG_LONG 8
#define HAVE_UNISTD_H 1
#define HAVE_STDINT_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_SYS_STAT_H 1
#define HAVE_POPEN 1
#define HAVE_SNPRINTF 1
#define HAVE_PERROR 1
#define HAVE_LIBM 1
#define HAVE_PTHREAD_H 1
#define RETSIGTYPE void
#define AD_BACKEND_NONE 1
