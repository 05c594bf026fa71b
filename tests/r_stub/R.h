/* stand-in for <R.h> (see Rinternals.h in this directory) */
#ifndef ZZ_R_STUB_R_H
#define ZZ_R_STUB_R_H
#include <stdint.h>
#include <stdlib.h>
#endif
