/* tests/r_stub/R.h — TEST INFRASTRUCTURE: stand-in for R's <R.h>, see Rinternals.h in this directory. */
#ifndef PSY_R_STUB_R_H
#define PSY_R_STUB_R_H
#include <stddef.h>
#include "Rinternals.h"
#endif
