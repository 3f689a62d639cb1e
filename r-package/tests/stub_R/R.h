/* stand-in for R.h (see Rinternals.h in this directory) */
#ifndef STUB_R_H
#define STUB_R_H
#include <stdlib.h>
#endif
