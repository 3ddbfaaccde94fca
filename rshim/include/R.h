/* rshim: stand-in for <R.h>. */
#ifndef RSHIM_R_H
#define RSHIM_R_H
#include "Rinternals.h"
#endif
