/* GSL shim (test infrastructure, see gsl_shim.h): <gsl/gsl_randist.h> */
#include "gsl_shim.h"
