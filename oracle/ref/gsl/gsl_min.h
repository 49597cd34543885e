/* GSL shim (test infrastructure, see gsl_shim.h): <gsl/gsl_min.h> */
#include "gsl_shim.h"
