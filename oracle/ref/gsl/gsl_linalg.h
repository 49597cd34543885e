/* GSL shim (test infrastructure, see gsl_shim.h): <gsl/gsl_linalg.h> */
#include "gsl_shim.h"
