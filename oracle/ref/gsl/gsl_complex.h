/* GSL shim (test infrastructure, see gsl_shim.h): <gsl/gsl_complex.h> */
#include "gsl_shim.h"
