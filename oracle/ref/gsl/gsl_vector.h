/* GSL shim (test infrastructure, see gsl_shim.h): <gsl/gsl_vector.h> */
#include "gsl_shim.h"
