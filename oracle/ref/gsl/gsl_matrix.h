/* GSL shim (test infrastructure, see gsl_shim.h): <gsl/gsl_matrix.h> */
#include "gsl_shim.h"
