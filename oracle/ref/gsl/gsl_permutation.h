/* GSL shim (test infrastructure, see gsl_shim.h): <gsl/gsl_permutation.h> */
#include "gsl_shim.h"
