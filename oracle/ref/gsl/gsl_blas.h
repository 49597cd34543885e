/* GSL shim (test infrastructure, see gsl_shim.h): <gsl/gsl_blas.h> */
#include "gsl_shim.h"
