/* GSL shim (test infrastructure, see gsl_shim.h): <gsl/gsl_eigen.h> */
#include "gsl_shim.h"
