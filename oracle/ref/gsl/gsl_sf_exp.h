/* GSL shim (test infrastructure, see gsl_shim.h): <gsl/gsl_sf_exp.h> */
#include "gsl_shim.h"
