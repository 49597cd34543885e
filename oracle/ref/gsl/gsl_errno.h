/* GSL shim (test infrastructure, see gsl_shim.h): <gsl/gsl_errno.h> */
#include "gsl_shim.h"
