/* stand-in for <gsl/gsl_spblas.h>; see gsl_standin.h */
#include "gsl_standin.h"
