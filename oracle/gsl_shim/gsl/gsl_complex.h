/* stand-in for <gsl/gsl_complex.h>; see gsl_standin.h */
#include "gsl_standin.h"
