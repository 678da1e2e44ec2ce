/* stand-in for <gsl/gsl_complex_math.h>; see gsl_standin.h */
#include "gsl_standin.h"
