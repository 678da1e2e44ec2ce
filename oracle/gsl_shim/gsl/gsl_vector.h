/* stand-in for <gsl/gsl_vector.h>; see gsl_standin.h */
#include "gsl_standin.h"
