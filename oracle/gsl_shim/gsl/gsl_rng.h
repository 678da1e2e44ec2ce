/* stand-in for <gsl/gsl_rng.h>; see gsl_standin.h */
#include "gsl_standin.h"
