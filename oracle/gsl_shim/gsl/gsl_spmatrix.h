/* stand-in for <gsl/gsl_spmatrix.h>; see gsl_standin.h */
#include "gsl_standin.h"
