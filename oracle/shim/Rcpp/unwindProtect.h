/* stand-in: everything lives in the shim Rcpp.h (oracle test infrastructure only) */
#include "../Rcpp.h"
