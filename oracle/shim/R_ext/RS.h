/* stand-in for <R_ext/RS.h> (oracle test infrastructure only): Fortran name mangling */
#ifndef F77_CALL
#define F77_CALL(x) x##_
#define F77_NAME(x) x##_
#endif
