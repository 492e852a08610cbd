/* stand-in for <Rinternals.h> (oracle test infrastructure only) */
