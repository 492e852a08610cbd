/* stand-in for <R.h> (oracle test infrastructure only) */
