/* forwards to the shim one directory up; this directory exists so the REFERENCE's own nsparse.h
 * (not ours) is picked up when oracle/Makefile compiles the reference host code. */
#include "../helper_cuda.h"
