/* Stub of the few R API entry points bindings/R/src/Rwrapper_b200.c uses - syntax / type check only (no R in the image). */
#ifndef R_STUB_H
#define R_STUB_H
#include <stddef.h>
#include <stdlib.h>
void error(const char *fmt, ...);
#define R_Calloc(n, T) ((T*) calloc((size_t) (n), sizeof(T)))
#define R_Free(p) free(p)
#endif
