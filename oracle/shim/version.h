/* Shim for the build-generated version.h of PINC (makefile target "version").
 * TEST INFRASTRUCTURE ONLY: used to compile the reference sources into oracle/_ref. */
#ifndef PINC_SHIM_VERSION_H
#define PINC_SHIM_VERSION_H
#define VERSION "reference-under-shim"
#endif
