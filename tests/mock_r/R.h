/* Mock of R's C API for tests/test_shim.py: just enough of R.h / Rinternals.h / R_ext/Rdynload.h to compile and run
 * r-shim/src/shim.c without R (the build container has none).  SEXPs are malloc'ed structs; nothing is garbage collected,
 * PROTECT / UNPROTECT only keep a balance counter that the driver checks.  TEST INFRASTRUCTURE, not part of the product. */
#ifndef MOCK_R_H
#define MOCK_R_H
#include <stddef.h>
#include <stdint.h>
#endif
