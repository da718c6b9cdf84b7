/* oracle/io_bridge_stubs.c -- TEST INFRASTRUCTURE ONLY.
 * _ref/librebx_io_bridge.so = the reference's src/output.c + src/input.c (binary save/load, out of
 * scope for this build) compiled unmodified and linked AGAINST libreboundx_b200.so, to test the
 * drop-in claim of INTEGRATION.md section 3: untouched reference objects working on top of this
 * library's registry / parameter / force API and struct layouts.  The operator half of the
 * reference API is not provided by the library; files written by the test contain no operators. */
#include <stddef.h>
void* rebx_load_operator(void* rebx, const char* name){ (void)rebx; (void)name; return NULL; }
void* rebx_get_operator(void* rebx, const char* name){ (void)rebx; (void)name; return NULL; }
int rebx_add_operator_step(void* rebx, void* op, double dt_fraction, int timing){ (void)rebx; (void)op; (void)dt_fraction; (void)timing; return 0; }
