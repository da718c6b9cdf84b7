/* Stand-in for REBOUND's src/transformations.h (included by the reference's src/gr.c:61).
 * The prototypes live in the shim rebound.h. */
#ifndef REBOUND_SHIM_TRANSFORMATIONS_H
#define REBOUND_SHIM_TRANSFORMATIONS_H
#include "rebound.h"
#endif
