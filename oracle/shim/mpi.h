/* Test-infrastructure shim (NOT product code): in-process stand-in for <mpi.h>.
 * See shim_core.h. */
#include "shim_core.h"
