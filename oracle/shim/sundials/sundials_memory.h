/* Test-infrastructure shim (NOT product code): stands in for the SUNDIALS 6.2 header
 * <sundials/sundials_memory.h> so that /root/reference/src/{utilities.cpp,euler3D.hpp} compile unmodified
 * in a container without SUNDIALS.  Everything lives in shim_core.h. */
#include "../shim_core.h"
