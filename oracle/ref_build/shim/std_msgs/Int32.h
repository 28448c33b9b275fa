// Shim: forwards to refshim.h (see oracle/ref_build/README.md).  Not a copy of any third-party header.
#include "refshim.h"
