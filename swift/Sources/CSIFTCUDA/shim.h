// Umbrella header of the CSIFTCUDA system-library target (was: include/MetalShaders.h:6-11).
#include "siftcuda.h"
