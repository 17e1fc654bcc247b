// <cuda_runtime.h> of the SIMT emulator (tools/simt): test infrastructure only.
#include "simt.h"
