#include "CudaContext.h"
