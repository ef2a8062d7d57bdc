// attention.cu — fused attention (bring-up stub, filled in next)
#include "common.cuh"
