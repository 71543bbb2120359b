// Second compilation of the kernels (kernels.cuh) for indexes in the alphabet-generic layout (protein,
// BASELINE configs[3]); api.cu holds the DNA-only set and all host code.
#define MGB_WIDE_ONLY 1
#define MGB_BASIC_ONLY 1
#define MGB_KERNEL_NS kern_any
#include "kernels.cuh"
