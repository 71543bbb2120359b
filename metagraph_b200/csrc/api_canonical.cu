// Third compilation of k_align (kernels.cuh): DNA block layout, CANONICAL-mode graphs (`metagraph build --mode
// canonical`, dbg_aligner.cpp:224-226, 646-722). The BASIC-mode kernels (api.cu, api_generic.cu) compile this
// mode's branches out, so serving it costs them no code, registers or stack.
#define MGB_NARROW_ONLY 1
#define MGB_CANONICAL_ONLY 1
#define MGB_ALIGN_KERNEL_ONLY 1
#define MGB_KERNEL_NS kern_canon
#include "kernels.cuh"
