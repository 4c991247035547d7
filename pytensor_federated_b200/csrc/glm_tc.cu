// placeholder until the tcgen05 kernel lands (next commit)
#include "fed_comm.cuh"
#include "models.h"
extern "C" int b200_launch_glm_tc(const FedComm*, const GlmSegment*, const GlmParams*, const void*, int, cudaStream_t) { return -1; }
extern "C" int b200_glm_tc_prepare(const GlmSegment*, int, const GlmParams*, void**) { return -1; }
