// stubs.cu -- entry points of include/sqgpu.h whose kernels are not written yet.
// Every function fails loudly; nothing here computes on the CPU.
#include "common.cuh"

#define NOT_YET(name)                                             \
    sq_set_error(name " is not implemented in this build yet");  \
    return SQ_E_LIMIT

extern "C" {
int sq_synth_illumina(sq_ctx *, uint8_t *, uint64_t, uint64_t, uint32_t, uint64_t, uint64_t *) { NOT_YET("sq_synth_illumina"); }
}
