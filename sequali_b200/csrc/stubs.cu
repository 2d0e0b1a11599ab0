// stubs.cu -- entry points of include/sqgpu.h whose kernels are not written yet.
// Every function fails loudly; nothing here computes on the CPU.
#include "common.cuh"

#define NOT_YET(name)                                             \
    sq_set_error(name " is not implemented in this build yet");  \
    return SQ_E_LIMIT

extern "C" {
int sq_batch_from_bam(sq_ctx *, const uint8_t *, uint64_t, const uint64_t *, uint64_t, sq_batch **out, uint64_t *) { *out = nullptr; NOT_YET("sq_batch_from_bam"); }


int sq_pertile_create(sq_ctx *, sq_pertile **out) { *out = nullptr; NOT_YET("sq_pertile_create"); }
void sq_pertile_destroy(sq_pertile *) {}
int sq_pertile_add(sq_pertile *, sq_batch *) { NOT_YET("sq_pertile_add"); }
int sq_pertile_sync(sq_pertile *, sq_pertile_info *) { NOT_YET("sq_pertile_sync"); }
int sq_pertile_skipped_name(sq_pertile *, uint8_t *, uint64_t, uint64_t *) { NOT_YET("sq_pertile_skipped_name"); }
int sq_pertile_read(sq_pertile *, uint64_t *, double *, uint64_t *) { NOT_YET("sq_pertile_read"); }



int sq_nanostats_create(sq_ctx *, sq_nanostats **out) { *out = nullptr; NOT_YET("sq_nanostats_create"); }
void sq_nanostats_destroy(sq_nanostats *) {}
int sq_nanostats_add(sq_nanostats *, sq_batch *) { NOT_YET("sq_nanostats_add"); }
int sq_nanostats_sync(sq_nanostats *, sq_nanostats_info *) { NOT_YET("sq_nanostats_sync"); }
int sq_nanostats_skipped_name(sq_nanostats *, uint8_t *, uint64_t, uint64_t *) { NOT_YET("sq_nanostats_skipped_name"); }
int sq_nanostats_read(sq_nanostats *, sq_nanoinfo *) { NOT_YET("sq_nanostats_read"); }

int sq_insert_create(sq_ctx *, uint64_t, sq_insert **out) { *out = nullptr; NOT_YET("sq_insert_create"); }
void sq_insert_destroy(sq_insert *) {}
int sq_insert_add_pair(sq_insert *, sq_batch *, sq_batch *) { NOT_YET("sq_insert_add_pair"); }
int sq_insert_sync(sq_insert *, sq_insert_info *) { NOT_YET("sq_insert_sync"); }
int sq_insert_read_sizes(sq_insert *, uint64_t *) { NOT_YET("sq_insert_read_sizes"); }
int sq_insert_read_adapters(sq_insert *, int, uint8_t *, uint64_t *, uint64_t *) { NOT_YET("sq_insert_read_adapters"); }

int sq_synth_illumina(sq_ctx *, uint8_t *, uint64_t, uint64_t, uint32_t, uint64_t, uint64_t *) { NOT_YET("sq_synth_illumina"); }
}
