// pxz_host.h — host-side internals of libpixlzr_b200 (not part of the public ABI).
#pragma once

#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "pxz_internal.h"

namespace pxz {

// tables.cpp
bool build_axis_table(uint32_t n_in, uint32_t n_out, int filter, std::vector<uint32_t>* pool, AxisTab* tab);
bool build_axis_table_fir(uint32_t n_in, uint32_t n_out, int alg, std::vector<uint32_t>* pool, AxisTab* tab);
void build_level_thresholds(LevelThresholds* out);

// nccl_dyn.cpp — NCCL is loaded lazily (dlopen) so the library has no link-time dependency on it
struct NcclApi;
const NcclApi* nccl_api(std::string* err);
int nccl_get_unique_id(uint8_t id[PXZ_COMM_ID_BYTES], std::string* err);
int nccl_comm_init(void** comm, int nranks, int rank, const uint8_t id[PXZ_COMM_ID_BYTES], std::string* err);
int nccl_allreduce_min_f32(void* comm, float* dev_buf, size_t count, cudaStream_t stream, std::string* err);
void nccl_comm_destroy(void* comm);

}  // namespace pxz
