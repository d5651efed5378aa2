/* Shim so that oracle/flat_host.c and oracle/hnsw_oracle.c compile WITHOUT /root/reference:
 * the types come from this repo's own C-ABI header (layout-identical to reference embedding.h:17-42)
 * and the six storage callbacks the algorithm calls upward are declared here
 * (reference embedding.h:44, :48-53).  TEST INFRASTRUCTURE ONLY. */
#pragma once
#include "pgemb_b200.h"

bool hnsw_begin_read(HnswMetadata* meta, idx_t idx, idx_t** indexes, coord_t** coords, label_t* label);
void hnsw_end_read(HnswMetadata* meta);
void hnsw_begin_write(HnswMetadata* meta, idx_t idx, idx_t** indexes, coord_t** coords, label_t* label);
void hnsw_end_write(HnswMetadata* meta);
void hnsw_prefetch(HnswMetadata* meta, idx_t idx);
