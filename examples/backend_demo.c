/*
 * backend_demo.c -- what the algorithm-facing part of a Postgres backend does, in plain C against libpgemb_client.so:
 * the call sites of reference embedding.c kept as they are (hnsw_bind_point at :695, hnsw_search at :317, free at :327),
 * plus the three mirror calls INTEGRATION.md section 7 adds.  It replays test/sql/knn.sql: four rows, the query {3,3,3},
 * and prints the rows in index order -- test/expected/knn.out:13-20 for `<->`.
 *
 *     gcc -Iinclude examples/backend_demo.c -Lpg_embedding_b200 -lpgemb_client -o backend_demo
 *     PGEMB_SIDECAR_SHM=/pgemb ./backend_demo l2        (a pgemb_sidecar must serve /pgemb)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pgemb_client.h"

/* the derived fields of HnswMetadata exactly as hnsw_get_index computes them (embedding.c:222-235) */
static void meta_init(HnswMetadata *m, size_t dims, size_t M, size_t efc, size_t efs, dist_func_t dist)
{
	memset(m, 0, sizeof(*m));
	m->dim = dims;
	m->M = M;
	m->maxM = M * 2;
	m->data_size = dims * sizeof(coord_t);
	m->offset_data = (m->maxM + 1) * sizeof(idx_t);
	m->offset_label = m->offset_data + m->data_size;
	m->size_data_per_element = m->offset_label + sizeof(label_t);
	m->elems_per_page = (8192 - 24 - 4) / (m->size_data_per_element + 4);
	m->efConstruction = efc;
	m->efSearch = efs;
	m->dist_func = dist;
}

int main(int argc, char **argv)
{
	const char		 *metric = argc > 1 ? argv[1] : "l2";
	const dist_func_t dist = !strcmp(metric, "cosine") ? DIST_COSINE : (!strcmp(metric, "manhattan") ? DIST_MANHATTAN : DIST_L2);
	static const float rows[4][3] = {{0, 1, 2}, {1, 2, 3}, {1, 1, 1}, {1, 2, 4}};
	const float		   query[3] = {3, 3, 3};

	hnsw_init_dist_func(); /* _PG_init (embedding.c:150): connects to $PGEMB_SIDECAR_SHM */

	PgembClientIndex h; /* HnswIndex: metadata first, then what identifies the relation */
	meta_init(&h.meta, 3, 3, 16, 64, dist);
	h.rel_key = 0x100000000ull | (uint64_t) dist;
	size_t have = 0;
	if (pgemb_client_attach(&h, 64, &have, NULL) != 0)
	{
		fprintf(stderr, "attach: %s\n", pgemb_client_last_error());
		return 1;
	}
	unsigned char *rec = calloc(1, h.meta.size_data_per_element);
	for (idx_t cur = (idx_t) have; cur < 4; cur++)
	{
		/* hnsw_add_point (embedding.c:606-701): the record is stored with zeroed links (:619-621) ... */
		memset(rec, 0, h.meta.size_data_per_element);
		memcpy(rec + h.meta.offset_data, rows[cur], h.meta.data_size);
		const label_t label = (label_t) (cur + 1) << 32; /* ItemPointer (0, cur+1) */
		memcpy(rec + h.meta.offset_label, &label, sizeof(label));
		if (pgemb_client_append_records(&h, 1, rec, h.meta.size_data_per_element) != 0)
		{
			fprintf(stderr, "append: %s\n", pgemb_client_last_error());
			return 1;
		}
		/* ... and bound: the reference's call, unchanged (:695) */
		if (!hnsw_bind_point(&h.meta, rows[cur], cur))
		{
			fprintf(stderr, "HNSW index insert failed\n");
			return 1;
		}
	}
	free(rec);

	/* hnsw_gettuple (embedding.c:317): the reference's call, unchanged */
	size_t	 n_results = 0;
	label_t *results = NULL;
	if (!hnsw_search(&h.meta, query, &n_results, &results))
	{
		fprintf(stderr, "HNSW index search failed: %s\n", pgemb_client_last_error());
		return 1;
	}
	for (size_t i = 0; i < n_results; i++)
	{
		const unsigned pos = (unsigned) (results[i] >> 32);
		printf("{%g,%g,%g}\n", rows[pos - 1][0], rows[pos - 1][1], rows[pos - 1][2]);
	}
	free(results); /* embedding.c:327 */
	return 0;
}
