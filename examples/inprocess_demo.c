/*
 * inprocess_demo.c -- the drop-in boundary in plain C, in-process variant: libpgemb_b200.so on the link line where the
 * reference has `hnswalg.o distfunc.o` (reference Makefile:6).  The reference's call sites are kept (hnsw_bind_point at
 * embedding.c:695, hnsw_search at :317, free at :327); the HBM mirror is kept current with the bulk entry points of
 * include/pgemb_b200.h (INTEGRATION.md sections 1-4).  Replays test/sql/knn.sql and prints the rows in index order
 * (test/expected/knn.out:13-20 for `<->`).  Without a CUDA device it fails loudly: there is no CPU path behind these symbols.
 *
 *     gcc -Iinclude examples/inprocess_demo.c -Lpg_embedding_b200 -lpgemb_b200 -o inprocess_demo && ./inprocess_demo cosine
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pgemb_b200.h"

int main(int argc, char **argv)
{
	const char		 *metric = argc > 1 ? argv[1] : "l2";
	const dist_func_t dist = !strcmp(metric, "cosine") ? DIST_COSINE : (!strcmp(metric, "manhattan") ? DIST_MANHATTAN : DIST_L2);
	static const float rows[4][3] = {{0, 1, 2}, {1, 2, 3}, {1, 1, 1}, {1, 2, 4}};
	const float		   query[3] = {3, 3, 3};

	hnsw_init_dist_func(); /* _PG_init, embedding.c:150 */
	if (pgemb_device_count() < 1)
	{
		fprintf(stderr, "no CUDA device: the hnsw hot path has no CPU fallback\n");
		return 3;
	}
	PgembHostIndex h; /* the reference's HnswIndex prefix: metadata first (embedding.c:65-75) */
	if (pgemb_meta_init(&h.meta, 3, 3, 16, 64, dist) != PGEMB_OK || pgemb_index_create(&h.meta, 64, 0, &h.dev) != PGEMB_OK)
	{
		fprintf(stderr, "%s\n", pgemb_last_error());
		return 1;
	}
	unsigned char *rec = calloc(1, h.meta.size_data_per_element);
	idx_t		  *links = calloc(h.meta.maxM + 1, sizeof(idx_t));
	for (idx_t cur = 0; cur < 4; cur++)
	{
		memset(rec, 0, h.meta.size_data_per_element); /* stored with zeroed links (embedding.c:619-621) */
		memcpy(rec + h.meta.offset_data, rows[cur], h.meta.data_size);
		const label_t label = (label_t) (cur + 1) << 32;
		memcpy(rec + h.meta.offset_label, &label, sizeof(label));
		if (pgemb_index_append_records(h.dev, 1, rec, h.meta.size_data_per_element) != PGEMB_OK || !hnsw_bind_point(&h.meta, rows[cur], cur))
		{
			fprintf(stderr, "HNSW index insert failed: %s\n", pgemb_last_error());
			return 1;
		}
		/* write-back of the new node's list to its page */
		if (pgemb_index_get_links(h.dev, cur, 1, links) != PGEMB_OK) return 1;
	}
	size_t	 n_results = 0;
	label_t *results = NULL;
	if (!hnsw_search(&h.meta, query, &n_results, &results))
	{
		fprintf(stderr, "HNSW index search failed: %s\n", pgemb_last_error());
		return 1;
	}
	for (size_t i = 0; i < n_results; i++)
	{
		const unsigned pos = (unsigned) (results[i] >> 32);
		printf("{%g,%g,%g}\n", rows[pos - 1][0], rows[pos - 1][1], rows[pos - 1][2]);
	}
	free(results); /* embedding.c:327 */
	free(rec);
	free(links);
	pgemb_index_destroy(h.dev);
	return 0;
}
