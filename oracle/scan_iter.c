/*
 * oracle/scan_iter.c -- CPU restatement of the reference's index-scan iteration, hnsw_gettuple
 * (embedding.c:285-370), over the flat-memory host.  TEST INFRASTRUCTURE ONLY: linked into both
 * checkers (oracle/_ref: the loop below drives the UNMODIFIED reference hnsw_search; oracle/_build: the
 * C restatement), never into the product.
 *
 * What the reference does, statement by statement:
 *   first call (so->curr == 0)            embedding.c:297-328   hnsw_search with efSearch; keep the 6-byte TIDs
 *                                                               (memcpy of sizeof(ItemPointerData) = the label's low 6
 *                                                               bytes: flags dropped); no_more = n < efSearch
 *   exhausted (curr >= n_results)         :329-366              no_more -> false; efSearch *= 2 IN PLACE (:334);
 *                                                               search again; `if (n_results <= so->n_results) return
 *                                                               false` compares the NEW search's count with the
 *                                                               ACCUMULATED count (:338); no_more = n < efSearch (:343);
 *                                                               pg_qsort the accumulated TIDs by ItemPointerCompare
 *                                                               (:354); append every new TID that bsearch does not find
 *                                                               (:357-363) -- the bsearch range is so->n_results, which
 *                                                               GROWS while appending, so the searched array is a sorted
 *                                                               prefix + an unsorted suffix.  A probe can therefore miss
 *                                                               a TID that IS in the prefix and the scan returns it a
 *                                                               second time; that quirk is part of the behaviour and is
 *                                                               restated with glibc's bsearch probe sequence.
 *   return so->results[so->curr++]        :367
 * ItemPointerCompare (PostgreSQL storage/itemptr.c): block number ((bi_hi << 16) | bi_lo) first, then ip_posid.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef float	 coord_t;
typedef uint64_t label_t;

extern long flat_search(void *fi, const coord_t *q, size_t efSearch, label_t *labels_out);

typedef struct
{
	unsigned char b[6];
} tid6; /* ItemPointerData: uint16 bi_hi, bi_lo, ip_posid */

typedef struct
{
	void	*host;
	coord_t *key;
	size_t	 dim, ef;	/* so->hnsw->meta.efSearch (per-scan copy, embedding.c:254) */
	tid6	*results;
	size_t	 n_results, curr;
	int		 no_more;
	uint64_t searches;	/* how many hnsw_search calls this scan made */
} FlatScan;

static int
tid_compare(const void *pa, const void *pb)
{
	const unsigned char *a = pa, *b = pb;
	uint16_t			 a_hi, a_lo, a_pos, b_hi, b_lo, b_pos;

	memcpy(&a_hi, a, 2); memcpy(&a_lo, a + 2, 2); memcpy(&a_pos, a + 4, 2);
	memcpy(&b_hi, b, 2); memcpy(&b_lo, b + 2, 2); memcpy(&b_pos, b + 4, 2);
	{
		uint32_t ba = ((uint32_t) a_hi << 16) | a_lo, bb = ((uint32_t) b_hi << 16) | b_lo;

		if (ba < bb) return -1;
		if (ba > bb) return 1;
	}
	if (a_pos < b_pos) return -1;
	if (a_pos > b_pos) return 1;
	return 0;
}

/* glibc bsearch (bits/stdlib-bsearch.h): l = 0, u = n; idx = (l + u) / 2; <0 -> u = idx; >0 -> l = idx + 1 */
static const void *
probe(const void *key, const tid6 *base, size_t n)
{
	size_t l = 0, u = n;

	while (l < u)
	{
		size_t idx = (l + u) / 2;
		int	   c = tid_compare(key, &base[idx]);

		if (c < 0) u = idx;
		else if (c > 0) l = idx + 1;
		else return &base[idx];
	}
	return NULL;
}

void *
flat_scan_begin(void *host, const coord_t *key, size_t dim, size_t efSearch)
{
	FlatScan *so = calloc(1, sizeof(FlatScan));

	if (!so) return NULL;
	so->host = host;
	so->dim = dim;
	so->ef = efSearch;
	so->key = malloc(dim * sizeof(coord_t));
	memcpy(so->key, key, dim * sizeof(coord_t));
	so->no_more = 1; /* embedding.c:258 */
	return so;
}

/* 1 = *tid_out holds the next TID (low 6 bytes of a label, flags zero); 0 = no more tuples; -1 = search failed */
int
flat_scan_next(void *scan, label_t *tid_out)
{
	FlatScan *so = scan;
	label_t	 *res;
	long	  n;

	if (so->curr == 0)
	{
		res = malloc((so->ef ? so->ef : 1) * sizeof(label_t));
		n = flat_search(so->host, so->key, so->ef, res);
		so->searches++;
		if (n < 0) { free(res); return -1; }
		free(so->results);
		so->results = malloc(((size_t) n ? (size_t) n : 1) * sizeof(tid6));
		so->n_results = (size_t) n;
		so->no_more = (size_t) n < so->ef;
		for (long i = 0; i < n; i++) memcpy(&so->results[i], &res[i], sizeof(tid6));
		free(res);
	}
	if (so->curr >= so->n_results)
	{
		if (so->no_more) return 0;
		so->ef *= 2;
		res = malloc(so->ef * sizeof(label_t));
		n = flat_search(so->host, so->key, so->ef, res);
		so->searches++;
		if (n < 0) { free(res); return -1; }
		if ((size_t) n <= so->n_results) { free(res); return 0; } /* "No new results found" */
		so->no_more = (size_t) n < so->ef;
		so->results = realloc(so->results, ((size_t) n + so->n_results) * sizeof(tid6));
		qsort(so->results, so->n_results, sizeof(tid6), tid_compare);
		for (long i = 0; i < n; i++)
			if (!probe(&res[i], so->results, so->n_results)) memcpy(&so->results[so->n_results++], &res[i], sizeof(tid6));
		free(res);
	}
	*tid_out = 0;
	memcpy(tid_out, &so->results[so->curr++], sizeof(tid6));
	return 1;
}

uint64_t
flat_scan_searches(void *scan)
{
	return ((FlatScan *) scan)->searches;
}

size_t
flat_scan_ef(void *scan)
{
	return ((FlatScan *) scan)->ef;
}

void
flat_scan_end(void *scan)
{
	FlatScan *so = scan;

	if (!so) return;
	free(so->key);
	free(so->results);
	free(so);
}
