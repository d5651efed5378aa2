/*
 * flat_host.c -- flat-memory storage host + test harness for the HNSW hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (pg_embedding_b200/) may link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do.
 *
 * What it is: the reference algorithm (hnswalg.cpp) never touches memory directly; it calls six
 * storage callbacks that the Postgres glue implements on 8 KB buffer pages
 * (reference embedding.h:44-53, embedding.c:704-850, :948-953).  This file implements the same six
 * callbacks over ONE flat malloc'd array of records that use the reference's record layout
 *     [ u32 count | u32 links[maxM] | f32 coords[dim] | u64 label ]      (embedding.c:224-228, :619-621)
 * so the algorithm above it (either the unmodified reference objects -> oracle/_ref/libpgemb_ref.so,
 * or the C restatement hnsw_oracle.c -> oracle/_build/libpgemb_port.so) runs without Postgres.
 * The handle is a struct whose FIRST member is HnswMetadata, exactly like the reference's HnswIndex
 * (embedding.c:65-75); callbacks down-cast the meta pointer the same way (embedding.c:706).
 *
 * It also provides the harness the tests/bench need: append+bind (the hnsw_add_point contract,
 * embedding.c:606-701), raw record loading (to search a graph built elsewhere), a multi-threaded
 * timed search loop, and per-query work counters (node-vector reads, link-list reads, link words)
 * from which SURVEY.md section 8(d)'s algorithmic bytes are derived.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stddef.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <time.h>

#include "embedding.h" /* the reference's own boundary header (-I/root/reference) or our mirror */

#define DELETED_FLAG 1 /* embedding.c:44 */

typedef struct
{
	uint64_t coord_reads; /* begin_read(.., coords != NULL)  ~ one node vector fetched            */
	uint64_t link_reads;  /* begin_read(.., indexes != NULL) ~ one node expansion (hnswalg.cpp:76) */
	uint64_t link_words;  /* sum over expansions of (1 + count) u32 words                         */
	uint64_t label_reads;
	uint64_t writes;
} FlatCounters;

typedef struct
{
	HnswMetadata meta; /* MUST be first (down-cast, embedding.c:65-75, :706) */
	char        *records;
	size_t       n_items;
	size_t       capacity;
} FlatIndex;

static __thread FlatCounters tl_counters;

/* ------------------------------------------------------------------------------------------
 * The six storage callbacks (reference embedding.h:44, :48-53).
 * ---------------------------------------------------------------------------------------- */

static inline char *
flat_record(FlatIndex *fi, idx_t idx)
{
	return fi->records + (size_t) idx * fi->meta.size_data_per_element;
}

/* embedding.c:704-757: returns false only when idx is beyond the last stored item. */
bool
hnsw_begin_read(HnswMetadata *meta, idx_t idx, idx_t **indexes, coord_t **coords, label_t *label)
{
	FlatIndex *fi = (FlatIndex *) meta;
	char	  *item;

	if ((size_t) idx >= fi->n_items)
		return false;
	item = flat_record(fi, idx);
	if (indexes)
	{
		*indexes = (idx_t *) item;
		tl_counters.link_reads += 1;
		tl_counters.link_words += 1 + (uint64_t) ((idx_t *) item)[0];
	}
	if (coords)
	{
		*coords = (coord_t *) (item + meta->offset_data);
		tl_counters.coord_reads += 1;
	}
	if (label)
	{
		memcpy(label, item + meta->offset_label, sizeof(*label));
		tl_counters.label_reads += 1;
	}
	return true;
}

void
hnsw_end_read(HnswMetadata *meta)
{
	(void) meta;
}

/* embedding.c:769-820 */
void
hnsw_begin_write(HnswMetadata *meta, idx_t idx, idx_t **indexes, coord_t **coords, label_t *label)
{
	FlatIndex *fi = (FlatIndex *) meta;
	char	  *item = flat_record(fi, idx);

	if ((size_t) idx >= fi->n_items)
	{
		fprintf(stderr, "flat_host: begin_write beyond end (%u >= %zu)\n", idx, fi->n_items);
		abort();
	}
	tl_counters.writes += 1;
	if (indexes)
		*indexes = (idx_t *) item;
	if (coords)
		*coords = (coord_t *) (item + meta->offset_data);
	if (label)
		memcpy(label, item + meta->offset_label, sizeof(*label));
}

void
hnsw_end_write(HnswMetadata *meta)
{
	(void) meta;
}

void
hnsw_prefetch(HnswMetadata *meta, idx_t idx)
{
	FlatIndex *fi = (FlatIndex *) meta;

	if ((size_t) idx < fi->n_items)
		__builtin_prefetch(flat_record(fi, idx) + meta->offset_data);
}

/* embedding.c:948-953: flags live in the top 16 bits of the 8-byte label (HnswLabel, :50-56). */
bool
hnsw_is_deleted(label_t label)
{
	return ((label >> 48) & DELETED_FLAG) != 0;
}

/* ------------------------------------------------------------------------------------------
 * Harness API (ctypes-friendly, plain C types).
 * ---------------------------------------------------------------------------------------- */

static pthread_once_t init_once = PTHREAD_ONCE_INIT;

static void
init_dist(void)
{
	hnsw_init_dist_func(); /* embedding.c:150 calls this once from _PG_init */
}

/* Mirrors hnsw_get_index's derivation of the layout from options (embedding.c:222-235), except
 * elems_per_page which is meaningless for a flat array (set to capacity). */
FlatIndex *
flat_create(size_t dim, size_t M, size_t efConstruction, size_t efSearch, int dist_func, size_t capacity)
{
	FlatIndex *fi = (FlatIndex *) calloc(1, sizeof(FlatIndex));

	pthread_once(&init_once, init_dist);
	if (!fi)
		return NULL;
	fi->meta.dim = dim;
	fi->meta.M = M;
	fi->meta.maxM = M * 2;
	fi->meta.data_size = dim * sizeof(coord_t);
	fi->meta.offset_data = (fi->meta.maxM + 1) * sizeof(idx_t);
	fi->meta.offset_label = fi->meta.offset_data + fi->meta.data_size;
	fi->meta.size_data_per_element = fi->meta.offset_label + sizeof(label_t);
	fi->meta.elems_per_page = capacity ? capacity : 1;
	fi->meta.efConstruction = efConstruction;
	fi->meta.efSearch = efSearch;
	fi->meta.dist_func = (dist_func_t) dist_func;
	fi->meta.enterpoint_node = 0;
	fi->capacity = capacity;
	fi->n_items = 0;
	fi->records = (char *) calloc(capacity ? capacity : 1, fi->meta.size_data_per_element);
	if (!fi->records)
	{
		free(fi);
		return NULL;
	}
	return fi;
}

void
flat_destroy(FlatIndex *fi)
{
	if (fi)
	{
		free(fi->records);
		free(fi);
	}
}

size_t flat_size(FlatIndex *fi) { return fi->n_items; }
size_t flat_record_size(FlatIndex *fi) { return fi->meta.size_data_per_element; }
char  *flat_records(FlatIndex *fi) { return fi->records; }
void   flat_truncate(FlatIndex *fi) { memset(fi->records, 0, fi->n_items * fi->meta.size_data_per_element); fi->n_items = 0; }

static int
flat_store(FlatIndex *fi, const coord_t *coords, label_t label, const idx_t *links)
{
	char *item;

	if (fi->n_items >= fi->capacity)
		return -1;
	item = fi->records + fi->n_items * fi->meta.size_data_per_element;
	if (links)
		memcpy(item, links, fi->meta.offset_data);
	else
		memset(item, 0, fi->meta.offset_data);
	memcpy(item + fi->meta.offset_data, coords, fi->meta.data_size);
	memcpy(item + fi->meta.offset_label, &label, sizeof(label));
	fi->n_items += 1;
	return 0;
}

/* The hnsw_add_point contract (embedding.c:606-701): store [zeroed links | coords | label], derive the
 * dense id, then hnsw_bind_point.  Returns 0 on success. */
int
flat_add(FlatIndex *fi, const coord_t *coords, label_t label)
{
	idx_t cur;

	if (flat_store(fi, coords, label, NULL) != 0)
		return -1;
	cur = (idx_t) (fi->n_items - 1);
	return hnsw_bind_point(&fi->meta, coords, cur) ? 0 : -2;
}

/* Append n records without binding: used to load a graph built elsewhere (links may be NULL). */
int
flat_append_raw(FlatIndex *fi, size_t n, const coord_t *coords, const label_t *labels, const idx_t *links)
{
	for (size_t i = 0; i < n; i++)
	{
		if (flat_store(fi, coords + i * fi->meta.dim, labels ? labels[i] : (label_t) (fi->n_items),
					   links ? links + i * (fi->meta.maxM + 1) : NULL) != 0)
			return -1;
	}
	return 0;
}

/* Append n records already in the reference's AoS layout (e.g. exported from the GPU index). */
int
flat_append_records(FlatIndex *fi, size_t n, const char *recs, size_t stride)
{
	if (fi->n_items + n > fi->capacity || stride < fi->meta.size_data_per_element)
		return -1;
	if (stride == fi->meta.size_data_per_element)
		memcpy(fi->records + fi->n_items * stride, recs, n * stride);
	else
		for (size_t i = 0; i < n; i++)
			memcpy(fi->records + (fi->n_items + i) * fi->meta.size_data_per_element, recs + i * stride,
				   fi->meta.size_data_per_element);
	fi->n_items += n;
	return 0;
}

/* Bind an already stored record (used when records were appended raw with zero links). */
int
flat_bind(FlatIndex *fi, idx_t idx)
{
	return hnsw_bind_point(&fi->meta, (coord_t *) (flat_record(fi, idx) + fi->meta.offset_data), idx) ? 0 : -2;
}

void
flat_get_links(FlatIndex *fi, size_t first, size_t n, idx_t *out)
{
	for (size_t i = 0; i < n; i++)
		memcpy(out + i * (fi->meta.maxM + 1), flat_record(fi, (idx_t) (first + i)), fi->meta.offset_data);
}

void
flat_set_links(FlatIndex *fi, size_t first, size_t n, const idx_t *in)
{
	for (size_t i = 0; i < n; i++)
		memcpy(flat_record(fi, (idx_t) (first + i)), in + i * (fi->meta.maxM + 1), fi->meta.offset_data);
}

void
flat_get_coords(FlatIndex *fi, size_t first, size_t n, coord_t *out)
{
	for (size_t i = 0; i < n; i++)
		memcpy(out + i * fi->meta.dim, flat_record(fi, (idx_t) (first + i)) + fi->meta.offset_data, fi->meta.data_size);
}

void
flat_get_labels(FlatIndex *fi, size_t first, size_t n, label_t *out)
{
	for (size_t i = 0; i < n; i++)
		memcpy(out + i, flat_record(fi, (idx_t) (first + i)) + fi->meta.offset_label, sizeof(label_t));
}

void
flat_set_label(FlatIndex *fi, idx_t idx, label_t label)
{
	memcpy(flat_record(fi, idx) + fi->meta.offset_label, &label, sizeof(label));
}

/* What ambulkdelete does to an entry (embedding.c:912-922): set DELETED_FLAG in the label's flags. */
void
flat_mark_deleted(FlatIndex *fi, idx_t idx, int deleted)
{
	label_t l;

	memcpy(&l, flat_record(fi, idx) + fi->meta.offset_label, sizeof(l));
	if (deleted)
		l |= ((label_t) DELETED_FLAG) << 48;
	else
		l &= ~(((label_t) DELETED_FLAG) << 48);
	flat_set_label(fi, idx, l);
}

void
flat_set_ef(FlatIndex *fi, size_t efConstruction, size_t efSearch)
{
	fi->meta.efConstruction = efConstruction;
	fi->meta.efSearch = efSearch;
}

/* One hnsw_search call exactly as hnsw_gettuple issues it (embedding.c:317): labels ascending by
 * distance, malloc'd by the callee and freed here.  labels_out must hold efSearch entries.
 * Returns n (>=0) or -1 on failure. */
long
flat_search(FlatIndex *fi, const coord_t *q, size_t efSearch, label_t *labels_out)
{
	FlatIndex local = *fi; /* per-call meta copy so concurrent callers may differ in efSearch */
	size_t	  n = 0;
	label_t	 *res = NULL;

	local.meta.efSearch = efSearch;
	if (!hnsw_search(&local.meta, q, &n, &res))
		return -1;
	memcpy(labels_out, res, n * sizeof(label_t));
	free(res);
	return (long) n;
}

dist_t
flat_dist(int dist_func, const coord_t *a, const coord_t *b, size_t dim)
{
	pthread_once(&init_once, init_dist);
	return hnsw_dist_func((dist_func_t) dist_func, a, b, dim);
}

void
flat_dist_many(int dist_func, const coord_t *a, const coord_t *b, size_t dim, size_t n, int broadcast_a, dist_t *out)
{
	pthread_once(&init_once, init_dist);
	for (size_t i = 0; i < n; i++)
		out[i] = hnsw_dist_func((dist_func_t) dist_func, broadcast_a ? a : a + i * dim, b + i * dim, dim);
}

void flat_counters_reset(void) { memset(&tl_counters, 0, sizeof(tl_counters)); }
void flat_counters_get(uint64_t out[5])
{
	out[0] = tl_counters.coord_reads;
	out[1] = tl_counters.link_reads;
	out[2] = tl_counters.link_words;
	out[3] = tl_counters.label_reads;
	out[4] = tl_counters.writes;
}

/* ------------------------------------------------------------------------------------------
 * Multi-threaded timed search: nthreads independent readers over a shared read-only graph
 * (the search path keeps all state local to searchBaseLayer, hnswalg.cpp:45-53).
 * ---------------------------------------------------------------------------------------- */

typedef struct
{
	FlatIndex	  *fi;
	const coord_t *queries;
	size_t		   nq, ef;
	size_t		   t, nthreads;
	label_t		  *labels;	/* nq * ef or NULL */
	int32_t		  *n_out;	/* nq or NULL */
	uint64_t	  *counters; /* nq * 3 (coord_reads, link_reads, link_words) or NULL */
	int			   reps;
	int			   failed;
} Worker;

static void *
worker_main(void *arg)
{
	Worker	  *w = (Worker *) arg;
	FlatIndex  local = *w->fi;
	label_t	  *res;
	size_t	   n;

	local.meta.efSearch = w->ef;
	for (int rep = 0; rep < w->reps; rep++)
		for (size_t i = w->t; i < w->nq; i += w->nthreads)
		{
			if (w->counters)
				flat_counters_reset();
			res = NULL;
			n = 0;
			if (!hnsw_search(&local.meta, w->queries + i * local.meta.dim, &n, &res))
			{
				w->failed = 1;
				continue;
			}
			if (w->labels)
			{
				memcpy(w->labels + i * w->ef, res, n * sizeof(label_t));
				for (size_t k = n; k < w->ef; k++)
					w->labels[i * w->ef + k] = ~(label_t) 0;
			}
			if (w->n_out)
				w->n_out[i] = (int32_t) n;
			if (w->counters)
			{
				w->counters[i * 3 + 0] = tl_counters.coord_reads;
				w->counters[i * 3 + 1] = tl_counters.link_reads;
				w->counters[i * 3 + 2] = tl_counters.link_words;
			}
			free(res);
		}
	return NULL;
}

/* Returns elapsed seconds (CLOCK_MONOTONIC) for reps passes over all nq queries, or -1. */
double
flat_search_many(FlatIndex *fi, const coord_t *queries, size_t nq, size_t ef, int nthreads, int reps,
				 label_t *labels, int32_t *n_out, uint64_t *counters)
{
	pthread_t	   *th;
	Worker		   *ws;
	struct timespec t0, t1;
	int				failed = 0;

	if (nthreads < 1)
		nthreads = 1;
	th = (pthread_t *) malloc(sizeof(pthread_t) * nthreads);
	ws = (Worker *) calloc(nthreads, sizeof(Worker));
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int t = 0; t < nthreads; t++)
	{
		ws[t] = (Worker){fi, queries, nq, ef, (size_t) t, (size_t) nthreads, labels, n_out, counters, reps, 0};
		pthread_create(&th[t], NULL, worker_main, &ws[t]);
	}
	for (int t = 0; t < nthreads; t++)
	{
		pthread_join(th[t], NULL);
		failed |= ws[t].failed;
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	free(th);
	free(ws);
	if (failed)
		return -1.0;
	return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

/* Sequential build of n points in id order through flat_add; returns elapsed seconds or -1. */
double
flat_build(FlatIndex *fi, size_t n, const coord_t *coords, const label_t *labels)
{
	struct timespec t0, t1;

	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (size_t i = 0; i < n; i++)
		if (flat_add(fi, coords + i * fi->meta.dim, labels ? labels[i] : (label_t) fi->n_items) != 0)
			return -1.0;
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}
