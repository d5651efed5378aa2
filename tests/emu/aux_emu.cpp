// The small kernels around the traversal (aux_kernels.cuh) on the host SIMT emulator (tests/emu):
// distance pairs (= hnsw_dist_func), cached norms, exact scan (distance step + running top-k), shard merge.
#include <cuda_runtime.h>  // the stand-in in tests/emu/fake_cuda

#include "../../pg_embedding_b200/csrc/aux_kernels.cuh"

using namespace pgemb;

extern "C" void emu_dist_pairs(int metric, const float *a, const float *b, uint32_t dim, uint32_t n, int broadcast_a, float *out)
{
	const uint32_t threads = 128, lanes = (metric == 0) ? 8 : 4;
	const uint32_t blocks = (n * lanes + threads - 1) / threads;
	if (n == 0) return;
	if (metric == 0) emu::launch(dim3(blocks), threads, 0, [=]() { dist_pairs_kernel<M_L2>(a, b, dim, dim, dim, n, broadcast_a, out); });
	else if (metric == 1) emu::launch(dim3(blocks), threads, 0, [=]() { dist_pairs_kernel<M_COS>(a, b, dim, dim, dim, n, broadcast_a, out); });
	else emu::launch(dim3(blocks), threads, 0, [=]() { dist_pairs_kernel<M_MAN>(a, b, dim, dim, dim, n, broadcast_a, out); });
}

extern "C" void emu_norms(const float *vectors, uint32_t row_f, uint32_t dim, uint32_t first, uint32_t n, float *norms)
{
	if (n == 0) return;
	emu::launch(dim3((n * 4 + 127) / 128), 128, 0, [=]() { norms_kernel(vectors, row_f, dim, first, n, norms); });
}

// exact k-NN over rows [0,N) in chunks, as pgemb_scan_topk drives the two kernels
extern "C" void emu_scan_topk(int metric, const float *vectors, const float *norms, const uint64_t *labels, uint32_t row_f, uint32_t dim,
							  uint32_t N, const float *queries, uint32_t nq, uint32_t k, uint32_t chunk, uint32_t *top_d, uint64_t *top_l,
							  uint32_t *top_n)
{
	std::vector<float>	  dist((size_t) nq * chunk);
	std::vector<uint32_t> tmp_d((size_t) nq * k);
	std::vector<uint64_t> tmp_l((size_t) nq * k);
	float	 *dp = dist.data();
	uint32_t *sd = tmp_d.data();
	uint64_t *sl = tmp_l.data();
	for (uint32_t q = 0; q < nq; q++) top_n[q] = 0;
	const uint32_t lanes = (metric == 0) ? 8 : 4;
	for (uint32_t r0 = 0; r0 < N; r0 += chunk)
	{
		const uint32_t nr = (N - r0 < chunk) ? (N - r0) : chunk;
		const uint32_t blocks = (uint32_t) (((size_t) nq * nr * lanes + 127) / 128);
		if (metric == 0) emu::launch(dim3(blocks), 128, 0, [=]() { scan_dist_kernel<M_L2>(vectors, norms, row_f, dim, queries, dim, nq, r0, nr, dp); });
		else if (metric == 1) emu::launch(dim3(blocks), 128, 0, [=]() { scan_dist_kernel<M_COS>(vectors, norms, row_f, dim, queries, dim, nq, r0, nr, dp); });
		else emu::launch(dim3(blocks), 128, 0, [=]() { scan_dist_kernel<M_MAN>(vectors, norms, row_f, dim, queries, dim, nq, r0, nr, dp); });
		emu::launch(dim3((nq + 3) / 4), 128, 0, [=]() { scan_select_kernel(dp, labels, nq, r0, nr, k, top_d, top_l, top_n, sd, sl); });
	}
}

extern "C" void emu_merge_topk(uint32_t nq, uint32_t n_shards, uint32_t k, const float *din, const uint64_t *lin, const int32_t *nin,
							   float *dout, uint64_t *lout, int32_t *nout)
{
	if (nq == 0) return;
	ShardLists in;
	memset(&in, 0, sizeof(in));
	for (uint32_t s = 0; s < n_shards; s++)	 // [shard][query][k], as pgemb_merge_topk_device lays them out
	{
		in.dist[s] = din + (size_t) s * nq * k;
		in.lab[s] = lin + (size_t) s * nq * k;
		in.cnt[s] = nin + (size_t) s * nq;
	}
	emu::launch(dim3((nq * 32 + 127) / 128), 128, 0, [=]() { merge_topk_lists_kernel(nq, n_shards, k, in, nullptr, 0u, 0u, dout, lout, nout, nullptr); });
}
