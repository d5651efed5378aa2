// pgemb::scan_tile_kernel on the host SIMT emulator (tests/emu).
#include <cuda_runtime.h>  // the stand-in in tests/emu/fake_cuda

#include "../../pg_embedding_b200/csrc/scan_tile_kernel.cuh"

using namespace pgemb;

template <int METRIC>
static void run(const float *vectors, const float *norms, uint32_t row_f, uint32_t dim, const float *queries, uint32_t q_stride,
				const float *qnorms, uint32_t nq, uint32_t r0, uint32_t nr, float *out)
{
	const dim3 grid((nq + ScanTile<METRIC>::TQ - 1) / ScanTile<METRIC>::TQ, (nr + kScanTileRows - 1) / kScanTileRows);
	emu::launch(grid, kScanThreads, 0, [=]() { scan_tile_kernel<METRIC>(vectors, norms, row_f, dim, queries, q_stride, qnorms, nq, r0, nr, out); });
}

extern "C" void emu_scan_tile(int metric, const float *vectors, const float *norms, uint32_t row_f, uint32_t dim, const float *queries,
							  uint32_t q_stride, const float *qnorms, uint32_t nq, uint32_t r0, uint32_t nr, float *out)
{
	if (metric == 0) run<M_L2>(vectors, norms, row_f, dim, queries, q_stride, qnorms, nq, r0, nr, out);
	else if (metric == 1) run<M_COS>(vectors, norms, row_f, dim, queries, q_stride, qnorms, nq, r0, nr, out);
	else run<M_MAN>(vectors, norms, row_f, dim, queries, q_stride, qnorms, nq, r0, nr, out);
}
