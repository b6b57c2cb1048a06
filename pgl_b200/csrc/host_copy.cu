// host_copy.cu -- strided copy between pinned HOST memory and device memory done by a kernel over the
// unified address space (zero-copy loads / stores across PCIe) instead of the copy engine.
// EXPERIMENTAL (PGLB_HOST_COPY=kernel), written after round 1's GPU budget was spent, not yet run.
//
// Why: the host-buffer entry of the aggregation (Graph.send_recv_host) streams
// column blocks of a row-major [N, D] matrix.  cudaMemcpy2DAsync moves 256-byte-wide blocks at full
// PCIe speed but 128-byte-wide ones at about half (measured: 2 blocks 168 ms, 4 blocks 192 ms per step),
// which pins the pipeline at two blocks and leaves a 45 ms un-overlapped head and tail.  A kernel that
// reads / writes the pinned buffer directly issues one 128-byte PCIe request per warp-row whatever the
// block width, so finer blocks should not lose bandwidth.  A handful of CTAs is enough to saturate the
// link (the aggregation kernel keeps the rest of the machine).
#include "common.cuh"

namespace pglb {

// one warp moves `width_bytes` (multiple of 16) of one row per iteration, 16 bytes per lane; rows are
// distributed over all warps of the grid; 4 rows in flight per warp
__global__ void __launch_bounds__(256) copy2d_kernel(char *__restrict__ dst, size_t dpitch,
                                                     const char *__restrict__ src, size_t spitch,
                                                     int width_bytes, int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int chunks = width_bytes >> 4;
    for (int64_t r0 = warp * 4; r0 < rows; r0 += nwarps * 4) {
        for (int c = lane; c < chunks; c += 32) {
            uint4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (r0 + i < rows) v[i] = *reinterpret_cast<const uint4 *>(src + (r0 + i) * spitch + (size_t)c * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (r0 + i < rows) *reinterpret_cast<uint4 *>(dst + (r0 + i) * dpitch + (size_t)c * 16) = v[i];
        }
    }
}

}  // namespace pglb

extern "C" int pglb_copy2d_kernel_async(void *dst, size_t dpitch, const void *src, size_t spitch,
                                        size_t width_bytes, int64_t height, int ctas, void *stream) {
    using namespace pglb;
    PGLB_CHECK_ARG(height >= 0, PGLB_EINVAL, "pglb_copy2d_kernel_async: negative height");
    if (width_bytes == 0 || height == 0) return PGLB_OK;
    PGLB_CHECK_ARG(dst && src, PGLB_EINVAL, "pglb_copy2d_kernel_async: NULL pointer");
    PGLB_CHECK_ARG(width_bytes % 16 == 0 && dpitch % 16 == 0 && spitch % 16 == 0 &&
                       (reinterpret_cast<uintptr_t>(dst) % 16) == 0 && (reinterpret_cast<uintptr_t>(src) % 16) == 0,
                   PGLB_ESHAPE, "pglb_copy2d_kernel_async: width, pitches and pointers must be multiples of 16 bytes");
    PGLB_CHECK_ARG(width_bytes <= dpitch && width_bytes <= spitch && width_bytes < (1u << 30), PGLB_ESHAPE,
                   "pglb_copy2d_kernel_async: width larger than a pitch");
    if (ctas <= 0) ctas = 16;
    copy2d_kernel<<<(unsigned)ctas, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<char *>(dst), dpitch, reinterpret_cast<const char *>(src), spitch, (int)width_bytes, height);
    PGLB_LAUNCH_CHECK("copy2d_kernel");
    return PGLB_OK;
}
