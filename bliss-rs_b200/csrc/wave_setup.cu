// wave_setup.cu -- first kernel of every wave: pulls the wave's descriptors (SongDesc array +
// prefix arrays) from pinned host memory into the wave set's device blob with plain loads over
// PCIe, and zeroes the two per-song counters the later kernels accumulate into.
//
// Why a kernel and not cudaMemcpyAsync/cudaMemsetAsync: on the host path the copy engine is busy
// with back-to-back 64..512 MB PCM copies, and a small H2D copy enqueued on another stream is only
// served when that queue runs dry -- measured (BLISS_B200_TRACE_CHUNKS): the kernels of chunks 0..2
// started after chunk 3's copy, 12 ms late, and every later group of chunks likewise.  SM loads from
// mapped host memory do not queue behind the DMA engine.
#include "common.cuh"

namespace bliss {

__global__ void wave_setup_kernel(const uint4 *__restrict__ src_host, uint4 *__restrict__ dst, unsigned int n16,
                                  unsigned int *__restrict__ zcr_count, unsigned int *__restrict__ cand_count,
                                  unsigned int n_songs) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n16) dst[i] = src_host[i];
    if (i < n_songs) {
        zcr_count[i] = 0u;
        cand_count[i] = 0u;
    }
}

int launch_wave_setup(const void *src_host, void *dst, size_t bytes, unsigned int *zcr_count,
                      unsigned int *cand_count, unsigned int n_songs, cudaStream_t st) {
    const unsigned int n16 = (unsigned int)((bytes + 15) / 16);
    const unsigned int n = n16 > n_songs ? n16 : n_songs;
    if (n == 0) return 0;
    wave_setup_kernel<<<(n + 255u) / 256u, 256, 0, st>>>(reinterpret_cast<const uint4 *>(src_host),
                                                         reinterpret_cast<uint4 *>(dst), n16, zcr_count, cand_count,
                                                         n_songs);
    return 1;
}

}  // namespace bliss
