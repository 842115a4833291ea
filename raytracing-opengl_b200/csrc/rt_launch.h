/* rt_launch.h — launch geometry shared by the kernels and the host API. */
#ifndef RT_LAUNCH_H
#define RT_LAUNCH_H

#include <cuda_runtime.h>
#include "rt_params.h"

#define QUAD_THREADS 128
/* persistent kernel: ONE 20-warp CTA per SM (all warps share one staged scene and, while the frame drains, one job pool).
 * 640 threads cap ptxas at 96 registers (a 250-byte spill outside the hot loops); measured against 512 threads / 124
 * registers: mixed1024@4K 441 -> 437 ms, spheres4k 35.2 -> 33.5 ms; 768 threads / 80 registers is slower (445 ms). */
#ifndef PERSIST_THREADS
#define PERSIST_THREADS 640
#endif
/* scenes without tori: 24 warps / 80 registers (fused build: spheres4k 17.9 -> 16.7 ms, profiles/r2_fused_thread_variants_ab.jsonl) */
#define PERSIST_THREADS_WIDE 768
#ifndef PERSIST_MIN_BLOCKS
#define PERSIST_MIN_BLOCKS 1
#endif
/* dynamic shared memory of the persistent kernel: staged scene (rounded up to 128 B) + one 32-B job and one 32-B result per thread */
#define PERSIST_SMEM_BYTES_T(scene_bytes, threads) ((((size_t)(scene_bytes) + 127u) & ~(size_t)127u) + (size_t)(threads) * 64u)
#define PERSIST_SMEM_BYTES(scene_bytes) PERSIST_SMEM_BYTES_T(scene_bytes, PERSIST_THREADS)
#define RTB_LAUNCH_QUAD 1
#define RTB_LAUNCH_PERSISTENT 2
#define RTB_LAUNCH_PERSISTENT_WIDE 3

#ifdef __cplusplus
extern "C" {
#endif
/* strict build (-fmad=false) */
int rtb_strict_launch_pack(const FrameParams* P, uint8_t* dst, cudaStream_t st);
int rtb_strict_launch(const FrameParams* P, int kernel, int counted, int grid, int threads, size_t smem, cudaStream_t st);
int rtb_strict_occupancy(int kernel, size_t smem, int* blocks_per_sm);
/* fast build */
int rtb_fast_launch_pack(const FrameParams* P, uint8_t* dst, cudaStream_t st);
int rtb_fast_launch(const FrameParams* P, int kernel, int counted, int grid, int threads, size_t smem, cudaStream_t st);
int rtb_fast_occupancy(int kernel, size_t smem, int* blocks_per_sm);
#ifdef __cplusplus
}
#endif
#endif
