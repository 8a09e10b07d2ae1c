/* j2k_kernels.h -- launchers of the sm_100a kernels (ht_block.cu, idwt.cu), host-callable. */
#ifndef B200J2K_KERNELS_H
#define B200J2K_KERNELS_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "j2k_plan.h"

#define B200J2K_KERR_INVALID (-1094995529)

/* status[] convention of the block kernels: one word per code-block (1 decoded, 0 nothing to decode,
 * B200J2K_KERR_INVALID zero-filled because its segment was damaged) and, IN FRONT of the array, status[-1] = how
 * many blocks of the frame ended up INVALID -- what the reference reports block by block with av_log
 * (jpeg2000dec.c:2275-2278).  The host zeroes the counter per frame and reads it back with the picture. */
#ifdef __CUDACC__
static __device__ __forceinline__ void j2k_block_status(int32_t *status, uint32_t bi, int32_t value)
{
    if (!status)
        return;
    status[bi] = value;
    if (value < 0)
        atomicAdd(status - 1, 1);
}
#endif

namespace b200j2k {

/* HT block decode + dequantisation.  order: optional list of descriptor indices (nullptr = 0..n-1).
 * refine = blocks carrying SigProp/MagRef passes (needs the shared-memory sample copy).
 * maxw: widest block; max_area: largest w*h; max_state: largest (w+3)*(h+3) (both only matter for refine).
 * Returns 1 when a kernel was queued, 0 when there was nothing to do, < 0 on launch failure. */
int launch_ht_decode(cudaStream_t stream, bool refine, const CblkDesc *descs, const uint32_t *order, int nblocks,
                     const uint8_t *bytes, const BandDesc *bands, void *coef, int32_t *status,
                     int max_Lcup, int maxw, int max_area, int max_state);
size_t ht_smem_bytes(bool refine, int stream_words, int maxQW, int samp_words, int st_words);

/* Split cleanup pipeline (ht_cleanup.cu) for cleanup-only blocks up to 64 samples wide.
 * Words of one un-stuffed VLC or MEL bit array of a block whose suffix is Scup bytes long. */
static inline __host__ __device__ int ht_stream_words(int Scup) { return (Scup + 3) / 4 + 3; }
/* Host: words of HBM scratch a block needs (un-stuffed bit arrays sized for the largest possible suffix,
 * one word per quad and -- with refinement passes -- the Dref bit arrays and three significance maps);
 * always even, 0 for blocks without data. */
uint32_t ht_cleanup_scratch_words(const CblkDesc &d);
/* bounds assertions that failed in the kernels of ht_cleanup.cu since the last call (always 0 unless the file was
 * built with -DB200J2K_DEBUG_BOUNDS); synchronises the device */
int ht_cleanup_bounds_failures();
/* scr_off: device array of scratch word offsets, indexed like descs.  order[] = [refine | coded | zero]:
 *   refine  blocks with SigProp / MagRef passes (<= 64 samples wide): un-stuffed and VLC-decoded here,
 *           finished by launch_ht_refine
 *   coded   cleanup-only blocks with data (headers are validated on the device)
 *   zero    blocks that only need zeros
 * Returns the number of kernels queued or < 0. */
int launch_ht_cleanup(cudaStream_t stream, const CblkDesc *descs, const uint32_t *order, const uint32_t *scr_off,
                      int n_refine, int n_coded, int n_zero, const uint8_t *bytes, const BandDesc *bands, void *scratch,
                      void *coef, int32_t *status, int max_Lcup, int max_Lref, cudaEvent_t after_unstuff = nullptr,
                      cudaEvent_t after_vlc = nullptr);
/* MagSgn (raw samples + significance map), SigProp (lane per block), MagRef + apply + dequantise for
 * order[0..n_refine); status must be a device array (the later kernels skip blocks the earlier ones rejected) */
int launch_ht_refine(cudaStream_t stream, const CblkDesc *descs, const uint32_t *order, const uint32_t *scr_off,
                     int n_refine, const uint8_t *bytes, const BandDesc *bands, void *scratch, void *coef,
                     int32_t *status, int max_Lcup);

/* Part-1 (EBCOT / MQ) code-blocks of style 0 (mq_block.cu): order[0..nblocks) lists descriptors flagged
 * J2K_CBLK_MQ that carry data; one warp per block, dequantisation fused into the write-back.  state_words: the
 * largest mq_state_words_of() over those blocks. */
int launch_mq_blocks(cudaStream_t stream, const CblkDesc *descs, const uint32_t *order, int nblocks, const uint8_t *bytes,
                     const BandDesc *bands, void *coef, int32_t *status, int state_words);
/* words of shared-memory state a w x h block needs there: two arrays of one word per 4-row stripe and column, framed */
static inline __host__ __device__ int mq_state_words_of(int w, int h) { return 2 * (((h + 7) >> 2) + 1) * (w + 2); }

/* Inverse DWT levels + final MCT / level shift / clip / pack.
 * tcs/groups: device copies of the plan tables.  coefA: HT output (Mallat layout), coefB/coefC: LL
 * ping-pong, picture: packed output.  Returns the number of kernels queued or < 0. */
struct IdwtStreamCache;           /* idwt_stream.cu: memoised launch parameters (tensor maps) of one geometry */
IdwtStreamCache *idwt_stream_cache_create();
void idwt_stream_cache_destroy(IdwtStreamCache *c);

struct IdwtFrame {
    const TileCompDesc *d_tcs;    /* device */
    const PackGroup *d_groups;    /* device */
    const TileCompDesc *h_tcs;    /* host copies, for grid sizing */
    const PackGroup *h_groups;
    int ntcs, ngroups;
    void *coefA, *coefB, *coefC;
    uint8_t *picture;
    uint64_t plane_off[4];
    int plane_row_bytes[4];
    cudaEvent_t before_final;     /* optional: recorded between the level kernels and the fused last level */
    IdwtStreamCache *stream_cache;/* optional: owned by the slot; valid while geom_serial and the buffers stay the same */
    uint64_t geom_serial;         /* Parser::geom_serial of the tables above */
};
void idwt_stream_cache_begin(IdwtStreamCache *c, uint64_t serial, const IdwtFrame &f);
int launch_idwt_frame(cudaStream_t stream, const IdwtFrame &f);

/* Register-streaming fast path (idwt_stream.cu).  The *_ok functions decide on the host, per launch,
 * whether every tile-component / pack group of transform `type` fits it (even origins, aligned rows,
 * packed 16-bit output); launch_idwt_frame falls back to the shared-memory tiled kernels otherwise. */
bool idwt_stream_levels_ok(const TileCompDesc *h_tcs, int ntcs, int type);
bool idwt_stream_final_ok(const IdwtFrame &f, int type, int *ncomp);
int launch_idwt_stream_level(cudaStream_t stream, int type, const IdwtFrame &f, int j, int max_w, int max_h,
                             const void *in, void *out);
int launch_idwt_stream_final(cudaStream_t stream, int type, int ncomp, const IdwtFrame &f, int max_w, int max_h);

/* one plane, all levels, result left in `plane` (stage-level entry point for parity tests) */
int launch_idwt_plane(cudaStream_t stream, int type, const int32_t *border, int levels, void *plane,
                      void *scratchB, void *scratchC, TileCompDesc *d_tc_scratch);

}  // namespace b200j2k
#endif
