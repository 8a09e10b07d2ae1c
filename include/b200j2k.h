/*
 * b200j2k.h -- C ABI of the B200-native HTJ2K decode path.
 *
 * This is the boundary a libavcodec maintainer binds to replace the hot path of the native "jpeg2000"
 * decoder of sandflow/ffmpeg-ht (FFCodec ff_jpeg2000_decoder, libavcodec/jpeg2000dec.c:2926-2939).
 * The FFCodec shim that calls it is ffmpeg-ht_b200/ffcodec/jpeg2000_b200.c; INTEGRATION.md shows the
 * build glue.  Plain C: pointers and sizes only, no C++ or torch types, no exceptions, no global state.
 *
 * Division of labour (BASELINE.json north_star):
 *   host, serial  : codestream/marker parsing, geometry, tag trees, packet headers
 *                   (jpeg2000dec.c:197-1869, jpeg2000.c:214-577)  -> flat code-block descriptor table +
 *                   packed segment buffer
 *   sm_100a CUDA  : HT block decode + dequantisation (jpeg2000htdec.c:1188-1336, jpeg2000dec.c:2098-2181),
 *                   inverse DWT (jpeg2000dwt.c:309-620), inverse RCT/ICT (jpeg2000dsp.c:43-91),
 *                   level shift / clip / pack (jpeg2000dec.c:2301-2366)
 * There is no CPU fallback: every entry point that needs the device fails with B200J2K_ERR_DEVICE when
 * no CUDA device is usable.
 *
 * Return values: >= 0 success, < 0 one of the B200J2K_ERR_* codes, which are numerically the AVERROR
 * values the reference returns in the same situation (jpeg2000dec.c:2905-2907).
 */
#ifndef B200J2K_H
#define B200J2K_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200J2K_ERR_INVALIDDATA  (-1094995529)   /* AVERROR_INVALIDDATA */
#define B200J2K_ERR_PATCHWELCOME (-1163346256)   /* AVERROR_PATCHWELCOME */
#define B200J2K_ERR_NOMEM        (-12)           /* AVERROR(ENOMEM) */
#define B200J2K_ERR_EINVAL       (-22)           /* AVERROR(EINVAL) */
#define B200J2K_ERR_UNSPECIFIED  (-1)            /* the reference's bare -1: marker segment length mismatch
                                                  * (jpeg2000dec.c:2633), unknown wavelet type (jpeg2000dwt.c:590) */
#define B200J2K_ERR_DEVICE       (-19)           /* AVERROR(ENODEV): no usable CUDA device / CUDA failure */

/* b200j2k_create flags */
#define B200J2K_FLAG_BITEXACT 1   /* AV_CODEC_FLAG_BITEXACT: 9/7 runs the fixed-point path (jpeg2000dec.c:543-544) */
/* AVCodecContext.lowres / the decoder's "lowres" option (jpeg2000dec.c:2807-2823, :2913-2924): drop the n
 * highest resolution levels; picture dimensions shrink by 2^n.  n = 0..15 in bits 8-11 of the flags. */
#define B200J2K_FLAG_LOWRES(n) (((n) & 15) << 8)
#define B200J2K_FLAG_GET_LOWRES(flags) (((flags) >> 8) & 15)
/* Tile subset: this decoder context reconstructs only the tiles whose index t satisfies t % step == first, writes only
 * their rectangles of the output planes and uploads only their tile-parts.  It is how the tiles of an oversized frame
 * are spread over several GPUs (one context per GPU, `first` = its rank, `step` = the number of GPUs, all writing
 * into the same AVFrame): the reference's per-tile jobs (jpeg2000_decode_tile through avctx->execute2,
 * jpeg2000dec.c:2368-2395, :2894) partitioned by device instead of by slice thread.  step <= 1: every tile. */
#define B200J2K_FLAG_TILES(first, step) ((((step) & 255) << 16) | (((first) & 255) << 24))
#define B200J2K_FLAG_GET_TILE_STEP(flags)  (((flags) >> 16) & 255)
#define B200J2K_FLAG_GET_TILE_FIRST(flags) (((unsigned)(flags) >> 24) & 255)

/* Output memory layouts.  b200j2k_frame_info.av_pix_fmt_name gives the exact libavutil format name; the FFCodec
 * shim maps it to AVPixelFormat with av_get_pix_fmt().  Selection follows jpeg2000dec.c:134-193,:328-412,
 * including the JP2 colour specification (sRGB / greyscale / sYCC lists) and channel definitions. */
enum b200j2k_pix_fmt {
    B200J2K_PIX_NONE = 0,
    B200J2K_PIX_GRAY8, B200J2K_PIX_GRAY16, B200J2K_PIX_YA8, B200J2K_PIX_YA16,
    B200J2K_PIX_RGB24, B200J2K_PIX_RGB48, B200J2K_PIX_RGBA, B200J2K_PIX_RGBA64,
    B200J2K_PIX_PLANAR8, B200J2K_PIX_PLANAR16
};
/* Palettised JP2 files (one component of at most 8 bits + a 'pclr' box, jpeg2000dec.c:2711-2752): av_pix_fmt_name is
 * "pal8", pix_fmt is B200J2K_PIX_GRAY8 (the index plane) and nplanes is 2 -- plane 1 is AVFrame.data[1] of a PAL8
 * frame: one row of 256 native-endian 0xAARRGGBB words (plane_row_bytes[1] = 1024; its linesize is not looked at). */

/* What jpeg2000_decode_frame() publishes on AVCodecContext / AVFrame before the pixels exist
 * (jpeg2000dec.c:213, :326, :354-363, :420, :546). */
typedef struct b200j2k_frame_info {
    int32_t width, height;          /* ff_set_dimensions() arguments */
    int32_t pix_fmt;                /* enum b200j2k_pix_fmt */
    int32_t ncomponents;
    int32_t precision;              /* bits_per_raw_sample */
    int32_t lossless;               /* FF_CODEC_PROPERTY_LOSSLESS */
    int32_t profile;                /* Rsiz */
    int32_t nplanes;
    int32_t plane_width[4];         /* samples per row of each output plane */
    int32_t plane_height[4];
    int32_t plane_row_bytes[4];     /* minimum linesize */
    int32_t ntiles;
    int32_t ncodeblocks;            /* code-blocks handed to the device */
    int64_t bytes_consumed;         /* decode() return value of the reference: bytestream2_tell() */
    char    av_pix_fmt_name[16];    /* libavutil name of the AVPixelFormat the reference would choose for this
                                     * stream ("rgb48le", "yuv422p10le", ...; jpeg2000dec.c:328-412): pix_fmt
                                     * above is only its memory-layout class */
    int32_t sar_num, sar_den;       /* AVCodecContext.sample_aspect_ratio from the capture / display resolution box
                                     * of a JP2 file ('res ', jpeg2000dec.c:2762-2796, :2867-2868); 0 / 0 = leave it */
    int32_t nblocks_failed;         /* code-blocks zero-filled because their segment was damaged: what the reference
                                     * logs block by block and carries on from (jpeg2000dec.c:2275-2278).  Known once
                                     * the frame is done: filled by b200j2k_decode(), else b200j2k_slot_blocks_failed() */
    int32_t reserved;
} b200j2k_frame_info;

typedef struct b200j2k_decoder b200j2k_decoder;

/* Number of CUDA devices usable by this library (0 when none: nothing else will work). */
int b200j2k_device_count(void);

/* jpeg2000_decode_init (jpeg2000dec.c:2807-2823) + device context: one decoder per AVCodecContext.
 * `slots` frames may be in flight at once (>= 1); each slot owns a CUDA stream, pinned staging and
 * device buffers that grow on demand. */
int b200j2k_create(b200j2k_decoder **out, int device, int slots, int flags);

/* FFCodec.close: releases streams, pinned and device memory. */
void b200j2k_destroy(b200j2k_decoder *dec);

/* Header-only pass (main header + tile-part headers, no packet parsing): what the caller needs to
 * allocate the AVFrame (ff_thread_get_buffer, jpeg2000dec.c:2877).  Works without a device. */
int b200j2k_probe(const uint8_t *pkt, int64_t pkt_size, int flags, b200j2k_frame_info *info);

/* Asynchronous decode of one packet (= one codestream, AVPacket.data/size) into caller-owned planes
 * (AVFrame.data[]/linesize[]).  Host parsing happens in the call; H2D, kernels and D2H are queued on the
 * slot's stream.  Returns the slot index (>= 0) to pass to b200j2k_wait().  If every slot is busy the
 * oldest one is waited for first.  A packet in ordinary (pageable) memory is copied into pinned staging and
 * may be released when the call returns; a packet in pinned memory (b200j2k_host_alloc / _register) is
 * DMA'd from where it lies and must stay valid until b200j2k_wait().  The planes must stay valid until
 * b200j2k_wait().  Planes that were registered with b200j2k_host_register() (or allocated with
 * b200j2k_host_alloc()) receive the device->host copy directly; others go through pinned staging + memcpy. */
int b200j2k_submit(b200j2k_decoder *dec, const uint8_t *pkt, int64_t pkt_size,
                   uint8_t *const data[4], const int linesize[4], b200j2k_frame_info *info);

/* Completes the frame in `slot`: after it returns 0 the planes hold the picture. */
int b200j2k_wait(b200j2k_decoder *dec, int slot);

/* FF_CODEC_DECODE_CB equivalent: submit + wait.  Returns bytes consumed (>= 0) like the reference. */
int64_t b200j2k_decode(b200j2k_decoder *dec, const uint8_t *pkt, int64_t pkt_size,
                       uint8_t *const data[4], const int linesize[4], b200j2k_frame_info *info);

/* Device-resident variants used by tests and by the kernel-only leg of bench.py: the packet has been
 * parsed and staged once (b200j2k_stage), the kernels are re-run on the resident inputs
 * (b200j2k_run_staged) with no host<->device traffic, and the output stays on the device. */
int b200j2k_stage(b200j2k_decoder *dec, int slot, const uint8_t *pkt, int64_t pkt_size, b200j2k_frame_info *info);
int b200j2k_run_staged(b200j2k_decoder *dec, int slot);
int b200j2k_sync(b200j2k_decoder *dec, int slot);
/* like run_staged, but brackets the stages with CUDA events on the slot's stream and waits.  ms[6]:
 * [0] HT un-stuff (VLC + MEL bit arrays), [1] HT MEL/CxtVLC/U-VLC (lane per block), [2] HT MagSgn + dequantisation
 * (+ blocks wider than 64 samples on the general kernel), [3] HT blocks with refinement passes,
 * [4] intermediate IDWT levels, [5] fused last level + MCT + pack */
int b200j2k_run_staged_timed(b200j2k_decoder *dec, int slot, float *ms);
/* copies the staged slot's picture to host planes (synchronous) */
int b200j2k_fetch(b200j2k_decoder *dec, int slot, uint8_t *const data[4], const int linesize[4]);
/* CUDA stream of a slot (cudaStream_t as void*), for event timing on the launching stream */
void *b200j2k_slot_stream(b200j2k_decoder *dec, int slot);
/* device pointer + byte size of the slot's packed output picture, or NULL */
void *b200j2k_slot_output(b200j2k_decoder *dec, int slot, int64_t *bytes);
/* bytes the last b200j2k_submit/_stage of `slot` copied host->device (tables + codestream) */
int64_t b200j2k_slot_h2d_bytes(b200j2k_decoder *dec, int slot);
/* number of kernel launches the last run of `slot` queued */
int b200j2k_slot_launches(b200j2k_decoder *dec, int slot);
/* Debug builds / B200J2K_REDZONE=1 only: dirties the guard zone behind the slot's picture buffer so that the checker
 * can be tested (tests/test_gpu_debug_bounds.py); EINVAL when guard zones are off. */
int b200j2k_debug_poke(b200j2k_decoder *dec, int slot);
/* After b200j2k_wait(): how many code-blocks of the slot's frame were zero-filled because their segment was damaged
 * (the reference logs each one and carries on, jpeg2000dec.c:2275-2278); 0 while the slot is still busy. */
int b200j2k_slot_blocks_failed(b200j2k_decoder *dec, int slot);

/* Stage-level entry points (parity tests against the oracle; same argument meaning as the reference's
 * ff_jpeg2000_decode_htj2k / ff_dwt_decode): all buffers are HOST pointers, the work runs on the device.
 * blocks: n code-blocks described by par[n][10] = {npasses, Lcup, Lref, zbp, w, h, M_b, roi_shift, vsc,
 * data_offset}; data = concatenated Dcup||Dref segments; out = concatenated w*h int32 sign-magnitude
 * coefficients; status[n] = 1 decoded / 0 nothing / <0 error (block zeroed, quirk Q1). */
int b200j2k_ht_blocks(int device, const uint8_t *data, int64_t data_size, const int32_t *par, int n,
                      int32_t *out, int32_t *status);
/* type: 0 = 9/7 float, 1 = 5/3, 2 = 9/7 fixed point; border = {x0,x1,y0,y1}; plane in Mallat layout */
int b200j2k_idwt(int device, int type, const int32_t *border, int levels, void *plane);

/* Device-resident output ("zero-copy output / hw frames"): the planes handed to
 * b200j2k_submit / _decode may be DEVICE pointers (an AV_PIX_FMT_CUDA frame's data[] / linesize[]); the
 * picture is then copied device-to-device on the slot's stream and never crosses PCIe.  Helpers for callers
 * without their own CUDA allocator: */
void *b200j2k_device_alloc(int device, int64_t bytes);
void b200j2k_device_free(void *p);
int b200j2k_device_download(void *dst_host, const void *src_device, int64_t bytes);

/* After b200j2k_submit(): returns once the packet has been read (its host->device copies are done), i.e. when a
 * packet in page-locked memory may be released although the frame is still being decoded.  A pageable packet has been
 * copied by the time submit returns.  The FFCodec shim with async_depth > 1 needs it: libavcodec keeps an AVPacket
 * valid only for the duration of the decode callback. */
int b200j2k_wait_upload(b200j2k_decoder *dec, int slot);
/* 1 when the pointer lies in page-locked (cudaHostAlloc'ed or registered) host memory */
int b200j2k_pointer_is_pinned(const void *p);

/* Device ordinal of the GPU a pointer's memory lives on, -1 for host memory: lets the FFCodec shim put its decoder
 * context next to the AV_PIX_FMT_CUDA frames the caller hands it (AVHWFramesContext, libavutil/hwcontext_cuda.c:204). */
int b200j2k_pointer_device(const void *p);

/* pinned host memory helpers for AVFrame pools (get_buffer2) */
void *b200j2k_host_alloc(int64_t bytes);
void b200j2k_host_free(void *p);
int b200j2k_host_register(void *p, int64_t bytes);
int b200j2k_host_unregister(void *p);

/* Introspection of the host-side plan (CPU only, no device needed): lets tests check the parser's flat
 * tables.  which: 0 CblkDesc, 1 BandDesc, 2 TileCompDesc, 3 PackGroup (ffmpeg-ht_b200/csrc/j2k_plan.h),
 * 4 gathered bytes of multi-contribution blocks, 5 the palette of a PAL8 frame (256 words).  Pointers stay valid until the next plan_build on the
 * calling thread. */
int b200j2k_plan_build(const uint8_t *pkt, int64_t pkt_size, int flags, b200j2k_frame_info *info);
const void *b200j2k_plan_table(int which, int64_t *count, int32_t *elem_size);
int64_t b200j2k_plan_scalar(int which);

/* ---- frame splitter: the AVCodecParser "jpeg2000" (libavcodec/jpeg2000_parser.c:36-215) ------------------
 * Cuts a byte stream of concatenated codestreams / JP2 files, fed in arbitrary pieces, into the packets
 * b200j2k_submit takes.  Host-only (a length-skipping serial scan).
 *   create(complete_frames)   complete_frames != 0 = PARSER_FLAG_COMPLETE_FRAMES: every input is one frame
 *   find_frame_end            jpeg2000_parser.c:91-175 alone, for a binding that keeps libavcodec's own
 *                             ff_combine_frame: offset of the next frame's first byte in buf (may be < 0:
 *                             the frame ended before this piece), or B200J2K_SPLIT_END_NOT_FOUND
 *   parse                     jpeg2000_parse (:177-208) incl. ff_combine_frame (parser.c:203-288): returns
 *                             the number of input bytes consumed (may be < 0, callers clamp to 0 as
 *                             av_parser_parse2 does); *out / *out_size = one complete frame or NULL / 0.
 *                             *out points into buf or into the splitter's own storage (valid until the next
 *                             call).  A call with size == 0 flushes the last frame at end of input. */
#define B200J2K_SPLIT_END_NOT_FOUND (-100)
typedef struct b200j2k_splitter b200j2k_splitter;
int b200j2k_splitter_create(b200j2k_splitter **out, int complete_frames);
void b200j2k_splitter_destroy(b200j2k_splitter *sp);
int b200j2k_splitter_find_frame_end(b200j2k_splitter *sp, const uint8_t *buf, int size);
int b200j2k_splitter_parse(b200j2k_splitter *sp, const uint8_t **out, int *out_size, const uint8_t *buf, int size);

const char *b200j2k_pix_fmt_name(int pix_fmt);
const char *b200j2k_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200J2K_H */
