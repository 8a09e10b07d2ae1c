/*
 * oracle/hwcuda_standin.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Stand-in for libavutil/hwcontext_cuda.c of the reference (which needs the external nv-codec-headers and is
 * therefore not compiled here): the HWContextType "cuda" that libavutil/hwcontext.c -- the UNMODIFIED generic
 * layer, compiled from /root/reference by oracle/Makefile -- dispatches to.  It does what the real one does for a
 * frame pool (hwcontext_cuda.c:81-260): frames are one device allocation each, planes laid out by
 * av_image_fill_arrays() with CUDA_FRAME_ALIGNMENT, AVFrame.data[] are device pointers, format AV_PIX_FMT_CUDA;
 * device memory comes from the product's C ABI (b200j2k_device_alloc = cudaMalloc on the primary context, which is
 * what AV_CUDA_USE_PRIMARY_CONTEXT gives the real implementation).  With it tests/test_gpu_ffcodec.py drives the
 * FFCodec shim exactly the way an application with a CUDA device / frames context would.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "libavutil/buffer.h"
#include "libavutil/frame.h"
#include "libavutil/hwcontext.h"
#include "libavutil/hwcontext_internal.h"
#include "libavutil/imgutils.h"
#include "libavutil/mem.h"
#include "libavutil/pixdesc.h"

#include "b200j2k.h"

#define CUDA_FRAME_ALIGNMENT 256        /* hwcontext_cuda.c:37 */

typedef struct StandinDevice {
    int ordinal;
} StandinDevice;

static void standin_buffer_free(void *opaque, uint8_t *data)
{
    (void)opaque;
    b200j2k_device_free(data);
}

static AVBufferRef *standin_pool_alloc(void *opaque, size_t size)
{
    AVHWFramesContext *ctx = opaque;
    StandinDevice *dev = ctx->device_ctx->hwctx;
    uint8_t *p = b200j2k_device_alloc(dev->ordinal, (int64_t)size);
    AVBufferRef *ref;
    if (!p)
        return NULL;
    ref = av_buffer_create(p, size, standin_buffer_free, NULL, 0);
    if (!ref)
        b200j2k_device_free(p);
    return ref;
}

static int standin_frames_init(AVHWFramesContext *ctx)
{
    if (!ctx->pool) {
        int size = av_image_get_buffer_size(ctx->sw_format, ctx->width, ctx->height, CUDA_FRAME_ALIGNMENT);
        if (size < 0)
            return size;
        ffhwframesctx(ctx)->pool_internal = av_buffer_pool_init2(size, ctx, standin_pool_alloc, NULL);
        if (!ffhwframesctx(ctx)->pool_internal)
            return AVERROR(ENOMEM);
    }
    return 0;
}

static int standin_get_buffer(AVHWFramesContext *ctx, AVFrame *frame)
{
    int res;
    frame->buf[0] = av_buffer_pool_get(ctx->pool);
    if (!frame->buf[0])
        return AVERROR(ENOMEM);
    res = av_image_fill_arrays(frame->data, frame->linesize, frame->buf[0]->data, ctx->sw_format, ctx->width,
                               ctx->height, CUDA_FRAME_ALIGNMENT);
    if (res < 0)
        return res;
    frame->format = AV_PIX_FMT_CUDA;
    frame->width  = ctx->width;
    frame->height = ctx->height;
    return 0;
}

static int standin_device_create(AVHWDeviceContext *ctx, const char *device, AVDictionary *opts, int flags)
{
    StandinDevice *dev = ctx->hwctx;
    (void)opts; (void)flags;
    dev->ordinal = device ? atoi(device) : 0;
    return dev->ordinal < b200j2k_device_count() ? 0 : AVERROR(ENODEV);
}

/* download of one frame into tightly packed host memory (what av_hwframe_transfer_data does plane by plane) */
int standin_download(const AVFrame *f, enum AVPixelFormat sw_format, uint8_t *out, int64_t cap)
{
    int linesizes[4];
    int64_t pos = 0;
    const AVPixFmtDescriptor *d = av_pix_fmt_desc_get(sw_format);
    if (av_image_fill_linesizes(linesizes, sw_format, f->width) < 0)
        return -1;
    for (int p = 0; p < 4 && f->data[p]; p++) {
        int h = f->height;
        if (p == 1 || p == 2)
            h = AV_CEIL_RSHIFT(f->height, d->log2_chroma_h);
        for (int y = 0; y < h; y++) {
            if (pos + linesizes[p] > cap)
                return -2;
            if (b200j2k_device_download(out + pos, f->data[p] + (int64_t)y * f->linesize[p], linesizes[p]) < 0)
                return -3;
            pos += linesizes[p];
        }
    }
    return (int)pos;
}

static const enum AVPixelFormat standin_pix_fmts[] = { AV_PIX_FMT_CUDA, AV_PIX_FMT_NONE };

const HWContextType ff_hwcontext_type_cuda = {
    .type              = AV_HWDEVICE_TYPE_CUDA,
    .name              = "CUDA (test stand-in over the b200j2k C ABI)",
    .device_hwctx_size = sizeof(StandinDevice),
    .device_create     = standin_device_create,
    .frames_init       = standin_frames_init,
    .frames_get_buffer = standin_get_buffer,
    .pix_fmts          = standin_pix_fmts,
};

/* libavutil/frame.c and dict.c come along with hwcontext.c; two of their string-parsing helpers would pull in
 * opt.c / eval.c / avsscanf.c, which nothing on this path calls */
int av_opt_get_key_value(const char **ropts, const char *key_val_sep, const char *pairs_sep, unsigned flags,
                         char **rkey, char **rval)
{
    (void)ropts; (void)key_val_sep; (void)pairs_sep; (void)flags; (void)rkey; (void)rval;
    return AVERROR(ENOSYS);
}
int av_sscanf(const char *string, const char *format, ...)
{
    (void)string; (void)format;
    return 0;
}
