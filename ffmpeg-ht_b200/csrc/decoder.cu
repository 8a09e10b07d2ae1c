/*
 * decoder.cu -- the C ABI (include/b200j2k.h): decoder contexts, slots, staging and the launch sequence.
 *
 * One b200j2k_decoder corresponds to one AVCodecContext of the reference (jpeg2000_decode_init /
 * jpeg2000_decode_frame, libavcodec/jpeg2000dec.c:2807-2908).  Per slot: a CUDA stream, one pinned
 * staging blob (descriptor tables + codestream bytes), device buffers that only grow, and the queue
 *     H2D blob -> ht_decode (cleanup-only blocks) -> ht_decode (blocks with refinement passes)
 *              -> idwt levels -> fused last level + MCT + pack -> D2H picture.
 * Nothing here falls back to the CPU: without a device every call returns B200J2K_ERR_DEVICE.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "../../include/b200j2k.h"
#include "j2k_kernels.h"
#include "j2k_parse.h"
#include "j2k_plan.h"

using namespace b200j2k;

namespace {

/* Device buffers grow on demand and are never shrunk.  With B200J2K_REDZONE=1 in the environment (debugging: the
 * fuzz tests run that way, compute-sanitizer being unavailable on the GPU pool) every buffer gets a guard zone on
 * either side, filled with a pattern that finish_slot() checks after each frame: a kernel writing next to its
 * buffer fails the decode loudly instead of corrupting a neighbour. */
constexpr size_t REDZONE = 4096;
constexpr int REDZONE_BYTE = 0xC5;
inline bool redzones_on()
{
    static const bool on = getenv("B200J2K_REDZONE") && atoi(getenv("B200J2K_REDZONE"));
    return on;
}

struct DevBuf {
    void *p = nullptr;      /* what the kernels see */
    void *base = nullptr;   /* what cudaMalloc returned (p - REDZONE when guard zones are on) */
    size_t cap = 0;
    int ensure(size_t n)
    {
        if (n <= cap)
            return 0;
        if (base)
            cudaFree(base);
        p = base = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 4096;
        const size_t guard = redzones_on() ? REDZONE : 0;
        if (guard)
            want = (want + 255) & ~(size_t)255;
        if (cudaMalloc(&base, want + 2 * guard) != cudaSuccess) {
            cudaGetLastError();
            base = nullptr;
            return B200J2K_ERR_NOMEM;
        }
        p = (uint8_t *)base + guard;
        cap = want;
        if (guard) {
            cudaMemset(base, REDZONE_BYTE, guard);
            cudaMemset((uint8_t *)p + cap, REDZONE_BYTE, guard);
        }
        return 0;
    }
    /* number of guard bytes a kernel has overwritten (0 when guard zones are off) */
    size_t violated(cudaStream_t st) const
    {
        if (!redzones_on() || !base)
            return 0;
        static thread_local uint8_t host[2 * REDZONE];
        if (cudaMemcpyAsync(host, base, REDZONE, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(host + REDZONE, (const uint8_t *)p + cap, REDZONE, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            return 1;
        size_t bad = 0;
        for (size_t i = 0; i < 2 * REDZONE; i++)
            bad += host[i] != REDZONE_BYTE;
        return bad;
    }
    void release()
    {
        if (base)
            cudaFree(base);
        p = base = nullptr;
        cap = 0;
    }
};

struct HostBuf {
    uint8_t *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n)
    {
        if (n <= cap)
            return 0;
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 4096;
        if (cudaHostAlloc((void **)&p, want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return B200J2K_ERR_NOMEM;
        }
        cap = want;
        return 0;
    }
    void release()
    {
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

/* How the code-blocks of one frame are spread over the HT kernels.  order[] lists descriptor indices as
 * [refine | coded | zero | wide | rwide | mq]:
 *   refine  blocks with SigProp / MagRef passes, <= 64 samples wide -> split pipeline + refinement kernels
 *   coded   cleanup-only blocks with data, <= 64 samples wide      -> split pipeline (ht_cleanup.cu)
 *   zero    blocks without data                                    -> zero-filled by the MagSgn kernel
 *   wide    cleanup-only blocks wider than 64 samples              -> general warp-per-block kernel
 *   rwide   refinement blocks wider than 64 samples                -> general kernel, refinement variant
 *   mq      Part-1 (EBCOT / MQ) blocks with data                   -> mq_block.cu (those without data: zero)
 * refine and coded are sorted by shape so that the 32 blocks sharing a warp of the lane-per-block kernels
 * run in step. */
struct HtPlan {
    std::vector<uint32_t> order, scr_off;
    int n_refine = 0, n_coded = 0, n_zero = 0, n_wide = 0, n_rwide = 0, n_mq = 0;
    int maxL_fast = 0, maxLref = 0, maxL_wide = 0, maxL_rwide = 0, mq_state_words = 0;
    size_t scratch_words = 0;
    /* descriptor indices sorted by quad-grid shape; depends on the geometry only */
    std::vector<uint32_t> by_shape;
    uint64_t shape_serial = ~0ull;

    static int klass(const CblkDesc &d)
    {
        if (d.flags & J2K_CBLK_SKIP)
            return 6;                               /* dropped by lowres: no kernel touches it */
        if (d.flags & J2K_CBLK_MQ)                  /* Part-1 block: own kernel; without data it only needs zeros */
            return d.npasses && d.length ? 5 : 2;
        const bool refine = d.npasses && (d.npasses % 3) != 1;
        /* the split pipeline sizes its shared memory by the longest segment of its class: a block whose lengths
         * are beyond anything 4096 samples can need (damaged or crafted packet headers) must not cost every other
         * block of the frame its occupancy, so it goes with the wide blocks */
        if (d.w > 64 || d.Lcup > 16384 || d.Lref > 8192)
            return refine ? 4 : 3;
        if (refine)
            return 0;
        return d.npasses ? 1 : 2;
    }

    void build(const CblkDesc *cblk, size_t n, uint64_t serial)
    {
        if (serial != shape_serial || by_shape.size() != n) {
            std::vector<uint32_t> key(n);
            by_shape.resize(n);
            for (size_t i = 0; i < n; i++) {
                by_shape[i] = (uint32_t)i;
                key[i] = ((uint32_t)(cblk[i].w + 1) >> 1) << 16 | ((uint32_t)(cblk[i].h + 1) >> 1);
            }
            std::stable_sort(by_shape.begin(), by_shape.end(), [&](uint32_t a, uint32_t b) { return key[a] > key[b]; });
            shape_serial = serial;
        }
        order.resize(n);
        scr_off.resize(n);
        maxL_fast = maxLref = maxL_wide = maxL_rwide = mq_state_words = 0;
        scratch_words = 0;
        size_t cnt[7] = { 0, 0, 0, 0, 0, 0, 0 };
        for (size_t i = 0; i < n; i++)
            cnt[klass(cblk[i])]++;
        size_t at[7];
        at[0] = 0;
        for (int k = 1; k < 7; k++)
            at[k] = at[k - 1] + cnt[k - 1];
        for (size_t k = 0; k < n; k++) {
            const uint32_t i = by_shape[k];
            const CblkDesc &d = cblk[i];
            const int c = klass(d);
            order[at[c]++] = i;
            scr_off[i] = 0;
            if (c <= 1) {
                scr_off[i] = (uint32_t)scratch_words;
                scratch_words += ht_cleanup_scratch_words(d);
                if (d.Lcup > maxL_fast) maxL_fast = d.Lcup;
                if (c == 0 && d.Lref > maxLref) maxLref = d.Lref;
            } else if (c == 3) {
                if (d.Lcup > maxL_wide) maxL_wide = d.Lcup;
            } else if (c == 4) {
                if (d.Lcup > maxL_rwide) maxL_rwide = d.Lcup;
            } else if (c == 5) {
                const int words = mq_state_words_of(d.w, d.h);
                if (words > mq_state_words) mq_state_words = words;
            }
        }
        n_refine = (int)cnt[0];
        n_coded = (int)cnt[1];
        n_zero = (int)cnt[2];
        n_wide = (int)cnt[3];
        n_rwide = (int)cnt[4];
        n_mq = (int)cnt[5];
        if (n_mq > 1) {
            /* a Part-1 block costs about what its segment is long: start the long ones first, so that the last
             * wave of warps is made of short blocks (longest-processing-time-first) */
            uint32_t *mq = order.data() + (n - cnt[6] - cnt[5]);
            std::stable_sort(mq, mq + n_mq, [&](uint32_t a, uint32_t b) { return cblk[a].length > cblk[b].length; });
        }
    }
};

/* queue the HT kernels of one plan; returns launches or < 0 */
int run_ht(cudaStream_t st, const HtPlan &hp, const CblkDesc *d_cblk, const uint32_t *d_order, const uint32_t *d_scr,
           const uint8_t *d_bytes, const BandDesc *d_bands, void *scratch, void *coef, int32_t *status, int max_w,
           int max_area, int max_state, cudaEvent_t *ev = nullptr)
{
    /* ev (optional): [1] after un-stuff, [2] after VLC, [3] after MagSgn + wide blocks */
    int launches = 0, r;
    r = launch_ht_cleanup(st, d_cblk, d_order, d_scr, hp.n_refine, hp.n_coded, hp.n_zero, d_bytes, d_bands, scratch, coef,
                          status, hp.maxL_fast, hp.maxLref, ev ? ev[1] : nullptr, ev ? ev[2] : nullptr);
    if (r < 0) return r;
    launches += r;
    const uint32_t *o_wide = d_order + hp.n_refine + hp.n_coded + hp.n_zero;
    r = launch_ht_decode(st, false, d_cblk, o_wide, hp.n_wide, d_bytes, d_bands, coef, status, hp.maxL_wide, max_w, max_area,
                         max_state);
    if (r < 0) return r;
    launches += r;
    r = launch_mq_blocks(st, d_cblk, o_wide + hp.n_wide + hp.n_rwide, hp.n_mq, d_bytes, d_bands, coef, status,
                         hp.mq_state_words);
    if (r < 0) return r;
    launches += r;
    if (ev) cudaEventRecord(ev[3], st);
    r = launch_ht_refine(st, d_cblk, d_order, d_scr, hp.n_refine, d_bytes, d_bands, scratch, coef, status, hp.maxL_fast);
    if (r < 0) return r;
    launches += r;
    r = launch_ht_decode(st, true, d_cblk, o_wide + hp.n_wide, hp.n_rwide, d_bytes, d_bands, coef, status, hp.maxL_rwide,
                         max_w, max_area, max_state);
    if (r < 0) return r;
    return launches + r;
}

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;       /* recorded behind the host->device copies of a frame (b200j2k_wait_upload) */
    Parser parser;
    HostBuf stage;          /* pinned: tables + bytes */
    HostBuf out_stage;      /* pinned: picture when the caller's planes are pageable */
    DevBuf d_blob, d_coefA, d_coefB, d_coefC, d_picture, d_status, d_scratch;
    HtPlan ht;
    IdwtStreamCache *idwt_cache = nullptr;
    size_t off_scr = 0;
    /* offsets inside the blob */
    size_t off_cblk = 0, off_bands = 0, off_tcs = 0, off_groups = 0, off_order = 0, off_bytes = 0, blob_bytes = 0;
    bool busy = false, staged = false;
    int32_t *h_failed = nullptr;   /* pinned: status[-1] of the last frame (blocks zero-filled for damaged data) */
    size_t h2d_bytes = 0;
    int launches = 0;
    /* pending output */
    uint8_t *dst[4] = { nullptr, nullptr, nullptr, nullptr };
    int dst_linesize[4] = { 0, 0, 0, 0 };
    bool direct = false, subset_out = false;
    uint64_t seq = 0;
};

}  // namespace

struct b200j2k_decoder {
    int device = 0;
    int flags = 0;
    std::vector<Slot> slots;
    uint64_t seq = 0;
};

namespace {

int device_ok()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void fill_info(const Parser &p, b200j2k_frame_info *info)
{
    if (!info)
        return;
    memset(info, 0, sizeof(*info));
    info->width = p.hdr.out_w;
    info->height = p.hdr.out_h;
    info->pix_fmt = p.hdr.pix;
    info->ncomponents = p.hdr.ncomp;
    info->precision = p.hdr.precision;
    info->lossless = p.hdr.lossless;
    info->profile = p.hdr.rsiz;
    info->nplanes = p.nplanes;
    for (int i = 0; i < 4; i++) {
        info->plane_width[i] = p.plane_w[i];
        info->plane_height[i] = p.plane_h[i];
        info->plane_row_bytes[i] = p.plane_row_bytes[i];
    }
    info->ntiles = p.hdr.ntx * p.hdr.nty;
    info->ncodeblocks = (int32_t)p.cblk.size();
    info->bytes_consumed = p.consumed;
    strncpy(info->av_pix_fmt_name, p.hdr.pix_name, sizeof(info->av_pix_fmt_name) - 1);
    info->sar_num = p.hdr.sar_num && p.hdr.sar_den ? p.hdr.sar_num : 0;      /* published only when both are set (:2867) */
    info->sar_den = p.hdr.sar_num && p.hdr.sar_den ? p.hdr.sar_den : 0;
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

/* memory the copy engines can write without staging: pinned host memory, or device / managed memory
 * (AV_PIX_FMT_CUDA-style frames whose data[] are device pointers: the picture never crosses PCIe) */
bool is_dma_target(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

/* host parse + build the staging blob + H2D */
int stage_frame(b200j2k_decoder *dec, Slot &s, const uint8_t *pkt, int64_t size, b200j2k_frame_info *info)
{
    int ret = s.parser.parse(pkt, size, dec->flags, false);
    if (ret < 0)
        return ret;
    Parser &p = s.parser;
    fill_info(p, info);
    const size_t ncb = p.cblk.size();
    s.ht.build(p.cblk.data(), ncb, p.geom_serial);
    const std::vector<uint32_t> &order = s.ht.order;
    /* blob layout on the device: [tables][codestream bytes][gathered bytes][64 B zero slack].  The tables
     * always go through the pinned staging buffer; the codestream is DMA'd straight from the caller's
     * packet when that memory is pinned (registered), else it is copied into the staging buffer first. */
    const bool direct_in = is_pinned(pkt);
    size_t off = 0;
    s.off_cblk = off;   off = align_up(off + ncb * sizeof(CblkDesc), 256);
    s.off_bands = off;  off = align_up(off + p.bands.size() * sizeof(BandDesc), 256);
    s.off_tcs = off;    off = align_up(off + p.tcs.size() * sizeof(TileCompDesc), 256);
    s.off_groups = off; off = align_up(off + p.groups.size() * sizeof(PackGroup), 256);
    s.off_order = off;  off = align_up(off + ncb * sizeof(uint32_t), 256);
    s.off_scr = off;    off = align_up(off + ncb * sizeof(uint32_t), 256);
    const bool pal8 = p.nplanes == 2 && !strcmp(p.hdr.pix_name, "pal8");
    const size_t off_pal = off;
    if (pal8)                                             /* the palette travels with the tables */
        off = align_up(off + sizeof(p.hdr.palette), 256);
    s.off_bytes = off;
    const size_t tail = p.extra.size() + 64;              /* gathered bytes + slack, placed behind the codestream */
    const size_t off_tail = s.off_bytes + (size_t)size;
    off = align_up(off_tail + tail, 256);
    s.blob_bytes = off;
    /* staging: tables, then (staged path) the codestream, then the tail */
    const size_t stage_need = direct_in ? s.off_bytes + tail : off;
    if ((ret = s.stage.ensure(stage_need)) < 0)
        return ret;
    uint8_t *h = s.stage.p;
    memcpy(h + s.off_cblk, p.cblk.data(), ncb * sizeof(CblkDesc));
    memcpy(h + s.off_bands, p.bands.data(), p.bands.size() * sizeof(BandDesc));
    memcpy(h + s.off_tcs, p.tcs.data(), p.tcs.size() * sizeof(TileCompDesc));
    memcpy(h + s.off_groups, p.groups.data(), p.groups.size() * sizeof(PackGroup));
    memcpy(h + s.off_order, order.data(), ncb * sizeof(uint32_t));
    memcpy(h + s.off_scr, s.ht.scr_off.data(), ncb * sizeof(uint32_t));
    if (pal8)
        memcpy(h + off_pal, p.hdr.palette, sizeof(p.hdr.palette));
    uint8_t *htail = direct_in ? h + s.off_bytes : h + off_tail;
    if (!p.extra.empty())
        memcpy(htail, p.extra.data(), p.extra.size());
    memset(htail + p.extra.size(), 0, 64);

    if ((ret = s.d_blob.ensure(off)) < 0) return ret;
    if ((ret = s.d_coefA.ensure(p.coef_elems * 4 + 256)) < 0) return ret;
    if (p.max_levels >= 2) {
        if ((ret = s.d_coefB.ensure(p.ll_elems * 4 + 256)) < 0) return ret;
        if ((ret = s.d_coefC.ensure(p.ll_elems * 4 + 256)) < 0) return ret;
    }
    if ((ret = s.d_picture.ensure(p.picture_bytes + 256)) < 0) return ret;
    if ((ret = s.d_status.ensure(ncb * 4 + 256 + 16)) < 0) return ret;     /* 16 bytes in front: the failure counter */
    if ((ret = s.d_scratch.ensure(s.ht.scratch_words * 4 + 256)) < 0) return ret;
    uint8_t *dblob = (uint8_t *)s.d_blob.p;
    cudaError_t e;
    /* the part of the packet the kernels will read: all of it, or -- for a context that decodes a subset of the tiles
     * (B200J2K_FLAG_TILES) -- the tile-parts of its own tiles, at unchanged offsets */
    std::vector<std::pair<uint32_t, uint32_t>> whole(1, std::make_pair(0u, (uint32_t)size));
    const std::vector<std::pair<uint32_t, uint32_t>> &ranges = p.tile_subset ? p.needed_ranges : whole;
    size_t uploaded = 0;
    e = cudaMemcpyAsync(dblob, h, s.off_bytes, cudaMemcpyHostToDevice, s.stream);       /* the tables go first: they are ready */
    for (const auto &rg : ranges) {
        const size_t at0 = rg.first, len = rg.second;
        uploaded += len;
        if (direct_in) {
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(dblob + s.off_bytes + at0, pkt + at0, len, cudaMemcpyHostToDevice, s.stream);
            continue;
        }
        /* pageable packet: copied into the pinned staging buffer piece by piece, each piece's DMA running while the
         * next one is copied */
        const size_t piece = (size_t)4 << 20;
        for (size_t at = at0; at < at0 + len && e == cudaSuccess; at += piece) {
            const size_t n = at0 + len - at < piece ? at0 + len - at : piece;
            memcpy(h + s.off_bytes + at, pkt + at, n);
            e = cudaMemcpyAsync(dblob + s.off_bytes + at, h + s.off_bytes + at, n, cudaMemcpyHostToDevice, s.stream);
        }
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dblob + off_tail, htail, tail, cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess)
        e = cudaEventRecord(s.done, s.stream);
    /* PAL8: plane 1 of the picture is the palette (AVFrame.data[1], jpeg2000dec.c:2900-2901); no kernel writes it */
    if (e == cudaSuccess && pal8)
        e = cudaMemcpyAsync((uint8_t *)s.d_picture.p + p.plane_off[1], dblob + off_pal, sizeof(p.hdr.palette),
                            cudaMemcpyDeviceToDevice, s.stream);
    if (e != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    s.h2d_bytes = s.off_bytes + uploaded + tail;
    s.staged = true;
    return 0;
}

/* queue every kernel of the frame on the slot's stream; ev (optional, 7 events) brackets the stages:
 * [0] start [1] un-stuffed [2] VLC done [3] MagSgn (+ wide blocks) done [4] refinement blocks done
 * [5] IDWT levels done [6] end */
int run_kernels(Slot &s, cudaEvent_t *ev = nullptr)
{
    Parser &p = s.parser;
    const uint8_t *blob = (const uint8_t *)s.d_blob.p;
    const CblkDesc *d_cblk = (const CblkDesc *)(blob + s.off_cblk);
    const BandDesc *d_bands = (const BandDesc *)(blob + s.off_bands);
    const TileCompDesc *d_tcs = (const TileCompDesc *)(blob + s.off_tcs);
    const PackGroup *d_groups = (const PackGroup *)(blob + s.off_groups);
    const uint32_t *d_order = (const uint32_t *)(blob + s.off_order);
    const uint8_t *d_bytes = blob + s.off_bytes;
    int launches = 0, r;
    if (cudaMemsetAsync((int32_t *)s.d_status.p + 3, 0, 4, s.stream) != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    if (ev) cudaEventRecord(ev[0], s.stream);
    r = run_ht(s.stream, s.ht, d_cblk, d_order, (const uint32_t *)(blob + s.off_scr), d_bytes, d_bands, s.d_scratch.p,
               s.d_coefA.p, (int32_t *)s.d_status.p + 4, (int)p.max_w, (int)p.max_area, (int)p.max_state, ev);
    if (r < 0) return B200J2K_ERR_DEVICE;
    launches += r;
    if (ev) cudaEventRecord(ev[4], s.stream);
    IdwtFrame f;
    memset(&f, 0, sizeof(f));
    f.before_final = ev ? ev[5] : nullptr;
    f.d_tcs = d_tcs;
    f.d_groups = d_groups;
    f.h_tcs = p.tcs.data();
    f.h_groups = p.groups.data();
    f.ntcs = (int)p.tcs.size();
    f.ngroups = (int)p.groups.size();
    f.coefA = s.d_coefA.p;
    f.coefB = s.d_coefB.p;
    f.coefC = s.d_coefC.p;
    f.picture = (uint8_t *)s.d_picture.p;
    if (!s.idwt_cache)
        s.idwt_cache = idwt_stream_cache_create();
    f.stream_cache = s.idwt_cache;
    f.geom_serial = p.geom_serial;
    for (int i = 0; i < 4; i++) {
        f.plane_off[i] = p.plane_off[i];
        f.plane_row_bytes[i] = p.plane_row_bytes[i];
    }
    r = launch_idwt_frame(s.stream, f);
    if (r < 0) return B200J2K_ERR_DEVICE;
    launches += r;
    if (ev) cudaEventRecord(ev[6], s.stream);
    s.launches = launches;
    return 0;
}

int queue_output(Slot &s, uint8_t *const data[4], const int linesize[4])
{
    Parser &p = s.parser;
    bool direct = true;
    if (!s.h_failed && cudaHostAlloc((void **)&s.h_failed, 16, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        s.h_failed = nullptr;
        return B200J2K_ERR_NOMEM;
    }
    if (cudaMemcpyAsync(s.h_failed, (int32_t *)s.d_status.p + 3, 4, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    for (int i = 0; i < p.nplanes; i++) {
        if (!data[i] || (linesize[i] < p.plane_row_bytes[i] && p.plane_h[i] > 1))   /* a one-row plane (palette)
                                                                                     * has no meaningful linesize */
            return B200J2K_ERR_EINVAL;
        direct = direct && is_dma_target(data[i]);
        s.dst[i] = data[i];
        s.dst_linesize[i] = linesize[i];
    }
    s.direct = direct;
    const uint8_t *pic = (const uint8_t *)s.d_picture.p;
    if (p.tile_subset) {
        /* only this context's tiles were reconstructed: only their rectangles leave the device, and only into memory
         * the copy engines reach directly (several contexts fill one frame side by side) */
        if (!direct)
            return B200J2K_ERR_EINVAL;
        s.subset_out = true;
        for (const PackGroup &g : p.groups) {
            for (int c = 0; c < g.ncomp; c++) {
                const TileCompDesc &tc = p.tcs[g.tc[c]];
                if (tc.transform < 0)
                    break;
                const bool packed = g.pixel_stride == g.ncomp && g.ncomp > 1;
                if (packed && c > 0)
                    break;                                            /* one rectangle holds all its components */
                const int pl = packed || g.ncomp == 1 ? g.plane : g.comp_plane[c];
                const LevelDesc &o = tc.lev[tc.nlevels];
                const size_t bpp = (size_t)(g.wide ? 2 : 1) * (packed ? g.ncomp : 1);
                const size_t wbytes = (size_t)(o.u1 - o.u0) * bpp, rows = (size_t)(o.v1 - o.v0);
                if (!wbytes || !rows)
                    continue;
                const size_t src = p.plane_off[pl] + (size_t)tc.out_y * p.plane_row_bytes[pl] + (size_t)tc.out_x * bpp;
                uint8_t *dst = data[pl] + (size_t)tc.out_y * linesize[pl] + (size_t)tc.out_x * bpp;
                if (cudaMemcpy2DAsync(dst, (size_t)linesize[pl], pic + src, (size_t)p.plane_row_bytes[pl], wbytes, rows,
                                      cudaMemcpyDefault, s.stream) != cudaSuccess)
                    return B200J2K_ERR_DEVICE;
            }
        }
        return 0;
    }
    if (direct) {
        for (int i = 0; i < p.nplanes; i++) {
            cudaError_t e;
            if (linesize[i] == p.plane_row_bytes[i] || p.plane_h[i] == 1)     /* no row padding: one linear copy */
                e = cudaMemcpyAsync(data[i], pic + p.plane_off[i], (size_t)p.plane_row_bytes[i] * p.plane_h[i],
                                    cudaMemcpyDefault, s.stream);
            else
                e = cudaMemcpy2DAsync(data[i], (size_t)linesize[i], pic + p.plane_off[i], (size_t)p.plane_row_bytes[i],
                                      (size_t)p.plane_row_bytes[i], (size_t)p.plane_h[i], cudaMemcpyDefault, s.stream);
            if (e != cudaSuccess)
                return B200J2K_ERR_DEVICE;
        }
    } else {
        int ret = s.out_stage.ensure(p.picture_bytes);
        if (ret < 0)
            return ret;
        if (cudaMemcpyAsync(s.out_stage.p, pic, p.picture_bytes, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess)
            return B200J2K_ERR_DEVICE;
    }
    return 0;
}

int finish_slot(Slot &s)
{
    if (!s.busy)
        return 0;
    cudaError_t e = cudaStreamSynchronize(s.stream);
    s.busy = false;
    if (e != cudaSuccess) {
        cudaGetLastError();
        return B200J2K_ERR_DEVICE;
    }
    if (redzones_on()) {
        const DevBuf *bufs[] = { &s.d_blob, &s.d_coefA, &s.d_coefB, &s.d_coefC, &s.d_picture, &s.d_status, &s.d_scratch };
        const char *names[] = { "blob", "coefA", "coefB", "coefC", "picture", "status", "scratch" };
        size_t bad = 0;
        for (int i = 0; i < 7; i++) {
            const size_t b = bufs[i]->violated(s.stream);
            if (b)
                fprintf(stderr, "b200j2k: %zu guard bytes of the %s buffer were overwritten\n", b, names[i]);
            bad += b;
        }
        const int asserts = ht_cleanup_bounds_failures();
        if (asserts) {
            fprintf(stderr, "b200j2k: %d in-kernel bounds assertions failed\n", asserts);
            bad += (size_t)(asserts > 0 ? asserts : 1);
        }
        if (bad)
            return B200J2K_ERR_DEVICE;
    }
    if (!s.direct && s.dst[0]) {
        Parser &p = s.parser;
        for (int i = 0; i < p.nplanes; i++) {
            const uint8_t *src = s.out_stage.p + p.plane_off[i];
            if (s.dst_linesize[i] == p.plane_row_bytes[i] || p.plane_h[i] == 1) {
                memcpy(s.dst[i], src, (size_t)p.plane_row_bytes[i] * p.plane_h[i]);
                continue;
            }
            for (int y = 0; y < p.plane_h[i]; y++)
                memcpy(s.dst[i] + (size_t)y * s.dst_linesize[i], src + (size_t)y * p.plane_row_bytes[i],
                       (size_t)p.plane_row_bytes[i]);
        }
    }
    s.dst[0] = nullptr;
    return 0;
}

}  // namespace

extern "C" {

int b200j2k_device_count(void) { return device_ok(); }

int b200j2k_create(b200j2k_decoder **out, int device, int slots, int flags)
{
    if (!out)
        return B200J2K_ERR_EINVAL;
    *out = nullptr;
    const int n = device_ok();
    if (n <= 0 || device < 0 || device >= n)
        return B200J2K_ERR_DEVICE;
    if (slots < 1) slots = 1;
    if (slots > 64) slots = 64;
    if (cudaSetDevice(device) != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    b200j2k_decoder *d = new (std::nothrow) b200j2k_decoder();
    if (!d)
        return B200J2K_ERR_NOMEM;
    d->device = device;
    d->flags = flags;
    d->slots.resize(slots);
    for (auto &s : d->slots) {
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) {
            b200j2k_destroy(d);
            return B200J2K_ERR_DEVICE;
        }
    }
    *out = d;
    return 0;
}

void b200j2k_destroy(b200j2k_decoder *d)
{
    if (!d)
        return;
    cudaSetDevice(d->device);
    for (auto &s : d->slots) {
        if (s.stream)
            cudaStreamSynchronize(s.stream);
        s.stage.release();
        s.out_stage.release();
        if (s.h_failed)
            cudaFreeHost(s.h_failed);
        s.h_failed = nullptr;
        s.d_blob.release(); s.d_coefA.release(); s.d_coefB.release(); s.d_coefC.release();
        s.d_picture.release(); s.d_status.release(); s.d_scratch.release();
        idwt_stream_cache_destroy(s.idwt_cache);
        s.idwt_cache = nullptr;
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete d;
}

int b200j2k_probe(const uint8_t *pkt, int64_t pkt_size, int flags, b200j2k_frame_info *info)
{
    Parser p;
    int ret = p.parse(pkt, pkt_size, flags, true);
    if (ret < 0)
        return ret;
    fill_info(p, info);
    return 0;
}

int b200j2k_submit(b200j2k_decoder *d, const uint8_t *pkt, int64_t pkt_size, uint8_t *const data[4],
                   const int linesize[4], b200j2k_frame_info *info)
{
    if (!d || !pkt || !data || !linesize)
        return B200J2K_ERR_EINVAL;
    if (cudaSetDevice(d->device) != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    /* pick a free slot, else the oldest busy one */
    int pick = -1;
    for (size_t i = 0; i < d->slots.size(); i++)
        if (!d->slots[i].busy) { pick = (int)i; break; }
    if (pick < 0) {
        pick = 0;
        for (size_t i = 1; i < d->slots.size(); i++)
            if (d->slots[i].seq < d->slots[pick].seq) pick = (int)i;
        int r = finish_slot(d->slots[pick]);
        if (r < 0)
            return r;
    }
    Slot &s = d->slots[pick];
    int ret = stage_frame(d, s, pkt, pkt_size, info);
    if (ret >= 0 && (ret = run_kernels(s)) >= 0)
        ret = queue_output(s, data, linesize);
    if (ret < 0) {
        /* copies from the caller's packet or from the staging buffer may be in flight: nobody may touch or
         * free either before the stream has drained */
        if (cudaStreamSynchronize(s.stream) != cudaSuccess)
            cudaGetLastError();
        s.staged = false;
        return ret;
    }
    s.busy = true;
    s.seq = ++d->seq;
    return pick;
}

int b200j2k_wait(b200j2k_decoder *d, int slot)
{
    if (!d || slot < 0 || slot >= (int)d->slots.size())
        return B200J2K_ERR_EINVAL;
    if (cudaSetDevice(d->device) != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    return finish_slot(d->slots[slot]);
}

int b200j2k_wait_upload(b200j2k_decoder *d, int slot)
{
    if (!d || slot < 0 || slot >= (int)d->slots.size())
        return B200J2K_ERR_EINVAL;
    if (cudaSetDevice(d->device) != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    return cudaEventSynchronize(d->slots[slot].done) == cudaSuccess ? 0 : B200J2K_ERR_DEVICE;
}

int b200j2k_pointer_is_pinned(const void *p) { return device_ok() > 0 && is_pinned(p); }

int64_t b200j2k_decode(b200j2k_decoder *d, const uint8_t *pkt, int64_t pkt_size, uint8_t *const data[4],
                       const int linesize[4], b200j2k_frame_info *info)
{
    b200j2k_frame_info local;
    int slot = b200j2k_submit(d, pkt, pkt_size, data, linesize, &local);
    if (slot < 0)
        return slot;
    int r = b200j2k_wait(d, slot);
    if (r < 0)
        return r;
    local.nblocks_failed = b200j2k_slot_blocks_failed(d, slot);
    if (info)
        *info = local;
    return local.bytes_consumed;
}

int b200j2k_stage(b200j2k_decoder *d, int slot, const uint8_t *pkt, int64_t pkt_size, b200j2k_frame_info *info)
{
    if (!d || slot < 0 || slot >= (int)d->slots.size())
        return B200J2K_ERR_EINVAL;
    if (cudaSetDevice(d->device) != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    Slot &s = d->slots[slot];
    int r = finish_slot(s);
    if (r < 0)
        return r;
    r = stage_frame(d, s, pkt, pkt_size, info);
    if (r < 0)
        return r;
    return cudaStreamSynchronize(s.stream) == cudaSuccess ? 0 : B200J2K_ERR_DEVICE;
}

int b200j2k_run_staged(b200j2k_decoder *d, int slot)
{
    if (!d || slot < 0 || slot >= (int)d->slots.size() || !d->slots[slot].staged)
        return B200J2K_ERR_EINVAL;
    if (cudaSetDevice(d->device) != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    return run_kernels(d->slots[slot]);
}

int b200j2k_run_staged_timed(b200j2k_decoder *d, int slot, float *ms)
{
    if (!d || !ms || slot < 0 || slot >= (int)d->slots.size() || !d->slots[slot].staged)
        return B200J2K_ERR_EINVAL;
    if (cudaSetDevice(d->device) != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    cudaEvent_t ev[7];
    for (auto &e : ev)
        if (cudaEventCreate(&e) != cudaSuccess)
            return B200J2K_ERR_DEVICE;
    Slot &s = d->slots[slot];
    int r = run_kernels(s, ev);
    if (r < 0)
        fprintf(stderr, "b200j2k_run_staged_timed: launch failed (%d): %s\n", r, cudaGetErrorString(cudaGetLastError()));
    cudaError_t se = cudaSuccess;
    if (r >= 0 && (se = cudaStreamSynchronize(s.stream)) != cudaSuccess) {
        fprintf(stderr, "b200j2k_run_staged_timed: CUDA error: %s\n", cudaGetErrorString(se));
        r = B200J2K_ERR_DEVICE;
    }
    if (r >= 0)
        for (int i = 0; i < 6; i++)
            cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
    for (auto &e : ev)
        cudaEventDestroy(e);
    return r;
}

int b200j2k_sync(b200j2k_decoder *d, int slot)
{
    if (!d || slot < 0 || slot >= (int)d->slots.size())
        return B200J2K_ERR_EINVAL;
    cudaError_t e = cudaStreamSynchronize(d->slots[slot].stream);
    if (e != cudaSuccess) {
        fprintf(stderr, "b200j2k: CUDA error: %s\n", cudaGetErrorString(e));
        return B200J2K_ERR_DEVICE;
    }
    return 0;
}

int b200j2k_fetch(b200j2k_decoder *d, int slot, uint8_t *const data[4], const int linesize[4])
{
    if (!d || slot < 0 || slot >= (int)d->slots.size() || !d->slots[slot].staged)
        return B200J2K_ERR_EINVAL;
    if (cudaSetDevice(d->device) != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    Slot &s = d->slots[slot];
    int r = queue_output(s, data, linesize);
    if (r < 0)
        return r;
    s.busy = true;
    return finish_slot(s);
}

void *b200j2k_slot_stream(b200j2k_decoder *d, int slot)
{
    if (!d || slot < 0 || slot >= (int)d->slots.size())
        return nullptr;
    return (void *)d->slots[slot].stream;
}

void *b200j2k_slot_output(b200j2k_decoder *d, int slot, int64_t *bytes)
{
    if (!d || slot < 0 || slot >= (int)d->slots.size())
        return nullptr;
    if (bytes)
        *bytes = (int64_t)d->slots[slot].parser.picture_bytes;
    return d->slots[slot].d_picture.p;
}

int64_t b200j2k_slot_h2d_bytes(b200j2k_decoder *d, int slot)
{
    if (!d || slot < 0 || slot >= (int)d->slots.size())
        return 0;
    return (int64_t)d->slots[slot].h2d_bytes;
}

int b200j2k_slot_blocks_failed(b200j2k_decoder *d, int slot)
{
    if (!d || slot < 0 || slot >= (int)d->slots.size())
        return B200J2K_ERR_EINVAL;
    const Slot &s = d->slots[slot];
    return s.busy || !s.h_failed ? 0 : *s.h_failed;
}

/* debugging aid for the guard-zone checker (tests/test_gpu_debug_bounds.py): writes one byte right behind the slot's
 * picture buffer, i.e. into its guard zone when B200J2K_REDZONE=1 -- refused otherwise */
int b200j2k_debug_poke(b200j2k_decoder *d, int slot)
{
    if (!d || slot < 0 || slot >= (int)d->slots.size() || !redzones_on() || !d->slots[slot].d_picture.p)
        return B200J2K_ERR_EINVAL;
    const DevBuf &b = d->slots[slot].d_picture;
    return cudaMemset((uint8_t *)b.p + b.cap + 7, 0, 1) == cudaSuccess ? 0 : B200J2K_ERR_DEVICE;
}

int b200j2k_slot_launches(b200j2k_decoder *d, int slot)
{
    if (!d || slot < 0 || slot >= (int)d->slots.size())
        return 0;
    return d->slots[slot].launches;
}

/* ---- stage-level entry points -------------------------------------------------------------------------- */
int b200j2k_ht_blocks(int device, const uint8_t *data, int64_t data_size, const int32_t *par, int n,
                      int32_t *out, int32_t *status)
{
    if (device_ok() <= 0 || cudaSetDevice(device) != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    if (n <= 0)
        return 0;
    std::vector<CblkDesc> descs(n);
    BandDesc band;
    memset(&band, 0, sizeof(band));
    band.i_stepsize = 32768;
    size_t total = 0;
    int maxw = 1, max_area = 1, max_state = 16;
    for (int i = 0; i < n; i++) {
        const int32_t *q = par + 10 * i;
        CblkDesc &d = descs[i];
        memset(&d, 0, sizeof(d));
        d.npasses = (uint8_t)q[0];
        d.Lcup = (uint16_t)q[1];
        d.Lref = (uint16_t)q[2];
        d.zbp = (uint8_t)q[3];
        d.w = (uint16_t)q[4];
        d.h = (uint16_t)q[5];
        /* the stage-level entry wants t1->data, i.e. the sample BEFORE dequantisation but after the ROI
         * shift: dequant mode 3 = raw (handled in the kernels) */
        d.M_b = (uint8_t)q[6];
        d.roi_shift = (uint8_t)q[7];
        d.flags = (uint8_t)((q[8] ? 1 : 0) | (3 << 1));
        d.data_off = (uint32_t)q[9];
        d.length = (uint32_t)(q[1] + q[2]);
        d.dst_off = (uint32_t)total;
        d.dst_stride = d.w;
        total += (size_t)d.w * d.h;
        if (d.w > maxw) maxw = d.w;
        if (d.w * d.h > max_area) max_area = d.w * d.h;
        if ((d.w + 3) * (d.h + 3) > max_state) max_state = (d.w + 3) * (d.h + 3);
        if ((int64_t)d.data_off + d.length > data_size)
            return B200J2K_ERR_EINVAL;
    }
    HtPlan hp;
    hp.build(descs.data(), (size_t)n, 0);
    uint8_t *d_bytes = nullptr;
    CblkDesc *d_descs = nullptr;
    uint32_t *d_order = nullptr, *d_scr = nullptr, *d_out = nullptr, *d_scratch = nullptr;
    BandDesc *d_band = nullptr;
    int32_t *d_status = nullptr;
    int ret = 0;
    cudaStream_t st = nullptr;
    if (cudaStreamCreate(&st) != cudaSuccess) return B200J2K_ERR_DEVICE;
    /* 16 bytes in front of the segments: the kernels fetch bytes as aligned words and may touch up to 6 bytes
     * before a segment (in a frame the descriptor tables sit there) */
    uint8_t *d_bytes_alloc = nullptr;
    if (cudaMalloc(&d_bytes_alloc, (size_t)data_size + 16 + 64) != cudaSuccess || cudaMalloc(&d_descs, n * sizeof(CblkDesc)) != cudaSuccess ||
        cudaMalloc(&d_order, n * 4) != cudaSuccess || cudaMalloc(&d_scr, n * 4) != cudaSuccess ||
        cudaMalloc(&d_out, total * 4 + 4) != cudaSuccess || cudaMalloc(&d_scratch, hp.scratch_words * 4 + 256) != cudaSuccess ||
        cudaMalloc(&d_band, sizeof(BandDesc)) != cudaSuccess || cudaMalloc(&d_status, n * 4 + 16) != cudaSuccess) {
        ret = B200J2K_ERR_NOMEM;
    } else {
        d_bytes = d_bytes_alloc + 16;
        cudaMemsetAsync(d_bytes_alloc, 0, (size_t)data_size + 16 + 64, st);
        cudaMemcpyAsync(d_bytes, data, (size_t)data_size, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_descs, descs.data(), n * sizeof(CblkDesc), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_order, hp.order.data(), n * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_scr, hp.scr_off.data(), n * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_band, &band, sizeof(band), cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(d_out, 0x55, total * 4 + 4, st);
        cudaMemsetAsync(d_status, 0, 16, st);
        if (run_ht(st, hp, d_descs, d_order, d_scr, d_bytes, d_band, d_scratch, d_out, d_status + 4, maxw, max_area, max_state) < 0)
            ret = B200J2K_ERR_DEVICE;
        cudaMemcpyAsync(out, d_out, total * 4, cudaMemcpyDeviceToHost, st);
        if (status) cudaMemcpyAsync(status, d_status + 4, n * 4, cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            fprintf(stderr, "b200j2k_ht_blocks: CUDA error: %s\n", cudaGetErrorString(e));
            ret = B200J2K_ERR_DEVICE;
        }
    }
    if (ret == 0 && getenv("B200J2K_DUMP_SCRATCH")) {
        /* kernel debugging: the intermediate per-block scratch (un-stuffed VLC / MEL bits, quad words) of this
         * call, as [n][scr_off...][words...] 32-bit words */
        std::vector<uint32_t> h(hp.scratch_words + 64);
        cudaMemcpy(h.data(), d_scratch, hp.scratch_words * 4, cudaMemcpyDeviceToHost);
        if (FILE *f = fopen(getenv("B200J2K_DUMP_SCRATCH"), "wb")) {
            const uint32_t hdr[2] = { (uint32_t)n, (uint32_t)hp.scratch_words };
            fwrite(hdr, 4, 2, f);
            fwrite(hp.scr_off.data(), 4, (size_t)n, f);
            fwrite(h.data(), 4, hp.scratch_words, f);
            fclose(f);
        }
    }
    cudaFree(d_bytes_alloc); cudaFree(d_descs); cudaFree(d_order); cudaFree(d_scr); cudaFree(d_out); cudaFree(d_scratch);
    cudaFree(d_band); cudaFree(d_status);
    cudaStreamDestroy(st);
    return ret;
}

int b200j2k_idwt(int device, int type, const int32_t *border, int levels, void *plane)
{
    if (device_ok() <= 0 || cudaSetDevice(device) != cudaSuccess)
        return B200J2K_ERR_DEVICE;
    if (type < 0 || type > 2 || levels < 0 || levels > J2K_MAX_LEVELS)
        return B200J2K_ERR_EINVAL;
    const size_t n = (size_t)(border[1] - border[0]) * (size_t)(border[3] - border[2]);
    if (!n || !levels)
        return 0;
    void *dA = nullptr, *dB = nullptr, *dC = nullptr;
    TileCompDesc *dtc = nullptr;
    int ret = 0;
    if (cudaMalloc(&dA, n * 4 + 256) != cudaSuccess || cudaMalloc(&dB, n * 4 + 256) != cudaSuccess ||
        cudaMalloc(&dC, n * 4 + 256) != cudaSuccess || cudaMalloc(&dtc, sizeof(TileCompDesc)) != cudaSuccess) {
        ret = B200J2K_ERR_NOMEM;
    } else {
        cudaMemcpy(dA, plane, n * 4, cudaMemcpyHostToDevice);
        if (launch_idwt_plane(nullptr, type, border, levels, dA, dB, dC, dtc) < 0)
            ret = B200J2K_ERR_DEVICE;
        cudaError_t e = cudaMemcpy(plane, dA, n * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            fprintf(stderr, "b200j2k_idwt: CUDA error: %s\n", cudaGetErrorString(e));
            ret = B200J2K_ERR_DEVICE;
        }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dtc);
    return ret;
}

void *b200j2k_device_alloc(int device, int64_t bytes)
{
    void *p = nullptr;
    if (device_ok() <= 0 || cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p, (size_t)bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void b200j2k_device_free(void *p) { if (p) cudaFree(p); }
int b200j2k_device_download(void *dst_host, const void *src_device, int64_t bytes)
{
    return cudaMemcpy(dst_host, src_device, (size_t)bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : B200J2K_ERR_DEVICE;
}

int b200j2k_pointer_device(const void *p)
{
    cudaPointerAttributes a;
    if (device_ok() <= 0 || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged ? a.device : -1;
}

void *b200j2k_host_alloc(int64_t bytes)
{
    void *p = nullptr;
    if (device_ok() <= 0 || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void b200j2k_host_free(void *p) { if (p) cudaFreeHost(p); }
int b200j2k_host_register(void *p, int64_t bytes)
{
    if (device_ok() <= 0)
        return B200J2K_ERR_DEVICE;
    return cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault) == cudaSuccess ? 0 : B200J2K_ERR_EINVAL;
}
int b200j2k_host_unregister(void *p) { return cudaHostUnregister(p) == cudaSuccess ? 0 : B200J2K_ERR_EINVAL; }

const char *b200j2k_pix_fmt_name(int pix) { return pix_fmt_name(pix); }
const char *b200j2k_version(void) { return "b200j2k 0.1 (sm_100a)"; }

/* ---- introspection of the host plan (CPU only; used by the not-gpu tests to check the parser) ---------- */
static thread_local Parser *g_plan = nullptr;

int b200j2k_plan_build(const uint8_t *pkt, int64_t pkt_size, int flags, b200j2k_frame_info *info)
{
    if (!g_plan)
        g_plan = new Parser();
    int ret = g_plan->parse(pkt, pkt_size, flags, false);
    if (ret < 0)
        return ret;
    fill_info(*g_plan, info);
    return 0;
}

/* which: 0 CblkDesc, 1 BandDesc, 2 TileCompDesc, 3 PackGroup, 4 gathered extra bytes */
const void *b200j2k_plan_table(int which, int64_t *count, int32_t *elem_size)
{
    if (!g_plan)
        return nullptr;
    switch (which) {
    case 0: *count = (int64_t)g_plan->cblk.size(); *elem_size = sizeof(CblkDesc); return g_plan->cblk.data();
    case 1: *count = (int64_t)g_plan->bands.size(); *elem_size = sizeof(BandDesc); return g_plan->bands.data();
    case 2: *count = (int64_t)g_plan->tcs.size(); *elem_size = sizeof(TileCompDesc); return g_plan->tcs.data();
    case 3: *count = (int64_t)g_plan->groups.size(); *elem_size = sizeof(PackGroup); return g_plan->groups.data();
    case 4: *count = (int64_t)g_plan->extra.size(); *elem_size = 1; return g_plan->extra.data();
    case 5: *count = 256; *elem_size = 4; return g_plan->hdr.palette;
    }
    return nullptr;
}

int64_t b200j2k_plan_scalar(int which)
{
    if (!g_plan)
        return -1;
    switch (which) {
    case 0: return (int64_t)g_plan->coef_elems;
    case 1: return (int64_t)g_plan->picture_bytes;
    case 2: return g_plan->max_levels;
    case 3: return g_plan->plane_off[1];
    case 4: return g_plan->plane_off[2];
    case 5: return g_plan->plane_off[3];
    }
    return -1;
}

}  // extern "C"
