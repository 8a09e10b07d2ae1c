/*
 * j2k_splitter.cc -- the frame splitter in front of the decoder: cuts a byte stream (concatenated raw
 * codestreams and/or JP2 / JPH files, arriving in arbitrary pieces) into the packets the decode call takes.
 * Host code: the work is a serial scan that skips most bytes by the lengths it reads (marker segments by
 * Lxxx, tile-parts by Psot), so there is nothing for the GPU in it -- SURVEY.md section 8(f) row N3.
 *
 * Behaviour follows the reference's AVCodecParser "jpeg2000" bit for bit, quirks included:
 *   libavcodec/jpeg2000_parser.c:36-60   JPEG2000ParserContext / reset_context
 *   libavcodec/jpeg2000_parser.c:64-85   info_marker (which markers carry a length)
 *   libavcodec/jpeg2000_parser.c:91-175  find_frame_end (the state machine)
 *   libavcodec/jpeg2000_parser.c:177-208 jpeg2000_parse
 *   libavcodec/parser.c:203-288          ff_combine_frame (reassembly across calls, "overread" bytes when a
 *                                        frame end is only recognised some bytes into the next frame)
 * Checked against the compiled reference on chunked streams (tests/test_splitter.py).
 */
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "b200j2k.h"

namespace {

constexpr int kEndNotFound = -100;      /* END_NOT_FOUND, parser.h:40 */
constexpr int kPadding = 64;            /* AV_INPUT_BUFFER_PADDING_SIZE */

enum class Kind : uint8_t { None = 0, Jp2File = 1, Codestream = 2 };

/* A marker 0xFFxx is followed by a two-byte length unless it is SOC, SOT (handled apart), SOD, EPH or EOC. */
inline bool marker_has_length(uint32_t m)
{
    if (m < 0xFF00)
        return false;
    switch (m & 0xFF) {
    case 0x4F: case 0x90: case 0x92: case 0x93: case 0xD9:
        return false;
    default:
        return true;
    }
}

struct Splitter {
    /* reassembly (ParseContext) */
    std::vector<uint8_t> store;
    int fill = 0;              /* bytes of the unfinished frame held in `store` */
    int last_fill = 0;
    int overread = 0;          /* bytes of the NEXT frame already swallowed into `store` ... */
    int overread_at = 0;       /* ... and where they sit */
    uint64_t window = 0;       /* the last eight bytes seen, most significant first */
    bool started = false;      /* frame_start_found */
    /* scan state (JPEG2000ParserContext) */
    uint64_t seen = 0;         /* bytes since the last reset */
    uint32_t skip = 0;         /* bytes to pass over without looking */
    Kind kind = Kind::None;
    uint8_t sig_countdown = 0; /* bytes until the JP2 signature check */
    uint8_t sot_countdown = 0; /* bytes until Psot is complete in `window` */
    bool cs_done = false;      /* a JP2 file's codestream has ended */
    bool in_cs = false;
    bool whole_frames = false; /* PARSER_FLAG_COMPLETE_FRAMES */

    void rearm()
    {
        started = false;
        window = 0;
        seen = 0;
        kind = Kind::None;
        cs_done = false;
        sig_countdown = 0;
        skip = 0;
        sot_countdown = 0;
        in_cs = false;
    }

    /* position of the first byte of the next frame relative to buf (may be negative), or kEndNotFound */
    int scan(const uint8_t *buf, int n)
    {
        if (n == 0)
            return 0;
        uint64_t w = window, cnt = seen;
        for (int i = 0; i < n; i++) {
            w = w << 8 | buf[i];
            cnt++;
            if (skip) {
                if (skip > 8) {             /* long jump, keeping the last eight bytes flowing through `w` */
                    int64_t hop = (int64_t)skip - 8;
                    if (hop > n - i - 9)
                        hop = n - i - 9;
                    if (hop > INT_MAX)
                        hop = INT_MAX;
                    if (hop > 0) {
                        skip -= (uint32_t)hop;
                        i += (int)hop;
                        cnt += (uint64_t)hop;
                    }
                }
                skip--;
                continue;
            }
            if (sot_countdown) {
                if (sot_countdown == 1) {
                    /* Psot counts from the SOT marker: 9 bytes of it are behind us.  The reference compares
                     * an unsigned difference with 0, so Psot < 9 wraps instead of clamping; keep that. */
                    const uint32_t psot = (uint32_t)(w & 0xFFFFFFFFu);
                    skip = psot - 9u > 0u ? psot - 9u : 0u;
                }
                sot_countdown--;
                continue;
            }
            if (sig_countdown) {
                if (sig_countdown == 1 && w == 0x6A5020200D0A870AULL) {       /* "jP  \r\n\x87\n" */
                    if (started) {          /* a new file begins: the frame ended 12 bytes ago */
                        started = false;
                        rearm();
                        return i - 11;
                    }
                    started = true;
                    kind = Kind::Jp2File;
                }
                sig_countdown--;
            }
            const uint32_t lo32 = (uint32_t)(w & 0xFFFFFFFFu), lo16 = lo32 & 0xFFFFu;
            if (lo32 == 0x0000000Cu && cnt >= 3) {
                sig_countdown = 8;          /* length field of a signature box */
            } else if (lo16 == 0xFF4F) {    /* SOC */
                in_cs = true;
                if (!started) {
                    started = true;
                    kind = Kind::Codestream;
                } else if (kind == Kind::Jp2File && cs_done) {
                    rearm();
                    return i - 1;
                }
            } else if (lo16 == 0xFFD9) {    /* EOC */
                if (started && kind == Kind::Jp2File) {
                    cs_done = true;
                } else if (started && kind == Kind::Codestream) {
                    rearm();
                    return i + 1;
                }
                in_cs = false;
            } else if (in_cs) {
                if (lo16 == 0xFF90) {       /* SOT: Lsot(2) Isot(2) Psot(4) follow */
                    sot_countdown = 8;
                } else if (marker_has_length(lo32 >> 16) && started && lo16) {
                    skip = lo16 - 1;        /* to the last byte of this marker segment */
                    if ((int64_t)i + skip + 1 < n) {
                        const uint32_t nxt = (uint32_t)buf[i + skip] << 8 | buf[i + skip + 1];
                        if (marker_has_length(nxt))
                            skip += 2;      /* land behind the next marker code, on its length */
                    }
                }
            }
        }
        window = w;
        seen = cnt;
        return kEndNotFound;
    }

    /* ff_combine_frame: < 0 = no complete frame yet (or no memory), 0 = *buf/*size describe one frame */
    int combine(int next, const uint8_t **buf, int *size)
    {
        for (; overread > 0; overread--)
            store[fill++] = store[overread_at++];
        if (next > *size)
            return B200J2K_ERR_EINVAL;
        if (*size == 0 && next == kEndNotFound)
            next = 0;                       /* flush at end of input */
        last_fill = fill;
        if (next == kEndNotFound) {
            store.resize((size_t)fill + *size + kPadding);
            memcpy(store.data() + fill, *buf, (size_t)*size);
            memset(store.data() + fill + *size, 0, kPadding);
            fill += *size;
            return -1;
        }
        *size = overread_at = fill + next;
        if (fill) {
            if (store.size() < (size_t)fill + (next > 0 ? next : 0) + kPadding)
                store.resize((size_t)fill + (next > 0 ? next : 0) + kPadding);
            /* the reference copies next + padding bytes here, i.e. up to 64 bytes past the frame out of the
             * caller's buffer (libavcodec pads its packets); only the frame's own bytes are taken and the
             * padding behind them is zeroed, which is what a padded packet would have held */
            if (next > 0)
                memcpy(store.data() + fill, *buf, (size_t)next);
            memset(store.data() + fill + (next > 0 ? next : 0), 0, kPadding);
            fill = 0;
            *buf = store.data();
        }
        if (next < -8) {
            overread += -8 - next;
            next = -8;
        }
        for (; next < 0; next++) {
            window = window << 8 | store[(size_t)(last_fill + next)];
            overread++;
        }
        return 0;
    }
};

}  // namespace

extern "C" {

struct b200j2k_splitter {
    Splitter s;
};

int b200j2k_splitter_create(b200j2k_splitter **out, int complete_frames)
{
    if (!out)
        return B200J2K_ERR_EINVAL;
    b200j2k_splitter *p = new (std::nothrow) b200j2k_splitter();
    if (!p)
        return B200J2K_ERR_NOMEM;
    p->s.whole_frames = complete_frames != 0;
    *out = p;
    return 0;
}

void b200j2k_splitter_destroy(b200j2k_splitter *sp)
{
    delete sp;
}

int b200j2k_splitter_find_frame_end(b200j2k_splitter *sp, const uint8_t *buf, int size)
{
    if (!sp || size < 0 || (size && !buf))
        return B200J2K_ERR_EINVAL;
    return sp->s.scan(buf, size);
}

int b200j2k_splitter_parse(b200j2k_splitter *sp, const uint8_t **out, int *out_size, const uint8_t *buf, int size)
{
    if (!sp || !out || !out_size || size < 0 || (size && !buf))
        return B200J2K_ERR_EINVAL;
    Splitter &s = sp->s;
    int next;
    try {
        if (s.whole_frames) {
            next = size;
        } else {
            next = s.scan(buf, size);
            if (s.combine(next, &buf, &size) < 0) {
                *out = nullptr;
                *out_size = 0;
                return size;
            }
        }
    } catch (const std::bad_alloc &) {
        s.fill = 0;
        *out = nullptr;
        *out_size = 0;
        return size;
    }
    *out = buf;
    *out_size = size;
    return next;
}

}  // extern "C"
