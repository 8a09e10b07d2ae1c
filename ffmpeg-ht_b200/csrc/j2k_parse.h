/*
 * j2k_parse.h -- host side of the path: codestream -> FramePlan.
 *
 * Keeps on the CPU exactly what BASELINE.json's north_star keeps there: marker segments
 * (libavcodec/jpeg2000dec.c:197-1014, :2425-2637), geometry (libavcodec/jpeg2000.c:214-577,
 * jpeg2000dec.c:1016-1070), tag trees and packet headers (jpeg2000dec.c:70-131, :1136-1869).
 * Instead of the reference's pointer tree with per-block allocations it emits flat tables
 * (j2k_plan.h) and references code-block bytes in place.  Geometry depends only on the main header,
 * so it is cached across the frames of a sequence.
 */
#ifndef B200J2K_PARSE_H
#define B200J2K_PARSE_H

#include <cstdint>
#include <utility>
#include <vector>

#include "j2k_plan.h"

namespace b200j2k {

struct CodSty {
    int nres = 0, l2cw = 0, l2ch = 0, cblk_style = 0, transform = 0, csty = 0, nlayers = 0, prog = 0, mct = 0;
    int init = 0;
    uint8_t ppx[J2K_MAX_LEVELS + 1], ppy[J2K_MAX_LEVELS + 1];
};

struct QntSty {
    uint8_t expn[100];
    uint16_t mant[100];
    int style = 0, nguard = 0;
};

struct Rect { int x0, x1, y0, y1; };

/* geometry of one precinct of one band, and of one resolution of one tile-component */
struct PrecGeom {
    int ncw, nch;           /* code-block grid */
    uint32_t cblk_base;     /* index of its first code-block in the frame tables */
    uint32_t tag_base;      /* index of its first tag-tree node (inclusion tree; zero-bit tree follows) */
    uint32_t tag_nodes;     /* nodes per tree */
};

struct BandGeom {
    Rect c;
    int empty;
    uint32_t prec_base;     /* index into FrameGeom::precs */
    int band_index;         /* index into bands[] */
};

struct ResGeom {
    Rect c;
    int ppx, ppy, npx, npy, nbands;
    BandGeom band[3];
};

struct TileGeom {
    Rect c;
    uint32_t res_base[J2K_MAX_COMP];   /* index into FrameGeom::res, nres entries per component */
    Rect comp[J2K_MAX_COMP];
};

/* per-code-block tier-2 state (what Jpeg2000Cblk carries between packets) */
struct CblkState {
    uint8_t incl, lblock, plhd, npasses;
    uint8_t zbp;
    uint8_t ncontrib;
    uint8_t nzb;             /* cblk->nonzerobits as computed at first inclusion (jpeg2000dec.c:1186-1194) */
    uint8_t modes;           /* cblk->modes: SPcod copy (:1177), bit 6 cleared once a MIXED block turns out to be Part-1 */
    uint32_t Lcup, Lref, length;
    uint32_t data_off;
};

struct TagNodeState {
    int16_t val;
    uint8_t vis;
};

struct Header {
    int W = 0, H = 0, X0 = 0, Y0 = 0, TW = 0, TH = 0, TX0 = 0, TY0 = 0, ncomp = 0, precision = 0, rsiz = 0;
    int cbps[J2K_MAX_COMP] = {0}, cdx[J2K_MAX_COMP] = {0}, cdy[J2K_MAX_COMP] = {0}, roi[J2K_MAX_COMP] = {0};
    int isHT = 0, cap_b05 = 0, cap_b11 = 0, cap_b12 = 0, cap_b14 = 0;
    CodSty cod[J2K_MAX_COMP];
    QntSty qnt[J2K_MAX_COMP];
    int had_coc[J2K_MAX_COMP] = {0}, had_qcc[J2K_MAX_COMP] = {0};
    int ntx = 0, nty = 0;
    int pix = 0, out_w = 0, out_h = 0;
    int lossless = 0;
    /* POC marker segments of the main header (get_poc, jpeg2000dec.c:760-818): progression order changes */
    struct Poc { int RS, CS, LYE, RE, CE, P; };
    int npoc = 0;
    Poc poc[32];
    /* JP2 wrapper (jp2_find_codestream, jpeg2000dec.c:2658-2804) and the pixel format it steers */
    int colour_space = 0, pal8 = 0;
    int sar_num = 0, sar_den = 0;     /* s->sar from the 'res ' box (jpeg2000dec.c:2762-2796) */
    uint32_t palette[256] = {0};      /* s->palette (jpeg2000dec.h:94), 0xAARRGGBB per entry; delivered as plane 1 of pal8 */
    int cdef[J2K_MAX_COMP] = { -1, -1, -1, -1 };
    const char *pix_name = "none";      /* libavutil name of the chosen AVPixelFormat */
    int chroma_lw = 0, chroma_lh = 0;   /* its log2_chroma_w / log2_chroma_h */
};

/* coding / quantisation / progression styles of one tile: the main header's at its first tile-part (get_sot,
 * jpeg2000dec.c:862-869), then whatever COD / COC / QCD / QCC / POC / RGN its tile-part headers carry */
struct TileSty {
    CodSty cod[J2K_MAX_COMP];
    QntSty qnt[J2K_MAX_COMP];
    int roi[J2K_MAX_COMP] = {0};            /* tile RGN; 0 = the main header's value (init_tile, :1052-1053) */
    uint8_t had_coc[J2K_MAX_COMP] = {0}, had_qcc[J2K_MAX_COMP] = {0};   /* tile->properties */
    int npoc = 0;
    bool poc_default = true;
    Header::Poc poc[32];
};

struct TilePart { const uint8_t *p, *end, *start; const uint8_t *hp, *hend; };   /* hp/hend: its run of PPM packet headers */

class Parser {
public:
    /* results of the last parse() */
    Header hdr;
    std::vector<CblkDesc> cblk;
    std::vector<BandDesc> bands;
    std::vector<uint8_t> band_negative_Mb_;   /* per entry of bands: exponent + guard bits - 1 < 0 */
    std::vector<TileCompDesc> tcs;
    std::vector<PackGroup> groups;
    std::vector<uint8_t> extra;       /* gathered multi-contribution segments, addressed at pkt_size + offset */
    uint64_t coef_elems = 0;          /* elements of coefA */
    uint64_t ll_elems = 0;            /* elements needed by each of coefB / coefC */
    int max_levels = 0;
    int nplanes = 0;
    int plane_w[4] = {0}, plane_h[4] = {0}, plane_row_bytes[4] = {0};
    uint64_t plane_off[4] = {0};      /* byte offset of each plane in the packed picture */
    uint64_t picture_bytes = 0;
    int64_t consumed = 0;
    uint32_t max_w = 0, max_h = 0;    /* largest code-block in the frame */
    uint32_t max_area = 0, max_state = 0;  /* largest w*h and (w+3)*(h+3) */
    uint32_t max_Lcup = 0, max_Lref = 0;
    int any_refine = 0;
    uint64_t geom_serial = 0;         /* bumped whenever the geometry tables were rebuilt */
    /* tile subset (B200J2K_FLAG_TILES): byte ranges of the packet the selected tiles' code-blocks live in -- empty =
     * the whole packet is needed */
    std::vector<std::pair<uint32_t, uint32_t>> needed_ranges;     /* (offset, length) */
    bool tile_subset = false;
    static bool tile_selected(int flags, int tileno);

    /* flags: B200J2K_FLAG_* ; headers_only: stop after the marker segments */
    int parse(const uint8_t *pkt, int64_t size, int flags, bool headers_only);

private:
    /* geometry cache, valid while the main header bytes do not change */
    std::vector<uint8_t> geom_key_;
    int geom_flags_ = -1;
    std::vector<TileGeom> tiles_;
    std::vector<ResGeom> res_;
    std::vector<PrecGeom> precs_;
    std::vector<CblkDesc> cblk_template_;
    uint32_t tag_total_ = 0;

    /* per-frame tier-2 state */
    std::vector<CblkState> state_;
    std::vector<TagNodeState> tags_;
    std::vector<uint8_t> decoded_layers_;   /* per precinct of band 0 */
    std::vector<std::vector<TilePart>> parts_;
    /* packed packet headers: PPM (main header, jpeg2000dec.c:958-980) and PPT (tile-part header, :982-1014) */
    std::vector<uint8_t> ppm_;
    bool has_ppm_ = false;
    std::vector<std::vector<uint8_t>> ppt_;   /* per tile */
    std::vector<uint8_t> has_ppt_;
    std::vector<TileSty> tsty_;
    std::vector<uint8_t> geom_extra_, geom_extra_key_;   /* tile-header segments the geometry depends on */
    int roi_of(int tileno, int c) const { return tsty_[tileno].roi[c] ? tsty_[tileno].roi[c] : hdr.roi[c]; }
    std::vector<uint8_t> tile_first_;          /* per tile: a tile-part with TPsot == 0 was seen */
    std::vector<size_t> ppt_pos_;             /* tile->packed_headers_stream as an offset into ppt_[tile] */
    std::vector<size_t> ppt_live_;            /* ... and its size: what had arrived at the SOD of the first tile-part */

    /* packet-header bit reader */
    const uint8_t *bp_ = nullptr, *bend_ = nullptr;
    int bit_index_ = 8;
    const uint8_t *pkt_ = nullptr;
    int64_t pkt_size_ = 0;

    int parse_headers(const uint8_t *pkt, int64_t size, int flags, size_t *main_header_len);
    int build_geometry(int flags);
    int decode_packets(int tileno);
    int po_iteration(int tileno, int RS, int CS, int LYE, int RE, int CE, int P, int *tp);
    const uint8_t *cur_part_start_ = nullptr;
    int decode_packet(int tileno, int comp, int r, int precno, int layno, int *tp);
    int precinct_at(int tileno, int c, int r, int x, int y, bool cprl_rule) const;
    void select_part(int tileno, int *tp, int csty);
    void select_header(int tileno, int *tp);
    inline int bits(int n);
    inline void flush();
    int tagtree_decode(uint32_t base, int ncw, int nch, int leaf, int threshold);
};

const char *pix_fmt_name(int pix);

}  // namespace b200j2k
#endif
