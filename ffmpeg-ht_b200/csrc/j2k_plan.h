/*
 * j2k_plan.h -- the flat tables the host parser hands to the kernels (plain structs shared by host C++
 * and CUDA device code).  This is the "inner C ABI" of SURVEY.md section 8(b): one FramePlan per packet.
 *
 * Data layout in HBM for one frame (one slot):
 *   bytes[]   the codestream as received (code-block segments are referenced in place by offset; blocks
 *             with several contributions are gathered behind it by the host)
 *   cblk[]    CblkDesc, one per code-block of every tile-component (also the empty ones: the kernel writes
 *             their zeros, so the coefficient planes never need a memset)
 *   coefA     per tile-component a Mallat-layout int32/float plane, row stride = component width
 *             (what Jpeg2000Component.i_data / f_data is in the reference, jpeg2000.c:499-511)
 *   coefB/C   ping-pong buffers holding the reconstructed LL of the previous DWT level
 *   picture   the packed output picture (AVFrame layout without padding)
 */
#ifndef B200J2K_PLAN_H
#define B200J2K_PLAN_H

#include <stdint.h>

#define J2K_MAX_COMP 4
#define J2K_MAX_LEVELS 32

enum { J2K_DWT97 = 0, J2K_DWT53 = 1, J2K_DWT97_INT = 2 };   /* enum DWTType, jpeg2000dwt.h:37-41 */

#define J2K_CBLK_SKIP 0x80
#define J2K_CBLK_MQ   0x08  /* Part-1 (EBCOT / MQ) code-block, style 0: decoded by mq_block_kernel (decode_cblk,
                             * jpeg2000dec.c:1993-2089).  Its descriptor carries nonzerobits in Lcup and the band
                             * orientation (0 LL, 1 HL, 2 LH, 3 HH) in Lref; data = the single codeword segment */

#define J2K_CBLK_MQ_RESET  0x10  /* Part-1 block: JPEG2000_CBLK_RESET, context states re-initialised after every pass */
#define J2K_CBLK_MQ_SEGSYM 0x20  /* Part-1 block: JPEG2000_CBLK_SEGSYM, a segmentation symbol behind every clean-up pass
                                  * (flags bit 0, stripe-causal, means vertically causal contexts for a Part-1 block) */

/* one HT code-block: a row of the table H9 (Jpeg2000Cblk, jpeg2000.h:183-205) */
typedef struct CblkDesc {
    uint32_t data_off;      /* offset of Dcup in bytes[]; Dref follows at data_off + Lcup */
    uint32_t dst_off;       /* element index of the block's top-left sample in coefA */
    uint32_t dst_stride;    /* row stride of that plane, elements */
    uint16_t Lcup, Lref;    /* pass_lengths[0], pass_lengths[1] */
    uint16_t w, h;
    uint8_t  npasses;
    uint8_t  zbp;
    uint8_t  M_b;           /* expn + guard bits - 1 (jpeg2000dec.c:2238) */
    uint8_t  flags;         /* bit0: stripe-causal (JPEG2000_CBLK_VSC)  bits 1-2: dequant mode (DWT type)
                             * bit7: J2K_CBLK_SKIP, resolution dropped by lowres: parsed, never decoded */
    uint8_t  roi_shift;
    uint8_t  pad;
    uint16_t band;          /* index into BandDesc[] */
    uint32_t length;        /* cblk->length: where the decoder plants its 0xFF sentinel (jpeg2000htdec.c:1258) */
} CblkDesc;

/* per sub-band dequantisation constants, computed on the host exactly as the reference does */
typedef struct BandDesc {
    float   fscale;         /* 9/7 float: f_stepsize / (1 << (31 - M_b))         (jpeg2000dec.c:2104-2106) */
    int32_t iscale;         /* 9/7 int: (int)(fscale * 2^6 * 2^24 + 0.5)         (:2164-2168) */
    int32_t i_stepsize;     /* 5/3: floorf(f_stepsize * 2^15), 32768 = no scaling (jpeg2000.c:271) */
    int32_t pad;
} BandDesc;

/* one DWT level of one tile-component (DWTContext.linelen/mod, jpeg2000dwt.c:539-560) */
typedef struct LevelDesc {
    int32_t u0, u1, v0, v1; /* absolute bounds of the resolution this level reconstructs */
} LevelDesc;

/* one tile-component = one independent inverse-DWT domain */
typedef struct TileCompDesc {
    uint32_t coef_off;      /* element offset of its plane in coefA (same offset in coefB / coefC) */
    int32_t  width, height; /* plane size = full-resolution size of the tile-component */
    int32_t  x0, y0;        /* absolute coordinates of its top-left sample (comp->coord) */
    int32_t  nlevels;       /* decomposition levels to undo */
    int32_t  transform;     /* J2K_DWT* */
    int32_t  comp;          /* component index */
    int32_t  out_x, out_y;  /* where its top-left sample lands in the component's output plane */
    int32_t  cbps;
    int32_t  group;         /* index of the MCT/pack group this plane belongs to */
    LevelDesc lev[J2K_MAX_LEVELS + 1];  /* lev[k] = bounds after k levels were undone; lev[nlevels] = full size */
} TileCompDesc;

/* output side: one entry per (tile, set of components written together) */
typedef struct PackGroup {
    int32_t ncomp;          /* 1..4 tile-components written by one kernel instance */
    int32_t tc[J2K_MAX_COMP];      /* TileCompDesc indices */
    int32_t mct;            /* apply RCT/ICT to tc[0..2] */
    int32_t plane;          /* output plane */
    int32_t pixel_stride;   /* samples between horizontally adjacent pixels of one component (ncomp if packed) */
    int32_t sample_off[J2K_MAX_COMP]; /* sample offset of each component inside a pixel */
    int32_t shift[J2K_MAX_COMP];      /* precision - cbps */
    int32_t wide;           /* 16-bit samples */
    int32_t comp_plane[J2K_MAX_COMP]; /* output plane of each component (planar formats; packed: all = plane) */
} PackGroup;

#endif
