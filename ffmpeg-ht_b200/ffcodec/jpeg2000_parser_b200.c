/*
 * jpeg2000_parser_b200.c -- the AVCodecParser "jpeg2000" (libavcodec/jpeg2000_parser.c:210-215) on top of the
 * C ABI of libb200j2k (include/b200j2k.h, b200j2k_splitter_*): same codec id, same parser_parse contract
 * (return value, *poutbuf / *poutbuf_size, PARSER_FLAG_COMPLETE_FRAMES), plus a parser_close that releases
 * the splitter.  A maintainer lists it in libavcodec/parsers.c in place of ff_jpeg2000_parser
 * (INTEGRATION.md).  Compiled against the reference's headers by oracle/Makefile for the tests.
 */
#include "libavcodec/avcodec.h"

#include "b200j2k.h"

typedef struct B200ParserContext {
    b200j2k_splitter *sp;
} B200ParserContext;

static int b200_parse(AVCodecParserContext *s, AVCodecContext *avctx,
                      const uint8_t **poutbuf, int *poutbuf_size,
                      const uint8_t *buf, int buf_size)
{
    B200ParserContext *c = s->priv_data;
    (void)avctx;
    if (!c->sp && b200j2k_splitter_create(&c->sp, s->flags & PARSER_FLAG_COMPLETE_FRAMES) < 0) {
        *poutbuf      = NULL;
        *poutbuf_size = 0;
        return buf_size;
    }
    return b200j2k_splitter_parse(c->sp, poutbuf, poutbuf_size, buf, buf_size);
}

static void b200_close(AVCodecParserContext *s)
{
    B200ParserContext *c = s->priv_data;
    b200j2k_splitter_destroy(c->sp);
    c->sp = NULL;
}

const AVCodecParser ff_jpeg2000_b200_parser = {
    .codec_ids      = { AV_CODEC_ID_JPEG2000 },
    .priv_data_size = sizeof(B200ParserContext),
    .parser_parse   = b200_parse,
    .parser_close   = b200_close,
};
