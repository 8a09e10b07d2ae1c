// fuzz_host.cc -- robustness of the host side against damaged input: mutates seed codestreams (byte / bit flips,
// truncation, insertions, deletions, biased towards the headers) and runs the plan builder and the frame splitter
// on exact-size heap copies, so that AddressSanitizer sees any over-read; also checks that every code-block
// descriptor the plan hands to the device stays inside the packet.  No CUDA involved.
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined,float-cast-overflow -Iffmpeg-ht_b200/csrc -Iinclude \
//       -I/usr/local/cuda/include tools/fuzz_host.cc ffmpeg-ht_b200/csrc/j2k_parse.cc \
//       ffmpeg-ht_b200/csrc/j2k_splitter.cc -o /tmp/fuzz_host
//   /tmp/fuzz_host <seed> <iterations> seed0.j2k seed1.j2k ...      (tests/test_host_fuzz.py does both)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "j2k_parse.h"
#include "b200j2k.h"
using namespace b200j2k;
static uint64_t rs = 88172645463325252ull;
static uint32_t rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }
int main(int argc, char **argv)
{
    int iters = argc > 2 ? atoi(argv[2]) : 2000;
    std::vector<std::vector<uint8_t>> seeds;
    for (int a = 3; a < argc; a++) {
        FILE *f = fopen(argv[a], "rb"); if (!f) continue;
        std::vector<uint8_t> d; uint8_t buf[65536]; size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + n);
        fclose(f); seeds.push_back(d);
    }
    rs ^= (uint64_t)atoi(argv[1]) * 0x9E3779B97F4A7C15ull;
    Parser P;
    int ok = 0, bad = 0;
    for (int it = 0; it < iters; it++) {
        std::vector<uint8_t> d = seeds[rnd() % seeds.size()];
        int nm = 1 + rnd() % 6;
        for (int m = 0; m < nm; m++) {
            size_t hdr = d.size() < 400 ? d.size() : 400;
            size_t pos = (rnd() & 3) ? rnd() % hdr : rnd() % d.size();     // bias towards the headers
            switch (rnd() % 5) {
            case 0: d[pos] = (uint8_t)rnd(); break;
            case 1: d[pos] ^= 1u << (rnd() & 7); break;
            case 2: d.resize(pos + 1); break;
            case 3: d.insert(d.begin() + pos, (uint8_t)rnd()); break;
            case 4: if (pos + 2 < d.size()) d.erase(d.begin() + pos, d.begin() + pos + 1 + rnd() % 2); break;
            }
            if (d.empty()) d.push_back(0);
        }
        // exact-size heap copy so that ASAN sees any over-read (the library contract asks for no padding)
        uint8_t *p = (uint8_t *)malloc(d.size()); memcpy(p, d.data(), d.size());
        int flags = (rnd() & 1) | ((rnd() % 8 == 0 ? rnd() % 4 : 0) << 8);
        if (rnd() % 5 == 0) {                       // a context that owns every step-th tile (B200J2K_FLAG_TILES)
            const int step = 2 + rnd() % 4;
            flags |= B200J2K_FLAG_TILES(rnd() % step, step);
        }
        int r = P.parse(p, (int64_t)d.size(), flags, false);
        if (r >= 0) {
            ok++;
            // every descriptor must stay inside pkt + extra
            for (auto &c : P.cblk) {
                uint64_t lim = (uint64_t)d.size() + P.extra.size();
                if (c.length && (uint64_t)c.data_off + c.length > lim) { printf("OOB descriptor\n"); abort(); }
                if (!(c.flags & J2K_CBLK_MQ) && (uint32_t)c.Lcup + c.Lref > c.length && c.npasses) { printf("Lcup+Lref > length\n"); abort(); }
            }
        } else bad++;
        // splitter on the same bytes in random pieces
        b200j2k_splitter *sp; b200j2k_splitter_create(&sp, 0);
        size_t off = 0;
        while (off < d.size()) {
            size_t n = 1 + rnd() % 300; if (n > d.size() - off) n = d.size() - off;
            uint8_t *q = (uint8_t *)malloc(n); memcpy(q, p + off, n);
            size_t o2 = 0; int guard = 0;
            while (o2 < n && guard++ < 1000) {
                const uint8_t *out; int osz;
                int c = b200j2k_splitter_parse(sp, &out, &osz, q + o2, (int)(n - o2));
                if (osz > 0) { volatile uint8_t s = 0; for (int k = 0; k < osz; k++) s ^= out[k]; }
                if (c <= 0 && osz == 0) break;
                o2 += c > 0 ? c : 0;
            }
            free(q); off += n;
        }
        { const uint8_t *out; int osz; b200j2k_splitter_parse(sp, &out, &osz, nullptr, 0); }
        b200j2k_splitter_destroy(sp);
        free(p);
    }
    printf("ok %d rejected %d\n", ok, bad);
    return 0;
}
