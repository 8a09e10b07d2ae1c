# marginal cost of the kernels in the pipelined run: B200J2K_SKIP bits 1 un-stuff, 2 VLC, 4 MagSgn, 8 IDWT
for s in ${SKIPS:-0 1 2 4 8 7 11 14}; do echo -n "skip=$s: "; B200J2K_SKIP=$s python tools/pipelined_fps.py cfg4 12 384 2>&1 | tail -1; done
