"""ffmpeg-ht_b200 -- Python view (ctypes) of the B200-native HTJ2K decode path.

The product is the C-ABI shared library ``libb200j2k.so`` (include/b200j2k.h): host parser in C++,
hand-written sm_100a CUDA kernels, no CPU fallback.  This module is what tests and bench.py call; it
mirrors the reference's decoder interface for this path (``FFCodec ff_jpeg2000_decoder``:
``init`` -> :class:`Jpeg2000Decoder`, ``decode(avctx, frame, got_frame, pkt)`` -> :meth:`decode`,
errors as negative AVERROR codes -> :class:`DecodeError` carrying the same number).

Importable as ``ffmpeg_ht_b200`` through the shim module at the repo root (the directory name has a
hyphen).  Nothing here imports or calls anything under ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("B200J2K_LIB", _DIR / "libb200j2k.so"))   # env override: kernel tuning builds only

ERR_INVALIDDATA = -1094995529
ERR_PATCHWELCOME = -1163346256
ERR_NOMEM = -12
ERR_EINVAL = -22
ERR_DEVICE = -19
FLAG_BITEXACT = 1

PIX_NAMES = {1: "gray", 2: "gray16le", 3: "ya8", 4: "ya16le", 5: "rgb24", 6: "rgb48le", 7: "rgba",
             8: "rgba64le", 9: "planar8", 10: "planar16"}


class FrameInfo(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("pix_fmt", ctypes.c_int32),
                ("ncomponents", ctypes.c_int32), ("precision", ctypes.c_int32), ("lossless", ctypes.c_int32),
                ("profile", ctypes.c_int32), ("nplanes", ctypes.c_int32),
                ("plane_width", ctypes.c_int32 * 4), ("plane_height", ctypes.c_int32 * 4),
                ("plane_row_bytes", ctypes.c_int32 * 4), ("ntiles", ctypes.c_int32),
                ("ncodeblocks", ctypes.c_int32), ("bytes_consumed", ctypes.c_int64),
                ("av_pix_fmt_name", ctypes.c_char * 16), ("sar_num", ctypes.c_int32), ("sar_den", ctypes.c_int32),
                ("nblocks_failed", ctypes.c_int32), ("reserved", ctypes.c_int32)]

    @property
    def pix_name(self) -> str:
        """libavutil name of the AVPixelFormat the reference would choose ("rgb48le", "yuv420p10le", ...)"""
        return self.av_pix_fmt_name.decode()


class DecodeError(RuntimeError):
    def __init__(self, code: int, what: str = ""):
        super().__init__(f"b200j2k error {code} {what}".strip())
        self.code = code


def build(force: bool = False) -> Path:
    """Compile libb200j2k.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-s", "-C", str(_DIR), "clean"])
    subprocess.check_call(["make", "-s", "-C", str(_DIR), "-j8"])
    return LIB_PATH


_lib = None

_PP = ctypes.POINTER(ctypes.c_void_p)
_PI = ctypes.POINTER(ctypes.c_int)


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(f"{LIB_PATH} missing: run `make -C ffmpeg-ht_b200` (or __graft_entry__.build())")
        L = ctypes.CDLL(str(LIB_PATH))
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        info_p = ctypes.POINTER(FrameInfo)
        L.b200j2k_device_count.restype = i32
        L.b200j2k_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32]
        L.b200j2k_destroy.argtypes = [vp]
        L.b200j2k_destroy.restype = None
        L.b200j2k_probe.argtypes = [vp, i64, i32, info_p]
        L.b200j2k_submit.argtypes = [vp, vp, i64, _PP, _PI, info_p]
        L.b200j2k_wait.argtypes = [vp, i32]
        L.b200j2k_decode.argtypes = [vp, vp, i64, _PP, _PI, info_p]
        L.b200j2k_decode.restype = i64
        L.b200j2k_stage.argtypes = [vp, i32, vp, i64, info_p]
        L.b200j2k_run_staged.argtypes = [vp, i32]
        L.b200j2k_sync.argtypes = [vp, i32]
        L.b200j2k_run_staged_timed.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_float)]
        L.b200j2k_fetch.argtypes = [vp, i32, _PP, _PI]
        L.b200j2k_slot_stream.argtypes = [vp, i32]
        L.b200j2k_slot_stream.restype = vp
        L.b200j2k_slot_output.argtypes = [vp, i32, ctypes.POINTER(i64)]
        L.b200j2k_slot_output.restype = vp
        L.b200j2k_slot_launches.argtypes = [vp, i32]
        L.b200j2k_slot_blocks_failed.argtypes = [vp, i32]
        L.b200j2k_debug_poke.argtypes = [vp, i32]
        L.b200j2k_slot_h2d_bytes.argtypes = [vp, i32]
        L.b200j2k_slot_h2d_bytes.restype = i64
        L.b200j2k_ht_blocks.argtypes = [i32, vp, i64, vp, i32, vp, vp]
        L.b200j2k_idwt.argtypes = [i32, i32, vp, i32, vp]
        L.b200j2k_host_alloc.argtypes = [i64]
        L.b200j2k_host_alloc.restype = vp
        L.b200j2k_device_alloc.argtypes = [i32, i64]
        L.b200j2k_device_alloc.restype = vp
        L.b200j2k_device_free.argtypes = [vp]
        L.b200j2k_device_free.restype = None
        L.b200j2k_device_download.argtypes = [vp, vp, i64]
        L.b200j2k_host_free.argtypes = [vp]
        L.b200j2k_host_free.restype = None
        L.b200j2k_host_register.argtypes = [vp, i64]
        L.b200j2k_host_unregister.argtypes = [vp]
        L.b200j2k_pix_fmt_name.argtypes = [i32]
        L.b200j2k_pix_fmt_name.restype = ctypes.c_char_p
        L.b200j2k_version.restype = ctypes.c_char_p
        L.b200j2k_plan_build.argtypes = [vp, i64, i32, info_p]
        L.b200j2k_plan_table.argtypes = [i32, ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_int32)]
        L.b200j2k_plan_table.restype = vp
        L.b200j2k_plan_scalar.argtypes = [i32]
        L.b200j2k_plan_scalar.restype = i64
        L.b200j2k_splitter_create.argtypes = [ctypes.POINTER(vp), i32]
        L.b200j2k_splitter_destroy.argtypes = [vp]
        L.b200j2k_splitter_destroy.restype = None
        L.b200j2k_splitter_find_frame_end.argtypes = [vp, vp, i32]
        L.b200j2k_splitter_parse.argtypes = [vp, ctypes.POINTER(vp), _PI, vp, i32]
        _lib = L
    return _lib


EXPORTS = ["b200j2k_device_count", "b200j2k_create", "b200j2k_destroy", "b200j2k_probe", "b200j2k_submit",
           "b200j2k_wait", "b200j2k_decode", "b200j2k_stage", "b200j2k_run_staged", "b200j2k_run_staged_timed",
           "b200j2k_sync",
           "b200j2k_fetch", "b200j2k_slot_stream", "b200j2k_slot_output", "b200j2k_slot_launches", "b200j2k_slot_blocks_failed", "b200j2k_pointer_device", "b200j2k_debug_poke", "b200j2k_wait_upload", "b200j2k_pointer_is_pinned",
           "b200j2k_slot_h2d_bytes",
           "b200j2k_ht_blocks", "b200j2k_idwt", "b200j2k_host_alloc", "b200j2k_host_free",
           "b200j2k_host_register", "b200j2k_host_unregister", "b200j2k_device_alloc", "b200j2k_device_free",
           "b200j2k_device_download", "b200j2k_pix_fmt_name", "b200j2k_version",
           "b200j2k_plan_build", "b200j2k_plan_table", "b200j2k_plan_scalar",
           "b200j2k_splitter_create", "b200j2k_splitter_destroy", "b200j2k_splitter_find_frame_end",
           "b200j2k_splitter_parse"]


def _check(ret: int, what: str = "") -> int:
    if ret < 0:
        raise DecodeError(int(ret), what)
    return int(ret)


def _buf(b):
    """bytes / bytearray / uint8 ndarray -> (address, length, keep-alive)"""
    if isinstance(b, PinnedPacket):
        return b._ptr, b.size, b
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b, dtype=np.uint8)
        return a.ctypes.data, a.size, a
    a = np.frombuffer(b, dtype=np.uint8)
    return a.ctypes.data, a.size, a


def flag_lowres(n: int) -> int:
    """B200J2K_FLAG_LOWRES(n): AVCodecContext.lowres, drop the n highest resolution levels."""
    return (int(n) & 15) << 8


def probe(pkt, flags: int = 0) -> FrameInfo:
    """Header-only pass: frame geometry and pixel format (no device needed)."""
    addr, n, keep = _buf(pkt)
    info = FrameInfo()
    _check(lib().b200j2k_probe(addr, n, flags, ctypes.byref(info)), "probe")
    return info


class PinnedFrame:
    """A pinned host picture (what an AVFrame from a pinned get_buffer2 pool would be)."""

    def __init__(self, info: FrameInfo):
        self.info = info
        sizes = [info.plane_row_bytes[i] * info.plane_height[i] for i in range(info.nplanes)]
        self.nbytes = sum(sizes)
        self._ptr = lib().b200j2k_host_alloc(max(self.nbytes, 1))
        if not self._ptr:
            raise DecodeError(ERR_DEVICE, "host_alloc")
        self.buffer = np.ctypeslib.as_array((ctypes.c_uint8 * self.nbytes).from_address(self._ptr))
        self.data = (ctypes.c_void_p * 4)()
        self.linesize = (ctypes.c_int * 4)()
        off = 0
        self.planes = []
        for i in range(info.nplanes):
            self.data[i] = self._ptr + off
            self.linesize[i] = info.plane_row_bytes[i]
            self.planes.append(self.buffer[off: off + sizes[i]])
            off += sizes[i]

    def close(self):
        if self._ptr:
            lib().b200j2k_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PageableFrame:
    """A picture in ordinary (pageable) host memory -- what an AVFrame from libavcodec's default get_buffer2 is.
    The library copies device -> pinned staging -> this memory (b200j2k_submit's note on unregistered planes)."""

    def __init__(self, info: FrameInfo):
        self.info = info
        sizes = [info.plane_row_bytes[i] * info.plane_height[i] for i in range(info.nplanes)]
        self.nbytes = sum(sizes)
        self.buffer = np.zeros(max(self.nbytes, 1), np.uint8)
        self.data = (ctypes.c_void_p * 4)()
        self.linesize = (ctypes.c_int * 4)()
        off = 0
        for i in range(info.nplanes):
            self.data[i] = self.buffer.ctypes.data + off
            self.linesize[i] = info.plane_row_bytes[i]
            off += sizes[i]

    def close(self):
        pass


class DeviceFrame:
    """A picture in DEVICE memory (what an AV_PIX_FMT_CUDA frame would be): submit() writes it with a
    device-to-device copy, the pixels never cross PCIe.  download() fetches them for checking."""

    def __init__(self, info: FrameInfo, device: int = 0, pitch_align: int = 1):
        self.info = info
        self.linesize = (ctypes.c_int * 4)()
        self.data = (ctypes.c_void_p * 4)()
        sizes = []
        for i in range(info.nplanes):
            ls = -(-info.plane_row_bytes[i] // pitch_align) * pitch_align
            self.linesize[i] = ls
            sizes.append(ls * info.plane_height[i])
        self.nbytes = sum(sizes)
        self._ptr = lib().b200j2k_device_alloc(device, max(self.nbytes, 1))
        if not self._ptr:
            raise DecodeError(ERR_DEVICE, "device_alloc")
        off = 0
        self._off = []
        for i in range(info.nplanes):
            self.data[i] = self._ptr + off
            self._off.append(off)
            off += sizes[i]

    def download(self) -> np.ndarray:
        """-> packed picture bytes (row padding removed), same layout Jpeg2000Decoder.decode() returns"""
        raw = np.empty(max(self.nbytes, 1), np.uint8)
        _check(lib().b200j2k_device_download(raw.ctypes.data, self._ptr, self.nbytes), "device_download")
        out = []
        for i in range(self.info.nplanes):
            ls, rb, h = self.linesize[i], self.info.plane_row_bytes[i], self.info.plane_height[i]
            out.append(raw[self._off[i]: self._off[i] + ls * h].reshape(h, ls)[:, :rb].reshape(-1))
        return np.concatenate(out)

    def close(self):
        if self._ptr:
            lib().b200j2k_device_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedPacket:
    """A codestream held in pinned host memory: b200j2k_submit DMAs it without a staging copy."""

    def __init__(self, data):
        addr, n, keep = _buf(data)
        self.size = n
        self._ptr = lib().b200j2k_host_alloc(max(n, 1))
        if not self._ptr:
            raise DecodeError(ERR_DEVICE, "host_alloc")
        ctypes.memmove(self._ptr, addr, n)
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(self._ptr))

    def close(self):
        if self._ptr:
            lib().b200j2k_host_free(self._ptr)
            self._ptr = None

    def __len__(self):
        return self.size

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def frame_to_arrays(info: FrameInfo, planes):
    """Packed bytes -> ndarray(s) in the natural shape of the pixel format."""
    dt = np.uint16 if info.precision > 8 else np.uint8
    if info.nplanes == 1:
        a = np.frombuffer(planes[0], dtype=dt) if not isinstance(planes[0], np.ndarray) else planes[0].view(dt)
        return a.reshape(info.height, info.width, info.ncomponents)
    return [p.view(dt).reshape(info.plane_height[i], info.plane_width[i]) for i, p in enumerate(planes)]


class Jpeg2000Decoder:
    """One decoder context (= one AVCodecContext of the reference) on one CUDA device."""

    def __init__(self, device: int = 0, slots: int = 1, bitexact: bool = False, lowres: int = 0, flags: int = 0):
        """flags: further B200J2K_FLAG_* bits, e.g. sharding.tile_flags(rank, world) for a tile subset"""
        self._h = ctypes.c_void_p()
        self.flags = (FLAG_BITEXACT if bitexact else 0) | flag_lowres(lowres) | flags
        _check(lib().b200j2k_create(ctypes.byref(self._h), device, slots, self.flags), "create")
        self.device = device
        self.slots = slots
        self._keep = {}

    def close(self):
        if self._h:
            lib().b200j2k_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the drop-in call: packet in, picture out -------------------------------------------------
    def decode(self, pkt, out: np.ndarray | None = None):
        """decode(pkt) -> (FrameInfo, packed picture bytes as uint8 ndarray).  Pageable output."""
        info = probe(pkt, self.flags)
        sizes = [info.plane_row_bytes[i] * info.plane_height[i] for i in range(info.nplanes)]
        total = sum(sizes)
        if out is None:
            out = np.empty(total, np.uint8)
        data = (ctypes.c_void_p * 4)()
        linesize = (ctypes.c_int * 4)()
        off = 0
        for i in range(info.nplanes):
            data[i] = out.ctypes.data + off
            linesize[i] = info.plane_row_bytes[i]
            off += sizes[i]
        addr, n, keep = _buf(pkt)
        info2 = FrameInfo()
        _check(lib().b200j2k_decode(self._h, addr, n, data, linesize, ctypes.byref(info2)), "decode")
        return info2, out

    # -- pipelined form -------------------------------------------------------------------------------
    def submit(self, pkt, frame: PinnedFrame) -> int:
        addr, n, keep = _buf(pkt)
        info = FrameInfo()
        slot = _check(lib().b200j2k_submit(self._h, addr, n, frame.data, frame.linesize, ctypes.byref(info)), "submit")
        self._keep[slot] = (keep, frame)
        return slot

    def wait(self, slot: int):
        _check(lib().b200j2k_wait(self._h, slot), "wait")
        self._keep.pop(slot, None)

    # -- device-resident form (kernel-only timing, tests) -------------------------------------------
    def stage(self, slot: int, pkt) -> FrameInfo:
        addr, n, keep = _buf(pkt)
        info = FrameInfo()
        _check(lib().b200j2k_stage(self._h, slot, addr, n, ctypes.byref(info)), "stage")
        return info

    def run_staged(self, slot: int):
        _check(lib().b200j2k_run_staged(self._h, slot), "run_staged")

    def run_staged_timed(self, slot: int):
        """-> ms of [HT un-stuff, HT VLC, HT MagSgn + dequant, HT refinement blocks, IDWT levels, fused last level]"""
        ms = (ctypes.c_float * 6)()
        _check(lib().b200j2k_run_staged_timed(self._h, slot, ms), "run_staged_timed")
        return [float(v) for v in ms]

    def sync(self, slot: int):
        _check(lib().b200j2k_sync(self._h, slot), "sync")

    def fetch(self, slot: int, info: FrameInfo) -> np.ndarray:
        sizes = [info.plane_row_bytes[i] * info.plane_height[i] for i in range(info.nplanes)]
        out = np.empty(sum(sizes), np.uint8)
        data = (ctypes.c_void_p * 4)()
        linesize = (ctypes.c_int * 4)()
        off = 0
        for i in range(info.nplanes):
            data[i] = out.ctypes.data + off
            linesize[i] = info.plane_row_bytes[i]
            off += sizes[i]
        _check(lib().b200j2k_fetch(self._h, slot, data, linesize), "fetch")
        return out

    def stream(self, slot: int) -> int:
        return int(lib().b200j2k_slot_stream(self._h, slot) or 0)

    def h2d_bytes(self, slot: int) -> int:
        return int(lib().b200j2k_slot_h2d_bytes(self._h, slot))

    def launches(self, slot: int) -> int:
        return int(lib().b200j2k_slot_launches(self._h, slot))

    def blocks_failed(self, slot: int) -> int:
        """After wait(): code-blocks of the slot's frame the kernels zero-filled because their data was damaged."""
        return int(lib().b200j2k_slot_blocks_failed(self._h, slot))


# -- stage-level entry points (parity tests) ---------------------------------------------------------
class FrameSplitter:
    """Mirror of the reference's AVCodecParser "jpeg2000" (libavcodec/jpeg2000_parser.c): feed a byte stream in
    arbitrary pieces, get the codestreams / JP2 files back one by one."""

    def __init__(self, complete_frames: bool = False):
        self._h = ctypes.c_void_p()
        _check(lib().b200j2k_splitter_create(ctypes.byref(self._h), int(complete_frames)), "splitter_create")

    def close(self):
        if self._h:
            lib().b200j2k_splitter_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def parse(self, data):
        """One parser_parse call on `data` (bytes-like).  -> (return value, frame bytes or None)."""
        a = np.frombuffer(data, np.uint8) if len(data) else np.zeros(0, np.uint8)
        out, n = ctypes.c_void_p(), ctypes.c_int(0)
        r = lib().b200j2k_splitter_parse(self._h, ctypes.byref(out), ctypes.byref(n), a.ctypes.data if a.size else None,
                                         int(a.size))
        frame = ctypes.string_at(out.value, n.value) if out.value and n.value > 0 else None
        return int(r), frame

    def feed(self, data, flush: bool = False):
        """av_parser_parse2's calling loop: consume all of `data` (and, with flush, the end-of-input call);
        returns the list of completed frames."""
        frames = []
        mv = memoryview(bytes(data))
        pos = 0
        while pos < len(mv):
            r, f = self.parse(mv[pos:])
            if f is not None:
                frames.append(f)
            pos += max(r, 0)
            if r <= 0 and f is None:
                break
        if flush:
            r, f = self.parse(b"")
            if f is not None:
                frames.append(f)
        return frames


def ht_blocks(blocks, device: int = 0):
    """blocks: list of dicts(data, npasses, Lcup, Lref, zbp, w, h, M_b, roi, vsc) ->
    (list of int32 arrays h x w in the reference's sign-magnitude form, status array)."""
    n = len(blocks)
    par = np.zeros((n, 10), np.int32)
    chunks, off = [], 0
    for i, b in enumerate(blocks):
        d = bytes(b["data"])
        par[i] = [b["npasses"], b["Lcup"], b["Lref"], b["zbp"], b["w"], b["h"], b["M_b"], b.get("roi", 0),
                  b.get("vsc", 0), off]
        chunks.append(d)
        pad = (-len(d)) % 16 + 16
        chunks.append(b"\0" * pad)
        off += len(d) + pad
    data = np.frombuffer(b"".join(chunks), np.uint8)
    total = int(sum(int(b["w"]) * int(b["h"]) for b in blocks))
    out = np.zeros(max(total, 1), np.int32)
    status = np.zeros(max(n, 1), np.int32)
    _check(lib().b200j2k_ht_blocks(device, data.ctypes.data, data.size, par.ctypes.data, n, out.ctypes.data,
                                   status.ctypes.data), "ht_blocks")
    res, pos = [], 0
    for b in blocks:
        k = int(b["w"]) * int(b["h"])
        res.append(out[pos: pos + k].reshape(int(b["h"]), int(b["w"])))
        pos += k
    return res, status[:n]


def idwt(dwt_type: int, border, levels: int, plane: np.ndarray, device: int = 0) -> np.ndarray:
    """In-place inverse DWT of a Mallat-layout plane (float32 for type 0, int32 otherwise)."""
    border = np.ascontiguousarray(border, np.int32)
    assert plane.flags.c_contiguous and plane.dtype.itemsize == 4
    _check(lib().b200j2k_idwt(device, dwt_type, border.ctypes.data, levels, plane.ctypes.data), "idwt")
    return plane


# -- host plan introspection (CPU only) -------------------------------------------------------------------
CBLK_DTYPE = np.dtype([("data_off", "<u4"), ("dst_off", "<u4"), ("dst_stride", "<u4"), ("Lcup", "<u2"),
                       ("Lref", "<u2"), ("w", "<u2"), ("h", "<u2"), ("npasses", "u1"), ("zbp", "u1"), ("M_b", "u1"),
                       ("flags", "u1"), ("roi_shift", "u1"), ("pad", "u1"), ("band", "<u2"), ("length", "<u4")])
BAND_DTYPE = np.dtype([("fscale", "<f4"), ("iscale", "<i4"), ("i_stepsize", "<i4"), ("pad", "<i4")])
LEVEL_DTYPE = np.dtype([("u0", "<i4"), ("u1", "<i4"), ("v0", "<i4"), ("v1", "<i4")])
TC_DTYPE = np.dtype([("coef_off", "<u4"), ("width", "<i4"), ("height", "<i4"), ("x0", "<i4"), ("y0", "<i4"),
                     ("nlevels", "<i4"), ("transform", "<i4"), ("comp", "<i4"), ("out_x", "<i4"), ("out_y", "<i4"),
                     ("cbps", "<i4"), ("group", "<i4"), ("lev", LEVEL_DTYPE, (33,))])
GROUP_DTYPE = np.dtype([("ncomp", "<i4"), ("tc", "<i4", (4,)), ("mct", "<i4"), ("plane", "<i4"),
                        ("pixel_stride", "<i4"), ("sample_off", "<i4", (4,)), ("shift", "<i4", (4,)), ("wide", "<i4"),
                        ("comp_plane", "<i4", (4,))])


def plan(pkt, flags: int = 0):
    """Host parser only: -> dict(info, cblk, bands, tcs, groups, extra, coef_elems, picture_bytes, ...)."""
    addr, n, keep = _buf(pkt)
    info = FrameInfo()
    _check(lib().b200j2k_plan_build(addr, n, flags, ctypes.byref(info)), "plan_build")
    out = {"info": info}
    for which, name, dt in ((0, "cblk", CBLK_DTYPE), (1, "bands", BAND_DTYPE), (2, "tcs", TC_DTYPE),
                            (3, "groups", GROUP_DTYPE), (4, "extra", np.dtype("u1")), (5, "palette", np.dtype("<u4"))):
        cnt, es = ctypes.c_int64(), ctypes.c_int32()
        p = lib().b200j2k_plan_table(which, ctypes.byref(cnt), ctypes.byref(es))
        assert es.value == dt.itemsize, (name, es.value, dt.itemsize)
        if cnt.value:
            raw = (ctypes.c_uint8 * (cnt.value * es.value)).from_address(p)
            out[name] = np.frombuffer(bytes(raw), dtype=dt).copy()
        else:
            out[name] = np.zeros(0, dt)
    out["coef_elems"] = lib().b200j2k_plan_scalar(0)
    out["picture_bytes"] = lib().b200j2k_plan_scalar(1)
    out["max_levels"] = lib().b200j2k_plan_scalar(2)
    out["plane_off"] = [0, lib().b200j2k_plan_scalar(3), lib().b200j2k_plan_scalar(4), lib().b200j2k_plan_scalar(5)]
    return out
