import os, sys, json
sys.path[:0] = ['/root/repo', '/root/repo/tools']
os.environ['B200SHIM_TRACE'] = '1'; os.environ['REF_LOG'] = '16'
import bench, ctypes, numpy as np
frames = bench.make_frames(4)
lib = ctypes.CDLL('/root/repo/oracle/_ref/libj2kb200shim.so')
fn = lib.b200shim_decode_bench; fn.restype = ctypes.c_double
fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
n = 16
bufs = [np.frombuffer(frames[i % 4], np.uint8) for i in range(n)]
ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs]); lens = (ctypes.c_int64 * n)(*[b.size for b in bufs])
print("cores", os.cpu_count(), flush=True)
for pinned in ('0', '1'):
    for depth in ('1', '2'):
        for t in (1, 2):
            os.environ['REF_PINNED_PACKETS'] = pinned; os.environ['REF_OPT_async_depth'] = depth
            s = fn(ptrs, lens, n, t, 4, 0)
            print(f"pinned={pinned} depth={depth} ctx={t}: {n*4/s:.0f} fps", flush=True)
