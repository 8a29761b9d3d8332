"""Isolated timing of the DMMA tile GEMM and the POTRF tile kernel through the development-only symbols gpx_debug_gemm /
gpx_debug_potrf (exported only by `make -C thesis_code_b200/csrc DEBUG_API=1`)."""
import ctypes, sys
sys.path.insert(0, '.')
from thesis_code_b200 import _gpx
h = _gpx.Handle(0)
f = h.lib.gpx_debug_gemm
f.restype = ctypes.c_double
f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
g = h.lib.gpx_debug_potrf; g.restype = ctypes.c_double; g.argtypes = [ctypes.c_void_p, ctypes.c_int]
print("potrf_tile_kernel (128x128 Cholesky + inverse): %.1f us per launch" % g(h._h, 50))
print("lib:", _gpx.LIB_PATH)
shapes = [(148, 64, 1, 0, 16), (148, 64, 2, 0, 16), (148, 64, 4, 0, 16), (148, 64, 8, 0, 16), (148, 64, 16, 0, 16),
          (148, 64, 4, 1, 16), (148, 64, 1, 1, 16), (148, 64, 4, 0, 1), (148, 64, 4, 0, 4), (148, 64, 4, 0, 64),
          (148, 8, 4, 0, 8), (148, 1, 4, 0, 1), (148, 1, 16, 0, 1), (148, 1, 64, 0, 1), (296, 1, 64, 0, 1), (131, 300, 16, 0, 16)]
print("rows cols K overwrite chunk :  strip 16: ms TFLOP/s | strip 8: ms TFLOP/s")
for (rows, cols, K, ow, ch) in shapes:
    out = []
    for strip in (16, 8):
        h.set_option("gemm_strip", strip)
        ms = min(f(h._h, rows, cols, K, ow, ch, 5) for _ in range(2))
        fl = 2.0 * 128 ** 3 * rows * cols * K
        out.append((ms, fl / ms / 1e9))
    print("%4d %4d %3d %d %3d : %8.3f  %6.2f | %8.3f  %6.2f" % (rows, cols, K, ow, ch, out[0][0], out[0][1], out[1][0], out[1][1]))
