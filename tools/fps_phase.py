"""Development aid: per-phase clock stamps inside the FPS kernel (CTA 0, warps 0/1, iterations 64..79)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stratified_transformer_b200 import pointops, _cabi
lib = _cabi.load()
lib.stb200_fps_debug_buffer.argtypes = [ctypes.c_void_p]
lib.stb200_fps_debug_buffer.restype = None
b, n = int(os.environ.get("FPS_PHASE_SCENES", "8")), int(sys.argv[1]) if len(sys.argv) > 1 else 80000
first = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lib.stb200_fps_debug_window.argtypes = [ctypes.c_int]
lib.stb200_fps_debug_window.restype = None
lib.stb200_fps_debug_window(first)
xyz = torch.rand(b * n, 3, device="cuda") * 6
off = (torch.arange(1, b + 1, device="cuda") * n).int()
new_off = (torch.arange(1, b + 1, device="cuda") * (n // 8 + 1)).int()
pointops.furthestsampling(xyz, off, new_off)
dbg = torch.zeros(16 * 2 * 8, dtype=torch.int64, device="cuda")
lib.stb200_fps_debug_buffer(dbg.data_ptr())
pointops.furthestsampling(xyz, off, new_off)
torch.cuda.synchronize()
lib.stb200_fps_debug_buffer(None)
d = dbg.cpu().numpy().reshape(16, 2, 8)
names = ["top", "computed", "sent", "inbox_done", "end", "-", "-", "-"]
for w in (0, 1):
    print(f"warp {w}: iterations {first}..{first + 15}, mean cycles since loop top (threads={os.environ.get('STB200_FPS_THREADS','256')}, n={n})")
    for k in range(1, 8):
        v = d[:, w, k] - d[:, w, 0]
        v = v[(d[:, w, k] > 0)]
        if v.size:
            print(f"   {names[k]:16s} {v.mean():8.0f}  (min {v.min()}, max {v.max()})")
    it = np.diff(d[:, w, 0])
    print(f"   iteration period {it.mean():.0f} cycles")
