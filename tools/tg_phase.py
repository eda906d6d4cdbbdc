"""Development aid: per-phase clock stamps of the table-gradient kernel (CTA 0, first 64 tiles) on the layer-0 shape.
Needs a library built with the TG_STAMP instrumentation (see git history of this tool); run with STB200_LIB=<that .so>."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from stratified_transformer_b200 import _cabi

class A: scenes = 8; points = 80000
dev = torch.device("cuda")
levels, _ = bench.build_inputs(A, 0, dev)
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
bench.device_step(levels[lvl:lvl + 1], [])      # the last launch of this call is table_grad_t[logits_bwd_gtk] of that level
torch.cuda.synchronize()
lib = _cabi.load()
buf = (ctypes.c_longlong * (64 * 8))()
lib.stb200_tg_debug_read.argtypes = [ctypes.c_void_p]
assert lib.stb200_tg_debug_read(buf) == 0
d = np.array(buf, dtype=np.int64).reshape(64, 8)
d = d[d[:, 0] > 0]
names = ["top", "zero+prologue issued", "B1 passed", "X staged (issued)", "chunk0 staged (issued)", "B2 passed", "hist done", "B3 passed"]
print(f"level {lvl}: {len(d)} tiles stamped; mean cycles since tile top")
for k in range(1, 8):
    print(f"   {names[k]:26s} {np.mean(d[:, k] - d[:, 0]):9.0f}")
print(f"   tile period                {np.mean(np.diff(d[:, 0])):9.0f}   (MMA + loop = period - B3)")
