import sys, torch, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda")
for _ in range(3):
    print(json.dumps({k: v for k, v in bench.ref_layer_leg(dev).items() if k in ("ref_ms", "ours_ms", "speedup")}))
