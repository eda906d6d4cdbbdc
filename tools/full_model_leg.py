"""bench.py's full-model leg alone (BASELINE configs[1] as written: whole S3DIS network, 8 x 80k points)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prefetch = len(sys.argv) > 2 and sys.argv[2] == "prefetch"
print(json.dumps(bench.full_model_leg(torch.device("cuda"), scenes, 80000, prefetch=prefetch), indent=1))
