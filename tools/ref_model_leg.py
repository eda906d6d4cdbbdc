"""bench.py's whole-network reference leg alone (reference Stratified on its own code and kernels vs model.Stratified)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print(json.dumps(bench.ref_model_leg(torch.device("cuda")), indent=1))
