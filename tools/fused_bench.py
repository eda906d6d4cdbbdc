"""Development aid: forward time of the per-pair path vs the fused per-window kernel on S3DIS layer shapes."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stratified_transformer_b200 import _cabi, index, pointops2_cuda as ext
from stratified_transformer_b200.synthetic import make_batch

def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
xyz, _, offset = make_batch(scenes, 80000)
xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
h, d, w, quant = 3, 16, 0.16, 0.01
L = 2 * int((2 * w + 1e-4) // quant)
li = index.build_layer_index(xd, od, w, quant, 8)
N = xd.shape[0]
q, k, v = (torch.randn(N, h, d, device="cuda") for _ in range(3))
tq, tk, tv = (torch.randn(L, h, d, 3, device="cuda") * 0.02 for _ in range(3))
stream = torch.cuda.current_stream().cuda_stream
for parity in (0, 1):
    pi = li.for_block(parity)
    M = pi.M
    flags, rows = pi.fused_plan()
    ix = pi.c_struct(L)
    s = torch.empty(M, h, device="cuda"); p = torch.empty(M, h, device="cuda"); out = torch.empty(N, h, d, device="cuda")
    def per_pair():
        _cabi.call("stb200_window_logits_forward", ctypes.byref(ix), h, d, L, q.data_ptr(), k.data_ptr(), tq.data_ptr(), tk.data_ptr(), s.data_ptr(), stream)
        ext.segment_softmax_forward_cuda(N, M, h, s, None, pi.index_0_offsets, p)
        _cabi.call("stb200_window_aggregate_forward", ctypes.byref(ix), h, d, L, p.data_ptr(), v.data_ptr(), tv.data_ptr(), out.data_ptr(), stream)
    def fused():
        _cabi.call("stb200_window_attention_forward_fused", ctypes.byref(ix), pi.n_win, pi.win_offsets.data_ptr(), flags.data_ptr(), h, d, L,
                   q.data_ptr(), k.data_ptr(), v.data_ptr(), tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), out.data_ptr(), p.data_ptr(), stream)
    frac_rows = 1.0 - rows.numel() / N
    print(f"parity {parity}: N={N} M={M} windows={pi.n_win} fused windows={int(flags.sum())} ({frac_rows:.1%} of rows)  "
          f"per-pair fwd {timed(per_pair):.3f} ms   fused kernel (eligible windows only) {timed(fused):.3f} ms", flush=True)

# bf16-storage forward (inference) vs fp32 per-pair forward, parity 0
pi = li.for_block(0); M = pi.M; ix = pi.c_struct(L)
q16, k16, v16 = (t.to(torch.bfloat16).contiguous() for t in (q, k, v))
s = torch.empty(M, h, device="cuda"); p = torch.empty(M, h, device="cuda"); out = torch.empty(N, h, d, device="cuda")
def fwd32():
    _cabi.call("stb200_window_logits_forward", ctypes.byref(ix), h, d, L, q.data_ptr(), k.data_ptr(), tq.data_ptr(), tk.data_ptr(), s.data_ptr(), stream)
    ext.segment_softmax_forward_cuda(N, M, h, s, None, pi.index_0_offsets, p)
    _cabi.call("stb200_window_aggregate_forward", ctypes.byref(ix), h, d, L, p.data_ptr(), v.data_ptr(), tv.data_ptr(), out.data_ptr(), stream)
def fwd16():
    _cabi.call("stb200_window_logits_forward_bf16", ctypes.byref(ix), h, d, L, q16.data_ptr(), k16.data_ptr(), tq.data_ptr(), tk.data_ptr(), s.data_ptr(), stream)
    ext.segment_softmax_forward_cuda(N, M, h, s, None, pi.index_0_offsets, p)
    _cabi.call("stb200_window_aggregate_forward_bf16", ctypes.byref(ix), h, d, L, p.data_ptr(), v16.data_ptr(), tv.data_ptr(), out.data_ptr(), stream)
print(f"forward, layer-0 shape: fp32 {timed(fwd32):.3f} ms   bf16 storage {timed(fwd16):.3f} ms", flush=True)
