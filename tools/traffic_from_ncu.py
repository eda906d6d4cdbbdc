"""Turn an ncu per-launch CSV of `bench.py --no-overlap --no-e2e --no-cpu-baseline` into profiles/*_traffic.json.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:'seg_|table_grad|softmax|permute_rows' --csv --log-file gpurun_out/traffic.csv \
        python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-overlap --no-profile
    python tools/traffic_from_ncu.py gpurun_out/traffic.csv profiles/r1_traffic.json

The library's profiler names (`seg_dot[logits_fwd]`, ...) are recovered from the kernel's template arguments and
its position in the fixed per-block launch order of bench.device_step (forward: logits, softmax, aggregate;
backward: grad_attn, grad_table_v, grad_v, softmax bwd, grad_q, grad_table_q, permute, grad_k, grad_table_k).
"""
import collections
import csv
import json
import re
import sys


def profiler_name(kernel: str, seen: dict) -> str | None:
    k = re.sub(r"\(bool\)|\(int\)|stb200::|void ", "", kernel)
    if k.startswith("permute_rows_kernel"):
        return "permute_rows[logits_bwd]"
    m = re.match(r"(\w+)<([^>]*)>", k)
    if not m:
        return None
    fn, args = m.group(1), [a.strip() for a in m.group(2).split(",")]
    flags = [a in ("1", "true") for a in args]
    if fn == "seg_dot_kernel":      # <D, HG, XY, EX, EY>
        return "seg_dot[logits_fwd]" if flags[4] else "seg_dot[aggregate_bwd_gattn]"
    if fn == "seg_reduce_kernel":   # <D, HG, HAS_Y, HAS_T, PERM>
        if flags[4]:
            return "seg_reduce_t[logits_bwd_gk]" if flags[3] else "seg_reduce_t[aggregate_bwd_gv]"
        seen["sr"] = seen.get("sr", 0) + 1          # aggregate_fwd and grad_q share one instantiation: they alternate
        return "seg_reduce[aggregate_fwd]" if seen["sr"] % 2 == 1 else "seg_reduce[logits_bwd_gq]"
    if fn == "table_grad_kernel":   # <D, HGC, PERM, MULTI>
        if flags[2]:
            return "table_grad_t[logits_bwd_gtk]"
        seen["tg"] = seen.get("tg", 0) + 1
        return "table_grad[aggregate_bwd_gtv]" if seen["tg"] % 2 == 1 else "table_grad[logits_bwd_gtq]"
    if fn.startswith("segment_softmax_fwd"):
        return "segment_softmax_fwd"
    if fn.startswith("segment_softmax_bwd"):
        return "segment_softmax_bwd"
    return None


def main(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ci = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[ci["ID"]], {"kernel": r[ci["Kernel Name"]]})
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        d[r[ci["Metric Name"]]] = v
    seen, agg = {}, collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in launches.values():
        name = profiler_name(d["kernel"], seen)
        if name is None:
            continue
        a = agg[name]
        a[0] += 1
        a[1] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a[2] += d.get("gpu__time_duration.sum", 0.0)
    out = {k: {"dram_bytes_per_launch": v[1] / v[0], "launches_captured": v[0], "ms_per_launch_under_ncu": v[2] / v[0] / 1e6}
           for k, v in agg.items()}
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in out.items():
        print(f"{k:36s} {v['launches_captured']:4d} launches  {v['dram_bytes_per_launch'] / 1e6:9.1f} MB  {v['ms_per_launch_under_ncu']:.3f} ms")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
