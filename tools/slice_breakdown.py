#!/usr/bin/env python
"""Per-step device time of one slice of the bench workload (immediate mode, one event pair
per step), aggregated by kernel class and GEMM-view shape."""
import collections, math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa
from picoquant_jl_b200.host import create_RQC
from picoquant_jl_b200.host.backends import parse_dsl
from picoquant_jl_b200.host.b200_backend import B200Backend
from picoquant_jl_b200.host.planner import sweep_plan
from picoquant_jl_b200.host.sliced import record_sliced_contraction
dtype = np.complex64 if (len(sys.argv) > 1 and sys.argv[1] == "c64") else np.complex128
circ = create_RQC(7, 7, 24, seed=0)
rec = record_sliced_contraction(circ, 64, 1, plan_fn=lambda tn, s: sweep_plan(tn, 7, 7, sliced_bonds=s),
                                output_config="0" * 49)
b = B200Backend(dtype)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for rep in range(2):
    shapes = {}
    for cmd, a in parse_dsl(rec.text):
        if cmd == "tensor":
            d = rec.store.read(a["key"]); shapes[a["t"]] = list(d.shape); b.save_tensor_data(a["t"], d)
        elif cmd == "del":
            shapes.pop(a["t"], None); b.delete_tensor(a["t"])
        elif cmd == "view":
            s = list(shapes[a["t"]]); s[a["axis"] - 1] = len(a["idx"]); shapes[a["v"]] = s
            b.view_tensor(a["v"], a["t"], a["axis"], a["idx"])
        elif cmd == "ncon":
            sa, sb = shapes[a["A"]], shapes[a["B"]]
            bset, aset = set(a["b_idx"]), set(a["a_idx"])
            K = int(np.prod([d for d, l in zip(sa, a["a_idx"]) if l in bset] or [1]))
            M = int(np.prod(sa)) // K; N = int(np.prod(sb)) // K
            shapes[a["C"]] = [d for d, l in zip(sa, a["a_idx"]) if l not in bset] + \
                             [d for d, l in zip(sb, a["b_idx"]) if l not in aset]
            b.profile_enable(True)
            b.contract_tensors(a["A"], a["a_idx"], a["B"], a["b_idx"], a["C"])
            prof = b.profile_read(); b.profile_enable(False)
            if rep == 1:
                for cls, r in prof.items():
                    key = (cls, int(math.log2(M)), int(math.log2(N)), int(math.log2(K)))
                    agg[key][0] += r["launches"]; agg[key][1] += r["ms"]; agg[key][2] += r["bytes"]; agg[key][3] += r["flops"]
tot = sum(v[1] for v in agg.values())
print("total eager ms per slice: %.3f" % tot)
big = [kv for kv in agg.items() if kv[1][1] / kv[1][0] > 0.012]
print("launches with avg > 12 us: %d, %.3f ms" % (sum(v[0] for _, v in big), sum(v[1] for _, v in big)))
for key, v in sorted(big, key=lambda kv: -kv[1][1]) + [(("--", 0, 0, 0), [1, 1e-9, 0, 0])] + sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    cls, m, n, k = key
    print("%-16s M=2^%-2d N=2^%-2d K=2^%-2d launches=%4d ms=%7.3f (%4.1f%%) avg_us=%7.1f GB/s=%7.0f TF=%5.1f"
          % (cls, m, n, k, v[0], v[1], 100 * v[1] / tot, 1e3 * v[1] / v[0], v[2] / v[1] / 1e6, v[3] / v[1] / 1e9))
