"""Developer tool: digest of a chrome trace written by tools/trace_forward.py (per-kernel sums, per-stream busy, long kernels)."""
import collections
import json
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_bf16.json"
ev = json.load(open(path))["traceEvents"]
k = sorted([e for e in ev if e.get("cat") == "kernel"], key=lambda e: e["ts"])
t0 = k[0]["ts"]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in k:
    a = agg[re.sub(r"\(.*", "", e["name"])]
    a[0] += 1
    a[1] += e["dur"]
for n, (c, d) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print(f"{d:9.1f} us {c:4d}x {n}")
st = collections.defaultdict(float)
for e in k:
    st[e["args"].get("stream")] += e["dur"]
print({s: round(v) for s, v in st.items()}, "span", round(max(e["ts"] + e["dur"] for e in k) - t0))
if len(sys.argv) > 3:
    for e in k:
        if str(e["args"].get("stream")) == sys.argv[3]:
            print(f'{e["ts"] - t0:8.0f} {e["dur"]:7.1f} {re.sub(r"\(.*", "", e["name"])[:60]} g{e["args"].get("grid")[0]}')
