"""Quick A/B helper: Plymouth at U universes (env U, default 10000) on cuda:0, back-to-back step time + per-launch stage times.
Env: OPTS="9=0,4=2" sets colq_query options, COLQ_LIB picks another build, TAG labels the output line.
Used for the A/B tables in profiles/ (r02_root_plan_ab.txt, r02_scan_rows_tma_ab.txt)."""
import os, sys, json
import pathlib; ROOT = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT / "java-columnar-query-engine_b200")); sys.path.insert(0, str(ROOT / "oracle"))
import torch
from colq import _ffi, geography as G
from colq.device_data import build_geography_on_device, plymouth_colq_query
from colq.engine import ColqContext
U = int(os.environ.get("U", "10000"))
ctx = ColqContext(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
geo = build_geography_on_device(ctx, U)
q = plymouth_colq_query(ctx)
for k, v in (os.environ.get("OPTS", "") and [kv.split("=") for kv in os.environ["OPTS"].split(",")] or []):
    q.set_option(int(k), int(v))
res = q.execute(want_indices=True, index_capacity=31 * U + 16)
with torch.cuda.stream(stream):
    for _ in range(5): q.execute_async()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(50): q.execute_async()
    e1.record(stream); stream.synchronize()
ms = e0.elapsed_time(e1) / 50
q.set_option(_ffi.OPT_PROFILE, 1)
acc = {}
for _ in range(10):
    q.execute(want_indices=False)
    for name, t, r, b in q.profile():
        if t >= 0: acc.setdefault(name, []).append(t)
print(json.dumps({"tag": os.environ.get("TAG", ""), "count": res.count, "ms_step": round(ms, 4), "stages": {k: round(sum(v) / len(v), 4) for k, v in acc.items()}}))
