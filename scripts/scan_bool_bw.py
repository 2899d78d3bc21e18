"""scan_bool bandwidth: 2^30 boolean rows on cuda:0 (1.07 GB read + 134 MB mask written per launch)."""
import sys, json, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "java-columnar-query-engine_b200"))
import numpy as np, torch
from colq import _ffi
from colq.engine import ColqContext
n = 1 << 30
ctx = ColqContext(0)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
rng = np.random.default_rng(5)
flags = (rng.integers(0, 1000, size=n, dtype=np.int16) == 0).astype(np.uint8)   # 0.1 % TRUE
t = ctx.table_create(n)
ctx.col_bool(t, 0, flags)
ctx.register("t", t)
q = ctx.query("t")
q.criteria_bool(0, 0, False, True)
res = q.execute(want_indices=True, index_capacity=int(flags.sum()) + 16)
assert res.count == int(flags.sum()), (res.count, int(flags.sum()))
assert np.array_equal(res.indices, np.flatnonzero(flags).astype(np.int32))
q.set_option(_ffi.OPT_PROFILE, 1)
acc = {}
for _ in range(10):
    q.execute(want_indices=False)
    for name, tt, r, b in q.profile():
        if tt >= 0: acc.setdefault(name, []).append(tt)
st = {k: round(sum(v) / len(v), 4) for k, v in acc.items()}
bytes_alg = n + n // 8
print(json.dumps({"rows": n, "count": res.count, "stages_ms": st, "scan_bool_GBps": round(bytes_alg / (st.get("scan_bool", 1e9) * 1e-3) / 1e9, 1)}))
