"""Per-CTA phase timeline of root_fused_kernel.  Needs a debug build of the library:
    nvcc ... -DCOLQ_RF_DEBUG -shared -o lib/libcolq_dbg.so csrc/colq.cu   (same flags as the Makefile)
    COLQ_LIB=.../lib/libcolq_dbg.so python scripts/root_fused_timeline.py
Output: profiles/r02_root_fused_cta_timeline_*.txt."""
import os, sys, json, ctypes as C
import pathlib; ROOT = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT / "java-columnar-query-engine_b200")); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np, torch
from colq import _ffi
from colq.device_data import build_geography_on_device, plymouth_colq_query
from colq.engine import ColqContext
U = int(os.environ.get("U", "10000"))
ctx = ColqContext(0)
geo = build_geography_on_device(ctx, U)
q = plymouth_colq_query(ctx)
for k, v in (os.environ.get("OPTS", "") and [kv.split("=") for kv in os.environ["OPTS"].split(",")] or []):
    q.set_option(int(k), int(v))   # e.g. OPTS="9=2": the timeline of root_finish_kernel (stamps: start, pre, prefix, chains, look-back, write, tail, exit)
for _ in range(5):
    r = q.execute(want_indices=False)
buf = np.zeros((4096, 8), dtype=np.uint64)
fn = ctx.lib.colq_debug_rf_times
fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
n = fn(q.handle, buf.ctypes.data_as(C.c_void_p), 4096)
t = buf[:n].astype(np.int64)
t0 = t[:, 0].min()
t = (t - t0) / 1000.0   # us
names = ["start", "A_end", "pre_end", "chains_end", "lookback_end", "write_end", "tail_atomic", "exit"]
if "9=2" in os.environ.get("OPTS", ""):
    names = ["start", "pre_end", "prefix_end", "chains_end", "lookback_end", "write_end", "tail_atomic", "exit"]
print("ctas", n, "count", r.count)
for k, nm in enumerate(names):
    c = t[:, k]
    print(f"{nm:14s} min {c.min():8.1f}  p50 {np.median(c):8.1f}  p90 {np.percentile(c, 90):8.1f}  max {c.max():8.1f}")
d = np.diff(t, axis=1)
for k in range(7):
    c = d[:, k]
    print(f"{names[k]:>12s}->{names[k+1]:12s} min {c.min():7.1f} p50 {np.median(c):7.1f} p90 {np.percentile(c, 90):7.1f} max {c.max():7.1f}")
