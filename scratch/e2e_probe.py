import sys, time
sys.path.insert(0, "java-columnar-query-engine_b200"); sys.path.insert(0, "oracle")
import numpy as np, torch
from colq import _ffi, geography as G
from colq.device_data import build_geography_on_device, plymouth_colq_query
from colq.engine import ColqContext
U = 10000
ctx = ColqContext(0)
base = G.load_base()
geo = build_geography_on_device(ctx, U, base=base)
host = {}
for k, t in geo.tensors.items():
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True); h.copy_(t); host[k] = h.numpy()
torch.cuda.synchronize()
for tb in (geo.zips, geo.cities, geo.states): ctx.table_destroy(tb)
geo.tensors.clear(); ctx._keepalive.clear(); torch.cuda.empty_cache()
nz, nc, nb = geo.n_zip_rows, geo.n_city_rows, geo.name_bytes
def step(trace):
    T = [time.perf_counter()]
    def mark(name):
        T.append(time.perf_counter()); trace.append((name, (T[-1]-T[-2])*1e3))
    states = ctx.table_create(51, 0, 0); cities = ctx.table_create(nc, 0, 0); zips = ctx.table_create(nz, 0, 0)
    ctx.col_str(states, 0, base["state_code_offsets"], base["state_code_bytes"])
    ctx.col_str(states, 1, base["state_name_offsets"], base["state_name_bytes"]); mark("states")
    ctx.col_str(cities, 0, host["city_name_offsets"][: nc + 1].view(np.uint32), host["city_name_bytes"][:nb]); mark("city names 3.2GB")
    ctx.associate_fk(cities, 1, states, 2, host["city_state"][:nc]); mark("city_state 1.03GB")
    ctx.col_i32(zips, 0, host["zip_code"][:nz]); mark("zip_code 1.17GB")
    ctx.col_i32(zips, 1, host["zip_pop"][:nz]); mark("zip_pop 1.17GB")
    ctx.associate_fk(zips, 2, cities, 2, host["zip_city"][:nz]); mark("zip_city 1.17GB")
    ctx.associate_csr(states, 3, states, 4, base["adj_offsets"].astype(np.int64), base["adj_targets"])
    for name, tb in (("states", states), ("cities", cities), ("zips", zips)): ctx.register(name, tb)
    mark("adj+register")
    qq = plymouth_colq_query(ctx); mark("query build")
    r = qq.execute(want_indices=True, index_capacity=31 * U + 16); mark("execute+fetch")
    qq.close(); mark("query close")
    for tb in (zips, cities, states): ctx.table_destroy(tb)
    mark("destroy")
    return r
for i in range(3):
    tr = []; t0 = time.perf_counter(); r = step(tr); dt = (time.perf_counter()-t0)*1e3
    print(i, round(dt,1), "ms", [(n, round(v,1)) for n, v in tr], r.count)
