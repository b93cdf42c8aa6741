import time, numpy as np, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from pcc_geo_cnn_v2_b200 import ops
from pcc_geo_cnn_v2_b200.entropy_models import gaussian_tables, make_scale_table
tab = gaussian_tables(make_scale_table())
rng = np.random.default_rng(0)
n=32; per=32768
idx = rng.integers(0, 37, size=n*per).astype(np.int32)
center = -tab['offset'][idx]
sym = np.rint(rng.normal(size=n*per)*(center/2.9)).astype(np.int32)
offs = (np.arange(n+1)*per).astype(np.int64)
for th in (1,4,16):
    t0=time.perf_counter(); s=ops.range_encode(sym, offs, tab, indexes=idx, threads=th); t1=time.perf_counter()
    d=ops.range_decode(s, offs, tab, indexes=idx, threads=th); t2=time.perf_counter()
    assert np.array_equal(d,sym)
    print(f'threads={th}: encode {1e9*(t1-t0)/(n*per)*th:.1f} ns/sym/thread ({(t1-t0)*1e3:.1f} ms), decode {1e9*(t2-t1)/(n*per)*th:.1f} ns/sym/thread ({(t2-t1)*1e3:.1f} ms), bytes/stream {sum(map(len,s))/n:.0f}')
