"""Quick GPU probe: time the hot path on the RQC configs (not the bench)."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qxb200 as q
from qxb200.executor import Graph, init, bits_from_strings
init(0)
def case(r, c, d, ns, namp, dtype, seed=42):
    t0 = time.time()
    circ = q.create_rqc_circuit(r, c, d, seed)
    tnc = q.convert_to_tnc(circ)
    bg, plan, meta = q.contraction_scheme(tnc, ns, time=0)
    cg = q.build_compute_graph(tnc, plan, bg)
    t1 = time.time()
    g = Graph.from_dsl(cg.dsl(), cg.tensors, dtype).compile()
    t2 = time.time()
    bits = np.random.default_rng(0).integers(0, 2, (namp, r * c)).astype(np.uint8)
    out = g.amplitudes(bits)
    t3 = time.time()
    best = 1e9
    for _ in range(3):
        t = time.time(); out = g.amplitudes(bits); best = min(best, time.time() - t)
    st = g.stats()
    print(f"{r}x{c} d{d} s{ns} {dtype} namp={namp}: plan {t1-t0:.2f}s compile {t2-t1:.2f}s first {t3-t2:.3f}s best {best*1e3:.2f} ms "
          f"-> {namp/best:.0f} amp/s | launches {st['kernel_launches']} bytes {st['bytes']/1e9:.2f} GB ({st['bytes']/best/1e9:.0f} GB/s) "
          f"flops {st['flops']/1e9:.2f} G ws {st['workspace_bytes']/1e9:.2f} GB batch {st['amp_batch']} sum|a|^2*2^n {np.sum(np.abs(out)**2)*2.0**(r*c)/namp:.3f}", flush=True)
    return g
for args in [(4,4,12,4,64,'c64'), (6,6,16,6,1024,'c32'), (6,6,16,6,1024,'c64'), (7,7,20,12,64,'c64'), (7,7,20,12,1024,'c64'), (7,7,20,12,1024,'c32'), (7,7,20,12,4096,'c64')]:
    try:
        case(*args)
    except Exception as e:
        print(args, "FAILED", e, flush=True)
