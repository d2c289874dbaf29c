import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
import test_vr as T
R = T.R; G = T.G
def check(name, in_rate, actions, ch=1):
    sc = dict(name=name, in_rate=in_rate, out_rate=1.0, ch=ch, dtype="float32", seed=42, scale=1.0, actions=actions)
    want = G.run(sc, lambda s: R.RefStream(s["in_rate"], s["out_rate"], s["ch"], s["dtype"], "HQ", vr=True), lambda rs, x, last: rs.process_vr(x, last))
    got = T._run_ours(sc)
    lw = [o.shape[0] for o in want]; lg = [o.shape[0] for o in got]
    if lw != lg: print(name, "LENGTHS", lg, lw); return
    w = np.concatenate([o.reshape(o.shape[0], ch) for o in want]); g = np.concatenate([o.reshape(o.shape[0], ch) for o in got])
    d = np.abs(g.astype(np.float64) - w).max(axis=1); bad = np.nonzero(d)[0]
    if bad.size == 0: print(name, "bit-exact", lw)
    else: print(name, "first diff", int(bad[0]), "last", int(bad[-1]), "n", bad.size, "max", d.max(), lw)
C = lambda n, last=False: ["chunk", n, last]
check("up 1->2 immediate small chunks", 6, [["ratio", 3, 1, 0]] + [C(1000)] * 4 + [["ratio", 5, 1, 0]] + [C(1000)] * 6 + [C(500, True)])
check("up 1->2 immediate big chunks", 6, [["ratio", 3, 1, 0]] + [C(6000)] * 2 + [["ratio", 5, 1, 0]] + [C(6000)] * 2 + [C(500, True)])
check("up 1->2 slew", 6, [["ratio", 3, 1, 0]] + [C(1000)] * 4 + [["ratio", 5, 1, 3000]] + [C(1000)] * 12 + [C(500, True)])
check("up 0->1 immediate", 6, [["ratio", 1.5, 1, 0]] + [C(1000)] * 4 + [["ratio", 2.5, 1, 0]] + [C(1000)] * 6 + [C(500, True)])
check("up 0->1->2 slew", 6, [["ratio", 1.5, 1, 0]] + [C(1000)] * 4 + [["ratio", 5.5, 1, 4000]] + [C(2000)] * 12 + [C(500, True)])
check("down 2->1->0 slew", 6, [C(2000)] * 3 + [["ratio", 1.2, 1, 4000]] + [C(2000)] * 8 + [C(500, True)])
check("up max 12: 1->2", 12, [["ratio", 3, 1, 0]] + [C(1000)] * 4 + [["ratio", 5, 1, 0]] + [C(1000)] * 6 + [C(500, True)])
check("up max 12: 2->3", 12, [["ratio", 5, 1, 0]] + [C(2000)] * 4 + [["ratio", 11, 1, 0]] + [C(4000)] * 6 + [C(500, True)])
check("golden-like tail", 6, [["ratio", 0.35, 1, 0], C(2000), ["ratio", 5.9, 1, 4000], C(6000), C(6000), C(6000, True)], ch=2)
check("golden-like tail small chunks", 6, [["ratio", 0.35, 1, 0], C(2000), ["ratio", 5.9, 1, 4000]] + [C(1000)] * 18 + [C(500, True)], ch=2)
