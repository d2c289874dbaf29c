import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
import test_vr as T
for sc in T._meta():
    want = np.load(os.path.join(T.GOLDEN, "vr_vectors.npz"))[sc["name"]]
    outs = T._run_ours(sc)
    lens = [int(o.shape[0]) for o in outs]
    got = np.concatenate([o.reshape(o.shape[0], sc["ch"]) for o in outs], axis=0)
    if lens != sc["lengths"]:
        print(sc["name"], "LENGTHS", lens, sc["lengths"]); continue
    d = np.abs(got.astype(np.float64) - want.astype(np.float64)).max(axis=1)
    bad = np.nonzero(d)[0]
    if bad.size == 0:
        print(sc["name"], "bit-exact", got.shape); continue
    cum = np.cumsum([0] + lens)
    call = int(np.searchsorted(cum, bad[0], side='right') - 1)
    print(sc["name"], "first diff at frame", int(bad[0]), "(call %d, offset %d of %d)" % (call, bad[0] - cum[call], lens[call]),
          "n_bad", bad.size, "last bad", int(bad[-1]), "max", d.max())
    # runs of bad samples
    runs = np.split(bad, np.nonzero(np.diff(bad) > 1)[0] + 1)
    print("  bad runs:", [(int(r[0]), int(r[-1])) for r in runs[:12]], "n_runs", len(runs))
