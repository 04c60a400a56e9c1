import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from geoflowslam_b200 import ORBextractor, synth
from oracle import oracle as O
cfg = (1000, 1.2, 8, 25, 7)
frames = synth.orb_frames(2)
ex = ORBextractor(*cfg, max_size=(752, 480), max_batch=2)
o = O.OrbOracle(*cfg)
img = frames[0]
ko, do, mo = o.extract(img)
mono, kg, dg = ex(img)
print("n", len(kg), len(ko), "mono", mono, mo)
for l in range(8):
    pg, po = ex.image_pyramid_level(l), o.level(l)
    print("L", l, "pyr mismatch", int((pg != po).sum()), end=" ")
    bg, bo = ex.image_pyramid_level(l, blurred=True), o.level(l, blurred=True)
    print("blur mismatch", int((bg != bo).sum()), end=" ")
    cg, co = ex.fast_candidates(l), o.candidates(l)
    same = cg.shape == co.shape and np.array_equal(cg, co)
    print("cand", len(cg), len(co), "same" if same else "DIFF", end=" ")
    if not same:
        sg = set(map(tuple, cg.tolist())); so = set(map(tuple, co.tolist()))
        print("only_gpu", sorted(sg - so)[:5], "only_orc", sorted(so - sg)[:5], end=" ")
        if sg == so:
            for i in range(min(len(cg), len(co))):
                if not np.array_equal(cg[i], co[i]):
                    print("first order diff at", i, cg[i], co[i]); break
    ng = int((kg["octave"] == l).sum()); no = int((ko["octave"] == l).sum())
    print("sel", ng, no)
n = min(len(kg), len(ko))
for f in ("x", "y", "size", "angle", "response", "octave"):
    d = np.nonzero(kg[f][:n] != ko[f][:n])[0]
    print(f, "diffs", len(d), d[:5], kg[f][d[:3]], ko[f][d[:3]])
dd = np.nonzero((dg[:n] != do[:n]).any(1))[0]
print("desc diffs", len(dd), dd[:10])
