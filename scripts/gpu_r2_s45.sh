#!/bin/bash
# Round 2, GPU session 45: compute-sanitizer (memcheck, racecheck) over the kernels added late in the round:
# dense-grid GICP (align + tracking mode, incl. a clamped region), KLT with tensor-map loads, ORB FAST with a tensor-map load.
set -x
O=gpurun_out
mkdir -p $O
cat > /tmp/san_gicp.py <<'PY'
import numpy as np, sys, os
sys.path.insert(0, '.')
from geoflowslam_b200 import RegistrationGICP, synth
pairs = [synth.gicp_pair(2080 + i, n_target=3000) for i in range(2)]
stride = max(max(len(t), len(s)) for t, s, _ in pairs)
def pack(cl):
    a = np.zeros((len(cl), stride, 4), np.float32); n = np.zeros(len(cl), np.int32)
    for i, c in enumerate(cl): a[i, :len(c)] = c; n[i] = len(c)
    return a, n
tg, nt = pack([p[0] for p in pairs]); sr, ns = pack([p[1] for p in pairs])
T0 = np.tile(np.eye(4), (2, 1, 1))
reg = RegistrationGICP(max_points=stride, max_pairs=2)
r = reg.align_batch(tg, nt, sr, ns, T0)
reg.track_reset(); reg.track_batch(tg, nt); q = reg.track_batch(sr, ns, T0)
assert r.tobytes() == q.tobytes()
print("gicp ok", r["iterations"], r["num_inliers"])
PY
cat > /tmp/san_klt.py <<'PY'
import numpy as np, sys, cv2
sys.path.insert(0, '.')
from geoflowslam_b200 import KltTracker, synth
f = synth.orb_frames(2, 640, 480, group=2, seed0=1000)
pts = cv2.goodFeaturesToTrack(f[0], 64, 0.01, 7).reshape(-1, 2).astype(np.float32)
pr, st = KltTracker(max_points=128, max_batch=1).fbKltTracking(f[0], f[1], pts, pts)
print("klt ok", st.sum())
PY
cat > /tmp/san_orb.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from geoflowslam_b200 import ORBextractor, synth
f = synth.orb_frames(2, 640, 480, group=2, seed0=1000)
orb = ORBextractor(1000, 1.2, 8, 25, 7, max_size=(640, 480), max_batch=2)
out = orb.extract_batch(f)
print("orb ok", [len(o[0]) for o in out] if isinstance(out, list) else type(out))
PY
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_gicp.py > $O/r2s45_sanitizer_gicp_$tool.log 2>&1; echo "gicp $tool rc=$?"; tail -3 $O/r2s45_sanitizer_gicp_$tool.log
  GFS_GICP_DENSE_CAP=512 timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_gicp.py > $O/r2s45_sanitizer_gicp_clamped_$tool.log 2>&1; echo "gicp clamped $tool rc=$?"; tail -3 $O/r2s45_sanitizer_gicp_clamped_$tool.log
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_klt.py > $O/r2s45_sanitizer_klt_$tool.log 2>&1; echo "klt $tool rc=$?"; tail -3 $O/r2s45_sanitizer_klt_$tool.log
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_orb.py > $O/r2s45_sanitizer_orb_$tool.log 2>&1; echo "orb $tool rc=$?"; tail -3 $O/r2s45_sanitizer_orb_$tool.log
done
