"""CPU-only checks of host-side logic that ships in libgfs_b200 (no compute kernels are called)."""
import ctypes as C
import json
import re
import os

import numpy as np

from geoflowslam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "gfs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gfs_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(L, name), "libgfs_b200.so does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes signature for %s" % name
    assert L.gfs_version().decode().startswith("gfs_b200")


def test_gcc_sort_restatement_equals_std_sort():
    """csrc/gcc_sort.h (used by the quadtree kernel) must reproduce libstdc++ std::sort's exact
    permutation on tie-heavy inputs (compareNodes ties on equal size and equal UL.x)."""
    from oracle import oracle as O
    OL = O.lib()
    OL.gfo_std_sort_pairs.argtypes = [C.c_void_p] * 3 + [C.c_int]
    L = _lib.lib()
    rng = np.random.default_rng(0)
    sizes = list(range(0, 40)) + [63, 64, 65, 100, 217, 500, 1000, 4000]
    for n in sizes:
        for trial in range(6):
            hi = [2, 4, 9, 50, 1000, 3][trial]
            first = rng.integers(0, hi, n).astype(np.int32)
            second = rng.permutation(n).astype(np.int32)
            ulx = (rng.integers(0, 6, max(n, 1)) * 38).astype(np.int32)
            if trial == 5:  # organ-pipe-ish input: the classic median-of-3 stress
                first = np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]]).astype(np.int32)
            f1, s1 = first.copy(), second.copy()
            f2, s2 = first.copy(), second.copy()
            OL.gfo_std_sort_pairs(_lib.ptr(f1), _lib.ptr(s1), _lib.ptr(ulx), n)
            hs = C.c_int(0)
            assert L.gfs_debug_gcc_sort(_lib.ptr(f2), _lib.ptr(s2), _lib.ptr(ulx), n, C.byref(hs)) == 0
            assert np.array_equal(f1, f2) and np.array_equal(s1, s2), (n, trial)


def test_gcc_sort_heapsort_fallback_matches_std_sort():
    """Median-of-3 killer sequence drives introsort into its heapsort fallback."""
    from oracle import oracle as O
    OL = O.lib()
    OL.gfo_std_sort_pairs.argtypes = [C.c_void_p] * 3 + [C.c_int]
    L = _lib.lib()
    n = 4096
    k = n // 2
    a = np.zeros(n, np.int32)  # Musser's adversary for median-of-3 quicksort
    for i in range(1, k + 1):
        a[i - 1] = i if i % 2 == 1 else k + i - 1
        a[k + i - 1] = 2 * i
    second = np.arange(n, dtype=np.int32)
    ulx = np.zeros(n, np.int32)
    f1, s1, f2, s2 = a.copy(), second.copy(), a.copy(), second.copy()
    OL.gfo_std_sort_pairs(_lib.ptr(f1), _lib.ptr(s1), _lib.ptr(ulx), n)
    hs = C.c_int(0)
    L.gfs_debug_gcc_sort(_lib.ptr(f2), _lib.ptr(s2), _lib.ptr(ulx), n, C.byref(hs))
    assert np.array_equal(f1, f2) and np.array_equal(s1, s2)
    assert np.all(np.diff(f2) >= 0)


def test_no_device_fails_loudly():
    import pytest
    L = _lib.lib()
    if L.gfs_device_check() == 0:
        pytest.skip("a GPU is present")
    from geoflowslam_b200 import GfsError, ORBextractor
    with pytest.raises(GfsError):
        ORBextractor(1000, 1.2, 8, 25, 7)
    assert b"no CPU fallback" in L.gfs_last_error()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "geoflowslam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "libgfs_oracle" not in src, f


def test_bench_reference_arm_runs_the_whole_step_on_the_cpu():
    """bench.py's CPU arm (cpu_baseline / --impl reference) for the headline workload: data of the named shapes, every stage
    of the step timed per frame, a positive rate.  One sequence, one worker: a few seconds."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    data = bench.make_track_data(1, 1000, 1)
    assert data["gray"].shape == (bench.RING, 1, 480, 640) and data["depth"].shape == (bench.RING, 1, 480, 640)
    assert data["depth"].dtype == np.uint16 and data["imu"].shape == (bench.RING, 1, 7, 7)
    n_valid = int((data["depth"][0, 0, ::2, ::2] > 0).sum())
    assert 45000 < n_valid < 55000                                   # configs[2]: ~50k-point clouds at stride 2
    fps, dt, stage = bench.cpu_track_frames_per_sec(data, 1, 1)
    assert fps > 0 and set(stage) == {"orb_match", "klt", "imu_pose_inertial", "depth_cloud_gicp", "local_inertial_ba"}
    assert stage["depth_cloud_gicp"] > stage["imu_pose_inertial"] > 0
    assert bench.TRACK_METRIC == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
