"""CPU: host-side helpers of bench.py that run before any GPU work."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rank_cpu_slices_are_disjoint_and_cover_whole_cores():
    """Two ranks of one host pin themselves to disjoint shares of the allowed CPUs (bench.bind_to_rank_cpu_slice); a single
    rank is left alone.  Run in child processes: the affinity change must not leak into the test session."""
    code = ("import os, sys, json; sys.path.insert(0, %r); import bench; "
            "d = bench.bind_to_rank_cpu_slice(int(sys.argv[1]), int(sys.argv[2])); "
            "print(json.dumps([d, sorted(os.sched_getaffinity(0))]))" % ROOT)
    import json
    allowed = sorted(os.sched_getaffinity(0))
    got = []
    for rank in (0, 1):
        out = subprocess.run([sys.executable, "-c", code, str(rank), "2"], capture_output=True, text=True, check=True).stdout
        got.append(json.loads(out.strip().splitlines()[-1]))
    if len(allowed) >= 2 and got[0][0] and "rank slice" in got[0][0]:
        a, b = set(got[0][1]), set(got[1][1])
        assert a and b and not (a & b) and (a | b) <= set(allowed)
    else:      # one CPU, or no permission: nothing was changed
        assert got[0][1] == allowed or got[0][0] is None or "unchanged" in got[0][0]
    out = subprocess.run([sys.executable, "-c", code, "0", "1"], capture_output=True, text=True, check=True).stdout
    d, cpus = json.loads(out.strip().splitlines()[-1])
    assert d is None and cpus == allowed
