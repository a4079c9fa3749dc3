"""Debug driver (GPU box): lock-step the CUDA engine against the CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from helpers import build_scenario, oracle_engine, signal_inter_indices, compare_snapshots
from pytsc_b200.binding import Engine

def run(name, n_ticks, mode, B=2, flow=None):
    kw = {}
    if flow: kw["cityflow"] = {"flow_file": flow}
    cfg, parser, cs = build_scenario(name, **kw)
    orc = oracle_engine(cfg)
    eng = Engine(cs, B, 0, vehicle_capacity=int(os.environ.get("VCAP", "1024")))
    print(name, "kernel", eng.kernel_info(), flush=True)
    inter = signal_inter_indices(parser)
    A = eng.A
    rng = np.random.RandomState(0)
    nraw = cs.sig_n_raw_phases
    raw = np.ones((B, A), np.int32)
    first_bad = None
    t0 = time.time()
    for t in range(n_ticks):
        if t % 5 == 0:
            if mode == "random":
                r = np.array([rng.randint(0, nraw[a]) for a in range(A)], np.int32)
            else:  # cyclic plan: 25 s green, 5 s yellow(0)
                k = (t // 30) % 8
                r = np.full(A, (k + 1) if (t % 30) < 25 else 0, np.int32)
            raw[:] = r
            eng.set_phase(torch.from_numpy(raw).cuda())
            for a in range(A):
                orc.set_tl_phase_idx(inter[a], int(r[a]))
        orc.next_step()
        eng.step(1)
        if t % int(os.environ.get("CHECK_EVERY", "1")) == 0 or t == n_ticks - 1:
            so = orc.snapshot()
            for b in (0, B - 1):
                sg = eng.snapshot(b)
                msg = compare_snapshots(so, sg)
                if msg:
                    print(f"[{name}/{mode}] tick {t} replica {b}: {msg}", flush=True)
                    first_bad = t
                    break
            if first_bad is not None:
                break
    try:
        eng.check()
    except Exception as e:
        print("check:", e)
    c = eng.counters()
    print(f"[{name}/{mode}] ticks={t+1} ok={first_bad is None} running={c['n_running'][0]} finished={c['n_finished'][0]} "
          f"oracle running={orc.get_vehicle_count()} finished={orc.get_finished_vehicle_count()} nonfifo={orc.non_fifo_events()} "
          f"wall={time.time()-t0:.1f}s", flush=True)
    return first_bad is None

if __name__ == "__main__":
    ok = True
    n = int(os.environ.get("TICKS", "600"))
    for name, flow in [("syn_1x1", None), ("hangzhou_4_4", None)]:
        for mode in ("cyclic", "random"):
            ok &= run(name, n, mode, flow=flow)
    print("ALL OK" if ok else "MISMATCH")
