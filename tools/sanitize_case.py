"""One short fused episode for compute-sanitizer (tools/gpu_sanitize.sh): 2 replicas of the heavy Hangzhou
flow file, `ticks` ticks of the in-kernel controller + every retrieve output, then tsc_check.

    compute-sanitizer --tool racecheck python tools/sanitize_case.py --capacity 600 --controller fixed_time
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--capacity", type=int, default=600)
    ap.add_argument("--ticks", type=int, default=150)
    ap.add_argument("--controller", default="fixed_time")
    ap.add_argument("--scenario", default="hangzhou_4_4")
    ap.add_argument("--obs", default="lane_features")
    ap.add_argument("--registered", action="store_true", help="step through the registered host path (packets + host threads)")
    args = ap.parse_args()
    import torch
    from helpers import build_scenario
    from pytsc_b200.binding import CONTROLLERS, Engine
    kw = dict(signal=dict(observation_space=args.obs))
    if args.scenario == "hangzhou_4_4":
        kw["cityflow"] = {"flow_file": "anon_4_4_hangzhou_real_5816.json"}
    cfg, parser, cs = build_scenario(args.scenario, **kw)
    eng = Engine(cs, 2, 0, vehicle_capacity=args.capacity)
    bufs = eng.alloc_outputs()
    eng.init_program(0)
    arg = 25 if args.controller == "fixed_time" else 3
    if args.registered:
        import numpy as np
        d = eng.dims
        eng.host_register(obs=np.empty((d["B"], d["A"], d["obs_dim"]), np.float32), reward=np.empty((d["B"], d["A"]), np.float32),
                          mask=np.empty((d["B"], d["A"], d["n_actions"]), np.uint8), reward_global=np.empty((d["B"],), np.float32))
    for _ in range(args.ticks // 5):
        if args.registered:
            eng.env_step_registered(None, n_ticks=5, controller=CONTROLLERS[args.controller], controller_arg=arg)
        else:
            eng.env_step(None, bufs, n_ticks=5, controller=CONTROLLERS[args.controller], controller_arg=arg)
    torch.cuda.synchronize()
    eng.check()
    c = eng.counters()
    print("sanitize_case ok", eng.kernel_info(), "running", int(c["n_running"][0]), "finished", int(c["n_finished"][0]), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
