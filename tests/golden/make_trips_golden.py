#!/usr/bin/env python
"""Generate ``tests/golden/trips_*.npz``: the flow lists the reference's own
``CityFlowTripGenerator`` (pytsc/backends/cityflow/trip_generator.py:45-286) writes.

Run in the BUILD CONTAINER only (imports the reference checkout):

    python tests/golden/make_trips_golden.py [--ref /root/reference]

Stored per case: vehicle start times and routes (CSR of indices into the roadnet's road list), the
generator arguments, and the vehicle template.  ``tests/test_generators.py`` requires
``pytsc_b200.generators.GridTripGenerator`` to reproduce them exactly.
"""
import argparse
import glob
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

# case -> (scenario, start, end, inter_mu, inter_sigma, seed, turn_probs)
CASES = {
    "trips_syn_3x3": ("syn_3x3", 0, 3600, 6.0, 0.8, 0, [0.1, 0.3, 0.6]),
    "trips_syn_1x1_heavy": ("syn_1x1", 0, 1800, 4.0, 0.8, 3, [0.2, 0.2, 0.6]),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    from make_golden import setup_reference
    setup_reference(args.ref)
    from pytsc.backends.cityflow.trip_generator import CityFlowTripGenerator
    for name, (scenario, t0, t1, mu, sigma, seed, probs) in CASES.items():
        gen = CityFlowTripGenerator(scenario, t0, t1, mu, sigma, seed=seed, turn_probs=probs)
        with tempfile.TemporaryDirectory() as d:
            gen.generate_flows(d)
            flows = json.load(open(glob.glob(os.path.join(d, "*.json"))[0]))
        roads = [r["id"] for r in gen.parsed_network.roads]
        ridx = {r: i for i, r in enumerate(roads)}
        off = np.zeros(len(flows) + 1, np.int32)
        flat = []
        for i, f in enumerate(flows):
            flat += [ridx[r] for r in f["route"]]
            off[i + 1] = len(flat)
            assert f["vehicle"] == flows[0]["vehicle"] and f["interval"] == 1.0 and f["endTime"] == f["startTime"]
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, start=np.asarray([f["startTime"] for f in flows], np.int32), route_off=off,
                            route=np.asarray(flat, np.int32), roads=np.asarray(roads), scenario=np.asarray(scenario),
                            args=np.asarray(json.dumps(dict(start_time=t0, end_time=t1, inter_mu=mu, inter_sigma=sigma,
                                                            seed=seed, turn_probs=probs))),
                            vehicle=np.asarray(json.dumps(flows[0]["vehicle"])),
                            max_trip_length=np.asarray(gen.max_trip_length))
        print(f"{name}: {len(flows)} vehicles, max_trip_length {gen.max_trip_length} -> {os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
