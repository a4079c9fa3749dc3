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


# the rest of the generator family (trip_generator.py:289-1027): case -> (class name, scenario, constructor kwargs,
# generate_flows kwargs, numpy seed set before construction for the classes that never seed themselves)
FAMILY_CASES = {
    "trips_link_disrupted_syn_3x3": ("LinkDisruptedCityFlowTripGenerator", "syn_3x3",
                                     dict(start_time=0, end_time=1800, inter_mu=6.0, inter_sigma=0.8, disruption_ratio=0.25, seed=4), {}, None),
    "trips_flow_disrupted_syn_3x3": ("FlowDisruptedCityFlowTripGenerator", "syn_3x3",
                                     dict(start_time=0, end_time=2400, inter_mu=6.0, inter_sigma=0.8, disruption_ratio=0.1, seed=2), {}, None),
    "trips_interval_syn_3x3": ("IntervalCityFlowTripGenerator", "syn_3x3",
                               dict(start_time=0, end_time=1440, inter_mu=7.0, inter_sigma=0.8, seed=5),
                               dict(replicate_no=1, interval_duration=360, shape=1.5, scale=300), None),
    "trips_variable_demand_syn_3x3": ("VariableDemandTripGenerator", "syn_3x3",
                                      dict(start_time=0, end_time=3000, inter_mus="EDGES:6.0", inter_sigmas="EDGES:0.8", edge_weights=None), {}, 11),
    "trips_oneway_syn_5x5": ("CityFlowOneWayTripGenerator", "syn_5x5_oneway",
                             dict(start_time=0, end_time=1800, inter_mu_ns=7.2, inter_sigma_ns=0.8, inter_mu_ew=4.8, inter_sigma_ew=0.8), {}, None),
    "trips_randomized_hangzhou": ("CityFlowRandomizedTripGenerator", "hangzhou_4_4", dict(start_time=0, end_time=1200),
                                  dict(flow_type="medium"), 7),
}


def family_cases(ref):
    import importlib
    tg = importlib.import_module("pytsc.backends.cityflow.trip_generator")
    for name, (cls, scenario, kw, gen_kw, np_seed) in FAMILY_CASES.items():
        kw = dict(kw)
        if np_seed is not None:
            np.random.seed(np_seed)
        probe = tg.CityFlowTripGenerator(scenario, 0, 10, 5.0, 1.0) if any(isinstance(v, str) and v.startswith("EDGES:") for v in kw.values()) else None
        for k, v in list(kw.items()):
            if isinstance(v, str) and v.startswith("EDGES:"):      # the same value for every second incoming fringe road
                inc, _ = probe._find_fringe_edges()
                kw[k] = {e: float(v[6:]) * (1 + 0.25 * (i % 3)) for i, e in enumerate(inc) if i % 2 == 0}
        if np_seed is not None:
            np.random.seed(np_seed)
        gen = getattr(tg, cls)(scenario, **kw)
        with tempfile.TemporaryDirectory() as d:
            gen.generate_flows(d, **gen_kw)
            flows = json.load(open(glob.glob(os.path.join(d, "**", "*.json"), recursive=True)[0]))
        roads = [r["id"] for r in gen.parsed_network.roads]
        ridx = {r: i for i, r in enumerate(roads)}
        off = np.zeros(len(flows) + 1, np.int32)
        flat = []
        for i, f in enumerate(flows):
            flat += [ridx[r] for r in f["route"]]
            off[i + 1] = len(flat)
        extra = {}
        for attr in ("disrupted_links", "burst_timings"):
            if hasattr(gen, attr):
                v = getattr(gen, attr)
                extra[attr] = json.dumps(sorted(v) if isinstance(v, set) else {k: list(t) for k, t in v.items()})
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, start=np.asarray([f["startTime"] for f in flows], np.int32), route_off=off,
                            route=np.asarray(flat, np.int32), roads=np.asarray(roads), scenario=np.asarray(scenario),
                            cls=np.asarray(cls), args=np.asarray(json.dumps(kw)), gen_args=np.asarray(json.dumps(gen_kw)),
                            np_seed=np.asarray(-1 if np_seed is None else np_seed), extra=np.asarray(json.dumps(extra)),
                            max_trip_length=np.asarray(int(gen.max_trip_length)))
        print(f"{name}: {cls} -> {len(flows)} vehicles, max_trip_length {gen.max_trip_length}, {os.path.getsize(path) / 1024:.0f} KB {extra}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    from make_golden import setup_reference
    setup_reference(args.ref)
    family_cases(args.ref)
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
