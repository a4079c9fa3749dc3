"""CPU: scenario generation pinned against what the reference ships / writes.

* ``grid_roadnet`` must reproduce the grid roadnets the reference bundles
  (syn_1x1, syn_3x3 -- output of CityFlow's grid generator, which the reference
  shells out to; carried here as lossless .npz bundles) key for key, lane-link
  geometry to 1e-9.
* ``GridTripGenerator`` must reproduce the flow lists recorded from the
  reference's ``CityFlowTripGenerator`` (tests/golden/trips_*.npz, written by
  tests/golden/make_trips_golden.py) exactly: same start times, same routes.
* a generated 4x4 heavy-demand scenario compiles and the CPU oracle runs it
  (the 16x16 grid of BASELINE config 4 is exercised on the GPU).
"""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, ROOT
from pytsc_b200 import bundle
from pytsc_b200.generators import GridTripGenerator, grid_roadnet, synthetic_grid_scenario


def _same(a, b, path=""):
    if isinstance(a, dict):
        assert isinstance(b, dict) and set(a) == set(b), path
        for k in a:
            _same(a[k], b[k], path + "/" + k)
    elif isinstance(a, list):
        assert isinstance(b, list) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    elif isinstance(a, float) or isinstance(b, float):
        assert abs(a - b) <= 1e-9, (path, a, b)
    else:
        assert a == b, (path, a, b)


@pytest.mark.parametrize("name,file,shape", [("syn_1x1", "1x1_roadnet.npz", (1, 1)), ("syn_3x3", "3x3_roadnet.npz", (3, 3))])
def test_grid_roadnet_reproduces_shipped_grids(name, file, shape):
    shipped = bundle.load_roadnet(os.path.join(ROOT, "pytsc_b200", "scenarios", name, file))
    _same(shipped, grid_roadnet(*shape))


@pytest.mark.parametrize("case,shape", [("trips_syn_3x3", (3, 3)), ("trips_syn_1x1_heavy", (1, 1))])
def test_trip_generator_reproduces_reference(case, shape):
    with np.load(os.path.join(GOLDEN, case + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    net = grid_roadnet(*shape)
    roads = [r["id"] for r in net["roads"]]
    assert roads == [str(x) for x in g["roads"]]
    gen = GridTripGenerator(net, **json.loads(str(g["args"])))
    assert gen.max_trip_length == int(g["max_trip_length"])
    flows = gen.generate()
    assert len(flows) == len(g["start"])
    assert [f["startTime"] for f in flows] == list(g["start"])
    ridx = {r: i for i, r in enumerate(roads)}
    flat = [ridx[r] for f in flows for r in f["route"]]
    assert np.array_equal(np.cumsum([0] + [len(f["route"]) for f in flows]), g["route_off"])
    assert np.array_equal(np.asarray(flat), g["route"])
    assert flows[0]["vehicle"] == json.loads(str(g["vehicle"]))


def test_generated_grid_runs_on_the_oracle(tmp_path):
    from oracle.engine import Engine
    net, flows = synthetic_grid_scenario(4, 4, vehicles_per_hour_per_road=900, horizon=600, seed=1)
    assert len(net["intersections"]) == 16 + 16 and len(flows) > 1500
    (tmp_path / "roadnet.json").write_text(json.dumps(net))
    (tmp_path / "flow.json").write_text(json.dumps(flows))
    cfg = dict(dir=str(tmp_path) + os.sep, roadnetFile="roadnet.json", flowFile="flow.json", interval=1.0,
               rlTrafficLight=True, laneChange=False, seed=0, saveReplay=False)
    (tmp_path / "cfg.json").write_text(json.dumps(cfg))
    eng = Engine(str(tmp_path / "cfg.json"))
    for t in range(300):
        if t % 30 == 0:
            for k, it in enumerate(i for i in net["intersections"] if not i["virtual"]):
                eng.set_tl_phase(it["id"], 1 + (t // 30 + k) % 8)
        eng.next_step()
    assert eng.get_vehicle_count() > 200
    assert eng.get_finished_vehicle_count() > 0


# ---- the rest of the generator family (trip_generator.py:289-1027), each pinned to a flow list recorded from the
#      reference's own class (tests/golden/make_trips_golden.py) ----
def _family_generator(g):
    from pytsc_b200 import generators as G
    from pytsc_b200.backend.config import Config
    scenario, cls = str(g["scenario"]), str(g["cls"])
    kw, gen_kw = json.loads(str(g["args"])), json.loads(str(g["gen_args"]))
    np_seed = int(g["np_seed"])
    cfg = Config(scenario, cityflow=dict(flow_rate_type="constant"))
    net = bundle.load_roadnet(cfg.cityflow_roadnet_file)
    t0, t1 = kw.pop("start_time"), kw.pop("end_time")
    if cls == "LinkDisruptedCityFlowTripGenerator":
        gen = G.LinkDisruptedTripGenerator(net, t0, t1, kw["inter_mu"], kw["inter_sigma"], disruption_ratio=kw["disruption_ratio"], seed=kw["seed"])
    elif cls == "FlowDisruptedCityFlowTripGenerator":
        gen = G.FlowDisruptedTripGenerator(net, t0, t1, kw["inter_mu"], kw["inter_sigma"], disruption_ratio=kw["disruption_ratio"], seed=kw["seed"])
    elif cls == "IntervalCityFlowTripGenerator":
        gen = G.IntervalTripGenerator(net, t0, t1, kw["inter_mu"], kw["inter_sigma"], seed=kw["seed"])
        gen_kw.pop("replicate_no")
    elif cls == "VariableDemandTripGenerator":
        gen = G.VariableDemandTripGenerator(net, t0, t1, kw["inter_mus"], kw["inter_sigmas"], kw["edge_weights"], seed=np_seed,
                                            config_seed=cfg.simulator["seed"])
    elif cls == "CityFlowOneWayTripGenerator":
        gen = G.OneWayTripGenerator(net, t0, t1, kw["inter_mu_ns"], kw["inter_sigma_ns"], kw["inter_mu_ew"], kw["inter_sigma_ew"])
    elif cls == "CityFlowRandomizedTripGenerator":
        base = bundle.load_flow(cfg.resolve_flow_file(cfg.simulator["flow_file"]))
        gen = G.RandomizedTripGenerator(net, base, t0, t1, seed=np_seed, config_seed=cfg.simulator["seed"])
    else:
        raise AssertionError(cls)
    return net, gen, gen_kw


@pytest.mark.parametrize("case", ["trips_link_disrupted_syn_3x3", "trips_flow_disrupted_syn_3x3", "trips_interval_syn_3x3",
                                  "trips_variable_demand_syn_3x3", "trips_oneway_syn_5x5", "trips_randomized_hangzhou"])
def test_generator_family_reproduces_reference(case):
    with np.load(os.path.join(GOLDEN, case + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    net, gen, gen_kw = _family_generator(g)
    roads = [r["id"] for r in net["roads"]]
    assert roads == [str(x) for x in g["roads"]]
    assert gen.max_trip_length == int(g["max_trip_length"])
    extra = json.loads(str(g["extra"]))
    if "disrupted_links" in extra:
        assert sorted(gen.disrupted_links) == json.loads(extra["disrupted_links"])
    if "burst_timings" in extra:
        assert {k: list(v) for k, v in gen.burst_timings.items()} == json.loads(extra["burst_timings"])
    flows = gen.generate(**gen_kw)
    assert len(flows) == len(g["start"])
    assert [f["startTime"] for f in flows] == list(g["start"])
    ridx = {r: i for i, r in enumerate(roads)}
    assert np.array_equal(np.cumsum([0] + [len(f["route"]) for f in flows]), g["route_off"])
    assert np.array_equal(np.asarray([ridx[r] for f in flows for r in f["route"]]), g["route"])
