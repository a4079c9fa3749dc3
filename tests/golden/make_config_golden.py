#!/usr/bin/env python
"""Generate ``tests/golden/scenario_configs.json``: what the REFERENCE's own Config + CityFlowNetworkParser make of every
bundled scenario's config.yaml (action space, round robin, reward, flow rate type, per-signal phase plan).

Run in the BUILD CONTAINER only (imports the reference checkout):  python tests/golden/make_config_golden.py
``tests/test_scenario_configs.py`` requires pytsc_b200's Config / NetworkParser on the bundled scenarios to agree.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

SCENARIOS = ["hangzhou_4_4", "jinan_3_4", "manhattan_16_3", "syn_1x1", "syn_3x3", "syn_1x3_gaussian", "syn_5x5_oneway",
             "new_york_arterial"]


def main():
    from make_golden import setup_reference
    setup_reference("/root/reference")
    from pytsc.backends.cityflow.config import Config
    from pytsc.backends.cityflow.network_parser import NetworkParser
    out = {}
    for s in SCENARIOS:
        cfg = Config(s)
        parser = NetworkParser(cfg)
        ts = parser.traffic_signals
        out[s] = {
            "signal": {k: cfg.signal[k] for k in ("action_space", "round_robin", "reward_function", "observation_space")},
            "flow_rate_type": cfg.simulator.get("flow_rate_type", "constant"),
            "phase_sequence": cfg.simulator.get("phase_sequence"),
            "signals": [[t, {"phases": list(c["phases"]), "n_phases": c["n_phases"],
                            "green_phase_indices": list(c["green_phase_indices"]),
                            "yellow_phase_indices": list(c["yellow_phase_indices"]),
                             "n_incoming": len(c["incoming_lanes"]), "n_outgoing": len(c["outgoing_lanes"])}]
                        for t, c in ts.items()],      # a list: agent order (roadnet order) matters
        }
        print(s, out[s]["signal"], out[s]["flow_rate_type"], len(ts), "signals")
    with open(os.path.join(HERE, "scenario_configs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
