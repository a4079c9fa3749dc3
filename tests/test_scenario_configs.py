"""CPU: the bundled scenarios carry the reference's own config.yaml (signal section, cityflow.phase_sequence, flow rate
type): what pytsc_b200's Config / NetworkParser derive from them equals what the reference's Config / CityFlowNetworkParser
derive from the reference's files (tests/golden/scenario_configs.json, written by tests/golden/make_config_golden.py)."""
import json
import os

import pytest

from helpers import GOLDEN
from pytsc_b200.backend.config import Config
from pytsc_b200.backend.network_parser import NetworkParser

with open(os.path.join(GOLDEN, "scenario_configs.json")) as f:
    EXPECT = json.load(f)


@pytest.mark.parametrize("scenario", sorted(EXPECT))
def test_bundled_scenario_reproduces_reference_config(scenario):
    want = EXPECT[scenario]
    cfg = Config(scenario)
    for k, v in want["signal"].items():
        assert cfg.signal[k] == v, (scenario, k)
    assert cfg.simulator.get("phase_sequence") == want["phase_sequence"]
    # the flow rate type is the reference's wherever the flow files it lists are really shipped
    if cfg.simulator.get("flow_files"):
        assert cfg.simulator.get("flow_rate_type", "constant") == want["flow_rate_type"]
    parser = NetworkParser(cfg)
    assert list(parser.traffic_signals) == [t for t, _ in want["signals"]]      # agent order = roadnet order
    for (t, c), (_, w) in zip(parser.traffic_signals.items(), want["signals"]):
        assert list(c["phases"]) == w["phases"] and c["n_phases"] == w["n_phases"], (scenario, t)
        assert list(c["green_phase_indices"]) == w["green_phase_indices"]
        assert list(c["yellow_phase_indices"]) == w["yellow_phase_indices"]
        assert len(c["incoming_lanes"]) == w["n_incoming"] and len(c["outgoing_lanes"]) == w["n_outgoing"]


def test_phase_switch_scenario_compiles_to_two_actions():
    from helpers import build_scenario
    cfg, parser, cs = build_scenario("syn_5x5_oneway")
    assert cfg.signal["action_space"] == "phase_switch" and cs.n_actions == 2 and cs.round_robin == 1
    cfg, parser, cs = build_scenario("new_york_arterial")
    assert cs.n_actions == max(c["n_phases"] for c in parser.traffic_signals.values()) == 8
