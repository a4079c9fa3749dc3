"""CPU: structural invariants of the compiled scenario tables the CUDA engine trusts (host logic,
pytsc_b200/scenario.py + roadnet.py), for every bundled scenario: route sequences alternate lane /
lane-link and are connected, cross tables are reciprocal and sorted, CSR offsets are monotone, spawn lists
are per-lane FIFO in creation order, and the pytsc tables agree with the network parser."""
import numpy as np
import pytest

from helpers import build_scenario

SCENARIOS = ["syn_1x1", "syn_3x3", "hangzhou_4_4", "jinan_3_4", "manhattan_16_3",
             "syn_1x3_gaussian", "syn_5x5_oneway", "new_york_arterial"]


@pytest.fixture(scope="module", params=SCENARIOS)
def compiled(request):
    return request.param, build_scenario(request.param)


def test_routes_alternate_and_connect(compiled):
    name, (cfg, parser, cs) = compiled
    L, K = cs.n_lanes, cs.n_lanelinks
    seq = np.asarray(cs.route_seq)
    start = np.asarray(cs.veh_seq_start)
    sl, el = np.asarray(cs.ll_start_lane), np.asarray(cs.ll_end_lane)
    assert (seq < L + K).all() and (seq >= -1).all()
    checked = set()
    for v in range(cs.n_vehicles):
        q = int(start[v])
        if q in checked:
            continue
        checked.add(q)
        assert 0 <= seq[q] < L, "a route starts on a lane"
        assert q == 0 or seq[q - 1] == -1, "sequences are separated by -1"
        while seq[q + 1] != -1:
            a, b = int(seq[q]), int(seq[q + 1])
            if a < L:
                assert b >= L and sl[b - L] == a, (name, v, "lane -> a lane-link that starts there")
            else:
                assert b < L and el[a - L] == b, (name, v, "lane-link -> its end lane")
            q += 1
        assert seq[q] < L, "a route ends on a lane"


def test_cross_tables_are_reciprocal_and_sorted(compiled):
    name, (cfg, parser, cs) = compiled
    K = cs.n_lanelinks
    off = np.asarray(cs.ll_cross_off)
    dist, foe, fdist = np.asarray(cs.xr_dist), np.asarray(cs.xr_foe_ll), np.asarray(cs.xr_foe_dist)
    assert off[0] == 0 and off[K] == cs.n_cross_entries and (np.diff(off) >= 0).all()
    length = np.asarray(cs.drv_length)[cs.n_lanes:]
    sig = np.asarray(cs.ll_signal)
    for k in range(K):
        d = dist[off[k]:off[k + 1]]
        assert (np.diff(d) >= 0).all(), "crosses of a link in ascending distance"
        assert (d >= -1e-9).all() and (d <= length[k] + 1e-9).all()
        for x in range(off[k], off[k + 1]):
            f = int(foe[x])
            assert f != k and sig[f] == sig[k], "a cross joins two links of the same intersection"
            back = [y for y in range(off[f], off[f + 1]) if foe[y] == k and dist[y] == fdist[x] and fdist[y] == dist[x]]
            assert back, (name, k, f, "the other link lists the same cross with the distances swapped")
    assert (np.diff(off) <= 255).all(), "the flat cross phase packs the cross position into 8 bits"


def test_csr_tables_and_spawn_lists(compiled):
    name, (cfg, parser, cs) = compiled
    L, A, N = cs.n_lanes, cs.n_signals, cs.n_vehicles
    for offs, n, total in ((cs.lane_ll_off, L, len(cs.lane_ll)), (cs.lane_spawn_off, L, N), (cs.sig_in_off, A, cs.n_in_total),
                           (cs.sig_out_off, A, cs.n_out_total), (cs.nbr_off, A, cs.n_nbr_total)):
        o = np.asarray(offs)
        assert len(o) == n + 1 and o[0] == 0 and o[n] == total and (np.diff(o) >= 0).all()
    # every lane-link is listed once, under its start lane
    sl = np.asarray(cs.ll_start_lane)
    llo, ll = np.asarray(cs.lane_ll_off), np.asarray(cs.lane_ll)
    assert sorted(ll.tolist()) == list(range(cs.n_lanelinks))
    for l in range(L):
        assert (sl[ll[llo[l]:llo[l + 1]]] == l).all()
    # spawn lists: every vehicle once, on the first lane of its route, creation ticks non-decreasing per lane
    so, sv = np.asarray(cs.lane_spawn_off), np.asarray(cs.lane_spawn_vid)
    tick, start, seq = np.asarray(cs.veh_tick), np.asarray(cs.veh_seq_start), np.asarray(cs.route_seq)
    assert sorted(sv.tolist()) == list(range(N))
    for l in range(L):
        vs = sv[so[l]:so[l + 1]]
        assert (seq[start[vs]] == l).all()
        assert (np.diff(tick[vs]) >= 0).all()
    assert tick.min() >= 0 and tick.max() <= cs.horizon_ticks
    assert len(set(np.asarray(cs.veh_priority).tolist())) == N, "priorities are a total order (canPass's last tie-break)"


def test_pytsc_tables_agree_with_the_network_parser(compiled):
    name, (cfg, parser, cs) = compiled
    lane_index = {l: k for k, l in enumerate(cs.lane_ids)}
    for a, ts in enumerate(cs.signal_ids):
        tcfg = parser.traffic_signals[ts]
        i0, i1 = cs.sig_in_off[a], cs.sig_in_off[a + 1]
        o0, o1 = cs.sig_out_off[a], cs.sig_out_off[a + 1]
        assert [cs.lane_ids[l] for l in cs.sig_in_lane[i0:i1]] == list(tcfg["incoming_lanes"])
        assert [cs.lane_ids[l] for l in cs.sig_out_lane[o0:o1]] == list(tcfg["outgoing_lanes"])
        assert cs.sig_n_phases[a] == tcfg["n_phases"]
        P = cs.max_phases
        green = np.asarray(cs.sig_phase_green)[a * P:a * P + tcfg["n_phases"]]
        assert np.flatnonzero(green).tolist() == list(tcfg["green_phase_indices"])
    plen = np.asarray(cs.lane_pytsc_length)
    for l, k in lane_index.items():
        assert plen[k] == pytest.approx(parser.lane_lengths[l], rel=0, abs=1e-9)
    assert cs.obs_dim == cs.max_lanes_per_signal * (12 if cs.obs_type == 0 else cs.visibility + 9) + cs.max_obs_phases
