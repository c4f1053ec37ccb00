"""The roofline numerator bench.py reports: algorithmic bytes per env-step (SURVEY.md 8d -- reference dtypes, every state byte
read once and written once, actions read, rewards / flags written) for the workloads DESIGN.md quotes.  Pinned so that the
roofline fraction cannot move through its bookkeeping."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402


def _bytes(workload):
    comp, cfg = bench.workload_config(workload=workload)
    return bench.algorithmic_bytes_per_env_step(comp, cfg, bench._packed_state_words(comp, cfg))


def test_headline_workload_bytes():
    """ToyCtf (12,10), MARLon pair, dense int8 masks: 12 675 B per env-step, 89 % of it the connect mask."""
    ab = _bytes("toyctf")
    assert ab == dict(attacker_obs=716, masks=11268, defender_obs=143, state_rw=448, actions_rewards_flags=100, total=12675)
    # the masks at their reference sizes: local N*L, remote N*N*R, connect N*N*P*C bytes
    assert ab["masks"] == 12 * 3 + 12 * 12 * 8 + 12 * 12 * 7 * 10


def test_chain100_workload_bytes():
    """config 4: Chain-100 (102,102), factored masks (the owned-node bitset replaces 8.5 MB of dense masks per env)."""
    ab = _bytes("chain100")
    assert ab["total"] == 11896 and ab["masks"] == 16 and ab["attacker_obs"] == 7044 and ab["defender_obs"] == 1528
    assert ab["state_rw"] == 3208 and ab["actions_rewards_flags"] == 100


def test_totals_are_sums_and_state_is_read_and_written_once():
    for w in ("toyctf", "chain100", "toyctf_scan"):
        comp, cfg = bench.workload_config(workload=w)
        S = bench._packed_state_words(comp, cfg)
        ab = bench.algorithmic_bytes_per_env_step(comp, cfg, S)
        assert ab["total"] == sum(v for k, v in ab.items() if k != "total"), w
        assert ab["state_rw"] == 2 * 4 * S, w


def test_in_place_kernel_counts_only_what_a_step_moves():
    """cbx_wide_kernel (Chain-100): the numerator is LOWER than SURVEY 8(d)'s -- state words a step does not touch and the
    defender's static rows (written at reset only) are not counted."""
    comp, cfg = bench.workload_config(workload="chain100")
    ab = bench.in_place_kernel_bytes(comp, cfg, _bytes("chain100"))
    assert ab["survey_8d_total"] == 11896
    assert ab["defender_obs"] == 102 and ab["attacker_obs"] == 7044 and ab["masks"] == 16
    assert ab["state_rw"] == 4 * (2 * (13 + 26 + 2) + (4 + 7 + 52 + 102 + 51)) == 1192
    assert ab["total"] == 7044 + 16 + 102 + 1192 + 100 == 8454
