"""GPU parity: the CUDA library (through the C ABI) against (a) the golden tapes recorded from the reference and
(b) the C oracle on seeded inputs at the configs' sizes.  Integer / byte / index work bit-exact; rewards within
1e-6 relative (BASELINE.json north_star), they are in fact exact."""
import numpy as np
import pytest

import helpers
from marlon_b200 import _abi, config, scenario, scenarios

pytestmark = pytest.mark.gpu


def _batch(comp, cfg, n):
    from marlon_b200.batch import Batch

    return Batch(comp, cfg, n)


@pytest.mark.parametrize("name", helpers.golden_tapes())
def test_cuda_matches_reference_tape(name):
    meta, z = helpers.load_tape(name)
    comp, cfg = helpers.config_from_meta(meta)
    b = _batch(comp, cfg, meta["n_tapes"])
    steps = helpers.replay(meta, z, b, b.numpy, b.export_state)
    assert steps == meta["steps"]
    b.close()


def _compare_all(b, o, step, names=None):
    for k, t in b.tensors.items():
        if names and k not in names:
            continue
        got, want = t.cpu().numpy(), o.arrays[k]  # term_* too: both sides write a terminal row only when its env finishes
        assert got.shape == want.shape, k
        if got.dtype.kind == "f":
            assert np.allclose(got, want, rtol=helpers.REWARD_RTOL, atol=helpers.REWARD_RTOL), (step, k)
        else:
            assert np.array_equal(got, want), (step, k, np.argwhere(got != want)[:4])
    assert np.array_equal(b.export_state(), o.export_state()), (step, "state")


def _run_against_oracle(comp, cfg, n, steps, seed, uniform_every=3, check_every=8, tape_rng=None, expect_order=None, i16=False,
                        reset_at=None, stats_rtol=1e-9):
    """`expect_order`: the tile order the batch must be running ("dynamic" / "static"); the ticket counter must then be back at
    zero after every launch.  `i16`: int16 device actions (bulk action loads of half the size).  `reset_at`: a masked explicit
    reset of half the envs after that step, followed by more steps."""
    import torch

    from oracle import OracleBatch

    b = _batch(comp, cfg, n)
    if expect_order:
        assert b.kernel_info()["tile_order"] == expect_order, b.kernel_info()
    o = OracleBatch(comp, cfg, n)
    b.reset()
    o.reset()
    _compare_all(b, o, -1)
    if expect_order:
        assert b.tile_counter() == (0, 0)
    rng = np.random.default_rng(seed)
    marlon = cfg.mode == _abi.MODE_MARLON
    cap = max(cfg.scan_capacity, 1)
    for s in range(steps):
        att, dfn = b.sample_actions(seed=seed)
        att = att.cpu().numpy()
        dfn = dfn.cpu().numpy() if dfn is not None else None
        if uniform_every and s % uniform_every == 0:  # sprinkle uniform (mostly invalid) actions: penalty / intercept branches
            pick = rng.random(n) < 0.5
            if marlon:
                nvec = np.array([3] + [cfg.maximum_node_count] * 9)
                lay = config.attacker_action_layout(cfg)
                for kind, (a0, a1) in lay.items():
                    dims = {0: [b.views.N, b.views.L], 1: [b.views.N, b.views.N, b.views.R],
                            2: [b.views.N, b.views.N, b.views.P, b.views.C]}[kind]
                    nvec[a0:a1] = dims
                uni = (rng.random((n, 10)) * nvec).astype(np.int32)
            else:
                kind = rng.integers(0, 3, n)
                uni = np.zeros((n, 5), dtype=np.int32)
                uni[:, 0] = kind
                uni[:, 1] = rng.integers(0, b.views.N, n)
                uni[:, 2] = np.where(kind == 0, rng.integers(0, b.views.L, n), rng.integers(0, b.views.N, n))
                uni[:, 3] = np.where(kind == 1, rng.integers(0, b.views.R, n), rng.integers(0, b.views.P, n))
                uni[:, 4] = rng.integers(0, b.views.C, n)
            att = np.where(pick[:, None], uni, att)
            if dfn is not None:
                dfn[rng.random(n) < 0.3, 0] = -1  # empty defender action
        su = du = None
        if tape_rng is not None and cfg.builtin_defender:
            su, du = tape_rng.random((n, cap)), tape_rng.random((n, cap))
        if i16 and su is None:
            b.step(torch.as_tensor(att).to(torch.int16).cuda(), None if dfn is None else torch.as_tensor(dfn).to(torch.int16).cuda())
        else:
            b.step(att, dfn, su, du)
        o.step(att, dfn, su, du)
        if expect_order:
            assert b.tile_counter() == (0, 0), s
        if reset_at is not None and s == reset_at:
            m = (np.arange(n) % 2 == 0).astype(np.uint8)
            b.reset(mask=m)
            o.reset(mask=m)
            _compare_all(b, o, s)
        if s % check_every == 0 or s == steps - 1:
            _compare_all(b, o, s)
    assert np.allclose(b.stats(), o.stats, rtol=stats_rtol, atol=1e-6), (b.stats(), o.stats)
    assert b.stats()[_abi.STAT_ENV_STEPS] == n * steps
    b.close()


def _toyctf_pair_cfg(**kw):
    args = dict(maximum_node_count=12, maximum_total_credentials=10, maximum_discoverable_credentials_per_action=5,
                throws_on_invalid_actions=False, attacker_goal=config.AttackerGoal(own_atleast=6),
                defender_constraint=config.DefenderConstraint(0.60), losing_reward=-5000.0, defender_enabled=True,
                defender_max_timesteps=2000, defender_invalid_action_reward=-1, attacker_max_timesteps=2000)
    args.update(kw)
    return config.make_config(_abi.MODE_MARLON, **args)


@pytest.mark.parametrize("n", [257, 513, 4099])
@pytest.mark.parametrize("mask_mode", [_abi.MASK_DENSE, _abi.MASK_FACTORED])
def test_pipe_kernel_dynamic_tile_order_vs_oracle(monkeypatch, n, mask_mode):
    """The pipelined kernel's DYNAMIC tile order (global ticket counter, stop markers, counter reset by the last CTA) with
    SERIALISED launches -- round 1's configuration above 113 664 envs per GPU.  Forced on here at small ragged sizes: every
    array and the state against the oracle, int16 bulk actions, a masked reset mid-run, short episodes so that terminal
    observations and auto-resets occur, and the ticket counter back at {0, 0} after every launch."""
    monkeypatch.setenv("CBX_PIPE_DYNAMIC", "1")
    monkeypatch.setenv("CBX_PIPE_OVERLAP", "0")
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = _toyctf_pair_cfg(mask_mode=mask_mode, emit_terminal_obs=True, attacker_max_timesteps=23, defender_max_timesteps=17)
    _run_against_oracle(comp, cfg, n, 90, seed=41 + n, check_every=5, expect_order="dynamic", i16=(n != 513), reset_at=31)


@pytest.mark.parametrize("mode", ["overlap-dynamic", "overlap-static", "serial-static"])
def test_pipe_kernel_launch_modes_vs_oracle(monkeypatch, mode):
    """Every combination of launch mode (overlapped: programmatic dependent launch + per-tile completion counters + publisher
    warp / serialised) and tile order the library can be put in, against the oracle (overlap-dynamic is the default)."""
    ov, order = mode.split("-")
    monkeypatch.setenv("CBX_PIPE_OVERLAP", "1" if ov == "overlap" else "0")
    monkeypatch.setenv("CBX_PIPE_DYNAMIC", "1" if order == "dynamic" else "0")
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = _toyctf_pair_cfg(emit_terminal_obs=True, attacker_max_timesteps=23, defender_max_timesteps=17)
    _run_against_oracle(comp, cfg, 4099, 60, seed=47, check_every=5, expect_order=order, i16=True, reset_at=20)


def test_pipe_kernel_dynamic_tile_order_default_threshold_vs_oracle():
    """131 072 envs on one GPU (the per-GPU share of the 1M-env / 8-GPU runs), default settings: overlapped launches, dynamic
    tile order.  Every array of every env against the oracle."""
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    _run_against_oracle(comp, _toyctf_pair_cfg(), 131072, 14, seed=43, check_every=13, expect_order="dynamic")


def test_default_launch_mode_at_bench_size(monkeypatch):
    from marlon_b200.batch import Batch

    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    b = Batch(comp, _toyctf_pair_cfg(), 65536)
    info = b.kernel_info()
    assert info["name"] == "cbx_pipe_kernel" and info["overlapped_launches"] and info["tile_order"] == "dynamic", info
    b.close()
    monkeypatch.setenv("CBX_PIPE_OVERLAP", "0")  # serialised launches: the static order below 24 tiles per CTA
    b = Batch(comp, _toyctf_pair_cfg(), 65536)
    info = b.kernel_info()
    assert not info["overlapped_launches"] and info["tile_order"] == "static", info
    b.close()


@pytest.mark.parametrize("n,steps", [(4099, 700), (65536, 150), (96, 300)])
def test_overlapped_launches_back_to_back_match_serialised(monkeypatch, n, steps):
    """The overlap protocol under real overlap: `steps` steps enqueued back to back (no host synchronisation in between, actions
    from a device-resident tape) on a batch with overlapped launches must end exactly where a batch with serialised launches
    ends -- state, every observation array, statistics -- and both must agree with the oracle.  A launch that took a tile
    before the previous launch's writes to it had landed would show up here.  4 099 envs = 129 CTAs on 148 SMs and 96 envs =
    3 CTAs: grids smaller than the machine, where MANY launches are in flight at once (each launch draws its tile tickets from
    its own counter; 700 launches also wrap the ring of those counters); 65 536 envs = the bench size."""
    import torch

    from marlon_b200.batch import Batch
    from oracle import OracleBatch

    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = _toyctf_pair_cfg(attacker_max_timesteps=29, defender_max_timesteps=31)
    monkeypatch.setenv("CBX_PIPE_OVERLAP", "0")
    serial = Batch(comp, cfg, n)
    assert not serial.kernel_info()["overlapped_launches"]
    serial.reset()
    tape = []
    for s in range(steps):  # record: valid actions for the evolving state
        att, dfn = serial.sample_actions(seed=51)
        tape.append((att.clone(), dfn.clone()))
        serial.step(att, dfn)
    monkeypatch.setenv("CBX_PIPE_OVERLAP", "1")
    ov = Batch(comp, cfg, n)
    assert ov.kernel_info()["overlapped_launches"]
    ov.reset()
    torch.cuda.synchronize()
    for att, dfn in tape:  # replay back to back
        ov.step(att, dfn)
    torch.cuda.synchronize()
    assert ov.tile_counter() == (0, 0)
    assert np.array_equal(ov.export_state(), serial.export_state())
    for k in serial.tensors:
        assert torch.equal(ov.tensors[k], serial.tensors[k]), k
    assert np.array_equal(ov.stats(), serial.stats())
    if n <= 8192:
        o = OracleBatch(comp, cfg, n)
        o.reset()
        for att, dfn in tape:
            o.step(att.cpu().numpy(), dfn.cpu().numpy())
        _compare_all(ov, o, steps)
    serial.close(); ov.close()


@pytest.mark.parametrize("mode", ["marlon", "cyberbattle"])
def test_device_sampler_equals_oracle_sampler(mode):
    """cbx_sample_kernel against the oracle's restatement of sample_valid_action (rejection sampling over whole proposals,
    swapped-kind quirk; the oracle's frequencies are pinned on the live reference by tests/test_sampler_distribution.py): the
    same actions for the same state and call number, over runs that visit states with and without cached credentials."""
    from oracle import OracleBatch

    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    if mode == "marlon":
        cfg = _toyctf_pair_cfg(attacker_max_timesteps=40, defender_max_timesteps=35)
    else:
        cfg = config.make_config(_abi.MODE_CYBERBATTLE, maximum_node_count=12, maximum_total_credentials=10,
                                 throws_on_invalid_actions=False, attacker_goal=config.AttackerGoal(own_atleast=6),
                                 defender_agent=config.ScanAndReimageCompromisedMachines(0.6, 2, 5),
                                 defender_constraint=config.DefenderConstraint(0.80), seed=5, auto_reset=True)
    n = 777
    b, o = _batch(comp, cfg, n), OracleBatch(comp, cfg, n)
    b.reset(); o.reset()
    kinds = np.zeros(3, dtype=np.int64)
    for s in range(120):
        att, dfn = b.sample_actions(seed=77)
        oa, od = o.sample_actions(seed=77)
        att = att.cpu().numpy()
        assert np.array_equal(att, oa), (s, np.argwhere(att != oa)[:4])
        if dfn is not None:
            dfn = dfn.cpu().numpy()
            assert np.array_equal(dfn, od), s
        kinds += np.bincount(np.array(cfg.kind_of_index)[att[:, 0]] if mode == "marlon" else att[:, 0], minlength=3)
        b.step(att, dfn)
        o.step(att, dfn)
    assert np.array_equal(b.export_state(), o.export_state())
    assert b.stats()[_abi.STAT_ATT_INVALID] == 0  # every sampled action passed the wrapper's range check
    assert kinds.min() > 0 and kinds[_abi.KIND_LOCAL] < kinds[_abi.KIND_REMOTE]  # local proposals are rejected more often
    b.close()


@pytest.mark.parametrize("kernel", ["cbx_pipe_kernel", "cbx_step_kernel"])
@pytest.mark.parametrize("scn", ["toyctf", "chain10"])
def test_live_defender_binding_vs_oracle(scn, kernel, monkeypatch):
    """the LIVE LearningDefender binding -- re-imaging, block_traffic and allow_traffic act on the environment
    the attacker plays in; firewall rule lists are per-env state, one per alias group.  The oracle's version (explicit rule
    lists) is pinned on tapes recorded from the reference's classes with their binding refreshed at every reset
    (tests/golden/marlon_*_live*.npz); here the CUDA version (two bits per group and port name) against it at scale -- on the
    pipelined kernel (the default: the logic thread packs the rule bits into the encoder descriptor after the defender's move,
    an encoder warp writes the tile's firewall rows) and on the fused kernel (CBX_PIPE=0: rows straight from the state tile)."""
    if kernel == "cbx_step_kernel":
        monkeypatch.setenv("CBX_PIPE", "0")
    if scn == "toyctf":
        comp = scenario.compile_scenario(scenarios.toyctf_environment())
        cfg = _toyctf_pair_cfg(defender_binding="live", attacker_max_timesteps=70, defender_max_timesteps=55, emit_terminal_obs=True)
    else:
        comp = scenario.compile_scenario(scenarios.chain_environment(10))
        cfg = config.make_config(_abi.MODE_MARLON, maximum_node_count=12, maximum_total_credentials=12, throws_on_invalid_actions=False,
                                 defender_constraint=config.DefenderConstraint(0.60), losing_reward=-5000.0, defender_enabled=True,
                                 defender_max_timesteps=90, attacker_max_timesteps=120, defender_invalid_action_reward=-1,
                                 defender_reset_on_constraint_broken=False, defender_binding="live", emit_terminal_obs=True)
    from marlon_b200.batch import Batch

    b = Batch(comp, cfg, 32)
    assert b.kernel_info()["name"] == kernel
    b.close()
    # Chain-10 here keeps playing after an SLA breach: the worsening penalty -200 * k / 12 is not a whole number, and the running
    # returns are fp32 on the device (fp64 in the oracle): per-step rewards agree to 1e-6 relative, sums of returns likewise
    _run_against_oracle(comp, cfg, 2051, 260, seed=61, check_every=7, reset_at=100, stats_rtol=1e-6)


class _TiledTape(dict):
    """A reference tape with every per-env array repeated `reps` times along the env axis, one step at a time (nothing big is
    materialised): env e of the wide batch replays tape column e % n_tapes."""

    class _Rows:
        def __init__(self, a, reps):
            self.a, self.reps = a, reps

        def __getitem__(self, s):
            row = self.a[s]
            return np.tile(row, (self.reps,) + (1,) * (row.ndim - 1))

    def __init__(self, z, reps, steps):
        super().__init__((k, (self._Rows(v, reps) if getattr(v, "ndim", 0) >= 2 and v.shape[0] == steps else v)) for k, v in z.items())

    @property
    def files(self):
        return list(self)


def test_chain10_attacker_only_4096_envs_replay_the_reference_tape():
    """BASELINE.json configs[1] to the letter: CyberBattleChain-10 attacker-only, 4096 batched envs on one GPU, bit-exact against
    the tape recorded from the unmodified reference (8 columns x 600 steps: env e replays column e % 8) -- every reward, flag,
    observation field, mask checksum, terminal observation and the canonical state, every fourth step."""
    meta, z = helpers.load_tape("marlon_chain10_attacker_only")
    comp, cfg = helpers.config_from_meta(meta)
    reps = 4096 // meta["n_tapes"]
    b = _batch(comp, cfg, 4096)
    assert b.kernel_info()["name"] == "cbx_pipe_kernel"
    steps = helpers.replay(meta, _TiledTape(z, reps, meta["steps"]), b, b.numpy, b.export_state, check_every=4)
    assert steps == meta["steps"]
    b.close()


def test_chain10_attacker_only_4096_envs_vs_oracle():
    """BASELINE.json configs[1]: CyberBattleChain-10 attacker-only, 4096 batched envs, bit-exact."""
    comp = scenario.compile_scenario(scenarios.chain_environment(10))
    cfg = config.make_config(_abi.MODE_CYBERBATTLE, maximum_node_count=12, maximum_total_credentials=12,
                             throws_on_invalid_actions=False, attacker_goal=config.AttackerGoal(own_atleast_percent=1.0),
                             auto_reset=True, emit_terminal_obs=True)
    _run_against_oracle(comp, cfg, 4096, 160, seed=7)


def test_toyctf_scan_and_reimage_philox_vs_oracle():
    """configs[2] parity leg: ToyCtf + ScanAndReimage(0.6, 2, 5), SLA 0.80, Philox draws on both sides."""
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = config.make_config(_abi.MODE_CYBERBATTLE, maximum_node_count=12, maximum_total_credentials=10,
                             throws_on_invalid_actions=False, attacker_goal=config.AttackerGoal(own_atleast=6),
                             defender_agent=config.ScanAndReimageCompromisedMachines(0.6, 2, 5),
                             defender_constraint=config.DefenderConstraint(0.80), seed=1234, auto_reset=True, emit_terminal_obs=True)
    _run_against_oracle(comp, cfg, 2048, 200, seed=11)


def test_chain10_scan_tape_draws_vs_oracle():
    comp = scenario.compile_scenario(scenarios.chain_environment(10))
    cfg = config.make_config(_abi.MODE_CYBERBATTLE, maximum_node_count=12, maximum_total_credentials=12,
                             throws_on_invalid_actions=False, defender_agent=config.ScanAndReimageCompromisedMachines(0.9, 3, 2),
                             defender_constraint=config.DefenderConstraint(0.5), auto_reset=True, emit_terminal_obs=True)
    _run_against_oracle(comp, cfg, 1000, 200, seed=13, tape_rng=np.random.default_rng(99))


def test_toyctf_marlon_pair_vs_oracle():
    """The bench workload: ToyCtf (12,10) MARLon attacker+defender pair step, dense masks."""
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = config.make_config(_abi.MODE_MARLON, maximum_node_count=12, maximum_total_credentials=10,
                             maximum_discoverable_credentials_per_action=5, throws_on_invalid_actions=False,
                             attacker_goal=config.AttackerGoal(own_atleast=6), defender_constraint=config.DefenderConstraint(0.60),
                             losing_reward=-5000.0, defender_enabled=True, defender_max_timesteps=2000,
                             defender_invalid_action_reward=-1, attacker_max_timesteps=2000, emit_terminal_obs=True)
    _run_against_oracle(comp, cfg, 4099, 300, seed=17)  # 4099: a ragged last tile


def test_toyctf_marlon_pair_full_bench_size_vs_oracle():
    """configs[2] at its full size: 65 536 envs on one GPU (what bench.py times), every array of every env against the oracle."""
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = config.make_config(_abi.MODE_MARLON, maximum_node_count=12, maximum_total_credentials=10,
                             maximum_discoverable_credentials_per_action=5, throws_on_invalid_actions=False,
                             attacker_goal=config.AttackerGoal(own_atleast=6), defender_constraint=config.DefenderConstraint(0.60),
                             losing_reward=-5000.0, defender_enabled=True, defender_max_timesteps=2000,
                             defender_invalid_action_reward=-1, attacker_max_timesteps=2000)
    _run_against_oracle(comp, cfg, 65536, 24, seed=23, check_every=12)


def test_toyctf_factored_masks_vs_oracle():
    """ToyCtf with factored masks: the pipelined kernel without any dense-mask work (encoder warps only emit the defender tile)."""
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = config.make_config(_abi.MODE_MARLON, maximum_node_count=12, maximum_total_credentials=10,
                             throws_on_invalid_actions=False, defender_constraint=config.DefenderConstraint(0.60),
                             losing_reward=-5000.0, defender_enabled=True, defender_max_timesteps=60,
                             attacker_max_timesteps=60, mask_mode=_abi.MASK_FACTORED, emit_terminal_obs=True)
    _run_against_oracle(comp, cfg, 1000, 150, seed=29)


def test_chain100_factored_masks_vs_oracle():
    """configs[3] shape: Chain-100 (102,102) attacker+defender, factored masks (dense would be 8.5 MB/env)."""
    comp = scenario.compile_scenario(scenarios.chain_environment(100))
    cfg = config.make_config(_abi.MODE_MARLON, maximum_node_count=102, maximum_total_credentials=102,
                             throws_on_invalid_actions=False, defender_constraint=config.DefenderConstraint(0.60),
                             losing_reward=-5000.0, defender_enabled=True, defender_max_timesteps=500,
                             attacker_max_timesteps=500, mask_mode=_abi.MASK_FACTORED)
    _run_against_oracle(comp, cfg, 1024, 200, seed=19)


def test_chain100_dense_small_batch_vs_oracle():
    comp = scenario.compile_scenario(scenarios.chain_environment(100))
    cfg = config.make_config(_abi.MODE_CYBERBATTLE, maximum_node_count=102, maximum_total_credentials=102,
                             throws_on_invalid_actions=False, auto_reset=True)
    _run_against_oracle(comp, cfg, 3, 40, seed=23, check_every=4)


def test_odd_bounds_take_the_generic_encoder_path():
    """Bounds whose mask sizes are not multiples of 16 bytes (ragged vectors straddle envs)."""
    comp = scenario.compile_scenario(scenarios.chain_environment(4))
    cfg = config.make_config(_abi.MODE_CYBERBATTLE, maximum_node_count=7, maximum_total_credentials=5,
                             throws_on_invalid_actions=False, auto_reset=True, emit_terminal_obs=True)
    _run_against_oracle(comp, cfg, 77, 120, seed=29, check_every=2)


def test_default_bounds_single_env():
    """CyberBattleEnv defaults N=100, C=1000 (70 MB connect mask): one env, a few steps."""
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = config.make_config(_abi.MODE_CYBERBATTLE, throws_on_invalid_actions=False, auto_reset=False)
    _run_against_oracle(comp, cfg, 1, 12, seed=31, check_every=3)


def test_step_host_roundtrip():
    """The HOST-buffer entry point (what bench.py's e2e leg times) returns the same rewards/flags as the device views."""
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = config.make_config(_abi.MODE_MARLON, maximum_node_count=12, maximum_total_credentials=10, throws_on_invalid_actions=False,
                             attacker_goal=config.AttackerGoal(own_atleast=6), defender_constraint=config.DefenderConstraint(0.60),
                             losing_reward=-5000.0, defender_enabled=True)
    b = _batch(comp, cfg, 513)
    b.reset()
    for s in range(5):
        att, dfn = b.sample_actions(seed=3)
        out = b.step_host(att.cpu().numpy(), dfn.cpu().numpy())
        for k in ("att_reward", "def_reward", "att_terminated", "att_truncated", "def_terminated", "def_truncated"):
            assert np.array_equal(out[k], b.numpy(k)), k
    b.close()


def test_step_host_page_locked_buffers_and_int16_actions():
    """Page-locked callers: the kernel reads the actions in place and writes the result block itself (no copies).  int16
    action elements give the same step as int32 ones; 513 envs = 16 full tiles (bulk action loads) + a ragged one."""
    import torch

    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = config.make_config(_abi.MODE_MARLON, maximum_node_count=12, maximum_total_credentials=10, throws_on_invalid_actions=False,
                             attacker_goal=config.AttackerGoal(own_atleast=6), defender_constraint=config.DefenderConstraint(0.60),
                             losing_reward=-5000.0, defender_enabled=True)
    n = 513
    ref, b32, b16 = _batch(comp, cfg, n), _batch(comp, cfg, n), _batch(comp, cfg, n)
    for b in (ref, b32, b16):
        b.reset()
    p_att32, p_def32 = torch.empty((n, 10), dtype=torch.int32, pin_memory=True), torch.empty((n, 12), dtype=torch.int32, pin_memory=True)
    p_att16, p_def16 = torch.empty((n, 10), dtype=torch.int16, pin_memory=True), torch.empty((n, 12), dtype=torch.int16, pin_memory=True)
    keys = ("att_reward", "def_reward", "att_terminated", "att_truncated", "def_terminated", "def_truncated")
    for s in range(40):
        att, dfn = ref.sample_actions(seed=5)
        if s % 4 == 0:
            dfn[::3, 0] = -1  # empty defender actions (negative values must survive the narrowing)
        ref.step(att, dfn)
        p_att32.copy_(att); p_def32.copy_(dfn); p_att16.copy_(att.to(torch.int16)); p_def16.copy_(dfn.to(torch.int16))
        o32 = b32.step_host(p_att32.numpy(), p_def32.numpy())
        o16 = b16.step_host(p_att16.numpy(), p_def16.numpy())
        for k in keys:
            want = ref.numpy(k)
            assert np.array_equal(o32[k], want), (s, k, "int32 page-locked")
            assert np.array_equal(o16[k], want), (s, k, "int16 page-locked")
            assert np.array_equal(b16.numpy(k), want), (s, k, "device copy of the results")
        if s % 8 == 7:
            st = ref.export_state()
            assert np.array_equal(b32.export_state(), st) and np.array_equal(b16.export_state(), st), s
            for k, t in ref.tensors.items():
                assert torch.equal(t, b16.tensors[k]), (s, k)
    for b in (ref, b32, b16):
        b.close()


@pytest.mark.parametrize("case", ["chain100_factored", "odd_bounds_cyber"])
def test_int16_device_actions_match_int32(case):
    """cbx_batch_step_i16 on the other two kernels (warp-per-tile and fused) and in CyberBattleEnv mode ([n,5] actions)."""
    import torch

    if case == "chain100_factored":
        comp = scenario.compile_scenario(scenarios.chain_environment(100))
        cfg = config.make_config(_abi.MODE_MARLON, maximum_node_count=102, maximum_total_credentials=102,
                                 throws_on_invalid_actions=False, defender_constraint=config.DefenderConstraint(0.60),
                                 losing_reward=-5000.0, defender_enabled=True, defender_max_timesteps=500,
                                 attacker_max_timesteps=500, mask_mode=_abi.MASK_FACTORED)
        n = 200
    else:
        comp = scenario.compile_scenario(scenarios.chain_environment(4))
        cfg = config.make_config(_abi.MODE_CYBERBATTLE, maximum_node_count=7, maximum_total_credentials=5,
                                 throws_on_invalid_actions=False, auto_reset=True, emit_terminal_obs=True)
        n = 77
    ref, b16 = _batch(comp, cfg, n), _batch(comp, cfg, n)
    ref.reset(); b16.reset()
    for s in range(60):
        att, dfn = ref.sample_actions(seed=9)
        ref.step(att, dfn)
        b16.step(att.to(torch.int16), dfn.to(torch.int16) if dfn is not None else None)
        if s % 6 == 5:
            assert np.array_equal(b16.export_state(), ref.export_state()), s
            for k, t in ref.tensors.items():
                assert torch.equal(t, b16.tensors[k]), (s, k)
    ref.close(); b16.close()


# ---- configs[4]: several generated CyberBattleRandom networks in ONE batch (padded layout, scenario per env group) ----
def _slice_export(x, n_max, n_k, C, nsec_k):
    """multi-scenario export (sections sized for the largest scenario) -> the layout a single-scenario export of n_k nodes has"""
    H = _abi.X_HEADER_WORDS
    parts = [x[:, :H]] + [x[:, H + sec * n_max: H + sec * n_max + n_k] for sec in range(10)]
    base = H + 10 * n_max
    parts += [x[:, base: base + C], x[:, base + C: base + C + (nsec_k + 31) // 32]]
    return np.concatenate(parts, axis=1)


@pytest.mark.parametrize("mask_mode", [_abi.MASK_FACTORED, _abi.MASK_DENSE])
def test_random_networks_multi_scenario_batch_vs_oracle(mask_mode):
    """Three generated networks (different node / credential / service counts) side by side in one batch: every group of envs
    must behave exactly like a single-scenario oracle batch of its own network."""
    from marlon_b200 import random_network
    from marlon_b200.batch import Batch
    from oracle import OracleBatch

    seeds = [0, 3, 7]
    counts = [64, 96, 45] if mask_mode == _abi.MASK_FACTORED else [32, 32, 13]
    comps = [scenario.compile_scenario(random_network.random_environment(s)) for s in seeds]
    assert len({c.n_services for c in comps}) > 1 and len({len(c.triples) for c in comps}) > 1
    cfg = config.make_config(_abi.MODE_MARLON, maximum_node_count=72, maximum_total_credentials=136,
                             maximum_discoverable_credentials_per_action=32, throws_on_invalid_actions=False,
                             defender_constraint=config.DefenderConstraint(0.60), losing_reward=-5000.0, defender_enabled=True,
                             defender_max_timesteps=200, attacker_max_timesteps=200, mask_mode=mask_mode, emit_terminal_obs=True)
    b = Batch(comps, cfg, counts)
    n_max, nsvc_max = b.views.n_nodes, b.views.n_services
    assert n_max == max(c.n_nodes for c in comps) and nsvc_max == max(c.n_services for c in comps)
    offs = np.concatenate([[0], np.cumsum(counts)])
    oracles = []
    for k, comp in enumerate(comps):
        ck = _abi.Config.from_buffer_copy(bytes(cfg))
        ck.env_index_base = int(offs[k])
        oracles.append(OracleBatch(comp, ck, counts[k]))
    b.reset()
    for o in oracles:
        o.reset()
    rng = np.random.default_rng(5)

    def compare(step):
        x = b.export_state()
        for k, (comp, o) in enumerate(zip(comps, oracles)):
            sl = slice(int(offs[k]), int(offs[k + 1]))
            n_k, nsvc_k = comp.n_nodes, comp.n_services
            for name, t in b.tensors.items():
                if name.startswith("term_"):
                    continue
                got, want = t[sl].cpu().numpy(), o.arrays[name]
                if name == "def_infected_nodes":
                    assert not got[:, n_k:].any(); got = got[:, :n_k]
                elif name in ("def_incoming_firewall", "def_outgoing_firewall"):
                    assert not got[:, 6 * n_k:].any(); got = got[:, :6 * n_k]
                elif name == "def_services_status":
                    assert not got[:, nsvc_k:].any(); got = got[:, :nsvc_k]
                assert got.shape == want.shape, (name, got.shape, want.shape)
                if got.dtype.kind == "f":
                    assert np.allclose(got, want, rtol=helpers.REWARD_RTOL, atol=helpers.REWARD_RTOL), (step, k, name)
                else:
                    assert np.array_equal(got, want), (step, k, name, np.argwhere(got != want)[:4])
            want_x = o.export_state()
            got_x = _slice_export(x[sl], n_max, n_k, cfg.maximum_total_credentials, len(comp.secrets))
            assert np.array_equal(got_x, want_x), (step, k, "state", np.argwhere(got_x != want_x)[:4])

    compare(-1)
    n = sum(counts)
    for s in range(120):
        att, dfn = b.sample_actions(seed=11)
        att, dfn = att.cpu().numpy(), dfn.cpu().numpy()
        if s % 3 == 0:  # uniform actions over the padded spaces: out-of-range nodes, invalid defender targets
            pick = rng.random(n) < 0.5
            nvec = np.array([3] + [cfg.maximum_node_count] * 9)
            for kind, (a0, a1) in config.attacker_action_layout(cfg).items():
                nvec[a0:a1] = {0: [b.views.N, b.views.L], 1: [b.views.N, b.views.N, b.views.R],
                               2: [b.views.N, b.views.N, b.views.P, b.views.C]}[kind]
            uni = (rng.random((n, 10)) * nvec).astype(np.int32)
            att = np.where(pick[:, None], uni, att)
            dfn[rng.random(n) < 0.3, 0] = -1
        for k, o in enumerate(oracles):
            sl = slice(int(offs[k]), int(offs[k + 1]))
            dk = dfn[sl].copy()
            # the sampler draws defender node coordinates below the scenario's own node count, so they are valid oracle inputs
            o.step(att[sl], dk, None, None)
        b.step(att, dfn, None, None)
        if s % 6 == 0 or s == 119:
            compare(s)
    want_stats = sum(o.stats for o in oracles)
    assert np.allclose(b.stats(), want_stats, rtol=1e-9, atol=1e-6)
    b.close()


def test_full_size_runs_are_reproducible():
    """Two independent batches at the bench size fed the same actions for 120 steps end in identical states, statistics and
    observations (a data race between the logic and encoder warps of the pipelined kernel would show up here)."""
    import torch

    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = config.make_config(_abi.MODE_MARLON, maximum_node_count=12, maximum_total_credentials=10,
                             maximum_discoverable_credentials_per_action=5, throws_on_invalid_actions=False,
                             attacker_goal=config.AttackerGoal(own_atleast=6), defender_constraint=config.DefenderConstraint(0.60),
                             losing_reward=-5000.0, defender_enabled=True, defender_max_timesteps=2000,
                             defender_invalid_action_reward=-1, attacker_max_timesteps=2000)
    n = 65536
    a, b = _batch(comp, cfg, n), _batch(comp, cfg, n)
    a.reset(); b.reset()
    for s in range(120):
        att, dfn = a.sample_actions(seed=31)
        a.step(att, dfn)
        b.step(att, dfn)
        if s % 40 == 39:
            for k in a.tensors:
                assert torch.equal(a.tensors[k], b.tensors[k]), (s, k)
    assert np.array_equal(a.export_state(), b.export_state())
    assert np.array_equal(a.stats(), b.stats()) and a.stats()[_abi.STAT_ENV_STEPS] == n * 120
    a.close(); b.close()
