"""CUDA engine (through the C-ABI) against the golden reference trajectories and against the C oracle."""
import os

import numpy as np
import pytest

import canon
import oracle_py as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(f for f in os.listdir(GOLDEN) if f.startswith("ep_") and f.endswith(".npz"))


def _engine(*a, **k):
    from cage_challenge_4_b200 import BatchedCC4
    return BatchedCC4(*a, **k)


def _diff(exp, got):
    return "; ".join(canon.diff(canon.decode(exp), canon.decode(got))[:8])


def _red_of(f):
    fx = np.load(os.path.join(GOLDEN, f))
    return str(fx["red"]) if "red" in fx else "fsm"


STEPS = sorted({(int(np.load(os.path.join(GOLDEN, f))["steps"]), _red_of(f)) for f in FILES})


@pytest.mark.parametrize("steps,red", STEPS)
def test_tape_replay_of_reference_episodes(steps, red):
    """All golden episodes of one episode length and red agent class (FiniteStateRedAgent / DiscoveryFSRed) in ONE batch
    (ragged tick counts): state, observations, reward, done after every tick."""
    names = [f for f in FILES if int(np.load(os.path.join(GOLDEN, f))["steps"]) == steps and _red_of(f) == red]
    fxs = [np.load(os.path.join(GOLDEN, f)) for f in names]
    n = len(fxs)
    T = max(len(fx["actions"]) for fx in fxs)
    recs_all, segs = [], []
    base = 0
    n_segs = 1 + T * O.N_ACTORS
    for fx in fxs:
        recs, seg = O.tape_from_fixture(fx)
        full = np.full(n_segs + 1, seg[-1], dtype=np.uint32)
        full[:len(seg)] = seg
        segs.append(full + np.uint32(base))
        recs_all.append(recs)
        base += len(recs)
    env = _engine(n, device=0, steps=steps, rng="tape", red_agent=red)
    env.set_tape(np.concatenate(recs_all), np.stack(segs))
    obs, mask = env.reset()
    obs, mask = obs.cpu().numpy(), mask.cpu().numpy()
    for i, fx in enumerate(fxs):
        assert np.array_equal(obs[i], fx["obs"][0]), names[i]
        assert np.array_equal(mask[i].astype(np.uint8), fx["mask"]), names[i]
        off = fx["canon_off"]
        exp, got = fx["canon"][off[0]:off[1]], env.export_canon(i)
        assert np.array_equal(exp, got), f"{names[i]} reset: {_diff(exp, got)}"
    assert not env.error_flags().any()
    alive = np.ones(n, dtype=bool)
    for t in range(T):
        acts = np.zeros((n, 5), dtype=np.int32)
        msgs = np.zeros((n, 5), dtype=np.uint8)
        for i, fx in enumerate(fxs):
            alive[i] = t < len(fx["actions"])
            if alive[i]:
                acts[i], msgs[i] = fx["actions"][t], fx["msgs"][t]
            else:
                acts[i] = (49, 49, 49, 49, 145)
        o, r, d = env.step(torch.from_numpy(acts).cuda(), torch.from_numpy(msgs).cuda())
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        flags = env.error_flags() if (t % 25 == 0 or t == T - 1) else None
        for i, fx in enumerate(fxs):
            if not alive[i]:
                continue
            assert np.array_equal(o[i], fx["obs"][t + 1]), f"{names[i]} tick {t}: observation"
            assert r[i] == fx["reward"][t], f"{names[i]} tick {t}: reward {r[i]} != {fx['reward'][t]}"
            assert d[i] == fx["done"][t], f"{names[i]} tick {t}: done"
            if flags is not None:
                assert flags[i] == 0, f"{names[i]} tick {t}: error flags {flags[i]}"
            if t % 10 == 0 or t == len(fx["actions"]) - 1:
                off = fx["canon_off"]
                exp, got = fx["canon"][off[t + 1]:off[t + 2]], env.export_canon(i)
                assert np.array_equal(exp, got), f"{names[i]} tick {t}: {_diff(exp, got)}"


PX_FILES = sorted(f for f in os.listdir(GOLDEN) if f.startswith("px_") and f.endswith(".npz"))


@pytest.mark.parametrize("name", PX_FILES)
def test_production_mode_equals_reference_on_the_philox_seam(name):
    """PRODUCTION-mode parity against the LIVE reference (not only against the oracle): px_*.npz are reference episodes
    whose random source served the engine's own Philox words (ref_harness.PhiloxSeamGenerator).  The CUDA engine in
    rng="philox" mode with the same (seed, episode) must reproduce them: observations, reward and done after every tick,
    canonical state every fifth tick.  The episode sits at batch index 1 of 3 where its id allows, so the keying by
    global episode id is exercised too."""
    fx = np.load(os.path.join(GOLDEN, name))
    seed, episode, steps, red = int(fx["seed"]), int(fx["episode"]), int(fx["steps"]), str(fx["red"])
    at = 1 if episode >= 1 else 0
    env = _engine(3, device=0, seed=seed, steps=steps, first_episode=episode - at, red_agent=red)
    obs, mask = env.reset()
    off = fx["canon_off"]
    assert np.array_equal(obs[at].cpu().numpy(), fx["obs"][0]) and np.array_equal(mask[at].cpu().numpy().astype(np.uint8), fx["mask"])
    exp, got = fx["canon"][off[0]:off[1]], env.export_canon(at)
    assert np.array_equal(exp, got), f"reset: {_diff(exp, got)}"
    T = len(fx["actions"])
    for t in range(T):
        acts = np.tile(np.array([49, 49, 49, 49, 145], dtype=np.int32), (3, 1))
        msgs = np.zeros((3, 5), dtype=np.uint8)
        acts[at], msgs[at] = fx["actions"][t], fx["msgs"][t]
        o, r, d = env.step(torch.from_numpy(acts).cuda(), torch.from_numpy(msgs).cuda())
        assert np.array_equal(o[at].cpu().numpy(), fx["obs"][t + 1]), f"tick {t}: observation"
        assert float(r[at]) == fx["reward"][t] and int(d[at]) == fx["done"][t], f"tick {t}: reward / done"
        if t % 5 == 0 or t == T - 1:
            exp, got = fx["canon"][off[t + 1]:off[t + 2]], env.export_canon(at)
            assert np.array_equal(exp, got), f"tick {t}: {_diff(exp, got)}"
    assert env.error_flags()[at] == 0


@pytest.mark.parametrize("n,ticks,seed", [(256, 120, 11), (37, 499, 5)])
def test_philox_mode_matches_oracle(n, ticks, seed):
    """Same seeded inputs through the CUDA path and the C oracle: bit-exact obs / reward / done every tick and
    identical canonical state at checkpoints (ragged batch size on purpose)."""
    env = _engine(n, device=0, seed=seed, steps=500, first_episode=100)
    obs, mask = env.reset()
    refs = [O.OracleEnv(seed=seed, episode=100 + i, steps=500) for i in range(n)]
    for i, r in enumerate(refs):
        r.reset()
    obs = obs.cpu().numpy()
    mask = mask.cpu().numpy()
    for i, r in enumerate(refs):
        assert np.array_equal(obs[i], r.observe()) and np.array_equal(mask[i].astype(np.uint8), r.mask())
        assert np.array_equal(env.export_canon(i), r.export()), f"reset {i}: {_diff(r.export(), env.export_canon(i))}"
    g = np.random.default_rng(seed)
    for t in range(ticks):
        acts = np.stack([g.integers(0, k, n) for k in (82, 82, 82, 82, 242)], 1).astype(np.int32)
        msgs = g.integers(0, 256, (n, 5)).astype(np.uint8)
        o, r, d = env.step(torch.from_numpy(acts).cuda(), torch.from_numpy(msgs).cuda())
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        for i, ref in enumerate(refs):
            rc, oo, rr, dd = ref.step(acts[i], msgs[i])
            assert np.array_equal(o[i], oo), f"episode {i} tick {t}: observation"
            assert r[i] == rr and d[i] == dd, f"episode {i} tick {t}: reward/done {r[i]} {rr}"
        if t % 20 == 0 or t == ticks - 1:
            for i in range(0, n, 7):
                exp, got = refs[i].export(), env.export_canon(i)
                assert np.array_equal(exp, got), f"episode {i} tick {t}: {_diff(exp, got)}"
    flags = env.error_flags()
    oflags = np.array([r.error()[0] for r in refs])
    assert np.array_equal(flags & ~np.uint32(4), oflags & ~4), (flags, oflags)   # capacity limits differ by design


def test_repeated_decoys_on_one_host_grow_the_process_list():
    """The reference lets a blue agent deploy a decoy on the same host every second tick (VsftpdDecoyFactory is always
    compatible, DecoyVsftpd.py:17-20): ~250 processes on one host by the end of an episode.  The engine's process arena
    must follow the oracle exactly through that growth (constant-action policies do this, e.g. an argmax of a
    random-init MLP), with and without other agents' random actions interleaved."""
    n, seed, ticks = 12, 17, 499
    env = _engine(n, device=0, seed=seed, steps=500)
    env.reset()
    refs = [O.OracleEnv(seed=seed, episode=i, steps=500) for i in range(n)]
    for r in refs:
        r.reset()
    g = np.random.default_rng(seed)
    decoy0 = np.array([66, 66, 66, 66, 194], dtype=np.int32)      # DeployDecoy on server_host_0 of the agent's first subnet
    for t in range(ticks):
        acts = np.tile(decoy0, (n, 1))
        rnd = np.stack([g.integers(0, k, n) for k in (82, 82, 82, 82, 242)], 1).astype(np.int32)
        acts[n // 2:, 1::2] = rnd[n // 2:, 1::2]                   # second half: agents 1 and 3 act randomly (Restore / Remove mixed in)
        o, r, d = env.step(torch.from_numpy(acts).cuda())
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        for i, ref in enumerate(refs):
            rc, oo, rr, dd = ref.step(acts[i])
            assert np.array_equal(o[i], oo) and r[i] == rr and d[i] == dd, f"episode {i} tick {t}"
        if t % 50 == 49 or t == ticks - 1:
            for i in range(n):
                exp, got = refs[i].export(), env.export_canon(i)
                assert np.array_equal(exp, got), f"episode {i} tick {t}: {_diff(exp, got)}"
    assert not env.error_flags().any()
    assert all(r.error()[0] == 0 for r in refs)
    c = canon.decode(env.export_canon(0))
    assert max(len(h["procs"]) for h in c["hosts"].values()) > 200


@pytest.mark.parametrize("n", [1, 5, 2049, 16389])
def test_host_buffer_step_ragged_batches(n):
    """cc4_step_host cuts the batch into pipelined chunks: sizes around the chunk boundaries must equal the device path."""
    a = _engine(n, device=0, seed=3)
    b = _engine(n, device=0, seed=3)
    a.reset(), b.reset()
    g = np.random.default_rng(1)
    for t in range(6):
        acts = np.stack([g.integers(0, k, n) for k in (82, 82, 82, 82, 242)], 1).astype(np.int32)
        msgs = g.integers(0, 256, (n, 5)).astype(np.uint8)
        o1, r1, d1 = a.step(torch.from_numpy(acts).cuda(), torch.from_numpy(msgs).cuda())
        o2, r2, d2 = b.step_host(torch.from_numpy(acts).pin_memory(), torch.from_numpy(msgs).pin_memory())
        assert torch.equal(o1.cpu(), o2) and torch.equal(r1.cpu(), r2) and torch.equal(d1.cpu(), d2)
    for i in {0, n // 2, n - 1}:
        assert np.array_equal(a.export_canon(i), b.export_canon(i))
    oh, mh = b.reset_host()
    od, md = a.reset()
    assert torch.equal(od.cpu(), oh) and torch.equal(md.cpu(), mh.bool())


def test_host_buffer_step_equals_device_step():
    n = 64
    a = _engine(n, device=0, seed=3)
    b = _engine(n, device=0, seed=3)
    a.reset(), b.reset()
    g = np.random.default_rng(0)
    for t in range(20):
        acts = np.stack([g.integers(0, k, n) for k in (82, 82, 82, 82, 242)], 1).astype(np.int32)
        o1, r1, d1 = a.step(torch.from_numpy(acts).cuda())
        o2, r2, d2 = b.step_host(torch.from_numpy(acts).pin_memory())
        assert torch.equal(o1.cpu(), o2) and torch.equal(r1.cpu(), r2) and torch.equal(d1.cpu(), d2)


def test_fused_rollout_equals_tick_by_tick_and_oracle():
    """k_rollout (state resident in shared memory across ticks, synthetic blue policy) == the oracle's
    cc4o_run_episodes on the same streams; full 499-tick episodes, round trip through done."""
    n, seed = 48, 21
    env = _engine(n, device=0, seed=seed, steps=500)
    env.reset()
    total = torch.zeros(n, dtype=torch.float32, device="cuda")
    env.rollout(200, policy_seed=seed, reward_sum=total)
    env.rollout(400, policy_seed=seed, reward_sum=total)     # stops at done (tick 499)
    torch.cuda.synchronize()
    ticks, ret = O.run_episodes(seed, 0, n, 500)
    assert ticks == n * 499
    assert abs(float(total.double().sum()) - ret) < 1e-6, (float(total.sum()), ret)
    assert not (env.error_flags() & ~np.uint32(4)).any()
    c = env.export_canon(0)
    assert c[1] == 499


def test_sharding_invariance():
    """Episodes are keyed by global id: two half-size handles reproduce one full-size handle (the multi-GPU split)."""
    n = 32
    full = _engine(n, device=0, seed=9)
    lo = _engine(n // 2, device=0, seed=9, first_episode=0)
    hi = _engine(n // 2, device=0, seed=9, first_episode=n // 2)
    for e in (full, lo, hi):
        e.reset()
        e.rollout(60, policy_seed=9)
    for i in range(n // 2):
        assert np.array_equal(full.export_canon(i), lo.export_canon(i))
        assert np.array_equal(full.export_canon(n // 2 + i), hi.export_canon(i))


# ------------------------------------------------------------------------------------------------
# BASELINE.json's full size: 65 536 episodes on one GPU (configs[2]); size-independent properties + oracle samples
# ------------------------------------------------------------------------------------------------
FULL = 65536


def test_full_batch_step_matches_oracle_on_sampled_episodes():
    """One 65 536-episode batch stepped through cc4_step; 96 episodes spread over the batch (first / last block, block
    boundaries, random) are replayed by the oracle with the same actions: obs / reward / done every tick, canonical state
    at the end.  Also: no error flags anywhere in the batch."""
    seed, ticks = 31, 40
    env = _engine(FULL, device=0, seed=seed, steps=500)
    obs, mask = env.reset()
    rng = np.random.default_rng(7)
    sample = sorted(set([0, 1, 49, 50, 51, 99, 100, FULL - 51, FULL - 50, FULL - 1]) | set(rng.integers(0, FULL, 86).tolist()))
    refs = {i: O.OracleEnv(seed=seed, episode=i, steps=500) for i in sample}
    obs_h, mask_h = obs.cpu().numpy(), mask.cpu().numpy()
    for i, r in refs.items():
        r.reset()
        assert np.array_equal(obs_h[i], r.observe()) and np.array_equal(mask_h[i].astype(np.uint8), r.mask()), f"reset {i}"
    g = torch.Generator(device="cuda").manual_seed(5)
    for t in range(ticks):
        acts = torch.stack([torch.randint(0, k, (FULL,), generator=g, device="cuda", dtype=torch.int32) for k in (82, 82, 82, 82, 242)], 1).contiguous()
        msgs = torch.randint(0, 256, (FULL, 5), generator=g, device="cuda", dtype=torch.uint8)
        o, r, d = env.step(acts, msgs)
        idx = torch.tensor(sample, device="cuda")
        oh, rh, dh = o[idx].cpu().numpy(), r[idx].cpu().numpy(), d[idx].cpu().numpy()
        ah, mh = acts[idx].cpu().numpy(), msgs[idx].cpu().numpy()
        for k, i in enumerate(sample):
            rc, oo, rr, dd = refs[i].step(ah[k], mh[k])
            assert np.array_equal(oh[k], oo) and rh[k] == rr and dh[k] == dd, f"episode {i} tick {t}"
    for i in sample[::6]:
        assert np.array_equal(env.export_canon(i), refs[i].export()), f"episode {i}: {_diff(refs[i].export(), env.export_canon(i))}"
    assert not env.error_flags().any()


def test_full_batch_rollout_is_tick_additive_and_sharding_invariant():
    """65 536 episodes: (a) one fused 30-tick rollout == thirty 1-tick rollouts (state resident across ticks vs stored
    and reloaded every tick); (b) two 32 768-episode shards keyed by global episode id == the full batch (the multi-GPU
    split); compared by per-episode reward sums, a sum over all hot records would hide compensating errors so sampled
    canonical states are compared too."""
    seed, ticks = 13, 30
    a = _engine(FULL, device=0, seed=seed)
    b = _engine(FULL, device=0, seed=seed)
    ra = torch.zeros(FULL, dtype=torch.float32, device="cuda")
    rb = torch.zeros(FULL, dtype=torch.float32, device="cuda")
    a.reset(), b.reset()
    a.rollout(ticks, policy_seed=seed, reward_sum=ra)
    for _ in range(ticks):
        b.rollout(1, policy_seed=seed, reward_sum=rb)
    torch.cuda.synchronize()
    assert torch.equal(ra, rb)
    for i in (0, 12345, FULL - 1):
        assert np.array_equal(a.export_canon(i), b.export_canon(i))
    del b
    half = FULL // 2
    rs = []
    for k in range(2):
        s = _engine(half, device=0, seed=seed, first_episode=k * half)
        r = torch.zeros(half, dtype=torch.float32, device="cuda")
        s.reset()
        s.rollout(ticks, policy_seed=seed, reward_sum=r)
        torch.cuda.synchronize()
        rs.append(r)
        assert np.array_equal(s.export_canon(77), a.export_canon(k * half + 77))
        del s
    assert torch.equal(torch.cat(rs), ra)
    assert float(ra.sum()) < 0


@pytest.mark.parametrize("n", [3, 1000])
def test_compact_observations_are_the_unpadded_vectors(n):
    """CC4_STEP_OBS_COMPACT: row = the five vectors without padding (92 x 4 + 210 = 578 bytes), device and host path."""
    from cage_challenge_4_b200.engine import split_compact, OBS_SIZES
    a, b, c = (_engine(n, device=0, seed=4) for _ in range(3))
    for e in (a, b, c):
        e.reset()
    g = np.random.default_rng(2)
    for t in range(12):
        acts = torch.from_numpy(np.stack([g.integers(0, k, n) for k in (82, 82, 82, 82, 242)], 1).astype(np.int32))
        msgs = torch.from_numpy(g.integers(0, 256, (n, 5)).astype(np.uint8))
        o1, r1, d1 = a.step(acts.cuda(), msgs.cuda())
        o2, r2, d2 = b.step_compact(acts.cuda(), msgs.cuda())
        o3, r3, d3 = c.step_host(acts.pin_memory(), msgs.pin_memory(), compact=True)
        assert o2.shape == (n, 578) and o3.shape == (n, 578)
        for k, v in enumerate(split_compact(o2.cpu())):
            assert torch.equal(v, o1[:, k, :OBS_SIZES[k]].cpu()), f"tick {t} agent {k}"
        assert torch.equal(o3, o2.cpu()) and torch.equal(r1.cpu(), r3) and torch.equal(r1, r2) and torch.equal(d1.cpu(), d3)


@pytest.mark.parametrize("n,stream", [(3, "0"), (9473, "0"), (9473, "1")])
def test_bit_row_observations_and_both_host_paths(n, stream, monkeypatch):
    """CC4_STEP_OBS_BITS: row = the compact row as bits + the mission phase; cc4_step_host moves these rows and expands them
    on the host (default) or copies byte rows (CC4_HOST_EXPAND=0): all the same bytes, padded and compact."""
    from cage_challenge_4_b200.engine import unpack_bits
    a, b, c, d = (_engine(n, device=0, seed=6) for _ in range(4))
    monkeypatch.setenv("CC4_HOST_EXPAND", "0")
    e, f = (_engine(n, device=0, seed=6) for _ in range(2))
    monkeypatch.delenv("CC4_HOST_EXPAND")
    monkeypatch.setenv("CC4_HOST_STREAM", stream)       # streaming stores of the expanded rows (the multi-rank default)
    for x in (a, b, c, d, e, f):
        x.reset()
    g = np.random.default_rng(5)
    for t in range(10):
        acts = torch.from_numpy(np.stack([g.integers(0, k, n) for k in (82, 82, 82, 82, 242)], 1).astype(np.int32))
        msgs = torch.from_numpy(g.integers(0, 256, (n, 5)).astype(np.uint8))
        o1, r1, d1 = a.step_compact(acts.cuda(), msgs.cuda())
        o2, r2, d2 = b.step_bits(acts.cuda(), msgs.cuda())
        assert o2.shape == (n, 20) and torch.equal(unpack_bits(o2), o1) and torch.equal(r1, r2) and torch.equal(d1, d2)
        o3 = c.step_host(acts.pin_memory(), msgs.pin_memory(), compact=True)[0]
        o5 = e.step_host(acts.pin_memory(), msgs.pin_memory(), compact=True)[0]
        assert torch.equal(o3, o1.cpu()) and torch.equal(o5, o1.cpu())
        o4 = d.step_host(acts.pin_memory(), msgs.pin_memory())[0]
        o6 = f.step_host(acts.pin_memory(), msgs.pin_memory())[0]
        assert torch.equal(o4, o6)
        for k, v in enumerate(split_compact_(o1.cpu())):
            assert torch.equal(v, o4[:, k, :v.shape[1]]) and not o4[:, k, v.shape[1]:].any(), f"tick {t} agent {k}"


def split_compact_(o):
    from cage_challenge_4_b200.engine import split_compact
    return split_compact(o)


def test_config2_state_transition_only_4096_envs():
    """BASELINE configs[1]: 4 096 episodes, FiniteStateRed + random blue, observation / reward outputs OFF (flags = 0):
    the state trajectory is the same as with the outputs on, and sampled episodes match the oracle."""
    n, seed, ticks = 4096, 17, 50
    a = _engine(n, device=0, seed=seed)
    b = _engine(n, device=0, seed=seed)
    a.reset(), b.reset()
    sample = [0, 1, 63, 64, 2047, 4095]
    refs = {i: O.OracleEnv(seed=seed, episode=i, steps=500) for i in sample}
    for r in refs.values():
        r.reset()
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(ticks):
        acts = torch.stack([torch.randint(0, k, (n,), generator=g, device="cuda", dtype=torch.int32) for k in (82, 82, 82, 82, 242)], 1).contiguous()
        a.step(acts, None, flags=0)             # no observation vectors, no reward
        b.step(acts)
        ah = acts[sample].cpu().numpy()
        for k, i in enumerate(sample):
            refs[i].step(ah[k])
    for i in sample:
        ca = a.export_canon(i)
        assert np.array_equal(ca, b.export_canon(i)) and np.array_equal(ca, refs[i].export()), f"episode {i}"
    assert not a.error_flags().any()


def test_discovery_red_agent_philox_mode_matches_oracle():
    """DiscoveryFSRed (FSMRedVariants.py:72-122) in production mode: CUDA engine == oracle on a seeded batch."""
    n, ticks, seed = 96, 150, 29
    env = _engine(n, device=0, seed=seed, steps=500, red_agent="discovery")
    env.reset()
    refs = [O.OracleEnv(seed=seed, episode=i, steps=500, red="discovery") for i in range(n)]
    for r in refs:
        r.reset()
    g = np.random.default_rng(seed)
    for t in range(ticks):
        acts = np.stack([g.integers(0, k, n) for k in (82, 82, 82, 82, 242)], 1).astype(np.int32)
        if t % 2:
            acts[:] = (49, 49, 49, 49, 145)          # let red run freely every other tick
        o, r, d = env.step(torch.from_numpy(acts).cuda())
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        for i, ref in enumerate(refs):
            rc, oo, rr, dd = ref.step(acts[i])
            assert np.array_equal(o[i], oo) and r[i] == rr and d[i] == dd, f"episode {i} tick {t}"
    for i in range(0, n, 7):
        assert np.array_equal(env.export_canon(i), refs[i].export()), f"episode {i}: {_diff(refs[i].export(), env.export_canon(i))}"
    assert not env.error_flags().any()


def test_random_select_red_agent_philox_mode_matches_oracle():
    """RandomSelectRedAgent + Withdraw (RandomSelectRedAgent.py:37-92, Withdraw.py:38-91) in production mode: CUDA engine ==
    oracle on a seeded batch over a whole episode, state compared at checkpoints (the tape and Philox-seam fixtures of this agent
    are replayed by the parametrised tests above)."""
    n, ticks, seed = 64, 499, 41
    env = _engine(n, device=0, seed=seed, steps=500, red_agent="random_select")
    env.reset()
    refs = [O.OracleEnv(seed=seed, episode=i, steps=500, red="random_select") for i in range(n)]
    for r in refs:
        r.reset()
    g = np.random.default_rng(seed)
    for t in range(ticks):
        acts = np.stack([g.integers(0, k, n) for k in (82, 82, 82, 82, 242)], 1).astype(np.int32)
        if t % 3:
            acts[:] = (49, 49, 49, 49, 145)          # let red keep most of what it takes, so that it has sessions to withdraw
        o, r, d = env.step(torch.from_numpy(acts).cuda())
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        for i, ref in enumerate(refs):
            rc, oo, rr, dd = ref.step(acts[i])
            assert np.array_equal(o[i], oo) and r[i] == rr and d[i] == dd, f"episode {i} tick {t}"
        if t % 50 == 0 or t == ticks - 1:
            for i in range(0, n, 5):
                assert np.array_equal(env.export_canon(i), refs[i].export()), f"episode {i} tick {t}: {_diff(refs[i].export(), env.export_canon(i))}"
    flags = env.error_flags()
    assert np.array_equal(flags & ~np.uint32(4), np.array([r.error()[0] for r in refs]) & ~4)
