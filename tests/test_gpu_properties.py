"""Size-independent properties at BASELINE.json's full batch size (65536 envs), where the
Python oracle cannot follow: determinism, batch-composition independence, bounds,
agreement of the reference-culled LiDAR with an exhaustive (exact-cull) cast, and the
oracle on a random sample of the big batch."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gym_auv_b200 import lidar_config, scenarios as S  # noqa: E402

pytestmark = pytest.mark.gpu
N_FULL = 65536


@pytest.fixture(scope="module")
def big(built_lib):
    from gym_auv_b200.vec_env import AUVVecEnv

    assert torch.cuda.is_available()
    cfg = lidar_config()
    scn = S.moving_obstacles(N_FULL, 16, 16, seed=0, n_paths=1024)
    gen = torch.Generator(device="cuda").manual_seed(7)
    lo = torch.tensor([-1.0, -0.15], device="cuda")
    hi = torch.tensor([1.0, 0.15], device="cuda")
    acts = [lo + (hi - lo) * torch.rand((N_FULL, 2), device="cuda", generator=gen) for _ in range(30)]
    env = AUVVecEnv(scn, N_FULL, cfg, test_mode=True, auto_reset=False, debug=True)
    env.reset()
    hist = []
    for a in acts:
        obs, rew, done, info = env.step(a)
        hist.append((obs.clone(), rew.clone(), done.clone(), env.get_attr("lidar_dist").clone()))
    return cfg, scn, acts, env, hist


def test_full_size_bounds(big):
    cfg, scn, acts, env, hist = big
    for obs, rew, done, dist in hist:
        assert torch.isfinite(obs).all() and obs.min() >= -1 and obs.max() <= 1
        assert torch.isfinite(rew).all()
        assert dist.min() >= 0 and dist.max() <= 150.0
        # closeness is a monotone function of the range: 0 at 150 m, 1 at 0 m
        cl = obs[:, 6:]
        assert (cl[dist == 150.0] == 0).all() and (dist[cl > 0] < 150.0).all()
    # collisions imply the fixed collision reward and done
    obs, rew, done, dist = hist[-1]
    coll = env._out["collision"].bool()
    assert (rew[coll] == -5000.0).all() and done[coll].all()
    assert ((dist.min(dim=1).values < 1.255) == coll).all()


def test_full_size_determinism_and_batch_independence(big):
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg, scn, acts, env, hist = big
    # same inputs, fresh env object: bit-identical outputs
    env2 = AUVVecEnv(scn, N_FULL, cfg, test_mode=True, auto_reset=False, debug=True)
    env2.reset()
    for t in range(10):
        obs, rew, done, _ = env2.step(acts[t])
        assert torch.equal(obs, hist[t][0]) and torch.equal(rew, hist[t][1]) and torch.equal(done, hist[t][2])
    # an env's trajectory does not depend on which batch it is in
    idx = np.array([0, 1, 31, 32, 33, 4095, 30000, 65535])
    sub = S.ScenarioSet(
        waypoints=scn.waypoints, path_id=scn.path_id[idx], vessel_init=scn.vessel_init[idx],
        mov_start=scn.mov_start[idx], mov_width=scn.mov_width[idx], mov_track=scn.mov_track[idx],
        vel_table=scn.vel_table, st_pos=scn.st_pos[idx], st_radius=scn.st_radius[idx], rewarder=scn.rewarder,
        post_generate_update=scn.post_generate_update)
    sub._bank = scn.bank
    env3 = AUVVecEnv(sub, len(idx), cfg, test_mode=True, auto_reset=False, debug=True)
    env3.reset()
    tidx = torch.as_tensor(idx, device="cuda")
    for t in range(10):
        obs, rew, done, _ = env3.step(acts[t][tidx])
        assert torch.equal(obs, hist[t][0][tidx]) and torch.equal(rew, hist[t][1][tidx])


def test_reference_culling_never_sees_more_than_exhaustive_cast(big):
    """With cull_mode='exact' every obstacle in range is tested by every ray of its angular
    window (no seam bug), so ranges can only be <= the reference-mode ranges; they are
    equal wherever the seam bug did not hide an obstacle."""
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg, scn, acts, env, hist = big
    env_ex = AUVVecEnv(scn, N_FULL, cfg, test_mode=True, auto_reset=False, debug=True, cull_mode="exact")
    env_ex.reset()
    n_equal = n_total = 0
    for t in range(6):
        env_ex.step(acts[t])
        d_ex = env_ex.get_attr("lidar_dist")
        d_ref = hist[t][3]
        assert (d_ex <= d_ref + 1e-4).all()
        n_equal += int((torch.abs(d_ex - d_ref) <= 1e-4).sum())
        n_total += d_ref.numel()
    # measured: ~1.5 % of all ray readings of this workload are affected by the seam bug
    assert n_equal / n_total > 0.97


def test_oracle_on_a_sample_of_the_full_batch(big):
    from tests._parity import compare, rollout_oracle

    cfg, scn, acts, env, hist = big
    # pick envs that have something in sensor range during the horizon
    seen = (hist[5][3].min(dim=1).values < 150).nonzero().flatten()[:5].cpu().numpy()
    idx = np.concatenate([seen, [17, 40000]]).astype(int)
    sub = S.ScenarioSet(
        waypoints=scn.waypoints, path_id=scn.path_id[idx], vessel_init=scn.vessel_init[idx],
        mov_start=scn.mov_start[idx], mov_width=scn.mov_width[idx], mov_track=scn.mov_track[idx],
        vel_table=scn.vel_table, st_pos=scn.st_pos[idx], st_radius=scn.st_radius[idx], rewarder=scn.rewarder,
        post_generate_update=scn.post_generate_update)
    sub._bank = scn.bank
    T = 12
    a = np.stack([acts[t][torch.as_tensor(idx, device="cuda")].cpu().numpy().astype(np.float64) for t in range(T)])
    ref = rollout_oracle(sub, cfg, a)
    tid = torch.as_tensor(idx, device="cuda")
    gpu = dict(
        obs=np.stack([hist[t][0][tid].cpu().numpy() for t in range(T)]),
        reward=np.stack([hist[t][1][tid].cpu().numpy() for t in range(T)]),
        done=np.stack([hist[t][2][tid].cpu().numpy().astype(bool) for t in range(T)]),
        dists=np.stack([hist[t][3][tid].cpu().numpy() for t in range(T)]),
    )
    alive = ref["alive"]
    err = np.abs(gpu["dists"] - ref["dists"])[alive]
    assert (err <= 1e-4 + 1e-4 * ref["dists"][alive]).all()
    assert np.abs(gpu["obs"] - ref["obs"])[alive].max() <= 1e-4
    assert np.abs(gpu["reward"] - ref["reward"])[alive].max() <= 1e-4 + 1e-4 * np.abs(ref["reward"][alive]).max()
    assert np.array_equal(gpu["done"][alive], ref["done"][alive])
