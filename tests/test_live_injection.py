"""The harness that injects the live state of a running env into the oracle
(tests/_parity.py::_oracle_from_live, used by the steady-state GPU parity tests) is itself
checked on the CPU: an oracle env stepped k steps, its state extracted the way the GPU tests
extract it, injected into a fresh oracle env -- both must continue identically."""
import numpy as np

from gym_auv_b200 import lidar_config, scenarios as S
from tests._parity import _oracle_from_live, oracle_cfg


def test_injected_oracle_continues_identically():
    from oracle.sim import OracleEnv

    cfg = lidar_config()
    scn = S.moving_obstacles(3, 6, 5, seed=17)
    rs = np.random.RandomState(2)
    for m in range(3):  # crowd the start so that the nearby list is not empty
        for j in range(6):
            ang, dist = rs.uniform(0, 2 * np.pi), rs.uniform(30, 140)
            scn.mov_start[m, j] = scn.vessel_init[m, :2] + dist * np.array([np.cos(ang), np.sin(ang)])
    acts = rs.uniform([-1, -0.15], [1, 0.15], size=(60, 2))
    for m in range(3):
        a = OracleEnv(scn.describe(m), oracle_cfg(cfg), test_mode=False)
        for t in range(31):  # past one nearby refresh, not on one
            a.step(acts[t])
        km, ks = scn.k_moving, scn.k_static
        pos, disp, cnt = np.zeros((km, 2)), np.zeros((km, 2)), np.zeros(km)
        mask = np.zeros(2, dtype=np.uint32)
        for j, ob in enumerate(a.obstacles):
            if not ob.static:
                pos[j], disp[j], cnt[j] = ob.position, (ob.dx, ob.dy), ob.counter
            if any(ob is nb for nb in a.vessel.nearby):
                mask[j >> 5] |= np.uint32(1 << (j & 31))
        b = _oracle_from_live(scn, cfg, m, a.vessel.state, a.vessel.step_counter, a.vessel.max_progress, a.t_step,
                              a.cumulative_reward, mask, pos, disp, cnt)
        assert len(b.vessel.nearby) == len(a.vessel.nearby)
        for t in range(31, 60):
            oa, ra, da, _ = a.step(acts[t])
            ob_, rb, db, _ = b.step(acts[t])
            assert np.array_equal(oa, ob_) and ra == rb and da == db, (m, t)
            if da:
                break
