"""Generate golden vectors from the REFERENCE'S OWN code (run in the build container,
where /root/reference is mounted; the fixtures it writes are committed because the
reference does not exist on the GPU box).

Only the four NumPy-only reference files can be executed here:
  gym_auv/utils/constants.py, gym_auv/utils/geomutils.py,
  gym_auv/objects/vessel/odesolver.py, gym_auv/utils/sector_partitioning.py
They are loaded by file path; ``_state_dot`` (vessel.py:561-570) and the action
scaling (vessel.py:572-578) are driven through them exactly as ``Vessel.step``
(vessel.py:226-247) does.

Usage:  python tests/golden/make_reference_goldens.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/gym_auv"
OUT = os.path.dirname(os.path.abspath(__file__))


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    const = _load("ref_constants", "utils/constants.py")
    geom = _load("ref_geomutils", "utils/geomutils.py")
    ode = _load("ref_odesolver", "objects/vessel/odesolver.py")
    sect = _load("ref_sector", "utils/sector_partitioning.py")

    def rollout(init, actions, dt):
        state = np.hstack([np.array(init, dtype=np.float64), np.zeros(3)])
        out = [state.copy()]
        for a in actions:
            tau = np.array([np.clip(a[0], 0, 1) * 2.0, 0, np.clip(a[1], -1, 1) * 0.15])

            def f(s):
                nu = s[3:]
                eta_dot = geom.Rz(geom.princip(s[2])).dot(nu)
                nu_dot = const.M_inv.dot(tau - const.D.dot(nu) - const.N(nu).dot(nu))
                return np.concatenate([eta_dot, nu_dot])

            _, q = ode.odesolver45(f, state, dt)
            state = q
            state[2] = geom.princip(state[2])
            out.append(state.copy())
        return np.array(out)

    rng = np.random.RandomState(20240)
    inits, acts, dts, trajs = [], [], [], []
    for case in range(16):
        init = np.array([rng.uniform(-500, 500), rng.uniform(-500, 500), rng.uniform(-np.pi, np.pi)])
        T = 64
        a = rng.uniform([-1, -0.15], [1, 0.15], size=(T, 2))
        if case % 4 == 3:  # out-of-box actions exercise the clipping
            a = rng.uniform(-2, 2, size=(T, 2))
        a = a.astype(np.float32).astype(np.float64)  # the C ABI takes float32 actions (gym Box dtype)
        dt = [1.0, 0.5, 0.1, 0.2][case % 4]
        inits.append(init)
        acts.append(a)
        dts.append(dt)
        trajs.append(rollout(init, a, dt))
    # the survey's probe case: from rest, a=[0.5,0.6], psi0=0.3, 5 steps
    probe = {h: rollout([0, 0, 0.3], np.tile([0.5, 0.6], (5, 1)), h)[-1] for h in (1.0, 0.5)}

    princip_in = np.concatenate([rng.uniform(-20, 20, 200), [np.pi, -np.pi, 3 * np.pi, 0.0, -3 * np.pi]])
    princip_out = np.array([geom.princip(x) for x in princip_in])

    env = types.SimpleNamespace(
        config=types.SimpleNamespace(vessel=types.SimpleNamespace(n_sensors_per_sector=20, n_sectors=9))
    )
    sectors_180 = np.array([sect.sector_partition_fun(env, i) for i in range(180)], dtype=np.int32)
    env2 = types.SimpleNamespace(
        config=types.SimpleNamespace(vessel=types.SimpleNamespace(n_sensors_per_sector=16, n_sectors=8))
    )
    sectors_128 = np.array([sect.sector_partition_fun(env2, i) for i in range(128)], dtype=np.int32)

    np.savez_compressed(
        os.path.join(OUT, "reference_dynamics.npz"),
        inits=np.array(inits),
        actions=np.array(acts),
        dts=np.array(dts),
        trajs=np.array(trajs),
        probe_h1=probe[1.0],
        probe_h05=probe[0.5],
        M_inv=const.M_inv,
        D=const.D,
        N_u1=const.N(np.array([1.0, 0.0, 0.0])),
        princip_in=princip_in,
        princip_out=princip_out,
        sectors_180=sectors_180,
        sectors_128=sectors_128,
    )
    print("wrote", os.path.join(OUT, "reference_dynamics.npz"))
    print("probe h=1:", probe[1.0])


if __name__ == "__main__":
    sys.exit(main())
