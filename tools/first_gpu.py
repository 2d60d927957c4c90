import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
from gym_auv_b200 import scenarios as S, lidar_config
from tests._parity import rollout_oracle, rollout_gpu, compare
print(torch.cuda.get_device_name(0))
cfg = lidar_config()
scn = S.moving_obstacles(6, 17, 11, seed=3)
T = 40
rng = np.random.RandomState(7)
actions = rng.uniform([-1, -0.15], [1, 0.15], size=(T, scn.n_scenarios, 2))
t = time.time(); ref = rollout_oracle(scn, cfg, actions); print('oracle s', time.time() - t)
gpu, env = rollout_gpu(scn, cfg, actions)
print(json.dumps(compare(ref, gpu, cfg, 'moving'), indent=1))
print('seg_tests gpu', gpu['seg_tests'], 'oracle', ref['n_tests'].sum())
