"""gym_mapf_b200 -- B200-native joint-transition engine behind gym-mapf's Python API.

Drop-in surface (same names and argument meaning as LevyvoNet/gym-mapf 0.4.6):
    gym_mapf_b200.envs                  constants + integer encodings     (reference gym_mapf/envs/__init__.py)
    gym_mapf_b200.envs.grid             MapfGrid, EmptyCell, ObstacleCell (reference gym_mapf/envs/grid.py)
    gym_mapf_b200.envs.mapf_env         MapfEnv, OptimizationCriteria ... (reference gym_mapf/envs/mapf_env.py)
    gym_mapf_b200.envs.utils            create_mapf_env, parsers ...      (reference gym_mapf/envs/utils.py)
New batched surface:
    gym_mapf_b200.envs.vec_env          VecMapfEnv (batched step / rollout / transition-table construction)

`MapfEnv.P`, `MapfEnv.step` and everything in `VecMapfEnv` run in the sm_100a CUDA kernels of
`csrc/libmapf_b200.so` (C ABI: include/mapf_b200.h).  There is no CPU fallback: without the built library or without
a CUDA device those calls raise.
"""
__version__ = "0.1.0"
