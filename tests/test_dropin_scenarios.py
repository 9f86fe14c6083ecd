"""The scenarios of the reference's own unit tests, run against the drop-in package (`gym_mapf_b200`) through the same
public API a user of gym-mapf calls -- `MapfEnv`, `MapfGrid`, `create_mapf_env`, `env.P[s][a]`, `env.step`, `copy(env)`.
`P` rows and steps come from the CUDA kernels; the expected values are the reference's known answers
(gym_mapf/tests/mapf_env_tests.py, utils_tests.py, action_execution_tests.py, mapf_grid_tests.py, parsers_tests.py;
each case cites the lines it restates).  Host-only cases (grid, parsers, pure helpers) also run without a GPU."""
import copy

import pytest

from gym_mapf_b200.envs import (DOWN, LEFT, RIGHT, STAY, UP, integer_to_vector, map_name_to_files, vector_to_integer)
from gym_mapf_b200.envs.grid import EmptyCell, MapfGrid, ObstacleCell
from gym_mapf_b200.envs.mapf_env import (MapfEnv, OptimizationCriteria, execute_action, integer_action_to_vector,
                                         vector_action_to_integer)
from gym_mapf_b200.envs.utils import create_mapf_env, parse_map_file, parse_scen_file

CLASH, GOAL, LIVING = -1000.0, 100.0, -1
SoC, Makespan = OptimizationCriteria.SoC, OptimizationCriteria.Makespan


def empty8():
    return MapfGrid(parse_map_file(map_name_to_files("empty-8-8", 1)[0]))


def open_grid(h, w):
    return MapfGrid(["." * w] * h)


def rounded(rows):
    return {((round(p, 2), c), s, r, d) for ((p, c), s, r, d) in rows}


# ---- host-only: grid, parsers, helpers ------------------------------------------------------------------------------
def test_grid_indexing_and_bounds():  # mapf_grid_tests.py:9-32
    g = MapfGrid(["....", "....", "....", "...."])
    assert g[0, 0] is EmptyCell and g[3, 3] is EmptyCell
    with pytest.raises(IndexError):
        g[4, 0]
    berlin = MapfGrid(parse_map_file(map_name_to_files("Berlin_1_256", 1)[0]))
    assert all(berlin[0, c] is ObstacleCell for c in range(105, 109)) and berlin[0, 104] is EmptyCell


def test_scen_parser():  # parsers_tests.py:10-15
    starts, goals = parse_scen_file(map_name_to_files("empty-8-8", 1)[1], 4)
    assert starts == ((0, 0), (5, 3), (1, 7), (0, 5)) and goals == ((1, 0), (5, 6), (6, 4), (7, 4))


def test_execute_action_moves_clamps_and_obstacles():  # action_execution_tests.py:13-56
    g = empty8()
    s = ((0, 0), (7, 7))
    assert execute_action(g, s, (RIGHT, UP)) == ((0, 1), (6, 7))
    assert execute_action(g, s, (DOWN, LEFT)) == ((1, 0), (7, 6))
    assert execute_action(g, s, (LEFT, RIGHT)) == s          # against the wall
    assert execute_action(g, s, (STAY, STAY)) == s
    walls = MapfGrid(["..@..", "..@..", ".....", "..@..", "..@.."])
    assert execute_action(walls, ((0, 1),), (RIGHT,)) == ((0, 1),)  # an obstacle: stay in place


def test_encodings_known_answers():  # utils_tests.py:37-82
    assert vector_action_to_integer((DOWN, STAY, UP)) == 28 and integer_action_to_vector(28, 3) == (DOWN, STAY, UP)
    assert vector_to_integer((10, 0), [12, 12], lambda x: x) == 10 and vector_to_integer((1, 1), [12, 12], lambda x: x) == 13
    assert integer_to_vector(143, [12, 12], 2, lambda x: x) == (11, 11)
    assert vector_to_integer((1, 2, 3), [2, 3, 4], lambda x: x) == 1 + 2 * 2 + 3 * 6   # heterogeneous radices
    for a in range(125):
        assert vector_action_to_integer(integer_action_to_vector(a, 3)) == a


def test_factory_start_states():  # utils_tests.py:14-35
    env = create_mapf_env(map_name="empty-8-8", scen_id=1, n_agents=2, fail_prob=0.2, reward_of_clash=-1000.0,
                          reward_of_goal=100.0, reward_of_living=0.0, optimization_criteria=Makespan)
    assert env.s == env.locations_to_state(((0, 0), (5, 3)))
    env = create_mapf_env(map_name="empty-48-48", scen_id=16, n_agents=2, fail_prob=0.2, reward_of_clash=-1000.0,
                          reward_of_goal=100.0, reward_of_living=0.0, optimization_criteria=Makespan)
    assert env.s == env.locations_to_state(((40, 42), (17, 2)))


# ---- the transition model: rows and steps come from the GPU ---------------------------------------------------------
@pytest.mark.gpu
def test_transition_function_empty_grid():  # mapf_env_tests.py:20-71
    env = MapfEnv(empty8(), 2, ((0, 0), (7, 7)), ((0, 2), (5, 7)), 0.2, CLASH, GOAL, LIVING, Makespan)
    st = env.locations_to_state
    a = vector_action_to_integer((RIGHT, UP))
    assert rounded(env.P[env.s][a]) == {
        ((0.64, False), st(((0, 1), (6, 7))), LIVING, False), ((0.08, False), st(((1, 0), (6, 7))), LIVING, False),
        ((0.08, False), st(((0, 0), (6, 7))), LIVING, False), ((0.08, False), st(((0, 1), (7, 7))), LIVING, False),
        ((0.08, False), st(((0, 1), (7, 6))), LIVING, False), ((0.01, False), st(((1, 0), (7, 7))), LIVING, False),
        ((0.01, False), st(((1, 0), (7, 6))), LIVING, False), ((0.01, False), st(((0, 0), (7, 7))), LIVING, False),
        ((0.01, False), st(((0, 0), (7, 6))), LIVING, False)}
    wish = st(((0, 1), (6, 7)))
    assert rounded(env.P[wish][a]) == {
        ((0.64, False), st(((0, 2), (5, 7))), LIVING + GOAL, True), ((0.08, False), st(((1, 1), (5, 7))), LIVING, False),
        ((0.08, False), st(((0, 1), (5, 7))), LIVING, False), ((0.08, False), st(((0, 2), (6, 7))), LIVING, False),
        ((0.08, False), st(((0, 2), (6, 6))), LIVING, False), ((0.01, False), st(((1, 1), (6, 7))), LIVING, False),
        ((0.01, False), st(((1, 1), (6, 6))), LIVING, False), ((0.01, False), st(((0, 1), (6, 7))), LIVING, False),
        ((0.01, False), st(((0, 1), (6, 6))), LIVING, False)}


@pytest.mark.gpu
def test_vertex_clash_is_terminal_with_negative_reward():  # mapf_env_tests.py:73-90
    env = MapfEnv(empty8(), 2, ((0, 0), (0, 2)), ((7, 7), (5, 5)), 0.2, CLASH, GOAL, LIVING, Makespan)
    rows = rounded(env.P[env.s][vector_action_to_integer((RIGHT, LEFT))])
    assert ((0.64, True), env.locations_to_state(((0, 1), (0, 1))), LIVING + CLASH, True) in rows


@pytest.mark.gpu
def test_copy_keeps_stepping():  # mapf_env_tests.py:92-105
    env = MapfEnv(open_grid(5, 4), 1, ((0, 0),), ((4, 0),), 0, CLASH, GOAL, LIVING, Makespan)
    env.step(vector_action_to_integer((RIGHT,)))
    twin = copy.copy(env)
    s, r, done, _ = twin.step(vector_action_to_integer((RIGHT,)))
    assert s == twin.locations_to_state(((0, 2),)) and r == LIVING and not done
    assert env.s == env.locations_to_state(((0, 1),))  # the original did not move


@pytest.mark.gpu
def test_step_from_terminal_state_is_a_no_op():  # mapf_env_tests.py:107-128
    env = MapfEnv(open_grid(2, 2), 1, ((0, 0),), ((1, 1),), 0, CLASH, GOAL, LIVING, Makespan)
    _, r, done, _ = env.step(vector_action_to_integer((RIGHT,)))
    assert (r, done) == (LIVING, False)
    state, r, done, _ = env.step(vector_action_to_integer((DOWN,)))
    assert (r, done) == (LIVING + GOAL, True)
    for a in ((UP,), (DOWN,)):
        s2, r2, d2, info = env.step(vector_action_to_integer(a))
        assert (s2, r2, d2) == (state, 0, True) and info == {"prob": 0}


@pytest.mark.gpu
def test_swapping_cells_is_a_collision():  # mapf_env_tests.py:130-143
    env = MapfEnv(MapfGrid([".."]), 2, ((0, 0), (0, 1)), ((0, 1), (0, 0)), 0, CLASH, GOAL, LIVING, Makespan)
    _, r, done, info = env.step(vector_action_to_integer((RIGHT, LEFT)))
    assert done and r == LIVING + CLASH and info["collision"]


@pytest.mark.gpu
def test_equal_outcomes_are_merged():  # mapf_env_tests.py:229-236
    env = MapfEnv(open_grid(2, 2), 1, ((0, 0),), ((1, 1),), 0.1, CLASH, GOAL, LIVING, Makespan)
    assert env.P[env.s][vector_action_to_integer((STAY, STAY))] == [((1, False), env.s, LIVING, False)]


@pytest.mark.gpu
@pytest.mark.parametrize("criterion,first,total", [(SoC, -3, 4 * LIVING + GOAL), (Makespan, -1, 2 * LIVING + GOAL)])
def test_three_agent_rewards(criterion, first, total):  # mapf_env_tests.py:247-330
    goals = ((0, 1), (1, 3), (1, 2))
    env = MapfEnv(open_grid(4, 4), 3, ((0, 0), (3, 3), (1, 1)), goals, 0, CLASH, GOAL, LIVING, criterion)
    _, r1, done, _ = env.step(vector_action_to_integer((RIGHT, UP, RIGHT)))
    assert r1 == first and not done
    s, r2, done, _ = env.step(vector_action_to_integer((STAY, UP, STAY)))
    assert s == env.locations_to_state(goals) and done and r1 + r2 == total


@pytest.mark.gpu
def test_soc_charges_an_agent_that_stays_off_goal():  # mapf_env_tests.py:281-303
    env = MapfEnv(open_grid(4, 4), 3, ((0, 0), (3, 3), (1, 1)), ((0, 1), (1, 3), (1, 2)), 0, CLASH, GOAL, LIVING, SoC)
    _, r, _, _ = env.step(vector_action_to_integer((RIGHT, STAY, STAY)))
    assert r == -3


@pytest.mark.gpu
@pytest.mark.parametrize("criterion", [SoC, Makespan])
def test_single_agent_rewards(criterion):  # mapf_env_tests.py:332-387
    env = MapfEnv(open_grid(5, 4), 1, ((0, 0),), ((4, 0),), 0, CLASH, GOAL, LIVING, criterion)
    down, total = vector_action_to_integer((DOWN,)), 0
    for _ in range(4):
        s, r, done, _ = env.step(down)
        total += r
    assert s == env.locations_to_state(((4, 0),)) and r == LIVING + GOAL and total == GOAL + 4 * LIVING


@pytest.mark.gpu
def test_predecessors_on_the_device_match_the_host_helper():  # mapf_env_tests.py:145-227
    import torch
    from gym_mapf_b200.envs.vec_env import VecMapfEnv
    env = MapfEnv(empty8(), 2, ((0, 0), (7, 7)), ((0, 2), (5, 7)), 0.2, CLASH, GOAL, LIVING, Makespan)
    vec = VecMapfEnv(env, 1)
    states = [env.locations_to_state(((0, 0), (7, 7))), env.locations_to_state(((3, 3), (4, 4)))]
    row_ptr, pred = vec.predecessors(vec.states_from_ints(states))
    got = vec.states_to_ints(pred)
    for i, s in enumerate(states):
        assert set(got[int(row_ptr[i]):int(row_ptr[i + 1])]) == env.predecessors(s)
    # the corner agents have 3 distinct predecessor cells each (the cell itself twice through the walls)
    assert int(row_ptr[1]) == 9 and int(row_ptr[2]) - int(row_ptr[1]) == 25
    assert isinstance(pred, torch.Tensor)
