"""Action alphabet and the integer <-> vector encodings of gym-mapf (reference gym_mapf/envs/__init__.py).

Same public names, argument order and results as the reference; the batched, device-side versions of the
state encodings live in `VecMapfEnv.state_to_cells` / `cells_to_state` (csrc: k_decode / k_encode).
"""
import os

from .maps import MAPS_PATH, ensure_map_files

UP, RIGHT, DOWN, LEFT, STAY = "UP", "RIGHT", "DOWN", "LEFT", "STAY"
ALL_STAY_JOINT_ACTION = 0
# digit order of a joint action (reference __init__.py:26)
ACTIONS = [STAY, UP, RIGHT, DOWN, LEFT]
ACTIONS_TO_INT = {name: idx for idx, name in enumerate(ACTIONS)}
# (right-slip, left-slip) of each intended action (reference __init__.py:19-25)
POSSIBILITIES = {UP: (RIGHT, LEFT), DOWN: (LEFT, RIGHT), LEFT: (UP, DOWN), RIGHT: (DOWN, UP), STAY: (STAY, STAY)}


def map_name_to_files(map_name, scen_id):
    """Paths of `<map>.map` and `<map>-even-<scen_id>.scen` (reference __init__.py:6-10).  The shipped MovingAI
    files are materialised from the packed bundle on first use."""
    ensure_map_files(map_name)
    folder = os.path.join(MAPS_PATH, map_name)
    return (os.path.join(folder, "{}.map".format(map_name)),
            os.path.join(folder, "{}-even-{}.scen".format(map_name, scen_id)))


def integer_to_vector_multiple_numbers(x, n_options_per_element, n_elements, index_to_element):
    """Little-endian mixed-radix digits of x, element 0 least significant, each mapped through
    `index_to_element` (reference __init__.py:50-67)."""
    digits = []
    for i in range(n_elements):
        x, d = divmod(x, n_options_per_element[i])
        digits.append(index_to_element(d))
    return tuple(digits)


def vector_to_integer_multiple_numbers(v, n_options_per_element, element_to_index):
    """Inverse of `integer_to_vector_multiple_numbers` (reference __init__.py:70-79)."""
    total, weight = 0, 1
    for i, element in enumerate(v):
        if i:
            weight *= n_options_per_element[i - 1]
        total += element_to_index(element) * weight
    return total


def integer_to_vector(x, options_per_element, n_elements, index_to_element):
    return integer_to_vector_multiple_numbers(x, options_per_element, n_elements, index_to_element)


def vector_to_integer(v, options_per_element, element_to_index):
    return vector_to_integer_multiple_numbers(v, options_per_element, element_to_index)
