"""ctypes binding of the C ABI declared in include/wurm_b200.h.

There is NO fallback: if the CUDA library has not been built this module raises, and so does every
env constructor.  (The CPU implementation of this path is the reference itself.)
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# WURM_B200_LIB: a tuning variant of the library built by scripts/build_variant.sh (A/B experiments only)
LIB_PATH = os.environ.get('WURM_B200_LIB') or os.path.join(_HERE, '_C', 'libwurm_b200.so')

ABI_VERSION = 6

OK, E_INVALID, E_UNSUPPORTED, E_CUDA = 0, 1, 2, 3
ST_MULTI_HEAD, ST_NO_HEAD_PARTIAL, ST_NO_SPAWN, ST_OVERLAP, ST_NOT_COMPACT = 1, 2, 4, 8, 16
STATS_SLOTS, STATS_FIELDS = 32, 5
STAT_NAMES = ['env_steps', 'episodes', 'reward', 'self_collisions', 'edge_collisions']
PACKED_DONE, PACKED_SELF, PACKED_EDGE, PACKED_REWARD_SHIFT = 1, 2, 4, 3
OBS_NONE, OBS_DEFAULT, OBS_RAW, OBS_ONE_CHANNEL, OBS_POSITIONS, OBS_PARTIAL = -1, 0, 1, 2, 3, 4

# every symbol include/wurm_b200.h declares (tests/test_abi.py checks header and library agree)
SYMBOLS = ['wurm_abi_version', 'wurm_last_error', 'wurm_single_obs_elems', 'wurm_single_step', 'wurm_single_step_reset',
           'wurm_single_reset', 'wurm_single_compact_step', 'wurm_single_compact_reset', 'wurm_single_compact_observe',
           'wurm_single_compact_check', 'wurm_single_compact', 'wurm_single_expand',
           'wurm_single_observe', 'wurm_multi_obs_elems', 'wurm_multi_step', 'wurm_multi_step_reset', 'wurm_multi_reset', 'wurm_multi_observe',
           'wurm_multi_env_images', 'wurm_multi_compact', 'wurm_multi_expand', 'wurm_single_check', 'wurm_multi_check', 'wurm_grid_step', 'wurm_grid_reset',
           'wurm_grid_observe', 'wurm_a2c_returns']
CHECK_REPORT = 4
# WURM_CHK_* bits in the order the reference tests them, with the reference's messages
CHECK_MESSAGES = [
    (1, 'An environment has an invalid food pixel'),
    (2, 'An environment has multiple num_heads for a single snake.'),
    (4, "environments don't contain a snake."),
    (8, "An environment has a snake with it's head not at the end of the body."),
    (16, 'An environment has a body with inconsistent values i.e. not range(n)'),
    (32, 'A snake has size of less than 3.'),
    (64, 'A food and head pixel is overlapping'),
    (128, "An environment doesn't contain exactly one food instance"),
    (256, 'An environment contains overlapping snakes'),
    (512, 'Dead snake contains non-zero elements.'),
]
MULTI_MAX_SNAKES = 32
MOBS_NONE, MOBS_FULL, MOBS_PARTIAL = -1, 0, 1


class WurmSingleCfg(ctypes.Structure):
    _fields_ = [('num_envs', ctypes.c_int32), ('size', ctypes.c_int32), ('obs_mode', ctypes.c_int32),
                ('obs_n', ctypes.c_int32)]


class WurmGridCfg(ctypes.Structure):
    _fields_ = [('num_envs', ctypes.c_int32), ('size', ctypes.c_int32), ('obs_mode', ctypes.c_int32),
                ('start_y', ctypes.c_int32), ('start_x', ctypes.c_int32)]


class WurmMultiCfg(ctypes.Structure):
    _fields_ = [('num_envs', ctypes.c_int32), ('num_snakes', ctypes.c_int32), ('size', ctypes.c_int32),
                ('obs_mode', ctypes.c_int32), ('obs_n', ctypes.c_int32), ('boost', ctypes.c_int32),
                ('food_on_death', ctypes.c_int32), ('death_threshold', ctypes.c_float),
                ('boost_cost_prob', ctypes.c_float), ('food_mode', ctypes.c_int32), ('food_rate', ctypes.c_float),
                ('reward_on_death', ctypes.c_float), ('respawn_any', ctypes.c_int32), ('colour_random', ctypes.c_int32)]


class WurmMultiState(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('foods', 'heads', 'bodies', 'dones', 'orientations', 'boost_this_step',
                                               'agent_colours', 'head_hints', 'cells')] + [('cells_valid', ctypes.c_int32)]


class WurmMultiStepDraws(ctypes.Structure):
    _fields_ = [('boost_phase_ran', ctypes.c_int32)] + [(n, ctypes.c_void_p) for n in ('u_boost', 'u_cost', 'u_reg',
                                                                                        'food_cell', 'u_rate')]


class WurmMultiStepOut(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('rewards', 'snake_collision', 'edge_collision', 'food', 'size', 'dones',
                                               'boost', 'all_done', 'obs')]


class WurmMultiResetDraws(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('create', 'respawn', 'colours')]


class WurmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f'wurm_b200 error {code}: {message}')
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f'{LIB_PATH} is missing: build it with `python -m wurm_b200.build` (needs nvcc). '
                          'wurm_b200 has no CPU or PyTorch fallback.')
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
    cfg = ctypes.POINTER(WurmSingleCfg)
    L.wurm_abi_version.restype = i32
    L.wurm_last_error.restype = ctypes.c_char_p
    L.wurm_single_obs_elems.restype = ctypes.c_int64
    L.wurm_single_obs_elems.argtypes = [cfg]
    L.wurm_single_step.restype = i32
    L.wurm_single_step.argtypes = [cfg, vp, vp, i32, vp, u64, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.wurm_single_step_reset.restype = i32
    L.wurm_single_step_reset.argtypes = [cfg, vp, vp, i32, vp, vp, u64, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.wurm_single_reset.restype = i32
    L.wurm_single_reset.argtypes = [cfg, vp, vp, vp, u64, u64, vp, vp, vp]
    L.wurm_single_observe.restype = i32
    L.wurm_single_observe.argtypes = [cfg, vp, vp, vp, vp]
    L.wurm_single_compact_step.restype = i32
    L.wurm_single_compact_step.argtypes = [cfg, vp, vp, vp, i32, vp, i32, vp, u64, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.wurm_single_compact_reset.restype = i32
    L.wurm_single_compact_reset.argtypes = [cfg, vp, vp, vp, vp, u64, u64, vp, vp]
    L.wurm_single_compact_observe.restype = i32
    L.wurm_single_compact_observe.argtypes = [cfg, vp, vp, vp, vp, vp]
    L.wurm_single_compact_check.restype = i32
    L.wurm_single_compact_check.argtypes = [cfg, vp, vp, vp, vp]
    L.wurm_single_compact.restype = i32
    L.wurm_single_compact.argtypes = [cfg, vp, vp, vp, vp, vp]
    L.wurm_single_expand.restype = i32
    L.wurm_single_expand.argtypes = [cfg, vp, vp, vp, vp]
    mcfg, mst = ctypes.POINTER(WurmMultiCfg), ctypes.POINTER(WurmMultiState)
    L.wurm_multi_obs_elems.restype = ctypes.c_int64
    L.wurm_multi_obs_elems.argtypes = [mcfg]
    L.wurm_multi_step.restype = i32
    L.wurm_multi_step.argtypes = [mcfg, mst, ctypes.POINTER(vp), i32, ctypes.POINTER(WurmMultiStepDraws), u64, u64, vp,
                                  ctypes.POINTER(WurmMultiStepOut), vp, vp, vp]
    L.wurm_multi_step_reset.restype = i32
    L.wurm_multi_step_reset.argtypes = [mcfg, mst, ctypes.POINTER(vp), i32, ctypes.POINTER(WurmMultiStepDraws),
                                        ctypes.POINTER(WurmMultiResetDraws), u64, u64, vp, ctypes.POINTER(WurmMultiStepOut),
                                        vp, vp, vp]
    L.wurm_multi_reset.restype = i32
    L.wurm_multi_reset.argtypes = [mcfg, mst, vp, ctypes.POINTER(WurmMultiResetDraws), u64, u64, vp, vp, vp]
    L.wurm_multi_observe.restype = i32
    L.wurm_multi_observe.argtypes = [mcfg, mst, vp, vp, vp]
    L.wurm_multi_env_images.restype = i32
    L.wurm_multi_env_images.argtypes = [mcfg, mst, vp, vp, vp]
    L.wurm_multi_compact.restype = i32
    L.wurm_multi_compact.argtypes = [mcfg, mst, vp, vp]
    L.wurm_multi_expand.restype = i32
    L.wurm_multi_expand.argtypes = [mcfg, mst, vp]
    gcfg = ctypes.POINTER(WurmGridCfg)
    L.wurm_grid_step.restype = i32
    L.wurm_grid_step.argtypes = [gcfg, vp, vp, i32, vp, u64, u64, vp, vp, vp, vp, vp, vp, vp]
    L.wurm_grid_reset.restype = i32
    L.wurm_grid_reset.argtypes = [gcfg, vp, vp, vp, u64, u64, vp, vp]
    L.wurm_grid_observe.restype = i32
    L.wurm_grid_observe.argtypes = [gcfg, vp, vp, vp]
    L.wurm_a2c_returns.restype = i32
    L.wurm_a2c_returns.argtypes = [i32, i32, ctypes.c_double, ctypes.c_double, vp, vp, vp, vp, vp, vp]
    L.wurm_single_check.restype = i32
    L.wurm_single_check.argtypes = [cfg, vp, vp, vp, vp]
    L.wurm_multi_check.restype = i32
    L.wurm_multi_check.argtypes = [mcfg, mst, vp, vp]
    if L.wurm_abi_version() != ABI_VERSION:
        raise ImportError(f'{LIB_PATH} has ABI version {L.wurm_abi_version()}, expected {ABI_VERSION}: rebuild it')
    _lib = L
    return L


class _NoGuard(object):
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(dev):
    """`with device_guard(dev):` makes `dev` the current CUDA device for a launch, like `torch.cuda.device(dev)`, but costs
    nothing when it already is (the common case: a few microseconds matter at launch-bound batch sizes)."""
    import torch
    if dev.index is None or torch.cuda.current_device() == dev.index:
        return _NO_GUARD
    return torch.cuda.device(dev)


def check(rc):
    if rc != OK:
        raise WurmError(rc, lib().wurm_last_error().decode())


def raise_on_report(report, only=None):
    """report: the 3 ints a check kernel produced.  Raises the reference's RuntimeError for the first failing
    check (in the reference's order); `only` restricts to a subset of bits."""
    bits, count, first = int(report[0]), int(report[1]), int(report[2])
    if only is not None:
        bits &= only
    for bit, message in CHECK_MESSAGES:
        if bits & bit:
            raise RuntimeError(f'{message} ({count} violating, first at index {first})')
