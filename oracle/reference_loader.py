"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference (oscarknagg/wurm) for pinning the oracle.

Used by `oracle/gen_golden.py` and `oracle/validate_vs_reference.py` (pinning the oracle), by
`bench.py --impl reference` / its `cpu_baseline` and `reference_cuda` legs (timing the reference's own
PyTorch implementation beside the CUDA path) and by `tests/test_reference_suite_gpu.py`; nothing in
`wurm_b200/` imports it.  The reference tree is looked for at `$WURM_REFERENCE_PATH`, then
`/root/reference` (the build container's read-only copy), then `baseline/_ref/` in this repo -- a
git-ignored verbatim copy that `__graft_entry__.build()` makes whenever `/root/reference` is present, so
that it travels to the GPU box with the snapshot (the reference has no setup.py / pyproject, so
`pip install --target baseline/_ref` is not possible: the tree is copied instead).

The reference was written for torch 1.1 / python 3.6.  It is imported *unmodified*; what is patched
is the interpreter around it, restoring the semantics the reference was written against
(SURVEY.md Appendix B.1):

  1. `gym` / `matplotlib` are absent          -> empty stub modules (rendering is never exercised)
  2. `collections.Iterable` moved             -> alias to `collections.abc.Iterable`
  3. `config.DEFAULT_DEVICE == 'cuda'`        -> patched to 'cpu' *before* `wurm.*` binds ctor defaults
  4. uint8 masks: `~mask` was logical-not     -> `torch.uint8 = torch.bool`, `Tensor.byte() -> .bool()`
  5. integer `/` was floor/trunc division     -> `Tensor.__truediv__` truncates for integer tensors

Because (4) and (5) monkey-patch torch process-wide, import this module only from a dedicated
process (the two scripts above do that).

Draw recording: the names `drop_duplicates` and `torch` in the namespaces of
`wurm.envs.single_snake` / `wurm.envs.multi_snake` are wrapped so that every random draw the
reference makes is appended to `TAPE` tagged with its call-site line number.
"""
import collections
import collections.abc
import inspect
import os
import sys
import types

import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIPPED_PATH = os.path.join(_REPO, 'baseline', '_ref')


def find_reference():
    """Path of the reference tree, or None."""
    for cand in (os.environ.get('WURM_REFERENCE_PATH'), '/root/reference', SHIPPED_PATH):
        if cand and os.path.isfile(os.path.join(cand, 'wurm', 'envs', 'single_snake.py')):
            return cand
    return None


REFERENCE_PATH = find_reference() or '/root/reference'


def ship(dest=SHIPPED_PATH, source='/root/reference'):
    """Copies the reference's Python tree (wurm/, tests/, experiments/, config.py) verbatim into the git-ignored
    baseline/_ref/ so that it exists on the GPU box.  No-op when the source is absent (the GPU box) or unchanged."""
    import filecmp
    import shutil
    if not os.path.isdir(source):
        return os.path.isdir(dest)
    for name in ('wurm', 'tests', 'experiments'):
        src, dst = os.path.join(source, name), os.path.join(dest, name)
        if os.path.isdir(dst):
            cmp = filecmp.dircmp(src, dst)
            if not (cmp.left_only or cmp.right_only or cmp.diff_files):
                continue
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    os.makedirs(dest, exist_ok=True)
    for name in ('config.py', 'README.md', 'requirements.txt'):
        if os.path.isfile(os.path.join(source, name)):
            shutil.copyfile(os.path.join(source, name), os.path.join(dest, name))
    return True

TAPE = []          # list of (kind, lineno, payload) appended by the recording proxies
_loaded = {}


def _stub(name):
    mod = types.ModuleType(name)
    mod.__path__ = []
    sys.modules[name] = mod
    return mod


def _install_shims():
    # (1) absent packages
    for name in ['gym', 'gym.envs', 'gym.envs.classic_control', 'gym.envs.classic_control.rendering',
                 'gym.wrappers', 'gym.wrappers.monitoring', 'gym.wrappers.monitoring.video_recorder',
                 'matplotlib', 'matplotlib.pyplot']:
        if name not in sys.modules:
            _stub(name)
    sys.modules['gym.envs.classic_control'].rendering = sys.modules['gym.envs.classic_control.rendering']
    # MultiSnake.render constructs a viewer even for mode='rgb_array' (multi_snake.py:229-231): an inert one
    sys.modules['gym.envs.classic_control.rendering'].SimpleImageViewer = type(
        'SimpleImageViewer', (), {'imshow': lambda self, img: None, 'isopen': True})
    sys.modules['gym.wrappers.monitoring.video_recorder'].VideoRecorder = type('VideoRecorder', (), {})
    # (2)
    if not hasattr(collections, 'Iterable'):
        collections.Iterable = collections.abc.Iterable
    # (4) torch-1.1 mask semantics
    torch.uint8 = torch.bool
    torch.Tensor.byte = lambda t: t.bool()
    # (5) torch-1.1 integer division
    _truediv = torch.Tensor.__truediv__

    def _int_truediv(self, other):
        if not self.is_floating_point() and isinstance(other, int):
            return torch.div(self, other, rounding_mode='trunc')
        return _truediv(self, other)
    torch.Tensor.__truediv__ = _int_truediv


class _TorchProxy(object):
    """Stands in for the name `torch` inside a reference module; records random draws."""

    def __init__(self, tag):
        self._tag = tag

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def _line():
        return inspect.stack()[2].lineno

    def rand(self, *a, **k):
        out = torch.rand(*a, **k)
        TAPE.append(('rand', self._line(), out.clone()))
        return out

    def rand_like(self, *a, **k):
        out = torch.rand_like(*a, **k)
        TAPE.append(('rand_like', self._line(), out.clone()))
        return out

    def randint(self, *a, **k):
        out = torch.randint(*a, **k)
        TAPE.append(('randint', self._line(), out.clone()))
        return out


def load(record=True, device='cpu'):
    """Returns a namespace with the reference's SingleSnake, MultiSnake, utils and config modules.
    `record=False` leaves the reference's random draws unrecorded (the timing runs: the recording proxies clone
    every draw and walk the interpreter stack).  `device`: what config.DEFAULT_DEVICE is patched to."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not os.path.isdir(REFERENCE_PATH):
        raise RuntimeError(f'reference tree not found at {REFERENCE_PATH} (nor under $WURM_REFERENCE_PATH or baseline/_ref)')
    _install_shims()
    sys.path.insert(0, REFERENCE_PATH)
    import config as ref_config                      # (3) patch before wurm.* is imported
    assert os.path.realpath(ref_config.__file__).startswith(os.path.realpath(REFERENCE_PATH))
    ref_config.DEFAULT_DEVICE = device
    import wurm.utils as ref_utils
    import wurm.envs.single_snake as ref_single
    import wurm.envs.multi_snake as ref_multi
    import wurm.envs.simple_gridworld as ref_grid

    for mod in (ref_single, ref_multi, ref_grid) if record else ():
        orig = mod.drop_duplicates

        def recording_drop_duplicates(tensor, column, random=True, _orig=orig):
            out = _orig(tensor, column, random)
            TAPE.append(('drop_duplicates', inspect.stack()[1].lineno, out.clone()))
            return out
        mod.drop_duplicates = recording_drop_duplicates
        mod.torch = _TorchProxy(mod.__name__)

    _loaded.update(dict(config=ref_config, utils=ref_utils, single=ref_single, multi=ref_multi,
                        SingleSnake=ref_single.SingleSnake, MultiSnake=ref_multi.MultiSnake,
                        SimpleGridworld=ref_grid.SimpleGridworld))
    return types.SimpleNamespace(**_loaded)


def take_tape():
    """Returns and clears the draws recorded since the last call."""
    out = list(TAPE)
    del TAPE[:]
    return out


def instrument_multi(env):
    """Records, at the moment the reference's MultiSnake._add_food runs (multi_snake.py:368-410), which
    envs it selects -- the row order of the draws made there (drop_duplicates :447 / rand :401)."""
    original = env._add_food

    def recording_add_food():
        total = env.foods.view(env.num_envs, -1).sum(dim=-1)
        selected = total < (1e-6 if env.food_mode == 'only_one' else env.max_food)
        TAPE.append(('add_food_selected', 0, selected.clone()))
        return original()
    env._add_food = recording_add_food
    return env
