"""CPU: the C-ABI library loads, exports every symbol include/wurm_b200.h declares, and rejects bad
arguments before touching the GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest

from wurm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'wurm_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(wurm_[a-z_0-9]+)\s*\(', text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in header_symbols():
        assert hasattr(L, name), f'{name} declared in include/wurm_b200.h but not exported'
    assert L.wurm_abi_version() == _lib.ABI_VERSION


def test_obs_elems():
    L = _lib.lib()
    for mode, n, expect in [(_lib.OBS_DEFAULT, 0, 3 * 81), (_lib.OBS_RAW, 0, 3 * 81), (_lib.OBS_ONE_CHANNEL, 0, 81),
                            (_lib.OBS_POSITIONS, 0, 4), (_lib.OBS_PARTIAL, 2, 75), (_lib.OBS_NONE, 0, 0)]:
        cfg = _lib.WurmSingleCfg(4, 9, mode, n)
        assert L.wurm_single_obs_elems(ctypes.byref(cfg)) == expect


@pytest.mark.parametrize('cfg,code', [
    ((0, 9, 0, 0), _lib.E_INVALID),          # no envs
    ((4, 8, 0, 0), _lib.E_INVALID),          # size <= 8 (reference single_snake.py:346)
    ((4, 9, 7, 0), _lib.E_INVALID),          # unknown observation mode
    ((4, 300, 0, 0), _lib.E_UNSUPPORTED),    # cell indices no longer fit the 16-bit index arithmetic (sides up to 255 run)
])
def test_bad_config_is_rejected_on_the_host(cfg, code):
    L = _lib.lib()
    c = _lib.WurmSingleCfg(*cfg)
    rc = L.wurm_single_reset(ctypes.byref(c), None, None, None, 0, 0, None, None, None)
    assert rc == code
    assert L.wurm_last_error()
    with pytest.raises(_lib.WurmError):
        _lib.check(rc)


def test_null_pointers_are_rejected():
    L = _lib.lib()
    c = _lib.WurmSingleCfg(4, 9, 0, 0)
    assert L.wurm_single_reset(ctypes.byref(c), None, None, None, 0, 0, None, None, None) == _lib.E_INVALID
    assert L.wurm_single_observe(ctypes.byref(c), None, None, None, None) == _lib.E_INVALID
    assert L.wurm_single_step(ctypes.byref(c), None, None, 8, None, 0, 0, None, None, None, None, None, None, None,
                              None, None, None, None) == _lib.E_INVALID


def test_env_refuses_to_run_without_cuda_device():
    """There is no CPU fallback: constructing an env on a non-CUDA device raises."""
    from wurm_b200.envs import SingleSnake
    with pytest.raises(RuntimeError):
        SingleSnake(num_envs=2, size=9, device='cpu')


def test_a2c_returns_rejects_bad_arguments():
    L = _lib.lib()
    assert L.wurm_a2c_returns(0, 4, 0.99, -1.0, None, None, None, None, None, None) == _lib.E_INVALID
    assert L.wurm_a2c_returns(5, 4, 0.99, -1.0, None, None, None, None, None, None) == _lib.E_INVALID


def test_trajectory_store_mirrors_the_reference():
    """wurm/rl/trajectory_store.py:4-89: append() keeps what it is given, properties stack to (T, N, ...), clear() empties."""
    import torch
    from wurm_b200.trajectory_store import TrajectoryStore
    store = TrajectoryStore()
    for t in range(3):
        store.append(action=torch.full((4,), t), log_prob=torch.zeros(4, 1, requires_grad=True), reward=torch.full((4, 1), float(t)),
                     value=torch.ones(4, 1), done=torch.zeros(4, 1, dtype=torch.bool), entropy=torch.tensor(0.5))
    assert store.rewards.shape == (3, 4, 1) and store.actions.shape == (3, 4) and store.entropies.shape == (3,)
    assert store.log_probs.requires_grad and store.dones.dtype == torch.bool and len(store) == 3
    assert float(store.rewards[2, 0, 0]) == 2.0
    with pytest.raises(RuntimeError):
        store.states                      # nothing appended: torch.stack of an empty list, as in the reference
    store.clear()
    assert len(store) == 0


def test_a2c_has_no_cpu_fallback():
    """The return scan runs on the device only: CPU tensors are refused (the CPU implementation is the reference itself)."""
    import torch
    from wurm_b200.rl import A2C
    T, N = 3, 4
    with pytest.raises(RuntimeError):
        A2C(gamma=0.99).loss(torch.zeros(N, 1), torch.zeros(T, N, 1), torch.zeros(T, N, 1), torch.zeros(T, N, 1),
                             torch.zeros(T, N, 1, dtype=torch.bool))


def test_ctypes_structs_mirror_the_header():
    """The binding's ctypes structures against include/wurm_b200.h itself: a C program compiled from the header prints
    sizeof and every member's offset; names, order, offsets and sizes must agree (the header is plain C by contract)."""
    import re
    import subprocess
    import tempfile
    header = os.path.join(ROOT, 'include', 'wurm_b200.h')
    text = re.sub(r'/\*.*?\*/', '', open(header).read(), flags=re.S)
    structs = {}
    for name, body in re.findall(r'typedef struct (\w+) \{(.*?)\} \1;', text, flags=re.S):
        members = []
        for decl in body.split(';'):
            decl = decl.strip()
            if decl:
                members.append(re.findall(r'(\w+)\s*(?:\[[^\]]*\])?$', decl)[0])
        structs[name] = members
    assert set(structs) == {'WurmSingleCfg', 'WurmMultiCfg', 'WurmMultiState', 'WurmMultiStepDraws', 'WurmMultiStepOut',
                            'WurmMultiResetDraws', 'WurmGridCfg'}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{header}"', 'int main(void) {']
    for name, members in structs.items():
        lines.append(f'  printf("{name} sizeof %zu\\n", sizeof({name}));')
        for m in members:
            lines.append(f'  printf("{name} {m} %zu\\n", offsetof({name}, {m}));')
    lines += ['  return 0;', '}']
    with tempfile.TemporaryDirectory() as tmp:
        src, exe = os.path.join(tmp, 'layout.c'), os.path.join(tmp, 'layout')
        open(src, 'w').write('\n'.join(lines))
        subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-o', exe, src])
        out = subprocess.check_output([exe], text=True)
    for line in out.splitlines():
        name, member, value = line.split()
        cls = getattr(_lib, name)
        if member == 'sizeof':
            assert ctypes.sizeof(cls) == int(value), f'sizeof({name})'
        else:
            assert getattr(cls, member).offset == int(value), f'offsetof({name}, {member})'
    for name, members in structs.items():
        assert [f[0] for f in getattr(_lib, name)._fields_] == members, f'{name}: member names / order'
