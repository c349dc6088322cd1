"""Helpers of the env contract (reference wurm/utils.py:24-178): channel accessors, orientation
recovery and the invariant checks the reference's drivers and tests call.  Plain torch, off the hot
path (the step kernels derive orientations themselves); SURVEY.md section 8(f) lists a fused checker
as the next row."""
import torch

from .config import FOOD_CHANNEL, HEAD_CHANNEL, BODY_CHANNEL


def food(envs: torch.Tensor) -> torch.Tensor:
    return envs[:, FOOD_CHANNEL:FOOD_CHANNEL + 1]


def head(envs: torch.Tensor) -> torch.Tensor:
    return envs[:, HEAD_CHANNEL:HEAD_CHANNEL + 1]


def body(envs: torch.Tensor) -> torch.Tensor:
    return envs[:, BODY_CHANNEL:BODY_CHANNEL + 1]


def determine_orientations(envs: torch.Tensor) -> torch.Tensor:
    """Orientation {0,1,2,3} of the snake of each env: k such that head = neck + OFF[k],
    OFF = [(-1,0),(0,1),(1,0),(0,-1)]; 0 when head and neck are not adjacent.  Equals the reference's
    filter-response argmax (wurm/utils.py:36-65) on snakes whose two largest body values are unique."""
    n, _, S, _ = envs.shape
    flat = envs[:, BODY_CHANNEL].reshape(n, -1)
    sizes, head_idx = flat.max(dim=1)
    neck_idx = (flat == (sizes - 1).unsqueeze(1)).float().argmax(dim=1)
    dy = torch.div(head_idx, S, rounding_mode='floor') - torch.div(neck_idx, S, rounding_mode='floor')
    dx = head_idx % S - neck_idx % S
    out = torch.zeros(n, dtype=torch.long, device=envs.device)
    out[(dy == 0) & (dx == 1)] = 1
    out[(dy == 1) & (dx == 0)] = 2
    out[(dy == 0) & (dx == -1)] = 3
    return out


# (orientation, body cells tail..head, food cell) of the reference's four hand-made single-snake fixtures
_TEST_ENVS = {
    'up': ([(3, 3), (3, 4), (4, 4), (5, 4)], (6, 6)),
    'right': ([(3, 3), (3, 4), (4, 4), (4, 5)], (6, 9)),
    'down': ([(8, 8), (7, 8), (6, 8), (5, 8)], (7, 2)),
    'left': ([(8, 7), (7, 7), (6, 7), (6, 6)], (1, 2)),
}


def get_test_env(size: int, orientation: str = 'up') -> torch.Tensor:
    """The reference's predetermined single-snake environments (wurm/utils.py:68-110): a CPU (1,3,size,size) tensor with
    a length-4 snake and one food cell, which its tests move to the device and assign to `env.envs`."""
    if orientation not in _TEST_ENVS:
        raise Exception
    cells, food_cell = _TEST_ENVS[orientation]
    env = torch.zeros((1, 3, size, size))
    for value, (y, x) in enumerate(cells, start=1):
        env[0, BODY_CHANNEL, y, x] = value
    env[0, HEAD_CHANNEL, cells[-1][0], cells[-1][1]] = 1
    env[0, FOOD_CHANNEL, food_cell[0], food_cell[1]] = 1
    return env


def _fused_check(envs: torch.Tensor, skip=None):
    """One launch of wurm_single_check on a contiguous fp32 CUDA batch; returns the 3-int report (one sync)."""
    import ctypes
    from . import _lib
    n, _, S, _ = envs.shape
    report = torch.tensor([0, 0, 2 ** 31 - 1, 0], dtype=torch.int32, device=envs.device)
    cfg = _lib.WurmSingleCfg(n, S, _lib.OBS_NONE, 0)
    if skip is not None:
        skip = skip.reshape(-1)
        if skip.dtype != torch.bool or skip.device != envs.device or not skip.is_contiguous():
            skip = (skip != 0).to(envs.device).contiguous()
    with torch.cuda.device(envs.device):
        _lib.check(_lib.lib().wurm_single_check(ctypes.byref(cfg), envs.data_ptr(), None if skip is None else skip.data_ptr(),
                                                report.data_ptr(),
                                                ctypes.c_void_p(torch.cuda.current_stream(envs.device).cuda_stream)))
    return report.tolist()


def _is_fusable(envs):
    return envs.is_cuda and envs.dtype == torch.float32 and envs.dim() == 4 and envs.shape[1] == 3 and envs.is_contiguous()


def snake_consistency(envs: torch.Tensor):
    """Invariants of a 3-channel single-snake view (reference wurm/utils.py:113-164).  Contiguous fp32 CUDA
    batches go through the fused checker kernel (one launch, one sync); anything else through torch ops."""
    n = envs.shape[0]
    if n == 0:
        return
    if _is_fusable(envs):
        from . import _lib
        _lib.raise_on_report(_fused_check(envs), only=127)
        return
    f, h, b = food(envs), head(envs), body(envs)
    if not torch.all((f == 0) | (f == 1)):
        raise RuntimeError('An environment has an invalid food pixel')
    if not torch.all(h.reshape(n, -1).sum(dim=-1) == 1):
        raise RuntimeError('An environment has multiple num_heads for a single snake.')
    totals = b.reshape(n, -1).sum(dim=-1)
    if not torch.all(totals > 0):
        raise RuntimeError(f'{(totals <= 0).sum()} environments don\'t contain a snake.')
    sizes = b.reshape(n, -1).max(dim=1)[0]
    if not torch.equal(sizes, (h * b).reshape(n, -1).sum(dim=-1)):
        raise RuntimeError('An environment has a snake with it\'s head not at the end of the body.')
    if not torch.equal((torch.sqrt(8 * totals + 1) - 1) / 2, sizes):
        raise RuntimeError('An environment has a body with inconsistent values i.e. not range(n)')
    if not torch.all(totals >= 6):
        raise RuntimeError('A snake has size of less than 3.')
    overlap = (h * f).reshape(n, -1).sum(dim=-1)
    if not torch.all(overlap == 0):
        raise RuntimeError(f'A food and head pixel is overlapping in {int(overlap.sum().item())} env(s).')


def env_consistency(envs: torch.Tensor):
    """snake_consistency plus exactly one food per env (reference wurm/utils.py:167-178)."""
    n = envs.shape[0]
    if n == 0:
        return
    if _is_fusable(envs):
        from . import _lib
        _lib.raise_on_report(_fused_check(envs))
        return
    snake_consistency(envs)
    if not torch.all(food(envs).reshape(n, -1).sum(dim=-1) == 1):
        raise RuntimeError('An environment doesn\'t contain exactly one food instance')
