"""TEST INFRASTRUCTURE ONLY -- turns the draws recorded from the reference into replay tapes.

The reference's random decisions are recorded by
oracle/reference_loader.py as (kind, call-site line, tensor) tuples.  The functions here convert
one step's / one reset's worth of those records into the dense per-env arrays that both the oracle
and the CUDA library accept as "injected draws" (include/wurm_b200.h).  Container-only: used by
gen_golden.py and validate_vs_reference.py.
"""
import numpy as np


def _rows_to_cells(rows, env_ids, N, S):
    """drop_duplicates rows (rank, 0, y, x) -> (N,) int32 cell index, -1 where no row."""
    cells = np.full(N, -1, np.int32)
    rows = rows.numpy()
    for rank, _, y, x in rows:
        cells[env_ids[rank]] = y * S + x
    return cells


def single_step_tape(tape, reward, N, S):
    """Tape of one SingleSnake.step -> food_cell (N,) int32 (single_snake.py:277-282)."""
    ate = np.flatnonzero(np.asarray(reward).reshape(-1) != 0)
    recs = [t for t in tape if t[0] == 'drop_duplicates']
    assert len(recs) <= 1 and all(t[0] == 'drop_duplicates' for t in tape)
    if not recs:
        assert len(ate) == 0
        return np.full(N, -1, np.int32)
    return _rows_to_cells(recs[0][2], ate, N, S)


def single_reset_tape(tape, done, N, S):
    """Tape of one SingleSnake.reset / __init__ -> spawn (N,4) int32 rows (y, x, dir, food_cell)
    (single_snake.py:358,359,366,384)."""
    ids = np.flatnonzero(np.asarray(done).reshape(-1))
    spawn = np.full((N, 4), -1, np.int32)
    if len(ids) == 0:
        assert not tape
        return spawn
    kinds = [(t[0], t[1]) for t in tape]
    assert kinds == [('randint', 358), ('randint', 359), ('randint', 366), ('drop_duplicates', 315)], kinds
    spawn[ids, 0] = tape[0][2].numpy()
    spawn[ids, 1] = tape[1][2].numpy()
    spawn[ids, 2] = tape[2][2].numpy()
    spawn[:, 3] = _rows_to_cells(tape[3][2], ids, N, S)
    return spawn
