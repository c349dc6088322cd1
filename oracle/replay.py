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


# ------------------------------------------------------------------------------------------------
# MultiSnake (call-site lines in wurm/envs/multi_snake.py)
# ------------------------------------------------------------------------------------------------
def multi_step_tape(tape, E, K, S):
    """Tape of one MultiSnake.step -> dict for oracle.multi_step(draws=...) plus dense per-env arrays
    for the CUDA library (`food_cell` (E,), `u_rate_dense` (E,S,S))."""
    draws = {'boost_phase_ran': False, 'u_boost': None, 'u_cost': None, 'u_reg': None, 'food_cell': None, 'u_rate': None,
             'u_rate_dense': None, 'selected': None}
    death = [t[2] for t in tape if t[0] == 'rand_like' and t[1] == 424]
    cost = [t[2] for t in tape if t[0] == 'rand' and t[1] == 579]
    rate = [t[2] for t in tape if t[0] == 'rand' and t[1] == 401]
    food = [t[2] for t in tape if t[0] == 'drop_duplicates' and t[1] == 447]
    selected = [t[2] for t in tape if t[0] == 'add_food_selected']
    assert len(selected) == 1 and len(cost) <= 1 and len(rate) <= 1 and len(food) <= 1 and len(death) <= 2
    assert len(tape) == len(death) + len(cost) + len(rate) + len(food) + 1, [(t[0], t[1]) for t in tape]
    if cost:
        draws['boost_phase_ran'] = True
        draws['u_cost'] = cost[0].numpy().reshape(E * K)
        if len(death) == 2:
            draws['u_boost'] = death[0].numpy().reshape(E, S, S)
    if death:
        draws['u_reg'] = death[-1].numpy().reshape(E, S, S)
    assert len(death) <= (2 if cost else 1)
    draws['selected'] = selected[0].numpy().astype(np.uint8)
    sel = np.flatnonzero(selected[0].numpy())
    if rate:
        rows = rate[0].numpy().reshape(-1, S, S)
        assert rows.shape[0] == len(sel)
        draws['u_rate'] = rows
        dense = np.ones((E, S, S), np.float32)
        dense[sel] = rows
        draws['u_rate_dense'] = dense
    else:
        cells = np.full(E, -1, np.int32)
        if food:
            for rank, _, y, x in food[0].numpy():
                cells[sel[rank]] = y * S + x
        draws['food_cell'] = cells
    return draws


def multi_create_tape(tape, env_ids, E, K, S):
    """Consumes the records of one _create_envs call (:996-1019) from the front of `tape`.
    Returns (create (E,K+1,2) int32, remaining tape)."""
    create = np.full((E, K + 1, 2), -1, np.int32)
    pos = 0
    for k in range(K):
        assert (tape[pos][0], tape[pos][1]) == ('drop_duplicates', 951), (tape[pos][0], tape[pos][1])
        assert (tape[pos + 1][0], tape[pos + 1][1]) == ('randint', 960)
        for rank, _, y, x in tape[pos][2].numpy():
            create[env_ids[rank], k, 0] = y * S + x
        create[env_ids, k, 1] = tape[pos + 1][2].numpy()
        pos += 2
    assert (tape[pos][0], tape[pos][1]) == ('drop_duplicates', 447)
    for rank, _, y, x in tape[pos][2].numpy():
        create[env_ids[rank], K, 0] = y * S + x
    create[env_ids, K, 1] = 0
    return create, tape[pos + 1:]


def multi_reset_tape(tape, env_done, dones_before, colours_after, respawn_any, E, K, S):
    """Tape of one MultiSnake.reset -> dict for oracle.multi_reset(draws=...) (multi_snake.py:771-831).
    dones_before: the env's `dones` before the call; colours_after: `agent_colours` after it."""
    env_done = np.asarray(env_done).reshape(-1).astype(bool)
    draws = {'create': np.full((E, K + 1, 2), -1, np.int32), 'respawn': np.full((E, 2), -1, np.int32),
             'colours': np.asarray(colours_after).astype(np.int16)}
    if env_done.any():
        draws['create'], tape = multi_create_tape(tape, np.flatnonzero(env_done), E, K, S)
    if tape and tape[0][0] == 'rand' and tape[0][1] == 164:      # colours of the still-dead agents (:802)
        tape = tape[1:]
    dones = np.asarray(dones_before).reshape(E, K).astype(bool) & ~env_done[:, None]
    if respawn_any and dones.any():
        ids = np.flatnonzero(dones.any(axis=1))
        assert (tape[0][0], tape[0][1]) == ('drop_duplicates', 870) and (tape[1][0], tape[1][1]) == ('randint', 881)
        for rank, _, y, x in tape[0][2].numpy():
            draws['respawn'][ids[rank], 0] = y * S + x
        draws['respawn'][ids, 1] = tape[1][2].numpy()
        tape = tape[2:]
    assert not tape, [(t[0], t[1]) for t in tape]
    return draws


# ------------------------------------------------------------------------------------------------
# SimpleGridworld (wurm/envs/simple_gridworld.py:218)
# ------------------------------------------------------------------------------------------------
def grid_food_tape(tape, env_mask, N, S):
    """Tape of one SimpleGridworld.step (envs that ate) or reset/_create_envs (done envs) -> food_cell (N,) int32."""
    ids = np.flatnonzero(np.asarray(env_mask).reshape(-1))
    recs = [t for t in tape if t[0] == 'drop_duplicates']
    assert len(recs) == len(tape) and len(recs) <= 1
    if not recs:
        assert len(ids) == 0
        return np.full(N, -1, np.int32)
    assert recs[0][1] == 218
    return _rows_to_cells(recs[0][2], ids, N, S)
