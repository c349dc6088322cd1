/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of oscarknagg/wurm's batched env step.
 *
 * Nothing under wurm_b200/ may include, link or call this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, as the checker
 * and as the timed CPU port -- never as the product path.
 *
 * Parity status: PINNED.  Every function here is checked against the unmodified reference
 * (imported in the build container through oracle/reference_loader.py) by
 * oracle/validate_vs_reference.py, and against the golden vectors under tests/golden/ (generated
 * from the reference by oracle/gen_golden.py) plus the known-answer scenarios of the reference's
 * own tests (tests/test_single_snake_env.py, tests/test_multi_snake_env.py in the reference tree).
 *
 * Style: one environment at a time, plain loops over the (S,S) grid on the reference's own fp32
 * state layout, following the reference's tensor formulas literally (zero-padded cross-correlation
 * with the fixed 3x3 filters becomes a neighbour lookup).  No attempt is made to be clever: this is
 * the specification the CUDA kernels are diffed against.  Citations are path:line in the reference
 * tree.
 *
 * Randomness: the reference draws from torch's CPU generator through an unstable argsort
 * (wurm/utils.py:188,224), which cannot be reproduced.  Every random decision is therefore an
 * INPUT: either replayed from a tape recorded from the reference, or derived from Philox4x32-10
 * keyed by (seed, step counter, env, stream) -- the same derivation the CUDA kernels use, restated
 * here independently.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EPS 1e-6f /* config.py:11 */

/* orientation k  <=>  head = neck + OFF[k]      (wurm/_filters.py:7-28 as cross-correlation taps) */
static const int OFF_Y[4] = {-1, 0, 1, 0};
static const int OFF_X[4] = {0, 1, 0, -1};

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11), restated from the paper.                               */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t v[4]; } philox4;

static philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    philox4 out = {{c0, c1, c2, c3}};
    return out;
}

void wurm_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    memcpy(out, r.v, sizeof(r.v));
}

/* Draw streams.  The i-th 32-bit draw of stream s for env e at call counter `step` is word (i % 4) of
 * philox(counter = (e, s | (i / 4) << 4, step_lo, step_hi), key = (seed_lo, seed_hi)). */
enum {
    STREAM_SINGLE_STEP_FOOD = 0, STREAM_SINGLE_RESET = 1,
    STREAM_MULTI_DEATH_BOOST = 2, STREAM_MULTI_BOOST_COST = 3, STREAM_MULTI_DEATH_REGULAR = 4,
    STREAM_MULTI_FOOD_ONE = 5, STREAM_MULTI_FOOD_RATE = 6, STREAM_MULTI_CREATE_SNAKE = 7,
    STREAM_MULTI_CREATE_FOOD = 8, STREAM_MULTI_RESPAWN = 9, STREAM_MULTI_COLOUR = 10
};

static philox4 draw(uint64_t seed, uint64_t step, uint32_t unit, uint32_t stream) {
    return philox4x32_10(unit, stream, (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
}

static uint32_t draw_i(uint64_t seed, uint64_t step, uint32_t unit, uint32_t stream, uint32_t i) {
    return draw(seed, step, unit, stream | ((i >> 2) << 4)).v[i & 3];
}

/* uniform integer in [0,n) by multiply-shift */
static uint32_t bounded(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

/* uniform float in [0,1) with 24 random bits (the resolution of torch.rand) */
static float unit_float(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

#define REJECTION_TRIES 32

/* ------------------------------------------------------------------------------------------ */
/* SingleSnake                                                                                  */
/* ------------------------------------------------------------------------------------------ */

/* wurm/utils.py:36-65 determine_orientations, one env.  `necks` is S*S scratch. */
static int orientation_of(int S, const float* body, float* necks) {
    int C = S * S;
    float size = body[0];
    for (int p = 1; p < C; ++p) size = body[p] > size ? body[p] : size;    /* :50 */
    float shift = size - 2.0f;                                               /* :51-52 */
    for (int p = 0; p < C; ++p) {
        float n = body[p] - shift;                                           /* :53 */
        n = n > 0.0f ? n : 0.0f;
        if (n > 0.0f) n -= 1.5f;                                             /* :54 */
        necks[p] = n * 2.0f;                                                 /* :55 */
    }
    int best_k = 0;
    float best = 0.0f;
    for (int k = 0; k < 4; ++k) {                                            /* :59 conv2d, padding=1 */
        float mk = -INFINITY;
        for (int y = 0; y < S; ++y)
            for (int x = 0; x < S; ++x) {
                int yy = y + OFF_Y[k], xx = x + OFF_X[k];
                float nb = (yy >= 0 && yy < S && xx >= 0 && xx < S) ? necks[yy * S + xx] : 0.0f;
                float r = nb - necks[y * S + x];
                mk = r > mk ? r : mk;                                        /* :63 max over cells */
            }
        if (k == 0 || mk > best) { best = mk; best_k = k; }                  /* :63 argmax, first max wins */
    }
    return best_k;
}

/* The r-th (raster order) interior cell with food+head+body < EPS; -1 if none.
 * single_snake.py:306-320 picks uniformly among those cells. */
static int count_free_single(int S, const float* env) {
    int C = S * S, n = 0;
    for (int y = 1; y < S - 1; ++y)
        for (int x = 1; x < S - 1; ++x) {
            int p = y * S + x;
            if (env[p] + env[C + p] + env[2 * C + p] < EPS) ++n;
        }
    return n;
}

static int nth_free_single(int S, const float* env, int r) {
    int C = S * S;
    for (int y = 1; y < S - 1; ++y)
        for (int x = 1; x < S - 1; ++x) {
            int p = y * S + x;
            if (env[p] + env[C + p] + env[2 * C + p] < EPS) {
                if (r == 0) return p;
                --r;
            }
        }
    return -1;
}

/* single_snake.py:197-304 for one env.  food_cell: >=0 replayed cell, -1 replay "no cell",
 * -2 derive from Philox. */
static void single_step_env(int S, float* env, int64_t* action, int food_cell, uint64_t seed, uint64_t step,
                            uint32_t e, float* reward, uint8_t* done, uint8_t* self_col, uint8_t* edge_col,
                            float* scratch) {
    int C = S * S;
    float* food = env;
    float* head = env + C;
    float* body = env + 2 * C;
    float* necks = scratch;
    float* moved = scratch + C;

    float size = body[0];                                                    /* :210 */
    for (int p = 1; p < C; ++p) size = body[p] > size ? body[p] : size;

    int k = orientation_of(S, body, necks);                                  /* :212 */
    int64_t a = *action;
    a = (a + ((int64_t)k == a ? 2 : 0)) % 4;                                 /* :221-222, written back */
    *action = a;

    /* :225-233  head += conv2d(head, ORIENTATION_FILTERS)[a]; round */
    for (int y = 0; y < S; ++y)
        for (int x = 0; x < S; ++x) {
            int p = y * S + x;
            float delta = 0.0f;
            if (a >= 0) {
                int yy = y + OFF_Y[a], xx = x + OFF_X[a];
                float nb = (yy >= 0 && yy < S && xx >= 0 && xx < S) ? head[yy * S + xx] : 0.0f;
                delta = nb - head[p];
            }
            moved[p] = nearbyintf(head[p] + delta);
        }
    memcpy(head, moved, sizeof(float) * C);

    float overlap = 0.0f;                                                    /* :242 */
    for (int p = 0; p < C; ++p) overlap += head[p] * food[p];

    if (overlap == 0.0f)                                                     /* :246-249 */
        for (int p = 0; p < C; ++p) {
            float b = body[p] - 1.0f;
            body[p] = b > 0.0f ? b : 0.0f;
        }

    float hb = 0.0f;                                                         /* :252 */
    for (int p = 0; p < C; ++p) hb += head[p] * body[p];
    int sc = hb > EPS;

    for (int p = 0; p < C; ++p) body[p] += head[p] * (size + overlap);       /* :258-262 */

    float removed = 0.0f;                                                    /* :270-272 */
    for (int p = 0; p < C; ++p) {
        float rem = head[p] * food[p] * -1.0f;
        removed += rem;
        food[p] += rem;
    }
    *reward = 0.0f - removed;

    if (removed * -1.0f != 0.0f) {                                           /* :277-282 */
        int cell = food_cell;
        if (food_cell == -2) {
            /* uniform over the free interior cells by rejection: draw interior cells until one is
             * free; after REJECTION_TRIES misses rank the free cells explicitly */
            int I = S - 2;
            cell = -1;
            for (uint32_t t = 0; t < REJECTION_TRIES && cell < 0; ++t) {
                int cand = (int)bounded(draw_i(seed, step, e, STREAM_SINGLE_STEP_FOOD, t), (uint32_t)(I * I));
                int q = (1 + cand / I) * S + 1 + cand % I;
                if (env[q] + env[C + q] + env[2 * C + q] < EPS) cell = q;
            }
            if (cell < 0) {
                int nfree = count_free_single(S, env);
                if (nfree > 0) cell = nth_free_single(S, env, (int)bounded(draw_i(seed, step, e, STREAM_SINGLE_STEP_FOOD, REJECTION_TRIES), (uint32_t)nfree));
            }
        }
        if (cell >= 0) food[cell] += 1.0f;
    }

    float interior = 0.0f;                                                   /* :290-293 */
    for (int y = 1; y < S - 1; ++y)
        for (int x = 1; x < S - 1; ++x) interior += head[y * S + x];
    int ec = interior < EPS;

    for (int p = 0; p < 3 * C; ++p) env[p] = nearbyintf(env[p]);             /* :300 */

    *self_col = (uint8_t)sc;
    *edge_col = (uint8_t)ec;
    *done = (uint8_t)(sc | ec);
}

void wurm_oracle_single_step(int N, int S, float* envs, int64_t* actions, const int32_t* food_cell_replay,
                             uint64_t seed, uint64_t step, float* reward, uint8_t* done, uint8_t* self_col,
                             uint8_t* edge_col) {
    int C = S * S;
#pragma omp parallel
    {
        float* scratch = (float*)malloc(sizeof(float) * 2 * C);
#pragma omp for schedule(static)
        for (int e = 0; e < N; ++e)
            single_step_env(S, envs + (size_t)e * 3 * C, actions + e, food_cell_replay ? food_cell_replay[e] : -2,
                            seed, step, (uint32_t)e, reward + e, done + e, self_col + e, edge_col + e, scratch);
        free(scratch);
    }
}

/* single_snake.py:344-387 for one env.  spawn = (y, x, dir, food_cell) replayed, or NULL -> Philox. */
static void single_create_env(int S, float* env, const int32_t* spawn, uint64_t seed, uint64_t step, uint32_t e) {
    int C = S * S;
    int y, x, d, cell;
    memset(env, 0, sizeof(float) * 3 * C);                                   /* :352 */
    philox4 r = draw(seed, step, e, STREAM_SINGLE_RESET);
    if (spawn) {
        y = spawn[0]; x = spawn[1]; d = spawn[2];
    } else {
        y = 4 + (int)bounded(r.v[0], (uint32_t)(S - 8));                     /* :358 randint(4, S-4) */
        x = 4 + (int)bounded(r.v[1], (uint32_t)(S - 8));                     /* :359 */
        d = (int)(r.v[2] >> 30);                                             /* :366 randint(4) */
    }
    /* :372-375 LENGTH_3_SNAKES[d] stamped at the seed: tail 1, seed 2, head 3 */
    env[2 * C + (y - OFF_Y[d]) * S + (x - OFF_X[d])] = 1.0f;
    env[2 * C + y * S + x] = 2.0f;
    env[2 * C + (y + OFF_Y[d]) * S + (x + OFF_X[d])] = 3.0f;
    env[C + (y + OFF_Y[d]) * S + (x + OFF_X[d])] = 1.0f;                     /* :379-381 head where body == max */
    if (spawn) {
        cell = spawn[3];
    } else {
        int nfree = count_free_single(S, env);                               /* :384 */
        cell = nfree > 0 ? nth_free_single(S, env, (int)bounded(r.v[3], (uint32_t)nfree)) : -1;
    }
    if (cell >= 0) env[cell] += 1.0f;                                        /* :385 */
}

/* single_snake.py:322-337 (the observation at :342 is wurm_oracle_single_observe) */
void wurm_oracle_single_reset(int N, int S, float* envs, const uint8_t* done, const int32_t* spawn_replay,
                              uint64_t seed, uint64_t step) {
    int C = S * S;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < N; ++e)
        if (done[e]) single_create_env(S, envs + (size_t)e * 3 * C, spawn_replay ? spawn_replay + 4 * (size_t)e : NULL, seed, step, (uint32_t)e);
}

/* single_snake.py:104-128 _get_rgb: int16 colour of one cell */
static void single_rgb(int S, const float* env, int y, int x, int16_t rgb[3]) {
    int C = S * S, p = y * S + x;
    rgb[0] = rgb[1] = rgb[2] = 255;                                          /* :106 */
    if (env[2 * C + p] > EPS) { rgb[0] = 0; rgb[1] = 127; rgb[2] = 0; }      /* :111-112, colour :99 */
    if (env[C + p] > EPS) { rgb[0] = 0; rgb[1] = 255; rgb[2] = 0; }          /* :114-115 */
    if (env[p] > EPS) { rgb[0] = 255; rgb[1] = 0; rgb[2] = 0; }              /* :117-118 */
    if (y == 0 || x == 0 || y == S - 1 || x == S - 1) rgb[0] = rgb[1] = rgb[2] = 0; /* :120-123 */
}

enum { OBS_DEFAULT = 0, OBS_RAW = 1, OBS_ONE_CHANNEL = 2, OBS_POSITIONS = 3, OBS_PARTIAL = 4 };

static int argmax_first(const float* v, int n) {
    int best = 0;
    for (int i = 1; i < n; ++i)
        if (v[i] > v[best]) best = i;
    return best;
}

/* single_snake.py:130-195.  Returns the number of envs whose partial window could not be formed
 * (no head cell: the reference raises a view-shape error there; the oracle writes zeros). */
int wurm_oracle_single_observe(int N, int S, const float* envs, int mode, int n, float* obs) {
    int C = S * S, W = 2 * n + 1, bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int e = 0; e < N; ++e) {
        const float* env = envs + (size_t)e * 3 * C;
        int16_t rgb[3];
        if (mode == OBS_DEFAULT) {                                           /* :131-138 */
            float* o = obs + (size_t)e * 3 * C;
            for (int y = 0; y < S; ++y)
                for (int x = 0; x < S; ++x) {
                    single_rgb(S, env, y, x, rgb);
                    for (int c = 0; c < 3; ++c) o[c * C + y * S + x] = (float)rgb[c] / 255.0f;
                }
        } else if (mode == OBS_RAW) {                                        /* :139-141 */
            memcpy(obs + (size_t)e * 3 * C, env, sizeof(float) * 3 * C);
        } else if (mode == OBS_ONE_CHANNEL) {                                /* :142-151 */
            float* o = obs + (size_t)e * C;
            for (int y = 0; y < S; ++y)
                for (int x = 0; x < S; ++x) {
                    int p = y * S + x;
                    float v = (env[2 * C + p] > EPS ? 1.0f : 0.0f) * 0.5f;
                    v += env[C + p] * 0.5f;
                    v += env[p] * 1.5f;
                    if (y == 0 || x == 0 || y == S - 1 || x == S - 1) v = -1.0f;
                    o[p] = v;
                }
        } else if (mode == OBS_POSITIONS) {                                  /* :152-165 */
            float* o = obs + (size_t)e * 4;
            int h = argmax_first(env + C, C), f = argmax_first(env, C);
            o[0] = (float)(h / S); o[1] = (float)(h % S); o[2] = (float)(f / S); o[3] = (float)(f % S);
        } else {                                                             /* :166-193 partial_n */
            float* o = obs + (size_t)e * 3 * W * W;
            int hp = -1, heads = 0;
            for (int p = 0; p < C; ++p)
                if (env[C + p] != 0.0f) { hp = p; ++heads; }
            if (heads != 1) {
                memset(o, 0, sizeof(float) * 3 * W * W);
                ++bad;
                continue;
            }
            int hy = hp / S, hx = hp % S;
            for (int i = 0; i < W; ++i)
                for (int j = 0; j < W; ++j) {
                    int y = hy - n + i, x = hx - n + j;
                    int inside = y >= 0 && y < S && x >= 0 && x < S;
                    if (inside) single_rgb(S, env, y, x, rgb);
                    for (int c = 0; c < 3; ++c) o[c * W * W + i * W + j] = inside ? (float)rgb[c] / 255.0f : 0.0f;
                }
        }
    }
    return bad;
}

/* ------------------------------------------------------------------------------------------ */
/* MultiSnake (wurm/envs/multi_snake.py)                                                         */
/* State of env e: foods[e] (S,S); heads[e*K+k], bodies[e*K+k] (S,S); dones, orientations,      */
/* boost_this_step [e*K+k]; agent_colours [e*K+k][3] int16.                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct WurmOracleMultiCfg {
    int32_t num_envs, num_snakes, size;
    int32_t boost;             /* self.boost                                  multi_snake.py:123 */
    int32_t food_on_death;     /* food_on_death_prob > 0                      :565,662 */
    float death_threshold;     /* float32(1 - food_on_death_prob)             :424 */
    float boost_cost_prob;     /* float32(boost_cost_prob)                    :579 */
    int32_t food_mode;         /* 0 'only_one', 1 'random_rate'               :369,380 */
    float food_rate;           /* float32(food_rate)                          :403 */
    float reward_on_death;     /*                                             :684 */
    int32_t respawn_any;       /* respawn_mode == 'any'                       :805 */
    int32_t colour_random;     /* colour_mode == 'random'                     :800 */
} WurmOracleMultiCfg;

/* The random draws of one MultiSnake.step.  replay == 0: Philox. */
typedef struct WurmOracleMultiStepDraws {
    int32_t replay;
    int32_t boost_phase_ran;   /* replay: the reference ran the boost phase (:503, batch-global)   */
    const float* u_boost;      /* (E,S,S)   rand_like at :424 via :574                             */
    const float* u_cost;       /* (E*K)     rand at :579                                           */
    const float* u_reg;        /* (E,S,S)   rand_like at :424 via :671                             */
    const int32_t* food_cell;  /* (E)       only_one: respawned cell or -1 (:447)                  */
    const float* u_rate;       /* (n,S,S)   random_rate: rand at :401, row = rank among selected   */
    int32_t n_rate_rows;
} WurmOracleMultiStepDraws;

/* multi_snake.py:341-353 _move_heads for one agent: head += conv2d(head, ORIENTATION_FILTERS)[dir] */
static void multi_move_head(int S, float* head, int dir, float* tmp) {
    for (int y = 0; y < S; ++y)
        for (int x = 0; x < S; ++x) {
            int yy = y + OFF_Y[dir], xx = x + OFF_X[dir], p = y * S + x;
            float nb = (yy >= 0 && yy < S && xx >= 0 && xx < S) ? head[yy * S + xx] : 0.0f;
            tmp[p] = head[p] + (nb - head[p]);
        }
    memcpy(head, tmp, sizeof(float) * S * S);
}

/* One phase of MultiSnake.step for env e on the agents flagged in `active`
 * (boost phase :509-576 with active = boosted agents; regular phase :613-673 with all agents). */
static void multi_phase(const WurmOracleMultiCfg* cfg, float* food, float* heads, float* bodies, uint8_t* done,
                        const int* mv, const uint8_t* active, float* sizes, float* rewards, float* food_cons,
                        uint8_t* scol, uint8_t* ecol, const float* U, uint32_t stream, uint64_t seed, uint64_t step,
                        uint32_t e, float* scratch) {
    int K = cfg->num_snakes, S = cfg->size, C = S * S;
    float* tmp = scratch;           /* C */
    float* total = scratch + C;     /* C */
    float* ovsum = scratch + 2 * C; /* K */
    uint8_t col[64], edge[64];

    for (int k = 0; k < K; ++k)                                              /* :509 / :613 */
        if (active[k]) multi_move_head(S, heads + (size_t)k * C, mv[k], tmp);

    /* :514-518 / :618-622 food overlap of ALL heads; food -= clamp(sum_k overlap, 0, 1) */
    for (int p = 0; p < C; ++p) total[p] = 0.0f;
    for (int k = 0; k < K; ++k) {
        ovsum[k] = 0.0f;
        for (int p = 0; p < C; ++p) {
            float o = (heads[(size_t)k * C + p] * food[p] > EPS) ? 1.0f : 0.0f;
            ovsum[k] += o;
            total[p] += o;
        }
    }
    for (int p = 0; p < C; ++p) food[p] -= total[p] < 0.0f ? 0.0f : (total[p] > 1.0f ? 1.0f : total[p]);

    /* :523-529 / :627-631 decay the active snakes that did not eat; reward the ones that did */
    for (int k = 0; k < K; ++k) {
        if (!active[k]) continue;
        if (ovsum[k] < EPS)
            for (int p = 0; p < C; ++p) {
                float b = bodies[(size_t)k * C + p] - 1.0f;
                bodies[(size_t)k * C + p] = b > 0.0f ? b : 0.0f;
            }
        float eaten = ovsum[k] > EPS ? 1.0f : 0.0f;
        rewards[k] += eaten;
        food_cons[k] += eaten;
    }

    /* :534-547 / :636-644 collisions: head of k against the heads of the others + ALL bodies */
    for (int k = 0; k < K; ++k) {
        col[k] = 0;
        if (!active[k]) continue;
        for (int p = 0; p < C; ++p) {
            float pathing = 0.0f;
            for (int j = 0; j < K; ++j)
                if (j != k) pathing += heads[(size_t)j * C + p];
            float allb = 0.0f;
            for (int j = 0; j < K; ++j) allb += bodies[(size_t)j * C + p];
            pathing += allb;
            if (heads[(size_t)k * C + p] * pathing > EPS) col[k] = 1;
        }
    }
    for (int k = 0; k < K; ++k)
        if (active[k]) { done[k] |= col[k]; scol[k] |= col[k]; }

    /* :552-555 / :649-652 grow the body at the head cell */
    for (int k = 0; k < K; ++k) {
        if (!active[k]) continue;
        float growth = ovsum[k] > EPS ? 1.0f : 0.0f;
        for (int p = 0; p < C; ++p) bodies[(size_t)k * C + p] += heads[(size_t)k * C + p] * (sizes[k] + growth);
        sizes[k] += growth;
    }

    /* :560-562 / :657-659 edge collisions (:412-414 with the mask of :155-161) */
    for (int k = 0; k < K; ++k) {
        edge[k] = 0;
        if (!active[k]) continue;
        for (int y = 0; y < S; ++y)
            for (int x = 0; x < S; ++x) {
                float m = (y == 0 || x == 0 || y == S - 1 || x == S - 1) ? 1.0f : 0.0f;
                if (heads[(size_t)k * C + y * S + x] * m > EPS) edge[k] = 1;
            }
        done[k] |= edge[k];
        ecol[k] |= edge[k];
    }

    /* :565-576 / :662-673 food where dead snakes lie (_food_from_death :416-428) */
    if (cfg->food_on_death)
        for (int y = 0; y < S; ++y)
            for (int x = 0; x < S; ++x) {
                int p = y * S + x;
                float dead = 0.0f, living = 0.0f;
                for (int k = 0; k < K; ++k) {
                    if (done[k]) dead += bodies[(size_t)k * C + p];
                    else living += bodies[(size_t)k * C + p];
                }
                if (y == 1 || x == 0 || y == S - 1 || x == S - 1) dead = 0.0f;   /* sic: row 1, not row 0 (:418) */
                dead = nearbyintf(dead) > 0.0f ? 1.0f : 0.0f;
                if (dead == 0.0f) continue;              /* 0 * u > 1-p is false for p <= 1 */
                float u = U ? U[p] : unit_float(draw_i(seed, step, e, stream, (uint32_t)p));
                if (dead * u > cfg->death_threshold && !(living > EPS)) food[p] += 1.0f;
            }
}

static void multi_delete_done(int K, int C, float* heads, float* bodies, const uint8_t* done) {
    for (int k = 0; k < K; ++k)                                              /* :595-596 / :676-677 */
        if (done[k]) {
            memset(heads + (size_t)k * C, 0, sizeof(float) * C);
            memset(bodies + (size_t)k * C, 0, sizeof(float) * C);
        }
}

static void multi_round(int K, int C, float* food, float* heads, float* bodies) {
    for (int p = 0; p < C; ++p) {                                            /* :599-605 / :688-694 */
        float f = nearbyintf(food[p]);
        food[p] = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
    }
    for (int p = 0; p < K * C; ++p) { heads[p] = nearbyintf(heads[p]); bodies[p] = nearbyintf(bodies[p]); }
}

static int multi_cell_free(int K, int C, const float* food, const float* heads, const float* bodies, int p) {
    float s = food[p];                                                       /* :432-440 cat + sum(dim=1) */
    for (int k = 0; k < K; ++k) s += heads[(size_t)k * C + p];
    for (int k = 0; k < K; ++k) s += bodies[(size_t)k * C + p];
    return s < EPS;
}

/* uniform free interior cell: rejection sampling, explicit ranking as the fallback; -1 if none */
static int multi_pick_free(int K, int S, const float* food, const float* heads, const float* bodies, uint64_t seed,
                           uint64_t step, uint32_t e, uint32_t stream) {
    int C = S * S, I = S - 2;
    for (uint32_t t = 0; t < REJECTION_TRIES; ++t) {
        int cand = (int)bounded(draw_i(seed, step, e, stream, t), (uint32_t)(I * I));
        int q = (1 + cand / I) * S + 1 + cand % I;
        if (multi_cell_free(K, C, food, heads, bodies, q)) return q;
    }
    int nfree = 0;
    for (int y = 1; y < S - 1; ++y)
        for (int x = 1; x < S - 1; ++x) nfree += multi_cell_free(K, C, food, heads, bodies, y * S + x);
    if (nfree == 0) return -1;
    int r = (int)bounded(draw_i(seed, step, e, stream, REJECTION_TRIES), (uint32_t)nfree);
    for (int y = 1; y < S - 1; ++y)
        for (int x = 1; x < S - 1; ++x)
            if (multi_cell_free(K, C, food, heads, bodies, y * S + x)) {
                if (r == 0) return y * S + x;
                --r;
            }
    return -1;
}

/* multi_snake.py:462-694 (everything before the observation), all envs.
 * actions (E,K) int64 in [0,8): a % 4 = direction, a > 3 = boost.
 * Outputs in (E,K) layout: rewards, snake_col, edge_col, food_cons (info food_i), sizes (info size_i);
 * all_done (E).  rate_selected (E, nullable): envs that drew a random_rate row (:382).
 * Returns 0, or -1 if a replayed tape is inconsistent with the state. */
int wurm_oracle_multi_step(const WurmOracleMultiCfg* cfg, float* foods, float* heads, float* bodies, uint8_t* dones,
                           int64_t* orientations, uint8_t* boost_this_step, const int64_t* actions,
                           const WurmOracleMultiStepDraws* draws, uint64_t seed, uint64_t step, float* rewards,
                           uint8_t* snake_col, uint8_t* edge_col, float* food_cons, float* sizes_out,
                           uint8_t* all_done, uint8_t* rate_selected) {
    const int E = cfg->num_envs, K = cfg->num_snakes, S = cfg->size, C = S * S;
    if (K > 64) return -2;
    const int replay = draws && draws->replay;
    float* sizes = (float*)malloc(sizeof(float) * (size_t)E * K);
    uint8_t* boosted = (uint8_t*)malloc((size_t)E * K);
    int* mv = (int*)malloc(sizeof(int) * (size_t)E * K);
    uint8_t* done0 = (uint8_t*)malloc((size_t)E * K);
    int any_boost = 0;

    /* :475-502 */
#pragma omp parallel for schedule(static) reduction(| : any_boost)
    for (int e = 0; e < E; ++e)
        for (int k = 0; k < K; ++k) {
            int n = e * K + k;
            const float* b = bodies + (size_t)n * C;
            float sz = b[0];
            for (int p = 1; p < C; ++p) sz = b[p] > sz ? b[p] : sz;          /* :489 */
            sizes[n] = sz;
            done0[n] = dones[n];                                             /* :490 */
            int64_t a = actions[n];
            int64_t m = a % 4;                                               /* :483 */
            if (orientations[n] == m) m = (m + 2) % 4;                       /* :336-339 */
            mv[n] = (int)m;
            orientations[n] = (m + 2) % 4;                                   /* :355-357, dead agents too */
            boosted[n] = (a > 3) && (sz >= 4.0f);                            /* :484,497-498 */
            boost_this_step[n] = boosted[n];                                 /* :499 */
            any_boost |= boosted[n];
            rewards[n] = 0.0f; food_cons[n] = 0.0f; snake_col[n] = 0; edge_col[n] = 0;
        }
    const int run_boost = cfg->boost && any_boost;                           /* :503 batch-global */
    if (replay && run_boost != draws->boost_phase_ran) { free(sizes); free(boosted); free(mv); free(done0); return -1; }

#pragma omp parallel
    {
        float* scratch = (float*)malloc(sizeof(float) * (2 * C + K));
        uint8_t all_active[64];
        memset(all_active, 1, sizeof(all_active));
#pragma omp for schedule(static)
        for (int e = 0; e < E; ++e) {
            float* food = foods + (size_t)e * C;
            float* hd = heads + (size_t)e * K * C;
            float* bd = bodies + (size_t)e * K * C;
            uint8_t* dn = dones + (size_t)e * K;
            int o = e * K;
            if (run_boost) {
                multi_phase(cfg, food, hd, bd, dn, mv + o, boosted + o, sizes + o, rewards + o, food_cons + o,
                            snake_col + o, edge_col + o, replay ? draws->u_boost + (size_t)e * C : NULL,
                            STREAM_MULTI_DEATH_BOOST, seed, step, (uint32_t)e, scratch);
                /* :579-592 boost cost */
                uint8_t cost[64];
                int any_cost = 0;
                for (int k = 0; k < K; ++k) {
                    float u = replay ? draws->u_cost[o + k] : unit_float(draw_i(seed, step, (uint32_t)e, STREAM_MULTI_BOOST_COST, (uint32_t)k));
                    cost[k] = boosted[o + k] && (u < cfg->boost_cost_prob);
                    any_cost |= cost[k];
                }
                if (any_cost) {
                    for (int p = 0; p < C; ++p) {                            /* :583-586 tails become food */
                        float tails = 0.0f;
                        for (int k = 0; k < K; ++k)
                            if (cost[k] && bd[(size_t)k * C + p] == 1.0f) tails += 1.0f;
                        if (tails > EPS) food[p] += 1.0f;
                    }
                    for (int k = 0; k < K; ++k) {
                        if (!cost[k]) continue;
                        for (int p = 0; p < C; ++p) {                        /* :588-589 */
                            float b = bd[(size_t)k * C + p] - 1.0f;
                            bd[(size_t)k * C + p] = b > 0.0f ? b : 0.0f;
                        }
                        rewards[o + k] -= 1.0f;                              /* :590 */
                        sizes[o + k] -= 1.0f;                                /* :591 */
                    }
                }
                multi_delete_done(K, C, hd, bd, dn);                         /* :595-596 */
                multi_round(K, C, food, hd, bd);                             /* :599-605 */
            }
            multi_phase(cfg, food, hd, bd, dn, mv + o, all_active, sizes + o, rewards + o, food_cons + o,
                        snake_col + o, edge_col + o, replay ? draws->u_reg + (size_t)e * C : NULL,
                        STREAM_MULTI_DEATH_REGULAR, seed, step, (uint32_t)e, scratch);
            multi_delete_done(K, C, hd, bd, dn);                             /* :676-677 */
        }
        free(scratch);
    }

    /* :680 _add_food (:368-410).  random_rate rows are indexed by rank among the selected envs. */
    int32_t* rate_row = (int32_t*)malloc(sizeof(int32_t) * (size_t)E);
    int rows = 0;
    for (int e = 0; e < E; ++e) {
        rate_row[e] = -1;
        if (cfg->food_mode == 1) {
            float total = 0.0f;
            for (int p = 0; p < C; ++p) total += foods[(size_t)e * C + p];
            if (total < (float)(K * 8)) rate_row[e] = rows++;                /* :382, max_food :127 */
        }
        if (rate_selected) rate_selected[e] = rate_row[e] >= 0;
    }
    int bad = replay && cfg->food_mode == 1 && rows != draws->n_rate_rows;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < E; ++e) {
        float* food = foods + (size_t)e * C;
        float* hd = heads + (size_t)e * K * C;
        float* bd = bodies + (size_t)e * K * C;
        int o = e * K;
        if (cfg->food_mode == 0) {                                           /* only_one :369-379 */
            float total = 0.0f;
            for (int p = 0; p < C; ++p) total += food[p];
            if (total < EPS) {
                int cell = replay ? draws->food_cell[e] : multi_pick_free(K, S, food, hd, bd, seed, step, (uint32_t)e, STREAM_MULTI_FOOD_ONE);
                if (cell >= 0) food[cell] += 1.0f;
            }
        } else if (rate_row[e] >= 0 && !bad) {                               /* random_rate :380-408 */
            for (int y = 1; y < S - 1; ++y)
                for (int x = 1; x < S - 1; ++x) {
                    int p = y * S + x;
                    if (!multi_cell_free(K, C, food, hd, bd, p)) continue;
                    float u = replay ? draws->u_rate[(size_t)rate_row[e] * C + p]
                                     : unit_float(draw_i(seed, step, (uint32_t)e, STREAM_MULTI_FOOD_RATE, (uint32_t)p));
                    if (u < cfg->food_rate) food[p] += 1.0f;
                }
        }
        int all = 1;
        for (int k = 0; k < K; ++k) {
            int n = o + k;
            if (dones[n] && !done0[n]) rewards[n] += cfg->reward_on_death;   /* :683-685 */
            sizes_out[n] = sizes[n];
            all &= dones[n];
        }
        all_done[e] = (uint8_t)all;                                          /* :703 (the lifetime cap :705 never fires) */
        multi_round(K, C, food, hd, bd);                                     /* :688-694 */
    }
    free(rate_row); free(sizes); free(boosted); free(mv); free(done0);
    return bad ? -1 : 0;
}

/* multi_snake.py:194-227 _get_env_images: int16 colour of one cell of one env */
static void multi_env_pixel(int K, int S, const float* food, const float* heads, const float* bodies,
                            const uint8_t* boost, const int16_t* colours, int y, int x, int16_t rgb[3]) {
    int C = S * S, p = y * S + x;
    float acc[3] = {0.0f, 0.0f, 0.0f};
    for (int k = 0; k < K; ++k) {
        float inten = (bodies[(size_t)k * C + p] > EPS ? 1.0f : 0.0f) * 1.0f / 3.0f +
                      (heads[(size_t)k * C + p] > EPS ? 1.0f : 0.0f) * 1.0f / 3.0f;          /* :197 */
        inten *= 1.0f + 0.5f * (boost[k] ? 1.0f : 0.0f);                                      /* :198 */
        for (int c = 0; c < 3; ++c) acc[c] += inten * (float)colours[3 * k + c];              /* :201-205 */
    }
    for (int c = 0; c < 3; ++c) rgb[c] = (int16_t)acc[c];                                     /* :206 .short() truncates */
    if (food[p] > EPS) rgb[0] = (int16_t)(rgb[0] + 255);                                      /* :208-209 */
    if (rgb[0] == 0 && rgb[1] == 0 && rgb[2] == 0) rgb[0] = rgb[1] = rgb[2] = 255;            /* :214-219 */
    if (y == 0 || x == 0 || y == S - 1 || x == S - 1) rgb[0] = rgb[1] = rgb[2] = 0;           /* :225 */
}

/* multi_snake.py:283-334 _observe.  mode 0 'full' -> obs (K,E,3,S,S); mode 1 'partial_n' -> (K,E,3,W,W).
 * (agent-major: obs[k] is the tensor the reference returns under 'agent_k'.)
 * Returns the number of living agents without exactly one head cell (the reference would raise). */
int wurm_oracle_multi_observe(const WurmOracleMultiCfg* cfg, const float* foods, const float* heads, const float* bodies,
                              const uint8_t* dones, const uint8_t* boost_this_step, const int16_t* colours, int mode,
                              int n, float* obs) {
    const int E = cfg->num_envs, K = cfg->num_snakes, S = cfg->size, C = S * S, W = 2 * n + 1;
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int e = 0; e < E; ++e) {
        const float* food = foods + (size_t)e * C;
        const float* hd = heads + (size_t)e * K * C;
        const float* bd = bodies + (size_t)e * K * C;
        int16_t rgb[3];
        for (int k = 0; k < K; ++k) {
            if (mode == 0) {                                                 /* :268-281 _observe_agent */
                float* o = obs + ((size_t)k * E + e) * 3 * C;
                for (int y = 0; y < S; ++y)
                    for (int x = 0; x < S; ++x) {
                        int p = y * S + x;
                        float ob = 0.0f, oh = 0.0f;
                        for (int j = 0; j < K; ++j)
                            if (j != k) { ob += bd[(size_t)j * C + p]; oh += hd[(size_t)j * C + p]; }
                        rgb[0] = rgb[1] = rgb[2] = 255;                      /* :176 */
                        if (food[p] > EPS) { rgb[0] = 255; rgb[1] = 0; rgb[2] = 0; }               /* :275 */
                        if (bd[(size_t)k * C + p] > EPS) { rgb[0] = 0; rgb[1] = 96; rgb[2] = 0; }  /* :276 self/2 */
                        if (hd[(size_t)k * C + p] > EPS) { rgb[0] = 0; rgb[1] = 192; rgb[2] = 0; } /* :277 */
                        if (ob > EPS) { rgb[0] = 0; rgb[1] = 0; rgb[2] = 96; }                     /* :278 other/2 */
                        if (oh > EPS) { rgb[0] = 0; rgb[1] = 0; rgb[2] = 192; }                    /* :279 */
                        if (y == 0 || x == 0 || y == S - 1 || x == S - 1) rgb[0] = rgb[1] = rgb[2] = 0;
                        for (int c = 0; c < 3; ++c) o[c * C + p] = (float)rgb[c] / 255.0f;         /* :281 */
                    }
            } else {                                                         /* :289-332 partial_n */
                float* o = obs + ((size_t)k * E + e) * 3 * W * W;
                memset(o, 0, sizeof(float) * 3 * W * W);
                if (dones[e * K + k]) continue;                              /* :320-323 zeros for the dead */
                int hp = -1, nh = 0;
                for (int p = 0; p < C; ++p)
                    if (hd[(size_t)k * C + p] != 0.0f) { hp = p; ++nh; }
                if (nh != 1) { ++bad; continue; }
                int hy = hp / S, hx = hp % S;
                for (int i = 0; i < W; ++i)
                    for (int j = 0; j < W; ++j) {
                        int y = hy - n + i, x = hx - n + j;
                        if (y < 0 || y >= S || x < 0 || x >= S) continue;    /* zero padding :301-302 */
                        multi_env_pixel(K, S, food, hd, bd, boost_this_step + e * K, colours + 3 * (size_t)e * K, y, x, rgb);
                        for (int c = 0; c < 3; ++c) o[c * W * W + i * W + j] = (float)rgb[c] / 255.0f;  /* :296 */
                    }
            }
        }
    }
    return bad;
}

/* multi_snake.py:194-227 as an (E,3,S,S) int16 image (used by render and by tests) */
void wurm_oracle_multi_env_images(const WurmOracleMultiCfg* cfg, const float* foods, const float* heads,
                                  const float* bodies, const uint8_t* boost_this_step, const int16_t* colours,
                                  int16_t* img) {
    const int E = cfg->num_envs, K = cfg->num_snakes, S = cfg->size, C = S * S;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < E; ++e) {
        int16_t rgb[3];
        for (int y = 0; y < S; ++y)
            for (int x = 0; x < S; ++x) {
                multi_env_pixel(K, S, foods + (size_t)e * C, heads + (size_t)e * K * C, bodies + (size_t)e * K * C,
                                boost_this_step + e * K, colours + 3 * (size_t)e * K, y, x, rgb);
                for (int c = 0; c < 3; ++c) img[((size_t)e * 3 + c) * C + y * S + x] = rgb[c];
            }
    }
}

/* Availability map for a new snake (:846-858 / :927-941): not within the 3x3 neighbourhood of an
 * occupied cell, and at least 2 cells away from the wall.  Returns the number of available cells;
 * if r >= 0 also returns (through *cell) the r-th available cell in raster order. */
static int multi_spawn_cells(int K, int S, const float* food, const float* heads, const float* bodies, int r, int* cell) {
    int C = S * S, n = 0;
    for (int y = 2; y < S - 2; ++y)
        for (int x = 2; x < S - 2; ++x) {
            int ok = 1;
            for (int dy = -1; dy <= 1 && ok; ++dy)
                for (int dx = -1; dx <= 1 && ok; ++dx)
                    if (!multi_cell_free(K, C, food, heads, bodies, (y + dy) * S + x + dx)) ok = 0;
            if (!ok) continue;
            if (n == r && cell) *cell = y * S + x;
            ++n;
        }
    return n;
}

static void multi_stamp_snake(int S, float* head, float* body, int cell, int d) {
    int y = cell / S, x = cell % S;                                          /* LENGTH_3_SNAKES, as single_create_env */
    body[(y - OFF_Y[d]) * S + (x - OFF_X[d])] = 1.0f;
    body[y * S + x] = 2.0f;
    body[(y + OFF_Y[d]) * S + (x + OFF_X[d])] = 3.0f;
    head[(y + OFF_Y[d]) * S + (x + OFF_X[d])] = 1.0f;
}

/* Colour of one agent from three uniforms (:163-169), float32 arithmetic in this fixed order. */
static void multi_colour(float c0, float c1, float c2, int16_t out[3]) {
    c0 = c0 / 1.5f;
    float norm = sqrtf(c0 * c0 + c1 * c1 + c2 * c2);
    out[0] = (int16_t)(c0 / norm * 192.0f);
    out[1] = (int16_t)(c1 / norm * 192.0f);
    out[2] = (int16_t)(c2 / norm * 192.0f);
}

/* The random draws of one MultiSnake.reset / __init__.  replay == 0: Philox.
 *   create (E, K+1, 2) int32: for envs being re-created, per snake (seed cell, direction), then (food cell, 0)
 *   respawn (E, 2) int32:     (seed cell or -1, direction) for the env's first dead agent
 *   colours (E*K, 3) int16:   new colour of every agent that is (still) dead                         */
typedef struct WurmOracleMultiResetDraws {
    int32_t replay;
    const int32_t* create;
    const int32_t* respawn;
    const int16_t* colours;
} WurmOracleMultiResetDraws;

/* multi_snake.py:771-831 (+ _create_envs :996-1019, _add_snake :911-994, _get_snake_addition :838-909).
 * env_done (E): envs to re-create.  Returns the number of envs where a snake could not be created
 * (the reference raises RuntimeError :865,947). */
int wurm_oracle_multi_reset(const WurmOracleMultiCfg* cfg, float* foods, float* heads, float* bodies, uint8_t* dones,
                            int64_t* orientations, int16_t* colours, const uint8_t* env_done,
                            const WurmOracleMultiResetDraws* draws, uint64_t seed, uint64_t step) {
    const int E = cfg->num_envs, K = cfg->num_snakes, S = cfg->size, C = S * S;
    const int replay = draws && draws->replay;
    int failed = 0;
#pragma omp parallel for schedule(static) reduction(+ : failed)
    for (int e = 0; e < E; ++e) {
        float* food = foods + (size_t)e * C;
        float* hd = heads + (size_t)e * K * C;
        float* bd = bodies + (size_t)e * K * C;
        uint8_t* dn = dones + (size_t)e * K;
        if (env_done[e]) {                                                   /* :787-798 */
            memset(food, 0, sizeof(float) * C);
            memset(hd, 0, sizeof(float) * K * C);
            memset(bd, 0, sizeof(float) * K * C);
            for (int k = 0; k < K; ++k) {                                    /* :1004-1006 */
                int cell = -1, d;
                if (replay) {
                    cell = draws->create[((size_t)e * (K + 1) + k) * 2];
                    d = draws->create[((size_t)e * (K + 1) + k) * 2 + 1];
                } else {
                    philox4 r = draw(seed, step, (uint32_t)e, STREAM_MULTI_CREATE_SNAKE | ((uint32_t)k << 4));
                    int n = multi_spawn_cells(K, S, food, hd, bd, -1, NULL);
                    if (n > 0) multi_spawn_cells(K, S, food, hd, bd, (int)bounded(r.v[0], (uint32_t)n), &cell);
                    d = (int)(r.v[1] >> 30);
                }
                if (cell < 0) { ++failed; continue; }
                multi_stamp_snake(S, hd + (size_t)k * C, bd + (size_t)k * C, cell, d);
                orientations[e * K + k] = d;                                 /* :793,1019 */
            }
            int fcell;
            if (replay) fcell = draws->create[((size_t)e * (K + 1) + K) * 2];
            else {                                                           /* :1016 one food on a free interior cell */
                int nfree = 0, r;
                for (int y = 1; y < S - 1; ++y)
                    for (int x = 1; x < S - 1; ++x) nfree += multi_cell_free(K, C, food, hd, bd, y * S + x);
                fcell = -1;
                r = nfree > 0 ? (int)bounded(draw(seed, step, (uint32_t)e, STREAM_MULTI_CREATE_FOOD).v[0], (uint32_t)nfree) : -1;
                for (int y = 1; y < S - 1 && fcell < 0 && r >= 0; ++y)
                    for (int x = 1; x < S - 1; ++x)
                        if (multi_cell_free(K, C, food, hd, bd, y * S + x)) {
                            if (r == 0) { fcell = y * S + x; break; }
                            --r;
                        }
            }
            if (fcell >= 0) food[fcell] += 1.0f;
            for (int k = 0; k < K; ++k) dn[k] = 0;                           /* :798 */
        }
        if (cfg->colour_random)                                              /* :800-803 */
            for (int k = 0; k < K; ++k) {
                if (!dn[k]) continue;
                int16_t* col = colours + 3 * ((size_t)e * K + k);
                if (replay) memcpy(col, draws->colours + 3 * ((size_t)e * K + k), 3 * sizeof(int16_t));
                else {
                    philox4 r = draw(seed, step, (uint32_t)e, STREAM_MULTI_COLOUR | ((uint32_t)k << 4));
                    multi_colour(unit_float(r.v[0]), unit_float(r.v[1]), unit_float(r.v[2]), col);
                }
            }
        if (cfg->respawn_any) {                                              /* :805-829 first dead snake of the env */
            int k = 0;
            while (k < K && !dn[k]) ++k;
            if (k < K) {
                int cell = -1, d;
                if (replay) {
                    cell = draws->respawn[2 * (size_t)e];
                    d = draws->respawn[2 * (size_t)e + 1];
                } else {
                    philox4 r = draw(seed, step, (uint32_t)e, STREAM_MULTI_RESPAWN);
                    int n = multi_spawn_cells(K, S, food, hd, bd, -1, NULL);
                    if (n > 0) multi_spawn_cells(K, S, food, hd, bd, (int)bounded(r.v[0], (uint32_t)n), &cell);
                    d = (int)(r.v[1] >> 30);
                }
                memset(hd + (size_t)k * C, 0, sizeof(float) * C);            /* :826-827 new tensors replace the old */
                memset(bd + (size_t)k * C, 0, sizeof(float) * C);
                if (cell >= 0) multi_stamp_snake(S, hd + (size_t)k * C, bd + (size_t)k * C, cell, d);
                orientations[e * K + k] = d;                                 /* :828 even when the spawn failed */
                dn[k] = cell < 0;                                            /* :829 */
            }
        }
    }
    return failed;
}

/* ------------------------------------------------------------------------------------------ */
/* SimpleGridworld (wurm/envs/simple_gridworld.py): 2 channels, food and agent ("head").          */
/* ------------------------------------------------------------------------------------------ */
enum { STREAM_GRID_STEP_FOOD = 11, STREAM_GRID_RESET = 12 };

static int grid_pick_free(int S, const float* env, uint64_t seed, uint64_t step, uint32_t e, uint32_t stream) {
    int C = S * S, I = S - 2;                                                /* :208-216 free = food + agent < EPS, interior */
    for (uint32_t t = 0; t < REJECTION_TRIES; ++t) {
        int cand = (int)bounded(draw_i(seed, step, e, stream, t), (uint32_t)(I * I));
        int q = (1 + cand / I) * S + 1 + cand % I;
        if (env[q] + env[C + q] < EPS) return q;
    }
    int nfree = 0;
    for (int y = 1; y < S - 1; ++y)
        for (int x = 1; x < S - 1; ++x) nfree += env[y * S + x] + env[C + y * S + x] < EPS;
    if (nfree == 0) return -1;
    int r = (int)bounded(draw_i(seed, step, e, stream, REJECTION_TRIES), (uint32_t)nfree);
    for (int y = 1; y < S - 1; ++y)
        for (int x = 1; x < S - 1; ++x)
            if (env[y * S + x] + env[C + y * S + x] < EPS && r-- == 0) return y * S + x;
    return -1;
}

/* simple_gridworld.py:135-201.  food_cell_replay as for the snake envs; actions are NOT sanitised. */
void wurm_oracle_grid_step(int N, int S, float* envs, const int64_t* actions, const int32_t* food_cell_replay,
                           uint64_t seed, uint64_t step, float* reward, uint8_t* done) {
    int C = S * S;
#pragma omp parallel
    {
        float* moved = (float*)malloc(sizeof(float) * C);
#pragma omp for schedule(static)
        for (int e = 0; e < N; ++e) {
            float* food = envs + (size_t)e * 2 * C;
            float* head = food + C;
            int a = (int)actions[e];
            for (int y = 0; y < S; ++y)                                      /* :149-158 head += conv2d(head)[a]; round */
                for (int x = 0; x < S; ++x) {
                    int yy = y + OFF_Y[a], xx = x + OFF_X[a], p = y * S + x;
                    float nb = (yy >= 0 && yy < S && xx >= 0 && xx < S) ? head[yy * S + xx] : 0.0f;
                    moved[p] = nearbyintf(head[p] + (nb - head[p]));
                }
            memcpy(head, moved, sizeof(float) * C);
            float removed = 0.0f;                                            /* :169-171 */
            for (int p = 0; p < C; ++p) {
                float rem = head[p] * food[p] * -1.0f;
                removed += rem;
                food[p] += rem;
            }
            reward[e] = 0.0f - removed;
            if (removed * -1.0f != 0.0f) {                                   /* :176-181 */
                int cell = food_cell_replay ? food_cell_replay[e] : grid_pick_free(S, food, seed, step, (uint32_t)e, STREAM_GRID_STEP_FOOD);
                if (cell >= 0) food[cell] += 1.0f;
            }
            float interior = 0.0f;                                           /* :189-192 */
            for (int y = 1; y < S - 1; ++y)
                for (int x = 1; x < S - 1; ++x) interior += head[y * S + x];
            done[e] = interior < EPS;
            for (int p = 0; p < 2 * C; ++p) food[p] = nearbyintf(food[p]);   /* :197 */
        }
        free(moved);
    }
}

/* simple_gridworld.py:222-262: done envs get the agent at (start_y, start_x) and one food. */
void wurm_oracle_grid_reset(int N, int S, float* envs, const uint8_t* done, int start_y, int start_x,
                            const int32_t* food_cell_replay, uint64_t seed, uint64_t step) {
    int C = S * S;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < N; ++e) {
        if (!done[e]) continue;
        float* env = envs + (size_t)e * 2 * C;
        memset(env, 0, sizeof(float) * 2 * C);
        env[C + start_y * S + start_x] = 1.0f;                               /* :255 */
        int cell = food_cell_replay ? food_cell_replay[e] : grid_pick_free(S, env, seed, step, (uint32_t)e, STREAM_GRID_RESET);
        if (cell >= 0) env[cell] += 1.0f;                                    /* :258-259 */
    }
}

/* simple_gridworld.py:88-133.  mode 0 'default' (N,3,S,S): BLACK background (:90 zeros * 255), agent
 * (0,255,0), food (255,0,0), edges 0, /255; mode 1 'raw' (N,2,S,S); mode 3 'positions' (N,4) (the
 * reference builds (1,4) and only works for N == 1). */
void wurm_oracle_grid_observe(int N, int S, const float* envs, int mode, float* obs) {
    int C = S * S;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < N; ++e) {
        const float* env = envs + (size_t)e * 2 * C;
        if (mode == OBS_DEFAULT) {
            float* o = obs + (size_t)e * 3 * C;
            for (int y = 0; y < S; ++y)
                for (int x = 0; x < S; ++x) {
                    int p = y * S + x;
                    int16_t rgb[3] = {0, 0, 0};
                    if (env[C + p] > EPS) { rgb[0] = 0; rgb[1] = 255; rgb[2] = 0; }
                    if (env[p] > EPS) { rgb[0] = 255; rgb[1] = 0; rgb[2] = 0; }
                    if (y == 0 || x == 0 || y == S - 1 || x == S - 1) rgb[0] = rgb[1] = rgb[2] = 0;
                    for (int c = 0; c < 3; ++c) o[c * C + p] = (float)rgb[c] / 255.0f;
                }
        } else if (mode == OBS_RAW) {
            memcpy(obs + (size_t)e * 2 * C, env, sizeof(float) * 2 * C);
        } else {
            float* o = obs + (size_t)e * 4;
            int h = argmax_first(env + C, C), f = argmax_first(env, C);
            o[0] = (float)(h / S); o[1] = (float)(h % S); o[2] = (float)(f / S); o[3] = (float)(f % S);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* A2C return scan (wurm/rl/a2c.py:49-63), fp32, the reference's operation order.                */
/* rewards, values, dones, returns: (T,N) row-major; bootstrap: (N).  lambda < 0: n-step returns  */
/* (:58-61), else generalised advantage estimation (:50-57).                                     */
/* ------------------------------------------------------------------------------------------ */
void wurm_oracle_a2c_returns(int T, int N, double gamma_d, double lambda_d, const float* bootstrap, const float* rewards,
                             const float* values, const uint8_t* dones, float* returns) {
    /* the reference's scalars are Python doubles: each meets an fp32 tensor as its own fp32 rounding, and
       `self.gamma * self.gae_lambda` (:56) is multiplied in double BEFORE it is rounded */
    const float gamma = (float)gamma_d, gamma_lambda = (float)(gamma_d * lambda_d);
    const int use_gae = lambda_d >= 0.0;
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
        if (!use_gae) {
            float R = bootstrap[n] * (dones[(size_t)(T - 1) * N + n] ? 0.0f : 1.0f);                 /* :58 */
            for (int t = T - 1; t >= 0; --t) {
                float m = dones[(size_t)t * N + n] ? 0.0f : 1.0f;
                R = rewards[(size_t)t * N + n] + gamma * R * m;                                      /* :60 */
                returns[(size_t)t * N + n] = R;
            }
        } else {
            float gae = 0.0f;
            for (int t = T - 1; t >= 0; --t) {
                float m = dones[(size_t)t * N + n] ? 0.0f : 1.0f;
                float next = t == T - 1 ? bootstrap[n] : values[(size_t)(t + 1) * N + n];
                float delta = rewards[(size_t)t * N + n] + gamma * next * m - values[(size_t)t * N + n];   /* :52-55 */
                gae = delta + gamma_lambda * m * gae;                                                /* :56 */
                returns[(size_t)t * N + n] = gae + values[(size_t)t * N + n];                        /* :57 */
            }
        }
    }
}
