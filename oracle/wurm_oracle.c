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

/* Draw streams: counter = (unit, stream, step_lo, step_hi), key = (seed_lo, seed_hi). */
enum { STREAM_SINGLE_STEP_FOOD = 0, STREAM_SINGLE_RESET = 1 };

static philox4 draw(uint64_t seed, uint64_t step, uint32_t unit, uint32_t stream) {
    return philox4x32_10(unit, stream, (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
}

/* uniform integer in [0,n) by multiply-shift */
static uint32_t bounded(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

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
            int nfree = count_free_single(S, env);
            cell = nfree > 0 ? nth_free_single(S, env, (int)bounded(draw(seed, step, e, STREAM_SINGLE_STEP_FOOD).v[0], (uint32_t)nfree)) : -1;
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
