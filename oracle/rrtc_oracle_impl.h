/* CPU ORACLE (test infrastructure, NOT product code): restatement of the batched RRT-Connect front end
 * (og.RRTConnect as configured at planning.py:151-156,190; OMPL defaults of SURVEY.md App. D) in fp32, one query
 * at a time, mirroring the decision sequence of the device planner: same counter-based sample stream
 * (Philox-4x32-10 keyed by (seed, "RRTC"), counter (iteration, search, block, 1)), first extension aimed at the
 * goal, brute-force nearest neighbour with lowest-index tie break, steer by `range`, DiscreteMotionValidator at
 * `resolution`, trees swapped every iteration, greedy CONNECT, then deterministic farthest-first shortcutting.
 * With identical verdicts the returned path is bit-identical to the device's; verdicts can differ only for states
 * inside the 1e-4 m contact band.  Included by panda_oracle.c after the f32 instantiation.
 */
#include <stdint.h>
#include <stdlib.h>

static void po_philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

static void po_rrtc_sample(const po_model_f32 *m, uint32_t seed, uint32_t search, uint32_t it, float *q, float *extra) {
    float u[12];
    for (uint32_t blk = 0; blk < 3; ++blk) {
        uint32_t c[4] = {it, search, blk, 1u};
        po_philox(c, seed, 0x52525443u);
        for (int j = 0; j < 4; ++j) u[4 * blk + j] = (float)(c[j] >> 8) * 5.9604644775390625e-08f;
    }
    for (int j = 0; j < 9; ++j) q[j] = fmaf(u[j], m->q_upper[j] - m->q_lower[j], m->q_lower[j]);
    *extra = u[9];
}

typedef struct {
    const po_model_f32 *m;
    const float *obb;
    int n_obb;
    float table_z;
    const float *base;
    int attached, flags;
    float resolution;
    long long checks;
} po_rrtc_ctx;

/* motion validity of a -> b; counts the states a full (no early exit) validation would touch up to the failing
 * "round" of 32 like the device does */
static int po_rrtc_motion_valid(po_rrtc_ctx *c, const float *a, const float *b) {
    float de[9], d2 = 0.f;
    for (int k = 0; k < 9; ++k) {
        de[k] = b[k] - a[k];
        d2 = fmaf(de[k], de[k], d2);
    }
    int nd = (int)ceilf(sqrtf(d2) / c->resolution);
    if (nd < 1) nd = 1;
    const int rounds = (nd + 31) >> 5;
    const float inv_nd = 1.0f / (float)nd;
    for (int r = 0; r < rounds; ++r) {
        int hit = 0;
        int left = nd - r * 32;
        c->checks += left < 32 ? left : 32;
        for (int lane = 0; lane < 32; ++lane) {
            int k = nd - (lane * rounds + r);
            if (k < 1) continue;
            float q[9];
            const float t = (float)k * inv_nd;
            for (int j = 0; j < 9; ++j) q[j] = (k == nd) ? b[j] : fmaf(t, de[j], a[j]);
            if (state_margin_one_f32(c->m, c->obb, c->n_obb, c->table_z, c->base, c->attached, c->flags, q) < 0.f) {
                hit = 1;
                break;
            }
        }
        if (hit) return 0;
    }
    return 1;
}

/* returns the path length (0 = no solution); path_out is [max_path][9] */
int po_rrtc_f32(const po_model_f32 *m, const float *obb, int n_obb, float table_z, const float *base, int attached,
                int flags, const float *start, const float *goal, float range, float resolution, int max_iters,
                int max_nodes, int max_path, uint32_t seed, uint32_t search, int shortcut_passes, int planner,
                float *path_out, int *iters_out, long long *checks_out) {
    po_rrtc_ctx ctx = {m, obb, n_obb, table_z, base, attached, flags, resolution, 0};
    const int M = max_nodes;
    float *tq = (float *)malloc(sizeof(float) * 2 * 9 * (size_t)M); /* [tree][node][9] */
    int *par = (int *)malloc(sizeof(int) * 2 * (size_t)M);
    int size[2] = {1, 1};
    memcpy(tq, start, 9 * sizeof(float));
    memcpy(tq + (size_t)9 * M, goal, 9 * sizeof(float));
    par[0] = -1;
    par[M] = -1;
    int cur = 0, it = 0, solved = 0, added_idx = 0, conn_idx = -1, path_n = 0;
    float target[9];
    enum { EXTEND, CONNECT } phase = EXTEND;
    for (;;) {
        float goal_q[9] = {0}, ea[9], eb[9];
        int tree, aim_goal = 0;
        if (phase == EXTEND) {
            if (it >= max_iters || size[0] >= M - 1 || size[1] >= M - 1) break;
            tree = cur;
            float u9 = 1.f;
            if (it > 0) po_rrtc_sample(m, seed, search, (uint32_t)it, goal_q, &u9);
            aim_goal = (it == 0) || (planner == 1 && u9 < 0.05f);
            if (aim_goal) memcpy(goal_q, tq + (size_t)9 * M, sizeof(goal_q));
        } else {
            tree = cur ^ 1;
            memcpy(goal_q, target, sizeof(goal_q));
        }
        const float *tt = tq + (size_t)tree * 9 * M;
        float bd = 3.0e38f;
        int bi = 0;
        for (int i = 0; i < size[tree]; ++i) {
            float d2 = 0.f;
            for (int k = 0; k < 9; ++k) {
                float d = tt[9 * i + k] - goal_q[k];
                d2 = fmaf(d, d, d2);
            }
            if (d2 < bd) {
                bd = d2;
                bi = i;
            }
        }
        const float d = sqrtf(bd);
        float f = 1.0f;
        int reach = 1;
        if (d > range) {
            f = range / d;
            reach = 0;
        }
        for (int k = 0; k < 9; ++k) {
            ea[k] = tt[9 * bi + k];
            eb[k] = reach ? goal_q[k] : fmaf(f, goal_q[k] - ea[k], ea[k]);
        }
        if (po_rrtc_motion_valid(&ctx, ea, eb)) {
            const int ni = size[tree];
            memcpy(tq + ((size_t)tree * M + ni) * 9, eb, 9 * sizeof(float));
            par[tree * M + ni] = bi;
            size[tree] = ni + 1;
            if (phase == EXTEND && planner == 1) {
                added_idx = ni;
                if (reach && aim_goal) {
                    solved = 1;
                    break;
                }
                ++it;
            } else if (phase == EXTEND) {
                memcpy(target, eb, sizeof(target));
                added_idx = ni;
                phase = CONNECT;
            } else if (reach) {
                conn_idx = ni;
                solved = 1;
                break;
            } else if (size[tree] >= M - 1) {
                break;
            }
        } else {
            phase = EXTEND;
            if (planner == 0) cur ^= 1;
            ++it;
        }
    }
    if (solved) {
        const int is = cur == 0 ? added_idx : conn_idx, ig = cur == 0 ? conn_idx : added_idx;
        int ds = 0, dg = 0;
        for (int x = is; x >= 0; x = par[x]) ++ds;
        if (planner == 0)
            for (int x = par[M + ig]; x >= 0; x = par[M + x]) ++dg;
        path_n = ds + dg;
        if (path_n > max_path) {
            solved = 0;
            path_n = 0;
        } else {
            int x = is;
            for (int k = ds - 1; k >= 0; --k) {
                memcpy(path_out + 9 * k, tq + (size_t)9 * x, 9 * sizeof(float));
                x = par[x];
            }
            x = planner == 0 ? par[M + ig] : -1;
            for (int k = 0; k < dg; ++k) {
                memcpy(path_out + 9 * (ds + k), tq + ((size_t)M + x) * 9, 9 * sizeof(float));
                x = par[M + x];
            }
        }
    }
    if (solved && shortcut_passes > 0 && path_n > 2) {
        int pass = 0, i = 0, j = path_n - 1, budget = 96;
        for (;;) {
            const int ok = po_rrtc_motion_valid(&ctx, path_out + 9 * i, path_out + 9 * j);
            --budget;
            if (ok) {
                const int drop = j - i - 1;
                memmove(path_out + 9 * (i + 1), path_out + 9 * j, (size_t)(path_n - j) * 9 * sizeof(float));
                path_n -= drop;
                ++i;
                j = path_n - 1;
            } else {
                --j;
            }
            if (j < i + 2) {
                ++i;
                j = path_n - 1;
            }
            if (i + 2 >= path_n && j < i + 2) {
                ++pass;
                i = 0;
                j = path_n - 1;
            }
            if (pass >= shortcut_passes || path_n <= 2 || budget <= 0) break;
        }
    }
    if (iters_out) *iters_out = solved ? it + 1 : it;
    if (checks_out) *checks_out = ctx.checks;
    free(tq);
    free(par);
    return solved ? path_n : 0;
}

/* Motion validator with the callback signature of pv_simplify_path_cb (include/panda_validity.h): lets the CPU arm of
 * bench.py and the CPU tests run the product's batched simplifier with THIS oracle answering, no Python in the loop.
 * user = po_edge_ctx_f32. */
typedef struct {
    const po_model_f32 *m;
    const float *obb;
    int n_obb;
    float table_z;
    const float *base;
    int attached, flags;
    float resolution;
    long long motions, states;
} po_edge_ctx_f32;

int po_edge_callback_f32(void *user, const float *a, const float *b, int n, unsigned char *ok) {
    po_edge_ctx_f32 *c = (po_edge_ctx_f32 *)user;
    for (int e = 0; e < n; ++e) {
        float m;
        long cnt = 0;
        po_edge_margin_f32(c->m, c->obb, c->n_obb, c->table_z, c->base, c->attached, c->flags, a + 9 * e, b + 9 * e, 1, 0,
                           c->resolution, 1, &m, &cnt, 1);
        ok[e] = m >= 0.f ? 1 : 0;
        c->states += cnt;
    }
    c->motions += n;
    return 0;
}
