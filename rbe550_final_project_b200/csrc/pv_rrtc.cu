// pv_rrtc.cu -- batched multi-query RRT-Connect on the device (K4): one warp per search, trees and paths
// in HBM, the whole solve + path extraction + shortcutting in ONE kernel launch.
//
// Restates og.RRTConnect as the reference configures it (planning.py:151-156,190: SimpleSetup defaults):
// RealVectorStateSpace(9) with L2 distance and linear interpolation, bounds = joint limits
// (planning.py:139-150), range = 20 % of the space extent, motion validity = DiscreteMotionValidator at 1 %
// of the extent.  Per iteration: sample uniformly, EXTEND the active tree towards the
// sample, then CONNECT the other tree greedily towards the new node; trees swap every iteration.
//
// Warp roles: lanes stride over tree nodes for the brute-force nearest-neighbour search (trees are 10^1..10^3
// nodes, SoA so the loads coalesce) and over the interpolated states of the edge being validated (one
// pv_check_config per lane, __any_sync early exit).  The solve is a small state machine so the ~10k-instruction
// state check is instantiated exactly once.
//
// Memory: a search's arena is 2 trees x 9 x M floats + parents + a path buffer.  Sized for the
// caller's max_nodes (2 048) that is 168 KB per search while the median search uses 2 nodes, so a batch is planned in
// two tiers: tier 1 gives every search small trees (RRTC_T1_NODES per tree); a query one of whose searches hit that
// cap BEFORE any search of the query had connected is planned again in tier 2 with the full max_nodes.  A search is a
// deterministic function of (seed, global search id) and of nothing else, and the tree capacity only decides where
// it gives up, so both tiers return exactly what a single full-size run returns.  Batches are cut into chunks of at most
// RRTC_CHUNK_SEARCHES searches (tier 1) / RRTC_T2_ARENA_BYTES of trees (tier 2), so the arena is bounded whatever
// the batch size, and only the USED rows of each path are gathered (packed) and copied to the host.
//
// Replicas (several searches per query) are OR-parallel but DETERMINISTIC: the winner is the search with the smallest
// key (iterations, replica id).  A running search gives up as soon as a published key is smaller than any key it can
// still reach, which never removes the eventual minimum, so the answer does not depend on warp scheduling.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/panda_validity.h"
#include "pv_device.cuh"
#include "pv_handle.h"

#define RRTC_THREADS 128
#define RRTC_MAX_SHORTCUT_CHECKS 96
#ifndef RRTC_T1_NODES
#define RRTC_T1_NODES 64
#endif
#ifndef RRTC_CHUNK_SEARCHES
#define RRTC_CHUNK_SEARCHES 131072  // 32 768: 7.1 M queries/s on 2^18..2^20 queries, 131 072: 8.8-8.9 (arena 1.3 GB + rows <= 1 GB)
#endif
#ifndef RRTC_T2_ARENA_BYTES
#define RRTC_T2_ARENA_BYTES ((size_t)2 << 30)
#endif
#define RRTC_SMALL_QUERIES 64  // calls up to this size use host-mapped inputs / results: no copy operations at all
#define RRTC_MAX_REPLICAS 256
#define RRTC_NO_KEY 0xffffffffu

enum { PH_EXTEND = 0, PH_CONNECT = 1, PH_EXTRACT = 2, PH_SHORTCUT = 3, PH_DONE = 4 };
// how a search ended (s_status)
enum { ST_NONE = 0, ST_SOLVED = 1, ST_ITERCAP = 2, ST_NODECAP = 3, ST_PATHCAP = 4, ST_ABORTED = 5, ST_BADEND = 16 };
// what the collect kernel decided for a query (q_status)
enum { QS_FINAL = 1, QS_TIER2 = 2 };

struct RrtcArgs {
    const float* starts;  // [rows of the chunk][9]
    const float* goals;
    const int* qmap;      // tier 2: launch-local query -> row of the chunk (null = identity)
    int n_queries;        // queries of this launch
    float range, resolution;
    int max_iters, max_path, replicas, shortcut_passes, check_endpoints;
    int planner;          // 0 = RRTConnect (two trees), 1 = RRT (start tree only, 5 % goal bias: og.RRT defaults)
    int tree_nodes;       // M: capacity of one tree in THIS launch's arena (tier 1: RRTC_T1_NODES, tier 2: max_nodes)
    int path_rows;        // rows of one search's path buffer: min(max_path, 2 M)
    int tier1;            // 1: a node-cap failure may be an artefact of the small arena (collect flags the query)
    unsigned seed;
    unsigned query_base;  // global id of row 0 of the chunk: the RNG is keyed by the GLOBAL search id
    float* tree_q;        // [search][2][9][M]
    int* parent;          // [search][2][M]
    float* path_tmp;      // [search][path_rows][9]
    // per search (launch-local), consumed by the collect kernel
    int* s_status;
    int* s_iters;
    long long* s_checks;
    int* s_plen;
    unsigned* best_key;   // [rows]: min over the connected searches of a query of (iteration << 8 | replica)
    // per row of the chunk, written by the collect kernel (device or host-mapped memory)
    int* q_status;
    int* q_len;
    int* q_off;           // first row of the query's path in `rows`
    int* q_iters;
    long long* q_checks;
    float* rows;          // packed path states [sum of lengths][9]
    unsigned* cursor;     // rows handed out so far (device)
    unsigned* cursor_next;  // the other call parity's cursor: zeroed here so the next call needs no memset
    // a single query travels in the kernel parameters themselves (constant bank): nothing is read over PCIe before the
    // searches start
    int use_inline;
    float inline_q[18];
};

// nearest node of one tree (SoA [9][max_nodes]) to `t`: lanes stride over nodes, warp arg-min (ties -> lowest index)
__device__ __forceinline__ int rrtc_nearest(const float* __restrict__ tq, int size, int max_nodes, const float* t,
                                            int lane, float& best_d2) {
    float bd = 3.0e38f;
    int bi = 0x7fffffff;
    for (int i = lane; i < size; i += 32) {
        float d2 = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            float d = tq[(size_t)k * max_nodes + i] - t[k];
            d2 = fmaf(d, d, d2);
        }
        if (d2 < bd) {
            bd = d2;
            bi = i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float od = __shfl_xor_sync(0xffffffffu, bd, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < bd || (od == bd && oi < bi)) {
            bd = od;
            bi = oi;
        }
    }
    best_d2 = bd;
    return bi;
}

template <bool CARRY, bool YAW = false>
__global__ void __launch_bounds__(RRTC_THREADS, 3) pv_rrtc_kernel(const __grid_constant__ PvScene S,
                                                                   const __grid_constant__ RrtcArgs A) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int search = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (search >= A.n_queries * A.replicas) return;
    const int lq = search / A.replicas;
    const unsigned rep = (unsigned)(search - lq * A.replicas);
    const int row = A.qmap ? A.qmap[lq] : lq;
    const unsigned gsearch = (A.query_base + (unsigned)row) * (unsigned)A.replicas + rep;
    const int M = A.tree_nodes;
    float* tq = A.tree_q + (size_t)search * 2 * 9 * M;
    int* par = A.parent + (size_t)search * 2 * M;
    float* path = A.path_tmp + (size_t)search * A.path_rows * 9;
    float qs_[9], qg_[9];  // warp-uniform copies of the query (registers)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        qs_[k] = A.use_inline ? A.inline_q[k] : A.starts[(size_t)row * 9 + k];
        qg_[k] = A.use_inline ? A.inline_q[9 + k] : A.goals[(size_t)row * 9 + k];
    }
    if (lane < 9) {
        float vs = qs_[0], vg = qg_[0];
#pragma unroll
        for (int k = 1; k < 9; ++k) {
            vs = (lane == k) ? qs_[k] : vs;
            vg = (lane == k) ? qg_[k] : vg;
        }
        tq[(size_t)lane * M] = vs;
        tq[(size_t)(9 + lane) * M] = vg;
    }
    if (lane == 0) {
        par[0] = -1;
        par[M] = -1;
    }
    __syncwarp();

    int status = ST_NONE;
    int it = 0;
    long long n_checks = 0;
    int path_n = 0;

    if (A.check_endpoints) {
        // OMPL drops out-of-bounds / invalid start and goal states at intake (planning.py:163-187): lane 0 judges
        // the start, the other lanes the goal (the state check itself enforces the bounds)
        float qe[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) qe[k] = (lane == 0) ? qs_[k] : qg_[k];
        PvAcc<PV_MODE_BITS> acc0;
        pv_check_config<PV_MODE_BITS, true, PV_EXIT_NONE, 0, false, CARRY, false, false, YAW>(qe, S, acc0);
        const bool bad = acc0.hit;
        const int code = (__shfl_sync(FULL, bad ? 1 : 0, 0) ? 1 : 0) | (__shfl_sync(FULL, bad ? 1 : 0, 1) ? 2 : 0);
        if (code) status = ST_BADEND + code;
    }

    int size0 = 1, size1 = 1;  // tree sizes (warp-uniform)
    int cur = 0;               // tree grown by EXTEND in this iteration (0 = start tree)
    int phase = status ? PH_DONE : PH_EXTEND;
    float target[9];           // CONNECT target = state of the node just added by EXTEND
    int added_idx = 0;         // its index in tree `cur`
    int conn_idx = -1;         // node of the other tree that reached the target
    int sc_pass = 0, sc_i = 0, sc_j = 0, sc_budget = RRTC_MAX_SHORTCUT_CHECKS;

    while (phase != PH_DONE) {
        // OR-parallel replicas: give up once a published key beats every key this search can still reach (its key can
        // only grow with `it`), which never removes the eventual minimum -> the winner is independent of scheduling
        if (A.replicas > 1) {
            unsigned b = RRTC_NO_KEY;
            if (lane == 0) b = *((volatile unsigned*)(A.best_key + row));
            b = __shfl_sync(FULL, b, 0);
            if (b < (((unsigned)it << 8) | rep)) {
                status = ST_ABORTED;
                break;
            }
        }
        float ea[9], eb[9];
        int from_idx = 0, tree = 0;
        bool reach = true, aim_goal = false;
        if (phase == PH_EXTEND || phase == PH_CONNECT) {
            float goal_q[9];
            if (phase == PH_EXTEND) {
                if (it >= A.max_iters) {
                    status = ST_ITERCAP;
                    break;
                }
                if (size0 >= M - 1 || size1 >= M - 1) {
                    status = ST_NODECAP;
                    break;
                }
                tree = cur;
                float u9 = 1.f;
                if (it > 0) rrtc_sample(A.seed, gsearch, (unsigned)it, goal_q, u9);
                // first extension aims at the goal itself (cheap straight-line attempt); the single-tree planner also
                // does so with OMPL's default goal bias of 5 %
                aim_goal = (it == 0) || (A.planner == 1 && u9 < 0.05f);
                if (aim_goal) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) goal_q[k] = tq[(size_t)(9 + k) * M];
                }
            } else {
                tree = cur ^ 1;
#pragma unroll
                for (int k = 0; k < 9; ++k) goal_q[k] = target[k];
            }
            const float* tt = tq + (size_t)tree * 9 * M;
            float d2;
            from_idx = rrtc_nearest(tt, tree ? size1 : size0, M, goal_q, lane, d2);
            const float d = sqrtf(d2);
            float f = 1.0f;
            if (d > A.range) {
                f = A.range / d;
                reach = false;
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                ea[k] = tt[(size_t)k * M + from_idx];
                eb[k] = reach ? goal_q[k] : fmaf(f, goal_q[k] - ea[k], ea[k]);
            }
        } else if (phase == PH_EXTRACT) {
            // path = start-tree branch (root .. node) + goal-tree branch (parent of the duplicate .. root)
            int is = cur == 0 ? added_idx : conn_idx;
            int ig = cur == 0 ? conn_idx : added_idx;
            int ds = 0, dg = 0;
            for (int x = is; x >= 0; x = par[x]) ++ds;
            if (A.planner == 0)
                for (int x = par[M + ig]; x >= 0; x = par[M + x]) ++dg;
            path_n = ds + dg;
            if (path_n > A.max_path) {
                status = ST_PATHCAP;
                path_n = 0;
                break;
            }
            int x = is;
            for (int k = ds - 1; k >= 0; --k) {
                if (lane < 9) path[k * 9 + lane] = tq[(size_t)lane * M + x];
                x = par[x];
            }
            x = A.planner == 0 ? par[M + ig] : -1;
            for (int k = 0; k < dg; ++k) {
                if (lane < 9) path[(ds + k) * 9 + lane] = tq[(size_t)(9 + lane) * M + x];
                x = par[M + x];
            }
            __syncwarp();
            status = ST_SOLVED;
            if (A.replicas > 1 && lane == 0) atomicMin(A.best_key + row, ((unsigned)it << 8) | rep);
            sc_pass = 0;
            sc_i = 0;
            sc_j = path_n - 1;
            phase = (A.shortcut_passes > 0 && path_n > 2) ? PH_SHORTCUT : PH_DONE;
            continue;
        } else {  // PH_SHORTCUT: try to replace path[sc_i .. sc_j] by a straight segment, farthest first
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                ea[k] = path[sc_i * 9 + k];
                eb[k] = path[sc_j * 9 + k];
            }
        }

        // ---- motion validity of (ea -> eb): OMPL DiscreteMotionValidator restated, lanes = states ----------
        float de[9], d2 = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            de[k] = eb[k] - ea[k];
            d2 = fmaf(de[k], de[k], d2);
        }
        const int nd = max(1, (int)ceilf(sqrtf(d2) / A.resolution));
        const int rounds = (nd + 31) >> 5;
        const float inv_nd = 1.0f / (float)nd;
        bool hit = false;
        for (int r = 0; r < rounds; ++r) {
            int k = nd - (lane * rounds + r);
            if (k < 1) k = nd;
            const float t = (float)k * inv_nd;
            float q[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) q[c] = (k == nd) ? eb[c] : fmaf(t, de[c], ea[c]);
            PvAcc<PV_MODE_BITS> acc;
            if (pv_check_config<PV_MODE_BITS, true, PV_EXIT_ANY, 0, false, CARRY, false, (PV_COLD_SCENE_WARP != 0), YAW>(q, S, acc)) {
                PvReloadLerp rl;
#pragma unroll
                for (int c = 0; c < 9; ++c) {
                    rl.ea[c] = ea[c];
                    rl.eb[c] = eb[c];
                }
                rl.t = t;
                rl.at_end = (k == nd);
                acc.hit |= pv_scene_cold<true, PV_EXIT_ANY, false, CARRY, false>(rl, S);
            }
            n_checks += min(nd - r * 32, 32);
            if (__any_sync(FULL, acc.hit)) {
                hit = true;
                break;
            }
        }

        // ---- transitions ------------------------------------------------------------------------------------
        if (phase == PH_EXTEND || phase == PH_CONNECT) {
            int& sz = tree ? size1 : size0;
            if (!hit) {
                const int ni = sz;
                if (lane < 9) {
                    // per-lane select of eb[lane] without dynamic register indexing
                    float v = eb[0];
#pragma unroll
                    for (int k = 1; k < 9; ++k) v = (lane == k) ? eb[k] : v;
                    tq[(size_t)(tree * 9 + lane) * M + ni] = v;
                }
                if (lane == 0) par[tree * M + ni] = from_idx;
                __syncwarp();
                sz = ni + 1;
                if (phase == PH_EXTEND && A.planner == 1) {
                    // single tree: done when the goal itself was reached, else next sample
                    added_idx = ni;
                    if (reach && aim_goal) phase = PH_EXTRACT;
                    else ++it;
                } else if (phase == PH_EXTEND) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) target[k] = eb[k];
                    added_idx = ni;
                    phase = PH_CONNECT;
                } else if (reach) {
                    conn_idx = ni;
                    phase = PH_EXTRACT;
                } else if (sz >= M - 1) {
                    status = ST_NODECAP;
                    break;
                }
            } else {
                // TRAPPED: next iteration (RRTConnect swaps trees every iteration)
                phase = PH_EXTEND;
                if (A.planner == 0) cur ^= 1;
                ++it;
            }
        } else {  // PH_SHORTCUT
            --sc_budget;
            if (!hit) {
                // drop path[sc_i+1 .. sc_j-1]
                const int drop = sc_j - sc_i - 1;
                for (int k = sc_j; k < path_n; ++k) {
                    float v = 0.f;
                    if (lane < 9) v = path[k * 9 + lane];
                    __syncwarp();
                    if (lane < 9) path[(k - drop) * 9 + lane] = v;
                }
                __syncwarp();
                path_n -= drop;
                ++sc_i;
                sc_j = path_n - 1;
            } else {
                --sc_j;
            }
            if (sc_j < sc_i + 2) {
                ++sc_i;
                sc_j = path_n - 1;
            }
            if (sc_i + 2 >= path_n + 0 && sc_j < sc_i + 2) {
                ++sc_pass;
                sc_i = 0;
                sc_j = path_n - 1;
            }
            if (sc_pass >= A.shortcut_passes || path_n <= 2 || sc_budget <= 0) phase = PH_DONE;
        }
    }

    if (lane == 0) {
        A.s_status[search] = status;
        A.s_iters[search] = status == ST_SOLVED ? it + 1 : it;
        A.s_checks[search] = n_checks;
        A.s_plen[search] = status == ST_SOLVED ? path_n : 0;
    }
}

// One warp per query, after all its searches have ended: pick the winner (smallest (iterations, replica) among the
// connected searches), decide whether the small tier-1 arena may have changed the answer, hand out rows of the packed
// path buffer and copy the winner's path there.
__global__ void __launch_bounds__(128) pv_rrtc_collect_kernel(const __grid_constant__ RrtcArgs A) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int lq = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (lq == 0 && lane == 0 && A.cursor_next) *A.cursor_next = 0u;
    if (lq >= A.n_queries) return;
    const int row = A.qmap ? A.qmap[lq] : lq;
    unsigned best = RRTC_NO_KEY, capkey = RRTC_NO_KEY;
    long long sum_checks = 0;
    int max_it = 0;
    for (int rep = lane; rep < A.replicas; rep += 32) {
        const int s = lq * A.replicas + rep;
        const int st = A.s_status[s], si = A.s_iters[s];
        if (st == ST_SOLVED) best = min(best, ((unsigned)(si - 1) << 8) | (unsigned)rep);
        if (st == ST_NODECAP) capkey = min(capkey, ((unsigned)si << 8) | (unsigned)rep);
        sum_checks += A.s_checks[s];
        max_it = max(max_it, si);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        best = min(best, __shfl_xor_sync(FULL, best, o));
        capkey = min(capkey, __shfl_xor_sync(FULL, capkey, o));
        sum_checks += __shfl_xor_sync(FULL, sum_checks, o);
        max_it = max(max_it, __shfl_xor_sync(FULL, max_it, o));
    }
    const int st0 = A.s_status[lq * A.replicas];
    if (lane == 0) A.best_key[row] = RRTC_NO_KEY;  // left clean for the next launch that uses this row
    if (st0 >= ST_BADEND) {
        if (lane == 0) {
            A.q_status[row] = QS_FINAL;
            A.q_len[row] = 0;
            A.q_off[row] = 0;
            A.q_iters[row] = -(st0 - ST_BADEND);
            A.q_checks[row] = 0;
        }
        return;
    }
    if (A.tier1 && capkey < best) {
        // a search ran out of its SMALL trees at a point where it could still have become the winner: plan the query
        // again with full-size trees
        if (lane == 0) A.q_status[row] = QS_TIER2;
        return;
    }
    if (best == RRTC_NO_KEY) {
        if (lane == 0) {
            A.q_status[row] = QS_FINAL;
            A.q_len[row] = 0;
            A.q_off[row] = 0;
            A.q_iters[row] = max_it;
            A.q_checks[row] = sum_checks;
        }
        return;
    }
    const int s = lq * A.replicas + (int)(best & 255u);
    const int len = A.s_plen[s];
    unsigned off = 0;
    if (lane == 0) off = atomicAdd(A.cursor, (unsigned)len);
    off = __shfl_sync(FULL, off, 0);
    const float* src = A.path_tmp + (size_t)s * A.path_rows * 9;
    float* dst = A.rows + (size_t)off * 9;
    for (int k = lane; k < len * 9; k += 32) dst[k] = src[k];
    if (lane == 0) {
        A.q_status[row] = QS_FINAL;
        A.q_len[row] = len;
        A.q_off[row] = (int)off;
        A.q_iters[row] = A.s_iters[s];
        A.q_checks[row] = A.s_checks[s];
    }
}

// =========================================================================================================
// host side
// =========================================================================================================
#define RR_CUDA(expr)                                                                                        \
    do {                                                                                                     \
        cudaError_t e_ = (expr);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            snprintf(h->err, sizeof(h->err), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                     __LINE__);                                                                              \
            return PV_ERR_CUDA;                                                                              \
        }                                                                                                    \
    } while (0)

static inline size_t rr_al(size_t x) { return (x + 255) & ~(size_t)255; }

// grow-only device / pinned host buffers of the handle
static int rr_reserve_dev(PvHandle* h, void** buf, size_t* have, size_t need) {
    if (need <= *have) return PV_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
    RR_CUDA(cudaMalloc(buf, need));
    *have = need;
    return PV_OK;
}
static int rr_reserve_host(PvHandle* h, void** buf, size_t* have, size_t need) {
    if (need <= *have) return PV_OK;
    if (*buf) cudaFreeHost(*buf);
    *buf = nullptr;
    *have = 0;
    RR_CUDA(cudaHostAlloc(buf, need, cudaHostAllocMapped));
    *have = need;
    return PV_OK;
}

int pv_rrtc_validate(PvHandle* h, const PvRrtcParams* p, RrtcArgs* a) {
    memset(a, 0, sizeof(*a));
    a->range = p->range > 0.f ? p->range : PV_RRTC_RANGE;
    a->resolution = p->resolution > 0.f ? p->resolution : PV_VALIDITY_RESOLUTION;
    a->max_iters = p->max_iters == 0 ? 2000 : p->max_iters;
    const int max_nodes = p->max_nodes == 0 ? 2048 : p->max_nodes;
    a->max_path = p->max_path == 0 ? 128 : p->max_path;
    a->replicas = p->replicas == 0 ? 1 : p->replicas;
    a->shortcut_passes = p->shortcut_passes;
    // out-of-range capacities are an error, not a silent default: a caller who sized its path buffer from
    // its own max_path must never be written beyond it
    if (a->max_iters < 1 || a->max_iters >= (1 << 23) || max_nodes < 8 || max_nodes > (1 << 22) || a->max_path < 2 ||
        a->max_path > (1 << 20) || a->replicas < 1 || a->replicas > RRTC_MAX_REPLICAS || a->shortcut_passes < 0 ||
        p->query_offset < 0 || (p->planner != 0 && p->planner != 1)) {
        snprintf(h->err, sizeof(h->err),
                 "pv_rrtc_batch: parameter out of range (max_iters 1..2^23-1, max_nodes 8..2^22, max_path 2..2^20, "
                 "replicas 1..%d, shortcut_passes >= 0, query_offset >= 0, planner 0|1; 0 selects the default of the "
                 "first four)", RRTC_MAX_REPLICAS);
        return PV_ERR_BAD_ARG;
    }
    a->check_endpoints = p->check_endpoints ? 1 : 0;
    a->planner = p->planner;
    a->seed = p->seed;
    a->tree_nodes = max_nodes;  // the caller's capacity; the launches below set the tier's own
    return PV_OK;
}

// Plans queries [0, n) (host AoS rows).  `sink(first, count, len, off, iters, checks, rows)` receives each chunk's
// results: per-query arrays indexed from the chunk's first query and the packed path rows of the chunk.
// `after_first_launch` (optional) runs right after the first chunk's kernels are queued and before the host waits for
// them: pv_plan_path queues its speculative straight-line validation there, so an easy plan costs ONE synchronisation.
int pv_rrtc_run(PvHandle* h, const float* h_starts, const float* h_goals, int n, const PvRrtcParams* params,
                const PvRrtcSink& sink, const std::function<void(cudaStream_t)>* after_first_launch) {
    RrtcArgs base;
    int rc = pv_rrtc_validate(h, params, &base);
    if (rc) return rc;
    const int max_nodes = base.tree_nodes;
    const int R = base.replicas;
    cudaStream_t st = h->streams[0];

    const int t1_nodes = max_nodes < RRTC_T1_NODES ? max_nodes : RRTC_T1_NODES;
    const bool two_tier = t1_nodes < max_nodes;
    auto path_rows_for = [&](int M) { return base.max_path < 2 * M ? base.max_path : 2 * M; };
    int chunk_q = RRTC_CHUNK_SEARCHES / R;
    // the packed-row buffer of a chunk (worst case: every path at full length) stays below 1 GB
    const size_t rows_budget = ((size_t)1 << 30) / ((size_t)path_rows_for(max_nodes) * 9 * sizeof(float));
    if ((size_t)chunk_q > rows_budget) chunk_q = (int)rows_budget;
    if (chunk_q < 1) chunk_q = 1;
    if (chunk_q > n) chunk_q = n;
    const bool small = n <= RRTC_SMALL_QUERIES && (size_t)n * path_rows_for(max_nodes) * 9 * sizeof(float) <= ((size_t)4 << 20);

    auto arena_per_search = [&](int M) {
        return (size_t)2 * 9 * M * sizeof(float) + (size_t)2 * M * sizeof(int) + (size_t)path_rows_for(M) * 9 * sizeof(float);
    };
    // tier-2 sub-chunks: as many queries as fit the arena budget (at least one)
    size_t t2_q = two_tier ? RRTC_T2_ARENA_BYTES / (arena_per_search(max_nodes) * R) : 0;
    if (two_tier && t2_q < 1) t2_q = 1;
    if (t2_q > (size_t)chunk_q) t2_q = chunk_q;

    // ---- carve the handle's grow-only buffers ---------------------------------------------------------------
    const size_t n_s1 = (size_t)chunk_q * R, n_s2 = t2_q * R;
    const size_t n_s = n_s1 > n_s2 ? n_s1 : n_s2;
    size_t arena = n_s1 * arena_per_search(t1_nodes);
    if (n_s2 * arena_per_search(max_nodes) > arena) arena = n_s2 * arena_per_search(max_nodes);
    const size_t rows_cap = (size_t)chunk_q * (size_t)path_rows_for(max_nodes);
    const size_t b_q9 = rr_al((size_t)chunk_q * 9 * sizeof(float));
    const size_t b_qi = rr_al((size_t)chunk_q * sizeof(int));
    const size_t b_ql = rr_al((size_t)chunk_q * sizeof(long long));
    const size_t b_si = rr_al(n_s * sizeof(int));
    const size_t b_sl = rr_al(n_s * sizeof(long long));
    const size_t b_meta = 4 * b_qi + b_ql;  // q_status, q_len, q_off, q_iters | q_checks
    const size_t b_rows = rr_al(rows_cap * 9 * sizeof(float));
    // counters and the per-row keys sit at FIXED offsets at the front of the buffer (their "left clean" invariant must
    // survive calls of different sizes, which carve the rest differently)
    const size_t b_ctr = 256;
    const size_t b_keys = rr_al((size_t)RRTC_CHUNK_SEARCHES * sizeof(unsigned));
    const size_t dev_need = b_ctr + b_keys + rr_al(arena + 256) + 2 * b_q9 + b_qi /*qmap*/ + 3 * b_si + b_sl +
                            (small ? 0 : b_meta + b_rows);
    const bool fresh = dev_need > h->rrtc_bytes;
    rc = rr_reserve_dev(h, &h->rrtc_buf, &h->rrtc_bytes, dev_need);
    if (rc) return rc;
    const size_t host_need = 2 * b_q9 + b_meta + (small ? b_rows : 0) + b_qi;
    rc = rr_reserve_host(h, &h->rrtc_host, &h->rrtc_host_bytes, host_need);
    if (rc) return rc;
    // the packed rows of a large chunk come back in a second copy whose size is only known after the first
    if (!small) {
        rc = rr_reserve_host(h, &h->rrtc_rows_host, &h->rrtc_rows_host_bytes, (size_t)1 << 20);
        if (rc) return rc;
    }

    char* p = (char*)h->rrtc_buf;
    unsigned* ctr = (unsigned*)p; p += b_ctr;  // [0], [1]: row cursors of even / odd launches
    base.best_key = (unsigned*)p; p += b_keys;
    char* arena_p = p; p += rr_al(arena + 256);
    float* d_starts = (float*)p; p += b_q9;
    float* d_goals = (float*)p; p += b_q9;
    int* d_qmap = (int*)p; p += b_qi;
    base.s_status = (int*)p; p += b_si;
    base.s_iters = (int*)p; p += b_si;
    base.s_plen = (int*)p; p += b_si;
    base.s_checks = (long long*)p; p += b_sl;
    char* d_meta = p;
    float* d_rows = nullptr;
    if (!small) {
        p += b_meta;
        d_rows = (float*)p; p += b_rows;
    }
    char* hp = (char*)h->rrtc_host;
    float* hm_starts = (float*)hp; hp += b_q9;
    float* hm_goals = (float*)hp; hp += b_q9;
    char* h_meta = hp; hp += b_meta;
    int* h_qmap = (int*)hp; hp += b_qi;
    float* hm_rows = small ? (float*)hp : nullptr;

    if (fresh) {
        // a new arena: counters and keys start clean; from then on the collect kernel leaves them clean
        RR_CUDA(cudaMemsetAsync(ctr, 0, b_ctr, st));
        RR_CUDA(cudaMemsetAsync(base.best_key, 0xFF, b_keys, st));
        h->rrtc_parity = 0;
    }
    auto meta_ptrs = [&](RrtcArgs& a, char* m) {
        a.q_status = (int*)m;
        a.q_len = (int*)(m + b_qi);
        a.q_off = (int*)(m + 2 * b_qi);
        a.q_iters = (int*)(m + 3 * b_qi);
        a.q_checks = (long long*)(m + 4 * b_qi);
    };
    auto carve_arena = [&](RrtcArgs& a, int M, size_t n_search) {
        char* q = arena_p;
        a.tree_nodes = M;
        a.path_rows = path_rows_for(M);
        a.tree_q = (float*)q; q += rr_al(n_search * 2 * 9 * M * sizeof(float));
        a.parent = (int*)q; q += rr_al(n_search * 2 * M * sizeof(int));
        a.path_tmp = (float*)q;
    };
    auto launch = [&](const RrtcArgs& a) {
        const size_t n_search = (size_t)a.n_queries * R;
        const int wpb = RRTC_THREADS / 32;
        const int grid = (int)((n_search + wpb - 1) / wpb);
        if (h->scene.carry) pv_rrtc_kernel<true><<<grid, RRTC_THREADS, 0, st>>>(h->scene, a);
        else if (h->all_yaw) pv_rrtc_kernel<false, true><<<grid, RRTC_THREADS, 0, st>>>(h->scene, a);
        else pv_rrtc_kernel<false><<<grid, RRTC_THREADS, 0, st>>>(h->scene, a);
        pv_rrtc_collect_kernel<<<(a.n_queries + 3) / 4, 128, 0, st>>>(a);
        h->launches += 2;
    };

    for (int first = 0; first < n; first += chunk_q) {
        const int nq = n - first < chunk_q ? n - first : chunk_q;
        RrtcArgs a = base;
        a.n_queries = nq;
        a.query_base = (unsigned)params->query_offset + (unsigned)first;
        a.qmap = nullptr;
        a.tier1 = two_tier ? 1 : 0;
        const int par = h->rrtc_parity & 1;
        h->rrtc_parity ^= 1;
        a.cursor = ctr + par;
        a.cursor_next = ctr + (par ^ 1);
        carve_arena(a, t1_nodes, (size_t)nq * R);
        memcpy(hm_starts, h_starts + (size_t)first * 9, (size_t)nq * 9 * sizeof(float));
        memcpy(hm_goals, h_goals + (size_t)first * 9, (size_t)nq * 9 * sizeof(float));
        if (small) {
            // zero-copy: the kernels read the queries from, and write the results to, host-mapped pinned memory
            a.starts = hm_starts;
            a.goals = hm_goals;
            if (n == 1) {
                a.use_inline = 1;
                memcpy(a.inline_q, hm_starts, 9 * sizeof(float));
                memcpy(a.inline_q + 9, hm_goals, 9 * sizeof(float));
            }
            meta_ptrs(a, h_meta);
            a.rows = hm_rows;
        } else {
            RR_CUDA(cudaMemcpyAsync(d_starts, hm_starts, (size_t)nq * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
            RR_CUDA(cudaMemcpyAsync(d_goals, hm_goals, (size_t)nq * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
            a.starts = d_starts;
            a.goals = d_goals;
            meta_ptrs(a, d_meta);
            a.rows = d_rows;
        }
        launch(a);
        RR_CUDA(cudaGetLastError());
        if (first == 0 && after_first_launch && *after_first_launch) (*after_first_launch)(st);
        if (!small) RR_CUDA(cudaMemcpyAsync(h_meta, d_meta, b_meta, cudaMemcpyDeviceToHost, st));
        RR_CUDA(cudaStreamSynchronize(st));
        RrtcArgs hm;  // host view of the meta block
        meta_ptrs(hm, h_meta);

        if (two_tier) {
            int n2 = 0;
            for (int k = 0; k < nq; ++k)
                if (hm.q_status[k] == QS_TIER2) h_qmap[n2++] = k;
            if (n2 > 0) {
                for (int f2 = 0; f2 < n2; f2 += (int)t2_q) {
                    const int m2 = n2 - f2 < (int)t2_q ? n2 - f2 : (int)t2_q;
                    RrtcArgs b = a;
                    b.n_queries = m2;
                    b.tier1 = 0;
                    b.cursor_next = nullptr;  // the same call parity keeps handing out rows behind tier 1's
                    RR_CUDA(cudaMemcpyAsync(d_qmap, h_qmap + f2, (size_t)m2 * sizeof(int), cudaMemcpyHostToDevice, st));
                    b.qmap = d_qmap;
                    carve_arena(b, max_nodes, (size_t)m2 * R);
                    launch(b);
                    RR_CUDA(cudaGetLastError());
                    // d_qmap and the arena are reused by the next sub-chunk
                    RR_CUDA(cudaStreamSynchronize(st));
                }
                if (!small) {
                    RR_CUDA(cudaMemcpyAsync(h_meta, d_meta, b_meta, cudaMemcpyDeviceToHost, st));
                    RR_CUDA(cudaStreamSynchronize(st));
                }
            }
        }
        const float* rows = hm_rows;
        if (!small) {
            size_t total = 0;
            for (int k = 0; k < nq; ++k) {
                const size_t end = (size_t)hm.q_off[k] + (size_t)hm.q_len[k];
                if (hm.q_len[k] > 0 && end > total) total = end;
            }
            if (total > 0) {
                rc = rr_reserve_host(h, &h->rrtc_rows_host, &h->rrtc_rows_host_bytes, total * 9 * sizeof(float));
                if (rc) return rc;
                RR_CUDA(cudaMemcpyAsync(h->rrtc_rows_host, d_rows, total * 9 * sizeof(float), cudaMemcpyDeviceToHost, st));
                RR_CUDA(cudaStreamSynchronize(st));
            }
            rows = (const float*)h->rrtc_rows_host;
        }
        sink(first, nq, hm.q_len, hm.q_off, hm.q_iters, hm.q_checks, rows);
    }
    return PV_OK;
}

static int rr_check_call(PvHandle* h, int n_queries, const PvRrtcParams* params, bool ptrs_ok) {
    if (!h || h->magic != PV_HANDLE_MAGIC) return PV_ERR_BAD_HANDLE;
    if (!h->has_scene) {
        snprintf(h->err, sizeof(h->err), "no scene set (pv_set_scene)");
        return PV_ERR_NO_SCENE;
    }
    if (n_queries < 0 || !params || (n_queries > 0 && !ptrs_ok)) {
        snprintf(h->err, sizeof(h->err), "pv_rrtc_batch: bad arguments");
        return PV_ERR_BAD_ARG;
    }
    return PV_OK;
}

extern "C" int pv_rrtc_batch(PvHandle* h, const float* h_starts, const float* h_goals, int n_queries,
                             const PvRrtcParams* params, float* h_path_out, int* h_path_len, int* h_iters,
                             long long* h_checks) {
    int rc = rr_check_call(h, n_queries, params, h_starts && h_goals && h_path_out && h_path_len);
    if (rc) return rc;
    if (n_queries == 0) {
        RrtcArgs tmp;
        return pv_rrtc_validate(h, params, &tmp);
    }
    const size_t row = (size_t)(params->max_path == 0 ? 128 : params->max_path) * 9;
    PvDeviceGuard guard(h->device);
    // only the used rows of each path are written: rows beyond h_path_len[k] of the caller's buffer stay as they were
    return pv_rrtc_run(h, h_starts, h_goals, n_queries, params,
                       [&](int first, int nq, const int* len, const int* off, const int* iters, const long long* checks,
                           const float* rows) {
                           for (int k = 0; k < nq; ++k) {
                               h_path_len[first + k] = len[k];
                               if (h_iters) h_iters[first + k] = iters[k];
                               if (h_checks) h_checks[first + k] = checks[k];
                               if (len[k] > 0)
                                   memcpy(h_path_out + (size_t)(first + k) * row, rows + (size_t)off[k] * 9,
                                          (size_t)len[k] * 9 * sizeof(float));
                           }
                       },
                       nullptr);
}

extern "C" int pv_rrtc_batch_packed(PvHandle* h, const float* h_starts, const float* h_goals, int n_queries,
                                    const PvRrtcParams* params, float* h_states, long long state_capacity,
                                    long long* h_path_off, int* h_path_len, int* h_iters, long long* h_checks,
                                    long long* n_states) {
    int rc = rr_check_call(h, n_queries, params, h_starts && h_goals && h_path_off && h_path_len && (h_states || state_capacity == 0));
    if (rc) return rc;
    if (state_capacity < 0) return PV_ERR_BAD_ARG;
    long long used = 0;
    bool overflow = false;
    if (n_queries > 0) {
        PvDeviceGuard guard(h->device);
        rc = pv_rrtc_run(h, h_starts, h_goals, n_queries, params,
                         [&](int first, int nq, const int* len, const int* off, const int* iters, const long long* checks,
                             const float* rows) {
                             for (int k = 0; k < nq; ++k) {
                                 h_path_len[first + k] = len[k];
                                 h_path_off[first + k] = used;
                                 if (h_iters) h_iters[first + k] = iters[k];
                                 if (h_checks) h_checks[first + k] = checks[k];
                                 if (len[k] > 0) {
                                     if (used + len[k] <= state_capacity)
                                         memcpy(h_states + (size_t)used * 9, rows + (size_t)off[k] * 9,
                                                (size_t)len[k] * 9 * sizeof(float));
                                     else
                                         overflow = true;
                                     used += len[k];
                                 }
                             }
                         },
                         nullptr);
        if (rc) return rc;
    } else {
        RrtcArgs tmp;
        rc = pv_rrtc_validate(h, params, &tmp);
        if (rc) return rc;
    }
    if (n_states) *n_states = used;
    if (overflow) {
        snprintf(h->err, sizeof(h->err), "pv_rrtc_batch_packed: %lld path states do not fit the capacity of %lld "
                 "(lengths and offsets are complete; call again with a larger buffer)", used, state_capacity);
        return PV_ERR_CAPACITY;
    }
    return PV_OK;
}
