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
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/panda_validity.h"
#include "pv_device.cuh"
#include "pv_handle.h"

#define RRTC_THREADS 128
#define RRTC_MAX_SHORTCUT_CHECKS 96

enum { PH_EXTEND = 0, PH_CONNECT = 1, PH_EXTRACT = 2, PH_SHORTCUT = 3, PH_DONE = 4 };

struct RrtcArgs {
    const float* starts;  // [nq][9]
    const float* goals;   // [nq][9]
    int n_queries;
    float range, resolution;
    int max_iters, max_nodes, max_path, replicas, shortcut_passes, check_endpoints;
    int planner;  // 0 = RRTConnect (two trees), 1 = RRT (start tree only, 5 % goal bias: og.RRT defaults)
    unsigned seed;
    unsigned search_offset;  // query_offset * replicas: the RNG is keyed by the GLOBAL search id
    float* tree_q;      // [search][2][9][max_nodes]
    int* parent;        // [search][2][max_nodes]
    float* path_tmp;    // [search][max_path][9]
    float* path_out;    // [nq][max_path][9]
    int* path_len;      // [nq]
    int* iters_out;     // [nq]
    long long* checks;  // [nq]
    int* winner;        // [nq], -1 until a search of that query finishes
};

__device__ __forceinline__ void rrtc_sample(unsigned seed, unsigned search, unsigned it, float* q, float& extra) {
    const float lo[9] = PV_Q_LOWER, hi[9] = PV_Q_UPPER;
    float u[12];
    const uint2 key = make_uint2(seed, 0x52525443u);
#pragma unroll
    for (int blk = 0; blk < 3; ++blk) {
        uint4 r = pv_philox(make_uint4(it, search, blk, 1u), key);
        u[4 * blk + 0] = (float)(r.x >> 8) * 5.9604644775390625e-08f;
        u[4 * blk + 1] = (float)(r.y >> 8) * 5.9604644775390625e-08f;
        u[4 * blk + 2] = (float)(r.z >> 8) * 5.9604644775390625e-08f;
        u[4 * blk + 3] = (float)(r.w >> 8) * 5.9604644775390625e-08f;
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) q[j] = __fmaf_rn(u[j], hi[j] - lo[j], lo[j]);
    extra = u[9];  // a tenth uniform draw of the same counter: goal bias of the single-tree planner
}

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

template <bool CARRY>
__global__ void __launch_bounds__(RRTC_THREADS, 3) pv_rrtc_kernel(const __grid_constant__ PvScene S,
                                                                   const __grid_constant__ RrtcArgs A) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int search = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (search >= A.n_queries * A.replicas) return;
    const int query = search / A.replicas;
    const int M = A.max_nodes;
    float* tq = A.tree_q + (size_t)search * 2 * 9 * M;
    int* par = A.parent + (size_t)search * 2 * M;
    float* path = A.path_tmp + (size_t)search * A.max_path * 9;

    if (lane < 9) {
        tq[(size_t)lane * M] = A.starts[query * 9 + lane];
        tq[(size_t)(9 + lane) * M] = A.goals[query * 9 + lane];
    }
    if (lane == 0) {
        par[0] = -1;
        par[M] = -1;
    }
    __syncwarp();

    if (A.check_endpoints) {
        // OMPL drops out-of-bounds / invalid start and goal states at intake (planning.py:163-187): lane 0 judges
        // the start, the other lanes the goal
        const float lo[9] = PV_Q_LOWER, hi[9] = PV_Q_UPPER;
        float qe[9];
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            qe[k] = (lane == 0) ? A.starts[query * 9 + k] : A.goals[query * 9 + k];
            bad |= (qe[k] < lo[k]) || (qe[k] > hi[k]);
        }
        PvAcc<PV_MODE_BITS> acc0;
        pv_check_config<PV_MODE_BITS, true, PV_EXIT_NONE, 0, false, CARRY>(qe, S, acc0);
        bad |= acc0.hit;
        const int code = (__shfl_sync(FULL, bad ? 1 : 0, 0) ? 1 : 0) | (__shfl_sync(FULL, bad ? 1 : 0, 1) ? 2 : 0);
        if (code) {
            if (lane == 0 && search % A.replicas == 0) {
                A.iters_out[query] = -code;
                A.path_len[query] = 0;
            }
            return;
        }
    }

    int size0 = 1, size1 = 1;  // tree sizes (warp-uniform)
    int cur = 0;               // tree grown by EXTEND in this iteration (0 = start tree)
    int it = 0;
    int phase = PH_EXTEND;
    long long n_checks = 0;
    float target[9];           // CONNECT target = state of the node just added by EXTEND
    int added_idx = 0;         // its index in tree `cur`
    int conn_idx = -1;         // node of the other tree that reached the target
    int path_n = 0;
    int sc_pass = 0, sc_i = 0, sc_j = 0, sc_budget = RRTC_MAX_SHORTCUT_CHECKS;
    bool solved = false;

    while (phase != PH_DONE) {
        // another replica of this query already finished
        if (A.replicas > 1 && phase <= PH_CONNECT) {
            int w = 0;
            if (lane == 0) w = *((volatile int*)(A.winner + query));
            if (__shfl_sync(FULL, w, 0) != -1) return;
        }
        float ea[9], eb[9];
        int from_idx = 0, tree = 0;
        bool reach = true, aim_goal = false;
        if (phase == PH_EXTEND || phase == PH_CONNECT) {
            float goal_q[9];
            if (phase == PH_EXTEND) {
                if (it >= A.max_iters || size0 >= M - 1 || size1 >= M - 1) break;
                tree = cur;
                float u9 = 1.f;
                if (it > 0) rrtc_sample(A.seed, (unsigned)search + A.search_offset, (unsigned)it, goal_q, u9);
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
                solved = false;
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
            solved = true;
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
            pv_check_config<PV_MODE_BITS, true, PV_EXIT_ANY, 0, false, CARRY>(q, S, acc);
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

    if (!solved) {
        // report the effort of failed searches only if no replica succeeds (winner stays -1)
        if (A.replicas == 1 && lane == 0) {
            A.iters_out[query] = it;
            A.checks[query] = n_checks;
        }
        return;
    }
    int won = 0;
    if (lane == 0) won = (atomicCAS(A.winner + query, -1, search) == -1);
    won = __shfl_sync(FULL, won, 0);
    if (!won) return;
    float* po = A.path_out + (size_t)query * A.max_path * 9;
    for (int k = lane; k < path_n * 9; k += 32) po[k] = path[k];
    if (lane == 0) {
        A.path_len[query] = path_n;
        A.iters_out[query] = it + 1;
        A.checks[query] = n_checks;
    }
}

// =========================================================================================================
extern "C" int pv_rrtc_batch(PvHandle* h, const float* h_starts, const float* h_goals, int n_queries,
                             const PvRrtcParams* params, float* h_path_out, int* h_path_len, int* h_iters,
                             long long* h_checks) {
    if (!h || h->magic != PV_HANDLE_MAGIC) return PV_ERR_BAD_HANDLE;
    if (!h->has_scene) {
        snprintf(h->err, sizeof(h->err), "no scene set (pv_set_scene)");
        return PV_ERR_NO_SCENE;
    }
    if (n_queries < 0 || !params || (n_queries > 0 && (!h_starts || !h_goals || !h_path_out || !h_path_len))) {
        snprintf(h->err, sizeof(h->err), "pv_rrtc_batch: bad arguments");
        return PV_ERR_BAD_ARG;
    }
    if (n_queries == 0) return PV_OK;
    RrtcArgs a;
    memset(&a, 0, sizeof(a));
    a.n_queries = n_queries;
    a.range = params->range > 0.f ? params->range : PV_RRTC_RANGE;
    a.resolution = params->resolution > 0.f ? params->resolution : PV_VALIDITY_RESOLUTION;
    a.max_iters = params->max_iters > 0 ? params->max_iters : 2000;
    a.max_nodes = params->max_nodes >= 8 ? params->max_nodes : 2048;
    a.max_path = params->max_path >= 2 ? params->max_path : 128;
    a.replicas = params->replicas >= 1 ? params->replicas : 1;
    a.shortcut_passes = params->shortcut_passes >= 0 ? params->shortcut_passes : 0;
    a.check_endpoints = params->check_endpoints ? 1 : 0;
    a.planner = params->planner == 1 ? 1 : 0;
    a.seed = params->seed;
    a.search_offset = (unsigned)(params->query_offset > 0 ? params->query_offset : 0) * (unsigned)a.replicas;
    const size_t n_search = (size_t)n_queries * a.replicas;

#define RR_CUDA(expr)                                                                                        \
    do {                                                                                                     \
        cudaError_t e_ = (expr);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            snprintf(h->err, sizeof(h->err), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                     __LINE__);                                                                              \
            return PV_ERR_CUDA;                                                                              \
        }                                                                                                    \
    } while (0)
    RR_CUDA(cudaSetDevice(h->device));

    // one grow-only arena in the handle, carved up per call
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_sg = al((size_t)n_queries * 9 * sizeof(float));
    const size_t b_tree = al(n_search * 2 * 9 * a.max_nodes * sizeof(float));
    const size_t b_par = al(n_search * 2 * a.max_nodes * sizeof(int));
    const size_t b_ptmp = al(n_search * a.max_path * 9 * sizeof(float));
    const size_t b_pout = al((size_t)n_queries * a.max_path * 9 * sizeof(float));
    const size_t b_i = al((size_t)n_queries * sizeof(int));
    const size_t b_ll = al((size_t)n_queries * sizeof(long long));
    const size_t total = 2 * b_sg + b_tree + b_par + b_ptmp + b_pout + 3 * b_i + b_ll;
    if (total > h->rrtc_bytes) {
        if (h->rrtc_buf) cudaFree(h->rrtc_buf);
        h->rrtc_buf = nullptr;
        h->rrtc_bytes = 0;
        RR_CUDA(cudaMalloc(&h->rrtc_buf, total));
        h->rrtc_bytes = total;
    }
    char* p = (char*)h->rrtc_buf;
    float* d_starts = (float*)p; p += b_sg;
    float* d_goals = (float*)p; p += b_sg;
    a.tree_q = (float*)p; p += b_tree;
    a.parent = (int*)p; p += b_par;
    a.path_tmp = (float*)p; p += b_ptmp;
    // results are contiguous so that one memset clears them and one D2H fetches them
    char* res0 = p;
    a.path_len = (int*)p; p += b_i;
    a.iters_out = (int*)p; p += b_i;
    a.checks = (long long*)p; p += b_ll;
    a.path_out = (float*)p; p += b_pout;
    const size_t res_bytes = (size_t)(p - res0);
    a.winner = (int*)p; p += b_i;
    a.starts = d_starts;
    a.goals = d_goals;

    // pinned host mirror of the result block (grow-only)
    if (res_bytes > h->rrtc_host_bytes) {
        if (h->rrtc_host) cudaFreeHost(h->rrtc_host);
        h->rrtc_host = nullptr;
        h->rrtc_host_bytes = 0;
        RR_CUDA(cudaMallocHost(&h->rrtc_host, res_bytes));
        h->rrtc_host_bytes = res_bytes;
    }

    cudaStream_t st = h->streams[0];
    RR_CUDA(cudaMemcpyAsync(d_starts, h_starts, (size_t)n_queries * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    RR_CUDA(cudaMemcpyAsync(d_goals, h_goals, (size_t)n_queries * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    RR_CUDA(cudaMemsetAsync(a.path_len, 0, 2 * b_i + b_ll, st));  // path_len, iters, checks
    RR_CUDA(cudaMemsetAsync(a.winner, 0xFF, b_i, st));
    const int warps_per_block = RRTC_THREADS / 32;
    const int grid = (int)((n_search + warps_per_block - 1) / warps_per_block);
    if (h->scene.carry) pv_rrtc_kernel<true><<<grid, RRTC_THREADS, 0, st>>>(h->scene, a);
    else pv_rrtc_kernel<false><<<grid, RRTC_THREADS, 0, st>>>(h->scene, a);
    h->launches++;
    RR_CUDA(cudaGetLastError());
    // small batches: one packed copy; large batches: skip the unused tail of each path? (paths are max_path long)
    RR_CUDA(cudaMemcpyAsync(h->rrtc_host, res0, res_bytes, cudaMemcpyDeviceToHost, st));
    RR_CUDA(cudaStreamSynchronize(st));
    const char* hp = (const char*)h->rrtc_host;
    memcpy(h_path_len, hp, (size_t)n_queries * sizeof(int));
    if (h_iters) memcpy(h_iters, hp + b_i, (size_t)n_queries * sizeof(int));
    if (h_checks) memcpy(h_checks, hp + 2 * b_i, (size_t)n_queries * sizeof(long long));
    // only the used rows of each path: rows beyond h_path_len[k] of the caller's buffer are left as they were (the
    // device rows there are scratch of the search), and a large, freshly allocated buffer is not paged in for nothing
    const float* hpaths = (const float*)(hp + 2 * b_i + b_ll);
    const size_t row = (size_t)a.max_path * 9;
    for (int k = 0; k < n_queries; ++k) {
        const int len = h_path_len[k] < a.max_path ? h_path_len[k] : a.max_path;
        if (len > 0) memcpy(h_path_out + k * row, hpaths + k * row, (size_t)len * 9 * sizeof(float));
    }
    return PV_OK;
#undef RR_CUDA
}
