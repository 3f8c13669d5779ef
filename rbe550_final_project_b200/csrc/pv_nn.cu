// pv_nn.cu -- nearest-tree-node candidates for a tree whose nodes are SHARDED over the GPUs of one box (SURVEY.md 8e,
// BASELINE north_star: "NCCL ... to gather verdicts and nearest-tree candidates for a batched multi-query RRT-Connect
// front end").  Each rank keeps every R-th node of every tree (global node g lives on rank g % world at slot g / world);
// this kernel answers, for a batch of (tree, target) pairs, "which of MY nodes is closest?" -- squared L2 distance in
// R^9 (og.RRTConnect's metric, planning.py:156), the node's GLOBAL index and its state.  The ranks' candidates are then
// all-gathered (8 + 36 B per pair per rank) and reduced by (distance, global index), which reproduces the nearest-
// neighbour decision of the single-GPU planner (pv_rrtc.cu: rrtc_nearest, ties to the lowest index) bit for bit: the
// distances are computed with the same fmaf chain.
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/panda_validity.h"
#include "pv_handle.h"

#define NN_THREADS 128

// trees: [n_trees][9][capacity] SoA (this rank's slots); sizes[t] = slots in use; targets [n_trees][9];
// out [n_trees][11] = squared distance, GLOBAL node index (int bits), node state -- the record that is all-gathered.
// Pair t searches tree tree_of[t] (null: tree t), so a batch can address any subset of the trees.
// Fused gather (peers / mc non-null): the record is ALSO stored into every rank's symmetric buffer, at rank-major
// position [rank][pair][11] -- one multimem.st per word through the NVSwitch multicast address, or one peer store per rank
// -- so the all-gather of the candidates happens inside the kernel that finds them (no collective launch; the ranks meet at
// the symmetric-memory barrier that follows).
__global__ void __launch_bounds__(NN_THREADS) pv_nn_kernel(const float* __restrict__ trees, const int* __restrict__ sizes,
                                                           const int* __restrict__ tree_of,
                                                           const float* __restrict__ targets, int capacity, int rank,
                                                           int world, float* __restrict__ out /* [n_pairs][11] or null */,
                                                           uint32_t* const* __restrict__ peers, uint32_t* __restrict__ mc,
                                                           int n_peers) {
    const int t = blockIdx.x;
    const int tree = tree_of ? tree_of[t] : t;
    const float* tq = trees + (size_t)tree * 9 * capacity;
    const int size = sizes[tree];
    float tg[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) tg[k] = targets[(size_t)t * 9 + k];
    float bd = 3.0e38f;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < size; i += NN_THREADS) {
        float d2 = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float d = tq[(size_t)k * capacity + i] - tg[k];
            d2 = fmaf(d, d, d2);
        }
        if (d2 < bd) {  // slots ascend with the global index, so the first minimum is the lowest index
            bd = d2;
            bi = i;
        }
    }
    __shared__ float s_d[NN_THREADS / 32];
    __shared__ int s_i[NN_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < bd || (od == bd && oi < bi)) {
            bd = od;
            bi = oi;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        s_d[threadIdx.x >> 5] = bd;
        s_i[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    __shared__ float s_rec[11];
    if (threadIdx.x == 0) {
        for (int w = 1; w < NN_THREADS / 32; ++w)
            if (s_d[w] < bd || (s_d[w] == bd && s_i[w] < bi)) {
                bd = s_d[w];
                bi = s_i[w];
            }
        s_rec[0] = bd;
        s_rec[1] = __int_as_float(size > 0 ? bi * world + rank : 0x7fffffff);
        s_i[0] = bi;
    }
    __syncthreads();
    const int slot = s_i[0];
    if (threadIdx.x < 9) s_rec[2 + threadIdx.x] = size > 0 ? tq[(size_t)threadIdx.x * capacity + slot] : 0.f;
    __syncthreads();
    if (threadIdx.x < 11) {
        const float v = s_rec[threadIdx.x];
        if (out) out[(size_t)t * 11 + threadIdx.x] = v;
        const size_t W = ((size_t)rank * gridDim.x + t) * 11 + threadIdx.x;
        if (mc) {
            asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(mc + W), "r"(__float_as_uint(v)) : "memory");
        } else if (peers) {
            for (int p = 0; p < n_peers; ++p) peers[p][W] = __float_as_uint(v);
        }
    }
}

static int nn_launch(PvHandle* h, const float* d_trees, const int* d_sizes, const int* d_tree_of, const float* d_targets,
                     int n_trees, int capacity, int rank, int world, float* d_out, const void* d_peer_ptrs, int n_peers,
                     void* d_multicast, void* stream, const char* who) {
    if (!h || h->magic != PV_HANDLE_MAGIC) return PV_ERR_BAD_HANDLE;
    if (n_trees < 0 || capacity < 1 || world < 1 || rank < 0 || rank >= world || n_peers < 0 || n_peers > 32 ||
        (n_trees > 0 && (!d_trees || !d_sizes || !d_targets || (!d_out && !d_peer_ptrs && !d_multicast)))) {
        snprintf(h->err, sizeof(h->err), "%s: bad arguments", who);
        return PV_ERR_BAD_ARG;
    }
    if (n_trees == 0) return PV_OK;
    PvDeviceGuard guard(h->device);
    pv_nn_kernel<<<n_trees, NN_THREADS, 0, (cudaStream_t)stream>>>(d_trees, d_sizes, d_tree_of, d_targets, capacity, rank, world,
                                                                   d_out, (uint32_t* const*)d_peer_ptrs, (uint32_t*)d_multicast,
                                                                   n_peers);
    h->launches++;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(h->err, sizeof(h->err), "pv_nn_kernel launch failed: %s", cudaGetErrorString(e));
        return PV_ERR_CUDA;
    }
    return PV_OK;
}

extern "C" int pv_nn_candidates(PvHandle* h, const float* d_trees, const int* d_sizes, const int* d_tree_of,
                                const float* d_targets, int n_trees, int capacity, int rank, int world, float* d_out,
                                void* stream) {
    if (n_trees > 0 && !d_out) return PV_ERR_BAD_ARG;
    return nn_launch(h, d_trees, d_sizes, d_tree_of, d_targets, n_trees, capacity, rank, world, d_out, nullptr, 0, nullptr,
                     stream, "pv_nn_candidates");
}

extern "C" int pv_nn_candidates_gather(PvHandle* h, const float* d_trees, const int* d_sizes, const int* d_tree_of,
                                       const float* d_targets, int n_pairs, int capacity, int rank, int world,
                                       const void* d_peer_ptrs, int n_peers, void* d_multicast, void* stream) {
    if (n_pairs > 0 && !d_peer_ptrs && !d_multicast) return PV_ERR_BAD_ARG;
    return nn_launch(h, d_trees, d_sizes, d_tree_of, d_targets, n_pairs, capacity, rank, world, nullptr, d_peer_ptrs, n_peers,
                     d_multicast, stream, "pv_nn_candidates_gather");
}

// ---- the two other device steps of the sharded-tree front end (distributed.ShardedTreePlanner) --------------------
// samples: exactly the stream of the single-GPU planner (rrtc_sample keyed by (seed, global search id, iteration))
__global__ void pv_rrtc_samples_kernel(unsigned seed, const unsigned* __restrict__ gsearch, const int* __restrict__ it, int n,
                                       float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float q[9], u9;
    rrtc_sample(seed, gsearch[i], (unsigned)it[i], q, u9);
#pragma unroll
    for (int k = 0; k < 9; ++k) out[(size_t)i * 9 + k] = q[k];
}

// steer: reduce the ranks' candidates by (distance, global index) and form the motion to validate -- from the nearest
// node towards the target, at most `range` long -- with the very expressions of pv_rrtc_kernel, so that the sharded front
// end takes the decisions of the single-GPU planner bit for bit.  cand: [world][n][11] = d2, gidx (int bits), state[9]
__global__ void pv_rrtc_steer_kernel(const float* __restrict__ cand, int world, int n, const float* __restrict__ targets,
                                     float range, int* __restrict__ from_gidx, float* __restrict__ ea_out,
                                     float* __restrict__ eb_out, int* __restrict__ reach_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float bd = 3.0e38f;
    int bi = 0x7fffffff, br = 0;
    for (int r = 0; r < world; ++r) {
        const float* c = cand + ((size_t)r * n + i) * 11;
        const float d2 = c[0];
        const int gi = __float_as_int(c[1]);
        if (d2 < bd || (d2 == bd && gi < bi)) {
            bd = d2;
            bi = gi;
            br = r;
        }
    }
    const float* c = cand + ((size_t)br * n + i) * 11;
    const float d = sqrtf(bd);
    float f = 1.0f;
    bool reach = true;
    if (d > range) {
        f = range / d;
        reach = false;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float ea = c[2 + k], g = targets[(size_t)i * 9 + k];
        ea_out[(size_t)i * 9 + k] = ea;
        eb_out[(size_t)i * 9 + k] = reach ? g : fmaf(f, g - ea, ea);
    }
    from_gidx[i] = bi;
    reach_out[i] = reach ? 1 : 0;
}

extern "C" int pv_rrtc_samples(PvHandle* h, uint32_t seed, const unsigned* d_gsearch, const int* d_it, int n, float* d_out,
                               void* stream) {
    if (!h || h->magic != PV_HANDLE_MAGIC) return PV_ERR_BAD_HANDLE;
    if (n < 0 || (n > 0 && (!d_gsearch || !d_it || !d_out))) return PV_ERR_BAD_ARG;
    if (n == 0) return PV_OK;
    PvDeviceGuard guard(h->device);
    pv_rrtc_samples_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, d_gsearch, d_it, n, d_out);
    h->launches++;
    return cudaGetLastError() == cudaSuccess ? PV_OK : PV_ERR_CUDA;
}

extern "C" int pv_rrtc_steer(PvHandle* h, const float* d_cand, int world, int n, const float* d_targets, float range,
                             int* d_from_gidx, float* d_ea, float* d_eb, int* d_reach, void* stream) {
    if (!h || h->magic != PV_HANDLE_MAGIC) return PV_ERR_BAD_HANDLE;
    if (n < 0 || world < 1 || (n > 0 && (!d_cand || !d_targets || !d_from_gidx || !d_ea || !d_eb || !d_reach))) return PV_ERR_BAD_ARG;
    if (n == 0) return PV_OK;
    PvDeviceGuard guard(h->device);
    pv_rrtc_steer_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_cand, world, n, d_targets,
                                                                           range > 0.f ? range : PV_RRTC_RANGE, d_from_gidx,
                                                                           d_ea, d_eb, d_reach);
    h->launches++;
    return cudaGetLastError() == cudaSuccess ? PV_OK : PV_ERR_CUDA;
}
