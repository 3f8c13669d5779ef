// pv_ik.cu -- batched, collision-aware inverse kinematics for the Panda hand (next-row component).
//
// Replaces robot.inverse_kinematics(link=hand, pos, quat) as the motion primitives call it right before every
// plan_path (motion_primitives.py:131-134, 273, 278, 388, ...).  One block per target pose, one thread per seed:
// every thread runs damped least squares on the 7 arm joints (geometric Jacobian from the same FK the validity
// kernels use, 6x6 normal equations solved by an unrolled Cholesky in registers), converged candidates are pushed
// through the state-validity rule (pv_check_config), and the block returns the VALID solution closest to the
// initial configuration -- the assignment's "validate the IK solution for collisions" (Project5.pdf p.2) for free.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/panda_validity.h"
#include "pv_device.cuh"
#include "pv_handle.h"

#define IK_MAX_SEEDS 256

struct IkArgs {
    const float* pos;     // [n][3]
    const float* quat;    // [n][4] wxyz
    const float* q_init;  // [9]
    int n_targets, n_seeds, max_iters;
    float pos_tol, rot_tol, damping;
    unsigned seed;
    float* q_out;   // [n][9]
    int* status;    // [n]
    float* err;     // [n][2]
};

__device__ __forceinline__ float3 v_cross(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

template <bool CARRY>
__global__ void __launch_bounds__(IK_MAX_SEEDS, 1) pv_ik_kernel(const __grid_constant__ PvScene S,
                                                                const __grid_constant__ IkArgs A) {
    const int target = blockIdx.x;
    const int tid = threadIdx.x;
    const float lo[9] = PV_Q_LOWER, hi[9] = PV_Q_UPPER;

    // target pose
    const float3 pt = make_float3(A.pos[3 * target], A.pos[3 * target + 1], A.pos[3 * target + 2]);
    float qw = A.quat[4 * target], qx = A.quat[4 * target + 1], qy = A.quat[4 * target + 2], qz = A.quat[4 * target + 3];
    {
        const float n = rsqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
        qw *= n; qx *= n; qy *= n; qz *= n;
    }
    const float3 tX = make_float3(1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy + qz * qw), 2 * (qx * qz - qy * qw));
    const float3 tY = make_float3(2 * (qx * qy - qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz + qx * qw));
    const float3 tZ = make_float3(2 * (qx * qz + qy * qw), 2 * (qy * qz - qx * qw), 1 - 2 * (qx * qx + qy * qy));

    // seed: thread 0 starts from q_init, the others from uniform samples in the joint limits
    float q[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) q[j] = A.q_init[j];
    if (tid > 0) {
        float r[9];
        pv_sweep_config((uint64_t)target * IK_MAX_SEEDS + tid, A.seed ^ 0x494B0000u, false, r);
#pragma unroll
        for (int j = 0; j < 7; ++j) q[j] = r[j];
    }

    bool converged = false;
    float ep_n = 1e30f, er_n = 1e30f;
    for (int it = 0; it <= A.max_iters; ++it) {
        float3 jp[8], jz[8], hp, hX, hY, hZ;
        pv_fk_visit(q, S.base[0], S.base[1], S.base[2], [&](auto lc, float3 p, float3 X, float3 Y, float3 Z) {
            constexpr int l = decltype(lc)::value;
            if constexpr (l >= 1 && l <= 7) {
                jp[l] = p;
                jz[l] = Z;
            }
            if constexpr (l == 8) {
                hp = p; hX = X; hY = Y; hZ = Z;
            }
        });
        const float3 ep = v_sub(pt, hp);
        float3 er = v_cross(hX, tX);
        {
            const float3 b = v_cross(hY, tY), c = v_cross(hZ, tZ);
            er = make_float3(0.5f * (er.x + b.x + c.x), 0.5f * (er.y + b.y + c.y), 0.5f * (er.z + b.z + c.z));
        }
        // a half-turn error has a vanishing cross-product sum: use the trace to tell it from "aligned"
        const float tr = v_dot(hX, tX) + v_dot(hY, tY) + v_dot(hZ, tZ);
        ep_n = sqrtf(v_dot(ep, ep));
        er_n = sqrtf(v_dot(er, er));
        if (tr < 0.f) er_n = fmaxf(er_n, 1.0f);
        converged = (ep_n < A.pos_tol) && (er_n < A.rot_tol) && (tr > 0.f);
        if (converged || it == A.max_iters) break;

        // geometric Jacobian, columns i = 1..7: [z_i x (p_h - p_i); z_i]
        float J[6][7];
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const float3 c = v_cross(jz[i + 1], v_sub(hp, jp[i + 1]));
            J[0][i] = c.x; J[1][i] = c.y; J[2][i] = c.z;
            J[3][i] = jz[i + 1].x; J[4][i] = jz[i + 1].y; J[5][i] = jz[i + 1].z;
        }
        float e[6] = {ep.x, ep.y, ep.z, er.x, er.y, er.z};
        // adaptive damping: larger while far away, small near the solution
        const float lam2 = A.damping * A.damping + 0.05f * fminf(1.0f, ep_n * ep_n + er_n * er_n);
        float M[6][6];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) {
                float s = (r == c) ? lam2 : 0.f;
#pragma unroll
                for (int i = 0; i < 7; ++i) s = fmaf(J[r][i], J[c][i], s);
                M[r][c] = s;
            }
        // Cholesky M = L L^T (in place, lower), then solve L L^T y = e
#pragma unroll
        for (int c = 0; c < 6; ++c) {
#pragma unroll
            for (int k = 0; k < c; ++k) M[c][c] = fmaf(-M[c][k], M[c][k], M[c][c]);
            const float inv = rsqrtf(fmaxf(M[c][c], 1e-12f));
            M[c][c] = inv;  // store 1 / L_cc
#pragma unroll
            for (int r = c + 1; r < 6; ++r) {
#pragma unroll
                for (int k = 0; k < c; ++k) M[r][c] = fmaf(-M[r][k], M[c][k], M[r][c]);
                M[r][c] *= inv;
            }
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int k = 0; k < r; ++k) e[r] = fmaf(-M[r][k], e[k], e[r]);
            e[r] *= M[r][r];
        }
#pragma unroll
        for (int r = 5; r >= 0; --r) {
#pragma unroll
            for (int k = r + 1; k < 6; ++k) e[r] = fmaf(-M[k][r], e[k], e[r]);
            e[r] *= M[r][r];
        }
        float dq[7], big = 0.f;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < 6; ++r) s = fmaf(J[r][i], e[r], s);
            dq[i] = s;
            big = fmaxf(big, fabsf(s));
        }
        const float scale = big > 0.5f ? 0.5f / big : 1.0f;  // trust region: at most 0.5 rad per joint per step
#pragma unroll
        for (int i = 0; i < 7; ++i) q[i] = fminf(fmaxf(fmaf(scale, dq[i], q[i]), lo[i]), hi[i]);
    }

    // validity of the candidate (all lanes call together; non-converged lanes are masked afterwards)
    PvAcc<PV_MODE_BITS> acc;
    pv_check_config<PV_MODE_BITS, true, PV_EXIT_NONE, 0, false, CARRY>(q, S, acc);
    const bool good = converged && !acc.hit;

    // the valid candidate closest to q_init (L2 over the arm joints); ties -> lowest thread index
    float cost = 3.0e38f;
    if (good) {
        cost = 0.f;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const float d = q[i] - A.q_init[i];
            cost = fmaf(d, d, cost);
        }
    }
    __shared__ float s_cost[IK_MAX_SEEDS];
    __shared__ int s_idx[IK_MAX_SEEDS];
    s_cost[tid] = cost;
    s_idx[tid] = tid;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (tid < o) {
            const float oc = s_cost[tid + o];
            const int oi = s_idx[tid + o];
            if (oc < s_cost[tid] || (oc == s_cost[tid] && oi < s_idx[tid])) {
                s_cost[tid] = oc;
                s_idx[tid] = oi;
            }
        }
        __syncthreads();
    }
    const bool any_good = s_cost[0] < 3.0e38f;
    if (tid == s_idx[0]) {
        if (any_good) {
#pragma unroll
            for (int j = 0; j < 9; ++j) A.q_out[9 * target + j] = q[j];
            A.err[2 * target] = ep_n;
            A.err[2 * target + 1] = er_n;
        }
        A.status[target] = any_good ? 1 : 0;
    }
}

extern "C" int pv_ik_batch(PvHandle* h, const float* h_pos, const float* h_quat, int n_targets, const float* h_q_init,
                           int n_seeds, int max_iters, float pos_tol, float rot_tol, uint32_t seed, float* h_q_out,
                           int* h_status, float* h_err) {
    if (!h || h->magic != PV_HANDLE_MAGIC) return PV_ERR_BAD_HANDLE;
    if (!h->has_scene) {
        snprintf(h->err, sizeof(h->err), "no scene set (pv_set_scene)");
        return PV_ERR_NO_SCENE;
    }
    if (n_targets < 0 || !h_q_init || (n_targets > 0 && (!h_pos || !h_quat || !h_q_out || !h_status))) {
        snprintf(h->err, sizeof(h->err), "pv_ik_batch: bad arguments");
        return PV_ERR_BAD_ARG;
    }
    if (n_targets == 0) return PV_OK;
    if (n_seeds < 32) n_seeds = 32;
    if (n_seeds > IK_MAX_SEEDS) n_seeds = IK_MAX_SEEDS;
    int p2 = 32;
    while (p2 < n_seeds) p2 <<= 1;  // the block reduction wants a power of two
    n_seeds = p2;
#define IK_CUDA(expr)                                                                                        \
    do {                                                                                                     \
        cudaError_t e_ = (expr);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            snprintf(h->err, sizeof(h->err), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                     __LINE__);                                                                              \
            return PV_ERR_CUDA;                                                                              \
        }                                                                                                    \
    } while (0)
    PvDeviceGuard guard(h->device);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_pos = al((size_t)n_targets * 3 * 4), b_quat = al((size_t)n_targets * 4 * 4), b_qi = al(9 * 4);
    const size_t b_q = al((size_t)n_targets * 9 * 4), b_st = al((size_t)n_targets * 4), b_err = al((size_t)n_targets * 8);
    const size_t total = b_pos + b_quat + b_qi + b_q + b_st + b_err;
    if (total > h->ik_bytes) {
        if (h->ik_buf) cudaFree(h->ik_buf);
        h->ik_buf = nullptr;
        h->ik_bytes = 0;
        IK_CUDA(cudaMalloc(&h->ik_buf, total));
        h->ik_bytes = total;
    }
    char* p = (char*)h->ik_buf;
    IkArgs a;
    memset(&a, 0, sizeof(a));
    float* d_pos = (float*)p; p += b_pos;
    float* d_quat = (float*)p; p += b_quat;
    float* d_qi = (float*)p; p += b_qi;
    a.q_out = (float*)p; p += b_q;
    a.status = (int*)p; p += b_st;
    a.err = (float*)p; p += b_err;
    a.pos = d_pos; a.quat = d_quat; a.q_init = d_qi;
    a.n_targets = n_targets; a.n_seeds = n_seeds;
    a.max_iters = max_iters > 0 ? max_iters : 64;
    a.pos_tol = pos_tol > 0.f ? pos_tol : 1e-4f;
    a.rot_tol = rot_tol > 0.f ? rot_tol : 1e-3f;
    a.damping = 0.01f;
    a.seed = seed;
    cudaStream_t st = h->streams[0];
    IK_CUDA(cudaMemcpyAsync(d_pos, h_pos, (size_t)n_targets * 12, cudaMemcpyHostToDevice, st));
    IK_CUDA(cudaMemcpyAsync(d_quat, h_quat, (size_t)n_targets * 16, cudaMemcpyHostToDevice, st));
    IK_CUDA(cudaMemcpyAsync(d_qi, h_q_init, 36, cudaMemcpyHostToDevice, st));
    IK_CUDA(cudaMemsetAsync(a.q_out, 0, b_q + b_st + b_err, st));
    if (h->scene.carry) pv_ik_kernel<true><<<n_targets, n_seeds, 0, st>>>(h->scene, a);
    else pv_ik_kernel<false><<<n_targets, n_seeds, 0, st>>>(h->scene, a);
    h->launches++;
    IK_CUDA(cudaGetLastError());
    IK_CUDA(cudaMemcpyAsync(h_q_out, a.q_out, (size_t)n_targets * 36, cudaMemcpyDeviceToHost, st));
    IK_CUDA(cudaMemcpyAsync(h_status, a.status, (size_t)n_targets * 4, cudaMemcpyDeviceToHost, st));
    if (h_err) IK_CUDA(cudaMemcpyAsync(h_err, a.err, (size_t)n_targets * 8, cudaMemcpyDeviceToHost, st));
    IK_CUDA(cudaStreamSynchronize(st));
    return PV_OK;
#undef IK_CUDA
}
