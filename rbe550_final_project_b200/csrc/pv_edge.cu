// pv_edge.cu -- K3 motion-validity kernels (fused interpolation + state check + any-hit early exit) and
// their C-ABI entry points.  Split from pv_kernels.cu so the two translation units compile in parallel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/panda_validity.h"
#include "pv_device.cuh"
#include "pv_handle.h"

// verdict bits of edges from the hardware sin/cos, like the state kernels (pv_device.cuh, PV_FAST_TRIG); margins keep
// the accurate form
#ifndef PV_EDGE_FAST_TRIG
#define PV_EDGE_FAST_TRIG 1
#endif


#define PV_CUDA(h, expr)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            snprintf((h)->err, sizeof((h)->err), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                                             \
            return PV_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

#define PV_PRECHECK(h, n)                                                   \
    if (!(h) || (h)->magic != PV_HANDLE_MAGIC) return PV_ERR_BAD_HANDLE;    \
    if (!(h)->has_scene) {                                                  \
        snprintf((h)->err, sizeof((h)->err), "no scene set (pv_set_scene)"); \
        return PV_ERR_NO_SCENE;                                             \
    }                                                                       \
    if ((n) < 0) {                                                          \
        snprintf((h)->err, sizeof((h)->err), "negative count");             \
        return PV_ERR_BAD_ARG;                                              \
    }                                                                       \
    if ((n) == 0) return PV_OK;                                             \
    PvDeviceGuard pv_guard_((h)->device);

// K3: one warp per edge, lanes = interpolation states, coarse-to-fine rounds, any-hit early exit.
// Each warp owns `epw` consecutive edges at a time: 32 for large batches (it then emits one whole verdict word, no
// atomics), 1 for small ones (a plan's 150 waypoint motions, a simplifier batch), where 32 edges per warp would leave
// the GPU to a handful of warps -- the verdict then goes to ok_bytes[e] (pv_plan.cu) or is OR-ed into a zeroed bit word.
// PV_E_LOCKSTEP = 1 makes the warps of a block advance ROUND by round behind a block barrier (one pv_check_config per
// warp per round, the loop ends when __syncthreads_or says no warp has work left) so that they share instruction
// fetches like the state kernel.  That was neutral while a round cost ~3 700 instructions; since the scene-level cull
// (a round: ~1 400) the barrier is the top stall (2.7 warp-cycles per issue, issue-active 39 %) and free-running warps
// are 20-25 % faster (config 3: 305 -> 367 M edges/s), so it is off.
#ifndef PV_E_THREADS
#define PV_E_THREADS 384
#endif
#ifndef PV_E_LOCKSTEP
#define PV_E_LOCKSTEP 0
#endif
// Verdict words of large batches also go to the handle's fused-gather target (pv_set_gather), like the state kernels' --
// from a SEPARATE instantiation that is only launched while a gather is configured: with the epilogue compiled into the
// one kernel the single-GPU rate fell by 3-4.5 % (402 -> 384 M edges/s on the pentagon scene; the layout of this
// fetch-bound kernel is that sensitive), so the default instantiation carries neither the parameter nor the code.
template <bool ON>
struct PvGatherOpt {};
template <>
struct PvGatherOpt<true> {
    PvGather g;
};

template <bool CULL, int MODE, bool CARRY, bool GATHER = false, bool YAW = false>
__global__ void __launch_bounds__(PV_E_THREADS, 1)
    pv_edge_kernel(const __grid_constant__ PvScene S, const float4* __restrict__ aA, const float4* __restrict__ aB,
                   const float* __restrict__ a9, const float4* __restrict__ bA, const float4* __restrict__ bB,
                   const float* __restrict__ b9, const float* __restrict__ a_aos, const float* __restrict__ b_aos,
                   int64_t n_edges, int n_steps, float resolution, uint32_t* __restrict__ bits,
                   float* __restrict__ margin, int epw, unsigned char* __restrict__ ok_bytes,
                   const __grid_constant__ PvGatherOpt<GATHER> GO) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t n_words = (n_edges + epw - 1) / epw;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;

    // warp state machine: word w, edge j of the word, round r of the edge.  Only the end points of the edge being
    // validated are kept in registers (warp-uniform, fetched with broadcast loads that hit L1), which took the kernel
    // from 95 to 132 M edges/s at 384 threads (512-thread blocks spill and are slower; block sizes swept 384..512).
    float ea[9], eb[9];
    int j = 0, r = 0, n_here = 0, nd = 1, rounds = 1;
    float inv_nd = 1.f, edge_m = 1e30f;
    unsigned word = 0;
    bool need_word = true, need_edge = true;

    for (;;) {
        const bool have = w < n_words;
#if PV_E_LOCKSTEP == 1
        if (!__syncthreads_or(have ? 1 : 0)) break;
        if (!have) continue;
#elif PV_E_LOCKSTEP == 2
        // lockstep only among the warps that share a scheduler -- and with it an L0 instruction cache: warp ids w, w + 4,
        // w + 8 meet at a named barrier once per round and walk the round's code together
        {
            unsigned any_;
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(any_)
                         : "r"(have ? 1u : 0u), "r"(1 + ((threadIdx.x >> 5) & 3)), "n"(PV_E_THREADS / 4)
                         : "memory");
            if (!any_) break;
            if (!have) continue;
        }
#else
        if (!have) break;
#endif
        if (need_word) {
            n_here = (int)min((int64_t)epw, n_edges - w * epw);
            word = 0;
            j = 0;
            need_word = false;
            need_edge = true;
        }
        if (need_edge) {
            const int64_t e = w * epw + j;
            if (a_aos) {
                pv_load_aos(a_aos, e, ea);
                pv_load_aos(b_aos, e, eb);
            } else {
                pv_load_soa(aA, aB, a9, e, ea);
                pv_load_soa(bA, bB, b9, e, eb);
            }
            float d2 = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float de = eb[k] - ea[k];
                d2 = fmaf(de, de, d2);
            }
            nd = n_steps;
            if (nd <= 0) nd = max(1, (int)ceilf(sqrtf(d2) / resolution));
            rounds = (nd + 31) >> 5;
            inv_nd = 1.0f / (float)nd;
            r = 0;
            edge_m = 1e30f;
            need_edge = false;
        }
        // one round: lane -> state k (coarse to fine; idle lanes re-check the end point)
        int k = nd - (lane * rounds + r);
        if (k < 1) k = nd;
        const float t = (float)k * inv_nd;
        float q[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) q[c] = (k == nd) ? eb[c] : fmaf(t, eb[c] - ea[c], ea[c]);
        PvAcc<MODE> acc;
        constexpr bool COLD = PV_COLD_SCENE_WARP && CULL && MODE == PV_MODE_BITS;
        constexpr int EX = (MODE == PV_MODE_BITS ? PV_EXIT_ANY : PV_EXIT_NONE);
        constexpr bool FT = (PV_EDGE_FAST_TRIG && MODE == PV_MODE_BITS);
        if (pv_check_config<MODE, CULL, EX, 0, false, CARRY, FT, COLD, YAW>(q, S, acc)) {
            // warp-uniform (the check votes before it reports): the whole warp runs the scene section out of line
            if constexpr (COLD) {
                PvReloadLerp rl;
#pragma unroll
                for (int c = 0; c < 9; ++c) {
                    rl.ea[c] = ea[c];
                    rl.eb[c] = eb[c];
                }
                rl.t = t;
                rl.at_end = (k == nd);
                acc.hit |= pv_scene_cold<CULL, EX, false, CARRY, FT>(rl, S);
            }
        }
        bool edge_done;
        bool edge_hit = false;
        if constexpr (MODE == PV_MODE_BITS) {
            edge_hit = __any_sync(FULL, acc.hit);
            edge_done = edge_hit || (r + 1 >= rounds);
        } else {
            edge_m = fminf(edge_m, acc.m);
            edge_done = (r + 1 >= rounds);
        }
        if (!edge_done) {
            ++r;
            continue;
        }
        if constexpr (MODE == PV_MODE_BITS) {
            word |= (edge_hit ? 0u : 1u) << j;
        } else {
            float m = edge_m;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(FULL, m, o));
            if (lane == 0) margin[w * epw + j] = m;
        }
        ++j;
        need_edge = true;
        if (j >= n_here) {
            if constexpr (MODE == PV_MODE_BITS) {
                if constexpr (GATHER) {
                    if (epw == 32) pv_emit_word(bits, GO.g, w, word, lane);  // local word + peer / multicast stores
                }
                if (lane == 0) {
                    if (epw == 32) {
                        if constexpr (!GATHER) bits[w] = word;
                    } else if (ok_bytes) ok_bytes[w] = (unsigned char)(word & 1u);  // epw == 1
                    else if (word & 1u) atomicOr(bits + (w >> 5), 1u << (w & 31));  // epw == 1, bits zeroed by the launcher
                }
            }
            w += n_warps;
            need_word = true;
        }
    }
}

// batches up to this size are validated one edge per warp (see the kernel's header)
#ifndef PV_E_SMALL_BATCH
#define PV_E_SMALL_BATCH 16384
#endif

int pv_launch_edges(PvHandle* h, const float* aA, const float* aB, const float* a9, const float* bA,
                           const float* bB, const float* b9, const float* a_aos, const float* b_aos, int64_t n,
                           int n_steps, float resolution, uint32_t* d_bits, float* d_margin, cudaStream_t st,
                           unsigned char* d_ok_bytes, bool allow_gather) {
    if (n_steps < 0 || (n_steps == 0 && !(resolution > 0.f))) {
        snprintf(h->err, sizeof(h->err), "edge check needs n_steps > 0 or resolution > 0");
        return PV_ERR_BAD_ARG;
    }
    const int epw = (d_ok_bytes || (d_bits && n <= PV_E_SMALL_BATCH)) ? 1 : 32;
    const int64_t words = (n + epw - 1) / epw;
    if (epw == 1 && d_bits && !d_ok_bytes) PV_CUDA(h, cudaMemsetAsync(d_bits, 0, (size_t)((n + 31) / 32) * sizeof(uint32_t), st));
    // the fused-gather target applies to whole verdict words only (large batches); small batches and margins never gather
    // (and device-buffer calls only: a host-buffer call numbers its words per chunk, see pv_launch_state_bits)
    const bool gather_on = allow_gather && epw == 32 && d_bits && h->gather.n_peers > 0;
#define PV_LAUNCH_E(CULL, MODE, CARRY, GATHER, GARG, YAW_)                                                        \
    {                                                                                                             \
        int grid = pv_grid_for(h, (const void*)pv_edge_kernel<CULL, MODE, CARRY, GATHER, YAW_>, PV_E_THREADS, words); \
        pv_edge_kernel<CULL, MODE, CARRY, GATHER, YAW_><<<grid, PV_E_THREADS, 0, st>>>(                           \
            h->scene, (const float4*)aA, (const float4*)aB, a9, (const float4*)bA, (const float4*)bB, b9, a_aos,  \
            b_aos, n, n_steps, resolution, d_bits, d_margin, epw, d_ok_bytes, GARG);                              \
    }
    if (gather_on) {
        const PvGatherOpt<true> go = {h->gather};
        if (h->scene.carry) PV_LAUNCH_E(true, PV_MODE_BITS, true, true, go, false)
        else if (h->all_yaw) PV_LAUNCH_E(true, PV_MODE_BITS, false, true, go, true)
        else PV_LAUNCH_E(true, PV_MODE_BITS, false, true, go, false)
    } else if (d_bits || d_ok_bytes) {
        if (h->scene.carry) PV_LAUNCH_E(true, PV_MODE_BITS, true, false, PvGatherOpt<false>{}, false)
        else if (h->all_yaw) PV_LAUNCH_E(true, PV_MODE_BITS, false, false, PvGatherOpt<false>{}, true)
        else PV_LAUNCH_E(true, PV_MODE_BITS, false, false, PvGatherOpt<false>{}, false)
    } else {  // margins: always brute force
        if (h->scene.carry) PV_LAUNCH_E(false, PV_MODE_MARGIN, true, false, PvGatherOpt<false>{}, false)
        else PV_LAUNCH_E(false, PV_MODE_MARGIN, false, false, PvGatherOpt<false>{}, false)
    }
#undef PV_LAUNCH_E
    h->launches++;
    PV_CUDA(h, cudaGetLastError());
    return PV_OK;
}

extern "C" int pv_check_edges(PvHandle* h, const float* d_aA, const float* d_aB, const float* d_a9, const float* d_bA,
                   const float* d_bB, const float* d_b9, int64_t n_edges, int n_steps, float resolution,
                   uint32_t* d_bits, void* stream) {
    PV_PRECHECK(h, n_edges);
    if (!d_aA || !d_aB || !d_bA || !d_bB || !d_bits) return PV_ERR_BAD_ARG;
    return pv_launch_edges(h, d_aA, d_aB, d_a9, d_bA, d_bB, d_b9, nullptr, nullptr, n_edges, n_steps, resolution,
                           d_bits, nullptr, (cudaStream_t)stream, nullptr, true);
}

extern "C" int pv_edge_margins(PvHandle* h, const float* d_aA, const float* d_aB, const float* d_a9, const float* d_bA,
                    const float* d_bB, const float* d_b9, int64_t n_edges, int n_steps, float resolution,
                    float* d_margin, void* stream) {
    PV_PRECHECK(h, n_edges);
    if (!d_aA || !d_aB || !d_bA || !d_bB || !d_margin) return PV_ERR_BAD_ARG;
    return pv_launch_edges(h, d_aA, d_aB, d_a9, d_bA, d_bB, d_b9, nullptr, nullptr, n_edges, n_steps, resolution,
                           nullptr, d_margin, (cudaStream_t)stream, nullptr);
}

