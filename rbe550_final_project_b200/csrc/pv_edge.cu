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

// LIST: the motions to validate are the entries of `list` (their count is read from *n_list: the certificate pass
// below wrote both), and a valid motion sets its own bit in the zero-initialised word with an atomic.
// SECT (LIST only): the sections of the check this instantiation contains (pv_check_config; 3 = all).
// APPEND (LIST only): a motion found valid is not final yet -- it is appended to out_list (length *out_n) for the kernel
// that holds the other section, instead of setting its bit.
template <bool CULL, int MODE, bool CARRY, bool GATHER = false, bool YAW = false, bool LIST = false, int SECT = 3,
          bool APPEND = false>
__global__ void __launch_bounds__(PV_E_THREADS, 1)
    pv_edge_kernel(const __grid_constant__ PvScene S, const float4* __restrict__ aA, const float4* __restrict__ aB,
                   const float* __restrict__ a9, const float4* __restrict__ bA, const float4* __restrict__ bB,
                   const float* __restrict__ b9, const float* __restrict__ a_aos, const float* __restrict__ b_aos,
                   int64_t n_edges, int n_steps, float resolution, uint32_t* __restrict__ bits,
                   float* __restrict__ margin, int epw, unsigned char* __restrict__ ok_bytes,
                   const __grid_constant__ PvGatherOpt<GATHER> GO, const unsigned* __restrict__ list = nullptr,
                   const unsigned* __restrict__ n_list = nullptr, unsigned* __restrict__ next_group = nullptr,
                   unsigned* __restrict__ out_list = nullptr, unsigned* __restrict__ out_n = nullptr) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    if constexpr (LIST) n_edges = (int64_t)__ldg(n_list);
    unsigned my_e = 0;  // LIST: lane i holds the i-th motion of the group of 32 this warp works on
    const int64_t n_words = (n_edges + epw - 1) / epw;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // LIST: the listed motions are the expensive ones and their cost varies, so the warps draw small groups (epw = 4)
    // from a counter instead of striding over the list: no warp is left with a long tail
    auto next_w = [&]() -> int64_t {
        unsigned g_ = 0;
        if (lane == 0) g_ = atomicAdd(next_group, 1u);
        return (int64_t)__shfl_sync(0xffffffffu, g_, 0);
    };
    // ... and stay ONE group ahead: the counter draw and the group's list entries (two dependent global accesses) are
    // issued while the previous group is being validated, and the end points of the next motion are prefetched into
    // L2 / L1 when the current one starts (with 12 warps per SM nothing else hides those latencies)
    int64_t w_nx = 0;
    unsigned e_nx = 0;
    auto load_group = [&](int64_t g_) -> unsigned {
        const int64_t i_ = g_ * epw + lane;
        return (lane < epw && i_ < n_edges) ? __ldg(list + i_) : 0u;
    };
    if constexpr (LIST) {
        w = next_w();
        my_e = load_group(w);
        w_nx = next_w();
        e_nx = load_group(w_nx);
    }

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
            int64_t e = w * epw + j;
            if constexpr (LIST) {
                e = (int64_t)__shfl_sync(FULL, my_e, j);
                // the motion after this one (next entry of the group, else the first one of the next group)
                const unsigned en_ = (j + 1 < n_here) ? __shfl_sync(FULL, my_e, (j + 1) & 31) : __shfl_sync(FULL, e_nx, 0);
                if (lane < 4) {
                    const float4* p_ = (lane == 0 ? aA : lane == 1 ? aB : lane == 2 ? bA : bB) + en_;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p_));
                }
            }
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
        if (pv_check_config<MODE, CULL, EX, 0, false, CARRY, FT, COLD, YAW, SECT>(q, S, acc)) {
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
                if constexpr (LIST) {  // epw == 32: lane i owns the i-th entry of this group
                    const bool ok_ = lane < n_here && ((word >> lane) & 1u);
                    if constexpr (APPEND) {
                        const unsigned om_ = __ballot_sync(FULL, ok_);
                        if (om_) {
                            unsigned base_ = 0;
                            if (lane == 0) base_ = atomicAdd(out_n, (unsigned)__popc(om_));
                            base_ = __shfl_sync(FULL, base_, 0);
                            if (ok_) out_list[base_ + __popc(om_ & ((1u << lane) - 1u))] = my_e;
                        }
                    } else if (ok_) {
                        atomicOr(bits + (my_e >> 5), 1u << (my_e & 31u));
                    }
                } else if (lane == 0) {
                    if (epw == 32) {
                        if constexpr (!GATHER) bits[w] = word;
                    } else if (ok_bytes) ok_bytes[w] = (unsigned char)(word & 1u);  // epw == 1
                    else if (word & 1u) atomicOr(bits + (w >> 5), 1u << (w & 31));  // epw == 1, bits zeroed by the launcher
                }
            }
            if constexpr (LIST) {
                w = w_nx;
                my_e = e_nx;
                w_nx = next_w();
                e_nx = load_group(w_nx);
            } else {
                w += n_warps;
            }
            need_word = true;
        }
    }
}

// ---- motion certificates (pv_set_culling(2), the default; large batches of motions without a carried box) ----------
// A motion of nd > 32 interpolation states is first looked at COARSELY by this small kernel: half a warp per motion, 16
// states t_i = nd - floor((s - 1) / 2) - i s with s = ceil(nd / 16), so that every state 1..nd lies within h = floor(s / 2)
// interpolation steps of a tested one.  Between two configurations dq apart no robot point moves farther than
// sum_j |dq_j| R_j + |dq_8| + |dq_9| (panda_model.motion_reach_bounds; tests/test_model_pruning.py), so with
// dl = h steps of that travel: if all 16 tested states are inside the joint limits and clear every culling test and the
// ground plane by more than dl (pv_cull_clear), every state of the motion passes all of its culls -- none of them is in
// contact, the start state's joints being inside the limits too -- and the motion is VALID without a single primitive
// test.  52 % of the config-3 motions (gaussian steps, sigma 0.3 rad, 64 states) finish here at a quarter of the states
// and a fraction of the instructions; the others go, as a compacted list, through pv_edge_kernel<LIST> unchanged.
// Exact, not a heuristic: the verdict words are those of the exhaustive validator (pv_set_culling(1); tests/
// test_gpu_parity.py::test_motion_certificates_do_not_change_verdicts).  The kernel is 10 KB of code at <= 80 registers:
// none of the instruction-fetch trouble of the full check (profiles/r2_notes.md).
#ifndef PV_EDGE_CERT
#define PV_EDGE_CERT 1
#endif
#ifndef PV_CERT_MIN_ND
#define PV_CERT_MIN_ND 33
#endif
#ifndef PV_CERT_THREADS
#define PV_CERT_THREADS 256
#endif
__global__ void __launch_bounds__(PV_CERT_THREADS, 3)
    pv_edge_cert_kernel(const __grid_constant__ PvScene S, const float4* __restrict__ aA, const float4* __restrict__ aB,
                        const float* __restrict__ a9, const float4* __restrict__ bA, const float4* __restrict__ bB,
                        const float* __restrict__ b9, int64_t n_edges, int n_steps, float resolution,
                        uint32_t* __restrict__ bits, unsigned* __restrict__ lists, unsigned* __restrict__ n_list) {
    // lists: two arrays of n_edges entries each -- the motions that provably never come near the scene boxes and need
    // the self-collision section only (class 1), and the others (class 2: the whole check); n_list[c - 1] = their lengths
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, sub = lane & 15, half = lane >> 4;
    const int64_t n_words = (n_edges + 31) >> 5;  // (n_edges < 2^32: the launcher checks)
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words; w += n_warps) {
        unsigned word = 0;   // bit j: motion 32 w + j is certified valid
        // lane-parallel setup: lane L looks at motion 32 w + L -- its number of states, whether a certificate is sought
        // (more than 32 states, both ends inside the joint limits, slack within the cap) and with which slack
        const int64_t mine = w * 32 + lane;
        float dl_m = 0.f;
        int nd_m = 1;
        if (mine < n_edges) {
            float ea[9], eb[9];
            pv_load_soa(aA, aB, a9, mine, ea);
            pv_load_soa(bA, bB, b9, mine, eb);
            const float reach[7] = PV_MOTION_REACH, lo[9] = PV_Q_LOWER, hi[9] = PV_Q_UPPER;
            float d2 = 0.f, trav = fabsf(eb[7] - ea[7]) + fabsf(eb[8] - ea[8]);
            bool in_ = true;  // both ends inside the joint limits: so is every state between them
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                const float de = eb[c] - ea[c];
                d2 = fmaf(de, de, d2);
                if (c < 7) trav = fmaf(fabsf(de), reach[c], trav);
                in_ = in_ && ea[c] >= lo[c] && ea[c] <= hi[c] && eb[c] >= lo[c] && eb[c] <= hi[c];
            }
            int nd = n_steps;
            if (nd <= 0) nd = max(1, (int)ceilf(sqrtf(d2) / resolution));  // as pv_edge_kernel counts them
            const int h = ((nd + 15) >> 4) >> 1;
            const float dl = fmaf(trav / (float)nd, (float)h * 1.0002f, 2e-5f);
            nd_m = nd;
            if (nd >= PV_CERT_MIN_ND && in_ && dl <= PV_MOTION_CERT_MAX_SLACK) dl_m = dl;
        }
        const unsigned seek = __ballot_sync(FULL, dl_m > 0.f);
        // (cls0, cls1): bit j = class of motion 32 w + j, low and high bit; no certificate sought = the whole check
        unsigned cls0 = 0, cls1 = ~seek;
        for (int it = 0; it < 16; ++it) {
            if (!((seek >> (2 * it)) & 3u)) continue;  // warp-uniform: neither motion of this pair seeks one
            const int src = 2 * it + half;
            const float dl = __shfl_sync(FULL, dl_m, src);
            const int nd = __shfl_sync(FULL, nd_m, src);
            unsigned st = 0;
            if (dl > 0.f) {
                const int64_t e = w * 32 + src;
                float ea[9], eb[9];
                pv_load_soa(aA, aB, a9, e, ea);  // (L1 hits: the owner lane has just read them)
                pv_load_soa(bA, bB, b9, e, eb);
                const int s = (nd + 15) >> 4;
                int k = nd - ((s - 1) >> 1) - sub * s;
                if (k < 1) k = 1;
                const float t = (float)k / (float)nd;
                float q[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) q[c] = (k == nd) ? eb[c] : fmaf(t, eb[c] - ea[c], ea[c]);
                st = pv_cull_status<(PV_EDGE_FAST_TRIG != 0), false>(q, S, dl);
            }
            // class of a motion = OR over its 16 states
            const unsigned m0 = __ballot_sync(FULL, st & 1u), m1 = __ballot_sync(FULL, st & 2u);
            const unsigned a0 = (m0 & 0xffffu) ? 1u : 0u, a1 = (m1 & 0xffffu) ? 1u : 0u;
            const unsigned b0 = (m0 >> 16) ? 1u : 0u, b1 = (m1 >> 16) ? 1u : 0u;
            cls0 |= (a0 << (2 * it)) | (b0 << (2 * it + 1));
            cls1 |= (a1 << (2 * it)) | (b1 << (2 * it + 1));
        }
        word = ~(cls0 | cls1);
        if (mine + 32 > n_edges) word &= (n_edges - w * 32 >= 32) ? 0xffffffffu : ((1u << (unsigned)(n_edges - w * 32)) - 1u);
        if (lane == 0) bits[w] = word;
        // the rest goes to the validators: append to the list of its class (one atomic per warp and class)
        const unsigned c = mine < n_edges ? (((cls1 >> lane) & 1u) ? 2u : ((cls0 >> lane) & 1u)) : 0u;
#pragma unroll
        for (unsigned k = 1; k <= 2; ++k) {
            const unsigned tm = __ballot_sync(FULL, c == k);
            if (!tm) continue;
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(n_list + (k - 1), (unsigned)__popc(tm));
            base = __shfl_sync(FULL, base, 0);
            if (c == k) lists[(size_t)(k - 1) * (size_t)n_edges + base + __popc(tm & ((1u << lane) - 1u))] = (unsigned)mine;
        }
    }
}

// Second-tier certificate (self-collision section): the motions the first pass could not finish but proved free of the
// scene (list 1: class 1 and the survivors of the scene-only validator) keep some self-collision cull busy along the whole
// motion -- 14 of their 16 coarse cells on average -- yet most of the valid ones are valid COMFORTABLY.  This kernel runs the
// self-collision section itself at the same 16 coarse states, two motions per warp, with `dl` metres of slack on every cull
// and every test (the SLK form of pv_check_config; dl = the travel bound over the half-spacing of the coarse states, as
// in pv_edge_cert_kernel):
//   * a contact at a coarse state -- the very comparison the validator makes at that state (same interpolation
//     formula, same instantiation flags) -- makes the motion invalid: done, its bit stays 0;
//   * every coarse state clears everything by dl: every state of the motion is free of contact: done, bit set;
//   * otherwise the motion goes on to the exhaustive self-collision-only validator (out_list).
// Offline (tools/probes/edge_round_cert_model.py, config 3): 41 % of list 1 certifies and most of its 41 % invalid motions
// show their contact at a coarse state, so about a quarter is left for the 64-state validator.
#ifndef PV_EDGE_CERT2
#define PV_EDGE_CERT2 1
#endif
#ifndef PV_CERT2_THREADS
#define PV_CERT2_THREADS 512  // 123 registers, no spills (384: 153 registers, -1.3 %; 640: 96 registers with 36 B of spills, same)
#endif
template <bool YAW>
__global__ void __launch_bounds__(PV_CERT2_THREADS, 1)
    pv_edge_cert2_kernel(const __grid_constant__ PvScene S, const float4* __restrict__ aA, const float4* __restrict__ aB,
                         const float* __restrict__ a9, const float4* __restrict__ bA, const float4* __restrict__ bB,
                         const float* __restrict__ b9, int n_steps, float resolution, uint32_t* __restrict__ bits,
                         const unsigned* __restrict__ list, const unsigned* __restrict__ n_list,
                         unsigned* __restrict__ next_group, unsigned* __restrict__ out_list, unsigned* __restrict__ out_n) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, sub = lane & 15, half = lane >> 4;
    const unsigned n = __ldg(n_list), n_groups = (n + 31u) >> 5;
    // a test of radius (sum) r clears by dl when d^2 - r^2 >= (2 r + dl) dl; r <= PV_SELF_R_MAX, dl <= the cap
    const float slk_k = (2.0f * PV_SELF_R_MAX + PV_MOTION_CERT_MAX_SLACK) * 1.001f;
    for (;;) {
        unsigned g = 0;
        if (lane == 0) g = atomicAdd(next_group, 1u);
        g = __shfl_sync(FULL, g, 0);
        if (g >= n_groups) break;
        const unsigned at = g * 32u + (unsigned)lane;
        const bool have = at < n;
        const unsigned mine = have ? __ldg(list + at) : 0u;
        // lane-parallel setup, as in pv_edge_cert_kernel
        float dl_m = 0.f;
        int nd_m = 1;
        if (have) {
            float ea[9], eb[9];
            pv_load_soa(aA, aB, a9, (int64_t)mine, ea);
            pv_load_soa(bA, bB, b9, (int64_t)mine, eb);
            const float reach[7] = PV_MOTION_REACH, lo[9] = PV_Q_LOWER, hi[9] = PV_Q_UPPER;
            float d2 = 0.f, trav = fabsf(eb[7] - ea[7]) + fabsf(eb[8] - ea[8]);
            bool in_ = true;
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                const float de = eb[c] - ea[c];
                d2 = fmaf(de, de, d2);
                if (c < 7) trav = fmaf(fabsf(de), reach[c], trav);
                in_ = in_ && ea[c] >= lo[c] && ea[c] <= hi[c] && eb[c] >= lo[c] && eb[c] <= hi[c];
            }
            int nd = n_steps;
            if (nd <= 0) nd = max(1, (int)ceilf(sqrtf(d2) / resolution));
            const int h = ((nd + 15) >> 4) >> 1;
            const float dl = fmaf(trav / (float)nd, (float)h * 1.0002f, 2e-5f);
            nd_m = nd;
            if (nd >= PV_CERT_MIN_ND && in_ && dl <= PV_MOTION_CERT_MAX_SLACK) dl_m = dl;
        }
        const unsigned seek = __ballot_sync(FULL, dl_m > 0.f);
        unsigned hit_m = 0, near_m = ~seek;  // bit j: the motion of lane j has a contact / is not cleared (or seeks nothing)
        for (int it = 0; it < 16; ++it) {
            if (!((seek >> (2 * it)) & 3u)) continue;  // warp-uniform
            const int src = 2 * it + half;
            const float dl = __shfl_sync(FULL, dl_m, src);
            const int nd = __shfl_sync(FULL, nd_m, src);
            const unsigned e = __shfl_sync(FULL, mine, src);
            bool hit = false, near = false;
            if (dl > 0.f) {
                float ea[9], eb[9];
                pv_load_soa(aA, aB, a9, (int64_t)e, ea);  // (L1 hits: the owner lane has just read them)
                pv_load_soa(bA, bB, b9, (int64_t)e, eb);
                const int s = (nd + 15) >> 4;
                int k = nd - ((s - 1) >> 1) - sub * s;
                if (k < 1) k = 1;
                const float t = (float)k * (1.0f / (float)nd);  // the validator's expression: state k bit for bit
                float q[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) q[c] = (k == nd) ? eb[c] : fmaf(t, eb[c] - ea[c], ea[c]);
                PvAcc<PV_MODE_BITS> acc;
                acc.dl = dl;
                pv_check_config<PV_MODE_BITS, true, PV_EXIT_NONE, 0, false, false, (PV_EDGE_FAST_TRIG != 0), false, YAW, 1, true>(q, S, acc);
                hit = acc.hit || acc.mslk < 0.f;
                near = acc.nearp || !(acc.mslk >= slk_k * dl);
            }
            const unsigned hb = __ballot_sync(FULL, hit), nb = __ballot_sync(FULL, near);
            hit_m |= (((hb & 0xffffu) ? 1u : 0u) << (2 * it)) | (((hb >> 16) ? 1u : 0u) << (2 * it + 1));
            near_m |= (((nb & 0xffffu) ? 1u : 0u) << (2 * it)) | (((nb >> 16) ? 1u : 0u) << (2 * it + 1));
        }
        const bool my_hit = (hit_m >> lane) & 1u, my_near = (near_m >> lane) & 1u;
        if (have && !my_hit && !my_near) atomicOr(bits + (mine >> 5), 1u << (mine & 31u));
        const unsigned fwd = __ballot_sync(FULL, have && !my_hit && my_near);
        if (fwd) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(out_n, (unsigned)__popc(fwd));
            base = __shfl_sync(FULL, base, 0);
            if ((fwd >> lane) & 1u) out_list[base + __popc(fwd & ((1u << lane) - 1u))] = mine;
        }
    }
}

// Fused verdict gather behind the certificate pipeline: its verdict words are complete only after the last list kernel,
// so a small kernel then stores them into every rank's gather buffer (one multimem.st per word through the NVSwitch
// multicast address, or one peer store per rank) -- still no collective launch and no host round trip.
__global__ void __launch_bounds__(256) pv_edge_emit_kernel(const uint32_t* __restrict__ bits, const __grid_constant__ PvGather G,
                                                           int64_t n_words) {
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words && w < G.word_cap;
         w += (int64_t)gridDim.x * blockDim.x) {
        const unsigned word = bits[w];
        if (G.mc) {
            asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(G.mc + G.word_off + w), "r"(word) : "memory");
        } else if (G.peers) {
            for (int p = 0; p < G.n_peers; ++p) G.peers[p][G.word_off + w] = word;
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
#if PV_EDGE_CERT
    // certificate pass + full validation of the rest (see pv_edge_cert_kernel), for motions cut into a fixed number of
    // more than 32 states.  (Under the resolution rule the planners' motions are a handful of states long -- nothing to
    // skip -- and the pass would only cost its loads, so n_steps == 0 keeps the single kernel.)  Scratch comes from the
    // stream-ordered allocator, so any number of streams and handles may be in flight.
    if (h->cull == 2 && epw == 32 && d_bits && !h->scene.carry && !a_aos && n < ((int64_t)1 << 32) &&
        n_steps >= PV_CERT_MIN_ND) {
        unsigned* scratch = nullptr;
        if (!h->pool) {  // a pool of the handle's own that keeps its memory between calls
            cudaMemPoolProps pp = {};
            pp.allocType = cudaMemAllocationTypePinned;
            pp.location.type = cudaMemLocationTypeDevice;
            pp.location.id = h->device;
            PV_CUDA(h, cudaMemPoolCreate(&h->pool, &pp));
            uint64_t keep = ~0ull;
            PV_CUDA(h, cudaMemPoolSetAttribute(h->pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        PV_CUDA(h, cudaMallocFromPoolAsync((void**)&scratch, (3 * (size_t)n + 64) * sizeof(unsigned), h->pool, st));
        // scratch[0], [1], [4] = lengths of lists 1, 2, 3; [2], [3], [5] = their group counters, [6] = the group counter of
        // the second-tier certificate pass over list 1; then the three lists of n entries each
        unsigned* d_list = scratch + 64;
        PV_CUDA(h, cudaMemsetAsync(scratch, 0, 8 * sizeof(unsigned), st));
        const int cgrid = pv_grid_for(h, (const void*)pv_edge_cert_kernel, PV_CERT_THREADS, words);
        pv_edge_cert_kernel<<<cgrid, PV_CERT_THREADS, 0, st>>>(h->scene, (const float4*)aA, (const float4*)aB, a9,
                                                              (const float4*)bA, (const float4*)bB, b9, n, n_steps,
                                                              resolution, d_bits, d_list, scratch);
        // class 2 first, through the instantiation that holds the scene section only: the motions it finds free of
        // the scene (and above the plane) join list 1; the second-tier certificate finishes most of list 1 at 16 states
        // a motion and passes the rest on (list 3) to the self-collision-only instantiation
#define PV_LAUNCH_EL(YAW_, SECT_, APPEND_, LIST_, NLIST_, GROUP_)                                                   \
    {                                                                                                               \
        auto kern_ = pv_edge_kernel<true, PV_MODE_BITS, false, false, YAW_, true, SECT_, APPEND_>;                  \
        int grid = pv_grid_for(h, (const void*)kern_, PV_E_THREADS, words * 8);                                     \
        kern_<<<grid, PV_E_THREADS, 0, st>>>(                                                                       \
            h->scene, (const float4*)aA, (const float4*)aB, a9, (const float4*)bA, (const float4*)bB, b9, nullptr,  \
            nullptr, n, n_steps, resolution, d_bits, nullptr, 4, nullptr, PvGatherOpt<false>{}, LIST_, NLIST_,      \
            GROUP_, d_list, scratch);                                                                               \
    }
#define PV_LAUNCH_C2(YAW_)                                                                                          \
    {                                                                                                               \
        int grid = pv_grid_for(h, (const void*)pv_edge_cert2_kernel<YAW_>, PV_CERT2_THREADS, words * 8);            \
        pv_edge_cert2_kernel<YAW_><<<grid, PV_CERT2_THREADS, 0, st>>>(                                              \
            h->scene, (const float4*)aA, (const float4*)aB, a9, (const float4*)bA, (const float4*)bB, b9, n_steps,  \
            resolution, d_bits, d_list, scratch, scratch + 6, d_list + 2 * (size_t)n, scratch + 4);                 \
    }
        const bool tier2 = PV_EDGE_CERT2 && h->edge_cert2;
        unsigned* l1 = tier2 ? d_list + 2 * (size_t)n : d_list;  // what the self-collision-only validator works off
        unsigned* l1n = tier2 ? scratch + 4 : scratch;
        unsigned* l1g = tier2 ? scratch + 5 : scratch + 2;
        if (h->all_yaw) {
            PV_LAUNCH_EL(true, 2, true, d_list + (size_t)n, scratch + 1, scratch + 3)
            if (tier2) PV_LAUNCH_C2(true)
            PV_LAUNCH_EL(true, 1, false, l1, l1n, l1g)
        } else {
            PV_LAUNCH_EL(false, 2, true, d_list + (size_t)n, scratch + 1, scratch + 3)
            if (tier2) PV_LAUNCH_C2(false)
            PV_LAUNCH_EL(false, 1, false, l1, l1n, l1g)
        }
#undef PV_LAUNCH_C2
#undef PV_LAUNCH_EL
        if (gather_on) {
            int egrid = (int)((words + 255) / 256);
            if (egrid > h->sm_count * 8) egrid = h->sm_count * 8;
            pv_edge_emit_kernel<<<egrid, 256, 0, st>>>(d_bits, h->gather, words);
            h->launches++;
        }
        PV_CUDA(h, cudaGetLastError());
        PV_CUDA(h, cudaFreeAsync(scratch, st));
        h->launches += tier2 ? 4 : 3;
        return PV_OK;
    }
#endif
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

