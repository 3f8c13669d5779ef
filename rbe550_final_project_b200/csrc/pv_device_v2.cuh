// pv_device_v2.cuh -- warp-cooperative state validity: per-thread FK + broad phase, shared-memory work
// queues, warp-wide narrow phase.
//
// Why: for uncorrelated configurations a per-lane bounding-ball cull inside a thread-per-config kernel does
// not pay (a warp runs a block of tests as soon as ONE of its 32 lanes needs it), and the fully unrolled
// narrow phase was instruction-fetch bound (ncu: stall_no_instruction, profiles/r1b).  Measured pass rates of
// the broad phase are tiny (about 1 self link-pair, 0.8 link-vs-gripper and 0.05-0.4 link-vs-scene-box
// candidates per configuration), so here every lane
//   1. runs FK for its own configuration, parks the 33 sphere centres + 3 gripper boxes in shared memory
//      (lane-contiguous SoA, conflict-free), tests the ground plane, and
//   2. pushes the few (primitive block, partner) candidates that survive conservative bounding-ball tests
//      into per-warp queues (chunked so all entries cost about the same), then
//   3. the 32 lanes drain the queues together: lane e takes entry e, whichever configuration it belongs to,
//      runs the exact narrow-phase test from small tables and ORs the verdict into that configuration's bit.
// The narrow-phase arithmetic is the same device functions as the brute-force kernel, so verdicts are
// bit-identical to it.  All culls are conservative (radius + 1e-4 m slack): they can never hide a contact.
#pragma once
#include "pv_device.cuh"

#define PV2_WARPS 4
#define PV2_Q_SS 320
#define PV2_Q_SB 192
#define PV2_Q_ENV 192
#define PV2_Q_SAT 96
#define PV2_SS_CHUNK 4
#define PV2_SB_CHUNK 2
#define PV2_ENV_CHUNK 2

__device__ const float pv2_c_sph_r[PV_N_SPHERES] = PV2_SPHERE_R;
__device__ const float pv2_c_ss_rr2[PV_N_SS_PAIRS] = PV2_SS_RR2;
__device__ const unsigned short pv2_c_ss_ab[PV_N_SS_PAIRS] = PV2_SS_AB;
__device__ const unsigned char pv2_c_sb_a[PV_N_SB_PAIRS] = PV2_SB_A;
__device__ const unsigned char pv2_c_sb_k[PV_N_SB_PAIRS] = PV2_SB_K;
__device__ const float pv2_c_box_half[3][3] = {
#define PV2_BH(k, link, cx, cy, cz, hx, hy, hz, br) {hx, hy, hz},
    PV_BOXES(PV2_BH)
#undef PV2_BH
};

struct Pv2Tables {  // one per block
    float sph_r[PV_N_SPHERES];
    float ss_rr2[PV_N_SS_PAIRS];
    unsigned short ss_ab[PV_N_SS_PAIRS + (PV_N_SS_PAIRS & 1)];
    unsigned char sb_a[(PV_N_SB_PAIRS + 3) & ~3], sb_k[(PV_N_SB_PAIRS + 3) & ~3];
    float box_half[3][3];
    float obb[PV_MAX_OBB][16];
};

struct Pv2Warp {  // one per warp
    float sx[PV_N_SPHERES][32], sy[PV_N_SPHERES][32], sz[PV_N_SPHERES][32];
    float bx[3][32], by[3][32], bz[3][32];
    float ax[9][32];  // gripper axes: X.xyz, Y.xyz, Z.xyz
    unsigned q_env[PV2_Q_ENV];
    unsigned short q_ss[PV2_Q_SS];
    unsigned short q_sb[PV2_Q_SB];
    unsigned short q_sat[PV2_Q_SAT];
    int n_ss, n_sb, n_env, n_sat;
    unsigned hitmask;
    unsigned pad_[3];
};

__device__ __forceinline__ void pv2_load_tables(Pv2Tables& T, const PvScene& S) {
    for (int i = threadIdx.x; i < PV_N_SPHERES; i += blockDim.x) T.sph_r[i] = pv2_c_sph_r[i];
    for (int i = threadIdx.x; i < PV_N_SS_PAIRS; i += blockDim.x) {
        T.ss_rr2[i] = pv2_c_ss_rr2[i];
        T.ss_ab[i] = pv2_c_ss_ab[i];
    }
    for (int i = threadIdx.x; i < PV_N_SB_PAIRS; i += blockDim.x) {
        T.sb_a[i] = pv2_c_sb_a[i];
        T.sb_k[i] = pv2_c_sb_k[i];
    }
    if (threadIdx.x < 9) T.box_half[threadIdx.x / 3][threadIdx.x % 3] = pv2_c_box_half[threadIdx.x / 3][threadIdx.x % 3];
    for (int i = threadIdx.x; i < S.n_obb * 16; i += blockDim.x) T.obb[i >> 4][i & 15] = S.obb[i >> 4][i & 15];
    __syncthreads();
}

// ---- narrow-phase workers (t = lane index of the configuration the entry belongs to) --------------------
__device__ __forceinline__ bool pv2_narrow_ss(const Pv2Warp& W, const Pv2Tables& T, int t, int start, int cnt) {
    PvAcc<PV_MODE_BITS> acc;
    for (int j = 0; j < cnt; ++j) {
        const unsigned ab = T.ss_ab[start + j];
        const int a = ab & 255, b = ab >> 8;
        pv_sphere_sphere<PV_MODE_BITS>(acc, make_float3(W.sx[a][t], W.sy[a][t], W.sz[a][t]),
                                       make_float3(W.sx[b][t], W.sy[b][t], W.sz[b][t]), T.ss_rr2[start + j], 0.f, 0);
    }
    return acc.hit;
}

__device__ __forceinline__ bool pv2_narrow_sb(const Pv2Warp& W, const Pv2Tables& T, int t, int start, int cnt) {
    PvAcc<PV_MODE_BITS> acc;
    const float3 hX = make_float3(W.ax[0][t], W.ax[1][t], W.ax[2][t]);
    const float3 hY = make_float3(W.ax[3][t], W.ax[4][t], W.ax[5][t]);
    const float3 hZ = make_float3(W.ax[6][t], W.ax[7][t], W.ax[8][t]);
    for (int j = 0; j < cnt; ++j) {
        const int a = T.sb_a[start + j], k = T.sb_k[start + j];
        const float r = T.sph_r[a];
        pv_sphere_box<PV_MODE_BITS>(acc, make_float3(W.sx[a][t], W.sy[a][t], W.sz[a][t]), r, r * r,
                                    make_float3(W.bx[k][t], W.by[k][t], W.bz[k][t]),
                                    make_float3(T.box_half[k][0], T.box_half[k][1], T.box_half[k][2]), hX, hY, hZ, 0);
    }
    return acc.hit;
}

__device__ __forceinline__ bool pv2_narrow_env(const Pv2Warp& W, const Pv2Tables& T, int t, int b, int start, int cnt) {
    PvAcc<PV_MODE_BITS> acc;
    const float* o = T.obb[b];
    const float3 oc = make_float3(o[0], o[1], o[2]), oh = make_float3(o[3], o[4], o[5]);
    const float3 BX = make_float3(o[6], o[9], o[12]), BY = make_float3(o[7], o[10], o[13]), BZ = make_float3(o[8], o[11], o[14]);
    for (int j = 0; j < cnt; ++j) {
        const int a = start + j;
        const float r = T.sph_r[a];
        pv_sphere_box<PV_MODE_BITS>(acc, make_float3(W.sx[a][t], W.sy[a][t], W.sz[a][t]), r, r * r, oc, oh, BX, BY, BZ, 0);
    }
    return acc.hit;
}

__device__ __forceinline__ bool pv2_narrow_sat(const Pv2Warp& W, const Pv2Tables& T, int t, int k, int b) {
    PvAcc<PV_MODE_BITS> acc;
    const float* o = T.obb[b];
    pv_box_box<PV_MODE_BITS>(acc, make_float3(W.bx[k][t], W.by[k][t], W.bz[k][t]),
                             make_float3(T.box_half[k][0], T.box_half[k][1], T.box_half[k][2]),
                             make_float3(W.ax[0][t], W.ax[1][t], W.ax[2][t]), make_float3(W.ax[3][t], W.ax[4][t], W.ax[5][t]),
                             make_float3(W.ax[6][t], W.ax[7][t], W.ax[8][t]), make_float3(o[0], o[1], o[2]),
                             make_float3(o[3], o[4], o[5]), make_float3(o[6], o[9], o[12]), make_float3(o[7], o[10], o[13]),
                             make_float3(o[8], o[11], o[14]), 0);
    return acc.hit;
}

// ---- the warp-cooperative check -------------------------------------------------------------------------------
// All 32 lanes call it together, each with its own configuration.  Returns this lane's "in collision" flag.
// EXIT == PV_EXIT_ANY: return early (value true for every lane that... see below) as soon as any lane is in
// collision -- used when the 32 lanes are states of ONE edge, where a single hit decides the edge.
template <int EXIT>
__device__ __forceinline__ bool pv2_check_warp(const float* q, const PvScene& S, Pv2Warp& W, const Pv2Tables& T,
                                               const int lane) {
    const unsigned FULL = 0xffffffffu;
    if (lane == 0) {
        W.n_ss = 0;
        W.n_sb = 0;
        W.n_env = 0;
        W.n_sat = 0;
        W.hitmask = 0;
    }
    __syncwarp();
    bool hit = false;
    if (S.flags & PV_FLAG_LIMITS) {
        const float lo[9] = PV_Q_LOWER, hi[9] = PV_Q_UPPER;
#pragma unroll
        for (int j = 0; j < 9; ++j) hit |= (q[j] < lo[j]) || (q[j] > hi[j]);
    }

    // ---- phase 1: FK, park primitives in shared memory, ground-plane tests --------------------------------
    const float tz = S.table_z;
    float3 gc[8], bc[3], hX, hY, hZ;
    float3 s[PV_N_SPHERES];  // statically indexed scratch: lives in registers only while its link is processed
    pv_fk_visit(q, S.base[0], S.base[1], S.base[2], [&](auto lc, float3 p, float3 X, float3 Y, float3 Z) {
        constexpr int l = decltype(lc)::value;
#define PV2_PARK(i, link, cx, cy, cz, r)               \
    W.sx[i][lane] = s[i].x;                            \
    W.sy[i][lane] = s[i].y;                            \
    W.sz[i][lane] = s[i].z;                            \
    if (link != 0) hit |= (s[i].z - (r) - tz < 0.f);
        if constexpr (l == 0) { PV_PLACE_LINK0(s, p, X, Y, Z) PV_SPHERES_LINK0(PV2_PARK) }
        if constexpr (l == 1) { PV_PLACE_LINK1(s, p, X, Y, Z) PV_SPHERES_LINK1(PV2_PARK) }
        if constexpr (l == 2) { PV_PLACE_LINK2(s, p, X, Y, Z) PV_SPHERES_LINK2(PV2_PARK) }
        if constexpr (l == 3) { PV_PLACE_LINK3(s, p, X, Y, Z) PV_SPHERES_LINK3(PV2_PARK) }
        if constexpr (l == 4) { PV_PLACE_LINK4(s, p, X, Y, Z) PV_SPHERES_LINK4(PV2_PARK) }
        if constexpr (l == 5) { PV_PLACE_LINK5(s, p, X, Y, Z) PV_SPHERES_LINK5(PV2_PARK) }
        if constexpr (l == 6) { PV_PLACE_LINK6(s, p, X, Y, Z) PV_SPHERES_LINK6(PV2_PARK) }
        if constexpr (l == 7) { PV_PLACE_LINK7(s, p, X, Y, Z) PV_SPHERES_LINK7(PV2_PARK) }
#undef PV2_PARK
#define PV2_GC(gl, cs, br) \
    if constexpr (l == gl) gc[gl] = s[cs];
        PV_LINK_GROUPS(PV2_GC)
#undef PV2_GC
        if constexpr (l == 8) {
            hX = X; hY = Y; hZ = Z;
            W.ax[0][lane] = X.x; W.ax[1][lane] = X.y; W.ax[2][lane] = X.z;
            W.ax[3][lane] = Y.x; W.ax[4][lane] = Y.y; W.ax[5][lane] = Y.z;
            W.ax[6][lane] = Z.x; W.ax[7][lane] = Z.y; W.ax[8][lane] = Z.z;
        }
#define PV2_BOX_PLACE(k, link, cx, cy, cz, hx, hy, hz, br)                                               \
    if constexpr (l == link) {                                                                           \
        bc[k] = v_fma(Z, cz, v_fma(Y, cy, v_fma(X, cx, p)));                                             \
        W.bx[k][lane] = bc[k].x;                                                                         \
        W.by[k][lane] = bc[k].y;                                                                         \
        W.bz[k][lane] = bc[k].z;                                                                         \
        float ext = fmaf(fabsf(hZ.z), hz, fmaf(fabsf(hY.z), hy, fabsf(hX.z) * hx));                      \
        hit |= (bc[k].z - ext - tz < 0.f);                                                               \
    }
        PV_BOXES(PV2_BOX_PLACE)
#undef PV2_BOX_PLACE
    });
    if (EXIT == PV_EXIT_ANY && __any_sync(FULL, hit)) return true;

    // ---- phase 2: broad phase -> queues ------------------------------------------------------------------------
    // pushes beyond a queue's capacity are evaluated on the spot by the pushing lane (slow path, still exact)
#define PV2_PUSH_CHUNKS(QARR, QN, QCAP, CHUNK, START, COUNT, ENC, INLINE_CALL)                 \
    {                                                                                          \
        constexpr int nch_ = ((COUNT) + (CHUNK)-1) / (CHUNK);                                  \
        const int k0_ = atomicAdd(&W.QN, nch_);                                                \
        _Pragma("unroll") for (int c_ = 0; c_ < nch_; ++c_) {                                  \
            const int st_ = (START) + c_ * (CHUNK);                                            \
            const int cn_ = ((COUNT)-c_ * (CHUNK)) < (CHUNK) ? ((COUNT)-c_ * (CHUNK)) : (CHUNK); \
            if (k0_ + c_ < (QCAP)) W.QARR[k0_ + c_] = ENC;                                     \
            else hit |= INLINE_CALL;                                                           \
        }                                                                                      \
    }
    if (S.flags & PV_FLAG_SELF) {
#define PV2_LP(la, lb, ca, cb, cull2, start, count)                                                            \
    {                                                                                                          \
        float3 d_ = v_sub(gc[la], gc[lb]);                                                                     \
        if (v_dot(d_, d_) < cull2)                                                                             \
            PV2_PUSH_CHUNKS(q_ss, n_ss, PV2_Q_SS, PV2_SS_CHUNK, start, count,                                  \
                            (unsigned short)(lane | (st_ << 5) | (cn_ << 12)), pv2_narrow_ss(W, T, lane, st_, cn_)) \
    }
        PV2_LPS(PV2_LP)
#undef PV2_LP
#define PV2_LB(la, k, ca, cull2, start, count)                                                                 \
    {                                                                                                          \
        float3 d_ = v_sub(gc[la], bc[k]);                                                                      \
        if (v_dot(d_, d_) < cull2)                                                                             \
            PV2_PUSH_CHUNKS(q_sb, n_sb, PV2_Q_SB, PV2_SB_CHUNK, start, count,                                  \
                            (unsigned short)(lane | (st_ << 5) | (cn_ << 11)), pv2_narrow_sb(W, T, lane, st_, cn_)) \
    }
        PV2_LBS(PV2_LB)
#undef PV2_LB
    }
    {
        constexpr int gstart[8] = PV2_GROUP_START, gcount[8] = PV2_GROUP_COUNT;
        const float bbr[3] = {
#define PV2_BR(k, link, cx, cy, cz, hx, hy, hz, br) br,
            PV_BOXES(PV2_BR)
#undef PV2_BR
        };
        const int nb = S.n_obb;
        for (int b = 0; b < nb; ++b) {
            const float3 oc = make_float3(S.obb[b][0], S.obb[b][1], S.obb[b][2]);
            const float obr = S.obb[b][15];
            const unsigned rmask = S.reach_mask[b];
#define PV2_ENV_GROUP(l, cs, br)                                                                               \
    if (rmask & (1u << l)) {                                                                                   \
        float3 d_ = v_sub(gc[l], oc);                                                                          \
        float rr_ = (br + PV_CULL_SLACK) + obr;                                                                \
        if (v_dot(d_, d_) < rr_ * rr_)                                                                         \
            PV2_PUSH_CHUNKS(q_env, n_env, PV2_Q_ENV, PV2_ENV_CHUNK, gstart[l], gcount[l],                      \
                            (unsigned)(lane | (b << 5) | (st_ << 10) | (cn_ << 16)),                           \
                            pv2_narrow_env(W, T, lane, b, st_, cn_))                                           \
    }
            PV_LINK_GROUPS(PV2_ENV_GROUP)
#undef PV2_ENV_GROUP
            if (b != S.attached) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (!((rmask >> (8 + k)) & 1u)) continue;
                    float3 d_ = v_sub(bc[k], oc);
                    float rr_ = (bbr[k] + PV_CULL_SLACK) + obr;
                    if (v_dot(d_, d_) < rr_ * rr_) {
                        const int k0_ = atomicAdd(&W.n_sat, 1);
                        if (k0_ < PV2_Q_SAT) W.q_sat[k0_] = (unsigned short)(lane | (k << 5) | (b << 7));
                        else hit |= pv2_narrow_sat(W, T, lane, k, b);
                    }
                }
            }
        }
    }
#undef PV2_PUSH_CHUNKS
    __syncwarp();

    // ---- phase 3: the warp drains the queues together --------------------------------------------------------------
    {
        const int n = min(W.n_ss, PV2_Q_SS);
        for (int e = lane; e < n; e += 32) {
            const unsigned ent = W.q_ss[e];
            if (pv2_narrow_ss(W, T, ent & 31, (ent >> 5) & 127, ent >> 12)) atomicOr(&W.hitmask, 1u << (ent & 31));
        }
    }
    {
        const int n = min(W.n_sb, PV2_Q_SB);
        for (int e = lane; e < n; e += 32) {
            const unsigned ent = W.q_sb[e];
            if (pv2_narrow_sb(W, T, ent & 31, (ent >> 5) & 63, ent >> 11)) atomicOr(&W.hitmask, 1u << (ent & 31));
        }
    }
    {
        const int n = min(W.n_env, PV2_Q_ENV);
        for (int e = lane; e < n; e += 32) {
            const unsigned ent = W.q_env[e];
            if (pv2_narrow_env(W, T, ent & 31, (ent >> 5) & 31, (ent >> 10) & 63, ent >> 16))
                atomicOr(&W.hitmask, 1u << (ent & 31));
        }
    }
    {
        const int n = min(W.n_sat, PV2_Q_SAT);
        for (int e = lane; e < n; e += 32) {
            const unsigned ent = W.q_sat[e];
            if (pv2_narrow_sat(W, T, ent & 31, (ent >> 5) & 3, ent >> 7)) atomicOr(&W.hitmask, 1u << (ent & 31));
        }
    }
    __syncwarp();
    hit |= (W.hitmask >> lane) & 1u;
    __syncwarp();  // the next call resets the queues
    return hit;
}
