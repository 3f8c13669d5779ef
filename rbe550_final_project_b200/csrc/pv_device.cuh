// pv_device.cuh -- device-side Panda state validity: FK + primitive tests, specialised at compile time
// to the frozen model in panda_model_gen.h.  sm_100a, FP32 CUDA-core work (no dense contraction, so no
// tensor cores).  One thread evaluates one configuration; all robot primitives live in registers.
//
// Verdict rule restated from the reference: planning.py:209-219 (_is_ompl_state_valid: any robot
// contact invalidates), planning.py:221-230 (contacts of hand/fingers with the attached box are
// forgiven), planning.py:139-150 (joint limits), Genesis pair filter as in SURVEY.md App. C.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "panda_model_gen.h"

#ifndef PV_MAX_OBB
#define PV_MAX_OBB 32
#endif
#define PV_FLAG_SELF 1u
#define PV_FLAG_LIMITS 2u

// Scene snapshot, passed to every kernel BY VALUE as a __grid_constant__ parameter: it then lives in
// the constant bank, every thread reads the same box at the same time (constant-cache broadcast, or a
// uniform-register load), and concurrent streams / handles never share mutable state.
struct PvScene {
    float obb[PV_MAX_OBB][16];  // c.xyz, half.xyz, R row-major (world-from-box), bounding radius
    float table_z;
    float base[3];
    int n_obb;
    int attached;  // scene-box index forgiven for hand / fingers, -1 = none
    unsigned flags;
    unsigned yaw_only_mask;  // bit b: box b is rotated about world z only
    // bit l (0..7): link group l can reach box b at all; bit 8+k: gripper box k can (static, joint-independent)
    unsigned short reach_mask[PV_MAX_OBB];
    // carry mode (not reference behaviour): box `attached` is held rigidly by the hand and moves with
    // it instead of staying a static obstacle.  Pose of the box in the hand frame: centre carry_t, axes = COLUMNS of
    // carry_R (row-major hand-from-box); carry_h = half extents of the scene record shrunk by the contact allowance
    // (a block resting on the table or on another block touches it: that is not a collision), carry_br = |carry_h|.
    int carry;
    float carry_t[3];
    float carry_R[9];
    float carry_h[3];
    float carry_br;
    // scene-level cull (exact, see pv_check_config): axis-aligned bounds of all scene boxes (their true extents
    // |R| h); the same bounds padded by each link group's bounding radius (+ slack): lo - r in [0..2], hi + r in
    // [3..5]; bit l of group_any: link group l can reach some box at all, bit 8: the gripper can
    float aabb_lo[3], aabb_hi[3];
    float gpad[8][6];
    unsigned group_any;
};
#define PV_LINK_CARRIED 11  // link id reported for the carried box in culprit / contact codes

enum { PV_MODE_BITS = 0, PV_MODE_MARGIN = 1, PV_MODE_LIST = 2 };
#define PV_MAX_CONTACTS 32
enum { PV_EXIT_NONE = 0, PV_EXIT_ALL = 1, PV_EXIT_ANY = 2 };

#define PV_CODE(kind, a, b) (((kind) << 16) | ((a) << 8) | (b))
// link of arm sphere i (compile-time lookups: the indices at the call sites are literals)
__device__ constexpr int pv_sphere_link[PV_N_SPHERES] = PV_SPHERE_LINK;
#define PV_SELF_CODE(a, lb) PV_CODE(3, pv_sphere_link[a], lb)

template <int MODE>
struct PvAcc;
// PV_PACK (opt-in experiment, off): two sphere-sphere tests per instruction with the packed FP32 forms of sm_100a
// (FADD2 / FFMA2, PTX *.f32x2).  A packed instruction takes one issue slot for two results (tools/probes/
// f32x2_probe.cu: FFMA2 sustains the FLOP/s of FFMA at half the issue rate) and the self-collision block shrinks from 88
// scalar to 51 packed tests, yet the kernel came out 1 % SLOWER on the same box (9.04 vs 9.14 G checks/s, verdicts
// bit-identical): with 4 warps per scheduler the check is bound by dependent-instruction latency as much as by issue
// slots (profiles/r1_notes.md), and the packed forms do not shorten the chains.
#ifndef PV_PACK
#define PV_PACK 0
#endif
#ifndef PV_PACK_ACC
#define PV_PACK_ACC 1
#endif
template <>
struct PvAcc<PV_MODE_BITS> {
    bool hit = false;
    // SLK instantiations of pv_check_config (second-tier motion certificates, pv_edge_cert2_kernel): the self-collision
    // tests accumulate min(d^2 - r^2) in mslk instead of setting hit -- a contact is mslk < 0, exactly the comparison the
    // other instantiations make, and a test that clears by less than dl metres is mslk < (2 PV_SELF_R_MAX + cap) dl; the
    // culls in front of them and the plane tests compare with dl of slack (nearp: the plane is within dl)
    float mslk = 1e30f, dl = 0.f;
    bool nearp = false;
    // packed tests accumulate min(d^2 - r^2) here (one FMNMX3 per two tests); hit |= min(m) < 0 at the end.  PV_PACK_ACC
    // independent accumulators keep the FMNMX3 chain from serialising the tests.
    float m[4] = {1e30f, 1e30f, 1e30f, 1e30f};
    __device__ __forceinline__ bool packed_hit() const { return fminf(fminf(m[0], m[1]), fminf(m[2], m[3])) < 0.f; }
};

typedef unsigned long long pv_f2;  // two floats in an aligned register pair
__device__ __forceinline__ pv_f2 pv_pk(float lo, float hi) {
    pv_f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void pv_upk(pv_f2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ pv_f2 pv_fma2(pv_f2 a, pv_f2 b, pv_f2 c) {
    pv_f2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ pv_f2 pv_sub2(pv_f2 a, pv_f2 b) {
    pv_f2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ pv_f2 pv_add2(pv_f2 a, pv_f2 b) {
    pv_f2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ pv_f2 pv_mul2(pv_f2 a, pv_f2 b) {
    pv_f2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// sphere a against spheres b0 and b1 at once: n0, n1 = -(ra + rb)^2 (0 masks a half)
template <int K>
__device__ __forceinline__ void pv_sphere_sphere2(PvAcc<PV_MODE_BITS>& acc, float3 a, float3 b0, float3 b1, float n0,
                                                  float n1) {
    pv_f2 dx = pv_sub2(pv_pk(a.x, a.x), pv_pk(b0.x, b1.x));
    pv_f2 dy = pv_sub2(pv_pk(a.y, a.y), pv_pk(b0.y, b1.y));
    pv_f2 dz = pv_sub2(pv_pk(a.z, a.z), pv_pk(b0.z, b1.z));
    pv_f2 d2 = pv_fma2(dz, dz, pv_fma2(dy, dy, pv_fma2(dx, dx, pv_pk(n0, n1))));
    float lo, hi;
    pv_upk(d2, lo, hi);
    acc.m[K % PV_PACK_ACC] = fminf(acc.m[K % PV_PACK_ACC], fminf(lo, hi));
}
// contact list (diagnostics): the codes of every test in penetration, like the pair list detect_collision returns
template <>
struct PvAcc<PV_MODE_LIST> {
    int n = 0;
    int codes[PV_MAX_CONTACTS];
    __device__ __forceinline__ void take(float g, int c) {
        if (g < 0.f) {
            for (int k = 0; k < n && k < PV_MAX_CONTACTS; ++k)
                if (codes[k] == c) return;  // one entry per (link, other) pair
            if (n < PV_MAX_CONTACTS) codes[n] = c;
            ++n;
        }
    }
};
template <>
struct PvAcc<PV_MODE_MARGIN> {
    float m = 1e30f;
    int code = 0;
    __device__ __forceinline__ void take(float g, int c) {
        if (g < m) {
            m = g;
            code = c;
        }
    }
};

// ---- small vector helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float3 v_sub(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 v_neg(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float v_dot(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
// a * s + b
__device__ __forceinline__ float3 v_fma(float3 a, float s, float3 b) {
    return make_float3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z));
}
// joint rotation about the (pre-rotated) z axis:  X = c Xp + s Yp,  Y = c Yp - s Xp
__device__ __forceinline__ void v_rotz(float3 Xp, float3 Yp, float c, float s, float3& X, float3& Y) {
    X = make_float3(fmaf(s, Yp.x, c * Xp.x), fmaf(s, Yp.y, c * Xp.y), fmaf(s, Yp.z, c * Xp.z));
    Y = make_float3(fmaf(-s, Xp.x, c * Yp.x), fmaf(-s, Xp.y, c * Yp.y), fmaf(-s, Xp.z, c * Yp.z));
}

// sin/cos for joint angles (|x| of a few rad; still correct to ~1e-7 up to |x| ~ 1e4): Cody-Waite
// reduction by pi/2 and the classic single-precision minimax polynomials on [-pi/4, pi/4].
__device__ __forceinline__ void pv_sincos(float x, float& s, float& c) {
    float kf = rintf(x * 0.63661977236758134f);
    int k = (int)kf;
    float r = fmaf(kf, -1.5707962512969971f, x);
    r = fmaf(kf, -7.5497894158615964e-08f, r);
    r = fmaf(kf, -5.3903029534742384e-15f, r);
    float r2 = r * r;
    float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
    float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f),
                    r2 * r2, fmaf(-0.5f, r2, 1.0f));
    float a = (k & 1) ? cp : sp;
    float b = (k & 1) ? sp : cp;
    s = (k & 2) ? -a : a;
    c = ((k + 1) & 2) ? -b : b;
}

// FAST: the hardware approximations (MUFU.SIN / MUFU.COS, abs. error ~4e-7), see PV_FAST_TRIG
// REDUCE: bring x into [-pi, pi] first (exact for |x| <= pi, where k = 0): the hardware's own reduction loses accuracy in
// proportion to |x|.  Only joint 6 (upper limit 3.7525) can exceed pi inside the joint limits, and a state outside the
// limits is invalid whatever its kinematics are, so the other joints go to MUFU as they are (same bits: k would be 0).
template <bool FAST, bool REDUCE = true>
__device__ __forceinline__ void pv_sincos_sel(float x, float& s, float& c) {
    if constexpr (FAST) {
        float r = x;
        if constexpr (REDUCE) {
            const float k = rintf(x * 0.15915494309189535f);
            r = fmaf(k, -6.2831854820251465f, x);
            r = fmaf(k, 1.7484555e-7f, r);
        }
        __sincosf(r, &s, &c);
    } else {
        pv_sincos(x, s, c);
    }
}

// ---- primitive tests ---------------------------------------------------------------------------------
// sphere (centre c, radius r) vs box (centre oc, half oh, box axes in world = columns of R)
template <int MODE>
__device__ __forceinline__ void pv_sphere_box(PvAcc<MODE>& acc, float3 c, float r, float r2, float3 oc, float3 oh,
                                              float3 ax, float3 ay, float3 az, int code) {
    float3 d = v_sub(c, oc);
    float ex = fabsf(v_dot(d, ax)) - oh.x;
    float ey = fabsf(v_dot(d, ay)) - oh.y;
    float ez = fabsf(v_dot(d, az)) - oh.z;
    float px = fmaxf(ex, 0.f), py = fmaxf(ey, 0.f), pz = fmaxf(ez, 0.f);
    float s2 = fmaf(pz, pz, fmaf(py, py, px * px));
    if constexpr (MODE == PV_MODE_BITS) {
        acc.hit |= (s2 < r2);
    } else {
        float g = (s2 > 0.f ? sqrtf(s2) : fmaxf(ex, fmaxf(ey, ez))) - r;
        acc.take(g, code);
    }
}

// sphere vs a box rotated about world z only: axes (cy, sy, 0), (-sy, cy, 0), (0, 0, 1)
template <int MODE>
__device__ __forceinline__ void pv_sphere_box_yaw(PvAcc<MODE>& acc, float3 c, float r, float r2, float3 oc, float3 oh,
                                                  float cy, float sy, int code) {
    float dx = c.x - oc.x, dy = c.y - oc.y, dz = c.z - oc.z;
    float ex = fabsf(fmaf(dy, sy, dx * cy)) - oh.x;
    float ey = fabsf(fmaf(dy, cy, -dx * sy)) - oh.y;
    float ez = fabsf(dz) - oh.z;
    float px = fmaxf(ex, 0.f), py = fmaxf(ey, 0.f), pz = fmaxf(ez, 0.f);
    float s2 = fmaf(pz, pz, fmaf(py, py, px * px));
    if constexpr (MODE == PV_MODE_BITS) {
        acc.hit |= (s2 < r2);
    } else {
        float g = (s2 > 0.f ? sqrtf(s2) : fmaxf(ex, fmaxf(ey, ez))) - r;
        acc.take(g, code);
    }
}

// FFMA-chain forms of the two tests above: k = (oc.ax, oc.ay, oc.az) is the box centre in the box frame, computed once
// per box, so a sphere centre goes into the box frame as R^T c - k (three FFMA chains) instead of a subtraction followed
// by dot products: 19 / 15 instead of 22 / 17 instructions per test.  Used where the scene-box tests dominate (state and
// sweep kernels: +3.6 %); the edge / RRT kernels mostly skip those tests and are faster with the classic form (the extra
// 9 instructions per box cost them 5 %).  The two forms differ by fp32 rounding only (~1e-8 m).
template <int MODE>
__device__ __forceinline__ void pv_sphere_box_k(PvAcc<MODE>& acc, float3 c, float r, float r2, float3 k, float3 oh,
                                                float3 ax, float3 ay, float3 az, int code) {
    float ex = fabsf(fmaf(c.z, ax.z, fmaf(c.y, ax.y, fmaf(c.x, ax.x, -k.x)))) - oh.x;
    float ey = fabsf(fmaf(c.z, ay.z, fmaf(c.y, ay.y, fmaf(c.x, ay.x, -k.y)))) - oh.y;
    float ez = fabsf(fmaf(c.z, az.z, fmaf(c.y, az.y, fmaf(c.x, az.x, -k.z)))) - oh.z;
    float px = fmaxf(ex, 0.f), py = fmaxf(ey, 0.f), pz = fmaxf(ez, 0.f);
    float s2 = fmaf(pz, pz, fmaf(py, py, px * px));
    if constexpr (MODE == PV_MODE_BITS) {
        acc.hit |= (s2 < r2);
    } else {
        float g = (s2 > 0.f ? sqrtf(s2) : fmaxf(ex, fmaxf(ey, ez))) - r;
        acc.take(g, code);
    }
}
template <int MODE>
__device__ __forceinline__ void pv_sphere_box_yaw_k(PvAcc<MODE>& acc, float3 c, float r, float r2, float3 k, float3 oh,
                                                    float cy, float sy, int code) {
    float ex = fabsf(fmaf(c.y, sy, fmaf(c.x, cy, -k.x))) - oh.x;
    float ey = fabsf(fmaf(c.y, cy, fmaf(c.x, -sy, -k.y))) - oh.y;
    float ez = fabsf(c.z - k.z) - oh.z;
    float px = fmaxf(ex, 0.f), py = fmaxf(ey, 0.f), pz = fmaxf(ez, 0.f);
    float s2 = fmaf(pz, pz, fmaf(py, py, px * px));
    if constexpr (MODE == PV_MODE_BITS) {
        acc.hit |= (s2 < r2);
    } else {
        float g = (s2 > 0.f ? sqrtf(s2) : fmaxf(ex, fmaxf(ey, ez))) - r;
        acc.take(g, code);
    }
}

// sphere already expressed in the box's frame (loc = R^T (c - origin)) vs a box with centre bcl in that frame
template <int MODE, bool SLK = false>
__device__ __forceinline__ void pv_sphere_box_local(PvAcc<MODE>& acc, float3 loc, float r, float r2, float3 bcl, float3 oh,
                                                    int code) {
    float ex = fabsf(loc.x - bcl.x) - oh.x;
    float ey = fabsf(loc.y - bcl.y) - oh.y;
    float ez = fabsf(loc.z - bcl.z) - oh.z;
    float px = fmaxf(ex, 0.f), py = fmaxf(ey, 0.f), pz = fmaxf(ez, 0.f);
    float s2 = fmaf(pz, pz, fmaf(py, py, px * px));
    if constexpr (MODE == PV_MODE_BITS && SLK) {
        acc.mslk = fminf(acc.mslk, s2 - r2);  // (s2 - r2 < 0 exactly when s2 < r2)
    } else if constexpr (MODE == PV_MODE_BITS) {
        acc.hit |= (s2 < r2);
    } else {
        float g = (s2 > 0.f ? sqrtf(s2) : fmaxf(ex, fmaxf(ey, ez))) - r;
        acc.take(g, code);
    }
}

template <int MODE, bool SLK = false>
__device__ __forceinline__ void pv_sphere_sphere(PvAcc<MODE>& acc, float3 a, float3 b, float rr2, float rr, int code) {
    float3 d = v_sub(a, b);
    float d2 = v_dot(d, d);
    if constexpr (MODE == PV_MODE_BITS && SLK) {
        acc.mslk = fminf(acc.mslk, d2 - rr2);  // (d2 - rr2 < 0 exactly when d2 < rr2)
    } else if constexpr (MODE == PV_MODE_BITS) {
        acc.hit |= (d2 < rr2);
    } else {
        acc.take(sqrtf(d2) - rr, code);
    }
}

template <int MODE, bool SLK = false>
__device__ __forceinline__ void pv_plane(PvAcc<MODE>& acc, float lowest, float table_z, int code) {
    float g = lowest - table_z;
    if constexpr (MODE == PV_MODE_BITS) {
        acc.hit |= (g < 0.f);
        if constexpr (SLK) acc.nearp |= !(g >= acc.dl);
    } else {
        acc.take(g, code);
    }
}

// box A (centre ca, half ha, axes AX/AY/AZ) vs box B (centre cb, half hb, axes BX/BY/BZ): 15-axis SAT.
// Gap = largest separation (NORMALISE: cross-axis separations divided by |a_i x b_j|, i.e. metres); cross axes with
// |a_i x b_j|^2 <= 1e-4 are skipped.  The out-of-line body RETURNS the gap: it used to take the accumulator by
// reference, which made `acc` address-taken and put acc.hit in local memory for the whole check (33 STL + 17 LDL in
// the hot loop of the state kernel, VERDICT r1 / profiles/r1_sass_tally.json).
template <bool NORMALISE>
__device__ __noinline__ float pv_box_box_gap(float3 ca, float3 ha, float3 AX, float3 AY, float3 AZ, float3 cb, float3 hb,
                                             float3 BX, float3 BY, float3 BZ) {
    float Rm[3][3], A[3][3], t[3], h_a[3] = {ha.x, ha.y, ha.z}, h_b[3] = {hb.x, hb.y, hb.z};
    float3 d = v_sub(cb, ca);
    const float3 Aax[3] = {AX, AY, AZ}, Bax[3] = {BX, BY, BZ};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            Rm[i][j] = v_dot(Aax[i], Bax[j]);
            A[i][j] = fabsf(Rm[i][j]);
        }
        t[i] = v_dot(Aax[i], d);
    }
    float best = -1e30f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float rb = fmaf(A[i][2], h_b[2], fmaf(A[i][1], h_b[1], A[i][0] * h_b[0]));
        best = fmaxf(best, fabsf(t[i]) - h_a[i] - rb);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float ra = fmaf(A[2][j], h_a[2], fmaf(A[1][j], h_a[1], A[0][j] * h_a[0]));
        float tl = fmaf(t[2], Rm[2][j], fmaf(t[1], Rm[1][j], t[0] * Rm[0][j]));
        best = fmaxf(best, fabsf(tl) - ra - h_b[j]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            float len2 = fmaf(-Rm[i][j], Rm[i][j], 1.0f);
            float ra = fmaf(h_a[i2], A[i1][j], h_a[i1] * A[i2][j]);
            float rb = fmaf(h_b[j2], A[i][j1], h_b[j1] * A[i][j2]);
            float tl = fabsf(fmaf(t[i2], Rm[i1][j], -t[i1] * Rm[i2][j]));
            float g = tl - ra - rb;
            if constexpr (NORMALISE) g = g * rsqrtf(fmaxf(len2, 1e-4f));
            if (len2 > 1e-4f) best = fmaxf(best, g);
        }
    }
    return best;
}
template <int MODE>
__device__ __forceinline__ void pv_box_box(PvAcc<MODE>& acc, float3 ca, float3 ha, float3 AX, float3 AY, float3 AZ,
                                           float3 cb, float3 hb, float3 BX, float3 BY, float3 BZ, int code) {
    const float best = pv_box_box_gap<MODE == PV_MODE_MARGIN>(ca, ha, AX, AY, AZ, cb, hb, BX, BY, BZ);
    if constexpr (MODE == PV_MODE_BITS) {
        acc.hit |= (best < 0.f);
    } else {
        acc.take(best, code);
    }
}

// ---- forward kinematics --------------------------------------------------------------------------------
// Frames of link1..link7, hand; fingers share the hand's axes.  Chain constants: SURVEY.md App. A
// (Menagerie panda.xml, scenes.py:85); the six +-90 deg x pre-rotations are axis permutations.
struct PvFrames {
    float3 p[11];
    float3 X[9], Y[9], Z[9];  // links 0..8 (fingers: axes of link 8; right finger has X, Y negated)
};

// Visitor-driven FK: `on_link(l, p, X, Y, Z)` is called as soon as link l's frame exists, so callers
// that only need sphere centres never keep more than one frame live.
#ifndef PV_TABLE_MIN
#define PV_TABLE_MIN 1
#endif
// The verdict-bit kernels (FMAK: state, sweep) evaluate the joint sines / cosines with the hardware approximations
// (MUFU.SIN / MUFU.COS through __sincosf: ~120 fewer instructions per check, +10 % throughput).  Their absolute error
// (~4e-7) moves a link by at most a few micrometres (measured: tests/test_gpu_parity.py::test_fk_verdict_path), far
// inside the 1e-4 m band within which verdicts may differ from the fp64 oracle.  The pose kernel pv_fk, the margin /
// contact kernels and the planners keep the accurate pv_sincos.
#ifndef PV_FAST_TRIG
#define PV_FAST_TRIG 1
#endif
template <bool FAST = false, class F>
__device__ __forceinline__ void pv_fk_visit(const float* q, float bx, float by, float bz, F&& on_link) {
    float s, c;
    float3 p = make_float3(bx, by, bz);
    float3 X = make_float3(1.f, 0.f, 0.f), Y = make_float3(0.f, 1.f, 0.f), Z = make_float3(0.f, 0.f, 1.f);
    on_link(std::integral_constant<int, 0>{}, p, X, Y, Z);
    // link1: pos (0,0,0.333), no pre-rotation
    pv_sincos_sel<FAST, false>(q[0], s, c);
    p.z += 0.333f;
    X = make_float3(c, s, 0.f);
    Y = make_float3(-s, c, 0.f);
    on_link(std::integral_constant<int, 1>{}, p, X, Y, Z);
    float3 Xp, Yp, Zp;
    // link2: pos 0, Rx(-90): X' = X, Y' = -Z, Z' = Y
    pv_sincos_sel<FAST, false>(q[1], s, c);
    {
        float3 X1 = X, Y1 = Y;
        X = make_float3(c * X1.x, c * X1.y, -s);
        Y = make_float3(-s * X1.x, -s * X1.y, -c);
        Z = Y1;
    }
    on_link(std::integral_constant<int, 2>{}, p, X, Y, Z);
    // link3: pos (0,-0.316,0), Rx(+90): X' = X, Y' = Z, Z' = -Y
    pv_sincos_sel<FAST, false>(q[2], s, c);
    p = v_fma(Y, -0.316f, p);
    Xp = X; Yp = Z; Zp = v_neg(Y);
    v_rotz(Xp, Yp, c, s, X, Y); Z = Zp;
    on_link(std::integral_constant<int, 3>{}, p, X, Y, Z);
    // link4: pos (0.0825,0,0), Rx(+90)
    pv_sincos_sel<FAST, false>(q[3], s, c);
    p = v_fma(X, 0.0825f, p);
    Xp = X; Yp = Z; Zp = v_neg(Y);
    v_rotz(Xp, Yp, c, s, X, Y); Z = Zp;
    on_link(std::integral_constant<int, 4>{}, p, X, Y, Z);
    // link5: pos (-0.0825,0.384,0), Rx(-90): X' = X, Y' = -Z, Z' = Y
    pv_sincos_sel<FAST, false>(q[4], s, c);
    p = v_fma(Y, 0.384f, v_fma(X, -0.0825f, p));
    Xp = X; Yp = v_neg(Z); Zp = Y;
    v_rotz(Xp, Yp, c, s, X, Y); Z = Zp;
    on_link(std::integral_constant<int, 5>{}, p, X, Y, Z);
    // link6: pos 0, Rx(+90)
    pv_sincos_sel<FAST>(q[5], s, c);
    Xp = X; Yp = Z; Zp = v_neg(Y);
    v_rotz(Xp, Yp, c, s, X, Y); Z = Zp;
    on_link(std::integral_constant<int, 6>{}, p, X, Y, Z);
    // link7: pos (0.088,0,0), Rx(+90)
    pv_sincos_sel<FAST, false>(q[6], s, c);
    p = v_fma(X, 0.088f, p);
    Xp = X; Yp = Z; Zp = v_neg(Y);
    v_rotz(Xp, Yp, c, s, X, Y); Z = Zp;
    on_link(std::integral_constant<int, 7>{}, p, X, Y, Z);
    // hand: pos (0,0,0.107), fixed rotation about z by -45 deg
    p = v_fma(Z, 0.107f, p);
    Xp = X; Yp = Y;
    v_rotz(Xp, Yp, PV_HAND_COS, PV_HAND_SIN, X, Y);
    on_link(std::integral_constant<int, 8>{}, p, X, Y, Z);
    // fingers: pos (0,0,0.0584), slide along own +y; right finger frame is the hand's turned 180 deg about z
    float3 pf = v_fma(Z, 0.0584f, p);
    on_link(std::integral_constant<int, 9>{}, v_fma(Y, q[7], pf), X, Y, Z);
    on_link(std::integral_constant<int, 10>{}, v_fma(Y, -q[8], pf), v_neg(X), v_neg(Y), Z);
}

// ---- the state check ----------------------------------------------------------------------------------
// Everything the tests need about the robot in ONE configuration (registers: every index below is a literal).
struct PvPlaced {
    float3 s[PV_N_SPHERES];      // arm sphere centres
    float3 hX, hY, hZ, hP;       // hand frame
    float3 bc[3];                // gripper box centres (hand, left finger, right finger; axes = the hand's)
    float grip_r;                // ball around bc[0] that contains all three gripper boxes in this configuration
    float3 cC, cH, cX, cY, cZ;   // carried box (CARRY): centre, half extents, axes
    float cbr;
};
__device__ constexpr float pv_bh[3][3] = {
#define PV_BOX_HALF(k, link, cx, cy, cz, hx, hy, hz, br) {hx, hy, hz},
    PV_BOXES(PV_BOX_HALF)
#undef PV_BOX_HALF
};
__device__ constexpr float pv_bbr[3] = {
#define PV_BOX_BR(k, link, cx, cy, cz, hx, hy, hz, br) br,
    PV_BOXES(PV_BOX_BR)
#undef PV_BOX_BR
};
__device__ constexpr int pv_blink[3] = {
#define PV_BOX_LK(k, link, cx, cy, cz, hx, hy, hz, br) link,
    PV_BOXES(PV_BOX_LK)
#undef PV_BOX_LK
};

// FK -> sphere centres + gripper boxes (+ the carried box)
template <bool FTRIG, bool CARRY>
__device__ __forceinline__ void pv_place(const float* q, const PvScene& S, PvPlaced& P) {
    pv_fk_visit<(PV_FAST_TRIG && FTRIG)>(q, S.base[0], S.base[1], S.base[2], [&](auto lc, float3 p, float3 X, float3 Y, float3 Z) {
        constexpr int l = decltype(lc)::value;
        if constexpr (l == 0) { PV_PLACE_LINK0(P.s, p, X, Y, Z) }
        if constexpr (l == 1) { PV_PLACE_LINK1(P.s, p, X, Y, Z) }
        if constexpr (l == 2) { PV_PLACE_LINK2(P.s, p, X, Y, Z) }
        if constexpr (l == 3) { PV_PLACE_LINK3(P.s, p, X, Y, Z) }
        if constexpr (l == 4) { PV_PLACE_LINK4(P.s, p, X, Y, Z) }
        if constexpr (l == 5) { PV_PLACE_LINK5(P.s, p, X, Y, Z) }
        if constexpr (l == 6) { PV_PLACE_LINK6(P.s, p, X, Y, Z) }
        if constexpr (l == 7) { PV_PLACE_LINK7(P.s, p, X, Y, Z) }
        if constexpr (l == 8) { P.hX = X; P.hY = Y; P.hZ = Z; P.hP = p; }
#define PV_BOX_PLACE(k, link, cx, cy, cz, hx, hy, hz, br) \
    if constexpr (l == link) P.bc[k] = v_fma(Z, cz, v_fma(Y, cy, v_fma(X, cx, p)));
        PV_BOXES(PV_BOX_PLACE)
#undef PV_BOX_PLACE
    });
    // radius of a ball centred on the hand box that contains all three gripper boxes for THIS configuration
    // (with both fingers inside their travel the hand box's own bounding ball contains them -- asserted by the model
    // generator -- so the general form only runs for out-of-limit finger values, i.e. for states that are invalid anyway)
    P.grip_r = pv_bbr[0] + 2.0f * PV_CULL_SLACK;
    if (!(fabsf(q[7]) <= PV_GRIP_CONST_MAXQ && fabsf(q[8]) <= PV_GRIP_CONST_MAXQ)) {
        float3 d1 = v_sub(P.bc[1], P.bc[0]), d2 = v_sub(P.bc[2], P.bc[0]);
        P.grip_r = fmaxf(pv_bbr[0], fmaxf(sqrtf(v_dot(d1, d1)) + pv_bbr[1], sqrtf(v_dot(d2, d2)) + pv_bbr[2])) + 2.0f * PV_CULL_SLACK;
    }
    P.cC = P.hP;
    P.cH = make_float3(0.f, 0.f, 0.f);
    P.cX = P.hX;
    P.cY = P.hY;
    P.cZ = P.hZ;
    P.cbr = 0.f;
    if constexpr (CARRY) {
        P.cH = make_float3(S.carry_h[0], S.carry_h[1], S.carry_h[2]);
        P.cbr = S.carry_br;
        P.cC = v_fma(P.hZ, S.carry_t[2], v_fma(P.hY, S.carry_t[1], v_fma(P.hX, S.carry_t[0], P.hP)));
#define PV_CARRY_AXIS(j) \
    v_fma(P.hZ, S.carry_R[6 + j], v_fma(P.hY, S.carry_R[3 + j], make_float3(P.hX.x * S.carry_R[j], P.hX.y * S.carry_R[j], P.hX.z * S.carry_R[j])))
        P.cX = PV_CARRY_AXIS(0);
        P.cY = PV_CARRY_AXIS(1);
        P.cZ = PV_CARRY_AXIS(2);
#undef PV_CARRY_AXIS
    }
}

#define PV_EARLY_EXIT_RET(retval)                                                          \
    if constexpr (MODE == PV_MODE_BITS && EXIT != PV_EXIT_NONE) {                          \
        if constexpr (PV_PACK) acc.hit |= acc.packed_hit();                                \
        bool h_ = acc.hit;                                                                 \
        if (EXIT == PV_EXIT_ALL ? __all_sync(0xffffffffu, h_) : __any_sync(0xffffffffu, h_)) return retval; \
    }

// ---- robot vs scene boxes (the section behind the scene-level cull) -----------------------------------------
// YAW: every box of the scene is rotated about the world's z axis only (true of every reference scene: blocks resting on
// the table or on each other); the kernels then come in an instantiation WITHOUT the general sphere-vs-box tests.  That
// code never ran for such scenes anyway, but it sat in the middle of the box loop: these kernels are bound by instruction
// fetch (a third of the loop body was dead weight between the culls: 23.6 -> 11.9 KB), and without it the edge kernel
// gains 5.7 % and the state kernel 3.4 % on the same box (profiles/r2_notes.md).  pv_set_scene decides (PvHandle::all_yaw).
template <int MODE, bool CULL, int EXIT, int SYNC, bool FMAK, bool CARRY, bool YAW = false>
__device__ __forceinline__ void pv_scene_section(PvAcc<MODE>& acc, const PvPlaced& P, const PvScene& S, int b0 = 0,
                                                 int bstep = 1) {
    // (b0, bstep): the boxes b0, b0 + bstep, ... only -- the dense owed-section passes of the sorted kernels deal one
    // configuration's boxes out to several lanes when they have lanes to spare (everything below is per box)
    const int nb = S.n_obb;
    for (int b = b0; b < nb; b += bstep) {
        if constexpr (SYNC >= 3) __syncthreads();
        if constexpr (CARRY) {
            if (b == S.attached) continue;  // it is where the hand is, not where the snapshot saw it
        }
        const float3 oc = make_float3(S.obb[b][0], S.obb[b][1], S.obb[b][2]);
        const float3 oh = make_float3(S.obb[b][3], S.obb[b][4], S.obb[b][5]);
        const float obr = S.obb[b][15];
        const float3 BX = make_float3(S.obb[b][6], S.obb[b][9], S.obb[b][12]);
        const float3 BY = make_float3(S.obb[b][7], S.obb[b][10], S.obb[b][13]);
        const float3 BZ = make_float3(S.obb[b][8], S.obb[b][11], S.obb[b][14]);
        const bool yaw_only = YAW ? true : (bool)((S.yaw_only_mask >> b) & 1u);
        float3 ok = oc;
        if constexpr (FMAK) ok = make_float3(v_dot(oc, BX), v_dot(oc, BY), yaw_only ? oc.z : v_dot(oc, BZ));
        const unsigned rmask = S.reach_mask[b];
        if constexpr (CARRY) {
            float3 d_ = v_sub(P.cC, oc);
            float rr_ = (P.cbr + PV_CULL_SLACK) + obr;
            if (MODE == PV_MODE_MARGIN || v_dot(d_, d_) < rr_ * rr_)
                pv_box_box<MODE>(acc, P.cC, P.cH, P.cX, P.cY, P.cZ, oc, oh, BX, BY, BZ, PV_CODE(2, PV_LINK_CARRIED, b));
        }
        // one uniform yaw / general decision per GROUP keeps the hot (yaw-only) sphere tests contiguous in the code
#define PV_ENV_SPHERE_YAW(i, link, cx, cy, cz, r)                                                         \
    if constexpr (FMAK) pv_sphere_box_yaw_k<MODE>(acc, P.s[i], r, (r) * (r), ok, oh, BX.x, BX.y, PV_CODE(2, link, b)); \
    else pv_sphere_box_yaw<MODE>(acc, P.s[i], r, (r) * (r), oc, oh, BX.x, BX.y, PV_CODE(2, link, b));
#define PV_ENV_SPHERE_GEN(i, link, cx, cy, cz, r)                                                         \
    if constexpr (FMAK) pv_sphere_box_k<MODE>(acc, P.s[i], r, (r) * (r), ok, oh, BX, BY, BZ, PV_CODE(2, link, b)); \
    else pv_sphere_box<MODE>(acc, P.s[i], r, (r) * (r), oc, oh, BX, BY, BZ, PV_CODE(2, link, b));
#define PV_ENV_GROUP(l, cs, br)                                 \
    if (MODE == PV_MODE_MARGIN || (rmask & (1u << l))) {        \
        float3 d_ = v_sub(P.s[cs], oc);                         \
        float rr_ = (br + PV_CULL_SLACK) + obr;                 \
        if (!CULL || v_dot(d_, d_) < rr_ * rr_) {               \
            if (yaw_only) {                                     \
                PV_SPHERES_LINK##l(PV_ENV_SPHERE_YAW)           \
            } else {                                            \
                PV_SPHERES_LINK##l(PV_ENV_SPHERE_GEN)           \
            }                                                   \
        }                                                       \
    }
        PV_LINK_GROUPS(PV_ENV_GROUP)
#undef PV_ENV_GROUP
#undef PV_ENV_SPHERE_YAW
#undef PV_ENV_SPHERE_GEN
        // one bounding ball around the whole gripper (hand + both fingers, radius from this configuration's finger
        // openings) goes first: the three per-box culls and SATs behind it are reached by ~1 % of the lanes
        bool near_gripper = true;
        if constexpr (MODE != PV_MODE_MARGIN) {
            float3 dg_ = v_sub(P.bc[0], oc);
            float rg_ = P.grip_r + obr;
            near_gripper = ((rmask >> 8) & 7u) && v_dot(dg_, dg_) < rg_ * rg_;
        }
        if (b != S.attached && near_gripper) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (MODE != PV_MODE_MARGIN && !((rmask >> (8 + k)) & 1u)) continue;
                // bounding-ball cull in front of the 15-axis SAT: always on (exact, and the SAT is ~250 instructions)
                float3 d_ = v_sub(P.bc[k], oc);
                float rr_ = (pv_bbr[k] + PV_CULL_SLACK) + obr;
                if (MODE == PV_MODE_MARGIN || v_dot(d_, d_) < rr_ * rr_) {
                    pv_box_box<MODE>(acc, P.bc[k], make_float3(pv_bh[k][0], pv_bh[k][1], pv_bh[k][2]), P.hX, P.hY, P.hZ, oc, oh,
                                     BX, BY, BZ, PV_CODE(2, pv_blink[k], b));
                }
            }
        }
        PV_EARLY_EXIT_RET()
    }
}

// The scene section OUT OF LINE (verdict-bit kernels, PV_COLD_SCENE): only ~2 % of random configurations come near any
// scene box, so the box loop -- a third of the kernel's code -- is kept out of the instruction stream every warp runs
// through.  pv_check_config<..., DEFER = true> then only REPORTS that the section is needed and the kernel calls
// pv_scene_cold afterwards with the configuration re-read from where it came (shared-memory staging, the edge's end
// points): nothing stays live across the hot path for the sake of the rare call, and the function re-derives the
// placement from q with the same code, hence the same bits (profiles/r2_notes.md).
#ifndef PV_COLD_SCENE
#define PV_COLD_SCENE 1
#endif
// The warp-per-edge kernels (edge validation, planner) could call the section out of line as pv_scene_cold below; measured
// on the same box it does not pay there (config 3: 393 -> 367 M edges/s in the pentagon scene, goal-1 unchanged): their
// warps take the section TOGETHER (the lanes are states of one edge) far more often than 2 % of the time, and every
// call re-derives the placement.  Off.
#ifndef PV_COLD_SCENE_WARP
#define PV_COLD_SCENE_WARP 0
#endif
// LOAD: a small trivially-copyable callable that re-creates the configuration (load(q) fills q[9])
template <bool CULL, int EXIT, bool FMAK, bool CARRY, bool FTRIG, class LOAD>
__device__ __noinline__ bool pv_scene_cold(LOAD load, const PvScene& S) {
    float q[9];
    load(q);
    PvPlaced P;
    pv_place<FTRIG, CARRY>(q, S, P);
    PvAcc<PV_MODE_BITS> acc;
    pv_scene_section<PV_MODE_BITS, CULL, EXIT, 0, FMAK, CARRY>(acc, P, S);
    if constexpr (PV_PACK) acc.hit |= acc.packed_hit();
    return acc.hit;
}

// slack thresholds of the motion certificates (pv_cull_status and the SLK form of pv_check_config):
template <bool SLK, class ACC>
__device__ __forceinline__ float pv_slk_dl(const ACC& a) {
    if constexpr (SLK) return a.dl;
    else return 0.f;
}
// (r + dl)^2 <= r^2 + (2 r + PV_MOTION_CERT_MAX_SLACK) dl for 0 <= dl <= PV_MOTION_CERT_MAX_SLACK; r2 is a literal, so
// the factor folds to a constant (the 1.001 covers the rounding of the fold and of the FFMA)
#define PV_SLK_THR(r2, dl) fmaf((2.0f * sqrtf(r2) + PV_MOTION_CERT_MAX_SLACK) * 1.001f, dl, r2)

// Returns through `acc`.  All 32 lanes of a warp must call this together when EXIT != PV_EXIT_NONE.
// SYNC: every warp of the block calls this together and block-level barriers keep the warps within one
// code region of each other, so the (large, straight-line) instruction stream is fetched once per SM instead
// of once per warp (ncu showed stall_no_instruction as the top stall of the free-running version).
// DEFER: do not run the scene section, return whether it is needed (see pv_scene_cold); the return value is false
// whenever the section has been dealt with here.
template <int MODE, bool CULL, int EXIT, int SYNC = 0, bool FMAK = false, bool CARRY = false, bool FTRIG = FMAK,
          bool DEFER = false, bool YAW = false, int SECT = 3, bool SLK = false>
__device__ __forceinline__ bool pv_check_config(const float* q, const PvScene& S, PvAcc<MODE>& acc) {
    // SLK (self-collision-only verdict bits, no early exit): the check ALSO says whether the configuration clears every
    // test by acc.dl metres -- see PvAcc<PV_MODE_BITS>::mslk.  The caller reads acc.hit | (acc.mslk < 0) for the verdict
    // (bit-identical to the other instantiations') and acc.nearp | (acc.mslk < slack threshold) for the clearance.
    static_assert(!SLK || (SECT == 1 && MODE == PV_MODE_BITS && EXIT == PV_EXIT_NONE && !PV_PACK), "slack form: self-collision section, bits");
    // SECT: which sections exist at all in this instantiation -- bit 0 the self-collision section, bit 1 the scene-level
    // test and the scene section.  The motion validator routes motions that provably never need one of them (their
    // coarse states clear all of its culls with slack: pv_cull_status) to a kernel compiled without it: 30 KB of code
    // instead of 53 KB, which is what these instruction-fetch-bound kernels respond to.
    static_assert(SECT == 3 || (CULL && MODE == PV_MODE_BITS && !CARRY && !DEFER), "section-less forms: culling verdict bits");
    static_assert(!SYNC || EXIT == PV_EXIT_NONE, "block barriers and warp-level early exit do not mix");
    static_assert(!DEFER || (CULL && MODE == PV_MODE_BITS), "only the culling verdict-bit form defers the scene section");
    const unsigned FULL = 0xffffffffu;
#define PV_LOCKSTEP(level) \
    if constexpr (SYNC >= level) __syncthreads();
#define PV_EARLY_EXIT() PV_EARLY_EXIT_RET(false)

    {
        // Joint limits are part of the model's validity domain and are ALWAYS enforced (PV_FLAG_LIMITS is kept in the
        // ABI but no longer switches anything): the self-collision pair lists are pruned by a never-collide certificate
        // that holds inside the limits only (tools/certify_never_collide.py), and the static reach masks assume the
        // finger travel.  OMPL never hands the reference's callback a state outside the bounds either
        // (planning.py:139-150: RealVectorBounds; out-of-bounds start / goal states are dropped at intake).  The
        // negated form also rejects non-finite joint values, which every later comparison would wave through.
        const float lo[9] = PV_Q_LOWER, hi[9] = PV_Q_UPPER;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const bool out = !(q[j] >= lo[j] && q[j] <= hi[j]);
            if constexpr (MODE == PV_MODE_BITS) {
                acc.hit |= out;
            } else if (out) {
                acc.take(-1e30f, PV_CODE(4, j, 0));
            }
        }
    }

    PvPlaced P;
    pv_place<FTRIG, CARRY>(q, S, P);
    float3(&s)[PV_N_SPHERES] = P.s;
    const float3 hX = P.hX, hY = P.hY, hZ = P.hZ, hP = P.hP;
    float3(&bc)[3] = P.bc;
    const float grip_r = P.grip_r;

    // ---- robot vs ground plane (link0 is fixed to the world: pair filtered) -------------
    const float tz = S.table_z;
    if constexpr (MODE == PV_MODE_BITS && PV_TABLE_MIN) {
        pv_plane<MODE, SLK>(acc, PV_TABLE_LOWEST(s), tz, 0);
    } else {
#define PV_TABLE_SPHERE(i, link, cx, cy, cz, r) \
    if (link != 0) pv_plane<MODE>(acc, s[i].z - r, tz, PV_CODE(1, link, 0));
        PV_SPHERES(PV_TABLE_SPHERE)
#undef PV_TABLE_SPHERE
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float ext = fmaf(fabsf(hZ.z), pv_bh[k][2], fmaf(fabsf(hY.z), pv_bh[k][1], fabsf(hX.z) * pv_bh[k][0]));
        pv_plane<MODE, SLK>(acc, bc[k].z - ext, tz, PV_CODE(1, pv_blink[k], 0));
    }
    // ---- carried box: placed by the hand, checked against the plane and the arm spheres of link0..link6 --------
    if constexpr (CARRY) {
        float ext = fmaf(fabsf(P.cZ.z), P.cH.z, fmaf(fabsf(P.cY.z), P.cH.y, fabsf(P.cX.z) * P.cH.x));
        pv_plane<MODE>(acc, P.cC.z - ext, tz, PV_CODE(1, PV_LINK_CARRIED, 0));
#define PV_CARRY_SPHERE(i, link, cx, cy, cz, r) \
    if (link <= 6) pv_sphere_box<MODE>(acc, s[i], r, (r) * (r), P.cC, P.cH, P.cX, P.cY, P.cZ, PV_CODE(3, link, PV_LINK_CARRIED));
        PV_SPHERES(PV_CARRY_SPHERE)
#undef PV_CARRY_SPHERE
    }
    PV_EARLY_EXIT()
    PV_LOCKSTEP(2)

    // ---- self collision ------------------------------------------------------------------------------
    if ((SECT & 1) && (S.flags & PV_FLAG_SELF)) {
#define PV_SS(a, b, rr2, rr) pv_sphere_sphere<MODE, SLK>(acc, s[a], s[b], rr2, rr, PV_SELF_CODE(a, pv_sphere_link[b]));
#define PV_SS2(a, b0, b1, n0, n1, k) pv_sphere_sphere2<k>(acc, s[a], s[b0], s[b1], n0, n1);
#define PV_LP(la, lb, ca, cb, cull2)                         \
    {                                                        \
        float3 d_ = v_sub(s[ca], s[cb]);                     \
        if (!CULL || v_dot(d_, d_) < (SLK ? PV_SLK_THR(cull2, pv_slk_dl<SLK>(acc)) : cull2)) { \
            if constexpr (MODE == PV_MODE_BITS && PV_PACK) { \
                PV_SS2_PAIRS_##la##_##lb(PV_SS2)             \
            } else {                                         \
                PV_SS_PAIRS_##la##_##lb(PV_SS)               \
            }                                                \
        }                                                    \
    }
        PV_SS_LINKPAIRS(PV_LP)
#undef PV_LP
#undef PV_SS
#undef PV_SS2
        PV_EARLY_EXIT()
        PV_LOCKSTEP(1)
        // sphere-vs-gripper pairs: the three gripper boxes share the hand's axes, so each proximal sphere is moved
        // into the hand frame once (12 instr.) and then meets axis-aligned boxes (10 instr. each)
        float3 bcl[3], bhl[3];
#define PV_BOX_HF(k, cx, cy0, cz, sgn, qi, hx, hy, hz) \
    bcl[k] = make_float3(cx, fmaf(sgn, q[qi], cy0), cz); \
    bhl[k] = make_float3(hx, hy, hz);
        PV_BOXES_HANDFRAME(PV_BOX_HF)
#undef PV_BOX_HF
        float3 loc_;
        float lr_, lr2_;
        float3 hk = make_float3(0.f, 0.f, 0.f);
        if constexpr (FMAK) hk = make_float3(v_dot(hP, hX), v_dot(hP, hY), v_dot(hP, hZ));  // hand origin, hand frame
#define PV_SBH_S(a, r, r2)                                                                              \
    {                                                                                                   \
        if constexpr (FMAK) {                                                                           \
            loc_ = make_float3(fmaf(s[a].z, hX.z, fmaf(s[a].y, hX.y, fmaf(s[a].x, hX.x, -hk.x))),       \
                               fmaf(s[a].z, hY.z, fmaf(s[a].y, hY.y, fmaf(s[a].x, hY.x, -hk.y))),       \
                               fmaf(s[a].z, hZ.z, fmaf(s[a].y, hZ.y, fmaf(s[a].x, hZ.x, -hk.z))));      \
        } else {                                                                                        \
            float3 d_ = v_sub(s[a], hP);                                                                \
            loc_ = make_float3(v_dot(d_, hX), v_dot(d_, hY), v_dot(d_, hZ));                            \
        }                                                                                               \
        lr_ = r;                                                                                        \
        lr2_ = r2;                                                                                      \
    }
#define PV_SBH_B(a, k) pv_sphere_box_local<MODE, SLK>(acc, loc_, lr_, lr2_, bcl[k], bhl[k], PV_SELF_CODE(a, 8 + k));
        // links that pair with all three gripper boxes are culled against the one gripper ball, the others against
        // the bounding balls of the boxes they pair with
#define PV_LB(la, ca, c0, c1, c2, rla)                                                                  \
    {                                                                                                   \
        bool near_;                                                                                     \
        if ((c0) > 0.f) {                                                                               \
            float3 d0_ = v_sub(s[ca], bc[0]);                                                           \
            float rr_ = (rla) + grip_r;                                                                 \
            if constexpr (SLK) rr_ += pv_slk_dl<SLK>(acc);                                                      \
            near_ = v_dot(d0_, d0_) < rr_ * rr_;                                                        \
        } else {                                                                                        \
            float3 d1_ = v_sub(s[ca], bc[1]), d2_ = v_sub(s[ca], bc[2]);                                \
            near_ = v_dot(d1_, d1_) < (SLK ? PV_SLK_THR(c1, pv_slk_dl<SLK>(acc)) : c1) ||                            \
                    v_dot(d2_, d2_) < (SLK ? PV_SLK_THR(c2, pv_slk_dl<SLK>(acc)) : c2);                              \
        }                                                                                               \
        if (!CULL || near_) {                                                                           \
            PV_SBH_##la(PV_SBH_S, PV_SBH_B)                                                             \
        }                                                                                               \
    }
        PV_SBH_LINKS(PV_LB)
#undef PV_LB
#undef PV_SBH_B
#undef PV_SBH_S
        PV_EARLY_EXIT()
    }

    // ---- robot vs scene boxes ------------------------------------------------------------------------
    // Scene-level cull in front of the whole box loop: a link group whose bounding ball stays outside the axis-aligned
    // bounds of ALL boxes (tested as the ball's centre against the bounds padded by its radius, which is conservative)
    // cannot touch any of them, and neither can the gripper ball or the carried box's ball.  Only ~2 % of random
    // configurations come near the goal-1 blocks at all, so a warp whose 32 lanes all pass skips the box loads and every
    // per-box cull (profiles/r1_notes.md: a third of the instructions of the check); the sorted kernel orders its
    // configurations so that whole warps agree (pv_sort_key).
    bool near_scene = true;
    if constexpr (!(SECT & 2)) {
        near_scene = false;
    } else if constexpr (CULL && MODE == PV_MODE_BITS) {
        near_scene = false;
#define PV_SCENE_GROUP(l, cs, br)                                                                                  \
    if (S.group_any & (1u << l))                                                                                   \
        near_scene |= s[cs].x > S.gpad[l][0] && s[cs].x < S.gpad[l][3] && s[cs].y > S.gpad[l][1] &&                \
                      s[cs].y < S.gpad[l][4] && s[cs].z > S.gpad[l][2] && s[cs].z < S.gpad[l][5];
        PV_LINK_GROUPS(PV_SCENE_GROUP)
#undef PV_SCENE_GROUP
        if (S.group_any & 0x100u)
            near_scene |= bc[0].x + grip_r > S.aabb_lo[0] && bc[0].x - grip_r < S.aabb_hi[0] &&
                          bc[0].y + grip_r > S.aabb_lo[1] && bc[0].y - grip_r < S.aabb_hi[1] &&
                          bc[0].z + grip_r > S.aabb_lo[2] && bc[0].z - grip_r < S.aabb_hi[2];
        if constexpr (CARRY) {
            const float rc_ = P.cbr + PV_CULL_SLACK;
            near_scene |= P.cC.x + rc_ > S.aabb_lo[0] && P.cC.x - rc_ < S.aabb_hi[0] && P.cC.y + rc_ > S.aabb_lo[1] &&
                          P.cC.y - rc_ < S.aabb_hi[1] && P.cC.z + rc_ > S.aabb_lo[2] && P.cC.z - rc_ < S.aabb_hi[2];
        }
        // kernels with warp votes inside the loop (EXIT) must take the branch as a whole warp
        if constexpr (EXIT != PV_EXIT_NONE) near_scene = __any_sync(FULL, near_scene);
    }
    if constexpr (DEFER) {
        if constexpr (PV_PACK) acc.hit |= acc.packed_hit();
        return near_scene;
    } else {
        if (near_scene) pv_scene_section<MODE, CULL, EXIT, SYNC, FMAK, CARRY, YAW>(acc, P, S);
        if constexpr (MODE == PV_MODE_BITS && PV_PACK) acc.hit |= acc.packed_hit();
        return false;
    }
#undef PV_EARLY_EXIT
#undef PV_LOCKSTEP
}

// ---- motion certificates: the culls with slack ------------------------------------------------------------------
// Nothing is flagged when the configuration is inside the joint limits and EVERY culling test of pv_check_config -- the link-pair and
// gripper balls of the self-collision section, the scene-level test -- as well as the ground-plane test clears by more
// than dl metres.  Every configuration whose robot points lie within dl of this one's (in the world, and relative to
// any link proximal to them) then skips all of its culls too and stays above the plane: it is free of contact without
// being tested (pv_edge_cert_kernel, csrc/pv_edge.cu; panda_model.motion_reach_bounds bounds the travel).  The
// comparisons are written so that a non-finite value never certifies.  The placement is pv_place's: only the handful of
// centres the culls look at and the z coordinates survive dead-code elimination (~520 instructions, 10 KB of code).
// Returns 0 when everything is clear; bit 0: limits, ground plane or a self-collision cull within dl; bit 1: the
// scene-level test within dl.
// LIMITS = false: the caller vouches for the joint limits (both ends of a motion inside them: so is every state between).
template <bool FTRIG, bool LIMITS = true>
__device__ __forceinline__ unsigned pv_cull_status(const float* q, const PvScene& S, float dl) {
    const float lo[9] = PV_Q_LOWER, hi[9] = PV_Q_UPPER;
    bool ok = true;
    if constexpr (LIMITS) {
#pragma unroll
        for (int j = 0; j < 9; ++j) ok = ok && q[j] >= lo[j] && q[j] <= hi[j];
    }
    PvPlaced P;
    pv_place<FTRIG, false>(q, S, P);
    float3(&s)[PV_N_SPHERES] = P.s;
    float3(&bc)[3] = P.bc;
    const float grip_r = P.grip_r, tz = S.table_z;
    ok = ok && (PV_TABLE_LOWEST(s) - tz) >= dl;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float ext = fmaf(fabsf(P.hZ.z), pv_bh[k][2], fmaf(fabsf(P.hY.z), pv_bh[k][1], fabsf(P.hX.z) * pv_bh[k][0]));
        ok = ok && ((bc[k].z - ext) - tz) >= dl;
    }
    if (S.flags & PV_FLAG_SELF) {
#define PV_LPS(la, lb, ca, cb, cull2)                          \
    {                                                          \
        float3 d_ = v_sub(s[ca], s[cb]);                       \
        ok = ok && v_dot(d_, d_) >= PV_SLK_THR(cull2, dl);     \
    }
        PV_SS_LINKPAIRS(PV_LPS)
#undef PV_LPS
#define PV_LBS(la, ca, c0, c1, c2, rla)                                                                \
    {                                                                                                  \
        if ((c0) > 0.f) {                                                                              \
            float3 d0_ = v_sub(s[ca], bc[0]);                                                          \
            float rr_ = ((rla) + grip_r) + dl;                                                         \
            ok = ok && v_dot(d0_, d0_) >= rr_ * rr_;                                                   \
        } else {                                                                                       \
            float3 d1_ = v_sub(s[ca], bc[1]), d2_ = v_sub(s[ca], bc[2]);                               \
            ok = ok && v_dot(d1_, d1_) >= PV_SLK_THR(c1, dl) && v_dot(d2_, d2_) >= PV_SLK_THR(c2, dl); \
        }                                                                                              \
    }
        PV_SBH_LINKS(PV_LBS)
#undef PV_LBS
    }
    // scene-level test, reach inflated by dl: outside the padded bounds along at least one axis
    const bool ok_self = ok;
    ok = true;
#define PV_SCENE_GROUP(l, cs, br)                                                                                  \
    if (S.group_any & (1u << l))                                                                                   \
        ok = ok && (s[cs].x + dl <= S.gpad[l][0] || s[cs].x - dl >= S.gpad[l][3] || s[cs].y + dl <= S.gpad[l][1] || \
                    s[cs].y - dl >= S.gpad[l][4] || s[cs].z + dl <= S.gpad[l][2] || s[cs].z - dl >= S.gpad[l][5]);
    PV_LINK_GROUPS(PV_SCENE_GROUP)
#undef PV_SCENE_GROUP
    const float grs_ = grip_r + dl;
    if (S.group_any & 0x100u)
        ok = ok && (bc[0].x + grs_ <= S.aabb_lo[0] || bc[0].x - grs_ >= S.aabb_hi[0] || bc[0].y + grs_ <= S.aabb_lo[1] ||
                    bc[0].y - grs_ >= S.aabb_hi[1] || bc[0].z + grs_ <= S.aabb_lo[2] || bc[0].z - grs_ >= S.aabb_hi[2]);
    return (ok_self ? 0u : 1u) | (ok ? 0u : 2u);
}

// ---- re-loaders of a configuration for pv_scene_cold --------------------------------------------------------
struct PvReloadSoa {  // two float4 (+ optional ninth value) wherever they live (global planes, shared-memory staging)
    const float4* a;
    const float4* b;
    const float* c;
    __device__ __forceinline__ void operator()(float* q) const {
        const float4 x = *a, y = *b;
        q[0] = x.x; q[1] = x.y; q[2] = x.z; q[3] = x.w;
        q[4] = y.x; q[5] = y.y; q[6] = y.z; q[7] = y.w;
        q[8] = c ? *c : y.w;
    }
};
struct PvReloadStrided {  // q[j] = p[j * stride] for j < np; the rest (fingers fixed open) = 0.04
    const float* p;
    int stride, np;
    __device__ __forceinline__ void operator()(float* q) const {
#pragma unroll
        for (int j = 0; j < 9; ++j) q[j] = j < np ? p[j * stride] : 0.04f;
    }
};
struct PvReloadLerp {  // state k of nd on the motion ea -> eb, exactly as the edge / planner kernels form it
    float ea[9], eb[9];
    float t;
    bool at_end;
    __device__ __forceinline__ void operator()(float* q) const {
#pragma unroll
        for (int c = 0; c < 9; ++c) q[c] = at_end ? eb[c] : fmaf(t, eb[c] - ea[c], ea[c]);
    }
};

// ---- config loads ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pv_load_soa(const float4* __restrict__ qA, const float4* __restrict__ qB,
                                            const float* __restrict__ q9, int64_t i, float* q) {
    float4 a = __ldg(qA + i), b = __ldg(qB + i);
    q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w;
    q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    q[8] = q9 ? __ldg(q9 + i) : b.w;
}
__device__ __forceinline__ void pv_load_aos(const float* __restrict__ qa, int64_t i, float* q) {
#pragma unroll
    for (int j = 0; j < 9; ++j) q[j] = __ldg(qa + 9 * i + j);
}

// ---- Philox-4x32-10 (counter-based RNG for the device-generated sweeps) ---------------------------------
__device__ __forceinline__ uint4 pv_philox(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// sample `it` of search `search` of the planner's counter-based stream (pv_rrtc.cu; also the sharded-tree front end)
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

__device__ __forceinline__ void pv_sweep_config(uint64_t i, unsigned seed, bool fingers_open, float* q) {
    const float lo[9] = PV_Q_LOWER, hi[9] = PV_Q_UPPER;
    float u[12];
    const uint2 key = make_uint2(seed, 0x50414E44u);
#pragma unroll
    for (int blk = 0; blk < 3; ++blk) {
        uint4 r = pv_philox(make_uint4((unsigned)(i & 0xffffffffull), (unsigned)(i >> 32), blk, 0u), key);
        u[4 * blk + 0] = (float)(r.x >> 8) * 5.9604644775390625e-08f;
        u[4 * blk + 1] = (float)(r.y >> 8) * 5.9604644775390625e-08f;
        u[4 * blk + 2] = (float)(r.z >> 8) * 5.9604644775390625e-08f;
        u[4 * blk + 3] = (float)(r.w >> 8) * 5.9604644775390625e-08f;
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) q[j] = __fmaf_rn(u[j], hi[j] - lo[j], lo[j]);
    if (fingers_open) {
        q[7] = 0.04f;
        q[8] = 0.04f;
    }
}
