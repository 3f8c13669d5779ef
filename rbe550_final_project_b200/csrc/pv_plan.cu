// pv_plan.cu -- ONE C entry for the whole of PlannerInterface.plan_path (planning.py:59-207): intake checks + solve
// (device RRT-Connect, pv_rrtc.cu) + ss.simplifySolution() (planning.py:195-196) + path.interpolate(num_waypoints)
// (planning.py:198) + a full-density validation of what is handed back.  Host logic in C++, every validity question
// answered by the edge kernel in batches; nothing here computes a verdict on the CPU.
//
// simplifySolution() is OMPL's PathSimplifier::simplifyMax: partialShortcutPath (up to five
// times while it helps), smoothBSpline(3 steps, minimum change = length / 100), checkAndRepair, reduceVertices,
// collapseCloseVertices.  OMPL draws one candidate at a time and asks one checkMotion at a time; here each of those
// passes proposes a BATCH of candidates, validates them in one launch of the edge kernel, and accepts the non-overlapping
// ones with the largest saving:
//   * vertex shortcuts (reduceVertices / collapseCloseVertices: both only ever join two existing vertices by a straight
//     motion) -- exhaustive, farthest first, inside the solve kernel (pv_rrtc.cu PH_SHORTCUT) and once more here after
//     the smoothing, over all vertex pairs;
//   * partial shortcuts between points INSIDE segments (partialShortcutPath: range ratio 0.33, snap to vertex 0.005);
//   * B-spline smoothing exactly as smoothBSpline does it: subdivide, then move every original interior vertex to
//     1/4 a + 1/2 v + 1/4 b when both new motions are valid and it moves by more than the minimum change (the moves of
//     one step do not depend on each other, so a step is one batch);
//   * checkAndRepair becomes the final validation of the RESAMPLED waypoints -- every returned waypoint and every
//     segment between consecutive waypoints -- which is denser than the planner's 1 % resolution.  A path that fails it
//     is replaced by the unsimplified solution, and if that fails too the query is planned again with a new seed at
//     half the motion-validation resolution (repair, do not shrug).
// Random choices come from the same counter-based Philox stream as the planner, keyed by (seed, pass), so a plan is a
// deterministic function of its arguments.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <functional>
#include <vector>

#include "../../include/panda_validity.h"
#include "pv_handle.h"

typedef std::array<double, 9> Q9;
typedef std::vector<Q9> Path;
// validates n motions a[k] -> b[k] (fp32 rows) at the motion-validation resolution; ok[k] = 1 when valid
typedef std::function<int(const float* a, const float* b, int n, unsigned char* ok)> EdgeFn;

static double q_dist(const Q9& a, const Q9& b) {
    double s = 0;
    for (int j = 0; j < 9; ++j) s += (b[j] - a[j]) * (b[j] - a[j]);
    return sqrt(s);
}
static Q9 q_lerp(const Q9& a, const Q9& b, double t) {
    Q9 r;
    for (int j = 0; j < 9; ++j) r[j] = a[j] + t * (b[j] - a[j]);
    return r;
}
static double path_len(const Path& p) {
    double s = 0;
    for (size_t i = 0; i + 1 < p.size(); ++i) s += q_dist(p[i], p[i + 1]);
    return s;
}

// Philox-4x32-10, the generator of the device planner (pv_device.cuh: pv_philox)
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1;
        c[3] = (uint32_t)p0;
        c[0] = n0;
        c[2] = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
static void uniform4(uint32_t seed, uint32_t pass, uint32_t k, double u[4]) {
    uint32_t c[4] = {k, pass, 0x53494D50u /* "SIMP" */, 2u};
    philox(c, seed, 0x52525443u);
    for (int j = 0; j < 4; ++j) u[j] = (double)(c[j] >> 8) * 5.9604644775390625e-08;
}

// ---- PathGeometric::interpolate(count) for a RealVectorStateSpace (planning.py:198) -------------------------------
// `count` states spread over the segments in proportion to their length, every original vertex kept, first = start,
// last = goal.  A path with more than `count` states (or fewer than 2) is returned unchanged.
static Path interpolate_path(const Path& pts, int count) {
    const int n_in = (int)pts.size();
    if (count < n_in || n_in < 2) return pts;
    std::vector<double> seg(n_in - 1);
    double remaining = 0;
    for (int i = 0; i + 1 < n_in; ++i) remaining += (seg[i] = q_dist(pts[i], pts[i + 1]));
    int budget = count;
    Path out;
    out.reserve(count);
    const int last = n_in - 1;
    for (int i = 0; i < last; ++i) {
        out.push_back(pts[i]);
        const int room = budget + i - n_in;  // interior states this segment may still take
        if (room > 0) {
            int want = 2;
            if (i + 1 == last) want = room + 2;
            else if (remaining > 0.0) want = (int)floor(0.5 + (double)budget * seg[i] / remaining) + 1;
            int inner = 0;
            if (want > 2) {
                inner = std::min(want - 2, room);
                for (int k = 1; k <= inner; ++k) out.push_back(q_lerp(pts[i], pts[i + 1], (double)k / (double)(inner + 1)));
            }
            budget -= inner + 1;
            remaining -= seg[i];
        } else {
            budget -= 1;
        }
    }
    out.push_back(pts[last]);
    return out;
}

// ---- the simplifier ------------------------------------------------------------------------------------------------
struct Simplifier {
    EdgeFn check;
    uint32_t seed = 1;
    int partial_rounds = 0, bspline_steps = 0, reduce_rounds = 0, batches = 0;
    long long edges = 0;
    int rc = PV_OK;

    bool run_batch(const Path& a, const Path& b, std::vector<unsigned char>& ok) {
        const int n = (int)a.size();
        ok.assign(n, 0);
        if (n == 0) return true;
        std::vector<float> fa((size_t)n * 9), fb((size_t)n * 9);
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < 9; ++j) {
                fa[(size_t)k * 9 + j] = (float)a[k][j];
                fb[(size_t)k * 9 + j] = (float)b[k][j];
            }
        rc = check(fa.data(), fb.data(), n, ok.data());
        ++batches;
        edges += n;
        return rc == PV_OK;
    }

    // point at arc length s of the path (L = cumulative lengths); seg receives the segment index
    static Q9 point_at(const Path& p, const std::vector<double>& L, double s, int& seg) {
        const int n = (int)p.size();
        int i = (int)(std::upper_bound(L.begin(), L.end(), s) - L.begin()) - 1;
        i = std::max(0, std::min(i, n - 2));
        seg = i;
        const double d = L[i + 1] - L[i];
        return d > 0 ? q_lerp(p[i], p[i + 1], std::min(1.0, std::max(0.0, (s - L[i]) / d))) : p[i];
    }

    // partialShortcutPath, batched: K random pairs of points on the path (second within +-33 % of the length of the
    // first, both snapped to a vertex when closer than 0.5 % of the length).
    // A candidate replaces the piece of path between its points a (on segment i1) and b (on segment i2) by the straight
    // motion a -> b; what is left of the two segments, p[i1] -> a and b -> p[i2 + 1], are NEW motions as well (a segment
    // that was valid at the validator's samples need not be valid at the samples of a part of it), so a candidate is
    // three motions of the batch -- the first of which also checks the state a -- and is taken only if all three hold.
    bool partial_round(Path& p, int pass) {
        const int n = (int)p.size();
        if (n < 3) return false;
        std::vector<double> L(n, 0.0);
        for (int i = 1; i < n; ++i) L[i] = L[i - 1] + q_dist(p[i - 1], p[i]);
        const double total = L[n - 1];
        if (!(total > 0)) return false;
        const int K = 24;
        struct Cand { double s1, s2, saving; Q9 a, b; int i1, i2; bool a_vertex, b_vertex; };
        std::vector<Cand> cands;
        for (int k = 0; k < K; ++k) {
            double u[4];
            uniform4(seed, (uint32_t)pass, (uint32_t)k, u);
            double s1 = u[0] * total;
            double s2 = s1 + (2.0 * u[1] - 1.0) * 0.33 * total;
            s2 = std::min(total, std::max(0.0, s2));
            if (s1 > s2) std::swap(s1, s2);
            Cand c;
            c.a_vertex = c.b_vertex = false;
            for (int i = 0; i < n; ++i) {  // snap to vertices
                if (fabs(s1 - L[i]) < 0.005 * total) { s1 = L[i]; c.a_vertex = true; c.i1 = i; }
                if (fabs(s2 - L[i]) < 0.005 * total) { s2 = L[i]; c.b_vertex = true; c.i2 = i - 1; }
            }
            if (s2 - s1 < 1e-3 * total) continue;
            int seg;
            c.a = point_at(p, L, s1, seg);
            if (c.a_vertex) c.a = p[c.i1]; else c.i1 = seg;
            c.b = point_at(p, L, s2, seg);
            if (c.b_vertex) c.b = p[c.i2 + 1]; else c.i2 = seg;
            // both on one straight segment (or its end points): nothing to gain
            bool bend = false;
            for (int i = 1; i + 1 < n; ++i) bend |= (L[i] > s1 + 1e-12 && L[i] < s2 - 1e-12);
            if (!bend) continue;
            c.s1 = s1;
            c.s2 = s2;
            c.saving = (s2 - s1) - q_dist(c.a, c.b);
            if (c.saving > 1e-6 * total) cands.push_back(c);
        }
        if (cands.empty()) return false;
        Path A, B;
        for (auto& c : cands) {
            A.push_back(c.a_vertex ? c.a : p[c.i1]);  // what remains of the segment a lies on (checks the state a)
            B.push_back(c.a);
            A.push_back(c.a);
            B.push_back(c.b);
            A.push_back(c.b);                          // what remains of the segment b lies on
            B.push_back(c.b_vertex ? c.b : p[c.i2 + 1]);
        }
        std::vector<unsigned char> ok;
        if (!run_batch(A, B, ok)) return false;
        ++partial_rounds;
        std::vector<int> order;
        for (int k = 0; k < (int)cands.size(); ++k)
            if (ok[3 * k] && ok[3 * k + 1] && ok[3 * k + 2]) order.push_back(k);
        std::sort(order.begin(), order.end(), [&](int x, int y) {
            return cands[x].saving != cands[y].saving ? cands[x].saving > cands[y].saving : x < y;
        });
        // two accepted shortcuts never share a segment (the piece between them would be one more unvalidated motion):
        // the clash test runs on the vertex-to-vertex extents
        auto lo_of = [&](const Cand& c) { return L[c.a_vertex ? c.i1 : c.i1]; };
        auto hi_of = [&](const Cand& c) { return L[c.i2 + 1]; };
        std::vector<int> taken;
        for (int k : order) {
            bool clash = false;
            for (int t : taken) clash |= !(hi_of(cands[k]) <= lo_of(cands[t]) || lo_of(cands[k]) >= hi_of(cands[t]));
            if (!clash) taken.push_back(k);
        }
        if (taken.empty()) return false;
        std::sort(taken.begin(), taken.end(), [&](int x, int y) { return cands[x].s1 < cands[y].s1; });
        Path out;
        auto push = [&](const Q9& q) {
            if (out.empty() || q_dist(out.back(), q) > 0) out.push_back(q);
        };
        int i = 0;
        for (int t : taken) {
            const Cand& c = cands[t];
            for (; i <= c.i1; ++i) push(p[i]);  // up to and including the vertex in front of (or at) a
            push(c.a);
            push(c.b);
            i = c.i2 + 1;                        // the vertex behind (or at) b comes next
        }
        for (; i < n; ++i) push(p[i]);
        if (out.size() < 2) return false;
        out.front() = p.front();
        out.back() = p.back();
        p.swap(out);
        return true;
    }

    // one step of smoothBSpline; returns the number of vertices moved.  subdivide() puts a midpoint m_i on every
    // segment; every original interior vertex v_i then moves to c_i = 1/4 m_(i-1) + 1/2 v_i + 1/4 m_i when both new
    // motions m_(i-1) -> c_i and c_i -> m_i are valid and it moves by more than the minimum change.  The halves
    // v_i -> m_i and m_i -> v_(i+1) are new motions too; one batch carries them all.  A midpoint both of whose halves
    // hold is kept (whichever neighbours move, all its motions have then been validated); a midpoint with a failing
    // half (its segment grazes an obstacle between the validator's samples) pins its two neighbours and is dropped again,
    // which restores the original segment.
    int bspline_step(Path& p, double min_change) {
        const int n = (int)p.size();
        if (n < 3) return 0;
        Path s;
        s.reserve(2 * n - 1);
        for (int i = 0; i + 1 < n; ++i) {  // PathGeometric::subdivide
            s.push_back(p[i]);
            s.push_back(q_lerp(p[i], p[i + 1], 0.5));
        }
        s.push_back(p[n - 1]);
        const int ns = (int)s.size();
        Path A, B, C;
        std::vector<int> idx;
        for (int i = 2; i + 1 < ns; i += 2) {
            const Q9 t1 = q_lerp(s[i - 1], s[i], 0.5), t2 = q_lerp(s[i], s[i + 1], 0.5);
            const Q9 c = q_lerp(t1, t2, 0.5);
            if (!(q_dist(s[i], c) > min_change)) continue;
            idx.push_back(i);
            C.push_back(c);
            A.push_back(s[i - 1]);
            B.push_back(c);
            A.push_back(c);
            B.push_back(s[i + 1]);
        }
        if (idx.empty()) return 0;
        const int n_motion = (int)A.size();
        for (int i = 1; i < ns; i += 2) {  // the two halves of every subdivided segment
            A.push_back(s[i - 1]);
            B.push_back(s[i]);
            A.push_back(s[i]);
            B.push_back(s[i + 1]);
        }
        std::vector<unsigned char> ok;
        if (!run_batch(A, B, ok)) return 0;
        ++bspline_steps;
        auto mid_ok = [&](int i) { return ok[n_motion + (i - 1)] && ok[n_motion + (i - 1) + 1]; };  // i odd
        int moved = 0;
        for (size_t k = 0; k < idx.size(); ++k)
            if (ok[2 * k] && ok[2 * k + 1] && mid_ok(idx[k] - 1) && mid_ok(idx[k] + 1)) {
                s[idx[k]] = C[k];
                ++moved;
            }
        if (!moved) return 0;  // a step that moves nothing leaves the path as it was (no growth for nothing)
        Path out;
        for (int i = 0; i < ns; ++i)
            if ((i & 1) == 0 || mid_ok(i)) out.push_back(s[i]);
        p.swap(out);
        return moved;
    }

    // reduceVertices / collapseCloseVertices, batched: every vertex pair (i, j >= i + 2), capped at 96 candidates drawn
    // from the random stream when there are more
    bool reduce_round(Path& p, int pass) {
        const int n = (int)p.size();
        if (n < 3) return false;
        std::vector<double> L(n, 0.0);
        for (int i = 1; i < n; ++i) L[i] = L[i - 1] + q_dist(p[i - 1], p[i]);
        struct Cand { int i, j; double saving; };
        std::vector<Cand> cands;
        const long long all = (long long)(n - 1) * (n - 2) / 2;
        const int K = 96;
        if (all <= K) {
            for (int i = 0; i + 2 < n; ++i)
                for (int j = i + 2; j < n; ++j) cands.push_back({i, j, (L[j] - L[i]) - q_dist(p[i], p[j])});
        } else {
            const int range = std::max(2, (int)(0.33 * n));
            for (int k = 0; k < K; ++k) {
                double u[4];
                uniform4(seed, (uint32_t)pass, (uint32_t)k, u);
                const int i = std::min(n - 1, (int)(u[0] * n));
                int j = i + (int)((2.0 * u[1] - 1.0) * range);
                j = std::max(0, std::min(n - 1, j));
                const int lo = std::min(i, j), hi = std::max(i, j);
                if (hi - lo < 2) continue;
                cands.push_back({lo, hi, (L[hi] - L[lo]) - q_dist(p[lo], p[hi])});
            }
        }
        cands.erase(std::remove_if(cands.begin(), cands.end(), [&](const Cand& c) { return !(c.saving > 1e-9 * (L[n - 1] + 1e-30)); }),
                    cands.end());
        if (cands.empty()) return false;
        Path A, B;
        for (auto& c : cands) {
            A.push_back(p[c.i]);
            B.push_back(p[c.j]);
        }
        std::vector<unsigned char> ok;
        if (!run_batch(A, B, ok)) return false;
        ++reduce_rounds;
        std::vector<int> order;
        for (int k = 0; k < (int)cands.size(); ++k)
            if (ok[k]) order.push_back(k);
        std::sort(order.begin(), order.end(), [&](int x, int y) {
            return cands[x].saving != cands[y].saving ? cands[x].saving > cands[y].saving : x < y;
        });
        std::vector<char> drop(n, 0);
        std::vector<int> taken;
        for (int k : order) {
            bool clash = false;
            for (int t : taken) clash |= !(cands[k].j <= cands[t].i || cands[k].i >= cands[t].j);
            if (clash) continue;
            taken.push_back(k);
            for (int v = cands[k].i + 1; v < cands[k].j; ++v) drop[v] = 1;
        }
        if (taken.empty()) return false;
        Path out;
        for (int i = 0; i < n; ++i)
            if (!drop[i]) out.push_back(p[i]);
        p.swap(out);
        return true;
    }

    // PathSimplifier::simplifyMax in batches
    void simplify(Path& p, int max_vertices) {
        if (p.size() < 3) return;
        int pass = 0;
        for (int times = 0; times < 5 && rc == PV_OK; ++times)
            if (!partial_round(p, pass++)) break;
        const double min_change = path_len(p) / 100.0;
        for (int s = 0; s < 3 && rc == PV_OK; ++s) {
            if (2 * (int)p.size() - 1 > max_vertices) break;
            if (bspline_step(p, min_change) == 0) break;
        }
        for (int times = 0; times < 5 && rc == PV_OK; ++times)
            if (!reduce_round(p, 100 + pass++)) break;
    }
};

// ---- edge batches on the device, through host-mapped staging (no copy operations: one launch + one sync) ----------
#define PL_CUDA(expr)                                                                                        \
    do {                                                                                                     \
        cudaError_t e_ = (expr);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            snprintf(h->err, sizeof(h->err), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                     __LINE__);                                                                              \
            return PV_ERR_CUDA;                                                                              \
        }                                                                                                    \
    } while (0)

#define PLAN_EDGE_CAP 4096  // edges per staged batch (two slots: the speculative batch and the working one)

struct PlanStage {
    float* a;
    float* b;
    unsigned char* ok;  // one byte per motion (the edge kernel's one-edge-per-warp mode)
};
static int plan_stage(PvHandle* h, int slot, PlanStage* s) {
    const size_t per = (size_t)PLAN_EDGE_CAP * 9 * sizeof(float) * 2 + (size_t)PLAN_EDGE_CAP;
    if (!h->plan_host) {
        PL_CUDA(cudaHostAlloc(&h->plan_host, 2 * per, cudaHostAllocMapped));
        h->plan_host_bytes = 2 * per;
    }
    char* p = (char*)h->plan_host + (size_t)slot * per;
    s->a = (float*)p;
    s->b = s->a + (size_t)PLAN_EDGE_CAP * 9;
    s->ok = (unsigned char*)(s->b + (size_t)PLAN_EDGE_CAP * 9);
    return PV_OK;
}
// queue the validation of n <= PLAN_EDGE_CAP staged motions (results land in stage.ok at stream completion)
static int plan_queue_edges(PvHandle* h, const PlanStage& s, int n, float resolution, cudaStream_t st) {
    return pv_launch_edges(h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, s.a, s.b, n, 0, resolution, nullptr,
                           nullptr, st, s.ok);
}
static int plan_check_edges(PvHandle* h, float resolution, const float* a, const float* b, int n, unsigned char* ok) {
    PlanStage s;
    int rc = plan_stage(h, 1, &s);
    if (rc) return rc;
    cudaStream_t st = h->streams[0];
    for (int done = 0; done < n; done += PLAN_EDGE_CAP) {
        const int m = std::min(PLAN_EDGE_CAP, n - done);
        memcpy(s.a, a + (size_t)done * 9, (size_t)m * 9 * sizeof(float));
        memcpy(s.b, b + (size_t)done * 9, (size_t)m * 9 * sizeof(float));
        rc = plan_queue_edges(h, s, m, resolution, st);
        if (rc) return rc;
        PL_CUDA(cudaStreamSynchronize(st));
        memcpy(ok + done, s.ok, (size_t)m);
    }
    return PV_OK;
}

// waypoints (fp32 rows, as handed to the caller) -> the motions that validate them: (w0 -> w0) checks the first state
// itself (a motion check assumes its start valid), then every consecutive pair
static void waypoint_motions(const std::vector<float>& w, int n, std::vector<float>& a, std::vector<float>& b) {
    a.resize((size_t)n * 9);
    b.resize((size_t)n * 9);
    memcpy(a.data(), w.data(), 9 * sizeof(float));
    memcpy(b.data(), w.data(), 9 * sizeof(float));
    if (n > 1) {
        memcpy(a.data() + 9, w.data(), (size_t)(n - 1) * 9 * sizeof(float));
        memcpy(b.data() + 9, w.data() + 9, (size_t)(n - 1) * 9 * sizeof(float));
    }
}
static void to_rows(const Path& p, std::vector<float>& rows) {
    rows.resize(p.size() * 9);
    for (size_t k = 0; k < p.size(); ++k)
        for (int j = 0; j < 9; ++j) rows[k * 9 + j] = (float)p[k][j];
}

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" {

int pv_obb_from_poses(const double* pos, const double* quat, const double* half, int n, float* obb_out) {
    if (n < 0 || (n > 0 && (!pos || !half || !obb_out))) return PV_ERR_BAD_ARG;
    for (int k = 0; k < n; ++k) {
        double w = 1, x = 0, y = 0, z = 0;
        if (quat) {
            w = quat[4 * k], x = quat[4 * k + 1], y = quat[4 * k + 2], z = quat[4 * k + 3];
            const double nq = sqrt(w * w + x * x + y * y + z * z);
            if (!(nq > 0)) return PV_ERR_BAD_ARG;
            w /= nq, x /= nq, y /= nq, z /= nq;
        }
        float* o = obb_out + 16 * k;
        const double hx = half[3 * k], hy = half[3 * k + 1], hz = half[3 * k + 2];
        o[0] = (float)pos[3 * k], o[1] = (float)pos[3 * k + 1], o[2] = (float)pos[3 * k + 2];
        o[3] = (float)hx, o[4] = (float)hy, o[5] = (float)hz;
        o[6] = (float)(1 - 2 * (y * y + z * z)), o[7] = (float)(2 * (x * y - z * w)), o[8] = (float)(2 * (x * z + y * w));
        o[9] = (float)(2 * (x * y + z * w)), o[10] = (float)(1 - 2 * (x * x + z * z)), o[11] = (float)(2 * (y * z - x * w));
        o[12] = (float)(2 * (x * z - y * w)), o[13] = (float)(2 * (y * z + x * w)), o[14] = (float)(1 - 2 * (x * x + y * y));
        o[15] = (float)sqrt(hx * hx + hy * hy + hz * hz);
    }
    return PV_OK;
}

int pv_interpolate_path(const double* states, int n_states, int count, double* out, int capacity, int* n_out) {
    if (!states || n_states < 0 || !out || !n_out) return PV_ERR_BAD_ARG;
    Path p(n_states);
    for (int k = 0; k < n_states; ++k) memcpy(p[k].data(), states + (size_t)k * 9, 9 * sizeof(double));
    const Path r = interpolate_path(p, count);
    *n_out = (int)r.size();
    if ((int)r.size() > capacity) return PV_ERR_CAPACITY;
    for (size_t k = 0; k < r.size(); ++k) memcpy(out + k * 9, r[k].data(), 9 * sizeof(double));
    return PV_OK;
}

int pv_simplify_path_cb(const double* states, int n_states, uint32_t seed, pv_edge_callback cb, void* user, double* out,
                        int capacity, int* n_out, int* counters) {
    if (!states || n_states < 0 || !cb || !out || !n_out) return PV_ERR_BAD_ARG;
    Path p(n_states);
    for (int k = 0; k < n_states; ++k) memcpy(p[k].data(), states + (size_t)k * 9, 9 * sizeof(double));
    Simplifier S;
    S.seed = seed;
    S.check = [&](const float* a, const float* b, int n, unsigned char* ok) { return cb(user, a, b, n, ok); };
    S.simplify(p, capacity);
    if (S.rc) return S.rc;
    *n_out = (int)p.size();
    if (counters) {
        counters[0] = S.partial_rounds;
        counters[1] = S.bspline_steps;
        counters[2] = S.reduce_rounds;
        counters[3] = (int)S.edges;
    }
    if ((int)p.size() > capacity) return PV_ERR_CAPACITY;
    for (size_t k = 0; k < p.size(); ++k) memcpy(out + k * 9, p[k].data(), 9 * sizeof(double));
    return PV_OK;
}

int pv_simplify_path(PvHandle* h, const double* states, int n_states, uint32_t seed, float resolution, double* out,
                     int capacity, int* n_out, int* counters) {
    if (!h || h->magic != PV_HANDLE_MAGIC) return PV_ERR_BAD_HANDLE;
    if (!h->has_scene) {
        snprintf(h->err, sizeof(h->err), "no scene set (pv_set_scene)");
        return PV_ERR_NO_SCENE;
    }
    PvDeviceGuard guard(h->device);
    const float res = resolution > 0.f ? resolution : PV_VALIDITY_RESOLUTION;
    struct Ctx { PvHandle* h; float res; } ctx = {h, res};
    return pv_simplify_path_cb(
        states, n_states, seed,
        [](void* u, const float* a, const float* b, int n, unsigned char* ok) {
            Ctx* c = (Ctx*)u;
            return plan_check_edges(c->h, c->res, a, b, n, ok);
        },
        &ctx, out, capacity, n_out, counters);
}

int pv_plan_path(PvHandle* h, const double* start, const double* goal, int num_waypoints, const PvPlanParams* prm,
                 float* h_waypoints, int capacity, int* n_waypoints, PvPlanStats* stats) {
    if (!h || h->magic != PV_HANDLE_MAGIC) return PV_ERR_BAD_HANDLE;
    if (!h->has_scene) {
        snprintf(h->err, sizeof(h->err), "no scene set (pv_set_scene)");
        return PV_ERR_NO_SCENE;
    }
    if (!start || !goal || !prm || !h_waypoints || !n_waypoints || capacity < 2) {
        snprintf(h->err, sizeof(h->err), "pv_plan_path: bad arguments");
        return PV_ERR_BAD_ARG;
    }
    PvDeviceGuard guard(h->device);
    const double t0 = now_ms();
    PvPlanStats S;
    memset(&S, 0, sizeof(S));
    *n_waypoints = 0;
    const long long launches0 = h->launches;
    const float base_res = prm->resolution > 0.f ? prm->resolution : PV_VALIDITY_RESOLUTION;
    const int max_attempts = prm->max_attempts > 0 ? prm->max_attempts : 4;
    const int max_path = 256;

    float sg[18];
    for (int j = 0; j < 9; ++j) {
        sg[j] = (float)start[j];
        sg[9 + j] = (float)goal[j];
    }
    Q9 q_start, q_goal;
    memcpy(q_start.data(), start, sizeof(Q9));
    memcpy(q_goal.data(), goal, sizeof(Q9));

    // speculative: the straight line start -> goal, resampled, is validated in the shadow of the first solve; most
    // plans of the reference's primitives are straight lines, and then the whole plan costs one synchronisation
    std::vector<float> spec_rows;
    int spec_n = 0;
    PlanStage spec;
    int rc = plan_stage(h, 0, &spec);
    if (rc) return rc;
    const bool resample = num_waypoints > 0;
    if (prm->validate && resample && num_waypoints <= PLAN_EDGE_CAP && num_waypoints <= capacity) {
        Path line = {q_start, q_goal};
        to_rows(interpolate_path(line, num_waypoints), spec_rows);
        spec_n = (int)spec_rows.size() / 9;
        std::vector<float> a, b;
        waypoint_motions(spec_rows, spec_n, a, b);
        memcpy(spec.a, a.data(), a.size() * sizeof(float));
        memcpy(spec.b, b.data(), b.size() * sizeof(float));
    }

    std::vector<float> out_rows;
    bool delivered = false;
    // Fast path: ONE launch answers the questions the common plan consists of -- is the straight motion start -> goal
    // valid at the planner's resolution (the first extension of the solve aims at the goal itself, so this IS its first
    // decision), and is every waypoint of its resampling valid (w0 -> w0 covers the start state, the last waypoint is the
    // goal)?  Most plans of the reference's primitives are such lines (approach, descend, lift); they cost one launch and
    // one synchronisation.  Any failure falls through to the full pipeline, which also reports WHY an end point is bad.
    bool spec_ready = false;
    if (spec_n > 0 && spec_n + 1 <= PLAN_EDGE_CAP) {
        memcpy(spec.a + (size_t)spec_n * 9, sg, 9 * sizeof(float));
        memcpy(spec.b + (size_t)spec_n * 9, sg + 9, 9 * sizeof(float));
        const double t_fast = now_ms();
        rc = plan_queue_edges(h, spec, spec_n + 1, base_res, h->streams[0]);
        if (rc) return rc;
        PL_CUDA(cudaStreamSynchronize(h->streams[0]));
        spec_ready = true;
        bool all_ok = true;
        for (int k = 0; k <= spec_n; ++k) all_ok &= spec.ok[k] != 0;
        S.ms_solve += (float)(now_ms() - t_fast);
        if (all_ok) {
            float d2 = 0.f;
            for (int j = 0; j < 9; ++j) d2 = fmaf(sg[9 + j] - sg[j], sg[9 + j] - sg[j], d2);
            const int nd = std::max(1, (int)ceilf(sqrtf(d2) / base_res));
            out_rows = spec_rows;
            delivered = true;
            S.attempts = 1;
            S.iters = 1;
            S.checks = nd;
            S.vertices_raw = S.vertices = 2;
            S.validated = 1;
            S.speculative_hit = 1;
        }
    }
    for (int attempt = 0; attempt < max_attempts && !delivered; ++attempt) {
        if (attempt > 0 && prm->timeout_s > 0 && (now_ms() - t0) * 1e-3 > prm->timeout_s) break;
        S.attempts = attempt + 1;
        // attempts after a failed VALIDATION halve the motion-validation resolution (S.refinements counts them)
        const float res = base_res / (float)(1 << std::min(S.refinements, 4));
        PvRrtcParams rp;
        memset(&rp, 0, sizeof(rp));
        rp.range = prm->range;
        rp.resolution = res;
        rp.max_iters = prm->max_iters;
        rp.max_nodes = prm->max_nodes;
        rp.max_path = max_path;
        rp.seed = prm->seed + 7919u * (uint32_t)attempt;
        rp.replicas = prm->replicas;
        rp.shortcut_passes = prm->smooth ? 2 : 0;
        rp.check_endpoints = 1;
        rp.planner = prm->planner;
        Path raw;
        int iters = 0;
        long long checks = 0;
        bool spec_queued = false;
        // the speculative validation does not depend on the solve: it runs beside it on a second stream
        std::function<void(cudaStream_t)> hook = [&](cudaStream_t) {
            if (attempt == 0 && spec_n > 0)
                spec_queued = spec_ready || plan_queue_edges(h, spec, spec_n, base_res, h->streams[1]) == PV_OK;
        };
        const double t_solve = now_ms();
        rc = pv_rrtc_run(h, sg, sg + 9, 1, &rp,
                         [&](int, int, const int* len, const int* off, const int* it, const long long* ch, const float* rows) {
                             iters = it[0];
                             checks = ch[0];
                             raw.resize(len[0]);
                             for (int k = 0; k < len[0]; ++k)
                                 for (int j = 0; j < 9; ++j) raw[k][j] = (double)rows[(size_t)(off[0] + k) * 9 + j];
                         },
                         &hook);
        if (spec_queued && !spec_ready && attempt == 0) PL_CUDA(cudaStreamSynchronize(h->streams[1]));
        S.ms_solve += (float)(now_ms() - t_solve);
        if (rc) return rc;
        if (iters < 0) {  // start / goal out of bounds or in collision: OMPL finds no valid start / goal -> no solution
            S.endpoint_status = -iters;
            break;
        }
        S.iters += iters;
        S.checks += checks;
        if (raw.size() < 2) continue;  // iteration / node budget exhausted: next attempt, next seed
        raw.front() = q_start;         // exact end points, as OMPL keeps them in fp64
        raw.back() = q_goal;
        S.vertices_raw = (int)raw.size();

        // candidates for delivery, best first: the simplified path, then the solution as the trees found it
        std::vector<Path> cands;
        if (prm->smooth && raw.size() > 2) {
            const double t_simpl = now_ms();
            Simplifier simp;
            simp.seed = rp.seed;
            simp.check = [&](const float* a, const float* b, int n, unsigned char* ok) {
                return plan_check_edges(h, res, a, b, n, ok);
            };
            Path sp = raw;
            simp.simplify(sp, std::max(3, std::min(max_path, resample ? num_waypoints : capacity)));
            S.ms_simplify += (float)(now_ms() - t_simpl);
            if (simp.rc) return simp.rc;
            S.partial_rounds += simp.partial_rounds;
            S.bspline_steps += simp.bspline_steps;
            S.reduce_rounds += simp.reduce_rounds;
            S.simplify_motions += (int)simp.edges;
            cands.push_back(std::move(sp));
        }
        cands.push_back(raw);
        const double t_post = now_ms();
        for (size_t c = 0; c < cands.size() && !delivered; ++c) {
            const Path& path = cands[c];
            std::vector<float> rows;
            to_rows(resample ? interpolate_path(path, num_waypoints) : path, rows);
            const int n = (int)rows.size() / 9;
            if (n > capacity) {
                snprintf(h->err, sizeof(h->err), "pv_plan_path: %d waypoints do not fit the capacity of %d", n, capacity);
                *n_waypoints = n;
                return PV_ERR_CAPACITY;
            }
            bool valid = true;
            if (prm->validate) {
                const bool straight = path.size() == 2 && spec_queued && n == spec_n &&
                                      memcmp(rows.data(), spec_rows.data(), rows.size() * sizeof(float)) == 0;
                if (straight) {  // already validated in the shadow of the solve
                    for (int k = 0; k < n; ++k) valid &= spec.ok[k] != 0;
                    S.speculative_hit = 1;
                } else {
                    std::vector<float> a, b;
                    std::vector<unsigned char> ok(n);
                    waypoint_motions(rows, n, a, b);
                    rc = plan_check_edges(h, base_res, a.data(), b.data(), n, ok.data());
                    if (rc) return rc;
                    for (int k = 0; k < n; ++k) valid &= ok[k] != 0;
                }
            }
            if (valid) {
                out_rows.swap(rows);
                delivered = true;
                S.vertices = (int)path.size();
                S.fallback_unsimplified = (c > 0 && cands.size() > 1) ? 1 : 0;
                S.validated = prm->validate ? 1 : 0;
            }
        }
        S.ms_post += (float)(now_ms() - t_post);
        if (!delivered) ++S.refinements;  // a solution that does not survive the dense check: plan again, finer
    }
    if (delivered) {
        memcpy(h_waypoints, out_rows.data(), out_rows.size() * sizeof(float));
        *n_waypoints = (int)out_rows.size() / 9;
        S.solved = 1;
    }
    S.launches = (int)(h->launches - launches0);
    S.ms_total = (float)(now_ms() - t0);
    if (stats) *stats = S;
    return PV_OK;
}

}  // extern "C"
