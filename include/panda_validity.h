/* panda_validity.h -- C-ABI of libpanda_validity.so (sm_100a CUDA kernels behind plain C entry points).
 *
 * This is the drop-in boundary for the reference's motion-planning hot path.  Each entry point names
 * the reference interface it replaces (file:line under /root/reference/code).  No torch / C++ types in
 * any signature: plain pointers and sizes.  Pointers named d_* are DEVICE pointers (the caller owns the
 * buffers, e.g. torch tensors handed over as data_ptr()); pointers named h_* are HOST pointers.
 * `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Device-pointer calls
 * are asynchronous on `stream`; host-pointer calls return when the result is in the host buffer.
 *
 * Every function returns 0 on success or a negative PV_ERR_* code; pv_last_error() gives the message.
 * One handle per device; a handle is not thread-safe.
 *
 * Layouts
 *   configuration, SoA:  d_qA[i] = float4(q1,q2,q3,q4), d_qB[i] = float4(q5,q6,q7,q8), d_q9[i] = q9
 *                        (d_q9 == NULL means q9 = q8: symmetric gripper)
 *   configuration, AoS:  h_q[i*9 + j], j = 0..8 -- the qpos vectors the reference passes around
 *                        (planning.py:131-135, shape (n_qs,) = (9,))
 *   verdict bits:        bit (i & 31) of word (i >> 5); 1 = valid.  ceil(n/32) words, tail bits 0.
 *   scene box (OBB):     16 floats: centre xyz, half extents xyz, world-from-box rotation (row-major 9),
 *                        1 pad float (overwritten with the bounding radius).
 */
#ifndef PANDA_VALIDITY_H
#define PANDA_VALIDITY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PV_MAX_OBB 32

#define PV_OK 0
#define PV_ERR_BAD_HANDLE (-1)
#define PV_ERR_BAD_ARG (-2)
#define PV_ERR_CUDA (-3)
#define PV_ERR_NO_DEVICE (-4)
#define PV_ERR_NO_SCENE (-5)
#define PV_ERR_CAPACITY (-6) /* a caller-supplied output buffer is too small (the needed size is reported) */

/* validity rule switches */
#define PV_FLAG_SELF 1u   /* robot-vs-robot pairs on (Genesis enable_self_collision) */
#define PV_FLAG_LIMITS 2u /* accepted, changes nothing: lower <= q <= upper (planning.py:139-150, 165-173) is ALWAYS
                             required.  The joint limits are part of the model's validity domain -- the self-collision
                             pair lists are pruned by a certificate that holds inside them only -- and OMPL never hands
                             the reference's callback a state outside its RealVectorBounds.  A state outside the limits
                             (or with a non-finite joint value) is invalid, culprit kind 4. */

typedef struct PvHandle PvHandle;

/* Library/handle life cycle.  Replaces PlannerInterface.__init__ (planning.py:24-30): binds the frozen
 * Panda model to one CUDA device.  Fails with PV_ERR_NO_DEVICE when no sm_100 device is visible --
 * there is no CPU fallback.  (Developer switch, read here: the environment variable PV_EDGE_CERT2=0 creates a handle
 * whose motion validator leaves out the second-tier certificate pass -- same verdict words, for A/B timing.) */
int pv_create(int device, PvHandle **out);
void pv_destroy(PvHandle *h);
const char *pv_last_error(const PvHandle *h); /* h may be NULL: last pv_create error */
const char *pv_version(void);

/* Model constants the host side needs (frozen with the kernels). */
int pv_model_info(int *n_spheres, int *n_boxes, int *n_ss_pairs, int *n_sb_pairs);
int pv_joint_limits(float lower[9], float upper[9]); /* robot.q_limit, planning.py:139-140 */

/* Scene snapshot: what robot.detect_collision() sees through Genesis (planning.py:211): the block
 * entities (scenes.py:52-83), the ground plane (scenes.py:49) and the robot base pose (scenes.py:29-34).
 * h_obb is [n_obb][16] on the host. */
int pv_set_scene(PvHandle *h, const float *h_obb, int n_obb, float table_z, const float base_xyz[3]);

/* self.attached_object = attached_object (planning.py:153): scene-box index whose contacts with hand /
 * left_finger / right_finger are forgiven (planning.py:221-230); -1 = none. */
int pv_set_attached(PvHandle *h, int obb_index);

/* Physically-correct alternative to pv_set_attached (NOT what planning.py does: there the
 * grasped block stays a static obstacle where it was, planning.py:216-230, although the simulation moves it with the
 * gripper, motion_primitives.py:367-376).  Scene box obb_index is held rigidly by the hand: hand_from_box is its pose
 * in the hand frame, 9 floats of row-major rotation then 3 of translation.  The box is no longer an obstacle at its
 * snapshot pose; instead it rides on the hand and is checked against the plane, every other scene box and the arm
 * links link0..link6 (culprit / contact link id 11).  In those tests its half extents are shrunk by
 * contact_allowance (metres, >= 0): a block that rests on the table or on another block touches it, which is not a
 * collision.  -1 (pose may be NULL) or pv_set_attached() leave the mode.  A new pv_set_scene keeps the mode while
 * obb_index still exists (the pose in the hand frame does not depend on the snapshot). */
int pv_set_carried(PvHandle *h, int obb_index, const float *hand_from_box, float contact_allowance);
int pv_set_flags(PvHandle *h, unsigned flags);
/* Which state / sweep kernel variant answers (all return bit-identical verdict words; a test and tuning knob, no
 * reference counterpart): 0 brute force over every kept pair, 1 per-lane bounding-volume culling, 2 (default) = 1 with
 * each block's share of the batch visited in sorted order. */
int pv_set_culling(PvHandle *h, int mode);
/* Back-to-back pv_check_states launches on one stream (a planner validating batch after batch; no reference counterpart):
 * with overlap on (the default) a launch may start reading its configurations while the PREVIOUS pv_check_states launch on
 * the stream is still finishing (CUDA programmatic dependent launch), which hides the launch gap and the uneven tail of
 * the previous launch.  Stream order is kept for everything a caller can observe: the launch never starts before any other
 * kind of earlier work on the stream has completed, and it writes its verdict words only after ALL earlier work has. */
int pv_set_launch_overlap(PvHandle *h, int on);

/* Fused verdict gather for multi-GPU runs: after this call pv_check_states, pv_sweep and pv_check_edges
 * (batches above 16 384 motions, which are written as whole verdict words) also store
 * every verdict word w of this rank at word (word_offset + w) of EVERY rank's gather buffer, from inside the
 * kernel, over peer memory: d_peer_ptrs is a device array of n_peers buffer base pointers (symmetric memory, one per
 * rank, this rank included); d_multicast, if not NULL, is the NVSwitch multicast address of the same buffer and is
 * used instead (one store, replicated by the switch).  The caller synchronises the ranks (e.g. the symmetric-memory
 * barrier) before reading.  Words at or beyond word_capacity (this rank's slot size) are not forwarded.  n_peers = 0
 * switches the gather off.  d_bits of pv_check_states / pv_sweep may then be NULL. */
int pv_set_gather(PvHandle *h, const void *d_peer_ptrs, int n_peers, void *d_multicast, long long word_offset,
                  long long word_capacity);

/* robot.set_qpos(q) -> link poses (planning.py:210; Genesis FK).  d_pose_out is [n][11][12]:
 * per link position xyz then rotation row-major. */
int pv_fk(PvHandle *h, const float *d_qA, const float *d_qB, const float *d_q9, int64_t n,
          float *d_pose_out, void *stream);
/* The same link poses as pv_check_states / pv_sweep evaluate them internally (joint sines and cosines from the
 * hardware approximations): for measuring how far the verdict kernels' kinematics are from pv_fk.  Same layout. */
int pv_fk_verdict_path(PvHandle *h, const float *d_qA, const float *d_qB, const float *d_q9, int64_t n,
                       float *d_pose_out, void *stream);

/* _is_ompl_state_valid(state) for n states at once (planning.py:209-219). */
int pv_check_states(PvHandle *h, const float *d_qA, const float *d_qB, const float *d_q9, int64_t n,
                    uint32_t *d_bits, void *stream);

/* Same rule, reporting the signed clearance in metres (valid iff >= 0) and, optionally, a culprit code
 * (diagnose_valid_violation, planning.py:43-57).  culprit = kind << 16 | a << 8 | b with kind 1 = table
 * (a = link), 2 = scene box (a = link, b = box), 3 = self (a, b = links), 4 = joint limit (a = joint);
 * 0 = none.  d_culprit may be NULL. */
int pv_state_margins(PvHandle *h, const float *d_qA, const float *d_qB, const float *d_q9, int64_t n,
                     float *d_margin, int32_t *d_culprit, void *stream);

/* robot.detect_collision() for n states (planning.py:47, 211): every colliding pair, not only the deepest one.
 * d_codes is [n][32] culprit codes as above, one per distinct (link, other) pair; d_count[n] the number of
 * distinct pairs found (only the first 32 are stored). */
int pv_state_contacts(PvHandle *h, const float *d_qA, const float *d_qB, const float *d_q9, int64_t n,
                      int32_t *d_codes, int32_t *d_count, void *stream);

/* si.checkMotion(a, b) for n edges at once (OMPL DiscreteMotionValidator installed by SimpleSetup,
 * planning.py:151-156).  States q(t) = a + t (b - a), t = k/nd, k = 1..nd (a is assumed valid).
 * n_steps > 0: nd = n_steps for every edge.  n_steps == 0: nd = ceil(|b - a|_2 / resolution). */
int pv_check_edges(PvHandle *h, const float *d_aA, const float *d_aB, const float *d_a9,
                   const float *d_bA, const float *d_bB, const float *d_b9, int64_t n_edges, int n_steps,
                   float resolution, uint32_t *d_bits, void *stream);
int pv_edge_margins(PvHandle *h, const float *d_aA, const float *d_aB, const float *d_a9,
                    const float *d_bA, const float *d_bB, const float *d_b9, int64_t n_edges, int n_steps,
                    float resolution, float *d_margin, void *stream);

/* Host-buffer forms of the two checks (the call a reference-side binding makes): AoS qpos rows in,
 * verdict bits out, host<->device copies pipelined inside.  h_q* may be pinned or pageable. */
int pv_check_states_host(PvHandle *h, const float *h_q, int64_t n, uint32_t *h_bits);
/* The same verdicts for n ARM configurations that share one gripper opening: rows of the 7 revolute joint values, the two
 * finger joints (q8, q9 of planning.py's 9-vector) given once.  That is the shape of the reference's own batches -- the
 * motion primitives never plan finger motion (motion_primitives.py:169-170 overwrite the finger entries of every waypoint)
 * and BASELINE config 2 fixes q8 = q9 = 0.04 -- and the call is PCIe-bound, so 28 instead of 36 bytes per configuration
 * cross the link.  Bit-identical to pv_check_states_host on the rows [q7, finger_left, finger_right]. */
int pv_check_states_host_arm(PvHandle *h, const float *h_q7, int64_t n, float finger_left, float finger_right,
                             uint32_t *h_bits);
int pv_check_edges_host(PvHandle *h, const float *h_qa, const float *h_qb, int64_t n_edges, int n_steps,
                        float resolution, uint32_t *h_bits);

/* Device-generated sweep (BASELINE config 5): configs first .. first+n-1 of the counter-based stream
 * `seed` (Philox-4x32-10; q_j = fma(u_j, upper_j - lower_j, lower_j)), checked in place.
 * d_bits gets ceil(n/32) words; *d_n_valid (device, may be NULL) is incremented by the valid count.
 * `first` must be a multiple of 32.  d_q_out (may be NULL) receives the configs as AoS [n][9]. */
int pv_sweep(PvHandle *h, uint64_t first, int64_t n, uint32_t seed, int fingers_open, uint32_t *d_bits,
             unsigned long long *d_n_valid, float *d_q_out, void *stream);

/* Batched multi-query RRT-Connect (og.RRTConnect + ss.solve, planning.py:156,190), one warp per query,
 * trees and paths in device memory.  See PvRrtcParams. */
typedef struct {
    float range;         /* RRTConnect range; <= 0 -> 0.2 * extent (OMPL default) */
    float resolution;    /* motion-validity resolution; <= 0 -> 0.01 * extent */
    int max_iters;       /* iteration cap per query */
    int max_nodes;       /* per tree */
    int max_path;        /* capacity of one output path (states) */
    uint32_t seed;
    int replicas;        /* independent searches per (start, goal); the first to connect wins */
    int shortcut_passes; /* deterministic shortcutting passes after the solve (simplifySolution) */
    int check_endpoints; /* != 0: test start and goal first (bounds + validity, planning.py:163-183); a query whose
                            start / goal / both fail returns path length 0 and iters = -1 / -2 / -3 */
    int planner;         /* 0 = RRTConnect (the reference's default, planning.py:67), 1 = RRT (single tree, 5 % goal bias) */
    int query_offset;    /* id of h_starts[0] inside the caller's whole batch.  The random stream of a search is keyed by
                            (seed, global query id, replica), so a batch split over several calls or GPUs returns exactly
                            what the unsplit call returns */
} PvRrtcParams;

/* h_starts/h_goals: [n_queries][9] host AoS.  h_path_out: [n_queries][max_path][9], of which only the first
 * h_path_len[k] rows of path k are written; h_path_len: states per path (0 = no solution); h_iters: iterations
 * used; h_checks: state checks issued (may be NULL).
 * Parameters: 0 selects the default of max_iters (2000), max_nodes (2048), max_path (128) and replicas (1); values
 * outside max_iters 1..2^23-1, max_nodes 8..2^22, max_path 2..2^20, replicas 1..256 are PV_ERR_BAD_ARG (never silently
 * replaced: the caller sized h_path_out from ITS max_path).  With replicas > 1 the winner is the search with the
 * smallest (iterations, replica id), so the result is the same on every run.  Memory on the device is bounded
 * whatever n_queries is (the batch is planned in chunks; searches start with small trees and only those that outgrow
 * them are re-planned with max_nodes), results are identical to one full-size run. */
int pv_rrtc_batch(PvHandle *h, const float *h_starts, const float *h_goals, int n_queries,
                  const PvRrtcParams *params, float *h_path_out, int *h_path_len, int *h_iters,
                  long long *h_checks);

/* The same planner with PACKED paths: the states of all paths back to back in h_states ([state_capacity][9]); path k is
 * rows h_path_off[k] .. h_path_off[k] + h_path_len[k] - 1.  *n_states (may be NULL) receives the total number of rows.
 * The dense form above needs max_path x 36 B per query on the caller's side (4.6 KB for typically 2..4 states); this
 * one scales to 10^6 queries.  If the rows do not fit, lengths / offsets / *n_states are still complete and
 * PV_ERR_CAPACITY is returned. */
int pv_rrtc_batch_packed(PvHandle *h, const float *h_starts, const float *h_goals, int n_queries,
                         const PvRrtcParams *params, float *h_states, long long state_capacity,
                         long long *h_path_off, int *h_path_len, int *h_iters, long long *h_checks,
                         long long *n_states);

/* PlannerInterface.plan_path in one call (planning.py:59-207): intake check of start and goal (planning.py:163-183),
 * ss.solve (planning.py:190), ss.simplifySolution() when `smooth` (planning.py:195-196: OMPL simplifyMax = partial
 * shortcuts, B-spline smoothing, vertex reduction -- each pass validated as one batch by the edge kernel),
 * path.interpolate(num_waypoints) (planning.py:198), and a validation of everything handed back: the first waypoint and
 * every motion between consecutive waypoints, which is denser than the planner's own 1 % resolution.  A path that fails
 * it is replaced by the unsimplified solution; if that fails too the query is planned again with a new seed at half the
 * motion-validation resolution, up to max_attempts times / until timeout_s.
 * start, goal: 9 doubles (the reference plans in fp64 and keeps the end points exact; the waypoints come back as fp32
 * rows like the tensors of planning.py:232-242).  num_waypoints <= 0: the path vertices are returned as they are.
 * h_waypoints: [capacity][9]; *n_waypoints = rows written (0 = no solution; planning.py:201-202 returns []).  A path
 * with more vertices than num_waypoints is returned unchanged (PathGeometric::interpolate), so capacity should be
 * >= max(num_waypoints, 256); PV_ERR_CAPACITY (with the needed count in *n_waypoints) otherwise. */
typedef struct {
    float range;       /* <= 0 -> OMPL default (0.2 x extent) */
    float resolution;  /* <= 0 -> OMPL default (0.01 x extent) */
    int max_iters;     /* per attempt; 0 -> 2000 */
    int max_nodes;     /* per tree; 0 -> 2048 */
    uint32_t seed;
    int replicas;      /* OR-parallel searches per attempt; 0 -> 1.  Deterministic winner (see pv_rrtc_batch) */
    int smooth;        /* smooth_path (planning.py:62, 195) */
    int planner;       /* 0 RRTConnect, 1 RRT */
    int validate;      /* != 0: dense validation of the returned waypoints, with fallback and re-planning */
    int max_attempts;  /* 0 -> 4 */
    double timeout_s;  /* ss.solve(timeout) (planning.py:190); checked between attempts */
} PvPlanParams;

typedef struct {
    int solved;
    int endpoint_status;      /* bit 0: start invalid / out of bounds, bit 1: goal (planning.py:165-183) */
    int attempts;             /* solves issued */
    int refinements;          /* attempts that halved the resolution after a failed dense validation */
    int iters;                /* planner iterations over all attempts */
    long long checks;         /* state checks of the planner over all attempts */
    int vertices_raw;         /* vertices of the solution as the trees found it (after the in-kernel vertex shortcuts) */
    int vertices;             /* vertices of the path that was resampled and returned */
    int partial_rounds, bspline_steps, reduce_rounds; /* batches each simplifier pass ran */
    int simplify_motions;     /* candidate motions the simplifier had validated */
    int fallback_unsimplified; /* 1: the simplified path failed the dense validation, the raw solution was returned */
    int validated;            /* 1: the returned waypoints passed the dense validation */
    int speculative_hit;      /* 1: the straight-line answer had been validated in the shadow of the solve */
    int launches;             /* kernel launches of this call */
    float ms_solve, ms_simplify, ms_post, ms_total; /* host wall clock inside the call */
} PvPlanStats;

int pv_plan_path(PvHandle *h, const double *start, const double *goal, int num_waypoints, const PvPlanParams *params,
                 float *h_waypoints, int capacity, int *n_waypoints, PvPlanStats *stats);

/* Scene records for pv_set_scene from what a simulator reports per box entity (entity.get_pos(), entity.get_quat() wxyz,
 * half of entity.morph.size; scenes.py:52-83): host arithmetic only.  pos [n][3], quat [n][4] (NULL = identity),
 * half [n][3] fp64 -> obb_out [n][16] fp32. */
int pv_obb_from_poses(const double *pos, const double *quat, const double *half, int n, float *obb_out);

/* path.interpolate(count) (planning.py:198; OMPL PathGeometric::interpolate for a RealVectorStateSpace) on its own:
 * host arithmetic only (no device, no handle).  states [n_states][9] fp64 -> out [capacity][9]; *n_out = rows. */
int pv_interpolate_path(const double *states, int n_states, int count, double *out, int capacity, int *n_out);

/* ss.simplifySolution() (planning.py:196) on a caller-supplied vertex list: the simplifier of pv_plan_path on its own,
 * every candidate motion validated by the edge kernel against the current scene.  states [n_states][9] fp64 ->
 * out [capacity][9]; resolution <= 0 -> OMPL default.  counters as for pv_simplify_path_cb. */
int pv_simplify_path(PvHandle *h, const double *states, int n_states, uint32_t seed, float resolution, double *out,
                     int capacity, int *n_out, int *counters);

/* The simplifier of pv_plan_path with the motion validator supplied by the caller -- a hook for testing the pass
 * logic against another validator (tests/ hands it the CPU oracle); the product path never uses it.
 * cb(user, a, b, n, ok): validate n motions a[k] -> b[k] (fp32 rows of 9), ok[k] = 1 when valid; returns 0.
 * counters (may be NULL): {partial rounds, B-spline steps, reduce rounds, motions validated}. */
typedef int (*pv_edge_callback)(void *user, const float *a, const float *b, int n, unsigned char *ok);
int pv_simplify_path_cb(const double *states, int n_states, uint32_t seed, pv_edge_callback cb, void *user,
                        double *out, int capacity, int *n_out, int *counters);

/* Device steps of the SHARDED-TREE planner front end (distributed.ShardedTreePlanner): for trees too large
 * for one GPU -- or simply to use the node memory and nearest-neighbour bandwidth of all of them -- every rank keeps
 * every world-th node of every tree (global node g at slot g / world of rank g % world) and the ranks exchange, per
 * extension, their nearest-node candidates and the motion verdicts over NCCL.  All pointers are DEVICE pointers.
 *   pv_nn_candidates  pair t: nearest of this rank's nodes of tree d_tree_of[t] (NULL: tree t) to d_targets[t]; trees are
 *                     [n][9][capacity] SoA, d_sizes[tree] = slots in use; d_out[t] = 11 floats: squared distance, GLOBAL
 *                     node index (int bits; INT_MAX when the rank holds no node of the tree), the node's 9 joint values
 *   pv_rrtc_steer     d_cand [world][n][11] (the all-gathered records): reduce by (distance, global index) and form the
 *                     motion from the nearest node towards the target, at most `range` long (<= 0: OMPL default)
 *   pv_rrtc_samples   sample d_it[i] of global search d_gsearch[i] of the planner's random stream
 * With these three the front end takes the decisions of pv_rrtc_batch bit for bit (og.RRTConnect as configured at
 * planning.py:151-156). */
int pv_nn_candidates(PvHandle *h, const float *d_trees, const int *d_sizes, const int *d_tree_of, const float *d_targets,
                     int n_pairs, int capacity, int rank, int world, float *d_out, void *stream);
/* pv_nn_candidates with the all-gather fused into it: record t of this rank goes, word by word, into EVERY rank's
 * symmetric buffer at [rank][t][11] (buffers of world * n_pairs * 11 32-bit words; d_peer_ptrs = device array of the
 * n_peers buffer base pointers, d_multicast = the NVSwitch multicast address of the same buffer or NULL), so that after the
 * symmetric-memory barrier every rank holds what the all-gather would have delivered -- pv_rrtc_steer's input. */
int pv_nn_candidates_gather(PvHandle *h, const float *d_trees, const int *d_sizes, const int *d_tree_of,
                            const float *d_targets, int n_pairs, int capacity, int rank, int world, const void *d_peer_ptrs,
                            int n_peers, void *d_multicast, void *stream);
int pv_rrtc_steer(PvHandle *h, const float *d_cand, int world, int n, const float *d_targets, float range,
                  int *d_from_gidx, float *d_ea, float *d_eb, int *d_reach, void *stream);
int pv_rrtc_samples(PvHandle *h, uint32_t seed, const unsigned *d_gsearch, const int *d_it, int n, float *d_out,
                    void *stream);

/* robot.inverse_kinematics(link=hand, pos, quat) as the motion primitives call it before every plan_path
 * (motion_primitives.py:131-134), batched and collision-aware.  For each of n_targets hand poses (world
 * position, quaternion wxyz) n_seeds damped-least-squares searches run in parallel (seed 0 = h_q_init, the rest
 * uniform in the joint limits from a counter-based RNG); converged candidates are filtered by the state-validity
 * rule of the current scene and the valid one closest to h_q_init is returned.  Finger joints are copied from
 * h_q_init.  h_status[i] = 1 if a solution was found; h_err[i] = {position error m, rotation error rad} (may be
 * NULL).  pos_tol / rot_tol <= 0 select 1e-4 m / 1e-3 rad; n_seeds is rounded up to a power of two in 32..256. */
int pv_ik_batch(PvHandle *h, const float *h_pos, const float *h_quat, int n_targets, const float *h_q_init,
                int n_seeds, int max_iters, float pos_tol, float rot_tol, uint32_t seed, float *h_q_out,
                int *h_status, float *h_err);

/* FP32 FMA issue-rate micro-benchmark (roofline denominator; MEASURED_PEAKS.json has no FP32 entry).
 * Returns achieved TFLOP/s over `iters` unrolled FFMA rounds; ms (may be NULL) gets the kernel time. */
int pv_fp32_peak(PvHandle *h, int iters, double *tflops, float *ms);

/* Number of kernel launches issued through this handle since creation (bench.py's gpu_launches). */
long long pv_launch_count(const PvHandle *h);

#ifdef __cplusplus
}
#endif
#endif
