/* CPU ORACLE body (test infrastructure, NOT product code); included twice by panda_oracle.c with
 * REAL = double (suffix _f64) and REAL = float (suffix _f32).
 *
 * PARITY UNPINNED vs the reference's Genesis/OMPL verdicts (no reference tests / golden vectors; Genesis,
 * OMPL and the Panda meshes are not installable here -- SURVEY.md 4, 8c).  Pinned: FK against the
 * analytic known-answer values of SURVEY.md App. A; everything else against oracle/panda_oracle.py.
 *
 * Restates: planning.py:209-219 (verdict rule), planning.py:221-230 (attached-object forgiveness),
 * planning.py:139-150 (joint limits), planning.py:151-156 -> OMPL DiscreteMotionValidator (App. D),
 * scenes.py:85 -> Menagerie panda.xml chain (App. A), scenes.py:29-34 (base lift).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

typedef struct {
    int n_spheres;
    const int *sphere_link;
    const REAL *sphere_center; /* [S][3] */
    const REAL *sphere_radius; /* [S] */
    int n_boxes;
    const int *box_link;
    const REAL *box_center; /* [H][3] */
    const REAL *box_half;   /* [H][3] */
    int n_ss;
    const int *ss_pairs; /* [P][2] */
    int n_sb;
    const int *sb_pairs; /* [P2][2] sphere, box */
    const REAL *q_lower, *q_upper; /* [9] */
} FN(po_model);

static void FN(mat_mul)(const REAL *A, const REAL *B, REAL *C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

static void FN(quat_mat)(const double *q, REAL *R) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    R[0] = (REAL)(1 - 2 * (y * y + z * z)); R[1] = (REAL)(2 * (x * y - z * w)); R[2] = (REAL)(2 * (x * z + y * w));
    R[3] = (REAL)(2 * (x * y + z * w)); R[4] = (REAL)(1 - 2 * (x * x + z * z)); R[5] = (REAL)(2 * (y * z - x * w));
    R[6] = (REAL)(2 * (x * z - y * w)); R[7] = (REAL)(2 * (y * z + x * w)); R[8] = (REAL)(1 - 2 * (x * x + y * y));
}

/* FK of the 11 bodies: child = parent * Trans(pos) * Rot(quat) * [Rot_z(q) | Trans_y(q)]. */
static void FN(fk_one)(const REAL *q, const REAL *base, REAL *R /*[11][9]*/, REAL *p /*[11][3]*/) {
    for (int i = 0; i < 11; ++i) {
        REAL Rl[9], Rw[9], pw[3];
        FN(quat_mat)(PO_QUAT[i], Rl);
        if (PO_PARENT[i] < 0) {
            memcpy(Rw, Rl, sizeof(Rw));
            for (int k = 0; k < 3; ++k) pw[k] = base[k] + (REAL)PO_POS[i][k];
        } else {
            const REAL *Rp = R + 9 * PO_PARENT[i], *pp = p + 3 * PO_PARENT[i];
            FN(mat_mul)(Rp, Rl, Rw);
            for (int k = 0; k < 3; ++k)
                pw[k] = pp[k] + Rp[3 * k] * (REAL)PO_POS[i][0] + Rp[3 * k + 1] * (REAL)PO_POS[i][1] +
                        Rp[3 * k + 2] * (REAL)PO_POS[i][2];
        }
        if (PO_JTYPE[i] == 1) {
            REAL c = (REAL)cos((double)q[PO_JIDX[i]]), s = (REAL)sin((double)q[PO_JIDX[i]]);
            REAL Rz[9] = {c, -s, 0, s, c, 0, 0, 0, 1}, T[9];
            FN(mat_mul)(Rw, Rz, T);
            memcpy(Rw, T, sizeof(T));
        } else if (PO_JTYPE[i] == 2) {
            for (int k = 0; k < 3; ++k) pw[k] += Rw[3 * k + 1] * q[PO_JIDX[i]];
        }
        memcpy(R + 9 * i, Rw, sizeof(Rw));
        memcpy(p + 3 * i, pw, sizeof(pw));
    }
}

/* signed clearance of a sphere (centre c, radius r) from a box (centre bc, half bh, world-from-box bR) */
static REAL FN(sphere_obb)(const REAL *c, REAL r, const REAL *bc, const REAL *bh, const REAL *bR) {
    REAL d[3] = {c[0] - bc[0], c[1] - bc[1], c[2] - bc[2]};
    REAL s2 = 0, inside = -(REAL)1e30;
    for (int i = 0; i < 3; ++i) {
        REAL loc = bR[i] * d[0] + bR[3 + i] * d[1] + bR[6 + i] * d[2];
        REAL e = (REAL)fabs((double)loc) - bh[i];
        if (e > inside) inside = e;
        if (e > 0) s2 += e * e;
    }
    return (s2 > 0 ? (REAL)sqrt((double)s2) : inside) - r;
}

/* SAT margin of two boxes: largest normalised separation over the 15 axes; cross axes whose squared
 * length is below 1e-4 (near-parallel edges) are skipped. */
static REAL FN(obb_obb)(const REAL *ca, const REAL *ha, const REAL *Ra, const REAL *cb, const REAL *hb,
                        const REAL *Rb) {
    REAL Rm[9], A[9], t[3], d[3] = {cb[0] - ca[0], cb[1] - ca[1], cb[2] - ca[2]};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            Rm[3 * i + j] = Ra[i] * Rb[j] + Ra[3 + i] * Rb[3 + j] + Ra[6 + i] * Rb[6 + j];
            A[3 * i + j] = (REAL)fabs((double)Rm[3 * i + j]);
        }
        t[i] = Ra[i] * d[0] + Ra[3 + i] * d[1] + Ra[6 + i] * d[2];
    }
    REAL best = -(REAL)1e30;
    for (int i = 0; i < 3; ++i) {
        REAL rb = A[3 * i] * hb[0] + A[3 * i + 1] * hb[1] + A[3 * i + 2] * hb[2];
        REAL s = (REAL)fabs((double)t[i]) - ha[i] - rb;
        if (s > best) best = s;
    }
    for (int j = 0; j < 3; ++j) {
        REAL ra = A[j] * ha[0] + A[3 + j] * ha[1] + A[6 + j] * ha[2];
        REAL tl = t[0] * Rm[j] + t[1] * Rm[3 + j] + t[2] * Rm[6 + j];
        REAL s = (REAL)fabs((double)tl) - ra - hb[j];
        if (s > best) best = s;
    }
    for (int i = 0; i < 3; ++i) {
        int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        for (int j = 0; j < 3; ++j) {
            int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            REAL len2 = (REAL)1 - Rm[3 * i + j] * Rm[3 * i + j];
            if (!(len2 > (REAL)1e-4)) continue;
            REAL ra = ha[i1] * A[3 * i2 + j] + ha[i2] * A[3 * i1 + j];
            REAL rb = hb[j1] * A[3 * i + j2] + hb[j2] * A[3 * i + j1];
            REAL tl = (REAL)fabs((double)(t[i2] * Rm[3 * i1 + j] - t[i1] * Rm[3 * i2 + j]));
            REAL s = (tl - ra - rb) / (REAL)sqrt((double)len2);
            if (s > best) best = s;
        }
    }
    return best;
}

#define PO_FLAG_SELF 1
#define PO_FLAG_LIMITS 2
/* carry mode (not reference behaviour): scene box `attached` is held rigidly by the hand.  Its pose
 * in the hand frame is the extra record obb[n_obb] = [t.xyz, half.xyz, R row-major (hand-from-box), pad]. */
#define PO_FLAG_CARRY 4
#define PO_HAND_LINK 8
#define PO_CARRY_LAST_ARM_LINK 6

/* Minimum signed clearance of one configuration (valid iff >= 0).  An out-of-bounds configuration
 * returns -1e30. */
static REAL FN(state_margin_one)(const FN(po_model) * m, const REAL *obb, int n_obb, REAL table_z,
                                 const REAL *base, int attached, int flags, const REAL *q) {
    REAL R[11 * 9], p[11 * 3];
    REAL wc[64 * 3], bw[8 * 3];
    REAL best = (REAL)1e30;
    /* Joint limits are part of the validity domain and always enforced (PO_FLAG_LIMITS is accepted and ignored):
     * OMPL only ever hands the callback states inside the RealVectorBounds of planning.py:139-150, and the pruned
     * self-collision pair lists of the model are certified inside the limits only.  The comparison is made on the
     * fp32 value of the joint against the fp32 limits -- what the kernels, whose inputs are fp32, see -- so a state
     * AT a limit is inside whichever precision it is handed over in; the negated form also rejects non-finite values. */
    for (int j = 0; j < 9; ++j) {
        const float qf = (float)q[j];
        if (!(qf >= (float)m->q_lower[j] && qf <= (float)m->q_upper[j])) return -(REAL)1e30;
    }
    FN(fk_one)(q, base, R, p);
    for (int i = 0; i < m->n_spheres; ++i) {
        const REAL *Rl = R + 9 * m->sphere_link[i], *pl = p + 3 * m->sphere_link[i], *c = m->sphere_center + 3 * i;
        for (int k = 0; k < 3; ++k) wc[3 * i + k] = pl[k] + Rl[3 * k] * c[0] + Rl[3 * k + 1] * c[1] + Rl[3 * k + 2] * c[2];
        if (m->sphere_link[i] != 0) { /* link0 and the plane are both fixed: pair filtered */
            REAL s = wc[3 * i + 2] - m->sphere_radius[i] - table_z;
            if (s < best) best = s;
        }
    }
    for (int k = 0; k < m->n_boxes; ++k) {
        const REAL *Rl = R + 9 * m->box_link[k], *pl = p + 3 * m->box_link[k], *c = m->box_center + 3 * k;
        const REAL *h = m->box_half + 3 * k;
        for (int a = 0; a < 3; ++a) bw[3 * k + a] = pl[a] + Rl[3 * a] * c[0] + Rl[3 * a + 1] * c[1] + Rl[3 * a + 2] * c[2];
        REAL ext = (REAL)fabs((double)Rl[6]) * h[0] + (REAL)fabs((double)Rl[7]) * h[1] + (REAL)fabs((double)Rl[8]) * h[2];
        REAL s = bw[3 * k + 2] - ext - table_z;
        if (s < best) best = s;
    }
    const int carry = (flags & PO_FLAG_CARRY) && attached >= 0 && attached < n_obb;
    REAL cc[3], cR[9];
    const REAL *chalf = obb + 16 * n_obb + 3;
    if (carry) {
        const REAL *rec = obb + 16 * n_obb, *Rh = R + 9 * PO_HAND_LINK, *ph = p + 3 * PO_HAND_LINK;
        FN(mat_mul)(Rh, rec + 6, cR);
        for (int a = 0; a < 3; ++a) cc[a] = ph[a] + Rh[3 * a] * rec[0] + Rh[3 * a + 1] * rec[1] + Rh[3 * a + 2] * rec[2];
        REAL ext = (REAL)fabs((double)cR[6]) * chalf[0] + (REAL)fabs((double)cR[7]) * chalf[1] +
                   (REAL)fabs((double)cR[8]) * chalf[2];
        REAL s = cc[2] - ext - table_z;
        if (s < best) best = s;
        for (int i = 0; i < m->n_spheres; ++i) {
            if (m->sphere_link[i] > PO_CARRY_LAST_ARM_LINK) continue;
            s = FN(sphere_obb)(wc + 3 * i, m->sphere_radius[i], cc, chalf, cR);
            if (s < best) best = s;
        }
    }
    for (int b = 0; b < n_obb; ++b) {
        const REAL *o = obb + 16 * b;
        if (carry && b == attached) continue; /* it moves with the hand */
        if (carry) {
            REAL s = FN(obb_obb)(cc, chalf, cR, o, o + 3, o + 6);
            if (s < best) best = s;
        }
        for (int i = 0; i < m->n_spheres; ++i) {
            REAL s = FN(sphere_obb)(wc + 3 * i, m->sphere_radius[i], o, o + 3, o + 6);
            if (s < best) best = s;
        }
        if (b == attached) continue; /* hand / finger contacts with the attached box are forgiven */
        for (int k = 0; k < m->n_boxes; ++k) {
            REAL s = FN(obb_obb)(bw + 3 * k, m->box_half + 3 * k, R + 9 * m->box_link[k], o, o + 3, o + 6);
            if (s < best) best = s;
        }
    }
    if (flags & PO_FLAG_SELF) {
        for (int k = 0; k < m->n_ss; ++k) {
            int a = m->ss_pairs[2 * k], b = m->ss_pairs[2 * k + 1];
            REAL dx = wc[3 * a] - wc[3 * b], dy = wc[3 * a + 1] - wc[3 * b + 1], dz = wc[3 * a + 2] - wc[3 * b + 2];
            REAL s = (REAL)sqrt((double)(dx * dx + dy * dy + dz * dz)) - (m->sphere_radius[a] + m->sphere_radius[b]);
            if (s < best) best = s;
        }
        for (int k = 0; k < m->n_sb; ++k) {
            int a = m->sb_pairs[2 * k], bx = m->sb_pairs[2 * k + 1];
            REAL s = FN(sphere_obb)(wc + 3 * a, m->sphere_radius[a], bw + 3 * bx, m->box_half + 3 * bx,
                                    R + 9 * m->box_link[bx]);
            if (s < best) best = s;
        }
    }
    return best;
}

void FN(po_fk)(const REAL *q, long n, const REAL *base, REAL *R_out, REAL *p_out) {
    for (long i = 0; i < n; ++i) FN(fk_one)(q + 9 * i, base, R_out + 99 * i, p_out + 33 * i);
}

void FN(po_state_margin)(const FN(po_model) * m, const REAL *obb, int n_obb, REAL table_z, const REAL *base,
                         int attached, int flags, const REAL *q, long n, REAL *out, int nthreads) {
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static, 256)
    for (long i = 0; i < n; ++i)
        out[i] = FN(state_margin_one)(m, obb, n_obb, table_z, base, attached, flags, q + 9 * i);
}

/* Edge: min clearance over q(t) = qa + t (qb - qa), t = k/nd, k = 1..nd.  n_steps > 0 fixes nd; n_steps == 0
 * applies OMPL's DiscreteMotionValidator count nd = ceil(|qb - qa| / resolution) (>= 1).
 * early_exit != 0 stops an edge at its first colliding state, testing b first and then the interior
 * states in ascending order (the verdict -- sign of the result -- is unchanged; the value is then the first
 * negative margin met, not the minimum). */
void FN(po_edge_margin)(const FN(po_model) * m, const REAL *obb, int n_obb, REAL table_z, const REAL *base,
                        int attached, int flags, const REAL *qa, const REAL *qb, long n, int n_steps,
                        REAL resolution, int early_exit, REAL *out, long *n_checked, int nthreads) {
    long total = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static, 64) reduction(+ : total)
    for (long e = 0; e < n; ++e) {
        const REAL *a = qa + 9 * e, *b = qb + 9 * e;
        int nd = n_steps;
        if (nd <= 0) {
            double d2 = 0;
            for (int j = 0; j < 9; ++j) d2 += ((double)b[j] - a[j]) * ((double)b[j] - a[j]);
            nd = (int)ceil(sqrt(d2) / (double)resolution);
            if (nd < 1) nd = 1;
        }
        REAL best = (REAL)1e30;
        for (int kk = 0; kk < nd; ++kk) {
            int k = (kk == 0) ? nd : kk; /* endpoint first */
            REAL t = (REAL)k / (REAL)nd, qs[9];
            for (int j = 0; j < 9; ++j) qs[j] = a[j] + t * (b[j] - a[j]);
            REAL s = FN(state_margin_one)(m, obb, n_obb, table_z, base, attached, flags, qs);
            ++total;
            if (s < best) best = s;
            if (early_exit && s < 0) break;
        }
        out[e] = best;
    }
    if (n_checked) *n_checked = total;
}

#undef FN
#undef CAT
#undef CAT_
