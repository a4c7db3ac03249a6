/*
 * oracle/sim_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by pienerf_b200/).
 *
 * fp64 CPU restatement of the reference's Q-GMLS simulator for the hot path
 * "Simulator.stepforward + get_IP_info" and the init that feeds it.
 *
 * PARITY PINNED by the reference's own source: tests/golden/ref_sim_block{64,512}.npz are written by
 * tests/golden/make_golden_sim.py, which imports the UNMODIFIED /root/reference/simulator/{func_utils,cpu_utils,
 * cuda_utils,solver}.py on numpy stand-ins for warp / kornia / plyfile (tests/golden/warp_shim.py; none of the three
 * is installable here) and records init + a 10-step sequence with a drag force.  tests/test_sim_golden.py holds this
 * file to those fixtures: topology bit-equal, shape functions / matrices / rhs <= 1e-9, DOFs <= 1e-9 (measured 2e-11),
 * velocities <= 1e-7 relative.  The only arithmetic the fixtures do not contain is wp.svd3's own iteration (the shim
 * uses an exact SVD in Warp's convention).  The GMLS identities in tests/test_sim_oracle.py stay as a second net.
 *
 * Third-party arithmetic restated here (not under /root/reference):
 *   wp.svd3 (warp-lang 0.13.0, call site simulator/cuda_utils.py:107) ->
 *     exact Jacobi SVD, U,V proper rotations, sign on the smallest sigma.
 *   torch.linalg.inv / Tensor.inverse (simulator/solver.py:357,508) ->
 *     Gauss-Jordan with partial pivoting.
 *   kornia create_meshgrid3d (+[1,2] swap, solver.py:162-169,235-240) ->
 *     integer cell coordinates (i,j,k) in C-order of the mask.
 *
 * Each function cites the reference file:line it follows.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    double dt, dx, stiff, kdx;
    int iters, kres;
    int res[3];
    double base[3], gravity[3];
    int n_pts, n_ip, n_k, n;            /* n = 10 * n_k scalar unknowns per component */
    double *pos, *mass, *mu, *lam;
    unsigned char *is_pin;
    int *pts_ip, *pts_kernel, *ip_kernel, *ip_grid;
    double *ip_pos, *kernel_pos;
    double *pts_Nx;                      /* [n_pts,8,10] */
    double *ip_Nx, *ip_dNx, *ip_ddNx;    /* [n_ip,8,10] [n_ip,8,3,10] [n_ip,8,3,3,10] */
    double *ip_mu, *ip_lam, *ip_rho;
    double *A, *Ainv, *M;                /* [n,n] scalar (x I3) */
    unsigned char *active;               /* [n_k] */
    double *dof, *dof_rest, *dof_vel, *dof_f, *rhs_rest, *rhs_gravity; /* [3n] */
    int threads;
} QO;

/* ---------- func_utils.py ---------- */

/* simulator/func_utils.py:73-81 */
static int idx2(int x, int y) {
    if (x > y) { int t = x; x = y; y = t; }
    return x == 0 ? 4 + y : 5 + x + y;
}
/* simulator/func_utils.py:84-92 */
static void P0(const double *p, double *a) {
    a[0] = 1; a[1] = p[0]; a[2] = p[1]; a[3] = p[2];
    a[4] = p[0]*p[0]; a[5] = p[0]*p[1]; a[6] = p[0]*p[2];
    a[7] = p[1]*p[1]; a[8] = p[1]*p[2]; a[9] = p[2]*p[2];
}
/* simulator/func_utils.py:95-103 */
static void P1(const double *p, int j, double *a) {
    for (int i = 0; i < 10; i++) a[i] = 0;
    a[j + 1] = 1.0;
    for (int i = 0; i < 3; i++) a[idx2(i, j)] = p[i];
    a[idx2(j, j)] += p[j];
}
/* simulator/func_utils.py:106-112 */
static void P2(int j, int k, double *a) {
    for (int i = 0; i < 10; i++) a[i] = 0;
    a[idx2(j, k)] = 1;
    if (j == k) a[idx2(j, k)] += 1;
}
/* simulator/func_utils.py:43-70: weight (1-d^2)^3 and analytic derivatives */
static void weight_fn(double r, const double *p, const double *q, double *w, double *dw, double *ddw) {
    double e[3] = { p[0]-q[0], p[1]-q[1], p[2]-q[2] };
    double d = sqrt(e[0]*e[0] + e[1]*e[1] + e[2]*e[2]) / r;
    if (d >= 1) {
        *w = 0; for (int i = 0; i < 3; i++) dw[i] = 0; for (int i = 0; i < 9; i++) ddw[i] = 0;
        return;
    }
    double s = 1.0 - d*d, r2 = r*r;
    *w = s*s*s;
    for (int i = 0; i < 3; i++) dw[i] = -6.0 * (s*s) * e[i] / r2;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        ddw[i*3+j] = -6.0 * (s*s) * (i == j ? 1.0 : 0.0) / r2 + 24.0 * s * (e[i]/r2) * (e[j]/r2);
}
/* simulator/func_utils.py:21-40 */
static void volume_invariant_project(const double *sig, double *out) {
    double D[3] = {0, 0, 0};
    for (int it = 0; it < 3; it++) {
        double a = sig[0]+D[0], b = sig[1]+D[1], c = sig[2]+D[2];
        double C = a*b*c - 1.0;
        double dC[3] = { b*c, a*c, a*b };
        double dCTD = dC[0]*D[0] + dC[1]*D[1] + dC[2]*D[2];
        double coef = (dCTD - C) / (dC[0]*dC[0] + dC[1]*dC[1] + dC[2]*dC[2]);
        D[0] = coef*dC[0]; D[1] = coef*dC[1]; D[2] = coef*dC[2];
    }
    out[0] = sig[0]+D[0]; out[1] = sig[1]+D[1]; out[2] = sig[2]+D[2];
}

/* ---------- small dense helpers ---------- */
static void mv10(const double *A, const double *v, double *o) {
    for (int i = 0; i < 10; i++) { double s = 0; for (int j = 0; j < 10; j++) s += A[i*10+j]*v[j]; o[i] = s; }
}
static double dot10(const double *a, const double *b) { double s = 0; for (int i = 0; i < 10; i++) s += a[i]*b[i]; return s; }

/* Gauss-Jordan inverse with partial pivoting (stands in for torch.linalg.inv, solver.py:357,508). */
static int invert_dense(double *A, double *Ai, int n, int threads) {
    (void)threads;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Ai[(size_t)i*n+j] = (i == j);
    for (int c = 0; c < n; c++) {
        int piv = c; double best = fabs(A[(size_t)c*n+c]);
        for (int r = c+1; r < n; r++) { double v = fabs(A[(size_t)r*n+c]); if (v > best) { best = v; piv = r; } }
        if (best == 0.0) return -1;
        if (piv != c) for (int j = 0; j < n; j++) {
            double t = A[(size_t)c*n+j]; A[(size_t)c*n+j] = A[(size_t)piv*n+j]; A[(size_t)piv*n+j] = t;
            t = Ai[(size_t)c*n+j]; Ai[(size_t)c*n+j] = Ai[(size_t)piv*n+j]; Ai[(size_t)piv*n+j] = t;
        }
        double ip = 1.0 / A[(size_t)c*n+c];
        for (int j = 0; j < n; j++) { A[(size_t)c*n+j] *= ip; Ai[(size_t)c*n+j] *= ip; }
        #pragma omp parallel for schedule(static) if (n > 256)
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            double f = A[(size_t)r*n+c];
            if (f == 0.0) continue;
            double *Ar = A + (size_t)r*n, *Air = Ai + (size_t)r*n;
            const double *Ac = A + (size_t)c*n, *Aic = Ai + (size_t)c*n;
            for (int j = 0; j < n; j++) { Ar[j] -= f*Ac[j]; Air[j] -= f*Aic[j]; }
        }
    }
    return 0;
}

/* Exact 3x3 SVD via cyclic Jacobi on F^T F.  Convention of wp.svd3: U, V proper
 * rotations, any reflection carried by the (smallest-magnitude) singular value.
 * F, U, V row-major.  (call site: simulator/cuda_utils.py:104-107) */
static void svd3(const double *F, double *U, double *sig, double *V) {
    double S[9], Vv[9] = {1,0,0, 0,1,0, 0,0,1};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += F[k*3+i]*F[k*3+j]; S[i*3+j] = s;
    }
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = fabs(S[1]) + fabs(S[2]) + fabs(S[5]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++) for (int q = p+1; q < 3; q++) {
            double apq = S[p*3+q];
            if (fabs(apq) < 1e-300) continue;
            double theta = (S[q*3+q] - S[p*3+p]) / (2.0*apq);
            double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta*theta + 1.0));
            double c = 1.0 / sqrt(t*t + 1.0), s = t*c;
            for (int k = 0; k < 3; k++) { /* S <- S J */
                double skp = S[k*3+p], skq = S[k*3+q];
                S[k*3+p] = c*skp - s*skq; S[k*3+q] = s*skp + c*skq;
            }
            for (int k = 0; k < 3; k++) { /* S <- J^T S */
                double spk = S[p*3+k], sqk = S[q*3+k];
                S[p*3+k] = c*spk - s*sqk; S[q*3+k] = s*spk + c*sqk;
            }
            for (int k = 0; k < 3; k++) {
                double vkp = Vv[k*3+p], vkq = Vv[k*3+q];
                Vv[k*3+p] = c*vkp - s*vkq; Vv[k*3+q] = s*vkp + c*vkq;
            }
        }
    }
    /* sort eigenvalues descending */
    double lam[3] = { S[0], S[4], S[8] };
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; i++) for (int j = i+1; j < 3; j++) if (lam[ord[j]] > lam[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    double Vs[9];
    for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) Vs[k*3+c] = Vv[k*3+ord[c]];
    /* make V a proper rotation */
    double detV = Vs[0]*(Vs[4]*Vs[8]-Vs[5]*Vs[7]) - Vs[1]*(Vs[3]*Vs[8]-Vs[5]*Vs[6]) + Vs[2]*(Vs[3]*Vs[7]-Vs[4]*Vs[6]);
    if (detV < 0) for (int k = 0; k < 3; k++) Vs[k*3+2] = -Vs[k*3+2];
    for (int c = 0; c < 3; c++) { double l = lam[ord[c]]; sig[c] = l > 0 ? sqrt(l) : 0.0; }
    /* U columns: F v_c / sigma_c for the two largest, third = cross so det(U)=+1 */
    double B[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { double s = 0; for (int k = 0; k < 3; k++) s += F[r*3+k]*Vs[k*3+c]; B[r*3+c] = s; }
    double u0[3], u1[3], u2[3];
    double n0 = sqrt(B[0]*B[0]+B[3]*B[3]+B[6]*B[6]);
    if (n0 > 0) { u0[0]=B[0]/n0; u0[1]=B[3]/n0; u0[2]=B[6]/n0; } else { u0[0]=1; u0[1]=0; u0[2]=0; }
    double d01 = u0[0]*B[1]+u0[1]*B[4]+u0[2]*B[7];
    u1[0]=B[1]-d01*u0[0]; u1[1]=B[4]-d01*u0[1]; u1[2]=B[7]-d01*u0[2];
    double n1 = sqrt(u1[0]*u1[0]+u1[1]*u1[1]+u1[2]*u1[2]);
    if (n1 > 1e-300) { u1[0]/=n1; u1[1]/=n1; u1[2]/=n1; }
    else { /* any unit vector orthogonal to u0 */
        double a[3] = {0,0,0}; int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
        a[m] = 1; double d = u0[m];
        u1[0]=a[0]-d*u0[0]; u1[1]=a[1]-d*u0[1]; u1[2]=a[2]-d*u0[2];
        n1 = sqrt(u1[0]*u1[0]+u1[1]*u1[1]+u1[2]*u1[2]); u1[0]/=n1; u1[1]/=n1; u1[2]/=n1;
    }
    u2[0]=u0[1]*u1[2]-u0[2]*u1[1]; u2[1]=u0[2]*u1[0]-u0[0]*u1[2]; u2[2]=u0[0]*u1[1]-u0[1]*u1[0];
    /* sigma_2 carries the sign: sigma_2 = u2 . (F v2) */
    sig[2] = u2[0]*B[2] + u2[1]*B[5] + u2[2]*B[8];
    for (int r = 0; r < 3; r++) { U[r*3+0] = u0[r]; U[r*3+1] = u1[r]; U[r*3+2] = u2[r]; }
    memcpy(V, Vs, sizeof(Vs));
}

/* exported for the numpy cross-check in tests */
void qo_svd3(const double *F, double *U, double *sig, double *V) { svd3(F, U, sig, V); }
void qo_volume_project(const double *s, double *o) { volume_invariant_project(s, o); }
int qo_invert(double *A, double *Ai, int n) { return invert_dense(A, Ai, n, 0); }

/* ---------- cpu_utils.py: shape functions ---------- */

/* simulator/cpu_utils.py:3-152 (calc_G, calc_Gp, calc_weight) + solver.py:334-399 (init_GMLS)
 * for one point.  Nx[8][10], dNx[8][3][10], ddNx[8][3][3][10] are zero-filled by the caller. */
static int gmls_point(double r, const double *p, const int *topo, const double *kernel_pos,
                      double *Nx, double *dNx, double *ddNx) {
    double G[100], dG[3][100], ddG[9][100], Gi[100], Gw[100];
    memset(G, 0, sizeof(G)); memset(dG, 0, sizeof(dG)); memset(ddG, 0, sizeof(ddG));
    double wgt[8], dwg[8][3], ddwg[8][9];
    for (int i = 0; i < 8; i++) {
        const double *q = kernel_pos + 3*topo[i];
        weight_fn(r, p, q, &wgt[i], dwg[i], ddwg[i]);
        if (wgt[i] <= 0.0) continue;                      /* cpu_utils.py:27-28 */
        double prim[100], a[10];
        P0(q, a);
        for (int x = 0; x < 10; x++) for (int y = 0; y < 10; y++) prim[x*10+y] = a[x]*a[y];
        for (int j = 0; j < 3; j++) {
            P1(q, j, a);
            for (int x = 0; x < 10; x++) for (int y = 0; y < 10; y++) prim[x*10+y] += a[x]*a[y];
            for (int k = 0; k < 3; k++) {
                P2(j, k, a);
                for (int x = 0; x < 10; x++) for (int y = 0; y < 10; y++) prim[x*10+y] += a[x]*a[y];
            }
        }
        for (int e = 0; e < 100; e++) {
            G[e] += wgt[i]*prim[e];
            for (int x = 0; x < 3; x++) {
                dG[x][e] += dwg[i][x]*prim[e];
                for (int y = 0; y < 3; y++) ddG[x*3+y][e] += ddwg[i][x*3+y]*prim[e];
            }
        }
    }
    memcpy(Gw, G, sizeof(G));
    if (invert_dense(Gw, Gi, 10, 0) != 0) return -1;     /* solver.py:357 */

    /* calc_Gp, literal operator order of cpu_utils.py:69-87 */
    double Pv[10], Gp[10], dGp[3][10], ddGp[9][10], t0[10], t1[10], t2[10], t3[10], t4[10];
    P0(p, Pv);
    mv10(Gi, Pv, Gp);
    for (int x = 0; x < 3; x++) {
        double dPv[10]; P1(p, x, dPv);
        mv10(Gi, dPv, t0);                               /* G_i dPv */
        mv10(dG[x], Gp, t1); mv10(Gi, t1, t2);           /* G_i dG_x G_i Pv */
        for (int e = 0; e < 10; e++) dGp[x][e] = t0[e] - t2[e];
    }
    for (int x = 0; x < 3; x++) {
        double dPx[10]; P1(p, x, dPx);
        for (int y = 0; y < 3; y++) {
            double ddPv[10], dPy[10]; P2(x, y, ddPv); P1(p, y, dPy);
            double acc[10];
            mv10(Gi, ddPv, acc);
            mv10(Gi, dPy, t0); mv10(dG[x], t0, t1); mv10(Gi, t1, t2);      /* G_i dGx G_i Pj(p,y) */
            for (int e = 0; e < 10; e++) acc[e] -= t2[e];
            mv10(Gi, dPx, t0); mv10(dG[y], t0, t1); mv10(Gi, t1, t2);      /* G_i dGy G_i dPv */
            for (int e = 0; e < 10; e++) acc[e] -= t2[e];
            mv10(ddG[x*3+y], Gp, t1); mv10(Gi, t1, t2);                    /* G_i ddG G_i Pv */
            for (int e = 0; e < 10; e++) acc[e] -= t2[e];
            mv10(dG[x], Gp, t1); mv10(Gi, t1, t2); mv10(dG[y], t2, t3); mv10(Gi, t3, t4); /* G_i dGy G_i dGx G_i Pv */
            for (int e = 0; e < 10; e++) acc[e] += t4[e];
            mv10(dG[y], Gp, t1); mv10(Gi, t1, t2); mv10(dG[x], t2, t3); mv10(Gi, t3, t4); /* G_i dGx G_i dGy G_i Pv */
            for (int e = 0; e < 10; e++) acc[e] += t4[e];
            memcpy(ddGp[x*3+y], acc, sizeof(acc));
        }
    }
    /* calc_weight, cpu_utils.py:108-152 */
    for (int i = 0; i < 8; i++) {
        if (wgt[i] <= 0.0) continue;
        const double *q = kernel_pos + 3*topo[i];
        for (int slot_src = 0; slot_src < 13; slot_src++) {
            double a[10]; int slot;
            if (slot_src == 0) { P0(q, a); slot = 0; }
            else if (slot_src < 4) { P1(q, slot_src-1, a); slot = slot_src; }
            else { int x = (slot_src-4)/3, y = (slot_src-4)%3; P2(x, y, a); slot = idx2(x, y); }
            double g = dot10(Gp, a);
            Nx[i*10+slot] += g * wgt[i];
            for (int j = 0; j < 3; j++) {
                double gj = dot10(dGp[j], a);
                dNx[(i*3+j)*10+slot] += g*dwg[i][j] + gj*wgt[i];
                for (int k = 0; k < 3; k++)
                    ddNx[((i*3+j)*3+k)*10+slot] += g*ddwg[i][j*3+k] + dot10(dGp[k], a)*dwg[i][j]
                                                 + gj*dwg[i][k] + dot10(ddGp[j*3+k], a)*wgt[i];
            }
        }
    }
    return 0;
}

/* ---------- solver.py ---------- */

QO *qo_create(double dt, int iters, const double *bbox, int kres, double dx,
              const double *gravity, double stiff, const double *base) {
    /* simulator/solver.py:13-39; bbox/base arrive already scaled by 1.02/1.01
     * (the caller does that in the caller's dtype, as the reference does in place). */
    QO *s = (QO*)calloc(1, sizeof(QO));
    s->dt = dt; s->iters = iters; s->kres = kres; s->dx = dx; s->stiff = stiff;
    for (int i = 0; i < 3; i++) {
        s->res[i] = (int)floor(bbox[i] / dx);             /* solver.py:32 */
        s->base[i] = base[i]; s->gravity[i] = gravity[i];
    }
    return s;
}

void qo_destroy(QO *s) {
    if (!s) return;
    free(s->pos); free(s->mass); free(s->mu); free(s->lam); free(s->is_pin);
    free(s->pts_ip); free(s->pts_kernel); free(s->ip_kernel); free(s->ip_grid);
    free(s->ip_pos); free(s->kernel_pos); free(s->pts_Nx);
    free(s->ip_Nx); free(s->ip_dNx); free(s->ip_ddNx);
    free(s->ip_mu); free(s->ip_lam); free(s->ip_rho);
    free(s->A); free(s->Ainv); free(s->M); free(s->active);
    free(s->dof); free(s->dof_rest); free(s->dof_vel); free(s->dof_f); free(s->rhs_rest); free(s->rhs_gravity);
    free(s);
}

/* simulator/cuda_utils.py:83-151 (calc_elastic + collect_rhs_IP) via solver.py:541-571 */
static void build_rhs(const QO *s, const double *dof, double *rhs) {
    int n3 = 3*s->n;
    for (int i = 0; i < n3; i++) rhs[i] = 0;
    double dx3 = s->dx*s->dx*s->dx;
    /* the reference scatters with fp64 atomics (order-nondeterministic); here each
     * thread owns a private accumulator that is reduced in thread order. */
    #pragma omp parallel if (s->n_ip > 512)
    {
    double *acc = rhs;
    int shared_acc = 1;
#ifdef _OPENMP
    if (omp_get_num_threads() > 1) { acc = (double*)calloc(n3, sizeof(double)); shared_acc = 0; }
#endif
    #pragma omp for schedule(static)
    for (int v = 0; v < s->n_ip; v++) {
        const int *topo = s->ip_kernel + 8*v;
        const double *dN = s->ip_dNx + (size_t)v*240;
        double F[9] = {0};
        for (int i = 0; i < 8; i++) for (int x = 0; x < 10; x++) {
            const double *d = dof + 3*(topo[i]*10 + x);
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) F[r*3+c] += d[r]*dN[(i*3+c)*10+x];
        }
        double U[9], sg[3], V[9], sp[3], R[9], W[9];
        svd3(F, U, sg, V);
        volume_invariant_project(sg, sp);
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
            double a = 0, b = 0;
            for (int k = 0; k < 3; k++) { a += U[r*3+k]*V[c*3+k]; b += U[r*3+k]*sp[k]*V[c*3+k]; }
            R[r*3+c] = a; W[r*3+c] = b;
        }
        double mu = s->ip_mu[v], lam = s->ip_lam[v], Mx[9];
        for (int e = 0; e < 9; e++) Mx[e] = dx3 * (mu*R[e] + lam*W[e]);
        for (int i = 0; i < 8; i++) for (int x = 0; x < 10; x++) {
            double *o = acc + 3*(topo[i]*10 + x);
            for (int r = 0; r < 3; r++)
                o[r] += Mx[r*3+0]*dN[(i*3+0)*10+x] + Mx[r*3+1]*dN[(i*3+1)*10+x] + Mx[r*3+2]*dN[(i*3+2)*10+x];
        }
    }
    if (!shared_acc) {
        #pragma omp critical
        for (int i = 0; i < n3; i++) rhs[i] += acc[i];
        free(acc);
    }
    }
}

/* y[n,3] = Mat[n,n] x[n,3]  (the reference multiplies by Mat (x) I3, solver.py:493-496,576,600) */
static void matvec3(const double *Mat, const double *x, double *y, int n) {
    #pragma omp parallel for schedule(static) if (n > 256)
    for (int r = 0; r < n; r++) {
        const double *row = Mat + (size_t)r*n; double a = 0, b = 0, c = 0;
        for (int j = 0; j < n; j++) { double m = row[j]; a += m*x[3*j]; b += m*x[3*j+1]; c += m*x[3*j+2]; }
        y[3*r] = a; y[3*r+1] = b; y[3*r+2] = c;
    }
}

/* simulator/cuda_utils.py:22-55 build_IP_global */
static void assemble(const QO *s, const double *mu, const double *lam, double *mat) {
    int n = s->n; double dx = s->dx, dt = s->dt;
    double dx3 = dx*dx*dx, dx5 = dx3*dx*dx, dt2 = dt*dt;
    memset(mat, 0, sizeof(double)*(size_t)n*n);
    for (int v = 0; v < s->n_ip; v++) {
        const int *topo = s->ip_kernel + 8*v;
        const double *N = s->ip_Nx + (size_t)v*80, *dN = s->ip_dNx + (size_t)v*240, *ddN = s->ip_ddNx + (size_t)v*720;
        double rho = s->ip_rho[v], ml = (mu ? mu[v] : 0.0) + (lam ? lam[v] : 0.0);
        double c0 = rho*dx3/dt2, c1 = dx3*(rho*(dx*dx)/12.0/dt2 + ml), c2 = dx5*ml/12.0;
        for (int i = 0; i < 8; i++) for (int x = 0; x < 10; x++) {
            int r = topo[i]*10 + x;
            for (int j = 0; j < 8; j++) for (int y = 0; y < 10; y++) {
                int c = topo[j]*10 + y;
                double acc = c0 * N[i*10+x]*N[j*10+y];
                for (int p = 0; p < 3; p++) {
                    acc += c1 * dN[(i*3+p)*10+x]*dN[(j*3+p)*10+y];
                    for (int q = 0; q < 3; q++) acc += c2 * ddN[((i*3+p)*3+q)*10+x]*ddN[((j*3+p)*3+q)*10+y];
                }
                mat[(size_t)r*n+c] += acc;
            }
        }
    }
}

int qo_initialize(QO *s, int n_pts, const double *pos, const double *mass, const double *mu,
                  const double *lam, const unsigned char *is_pin) {
    /* simulator/solver.py:139-331 */
    s->n_pts = n_pts;
    s->pos = (double*)malloc(sizeof(double)*3*n_pts); memcpy(s->pos, pos, sizeof(double)*3*n_pts);
    s->mass = (double*)malloc(sizeof(double)*n_pts); memcpy(s->mass, mass, sizeof(double)*n_pts);
    s->mu = (double*)malloc(sizeof(double)*n_pts); memcpy(s->mu, mu, sizeof(double)*n_pts);
    s->lam = (double*)malloc(sizeof(double)*n_pts); memcpy(s->lam, lam, sizeof(double)*n_pts);
    s->is_pin = (unsigned char*)malloc(n_pts); memcpy(s->is_pin, is_pin, n_pts);
    const int R0 = s->res[0], R1 = s->res[1], R2 = s->res[2], K = s->kres;
    size_t ncell = (size_t)R0*R1*R2;
    int *ip_idx = (int*)malloc(sizeof(int)*ncell);
    unsigned char *mask = (unsigned char*)calloc(ncell, 1);
    int *gi = (int*)malloc(sizeof(int)*3*n_pts);
    for (int p = 0; p < n_pts; p++) {                      /* solver.py:141-148 */
        for (int c = 0; c < 3; c++) gi[3*p+c] = (int)floor((pos[3*p+c] - s->base[c]) / s->dx);
        if (gi[3*p] < 0 || gi[3*p] >= R0 || gi[3*p+1] < 0 || gi[3*p+1] >= R1 || gi[3*p+2] < 0 || gi[3*p+2] >= R2) return -2;
        mask[((size_t)gi[3*p]*R1 + gi[3*p+1])*R2 + gi[3*p+2]] = 1;
    }
    int n_ip = 0;
    for (size_t c = 0; c < ncell; c++) ip_idx[c] = mask[c] ? n_ip++ : -1;   /* solver.py:150-154 */
    s->n_ip = n_ip;
    s->pts_ip = (int*)malloc(sizeof(int)*n_pts);
    for (int p = 0; p < n_pts; p++) s->pts_ip[p] = ip_idx[((size_t)gi[3*p]*R1 + gi[3*p+1])*R2 + gi[3*p+2]];
    s->ip_grid = (int*)malloc(sizeof(int)*3*n_ip);
    s->ip_pos = (double*)malloc(sizeof(double)*3*n_ip);
    for (int i = 0, k = 0; i < R0; i++) for (int j = 0; j < R1; j++) for (int l = 0; l < R2; l++)
        if (mask[((size_t)i*R1 + j)*R2 + l]) {             /* solver.py:162-177 */
            int g[3] = {i, j, l};
            /* solver.py:177: `(IP_grid + 0.5) * dx` is int32 + python float -> float32 in torch, widened by `+ base` */
            for (int c = 0; c < 3; c++) { s->ip_grid[3*k+c] = g[c]; s->ip_pos[3*k+c] = (double)(((float)g[c] + 0.5f)*(float)s->dx) + s->base[c]; }
            k++;
        }
    int rmax = R0 > R1 ? (R0 > R2 ? R0 : R2) : (R1 > R2 ? R1 : R2);
    /* solver.py:184: `res.max() * dx / (kres-1)` is an int32 0-dim tensor times python floats, which torch
     * evaluates in float32 (default dtype); the value is then used as a double everywhere. */
    const float kdx32 = ((float)rmax * (float)s->dx) / (float)(K - 1);
    s->kdx = (double)kdx32;
    unsigned char *kmask = (unsigned char*)calloc((size_t)K*K*K, 1);
    int *kidx = (int*)calloc((size_t)K*K*K, sizeof(int)); /* 0 (not -1) where unmasked, solver.py:204-208 */
    int *ip2k = (int*)malloc(sizeof(int)*3*n_ip);
    for (int v = 0; v < n_ip; v++) {
        for (int c = 0; c < 3; c++) ip2k[3*v+c] = (int)floor((s->ip_pos[3*v+c] - s->base[c]) / s->kdx);
        for (int S = 0; S < 8; S++) {                      /* solver.py:193-202 */
            int a = ip2k[3*v] + (S>>2&1), b = ip2k[3*v+1] + (S>>1&1), c = ip2k[3*v+2] + (S&1);
            if (a < 0 || a >= K || b < 0 || b >= K || c < 0 || c >= K) return -3;
            kmask[(a*K + b)*K + c] = 1;
        }
    }
    int n_k = 0;
    for (int c = 0; c < K*K*K; c++) if (kmask[c]) kidx[c] = n_k++;
    s->n_k = n_k; s->n = 10*n_k;
    s->kernel_pos = (double*)malloc(sizeof(double)*3*n_k);
    for (int a = 0, k = 0; a < K; a++) for (int b = 0; b < K; b++) for (int c = 0; c < K; c++)
        if (kmask[(a*K + b)*K + c]) {                      /* solver.py:235-248 */
            /* solver.py:248: int32 grid * float32 0-dim kdx is a float32 product, widened by `+ base` (fp64) */
            s->kernel_pos[3*k] = (double)((float)a*kdx32) + s->base[0];
            s->kernel_pos[3*k+1] = (double)((float)b*kdx32) + s->base[1];
            s->kernel_pos[3*k+2] = (double)((float)c*kdx32) + s->base[2];
            k++;
        }
    s->ip_kernel = (int*)malloc(sizeof(int)*8*n_ip);
    s->pts_kernel = (int*)malloc(sizeof(int)*8*n_pts);
    for (int v = 0; v < n_ip; v++) for (int S = 0; S < 8; S++)
        s->ip_kernel[8*v+S] = kidx[((ip2k[3*v] + (S>>2&1))*K + ip2k[3*v+1] + (S>>1&1))*K + ip2k[3*v+2] + (S&1)];
    for (int p = 0; p < n_pts; p++) {                      /* solver.py:215-233 */
        int k3[3];
        for (int c = 0; c < 3; c++) k3[c] = (int)floor((pos[3*p+c] - s->base[c]) / s->kdx);
        for (int S = 0; S < 8; S++) {
            int a = k3[0] + (S>>2&1), b = k3[1] + (S>>1&1), c = k3[2] + (S&1);
            if (a < 0 || a >= K || b < 0 || b >= K || c < 0 || c >= K) return -3;
            s->pts_kernel[8*p+S] = kidx[(a*K + b)*K + c];
        }
    }
    free(ip_idx); free(mask); free(gi); free(kmask); free(kidx); free(ip2k);

    /* init_GMLS for sample points (only Nx is ever used) and IPs: solver.py:250-252 */
    s->pts_Nx = (double*)calloc((size_t)n_pts*80, sizeof(double));
    s->ip_Nx = (double*)calloc((size_t)n_ip*80, sizeof(double));
    s->ip_dNx = (double*)calloc((size_t)n_ip*240, sizeof(double));
    s->ip_ddNx = (double*)calloc((size_t)n_ip*720, sizeof(double));
    int bad = 0;
    #pragma omp parallel for schedule(dynamic, 16) reduction(+:bad)
    for (int p = 0; p < n_pts; p++) {
        double d1[240], d2[720]; memset(d1, 0, sizeof(d1)); memset(d2, 0, sizeof(d2));
        if (gmls_point(s->kdx, s->pos + 3*p, s->pts_kernel + 8*p, s->kernel_pos, s->pts_Nx + (size_t)p*80, d1, d2)) bad++;
    }
    #pragma omp parallel for schedule(dynamic, 16) reduction(+:bad)
    for (int v = 0; v < n_ip; v++)
        if (gmls_point(s->kdx, s->ip_pos + 3*v, s->ip_kernel + 8*v, s->kernel_pos,
                       s->ip_Nx + (size_t)v*80, s->ip_dNx + (size_t)v*240, s->ip_ddNx + (size_t)v*720)) bad++;
    if (bad) return -4;

    /* collect_IP: solver.py:427-450, cuda_utils.py:3-19 */
    s->ip_mu = (double*)calloc(n_ip, sizeof(double)); s->ip_lam = (double*)calloc(n_ip, sizeof(double)); s->ip_rho = (double*)calloc(n_ip, sizeof(double));
    for (int p = 0; p < n_pts; p++) {
        int v = s->pts_ip[p];
        s->ip_mu[v] += mu[p]*mass[p]; s->ip_lam[v] += lam[p]*mass[p]; s->ip_rho[v] += mass[p];
    }
    double dx3 = s->dx*s->dx*s->dx;
    for (int v = 0; v < n_ip; v++) { s->ip_mu[v] /= s->ip_rho[v]; s->ip_lam[v] /= s->ip_rho[v]; s->ip_rho[v] /= dx3; }

    /* build_global: solver.py:453-538 */
    int n = s->n;
    s->A = (double*)malloc(sizeof(double)*(size_t)n*n);
    s->M = (double*)malloc(sizeof(double)*(size_t)n*n);
    s->Ainv = (double*)calloc((size_t)n*n, sizeof(double));
    s->active = (unsigned char*)calloc(n_k, 1);
    assemble(s, s->ip_mu, s->ip_lam, s->A);
    for (int p = 0; p < n_pts; p++) if (is_pin[p]) {       /* cuda_utils.py:58-81 */
        const int *topo = s->pts_kernel + 8*p; const double *N = s->pts_Nx + (size_t)p*80;
        for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) for (int x = 0; x < 10; x++) for (int y = 0; y < 10; y++)
            s->A[(size_t)(topo[i]*10+x)*n + topo[j]*10+y] += s->stiff * N[i*10+x]*N[j*10+y];
    }
    int na = 0; int *lst = (int*)malloc(sizeof(int)*n);
    for (int k = 0; k < n_k; k++) if (s->A[(size_t)(k*10)*n + k*10] > 0.0) {   /* solver.py:499-503 */
        s->active[k] = 1; for (int x = 0; x < 10; x++) lst[na++] = k*10 + x;
    }
    double *sub = (double*)malloc(sizeof(double)*(size_t)na*na), *subi = (double*)malloc(sizeof(double)*(size_t)na*na);
    for (int i = 0; i < na; i++) for (int j = 0; j < na; j++) sub[(size_t)i*na+j] = s->A[(size_t)lst[i]*n + lst[j]] + (i == j ? 1e-3 : 0.0);
    if (invert_dense(sub, subi, na, 0)) return -5;         /* solver.py:505-511 */
    for (int i = 0; i < na; i++) for (int j = 0; j < na; j++) s->Ainv[(size_t)lst[i]*n + lst[j]] = subi[(size_t)i*na+j];
    free(sub); free(subi); free(lst);
    assemble(s, NULL, NULL, s->M);                          /* solver.py:513-538 */

    /* dof layout: solver.py:258-277 */
    int n3 = 3*n;
    s->dof = (double*)calloc(n3, sizeof(double)); s->dof_rest = (double*)calloc(n3, sizeof(double));
    s->dof_vel = (double*)calloc(n3, sizeof(double)); s->dof_f = (double*)calloc(n3, sizeof(double));
    s->rhs_rest = (double*)calloc(n3, sizeof(double)); s->rhs_gravity = (double*)calloc(n3, sizeof(double));
    for (int k = 0; k < n_k; k++) for (int x = 0; x < 3; x++) {
        s->dof[k*30 + x] = s->kernel_pos[3*k+x];
        s->dof[k*30 + 3 + x*3 + x] = 1;
    }
    memcpy(s->dof_rest, s->dof, sizeof(double)*n3);
    /* rhs_rest: solver.py:314 */
    double *tmp = (double*)malloc(sizeof(double)*n3);
    build_rhs(s, s->dof, s->rhs_rest);
    matvec3(s->M, s->dof, tmp, n);
    for (int i = 0; i < n3; i++) s->rhs_rest[i] += tmp[i];
    free(tmp);
    /* gravity: solver.py:316-331, cuda_utils.py:262-279 */
    for (int v = 0; v < n_ip; v++) {
        double m = s->ip_rho[v]*s->dx*s->dx*s->dx;
        for (int i = 0; i < 8; i++) for (int x = 0; x < 10; x++) for (int c = 0; c < 3; c++)
            s->rhs_gravity[3*(s->ip_kernel[8*v+i]*10 + x) + c] += m * s->ip_Nx[(size_t)v*80 + i*10 + x] * s->gravity[c];
    }
    return 0;
}

/* simulator/solver.py:574-576, 595-602 */
void qo_step(QO *s) {
    int n = s->n, n3 = 3*n;
    double *tilde = (double*)malloc(sizeof(double)*n3), *mom = (double*)malloc(sizeof(double)*n3);
    double *last = (double*)malloc(sizeof(double)*n3), *rhs = (double*)malloc(sizeof(double)*n3), *x = (double*)malloc(sizeof(double)*n3);
    for (int i = 0; i < n3; i++) tilde[i] = s->dof[i] + s->dt*s->dof_vel[i];
    matvec3(s->M, tilde, mom, n);
    for (int i = 0; i < n3; i++) mom[i] += s->dof_f[i] + s->rhs_gravity[i];
    memcpy(last, s->dof, sizeof(double)*n3);
    for (int it = 0; it < s->iters; it++) {
        build_rhs(s, s->dof, rhs);
        for (int i = 0; i < n3; i++) rhs[i] = mom[i] + rhs[i] - s->rhs_rest[i];
        matvec3(s->Ainv, rhs, x, n);
        for (int i = 0; i < n3; i++) s->dof[i] = s->dof_rest[i] + x[i];
    }
    for (int i = 0; i < n3; i++) s->dof_vel[i] = (s->dof[i] - last[i]) / s->dt * 0.998;
    free(tilde); free(mom); free(last); free(rhs); free(x);
}

/* simulator/solver.py:402-424 + cuda_utils.py:206-233; fp32 outputs in the renderer's layouts:
 * F[a*3+b] = F[b][a],  dF[c*9+r*3+j] = dF_j[r][c]. Also returns fp64 positions if pos64 != NULL. */
void qo_ip_info(const QO *s, float *pos, float *F, float *dF, double *pos64) {
    for (int v = 0; v < s->n_ip; v++) {
        const int *topo = s->ip_kernel + 8*v;
        const double *N = s->ip_Nx + (size_t)v*80, *dN = s->ip_dNx + (size_t)v*240, *ddN = s->ip_ddNx + (size_t)v*720;
        double p[3] = {0}, Fm[9] = {0}, dFm[27] = {0};
        for (int i = 0; i < 8; i++) for (int x = 0; x < 10; x++) {
            const double *d = s->dof + 3*(topo[i]*10 + x);
            for (int r = 0; r < 3; r++) {
                p[r] += N[i*10+x]*d[r];
                for (int c = 0; c < 3; c++) {
                    Fm[r*3+c] += d[r]*dN[(i*3+c)*10+x];
                    for (int j = 0; j < 3; j++) dFm[(j*3+r)*3+c] += d[r]*ddN[((i*3+j)*3+c)*10+x];
                }
            }
        }
        for (int r = 0; r < 3; r++) { pos[3*v+r] = (float)p[r]; if (pos64) pos64[3*v+r] = p[r]; }
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) F[9*v + a*3+b] = (float)Fm[b*3+a];
        for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) for (int j = 0; j < 3; j++)
            dF[27*v + c*9 + r*3 + j] = (float)dFm[(j*3+r)*3+c];
    }
}

/* simulator/solver.py:578-593 */
void qo_update_force(QO *s, int vid, const double *f) {
    int n3 = 3*s->n;
    for (int i = 0; i < n3; i++) s->dof_f[i] = 0;
    if (vid < 0) return;
    double m = s->ip_rho[vid]*s->dx*s->dx*s->dx;
    for (int i = 0; i < 8; i++) for (int j = 0; j < 10; j++) for (int c = 0; c < 3; c++)
        s->dof_f[3*(s->ip_kernel[8*vid+i]*10 + j) + c] += m * s->ip_Nx[(size_t)vid*80 + i*10 + j] * f[c];
}

/* simulator/solver.py:604-617, cuda_utils.py:191-203 */
void qo_update_pos(const QO *s, double *out) {
    for (int p = 0; p < s->n_pts; p++) {
        double a[3] = {0};
        for (int i = 0; i < 8; i++) for (int j = 0; j < 10; j++) for (int c = 0; c < 3; c++)
            a[c] += s->pts_Nx[(size_t)p*80 + i*10 + j] * s->dof[3*(s->pts_kernel[8*p+i]*10 + j) + c];
        out[3*p] = a[0]; out[3*p+1] = a[1]; out[3*p+2] = a[2];
    }
}

/* accessors for ctypes */
int qo_n_ip(const QO *s) { return s->n_ip; }
int qo_n_k(const QO *s) { return s->n_k; }
int qo_n_pts(const QO *s) { return s->n_pts; }
double qo_kdx(const QO *s) { return s->kdx; }
void qo_res(const QO *s, int *o) { o[0] = s->res[0]; o[1] = s->res[1]; o[2] = s->res[2]; }
const void *qo_ptr(const QO *s, const char *name) {
#define F(x) if (!strcmp(name, #x)) return s->x;
    F(pts_ip) F(pts_kernel) F(ip_kernel) F(ip_grid) F(ip_pos) F(kernel_pos) F(pts_Nx) F(ip_Nx) F(ip_dNx) F(ip_ddNx)
    F(ip_mu) F(ip_lam) F(ip_rho) F(A) F(Ainv) F(M) F(active) F(dof) F(dof_rest) F(dof_vel) F(dof_f) F(rhs_rest) F(rhs_gravity)
#undef F
    return NULL;
}
void qo_set_dof(QO *s, const double *dof, const double *vel) {
    if (dof) memcpy(s->dof, dof, sizeof(double)*3*s->n);
    if (vel) memcpy(s->dof_vel, vel, sizeof(double)*3*s->n);
}
void qo_build_rhs(const QO *s, double *out) { build_rhs(s, s->dof, out); }
/* bench.py --impl reference under torchrun: the launcher exports OMP_NUM_THREADS=1; the reference arm must use every host core */
void qo_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int qo_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
