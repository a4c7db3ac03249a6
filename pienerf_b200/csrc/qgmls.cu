// Q-GMLS elastodynamics kernels (fp64) — the `_qgmls` operator set.
//
// The reference runs these as Warp kernels launched from Python (simulator/cuda_utils.py, cpu_utils.py) plus
// dense torch mat-vecs on (Mat (x) I3) (simulator/solver.py:574-602).  Here:
//   * shape functions / assembly (init): one thread per point resp. per (IP, row) — init-time, not tuned;
//   * local step: one WARP per IP gathers its 80 DOF vectors, reduces F with shuffles, lane 0 does the 3x3
//     SVD + projections and writes the 3x3 stress  dx^3 (mu R + lam V);
//   * rhs: deterministic gather over the kernel->(IP,corner) CSR (the adjacency the reference builds at
//     solver.py:279-313 but never uses) — one warp per DOF row, no fp64 atomics, bit-reproducible;
//   * global step: compact [n,n] x [n,3] mat-vec (the reference streams the 9x larger (x)I3 matrix),
//     one warp per row, 128-bit loads, shuffle reduction;
//   * optional PCG on the assembled system instead of the pre-inverted matrix (warp-shuffle dot products);
//   * ip_info: emits the renderer's packed fp32 IP state (pos, F, dF layouts) directly.
#include "common.cuh"

namespace {

constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------- func_utils.py equivalents
__device__ __forceinline__ int sym_slot(int x, int y) {  // func_utils.py:73-81
    if (x > y) { const int t = x; x = y; y = t; }
    return x == 0 ? 4 + y : 5 + x + y;
}
__device__ __forceinline__ void basis0(const double *p, double *a) {  // func_utils.py:84-92
    a[0] = 1; a[1] = p[0]; a[2] = p[1]; a[3] = p[2];
    a[4] = p[0] * p[0]; a[5] = p[0] * p[1]; a[6] = p[0] * p[2];
    a[7] = p[1] * p[1]; a[8] = p[1] * p[2]; a[9] = p[2] * p[2];
}
__device__ __forceinline__ void basis1(const double *p, int j, double *a) {  // func_utils.py:95-103
    for (int i = 0; i < 10; i++) a[i] = 0;
    a[j + 1] = 1.0;
    for (int i = 0; i < 3; i++) a[sym_slot(i, j)] = p[i];
    a[sym_slot(j, j)] += p[j];
}
__device__ __forceinline__ void basis2(int j, int k, double *a) {  // func_utils.py:106-112
    for (int i = 0; i < 10; i++) a[i] = 0;
    a[sym_slot(j, k)] = (j == k) ? 2.0 : 1.0;
}
__device__ __forceinline__ void mv10(const double *A, const double *v, double *o) {
    for (int i = 0; i < 10; i++) {
        double s = 0;
        for (int j = 0; j < 10; j++) s += A[i * 10 + j] * v[j];
        o[i] = s;
    }
}
__device__ __forceinline__ double dot10(const double *a, const double *b) {
    double s = 0;
    for (int i = 0; i < 10; i++) s += a[i] * b[i];
    return s;
}
// in-place Gauss-Jordan inverse of a 10x10 (partial pivoting); returns false when singular
__device__ bool invert10(double *A, double *Ai) {
    for (int i = 0; i < 100; i++) Ai[i] = (i / 10 == i % 10) ? 1.0 : 0.0;
    for (int c = 0; c < 10; c++) {
        int piv = c;
        double best = fabs(A[c * 10 + c]);
        for (int r = c + 1; r < 10; r++) { const double v = fabs(A[r * 10 + c]); if (v > best) { best = v; piv = r; } }
        if (best == 0.0) return false;
        if (piv != c)
            for (int j = 0; j < 10; j++) {
                double t = A[c * 10 + j]; A[c * 10 + j] = A[piv * 10 + j]; A[piv * 10 + j] = t;
                t = Ai[c * 10 + j]; Ai[c * 10 + j] = Ai[piv * 10 + j]; Ai[piv * 10 + j] = t;
            }
        const double ip = 1.0 / A[c * 10 + c];
        for (int j = 0; j < 10; j++) { A[c * 10 + j] *= ip; Ai[c * 10 + j] *= ip; }
        for (int r = 0; r < 10; r++) {
            if (r == c) continue;
            const double f = A[r * 10 + c];
            if (f == 0.0) continue;
            for (int j = 0; j < 10; j++) { A[r * 10 + j] -= f * A[c * 10 + j]; Ai[r * 10 + j] -= f * Ai[c * 10 + j]; }
        }
    }
    return true;
}

// ---------------------------------------------------------------- shape functions (init)
// cpu_utils.py:3-152 for one point per thread.  Uses the algebraically equivalent compact form
//   dGp_x = Gi (P_x - dG_x Gp),  ddGp_xy = Gi (P_xy - dG_x dGp_y - dG_y dGp_x - ddG_xy Gp)
// of the reference's expanded product-rule chain (cpu_utils.py:75-87).
__global__ void __launch_bounds__(64) shape_kernel(double r, const double *__restrict__ pos, const int *__restrict__ topo,
                                                   const double *__restrict__ kpos, int n, double *__restrict__ Nx,
                                                   double *__restrict__ dNx, double *__restrict__ ddNx,
                                                   int *__restrict__ status) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const double p[3] = {pos[3 * v], pos[3 * v + 1], pos[3 * v + 2]};
    double G[100], dG[3][100], ddG[6][100];  // ddG symmetric in (x,y): slots via sym index 0..5
    for (int i = 0; i < 100; i++) {
        G[i] = 0;
        for (int x = 0; x < 3; x++) dG[x][i] = 0;
        for (int x = 0; x < 6; x++) ddG[x][i] = 0;
    }
    auto s6 = [](int x, int y) { if (x > y) { int t = x; x = y; y = t; } return x == 0 ? y : x + y + 1; };  // 00,01,02,11,12,22
    double wgt[8], dwg[8][3], ddwg[8][6];
    const double r2 = r * r;
    for (int i = 0; i < 8; i++) {
        const double *q = kpos + 3 * topo[8 * v + i];
        const double e[3] = {p[0] - q[0], p[1] - q[1], p[2] - q[2]};
        const double d = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) / r;
        if (d >= 1) { wgt[i] = 0; continue; }          // func_utils.py:43-70
        const double s = 1.0 - d * d;
        wgt[i] = s * s * s;
        for (int x = 0; x < 3; x++) dwg[i][x] = -6.0 * (s * s) * e[x] / r2;
        for (int x = 0; x < 3; x++)
            for (int y = x; y < 3; y++)
                ddwg[i][s6(x, y)] = -6.0 * (s * s) * (x == y ? 1.0 : 0.0) / r2 + 24.0 * s * (e[x] / r2) * (e[y] / r2);
        if (wgt[i] <= 0.0) continue;
        double prim[100], a[10];
        basis0(q, a);
        for (int x = 0; x < 10; x++) for (int y = 0; y < 10; y++) prim[x * 10 + y] = a[x] * a[y];
        for (int j = 0; j < 3; j++) {
            basis1(q, j, a);
            for (int x = 0; x < 10; x++) for (int y = 0; y < 10; y++) prim[x * 10 + y] += a[x] * a[y];
            for (int k = 0; k < 3; k++) {
                basis2(j, k, a);
                for (int x = 0; x < 10; x++) for (int y = 0; y < 10; y++) prim[x * 10 + y] += a[x] * a[y];
            }
        }
        for (int t = 0; t < 100; t++) {
            G[t] += wgt[i] * prim[t];
            for (int x = 0; x < 3; x++) dG[x][t] += dwg[i][x] * prim[t];
            for (int x = 0; x < 6; x++) ddG[x][t] += ddwg[i][x] * prim[t];
        }
    }
    double Gi[100];
    {
        double Gw[100];
        for (int i = 0; i < 100; i++) Gw[i] = G[i];
        if (!invert10(Gw, Gi)) { atomicExch(status, 1); return; }
    }
    double Pv[10], Gp[10], dGp[3][10], ddGp[6][10], t0[10], t1[10];
    basis0(p, Pv);
    mv10(Gi, Pv, Gp);
    for (int x = 0; x < 3; x++) {
        basis1(p, x, t0);
        mv10(dG[x], Gp, t1);
        for (int i = 0; i < 10; i++) t0[i] -= t1[i];
        mv10(Gi, t0, dGp[x]);
    }
    for (int x = 0; x < 3; x++)
        for (int y = x; y < 3; y++) {
            basis2(x, y, t0);
            mv10(dG[x], dGp[y], t1);
            for (int i = 0; i < 10; i++) t0[i] -= t1[i];
            mv10(dG[y], dGp[x], t1);
            for (int i = 0; i < 10; i++) t0[i] -= t1[i];
            mv10(ddG[s6(x, y)], Gp, t1);
            for (int i = 0; i < 10; i++) t0[i] -= t1[i];
            mv10(Gi, t0, ddGp[s6(x, y)]);
        }
    // calc_weight (cpu_utils.py:108-152): slots 0, 1..3 and the 9 (x,y) second-order terms (off-diagonals twice)
    for (int i = 0; i < 8; i++) {
        double *N = Nx + ((size_t)v * 8 + i) * 10;
        double *dN = dNx ? dNx + ((size_t)v * 8 + i) * 30 : nullptr;
        double *ddN = ddNx ? ddNx + ((size_t)v * 8 + i) * 90 : nullptr;
        for (int t = 0; t < 10; t++) N[t] = 0;
        if (dN) for (int t = 0; t < 30; t++) dN[t] = 0;
        if (ddN) for (int t = 0; t < 90; t++) ddN[t] = 0;
        if (wgt[i] <= 0.0) continue;
        const double *q = kpos + 3 * topo[8 * v + i];
        for (int src = 0; src < 13; src++) {
            double a[10];
            int slot;
            if (src == 0) { basis0(q, a); slot = 0; }
            else if (src < 4) { basis1(q, src - 1, a); slot = src; }
            else { const int x = (src - 4) / 3, y = (src - 4) % 3; basis2(x, y, a); slot = sym_slot(x, y); }
            const double g = dot10(Gp, a);
            N[slot] += g * wgt[i];
            if (!dN) continue;
            double gj[3];
            for (int j = 0; j < 3; j++) { gj[j] = dot10(dGp[j], a); dN[j * 10 + slot] += g * dwg[i][j] + gj[j] * wgt[i]; }
            if (!ddN) continue;
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 3; k++)
                    ddN[(j * 3 + k) * 10 + slot] += g * ddwg[i][s6(j, k)] + gj[k] * dwg[i][j] + gj[j] * dwg[i][k] +
                                                    dot10(ddGp[s6(j, k)], a) * wgt[i];
        }
    }
}

// ---------------------------------------------------------------- collect_param / assembly (init)
__global__ void collect_param_kernel(const int *__restrict__ pts_ip, const double *__restrict__ mu,
                                     const double *__restrict__ lam, const double *__restrict__ mass, int n_pts,
                                     double *ip_mu, double *ip_lam, double *ip_rho) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_pts) return;
    const int ip = pts_ip[v];
    atomicAdd(ip_mu + ip, mu[v] * mass[v]);
    atomicAdd(ip_lam + ip, lam[v] * mass[v]);
    atomicAdd(ip_rho + ip, mass[v]);
}
__global__ void finish_param_kernel(int n_ip, double dx3, double *ip_mu, double *ip_lam, double *ip_rho) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_ip) return;
    const double m = ip_rho[v];
    ip_mu[v] /= m; ip_lam[v] /= m; ip_rho[v] = m / dx3;      // solver.py:450
}

// one thread per (IP, local row (i,x)); loops over the 80 local columns (cuda_utils.py:22-55)
__global__ void __launch_bounds__(128) build_ip_global_kernel(double dx, double dt, const int *__restrict__ topo,
                                                              const double *__restrict__ mu, const double *__restrict__ lam,
                                                              const double *__restrict__ rho, const double *__restrict__ Nx,
                                                              const double *__restrict__ dNx, const double *__restrict__ ddNx,
                                                              int n_ip, int n, double *mat) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_ip * 80) return;
    const int v = id / 80, i = (id % 80) / 10, x = id % 10;
    const double ml = (mu ? mu[v] : 0.0) + (lam ? lam[v] : 0.0), rh = rho[v];
    const double dx3 = dx * dx * dx, dx5 = dx3 * dx * dx, dt2 = dt * dt;
    const double c0 = rh * dx3 / dt2, c1 = dx3 * (rh * (dx * dx) / 12.0 / dt2 + ml), c2 = dx5 * ml / 12.0;
    const double *N = Nx + (size_t)v * 80, *dN = dNx + (size_t)v * 240, *ddN = ddNx + (size_t)v * 720;
    const int r = topo[8 * v + i] * 10 + x;
    double ni = N[i * 10 + x], dni[3], ddni[9];
    for (int p = 0; p < 3; p++) dni[p] = dN[(i * 3 + p) * 10 + x];
    for (int p = 0; p < 9; p++) ddni[p] = ddN[(i * 9 + p) * 10 + x];
    for (int j = 0; j < 8; j++)
        for (int y = 0; y < 10; y++) {
            double acc = c0 * ni * N[j * 10 + y];
            for (int p = 0; p < 3; p++) acc += c1 * dni[p] * dN[(j * 3 + p) * 10 + y];
            for (int p = 0; p < 9; p++) acc += c2 * ddni[p] * ddN[(j * 9 + p) * 10 + y];
            atomicAdd(mat + (size_t)r * n + topo[8 * v + j] * 10 + y, acc);
        }
}

__global__ void build_pin_global_kernel(double stiff, const int *__restrict__ vidx, int n_pin, const int *__restrict__ topo,
                                        const double *__restrict__ Nx, int n, double *mat) {  // cuda_utils.py:58-81
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_pin * 64) return;
    const int v = vidx[id / 64], i = (id % 64) / 8, j = id % 8;
    const int ki = topo[8 * v + i], kj = topo[8 * v + j];
    const double *Ni = Nx + ((size_t)v * 8 + i) * 10, *Nj = Nx + ((size_t)v * 8 + j) * 10;
    for (int x = 0; x < 10; x++)
        for (int y = 0; y < 10; y++) atomicAdd(mat + (size_t)(ki * 10 + x) * n + kj * 10 + y, stiff * Ni[x] * Nj[y]);
}

__global__ void collect_gravity_kernel(double dx, const int *__restrict__ topo, const double *__restrict__ Nx, double g0,
                                       double g1, double g2, const double *__restrict__ rho, int n_ip, double *rhs) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;  // cuda_utils.py:262-279, one thread per (IP, corner, slot)
    if (id >= n_ip * 80) return;
    const int v = id / 80, i = (id % 80) / 10, x = id % 10;
    const double m = rho[v] * dx * dx * dx * Nx[(size_t)v * 80 + i * 10 + x];
    double *o = rhs + 3 * (size_t)(topo[8 * v + i] * 10 + x);
    atomicAdd(o, m * g0); atomicAdd(o + 1, m * g1); atomicAdd(o + 2, m * g2);
}

// ---------------------------------------------------------------- local step
// 3x3 SVD by cyclic Jacobi on F^T F; U,V proper rotations, the sign rides on sigma_2 (the wp.svd3 convention).
__device__ void svd3x3(const double *F, double *U, double *sg, double *V) {
    double S[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) S[i * 3 + j] = F[i] * F[j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j];
    double Q[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    // (warm-starting Q from the previous local-global iteration was measured: no gain once the convergence test below is sane,
    //  and it costs the exact rest-state fixed point — with F = I the eigenvectors are arbitrary, so a carried Q changes U V^T by a ulp)
    // the three rotations of a sweep are spelled out with constant indices so that S and Q stay in registers
    // (a (p, q) loop indexes them dynamically -> local memory on the serial critical path of every step)
#define PN_JACOBI(p, q)                                                                                                   \
    {                                                                                                                     \
        const double apq = S[p * 3 + q];                                                                                  \
        if (apq != 0.0) {                                                                                                 \
            const double th = (S[q * 3 + q] - S[p * 3 + p]) / (2.0 * apq);                                                \
            const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));                                   \
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;                                                          \
            _Pragma("unroll") for (int k = 0; k < 3; k++) { const double a = S[k * 3 + p], b = S[k * 3 + q]; S[k * 3 + p] = c * a - s * b; S[k * 3 + q] = s * a + c * b; } \
            _Pragma("unroll") for (int k = 0; k < 3; k++) { const double a = S[p * 3 + k], b = S[q * 3 + k]; S[p * 3 + k] = c * a - s * b; S[q * 3 + k] = s * a + c * b; } \
            _Pragma("unroll") for (int k = 0; k < 3; k++) { const double a = Q[k * 3 + p], b = Q[k * 3 + q]; Q[k * 3 + p] = c * a - s * b; Q[k * 3 + q] = s * a + c * b; } \
        }                                                                                                                 \
    }
    for (int sweep = 0; sweep < 12; sweep++) {
        const double off = fabs(S[1]) + fabs(S[2]) + fabs(S[5]);
        // converged when the off-diagonal mass is below fp64 resolution of the diagonal (1e-17 relative: a further sweep cannot
        // change U V^T or the singular values by a bit that matters; a 1e-30 target is unreachable and costs all 12 sweeps)
        if (off <= 1e-17 * (fabs(S[0]) + fabs(S[4]) + fabs(S[8]))) break;
        PN_JACOBI(0, 1)
        PN_JACOBI(0, 2)
        PN_JACOBI(1, 2)
    }
#undef PN_JACOBI
    double lam[3] = {S[0], S[4], S[8]};
    int o0 = 0, o1 = 1, o2 = 2;
    if (lam[o1] > lam[o0]) { const int t = o0; o0 = o1; o1 = t; }
    if (lam[o2] > lam[o0]) { const int t = o0; o0 = o2; o2 = t; }
    if (lam[o2] > lam[o1]) { const int t = o1; o1 = o2; o2 = t; }
    const int ord[3] = {o0, o1, o2};
    for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) V[k * 3 + c] = Q[k * 3 + ord[c]];
    const double det = V[0] * (V[4] * V[8] - V[5] * V[7]) - V[1] * (V[3] * V[8] - V[5] * V[6]) + V[2] * (V[3] * V[7] - V[4] * V[6]);
    if (det < 0) for (int k = 0; k < 3; k++) V[k * 3 + 2] = -V[k * 3 + 2];
    double B[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) B[r * 3 + c] = F[r * 3] * V[c] + F[r * 3 + 1] * V[3 + c] + F[r * 3 + 2] * V[6 + c];
    double u0[3], u1[3], u2[3];
    double n0 = sqrt(B[0] * B[0] + B[3] * B[3] + B[6] * B[6]);
    if (n0 > 0) { u0[0] = B[0] / n0; u0[1] = B[3] / n0; u0[2] = B[6] / n0; } else { u0[0] = 1; u0[1] = 0; u0[2] = 0; }
    const double d01 = u0[0] * B[1] + u0[1] * B[4] + u0[2] * B[7];
    u1[0] = B[1] - d01 * u0[0]; u1[1] = B[4] - d01 * u0[1]; u1[2] = B[7] - d01 * u0[2];
    double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    if (n1 > 1e-300) { u1[0] /= n1; u1[1] /= n1; u1[2] /= n1; }
    else {
        const int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
        double a[3] = {0, 0, 0};
        a[m] = 1;
        const double d = u0[m];
        u1[0] = a[0] - d * u0[0]; u1[1] = a[1] - d * u0[1]; u1[2] = a[2] - d * u0[2];
        n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        u1[0] /= n1; u1[1] /= n1; u1[2] /= n1;
    }
    u2[0] = u0[1] * u1[2] - u0[2] * u1[1]; u2[1] = u0[2] * u1[0] - u0[0] * u1[2]; u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
    sg[0] = lam[o0] > 0 ? sqrt(lam[o0]) : 0.0;
    sg[1] = lam[o1] > 0 ? sqrt(lam[o1]) : 0.0;
    sg[2] = u2[0] * B[2] + u2[1] * B[5] + u2[2] * B[8];
    for (int r = 0; r < 3; r++) { U[r * 3] = u0[r]; U[r * 3 + 1] = u1[r]; U[r * 3 + 2] = u2[r]; }
}

__device__ __forceinline__ void volume_project(const double *sig, double *out) {  // func_utils.py:21-40
    double D0 = 0, D1 = 0, D2 = 0;
    for (int it = 0; it < 3; it++) {
        const double a = sig[0] + D0, b = sig[1] + D1, c = sig[2] + D2;
        const double C = a * b * c - 1.0;
        const double e0 = b * c, e1 = a * c, e2 = a * b;
        const double coef = ((e0 * D0 + e1 * D1 + e2 * D2) - C) / (e0 * e0 + e1 * e1 + e2 * e2);
        D0 = coef * e0; D1 = coef * e1; D2 = coef * e2;
    }
    out[0] = sig[0] + D0; out[1] = sig[1] + D1; out[2] = sig[2] + D2;
}

// calc_elastic (cuda_utils.py:83-121): 8 lanes per IP (one per kernel corner), 4 IPs per warp.  Each lane folds its
// corner's 10 DOF vectors against the corner's contiguous [3][10] gradient block, F is reduced over the 8 lanes with
// shuffles, and the 4 group leaders run the 3x3 SVD + projections side by side.
// stress[v] = dx^3 (mu R + lam V)  (row-major 3x3)
__global__ void __launch_bounds__(128) ip_stress_kernel(double dx3, const int *__restrict__ topo, const double *__restrict__ mu,
                                                        const double *__restrict__ lam, const double *__restrict__ dNx,
                                                        const double *__restrict__ dof, int n_ip,
                                                        double *__restrict__ stress) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = t >> 3, i = t & 7;
    const bool live = v < n_ip;
    double F[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (live) {
        const double *dN = dNx + (size_t)v * 240 + i * 30;   // [c][x]
        const double *d = dof + 3 * (size_t)(topo[8 * v + i] * 10);
#pragma unroll
        for (int x = 0; x < 10; x++) {
            const double g0 = dN[x], g1 = dN[10 + x], g2 = dN[20 + x];
            const double d0 = d[3 * x], d1 = d[3 * x + 1], d2 = d[3 * x + 2];
            F[0] += d0 * g0; F[1] += d0 * g1; F[2] += d0 * g2;
            F[3] += d1 * g0; F[4] += d1 * g1; F[5] += d1 * g2;
            F[6] += d2 * g0; F[7] += d2 * g1; F[8] += d2 * g2;
        }
    }
#pragma unroll
    for (int k = 0; k < 9; k++)
        for (int o = 4; o > 0; o >>= 1) F[k] += __shfl_xor_sync(kFull, F[k], o);
    if (live && i == 0) {
        double U[9], sg[3], V[9], sp[3];
        svd3x3(F, U, sg, V);
        volume_project(sg, sp);
        const double m = mu[v], l = lam[v];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                double a = 0, b = 0;
                for (int k = 0; k < 3; k++) { a += U[r * 3 + k] * V[c * 3 + k]; b += U[r * 3 + k] * sp[k] * V[c * 3 + k]; }
                stress[(size_t)v * 9 + r * 3 + c] = dx3 * (m * a + l * b);
            }
    }
}

// collect_rhs_IP as a gather (cuda_utils.py:124-151 semantics, :153-188 structure), two deterministic passes:
//  1. rhs_partial_kernel: CTA (k, s) takes the s-th 128-entry slice of kernel k's (ip,corner) adjacency, one entry per
//     thread: reads the IP's 3x3 stress and the corner's contiguous [3][10] gradient block, forms all 10 slots x 3
//     components, and reduces them over the CTA with a fixed shuffle + shared-memory tree -> partial[k][s][30];
//  2. rhs_final_kernel: sums the slices in order and applies `momentum + rhs - rhs_rest` (solver.py:599).
// No fp64 atomics anywhere: steps are bit-reproducible.
__global__ void __launch_bounds__(128) rhs_partial_kernel(const int *__restrict__ adj_bgn, const int *__restrict__ adj,
                                                          const double *__restrict__ stress, const double *__restrict__ dNx,
                                                          int slices, double *__restrict__ partial) {
    __shared__ double part[4][30];
    const int k = blockIdx.x, sl = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int e = adj_bgn[k] + sl * 128 + threadIdx.x;
    double acc[30];
#pragma unroll
    for (int i = 0; i < 30; i++) acc[i] = 0.0;
    if (e < adj_bgn[k + 1]) {
        const int code = adj[e], v = code >> 3, i = code & 7;
        const double *S = stress + (size_t)v * 9;
        const double *dN = dNx + (size_t)v * 240 + i * 30;   // [c][x], 30 contiguous doubles
        const double s0 = S[0], s1 = S[1], s2 = S[2], s3 = S[3], s4 = S[4], s5 = S[5], s6 = S[6], s7 = S[7], s8 = S[8];
#pragma unroll
        for (int x = 0; x < 10; x++) {
            const double g0 = dN[x], g1 = dN[10 + x], g2 = dN[20 + x];
            acc[3 * x] = s0 * g0 + s1 * g1 + s2 * g2;
            acc[3 * x + 1] = s3 * g0 + s4 * g1 + s5 * g2;
            acc[3 * x + 2] = s6 * g0 + s7 * g1 + s8 * g2;
        }
    }
#pragma unroll
    for (int i = 0; i < 30; i++) {
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(kFull, acc[i], o);
        if (lane == 0) part[wid][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 30)
        partial[((size_t)k * slices + sl) * 30 + threadIdx.x] =
            ((part[0][threadIdx.x] + part[1][threadIdx.x]) + part[2][threadIdx.x]) + part[3][threadIdx.x];
}

__global__ void __launch_bounds__(256) rhs_final_kernel(const double *__restrict__ partial, int n_k, int slices,
                                                        const double *__restrict__ base_add, const double *__restrict__ base_sub,
                                                        double *__restrict__ out) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;    // (k*10 + x)*3 + r
    if (id >= n_k * 30) return;
    const int k = id / 30, i = id % 30;
    double sum = 0.0;
    for (int s = 0; s < slices; s++) sum += partial[((size_t)k * slices + s) * 30 + i];
    out[id] = base_add ? (base_add[id] + sum) - base_sub[id] : sum;
}

// ---------------------------------------------------------------- global step
// y = alpha_add[row] + Mat x  (optional add vectors).  One warp per row, 128-bit row loads.
__global__ void __launch_bounds__(256) matvec3_kernel(const double *__restrict__ mat, const double *__restrict__ x, int n,
                                                      const double *__restrict__ add0, const double *__restrict__ add1,
                                                      const double *__restrict__ add2, double *__restrict__ y) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    const double *m = mat + (size_t)row * n;
    double a0 = 0, a1 = 0, a2 = 0;
    if ((n & 1) == 0) {
        const double2 *m2 = reinterpret_cast<const double2 *>(m);
        for (int j = lane; j < n / 2; j += 32) {
            const double2 w = __ldg(m2 + j);
            const double *xa = x + 6 * (size_t)j;
            a0 += w.x * xa[0] + w.y * xa[3]; a1 += w.x * xa[1] + w.y * xa[4]; a2 += w.x * xa[2] + w.y * xa[5];
        }
    } else {
        for (int j = lane; j < n; j += 32) {
            const double w = __ldg(m + j);
            a0 += w * x[3 * j]; a1 += w * x[3 * j + 1]; a2 += w * x[3 * j + 2];
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(kFull, a0, o); a1 += __shfl_xor_sync(kFull, a1, o); a2 += __shfl_xor_sync(kFull, a2, o);
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double v = c == 0 ? a0 : (c == 1 ? a1 : a2);
            if (add0) v += add0[3 * row + c];
            if (add1) v += add1[3 * row + c];
            if (add2) v += add2[3 * row + c];
            y[3 * row + c] = v;
        }
    }
}

// Small systems (n <= 512): the vector a mat-vec multiplies is rebuilt by every CTA in shared memory instead of being produced
// by a launch of its own — 3 launches per local-global iteration instead of 4, and no launch at all for dof_tilde / the final
// velocity.  Row arithmetic is matvec3_kernel's (same lane partition, same shuffle tree): identical bits.
//   MODE 0: x = dof + dt vel (solver.py:575); y = mom = M x + dof_f + rhs_gravity (576); CTA 0 also saves dof_last (597)
//   MODE 1: x = mom + sum_s partial[.][s] - rhs_rest (599); y = dof = dof_rest + Ainv x (600-601);
//           if vel_out: vel = (dof - dof_last) / dt * 0.998 for the same rows (602)
template <int MODE>
__global__ void __launch_bounds__(256) fused_matvec3_kernel(const double *__restrict__ mat, int n, const double *__restrict__ v0,
                                                            const double *__restrict__ v1, const double *__restrict__ v2, int slices, double dt,
                                                            const double *__restrict__ add0, const double *__restrict__ add1,
                                                            double *__restrict__ y, double *__restrict__ last, double *__restrict__ vel_out) {
    extern __shared__ __align__(16) double xs[];
    const int n3 = 3 * n;
    for (int id = threadIdx.x; id < n3; id += blockDim.x) {
        if (MODE == 0) {
            const double d = v0[id];
            xs[id] = d + dt * v1[id];
            if (blockIdx.x == 0) last[id] = d;
        } else {
            const int k = id / 30, i = id % 30;
            double sum = 0.0;
            for (int sl = 0; sl < slices; sl++) sum += v0[((size_t)k * slices + sl) * 30 + i];
            xs[id] = (v1[id] + sum) - v2[id];
        }
    }
    __syncthreads();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    const double *m = mat + (size_t)row * n;
    double a0 = 0, a1 = 0, a2 = 0;
    if ((n & 1) == 0) {
        const double2 *m2 = reinterpret_cast<const double2 *>(m);
        for (int j = lane; j < n / 2; j += 32) {
            const double2 w = __ldg(m2 + j);
            const double *xa = xs + 6 * (size_t)j;
            a0 += w.x * xa[0] + w.y * xa[3]; a1 += w.x * xa[1] + w.y * xa[4]; a2 += w.x * xa[2] + w.y * xa[5];
        }
    } else {
        for (int j = lane; j < n; j += 32) {
            const double w = __ldg(m + j);
            a0 += w * xs[3 * j]; a1 += w * xs[3 * j + 1]; a2 += w * xs[3 * j + 2];
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(kFull, a0, o); a1 += __shfl_xor_sync(kFull, a1, o); a2 += __shfl_xor_sync(kFull, a2, o);
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double v = c == 0 ? a0 : (c == 1 ? a1 : a2);
            if (add0) v += add0[3 * row + c];
            if (add1) v += add1[3 * row + c];
            y[3 * row + c] = v;
            if (MODE == 1 && vel_out) vel_out[3 * row + c] = (v - last[3 * row + c]) / dt * 0.998;
        }
    }
}

__global__ void axpy_tilde_kernel(const double *__restrict__ dof, const double *__restrict__ vel, double dt, int n3,
                                  double *__restrict__ tilde, double *__restrict__ last) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    tilde[i] = dof[i] + dt * vel[i];   // solver.py:575
    last[i] = dof[i];                  // solver.py:597
}
__global__ void finish_step_kernel(const double *__restrict__ dof, const double *__restrict__ last, double dt, int n3,
                                   double *__restrict__ vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) vel[i] = (dof[i] - last[i]) / dt * 0.998;  // solver.py:602
}

// ---------------------------------------------------------------- the whole step as ONE thread-block-cluster kernel
// stepforward (solver.py:574-602) is a strictly sequential chain — per local-global iteration: per-IP F + SVD, rhs gather,
// one dense mat-vec — over a few thousand IPs and a few hundred DOFs.  As 1 + 3*iters launches it is latency-bound, and on the
// GPU that also renders every one of those launches queues behind persistent render CTAs (0.21 ms alone, 0.5 ms beside the
// renderer).  Here ONE cluster of 16 (or 8) CTAs — one GPC — runs the entire step:
//   * every CTA owns a contiguous run of IPs and a contiguous run of DOF rows for the whole step;
//   * the IPs' gradient blocks dNx (the only large operand, 1920 B / IP, constant) are streamed through a per-warp shared-memory
//     stage by TMA bulk copies (cp.async.bulk on an mbarrier), the next block being fetched while the current one is used —
//     the block for the first round of iteration i+1 is already in flight during the SVDs of iteration i;
//   * the rhs is gathered IP-centrically: each CTA sums the contributions of ITS IPs per kernel in a fixed order (a CTA-local
//     kernel -> (ip, corner) CSR built once in shared memory), the C partial vectors are added in rank order: deterministic,
//     no atomics, and the stresses never leave shared memory;
//   * the CTA's rows of the pre-inverted matrix stay in shared memory for all iterations when they fit (n <= ~400);
//   * two hardware cluster barriers (barrier.cluster, release / acquire) per iteration; what crosses CTAs goes through L2
//     (plain stores, ld.global.cg after the barrier).
// Same algorithm and per-IP / per-row arithmetic as the kernels above; only the rhs summation order differs (fp64 round-off).
struct StepClusterArgs {
    int n_ip, n_k, n, iters;
    double dt, dx3;
    const int *topo; const double *mu, *lam, *dNx;
    const double *Ainv, *M, *dof_rest, *dof_f, *rhs_rest, *rhs_gravity;
    double *dof, *dof_vel;
    double *mom, *partialC;                    // scratch in global memory (L2): [n3], [C][n3]
    long long *prof;                           // [8] cycles of CTA 0 per phase, summed over the iterations (diagnostics)
    int ainv_in_smem;
};
constexpr int kStepThreads = 512;
constexpr int kStepWarps = kStepThreads / 32;
constexpr int kGatherUnroll = 4;           // entries in flight per lane group in the rhs gather

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init1(uint64_t *bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar))); }
__device__ __forceinline__ void mbar_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {   // TMA 1-D bulk copy, 16-byte multiples
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// one row of y = Mat x (+ add0 + add1): matvec3_kernel's lane partition and shuffle tree; `m` may point to shared or global memory
__device__ __forceinline__ void cluster_row(const double *m, const double *x_s, int n, int lane, double &a0, double &a1, double &a2) {
    a0 = a1 = a2 = 0;
    if ((n & 1) == 0) {
        const double2 *m2 = reinterpret_cast<const double2 *>(m);
        const int h = n / 2;
        for (int j0 = lane; j0 < h; j0 += 128) {
            double2 w[4];
#pragma unroll
            for (int u = 0; u < 4; u++) w[u] = (j0 + 32 * u < h) ? m2[j0 + 32 * u] : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (j0 + 32 * u < h) {
                    const double *xa = x_s + 6 * (size_t)(j0 + 32 * u);
                    a0 += w[u].x * xa[0] + w[u].y * xa[3]; a1 += w[u].x * xa[1] + w[u].y * xa[4]; a2 += w[u].x * xa[2] + w[u].y * xa[5];
                }
            }
        }
    } else {
        for (int j = lane; j < n; j += 32) {
            const double w = m[j];
            a0 += w * x_s[3 * j]; a1 += w * x_s[3 * j + 1]; a2 += w * x_s[3 * j + 2];
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(kFull, a0, o); a1 += __shfl_xor_sync(kFull, a1, o); a2 += __shfl_xor_sync(kFull, a2, o);
    }
}

struct StepSmemLayout { size_t dof, vec, last, F, S, ainv, stage, bars, off, ent, total; };
__host__ __device__ inline StepSmemLayout step_smem_layout(int n_ip, int n, int n_k, int C, int ainv_in_smem) {
    const int ips = (((n_ip + C - 1) / C) + 3) & ~3, rows = (n + C - 1) / C;
    StepSmemLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t r = o; o += (bytes + 15) & ~size_t(15); return r; };
    L.dof = take(8 * 3 * (size_t)n); L.vec = take(8 * 3 * (size_t)n); L.last = take(8 * 3 * (size_t)rows);
    L.F = take(8 * 9 * (size_t)ips); L.S = take(8 * 9 * (size_t)ips);
    L.ainv = take(ainv_in_smem ? 8 * (size_t)rows * n : 0);
    L.stage = take(8 * 960 * (size_t)kStepWarps);
    L.bars = take(8 * kStepWarps);
    L.off = take(4 * (2 * (size_t)n_k + 2));
    L.ent = take(2 * 8 * (size_t)ips);
    L.total = o;
    return L;
}

__global__ void __launch_bounds__(kStepThreads, 1) qgmls_step_cluster_kernel(const StepClusterArgs a) {
    extern __shared__ __align__(16) unsigned char step_smem[];
    const int n = a.n, n3 = 3 * a.n;
    const int C = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
    const StepSmemLayout L = step_smem_layout(a.n_ip, n, a.n_k, C, a.ainv_in_smem);
    double *dof_s = reinterpret_cast<double *>(step_smem + L.dof);      // [n3] current DOFs (every CTA holds a copy)
    double *vec_s = reinterpret_cast<double *>(step_smem + L.vec);      // [n3] dof_tilde, then the rhs of each iteration
    double *last_s = reinterpret_cast<double *>(step_smem + L.last);    // [rows,3] DOFs of this CTA's rows at the start of the step
    double *F_s = reinterpret_cast<double *>(step_smem + L.F);          // [ips,9]
    double *S_s = reinterpret_cast<double *>(step_smem + L.S);          // [ips,9] stresses dx^3 (mu R + lam V)
    double *ainv_s = reinterpret_cast<double *>(step_smem + L.ainv);    // [rows,n] this CTA's rows of the pre-inverted matrix
    double *stage_s = reinterpret_cast<double *>(step_smem + L.stage);  // [warps][960] gradient blocks of 4 IPs (7680 B per warp)
    uint64_t *bars = reinterpret_cast<uint64_t *>(step_smem + L.bars);  // one mbarrier per warp
    int *off_s = reinterpret_cast<int *>(step_smem + L.off);            // [n_k+1] CSR offsets, then [n_k+1] counters
    int *cnt_s = off_s + a.n_k + 1;
    unsigned short *ent_s = reinterpret_cast<unsigned short *>(step_smem + L.ent);   // local (ip, corner) codes grouped by kernel
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ips_per_cta = (((a.n_ip + C - 1) / C) + 3) & ~3;
    const int ip0 = min(a.n_ip, rank * ips_per_cta), ip1 = min(a.n_ip, ip0 + ips_per_cta), nloc = ip1 - ip0;
    const int rows_per_cta = (n + C - 1) / C, row0 = min(n, rank * rows_per_cta), row1 = min(n, row0 + rows_per_cta);
    const int n_rounds = (ips_per_cta / 4 + kStepWarps - 1) / kStepWarps;     // gradient blocks (4 IPs) per warp and iteration
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc = clock64();
#define PN_STEP_TICK(k) { const long long t_ = clock64(); pc[k] += t_ - tc; tc = t_; }

    // this warp's gradient block of round r: IPs ip0 + 4 * (r * warps + wid) ..+3; bytes = what exists of it
    auto block_bytes = [&](int r) -> uint32_t {
        const int v = ip0 + 4 * (r * kStepWarps + wid);
        return (r < n_rounds && v < ip1) ? (uint32_t)(min(4, ip1 - v) * 1920) : 0u;
    };
    auto fetch = [&](int r) {                                           // whole warp calls; lane 0 issues
        const uint32_t bytes = block_bytes(r);
        if (lane == 0 && bytes) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage's previous readers (generic proxy) come first
            mbar_expect(&bars[wid], bytes);
            bulk_load(stage_s + (size_t)wid * 960, a.dNx + (size_t)(ip0 + 4 * (r * kStepWarps + wid)) * 240, bytes, &bars[wid]);
        }
    };
    if (lane == 0) mbar_init1(&bars[wid]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t stage_phase = 0;
    int my_rounds = 0;
    for (int r = 0; r < n_rounds; r++) my_rounds += block_bytes(r) != 0;
    fetch(0);

    // ---- dof_tilde = dof + dt vel (solver.py:575), dof_last (597); CTA-local kernel -> (ip, corner) lists; resident matrix rows
    for (int i = tid; i < n3; i += kStepThreads) {
        const double d = a.dof[i];
        dof_s[i] = d;
        vec_s[i] = d + a.dt * a.dof_vel[i];
    }
    for (int i = tid; i < a.n_k + 1; i += kStepThreads) { off_s[i] = 0; cnt_s[i] = 0; }
    if (a.ainv_in_smem) {
        const double2 *src = reinterpret_cast<const double2 *>(a.Ainv + (size_t)row0 * n);
        double2 *dst = reinterpret_cast<double2 *>(ainv_s);
        const int tot2 = (row1 - row0) * n / 2;                         // n is even when ainv_in_smem (checked on the host)
        for (int i = tid; i < tot2; i += kStepThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    for (int i = tid; i < 3 * (row1 - row0); i += kStepThreads) last_s[i] = dof_s[3 * row0 + i];
    for (int e = tid; e < nloc * 8; e += kStepThreads) atomicAdd(&cnt_s[a.topo[8 * (size_t)ip0 + e]], 1);
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < a.n_k; k++) { off_s[k] = run; run += cnt_s[k]; cnt_s[k] = 0; }
        off_s[a.n_k] = run;
    }
    __syncthreads();
    for (int e = tid; e < nloc * 8; e += kStepThreads) {
        const int k = a.topo[8 * (size_t)ip0 + e];
        ent_s[off_s[k] + atomicAdd(&cnt_s[k], 1)] = (unsigned short)e;
    }
    __syncthreads();
    for (int k = tid; k < a.n_k; k += kStepThreads) {                     // ascending (ip, corner) inside a kernel: a fixed summation order
        unsigned short *l = ent_s + off_s[k];
        const int m = off_s[k + 1] - off_s[k];
        for (int i = 1; i < m; i++) {
            const unsigned short v = l[i];
            int j = i - 1;
            while (j >= 0 && l[j] > v) { l[j + 1] = l[j]; j--; }
            l[j + 1] = v;
        }
    }
    // ---- momentum = M/dt^2 dof_tilde + dof_f + rhs_gravity (solver.py:576) for this CTA's rows
    for (int row = row0 + wid; row < row1; row += kStepWarps) {
        double a0, a1, a2;
        cluster_row(a.M + (size_t)row * n, vec_s, n, lane, a0, a1, a2);
        if (lane == 0) {
            a.mom[3 * row] = (a0 + a.dof_f[3 * row]) + a.rhs_gravity[3 * row];
            a.mom[3 * row + 1] = (a1 + a.dof_f[3 * row + 1]) + a.rhs_gravity[3 * row + 1];
            a.mom[3 * row + 2] = (a2 + a.dof_f[3 * row + 2]) + a.rhs_gravity[3 * row + 2];
        }
    }
    __syncthreads();
    PN_STEP_TICK(0)

    for (int it = 0; it < a.iters; it++) {
        // ---- calc_elastic, part 1: F per IP, 8 lanes per IP (ip_stress_kernel's mapping and shuffle tree), gradients from the stage
        for (int r = 0; r < n_rounds; r++) {
            if (block_bytes(r) == 0) continue;                              // warp-uniform
            if (my_rounds > 1 || (it == 0 && r == 0)) { mbar_wait_parity(&bars[wid], stage_phase); stage_phase ^= 1; }
            const int lv = 4 * (r * kStepWarps + wid) + (lane >> 3), v = ip0 + lv, i = lane & 7;
            const bool live = v < ip1;
            double F[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (live) {
                const double *dN = stage_s + (size_t)wid * 960 + (size_t)lane * 30;
                const double *d = dof_s + 3 * (size_t)(a.topo[8 * v + i] * 10);
#pragma unroll
                for (int x = 0; x < 10; x++) {
                    const double g0 = dN[x], g1 = dN[10 + x], g2 = dN[20 + x];
                    const double d0 = d[3 * x], d1 = d[3 * x + 1], d2 = d[3 * x + 2];
                    F[0] += d0 * g0; F[1] += d0 * g1; F[2] += d0 * g2;
                    F[3] += d1 * g0; F[4] += d1 * g1; F[5] += d1 * g2;
                    F[6] += d2 * g0; F[7] += d2 * g1; F[8] += d2 * g2;
                }
            }
#pragma unroll
            for (int k = 0; k < 9; k++)
                for (int o = 4; o > 0; o >>= 1) F[k] += __shfl_xor_sync(kFull, F[k], o);
            if (live && i == 0) {
#pragma unroll
                for (int k = 0; k < 9; k++) F_s[9 * lv + k] = F[k];
            }
            __syncwarp();
            if (my_rounds > 1) {                                             // next block (of the next iteration after the last round)
                int nr = r + 1;
                while (nr < n_rounds && block_bytes(nr) == 0) nr++;
                if (nr < n_rounds) fetch(nr);
                else if (it + 1 < a.iters) fetch(0);
            }
        }
        __syncthreads();
        PN_STEP_TICK(1)
        // ---- part 2: one thread per IP: SVD, R = U V^T, V = U proj(sigma) V^T, stress = dx^3 (mu R + lam V) -> shared memory
        for (int lv = tid; lv < nloc; lv += kStepThreads) {
            const int v = ip0 + lv;
            double F[9], U[9], sg[3], V[9], sp[3];
#pragma unroll
            for (int k = 0; k < 9; k++) F[k] = F_s[9 * lv + k];
            svd3x3(F, U, sg, V);
            volume_project(sg, sp);
            const double m = a.mu[v], l = a.lam[v];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) {
                    double x = 0, y = 0;
                    for (int k = 0; k < 3; k++) { x += U[r * 3 + k] * V[c * 3 + k]; y += U[r * 3 + k] * sp[k] * V[c * 3 + k]; }
                    S_s[9 * lv + r * 3 + c] = a.dx3 * (m * x + l * y);
                }
        }
        __syncthreads();
        PN_STEP_TICK(2)
        // ---- collect_rhs_IP, IP-centric: a warp per kernel; lane group g = lane / 10 walks the kernel's local entries g, g + 3, ...
        // in order, lane x = lane % 10 owns DOF slot x (coalesced 80-byte gradient rows); (g0 + g1) + g2; no atomics
        for (int k = wid; k < a.n_k; k += kStepWarps) {
            const int b = off_s[k], cnt = off_s[k + 1] - b;
            const int g = lane / 10, x = lane - 10 * g;
            double r0 = 0, r1 = 0, r2 = 0;
            if (g < 3) {
                for (int j0 = g; j0 < cnt; j0 += 3 * kGatherUnroll) {
                    double G[kGatherUnroll][3];
                    int lvs[kGatherUnroll];
#pragma unroll
                    for (int u = 0; u < kGatherUnroll; u++) {
                        const int j = j0 + 3 * u;
                        lvs[u] = -1;
                        if (j < cnt) {
                            const int e = ent_s[b + j];
                            lvs[u] = e >> 3;
                            const double *dN = a.dNx + (size_t)(ip0 + (e >> 3)) * 240 + (e & 7) * 30 + x;
                            G[u][0] = __ldg(dN); G[u][1] = __ldg(dN + 10); G[u][2] = __ldg(dN + 20);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kGatherUnroll; u++) {
                        if (lvs[u] >= 0) {
                            const double *S = S_s + 9 * lvs[u];
                            r0 += S[0] * G[u][0] + S[1] * G[u][1] + S[2] * G[u][2];
                            r1 += S[3] * G[u][0] + S[4] * G[u][1] + S[5] * G[u][2];
                            r2 += S[6] * G[u][0] + S[7] * G[u][1] + S[8] * G[u][2];
                        }
                    }
                }
            }
            const double b0 = __shfl_sync(kFull, r0, (lane + 10) & 31), b1 = __shfl_sync(kFull, r1, (lane + 10) & 31), b2 = __shfl_sync(kFull, r2, (lane + 10) & 31);
            const double c0 = __shfl_sync(kFull, r0, (lane + 20) & 31), c1 = __shfl_sync(kFull, r1, (lane + 20) & 31), c2 = __shfl_sync(kFull, r2, (lane + 20) & 31);
            if (lane < 10) {
                double *o = a.partialC + (size_t)rank * n3 + (size_t)k * 30 + 3 * lane;
                o[0] = (r0 + b0) + c0; o[1] = (r1 + b1) + c1; o[2] = (r2 + b2) + c2;
            }
        }
        PN_STEP_TICK(4)
        cluster_sync_all();
        PN_STEP_TICK(3)
        // ---- rhs = momentum + gathered - rhs_rest (solver.py:599): the C partial vectors in rank order, every CTA its own copy
        for (int id = tid; id < n3; id += kStepThreads) {
            double sum = 0.0;
            for (int c = 0; c < C; c++) sum += __ldcg(a.partialC + (size_t)c * n3 + id);
            vec_s[id] = (__ldcg(a.mom + id) + sum) - a.rhs_rest[id];
        }
        __syncthreads();
        // ---- dof = dof_rest + Ainv rhs (solver.py:600-601) for this CTA's rows; the last iteration also writes the velocity (602)
        for (int row = row0 + wid; row < row1; row += kStepWarps) {
            double a0, a1, a2;
            cluster_row(a.ainv_in_smem ? ainv_s + (size_t)(row - row0) * n : a.Ainv + (size_t)row * n, vec_s, n, lane, a0, a1, a2);
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const double v = (c == 0 ? a0 : (c == 1 ? a1 : a2)) + a.dof_rest[3 * row + c];
                    a.dof[3 * row + c] = v;
                    if (it == a.iters - 1) a.dof_vel[3 * row + c] = (v - last_s[3 * (row - row0) + c]) / a.dt * 0.998;
                }
            }
        }
        PN_STEP_TICK(5)
        if (it + 1 < a.iters) {
            cluster_sync_all();
            PN_STEP_TICK(3)
            for (int i = tid; i < n3; i += kStepThreads) dof_s[i] = __ldcg(a.dof + i);
            __syncthreads();
            PN_STEP_TICK(6)
        }
    }
    cluster_sync_all();        // nobody leaves (and frees its shared memory / lets the next launch overwrite scratch) before everyone is done
#undef PN_STEP_TICK
    if (a.prof && rank == 0 && tid == 0)
        for (int k = 0; k < 8; k++) a.prof[k] = pc[k];
}

size_t step_cluster_smem(int n_ip, int n, int n_k, int C, int ainv_in_smem) { return step_smem_layout(n_ip, n, n_k, C, ainv_in_smem).total + 16; }
int step_ainv_fits(int n, int C) { return (n % 2 == 0) && (size_t)((n + C - 1) / C) * n * 8 <= 48 * 1024; }

// ---------------------------------------------------------------- PCG on (A + 1e-3 I) x = b, 3 right-hand sides
// Jacobi-preconditioned CG over the active rows; one cooperative single-block driver per iteration would serialise,
// so each CG iteration is: matvec (n warps) + two fused dot/update kernels with warp-shuffle + atomic reductions.
__global__ void __launch_bounds__(256) pcg_matvec_kernel(const double *__restrict__ A, const unsigned char *__restrict__ act,
                                                         const double *__restrict__ p, int n, double *__restrict__ Ap,
                                                         double *__restrict__ pAp) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    double a0 = 0, a1 = 0, a2 = 0;
    const bool on = act[row / 10] != 0;
    if (on) {
        const double *m = A + (size_t)row * n;
        for (int j = lane; j < n; j += 32) {
            if (!act[j / 10]) continue;
            const double w = __ldg(m + j) + (j == row ? 1e-3 : 0.0);
            a0 += w * p[3 * j]; a1 += w * p[3 * j + 1]; a2 += w * p[3 * j + 2];
        }
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(kFull, a0, o); a1 += __shfl_xor_sync(kFull, a1, o); a2 += __shfl_xor_sync(kFull, a2, o);
        }
    }
    if (lane == 0) {
        Ap[3 * row] = a0; Ap[3 * row + 1] = a1; Ap[3 * row + 2] = a2;
        if (on) {
            atomicAdd(pAp, p[3 * row] * a0); atomicAdd(pAp + 1, p[3 * row + 1] * a1); atomicAdd(pAp + 2, p[3 * row + 2] * a2);
        }
    }
}
// x += alpha p ; r -= alpha Ap ; z = r / diag ; rz_new = sum r z
__global__ void __launch_bounds__(256) pcg_update_kernel(const double *__restrict__ A, const unsigned char *__restrict__ act, int n,
                                                         const double *__restrict__ rz, const double *__restrict__ pAp,
                                                         const double *__restrict__ p, const double *__restrict__ Ap,
                                                         double *__restrict__ x, double *__restrict__ r, double *__restrict__ z,
                                                         double *__restrict__ rz_new) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[3] = {0, 0, 0};
    if (i < n && act[i / 10]) {
        const double dinv = 1.0 / (A[(size_t)i * n + i] + 1e-3);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double alpha = pAp[c] != 0.0 ? rz[c] / pAp[c] : 0.0;
            x[3 * i + c] += alpha * p[3 * i + c];
            const double rr = r[3 * i + c] - alpha * Ap[3 * i + c];
            r[3 * i + c] = rr;
            const double zz = rr * dinv;
            z[3 * i + c] = zz;
            acc[c] = rr * zz;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(kFull, acc[c], o);
        if ((threadIdx.x & 31) == 0 && acc[c] != 0.0) atomicAdd(rz_new + c, acc[c]);
    }
}
// p = z + (rz_new/rz) p ; then rotate scalars: rz <- rz_new, zero rz_new and pAp for the next iteration
__global__ void __launch_bounds__(256) pcg_direction_kernel(const unsigned char *__restrict__ act, int n, const double *__restrict__ z,
                                                            double *__restrict__ p, const double *__restrict__ rz,
                                                            const double *__restrict__ rz_new) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !act[i / 10]) return;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double beta = rz[c] != 0.0 ? rz_new[c] / rz[c] : 0.0;
        p[3 * i + c] = z[3 * i + c] + beta * p[3 * i + c];
    }
}
__global__ void pcg_rotate_kernel(double *rz, double *rz_new, double *pAp) {
    if (threadIdx.x < 3) { rz[threadIdx.x] = rz_new[threadIdx.x]; rz_new[threadIdx.x] = 0.0; pAp[threadIdx.x] = 0.0; }
}
// x = 0 ; r = b (active rows) ; z = r/diag ; p = z ; rz = sum r z
__global__ void __launch_bounds__(256) pcg_init_kernel(const double *__restrict__ A, const unsigned char *__restrict__ act, int n,
                                                       const double *__restrict__ b, double *__restrict__ x, double *__restrict__ r,
                                                       double *__restrict__ z, double *__restrict__ p, double *__restrict__ rz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[3] = {0, 0, 0};
    if (i < n) {
        const bool on = act[i / 10] != 0;
        const double dinv = on ? 1.0 / (A[(size_t)i * n + i] + 1e-3) : 0.0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double rr = on ? b[3 * i + c] : 0.0;
            x[3 * i + c] = 0.0; r[3 * i + c] = rr;
            const double zz = rr * dinv;
            z[3 * i + c] = zz; p[3 * i + c] = zz;
            acc[c] = rr * zz;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(kFull, acc[c], o);
        if ((threadIdx.x & 31) == 0 && acc[c] != 0.0) atomicAdd(rz + c, acc[c]);
    }
}
__global__ void add_rest_kernel(const double *__restrict__ rest, const double *__restrict__ x, int n3, double *__restrict__ dof) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) dof[i] = rest[i] + x[i];
}

// ---------------------------------------------------------------- IP info / positions / force
// update_F_kernel + the fp32 re-layout of get_IP_info (cuda_utils.py:206-233, solver.py:402-424): one warp per IP
__global__ void __launch_bounds__(128) ip_info_kernel(const int *__restrict__ topo, const double *__restrict__ dof,
                                                      const double *__restrict__ Nx, const double *__restrict__ dNx,
                                                      const double *__restrict__ ddNx, int n_ip, float *__restrict__ pos,
                                                      float *__restrict__ Fo, float *__restrict__ dFo) {
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= n_ip) return;
    const double *N = Nx + (size_t)v * 80, *dN = dNx + (size_t)v * 240, *ddN = ddNx + (size_t)v * 720;
    double acc[39];  // pos[3], F[r*3+c] (9), dF[(j*3+r)*3+c] (27)
#pragma unroll
    for (int k = 0; k < 39; k++) acc[k] = 0;
    for (int e = lane; e < 80; e += 32) {
        const int i = e / 10, x = e % 10;
        const double *d = dof + 3 * (size_t)(topo[8 * v + i] * 10 + x);
        const double dv[3] = {d[0], d[1], d[2]};
        const double nv = N[i * 10 + x];
        double g[3], h[9];
#pragma unroll
        for (int c = 0; c < 3; c++) g[c] = dN[(i * 3 + c) * 10 + x];
#pragma unroll
        for (int c = 0; c < 9; c++) h[c] = ddN[(i * 9 + c) * 10 + x];  // h[j*3+c]
#pragma unroll
        for (int r = 0; r < 3; r++) {
            acc[r] += nv * dv[r];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                acc[3 + r * 3 + c] += dv[r] * g[c];
#pragma unroll
                for (int j = 0; j < 3; j++) acc[12 + (j * 3 + r) * 3 + c] += dv[r] * h[j * 3 + c];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 39; k++)
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(kFull, acc[k], o);
    if (lane == 0) {
        for (int r = 0; r < 3; r++) pos[3 * v + r] = (float)acc[r];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Fo[9 * v + a * 3 + b] = (float)acc[3 + b * 3 + a];
        for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) for (int j = 0; j < 3; j++)
            dFo[27 * v + c * 9 + r * 3 + j] = (float)acc[12 + (j * 3 + r) * 3 + c];
    }
}

__global__ void update_pos_kernel(const int *__restrict__ topo, const double *__restrict__ dof, const double *__restrict__ Nx,
                                  int n_pts, double *__restrict__ pos) {  // cuda_utils.py:191-203
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_pts) return;
    double a0 = 0, a1 = 0, a2 = 0;
    for (int i = 0; i < 8; i++) {
        const double *d = dof + 3 * (size_t)(topo[8 * v + i] * 10);
        for (int j = 0; j < 10; j++) {
            const double w = Nx[(size_t)v * 80 + i * 10 + j];
            a0 += w * d[3 * j]; a1 += w * d[3 * j + 1]; a2 += w * d[3 * j + 2];
        }
    }
    pos[3 * v] = a0; pos[3 * v + 1] = a1; pos[3 * v + 2] = a2;
}

__global__ void update_force_kernel(int vid, double f0, double f1, double f2, const int *__restrict__ topo,
                                    const double *__restrict__ Nx, const double *__restrict__ rho, double dx3,
                                    double *__restrict__ dof_f) {  // solver.py:578-588 (dof_f already zeroed)
    const int e = threadIdx.x;
    if (e >= 80) return;
    const int i = e / 10, j = e % 10;
    const double m = rho[vid] * dx3 * Nx[(size_t)vid * 80 + i * 10 + j];
    double *o = dof_f + 3 * (size_t)(topo[8 * vid + i] * 10 + j);
    // kernels of one IP are distinct, so rows are disjoint; atomics keep it safe if two corners alias kernel 0
    atomicAdd(o, m * f0); atomicAdd(o + 1, m * f1); atomicAdd(o + 2, m * f2);
}

}  // namespace

// =============================================================================================== C-ABI
extern "C" int pn_qgmls_shape_functions(double r, const double *pos, const int *topo, const double *kernel_pos, int n,
                                        double *Nx, double *dNx, double *ddNx, int *status, void *stream) {
    PN_REQUIRE(pos && topo && kernel_pos && Nx && status, "null pointer");
    PN_REQUIRE(!ddNx || dNx, "ddNx requires dNx");
    if (n <= 0) return PN_OK;
    cudaStream_t st = PN_STREAM(stream);
    PN_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
    shape_kernel<<<div_up(n, 64), 64, 0, st>>>(r, pos, topo, kernel_pos, n, Nx, dNx, ddNx, status);
    PN_LAUNCH_CHECK("shape_kernel");
    return PN_OK;
}

extern "C" int pn_qgmls_collect_param(const int *pts_ip, const double *mu, const double *lam, const double *mass,
                                      int n_pts, int n_ip, double dx, double *ip_mu, double *ip_lam, double *ip_rho,
                                      void *stream) {
    PN_REQUIRE(pts_ip && mu && lam && mass && ip_mu && ip_lam && ip_rho, "null pointer");
    cudaStream_t st = PN_STREAM(stream);
    PN_CUDA(cudaMemsetAsync(ip_mu, 0, sizeof(double) * n_ip, st));
    PN_CUDA(cudaMemsetAsync(ip_lam, 0, sizeof(double) * n_ip, st));
    PN_CUDA(cudaMemsetAsync(ip_rho, 0, sizeof(double) * n_ip, st));
    collect_param_kernel<<<div_up(n_pts, 256), 256, 0, st>>>(pts_ip, mu, lam, mass, n_pts, ip_mu, ip_lam, ip_rho);
    finish_param_kernel<<<div_up(n_ip, 256), 256, 0, st>>>(n_ip, dx * dx * dx, ip_mu, ip_lam, ip_rho);
    PN_LAUNCH_CHECK("collect_param");
    return PN_OK;
}

extern "C" int pn_qgmls_build_ip_global(double dx, double dt, const int *topo, const double *mu, const double *lam,
                                        const double *rho, const double *Nx, const double *dNx, const double *ddNx,
                                        int n_ip, int n, double *mat, void *stream) {
    PN_REQUIRE(topo && rho && Nx && dNx && ddNx && mat, "null pointer");
    if (n_ip <= 0) return PN_OK;
    build_ip_global_kernel<<<div_up(n_ip * 80, 128), 128, 0, PN_STREAM(stream)>>>(dx, dt, topo, mu, lam, rho, Nx, dNx, ddNx, n_ip, n, mat);
    PN_LAUNCH_CHECK("build_ip_global_kernel");
    return PN_OK;
}

extern "C" int pn_qgmls_build_pin_global(double stiff, const int *vidx, int n_pin, const int *topo, const double *Nx,
                                         int n, double *mat, void *stream) {
    if (n_pin <= 0) return PN_OK;
    PN_REQUIRE(vidx && topo && Nx && mat, "null pointer");
    build_pin_global_kernel<<<div_up(n_pin * 64, 128), 128, 0, PN_STREAM(stream)>>>(stiff, vidx, n_pin, topo, Nx, n, mat);
    PN_LAUNCH_CHECK("build_pin_global_kernel");
    return PN_OK;
}

extern "C" int pn_qgmls_collect_gravity(double dx, const int *topo, const double *Nx, const double *g, const double *rho,
                                        int n_ip, double *rhs, void *stream) {
    PN_REQUIRE(topo && Nx && g && rho && rhs, "null pointer");
    if (n_ip <= 0) return PN_OK;
    collect_gravity_kernel<<<div_up(n_ip * 80, 256), 256, 0, PN_STREAM(stream)>>>(dx, topo, Nx, g[0], g[1], g[2], rho, n_ip, rhs);
    PN_LAUNCH_CHECK("collect_gravity_kernel");
    return PN_OK;
}

extern "C" int pn_qgmls_build_rhs(double dx, const int *topo, const double *mu, const double *lam, const double *dNx,
                                  const double *dof, int n_ip, int n_k, const int *adj_bgn, const int *adj, int slices,
                                  double *ip_stress, double *partial, double *rhs, void *stream) {
    PN_REQUIRE(topo && mu && lam && dNx && dof && adj_bgn && adj && ip_stress && partial && rhs, "null pointer");
    PN_REQUIRE(slices >= 1, "adjacency slices must be >= 1");
    cudaStream_t st = PN_STREAM(stream);
    ip_stress_kernel<<<div_up(n_ip * 8, 128), 128, 0, st>>>(dx * dx * dx, topo, mu, lam, dNx, dof, n_ip, ip_stress);
    rhs_partial_kernel<<<dim3(n_k, slices), 128, 0, st>>>(adj_bgn, adj, ip_stress, dNx, slices, partial);
    rhs_final_kernel<<<div_up(n_k * 30, 256), 256, 0, st>>>(partial, n_k, slices, nullptr, nullptr, rhs);
    PN_LAUNCH_CHECK("build_rhs");
    return PN_OK;
}

extern "C" int pn_qgmls_matvec3(const double *mat, const double *x, int n, double *y, void *stream) {
    PN_REQUIRE(mat && x && y, "null pointer");
    if (n <= 0) return PN_OK;
    matvec3_kernel<<<div_up(n * 32, 256), 256, 0, PN_STREAM(stream)>>>(mat, x, n, nullptr, nullptr, nullptr, y);
    PN_LAUNCH_CHECK("matvec3_kernel");
    return PN_OK;
}

extern "C" uint64_t pn_qgmls_step_scratch_doubles(int n_ip, int n_k, int adj_slices) {
    // stress | tilde last mom rhs x | pcg | partial | 8 phase cycle counters of the cluster kernel | its 16 per-CTA partial rhs vectors
    return 9ull * n_ip + 9ull * 30 * n_k + 16 + 30ull * n_k * (adj_slices > 0 ? adj_slices : 1) + 8 + 16ull * 30 * n_k;
}

// Cluster size for the one-kernel step: 16 CTAs (non-portable size, one GPC) if the device can co-schedule them, else 8;
// 0 = use the multi-kernel path (big systems: the dense mat-vec wants the whole GPU's bandwidth, not one GPC's).
static int g_step_force_multi = 1;   // measured on B200: the cluster kernel is slower (469 vs 219 us, chair body) -> opt-in
extern "C" int pn_qgmls_step_mode(int force_multi_kernel) { g_step_force_multi = force_multi_kernel; return PN_OK; }
static int step_cluster_size(int n_ip, int n, int n_k) {
    static int cached_key_n = -1, cached_key_ip = -1, cached = 0;
    if (n > 1280 || n_ip > 16 * 1024) return 0;              // matrix > 13 MB (x (1 + iters) reads per step) / 16-bit local entry codes
    if (cached_key_n == n && cached_key_ip == n_ip) return cached;
    int best = 0;
    for (int C : {16, 8}) {
        const size_t smem = step_cluster_smem(n_ip, n, n_k, C, step_ainv_fits(n, C));
        if (smem > 220 * 1024) continue;
        if (cudaFuncSetAttribute(qgmls_step_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); continue; }
        if (C > 8 && cudaFuncSetAttribute(qgmls_step_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); continue; }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(C); cfg.blockDim = dim3(kStepThreads); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, qgmls_step_cluster_kernel, &cfg) == cudaSuccess && nclusters >= 1) { best = C; break; }
        cudaGetLastError();
    }
    cached_key_n = n; cached_key_ip = n_ip; cached = best;
    return best;
}

extern "C" int pn_qgmls_step_launches(int n_ip, int n_k, int iters, int solver, int pcg_iters) {
    if (solver == 0 && !g_step_force_multi && step_cluster_size(n_ip, 10 * n_k, n_k) > 0) return 1;
    if (solver == 0 && 10 * n_k <= 512) return 1 + 3 * iters;
    return 3 + iters * (solver == 0 ? 4 : 6 + 4 * pcg_iters);
}

extern "C" int pn_qgmls_step(const pn_qgmls_step_t *s, int solver, void *stream) {
    PN_REQUIRE(s && s->topo && s->mu && s->lam && s->dNx && s->adj_bgn && s->adj && s->M && s->dof_rest && s->dof_f &&
                   s->rhs_rest && s->rhs_gravity && s->dof && s->dof_vel && s->scratch, "null pointer");
    PN_REQUIRE(solver == 0 || solver == 1, "solver must be 0 (dense inverse) or 1 (PCG)");
    PN_REQUIRE(s->adj_slices >= 1, "adj_slices must be >= 1");
    PN_REQUIRE(solver == 1 ? (s->A && s->active) : (s->Ainv != nullptr), "missing system matrix for the chosen solver");
    cudaStream_t st = PN_STREAM(stream);
    const int n = 10 * s->n_k, n3 = 3 * n;
    double *stress = s->scratch;
    double *tilde = stress + 9 * (size_t)s->n_ip;
    double *last = tilde + n3, *mom = last + n3, *rhs = mom + n3, *x = rhs + n3;
    double *pcg = x + n3;  // r, z, p, Ap (4*n3) + 9 scalars, only touched by the PCG path
    double *partial = pcg + 4 * (size_t)n3 + 16;
    const double dx3 = s->dx * s->dx * s->dx;
    const int eb = div_up(n3, 256);
    if (solver == 0 && !g_step_force_multi) {
        const int C = step_cluster_size(s->n_ip, n, s->n_k);
        if (C > 0) {
            StepClusterArgs a{};
            a.n_ip = s->n_ip; a.n_k = s->n_k; a.n = n; a.iters = s->iters; a.dt = s->dt; a.dx3 = dx3;
            a.topo = s->topo; a.mu = s->mu; a.lam = s->lam; a.dNx = s->dNx;
            a.Ainv = s->Ainv; a.M = s->M; a.dof_rest = s->dof_rest; a.dof_f = s->dof_f; a.rhs_rest = s->rhs_rest; a.rhs_gravity = s->rhs_gravity;
            a.dof = s->dof; a.dof_vel = s->dof_vel; a.mom = mom;
            a.prof = reinterpret_cast<long long *>(partial + 30 * (size_t)s->n_k * s->adj_slices);
            a.partialC = partial + 30 * (size_t)s->n_k * s->adj_slices + 8;
            a.ainv_in_smem = step_ainv_fits(n, C);
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(C); cfg.blockDim = dim3(kStepThreads); cfg.dynamicSmemBytes = step_cluster_smem(s->n_ip, n, s->n_k, C, a.ainv_in_smem); cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            PN_CUDA(cudaLaunchKernelEx(&cfg, qgmls_step_cluster_kernel, a));
            return PN_OK;
        }
    }
    const bool fuse = solver == 0 && n <= 512;                              // small systems: 1 + 3*iters launches (see fused_matvec3_kernel; measured slower at n = 1250)
    const size_t xs_bytes = sizeof(double) * (size_t)n3;
    if (fuse) {
        fused_matvec3_kernel<0><<<div_up(n * 32, 256), 256, xs_bytes, st>>>(s->M, n, s->dof, s->dof_vel, nullptr, 0, s->dt, s->dof_f, s->rhs_gravity, mom, last, nullptr);
    } else {
        axpy_tilde_kernel<<<eb, 256, 0, st>>>(s->dof, s->dof_vel, s->dt, n3, tilde, last);
        // momentum = M/dt^2 @ dof_tilde + dof_f + rhs_gravity  (solver.py:576)
        matvec3_kernel<<<div_up(n * 32, 256), 256, 0, st>>>(s->M, tilde, n, s->dof_f, s->rhs_gravity, nullptr, mom);
    }
    for (int it = 0; it < s->iters; it++) {
        ip_stress_kernel<<<div_up(s->n_ip * 8, 128), 128, 0, st>>>(dx3, s->topo, s->mu, s->lam, s->dNx, s->dof, s->n_ip, stress);
        rhs_partial_kernel<<<dim3(s->n_k, s->adj_slices), 128, 0, st>>>(s->adj_bgn, s->adj, stress, s->dNx, s->adj_slices, partial);
        if (fuse) {
            fused_matvec3_kernel<1><<<div_up(n * 32, 256), 256, xs_bytes, st>>>(s->Ainv, n, partial, mom, s->rhs_rest, s->adj_slices, s->dt, s->dof_rest, nullptr, s->dof,
                                                                                 last, it == s->iters - 1 ? s->dof_vel : nullptr);
            continue;
        }
        rhs_final_kernel<<<div_up(n3, 256), 256, 0, st>>>(partial, s->n_k, s->adj_slices, mom, s->rhs_rest, rhs);
        if (solver == 0) {
            // dof = dof_rest + Ainv rhs  (solver.py:600-601)
            matvec3_kernel<<<div_up(n * 32, 256), 256, 0, st>>>(s->Ainv, rhs, n, s->dof_rest, nullptr, nullptr, s->dof);
        } else {
            double *r = pcg, *z = r + n3, *p = z + n3, *Ap = p + n3, *sc = Ap + n3;  // sc: rz[3], rz_new[3], pAp[3]
            PN_CUDA(cudaMemsetAsync(sc, 0, 9 * sizeof(double), st));
            const int nb = div_up(n, 256);
            pcg_init_kernel<<<nb, 256, 0, st>>>(s->A, s->active, n, rhs, x, r, z, p, sc);
            for (int k = 0; k < s->pcg_iters; k++) {
                pcg_matvec_kernel<<<div_up(n * 32, 256), 256, 0, st>>>(s->A, s->active, p, n, Ap, sc + 6);
                pcg_update_kernel<<<nb, 256, 0, st>>>(s->A, s->active, n, sc, sc + 6, p, Ap, x, r, z, sc + 3);
                pcg_direction_kernel<<<nb, 256, 0, st>>>(s->active, n, z, p, sc, sc + 3);
                pcg_rotate_kernel<<<1, 32, 0, st>>>(sc, sc + 3, sc + 6);
            }
            add_rest_kernel<<<eb, 256, 0, st>>>(s->dof_rest, x, n3, s->dof);
        }
    }
    if (fuse) {
        PN_LAUNCH_CHECK("qgmls_step");
        return PN_OK;
    }
    finish_step_kernel<<<eb, 256, 0, st>>>(s->dof, last, s->dt, n3, s->dof_vel);
    PN_LAUNCH_CHECK("qgmls_step");
    return PN_OK;
}

extern "C" int pn_qgmls_ip_info(const int *topo, const double *dof, const double *Nx, const double *dNx,
                                const double *ddNx, int n_ip, float *pos, float *F, float *dF, void *stream) {
    PN_REQUIRE(topo && dof && Nx && dNx && ddNx && pos && F && dF, "null pointer");
    if (n_ip <= 0) return PN_OK;
    ip_info_kernel<<<div_up(n_ip * 32, 128), 128, 0, PN_STREAM(stream)>>>(topo, dof, Nx, dNx, ddNx, n_ip, pos, F, dF);
    PN_LAUNCH_CHECK("ip_info_kernel");
    return PN_OK;
}

extern "C" int pn_qgmls_update_pos(const int *topo, const double *dof, const double *Nx, int n_pts, double *pos,
                                   void *stream) {
    PN_REQUIRE(topo && dof && Nx && pos, "null pointer");
    if (n_pts <= 0) return PN_OK;
    update_pos_kernel<<<div_up(n_pts, 128), 128, 0, PN_STREAM(stream)>>>(topo, dof, Nx, n_pts, pos);
    PN_LAUNCH_CHECK("update_pos_kernel");
    return PN_OK;
}

extern "C" int pn_qgmls_update_force(int vid, const double *f, const int *topo, const double *Nx, const double *rho,
                                     double dx, int n, double *dof_f, void *stream) {
    PN_REQUIRE(topo && Nx && rho && dof_f, "null pointer");
    cudaStream_t st = PN_STREAM(stream);
    PN_CUDA(cudaMemsetAsync(dof_f, 0, sizeof(double) * 3 * (size_t)n, st));
    if (vid < 0) return PN_OK;
    PN_REQUIRE(f, "null force");
    update_force_kernel<<<1, 96, 0, st>>>(vid, f[0], f[1], f[2], topo, Nx, rho, dx * dx * dx, dof_f);
    PN_LAUNCH_CHECK("update_force_kernel");
    return PN_OK;
}
