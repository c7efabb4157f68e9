/* Plain-C caller of libmgb_b200.so - no Python, no torch, no CUDA headers: what a foreign-language host (the Julia
 * `ccall` shim) does.  Builds a fem1d-type problem (2-node broken elements on [-1,1], operator table
 * [u.id; u.dx; s.id], R = blockdiag(R_u Dirichlet, R_s full)), stores every operator the way the reference stores an
 * HPCSparseMatrix row block - 1-based `colptr` over local rows, COMPRESSED `rowval`, `col_indices`
 * (reference src/MultiGridBarrierMPI.jl:216-221) - for TWO row blocks ("ranks"), creates one plan per block with
 * mgb_plan_create_local, assembles through host buffers (mgb_assemble_host) and checks  sum over blocks  of gradient
 * and R'HR against a dense evaluation written here from the formulas of the reference's f2 loop
 * (test/test_map_rows_compare.jl:102-123,165-171).
 *   c_driver            numeric run (needs a B200)
 *   c_driver --symbolic pattern / info only (ctx = NULL): runs anywhere                                            */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mgb_b200.h"

#define E 24           /* elements */
#define N (2 * E)      /* quadrature points = broken nodes */
#define MU (E - 1)     /* u dofs: interior vertices */
#define MS (E + 1)     /* s dofs: all vertices */
#define M (MU + MS)
#define ND 3

#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, mgb_last_error()); return 1; } } while (0)

static double Dm[ND][N][2 * N];  /* dense operators */
static double Rm[2 * N][M];
static double Em[ND][N][M];      /* D_k R */

/* one row block of a dense operator in the reference's HPCSparseMatrix storage (1-based, compressed columns) */
typedef struct { int32_t colptr[N + 1], rowval[4 * N], col_indices[2 * N]; double nzval[4 * N]; mgb_hpc_block blk; } Block;

static void make_block(const double A[N][2 * N], int r0, int r1, Block* b) {
    int used[2 * N], comp[2 * N], nc = 0, nz = 0;
    memset(used, 0, sizeof(used));
    for (int i = r0; i < r1; ++i) for (int j = 0; j < 2 * N; ++j) if (A[i][j] != 0.0) used[j] = 1;
    for (int j = 0; j < 2 * N; ++j) if (used[j]) { comp[j] = nc; b->col_indices[nc++] = j + 1; }
    b->colptr[0] = 1;
    for (int i = r0; i < r1; ++i) {
        for (int j = 0; j < 2 * N; ++j) if (A[i][j] != 0.0) { b->rowval[nz] = comp[j] + 1; b->nzval[nz++] = A[i][j]; }
        b->colptr[i - r0 + 1] = nz + 1;
    }
    b->blk.nrows_local = r1 - r0; b->blk.ncols_compressed = nc; b->blk.ncols_global = 2 * N; b->blk.row0 = r0;
    b->blk.colptr = b->colptr; b->blk.rowval = b->rowval; b->blk.nzval = b->nzval; b->blk.col_indices = b->col_indices;
    b->blk.index_base = 1;
}

int main(int argc, char** argv) {
    const int symbolic = argc > 1 && strcmp(argv[1], "--symbolic") == 0;
    const double h = 2.0 / E, p = 1.0, t = 0.7;
    static double x[N], w[N], c[N][ND], z0[2 * N], s[M];
    /* geometry + operators */
    for (int e = 0; e < E; ++e)
        for (int l = 0; l < 2; ++l) {
            const int i = 2 * e + l;
            x[i] = -1.0 + h * (e + l); w[i] = h / 2;
            Dm[0][i][i] = 1.0;                                   /* u.id */
            Dm[1][i][2 * e] = -1.0 / h; Dm[1][i][2 * e + 1] = 1.0 / h;   /* u.dx */
            Dm[2][i][N + i] = 1.0;                               /* s.id */
            const int v = e + l;                                 /* vertex of this broken node */
            if (v >= 1 && v <= E - 1) Rm[i][v - 1] = 1.0;        /* u: Dirichlet subspace */
            Rm[N + i][MU + v] = 1.0;                             /* s: full subspace */
            z0[i] = x[i]; z0[N + i] = 2.0;                       /* boundary lift g(x) = [x, 2] (the reference's 1-D default shape) */
            c[i][0] = 0.5; c[i][1] = 0.0; c[i][2] = 1.0;         /* f(x) = [0.5, 0, 1] */
        }
    srand(20261018);
    for (int a = 0; a < M; ++a) s[a] = 1e-3 * (2.0 * rand() / RAND_MAX - 1.0);
    for (int k = 0; k < ND; ++k)
        for (int i = 0; i < N; ++i)
            for (int a = 0; a < M; ++a) { double v = 0; for (int j = 0; j < 2 * N; ++j) v += Dm[k][i][j] * Rm[j][a]; Em[k][i][a] = v; }
    /* dense reference: Dz = D (z0 + R s); F = -log(s^2 - q^2) on (q, s) = Dz[1], Dz[2] (p = 1) */
    static double g_ref[M], H_ref[M][M];
    double f0_ref = 0.0;
    for (int i = 0; i < N; ++i) {
        double dz[ND];
        for (int k = 0; k < ND; ++k) {
            double v = 0;
            for (int j = 0; j < 2 * N; ++j) if (Dm[k][i][j] != 0.0) { double zj = z0[j]; for (int a = 0; a < M; ++a) zj += Rm[j][a] * s[a]; v += Dm[k][i][j] * zj; }
            dz[k] = v;
        }
        const double q = dz[1], sv = dz[2], phi = sv * sv - q * q;
        if (!(phi > 0 && sv > 0)) { fprintf(stderr, "driver iterate infeasible\n"); return 1; }
        double y1[ND] = {0, 2 * q / phi, -2 * sv / phi};
        double y2[ND][ND] = {{0}};
        y2[1][1] = 4 * q * q / (phi * phi) + 2 / phi;
        y2[1][2] = y2[2][1] = -2 * q * 2 * sv / (phi * phi);
        y2[2][2] = -2 / phi + 4 * sv * sv / (phi * phi);
        f0_ref += w[i] * (-log(phi));
        for (int k = 0; k < ND; ++k) f0_ref += w[i] * t * c[i][k] * dz[k];
        for (int a = 0; a < M; ++a)
            for (int k = 0; k < ND; ++k) {
                if (Em[k][i][a] == 0.0) continue;
                g_ref[a] += Em[k][i][a] * w[i] * (y1[k] + t * c[i][k]);
                for (int k2 = 0; k2 < ND; ++k2)
                    for (int b = 0; b < M; ++b) H_ref[a][b] += Em[k][i][a] * w[i] * y2[k][k2] * Em[k2][i][b];
            }
    }
    /* R as a replicated 1-based CSR */
    static int32_t rp[2 * N + 1], rc[2 * N]; static double rv[2 * N];
    int nzr = 0; rp[0] = 1;
    for (int j = 0; j < 2 * N; ++j) { for (int a = 0; a < M; ++a) if (Rm[j][a] != 0.0) { rc[nzr] = a + 1; rv[nzr++] = Rm[j][a]; } rp[j + 1] = nzr + 1; }
    mgb_csr R = {2 * N, M, nzr, rp, rc, rv, 1};
    mgb_barrier bar; memset(&bar, 0, sizeof(bar));
    bar.kind = MGB_BARRIER_EUCLIDIAN_POWER; bar.nidx = 2; bar.idx[0] = 1; bar.idx[1] = 2; bar.p = p;

    mgb_ctx* ctx = NULL;
    if (!symbolic) CHECK(mgb_ctx_create(0, NULL, &ctx));
    static double g_sum[M], H_sum[M][M];
    double f0_sum = 0.0, cdz = 0.0;
    const int cut = 2 * (E / 3);   /* two unequal row blocks on an element boundary */
    const int r0s[2] = {0, cut}, r1s[2] = {cut, N};
    for (int rank = 0; rank < 2; ++rank) {
        static Block blocks[ND];
        mgb_hpc_block Dblk[ND];
        const int r0 = r0s[rank], r1 = r1s[rank], nl = r1 - r0;
        for (int k = 0; k < ND; ++k) { make_block(Dm[k], r0, r1, &blocks[k]); Dblk[k] = blocks[k].blk; }
        mgb_plan* plan = NULL;
        CHECK(mgb_plan_create_local(ctx, N, ND, Dblk, &R, 1, x + r0, w + r0, &bar, 0, &plan));
        int64_t info[16];
        CHECK(mgb_plan_info(plan, info, 16));
        if (info[0] != MGB_PATH_ELEMENT || info[1] != nl || info[3] != M || info[6] != 2) { fprintf(stderr, "unexpected plan info\n"); return 1; }
        const int nnz = (int)info[4];
        int32_t* prow = malloc((M + 1) * sizeof(int32_t)); int32_t* pcol = malloc((nnz + 1) * sizeof(int32_t));
        CHECK(mgb_plan_pattern(plan, prow, pcol));
        if (prow[0] != 0 || prow[M] != nnz) { fprintf(stderr, "bad pattern\n"); return 1; }
        if (!symbolic) {
            /* host inputs: local rows of Dz0 = D z0 and of c, column-major nl x ND */
            double* Dz0 = calloc((size_t)nl * ND, sizeof(double)); double* cl = calloc((size_t)nl * ND, sizeof(double));
            for (int k = 0; k < ND; ++k)
                for (int i = 0; i < nl; ++i) {
                    double v = 0; for (int j = 0; j < 2 * N; ++j) v += Dm[k][r0 + i][j] * z0[j];
                    Dz0[k * nl + i] = v; cl[k * nl + i] = c[r0 + i][k];
                }
            double scal[4], *grad = calloc(M, sizeof(double)), *hval = calloc(nnz, sizeof(double));
            CHECK(mgb_assemble_host(plan, s, Dz0, cl, 1, t, MGB_WANT_F0 | MGB_WANT_GRAD | MGB_WANT_HESS, scal, grad, hval, NULL));
            if (scal[1] != 1.0) { fprintf(stderr, "library reports a non-finite iterate\n"); return 1; }
            f0_sum += scal[0]; cdz += scal[2];
            for (int a = 0; a < M; ++a) { g_sum[a] += grad[a]; for (int q = prow[a]; q < prow[a + 1]; ++q) H_sum[a][pcol[q]] += hval[q]; }
            free(Dz0); free(cl); free(grad); free(hval);
        }
        free(prow); free(pcol);
        CHECK(mgb_plan_destroy(plan));
    }
    if (!symbolic) {
        double eg = 0, gn = 0, eh = 0, hn = 0;
        for (int a = 0; a < M; ++a) {
            eg = fmax(eg, fabs(g_sum[a] - g_ref[a])); gn = fmax(gn, fabs(g_ref[a]));
            for (int b = 0; b < M; ++b) { eh = fmax(eh, fabs(H_sum[a][b] - H_ref[a][b])); hn = fmax(hn, fabs(H_ref[a][b])); }
        }
        const double ef = fabs(f0_sum - f0_ref) / fabs(f0_ref);
        printf("f0 %.15g (ref %.15g) rel %.2e | grad rel %.2e | hess rel %.2e\n", f0_sum, f0_ref, ef, eg / gn, eh / hn);
        if (!(ef < 1e-12 && eg / gn < 1e-12 && eh / hn < 1e-12)) { fprintf(stderr, "MISMATCH\n"); return 1; }
        CHECK(mgb_ctx_destroy(ctx));
    }
    printf("C_DRIVER_OK\n");
    return 0;
}
