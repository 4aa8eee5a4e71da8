/*
 * Oracle (TEST INFRASTRUCTURE): sequential marching cubes on the CPU.
 *
 * PARITY UNPINNED against scikit-image 0.18.2 `marching_cubes(method='lewiner')` (ref predict.py:172-177): the
 * Lewiner look-up tables are in an un-vendored third-party package that is not installed here.  This file restates
 * the parts of that implementation that can be stated from its published description -- cell scan order
 * (axis0 -> axis1 -> axis2), corner/edge numbering, "bit i set iff v_i - level > 0", per-edge vertex de-duplication
 * with first-use numbering, inverse-|value| weighted vertex placement with the FLT_EPSILON guard, the face test
 * (asymptotic decider) for ambiguous faces -- and replaces the recited MC33 tilings by their constructive
 * definition: the iso-contour segments of the six cube faces are chained into closed loops, each loop is
 * fan-triangulated.  Unlike the CUDA implementation, nothing is tabulated: segments and loops are traced per cell
 * at run time and vertices are numbered by a literal first-use cache, so the two implementations only share the
 * rules, not code.
 *
 * Build: gcc -O2 -shared -fPIC -o _build/libmc_oracle.so mc_oracle.c -lm   (oracle/Makefile)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float* coords;   /* [nv,3] voxel-unit coordinates (axis0, axis1, axis2), float32 */
    int32_t* faces;  /* [nf,3] */
    float* normals;  /* [nv,3] */
    float* values;   /* [nv]   */
    int64_t nv, nf;
    int status;      /* 0 ok, -4 level outside data range */
} mc_result;

static const int CORNER[8][3] = {/* dx, dy, dz */ {0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                                 {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const int EDGE_CORNER[12][2] = {{0, 1}, {1, 2}, {3, 2}, {0, 3}, {4, 5}, {5, 6}, {7, 6}, {4, 7},
                                       {0, 4}, {1, 5}, {2, 6}, {3, 7}};
/* faces, corners counter-clockwise seen from outside the cube */
static const int FACE_CORNER[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {3, 7, 6, 2}, {0, 4, 7, 3}, {1, 2, 6, 5}};

static int edge_between(int a, int b) {
    for (int e = 0; e < 12; ++e)
        if ((EDGE_CORNER[e][0] == a && EDGE_CORNER[e][1] == b) || (EDGE_CORNER[e][0] == b && EDGE_CORNER[e][1] == a))
            return e;
    return -1;
}

static float vol_at(const float* v, int D, int H, int W, int z, int y, int x) {
    z = z < 0 ? 0 : (z > D - 1 ? D - 1 : z);
    y = y < 0 ? 0 : (y > H - 1 ? H - 1 : y);
    x = x < 0 ? 0 : (x > W - 1 ? W - 1 : x);
    return v[((int64_t)z * H + y) * W + x];
}

typedef struct { float *coords, *normals, *values; int32_t* faces; int64_t nv, nf, cv, cf; } grow;

static void push_vertex(grow* g, const float c[3], const float n[3], float val) {
    if (g->nv == g->cv) {
        g->cv = g->cv ? g->cv * 2 : 1024;
        g->coords = (float*)realloc(g->coords, sizeof(float) * 3 * g->cv);
        g->normals = (float*)realloc(g->normals, sizeof(float) * 3 * g->cv);
        g->values = (float*)realloc(g->values, sizeof(float) * g->cv);
    }
    memcpy(g->coords + 3 * g->nv, c, sizeof(float) * 3);
    memcpy(g->normals + 3 * g->nv, n, sizeof(float) * 3);
    g->values[g->nv] = val;
    g->nv++;
}
static void push_face(grow* g, int a, int b, int c) {
    if (g->nf == g->cf) {
        g->cf = g->cf ? g->cf * 2 : 1024;
        g->faces = (int32_t*)realloc(g->faces, sizeof(int32_t) * 3 * g->cf);
    }
    g->faces[3 * g->nf] = a; g->faces[3 * g->nf + 1] = b; g->faces[3 * g->nf + 2] = c;
    g->nf++;
}

/* float32 voxel-unit coordinates (axis0,axis1,axis2) of the iso-vertex on cube edge e; val = corner values - level */
static int edge_vertex(const double* val, int x, int y, int z, int e, float c[3], double* t, int lo[3], int hi[3]) {
    int c0 = EDGE_CORNER[e][0], c1 = EDGE_CORNER[e][1];
    lo[0] = z + CORNER[c0][2]; lo[1] = y + CORNER[c0][1]; lo[2] = x + CORNER[c0][0];
    hi[0] = z + CORNER[c1][2]; hi[1] = y + CORNER[c1][1]; hi[2] = x + CORNER[c1][0];
    int ax = hi[2] != lo[2] ? 0 : (hi[1] != lo[1] ? 1 : 2);
    double a0 = fabs(val[c0]), a1 = fabs(val[c1]);
    double w0 = 1.0 / ((double)FLT_EPSILON + a0), w1 = 1.0 / ((double)FLT_EPSILON + a1);
    *t = w1 / (w0 + w1);
    c[0] = (float)lo[0]; c[1] = (float)lo[1]; c[2] = (float)lo[2];
    c[2 - ax] = (float)((double)lo[2 - ax] + *t);
    return ax;
}
static void normalise(float gn[3]) {
    volatile float s0 = gn[0] * gn[0], s1 = gn[1] * gn[1], s2 = gn[2] * gn[2];
    volatile float s01 = s0 + s1;
    float nrm = sqrtf(s01 + s2);
    if (nrm > 0.f) for (int q = 0; q < 3; ++q) gn[q] = gn[q] / nrm;
}

/* ---- loop triangulation ------------------------------------------------------------------------------------
 * A diagonal between two loop vertices whose cube edges lie on a common cube face would lie inside that face and
 * collide with the neighbouring cell, so it is forbidden.  Rule (shared with the CUDA table builder): split the chain
 * i..j at the SMALLEST apex k for which both sub-chains are feasible; emit (i,k,j), then the left chain, then the
 * right chain.  If the loop has no feasible triangulation, fan it around an extra centre vertex (mean of the loop's
 * vertices), numbered like any other vertex at its first use. */
static int FACE_EDGE[6][4];
static int cofacial(int e1, int e2) {
    for (int f = 0; f < 6; ++f) {
        int a = 0, b = 0;
        for (int i = 0; i < 4; ++i) { a |= FACE_EDGE[f][i] == e1; b |= FACE_EDGE[f][i] == e2; }
        if (a && b) return 1;
    }
    return 0;
}
static int chord_ok(const int* poly, int n, int a, int b) {
    if (b - a == 1 || (a == 0 && b == n - 1)) return 1;
    return !cofacial(poly[a], poly[b]);
}
static int feasible(const int* poly, int n, int i, int j, int* apex /* [12*12] memo: -2 unknown, -1 impossible */) {
    if (j - i < 2) return 1;
    int* m = &apex[i * 12 + j];
    if (*m != -2) return *m >= 0;
    *m = -1;
    for (int k = i + 1; k < j; ++k)
        if (chord_ok(poly, n, i, k) && chord_ok(poly, n, k, j) && feasible(poly, n, i, k, apex) &&
            feasible(poly, n, k, j, apex)) { *m = k; break; }
    return *m >= 0;
}
static void emit_chain(const int* poly, int i, int j, const int* apex, int (*tris)[3], int* nt) {
    if (j - i < 2) return;
    int k = apex[i * 12 + j];
    tris[*nt][0] = poly[i]; tris[*nt][1] = poly[k]; tris[*nt][2] = poly[j];
    (*nt)++;
    emit_chain(poly, i, k, apex, tris, nt);
    emit_chain(poly, k, j, apex, tris, nt);
}

/* ---- MC33 interior ("tunnel") ambiguity --------------------------------------------------------------------------
 * After the face tests the iso-contour loops cut the cube surface into regions of one sign.  A region with exactly two
 * boundary loops is an ANNULUS; the two regions of the opposite sign on either side of it may or may not be joined through
 * the inside of the cube.  The published MC33 sub-cases with that ambiguity are exactly (enumerated, tests/test_mc33.py):
 *   - the annuli whose two neighbours contain a pair of body-diagonal corners: 4.1, 6.1, 7.4, 10.1, 12.1 (and their
 *     sign-inverted forms) -- 132 (cube code, face decision) configurations;
 *   - case 13.5 (three connected faces around one corner): two nested annuli, each tested with its face-diagonal pairs.
 * Interior test (Chernyaev's criterion, the form Lewiner's test_interior uses for cases 4 and 10): sweep planes
 * perpendicular to an axis; A(t), C(t) = the field along the two cube edges that carry the candidate corners p and q
 * (diagonal in every plane), B(t), D(t) the other two; the corners are joined iff at the extremum t* in (0,1) of
 * A C - B D (a maximum) A(t*), C(t*) have the corners' sign and A C - B D >= FLT_EPSILON there (or B / D has it too).
 * Checked against brute-force connectivity of the sampled trilinear interpolant in tests/test_mc33.py.
 * A joined pair replaces the two caps by a TUBE between the two loops: the cheapest (sum of squared chord lengths at edge
 * midpoints) triangulation of the annulus without in-face diagonals and without extra vertices if one exists (4.1.2:
 * 6 triangles, 6.1.2: 7), else two fans around two extra vertices (7.4.2, 10.1.2, 12.1.2, 13.5.2). */
static int manhattan(int a, int b) {
    return abs(CORNER[a][0] - CORNER[b][0]) + abs(CORNER[a][1] - CORNER[b][1]) + abs(CORNER[a][2] - CORNER[b][2]);
}
static int corner_at(int x, int y, int z) {
    for (int c = 0; c < 8; ++c) if (CORNER[c][0] == x && CORNER[c][1] == y && CORNER[c][2] == z) return c;
    return -1;
}
static int shifted(int c, int axis) {
    int p[3] = {CORNER[c][0], CORNER[c][1], CORNER[c][2]};
    p[axis] ^= 1;
    return corner_at(p[0], p[1], p[2]);
}
/* corners p, q of sign sigma (face- or body-diagonal): joined through the inside of the cube? */
static int interior_joined(const double* val, int p, int q, double sigma) {
    int axis;
    if (manhattan(p, q) == 3) axis = 2;
    else axis = CORNER[p][0] == CORNER[q][0] ? 0 : (CORNER[p][1] == CORNER[q][1] ? 1 : 2);
    int A0 = p, A1 = shifted(p, axis);
    int C0 = CORNER[q][axis] == CORNER[p][axis] ? q : shifted(q, axis), C1 = shifted(C0, axis);
    int B0 = -1, D0 = -1;
    for (int c = 0; c < 8; ++c)
        if (CORNER[c][axis] == CORNER[p][axis] && c != A0 && c != C0) { if (B0 < 0) B0 = c; else D0 = c; }
    int B1 = shifted(B0, axis), D1 = shifted(D0, axis);
    double a0 = sigma * val[A0], a1 = sigma * val[A1], b0 = sigma * val[B0], b1 = sigma * val[B1];
    double c0 = sigma * val[C0], c1 = sigma * val[C1], d0 = sigma * val[D0], d1 = sigma * val[D1];
    double da = a1 - a0, db = b1 - b0, dc = c1 - c0, dd = d1 - d0;
    double a = da * dc - db * dd;
    double b = ((c0 * da + a0 * dc) - d0 * db) - b0 * dd;
    if (!(a < 0.0)) return 0;
    double t = -b / (2.0 * a);
    if (!(t > 0.0 && t < 1.0)) return 0;
    double At = a0 + da * t, Bt = b0 + db * t, Ct = c0 + dc * t, Dt = d0 + dd * t;
    if (At < 0.0 || Ct < 0.0) return 0;
    if (Bt >= 0.0 || Dt >= 0.0) return 1;
    return At * Ct - Bt * Dt >= (double)FLT_EPSILON;
}

/* edge midpoints in units of 1/5040 (every mean over 3..10 of them stays an integer) */
#define MIDU 5040
static void edge_mid(int e, long long m[3]) {
    for (int i = 0; i < 3; ++i) m[i] = (long long)(CORNER[EDGE_CORNER[e][0]][i] + CORNER[EDGE_CORNER[e][1]][i]) * (MIDU / 2);
}
static long long mid_d2(int e1, int e2) {
    long long a[3], b[3];
    edge_mid(e1, a); edge_mid(e2, b);
    return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
}

/* all triangulations of the cut annulus polygon P[0..m) (P may hold a vertex twice: the bridge end points); keeps the
 * best by (cost, emitted vertex sequence) among those that use every directed edge at most once */
typedef struct { const int* P; int m; long long best_cost; int best[16][3]; int have; int cur[16][3]; } tube_enum;
static int te_chord(const tube_enum* te, int a, int b) {
    if (b == a + 1 || (a == 0 && b == te->m - 1)) return 1;
    return te->P[a] != te->P[b] && !cofacial(te->P[a], te->P[b]);
}
static long long te_w(const tube_enum* te, int a, int b) {
    if (b == a + 1 || (a == 0 && b == te->m - 1)) return 0;
    return mid_d2(te->P[a], te->P[b]);
}
static void te_finish(tube_enum* te, int nt, long long cost, long long bridge_cost) {
    cost += bridge_cost;
    /* simplicial: no directed edge twice */
    for (int i = 0; i < nt; ++i)
        for (int k = 0; k < 3; ++k) {
            int u = te->cur[i][k], v = te->cur[i][(k + 1) % 3];
            if (u == v) return;
            for (int j = 0; j < i; ++j)
                for (int l = 0; l < 3; ++l)
                    if (te->cur[j][l] == u && te->cur[j][(l + 1) % 3] == v) return;
        }
    int better = !te->have || cost < te->best_cost;
    if (!better && cost == te->best_cost) {   /* ties: lexicographically smallest emitted vertex sequence */
        const int* a = &te->cur[0][0];
        const int* b = &te->best[0][0];
        for (int i = 0; i < 3 * nt; ++i)
            if (a[i] != b[i]) { better = a[i] < b[i]; break; }
    }
    if (better) { te->have = 1; te->best_cost = cost; memcpy(te->best, te->cur, sizeof(int) * 3 * nt); }
}
/* work list of open intervals (a,b); depth-first: split the first open interval at every admissible apex */
static void te_rec(tube_enum* te, int (*open)[2], int nopen, int nt, long long cost, long long bridge_cost) {
    if (nopen == 0) { te_finish(te, nt, cost, bridge_cost); return; }
    int a = open[nopen - 1][0], b = open[nopen - 1][1];
    if (b - a < 2) { te_rec(te, open, nopen - 1, nt, cost, bridge_cost); return; }
    for (int k = a + 1; k < b; ++k) {
        if (!te_chord(te, a, k) || !te_chord(te, k, b)) continue;
        if (te->P[a] == te->P[k] || te->P[k] == te->P[b] || te->P[a] == te->P[b]) continue;
        te->cur[nt][0] = te->P[a]; te->cur[nt][1] = te->P[k]; te->cur[nt][2] = te->P[b];
        int nxt[16][2];
        memcpy(nxt, open, sizeof(int) * 2 * (nopen - 1));
        /* left chain first (stack: push right, then left) */
        nxt[nopen - 1][0] = k; nxt[nopen - 1][1] = b;
        nxt[nopen][0] = a; nxt[nopen][1] = k;
        te_rec(te, nxt, nopen + 1, nt + 1, cost + te_w(te, a, k) + te_w(te, k, b), bridge_cost);
    }
}
/* tube between directed loops L1 (n1) and L2 (n2).  tris use edge ids 0..11 and centre ids 12 + c0, 12 + c0 + 1;
 * cen_n / cen_loop receive the definitions of the extra vertices.  Returns the triangle count. */
static int tube_triangulate(const int* L1, int n1, const int* L2, int n2, int (*tris)[3], int c0, int* ncen, int* cen_n,
                            int (*cen_loop)[12]) {
    tube_enum te;
    memset(&te, 0, sizeof(te));
    int bestP_nt = 0, have = 0;
    long long best_cost = 0;
    int best[16][3];
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j) {
            if (cofacial(L1[i], L2[j])) continue;
            int P[16], m = 0;
            for (int k = 0; k < n1; ++k) P[m++] = L1[(i + k) % n1];
            P[m++] = L1[i];
            for (int k = 0; k < n2; ++k) P[m++] = L2[(j + k) % n2];
            P[m++] = L2[j];
            te.P = P; te.m = m; te.have = 0;
            int open[16][2] = {{0, m - 1}};
            te_rec(&te, open, 1, 0, 0, mid_d2(L1[i], L2[j]));
            if (te.have && (!have || te.best_cost < best_cost)) {   /* ties: the first bridge (i, j) wins */
                have = 1; best_cost = te.best_cost; bestP_nt = m - 2;
                memcpy(best, te.best, sizeof(int) * 3 * (m - 2));
            }
        }
    if (have) { memcpy(tris, best, sizeof(int) * 3 * bestP_nt); return bestP_nt; }
    /* two fans: centre 1 over the link  L1[a..a+k1], L2[p..p+k2]; centre 2 over the rest */
    long long bc = 0;
    int ba = -1, bk1 = 0, bp = 0, bk2 = 0;
    for (int a = 0; a < n1; ++a)
        for (int k1 = 1; k1 < n1; ++k1)
            for (int p = 0; p < n2; ++p)
                for (int k2 = 1; k2 < n2; ++k2) {
                    int b = (a + k1) % n1, q = (p + k2) % n2;
                    if (cofacial(L1[b], L2[p]) || cofacial(L2[q], L1[a])) continue;
                    int link[16], d1 = 0, rest[16], d2 = 0;
                    for (int t = 0; t <= k1; ++t) link[d1++] = L1[(a + t) % n1];
                    for (int t = 0; t <= k2; ++t) link[d1++] = L2[(p + t) % n2];
                    for (int t = 0; t <= n1 - k1; ++t) rest[d2++] = L1[(b + t) % n1];
                    for (int t = 0; t <= n2 - k2; ++t) rest[d2++] = L2[(q + t) % n2];
                    long long cost = mid_d2(L1[b], L2[p]) + mid_d2(L2[q], L1[a]);
                    for (int pass = 0; pass < 2; ++pass) {
                        const int* v = pass ? rest : link;
                        int d = pass ? d2 : d1;
                        long long c[3] = {0, 0, 0}, m3[3];
                        for (int t = 0; t < d; ++t) { edge_mid(v[t], m3); c[0] += m3[0]; c[1] += m3[1]; c[2] += m3[2]; }
                        c[0] /= d; c[1] /= d; c[2] /= d;
                        for (int t = 0; t < d; ++t) {
                            edge_mid(v[t], m3);
                            cost += (c[0] - m3[0]) * (c[0] - m3[0]) + (c[1] - m3[1]) * (c[1] - m3[1]) + (c[2] - m3[2]) * (c[2] - m3[2]);
                        }
                    }
                    if (ba < 0 || cost < bc) { bc = cost; ba = a; bk1 = k1; bp = p; bk2 = k2; }
                }
    if (ba < 0) return -1;
    int nt = 0;
    {
        int a = ba, k1 = bk1, p = bp, k2 = bk2, b = (a + k1) % n1, q = (p + k2) % n2;
        int link[16], d1 = 0, rest[16], d2 = 0;
        for (int t = 0; t <= k1; ++t) link[d1++] = L1[(a + t) % n1];
        for (int t = 0; t <= k2; ++t) link[d1++] = L2[(p + t) % n2];
        for (int t = 0; t <= n1 - k1; ++t) rest[d2++] = L1[(b + t) % n1];
        for (int t = 0; t <= n2 - k2; ++t) rest[d2++] = L2[(q + t) % n2];
        for (int pass = 0; pass < 2; ++pass) {
            const int* v = pass ? rest : link;
            int d = pass ? d2 : d1, cid = c0 + pass;
            cen_n[cid] = d;
            for (int t = 0; t < d; ++t) cen_loop[cid][t] = v[t];
            for (int t = 0; t < d; ++t, ++nt) { tris[nt][0] = 12 + cid; tris[nt][1] = v[t]; tris[nt][2] = v[(t + 1) % d]; }
        }
        *ncen += 2;
    }
    return nt;
}

static int64_t g_tunnel_cells = 0;
int64_t mc_oracle_tunnel_cells(void) { return g_tunnel_cells; }   /* cells of the last call that took a tunnel tiling */

int mc_oracle(const float* v, int D, int H, int W, float level, int ascent, mc_result* out) {
    memset(out, 0, sizeof(*out));
    g_tunnel_cells = 0;
    const int64_t vol_n = (int64_t)D * H * W;
    float lo = INFINITY, hi = -INFINITY;
    for (int64_t i = 0; i < vol_n; ++i) { lo = fminf(lo, v[i]); hi = fmaxf(hi, v[i]); }
    if (level < lo || level > hi) { out->status = -4; return -4; }
    int32_t* cache = (int32_t*)malloc(sizeof(int32_t) * 3 * vol_n); /* edge (axis, lower grid point) -> vertex id */
    for (int64_t i = 0; i < 3 * vol_n; ++i) cache[i] = -1;
    grow g;
    memset(&g, 0, sizeof(g));
    int (*face_edge)[4] = FACE_EDGE;
    for (int f = 0; f < 6; ++f)
        for (int i = 0; i < 4; ++i) face_edge[f][i] = edge_between(FACE_CORNER[f][i], FACE_CORNER[f][(i + 1) & 3]);

    for (int z = 0; z < D - 1; ++z)
        for (int y = 0; y < H - 1; ++y)
            for (int x = 0; x < W - 1; ++x) {
                double val[8];
                int s[8], idx = 0;
                for (int i = 0; i < 8; ++i) {
                    val[i] = (double)v[((int64_t)(z + CORNER[i][2]) * H + (y + CORNER[i][1])) * W + (x + CORNER[i][0])] -
                             (double)level;
                    s[i] = val[i] > 0.0;
                    idx |= s[i] << i;
                }
                if (idx == 0 || idx == 255) continue;
                /* directed iso-contour segments on the faces: positive side on the left seen from outside */
                int succ[12], uf[8];
                for (int e = 0; e < 12; ++e) succ[e] = -1;
                for (int c = 0; c < 8; ++c) uf[c] = c;
#define UF_FIND(r, c) do { r = (c); while (uf[r] != r) r = uf[r]; } while (0)
#define UF_JOIN(a_, b_) do { int ra, rb; UF_FIND(ra, a_); UF_FIND(rb, b_); if (ra != rb) uf[ra > rb ? ra : rb] = ra > rb ? rb : ra; } while (0)
                for (int e = 0; e < 12; ++e)
                    if (s[EDGE_CORNER[e][0]] == s[EDGE_CORNER[e][1]]) UF_JOIN(EDGE_CORNER[e][0], EDGE_CORNER[e][1]);
                for (int f = 0; f < 6; ++f) {
                    const int* fc = FACE_CORNER[f];
                    const int* fe = face_edge[f];
                    int fs[4], np = 0;
                    for (int i = 0; i < 4; ++i) { fs[i] = s[fc[i]]; np += fs[i]; }
                    if (np == 0 || np == 4) continue;
                    int ambiguous = fs[0] == fs[2] && fs[1] == fs[3] && fs[0] != fs[1];
                    if (!ambiguous) {
                        int i0 = -1, j0 = -1;
                        for (int i = 0; i < 4; ++i) {
                            if (fs[i] && !fs[(i + 3) & 3]) i0 = i; /* first corner of the positive run */
                            if (fs[i] && !fs[(i + 1) & 3]) j0 = i; /* last corner of the positive run  */
                        }
                        succ[fe[j0]] = fe[(i0 + 3) & 3];
                    } else {
                        /* face test (asymptotic decider with Lewiner's FLT_EPSILON band): the positive corners are
                           connected iff (product of the positive pair) - (product of the negative pair) > -FLT_EPSILON */
                        double a = val[fc[0]], b = val[fc[1]], c = val[fc[2]], d = val[fc[3]];
                        double pp = fs[0] ? a * c : b * d, nn = fs[0] ? b * d : a * c;
                        if (pp - nn > -(double)FLT_EPSILON) {
                            for (int n = 0; n < 4; ++n) if (!fs[n]) succ[fe[(n + 3) & 3]] = fe[n];
                            UF_JOIN(fc[fs[0] ? 0 : 1], fc[fs[0] ? 2 : 3]);
                        } else {
                            for (int p = 0; p < 4; ++p) if (fs[p]) succ[fe[p]] = fe[(p + 3) & 3];
                            UF_JOIN(fc[fs[0] ? 1 : 0], fc[fs[0] ? 3 : 2]);
                        }
                    }
                }
                /* loops in order of their lowest edge */
                int seen[12] = {0}, loops[4][12], ln[4], nl = 0;
                for (int e0 = 0; e0 < 12; ++e0) {
                    if (succ[e0] < 0 || seen[e0]) continue;
                    ln[nl] = 0;
                    for (int e = e0; !seen[e]; e = succ[e]) { seen[e] = 1; loops[nl][ln[nl]++] = e; }
                    ++nl;
                }
                /* interior test: annular regions */
                int tun_a = -1, tun_b = -1;
                if (nl >= 2) {
                    int lreg[4][2]; /* the two regions (roots) a loop separates: [0] = side of the edge's corner 0 ... */
                    for (int l = 0; l < nl; ++l) {
                        int r0, r1;
                        UF_FIND(r0, EDGE_CORNER[loops[l][0]][0]);
                        UF_FIND(r1, EDGE_CORNER[loops[l][0]][1]);
                        lreg[l][0] = r0; lreg[l][1] = r1;
                    }
                    int ann[4][3], na = 0; /* region root, loop a, loop b */
                    for (int r = 0; r < 8; ++r) {
                        if (uf[r] != r) continue;
                        int cnt = 0, la = -1, lb = -1;
                        for (int l = 0; l < nl; ++l)
                            if (lreg[l][0] == r || lreg[l][1] == r) { if (cnt == 0) la = l; else lb = l; ++cnt; }
                        if (cnt == 2) { ann[na][0] = r; ann[na][1] = la; ann[na][2] = lb; ++na; }
                    }
                    for (int k = 0; k < na && tun_a < 0; ++k) {
                        int r = ann[k][0], la = ann[k][1], lb = ann[k][2];
                        int ra = lreg[la][0] == r ? lreg[la][1] : lreg[la][0];
                        int rb = lreg[lb][0] == r ? lreg[lb][1] : lreg[lb][0];
                        int want = 0;
                        for (int p = 0; p < 8 && !want; ++p)
                            for (int q = 0; q < 8; ++q) {
                                int rp, rq;
                                UF_FIND(rp, p); UF_FIND(rq, q);
                                if (rp == ra && rq == rb && manhattan(p, q) == 3) { want = 3; break; }
                            }
                        if (!want && na == 2) want = 2;
                        if (!want) continue;
                        for (int p = 0; p < 8 && tun_a < 0; ++p)
                            for (int q = 0; q < 8; ++q) {
                                int rp, rq;
                                UF_FIND(rp, p); UF_FIND(rq, q);
                                if (rp != ra || rq != rb || manhattan(p, q) != want) continue;
                                if (interior_joined(val, p, q, s[p] ? 1.0 : -1.0)) { tun_a = la; tun_b = lb; break; }
                            }
                    }
                }
                /* cell-local triangle list (vertex ids: 0..11 cube edges, 12.. extra vertices) */
                int tris[16][3], nt = 0, ncen = 0, cen_n[4], cen_loop[4][12];
                for (int l = 0; l < nl; ++l) {
                    if (l == tun_b) continue;
                    if (l == tun_a) {
                        ++g_tunnel_cells;
                        int k = tube_triangulate(loops[tun_a], ln[tun_a], loops[tun_b], ln[tun_b], tris + nt, ncen, &ncen, cen_n,
                                                 cen_loop);
                        if (k < 0) { free(cache); return -9; }
                        nt += k;
                        continue;
                    }
                    const int* poly = loops[l];
                    int n = ln[l], apex[144];
                    for (int i = 0; i < 144; ++i) apex[i] = -2;
                    if (feasible(poly, n, 0, n - 1, apex)) emit_chain(poly, 0, n - 1, apex, tris, &nt);
                    else {
                        cen_n[ncen] = n;
                        for (int i = 0; i < n; ++i) cen_loop[ncen][i] = poly[i];
                        for (int i = 0; i < n; ++i, ++nt) { tris[nt][0] = 12 + ncen; tris[nt][1] = poly[i]; tris[nt][2] = poly[(i + 1) % n]; }
                        ++ncen;
                    }
                }
                /* faces with first-use vertex creation */
                int centre_vid[4] = {-1, -1, -1, -1};
                for (int ti = 0; ti < nt; ++ti) {
                    int vid[3];
                    for (int k = 0; k < 3; ++k) {
                        int e = tris[ti][k];
                        if (e >= 12) {
                            int ci = e - 12;
                            if (centre_vid[ci] < 0) {
                                int n = cen_n[ci];
                                double acc[3] = {0, 0, 0};
                                for (int i = 0; i < n; ++i) {
                                    float c[3]; double t; int lo[3], hi[3];
                                    edge_vertex(val, x, y, z, cen_loop[ci][i], c, &t, lo, hi);
                                    acc[0] += (double)c[0]; acc[1] += (double)c[1]; acc[2] += (double)c[2];
                                }
                                float c[3] = {(float)(acc[0] / (double)n), (float)(acc[1] / (double)n), (float)(acc[2] / (double)n)};
                                float f8[8], vmax = -INFINITY;
                                for (int i = 0; i < 8; ++i) {
                                    f8[i] = v[((int64_t)(z + CORNER[i][2]) * H + (y + CORNER[i][1])) * W + (x + CORNER[i][0])];
                                    vmax = fmaxf(vmax, f8[i]);
                                }
                                float gn[3];
                                { volatile float a = (f8[4] - f8[0]) + (f8[5] - f8[1]); volatile float b = a + (f8[6] - f8[2]); volatile float c2 = b + (f8[7] - f8[3]); gn[0] = c2 * 0.25f; }
                                { volatile float a = (f8[3] - f8[0]) + (f8[2] - f8[1]); volatile float b = a + (f8[7] - f8[4]); volatile float c2 = b + (f8[6] - f8[5]); gn[1] = c2 * 0.25f; }
                                { volatile float a = (f8[1] - f8[0]) + (f8[2] - f8[3]); volatile float b = a + (f8[5] - f8[4]); volatile float c2 = b + (f8[6] - f8[7]); gn[2] = c2 * 0.25f; }
                                normalise(gn);
                                centre_vid[ci] = (int)g.nv;
                                push_vertex(&g, c, gn, vmax);
                            }
                            vid[k] = centre_vid[ci];
                            continue;
                        }
                        float c[3]; double t; int lo[3], hi[3];
                        int ax = edge_vertex(val, x, y, z, e, c, &t, lo, hi);
                        int64_t key = (int64_t)ax * vol_n + ((int64_t)lo[0] * H + lo[1]) * W + lo[2];
                        if (cache[key] < 0) {
                            float vmax = -INFINITY;
                            for (int dz = (ax == 2 ? 0 : -1); dz <= 1; ++dz)
                                for (int dy = (ax == 1 ? 0 : -1); dy <= 1; ++dy)
                                    for (int dx = (ax == 0 ? 0 : -1); dx <= 1; ++dx)
                                        vmax = fmaxf(vmax, vol_at(v, D, H, W, lo[0] + dz, lo[1] + dy, lo[2] + dx));
                            float tt = (float)t, gn[3];
                            for (int q = 0; q < 3; ++q) {
                                int dz = q == 0, dy = q == 1, dx = q == 2;
                                float g0 = vol_at(v, D, H, W, lo[0] + dz, lo[1] + dy, lo[2] + dx) - vol_at(v, D, H, W, lo[0] - dz, lo[1] - dy, lo[2] - dx);
                                float g1 = vol_at(v, D, H, W, hi[0] + dz, hi[1] + dy, hi[2] + dx) - vol_at(v, D, H, W, hi[0] - dz, hi[1] - dy, hi[2] - dx);
                                volatile float p0 = g0 * (1.0f - tt), p1 = g1 * tt; /* no fma */
                                gn[q] = p0 + p1;
                            }
                            normalise(gn);
                            cache[key] = (int32_t)g.nv;
                            push_vertex(&g, c, gn, vmax);
                        }
                        vid[k] = cache[key];
                    }
                    /* native winding: right-hand normal points towards LOWER values ('descent': object greater
                       than exterior -> outward); 'ascent' reverses the column order (np.fliplr in skimage). */
                    if (ascent) push_face(&g, vid[2], vid[1], vid[0]);
                    else push_face(&g, vid[0], vid[1], vid[2]);
                }
            }
    free(cache);
    out->coords = g.coords; out->faces = g.faces; out->normals = g.normals; out->values = g.values;
    out->nv = g.nv; out->nf = g.nf; out->status = 0;
    return 0;
}

void mc_free(mc_result* r) {
    free(r->coords); free(r->faces); free(r->normals); free(r->values);
    memset(r, 0, sizeof(*r));
}
