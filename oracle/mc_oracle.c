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

int mc_oracle(const float* v, int D, int H, int W, float level, int ascent, mc_result* out) {
    memset(out, 0, sizeof(*out));
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
                int succ[12];
                for (int e = 0; e < 12; ++e) succ[e] = -1;
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
                        /* face test: positive corners connected iff saddle value > 0 iff pos pair product > neg pair product */
                        double a = val[fc[0]], b = val[fc[1]], c = val[fc[2]], d = val[fc[3]];
                        double pp = fs[0] ? a * c : b * d, nn = fs[0] ? b * d : a * c;
                        if (pp > nn) {
                            for (int n = 0; n < 4; ++n) if (!fs[n]) succ[fe[(n + 3) & 3]] = fe[n];
                        } else {
                            for (int p = 0; p < 4; ++p) if (fs[p]) succ[fe[p]] = fe[(p + 3) & 3];
                        }
                    }
                }
                /* loops -> triangles -> faces with first-use vertex creation */
                int seen[12] = {0};
                for (int e0 = 0; e0 < 12; ++e0) {
                    if (succ[e0] < 0 || seen[e0]) continue;
                    int poly[12], n = 0;
                    for (int e = e0; !seen[e]; e = succ[e]) { seen[e] = 1; poly[n++] = e; }
                    int tris[12][3], nt = 0, apex[144];
                    for (int i = 0; i < 144; ++i) apex[i] = -2;
                    int centre = !feasible(poly, n, 0, n - 1, apex); /* vertex id 12 = the loop centre */
                    if (!centre) emit_chain(poly, 0, n - 1, apex, tris, &nt);
                    else for (int i = 0; i < n; ++i, ++nt) { tris[nt][0] = 12; tris[nt][1] = poly[i]; tris[nt][2] = poly[(i + 1) % n]; }
                    int centre_vid = -1;
                    for (int ti = 0; ti < nt; ++ti) {
                        int vid[3];
                        for (int k = 0; k < 3; ++k) {
                            int e = tris[ti][k];
                            if (e == 12) {
                                if (centre_vid < 0) {
                                    double acc[3] = {0, 0, 0};
                                    for (int i = 0; i < n; ++i) {
                                        float c[3]; double t; int lo[3], hi[3];
                                        edge_vertex(val, x, y, z, poly[i], c, &t, lo, hi);
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
                                    centre_vid = (int)g.nv;
                                    push_vertex(&g, c, gn, vmax);
                                }
                                vid[k] = centre_vid;
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
