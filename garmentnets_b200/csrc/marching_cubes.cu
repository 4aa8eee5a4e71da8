// Marching cubes on the device (ref predict.py:172-181: skimage.measure.marching_cubes(..., method='lewiner') followed
// by the ggm lookup at trunc(vert/spacing)).
//
// PARITY UNPINNED vs scikit-image 0.18.2: its Lewiner case/tiling tables are not available offline, so the
// triangulation is generated from first principles instead of recited: per cube configuration and per resolution of
// each ambiguous face (asymptotic decider = Lewiner's face test), the iso-contour segments on the six faces are
// chained into closed loops and each loop is fan-triangulated.  What is kept from the reference implementation:
//   * corner / edge numbering and "bit i set iff v_i - level > 0";
//   * cells scanned axis0 -> axis1 -> axis2 (fastest); faces emitted in cell order; vertices de-duplicated per edge
//     and numbered in first-use order of that sequential scan (reproduced in parallel with an owner-cell rule +
//     exclusive scans);
//   * vertex position = inverse-|value| weighted mean of the edge end points (FLT_EPSILON guard), double arithmetic,
//     stored as float32 in voxel units, then * spacing in double, cast to float32 (predict.py:193);
//   * 'ascent' and 'descent' differ by the face column order; ValueError when level is outside [min,max].
// MC33 structure: ambiguous faces are resolved by the face test (asymptotic decider with Lewiner's FLT_EPSILON band), the
// interior ambiguity of sub-cases 4.1 / 6.1 / 7.4 / 10.1 / 12.1 / 13.5 by Chernyaev's interior test (the form Lewiner's
// test_interior uses for cases 4 and 10, applied to every tunnel-capable configuration; it equals the true connectivity of
// the trilinear interpolant, tests/test_mc33.py); a joined pair of regions replaces two caps by a tube (4.1.2, 6.1.2, 7.4.2,
// 10.1.2, 12.1.2, 13.5.2).  Tilings never put an edge inside a cube face other than the iso-contour segments, so cells
// sharing a face cannot collide; loops / tubes without such a triangulation get extra centre vertices.
//
// Kernels: classify (cube code per cell + per-block vertex/face counts + min/max), single-CTA scan of the block
// counts, vertex emission by owner cells (+ edge -> vertex map), face emission.  All integer/byte work, HBM-bound:
// one coalesced pass over the volume, table lookups served from L2.
#include "common.cuh"
#include <float.h>
#include <string.h>

namespace gnb {

constexpr int MC_MAX_TRI = 14, MC_MAX_CEN = 2, MC_MAX_TUN = 160;
struct McEntry {
    uint8_t ntri, nedge, ncen, pad0;
    uint8_t order[12 + MC_MAX_CEN];      // distinct vertex ids in first-use order (0..11 cube edges, 12+ loop centres)
    uint8_t tri[3 * MC_MAX_TRI];         // ntri * 3 vertex ids
    uint8_t cen_n[MC_MAX_CEN];           // loop length of each centre vertex
    uint8_t cen_loop[MC_MAX_CEN][12];    // the loop's edges, in loop order
};
static_assert(sizeof(McEntry) == 4 + 14 + 3 * MC_MAX_TRI + 2 + 24, "McEntry layout");
// interior-test descriptor of one (cube index, face decisions) configuration with annular regions
struct McTunDesc {
    uint8_t nann, npair[2], pad;
    int8_t sigma[2];            // sign of the two regions the tunnel would join
    uint8_t col[2][3][8];       // per annulus, per candidate corner pair: corner ids A0 A1 B0 B1 C0 C1 D0 D1 (sweep columns)
};

__device__ McEntry d_mc_table[256 * 64];
__device__ McEntry d_mc_tun_table[MC_MAX_TUN];     // tunnel tilings: entry d_mc_tun_index[key] + annulus
__device__ McTunDesc d_mc_tun_desc[MC_MAX_TUN];    // descriptor at d_mc_tun_index[key]
__device__ int16_t d_mc_tun_index[256 * 64];       // -1: no interior ambiguity
__device__ uint8_t d_mc_cnt[256 * 64];             // triangles | centre vertices << 4 of d_mc_table (1 byte instead of 86)
__device__ uint8_t d_mc_tun_cnt[MC_MAX_TUN];
// per cube index, for the common cells without any ambiguity: {edge mask, counts of the fb = 0 tiling, needs_resolve}
struct McFast { uint16_t edgemask; uint8_t cnt0; uint8_t resolve; uint16_t eid0; uint16_t pad; };   // eid0: compact id of the fb = 0 tiling
__device__ McFast d_mc_fast[256];
// all tilings back to back (the 656 valid (index, face decision) entries, then the tunnel entries) for kernels that keep the
// whole table in shared memory; d_mc_entry_id maps index * 64 + face bits -> position, tunnel entry k of a key follows at
// d_mc_n_base + d_mc_tun_index[key] + k
constexpr int MC_MAX_ENTRIES = 700 + MC_MAX_TUN;
__device__ McEntry d_mc_compact[MC_MAX_ENTRIES];
// the same tilings packed for the emit kernels: 4-bit vertex ids (0..11 cube edges, 12+ centres), no per-byte loads.
//   order: first-use vertex order, slot j in bits 4j..4j+3 (14 slots), nedge in bits 56..59, ntri in bits 60..63
//   tri[w]: five triangles per word, triangle t in bits 12*(t%5) .. +11 of word t/5 (corner c in bits 4c..4c+3 of those)
struct McPacked { unsigned long long order, tri[3]; };
__device__ McPacked d_mc_packed[MC_MAX_ENTRIES];
__device__ uint16_t d_mc_entry_id[256 * 64];
__device__ int d_mc_n_base;
__device__ uint16_t d_mc_edgemask[256];
__device__ uint8_t d_mc_ambig[256];  // bit f set: face f is ambiguous for this cube index

// cube topology (corner c_i at (x,y,z) = bits of {0:000,1:100,2:110,3:010,4:001,5:101,6:111,7:011})
__constant__ int8_t c_corner_dx[8] = {0, 1, 1, 0, 0, 1, 1, 0};
__constant__ int8_t c_corner_dy[8] = {0, 0, 1, 1, 0, 0, 1, 1};
__constant__ int8_t c_corner_dz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
// faces: corners counter-clockwise seen from outside; edge i joins corner i and corner (i+1)%4
__constant__ int8_t c_face_corner[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {3, 7, 6, 2}, {0, 4, 7, 3}, {1, 2, 6, 5}};
static const int8_t h_face_corner[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {3, 7, 6, 2}, {0, 4, 7, 3}, {1, 2, 6, 5}};
static const int8_t h_face_edge[6][4] = {{3, 2, 1, 0}, {4, 5, 6, 7}, {0, 9, 4, 8}, {11, 6, 10, 2}, {8, 7, 11, 3}, {1, 10, 5, 9}};
static const int8_t h_edge_corner[12][2] = {{0, 1}, {1, 2}, {3, 2}, {0, 3}, {4, 5}, {5, 6}, {7, 6}, {4, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
// edge -> (axis: 0 = x/axis2, 1 = y/axis1, 2 = z/axis0 ; lower end point offset dx,dy,dz)
__constant__ int8_t c_edge_axis[12] = {0, 1, 0, 1, 0, 1, 0, 1, 2, 2, 2, 2};
__constant__ int8_t c_edge_dx[12] = {0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 0};
__constant__ int8_t c_edge_dy[12] = {0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1};
__constant__ int8_t c_edge_dz[12] = {0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0};

// ---- host: table generation -----------------------------------------------------------------------------
// Two cube edges are co-facial when one face contains both: a triangle edge joining their vertices would lie inside
// that face (and collide with the neighbouring cell's triangles), so such diagonals are forbidden.
static bool cofacial(int e1, int e2) {
    for (int f = 0; f < 6; ++f) {
        bool a = false, b = false;
        for (int i = 0; i < 4; ++i) { a |= h_face_edge[f][i] == e1; b |= h_face_edge[f][i] == e2; }
        if (a && b) return true;
    }
    return false;
}

// Deterministic triangulation of the loop poly[0..n) without forbidden diagonals: interval DP, smallest apex first,
// triangles emitted (i,k,j) then the left sub-chain then the right one.  Returns the triangle count, 0 if impossible.
static bool tl_ok[12][12];
static int tl_choice[12][12];
static bool tl_chord(const int* poly, int n, int a, int b) {
    if (b == a + 1 || (a == 0 && b == n - 1)) return true;
    return !cofacial(poly[a], poly[b]);
}
static void tl_emit(const int* poly, int i, int j, int (*tris)[3], int& nt) {
    if (j - i < 2) return;
    const int k = tl_choice[i][j];
    tris[nt][0] = poly[i]; tris[nt][1] = poly[k]; tris[nt][2] = poly[j];
    ++nt;
    tl_emit(poly, i, k, tris, nt);
    tl_emit(poly, k, j, tris, nt);
}
static int triangulate_loop(const int* poly, int n, int (*tris)[3]) {
    for (int len = 1; len < n; ++len)
        for (int i = 0; i + len < n; ++i) {
            const int j = i + len;
            if (len == 1) { tl_ok[i][j] = true; continue; }
            tl_ok[i][j] = false;
            for (int k = i + 1; k < j; ++k)
                if (tl_chord(poly, n, i, k) && tl_chord(poly, n, k, j) && tl_ok[i][k] && tl_ok[k][j]) {
                    tl_ok[i][j] = true;
                    tl_choice[i][j] = k;
                    break;
                }
        }
    if (!tl_ok[0][n - 1]) return 0;
    int nt = 0;
    tl_emit(poly, 0, n - 1, tris, nt);
    return nt;
}

// ---- host: MC33 configuration analysis ---------------------------------------------------------------------------
static const int8_t h_corner_xyz[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static int corner_dist(int a, int b) {
    int d = 0;
    for (int i = 0; i < 3; ++i) d += h_corner_xyz[a][i] != h_corner_xyz[b][i];
    return d;
}
static int corner_flip(int c, int axis) {
    int q[3] = {h_corner_xyz[c][0], h_corner_xyz[c][1], h_corner_xyz[c][2]};
    q[axis] ^= 1;
    for (int k = 0; k < 8; ++k)
        if (h_corner_xyz[k][0] == q[0] && h_corner_xyz[k][1] == q[1] && h_corner_xyz[k][2] == q[2]) return k;
    return -1;
}

struct CellConfig {
    int nloop = 0, len[4] = {0, 0, 0, 0}, loop[4][12];
    uint8_t region_of[8];        // region id of every corner = smallest corner index of its region
    uint8_t side[4][2];          // the two regions a loop separates
    int nann = 0, ann_loop[2][2], ann_npair[2] = {0, 0}, ann_pair[2][3][2];
};

// loops of iso-contour segments (directed: positive side on the left seen from outside), the regions they cut the cube
// surface into, the annular regions and the corner pairs the interior test examines
static CellConfig analyse_config(int idx, int fb, int am) {
    CellConfig cc;
    int succ[12];
    for (int e = 0; e < 12; ++e) succ[e] = -1;
    // regions as corner bitmasks, merged until stable
    uint8_t mask[8];
    for (int c = 0; c < 8; ++c) mask[c] = (uint8_t)(1u << c);
    auto merge = [&](int a, int b) {
        const uint8_t m = mask[a] | mask[b];
        for (int c = 0; c < 8; ++c)
            if ((m >> c) & 1) mask[c] = m;
    };
    for (int e = 0; e < 12; ++e) {
        const int a = h_edge_corner[e][0], b = h_edge_corner[e][1];
        if (((idx >> a) & 1) == ((idx >> b) & 1)) merge(a, b);
    }
    for (int f = 0; f < 6; ++f) {
        int s[4], np = 0;
        for (int i = 0; i < 4; ++i) { s[i] = (idx >> h_face_corner[f][i]) & 1; np += s[i]; }
        if (np == 0 || np == 4) continue;
        const int8_t* fe = h_face_edge[f];
        const int8_t* fc = h_face_corner[f];
        if (!((am >> f) & 1)) {
            int i0 = -1, j0 = -1;  // positive run i0..j0 (ccw)
            for (int i = 0; i < 4; ++i) {
                if (s[i] && !s[(i + 3) & 3]) i0 = i;
                if (s[i] && !s[(i + 1) & 3]) j0 = i;
            }
            succ[fe[j0]] = fe[(i0 + 3) & 3];
        } else if (!((fb >> f) & 1)) {  // positive corners separated: the negative diagonal is connected
            for (int q = 0; q < 4; ++q)
                if (s[q]) succ[fe[q]] = fe[(q + 3) & 3];
            merge(fc[s[0] ? 1 : 0], fc[s[0] ? 3 : 2]);
        } else {  // positive corners connected: the negative corners are cut off
            for (int n = 0; n < 4; ++n)
                if (!s[n]) succ[fe[(n + 3) & 3]] = fe[n];
            merge(fc[s[0] ? 0 : 1], fc[s[0] ? 2 : 3]);
        }
    }
    for (int c = 0; c < 8; ++c) {
        int r = 0;
        while (!((mask[c] >> r) & 1)) ++r;
        cc.region_of[c] = (uint8_t)r;
    }
    bool seen[12] = {false};
    for (int e0 = 0; e0 < 12; ++e0) {
        if (succ[e0] < 0 || seen[e0]) continue;
        int n = 0;
        for (int e = e0; !seen[e]; e = succ[e]) { seen[e] = true; cc.loop[cc.nloop][n++] = e; }
        cc.len[cc.nloop] = n;
        cc.side[cc.nloop][0] = cc.region_of[h_edge_corner[e0][0]];
        cc.side[cc.nloop][1] = cc.region_of[h_edge_corner[e0][1]];
        ++cc.nloop;
    }
    // annuli (regions bounded by exactly two loops), by region id
    int ann_region[4], ann_l[4][2], nann_all = 0;
    for (int r = 0; r < 8; ++r) {
        if (cc.region_of[r] != r) continue;
        int cnt = 0, l2[2] = {-1, -1};
        for (int l = 0; l < cc.nloop; ++l)
            if (cc.side[l][0] == r || cc.side[l][1] == r) { if (cnt < 2) l2[cnt] = l; ++cnt; }
        if (cnt == 2) { ann_region[nann_all] = r; ann_l[nann_all][0] = l2[0]; ann_l[nann_all][1] = l2[1]; ++nann_all; }
    }
    for (int k = 0; k < nann_all; ++k) {
        const int r = ann_region[k], la = ann_l[k][0], lb = ann_l[k][1];
        const int ra = cc.side[la][0] == r ? cc.side[la][1] : cc.side[la][0];
        const int rb = cc.side[lb][0] == r ? cc.side[lb][1] : cc.side[lb][0];
        int want = 0;
        for (int a = 0; a < 8; ++a)
            for (int b = 0; b < 8; ++b)
                if (cc.region_of[a] == ra && cc.region_of[b] == rb && corner_dist(a, b) == 3) want = 3;
        if (!want && nann_all == 2) want = 2;   // case 13.5: nested annuli, face-diagonal pairs
        if (!want) continue;
        const int o = cc.nann++;
        cc.ann_loop[o][0] = la; cc.ann_loop[o][1] = lb;
        for (int a = 0; a < 8; ++a)
            for (int b = 0; b < 8; ++b)
                if (cc.region_of[a] == ra && cc.region_of[b] == rb && corner_dist(a, b) == want) {
                    if (cc.ann_npair[o] >= 3) { fprintf(stderr, "mc table: too many interior-test pairs\n"); abort(); }
                    cc.ann_pair[o][cc.ann_npair[o]][0] = a; cc.ann_pair[o][cc.ann_npair[o]][1] = b;
                    ++cc.ann_npair[o];
                }
    }
    return cc;
}

// sweep columns of the interior test for the corner pair (p, q): A0 A1 B0 B1 C0 C1 D0 D1
static void sweep_columns(int p, int q, uint8_t col[8]) {
    int axis = 2;
    if (corner_dist(p, q) == 2)
        for (int i = 0; i < 3; ++i)
            if (h_corner_xyz[p][i] == h_corner_xyz[q][i]) axis = i;
    const int A0 = p, C0 = h_corner_xyz[q][axis] == h_corner_xyz[p][axis] ? q : corner_flip(q, axis);
    int other[2], no = 0;
    for (int c = 0; c < 8; ++c)
        if (h_corner_xyz[c][axis] == h_corner_xyz[p][axis] && c != A0 && c != C0) other[no++] = c;
    col[0] = (uint8_t)A0; col[1] = (uint8_t)corner_flip(A0, axis);
    col[2] = (uint8_t)other[0]; col[3] = (uint8_t)corner_flip(other[0], axis);
    col[4] = (uint8_t)C0; col[5] = (uint8_t)corner_flip(C0, axis);
    col[6] = (uint8_t)other[1]; col[7] = (uint8_t)corner_flip(other[1], axis);
}

// ---- host: tube between two loops -----------------------------------------------------------------------------------
// Edge midpoints in units of 1/5040 (the mean of 3..10 of them stays integral): integer costs, no rounding.
static void mid_of(int e, long long m[3]) {
    for (int i = 0; i < 3; ++i) m[i] = (long long)(h_corner_xyz[h_edge_corner[e][0]][i] + h_corner_xyz[h_edge_corner[e][1]][i]) * 2520;
}
static long long mid_dist2(int e1, int e2) {
    long long a[3], b[3], d = 0;
    mid_of(e1, a); mid_of(e2, b);
    for (int i = 0; i < 3; ++i) d += (a[i] - b[i]) * (a[i] - b[i]);
    return d;
}
struct Triangulation {
    int n = 0;
    long long cost = 0;
    uint8_t v[16][3];
    bool less_than(const Triangulation& o) const {   // (cost, emitted vertex sequence)
        if (cost != o.cost) return cost < o.cost;
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k)
                if (v[i][k] != o.v[i][k]) return v[i][k] < o.v[i][k];
        return false;
    }
    bool simplicial() const {                         // every oriented edge at most once, no degenerate triangle
        bool used[16][16] = {{false}};
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) {
                const int a = v[i][k], b = v[i][(k + 1) % 3];
                if (a == b || used[a][b]) return false;
                used[a][b] = true;
            }
        return true;
    }
};
// all triangulations of the polygon P[lo..hi] (closing chord lo-hi), appended to `cur` in the order (lo,k,hi), left, right;
// `todo` holds the intervals still to split (the last one is split next)
static void enumerate_polygon(const int* P, int m, int (*todo)[2], int ntodo, Triangulation& cur, Triangulation& best,
                              bool& have) {
    while (ntodo > 0 && todo[ntodo - 1][1] - todo[ntodo - 1][0] < 2) --ntodo;
    if (ntodo == 0) {
        if (cur.simplicial() && (!have || cur.less_than(best))) { best = cur; have = true; }
        return;
    }
    const int lo = todo[ntodo - 1][0], hi = todo[ntodo - 1][1];
    auto side_ok = [&](int a, int b) {
        if (b == a + 1 || (a == 0 && b == m - 1)) return true;
        return P[a] != P[b] && !cofacial(P[a], P[b]);
    };
    auto side_cost = [&](int a, int b) -> long long {
        if (b == a + 1 || (a == 0 && b == m - 1)) return 0;
        return mid_dist2(P[a], P[b]);
    };
    for (int k = lo + 1; k < hi; ++k) {
        if (!side_ok(lo, k) || !side_ok(k, hi)) continue;
        if (P[lo] == P[k] || P[k] == P[hi] || P[lo] == P[hi]) continue;
        int next[16][2];
        for (int i = 0; i < ntodo - 1; ++i) { next[i][0] = todo[i][0]; next[i][1] = todo[i][1]; }
        next[ntodo - 1][0] = k; next[ntodo - 1][1] = hi;
        next[ntodo][0] = lo; next[ntodo][1] = k;
        Triangulation t = cur;
        t.v[t.n][0] = (uint8_t)P[lo]; t.v[t.n][1] = (uint8_t)P[k]; t.v[t.n][2] = (uint8_t)P[hi];
        ++t.n;
        t.cost += side_cost(lo, k) + side_cost(k, hi);
        enumerate_polygon(P, m, next, ntodo + 1, t, best, have);
    }
}
// Triangulates the annulus between the directed loops L1, L2.  Vertex ids 0..11 are cube edges, 12 + c0 (+1) extra
// vertices whose definitions go to cen_n / cen_loop.  Returns false if no tiling exists (never for the MC33 cases).
static bool tube_between(const int* L1, int n1, const int* L2, int n2, int c0, Triangulation& out, int& ncen_used,
                         uint8_t* cen_n, uint8_t (*cen_loop)[12]) {
    bool have = false;
    Triangulation best;
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j) {   // bridge L1[i] - L2[j] cuts the annulus open into one polygon
            if (cofacial(L1[i], L2[j])) continue;
            int P[16], m = 0;
            for (int k = 0; k < n1; ++k) P[m++] = L1[(i + k) % n1];
            P[m++] = L1[i];
            for (int k = 0; k < n2; ++k) P[m++] = L2[(j + k) % n2];
            P[m++] = L2[j];
            Triangulation cur, b;
            cur.cost = mid_dist2(L1[i], L2[j]);
            bool h = false;
            int todo[16][2] = {{0, m - 1}};
            enumerate_polygon(P, m, todo, 1, cur, b, h);
            if (h && (!have || b.cost < best.cost)) { best = b; have = true; }   // ties: the first bridge wins
        }
    ncen_used = 0;
    if (have) { out = best; return true; }
    // two fans around two extra vertices: vertex 1 over L1[a..a+k1] + L2[p..p+k2], vertex 2 over the rest
    long long best_cost = -1;
    int sel[4] = {0, 0, 0, 0};
    auto fan_cost = [&](const int* ring, int d) {
        long long c[3] = {0, 0, 0}, m3[3], sum = 0;
        for (int t = 0; t < d; ++t) { mid_of(ring[t], m3); c[0] += m3[0]; c[1] += m3[1]; c[2] += m3[2]; }
        for (int i = 0; i < 3; ++i) c[i] /= d;
        for (int t = 0; t < d; ++t) {
            mid_of(ring[t], m3);
            for (int i = 0; i < 3; ++i) sum += (c[i] - m3[i]) * (c[i] - m3[i]);
        }
        return sum;
    };
    auto rings = [&](int a, int k1, int p, int k2, int* link, int& d1, int* rest, int& d2) {
        const int b = (a + k1) % n1, q = (p + k2) % n2;
        d1 = d2 = 0;
        for (int t = 0; t <= k1; ++t) link[d1++] = L1[(a + t) % n1];
        for (int t = 0; t <= k2; ++t) link[d1++] = L2[(p + t) % n2];
        for (int t = 0; t <= n1 - k1; ++t) rest[d2++] = L1[(b + t) % n1];
        for (int t = 0; t <= n2 - k2; ++t) rest[d2++] = L2[(q + t) % n2];
    };
    for (int a = 0; a < n1; ++a)
        for (int k1 = 1; k1 < n1; ++k1)
            for (int p = 0; p < n2; ++p)
                for (int k2 = 1; k2 < n2; ++k2) {
                    const int b = (a + k1) % n1, q = (p + k2) % n2;
                    if (cofacial(L1[b], L2[p]) || cofacial(L2[q], L1[a])) continue;
                    int link[16], rest[16], d1, d2;
                    rings(a, k1, p, k2, link, d1, rest, d2);
                    const long long c = mid_dist2(L1[b], L2[p]) + mid_dist2(L2[q], L1[a]) + fan_cost(link, d1) + fan_cost(rest, d2);
                    if (best_cost < 0 || c < best_cost) { best_cost = c; sel[0] = a; sel[1] = k1; sel[2] = p; sel[3] = k2; }
                }
    if (best_cost < 0) return false;
    int link[16], rest[16], d1, d2;
    rings(sel[0], sel[1], sel[2], sel[3], link, d1, rest, d2);
    out = Triangulation();
    for (int pass = 0; pass < 2; ++pass) {
        const int* ring = pass ? rest : link;
        const int d = pass ? d2 : d1, cid = c0 + pass;
        if (cid >= MC_MAX_CEN || d > 12) { fprintf(stderr, "mc table overflow (tube centres)\n"); abort(); }
        cen_n[cid] = (uint8_t)d;
        for (int t = 0; t < d; ++t) cen_loop[cid][t] = (uint8_t)ring[t];
        for (int t = 0; t < d; ++t) {
            out.v[out.n][0] = (uint8_t)(12 + cid); out.v[out.n][1] = (uint8_t)ring[t]; out.v[out.n][2] = (uint8_t)ring[(t + 1) % d];
            ++out.n;
        }
    }
    ncen_used = 2;
    return true;
}

// tiling of one configuration; tunnel >= 0 selects the annulus whose two loops are joined by a tube
static void build_entry(const CellConfig& cc, int tunnel, McEntry& en) {
    memset(&en, 0, sizeof(en));
    int nt = 0, nc = 0;
    const int ta = tunnel >= 0 ? cc.ann_loop[tunnel][0] : -1, tb = tunnel >= 0 ? cc.ann_loop[tunnel][1] : -1;
    for (int l = 0; l < cc.nloop; ++l) {
        if (l == tb) continue;
        const int* poly = cc.loop[l];
        const int n = cc.len[l];
        if (l == ta) {
            Triangulation t;
            int used = 0;
            if (!tube_between(cc.loop[ta], cc.len[ta], cc.loop[tb], cc.len[tb], nc, t, used, en.cen_n, en.cen_loop)) {
                fprintf(stderr, "mc table: no tube tiling\n"); abort();
            }
            if (nt + t.n > MC_MAX_TRI) { fprintf(stderr, "mc table overflow (tube)\n"); abort(); }
            for (int i = 0; i < t.n; ++i, ++nt)
                for (int q = 0; q < 3; ++q) en.tri[nt * 3 + q] = t.v[i][q];
            nc += used;
            continue;
        }
        int tris[12][3];
        const int k = triangulate_loop(poly, n, tris);
        if (k > 0) {
            if (nt + k > MC_MAX_TRI) { fprintf(stderr, "mc table overflow\n"); abort(); }
            for (int t = 0; t < k; ++t, ++nt)
                for (int q = 0; q < 3; ++q) en.tri[nt * 3 + q] = (uint8_t)tris[t][q];
        } else {  // no diagonal-safe triangulation: fan around an extra centre vertex
            if (nc >= MC_MAX_CEN || nt + n > MC_MAX_TRI) { fprintf(stderr, "mc table overflow\n"); abort(); }
            en.cen_n[nc] = (uint8_t)n;
            for (int i = 0; i < n; ++i) en.cen_loop[nc][i] = (uint8_t)poly[i];
            for (int i = 0; i < n; ++i, ++nt) {
                en.tri[nt * 3 + 0] = (uint8_t)(12 + nc);
                en.tri[nt * 3 + 1] = (uint8_t)poly[i];
                en.tri[nt * 3 + 2] = (uint8_t)poly[(i + 1) % n];
            }
            ++nc;
        }
    }
    en.ncen = (uint8_t)nc;
    en.ntri = (uint8_t)nt;
    bool used[12 + MC_MAX_CEN] = {false};
    int no = 0;
    for (int t = 0; t < nt * 3; ++t)
        if (!used[en.tri[t]]) { used[en.tri[t]] = true; en.order[no++] = en.tri[t]; }
    en.nedge = (uint8_t)no;
}

struct McTables {
    McEntry table[256 * 64];
    McEntry tun_table[MC_MAX_TUN];
    McTunDesc tun_desc[MC_MAX_TUN];
    int16_t tun_index[256 * 64];
    uint16_t edgemask[256];
    uint8_t ambig[256];
    int ntun = 0;
};

static void build_tables(McTables& T) {
    T.ntun = 0;
    memset(T.tun_table, 0, sizeof(T.tun_table));
    memset(T.tun_desc, 0, sizeof(T.tun_desc));
    for (int idx = 0; idx < 256; ++idx) {
        uint16_t em = 0;
        for (int e = 0; e < 12; ++e)
            if (((idx >> h_edge_corner[e][0]) & 1) != ((idx >> h_edge_corner[e][1]) & 1)) em |= (uint16_t)(1u << e);
        T.edgemask[idx] = em;
        uint8_t am = 0;
        for (int f = 0; f < 6; ++f) {
            int s[4];
            for (int i = 0; i < 4; ++i) s[i] = (idx >> h_face_corner[f][i]) & 1;
            if (s[0] == s[2] && s[1] == s[3] && s[0] != s[1]) am |= (uint8_t)(1u << f);
        }
        T.ambig[idx] = am;
        for (int fb = 0; fb < 64; ++fb) {
            McEntry& en = T.table[idx * 64 + fb];
            memset(&en, 0, sizeof(en));
            T.tun_index[idx * 64 + fb] = -1;
            if ((fb & ~am) || idx == 0 || idx == 255) continue;  // decision bits only exist for ambiguous faces
            const CellConfig cc = analyse_config(idx, fb, am);
            build_entry(cc, -1, en);
            if (cc.nann == 0) continue;
            if (T.ntun + cc.nann > MC_MAX_TUN) { fprintf(stderr, "mc table overflow (tunnel entries)\n"); abort(); }
            T.tun_index[idx * 64 + fb] = (int16_t)T.ntun;
            McTunDesc& d = T.tun_desc[T.ntun];
            d.nann = (uint8_t)cc.nann;
            for (int k = 0; k < cc.nann; ++k) {
                d.npair[k] = (uint8_t)cc.ann_npair[k];
                d.sigma[k] = ((idx >> cc.ann_pair[k][0][0]) & 1) ? 1 : -1;
                for (int q = 0; q < cc.ann_npair[k]; ++q) sweep_columns(cc.ann_pair[k][q][0], cc.ann_pair[k][q][1], d.col[k][q]);
                build_entry(cc, k, T.tun_table[T.ntun + k]);
            }
            T.ntun += cc.nann;
        }
    }
}

static int g_mc_n_entries = 0;   // entries of d_mc_compact (host copy: sizes the shared-memory table of the emit kernels)

static int ensure_tables() {
    static bool done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (done[dev]) return 0;
    static McTables* T = nullptr;
    if (!T) {
        T = new McTables();
        build_tables(*T);
    }
    if (cudaMemcpyToSymbol(d_mc_table, T->table, sizeof(T->table)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(d_mc_tun_table, T->tun_table, sizeof(T->tun_table)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(d_mc_tun_desc, T->tun_desc, sizeof(T->tun_desc)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(d_mc_tun_index, T->tun_index, sizeof(T->tun_index)) != cudaSuccess) return -1;
    {
        static uint8_t cnt[256 * 64], tcnt[MC_MAX_TUN];
        for (int i = 0; i < 256 * 64; ++i) cnt[i] = (uint8_t)(T->table[i].ntri | (T->table[i].ncen << 4));
        for (int i = 0; i < MC_MAX_TUN; ++i) tcnt[i] = (uint8_t)(T->tun_table[i].ntri | (T->tun_table[i].ncen << 4));
        if (cudaMemcpyToSymbol(d_mc_cnt, cnt, sizeof(cnt)) != cudaSuccess) return -1;
        if (cudaMemcpyToSymbol(d_mc_tun_cnt, tcnt, sizeof(tcnt)) != cudaSuccess) return -1;
    }
    {
        static McFast fast[256];
        for (int i = 0; i < 256; ++i) {
            fast[i].edgemask = T->edgemask[i];
            fast[i].cnt0 = (uint8_t)(T->table[i * 64].ntri | (T->table[i * 64].ncen << 4));
            fast[i].resolve = (uint8_t)((T->ambig[i] != 0 || T->tun_index[i * 64] >= 0) ? 1 : 0);
            // compact id of (i, fb = 0): the number of valid entries in front of it (same enumeration as below)
            int n = 0;
            for (int idx = 0; idx < i; ++idx)
                for (int fb = 0; fb < 64; ++fb)
                    if (!((fb & ~T->ambig[idx]) || idx == 0 || idx == 255)) ++n;
            fast[i].eid0 = (uint16_t)n;
            fast[i].pad = 0;
        }
        if (cudaMemcpyToSymbol(d_mc_fast, fast, sizeof(fast)) != cudaSuccess) return -1;
    }
    {
        static McEntry compact[MC_MAX_ENTRIES];
        static uint16_t eid[256 * 64];
        int n = 0;
        for (int idx = 0; idx < 256; ++idx)
            for (int fb = 0; fb < 64; ++fb) {
                eid[idx * 64 + fb] = 0;
                if ((fb & ~T->ambig[idx]) || idx == 0 || idx == 255) continue;
                eid[idx * 64 + fb] = (uint16_t)n;
                compact[n++] = T->table[idx * 64 + fb];
            }
        const int n_base = n;
        for (int i = 0; i < T->ntun; ++i) compact[n++] = T->tun_table[i];
        if (n > MC_MAX_ENTRIES) return -1;
        g_mc_n_entries = n;
        if (cudaMemcpyToSymbol(d_mc_compact, compact, sizeof(McEntry) * n) != cudaSuccess) return -1;
        static McPacked packed[MC_MAX_ENTRIES];
        for (int i = 0; i < n; ++i) {
            const McEntry& en = compact[i];
            McPacked pk = {0ull, {0ull, 0ull, 0ull}};
            for (int j = 0; j < en.nedge; ++j) pk.order |= (unsigned long long)(en.order[j] & 15) << (4 * j);
            pk.order |= (unsigned long long)en.nedge << 56;
            pk.order |= (unsigned long long)en.ntri << 60;
            for (int t = 0; t < en.ntri; ++t)
                for (int c = 0; c < 3; ++c) pk.tri[t / 5] |= (unsigned long long)(en.tri[t * 3 + c] & 15) << (12 * (t % 5) + 4 * c);
            packed[i] = pk;
        }
        if (cudaMemcpyToSymbol(d_mc_packed, packed, sizeof(McPacked) * n) != cudaSuccess) return -1;
        if (cudaMemcpyToSymbol(d_mc_entry_id, eid, sizeof(eid)) != cudaSuccess) return -1;
        if (cudaMemcpyToSymbol(d_mc_n_base, &n_base, sizeof(int)) != cudaSuccess) return -1;
    }
    if (cudaMemcpyToSymbol(d_mc_edgemask, T->edgemask, sizeof(T->edgemask)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(d_mc_ambig, T->ambig, sizeof(T->ambig)) != cudaSuccess) return -1;
    done[dev] = true;
    return 0;
}

// ---- workspace --------------------------------------------------------------------------------------------
// One block of gnb_mc_workspace_bytes(D,H,W) bytes per volume (batch: N blocks `ws_stride` bytes apart).
// A cell is named by the voxel index of its lower corner, c = (z*H + y)*W + x ("padded" cell space: the cells with
// x = W-1, y = H-1 or z = D-1 do not exist and stay inactive), so rows of cells are rows of the volume.
// Work is cut into ITEMS = (volume row, 128-value segment): one warp classifies one item with four 16-byte loads per
// lane (the rows y, y+1 of the slices z, z+1) and gets the x+1 corners of its last cell from the next lane by shuffle.
// Counts (vertices / faces / active cells) are kept per ITEM, so a warp never synchronises with the rest of its CTA.
constexpr int MC_WARPS = 8, MC_BLOCK = MC_WARPS * 32, MC_SEG = 128;
struct McRec {          // 512-byte record, read back by the host in ONE copy for the whole batch
    int64_t V, F, A;    // vertices, faces, active cells (cube index not 0 / 255)
    int64_t vbase, fbase;  // first row of this volume in the concatenated vertex / face outputs of the batch
    int64_t pad0[27];
    unsigned min_enc, max_enc;  // byte 256: order-preserving encodings of the data range
    int n_resolve;              // cells whose tiling needs the face / interior tests (listed in the `active` area during counting)
    unsigned pad1[61];
};
static_assert(sizeof(McRec) == 512, "McRec layout");
struct McWs {
    uint16_t* codes;    // [D*H*W] cube index | face bits << 8 | tunnel << 14 (0 for the padding cells)
    int32_t* blockV;    // [nb = nitems] per-item counts -> exclusive offsets after the scan
    int32_t* blockF;    // [nb]
    int32_t* blockA;    // [nb]
    McRec* rec;
    int4* active;       // [A] compacted active cells in scan order: {cell, first vertex id, first face id, code}
    int32_t* edge_map;  // [(3+MC_MAX_CEN)*D*H*W] global edge / cell centre -> vertex id
};
struct McGeom { int64_t nvox, nseg, nitems, nb; int64_t o_blockV, o_blockF, o_blockA, o_rec, o_active, o_edge, total; };
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static McGeom geom(int D, int H, int W) {
    McGeom g;
    g.nvox = (int64_t)D * H * W;
    g.nseg = ceil_div<int64_t>(W, MC_SEG);
    g.nitems = (int64_t)D * H * g.nseg;
    g.nb = g.nitems;
    size_t p = align256(sizeof(uint16_t) * g.nvox);
    g.o_blockV = p; p += align256(sizeof(int32_t) * g.nb);
    g.o_blockF = p; p += align256(sizeof(int32_t) * g.nb);
    g.o_blockA = p; p += align256(sizeof(int32_t) * g.nb);
    g.o_rec = p; p += sizeof(McRec);
    g.o_active = p; p += align256(sizeof(int4) * g.nvox);
    g.o_edge = p; p += align256(sizeof(int32_t) * (3 + MC_MAX_CEN) * (size_t)g.nvox);
    g.total = p;
    return g;
}
struct McBatch { char* base; int64_t stride; McGeom g; };
__host__ __device__ __forceinline__ McWs carve(const McBatch& b, int vol) {
    char* p = b.base + (int64_t)vol * b.stride;
    McWs w;
    w.codes = reinterpret_cast<uint16_t*>(p);
    w.blockV = reinterpret_cast<int32_t*>(p + b.g.o_blockV);
    w.blockF = reinterpret_cast<int32_t*>(p + b.g.o_blockF);
    w.blockA = reinterpret_cast<int32_t*>(p + b.g.o_blockA);
    w.rec = reinterpret_cast<McRec*>(p + b.g.o_rec);
    w.active = reinterpret_cast<int4*>(p + b.g.o_active);
    w.edge_map = reinterpret_cast<int32_t*>(p + b.g.o_edge);
    return w;
}

// ---- device helpers -----------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned mc_enc(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// edges whose vertex is created by THIS cell in a sequential z->y->x scan (first cell that contains the edge)
__device__ __forceinline__ unsigned owned_mask(int z, int y, int x) {
    unsigned m = (1u << 5) | (1u << 6) | (1u << 10);
    if (z == 0) m |= (1u << 1) | (1u << 2);
    if (y == 0) m |= (1u << 4) | (1u << 9);
    if (x == 0) m |= (1u << 7) | (1u << 11);
    if (z == 0 && y == 0) m |= 1u << 0;
    if (z == 0 && x == 0) m |= 1u << 3;
    if (y == 0 && x == 0) m |= 1u << 8;
    return m;
}

struct CellPos { int z, y, x; };
__device__ __forceinline__ CellPos cell_pos(int64_t c, int H, int W) {   // c = (z*H + y)*W + x
    CellPos p;
    if (c < (1ll << 31)) {   // 32-bit divisions (a 128^3 volume has 2 M cells); the 64-bit ones are emulated and ~5x slower
        const unsigned cu = (unsigned)c, row = cu / (unsigned)W, z = row / (unsigned)H;
        p.x = (int)(cu - row * (unsigned)W);
        p.y = (int)(row - z * (unsigned)H);
        p.z = (int)z;
        return p;
    }
    p.x = (int)(c % W);
    p.y = (int)((c / W) % H);
    p.z = (int)(c / ((int64_t)W * H));
    return p;
}

// cube code word: bits 0-7 cube index, 8-13 face decisions, 14-15 tunnel (0 none, k+1 = annulus k)
__device__ __forceinline__ const McEntry& mc_entry(int code) {
    const int key = (code & 255) * 64 + ((code >> 8) & 63);
    const int tun = code >> 14;
    return tun ? d_mc_tun_table[d_mc_tun_index[key] + tun - 1] : d_mc_table[key];
}
// (triangles | centre vertices << 4) of a code word from the compact count tables
__device__ __forceinline__ unsigned mc_counts(int code) {
    const int key = (code & 255) * 64 + ((code >> 8) & 63);
    const int tun = code >> 14;
    return tun ? d_mc_tun_cnt[d_mc_tun_index[key] + tun - 1] : d_mc_cnt[key];
}

// Interior test of one corner pair (columns A0 A1 B0 B1 C0 C1 D0 D1 of the sweep, see build_tables): every operation
// individually rounded in double (no FMA contraction), the same sequence as oracle/mc_oracle.c.
#ifndef __CUDA_ARCH__   // host build of the same functions (gnb_mc_cell_tiling_host): x86-64 without -mfma never contracts
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
#endif
__host__ __device__ __forceinline__ bool interior_joined(const float* val, float level, const uint8_t* col, double sigma) {
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = __dmul_rn(sigma, __dsub_rn((double)val[col[i]], (double)level));
    const double a0 = c[0], b0 = c[2], c0 = c[4], d0 = c[6];
    const double da = __dsub_rn(c[1], a0), db = __dsub_rn(c[3], b0), dc = __dsub_rn(c[5], c0), dd = __dsub_rn(c[7], d0);
    const double a = __dsub_rn(__dmul_rn(da, dc), __dmul_rn(db, dd));
    const double b = __dsub_rn(__dsub_rn(__dadd_rn(__dmul_rn(c0, da), __dmul_rn(a0, dc)), __dmul_rn(d0, db)), __dmul_rn(b0, dd));
    if (!(a < 0.0)) return false;
    const double t = __ddiv_rn(-b, __dmul_rn(2.0, a));
    if (!(t > 0.0 && t < 1.0)) return false;
    const double At = __dadd_rn(a0, __dmul_rn(da, t)), Bt = __dadd_rn(b0, __dmul_rn(db, t));
    const double Ct = __dadd_rn(c0, __dmul_rn(dc, t)), Dt = __dadd_rn(d0, __dmul_rn(dd, t));
    if (At < 0.0 || Ct < 0.0) return false;
    if (Bt >= 0.0 || Dt >= 0.0) return true;
    return __dsub_rn(__dmul_rn(At, Ct), __dmul_rn(Bt, Dt)) >= (double)FLT_EPSILON;
}

// face test of one ambiguous face given its four corner values (ccw): positive corners connected?
__host__ __device__ __forceinline__ bool face_connected(float v0, float v1, float v2, float v3, float level) {
    const double a = (double)v0 - (double)level, b = (double)v1 - (double)level;
    const double cc = (double)v2 - (double)level, d = (double)v3 - (double)level;
    // asymptotic decider with Lewiner's FLT_EPSILON band: connected iff (product of the positive pair) -
    // (product of the negative pair) > -FLT_EPSILON; fp32 x fp32 products are exact in double.
    const bool pos02 = a > 0.0;
    const double pp = pos02 ? __dmul_rn(a, cc) : __dmul_rn(b, d), nn = pos02 ? __dmul_rn(b, d) : __dmul_rn(a, cc);
    return __dsub_rn(pp, nn) > -(double)FLT_EPSILON;
}

// face / interior decisions of one ACTIVE cell (the rare slow path of the classifier); val[] in corner order
__device__ __noinline__ int mc_resolve(const float* val, float level, int idx) {
    unsigned fb = 0, tun = 0;
    const unsigned am = d_mc_ambig[idx];
    for (int f = 0; f < 6; ++f)
        if ((am >> f) & 1)
            if (face_connected(val[c_face_corner[f][0]], val[c_face_corner[f][1]], val[c_face_corner[f][2]],
                               val[c_face_corner[f][3]], level)) fb |= 1u << f;
    const int ti = d_mc_tun_index[idx * 64 + fb];
    if (ti >= 0) {   // interior ambiguity: first annulus whose regions are joined through the cell
        const McTunDesc& td = d_mc_tun_desc[ti];
        for (int k = 0; k < td.nann && tun == 0; ++k)
            for (int q = 0; q < td.npair[k]; ++q)
                if (interior_joined(val, level, td.col[k][q], (double)td.sigma[k])) { tun = k + 1; break; }
    }
    return idx | (fb << 8) | (tun << 14);
}

// item -> (row = z*H + y, first x of the lane); returns false beyond the last item
struct ItemPos { int z, y, x0; int64_t row; bool ok; };
__device__ __forceinline__ ItemPos item_pos(int64_t item, int lane, int H, const McGeom& g) {
    ItemPos p;
    p.ok = item < g.nitems;
    // 32-bit divisions: a volume has fewer than 2^31 voxels (checked at the entry points), hence fewer items
    const unsigned it = p.ok ? (unsigned)item : 0u, nseg = (unsigned)g.nseg;
    const unsigned row = it / nseg, seg = it - row * nseg, z = row / (unsigned)H;
    p.row = row;
    p.z = (int)z;
    p.y = (int)(row - z * (unsigned)H);
    p.x0 = (int)seg * MC_SEG + lane * 4;
    return p;
}

// ---- kernel 1: classify --------------------------------------------------------------------------------------
// One warp walks a STRIP of MC_STRIP consecutive rows (same slab pair z, z+1, same 128-cell segment) with a two-row window:
// every row is loaded once per slab pair instead of four times, and what is kept of it is one bit per value (v > level).  An
// item whose 4 x 129 corner bits are all equal (three items out of four in a garment volume) is recognised with a handful of
// bit operations and one ballot; the cube indices, table look-ups and counts are computed only for the items the surface
// crosses.  No CTA-wide synchronisation inside the loop: counts are kept per item (one REDUX per item), the data range is
// accumulated in registers and leaves the CTA once.
// VEC: W % 4 == 0 and a 16-byte aligned volume -> float4 loads and 8-byte code stores
constexpr int MC_STRIP = 16;
// nv (<= 128 * 14), nf (<= 128 * 14), na (<= 128) of one item in one 32-bit word
__device__ __forceinline__ unsigned pack3s(int nv, int nf, int na) { return (unsigned)nv | ((unsigned)nf << 11) | ((unsigned)na << 22); }

template <bool VEC>
__global__ void __launch_bounds__(MC_BLOCK, 4)
mc_classify_kernel(const float* __restrict__ vols, int D, int H, int W, float level, McBatch batch) {
    const McWs ws = carve(batch, blockIdx.y);
    const int64_t nvox = batch.g.nvox;
    const float* __restrict__ v = vols + (int64_t)blockIdx.y * nvox;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the per-index table of the cells without ambiguity (99 % of the active ones) lives in shared memory: a global
    // table look-up per active cell made this kernel latency-bound
    __shared__ McFast s_fast[256];
    __shared__ float s_lo[MC_WARPS], s_hi[MC_WARPS];
    s_fast[threadIdx.x] = d_mc_fast[threadIdx.x];
    __syncthreads();
    const unsigned nseg = (unsigned)batch.g.nseg, ygroups = (unsigned)ceil_div(H, MC_STRIP);
    const int64_t nstrips = (int64_t)D * ygroups * nseg;
    const int64_t HW = (int64_t)H * W;
    // data range: every voxel is in row (z, y) of exactly one item
    float lo = INFINITY, hi = -INFINITY;

    // four values of row `p` at this lane's x0 (0 beyond the row)
    auto load4 = [&](const float* __restrict__ p, int x0, float (&q)[4]) {
        if (VEC) {
            if (x0 + 3 < W) {
                const float4 t = *reinterpret_cast<const float4*>(p + x0);
                q[0] = t.x; q[1] = t.y; q[2] = t.z; q[3] = t.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) q[k] = 0.f;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) q[k] = x0 + k < W ? p[x0 + k] : 0.f;
        }
    };
    // (v > level) == ((double)v - (double)level > 0): both operands are exact in double
    auto bits4 = [&](const float (&q)[4]) {
        return (q[0] > level ? 1u : 0u) | (q[1] > level ? 2u : 0u) | (q[2] > level ? 4u : 0u) | (q[3] > level ? 8u : 0u);
    };
    // 5-bit masks of the rows (z, y) and (z+1, y) at this lane (bit 4 = the next lane's first value), packed lo | hi << 8
    auto row_masks = [&](const float* __restrict__ pz, const float* __restrict__ pz1, int x0, const float (&qa)[4], const float (&qb)[4]) {
        const unsigned m4 = bits4(qa) | (bits4(qb) << 8);
        unsigned nb = __shfl_down_sync(0xffffffffu, m4, 1);
        if (lane == 31) nb = x0 + 4 < W ? ((pz[x0 + 4] > level ? 1u : 0u) | (pz1[x0 + 4] > level ? 0x100u : 0u)) : 0u;
        return m4 | ((nb & 0x101u) << 4);
    };

    for (int64_t strip = (int64_t)blockIdx.x * MC_WARPS + warp; strip < nstrips; strip += (int64_t)gridDim.x * MC_WARPS) {
        const unsigned su = (unsigned)strip, seg = su % nseg, t = su / nseg, yg = t % ygroups;
        const int z = (int)(t / ygroups);
        const int y0 = (int)yg * MC_STRIP, y1 = y0 + MC_STRIP < H ? y0 + MC_STRIP : H;
        const int x0 = (int)seg * MC_SEG + lane * 4;
        unsigned valid = 0u;   // cells (x, x+1) that exist along x
#pragma unroll
        for (int k = 0; k < 4; ++k) valid |= (x0 + k < W - 1 ? 1u : 0u) << k;
        const float* __restrict__ rz = v + (int64_t)z * HW;   // slab z
        const bool zin = z < D - 1;
        float qa[4], qb[4];
        load4(rz + (int64_t)y0 * W, x0, qa);
        unsigned mprev = 0u;
        if (zin) {
            load4(rz + HW + (int64_t)y0 * W, x0, qb);
            mprev = row_masks(rz + (int64_t)y0 * W, rz + HW + (int64_t)y0 * W, x0, qa, qb);
        }
        for (int y = y0; y < y1; ++y) {
            const int64_t item = ((int64_t)z * H + y) * nseg + seg;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (x0 + k < W) { lo = fminf(lo, qa[k]); hi = fmaxf(hi, qa[k]); }
            unsigned long long codes = 0ull;   // the lane's four 16-bit code words
            unsigned tot = 0u;
            if (zin && y < H - 1) {   // warp-uniform
                const float* __restrict__ pc = rz + (int64_t)(y + 1) * W;
                load4(pc, x0, qa);
                load4(pc + HW, x0, qb);
                const unsigned mnext = row_masks(pc, pc + HW, x0, qa, qb);
                // rows r = dy + 2 dz: r0 = (z,y) r1 = (z,y+1) r2 = (z+1,y) r3 = (z+1,y+1)
                const unsigned m0 = mprev & 0x1Fu, m2 = mprev >> 8, m1 = mnext & 0x1Fu, m3 = mnext >> 8;
                const unsigned any = m0 | m1 | m2 | m3, all = m0 & m1 & m2 & m3;
                unsigned act = (any | (any >> 1)) & ~(all & (all >> 1)) & valid;   // bit k: cell k is neither empty nor full
                if (__ballot_sync(0xffffffffu, act != 0u) != 0u) {
                    unsigned cnt3 = 0u;
                    // one copy of the per-cell code (a loop over the set bits, not four unrolled bodies): the classifier must
                    // stay within 64 registers for 32 resident warps per SM
#pragma unroll 1
                    while (act) {
                        const int k = __ffs(act) - 1;
                        act &= act - 1u;
                        // corner i at (dx,dy,dz): 0:000 1:100 2:110 3:010 4:001 5:101 6:111 7:011
                        const unsigned b0 = m0 >> k, b1 = m1 >> k, b2 = m2 >> k, b3 = m3 >> k;
                        const int idx = (int)((b0 & 3u) | ((b1 & 2u) << 1) | ((b1 & 1u) << 3) | ((b2 & 3u) << 4) | ((b3 & 2u) << 5) | ((b3 & 1u) << 7));
                        const McFast fe = s_fast[idx];
                        const int code = idx;
                        const unsigned cnt = fe.cnt0;
                        if (fe.resolve) {
                            // ambiguous faces / interior ambiguity (about 1 % of the active cells): the face and interior tests
                            // run in mc_resolve_kernel, which patches the code word and the item's counts; here the cell is only
                            // listed (the `active` area is free until the compaction) and counted with its fb = 0 tiling
                            reinterpret_cast<int*>(ws.active)[atomicAdd(&ws.rec->n_resolve, 1)] = (int)(((int64_t)z * H + y) * W + x0 + k);
                        }
                        codes |= (unsigned long long)(unsigned)code << (16 * k);
                        cnt3 += pack3s(__popc(fe.edgemask & owned_mask(z, y, x0 + k)) + (int)(cnt >> 4), (int)(cnt & 15), 1);
                    }
                    tot = __reduce_add_sync(0xffffffffu, cnt3);
                }
                mprev = mnext;
            } else if (y + 1 < y1) {
                load4(rz + (int64_t)(y + 1) * W, x0, qa);   // last slab: only the data range needs the rows
            }
            {   // code words of the lane's four cells (padding cells / rows: 0)
                uint16_t* __restrict__ dst = ws.codes + ((int64_t)z * H + y) * W + x0;
                if (VEC) {
                    if (x0 + 3 < W) *reinterpret_cast<uint2*>(dst) = make_uint2((unsigned)codes, (unsigned)(codes >> 32));
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (x0 + k < W) dst[k] = (uint16_t)(codes >> (16 * k));
                }
            }
            if (lane == 0) {
                ws.blockV[item] = (int)(tot & 0x7FFu);
                ws.blockF[item] = (int)((tot >> 11) & 0x7FFu);
                ws.blockA[item] = (int)(tot >> 22);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < MC_WARPS; ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
        if (lo <= hi) {   // one atomic pair per CTA
            atomicMin(&ws.rec->min_enc, mc_enc(lo));
            atomicMax(&ws.rec->max_enc, mc_enc(hi));
        }
    }
}

// ---- kernel 1b: face / interior tests of the listed cells ----------------------------------------------------------
// One thread per listed cell: the resolved code word replaces the provisional one (cube index only) and the differences of
// its triangle / centre-vertex counts go to the item's counts with integer atomics (commutative: the result is deterministic).
__global__ void __launch_bounds__(128)
mc_resolve_kernel(const float* __restrict__ vols, int D, int H, int W, float level, McBatch batch) {
    const McWs ws = carve(batch, blockIdx.y);
    const float* __restrict__ v = vols + (int64_t)blockIdx.y * batch.g.nvox;
    const int n = ws.rec->n_resolve;
    const int* __restrict__ list = reinterpret_cast<const int*>(ws.active);
    const int64_t HW = (int64_t)H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int cell = list[i];
        const int idx = ws.codes[cell] & 255;
        const float* __restrict__ c0 = v + cell;
        const float val[8] = {c0[0], c0[1], c0[W + 1], c0[W], c0[HW], c0[HW + 1], c0[HW + W + 1], c0[HW + W]};
        const int code = mc_resolve(val, level, idx);
        if (code == idx) continue;
        ws.codes[cell] = (uint16_t)code;
        const unsigned c_new = mc_counts(code), c_old = d_mc_fast[idx].cnt0;
        const CellPos p = cell_pos(cell, H, W);
        const int64_t item = ((int64_t)p.z * H + p.y) * batch.g.nseg + p.x / MC_SEG;
        const int dv = (int)(c_new >> 4) - (int)(c_old >> 4), df = (int)(c_new & 15) - (int)(c_old & 15);
        if (dv) atomicAdd(&ws.blockV[item], dv);
        if (df) atomicAdd(&ws.blockF[item], df);
    }
}

// ---- kernel 2: one CTA per volume, exclusive scan of the block counts --------------------------------------------
__global__ void __launch_bounds__(1024)
mc_scan_kernel(McBatch batch) {
    const McWs ws = carve(batch, blockIdx.x);
    const int64_t nb = batch.g.nb;
    __shared__ long long wsum[3][32];
    __shared__ long long ctot[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long carry[3] = {0, 0, 0};
    int32_t* arr[3] = {ws.blockV, ws.blockF, ws.blockA};
    for (int64_t base = 0; base < nb; base += 1024) {
        const int64_t i = base + tid;
        long long val[3], incl[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            val[k] = i < nb ? arr[k][i] : 0;
            incl[k] = val[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, incl[k], o);
                if (lane >= o) incl[k] += t;
            }
            if (lane == 31) wsum[k][warp] = incl[k];
        }
        __syncthreads();
        if (warp < 3) {
            const long long w = wsum[warp][lane];
            long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            wsum[warp][lane] = wi - w;
            if (lane == 31) ctot[warp] = wi;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (i < nb) arr[k][i] = (int32_t)(carry[k] + wsum[k][warp] + incl[k] - val[k]);
            carry[k] += ctot[k];
        }
        __syncthreads();
    }
    if (tid == 0) { ws.rec->V = carry[0]; ws.rec->F = carry[1]; ws.rec->A = carry[2]; }
}

// first output row of every volume in the concatenated vertex / face arrays of the batch
__global__ void mc_bases_kernel(McBatch batch, int N) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int64_t v = 0, f = 0;
    for (int i = 0; i < N; ++i) {
        McRec* r = carve(batch, i).rec;
        r->vbase = v; r->fbase = f;
        v += r->V; f += r->F;
    }
}

// ---- kernel 3: compaction of the active cells (order preserving) + the vertex work list ---------------------------
// Same item mapping as the classifier: one warp per item, no CTA-wide synchronisation in the loop.  Besides the active-cell
// records it leaves, in the (still unwritten) output row of every vertex, the pair (cell, vertex slot) that the vertex
// kernel consumes: one thread per VERTEX there, no divergence over cells that own 0..5 vertices.
template <bool VEC>
__global__ void __launch_bounds__(MC_BLOCK, 4)
mc_compact_kernel(int D, int H, int W, McBatch batch, int n_entries, float* __restrict__ verts) {
    // the first-use vertex order of every tiling (one 64-bit word each) and the per-index fast table live in shared memory
    extern __shared__ __align__(16) unsigned long long s_order[];
    __shared__ McFast s_fast[256];
    __shared__ int4 s_cells[MC_WARPS][MC_SEG];   // per warp: the active cells of the item in flight
    for (int i = threadIdx.x; i < n_entries; i += MC_BLOCK) s_order[i] = d_mc_packed[i].order;
    s_fast[threadIdx.x] = d_mc_fast[threadIdx.x];
    __syncthreads();
    const McWs ws = carve(batch, blockIdx.y);
    const int64_t nitems = batch.g.nitems;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_base = d_mc_n_base;
    const int64_t vbase = ws.rec->vbase;
    const int total_active = (int)ws.rec->A;
    for (int64_t item = (int64_t)blockIdx.x * MC_WARPS + warp; item < nitems; item += (int64_t)gridDim.x * MC_WARPS) {
        // every load of the item is issued before the first one is needed (the loop is bound by memory latency): the
        // active-cell offsets that decide whether the item has work, its vertex / face bases and its code words
        const ItemPos ip = item_pos(item, lane, H, batch.g);
        const int a0 = ws.blockA[item];
        const int a1 = item + 1 < nitems ? ws.blockA[item + 1] : total_active;
        const int bv = ws.blockV[item], bf = ws.blockF[item];
        int code[4] = {0, 0, 0, 0}, cv[4] = {0, 0, 0, 0}, cf[4] = {0, 0, 0, 0};
        {
            const uint16_t* __restrict__ src = ws.codes + ip.row * W + ip.x0;
            if (VEC) {
                if (ip.x0 + 3 < W) {
                    const uint2 q = *reinterpret_cast<const uint2*>(src);
                    code[0] = q.x & 0xFFFF; code[1] = q.x >> 16; code[2] = q.y & 0xFFFF; code[3] = q.y >> 16;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (ip.x0 + k < W - 1) code[k] = src[k];
            }
        }
        if (a1 == a0) continue;  // no active cell in this item (warp-uniform)
        // owned-edge mask of the item's cells: the (z, y) part is the same for the whole row, x = 0 adds three edges
        const unsigned own_row = owned_mask(ip.z, ip.y, 1), own_x0 = owned_mask(ip.z, ip.y, 0);
        int nv = 0, nf = 0, na = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = code[k] & 255;
            if (idx != 0 && idx != 255) {
                const McFast fe = s_fast[idx];
                const unsigned cnt = (code[k] >> 8) ? mc_counts(code[k]) : fe.cnt0;
                cf[k] = cnt & 15;
                cv[k] = __popc(fe.edgemask & (ip.x0 + k == 0 ? own_x0 : own_row)) + (cnt >> 4);
                nv += cv[k]; nf += cf[k]; na += 1;
            }
        }
        const unsigned mine = pack3s(nv, nf, na);
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned ex = incl - mine;
        // 1. the item's active cells, compacted into this warp's staging row: {cell, first vertex, first face, code}
        {
            int ov = bv + (int)(ex & 0x7FFu), of = bf + (int)((ex >> 11) & 0x7FFu), j = (int)(ex >> 22);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = code[k] & 255;
                if (idx == 0 || idx == 255) continue;
                s_cells[warp][j++] = make_int4((int)(ip.row * W) + ip.x0 + k, ov, of, code[k]);
                ov += cv[k]; of += cf[k];
            }
        }
        __syncwarp();
        // 2. one lane per active cell: the record of the face kernel (coalesced 16-byte stores) and the vertex work list
        const int n_act = a1 - a0;
        for (int j = lane; j < n_act; j += 32) {
            const int4 c = s_cells[warp][j];
            const int idx = c.w & 255;
            // the face kernel only needs the tiling: its position in the packed table goes into .w
            const int key = idx * 64 + ((c.w >> 8) & 63), tun = c.w >> 14;
            const McFast fe = s_fast[idx];
            const int eid = tun ? n_base + d_mc_tun_index[key] + tun - 1 : ((c.w >> 8) ? d_mc_entry_id[key] : fe.eid0);
            ws.active[a0 + j] = make_int4(c.x, c.y, c.z, eid);
            // vertices this cell creates, in first-use order of its tiling (centres 12, 13 are always the cell's own)
            unsigned long long ow = s_order[eid];
            const int nedge = (int)((ow >> 56) & 15);
            const unsigned own = (fe.edgemask & (c.x == (int)(ip.row * W) ? own_x0 : own_row)) | 0x3000u;
            float* __restrict__ row = verts + (vbase + c.y) * 3;
            for (int q = 0; q < nedge; ++q, ow >>= 4) {
                const unsigned e = (unsigned)ow & 15u;
                if (!((own >> e) & 1u)) continue;
                row[0] = __int_as_float(c.x);
                row[1] = __int_as_float((int)e);
                row += 3;
            }
        }
        __syncwarp();   // the staging row is reused by the warp's next item
    }
}

// ---- kernel 4: vertices --------------------------------------------------------------------------------------
__device__ __forceinline__ float vol_at(const float* __restrict__ v, int D, int H, int W, int z, int y, int x) {
    z = z < 0 ? 0 : (z > D - 1 ? D - 1 : z);
    y = y < 0 ? 0 : (y > H - 1 ? H - 1 : y);
    x = x < 0 ? 0 : (x > W - 1 ? W - 1 : x);
    return v[((int64_t)z * H + y) * W + x];
}

struct Spacing { double s[3]; };

// float32 voxel-unit coordinates (axis0, axis1, axis2) of the iso-vertex on cube edge e of cell p, and the blend t
__device__ __forceinline__ void edge_vertex(const float* __restrict__ v, int H, int W, float level, const CellPos& p,
                                            int e, float c[3], double& t, int lo[3], int hi[3]) {
    const int ax = c_edge_axis[e];
    lo[0] = p.z + c_edge_dz[e]; lo[1] = p.y + c_edge_dy[e]; lo[2] = p.x + c_edge_dx[e];
    hi[0] = lo[0] + (ax == 2); hi[1] = lo[1] + (ax == 1); hi[2] = lo[2] + (ax == 0);
    const float f0 = v[((int64_t)lo[0] * H + lo[1]) * W + lo[2]], f1 = v[((int64_t)hi[0] * H + hi[1]) * W + hi[2]];
    const double a0 = fabs((double)f0 - (double)level), a1 = fabs((double)f1 - (double)level);
    const double w0 = 1.0 / ((double)FLT_EPSILON + a0), w1 = 1.0 / ((double)FLT_EPSILON + a1);
    t = w1 / (w0 + w1);  // (0*w0 + 1*w1)/(w0+w1)
    c[0] = (float)lo[0]; c[1] = (float)lo[1]; c[2] = (float)lo[2];
    const int k = 2 - ax;  // coordinate slot that moves along the edge
    c[k] = (float)((double)lo[k] + t);
}

__device__ __forceinline__ void store_vertex(int64_t row, const float c[3], float g[3], float val, const Spacing& sp,
                                             const float* __restrict__ ggm, int D, int H, int W,
                                             float* __restrict__ verts, float* __restrict__ normals,
                                             float* __restrict__ values, float* __restrict__ ggm_at) {
    const double o0 = (double)c[0] * sp.s[0], o1 = (double)c[1] * sp.s[1], o2 = (double)c[2] * sp.s[2];
    verts[row * 3 + 0] = (float)o0;
    verts[row * 3 + 1] = (float)o1;
    verts[row * 3 + 2] = (float)o2;
    if (ggm_at != nullptr && ggm != nullptr) {
        // predict.py:179-181: (mc_verts / voxel_spacing).astype(np.uint32), float64 arithmetic
        long long i0 = (long long)(o0 / sp.s[0]), i1 = (long long)(o1 / sp.s[1]), i2 = (long long)(o2 / sp.s[2]);
        i0 = i0 < 0 ? 0 : (i0 > D - 1 ? D - 1 : i0);
        i1 = i1 < 0 ? 0 : (i1 > H - 1 ? H - 1 : i1);
        i2 = i2 < 0 ? 0 : (i2 > W - 1 ? W - 1 : i2);
        ggm_at[row] = ggm[(i0 * H + i1) * W + i2];
    }
    if (normals != nullptr) {
        const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(g[0], g[0]), __fmul_rn(g[1], g[1])), __fmul_rn(g[2], g[2]));
        const float nrm = __fsqrt_rn(n2);
        if (nrm > 0.f) { g[0] = __fdiv_rn(g[0], nrm); g[1] = __fdiv_rn(g[1], nrm); g[2] = __fdiv_rn(g[2], nrm); }
        normals[row * 3 + 0] = g[0];
        normals[row * 3 + 1] = g[1];
        normals[row * 3 + 2] = g[2];
    }
    if (values != nullptr) values[row] = val;
}

// one thread per VERTEX (work list left in the vertex rows by the compaction kernel)
__global__ void __launch_bounds__(MC_BLOCK)
mc_vertices_kernel(const float* __restrict__ vols, int D, int H, int W, float level, Spacing sp,
                   const float* __restrict__ ggms, McBatch batch, float* __restrict__ verts, float* __restrict__ normals,
                   float* __restrict__ values, float* __restrict__ ggm_at) {
    const McWs ws = carve(batch, blockIdx.y);
    const int64_t vid = (int64_t)blockIdx.x * MC_BLOCK + threadIdx.x;
    if (vid >= ws.rec->V) return;
    const int64_t vol_n = batch.g.nvox;
    const float* __restrict__ v = vols + (int64_t)blockIdx.y * vol_n;
    const float* __restrict__ ggm = ggms ? ggms + (int64_t)blockIdx.y * vol_n : nullptr;
    const int64_t row = ws.rec->vbase + vid;
    const int cell = __float_as_int(verts[row * 3 + 0]), e = __float_as_int(verts[row * 3 + 1]);
    const CellPos p = cell_pos(cell, H, W);
    if (e >= 12) {
        // centre vertex of a loop / tube fan: mean of the listed iso-vertices (double sum in list order, float32 result)
        const McEntry& en = mc_entry(ws.codes[cell]);
        const int ci = e - 12, n = en.cen_n[ci];
        double acc[3] = {0.0, 0.0, 0.0};
        for (int i = 0; i < n; ++i) {
            float cc[3]; double t; int lo[3], hi[3];
            edge_vertex(v, H, W, level, p, en.cen_loop[ci][i], cc, t, lo, hi);
            acc[0] += (double)cc[0]; acc[1] += (double)cc[1]; acc[2] += (double)cc[2];
        }
        const float cc[3] = {(float)(acc[0] / (double)n), (float)(acc[1] / (double)n), (float)(acc[2] / (double)n)};
        float val[8], vmax = -INFINITY;
        float g[3] = {0.f, 0.f, 0.f};
        if (normals != nullptr || values != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            val[i] = v[((int64_t)(p.z + c_corner_dz[i]) * H + (p.y + c_corner_dy[i])) * W + (p.x + c_corner_dx[i])];
            vmax = fmaxf(vmax, val[i]);
        }
        // cell-centre gradient: mean of the four parallel edge differences per axis
        g[0] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(val[4] - val[0], val[5] - val[1]), val[6] - val[2]), val[7] - val[3]), 0.25f);
        g[1] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(val[3] - val[0], val[2] - val[1]), val[7] - val[4]), val[6] - val[5]), 0.25f);
        g[2] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(val[1] - val[0], val[2] - val[3]), val[5] - val[4]), val[6] - val[7]), 0.25f);
        }
        store_vertex(row, cc, g, vmax, sp, ggm, D, H, W, verts, normals, values, ggm_at);
        ws.edge_map[(int64_t)(3 + ci) * vol_n + cell] = (int)vid;
        return;
    }
    const int ax = c_edge_axis[e];
    float cc[3]; double t; int lo[3], hi[3];
    edge_vertex(v, H, W, level, p, e, cc, t, lo, hi);
    // values: max of the data over the cells that share the edge (local maximum near the vertex)
    float vmax = -INFINITY;
    if (values != nullptr)
    for (int dz = (ax == 2 ? 0 : -1); dz <= 1; ++dz)
        for (int dy = (ax == 1 ? 0 : -1); dy <= 1; ++dy)
            for (int dx = (ax == 0 ? 0 : -1); dx <= 1; ++dx)
                vmax = fmaxf(vmax, vol_at(v, D, H, W, lo[0] + dz, lo[1] + dy, lo[2] + dx));
    // normals: central-difference gradient at the two end points, blended with t, normalised
    float g[3] = {0.f, 0.f, 0.f};
    const float tt = (float)t;
    if (normals != nullptr)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int dz = q == 0, dy = q == 1, dx = q == 2;
        const float g0 = vol_at(v, D, H, W, lo[0] + dz, lo[1] + dy, lo[2] + dx) - vol_at(v, D, H, W, lo[0] - dz, lo[1] - dy, lo[2] - dx);
        const float g1 = vol_at(v, D, H, W, hi[0] + dz, hi[1] + dy, hi[2] + dx) - vol_at(v, D, H, W, hi[0] - dz, hi[1] - dy, hi[2] - dx);
        g[q] = __fadd_rn(__fmul_rn(g0, 1.0f - tt), __fmul_rn(g1, tt));
    }
    store_vertex(row, cc, g, vmax, sp, ggm, D, H, W, verts, normals, values, ggm_at);
    ws.edge_map[(int64_t)ax * vol_n + ((int64_t)lo[0] * H + lo[1]) * W + lo[2]] = (int)vid;
}

// ---- kernel 5: faces -------------------------------------------------------------------------------------------
// One thread per active cell.  The tiling comes as four 64-bit words (McPacked: two 16-byte loads, L1-resident table), every
// triangle corner is a 4-bit vertex id decoded with shifts and looked up in the edge map directly -- no per-byte table reads
// and no locally indexed vertex array (both bound the previous form of this kernel: L1 at 80 %, lg / mio throttle).
constexpr unsigned mc_bits12(const int8_t (&t)[12], int bit) {
    unsigned m = 0;
    for (int e = 0; e < 12; ++e) m |= (unsigned)((t[e] >> bit) & 1) << e;
    return m;
}
constexpr int8_t k_edge_axis[12] = {0, 1, 0, 1, 0, 1, 0, 1, 2, 2, 2, 2};
constexpr int8_t k_edge_dx[12] = {0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 0};
constexpr int8_t k_edge_dy[12] = {0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1};
constexpr int8_t k_edge_dz[12] = {0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0};
constexpr unsigned MC_AX0 = mc_bits12(k_edge_axis, 0), MC_AX1 = mc_bits12(k_edge_axis, 1), MC_DX = mc_bits12(k_edge_dx, 0),
                   MC_DY = mc_bits12(k_edge_dy, 0), MC_DZ = mc_bits12(k_edge_dz, 0);

__global__ void __launch_bounds__(MC_BLOCK)
mc_faces_kernel(int D, int H, int W, int ascent, McBatch batch, int32_t* __restrict__ faces) {
    const McWs ws = carve(batch, blockIdx.y);
    const int64_t a = (int64_t)blockIdx.x * MC_BLOCK + threadIdx.x;
    if (a >= ws.rec->A) return;
    const int4 act = ws.active[a];
    const uint4* __restrict__ pk = reinterpret_cast<const uint4*>(&d_mc_packed[act.w]);
    const uint4 q0 = __ldg(pk), q1 = __ldg(pk + 1);
    const unsigned long long tri[3] = {(unsigned long long)q0.z | ((unsigned long long)q0.w << 32),
                                       (unsigned long long)q1.x | ((unsigned long long)q1.y << 32),
                                       (unsigned long long)q1.z | ((unsigned long long)q1.w << 32)};
    const int nf = (int)(q0.y >> 28);
    if (nf == 0) return;
    const int64_t vol_n = batch.g.nvox;
    const int HW = H * W;
    const int32_t* __restrict__ em = ws.edge_map;
    // vertex id of vertex e of this cell (cube edge: the edge map entry of its lower end point; centre: the cell's own plane)
    auto vid_of = [&](unsigned e) {
        if (e >= 12u) return em[(int64_t)(3 + (e - 12u)) * vol_n + act.x];
        const unsigned ax = ((MC_AX0 >> e) & 1u) | (((MC_AX1 >> e) & 1u) << 1);
        const int off = (int)((MC_DZ >> e) & 1u) * HW + (int)((MC_DY >> e) & 1u) * W + (int)((MC_DX >> e) & 1u);
        return em[(int64_t)ax * vol_n + act.x + off];
    };
    int32_t* __restrict__ out = faces + (ws.rec->fbase + act.z) * 3;
    unsigned long long w = tri[0];
    for (int t = 0; t < nf; ++t) {
        if (t == 5) w = tri[1];
        if (t == 10) w = tri[2];
        const unsigned c3 = (unsigned)w & 0xFFFu;
        w >>= 12;
        const int a0 = vid_of(c3 & 15u), b = vid_of((c3 >> 4) & 15u), cc = vid_of(c3 >> 8);
        // native winding: right-hand normal towards lower values ('descent'); 'ascent' reverses the columns
        out[t * 3 + 0] = ascent ? cc : a0;
        out[t * 3 + 1] = b;
        out[t * 3 + 2] = ascent ? a0 : cc;
    }
}

// resets the per-volume records (totals, bases, data range) before a classification pass
__global__ void mc_init_kernel(McBatch batch, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    McRec* r = carve(batch, i).rec;
    r->V = r->F = r->A = r->vbase = r->fbase = 0;
    r->n_resolve = 0;
    r->min_enc = 0xFFFFFFFFu;
    r->max_enc = 0u;
}

static int32_t count_batch(const float* v, int N, int D, int H, int W, float level, void* ws_, int64_t ws_stride,
                           cudaStream_t st) {
    if (ensure_tables() != 0) { set_error("gnb_mc_count: table upload failed"); return GNB_ERR_CUDA; }
    McBatch b = {reinterpret_cast<char*>(ws_), ws_stride, geom(D, H, W)};
    mc_init_kernel<<<ceil_div(N, 128), 128, 0, st>>>(b, N);
    const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0) && ((reinterpret_cast<uintptr_t>(ws_) & 7) == 0) &&
                     (ws_stride % 8 == 0);
    // one strip per warp: strips differ a lot in cost (the surface crosses some rows and not others), so the grid is left
    // to the block scheduler instead of a fixed striding (a persistent grid idled a quarter of the SMs in the tail)
    const int64_t strip_ctas = ceil_div<int64_t>((int64_t)D * ceil_div(H, MC_STRIP) * b.g.nseg, MC_WARPS);
    const dim3 grid((unsigned)strip_ctas, N);
    if (vec) mc_classify_kernel<true><<<grid, MC_BLOCK, 0, st>>>(v, D, H, W, level, b);
    else mc_classify_kernel<false><<<grid, MC_BLOCK, 0, st>>>(v, D, H, W, level, b);
    mc_resolve_kernel<<<dim3(16, N), 128, 0, st>>>(v, D, H, W, level, b);
    mc_scan_kernel<<<N, 1024, 0, st>>>(b);
    mc_bases_kernel<<<1, 32, 0, st>>>(b, N);
    return check_launch("gnb_mc_count");
}

static int32_t emit_batch(const float* v, int N, int D, int H, int W, float level, const double* spacing_host,
                          int ascent, const float* ggm, void* ws_, int64_t ws_stride, int64_t max_active, int64_t max_verts,
                          float* verts, int32_t* faces, float* normals, float* values, float* ggm_at, cudaStream_t st) {
    McBatch b = {reinterpret_cast<char*>(ws_), ws_stride, geom(D, H, W)};
    if (max_active <= 0 || max_verts <= 0) return GNB_OK;
    if (max_active > b.g.nvox) max_active = b.g.nvox;
    Spacing sp = {{spacing_host[0], spacing_host[1], spacing_host[2]}};
    if (ensure_tables() != 0) { set_error("gnb_mc_emit: table upload failed"); return GNB_ERR_CUDA; }
    const int n_entries = g_mc_n_entries;
    const int tab_smem = n_entries * (int)sizeof(unsigned long long);
    {
        // ~16 CTAs per SM over the batch (four waves of the 4 resident ones): the tables are staged per CTA, so fewer CTAs
        // than items, but enough of them that the block scheduler evens out the volumes
        const int64_t per_vol = ceil_div<int64_t>((int64_t)sm_count() * 16, N);
        const int64_t item_ctas = ceil_div<int64_t>(b.g.nitems, MC_WARPS);
        const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(ws_) & 7) == 0) && (ws_stride % 8 == 0);
        const dim3 grid((unsigned)(item_ctas < per_vol ? item_ctas : per_vol), N);
        if (vec) mc_compact_kernel<true><<<grid, MC_BLOCK, tab_smem, st>>>(D, H, W, b, n_entries, verts);
        else mc_compact_kernel<false><<<grid, MC_BLOCK, tab_smem, st>>>(D, H, W, b, n_entries, verts);
    }
    mc_vertices_kernel<<<dim3((unsigned)ceil_div<int64_t>(max_verts, MC_BLOCK), N), MC_BLOCK, 0, st>>>(
        v, D, H, W, level, sp, ggm, b, verts, normals, values, ggm_at);
    mc_faces_kernel<<<dim3((unsigned)ceil_div<int64_t>(max_active, MC_BLOCK), N), MC_BLOCK, 0, st>>>(D, H, W, ascent, b, faces);
    return check_launch("gnb_mc_emit");
}

}  // namespace gnb

using namespace gnb;

extern "C" {

int64_t gnb_mc_workspace_bytes(int32_t D, int32_t H, int32_t W) {
    if (D < 2 || H < 2 || W < 2) return 512;
    return geom(D, H, W).total;
}

int64_t gnb_mc_totals_offset(int32_t D, int32_t H, int32_t W) {
    if (D < 2 || H < 2 || W < 2) return 0;
    return geom(D, H, W).o_rec;
}

int32_t gnb_mc_count(const float* v, int32_t D, int32_t H, int32_t W, float level, void* ws_, int64_t* counts_host,
                     void* stream) {
    GNB_REQUIRE(v && ws_, "gnb_mc_count: null pointer");
    GNB_REQUIRE(D >= 2 && H >= 2 && W >= 2, "gnb_mc_count: volume must be at least 2x2x2");
    GNB_REQUIRE((int64_t)D * H * W < (1ll << 31), "gnb_mc_count: volumes of 2^31 voxels or more are not supported");
    cudaStream_t st = as_stream(stream);
    const McGeom g = geom(D, H, W);
    int32_t rc = count_batch(v, 1, D, H, W, level, ws_, g.total, st);
    if (rc != GNB_OK) return rc;
    if (counts_host == nullptr) return GNB_OK;  // asynchronous form: the caller reads the record itself
    McRec h;
    GNB_CUDA(cudaMemcpyAsync(&h, reinterpret_cast<char*>(ws_) + g.o_rec, sizeof(h), cudaMemcpyDeviceToHost, st));
    GNB_CUDA(cudaStreamSynchronize(st));
    counts_host[0] = h.V;
    counts_host[1] = h.F;
    counts_host[2] = h.A;
    auto dec = [](unsigned e) { unsigned b = (e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e; float f; memcpy(&f, &b, 4); return f; };
    const float lo = dec(h.min_enc), hi = dec(h.max_enc);
    if (level < lo || level > hi) {
        set_error("Surface level must be within volume data range.");
        return GNB_ERR_NO_SURFACE;
    }
    return GNB_OK;
}

int32_t gnb_mc_emit(const float* v, int32_t D, int32_t H, int32_t W, float level, const double* spacing_host,
                    int32_t ascent, const float* ggm, void* ws_, int64_t n_active, int64_t n_verts, float* verts,
                    int32_t* faces, float* normals, float* values, float* ggm_at_verts, void* stream) {
    GNB_REQUIRE(v && ws_ && verts && faces && spacing_host, "gnb_mc_emit: null pointer");
    GNB_REQUIRE(D >= 2 && H >= 2 && W >= 2, "gnb_mc_emit: volume must be at least 2x2x2");
    const McGeom g = geom(D, H, W);
    return emit_batch(v, 1, D, H, W, level, spacing_host, ascent, ggm, ws_, g.total, n_active, n_verts, verts, faces, normals,
                      values, ggm_at_verts, as_stream(stream));
}

int32_t gnb_mc_count_batch(const float* v, int32_t N, int32_t D, int32_t H, int32_t W, float level, void* ws,
                           int64_t ws_stride, void* stream) {
    GNB_REQUIRE(v && ws, "gnb_mc_count_batch: null pointer");
    GNB_REQUIRE(N >= 1 && N <= 65535, "gnb_mc_count_batch: 1 <= N <= 65535 volumes");
    GNB_REQUIRE(D >= 2 && H >= 2 && W >= 2, "gnb_mc_count_batch: volumes must be at least 2x2x2");
    GNB_REQUIRE((int64_t)D * H * W < (1ll << 31), "gnb_mc_count_batch: volumes of 2^31 voxels or more are not supported");
    GNB_REQUIRE(ws_stride >= geom(D, H, W).total && ws_stride % 256 == 0,
                "gnb_mc_count_batch: ws_stride must be a multiple of 256 >= gnb_mc_workspace_bytes");
    return count_batch(v, N, D, H, W, level, ws, ws_stride, as_stream(stream));
}

int32_t gnb_mc_emit_batch(const float* v, int32_t N, int32_t D, int32_t H, int32_t W, float level,
                          const double* spacing_host, int32_t ascent, const float* ggm, void* ws, int64_t ws_stride,
                          int64_t max_active, int64_t max_verts, float* verts, int32_t* faces, float* normals,
                          float* values, float* ggm_at_verts, void* stream) {
    GNB_REQUIRE(v && ws && verts && faces && spacing_host, "gnb_mc_emit_batch: null pointer");
    GNB_REQUIRE(N >= 1 && N <= 65535, "gnb_mc_emit_batch: 1 <= N <= 65535 volumes");
    GNB_REQUIRE(D >= 2 && H >= 2 && W >= 2, "gnb_mc_emit_batch: volumes must be at least 2x2x2");
    return emit_batch(v, N, D, H, W, level, spacing_host, ascent, ggm, ws, ws_stride, max_active, max_verts, verts, faces,
                      normals, values, ggm_at_verts, as_stream(stream));
}

int32_t gnb_mc_cell_tiling_host(const float* corner_values, float level, int32_t* code_out, int32_t* ntri_out,
                                uint8_t* tri_out, int32_t* nvert_out, uint8_t* order_out, int32_t* ncen_out, uint8_t* cen_n_out,
                                uint8_t* cen_loop_out) {
    GNB_REQUIRE(corner_values && code_out && ntri_out && tri_out && nvert_out && order_out, "gnb_mc_cell_tiling_host: null pointer");
    static McTables* T = nullptr;
    if (!T) { T = new McTables(); build_tables(*T); }
    int idx = 0;
    for (int i = 0; i < 8; ++i)
        if (corner_values[i] > level) idx |= 1 << i;
    unsigned fb = 0, tun = 0;
    for (int f = 0; f < 6; ++f)
        if ((T->ambig[idx] >> f) & 1)
            if (face_connected(corner_values[h_face_corner[f][0]], corner_values[h_face_corner[f][1]],
                               corner_values[h_face_corner[f][2]], corner_values[h_face_corner[f][3]], level)) fb |= 1u << f;
    const int ti = (idx != 0 && idx != 255) ? T->tun_index[idx * 64 + fb] : -1;
    if (ti >= 0) {
        const McTunDesc& td = T->tun_desc[ti];
        for (int k = 0; k < td.nann && tun == 0; ++k)
            for (int q = 0; q < td.npair[k]; ++q)
                if (interior_joined(corner_values, level, td.col[k][q], (double)td.sigma[k])) { tun = k + 1; break; }
    }
    const McEntry& en = tun ? T->tun_table[ti + tun - 1] : T->table[idx * 64 + fb];
    *code_out = idx | (fb << 8) | (tun << 14);
    *ntri_out = en.ntri;
    memcpy(tri_out, en.tri, 3 * MC_MAX_TRI);
    *nvert_out = en.nedge;
    memcpy(order_out, en.order, 12 + MC_MAX_CEN);
    if (ncen_out) *ncen_out = en.ncen;
    if (cen_n_out) memcpy(cen_n_out, en.cen_n, MC_MAX_CEN);
    if (cen_loop_out) memcpy(cen_loop_out, en.cen_loop, MC_MAX_CEN * 12);
    return GNB_OK;
}

}  // extern "C"

#ifdef GNB_MC_TABLE_MAIN
// host-only self check of the generated tables: nvcc -DGNB_MC_TABLE_MAIN marching_cubes.cu capi.cu -o mctab
int main() {
    static gnb::McTables T;
    gnb::build_tables(T);
    int max_tri = 0, max_cen = 0, with_cen = 0, valid = 0, tun_cfg = 0;
    for (int idx = 1; idx < 255; ++idx)
        for (int fb = 0; fb < 64; ++fb) {
            if (fb & ~T.ambig[idx]) continue;
            const gnb::McEntry& e = T.table[idx * 64 + fb];
            ++valid;
            if (e.ntri > max_tri) max_tri = e.ntri;
            if (e.ncen > max_cen) max_cen = e.ncen;
            with_cen += e.ncen > 0;
            tun_cfg += T.tun_index[idx * 64 + fb] >= 0;
        }
    int tmax = 0;
    for (int i = 0; i < T.ntun; ++i) if (T.tun_table[i].ntri > tmax) tmax = T.tun_table[i].ntri;
    printf("entries %d max_tri %d max_cen %d entries_with_centre %d tunnel_configs %d tunnel_entries %d max_tunnel_tri %d\n", valid,
           max_tri, max_cen, with_cen, tun_cfg, T.ntun, tmax);
    return 0;
}
#endif
