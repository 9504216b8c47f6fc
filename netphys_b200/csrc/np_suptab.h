// netphys_b200/csrc/np_suptab.h -- host side of the support-direction tables (device side and the argument: np_narrow.cuh, support_tab).
//
// For one shape (its vertex list as the reference's PhysUtil_GetPointFurthestInDirection scans it, physics_util.cpp:5-35) build a cube
// map of directions: 6 faces x N x N cells, each cell the ascending list of every vertex that can be the scan's winner for SOME
// direction in the cell, given that binary32 rounding can move a vertex's score by at most mu/2 (in units of |direction|).
// Conservative by construction: every cell is covered by S x S sample directions c with a covering radius rho (chord length between
// unit vectors; radial projection of the cube face onto the unit sphere is a contraction, so the half-diagonal of a sub-cell on the
// face bounds it); u . p changes by at most rho * r over a sub-cell, so  { k : c . p_k >= max_j c . p_j - (mu + 2 rho r) }  contains the
// candidates of every direction the sample covers.  The same bound with the whole cell's covering radius first narrows the vertices
// to those that matter anywhere in the cell (the winner of every direction of the cell is among them), so the fine samples only
// score a dozen vertices.  Evaluated in double precision with extra slack; the extra dilation eps_dir covers the device's own
// rounding when it picks a cell.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace np {

struct SupTable {
    int log_n = 0;                       // N = 1 << log_n cells per face edge
    float tlim = 0.0f;                   // usable while |t|_inf <= tlim (np_narrow.cuh: tab_arm)
    std::vector<uint64_t> cells;         // 6 * N * N: bytes 0..6 candidate vertex indices (ascending), byte 7 = count (0xff: scan)
    double mean_candidates = 0.0; int scan_cells = 0;
};

// false: the shape gets no table (fewer than 5 or more than 255 vertices, degenerate or non-finite coordinates)
inline bool build_support_table(const std::vector<float>& xyz, SupTable& out)
{
    const int n = (int)(xyz.size() / 3);
    if (n < 5 || n > 255) return false;
    double r = 0.0;
    for (int k = 0; k < n; k++) {
        const double x = xyz[3 * k], y = xyz[3 * k + 1], z = xyz[3 * k + 2];
        if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) return false;
        r = std::fmax(r, std::sqrt(x * x + y * y + z * z));
    }
    if (!(r >= 1e-6 && r <= 1e6)) return false;
    const int log_n = n <= 24 ? 4 : (n <= 80 ? 5 : 6), N = 1 << log_n;
    const int S = (n <= 80 ? 256 : 1024) / N;                        // samples per cell edge
    // Rounding margin the table is built for, and what it allows (np_narrow.cuh, support_tab): 2 delta / |R^T d| <= mu with
    // delta = 7.001 u |d|_1 (1.001 r + |t|_inf), |d|_1 <= sqrt(3) |d|, |R^T d| >= 0.9985 |d|, u = 2^-24:
    //   r + |t|_inf <= mu / (24.32 * 2^-24) = mu * 689 800;  the table promises two thirds of that.
    const double mu = r / 256.0;
    const double reach = mu * 460000.0;
    const double eps_dir = 2e-5;                                     // cell-picking error of the device, as a chord length (it is < 1e-6)
    const double cell_edge = 2.0 / (double)N, sub_edge = cell_edge / (double)S;
    const double rho0 = 0.70710678118654757 * cell_edge * (1.0 + 1e-9) + eps_dir;      // covering radius of a whole cell around its centre
    const double rho = 0.70710678118654757 * sub_edge * (1.0 + 1e-9) + eps_dir;        // ... of one sample
    const double slack0 = (mu + 2.0 * rho0 * r) * (1.0 + 1e-9) + 1e-12 * r;
    const double slack = (mu + 2.0 * rho * r) * (1.0 + 1e-9) + 1e-12 * r;
    out.log_n = log_n;
    out.tlim = (float)((reach - r) * 0.999999);
    out.cells.assign((size_t)6 * N * N, 0);
    std::vector<double> score((size_t)n);
    std::vector<unsigned char> in((size_t)n);
    std::vector<int> pre; pre.reserve((size_t)n);
    double total = 0.0; int scans = 0;
    auto direction = [](int f, double a, double b, double* v) {
        const int axis = f >> 1;
        v[axis] = (f & 1) ? -1.0 : 1.0; v[(axis + 1) % 3] = a; v[(axis + 2) % 3] = b;
        const double inv = 1.0 / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        v[0] *= inv; v[1] *= inv; v[2] *= inv;
    };
    for (int f = 0; f < 6; f++) {
        for (int ib = 0; ib < N; ib++) for (int ia = 0; ia < N; ia++) {
            // every vertex that can win, or come within mu of winning, anywhere in the cell (the winner of each direction included)
            double v[3];
            direction(f, -1.0 + cell_edge * ((double)ia + 0.5), -1.0 + cell_edge * ((double)ib + 0.5), v);
            double best = -1e300;
            for (int k = 0; k < n; k++) { score[(size_t)k] = v[0] * xyz[3 * k] + v[1] * xyz[3 * k + 1] + v[2] * xyz[3 * k + 2]; best = std::fmax(best, score[(size_t)k]); }
            pre.clear();
            for (int k = 0; k < n; k++) if (score[(size_t)k] >= best - slack0) pre.push_back(k);
            std::fill(in.begin(), in.end(), 0);
            if (pre.size() == 1) in[(size_t)pre[0]] = 1;
            else for (int sb = 0; sb < S; sb++) for (int sa = 0; sa < S; sa++) {
                direction(f, -1.0 + cell_edge * (double)ia + sub_edge * ((double)sa + 0.5), -1.0 + cell_edge * (double)ib + sub_edge * ((double)sb + 0.5), v);
                best = -1e300;
                for (int k : pre) { score[(size_t)k] = v[0] * xyz[3 * k] + v[1] * xyz[3 * k + 1] + v[2] * xyz[3 * k + 2]; best = std::fmax(best, score[(size_t)k]); }
                for (int k : pre) if (score[(size_t)k] >= best - slack) in[(size_t)k] = 1;
            }
            int cnt = 0; uint64_t word = 0;
            for (int k = 0; k < n; k++) if (in[(size_t)k]) { if (cnt < 7) word |= (uint64_t)k << (8 * cnt); cnt++; }
            if (cnt > 7) { word = 0xffull << 56; scans++; } else { word |= (uint64_t)cnt << 56; total += cnt; }
            out.cells[((size_t)f * N + ib) * N + ia] = word;
        }
    }
    out.scan_cells = scans;
    out.mean_candidates = total / std::fmax(1.0, (double)(6 * N * N - scans));
    return true;
}

}  // namespace np
