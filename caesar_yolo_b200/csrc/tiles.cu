// Host-side tile grid + neighbour lists (native replacement of the reference's O(T^2) Python double loop).
// Reference: utils.generate_tiles (caesar_yolo/utils.py:622-697), SFinder.create_tile_tasks neighbour search
// (caesar_yolo/inference.py:1034-1071) with TileTask.is_task_tile_{adjacent,overlapping,neighbor} (:123-163).
#include "common.h"
#include <math.h>
#include <vector>

static bool adjacent1d(int amin, int amax, int bmin, int bmax) {
    return amax == bmin - 1 || amin == bmax + 1 || (amin == bmin && amax == bmax);
}
static bool overlap1d(int amin, int amax, int bmin, int bmax) { return !(amax < bmin) && !(amin > bmax); }

extern "C" int cy_generate_tiles(int img_xmin, int img_xmax, int img_ymin, int img_ymax, int tile_x, int tile_y,
                                 double step_x, double step_y, cy_tile* tiles_host, int capacity, int* ntiles) {
    if (img_xmax <= img_xmin || img_ymax <= img_ymin) return cy::set_error(CY_ERR_INVALID, "xmax/ymax must be > xmin/ymin");
    if (tile_x <= 0 || tile_y <= 0) return cy::set_error(CY_ERR_INVALID, "invalid tile size");
    if (step_x <= 0 || step_y <= 0 || step_x > 1 || step_y > 1) return cy::set_error(CY_ERR_INVALID, "invalid grid step");
    const int Nx = img_xmax - img_xmin + 1, Ny = img_ymax - img_ymin + 1;
    if (tile_x > Nx || tile_y > Ny) return cy::set_error(CY_ERR_INVALID, "tile larger than image");
    const int sx = (int)nearbyint(step_x * tile_x), sy = (int)nearbyint(step_y * tile_y);  // np.round: half to even
    if (sx <= 0 || sy <= 0) return cy::set_error(CY_ERR_INVALID, "grid step rounds to zero");
    std::vector<int> x0, x1, y0, y1;
    for (int iy = 0; iy <= Ny; iy += sy) {
        const int off = std::min(tile_y, Ny - iy);
        if (iy >= Ny || off == 0) break;
        y0.push_back(iy);
        y1.push_back(iy + off);
    }
    for (int ix = 0; ix <= Nx; ix += sx) {
        const int off = std::min(tile_x, Nx - ix);
        if (ix >= Nx || off == 0) break;
        x0.push_back(ix);
        x1.push_back(ix + off);
    }
    const long long T = (long long)x0.size() * y0.size();
    *ntiles = (int)T;
    if (!tiles_host) return CY_OK;  // size query
    if (T > capacity) return cy::set_error(CY_ERR_INVALID, "tile capacity %d < %lld", capacity, T);
    int k = 0;
    for (size_t j = 0; j < y0.size(); ++j)
        for (size_t i = 0; i < x0.size(); ++i) {
            cy_tile t;
            t.xmin = img_xmin + x0[i]; t.xmax = img_xmin + x1[i]; t.ymin = img_ymin + y0[j]; t.ymax = img_ymin + y1[j];
            tiles_host[k++] = t;
        }
    return CY_OK;
}

// CSR neighbour lists, ascending tile id, self excluded.  nb_off_host has T+1 entries.  If nb_idx_host is null only
// the offsets/total are computed (size query).
extern "C" int cy_tile_neighbors(const cy_tile* tiles_host, int T, int* nb_off_host, int* nb_idx_host, int capacity,
                                 int* total) {
    if (T <= 0) return cy::set_error(CY_ERR_INVALID, "no tiles");
    // detect the row-major grid structure produced by generate_tiles
    int nx = 1;
    while (nx < T && tiles_host[nx].ymin == tiles_host[0].ymin && tiles_host[nx].ymax == tiles_host[0].ymax) ++nx;
    bool grid = (T % nx == 0);
    const int ny = grid ? T / nx : 0;
    for (int j = 0; grid && j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const cy_tile& t = tiles_host[j * nx + i];
            if (t.xmin != tiles_host[i].xmin || t.xmax != tiles_host[i].xmax || t.ymin != tiles_host[j * nx].ymin ||
                t.ymax != tiles_host[j * nx].ymax) {
                grid = false;
                break;
            }
        }
    long long cnt = 0;
    auto emit = [&](int i, int j) {
        if (nb_idx_host) {
            if (cnt < capacity) nb_idx_host[cnt] = j;
        }
        ++cnt;
    };
    if (grid) {
        std::vector<unsigned char> ax((size_t)nx * nx), ox((size_t)nx * nx), ay((size_t)ny * ny), oy((size_t)ny * ny);
        for (int a = 0; a < nx; ++a)
            for (int b = 0; b < nx; ++b) {
                const cy_tile &p = tiles_host[a], &q = tiles_host[b];
                ax[(size_t)a * nx + b] = adjacent1d(p.xmin, p.xmax, q.xmin, q.xmax);
                ox[(size_t)a * nx + b] = overlap1d(p.xmin, p.xmax, q.xmin, q.xmax);
            }
        for (int a = 0; a < ny; ++a)
            for (int b = 0; b < ny; ++b) {
                const cy_tile &p = tiles_host[(size_t)a * nx], &q = tiles_host[(size_t)b * nx];
                ay[(size_t)a * ny + b] = adjacent1d(p.ymin, p.ymax, q.ymin, q.ymax);
                oy[(size_t)a * ny + b] = overlap1d(p.ymin, p.ymax, q.ymin, q.ymax);
            }
        // per-axis candidate lists
        std::vector<std::vector<int>> cx(nx), cyv(ny);
        for (int a = 0; a < nx; ++a)
            for (int b = 0; b < nx; ++b)
                if (ax[(size_t)a * nx + b] || ox[(size_t)a * nx + b]) cx[a].push_back(b);
        for (int a = 0; a < ny; ++a)
            for (int b = 0; b < ny; ++b)
                if (ay[(size_t)a * ny + b] || oy[(size_t)a * ny + b]) cyv[a].push_back(b);
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int t = j * nx + i;
                nb_off_host[t] = (int)cnt;
                for (int jj : cyv[j])
                    for (int ii : cx[i]) {
                        const int u = jj * nx + ii;
                        if (u == t) continue;
                        const bool adj = ax[(size_t)i * nx + ii] && ay[(size_t)j * ny + jj];
                        const bool ov = ox[(size_t)i * nx + ii] && oy[(size_t)j * ny + jj];
                        if (adj || ov) emit(t, u);
                    }
            }
    } else {
        for (int t = 0; t < T; ++t) {
            nb_off_host[t] = (int)cnt;
            const cy_tile& p = tiles_host[t];
            for (int u = 0; u < T; ++u) {
                if (u == t) continue;
                const cy_tile& q = tiles_host[u];
                const bool adj = adjacent1d(p.xmin, p.xmax, q.xmin, q.xmax) && adjacent1d(p.ymin, p.ymax, q.ymin, q.ymax);
                const bool ov = overlap1d(p.xmin, p.xmax, q.xmin, q.xmax) && overlap1d(p.ymin, p.ymax, q.ymin, q.ymax);
                if (adj || ov) emit(t, u);
            }
        }
    }
    nb_off_host[T] = (int)cnt;
    *total = (int)cnt;
    if (nb_idx_host && cnt > capacity) return cy::set_error(CY_ERR_INVALID, "neighbour capacity %d < %lld", capacity, cnt);
    return CY_OK;
}
