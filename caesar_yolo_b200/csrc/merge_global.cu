// Catalog assembly on the device: per-tile JSON-record construction (int truncation + tile offsets + local edge
// flag), edge flagging against neighbour tiles, and the cross-tile overlap merge (graph over edge sources,
// connected components in the reference's DFS order, hull + largest-member attributes).
//
// Replaces Analyzer.make_json_results (caesar_yolo/evaluation.py:418-469), SFinder.find_sources_at_edge
// (caesar_yolo/inference.py:663-726) and SFinder.merge_edge_sources (caesar_yolo/inference.py:731-931) with
// utils.get_merged_bbox (caesar_yolo/utils.py:110-119).  The reference's O(E^2) Python pair loop becomes a
// neighbour-tile-restricted pair search; results (membership, order, tie-breaks) are identical.
#include "merge_global.h"
#include "common.h"

namespace cy {

// ------------------------------------------------------------------------------------------ exclusive scan (int32)
static constexpr int kScanThreads = 256;
static constexpr int kScanItems = 4;
static constexpr int kScanBlock = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int s_w[kScanThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_w[w] = x;
    __syncthreads();
    if (w == 0) {
        int s = lane < kScanThreads / 32 ? s_w[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < kScanThreads / 32) s_w[lane] = s;
    }
    __syncthreads();
    const int base = w ? s_w[w - 1] : 0;
    *total = s_w[kScanThreads / 32 - 1];
    __syncthreads();
    return base + x - v;
}

__global__ void scan_local_kernel(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ sums, int n) {
    const int base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
    int v[kScanItems], s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    int total;
    int ex = block_exclusive_scan(s, &total);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
// single block: exclusive scan of block sums in place; writes grand total to sums[nb]
__global__ void scan_sums_kernel(int* sums, int nb) {
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, &total);
        if (i < nb) sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[nb] = carry;
}
__global__ void scan_add_kernel(int* out, const int* sums, int n, int* total_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += sums[i / kScanBlock];
    if (i == 0 && total_out) *total_out = sums[(n + kScanBlock - 1) / kScanBlock];
}
// out[i] = sum_{k<i} in[k]; out[n] is NOT written; *total (device) receives the sum.  sums: >= n/1024+2 ints.
static void exclusive_scan(const int* in, int* out, int n, int* sums, int* total, cudaStream_t st) {
    if (n <= 0) {
        if (total) cudaMemsetAsync(total, 0, sizeof(int), st);
        return;
    }
    const int nb = (n + kScanBlock - 1) / kScanBlock;
    scan_local_kernel<<<nb, kScanThreads, 0, st>>>(in, out, sums, n);
    scan_sums_kernel<<<1, kScanThreads, 0, st>>>(sums, nb);
    scan_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(out, sums, n, total);
}

// ------------------------------------------------------------------------------------------ records

// evaluation.py:418-469: int() truncation, + tile origin, tile-local edge flag.  One thread per kept detection.
__global__ void make_records_kernel(const float* __restrict__ dets, const int* __restrict__ keep_idx,
                                    const int* __restrict__ nkeep, const int* __restrict__ status, int det_stride,
                                    const cy_tile* __restrict__ tiles, const int* __restrict__ tile_ids, int B,
                                    cy_det_record* __restrict__ recs, int* __restrict__ nrec) {
    const int b = blockIdx.x;
    const int tid = tile_ids[b];
    const cy_tile tl = tiles[tid];
    const int nx = tl.xmax - tl.xmin, ny = tl.ymax - tl.ymin;
    const int n = (status && status[b] != 0) ? 0 : nkeep[b];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float* d = dets + ((long long)b * det_stride + keep_idx[(long long)b * det_stride + i]) * 6;
        const int x1 = (int)d[0], y1 = (int)d[1], x2 = (int)d[2], y2 = (int)d[3];
        int edge = 0;
        if (x1 <= 0 || x1 >= nx - 1 || x2 <= 0 || x2 >= nx - 1) edge = 1;
        if (y1 <= 0 || y1 >= ny - 1 || y2 <= 0 || y2 >= ny - 1) edge = 1;
        cy_det_record r;
        r.x1 = (float)(tl.xmin + x1);
        r.y1 = (float)(tl.ymin + y1);
        r.x2 = (float)(tl.xmin + x2);
        r.y2 = (float)(tl.ymin + y2);
        r.score = d[4];
        r.cls = (int)d[5];
        r.tile_id = tid;
        r.flags = edge;
        recs[(long long)tid * det_stride + i] = r;
    }
    if (threadIdx.x == 0) nrec[tid] = n;
}

__global__ void compact_records_kernel(const cy_det_record* __restrict__ slots, const int* __restrict__ counts,
                                       const int* __restrict__ offs, int slot_stride, cy_det_record* __restrict__ out) {
    const int t = blockIdx.x;
    const int n = counts[t], o = offs[t];
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[o + i] = slots[(long long)t * slot_stride + i];
}

// ------------------------------------------------------------------------------------------ edge flags (a23)

__global__ void edge_flag_kernel(cy_det_record* __restrict__ recs, int n, const cy_tile* __restrict__ tiles,
                                 const int* __restrict__ nb_off, const int* __restrict__ nb_idx,
                                 int* __restrict__ is_edge, int* __restrict__ is_plain) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cy_det_record r = recs[i];
    const cy_tile tl = tiles[r.tile_id];
    int edge = r.flags & 1;
    // inference.py:696-702: touches the tile rectangle (xmax/ymax are the EXCLUSIVE generate_tiles bounds)
    if (r.x1 == (float)tl.xmin || r.x2 == (float)tl.xmax || r.y1 == (float)tl.ymin || r.y2 == (float)tl.ymax) {
        edge = 1;
    } else {
        for (int k = nb_off[r.tile_id]; k < nb_off[r.tile_id + 1]; ++k) {
            const cy_tile nt = tiles[nb_idx[k]];
            if (r.x2 < (float)nt.xmin || r.x1 > (float)nt.xmax || r.y2 < (float)nt.ymin || r.y1 > (float)nt.ymax)
                continue;
            edge = 1;
            break;
        }
    }
    r.flags = (r.flags & ~1) | edge;
    recs[i].flags = r.flags;
    is_edge[i] = edge;
    is_plain[i] = 1 - edge;
}

// V list (edge sources in order) + per-tile first vertex
__global__ void build_vertices_kernel(const cy_det_record* __restrict__ recs, int n, const int* __restrict__ is_edge,
                                      const int* __restrict__ epos, int* __restrict__ vert_rec,
                                      int* __restrict__ tile_vcount) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !is_edge[i]) return;
    vert_rec[epos[i]] = i;
    atomicAdd(&tile_vcount[recs[i].tile_id], 1);
}

__device__ __forceinline__ bool rect_overlap(const cy_det_record& a, const cy_det_record& b) {
    return !(a.x2 < b.x1 || a.x1 > b.x2 || a.y2 < b.y1 || a.y1 > b.y2);
}

// pass 0: degree; pass 1: fill adjacency (ascending vertex order by construction)
template <bool FILL>
__global__ void adjacency_kernel(const cy_det_record* __restrict__ recs, const int* __restrict__ vert_rec, int nv,
                                 const int* __restrict__ tile_vstart, const int* __restrict__ tile_vcount,
                                 const int* __restrict__ nb_off, const int* __restrict__ nb_idx,
                                 int* __restrict__ deg, const int* __restrict__ adj_off, int* __restrict__ adj) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const cy_det_record a = recs[vert_rec[v]];
    int cnt = 0;
    int* dst = FILL ? adj + adj_off[v] : nullptr;
    for (int k = nb_off[a.tile_id]; k < nb_off[a.tile_id + 1]; ++k) {  // neighbour tiles, ascending id
        const int u = nb_idx[k];
        const int s = tile_vstart[u], e = s + tile_vcount[u];
        for (int w = s; w < e; ++w) {
            if (rect_overlap(a, recs[vert_rec[w]])) {
                if (FILL) dst[cnt] = w;
                ++cnt;
            }
        }
    }
    if (!FILL) deg[v] = cnt;
}

// ------------------------------------------------------------------------------------------ components

__device__ __forceinline__ int uf_find(int* parent, int x) {
    int p = parent[x];
    while (p != x) {
        const int g = parent[p];
        if (g != p) parent[x] = g;  // path halving (benign race)
        x = p;
        p = parent[x];
    }
    return x;
}
__global__ void uf_init_kernel(int* parent, int nv) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nv) parent[v] = v;
}
__global__ void uf_union_kernel(int* parent, int nv, const int* __restrict__ adj_off, const int* __restrict__ adj) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    for (int k = adj_off[v]; k < adj_off[v + 1]; ++k) {
        int a = v, b = adj[k];
        if (b < a) continue;  // each undirected edge once
        while (true) {
            a = uf_find(parent, a);
            b = uf_find(parent, b);
            if (a == b) break;
            if (a > b) {
                const int t = a;
                a = b;
                b = t;
            }
            if (atomicCAS(&parent[b], b, a) == b) break;  // hook larger root under smaller
        }
    }
}
__global__ void uf_flatten_kernel(int* parent, int nv, int* __restrict__ is_root, int* __restrict__ comp_size) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const int r = uf_find(parent, v);
    parent[v] = r;
    is_root[v] = (r == v);
    atomicAdd(&comp_size[r], 1);
}

// Per-component statistics without walking the graph: largest member area, how many members reach it, the smallest
// vertex id among those, and the hull.  When the largest area is unique (the common case) the reference's winner
// ("first member in DFS preorder with strictly largest area") is that member whatever the DFS order, so giant
// components (50 % overlap chains thousands of boxes) need no serial walk; ties fall back to the DFS replay.
__device__ __forceinline__ int f2ord(float f) {
    const int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ unsigned long long rec_area(const cy_det_record& r) {
    const long long a = ((long long)r.x2 - (long long)r.x1) * ((long long)r.y2 - (long long)r.y1);
    return (unsigned long long)(a + (1ll << 62));  // order-preserving for negative areas as well
}
__global__ void comp_stats_init_kernel(unsigned long long* amax, int* tie, int* bestv, int* hull, int nv) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    amax[v] = 0ull;
    tie[v] = 0;
    bestv[v] = 0x7fffffff;
    hull[v] = 0x7fffffff;
    hull[nv + v] = 0x7fffffff;
    hull[2 * nv + v] = (int)0x80000000;
    hull[3 * nv + v] = (int)0x80000000;
}
__global__ void comp_stats_kernel(const cy_det_record* __restrict__ recs, const int* __restrict__ vert_rec, int nv,
                                  const int* __restrict__ parent, unsigned long long* amax, int* hull) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const cy_det_record r = recs[vert_rec[v]];
    const int root = parent[v];
    atomicMax(&amax[root], rec_area(r));
    atomicMin(&hull[root], f2ord(r.x1));
    atomicMin(&hull[nv + root], f2ord(r.y1));
    atomicMax(&hull[2 * nv + root], f2ord(r.x2));
    atomicMax(&hull[3 * nv + root], f2ord(r.y2));
}
__global__ void comp_tie_kernel(const cy_det_record* __restrict__ recs, const int* __restrict__ vert_rec, int nv,
                                const int* __restrict__ parent, const unsigned long long* __restrict__ amax, int* tie,
                                int* bestv) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const int root = parent[v];
    if (rec_area(recs[vert_rec[v]]) == amax[root]) {
        atomicAdd(&tie[root], 1);
        atomicMin(&bestv[root], v);
    }
}

// One thread per component root: recursive-DFS emulation (graph.py:9-23) with per-vertex adjacency cursors.
// Winner = first member in DFS preorder with strictly largest (x2-x1)*(y2-y1) (inference.py:838-851), hull bbox.
__global__ void component_kernel(const cy_det_record* __restrict__ recs, const int* __restrict__ vert_rec, int nv,
                                 const int* __restrict__ parent, const int* __restrict__ is_root,
                                 const int* __restrict__ comp_rank, const int* __restrict__ comp_size,
                                 const int* __restrict__ stack_off, const int* __restrict__ adj_off,
                                 const int* __restrict__ adj, int* __restrict__ cursor, unsigned char* __restrict__ visited,
                                 int* __restrict__ stack, int n_plain, const int* __restrict__ tie,
                                 const int* __restrict__ bestv, const int* __restrict__ hull,
                                 cy_source* __restrict__ out) {
    const int v0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (v0 >= nv || !is_root[v0]) return;
    cy_source s;
    const cy_det_record r0 = recs[vert_rec[v0]];
    if (comp_size[v0] == 1) {
        s.x1 = r0.x1; s.y1 = r0.y1; s.x2 = r0.x2; s.y2 = r0.y2; s.score = r0.score; s.cls = r0.cls;
        s.flags = (r0.flags & 1);  // edge stays as flagged, merged = False
        s.tile_id = r0.tile_id;
        out[n_plain + comp_rank[v0]] = s;
        return;
    }
    if (tie[v0] == 1) {  // unique largest member: no walk needed
        const cy_det_record rb = recs[vert_rec[bestv[v0]]];
        s.x1 = ord2f(hull[v0]); s.y1 = ord2f(hull[nv + v0]); s.x2 = ord2f(hull[2 * nv + v0]);
        s.y2 = ord2f(hull[3 * nv + v0]);
        s.score = rb.score; s.cls = rb.cls;
        s.flags = 3;
        s.tile_id = -1;
        out[n_plain + comp_rank[v0]] = s;
        return;
    }
    int* stk = stack + stack_off[v0];
    int sp = 0;
    long long best_area = -1;
    int best = -1;
    float hx1 = r0.x1, hy1 = r0.y1, hx2 = r0.x2, hy2 = r0.y2;
    auto visit = [&](int v) {
        visited[v] = 1;
        const cy_det_record r = recs[vert_rec[v]];
        const long long area = ((long long)r.x2 - (long long)r.x1) * ((long long)r.y2 - (long long)r.y1);
        if (area > best_area) {
            best_area = area;
            best = v;
        }
        hx1 = fminf(hx1, r.x1); hy1 = fminf(hy1, r.y1); hx2 = fmaxf(hx2, r.x2); hy2 = fmaxf(hy2, r.y2);
        stk[sp++] = v;
    };
    visit(v0);
    while (sp > 0) {
        const int u = stk[sp - 1];
        int c = cursor[u];
        const int e = adj_off[u + 1];
        while (c < e && visited[adj[c]]) ++c;
        if (c >= e) {
            cursor[u] = c;
            --sp;
            continue;
        }
        const int nxt = adj[c];
        cursor[u] = c + 1;
        visit(nxt);
    }
    const cy_det_record rb = recs[vert_rec[best]];
    s.x1 = hx1; s.y1 = hy1; s.x2 = hx2; s.y2 = hy2; s.score = rb.score; s.cls = rb.cls;
    s.flags = 3;  // edge | merged
    s.tile_id = -1;
    out[n_plain + comp_rank[v0]] = s;
}

__global__ void plain_out_kernel(const cy_det_record* __restrict__ recs, int n, const int* __restrict__ is_plain,
                                 const int* __restrict__ ppos, cy_source* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !is_plain[i]) return;
    const cy_det_record r = recs[i];
    cy_source s;
    s.x1 = r.x1; s.y1 = r.y1; s.x2 = r.x2; s.y2 = r.y2; s.score = r.score; s.cls = r.cls; s.flags = 0;
    s.tile_id = r.tile_id;
    out[ppos[i]] = s;
}

__global__ void init_cursor_kernel(int* cursor, const int* adj_off, int nv) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nv) cursor[v] = adj_off[v];
}
__global__ void masked_size_kernel(const int* is_root, const int* comp_size, int* out, int nv) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nv) out[v] = is_root[v] ? comp_size[v] : 0;
}
__global__ void set_last_kernel(int* arr, int idx, const int* total) { arr[idx] = *total; }
__global__ void final_count_kernel(const int* n_plain, const int* n_comp, long long* nout) {
    *nout = (long long)*n_plain + *n_comp;
}

// ------------------------------------------------------------------------------------------ host orchestration

int make_records(const float* dets, const int* keep_idx, const int* nkeep, const int* status, int det_stride,
                 const cy_tile* tiles, const int* tile_ids, int B, cy_det_record* recs, int* nrec, cudaStream_t st) {
    make_records_kernel<<<B, 128, 0, st>>>(dets, keep_idx, nkeep, status, det_stride, tiles, tile_ids, B, recs, nrec);
    return (int)cudaGetLastError();
}

size_t compact_scratch_bytes(int T) { return ((size_t)T + T / kScanBlock + 8) * sizeof(int); }

int compact_records(const cy_det_record* slots, const int* counts, int T, int slot_stride, cy_det_record* out,
                    int* total, void* scratch, cudaStream_t st) {
    int* offs = (int*)scratch;
    int* sums = offs + T;
    exclusive_scan(counts, offs, T, sums, total, st);
    if (T > 0) compact_records_kernel<<<T, 128, 0, st>>>(slots, counts, offs, slot_stride, out);
    return (int)cudaGetLastError();
}

int merge_global(cy_det_record* recs, int n, const cy_tile* tiles, int T, const int* nb_off, const int* nb_idx,
                 cy_source* out, long long* nout, cudaStream_t st) {
    if (n == 0) {
        cudaMemsetAsync(nout, 0, sizeof(long long), st);
        return 0;
    }
    // workspace (stream-ordered allocation keeps the entry point allocation-free for callers); keep freed blocks in
    // the pool instead of returning them to the driver at every synchronisation
    static std::atomic<unsigned long long> pool_set{0};
    if (first_use_on_device(pool_set)) {
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    const int nsums = n / kScanBlock + T / kScanBlock + 16;
    int *is_edge, *is_plain, *epos, *ppos, *vert_rec, *tile_vcount, *tile_vstart, *deg, *adj_off, *parent, *is_root,
        *comp_size, *comp_rank, *stack_off, *cursor, *stack, *sums, *scalars, *msize;
    unsigned char* visited;
    size_t ints = (size_t)n * 24 + (size_t)T * 2 + nsums + 64;  // 23 n-sized arrays + adj_off[n+1]
    int* ws;
    if (cudaMallocAsync(&ws, ints * sizeof(int) + (size_t)n + 16, st) != cudaSuccess) return -3;
    int* p = ws;
    is_edge = p; p += n;
    is_plain = p; p += n;
    epos = p; p += n;
    ppos = p; p += n;
    vert_rec = p; p += n;
    deg = p; p += n;
    adj_off = p; p += n + 1;
    parent = p; p += n;
    is_root = p; p += n;
    comp_size = p; p += n;
    comp_rank = p; p += n;
    stack_off = p; p += n;
    cursor = p; p += n;
    stack = p; p += n;
    msize = p; p += n;
    int* tie = p; p += n;
    int* bestv = p; p += n;
    int* hull = p; p += 4 * (size_t)n;
    p += ((size_t)(p - ws) & 1);  // 8-byte alignment for the 64-bit area array
    unsigned long long* amax = (unsigned long long*)p; p += 2 * (size_t)n;
    tile_vcount = p; p += T;
    tile_vstart = p; p += T;
    sums = p; p += nsums;
    scalars = p; p += 8;  // [0]=n_edge [1]=n_plain [2]=n_adj [3]=n_comp [4]=stack total
    visited = (unsigned char*)(ws + ints);
    const int nb = (n + 255) / 256;
    cudaMemsetAsync(tile_vcount, 0, (size_t)T * sizeof(int), st);
    edge_flag_kernel<<<nb, 256, 0, st>>>(recs, n, tiles, nb_off, nb_idx, is_edge, is_plain);
    exclusive_scan(is_edge, epos, n, sums, scalars + 0, st);
    exclusive_scan(is_plain, ppos, n, sums, scalars + 1, st);
    plain_out_kernel<<<nb, 256, 0, st>>>(recs, n, is_plain, ppos, out);
    build_vertices_kernel<<<nb, 256, 0, st>>>(recs, n, is_edge, epos, vert_rec, tile_vcount);
    exclusive_scan(tile_vcount, tile_vstart, T, sums, nullptr, st);
    // the number of vertices is data dependent: bring it to the host (one small sync per mosaic)
    int h[2];
    cudaMemcpyAsync(h, scalars, 2 * sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    const int nv = h[0], n_plain = h[1];
    int rc = 0;
    if (nv > 0) {
        const int vb = (nv + 255) / 256;
        adjacency_kernel<false><<<vb, 256, 0, st>>>(recs, vert_rec, nv, tile_vstart, tile_vcount, nb_off, nb_idx, deg,
                                                    nullptr, nullptr);
        exclusive_scan(deg, adj_off, nv, sums, scalars + 2, st);
        set_last_kernel<<<1, 1, 0, st>>>(adj_off, nv, scalars + 2);
        int nadj;
        cudaMemcpyAsync(&nadj, scalars + 2, sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        int* adj = nullptr;
        if (cudaMallocAsync(&adj, ((size_t)nadj + 1) * sizeof(int), st) != cudaSuccess) {
            cudaFreeAsync(ws, st);
            return -3;
        }
        adjacency_kernel<true><<<vb, 256, 0, st>>>(recs, vert_rec, nv, tile_vstart, tile_vcount, nb_off, nb_idx, deg,
                                                   adj_off, adj);
        uf_init_kernel<<<vb, 256, 0, st>>>(parent, nv);
        uf_union_kernel<<<vb, 256, 0, st>>>(parent, nv, adj_off, adj);
        cudaMemsetAsync(comp_size, 0, (size_t)nv * sizeof(int), st);
        uf_flatten_kernel<<<vb, 256, 0, st>>>(parent, nv, is_root, comp_size);
        exclusive_scan(is_root, comp_rank, nv, sums, scalars + 3, st);
        masked_size_kernel<<<vb, 256, 0, st>>>(is_root, comp_size, msize, nv);
        exclusive_scan(msize, stack_off, nv, sums, scalars + 4, st);
        init_cursor_kernel<<<vb, 256, 0, st>>>(cursor, adj_off, nv);
        cudaMemsetAsync(visited, 0, (size_t)nv, st);
        comp_stats_init_kernel<<<vb, 256, 0, st>>>(amax, tie, bestv, hull, nv);
        comp_stats_kernel<<<vb, 256, 0, st>>>(recs, vert_rec, nv, parent, amax, hull);
        comp_tie_kernel<<<vb, 256, 0, st>>>(recs, vert_rec, nv, parent, amax, tie, bestv);
        component_kernel<<<vb, 256, 0, st>>>(recs, vert_rec, nv, parent, is_root, comp_rank, comp_size, stack_off,
                                             adj_off, adj, cursor, visited, stack, n_plain, tie, bestv, hull, out);
        final_count_kernel<<<1, 1, 0, st>>>(scalars + 1, scalars + 3, nout);
        cudaFreeAsync(adj, st);
    } else {
        cudaMemsetAsync(scalars + 3, 0, sizeof(int), st);
        final_count_kernel<<<1, 1, 0, st>>>(scalars + 1, scalars + 3, nout);
    }
    rc = (int)cudaGetLastError();
    cudaFreeAsync(ws, st);
    return rc;
}

}  // namespace cy
