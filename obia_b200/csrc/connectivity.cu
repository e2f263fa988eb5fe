// K3: enforce connectivity (replaces skimage
// `_enforce_label_connectivity_cython`, reached from
// obia/segmentation/segment_boundaries.py:51).
//
// The reference is one sequential raster scan with a BFS per component.  The
// same result is produced here by data-parallel phases (host model with the
// proof-by-test of equivalence: tests/cc_parallel_model.py):
//   1. union-find connected components (4-connectivity, equal label) with
//      min-index roots: T[p] = first raster pixel of p's component;
//   2. components larger than max_size: one thread per component replays the
//      capped BFS and splits it into pieces (T[p] = piece start);
//   3. pieces smaller than min_size: one thread per piece replays the BFS of
//      the reference to find `adjacent` = last-seen neighbour that was already
//      labelled when the scan reached the piece (T[q] < t), iterated to a
//      fixed point for the start_label=1 corner case (label 0 == mask label);
//   4. kept pieces are numbered by the rank of their start pixel (prefix sum),
//      merged pieces follow their adjacent chain.
#include <stdlib.h>

#include "common.cuh"

namespace obia {

constexpr int32_t kTInf = 0x7fffffff;
constexpr int kBitChunk = 1024;    // bitmap words per numbering block (32768 pixels)

// Workspace.  `T` is the union-find parent array first and the piece-start array afterwards (every
// pixel points at the first raster pixel of its piece); `psize` is only meaningful at piece starts;
// `fin` (final label per piece start) shares memory with the BFS queue, which is idle by then;
// `roots` (list of component roots, consumed by the classification) shares it as well.
struct CcWs {
    int32_t *T, *psize, *adj, *aux, *queue, *list, *ctr, *stamp, *dirty0, *dirty1;
    uint32_t *bits;       // kept-piece starts, one bit per pixel
    int32_t *chunksum;    // per kBitChunk words
    uint8_t *visit, *flag;
    int64_t nwords, nchunks;
    int64_t bytes;
};
// ctr words
enum { CTR_NSMALL = 0, CTR_NOVER = 1, CTR_CURSOR = 2, CTR_CHANGED = 3, CTR_NKEPT = 4, CTR_ERR = 5, CTR_NDIRTY0 = 6,
       CTR_NDIRTY1 = 7, CTR_ROUNDS = 8, CTR_NROOTS = 9, CTR_KBEFORE = 10, CTR_KCORE = 11, CTR_FAIL = 12, CTR_WORDS = 16 };
// flag bits (strip mode: which results depend on pixels outside the strip)
enum { FLAG_CUT = 1, FLAG_ADJ_UNKNOWN = 2, FLAG_TFIX_UNKNOWN = 4, FLAG_LABEL_UNKNOWN = 8 };

static CcWs cc_ws_layout(void *base, int64_t N)
{
    CcWs w;
    char *p = (char *)base;
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        char *r = p + off;
        off += round_up(bytes, 256);
        return r;
    };
    w.T = (int32_t *)take(N * 4);
    w.psize = (int32_t *)take(N * 4);
    w.adj = (int32_t *)take(N * 4);
    w.aux = (int32_t *)take(N * 4);
    w.queue = (int32_t *)take(2 * N * 4);
    w.list = (int32_t *)take(N * 4);
    w.nwords = ceil_div(N, 32);
    w.nchunks = ceil_div(w.nwords, kBitChunk);
    w.bits = (uint32_t *)take(w.nchunks * kBitChunk * 4);
    w.chunksum = (int32_t *)take((w.nchunks + 1) * 4);
    w.ctr = (int32_t *)take(CTR_WORDS * 4);
    w.visit = (uint8_t *)take(N);
    w.flag = (uint8_t *)take(N);
    w.stamp = (int32_t *)take(N * 4);    // round in which a piece was last queued for re-evaluation
    w.dirty0 = (int32_t *)take(N * 4);   // ping-pong lists of pieces to re-evaluate
    w.dirty1 = (int32_t *)take(N * 4);
    w.bytes = off;
    return w;
}

__device__ __forceinline__ int32_t uf_find(const int32_t *parent, int32_t x)
{
    int32_t p = __ldcg(parent + x);
    while (p != x) {
        x = p;
        p = __ldcg(parent + x);
    }
    return x;
}

__device__ __forceinline__ void uf_union(int32_t *parent, int32_t a, int32_t b)
{
    bool done;
    do {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a < b) {
            const int32_t old = atomicMin(parent + b, a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            const int32_t old = atomicMin(parent + a, b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// phase 1a: union-find inside a 32 x 32 tile, entirely in shared memory (short chains, no global
// atomics); every pixel then points at the global index of its tile-local root.  Local and
// global raster order agree inside a tile, so the local min-index root is the tile's first pixel
// of the component.
constexpr int kTile = 32;

__device__ __forceinline__ int uf_find_s(const volatile int *parent, int x)
{
    int p = parent[x];
    while (p != x) {
        x = p;
        p = parent[x];
    }
    return x;
}

__device__ __forceinline__ void uf_union_s(int *parent, int a, int b)
{
    bool done;
    do {
        a = uf_find_s(parent, a);
        b = uf_find_s(parent, b);
        if (a < b) {
            const int old = atomicMin(parent + b, a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            const int old = atomicMin(parent + a, b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

__global__ void __launch_bounds__(256)
cc_local_kernel(const int32_t *__restrict__ lab, int32_t *__restrict__ parent, int32_t *__restrict__ psize,
                int H, int W, int32_t mask_label)
{
    __shared__ int32_t s_lab[kTile * kTile];
    __shared__ int s_par[kTile * kTile];
    __shared__ int s_cnt[kTile * kTile];   // pixels per tile-local root
    __shared__ unsigned char s_len[kTile * kTile];   // run length at run heads
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // a warp = one tile row: every pixel is linked to the first pixel of its horizontal run
    // with a ballot (no atomics)
    for (int ly = warp; ly < kTile; ly += 8) {
        const int y = y0 + ly, x = x0 + lane;
        const int32_t l = (y < H && x < W) ? lab[(int64_t)y * W + x] : mask_label;
        const int32_t lp = __shfl_up_sync(0xffffffffu, l, 1);
        const bool head = (lane == 0) || (l != lp) || (l == mask_label);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
        const unsigned later = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
        s_lab[ly * kTile + lane] = l;
        s_par[ly * kTile + lane] = ly * kTile + start;
        s_cnt[ly * kTile + lane] = 0;
        s_len[ly * kTile + lane] = (unsigned char)((later ? (__ffs(later) - 1) : 32) - lane);
    }
    __syncthreads();
    // vertical joins, once per pair of overlapping runs (at the first column of the overlap)
    for (int i = threadIdx.x + kTile; i < kTile * kTile; i += 256) {
        const int32_t l = s_lab[i];
        if (l == mask_label || s_lab[i - kTile] != l) continue;
        const bool cur_head = (i % kTile) == 0 || s_lab[i - 1] != l;   // from labels: s_par is being rewritten
        const int up = i - kTile;
        const bool up_head = (up % kTile) == 0 || s_lab[up - 1] != l;
        if (cur_head || up_head) uf_union_s(s_par, i, up);
    }
    __syncthreads();
    // Only run heads can be roots (every other pixel points at its run head and is never
    // re-parented), so the chains are walked once per run, not once per pixel: heads are
    // compressed to their root first, then every pixel is two loads away from it.
    int head_root[kTile * kTile / 256];
#pragma unroll
    for (int k = 0; k < kTile * kTile / 256; ++k) {
        const int i = threadIdx.x + k * 256;
        const int32_t l = s_lab[i];
        const bool is_head = (i % kTile) == 0 || s_lab[i - 1] != l || l == mask_label;
        head_root[k] = is_head ? uf_find_s(s_par, i) : -1;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTile * kTile / 256; ++k) {
        if (head_root[k] >= 0) {
            s_par[threadIdx.x + k * 256] = head_root[k];
            atomicAdd(&s_cnt[head_root[k]], (int)s_len[threadIdx.x + k * 256]);   // one add per run
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTile * kTile / 256; ++k) {
        const int i = threadIdx.x + k * 256;
        const int ly = i / kTile, lx = i % kTile;
        const int y = y0 + ly, x = x0 + lx;
        if (y >= H || x >= W) continue;
        const int64_t g = (int64_t)y * W + x;
        if (s_lab[i] == mask_label) {
            parent[g] = -1;
            psize[g] = 0;
            continue;
        }
        const int r = s_par[s_par[i]];
        parent[g] = (int32_t)((int64_t)(y0 + r / kTile) * W + x0 + r % kTile);
        psize[g] = (r == i) ? s_cnt[i] : 0;    // tile-local component size at its tile-local root
    }
}

// phase 1b: join components across tile borders (global union-find with atomicMin)
__global__ void __launch_bounds__(256)
cc_border_kernel(const int32_t *__restrict__ lab, int32_t *parent, int H, int W, int32_t mask_label)
{
    // one thread per border pixel: left columns of tiles (x % 32 == 0, x > 0), then top rows
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ncolb = (W - 1) / kTile;          // number of vertical borders
    const int nrowb = (H - 1) / kTile;          // number of horizontal borders
    const int64_t nv = (int64_t)ncolb * H;
    const int64_t nh = (int64_t)nrowb * W;
    int y, x;
    bool vertical;
    if (t < nv) {
        vertical = true;
        y = (int)(t / ncolb);
        x = (int)(t % ncolb + 1) * kTile;
    } else if (t < nv + nh) {
        vertical = false;
        const int64_t u = t - nv;
        y = (int)(u / W + 1) * kTile;
        x = (int)(u % W);
    } else {
        return;
    }
    const int64_t i = (int64_t)y * W + x;
    const int32_t l = lab[i];
    if (l == mask_label) return;
    if (vertical) {
        if (lab[i - 1] == l) uf_union(parent, (int32_t)i, (int32_t)(i - 1));
    } else {
        if (lab[i - W] == l) uf_union(parent, (int32_t)i, (int32_t)(i - W));
    }
}

// phase 1c: every pixel points at its component root (T = root, written over the parent array: a
// concurrent find that reads the new value just skips part of its chain); tile-local sizes are summed
// into the root (one atomic per tile-local component); roots are appended to a compact list.
__global__ void __launch_bounds__(256)
cc_flatten_kernel(int32_t *T, int32_t *psize, int32_t *roots, int32_t *ctr, int64_t N)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool is_root = false;
    if (i < N) {
        const int32_t p = __ldcg(T + i);
        if (p >= 0) {
            const int32_t root = uf_find(T, p);
            if (root != p) T[i] = root;
            const int32_t mine = psize[i];
            if (mine > 0 && root != (int32_t)i) atomicAdd(psize + root, mine);
            is_root = root == (int32_t)i;
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, is_root);
    if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(ctr + CTR_NROOTS, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (is_root) roots[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)i;
    }
}

// phase 2/3 lists: classify the components by size (over the compact root list): larger than
// max_size -> split list (grows from the end of `list`), smaller than min_size -> small list,
// otherwise kept: its start pixel is marked in the bitmap.
__global__ void __launch_bounds__(256)
cc_classify_kernel(const int32_t *__restrict__ roots, const int32_t *__restrict__ psize, int32_t *list,
                   int32_t *adj, int32_t *aux, uint32_t *bits, int32_t *ctr, int64_t N, int64_t min_size,
                   int64_t max_size)
{
    const int n_roots = ctr[CTR_NROOTS];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_roots; e += gridDim.x * blockDim.x) {
        const int32_t i = roots[e];
        const int64_t sz = psize[i];
        if (sz > max_size) {
            list[N - 1 - atomicAdd(ctr + CTR_NOVER, 1)] = i;
        } else if (sz < min_size) {
            list[atomicAdd(ctr + CTR_NSMALL, 1)] = i;
            adj[i] = -1;
            aux[i] = i;  // tfix: optimistic "labelled at its own time"
        } else {
            atomicOr(bits + (i >> 5), 1u << (i & 31));
        }
    }
}

__device__ __forceinline__ bool nbr(int dir, int py, int px, int H, int W, int32_t &q)
{
    // reference neighbour order: x+1, x-1, y+1, y-1
    int yy = py, xx = px;
    if (dir == 0) xx += 1;
    else if (dir == 1) xx -= 1;
    else if (dir == 2) yy += 1;
    else yy -= 1;
    if (xx < 0 || xx >= W || yy < 0 || yy >= H) return false;
    q = yy * W + xx;
    return true;
}

// phase 2: replay the BFS cap on oversized components, one thread each
__global__ void __launch_bounds__(128)
cc_split_kernel(const int32_t *__restrict__ lab, const int32_t *__restrict__ parent, int32_t *T,
                int32_t *psize, int32_t *queue, const int32_t *__restrict__ list, int32_t *ctr,
                uint8_t *visit, int64_t N, int H, int W, int64_t max_size)
{
    const int n_over = ctr[CTR_NOVER];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_over; e += gridDim.x * blockDim.x) {
        const int32_t r = list[N - 1 - e];
        const int32_t L = lab[r];
        const int32_t n = psize[r];
        int32_t *qu = queue + atomicAdd(ctr + CTR_CURSOR, n);
        // uncapped BFS: mark members (visit = 2), bounding box
        int cnt = 1, head = 0;
        qu[0] = r;
        visit[r] = 2;
        int ymin = r / W, ymax = ymin, xmin = r % W, xmax = xmin;
        while (head < cnt) {
            const int32_t p = qu[head++];
            const int py = p / W, px = p % W;
            for (int d = 0; d < 4; ++d) {
                int32_t q;
                if (!nbr(d, py, px, H, W, q)) continue;
                if (lab[q] == L && visit[q] == 0) {
                    visit[q] = 2;
                    qu[cnt++] = q;
                    const int qy = q / W, qx = q % W;
                    ymin = min(ymin, qy); ymax = max(ymax, qy);
                    xmin = min(xmin, qx); xmax = max(xmax, qx);
                }
            }
        }
        // pieces in raster order of their first unassigned member
        for (int y = ymin; y <= ymax; ++y) {
            for (int x = xmin; x <= xmax; ++x) {
                const int32_t s = y * W + x;
                if (parent[s] != r || visit[s] != 2) continue;
                int pc = 1, ph = 0;
                qu[0] = s;
                visit[s] = 1;
                while (ph < pc && (int64_t)pc < max_size) {
                    const int32_t p = qu[ph];
                    const int py = p / W, px = p % W;
                    for (int d = 0; d < 4; ++d) {
                        int32_t q;
                        if (!nbr(d, py, px, H, W, q)) continue;
                        if (lab[q] == L && visit[q] == 2) {
                            visit[q] = 1;
                            qu[pc++] = q;
                            if ((int64_t)pc >= max_size) break;
                        }
                    }
                    ++ph;
                }
                for (int i = 0; i < pc; ++i) {
                    T[qu[i]] = s;
                    visit[qu[i]] = 0;
                }
                psize[s] = pc;
            }
        }
    }
}

// phase 3 list: starts of pieces smaller than min_size (one global atomic per CTA)
__global__ void __launch_bounds__(256)
cc_list_small_kernel(const int32_t *__restrict__ T, const int32_t *__restrict__ psize, int32_t *list,
                     int32_t *adj, int32_t *aux, int32_t *ctr, int64_t N, int64_t min_size)
{
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool is_small = (i < N) && T[i] == (int32_t)i && (int64_t)psize[i] < min_size;
    const unsigned m = __ballot_sync(0xffffffffu, is_small);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) {
            const int c = s_warp[w];
            s_warp[w] = tot;
            tot += c;
        }
        s_base = tot ? atomicAdd(ctr + CTR_NSMALL, tot) : 0;
    }
    __syncthreads();
    if (is_small) {
        const int e = s_base + s_warp[warp] + __popc(m & ((1u << lane) - 1u));
        list[e] = (int32_t)i;
        adj[i] = -1;
        aux[i] = (int32_t)i;  // tfix: optimistic "labelled at its own time"
    }
}

struct CcParams {
    int H, W;
    int64_t min_size, max_size;
    int32_t mask_label, start_label;
    int32_t optimistic;   // round 1: every piece counts as labelled from its own start pixel
};

// Scan position at which pixel q (not in piece t) receives a label > mask label, kTInf if never:
// kept pieces at their start; merged pieces at their start for start_label 0 (they always carry
// a label >= 0), at `tfix` for start_label 1 (label 0 == mask label until a later re-scan).
__device__ __forceinline__ int32_t label_time(const int32_t *lab, const int32_t *T, const int32_t *psize,
                                              const int32_t *aux, const CcParams &P, int32_t q, int32_t t)
{
    if (P.optimistic) {
        // one load per neighbour: T is -1 on masked pixels, and in round 1 both kept and merged
        // pieces are taken as labelled from their start (merged ones are re-checked in round 2)
        const int32_t tq = T[q];
        return (tq < 0 || tq == t) ? kTInf : tq;
    }
    if (lab[q] == P.mask_label) return kTInf;
    const int32_t tq = T[q];
    if (tq == t) return kTInf;
    if ((int64_t)psize[tq] >= P.min_size) return tq;
    if (P.start_label == 0) return tq;
    return __ldcg(aux + tq);
}

__device__ __forceinline__ bool labelled_at(const int32_t *lab, const int32_t *T, const int32_t *psize,
                                            const int32_t *aux, const CcParams &P, int32_t q, int32_t t,
                                            int32_t now)
{
    return label_time(lab, T, psize, aux, P, q, t) < now;
}

// replay of the reference BFS restricted to piece t, started at pixel s
__device__ int bfs_piece(const int32_t *lab, const int32_t *T, const int32_t *psize, const int32_t *aux,
                         uint8_t *visit, int32_t *qu, const CcParams &P, int32_t t, int32_t s,
                         int32_t L, int32_t &adj_out, int32_t &first_time)
{
    int cnt = 1, head = 0;
    int32_t adjq = -1;
    int32_t tmin = kTInf;   // earliest scan position at which any examined neighbour is labelled
    qu[0] = s;
    visit[s] = 1;
    while (head < cnt && (int64_t)cnt < P.max_size) {
        const int32_t p = qu[head];
        const int py = p / P.W, px = p % P.W;
        for (int d = 0; d < 4; ++d) {
            int32_t q;
            if (!nbr(d, py, px, P.H, P.W, q)) continue;
            const bool same = T[q] == t;   // a piece is a set of equal-label pixels: T alone identifies it
            if (same) {
                if (!visit[q]) {
                    visit[q] = 1;
                    qu[cnt++] = q;
                    if ((int64_t)cnt >= P.max_size) break;
                }
            } else {
                const int32_t lt = label_time(lab, T, psize, aux, P, q, t);
                tmin = min(tmin, lt);
                if (lt < s) adjq = q;
            }
        }
        ++head;
    }
    adj_out = adjq;
    first_time = tmin;
    return cnt;
}

struct CcArrays {
    const int32_t *lab, *T, *psize, *list;
    int32_t *adj, *aux, *queue, *ctr, *stamp;
    uint8_t *visit;
};

// A small piece that touches pixel p of a piece whose `tfix` just changed may have relied on the
// old value -- later pieces at their start, EARLIER pieces at one of their re-scan times -- so
// every small neighbour is queued for re-evaluation (once per round).
__device__ __forceinline__ void push_dependents(const CcArrays &A, const CcParams &P, int32_t p, int32_t t,
                                                int round_id, int32_t *dirty_next, int32_t *n_next)
{
    const int py = p / P.W, px = p % P.W;
    for (int d = 0; d < 4; ++d) {
        int32_t q;
        if (!nbr(d, py, px, P.H, P.W, q)) continue;
        if (A.lab[q] == P.mask_label) continue;
        const int32_t tq = A.T[q];
        if (tq == t || (int64_t)A.psize[tq] >= P.min_size) continue;
        if (atomicExch(A.stamp + tq, round_id) != round_id) dirty_next[atomicAdd(n_next, 1)] = tq;
    }
}

// `adjacent` of one small piece (start pixel t) under the current knowledge of its earlier
// neighbours; `qu` = queue space for max(psize[t], 1) pixels (+ psize[t] more for the re-scan copy,
// claimed from the global cursor on demand).  Returns true when the piece's tfix changed.
__device__ bool small_piece_adjacent(const CcArrays &A, const CcParams &P, int32_t t, int32_t *qu,
                                     int round_id, int32_t *dirty_next, int32_t *n_next)
{
    const int32_t L = A.lab[t];
    const int32_t n = A.psize[t];
    int32_t a = -1;
    int32_t fix = t;
    int cnt = 1;
    const int32_t *members = qu;   // pixels of the first BFS
    if (n == 1) {
        const int py = t / P.W, px = t % P.W;
        // (with max_size <= 1 the reference's BFS loop never runs: no neighbour is looked at)
        for (int d = 0; d < 4 && P.max_size > 1; ++d) {
            int32_t q;
            if (!nbr(d, py, px, P.H, P.W, q)) continue;
            if (labelled_at(A.lab, A.T, A.psize, A.aux, P, q, t, t)) a = q;
        }
        if (a < 0 && P.start_label == 1) fix = kTInf;  // stays label 0: no later pixel to re-enter at
        qu[0] = t;
    } else {
        int32_t tmin;
        cnt = bfs_piece(A.lab, A.T, A.psize, A.aux, A.visit, qu, P, t, t, L, a, tmin);
        for (int i = 0; i < cnt; ++i) A.visit[qu[i]] = 0;
        if (a < 0 && P.start_label == 1) {
            // Merged to 0 == mask label: the raster scan re-enters the piece at each of its later
            // pixels (ascending) until the BFS from there sees a labelled neighbour.
            fix = kTInf;
            // (queue demand stays <= 2N: psize per piece + one copy per re-scanned piece)
            int32_t *cand = A.queue + atomicAdd(A.ctr + CTR_CURSOR, cnt);
            int32_t *qu2 = qu;
            for (int i = 0; i < cnt; ++i) cand[i] = qu[i];
            members = cand;
            if ((int64_t)n < P.max_size) {
                // The BFS is not cut by the size cap, so from ANY start it examines every
                // neighbour of the piece: the first successful re-scan is at the first member
                // pixel after the earliest labelling time among those neighbours.
                int32_t s = kTInf;
                if (tmin != kTInf)
                    for (int i = 0; i < cnt; ++i)
                        if (cand[i] > tmin && cand[i] < s) s = cand[i];
                if (s != kTInf) {
                    int32_t a2, tm2;
                    const int c2 = bfs_piece(A.lab, A.T, A.psize, A.aux, A.visit, qu2, P, t, s, L, a2, tm2);
                    for (int i = 0; i < c2; ++i) A.visit[qu2[i]] = 0;
                    a = a2;
                    fix = (a2 >= 0) ? s : kTInf;
                }
            } else {
                // size cap can truncate the BFS (min_size > max_size): replay every re-scan
                int32_t last = t;
                while (true) {
                    int32_t s = kTInf;
                    for (int i = 0; i < cnt; ++i)
                        if (cand[i] > last && cand[i] < s) s = cand[i];
                    if (s == kTInf) break;
                    last = s;
                    int32_t a2, tm2;
                    const int c2 = bfs_piece(A.lab, A.T, A.psize, A.aux, A.visit, qu2, P, t, s, L, a2, tm2);
                    for (int i = 0; i < c2; ++i) A.visit[qu2[i]] = 0;
                    if (a2 >= 0) {
                        a = a2;
                        fix = s;
                        break;
                    }
                }
            }
        }
    }
    A.adj[t] = a;
    if (A.aux[t] == fix) return false;
    A.aux[t] = fix;
    if (P.start_label == 0) return true;   // no label-0 ambiguity: nobody depends on tfix
    // tell later small neighbours of the piece to look again
    for (int i = 0; i < cnt; ++i) push_dependents(A, P, members[i], t, round_id, dirty_next, n_next);
    return true;
}

// round 1: every small piece, in parallel (optimistic: every merged neighbour counts as labelled
// from its own start time).  Pieces whose tfix turns out different queue their dependents.
__global__ void __launch_bounds__(128)
cc_small_adjacent_kernel(CcArrays A, CcParams P, int32_t *dirty_next)
{
    const int n_small = A.ctr[CTR_NSMALL];
    const int lane = threadIdx.x & 31;
    const int stride = gridDim.x * blockDim.x;
    // warp-uniform trip count: the queue space of a warp's pieces is claimed with one atomic
    for (int e0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); e0 < n_small; e0 += stride) {
        const int e = e0 + lane;
        const bool active = e < n_small;
        int32_t t = 0, n = 0;
        if (active) {
            t = A.list[e];
            n = A.psize[t];
        }
        const int need = active ? n : 0;
        int incl = need;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (lane == 0 && total > 0) base = atomicAdd(A.ctr + CTR_CURSOR, total);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (!active) continue;
        small_piece_adjacent(A, P, t, A.queue + base + incl - need, 1, dirty_next, A.ctr + CTR_NDIRTY0);
    }
}

// rounds 2..: re-evaluate only the queued pieces (the start_label = 1 "label 0 == mask label"
// chains; usually there are none).  One launch per round; the lists ping-pong.  The fixed point
// is unique: a decision at a scan position depends only on earlier scan positions.
__global__ void __launch_bounds__(128)
cc_small_round_kernel(CcArrays A, CcParams P, const int32_t *__restrict__ cur, int32_t *nxt, int cur_ctr,
                      int nxt_ctr, int round_id)
{
    const int n = A.ctr[cur_ctr];
    const int lane = threadIdx.x & 31;
    const int stride = gridDim.x * blockDim.x;
    for (int e0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); e0 < n; e0 += stride) {
        const int e = e0 + lane;
        const bool active = e < n;
        int32_t t = 0, need = 0;
        if (active) {
            t = cur[e];
            need = A.psize[t];
        }
        int incl = need;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (lane == 0 && total > 0) base = atomicAdd(A.ctr + CTR_CURSOR, total);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (!active) continue;
        small_piece_adjacent(A, P, t, A.queue + base + incl - need, round_id, nxt, A.ctr + nxt_ctr);
    }
}

// phase 4a/b/c: rank of kept piece starts (chunked prefix sum)
__device__ __forceinline__ bool is_kept_start(const int32_t *T, const int32_t *psize, int64_t i,
                                              int64_t min_size)
{
    return T[i] == (int32_t)i && (int64_t)psize[i] >= min_size;
}

__global__ void __launch_bounds__(256)
cc_count_kernel(const int32_t *__restrict__ T, const int32_t *__restrict__ psize, int32_t *blocksum,
                int64_t N, int64_t min_size)
{
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kScanChunk;
    int c = 0;
    for (int j = threadIdx.x; j < kScanChunk; j += 256) {
        const int64_t i = base + j;
        if (i < N && is_kept_start(T, psize, i, min_size)) ++c;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) blocksum[blockIdx.x] = s_cnt;
}

__global__ void __launch_bounds__(1024)
cc_scan_blocks_kernel(int32_t *blocksum, int64_t nblocks, int32_t *ctr)
{
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t b0 = 0; b0 < nblocks; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const int v = (i < nblocks) ? blocksum[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int carry = s_carry;
        const int excl = carry + (warp ? s_warp[warp - 1] : 0) + incl - v;
        if (i < nblocks) blocksum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) ctr[CTR_NKEPT] = s_carry;
}

__global__ void __launch_bounds__(256)
cc_number_kernel(const int32_t *__restrict__ T, const int32_t *__restrict__ psize,
                 const int32_t *__restrict__ blocksum, int32_t *aux, int64_t N, int64_t min_size,
                 int32_t start_label)
{
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = blocksum[blockIdx.x];
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kScanChunk;
    for (int j0 = 0; j0 < kScanChunk; j0 += 256) {
        const int64_t i = base + j0 + threadIdx.x;
        const bool k = (i < N) && is_kept_start(T, psize, i, min_size);
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (k) aux[i] = start_label + before + __popc(m & ((1u << lane) - 1u));
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
}

// phase 4d: final labels
__global__ void __launch_bounds__(256)
cc_resolve_kernel(const int32_t *__restrict__ lab, const int32_t *__restrict__ T,
                  const int32_t *__restrict__ psize, const int32_t *__restrict__ adj,
                  const int32_t *__restrict__ aux, int32_t *__restrict__ out, int32_t *ctr, int64_t N,
                  int64_t min_size, int32_t mask_label)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (lab[i] == mask_label) {
        out[i] = mask_label;
        return;
    }
    int32_t t = T[i];
    int32_t r = mask_label;
    bool done = false;
    for (int hop = 0; hop < (1 << 24) && !done; ++hop) {   // chains are short; never spin
        if ((int64_t)psize[t] >= min_size) {
            r = aux[t];
            done = true;
        } else {
            const int32_t a = adj[t];
            if (a < 0) {
                r = 0;  // `adjacent` initial value
                done = true;
            } else {
                t = T[a];
            }
        }
    }
    if (!done) atomicExch(ctr + CTR_ERR, 2);
    out[i] = r;
}

}  // namespace obia

using namespace obia;

extern "C" int64_t obia_b200_connectivity_workspace_bytes(int64_t H, int64_t W)
{
    if (H <= 0 || W <= 0) return -1;
    return cc_ws_layout(nullptr, H * W).bytes;
}

extern "C" int obia_b200_enforce_connectivity(const int32_t *labels_in, int32_t *labels_out,
                                              void *workspace, int64_t H, int64_t W, int64_t min_size,
                                              int64_t max_size, int32_t start_label,
                                              int64_t *n_labels_host, void *stream)
{
    if (!labels_in || !labels_out || !workspace || H <= 0 || W <= 0)
        return set_err(OBIA_B200_ERR_ARG, "enforce_connectivity: bad argument");
    if (max_size < 1) return set_err(OBIA_B200_ERR_ARG, "enforce_connectivity: max_size must be >= 1");
    if (start_label != 0 && start_label != 1) return set_err(OBIA_B200_ERR_ARG, "start_label should be 0 or 1.");
    const int64_t N = H * W;
    if (N >= 0x7fffffffLL) return set_err(OBIA_B200_ERR_UNSUPPORTED, "enforce_connectivity: H*W exceeds int32");
    cudaStream_t st = (cudaStream_t)stream;
    CcWs w = cc_ws_layout(workspace, N);
    const int32_t mask_label = start_label - 1;
    const unsigned gridN = (unsigned)ceil_div(N, 256);

    OBIA_CUDA_CHECK(cudaMemsetAsync(w.ctr, 0, CTR_WORDS * 4, st));
    {
        dim3 tiles((unsigned)ceil_div(W, kTile), (unsigned)ceil_div(H, kTile));
        cc_local_kernel<<<tiles, 256, 0, st>>>(labels_in, w.parent, w.psize, w.visit, w.stamp, (int)H, (int)W,
                                               mask_label);
        OBIA_LAUNCH_CHECK();
        const int64_t nborder = ((W - 1) / kTile) * H + ((H - 1) / kTile) * W;
        if (nborder > 0) {
            cc_border_kernel<<<(unsigned)ceil_div(nborder, 256), 256, 0, st>>>(labels_in, w.parent, (int)H, (int)W,
                                                                              mask_label);
            OBIA_LAUNCH_CHECK();
        }
    }
    cc_flatten_kernel<<<gridN, 256, 0, st>>>(labels_in, w.parent, w.T, w.psize, N, mask_label);
    OBIA_LAUNCH_CHECK();
    cc_list_over_kernel<<<gridN, 256, 0, st>>>(w.T, w.psize, w.list, w.ctr, N, max_size);
    OBIA_LAUNCH_CHECK();
    cc_split_kernel<<<kNumSMs * 2, 128, 0, st>>>(labels_in, w.parent, w.T, w.psize, w.queue, w.list, w.ctr,
                                                 w.visit, N, (int)H, (int)W, max_size);
    OBIA_LAUNCH_CHECK();
    cc_list_small_kernel<<<gridN, 256, 0, st>>>(w.T, w.psize, w.list, w.adj, w.aux, w.ctr, N, min_size);
    OBIA_LAUNCH_CHECK();

    CcParams P;
    P.H = (int)H; P.W = (int)W; P.min_size = min_size; P.max_size = max_size;
    P.mask_label = mask_label; P.start_label = start_label; P.optimistic = 0;
    int32_t hctr[CTR_WORDS];
    {
        CcArrays A;
        A.lab = labels_in; A.T = w.T; A.psize = w.psize; A.list = w.list;
        A.adj = w.adj; A.aux = w.aux; A.queue = w.queue; A.ctr = w.ctr; A.stamp = w.stamp; A.visit = w.visit;
        OBIA_CUDA_CHECK(cudaMemsetAsync(w.ctr + CTR_CURSOR, 0, 4, st));
        P.optimistic = 1;
        cc_small_adjacent_kernel<<<kNumSMs * 8, 128, 0, st>>>(A, P, w.dirty0);
        OBIA_LAUNCH_CHECK();
        P.optimistic = 0;
        if (start_label == 1) {   // start_label 0 has no label-0 ambiguity: round 1 is exact
            int round_id = 1;
            int cur = 0;   // dirty list written by the previous round
            while (true) {
                OBIA_CUDA_CHECK(cudaMemcpyAsync(hctr, w.ctr, sizeof(hctr), cudaMemcpyDeviceToHost, st));
                OBIA_CUDA_CHECK(cudaStreamSynchronize(st));
                const int n_dirty = hctr[cur ? CTR_NDIRTY1 : CTR_NDIRTY0];
                if (getenv("OBIA_B200_DEBUG"))
                    fprintf(stderr, "[obia_b200] connectivity: round %d, %d small pieces, %d queued\n", round_id,
                            hctr[CTR_NSMALL], n_dirty);
                if (n_dirty == 0) break;
                if (round_id > (1 << 28)) return set_err(OBIA_B200_ERR_CUDA, "enforce_connectivity: no fixed point");
                // a batch of rounds without host round trips (an empty round costs a few microseconds)
                const int batch = 16;
                for (int b = 0; b < batch; ++b) {
                    ++round_id;
                    const int cur_ctr = cur ? CTR_NDIRTY1 : CTR_NDIRTY0, nxt_ctr = cur ? CTR_NDIRTY0 : CTR_NDIRTY1;
                    OBIA_CUDA_CHECK(cudaMemsetAsync(w.ctr + nxt_ctr, 0, 4, st));
                    OBIA_CUDA_CHECK(cudaMemsetAsync(w.ctr + CTR_CURSOR, 0, 4, st));   // queue space is recycled
                    cc_small_round_kernel<<<kNumSMs * 4, 128, 0, st>>>(A, P, cur ? w.dirty1 : w.dirty0,
                                                                       cur ? w.dirty0 : w.dirty1, cur_ctr, nxt_ctr,
                                                                       round_id);
                    OBIA_LAUNCH_CHECK();
                    cur ^= 1;
                }
            }
        }
    }

    cc_count_kernel<<<(unsigned)w.nblocks, 256, 0, st>>>(w.T, w.psize, w.blocksum, N, min_size);
    OBIA_LAUNCH_CHECK();
    cc_scan_blocks_kernel<<<1, 1024, 0, st>>>(w.blocksum, w.nblocks, w.ctr);
    OBIA_LAUNCH_CHECK();
    cc_number_kernel<<<(unsigned)w.nblocks, 256, 0, st>>>(w.T, w.psize, w.blocksum, w.aux, N, min_size,
                                                          start_label);
    OBIA_LAUNCH_CHECK();
    cc_resolve_kernel<<<gridN, 256, 0, st>>>(labels_in, w.T, w.psize, w.adj, w.aux, labels_out, w.ctr, N,
                                             min_size, mask_label);
    OBIA_LAUNCH_CHECK();
    OBIA_CUDA_CHECK(cudaMemcpyAsync(hctr, w.ctr, sizeof(hctr), cudaMemcpyDeviceToHost, st));
    OBIA_CUDA_CHECK(cudaStreamSynchronize(st));
    if (n_labels_host) *n_labels_host = hctr[CTR_NKEPT];
    if (hctr[CTR_ERR])
        return set_err(OBIA_B200_ERR_CUDA, "enforce_connectivity: internal consistency check failed (%d)",
                       hctr[CTR_ERR]);
    return OBIA_B200_OK;
}
