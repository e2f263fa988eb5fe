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

#include <algorithm>

#include "common.cuh"

#ifndef OBIA_CC_ADJ_CTAS
#define OBIA_CC_ADJ_CTAS 8      // CTAs per SM (128 threads) of the small-piece adjacency kernel
#endif
#ifndef OBIA_CC_FIN_CTAS
#define OBIA_CC_FIN_CTAS 4      // CTAs per SM (256 threads) of the merge-chain kernel
#endif

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
       CTR_NDIRTY1 = 7, CTR_ROUNDS = 8, CTR_NROOTS = 9, CTR_KBEFORE = 10, CTR_KCORE = 11, CTR_FAIL = 12, CTR_HASZERO = 13,
       CTR_CUTMIN = 14, CTR_KVALID = 15, CTR_WORDS = 16 };
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
    // 1024 pixels per CTA, one append to the root list per CTA (a counter shared by every warp of the
    // grid serialises in L2: 3 M single-address atomics cost 1.5 ms on a 10^8-pixel raster)
    constexpr int PER = 4;
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * (256 * PER);
    bool is_root[PER];
    int nroot = 0;
    {
        // the four pixels of a thread walk their parent chains in lock step: the loads of one step are
        // requested together (the kernel waits on memory: one dependent round trip per chain link)
        int32_t par[PER], x[PER], mine[PER];
        bool act[PER], fin[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int64_t i = base + k * 256 + threadIdx.x;
            par[k] = (i < N) ? __ldcg(T + i) : -1;
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int64_t i = base + k * 256 + threadIdx.x;
            act[k] = par[k] >= 0;
            mine[k] = act[k] ? psize[i] : 0;
            x[k] = par[k];
            fin[k] = !act[k];
        }
        bool go = true;
        while (go) {
            go = false;
            int32_t nx[PER];
#pragma unroll
            for (int k = 0; k < PER; ++k) nx[k] = fin[k] ? x[k] : __ldcg(T + x[k]);
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                if (fin[k]) continue;
                if (nx[k] == x[k]) {
                    fin[k] = true;
                } else {
                    x[k] = nx[k];
                    go = true;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int64_t i = base + k * 256 + threadIdx.x;
            is_root[k] = false;
            if (act[k]) {
                const int32_t root = x[k];
                if (root != par[k]) T[i] = root;
                if (mine[k] > 0 && root != (int32_t)i) atomicAdd(psize + root, mine[k]);
                is_root[k] = root == (int32_t)i;
            }
            nroot += is_root[k];
        }
    }
    int incl = nroot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) {
            const int c = s_warp[w];
            s_warp[w] = tot;
            tot += c;
        }
        s_base = tot ? atomicAdd(ctr + CTR_NROOTS, tot) : 0;
    }
    __syncthreads();
    int pos = s_base + s_warp[warp] + incl - nroot;
#pragma unroll
    for (int k = 0; k < PER; ++k)
        if (is_root[k]) roots[pos++] = (int32_t)(base + k * 256 + threadIdx.x);
}

// phase 2/3 lists: classify the components by size (over the compact root list): larger than
// max_size -> split list (grows from the end of `list`), smaller than min_size -> small list,
// otherwise kept: its start pixel is marked in the bitmap.
__global__ void __launch_bounds__(256)
cc_classify_kernel(const int32_t *__restrict__ roots, const int32_t *__restrict__ psize, int32_t *list,
                   int32_t *adj, int32_t *aux, int32_t *stamp, uint32_t *bits, const uint8_t *__restrict__ flag,
                   int32_t *ctr, int64_t N, int64_t min_size, int64_t max_size, const int2 *__restrict__ wsizes,
                   int win_px)
{
    const int n_roots = ctr[CTR_NROOTS];
    const int lane = threadIdx.x & 31;
    // Four roots per lane and trip: the kernel is a chain of dependent loads (root -> size / flag -> scattered
    // stores), so the loads of the four are requested together.  Warp-uniform trip count: one append per warp
    // and group of 32 roots to the small list.
    constexpr int U = 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * U;
    for (int64_t e0 = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * U; e0 < n_roots; e0 += stride) {
        int32_t iu[U];
        int64_t szu[U], mnu[U], mxu[U];
        uint8_t fu[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = e0 + u * 32 + lane;
            iu[u] = (e < n_roots) ? roots[e] : -1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            szu[u] = 0;
            fu[u] = 0;
            mnu[u] = min_size;
            mxu[u] = max_size;
            if (iu[u] >= 0) {
                szu[u] = psize[iu[u]];
                if (flag) fu[u] = flag[iu[u]];
                if (wsizes) {   // slab of windows (tiled driver): the sizes of the window the component lies in
                    const int2 ws = wsizes[iu[u] / win_px];
                    mnu[u] = ws.x;
                    mxu[u] = ws.y;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int32_t i = iu[u];
            int64_t sz = szu[u];
            const int64_t mn = mnu[u], mx = mxu[u];
            // Strip mode: a component cut by an open strip edge is unknown anyway (FLAG_CUT: every result
            // that depends on it is reported incomplete).  It is booked as ONE kept piece -- the fragments
            // along an edge row would otherwise form long chains of "small pieces without an earlier
            // neighbour" whose fixed point needs hundreds of rounds.
            if (i >= 0 && (fu[u] & FLAG_CUT)) sz = mn > mx ? mx : mn;
            const bool small = i >= 0 && sz <= mx && sz < mn;   // (oversized components are split first)
            const unsigned m = __ballot_sync(0xffffffffu, small);
            int base = 0;
            if (lane == 0 && m) base = atomicAdd(ctr + CTR_NSMALL, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (i < 0) continue;
            if (sz > mx) {
                list[N - 1 - atomicAdd(ctr + CTR_NOVER, 1)] = i;
            } else if (small) {
                list[base + __popc(m & ((1u << lane) - 1u))] = i;
                adj[i] = -1;
                aux[i] = i;  // tfix: optimistic "labelled at its own time"
                stamp[i] = 0;
            } else {
                atomicOr(bits + (i >> 5), 1u << (i & 31));
            }
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

// phase 2: replay the BFS cap on oversized components, one thread each.  T still holds the component
// root for every unassigned member; every piece is classified as it is created (small list / kept bit).
__global__ void __launch_bounds__(128)
cc_split_kernel(const int32_t *__restrict__ lab, int32_t *T, int32_t *psize, int32_t *queue, int32_t *list,
                int32_t *adj, int32_t *aux, int32_t *stamp, uint32_t *bits, uint8_t *flag, int32_t *ctr,
                uint8_t *visit, int64_t N, int H, int W, int64_t min_size, int64_t max_size,
                const int2 *__restrict__ wsizes, int win_px)
{
    const int n_over = ctr[CTR_NOVER];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_over; e += gridDim.x * blockDim.x) {
        const int32_t r = list[N - 1 - e];
        if (wsizes) {
            const int2 ws = wsizes[r / win_px];
            min_size = ws.x;
            max_size = ws.y;
        }
        const int32_t L = lab[r];
        const int32_t n = psize[r];
        const uint8_t cut = flag ? (flag[r] & FLAG_CUT) : 0;
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
                if (visit[s] != 2 || T[s] != r) continue;
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
                // (members still marked 2 keep T == r until their own piece is formed, so the
                //  membership test above stays valid while T is rewritten piece by piece)
                for (int i = 0; i < pc; ++i) {
                    T[qu[i]] = s;
                    visit[qu[i]] = 0;
                }
                psize[s] = pc;
                if (flag) flag[s] = cut;
                if ((int64_t)pc < min_size) {
                    list[atomicAdd(ctr + CTR_NSMALL, 1)] = s;
                    adj[s] = -1;
                    aux[s] = s;
                    stamp[s] = 0;
                } else {
                    atomicOr(bits + (s >> 5), 1u << (s & 31));
                }
            }
        }
    }
}

struct CcParams {
    int H, W;
    int64_t min_size, max_size;
    int32_t mask_label, start_label;
    int32_t optimistic;   // round 1: every piece counts as labelled from its own start pixel
    const int2 *wsizes;   // slab of windows (tiled driver): (min_size, max_size) per window, else NULL
    int32_t win_px;       // pixels per window block of the slab
};

__device__ __forceinline__ int64_t min_size_at(const CcParams &P, int32_t p)
{
    return P.wsizes ? (int64_t)P.wsizes[p / P.win_px].x : P.min_size;
}
__device__ __forceinline__ int64_t max_size_at(const CcParams &P, int32_t p)
{
    return P.wsizes ? (int64_t)P.wsizes[p / P.win_px].y : P.max_size;
}

struct CcArrays {
    const int32_t *lab, *T, *psize, *list;
    int32_t *adj, *aux, *queue, *ctr, *stamp;
    uint8_t *visit;
    uint8_t *flag;        // strip mode only (else NULL): FLAG_* per piece start
};

// Scan position at which pixel q (not in piece t) receives a label > mask label, kTInf if never:
// kept pieces at their start; merged pieces at their start for start_label 0 (they always carry
// a label >= 0), at `tfix` for start_label 1 (label 0 == mask label until a later re-scan).
// Strip mode: `unk` is raised when that time is not known inside the strip (the neighbour's component
// is cut by the strip edge, or it is a merged piece whose own re-scan time is unknown).
// (tq = T[q], -1 on masked pixels.)  The neighbour's flag byte and piece size are loaded by nb_prefetch for the four
// neighbours of a pixel together instead of one dependent round trip after the other.
__device__ __forceinline__ void nb_prefetch(const CcArrays &A, const CcParams &P, int32_t tq, int32_t t, uint8_t &f,
                                            int32_t &ps)
{
    f = 0;
    ps = 0;
    if (tq < 0 || tq == t) return;
    if (A.flag) f = A.flag[tq];
    if (!P.optimistic) ps = A.psize[tq];
}
__device__ __forceinline__ int32_t label_time_pre(const CcArrays &A, const CcParams &P, int32_t tq, int32_t t, uint8_t f,
                                                  int32_t ps, bool &unk)
{
    if (tq < 0 || tq == t) return kTInf;
    if (f & FLAG_CUT) unk = true;
    else if ((f & FLAG_TFIX_UNKNOWN) && P.start_label == 1) unk = true;
    if (P.optimistic) return tq;
    if ((int64_t)ps >= min_size_at(P, tq)) return tq;
    if (P.start_label == 0) return tq;
    return __ldcg(A.aux + tq);
}

// replay of the reference BFS restricted to piece t, started at pixel s.  `unk_any`: some examined
// neighbour has an unknown labelling time; `known_before`: some examined neighbour is known to be
// labelled before s.
__device__ int bfs_piece(const CcArrays &A, int32_t *qu, const CcParams &P, int32_t t, int32_t s,
                         int32_t &adj_out, int32_t &first_time, bool &unk_any, bool &known_before)
{
    int cnt = 1, head = 0;
    int32_t adjq = -1;
    int32_t tmin = kTInf;   // earliest scan position at which any examined neighbour is labelled
    const int64_t max_size = max_size_at(P, t);
    qu[0] = s;
    A.visit[s] = 1;
    while (head < cnt && (int64_t)cnt < max_size) {
        const int32_t p = qu[head];
        const int py = p / P.W, px = p - py * P.W;
        // the four T loads of a pixel do not depend on each other: request them together (the walk is a chain of
        // memory round trips, one thread per piece)
        int32_t qs[4], tqs[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            qs[d] = -1;
            tqs[d] = -1;
            int32_t q;
            if (nbr(d, py, px, P.H, P.W, q)) {
                qs[d] = q;
                tqs[d] = A.T[q];
            }
        }
        uint8_t nf[4];
        int32_t nps[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) nb_prefetch(A, P, tqs[d], t, nf[d], nps[d]);
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int32_t q = qs[d];
            if (q < 0) continue;
            const bool same = tqs[d] == t;   // a piece is a set of equal-label pixels: T alone identifies it
            if (same) {
                if (!A.visit[q]) {
                    A.visit[q] = 1;
                    qu[cnt++] = q;
                    if ((int64_t)cnt >= max_size) break;
                }
            } else {
                bool unk = false;
                const int32_t lt = label_time_pre(A, P, tqs[d], t, nf[d], nps[d], unk);
                if (unk) unk_any = true;
                tmin = min(tmin, lt);
                if (lt < s) {
                    adjq = q;
                    if (!unk) known_before = true;
                }
            }
        }
        ++head;
    }
    adj_out = adjq;
    first_time = tmin;
    return cnt;
}

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
        const int32_t tq = A.T[q];
        if (tq < 0 || tq == t || (int64_t)A.psize[tq] >= min_size_at(P, tq)) continue;
        if (atomicExch(A.stamp + tq, round_id) != round_id) dirty_next[atomicAdd(n_next, 1)] = tq;
    }
}

// Books the result of one small piece (start pixel t): `adjacent`, the strip-mode flags and the re-scan time
// `fix` (aux_old = the value stored so far); queues the small neighbours when the re-scan time changed.
__device__ __forceinline__ bool piece_commit(const CcArrays &A, const CcParams &P, int32_t t, int32_t a, int32_t fix,
                                             int32_t aux_old, int cnt, const int32_t *members, bool unk_any,
                                             bool known_before, int round_id, int32_t *dirty_next, int32_t *n_next)
{
    A.adj[t] = a;
    bool flag_changed = false;
    if (A.flag && unk_any) {
        // conservative: any unknown neighbour makes `adjacent` unknown; the re-scan time stays known
        // when a KNOWN neighbour is labelled before the piece's own start (then tfix == start)
        uint8_t f = A.flag[t];
        uint8_t nf = f | FLAG_ADJ_UNKNOWN;
        if (!known_before && P.start_label == 1) nf |= FLAG_TFIX_UNKNOWN;
        if (nf != f) {
            A.flag[t] = nf;
            flag_changed = ((nf ^ f) & FLAG_TFIX_UNKNOWN) != 0;
        }
    }
    if (aux_old == fix && !flag_changed) return false;
    A.aux[t] = fix;
    if (P.start_label == 0) return true;   // no label-0 ambiguity: nobody depends on tfix
    // tell later small neighbours of the piece to look again
    for (int i = 0; i < cnt; ++i) push_dependents(A, P, members[i], t, round_id, dirty_next, n_next);
    return true;
}

// `adjacent` of one small piece (start pixel t) under the current knowledge of its earlier
// neighbours; `qu` = queue space for max(psize[t], 1) pixels (+ psize[t] more for the re-scan copy,
// claimed from the global cursor on demand).  Returns true when the piece's tfix changed.
__device__ bool small_piece_adjacent(const CcArrays &A, const CcParams &P, int32_t t, int32_t *qu,
                                     int round_id, int32_t *dirty_next, int32_t *n_next)
{
    const int32_t n = A.psize[t];
    int32_t a = -1;
    int32_t fix = t;
    int cnt = 1;
    bool unk_any = false, known_before = false;
    const int64_t max_size = max_size_at(P, t);
    const int32_t *members = qu;   // pixels of the first BFS
    if (n == 1) {
        const int py = t / P.W, px = t - py * P.W;
        // (with max_size <= 1 the reference's BFS loop never runs: no neighbour is looked at)
        int32_t qs[4], tqs[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {       // four independent loads, in flight together
            qs[d] = -1;
            tqs[d] = -1;
            int32_t q;
            if (max_size > 1 && nbr(d, py, px, P.H, P.W, q)) {
                qs[d] = q;
                tqs[d] = A.T[q];
            }
        }
        uint8_t nf[4];
        int32_t nps[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) nb_prefetch(A, P, tqs[d], t, nf[d], nps[d]);
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (qs[d] < 0) continue;
            bool unk = false;
            if (label_time_pre(A, P, tqs[d], t, nf[d], nps[d], unk) < t) {
                a = qs[d];
                if (!unk) known_before = true;
            }
            if (unk) unk_any = true;
        }
        if (a < 0 && P.start_label == 1) fix = kTInf;  // stays label 0: no later pixel to re-enter at
        qu[0] = t;
    } else {
        int32_t tmin;
        cnt = bfs_piece(A, qu, P, t, t, a, tmin, unk_any, known_before);
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
            bool u2 = false, k2 = false;
            if ((int64_t)n < max_size) {
                // The BFS is not cut by the size cap, so from ANY start it examines every
                // neighbour of the piece: the first successful re-scan is at the first member
                // pixel after the earliest labelling time among those neighbours.
                int32_t s = kTInf;
                if (tmin != kTInf)
                    for (int i = 0; i < cnt; ++i)
                        if (cand[i] > tmin && cand[i] < s) s = cand[i];
                if (s != kTInf) {
                    int32_t a2, tm2;
                    const int c2 = bfs_piece(A, qu2, P, t, s, a2, tm2, u2, k2);
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
                    const int c2 = bfs_piece(A, qu2, P, t, s, a2, tm2, u2, k2);
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
    return piece_commit(A, P, t, a, fix, A.aux[t], cnt, members, unk_any, known_before, round_id, dirty_next, n_next);
}

// round 1: every small piece, in parallel (optimistic: every merged neighbour counts as labelled
// from its own start time).  Pieces whose tfix turns out different queue their dependents.
#ifndef OBIA_CC_ADJ_U
#define OBIA_CC_ADJ_U 4       // list entries per lane and trip
#endif
// One thread per piece is a chain of dependent loads (list -> size -> piece starts of the neighbours -> flags),
// and most small pieces of a fragmented raster are single pixels: a lane takes OBIA_CC_ADJ_U list entries per
// trip, walks their one-pixel pieces in lock step (the loads of every step requested together) and then replays
// the larger ones one after the other.
__global__ void __launch_bounds__(128)
cc_small_adjacent_kernel(CcArrays A, CcParams P, int32_t *dirty_next)
{
    constexpr int U = OBIA_CC_ADJ_U;
    const int n_small = A.ctr[CTR_NSMALL];
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * U;
    // warp-uniform trip count: the queue space of a warp's pieces is claimed with one atomic
    for (int64_t e0 = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * U; e0 < n_small; e0 += stride) {
        int32_t t[U], n[U], auxv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = e0 + u * 32 + lane;
            t[u] = (e < n_small) ? A.list[e] : -1;
        }
        int need = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            n[u] = (t[u] >= 0) ? A.psize[t[u]] : 0;
            auxv[u] = (t[u] >= 0) ? A.aux[t[u]] : 0;
            need += n[u];
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
        int32_t *qu = A.queue + base + incl - need;     // this lane's queue space, entry after entry

        // ---- one-pixel pieces, in lock step (this is round 1: P.optimistic, a neighbour piece counts as
        //      labelled from its own start) --------------------------------------------------------------
        int32_t tq[U][4];
        unsigned have[U];      // bit d: neighbour d exists and is looked at
#pragma unroll
        for (int u = 0; u < U; ++u) {
            have[u] = 0;
#pragma unroll
            for (int d = 0; d < 4; ++d) tq[u][d] = -1;
            // (with max_size <= 1 the reference's BFS loop never runs: no neighbour is looked at)
            if (n[u] == 1 && max_size_at(P, t[u]) > 1) {
                const int py = t[u] / P.W, px = t[u] - py * P.W;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    int32_t q;
                    if (nbr(d, py, px, P.H, P.W, q)) {
                        have[u] |= 1u << d;
                        tq[u][d] = A.T[q];
                    }
                }
            }
        }
        uint8_t nf[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int d = 0; d < 4; ++d)
                nf[u][d] = (A.flag && tq[u][d] >= 0 && tq[u][d] != t[u]) ? A.flag[tq[u][d]] : (uint8_t)0;
        int off = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (n[u] == 1) {
                int32_t a = -1;
                bool unk_any = false, known_before = false;
#pragma unroll
                for (int d = 0; d < 4; ++d) {     // reference neighbour order: x+1, x-1, y+1, y-1
                    const int32_t v = tq[u][d];
                    if (!((have[u] >> d) & 1u) || v < 0 || v == t[u]) continue;
                    const bool unk = (nf[u][d] & FLAG_CUT) || ((nf[u][d] & FLAG_TFIX_UNKNOWN) && P.start_label == 1);
                    if (v < t[u]) {
                        a = t[u] + (d == 0 ? 1 : d == 1 ? -1 : d == 2 ? P.W : -P.W);
                        if (!unk) known_before = true;
                    }
                    if (unk) unk_any = true;
                }
                const int32_t fix = (a < 0 && P.start_label == 1) ? kTInf : t[u];   // stays label 0: no later pixel
                qu[off] = t[u];
                piece_commit(A, P, t[u], a, fix, auxv[u], 1, qu + off, unk_any, known_before, 1, dirty_next,
                             A.ctr + CTR_NDIRTY0);
            }
            off += n[u];
        }
        // ---- larger pieces: the BFS replay, one after the other ---------------------------------------------
        // (ONE call site: U inlined copies of the replay do not fit the instruction cache -- 6.2 -> 13 ms)
        int32_t mt[U], mo[U];
        int nm = 0;
        off = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (n[u] > 1) {
                mt[nm] = t[u];
                mo[nm] = off;
                ++nm;
            }
            off += n[u];
        }
#pragma unroll 1
        for (int i = 0; i < nm; ++i) small_piece_adjacent(A, P, mt[i], qu + mo[i], 1, dirty_next, A.ctr + CTR_NDIRTY0);
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
    if (blockIdx.x == 0 && threadIdx.x == 0) A.ctr[CTR_ROUNDS] = round_id;
    const int lane = threadIdx.x & 31;
    // A short list (the usual case: the tail of a strip's unknown-flag propagation is a few dozen pieces per
    // round) is spread ONE PIECE PER WARP: 32 different BFS replays inside one warp take turns at its issue
    // slot, and the round lasts as long as its slowest warp.
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const bool sparse = n <= nwarps;
    const int stride = sparse ? nwarps : gridDim.x * blockDim.x;
    const int first = sparse ? blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)
                             : blockIdx.x * blockDim.x + (threadIdx.x & ~31);
    for (int e0 = first; e0 < n; e0 += stride) {
        const int e = sparse ? e0 : e0 + lane;
        const bool active = sparse ? (lane == 0) : (e < n);
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

__global__ void cc_reset_round_kernel(int32_t *ctr, int nxt_ctr)
{
    ctr[nxt_ctr] = 0;
    ctr[CTR_CURSOR] = 0;   // queue space is recycled
}

// phase 4a/b/c: kept pieces are numbered by the rank of their start pixel = prefix population count
// of the start-pixel bitmap (N / 8 bytes instead of two passes over T and psize)
__global__ void __launch_bounds__(256)
cc_bits_count_kernel(const uint32_t *__restrict__ bits, int32_t *chunksum, int64_t valid_lo, int64_t core_lo,
                     int64_t core_hi, int32_t *ctr)
{
    __shared__ int s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t w0 = (int64_t)blockIdx.x * kBitChunk;
    int c = 0, before = 0, core = 0, valid = 0;
    for (int j = threadIdx.x; j < kBitChunk; j += 256) {
        const uint32_t v = bits[w0 + j];
        c += __popc(v);
        // kept pieces that start before / inside the core rows of a strip (pixel range [core_lo, core_hi));
        // `valid`: before the core but below the unknown band of the upper halo
        const int64_t p0 = (w0 + j) * 32;
        if (v) {
            for (uint32_t m = v; m; m &= m - 1) {
                const int64_t px = p0 + __ffs(m) - 1;
                before += px < core_lo;
                valid += (px >= valid_lo && px < core_lo);
                core += (px >= core_lo && px < core_hi);
            }
        }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    before = __reduce_add_sync(0xffffffffu, before);
    core = __reduce_add_sync(0xffffffffu, core);
    valid = __reduce_add_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0) {
        if (c) atomicAdd(&s_cnt[0], c);
        if (before) atomicAdd(&s_cnt[1], before);
        if (core) atomicAdd(&s_cnt[2], core);
        if (valid) atomicAdd(&s_cnt[3], valid);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        chunksum[blockIdx.x] = s_cnt[0];
        if (s_cnt[1]) atomicAdd(ctr + CTR_KBEFORE, s_cnt[1]);
        if (s_cnt[2]) atomicAdd(ctr + CTR_KCORE, s_cnt[2]);
        if (s_cnt[3]) atomicAdd(ctr + CTR_KVALID, s_cnt[3]);
    }
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

// label of every kept piece: start_label + label_offset + rank of its start pixel (label_offset != 0
// only for strips: global rank of the strip's first core piece minus its local rank)
__global__ void __launch_bounds__(256)
cc_bits_number_kernel(const uint32_t *__restrict__ bits, const int32_t *__restrict__ chunksum, int32_t *fin,
                      int32_t label_base)
{
    __shared__ int s_warp[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w0 = (int64_t)blockIdx.x * kBitChunk;
    int running = chunksum[blockIdx.x];
    for (int j0 = 0; j0 < kBitChunk; j0 += 256) {
        const uint32_t v = bits[w0 + j0 + threadIdx.x];
        const int c = __popc(v);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int before = running + incl - c;
        int tot = 0;
        for (int w = 0; w < 8; ++w) {
            if (w < warp) before += s_warp[w];
            tot += s_warp[w];
        }
        const int64_t p0 = (w0 + j0 + threadIdx.x) * 32;
        int k = 0;
        for (uint32_t m = v; m; m &= m - 1) fin[p0 + __ffs(m) - 1] = label_base + before + k++;
        running += tot;
        __syncthreads();
    }
}

// phase 4d: merged pieces follow their adjacent chain to a kept piece (or to 0, the initial value of
// `adjacent`); strip mode also records whether anything on the chain is unknown inside the strip
__global__ void __launch_bounds__(256)
cc_small_final_kernel(CcArrays A, const uint32_t *__restrict__ bits, int32_t *fin, int max_hops, int32_t leftover)
{
    const int n_small = A.ctr[CTR_NSMALL];
    // Four pieces per thread, hop by hop: every hop is a chain of dependent loads (kept bit -> adjacent or
    // label -> piece start of the adjacent pixel), so the loads of the four pieces are requested together.
    constexpr int U = 4;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int32_t cutmin = A.flag ? A.ctr[CTR_CUTMIN] : kTInf;
    for (int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < n_small; e0 += nthreads * U) {
        int32_t t0[U], t[U], r[U];
        bool unk[U], done[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = e0 + u * nthreads;
            t0[u] = (e < n_small) ? A.list[e] : -1;
            t[u] = t0[u];
            r[u] = 0;
            unk[u] = false;
            done[u] = t0[u] < 0;
        }
        for (int hop = 0; hop < max_hops; ++hop) {   // chains are short; never spin
            if (done[0] && done[1] && done[2] && done[3]) break;
            uint32_t bw[U];
#pragma unroll
            for (int u = 0; u < U; ++u) bw[u] = done[u] ? 0u : bits[t[u] >> 5];
            int32_t val[U];
            uint8_t fl[U];
            bool kept[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                kept[u] = false;
                val[u] = -1;
                fl[u] = 0;
                if (done[u]) continue;
                kept[u] = (bw[u] >> (t[u] & 31)) & 1u;
                val[u] = kept[u] ? fin[t[u]] : A.adj[t[u]];
                if (A.flag) fl[u] = A.flag[t[u]];
            }
            int32_t nxt[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                nxt[u] = -1;
                if (done[u]) continue;
                if (kept[u]) {
                    r[u] = val[u];
                    if (A.flag && ((fl[u] & FLAG_CUT) || t[u] >= cutmin)) unk[u] = true;
                    done[u] = true;
                } else {
                    if (fl[u] & (FLAG_CUT | FLAG_ADJ_UNKNOWN)) unk[u] = true;
                    if (val[u] < 0) {
                        r[u] = leftover;  // `adjacent` initial value: label 0 (of the window, see the windows entry)
                        done[u] = true;
                        A.ctr[CTR_HASZERO] = 1;   // a leftover piece carries label 0 (benign race: same value)
                    } else {
                        nxt[u] = val[u];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (nxt[u] >= 0) t[u] = A.T[nxt[u]];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (t0[u] < 0) continue;
            if (!done[u]) atomicExch(A.ctr + CTR_ERR, 2);
            fin[t0[u]] = r[u];            // kept and merged piece starts are disjoint: one table for both
            if (unk[u]) A.flag[t0[u]] |= FLAG_LABEL_UNKNOWN;
        }
    }
}

// phase 4e: final labels of rows [row_lo, row_hi) of the strip, written to `out` (row 0 of out = row_lo)
__global__ void __launch_bounds__(256)
cc_resolve_kernel(const int32_t *__restrict__ T, const uint32_t *__restrict__ bits, const int32_t *__restrict__ fin,
                  const uint8_t *__restrict__ flag, int32_t *__restrict__ out, int32_t *ctr, int64_t p_lo,
                  int64_t p_hi, int32_t mask_label)
{
    const int64_t i = p_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p_hi) return;
    const int32_t t = T[i];
    int32_t r = mask_label;
    if (t >= 0) {
        r = fin[t];
        if (flag) {
            const bool kept = (bits[t >> 5] >> (t & 31)) & 1u;
            const uint8_t f = flag[t];
            if (kept ? ((f & FLAG_CUT) || t >= ctr[CTR_CUTMIN]) : (f & (FLAG_CUT | FLAG_LABEL_UNKNOWN)))
                atomicExch(ctr + CTR_FAIL, 1);
        }
    }
    out[i - p_lo] = r;
}

// strip mode: components with a pixel in the outer band of an open halo (its outer third) may continue
// outside the strip, or depend on pieces that do: they are all booked unknown (FLAG_CUT).  A band
// instead of the edge row alone keeps the small-piece fixed point short: under a single cut row the
// "unknown re-scan time" flag crawls from fragment to fragment across the whole strip width.
__global__ void __launch_bounds__(256)
cc_mark_cut_kernel(const int32_t *__restrict__ T, uint8_t *flag, int32_t *ctr, int64_t N, int W, int band_top,
                   int band_bottom)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_top = (int64_t)band_top * W, n_bot = (int64_t)band_bottom * W;
    if (i < n_top) {
        const int32_t t = T[i];
        if (t >= 0) flag[t] = FLAG_CUT;
    } else if (i < n_top + n_bot) {
        const int32_t t = T[N - n_bot + (i - n_top)];
        if (t >= 0) {
            flag[t] = FLAG_CUT;
            // kept pieces that start after the first bottom-band component have an unreliable rank (band
            // components are booked as kept whatever their true size)
            atomicMin(ctr + CTR_CUTMIN, t);
        }
    }
}

__global__ void cc_init_ctr_kernel(int32_t *ctr) { ctr[CTR_CUTMIN] = kTInf; }

}  // namespace obia

using namespace obia;

extern "C" int64_t obia_b200_connectivity_workspace_bytes(int64_t H, int64_t W)
{
    if (H <= 0 || W <= 0) return -1;
    return cc_ws_layout(nullptr, H * W).bytes;
}

namespace {

struct CcRun {
    const int32_t *lab;
    CcWs w;
    CcParams P;
    int64_t N;
    int32_t *fin;
    bool strip;
    cudaStream_t st;
};

// phases 1-3: components, split, small-piece adjacency (round 1 + two speculative rounds), counts
int cc_phase_a(CcRun &R, int top_open, int bottom_open, int64_t core_lo, int64_t core_hi)
{
    const CcWs &w = R.w;
    const CcParams &P0 = R.P;
    const int64_t N = R.N;
    cudaStream_t st = R.st;
    const int H = P0.H, W = P0.W;
    const unsigned gridN = (unsigned)ceil_div(N, 256);
    OBIA_CUDA_CHECK(cudaMemsetAsync(w.ctr, 0, CTR_WORDS * 4, st));
    cc_init_ctr_kernel<<<1, 1, 0, st>>>(w.ctr);
    OBIA_LAUNCH_CHECK();
    OBIA_CUDA_CHECK(cudaMemsetAsync(w.visit, 0, (size_t)N, st));
    OBIA_CUDA_CHECK(cudaMemsetAsync(w.bits, 0, (size_t)w.nchunks * kBitChunk * 4, st));
    if (R.strip) OBIA_CUDA_CHECK(cudaMemsetAsync(w.flag, 0, (size_t)N, st));
    {
        dim3 tiles((unsigned)ceil_div(W, kTile), (unsigned)ceil_div(H, kTile));
        cc_local_kernel<<<tiles, 256, 0, st>>>(R.lab, w.T, w.psize, H, W, P0.mask_label);
        OBIA_LAUNCH_CHECK();
        const int64_t nborder = (int64_t)((W - 1) / kTile) * H + (int64_t)((H - 1) / kTile) * W;
        if (nborder > 0) {
            cc_border_kernel<<<(unsigned)ceil_div(nborder, 256), 256, 0, st>>>(R.lab, w.T, H, W, P0.mask_label);
            OBIA_LAUNCH_CHECK();
        }
    }
    int32_t *roots = w.queue;   // idle until the split / adjacency kernels
    cc_flatten_kernel<<<(unsigned)ceil_div(N, 1024), 256, 0, st>>>(w.T, w.psize, roots, w.ctr, N);
    OBIA_LAUNCH_CHECK();
    if (R.strip && (top_open || bottom_open)) {
        // outer third of each open halo (core rows are [core_lo / W, core_hi / W))
        const int halo_top = (int)(core_lo / W), halo_bottom = (int)((N - core_hi) / W);
        const int band_top = top_open ? std::max(1, halo_top / 3) : 0;
        const int band_bottom = bottom_open ? std::max(1, halo_bottom / 3) : 0;
        const int64_t n = (int64_t)(band_top + band_bottom) * W;
        cc_mark_cut_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(w.T, w.flag, w.ctr, N, W, band_top, band_bottom);
        OBIA_LAUNCH_CHECK();
    }
    cc_classify_kernel<<<kNumSMs * 8, 256, 0, st>>>(roots, w.psize, w.list, w.adj, w.aux, w.stamp, w.bits,
                                                    R.strip ? w.flag : nullptr, w.ctr, N, P0.min_size, P0.max_size,
                                                    P0.wsizes, P0.win_px);
    OBIA_LAUNCH_CHECK();
    cc_split_kernel<<<kNumSMs * 2, 128, 0, st>>>(R.lab, w.T, w.psize, w.queue, w.list, w.adj, w.aux, w.stamp, w.bits,
                                                 R.strip ? w.flag : nullptr, w.ctr, w.visit, N, H, W, P0.min_size,
                                                 P0.max_size, P0.wsizes, P0.win_px);
    OBIA_LAUNCH_CHECK();

    CcArrays A;
    A.lab = R.lab; A.T = w.T; A.psize = w.psize; A.list = w.list;
    A.adj = w.adj; A.aux = w.aux; A.queue = w.queue; A.ctr = w.ctr; A.stamp = w.stamp; A.visit = w.visit;
    A.flag = R.strip ? w.flag : nullptr;
    CcParams P = P0;
    cc_reset_round_kernel<<<1, 1, 0, st>>>(w.ctr, CTR_NDIRTY0);
    OBIA_LAUNCH_CHECK();
    P.optimistic = 1;
    cc_small_adjacent_kernel<<<kNumSMs * OBIA_CC_ADJ_CTAS, 128, 0, st>>>(A, P, w.dirty0);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

// rounds 2..: `n` launches without a host round trip (an empty round costs a few microseconds);
// `cur` = index of the dirty list written last
int cc_rounds(CcRun &R, int n, int &round_id, int &cur)
{
    const CcWs &w = R.w;
    CcArrays A;
    A.lab = R.lab; A.T = w.T; A.psize = w.psize; A.list = w.list;
    A.adj = w.adj; A.aux = w.aux; A.queue = w.queue; A.ctr = w.ctr; A.stamp = w.stamp; A.visit = w.visit;
    A.flag = R.strip ? w.flag : nullptr;
    CcParams P = R.P;
    P.optimistic = 0;
    for (int b = 0; b < n; ++b) {
        ++round_id;
        const int cur_ctr = cur ? CTR_NDIRTY1 : CTR_NDIRTY0, nxt_ctr = cur ? CTR_NDIRTY0 : CTR_NDIRTY1;
        cc_reset_round_kernel<<<1, 1, 0, R.st>>>(w.ctr, nxt_ctr);
        OBIA_LAUNCH_CHECK();
        cc_small_round_kernel<<<kNumSMs * 4, 128, 0, R.st>>>(A, P, cur ? w.dirty1 : w.dirty0,
                                                             cur ? w.dirty0 : w.dirty1, cur_ctr, nxt_ctr, round_id);
        OBIA_LAUNCH_CHECK();
        cur ^= 1;
    }
    return OBIA_B200_OK;
}

int cc_count(CcRun &R, int64_t core_lo, int64_t core_hi, int64_t valid_lo = 0)
{
    const CcWs &w = R.w;
    OBIA_CUDA_CHECK(cudaMemsetAsync(w.ctr + CTR_KBEFORE, 0, 8, R.st));
    OBIA_CUDA_CHECK(cudaMemsetAsync(w.ctr + CTR_KVALID, 0, 4, R.st));
    cc_bits_count_kernel<<<(unsigned)w.nchunks, 256, 0, R.st>>>(w.bits, w.chunksum, valid_lo, core_lo, core_hi, w.ctr);
    OBIA_LAUNCH_CHECK();
    cc_scan_blocks_kernel<<<1, 1024, 0, R.st>>>(w.chunksum, w.nchunks, w.ctr);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

// phase 4: numbering + final labels of pixel range [p_lo, p_hi)
int cc_phase_b(CcRun &R, int32_t label_base, int32_t *out, int64_t p_lo, int64_t p_hi, int max_hops)
{
    const CcWs &w = R.w;
    OBIA_CUDA_CHECK(cudaMemsetAsync(w.ctr + CTR_ERR, 0, 4, R.st));
    OBIA_CUDA_CHECK(cudaMemsetAsync(w.ctr + CTR_FAIL, 0, 8, R.st));   // CTR_FAIL, CTR_HASZERO
    CcArrays A;
    A.lab = R.lab; A.T = w.T; A.psize = w.psize; A.list = w.list;
    A.adj = w.adj; A.aux = w.aux; A.queue = w.queue; A.ctr = w.ctr; A.stamp = w.stamp; A.visit = w.visit;
    A.flag = R.strip ? w.flag : nullptr;
    cc_bits_number_kernel<<<(unsigned)w.nchunks, 256, 0, R.st>>>(w.bits, w.chunksum, R.fin, label_base);
    OBIA_LAUNCH_CHECK();
    // slab of windows with start_label 0: "label 0" is the first kept piece of the piece's own WINDOW, which the
    // caller resolves (marker -2); with start_label 1 it is the mask label 0 everywhere
    const int32_t leftover = (R.P.wsizes && R.P.start_label == 0) ? -2 : 0;
    cc_small_final_kernel<<<kNumSMs * OBIA_CC_FIN_CTAS, 256, 0, R.st>>>(A, w.bits, R.fin, max_hops, leftover);
    OBIA_LAUNCH_CHECK();
    cc_resolve_kernel<<<(unsigned)ceil_div(p_hi - p_lo, 256), 256, 0, R.st>>>(
        w.T, w.bits, R.fin, R.strip ? w.flag : nullptr, out, w.ctr, p_lo, p_hi, R.P.mask_label);
    OBIA_LAUNCH_CHECK();
    return OBIA_B200_OK;
}

}  // namespace

static int enforce_connectivity_impl(const int32_t *labels_in, int32_t *labels_out, void *workspace, int64_t H,
                                     int64_t W, int64_t min_size, int64_t max_size, const int2 *wsizes,
                                     int64_t win_rows, int32_t start_label, int64_t *n_labels_host, void *stream)
{
    if (!labels_in || !labels_out || !workspace || H <= 0 || W <= 0)
        return set_err(OBIA_B200_ERR_ARG, "enforce_connectivity: bad argument");
    if (max_size < 1) return set_err(OBIA_B200_ERR_ARG, "enforce_connectivity: max_size must be >= 1");
    if (start_label != 0 && start_label != 1) return set_err(OBIA_B200_ERR_ARG, "start_label should be 0 or 1.");
    const int64_t N = H * W;
    if (N >= 0x7fffffffLL) return set_err(OBIA_B200_ERR_UNSUPPORTED, "enforce_connectivity: H*W exceeds int32");
    CcRun R;
    R.lab = labels_in;
    R.w = cc_ws_layout(workspace, N);
    R.N = N;
    R.fin = R.w.queue + N;   // second half of the queue: only the re-scan copies of a later round reach it
    R.strip = false;
    R.st = (cudaStream_t)stream;
    R.P.H = (int)H; R.P.W = (int)W; R.P.min_size = min_size; R.P.max_size = max_size;
    R.P.mask_label = start_label - 1; R.P.start_label = start_label; R.P.optimistic = 0;
    R.P.wsizes = wsizes; R.P.win_px = (int32_t)std::max<int64_t>(1, win_rows * W);
    int rc = cc_phase_a(R, 0, 0, 0, N);
    if (rc) return rc;
    // Rounds 2.. exist only for start_label 1 (label-0 chains) and are almost always empty: two are
    // enqueued speculatively and the result is finished without a host round trip (merge chains walked
    // for at most kSpecHops hops); the single read-back at the end tells whether more rounds, or a
    // second finish with unbounded chains, are needed.
    constexpr int kSpecHops = 256, kAllHops = 1 << 30;
    int round_id = 1, cur = 0;
    int32_t hctr[CTR_WORDS];
    auto read_back = [&]() -> int {
        OBIA_CUDA_CHECK(cudaMemcpyAsync(hctr, R.w.ctr, sizeof(hctr), cudaMemcpyDeviceToHost, R.st));
        OBIA_CUDA_CHECK(cudaStreamSynchronize(R.st));
        return OBIA_B200_OK;
    };
    if (start_label == 1) {
        rc = cc_rounds(R, 2, round_id, cur);
        if (rc) return rc;
    }
    rc = cc_count(R, 0, N);
    if (!rc) rc = cc_phase_b(R, start_label, labels_out, 0, N, kSpecHops);
    if (!rc) rc = read_back();
    if (rc) return rc;
    bool dirty = start_label == 1 && hctr[cur ? CTR_NDIRTY1 : CTR_NDIRTY0] != 0;
    if (dirty || hctr[CTR_ERR]) {
        while (dirty) {   // label times were still changing: rounds in batches until the list is empty
            rc = cc_rounds(R, 16, round_id, cur);
            if (!rc) rc = read_back();
            if (rc) return rc;
            dirty = hctr[cur ? CTR_NDIRTY1 : CTR_NDIRTY0] != 0;
            if (round_id > (1 << 28)) return set_err(OBIA_B200_ERR_CUDA, "enforce_connectivity: no fixed point");
        }
        rc = cc_count(R, 0, N);
        if (!rc) rc = cc_phase_b(R, start_label, labels_out, 0, N, kAllHops);
        if (!rc) rc = read_back();
        if (rc) return rc;
    }
    if (n_labels_host) *n_labels_host = hctr[CTR_NKEPT];
    if (hctr[CTR_ERR])
        return set_err(OBIA_B200_ERR_CUDA, "enforce_connectivity: internal consistency check failed (%d)",
                       hctr[CTR_ERR]);
    return OBIA_B200_OK;
}

extern "C" int obia_b200_enforce_connectivity(const int32_t *labels_in, int32_t *labels_out,
                                              void *workspace, int64_t H, int64_t W, int64_t min_size,
                                              int64_t max_size, int32_t start_label,
                                              int64_t *n_labels_host, void *stream)
{
    return enforce_connectivity_impl(labels_in, labels_out, workspace, H, W, min_size, max_size, nullptr, 0, start_label,
                                     n_labels_host, stream);
}

// Slab of windows (tiled driver, batch.cuh): rows [i * win_rows, (i + 1) * win_rows) belong to window i, whose
// (min_size, max_size) are window_sizes[i] (device, int32 pairs); windows are separated by mask-label rows, so
// the result inside every window is the one obia_b200_enforce_connectivity gives on that window alone, with the
// kept pieces numbered start_label.. in slab raster order (= window by window).  start_label 0: pieces the
// reference merges into its `adjacent = 0` default (label 0 = the window's FIRST kept piece) are written as -2.
extern "C" int obia_b200_enforce_connectivity_windows(const int32_t *labels_in, int32_t *labels_out, void *workspace,
                                                      int64_t H, int64_t W, const int32_t *window_sizes,
                                                      int64_t win_rows, int32_t start_label,
                                                      int64_t *n_labels_host, void *stream)
{
    if (!window_sizes || win_rows <= 0 || H % win_rows)
        return set_err(OBIA_B200_ERR_ARG, "enforce_connectivity_windows: bad argument");
    return enforce_connectivity_impl(labels_in, labels_out, workspace, H, W, 1, 1, (const int2 *)window_sizes, win_rows,
                                     start_label, n_labels_host, stream);
}

// ---- strip mode (one raster sharded by row strips across GPUs) ----------------------------------------
static int cc_strip_setup(CcRun &R, const int32_t *labels_ext, void *workspace, int64_t H_ext, int64_t W,
                          int64_t core_row0, int64_t core_rows, int64_t min_size, int64_t max_size,
                          int32_t start_label, void *stream)
{
    if (!labels_ext || !workspace || H_ext <= 0 || W <= 0 || core_row0 < 0 || core_rows <= 0 ||
        core_row0 + core_rows > H_ext)
        return set_err(OBIA_B200_ERR_ARG, "connectivity_strip: bad argument");
    if (max_size < 1) return set_err(OBIA_B200_ERR_ARG, "connectivity_strip: max_size must be >= 1");
    if (start_label != 0 && start_label != 1) return set_err(OBIA_B200_ERR_ARG, "start_label should be 0 or 1.");
    const int64_t N = H_ext * W;
    if (N >= 0x7fffffffLL) return set_err(OBIA_B200_ERR_UNSUPPORTED, "connectivity_strip: H_ext*W exceeds int32");
    R.lab = labels_ext;
    R.w = cc_ws_layout(workspace, N);
    R.N = N;
    R.fin = R.w.queue + N;
    R.strip = true;
    R.st = (cudaStream_t)stream;
    R.P.H = (int)H_ext; R.P.W = (int)W; R.P.min_size = min_size; R.P.max_size = max_size;
    R.P.mask_label = start_label - 1; R.P.start_label = start_label; R.P.optimistic = 0;
    R.P.wsizes = nullptr; R.P.win_px = 1;
    return OBIA_B200_OK;
}

extern "C" int obia_b200_connectivity_strip_begin(const int32_t *labels_ext, void *workspace, int64_t H_ext,
                                                  int64_t W, int64_t core_row0, int64_t core_rows,
                                                  int32_t top_open, int32_t bottom_open, int64_t min_size,
                                                  int64_t max_size, int32_t start_label, int64_t *counts_host,
                                                  void *stream)
{
    CcRun R;
    int rc = cc_strip_setup(R, labels_ext, workspace, H_ext, W, core_row0, core_rows, min_size, max_size,
                            start_label, stream);
    if (rc) return rc;
    if (!counts_host) return set_err(OBIA_B200_ERR_ARG, "connectivity_strip_begin: bad argument");
    const int64_t core_lo = core_row0 * W, core_hi = (core_row0 + core_rows) * W;
    rc = cc_phase_a(R, top_open, bottom_open, core_lo, core_hi);
    if (rc) return rc;
    int round_id = 1, cur = 0;
    int32_t hctr[CTR_WORDS];
    const bool debug = getenv("OBIA_B200_DEBUG") != nullptr;   // read once per call, outside the round loop
    while (true) {
        if (debug) {
            OBIA_CUDA_CHECK(cudaMemcpyAsync(hctr, R.w.ctr, sizeof(hctr), cudaMemcpyDeviceToHost, R.st));
            OBIA_CUDA_CHECK(cudaStreamSynchronize(R.st));
            fprintf(stderr, "[obia_b200] strip connectivity: round %d, %d small pieces, %d over, dirty %d / %d, roots %d\n",
                    round_id, hctr[CTR_NSMALL], hctr[CTR_NOVER], hctr[CTR_NDIRTY0], hctr[CTR_NDIRTY1], hctr[CTR_NROOTS]);
        }
        if (start_label == 1) {
            rc = cc_rounds(R, round_id == 1 ? 12 : (debug ? 1 : 16), round_id, cur);   // (an empty round costs a few microseconds)
            if (rc) return rc;
        }
        // (same band as cc_phase_a marks: the outer third of an open upper halo)
        rc = cc_count(R, core_lo, core_hi, top_open ? (int64_t)std::max<int64_t>(1, core_row0 / 3) * W : 0);
        if (rc) return rc;
        OBIA_CUDA_CHECK(cudaMemcpyAsync(hctr, R.w.ctr, sizeof(hctr), cudaMemcpyDeviceToHost, R.st));
        OBIA_CUDA_CHECK(cudaStreamSynchronize(R.st));
        if (start_label == 0 || hctr[cur ? CTR_NDIRTY1 : CTR_NDIRTY0] == 0) break;
        if (round_id > (1 << 28)) return set_err(OBIA_B200_ERR_CUDA, "connectivity_strip: no fixed point");
    }
    counts_host[0] = hctr[CTR_KBEFORE];
    counts_host[1] = hctr[CTR_KCORE];
    counts_host[2] = hctr[CTR_NKEPT];
    counts_host[3] = round_id;
    counts_host[4] = hctr[CTR_KVALID];
    return OBIA_B200_OK;
}

extern "C" int obia_b200_connectivity_strip_finish(const int32_t *labels_ext, int32_t *labels_out_core,
                                                   void *workspace, int64_t H_ext, int64_t W, int64_t core_row0,
                                                   int64_t core_rows, int64_t min_size, int64_t max_size,
                                                   int32_t start_label, int64_t label_offset,
                                                   int32_t *incomplete_host, void *stream)
{
    CcRun R;
    int rc = cc_strip_setup(R, labels_ext, workspace, H_ext, W, core_row0, core_rows, min_size, max_size,
                            start_label, stream);
    if (rc) return rc;
    if (!labels_out_core || !incomplete_host)
        return set_err(OBIA_B200_ERR_ARG, "connectivity_strip_finish: bad argument");
    const int64_t core_lo = core_row0 * W, core_hi = (core_row0 + core_rows) * W;
    rc = cc_phase_b(R, (int32_t)(start_label + label_offset), labels_out_core, core_lo, core_hi, 1 << 30);
    if (rc) return rc;
    int32_t hctr[CTR_WORDS];
    OBIA_CUDA_CHECK(cudaMemcpyAsync(hctr, R.w.ctr, sizeof(hctr), cudaMemcpyDeviceToHost, R.st));
    OBIA_CUDA_CHECK(cudaStreamSynchronize(R.st));
    incomplete_host[0] = hctr[CTR_FAIL];
    incomplete_host[1] = hctr[CTR_HASZERO];
    if (hctr[CTR_ERR])
        return set_err(OBIA_B200_ERR_CUDA, "connectivity_strip: internal consistency check failed (%d)", hctr[CTR_ERR]);
    return OBIA_B200_OK;
}
