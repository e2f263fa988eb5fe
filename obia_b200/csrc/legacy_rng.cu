// Host-side helper of the maskSLIC initialisation: the two sample draws of scikit-image's
// `_get_mask_centroids` (reached from obia/segmentation/segment_boundaries.py:51 whenever `mask=` is
// passed, i.e. for every tile of obia/utils/tiling.py:137-143, :275-281):
//
//     rng = np.random.RandomState(123)
//     idx       = np.sort(rng.choice(np.arange(n_coord), min(n_segments, n_coord), replace=False))
//     idx_dense = np.sort(rng.choice(np.arange(n_coord), min(100 * n_segments, n_coord), replace=False))
//
// numpy's legacy `choice(replace=False)` is `permutation(n)[:k]`, `permutation` is `arange` + the legacy
// `shuffle`: for i = n-1 .. 1 swap x[i] with x[random_interval(i)], where `random_interval` draws 32-bit
// MT19937 outputs masked to the bit length of i until one is <= i (64-bit outputs beyond 2^32).  The
// algorithm is inherently sequential in i, O(n_coord) per draw, and the result depends only on
// (n_coord, n_segments) -- restated here in C++ so that the tiled driver can run it for many tiles in
// parallel host threads (ctypes releases the GIL) while the GPU works on earlier tiles.  No CUDA here.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "common.cuh"

namespace {

struct MT19937 {
    uint32_t key[624];
    int pos;
    explicit MT19937(uint32_t seed)
    {
        // numpy `_legacy_seeding(int)` -> mt19937_seed == init_genrand
        for (int i = 0; i < 624; ++i) {
            key[i] = seed;
            seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
        }
        pos = 624;
    }
    void gen()
    {
        constexpr uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, A = 0x9908b0dfu;
        int i = 0;
        for (; i < 624 - 397; ++i) {
            const uint32_t y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        }
        for (; i < 623; ++i) {
            const uint32_t y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        }
        const uint32_t y = (key[623] & UPPER) | (key[0] & LOWER);
        key[623] = key[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        pos = 0;
    }
    uint32_t next32()
    {
        if (pos == 624) gen();
        uint32_t y = key[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    uint64_t next64() { const uint64_t hi = next32(); return (hi << 32) | next32(); }
    // numpy legacy `random_interval`: uniform integer in [0, max]
    uint64_t interval(uint64_t max)
    {
        if (max == 0) return 0;
        uint64_t mask = max;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
        mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
        uint64_t v;
        if (max <= 0xffffffffull) {
            while ((v = (next32() & mask)) > max) {}
        } else {
            while ((v = (next64() & mask)) > max) {}
        }
        return v;
    }
};

// The same draw for n < 2^32, tuned (the tiled driver makes one draw per window: tens of thousands of
// O(n_coord) sequential shuffles): tempered outputs are produced 624 at a time, the rejection mask of
// `random_interval(i)` is carried along while i falls, and the k selected values are collected through a
// bitmap instead of a sort.  `work` / `bits` are scratch reused between draws.
struct MTBlock {
    MT19937 &g;
    uint32_t out[624];
    int pos;
    explicit MTBlock(MT19937 &gen) : g(gen), pos(624)
    {
        // continue exactly where the scalar interface stopped
        if (g.pos < 624) {
            for (int i = g.pos; i < 624; ++i) out[i] = temper(g.key[i]);
            pos = g.pos;
        }
    }
    static uint32_t temper(uint32_t y)
    {
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    void refill()
    {
        g.gen();
        for (int i = 0; i < 624; ++i) out[i] = temper(g.key[i]);
        pos = 0;
    }
    void done() { g.pos = pos; }   // hand the stream position back
};

void choice_sorted_u32(MT19937 &rng, uint32_t n, uint32_t k, int64_t *outv, std::vector<uint32_t> &work,
                       std::vector<uint64_t> &bits)
{
    work.resize(n);
    uint32_t *x = work.data();
    for (uint32_t i = 0; i < n; ++i) x[i] = i;
    MTBlock b(rng);
    if (n > 1) {
        uint32_t mask = n - 1;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        // Two passes per block of generator outputs, both free of data-dependent branches (the rejection test
        // `v > i` of random_interval is unpredictable): (1) walk the outputs, keep the accepted values -- output
        // v serves index i iff (v & mask) <= i, then i falls by one; (2) do the swaps x[i] <-> x[v].
        uint32_t acc[624];
        uint32_t i = n - 1;
        while (i >= 1) {
            if (b.pos == 624) b.refill();
            const int avail = 624 - b.pos;
            const uint32_t *o = b.out + b.pos;
            const uint32_t i0 = i;
            uint32_t cnt = 0;
            int c = 0;
            for (; c < avail && i >= 1; ++c) {
                while ((mask >> 1) >= i) mask >>= 1;      // smallest 2^b - 1 that is >= i
                const uint32_t v = o[c] & mask;
                const uint32_t ok = v <= i;
                acc[cnt] = v;
                cnt += ok;
                i -= ok;
            }
            b.pos += c;
            for (uint32_t q = 0; q < cnt; ++q) {
                const uint32_t ii = i0 - q, v = acc[q];
                const uint32_t t = x[ii];
                x[ii] = x[v];
                x[v] = t;
            }
        }
    }
    b.done();
    if (k <= 64) {
        std::sort(x, x + k);
        for (uint32_t i = 0; i < k; ++i) outv[i] = x[i];
        return;
    }
    bits.assign(((size_t)n + 63) / 64, 0ull);
    for (uint32_t i = 0; i < k; ++i) bits[x[i] >> 6] |= 1ull << (x[i] & 63);
    size_t o = 0;
    for (size_t w = 0; w < bits.size(); ++w) {
        uint64_t m = bits[w];
        while (m) {
            outv[o++] = (int64_t)(w * 64 + (size_t)__builtin_ctzll(m));
            m &= m - 1;
        }
    }
}

// first k entries of RandomState.permutation(n), sorted
template <typename T>
void choice_sorted(MT19937 &rng, int64_t n, int64_t k, int64_t *out)
{
    std::vector<T> x((size_t)n);
    for (int64_t i = 0; i < n; ++i) x[(size_t)i] = (T)i;
    for (int64_t i = n - 1; i >= 1; --i) {
        const int64_t j = (int64_t)rng.interval((uint64_t)i);
        std::swap(x[(size_t)i], x[(size_t)j]);
    }
    std::sort(x.begin(), x.begin() + k);
    for (int64_t i = 0; i < k; ++i) out[i] = (int64_t)x[(size_t)i];
}

}  // namespace

extern "C" int obia_b200_mask_sample_indices(int64_t n_coord, int64_t n_segments, int64_t *idx,
                                             int64_t *idx_dense)
{
    const int64_t k1 = std::min(n_segments, n_coord);
    const int64_t k2 = std::min((int64_t)100 * n_segments, n_coord);
    // idx_dense may be NULL when the second draw selects every pixel (k2 == n_coord): nothing to compute
    if (n_coord <= 0 || n_segments <= 0 || !idx || (!idx_dense && k2 != n_coord))
        return obia::set_err(OBIA_B200_ERR_ARG, "mask_sample_indices: bad argument");
    MT19937 rng(123u);
    // the second draw selects every pixel when 100 * n_segments >= n_coord: sorted, that is 0 .. n_coord-1
    // whatever the permutation was (the generator is not used afterwards)
    const bool dense_all = k2 == n_coord;
    if (dense_all && idx_dense)
        for (int64_t i = 0; i < n_coord; ++i) idx_dense[i] = i;
    if (n_coord <= 0x7fffffffLL) {
        static thread_local std::vector<uint32_t> work;
        static thread_local std::vector<uint64_t> bits;
        choice_sorted_u32(rng, (uint32_t)n_coord, (uint32_t)k1, idx, work, bits);
        if (!dense_all) choice_sorted_u32(rng, (uint32_t)n_coord, (uint32_t)k2, idx_dense, work, bits);
    } else {
        choice_sorted<int64_t>(rng, n_coord, k1, idx);
        if (!dense_all) choice_sorted<int64_t>(rng, n_coord, k2, idx_dense);
    }
    return OBIA_B200_OK;
}
