// radix_sort.cu — K2: onesweep-style LSD radix sort of (u64 key, u32 value) pairs.
//
// Replaces the reference's HOST thrust::sort_by_key(mortons, triangles)
// (reference load_obj.h:107). Stable, so equal keys keep face (ID) order.
//
// Structure (Adinets & Merrill, "Onesweep", 2022 — restated, not library code):
//   1. rs_histogram  one read of the keys builds the digit histograms of ALL passes
//                    (shared-memory atomics, then one global atomic per bin);
//   2. rs_scan       exclusive scan of each pass's 256 bins -> global digit bases;
//   3. rs_pass x P   one kernel per digit. Each CTA (512 threads x 8 items) takes a tile of 4096 items
//                    (ticket from an atomic counter, so look-back never waits on a
//                    tile that has not started), builds the tile's digit histogram and
//                    publishes it, resolves its global offsets by DECOUPLED LOOK-BACK over
//                    the previous tiles' status words BEFORE the ranking (so the chain of
//                    inclusive prefixes moves at histogram speed), ranks its keys with a
//                    warp-level multi-split (one ballot per digit bit + per-warp shared
//                    histograms; __match_any_sync issues far too slowly on B200), stages
//                    the tile in shared memory in digit order and writes it out in
//                    coalesced runs. Keys and values are read once and written once per pass.
//   4. rs_fixup      hybrid sort: only the top digits go through passes, the low bits are
//                    ordered per run of equal high bits (see below); conditional fallback passes.
//   rs_pass<.., SPLIT> is also the multi-GPU range partition: digit = number of splitters <= key
//                    (binary search in shared memory), destination = the owning rank's buffers
//                    through peer memory.
//
// HBM traffic per pass: 12 B read + 12 B written per item (8+8 keys-only); the
// histogram adds one 8 B read. Look-back words: 1 KiB per tile per pass.
#include <algorithm>
#include <cstdlib>

#include <cooperative_groups.h>

#include "common.cuh"

namespace b200cd {

#ifdef RS_PROFILE  // scratch builds only: cycles per phase of rs_pass, summed over tiles (thread 0)
__device__ unsigned long long g_rs_prof[16];
extern "C" __attribute__((visibility("default"))) void b200cd_debug_rs_prof(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, g_rs_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_rs_prof, z, sizeof z); }
}
#define RS_MARK(i) do { if (threadIdx.x == 0) { long long t__ = clock64(); atomicAdd(&g_rs_prof[i], (unsigned long long)(t__ - t_prev)); t_prev = t__; } } while (0)
#else
#define RS_MARK(i) do { } while (0)
#endif

namespace {

// tile shape (overridable for tuning builds: -DRS_THREADS_V=256 -DRS_IPT_V=8 -DRS_MINB_V=4)
#ifndef RS_THREADS_V
#define RS_THREADS_V 512
#endif
#ifndef RS_IPT_V
#define RS_IPT_V 8
#endif
#ifndef RS_MINB_V
#define RS_MINB_V 2
#endif
#ifndef RS_BULK
#define RS_BULK 0  // 1: tile loads by cp.async.bulk + mbarrier (tuning build, see rs_pass_body)
#endif
constexpr int RS_THREADS = RS_THREADS_V;      // >= 256: one thread per digit publishes / looks back
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_IPT = RS_IPT_V;              // items per thread
constexpr int RS_TILE = RS_THREADS * RS_IPT;  // 4096 items per tile
constexpr int RS_BITS = 8;                    // digit width (9 bits = 7 passes was measured slower: 1.51 vs 1.29 ms at 16 M)
constexpr int RS_RADIX = 1 << RS_BITS;
constexpr uint32_t ST_PARTIAL = 1u << 30;     // status word = flag | count  (count < 2^30)
constexpr uint32_t ST_INCLUSIVE = 2u << 30;
constexpr uint32_t ST_MASK = (1u << 30) - 1;
constexpr int RS_MAX_PASS = 8;
constexpr int RS_LOOKBACK = 8;                // predecessor status words fetched per look-back round

struct PassList {
    int npass;
    int shift[RS_MAX_PASS];
    uint32_t mask[RS_MAX_PASS];
};

// Digit of a key: a bit field (radix passes) or, for the multi-GPU range partition, the number of
// splitters <= key (splitters ascending, at most RS_MAX_SPLIT of them, read from device memory
// because they are computed on the device).
constexpr int RS_MAX_SPLIT = 15;  // == RS_MAX_SPLIT_P1 - 1 (common.cuh)
template <bool SPLIT>
struct Digit;
template <>
struct Digit<false> {  // bit field
    int shift;
    uint32_t mask;
    __device__ __forceinline__ uint32_t operator()(uint64_t k) const { return (uint32_t)((k >> shift) & mask); }
};
template <>
struct Digit<true> {   // range partition: number of splitters <= key
    // 16 ascending u64 in SHARED memory: the nsplit splitters, padded with ~0 (no key reaches it: keys use 63 bits).
    // Branch-free binary search, 4 shared loads + compares per key (a linear scan over 15 splitters held in registers
    // cost 45+ instructions per call, three calls per key: a third of the fused partition + exchange kernel's issue slots)
    const uint64_t* sp;
    __device__ __forceinline__ uint32_t operator()(uint64_t k) const {
        uint32_t d = (sp[7] <= k) ? 8u : 0u;
        d += (sp[d + 3] <= k) ? 4u : 0u;
        d += (sp[d + 1] <= k) ? 2u : 0u;
        d += (sp[d] <= k) ? 1u : 0u;
        return d;
    }
};
__device__ __forceinline__ void load_splitters(uint64_t* s_split, const uint64_t* __restrict__ splitters, int nsplit) {
    if (threadIdx.x < 16) s_split[threadIdx.x] = ((int)threadIdx.x < nsplit) ? __ldg(splitters + threadIdx.x) : ~0ull;
    __syncthreads();
}
template <bool SPLIT>
__device__ __forceinline__ Digit<SPLIT> make_digit(int shift, uint32_t mask, const uint64_t* s_split);
template <>
__device__ __forceinline__ Digit<false> make_digit<false>(int shift, uint32_t mask, const uint64_t*) {
    Digit<false> dg;
    dg.shift = shift; dg.mask = mask;
    return dg;
}
template <>
__device__ __forceinline__ Digit<true> make_digit<true>(int, uint32_t, const uint64_t* s_split) {
    Digit<true> dg;
    dg.sp = s_split;
    return dg;
}

// ---- 1. histograms of every pass in one sweep
constexpr int RH_THREADS = 256;
// histogram of ONE custom digit (range partition by splitters)
__global__ void __launch_bounds__(RH_THREADS) rs_histogram_split(const uint64_t* __restrict__ keys, uint32_t n,
                                                                 const uint64_t* __restrict__ splitters, int nsplit,
                                                                 uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[RS_MAX_SPLIT + 1];
    __shared__ uint64_t s_split[16];
    if (threadIdx.x <= RS_MAX_SPLIT) sh[threadIdx.x] = 0;
    load_splitters(s_split, splitters, nsplit);
    const Digit<true> dg = make_digit<true>(0, 0, s_split);
    for (uint32_t i = blockIdx.x * RH_THREADS + threadIdx.x; i < n; i += gridDim.x * RH_THREADS) {
        const uint32_t d = dg(__ldg(keys + i));
        const uint32_t peers = __match_any_sync(__activemask(), d);
        if ((threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&sh[d], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x <= RS_MAX_SPLIT && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// cond (optional): device word; the kernel does nothing when it is 0 (fallback path of the hybrid sort, below).
// clear / clear_words: status words to zero first (grid-stride) - only the conditional launch uses it.
__global__ void __launch_bounds__(RH_THREADS) rs_histogram(const uint64_t* __restrict__ keys, uint32_t n,
                                                           PassList pl, uint32_t* __restrict__ hist,
                                                           const uint32_t* __restrict__ cond, uint4* __restrict__ clear,
                                                           uint64_t clear_vec) {
    __shared__ uint32_t sh[RS_MAX_PASS * RS_RADIX];
    if (cond && *cond == 0) return;
    if (clear)
        for (uint64_t i = (uint64_t)blockIdx.x * RH_THREADS + threadIdx.x; i < clear_vec; i += (uint64_t)gridDim.x * RH_THREADS)
            clear[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < pl.npass * RS_RADIX; i += RH_THREADS) sh[i] = 0;
    __syncthreads();
    // grid-stride over 2-key vectors (16-B loads)
    const uint32_t nvec = n >> 1;
    const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(keys);
    for (uint32_t i = blockIdx.x * RH_THREADS + threadIdx.x; i < nvec; i += gridDim.x * RH_THREADS) {
        ulonglong2 k = __ldg(k2 + i);
#pragma unroll
        for (int p = 0; p < RS_MAX_PASS; ++p) {
            if (p < pl.npass) {
                atomicAdd(&sh[p * RS_RADIX + ((k.x >> pl.shift[p]) & pl.mask[p])], 1u);
                atomicAdd(&sh[p * RS_RADIX + ((k.y >> pl.shift[p]) & pl.mask[p])], 1u);
            }
        }
    }
    if ((n & 1u) && blockIdx.x == 0 && threadIdx.x == 0) {
        uint64_t k = keys[n - 1];
        for (int p = 0; p < pl.npass; ++p) atomicAdd(&sh[p * RS_RADIX + ((k >> pl.shift[p]) & pl.mask[p])], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < pl.npass * RS_RADIX; i += RH_THREADS) {
        uint32_t c = sh[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ---- 2. exclusive scan of 256 bins, one block per pass
__global__ void __launch_bounds__(RS_RADIX) rs_scan(uint32_t* __restrict__ hist, const uint32_t* __restrict__ cond) {
    __shared__ uint32_t wsum[RS_RADIX / 32];
    if (cond && *cond == 0) return;
    uint32_t* h = hist + blockIdx.x * RS_RADIX;
    uint32_t v = h[threadIdx.x];
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) base += wsum[w];
    h[threadIdx.x] = base + incl - v;
}

// ---- 3. one digit pass
__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 512 threads x 8 items, two CTAs per SM (<= 64 registers): 32 resident warps hide the
// load -> rank -> look-back -> scatter latency chain of each tile behind the other tile's.
// COND: persistent variant for the hybrid sort's fallback - a fixed grid that does nothing when *cond == 0 and
// otherwise works through all the tiles (ticket loop). The plain variant runs one tile per CTA.
template <bool HAS_VALUES, bool SPLIT, bool COND>
__device__ __forceinline__ void
rs_pass_body(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out, const uint32_t* __restrict__ vals_in,
             uint32_t* __restrict__ vals_out, uint32_t n, int shift, uint32_t mask, int iota_values, uint32_t iota_base,
             const uint64_t* __restrict__ splitters, int nsplit,
             const PeerTable* __restrict__ peers,        // non-null: bucket d is written into peer d's buffers (NVLink stores)
             const uint32_t* __restrict__ digit_base_g,  // [256] exclusive global digit offsets of this pass
             uint32_t* __restrict__ status,              // [ntiles][256] look-back words of this pass
             uint32_t* __restrict__ ticket, uint32_t ntiles) {
    // dynamic shared memory (> 48 KiB): [stage_k 32 KiB][stage_v 16 KiB if values][warp_hist][digit_base][warp_tot][tile][tile_hist]
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t* stage_k = reinterpret_cast<uint64_t*>(rs_smem);
    uint32_t* stage_v = reinterpret_cast<uint32_t*>(rs_smem + RS_TILE * 8);
    uint32_t (*warp_hist)[RS_RADIX + 1] = reinterpret_cast<uint32_t (*)[RS_RADIX + 1]>(
        rs_smem + RS_TILE * 8 + (HAS_VALUES ? RS_TILE * 4 : 0));  // +1: bin 256 collects out-of-range padding
    uint32_t* digit_base = &warp_hist[RS_WARPS][0];
    uint32_t* warp_tot = digit_base + RS_RADIX;
    uint32_t& s_tile = warp_tot[RS_RADIX / 32];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef RS_PROFILE
    long long t_prev = clock64();
#endif
    __shared__ uint64_t s_split[16];
    if (SPLIT) load_splitters(s_split, splitters, nsplit);
    const Digit<SPLIT> digit_of = make_digit<SPLIT>(shift, mask, s_split);
  for (;;) {  // one tile per CTA, or (COND) a ticket loop over all tiles
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < RS_WARPS * (RS_RADIX + 1); i += RS_THREADS) (&warp_hist[0][0])[i] = 0;
    if (tid < RS_RADIX) (warp_tot + RS_RADIX / 32 + 1)[tid] = 0;  // tile_hist
    __syncthreads();
    const uint32_t tile = s_tile;
    if (COND && tile >= ntiles) break;
    const uint32_t tile_base = tile * RS_TILE;
    const uint32_t tile_items = min((uint32_t)RS_TILE, n - tile_base);
    const uint32_t my_base = tile_base + warp * (32 * RS_IPT) + lane;  // warp-striped: item k at my_base + 32k

    // load keys (and values: their latency overlaps the ranking)
    uint64_t key[RS_IPT];
    uint32_t val[RS_IPT];
#if RS_BULK
    // Tuning build (-DRS_BULK=1): the tile's keys (and values) arrive by ONE bulk asynchronous copy each
    // (cp.async.bulk + mbarrier, the 1-D TMA path) into the staging buffers, and the threads pick their items up from
    // shared memory. Measured (profiles/r02_sort_bulk.md): no faster than the coalesced loads below - the pass is bound by
    // its look-back chain and by 2 CTAs of 64 registers per SM, not by load issue.
    const bool full_tile = tile_items == (uint32_t)RS_TILE && !SPLIT;
    if (full_tile) {
        __shared__ __align__(8) unsigned long long s_mbar;
        const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
        const uint32_t bytes = RS_TILE * 8u + ((HAS_VALUES && !iota_values) ? RS_TILE * 4u : 0u);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(stage_k)),
                         "l"(keys_in + tile_base), "r"((uint32_t)(RS_TILE * 8u)), "r"(mbar)
                         : "memory");
            if (HAS_VALUES && !iota_values)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 (uint32_t)__cvta_generic_to_shared(stage_v)),
                             "l"(vals_in + tile_base), "r"((uint32_t)(RS_TILE * 4u)), "r"(mbar)
                             : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(mbar) : "memory");
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k) {
            const uint32_t l = warp * (32 * RS_IPT) + lane + 32 * k;
            key[k] = stage_k[l];
            if (HAS_VALUES) val[k] = iota_values ? iota_base + tile_base + l : stage_v[l];
        }
        __syncthreads();  // everybody holds its items: the staging buffers are free again (and s_mbar may be re-initialised)
        if (tid == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mbar) : "memory");
    } else
#endif
    {
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k) {
            const uint32_t g = my_base + 32 * k;
            key[k] = (g < n) ? __ldg(keys_in + g) : ~0ull;
        }
        if (HAS_VALUES) {
#pragma unroll
            for (int k = 0; k < RS_IPT; ++k) {
                const uint32_t g = my_base + 32 * k;
                val[k] = iota_values ? iota_base + g : ((g < n) ? __ldg(vals_in + g) : 0u);
            }
        }
    }

    RS_MARK(0);  // ticket + zero + issue loads
    // ---- early counts (onesweep): the tile's digit histogram first, so that its PARTIAL status is out and its
    // look-back is done BEFORE the (long) ranking. Tiles start in ticket order; with the inclusive prefixes
    // published this early a tile rarely has to look further back than a few predecessors. (Ranking first
    // - the previous version - delayed every tile's inclusive word by the ranking time and the look-back
    // chains grew to the number of co-resident tiles: 15 us per tile, measured.)
    uint32_t* tile_hist = warp_tot + RS_RADIX / 32 + 1;  // [RS_RADIX]
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        const uint32_t g = my_base + 32 * k;
        if (g < n) atomicAdd(&tile_hist[digit_of(key[k])], 1u);
    }
    __syncthreads();
    RS_MARK(1);  // keys arrive + tile histogram
    uint32_t sum = 0, excl = 0, incl = 0;
    if (tid < RS_RADIX) {
        const uint32_t d = tid;
        sum = tile_hist[d];
        // publish this tile's count for digit d, then sum the counts of earlier tiles
        uint32_t* my_status = status + (size_t)tile * RS_RADIX + d;
        st_volatile(my_status, (tile == 0 ? ST_INCLUSIVE : ST_PARTIAL) | sum);
        if (tile > 0) {
            int t = (int)tile - 1;
            bool done = false;
            while (!done) {
#ifdef RS_PROFILE
                if (tid == 0) atomicAdd(&g_rs_prof[8], 1ull);  // look-back rounds
#endif
                // fetch several predecessors at once: their latencies overlap; consume in order
                uint32_t sv[RS_LOOKBACK];
#pragma unroll
                for (int j = 0; j < RS_LOOKBACK; ++j)
                    sv[j] = (t - j >= 0) ? ld_volatile(status + (size_t)(t - j) * RS_RADIX + d) : ST_INCLUSIVE;
#pragma unroll
                for (int j = 0; j < RS_LOOKBACK; ++j) {
                    if (done) break;
                    if ((sv[j] & ~ST_MASK) == 0) {  // not published yet (tile is running: tickets are ordered) - refetch from here
#ifdef RS_PROFILE
                        if (tid == 0) atomicAdd(&g_rs_prof[10], 1ull);  // rounds cut short by an unpublished predecessor
#endif
                        break;
                    }
#ifdef RS_PROFILE
                    if (tid == 0) atomicAdd(&g_rs_prof[9], 1ull);  // predecessors consumed
#endif
                    excl += sv[j] & ST_MASK;
                    --t;
                    if (sv[j] & ST_INCLUSIVE) done = true;
                }
            }
            st_volatile(my_status, ST_INCLUSIVE | (excl + sum));
        }
        // exclusive scan of `sum` over the 256 digits -> where digit d starts inside the tile
        incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t2 = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t2;
        }
        if (lane == 31) warp_tot[warp] = incl;
    }

    RS_MARK(2);  // publish + look-back (thread 0's digit)
    // ---- warp-level multi-split: rank of every key among the warp's keys with the same digit
    uint32_t rank[RS_IPT];
    uint32_t* wh = warp_hist[warp];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        const uint32_t g = my_base + 32 * k;
        const uint32_t d = (g < n) ? digit_of(key[k]) : (uint32_t)RS_RADIX;
        // lanes holding the same digit, from RS_BITS + 1 ballots (digit bits + "in range"). MATCH.ANY gives the same
        // mask in one instruction but issues only about once per 130 cycles per SM sub-partition on B200
        // (measured: 8.5 k of a tile's 30 k cycles went into 8 rounds of it).
        uint32_t peers = 0xffffffffu;
        // (range partition: at most 16 buckets -> 4 digit bits + the "in range" bit 8 are enough)
        constexpr int NBALLOT = SPLIT ? 5 : RS_BITS + 1;
#pragma unroll
        for (int bb = 0; bb < NBALLOT; ++bb) {
            const int b = (SPLIT && bb == 4) ? RS_BITS : bb;
            const bool bit = (d >> b) & 1u;
            const uint32_t vote = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? vote : ~vote;
        }
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if ((int)lane == leader) {
            prev = wh[d];
            wh[d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[k] = prev + __popc(peers & lt);
        __syncwarp();
    }
    RS_MARK(3);  // ranking (warp 0)
    __syncthreads();
    RS_MARK(4);  // wait for the other warps
    // per digit: where the digit starts inside the tile, exclusive prefix over the warps
    if (tid < RS_RADIX) {
        const uint32_t d = tid;
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < RS_RADIX / 32; ++w) wbase += (w < (int)warp) ? warp_tot[w] : 0u;
        const uint32_t local_start = wbase + incl - sum;
        // final position of the item with in-tile slot r and digit d: digit_base[d] + r  (mod 2^32)
        digit_base[d] = __ldg(digit_base_g + d) + excl - local_start;
        uint32_t run = local_start;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = warp_hist[w][d];
            warp_hist[w][d] = run;
            run += c;
        }
    }
    __syncthreads();

    RS_MARK(5);  // digit bases
    // stage keys (and values) in digit order
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        const uint32_t g = my_base + 32 * k;
        if (g < n) {
            const uint32_t d = digit_of(key[k]);
            const uint32_t slot = wh[d] + rank[k];
            stage_k[slot] = key[k];
            if (HAS_VALUES) stage_v[slot] = val[k];
        }
    }
    __syncthreads();
    RS_MARK(6);  // staging
    // write out: consecutive slots of one digit land on consecutive addresses
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i) {
        const uint32_t r = i * RS_THREADS + tid;
        if (r < tile_items) {
            const uint64_t kk = stage_k[r];
            const uint32_t pos = digit_base[digit_of(kk)] + r;
            if (SPLIT && peers) {  // pos is relative to the start of my segment in the destination rank's buffer
                const uint32_t d = digit_of(kk);
                peers->keys[d][pos] = kk;
                if (HAS_VALUES) peers->ids[d][pos] = stage_v[r];
            } else {
                keys_out[pos] = kk;
                if (HAS_VALUES) vals_out[pos] = stage_v[r];
            }
        }
    }
    RS_MARK(7);  // write-out issue
    if (!COND) break;
    __syncthreads();
  }
}

template <bool HAS_VALUES, bool SPLIT>
__global__ void __launch_bounds__(RS_THREADS, RS_MINB_V)
rs_pass(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out, const uint32_t* __restrict__ vals_in,
        uint32_t* __restrict__ vals_out, uint32_t n, int shift, uint32_t mask, int iota_values, uint32_t iota_base,
        const uint64_t* __restrict__ splitters, int nsplit, const PeerTable* __restrict__ peers,
        const uint32_t* __restrict__ digit_base_g, uint32_t* __restrict__ status, uint32_t* __restrict__ ticket) {
    rs_pass_body<HAS_VALUES, SPLIT, false>(keys_in, keys_out, vals_in, vals_out, n, shift, mask, iota_values, iota_base, splitters,
                                           nsplit, peers, digit_base_g, status, ticket, 0u);
}

// one conditional pass per launch: only used where the cooperative launch below is refused (e.g. a partitioned GPU)
__global__ void __launch_bounds__(RS_THREADS, RS_MINB_V)
rs_pass_cond(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out, const uint32_t* __restrict__ vals_in,
             uint32_t* __restrict__ vals_out, uint32_t n, int shift, uint32_t mask, const uint32_t* __restrict__ digit_base_g,
             uint32_t* __restrict__ status, uint32_t* __restrict__ ticket, const uint32_t* __restrict__ cond, uint32_t ntiles) {
    if (*cond == 0) return;
    rs_pass_body<true, false, true>(keys_in, keys_out, vals_in, vals_out, n, shift, mask, 0, 0u, nullptr, 0, nullptr, digit_base_g,
                                    status, ticket, ntiles);
}

// The hybrid sort's fallback: ALL passes of a full sort in ONE cooperative launch (a fixed grid of co-resident CTAs
// works through every pass's tiles by ticket, grid-wide barrier between passes). Launched behind every fix-up and
// idle - one launch that returns at once - unless the fix-up met a run it does not handle (*cond != 0). Before: eight
// conditional launches per build, each idle but not free (~25 us per build; 4 % of a 1.26 M-triangle step).
__global__ void __launch_bounds__(RS_THREADS, RS_MINB_V)
rs_fallback_kernel(uint64_t* __restrict__ k0, uint64_t* __restrict__ k1, uint32_t* __restrict__ v0, uint32_t* __restrict__ v1, uint32_t n,
                   PassList all, const uint32_t* __restrict__ digit_bases /* [npass][256] */, uint32_t* __restrict__ status_all,
                   uint32_t* __restrict__ tickets, const uint32_t* __restrict__ cond, uint32_t ntiles) {
    if (*cond == 0) return;  // (every CTA takes the same way: nobody is left waiting at a barrier)
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    for (int p = 0; p < all.npass; ++p) {
        const bool even = (p & 1) == 0;
        rs_pass_body<true, false, true>(even ? k0 : k1, even ? k1 : k0, even ? v0 : v1, even ? v1 : v0, n, all.shift[p], all.mask[p], 0,
                                        0u, nullptr, 0, nullptr, digit_bases + p * RS_RADIX, status_all + (size_t)p * ntiles * RS_RADIX,
                                        tickets + p, ntiles);
        __threadfence();
        grid.sync();
    }
}

// ---- 4. hybrid sort: fix-up of the low bits after sorting only the high digits
// After stable passes over the digits at and above `lowbit`, the items are ordered by their high bits and every
// run of equal high bits still has to be ordered by its low bits (ties keep their order). With lowbit chosen so
// that the high bits alone separate nearly all keys (Morton keys of N triangles: ~log2 N + a few bits), the runs
// are 1-3 items long and ONE coalesced read of the keys replaces lowbit/8 full passes. One thread per run head:
// stable insertion sort of the run in place. A run longer than RS_MAXRUN raises fix[0]; the conditional passes
// launched behind this kernel then redo the sort over every digit (correct for any input, e.g. all keys equal).
// fix[1] = longest run seen, fix[2] = items in runs of 2 or more (the host adapts lowbit for the next build).
constexpr int RS_MAXRUN = 48;
constexpr int RF_THREADS = 256;
constexpr int RF_IPT = 8;  // items per thread: one pair of (same-address) statistics atomics per 2048 items
__global__ void __launch_bounds__(RF_THREADS)
rs_fixup(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t n, int lowbit, int top, uint32_t* __restrict__ fix) {
    __shared__ uint32_t s_max, s_sum, s_bits;
    if (threadIdx.x == 0) { s_max = 0; s_sum = 0; s_bits = 0; }
    __syncthreads();
    uint32_t tmax = 0, tsum = 0;
    uint64_t seen = 0;  // OR of this thread's keys: fix[3] = significant key bits (where the next build puts its digit window)
    // all the block's loads first (two coalesced reads per item), then the rare per-run work
    const uint32_t base = blockIdx.x * (RF_IPT * RF_THREADS) + threadIdx.x;
    uint64_t kk[RF_IPT], kp[RF_IPT];
#pragma unroll
    for (int r = 0; r < RF_IPT; ++r) {
        const uint32_t i = base + r * RF_THREADS;
        kk[r] = i < n ? keys[i] : 0ull;
        kp[r] = (i < n && i > 0) ? keys[i - 1] : 0ull;
        seen |= kk[r];
    }
    // a key with bits at or above `top` was not ordered by the passes (their window ends there): redo everything
    if (top < 64 && (seen >> top)) atomicOr(fix + 0, 2u);
    // (a run's head may already be reordering it while others take their snapshot: harmless, the HIGH bits - all that
    // the head test looks at - are the same for every item of a run, and 8-byte accesses do not tear)
#pragma unroll
    for (int r = 0; r < RF_IPT; ++r) {
        const uint32_t i = base + r * RF_THREADS;
        if (i >= n) continue;
        const uint64_t hi = kk[r] >> lowbit;
        const bool head = i == 0 || (kp[r] >> lowbit) != hi;
        if (!head) continue;
        uint32_t j = i + 1;
        while (j < n && j - i <= (uint32_t)RS_MAXRUN && (keys[j] >> lowbit) == hi) ++j;
        const uint32_t len = j - i;
        if (len > (uint32_t)RS_MAXRUN) {
            atomicOr(fix + 0, 1u);  // bit 0: a run was too long; bit 1: a key reached above the digit window
        } else if (len > 1) {
            for (uint32_t a = i + 1; a < j; ++a) {  // stable: an item only moves past strictly larger keys
                const uint64_t ka = keys[a];
                const uint32_t va = vals ? vals[a] : 0u;
                uint32_t b = a;
                while (b > i) {
                    const uint64_t kb = keys[b - 1];
                    if (kb <= ka) break;
                    keys[b] = kb;
                    if (vals) vals[b] = vals[b - 1];
                    --b;
                }
                if (b != a) {
                    keys[b] = ka;
                    if (vals) vals[b] = va;
                }
            }
            tmax = max(tmax, len);
            tsum += len;
        }
    }
    // statistics: shared-memory reduction, then one pair of global atomics per block that saw a run of 2 or more
    if (tmax) { atomicMax(&s_max, tmax); atomicAdd(&s_sum, tsum); }
    const uint32_t wbits = __reduce_max_sync(0xffffffffu, seen ? 64u - (uint32_t)__clzll((long long)seen) : 0u);
    if ((threadIdx.x & 31) == 0 && wbits) atomicMax(&s_bits, wbits);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_max) {
            atomicMax(fix + 1, s_max);
            atomicAdd(fix + 2, s_sum);
        }
        if (s_bits) atomicMax(fix + 3, s_bits);
    }
}

constexpr size_t rs_smem_bytes(bool has_values) {
    return (size_t)RS_TILE * 8 + (has_values ? (size_t)RS_TILE * 4 : 0) + sizeof(uint32_t) * (RS_WARPS * (RS_RADIX + 1) + RS_RADIX + RS_RADIX / 32 + 1 + RS_RADIX);
}

}  // namespace

// > 48 KiB of dynamic shared memory must be opted into, once per function and device (NOT once per launch: the
// host calls sit in latency-bound steps)
static void opt_in_shared_memory() {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && attr_set[dev]) return;
    cudaFuncSetAttribute(rs_pass<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(true));
    cudaFuncSetAttribute(rs_pass<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(false));
    cudaFuncSetAttribute(rs_fallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(true));
    cudaFuncSetAttribute(rs_pass_cond, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(true));
    cudaFuncSetAttribute(rs_pass<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(true));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
}

// ---- 5. small lists: ONE block sorts up to SS_MAX packed pairs in shared memory (bitonic network) instead of a
// histogram + scan + 2 x ceil(id_bits / 8) radix launches whose cost is all launch latency (a 384-pair result - the
// flag mesh - took 10 launches, ~60 us, for 3 KB of data). Order: lexicographic by (first ID, second ID), the same as
// the LSD passes over the two halves of the packed word give.
namespace {
constexpr int SS_THREADS = 1024;
constexpr int SS_MAX = 8192;  // 64 KiB of shared memory
__global__ void __launch_bounds__(SS_THREADS)
small_pair_sort_kernel(uint2* __restrict__ pairs, uint32_t count, uint32_t padded /* power of two >= count */) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    uint64_t* key = reinterpret_cast<uint64_t*>(ss_smem);
    for (uint32_t i = threadIdx.x; i < padded; i += SS_THREADS) {
        uint64_t k = ~0ull;  // padding sorts last (a real pair has first ID < second ID, never both all-ones)
        if (i < count) {
            const uint2 p = pairs[i];
            k = ((uint64_t)p.x << 32) | p.y;
        }
        key[i] = k;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= padded; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < padded / 2; t += SS_THREADS) {
                const uint32_t lo = 2 * t - (t & (stride - 1));  // partner pairs (lo, lo + stride)
                const uint32_t hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint64_t a = key[lo], b = key[hi];
                if ((a > b) == up) { key[lo] = b; key[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < count; i += SS_THREADS) {
        const uint64_t k = key[i];
        pairs[i] = make_uint2((uint32_t)(k >> 32), (uint32_t)k);
    }
}
}  // namespace

bool small_pair_sort(uint2* d_pairs, uint64_t count, cudaStream_t s) {
    if (count < 2) return true;
    if (count > (uint64_t)SS_MAX) return false;
    uint32_t padded = 2;
    while (padded < count) padded <<= 1;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaFuncSetAttribute(small_pair_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SS_MAX * 8);
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    small_pair_sort_kernel<<<1, SS_THREADS, (size_t)padded * 8, s>>>(d_pairs, (uint32_t)count, padded);
    count_launch();
    trace_mark("small_pair_sort", s);
    return true;
}

int radix_digit_bits() { return RS_BITS; }

uint32_t radix_hist_words(int npass) { return (uint32_t)(npass * RS_RADIX + RS_MAX_PASS); }  // + tickets

uint64_t radix_tile_status_words(uint32_t n, int npass) {
    uint64_t tiles = ((uint64_t)n + RS_TILE - 1) / RS_TILE;
    return tiles * RS_RADIX * (uint64_t)npass;
}

// which passes run first (unconditionally), and where the hybrid sort's window of digits ends
static void make_plan(const RadixPass* passes, int npass, bool has_values, int high_passes, bool has_fix, int top_bits,
                      PassList& pl, PassList& all, bool& hybrid, int& top) {
    hybrid = high_passes > 0 && high_passes < npass && has_fix && has_values && npass % 2 == 0;
    const int first = hybrid ? npass - high_passes : 0;  // passes [first, npass) run unconditionally
    pl = PassList{};
    all = PassList{};
    all.npass = npass;
    for (int p = 0; p < npass; ++p) {
        all.shift[p] = passes[p].shift;
        all.mask[p] = (1u << passes[p].bits) - 1u;
    }
    pl.npass = npass - first;
    for (int p = first; p < npass; ++p) {
        pl.shift[p - first] = all.shift[p];
        pl.mask[p - first] = all.mask[p];
    }
    // hybrid: the window of sorted digits ends at the keys' highest significant bit (as the previous build saw it; in-box
    // Morton keys use 60 of the 63 bits), so all of its 8 * high_passes bits separate keys. The fix-up checks that no key
    // reaches above the window and otherwise hands over to the fallback passes.
    const int key_top = all.shift[npass - 1] + (32 - __builtin_clz(all.mask[npass - 1]));  // bits the full plan covers
    top = hybrid ? std::max(8 * high_passes, std::min(top_bits > 0 ? top_bits : key_top, key_top)) : key_top;
    if (hybrid) {
        for (int p = 0; p < high_passes; ++p) {
            pl.shift[p] = top - 8 * (high_passes - p);
            pl.mask[p] = 0xffu;
        }
    }
}

void radix_hist_plan(const RadixPass* passes, int npass, bool has_values, int high_passes, bool has_fix, int top_bits,
                     RadixHistPlan* out) {
    PassList pl, all;
    bool hybrid;
    int top;
    make_plan(passes, npass, has_values, high_passes, has_fix, top_bits, pl, all, hybrid, top);
    out->npass = pl.npass;
    for (int p = 0; p < RS_MAX_PASS; ++p) {
        out->shift[p] = pl.shift[p];
        out->mask[p] = pl.mask[p];
    }
}

int radix_sort(uint64_t* keys[2], uint32_t* vals[2], uint32_t n, const RadixPass* passes, int npass,
               bool iota_values, uint32_t* d_hist, uint32_t* d_tile_status, uint64_t tile_status_words, int sms,
               cudaStream_t s, int high_passes, uint32_t* d_fix, int top_bits, cudaEvent_t* ev4, bool hist_done) {
    (void)tile_status_words;
    if (n == 0 || npass == 0) return 0;
    PassList pl, all;
    bool hybrid;
    int top;
    make_plan(passes, npass, vals != nullptr, high_passes, d_fix != nullptr, top_bits, pl, all, hybrid, top);
    const uint32_t tiles = (n + RS_TILE - 1) / RS_TILE;
    opt_in_shared_memory();
    const uint32_t hblocks = min((n / 2 + RH_THREADS - 1) / RH_THREADS + 1, (uint32_t)sms * 8u);
    uint32_t* d_ticket = d_hist + npass * RS_RADIX;
    if (!hist_done) cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * radix_hist_words(npass), s);
    cudaMemsetAsync(d_tile_status, 0, sizeof(uint32_t) * (size_t)tiles * RS_RADIX * pl.npass, s);
    if (hybrid) cudaMemsetAsync(d_fix, 0, sizeof(uint32_t) * 4, s);
    if (!hist_done) {  // (otherwise K1 counted the digits while it had the keys in registers)
        rs_histogram<<<hblocks, RH_THREADS, 0, s>>>(keys[0], n, pl, d_hist, nullptr, nullptr, 0);
        count_launch();
    }
    rs_scan<<<pl.npass, RS_RADIX, 0, s>>>(d_hist, nullptr);
    count_launch();
    trace_mark("rs_histogram+scan", s);
    int cur = 0;
    for (int p = 0; p < pl.npass; ++p) {
        uint32_t* status = d_tile_status + (size_t)p * tiles * RS_RADIX;
        const int iota = (iota_values && p == 0) ? 1 : 0;
        if (hybrid && ev4 && p == pl.npass - 1) cudaEventRecord(ev4[0], s);
        if (vals)
            rs_pass<true, false><<<tiles, RS_THREADS, rs_smem_bytes(true), s>>>(keys[cur], keys[cur ^ 1], vals[cur], vals[cur ^ 1], n,
                                                       pl.shift[p], pl.mask[p], iota, 0u, nullptr, 0, nullptr,
                                                       d_hist + p * RS_RADIX, status, d_ticket + p);
        else
            rs_pass<false, false><<<tiles, RS_THREADS, rs_smem_bytes(false), s>>>(keys[cur], keys[cur ^ 1], nullptr, nullptr, n, pl.shift[p],
                                                        pl.mask[p], 0, 0u, nullptr, 0, nullptr, d_hist + p * RS_RADIX, status, d_ticket + p);
        count_launch();
        trace_mark("rs_pass", s);
        if (hybrid && ev4 && p == pl.npass - 1) cudaEventRecord(ev4[1], s);
        cur ^= 1;
    }
    if (!hybrid) return cur;
    // low bits: per-run fix-up; if a run was too long, the conditional kernels below sort the (already permuted,
    // ties still in their original order) items again over every digit. They exit at once when fix[0] == 0.
    if (ev4) cudaEventRecord(ev4[2], s);
    rs_fixup<<<(n + RF_THREADS * RF_IPT - 1) / (RF_THREADS * RF_IPT), RF_THREADS, 0, s>>>(keys[cur], vals[cur], n, pl.shift[0], top, d_fix);
    count_launch();
    if (ev4) cudaEventRecord(ev4[3], s);
    trace_mark("rs_fixup", s);
    cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * radix_hist_words(npass), s);  // small; the status words are cleared conditionally
    rs_histogram<<<hblocks, RH_THREADS, 0, s>>>(keys[cur], n, all, d_hist, d_fix, reinterpret_cast<uint4*>(d_tile_status),
                                                 (uint64_t)tiles * RS_RADIX * npass / 4);
    rs_scan<<<npass, RS_RADIX, 0, s>>>(d_hist, d_fix);
    count_launch(2);
    {   // an even number of passes: the result lands in the buffer the fix-up worked on
        static int resident[64] = {};  // co-resident CTAs of the cooperative launch, per device
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) dev = 0;
        if (!resident[dev]) {
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rs_fallback_kernel, RS_THREADS, rs_smem_bytes(true));
            resident[dev] = std::max(1, std::min(per_sm, 2)) * std::max(sms, 1);
        }
        uint64_t* k0 = keys[cur];
        uint64_t* k1 = keys[cur ^ 1];
        uint32_t* v0 = vals[cur];
        uint32_t* v1 = vals[cur ^ 1];
        const uint32_t* bases = d_hist;
        uint32_t* status_all = d_tile_status;
        uint32_t* tickets = d_ticket;
        const uint32_t* cond = d_fix;
        uint32_t ntiles = tiles;
        void* args[] = {&k0, &k1, &v0, &v1, &n, &all, &bases, &status_all, &tickets, &cond, &ntiles};
        static int coop = -1;  // B200CD_COOP=0: never use the cooperative launch (tests of the path below; read once)
        if (coop < 0) {
            const char* ev = getenv("B200CD_COOP");
            coop = (ev && ev[0] == '0') ? 0 : 1;
        }
        const cudaError_t e = !coop ? cudaErrorCooperativeLaunchTooLarge
                                    : cudaLaunchCooperativeKernel(reinterpret_cast<void*>(rs_fallback_kernel), dim3(resident[dev]),
                                                                  dim3(RS_THREADS), args, rs_smem_bytes(true), s);
        count_launch();
        if (e != cudaSuccess) {  // refused: the same passes as eight conditional launches (the result must never depend on it)
            cudaGetLastError();
            int c = cur;
            for (int p = 0; p < npass; ++p) {
                rs_pass_cond<<<2 * sms, RS_THREADS, rs_smem_bytes(true), s>>>(keys[c], keys[c ^ 1], vals[c], vals[c ^ 1], n, all.shift[p],
                                                                               all.mask[p], d_hist + p * RS_RADIX,
                                                                               d_tile_status + (size_t)p * tiles * RS_RADIX, d_ticket + p,
                                                                               d_fix, tiles);
                count_launch();
                c ^= 1;
            }
        }
    }
    trace_mark("rs_fallback (idle unless a run was too long)", s);
    return cur;
}

// Stable partition of (key, value) by key range: item goes to bucket d = #splitters <= key.
// d_splitters: nsplit ascending keys in device memory. The per-bucket counts are left in
// d_hist[256 .. 256+nsplit] (d_hist[0..] holds the exclusive bucket offsets). values_in may be null:
// value = iota_base + index.
void radix_partition(const uint64_t* keys_in, const uint32_t* vals_in, uint32_t iota_base, uint64_t* keys_out,
                     uint32_t* vals_out, uint32_t n, const uint64_t* d_splitters, int nsplit, uint32_t* d_hist,
                     uint32_t* d_tile_status, int sms, cudaStream_t s) {
    if (n == 0) {
        cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * 2 * RS_RADIX, s);
        return;
    }
    opt_in_shared_memory();
    const uint32_t tiles = (n + RS_TILE - 1) / RS_TILE;
    uint32_t* d_ticket = d_hist + 2 * RS_RADIX;
    cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * (2 * RS_RADIX + 1), s);
    cudaMemsetAsync(d_tile_status, 0, sizeof(uint32_t) * (size_t)tiles * RS_RADIX, s);
    const uint32_t hblocks = min((n + RH_THREADS - 1) / RH_THREADS, (uint32_t)sms * 8u);
    rs_histogram_split<<<hblocks, RH_THREADS, 0, s>>>(keys_in, n, d_splitters, nsplit, d_hist);
    count_launch();
    cudaMemcpyAsync(d_hist + RS_RADIX, d_hist, sizeof(uint32_t) * RS_RADIX, cudaMemcpyDeviceToDevice, s);  // keep the counts
    rs_scan<<<1, RS_RADIX, 0, s>>>(d_hist, nullptr);
    count_launch();
    rs_pass<true, true><<<tiles, RS_THREADS, rs_smem_bytes(true), s>>>(keys_in, keys_out, vals_in, vals_out, n, 0, 0u,
                                                                vals_in ? 0 : 1, iota_base, d_splitters, nsplit, nullptr,
                                                                d_hist, d_tile_status, d_ticket);
    count_launch();
}

// Fused partition + exchange, step 1: how many of my keys go to each rank (u32 counts[nsplit+1] in device memory).
void radix_partition_counts(const uint64_t* keys_in, uint32_t n, const uint64_t* d_splitters, int nsplit, uint32_t* d_counts,
                            int sms, cudaStream_t s) {
    cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * (nsplit + 1), s);
    if (n == 0) return;
    const uint32_t hblocks = min((n + RH_THREADS - 1) / RH_THREADS, (uint32_t)sms * 8u);
    rs_histogram_split<<<hblocks, RH_THREADS, 0, s>>>(keys_in, n, d_splitters, nsplit, d_counts);
    count_launch();
}

// Step 2: one pass over my keys; every (key, id) is stored directly into its owner's receive buffer
// (d_peers, device-resident PeerTable) at d_recv_offsets[d] + its stable rank inside my bucket d.
// The stores to other ranks travel over NVLink while the tile is still being ranked.
void radix_partition_to_peers(const uint64_t* keys_in, uint32_t iota_base, uint32_t n, const uint64_t* d_splitters, int nsplit,
                              const PeerTable* d_peers, const uint32_t* d_recv_offsets, uint32_t* d_hist,
                              uint32_t* d_tile_status, cudaStream_t s) {
    if (n == 0) return;
    opt_in_shared_memory();
    const uint32_t tiles = (n + RS_TILE - 1) / RS_TILE;
    uint32_t* d_ticket = d_hist + 2 * RS_RADIX;
    // digit bases = where my segment starts in each destination buffer; unused digits stay 0
    cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * (2 * RS_RADIX + 1), s);
    cudaMemcpyAsync(d_hist, d_recv_offsets, sizeof(uint32_t) * (nsplit + 1), cudaMemcpyDeviceToDevice, s);
    cudaMemsetAsync(d_tile_status, 0, sizeof(uint32_t) * (size_t)tiles * RS_RADIX, s);
    rs_pass<true, true><<<tiles, RS_THREADS, rs_smem_bytes(true), s>>>(keys_in, nullptr, nullptr, nullptr, n, 0, 0u, 1, iota_base,
                                                                d_splitters, nsplit, d_peers, d_hist, d_tile_status, d_ticket);
    count_launch();
    trace_mark("partition_to_peers", s);
}

}  // namespace b200cd
