// Per-row streaming selection shared by the two Hamming-scan variants: the running bound tau, the
// append-only candidate buffer in global memory and its exact in-place prune.
//
// Each (row, column stream) -- one thread -- appends the composite keys (mismatch << 32 | cellId) of
// accepted candidates to a private region of global memory (L1/L2 resident, touched only on the rare
// accepted candidate).  When a region runs out of slack it is pruned in place to its k best keys and tau
// becomes the k-th best mismatch count.  Columns of a stream are visited in increasing cell id, so a new
// candidate's id is larger than every kept id: it can displace a kept key only if its mismatch count is
// STRICTLY smaller than the k-th best.  The hot-path filter is therefore the single compare `ham < tau`,
// and ties at the k-th place resolve to the smaller id, exactly as the (mismatch asc, id asc) order of the
// reference's deterministic selection (reference src/ExpressionMatrixLshGpu.cpp:148-149) requires.
//
// The prune is WARP-COOPERATIVE: at a point where the warp is converged, lanes whose buffers are nearly
// full are served one after the other by all 32 lanes (count by ballot/reduce, stable compaction by
// ballot prefix sums), ~400 issue slots instead of ~5000 for a single-thread prune under divergence.
// (Measured alternatives: single-thread prune -- 1.5-4x slower scans on clustered data; a k-entry max-heap
// with an exact bound -- 40% fewer accepted candidates but a 6-level dependent sift through L1 each.)
#pragma once
#include <cstdint>

namespace em2 {

constexpr uint32_t kPruneSlack = 32;    // appends a thread may make between two prune points

struct RowState {
    uint64_t* buf;
    uint32_t count;
    uint32_t tau;       // this stream's own bound: every stored mismatch count is < tau
    uint32_t lim;       // filter: accept iff ham < lim; lim <= tau (a kernel may tighten it with bounds
                        // learned by other streams of the same row)
    uint32_t rowId;
    uint32_t appended;
};

// capacity of a candidate region for a given k
__host__ __device__ inline uint32_t candidateCapacity(uint32_t k) { return 2 * k + kPruneSlack; }
// The capacity the scans use: as large as the register prune allows (256 keys), up to 4k + slack -- a prune costs
// the same for 132 or 232 keys, so fewer, larger prunes are cheaper (config 2: 8.33 -> 8.01 ms; beyond 256 keys the
// memory-walking prune takes over and the scan is slower, 9.6 ms).  extra > 0 (option "cand_cap_extra") forces
// (2 + extra) k + slack.
__host__ __device__ inline uint32_t scanCandidateCapacity(uint32_t k, uint32_t extra)
{
    const uint32_t base = candidateCapacity(k);
    if (extra) return base + k * extra;
    if (base > 256) return base;
    const uint32_t wide = 4 * k + kPruneSlack < 256 ? 4 * k + kPruneSlack : 256;
    return wide > base ? wide : base;
}

// Plain append; the caller guarantees at most kPruneSlack appends between two warpPruneIfNeeded() calls.
static __device__ __forceinline__ void consider(RowState& st, uint32_t ham, uint32_t id, uint32_t colEnd)
{
    if (ham < st.lim && id < colEnd && id != st.rowId) {
        st.buf[st.count++] = (uint64_t(ham) << 32) | id;
        st.appended++;
    }
}

// Exact in-place prune of ONE candidate region to its k smallest (mismatch, id) keys, executed by the
// whole (converged) warp; buf/count/tau are warp-uniform.  Invariant kept: among entries with equal
// mismatch count ids are in increasing order (appends arrive in increasing id; the compaction is stable),
// so "the r smallest ids among the ties" are simply the first r ties.  Returns the new bound; the new count
// is k.
static __device__ __noinline__ uint32_t warpPrune(uint64_t* buf, uint32_t count, uint32_t k, uint32_t tau)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t lo = 0, hi = tau - 1;          // every stored mismatch count is < tau
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        uint32_t c = 0;
        for (uint32_t i = lane; i < count; i += 32) c += (uint32_t(buf[i] >> 32) <= mid);
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= k) hi = mid;
        else lo = mid + 1;
    }
    const uint32_t h = lo;
    uint32_t less = 0;
    for (uint32_t i = lane; i < count; i += 32) less += (uint32_t(buf[i] >> 32) < h);
    less = __reduce_add_sync(0xffffffffu, less);
    const uint32_t r = k - less;            // ties at h that still fit
    uint32_t out = 0, tiesBefore = 0;
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t base = 0; base < count; base += 32) {
        const uint32_t i = base + lane;
        const uint64_t key = i < count ? buf[i] : ~0ull;
        const uint32_t m = uint32_t(key >> 32);
        const bool tie = (i < count) && (m == h);
        const uint32_t tieMask = __ballot_sync(0xffffffffu, tie);
        const bool keep = (i < count) && (m < h || (tie && tiesBefore + __popc(tieMask & lt) < r));
        const uint32_t keepMask = __ballot_sync(0xffffffffu, keep);
        if (keep) buf[out + __popc(keepMask & lt)] = key;      // out + rank <= i: never overtakes the reads
        out += __popc(keepMask);
        tiesBefore += __popc(tieMask);
    }
    __syncwarp();
    return h;                               // later ids are larger: ties at h can no longer enter
}

// Same contract, for regions of at most 32*EPL keys: the region is read ONCE into registers (EPL independent
// loads per lane, one memory round trip), the k-th smallest mismatch count is found by bisection on registers
// only, and the stable compaction writes straight from registers.  ~4x shorter than the version above, whose
// every bisection step goes back to memory -- and prune latency is what stalls the MMA pipeline of the scan
// on clustered data.
template <int EPL>
static __device__ __noinline__ uint32_t warpPruneRegs(uint64_t* buf, uint32_t count, uint32_t k, uint32_t tau)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t m[EPL], id[EPL];
#pragma unroll
    for (int e = 0; e < EPL; e++) {
        const uint32_t i = e * 32 + lane;
        const uint64_t key = i < count ? buf[i] : ~0ull;
        m[e] = uint32_t(key >> 32);          // 0xffffffff for the slots beyond count: never selected
        id[e] = uint32_t(key);
    }
    uint32_t lo = 0, hi = tau - 1;           // every stored mismatch count is < tau
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        uint32_t c = 0;
#pragma unroll
        for (int e = 0; e < EPL; e++) c += (m[e] <= mid);
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= k) hi = mid;
        else lo = mid + 1;
    }
    const uint32_t h = lo;
    uint32_t less = 0;
#pragma unroll
    for (int e = 0; e < EPL; e++) less += (m[e] < h);
    less = __reduce_add_sync(0xffffffffu, less);
    const uint32_t r = k - less;             // ties at h that still fit
    uint32_t out = 0, tiesBefore = 0;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int e = 0; e < EPL; e++) {
        if (uint32_t(e) * 32 < count) {      // warp-uniform
            const bool tie = m[e] == h;
            const uint32_t tieMask = __ballot_sync(0xffffffffu, tie);
            const bool keep = m[e] < h || (tie && tiesBefore + __popc(tieMask & lt) < r);
            const uint32_t keepMask = __ballot_sync(0xffffffffu, keep);
            if (keep) buf[out + __popc(keepMask & lt)] = (uint64_t(m[e]) << 32) | id[e];
            out += __popc(keepMask);
            tiesBefore += __popc(tieMask);
        }
    }
    __syncwarp();
    return h;
}

constexpr int kPruneRegsPerLane = 8;         // register prune for regions of up to 256 keys

// Call with the warp converged.  Serves every lane whose region has less than kPruneSlack free slots.
static __device__ __forceinline__ void warpPruneIfNeeded(RowState& st, uint32_t k, uint32_t cap)
{
    uint32_t need = __ballot_sync(0xffffffffu, st.count + kPruneSlack > cap);
    if (need) __syncwarp();                 // the owners' appends become visible to the helping lanes
    while (need) {
        const int src = __ffs(int(need)) - 1;
        need &= need - 1;
        const uint64_t b = __shfl_sync(0xffffffffu, reinterpret_cast<uint64_t>(st.buf), src);
        const uint32_t c = __shfl_sync(0xffffffffu, st.count, src);
        const uint32_t t = __shfl_sync(0xffffffffu, st.tau, src);
        const uint32_t h = cap <= 32 * kPruneRegsPerLane
                               ? warpPruneRegs<kPruneRegsPerLane>(reinterpret_cast<uint64_t*>(b), c, k, t)
                               : warpPrune(reinterpret_cast<uint64_t*>(b), c, k, t);
        if (int(threadIdx.x & 31) == src) {
            st.count = k;
            st.tau = h;
            st.lim = st.lim < h ? st.lim : h;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Order-independent forms, used by the symmetric scan (scan_mma.cu), whose streams do NOT visit candidates in
// increasing id: a later candidate that ties with the k-th best mismatch count may carry a smaller id and win.
// Bounds are therefore kept in EXCLUSIVE form over the mismatch count alone -- after a prune that found h as the
// k-th smallest count, tau = h + 1 (a tie is still accepted) -- and the prune keeps the k smallest full keys
// (mismatch, id), choosing among the ties at h by id with a second bisection.
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t kTieSurplus = 16;

template <int EPL>
static __device__ __noinline__ uint32_t warpPruneRegsAnyOrder(uint64_t* buf, uint32_t count, uint32_t k, uint32_t tau,
                                                              const uint32_t* __restrict__ idOf, uint32_t* newCount)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t m[EPL], id[EPL];
#pragma unroll
    for (int e = 0; e < EPL; e++) {
        const uint32_t i = e * 32 + lane;
        const uint64_t key = i < count ? buf[i] : ~0ull;
        m[e] = uint32_t(key >> 32);
        id[e] = uint32_t(key);
    }
    uint32_t lo = 0, hi = tau - 1;           // every stored mismatch count is < tau
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        uint32_t c = 0;
#pragma unroll
        for (int e = 0; e < EPL; e++) c += (m[e] <= mid);
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= k) hi = mid;
        else lo = mid + 1;
    }
    const uint32_t h = lo;
    uint32_t less = 0, ties = 0;
#pragma unroll
    for (int e = 0; e < EPL; e++) {
        less += (m[e] < h);
        ties += (m[e] == h);
    }
    less = __reduce_add_sync(0xffffffffu, less);
    ties = __reduce_add_sync(0xffffffffu, ties);
    const uint32_t r = k - less;             // ties at h that still fit: those with the r smallest ids
    // A few surplus ties are simply kept (the region then holds a little more than k keys); only when they would eat
    // into the room for new candidates are the r smallest ids among them selected.
    if (ties > r + (k < kTieSurplus ? k : kTieSurplus)) {      // kept keys <= 2k: a slack of appends always fits
        // The low word of a key is a scan POSITION (the hot path appends without looking anything up); the tie-break
        // is on cell ids, looked up here -- for the ties only, all loads in flight at once.
        uint32_t pos[EPL];
#pragma unroll
        for (int e = 0; e < EPL; e++) {
            pos[e] = id[e];
            if (idOf && m[e] == h) id[e] = idOf[pos[e]];
        }
        uint32_t a = 0, b = 0xffffffffu;
        while (a < b) {
            const uint32_t mid = a + ((b - a) >> 1);
            uint32_t c = 0;
#pragma unroll
            for (int e = 0; e < EPL; e++) c += (m[e] == h && id[e] <= mid);
            c = __reduce_add_sync(0xffffffffu, c);
            if (c >= r) b = mid;
            else a = mid + 1;
        }
        const uint32_t idCut = a;             // ids are unique: exactly r ties have id <= idCut
        // back to positions, with the losing ties marked
#pragma unroll
        for (int e = 0; e < EPL; e++) {
            if (m[e] == h && id[e] > idCut) m[e] = 0xffffffffu;
            id[e] = pos[e];
        }
    }
    uint32_t out = 0;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int e = 0; e < EPL; e++) {
        if (uint32_t(e) * 32 < count) {      // warp-uniform
            const bool keep = m[e] <= h;         // losing ties were marked above
            const uint32_t keepMask = __ballot_sync(0xffffffffu, keep);
            if (keep) buf[out + __popc(keepMask & lt)] = (uint64_t(m[e]) << 32) | id[e];
            out += __popc(keepMask);
        }
    }
    __syncwarp();
    *newCount = out;
    return h;
}

// Call with the warp converged; regions of at most 32 * kPruneRegsPerLane keys.  The owner's new bound is also
// published (atomicMin) to *shared, the row's entry of a global bound array read by other CTAs.
// force: prune every region that holds at least k keys (end of the near window: every row publishes the bound it has
// learned, whether or not its region ever filled up).
static __device__ __forceinline__ void warpPruneIfNeededAnyOrder(RowState& st, uint32_t k, uint32_t cap, uint32_t* shared,
                                                                 const uint32_t* __restrict__ idOf, bool force = false)
{
    uint32_t need = __ballot_sync(0xffffffffu, force ? (st.count >= k) : (st.count + kPruneSlack > cap));
    if (need) __syncwarp();
    while (need) {
        const int src = __ffs(int(need)) - 1;
        need &= need - 1;
        const uint64_t b = __shfl_sync(0xffffffffu, reinterpret_cast<uint64_t>(st.buf), src);
        const uint32_t c = __shfl_sync(0xffffffffu, st.count, src);
        const uint32_t t = __shfl_sync(0xffffffffu, st.tau, src);
        uint32_t kept;
        const uint32_t h = warpPruneRegsAnyOrder<kPruneRegsPerLane>(reinterpret_cast<uint64_t*>(b), c, k, t, idOf, &kept);
        if (int(threadIdx.x & 31) == src) {
            st.count = kept;          // k, or up to k + kTieSurplus when ties at h were kept
            st.tau = h + 1;
            st.lim = st.lim < h + 1 ? st.lim : h + 1;
            atomicMin(shared, h + 1);
        }
    }
}

}  // namespace em2
