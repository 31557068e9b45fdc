// Per-row candidate state shared by the two Hamming-scan variants: the running bound tau, the
// append-only candidate buffer in global memory and its exact in-place prune.
#pragma once
#include <cstdint>

namespace em2 {

// Exact in-place prune of a per-row candidate buffer to its k smallest (mismatch, id) keys.
// Invariant kept: among entries with equal mismatch count, ids are in increasing order (appends arrive
// in increasing id; the compaction below is stable), so "the r smallest ids among the ties" are simply
// the first r ties.  Returns (new count, new tau); state is passed by value so that it stays in registers
// at the (rare) call sites.
static __device__ __noinline__ uint2 pruneCandidates(uint64_t* buf, uint32_t count, uint32_t k, uint32_t tau)
{
    uint32_t lo = 0, hi = tau - 1;          // every stored mismatch count is < tau
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        uint32_t c = 0;
        for (uint32_t i = 0; i < count; i++) c += (uint32_t(buf[i] >> 32) <= mid);
        if (c >= k) hi = mid;
        else lo = mid + 1;
    }
    const uint32_t h = lo;
    uint32_t less = 0;
    for (uint32_t i = 0; i < count; i++) less += (uint32_t(buf[i] >> 32) < h);
    uint32_t r = k - less;                  // ties at h that still fit
    uint32_t j = 0;
    for (uint32_t i = 0; i < count; i++) {
        const uint64_t key = buf[i];
        const uint32_t m = uint32_t(key >> 32);
        bool keep = m < h;
        if (m == h && r > 0) {
            keep = true;
            r--;
        }
        if (keep) buf[j++] = key;
    }
    return make_uint2(j, h);                // later ids are larger: ties at h can no longer enter
}

struct RowState {
    uint64_t* buf;
    uint32_t count;
    uint32_t tau;
    uint32_t rowId;
    uint32_t appended;
};

static __device__ __forceinline__ void consider(RowState& st, uint32_t ham, uint32_t id, uint32_t colEnd, uint32_t k,
                                                uint32_t cap)
{
    if (ham < st.tau && id < colEnd && id != st.rowId) {
        st.buf[st.count++] = (uint64_t(ham) << 32) | id;
        st.appended++;
        if (st.count == cap) {
            const uint2 r = pruneCandidates(st.buf, st.count, k, st.tau);
            st.count = r.x;
            st.tau = r.y;
        }
    }
}

}  // namespace em2
