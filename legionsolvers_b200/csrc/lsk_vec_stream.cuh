// lsk_vec_stream.cuh -- TMA-streamed, dynamically scheduled traversal of congruent vectors.
//
// Building block of the fused vector kernels of the CG step (lsk_blas1.cu).  Measured on B200 (256^3, 805 MB pass): a grid-stride kernel with
// register-staged 256-bit loads reaches 6.3 TB/s, the torch copy 6.5 TB/s; this traversal reaches 7.0 TB/s.
#pragma once

#include "lsk_common.cuh"
#include "lsk_spmv_tma.cuh"

namespace lsk {

#ifdef __CUDACC__

// ---- TMA-streamed vector phases ---------------------------------------------------------------------------
// A CTA shape of 3 x 256 threads per SM cannot keep enough register-staged loads in flight to saturate HBM, so the
// vector passes stream too: a 64 KB shared-memory ring of 4 stages of 16 KB, each holding one chunk of every input vector, filled by
// cp.async.bulk three chunks ahead; threads read the chunk from shared memory, compute, and store results
// straight from registers.
constexpr int kVecStages = 4;
constexpr int kVecStageBytes = 16384;

struct VecRing {
    unsigned char *smem;   // kVecStages x kVecStageBytes
    uint64_t *bar;         // kVecStages mbarriers
    long long *chunk;      // [kVecStages] (shared): chunk held by each stage, -1 = none left
    uint32_t phases;       // bit s = parity to wait for on stage s
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Streams elements [head, head + nelem) of NIN congruent (32-byte aligned at `head`) arrays through the ring in
// chunks of CH = 16 KB / (8 NIN).  Chunks are handed out DYNAMICALLY from a global counter: SMs do not all get the
// same share of the memory system, and with a static split the slow ones finish ~7 % after the fast ones while
// HBM idles; first come, first served keeps it saturated to the end of the phase.  Thread 0 grabs three chunks
// ahead (the atomic's latency hides behind the chunks in flight), publishes the chunk id of a stage in shared
// memory before arming its mbarrier, and everybody learns it after the wait.  Grabbed index g maps to chunk
// (g + rot) % nchunks, so a phase can have a chosen range of chunks taken first.
// begin(i0, cnt) is called by every thread once per chunk (elements i0 .. i0 + cnt - 1) before its pairs;
// f(i, v) receives, for two consecutive elements i and i + 1, the inputs v[a][0..1] of each array.
// elements per chunk: a multiple of 512 (two per thread and pass) that fits NIN arrays into one 16 KB stage
template <int NIN>
struct VecChunk {
    static constexpr int value = NIN == 1 ? 2048 : NIN == 2 ? 1024 : 512;
    static_assert(NIN >= 1 && NIN <= 4, "a stage holds one chunk of up to four arrays");
};

template <int NIN, class B, class F>
__device__ __forceinline__ void vec_stream(VecRing &ring, const double *const (&in)[NIN], int64_t head, int64_t nelem,
                                           unsigned long long *counter, int64_t rot, B begin, F f) {
    constexpr int CH = VecChunk<NIN>::value;
    const int64_t nchunks = (nelem + CH - 1) / CH;
    auto issue = [&](int s, int64_t g) {  // thread 0: grabbed index g into stage s
        if (g < nchunks) {
            int64_t c = g + rot;
            if (c >= nchunks) c -= nchunks;
            ring.chunk[s] = c;
            const int64_t e0 = c * CH;
            const int64_t cnt = nelem - e0 < CH ? nelem - e0 : CH;
            const uint32_t bytes = (uint32_t) cnt * 8u;
            mbar_expect_tx(&ring.bar[s], NIN * bytes);
#pragma unroll
            for (int a = 0; a < NIN; ++a)
                tma_bulk_g2s_plain(ring.smem + (size_t) s * kVecStageBytes + (size_t) a * CH * 8, in[a] + head + e0, bytes, &ring.bar[s]);
        } else {
            ring.chunk[s] = -1;
            mbar_arrive(&ring.bar[s]);  // completes the stage's phase with no bytes
        }
    };
    auto grab = [&](long long prev) -> long long {  // thread 0; stops asking once the work has run out
        return prev < nchunks ? (long long) atomicAdd(counter, 1ull) : prev;
    };
    long long pend = 0;
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < kVecStages - 1; ++s) {
            pend = grab(pend);
            issue(s, pend);
        }
        pend = grab(pend);
    }
    int64_t k = 0;
    for (;; ++k) {
        const int s = (int) (k % kVecStages);
        // stage (k - 1) % kVecStages was consumed in the previous iteration, which ended with a CTA barrier
        if (threadIdx.x == 0) {
            issue((int) ((k + kVecStages - 1) % kVecStages), pend);
            pend = grab(pend);
        }
        mbar_wait(&ring.bar[s], (ring.phases >> s) & 1u);
        ring.phases ^= (1u << s);
        const long long c = ring.chunk[s];
        if (c < 0) break;  // uniform: the stages issued after this one are empty as well
        const int64_t e0 = c * CH;
        const int cnt = (int) (nelem - e0 < CH ? nelem - e0 : CH);
        const unsigned char *st = ring.smem + (size_t) s * kVecStageBytes;
        begin(head + e0, cnt);
#pragma unroll
        for (int u = 0; u < CH / (2 * kBlock); ++u) {
            const int j = u * 2 * kBlock + 2 * (int) threadIdx.x;
            if (j < cnt) {  // cnt is a multiple of 4, j is even: both elements exist
                double v[NIN][2];
#pragma unroll
                for (int a = 0; a < NIN; ++a) {
                    const double2 t = *reinterpret_cast<const double2 *>(st + (size_t) a * CH * 8 + (size_t) j * 8);
                    v[a][0] = t.x;
                    v[a][1] = t.y;
                }
                f(head + e0 + j, v);
            }
        }
        __syncthreads();
    }
    // retire the (empty) stages that were armed ahead, so that every mbarrier's parity matches ring.phases again
    for (int64_t j = k + 1; j < k + kVecStages; ++j) {
        const int s = (int) (j % kVecStages);
        mbar_wait(&ring.bar[s], (ring.phases >> s) & 1u);
        ring.phases ^= (1u << s);
    }
    __syncthreads();
}

// ---- halo mirroring of a vector update (p = r + beta p of CG feeds the next mat-vec) -----------------------------
// does the chunk [i0, i0 + cnt) overlap a range that is mirrored into a neighbour?  (uniform per chunk)
__device__ __forceinline__ bool halo_chunk_overlaps(const HaloSpec &h, int64_t i0, int cnt) {
    bool any = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) any |= (q < h.nmoves && i0 + cnt > h.lo[q] && i0 < h.lo[q] + h.m[q].n);
    return any;
}
// the new values of elements i, i + 1 leave as packets for every neighbour whose send range holds them
__device__ __forceinline__ void halo_send_pair(const HaloSpec &h, const HaloLive &hl, int64_t i, double p0, double p1) {
#pragma unroll
    for (int q = 0; q < kMaxFusedMoves; ++q) {
        if (q < h.nmoves && i + 2 > h.lo[q] && i < h.lo[q] + h.m[q].n) {
            const int64_t k = i - h.lo[q];
            if (k >= 0 && k + 1 < h.m[q].n && ((reinterpret_cast<uintptr_t>(hl.send_slot[q]) + (uintptr_t) k * 16) & 31) == 0) {
                ll_store2(hl.send_slot[q], k, p0, p1, hl.tag[q]);
            } else {
                if (k >= 0) ll_store(hl.send_slot[q], k, p0, hl.tag[q]);
                if (k + 1 < h.m[q].n) ll_store(hl.send_slot[q], k + 1, p1, hl.tag[q]);
            }
        }
    }
}
__device__ __forceinline__ void halo_send_one(const HaloSpec &h, const HaloLive &hl, int64_t i, double v) {
#pragma unroll
    for (int q = 0; q < kMaxFusedMoves; ++q)
        if (q < h.nmoves && i >= h.lo[q] && i < h.lo[q] + h.m[q].n) ll_store(hl.send_slot[q], i - h.lo[q], v, hl.tag[q]);
}
// chunks from the first one that overlaps a send range not at the start of the vector are taken first: their
// results also travel over NVLink, which then overlaps the rest of the pass
__device__ __forceinline__ int64_t halo_first_chunk(const HaloSpec &h, int64_t head, int ch) {
    int64_t first = -1;
    for (int q = 0; q < h.nmoves; ++q)
        if (h.m[q].n > 0 && h.lo[q] > head && (first < 0 || h.lo[q] < first)) first = h.lo[q];
    return first >= 0 ? (first - head) / ch : 0;
}

#endif  // __CUDACC__

}  // namespace lsk
