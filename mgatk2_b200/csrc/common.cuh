// mgatk2_b200 — shared definitions of the sm_100a kernels: the slot layouts that carry a read from the partition to
// the pileup, shared / global memory word accessors, and the mbarrier + bulk-copy (TMA) primitives.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mgatk2_b200.h"
#include "bitplane.cuh"

namespace mgatk {

typedef unsigned long long u64;

constexpr u32 kFull = 0xffffffffu;
constexpr u32 ERR_UNSORTED = 1, ERR_EXTENT = 2, ERR_OVERFLOW_CAP = 4, ERR_SATURATED = 8;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---------------------------------------------------------------------------------------------
// Slots. Stage 1 turns every record that passes the flag / whitelist filter into ONE self-contained slot: the
// dedup key, a few flags and the read's bases as three bit planes in REFERENCE coordinates anchored at its start
// (bit i <=> reference_start + i): V = "counts here" (aligned, A/C/G/T, base quality, distance-from-end window,
// pileup.py:52-86), B0 / B1 = the two bits of the base code under V. Nothing downstream touches SEQ, QUAL or CIGAR.
//
//   compact (max_read_extent <= 56), 32 bytes = one sector:
//       w0 pos   w1 |tlen|   w2 V[0:32)   w3 V[32:56) | meta << 24 | hi[8:11) << 29   w4 B0[0:32)
//       w5 B0[32:56) | tn5off << 24   w6 B1[0:32)   w7 B1[32:56) | hi[0:8) << 24
//       hi = high digit of the cell index, carried from the first to the second pass of a two-digit partition (more than
//       2048 cells); after the partition a compact slot's cell follows from its place (table of first slots per cell)
//   wide (W = ceil(extent / 32) words per plane, at most 8), 16 + 16 W bytes rounded up to whole sectors:
//       w0 pos   w1 |tlen|   w2 cell | meta << 24   w3 tn5off   then W groups { V, B0, B1, 0 } of 32 offsets each
//       an INDIRECT read (longer than 32 W) keeps { blob_off, l_seq | n_cigar << 16 } in its first group instead and
//       is counted base by base from the caller's blob.
// tn5off = len(SEQ) - 1, the distance of a reverse read's Tn5 site from its start (pileup.py:43-46).
// ---------------------------------------------------------------------------------------------
constexpr u32 SM_STRAND = 1, SM_PAIRED = 2, SM_MAPQ_OK = 4, SM_EMPTY = 8, SM_INDIRECT = 16;
constexpr int kCompactExtent = 56;
constexpr int kMaxPlaneWords = 8;

struct SlotFmt { int compact, words, bytes; };

__host__ __device__ inline SlotFmt slot_format(int extent) {
    SlotFmt f;
    f.compact = extent <= kCompactExtent;
    int w = (extent + 31) / 32;
    f.words = f.compact ? 2 : (w < 2 ? 2 : w > kMaxPlaneWords ? kMaxPlaneWords : w);
    f.bytes = f.compact ? 32 : (16 + 16 * f.words + 31) / 32 * 32;      // whole 32-byte sectors: scattered slot stores never straddle one
    return f;
}

// first 16 bytes of a slot = everything dedup and planning need
template <bool kCompact> struct SlotKey;
template <> struct SlotKey<true> {
    __device__ __forceinline__ static u32 meta(const uint4 &k) { return (k.w >> 24) & 31u; }
};
template <> struct SlotKey<false> {
    __device__ __forceinline__ static u32 meta(const uint4 &k) { return k.z >> 24; }
    __device__ __forceinline__ static int cell(const uint4 &k) { return (int)(k.z & 0xffffffu); }
};

// ---------------------------------------------------------------------------------------------
// word access to staging memory (bitplane.cuh's `M`)
// ---------------------------------------------------------------------------------------------
struct SharedMem {                   // by 32-bit shared address
    __device__ __forceinline__ u32 ld32(u32 a) const { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
    __device__ __forceinline__ void st128(u32 a, u32 x, u32 y, u32 z, u32 w) const {
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
    }
    __device__ __forceinline__ void ld128(u32 a, u32 (&v)[4]) const {
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(a));
    }
};

struct SharedMemPure {               // loads the compiler may merge and schedule (no `volatile`): only for data that is not
    __device__ __forceinline__ u32 ld32(u32 a) const { u32 v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }   // rewritten while in use
};

struct GlobalBlob {                  // one record's cigar|seq|qual in the caller's blob, by offset; nothing is read beyond `avail`
    const uint8_t *base; u32 avail;
    __device__ __forceinline__ u32 ld32(u32 a) const { return a + 4u <= avail ? __ldg(reinterpret_cast<const u32 *>(base + a)) : 0u; }
};

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk asynchronous copy (the TMA unit's linear mode): one elected thread arms the barrier with the byte
// count and issues `cp.async.bulk`; the consumers spin on the barrier's phase parity.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(u32 bar, u32 parity) {
    u32 ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) { while (!mbar_try_wait(bar, parity)) {} }
// the same with a suspend-time hint (ns): the thread sleeps in hardware instead of burning issue slots on the probe
__device__ __forceinline__ void mbar_wait_sleepy(u32 bar, u32 parity, u32 hint_ns) {
    u32 ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(hint_ns) : "memory");
    } while (!ok);
}
// wait, then hand out a zero that DEPENDS on the wait: added to a shared address it keeps reordering-free (non-volatile)
// loads of the data the barrier guards behind the wait
__device__ __forceinline__ u32 mbar_wait_token(u32 bar, u32 parity) {
    mbar_wait(bar, parity);
    u32 z;
    asm volatile("mov.u32 %0, 0;" : "=r"(z) :: "memory");
    return z;
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void bulk_load(u32 dst_shared, const void *src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_shared), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// generic-proxy accesses to shared memory ordered before later async-proxy (bulk copy) writes to it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace mgatk
