// Shared device/host helpers for the qlidar sm_100a kernels: error codes, PTX wrappers (mbarrier,
// cp.async, cp.async.bulk, tcgen05), the open-addressing coordinate hash, and the grid/linearisation rules.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/qlidar.h"

#define QL_TILE_M 128            // output rows per MMA tile == TMEM lanes == rulebook tile height
#define QL_NUM_SMS_DEFAULT 148
#define QL_KVOL_MAX 343           // 7^3
#define QL_MASK_WORDS_MAX 11     // ceil(QL_KVOL_MAX / 32) words of per-tile offset mask

#define QL_CUDA_CHECK_LAST()                                   \
    do {                                                       \
        cudaError_t e__ = cudaPeekAtLastError();               \
        if (e__ != cudaSuccess) return QL_ERR_CUDA;            \
    } while (0)

static inline int ql_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = QL_NUM_SMS_DEFAULT;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
// Grid description shared by every index kernel: coords are int32 [b, z, y, x]; the linear key is
// ((b*D + z)*H + y)*W + x and must fit 32 bits (checked on the host; QL_ERR_GRID_TOO_LARGE otherwise).
// ------------------------------------------------------------------------------------------------
struct QlGrid {
    int B, D, H, W;
};

__host__ __device__ __forceinline__ uint32_t ql_key(const QlGrid& g, int b, int z, int y, int x) {
    return ((uint32_t)((b * g.D + z) * g.H + y)) * (uint32_t)g.W + (uint32_t)x;
}

__device__ __forceinline__ int4 ql_unkey(const QlGrid& g, uint32_t key) {
    int4 c;
    c.w = (int)(key % (uint32_t)g.W); key /= (uint32_t)g.W;      // x
    c.z = (int)(key % (uint32_t)g.H); key /= (uint32_t)g.H;      // y
    c.y = (int)(key % (uint32_t)g.D); key /= (uint32_t)g.D;      // z
    c.x = (int)key;                                              // b
    return c;
}

// ------------------------------------------------------------------------------------------------
// Open-addressing hash: slot = {key, value} packed in a uint2 (8 B, one load per probe), linear probing,
// capacity a power of two >= 2 * entries (load factor <= 0.5).  EMPTY key = 0xFFFFFFFF.
// ------------------------------------------------------------------------------------------------
#define QL_HASH_EMPTY 0xFFFFFFFFu

// BUCKETISED: the four cells key & ~3 .. key | 3 (adjacent along x, the fastest axis of the key) hash to the four 8-byte
// slots of ONE 32-byte sector, so the kw = 3 probes of a rulebook line touch one or two sectors instead of three random
// ones (the probe kernels are bound by L2 random-sector rate).  Collisions step a whole bucket (ql_hash_next).
__device__ __forceinline__ uint32_t ql_hash_slot(uint32_t key, uint32_t cap_mask) {
    uint32_t h = (key >> 2) * 0x9E3779B1u;
    h ^= h >> 15;
    return ((h << 2) | (key & 3u)) & cap_mask;
}
__device__ __forceinline__ uint32_t ql_hash_next(uint32_t s, uint32_t cap_mask) { return (s + 4u) & cap_mask; }

// returns the slot index that holds `key` (inserting it if absent); value word untouched.
__device__ __forceinline__ uint32_t ql_hash_insert(uint2* __restrict__ table, uint32_t cap_mask, uint32_t key) {
    uint32_t s = ql_hash_slot(key, cap_mask);
    while (true) {
        uint32_t prev = atomicCAS(&table[s].x, QL_HASH_EMPTY, key);
        if (prev == QL_HASH_EMPTY || prev == key) return s;
        s = ql_hash_next(s, cap_mask);
    }
}

// returns slot index or 0xFFFFFFFF when absent
__device__ __forceinline__ uint32_t ql_hash_find_slot(const uint2* __restrict__ table, uint32_t cap_mask, uint32_t key) {
    uint32_t s = ql_hash_slot(key, cap_mask);
    while (true) {
        uint32_t k = __ldg(&table[s].x);
        if (k == key) return s;
        if (k == QL_HASH_EMPTY) return 0xFFFFFFFFu;
        s = ql_hash_next(s, cap_mask);
    }
}

// returns the value stored for key, or -1
__device__ __forceinline__ int ql_hash_lookup(const uint2* __restrict__ table, uint32_t cap_mask, uint32_t key) {
    uint32_t s = ql_hash_slot(key, cap_mask);
    while (true) {
        uint2 e = __ldg(&table[s]);
        if (e.x == key) return (int)e.y;
        if (e.x == QL_HASH_EMPTY) return -1;
        s = ql_hash_next(s, cap_mask);
    }
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ql_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void ql_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ql_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ql_mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ql_mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
// Blocking poll: the thread is suspended in hardware until the phase completes or ~`hint_ns` elapse (without the hint the
// time slice is so short that 16 polling warps took most of the issue slots of the one warp that feeds the tensor core).
__device__ __forceinline__ bool ql_mbar_try_wait(uint32_t bar, uint32_t parity, uint32_t hint_ns = 20000u) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU box (a hang is a strike).  After ~4 s of failed polls the
// CTA traps, which surfaces as a launch failure on the host instead.
__device__ __forceinline__ uint64_t ql_globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void ql_mbar_wait(uint32_t bar, uint32_t parity) {
    if (ql_mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!ql_mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0) {
            uint64_t t = ql_globaltimer_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ull) __trap();
        }
    }
}

// 32-bit shared-memory load by shared-window address (LDS; no generic-address resolution, freely pipelined)
__device__ __forceinline__ int ql_lds_s32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// TMA 1-D bulk copy global -> shared::cta, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void ql_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// one lane of the (fully active) warp
__device__ __forceinline__ bool ql_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void ql_tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void ql_tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void ql_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void ql_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ql_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ql_tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread t <-> lane base+t)
__device__ __forceinline__ void ql_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ql_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ql_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
