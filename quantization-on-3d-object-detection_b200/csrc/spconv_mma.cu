// Implicit gather-GEMM-scatter sparse convolution on tcgen05 tensor cores (sm_100a), fused epilogue.
//
// Replaces QConvNd.forward -> [EXT] spconv SubMConv3d/SparseConv3d forward (quant/quant.py:36-58) together with
// the BatchNorm1d / ReLU / residual-add that follow it (spconv_backbone.py:8-27,51-67).
//
// One persistent CTA per SM walks 128-row output tiles.  Per tile the conv is a sum of per-offset GEMMs
//     D[128, C_out] += A_k[128, C_in] . W_k[C_out, C_in]^T      for every kernel offset k that is non-empty in the tile
// (the rulebook's per-tile offset mask lists them; empty (tile, offset) slabs cost nothing).  The A operand never
// exists in memory and never touches shared memory: output row r of the tile is TMEM lane r, and the producer thread
// that owns lane r loads its neighbour's feature row straight from global memory (L2) into registers and writes it
// to tensor memory with tcgen05.st; the MMA reads A from TMEM (the ".ts" operand form) and only the weights from
// shared memory.  ncu on the previous cp.async -> SWIZZLE_128B smem version showed the shared-memory data pipe at
// 75 % (one write wavefront per returning 32-byte sector plus zero fills plus the tensor core's operand reads) with
// DRAM at 17 %, see profiles/r01_conv_v2_smem_gather.md; TMEM stores run at 256 B/clk and are off that pipe.
//
// K is cut into sub-chunks: one sub-chunk = one kernel offset x one <=128-byte segment of the input row
// (CH = 32 / 64 / 128 bytes => 1 / 2 / 4 MMA k-steps).  128/CH consecutive sub-chunks form a unit (32 TMEM columns =
// 32 registers per producer thread), kTeams units form a ring slot: one full/empty mbarrier pair and one pass of the
// single MMA-issuing thread per slot (that thread's instruction stream is the serial bottleneck of the kernel, so the
// work per barrier round trip is made as large as TMEM allows).  Roles (18 warps):
//   warps 0-11  gather producers : 3 teams x 4 warps; warp w owns TMEM lanes 32*(w%4)..+31; team t fills unit t of every
//                                  slot; per sub-chunk: nbr index (LDS from the tile's rulebook slab), CH bytes of the
//                                  neighbour row (or zeros); one tcgen05.st.x32 per unit, mbarrier arrive; the loads of a
//                                  team's next unit are issued before the current unit is stored
//   warps 12-15 epilogue         : tcgen05.ld accumulators -> dequant*BN scale/shift (+residual) (+ReLU)
//                                  -> fp16/fp32 rows (+ int8 re-quantised rows, + per-channel absmax)
//   warp  16    MMA issuer       : one lane issues tcgen05.mma (kind::f16 or kind::i8), accumulators in TMEM
//                                  (double buffered when they fit: tile i+1 accumulates while tile i drains)
//   warp  17    TMA loader       : cp.async.bulk of each chunk's weight slab (pre-swizzled image from
//                                  ql_pack_weights_host) into the chunk's B slot -- or, when the whole packed weight tensor
//                                  fits in shared memory (C <= 32 fp16, C <= 64 int8), of all of it once -- and of the
//                                  NEXT tile's non-empty rulebook slabs (512 B each) into a double-buffered copy.
//                                  One UBLKCP instruction costs its issuing warp ~225 ns whatever the size
//                                  (tools/microbench/bulk_copy_rate.cu), an extra active lane only ~26 ns: copies are issued
//                                  several lanes at a time, one copy per lane.
#include "ql_common.cuh"
#include <string.h>

namespace {

constexpr int kTeams = 3;
constexpr int kProducerWarps = kTeams * 4;               // 12
constexpr int kEpilogueThreads = 128;
constexpr int kEpilogueWarp0 = kProducerWarps;            // warps 12..15 (warp % 4 == TMEM lane quarter)
constexpr int kMmaWarp = kEpilogueWarp0 + 4;              // 16
constexpr int kLoaderWarp = kMmaWarp + 1;                 // 17
constexpr int kThreadsTotal = (kLoaderWarp + 1) * 32;     // 576
constexpr int kMaxSlots = 4;
constexpr int kSlotCols = 32 * kTeams;                    // TMEM columns per ring slot
constexpr int kMaskWords = 4;                             // kernel volumes up to 128 (3^3, 5^3)
constexpr int kTmemCols = 512;
constexpr int kSmemBudget = 232448;                       // 227 KB opt-in maximum per CTA
constexpr int kSmemFloor = 120 * 1024;                    // always ask for > half an SM: one CTA (one TMEM owner) per SM

struct ConvParams {
    const uint8_t* feats;
    const int* nbr;
    const uint32_t* kmask;  // [tiles][mask_words] or null (every offset)
    const int* n_out_dev;
    int64_t n_out_cap;
    int row_bytes;          // c_in * elem size
    int wide;               // rows (and the feature base) are 32-byte aligned: gather with 256-bit loads
    int c_out, kvol, nseg, mask_words;
    const uint8_t* w_packed;
    const float* scale;
    const float* shift;
    const float* act_scale_dev;
    const __half* residual;
    int relu;
    void* out;
    int out_dtype;
    int8_t* out_q;
    const float* out_qscale;
    float* absmax;
    int n_slots;            // A/B ring depth in slots
    int n_acc;              // accumulator buffers in TMEM (2, or 1 when 2*c_out does not fit beside the A ring)
    int a_col0;             // first TMEM column of the A ring
    int resident;           // 1: every weight chunk lives in shared memory for the whole kernel (no per-chunk B copies)
    int w_bytes;            // packed weight bytes (resident mode)
    int off_nbr;            // smem offset of the rulebook buffers: nbr_bufs x {16-byte tile mask, [kvol][128] int32}
    int nbr_bufs, nbr_log2; // 4 (or 2 when shared memory is short): the loader runs nbr_bufs-1 tiles ahead
    int nbr_stride;         // bytes per buffer
    int off_misc;           // smem offset of MiscSmem from the 1024-aligned base
};

struct MiscSmem {
    uint64_t full[kMaxSlots];
    uint64_t empty[kMaxSlots];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint64_t nbr_full[4];
    uint64_t nbr_empty[4];
    uint64_t w_full;
    uint32_t tmem_base;
    uint32_t pad[1];
    // followed by: float scale[c_out], float shift[c_out], uint32 absmax[c_out], float qscale[c_out]
};

template <bool kInt8>
__device__ __forceinline__ uint32_t make_idesc(int n) {
    uint32_t d = 0;
    if (kInt8) {
        d |= 2u << 4;        // D format S32
        d |= 1u << 7;        // A signed int8
        d |= 1u << 10;       // B signed int8
    } else {
        d |= 1u << 4;        // D format F32;  A,B formats 0 = F16
    }
    // a_major = b_major = 0 (K-major), no negate, dense
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(QL_TILE_M >> 4) << 24;
    return d;
}

// D[tmem] (+)= A[tmem] * B[smem desc]
template <bool kInt8>
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (kInt8) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

// K-major swizzled shared-memory matrix descriptor for a [rows x CH bytes] weight chunk (CH = 32 / 64 / 128):
// rows are CH bytes apart inside an 8-row swizzle atom, atoms are SBO = 8*CH bytes apart.
template <int CH>
__device__ __forceinline__ uint64_t umma_desc_b(uint32_t smem_addr) {
    constexpr uint64_t layout = CH == 128 ? 2 : (CH == 64 ? 4 : 6);    // SWIZZLE_128B / 64B / 32B
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                    // LBO (ignored for swizzled K-major)
    d |= (uint64_t)((8 * CH) >> 4) << 32;      // SBO
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= layout << 61;
    return d;
}

// 256-bit load (LDG.E.ENL2.256, sm_100): one full 32-byte sector per lane, half the L1 wavefronts of two 128-bit loads when
// every lane reads a different row.  Needs 32-byte alignment.
__device__ __forceinline__ void ldg32(const uint8_t* p, uint32_t (&v)[8]) {
    asm volatile("ld.global.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ uint4 ldg16(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <int NREG>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&v)[NREG]);
template <>
__device__ __forceinline__ void tmem_st<32>(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// position of the n-th (0-based) set bit of a tile mask
__device__ __forceinline__ int nth_set_bit(const uint32_t (&m)[kMaskWords], int n) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < kMaskWords; ++i) {
        const int c = __popc(m[i]);
        if (n >= 0 && n < c) k = i * 32 + (int)__fns(m[i], 0, n + 1);
        n -= c;                                          // goes negative once found: later words cannot match
    }
    return k;
}

__device__ __forceinline__ int load_tile_mask(const ConvParams& p, int64_t tile, uint32_t (&mask)[kMaskWords]) {
    int n = 0;
#pragma unroll
    for (int i = 0; i < kMaskWords; ++i) {
        uint32_t w = 0u;
        if (i < p.mask_words) {
            if (p.kmask) {
                w = __ldg(p.kmask + tile * p.mask_words + i);
            } else {
                const int rem = p.kvol - 32 * i;
                w = rem >= 32 ? 0xFFFFFFFFu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
            }
        }
        mask[i] = w;
        n += __popc(w);
    }
    if (n == 0) { mask[0] = 1u; n = 1; }       // a tile without pairs still has to zero its accumulators
    return n * p.nseg;
}

// the tile mask as the loader left it in the header of a rulebook buffer
__device__ __forceinline__ int load_tile_mask_smem(uint32_t addr, int nseg, uint32_t (&mask)[kMaskWords]) {
    int n = 0;
#pragma unroll
    for (int i = 0; i < kMaskWords; ++i) {
        mask[i] = (uint32_t)ql_lds_s32(addr + 4u * i);
        n += __popc(mask[i]);
    }
    return n * nseg;
}

template <bool kInt8, int CH, bool kResident>
__global__ void __launch_bounds__(kThreadsTotal, 1) k_spconv_ts(const ConvParams p) {
    constexpr int kAReg = CH / 4;                          // 32-bit TMEM columns (registers) per chunk
    constexpr int kGroup = 128 / CH;                       // sub-chunks per group == per ring slot (128 bytes of K, 32 TMEM columns)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base_u32 = (ql_smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle atoms need 1024-byte alignment
    uint8_t* smem = smem_raw + (smem_base_u32 - ql_smem_u32(smem_raw));
    MiscSmem* misc = reinterpret_cast<MiscSmem*>(smem + p.off_misc);
    float* s_scale = reinterpret_cast<float*>(misc + 1);
    float* s_shift = s_scale + p.c_out;
    uint32_t* s_absmax = reinterpret_cast<uint32_t*>(s_shift + p.c_out);
    float* s_qscale = reinterpret_cast<float*>(s_absmax + p.c_out);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int S = p.n_slots;

    const int64_t n_out = p.n_out_dev ? (int64_t)*p.n_out_dev : p.n_out_cap;
    const int64_t n_tiles = (n_out + QL_TILE_M - 1) / QL_TILE_M;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            ql_mbar_init(ql_smem_u32(&misc->full[s]), kProducerWarps + (kResident ? 0 : 1));   // every producer warp (+ the loader's expect_tx)
            ql_mbar_init(ql_smem_u32(&misc->empty[s]), 1);         // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            ql_mbar_init(ql_smem_u32(&misc->acc_full[i]), 1);
            ql_mbar_init(ql_smem_u32(&misc->acc_empty[i]), kEpilogueThreads);
        }
        for (int i = 0; i < 4; ++i) {
            ql_mbar_init(ql_smem_u32(&misc->nbr_full[i]), 1);
            ql_mbar_init(ql_smem_u32(&misc->nbr_empty[i]), kProducerWarps);
        }
        ql_mbar_init(ql_smem_u32(&misc->w_full), 1);
        ql_fence_mbar_init();
    }
    {
        const float act = p.act_scale_dev ? *p.act_scale_dev : 1.0f;
        for (int c = tid; c < p.c_out; c += kThreadsTotal) {
            s_scale[c] = p.scale[c] * act;
            s_shift[c] = p.shift[c];
            s_absmax[c] = 0u;
            s_qscale[c] = p.out_qscale ? p.out_qscale[c] : 0.f;
        }
    }
    if (warp == kMmaWarp) {
        ql_tmem_alloc(ql_smem_u32(&misc->tmem_base), kTmemCols);
        ql_tmem_relinquish();
    }
    ql_tc_fence_before();
    __syncthreads();
    ql_tc_fence_after();
    const uint32_t tmem_base = misc->tmem_base;
    const uint32_t b_sub_bytes = (uint32_t)p.c_out * CH;     // one weight sub-chunk: [c_out x CH bytes]
    constexpr int kSlotSubs = kGroup * kTeams;               // sub-chunks per ring slot

    if (warp < kProducerWarps) {
        // ============================ gather producers ============================
        const int q = warp & 3;                              // TMEM lane quarter
        const int team = warp >> 2;
        const int r = q * 32 + lane;                         // row in tile == TMEM lane
        const uint32_t a_lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)p.a_col0 + (uint32_t)(team * 32);
        const uint32_t nbr_s0 = smem_base_u32 + (uint32_t)p.off_nbr;            // buffer b: mask at +b*stride, slabs at +16
        const uint32_t nbr_s = nbr_s0 + 16u + (uint32_t)r * 4u;
        const uint32_t nbr_stride = (uint32_t)p.nbr_stride, nbmask = (uint32_t)p.nbr_bufs - 1u;

        // Walks the CTA's tiles slot by slot; this team owns unit `team` of every slot (a unit past the tile's last
        // sub-chunk is empty: nothing is loaded or stored, but the slot's barriers are still honoured, so every warp
        // sees every ring phase and the parity waits cannot alias).
        struct Unit { uint32_t ring, ph; int n; int c0; uint32_t buf_off; };
        int64_t tile = blockIdx.x;
        uint32_t it = 0, ring = 0, ph = 0;
        int n_sub = 0, n_slot_t = 0, sl = 0;
        bool have_tile = false, ready = false;
        // returns 1 = unit found, 0 = no more work, 2 = the next tile's rulebook slab has not landed yet (only if !blocking)
        auto next_unit = [&](bool blocking, Unit& out) -> int {
            while (true) {
                if (!have_tile) {
                    if (tile >= n_tiles) return 0;
                    have_tile = true; ready = false;
                }
                if (!ready) {
                    const uint32_t nb = it & nbmask, par = (it >> p.nbr_log2) & 1u;
                    const uint32_t bar = ql_smem_u32(&misc->nbr_full[nb]);
                    if (blocking) ql_mbar_wait(bar, par);
                    else if (!ql_mbar_test_wait(bar, par)) return 2;
                    uint32_t mask[kMaskWords];
                    n_sub = load_tile_mask_smem(nbr_s0 + nb * nbr_stride, p.nseg, mask);
                    n_slot_t = (n_sub + kSlotSubs - 1) / kSlotSubs;
                    sl = 0; ready = true;
                }
                if (sl < n_slot_t) {
                    out.ring = ring; out.ph = ph;
                    out.c0 = (sl * kTeams + team) * kGroup;
                    const int rem = n_sub - out.c0;
                    out.n = rem < 0 ? 0 : (rem < kGroup ? rem : kGroup);
                    out.buf_off = (it & nbmask) * nbr_stride;
                    ++sl;
                    if (++ring == (uint32_t)S) { ring = 0; ph ^= 1u; }
                    return 1;
                }
                // every index this warp needs from the tile's slab has been read: hand the buffer back
                __syncwarp();
                if (lane == 0) ql_mbar_arrive(ql_smem_u32(&misc->nbr_empty[it & nbmask]));
                tile += gridDim.x; ++it; have_tile = false;
            }
        };
        // loads of one unit: kGroup sub-chunks of CH bytes = 32 registers = the unit's 32 TMEM columns
        auto issue = [&](const Unit& u, uint32_t (&v)[32]) {
            int idx[kGroup], boff[kGroup];
#pragma unroll
            for (int j = 0; j < kGroup; ++j) {
                idx[j] = -1; boff[j] = 0;
                if (j < u.n) {
                    const int lc = u.c0 + j;
                    int ord = lc;                            // ordinal of the sub-chunk's offset among the tile's non-empty ones
                    if (CH == 128 && p.nseg > 1) { ord = lc / p.nseg; boff[j] = (lc - ord * p.nseg) * 128; }
                    idx[j] = ql_lds_s32(nbr_s + u.buf_off + (uint32_t)ord * (QL_TILE_M * 4u));
                }
            }
            if (p.wide) {                                    // rows are multiples of 32 bytes: 256-bit loads
#pragma unroll
                for (int j = 0; j < kGroup; ++j) {
                    const uint8_t* src = p.feats + (int64_t)(idx[j] < 0 ? 0 : idx[j]) * p.row_bytes + boff[j];
#pragma unroll
                    for (int t = 0; t < CH / 32; ++t) {
                        uint32_t x[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                        if (idx[j] >= 0 && boff[j] + t * 32 < p.row_bytes) ldg32(src + t * 32, x);
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[j * kAReg + 8 * t + e] = x[e];
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < kGroup; ++j) {
                    const uint8_t* src = p.feats + (int64_t)(idx[j] < 0 ? 0 : idx[j]) * p.row_bytes + boff[j];
#pragma unroll
                    for (int t = 0; t < CH / 16; ++t) {
                        uint4 x = make_uint4(0u, 0u, 0u, 0u);
                        if (idx[j] >= 0 && boff[j] + t * 16 < p.row_bytes) x = ldg16(src + t * 16);
                        const int o = j * kAReg + 4 * t;
                        v[o] = x.x; v[o + 1] = x.y; v[o + 2] = x.z; v[o + 3] = x.w;
                    }
                }
            }
        };
        auto store = [&](const Unit& u, const uint32_t (&v)[32]) {
            ql_mbar_wait(ql_smem_u32(&misc->empty[u.ring]), u.ph ^ 1u);         // the MMAs that read this slot have completed
            if (u.n > 0) {
                ql_tc_fence_after();
                tmem_st<32>(a_lane_base + u.ring * (uint32_t)kSlotCols, v);
                tmem_st_wait();
                ql_tc_fence_before();
            }
            __syncwarp();
            if (lane == 0) ql_mbar_arrive(ql_smem_u32(&misc->full[u.ring]));
        };
        // software pipeline, two register buffers: the next unit's loads are in flight while the current unit waits for
        // its ring slot.  A unit that is still held in registers is never kept waiting on a rulebook slab (that could
        // deadlock against the loader, which streams weights only as fast as stored units are consumed): if the next
        // tile's slab is not there yet the held unit is stored first.
        uint32_t va[32], vb[32];
        Unit ua, ub;
        int have_a = next_unit(true, ua);
        if (have_a == 1) issue(ua, va);
        while (have_a == 1) {
            int have_b = next_unit(false, ub);
            if (have_b == 1) issue(ub, vb);
            store(ua, va);
            if (have_b == 2) {
                have_b = next_unit(true, ub);
                if (have_b == 1) issue(ub, vb);
            }
            if (have_b != 1) break;
            have_a = next_unit(false, ua);
            if (have_a == 1) issue(ua, va);
            store(ub, vb);
            if (have_a == 2) {
                have_a = next_unit(true, ua);
                if (have_a == 1) issue(ua, va);
            }
        }
    } else if (warp < kMmaWarp) {
        // ================================ epilogue ================================
        const int w = warp - kEpilogueWarp0;                 // TMEM lane quarter (warp id % 4)
        const int et = tid - kEpilogueWarp0 * 32;            // 0..127 == row in tile
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int a = p.n_acc == 2 ? (int)(it & 1u) : 0;
            const uint32_t aph = p.n_acc == 2 ? ((it >> 1) & 1u) : (it & 1u);
            ql_mbar_wait(ql_smem_u32(&misc->acc_full[a]), aph);
            ql_tc_fence_after();
            const int64_t row = tile * QL_TILE_M + et;
            const bool row_ok = row < n_out;
            const uint32_t taddr = tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(a * p.c_out);
            for (int c0 = 0; c0 < p.c_out; c0 += 16) {
                uint32_t v[16];
                ql_tmem_ld16(taddr + (uint32_t)c0, v);
                ql_tmem_ld_wait();
                if (p.out_dtype == QL_S32) {
                    if (row_ok) {
                        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(p.out) + row * p.c_out + c0);
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) o[qd] = make_uint4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
                    }
                    continue;
                }
                float y[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float acc = kInt8 ? (float)(int)v[j] : __uint_as_float(v[j]);
                    y[j] = fmaf(acc, s_scale[c0 + j], s_shift[c0 + j]);
                }
                if (p.residual && row_ok) {
                    const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + row * p.c_out + c0);
                    uint4 ra = r4[0], rb = r4[1];
                    const __half2* h = reinterpret_cast<const __half2*>(&ra);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 f = __half22float2(h[j]);
                        y[2 * j] += f.x; y[2 * j + 1] += f.y;
                    }
                    h = reinterpret_cast<const __half2*>(&rb);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 f = __half22float2(h[j]);
                        y[8 + 2 * j] += f.x; y[8 + 2 * j + 1] += f.y;
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j], 0.f);
                }
                if (row_ok) {
                    if (p.out_dtype == QL_F16) {
                        uint32_t h[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            __half2 hh = __floats2half2_rn(y[2 * j], y[2 * j + 1]);
                            h[j] = *reinterpret_cast<uint32_t*>(&hh);
                        }
                        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + row * p.c_out + c0);
                        o[0] = make_uint4(h[0], h[1], h[2], h[3]);
                        o[1] = make_uint4(h[4], h[5], h[6], h[7]);
                    } else {
                        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * p.c_out + c0);
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) o[qd] = make_float4(y[4 * qd], y[4 * qd + 1], y[4 * qd + 2], y[4 * qd + 3]);
                    }
                    if (p.out_q) {
                        uint32_t qq[4];
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            uint32_t word = 0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float t = rintf(y[4 * qd + j] * s_qscale[c0 + 4 * qd + j]);
                                t = fminf(fmaxf(t, -127.f), 127.f);
                                word |= ((uint32_t)(uint8_t)(int8_t)(int)t) << (8 * j);
                            }
                            qq[qd] = word;
                        }
                        *reinterpret_cast<uint4*>(p.out_q + row * p.c_out + c0) = make_uint4(qq[0], qq[1], qq[2], qq[3]);
                    }
                }
                if (p.absmax) {
                    // warp-wide max per channel (redux.sync), then one shared-memory atomic per channel
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t m = __reduce_max_sync(0xffffffffu, row_ok ? __float_as_uint(fabsf(y[j])) : 0u);
                        if (lane == j) atomicMax(&s_absmax[c0 + j], m);
                    }
                }
            }
            ql_tc_fence_before();
            ql_mbar_arrive(ql_smem_u32(&misc->acc_empty[a]));
        }
    } else if (warp == kMmaWarp) {
        // =============================== MMA issuer ===============================
        // One elected lane runs the whole loop (nothing in it is warp-collective).  Every sub-chunk of every tile passes
        // through this single instruction stream, so it is kept short: ring position and phase are counters, the tile
        // mask is walked with ffs, descriptors differ only in their low word.
        if (ql_elect_one()) {
            const uint32_t idesc = make_idesc<kInt8>(p.c_out);
            const uint64_t bdesc0 = umma_desc_b<CH>(smem_base_u32);
            const uint32_t bdesc_hi = (uint32_t)(bdesc0 >> 32), bdesc_lo0 = (uint32_t)bdesc0;
            const uint32_t b_sub16 = b_sub_bytes >> 4;
            const uint32_t a_base = tmem_base + (uint32_t)p.a_col0;
            const uint32_t full0 = ql_smem_u32(&misc->full[0]), empty0 = ql_smem_u32(&misc->empty[0]);
            const int nseg = p.nseg;
            const bool narrow = p.mask_words == 1;                 // kernel volume <= 32: the mask is one word
            uint32_t ring = 0, ph = 0, it = 0;
            if (kResident && (int64_t)blockIdx.x < n_tiles) ql_mbar_wait(ql_smem_u32(&misc->w_full), 0);
            uint32_t mask_next[kMaskWords];
            int n_sub_next = (int64_t)blockIdx.x < n_tiles ? load_tile_mask(p, blockIdx.x, mask_next) : 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                uint32_t m0 = mask_next[0];
                uint64_t m_lo = (uint64_t)mask_next[0] | ((uint64_t)mask_next[1] << 32);
                uint64_t m_hi = (uint64_t)mask_next[2] | ((uint64_t)mask_next[3] << 32);
                const int n_sub = n_sub_next;
                if (tile + gridDim.x < n_tiles) n_sub_next = load_tile_mask(p, tile + gridDim.x, mask_next);   // in flight during this tile
                const int a = p.n_acc == 2 ? (int)(it & 1u) : 0;
                const uint32_t aph = p.n_acc == 2 ? ((it >> 1) & 1u) : (it & 1u);
                ql_mbar_wait(ql_smem_u32(&misc->acc_empty[a]), aph ^ 1u);
                ql_tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * p.c_out);
                uint32_t accumulate = 0u;
                int k = 0, seg = nseg;                           // seg == nseg: take the next offset from the mask
                for (int c0 = 0; c0 < n_sub; c0 += kSlotSubs) {
                    ql_mbar_wait(full0 + ring * 8u, ph);
                    ql_tc_fence_after();
                    const int n_in = n_sub - c0 < kSlotSubs ? n_sub - c0 : kSlotSubs;
                    const uint32_t a_slot = a_base + ring * (uint32_t)kSlotCols;
                    const uint32_t b_slot = ring * (uint32_t)kSlotSubs;
                    for (int j = 0; j < n_in; ++j) {
                        uint32_t b_idx = b_slot + (uint32_t)j;       // streamed: the sub-chunk's place in the ring slot
                        if (kResident) {                             // resident: its place in the packed tensor
                            if (seg == nseg) {
                                seg = 0;
                                if (narrow) { k = __ffs((int)m0) - 1; m0 &= m0 - 1u; }
                                else if (m_lo) { k = __ffsll((long long)m_lo) - 1; m_lo &= m_lo - 1; }
                                else { k = 63 + __ffsll((long long)m_hi); m_hi &= m_hi - 1; }
                            }
                            b_idx = (uint32_t)(k * nseg + seg);
                            ++seg;
                        }
                        const uint32_t blo = bdesc_lo0 + b_idx * b_sub16;
                        const uint32_t a_tmem = a_slot + (uint32_t)(j * kAReg);
#pragma unroll
                        for (int ks = 0; ks < CH / 32; ++ks) {       // one k-step = 32 bytes of K: +8 TMEM columns, +2 in the desc (addr >> 4)
                            const uint64_t bdesc = ((uint64_t)bdesc_hi << 32) | (uint64_t)(blo + (uint32_t)(ks * 2));
                            tc_mma_ts<kInt8>(d_tmem, a_tmem + (uint32_t)(ks * 8), bdesc, idesc, accumulate);
                            accumulate = 1u;
                        }
                    }
                    ql_tc_commit(empty0 + ring * 8u);
                    if (c0 + kSlotSubs >= n_sub) ql_tc_commit(ql_smem_u32(&misc->acc_full[a]));
                    if (++ring == (uint32_t)S) { ring = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else {
        // =============================== TMA loader ===============================
        const uint32_t nbr_s0 = smem_base_u32 + (uint32_t)p.off_nbr;
        const uint32_t nbr_stride = (uint32_t)p.nbr_stride, nbmask = (uint32_t)p.nbr_bufs - 1u;
        // Tile `tile` (the CTA's itn-th) -> rulebook buffer itn % nbr_bufs: its mask into the 16-byte header (so producers
        // need no global load of their own), its non-empty slabs packed in mask order behind it (slab j = j-th set bit);
        // lane j copies slabs j, j+32, ...
        auto prefetch_nbr = [&](int64_t tile, uint32_t itn, const uint32_t (&mask)[kMaskWords], int n_slabs) {
            const uint32_t nb = itn & nbmask;
            const uint32_t bar = ql_smem_u32(&misc->nbr_full[nb]);
            const uint32_t dst = nbr_s0 + nb * nbr_stride;
            ql_mbar_wait(ql_smem_u32(&misc->nbr_empty[nb]), ((itn >> p.nbr_log2) & 1u) ^ 1u);
            if (lane < kMaskWords) {
                uint32_t w = mask[0];
#pragma unroll
                for (int i = 1; i < kMaskWords; ++i)
                    if (lane == i) w = mask[i];
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + 4u * lane), "r"(w) : "memory");
            }
            __syncwarp();
            if (lane == 0) ql_mbar_arrive_expect_tx(bar, (uint32_t)n_slabs * (QL_TILE_M * 4u));   // release: orders the header stores
            __syncwarp();
            const int* src = p.nbr + tile * (int64_t)p.kvol * QL_TILE_M;
            for (int j0 = 0; j0 < n_slabs; j0 += 32) {
                const int j = j0 + lane;
                if (j < n_slabs) ql_bulk_g2s(dst + 16u + (uint32_t)j * (QL_TILE_M * 4u), src + nth_set_bit(mask, j) * QL_TILE_M, QL_TILE_M * 4u, bar);
                __syncwarp();
            }
        };
        uint32_t it = 0;
        if (kResident && (int64_t)blockIdx.x < n_tiles) {
            // the whole packed weight tensor, 32 lanes x (w_bytes / 32) bytes
            const uint32_t bar = ql_smem_u32(&misc->w_full);
            const uint32_t per_lane = (uint32_t)p.w_bytes / 32u;
            if (lane == 0) ql_mbar_arrive_expect_tx(bar, (uint32_t)p.w_bytes);
            __syncwarp();
            ql_bulk_g2s(smem_base_u32 + (uint32_t)lane * per_lane, p.w_packed + (size_t)lane * per_lane, per_lane, bar);
            __syncwarp();
        }
        // rulebook prefetch runs nbr_bufs-1 tiles ahead of the tile being streamed; pf = next tile to prefetch, its mask is
        // loaded one step early so the global-load latency is off the path
        int64_t pf_tile = blockIdx.x;
        uint32_t pf_it = 0;
        uint32_t pf_mask[kMaskWords];
        int pf_sub = pf_tile < n_tiles ? load_tile_mask(p, pf_tile, pf_mask) : 0;
        auto prefetch_step = [&]() {
            if (pf_tile >= n_tiles) return;
            uint32_t m[kMaskWords];
#pragma unroll
            for (int i = 0; i < kMaskWords; ++i) m[i] = pf_mask[i];
            const int n_slabs = pf_sub / p.nseg;
            const int64_t t = pf_tile;
            const uint32_t itn = pf_it;
            pf_tile += gridDim.x; ++pf_it;
            if (pf_tile < n_tiles) pf_sub = load_tile_mask(p, pf_tile, pf_mask);
            prefetch_nbr(t, itn, m, n_slabs);
        };
        for (int i = 0; i < p.nbr_bufs - 1; ++i) prefetch_step();
        // streamed weights: one slot per pass, lane j = the slot's j-th sub-chunk
        uint32_t ring = 0, ph = 0;
        uint32_t mask_next[kMaskWords];
        int n_sub_next = (!kResident && (int64_t)blockIdx.x < n_tiles) ? load_tile_mask(p, blockIdx.x, mask_next) : 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            prefetch_step();
            if constexpr (!kResident) {
            uint32_t mask[kMaskWords];
#pragma unroll
            for (int i = 0; i < kMaskWords; ++i) mask[i] = mask_next[i];
            const int n_sub = n_sub_next;
            if (tile + gridDim.x < n_tiles) n_sub_next = load_tile_mask(p, tile + gridDim.x, mask_next);
            for (int c0 = 0; c0 < n_sub; c0 += kSlotSubs) {
                const int sub = c0 + lane;
                const bool mine = lane < kSlotSubs && sub < n_sub;
                ql_mbar_wait(ql_smem_u32(&misc->empty[ring]), ph ^ 1u);
                if (lane == 0) {
                    const int n_in = n_sub - c0 < kSlotSubs ? n_sub - c0 : kSlotSubs;
                    ql_mbar_arrive_expect_tx(ql_smem_u32(&misc->full[ring]), (uint32_t)n_in * b_sub_bytes);
                }
                __syncwarp();
                if (mine) {
                    const int ord = sub / p.nseg, seg = sub - ord * p.nseg;
                    const int k = nth_set_bit(mask, ord);
                    ql_bulk_g2s(smem_base_u32 + (ring * (uint32_t)kSlotSubs + (uint32_t)lane) * b_sub_bytes,
                                p.w_packed + (int64_t)(k * p.nseg + seg) * b_sub_bytes, b_sub_bytes, ql_smem_u32(&misc->full[ring]));
                }
                __syncwarp();
                if (++ring == (uint32_t)S) { ring = 0; ph ^= 1u; }
            }
            }
        }
    }

    ql_tc_fence_before();
    __syncthreads();
    ql_tc_fence_after();
    if (p.absmax) {
        for (int c = tid; c < p.c_out; c += kThreadsTotal) {
            uint32_t v = s_absmax[c];
            if (v) atomicMax(reinterpret_cast<unsigned int*>(p.absmax) + c, v);
        }
    }
    if (warp == kMmaWarp) ql_tmem_dealloc(tmem_base, kTmemCols);
}

inline int elem_size(int dtype) { return dtype == QL_S8 ? 1 : (dtype == QL_F16 ? 2 : 0); }

// chunk geometry shared by the packer and the launcher
struct ChunkGeom {
    int ch;      // bytes of K per chunk (32 / 64 / 128), zero padded when the row (segment) is shorter
    int nseg;    // chunks per kernel offset
};
inline ChunkGeom chunk_geom(int row_bytes) {
    ChunkGeom g;
    if (row_bytes > 128) { g.ch = 128; g.nseg = (row_bytes + 127) / 128; }
    else { g.ch = row_bytes <= 32 ? 32 : (row_bytes <= 64 ? 64 : 128); g.nseg = 1; }
    return g;
}
// byte offset of 16-byte piece c16 of row r inside a K-major swizzled [rows x ch bytes] chunk image
inline uint32_t chunk_sw_offset(int ch, uint32_t r, uint32_t c16) {
    const uint32_t x = ch == 128 ? (r & 7u) : (ch == 64 ? ((r >> 1) & 3u) : ((r >> 2) & 1u));
    return (r >> 3) * (uint32_t)(8 * ch) + (r & 7u) * (uint32_t)ch + ((c16 ^ x) << 4);
}

template <bool kInt8, int CH, bool kResident>
cudaError_t launch2(const ConvParams& p, int grid, size_t smem_bytes, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_spconv_ts<kInt8, CH, kResident>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    k_spconv_ts<kInt8, CH, kResident><<<grid, kThreadsTotal, smem_bytes, st>>>(p);
    return cudaPeekAtLastError();                     // left pending for ql_last_cuda_error()
}
template <bool kInt8, int CH>
cudaError_t launch(const ConvParams& p, int grid, size_t smem_bytes, cudaStream_t st) {
    return p.resident ? launch2<kInt8, CH, true>(p, grid, smem_bytes, st) : launch2<kInt8, CH, false>(p, grid, smem_bytes, st);
}

}  // namespace

extern "C" size_t ql_packed_weight_bytes(int32_t c_in, int32_t c_out, int32_t kvol, int32_t elem_dtype) {
    int es = elem_size(elem_dtype);
    if (es == 0 || c_in <= 0 || c_out <= 0 || kvol <= 0) return 0;
    const ChunkGeom g = chunk_geom(c_in * es);
    return (size_t)kvol * g.nseg * (size_t)c_out * g.ch;
}

// w_host: [c_out][kvol][c_in] elements (== the reference layout (oc, kd, kh, kw, ic) flattened, quant/quant.py:37-39).
// packed: for every (offset k, segment s) chunk one [c_out x CH bytes] K-major swizzled image (SWIZZLE_32B/64B/128B by
// CH), zero padded -- exactly what the loader warp bulk-copies into the chunk's shared-memory slot.
extern "C" int ql_pack_weights_host(const void* w_host, int32_t elem_dtype, int32_t c_in, int32_t c_out, int32_t kvol,
                                    void* packed_host) {
    int es = elem_size(elem_dtype);
    if (!w_host || !packed_host || es == 0 || c_in <= 0 || c_out <= 0 || kvol <= 0) return QL_ERR_INVALID;
    if ((c_in * es) % 16 != 0 || c_out % 16 != 0 || c_out > 256) return QL_ERR_UNSUPPORTED;
    const int row_bytes = c_in * es;
    const ChunkGeom g = chunk_geom(row_bytes);
    const size_t chunk_bytes = (size_t)c_out * g.ch;
    memset(packed_host, 0, ql_packed_weight_bytes(c_in, c_out, kvol, elem_dtype));
    const uint8_t* src = (const uint8_t*)w_host;
    uint8_t* dst = (uint8_t*)packed_host;
    for (int oc = 0; oc < c_out; ++oc)
        for (int k = 0; k < kvol; ++k)
            for (int b = 0; b < row_bytes; b += 16) {
                const int seg = b / 128, c16 = (b % 128) / 16;
                memcpy(dst + (size_t)(k * g.nseg + seg) * chunk_bytes + chunk_sw_offset(g.ch, (uint32_t)oc, (uint32_t)c16),
                       src + ((size_t)oc * kvol + k) * row_bytes + b, 16);
            }
    return QL_OK;
}

extern "C" int ql_spconv_mma(const void* feats, int32_t in_dtype, const int32_t* nbr, const uint32_t* tile_kmask, int64_t n_out_cap,
                             const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t kvol, const void* w_packed,
                             const float* scale, const float* shift, const float* act_scale_dev, const void* residual_f16,
                             int32_t relu, void* out, int32_t out_dtype, int8_t* out_q, const float* out_qscale,
                             float* absmax, ql_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (!feats || !nbr || !w_packed || !scale || !shift || !out) return QL_ERR_INVALID;
    int es = elem_size(in_dtype);
    if (es == 0) return QL_ERR_INVALID;
    if (out_dtype != QL_F16 && out_dtype != QL_F32 && out_dtype != QL_S32) return QL_ERR_INVALID;
    if (out_q && !out_qscale) return QL_ERR_INVALID;
    if (c_in <= 0 || (c_in * es) % 16 != 0 || c_out < 16 || c_out % 16 != 0 || c_out > 256 || kvol <= 0 || kvol > 32 * kMaskWords)
        return QL_ERR_UNSUPPORTED;
    if (n_out_cap <= 0) return QL_OK;

    ConvParams p;
    p.feats = (const uint8_t*)feats; p.nbr = nbr; p.kmask = tile_kmask; p.n_out_dev = n_out_dev; p.n_out_cap = n_out_cap;
    p.row_bytes = c_in * es; p.c_out = c_out; p.kvol = kvol;
    p.wide = (p.row_bytes % 32 == 0 && ((uintptr_t)feats & 31) == 0) ? 1 : 0;
    const ChunkGeom g = chunk_geom(p.row_bytes);
    p.nseg = g.nseg; p.mask_words = (kvol + 31) / 32;
    p.w_packed = (const uint8_t*)w_packed; p.scale = scale; p.shift = shift; p.act_scale_dev = act_scale_dev;
    p.residual = (const __half*)residual_f16; p.relu = relu; p.out = out; p.out_dtype = out_dtype;
    p.out_q = out_q; p.out_qscale = out_qscale; p.absmax = absmax;

    // ring depth in slots (kTeams units of 32 TMEM columns): bounded by the TMEM columns left beside the accumulators
    // and, when the weights are streamed, by shared memory
    const int misc_bytes = (int)sizeof(MiscSmem) + 4 * c_out * 4;
    const int b_sub = c_out * g.ch;
    const int b_slot = kTeams * c_out * 128;                     // a slot's weight sub-chunks
    p.nbr_stride = (16 + kvol * QL_TILE_M * 4 + 127) & ~127;
    p.n_acc = (kTmemCols - 2 * c_out) / kSlotCols >= 2 ? 2 : 1;
    int S = (kTmemCols - p.n_acc * c_out) / kSlotCols;
    if (S > kMaxSlots) S = kMaxSlots;
    p.w_bytes = kvol * g.nseg * b_sub;
    for (p.nbr_bufs = 4; p.nbr_bufs >= 2; p.nbr_bufs >>= 1) {
        if (p.nbr_bufs == 4 && 4 * p.nbr_stride > 64 * 1024) continue;
        const int smem_free = kSmemBudget - 1024 - p.nbr_bufs * p.nbr_stride - ((misc_bytes + 127) & ~127);
        p.resident = (p.w_bytes <= smem_free && p.w_bytes % 512 == 0) ? 1 : 0;   // 32 lanes x 16-byte multiples
        if (p.resident || smem_free / b_slot >= 2) {
            if (!p.resident && S > smem_free / b_slot) S = smem_free / b_slot;
            break;
        }
    }
    if (p.nbr_bufs < 2 || S < 2) return QL_ERR_UNSUPPORTED;
    p.nbr_log2 = p.nbr_bufs == 4 ? 2 : 1;
    const int nbr_bytes = p.nbr_bufs * p.nbr_stride;
    p.n_slots = S;
    p.a_col0 = p.n_acc * c_out;
    p.off_nbr = p.resident ? ((p.w_bytes + 1023) & ~1023) : S * b_slot;
    p.off_misc = (p.off_nbr + nbr_bytes + 127) & ~127;
    size_t smem_bytes = 1024 + (size_t)p.off_misc + misc_bytes;
    if (smem_bytes < (size_t)kSmemFloor) smem_bytes = kSmemFloor;

    int64_t tiles = (n_out_cap + QL_TILE_M - 1) / QL_TILE_M;
    int grid = (int)(tiles < ql_num_sms() ? tiles : ql_num_sms());
    cudaError_t e;
    if (in_dtype == QL_S8) {
        e = g.ch == 32 ? launch<true, 32>(p, grid, smem_bytes, st)
          : g.ch == 64 ? launch<true, 64>(p, grid, smem_bytes, st) : launch<true, 128>(p, grid, smem_bytes, st);
    } else {
        e = g.ch == 32 ? launch<false, 32>(p, grid, smem_bytes, st)
          : g.ch == 64 ? launch<false, 64>(p, grid, smem_bytes, st) : launch<false, 128>(p, grid, smem_bytes, st);
    }
    return e == cudaSuccess ? QL_OK : QL_ERR_CUDA;
}
