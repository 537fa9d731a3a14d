// Implicit gather-GEMM-scatter sparse convolution on tcgen05 tensor cores (sm_100a), fused epilogue.
//
// Replaces QConvNd.forward -> [EXT] spconv SubMConv3d/SparseConv3d forward (quant/quant.py:36-58) together with
// the BatchNorm1d / ReLU / residual-add that follow it (spconv_backbone.py:8-27,51-67).
//
// One persistent CTA per SM walks 128-row output tiles.  Per tile the conv is a sum of per-offset GEMMs
//     D[128, C_out] += A_k[128, C_in] . W_k[C_out, C_in]^T      for every kernel offset k that is non-empty in the tile
// (the rulebook's per-tile offset mask lists them; empty (tile, offset) slabs cost nothing).  The A operand never
// exists in memory and never touches shared memory: output row r of the tile is TMEM lane r; gather producers load the
// neighbours' feature rows straight from global memory (L1/L2) into registers and write them to tensor memory with
// tcgen05.st; the MMA reads A from TMEM (the ".ts" operand form) and only the weights from shared memory.
//
// K is cut into sub-chunks: one sub-chunk = one kernel offset x one <=128-byte segment of the input row
// (CH = 32 / 64 / 128 bytes => 1 / 2 / 4 MMA k-steps).  128/CH consecutive sub-chunks form a UNIT = 32 TMEM columns x 128
// lanes (16 KB of A).  Units flow through a ring of up to 8 TMEM slots, each with its own full/empty mbarrier pair, so
// every producer warp works independently with exactly one unit in flight.  Roles (22 warps):
//   warps 0-15  gather producers : 4 teams x 4 warps; warp w owns TMEM lanes 32*(w%4)..+31; team t fills units t, t+4, ...
//                                  (global unit numbering across the CTA's tiles).  ncu on the first version of this kernel
//                                  (12 fat producer warps, two units in flight each, a slot state machine) showed the
//                                  producers instruction-latency bound at ~216 warp-instructions per unit and, before
//                                  that, the L1 data pipe at 79-86 % with one 32-byte sector per wavefront
//                                  (profiles/r01_conv_v6_*, r01_conv_v8_*).  Hence: thin warps, straight-line unit code,
//                                  loads that are never predicated (a missing neighbour reads a zero line instead), and
//                                  for rows >= 64 bytes a QUAD gather -- four lanes read one row's CH contiguous bytes
//                                  (8 rows / 8 wavefronts per load instruction) and the registers go to TMEM with
//                                  tcgen05.st.16x256b; the K order this leaves inside a sub-chunk is undone in the weight
//                                  packing (k_word_src).
//   warps 16-19 epilogue         : tcgen05.ld accumulators -> dequant*BN scale/shift (+residual) (+ReLU)
//                                  -> fp16/fp32 rows (+ int8 re-quantised rows, + per-channel absmax)
//   warp  20    MMA issuer       : one lane issues tcgen05.mma (kind::f16 or kind::i8), accumulators in TMEM
//                                  (double buffered when they fit: tile i+1 accumulates while tile i drains)
//   warp  21    loader           : cp.async.bulk of each tile's rulebook block ([kvol][128] int32, one copy per 16 KB, plus a
//                                  header with the tile mask and an ordinal -> offset table) up to 3 tiles ahead, and -- when
//                                  the packed weights do not fit in shared memory (C <= 32 fp16, C <= 64 int8 do) -- of every
//                                  unit's weight sub-chunks into the unit's B slot.  One UBLKCP instruction costs its issuing
//                                  warp ~225 ns whatever the size (tools/microbench/bulk_copy_rate.cu), an extra active lane
//                                  ~26 ns: weight copies for up to half the ring are issued by one instruction, one per lane.
#include "ql_common.cuh"
#include <string.h>

namespace {

constexpr int kTeams = 4;
constexpr int kProducerWarps = kTeams * 4;               // 16
constexpr int kEpilogueThreads = 128;
constexpr int kEpilogueWarp0 = kProducerWarps;            // warps 16..19 (warp % 4 == TMEM lane quarter)
constexpr int kMmaWarp = kEpilogueWarp0 + 4;              // 20
constexpr int kLoaderWarp = kMmaWarp + 1;                 // 21
constexpr int kThreadsTotal = (kLoaderWarp + 1) * 32;     // 704
constexpr int kMaxUnits = 8;                              // ring depth in units
constexpr int kUnitCols = 32;                             // TMEM columns per unit (128 bytes of K per lane)
constexpr int kMaskWords = 4;                             // kernel volumes up to 128 (3^3, 5^3)
constexpr int kTmemCols = 512;
constexpr int kSmemBudget = 232448;                       // 227 KB opt-in maximum per CTA
constexpr int kSmemFloor = 120 * 1024;                    // always ask for > half an SM: one CTA (one TMEM owner) per SM
constexpr int kNbrHeaderMin = 160;                        // rulebook buffer header: 16 B mask, n_off at +16, n_sub at +20, ord -> k
                                                          // table (u8) at +32, then (resident weights) one B-descriptor low
                                                          // word per sub-chunk at +160

__device__ __align__(128) uint8_t g_zero_line[128];       // what a missing neighbour reads (zero-initialised module memory)

struct ConvParams {
    const uint8_t* feats;
    const int* nbr;
    const uint32_t* kmask;  // [tiles][mask_words] or null (every offset)
    const int* row_perm;    // [tiles][128] tile slot -> output row (-1 = padding) of a GROUPED rulebook, or null (slot == row)
    const int* n_out_dev;
    int64_t n_out_cap;
    int row_bytes;          // c_in * elem size
    int wide;               // rows (and the feature base) are 32-byte aligned: gather with 256-bit loads
    int c_out, kvol, nseg, mask_words;
    uint32_t inv_nseg;      // ceil(65536 / nseg): ord = (sub * inv_nseg) >> 16 for sub < 4096
    int pair;               // 16-byte rows (int8, C_in = 16): a sub-chunk holds TWO kernel offsets, 2j in bytes 0-15 and 2j+1 in 16-31
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
    int n_ring;             // A/B ring depth in units
    int teams;              // active producer teams (<= n_ring)
    int n_acc;              // accumulator buffers in TMEM (2, or 1 when 2*c_out does not fit beside the A ring)
    int a_col0;             // first TMEM column of the A ring
    int resident;           // 1: every weight chunk lives in shared memory for the whole kernel (no per-unit B copies)
    int w_bytes;            // packed weight bytes (resident mode)
    int off_nbr;            // smem offset of the rulebook buffers: nbr_bufs x {header, [kvol][128] int32}
    int nbr_bufs, nbr_log2; // 4 (or 2 when shared memory is short): the loader runs nbr_bufs-1 tiles ahead
    int nbr_stride;         // bytes per buffer
    int nbr_hdr;            // header bytes in front of the [kvol][128] block
    int off_misc;           // smem offset of MiscSmem from the 1024-aligned base
};

struct MiscSmem {
    uint64_t full[kMaxUnits];
    uint64_t empty[kMaxUnits];
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

// 256-bit load (LDG.E.ENL2.256, sm_100).  Needs 32-byte alignment.
__device__ __forceinline__ void ldg32(const uint8_t* p, uint32_t* v) {
    asm volatile("ld.global.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void ldg16(const uint8_t* p, uint32_t* v) {
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(p));
}
__device__ __forceinline__ int lds_u8(uint32_t addr) {
    int v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, int v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// 32 lanes x 32 columns: thread t supplies columns 0..31 of TMEM lane base + t
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
}
// 16 lanes x (NREP x 256 bits): thread (t0 = lane%4, t1 = lane/4) supplies, for repeat v2 and half v1, the two 32-bit
// columns 8*v2 + 2*t0 + {0,1} of TMEM lane base + t1 + 8*v1, as registers [4*v2 + 2*v1 + {0,1}].
template <int NREP>
__device__ __forceinline__ void tmem_st_16x256b(uint32_t taddr, const uint32_t* v);
template <>
__device__ __forceinline__ void tmem_st_16x256b<2>(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
template <>
__device__ __forceinline__ void tmem_st_16x256b<4>(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
            taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// the two halves of load_tile_mask: issue the loads now, use them (count / fix-up) later -- a warp issues in order, so a popcount
// right behind its load costs the loader one global round trip per tile
__device__ __forceinline__ void load_tile_mask_raw(const ConvParams& p, int64_t tile, uint32_t (&mask)[kMaskWords]) {
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
    }
}
__device__ __forceinline__ int finish_tile_mask(uint32_t (&mask)[kMaskWords]) {
    int n = 0;
#pragma unroll
    for (int i = 0; i < kMaskWords; ++i) n += __popc(mask[i]);
    if (n == 0) { mask[0] = 1u; n = 1; }       // a tile without pairs still has to zero its accumulators
    return n;
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
    return n;                                  // non-empty kernel offsets
}

template <bool kInt8, int CH, bool kResident>
__global__ void __launch_bounds__(kThreadsTotal, 1) k_spconv_ts(const ConvParams p) {
    constexpr int kAReg = CH / 4;                          // 32-bit TMEM columns (registers) per sub-chunk
    constexpr int kGroup = 128 / CH;                       // sub-chunks per unit
    constexpr int kGroupLog2 = CH == 128 ? 0 : (CH == 64 ? 1 : 2);
    constexpr bool kQuad = CH >= 64;                       // 4 lanes per row + tcgen05.st.16x256b, else lane per row + 32x32b
    constexpr int kRep = kQuad ? CH / 32 : 2;              // 256-bit repeats per sub-chunk in the quad form
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
    const uint32_t R = (uint32_t)p.n_ring;

    const int64_t n_out = p.n_out_dev ? (int64_t)*p.n_out_dev : p.n_out_cap;
    const int64_t n_tiles = (n_out + QL_TILE_M - 1) / QL_TILE_M;

    if (tid == 0) {
        for (int s = 0; s < kMaxUnits; ++s) {
            ql_mbar_init(ql_smem_u32(&misc->full[s]), 4 + (kResident ? 0 : 1));   // the team's 4 warps (+ the loader's expect_tx)
            ql_mbar_init(ql_smem_u32(&misc->empty[s]), 1);                         // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            ql_mbar_init(ql_smem_u32(&misc->acc_full[i]), 1);
            ql_mbar_init(ql_smem_u32(&misc->acc_empty[i]), kEpilogueThreads);
        }
        for (int i = 0; i < 4; ++i) {
            ql_mbar_init(ql_smem_u32(&misc->nbr_full[i]), 1);
            ql_mbar_init(ql_smem_u32(&misc->nbr_empty[i]), 4 * p.teams + 1);   // producer warps + the MMA issuer
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
    const uint32_t nbr_s0 = smem_base_u32 + (uint32_t)p.off_nbr;
    const uint32_t nbr_stride = (uint32_t)p.nbr_stride, nbmask = (uint32_t)p.nbr_bufs - 1u;
    const uint32_t full0 = ql_smem_u32(&misc->full[0]), empty0 = ql_smem_u32(&misc->empty[0]);

    if (warp < kProducerWarps) {
        // ============================ gather producers ============================
        const int q = warp & 3;                              // TMEM lane quarter
        const uint32_t team = (uint32_t)(warp >> 2);
        const uint32_t T = (uint32_t)p.teams;
        if (team < T) {
            const uint32_t a_lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)p.a_col0;
            const int t0 = lane & 3, t1 = lane >> 2;         // quad form: lane t0 of the quad that serves rows t1 + 8*rr
            // byte offset of this thread's first row inside a [128] int32 slab
            const uint32_t row_off = (uint32_t)(q * 32 + (kQuad ? t1 : lane)) * 4u;
            const int tb0 = kQuad ? t0 * (CH / 4) : 0;       // this thread's bytes inside a sub-chunk's row segment
            const uint32_t row_bytes = (uint32_t)p.row_bytes;
            const uint8_t* const feats = p.feats;
            const uint8_t* const zero = g_zero_line;
            const bool wide = p.wide != 0;
            const uint32_t nseg = (uint32_t)p.nseg, inv_nseg = p.inv_nseg, hdr = (uint32_t)p.nbr_hdr, kvol = (uint32_t)p.kvol;
            const bool pair = !kQuad && p.pair != 0;

            uint32_t g = team;                               // next global unit of this team
            uint32_t G0 = 0;                                 // global number of the current tile's first unit
            uint32_t u = team, ph = 0;                       // ring slot / pass parity of unit g
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t nb = it & nbmask;
                const uint32_t buf = nbr_s0 + nb * nbr_stride;
                ql_mbar_wait(ql_smem_u32(&misc->nbr_full[nb]), (it >> p.nbr_log2) & 1u);
                const uint32_t n_sub = (uint32_t)ql_lds_s32(buf + 20u);
                const uint32_t Gend = G0 + ((n_sub + kGroup - 1) >> kGroupLog2);
                // The rulebook indices of a unit (table byte -> slab -> row index: two dependent shared-memory loads) are fetched
                // one unit ahead, right after the previous unit's gathers have been issued, so that chain hides under them.
                constexpr int kIdx = kQuad ? 4 * kGroup : kGroup;
                int idx[kIdx];
                int idx2[kQuad ? 1 : kGroup];                                         // pair mode: the row of kernel offset 2j+1
                uint32_t boff[kGroup];
                auto fetch_idx = [&](uint32_t c0) {
                    // the unit's kGroup table bytes (ordinal -> kernel offset) in one load
                    uint32_t tbl;
                    if constexpr (kGroup == 4) tbl = (uint32_t)ql_lds_s32(buf + 32u + c0);
                    else if constexpr (kGroup == 2) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(tbl) : "r"(buf + 32u + c0));
                    else {
                        uint32_t ord = c0;
                        boff[0] = 0;
                        if (nseg > 1) { ord = (c0 * inv_nseg) >> 16; boff[0] = (c0 - ord * nseg) * 128u; }
                        tbl = (uint32_t)lds_u8(buf + 32u + ord);
                    }
#pragma unroll
                    for (int j = 0; j < kGroup; ++j) {
                        const bool live = c0 + (uint32_t)j < n_sub;
                        const uint32_t k = live ? ((tbl >> (8 * j)) & 0xFFu) : 0u;
                        const uint32_t a = buf + hdr + k * (QL_TILE_M * 4u) + row_off;
                        if constexpr (kGroup > 1) boff[j] = 0;
                        if constexpr (kQuad) {
#pragma unroll
                            for (int rr = 0; rr < 4; ++rr) {
                                const int v = ql_lds_s32(a + (uint32_t)rr * 32u);
                                idx[4 * j + rr] = live ? v : -1;
                            }
                        } else if (!pair) {
                            const int v = ql_lds_s32(a);
                            idx[j] = live ? v : -1;
                        } else {
                            // k is a PAIR of kernel offsets (2k, 2k+1): two slabs, two rows
                            const uint32_t a2 = buf + hdr + (2u * k) * (QL_TILE_M * 4u) + row_off;
                            const int v0 = ql_lds_s32(a2);
                            const int v1 = (2u * k + 1u < kvol) ? ql_lds_s32(a2 + QL_TILE_M * 4u) : -1;
                            idx[j] = live ? v0 : -1;
                            idx2[j] = live ? v1 : -1;
                        }
                    }
                };
                if (g < Gend) fetch_idx((g - G0) << kGroupLog2);
                for (; g < Gend; g += T) {
                    uint32_t v[32];
                    if constexpr (kQuad) {
#pragma unroll
                        for (int j = 0; j < kGroup; ++j) {
                            const uint32_t tb = boff[j] + (uint32_t)tb0;
#pragma unroll
                            for (int rr = 0; rr < 4; ++rr) {
                                const int h = rr >> 1, v1 = rr & 1;
                                const int id = idx[4 * j + rr];
                                const bool ok = id >= 0 && tb < row_bytes;
                                const uint8_t* src = ok ? feats + ((uint64_t)(uint32_t)id * row_bytes + tb) : zero;
                                uint32_t x[8];
                                if constexpr (CH == 128) {
                                    if (wide) {
                                        ldg32(src, x);
                                    } else {
                                        ldg16(src, x);
                                        ldg16((ok && tb + 16u < row_bytes) ? src + 16 : zero, x + 4);
                                    }
                                } else {
                                    ldg16(src, x);
                                }
#pragma unroll
                                for (int v2 = 0; v2 < kRep; ++v2) {
                                    v[j * kAReg + h * (4 * kRep) + 4 * v2 + 2 * v1] = x[2 * v2];
                                    v[j * kAReg + h * (4 * kRep) + 4 * v2 + 2 * v1 + 1] = x[2 * v2 + 1];
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < kGroup; ++j) {
                            const int id = idx[j];                                       // CH = 32: one sub-chunk per kernel offset
                            const bool ok = id >= 0;
                            const uint8_t* src = ok ? feats + (uint64_t)(uint32_t)id * row_bytes : zero;
                            if (wide) {
                                ldg32(src, v + j * 8);
                            } else if (!pair) {
                                ldg16(src, v + j * 8);
                                ldg16((ok && 16u < row_bytes) ? src + 16 : zero, v + j * 8 + 4);
                            } else {
                                ldg16(src, v + j * 8);
                                ldg16(idx2[j] >= 0 ? feats + (uint64_t)(uint32_t)idx2[j] * row_bytes : zero, v + j * 8 + 4);
                            }
                        }
                    }
                    if (g + T < Gend) fetch_idx((g + T - G0) << kGroupLog2);             // next unit's indices, under this unit's gathers
                    ql_mbar_wait(empty0 + u * 8u, ph ^ 1u);              // the MMAs that read this ring slot have completed
                    ql_tc_fence_after();
                    const uint32_t a_unit = a_lane_base + u * (uint32_t)kUnitCols;
                    if constexpr (kQuad) {
#pragma unroll
                        for (int j = 0; j < kGroup; ++j)
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                tmem_st_16x256b<kRep>(a_unit + ((uint32_t)(h * 16) << 16) + (uint32_t)(j * kAReg), &v[j * kAReg + h * (4 * kRep)]);
                    } else {
                        tmem_st_32x32b_x32(a_unit, v);
                    }
                    tmem_st_wait();
                    ql_tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ql_mbar_arrive(full0 + u * 8u);
                    u += T;
                    if (u >= R) { u -= R; ph ^= 1u; }
                }
                G0 = Gend;
                __syncwarp();
                if (lane == 0) ql_mbar_arrive(ql_smem_u32(&misc->nbr_empty[nb]));   // this warp has read all it needs from the block
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
            // grouped rulebook: the tile's rows are scattered over the output; the slot -> row load hides under the wait
            const int64_t slot = tile * QL_TILE_M + et;
            int64_t row = slot;
            if (p.row_perm) row = slot < n_out ? (int64_t)__ldg(p.row_perm + slot) : -1;
            const bool row_ok = row >= 0 && row < n_out;
            // The residual row does not depend on the MMAs: (kind::f16 instantiations) its first 32 bytes are requested BEFORE the
            // accumulator wait and every further chunk one iteration ahead, so the global-load latency is off the epilogue's critical
            // path (the role trace, profiles/r01_conv_role_waits.md, showed the epilogue 80-90 % busy on every residual layer and the
            // MMA thread waiting for accumulators up to 26 % of its time).  The kind::i8 instantiations keep the load inside the
            // column loop: with the early loads their abs-max epilogue (dynamic W8A8) ran 25 % slower on every layer.
            const bool has_res = p.residual != nullptr && row_ok;
            const uint4* res4 = has_res ? reinterpret_cast<const uint4*>(p.residual + row * p.c_out) : nullptr;
            uint4 ra = make_uint4(0u, 0u, 0u, 0u), rb = ra;
            if constexpr (!kInt8) {
                if (has_res) { ra = __ldg(res4); rb = __ldg(res4 + 1); }
            }
            ql_mbar_wait(ql_smem_u32(&misc->acc_full[a]), aph);
            ql_tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(a * p.c_out);
            for (int c0 = 0; c0 < p.c_out; c0 += 16) {
                uint32_t v[16];
                ql_tmem_ld16(taddr + (uint32_t)c0, v);
                uint4 rc = ra, rd = rb;                                // this iteration's residual chunk
                if constexpr (!kInt8) {                                // ... and the next one requested now
                    if (has_res && c0 + 16 < p.c_out) { ra = __ldg(res4 + (c0 + 16) / 8); rb = __ldg(res4 + (c0 + 16) / 8 + 1); }
                }
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
                if (has_res) {
                    if constexpr (kInt8) { rc = res4[c0 / 8]; rd = res4[c0 / 8 + 1]; }
                    const __half2* h = reinterpret_cast<const __half2*>(&rc);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 f = __half22float2(h[j]);
                        y[2 * j] += f.x; y[2 * j + 1] += f.y;
                    }
                    h = reinterpret_cast<const __half2*>(&rd);
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
            const uint32_t n_acc = (uint32_t)p.n_acc, c_out = (uint32_t)p.c_out;
            uint32_t u = 0, ph = 0, it = 0;
            if (kResident && (int64_t)blockIdx.x < n_tiles) ql_mbar_wait(ql_smem_u32(&misc->w_full), 0);
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                // the tile's header in the rulebook buffer: n_sub and (resident weights) the B descriptor of every sub-chunk,
                // prepared by the loader's 32 lanes so that this one thread only loads and issues
                const uint32_t nb = it & nbmask;
                const uint32_t buf = nbr_s0 + nb * nbr_stride;
                ql_mbar_wait(ql_smem_u32(&misc->nbr_full[nb]), (it >> p.nbr_log2) & 1u);
                const uint32_t n_sub = (uint32_t)ql_lds_s32(buf + 20u);
                const uint32_t a = n_acc == 2 ? (it & 1u) : 0u;
                const uint32_t aph = n_acc == 2 ? ((it >> 1) & 1u) : (it & 1u);
                ql_mbar_wait(ql_smem_u32(&misc->acc_empty[a]), aph ^ 1u);
                ql_tc_fence_after();
                const uint32_t d_tmem = tmem_base + a * c_out;
                const uint32_t acc_bar = ql_smem_u32(&misc->acc_full[a]);
                uint32_t accumulate = 0u;
                for (uint32_t c0 = 0; c0 < n_sub; c0 += kGroup) {
                    uint32_t blo[kGroup];
                    if constexpr (kResident) {
                        const uint32_t ta = buf + (uint32_t)kNbrHeaderMin + 4u * c0;
                        if constexpr (kGroup == 4) {
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(blo[0]), "=r"(blo[1]), "=r"(blo[2]), "=r"(blo[3]) : "r"(ta));
                        } else if constexpr (kGroup == 2) {
                            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(blo[0]), "=r"(blo[1]) : "r"(ta));
                        } else {
                            blo[0] = (uint32_t)ql_lds_s32(ta);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < kGroup; ++j) blo[j] = bdesc_lo0 + (u * (uint32_t)kGroup + (uint32_t)j) * b_sub16;
                    }
                    ql_mbar_wait(full0 + u * 8u, ph);
                    ql_tc_fence_after();
                    const uint32_t a_unit = a_base + u * (uint32_t)kUnitCols;
#pragma unroll
                    for (int j = 0; j < kGroup; ++j) {
                        if (j == 0 || c0 + (uint32_t)j < n_sub) {
#pragma unroll
                            for (int ks = 0; ks < CH / 32; ++ks) {       // one k-step = 32 bytes of K: +8 TMEM columns, +2 in the desc (addr >> 4)
                                const uint64_t bdesc = ((uint64_t)bdesc_hi << 32) | (uint64_t)(blo[j] + (uint32_t)(ks * 2));
                                tc_mma_ts<kInt8>(d_tmem, a_unit + (uint32_t)(j * kAReg + ks * 8), bdesc, idesc, accumulate);
                                accumulate = 1u;
                            }
                        }
                    }
                    ql_tc_commit(empty0 + u * 8u);
                    if (c0 + kGroup >= n_sub) ql_tc_commit(acc_bar);
                    if (++u == R) { u = 0; ph ^= 1u; }
                }
                ql_mbar_arrive(ql_smem_u32(&misc->nbr_empty[nb]));
            }
        }
        __syncwarp();
    } else {
        // ================================= loader =================================
        // Tile `tile` (the CTA's itn-th) -> rulebook buffer itn % nbr_bufs: header {mask, n_off, ord -> k table} written with
        // plain stores (released by the arrive below), then the tile's whole [kvol][128] block, one bulk copy per 16 KB.
        auto prefetch_nbr = [&](int64_t tile, uint32_t itn, const uint32_t (&mask)[kMaskWords], int n_off) {   // n_off: live offsets (pairs in pair mode)
            const uint32_t nb = itn & nbmask;
            const uint32_t bar = ql_smem_u32(&misc->nbr_full[nb]);
            const uint32_t dst = nbr_s0 + nb * nbr_stride;
            ql_mbar_wait(ql_smem_u32(&misc->nbr_empty[nb]), ((itn >> p.nbr_log2) & 1u) ^ 1u);
            uint32_t vmask[kMaskWords];
#pragma unroll
            for (int i = 0; i < kMaskWords; ++i) vmask[i] = mask[i];
            if (p.pair) {
                // pair j is live when kernel offset 2j or 2j+1 is: lane l of pass i looks at pair 32*i + l
                n_off = 0;
#pragma unroll
                for (int i = 0; i < kMaskWords; ++i) {
                    const int vj = 32 * i + lane;                      // pair index; its two bits never straddle a word
                    uint32_t word = 0u;
                    if (2 * vj < 32 * kMaskWords) {
                        word = mask[0];
#pragma unroll
                        for (int q2 = 1; q2 < kMaskWords; ++q2)
                            if (((2 * vj) >> 5) == q2) word = mask[q2];
                    }
                    const bool on = 2 * vj < 32 * kMaskWords && ((word >> ((2 * vj) & 31)) & 3u) != 0u;
                    vmask[i] = __ballot_sync(0xffffffffu, on);
                    n_off += __popc(vmask[i]);
                }
            }
            int prefix = 0;
#pragma unroll
            for (int i = 0; i < kMaskWords; ++i) {
                const uint32_t w = vmask[i];
                if (lane == i) sts_u32(dst + 4u * i, w);
                if ((w >> lane) & 1u) sts_u8(dst + 32u + (uint32_t)(prefix + __popc(w & ((1u << lane) - 1u))), i * 32 + lane);
                prefix += __popc(w);
            }
            const uint32_t n_sub = (uint32_t)n_off * (uint32_t)p.nseg;
            if (lane == 0) { sts_u32(dst + 16u, (uint32_t)n_off); sts_u32(dst + 20u, n_sub); }
            __syncwarp();
            if constexpr (kResident) {
                // B descriptor (low word) of every sub-chunk of the tile, in processing order, for the MMA issuer
                const uint32_t bdesc_lo0 = (uint32_t)umma_desc_b<CH>(smem_base_u32);
                for (uint32_t sub = (uint32_t)lane; sub < n_sub; sub += 32u) {
                    uint32_t ord = sub, seg = 0;
                    if (CH == 128) { ord = (sub * p.inv_nseg) >> 16; seg = sub - ord * (uint32_t)p.nseg; }
                    const uint32_t k = (uint32_t)lds_u8(dst + 32u + ord);
                    sts_u32(dst + (uint32_t)kNbrHeaderMin + 4u * sub, bdesc_lo0 + (k * (uint32_t)p.nseg + seg) * (b_sub_bytes >> 4));
                }
                __syncwarp();
            }
            const uint32_t total = (uint32_t)p.kvol * (QL_TILE_M * 4u);
            if (lane == 0) ql_mbar_arrive_expect_tx(bar, total);     // release: orders the header stores
            __syncwarp();
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.nbr + tile * (int64_t)p.kvol * QL_TILE_M);
            const uint32_t off = (uint32_t)lane * 16384u;
            if (off < total) ql_bulk_g2s(dst + (uint32_t)p.nbr_hdr + off, src + off, total - off < 16384u ? total - off : 16384u, bar);
            __syncwarp();
        };
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
#pragma unroll
        for (int i = 0; i < kMaskWords; ++i) pf_mask[i] = 0u;
        if (pf_tile < n_tiles) load_tile_mask_raw(p, pf_tile, pf_mask);
        auto prefetch_step = [&]() {
            if (pf_tile >= n_tiles) return;
            uint32_t m[kMaskWords];
#pragma unroll
            for (int i = 0; i < kMaskWords; ++i) m[i] = pf_mask[i];
            const int n_off = finish_tile_mask(m);                 // first use of the words requested one step ago
            const int64_t t = pf_tile;
            const uint32_t itn = pf_it;
            pf_tile += gridDim.x; ++pf_it;
            if (pf_tile < n_tiles) load_tile_mask_raw(p, pf_tile, pf_mask);   // the next tile's words: requested, not looked at
            prefetch_nbr(t, itn, m, n_off);
        };
        for (int i = 0; i < p.nbr_bufs - 1; ++i) prefetch_step();
        // streamed weights: up to half the ring per pass, lane l = sub-chunk l of the pass (unit l / kGroup)
        uint32_t u = 0, ph = 0, it = 0;
        const uint32_t batch_subs = (R / 2u > 0u ? R / 2u : 1u) << kGroupLog2;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            prefetch_step();
            if constexpr (!kResident) {
                const uint32_t buf = nbr_s0 + (it & nbmask) * nbr_stride;     // this tile's header (already resident: prefetched earlier)
                const uint32_t n_sub = (uint32_t)ql_lds_s32(buf + 20u);
                for (uint32_t c0 = 0; c0 < n_sub; c0 += batch_subs) {
                    const uint32_t sub = c0 + (uint32_t)lane;
                    const bool mine = (uint32_t)lane < batch_subs && sub < n_sub;
                    const uint32_t bu = (uint32_t)lane >> kGroupLog2, j = (uint32_t)lane & (uint32_t)(kGroup - 1);
                    uint32_t uu = u + bu, pp = ph;
                    if (uu >= R) { uu -= R; pp ^= 1u; }
                    const uint32_t fbar = full0 + uu * 8u;
                    if (mine && j == 0) {
                        ql_mbar_wait(empty0 + uu * 8u, pp ^ 1u);
                        const uint32_t left = n_sub - sub;
                        ql_mbar_arrive_expect_tx(fbar, (left < (uint32_t)kGroup ? left : (uint32_t)kGroup) * b_sub_bytes);
                    }
                    __syncwarp();
                    if (mine) {
                        uint32_t ord = sub, seg = 0;
                        if (CH == 128) { ord = (sub * p.inv_nseg) >> 16; seg = sub - ord * (uint32_t)p.nseg; }
                        const uint32_t k = (uint32_t)lds_u8(buf + 32u + ord);
                        ql_bulk_g2s(smem_base_u32 + (uu * (uint32_t)kGroup + j) * b_sub_bytes,
                                    p.w_packed + (size_t)(k * (uint32_t)p.nseg + seg) * b_sub_bytes, b_sub_bytes, fbar);
                    }
                    __syncwarp();
                    const uint32_t left = n_sub - c0;
                    const uint32_t nu = ((left < batch_subs ? left : batch_subs) + (uint32_t)kGroup - 1u) >> kGroupLog2;
                    u += nu;
                    if (u >= R) { u -= R; ph ^= 1u; }
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
    int pair;    // 16-byte rows: one 32-byte chunk holds two consecutive kernel offsets (2j | 2j+1)
};
inline ChunkGeom chunk_geom(int row_bytes) {
    ChunkGeom g;
    if (row_bytes > 128) { g.ch = 128; g.nseg = (row_bytes + 127) / 128; }
    else { g.ch = row_bytes <= 32 ? 32 : (row_bytes <= 64 ? 64 : 128); g.nseg = 1; }
    g.pair = row_bytes == 16 ? 1 : 0;
    return g;
}
// byte offset of 16-byte piece c16 of row r inside a K-major swizzled [rows x ch bytes] chunk image
inline uint32_t chunk_sw_offset(int ch, uint32_t r, uint32_t c16) {
    const uint32_t x = ch == 128 ? (r & 7u) : (ch == 64 ? ((r >> 1) & 3u) : ((r >> 2) & 1u));
    return (r >> 3) * (uint32_t)(8 * ch) + (r & 7u) * (uint32_t)ch + ((c16 ^ x) << 4);
}

// K order inside a sub-chunk: TMEM column c (4 bytes of K) of the A operand holds source word k_word_src(ch, c) of the
// row segment -- the identity for the lane-per-row gather (CH = 32), the 16x256b quad-gather order for CH >= 64.
inline int k_word_src(int ch, int c) {
    if (ch < 64) return c;
    const int krep = ch / 32;
    const int v2 = c >> 3, t0 = (c >> 1) & 3, e = c & 1;
    return 2 * krep * t0 + 2 * v2 + e;
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


// Shared-memory / TMEM plan of one launch (also answers "are this layer's weights streamed?" for the host).
// Returns the dynamic shared-memory bytes, or 0 when the shape is unsupported.
size_t plan_conv(ConvParams& p, const ChunkGeom& g, int c_out, int kvol) {
    const int kv = g.pair ? (kvol + 1) / 2 : kvol;            // kernel offsets as the weight tensor / B descriptors see them
    // ring depth in units (32 TMEM columns each): bounded by the TMEM columns left beside the accumulators and, when
    // the weights are streamed, by shared memory (a unit's B slot = its 128/CH weight sub-chunks = c_out x 128 bytes)
    const int misc_bytes = (int)sizeof(MiscSmem) + 4 * c_out * 4;
    const int b_sub = c_out * g.ch;
    const int b_unit = c_out * 128;
    p.inv_nseg = (uint32_t)((65536 + g.nseg - 1) / g.nseg);
    p.n_acc = (kTmemCols - 2 * c_out) / kUnitCols >= kTeams ? 2 : 1;
    int R = (kTmemCols - p.n_acc * c_out) / kUnitCols;
    if (R > kMaxUnits) R = kMaxUnits;
    p.w_bytes = kv * g.nseg * b_sub;
    const int hdr_resident = kNbrHeaderMin + ((4 * kv * g.nseg + 15) & ~15);
    for (p.nbr_bufs = 4; p.nbr_bufs >= 2; p.nbr_bufs >>= 1) {
        // try with the resident-weights header first; fall back to streamed weights (short header) if they do not fit
        p.nbr_hdr = hdr_resident;
        p.nbr_stride = (p.nbr_hdr + kvol * QL_TILE_M * 4 + 127) & ~127;
        if (p.nbr_bufs == 4 && 4 * p.nbr_stride > 64 * 1024) continue;
        int smem_free = kSmemBudget - 1024 - p.nbr_bufs * p.nbr_stride - ((misc_bytes + 127) & ~127);
        p.resident = (p.w_bytes <= smem_free && p.w_bytes % 512 == 0) ? 1 : 0;   // 32 lanes x 16-byte multiples
        if (!p.resident) {
            p.nbr_hdr = kNbrHeaderMin;
            p.nbr_stride = (p.nbr_hdr + kvol * QL_TILE_M * 4 + 127) & ~127;
            smem_free = kSmemBudget - 1024 - p.nbr_bufs * p.nbr_stride - ((misc_bytes + 127) & ~127);
        }
        if (p.resident || smem_free / b_unit >= 2) {
            if (!p.resident && R > smem_free / b_unit) R = smem_free / b_unit;
            break;
        }
    }
    if (p.nbr_bufs < 2 || R < 2) return 0;
    p.nbr_log2 = p.nbr_bufs == 4 ? 2 : 1;
    const int nbr_bytes = p.nbr_bufs * p.nbr_stride;
    p.n_ring = R;
    p.teams = R < kTeams ? R : kTeams;
    p.a_col0 = p.n_acc * c_out;
    p.off_nbr = p.resident ? ((p.w_bytes + 1023) & ~1023) : ((R * b_unit + 1023) & ~1023);
    p.off_misc = (p.off_nbr + nbr_bytes + 127) & ~127;
    size_t smem_bytes = 1024 + (size_t)p.off_misc + misc_bytes;
    if (smem_bytes < (size_t)kSmemFloor) smem_bytes = kSmemFloor;

    return smem_bytes;
}

}  // namespace

extern "C" size_t ql_packed_weight_bytes(int32_t c_in, int32_t c_out, int32_t kvol, int32_t elem_dtype) {
    int es = elem_size(elem_dtype);
    if (es == 0 || c_in <= 0 || c_out <= 0 || kvol <= 0) return 0;
    const ChunkGeom g = chunk_geom(c_in * es);
    return (size_t)(g.pair ? (kvol + 1) / 2 : kvol) * g.nseg * (size_t)c_out * g.ch;
}

// w_host: [c_out][kvol][c_in] elements (== the reference layout (oc, kd, kh, kw, ic) flattened, quant/quant.py:37-39).
// packed: for every (offset k, segment s) chunk one [c_out x CH bytes] K-major swizzled image (SWIZZLE_32B/64B/128B by
// CH), zero padded -- exactly what the loader warp bulk-copies into the chunk's shared-memory slot.  Inside a chunk row
// the 4-byte K words follow the order in which the gather leaves them in tensor memory (k_word_src).
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
            for (int seg = 0; seg < g.nseg; ++seg)
                for (int c = 0; c < g.ch / 4; ++c) {
                    const int b = seg * 128 + 4 * k_word_src(g.ch, c);          // source byte of this 4-byte K word
                    if (b >= row_bytes) continue;                                // zero padding
                    // pair mode: offset k lives in chunk k/2, bytes 16*(k%2) .. +15 of the chunk row
                    const size_t chunk = g.pair ? (size_t)(k / 2) : (size_t)(k * g.nseg + seg);
                    const int cc = g.pair ? c + 4 * (k & 1) : c;
                    memcpy(dst + chunk * chunk_bytes + chunk_sw_offset(g.ch, (uint32_t)oc, (uint32_t)(cc >> 2)) + 4 * (cc & 3),
                           src + ((size_t)oc * kvol + k) * row_bytes + b, 4);
                }
    return QL_OK;
}

extern "C" int ql_spconv_mma(const void* feats, int32_t in_dtype, const int32_t* nbr, const uint32_t* tile_kmask, int64_t n_out_cap,
                             const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t kvol, const void* w_packed,
                             const float* scale, const float* shift, const float* act_scale_dev, const void* residual_f16,
                             int32_t relu, void* out, int32_t out_dtype, int8_t* out_q, const float* out_qscale,
                             float* absmax, ql_stream_t stream_) {
    return ql_spconv_mma_rows(feats, in_dtype, nbr, tile_kmask, nullptr, n_out_cap, n_out_dev, c_in, c_out, kvol, w_packed, scale, shift,
                              act_scale_dev, residual_f16, relu, out, out_dtype, out_q, out_qscale, absmax, stream_);
}

extern "C" int ql_spconv_mma_rows(const void* feats, int32_t in_dtype, const int32_t* nbr, const uint32_t* tile_kmask,
                                  const int32_t* row_perm, int64_t n_out_cap, const int32_t* n_out_dev, int32_t c_in, int32_t c_out,
                                  int32_t kvol, const void* w_packed, const float* scale, const float* shift,
                                  const float* act_scale_dev, const void* residual_f16, int32_t relu, void* out, int32_t out_dtype,
                                  int8_t* out_q, const float* out_qscale, float* absmax, ql_stream_t stream_) {
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
    p.feats = (const uint8_t*)feats; p.nbr = nbr; p.kmask = tile_kmask; p.row_perm = row_perm; p.n_out_dev = n_out_dev; p.n_out_cap = n_out_cap;
    p.row_bytes = c_in * es; p.c_out = c_out; p.kvol = kvol;
    p.wide = (p.row_bytes % 32 == 0 && ((uintptr_t)feats & 31) == 0) ? 1 : 0;
    const ChunkGeom g = chunk_geom(p.row_bytes);
    p.nseg = g.nseg; p.mask_words = (kvol + 31) / 32; p.pair = g.pair;
    p.w_packed = (const uint8_t*)w_packed; p.scale = scale; p.shift = shift; p.act_scale_dev = act_scale_dev;
    p.residual = (const __half*)residual_f16; p.relu = relu; p.out = out; p.out_dtype = out_dtype;
    p.out_q = out_q; p.out_qscale = out_qscale; p.absmax = absmax;

    const size_t smem_bytes = plan_conv(p, g, c_out, kvol);
    if (smem_bytes == 0) return QL_ERR_UNSUPPORTED;

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

// 1 when ql_spconv_mma streams this layer's weights per unit (they do not fit in shared memory) -- the layers whose L2 -> SM
// traffic is dominated by the weight stream.
extern "C" int32_t ql_spconv_weights_streamed(int32_t c_in, int32_t c_out, int32_t kvol, int32_t elem_dtype) {
    int es = elem_size(elem_dtype);
    if (es == 0 || c_in <= 0 || (c_in * es) % 16 != 0 || c_out < 16 || c_out % 16 != 0 || c_out > 256 || kvol <= 0 || kvol > 32 * kMaskWords)
        return 0;
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.row_bytes = c_in * es; p.c_out = c_out; p.kvol = kvol;
    const ChunkGeom g = chunk_geom(p.row_bytes);
    p.nseg = g.nseg; p.pair = g.pair;
    if (plan_conv(p, g, c_out, kvol) == 0) return 0;
    return p.resident ? 0 : 1;
}
