// Implicit gather-GEMM-scatter sparse convolution on tcgen05 tensor cores (sm_100a), fused epilogue.
//
// Replaces QConvNd.forward -> [EXT] spconv SubMConv3d/SparseConv3d forward (quant/quant.py:36-58) together with
// the BatchNorm1d / ReLU / residual-add that follow it (spconv_backbone.py:8-27,51-67).
//
// One persistent CTA per SM walks 128-row output tiles.  Per tile the conv is a sum of per-offset GEMMs
//     D[128, C_out] += A_k[128, C_in] . W_k[C_out, C_in]^T      for every kernel offset k that is non-empty in the tile
// (the rulebook's per-tile offset mask lists them; the rulebook is COMPACT: only the live (tile, offset) slabs exist).
// The A operand never exists in memory and never touches shared memory: output row r of the tile is TMEM lane r; gather
// producers load the neighbours' feature rows straight from global memory (L1/L2) into registers and write them to
// tensor memory with tcgen05.st; the MMA reads A from TMEM (the ".ts" operand form) and only the weights from shared memory.
//
// K is cut into sub-chunks: one sub-chunk = one kernel offset x one <=128-byte segment of the input row
// (CH = 32 / 64 / 128 bytes => 1 / 2 / 4 MMA k-steps, CH/4 TMEM columns).  The hand-off between the roles is the UNIT:
// up to U consecutive live sub-chunks of one tile = one SLOT of U*CH/4 TMEM columns (up to 224: a whole C = 16 tile, half a
// C = 32 tile, 6 offsets of C = 64, 2 of C = 128) with ONE full/empty mbarrier pair.  Round 1 handed over 32 columns at a
// time and spent two thirds of its time in that protocol (profiles/r01_conv_ablation.md: ~400 cycles per hand-off for the
// in-order MMA thread: mbarrier wait + tcgen05 fence + tcgen05.commit, plus the slot round trip of the producers); a unit
// now carries 4-7x the work per hand-off.  Roles (6 + 4T warps, T = 4 teams by default):
//   warps 0-3       epilogue    : tcgen05.ld accumulators -> dequant*BN scale/shift (+residual) (+ReLU)
//                                 -> fp16/fp32 rows (+ int8 re-quantised rows, + per-channel absmax)
//   warps 4..3+4T   gather      : warp = (team, TMEM lane quarter).  A unit's 32-column PIECES are dealt round-robin to the
//                                 teams; every producer warp arrives once per unit.  Loads are never predicated (a missing
//                                 neighbour reads a zero line); rows >= 64 bytes use a QUAD gather (four lanes read one
//                                 row's CH contiguous bytes, registers go to TMEM with tcgen05.st.16x256b; the K order this
//                                 leaves inside a sub-chunk is undone in the weight packing, k_word_src).
//   warp  4+4T      MMA issuer  : one lane; per unit one wait, all the unit's tcgen05.mma back to back, one commit.
//                                 Accumulators in TMEM, double buffered when they fit (tile i+1 accumulates while i drains).
//   warp  5+4T      loader      : turns the tile list into a STREAM OF UNIT DESCRIPTORS in a 4-deep shared-memory ring: header
//                                 {n_sub, first/last-of-tile flags, B descriptors} + the unit's rulebook slabs (one contiguous
//                                 cp.async.bulk of n_offsets x 512 bytes out of the compact rulebook).  Producers and the MMA
//                                 thread only ever look at the ring: no per-tile header work is left on the MMA thread's path.
//                                 When the packed weights do not fit in shared memory it also streams every unit's weight
//                                 sub-chunks into the slot's B area (one bulk copy per lane, two units behind the descriptors).
#include "ql_common.cuh"
#include <stdlib.h>
#include <string.h>

int ql_spconv_warp_try(const void* feats, int32_t in_dtype, const int32_t* nbr, const uint32_t* tile_kmask, const int32_t* row_perm,
                       int64_t n_out_cap, const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t kvol, const void* w_packed,
                       const float* scale, const float* shift, const float* act_scale_dev, const void* residual_f16, int32_t relu,
                       void* out, int32_t out_dtype, int8_t* out_q, const float* out_qscale, float* absmax, cudaStream_t st);   // spconv_warp.cu

namespace {

constexpr int kMaxTeams = 4;
constexpr int kEpilogueThreads = 128;
constexpr int kProducerWarp0 = 4;                         // warps 0-3: epilogue (warp % 4 == TMEM lane quarter for both roles)
constexpr int kThreadsMax = (6 + 4 * kMaxTeams) * 32;     // 704
constexpr int kMaxSlots = 4;                              // A/B slot ring depth
constexpr int kUbufs = 4;                                 // unit-descriptor ring depth (power of two)
constexpr int kMaxUnitSubs = 32;                          // sub-chunks per unit: one loader lane each
constexpr int kMaskWords = 4;                             // kernel volumes up to 128 (3^3, 5^3)
constexpr int kTmemCols = 512;
constexpr int kSmemBudget = 232448;                       // 227 KB opt-in maximum per CTA
constexpr int kSmemFloor = 120 * 1024;                    // always ask for > half an SM: one CTA (one TMEM owner) per SM
constexpr int kUbHdr = 16 + 4 * kMaxUnitSubs + 16;        // unit descriptor: n_sub, flags, sub0, pad | blo[32] | pad -> slabs at +160
constexpr uint32_t kFlagFirst = 1u, kFlagLast = 2u;

#ifdef QL_SPCONV_ABLATE
// test-time build flag (never in the product library): bit 1 = no gather loads, 2 = no tcgen05.st, 4 = no tcgen05.mma,
// 8 = no epilogue global traffic, 16 = no rulebook slab copies, 32 = no streamed-weight copies, 64 = no tcgen05.ld / epilogue arithmetic
__device__ int g_ablate = 0;
#define QL_ABL(bit) ((abl_ & (bit)) != 0)
#define QL_ABL_INIT const int abl_ = g_ablate
// role trace of the same test-time build: clock64() around every wait / work phase of each role's lead thread, summed over
// the CTAs of a launch into g_trace[launch % 64][role * 8 + counter] (tools/conv_sweep.py TRACE=1 prints the shares)
__device__ unsigned long long g_trace[64][32];
#define QL_TR_DECL(n) long long tr_[n] = {}; long long tr_t0_ = clock64(); const long long tr_start_ = tr_t0_
#define QL_TR(i) do { const long long t_ = clock64(); tr_[i] += t_ - tr_t0_; tr_t0_ = t_; } while (0)
#define QL_TR_FLUSH(role, n, lead) do { if (lead) { tr_[(n) - 1] = clock64() - tr_start_; \
        for (int i_ = 0; i_ < (n); ++i_) atomicAdd(&g_trace[p.trace_id & 63][(role) * 8 + i_], (unsigned long long)tr_[i_]); } } while (0)
#else
#define QL_ABL(bit) false
#define QL_ABL_INIT
#define QL_TR_DECL(n)
#define QL_TR(i)
#define QL_TR_FLUSH(role, n, lead)
#endif

struct ConvParams {
    const uint8_t* feats;   // [n_in][row_bytes], preceded by one all-zero row: feats[-1] is what a missing neighbour (index -1) reads
    const int* nbr;         // compact rulebook [tiles][kvol][128]: the first popc(kmask[tile]) slabs of a tile are its live offsets
    const uint32_t* kmask;  // [tiles][mask_words] or null (every offset live: the dense rulebook)
    const int* row_perm;    // [tiles][128] tile slot -> output row (-1 = padding) of a GROUPED rulebook, or null (slot == row)
    const int* n_out_dev;
    int64_t n_out_cap;
    int row_bytes;          // c_in * elem size
    int wide;               // rows (and the feature base) are 32-byte aligned: gather with 256-bit loads
    int c_out, kvol, nseg, mask_words;
    uint32_t inv_nseg;      // ceil(65536 / nseg): ord = (sub * inv_nseg) >> 16 for sub < 4096
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
    int teams;              // producer teams (4 warps each)
    int n_slots;            // A (TMEM) / B (shared memory, streamed weights) slot ring depth
    int unit_subs;          // U: sub-chunks per unit
    int unit_cols;          // S = U * CH / 4 TMEM columns per slot
    int n_acc;              // accumulator buffers in TMEM (2, or 1 when 2*c_out does not fit beside the A slots)
    int a_col0;             // first TMEM column of the A slots
    int resident;           // 1: every weight chunk lives in shared memory for the whole kernel (no per-unit B copies)
    int w_bytes;            // packed weight bytes (resident mode)
    int off_ub;             // smem offset of the unit-descriptor ring
    int ub_stride;          // bytes per ring entry
    int off_tab;            // smem offset of the loader's ordinal -> offset table (128 bytes)
    int off_misc;           // smem offset of MiscSmem from the 1024-aligned base
    int trace_id;           // (test-time build) launch number for the role trace
};

struct MiscSmem {
    uint64_t full[kMaxSlots];
    uint64_t empty[kMaxSlots];
    uint64_t ub_full[kUbufs];
    uint64_t ub_empty[kUbufs];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
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
// base + id * mul in ONE instruction (IMAD.WIDE, signed 32 x 32 + 64): the address of gathered row `id` (-1 = the zero row)
__device__ __forceinline__ const uint8_t* row_ptr(const uint8_t* base, int id, int mul) {
    uint64_t r;
    asm volatile("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(id), "r"(mul), "l"(base));
    return reinterpret_cast<const uint8_t*>(r);
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

// the two halves of reading a tile's offset mask: issue the loads now, use them (count / fix-up) later -- a warp issues in order,
// so a popcount right behind its load costs the loader one global round trip per tile
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

template <bool kInt8, int CH, bool kResident>
__global__ void __launch_bounds__(kThreadsMax, 1) k_spconv_ts(const ConvParams p) {
    constexpr int kAReg = CH / 4;                          // 32-bit TMEM columns (registers) per sub-chunk
    constexpr int kGroup = 128 / CH;                       // sub-chunks per 32-column piece
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
    const int T = p.teams;
    QL_ABL_INIT;
    const int mma_warp = kProducerWarp0 + 4 * T, loader_warp = mma_warp + 1;
    const uint32_t n_slots = (uint32_t)p.n_slots;
    const uint32_t S = (uint32_t)p.unit_cols;

    const int64_t n_out = p.n_out_dev ? min((int64_t)*p.n_out_dev, p.n_out_cap) : p.n_out_cap;
    const int64_t n_tiles = (n_out + QL_TILE_M - 1) / QL_TILE_M;

    if (tid == 0) {
        for (int s = 0; s < kMaxSlots; ++s) {
            ql_mbar_init(ql_smem_u32(&misc->full[s]), 4 * T + (kResident ? 0 : 1));   // every producer warp (+ the loader's expect_tx)
            ql_mbar_init(ql_smem_u32(&misc->empty[s]), 1);                             // tcgen05.commit
        }
        for (int i = 0; i < kUbufs; ++i) {
            ql_mbar_init(ql_smem_u32(&misc->ub_full[i]), 1);
            ql_mbar_init(ql_smem_u32(&misc->ub_empty[i]), 4 * T + 1);                  // producer warps + the MMA issuer
        }
        for (int i = 0; i < 2; ++i) {
            ql_mbar_init(ql_smem_u32(&misc->acc_full[i]), 1);
            ql_mbar_init(ql_smem_u32(&misc->acc_empty[i]), kEpilogueThreads);
        }
        ql_mbar_init(ql_smem_u32(&misc->w_full), 1);
        ql_fence_mbar_init();
    }
    {
        const float act = p.act_scale_dev ? *p.act_scale_dev : 1.0f;
        for (int c = tid; c < p.c_out; c += blockDim.x) {
            s_scale[c] = p.scale[c] * act;
            s_shift[c] = p.shift[c];
            s_absmax[c] = 0u;
            s_qscale[c] = p.out_qscale ? p.out_qscale[c] : 0.f;
        }
    }
    if (warp == mma_warp) {
        ql_tmem_alloc(ql_smem_u32(&misc->tmem_base), kTmemCols);
        ql_tmem_relinquish();
    }
    ql_tc_fence_before();
    __syncthreads();
    ql_tc_fence_after();
    const uint32_t tmem_base = misc->tmem_base;
    const uint32_t b_sub_bytes = (uint32_t)p.c_out * CH;     // one weight sub-chunk: [c_out x CH bytes]
    const uint32_t ub_s0 = smem_base_u32 + (uint32_t)p.off_ub;
    const uint32_t ub_stride = (uint32_t)p.ub_stride;
    const uint32_t full0 = ql_smem_u32(&misc->full[0]), empty0 = ql_smem_u32(&misc->empty[0]);
    const uint32_t ubfull0 = ql_smem_u32(&misc->ub_full[0]), ubempty0 = ql_smem_u32(&misc->ub_empty[0]);

    if (warp < kProducerWarp0) {
        // ================================ epilogue ================================
        const int w = warp;                                  // TMEM lane quarter (warp id % 4)
        const int et = tid;                                  // 0..127 == row in tile
        uint32_t it = 0;
        QL_TR_DECL(8);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int a = p.n_acc == 2 ? (int)(it & 1u) : 0;
            const uint32_t aph = p.n_acc == 2 ? ((it >> 1) & 1u) : (it & 1u);
            // grouped rulebook: the tile's rows are scattered over the output; the slot -> row load hides under the wait
            const int64_t slot = tile * QL_TILE_M + et;
            int64_t row = slot;
            if (p.row_perm) row = slot < n_out ? (int64_t)__ldg(p.row_perm + slot) : -1;
            const bool row_ok = row >= 0 && row < n_out && !QL_ABL(8);
            // The residual row does not depend on the MMAs: (kind::f16 instantiations) its first 32 bytes are requested BEFORE the
            // accumulator wait and every further chunk one iteration ahead, so the global-load latency is off the epilogue's critical
            // path.  The kind::i8 instantiations keep the load inside the column loop: with the early loads their abs-max epilogue
            // (dynamic W8A8) ran 25 % slower on every layer (round 1).
            const bool has_res = p.residual != nullptr && row_ok;
            const uint4* res4 = has_res ? reinterpret_cast<const uint4*>(p.residual + row * p.c_out) : nullptr;
            uint4 ra = make_uint4(0u, 0u, 0u, 0u), rb = ra;
            if constexpr (!kInt8) {
                if (has_res) { ra = __ldg(res4); rb = __ldg(res4 + 1); }
            }
            QL_TR(1);
            ql_mbar_wait(ql_smem_u32(&misc->acc_full[a]), aph);
            QL_TR(0);
            ql_tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(a * p.c_out);
            for (int c0 = 0; c0 < p.c_out; c0 += 16) {
                if (QL_ABL(64)) break;
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
            QL_TR(1);
        }
        QL_TR_FLUSH(0, 8, tid == 0);
    } else if (warp < mma_warp) {
        // ============================ gather producers ============================
        // Instruction diet (profiles/r02_conv_ablation.md: with loads, tcgen05.st, MMAs and the epilogue's traffic all switched
        // off the launches still took 71 % of their time, and halving the producer warps cost only +23 %: the kernel was bound by
        // instruction issue, ~160 SASS instructions per 4 KB piece).  Per gathered 16/32 bytes a thread now issues one LDS (row
        // index), one IMAD.WIDE (address) and the load: a missing neighbour (index -1) reads the all-zero row the caller keeps in
        // front of the feature rows (include/qlidar.h), so nothing is predicated or selected; units always hold whole pieces (the
        // loader pads with all -1 slabs), so nothing depends on the sub-chunk count.
        const int q = warp & 3;                              // TMEM lane quarter
        const uint32_t team = (uint32_t)((warp - kProducerWarp0) >> 2);
        const uint32_t Tu = (uint32_t)T;
        const uint32_t a_lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)p.a_col0;
        const int t0 = lane & 3, t1 = lane >> 2;             // quad form: lane t0 of the quad that serves rows t1 + 8*rr
        // byte offset of this thread's first row inside a [128] int32 slab
        const uint32_t row_off = (uint32_t)(q * 32 + (kQuad ? t1 : lane)) * 4u;
        const uint32_t row_bytes = (uint32_t)p.row_bytes;
        const uint32_t nseg = (uint32_t)p.nseg, inv_nseg = p.inv_nseg;
        const bool wide = p.wide != 0;
        // this thread's bytes inside a row segment; a thread whose bytes lie past the end of a short row (row_bytes < CH, one
        // segment) always reads the zero row: multiplier 0, base = the zero row
        const uint32_t tb0 = kQuad ? (uint32_t)t0 * (CH / 4) : 0u;
        const bool dead0 = nseg == 1u && tb0 >= row_bytes;
        const uint8_t* const feats = p.feats;
        const uint8_t* const zrow = feats - row_bytes;
        const uint8_t* base0 = dead0 ? zrow : feats + tb0;
        asm volatile("mov.b64 %0, %0;" : "+l"(base0));     // one 64-bit register pair: the addend of the IMAD.WIDE of every gather
        const int mul0 = dead0 ? 0 : (int)row_bytes;

        uint32_t slot = 0, sph = 0;                          // slot ring position / pass parity
        uint32_t gmod = 0;                                   // (pieces handed out so far) mod T: piece g goes to team g mod T
        QL_TR_DECL(8);
        for (uint32_t u = 0;; ++u) {
            const uint32_t ub = u & (uint32_t)(kUbufs - 1);
            const uint32_t ubuf = ub_s0 + ub * ub_stride;
            QL_TR(3);
            ql_mbar_wait(ubfull0 + ub * 8u, (u / kUbufs) & 1u);
            QL_TR(0);
            const uint32_t n_sub = (uint32_t)ql_lds_s32(ubuf);
            if (n_sub == 0u) break;                          // end of the unit stream
            const uint32_t sub0 = (uint32_t)ql_lds_s32(ubuf + 8u);
            const uint32_t ord0 = (CH == 128 && nseg > 1u) ? ((sub0 * inv_nseg) >> 16) : sub0;
            const uint32_t pieces = (n_sub + kGroup - 1) >> kGroupLog2;
            uint32_t pc = team >= gmod ? team - gmod : team + Tu - gmod;      // this team's first piece of the unit
            const uint32_t slabs = ubuf + (uint32_t)kUbHdr + row_off;
            bool waited = false;
            for (; pc < pieces; pc += Tu) {
                uint32_t v[32];
                if constexpr (kQuad) {
#pragma unroll
                    for (int j = 0; j < kGroup; ++j) {
                        uint32_t slab = (pc << kGroupLog2) + (uint32_t)j;        // sub-chunk inside the unit == slab for one-segment rows
                        const uint8_t* base = base0;
                        int mul = mul0;
                        uint32_t tb = tb0;
                        if constexpr (CH == 128) {
                            if (nseg > 1u) {                                     // rows longer than 128 bytes: (offset, segment) -> slab, byte offset
                                const uint32_t c = sub0 + slab;
                                const uint32_t ord = (c * inv_nseg) >> 16;
                                tb = (c - ord * nseg) * 128u + tb0;
                                slab = ord - ord0;
                                const bool dead = tb >= row_bytes;
                                base = dead ? zrow : feats + tb;
                                mul = dead ? 0 : (int)row_bytes;
                            }
                        }
                        const uint32_t a = slabs + slab * (QL_TILE_M * 4u);
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) {
                            const int h = rr >> 1, v1 = rr & 1;
                            const int id = QL_ABL(1) ? -1 : ql_lds_s32(a + (uint32_t)rr * 32u);
                            const uint8_t* src = row_ptr(base, id, mul);
                            uint32_t x[8];
                            if constexpr (CH == 128) {
                                if (wide) {
                                    ldg32(src, x);
                                } else {
                                    ldg16(src, x);
                                    ldg16(tb + 16u < row_bytes ? src + 16 : zrow, x + 4);   // rows of an odd number of 16-byte pieces
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
                    const uint32_t a = slabs + (pc << kGroupLog2) * (QL_TILE_M * 4u);
#pragma unroll
                    for (int j = 0; j < kGroup; ++j) {
                        const int id = QL_ABL(1) ? -1 : ql_lds_s32(a + (uint32_t)j * (QL_TILE_M * 4u));   // CH = 32: one sub-chunk per kernel offset
                        const uint8_t* src = row_ptr(feats, id, (int)row_bytes);
                        if (wide) {
                            ldg32(src, v + j * 8);
                        } else if (row_bytes > 16u) {
                            ldg16(src, v + j * 8);
                            ldg16(src + 16, v + j * 8 + 4);
                        } else {
                            ldg16(src, v + j * 8);                                // 16-byte rows: the upper half of the k-step is padding
                            v[j * 8 + 4] = 0u; v[j * 8 + 5] = 0u; v[j * 8 + 6] = 0u; v[j * 8 + 7] = 0u;
                        }
                    }
                }
                const uint32_t a_piece = a_lane_base + slot * S + pc * 32u;
                if (!waited) {
                    QL_TR(1);
                    ql_mbar_wait(empty0 + slot * 8u, sph ^ 1u);                      // the MMAs that read this slot have completed
                    QL_TR(2);
                    ql_tc_fence_after();
                    waited = true;
                }
#ifdef QL_SPCONV_ABLATE
                asm volatile("" ::"r"(v[0]), "r"(v[31]));                           // (trace) the loads have landed before the clock is read
                QL_TR(1);
#endif
                if (!QL_ABL(2)) {
                    if constexpr (kQuad) {
#pragma unroll
                        for (int j = 0; j < kGroup; ++j)
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                tmem_st_16x256b<kRep>(a_piece + ((uint32_t)(h * 16) << 16) + (uint32_t)(j * kAReg), &v[j * kAReg + h * (4 * kRep)]);
                    } else {
                        tmem_st_32x32b_x32(a_piece, v);
                    }
                }
            }
            // a warp without a piece in this unit still arrives, and must do so inside the slot's current phase
            QL_TR(3);
            if (!waited) ql_mbar_wait(empty0 + slot * 8u, sph ^ 1u);
            QL_TR(2);
            tmem_st_wait();
            ql_tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                ql_mbar_arrive(full0 + slot * 8u);
                ql_mbar_arrive(ubempty0 + ub * 8u);          // this warp has read all it needs from the descriptor
            }
            gmod += pieces;                                  // pieces <= 32, T <= 4
            gmod = T == 4 ? (gmod & 3u) : gmod % Tu;
            if (++slot == n_slots) { slot = 0; sph ^= 1u; }
        }
        QL_TR_FLUSH(1, 8, warp == kProducerWarp0 && lane == 0);
    } else if (warp == mma_warp) {
        // =============================== MMA issuer ===============================
        // One elected lane runs the whole loop (nothing in it is warp-collective): per unit one descriptor wait, one slot wait,
        // the unit's MMAs back to back, one commit.
        if (ql_elect_one()) {
            const uint32_t idesc = make_idesc<kInt8>(p.c_out);
            const uint64_t bdesc0 = umma_desc_b<CH>(smem_base_u32);
            const uint32_t bdesc_hi = (uint32_t)(bdesc0 >> 32), bdesc_lo0 = (uint32_t)bdesc0;
            const uint32_t b_sub16 = b_sub_bytes >> 4;
            const uint32_t a_base = tmem_base + (uint32_t)p.a_col0;
            const uint32_t n_acc = (uint32_t)p.n_acc, c_out = (uint32_t)p.c_out, U = (uint32_t)p.unit_subs;
            uint32_t slot = 0, sph = 0, it = 0, accumulate = 0u, d_tmem = tmem_base, acc_bar = 0;
            if (kResident && (int64_t)blockIdx.x < n_tiles) ql_mbar_wait(ql_smem_u32(&misc->w_full), 0);
            QL_TR_DECL(8);
            for (uint32_t u = 0;; ++u) {
                const uint32_t ub = u & (uint32_t)(kUbufs - 1);
                const uint32_t ubuf = ub_s0 + ub * ub_stride;
                QL_TR(4);
                ql_mbar_wait(ubfull0 + ub * 8u, (u / kUbufs) & 1u);
                QL_TR(0);
                uint32_t n_sub, flags;
                asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(n_sub), "=r"(flags) : "r"(ubuf));
                if (n_sub == 0u) break;
                uint32_t blo[4];
                auto load_blo = [&](uint32_t j0) {
                    if constexpr (kResident) {
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(blo[0]), "=r"(blo[1]), "=r"(blo[2]), "=r"(blo[3]) : "r"(ubuf + 16u + 4u * j0));
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) blo[i] = bdesc_lo0 + (slot * U + j0 + (uint32_t)i) * b_sub16;
                    }
                };
                load_blo(0);
                if (flags & kFlagFirst) {
                    const uint32_t a = n_acc == 2 ? (it & 1u) : 0u;
                    const uint32_t aph = n_acc == 2 ? ((it >> 1) & 1u) : (it & 1u);
                    QL_TR(4);
                    ql_mbar_wait(ql_smem_u32(&misc->acc_empty[a]), aph ^ 1u);
                    QL_TR(1);
                    d_tmem = tmem_base + a * c_out;
                    acc_bar = ql_smem_u32(&misc->acc_full[a]);
                    accumulate = 0u;
                }
                QL_TR(4);
                ql_mbar_wait(full0 + slot * 8u, sph);
                QL_TR(2);
                ql_tc_fence_after();
                const uint32_t a_unit = a_base + slot * S;
                for (uint32_t j0 = 0; j0 < n_sub; j0 += 4) {
                    uint32_t cur[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) cur[i] = blo[i];
                    if (j0 + 4 < n_sub) load_blo(j0 + 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (i == 0 || j0 + (uint32_t)i < n_sub) {
#pragma unroll
                            for (int ks = 0; ks < CH / 32; ++ks) {       // one k-step = 32 bytes of K: +8 TMEM columns, +2 in the desc (addr >> 4)
                                const uint64_t bdesc = ((uint64_t)bdesc_hi << 32) | (uint64_t)(cur[i] + (uint32_t)(ks * 2));
                                if (!QL_ABL(4)) tc_mma_ts<kInt8>(d_tmem, a_unit + (j0 + (uint32_t)i) * (uint32_t)kAReg + (uint32_t)(ks * 8), bdesc, idesc, accumulate);
                                accumulate = 1u;
                            }
                        }
                    }
                }
                QL_TR(3);
                ql_tc_commit(empty0 + slot * 8u);
                if (flags & kFlagLast) { ql_tc_commit(acc_bar); ++it; }
                ql_mbar_arrive(ubempty0 + ub * 8u);
                if (++slot == n_slots) { slot = 0; sph ^= 1u; }
            }
            QL_TR_FLUSH(2, 8, true);
        }
        __syncwarp();
    } else if (warp == loader_warp) {
        // ================================= loader =================================
        const uint32_t tab = smem_base_u32 + (uint32_t)p.off_tab;          // ordinal -> kernel offset of the tile being cut into units
        const uint32_t U = (uint32_t)p.unit_subs, nseg = (uint32_t)p.nseg, inv_nseg = p.inv_nseg;
        const uint32_t bdesc_lo0 = (uint32_t)umma_desc_b<CH>(smem_base_u32);
        const uint32_t b_sub16 = b_sub_bytes >> 4;
        if (kResident && (int64_t)blockIdx.x < n_tiles) {
            // the whole packed weight tensor, 32 lanes x (w_bytes / 32) bytes
            const uint32_t bar = ql_smem_u32(&misc->w_full);
            const uint32_t per_lane = (uint32_t)p.w_bytes / 32u;
            if (lane == 0) ql_mbar_arrive_expect_tx(bar, (uint32_t)p.w_bytes);
            __syncwarp();
            ql_bulk_g2s(smem_base_u32 + (uint32_t)lane * per_lane, p.w_packed + (size_t)lane * per_lane, per_lane, bar);
            __syncwarp();
        }
        // descriptor cursor: tile state
        int64_t next_tile = blockIdx.x;                 // next tile to start
        int64_t cur_tile = 0;
        uint32_t s0 = 0, n_sub_tile = 0;                // next sub-chunk of the current tile / its total (s0 >= total: start a tile)
        bool synth = false;                             // the current tile has no pairs: one all -1 slab is synthesised
        uint32_t pf_mask[kMaskWords];
#pragma unroll
        for (int i = 0; i < kMaskWords; ++i) pf_mask[i] = 0u;
        if (next_tile < n_tiles) load_tile_mask_raw(p, next_tile, pf_mask);
        uint32_t ur = 0;                                // units described so far
        QL_TR_DECL(8);
        // returns true when it pushed the end-of-stream descriptor
        auto push_desc = [&]() -> bool {
            const uint32_t ub = ur & (uint32_t)(kUbufs - 1);
            const uint32_t ubuf = ub_s0 + ub * ub_stride;
            const uint32_t bar = ubfull0 + ub * 8u;
            if (s0 >= n_sub_tile) {
                if (next_tile >= n_tiles) {
                    QL_TR(2);
                    ql_mbar_wait(ubempty0 + ub * 8u, ((ur / kUbufs) & 1u) ^ 1u);
                    QL_TR(0);
                    if (lane == 0) { sts_u32(ubuf, 0u); sts_u32(ubuf + 4u, 0u); ql_mbar_arrive(bar); }
                    __syncwarp();
                    ++ur;
                    return true;
                }
                // start the next tile: first use of the mask words requested one step ago
                uint32_t m[kMaskWords];
                int n_live = 0;
#pragma unroll
                for (int i = 0; i < kMaskWords; ++i) { m[i] = pf_mask[i]; n_live += __popc(m[i]); }
                synth = n_live == 0;
                if (synth) { m[0] = 1u; n_live = 1; }          // a tile without pairs still has to zero its accumulators
                cur_tile = next_tile;
                next_tile += gridDim.x;
                if (next_tile < n_tiles) load_tile_mask_raw(p, next_tile, pf_mask);   // the next tile's words: requested, not looked at
                __syncwarp();
                int prefix = 0;
#pragma unroll
                for (int i = 0; i < kMaskWords; ++i) {
                    const uint32_t w = m[i];
                    if ((w >> lane) & 1u) sts_u8(tab + (uint32_t)(prefix + __popc(w & ((1u << lane) - 1u))), i * 32 + lane);
                    prefix += __popc(w);
                }
                __syncwarp();
                s0 = 0;
                n_sub_tile = (uint32_t)n_live * nseg;
            }
            const uint32_t left = n_sub_tile - s0;
            const uint32_t n_sub = left < U ? left : U;
            const uint32_t ord_first = CH == 128 ? ((s0 * inv_nseg) >> 16) : s0;
            const uint32_t ord_last = CH == 128 ? (((s0 + n_sub - 1u) * inv_nseg) >> 16) : s0 + n_sub - 1u;
            const uint32_t slab_bytes = (ord_last - ord_first + 1u) * (QL_TILE_M * 4u);
            QL_TR(2);
            ql_mbar_wait(ubempty0 + ub * 8u, ((ur / kUbufs) & 1u) ^ 1u);
            QL_TR(0);
            if ((uint32_t)lane < n_sub) {
                const uint32_t c = s0 + (uint32_t)lane;
                uint32_t ord = c, seg = 0;
                if (CH == 128) { ord = (c * inv_nseg) >> 16; seg = c - ord * nseg; }
                const uint32_t chunk = (uint32_t)lds_u8(tab + ord) * nseg + seg;
                // resident weights: the B descriptor (low word) of the sub-chunk; streamed: the chunk number for the weight cursor
                sts_u32(ubuf + 16u + 4u * (uint32_t)lane, kResident ? bdesc_lo0 + chunk * b_sub16 : chunk);
            }
            if (lane == 0) {
                sts_u32(ubuf, n_sub);
                sts_u32(ubuf + 4u, (s0 == 0u ? kFlagFirst : 0u) | (s0 + n_sub >= n_sub_tile ? kFlagLast : 0u));
                sts_u32(ubuf + 8u, s0);
            }
            // whole pieces for the producers: the sub-chunks that pad the unit's last piece get all -1 slabs (CH < 128 only: one
            // segment per offset, slab == sub-chunk)
            if constexpr (kGroup > 1) {
                const uint32_t padded = (n_sub + (uint32_t)kGroup - 1u) & ~(uint32_t)(kGroup - 1);
                for (uint32_t sl = n_sub; sl < padded; ++sl)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ubuf + (uint32_t)kUbHdr + sl * (QL_TILE_M * 4u) + 16u * (uint32_t)lane), "r"(-1) : "memory");
            }
            if (synth || QL_ABL(16)) {
                for (uint32_t sl = 0; sl < (QL_ABL(16) ? ord_last - ord_first + 1u : 1u); ++sl)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ubuf + (uint32_t)kUbHdr + sl * (QL_TILE_M * 4u) + 16u * (uint32_t)lane), "r"(-1) : "memory");
                __syncwarp();
                if (lane == 0) ql_mbar_arrive(bar);
            } else {
                __syncwarp();
                if (lane == 0) {
                    ql_mbar_arrive_expect_tx(bar, slab_bytes);         // release: orders the header stores
                    ql_bulk_g2s(ubuf + (uint32_t)kUbHdr, p.nbr + (cur_tile * (int64_t)p.kvol + ord_first) * QL_TILE_M, slab_bytes, bar);
                }
            }
            __syncwarp();
            s0 += n_sub;
            ++ur;
            return false;
        };
        if constexpr (kResident) {
            while (!push_desc()) {}
        } else {
            // streamed weights: the descriptors (and their rulebook copies) run up to 2 units ahead of the weight copies, which
            // are bound to the slot ring (a unit's B area is free once the MMAs of the unit n_slots earlier have completed)
            bool done = false;
            uint32_t uw = 0, slot = 0, sph = 0;
            for (;;) {
                while (!done && ur < uw + 3u) done = push_desc();
                const uint32_t ubuf = ub_s0 + (uw & (uint32_t)(kUbufs - 1)) * ub_stride;
                const uint32_t n_sub = (uint32_t)ql_lds_s32(ubuf);
                if (n_sub == 0u) break;
                const uint32_t fbar = full0 + slot * 8u;
                QL_TR(2);
                ql_mbar_wait(empty0 + slot * 8u, sph ^ 1u);
                QL_TR(1);
                if (QL_ABL(32)) {                                   // (test-time build) no weight stream: the B slots keep whatever they hold
                    if (lane == 0) ql_mbar_arrive(fbar);
                } else {
                    if (lane == 0) ql_mbar_arrive_expect_tx(fbar, n_sub * b_sub_bytes);
                    __syncwarp();
                    if ((uint32_t)lane < n_sub) {
                        const uint32_t chunk = (uint32_t)ql_lds_s32(ubuf + 16u + 4u * (uint32_t)lane);
                        ql_bulk_g2s(smem_base_u32 + (slot * U + (uint32_t)lane) * b_sub_bytes, p.w_packed + (size_t)chunk * b_sub_bytes, b_sub_bytes, fbar);
                    }
                }
                __syncwarp();
                ++uw;
                if (++slot == n_slots) { slot = 0; sph ^= 1u; }
            }
        }
        QL_TR(2);
        QL_TR_FLUSH(3, 8, lane == 0);
    }

    ql_tc_fence_before();
    __syncthreads();
    ql_tc_fence_after();
    if (p.absmax) {
        for (int c = tid; c < p.c_out; c += blockDim.x) {
            uint32_t v = s_absmax[c];
            if (v) atomicMax(reinterpret_cast<unsigned int*>(p.absmax) + c, v);
        }
    }
    if (warp == mma_warp) ql_tmem_dealloc(tmem_base, kTmemCols);
}

#ifdef QL_SPCONV_ABLATE
int g_launch_id = 0;
#endif

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
    k_spconv_ts<kInt8, CH, kResident><<<grid, (6 + 4 * p.teams) * 32, smem_bytes, st>>>(p);
    return cudaPeekAtLastError();                     // left pending for ql_last_cuda_error()
}
template <bool kInt8, int CH>
cudaError_t launch(const ConvParams& p, int grid, size_t smem_bytes, cudaStream_t st) {
    return p.resident ? launch2<kInt8, CH, true>(p, grid, smem_bytes, st) : launch2<kInt8, CH, false>(p, grid, smem_bytes, st);
}

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// Shared-memory / TMEM plan of one launch (also answers "are this layer's weights streamed?" for the host).
// Returns the dynamic shared-memory bytes, or 0 when the shape is unsupported.
//   TMEM: n_acc accumulators of c_out columns, then n_slots A slots of S = U * CH/4 columns (S a multiple of 32).
//   smem: [resident weights | n_slots x U streamed weight sub-chunks] [4 unit descriptors: 160-byte header + the unit's
//         rulebook slabs] [ordinal table] [barriers, scale/shift/absmax/qscale]
size_t plan_conv(ConvParams& p, const ChunkGeom& g, int c_out, int kvol) {
    // tuning overrides (tools/conv_sweep.py); read per call so that one process can sweep them
    const int force_u = env_int("QL_SPCONV_UNIT_SUBS", 0), force_slots = env_int("QL_SPCONV_SLOTS", 0),
              force_teams = env_int("QL_SPCONV_TEAMS", 0), force_stream = env_int("QL_SPCONV_STREAM", 0);
    const int group = 128 / g.ch, areg = g.ch / 4;
    const int misc_bytes = (int)sizeof(MiscSmem) + 4 * c_out * 4;
    const int b_sub = c_out * g.ch;
    p.inv_nseg = (uint32_t)((65536 + g.nseg - 1) / g.nseg);
    p.teams = force_teams > 0 && force_teams <= kMaxTeams ? force_teams : kMaxTeams;
    p.n_acc = 2 * c_out <= 256 ? 2 : 1;
    const int cols_left = kTmemCols - p.n_acc * c_out;
    const int tile_subs = (kvol * g.nseg + group - 1) / group * group;        // a whole tile's sub-chunks, in whole pieces
    int n_slots = 2;
    int U = cols_left / n_slots / 32 * group;                                  // pieces per slot x sub-chunks per piece
    if (U > kMaxUnitSubs) U = kMaxUnitSubs / group * group;
    if (U > tile_subs) U = tile_subs;
    if (force_u > 0) {
        U = (force_u + group - 1) / group * group;
        if (U > cols_left / n_slots / 32 * group) U = cols_left / n_slots / 32 * group;
        if (U > kMaxUnitSubs) U = kMaxUnitSubs / group * group;
    }
    if (U < group) return 0;
    p.w_bytes = kvol * g.nseg * b_sub;
    auto ub_stride_of = [&](int u) {
        const int offs = g.nseg > 1 ? (u + g.nseg - 2) / g.nseg + 1 : u;       // kernel offsets a unit of u sub-chunks can span (u is whole pieces)
        return (kUbHdr + offs * QL_TILE_M * 4 + 127) & ~127;
    };
    const int fixed = 1024 + 128 + ((misc_bytes + 127) & ~127);
    int smem_free = kSmemBudget - fixed - kUbufs * ub_stride_of(U);
    p.resident = (!force_stream && p.w_bytes <= smem_free && p.w_bytes % 512 == 0) ? 1 : 0;   // 32 lanes x 16-byte multiples
    if (!p.resident) {
        // the B slots hold n_slots x U sub-chunks
        while (U > group && n_slots * U * b_sub > kSmemBudget - fixed - kUbufs * ub_stride_of(U)) U -= group;
        smem_free = kSmemBudget - fixed - kUbufs * ub_stride_of(U);
        if (n_slots * U * b_sub > smem_free) return 0;
    }
    // more, smaller-or-equal slots when TMEM (and the B area) allow: the producers run further ahead of the MMA thread
    while (n_slots < kMaxSlots && (n_slots + 1) * U * areg <= cols_left &&
           (p.resident || (n_slots + 1) * U * b_sub <= smem_free))
        ++n_slots;
    if (force_slots >= 2 && force_slots <= kMaxSlots && force_slots * U * areg <= cols_left &&
        (p.resident || force_slots * U * b_sub <= smem_free))
        n_slots = force_slots;
    p.n_slots = n_slots;
    p.unit_subs = U;
    p.unit_cols = U * areg;
    p.a_col0 = p.n_acc * c_out;
    p.ub_stride = ub_stride_of(U);
    p.off_ub = p.resident ? ((p.w_bytes + 1023) & ~1023) : ((n_slots * U * b_sub + 1023) & ~1023);
    p.off_tab = p.off_ub + kUbufs * p.ub_stride;
    p.off_misc = p.off_tab + 128;
    size_t smem_bytes = 1024 + (size_t)p.off_misc + misc_bytes;
    if (smem_bytes > (size_t)kSmemBudget) return 0;
    if (smem_bytes < (size_t)kSmemFloor) smem_bytes = kSmemFloor;
    return smem_bytes;
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
                    const size_t chunk = (size_t)(k * g.nseg + seg);
                    memcpy(dst + chunk * chunk_bytes + chunk_sw_offset(g.ch, (uint32_t)oc, (uint32_t)(c >> 2)) + 4 * (c & 3),
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
    {
        // narrow layers (rows <= 64 bytes): the register-gather kernel of spconv_warp.cu
        const int r = ql_spconv_warp_try(feats, in_dtype, nbr, tile_kmask, row_perm, n_out_cap, n_out_dev, c_in, c_out, kvol, w_packed, scale, shift,
                                         act_scale_dev, residual_f16, relu, out, out_dtype, out_q, out_qscale, absmax, st);
        if (r != 1) return r;
    }

    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.feats = (const uint8_t*)feats; p.nbr = nbr; p.kmask = tile_kmask; p.row_perm = row_perm; p.n_out_dev = n_out_dev; p.n_out_cap = n_out_cap;
    p.row_bytes = c_in * es; p.c_out = c_out; p.kvol = kvol;
    p.wide = (p.row_bytes % 32 == 0 && ((uintptr_t)feats & 31) == 0) ? 1 : 0;
    const ChunkGeom g = chunk_geom(p.row_bytes);
    p.nseg = g.nseg; p.mask_words = (kvol + 31) / 32;
    p.w_packed = (const uint8_t*)w_packed; p.scale = scale; p.shift = shift; p.act_scale_dev = act_scale_dev;
    p.residual = (const __half*)residual_f16; p.relu = relu; p.out = out; p.out_dtype = out_dtype;
    p.out_q = out_q; p.out_qscale = out_qscale; p.absmax = absmax;

    const size_t smem_bytes = plan_conv(p, g, c_out, kvol);
    if (smem_bytes == 0) return QL_ERR_UNSUPPORTED;
#ifdef QL_SPCONV_ABLATE
    p.trace_id = g_launch_id++;
#endif

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
    p.nseg = g.nseg;
    if (plan_conv(p, g, c_out, kvol) == 0) return 0;
    return p.resident ? 0 : 1;
}

#ifdef QL_SPCONV_ABLATE
extern "C" int ql_debug_set_ablate(int32_t mask) {
    return cudaMemcpyToSymbol(g_ablate, &mask, sizeof(int)) == cudaSuccess ? QL_OK : QL_ERR_CUDA;
}
// copies the role trace [64][32] to the host and clears it
extern "C" int ql_debug_read_trace(unsigned long long* out_host) {
    if (cudaDeviceSynchronize() != cudaSuccess) return QL_ERR_CUDA;
    if (cudaMemcpyFromSymbol(out_host, g_trace, sizeof(unsigned long long) * 64 * 32) != cudaSuccess) return QL_ERR_CUDA;
    static unsigned long long zeros[64 * 32];
    g_launch_id = 0;
    return cudaMemcpyToSymbol(g_trace, zeros, sizeof(zeros)) == cudaSuccess ? QL_OK : QL_ERR_CUDA;
}
#endif
