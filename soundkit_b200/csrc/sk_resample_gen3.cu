// sk_resample_gen3.cu — general-ratio polyphase FIR, third form: 4 consecutive outputs x 64 columns per warp, packed FFMA2 against
// shifted tap rows, the 8 accumulator lanes of an output split over 4 (lane columns) or 2 (period columns) threads.
//
// Same arithmetic contract as resample_gen2_kernel (sk_resample.cu): every output is rubato's
//     p0 + frac * (p1 - p0),   p_k = 8 lane accumulators (lane l owns taps l, l+8, ... with fused multiply-adds in tap order)
//                                    reduced as ((a0+a4)+(a1+a5)) + ((a2+a6)+(a3+a7))
// so results are bit-identical to the AVX path of rubato 0.14.1 `SincFixedIn` as soundkit calls it
// (soundkit/src/audio_pipeline.rs:474-491, soundkit-decoder/src/lib.rs:1939-2031).
//
// Why a third form.  The 2 x 2 tile of gen2 spends 13-14 shared-memory wavefronts per 32 FFMA (LSU-bound at ~0.57 of the FMA
// peak, measured 0.28).  Here a warp owns A = 4 CONSECUTIVE outputs (n .. n+3) of 64 COLUMNS:
//   * a column is an independent input window that shares the group's table rows: another (stream, channel) lane of the batch
//     and/or the same lane one or more whole phase periods later.  With >= 64 lanes in a batch (BASELINE configs[3]: 4096
//     streams x 2 channels) columns are simply lanes, every plan entry is shared exactly and the period structure of the
//     ratio is not needed at all; smaller batches fill the 64 columns with periods of the same lane (the gen2 idea);
//   * the windows of 4 consecutive outputs overlap almost completely, so ONE pass over the input words x[m] serves all four:
//     output a needs tap j = m - d_a, where d_a is its window offset inside the group.  The warp stages its 8 table rows
//     (4 outputs x 2 interpolation products) in shared memory SHIFTED by d_a and zero padded: T_a[m] = h_a[m - d_a] (TMA bulk
//     copies from word-shifted copies of the table).  Every output then consumes the same (x[m], x[m+1]) register pair
//     against (T_a[m], T_a[m+1]) with one packed FFMA2;
//   * accumulators are indexed by m & 7 ("virtual lanes", compile time).  Rubato's lane of tap j is j & 7 = (m - d_a) & 7, a
//     rotation by d_a & 7 that is undone in the reduction: the tree pairs lane l with l + 4 — a rotation-invariant pairing —
//     and then sums adjacent pairs, which depends on the parity of the rotation only (two selects per product).
//     Zero-padded taps add x * 0 to an accumulator that can never be -0: an exact no-op for finite x; tiles that stage
//     non-finite or huge (>= 2^113) floats take the per-output path instead;
//   * the 8 virtual lanes of every output are split over a lane quad (pair, for period columns) that shares 8 (4) columns, and
//     exchanged with shuffles after the walk: a thread tile is 8 tap rows x 8 columns of accumulator pairs (128 registers),
//     fed by 16 LDS.64 per 64 FFMA2.
//
// Lanes whose own plan entries differ from the group's reference entries (possible only for period columns, where the f64
// index walk lands within rounding of a table-row boundary for some periods but not others) recompute their outputs with the
// per-output arithmetic afterwards; sk_ctx_resample_fallback_outputs() counts them.
#include "sk_resample_dev.cuh"

constexpr int kG3Threads = 256;
constexpr int kG3Warps = kG3Threads / 32;
constexpr int kG3A = 4;   // consecutive outputs per thread
constexpr int kG3B = 4;   // columns per thread (a lane pair q, q + 16 shares them and splits the 8 accumulator lanes)
constexpr int kG3Cols = 16 * kG3B;
constexpr int kG3Rows = 2 * kG3A;

// Pre-shifted copies of the sinc table for 16-byte row staging: copy sh (0..3) of row r is h_r delayed by sh words inside a
// zero-padded row of kG3RowLen words: G[sh][r][i] = h_r[i - sh - kG3PadL] (0 outside [0, 256)).  A row shifted by d words is then
// the 16-byte aligned slice G[d & 3][r][kG3PadL - (d - (d & 3)) ...] — staged with cp.async, no register round trip.
constexpr int kG3PadL = 32;                       // largest staged shift (multiple of 4)
constexpr int kG3RowLen = kG3PadL + 288 + 4;      // 324 words: the longest tap walk (288) starting at any offset <= kG3PadL, + the shift

// The tap walk is NOT fully unrolled.  264 words = 33 blocks x (64 FFMA2 + 16 LDS) is 42 KB of straight-line code; streamed through
// the instruction cache it caps a scheduler at ~0.45 issued instructions per clock whatever the instruction mix
// (tools/ub/ub_lds_split.cu: ncu stall reason no_instructions, walk alone 41-45 TFLOP/s fully unrolled, 54-56 as a loop of
// 3-11 blocks).  The walk therefore runs as a short loop of g3_walk_unroll<MLEN>() blocks that stays resident.
// (A trip count that is not a multiple of the unroll factor gets a compiler-generated remainder loop.)
template <int MLEN> __host__ __device__ constexpr int g3_walk_unroll() { return (MLEN / 8) % 12 == 0 ? 12 : 11; }
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct Gen3Geom {
    int C;            // lanes (stream x channel) per CTA: power of two <= 64
    int P;            // periods per tile = 64 / C
    int period;       // outputs per period; P == 1: outputs per tile (any multiple of 4 * warps)
    int col_stride;   // words between the windows of consecutive lanes in shared memory
    int span;         // words staged per lane (window of the whole tile + overshoot of the padded tap walk)
    int tap_stride;   // words per staged (shifted, zero padded) table row; the tap walk covers m in [0, tap_stride)
    int xvec;         // 1: lane windows and group bases are 16-byte aligned -> LDS.128 input loads (P == 1 only)
    int log2c;        // C == 1 << log2c
    long long n_lanes;
};

__device__ __forceinline__ float gen3_single(const float* x, const float* __restrict__ h0, const float* __restrict__ h1, int d, float frac) {
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, c[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int k = 0; k < SINC_LEN; k += 8) {
        float xs[9];
#pragma unroll
        for (int l = 0; l < 9; l++) xs[l] = x[k + l];
        const float4 ha = __ldg(reinterpret_cast<const float4*>(h0 + k)), hb = __ldg(reinterpret_cast<const float4*>(h0 + k + 4));
        const float4 hc = __ldg(reinterpret_cast<const float4*>(h1 + k)), hd = __ldg(reinterpret_cast<const float4*>(h1 + k + 4));
        const float g0[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
        const float g1[8] = {hc.x, hc.y, hc.z, hc.w, hd.x, hd.y, hd.z, hd.w};
#pragma unroll
        for (int l = 0; l < 8; l++) {
            a[l] = __fmaf_rn(xs[l], g0[l], a[l]);
            c[l] = __fmaf_rn(d ? xs[l + 1] : xs[l], g1[l], c[l]);
        }
    }
    const float p0 = reduce8(a), p1 = reduce8(c);
    return __fadd_rn(p0, __fmul_rn(frac, __fsub_rn(p1, p0)));  // interp_lin
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

// rubato's reduction of 8 lane accumulators held as 8 VIRTUAL lanes v[(l + rot) & 7] = lane l
__device__ __forceinline__ float reduce8_rotated(const unsigned long long (&acc)[4], int rot) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; k++) unpack2(acc[k], v[2 * k], v[2 * k + 1]);
    const float s0 = __fadd_rn(v[0], v[4]), s1 = __fadd_rn(v[1], v[5]), s2 = __fadd_rn(v[2], v[6]), s3 = __fadd_rn(v[3], v[7]);
    // (a_l + a_{l+4}) = s[(l + rot) & 3]; the tree adds pairs (0,1) and (2,3) of those: adjacent pairs starting at rot & 3
    const bool odd = (rot & 1) != 0;
    const float t0 = odd ? __fadd_rn(s1, s2) : __fadd_rn(s0, s1);
    const float t1 = odd ? __fadd_rn(s3, s0) : __fadd_rn(s2, s3);
    // rot & 3 == 2 or 3 swaps t0 and t1, and swaps the operands inside each pair sum: additions commute bit for bit
    return ((rot & 2) != 0) ? __fadd_rn(t1, t0) : __fadd_rn(t0, t1);
}

__device__ __forceinline__ bool g3_nonfinite(float v) { return (__float_as_uint(v) & 0x7f800000u) == 0x7f800000u; }

// the per-output arithmetic for one output of the lane whose staged window starts at `wc` (frame w0)
__device__ __noinline__ float g3_exact(const float* wc, int w0, const int32_t* __restrict__ pi0, const uint8_t* __restrict__ ps0, long long n,
                                          const float* __restrict__ table, float fr) {
    const int i0o = pi0[n], s0o = ps0[n];
    int s1 = s0o + 1, d = 0;
    if (s1 >= OVERSAMPLING) { s1 -= OVERSAMPLING; d = 1; }  // get_nearest_times_2
    return gen3_single(wc + (i0o - w0), table + s0o * SINC_LEN, table + s1 * SINC_LEN, d, fr);
}

// Tile prologue shared by the exact and the fast kernel: per-lane source / destination pointers (the divisions by the channel
// count happen once per CTA), then the windows of the tile's lanes into shared memory.
__device__ __forceinline__ void g3_stage_tile(const SrcDesc& src, const OutDesc& dst, const Gen3Geom& geo, long long n0, long long lane0, int w0, int count,
                                              int lanes_per_stream, float* win, const uint8_t** s_col_src, float** s_col_dst) {
    const int warp = threadIdx.x >> 5, q = threadIdx.x & 31;
    // ---- per-lane pointers (the divisions by the channel count happen once per CTA)
    if (threadIdx.x < geo.C) {
        const long long lane = lane0 + threadIdx.x;
        const uint8_t* sp = nullptr; float* dp = nullptr;
        if (lane < geo.n_lanes) {
            const long long stream = lane / lanes_per_stream;
            const int ch = (int)(lane % lanes_per_stream);
            sp = src.base + stream * src.stream_stride + (src.width == 0 ? (long long)ch * src.lane_stride : (long long)ch * src.width) +
                 ((long long)w0 + src.origin) * (src.width == 0 ? 4 : (long long)src.channels * src.width);
            dp = dst.base + stream * dst.stream_stride + (long long)ch * dst.lane_stride + n0;
        }
        s_col_src[threadIdx.x] = sp; s_col_dst[threadIdx.x] = dp;
    }
    __syncthreads();

    // ---- stage the windows: a warp takes lanes warp, warp + 8, ... and walks each in chunks of 32 frames.  32-bit samples are
    // fetched with 4-byte cp.async (every copy of the tile in flight at once, no register round trip) and, where the format
    // needs it (big-endian / integer / non-finite scrub), converted in place by the thread that fetched them; 16 / 24-bit
    // samples go through registers, eight independent loads per lane at a time.
    {
        const int fstride = src.width == 0 ? 4 : src.channels * src.width;  // bytes between consecutive frames of a lane
        const int lo = (int)max(0ll, min((long long)count, src.valid_lo - w0)), hi = (int)max(0ll, min((long long)count, src.valid_hi - w0));
        if (src.width == 0 || (src.width == 4 && !src.byte_loads)) {
            const bool raw_float = src.width == 0 || (src.is_float && !src.big_endian && !src.scrub);
            for (int c = warp; c < geo.C; c += kG3Warps) {
                const uint8_t* sp = s_col_src[c];
                float* d = win + (size_t)c * geo.col_stride;
                const int lo_c = sp ? lo : count, hi_c = sp ? hi : count;  // a missing lane is all zeros
                for (int i = q; i < lo_c; i += 32) d[i] = 0.0f;
                // first i >= lo_c with i == q (mod 32); 32-bit byte offsets (a window is far below 2 GB)
                int i = q + ((max(lo_c - q, 0) + 31) & ~31);
                const uint32_t dst0 = smem_u32(d);
                for (; i < hi_c; i += 32)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + 4u * (uint32_t)i), "l"(sp + (uint32_t)i * (uint32_t)fstride) : "memory");
                for (; i < count; i += 32) d[i] = 0.0f;
            }
            cp_async_wait_all();
            if (!raw_float) {
                for (int c = warp; c < geo.C; c += kG3Warps) {  // the same elements this thread fetched
                    if (!s_col_src[c]) continue;
                    float* d = win + (size_t)c * geo.col_stride;
                    for (int i = max(lo, q + ((lo - q + 31) & ~31)); i < hi; i += 32) {
                        const uint32_t u = __float_as_uint(d[i]);
                        d[i] = decode_sample_bits(src.big_endian ? bswap32(u) : u, src);
                    }
                }
            }
        } else {
            constexpr int U = 8;
            for (int c = warp; c < geo.C; c += kG3Warps) {
                const uint8_t* sp = s_col_src[c];
                float* d = win + (size_t)c * geo.col_stride;
                for (int i0 = q; i0 < count; i0 += 32 * U) {
                    uint32_t bits[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int i = i0 + 32 * u;
                        bits[u] = 0;
                        if (sp && i >= lo && i < hi) {
                            const uint8_t* p = sp + (long long)i * fstride;
                            if (src.width != 2 || src.byte_loads) {
                                for (int bb = 0; bb < src.width; bb++) bits[u] |= (uint32_t)p[bb] << (8 * (src.big_endian ? src.width - 1 - bb : bb));
                            } else {
                                bits[u] = *reinterpret_cast<const uint16_t*>(p);
                                if (src.big_endian) bits[u] = ((bits[u] & 0xff) << 8) | (bits[u] >> 8);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int i = i0 + 32 * u;
                        if (i < count) d[i] = (sp && i >= lo && i < hi) ? decode_sample_bits(bits[u], src) : 0.0f;
                    }
                }
            }
        }
    }
    __syncthreads();

}

// XVEC == 8: lane columns only (P == 1, C == 64): every column shares every plan entry, windows and group bases are 16-byte
// aligned (vector input loads).  XVEC == 1: period columns (P > 1): scalar input loads, per-column plan entries checked
// against the group's reference entries; the accumulator lanes are split over lane PAIRS (q, q + 16; 4 columns per thread,
// LDS.128 tap loads).
// XVEC == 8 is the QUAD form: the 8 accumulator lanes of an output are split over FOUR threads (q, q + 8, q + 16, q + 24) that
// share 8 columns; quarter hh walks words m + 2 hh, m + 2 hh + 1 of every 8-word block, so every load is an LDS.64 and a thread
// tile is 8 tap rows x 8 columns: 16 x 8 bytes arrive in registers per 64 FFMA2 instead of 12 x 16 (the register tile is
// square, which minimises operand bytes per FMA for 64 accumulator pairs).  tools/ub/ub_lds_split.cu measures the walk alone:
// shared-memory loads cost by the bytes they return to the register file, not by wavefronts — LDS.128 twice an LDS.64 whatever
// the address pattern — and the quad walk runs 1.10x the pair walk.
template <int XVEC, int MLEN>
__global__ void __launch_bounds__(kG3Threads, 1)
resample_gen3_kernel(const __grid_constant__ SrcDesc src, const __grid_constant__ OutDesc dst, const __grid_constant__ Gen3Geom geo,
                     const int32_t* __restrict__ pi0, const uint8_t* __restrict__ ps0, const float* __restrict__ pfrac,
                     const float* __restrict__ table, const float* __restrict__ table4, int lanes_per_stream, unsigned long long* __restrict__ stats) {
    constexpr bool kLanes = XVEC == 8, kQuad = kLanes;
    constexpr int A = kG3A, B = kQuad ? 8 : kG3B;
    constexpr int NP = kQuad ? 1 : 2;   // accumulator pairs (of the thread's 2 or 4 virtual lanes) per product
    constexpr int HW = kQuad ? 2 : 4;   // words of every 8-word block this thread walks
    extern __shared__ __align__(16) float g3_smem[];
    __shared__ const uint8_t* s_col_src[kG3Cols];  // per lane of this CTA: first sample of the lane (channel plane for planar sources); null = no lane
    __shared__ float* s_col_dst[kG3Cols];          // output row of the lane
    __shared__ __align__(8) unsigned long long s_tap_bar[kG3Warps];  // per warp: completion of its 8 bulk row copies
    constexpr int TS = MLEN;  // words per staged row == length of the (fully unrolled) tap walk
    float* taps_all = g3_smem;                                   // [warps][8 rows][TS]
    float* win = g3_smem + kG3Warps * kG3Rows * TS;              // [C][col_stride]
    const int warp = threadIdx.x >> 5, q = threadIdx.x & 31;
    float* T = taps_all + warp * kG3Rows * TS;

    const long long tile_out = (long long)geo.period * geo.P;
    const long long n0 = (long long)blockIdx.x * tile_out;
    const int nt = (int)(min(dst.n_out, n0 + tile_out) - n0);  // outputs of this tile (tile-relative indices 0 .. nt - 1 from here on)
    pi0 += n0; ps0 += n0; pfrac += n0;
    const int w0 = pi0[0] - (kLanes ? (pi0[0] & 3) : 0);       // first staged frame (16-byte aligned window origin for LDS.128)
    const long long lane0 = (long long)blockIdx.y * geo.C;

    // frames staged per lane: the tile's windows plus the overshoot of the padded tap walk (zero taps meet these words: they must
    // be staged data)
    const int count = min(geo.span, pi0[nt - 1] - w0 + SINC_LEN + 2 + (TS - SINC_LEN) + 8);

    const uint32_t tap_bar = smem_u32(&s_tap_bar[warp]);
    if (q == 0) { mbar_init(tap_bar, 1); mbar_fence_init(); }  // visible to every lane after the barriers of the tile prologue
    uint32_t tap_phase = 0;
    g3_stage_tile(src, dst, geo, n0, lane0, w0, count, lanes_per_stream, win, s_col_src, s_col_dst);

    // ---- this thread's four columns.  The two half-warps share the 64 columns of the warp: lane pair (ql, ql + 16) owns columns
    // ql + 16 b (b < 4) and splits the 8 virtual accumulator lanes — half hh walks words m + 4 hh .. m + 4 hh + 3 of every 8-word
    // block.  A tap LDS.128 then carries two useful 16-byte rows (one per half-warp) in the two wavefronts a broadcast costs
    // anyway, and every tap register feeds 4 columns: 32 shared-memory wavefronts per 64 FFMA2 per thread.  Each thread keeps
    // the two columns it will finish in slots 0-1 and its partner's in slots 2-3 (slot b = column (b ^ 2 hh)), so the exchange
    // after the walk needs no selects.  column j -> (lane c = j % C, period k = j / C)
    // Quad form: threads (ql, hh = 0..3) share columns ql + 8 n (n < 8); slot b = column n = b ^ 2 hh, so a thread's own two
    // columns (n = 2 hh, 2 hh + 1) sit in slots 0-1, those of its second-step partner (hh ^ 1) in 2-3 and those of the
    // first-step partner's pair (hh ^ 2, hh ^ 3) in 4-7: both exchange steps send and receive the same slot numbers.
    constexpr int CW = kQuad ? 8 : 16;  // threads that own distinct columns
    const int ql = q & (CW - 1), hh = q / CW;
    // slot b of this thread = column j = ql + CW (b ^ 2 hh) -> lane c = j % C of the CTA, period k = j / C of the tile
    auto col_lane = [&](int b) { return (ql + CW * (b ^ (2 * hh))) & (geo.C - 1); };
    auto col_first = [&](int b) { return kLanes ? 0 : ((ql + CW * (b ^ (2 * hh))) >> geo.log2c) * geo.period; };  // first output of the period
    unsigned col_ok = 0;  // bit = slot: the lane exists
#pragma unroll
    for (int b = 0; b < B; b++) col_ok |= (lane0 + col_lane(b) < geo.n_lanes ? 1u : 0u) << b;
    const float* xcol[B];  // window of each slot's lane, + this half's word offset
#pragma unroll
    for (int b = 0; b < B; b++) xcol[b] = win + (size_t)col_lane(b) * geo.col_stride + HW * hh;

    // Reference plan entries of a group (period 0 of the tile; entries clamped at ragged ends are masked later), packed:
    //   base = window start of output 0 (moved down to a 16-byte boundary for the LDS.128 form: the taps shift up instead),
    //   s0p  = the four subindices, dp = the four tap shifts of product 0 (product 1: + 1 where the subindex is 255,
    //   get_nearest_times_2: row 0 at the next input sample); shift 255 marks a group the packed walk cannot take.
    struct GroupRef { int base; unsigned s0p, dp; };
    auto shift_of = [&](const GroupRef& r, int a, int p) { return (int)((r.dp >> (8 * a)) & 255u) + (p && ((r.s0p >> (8 * a)) & 255u) == 255u ? 1 : 0); };
    auto ref_ok = [&](const GroupRef& r) {
        int dmax = 0;
#pragma unroll
        for (int a = 0; a < A; a++) dmax = max(dmax, shift_of(r, a, 1));
        return SINC_LEN + dmax <= MLEN && dmax <= kG3PadL;
    };
    // raw plan entries of a group: the four window starts and the four subindices packed as GroupRef::s0p packs them — one
    // 128-bit and one 32-bit load when the group is whole and aligned, clamped scalar loads at ragged ends
    auto load_ref = [&](int ph0, int (&i0v)[A], unsigned& s0p) {
        static_assert(A == 4, "vector form of the plan loads");
        if (ph0 + A <= min(geo.period, nt) && ((reinterpret_cast<uintptr_t>(pi0 + ph0) & 15) | (reinterpret_cast<uintptr_t>(ps0 + ph0) & 3)) == 0) {
            const int4 v = __ldg(reinterpret_cast<const int4*>(pi0 + ph0));
            i0v[0] = v.x; i0v[1] = v.y; i0v[2] = v.z; i0v[3] = v.w;
            s0p = __ldg(reinterpret_cast<const unsigned*>(ps0 + ph0));
        } else {
            s0p = 0;
#pragma unroll
            for (int a = 0; a < A; a++) {
                const int nr = min(min(ph0 + a, geo.period - 1), nt - 1);
                i0v[a] = pi0[nr];
                s0p |= (unsigned)ps0[nr] << (8 * a);
            }
        }
    };
    auto make_ref = [&](const int (&i0v)[A], unsigned s0p, GroupRef& r) {
        const int fix = kLanes ? ((i0v[0] - w0) & 3) : 0;
        r.base = i0v[0] - fix; r.s0p = s0p; r.dp = 0;
#pragma unroll
        for (int a = 0; a < A; a++) r.dp |= (unsigned)min(255, max(0, i0v[a] - i0v[0] + fix)) << (8 * a);
    };
    // stage the 8 shifted rows of a group (warp private): one TMA bulk copy per row from the pre-shifted table copies (16-byte
    // aligned slices of MLEN words), issued by lane 0 and completed on the warp's mbarrier — ~40 instructions in one lane
    // instead of ~350 in every lane for 16-byte cp.async loops.  The caller has synchronised the warp after its last read
    // of the rows.
    auto stage_taps = [&](const GroupRef& r) {
        if (!ref_ok(r) || q != 0) return;
        mbar_expect_tx(tap_bar, kG3Rows * MLEN * (uint32_t)sizeof(float));
#pragma unroll
        for (int a = 0; a < A; a++)
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const int s0 = (int)((r.s0p >> (8 * a)) & 255u);
                const int row = p == 0 ? s0 : ((s0 + 1) & (OVERSAMPLING - 1));
                const int d = shift_of(r, a, p), sh = d & 3;
                const float* grow = table4 + ((size_t)(sh * OVERSAMPLING + row)) * kG3RowLen + (kG3PadL - (d - sh));
                tma_load_1d(smem_u32(T + (a * 2 + p) * TS), grow, MLEN * (uint32_t)sizeof(float), tap_bar);
            }
    };

    // The group loop is software pipelined: while group g is convolved, the reference entries of the warp's next group are in
    // flight; its tap rows are requested as soon as the walk has read the last word of the current rows, so they arrive under
    // the exchange / store phase, and the frac values the finish step needs are loaded before the walk.
    const int groups = min((geo.period + A - 1) / A, (nt + A - 1) / A);  // groups that exist in period 0 (later periods start later still)
    GroupRef cur = {0, 0u, 0u};
    if (warp < groups) {
        int i0v[A];
        unsigned s0p;
        load_ref(warp * A, i0v, s0p);
        make_ref(i0v, s0p, cur);
        stage_taps(cur);
    }
    for (int g = warp; g < groups; g += kG3Warps) {
        const int ph0 = g * A;
        const bool have_next = g + kG3Warps < groups;
        int i0n[A] = {0, 0, 0, 0};  // the next group's raw plan entries: requested now, consumed after the walk (five registers across it)
        unsigned s0n = 0;
        constexpr bool kCarry = kLanes;  // period columns hold per-column offsets across the walk: no registers to spare there
        if (have_next) {
            if constexpr (kCarry) load_ref((g + kG3Warps) * A, i0n, s0n);
            else { const int nr = min((g + kG3Warps) * A, nt - 1); prefetch_l1(pi0 + nr); prefetch_l1(ps0 + nr); }
        }
        const bool cur_ok = ref_ok(cur);
        const int fix = (int)(cur.dp & 255u);  // shift of output 0, product 0 == the alignment fix

        // x base of each column for this group (frames of the staged window); period columns: does the column follow the
        // reference pattern?  (lane columns share every plan entry: always)
        int xoff[kLanes ? 1 : B];
        unsigned follows = 0;  // bit = slot
        if (kLanes) {
            xoff[0] = cur.base - w0;
            const bool f = cur_ok && xoff[0] >= 0 && xoff[0] + MLEN <= count;
            if (!f) xoff[0] = 0;
            follows = f ? 0xffu : 0u;
        } else {
#pragma unroll
            for (int b = 0; b < B; b++) {
                bool f = cur_ok;
                int ibase = cur.base;
                const int rb = col_first(b) + ph0;
                const int nv = ((col_ok >> b) & 1u) ? max(0, min(A, min(geo.period - ph0, nt - rb))) : 0;
#pragma unroll
                for (int a = 0; a < A; a++) {
                    if (a < nv) {
                        const int rel = pi0[rb + a] - ((int)((cur.dp >> (8 * a)) & 255u) - fix);  // implied window start of output 0 of this column
                        if (a == 0) ibase = rel;
                        f = f && rel == ibase && (unsigned)ps0[rb + a] == ((cur.s0p >> (8 * a)) & 255u);
                    }
                }
                xoff[b] = ibase - w0;
                f = f && xoff[b] >= 0 && xoff[b] + MLEN <= count;
                if (!f) xoff[b] = 0;  // any staged window: the result is discarded
                follows |= (f ? 1u : 0u) << b;
            }
        }
        // the finish step's frac values (slots 0, 1): into L1 ahead of the walk
#pragma unroll
        for (int bb = 0; bb < (kLanes ? 1 : 2); bb++) prefetch_l1(pfrac + min(col_first(bb) + ph0, nt - 1));
        unsigned rot_odd = 0;  // bit a * 2 + p: parity of the accumulator-lane rotation of that product
#pragma unroll
        for (int a = 0; a < A; a++)
#pragma unroll
            for (int p = 0; p < 2; p++) rot_odd |= (unsigned)(shift_of(cur, a, p) & 1) << (a * 2 + p);

        if (cur_ok) { mbar_wait(tap_bar, tap_phase); tap_phase ^= 1u; }  // this group's rows
        __syncwarp();
        unsigned long long acc[A][B][2][NP];  // [output][slot][product][pair of this thread's 4 (2: quad form) virtual lanes]
        if (cur_ok) {
            // ---- the packed FIR walk over this half-warp's words
            const float* xb[B];
#pragma unroll
            for (int b = 0; b < B; b++) xb[b] = xcol[b] + xoff[kLanes ? 0 : b];
            const float* Th = T + HW * hh;
#pragma unroll
            for (int i = 0; i < A * B * 2 * NP; i++) (&acc[0][0][0][0])[i] = 0ull;
#pragma unroll(kLanes ? g3_walk_unroll<MLEN>() : MLEN / 8)
            for (int m = 0; m < MLEN; m += 8) {
                unsigned long long xp[B][NP];
#pragma unroll
                for (int b = 0; b < B; b++) {
                    if constexpr (kQuad) {
                        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(xp[b][0]) : "r"(smem_u32(xb[b] + m)));
                    } else {
                        xp[b][0] = pack2(xb[b][m], xb[b][m + 1]);
                        xp[b][NP - 1] = pack2(xb[b][m + 2], xb[b][m + 3]);
                    }
                }
#pragma unroll
                for (int a = 0; a < A; a++)
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        unsigned long long t[NP];
                        if constexpr (kQuad) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(t[0]) : "r"(smem_u32(Th + (a * 2 + p) * TS + m)));
                        else asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(t[0]), "=l"(t[NP - 1]) : "r"(smem_u32(Th + (a * 2 + p) * TS + m)));
#pragma unroll
                        for (int b = 0; b < B; b++)
#pragma unroll
                            for (int k = 0; k < NP; k++) ffma2(acc[a][b][p][k], xp[b][k], t[k]);
                    }
            }
        }
        // ---- the rows are consumed: request the next group's while this one is finished
        __syncwarp();
        if (have_next) {
            if constexpr (!kCarry) load_ref((g + kG3Warps) * A, i0n, s0n);
            make_ref(i0n, s0n, cur);
            stage_taps(cur);
        }
        // the finish step's frac values: requested now (L1 hits after the prefetch), consumed after the exchange
        float frv[A][kLanes ? 1 : 2];
#pragma unroll
        for (int bb = 0; bb < (kLanes ? 1 : 2); bb++) {
            const int rb = col_first(bb) + ph0;
#pragma unroll
            for (int a = 0; a < A; a++) frv[a][bb] = pfrac[min(rb + a, nt - 1)];
        }
        float* const dst_row[2] = {s_col_dst[col_lane(0)], s_col_dst[col_lane(1)]};  // requested ahead of the exchange, used by the stores
        float P0[A][2], P1[A][2];  // the two products of the 4 outputs of the 2 columns this thread finishes (slots 0, 1)
#pragma unroll
        for (int a = 0; a < A; a++)
#pragma unroll
            for (int bb = 0; bb < 2; bb++) { P0[a][bb] = 0.0f; P1[a][bb] = 0.0f; }
        if constexpr (kQuad) {
            if (cur_ok) {
                // ---- exchange and reduce, quad form.  Step 1 (partner lane ^ 16 = quarter hh ^ 2, which holds virtual lanes
                // v_{i+4} of this thread's v_i): send slots 4-7, receive the partner's 4-7 = this thread's slots 0-3, leaving
                // (s_i, s_{i+1}) = (v_i + v_{i+4}, v_{i+1} + v_{i+5}), i = 2 (hh & 1).  Step 2 (partner lane ^ 8 = quarter hh ^ 1,
                // which holds the other two s of the same columns): send slots 2-3, receive the partner's = this thread's
                // slots 0-1.  rubato's tree adds s pairs (0,1), (2,3) for an even rotation and (1,2), (3,0) for an odd one;
                // with own (m0, m1) and received (r0, r1) that is (m0 + m1) + (r0 + r1) or (m1 + r0) + (r1 + m0) for either
                // value of hh & 1, because every addition commutes bit for bit.
#pragma unroll
                for (int a = 0; a < A; a++)
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        float s[4][2];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const unsigned long long recv = __shfl_xor_sync(0xffffffffu, acc[a][4 + k][p][0], 16);
                            float m0, m1, r0, r1;
                            unpack2(acc[a][k][p][0], m0, m1);
                            unpack2(recv, r0, r1);
                            s[k][0] = __fadd_rn(m0, r0); s[k][1] = __fadd_rn(m1, r1);
                        }
                        const bool odd = ((rot_odd >> (a * 2 + p)) & 1u) != 0;
#pragma unroll
                        for (int bb = 0; bb < 2; bb++) {
                            const float r0 = __shfl_xor_sync(0xffffffffu, s[2 + bb][0], 8), r1 = __shfl_xor_sync(0xffffffffu, s[2 + bb][1], 8);
                            const float m0 = s[bb][0], m1 = s[bb][1];
                            const float t0 = odd ? __fadd_rn(m1, r0) : __fadd_rn(m0, m1);
                            const float t1 = odd ? __fadd_rn(r1, m0) : __fadd_rn(r0, r1);
                            const float r = __fadd_rn(t0, t1);
                            if (p == 0) P0[a][bb] = r; else P1[a][bb] = r;
                        }
                    }
            }
        } else if (cur_ok) {
            // ---- exchange and reduce: this thread holds virtual lanes 4 hh .. 4 hh + 3 of all four slots and finishes slots 0, 1;
            // its partner (lane ^ 16) holds the other four lanes and finishes this thread's slots 2, 3 (its own 0, 1).
            // s_i = v_i + v_{i+4} pairs one value of each half and additions commute, so the halves need no ordering.
#pragma unroll
            for (int a = 0; a < A; a++)
#pragma unroll
                for (int p = 0; p < 2; p++)
#pragma unroll
                    for (int bb = 0; bb < 2; bb++) {
                        float s4[4];
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const unsigned long long recv = __shfl_xor_sync(0xffffffffu, acc[a][2 + bb][p][k & (NP - 1)], 16);
                            float m0, m1, r0, r1;
                            unpack2(acc[a][bb][p][k & (NP - 1)], m0, m1);
                            unpack2(recv, r0, r1);
                            s4[2 * k] = __fadd_rn(m0, r0); s4[2 * k + 1] = __fadd_rn(m1, r1);
                        }
                        // (a_l + a_{l+4}) = s4[(l + rot) & 3]; the tree adds pairs (0,1) and (2,3) of those: adjacent pairs starting at rot & 3
                        const bool odd = ((rot_odd >> (a * 2 + p)) & 1u) != 0;
                        const float t0 = odd ? __fadd_rn(s4[1], s4[2]) : __fadd_rn(s4[0], s4[1]);
                        const float t1 = odd ? __fadd_rn(s4[3], s4[0]) : __fadd_rn(s4[2], s4[3]);
                        const float r = __fadd_rn(t0, t1);
                        if (p == 0) P0[a][bb] = r; else P1[a][bb] = r;
                    }
        }

        // ---- finish slots 0, 1: interp_lin with each output's own frac; a non-finite result is recomputed with the per-output
        // arithmetic (an Inf / NaN input word that met a zero-padded tap must not leak into an output whose true window excludes
        // it), as are columns that do not follow the reference pattern; one 16-byte store per column
#pragma unroll
        for (int bb = 0; bb < 2; bb++) {
            const int rb = col_first(bb) + ph0;
            const int nv = ((col_ok >> bb) & 1u) ? max(0, min(A, min(geo.period - ph0, nt - rb))) : 0;
            if (nv > 0) {
                const bool f = ((follows >> bb) & 1u) != 0;
                const int cl = col_lane(bb);
                const float* wc = win + (size_t)cl * geo.col_stride;
                float y[A];
                unsigned redo = 0;  // bit a: output a goes to the per-output arithmetic (rare: one branch for the whole column)
#pragma unroll
                for (int a = 0; a < A; a++) {
                    y[a] = __fadd_rn(P0[a][bb], __fmul_rn(frv[a][kLanes ? 0 : bb], __fsub_rn(P1[a][bb], P0[a][bb])));  // interp_lin
                    redo |= ((a < nv && (!f || g3_nonfinite(y[a]))) ? 1u : 0u) << a;
                }
                if (redo) {
#pragma unroll
                    for (int a = 0; a < A; a++)
                        if ((redo >> a) & 1u) y[a] = g3_exact(wc, w0, pi0, ps0, rb + a, table, frv[a][kLanes ? 0 : bb]);
                    atomicAdd(stats, (unsigned long long)__popc(redo));
                }
                float* op = dst_row[bb] + rb;
                if (nv == A && ((((uintptr_t)op) & 15) == 0)) {
                    *reinterpret_cast<float4*>(op) = make_float4(y[0], y[1], y[2], y[3]);
                } else {
#pragma unroll
                    for (int a = 0; a < A; a++)
                        if (a < nv) op[a] = y[a];
                }
            }
        }
    }
}

// ================================================================================================
// opt-in tolerance mode (sk_ctx_set_resample_mode(SK_RESAMPLE_FAST)): one product per output
// ================================================================================================
// rubato evaluates out = p0 + frac (p1 - p0) with two 256-tap products against table rows s and s + 1.  Algebraically that is ONE
// product against the interpolated row h = h_s + frac (h_{s+1} - h_s).  With lane columns every stream of the batch shares (s, frac)
// of output n, so the warp builds the 8 interpolated rows of its group once (8 x 272 fused multiply-adds) and the walk needs half
// the FFMA2 and half the tap traffic per output; the freed accumulators double the outputs per thread (8 consecutive outputs x 4
// columns, still 128 accumulator registers), which halves the input traffic per output as well.  The result is NOT bit-identical
// to rubato (the table interpolation is rounded once per tap instead of once per output, the lane sums are taken in a fixed
// order): it stays within the north-star contract — max-abs 1e-5 of full scale, SNR >= 120 dB against the bit-exact kernel —
// which tests/test_gpu_baseline_lengths.py asserts (measured: ~1e-7, > 135 dB).  Lane columns only (>= 64 lanes); anything
// else, and any non-finite result, takes the bit-exact paths.
constexpr int kG3FA = 8;  // consecutive outputs per thread in the fast kernel

template <int MLEN>
__global__ void __launch_bounds__(kG3Threads, 1)
resample_gen3_fast_kernel(const __grid_constant__ SrcDesc src, const __grid_constant__ OutDesc dst, const __grid_constant__ Gen3Geom geo,
                          const int32_t* __restrict__ pi0, const uint8_t* __restrict__ ps0, const float* __restrict__ pfrac,
                          const float* __restrict__ table, const float* __restrict__ table4, int lanes_per_stream, unsigned long long* __restrict__ stats) {
    constexpr int A = kG3FA, B = kG3B;
    extern __shared__ __align__(16) float g3_smem[];
    __shared__ const uint8_t* s_col_src[kG3Cols];
    __shared__ float* s_col_dst[kG3Cols];
    constexpr int TS = MLEN;
    float* taps_all = g3_smem;                                   // [warps][8 rows][TS]
    float* win = g3_smem + kG3Warps * A * TS;                    // [64][col_stride]
    const int warp = threadIdx.x >> 5, q = threadIdx.x & 31;
    float* T = taps_all + warp * A * TS;

    const long long n0 = (long long)blockIdx.x * geo.period;
    const int nt = (int)(min(dst.n_out, n0 + geo.period) - n0);
    pi0 += n0; ps0 += n0; pfrac += n0;
    const int w0 = pi0[0] - (pi0[0] & 3);
    const long long lane0 = (long long)blockIdx.y * geo.C;
    const int count = min(geo.span, pi0[nt - 1] - w0 + SINC_LEN + 2 + (TS - SINC_LEN) + 8);
    g3_stage_tile(src, dst, geo, n0, lane0, w0, count, lanes_per_stream, win, s_col_src, s_col_dst);

    const int ql = q & 15, hh = q >> 4;
    auto col_lane = [&](int b) { return (ql + 16 * (b ^ (2 * hh))) & (kG3Cols - 1); };  // slot b -> lane of the CTA (own columns in slots 0, 1)
    const float* xcol[B];
#pragma unroll
    for (int b = 0; b < B; b++) xcol[b] = win + (size_t)col_lane(b) * geo.col_stride + 4 * hh;
    const float* Th = T + 4 * hh;

    const int groups = (nt + A - 1) / A;
    for (int g = warp; g < groups; g += kG3Warps) {
        const int ph0 = g * A;
        // plan entries of the group's 8 outputs (shared by every column); entries past the tile's end are clamped and masked at the store
        int i0r[A], s0r[A];
        float fr[A];
#pragma unroll
        for (int a = 0; a < A; a++) {
            const int nr = min(ph0 + a, nt - 1);
            i0r[a] = pi0[nr]; s0r[a] = ps0[nr]; fr[a] = pfrac[nr];
        }
        const int fix = (i0r[0] - w0) & 3;
        const int xoff = i0r[0] - fix - w0;
        int dmax = 0;
#pragma unroll
        for (int a = 0; a < A; a++) dmax = max(dmax, i0r[a] - i0r[0] + fix + (s0r[a] == OVERSAMPLING - 1 ? 1 : 0));
        const bool ok = SINC_LEN + dmax <= MLEN && dmax <= kG3PadL && xoff >= 0 && xoff + MLEN <= count;

        float y[A];
#pragma unroll
        for (int a = 0; a < A; a++) y[a] = __int_as_float(0x7fc00000);  // NaN: a group the walk cannot take goes to the exact path below
        if (ok) {
            // ---- the 8 interpolated, shifted, zero-padded rows of this group (warp private).  Both source rows come from the
            // word-shifted table copies as 16-byte aligned slices; the accumulators are not live yet, so half the group's loads
            // (4 outputs x 3 vectors x 2 rows = 24 LDG.128) are in flight at once.
            __syncwarp();
            constexpr int NV = (MLEN / 4 + 31) / 32;  // float4 vectors per lane and row
#pragma unroll
            for (int half = 0; half < 2; half++) {
                float4 h0v[4][NV], h1v[4][NV];
#pragma unroll
                for (int aa = 0; aa < 4; aa++) {
                    const int a = 4 * half + aa;
                    const int d0 = i0r[a] - i0r[0] + fix, d1 = d0 + (s0r[a] == OVERSAMPLING - 1 ? 1 : 0);
                    const float* g0 = table4 + ((size_t)((d0 & 3) * OVERSAMPLING + s0r[a])) * kG3RowLen + (kG3PadL - (d0 - (d0 & 3)));
                    const float* g1 = table4 + ((size_t)((d1 & 3) * OVERSAMPLING + ((s0r[a] + 1) & (OVERSAMPLING - 1)))) * kG3RowLen + (kG3PadL - (d1 - (d1 & 3)));
#pragma unroll
                    for (int v = 0; v < NV; v++) {
                        const int i = q + 32 * v;
                        if (i < MLEN / 4) { h0v[aa][v] = __ldg(reinterpret_cast<const float4*>(g0) + i); h1v[aa][v] = __ldg(reinterpret_cast<const float4*>(g1) + i); }
                    }
                }
#pragma unroll
                for (int aa = 0; aa < 4; aa++) {
                    const int a = 4 * half + aa;
                    float* t = T + a * TS;
#pragma unroll
                    for (int v = 0; v < NV; v++) {
                        const int i = q + 32 * v;
                        if (i < MLEN / 4) {
                            const float4 x0 = h0v[aa][v], x1 = h1v[aa][v];
                            reinterpret_cast<float4*>(t)[i] = make_float4(__fmaf_rn(fr[a], __fsub_rn(x1.x, x0.x), x0.x), __fmaf_rn(fr[a], __fsub_rn(x1.y, x0.y), x0.y),
                                                                          __fmaf_rn(fr[a], __fsub_rn(x1.z, x0.z), x0.z), __fmaf_rn(fr[a], __fsub_rn(x1.w, x0.w), x0.w));
                        }
                    }
                }
            }
            __syncwarp();

            // ---- the packed walk over this half-warp's words: 8 outputs x 4 columns x 4 virtual lanes
            const float* xb[B];
#pragma unroll
            for (int b = 0; b < B; b++) xb[b] = xcol[b] + xoff;
            unsigned long long acc[A][B][2];
#pragma unroll
            for (int i = 0; i < A * B * 2; i++) (&acc[0][0][0])[i] = 0ull;
#pragma unroll(g3_walk_unroll<MLEN>())
            for (int m = 0; m < MLEN; m += 8) {
                unsigned long long xp[B][2];
#pragma unroll
                for (int b = 0; b < B; b++)
                    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(xp[b][0]), "=l"(xp[b][1]) : "r"(smem_u32(xb[b] + m)));
#pragma unroll
                for (int a = 0; a < A; a++) {
                    unsigned long long t0, t1;
                    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(t0), "=l"(t1) : "r"(smem_u32(Th + a * TS + m)));
#pragma unroll
                    for (int b = 0; b < B; b++) {
                        ffma2(acc[a][b][0], xp[b][0], t0);
                        ffma2(acc[a][b][1], xp[b][1], t1);
                    }
                }
            }
            __syncwarp();
            // ---- exchange: this thread finishes slots 0, 1; its partner (lane ^ 16) holds the other four lanes of them
#pragma unroll
            for (int a = 0; a < A; a++) {
                float v[2];
#pragma unroll
                for (int bb = 0; bb < 2; bb++) {
                    float s4[4];
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const unsigned long long recv = __shfl_xor_sync(0xffffffffu, acc[a][2 + bb][k], 16);
                        float m0, m1, r0, r1;
                        unpack2(acc[a][bb][k], m0, m1);
                        unpack2(recv, r0, r1);
                        s4[2 * k] = __fadd_rn(m0, r0); s4[2 * k + 1] = __fadd_rn(m1, r1);
                    }
                    v[bb] = __fadd_rn(__fadd_rn(s4[0], s4[1]), __fadd_rn(s4[2], s4[3]));
                }
                // outputs of slot 0 stay in y[], slot 1's go out right away below: keep both
                y[a] = v[0];
                fr[a] = v[1];  // (fr is dead after the row staging: reuse its registers for slot 1)
            }
        } else {
#pragma unroll
            for (int a = 0; a < A; a++) fr[a] = __int_as_float(0x7fc00000);
        }

        // ---- finish slots 0, 1: a non-finite result is recomputed with the bit-exact per-output arithmetic
        const int nv = min(A, nt - ph0);
#pragma unroll
        for (int bb = 0; bb < 2; bb++) {
            const int cl = col_lane(bb);
            if (lane0 + cl >= geo.n_lanes) continue;
            const float* wc = win + (size_t)cl * geo.col_stride;
            float o[A];
            int redo = 0;
#pragma unroll
            for (int a = 0; a < A; a++) {
                o[a] = bb == 0 ? y[a] : fr[a];
                if (a < nv && g3_nonfinite(o[a])) { o[a] = g3_exact(wc, w0, pi0, ps0, ph0 + a, table, pfrac[ph0 + a]); redo++; }
            }
            if (redo) atomicAdd(stats, (unsigned long long)redo);
            float* op = s_col_dst[cl] + ph0;
            if (nv == A && ((((uintptr_t)op) & 15) == 0)) {
                *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(op + 4) = make_float4(o[4], o[5], o[6], o[7]);
            } else {
#pragma unroll
                for (int a = 0; a < A; a++)
                    if (a < nv) op[a] = o[a];
            }
        }
    }
}

template <int MLEN>
static int launch_gen3_fast_t(sk_ctx* ctx, const SrcDesc& s2, const OutDesc& d2, const Gen3Geom& g2, const sk_plan* plan, const float* d_table,
                              const float* d_table4, int lanes_per_stream, dim3 grid, int smem) {
    SK_CUDA(ctx, cudaFuncSetAttribute(resample_gen3_fast_kernel<MLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    resample_gen3_fast_kernel<MLEN><<<grid, kG3Threads, smem, ctx->stream>>>(s2, d2, g2, plan->d_i0, plan->d_s0, plan->d_frac, d_table, d_table4,
                                                                           lanes_per_stream, ctx->d_stats);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

// ================================================================================================
// host: geometry and launch
// ================================================================================================
template <int XVEC, int MLEN>
static int launch_gen3_t(sk_ctx* ctx, const SrcDesc& s2, const OutDesc& d2, const Gen3Geom& g2, const sk_plan* plan, const float* d_table,
                         const float* d_table4, int lanes_per_stream, dim3 grid, int smem) {
    SK_CUDA(ctx, cudaFuncSetAttribute(resample_gen3_kernel<XVEC, MLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    resample_gen3_kernel<XVEC, MLEN><<<grid, kG3Threads, smem, ctx->stream>>>(s2, d2, g2, plan->d_i0, plan->d_s0, plan->d_frac, d_table, d_table4,
                                                                            lanes_per_stream, ctx->d_stats);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

int sk_get_shifted_sinc_table(sk_ctx* ctx, uint32_t in_rate, uint32_t out_rate, int pad_l, int row_len, const float** d_table4);  // sk_resample.cu

// returns SK_OK after a launch, 1 when this kernel does not apply to the plan / batch (the caller falls back to gen2 / gen)
int sk_launch_gen3(sk_ctx* ctx, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* d_table, uint32_t in_rate, uint32_t out_rate,
                   size_t n_streams, int lanes_per_stream) {
    const long long n_lanes = (long long)n_streams * lanes_per_stream;
    if (n_lanes <= 0 || !plan->n_out || !plan->d_i0 || plan->max_step3 < 0) return 1;
    const double t_ratio = (double)in_rate / (double)out_rate;
    Gen3Geom g;
    memset(&g, 0, sizeof(g));
    g.n_lanes = n_lanes;
    const bool lanes_only = n_lanes >= kG3Cols;
    const bool fast = ctx->resample_fast && lanes_only && plan->max_step7 >= 0 && SINC_LEN + plan->max_step7 + 1 + 3 <= 288;
    // largest tap shift inside a group: window advance over 3 outputs (exact, from the plan) + the subindex-255 shift + the
    // alignment fix of the LDS.128 form
    const int dmax = fast ? plan->max_step7 + 1 + 3 : plan->max_step3 + 1 + (lanes_only ? 3 : 0);
    if (dmax > (fast ? kG3PadL : 16)) return 1;
    const int mlen = fast ? (SINC_LEN + dmax <= 272 ? 272 : 288) : (SINC_LEN + dmax <= 264 ? 264 : (SINC_LEN + dmax <= 272 ? 272 : 0));
    if (!mlen) return 1;                                        // steep downsampling ratios: the padded walk would waste too much; gen2 / gen serve them
    g.tap_stride = mlen;
    const int smem_budget = 224 * 1024;
    const int taps_bytes = kG3Warps * kG3Rows * g.tap_stride * (int)sizeof(float);
    const int over = (g.tap_stride - SINC_LEN) + 8 + 4;         // overshoot of the padded walk past the last real window word
    if (lanes_only) {
        // lane columns only: no period structure needed; outputs per tile as large as the windows allow
        g.C = kG3Cols; g.P = 1; g.xvec = 1;
        int T = 256;
        for (; T >= 32; T -= 32) {
            const int span = (int)ceil((double)(T - 1) * t_ratio) + 2 + SINC_LEN + 2 + over;
            int stride = (span + 3) & ~3;
            if ((stride & 7) == 0) stride += 4;                  // stride / 4 odd: the 8 lanes of an LDS.128 quarter-warp hit 8 distinct bank quads
            if (taps_bytes + kG3Cols * stride * (int)sizeof(float) <= smem_budget) { g.period = T; g.span = span; g.col_stride = stride; break; }
        }
        if (T < 32) return 1;
    } else {
        // fewer than 64 lanes: fill the columns with whole phase periods of the same lanes
        if (plan->period2 <= 0) return 1;
        int C = 1;
        while (C * 2 <= n_lanes) C *= 2;
        g.C = C; g.P = kG3Cols / C; g.period = plan->period2; g.xvec = 0;
        if ((size_t)g.period * g.P > plan->n_out) return 1;      // shorter than one tile: gen2 / gen handle it
        const int span = (int)ceil((double)((long long)g.period * g.P - 1) * t_ratio) + 2 + SINC_LEN + 2 + over;
        g.span = span;
        // input frames per period (exact for true periods): the bank of column (c, k) is c * stride + k * M' (mod 32)
        const uint32_t gg = gcd_u32(in_rate, out_rate);
        const long long Mper = (long long)(in_rate / gg) * (g.period / (int)(out_rate / gg));
        int stride = span | 1;
        if (C < 32) {
            const int want = (int)(((32 / C) * (Mper % 32)) % 32);  // stride == (32 / C) * M' (mod 32): 32 consecutive columns hit 32 banks when M' is odd
            while ((stride % 32) != want) stride++;
        }
        g.col_stride = stride;
        if (taps_bytes + C * stride * (int)sizeof(float) > smem_budget) return 1;
    }
    const int smem = taps_bytes + g.C * g.col_stride * (int)sizeof(float);
    for (g.log2c = 0; (1 << g.log2c) < g.C; g.log2c++) {}
    if (int rc = sk_ensure_stats(ctx)) return rc;
    const float* d_table4 = nullptr;
    if (int rc = sk_get_shifted_sinc_table(ctx, in_rate, out_rate, kG3PadL, kG3RowLen, &d_table4)) return rc;
    const long long tile_out = (long long)g.period * g.P;
    const long long tiles = ((long long)plan->n_out + tile_out - 1) / tile_out;
    const long long lane_blocks = (n_lanes + g.C - 1) / g.C;
    if (tiles > 0x7fffffffll) return 1;
    for (long long b0 = 0; b0 < lane_blocks; b0 += 65535) {
        const long long nb = std::min<long long>(65535, lane_blocks - b0);
        SrcDesc s2 = src; OutDesc d2 = dst; Gen3Geom g2 = g;
        if (b0) {  // only reachable with more than 65535 * C lanes: the slab must start on a stream boundary
            const long long first_lane = b0 * g.C;
            if (first_lane % lanes_per_stream) SK_FAIL(ctx, SK_E_ARG, "batch too large for one call (%lld lanes); split it", n_lanes);
            const long long stream0 = first_lane / lanes_per_stream;
            s2.base += stream0 * src.stream_stride;
            d2.base += stream0 * dst.stream_stride;
            g2.n_lanes = n_lanes - first_lane;
        }
        const dim3 grid((unsigned)tiles, (unsigned)nb);
        int rc;
        if (fast) rc = mlen == 272 ? launch_gen3_fast_t<272>(ctx, s2, d2, g2, plan, d_table, d_table4, lanes_per_stream, grid, smem)
                                   : launch_gen3_fast_t<288>(ctx, s2, d2, g2, plan, d_table, d_table4, lanes_per_stream, grid, smem);
        else if (g.xvec) rc = mlen == 264 ? launch_gen3_t<8, 264>(ctx, s2, d2, g2, plan, d_table, d_table4, lanes_per_stream, grid, smem)
                                     : launch_gen3_t<8, 272>(ctx, s2, d2, g2, plan, d_table, d_table4, lanes_per_stream, grid, smem);
        else rc = mlen == 264 ? launch_gen3_t<1, 264>(ctx, s2, d2, g2, plan, d_table, d_table4, lanes_per_stream, grid, smem)
                              : launch_gen3_t<1, 272>(ctx, s2, d2, g2, plan, d_table, d_table4, lanes_per_stream, grid, smem);
        if (rc) return rc;
    }
    return SK_OK;
}
