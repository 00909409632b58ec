// sk_resample_gen3.cu — general-ratio polyphase FIR, third form: 4 outputs x 2 columns per thread, packed FFMA2, shifted tap rows.
//
// Same arithmetic contract as resample_gen2_kernel (sk_resample.cu): every output is rubato's
//     p0 + frac * (p1 - p0),   p_k = 8 lane accumulators (lane l owns taps l, l+8, ... with fused multiply-adds in tap order)
//                                    reduced as ((a0+a4)+(a1+a5)) + ((a2+a6)+(a3+a7))
// so results are bit-identical to the AVX path of rubato 0.14.1 `SincFixedIn` as soundkit calls it
// (soundkit/src/audio_pipeline.rs:474-491, soundkit-decoder/src/lib.rs:1939-2031).
//
// Why a third form.  The 2 x 2 tile of gen2 spends 13-14 shared-memory wavefronts per 32 FFMA (LSU-bound at ~0.57 of the FMA
// peak, measured 0.28).  Here a thread owns A = 4 CONSECUTIVE outputs (n .. n+3) of B = 2 COLUMNS:
//   * a column is an independent input window that shares the group's table rows: another (stream, channel) lane of the batch
//     and/or the same lane one or more whole phase periods later.  With >= 64 lanes in a batch (BASELINE configs[3]: 4096
//     streams x 2 channels) columns are simply lanes, every plan entry is shared exactly and the period structure of the
//     ratio is not needed at all; smaller batches fill the 64 columns with periods of the same lane (the gen2 idea);
//   * the windows of 4 consecutive outputs overlap almost completely, so ONE pass over the input words x[m] serves all four:
//     output a needs tap j = m - d_a, where d_a is its window offset inside the group.  The warp stages its 8 table rows
//     (4 outputs x 2 interpolation products) in shared memory SHIFTED by d_a and zero padded: T_a[m] = h_a[m - d_a].  Every
//     output then consumes the same (x[m], x[m+1]) register pair against (T_a[m], T_a[m+1]) with one packed FFMA2;
//   * accumulators are indexed by m & 7 ("virtual lanes", compile time).  Rubato's lane of tap j is j & 7 = (m - d_a) & 7, a
//     rotation by d_a & 7 that is undone in the reduction: the tree pairs lane l with l + 4 — a rotation-invariant pairing —
//     and then sums adjacent pairs, which depends on the parity of the rotation only (two selects per product).
//     Zero-padded taps add x * 0 to an accumulator that can never be -0: an exact no-op for finite x; tiles that stage
//     non-finite or huge (>= 2^113) floats take the per-output path instead.
// Per 4 input words and warp: 2 LDS.128 of inputs + 8 broadcast LDS.128 of taps feed 32 FFMA2 (64 FMA lanes-pairs):
// 10 shared-memory instructions / 16 wavefronts per 32 FFMA2 instead of 13-14 per 32 scalar FFMA.
//
// Lanes whose own plan entries differ from the group's reference entries (possible only for period columns, where the f64
// index walk lands within rounding of a table-row boundary for some periods but not others) recompute their outputs with the
// per-output arithmetic afterwards; sk_ctx_resample_fallback_outputs() counts them.
#include "sk_resample_dev.cuh"

constexpr int kG3Threads = 256;
constexpr int kG3Warps = kG3Threads / 32;
constexpr int kG3A = 4;   // consecutive outputs per thread
constexpr int kG3B = 2;   // columns per thread
constexpr int kG3Cols = 32 * kG3B;
constexpr int kG3Rows = 2 * kG3A;

struct Gen3Geom {
    int C;            // lanes (stream x channel) per CTA: power of two <= 64
    int P;            // periods per tile = 64 / C
    int period;       // outputs per period; P == 1: outputs per tile (any multiple of 4 * warps)
    int col_stride;   // words between the windows of consecutive lanes in shared memory
    int span;         // words staged per lane (window of the whole tile + overshoot of the padded tap walk)
    int tap_stride;   // words per staged (shifted, zero padded) table row; the tap walk covers m in [0, tap_stride)
    int xvec;         // 1: lane windows and group bases are 16-byte aligned -> LDS.128 input loads (P == 1 only)
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

template <int XVEC, int MLEN>
__global__ void __launch_bounds__(kG3Threads, 1)
resample_gen3_kernel(const __grid_constant__ SrcDesc src, const __grid_constant__ OutDesc dst, const __grid_constant__ Gen3Geom geo,
                     const int32_t* __restrict__ pi0, const uint8_t* __restrict__ ps0, const float* __restrict__ pfrac,
                     const float* __restrict__ table, int lanes_per_stream, unsigned long long* __restrict__ stats) {
    constexpr int A = kG3A, B = kG3B;
    extern __shared__ __align__(16) float g3_smem[];
    __shared__ int s_big;
    constexpr int TS = MLEN;  // words per staged row == length of the (fully unrolled) tap walk
    float* taps_all = g3_smem;                                   // [warps][8 rows][TS]
    float* win = g3_smem + kG3Warps * kG3Rows * TS;              // [C][col_stride]
    const int warp = threadIdx.x >> 5, q = threadIdx.x & 31;
    float* T = taps_all + warp * kG3Rows * TS;

    const long long tile_out = (long long)geo.period * geo.P;
    const long long n0 = (long long)blockIdx.x * tile_out;
    const long long n1 = min(dst.n_out, n0 + tile_out) - 1;  // last output of this tile
    const int w0 = pi0[n0] - (XVEC == 4 ? (pi0[n0] & 3) : 0);  // first staged frame (16-byte aligned window origin for LDS.128)
    const long long lane0 = (long long)blockIdx.y * geo.C;

    // frames staged per lane: the tile's windows plus the overshoot of the padded tap walk (zero taps meet these words: they must
    // be staged, finite data)
    const int count = min(geo.span, pi0[n1] - w0 + SINC_LEN + 2 + (TS - SINC_LEN) + 8);

    // ---- stage the windows: warp w takes lanes w, w + 8, ...; threads walk the frames
    if (threadIdx.x == 0) s_big = 0;
    __syncthreads();
    {
        bool big = false;
        for (int c = warp; c < geo.C; c += kG3Warps) {
            const long long lane = lane0 + c;
            float* wc = win + (size_t)c * geo.col_stride;
            if (lane >= geo.n_lanes) {
                for (int i = q; i < count; i += 32) wc[i] = 0.0f;
                continue;
            }
            const long long stream = lane / lanes_per_stream;
            const int ch = (int)(lane % lanes_per_stream);
            if (src.width == 0) {
                const float* p = reinterpret_cast<const float*>(src.base + stream * src.stream_stride + (long long)ch * src.lane_stride);
                for (int i = q; i < count; i += 32) {
                    const long long f = (long long)w0 + i;
                    const float v = (f >= src.valid_lo && f < src.valid_hi) ? __ldg(p + f + src.origin) : 0.0f;
                    big |= (__float_as_uint(v) & 0x7f800000u) >= 0x78000000u;
                    wc[i] = v;
                }
            } else {
                const uint8_t* sbase = src.base + stream * src.stream_stride;
                for (int i = q; i < count; i += 32) {
                    const long long f = (long long)w0 + i;
                    float v = 0.0f;
                    if (f >= src.valid_lo && f < src.valid_hi) v = decode_sample_bits(load_sample_bits(sbase, f, ch, src), src);
                    if (src.is_float) big |= (__float_as_uint(v) & 0x7f800000u) >= 0x78000000u;
                    wc[i] = v;
                }
            }
        }
        if (big) s_big = 1;
    }
    __syncthreads();
    const bool tile_big = s_big != 0;

    // ---- this thread's two columns: column j = q + 32 b -> (lane c = j % C, period k = j / C)
    int col_c[B];
    long long col_n0[B];   // first output of the column's period
    bool col_ok[B];
#pragma unroll
    for (int b = 0; b < B; b++) {
        const int j = q + 32 * b;
        col_c[b] = j & (geo.C - 1);
        col_n0[b] = n0 + (long long)(j / geo.C) * geo.period;
        col_ok[b] = lane0 + col_c[b] < geo.n_lanes;
    }
    const int groups = (geo.period + A - 1) / A;
    for (int g = warp; g < groups; g += kG3Warps) {
        const int ph0 = g * A;
        if (n0 + ph0 > n1) break;  // nothing of this group (or any later one) exists in period 0; later periods start later still
        // reference plan entries of the group (period 0 of the tile), clamped at ragged ends (clamped entries are masked below)
        int i0r[A], s0r[A];
#pragma unroll
        for (int a = 0; a < A; a++) {
            long long nr = n0 + min(ph0 + a, geo.period - 1);
            if (nr > n1) nr = n1;
            i0r[a] = pi0[nr];
            s0r[a] = ps0[nr];
        }
        const int fix = XVEC == 4 ? ((i0r[0] - w0) & 3) : 0;  // move the group base down to a 16-byte boundary, taps shift up instead
        int dsh[A][2];
        int dmax = 0;
#pragma unroll
        for (int a = 0; a < A; a++) {
            dsh[a][0] = i0r[a] - i0r[0] + fix;
            dsh[a][1] = dsh[a][0] + (s0r[a] == OVERSAMPLING - 1 ? 1 : 0);  // get_nearest_times_2: row 0 at the next input sample
            dmax = max(dmax, dsh[a][1]);
        }
        constexpr int mlen = MLEN;
        const bool group_ok = !tile_big && SINC_LEN + dmax <= MLEN;

        // own plan entries: valid outputs of a column are a < nv[b]
        int nv[B], xoff[B];
        bool follows[B];
        float fr[A][B];
#pragma unroll
        for (int b = 0; b < B; b++) {
            const long long nb = col_n0[b] + ph0;
            long long left = min((long long)(geo.period - ph0), n1 - nb + 1);
            nv[b] = col_ok[b] ? (int)max(0ll, min((long long)A, left)) : 0;
            follows[b] = group_ok;
            int ibase = i0r[0];
#pragma unroll
            for (int a = 0; a < A; a++) {
                fr[a][b] = 0.0f;
                if (a < nv[b]) {
                    const int i0o = pi0[nb + a], s0o = ps0[nb + a];
                    fr[a][b] = pfrac[nb + a];
                    const int rel = i0o - (i0r[a] - i0r[0]);  // implied window start of output 0 of this column
                    if (a == 0) ibase = rel;
                    follows[b] = follows[b] && rel == ibase && s0o == s0r[a];
                }
            }
            xoff[b] = ibase - w0 - fix;  // x base of the column for this group, in frames of the staged window
            follows[b] = follows[b] && xoff[b] >= 0 && xoff[b] + mlen <= count;
            if (!follows[b]) xoff[b] = 0;  // any staged window: the result is discarded
        }

        float y[A][B];
#pragma unroll
        for (int a = 0; a < A; a++)
#pragma unroll
            for (int b = 0; b < B; b++) y[a][b] = 0.0f;

        if (group_ok) {
            // ---- stage the 8 shifted rows of this group (warp private)
            __syncwarp();
#pragma unroll
            for (int a = 0; a < A; a++)
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const int row = p == 0 ? s0r[a] : ((s0r[a] + 1) & (OVERSAMPLING - 1));
                    const float* h = table + row * SINC_LEN;
                    float* t = T + (a * 2 + p) * TS;
                    const int d = dsh[a][p];
                    for (int m = q; m < mlen; m += 32) {
                        const int j = m - d;
                        t[m] = (j >= 0 && j < SINC_LEN) ? __ldg(h + j) : 0.0f;
                    }
                }
            __syncwarp();

            // ---- the packed FIR walk
            const float* xb[B];
#pragma unroll
            for (int b = 0; b < B; b++) xb[b] = win + (size_t)col_c[b] * geo.col_stride + xoff[b];
            unsigned long long acc[A][B][2][4];
#pragma unroll
            for (int i = 0; i < A * B * 8; i++) (&acc[0][0][0][0])[i] = 0ull;
#pragma unroll
            for (int m = 0; m < MLEN; m += 8) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int mm = m + 4 * h;
                    unsigned long long xp[B][2];
#pragma unroll
                    for (int b = 0; b < B; b++) {
                        if (XVEC == 4) {
                            asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(xp[b][0]), "=l"(xp[b][1]) : "r"(smem_u32(xb[b] + mm)));
                        } else {
                            xp[b][0] = pack2(xb[b][mm], xb[b][mm + 1]);
                            xp[b][1] = pack2(xb[b][mm + 2], xb[b][mm + 3]);
                        }
                    }
#pragma unroll
                    for (int a = 0; a < A; a++)
#pragma unroll
                        for (int p = 0; p < 2; p++) {
                            unsigned long long t0, t1;
                            asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(t0), "=l"(t1) : "r"(smem_u32(T + (a * 2 + p) * TS + mm)));
#pragma unroll
                            for (int b = 0; b < B; b++) {
                                ffma2(acc[a][b][p][2 * h], xp[b][0], t0);
                                ffma2(acc[a][b][p][2 * h + 1], xp[b][1], t1);
                            }
                        }
                }
            }
#pragma unroll
            for (int a = 0; a < A; a++)
#pragma unroll
                for (int b = 0; b < B; b++) {
                    const float p0 = reduce8_rotated(acc[a][b][0], dsh[a][0] & 7), p1 = reduce8_rotated(acc[a][b][1], dsh[a][1] & 7);
                    y[a][b] = __fadd_rn(p0, __fmul_rn(fr[a][b], __fsub_rn(p1, p0)));  // interp_lin with the output's own frac
                }
        }

        // ---- columns that do not follow the reference pattern (and whole groups / tiles on the exact slow path): per-output arithmetic
#pragma unroll
        for (int b = 0; b < B; b++) {
            if (nv[b] == 0 || follows[b]) continue;
            const long long nb = col_n0[b] + ph0;
#pragma unroll
            for (int a = 0; a < A; a++) {
                if (a >= nv[b]) continue;
                const int i0o = pi0[nb + a], s0o = ps0[nb + a];
                int s1 = s0o + 1, d = 0;
                if (s1 >= OVERSAMPLING) { s1 -= OVERSAMPLING; d = 1; }
                y[a][b] = gen3_single(win + (size_t)col_c[b] * geo.col_stride + (i0o - w0), table + s0o * SINC_LEN, table + s1 * SINC_LEN, d, fr[a][b]);
            }
            atomicAdd(stats, (unsigned long long)nv[b]);
        }

        // ---- store: 4 consecutive outputs per column
#pragma unroll
        for (int b = 0; b < B; b++) {
            if (nv[b] == 0) continue;
            const long long nb = col_n0[b] + ph0;
            const long long lane = lane0 + col_c[b];
            float* op = dst.base + (lane / lanes_per_stream) * dst.stream_stride + (lane % lanes_per_stream) * dst.lane_stride + nb;
            if (nv[b] == A && ((((uintptr_t)op) & 15) == 0)) {
                *reinterpret_cast<float4*>(op) = make_float4(y[0][b], y[1][b], y[2][b], y[3][b]);
            } else {
#pragma unroll
                for (int a = 0; a < A; a++)
                    if (a < nv[b]) op[a] = y[a][b];
            }
        }
    }
}

// ================================================================================================
// host: geometry and launch
// ================================================================================================
template <int XVEC, int MLEN>
static int launch_gen3_t(sk_ctx* ctx, const SrcDesc& s2, const OutDesc& d2, const Gen3Geom& g2, const sk_plan* plan, const float* d_table, int lanes_per_stream,
                         dim3 grid, int smem) {
    SK_CUDA(ctx, cudaFuncSetAttribute(resample_gen3_kernel<XVEC, MLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    resample_gen3_kernel<XVEC, MLEN><<<grid, kG3Threads, smem, ctx->stream>>>(s2, d2, g2, plan->d_i0, plan->d_s0, plan->d_frac, d_table, lanes_per_stream,
                                                                            ctx->d_stats);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

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
    // largest tap shift inside a group: window advance over 3 outputs (exact, from the plan) + the subindex-255 shift + the
    // alignment fix of the LDS.128 form
    const int dmax = plan->max_step3 + 1 + (lanes_only ? 3 : 0);
    const int mlen = SINC_LEN + dmax <= 264 ? 264 : (SINC_LEN + dmax <= 272 ? 272 : 0);
    if (!mlen) return 1;                                        // steep downsampling ratios: the padded walk would waste too much; gen2 / gen serve them
    g.tap_stride = mlen;
    const int smem_budget = 200 * 1024;
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
    if (int rc = sk_ensure_stats(ctx)) return rc;
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
        if (g.xvec) rc = mlen == 264 ? launch_gen3_t<4, 264>(ctx, s2, d2, g2, plan, d_table, lanes_per_stream, grid, smem)
                                     : launch_gen3_t<4, 272>(ctx, s2, d2, g2, plan, d_table, lanes_per_stream, grid, smem);
        else rc = mlen == 264 ? launch_gen3_t<1, 264>(ctx, s2, d2, g2, plan, d_table, lanes_per_stream, grid, smem)
                              : launch_gen3_t<1, 272>(ctx, s2, d2, g2, plan, d_table, lanes_per_stream, grid, smem);
        if (rc) return rc;
    }
    return SK_OK;
}
