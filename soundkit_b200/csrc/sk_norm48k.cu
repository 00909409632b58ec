// sk_norm48k.cu — SURVEY.md 8(f) row 2: soundkit's StreamingStereo48kNormalizer on the device.
//
// Replaces soundkit/src/audio_pipeline.rs:133-413: decoder-convention sample conversion (pcm_sample_to_f32,
// :176-227), mono duplicated / first two channels (:155-166), finite_pcm (:396-402) and the 2-tap linear
// interpolation to 48 kHz over the whole-stream timeline (:326-347) with the repeated-final-sample tail (:366-394).
//
// Numerics (bit-exact to the reference):
//   * source_position = output_frames as f64 * f64::from(rate) / 48_000.0 is evaluated per output with the same two
//     IEEE f64 operations (__dmul_rn, __ddiv_rn); `lower` is its floor, `fraction = (pos - lower) as f32`;
//   * out = lower + ((upper - lower) * fraction) with separate sub / mul / add (file built with -fmad=false);
//   * integer samples scale by exact powers of two, finite_pcm is applied to inputs and outputs.
// HBM bound: per output frame 8 B written + 2 * channels * width * rate / 48000 read.  One thread owns 4 consecutive
// output frames (two 128-bit stores) and gathers the few source frames they touch through L1; the f64 position
// arithmetic is done once per thread and reused for every stream of its grid.y slice.
#include <math.h>

#include <algorithm>
#include <new>

#include "sk_common.cuh"

struct N48Src {
    const uint8_t* base;      // first byte of stream 0
    long long stream_stride;  // bytes between streams
    int width;                // sample bytes 2 / 3 / 4
    int channels;
    int big_endian, is_float;
    unsigned long long start_frame;  // global index of the buffer's first frame (streaming push)
    unsigned long long frames;       // frames in the buffer
};

__device__ __forceinline__ float n48_finite_pcm(float s) {  // audio_pipeline.rs:396-402
    if ((__float_as_uint(s) & 0x7f800000u) == 0x7f800000u) return 0.0f;
    return s < -1.0f ? -1.0f : (s > 1.0f ? 1.0f : s);
}

// sample `ch` of local frame `f` -> f32 (pcm_sample_to_f32, :176-227) -> finite_pcm
__device__ __forceinline__ float n48_sample(const uint8_t* sbase, const N48Src& s, unsigned long long f, int ch) {
    const uint8_t* p = sbase + (f * (unsigned long long)s.channels + ch) * (unsigned long long)s.width;
    float v;
    if (s.width == 2) {
        uint32_t u = *reinterpret_cast<const uint16_t*>(p);
        if (s.big_endian) u = ((u & 0xff) << 8) | (u >> 8);
        v = __fmul_rn((float)(int16_t)u, 1.0f / 32768.0f);
    } else if (s.width == 3) {
        const uint32_t b0 = p[0], b1 = p[1], b2 = p[2];
        const uint32_t u = s.big_endian ? (b2 | (b1 << 8) | (b0 << 16)) : (b0 | (b1 << 8) | (b2 << 16));
        v = __fmul_rn((float)(((int32_t)(u << 8)) >> 8), 1.0f / 8388608.0f);
    } else {
        uint32_t u = *reinterpret_cast<const uint32_t*>(p);
        if (s.big_endian) u = bswap32(u);
        v = s.is_float ? __uint_as_float(u) : __fmul_rn(__int2float_rn((int32_t)u), 4.656612873077393e-10f);
    }
    return n48_finite_pcm(v);
}

constexpr int kN48Threads = 256;

// out_l/out_r: [stream][n_total] planes (out_stride floats between streams).  Outputs [0, n_lerp) are interpolated
// (global output index o_start + i), outputs [n_lerp, n_total) repeat the stream's final sample (finish()).
// prev: value of global frame start_frame - 1 (streaming push only; ignored when start_frame == 0).
// final_lr (optional): [stream][2] <- the last frame of each stream's buffer (the next push's `previous`).
__global__ void __launch_bounds__(kN48Threads)
norm48k_kernel(const __grid_constant__ N48Src src, unsigned long long o_start, unsigned long long n_lerp, unsigned long long n_total,
               unsigned rate, float prev_l, float prev_r, float* __restrict__ out_l, float* __restrict__ out_r, long long out_stride,
               unsigned n_streams, float* __restrict__ final_lr) {
    const unsigned long long i0 = ((unsigned long long)blockIdx.x * kN48Threads + threadIdx.x) * 4ull;
    const bool writes_final = final_lr != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    if (i0 >= n_total && !writes_final) return;
    const bool need_last = writes_final || i0 + 4 > n_lerp;  // tail outputs repeat the stream's final frame
    // positions of the (up to) 4 outputs this thread owns: shared by every stream
    long long lo[4];   // local frame index of `lower` (-1 = the previous push's final frame)
    float fr[4];
    const bool passthrough = rate == 48000u;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned long long o = o_start + i0 + k;
        if (passthrough) { lo[k] = (long long)(o - src.start_frame); fr[k] = 0.0f; continue; }
        const double pos = __ddiv_rn(__dmul_rn((double)o, (double)rate), 48000.0);
        const double fl = floor(pos);
        lo[k] = (long long)((unsigned long long)fl - src.start_frame);  // >= -1 by construction of n_lerp
        fr[k] = __double2float_rn(__dsub_rn(pos, fl));
    }
    const int c1 = src.channels > 1 ? 1 : 0;
    for (unsigned s = blockIdx.y; s < n_streams; s += gridDim.y) {
        const uint8_t* sbase = src.base + (long long)s * src.stream_stride;
        float last_l = prev_l, last_r = prev_r;
        if (need_last && src.frames) { last_l = n48_sample(sbase, src, src.frames - 1, 0); last_r = n48_sample(sbase, src, src.frames - 1, c1); }
        if (writes_final) { final_lr[2 * s] = last_l; final_lr[2 * s + 1] = last_r; }
        if (i0 >= n_total) continue;
        float ol[4], orr[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k >= n_lerp) { ol[k] = last_l; orr[k] = last_r; continue; }  // finish(): vec![previous; remaining]
            const long long f = lo[k];
            const float ll = f < 0 ? prev_l : n48_sample(sbase, src, (unsigned long long)f, 0);
            const float rl = f < 0 ? prev_r : n48_sample(sbase, src, (unsigned long long)f, c1);
            if (passthrough) { ol[k] = ll; orr[k] = rl; continue; }
            const float lu = n48_sample(sbase, src, (unsigned long long)(f + 1), 0);
            const float ru = n48_sample(sbase, src, (unsigned long long)(f + 1), c1);
            ol[k] = n48_finite_pcm(__fadd_rn(ll, __fmul_rn(__fsub_rn(lu, ll), fr[k])));
            orr[k] = n48_finite_pcm(__fadd_rn(rl, __fmul_rn(__fsub_rn(ru, rl), fr[k])));
        }
        float* pl = out_l + (long long)s * out_stride + i0;
        float* pr = out_r + (long long)s * out_stride + i0;
        if (i0 + 4 <= n_total && (((uintptr_t)pl | (uintptr_t)pr) & 15) == 0) {
            st_stream_v4(pl, make_uint4(__float_as_uint(ol[0]), __float_as_uint(ol[1]), __float_as_uint(ol[2]), __float_as_uint(ol[3])));
            st_stream_v4(pr, make_uint4(__float_as_uint(orr[0]), __float_as_uint(orr[1]), __float_as_uint(orr[2]), __float_as_uint(orr[3])));
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (i0 + k < n_total) { pl[k] = ol[k]; pr[k] = orr[k]; }
        }
    }
}

// Tiled form (the fast path): a CTA owns 256 x OPT consecutive outputs of one stream at a time.  The source bytes those
// outputs touch are one contiguous run, copied into shared memory with coalesced 128-bit loads (the gathers of the
// direct kernel above become shared-memory reads).  Positions AND the byte offsets of the frames they select are
// computed once per thread and reused for every stream of the CTA's grid.y slice, and the sample format is a template
// parameter, so the per-stream work of an output is 4 shared-memory loads, 4 conversions and the lerp.
// Outputs per thread and stream: one barrier, one mbarrier wait and one round of pointer arithmetic per OPT outputs.  Measured on
// s16 stereo 44.1 kHz (B200): 4 -> 0.86, 8 -> 0.93, 16 -> 0.96 of the HBM copy rate.  16 whenever a tile's run fits the per-CTA
// shared-memory budget, 4 for the widest formats / highest rates.
constexpr int kN48OptWide = 16, kN48OptNarrow = 4;
constexpr int kN48RingBytes = 48 << 10;   // shared memory a CTA aims to spend on its ring of staged runs: 3 runs in flight for s16 stereo.  Deeper is
                                          // slower (measured: 6 stages 0.90, 3 stages 0.96 of the HBM rate; 8 vs 4 stages at 4 outputs per thread 0.86 vs 0.88)

// W sample bytes at shared-memory address p (2- / 4-byte aligned for W = 2 / 4, any alignment for W = 3)
template <int W>
__device__ __forceinline__ uint32_t n48_lds(const uint8_t* p) {
    if (W == 2) return *reinterpret_cast<const uint16_t*>(p);
    if (W == 4) return *reinterpret_cast<const uint32_t*>(p);
    const uint32_t* q = reinterpret_cast<const uint32_t*>((uintptr_t)p & ~(uintptr_t)3);
    return __funnelshift_r(q[0], q[1], 8 * (int)((uintptr_t)p & 3));  // staged run has 16 spare bytes behind it
}
template <int W, bool BE, bool FLT>
__device__ __forceinline__ float n48_cvt(uint32_t u) {  // pcm_sample_to_f32 (:176-227) + finite_pcm
    if (W == 2) {
        if (BE) u = __byte_perm(u, 0, 0x3201);
        return __fmul_rn((float)(int16_t)(u & 0xffffu), 1.0f / 32768.0f);  // always finite, |x| <= 1
    }
    if (W == 3) {
        const uint32_t v = BE ? __byte_perm(u, 0, 0x0124) : __byte_perm(u, 0, 0x2104);  // sample in bytes 1..3
        return __fmul_rn((float)(((int32_t)v) >> 8), 1.0f / 8388608.0f);
    }
    if (BE) u = bswap32(u);
    if (FLT) return n48_finite_pcm(__uint_as_float(u));
    return __fmul_rn(__int2float_rn((int32_t)u), 4.656612873077393e-10f);
}

// one frame (left, right) at shared-memory address p.  ST = the stream has >= 2 channels (else right = left).
template <int W, bool BE, bool FLT, bool ST>
__device__ __forceinline__ void n48_frame(const uint8_t* p, float& l, float& r) {
    if (W == 2) {
        if (ST) {  // both channels in one aligned 32-bit word when the frame is 4 bytes; two halfword loads otherwise
            const uint32_t lo = *reinterpret_cast<const uint16_t*>(p), hi = *reinterpret_cast<const uint16_t*>(p + 2);
            l = n48_cvt<2, BE, FLT>(lo); r = n48_cvt<2, BE, FLT>(hi);
        } else { l = r = n48_cvt<2, BE, FLT>(*reinterpret_cast<const uint16_t*>(p)); }
    } else {
        l = n48_cvt<W, BE, FLT>(n48_lds<W>(p));
        r = ST ? n48_cvt<W, BE, FLT>(n48_lds<W>(p + W)) : l;
    }
}
__device__ __forceinline__ float n48_clamp1(float x) { return fminf(fmaxf(x, -1.0f), 1.0f); }  // finite_pcm of a finite value

// Integer formats, fast loop: the frame as UNSCALED floats (the integer value).  The exact power-of-two scale commutes with
// every rounding of `lo + ((up - lo) * frac)`, so it is applied once to the result instead of to the four inputs, and the
// result of interpolating between values in [-1, 1] needs no clamp.  fb4: the frame is one aligned 32-bit word (s16 stereo).
template <int W, bool BE, bool ST>
__device__ __forceinline__ void n48_frame_raw(const uint8_t* p, bool fb4, float& l, float& r) {
    if (W == 2) {
        uint32_t u;
        if (ST && fb4) u = *reinterpret_cast<const uint32_t*>(p);
        else u = *reinterpret_cast<const uint16_t*>(p) | (ST ? (uint32_t)*reinterpret_cast<const uint16_t*>(p + 2) << 16 : 0u);
        if (BE) u = __byte_perm(u, 0, 0x2301);
        // s16 -> float without I2F.S16 (a conversion-unit instruction that shares the memory-instruction queue with the LDS / STG
        // of this loop: 17 % of the stall samples): s ^ 0x8000 = s + 32768 as the mantissa of 2^23, minus (2^23 + 32768), exact
        l = __fsub_rn(__uint_as_float((u & 0xffffu) ^ 0x4B008000u), 8421376.0f);
        r = ST ? __fsub_rn(__uint_as_float((u >> 16) ^ 0x4B008000u), 8421376.0f) : l;
    } else if (W == 3) {
        const uint32_t a = n48_lds<3>(p);
        l = (float)(((int32_t)(BE ? __byte_perm(a, 0, 0x0124) : __byte_perm(a, 0, 0x2104))) >> 8);
        if (ST) { const uint32_t b = n48_lds<3>(p + 3); r = (float)(((int32_t)(BE ? __byte_perm(b, 0, 0x0124) : __byte_perm(b, 0, 0x2104))) >> 8); }
        else r = l;
    } else {
        uint32_t a = *reinterpret_cast<const uint32_t*>(p);
        if (BE) a = bswap32(a);
        l = __int2float_rn((int32_t)a);
        if (ST) { uint32_t b = *reinterpret_cast<const uint32_t*>(p + 4); if (BE) b = bswap32(b); r = __int2float_rn((int32_t)b); }
        else r = l;
    }
}
// Two consecutive packed s24 stereo frames (the lower and upper frame of one output: 12 bytes at any alignment) as four UNSCALED
// sample values: 4 aligned word loads, 3 funnel shifts and 3 byte permutes, instead of a pair of loads and a funnel shift per
// sample (s24 stereo ran at 0.70 of the HBM rate against 0.93-0.98 for the 16- and 32-bit formats: the loop is bound by its
// shared-memory instructions).  The staged run has 16 spare bytes behind it.
template <bool BE>
__device__ __forceinline__ void n48_s24_stereo_pair(const uint8_t* p, float& ll, float& rl, float& lu, float& ru) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>((uintptr_t)p & ~(uintptr_t)3);
    const int sh = 8 * (int)((uintptr_t)p & 3);
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
    const uint32_t u0 = __funnelshift_r(w0, w1, sh), u1 = __funnelshift_r(w1, w2, sh), u2 = __funnelshift_r(w2, w3, sh);  // bytes 0-3, 4-7, 8-11
    // every sample lands in the top three bytes of a word (its low byte is don't-care), then one arithmetic shift
    const uint32_t a = BE ? __byte_perm(u0, 0, 0x0122) : __byte_perm(u0, 0, 0x2100);     // L0 = bytes 0, 1, 2
    const uint32_t b = BE ? __byte_perm(u0, u1, 0x3455) : __byte_perm(u0, u1, 0x5433);   // R0 = bytes 3, 4, 5
    const uint32_t c = BE ? __byte_perm(u1, u2, 0x2344) : __byte_perm(u1, u2, 0x4322);   // L1 = bytes 6, 7, 8
    const uint32_t d = BE ? __byte_perm(u2, 0, 0x1233) : u2;                              // R1 = bytes 9, 10, 11
    ll = (float)(((int32_t)a) >> 8); rl = (float)(((int32_t)b) >> 8);
    lu = (float)(((int32_t)c) >> 8); ru = (float)(((int32_t)d) >> 8);
}
template <int W> __device__ __forceinline__ float n48_scale() { return W == 2 ? 1.0f / 32768.0f : (W == 3 ? 1.0f / 8388608.0f : 4.656612873077393e-10f); }

constexpr int kN48MaxStages = 4;         // staged runs in flight per CTA (ring of shared-memory buffers)
constexpr int kN48MaxBufBytes = 48 << 10;  // largest run the tiled kernel stages (two stages of it still leave 2 CTAs per SM)

__device__ __forceinline__ uint32_t n48_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void n48_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void n48_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void n48_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void n48_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tN48_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra N48_DONE;\n\tbra N48_WAIT;\n\tN48_DONE:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// Outputs of a tile are dealt to threads round-robin (thread t owns outputs t, t + 256, t + 512, ... of the tile):
// consecutive lanes read consecutive source frames (conflict-free shared-memory reads) and write consecutive floats
// (one 128-byte line per warp store).  The staged runs form a ring of `stages` buffers filled by TMA bulk copies
// (cp.async.bulk -> mbarrier, issued by one thread `stages - 1` streams ahead): no register staging, no shared-memory
// stores, and the global latency of a run is covered by the interpolation of the streams before it (history: a one-deep
// register prefetch left 39 % of the stall samples on long_scoreboard and 20 % on the barrier: 0.79 of the HBM rate).
template <int W, bool BE, bool FLT, bool ST, bool FB4, int OPT>  // FB4: a frame is exactly one aligned 32-bit word (16-bit stereo)
__global__ void __launch_bounds__(kN48Threads)
norm48k_tile_kernel(const __grid_constant__ N48Src src, unsigned long long o_start, unsigned long long n_lerp, unsigned long long n_total,
                    unsigned rate, float prev_l, float prev_r, float* __restrict__ out_l, float* __restrict__ out_r, long long out_stride,
                    unsigned n_streams, float* __restrict__ final_lr, int buf_bytes, int stages) {
    extern __shared__ __align__(16) uint8_t n48_raw[];
    constexpr int kN48Tile = kN48Threads * OPT, kN48Opt = OPT;
    __shared__ __align__(8) unsigned long long n48_bars[kN48MaxStages];
    const unsigned long long t0 = (unsigned long long)blockIdx.x * kN48Tile;   // first output of the tile
    const bool writes_final = final_lr != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    const bool passthrough = rate == 48000u;
    auto lower_of = [&](unsigned long long o, float* frac) -> long long {
        if (passthrough) { if (frac) *frac = 0.0f; return (long long)(o - src.start_frame); }
        const double pos = __ddiv_rn(__dmul_rn((double)o, (double)rate), 48000.0);
        const double fl = floor(pos);
        if (frac) *frac = __double2float_rn(__dsub_rn(pos, fl));
        return (long long)((unsigned long long)fl - src.start_frame);
    };
    // source frames of the tile's interpolated outputs: [f_lo, f_hi] (local indices, clipped to the buffer)
    const unsigned long long t_last = min(n_lerp, t0 + kN48Tile);  // one past the tile's last interpolated output
    long long f_lo = 0, f_hi = -1, first_lower = 0;
    if (t_last > t0) {
        first_lower = lower_of(o_start + t0, nullptr);
        f_lo = max(0ll, first_lower);
        f_hi = min((long long)src.frames - 1, lower_of(o_start + t_last - 1, nullptr) + (passthrough ? 0 : 1));
    }
    // the common tile: every output interpolated from frames inside this buffer -> no per-output special cases
    const bool plain = t0 + kN48Tile <= n_lerp && first_lower >= 0 && !writes_final && t0 + kN48Tile <= n_total;
    const int fb = src.channels * src.width;
    int off[kN48Opt];   // byte offset of the output's lower frame from frame f_lo; -1: the carried previous frame; -2: tail output
    float fr[kN48Opt];
#pragma unroll
    for (int k = 0; k < kN48Opt; k++) {
        const unsigned long long i = t0 + (unsigned)(k * kN48Threads) + threadIdx.x;
        const long long f = lower_of(o_start + i, &fr[k]);
        off[k] = (i >= n_lerp) ? -2 : (f < 0 ? -1 : (int)(f - f_lo) * fb);
    }
    bool need_last = writes_final;
#pragma unroll
    for (int k = 0; k < kN48Opt; k++) need_last |= off[k] == -2 && t0 + (unsigned)(k * kN48Threads) + threadIdx.x < n_total;

    // ---- staging: the run [f_lo, f_hi] of a stream as 16-byte vectors around it.  Base and stride are 16-byte aligned on
    // this path, so the run's offset inside a stream, its length and its skew are the same for every stream.
    const long long g_off = f_lo * fb;                        // byte offset of frame f_lo inside a stream
    const long long a_off = g_off & ~15ll;                    // ... rounded down to a vector
    const int skew = (int)(g_off - a_off);
    const int run = f_hi >= f_lo ? (int)((f_hi + 1) * fb - a_off) : 0;
    const long long safe = (long long)src.frames * fb - a_off;  // never read past the caller's buffer
    const int nvec = (int)min((long long)((run + 15) >> 4), safe >> 4);
    const uint32_t run_bytes = (uint32_t)nvec * 16u;
    const unsigned n_it = blockIdx.y < n_streams ? (n_streams - blockIdx.y + gridDim.y - 1) / gridDim.y : 0;
    auto issue = [&](unsigned it, int stage) {  // one thread: request the run of the CTA's it-th stream into ring slot `stage`
        const unsigned s = blockIdx.y + it * gridDim.y;
        const uint32_t bar = n48_smem_u32(&n48_bars[stage]);
        n48_mbar_expect_tx(bar, run_bytes);
        if (run_bytes) n48_bulk_load(n48_smem_u32(n48_raw + (size_t)stage * buf_bytes), src.base + (long long)s * src.stream_stride + a_off, run_bytes, bar);
    };
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; i++) n48_mbar_init(n48_smem_u32(&n48_bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < stages - 1 && (unsigned)i < n_it; i++) issue((unsigned)i, i);
    }
    int cur = 0, nxt = stages - 1;  // ring slot of this iteration / of the request made in it
    uint32_t phase = 0;
    unsigned s = blockIdx.y;
    for (unsigned it = 0; it < n_it; it++, s += gridDim.y) {
        uint8_t* buf = n48_raw + (size_t)cur * buf_bytes;
        if ((int)run_bytes < run) {  // the bytes after the last whole vector the caller's buffer holds (last tile of a stream only)
            const uint8_t* a0 = src.base + (long long)s * src.stream_stride + a_off;
            for (int b = (int)run_bytes + threadIdx.x; b < run; b += kN48Threads) buf[b] = a0[b];
        }
        __syncthreads();  // mbarriers initialised (first trip); every thread is done with the previous stream: its slot may be refilled
        if (threadIdx.x == 0 && it + (unsigned)(stages - 1) < n_it) issue(it + (unsigned)(stages - 1), nxt);
        n48_mbar_wait(n48_smem_u32(&n48_bars[cur]), phase);
        if (++cur == stages) { cur = 0; phase ^= 1u; }
        if (++nxt == stages) nxt = 0;
        const uint8_t* sm = buf + skew;  // frame f_lo
        float* pl = out_l + (long long)s * out_stride + t0 + threadIdx.x;
        float* pr = out_r + (long long)s * out_stride + t0 + threadIdx.x;
        if (plain && !FLT) {
            constexpr bool fb4 = FB4;
#pragma unroll
            for (int k = 0; k < kN48Opt; k++) {
                float ll, rl, lu, ru;
                if (W == 3 && ST && fb == 6 && !passthrough) {  // packed s24 stereo: both frames of the output from one 16-byte read
                    n48_s24_stereo_pair<BE>(sm + off[k], ll, rl, lu, ru);
                } else {
                    n48_frame_raw<W, BE, ST>(sm + off[k], fb4, ll, rl);
                    if (passthrough) { pl[k * kN48Threads] = __fmul_rn(ll, n48_scale<W>()); pr[k * kN48Threads] = __fmul_rn(rl, n48_scale<W>()); continue; }
                    n48_frame_raw<W, BE, ST>(sm + off[k] + fb, fb4, lu, ru);
                }
                pl[k * kN48Threads] = __fmul_rn(__fadd_rn(ll, __fmul_rn(__fsub_rn(lu, ll), fr[k])), n48_scale<W>());
                pr[k * kN48Threads] = __fmul_rn(__fadd_rn(rl, __fmul_rn(__fsub_rn(ru, rl), fr[k])), n48_scale<W>());
            }
        } else if (plain) {
#pragma unroll
            for (int k = 0; k < kN48Opt; k++) {
                float ll, rl, lu, ru;
                n48_frame<W, BE, FLT, ST>(sm + off[k], ll, rl);
                if (passthrough) { pl[k * kN48Threads] = ll; pr[k * kN48Threads] = rl; continue; }
                n48_frame<W, BE, FLT, ST>(sm + off[k] + fb, lu, ru);
                pl[k * kN48Threads] = n48_clamp1(__fadd_rn(ll, __fmul_rn(__fsub_rn(lu, ll), fr[k])));
                pr[k * kN48Threads] = n48_clamp1(__fadd_rn(rl, __fmul_rn(__fsub_rn(ru, rl), fr[k])));
            }
        } else {
            const uint8_t* sbase = src.base + (long long)s * src.stream_stride;
            float last_l = prev_l, last_r = prev_r;
            if (need_last && src.frames) { last_l = n48_sample(sbase, src, src.frames - 1, 0); last_r = n48_sample(sbase, src, src.frames - 1, src.channels > 1 ? 1 : 0); }
            if (writes_final) { final_lr[2 * s] = last_l; final_lr[2 * s + 1] = last_r; }
#pragma unroll
            for (int k = 0; k < kN48Opt; k++) {
                const unsigned long long i = t0 + (unsigned)(k * kN48Threads) + threadIdx.x;
                if (i >= n_total) continue;
                float o_l, o_r;
                if (off[k] == -2) { o_l = last_l; o_r = last_r; }  // finish(): vec![previous; remaining]
                else {
                    float ll = prev_l, rl = prev_r, lu, ru;
                    if (off[k] >= 0) n48_frame<W, BE, FLT, ST>(sm + off[k], ll, rl);
                    if (passthrough) { o_l = ll; o_r = rl; }
                    else {
                        n48_frame<W, BE, FLT, ST>(off[k] < 0 ? sm : sm + off[k] + fb, lu, ru);  // upper = lower + 1 (frame 0 after the carried one)
                        o_l = n48_clamp1(__fadd_rn(ll, __fmul_rn(__fsub_rn(lu, ll), fr[k])));
                        o_r = n48_clamp1(__fadd_rn(rl, __fmul_rn(__fsub_rn(ru, rl), fr[k])));
                    }
                }
                pl[k * kN48Threads] = o_l; pr[k * kN48Threads] = o_r;
            }
        }
    }
}

// ---- host arithmetic of the block / stream lengths (exact integer forms of the reference's f64 expressions) ----
// outputs a push may emit once `end` source frames are known: every o with floor(o*rate/48000) + 1 < end, i.e.
// o * rate < (end - 1) * 48000  ->  o < ceil((end - 1) * 48000 / rate).  (o * rate is exact in f64 below 2^53 and the
// quotient's fractional part is a multiple of 1/48000, far above half an ulp, so floor() of the f64 quotient equals
// the integer quotient.)
static unsigned long long n48_lerp_limit(unsigned long long end_frame, unsigned rate) {
    if (end_frame < 2) return 0;
    const unsigned __int128 num = (unsigned __int128)(end_frame - 1) * 48000u;
    return (unsigned long long)((num + rate - 1) / rate);
}
// finish(): ((source_frames as f64 * 48000.0) / rate as f64).round().max(1.0) as u64
static unsigned long long n48_target(unsigned long long frames, unsigned rate) {
    double t = round(((double)frames * 48000.0) / (double)rate);
    if (t < 1.0) t = 1.0;
    return (unsigned long long)t;
}

extern "C" size_t sk_normalize48k_out_frames(size_t frames, uint32_t rate) {
    if (!frames || !rate) return 0;
    if (rate == 48000u) return frames;
    const unsigned long long a = n48_lerp_limit(frames, rate), b = n48_target(frames, rate);
    return (size_t)(a > b ? a : b);
}

static int n48_fill_src(sk_ctx* ctx, N48Src* s, const void* in, size_t in_stride, int bits, int fmt, int big_endian, int channels,
                        unsigned long long start_frame, unsigned long long frames) {
    if (bits != 16 && bits != 24 && bits != 32) SK_FAIL(ctx, SK_E_ARG, "unsupported PCM bit depth %d", bits);
    if (channels <= 0) SK_FAIL(ctx, SK_E_ARG, "decoded audio contained no channels");
    if (bits != 24 && ((((uintptr_t)in) | in_stride) & (uintptr_t)(bits / 8 - 1))) SK_FAIL(ctx, SK_E_ARG, "input not aligned to the sample width");
    s->base = (const uint8_t*)in; s->stream_stride = (long long)in_stride;
    s->width = bits / 8; s->channels = channels; s->big_endian = big_endian != 0;
    s->is_float = (bits == 32 && fmt == SK_FMT_FLOAT);
    s->start_frame = start_frame; s->frames = frames;
    return SK_OK;
}

static int n48_launch(sk_ctx* ctx, const N48Src& src, unsigned long long o_start, unsigned long long n_lerp, unsigned long long n_total,
                      unsigned rate, float prev_l, float prev_r, float* out_l, float* out_r, long long out_stride, size_t n_streams,
                      float* final_lr) {
    // source bytes one tile of outputs can touch (+ alignment slack); the staged kernel needs them in shared memory
    auto span_bytes_of = [&](int opt) {
        const unsigned long long span_frames = ((unsigned long long)kN48Threads * opt * rate + 47999) / 48000 + 3;
        return span_frames * (unsigned long long)(src.width * src.channels) + 48;
    };
    const int opt = span_bytes_of(kN48OptWide) + 32 <= (unsigned long long)kN48MaxBufBytes ? kN48OptWide : kN48OptNarrow;
    const unsigned long long span_bytes = span_bytes_of(opt);
    const bool tiled = span_bytes + 32 <= (unsigned long long)kN48MaxBufBytes &&
                       (((uintptr_t)src.base | (uintptr_t)src.stream_stride) & 15) == 0;
    const unsigned long long groups = (n_total + 3) / 4;
    unsigned long long gx = (groups + kN48Threads - 1) / kN48Threads;
    if (gx < 1) gx = 1;
    if (gx > 0x7fffffffull) SK_FAIL(ctx, SK_E_ARG, "stream too long for one launch");
    // enough CTAs to fill the machine; beyond that, streams are looped inside the thread (position math amortised)
    unsigned gy = 1;
    const unsigned long long want = (unsigned long long)ctx->num_sms * 16;
    if (gx < want) gy = (unsigned)std::min<unsigned long long>(n_streams, (want + gx - 1) / gx);
    if (gy > 65535u) gy = 65535u;
    if (tiled) {
        const unsigned long long tile = (unsigned long long)kN48Threads * opt;
        gx = std::max<unsigned long long>(1, (n_total + tile - 1) / tile);
        gy = 1;
        if (gx < want) gy = (unsigned)std::min<unsigned long long>(n_streams, (want + gx - 1) / gx);
        if (gy > 65535u) gy = 65535u;
        const dim3 grid((unsigned)gx, gy);
        const int buf_bytes = (int)sk_round_up((size_t)span_bytes + 32, 16);
        // ring depth: as many runs in flight as the ring budget holds, between 2 and kN48MaxStages
        const int stages = (int)std::min<size_t>(kN48MaxStages, std::max<size_t>(2, (size_t)kN48RingBytes / (size_t)buf_bytes));
        const size_t smem = (size_t)stages * (size_t)buf_bytes;
#define N48_LAUNCH_OPT(W_, BE_, FLT_, ST_, FB4_, OPT_)                                                                                             \
        do {                                                                                                                                  \
            SK_CUDA(ctx, cudaFuncSetAttribute(norm48k_tile_kernel<W_, BE_, FLT_, ST_, FB4_, OPT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
            norm48k_tile_kernel<W_, BE_, FLT_, ST_, FB4_, OPT_><<<grid, kN48Threads, smem, ctx->stream>>>(src, o_start, n_lerp, n_total, rate, prev_l, prev_r, \
                                                                                              out_l, out_r, out_stride, (unsigned)n_streams, final_lr, buf_bytes, stages); \
        } while (0)
#define N48_LAUNCH(W_, BE_, FLT_, ST_, FB4_)                                                                                                        \
        do {                                                                                                                                  \
            if (opt == kN48OptWide) N48_LAUNCH_OPT(W_, BE_, FLT_, ST_, FB4_, kN48OptWide); else N48_LAUNCH_OPT(W_, BE_, FLT_, ST_, FB4_, kN48OptNarrow); \
        } while (0)
#define N48_LAUNCH_ST(W_, BE_, FLT_)                                                                  \
        do {                                                                                          \
            if (W_ == 2 && src.channels == 2) N48_LAUNCH(W_, BE_, FLT_, true, (W_ == 2));             \
            else if (src.channels > 1) N48_LAUNCH(W_, BE_, FLT_, true, false);                        \
            else N48_LAUNCH(W_, BE_, FLT_, false, false);                                             \
        } while (0)
        const bool be = src.big_endian != 0;
        if (src.width == 2) { if (be) N48_LAUNCH_ST(2, true, false); else N48_LAUNCH_ST(2, false, false); }
        else if (src.width == 3) { if (be) N48_LAUNCH_ST(3, true, false); else N48_LAUNCH_ST(3, false, false); }
        else if (src.is_float) { if (be) N48_LAUNCH_ST(4, true, true); else N48_LAUNCH_ST(4, false, true); }
        else { if (be) N48_LAUNCH_ST(4, true, false); else N48_LAUNCH_ST(4, false, false); }
#undef N48_LAUNCH_ST
#undef N48_LAUNCH
#undef N48_LAUNCH_OPT
    } else {
        norm48k_kernel<<<dim3((unsigned)gx, gy), kN48Threads, 0, ctx->stream>>>(src, o_start, n_lerp, n_total, rate, prev_l, prev_r, out_l, out_r,
                                                                               out_stride, (unsigned)n_streams, final_lr);
    }
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

// whole streams, device resident: normalize_blocks(&[whole]) of the reference tests for every stream of a batch
extern "C" int sk_normalize48k_batch_dev(sk_ctx* ctx, const void* in, size_t in_stride, size_t n_streams, size_t frames, int bits, int fmt,
                                         int big_endian, int channels, uint32_t rate, float* out, size_t out_stride, size_t out_stride_ch,
                                         size_t* n_out) {
    if (!ctx) return SK_E_ARG;
    if (rate == 0) SK_FAIL(ctx, SK_E_ARG, "decoded audio has invalid PCM geometry");
    N48Src src;
    int rc = n48_fill_src(ctx, &src, in, in_stride, bits, fmt, big_endian, channels, 0, frames);
    if (rc) return rc;
    const size_t total = sk_normalize48k_out_frames(frames, rate);
    if (n_out) *n_out = total;
    if (!total || !n_streams) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    if (n_streams > 0xffffffffull) SK_FAIL(ctx, SK_E_ARG, "too many streams");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned long long n_lerp = rate == 48000u ? frames : n48_lerp_limit(frames, rate);
    return n48_launch(ctx, src, 0, n_lerp, total, rate, 0.0f, 0.0f, out, out + out_stride_ch, (long long)out_stride, n_streams, nullptr);
}

// ---- stateful handle: StreamingStereo48kNormalizer (:256-394) -------------------------------------------------
struct sk_norm48k {
    sk_ctx* ctx;
    uint32_t rate = 0; int channels = 0;
    unsigned long long source_frames = 0, output_frames = 0;
    float prev_l = 0.0f, prev_r = 0.0f;
    bool has_previous = false, finished = false;
};

extern "C" int sk_norm48k_create(sk_ctx* ctx, sk_norm48k** out) {
    if (!ctx || !out) return SK_E_ARG;
    sk_norm48k* n = new (std::nothrow) sk_norm48k();
    if (!n) SK_FAIL(ctx, SK_E_ALLOC, "out of memory");
    n->ctx = ctx;
    *out = n;
    return SK_OK;
}
extern "C" void sk_norm48k_destroy(sk_norm48k* n) { delete n; }
extern "C" uint64_t sk_norm48k_source_frames(const sk_norm48k* n) { return n ? n->source_frames : 0; }
extern "C" uint64_t sk_norm48k_output_frames(const sk_norm48k* n) { return n ? n->output_frames : 0; }
extern "C" uint32_t sk_norm48k_source_sample_rate(const sk_norm48k* n) { return n ? n->rate : 0; }
extern "C" int sk_norm48k_source_channels(const sk_norm48k* n) { return n ? n->channels : 0; }

// maximum_output of push(): (frame_count * 48000).div_ceil(rate) + 4 (:318-320); frames for 48 kHz sources
extern "C" size_t sk_norm48k_max_output(size_t frames, uint32_t rate) {
    if (!rate) return 0;
    if (rate == 48000u) return frames;
    return (size_t)(((unsigned __int128)frames * 48000u + rate - 1) / rate) + 4;
}

extern "C" int sk_norm48k_push(sk_norm48k* n, const uint8_t* data, size_t bytes, int bits, int fmt, int big_endian, int channels, uint32_t rate,
                               float* left, float* right, size_t cap, size_t* n_out) {
    if (!n) return SK_E_ARG;
    sk_ctx* ctx = n->ctx;
    if (n_out) *n_out = 0;
    if (n->finished) SK_FAIL(ctx, SK_E_STATE, "48 kHz stereo normalizer is already finished");
    if (rate == 0 || channels <= 0) SK_FAIL(ctx, SK_E_ARG, "decoded audio has invalid PCM geometry");
    if (n->rate != 0 && (n->rate != rate || n->channels != channels))
        SK_FAIL(ctx, SK_E_STATE, "decoded PCM geometry changed from %u Hz/%d ch to %u Hz/%d ch", n->rate, n->channels, rate, channels);
    if (bits != 16 && bits != 24 && bits != 32) SK_FAIL(ctx, SK_E_ARG, "unsupported PCM bit depth %d", bits);
    const size_t fb = (size_t)(bits / 8) * (size_t)channels;
    if (bytes == 0 || bytes % fb) SK_FAIL(ctx, SK_E_ARG, "decoded audio contained an incomplete PCM frame");
    if (!data) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    if (n->rate == 0) { n->rate = rate; n->channels = channels; }
    const size_t frames = bytes / fb;
    const unsigned long long start = n->source_frames, end = start + frames;
    unsigned long long count;
    if (rate == 48000u) count = frames;
    else {
        const unsigned long long lim = n48_lerp_limit(end, rate);
        count = lim > n->output_frames ? lim - n->output_frames : 0;
    }
    if (count > cap) SK_FAIL(ctx, SK_E_ARG, "output buffers too small: need %llu frames, have %zu", count, cap);
    if (count && (!left || !right)) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    const size_t plane = sk_round_up((size_t)count, 4) + 4;
    // 48 kHz sources: output index == source frame index
    const unsigned long long o_start = rate == 48000u ? start : n->output_frames;
    float fin[2];
    char* const s_in = sk_stage_half(ctx, 0);
    char* const s_out = sk_stage_half(ctx, 1);
    if (s_in && bytes <= kSkStageHalf && (2 * plane + 4) * sizeof(float) <= kSkStageHalf) {
        // a pushed block is a latency-bound call: the kernel reads the block from the pinned staging buffer and writes both planes and
        // the carried frame straight back into it (zero copy): one launch, one synchronize instead of four
        memcpy(s_in, data, bytes);
        N48Src src;
        if ((rc = n48_fill_src(ctx, &src, s_in, 0, bits, fmt, big_endian, channels, start, frames))) return rc;
        float* d_out = reinterpret_cast<float*>(s_out);
        float* d_final = d_out + 2 * plane;
        if ((rc = n48_launch(ctx, src, o_start, count, count, rate, n->prev_l, n->prev_r, d_out, d_out + plane, 0, 1, d_final))) return rc;
        SK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(fin, d_final, sizeof(fin));
        if (count) {
            memcpy(left, d_out, (size_t)count * sizeof(float));
            memcpy(right, d_out + plane, (size_t)count * sizeof(float));
        }
    } else {
        if ((rc = sk_ensure_dev_in(ctx, bytes))) return rc;
        if ((rc = sk_ensure_dev_out(ctx, (2 * plane + 4) * sizeof(float)))) return rc;
        if ((rc = sk_h2d(ctx, ctx->d_in, data, bytes))) return rc;
        N48Src src;
        if ((rc = n48_fill_src(ctx, &src, ctx->d_in, 0, bits, fmt, big_endian, channels, start, frames))) return rc;
        float* d_out = (float*)ctx->d_out;
        float* d_final = d_out + 2 * plane;
        if ((rc = n48_launch(ctx, src, o_start, count, count, rate, n->prev_l, n->prev_r, d_out, d_out + plane, 0, 1, d_final))) return rc;
        if ((rc = sk_d2h(ctx, fin, d_final, sizeof(fin)))) return rc;
        if (count) {
            if ((rc = sk_d2h(ctx, left, d_out, (size_t)count * sizeof(float)))) return rc;
            if ((rc = sk_d2h(ctx, right, d_out + plane, (size_t)count * sizeof(float)))) return rc;
        }
    }
    n->output_frames += count;
    n->source_frames = end;
    n->prev_l = fin[0]; n->prev_r = fin[1];
    n->has_previous = true;
    if (n_out) *n_out = (size_t)count;
    return SK_OK;
}

// finish(): the tail is `remaining` copies of the retained final sample — state the host already holds, so no launch
extern "C" int sk_norm48k_finish(sk_norm48k* n, float* left, float* right, size_t cap, size_t* n_out) {
    if (!n) return SK_E_ARG;
    sk_ctx* ctx = n->ctx;
    if (n_out) *n_out = 0;
    if (n->finished) SK_FAIL(ctx, SK_E_STATE, "48 kHz stereo normalizer is already finished");
    n->finished = true;
    if (!n->has_previous || n->source_frames == 0 || n->rate == 0) SK_FAIL(ctx, SK_E_STATE, "decoded source contained no PCM");
    if (n->rate == 48000u) return SK_OK;
    const unsigned long long target = n48_target(n->source_frames, n->rate);
    const unsigned long long remaining = target > n->output_frames ? target - n->output_frames : 0;
    n->output_frames = target;
    if (remaining > cap) SK_FAIL(ctx, SK_E_ARG, "output buffers too small: need %llu frames, have %zu", remaining, cap);
    for (unsigned long long i = 0; i < remaining; i++) { left[i] = n->prev_l; right[i] = n->prev_r; }
    if (n_out) *n_out = (size_t)remaining;
    return SK_OK;
}
