// evk_hot.cu -- hot-spot (contention-aware) event-image scatter: a per-CTA write-combining cache
// in shared memory in front of the global reductions.
//
// Same semantics as evk_image.cu (events_to_image_torch, reference
// lib/representations/image.py:46-115); only the accumulation strategy differs.
//
// Why: on a Zipf-distributed stream (BASELINE config 4) a few pixels receive millions of events.
// Global reductions to ONE address serialise in its L2 slice (measured: 5.5 ms for 50 M events at
// s=1.0, 14.7 ms at s=1.2, against 0.33 ms for a uniform stream).  Each CTA therefore keeps a
// small direct-mapped table {cell -> partial sum} in shared memory: an event whose cell owns its
// slot is accumulated with a shared-memory atomic, everything else (slot taken by another cell)
// goes straight to L2, and the table is flushed with one reduction per occupied slot at the end.
// Hot cells are thereby reduced to one L2 reduction per CTA.  The kernel is ADAPTIVE: a prologue
// counts intra-warp duplicates (match.any) among the CTA's first events and turns the cache off
// for streams without contention, where it would only add shared-memory traffic.
#include "evk_common.cuh"

namespace evk {

struct HotArgs {
    const float *x, *y, *p;
    int64_t n;
    int H, W;
    int clip;
    float clipx, clipy;
    float *out;
    float *ws;         // bilinear only: block workspace ws[H][W][4] (TL,TR,BL,BR); NULL -> taps go to `out`
    unsigned *out_u32;
    unsigned long long *oob;
    int force_cache;  // 0 adaptive, 1 always on
    int vec4;         // x, y (and p) are 16-byte aligned
};

constexpr int kHotLog2 = 12;
constexpr int kHotSlots = 1 << kHotLog2;  // 4096 slots: 16 KB keys + 16 KB values
constexpr unsigned kEmpty = 0xffffffffu;

// `gs` = element stride of a cell in the global target (4 when the target is the TL slot of a block)
__device__ __forceinline__ void global_add(float *out, unsigned cell, float v, int gs) { red_add(out + (size_t)cell * gs, v); }
__device__ __forceinline__ void global_add(unsigned *out, unsigned cell, unsigned v, int gs) { red_add_u32(out + (size_t)cell * gs, v); }

template <typename V>
__device__ __forceinline__ void hot_add(unsigned *keys, V *vals, V *gout, bool use_cache, unsigned cell, V v, int gs = 1)
{
    if (use_cache) {
        const unsigned slot = (cell * 2654435761u) >> (32 - kHotLog2);
        unsigned k = keys[slot];
        if (k == kEmpty) {
            const unsigned old = atomicCAS(&keys[slot], kEmpty, cell);
            k = (old == kEmpty) ? cell : old;
        }
        if (k == cell) {
            atomicAdd(&vals[slot], v);  // shared memory: native for u32, CAS loop (ATOMS.CAST.SPIN) for f32
            return;
        }
    }
    global_add(gout, cell, v, gs);
}

enum { HOT_NEAREST = 0, HOT_BILINEAR = 1, HOT_COUNT = 2 };

// one event; CACHE is a compile-time switch so that the cache-off instantiation is the plain scatter code
template <int MODE, bool CACHE, typename V>
__device__ __forceinline__ void hot_event(const HotArgs &A, unsigned *keys, V *vals, V *gout, int gs, float x, float y, float pin,
                                          unsigned &oob)
{
    constexpr bool use_cache = CACHE;
    if (MODE == HOT_NEAREST || MODE == HOT_COUNT) {
        // image.py:88-95
        const bool keep = !A.clip || (!(x >= A.clipx) && !(y >= A.clipy));
        int ux, uy, xi, yi;
        if (!trunc_checked(x, ux) || !trunc_checked(y, uy)) { ++oob; return; }
        if (!keep) { ux = 0; uy = 0; }
        if (!wrap_int_index(ux, A.W, xi) || !wrap_int_index(uy, A.H, yi)) { ++oob; return; }
        const unsigned cell = (unsigned)yi * (unsigned)A.W + (unsigned)xi;
        if (MODE == HOT_COUNT) {
            hot_add<V>(keys, vals, gout, use_cache, cell, (V)1);
        } else {
            const float p = pin;
            if (p != 0.0f) hot_add<V>(keys, vals, gout, use_cache, cell, (V)p);
        }
    } else {
        // image.py:79-86 + 111-114
        const float p = pin;
        float m = 1.0f;
        if (A.clip) m = (x >= A.clipx ? 0.0f : 1.0f) * (y >= A.clipy ? 0.0f : 1.0f);
        const float pxf = floorf(x), pyf = floorf(y);
        const float dx = __fsub_rn(x, pxf), dy = __fsub_rn(y, pyf);
        int upx, upy, x0, x1, y0, y1;
        if (!trunc_checked(__fmul_rn(pxf, m), upx) || !trunc_checked(__fmul_rn(pyf, m), upy) ||
            !wrap_int_index(upx, A.W, x0) || !wrap_int_index(upx + 1, A.W, x1) ||
            !wrap_int_index(upy, A.H, y0) || !wrap_int_index(upy + 1, A.H, y1)) { ++oob; return; }
        const float w = __fmul_rn(p, m);
        const float ox = __fsub_rn(1.0f, dx), oy = __fsub_rn(1.0f, dy);
        const float wl = __fmul_rn(w, ox), wr = __fmul_rn(w, dx);
        const float v00 = __fmul_rn(wl, oy), v01 = __fmul_rn(wr, oy), v10 = __fmul_rn(wl, dy), v11 = __fmul_rn(wr, dy);
        const unsigned r0 = (unsigned)y0 * (unsigned)A.W, r1 = (unsigned)y1 * (unsigned)A.W;
        if (!use_cache && A.ws && x1 == x0 + 1 && y1 == y0 + 1) {
            // uncontended stream: the whole footprint as ONE vector reduction into its block
            if (v00 != 0.0f || v01 != 0.0f || v10 != 0.0f || v11 != 0.0f)
                red_add4(A.ws + ((size_t)r0 + x0) * 4, make_float4(v00, v01, v10, v11));
        } else {
            if (v00 != 0.0f) hot_add<V>(keys, vals, gout, use_cache, r0 + x0, (V)v00, gs);
            if (v01 != 0.0f) hot_add<V>(keys, vals, gout, use_cache, r0 + x1, (V)v01, gs);
            if (v10 != 0.0f) hot_add<V>(keys, vals, gout, use_cache, r1 + x0, (V)v10, gs);
            if (v11 != 0.0f) hot_add<V>(keys, vals, gout, use_cache, r1 + x1, (V)v11, gs);
        }
    }
}

// the event loop; VEC4: 16-byte loads (x, y, p all 16-byte aligned)
template <int MODE, bool CACHE, bool VEC4, typename V>
__device__ __forceinline__ void hot_loop(const HotArgs &A, unsigned *keys, V *vals, V *gout, int gs, int64_t tid, int64_t stride,
                                         unsigned &oob)
{
    if (VEC4) {
        const int64_t n4 = A.n >> 2;
        for (int64_t g = tid; g < n4; g += stride) {
            const float4 X = ld_stream4(A.x + 4 * g), Y = ld_stream4(A.y + 4 * g);
            const float4 P = (MODE == HOT_COUNT) ? make_float4(1.f, 1.f, 1.f, 1.f) : ld_stream4(A.p + 4 * g);
            hot_event<MODE, CACHE, V>(A, keys, vals, gout, gs, X.x, Y.x, P.x, oob);
            hot_event<MODE, CACHE, V>(A, keys, vals, gout, gs, X.y, Y.y, P.y, oob);
            hot_event<MODE, CACHE, V>(A, keys, vals, gout, gs, X.z, Y.z, P.z, oob);
            hot_event<MODE, CACHE, V>(A, keys, vals, gout, gs, X.w, Y.w, P.w, oob);
        }
        for (int64_t i = (n4 << 2) + tid; i < A.n; i += stride)
            hot_event<MODE, CACHE, V>(A, keys, vals, gout, gs, A.x[i], A.y[i], (MODE == HOT_COUNT) ? 1.0f : A.p[i], oob);
    } else {
        for (int64_t i = tid; i < A.n; i += stride)
            hot_event<MODE, CACHE, V>(A, keys, vals, gout, gs, ld_stream(A.x + i), ld_stream(A.y + i),
                                      (MODE == HOT_COUNT) ? 1.0f : ld_stream(A.p + i), oob);
    }
}


template <int MODE>
__global__ void __launch_bounds__(256) image_hot_kernel(const HotArgs A)
{
    using V = typename std::conditional<MODE == HOT_COUNT, unsigned, float>::type;
    __shared__ unsigned keys[kHotSlots];
    __shared__ V vals[kHotSlots];
    V *gout = (MODE == HOT_COUNT) ? (V *)A.out_u32 : (MODE == HOT_BILINEAR && A.ws) ? (V *)A.ws : (V *)A.out;
    const int gs = (MODE == HOT_BILINEAR && A.ws) ? 4 : 1;

    const int64_t tid = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * 256;

    // ---- prologue: is this stream contended?  (lanes whose first event shares its pixel with
    // another lane of the same warp)
    bool use_cache = A.force_cache != 0;
    if (!use_cache) {
        unsigned long long key = ~0ull - (threadIdx.x & 31);
        if (tid < A.n) {
            int ux, uy;
            if (trunc_checked(A.x[tid], ux) && trunc_checked(A.y[tid], uy)) key = ((unsigned long long)(unsigned)uy << 32) | (unsigned)ux;
        }
        // counted over the CTA by the barrier itself: one __syncthreads_count, no shared-memory traffic
        const int dup_lanes = __syncthreads_count(__popc(__match_any_sync(0xffffffffu, key)) > 1);
        use_cache = dup_lanes * 64 > 256;  // > 4 of 256 lanes collide inside their warp (uniform streams: ~0.01)
    }
    if (use_cache) {   // only CTAs that will use the table pay for initialising it
        for (int s = threadIdx.x; s < kHotSlots; s += 256) { keys[s] = kEmpty; vals[s] = (V)0; }
        __syncthreads();
    }

    unsigned oob = 0;
    // cache off -> the plain scatter code (vector loads when possible), cache on -> the cached code
    if (use_cache) hot_loop<MODE, true, false, V>(A, keys, vals, gout, gs, tid, stride, oob);
    else if (A.vec4) hot_loop<MODE, false, true, V>(A, keys, vals, gout, gs, tid, stride, oob);
    else hot_loop<MODE, false, false, V>(A, keys, vals, gout, gs, tid, stride, oob);
    __syncthreads();
    if (use_cache) {
        for (int s = threadIdx.x; s < kHotSlots; s += 256) {
            const unsigned k = keys[s];
            if (k != kEmpty && vals[s] != (V)0) global_add(gout, k, vals[s], gs);
        }
    }
    flush_oob(A.oob, oob);
}

int launch_image_hot(const float *x, const float *y, const float *p, int64_t n, int H, int W, int clip, float clipx,
                     float clipy, int mode, int force_cache, float *out, float *ws, unsigned *out_u32,
                     unsigned long long *oob, cudaStream_t st)
{
    if ((int64_t)H * W >= 0xffffffffLL) { set_error("hot-spot variant: image too large for 32-bit cell ids"); return EVK_E_UNSUPPORTED; }
    HotArgs A{};
    A.x = x; A.y = y; A.p = p; A.n = n; A.H = H; A.W = W;
    A.clip = clip; A.clipx = clipx; A.clipy = clipy;
    A.out = out; A.ws = ws; A.out_u32 = out_u32; A.oob = oob; A.force_cache = force_cache;
    A.vec4 = ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)(p ? p : x)) & 15) == 0) ? 1 : 0;
    if (n <= 0) return EVK_OK;
    ProfScope prof(st);
    prof_count(1);
    if (mode == HOT_NEAREST) image_hot_kernel<HOT_NEAREST><<<grid_for(image_hot_kernel<HOT_NEAREST>, 256, n, 256 * 16), 256, 0, st>>>(A);
    else if (mode == HOT_BILINEAR) image_hot_kernel<HOT_BILINEAR><<<grid_for(image_hot_kernel<HOT_BILINEAR>, 256, n, 256 * 16), 256, 0, st>>>(A);
    else image_hot_kernel<HOT_COUNT><<<grid_for(image_hot_kernel<HOT_COUNT>, 256, n, 256 * 16), 256, 0, st>>>(A);
    EVK_CUDA(cudaGetLastError());
    return EVK_OK;
}

}  // namespace evk
