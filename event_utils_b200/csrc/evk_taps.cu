// evk_taps.cu -- the reference's lower-level splat / gather helpers as stand-alone entry points:
//   interpolate_to_image            reference lib/representations/image.py:102-115
//   interpolate_to_derivative_img   reference lib/representations/image.py:117-136
//   image_to_event_weights          reference lib/representations/image.py:138-160
//   events_to_image_drv             reference lib/representations/image.py:162-217 (general Jacobians)
// These take arbitrary per-event Jacobians / precomputed indices, so they use plain scalar
// red.global.add.f32 taps; the fused linvel path (evk_cmax.cu) is the fast one.
#include "evk_common.cuh"

namespace evk {

__device__ __forceinline__ bool wrap_i64(long long i, int size, int &out)
{
    if (i < 0) i += size;
    if (i < 0 || i >= size) return false;
    out = (int)i;
    return true;
}

__global__ void __launch_bounds__(256) splat_idx_kernel(const long long *__restrict__ px, const long long *__restrict__ py,
                                                        const float *__restrict__ dx, const float *__restrict__ dy,
                                                        const float *__restrict__ w, int64_t n, int H, int W,
                                                        float *img, unsigned long long *oob_ctr)
{
    unsigned oob = 0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        int x0, x1, y0, y1;
        if (!wrap_i64(px[i], W, x0) || !wrap_i64(px[i] + 1, W, x1) || !wrap_i64(py[i], H, y0) || !wrap_i64(py[i] + 1, H, y1)) { ++oob; continue; }
        const float fx = dx[i], fy = dy[i], ww = w[i];
        const float ox = __fsub_rn(1.0f, fx), oy = __fsub_rn(1.0f, fy);
        const float wl = __fmul_rn(ww, ox), wr = __fmul_rn(ww, fx);
        red_add(img + (int64_t)y0 * W + x0, __fmul_rn(wl, oy));
        red_add(img + (int64_t)y0 * W + x1, __fmul_rn(wr, oy));
        red_add(img + (int64_t)y1 * W + x0, __fmul_rn(wl, fy));
        red_add(img + (int64_t)y1 * W + x1, __fmul_rn(wr, fy));
    }
    flush_oob(oob_ctr, oob);
}

// image.py:131-135 tap weights for one derivative plane
__device__ __forceinline__ void drv_taps(float *plane, int W, int x0, int x1, int y0, int y1, float a1, float a2,
                                         float fx, float fy, float ox, float oy)
{
    red_add(plane + (int64_t)y0 * W + x0, __fadd_rn(__fmul_rn(a1, -oy), __fmul_rn(a2, -ox)));
    red_add(plane + (int64_t)y0 * W + x1, __fadd_rn(__fmul_rn(a1, oy), __fmul_rn(a2, -fx)));
    red_add(plane + (int64_t)y1 * W + x0, __fadd_rn(__fmul_rn(a1, -fy), __fmul_rn(a2, ox)));
    red_add(plane + (int64_t)y1 * W + x1, __fadd_rn(__fmul_rn(a1, fy), __fmul_rn(a2, fx)));
}

__global__ void __launch_bounds__(256) splat_drv_idx_kernel(const long long *__restrict__ px, const long long *__restrict__ py,
                                                            const float *__restrict__ dx, const float *__restrict__ dy,
                                                            const float *__restrict__ w1, const float *__restrict__ w2,
                                                            int K, int64_t n, int H, int W, float *dimg,
                                                            unsigned long long *oob_ctr)
{
    unsigned oob = 0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        int x0, x1, y0, y1;
        if (!wrap_i64(px[i], W, x0) || !wrap_i64(px[i] + 1, W, x1) || !wrap_i64(py[i], H, y0) || !wrap_i64(py[i] + 1, H, y1)) { ++oob; continue; }
        const float fx = dx[i], fy = dy[i];
        const float ox = __fsub_rn(1.0f, fx), oy = __fsub_rn(1.0f, fy);
        for (int k = 0; k < K; ++k)
            drv_taps(dimg + (int64_t)k * H * W, W, x0, x1, y0, y1, w1[(int64_t)k * n + i], w2[(int64_t)k * n + i], fx, fy, ox, oy);
    }
    flush_oob(oob_ctr, oob);
}

__global__ void __launch_bounds__(256) image_drv_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                        const float *__restrict__ p, const float *__restrict__ jx,
                                                        const float *__restrict__ jy, int K, int64_t n, int H, int W,
                                                        int clip, float clipx, float clipy, float *img, float *dimg,
                                                        unsigned long long *oob_ctr)
{
    unsigned oob = 0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const float xe = x[i], ye = y[i];
        float m = 1.0f;
        if (clip) m = (xe >= clipx ? 0.0f : 1.0f) * (ye >= clipy ? 0.0f : 1.0f);
        const float pxf = floorf(xe), pyf = floorf(ye);
        const float fx = __fsub_rn(xe, pxf), fy = __fsub_rn(ye, pyf);
        int upx, upy, x0, x1, y0, y1;
        if (!trunc_checked(__fmul_rn(pxf, m), upx) || !trunc_checked(__fmul_rn(pyf, m), upy) ||
            !wrap_int_index(upx, W, x0) || !wrap_int_index(upx + 1, W, x1) ||
            !wrap_int_index(upy, H, y0) || !wrap_int_index(upy + 1, H, y1)) { ++oob; continue; }
        const float wp = __fmul_rn(p[i], m);
        const float ox = __fsub_rn(1.0f, fx), oy = __fsub_rn(1.0f, fy);
        const float wl = __fmul_rn(wp, ox), wr = __fmul_rn(wp, fx);
        red_add(img + (int64_t)y0 * W + x0, __fmul_rn(wl, oy));
        red_add(img + (int64_t)y0 * W + x1, __fmul_rn(wr, oy));
        red_add(img + (int64_t)y1 * W + x0, __fmul_rn(wl, fy));
        red_add(img + (int64_t)y1 * W + x1, __fmul_rn(wr, fy));
        for (int k = 0; k < K; ++k)
            drv_taps(dimg + (int64_t)k * H * W, W, x0, x1, y0, y1, __fmul_rn(jx[(int64_t)k * n + i], wp),
                     __fmul_rn(jy[(int64_t)k * n + i], wp), fx, fy, ox, oy);
    }
    flush_oob(oob_ctr, oob);
}

__global__ void __launch_bounds__(256) gather_bilinear_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                                              int64_t n, const double *__restrict__ img, int H, int W,
                                                              double *__restrict__ out, unsigned long long *oob_ctr)
{
    unsigned oob = 0;
    const double clipx = (double)(W - 1), clipy = (double)(H - 1);
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const double xe = x[i], ye = y[i];
        const double m = (xe >= clipx ? 0.0 : 1.0) * (ye >= clipy ? 0.0 : 1.0);
        const double pxd = floor(__dmul_rn(xe, m)), pyd = floor(__dmul_rn(ye, m));
        int x0, x1, y0, y1;
        if (!(fabs(pxd) < 2.0e9) || !(fabs(pyd) < 2.0e9) ||
            !wrap_i64((long long)pxd, W, x0) || !wrap_i64((long long)pxd + 1, W, x1) ||
            !wrap_i64((long long)pyd, H, y0) || !wrap_i64((long long)pyd + 1, H, y1)) { ++oob; out[i] = 0.0; continue; }
        const double fx = __dsub_rn(xe, pxd), fy = __dsub_rn(ye, pyd);
        const double ox = __dsub_rn(1.0, fx), oy = __dsub_rn(1.0, fy);
        double wgt = __dmul_rn(__dmul_rn(img[(int64_t)y0 * W + x0], ox), oy);
        wgt = __dadd_rn(wgt, __dmul_rn(__dmul_rn(img[(int64_t)y0 * W + x1], fx), oy));
        wgt = __dadd_rn(wgt, __dmul_rn(__dmul_rn(img[(int64_t)y1 * W + x0], ox), fy));
        wgt = __dadd_rn(wgt, __dmul_rn(__dmul_rn(img[(int64_t)y1 * W + x1], fx), fy));
        out[i] = __dmul_rn(wgt, m);
    }
    flush_oob(oob_ctr, oob);
}

}  // namespace evk

extern "C" {

int evk_splat_idx_f32(const int64_t *px, const int64_t *py, const float *dx, const float *dy, const float *w,
                      int64_t n, int H, int W, float *img, unsigned long long *oob, void *stream)
{
    using namespace evk;
    if (n < 0 || H < 2 || W < 2 || !img || (n > 0 && (!px || !py || !dx || !dy || !w))) { set_error("evk_splat_idx_f32: bad arguments"); return EVK_E_ARG; }
    if (n == 0) return EVK_OK;
    splat_idx_kernel<<<grid_for(splat_idx_kernel, 256, n, 256 * 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        (const long long *)px, (const long long *)py, dx, dy, w, n, H, W, img, oob);
    EVK_CUDA(cudaGetLastError());
    return EVK_OK;
}

int evk_splat_drv_idx_f32(const int64_t *px, const int64_t *py, const float *dx, const float *dy, const float *w1,
                          const float *w2, int K, int64_t n, int H, int W, float *dimg, unsigned long long *oob,
                          void *stream)
{
    using namespace evk;
    if (n < 0 || K < 1 || H < 2 || W < 2 || !dimg || (n > 0 && (!px || !py || !dx || !dy || !w1 || !w2))) { set_error("evk_splat_drv_idx_f32: bad arguments"); return EVK_E_ARG; }
    if (n == 0) return EVK_OK;
    splat_drv_idx_kernel<<<grid_for(splat_drv_idx_kernel, 256, n, 256 * 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        (const long long *)px, (const long long *)py, dx, dy, w1, w2, K, n, H, W, dimg, oob);
    EVK_CUDA(cudaGetLastError());
    return EVK_OK;
}

int evk_image_drv_f32(const float *x, const float *y, const float *p, const float *jx, const float *jy, int K,
                      int64_t n, int Himg, int Wimg, float clipx, float clipy, unsigned flags, float *img,
                      float *dimg, unsigned long long *oob, void *stream)
{
    using namespace evk;
    if (n < 0 || K < 0 || Himg < 2 || Wimg < 2 || !img || (K > 0 && (!dimg || !jx || !jy)) || (n > 0 && (!x || !y || !p))) {
        set_error("evk_image_drv_f32: bad arguments");
        return EVK_E_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t plane = (size_t)Himg * Wimg * sizeof(float);
    if (!(flags & EVK_ACCUMULATE)) {
        EVK_CUDA(cudaMemsetAsync(img, 0, plane, st));
        if (K > 0) EVK_CUDA(cudaMemsetAsync(dimg, 0, plane * K, st));
    }
    if (n > 0) {
        image_drv_kernel<<<grid_for(image_drv_kernel, 256, n, 256 * 4), 256, 0, st>>>(x, y, p, jx, jy, K, n, Himg, Wimg,
                                                                  (flags & EVK_CLIP) ? 1 : 0, clipx, clipy, img, dimg, oob);
        EVK_CUDA(cudaGetLastError());
    }
    return EVK_OK;
}

int evk_gather_bilinear_f64(const double *x, const double *y, int64_t n, const double *img, int H, int W, double *out,
                            unsigned long long *oob, void *stream)
{
    using namespace evk;
    if (n < 0 || H < 2 || W < 2 || !img || (n > 0 && (!x || !y || !out))) { set_error("evk_gather_bilinear_f64: bad arguments"); return EVK_E_ARG; }
    if (n == 0) return EVK_OK;
    gather_bilinear_kernel<<<grid_for(gather_bilinear_kernel, 256, n, 256 * 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, img, H, W, out, oob);
    EVK_CUDA(cudaGetLastError());
    return EVK_OK;
}

}  // extern "C"
