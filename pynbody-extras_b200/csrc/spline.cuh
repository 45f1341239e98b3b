// spline.cuh — fp32 cubic-spline (Springel W2) pair terms shared by the direct and tree kernels.
#pragma once

namespace pnbx {

// Branch-free W2 terms of one pair with r < h (kernel.rs:84-124): u = r/h, 1/u = h/r, both polynomial branches are
// evaluated and selected (no divergence, no IEEE division). kpot = W2(u)/h, kacc = W2'(u)/(h^2 r).
// r2t = r^2 (+ R2_TINY), rinv = rsqrt(r2t), hinv = 1/h. The unselected branch may overflow for r -> 0; it is
// discarded by the select.
__device__ __forceinline__ void w2_terms_f32(float r2t, float rinv, float h, float hinv, float& kpot, float& kacc) {
    const float u = r2t * rinv * hinv, uinv = h * rinv, u2 = u * u;
    const bool lo = u < 0.5f;
    const float wi = fmaf(u2, fmaf(u2, fmaf(6.4f, u, -9.6f), 16.0f / 3.0f), -2.8f);
    const float wo = fmaf(u2, fmaf(u, fmaf(u, fmaf(-32.0f / 15.0f, u, 9.6f), -16.0f), 32.0f / 3.0f),
                          fmaf(1.0f / 15.0f, uinv, -3.2f));
    kpot = (lo ? wi : wo) * hinv;
    const float pi = u * fmaf(u2, fmaf(32.0f, u, -38.4f), 32.0f / 3.0f);
    const float po = fmaf(u, fmaf(u, fmaf(u, fmaf(-32.0f / 3.0f, u, 38.4f), -48.0f), 64.0f / 3.0f),
                          (-1.0f / 15.0f) * uinv * uinv);
    kacc = (lo ? pi : po) * (hinv * hinv) * rinv;
}

}  // namespace pnbx
