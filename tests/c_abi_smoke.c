/* c_abi_smoke.c - a language-neutral caller of the device-pointer ABI (include/dpr.h): plain C, the CUDA runtime, raw
 * cudaMalloc pointers, the 15- and 20-argument calls marshalled exactly as a Julia `ccall` would marshal them
 * (INTEGRATION.md 2).  Stands in for the Julia glue, which cannot run in this image (no Julia).
 *
 * Case: the reference's first known-answer test (src/raster.jl:143-157): one 2-d point at the origin, identity
 * rotation, zero translation, 5 x 5 grid -> a single 1 at the centre cell.  Then the pullback of ds_dout = 1, 2, .. 25
 * (column-major): d_background = 325, d_out_weight = ds_dout[centre] = 13.
 * Build + run: see tests/test_abi.py::test_c_abi_smoke.  Exit code 0 = ok. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dpr.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); return 2; } } while (0)
#define CHECK_DPR(x) do { int rc_ = (x); if (rc_ != DPR_OK) { printf("libdpr status %d (%s) [%s] at %s\n", rc_, dpr_status_string(rc_), dpr_last_error_message(), #x); return 3; } } while (0)

int main(void) {
    const int64_t grid[2] = {5, 5};
    const int64_t P = 1, B = 1;
    const float h_points[2] = {0.f, 0.f}, h_rot[4] = {1.f, 0.f, 0.f, 1.f}, h_tr[2] = {0.f, 0.f};
    float h_ds[25], h_out[25];
    for (int i = 0; i < 25; ++i) h_ds[i] = (float)(i + 1);

    float *points, *rot, *tr, *out, *ds, *d_points, *d_rot, *d_tr, *d_bg, *d_ow, *d_pw;
    CHECK_CUDA(cudaMalloc((void**)&points, sizeof h_points));
    CHECK_CUDA(cudaMalloc((void**)&rot, sizeof h_rot));
    CHECK_CUDA(cudaMalloc((void**)&tr, sizeof h_tr));
    CHECK_CUDA(cudaMalloc((void**)&out, sizeof h_out));
    CHECK_CUDA(cudaMalloc((void**)&ds, sizeof h_ds));
    CHECK_CUDA(cudaMalloc((void**)&d_points, 2 * sizeof(float)));
    CHECK_CUDA(cudaMalloc((void**)&d_rot, 4 * sizeof(float)));
    CHECK_CUDA(cudaMalloc((void**)&d_tr, 2 * sizeof(float)));
    CHECK_CUDA(cudaMalloc((void**)&d_bg, sizeof(float)));
    CHECK_CUDA(cudaMalloc((void**)&d_ow, sizeof(float)));
    CHECK_CUDA(cudaMalloc((void**)&d_pw, sizeof(float)));
    CHECK_CUDA(cudaMemcpy(points, h_points, sizeof h_points, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(rot, h_rot, sizeof h_rot, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(tr, h_tr, sizeof h_tr, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(ds, h_ds, sizeof h_ds, cudaMemcpyHostToDevice));

    /* the caller owns the workspace: size from the library, first 256 bytes zeroed (binning-cache contract) */
    size_t wf = dpr_workspace_bytes(DPR_OP_FORWARD, 2, 2, grid, P, B, 4), wb = dpr_workspace_bytes(DPR_OP_PULLBACK, 2, 2, grid, P, B, 4);
    size_t wbytes = wf > wb ? wf : wb;
    if (wbytes < 256) wbytes = 256;
    void* ws;
    CHECK_CUDA(cudaMalloc(&ws, wbytes));
    CHECK_CUDA(cudaMemset(ws, 0, 256));
    cudaStream_t stream;
    CHECK_CUDA(cudaStreamCreate(&stream));

    if (dpr_version() < 100) { printf("bad version\n"); return 4; }
    /* forward: NULL background / weights = the FillArrays defaults */
    CHECK_DPR(dpr_raster_forward_f32(2, 2, grid, P, B, points, rot, tr, NULL, NULL, NULL, out, ws, wbytes, (dpr_stream_t)stream));
    CHECK_CUDA(cudaMemcpyAsync(h_out, out, sizeof h_out, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < 25; ++i) {
        const float want = (i == 12) ? 1.f : 0.f;
        if (fabsf(h_out[i] - want) > 1e-6f) { printf("forward: out[%d] = %g, want %g\n", i, h_out[i], want); return 5; }
    }
    CHECK_DPR(dpr_raster_pullback_f32(2, 2, grid, P, B, ds, points, rot, tr, NULL, NULL, d_points, d_rot, d_tr, d_bg, d_ow, d_pw, ws, wbytes,
                                      (dpr_stream_t)stream));
    float g_bg, g_ow, g_pw, g_tr[2], g_pts[2];
    CHECK_CUDA(cudaMemcpyAsync(&g_bg, d_bg, sizeof g_bg, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaMemcpyAsync(&g_ow, d_ow, sizeof g_ow, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaMemcpyAsync(&g_pw, d_pw, sizeof g_pw, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaMemcpyAsync(g_tr, d_tr, sizeof g_tr, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaMemcpyAsync(g_pts, d_points, sizeof g_pts, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaStreamSynchronize(stream));
    /* the point sits on a cell centre: dl = 1, so all the weight is on the upper corner, cell (2, 2) = 13; the gradient
     * of the coordinate is (G(2,2) - G(1,2), G(2,2) - G(2,1)) * scale = (13 - 12, 13 - 8) * 2.5 */
    if (fabsf(g_bg - 325.f) > 1e-3f || fabsf(g_ow - 13.f) > 1e-5f || fabsf(g_pw - 13.f) > 1e-5f) {
        printf("pullback: d_background %g (325) d_out_weight %g (13) d_point_weight %g (13)\n", g_bg, g_ow, g_pw);
        return 6;
    }
    if (fabsf(g_tr[0] - 2.5f) > 1e-5f || fabsf(g_tr[1] - 12.5f) > 1e-5f || fabsf(g_pts[0] - 2.5f) > 1e-5f || fabsf(g_pts[1] - 12.5f) > 1e-5f) {
        printf("pullback: d_translation (%g, %g) d_points (%g, %g), want (2.5, 12.5)\n", g_tr[0], g_tr[1], g_pts[0], g_pts[1]);
        return 7;
    }
    /* error convention: a status code, never an abort */
    if (dpr_raster_forward_f32(7, 2, grid, P, B, points, rot, tr, NULL, NULL, NULL, out, ws, wbytes, (dpr_stream_t)stream) != DPR_ERR_UNSUPPORTED) return 8;
    if (dpr_raster_forward_f32(2, 2, grid, P, B, NULL, rot, tr, NULL, NULL, NULL, out, ws, wbytes, (dpr_stream_t)stream) != DPR_ERR_NULL_POINTER) return 9;
    printf("c_abi_smoke ok: forward %s, pullback %s, %lld kernel launches\n", dpr_last_path(DPR_OP_FORWARD), dpr_last_path(DPR_OP_PULLBACK),
           (long long)dpr_kernel_launch_count());
    return 0;
}
